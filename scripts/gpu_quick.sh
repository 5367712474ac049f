#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/bench_latency.py > gpurun_out/bench_latency.jsonl 2> gpurun_out/bench_latency.err; echo "rc=$?"; cat gpurun_out/bench_latency.jsonl; tail -5 gpurun_out/bench_latency.err
