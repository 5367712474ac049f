#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/ab_step.py red=caco_set_gemm_resid_red:1 nored=caco_set_gemm_resid_red:0 e8=caco_set_gemm_variant:3 auto=caco_set_gemm_variant:0 2>&1 | tail -3
