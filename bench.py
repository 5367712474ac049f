#!/usr/bin/env python
"""bench.py — headline benchmark of the Cacophony inference hot path on B200 (driver contract: see the task prompt).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input: B = 256 audio-text pairs per GPU
(10 s @ 16 kHz clips + 32-token captions): waveform -> STFT/mel/patch frontend -> AudioMAE-ViT tower -> pooler,
ids -> RoBERTa tower -> pooler/projection, L2-normalise, (N > 1: one NCCL all-gather of the embeddings), and the
exp(logit_scale)-scaled cosine-similarity matrix — BASELINE.json configs[2] (configs[3] sharded when N > 1).

  value        pairs/s with the inputs already resident in HBM (device-timed, max over ranks, whole job)
  e2e          the same metric through the public API with HOST (pinned) inputs: per step H2D of waveforms/ids/mask and
               D2H of the logits block are inside the timed region (double-buffered so copies overlap compute)
  roofline     the dominant kernel family (tcgen05 GEMM): algorithmic FLOPs / CUDA-event time measured live in the timed
               region, against the MEASURED sustained bf16 GEMM peak (MEASURED_PEAKS.json)
  cpu_baseline the CPU port of the reference algorithm (oracle/) timed on this box's host cores on a bounded sample
  --impl reference   times that CPU implementation alone (the reference arm for this tier)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_SAMPLES = 160000          # 10 s @ 16 kHz
MAX_PATCHES = 500           # eval_caco_torch.py:573  (100*10*8//16)
TEXT_LEN = 32
METRIC = "audio-text pairs/sec (10s@16kHz)"
UNIT = "pairs/s"
GFLOP_PER_PAIR = 101.08     # SURVEY.md §8d (audio 95.53 + text 5.55)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_sustained": d.get("bf16_tflops_sustained"), "bf16_burst": d.get("bf16_tflops"),
                "hbm": d.get("hbm_gbs"), "src": "measured"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "src": "fallback"}


def gemm_traffic_from_profile():
    """DRAM bytes per GEMM launch from the committed `ncu --set full` capture (profiles/r01_ncu_full_summary.csv: the four
    GEMMs of one audio layer — QKV, out-proj, fc1, fc2 — dram__bytes_read.sum + dram__bytes_write.sum), averaged per
    launch.  Static evidence read from the repo, not measured in this run (ncu cannot run inside a timed bench)."""
    import csv
    p = os.path.join(ROOT, "profiles", "r01_ncu_full_summary.csv")
    if not os.path.exists(p):
        return None, None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, n = 0.0, 0
    with open(p, newline="") as f:
        for row in csv.DictReader(f):
            if "gemm_f16_kernel" not in row["kernel"]:
                continue
            for k, v in row.items():
                if k.startswith(("dram__bytes_read.sum", "dram__bytes_write.sum")) and v:
                    tot += float(v) * unit.get(k.split("[")[1].rstrip("]"), 1.0)
            n += 1
    return (tot / n, n) if n else (None, None)


def workload_config(batch: int, world: int):
    return {"workload": f"{batch} audio-text pairs per GPU: 10 s @ 16 kHz clips (S=500 patches) + {TEXT_LEN}-token captions, "
                        "frontend + AudioMAE-ViT + RoBERTa + cosine-sim" + (" + NCCL all-gather" if world > 1 else ""),
            "global_batch": batch * world, "parallelism": f"dp{world}", "random_init_weights": True,
            "l2": "inputs+activations per step (2.4 GB) exceed the 126 MB L2; no explicit flush"}


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region (pynvml; B200_PROFILING.md 'clocks line')."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                     0x80: "hw_power_brake_slowdown"}
            while not self._halt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, n in names.items():
                    if r & bit:
                        self.reasons.add(n)
                time.sleep(0.02)
        except Exception as e:          # never fail the bench because of telemetry
            self.reasons.add(f"unavailable:{type(e).__name__}")

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def synth_inputs(batch: int, seed: int):
    """Synthetic 10 s clips (uniform ±0.1 white noise) and 32-token captions (<s>=0 ... </s>=2, body in [3, vocab))."""
    import torch
    g = torch.Generator().manual_seed(1234 + seed)
    wave = (0.1 * (2.0 * torch.rand(batch, N_SAMPLES, generator=g) - 1.0)).float()
    ids = torch.randint(3, 50265, (batch, TEXT_LEN), generator=g, dtype=torch.int64)
    ids[:, 0] = 0
    ids[:, -1] = 2
    mask = torch.ones(batch, TEXT_LEN, dtype=torch.float32)
    return wave, ids, mask


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on the host cores (checker code used as a timed baseline)
# ----------------------------------------------------------------------------------------------------------------
def cpu_pairs_per_s(sd_cpu, sample_pairs: int, steps: int, warmup: int, budget_s: float = 0.0):
    """Times `steps` passes of `sample_pairs` pairs (after `warmup` untimed passes) through the CPU port of the reference
    algorithm with every host thread.  budget_s > 0: `steps` is a minimum and passes continue until about budget_s seconds
    of timed CPU work have been done (bench.py's cpu_baseline leg: ~10-30 s).  Returns (pairs/s, s per pass, passes)."""
    import torch
    from oracle import caco_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    wave, ids, mask = synth_inputs(sample_pairs, 99)
    times = []
    with torch.no_grad():
        i = 0
        while True:
            t0 = time.perf_counter()
            ab = O.prepare_audio_batch(list(wave.numpy()), MAX_PATCHES)
            O.forward(sd_cpu, ab["audio_patches"], ab["audio_time_inds"], ab["audio_freq_inds"], ab["audio_mask"], ids, mask)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
            i += 1
            if len(times) >= steps and (budget_s <= 0 or sum(times) >= budget_s or len(times) >= 40):
                break
    dt = sum(times) / len(times)
    return sample_pairs / dt, dt, len(times)


def cpu_model():
    try:
        import cpuinfo
        return cpuinfo.get_cpu_info().get("brand_raw", "unknown")
    except Exception:
        return "unknown"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    import cacophony_b200 as cb
    torch.manual_seed(0)
    sd = {k: v for k, v in cb.create_caco_model().state_dict().items()}
    sample = 16
    value, dt, _ = cpu_pairs_per_s(sd, sample, max(1, args.steps), max(1, min(args.warmup, 1)))
    cores = os.cpu_count() or 1
    desc = f"{sample} pairs per step (same clip/caption shape as the GPU arm), fp32 torch CPU ops, {cores} threads"
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus,
            "steps": max(1, args.steps), "warmup": max(1, min(args.warmup, 1)), "ms_per_step": round(dt * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args.batch, max(1, args.gpus)), sample=desc),
            "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                             "cpu": cpu_model()},
            "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import cacophony_b200 as cb
    from cacophony_b200 import _lib as L
    from cacophony_b200 import dist as cdist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = L.load()
    B = args.batch

    torch.manual_seed(0)                                   # same random-init weights on every rank
    model = cb.create_caco_model()
    sd_cpu = {k: v.clone() for k, v in model.state_dict().items()} if (rank == 0 and world == 1 and not args.no_cpu) else None
    model = model.to(dev)

    wave_h, ids_h, mask_h = synth_inputs(B, rank)
    wave_h, ids_h, mask_h = wave_h.pin_memory(), ids_h.pin_memory(), mask_h.pin_memory()
    wave_d, ids_d, mask_d = wave_h.to(dev), ids_h.to(dev), mask_h.to(dev)

    def step(w, i, m, serial=False):
        if args.serial_towers or serial:
            a = model.encode_audio(w, max_patches=MAX_PATCHES)
            t = model.encode_text(i, m)
        else:
            a, t = model.encode_pairs(w, i, m, max_patches=MAX_PATCHES)
        if world > 1:
            return cdist.sharded_contrastive_logits(model, a, t)
        return model.similarity(a, t)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out = step(wave_d, ids_d, mask_d)
    barrier()

    # ---- timed region 1: inputs resident in HBM ---------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.caco_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = step(wave_d, ids_d, mask_d)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = lib.caco_launch_count() - launches0
    # ---- roofline leg: the same K steps again with both towers on ONE stream, CUDA events around every GEMM launch
    # (with the text tower on its side stream, as in the region above, per-kernel event pairs would also count the time a
    # kernel waits for the other stream's CTAs to drain, so the kernel timing is taken on the serialised replay)
    import ctypes as C
    lib.caco_gemm_profile(1)
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for _ in range(args.steps):
        step(wave_d, ids_d, mask_d, serial=True)
    r1.record()
    g_ms, g_fl = C.c_double(0), C.c_double(0)
    n_gemm = lib.caco_gemm_profile_read(C.byref(g_ms), C.byref(g_fl))
    lib.caco_gemm_profile(0)
    barrier()
    ms_serial = r0.elapsed_time(r1)
    clocks = sampler.stop()

    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": round(ms_total / args.steps, 3), "gpu_launches": int(launches)}))
        return

    # ---- timed region 2: end to end from pinned host memory, double-buffered H2D, D2H of the logits ---------------
    copy_s = torch.cuda.Stream(device=dev)
    comp_s = torch.cuda.current_stream()
    bufs = [(torch.empty_like(wave_d), torch.empty_like(ids_d), torch.empty_like(mask_d)) for _ in range(2)]
    out_h = [torch.empty(out[0].shape, dtype=torch.float32).pin_memory() for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]

    def e2e_loop(k):
        for i in range(k):
            b = i & 1
            with torch.cuda.stream(copy_s):
                if i >= 2:
                    copy_s.wait_event(free[b])
                bufs[b][0].copy_(wave_h, non_blocking=True)
                bufs[b][1].copy_(ids_h, non_blocking=True)
                bufs[b][2].copy_(mask_h, non_blocking=True)
                ready[b].record(copy_s)
            comp_s.wait_event(ready[b])
            o = step(*bufs[b])
            free[b].record(comp_s)
            out_h[b].copy_(o[0], non_blocking=True)

    e2e_loop(2)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    e2e_loop(args.steps)
    s1.record()
    barrier()
    ms_e2e = s0.elapsed_time(s1)
    checksum = float(out_h[(args.steps - 1) & 1].double().sum())       # the host really reads the result

    # ---- max over ranks ------------------------------------------------------------------------------------------
    t = torch.tensor([ms_total, ms_e2e, g_ms.value, ms_serial], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, gemm_ms, ms_serial = [float(x) for x in t.cpu()]
    pairs = B * world * args.steps
    value = pairs / (ms_total / 1e3)
    e2e_value = pairs / (ms_e2e / 1e3)
    peaks = _peaks()
    achieved = g_fl.value / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
    traffic, traffic_n = gemm_traffic_from_profile()
    roof = {"bound": "tensor", "kernel": "gemm_f16_kernel (tcgen05.mma kind::f16, fp16 operands, fp32 accumulate)",
            "achieved": round(achieved, 1) if achieved else None, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
            "frac": round(achieved / peaks["bf16_sustained"], 4) if achieved else None,
            "traffic": round(traffic) if traffic else None,
            "traffic_note": (f"DRAM bytes per launch, mean of the {traffic_n} audio-layer GEMMs (QKV, out-proj, fc1, fc2) in "
                             "profiles/r01_ncu_full_summary.csv; algorithmic bytes of the same four: 1081e6 per launch") if traffic else None,
            "peak_source": f"{peaks['src']} sustained bf16 GEMM (burst {peaks['bf16_burst']})",
            "launches_per_step": n_gemm // max(1, args.steps), "share_of_step": round(gemm_ms / ms_serial, 4),
            "timed_on": f"serial-stream replay of the same {args.steps} steps ({ms_serial / args.steps:.2f} ms/step), right after the main region",
            "whole_step_frac": round(value / world * GFLOP_PER_PAIR / 1e3 / peaks["bf16_sustained"], 4)}

    # ---- HBM-bound side of the metric ("encoder HBM GB/s vs peak"): the LayerNorm kernel at the tower's shape, timed alone
    from cacophony_b200 import ops as cops
    xr = torch.randn(B * MAX_PATCHES, 768, device=dev)
    gam = torch.ones(768, device=dev)
    for _ in range(3):
        cops.layernorm(xr, gam, gam, want_f32=False, want_f16=True)
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    h0.record()
    for _ in range(20):
        cops.layernorm(xr, gam, gam, want_f32=False, want_f16=True)
    h1.record()
    torch.cuda.synchronize()
    ln_ms = h0.elapsed_time(h1) / 20
    ln_bytes = xr.numel() * 6                                # 4 B read + 2 B written per element
    ln_gbs = ln_bytes / (ln_ms / 1e3) / 1e9
    roof_hbm = {"bound": "hbm", "kernel": "layernorm_kernel (fp32 in, fp16 out; 24 launches per clip batch in the audio tower)",
                "achieved": round(ln_gbs, 1), "peak": peaks["hbm"], "unit": "GB/s", "frac": round(ln_gbs / peaks["hbm"], 4),
                "bytes_per_launch": ln_bytes, "us_per_launch": round(ln_ms * 1e3, 2),
                "timed_on": "20 back-to-back launches on a 393 MB input (> L2), CUDA events",
                "traffic_note": "ncu: 0.393 GB read + 0.168 GB written per launch (profiles/r01_ncu_full_summary.csv)"}
    del xr

    cpu = None
    if sd_cpu is not None:
        v, dt, n_pass = cpu_pairs_per_s(sd_cpu, 16, 2, 1, budget_s=12.0)
        cores = os.cpu_count() or 1
        cpu = {"value": round(v, 3), "unit": UNIT, "cores": cores, "kind": "port", "cpu": cpu_model(),
               "sample": f"16 pairs per pass, 1 warm-up + {n_pass} timed passes ({dt:.2f} s each, {dt * n_pass:.0f} s of CPU work), "
                         f"fp32 torch CPU ops, {cores} threads"}

    if rank == 0:
        h2d = wave_h.numel() * 4 + ids_h.numel() * 8 + mask_h.numel() * 4
        d2h = out_h[0].numel() * 4
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate+residual", "data": "synthetic",
                "config": workload_config(B, world),
                "clocks": clocks,
                "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": round(ms_e2e / args.steps, 3), "checksum": checksum},
                "gpu_launches": int(launches), "roofline": roof, "roofline_hbm": roof_hbm, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="pairs per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--serial-towers", action="store_true", help="run the text tower after the audio tower on one stream")
    ap.add_argument("--profile", action="store_true", help="profiling run: resident-input region only (for ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
