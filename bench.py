#!/usr/bin/env python
"""bench.py — headline benchmark of the Cacophony inference hot path on B200 (driver contract: see the task prompt).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--workload pairs|zeroshot]
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
  cpu_baseline the UNMODIFIED reference (oracle/_ref archive, kind "reference"; the oracle port only if the archive is
               missing, kind "port") timed on this box's host cores on a bounded sample
  gpu_library_baseline   the same unmodified reference model moved to this GPU (torch's cuBLAS / native-MHA kernels, the bar
               SURVEY.md 2 names: eval_caco_torch.py:549 --device cuda): fp32 without TF32, TF32, fp16 autocast; pairs/s at B
  --impl reference   times the reference's own CPU implementation alone (the reference arm for this tier)
  --workload zeroshot    BASELINE config 5 (400 x 5 s clips, 50 prompts of 100 tokens) as a first-class workload: clips/s
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
    import glob
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full_summary.csv")))
    if not cands:
        return None, None, None
    p = cands[-1]                                  # the latest round's capture
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
    return (tot / n, n, os.path.relpath(p, ROOT)) if n else (None, None, None)


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
def _reference_or_none():
    """(create_caco_model, eval module) of the unmodified reference from oracle/_ref, or None (then the port is timed)."""
    try:
        from oracle import ref_loader
        return ref_loader.load()
    except Exception as e:                    # archive missing on this box
        sys.stderr.write(f"bench.py: reference archive unavailable ({e}); timing the oracle port instead\n")
        return None


def cpu_pairs_per_s(sd_cpu, sample_pairs: int, steps: int, warmup: int, budget_s: float = 0.0):
    """Times `steps` passes of `sample_pairs` pairs (after `warmup` untimed passes) through the reference's own CPU
    implementation with every host thread: its frontend one clip at a time (eval_caco_torch.py:181-206, as its drivers call
    it) and ONE batched CACO.forward (caco.py:242-261).  budget_s > 0: `steps` is a minimum and passes continue until about
    budget_s seconds of timed CPU work have been done (bench.py's cpu_baseline leg: ~10-30 s).
    Returns (pairs/s, s per pass, passes, kind)."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    wave, ids, mask = synth_inputs(sample_pairs, 99)
    ref = _reference_or_none()
    if ref is not None:
        create_ref, E = ref
        model = create_ref().eval()
        missing, unexpected = model.load_state_dict(sd_cpu, strict=False)
        assert not unexpected and all(k.startswith("decoder_module") for k in missing)
        cfg = E.DatasetConfig(patches_seq_len=MAX_PATCHES)

        def one_pass():
            bs = [E.prepare_audio_batch(wave[i:i + 1], cfg, "cpu") for i in range(sample_pairs)]
            ab = {k: torch.cat([b[k] for b in bs]) for k in bs[0]}
            return model(**ab, text_input_ids=ids, text_mask=mask.long())
        kind = "reference"
    else:
        from oracle import caco_oracle as O

        def one_pass():
            ab = O.prepare_audio_batch(list(wave.numpy()), MAX_PATCHES)
            return O.forward(sd_cpu, ab["audio_patches"], ab["audio_time_inds"], ab["audio_freq_inds"], ab["audio_mask"], ids, mask)
        kind = "port"
    times = []
    with torch.no_grad():
        i = 0
        while True:
            t0 = time.perf_counter()
            one_pass()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
            i += 1
            if len(times) >= steps and (budget_s <= 0 or sum(times) >= budget_s or len(times) >= 40):
                break
    dt = sum(times) / len(times)
    return sample_pairs / dt, dt, len(times), kind


def gpu_library_baseline(sd_cpu, batch_d, ids_d, mask_d, dev, steps: int = 5):
    """The unmodified reference model on THIS GPU through torch's own kernels (cuBLAS GEMMs, the native multi-head-attention
    fast path, eager LayerNorm / GELU): model.to(device) as eval_caco_torch.py:549 --device cuda would run it, CACO.forward on
    the same B pairs (patches already on the device: the frontend is outside this number, which favours the reference).
    Three settings: fp32 with TF32 off (the reference's numerics), TF32 on, fp16 autocast.  Returns a dict or None."""
    import torch
    ref = _reference_or_none()
    if ref is None:
        return None
    create_ref, _ = ref
    out = {"unit": UNIT, "batch": int(ids_d.shape[0]), "what": "unmodified reference CACO.forward on this GPU, torch "
           + torch.__version__ + " library kernels, device-resident patches, CUDA events"}
    try:
        model = create_ref().eval()
        model.load_state_dict(sd_cpu, strict=False)
        model = model.to(dev)
        args = dict(batch_d, text_input_ids=ids_d, text_mask=mask_d.long())
        B = int(ids_d.shape[0])

        def timed(fn):
            with torch.no_grad():
                for _ in range(2):
                    r = fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    r = fn()
                e1.record()
                torch.cuda.synchronize()
            return B * steps / (e0.elapsed_time(e1) / 1e3), r
        prev = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        out["fp32"], r32 = timed(lambda: model(**args))
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
        out["tf32"], _ = timed(lambda: model(**args))

        def autocast():
            with torch.autocast("cuda", dtype=torch.float16):
                return model(**args)
        out["fp16_autocast"], _ = timed(autocast)
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev
        for k in ("fp32", "tf32", "fp16_autocast"):
            out[k] = round(out[k], 1)
        out["_logits_fp32"] = r32[0]
        del model
        torch.cuda.empty_cache()
        return out
    except Exception as e:                    # never fail the bench because of the comparison leg
        out["error"] = f"{type(e).__name__}: {e}"[:200]
        return out


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
    steps, warm = max(1, args.steps), max(1, min(args.warmup, 1))
    value, dt, _, kind = cpu_pairs_per_s(sd, sample, steps, warm)
    cores = os.cpu_count() or 1
    what = ("the unmodified reference (src/caco_torch + src/eval/eval_caco_torch.py from oracle/_ref): its frontend per clip, "
            "one batched CACO.forward") if kind == "reference" else "the oracle port of the reference algorithm"
    desc = f"{sample} pairs per step (same clip/caption shape as the GPU arm), {what}, fp32 torch CPU ops, {cores} threads"
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": round(dt * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args.batch, max(1, args.gpus)), sample=desc),
            "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": kind, "sample": desc,
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
            if world > 1:
                return cdist.sharded_contrastive_logits(model, a, t)
            return model.similarity(a, t)
        if world > 1 and not args.diag_local:   # text embeddings gathered under the audio tower, audio embeddings after it
            return cdist.sharded_pairs_logits(model, w, i, m, max_patches=MAX_PATCHES, use_peer_memory=args.exchange == "peer")
        a, t = model.encode_pairs(w, i, m, max_patches=MAX_PATCHES)
        return model.similarity(a, t)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # N > 1: the step as a software pipeline (dist.PipelinedPairs): step k's towers + publication, then the logits of step
    # k - 1; K steps still produce K logits blocks inside the timed region (the last one by flush()).
    pipe = None
    if world > 1 and args.exchange == "peer" and not args.diag_local and not args.serial_towers and not args.no_pipeline:
        try:
            pipe = cdist.PipelinedPairs(model, B, MAX_PATCHES)
        except RuntimeError as e:
            sys.stderr.write(f"bench.py: {e}; plain per-step exchange\n")

    def run_steps(k, w, i, m):
        """k steps on resident inputs; returns the logits of the last one."""
        if pipe is None:
            o = None
            for _ in range(k):
                o = step(w, i, m)
            return o
        for _ in range(k):
            pipe.step(w, i, m)
        return pipe.flush()

    out = run_steps(max(args.warmup, 3), wave_d, ids_d, mask_d)
    barrier()

    # ---- timed region 1: inputs resident in HBM ---------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.caco_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    out = run_steps(args.steps, wave_d, ids_d, mask_d)
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

    if args.diag_local:
        tt = torch.tensor([ms_total / args.steps], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(tt) for _ in range(world)]
        if world > 1:
            dist.all_gather(allr, tt)
        else:
            allr = [tt]
        if rank == 0:
            print(json.dumps({"diag_local": True, "n_gpus": world, "per_rank_ms_per_step": [round(float(x), 3) for x in allr],
                              "clocks_rank0": clocks}))
        if world > 1:
            dist.destroy_process_group()
        return
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
            o = step(*bufs[b]) if pipe is None else pipe.step(*bufs[b])
            free[b].record(comp_s)
            if o is not None:
                out_h[b].copy_(o[0], non_blocking=True)
        if pipe is not None:
            out_h[k & 1].copy_(pipe.flush()[0], non_blocking=True)

    e2e_loop(2)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    e2e_loop(args.steps)
    s1.record()
    barrier()
    ms_e2e = s0.elapsed_time(s1)
    checksum = float(out_h[(args.steps - 1) & 1].double().sum())       # the host really reads the result

    # ---- what follows the towers (N > 1: all-gather of the audio embeddings + the two row-block similarity launches; N = 1:
    # the similarity kernel), timed alone on fixed embeddings: names the un-overlappable tail of the step
    a_fix = torch.nn.functional.normalize(torch.randn(B, 768, device=dev), dim=-1)
    t_fix = torch.nn.functional.normalize(torch.randn(B, 768, device=dev), dim=-1)

    ex = cdist.peer_exchange(model, B) if (world > 1 and args.exchange == "peer") else None

    def tail():
        if world > 1 and ex is not None:
            ex.scatter(t_fix, ex.TEXT)
            ex.scatter(a_fix, ex.AUDIO)
            at_b, _ = model.similarity(ex.local_rows(ex.AUDIO), ex.gathered(ex.TEXT), want_ta=False)
            return at_b, model.similarity(ex.local_rows(ex.TEXT), ex.gathered(ex.AUDIO), want_ta=False)[0]
        if world > 1:
            t_all = cdist.gather_embedding(t_fix, None, "text")
            at_b, _ = model.similarity(a_fix, t_all, want_ta=False)
            a_all = cdist.gather_embedding(a_fix, None, "audio")
            return at_b, model.similarity(t_fix, a_all, want_ta=False)[0]
        return model.similarity(a_fix, t_fix)
    for _ in range(3):
        tail()
    barrier()
    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    q0.record()
    for _ in range(20):
        tail()
    q1.record()
    barrier()
    tail_ms = q0.elapsed_time(q1) / 20

    # ---- max over ranks ------------------------------------------------------------------------------------------
    t = torch.tensor([ms_total, ms_e2e, g_ms.value, ms_serial, tail_ms], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        allr = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
        per_rank = [round(float(x[0]) / args.steps, 3) for x in allr]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, gemm_ms, ms_serial, tail_ms = [float(x) for x in t.cpu()]
    pairs = B * world * args.steps
    value = pairs / (ms_total / 1e3)
    e2e_value = pairs / (ms_e2e / 1e3)
    peaks = _peaks()
    achieved = g_fl.value / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
    traffic, traffic_n, traffic_src = gemm_traffic_from_profile()
    roof = {"bound": "tensor", "kernel": "gemm_f16_kernel (tcgen05.mma kind::f16, fp16 operands, fp32 accumulate)",
            "achieved": round(achieved, 1) if achieved else None, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
            "frac": round(achieved / peaks["bf16_sustained"], 4) if achieved else None,
            "traffic": round(traffic) if traffic else None,
            "traffic_note": (f"DRAM bytes per launch, mean of the {traffic_n} audio-layer GEMMs (QKV, out-proj, fc1, fc2) in "
                             f"{traffic_src}; algorithmic bytes of the same four: 1081e6 per launch") if traffic else None,
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
                "traffic_note": "ncu: 0.393 GB read + 0.168 GB written per launch (profiles/r01_ncu_full_summary.csv; kernel unchanged)"}
    del xr

    cpu, lib_base = None, None
    if sd_cpu is not None:
        # the unmodified reference on THIS GPU through torch's library kernels, on the same inputs
        ab = cb.prepare_audio_batch(wave_d, cb.DatasetConfig(patches_seq_len=MAX_PATCHES), dev)
        lib_base = gpu_library_baseline(sd_cpu, ab, ids_d, mask_d, dev) if not args.no_lib else None
        if lib_base is not None and "_logits_fp32" in lib_base:
            ref_logits = lib_base.pop("_logits_fp32")
            ours = step(wave_d, ids_d, mask_d)[0]
            lib_base["max_abs_dlogit_vs_ours"] = round(float((ours - ref_logits).abs().max()), 6)
            lib_base["speedup_e2e_over"] = {k: round(e2e_value / lib_base[k], 2) for k in ("fp32", "tf32", "fp16_autocast")
                                            if lib_base.get(k)}
            del ref_logits
        del ab
        v, dt, n_pass, kind = cpu_pairs_per_s(sd_cpu, 16, 2, 1, budget_s=12.0)
        cores = os.cpu_count() or 1
        cpu = {"value": round(v, 3), "unit": UNIT, "cores": cores, "kind": kind, "cpu": cpu_model(),
               "sample": f"16 pairs per pass, 1 warm-up + {n_pass} timed passes ({dt:.2f} s each, {dt * n_pass:.0f} s of CPU work), "
                         + ("unmodified reference (oracle/_ref): frontend per clip + one batched CACO.forward, " if kind == "reference"
                            else "oracle port, ") + f"fp32 torch CPU ops, {cores} threads"}

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
                "gpu_launches": int(launches), "roofline": roof, "roofline_hbm": roof_hbm, "cpu_baseline": cpu,
                "gpu_library_baseline": lib_base,
                "tail": {"ms": round(tail_ms, 4), "what": (("exchange of both modalities' embeddings (" +
                         ("L2-norm kernels storing into every peer's matrix over NVLink + flag waits" if ex is not None else
                          "two NCCL all_gather_into_tensor, 0.79 MB per rank each") +
                         ") + two [B, N*B] similarity launches, timed alone; in the step the text exchange is hidden under "
                         "the audio tower") if world > 1 else "similarity kernel (both directions), timed alone"),
                         "exchange": (None if world == 1 else "peer-memory" if ex is not None else "nccl"),
                         "pipelined": pipe is not None},
                "per_rank_ms_per_step": per_rank}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------
# BASELINE config 5: zero-shot classification, ESC-50 shape: 400 clips of 80 000 samples (5 s), 50 prompts
# padded to 100 tokens with 8-12 valid ones; clips sharded across ranks, prompts encoded by every rank, top-1 on the device
# ----------------------------------------------------------------------------------------------------------------
def run_zeroshot(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import cacophony_b200 as cb
    from cacophony_b200 import _lib as L
    from cacophony_b200 import dist as cdist
    from cacophony_b200 import eval as ev

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = L.load()
    N_CLIPS, N_CLASSES, CLIP, T = 400, 50, 80000, 100
    torch.manual_seed(0)
    model = cb.create_caco_model()
    sd_cpu = {k: v.clone() for k, v in model.state_dict().items()} if rank == 0 else None
    model = model.to(dev)
    g = torch.Generator().manual_seed(4321)
    waves_h = (0.1 * (2.0 * torch.rand(N_CLIPS, CLIP, generator=g) - 1.0)).float().pin_memory()
    ids_h = torch.full((N_CLASSES, T), 1, dtype=torch.int64)
    mask_h = torch.zeros(N_CLASSES, T, dtype=torch.float32)
    for c in range(N_CLASSES):
        n = int(torch.randint(8, 13, (1,), generator=g))
        ids_h[c, :n] = torch.randint(3, 50265, (n,), generator=g)
        ids_h[c, 0], ids_h[c, n - 1] = 0, 2
        mask_h[c, :n] = 1
    ids_h, mask_h = ids_h.pin_memory(), mask_h.pin_memory()
    lo, hi = cdist.shard_range(N_CLIPS, rank, world)
    lens = torch.full((hi - lo,), CLIP, dtype=torch.int32, device=dev)
    cfg = cb.DatasetConfig(patches_seq_len=MAX_PATCHES)

    # trimmed shapes from HOST knowledge (clip length, prompt lengths): no device read-back inside a step
    p_trim = max(8, min(cfg.patches_seq_len, (((CLIP + 159) // 160) // 16) * 8))
    t_trim = min(T, -(-int(mask_h.sum(1).max()) // 8) * 8)

    def step(w_dev, ids_dev, mask_dev, trim):
        """class embeddings (every rank), this rank's clips, logits + top-1 on the device, predictions gathered."""
        if trim:
            t = model.encode_text(ids_dev[:, :t_trim].contiguous(), mask_dev[:, :t_trim].contiguous())
            a = model.encode_audio(w_dev, max_patches=p_trim, lengths=lens)
        else:
            t = model.encode_text(ids_dev, mask_dev)
            a = model.encode_audio(w_dev, max_patches=cfg.patches_seq_len, lengths=lens)
        top = ev.zero_shot_topk(model, a, t, 1)
        return cdist.gather_rows(top, N_CLIPS) if world > 1 else top

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    w_d, ids_d, mask_d = waves_h[lo:hi].to(dev), ids_h.to(dev), mask_h.to(dev)
    res = {}
    launches = 0
    for trim in (False, True):
        for _ in range(max(args.warmup, 3)):
            top = step(w_d, ids_d, mask_d, trim)
        barrier()
        l0 = lib.caco_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            top = step(w_d, ids_d, mask_d, trim)
        e1.record()
        barrier()
        if not trim:
            launches = lib.caco_launch_count() - l0
        # end to end: pinned host waveforms / ids -> device inside the timed region, predictions read back
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            wd = waves_h[lo:hi].to(dev, non_blocking=True)
            top_h = step(wd, ids_h.to(dev, non_blocking=True), mask_h.to(dev, non_blocking=True), trim).cpu()
        s1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1), s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[trim] = (float(t[0]) / args.steps, float(t[1]) / args.steps, top_h)
    agree = None
    if rank == 0:
        # top-1 agreement with the CPU oracle on a 16-clip sample (the oracle takes ~0.3 s per clip)
        from oracle import caco_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        sample = list(range(0, N_CLIPS, N_CLIPS // 16))[:16]
        with torch.no_grad():
            ab = O.prepare_audio_batch([waves_h[i].numpy() for i in sample], MAX_PATCHES)
            a_ref, _ = O.get_audio_embedding(sd_cpu, ab["audio_patches"], ab["audio_time_inds"], ab["audio_freq_inds"],
                                             ab["audio_mask"], normalize=True)
            t_ref, _ = O.get_text_embedding(sd_cpu, ids_h, mask_h, normalize=True)
            logits = float(np.exp(sd_cpu["logit_scale"].item())) * a_ref @ t_ref.T
        ref_top = logits.argmax(-1)
        srt = logits.sort(-1).values
        clear = (srt[:, -1] - srt[:, -2]) > 2e-2            # ties aside
        ours = res[False][2][sample, 0].long()
        ours_trim = res[True][2][sample, 0].long()
        agree = {"sample_clips": len(sample), "clear_margin": int(clear.sum()),
                 "top1_equal_untrimmed": int((ours[clear] == ref_top[clear]).sum()),
                 "top1_equal_trimmed": int((ours_trim[clear] == ref_top[clear]).sum()),
                 "trimmed_equals_untrimmed_all_400": bool(torch.equal(res[False][2], res[True][2]))}
    if rank == 0:
        ms, ms_e2e, _ = res[False]
        ms_t, ms_e2e_t, _ = res[True]
        line = {"metric": "zero-shot classification clips/sec (ESC-50 shape: 400 x 5 s clips, 50 prompts)", "unit": "clips/s",
                "value": round(N_CLIPS / (ms / 1e3), 1), "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f16 operands / f32 accumulate+residual", "data": "synthetic",
                "config": {"workload": "BASELINE config 5: 400 clips x 80000 samples (248 valid of 500 token slots, all 500 computed "
                                       "as the reference does), 50 prompts padded to 100 tokens (8-12 valid), class embeddings "
                                       "recomputed every step on every rank, logits + top-1 on the device",
                           "clips_per_gpu": hi - lo, "parallelism": f"dp{world}"},
                "trim_padding": {"value": round(N_CLIPS / (ms_t / 1e3), 1), "ms_per_step": round(ms_t, 3),
                                 "what": "tower run on the 248 valid token slots / 16 prompt columns only; same predictions"},
                "e2e": {"value": round(N_CLIPS / (ms_e2e / 1e3), 1), "unit": "clips/s", "ms_per_step": round(ms_e2e, 3),
                        "h2d_bytes_per_step": (hi - lo) * CLIP * 4 + N_CLASSES * T * 12, "d2h_bytes_per_step": N_CLIPS * 4,
                        "trim_padding_value": round(N_CLIPS / (ms_e2e_t / 1e3), 1)},
                "gpu_launches": int(launches), "oracle_agreement": agree}
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline and gpu_library_baseline legs")
    ap.add_argument("--no-lib", action="store_true", help="skip the gpu_library_baseline leg")
    ap.add_argument("--workload", default="pairs", choices=["pairs", "zeroshot"],
                    help="pairs = BASELINE configs 3/4 (headline); zeroshot = config 5 (400 x 5 s clips, 50 prompts of 100 tokens)")
    ap.add_argument("--serial-towers", action="store_true", help="run the text tower after the audio tower on one stream")
    ap.add_argument("--profile", action="store_true", help="profiling run: resident-input region only (for ncu)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: how the embeddings are exchanged: peer = stores into peer memory fused into the L2-norm kernel "
                         "(default; falls back to nccl when symmetric memory is unavailable), nccl = two all-gathers")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="N > 1: compute every step's logits inside the step (all ranks meet every step) instead of one step later")
    ap.add_argument("--diag-local", action="store_true",
                    help="diagnostic (N > 1): no exchange, every rank computes its local logits only and reports its own ms/step "
                         "— separates GPU-to-GPU spread from the cost of the collective; not a bench line")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "zeroshot":
        run_zeroshot(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
