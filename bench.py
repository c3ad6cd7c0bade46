#!/usr/bin/env python
"""Benchmark of the B200 ASR forward hot path (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload W] [--impl ours|reference|reference_cuda]
                    [--scaling strong|weak] [--precision bf16|fp16]

A *step* is one pass of the hot path over one batch of synthetic 16 kHz audio per GPU.  The metric is
BASELINE.json's: audio-seconds processed per wall-second, whole job (all N GPUs).  One JSON line is
printed by rank 0.

  value      inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the public module API with HOST (pinned) buffers: H2D copy of the audio and
             D2H read of the result inside the timed region
  roofline   dominant kernel, algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline  the oracle (a port of the reference's CPU path) timed on this box's host cores on a bounded
             sample of the same workload
  --impl reference   times the reference's CPU implementation of the path on the same config/metric: the UNMODIFIED
             reference modules from baseline/_ref (git-ignored copy made by baseline/install_ref.sh, it travels to the
             GPU box with the snapshot) or, if that directory is empty, the oracle port
  --impl reference_cuda   informative: the same unmodified modules .cuda() on this B200 (eager fp32 / autocast bf16)
  --gpus N > 1       strong scaling by default: the configuration's batch (256 / 128) is SHARDED over the ranks, no
             data-path collective, transcripts of all ranks gathered on rank 0 inside the e2e region; the line also
             carries `also.citrinet1024` (config 4) so that both networks of BASELINE.json's metric are on record
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SAMPLE_RATE = 16000
WORKLOADS = {
    # name: (description, per-GPU batch, seconds, nfilt)
    "features": ("FilterbankFeatures only: batch 64 x 20 s 16 kHz audio -> 64-bin log-mel (n_fft 512, win 320, hop 160)",
                 64, 20, 64),
    "quartznet15x5": ("QuartzNet 15x5 inference bf16, batch 256 x 15 s synthetic audio", 256, 15, 64),
    "citrinet1024": ("Citrinet-1024 with SqueezeExcite inference bf16, 1024-token vocab, batch 128 x 20 s", 128, 20, 80),
    "quartznet15x5_train": ("QuartzNet 15x5 training step (forward + backward + CTC loss + AdamW) with NCCL gradient "
                            "allreduce, per-GPU batch 32 x 15 s", 32, 15, 64),
}


#: bounded CPU sample (utterances per step) for the cpu_baseline / --impl reference legs: ~5-20 s of host work in total
CPU_SAMPLE_BATCH = {"features": 64, "quartznet15x5": 16, "citrinet1024": 4, "quartznet15x5_train": 4}


def traffic_per_launch(kernel: str, workload: str):
    """DRAM bytes (read + write) per launch of `kernel`, from the committed ncu pass of one forward
    (profiles/r01_dram_traffic.json, produced by tools/traffic_from_ncu.py); None if not measured."""
    p = os.path.join(ROOT, "profiles", "r01_dram_traffic.json")
    try:
        with open(p) as f:
            return json.load(f)[workload][kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def default_workload() -> str:
    try:
        from thunder_speech_b200 import runner  # noqa: F401
        return "quartznet15x5"
    except Exception:
        return "features"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int = 0):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            c = [v.strip() for v in r.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def _ref_harness():
    """baseline/ref_harness.py when the unmodified reference travelled with the repo (baseline/_ref/, git-ignored), else None."""
    try:
        from baseline import ref_harness as H
    except Exception:
        return None
    return H if H.available() else None


def cpu_reference_rate(workload: str, steps: int, warmup: int, sample_batch: int):
    """Times the reference's CPU implementation of the path on a bounded sample with all host threads: the UNMODIFIED
    reference modules from baseline/_ref when present (kind "reference"), else the oracle port (kind "port").
    Returns (audio-s/s, cores, kind, sample description, seconds per step)."""
    import torch

    desc, _, secs, nfilt = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    H = _ref_harness()
    if H is not None and workload != "quartznet15x5_train":
        from thunder_speech_b200 import synth

        model = H.ReferenceModel(workload, "cpu", seed=0)
        x = torch.from_numpy(synth.audio(sample_batch, secs * SAMPLE_RATE, 1234, "noise"))
        fn = lambda: model.predict(x)   # noqa: E731
        kind = "reference"
        sample = (f"B={sample_batch} x {secs} s, unmodified reference modules (baseline/_ref) on CPU fp32, predict() incl. "
                  "greedy decode")
    else:
        from oracle import ref_torch as RT

        fn, sample = RT.make_workload(workload, sample_batch, secs * SAMPLE_RATE, nfilt)
        kind = "port"
    for _ in range(max(1, warmup)):
        fn()
    ts = []
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    dt = float(np.mean(ts))
    return sample_batch * secs / dt, cores, kind, sample, dt


def run_reference(args):
    """`--impl reference`: the reference's own CPU path on the box's host cores (rank 0 only), bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    desc, B, secs, nfilt = WORKLOADS[args.workload]
    sb = CPU_SAMPLE_BATCH[args.workload]
    steps, warmup = min(args.steps, 5), min(args.warmup, 2)
    rate, cores, kind, sample, dt = cpu_reference_rate(args.workload, steps, warmup, sb)
    out = {
        "impl": "reference", "metric": "audio-sec/sec", "value": rate, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "global_batch": B, "seconds": secs,
                   "sample_batch_per_step": sb,
                   "note": f"CPU arm: each step is a bounded sample of {sb} utterances of the workload (not {B}); "
                           f"steps / warm-up clamped to {steps} / {warmup} so the run ends within minutes"},
        "cpu_baseline": {"value": rate, "unit": "audio-s/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def run_reference_cuda(args):
    """`--impl reference_cuda` (informative second baseline): the UNMODIFIED reference modules `.cuda()` on
    this B200 -- eager PyTorch (cuFFT / cuDNN / cuBLAS), fp32 and under autocast(bf16) -- same workload, same synthetic
    weights, predict() incl. greedy decode on host-resident results.  Rank 0 only."""
    import torch

    if int(os.environ.get("RANK", "0")) != 0:
        return
    H = _ref_harness()
    if H is None:
        print(json.dumps({"impl": "reference_cuda", "unavailable": "baseline/_ref is not populated (run baseline/install_ref.sh "
                          "in the build container)"}), flush=True)
        return
    from thunder_speech_b200 import synth

    desc, B, secs, nfilt = WORKLOADS[args.workload]
    B = args.batch or B
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    model = H.ReferenceModel(args.workload, dev, seed=0)
    x = torch.from_numpy(synth.audio(B, secs * SAMPLE_RATE, 1234, "noise")).to(dev)
    res = {}
    for tag, ctx in (("fp32", torch.autocast("cuda", enabled=False)), ("autocast_bf16", torch.autocast("cuda", dtype=torch.bfloat16))):
        try:
            with ctx:
                for _ in range(max(2, min(args.warmup, 3))):
                    model.predict(x)
                torch.cuda.synchronize()
                n = max(2, min(args.steps, 5))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(n):
                    model.predict(x)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            res[tag] = {"value": B * secs / (ms * 1e-3), "ms_per_step": ms, "steps": n}
        except Exception as e:  # e.g. a reference op without a bf16 CUDA kernel under autocast
            res[tag] = {"value": None, "error": f"{type(e).__name__}: {e}"[:200]}
    main = res["fp32"]
    out = {"impl": "reference_cuda", "metric": "audio-sec/sec", "value": main.get("value"), "unit": "audio-s/s", "n_gpus": 1,
           "steps": main.get("steps"), "warmup": max(2, min(args.warmup, 3)), "ms_per_step": main.get("ms_per_step"),
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": desc, "global_batch": B, "seconds": secs,
                      "note": "unmodified reference modules on cuda:0, eager PyTorch; inputs resident on the device"},
           "autocast_bf16": res["autocast_bf16"], "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def measure(args, workload, world, rank, local_rank, dev, barrier, with_cpu_baseline):
    """One workload on this job's ranks -> the metrics dict (rank 0) or None."""
    import torch
    import torch.distributed as dist

    from thunder_speech_b200 import _lib
    from thunder_speech_b200 import bench_workloads as BW
    from thunder_speech_b200.parallel import shard_bounds

    desc, B_total, secs, nfilt = WORKLOADS[workload]
    scaling = "weak" if workload == "quartznet15x5_train" else args.scaling    # config 5 is DEFINED per GPU (32 x 15 s each)
    if args.batch:
        B_total = args.batch
        desc += f" [batch overridden to {B_total}]"
    if scaling == "strong":
        lo, hi = shard_bounds(B_total, world, rank)      # BASELINE configs 3 / 4: "batch-sharded at 1/2/4/8"
        B = hi - lo
        global_batch = B_total
    else:
        B, global_batch = B_total, B_total * world
    assert B > 0, f"rank {rank} has no utterances: batch {B_total} over {world} ranks"
    N = secs * SAMPLE_RATE
    wl = BW.make(workload, B, N, nfilt, dev, rank, pcm16=args.pcm16)

    # ---- value: inputs resident in HBM ------------------------------------------------------------------
    for i in range(args.warmup):
        wl.step_device(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count() + wl.graph_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        wl.step_device(i)
    ev1.record()
    barrier()
    launches = _lib.launch_count() + wl.graph_launches() - launches0
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: host buffers, H2D + D2H (+ the transcript gather when world > 1) inside the timed region -
    wl.run_host(min(args.warmup, 3))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    wl.run_host(args.steps)
    e1.record()
    barrier()
    e2e_wall = time.perf_counter() - t0
    e2e_ms = max(e0.elapsed_time(e1), e2e_wall * 1e3)

    t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
    s = torch.tensor([float(wl.h2d_bytes), float(wl.d2h_bytes), float(launches)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    ms, e2e_ms = [float(v) for v in t.cpu()]
    h2d, d2h, launches_all = [int(v) for v in s.cpu()]

    # ---- per-kernel roofline pass (separate from the clean timed region) -------------------------------
    roof = wl.roofline(args.steps) if rank == 0 else None
    out = None
    if rank == 0:
        peaks, peak_src = measured_peaks()
        audio_s = global_batch * secs * args.steps
        value = audio_s / (ms * 1e-3)
        e2e = audio_s / (e2e_ms * 1e-3)
        if roof is not None:
            if roof["bound"] == "hbm":
                roof["peak"] = peaks["hbm_gbs"]
                roof["peak_kind"] = "measured copy bandwidth"
            else:
                # the GEMMs run inside a long, power-capped step: sustained cuBLAS peak; the burst figure is given beside it
                roof["peak"] = peaks["bf16_tflops_sustained"]
                roof["peak_kind"] = "sustained cuBLAS bf16 (kernel timed inside a long step)"
                roof["frac_of_burst_peak"] = roof["achieved"] / peaks["bf16_tflops"]
            roof["frac"] = roof["achieved"] / roof["peak"]
            roof["peak_source"] = peak_src
            roof["traffic"] = traffic_per_launch(roof["kernel"], workload)
        out = {
            "metric": "audio-sec/sec", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
            "config": {"workload": desc, "global_batch": global_batch, "per_gpu_batch": B, "seconds": secs,
                       "parallelism": f"batch-shard x{world}" + ("" if world == 1 else
                                                                   " (no data-path collective; transcripts gathered in e2e)"),
                       "precision": getattr(wl, "precision", wl.dtype),
                       "e2e_input": "int16 PCM from pinned host memory, scaled on the device (ts_pcm_ingest)" if args.pcm16
                       else "float32 audio from pinned host memory", "l2": wl.l2_note},
            "e2e": {"value": e2e, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches_all), "clocks": clocks, "roofline": roof,
        }
        if with_cpu_baseline:
            try:
                sb = CPU_SAMPLE_BATCH[workload]
                rate, cores, kind, sample, _ = cpu_reference_rate(workload, 3, 1, sb)
                out["cpu_baseline"] = {"value": rate, "unit": "audio-s/s", "cores": cores, "kind": kind, "sample": sample}
            except Exception as e:  # the baseline is informative; never lose the GPU numbers over it
                out["cpu_baseline"] = {"value": None, "unit": "audio-s/s", "cores": os.cpu_count(), "kind": "port",
                                       "sample": f"failed: {e}"}
    del wl
    import gc

    gc.collect()
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    import thunder_speech_b200 as tsb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.precision:
        tsb.set_default_precision(args.precision)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = measure(args, args.workload, world, rank, local_rank, dev, barrier,
                  with_cpu_baseline=(world == 1 and not args.no_cpu_baseline))
    # BASELINE.json's metric names BOTH networks: the default line carries Citrinet-1024 (config 4) as a second leg
    if args.workload == "quartznet15x5" and not args.no_also and not args.batch:
        also = measure(args, "citrinet1024", world, rank, local_rank, dev, barrier, with_cpu_baseline=False)
        if rank == 0:
            keep = ("value", "unit", "ms_per_step", "scaling", "dtype", "config", "e2e", "gpu_launches", "roofline")
            out["also"] = {"citrinet1024": {k: also[k] for k in keep}}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference_cuda"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): the configuration's batch is SHARDED over the GPUs (BASELINE configs 3 / 4: "
                         "256 -> 32 per GPU at 8); weak: every GPU runs the full per-GPU batch")
    ap.add_argument("--precision", default=None, choices=["bf16", "fp16"], help="row format of the inference path")
    ap.add_argument("--pcm16", action="store_true", help="e2e leg feeds int16 PCM (half the host-to-device bytes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the Citrinet-1024 leg of the default line")
    ap.add_argument("--batch", type=int, default=None, help="override the configuration's batch (experiments only)")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = default_workload()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference_cuda":
        run_reference_cuda(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
