#!/usr/bin/env python
"""Benchmark of the B200 ASR forward hot path (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload W] [--impl ours|reference]

A *step* is one pass of the hot path over one batch of synthetic 16 kHz audio per GPU.  The metric is
BASELINE.json's: audio-seconds processed per wall-second, whole job (all N GPUs).  One JSON line is
printed by rank 0.

  value      inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the public module API with HOST (pinned) buffers: H2D copy of the audio and
             D2H read of the result inside the timed region
  roofline   dominant kernel, algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline  the oracle (a port of the reference's CPU path) timed on this box's host cores on a bounded
             sample of the same workload
  --impl reference   times the reference's CPU implementation of the path (the oracle port: the reference is
             pure Python over PyTorch and cannot travel to the GPU box) on the same config/metric
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
def cpu_port_rate(workload: str, steps: int, warmup: int, sample_batch: int):
    """Times the CPU port of the reference path (oracle) on a bounded sample; returns (audio-s/s, cores, sample)."""
    import torch

    from oracle import ref_torch as RT

    desc, _, secs, nfilt = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fn, sample = RT.make_workload(workload, sample_batch, secs * SAMPLE_RATE, nfilt)
    for _ in range(max(1, warmup)):
        fn()
    ts = []
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    dt = float(np.mean(ts))
    return sample_batch * secs / dt, cores, sample, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    desc, B, secs, nfilt = WORKLOADS[args.workload]
    sb = CPU_SAMPLE_BATCH[args.workload]
    rate, cores, sample, dt = cpu_port_rate(args.workload, min(args.steps, 5), min(args.warmup, 2), sb)
    out = {
        "impl": "reference", "metric": "audio-sec/sec", "value": rate, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": min(args.steps, 5), "warmup": min(args.warmup, 2), "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "per_gpu_batch": B, "seconds": secs},
        "cpu_baseline": {"value": rate, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from thunder_speech_b200 import _lib, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    desc, B, secs, nfilt = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
        desc += f" [per-GPU batch overridden to {B}]"
    N = secs * SAMPLE_RATE
    from thunder_speech_b200 import bench_workloads as BW

    wl = BW.make(args.workload, B, N, nfilt, dev, rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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

    # ---- e2e: host buffers, H2D + D2H inside the timed region ------------------------------------------
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
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = [float(v) for v in t.cpu()]

    # ---- per-kernel roofline pass (separate from the clean timed region) -------------------------------
    roof = wl.roofline(args.steps) if rank == 0 else None

    if rank == 0:
        peaks, peak_src = measured_peaks()
        audio_s = B * secs * world * args.steps
        value = audio_s / (ms * 1e-3)
        e2e = audio_s / (e2e_ms * 1e-3)
        if roof is not None:
            if roof["bound"] == "hbm":
                roof["peak"] = peaks["hbm_gbs"]
            else:
                roof["peak"] = peaks["bf16_tflops_sustained"] * (1.0 if roof.get("unit") == "TFLOP/s" else 1.0)
            roof["frac"] = roof["achieved"] / roof["peak"]
            roof["peak_source"] = peak_src
            roof["traffic"] = traffic_per_launch(roof["kernel"], args.workload)
        out = {
            "metric": "audio-sec/sec", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
            "config": {"workload": desc, "per_gpu_batch": B, "seconds": secs, "parallelism": f"batch-shard x{world}",
                       "l2": wl.l2_note},
            "e2e": {"value": e2e, "unit": "audio-s/s", "h2d_bytes_per_step": wl.h2d_bytes, "d2h_bytes_per_step": wl.d2h_bytes},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                sb = CPU_SAMPLE_BATCH[args.workload]
                rate, cores, sample, _ = cpu_port_rate(args.workload, 3, 1, sb)
                out["cpu_baseline"] = {"value": rate, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample}
            except Exception as e:  # the baseline is informative; never lose the GPU numbers over it
                out["cpu_baseline"] = {"value": None, "unit": "audio-s/s", "cores": os.cpu_count(), "kind": "port",
                                       "sample": f"failed: {e}"}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=None, help="override the per-GPU batch (experiments only)")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = default_workload()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
