"""Workload objects used by ``bench.py`` (product side: synthetic inputs, module construction, step
functions, per-kernel roofline pass).  Nothing here touches ``oracle/``."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, synth
from .quartznet.transform import FilterbankFeatures


class _Base:
    dtype = "f32"
    l2_note = ""
    h2d_bytes = 0
    d2h_bytes = 0

    def graph_launches(self) -> int:
        """Kernel launches that happened inside CUDA-graph replays (not seen by ts_launch_count)."""
        return 0

    def run_host(self, steps):
        for i in range(steps):
            self.step_host(i)


class FeaturesWorkload(_Base):
    """BASELINE config 2: FilterbankFeatures only."""

    NBUF = 4

    def __init__(self, B, N, nfilt, dev, rank):
        self.B, self.N, self.nfilt, self.dev = B, N, nfilt, dev
        self.F = 1 + N // 160
        self.fb = FilterbankFeatures(nfilt=nfilt).eval().to(dev)
        host = torch.from_numpy(synth.audio(B, N, 1234 + rank, "noise"))
        self.host_audio = host.pin_memory()
        # rotate over NBUF distinct input buffers so that inputs (NBUF x 82 MB) exceed the 126 MB L2
        self.audio = [self.host_audio.to(dev) * (1.0 + 0.01 * i) for i in range(self.NBUF)]
        self.lengths = torch.full((B,), N, dtype=torch.int64, device=dev)
        self.host_out = torch.empty((B, nfilt, self.F), dtype=torch.float32).pin_memory()
        self.stage = torch.empty((B, N), dtype=torch.float32, device=dev)
        self.h2d_bytes = B * N * 4 + B * 8
        self.d2h_bytes = B * nfilt * self.F * 4
        self.l2_note = f"inputs rotate over {self.NBUF} distinct buffers ({self.NBUF * B * N * 4 / 1e6:.0f} MB) > 126 MB L2"
        self.dtype = "f32"

    def step_device(self, i):
        return self.fb(self.audio[i % self.NBUF], self.lengths)

    def step_host(self, i):
        self.stage.copy_(self.host_audio, non_blocking=True)
        lens = torch.full((self.B,), self.N, dtype=torch.int64).to(self.dev, non_blocking=True)
        f, fl = self.fb(self.stage, lens)
        self.host_out.copy_(f, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.host_out

    def roofline(self, steps):
        """CUDA events around the dominant kernel (logmel_kernel) alone, launched through the C ABI."""
        L = _lib.lib()
        t = self.fb._device_tables(self.dev)
        out = torch.empty((self.B, self.nfilt, self.F), dtype=torch.float32, device=self.dev)
        st = torch.cuda.current_stream().cuda_stream
        evs = []
        for i in range(steps + 3):
            a = self.audio[i % self.NBUF]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(L.ts_logmel(a.data_ptr(), self.B, self.N, 512, 160, 0.97, t["window_full"].data_ptr(),
                                   t["win_lo"], t["win_hi"], t["twiddle"].data_ptr(), t["mel_start"].data_ptr(),
                                   t["mel_count"].data_ptr(), t["mel_off"].data_ptr(), t["mel_w"].data_ptr(),
                                   self.nfilt, t["mel_w"].numel(), out.data_ptr(), st), "ts_logmel")
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        ms = float(np.mean([a.elapsed_time(b) for a, b in evs[3:]]))
        alg = self.B * self.N * 4 + self.B * self.nfilt * self.F * 4
        return {"kernel": "logmel_kernel", "bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "unit": "GB/s",
                "algorithmic_bytes": alg, "avg_kernel_ms": ms, "traffic": None}


def make(name, B, N, nfilt, dev, rank, pcm16: bool = False):
    if name == "features":
        return FeaturesWorkload(B, N, nfilt, dev, rank)
    from . import runner  # encoder workloads need the model runner

    return runner.make_bench_workload(name, B, N, nfilt, dev, rank, pcm16)
