"""Model builders for the benchmark / smoke configurations and the bench workload object.

``build_model(name)`` assembles the reference architectures named in BASELINE.json from this package's module
shells with random-init weights from ``synth`` (no checkpoints offline):

    quartznet5x5   QuartznetEncoder(repeat_blocks=1), 64 mel, V=29      (config 1)
    quartznet15x5  QuartznetEncoder(repeat_blocks=3), 64 mel, V=29      (config 3)
    citrinet1024   CitrinetEncoder(NeMo citrinet_1024 body), 80 mel, V=1025 (config 4)
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from . import ops, synth
from .blocks import conv1d_decoder
from .citrinet.blocks import CitrinetEncoder
from .module import CTCModule
from .quartznet.blocks import QuartznetEncoder
from .quartznet.transform import FilterbankFeatures
from .text_processing import BatchTextTransformer


def _load(module, state: Dict[str, np.ndarray]):
    module.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in state.items()}, strict=True)
    return module


def build_model(name: str, device, seed: int = 0) -> CTCModule:
    if name in ("quartznet5x5", "quartznet15x5"):
        rep = 1 if name == "quartznet5x5" else 3
        enc = _load(QuartznetEncoder(repeat_blocks=rep),
                    synth.encoder_state(synth.quartznet_block_list(repeat_blocks=rep), seed=seed))
        dec = conv1d_decoder(1024, 29)
        dec.load_state_dict({k: torch.from_numpy(v) for k, v in synth.decoder_state(1024, 29, seed + 1).items()})
        fb = FilterbankFeatures(nfilt=64)
        tt = BatchTextTransformer(synth.quartznet_vocab())
    elif name == "citrinet1024":
        c = synth.CITRINET_1024
        enc = _load(CitrinetEncoder(c["filters"], c["kernel_sizes"], c["strides"], feat_in=80),
                    synth.encoder_state(synth.citrinet_block_list(c["filters"], c["kernel_sizes"], c["strides"], 80),
                                        seed=seed, se=True))
        dec = conv1d_decoder(640, 1025)
        dec.load_state_dict({k: torch.from_numpy(v) for k, v in synth.decoder_state(640, 1025, seed + 1).items()})
        fb = FilterbankFeatures(nfilt=80)
        tt = BatchTextTransformer(synth.citrinet_vocab(1024))
    else:
        raise ValueError(name)
    return CTCModule(enc, dec, fb, tt).eval().to(device)


class ModelWorkload:
    """bench.py workload: ``predict()`` of a full model on a per-GPU batch of synthetic audio."""

    NBUF = 2
    dtype = "bf16"

    def __init__(self, name, B, N, nfilt, dev, rank, pcm16: bool = False):
        from . import get_default_precision

        self.name, self.B, self.N, self.dev = name, B, N, dev
        self.pcm16 = pcm16
        self.model = build_model(name, dev)
        self.precision = self.model.precision or get_default_precision()
        self.dtype = "f16" if self.precision == "fp16" else "bf16"
        host = torch.from_numpy(synth.audio(B, N, 1234 + rank, "noise"))
        self.host_audio = host.pin_memory()
        self.audio = [(self.host_audio.to(dev) * (1.0 + 0.01 * i)).contiguous() for i in range(self.NBUF)]
        self.stage = torch.empty((B, N), dtype=torch.float32, device=dev)
        # e2e input: float32 audio (the reference's predict() contract) or, opt-in, int16 PCM as stored in wav files
        self.host_pcm = (host * 32768.0).round().clamp(-32768, 32767).to(torch.int16).pin_memory() if pcm16 else None
        self.h2d_bytes = B * N * (2 if pcm16 else 4)
        for a in self.audio:      # resident input buffers: one graph each, read in place (no staging copy in the timed step)
            ids, col, cnt = self.model.predict_ids_graphed(a, in_place=True)
        torch.cuda.synchronize()
        self.T_out = col.shape[1]
        self.d2h_bytes = col.numel() * 8 + cnt.numel() * 4
        self.l2_note = (f"audio batch {B * N * 4 / 1e6:.0f} MB and every activation tensor exceed the 126 MB L2; "
                        f"inputs rotate over {self.NBUF} buffers")

    def graph_launches(self) -> int:
        return sum(g.kernels_per_replay * g.replays for g in self.model._graphs.values())

    def step_device(self, i):
        return self.model.predict_ids_graphed(self.audio[i % self.NBUF], in_place=True)

    def step_host(self, i):
        """The single-shot user call: host audio in, transcriptions out (copy, compute, decode in sequence)."""
        self.stage.copy_(self.host_audio, non_blocking=True)
        return self.model.predict_graphed(self.stage)

    def run_host(self, steps):
        """The serving-loop user call: ``predict_stream`` over ``steps`` pinned host batches -- every step still
        copies its own audio host->device and its token ids device->host inside the timed region; the copies
        overlap the neighbouring steps' compute."""
        from .parallel import sharded_predict_stream

        # every rank streams ITS shard of each batch; the transcripts of all ranks are gathered at the end of the stream
        src = self.host_pcm if self.pcm16 else self.host_audio
        out = sharded_predict_stream(self.model, (src for _ in range(steps)), presharded=True)
        return sum(len(t) for t in out)

    def roofline(self, steps):
        """Eager pass with CUDA events around every kernel launch; reports the kernel class with the largest
        share of the step against its own roofline (HBM for depthwise, tensor for the GEMMs)."""
        n = max(2, min(steps, 3))
        self.model.predict_ids(self.audio[0])
        torch.cuda.synchronize()
        ops.PROFILE = []
        try:
            for i in range(n):
                # keep the GPU busy while the CPU queues the ~160-300 launches of this pass, otherwise every
                # event pair would also time the host-side launch latency (the eager path is CPU-bound)
                torch.cuda._sleep(int(40e6))
                self.model.predict_ids(self.audio[i % self.NBUF])
            torch.cuda.synchronize()
            recs = ops.PROFILE
        finally:
            ops.PROFILE = None
        agg = {}
        for name, meta, e0, e1 in recs:
            a = agg.setdefault(name, dict(ms=0.0, bytes=0, flops=0, calls=0))
            a["ms"] += e0.elapsed_time(e1)
            a["bytes"] += meta["bytes"]
            a["flops"] += meta["flops"]
            a["calls"] += 1
        total_ms = sum(a["ms"] for a in agg.values())
        shares = {k: {"ms_per_step": a["ms"] / n, "share": a["ms"] / total_ms, "launches_per_step": a["calls"] // n,
                      "GBps": a["bytes"] / (a["ms"] * 1e-3) / 1e9, "TFLOPs": a["flops"] / (a["ms"] * 1e-3) / 1e12}
                  for k, a in agg.items()}
        top = max(agg, key=lambda k: agg[k]["ms"])
        a = agg[top]
        if top == "pw_gemm":
            out = {"kernel": "pw_gemm_pair_kernel", "bound": "tensor", "achieved": a["flops"] / (a["ms"] * 1e-3) / 1e12,
                   "unit": "TFLOP/s", "algorithmic_flops_per_step": a["flops"] // n}
        else:
            out = {"kernel": "dw_tma_kernel", "bound": "hbm", "achieved": a["bytes"] / (a["ms"] * 1e-3) / 1e9,
                   "unit": "GB/s", "algorithmic_bytes_per_step": a["bytes"] // n}
        out["avg_kernel_ms"] = a["ms"] / a["calls"]
        out["launches_per_step"] = a["calls"] // n
        out["traffic"] = None
        out["per_kernel"] = shares
        return out


class TrainWorkload:
    """bench.py workload (BASELINE config 5): one optimisation step of QuartzNet 15x5 -- features, train()-mode forward,
    CTC loss, backward, gradient all-reduce (NCCL when world > 1), AdamW -- on a per-GPU batch of synthetic audio with
    random 29-character transcripts."""

    NBUF = 2
    dtype = "bf16"
    LABEL_LEN = 120   # characters per 15 s utterance (~8 per second)

    def __init__(self, name, B, N, nfilt, dev, rank):
        from .train import CTCTrainStep

        self.name, self.B, self.N, self.dev = name, B, N, dev
        self.model = build_model("quartznet15x5", dev)
        self.model.train()     # like the reference's training_step: dither (1e-5) in the front-end, batch-stat BatchNorm
        self.trainer = CTCTrainStep(self.model, lr=1e-4)
        rng = np.random.Generator(np.random.PCG64(99 + rank))
        self.host_audio = torch.from_numpy(synth.audio(B, N, 1234 + rank, "noise")).pin_memory()
        self.host_lens = torch.from_numpy(synth.ragged_lengths(B, N, 7 + rank)).pin_memory()
        self.host_y = torch.from_numpy(rng.integers(0, 28, (B, self.LABEL_LEN)).astype(np.int64)).pin_memory()
        self.host_ylen = torch.from_numpy(rng.integers(self.LABEL_LEN // 2, self.LABEL_LEN + 1, B).astype(np.int64)).pin_memory()
        self.audio = [(self.host_audio.to(dev) * (1.0 + 0.01 * i)).contiguous() for i in range(self.NBUF)]
        self.lens, self.y, self.ylen = self.host_lens.to(dev), self.host_y.to(dev), self.host_ylen.to(dev)
        self.stage = torch.empty((B, N), dtype=torch.float32, device=dev)
        self.h2d_bytes = B * N * 4 + B * 8 + self.host_y.numel() * 8 + B * 8
        self.d2h_bytes = 4
        self.l2_note = ("the step's saved activations (several GB) exceed the 126 MB L2; audio rotates over "
                        f"{self.NBUF} buffers")
        self.last_loss = None

    def graph_launches(self) -> int:
        return self.trainer.graph_launches()

    def step_device(self, i):
        self.last_loss = self.trainer.step(self.audio[i % self.NBUF], self.lens, self.y, self.ylen)
        return self.last_loss

    def step_host(self, i):
        self.stage.copy_(self.host_audio, non_blocking=True)
        lens = self.host_lens.to(self.dev, non_blocking=True)
        y = self.host_y.to(self.dev, non_blocking=True)
        ylen = self.host_ylen.to(self.dev, non_blocking=True)
        return float(self.trainer.step(self.stage, lens, y, ylen).item())

    def run_host(self, steps):
        """The training-loop user call: ``CTCTrainStep.fit_stream`` over ``steps`` pinned host batches -- every step
        copies its own audio + labels host->device and its own loss device->host inside the timed region; the copies
        overlap the neighbouring step's compute."""
        batch = (self.host_audio, self.host_lens, self.host_y, self.host_ylen)
        for loss in self.trainer.fit_stream(batch for _ in range(steps)):
            self.last_loss = loss

    def roofline(self, steps):
        """Eager step with CUDA events around every kernel launch of this library (torch's own kernels -- decoder,
        log_softmax, CTC loss, AdamW, the [C]-vector arithmetic -- are not in the list; their share is reported as
        the graph replay time of forward + loss + backward is reported next to the eager per-kernel sum)."""
        n = 2
        tr = self.trainer
        fb = lambda i: tr._forward_backward(self.audio[i % self.NBUF], self.lens, self.y, self.ylen, update_running=False)
        fb(0)
        torch.cuda.synchronize()
        # graph replay time of the same work (what the timed region runs), for the share-of-step column
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for i in range(n):   # forward + loss + backward only: no collective on this rank-0-only pass
            tr.loss_and_grads(self.audio[i % self.NBUF], self.lens, self.y, self.ylen)
        g1.record()
        torch.cuda.synchronize()
        step_ms = g0.elapsed_time(g1) / n
        ops.PROFILE = []
        try:
            for i in range(n):
                # keep the GPU busy while the CPU queues the ~1400 launches of the eager pass (see ModelWorkload.roofline)
                torch.cuda._sleep(int(150e6))
                fb(i)
            torch.cuda.synchronize()
            recs = ops.PROFILE
        finally:
            ops.PROFILE = None
        agg = {}
        for name, meta, a0, a1 in recs:
            a = agg.setdefault(name, dict(ms=0.0, bytes=0, flops=0, calls=0))
            a["ms"] += a0.elapsed_time(a1)
            a["bytes"] += meta["bytes"]
            a["flops"] += meta["flops"]
            a["calls"] += 1
        total_ms = sum(a["ms"] for a in agg.values())
        shares = {k: {"ms_per_step": a["ms"] / n, "share_of_step": a["ms"] / n / step_ms, "launches_per_step": a["calls"] // n,
                      "GBps": a["bytes"] / (a["ms"] * 1e-3) / 1e9, "TFLOPs": a["flops"] / (a["ms"] * 1e-3) / 1e12}
                  for k, a in agg.items()}
        top = max(agg, key=lambda k: agg[k]["ms"])
        a = agg[top]
        tensor = top in ("pw_gemm", "pw_wgrad")
        out = {"kernel": {"pw_gemm": "pw_gemm_pair_kernel", "pw_wgrad": "pw_wgrad_kernel", "dw_conv": "dw_tma_kernel"}.get(top, top + "_kernel"),
               "bound": "tensor" if tensor else "hbm",
               "achieved": (a["flops"] / (a["ms"] * 1e-3) / 1e12) if tensor else (a["bytes"] / (a["ms"] * 1e-3) / 1e9),
               "unit": "TFLOP/s" if tensor else "GB/s", "avg_kernel_ms": a["ms"] / a["calls"],
               "launches_per_step": a["calls"] // n, "traffic": None, "per_kernel": shares,
               # the eager per-kernel times add up to MORE than the graph step: inside the graph the weight-gradient
               # kernels run on a forked stream and PDL overlaps prologues with the predecessors' tails
               "eager_kernel_sum_ms": total_ms / n, "graph_step_ms": step_ms}
        return out


def make_bench_workload(name, B, N, nfilt, dev, rank, pcm16: bool = False):
    if name == "quartznet15x5_train":
        return TrainWorkload(name, B, N, nfilt, dev, rank)
    return ModelWorkload(name, B, N, nfilt, dev, rank, pcm16)
