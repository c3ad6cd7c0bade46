"""``torch.library`` custom ops (namespace ``thunder_b200``) over the C ABI.

Each op takes/returns torch tensors, allocates outputs with torch (device memory and streams are
torch's job -- plumbing), and passes raw device pointers plus the CURRENT CUDA stream to
``libthunder_b200.so``.  Ops are registered for CUDA only; calling them with CPU tensors raises --
there is no CPU fallback.  Fake (meta) implementations give shapes to ``torch.jit.script`` /
``torch.compile`` / ``to_torchscript``.
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor

from . import _lib

NS = "thunder_b200"


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*tensors: Tensor) -> None:
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("thunder_b200 ops run on CUDA (sm_100a) tensors only; there is no CPU fallback")


def _ptr(t: Tensor) -> int:
    return t.data_ptr()


# ------------------------------------------------------------------------------------------- features
@torch.library.custom_op(f"{NS}::filterbank", mutates_args=())
def filterbank(audio: Tensor, lengths: Tensor, window_full: Tensor, twiddle: Tensor, mel_start: Tensor,
               mel_count: Tensor, mel_off: Tensor, mel_w: Tensor, hop: int, preemph: float, win_lo: int,
               win_hi: int, div_guard: float, out_bf16_pitch: int) -> Tuple[Tensor, Tensor]:
    """Eval-mode ``FilterbankFeatures`` (src/thunder/quartznet/transform.py:258-321).

    ``out_bf16_pitch == 0``: returns ``(features[B,nfilt,F] f32, feature_lengths[B] i64)`` exactly like the
    reference.  ``out_bf16_pitch > 0``: features are emitted as bf16 padded rows ``[B,nfilt,pitch]`` for the
    encoder kernels."""
    _need_cuda(audio, lengths, window_full, twiddle, mel_start, mel_count, mel_off, mel_w)
    if audio.dim() != 2:
        raise ValueError("audio must be [batch, time]")
    audio = audio.contiguous().float()
    B, N = audio.shape
    n_fft = window_full.numel()
    nfilt = mel_start.numel()
    F = 1 + N // hop
    lens64 = lengths.to(torch.int64).contiguous()
    L = _lib.lib()
    logmel = torch.empty((B, nfilt, F), device=audio.device, dtype=torch.float32)
    _lib.check(L.ts_logmel(_ptr(audio), B, N, n_fft, hop, preemph, _ptr(window_full), win_lo, win_hi,
                           _ptr(twiddle), _ptr(mel_start), _ptr(mel_count), _ptr(mel_off), _ptr(mel_w), nfilt,
                           mel_w.numel(), _ptr(logmel), _stream()), "ts_logmel")
    seq = torch.empty((B,), device=audio.device, dtype=torch.int64)
    if out_bf16_pitch > 0:
        out = torch.empty((B, nfilt, out_bf16_pitch), device=audio.device, dtype=torch.bfloat16)
        dt, pitch = _lib.TS_BF16, out_bf16_pitch
    else:
        out = torch.empty((B, nfilt, F), device=audio.device, dtype=torch.float32)
        dt, pitch = _lib.TS_F32, F
    _lib.check(L.ts_feature_normalize(_ptr(logmel), _ptr(lens64), B, nfilt, F, hop, div_guard, _ptr(out), dt,
                                      pitch, _ptr(seq), _stream()), "ts_feature_normalize")
    return out, seq


@filterbank.register_fake
def _(audio, lengths, window_full, twiddle, mel_start, mel_count, mel_off, mel_w, hop, preemph, win_lo, win_hi,
      div_guard, out_bf16_pitch):
    B, N = audio.shape
    nfilt = mel_start.numel()
    F = 1 + N // hop
    if out_bf16_pitch > 0:
        out = audio.new_empty((B, nfilt, out_bf16_pitch), dtype=torch.bfloat16)
    else:
        out = audio.new_empty((B, nfilt, F), dtype=torch.float32)
    return out, audio.new_empty((B,), dtype=torch.int64)
