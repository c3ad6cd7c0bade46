"""Audio ingest on the device: what ``AudioFileLoader.preprocess_audio``
(src/thunder/data/dataset.py:50-77) does per file on the CPU -- mono mix, DC removal, resampling to the model's rate --
for a whole padded batch, straight from int16 PCM if that is what the caller has (half the host-to-device bytes of
float32 audio).  File opening / decoding (``torchaudio.load``) stays outside: it is host I/O, not on the path."""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
from torch import Tensor, nn

from . import _lib

__all__ = ["AudioFileLoader", "sinc_resample_taps", "pcm_ingest", "resample"]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def sinc_resample_taps(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99
                       ) -> Tuple[np.ndarray, int, int, int]:
    """torchaudio's ``_get_sinc_resample_kernel`` (sinc_interp_hann): (taps [new', ntaps] f32, width, orig', new')."""
    if int(orig_freq) != orig_freq or int(new_freq) != new_freq:
        raise Exception("Frequencies must be of integer type to ensure quality resampling computation.")
    if lowpass_filter_width <= 0:
        raise ValueError("Low pass filter width should be positive.")
    g = math.gcd(int(orig_freq), int(new_freq))
    o, n = int(orig_freq) // g, int(new_freq) // g
    base = min(o, n) * rolloff
    width = math.ceil(lowpass_filter_width * o / base)
    idx = np.arange(-width, width + o, dtype=np.float64)[None, :] / o
    t = np.clip((np.arange(0, -n, -1, dtype=np.float64)[:, None] / n + idx) * base, -lowpass_filter_width,
                lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    return (k * window * (base / o)).astype(np.float32), width, o, n


def pcm_ingest(pcm: Tensor, lens: Optional[Tensor] = None, interleaved: bool = False, remove_dc: bool = True,
               out: Optional[Tensor] = None) -> Tensor:
    """``pcm``: int16 or float32 CUDA tensor ``[B, channels, N]`` (or ``[B, N, channels]`` with ``interleaved``) ->
    float32 ``[B, N]``: mono mix, int16 scaled by 1/32768, per-utterance DC removed, zero beyond ``lens``.  ``out``: an
    existing contiguous float32 ``[B, N]`` buffer to write into (e.g. the input buffer of a captured graph)."""
    if not pcm.is_cuda:
        raise RuntimeError("thunder_b200 ingest runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
    if pcm.dim() != 3:
        raise ValueError("pcm must be [B, channels, N] (or [B, N, channels] with interleaved=True)")
    if pcm.dtype not in (torch.int16, torch.float32):
        raise TypeError("pcm must be int16 or float32")
    pcm = pcm.contiguous()
    B = pcm.shape[0]
    C, N = (pcm.shape[2], pcm.shape[1]) if interleaved else (pcm.shape[1], pcm.shape[2])
    l32 = lens.to(device=pcm.device, dtype=torch.int32).contiguous() if lens is not None else None
    if out is None:
        out = torch.empty((B, N), device=pcm.device, dtype=torch.float32)
    elif out.dtype != torch.float32 or tuple(out.shape) != (B, N) or not out.is_contiguous() or out.device != pcm.device:
        raise ValueError("pcm_ingest: out must be a contiguous float32 [B, N] tensor on the device of pcm")
    nscr = B * ((N + 65535) // 65536)
    scratch = torch.empty((nscr,), device=pcm.device, dtype=torch.float64)
    _lib.check(_lib.lib().ts_pcm_ingest(pcm.data_ptr(), _lib.TS_I16 if pcm.dtype == torch.int16 else _lib.TS_F32, B, C, N,
                                        l32.data_ptr() if l32 is not None else None, int(interleaved), int(remove_dc),
                                        out.data_ptr(), N, scratch.data_ptr(), nscr, _stream()), "ts_pcm_ingest")
    return out


_TAPS: Dict[Tuple[int, int, int], tuple] = {}


def compress_taps(k: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """taps [new', ntaps] -> (taps_c [nt, new'], k0 [new']): per phase only the run from the first to the last tap whose
    magnitude is above fp32 denormal noise is kept (the Hann window is clamped to exactly 0 outside +-6 zero crossings, so
    for 441 -> 160 only ~35 of 475 taps per phase matter)."""
    nz = np.abs(k) > 1e-30
    first = np.where(nz.any(1), nz.argmax(1), 0)
    last = np.where(nz.any(1), k.shape[1] - 1 - nz[:, ::-1].argmax(1), 0)
    nt = int((last - first + 1).max())
    out = np.zeros((nt, k.shape[0]), np.float32)
    for p in range(k.shape[0]):
        seg = k[p, first[p]: last[p] + 1]
        out[: seg.size, p] = seg
    return out, first.astype(np.int32)


def resample(x: Tensor, orig_freq: int, new_freq: int, lens: Optional[Tensor] = None) -> Tuple[Tensor, Optional[Tensor]]:
    """``torchaudio.functional.resample`` for a padded float32 batch ``[B, N]`` on the device; every utterance is resampled
    as if it were alone (``lens`` samples, zero beyond).  Returns ``(y [B, ceil(new' N / orig')], new lens)``."""
    if not x.is_cuda:
        raise RuntimeError("thunder_b200 resample runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
    if not x.is_floating_point():
        raise TypeError(f"Expected floating point type for waveform tensor, but received {x.dtype}.")
    if int(orig_freq) == int(new_freq):
        return x, lens
    dev = x.device
    key = (int(orig_freq), int(new_freq), dev.index or 0)
    if key not in _TAPS:
        k, width, o, n = sinc_resample_taps(orig_freq, new_freq)
        tc, k0 = compress_taps(k)
        _TAPS[key] = (torch.from_numpy(tc).to(dev), torch.from_numpy(k0).to(dev), width, o, n)
    taps_c, k0, width, o, n = _TAPS[key]
    x = x.float().contiguous()
    B, N = x.shape
    N_out = -((-n * N) // o)
    l32 = lens.to(device=dev, dtype=torch.int32).contiguous() if lens is not None else None
    y = torch.empty((B, N_out), device=dev, dtype=torch.float32)
    _lib.check(_lib.lib().ts_resample(x.data_ptr(), B, N, N, l32.data_ptr() if l32 is not None else None, o, n,
                                      taps_c.data_ptr(), k0.data_ptr(), taps_c.shape[0], width, y.data_ptr(), N_out, N_out,
                                      _stream()), "ts_resample")
    new_lens = None
    if lens is not None:
        new_lens = torch.div(lens.to(torch.int64) * n + (o - 1), o, rounding_mode="floor")
    return y, new_lens


class AudioFileLoader(nn.Module):
    """Same constructor and ``preprocess_audio(audio, sample_rate)`` contract as the reference
    (src/thunder/data/dataset.py:22-77) for ``audio [channels, time]`` on the GPU, plus ``preprocess_batch`` for a padded
    batch.  ``open_audio`` (file decoding) is host I/O and not provided."""

    def __init__(self, force_mono: bool = True, sample_rate: int = 16000):
        super().__init__()
        self.force_mono = force_mono
        self.sample_rate = sample_rate

    def open_audio(self, item: str):
        raise NotImplementedError("file decoding is host I/O outside the device path; pass decoded PCM to preprocess_audio")

    def preprocess_batch(self, pcm: Tensor, sample_rate: int, lens: Optional[Tensor] = None, interleaved: bool = False
                         ) -> Tuple[Tensor, Optional[Tensor]]:
        """``pcm [B, channels, N]`` int16 / float32 -> ``(audio [B, N'] float32 at self.sample_rate, lens')``."""
        C = pcm.shape[2] if interleaved else pcm.shape[1]
        if C > 1 and not self.force_mono:
            # the reference's `audio - audio.mean(1)` only broadcasts for a single channel (dataset.py:69)
            raise RuntimeError("The size of tensor a must match the size of tensor b: DC removal needs mono audio "
                               "(force_mono=False with multi-channel input fails in the reference too)")
        mono = pcm_ingest(pcm, lens, interleaved, remove_dc=True)
        if int(self.sample_rate) != int(sample_rate):
            return resample(mono, int(sample_rate), int(self.sample_rate), lens)
        return mono, lens

    def preprocess_audio(self, audio: Tensor, sample_rate: int) -> Tensor:
        """``audio [channels, time]`` -> ``[1, time']`` (the reference's per-file entry point)."""
        y, _ = self.preprocess_batch(audio.unsqueeze(0), sample_rate)
        return y

    def forward(self, item):
        raise NotImplementedError("file decoding is host I/O outside the device path; use preprocess_audio / preprocess_batch")
