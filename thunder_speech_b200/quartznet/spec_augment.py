"""SpecAugment / SpecCutout on the device (mirror of src/thunder/quartznet/spec_augment.py:23-110).

Training-time only, identity in eval().  The reference draws every mask interval with two ``torch.rand(1)`` calls on the
HOST generator (torchaudio ``mask_along_axis``: ``v = rand*param``, ``v0 = rand*(size - v)``, mask ``[long(v0), long(v0) +
long(v))``, the same interval for every utterance).  Eager mode here makes the same draws in the same order, so a seeded run
masks exactly what the reference masks; while a CUDA graph is being captured the draws come from the device generator
instead (graph-safe: every replay gets fresh intervals).  The masking itself is one launch of ``ts_spec_mask`` for all
rectangles, in place on fp32 ``[B, C, T]`` features or on bf16 rows."""
from __future__ import annotations

from typing import List, Tuple

import torch
from torch import Tensor, nn

from .. import _lib

__all__ = ["SpecAugment", "SpecCutout", "apply_rects"]


def _host_interval(size: int, param: int) -> Tuple[int, int]:
    """torchaudio.functional.mask_along_axis / spec_augment._create_mask: two host draws -> [start, end)."""
    value = torch.rand(1) * param
    min_value = torch.rand(1) * (size - value)
    start = int(min_value.long())
    return start, start + int(value.long())


def _device_interval(u: Tensor, size: int, param: int) -> Tuple[Tensor, Tensor]:
    """The same formula on two draws ``u[0], u[1]`` of the CUDA generator; only scalar-constant kernels, so it can be
    captured in a CUDA graph (no host-to-device copy).  Returns 0-dim int32 tensors (start, end)."""
    value = u[0] * float(param)
    start = (u[1] * (float(size) - value)).long()
    return start.to(torch.int32), (start + value.long()).to(torch.int32)


def apply_rects(feat: Tensor, T: int, rects: Tensor) -> Tensor:
    """Zero ``feat[:, f0:f1, t0:t1]`` in place for every row ``(f0, f1, t0, t1)`` of the int32 device tensor ``rects``."""
    if not feat.is_cuda:
        raise RuntimeError("thunder_b200 SpecAugment runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
    if feat.dtype not in (torch.float32, torch.bfloat16) or not feat.is_contiguous():
        raise TypeError("features must be contiguous float32 [B, C, T] or bf16 rows [B, C, pitch]")
    B, C, pitch = feat.shape
    rects = rects.to(device=feat.device, dtype=torch.int32).contiguous()
    _lib.check(_lib.lib().ts_spec_mask(feat.data_ptr(), _lib.TS_F32 if feat.dtype == torch.float32 else _lib.TS_BF16, B, C, T,
                                       pitch, rects.data_ptr(), rects.shape[0], torch.cuda.current_stream().cuda_stream),
               "ts_spec_mask")
    return feat


class _MaskBase(nn.Module):
    def _plan(self, C: int, T: int) -> List[Tuple[str, int, int]]:
        """[(axis, size, param)] in the reference's draw order; axis 't' / 'f' / 'rect' (two intervals)."""
        raise NotImplementedError

    def rects(self, C: int, T: int, device) -> Tensor:
        plan = self._plan(C, T)
        if not plan:
            return torch.zeros((0, 4), dtype=torch.int32, device=device)
        if torch.cuda.is_current_stream_capturing():
            u = torch.rand((len(plan), 2), device=device)
            lo = torch.zeros((), device=device, dtype=torch.int32)
            full_f = (lo, torch.full((), C, device=device, dtype=torch.int32))
            full_t = (lo, torch.full((), T, device=device, dtype=torch.int32))
            rows, i = [], 0
            while i < len(plan):
                kind, size, param = plan[i]
                iv = _device_interval(u[i], size, param)
                if kind == "t":
                    rows.append(torch.stack(full_f + iv))
                    i += 1
                elif kind == "f":
                    rows.append(torch.stack(iv + full_t))
                    i += 1
                else:   # cutout: a frequency interval then a time interval
                    rows.append(torch.stack(iv + _device_interval(u[i + 1], plan[i + 1][1], plan[i + 1][2])))
                    i += 2
            return torch.stack(rows)
        rows, i = [], 0
        while i < len(plan):
            kind, size, param = plan[i]
            if kind == "t":
                rows.append((0, C) + _host_interval(size, param))
                i += 1
            elif kind == "f":
                rows.append(_host_interval(size, param) + (0, T))
                i += 1
            else:
                f = _host_interval(size, param)
                t = _host_interval(plan[i + 1][1], plan[i + 1][2])
                rows.append(f + t)
                i += 2
        return torch.tensor(rows, dtype=torch.int32).to(device)

    def forward(self, x: Tensor, T: int = None) -> Tensor:
        """``x``: fp32 ``[B, C, T]`` (a fresh tensor is returned, like the reference's masked_fill) or, with ``T`` given,
        bf16 rows ``[B, C, pitch]`` (masked in place)."""
        if not self.training:
            return x
        if T is None:
            x = x.contiguous().clone()
            T = x.shape[-1]
        with torch.no_grad():
            return apply_rects(x, T, self.rects(x.shape[1], T, x.device))


class SpecAugment(_MaskBase):
    """Time masks first, then frequency masks (spec_augment.py:23-62).  torchaudio skips a mask when ``param < 1``."""

    def __init__(self, freq_masks=0, time_masks=0, freq_width=10, time_width=10):
        super().__init__()
        self.freq_masks, self.time_masks = freq_masks, time_masks
        self.freq_width, self.time_width = freq_width, time_width

    def _plan(self, C, T):
        plan = []
        if self.time_width >= 1:
            plan += [("t", T, self.time_width)] * self.time_masks
        if self.freq_width >= 1:
            plan += [("f", C, self.freq_width)] * self.freq_masks
        return plan


class SpecCutout(_MaskBase):
    """Random rectangles (spec_augment.py:78-110).  Like the reference, BOTH sides of a rectangle are drawn with
    ``freq_width`` -- ``time_width`` is stored but unused there (spec_augment.py:106-107)."""

    def __init__(self, rect_masks: int = 0, time_width: int = 5, freq_width: int = 20):
        super().__init__()
        self.rect_masks, self.time_width, self.freq_width = rect_masks, time_width, freq_width

    def _plan(self, C, T):
        plan = []
        for _ in range(self.rect_masks):
            plan += [("rect", C, self.freq_width), ("rect2", T, self.freq_width)]
        return plan
