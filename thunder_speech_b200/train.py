"""Training step on the B200 kernels: train()-mode forward, backward and parameter gradients of
the QuartzNet encoder stack, mirroring what ``BaseCTCModule.training_step`` (src/thunder/module.py:102-127) does through
torch autograd.

Scope of this first version: separable QuartzNet blocks with stride-1 body (every block of QuartzNet 5x5 / 15x5), the
strided stem (no input gradient needed: the feature front-end is under ``no_grad``, transform.py:87-186) and the final
1x1 block.  The decoder (1024 -> V) and ``calculate_ctc`` (log_softmax + ``F.ctc_loss``, src/thunder/ctc_loss.py:15-47) run
through torch autograd on the encoder output ("use torch's ctc_loss as-is first"); the optimiser is
torch's AdamW (module.py:32); gradients are averaged across ranks with one flat NCCL all-reduce.

Everything heavy is a kernel of libthunder_b200.so:
    forward   ts_dw_conv (Toeplitz MMA) -> ts_pw_gemm (unfolded bf16 weights) -> ts_row_stats -> ts_bn_apply
    backward  ts_bn_bwd_reduce -> ts_bn_bwd_apply -> ts_pw_wgrad (tcgen05, K = batch x time) + ts_pw_gemm (W^T: dgrad)
              -> ts_dw_wgrad + ts_dw_conv (flipped taps: dgrad)
The only torch arithmetic is on per-channel ``[C]`` vectors (statistics -> scale/shift/coefficients) and the final sums
over batch slices of the deterministic partial reductions.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor, nn

from . import _lib, ops
from .blocks import conv_out_length
from .parallel import allreduce_mean_, ensure_grad_views, flat_grad_views

BN_EPS = 1e-3
BN_MOMENTUM = 0.1


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[Tensor]):
    return t.data_ptr() if t is not None else None


# ------------------------------------------------------------------------------------------------ thin kernel wrappers
def row_stats(z: Tensor, T: int) -> Tensor:
    """[C, 2] = (sum z, sum z^2) over batch and t < T."""
    B, C, pitch = z.shape
    st = torch.empty((B, C, 2), device=z.device, dtype=torch.float32)
    with ops._timed("row_stats", bytes=B * C * T * 2, flops=0):
        _lib.check(_lib.lib().ts_row_stats(_p(z), B, C, T, pitch, _p(st), _stream()), "ts_row_stats")
    return st.sum(0, dtype=torch.float64)


def row_stats_partial(z: Tensor, T: int) -> Tensor:
    """[B, C, 2] per-utterance (sum z, sum z^2) over t < T (summed over B by ts_bn_finalize)."""
    B, C, pitch = z.shape
    st = torch.empty((B, C, 2), device=z.device, dtype=torch.float32)
    with ops._timed("row_stats", bytes=B * C * T * 2, flops=0):
        _lib.check(_lib.lib().ts_row_stats(_p(z), B, C, T, pitch, _p(st), _stream()), "ts_row_stats")
    return st


def pw_gemm_stats(w: Tensor, x: Tensor, T: int) -> Tuple[Tensor, Tensor]:
    """z = W x (bf16 rows) and the BatchNorm partial sums of z, [B, Cout, slots, 2].  Cout > 128: the statistics come out
    of the GEMM epilogue (ts_pw_gemm_stats); otherwise a plain GEMM followed by ts_row_stats (slots = 1)."""
    B, cin, px = x.shape
    Cout = w.shape[0]
    if Cout > 128:
        pitch = ops.row_pitch(T)
        slots = 2 * ((pitch + 255) // 256)
        z = torch.empty((B, Cout, pitch), device=x.device, dtype=torch.bfloat16)
        st = torch.empty((B, Cout, slots, 2), device=x.device, dtype=torch.float32)
        with ops._timed("pw_gemm", bytes=2 * B * T * (cin + Cout) + 2 * Cout * cin, flops=2 * B * T * cin * Cout, K=cin,
                        C=Cout, T=T):
            rc = _lib.lib().ts_pw_gemm_stats(_p(w), _p(x), cin, px, B, Cout, T, _p(z), pitch, _p(st), slots, _stream())
        if rc != _lib.TS_ERR_UNSUPPORTED:
            _lib.check(rc, "ts_pw_gemm_stats")
            return z, st
    z = ops.pw_gemm(w, x, None, None, T, None, None, False, False, None, None, None)
    return z, row_stats_partial(z, T).unsqueeze(2)


def bn_finalize(part: Tensor, n: int, bn: nn.BatchNorm1d, update_running: bool):
    """(scale, shift, mean, inv) f32 [C] from the partial sums [NB, C, 2] or [NB, C, slots, 2], on the device; updates the
    running statistics in place like nn.BatchNorm1d(momentum) does in train() (unbiased variance for the running estimate)."""
    NB, C = part.shape[0], part.shape[1]
    slots = part.shape[2] if part.dim() == 4 else 1
    out = torch.empty((4, C), device=part.device, dtype=torch.float32)
    upd = update_running and bn.track_running_stats
    m = bn.momentum if bn.momentum is not None else BN_MOMENTUM
    _lib.check(_lib.lib().ts_bn_finalize(_p(part), NB, slots, C, float(n), _p(bn.weight), _p(bn.bias), float(bn.eps), float(m),
                                         _p(bn.running_mean) if upd else None, _p(bn.running_var) if upd else None,
                                         _p(out[0]), _p(out[1]), _p(out[2]), _p(out[3]), _stream()), "ts_bn_finalize")
    return out[0], out[1], out[2], out[3]


def bn_bwd_coef(part: Tensor, which: int, n: int, bn: nn.BatchNorm1d, mean: Tensor, inv: Tensor) -> Tensor:
    """coef [C, 3] of dz = a dym + b z + c; writes dgamma / dbeta straight into ``bn.weight.grad`` / ``bn.bias.grad``."""
    NB, C, _ = part.shape
    coef = torch.empty((C, 3), device=part.device, dtype=torch.float32)
    _lib.check(_lib.lib().ts_bn_bwd_coef(_p(part), NB, C, which, float(n), _p(bn.weight), _p(mean), _p(inv),
                                         _p(_grad(bn.weight)), _p(_grad(bn.bias)), _p(coef), _stream()), "ts_bn_bwd_coef")
    return coef


def ctc_loss(logits: Tensor, T: int, in_len: Tensor, targets: Tensor, tgt_len: Tensor, blank: int, Vp: int = 64,
             gscale: float = 1.0) -> Tuple[Tensor, Tensor]:
    """(loss [B] = nll_b / max(target_len_b, 1), zero when infinite;  d mean(loss) / d logits as bf16 rows [B, Vp, pitch])
    from f32 logit rows [B, V, pitch]: log_softmax + F.ctc_loss(reduction="mean", zero_infinity=True) and its backward."""
    B, V, pitch = logits.shape
    gp = ops.row_pitch(T)
    assert logits.dtype == torch.float32 and targets.dtype == torch.int64 and tgt_len.dtype == torch.int64
    assert in_len.dtype == torch.int32
    Lmax = targets.shape[1]
    Sp = (2 * Lmax + 1 + 3) // 4 * 4
    nfl = 2 * B * T * Sp + B * T + B
    scratch = torch.empty((nfl,), device=logits.device, dtype=torch.float32)
    loss = torch.empty((B,), device=logits.device, dtype=torch.float32)
    grad = torch.zeros((B, max(Vp, V), gp), device=logits.device, dtype=torch.bfloat16)
    with ops._timed("ctc_loss", bytes=B * V * T * 4 + 2 * 2 * B * T * Sp * 4, flops=0):
        _lib.check(_lib.lib().ts_ctc_loss(_p(logits), B, V, T, pitch, _p(in_len), _p(targets.contiguous()), Lmax,
                                          _p(tgt_len), int(blank), float(gscale), _p(scratch), nfl, _p(loss), _p(grad),
                                          grad.shape[1], gp, _stream()), "ts_ctc_loss")
    return loss, grad


def bn_apply_fused(z: Tensor, part: Tensor, bn: nn.BatchNorm1d, zr: Optional[Tensor], part_r: Optional[Tensor],
                   rbn: Optional[nn.BatchNorm1d], T: int, lens: Optional[Tensor], relu: bool, update_running: bool):
    """y = act(BN(z) [+ BN_r(zr)]) with batch statistics taken from the partial sums ``part`` [NB, C, slots, 2]; returns
    (y, stats [4, C], stats_r): stats = (scale, shift, mean, inv) for the backward pass.  Running statistics are updated in
    place like nn.BatchNorm1d(momentum) does in train() (unbiased variance for the running estimate)."""
    B, C, pitch = z.shape
    y = torch.empty_like(z)
    stats = torch.empty((4, C), device=z.device, dtype=torch.float32)
    stats_r = torch.empty((4, C), device=z.device, dtype=torch.float32) if zr is not None else None

    def run_ptrs(b):
        upd = update_running and b.track_running_stats
        return (_p(b.running_mean) if upd else None, _p(b.running_var) if upd else None)

    rm, rv = run_ptrs(bn)
    rmr, rvr = run_ptrs(rbn) if rbn is not None else (None, None)
    mom = bn.momentum if bn.momentum is not None else BN_MOMENTUM
    with ops._timed("bn_apply", bytes=B * C * T * 2 * (3 if zr is not None else 2), flops=0):
        _lib.check(_lib.lib().ts_bn_apply_fused(
            _p(z), _p(part), part.shape[0], part.shape[2], _p(bn.weight), _p(bn.bias), rm, rv, _p(stats),
            _p(zr), _p(part_r), part_r.shape[0] if part_r is not None else 0, part_r.shape[2] if part_r is not None else 0,
            _p(rbn.weight) if rbn is not None else None, _p(rbn.bias) if rbn is not None else None, rmr, rvr, _p(stats_r),
            float(bn.eps), float(mom), B, C, T, pitch, _p(lens), int(relu), _p(y), _stream()), "ts_bn_apply_fused")
    return y, stats, stats_r


def bn_bwd_apply_fused(dy: Tensor, z: Tensor, bn: nn.BatchNorm1d, stats: Tensor, zr: Optional[Tensor],
                       rbn: Optional[nn.BatchNorm1d], stats_r: Optional[Tensor], sums: Tensor, T: int, relu: bool = True):
    """(dz, dzr) of BatchNorm + ReLU from the partial sums of bn_bwd_reduce; dgamma / dbeta go straight into the
    parameters' gradient buffers."""
    B, C, pitch = z.shape
    dz = torch.empty_like(z)
    dzr = torch.empty_like(z) if zr is not None else None
    with ops._timed("bn_bwd_apply", bytes=B * C * T * 2 * (6 if zr is not None else 4), flops=0):
        _lib.check(_lib.lib().ts_bn_bwd_apply_fused(
            _p(dy), _p(z), _p(bn.weight), _p(stats), _p(_grad(bn.weight)), _p(_grad(bn.bias)), _p(dz), _p(zr),
            _p(rbn.weight) if rbn is not None else None, _p(stats_r),
            _p(_grad(rbn.weight)) if rbn is not None else None, _p(_grad(rbn.bias)) if rbn is not None else None, _p(dzr),
            _p(sums), B, C, T, pitch, int(relu), _stream()), "ts_bn_bwd_apply_fused")
    return dz, dzr


def scatter_rows(src: Tensor, T_src: int, S: int, T_dst: int, dst: Optional[Tensor] = None) -> Tensor:
    """Transpose of ``ops.gather_rows``: ``dst[:, :, t*S] (+)= src[:, :, t]``; a fresh zero-filled tensor unless ``dst`` is
    given (then accumulated in place)."""
    B, C, ps = src.shape
    acc = dst is not None
    if dst is None:
        dst = torch.empty((B, C, ops.row_pitch(T_dst)), device=src.device, dtype=torch.bfloat16)
    _lib.check(_lib.lib().ts_scatter_rows(_p(src), B, C, T_src, ps, S, _p(dst), T_dst, dst.shape[2], int(acc), _stream()),
               "ts_scatter_rows")
    return dst


def bn_apply_se(z, scale, shift, zr, scale_r, shift_r, gate, T, lens, relu=True) -> Tensor:
    """y = act(gate[b,c] * BN(z) [+ BN_r(zr)]): the last sub-block of a Citrinet block (SqueezeExcite scale)."""
    B, C, pitch = z.shape
    y = torch.empty_like(z)
    with ops._timed("bn_apply", bytes=B * C * T * 2 * (3 if zr is not None else 2), flops=0):
        _lib.check(_lib.lib().ts_bn_apply_se(_p(z), _p(scale), _p(shift), _p(zr), _p(scale_r), _p(shift_r), _p(gate), B, C, T,
                                             pitch, _p(lens), int(relu), _p(y), _stream()), "ts_bn_apply_se")
    return y


def bn_bwd_reduce_se(dy, z, zr, T, mask, gate, relu=True) -> Tensor:
    """Per-utterance partial sums [B, C, 3] = (sum dym, sum dym*z, sum dym*zr), the ReLU mask rebuilt with the SE gate."""
    B, C, pitch = z.shape
    sums = torch.empty((B, C, 3), device=z.device, dtype=torch.float32)
    with ops._timed("bn_bwd_reduce", bytes=B * C * T * 2 * (3 if zr is not None else 2), flops=0):
        _lib.check(_lib.lib().ts_bn_bwd_reduce_se(_p(dy), _p(z), _p(zr), B, C, T, pitch, int(relu), _p(sums), _p(mask[0]),
                                                  _p(mask[1]), _p(mask[2]), _p(mask[3]), _p(gate), _stream()),
                   "ts_bn_bwd_reduce_se")
    return sums


def bn_bwd_apply_se(dy, z, zr, coef, coef_r, T, mask, gate, addc, relu=True) -> Tuple[Tensor, Optional[Tensor]]:
    """dz = a (dym gate + addc) + b z + c0;  dzr = a_r dym + b_r zr + c0_r."""
    B, C, pitch = z.shape
    dz = torch.empty_like(z)
    dzr = torch.empty_like(z) if zr is not None else None
    with ops._timed("bn_bwd_apply", bytes=B * C * T * 2 * (6 if zr is not None else 4), flops=0):
        _lib.check(_lib.lib().ts_bn_bwd_apply_se(_p(dy), _p(z), _p(zr), _p(coef), _p(coef_r), B, C, T, pitch, int(relu),
                                                 _p(dz), _p(dzr), _p(mask[0]), _p(mask[1]), _p(mask[2]), _p(mask[3]),
                                                 _p(gate), _p(addc), _stream()), "ts_bn_bwd_apply_se")
    return dz, dzr


def bn_apply(z, scale, shift, zr, scale_r, shift_r, T, lens, relu=True) -> Tensor:
    B, C, pitch = z.shape
    y = torch.empty_like(z)
    with ops._timed("bn_apply", bytes=B * C * T * 2 * (3 if zr is not None else 2), flops=0):
        _lib.check(_lib.lib().ts_bn_apply(_p(z), _p(scale), _p(shift), _p(zr), _p(scale_r), _p(shift_r), B, C, T, pitch,
                                          _p(lens), int(relu), _p(y), _stream()), "ts_bn_apply")
    return y


def bn_bwd_reduce(dy, y, z, zr, T, relu=True, partial: bool = False, mask=None) -> Tensor:
    """``mask`` = (scale, shift, scale_r, shift_r) of the forward bn_apply: with ``y=None`` the ReLU mask is rebuilt from
    z (and zr) instead of being read from y."""
    B, C, pitch = z.shape
    ms = mask if mask is not None else (None, None, None, None)
    sums = torch.empty((B, C, 3), device=z.device, dtype=torch.float32)
    with ops._timed("bn_bwd_reduce", bytes=B * C * T * 2 * (4 if zr is not None else 3), flops=0):
        _lib.check(_lib.lib().ts_bn_bwd_reduce(_p(dy), _p(y), _p(z), _p(zr), B, C, T, pitch, int(relu), _p(sums),
                                               _p(ms[0]), _p(ms[1]), _p(ms[2]), _p(ms[3]), _stream()), "ts_bn_bwd_reduce")
    return sums if partial else sums.sum(0, dtype=torch.float64)


def bn_bwd_apply(dy, y, z, zr, coef, coef_r, T, relu=True, mask=None) -> Tuple[Tensor, Optional[Tensor]]:
    B, C, pitch = z.shape
    ms = mask if mask is not None else (None, None, None, None)
    dz = torch.empty_like(z)
    dzr = torch.empty_like(z) if zr is not None else None
    with ops._timed("bn_bwd_apply", bytes=B * C * T * 2 * (6 if zr is not None else 4), flops=0):
        _lib.check(_lib.lib().ts_bn_bwd_apply(_p(dy), _p(y), _p(z), _p(zr), _p(coef), _p(coef_r), B, C, T, pitch,
                                              int(relu), _p(dz), _p(dzr), _p(ms[0]), _p(ms[1]), _p(ms[2]), _p(ms[3]),
                                              _stream()), "ts_bn_bwd_apply")
    return dz, dzr


def pw_wgrad(dz: Tensor, a: Tensor, T: int, out: Optional[Tensor] = None) -> Tensor:
    """dW[co, ci] = sum_{b,t} dz[b,co,t] a[b,ci,t] (fp32); written into ``out`` (any shape with Cout*Cin elements)."""
    B, Cout, pz = dz.shape
    Cin, pa = a.shape[1], a.shape[2]
    tiles = ((Cout + 127) // 128) * ((Cin + 255) // 256)
    nsplit = max(1, min(B * ((T + 63) // 64), 148 // max(tiles, 1)))   # fill the 148 SMs
    if nsplit == 1 and out is not None and out.is_contiguous():
        part = out.view(1, Cout, Cin)
    else:
        part = torch.empty((nsplit, Cout, Cin), device=dz.device, dtype=torch.float32)
    with ops._timed("pw_wgrad", bytes=B * (Cout + Cin) * T * 2 + nsplit * Cout * Cin * 4, flops=2 * B * T * Cout * Cin):
        _lib.check(_lib.lib().ts_pw_wgrad(_p(dz), pz, _p(a), pa, B, Cout, Cin, T, nsplit, _p(part), _stream()),
                   "ts_pw_wgrad")
    if out is None:
        out = torch.empty((Cout, Cin), device=dz.device, dtype=torch.float32)
    if part.data_ptr() != out.data_ptr():
        n = Cout * Cin
        if n % 4 == 0 and out.is_contiguous() and out.data_ptr() % 16 == 0:
            _lib.check(_lib.lib().ts_pw_wgrad_reduce(_p(part), nsplit, n, _p(out), _stream()), "ts_pw_wgrad_reduce")
        else:
            torch.sum(part, 0, out=out.view(Cout, Cin))
    return out


def dw_wgrad(da: Tensor, T_out: int, x: Tensor, T_in: int, lens_in: Optional[Tensor], K: int, S: int, D: int, P: int,
             out: Optional[Tensor] = None, premasked: bool = False) -> Tensor:
    """dw[c, k] = sum_{b,t} da[b,c,t] x[b,c,t*S + k*D - P] (x masked to lens_in).  ``premasked``: both row tensors are
    already zero beyond the utterance lengths and in the pitch pad -> stride-1 layers run on the tensor cores."""
    B, C, po = da.shape
    pi = x.shape[2]
    mma = premasked and S == 1 and T_in == T_out and pi == po and K <= 128
    bchunk = B if mma else min(B, 8)   # SIMT kernels: 8 utterances per CTA amortise the staging / reduction phases
    nchunk = (B + bchunk - 1) // bchunk
    if nchunk == 1 and out is not None and out.is_contiguous():
        part = out.view(1, C, K)
    else:
        part = torch.empty((nchunk, C, K), device=da.device, dtype=torch.float32)
    with ops._timed("dw_wgrad", bytes=B * C * (T_out + T_in) * 2, flops=2 * B * C * T_out * K):
        _lib.check(_lib.lib().ts_dw_wgrad(_p(da), T_out, po, _p(x), T_in, pi, _p(lens_in), B, C, K, S, D, P, bchunk,
                                          _lib.TS_DW_INPUT_PREMASKED if premasked else 0, _p(part), _stream()),
                   "ts_dw_wgrad")
    if out is None:
        return part.sum(0)
    if part.data_ptr() != out.data_ptr():
        torch.sum(part, 0, out=out.view(C, K))
    return out


# ------------------------------------------------------------------------------------------------ block trainer
@dataclass
class _Sub:
    dw: Optional[nn.Conv1d]
    pw: nn.Conv1d
    bn: nn.BatchNorm1d
    K: int
    S: int
    D: int
    P: int


def _grad(p: nn.Parameter) -> Tensor:
    """The gradient buffer of `p` (created on first use).  Every parameter of the encoder receives exactly one contribution
    per step, so the backward kernels WRITE into these buffers (no zeroing pass, no accumulation); CTCTrainStep points them
    at slices of one flat buffer so that the all-reduce needs no packing."""
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    return p.grad


class WeightPack:
    """The compute-precision copies of the fp32 master weights, rebuilt by ONE kernel launch per step (ts_prep_weights):

      pointwise / decoder weight [Cout, Cin, 1]  ->  bf16 ``w`` [Cout, Cin] (forward GEMM) and bf16 ``wT`` [Cin, ld]
                                                     (input-gradient GEMM; ld = Cout rounded up to 64 for the decoder)
      depthwise taps [C, 1, K]                   ->  f32 ``w`` [C, K] rounded to bf16 (the Toeplitz tensor-core path holds
                                                     its taps in bf16; every path sees the same values) and ``wT`` = the
                                                     taps flipped along K (the input-gradient convolution)

    The buffers and the device-side table are allocated once; parameters are updated in place by the optimiser, so the
    table stays valid."""

    def __init__(self, entries: List[Tuple[str, nn.Parameter]]):
        self.entries = list(entries)
        self._build()

    def _build(self) -> None:
        entries = self.entries
        self.views: Dict[int, Tuple[Tensor, Tensor]] = {}
        rows_tab = []
        tiles = 0
        self._keep = []
        for kind, p in entries:
            dev = p.device
            if kind == "pw":
                rows, cols = p.shape[0], p.shape[1]
                # K of the input-gradient GEMM = channels of the dz rows; only a row pitch that is not 16-byte aligned
                # (the decoder's V = 29) is padded, and then the gradient rows are padded to the same count
                ld = rows if rows % 8 == 0 else (rows + 63) // 64 * 64
                w = torch.empty((rows, cols), device=dev, dtype=torch.bfloat16)
                wT = torch.zeros((cols, ld), device=dev, dtype=torch.bfloat16)
                k = 0
            else:
                rows, cols = p.shape[0], p.shape[2]
                ld = cols
                w = torch.empty((rows, cols), device=dev, dtype=torch.float32)
                wT = torch.empty((rows, cols), device=dev, dtype=torch.float32)
                k = 1
            if not p.is_contiguous():
                raise ValueError("WeightPack: parameters must be contiguous")
            rows_tab.append([p.data_ptr(), w.data_ptr(), wT.data_ptr(), rows, cols, ld, k, tiles])
            tiles += ((rows + 31) // 32) * ((cols + 31) // 32)
            self.views[id(p)] = (w, wT)
        self.n, self.tiles = len(rows_tab), tiles
        self.table = torch.tensor(rows_tab, dtype=torch.int64).to(entries[0][1].device)
        self._src_ptrs = [p.data_ptr() for _, p in entries]

    def refresh(self) -> None:
        # the table holds raw device pointers: rebuild it if a parameter's storage moved (module.to(...), re-assignment)
        if any(p.data_ptr() != q for (_, p), q in zip(self.entries, self._src_ptrs)):
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("WeightPack: a parameter was re-allocated while a CUDA graph is being captured")
            self._build()
        _lib.check(_lib.lib().ts_prep_weights(_p(self.table), self.n, self.tiles, _stream()), "ts_prep_weights")

    def get(self, p: nn.Parameter) -> Tuple[Tensor, Tensor]:
        return self.views[id(p)]


class _SideStream:
    """Weight gradients are off the critical path of the backward pass (which is the dx chain: BN backward -> W^T GEMM ->
    flipped-tap depthwise conv).  They are issued on a second stream, forked and joined with events, so that inside the
    captured CUDA graph they become parallel branches that fill the SMs left idle by the small per-layer kernels.
    Tensors handed to the side stream are kept alive until the join (no allocator reuse while a side kernel is pending)."""

    _streams: Dict[int, "torch.cuda.Stream"] = {}

    def __init__(self, device: torch.device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if idx not in _SideStream._streams:
            _SideStream._streams[idx] = torch.cuda.Stream(device=idx)
        self.stream = _SideStream._streams[idx]
        self.keep: List[Tensor] = []
        self.dirty = False

    def run(self, fn, *tensors: Tensor) -> "torch.cuda.Event":
        """Runs ``fn`` on the side stream after everything queued so far on the current stream; returns the event that marks
        its completion (wait on it before consuming results on the main stream, or call ``join``)."""
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        self.keep.extend(t for t in tensors if t is not None)
        done = torch.cuda.Event()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            fn()
            done.record(self.stream)
        self.dirty = True
        return done

    def join(self) -> None:
        if self.dirty:
            torch.cuda.current_stream().wait_stream(self.stream)
            self.dirty = False
        self.keep.clear()


class BlockTrainer:
    """train()-mode forward / backward of one QuartznetBlock on bf16 rows."""

    def __init__(self, block: nn.Module):
        from .quartznet.blocks import MaskedConv1d

        self.block = block
        self.subs: List[_Sub] = []
        self.se: Optional[nn.Module] = None
        self.res_stride = 1
        pending = []
        for layer in block.mconv.children():
            if isinstance(layer, MaskedConv1d):
                pending.append(layer)
                continue
            inner = layer.layer[0]
            if isinstance(inner, nn.BatchNorm1d):
                if block.separable:
                    dwl, pwl = pending
                    self.subs.append(_Sub(dwl.conv, pwl.conv, inner, dwl.kernel_size, dwl.stride, dwl.dilation, dwl.padding))
                else:
                    (pwl,) = pending
                    if pwl.kernel_size != 1 or pwl.stride != 1:
                        raise NotImplementedError("training: non-separable convs only for kernel_size=1")
                    self.subs.append(_Sub(None, pwl.conv, inner, 1, 1, 1, 0))
                pending = []
            elif hasattr(inner, "fc"):
                self.se = inner          # SqueezeExcite after the last BatchNorm (citrinet/blocks.py:154)
        self.res = None
        if block.res is not None:
            rl = list(block.res.children())
            self.res = (rl[0].conv, rl[1].layer[0])
            self.res_stride = int(rl[0].stride)
        self.pack: Optional[WeightPack] = None      # set by EncoderTrainer (one pack for the whole model)
        self._own_pack = False

    def pack_entries(self) -> List[Tuple[str, nn.Parameter]]:
        ent = []
        for sb in self.subs:
            if sb.dw is not None:
                ent.append(("dw", sb.dw.weight))
            ent.append(("pw", sb.pw.weight))
        if self.res is not None:
            ent.append(("pw", self.res[0].weight))
        return ent

    # -- forward ---------------------------------------------------------------------------------------
    def _nbt_buffers(self) -> List[Tensor]:
        nbt = [sb.bn.num_batches_tracked for sb in self.subs if sb.bn.track_running_stats]
        if self.res is not None and self.res[1].track_running_stats:
            nbt.append(self.res[1].num_batches_tracked)
        return nbt

    def forward(self, x: Tensor, T: int, lens: Optional[Tensor], zero_tail: bool, update_running: bool = True,
                bump_nbt: bool = True, side: Optional[_SideStream] = None):
        """``bump_nbt=False``: the caller increments ``num_batches_tracked`` for all blocks at once (EncoderTrainer).
        ``side``: run the residual branch's GEMM (it only depends on the block input) on this forked stream, in parallel with
        the main chain of sub-blocks; joined before the last BatchNorm apply."""
        if self.pack is None:            # stand-alone use (block-level tests): own pack, refreshed every forward
            self.pack, self._own_pack = WeightPack(self.pack_entries()), True
        if self._own_pack:
            self.pack.refresh()
        tape = dict(x=x, T=T, lens=lens, subs=[])
        # input of the residual 1x1 conv: every res_stride-th frame (citrinet/blocks.py:140-159)
        xr_in, Tr = x, T
        if self.res is not None and self.res_stride != 1:
            xr_in = ops.gather_rows(x, T, self.res_stride, lens)
            Tr = (T - 1) // self.res_stride + 1
        tape.update(xr=xr_in, Tr=Tr)
        res_out, res_done = {}, None
        if self.res is not None and side is not None:
            def _res_branch():
                res_out["z"], res_out["st"] = pw_gemm_stats(self.pack.get(self.res[0].weight)[0], xr_in, Tr)
            res_done = side.run(_res_branch, xr_in)

        def residual_gemm():
            """(zr, partial stats) of the residual 1x1 conv: from the side stream if it was forked, else computed here."""
            if res_done is not None:
                torch.cuda.current_stream().wait_event(res_done)
                return res_out["z"], res_out["st"]
            return pw_gemm_stats(self.pack.get(self.res[0].weight)[0], xr_in, Tr)

        cur, Tc, lc = x, T, lens
        B = x.shape[0]
        n = len(self.subs)
        y = None
        for r, sb in enumerate(self.subs):
            last = r == n - 1
            if sb.dw is not None:
                a = ops.dw_conv(cur, Tc, self.pack.get(sb.dw.weight)[0], sb.S, sb.D, sb.P, lc, True)
                Ta = conv_out_length(Tc, sb.K, sb.S, sb.P, sb.D)
                la = lc if (lc is None or (sb.S == 1 and 2 * sb.P == sb.D * (sb.K - 1))) else ops.conv_lengths(
                    lc, sb.K, sb.S, sb.D, sb.P)
            else:
                a, Ta, la = cur, Tc, lc
            wpw = self.pack.get(sb.pw.weight)[0]
            z, zst = pw_gemm_stats(wpw, a, Ta)
            nn_ = B * Ta
            rec = dict(x=cur, Tin=Tc, lin=lc, a=a, Ta=Ta, la=la, z=z, n=nn_)
            if last and self.se is not None:
                # SqueezeExcite: u = BN(z); gate = sigmoid(W2 relu(W1 mean_t u)); y = relu(gate * u [+ BN_r(zr)]).
                # mean_t u over ALL Ta frames (the reference pools the unmasked BatchNorm output, citrinet/blocks.py:70-83)
                # follows from the row sums the GEMM epilogue already produced: sum_t u = scale * sum_t z + Ta * shift.
                scale, shift, mean, inv = bn_finalize(zst, nn_, sb.bn, update_running)
                rowsum = zst[..., 0].sum(2) if zst.dim() == 4 else zst[..., 0]
                pool = torch.addcmul(shift * float(Ta), rowsum, scale).contiguous()            # [B, C] sum_t u
                w1 = self.se.fc[0].weight.detach().float().contiguous()
                w2 = self.se.fc[2].weight.detach().float().contiguous()
                gate = ops.se_fc(pool, Ta, w1, w2)
                zr = st_r = None
                sc_r = sh_r = None
                if self.res is not None:
                    rconv, rbn = self.res
                    if Tr != Ta:
                        raise ValueError(f"residual branch length {Tr} != main branch length {Ta}")
                    zr, zrst = residual_gemm()
                    sc_r, sh_r, mean_r, inv_r = bn_finalize(zrst, B * Tr, rbn, update_running)
                    st_r = torch.stack([sc_r, sh_r, mean_r, inv_r])
                y = bn_apply_se(z, scale, shift, zr, sc_r, sh_r, gate, Ta, la if zero_tail else None, True)
                rec.update(zr=zr, stats=torch.stack([scale, shift, mean, inv]), stats_r=st_r, gate=gate, pool=pool,
                           rowsum=rowsum)
            elif last and self.res is not None:
                rconv, rbn = self.res
                if Tr != Ta:
                    raise ValueError(f"residual branch length {Tr} != main branch length {Ta}")
                zr, zrst = residual_gemm()
                y, st, st_r = bn_apply_fused(z, zst, sb.bn, zr, zrst, rbn, Ta, la if zero_tail else None, True,
                                             update_running)
                rec.update(zr=zr, stats=st, stats_r=st_r)
            else:
                tail = la if (not last or zero_tail) else None
                y, st, _ = bn_apply_fused(z, zst, sb.bn, None, None, None, Ta, tail, True, update_running)
                rec.update(stats=st)
            rec["y"] = y
            tape["subs"].append(rec)
            cur, Tc, lc = y, Ta, la
        if update_running and bump_nbt:
            nbt = self._nbt_buffers()
            if nbt:
                torch._foreach_add_(nbt, 1)
        return y, Tc, lc, tape

    def _se_backward(self, g: Tensor, rec, sb: _Sub, zr: Optional[Tensor], mask, Ta: int, has_res: bool):
        """BatchNorm + SqueezeExcite + (residual) + ReLU backward of a Citrinet block's last sub-block.

        With u = BN(z) and dym = dy * relu':  d gate[b,c] = sum_t dym u = scale sum_t(dym z) + shift sum_t(dym)  (from the
        per-utterance partial sums, no extra pass);  the two FCs are [B, C] / [B, C/8] matrices -- torch.matmul;
        d u = dym gate + dm / Ta with dm = d loss / d mean_t(u), so the BatchNorm-backward sums become
        sum du = sum_b (gate s0 + dm),  sum du z = sum_b (gate s1 + dm/Ta rowsum_z)."""
        st, st_r = rec["stats"], rec.get("stats_r")
        gate, pool, rowsum = rec["gate"], rec["pool"], rec["rowsum"]
        sums = bn_bwd_reduce_se(g, rec["z"], zr, Ta, mask, gate)               # [B, C, 3]
        s0, s1 = sums[..., 0], sums[..., 1]
        dgate = st[0] * s1 + st[1] * s0
        w1, w2 = self.se.fc[0].weight.detach().float(), self.se.fc[2].weight.detach().float()
        m = pool / float(Ta)
        hpre = m @ w1.t()
        h = torch.relu(hpre)
        ds = dgate * gate * (1.0 - gate)
        _grad(self.se.fc[2].weight).copy_(ds.t() @ h)
        dh = (ds @ w2) * (hpre > 0)
        _grad(self.se.fc[0].weight).copy_(dh.t() @ m)
        dm = dh @ w1
        addc = (dm / float(Ta)).contiguous()
        part = torch.stack([gate * s0 + dm, gate * s1 + addc * rowsum, torch.zeros_like(s0)], dim=-1).contiguous()
        coef = bn_bwd_coef(part, 1, rec["n"], sb.bn, st[2], st[3])
        coef_r = bn_bwd_coef(sums, 2, rec["n"], self.res[1], st_r[2], st_r[3]) if has_res else None
        return bn_bwd_apply_se(g, rec["z"], zr, coef, coef_r, Ta, mask, gate, addc)

    # -- backward --------------------------------------------------------------------------------------
    def backward(self, tape, dy: Tensor, need_dx: bool = True, side: Optional[_SideStream] = None) -> Optional[Tensor]:
        """``side``: the caller's side stream (the caller joins it); None = own side stream, joined before returning."""
        own_side = side is None
        if own_side:
            side = _SideStream(dy.device)
        n = len(self.subs)
        dx_res = None
        g = dy
        for r in range(n - 1, -1, -1):
            sb, rec = self.subs[r], tape["subs"][r]
            last = r == n - 1
            has_res = last and self.res is not None
            Ta = rec["Ta"]
            # the ReLU mask is rebuilt from z with the forward scale / shift (y is not read again)
            st, st_r = rec["stats"], rec.get("stats_r")
            zr = rec.get("zr") if has_res else None
            mask = (st[0], st[1], st_r[0] if has_res else None, st_r[1] if has_res else None)
            if last and self.se is not None:
                dz, dzr = self._se_backward(g, rec, sb, zr, mask, Ta, has_res)
            else:
                sums = bn_bwd_reduce(g, None, rec["z"], zr, Ta, True, partial=True, mask=mask)
                dz, dzr = bn_bwd_apply_fused(g, rec["z"], sb.bn, st, zr, self.res[1] if has_res else None,
                                             st_r if has_res else None, sums, Ta, True)
            # pointwise conv: weight gradient on the tensor cores, input gradient = W^T dz (masked like `a` was)
            side.run(lambda dz=dz, rec=rec, sb=sb, Ta=Ta: pw_wgrad(dz, rec["a"], Ta, out=_grad(sb.pw.weight)), dz)
            first = r == 0
            need_da = sb.dw is not None or need_dx or not first
            da = None
            if need_da:
                wT = self.pack.get(sb.pw.weight)[1]
                da = ops.pw_gemm(wT, dz, None, None, Ta, None, rec["la"], False, False, None, None, None)
            if sb.dw is not None:
                side.run(lambda da=da, rec=rec, sb=sb, Ta=Ta: dw_wgrad(
                    da, Ta, rec["x"], rec["Tin"], rec["lin"], sb.K, sb.S, sb.D, sb.P, out=_grad(sb.dw.weight),
                    premasked=True), da)
                if (not first) or need_dx:
                    wflip = self.pack.get(sb.dw.weight)[1]
                    if sb.S != 1:
                        # transposed conv = stride-1 conv of the zero-upsampled gradient with flipped taps
                        da_up = scatter_rows(da, Ta, sb.S, rec["Tin"])
                        side.keep.append(da_up)
                        g = ops.dw_conv(da_up, rec["Tin"], wflip, 1, sb.D, sb.D * (sb.K - 1) - sb.P, rec["lin"], True)
                    else:
                        g = ops.dw_conv(da, Ta, wflip, 1, sb.D, sb.D * (sb.K - 1) - sb.P, rec["lin"], True)
                else:
                    g = None
            else:
                g = da
            if has_res:
                rconv, rbn = self.res
                side.run(lambda dzr=dzr, rconv=rconv: pw_wgrad(dzr, tape["xr"], tape["Tr"], out=_grad(rconv.weight)), dzr)
                if need_dx:
                    dx_res = dzr
        if need_dx and self.res is not None:
            # dx = dx_main + W_r^T dz_r, masked by the block-input lengths (epilogue: acc + 1 * y1)
            rconv, _ = self.res
            wrT = self.pack.get(rconv.weight)[1]
            if self.res_stride == 1:
                ones = torch.ones((g.shape[0], wrT.shape[0]), device=g.device, dtype=torch.float32)
                g = ops.pw_gemm(wrT, dx_res, None, None, tape["T"], None, tape["lens"], False, False, None, ones, g)
            else:   # strided residual: its input gradient lands on every res_stride-th frame of dx
                # frame q of the strided branch reads x[q * stride]: valid iff q < ceil(len / stride) = the output length
                dxr = ops.pw_gemm(wrT, dx_res, None, None, tape["Tr"], None, tape["subs"][-1]["la"], False, False, None, None,
                                  None)
                g = scatter_rows(dxr, tape["Tr"], self.res_stride, tape["T"], dst=g)
        if own_side:
            side.join()
        return g if need_dx else None


class EncoderTrainer:
    """train()-mode forward/backward of a whole QuartznetEncoder; gradients are accumulated into ``param.grad``."""

    def __init__(self, encoder: nn.Module):
        self.encoder = encoder
        self.blocks = [BlockTrainer(b) for b in encoder.children()]
        self.extra_entries: List[Tuple[str, nn.Parameter]] = []   # e.g. the decoder weight (CTCTrainStep)
        self.pack: Optional[WeightPack] = None

    def build_pack(self) -> WeightPack:
        ent = [e for bt in self.blocks for e in bt.pack_entries()] + self.extra_entries
        self.pack = WeightPack(ent)
        for bt in self.blocks:
            bt.pack, bt._own_pack = self.pack, False
        return self.pack

    def forward(self, rows: Tensor, T: int, lens: Optional[Tensor], update_running: bool = True):
        if self.pack is None:
            self.build_pack()
        self.pack.refresh()
        side = _SideStream(rows.device) if os.environ.get("THUNDER_B200_FWD_SIDE", "1") != "0" else None
        tapes = []
        for i, bt in enumerate(self.blocks):
            rows, T, lens, tape = bt.forward(rows, T, lens, zero_tail=(i != len(self.blocks) - 1),
                                             update_running=update_running, bump_nbt=False, side=side)
            tapes.append(tape)
        if side is not None:
            side.join()
        if update_running:   # nn.BatchNorm1d bookkeeping for every layer of the model in one multi-tensor launch
            nbt = [b for bt in self.blocks for b in bt._nbt_buffers()]
            if nbt:
                torch._foreach_add_(nbt, 1)
        return rows, T, lens, tapes

    def backward(self, tapes, dy: Tensor, side: Optional[_SideStream] = None) -> None:
        own_side = side is None
        if own_side:
            side = _SideStream(dy.device)
        g = dy
        for i in range(len(self.blocks) - 1, -1, -1):
            g = self.blocks[i].backward(tapes[i], g, need_dx=(i > 0), side=side)
        if own_side:
            side.join()


class FusedAdamW:
    """``torch.optim.AdamW`` (the reference's default optimizer, module.py:32) as ONE kernel launch over every parameter
    (``ts_adamw``): same update rule and defaults (lr 1e-3, betas (0.9, 0.999), eps 1e-8, weight_decay 1e-2).  Gradients are
    read from the flat buffer the training step writes (``param.grad`` are views of it); both moments are flat buffers of
    the same layout; the parameters stay where the module keeps them (device table of pointers, rebuilt if they move).
    After the update the parameters' version counters are bumped so that caches keyed on ``_version`` (the folded inference
    plans) notice the change."""

    def __init__(self, params, flat_grad: Tensor, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2):
        self.params = list(params)
        self.flat = flat_grad
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.exp_avg = torch.zeros_like(flat_grad)
        self.exp_avg_sq = torch.zeros_like(flat_grad)
        self.step_count = 0
        self._build()

    def _build(self) -> None:
        rows, off, tiles = [], 0, 0
        for p in self.params:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise TypeError("FusedAdamW: contiguous fp32 parameters expected")
            n = p.numel()
            rows.append([p.data_ptr(), off, n, tiles])
            off += n
            tiles += (n + 1023) // 1024
        if off != self.flat.numel():
            raise ValueError("FusedAdamW: the flat gradient buffer does not match the parameters")
        self._ptrs = [r[0] for r in rows]
        self.n, self.tiles = len(rows), tiles
        self.table = torch.tensor(rows, dtype=torch.int64).to(self.flat.device)

    def zero_grad(self, set_to_none: bool = False) -> None:
        """Not needed by CTCTrainStep (the backward kernels overwrite every gradient); provided for API familiarity."""
        self.flat.zero_()

    def step(self) -> None:
        if any(p.data_ptr() != q for p, q in zip(self.params, self._ptrs)):
            self._build()
        self.step_count += 1
        b1, b2 = self.betas
        _lib.check(_lib.lib().ts_adamw(_p(self.table), self.n, self.tiles, _p(self.flat), _p(self.exp_avg), _p(self.exp_avg_sq),
                                       self.lr, b1, b2, self.eps, self.weight_decay, 1.0 - b1 ** self.step_count,
                                       1.0 - b2 ** self.step_count, _stream()), "ts_adamw")
        torch.autograd.graph.increment_version(self.params)

    def state_dict(self) -> Dict[str, object]:
        return dict(step=self.step_count, exp_avg=self.exp_avg.clone(), exp_avg_sq=self.exp_avg_sq.clone(), lr=self.lr,
                    betas=self.betas, eps=self.eps, weight_decay=self.weight_decay)

    def load_state_dict(self, sd: Dict[str, object]) -> None:
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.lr, self.betas, self.eps, self.weight_decay = float(sd["lr"]), tuple(sd["betas"]), float(sd["eps"]), float(sd["weight_decay"])


class _CTCLossFn(torch.autograd.Function):
    """Bridge between the kernel-written gradients and torch.autograd (see CTCTrainStep.autograd_loss)."""

    @staticmethod
    def forward(ctx, step, audio, lengths, y, y_lengths, *params):
        flat = step.flat
        fresh = [p.grad is None for p in step.params]
        ensure_grad_views(step.params, flat)
        if all(fresh):
            prev = None                      # nothing accumulated so far: param.grad starts from zero
        else:
            for p, f in zip(step.params, fresh):
                if f:
                    p.grad.zero_()
            prev = flat.clone()              # gradients accumulated by earlier backward() calls survive this step
        loss = step.loss_and_grads(audio, lengths, y, y_lengths)       # the kernels overwrite `flat`
        ctx.grads = flat.clone()
        if prev is None:
            flat.zero_()
        else:
            flat.copy_(prev)
        ctx.step = step
        return loss.clone()

    @staticmethod
    def backward(ctx, gout):
        step, g = ctx.step, ctx.grads
        ctx.grads = None
        if g is None:
            raise RuntimeError("CTCTrainStep.autograd_loss: backward() called twice on the same loss")
        g.mul_(gout)
        out, off = [], 0
        for p in step.params:
            n = p.numel()
            out.append(g[off:off + n].view_as(p) if p.requires_grad else None)
            off += n
        return (None, None, None, None, None, *out)


class _StepGraph:
    """One captured forward + loss + backward for fixed input shapes."""

    def __init__(self, graph, audio, lengths, y, y_len, loss, kernels_per_replay):
        self.graph, self.audio, self.lengths, self.y, self.y_len, self.loss = graph, audio, lengths, y, y_len, loss
        self.kernels_per_replay = kernels_per_replay   # launches of this library's kernels captured in the graph
        self.replays = 0


class CTCTrainStep:
    """One optimisation step of a ``CTCModule`` (QuartzNet and Citrinet encoders), mirroring ``BaseCTCModule.training_step`` +
    ``configure_optimizers`` (src/thunder/module.py:102-127, :173-192): features (no grad) -> encoder forward (kernels) ->
    decoder GEMM -> CTC loss kernel -> decoder / encoder backward (kernels) -> gradient all-reduce (NCCL when initialised)
    -> AdamW (``FusedAdamW``: torch.optim.AdamW's update rule -- the reference's default optimizer -- in one kernel launch).

    All gradients live in ONE flat fp32 buffer (``param.grad`` are views): the all-reduce is a single in-place collective.
    With ``use_graph`` the whole forward + loss + backward (~890 launches for QuartzNet 15x5) is captured in a CUDA graph
    per input shape and replayed; inputs are copied into the graph's static buffers.  Graphs need FIXED OR BUCKETED shapes:
    every new ``(audio.shape, y.shape)`` costs two warm-up passes plus a capture and holds its own memory pool, so at most
    ``MAX_GRAPHS`` are kept (least recently used evicted) -- pad audio / targets to a few bucket lengths before calling, or
    pass ``use_graph=False`` for free-form shapes.

    Not implemented, and refused loudly instead of silently differing from the reference: dropout ``p > 0`` in train mode
    (the reference applies it after every ReLU, quartznet/blocks.py:225-228), BatchNorm modules switched to ``eval()``
    inside a training step (the fine-tuning callback's ``train_bn=False``) and frozen (``requires_grad=False``)
    parameters."""

    MAX_GRAPHS = 8

    def __init__(self, module, lr: float = 3e-4, blank_idx: Optional[int] = None,
                 optimizer: Optional[torch.optim.Optimizer] = None, use_graph: bool = True):
        self.m = module
        for name, mod in list(module.encoder.named_modules()) + list(module.decoder.named_modules()):
            if isinstance(mod, nn.Dropout) and mod.p > 0:
                raise NotImplementedError(
                    f"CTCTrainStep: dropout p={mod.p} at encoder.{name} is not implemented by the training kernels (the "
                    "reference applies it after every ReLU in train mode); build the encoder with dropout=0.0")
        frozen = [n for n, p_ in list(module.encoder.named_parameters()) + list(module.decoder.named_parameters())
                  if not p_.requires_grad]
        if frozen:
            raise NotImplementedError(
                f"CTCTrainStep: {len(frozen)} frozen parameters (requires_grad=False, first: {frozen[0]}): the fused "
                "optimizer updates every encoder / decoder parameter; partial fine-tuning is not implemented")
        self.enc = EncoderTrainer(module.encoder)
        self.enc.extra_entries.append(("pw", module.decoder.weight))
        self.params = [p for p in list(module.encoder.parameters()) + list(module.decoder.parameters())]
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("CTCTrainStep runs on CUDA (sm_100a) modules only; there is no CPU fallback")
        self.flat = flat_grad_views(self.params)
        # default: AdamW like the reference (module.py:32), as one launch of this library; any torch optimizer over
        # `self.params` can be passed instead
        self.opt = optimizer or FusedAdamW(self.params, self.flat, lr=lr)
        self._bn_buffers = [b for n_, b in list(module.encoder.named_buffers()) if "running_" in n_ or "num_batches" in n_]
        self.blank = blank_idx if blank_idx is not None else module.text_transform.vocab.blank_idx
        self.use_graph = use_graph
        self._graphs: Dict[tuple, _StepGraph] = {}

    # -- forward + loss + backward (all kernels of this library, no host synchronisation) ------------------------------
    def _forward_backward(self, audio: Tensor, lengths: Tensor, y: Tensor, y_lengths: Tensor,
                          update_running: bool = True) -> Tensor:
        m = self.m
        dec = m.decoder
        V = dec.weight.shape[0]
        Vp = V if V % 8 == 0 else (V + 63) // 64 * 64    # = WeightPack's leading dimension of the transposed weight
        with torch.no_grad():
            N = audio.shape[-1]
            hop = m.audio_transform[1].hop_length
            F = 1 + N // hop
            feats, feat_len = m.audio_transform.features(audio, lengths, bf16_pitch=ops.row_pitch(F))
            rows, T, l32o, tapes = self.enc.forward(feats, F, feat_len.to(torch.int32), update_running)
            # decoder: logits = W_d enc + b (f32 rows), then CTC
            wd, wdT = self.enc.pack.get(dec.weight)     # bf16 [V, Cd] and its transpose padded to [Cd, Vp]
            logits = ops.pw_gemm(wd, rows, None, None, T, dec.bias.detach(), None, True, False, None, None, None)
            loss_b, dlogits = ctc_loss(logits, T, l32o, y, y_lengths.to(torch.int64), self.blank, Vp)
            loss = loss_b.mean()
            # decoder gradients: d enc = W_d^T dlogits is on the critical path; dW_d = dlogits enc^T and db = sum dlogits run
            # on the forked stream together with the encoder's weight gradients
            side = _SideStream(rows.device)

            def _decoder_param_grads():
                dwd = pw_wgrad(dlogits, rows, T)
                _grad(dec.weight).copy_(dwd[:V].view_as(dec.weight))
                if dec.bias is not None:
                    _grad(dec.bias).copy_(row_stats_partial(dlogits, T).sum(0)[:V, 0])

            side.run(_decoder_param_grads, dlogits, rows)
            d_enc = ops.pw_gemm(wdT, dlogits, None, None, T, None, None, False, False, None, None, None)
            self.enc.backward(tapes, d_enc, side=side)
            side.join()
        return loss

    def _capture(self, audio: Tensor, lengths: Tensor, y: Tensor, y_lengths: Tensor) -> _StepGraph:
        sa, sl, sy, syl = audio.clone(), lengths.clone(), y.clone(), y_lengths.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):   # warm-up outside capture (lazy kernel attributes, allocator); leaves BN statistics alone
                self._forward_backward(sa, sl, sy, syl, update_running=False)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(graph):
            loss = self._forward_backward(sa, sl, sy, syl)
        return _StepGraph(graph, sa, sl, sy, syl, loss, _lib.launch_count() - n0)

    def loss_and_grads(self, audio: Tensor, lengths: Tensor, y: Tensor, y_lengths: Tensor) -> Tensor:
        """Mean CTC loss (device scalar); parameter gradients are left in ``param.grad``."""
        if not self.use_graph:
            self._check_bn_mode()
            loss = self._forward_backward(audio, lengths, y, y_lengths)
            self._bump_buffer_versions()
            return loss
        # user code may have dropped / replaced param.grad (zero_grad(set_to_none=True)); the captured graphs write into the
        # flat buffer itself, so re-attaching the views is all that is needed
        ensure_grad_views(self.params, self.flat)
        ptrs = tuple(p.data_ptr() for p in self.params)
        if ptrs != getattr(self, "_captured_ptrs", ptrs):
            self._graphs.clear()          # parameters were re-allocated: the captured graphs point at stale storage
        self._captured_ptrs = ptrs
        self._check_bn_mode()
        key = (tuple(audio.shape), tuple(y.shape))
        g = self._graphs.pop(key, None)
        if g is None:
            while len(self._graphs) >= self.MAX_GRAPHS:      # least recently used first (dicts keep insertion order)
                del self._graphs[next(iter(self._graphs))]
            g = self._capture(audio, lengths, y, y_lengths)
        self._graphs[key] = g                                # (re-)insert as most recently used
        g.audio.copy_(audio, non_blocking=True)
        g.lengths.copy_(lengths, non_blocking=True)
        g.y.copy_(y, non_blocking=True)
        g.y_len.copy_(y_lengths, non_blocking=True)
        g.graph.replay()
        g.replays += 1
        self._bump_buffer_versions()
        return g.loss.clone()     # the graph's static output is overwritten by the next replay

    def _check_bn_mode(self) -> None:
        for name, mod in self.m.encoder.named_modules():
            if isinstance(mod, nn.BatchNorm1d) and not mod.training:
                raise NotImplementedError(
                    f"CTCTrainStep: BatchNorm encoder.{name} is in eval() mode; the training kernels always normalise with "
                    "batch statistics and update the running ones (frozen-BN fine-tuning is not implemented)")

    def _bump_buffer_versions(self) -> None:
        """The kernels update BatchNorm running statistics through raw pointers; tell torch (caches keyed on `_version`)."""
        if self._bn_buffers:
            torch.autograd.graph.increment_version(self._bn_buffers)

    def graph_launches(self) -> int:
        """Kernel launches of this library executed through graph replays so far (ts_launch_count only sees eager ones)."""
        return sum(g.kernels_per_replay * g.replays for g in self._graphs.values())

    def allreduce_grads(self) -> None:
        allreduce_mean_(self.flat)

    def training_step(self, batch, batch_idx: int = 0) -> Tensor:
        """The reference's entry point (module.py:102-127): ``batch = (audio, audio_lengths, texts)``; the texts are encoded
        with ``text_transform.encode`` on the host, then one optimisation step runs.  Returns the loss."""
        audio, audio_lengths, texts = batch
        y, y_lengths = self.m.text_transform.encode(texts, device=audio.device)
        return self.step(audio, audio_lengths, y, y_lengths)

    def step(self, audio: Tensor, lengths: Tensor, y: Tensor, y_lengths: Tensor) -> Tensor:
        loss = self.loss_and_grads(audio, lengths, y, y_lengths)
        self.allreduce_grads()
        self.opt.step()
        return loss

    def autograd_loss(self, audio: Tensor, lengths: Tensor, y: Tensor, y_lengths: Tensor) -> Tensor:
        """The loss as a node of torch's autograd graph, for callers that drive the optimisation themselves the way
        Lightning's automatic optimisation drives the reference (``loss = training_step(...); loss.backward();
        optimizer.step(); optimizer.zero_grad()``): forward runs the whole captured forward + loss + backward, ``backward()``
        hands the finished parameter gradients to autograd, which ACCUMULATES them into ``param.grad`` as usual (gradient
        accumulation over micro-batches and ``zero_grad(set_to_none=True)`` both behave like torch)."""
        return _CTCLossFn.apply(self, audio, lengths, y, y_lengths, *self.params)

    def fit_stream(self, batches, depth: int = 2):
        """Training loop over HOST batches ``(audio, audio_lengths, y, y_lengths)`` of one fixed shape (pinned memory for
        true overlap) -- what Lightning's loop does around ``training_step`` (module.py:102-127) with a pinned-memory
        DataLoader: the host->device copies of batch i+1 run on a copy stream while step i computes, and the loss of
        step i is read back through pinned memory ``depth - 1`` steps late, so the host never idles the GPU.  Every step
        still copies its own inputs in and its own loss out.  Yields one float loss per step, in order."""
        depth = max(1, int(depth))
        dev = self.params[0].device
        copy_stream = torch.cuda.Stream(device=dev)
        stage, host_loss = None, None
        h2d_done = [torch.cuda.Event() for _ in range(depth)]
        slot_free = [torch.cuda.Event() for _ in range(depth)]
        done = [torch.cuda.Event() for _ in range(depth)]
        pending = []
        for i, batch in enumerate(batches):
            if stage is None:
                stage = [[torch.empty(t.shape, dtype=t.dtype, device=dev) for t in batch] for _ in range(depth)]
                host_loss = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(depth)]
            s = i % depth
            cur = torch.cuda.current_stream(dev)
            with torch.cuda.stream(copy_stream):
                if i >= depth:
                    copy_stream.wait_event(slot_free[s])
                for d, h in zip(stage[s], batch):
                    d.copy_(h, non_blocking=True)
                h2d_done[s].record(copy_stream)
            cur.wait_event(h2d_done[s])
            loss = self.step(*stage[s])
            slot_free[s].record(cur)
            host_loss[s].copy_(loss, non_blocking=True)
            done[s].record(cur)
            pending.append(s)
            if len(pending) == depth:
                j = pending.pop(0)
                done[j].synchronize()
                yield float(host_loss[j])
        for j in pending:
            done[j].synchronize()
            yield float(host_loss[j])
