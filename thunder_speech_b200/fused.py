"""Execution plans for Quartznet / Citrinet blocks on the B200 kernels.

A *plan* is what a block's parameters become for inference: depthwise taps as ``[C, K]`` f32, pointwise /
residual weights as bf16 ``[Cout, Cin]`` with the eval-mode BatchNorm scale folded in
(``scale = gamma / sqrt(running_var + 1e-3)``, quartznet/blocks.py:222), BatchNorm shifts summed into one
``shift[Cout]`` vector, SqueezeExcite FC weights as f32.  Plans are rebuilt when any parameter or buffer
changes (``_version`` / ``data_ptr`` key).

``run_block`` strings the ops together on bf16 padded rows ``[B, C, pitch]``:

    per sub-block:   dw_conv (masks input & output)  ->  pw_gemm (+BN shift, ReLU, zero tail)
    last sub-block:  pw_gemm with the residual 1x1 conv as a second K segment (+ReLU)           (QuartzNet)
                     pw_gemm(+pool) -> se_fc -> pw_gemm(residual, gate*y1 epilogue, ReLU)        (Citrinet)

which reproduces ``QuartznetBlock.forward`` (quartznet/blocks.py:317-338) / ``CitrinetBlock.forward``
(citrinet/blocks.py:177-197) in eval mode, including their masking rules: every conv sees its input zeroed
beyond the running length, SE and BN see unmasked frames.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch
from torch import Tensor, nn

from . import ops
from .blocks import conv_out_length

BN_EPS_DEFAULT = 1e-3


@dataclass
class SubPlan:
    dw_w: Optional[Tensor]      # [C, K] f32, None for a non-separable 1x1 conv
    K: int
    S: int
    D: int
    P: int
    pw_w: Tensor                # [Cout, Cin] bf16, BN scale folded ([Cout, Cin * K] for a full convolution)
    shift: Tensor               # [Cout] f32
    full: bool = False          # non-separable conv with kernel_size > 1: im2col rows + one GEMM over Cin * K channels


@dataclass
class BlockPlan:
    subs: List[SubPlan] = field(default_factory=list)
    res_w: Optional[Tensor] = None      # [Cout, Cin] bf16, BN scale folded
    res_shift: Optional[Tensor] = None  # [Cout] f32
    res_stride: int = 1
    ones: Optional[Tensor] = None       # [Cin, 1] f32 taps for the strided residual gather
    total_shift: Optional[Tensor] = None  # last sub-block shift + residual shift (QuartzNet fused epilogue)
    se_w1: Optional[Tensor] = None      # [H, C] f32
    se_w2: Optional[Tensor] = None      # [C, H] f32
    in_channels: int = 0
    out_channels: int = 0


def _bn_fold(bn: nn.BatchNorm1d) -> Tuple[Tensor, Tensor]:
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    shift = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    return scale, shift


def _fold_pw(conv: nn.Conv1d, bn: nn.BatchNorm1d, dtype: torch.dtype = torch.bfloat16) -> Tuple[Tensor, Tensor]:
    scale, shift = _bn_fold(bn)
    w = conv.weight.detach().float()[:, :, 0] * scale[:, None]
    if conv.bias is not None:
        shift = shift + conv.bias.detach().float() * scale
    if dtype == torch.float16:   # saturate like the kernels do (a folded weight beyond 65504 is pathological anyway)
        w = w.clamp(-65504.0, 65504.0)
    return w.to(dtype).contiguous(), shift.contiguous()


def params_key(module: nn.Module):
    return tuple((t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers()))


def build_block_plan(block: nn.Module, dtype: torch.dtype = torch.bfloat16) -> BlockPlan:
    """Reads a block laid out like the reference (``mconv`` = [dw, pw, BN, (ReLU, Dropout)]*, optional SE,
    ``res`` = [conv1x1, BN]) and folds it into a plan.  Non-separable convs: kernel_size 1 is a plain GEMM (the only
    non-separable layer of the models, quartznet/blocks.py:399-407); kernel_size > 1 (the block's DEFAULT, :232-243) runs as
    im2col rows + one GEMM over ``Cin * K`` channels."""
    from .quartznet.blocks import MaskedConv1d  # local import: blocks.py imports this module

    plan = BlockPlan()
    pending: List[nn.Module] = []
    for layer in block.mconv.children():
        if isinstance(layer, MaskedConv1d):
            pending.append(layer)
            continue
        inner = layer.layer[0]
        if isinstance(inner, nn.BatchNorm1d):
            if block.separable:
                dwl, pwl = pending  # [depthwise, pointwise] (quartznet/blocks.py:193-210)
            else:
                dwl, (pwl,) = None, pending
            conv = pwl.conv
            if conv.groups != 1:
                raise NotImplementedError(f"grouped non-depthwise convolutions are not implemented (groups={conv.groups})")
            if dwl is None and (conv.kernel_size[0] != 1 or conv.stride[0] != 1):
                # full convolution (QuartznetBlock's default separable=False, quartznet/blocks.py:212-219): the BN-folded
                # weight [Cout, Cin, K] IS the [Cout, Cin * K] operand of one GEMM over im2col rows
                scale, shift = _bn_fold(inner)
                w = conv.weight.detach().float() * scale[:, None, None]
                if conv.bias is not None:
                    shift = shift + conv.bias.detach().float() * scale
                if dtype == torch.float16:
                    w = w.clamp(-65504.0, 65504.0)
                w2 = w.reshape(w.shape[0], -1)
                if w2.shape[1] % 8:     # whole 16-byte groups along the GEMM's K (matches ops.im2col_rows' zero rows)
                    w2 = torch.nn.functional.pad(w2, (0, 8 - w2.shape[1] % 8))
                plan.subs.append(SubPlan(None, pwl.kernel_size, pwl.stride, pwl.dilation, pwl.padding,
                                         w2.to(dtype).contiguous(), shift.contiguous(), True))
                pending = []
                continue
            pw_w, shift = _fold_pw(conv, inner, dtype)
            if dwl is not None:
                dw_w = dwl.conv.weight.detach().float()[:, 0, :].contiguous()
                plan.subs.append(SubPlan(dw_w, dwl.kernel_size, dwl.stride, dwl.dilation, dwl.padding, pw_w, shift))
            else:
                plan.subs.append(SubPlan(None, 1, 1, 1, 0, pw_w, shift))
            pending = []
        elif hasattr(inner, "fc") and hasattr(inner, "pool"):  # SqueezeExcite
            plan.se_w1 = inner.fc[0].weight.detach().float().contiguous()
            plan.se_w2 = inner.fc[2].weight.detach().float().contiguous()
        # ReLU / Dropout(eval) carry no state: fused into the GEMM epilogues
    if block.res is not None:
        rlayers = list(block.res.children())
        rconv, rbn = rlayers[0], rlayers[1].layer[0]
        plan.res_w, plan.res_shift = _fold_pw(rconv.conv, rbn, dtype)
        plan.res_stride = rconv.stride
        if plan.res_stride > 1:
            plan.ones = torch.ones((rconv.conv.in_channels, 1), device=plan.res_w.device, dtype=torch.float32)
        plan.total_shift = (plan.subs[-1].shift + plan.res_shift).contiguous()
    first = next(iter(block.mconv.children()))
    plan.in_channels = first.conv.in_channels
    plan.out_channels = plan.subs[-1].pw_w.shape[0]
    return plan


def _next_lens(lens: Optional[Tensor], K: int, S: int, D: int, P: int) -> Optional[Tensor]:
    if lens is None:
        return None
    if S == 1 and 2 * P == D * (K - 1):
        return lens  # length preserving (odd kernel, "same" padding)
    return ops.conv_lengths(lens, K, S, D, P)


def run_block(plan: BlockPlan, x: Tensor, T: int, lens: Optional[Tensor], zero_tail: bool,
              pool: Optional[Tensor] = None) -> Tuple[Tensor, int, Optional[Tensor]]:
    """``x``: bf16 rows ``[B, Cin, pitch]`` holding ``T`` frames, already zero beyond ``lens`` (i32 ``[B]``;
    ``None`` = every frame valid).  Returns ``(y rows, T_out, lens_out)``.  ``zero_tail``: store the block output
    with frames beyond ``lens_out`` zeroed (legal whenever the consumer is another MaskedConv1d)."""
    x_in, T_in, lens_in = x, T, lens
    cur, Tc, lc = x, T, lens
    n = len(plan.subs)
    out = None
    for r, sb in enumerate(plan.subs):
        last = r == n - 1
        if sb.dw_w is not None:
            a = ops.dw_conv(cur, Tc, sb.dw_w, sb.S, sb.D, sb.P, lc, True)  # rows are zero beyond lc
            Ta = conv_out_length(Tc, sb.K, sb.S, sb.P, sb.D)
            la = _next_lens(lc, sb.K, sb.S, sb.D, sb.P)
        elif sb.full:
            a = ops.im2col_rows(cur, Tc, sb.K, sb.S, sb.D, sb.P, lc)      # [B, Cin * K, pitch'] (input masked by lc)
            Ta = conv_out_length(Tc, sb.K, sb.S, sb.P, sb.D)
            la = _next_lens(lc, sb.K, sb.S, sb.D, sb.P)
        else:
            a, Ta, la = cur, Tc, lc
        if not last:
            cur = ops.pw_gemm(sb.pw_w, a, None, None, Ta, sb.shift, la, False, True, None, None, None, True)
            Tc, lc = Ta, la
            continue
        out_lens = la if zero_tail else None
        xr = None
        if plan.res_w is not None:
            xr = x_in
            if plan.res_stride > 1:  # 1x1 conv with stride: gather every stride-th (masked) input frame
                xr = ops.gather_rows(x_in, T_in, plan.res_stride, lens_in)
        if plan.se_w1 is None:
            if xr is not None:
                out = ops.pw_gemm(sb.pw_w, a, plan.res_w, xr, Ta, plan.total_shift, out_lens, False, True, None, None,
                                  None, True)
            else:
                out = ops.pw_gemm(sb.pw_w, a, None, None, Ta, sb.shift, out_lens, False, True, None, None, None, True)
        else:
            B = a.shape[0]
            if pool is None:    # (an encoder hands every SE block a slice of ONE buffer it zeroed with a single fill)
                pool = torch.zeros((B, plan.out_channels), device=a.device, dtype=torch.int64)   # fixed-point sums
            y1 = ops.pw_gemm(sb.pw_w, a, None, None, Ta, sb.shift, None, False, False, pool, None, None, True)
            gate = ops.se_fc(pool, Ta, plan.se_w1, plan.se_w2)
            if xr is not None:
                out = ops.pw_gemm(plan.res_w, xr, None, None, Ta, plan.res_shift, out_lens, False, True, None, gate,
                                  y1, True)
            else:
                out = ops.se_apply(y1, gate, out_lens, True)
        Tc, lc = Ta, la
    return out, Tc, lc


class PlannedBlock(nn.Module):
    """Mixin-style base of ``QuartznetBlock`` / ``CitrinetBlock``: lazy plan + the three ways to run it."""

    def _plan(self, dtype: Optional[torch.dtype] = None) -> BlockPlan:
        """Folded operands in the row format ``dtype`` (None = the package default precision)."""
        if dtype is None:
            from . import row_dtype

            dtype = row_dtype()
        key = (params_key(self), dtype)
        cached = getattr(self, "_plan_cache", None)
        if cached is None or cached[0] != key:
            cached = (key, build_block_plan(self, dtype))
            object.__setattr__(self, "_plan_cache", cached)
            # the folded operands are complete before any kernel that is told they are constants (TS_PW_CONST_WEIGHTS) starts
            if cached[1].subs[0].pw_w.is_cuda and not torch.cuda.is_current_stream_capturing():
                torch.cuda.current_stream(cached[1].subs[0].pw_w.device).synchronize()
        return cached[1]

    def _check_eval(self):
        if self.training:
            raise NotImplementedError(
                "this entry point is the inference (eval) forward; call .eval() first.  The train()-mode forward with "
                "batch-statistics BatchNorm and the backward kernels run through CTCModule.training_step / "
                "thunder_speech_b200.train.CTCTrainStep (whole model) or train.BlockTrainer (one block).")

    def forward_rows(self, x: Tensor, T: int, lens: Optional[Tensor], zero_tail: bool, pool: Optional[Tensor] = None):
        """``pool``: a zeroed int64 ``[B, out_channels]`` SqueezeExcite accumulator owned by the caller (optional)."""
        self._check_eval()
        with torch.no_grad():
            return run_block(self._plan(x.dtype), x, T, lens, zero_tail, pool)

    def se_channels(self, dtype: Optional[torch.dtype] = None) -> int:
        """Channels of this block's SqueezeExcite pool (0 without SE): lets an encoder zero all pools with one fill."""
        plan = self._plan(dtype)
        return plan.out_channels if plan.se_w1 is not None else 0

    def out_lengths(self, lengths: Tensor) -> Tensor:
        """Lengths after the main branch, computed with the reference's own formula on the caller's tensor
        (dtype preserved, quartznet/blocks.py:142-156)."""
        for sb in self._plan().subs:
            if sb.dw_w is not None or sb.full:
                lengths = conv_out_length(lengths, sb.K, sb.S, sb.P, sb.D)
        return lengths

    def forward(self, x: Tensor, lengths: Tensor) -> Tuple[Tensor, Tensor]:
        """Drop-in ``(x[B,C,T], lengths[B]) -> (y[B,C',T'] f32, lengths')``."""
        self._check_eval()
        with torch.no_grad():
            from . import row_dtype

            l32 = ops.lengths_i32(lengths)
            rows = ops.pack_rows(x, l32, row_dtype() == torch.float16)
            y, T_out, _ = run_block(self._plan(rows.dtype), rows, x.shape[-1], l32, False)
            return ops.unpack_rows(y, T_out), self.out_lengths(lengths)
