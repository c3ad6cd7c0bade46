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

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor, nn

from . import _lib, ops
from .blocks import conv_out_length

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
    _lib.check(_lib.lib().ts_row_stats(_p(z), B, C, T, pitch, _p(st), _stream()), "ts_row_stats")
    return st.sum(0, dtype=torch.float64)


def bn_apply(z, scale, shift, zr, scale_r, shift_r, T, lens, relu=True) -> Tensor:
    B, C, pitch = z.shape
    y = torch.empty_like(z)
    _lib.check(_lib.lib().ts_bn_apply(_p(z), _p(scale), _p(shift), _p(zr), _p(scale_r), _p(shift_r), B, C, T, pitch,
                                      _p(lens), int(relu), _p(y), _stream()), "ts_bn_apply")
    return y


def bn_bwd_reduce(dy, y, z, zr, T, relu=True) -> Tensor:
    B, C, pitch = z.shape
    sums = torch.empty((B, C, 3), device=z.device, dtype=torch.float32)
    _lib.check(_lib.lib().ts_bn_bwd_reduce(_p(dy), _p(y), _p(z), _p(zr), B, C, T, pitch, int(relu), _p(sums), _stream()),
               "ts_bn_bwd_reduce")
    return sums.sum(0, dtype=torch.float64)


def bn_bwd_apply(dy, y, z, zr, coef, coef_r, T, relu=True) -> Tuple[Tensor, Optional[Tensor]]:
    B, C, pitch = z.shape
    dz = torch.empty_like(z)
    dzr = torch.empty_like(z) if zr is not None else None
    _lib.check(_lib.lib().ts_bn_bwd_apply(_p(dy), _p(y), _p(z), _p(zr), _p(coef), _p(coef_r), B, C, T, pitch, int(relu),
                                          _p(dz), _p(dzr), _stream()), "ts_bn_bwd_apply")
    return dz, dzr


def pw_wgrad(dz: Tensor, a: Tensor, T: int) -> Tensor:
    """dW[co, ci] = sum_{b,t} dz[b,co,t] a[b,ci,t] (fp32)."""
    B, Cout, pz = dz.shape
    Cin, pa = a.shape[1], a.shape[2]
    tiles = ((Cout + 127) // 128) * ((Cin + 255) // 256)
    nsplit = max(1, min(B, 148 // max(tiles, 1)))
    part = torch.empty((nsplit, Cout, Cin), device=dz.device, dtype=torch.float32)
    _lib.check(_lib.lib().ts_pw_wgrad(_p(dz), pz, _p(a), pa, B, Cout, Cin, T, nsplit, _p(part), _stream()), "ts_pw_wgrad")
    return part.sum(0)


def dw_wgrad(da: Tensor, T_out: int, x: Tensor, T_in: int, lens_in: Optional[Tensor], K: int, S: int, D: int, P: int
             ) -> Tensor:
    B, C, po = da.shape
    pi = x.shape[2]
    bchunk = max(1, min(B, (B + 7) // 8))   # ~8 batch slices per channel: C x 8 CTAs
    nchunk = (B + bchunk - 1) // bchunk
    part = torch.empty((nchunk, C, K), device=da.device, dtype=torch.float32)
    _lib.check(_lib.lib().ts_dw_wgrad(_p(da), T_out, po, _p(x), T_in, pi, _p(lens_in), B, C, K, S, D, P, bchunk, _p(part),
                                      _stream()), "ts_dw_wgrad")
    return part.sum(0)


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


def _bn_forward_stats(bn: nn.BatchNorm1d, st: Tensor, n: int, update_running: bool):
    """batch mean / biased var from (sum, sumsq); returns (scale, shift, mean, inv) as f32 [C]; updates running stats like
    nn.BatchNorm1d(momentum=0.1) does in train() (unbiased variance for the running estimate)."""
    mean = st[:, 0] / n
    var = torch.clamp(st[:, 1] / n - mean * mean, min=0.0)
    inv = torch.rsqrt(var + bn.eps)
    g = bn.weight.detach().double()
    scale = g * inv
    shift = bn.bias.detach().double() - mean * scale
    if update_running and bn.track_running_stats:
        with torch.no_grad():
            m = bn.momentum if bn.momentum is not None else BN_MOMENTUM
            bn.running_mean.mul_(1 - m).add_(m * mean.float())
            bn.running_var.mul_(1 - m).add_(m * (var * n / max(n - 1, 1)).float())
            bn.num_batches_tracked += 1
    return scale.float().contiguous(), shift.float().contiguous(), mean, inv


def _bn_backward_coef(bn: nn.BatchNorm1d, s0: Tensor, s1: Tensor, mean: Tensor, inv: Tensor, n: int):
    """(dgamma, dbeta, coef[C,3]) with dz = coef0 * dym + coef1 * z + coef2."""
    g = bn.weight.detach().double()
    dbeta = s0
    dgamma = inv * (s1 - mean * s0)
    a = g * inv
    b = -g * inv * inv * dgamma / n
    c = -a * dbeta / n - b * mean
    coef = torch.stack([a, b, c], dim=1).float().contiguous()
    return dgamma.float(), dbeta.float(), coef


def _accum(p: nn.Parameter, g: Tensor):
    g = g.reshape(p.shape).to(p.dtype)
    if p.grad is None:
        p.grad = g.clone()
    else:
        p.grad.add_(g)


def _bf16_taps(w: Tensor) -> Tensor:
    """Depthwise taps [C, K] as the training step computes with them: rounded to bf16 (the Toeplitz tensor-core path
    holds its taps in bf16; the SIMT paths get the same rounded values so every layer sees one set of weights)."""
    return w.detach()[:, 0, :].to(torch.bfloat16).float().contiguous()


class BlockTrainer:
    """train()-mode forward / backward of one QuartznetBlock on bf16 rows."""

    def __init__(self, block: nn.Module):
        from .quartznet.blocks import MaskedConv1d

        self.block = block
        self.subs: List[_Sub] = []
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
                raise NotImplementedError("training step: SqueezeExcite (Citrinet) backward is not implemented yet")
        self.res = None
        if block.res is not None:
            rl = list(block.res.children())
            if rl[0].stride != 1:
                raise NotImplementedError("training step: strided residual branches are not implemented yet")
            self.res = (rl[0].conv, rl[1].layer[0])

    # -- forward ---------------------------------------------------------------------------------------
    def forward(self, x: Tensor, T: int, lens: Optional[Tensor], zero_tail: bool, update_running: bool = True):
        tape = dict(x=x, T=T, lens=lens, subs=[])
        cur, Tc, lc = x, T, lens
        B = x.shape[0]
        n = len(self.subs)
        y = None
        for r, sb in enumerate(self.subs):
            last = r == n - 1
            if sb.dw is not None:
                w = _bf16_taps(sb.dw.weight)
                a = ops.dw_conv(cur, Tc, w, sb.S, sb.D, sb.P, lc, True)
                Ta = conv_out_length(Tc, sb.K, sb.S, sb.P, sb.D)
                la = lc if (lc is None or (sb.S == 1 and 2 * sb.P == sb.D * (sb.K - 1))) else ops.conv_lengths(
                    lc, sb.K, sb.S, sb.D, sb.P)
            else:
                a, Ta, la = cur, Tc, lc
            wpw = sb.pw.weight.detach()[:, :, 0].to(torch.bfloat16).contiguous()
            z = ops.pw_gemm(wpw, a, None, None, Ta, None, None, False, False, None, None, None)
            nn_ = B * Ta
            scale, shift, mean, inv = _bn_forward_stats(sb.bn, row_stats(z, Ta), nn_, update_running)
            rec = dict(x=cur, Tin=Tc, lin=lc, a=a, Ta=Ta, la=la, z=z, mean=mean, inv=inv, n=nn_, wpw=wpw)
            if last and self.res is not None:
                rconv, rbn = self.res
                wr = rconv.weight.detach()[:, :, 0].to(torch.bfloat16).contiguous()
                zr = ops.pw_gemm(wr, x, None, None, T, None, None, False, False, None, None, None)
                scale_r, shift_r, mean_r, inv_r = _bn_forward_stats(rbn, row_stats(zr, T), B * T, update_running)
                y = bn_apply(z, scale, shift, zr, scale_r, shift_r, Ta, la if zero_tail else None, True)
                rec.update(zr=zr, mean_r=mean_r, inv_r=inv_r, wr=wr)
            else:
                tail = la if (not last or zero_tail) else None
                y = bn_apply(z, scale, shift, None, None, None, Ta, tail, True)
            rec["y"] = y
            tape["subs"].append(rec)
            cur, Tc, lc = y, Ta, la
        return y, Tc, lc, tape

    # -- backward --------------------------------------------------------------------------------------
    def backward(self, tape, dy: Tensor, need_dx: bool = True) -> Optional[Tensor]:
        n = len(self.subs)
        dx_res = None
        g = dy
        for r in range(n - 1, -1, -1):
            sb, rec = self.subs[r], tape["subs"][r]
            last = r == n - 1
            has_res = last and self.res is not None
            Ta = rec["Ta"]
            sums = bn_bwd_reduce(g, rec["y"], rec["z"], rec.get("zr") if has_res else None, Ta, True)
            dgamma, dbeta, coef = _bn_backward_coef(sb.bn, sums[:, 0], sums[:, 1], rec["mean"], rec["inv"], rec["n"])
            _accum(sb.bn.weight, dgamma)
            _accum(sb.bn.bias, dbeta)
            coef_r = None
            if has_res:
                rconv, rbn = self.res
                dgr, dbr, coef_r = _bn_backward_coef(rbn, sums[:, 0], sums[:, 2], rec["mean_r"], rec["inv_r"], rec["n"])
                _accum(rbn.weight, dgr)
                _accum(rbn.bias, dbr)
            dz, dzr = bn_bwd_apply(g, rec["y"], rec["z"], rec.get("zr") if has_res else None, coef, coef_r, Ta, True)
            # pointwise conv: weight gradient on the tensor cores, input gradient = W^T dz (masked like `a` was)
            _accum(sb.pw.weight, pw_wgrad(dz, rec["a"], Ta))
            first = r == 0
            need_da = sb.dw is not None or need_dx or not first
            da = None
            if need_da:
                wT = rec["wpw"].t().contiguous()
                da = ops.pw_gemm(wT, dz, None, None, Ta, None, rec["la"], False, False, None, None, None)
            if sb.dw is not None:
                _accum(sb.dw.weight, dw_wgrad(da, Ta, rec["x"], rec["Tin"], rec["lin"], sb.K, sb.S, sb.D, sb.P))
                if (not first) or need_dx:
                    if sb.S != 1:
                        raise NotImplementedError("training step: input gradient of a strided depthwise conv")
                    wflip = _bf16_taps(sb.dw.weight).flip(-1).contiguous()
                    g = ops.dw_conv(da, Ta, wflip, 1, sb.D, sb.D * (sb.K - 1) - sb.P, rec["lin"], True)
                else:
                    g = None
            else:
                g = da
            if has_res:
                rconv, rbn = self.res
                _accum(rconv.weight, pw_wgrad(dzr, tape["x"], tape["T"]))
                if need_dx:
                    dx_res = dzr
        if need_dx and self.res is not None:
            # dx = dx_main + W_r^T dz_r, masked by the block-input lengths (epilogue: acc + 1 * y1)
            rconv, _ = self.res
            wrT = tape["subs"][-1]["wr"].t().contiguous()
            ones = torch.ones((g.shape[0], wrT.shape[0]), device=g.device, dtype=torch.float32)
            g = ops.pw_gemm(wrT, dx_res, None, None, tape["T"], None, tape["lens"], False, False, None, ones, g)
        return g if need_dx else None


class EncoderTrainer:
    """train()-mode forward/backward of a whole QuartznetEncoder; gradients are accumulated into ``param.grad``."""

    def __init__(self, encoder: nn.Module):
        self.encoder = encoder
        self.blocks = [BlockTrainer(b) for b in encoder.children()]

    def forward(self, rows: Tensor, T: int, lens: Optional[Tensor], update_running: bool = True):
        tapes = []
        for i, bt in enumerate(self.blocks):
            rows, T, lens, tape = bt.forward(rows, T, lens, zero_tail=(i != len(self.blocks) - 1),
                                             update_running=update_running)
            tapes.append(tape)
        return rows, T, lens, tapes

    def backward(self, tapes, dy: Tensor) -> None:
        g = dy
        for i in range(len(self.blocks) - 1, -1, -1):
            g = self.blocks[i].backward(tapes[i], g, need_dx=(i > 0))


class CTCTrainStep:
    """One optimisation step of a ``CTCModule`` (QuartzNet family): features (no grad) -> encoder (kernels) -> decoder +
    CTC loss (torch autograd) -> encoder backward (kernels) -> gradient all-reduce (NCCL when initialised) -> AdamW."""

    def __init__(self, module, lr: float = 3e-4, blank_idx: Optional[int] = None, optimizer: Optional[torch.optim.Optimizer] = None):
        self.m = module
        self.enc = EncoderTrainer(module.encoder)
        self.params = [p for p in list(module.encoder.parameters()) + list(module.decoder.parameters())]
        self.opt = optimizer or torch.optim.AdamW(self.params, lr=lr)
        self.blank = blank_idx if blank_idx is not None else module.text_transform.vocab.blank_idx

    def loss_and_grads(self, audio: Tensor, lengths: Tensor, y: Tensor, y_lengths: Tensor) -> Tensor:
        m = self.m
        for p in self.params:
            p.grad = None
        with torch.no_grad():
            N = audio.shape[-1]
            hop = m.audio_transform[1].hop_length
            F = 1 + N // hop
            feats, feat_len = m.audio_transform.features(audio, lengths, bf16_pitch=ops.row_pitch(F))
            l32 = feat_len.to(torch.int32)
            rows, T, l32o, tapes = self.enc.forward(feats, F, l32)
            enc_out = ops.unpack_rows(rows, T)                      # [B, C, T] f32 (layout change for torch autograd)
        enc_out.requires_grad_(True)
        logits = torch.nn.functional.conv1d(enc_out, m.decoder.weight, m.decoder.bias)
        logprobs = torch.nn.functional.log_softmax(logits.permute(2, 0, 1), dim=2)
        loss = torch.nn.functional.ctc_loss(logprobs, y, l32o.long(), y_lengths, blank=self.blank, reduction="mean",
                                            zero_infinity=True)
        loss.backward()
        with torch.no_grad():
            self.enc.backward(tapes, ops.pack_rows(enc_out.grad))
        return loss.detach()

    def allreduce_grads(self) -> None:
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        flat = torch.cat([p.grad.reshape(-1) for p in self.params])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(dist.get_world_size())
        off = 0
        for p in self.params:
            k = p.numel()
            p.grad.copy_(flat[off:off + k].view_as(p.grad))
            off += k

    def step(self, audio: Tensor, lengths: Tensor, y: Tensor, y_lengths: Tensor) -> Tensor:
        loss = self.loss_and_grads(audio, lengths, y, y_lengths)
        self.allreduce_grads()
        self.opt.step()
        return loss
