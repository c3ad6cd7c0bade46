"""QuartzNet building blocks (mirrors the public names of ``src/thunder/quartznet/blocks.py``).

The classes keep the reference's constructor signatures, attribute layout and therefore ``state_dict``
keys (``mconv.<i>.conv.weight``, ``mconv.<i>.layer.0.*``, ``res.0.conv.weight``, ``res.1.layer.0.*``), so
checkpoints converted by the reference's loaders (quartznet/compatibility.py:127-158) load strictly.
The modules are parameter holders; ``forward`` runs the CUDA plan of ``thunder_speech_b200.fused``.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import Tensor, nn

from .. import ops
from ..blocks import Masked, MultiSequential, conv_out_length, get_same_padding
from ..fused import PlannedBlock

__all__ = ["MaskedConv1d", "QuartznetBlock", "stem", "body", "QuartznetEncoder", "EncoderBase"]


class MaskedConv1d(nn.Module):
    """``nn.Conv1d`` preceded by length masking (quartznet/blocks.py:93-182)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size, stride=1, padding=0, dilation=1,
                 groups: int = 1, bias: bool = False, use_mask: bool = True):
        super().__init__()
        self.use_mask = use_mask
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                              dilation=dilation, groups=groups, bias=bias)
        self.padding = self.conv.padding[0]
        self.dilation = self.conv.dilation[0]
        self.kernel_size = self.conv.kernel_size[0]
        self.stride = self.conv.stride[0]

    def get_seq_len(self, lengths: Tensor) -> Tensor:
        return conv_out_length(lengths, self.kernel_size, self.stride, self.padding, self.dilation)

    def forward(self, x: Tensor, lengths: Tensor) -> Tuple[Tensor, Tensor]:
        """Stand-alone use (inside blocks the conv is fused into the block plan).  Depthwise and 1x1 convolutions
        are implemented; output frames are NOT masked, like the reference."""
        conv = self.conv
        with torch.no_grad():
            l32 = ops.lengths_i32(lengths) if self.use_mask else None
            from .. import row_dtype

            T = x.shape[-1]
            rows = ops.pack_rows(x, l32, row_dtype() == torch.float16)
            if conv.groups == conv.in_channels and conv.out_channels == conv.in_channels and conv.groups > 1:
                w = conv.weight.detach().float()[:, 0, :].contiguous()
                # the kernel masks its output for the following pointwise conv; stand-alone semantics are
                # unmasked outputs, so run without lengths on the already masked input
                y = ops.dw_conv(rows, T, w, self.stride, self.dilation, self.padding, None)
                T_out = conv_out_length(T, self.kernel_size, self.stride, self.padding, self.dilation)
                out = ops.unpack_rows(y, T_out)
                if conv.bias is not None:
                    out = out + conv.bias.detach().float()[None, :, None]
            elif self.kernel_size == 1 and self.stride == 1 and conv.groups == 1 and self.padding == 0:
                w = conv.weight.detach()[:, :, 0].to(rows.dtype).contiguous()
                bias = conv.bias.detach().float().contiguous() if conv.bias is not None else None
                out = ops.pw_gemm(w, rows, None, None, T, bias, None, True, False, None, None, None)
            elif conv.groups == 1:     # full convolution: im2col rows + one GEMM over Cin * K channels
                cols = ops.im2col_rows(rows, T, self.kernel_size, self.stride, self.dilation, self.padding, None)
                w = conv.weight.detach().float().reshape(conv.out_channels, -1)
                if w.shape[1] != cols.shape[1]:
                    w = torch.nn.functional.pad(w, (0, cols.shape[1] - w.shape[1]))
                bias = conv.bias.detach().float().contiguous() if conv.bias is not None else None
                T_out = conv_out_length(T, self.kernel_size, self.stride, self.padding, self.dilation)
                out = ops.pw_gemm(w.to(rows.dtype).contiguous(), cols, None, None, T_out, bias, None, True, False, None, None,
                                  None)
            else:
                raise NotImplementedError("stand-alone MaskedConv1d supports depthwise, 1x1 and ungrouped convolutions")
        return out, self.get_seq_len(lengths)


def _get_conv_bn_layer(in_channels: int, out_channels: int, kernel_size=11, separable: bool = False,
                       **conv_kwargs) -> List[nn.Module]:
    """[depthwise, pointwise, BN] or [conv, BN] -- no BN/ReLU between depthwise and pointwise
    (quartznet/blocks.py:185-224)."""
    if separable:
        layers = [
            MaskedConv1d(in_channels, in_channels, kernel_size, groups=in_channels, **conv_kwargs),
            MaskedConv1d(in_channels, out_channels, kernel_size=1, stride=1, dilation=1, padding=0,
                         bias=conv_kwargs.get("bias", False)),
        ]
    else:
        layers = [MaskedConv1d(in_channels, out_channels, kernel_size, **conv_kwargs)]
    layers.append(Masked(nn.BatchNorm1d(out_channels, eps=1e-3, momentum=0.1)))
    return layers


def _get_act_dropout_layer(drop_prob: float = 0.2) -> List[nn.Module]:
    return [Masked(nn.ReLU(True)), Masked(nn.Dropout(p=drop_prob))]


def _first(v) -> int:
    return int(v[0]) if isinstance(v, (tuple, list)) else int(v)


class QuartznetBlock(PlannedBlock):
    """``repeat x (dw, pw, BN[, ReLU, Dropout])`` + ``BN(conv1x1(x))`` residual + ReLU
    (quartznet/blocks.py:231-338).  Every sub-block is strided; the residual stride is ``stride**repeat``."""

    def __init__(self, in_channels: int, out_channels: int, repeat: int = 5, kernel_size=(11,), stride=(1,),
                 dilation=(1,), dropout: float = 0.0, residual: bool = True, separable: bool = False):
        super().__init__()
        k, s, d = _first(kernel_size), _first(stride), _first(dilation)
        padding_val = get_same_padding(k, s, d)
        self.separable = separable
        inplanes_loop = in_channels
        conv: List[nn.Module] = []
        for _ in range(repeat - 1):
            conv.extend(_get_conv_bn_layer(inplanes_loop, out_channels, kernel_size=k, stride=s, dilation=d,
                                           padding=padding_val, separable=separable, bias=False))
            conv.extend(_get_act_dropout_layer(drop_prob=dropout))
            inplanes_loop = out_channels
        conv.extend(_get_conv_bn_layer(inplanes_loop, out_channels, kernel_size=k, stride=s, dilation=d,
                                       padding=padding_val, separable=separable, bias=False))
        self.mconv = MultiSequential(*conv)
        if residual:
            stride_residual = s if s == 1 else s ** repeat
            self.res = MultiSequential(*_get_conv_bn_layer(in_channels, out_channels, kernel_size=1,
                                                           stride=stride_residual, bias=False))
        else:
            self.res = None
        self.mout = MultiSequential(*_get_act_dropout_layer(drop_prob=dropout))


def stem(feat_in: int) -> QuartznetBlock:
    """First block: ``feat_in -> 256``, K=33, stride 2, no residual (quartznet/blocks.py:341-358)."""
    return QuartznetBlock(feat_in, 256, repeat=1, stride=(2,), kernel_size=(33,), residual=False, separable=True)


def body(filters: List[int], kernel_size: List[int], repeat_blocks: int = 1, dropout: float = 0.0
         ) -> List[QuartznetBlock]:
    """Middle blocks + ``512, K=87, dilation 2`` + plain ``512 -> 1024, K=1`` (quartznet/blocks.py:361-410)."""
    layers = []
    f_in = 256
    for f, k in zip(filters, kernel_size):
        for _ in range(repeat_blocks):
            layers.append(QuartznetBlock(f_in, f, kernel_size=(k,), separable=True, dropout=dropout))
            f_in = f
    layers.extend([
        QuartznetBlock(f_in, 512, repeat=1, dilation=(2,), kernel_size=(87,), residual=False, separable=True,
                       dropout=dropout),
        QuartznetBlock(512, 1024, repeat=1, kernel_size=(1,), residual=False, separable=False, dropout=dropout),
    ])
    return layers


class EncoderBase(MultiSequential):
    """``MultiSequential`` of planned blocks that keeps activations in the kernels' bf16 row layout between
    blocks (one pack at the input, one unpack at the output)."""

    def forward_rows(self, rows: Tensor, T: int, lens: Optional[Tensor]):
        """bf16 rows in (zero beyond ``lens``) -> bf16 rows out.  Inter-block outputs are stored with their tail
        zeroed (every consumer is a MaskedConv1d); the final block's output is left unmasked, like the
        reference, because the decoder is a plain Conv1d."""
        blocks = list(self.children())
        # SqueezeExcite pools of all blocks: ONE zero-fill per forward instead of one per block (23 on Citrinet-1024)
        B = rows.shape[0]
        se_c = [blk.se_channels(rows.dtype) for blk in blocks]
        pools = torch.zeros((B * sum(se_c),), device=rows.device, dtype=torch.int64) if any(se_c) else None
        off = 0
        for i, blk in enumerate(blocks):
            pool = None
            if se_c[i]:
                pool = pools[off:off + B * se_c[i]].view(B, se_c[i])
                off += B * se_c[i]
            rows, T, lens = blk.forward_rows(rows, T, lens, zero_tail=(i != len(blocks) - 1), pool=pool)
        return rows, T, lens

    def out_lengths(self, lengths: Tensor) -> Tensor:
        for blk in self.children():
            lengths = blk.out_lengths(lengths)
        return lengths

    def forward(self, x: Tensor, lengths: Tensor) -> Tuple[Tensor, Tensor]:
        with torch.no_grad():
            from .. import row_dtype

            l32 = ops.lengths_i32(lengths)
            rows = ops.pack_rows(x, l32, row_dtype() == torch.float16)
            y, T_out, _ = self.forward_rows(rows, x.shape[-1], l32)
            return ops.unpack_rows(y, T_out), self.out_lengths(lengths)


def QuartznetEncoder(feat_in: int = 64, filters: List[int] = [256, 256, 512, 512, 512],
                     kernel_sizes: List[int] = [33, 39, 51, 63, 75], repeat_blocks: int = 1,
                     dropout: float = 0.0) -> nn.Module:
    """Quartznet5x5 (``repeat_blocks=1``) / Quartznet15x5 (``repeat_blocks=3``) encoder
    (quartznet/blocks.py:413-434)."""
    return EncoderBase(stem(feat_in), *body(filters, kernel_sizes, repeat_blocks, dropout))
