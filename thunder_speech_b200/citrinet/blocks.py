"""Citrinet building blocks (mirrors the public names of ``src/thunder/citrinet/blocks.py``)."""
from __future__ import annotations

from typing import List

from torch import Tensor, nn

from ..blocks import Masked, MultiSequential, get_same_padding
from ..fused import PlannedBlock
from ..quartznet.blocks import EncoderBase, _first, _get_act_dropout_layer, _get_conv_bn_layer

__all__ = ["SqueezeExcite", "CitrinetBlock", "stem", "body", "CitrinetEncoder"]


class SqueezeExcite(nn.Module):
    """``x * sigmoid(W2 relu(W1 mean_t(x)))``, no biases (citrinet/blocks.py:48-83).  Holder of ``fc.0.weight``
    / ``fc.2.weight``; inside a block the pooling is fused into the last pointwise GEMM's epilogue, the two
    FCs run in ``se_fc`` and the scale is applied in the residual GEMM's epilogue."""

    def __init__(self, channels: int, reduction_ratio: int):
        super().__init__()
        self.pool = nn.AdaptiveAvgPool1d(1)
        self.fc = nn.Sequential(
            nn.Linear(channels, channels // reduction_ratio, bias=False),
            nn.ReLU(True),
            nn.Linear(channels // reduction_ratio, channels, bias=False),
        )

    def forward(self, x: Tensor) -> Tensor:
        """Stand-alone use on ``[B, C, T]``: pooled sums with torch (plumbing), FC + scale on the kernels."""
        import torch

        from .. import ops

        with torch.no_grad():
            T = x.shape[-1]
            pool = x.float().sum(-1).contiguous()
            gate = ops.se_fc(pool, T, self.fc[0].weight.detach().float().contiguous(),
                             self.fc[2].weight.detach().float().contiguous())
            from .. import row_dtype

            rows = ops.pack_rows(x, None, row_dtype() == torch.float16)
            return ops.unpack_rows(ops.se_apply(rows, gate, None, False), T)


class CitrinetBlock(PlannedBlock):
    """Like ``QuartznetBlock`` but only the LAST sub-block is strided, SqueezeExcite follows the last BN, and
    the residual stride equals the block stride (citrinet/blocks.py:86-197)."""

    def __init__(self, in_channels: int, out_channels: int, repeat: int = 5, kernel_size=(11,), stride=(1,),
                 dilation=(1,), dropout: float = 0.0, residual: bool = True, separable: bool = False):
        super().__init__()
        k, s, d = _first(kernel_size), _first(stride), _first(dilation)
        self.separable = separable
        padding_val = get_same_padding(k, 1, d)
        inplanes_loop = in_channels
        conv: List[nn.Module] = []
        for _ in range(repeat - 1):
            conv.extend(_get_conv_bn_layer(inplanes_loop, out_channels, kernel_size=k, stride=1, dilation=d,
                                           padding=padding_val, separable=separable, bias=False))
            conv.extend(_get_act_dropout_layer(drop_prob=dropout))
            inplanes_loop = out_channels
        padding_val = get_same_padding(k, s, d)
        conv.extend(_get_conv_bn_layer(inplanes_loop, out_channels, kernel_size=k, stride=s, dilation=d,
                                       padding=padding_val, separable=separable, bias=False))
        conv.append(Masked(SqueezeExcite(out_channels, reduction_ratio=8)))
        self.mconv = MultiSequential(*conv)
        if residual:
            self.res = MultiSequential(*_get_conv_bn_layer(in_channels, out_channels, kernel_size=1, stride=s,
                                                           bias=False))
        else:
            self.res = None
        self.mout = MultiSequential(*_get_act_dropout_layer(drop_prob=dropout))


def stem(feat_in: int) -> CitrinetBlock:
    """``feat_in -> 256`` (hard-coded), K=5, no residual (citrinet/blocks.py:200-216)."""
    return CitrinetBlock(feat_in, 256, repeat=1, kernel_size=(5,), residual=False, separable=True)


def body(filters: List[int], kernel_size: List[int], strides: List[int], dropout: float = 0.0
         ) -> List[CitrinetBlock]:
    """Body blocks starting from 256 channels + ``-> 640, K=41`` epilogue (citrinet/blocks.py:219-255)."""
    layers = []
    f_in = 256
    for f, k, s in zip(filters, kernel_size, strides):
        layers.append(CitrinetBlock(f_in, f, kernel_size=(k,), stride=(s,), separable=True, dropout=dropout))
        f_in = f
    layers.append(CitrinetBlock(f_in, 640, repeat=1, kernel_size=(41,), residual=False, separable=True,
                                dropout=dropout))
    return layers


def CitrinetEncoder(filters: List[int], kernel_sizes: List[int], strides: List[int], feat_in: int = 80,
                    dropout: float = 0.0) -> nn.Module:
    """Citrinet encoder (citrinet/blocks.py:258-278)."""
    return EncoderBase(stem(feat_in), *body(filters, kernel_sizes, strides, dropout))
