"""Helpers shared by every model, mirroring the public names of ``src/thunder/blocks.py``.

Only what the forward hot path needs is here: the two-input container protocol
(``MultiSequential`` / ``Masked``, blocks.py:94-115), the length/mask helpers (blocks.py:156-196) and
the ``conv1d_decoder`` factory (blocks.py:199-216).
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor, nn

__all__ = ["MultiSequential", "Masked", "lengths_to_mask", "get_same_padding", "conv1d_decoder"]


class MultiSequential(nn.Sequential):
    """``nn.Sequential`` whose children map ``(x, lengths) -> (x, lengths)`` (blocks.py:94-102)."""

    def forward(self, audio: Tensor, audio_lengths: Tensor) -> Tuple[Tensor, Tensor]:
        for module in self.children():
            audio, audio_lengths = module(audio, audio_lengths)
        return audio, audio_lengths


class Masked(nn.Module):
    """Lifts single-input layers into the two-input protocol (blocks.py:105-115).  In this package it
    is mostly a *parameter holder* that keeps the reference's ``state_dict`` names
    (``...layer.0.weight``); the arithmetic of the wrapped layers is fused into the CUDA kernels."""

    def __init__(self, *layers: nn.Module):
        super().__init__()
        self.layer = nn.Sequential(*layers)

    def forward(self, audio: Tensor, audio_lengths: Tensor) -> Tuple[Tensor, Tensor]:
        return self.layer(audio), audio_lengths


def lengths_to_mask(lengths: Tensor, max_length: int) -> Tensor:
    """Boolean mask ``t < lengths[b]`` (blocks.py:156-170).  Host-side utility; the kernels build the
    mask from ``lengths`` on the fly and never materialise it."""
    lengths = lengths.type(torch.long)
    return torch.arange(max_length, device=lengths.device).expand(lengths.shape[0], max_length) < lengths.unsqueeze(1)


def get_same_padding(kernel_size: int, stride: int, dilation: int) -> int:
    """Padding giving ``ceil(T / stride)`` outputs (blocks.py:173-196).  Raises ``ValueError`` when both
    stride and dilation exceed 1, like the reference."""
    if stride > 1 and dilation > 1:
        raise ValueError("Only stride OR dilation may be greater than 1")
    if dilation > 1:
        return (dilation * (kernel_size - 1) + 1) // 2
    return kernel_size // 2


def conv_out_length(length, kernel_size: int, stride: int, padding: int, dilation: int):
    """``MaskedConv1d.get_seq_len`` (quartznet/blocks.py:142-156); works on ints and integer tensors."""
    num = length + 2 * padding - dilation * (kernel_size - 1) - 1
    if isinstance(num, Tensor):
        return torch.div(num, stride, rounding_mode="floor") + 1
    return num // stride + 1


def conv1d_decoder(decoder_input_channels: int, num_classes: int) -> nn.Module:
    """One 1x1 ``Conv1d`` with bias, xavier-uniform init (blocks.py:199-216).  Returned as a plain
    ``nn.Conv1d`` so that ``state_dict`` keys (``weight``, ``bias``) match; the fused model runner reads its
    parameters and runs the decoder GEMM + greedy argmax on the GPU."""
    decoder = nn.Conv1d(decoder_input_channels, num_classes, kernel_size=1, bias=True)
    nn.init.xavier_uniform_(decoder.weight, gain=1.0)
    return decoder
