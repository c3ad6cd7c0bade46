"""Deterministic synthetic inputs and random-init weights (numpy PCG64, torch-version independent).

There are no datasets or checkpoints offline, so every parity test, golden fixture and benchmark
uses audio and weights generated here from an integer seed.  Weights are returned as a flat
``{state_dict key: float32 ndarray}`` using the reference's own key names (SURVEY.md 3.4:
``<block>.mconv.<i>.conv.weight``, ``<block>.mconv.<i>.layer.0.{weight,bias,running_mean,running_var}``,
``<block>.res.0.conv.weight``, ``<block>.res.1.layer.0.*``, Citrinet SE at
``<block>.mconv.<i>.layer.0.fc.{0,2}.weight``), so the same dict loads strictly into the reference
modules, into this package's module shells, and into the numpy oracle.

Initialisation is variance-preserving (He-style) rather than the reference's default
``kaiming_uniform(a=sqrt(5))`` so that signal -- and therefore any kernel error -- actually
propagates through all 15x5 sub-blocks instead of decaying into the BatchNorm shifts; BatchNorm
affine/running statistics are randomised so that BN folding is exercised (a fresh BN is the identity).
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

F32 = np.float32


def audio(batch: int, samples: int, seed: int = 1234, kind: str = "noise") -> np.ndarray:
    """16 kHz synthetic audio ``[batch, samples]`` float32.

    ``noise``: ``0.1 * N(0,1)`` (worst case for STFT error).  ``tones``: sum of 5 random sinusoids
    in 100-7000 Hz plus ``0.01 * N(0,1)`` (speech-like, peaky spectrum)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if kind == "noise":
        return (0.1 * rng.standard_normal((batch, samples))).astype(F32)
    if kind == "tones":
        t = np.arange(samples, dtype=np.float64) / 16000.0
        out = np.zeros((batch, samples), np.float64)
        for b in range(batch):
            f = rng.uniform(100.0, 7000.0, 5)
            a = rng.uniform(0.02, 0.2, 5)
            ph = rng.uniform(0, 2 * np.pi, 5)
            out[b] = (a[:, None] * np.sin(2 * np.pi * f[:, None] * t[None, :] + ph[:, None])).sum(0)
        out += 0.01 * rng.standard_normal((batch, samples))
        return out.astype(F32)
    raise ValueError(kind)


def ragged_lengths(batch: int, samples: int, seed: int = 7) -> np.ndarray:
    """``asr_collate``-style lengths: first = full, rest random in [samples//2, samples], sorted
    descending (src/thunder/data/dataloader_utils.py:17-33)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = rng.integers(samples // 2, samples + 1, batch)
    lens[0] = samples
    return np.sort(lens)[::-1].astype(np.int64).copy()


def _uniform(rng, shape, var):
    b = np.sqrt(3.0 * var)
    return rng.uniform(-b, b, shape).astype(F32)


def _bn(rng, state: Dict[str, np.ndarray], prefix: str, c: int):
    state[prefix + ".weight"] = rng.uniform(0.5, 1.5, c).astype(F32)
    state[prefix + ".bias"] = (0.1 * rng.standard_normal(c)).astype(F32)
    state[prefix + ".running_mean"] = (0.1 * rng.standard_normal(c)).astype(F32)
    state[prefix + ".running_var"] = rng.uniform(0.5, 1.5, c).astype(F32)
    state[prefix + ".num_batches_tracked"] = np.zeros((), np.int64)


def block_state(rng, prefix: str, in_channels: int, out_channels: int, repeat: int, kernel_size: int,
                residual: bool, separable: bool, se: bool = False, se_ratio: int = 8) -> Dict[str, np.ndarray]:
    """Parameters of one Quartznet/Citrinet block under the reference key names."""
    st: Dict[str, np.ndarray] = {}
    cin = in_channels
    per = (3 if separable else 2) + 2
    for r in range(repeat):
        i = r * per
        if separable:
            st[f"{prefix}mconv.{i}.conv.weight"] = _uniform(rng, (cin, 1, kernel_size), 1.0 / kernel_size)
            st[f"{prefix}mconv.{i + 1}.conv.weight"] = _uniform(rng, (out_channels, cin, 1), 2.0 / cin)
            _bn(rng, st, f"{prefix}mconv.{i + 2}.layer.0", out_channels)
        else:
            st[f"{prefix}mconv.{i}.conv.weight"] = _uniform(rng, (out_channels, cin, kernel_size),
                                                            2.0 / (cin * kernel_size))
            _bn(rng, st, f"{prefix}mconv.{i + 1}.layer.0", out_channels)
        cin = out_channels
    if se:
        i_se = (repeat - 1) * per + (3 if separable else 2)
        hid = out_channels // se_ratio
        st[f"{prefix}mconv.{i_se}.layer.0.fc.0.weight"] = _uniform(rng, (hid, out_channels), 2.0 / out_channels)
        st[f"{prefix}mconv.{i_se}.layer.0.fc.2.weight"] = _uniform(rng, (out_channels, hid), 4.0 / hid)
    if residual:
        st[f"{prefix}res.0.conv.weight"] = _uniform(rng, (out_channels, in_channels, 1), 1.0 / in_channels)
        _bn(rng, st, f"{prefix}res.1.layer.0", out_channels)
    return st


def quartznet_block_list(feat_in: int = 64, filters: Sequence[int] = (256, 256, 512, 512, 512),
                         kernel_sizes: Sequence[int] = (33, 39, 51, 63, 75), repeat_blocks: int = 1) -> List[dict]:
    """Constructor arguments of every block of ``QuartznetEncoder``
    (src/thunder/quartznet/blocks.py:341-434)."""
    blocks = [dict(in_channels=feat_in, out_channels=256, repeat=1, kernel_size=33, stride=2, dilation=1,
                   residual=False, separable=True)]
    f_in = 256
    for f, k in zip(filters, kernel_sizes):
        for _ in range(repeat_blocks):
            blocks.append(dict(in_channels=f_in, out_channels=f, repeat=5, kernel_size=k, stride=1, dilation=1,
                               residual=True, separable=True))
            f_in = f
    blocks.append(dict(in_channels=f_in, out_channels=512, repeat=1, kernel_size=87, stride=1, dilation=2,
                       residual=False, separable=True))
    blocks.append(dict(in_channels=512, out_channels=1024, repeat=1, kernel_size=1, stride=1, dilation=1,
                       residual=False, separable=False))
    return blocks


def citrinet_block_list(filters: Sequence[int], kernel_sizes: Sequence[int], strides: Sequence[int],
                        feat_in: int = 80) -> List[dict]:
    """Constructor arguments of every block of ``CitrinetEncoder``
    (src/thunder/citrinet/blocks.py:200-278; stem hard-coded to 256 channels)."""
    blocks = [dict(in_channels=feat_in, out_channels=256, repeat=1, kernel_size=5, stride=1, dilation=1,
                   residual=False, separable=True)]
    f_in = 256
    for f, k, s in zip(filters, kernel_sizes, strides):
        blocks.append(dict(in_channels=f_in, out_channels=f, repeat=5, kernel_size=k, stride=s, dilation=1,
                           residual=True, separable=True))
        f_in = f
    blocks.append(dict(in_channels=f_in, out_channels=640, repeat=1, kernel_size=41, stride=1, dilation=1,
                       residual=False, separable=True))
    return blocks


#: NeMo ``stt_en_citrinet_1024`` body ([external]); the reference reads these from the
#: ``.nemo`` YAML at load time (src/thunder/citrinet/compatibility.py:71-85).
CITRINET_1024 = dict(
    filters=[1024] * 21,
    kernel_sizes=[11, 13, 15, 17, 19, 21, 13, 15, 17, 19, 21, 23, 25, 25, 27, 29, 31, 33, 35, 37, 39],
    strides=[2, 1, 1, 1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1],
)


def encoder_state(blocks: List[dict], seed: int = 0, se: bool = False) -> Dict[str, np.ndarray]:
    rng = np.random.Generator(np.random.PCG64(seed))
    st: Dict[str, np.ndarray] = {}
    for bi, b in enumerate(blocks):
        st.update(block_state(rng, f"{bi}.", b["in_channels"], b["out_channels"], b["repeat"],
                              b["kernel_size"], b["residual"], b["separable"], se=se))
    return st


def decoder_state(in_channels: int, num_classes: int, seed: int = 1) -> Dict[str, np.ndarray]:
    """``conv1d_decoder`` parameters (src/thunder/blocks.py:199-216): ``weight[V,C,1]``, ``bias[V]``."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return {
        "weight": _uniform(rng, (num_classes, in_channels, 1), 1.0 / in_channels),
        "bias": (0.1 * rng.standard_normal(num_classes)).astype(F32),
    }


def quartznet_vocab() -> List[str]:
    """``[" ", a-z, "'"]`` (tests/nemo_config_samples/QuartzNet15x5Base-En.yaml:229-257); the blank is
    appended by the vocabulary => V = 29."""
    return [" "] + [chr(c) for c in range(ord("a"), ord("z") + 1)] + ["'"]


def citrinet_vocab(n: int = 1024) -> List[str]:
    """1024 dummy sentencepiece-like tokens (+ blank => V = 1025)."""
    return [("▁" if i % 3 == 0 else "") + f"t{i}" for i in range(n)]
