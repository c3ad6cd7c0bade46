"""thunder_speech_b200: B200-native (sm_100a) implementation of the thunder-speech ASR forward hot path.

Public surface mirrors the reference's module names for this path only:

    thunder_speech_b200.quartznet.transform.FilterbankFeatures
    thunder_speech_b200.quartznet.blocks.{MaskedConv1d, QuartznetBlock, QuartznetEncoder}
    thunder_speech_b200.citrinet.blocks.{SqueezeExcite, CitrinetBlock, CitrinetEncoder}
    thunder_speech_b200.blocks.{MultiSequential, Masked, lengths_to_mask, get_same_padding, conv1d_decoder}
    thunder_speech_b200.module.CTCModule (= BaseCTCModule forward / predict)
    thunder_speech_b200.text_processing.{Vocabulary, BatchTextTransformer}

All arithmetic runs in ``libthunder_b200.so`` (hand-written CUDA, C ABI in include/thunder_b200.h) through the
``torch.ops.thunder_b200.*`` custom ops; there is no CPU or eager fallback.
"""
__version__ = "0.2.0"

import os as _os

#: storage format of the activation rows / GEMM operands of the INFERENCE path: "bf16" (default; the reference's autocast
#: type, 8-bit mantissa, fp32 range) or "fp16" (IEEE half: 11-bit mantissa, conversions saturate at +-65504).  Same kernels,
#: same bytes, same tensor-core rate (tcgen05 kind::f16 takes either); fp16 rows are what holds the 2e-2 logit parity
#: against the fp32 reference through the 90 convolutions of QuartzNet 15x5.  The training step
#: always runs bf16.
_PRECISIONS = ("bf16", "fp16")
_default_precision = _os.environ.get("THUNDER_B200_PRECISION", "bf16")
if _default_precision not in _PRECISIONS:
    raise ValueError(f"THUNDER_B200_PRECISION must be one of {_PRECISIONS}, got {_default_precision!r}")


def set_default_precision(precision: str) -> None:
    """Row format used by modules that were not given one explicitly (``CTCModule.set_precision``)."""
    global _default_precision
    if precision not in _PRECISIONS:
        raise ValueError(f"precision must be one of {_PRECISIONS}, got {precision!r}")
    _default_precision = precision


def get_default_precision() -> str:
    return _default_precision


def row_dtype(precision=None):
    """torch dtype of the activation rows for ``precision`` (None = the package default)."""
    import torch

    precision = precision or _default_precision
    if precision not in _PRECISIONS:
        raise ValueError(f"precision must be one of {_PRECISIONS}, got {precision!r}")
    return torch.float16 if precision == "fp16" else torch.bfloat16


#: STFT kernel of the feature front-end: "dft" = DFT-matrix contraction on the tensor cores (tcgen05, split-fp16 operands;
#: n_fft 512 with the window supported on [96, 416), float audio within [-255, 255]) with automatic fall-back to "fft" = the
#: shared-memory radix-8 FFT (any window / filter bank / amplitude).  Both hold the 1e-4 feature parity; see DESIGN.md for
#: the ncu A/B that picked the default.
_STFT_KERNELS = ("fft", "dft")
_stft_kernel = _os.environ.get("THUNDER_B200_STFT", "fft")
if _stft_kernel not in _STFT_KERNELS:
    raise ValueError(f"THUNDER_B200_STFT must be one of {_STFT_KERNELS}, got {_stft_kernel!r}")


def set_stft_kernel(kind: str) -> None:
    global _stft_kernel
    if kind not in _STFT_KERNELS:
        raise ValueError(f"stft kernel must be one of {_STFT_KERNELS}, got {kind!r}")
    _stft_kernel = kind


def get_stft_kernel() -> str:
    return _stft_kernel
