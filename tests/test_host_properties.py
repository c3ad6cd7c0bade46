"""Property tests (hypothesis) of the host-side arithmetic that every kernel launch depends on."""
import math

import numpy as np
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import ref_numpy as R
from thunder_speech_b200 import _lib
from thunder_speech_b200.blocks import conv_out_length, get_same_padding
from thunder_speech_b200.data import compress_taps, sinc_resample_taps
from thunder_speech_b200.parallel import shard_bounds


@settings(max_examples=200, deadline=None)
@given(T=st.integers(1, 400), K=st.integers(1, 41), S=st.integers(1, 3), D=st.integers(1, 3), P=st.integers(0, 60))
def test_conv_out_length_matches_torch_conv1d(T, K, S, D, P):
    if T + 2 * P < D * (K - 1) + 1:
        return
    y = torch.nn.functional.conv1d(torch.zeros(1, 1, T), torch.zeros(1, 1, K), None, S, P, D)
    assert conv_out_length(T, K, S, P, D) == y.shape[-1]


@settings(max_examples=100, deadline=None)
@given(K=st.integers(1, 99).filter(lambda k: k % 2 == 1), D=st.integers(1, 4), T=st.integers(1, 300))
def test_same_padding_preserves_length_for_stride_1(K, D, T):
    P = get_same_padding(K, 1, D)
    assert P == R.get_same_padding(K, 1, D)
    assert conv_out_length(T, K, 1, P, D) == T


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 1000), world=st.integers(1, 16))
def test_shard_bounds_partition_the_batch(n, world):
    spans = [shard_bounds(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1


@settings(max_examples=60, deadline=None)
@given(orig=st.sampled_from([8000, 11025, 22050, 32000, 44100, 48000, 16000, 24000]),
       new=st.sampled_from([8000, 16000, 22050, 44100]))
def test_compressed_taps_reconstruct_the_full_kernel(orig, new):
    if orig == new:
        return
    k, width, o, n = sinc_resample_taps(orig, new)
    assert k.shape == (n, 2 * width + o)
    kr, _, _, _ = R.sinc_resample_kernel(orig, new)
    assert np.array_equal(k, kr)
    tc, k0 = compress_taps(k)
    full = np.zeros_like(k)
    for p in range(n):
        seg = tc[:, p]
        m = min(seg.size, k.shape[1] - k0[p])
        full[p, k0[p]: k0[p] + m] = seg[:m]
    assert np.abs(full - k).max() <= 1e-30          # only exact (window-clamped) zeros were dropped
    assert tc.shape[0] <= 2 * width + 2 or tc.shape[0] == k.shape[1]


@settings(max_examples=100, deadline=None)
@given(T=st.integers(1, 100000))
def test_row_pitch_is_the_next_multiple_of_64(T):
    p = _lib.row_pitch(T)
    assert p % 64 == 0 and p >= T and p - T < 64


@settings(max_examples=50, deadline=None)
@given(length=st.integers(1, 3000), orig=st.sampled_from([8000, 22050, 44100, 48000]))
def test_resampled_length_formula(length, orig):
    g = math.gcd(orig, 16000)
    o, n = orig // g, 16000 // g
    x = np.zeros((1, length), np.float32)
    assert R.resample(x, orig, 16000).shape[-1] == -((-n * length) // o)
