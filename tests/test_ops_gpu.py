"""GPU parity tests of the individual encoder / decoder ops through the C ABI, against the numpy oracle.

Inputs are rounded to bf16 first (the kernels' storage type), the oracle then computes in float64, so the
remaining difference is fp32 accumulation order plus ONE bf16 rounding of the output: tolerance 2^-8 relative
to the largest magnitude (north_star: 2e-2 for bf16); f32 outputs are held to 1e-4."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import ref_numpy as R
from thunder_speech_b200 import ops

pytestmark = pytest.mark.gpu
BF16_TOL = 2.0 ** -8


def bf16_round(a):
    return torch.from_numpy(np.asarray(a, np.float32)).bfloat16().float().numpy()


def to_rows(a):
    """numpy [B,C,T] f32 -> device bf16 rows"""
    return ops.pack_rows(torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda())


def from_rows(rows, T):
    return ops.unpack_rows(rows, T).cpu().numpy()


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 5, 77)).astype(np.float32)
    rows = to_rows(x)
    assert rows.shape == (3, 5, 128) and rows.dtype == torch.bfloat16
    assert np.array_equal(from_rows(rows, 77), bf16_round(x))
    assert (rows[:, :, 77:].float() == 0).all()
    xb = torch.from_numpy(x).cuda().bfloat16()
    assert torch.equal(ops.pack_rows(xb)[:, :, :77], xb)


DW_FAST = [(33, 2, 1), (33, 1, 1), (39, 1, 1), (51, 1, 1), (63, 1, 1), (75, 1, 1), (87, 1, 2), (5, 1, 1), (11, 1, 1),
           (13, 1, 1), (15, 1, 1), (17, 1, 1), (19, 1, 1), (21, 1, 1), (23, 1, 1), (25, 1, 1), (27, 1, 1), (29, 1, 1),
           (31, 1, 1), (35, 1, 1), (37, 1, 1), (41, 1, 1), (11, 2, 1), (13, 2, 1), (25, 2, 1)]
DW_GENERIC = [(3, 1, 1, 1), (3, 2, 1, 1), (7, 1, 3, 9), (9, 1, 1, 0), (4, 1, 1, 2), (1, 1, 1, 0)]


def run_dw(x, w, K, S, D, P, lens, premasked=False):
    B, C, T = x.shape
    l32 = None if lens is None else torch.from_numpy(lens.astype(np.int32)).cuda()
    xr = ops.pack_rows(torch.from_numpy(x).cuda(), l32) if premasked else to_rows(x)
    y = ops.dw_conv(xr, T, torch.from_numpy(w).cuda(), S, D, P, l32, premasked)
    T_out = (T + 2 * P - D * (K - 1) - 1) // S + 1
    torch.cuda.synchronize()
    return from_rows(y, T_out), y


def oracle_dw(x, w, K, S, D, P, lens, w_bf16=None):
    B, C, T = x.shape
    xb = bf16_round(x)
    if w_bf16 is None:
        w_bf16 = S == 1 and D == 1 and K % 2 == 1 and P == K // 2
    if w_bf16:
        w = bf16_round(w)   # the tensor-core (Toeplitz MMA) paths hold the taps in bf16; the SIMT paths keep fp32
    if lens is None:
        lens = np.full((B,), T)
    y, yl = R.masked_conv1d(xb, lens, w[:, None, :], S, P, D, groups=C)
    m = R.lengths_to_mask(yl, y.shape[-1])[:, None, :]   # the mask the next (pointwise) MaskedConv1d applies
    return np.where(m, y, 0).astype(np.float32)


@pytest.mark.parametrize("K,S,D", DW_FAST)
@pytest.mark.parametrize("T", [751, 300, 2001])
def test_dw_conv_fast_paths(K, S, D, T):
    rng = np.random.default_rng(K * 7 + S + D + T)
    B, C = 3, 5
    P = R.get_same_padding(K, S, D)
    x = rng.standard_normal((B, C, T)).astype(np.float32)
    w = rng.uniform(-0.3, 0.3, (C, K)).astype(np.float32)
    lens = np.array([T, T * 2 // 3, 1], np.int64)
    for ln, pre in ((None, False), (lens, False), (lens, True)):   # unmasked / kernel-masked / caller-masked (TMA path)
        got, rows = run_dw(x, w, K, S, D, P, ln, pre)
        # dilated "same" convs run on the TMA Toeplitz kernel (bf16 taps) when the input is pre-masked / unmasked
        tma = S == 1 and 2 * P == D * (K - 1) and (ln is None or pre)
        ref = oracle_dw(x, w, K, S, D, P, ln, w_bf16=True if (tma and D > 1) else None)
        assert got.shape == ref.shape
        emax, _ = rel_err(got, ref)
        assert emax < BF16_TOL, (K, S, D, T, pre, emax)
        assert (rows[:, :, ref.shape[-1]:].float() == 0).all()      # pad frames are zero


@pytest.mark.parametrize("K,S,D,P", DW_GENERIC)
def test_dw_conv_generic_path(K, S, D, P):
    rng = np.random.default_rng(K + 10 * S + 100 * D)
    B, C, T = 2, 7, 91
    x = rng.standard_normal((B, C, T)).astype(np.float32)
    w = rng.uniform(-0.5, 0.5, (C, K)).astype(np.float32)
    lens = np.array([T, 40], np.int64)
    for ln in (None, lens):
        got, _ = run_dw(x, w, K, S, D, P, ln)
        ref = oracle_dw(x, w, K, S, D, P, ln)
        emax, _ = rel_err(got, ref)
        assert got.shape == ref.shape and emax < BF16_TOL, (K, S, D, P, emax)


def test_dw_conv_errors():
    x = torch.zeros((1, 4, 64), device="cuda", dtype=torch.bfloat16)
    w = torch.zeros((4, 3), device="cuda")
    with pytest.raises(ValueError):   # stride and dilation both > 1 (blocks.py:192-193)
        ops.dw_conv(x, 50, w, 2, 2, 1, None)


def gemm_ref(w0, x0, w1=None, x1=None, shift=None, relu=False, lens=None):
    y = np.einsum("mk,bkt->bmt", bf16_round(w0).astype(np.float64), bf16_round(x0).astype(np.float64))
    if w1 is not None:
        y += np.einsum("mk,bkt->bmt", bf16_round(w1).astype(np.float64), bf16_round(x1).astype(np.float64))
    if shift is not None:
        y += shift[None, :, None]
    if relu:
        y = np.maximum(y, 0)
    if lens is not None:
        y = np.where(R.lengths_to_mask(lens, y.shape[-1])[:, None, :], y, 0)
    return y.astype(np.float32)


def dev(a, dtype=torch.float32):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda().to(dtype)


GEMM_SHAPES = [
    # Cout, Cin, T, B
    (128, 64, 128, 1), (128, 64, 64, 1), (256, 256, 751, 2), (512, 256, 300, 2), (256, 512, 129, 3),
    (1024, 512, 70, 1), (640, 1024, 251, 2), (256, 80, 97, 2), (29, 1024, 100, 2), (1025, 640, 63, 1),
    (24, 16, 50, 3), (16, 8, 33, 1),
]


@pytest.mark.parametrize("Cout,Cin,T,B", GEMM_SHAPES)
def test_pw_gemm_plain(Cout, Cin, T, B):
    rng = np.random.default_rng(Cout + Cin + T)
    w = (rng.standard_normal((Cout, Cin)) / np.sqrt(Cin)).astype(np.float32)
    x = rng.standard_normal((B, Cin, T)).astype(np.float32)
    shift = rng.standard_normal(Cout).astype(np.float32)
    ref = gemm_ref(w, x, shift=shift)
    out = ops.pw_gemm(dev(w, torch.bfloat16), to_rows(x), None, None, T, dev(shift), None, False, False, None, None,
                      None)
    got = from_rows(out, T)
    emax, el2 = rel_err(got, ref)
    assert emax < BF16_TOL and el2 < BF16_TOL, (emax, el2)
    # f32 output (decoder mode): only fp32 accumulation-order error remains
    out32 = ops.pw_gemm(dev(w, torch.bfloat16), to_rows(x), None, None, T, dev(shift), None, True, False, None, None,
                        None)
    assert out32.shape == (B, Cout, T) and out32.dtype == torch.float32
    emax, el2 = rel_err(out32.cpu().numpy(), ref)
    assert emax < 1e-4 and el2 < 1e-4, (emax, el2)


def test_pw_gemm_residual_relu_mask():
    rng = np.random.default_rng(5)
    B, C0, C1, Cout, T = 3, 256, 128, 256, 333
    w0 = (rng.standard_normal((Cout, C0)) / 16).astype(np.float32)
    w1 = (rng.standard_normal((Cout, C1)) / 11).astype(np.float32)
    x0 = rng.standard_normal((B, C0, T)).astype(np.float32)
    x1 = rng.standard_normal((B, C1, T)).astype(np.float32)
    shift = rng.standard_normal(Cout).astype(np.float32)
    lens = np.array([333, 200, 0], np.int64)
    ref = gemm_ref(w0, x0, w1, x1, shift, True, lens)
    out = ops.pw_gemm(dev(w0, torch.bfloat16), to_rows(x0), dev(w1, torch.bfloat16), to_rows(x1), T, dev(shift),
                      dev(lens, torch.int32), True, True, None, None, None)
    emax, el2 = rel_err(out.cpu().numpy(), ref)
    assert emax < 1e-4 and el2 < 1e-4, (emax, el2)
    outb = ops.pw_gemm(dev(w0, torch.bfloat16), to_rows(x0), dev(w1, torch.bfloat16), to_rows(x1), T, dev(shift),
                       dev(lens, torch.int32), False, True, None, None, None)
    emax, _ = rel_err(from_rows(outb, T), ref)
    assert emax < BF16_TOL


def test_pw_gemm_squeeze_excite_epilogues():
    rng = np.random.default_rng(6)
    B, C, T = 2, 256, 140
    w0 = (rng.standard_normal((C, C)) / 16).astype(np.float32)
    x0 = rng.standard_normal((B, C, T)).astype(np.float32)
    shift = rng.standard_normal(C).astype(np.float32)
    # squeeze: pooled sums of (acc + shift) over ALL T frames (citrinet/blocks.py:77, no mask)
    pool = torch.zeros((B, C), device="cuda", dtype=torch.int64)     # fixed-point sums (2^-32 units), integer atomics
    y1 = ops.pw_gemm(dev(w0, torch.bfloat16), to_rows(x0), None, None, T, dev(shift), None, False, False, pool, None,
                     None)
    ref1 = gemm_ref(w0, x0, shift=shift)
    emax, _ = rel_err(ops.se_pool_to_float(pool).cpu().numpy(), ref1.sum(-1))
    assert emax < 1e-4
    with pytest.raises(TypeError):
        ops.pw_gemm(dev(w0, torch.bfloat16), to_rows(x0), None, None, T, dev(shift), None, False, False,
                    torch.zeros((B, C), device="cuda"), None, None)
    # order-independent accumulation: a long utterance (many frame tiles per channel) pools to the same bits every time, and
    # the excitation computed from the fixed-point sums equals the one computed from their float value
    xl = rng.standard_normal((3, C, 2001)).astype(np.float32)
    pools = []
    for _ in range(3):
        pl = torch.zeros((3, C), device="cuda", dtype=torch.int64)
        ops.pw_gemm(dev(w0, torch.bfloat16), to_rows(xl), None, None, 2001, dev(shift), None, False, False, pl, None, None)
        pools.append(pl)
    assert torch.equal(pools[0], pools[1]) and torch.equal(pools[0], pools[2])
    emax, _ = rel_err(ops.se_pool_to_float(pools[0]).cpu().numpy(), gemm_ref(w0, xl, shift=shift).sum(-1))
    assert emax < 1e-4
    fw1 = dev((rng.standard_normal((32, C)) / 16).astype(np.float32))
    fw2 = dev((rng.standard_normal((C, 32)) / 6).astype(np.float32))
    g_fix = ops.se_fc(pools[0], 2001, fw1, fw2)
    g_f32 = ops.se_fc(ops.se_pool_to_float(pools[0]), 2001, fw1, fw2)
    assert (g_fix - g_f32).abs().max() < 1e-6
    # excite: out = relu(acc_res + shift_res + gate * y1)
    wr = (rng.standard_normal((C, C)) / 16).astype(np.float32)
    xr = rng.standard_normal((B, C, T)).astype(np.float32)
    sr = rng.standard_normal(C).astype(np.float32)
    gate = rng.uniform(0.1, 0.9, (B, C)).astype(np.float32)
    out = ops.pw_gemm(dev(wr, torch.bfloat16), to_rows(xr), None, None, T, dev(sr), None, True, True, None, dev(gate),
                      y1)
    y1f = from_rows(y1, T).astype(np.float64)
    ref = np.maximum(gemm_ref(wr, xr, shift=sr).astype(np.float64) + gate[:, :, None] * y1f, 0)
    emax, el2 = rel_err(out.cpu().numpy(), ref)
    assert emax < 1e-4 and el2 < 1e-4, (emax, el2)


def test_pw_gemm_weights_in_tensor_memory_variant_is_bit_identical():
    """Option pw_ws: the weight-stationary pair GEMM (csrc/pwgemm4.cu: weights copied once into tensor memory, tcgen05.mma
    with A from TMEM, 256 x 128 tiles) for K <= 512 -- same k order, same epilogue, so the same bits as the streaming pair
    kernel (the SE pool agrees to fp32 rounding): plain / residual segment + ReLU + tail mask / SE pool / SE gate, ragged channel counts, both row formats, and the
    const-weights (copy before griddepcontrol.wait) and produced-weights orders."""
    from thunder_speech_b200 import _lib

    rng = np.random.default_rng(44)
    B, T = 5, 700
    lens = dev(np.array([700, 512, 129, 64, 0], np.int64), torch.int32)
    cases = [(256, 256, 0), (512, 384, 128), (264, 72, 0), (1024, 512, 0), (384, 256, 256)]
    try:
        for dt in (torch.bfloat16, torch.float16):
            for Cout, c0, c1 in cases:
                w0 = dev((rng.standard_normal((Cout, c0)) / np.sqrt(c0)).astype(np.float32), dt)
                x0 = to_rows(rng.standard_normal((B, c0, T)).astype(np.float32)).to(dt)
                w1 = dev((rng.standard_normal((Cout, c1)) / np.sqrt(c1)).astype(np.float32), dt) if c1 else None
                x1 = to_rows(rng.standard_normal((B, c1, T)).astype(np.float32)).to(dt) if c1 else None
                shift = dev(rng.standard_normal(Cout).astype(np.float32))
                gate = dev(rng.uniform(0.1, 0.9, (B, Cout)).astype(np.float32))
                outs = {}
                for ws in (0, 1):
                    _lib.set_option("pw_ws", ws)
                    pool = torch.zeros((B, Cout), device="cuda", dtype=torch.int64)
                    y = ops.pw_gemm(w0, x0, w1, x1, T, shift, lens, False, True, None, None, None, ws == 1)
                    y1 = ops.pw_gemm(w0, x0, None, None, T, shift, None, False, False, pool, None, None)
                    z = ops.pw_gemm(w0, x0, None, None, T, shift, lens, False, True, None, gate, y1, True)
                    outs[ws] = (y, y1, pool, z)
                torch.cuda.synchronize()
                for i, (a, b) in enumerate(zip(outs[0], outs[1])):
                    if i == 2:   # SE pool: per-thread fp32 partial sums cover 128 vs 64 frames before the fixed-point add
                        fa, fb = ops.se_pool_to_float(a), ops.se_pool_to_float(b)
                        assert (fa - fb).abs().max() <= 1e-5 * fa.abs().max(), (dt, Cout, c0, c1)
                    else:
                        assert torch.equal(a, b), (dt, Cout, c0, c1, i)
    finally:
        _lib.set_option("pw_ws", 0)


def test_se_fc_matches_oracle():
    rng = np.random.default_rng(7)
    B, C, H, T = 3, 64, 8, 41
    x = rng.standard_normal((B, C, T)).astype(np.float32)
    w1 = rng.standard_normal((H, C)).astype(np.float32) / 8
    w2 = rng.standard_normal((C, H)).astype(np.float32) / 3
    gate = ops.se_fc(dev(x.sum(-1)), T, dev(w1), dev(w2)).cpu().numpy()
    ref = R.squeeze_excite(x, w1, w2) / x     # = the gate, broadcast over T
    emax, _ = rel_err(gate, ref[:, :, 0])
    assert emax < 1e-5


def test_ctc_greedy_known_answers(golden_decode):
    g = golden_decode
    lg = g["argmax_logits"]
    ids, col, cnt = ops.ctc_greedy(dev(lg), lg.shape[2], -1)
    assert np.array_equal(ids.cpu().numpy(), g["argmax_ids"])          # ties -> first index, NaN maximal
    # collapse == torch.unique_consecutive (via the oracle), blanks kept
    rng = np.random.default_rng(8)
    B, V, T = 5, 29, 203
    idx = rng.integers(0, V, (B, T))
    idx[:, 1::2] = idx[:, ::2][:, : idx[:, 1::2].shape[1]]
    idx[3] = 28
    logits = np.full((B, V, T), -1.0, np.float32)
    np.put_along_axis(logits, idx[:, None, :], 1.0, axis=1)
    ids, col, cnt = [t.cpu().numpy() for t in ops.ctc_greedy(dev(logits), T, -1)]
    assert np.array_equal(ids, idx)
    for b, row in enumerate(R.ctc_collapse(idx)):
        assert cnt[b] == len(row) and np.array_equal(col[b, : cnt[b]], row) and (col[b, cnt[b]:] == -1).all()
    # device-side blank drop (optional mode)
    _, col2, cnt2 = [t.cpu().numpy() for t in ops.ctc_greedy(dev(logits), T, 28)]
    for b, row in enumerate(R.ctc_collapse(idx)):
        keep = row[row != 28]
        assert cnt2[b] == len(keep) and np.array_equal(col2[b, : cnt2[b]], keep)
    # bf16 padded-row logits
    rows = to_rows(logits)
    ids3, _, _ = ops.ctc_greedy(rows, T, -1)
    assert np.array_equal(ids3.cpu().numpy(), idx)
    # word-piece vocabularies (V >= 128: rows split over the warps of a CTA, partial results merged): torch.argmax's rules
    # survive the merge -- the FIRST maximal index wins across slices, the first NaN beats everything, -inf columns give 0
    for V2, T2 in ((1025, 251), (128, 33), (300, 97)):
        lw = rng.standard_normal((3, V2, T2)).astype(np.float32)
        lw[0, :, 5] = -np.inf                                   # all -inf -> index 0
        lw[0, [7, 8, 600 % V2, V2 - 1], 6] = 9.0                # four-way tie across different warps' slices -> 7
        lw[1, [V2 - 1, 3], 0] = [np.nan, np.nan]                # two NaNs -> the first one (3)
        lw[1, 100, 1] = np.nan; lw[1, 17, 1] = 50.0             # NaN beats a larger finite value
        lw[2, :, :] = np.round(lw[2] * 2) / 2                   # many exact ties
        want = np.empty((3, T2), np.int64)
        for b_ in range(3):
            for t_ in range(T2):
                col_ = lw[b_, :, t_]
                nn = np.nonzero(np.isnan(col_))[0]
                want[b_, t_] = nn[0] if nn.size else int(np.argmax(col_))
        got = ops.ctc_greedy(dev(lw), T2, -1)[0].cpu().numpy()
        assert np.array_equal(got, want), (V2, T2)
        assert np.array_equal(got, torch.from_numpy(lw).argmax(1).numpy())       # torch agrees (NaN maximal, first index)
        lb = torch.from_numpy(lw).cuda().bfloat16()                                # bf16 rows, pitch > T
        rows_w = torch.zeros((3, V2, ops.row_pitch(T2)), device="cuda", dtype=torch.bfloat16)
        rows_w[:, :, :T2] = lb
        gb = ops.ctc_greedy(rows_w, T2, -1)[0].cpu()
        assert torch.equal(gb, lb.float().cpu().argmax(1))
