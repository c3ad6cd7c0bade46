"""GPU parity of the training step (train()-mode forward, backward, parameter gradients, BatchNorm running statistics)
against goldens produced by the reference's own modules with torch autograd (tests/golden/train.npz) and against the
autograd-capable torch port.  bf16 activations / gradients: tolerances are relative L2 errors."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import ref_numpy as R
from oracle.make_golden_train import TRAIN_BLOCK_CASES, block_case, tiny_model_case
from thunder_speech_b200 import ops, synth
from thunder_speech_b200.blocks import conv1d_decoder
from thunder_speech_b200.module import CTCModule
from thunder_speech_b200.quartznet.blocks import QuartznetBlock, QuartznetEncoder
from thunder_speech_b200.quartznet.transform import FilterbankFeatures
from thunder_speech_b200.text_processing import BatchTextTransformer
from thunder_speech_b200.train import BlockTrainer, CTCTrainStep

pytestmark = pytest.mark.gpu

# Gradients are compared by relative L2 error.  With bf16 activations the ReLU masks of ~0.3 % of the elements (those with
# |y| below one bf16 ulp of the pre-activation) flip relative to the fp32 reference; each flip changes dy*mask by a whole
# element, so the gradient error is ~sqrt(fraction flipped) ~ 5 % on these 100-150 sample batches (the single
# sub-block cases agree to 0.3-2 %).
GRAD_TOL = 8e-2


@pytest.fixture(scope="module")
def gtrain():
    return np.load("tests/golden/train.npz", allow_pickle=False)


def l2(a, b):
    return rel_err(a, b)[1]


@pytest.mark.parametrize("ci", range(len(TRAIN_BLOCK_CASES)))
def test_block_training_step_vs_reference(gtrain, ci):
    name, cfg, st, x, lens, Rm = block_case(ci)
    blk = QuartznetBlock(cfg["in_channels"], cfg["out_channels"], repeat=cfg["repeat"], kernel_size=(cfg["kernel_size"],),
                         stride=(cfg["stride"],), dilation=(cfg["dilation"],), residual=cfg["residual"],
                         separable=cfg["separable"])
    blk.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}, strict=True)
    blk = blk.cuda().train()
    bt = BlockTrainer(blk)
    l32 = torch.from_numpy(lens.astype(np.int32)).cuda()
    rows = ops.pack_rows(torch.from_numpy(x).cuda(), l32)
    y, T_out, lo, tape = bt.forward(rows, x.shape[-1], l32, zero_tail=False)
    out = ops.unpack_rows(y, T_out).cpu().numpy()
    assert out.shape == gtrain[f"{name}.out"].shape
    assert np.array_equal(lo.cpu().numpy(), gtrain[f"{name}.out_lengths"])
    assert l2(out, gtrain[f"{name}.out"]) < 2e-2
    need_dx = cfg["stride"] == 1
    dx = bt.backward(tape, ops.pack_rows(torch.from_numpy(Rm).cuda()), need_dx=need_dx)
    torch.cuda.synchronize()
    if need_dx:
        e = l2(ops.unpack_rows(dx, x.shape[-1]).cpu().numpy(), gtrain[f"{name}.dx"])
        assert e < GRAD_TOL, (name, "dx", e)
    for k, p in blk.named_parameters():
        ref = gtrain[f"{name}.grad.{k}"]
        e = l2(p.grad.cpu().numpy(), ref)
        assert e < GRAD_TOL, (name, k, e)
    for k, b in blk.named_buffers():
        if "running" in k:
            assert rel_err(b.cpu().numpy(), gtrain[f"{name}.buf.{k}"])[0] < 1e-2, (name, k)


def test_pw_wgrad_and_dw_wgrad_ops():
    from thunder_speech_b200.train import dw_wgrad, pw_wgrad

    rng = np.random.default_rng(3)
    B, Cout, Cin, T = 5, 300, 140, 333
    dz = rng.standard_normal((B, Cout, T)).astype(np.float32)
    a = rng.standard_normal((B, Cin, T)).astype(np.float32)
    dzr, ar = ops.pack_rows(torch.from_numpy(dz).cuda()), ops.pack_rows(torch.from_numpy(a).cuda())
    got = pw_wgrad(dzr, ar, T).cpu().numpy()
    ref = np.einsum("bot,bit->oi", dzr.float().cpu().numpy()[:, :, :T].astype(np.float64),
                    ar.float().cpu().numpy()[:, :, :T].astype(np.float64))
    assert rel_err(got, ref)[0] < 1e-4
    for (K, S, D) in ((5, 1, 1), (33, 2, 1), (9, 1, 2)):
        P = R.get_same_padding(K, S, D)
        Tin = 201
        Tout = (Tin + 2 * P - D * (K - 1) - 1) // S + 1
        C = 7
        x = rng.standard_normal((B, C, Tin)).astype(np.float32)
        da = rng.standard_normal((B, C, Tout)).astype(np.float32)
        lens = np.array([201, 150, 100, 201, 7], np.int32)
        xr, dar = ops.pack_rows(torch.from_numpy(x).cuda()), ops.pack_rows(torch.from_numpy(da).cuda())
        got = dw_wgrad(dar, Tout, xr, Tin, torch.from_numpy(lens).cuda(), K, S, D, P).cpu().numpy()
        xm = np.where(np.arange(Tin)[None, None, :] < lens[:, None, None], xr.float().cpu().numpy()[:, :, :Tin], 0).astype(np.float64)
        xp = np.pad(xm, ((0, 0), (0, 0), (P, P + S)))
        dd = dar.float().cpu().numpy()[:, :, :Tout].astype(np.float64)
        ref = np.stack([(dd * xp[:, :, k * D: k * D + (Tout - 1) * S + 1: S]).sum((0, 2)) for k in range(K)], axis=1)
        assert rel_err(got, ref)[0] < 1e-4, (K, S, D)


def test_model_training_step_vs_reference(gtrain):
    filters, kernels, st, dec, x, lens, y, y_len = tiny_model_case()
    enc = QuartznetEncoder(filters=filters, kernel_sizes=kernels, repeat_blocks=1)
    enc.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}, strict=True)
    d = conv1d_decoder(1024, 29)
    d.load_state_dict({k: torch.from_numpy(v) for k, v in dec.items()})
    m = CTCModule(enc, d, FilterbankFeatures(nfilt=64, dither=0.0), BatchTextTransformer(synth.quartznet_vocab())).cuda()
    m.encoder.train(); m.decoder.train()
    step = CTCTrainStep(m, lr=1e-3)
    loss = step.loss_and_grads(torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda(), torch.from_numpy(y).cuda(),
                               torch.from_numpy(y_len).cuda())
    ref_loss = float(gtrain["model.loss"])
    assert abs(loss.item() - ref_loss) < 3e-2 * abs(ref_loss), (loss.item(), ref_loss)
    # the tiny golden case (62 BatchNorm samples per channel) is too ill-conditioned to compare deep gradients in bf16;
    # gradient direction is checked below against the oracle on a larger batch
    # an optimiser step runs and changes the parameters
    w0 = m.decoder.weight.detach().clone()
    step.opt.step()
    assert not torch.equal(w0, m.decoder.weight)


def _model_case():
    filters, kernels = [32, 32, 32, 32, 32], [5, 7, 9, 11, 13]
    st = synth.encoder_state(synth.quartznet_block_list(filters=filters, kernel_sizes=kernels, repeat_blocks=1), seed=31)
    dec = synth.decoder_state(1024, 29, seed=32)
    x = synth.audio(8, 32000, 33, "tones")
    lens = np.array([32000, 32000, 30000, 28000, 25000, 22222, 20000, 16000], np.int64)
    rng = np.random.default_rng(34)
    y = rng.integers(0, 28, (8, 12)).astype(np.int64)
    y_len = rng.integers(4, 13, 8).astype(np.int64)
    return filters, kernels, st, dec, x, lens, y, y_len


def _oracle_model(case, store=None, steps=0, lr=1e-3):
    """Autograd torch port of the training step on CPU (fp32, or fp32 with bf16 STORAGE simulated at the points where
    the device path writes bf16 rows).  Returns (losses, grads of the first step)."""
    from oracle import ref_torch as RT

    filters, kernels, st, dec, x, lens, y, y_len = case
    cfgs = R.quartznet_cfgs(filters=filters, kernel_sizes=kernels, repeat_blocks=1)
    stt = {k: torch.from_numpy(np.asarray(v)).clone() for k, v in st.items()}
    for k, v in stt.items():
        if v.dtype.is_floating_point and "running" not in k:
            v.requires_grad_(True)
    dw_, db_ = torch.from_numpy(dec["weight"]).clone().requires_grad_(True), torch.from_numpy(dec["bias"]).clone().requires_grad_(True)
    params = [v for v in stt.values() if v.requires_grad] + [dw_, db_]
    opt = torch.optim.AdamW(params, lr=lr)
    with torch.no_grad():
        f, fl = RT.features(torch.from_numpy(x), torch.from_numpy(lens))
    losses, grads = [], None
    for it in range(max(steps, 1)):
        opt.zero_grad()
        e, el = RT.encoder(f, fl, cfgs, stt, train=True, store=store)
        loss = RT.ctc_loss(torch.nn.functional.conv1d(e, dw_, db_), torch.from_numpy(y), el, torch.from_numpy(y_len), 28)
        loss.backward()
        losses.append(loss.item())
        if grads is None:
            grads = {k: v.grad.numpy().copy() for k, v in stt.items() if v.requires_grad}
            grads["decoder.weight"], grads["decoder.bias"] = dw_.grad.numpy().copy(), db_.grad.numpy().copy()
        if steps:
            opt.step()
    return losses, grads


def _device_model(case, lr=1e-3):
    filters, kernels, st, dec, x, lens, y, y_len = case
    enc = QuartznetEncoder(filters=filters, kernel_sizes=kernels, repeat_blocks=1)
    enc.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}, strict=True)
    d = conv1d_decoder(1024, 29)
    d.load_state_dict({k: torch.from_numpy(v) for k, v in dec.items()})
    m = CTCModule(enc, d, FilterbankFeatures(nfilt=64, dither=0.0), BatchTextTransformer(synth.quartznet_vocab())).cuda()
    m.encoder.train(); m.decoder.train()
    batch = (torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(y_len).cuda())
    return m, CTCTrainStep(m, lr=lr), batch


def _cos(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-30))


def test_model_gradients_vs_oracle_larger_batch():
    """Per-parameter gradients of the full training step (8 utterances x 2 s, 28 conv+BN+ReLU layers) against the autograd
    torch port.  BatchNorm + ReLU at random init is in the chaotic regime (a perturbation grows ~1.15x per layer: the fp32
    port and the SAME port with bf16 storage already differ by 17 % at the encoder output), so the criterion for the deep
    layers is relative: the device gradients must be as close to the fp32 oracle as the bf16-storage oracle is.  Kernel
    correctness proper is pinned by the block-level tests above (<= 2 % against the bf16-storage oracle)."""
    from oracle import ref_torch as RT

    case = _model_case()
    (ref_loss,), ref = _oracle_model(case)
    (sim_loss,), sim = _oracle_model(case, store=RT.bf16_store)
    m, step, batch = _device_model(case)
    loss = step.loss_and_grads(*batch)
    assert abs(loss.item() - ref_loss) < 2e-2 * abs(ref_loss), (loss.item(), ref_loss)
    got = {k: p.grad.float().cpu().numpy() for k, p in m.encoder.named_parameters()}
    got["decoder.weight"], got["decoder.bias"] = m.decoder.weight.grad.cpu().numpy(), m.decoder.bias.grad.cpu().numpy()
    for k in ref:
        na, nb = np.linalg.norm(got[k]), np.linalg.norm(ref[k])
        assert abs(na - nb) < 0.4 * nb + 1e-8, (k, na, nb)
    # per encoder block (all its parameters concatenated -- single 32-element BN vectors are too noisy to compare)
    groups = sorted({k.split(".")[0] for k in ref})
    cat = lambda g, grp: np.concatenate([g[k].ravel() for k in sorted(ref) if k.split(".")[0] == grp])
    ours = {grp: _cos(cat(got, grp), cat(ref, grp)) for grp in groups}
    base = {grp: _cos(cat(sim, grp), cat(ref, grp)) for grp in groups}
    print("cosine vs fp32 oracle per block: device", {k: round(v, 3) for k, v in ours.items()})
    print("                   bf16-storage oracle", {k: round(v, 3) for k, v in base.items()})
    for grp in groups:
        assert ours[grp] > base[grp] - 0.1, (grp, ours[grp], base[grp])
    assert np.mean(list(ours.values())) > np.mean(list(base.values())) - 0.03
    # shallow end of the backward pass: few layers of amplification, absolute bounds hold
    assert ours["decoder"] > 0.995 and ours["7"] > 0.97 and ours["6"] > 0.92


def test_model_loss_trajectory_vs_oracle():
    """Six AdamW steps on one batch: the loss curve of the device training step follows the fp32 oracle's (within 10 %
    during the steep descent: 24.8 -> 13.3 -> 6.4 -> 4.8 -> 5.0 measured vs 24.8 -> 13.9 -> 6.3 -> 5.1 -> 5.0)."""
    case = _model_case()
    ref_losses, _ = _oracle_model(case, steps=6, lr=2e-3)
    m, step, batch = _device_model(case, lr=2e-3)
    losses = [step.step(*batch).item() for _ in range(6)]
    print("losses device", losses, "oracle", ref_losses)
    assert ref_losses[-1] < 0.9 * ref_losses[0]
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) < 1e-1 * abs(b), (losses, ref_losses)


SIM_BLOCK_CASES = [  # cin, cout, K, repeat, B, T, residual
    (32, 32, 5, 1, 8, 101, False),
    (32, 32, 13, 5, 8, 101, True),
    (256, 32, 5, 5, 8, 101, True),
    (64, 128, 11, 3, 4, 300, True),
]


@pytest.mark.parametrize("cin,cout,K,rep,B,T,res", SIM_BLOCK_CASES)
def test_block_gradients_vs_bf16_storage_oracle(cin, cout, K, rep, B, T, res):
    """Forward and every gradient of one block against the autograd port with bf16 storage simulated (same ReLU masks and
    batch statistics as the device): what is left is backward-pass bf16 rounding."""
    from oracle import ref_torch as RT

    rng = np.random.Generator(np.random.PCG64(cin + K))
    st = synth.block_state(rng, "", cin, cout, rep, K, res, True)
    x = np.maximum(rng.standard_normal((B, cin, T)), 0).astype(np.float32)
    lens = np.sort(rng.integers(T // 2, T + 1, B))[::-1].astype(np.int64).copy()
    lens[0] = T
    m = (np.arange(T)[None, :] < lens[:, None])[:, None, :]
    Rm = np.where(m, rng.standard_normal((B, cout, T)), 0).astype(np.float32)
    cfg = R.BlockCfg(cin, cout, repeat=rep, kernel_size=K, residual=res, separable=True)
    stt = {k: torch.from_numpy(np.asarray(v)).clone() for k, v in st.items()}
    for k, v in stt.items():
        if v.dtype.is_floating_point and "running" not in k:
            v.requires_grad_(True)
    xt = torch.from_numpy(np.where(m, x, 0).astype(np.float32)).requires_grad_(True)
    y, _ = RT.block(xt, torch.from_numpy(lens), cfg, stt, "", train=True, store=RT.bf16_store)
    (y * torch.from_numpy(Rm)).sum().backward()
    blk = QuartznetBlock(cin, cout, repeat=rep, kernel_size=(K,), residual=res, separable=True)
    blk.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}, strict=True)
    blk = blk.cuda().train()
    bt = BlockTrainer(blk)
    l32 = torch.from_numpy(lens.astype(np.int32)).cuda()
    yy, T_out, lo, tape = bt.forward(ops.pack_rows(torch.from_numpy(x).cuda(), l32), T, l32, zero_tail=True)
    assert l2(ops.unpack_rows(yy, T_out).cpu().numpy(), np.where(m, y.detach().numpy(), 0)) < 5e-3
    dx = bt.backward(tape, ops.pack_rows(torch.from_numpy(Rm).cuda()), need_dx=True)
    # single-ulp bf16 differences in the forward cascade through the five sub-blocks (a handful of ReLU masks end up
    # differing by the last one): 0.3-0.5 % for 1-3 sub-blocks, 0.6-3 % for 5
    tol = 2e-2 if rep <= 3 else 5e-2
    assert l2(ops.unpack_rows(dx, T).cpu().numpy(), xt.grad.numpy()) < tol
    for k, p in blk.named_parameters():
        e = l2(p.grad.cpu().numpy(), stt[k].grad.numpy())
        assert e < tol, (k, e)


@pytest.mark.parametrize("B,V,T,Lmax,seed", [(6, 29, 101, 12, 0), (3, 29, 376, 120, 1), (4, 5, 40, 30, 2), (2, 200, 33, 7, 3),
                                              (3, 1025, 70, 20, 4)])
def test_ctc_loss_kernel_vs_torch(B, V, T, Lmax, seed):
    """ts_ctc_loss against log_softmax + F.ctc_loss(reduction="mean", zero_infinity=True) (src/thunder/ctc_loss.py:15-47)
    in float64 on the CPU: per-utterance loss and the gradient w.r.t. the logits.  Covers ragged input / target lengths,
    repeated labels (forced blanks), an empty target and an infeasible utterance (target longer than the input: infinite
    loss -> zeroed)."""
    from thunder_speech_b200.train import ctc_loss

    rng = np.random.default_rng(seed)
    blank = V - 1
    logits = (3.0 * rng.standard_normal((B, V, T))).astype(np.float32)
    y = rng.integers(0, V - 1, (B, Lmax)).astype(np.int64)
    y[0, : Lmax // 2] = y[0, 0]                        # long run of a repeated label
    tl = rng.integers(1, Lmax + 1, B).astype(np.int64)
    il = rng.integers(T // 2, T + 1, B).astype(np.int32)
    il[0] = T
    tl[0] = min(Lmax, T // 3)
    if B > 2:
        tl[1] = 0                                      # empty transcript
        il[2], tl[2] = 4, min(Lmax, 6)                 # infeasible -> inf -> zero_infinity
    for b in range(B):                                 # keep the others feasible
        if not (B > 2 and b == 2):
            tl[b] = min(tl[b], il[b] // 2)
    # reference (float64)
    lt = torch.from_numpy(logits).double().requires_grad_(True)
    lp = torch.nn.functional.log_softmax(lt.permute(2, 0, 1), dim=2)
    per = torch.nn.functional.ctc_loss(lp, torch.from_numpy(y), torch.from_numpy(il).long(), torch.from_numpy(tl), blank=blank,
                                       reduction="none", zero_infinity=True)
    ref_loss = per / torch.from_numpy(tl).clamp_min(1)
    ref_loss.mean().backward()
    # kernel
    pitch = ops.row_pitch(T)
    lg = torch.zeros((B, V, pitch), dtype=torch.float32, device="cuda")
    lg[:, :, :T] = torch.from_numpy(logits).cuda()
    lg[:, :, T:] = float("nan")                        # the pad must never be read
    loss, grad = ctc_loss(lg, T, torch.from_numpy(il).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(tl).cuda(), blank,
                          Vp=(V + 63) // 64 * 64)
    torch.cuda.synchronize()
    assert np.allclose(loss.cpu().numpy(), ref_loss.detach().numpy(), rtol=2e-5, atol=1e-5), (loss.cpu().numpy(), ref_loss)
    g = grad.float().cpu().numpy()
    assert np.all(g[:, V:, :] == 0) and np.all(g[:, :, T:] == 0)
    gr = lt.grad.numpy()
    # bf16 output: 2^-9 relative per element
    assert np.abs(g[:, :V, :T] - gr).max() <= 2 ** -8 * np.abs(gr).max() + 1e-9
    assert l2(g[:, :V, :T], gr) < 4e-3
    if B > 2:
        assert loss[2].item() == 0.0 and np.all(g[2] == 0)


def test_pw_gemm_stats_epilogue():
    """Statistics from the GEMM epilogue == statistics of the stored bf16 output, for a time length with a partial tile."""
    from thunder_speech_b200.train import pw_gemm_stats

    rng = np.random.default_rng(11)
    B, Cin, Cout, T = 5, 256, 384, 751
    w = torch.from_numpy((rng.standard_normal((Cout, Cin)) / 16).astype(np.float32)).cuda().to(torch.bfloat16)
    lens = torch.from_numpy(np.array([751, 700, 512, 300, 64], np.int32)).cuda()
    x = ops.pack_rows(torch.from_numpy(rng.standard_normal((B, Cin, T)).astype(np.float32)).cuda(), lens)
    z, st = pw_gemm_stats(w, x, T)
    assert st.shape == (B, Cout, 6, 2)
    zref = ops.pw_gemm(w, x, None, None, T, None, None, False, False, None, None, None)
    assert torch.equal(z, zref)
    zf = z.float()[:, :, :T].double()
    got = st.double().sum(2).cpu().numpy()
    assert rel_err(got[..., 0], zf.sum(-1).cpu().numpy())[0] < 1e-5
    assert rel_err(got[..., 1], (zf * zf).sum(-1).cpu().numpy())[0] < 1e-5


def test_bn_finalize_and_bwd_coef_kernels():
    from thunder_speech_b200.train import bn_bwd_coef, bn_finalize

    rng = np.random.default_rng(5)
    NB, C, n = 7, 300, 7 * 123
    part = rng.standard_normal((NB, C, 2)).astype(np.float32)
    part[:, :, 1] = np.abs(part[:, :, 1]) * 50 + 30
    bn = torch.nn.BatchNorm1d(C, eps=1e-3, momentum=0.1).cuda()
    with torch.no_grad():
        bn.weight.copy_(torch.from_numpy(rng.uniform(0.5, 1.5, C).astype(np.float32)))
        bn.bias.copy_(torch.from_numpy(rng.standard_normal(C).astype(np.float32)))
        bn.running_mean.copy_(torch.from_numpy(rng.standard_normal(C).astype(np.float32)))
    rm0, rv0 = bn.running_mean.cpu().numpy().astype(np.float64), bn.running_var.cpu().numpy().astype(np.float64)
    scale, shift, mean, inv = bn_finalize(torch.from_numpy(part).cuda(), n, bn, True)
    s = part.astype(np.float64).sum(0)
    m = s[:, 0] / n
    var = np.maximum(s[:, 1] / n - m * m, 0)
    iv = 1 / np.sqrt(var + 1e-3)
    g, b = bn.weight.detach().cpu().numpy().astype(np.float64), bn.bias.detach().cpu().numpy().astype(np.float64)
    assert rel_err(mean.cpu().numpy(), m)[0] < 1e-6 and rel_err(inv.cpu().numpy(), iv)[0] < 1e-6
    assert rel_err(scale.cpu().numpy(), g * iv)[0] < 1e-6 and rel_err(shift.cpu().numpy(), b - m * g * iv)[0] < 1e-6
    assert rel_err(bn.running_mean.cpu().numpy(), 0.9 * rm0 + 0.1 * m)[0] < 1e-6
    assert rel_err(bn.running_var.cpu().numpy(), 0.9 * rv0 + 0.1 * var * n / (n - 1))[0] < 1e-6
    part3 = rng.standard_normal((NB, C, 3)).astype(np.float32)
    for which in (1, 2):
        bn.weight.grad = None
        bn.bias.grad = None
        coef = bn_bwd_coef(torch.from_numpy(part3).cuda(), which, n, bn, mean, inv).cpu().numpy()
        s3 = part3.astype(np.float64).sum(0)
        dg = iv * (s3[:, which] - m * s3[:, 0])
        a = g * iv
        bb = -g * iv * iv * dg / n
        cc = -a * s3[:, 0] / n - bb * m
        assert rel_err(bn.weight.grad.cpu().numpy(), dg)[0] < 1e-5 and rel_err(bn.bias.grad.cpu().numpy(), s3[:, 0])[0] < 1e-6
        assert rel_err(coef, np.stack([a, bb, cc], 1))[0] < 1e-5


@pytest.mark.parametrize("K,D,C,B,T", [(33, 1, 7, 5, 201), (39, 1, 3, 9, 751), (75, 1, 150, 3, 100), (87, 2, 5, 4, 333),
                                        (5, 1, 4, 17, 64), (11, 1, 2, 2, 129), (127, 1, 3, 2, 300)])
def test_dw_wgrad_tensor_core_kernel(K, D, C, B, T):
    """dwwgrad_mma.cu (windows-as-K MMA + diagonal sums) against a float64 correlation of the same bf16 rows; ragged
    lengths, batch not a multiple of the utterance group, halos of 1-2 windows, dilation 2."""
    from thunder_speech_b200 import _lib
    from thunder_speech_b200.train import dw_wgrad

    rng = np.random.default_rng(K + C)
    P = D * (K - 1) // 2
    x = rng.standard_normal((B, C, T)).astype(np.float32)
    da = rng.standard_normal((B, C, T)).astype(np.float32)
    lens = rng.integers(T // 2, T + 1, B).astype(np.int32)
    lens[0] = T
    l32 = torch.from_numpy(lens).cuda()
    xr, dar = ops.pack_rows(torch.from_numpy(x).cuda(), l32), ops.pack_rows(torch.from_numpy(da).cuda(), l32)
    n0 = _lib.launch_count()
    got = dw_wgrad(dar, T, xr, T, l32, K, 1, D, P, premasked=True).cpu().numpy()
    assert _lib.launch_count() == n0 + 1
    xm = xr.float().cpu().numpy()[:, :, :T].astype(np.float64)
    dd = dar.float().cpu().numpy()[:, :, :T].astype(np.float64)
    xp = np.pad(xm, ((0, 0), (0, 0), (P, P)))
    ref = np.stack([(dd * xp[:, :, k * D: k * D + T]).sum((0, 2)) for k in range(K)], axis=1)
    assert rel_err(got, ref)[0] < 1e-4, (K, D, rel_err(got, ref))
    # and the SIMT kernel on the same data agrees
    simt = dw_wgrad(dar, T, xr, T, l32, K, 1, D, P, premasked=False).cpu().numpy()
    assert rel_err(simt, ref)[0] < 1e-4


def test_parameters_update_template_block_and_model():
    """The reference's `_test_parameters_update` template (tests/utils.py:38-50): after backward of `outputs.mean()` every
    trainable parameter has a non-zero gradient, and an optimiser step moves every parameter -- for one block through
    BlockTrainer and for the whole model through CTCTrainStep."""
    rng = np.random.Generator(np.random.PCG64(77))
    cin, cout, K, rep, B, T = 64, 128, 11, 3, 4, 200
    st = synth.block_state(rng, "", cin, cout, rep, K, True, True)
    blk = QuartznetBlock(cin, cout, repeat=rep, kernel_size=(K,), residual=True, separable=True)
    blk.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}, strict=True)
    blk = blk.cuda().train()
    bt = BlockTrainer(blk)
    x = torch.from_numpy(np.maximum(rng.standard_normal((B, cin, T)), 0).astype(np.float32)).cuda()
    y, T_out, _, tape = bt.forward(ops.pack_rows(x), T, None, zero_tail=False)
    dy = torch.full((B, cout, T_out), 1.0 / (B * cout * T_out), device="cuda")       # d mean(y) / dy
    # a constant dy is annihilated by the last BatchNorm's mean subtraction (dgamma ~ 0 by construction: the template's
    # reference model has the same property only up to ReLU masking) -> modulate it like a masked mean
    dy = dy * torch.from_numpy((rng.random((B, cout, T_out)) > 0.3).astype(np.float32)).cuda()
    bt.backward(tape, ops.pack_rows(dy), need_dx=True)
    torch.cuda.synchronize()
    opt = torch.optim.SGD(blk.parameters(), lr=0.1)
    before = {k: p.detach().clone() for k, p in blk.named_parameters()}
    for k, p in blk.named_parameters():
        assert p.grad is not None and float((p.grad ** 2).sum()) != 0.0, k
    opt.step()
    for k, p in blk.named_parameters():
        assert not torch.equal(before[k], p.detach()), k
    # whole model
    m, step, batch = _device_model(_model_case(), lr=1e-3)
    before = {k: p.detach().clone() for k, p in list(m.encoder.named_parameters()) + list(m.decoder.named_parameters())}
    step.step(*batch)
    torch.cuda.synchronize()
    for k, p in list(m.encoder.named_parameters()) + list(m.decoder.named_parameters()):
        assert p.grad is not None and float((p.grad ** 2).sum()) != 0.0, k
        assert not torch.equal(before[k], p.detach()), k
    # BatchNorm bookkeeping like nn.BatchNorm1d in train(): one step -> num_batches_tracked == 1
    for k, b in m.encoder.named_buffers():
        if k.endswith("num_batches_tracked"):
            assert int(b) == 1, k


@pytest.mark.parametrize("cin,cout,K,rep,B,T,res", [(64, 128, 11, 3, 4, 200, True), (32, 64, 5, 1, 6, 101, False),
                                                     (256, 256, 13, 5, 8, 251, True)])
def test_citrinet_block_with_squeeze_excite_training(cin, cout, K, rep, B, T, res):
    """train()-mode forward and every gradient (incl. the SqueezeExcite FCs) of a stride-1 CitrinetBlock against the autograd
    torch port, bf16 storage simulated (citrinet/blocks.py:48-83,177-197)."""
    from oracle import ref_torch as RT
    from thunder_speech_b200.citrinet.blocks import CitrinetBlock

    rng = np.random.Generator(np.random.PCG64(cin + K + rep))
    st = synth.block_state(rng, "", cin, cout, rep, K, res, True, se=True)
    x = np.maximum(rng.standard_normal((B, cin, T)), 0).astype(np.float32)
    lens = np.sort(rng.integers(T // 2, T + 1, B))[::-1].astype(np.int64).copy()
    lens[0] = T
    m = (np.arange(T)[None, :] < lens[:, None])[:, None, :]
    Rm = np.where(m, rng.standard_normal((B, cout, T)), 0).astype(np.float32)
    cfg = R.BlockCfg(cin, cout, repeat=rep, kernel_size=K, residual=res, separable=True, kind="citrinet")
    def oracle(store):
        sd = {k: torch.from_numpy(np.asarray(v)).clone() for k, v in st.items()}
        for k, v in sd.items():
            if v.dtype.is_floating_point and "running" not in k:
                v.requires_grad_(True)
        xi = torch.from_numpy(np.where(m, x, 0).astype(np.float32)).requires_grad_(True)
        yo, _ = RT.block(xi, torch.from_numpy(lens), cfg, sd, "", train=True, store=store)
        (yo * torch.from_numpy(Rm)).sum().backward()
        return sd, xi.grad.numpy(), yo.detach().numpy()

    stt, dx_sim, y_sim = oracle(RT.bf16_store)
    blk = CitrinetBlock(cin, cout, repeat=rep, kernel_size=(K,), residual=res, separable=True)
    blk.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}, strict=True)
    blk = blk.cuda().train()
    bt = BlockTrainer(blk)
    l32 = torch.from_numpy(lens.astype(np.int32)).cuda()
    yy, T_out, lo, tape = bt.forward(ops.pack_rows(torch.from_numpy(x).cuda(), l32), T, l32, zero_tail=True)
    assert l2(ops.unpack_rows(yy, T_out).cpu().numpy(), np.where(m, y_sim, 0)) < 6e-3
    dx = bt.backward(tape, ops.pack_rows(torch.from_numpy(Rm).cuda()), need_dx=True)
    dxn = ops.unpack_rows(dx, T).cpu().numpy()
    if rep <= 3:
        # shallow: what is left against the bf16-storage oracle is backward rounding (measured 0.2-0.5 %)
        assert l2(dxn, dx_sim) < 2e-2
        for k, p in blk.named_parameters():
            e = l2(p.grad.cpu().numpy(), stt[k].grad.numpy())
            assert e < 2e-2, (k, e)
    else:
        # 5 sub-blocks: the single-ulp forward cascade (see the QuartzNet test above) also passes through the SE gate, and
        # for the parameters fed by d loss / d mean_t(u) (last BN bias, fc.0) the bf16-storage oracle itself is 16-20 % away
        # from the fp32 oracle while the device is within 7 % of fp32: accept closeness to EITHER reference
        ref32, dx32, _ = oracle(None)
        assert min(l2(dxn, dx_sim), l2(dxn, dx32)) < GRAD_TOL
        for k, p in blk.named_parameters():
            g = p.grad.cpu().numpy()
            e = min(l2(g, stt[k].grad.numpy()), l2(g, ref32[k].grad.numpy()))
            assert e < GRAD_TOL, (k, e)
    for k, b in blk.named_buffers():
        if "running" in k:
            assert rel_err(b.cpu().numpy(), stt[k].detach().numpy())[0] < 1e-2, k


@pytest.mark.parametrize("cin,cout,K,rep,B,T,res", [(64, 128, 11, 3, 4, 201, True), (32, 64, 5, 1, 6, 100, True),
                                                     (256, 256, 13, 2, 4, 333, False)])
def test_strided_citrinet_block_training(cin, cout, K, rep, B, T, res):
    """Stride-2 CitrinetBlock (last sub-block strided, strided 1x1 residual, SqueezeExcite): forward, dx and every gradient
    against the autograd torch port with bf16 storage simulated.  Covers the transposed depthwise conv (zero-upsample +
    flipped taps), the strided weight gradient and ts_scatter_rows."""
    from oracle import ref_torch as RT
    from thunder_speech_b200.citrinet.blocks import CitrinetBlock

    rng = np.random.Generator(np.random.PCG64(7 * cin + K + rep))
    st = synth.block_state(rng, "", cin, cout, rep, K, res, True, se=True)
    x = np.maximum(rng.standard_normal((B, cin, T)), 0).astype(np.float32)
    lens = np.sort(rng.integers(T // 2, T + 1, B))[::-1].astype(np.int64).copy()
    lens[0] = T
    To = (T - 1) // 2 + 1
    lo_ref = (lens - 1) // 2 + 1
    m = (np.arange(T)[None, :] < lens[:, None])[:, None, :]
    mo = (np.arange(To)[None, :] < lo_ref[:, None])[:, None, :]
    Rm = np.where(mo, rng.standard_normal((B, cout, To)), 0).astype(np.float32)
    cfg = R.BlockCfg(cin, cout, repeat=rep, kernel_size=K, stride=2, residual=res, separable=True, kind="citrinet")
    stt = {k: torch.from_numpy(np.asarray(v)).clone() for k, v in st.items()}
    for k, v in stt.items():
        if v.dtype.is_floating_point and "running" not in k:
            v.requires_grad_(True)
    xt = torch.from_numpy(np.where(m, x, 0).astype(np.float32)).requires_grad_(True)
    y, yl = RT.block(xt, torch.from_numpy(lens), cfg, stt, "", train=True, store=RT.bf16_store)
    assert y.shape[-1] == To and np.array_equal(yl.numpy(), lo_ref)
    (y * torch.from_numpy(Rm)).sum().backward()
    blk = CitrinetBlock(cin, cout, repeat=rep, kernel_size=(K,), stride=(2,), residual=res, separable=True)
    blk.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}, strict=True)
    blk = blk.cuda().train()
    bt = BlockTrainer(blk)
    l32 = torch.from_numpy(lens.astype(np.int32)).cuda()
    yy, T_out, lo, tape = bt.forward(ops.pack_rows(torch.from_numpy(x).cuda(), l32), T, l32, zero_tail=True)
    assert T_out == To and np.array_equal(lo.cpu().numpy(), lo_ref)
    assert l2(ops.unpack_rows(yy, T_out).cpu().numpy(), np.where(mo, y.detach().numpy(), 0)) < 6e-3
    dx = bt.backward(tape, ops.pack_rows(torch.from_numpy(Rm).cuda()), need_dx=True)
    dxn = ops.unpack_rows(dx, T).cpu().numpy()
    assert np.all(np.where(m, 0, dxn) == 0)                      # nothing flows into frames beyond the utterance
    assert l2(dxn, xt.grad.numpy()) < 2e-2
    for k, p in blk.named_parameters():
        e = l2(p.grad.cpu().numpy(), stt[k].grad.numpy())
        assert e < 2e-2, (k, e)


def test_scatter_rows_is_the_transpose_of_gather_rows():
    from thunder_speech_b200.train import scatter_rows

    rng = np.random.default_rng(9)
    B, C, T, S = 3, 5, 201, 2
    Ts = (T - 1) // S + 1
    src = ops.pack_rows(torch.from_numpy(rng.standard_normal((B, C, Ts)).astype(np.float32)).cuda())
    up = scatter_rows(src, Ts, S, T)
    u = ops.unpack_rows(up, T).cpu().numpy()
    sv = ops.unpack_rows(src, Ts).cpu().numpy()
    assert np.array_equal(u[:, :, ::S], sv) and np.all(u[:, :, 1::S] == 0) and np.all(up[:, :, T:].float().cpu().numpy() == 0)
    base = ops.pack_rows(torch.from_numpy(rng.standard_normal((B, C, T)).astype(np.float32)).cuda())
    b0 = ops.unpack_rows(base, T).cpu().numpy()
    acc = ops.unpack_rows(scatter_rows(src, Ts, S, T, dst=base), T).cpu().numpy()
    exp = b0.copy()
    exp[:, :, ::S] += sv
    assert np.abs(acc - exp).max() <= 2 ** -8 * np.abs(exp).max()
    assert np.array_equal(ops.unpack_rows(ops.gather_rows(up, T, S, None), Ts).cpu().numpy(), sv)


def test_citrinet_model_training_step_vs_oracle():
    """A whole (small) Citrinet -- stem, stride-1 / stride-2 blocks with SqueezeExcite, 640-wide head, 257-token vocabulary --
    through CTCTrainStep: the loss equals the autograd torch port's, every parameter receives a gradient, the decoder-side
    gradients agree, and a CUDA-graph step runs and lowers the loss."""
    from oracle import ref_torch as RT
    from thunder_speech_b200.citrinet.blocks import CitrinetEncoder

    filters, kernels, strides, V = [64, 64, 64], [5, 7, 9], [1, 2, 1], 257
    blist = synth.citrinet_block_list(filters, kernels, strides, 80)
    st = synth.encoder_state(blist, seed=41, se=True)
    dec = synth.decoder_state(640, V, seed=42)
    x = synth.audio(6, 32000, 43, "noise")
    lens = np.array([32000, 32000, 30000, 27000, 24000, 20000], np.int64)
    rng = np.random.default_rng(44)
    y = rng.integers(0, V - 1, (6, 10)).astype(np.int64)
    y_len = rng.integers(3, 11, 6).astype(np.int64)
    # oracle (fp32 autograd)
    cfgs = R.citrinet_cfgs(filters, kernels, strides, feat_in=80)
    stt = {k: torch.from_numpy(np.asarray(v)).clone() for k, v in st.items()}
    for k, v in stt.items():
        if v.dtype.is_floating_point and "running" not in k:
            v.requires_grad_(True)
    dw_, db_ = torch.from_numpy(dec["weight"]).requires_grad_(True), torch.from_numpy(dec["bias"]).requires_grad_(True)
    with torch.no_grad():
        f, fl = RT.features(torch.from_numpy(x), torch.from_numpy(lens), nfilt=80)
    e, el = RT.encoder(f, fl, cfgs, stt, train=True)
    ref_loss = RT.ctc_loss(torch.nn.functional.conv1d(e, dw_, db_), torch.from_numpy(y), el, torch.from_numpy(y_len), V - 1)
    ref_loss.backward()
    # device
    enc = CitrinetEncoder(filters, kernels, strides, feat_in=80)
    enc.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}, strict=True)
    d = conv1d_decoder(640, V)
    d.load_state_dict({k: torch.from_numpy(v) for k, v in dec.items()})
    tokens = [f"t{i}" for i in range(V - 1)]
    m = CTCModule(enc, d, FilterbankFeatures(nfilt=80, dither=0.0), BatchTextTransformer(tokens)).cuda()
    m.encoder.train(); m.decoder.train()
    step = CTCTrainStep(m, lr=1e-3, use_graph=False)
    batch = (torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(y_len).cuda())
    loss = step.loss_and_grads(*batch)
    assert abs(loss.item() - ref_loss.item()) < 2e-2 * abs(ref_loss.item()), (loss.item(), ref_loss.item())
    for k, p in list(m.encoder.named_parameters()) + list(m.decoder.named_parameters()):
        assert p.grad is not None and torch.isfinite(p.grad).all() and float((p.grad ** 2).sum()) > 0, k
    assert _cos(m.decoder.weight.grad.cpu().numpy(), dw_.grad.numpy()) > 0.99
    assert _cos(m.decoder.bias.grad.cpu().numpy(), db_.grad.numpy()) > 0.99
    last = max(int(k.split(".")[0]) for k in stt)
    for k in stt:                                               # the head block: few layers of amplification
        if k.startswith(f"{last}.") and stt[k].requires_grad:
            assert _cos(dict(m.encoder.named_parameters())[k].grad.cpu().numpy(), stt[k].grad.numpy()) > 0.95, k
    # graph-captured steps lower the loss
    gstep = CTCTrainStep(m, lr=2e-3)
    losses = [gstep.step(*batch).item() for _ in range(5)]
    assert losses[-1] < 0.9 * losses[0], losses


def test_fused_adamw_matches_torch_adamw():
    """ts_adamw (one launch over all parameters) == torch.optim.AdamW with its default hyper-parameters, step for step."""
    from thunder_speech_b200.parallel import flat_grad_views
    from thunder_speech_b200.train import FusedAdamW

    torch.manual_seed(3)
    shapes = [(300, 17, 1), (5,), (129, 64), (1,), (1024, 33)]
    ours = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    flat = flat_grad_views(ours)
    opt = FusedAdamW(ours, flat, lr=3e-3)
    topt = torch.optim.AdamW(ref, lr=3e-3)
    v0 = [p._version for p in ours]
    for it in range(6):
        for p, q in zip(ours, ref):
            g = torch.randn_like(q) * (0.1 + it)
            p.grad.copy_(g)
            q.grad = g.clone()
        opt.step()
        topt.step()
    for p, q in zip(ours, ref):
        assert rel_err(p.detach().cpu().numpy(), q.detach().cpu().numpy())[0] < 2e-6
    assert all(p._version > v for p, v in zip(ours, v0))           # caches keyed on _version see the update
    sd = opt.state_dict()
    opt2 = FusedAdamW(ours, flat, lr=1.0)
    opt2.load_state_dict(sd)
    assert opt2.step_count == 6 and opt2.lr == 3e-3 and torch.equal(opt2.exp_avg, opt.exp_avg)


def test_inference_after_training_uses_the_updated_weights_and_statistics():
    """After CTCTrainStep.step() the eval()-mode forward must reflect the new parameters AND BatchNorm running statistics
    (both are updated by kernels through raw pointers): compare with a freshly built model that loads the trained state_dict."""
    case = _model_case()
    m, step, batch = _device_model(case, lr=2e-3)
    m.eval()
    x = batch[0][:3]
    before, _ = m(x, torch.full((3,), x.shape[1], device="cuda"))
    m.encoder.train(); m.decoder.train()
    for _ in range(3):
        step.step(*batch)
    m.eval()
    after, _ = m(x, torch.full((3,), x.shape[1], device="cuda"))
    assert not torch.equal(before, after)
    filters, kernels = case[0], case[1]
    enc = QuartznetEncoder(filters=filters, kernel_sizes=kernels, repeat_blocks=1)
    enc.load_state_dict(m.encoder.state_dict(), strict=True)
    d = conv1d_decoder(1024, 29)
    d.load_state_dict(m.decoder.state_dict())
    fresh = CTCModule(enc, d, FilterbankFeatures(nfilt=64, dither=0.0), BatchTextTransformer(synth.quartznet_vocab())).cuda().eval()
    expect, _ = fresh(x, torch.full((3,), x.shape[1], device="cuda"))
    assert torch.equal(after, expect)


def test_training_step_takes_texts_like_the_reference():
    """CTCTrainStep.training_step((audio, audio_lengths, texts)) == step() on the encoded labels (module.py:102-127)."""
    case = _model_case()
    m, step, batch = _device_model(case, lr=1e-3)
    texts = ["hello world", "a test", "speech", "b two hundred", "x", "quartz net", "ctc", "gpu"]
    y, yl = m.text_transform.encode(texts, device="cuda")
    m2, step2, _ = _device_model(case, lr=1e-3)
    l1 = step.training_step((batch[0], batch[1], texts))
    l2_ = step2.step(batch[0], batch[1], y, yl)
    assert torch.equal(l1, l2_) and torch.isfinite(l1)
    assert torch.equal(m.decoder.weight, m2.decoder.weight)


def test_training_step_is_an_autograd_node_like_the_reference():
    """CTCModule.training_step(batch) returns a loss with a grad_fn; loss.backward() ACCUMULATES the kernel gradients into
    param.grad exactly like torch autograd would (Lightning's automatic optimisation: backward, optimizer.step,
    zero_grad(set_to_none=True)); a stock torch.optim.AdamW then moves the weights like the fused AdamW kernel does."""
    case = _model_case()
    texts = ["hello world", "a test", "speech", "b two hundred", "x", "quartz net", "ctc", "gpu"]
    m1, step1, batch = _device_model(case, lr=1e-3)
    y, yl = m1.text_transform.encode(texts, device="cuda")
    want_loss = step1.loss_and_grads(batch[0], batch[1], y, yl).clone()
    want = step1.flat.clone()
    m2, _, _ = _device_model(case, lr=1e-3)
    params = list(m2.encoder.parameters()) + list(m2.decoder.parameters())
    opt = torch.optim.AdamW(params, lr=1e-3)
    loss = m2.training_step((batch[0], batch[1], texts), 0)
    assert loss.requires_grad and loss.grad_fn is not None and torch.equal(loss.detach(), want_loss)
    loss.backward()
    got = torch.cat([p.grad.reshape(-1) for p in params])
    assert torch.equal(got, want)
    with pytest.raises(RuntimeError):
        loss.backward()
    # second micro-batch without zero_grad: gradients accumulate (same weights, deterministic kernels -> exactly 2 g)
    (m2.training_step((batch[0], batch[1], texts), 1) * 0.5).backward()
    got = torch.cat([p.grad.reshape(-1) for p in params])
    assert torch.allclose(got, 1.5 * want, rtol=1e-6, atol=0)
    # Lightning's default zero_grad drops the grads; the next backward starts from zero again
    opt.zero_grad(set_to_none=True)
    assert all(p.grad is None for p in params)
    m2.training_step((batch[0], batch[1], texts), 2).backward()
    got = torch.cat([p.grad.reshape(-1) for p in params])
    assert torch.equal(got, want)
    graphs = len(m2.__dict__["_b200_step"]._graphs)
    assert graphs == 1                               # re-attaching the gradient views did not force a re-capture
    opt.step()
    step1.opt.step()
    for (k, a), (_, b) in zip(m2.named_parameters(), m1.named_parameters()):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), k
    # and the updated weights are what the next step trains on
    l3 = m2.training_step((batch[0], batch[1], texts), 3)
    assert torch.isfinite(l3) and float(l3.detach()) != float(want_loss)


def test_config5_full_size_training_step_properties():
    """BASELINE config 5 at its full size (QuartzNet 15x5, 32 x 15 s per GPU, ragged lengths, 120-character labels): the
    captured-graph step and the eager step of an identical model give the same loss and bit-identical gradients (894 kernel
    launches, forked weight-gradient stream, split-K reductions: no order dependence anywhere), the optimiser moves every
    parameter, and repeated steps on the batch bring the loss down."""
    from thunder_speech_b200.runner import build_model

    B, N, L = 32, 15 * 16000, 120
    rng = np.random.default_rng(5)
    audio = torch.from_numpy(synth.audio(B, N, 91, "noise")).cuda()
    lens = torch.from_numpy(synth.ragged_lengths(B, N, 7)).cuda()
    y = torch.from_numpy(rng.integers(0, 28, (B, L)).astype(np.int64)).cuda()
    yl = torch.from_numpy(rng.integers(L // 2, L + 1, B).astype(np.int64)).cuda()
    steps = []
    for use_graph in (True, False):
        m = build_model("quartznet15x5", torch.device("cuda"), seed=0)
        m.encoder.train(); m.decoder.train()             # front-end stays in eval(): no dither noise in this comparison
        steps.append((m, CTCTrainStep(m, lr=3e-4, use_graph=use_graph)))
    (m1, s1), (m2, s2) = steps
    l1 = s1.loss_and_grads(audio, lens, y, yl).clone()
    l2 = s2.loss_and_grads(audio, lens, y, yl).clone()
    assert torch.isfinite(l1) and torch.equal(l1, l2)
    assert torch.equal(s1.flat, s2.flat) and torch.isfinite(s1.flat).all() and float(s1.flat.norm()) > 0
    for (k, a), (_, b) in zip(m1.named_buffers(), m2.named_buffers()):
        assert torch.equal(a, b), k                     # BatchNorm running statistics / num_batches_tracked
    before = [p.detach().clone() for p in s1.params]
    s1.opt.step()
    assert all(not torch.equal(a, p) for a, p in zip(before, s1.params))
    losses = [float(l1)] + [float(s1.step(audio, lens, y, yl)) for _ in range(12)]
    assert all(np.isfinite(losses)) and min(losses[6:]) < losses[0], losses


def test_validation_step_loss_and_error_rates():
    """CTCModule.validation_step (module.py:130-163): eval() forward, CTC loss == F.ctc_loss(log_softmax(logits)) with the
    encoder's output lengths (reduction mean, zero_infinity), predictions decoded from every frame, CER / WER accumulated
    against the label strings."""
    from thunder_speech_b200.metrics import CharErrorRate, WordErrorRate

    case = _model_case()
    m, _, batch = _device_model(case)
    texts = ["hello world", "a test", "speech", "b two hundred", "x", "quartz net", "ctc", "gpu"]
    with pytest.raises(NotImplementedError):
        m.validation_step((batch[0], batch[1], texts), 0)          # still in train() mode
    m.eval()
    loss = m.validation_step((batch[0], batch[1], texts), 0)
    logits, out_len = m(batch[0], batch[1])
    y, yl = m.text_transform.encode(texts, device="cuda")
    ref = torch.nn.functional.ctc_loss(torch.log_softmax(logits.double(), 1).permute(2, 0, 1), y, out_len, yl,
                                       blank=m.text_transform.vocab.blank_idx, reduction="mean", zero_infinity=True)
    assert abs(float(loss) - float(ref)) < 1e-4 * abs(float(ref)), (float(loss), float(ref))
    preds = m.text_transform.decode_prediction(logits.argmax(1))    # the reference's formula on the same logits
    cer, wer = CharErrorRate(), WordErrorRate()
    cer(preds, texts); wer(preds, texts)
    assert m.validation_cer.compute() == cer.compute() and m.validation_wer.compute() == wer.compute()
    m.validation_step((batch[0], batch[1], texts), 1)              # accumulates over batches
    assert m.validation_cer.total == 2 * cer.total and m.validation_cer.compute() == cer.compute()


def test_fit_stream_equals_step_by_step():
    """CTCTrainStep.fit_stream (pinned host batches, copies on a side stream, loss read back one step late) produces the
    same losses and the same weights as calling step() on device tensors, for pipeline depths 1-3."""
    case = _model_case()
    m0, step0, batch = _device_model(case, lr=1e-3)
    want = [float(step0.step(*batch)) for _ in range(4)]
    host = tuple(t.cpu().pin_memory() for t in batch)
    for depth in (1, 2, 3):
        m, step, _ = _device_model(case, lr=1e-3)
        got = list(step.fit_stream((host for _ in range(4)), depth=depth))
        assert got == want, (depth, got, want)
        sd0 = m0.state_dict()
        for k, v in m.state_dict().items():
            assert torch.equal(v, sd0[k]), (depth, k)


def test_random_block_configurations_training_vs_oracle():
    """Random block configurations through BlockTrainer (forward, dx, every parameter gradient) against the autograd torch
    port with bf16 storage simulated: odd channel counts (the <= 128-row GEMM kernel, partial tiles), K = 1 .. 33, dilation 2,
    stride-2 Citrinet blocks, with / without residual and SqueezeExcite -- the fallback kernels get the same scrutiny as the
    QuartzNet-15x5 shapes."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    from oracle import ref_torch as RT
    from thunder_speech_b200.citrinet.blocks import CitrinetBlock

    @settings(max_examples=16, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(kind=st.sampled_from(["quartznet", "citrinet"]), cin=st.sampled_from([8, 16, 40, 64, 136]),
           cout=st.sampled_from([8, 24, 64, 144, 256]), rep=st.integers(1, 3), K=st.sampled_from([1, 3, 5, 11, 33]),
           stride=st.sampled_from([1, 1, 2]), dil=st.sampled_from([1, 1, 2]), res=st.booleans(), B=st.integers(2, 5),
           T=st.integers(40, 200), seed=st.integers(0, 1000))
    def check(kind, cin, cout, rep, K, stride, dil, res, B, T, seed):
        if stride > 1 and (dil > 1 or kind == "quartznet"):
            return     # QuartzNet strides every sub-block (only its stem, which needs no dx); stride with dilation is a ValueError
        rng = np.random.Generator(np.random.PCG64(seed))
        se = kind == "citrinet"
        st_ = synth.block_state(rng, "", cin, cout, rep, K, res, True, se=se)
        x = np.maximum(rng.standard_normal((B, cin, T)), 0).astype(np.float32)
        lens = np.sort(rng.integers(T // 2, T + 1, B))[::-1].astype(np.int64).copy()
        lens[0] = T
        To = (T - 1) // stride + 1
        lo_ref = (lens - 1) // stride + 1
        m = (np.arange(T)[None, :] < lens[:, None])[:, None, :]
        mo = (np.arange(To)[None, :] < lo_ref[:, None])[:, None, :]
        Rm = np.where(mo, rng.standard_normal((B, cout, To)), 0).astype(np.float32)
        cfg = R.BlockCfg(cin, cout, repeat=rep, kernel_size=K, stride=stride, dilation=dil, residual=res, separable=True,
                         kind=kind)
        sd = {k: torch.from_numpy(np.asarray(v)).clone() for k, v in st_.items()}
        for k, v in sd.items():
            if v.dtype.is_floating_point and "running" not in k:
                v.requires_grad_(True)
        xt = torch.from_numpy(np.where(m, x, 0).astype(np.float32)).requires_grad_(True)
        y, yl = RT.block(xt, torch.from_numpy(lens), cfg, sd, "", train=True, store=RT.bf16_store)
        (y * torch.from_numpy(Rm)).sum().backward()
        cls = QuartznetBlock if kind == "quartznet" else CitrinetBlock
        blk = cls(cin, cout, repeat=rep, kernel_size=(K,), stride=(stride,), dilation=(dil,), residual=res, separable=True)
        blk.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st_.items()}, strict=True)
        blk = blk.cuda().train()
        bt = BlockTrainer(blk)
        l32 = torch.from_numpy(lens.astype(np.int32)).cuda()
        yy, T_out, lo, tape = bt.forward(ops.pack_rows(torch.from_numpy(x).cuda(), l32), T, l32, zero_tail=True)
        tag = (kind, cin, cout, rep, K, stride, dil, res, B, T, seed)
        assert T_out == To and np.array_equal(lo.cpu().numpy(), lo_ref), tag
        assert l2(ops.unpack_rows(yy, T_out).cpu().numpy(), np.where(mo, y.detach().numpy(), 0)) < 8e-3, tag
        dx = bt.backward(tape, ops.pack_rows(torch.from_numpy(Rm).cuda()), need_dx=True)
        assert l2(ops.unpack_rows(dx, T).cpu().numpy(), xt.grad.numpy()) < 3e-2, tag
        for k, p in blk.named_parameters():
            e = l2(p.grad.cpu().numpy(), sd[k].grad.numpy())
            assert e < 3e-2, tag + (k, e)

    check()
