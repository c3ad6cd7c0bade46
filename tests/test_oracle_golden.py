"""Pins the numpy oracle (oracle/ref_numpy.py) against outputs of the reference's own PyTorch
modules (tests/golden/*.npz, produced by oracle/make_golden.py) and against the exact
known-answer cases of the reference test-suite.  CPU only."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import ref_numpy as R
from thunder_speech_b200 import synth

FEATURE_CASES = [("qn_noise", "noise"), ("qn_tones", "tones"), ("cn_noise", "noise"), ("short", "noise")]


@pytest.mark.parametrize("name,kind", FEATURE_CASES)
def test_features_match_reference(golden_features, name, kind):
    g = golden_features
    nfilt, B, N, seed = [int(v) for v in g[f"{name}.meta"]]
    x = synth.audio(B, N, seed, kind)
    lens = synth.ragged_lengths(B, N, seed + 100)
    assert (lens == g[f"{name}.in_lengths"]).all()
    feats, flen, inter = R.filterbank_features(x, lens, nfilt=nfilt, return_intermediate=True)
    assert (flen == g[f"{name}.lengths"]).all()
    assert feats.shape == g[f"{name}.features"].shape
    emax, el2 = rel_err(inter["logmel"], g[f"{name}.logmel"])
    assert emax < 2e-5 and el2 < 2e-5, (emax, el2)
    emax, el2 = rel_err(feats, g[f"{name}.features"])
    assert emax < 2e-5 and el2 < 2e-5, (emax, el2)
    # float lengths (asr_collate contract, data/dataloader_utils.py:29) give the same result
    feats2, flen2 = R.filterbank_features(x, lens.astype(np.float32), nfilt=nfilt)
    assert (flen2 == flen).all() and np.array_equal(feats2, feats)


def test_mel_filterbank_and_window(golden_features):
    g = golden_features
    for n in (64, 80):
        fb = R.mel_filterbank(257, n, 16000)
        assert fb.shape == (n, 257)
        # torchaudio builds the bank in float32: agreement to a few float32 ulps of the peak weight
        assert np.abs(fb - g[f"fb{n}"]).max() < 5e-6 * np.abs(g[f"fb{n}"]).max()
        # ~97% zeros, at most 2 filters per FFT bin
        assert ((fb > 0).sum(0) <= 2).all()
    assert np.abs(R.hann_window(320) - g["window320"]).max() < 1e-6


def test_feature_shape_laws():
    # tests/quartznet/test_transform_qn.py:179-189: bins = 1+n_fft//2, frames = 1+N//hop
    x = synth.audio(2, 1234, 3)
    f, fl = R.filterbank_features(x, np.array([1234, 1234]))
    assert f.shape == (2, 64, 1 + 1234 // 160)
    assert (fl == 1 + 1234 // 160).all()
    # tests/quartznet/test_transform_qn.py:43-51: normalised mean~0 / std~1
    assert np.abs(f.mean(-1)).max() < 0.1 and np.abs(f.std(-1) - 1).max() < 0.1
    with pytest.raises(ValueError):
        R.filterbank_features(x, np.array([1234, 1234]), n_window_size=0)


def test_mask_and_length_helpers(golden_helpers):
    g = golden_helpers
    assert np.array_equal(R.lengths_to_mask(g["mask_lengths"], 6), g["mask"])
    # tests/test_blocks.py:56-68 exact truth table
    assert np.array_equal(R.lengths_to_mask(np.array([1, 2, 3]), 3),
                          np.array([[1, 0, 0], [1, 1, 0], [1, 1, 1]], bool))
    for row, (k, s, d, p) in zip(g["seq_len_out"], g["same_padding"]):
        assert R.get_same_padding(int(k), int(s), int(d)) == int(p)
        assert np.array_equal(R.conv_out_len(g["seq_len_in"], int(k), int(s), int(p), int(d)), row)
    with pytest.raises(ValueError):  # tests/quartznet/test_blocks_qn.py:109-116
        R.get_same_padding(3, 2, 2)
    # tests/quartznet/test_blocks_qn.py:89-106: same padding => ceil(T/stride)
    for k in (1, 5, 33):
        for s in (1, 2):
            p = R.get_same_padding(k, s, 1)
            for T in (10, 11, 751):
                assert R.conv_out_len(np.array([T]), k, s, p, 1)[0] == -(-T // s)


BLOCK_CASES = [
    ("qn_res", "quartznet", dict(in_channels=16, out_channels=24, repeat=3, kernel_size=5, stride=1, dilation=1,
                                 residual=True, separable=True), 3, 50),
    ("qn_stem", "quartznet", dict(in_channels=8, out_channels=16, repeat=1, kernel_size=33, stride=2, dilation=1,
                                  residual=False, separable=True), 2, 101),
    ("qn_dil", "quartznet", dict(in_channels=16, out_channels=16, repeat=1, kernel_size=87, stride=1, dilation=2,
                                 residual=False, separable=True), 2, 120),
    ("qn_k1", "quartznet", dict(in_channels=16, out_channels=32, repeat=1, kernel_size=1, stride=1, dilation=1,
                                residual=False, separable=False), 2, 37),
    ("qn_stride_res", "quartznet", dict(in_channels=8, out_channels=8, repeat=2, kernel_size=3, stride=2,
                                        dilation=1, residual=True, separable=True), 2, 64),
    ("cn_res", "citrinet", dict(in_channels=16, out_channels=32, repeat=5, kernel_size=11, stride=1, dilation=1,
                                residual=True, separable=True), 3, 77),
    ("cn_stride", "citrinet", dict(in_channels=32, out_channels=32, repeat=5, kernel_size=13, stride=2,
                                   dilation=1, residual=True, separable=True), 2, 91),
    ("cn_stem", "citrinet", dict(in_channels=80, out_channels=256, repeat=1, kernel_size=5, stride=1, dilation=1,
                                 residual=False, separable=True), 2, 33),
    # non-separable convolutions with kernel_size > 1: QuartznetBlock's DEFAULT (quartznet/blocks.py:232-243)
    ("qn_full", "quartznet", dict(in_channels=16, out_channels=24, repeat=2, kernel_size=11, stride=1, dilation=1,
                                  residual=True, separable=False), 2, 61),
    ("qn_full_stride", "quartznet", dict(in_channels=8, out_channels=16, repeat=1, kernel_size=5, stride=2, dilation=1,
                                         residual=False, separable=False), 2, 50),
]


def block_case_inputs(ci):
    name, kind, cfg, B, T = BLOCK_CASES[ci]
    rng = np.random.Generator(np.random.PCG64(1000 + ci))
    st = synth.block_state(rng, "", cfg["in_channels"], cfg["out_channels"], cfg["repeat"], cfg["kernel_size"],
                           cfg["residual"], cfg["separable"], se=(kind == "citrinet"))
    x = rng.standard_normal((B, cfg["in_channels"], T)).astype(np.float32)
    lens = np.sort(rng.integers(T // 2, T + 1, B))[::-1].astype(np.int64).copy()
    lens[0] = T
    return name, kind, cfg, st, x, lens


@pytest.mark.parametrize("ci", range(len(BLOCK_CASES)))
def test_blocks_match_reference(golden_blocks, ci):
    name, kind, cfg, st, x, lens = block_case_inputs(ci)
    assert (lens == golden_blocks[f"{name}.in_lengths"]).all()
    bc = R.BlockCfg(kind=kind, **cfg)
    y, yl = R.block_forward(x, lens, bc, st)
    assert np.array_equal(yl, golden_blocks[f"{name}.out_lengths"])
    assert y.shape == golden_blocks[f"{name}.out"].shape
    emax, el2 = rel_err(y, golden_blocks[f"{name}.out"])
    assert emax < 1e-5 and el2 < 1e-5, (name, emax, el2)


def test_squeeze_excite_matches_reference(golden_blocks):
    rng = np.random.Generator(np.random.PCG64(2000))
    w1 = synth._uniform(rng, (4, 32), 2.0 / 32)
    w2 = synth._uniform(rng, (32, 4), 4.0 / 4)
    x = rng.standard_normal((3, 32, 41)).astype(np.float32)
    emax, _ = rel_err(R.squeeze_excite(x, w1, w2), golden_blocks["se.out"])
    assert emax < 1e-6


def qn5x5_model():
    blocks = synth.quartznet_block_list(repeat_blocks=1)
    return (R.quartznet_cfgs(repeat_blocks=1), synth.encoder_state(blocks, seed=5),
            synth.decoder_state(1024, 29, seed=6), R.Vocab(synth.quartznet_vocab()))


def cn_small_model():
    f, k, s = [64, 64, 96, 96], [11, 13, 15, 17], [2, 1, 2, 2]
    blocks = synth.citrinet_block_list(f, k, s, feat_in=80)
    return (R.citrinet_cfgs(f, k, s, feat_in=80), synth.encoder_state(blocks, seed=8, se=True),
            synth.decoder_state(640, 65, seed=9), R.Vocab(synth.citrinet_vocab(64)))


@pytest.mark.parametrize("model,nfilt,tag2", [("qn5x5", 64, 7777), ("cn", 80, 6001)])
def test_end_to_end_matches_reference(golden_e2e, model, nfilt, tag2):
    g = golden_e2e
    cfgs, st, dec, vocab = qn5x5_model() if model == "qn5x5" else cn_small_model()
    x = synth.audio(2, 12000, 21, "tones")
    for tag, lens in (("full", np.array([12000, 12000])), ("ragged", np.array([12000, tag2]))):
        f, fl = R.filterbank_features(x, lens, nfilt=nfilt)
        e, el = R.encoder_forward(f, fl, cfgs, st)
        logits = R.decoder_forward(e, dec["weight"], dec["bias"])
        assert np.array_equal(el, g[f"{model}.{tag}.out_lengths"])
        emax, el2 = rel_err(logits, g[f"{model}.{tag}.logits"])
        assert emax < 5e-5 and el2 < 5e-5, (model, tag, emax, el2)
        ids = R.greedy_argmax(logits)
        ref_ids = g[f"{model}.{tag}.ids"]
        # own-logits agreement (near-ties may flip); decoding the SAME ids must be exact
        assert (ids == ref_ids).mean() > 0.99
        assert R.decode_prediction(ref_ids, vocab) == list(g[f"{model}.{tag}.text"])
    # predict() == forward with full lengths (module.py:98-100)
    texts, allv = R.predict(x, cfgs, st, dec["weight"], dec["bias"], vocab, dict(nfilt=nfilt), return_all=True)
    assert np.array_equal(allv["out_lengths"], g[f"{model}.full.out_lengths"])


def test_greedy_decode_known_answers(golden_decode):
    g = golden_decode
    vocab = R.Vocab(synth.quartznet_vocab())
    assert vocab.blank_idx == 28 and len(vocab.itos) == 29
    assert R.decode_prediction(g["ids"], vocab) == list(g["text"])
    vocab2 = R.Vocab(synth.citrinet_vocab(64))
    assert R.decode_prediction(g["ids_bpe"], vocab2) == list(g["text_bpe"])
    assert np.array_equal(R.greedy_argmax(g["argmax_logits"]), g["argmax_ids"])
    # reference tests/text/test_transforms.py:59-91 (vocab with unk/bos/eos, blank==pad)
    v = R.Vocab([" "] + [chr(c) for c in range(97, 123)], "<blank>", "<blank>", "<unk>", "<bos>", "<eos>")
    V = len(v.itos)
    blank = np.zeros((1, V, 100), np.float32)
    blank[:, v.blank_idx, :] = 1
    assert R.decode_prediction(R.greedy_argmax(blank), v) == [""]
    a, b = v.itos.index("a"), v.itos.index("b")
    x = blank.copy(); x[:, a, :10] = 2; x[:, b, 15:20] = 2
    assert R.decode_prediction(R.greedy_argmax(x), v) == ["ab"]
    x = blank.copy(); x[:, a, :10] = 2; x[:, a, 15:20] = 2
    assert R.decode_prediction(R.greedy_argmax(x), v) == ["aa"]
    # tests/text/test_vocab.py:108-112: blank appended last
    assert R.Vocab(["a", "b", "c"]).blank_idx == 3


# ------------------------------------------------------------------------------------------------ training step
def test_torch_port_training_blocks_match_reference():
    """The autograd-enabled torch port (oracle of the training step) reproduces the reference's train()-mode block
    outputs, input/parameter gradients and BatchNorm running-statistic updates (tests/golden/train.npz)."""
    import torch

    from oracle import ref_torch as RT
    from oracle.make_golden_train import TRAIN_BLOCK_CASES, block_case

    g = np.load("tests/golden/train.npz", allow_pickle=False)
    for ci in range(len(TRAIN_BLOCK_CASES)):
        name, cfg, st, x, lens, Rm = block_case(ci)
        stt = {k: torch.from_numpy(np.asarray(v)).clone() for k, v in st.items()}
        for k, v in stt.items():
            if v.dtype.is_floating_point and "running" not in k:
                v.requires_grad_(True)
        xt = torch.from_numpy(x).requires_grad_(True)
        bc = R.BlockCfg(kind="quartznet", **cfg)
        y, yl = RT.block(xt, torch.from_numpy(lens), bc, stt, "", train=True)
        (y * torch.from_numpy(Rm)).sum().backward()
        assert np.array_equal(yl.numpy(), g[f"{name}.out_lengths"])
        assert rel_err(y.detach().numpy(), g[f"{name}.out"])[0] < 1e-5
        assert rel_err(xt.grad.numpy(), g[f"{name}.dx"])[0] < 1e-4
        for k, v in stt.items():
            if v.requires_grad:
                e = rel_err(v.grad.numpy(), g[f"{name}.grad.{k}"])[0]
                assert e < 2e-4, (name, k, e)
            elif "running" in k:
                assert rel_err(v.numpy(), g[f"{name}.buf.{k}"])[0] < 1e-5, (name, k)


def test_torch_port_training_model_matches_reference():
    import torch

    from oracle import ref_torch as RT
    from oracle.make_golden_train import tiny_model_case

    g = np.load("tests/golden/train.npz", allow_pickle=False)
    filters, kernels, st, dec, x, lens, y, y_len = tiny_model_case()
    cfgs = R.quartznet_cfgs(filters=filters, kernel_sizes=kernels, repeat_blocks=1)
    stt = {k: torch.from_numpy(np.asarray(v)).clone() for k, v in st.items()}
    for k, v in stt.items():
        if v.dtype.is_floating_point and "running" not in k:
            v.requires_grad_(True)
    dw = torch.from_numpy(dec["weight"]).requires_grad_(True)
    db = torch.from_numpy(dec["bias"]).requires_grad_(True)
    with torch.no_grad():
        f, fl = RT.features(torch.from_numpy(x), torch.from_numpy(lens))
    e, el = RT.encoder(f, fl, cfgs, stt, train=True)
    logits = torch.nn.functional.conv1d(e, dw, db)
    loss = RT.ctc_loss(logits, torch.from_numpy(y), el, torch.from_numpy(y_len), 28)
    loss.backward()
    assert abs(loss.item() - float(g["model.loss"])) < 1e-4 * abs(float(g["model.loss"]))
    assert np.array_equal(el.numpy(), g["model.out_lengths"])
    assert rel_err(logits.detach().numpy(), g["model.logits"])[0] < 1e-4
    rng = np.random.Generator(np.random.PCG64(9999))
    grads = {k: v.grad for k, v in stt.items() if v.requires_grad}
    grads["decoder.weight"], grads["decoder.bias"] = dw.grad, db.grad
    for k in g["model.param_names"]:
        gr = grads[str(k)].numpy()
        r = rng.standard_normal(gr.shape).astype(np.float32)
        proj, norm = g[f"model.gproj.{k}"]
        assert abs(np.sqrt((gr * gr).sum()) - norm) <= 1e-2 * norm + 1e-7, k
        # random projection of the gradient: two fp32 evaluations of a 25-layer train-mode BN stack agree only to a few 1e-2 of
        # the gradient norm on the earliest layers (block-level gradients above are pinned to 2e-4)
        assert abs(float((gr * r).sum()) - proj) <= 1e-1 * norm + 1e-6, k


def test_ingest_oracle_vs_reference_goldens():
    """oracle.ref_numpy.preprocess_audio (mono mix, DC removal, torchaudio-style sinc resample) against outputs of the
    reference's own AudioFileLoader.preprocess_audio (tests/golden/ingest.npz, oracle/make_golden_ingest.py)."""
    from oracle.make_golden_ingest import CASES, clip

    g = np.load("tests/golden/ingest.npz")
    for name, ch, sr, secs in CASES:
        pcm = clip(name, ch, sr, secs)
        y = R.preprocess_audio(pcm.astype(np.float32) / 32768.0, sr)
        ref = g[f"{name}.out"]
        assert y.shape == ref.shape, name
        assert np.abs(y - ref).max() <= 5e-5 * np.abs(ref).max(), (name, np.abs(y - ref).max() / np.abs(ref).max())
    k, width, o, n = R.sinc_resample_kernel(44100, 16000)
    assert (o, n, width, k.shape) == (441, 160, 17, (160, 475))
    with pytest.raises(RuntimeError):
        R.preprocess_audio(np.zeros((2, 100), np.float32), 16000, force_mono=False)


def test_spec_augment_oracle_vs_reference_goldens():
    """oracle spec_augment / spec_cutout fed with the host generator's draws == the reference's own modules under the same
    seed (tests/golden/augment.npz, oracle/make_golden_augment.py)."""
    import torch

    from oracle.make_golden_augment import CASES, case_input

    g = np.load("tests/golden/augment.npz")
    for name, kind, kw, shape, seed in CASES:
        x = case_input(name, shape)
        torch.manual_seed(seed)
        draw = lambda: float(torch.rand(1))
        y = R.spec_augment(x, draw, **kw) if kind == "augment" else R.spec_cutout(x, draw, **kw)
        assert np.array_equal(y, g[f"{name}.out"]), name
