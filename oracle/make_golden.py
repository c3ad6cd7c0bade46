"""Generate ``tests/golden/*.npz`` by running the REFERENCE's own PyTorch modules (CPU, fp32).

Run in the build container only (``python oracle/make_golden.py``): it imports the unmodified
reference from ``/root/reference/src`` through a stub for ``thunder/__init__.py`` (which otherwise
calls ``importlib.metadata.version("thunder-speech")`` on a package that is not installed,
src/thunder/__init__.py:1-6).  ``/root/reference`` does not exist on the GPU box, so the outputs are
committed as small fixtures; inputs and weights are NOT stored -- they are regenerated from integer
seeds by ``thunder_speech_b200.synth`` (numpy PCG64), which is what makes the fixtures small.

Nothing here is product code; nothing in the product imports it.
"""
from __future__ import annotations

import os
import sys
import types
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_SRC = os.environ.get("THUNDER_REF_SRC", "/root/reference/src")


def import_reference():
    stub = types.ModuleType("thunder")
    stub.__path__ = [os.path.join(REF_SRC, "thunder")]
    sys.modules["thunder"] = stub
    import thunder.blocks  # noqa: F401
    import thunder.citrinet.blocks  # noqa: F401
    import thunder.quartznet.blocks  # noqa: F401
    import thunder.quartznet.transform  # noqa: F401
    import thunder.text_processing.transform  # noqa: F401
    return sys.modules["thunder"]


def load_state(module: torch.nn.Module, state: dict, prefix: str = ""):
    sd = {}
    for k, v in state.items():
        if prefix and not k.startswith(prefix):
            continue
        sd[k[len(prefix):]] = torch.from_numpy(np.asarray(v))
    module.load_state_dict(sd, strict=True)
    return module.eval()


def main():
    warnings.filterwarnings("ignore")
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    import_reference()
    from thunder.blocks import conv1d_decoder, get_same_padding, lengths_to_mask
    from thunder.citrinet.blocks import CitrinetBlock, CitrinetEncoder, SqueezeExcite
    from thunder.quartznet.blocks import MaskedConv1d, QuartznetBlock, QuartznetEncoder
    from thunder.quartznet.transform import FilterbankFeatures
    from thunder.text_processing.transform import BatchTextTransformer

    from thunder_speech_b200 import synth

    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)

    # ---------------------------------------------------------------- features
    feats = {}
    for name, nfilt, kind, B, N, seed in [
        ("qn_noise", 64, "noise", 3, 4800, 11),
        ("qn_tones", 64, "tones", 2, 8000, 12),
        ("cn_noise", 80, "noise", 2, 3333, 13),   # N not a multiple of hop
        ("short", 64, "noise", 2, 700, 14),        # shorter than the reflect region of two frames
    ]:
        x = synth.audio(B, N, seed, kind)
        lens = synth.ragged_lengths(B, N, seed + 100)
        fbank = FilterbankFeatures(nfilt=nfilt).eval()
        f, fl = fbank(torch.from_numpy(x), torch.from_numpy(lens))
        # intermediate: un-normalised log-mel (first three stages)
        y, yl = x, torch.from_numpy(lens)
        y = torch.from_numpy(x)
        for m in list(fbank.children())[:3]:
            y, yl = m(y, yl)
        feats[f"{name}.features"] = f.numpy()
        feats[f"{name}.lengths"] = fl.numpy()
        feats[f"{name}.logmel"] = y.numpy()
        feats[f"{name}.meta"] = np.array([nfilt, B, N, seed], np.int64)
        feats[f"{name}.in_lengths"] = lens
    # float lengths (asr_collate contract) give the same result
    fb64 = FilterbankFeatures(nfilt=64).eval()
    feats["fb64"] = fb64[2].layer[0].fb[0].numpy()
    feats["fb80"] = FilterbankFeatures(nfilt=80).eval()[2].layer[0].fb[0].numpy()
    feats["window320"] = fb64[1].window.numpy()
    np.savez_compressed(os.path.join(out_dir, "features.npz"), **feats)

    # ---------------------------------------------------------------- helpers: exact known answers
    helpers = {}
    helpers["mask_lengths"] = np.array([3, 0, 5, 7, 2], np.int64)
    helpers["mask"] = lengths_to_mask(torch.from_numpy(helpers["mask_lengths"]), 6).numpy()
    pads = []
    for k in (1, 5, 11, 33, 39, 41, 87):
        for s, d in ((1, 1), (2, 1), (1, 2)):
            pads.append((k, s, d, get_same_padding(k, s, d)))
    helpers["same_padding"] = np.array(pads, np.int64)
    sl = []
    for (k, s, d, p) in pads:
        mc = MaskedConv1d(1, 1, k, stride=s, padding=p, dilation=d)
        lens = torch.tensor([1, 2, 17, 100, 751, 2001])
        sl.append(mc.get_seq_len(lens).numpy())
    helpers["seq_len_in"] = np.array([1, 2, 17, 100, 751, 2001], np.int64)
    helpers["seq_len_out"] = np.stack(sl)
    np.savez_compressed(os.path.join(out_dir, "helpers.npz"), **helpers)

    # ---------------------------------------------------------------- blocks
    blocks = {}
    block_cases = [
        # name, kind, cfg, B, T
        ("qn_res", "quartznet", dict(in_channels=16, out_channels=24, repeat=3, kernel_size=5, stride=1,
                                     dilation=1, residual=True, separable=True), 3, 50),
        ("qn_stem", "quartznet", dict(in_channels=8, out_channels=16, repeat=1, kernel_size=33, stride=2,
                                      dilation=1, residual=False, separable=True), 2, 101),
        ("qn_dil", "quartznet", dict(in_channels=16, out_channels=16, repeat=1, kernel_size=87, stride=1,
                                     dilation=2, residual=False, separable=True), 2, 120),
        ("qn_k1", "quartznet", dict(in_channels=16, out_channels=32, repeat=1, kernel_size=1, stride=1,
                                    dilation=1, residual=False, separable=False), 2, 37),
        ("qn_stride_res", "quartznet", dict(in_channels=8, out_channels=8, repeat=2, kernel_size=3, stride=2,
                                            dilation=1, residual=True, separable=True), 2, 64),
        ("cn_res", "citrinet", dict(in_channels=16, out_channels=32, repeat=5, kernel_size=11, stride=1,
                                    dilation=1, residual=True, separable=True), 3, 77),
        ("cn_stride", "citrinet", dict(in_channels=32, out_channels=32, repeat=5, kernel_size=13, stride=2,
                                       dilation=1, residual=True, separable=True), 2, 91),
        ("cn_stem", "citrinet", dict(in_channels=80, out_channels=256, repeat=1, kernel_size=5, stride=1,
                                     dilation=1, residual=False, separable=True), 2, 33),
        ("qn_full", "quartznet", dict(in_channels=16, out_channels=24, repeat=2, kernel_size=11, stride=1,
                                      dilation=1, residual=True, separable=False), 2, 61),
        ("qn_full_stride", "quartznet", dict(in_channels=8, out_channels=16, repeat=1, kernel_size=5, stride=2,
                                             dilation=1, residual=False, separable=False), 2, 50),
    ]
    for ci, (name, kind, cfg, B, T) in enumerate(block_cases):
        rng = np.random.Generator(np.random.PCG64(1000 + ci))
        st = synth.block_state(rng, "", cfg["in_channels"], cfg["out_channels"], cfg["repeat"],
                               cfg["kernel_size"], cfg["residual"], cfg["separable"], se=(kind == "citrinet"))
        x = rng.standard_normal((B, cfg["in_channels"], T)).astype(np.float32)
        lens = np.sort(rng.integers(T // 2, T + 1, B))[::-1].astype(np.int64).copy()
        lens[0] = T
        cls = QuartznetBlock if kind == "quartznet" else CitrinetBlock
        mod = cls(cfg["in_channels"], cfg["out_channels"], repeat=cfg["repeat"],
                  kernel_size=(cfg["kernel_size"],), stride=(cfg["stride"],), dilation=(cfg["dilation"],),
                  residual=cfg["residual"], separable=cfg["separable"])
        load_state(mod, st)
        y, yl = mod(torch.from_numpy(x), torch.from_numpy(lens))
        blocks[f"{name}.out"] = y.numpy()
        blocks[f"{name}.out_lengths"] = yl.numpy()
        blocks[f"{name}.in_lengths"] = lens
    # SqueezeExcite alone
    rng = np.random.Generator(np.random.PCG64(2000))
    se = SqueezeExcite(32, 8).eval()
    w1 = synth._uniform(rng, (4, 32), 2.0 / 32)
    w2 = synth._uniform(rng, (32, 4), 4.0 / 4)
    se.load_state_dict({"fc.0.weight": torch.from_numpy(w1), "fc.2.weight": torch.from_numpy(w2)})
    x = rng.standard_normal((3, 32, 41)).astype(np.float32)
    blocks["se.out"] = se(torch.from_numpy(x)).numpy()
    np.savez_compressed(os.path.join(out_dir, "blocks.npz"), **blocks)

    # ---------------------------------------------------------------- end to end (QuartzNet 5x5, Citrinet small)
    e2e = {}
    x = synth.audio(2, 12000, 21, "tones")
    lens_full = np.full((2,), 12000, np.int64)
    qn_blocks = synth.quartznet_block_list(repeat_blocks=1)
    st = synth.encoder_state(qn_blocks, seed=5)
    dec = synth.decoder_state(1024, 29, seed=6)
    enc = load_state(QuartznetEncoder(repeat_blocks=1), st)
    decoder = conv1d_decoder(1024, 29)
    decoder.load_state_dict({"weight": torch.from_numpy(dec["weight"]), "bias": torch.from_numpy(dec["bias"])})
    fbank = FilterbankFeatures(nfilt=64).eval()
    tt = BatchTextTransformer(tokens=synth.quartznet_vocab())
    for tag, lens in (("full", lens_full), ("ragged", np.array([12000, 7777], np.int64))):
        f, fl = fbank(torch.from_numpy(x), torch.from_numpy(lens))
        e, el = enc(f, fl)
        logits = decoder.eval()(e)
        ids = logits.argmax(1)
        e2e[f"qn5x5.{tag}.logits"] = logits.numpy()
        e2e[f"qn5x5.{tag}.out_lengths"] = el.numpy()
        e2e[f"qn5x5.{tag}.ids"] = ids.numpy()
        e2e[f"qn5x5.{tag}.text"] = np.array(tt.decode_prediction(ids))
        e2e[f"qn5x5.{tag}.enc_absmax"] = np.array(float(e.abs().max()))

    # Citrinet: reduced-width body with the 1024 variant's structure (3 strided blocks)
    cn_filters, cn_k, cn_s = [64, 64, 96, 96], [11, 13, 15, 17], [2, 1, 2, 2]
    cn_blocks = synth.citrinet_block_list(cn_filters, cn_k, cn_s, feat_in=80)
    st = synth.encoder_state(cn_blocks, seed=8, se=True)
    dec = synth.decoder_state(640, 65, seed=9)
    enc = load_state(CitrinetEncoder(cn_filters, cn_k, cn_s, feat_in=80), st)
    decoder = conv1d_decoder(640, 65)
    decoder.load_state_dict({"weight": torch.from_numpy(dec["weight"]), "bias": torch.from_numpy(dec["bias"])})
    fbank = FilterbankFeatures(nfilt=80).eval()
    tt = BatchTextTransformer(tokens=synth.citrinet_vocab(64))
    for tag, lens in (("full", lens_full), ("ragged", np.array([12000, 6001], np.int64))):
        f, fl = fbank(torch.from_numpy(x), torch.from_numpy(lens))
        e, el = enc(f, fl)
        logits = decoder.eval()(e)
        ids = logits.argmax(1)
        e2e[f"cn.{tag}.logits"] = logits.numpy()
        e2e[f"cn.{tag}.out_lengths"] = el.numpy()
        e2e[f"cn.{tag}.ids"] = ids.numpy()
        e2e[f"cn.{tag}.text"] = np.array(tt.decode_prediction(ids))
    np.savez_compressed(os.path.join(out_dir, "e2e.npz"), **e2e)

    # ---------------------------------------------------------------- greedy decode
    dec_g = {}
    rng = np.random.Generator(np.random.PCG64(3000))
    tt = BatchTextTransformer(tokens=synth.quartznet_vocab())
    ids = rng.integers(0, 29, (6, 60)).astype(np.int64)
    ids[rng.random((6, 60)) < 0.5] = 28          # plenty of blanks
    ids[:, 1::2] = ids[:, ::2]                   # plenty of repeats
    ids[4] = 28                                  # all blank
    ids[5] = 3                                   # one long run
    dec_g["ids"] = ids
    dec_g["text"] = np.array(tt.decode_prediction(torch.from_numpy(ids)))
    tt2 = BatchTextTransformer(tokens=synth.citrinet_vocab(64))
    ids2 = rng.integers(0, 65, (3, 40)).astype(np.int64)
    dec_g["ids_bpe"] = ids2
    dec_g["text_bpe"] = np.array(tt2.decode_prediction(torch.from_numpy(ids2)))
    # argmax tie / NaN behaviour of torch.argmax on [B,V,T]
    lg = rng.standard_normal((2, 7, 9)).astype(np.float32)
    lg[0, 2, 3] = lg[0, 5, 3] = 9.0              # tie -> first index
    lg[1, 4, 1] = np.nan                         # NaN is maximal
    lg[1, 1, 6] = np.nan
    lg[1, 3, 6] = np.nan                         # two NaN -> first
    lg[0, :, 8] = 1.5                            # all equal -> 0
    dec_g["argmax_logits"] = lg
    dec_g["argmax_ids"] = torch.from_numpy(lg).argmax(1).numpy()
    np.savez_compressed(os.path.join(out_dir, "decode.npz"), **dec_g)

    total = sum(os.path.getsize(os.path.join(out_dir, f)) for f in os.listdir(out_dir))
    print("golden fixtures written to", out_dir, "total bytes", total)


if __name__ == "__main__":
    main()
