"""Training-step goldens (tests/golden/train.npz) from the REFERENCE's own modules in train() mode with torch
autograd on CPU fp32: block outputs, input/parameter gradients, updated BatchNorm running statistics, CTC loss.
Run in the build container only (`python oracle/make_golden_train.py`); see make_golden.py for the import stub."""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import import_reference  # noqa: E402

TRAIN_BLOCK_CASES = [
    # name, cfg, B, T
    ("res", dict(in_channels=16, out_channels=24, repeat=3, kernel_size=5, stride=1, dilation=1, residual=True,
                 separable=True), 3, 50),
    ("stem", dict(in_channels=8, out_channels=16, repeat=1, kernel_size=33, stride=2, dilation=1, residual=False,
                  separable=True), 2, 101),
    ("dil", dict(in_channels=16, out_channels=16, repeat=1, kernel_size=9, stride=1, dilation=2, residual=False,
                 separable=True), 2, 70),
    ("k1", dict(in_channels=16, out_channels=32, repeat=1, kernel_size=1, stride=1, dilation=1, residual=False,
                separable=False), 2, 37),
]


def block_case(ci):
    from thunder_speech_b200 import synth

    name, cfg, B, T = TRAIN_BLOCK_CASES[ci]
    rng = np.random.Generator(np.random.PCG64(5000 + ci))
    st = synth.block_state(rng, "", cfg["in_channels"], cfg["out_channels"], cfg["repeat"], cfg["kernel_size"],
                           cfg["residual"], cfg["separable"])
    x = rng.standard_normal((B, cfg["in_channels"], T)).astype(np.float32)
    lens = np.sort(rng.integers(T // 2, T + 1, B))[::-1].astype(np.int64).copy()
    lens[0] = T
    T_out = T
    for _ in range(cfg["repeat"]):
        p = cfg["kernel_size"] // 2 if cfg["dilation"] == 1 else (cfg["dilation"] * (cfg["kernel_size"] - 1) + 1) // 2
        T_out = (T_out + 2 * p - cfg["dilation"] * (cfg["kernel_size"] - 1) - 1) // cfg["stride"] + 1
    R = rng.standard_normal((B, cfg["out_channels"], T_out)).astype(np.float32)
    return name, cfg, st, x, lens, R


def tiny_model_case():
    """QuartzNet-shaped tiny model (stem / 5 body blocks / dilated block / 1x1 block hard-coded like the reference),
    2 utterances, CTC targets."""
    from thunder_speech_b200 import synth

    filters, kernels = [32, 32, 32, 32, 32], [5, 7, 9, 11, 13]
    blocks = synth.quartznet_block_list(filters=filters, kernel_sizes=kernels, repeat_blocks=1)
    st = synth.encoder_state(blocks, seed=77)
    dec = synth.decoder_state(1024, 29, seed=78)
    x = synth.audio(2, 9600, 79, "tones")
    lens = np.array([9600, 7000], np.int64)
    y = np.array([[1, 5, 9, 3, 0, 7], [2, 2, 8, 0, 0, 0]], np.int64)
    y_len = np.array([6, 3], np.int64)
    return filters, kernels, st, dec, x, lens, y, y_len


def main():
    warnings.filterwarnings("ignore")
    import_reference()
    from thunder.blocks import conv1d_decoder
    from thunder.ctc_loss import calculate_ctc
    from thunder.quartznet.blocks import QuartznetBlock, QuartznetEncoder
    from thunder.quartznet.transform import FilterbankFeatures

    out = {}
    for ci in range(len(TRAIN_BLOCK_CASES)):
        name, cfg, st, x, lens, R = block_case(ci)
        mod = QuartznetBlock(cfg["in_channels"], cfg["out_channels"], repeat=cfg["repeat"],
                             kernel_size=(cfg["kernel_size"],), stride=(cfg["stride"],), dilation=(cfg["dilation"],),
                             residual=cfg["residual"], separable=cfg["separable"])
        mod.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}, strict=True)
        mod.train()
        xt = torch.from_numpy(x).requires_grad_(True)
        y, yl = mod(xt, torch.from_numpy(lens))
        loss = (y * torch.from_numpy(R)).sum()
        loss.backward()
        out[f"{name}.out"] = y.detach().numpy()
        out[f"{name}.out_lengths"] = yl.numpy()
        out[f"{name}.dx"] = xt.grad.numpy()
        for k, p in mod.named_parameters():
            out[f"{name}.grad.{k}"] = p.grad.numpy()
        for k, b in mod.named_buffers():
            if "running" in k:
                out[f"{name}.buf.{k}"] = b.numpy().copy()

    filters, kernels, st, dec, x, lens, y, y_len = tiny_model_case()
    enc = QuartznetEncoder(filters=filters, kernel_sizes=kernels, repeat_blocks=1)
    enc.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}, strict=True)
    decoder = conv1d_decoder(1024, 29)
    decoder.load_state_dict({k: torch.from_numpy(v) for k, v in dec.items()})
    fb = FilterbankFeatures(nfilt=64, dither=0.0)
    enc.train(); decoder.train(); fb.train()
    f, fl = fb(torch.from_numpy(x), torch.from_numpy(lens))
    e, el = enc(f, fl)
    logits = decoder(e)
    loss = calculate_ctc(logits, torch.from_numpy(y), el, torch.from_numpy(y_len), 28)
    loss.backward()
    out["model.loss"] = np.array(loss.item(), np.float64)
    out["model.out_lengths"] = el.numpy()
    out["model.logits"] = logits.detach().numpy()
    rng = np.random.Generator(np.random.PCG64(9999))
    names = []
    for k, p in list(enc.named_parameters()) + [("decoder." + k, p) for k, p in decoder.named_parameters()]:
        g = p.grad.numpy()
        r = rng.standard_normal(g.shape).astype(np.float32)
        out[f"model.gproj.{k}"] = np.array([float((g * r).sum()), float(np.sqrt((g * g).sum()))], np.float64)
        names.append(k)
        if g.size <= 2048:
            out[f"model.grad.{k}"] = g
    out["model.param_names"] = np.array(names)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "train.npz"), **out)
    print("train goldens:", len(out), "arrays,", os.path.getsize(os.path.join(ROOT, "tests", "golden", "train.npz")), "bytes; loss", loss.item())


if __name__ == "__main__":
    main()
