"""CPU port of the reference path built from the SAME library calls the reference makes
(``torch.stft``, ``F.conv1d``, ``F.batch_norm`` ... on CPU, all host threads).  TEST/BENCH
INFRASTRUCTURE ONLY -- used by ``bench.py`` for the ``cpu_baseline`` leg and ``--impl reference``
(the reference is pure Python over PyTorch and cannot travel to the GPU box), and by
tests/test_oracle_golden.py as a second checker.  Never imported by the product package.

Parity status: PINNED against tests/golden/*.npz (tests/test_oracle_golden.py::test_torch_port_*).
Citations are file:line under /root/reference.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import ref_numpy as R


def features(audio: torch.Tensor, lengths: torch.Tensor, nfilt: int = 64, n_fft: int = 512, hop: int = 160,
             win: int = 320, preemph: float = 0.97, fb: torch.Tensor = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """src/thunder/quartznet/transform.py:136-144,186-208,243-255,77-92 (eval mode)."""
    x = torch.cat((audio[:, :1], audio[:, 1:] - preemph * audio[:, :-1]), dim=1)
    window = torch.hann_window(win, periodic=False)
    spec = torch.stft(x, n_fft=n_fft, hop_length=hop, win_length=win, center=True, window=window,
                      return_complex=True)
    spec = torch.view_as_real(spec)
    p = torch.sqrt(spec.pow(2).sum(-1)).pow(2.0)
    if fb is None:
        fb = torch.from_numpy(R.mel_filterbank(1 + n_fft // 2, nfilt, 16000))
    mel = torch.log(torch.matmul(fb.unsqueeze(0), p) + 2 ** -24)
    seq = (torch.floor(lengths / hop) + 1).to(torch.long)
    mask = (torch.arange(mel.shape[-1]).expand(mel.shape[0], -1) < seq.unsqueeze(1)).unsqueeze(1)
    xm = mel.masked_fill(~mask, 0.0)
    n = mask.sum(-1, keepdim=True)
    mean = xm.sum(-1, keepdim=True) / n
    std = ((xm - mean).pow(2).sum(-1, keepdim=True) / n).sqrt()
    return ((xm - mean) / (std + 1e-5)).masked_fill(~mask, 0.0), seq


def _mask(x, lengths):
    m = torch.arange(x.shape[-1]).expand(x.shape[0], -1) < lengths.to(torch.long).unsqueeze(1)
    return x.masked_fill(~m.unsqueeze(1), 0)


def _bn(st, p, x, train=False):
    """eval: running statistics.  train: batch statistics over ALL (batch, time) positions incl. padded frames, and the
    running statistics in `st` are updated in place with momentum 0.1 (nn.BatchNorm1d(eps=1e-3, momentum=0.1))."""
    return F.batch_norm(x, st[p + ".running_mean"], st[p + ".running_var"], st[p + ".weight"], st[p + ".bias"],
                        train, 0.1, 1e-3)


def bf16_store(x: torch.Tensor) -> torch.Tensor:
    """Round to bf16 and back with a straight-through gradient: models a tensor that the device path STORES in bf16.
    With `block(..., store=bf16_store)` the oracle rounds at the points where the B200 path writes bf16 rows (block
    input, depthwise output, GEMM output, BN+ReLU output) and where it reads bf16 weights, so that ReLU masks and batch
    statistics agree with the device and what is left is backward-pass rounding only."""
    return x + (x.detach().to(torch.bfloat16).to(x.dtype) - x.detach())


def block(x, lengths, cfg: R.BlockCfg, st: Dict[str, torch.Tensor], prefix: str, train: bool = False, store=None):
    """quartznet/blocks.py:317-338 / citrinet/blocks.py:177-197; differentiable (torch autograd) when the tensors in
    `st` require grad -- the oracle of the training step (module.py:102-127).  `store` (default: identity) is applied
    to every tensor the device path materialises, see `bf16_store`."""
    q = store if store is not None else (lambda t: t)
    if store is not None:
        st = {k: (q(v) if k.endswith("conv.weight") else v) for k, v in st.items()}
    x = q(x)
    out, out_len = x, lengths
    strides = cfg.sub_strides()
    for r in range(cfg.repeat):
        i = cfg.mconv_index(r)
        s = strides[r]
        pad = R.get_same_padding(cfg.kernel_size, s, cfg.dilation)
        k = cfg.kernel_size
        if cfg.separable:
            out = F.conv1d(_mask(out, out_len), st[f"{prefix}mconv.{i}.conv.weight"], None, s, pad, cfg.dilation,
                           groups=out.shape[1])
            out_len = torch.div(out_len + 2 * pad - cfg.dilation * (k - 1) - 1, s, rounding_mode="floor") + 1
            out = q(F.conv1d(_mask(q(out), out_len), st[f"{prefix}mconv.{i + 1}.conv.weight"]))
            out = _bn(st, f"{prefix}mconv.{i + 2}.layer.0", out, train)
        else:
            out = F.conv1d(_mask(out, out_len), st[f"{prefix}mconv.{i}.conv.weight"], None, s, pad, cfg.dilation)
            out_len = torch.div(out_len + 2 * pad - cfg.dilation * (k - 1) - 1, s, rounding_mode="floor") + 1
            out = _bn(st, f"{prefix}mconv.{i + 1}.layer.0", q(out), train)
        if r != cfg.repeat - 1:
            out = q(F.relu(out))
    if cfg.kind == "citrinet":
        i_se = cfg.mconv_index(cfg.repeat - 1) + (3 if cfg.separable else 2)
        y = out.mean(-1)
        y = F.relu(y @ st[f"{prefix}mconv.{i_se}.layer.0.fc.0.weight"].T) @ st[f"{prefix}mconv.{i_se}.layer.0.fc.2.weight"].T
        out = out * torch.sigmoid(y).unsqueeze(-1)
    if cfg.residual:
        res = q(F.conv1d(_mask(x, lengths), st[f"{prefix}res.0.conv.weight"], None, cfg.residual_stride()))
        out = out + _bn(st, f"{prefix}res.1.layer.0", res, train)
    return q(F.relu(out)), out_len


def encoder(x, lengths, cfgs: List[R.BlockCfg], st, train: bool = False, store=None):
    for bi, cfg in enumerate(cfgs):
        x, lengths = block(x, lengths, cfg, st, f"{bi}.", train, store)
    return x, lengths


def _fold_bn(st, p):
    """eval BatchNorm1d(eps=1e-3) as a per-channel affine (quartznet/blocks.py:222-228)."""
    s = st[p + ".weight"] / torch.sqrt(st[p + ".running_var"] + 1e-3)
    return s, st[p + ".bias"] - st[p + ".running_mean"] * s


def rounder(fmt):
    """Storage-format model: ``None``/"f32" = identity, "bf16"/"fp16" = round to that format and back."""
    if fmt in (None, "f32"):
        return lambda t: t
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[fmt]
    return lambda t: t.to(dt).to(torch.float32)


def block_storage(x, lengths, cfg: R.BlockCfg, st, prefix: str, fmt="bf16"):
    """Eval-mode block (quartznet/blocks.py:317-338 / citrinet/blocks.py:177-197) with a 16-bit STORAGE format simulated at
    exactly the four places per sub-block where ANY 16-bit tensor-core implementation must round: the depthwise taps, the
    depthwise output (GEMM operand), the BN-folded pointwise / residual weights (GEMM operand) and the sub-block output
    (and, for Citrinet, the pre-gate tensor the SqueezeExcite scale is applied to).  Everything between two roundings is
    fp32, like the accumulators of the device path.  With ``fmt=None`` this is algebraically the reference block (BN
    folded).  It is the yardstick that separates what the FORMAT costs at depth from what the kernels add."""
    q = rounder(fmt)
    out, out_len = x, lengths
    strides = cfg.sub_strides()
    for r in range(cfg.repeat):
        i = cfg.mconv_index(r)
        s, k = strides[r], cfg.kernel_size
        pad = R.get_same_padding(k, s, cfg.dilation)
        if cfg.separable:
            out = F.conv1d(_mask(out, out_len), q(st[f"{prefix}mconv.{i}.conv.weight"]), None, s, pad, cfg.dilation,
                           groups=out.shape[1])
            out_len = torch.div(out_len + 2 * pad - cfg.dilation * (k - 1) - 1, s, rounding_mode="floor") + 1
            sc, sh = _fold_bn(st, f"{prefix}mconv.{i + 2}.layer.0")
            w = q(st[f"{prefix}mconv.{i + 1}.conv.weight"] * sc[:, None, None])
            out = F.conv1d(_mask(q(out), out_len), w) + sh[None, :, None]
        else:
            sc, sh = _fold_bn(st, f"{prefix}mconv.{i + 1}.layer.0")
            w = q(st[f"{prefix}mconv.{i}.conv.weight"] * sc[:, None, None])
            out = F.conv1d(_mask(out, out_len), w, None, s, pad, cfg.dilation) + sh[None, :, None]
            out_len = torch.div(out_len + 2 * pad - cfg.dilation * (k - 1) - 1, s, rounding_mode="floor") + 1
        if r != cfg.repeat - 1:
            out = q(F.relu(out))
    if cfg.kind == "citrinet":
        i_se = cfg.mconv_index(cfg.repeat - 1) + (3 if cfg.separable else 2)
        y = out.mean(-1)
        y = F.relu(y @ st[f"{prefix}mconv.{i_se}.layer.0.fc.0.weight"].T) @ st[f"{prefix}mconv.{i_se}.layer.0.fc.2.weight"].T
        out = q(out) * torch.sigmoid(y).unsqueeze(-1)
    if cfg.residual:
        sc, sh = _fold_bn(st, f"{prefix}res.1.layer.0")
        w = q(st[f"{prefix}res.0.conv.weight"] * sc[:, None, None])
        out = out + F.conv1d(_mask(x, lengths), w, None, cfg.residual_stride()) + sh[None, :, None]
    return q(F.relu(out)), out_len


def ctc_loss(logits: torch.Tensor, y: torch.Tensor, prob_lengths: torch.Tensor, y_lengths: torch.Tensor,
             blank_idx: int) -> torch.Tensor:
    """calculate_ctc (src/thunder/ctc_loss.py:15-47): permute -> log_softmax -> F.ctc_loss(mean, zero_infinity)."""
    logprobs = F.log_softmax(logits.permute(2, 0, 1), dim=2)
    return F.ctc_loss(logprobs, y, prob_lengths.long(), y_lengths, blank=blank_idx, reduction="mean",
                      zero_infinity=True)


def predict_ids(audio: torch.Tensor, cfgs, st, dec_w, dec_b, nfilt: int):
    """module.py:88-100 up to the argmax; the string part is oracle.ref_numpy.decode_prediction."""
    lengths = torch.full((audio.shape[0],), audio.shape[1], dtype=torch.long)
    f, fl = features(audio, lengths, nfilt=nfilt)
    e, el = encoder(f, fl, cfgs, st)
    logits = F.conv1d(e, dec_w, dec_b)
    return logits.argmax(1), logits, el


def to_torch(state: Dict[str, np.ndarray]) -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(np.asarray(v)) for k, v in state.items()}


def make_workload(name: str, batch: int, samples: int, nfilt: int):
    """Closure running one step of a bench workload on CPU + a description of the bounded sample."""
    from thunder_speech_b200 import synth

    x = torch.from_numpy(synth.audio(batch, samples, 1234, "noise"))
    lens = torch.full((batch,), samples, dtype=torch.long)
    if name == "features":
        fb = torch.from_numpy(R.mel_filterbank(257, nfilt, 16000))

        def feat_step():
            with torch.no_grad():
                return features(x, lens, nfilt=nfilt, fb=fb)

        return feat_step, f"B={batch} x {samples / 16000:.0f} s, torch CPU fp32"
    if name == "quartznet15x5_train":
        # BaseCTCModule.training_step + AdamW (src/thunder/module.py:102-127, :32): features under no_grad, train()-mode
        # encoder, decoder, calculate_ctc, backward, optimiser step -- all torch CPU fp32 autograd
        cfgs = R.quartznet_cfgs(repeat_blocks=3)
        st = to_torch(synth.encoder_state(synth.quartznet_block_list(repeat_blocks=3), seed=0))
        dec = to_torch(synth.decoder_state(1024, 29, seed=1))
        params = []
        for k, v in list(st.items()) + list(dec.items()):
            if v.dtype.is_floating_point and "running" not in k:
                params.append(v.requires_grad_(True))
        opt = torch.optim.AdamW(params, lr=1e-4)
        rng = np.random.Generator(np.random.PCG64(99))
        L = 120
        y = torch.from_numpy(rng.integers(0, 28, (batch, L)).astype(np.int64))
        ylen = torch.from_numpy(rng.integers(L // 2, L + 1, batch).astype(np.int64))
        rl = torch.from_numpy(synth.ragged_lengths(batch, samples, 7))

        def train_step():
            opt.zero_grad()
            with torch.no_grad():
                f, fl = features(x, rl, nfilt=nfilt)
            e, el = encoder(f, fl, cfgs, st, train=True)
            loss = ctc_loss(F.conv1d(e, dec["weight"], dec["bias"]), y, el, ylen, 28)
            loss.backward()
            opt.step()
            return float(loss.detach())

        return train_step, f"B={batch} x {samples / 16000:.0f} s training step (fwd+bwd+CTC+AdamW), torch CPU fp32 autograd"
    if name == "quartznet15x5":
        cfgs = R.quartznet_cfgs(repeat_blocks=3)
        st = to_torch(synth.encoder_state(synth.quartznet_block_list(repeat_blocks=3), seed=0))
        dec = to_torch(synth.decoder_state(1024, 29, seed=1))
        vocab = R.Vocab(synth.quartznet_vocab())
    elif name == "citrinet1024":
        c = synth.CITRINET_1024
        cfgs = R.citrinet_cfgs(c["filters"], c["kernel_sizes"], c["strides"], feat_in=80)
        st = to_torch(synth.encoder_state(synth.citrinet_block_list(c["filters"], c["kernel_sizes"], c["strides"], 80),
                                          seed=0, se=True))
        dec = to_torch(synth.decoder_state(640, 1025, seed=1))
        vocab = R.Vocab(synth.citrinet_vocab(1024))
    else:
        raise ValueError(name)

    def step():
        with torch.no_grad():
            ids, _, _ = predict_ids(x, cfgs, st, dec["weight"], dec["bias"], nfilt)
        return R.decode_prediction(ids.numpy(), vocab)

    return step, f"B={batch} x {samples / 16000:.0f} s predict() incl. greedy decode, torch CPU fp32"
