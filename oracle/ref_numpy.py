"""CPU oracle for the thunder-speech ASR forward hot path (TEST INFRASTRUCTURE ONLY).

This file is a from-scratch numpy restatement of what the reference computes on the path
``FilterbankFeatures -> Quartznet/Citrinet encoder -> conv1d decoder -> greedy CTC decode``.
It is the *checker* for the CUDA kernels: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product package
(``thunder_speech_b200``) never imports anything under ``oracle/``.

Parity status: PINNED.  Every function below is checked (tests/test_oracle_golden.py) against
outputs of the reference's own PyTorch modules run in the build container
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``), and against the exact known-answer cases of
the reference test-suite (``tests/text/test_transforms.py:59-91``, ``tests/test_blocks.py:56-68``,
``tests/quartznet/test_blocks_qn.py:89-143``).  The rows added around the path are pinned the same way, each by its own
generator importing the reference's code: audio ingest / resampling (``oracle/make_golden_ingest.py`` ->
``tests/golden/ingest.npz``, torchaudio 0.12's ``resample``), SpecAugment / SpecCutout under fixed seeds
(``oracle/make_golden_augment.py`` -> ``augment.npz``); the training-step oracle is ``oracle/ref_torch.py`` (autograd) pinned
by ``oracle/make_golden_train.py`` -> ``train.npz``.

The reference's arithmetic lives in third-party libraries that are not part of the reference tree:
PyTorch (pinned ``torch 1.12.0``, ``poetry.lock:1050``) for ``torch.stft``/``conv1d``/``batch_norm``
and torchaudio (pinned ``0.12.0``, ``poetry.lock:1061``) for ``melscale_fbanks``.  Their published
algorithms are restated here in numpy; internal accumulations run in float64 so the oracle is at
least as exact as the fp32 reference, results are returned as float32.

All citations are ``file:line`` relative to ``/root/reference``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np

F32 = np.float32
F64 = np.float64


# --------------------------------------------------------------------------------------------
# shared helpers  (src/thunder/blocks.py)
# --------------------------------------------------------------------------------------------
def lengths_to_mask(lengths: np.ndarray, max_length: int) -> np.ndarray:
    """``mask[b, t] = t < int(lengths[b])``  (src/thunder/blocks.py:156-170)."""
    lengths = np.asarray(lengths).astype(np.int64)  # .type(torch.long) truncates toward zero
    return np.arange(max_length, dtype=np.int64)[None, :] < lengths[:, None]


def get_same_padding(kernel_size: int, stride: int, dilation: int) -> int:
    """src/thunder/blocks.py:173-196."""
    if stride > 1 and dilation > 1:
        raise ValueError("Only stride OR dilation may be greater than 1")
    if dilation > 1:
        return (dilation * (kernel_size - 1) + 1) // 2
    return kernel_size // 2


def conv_out_len(lengths: np.ndarray, kernel_size: int, stride: int, padding: int, dilation: int) -> np.ndarray:
    """``floor((L + 2p - d(k-1) - 1) / s) + 1``  (src/thunder/quartznet/blocks.py:142-156)."""
    lengths = np.asarray(lengths)
    num = lengths + 2 * padding - dilation * (kernel_size - 1) - 1
    return np.floor_divide(num, stride) + 1


# --------------------------------------------------------------------------------------------
# feature front-end  (src/thunder/quartznet/transform.py)
# --------------------------------------------------------------------------------------------
def hann_window(win_length: int) -> np.ndarray:
    """``torch.hann_window(win_length, periodic=False)`` (transform.py:175)."""
    if win_length == 1:
        return np.ones(1, F32)
    n = np.arange(win_length, dtype=F64)
    return (0.5 - 0.5 * np.cos(2.0 * math.pi * n / (win_length - 1))).astype(F32)


def _hz_to_mel_slaney(freq: float) -> float:
    f_sp = 200.0 / 3
    mels = freq / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    if freq >= min_log_hz:
        mels = min_log_mel + math.log(freq / min_log_hz) / logstep
    return mels


def _mel_to_hz_slaney(mels: np.ndarray) -> np.ndarray:
    f_sp = 200.0 / 3
    freqs = f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    log_t = mels >= min_log_mel
    freqs = np.where(log_t, min_log_hz * np.exp(logstep * (mels - min_log_mel)), freqs)
    return freqs


def mel_filterbank(n_freqs: int, n_mels: int, sample_rate: int, f_min: float = 0.0,
                   f_max: Optional[float] = None) -> np.ndarray:
    """``torchaudio.functional.melscale_fbanks(norm="slaney", mel_scale="slaney")`` transposed to
    ``[n_mels, n_freqs]`` as used by ``MelScale`` (transform.py:227-240)."""
    if f_max is None:
        f_max = sample_rate / 2
    all_freqs = np.linspace(0, sample_rate // 2, n_freqs, dtype=F64)
    m_min = _hz_to_mel_slaney(f_min)
    m_max = _hz_to_mel_slaney(f_max)
    m_pts = np.linspace(m_min, m_max, n_mels + 2, dtype=F64)
    f_pts = _mel_to_hz_slaney(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]  # (n_freqs, n_mels + 2)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    enorm = 2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])
    fb = fb * enorm[None, :]
    return fb.T.astype(F32)  # [n_mels, n_freqs]


def preemphasis(x: np.ndarray, preemph: float = 0.97) -> np.ndarray:
    """``y[0]=x[0]; y[n]=x[n]-preemph*x[n-1]`` over the whole padded row (transform.py:136-144)."""
    x = np.asarray(x, F32)
    y = np.empty_like(x)
    y[:, 0] = x[:, 0]
    y[:, 1:] = x[:, 1:] - F32(preemph) * x[:, :-1]
    return y


def feature_lengths(lengths: np.ndarray, hop: int) -> np.ndarray:
    """``floor(len / hop) + 1`` as int64 (transform.py:182-184)."""
    lengths = np.asarray(lengths)
    return (np.floor(lengths.astype(F64) / hop) + 1).astype(np.int64)


def stft_power(y: np.ndarray, n_fft: int, hop: int, win_length: int, window: np.ndarray) -> np.ndarray:
    """``torch.stft(center=True, pad_mode="reflect", onesided)`` then ``sqrt(re^2+im^2)^2``
    (transform.py:194-207).  Returns ``[B, n_fft//2+1, 1 + N//hop]`` float32."""
    y = np.asarray(y, F64)
    B, N = y.shape
    pad = n_fft // 2
    yp = np.pad(y, ((0, 0), (pad, pad)), mode="reflect")
    n_frames = 1 + N // hop
    left = (n_fft - win_length) // 2
    wfull = np.zeros(n_fft, F64)
    wfull[left:left + win_length] = np.asarray(window, F64)
    idx = np.arange(n_frames)[:, None] * hop + np.arange(n_fft)[None, :]
    frames = yp[:, idx] * wfull[None, None, :]  # [B, F, n_fft]
    spec = np.fft.rfft(frames, n=n_fft, axis=-1)  # [B, F, n_fft//2+1]
    mag = np.sqrt(spec.real ** 2 + spec.imag ** 2)
    power = mag ** 2
    return np.ascontiguousarray(power.transpose(0, 2, 1)).astype(F32)


def mel_log(power: np.ndarray, fb: np.ndarray, log_scale: bool = True) -> np.ndarray:
    """``log(fb @ P + 2^-24)`` (transform.py:250-254). ``fb`` is ``[n_mels, n_freqs]``."""
    out = np.einsum("mf,bft->bmt", np.asarray(fb, F64), np.asarray(power, F64))
    if log_scale:
        out = np.log(out.astype(F32).astype(F64) + 2.0 ** -24)
    return out.astype(F32)


def normalize_features(x: np.ndarray, seq_len: np.ndarray, div_guard: float = 1e-5) -> np.ndarray:
    """Masked per-(batch, feature) normalisation over time with *biased* std and the guard added to
    the std, tail zeroed (transform.py:77-92 -> blocks.py:118-149)."""
    x = np.asarray(x, F64)
    mask = lengths_to_mask(seq_len, x.shape[-1])[:, None, :]  # [B,1,T]
    xm = np.where(mask, x, 0.0)
    n = mask.sum(axis=-1, keepdims=True).astype(F64)
    with np.errstate(divide="ignore", invalid="ignore"):
        mean = xm.sum(axis=-1, keepdims=True) / n
        num = ((xm - mean) ** 2).sum(axis=-1, keepdims=True)
        std = np.sqrt(num / n)
        out = (xm - mean) / (std + div_guard)
    return np.where(mask, out, 0.0).astype(F32)


def filterbank_features(audio: np.ndarray, lengths: np.ndarray, sample_rate: int = 16000,
                        n_window_size: int = 320, n_window_stride: int = 160, n_fft: int = 512,
                        preemph: float = 0.97, nfilt: int = 64,
                        window: Optional[np.ndarray] = None, fb: Optional[np.ndarray] = None,
                        return_intermediate: bool = False):
    """Eval-mode ``FilterbankFeatures`` (transform.py:258-321): dither is the identity in eval
    (transform.py:115-118).  Returns ``(features[B,nfilt,F] f32, feature_lengths[B] i64)``."""
    if n_window_size <= 0 or n_window_stride <= 0:
        raise ValueError("n_window_size and n_window_stride must be positive ints")
    if window is None:
        window = hann_window(n_window_size)
    if fb is None:
        fb = mel_filterbank(1 + n_fft // 2, nfilt, sample_rate, 0.0, sample_rate / 2)
    y = preemphasis(audio, preemph)
    power = stft_power(y, n_fft, n_window_stride, n_window_size, window)
    seq_len = feature_lengths(lengths, n_window_stride)
    logmel = mel_log(power, fb)
    feats = normalize_features(logmel, seq_len)
    if return_intermediate:
        return feats, seq_len, dict(preemph=y, power=power, logmel=logmel)
    return feats, seq_len


# --------------------------------------------------------------------------------------------
# encoder blocks  (src/thunder/quartznet/blocks.py, src/thunder/citrinet/blocks.py)
# --------------------------------------------------------------------------------------------
def conv1d(x: np.ndarray, w: np.ndarray, stride: int = 1, padding: int = 0, dilation: int = 1,
           groups: int = 1, bias: Optional[np.ndarray] = None) -> np.ndarray:
    """``torch.nn.functional.conv1d`` restated (cross-correlation, zero padding).
    ``x[B,Cin,T]``, ``w[Cout,Cin/groups,K]``; only ``groups in {1, Cin}`` are needed by the path."""
    x = np.asarray(x, F64)
    w = np.asarray(w, F64)
    B, Cin, T = x.shape
    Cout, Cg, K = w.shape
    T_out = (T + 2 * padding - dilation * (K - 1) - 1) // stride + 1
    if T_out <= 0:
        raise ValueError("conv1d: empty output")
    xp = np.pad(x, ((0, 0), (0, 0), (padding, padding)))
    out = np.zeros((B, Cout, T_out), F64)
    span = (T_out - 1) * stride + 1
    if groups == 1:
        assert Cg == Cin
        for k in range(K):
            xs = xp[:, :, k * dilation: k * dilation + span: stride]  # [B,Cin,T_out]
            out += np.einsum("oc,bct->bot", w[:, :, k], xs)
    elif groups == Cin and Cg == 1:
        mult = Cout // Cin
        assert mult == 1, "depthwise multiplier 1 only"
        for k in range(K):
            xs = xp[:, :, k * dilation: k * dilation + span: stride]
            out += w[None, :, 0, k, None] * xs
    else:
        raise NotImplementedError("groups must be 1 or Cin")
    if bias is not None:
        out += np.asarray(bias, F64)[None, :, None]
    return out.astype(F32)


def masked_conv1d(x, lengths, w, stride=1, padding=0, dilation=1, groups=1, bias=None, use_mask=True):
    """``MaskedConv1d.forward`` (quartznet/blocks.py:169-182): zero frames ``t >= len`` then conv."""
    x = np.asarray(x, F32)
    if use_mask:
        m = lengths_to_mask(lengths, x.shape[-1])[:, None, :]
        x = np.where(m, x, F32(0))
    out = conv1d(x, w, stride, padding, dilation, groups, bias)
    return out, conv_out_len(lengths, w.shape[-1], stride, padding, dilation)


def batchnorm_eval(x, weight, bias, running_mean, running_var, eps=1e-3):
    """Eval-mode ``BatchNorm1d(eps=1e-3)`` (quartznet/blocks.py:222)."""
    x = np.asarray(x, F64)
    inv = 1.0 / np.sqrt(np.asarray(running_var, F64) + eps)
    y = (x - np.asarray(running_mean, F64)[None, :, None]) * inv[None, :, None]
    y = y * np.asarray(weight, F64)[None, :, None] + np.asarray(bias, F64)[None, :, None]
    return y.astype(F32)


def squeeze_excite(x, w1, w2):
    """``SqueezeExcite.forward`` (citrinet/blocks.py:70-83): mean over ALL time steps (no mask),
    ``sigmoid(W2 relu(W1 mean))`` scale.  ``w1[C/r, C]``, ``w2[C, C/r]`` (nn.Linear layout)."""
    x64 = np.asarray(x, F64)
    y = x64.mean(axis=-1)  # [B,C]
    y = np.maximum(y @ np.asarray(w1, F64).T, 0.0)
    y = y @ np.asarray(w2, F64).T
    y = 1.0 / (1.0 + np.exp(-y))
    return (x64 * y[:, :, None]).astype(F32)


class BlockCfg:
    """Static description of one Quartznet/Citrinet block (constructor args of
    ``QuartznetBlock`` quartznet/blocks.py:232-243 / ``CitrinetBlock`` citrinet/blocks.py:87-98)."""

    def __init__(self, in_channels, out_channels, repeat=5, kernel_size=11, stride=1, dilation=1,
                 residual=True, separable=False, kind="quartznet"):
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.repeat = repeat
        self.kernel_size = kernel_size
        self.stride = stride
        self.dilation = dilation
        self.residual = residual
        self.separable = separable
        self.kind = kind  # "quartznet" | "citrinet"

    def sub_strides(self) -> List[int]:
        """Stride of each of the ``repeat`` sub-blocks: QuartzNet strides every sub-block
        (quartznet/blocks.py:266-296); Citrinet only the last (citrinet/blocks.py:128,140-152)."""
        if self.kind == "quartznet":
            return [self.stride] * self.repeat
        return [1] * (self.repeat - 1) + [self.stride]

    def residual_stride(self) -> int:
        """quartznet/blocks.py:301 (``stride**repeat``) vs citrinet/blocks.py:159 (``stride``)."""
        if self.stride == 1:
            return 1
        return self.stride ** self.repeat if self.kind == "quartznet" else self.stride

    def mconv_index(self, r: int) -> int:
        """Index inside ``mconv`` of the first conv of sub-block ``r``: each sub-block contributes
        ``[dw, pw, BN]`` (separable) or ``[conv, BN]`` followed by ``[ReLU, Dropout]`` except after
        the last one (quartznet/blocks.py:185-228,266-296)."""
        per = (3 if self.separable else 2) + 2
        return r * per


def _bn(state: Dict[str, np.ndarray], prefix: str, x):
    return batchnorm_eval(x, state[prefix + ".weight"], state[prefix + ".bias"],
                          state[prefix + ".running_mean"], state[prefix + ".running_var"])


def block_forward(x, lengths, cfg: BlockCfg, state: Dict[str, np.ndarray], prefix: str = ""):
    """``QuartznetBlock.forward`` (quartznet/blocks.py:317-338) / ``CitrinetBlock.forward``
    (citrinet/blocks.py:177-197) in eval mode, parameters addressed by the reference's own
    ``state_dict`` key names."""
    x = np.asarray(x, F32)
    lengths = np.asarray(lengths)
    out, out_len = x, lengths
    strides = cfg.sub_strides()
    for r in range(cfg.repeat):
        i = cfg.mconv_index(r)
        s = strides[r]
        pad = get_same_padding(cfg.kernel_size, s, cfg.dilation)
        if cfg.separable:
            cin = out.shape[1]
            out, out_len = masked_conv1d(out, out_len, state[f"{prefix}mconv.{i}.conv.weight"], s, pad,
                                         cfg.dilation, groups=cin)
            out, out_len = masked_conv1d(out, out_len, state[f"{prefix}mconv.{i + 1}.conv.weight"])
            out = _bn(state, f"{prefix}mconv.{i + 2}.layer.0", out)
        else:
            out, out_len = masked_conv1d(out, out_len, state[f"{prefix}mconv.{i}.conv.weight"], s, pad,
                                         cfg.dilation)
            out = _bn(state, f"{prefix}mconv.{i + 1}.layer.0", out)
        if r != cfg.repeat - 1:
            out = np.maximum(out, F32(0))
    if cfg.kind == "citrinet":
        i_se = cfg.mconv_index(cfg.repeat - 1) + (3 if cfg.separable else 2)
        out = squeeze_excite(out, state[f"{prefix}mconv.{i_se}.layer.0.fc.0.weight"],
                             state[f"{prefix}mconv.{i_se}.layer.0.fc.2.weight"])
    if cfg.residual:
        res, _ = masked_conv1d(x, lengths, state[f"{prefix}res.0.conv.weight"], cfg.residual_stride(), 0, 1)
        res = _bn(state, f"{prefix}res.1.layer.0", res)
        out = out + res
    out = np.maximum(out, F32(0))
    return out, out_len


def quartznet_cfgs(feat_in: int = 64, filters: Sequence[int] = (256, 256, 512, 512, 512),
                   kernel_sizes: Sequence[int] = (33, 39, 51, 63, 75), repeat_blocks: int = 1) -> List[BlockCfg]:
    """``QuartznetEncoder`` = stem + body (quartznet/blocks.py:341-434)."""
    cfgs = [BlockCfg(feat_in, 256, repeat=1, stride=2, kernel_size=33, residual=False, separable=True)]
    f_in = 256
    for f, k in zip(filters, kernel_sizes):
        for _ in range(repeat_blocks):
            cfgs.append(BlockCfg(f_in, f, kernel_size=k, separable=True))
            f_in = f
    cfgs.append(BlockCfg(f_in, 512, repeat=1, dilation=2, kernel_size=87, residual=False, separable=True))
    cfgs.append(BlockCfg(512, 1024, repeat=1, kernel_size=1, residual=False, separable=False))
    return cfgs


def citrinet_cfgs(filters: Sequence[int], kernel_sizes: Sequence[int], strides: Sequence[int],
                  feat_in: int = 80) -> List[BlockCfg]:
    """``CitrinetEncoder`` = stem + body (citrinet/blocks.py:200-278); stem is hard-coded to 256."""
    cfgs = [BlockCfg(feat_in, 256, repeat=1, kernel_size=5, residual=False, separable=True, kind="citrinet")]
    f_in = 256
    for f, k, s in zip(filters, kernel_sizes, strides):
        cfgs.append(BlockCfg(f_in, f, kernel_size=k, stride=s, separable=True, kind="citrinet"))
        f_in = f
    cfgs.append(BlockCfg(f_in, 640, repeat=1, kernel_size=41, residual=False, separable=True, kind="citrinet"))
    return cfgs


def encoder_forward(x, lengths, cfgs: List[BlockCfg], state: Dict[str, np.ndarray], prefix: str = ""):
    """``MultiSequential(stem, *body)`` (blocks.py:94-102)."""
    for bi, cfg in enumerate(cfgs):
        x, lengths = block_forward(x, lengths, cfg, state, prefix=f"{prefix}{bi}.")
    return x, lengths


def decoder_forward(x, weight, bias):
    """``conv1d_decoder``: ``Conv1d(C, V, 1, bias=True)`` (blocks.py:199-216)."""
    return conv1d(x, weight, bias=bias)


# --------------------------------------------------------------------------------------------
# greedy CTC decode  (module.py:100, text_processing/transform.py:93-122, vocab.py:85-130)
# --------------------------------------------------------------------------------------------
def greedy_argmax(logits: np.ndarray) -> np.ndarray:
    """``pred.argmax(1)`` on ``[B,V,T]``: first maximal index wins, NaN counts as maximal
    (SURVEY.md K4b probe of ``torch.argmax``)."""
    logits = np.asarray(logits)
    B, V, T = logits.shape
    out = np.zeros((B, T), np.int64)
    for b in range(B):
        col = logits[b]
        nan = np.isnan(col)
        has_nan = nan.any(axis=0)
        first_nan = nan.argmax(axis=0)
        with np.errstate(invalid="ignore"):
            am = np.where(nan, -np.inf, col).argmax(axis=0)
        out[b] = np.where(has_nan, first_nan, am)
    return out


def ctc_collapse(ids: np.ndarray) -> List[np.ndarray]:
    """Per-row ``torch.unique_consecutive`` (text_processing/transform.py:107-110).  Blanks are kept
    (the reference removes the blank *string* after joining, vocab.py:114-130)."""
    ids = np.asarray(ids)
    rows = []
    for row in ids:
        if row.size == 0:
            rows.append(row.copy())
            continue
        keep = np.ones(row.shape[0], bool)
        keep[1:] = row[1:] != row[:-1]
        rows.append(row[keep])
    return rows


class Vocab:
    """``Vocabulary`` token table (text_processing/vocab.py:18-67): special tokens are appended in
    the order blank, pad, unknown, start, end, skipping ``None`` and ones already present."""

    def __init__(self, tokens: List[str], blank_token="<blank>", pad_token=None, unknown_token=None,
                 start_token=None, end_token=None):
        self.blank_token = blank_token
        self.pad_token = pad_token or blank_token
        self.start_token = start_token
        self.end_token = end_token
        itos = list(tokens)
        for tok in (blank_token, pad_token, unknown_token, start_token, end_token):
            if tok and tok not in itos:
                itos.append(tok)
        self.itos = itos
        self.blank_idx = itos.index(self.blank_token)

    def remove_special_tokens(self, text: str) -> str:
        """vocab.py:114-130 (string replace, in this order)."""
        text = text.replace(self.blank_token, "")
        text = text.replace(self.pad_token, "")
        if self.start_token is not None:
            text = text.replace(self.start_token, "")
        if self.end_token is not None:
            text = text.replace(self.end_token, "")
        return text


def decode_prediction(ids: np.ndarray, vocab: Vocab, remove_repeated: bool = True) -> List[str]:
    """``BatchTextTransformer.decode_prediction`` (text_processing/transform.py:93-122)."""
    out = []
    rows = ctc_collapse(ids) if remove_repeated else list(np.asarray(ids))
    for row in rows:
        text = "".join(vocab.itos[int(i)] for i in row)
        text = text.replace("▁", " ").replace("|", " ")
        out.append(vocab.remove_special_tokens(text))
    return out


def predict(audio: np.ndarray, cfgs: List[BlockCfg], enc_state, dec_w, dec_b, vocab: Vocab,
            feat_kwargs: Optional[dict] = None, return_all: bool = False):
    """``BaseCTCModule.predict`` (module.py:88-100): all lengths = N, decode every frame."""
    audio = np.asarray(audio, F32)
    B, N = audio.shape
    lengths = np.full((B,), N, np.int64)
    feats, flen = filterbank_features(audio, lengths, **(feat_kwargs or {}))
    enc, elen = encoder_forward(feats, flen, cfgs, enc_state)
    logits = decoder_forward(enc, dec_w, dec_b)
    ids = greedy_argmax(logits)
    texts = decode_prediction(ids, vocab)
    if return_all:
        return texts, dict(features=feats, feature_lengths=flen, encoded=enc, out_lengths=elen,
                           logits=logits, ids=ids)
    return texts


# ------------------------------------------------------------------------------------------------ audio ingest
# AudioFileLoader.preprocess_audio (src/thunder/data/dataset.py:50-77): mono mix, DC removal,
# resample.  The resampler is torchaudio.functional.resample (torchaudio ^0.12, locked 0.12.0; not under /root/reference):
# windowed-sinc polyphase FIR, sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99 -- restated from its published
# algorithm (torchaudio/functional/functional.py: _get_sinc_resample_kernel / _apply_sinc_resample_kernel) and pinned
# against tests/golden/ingest.npz, which the reference's own module produced.

def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """(kernel [new', taps] float32, width, orig', new') with orig' = orig/gcd, new' = new/gcd, taps = 2*width + orig'."""
    import math

    g = math.gcd(int(orig_freq), int(new_freq))
    o, n = int(orig_freq) // g, int(new_freq) // g
    base = min(o, n) * rolloff
    width = math.ceil(lowpass_filter_width * o / base)
    idx = np.arange(-width, width + o, dtype=np.float64)[None, :] / o
    t = (np.arange(0, -n, -1, dtype=np.float64)[:, None] / n + idx) * base
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    k = k * window * (base / o)
    return k.astype(np.float32), width, o, n


def resample(x: np.ndarray, orig_freq: int, new_freq: int) -> np.ndarray:
    """x [..., length] -> [..., ceil(new' * length / orig')]; float64 accumulation of the float32 taps."""
    if int(orig_freq) == int(new_freq):
        return np.asarray(x, np.float32)
    k, width, o, n = sinc_resample_kernel(orig_freq, new_freq)
    x = np.asarray(x, np.float32)
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1]).astype(np.float64)
    length = x2.shape[1]
    xp = np.pad(x2, ((0, 0), (width, width + o)))
    nfr = (xp.shape[1] - k.shape[1]) // o + 1
    frames = np.lib.stride_tricks.sliding_window_view(xp, k.shape[1], axis=1)[:, ::o][:, :nfr]      # [R, nfr, taps]
    y = np.einsum("rft,pt->rfp", frames, k.astype(np.float64)).reshape(x2.shape[0], -1)              # interleave phases
    target = -((-n * length) // o)                                                                  # ceil(n*length/o)
    return y[:, :target].astype(np.float32).reshape(lead + (target,))


def preprocess_audio(audio: np.ndarray, sample_rate: int, force_mono: bool = True, target_rate: int = 16000) -> np.ndarray:
    """audio [channels, time] float -> [1 (or channels), time'] (dataset.py:50-77)."""
    a = np.asarray(audio, np.float32).astype(np.float64)
    if force_mono and a.shape[0] > 1:
        a = a.mean(0, keepdims=True)
    if a.shape[0] != 1:
        raise RuntimeError("audio - audio.mean(1) only broadcasts for one channel (dataset.py:69)")
    a = (a - a.mean(1)).astype(np.float32)
    return resample(a, sample_rate, target_rate) if int(sample_rate) != int(target_rate) else a


# ------------------------------------------------------------------------------------------------ SpecAugment / SpecCutout
# src/thunder/quartznet/spec_augment.py:23-110 (train() mode).  The mask intervals come from torch.rand(1) on the host
# generator (torchaudio.functional.mask_along_axis / _create_mask): v = rand * param, v0 = rand * (size - v), masked
# [long(v0), long(v0) + long(v)), the same interval for every utterance.  `draw()` must return those uniform numbers in order
# (tests pass `lambda: float(torch.rand(1))` after torch.manual_seed to reproduce the reference's stream).

def _mask_interval(size: int, param: int, draw):
    value = np.float32(draw()) * np.float32(param)
    min_value = np.float32(draw()) * (np.float32(size) - value)
    start = int(min_value)
    return start, start + int(value)


def spec_augment(x: np.ndarray, draw, time_masks=0, freq_masks=0, time_width=10, freq_width=10) -> np.ndarray:
    """x [B, C, T]: time masks first, then frequency masks; a mask with param < 1 is skipped (mask_along_axis)."""
    y = np.array(x, copy=True)
    for _ in range(time_masks):
        if time_width >= 1:
            a, b = _mask_interval(y.shape[2], time_width, draw)
            y[:, :, a:b] = 0
    for _ in range(freq_masks):
        if freq_width >= 1:
            a, b = _mask_interval(y.shape[1], freq_width, draw)
            y[:, a:b, :] = 0
    return y


def spec_cutout(x: np.ndarray, draw, rect_masks=0, time_width=5, freq_width=20) -> np.ndarray:
    """Rectangles; like the reference, BOTH extents are drawn with freq_width (spec_augment.py:106-107)."""
    y = np.array(x, copy=True)
    for _ in range(rect_masks):
        f0, f1 = _mask_interval(y.shape[1], freq_width, draw)
        t0, t1 = _mask_interval(y.shape[2], freq_width, draw)
        y[:, f0:f1, t0:t1] = 0
    return y
