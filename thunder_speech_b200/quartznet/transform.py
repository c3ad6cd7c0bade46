"""``FilterbankFeatures`` drop-in (mirrors ``src/thunder/quartznet/transform.py``).

Same constructor, same ``(audio[B,N], lengths[B]) -> (features[B,nfilt,F] f32, lengths[B] i64)``
contract, same ``state_dict`` keys (``1.window``, ``2.layer.0.fb``).  The four reference stages
(dither/pre-emphasis, power spectrum, mel+log, per-feature normalisation) are *holders* of their
configuration and buffers; the arithmetic runs in two CUDA kernels behind the
``thunder_b200::filterbank`` op (csrc/features.cu).  There is no CPU path.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch
from torch import Tensor, nn

from ..blocks import Masked

__all__ = ["patch_stft", "FeatureBatchNormalizer", "DitherAudio", "PreEmphasisFilter", "PowerSpectrum", "MelScale",
           "FilterbankFeatures", "mel_filterbank"]


def _fused_only(name: str):
    raise NotImplementedError(
        f"{name}.forward is fused into FilterbankFeatures.forward (thunder_b200::filterbank); "
        "call the FilterbankFeatures module instead")


def _hz_to_mel(freq: float) -> float:
    f_sp = 200.0 / 3
    if freq >= 1000.0:
        return 1000.0 / f_sp + math.log(freq / 1000.0) / (math.log(6.4) / 27.0)
    return freq / f_sp


def mel_filterbank(n_freqs: int, n_mels: int, sample_rate: int, f_min: float, f_max: float) -> Tensor:
    """Slaney-scale, slaney-normalised triangular filter bank ``[n_mels, n_freqs]`` -- what the reference
    obtains from ``torchaudio.functional.melscale_fbanks(norm="slaney", mel_scale="slaney")``
    (transform.py:227-240).  Computed in float64 on the host, stored as float32."""
    f_sp = 200.0 / 3
    logstep = math.log(6.4) / 27.0
    min_log_mel = 1000.0 / f_sp
    all_freqs = np.linspace(0, sample_rate // 2, n_freqs)
    m_pts = np.linspace(_hz_to_mel(f_min), _hz_to_mel(f_max), n_mels + 2)
    f_pts = np.where(m_pts >= min_log_mel, 1000.0 * np.exp(logstep * (m_pts - min_log_mel)), f_sp * m_pts)
    f_diff = np.diff(f_pts)
    slopes = f_pts[None, :] - all_freqs[:, None]
    fb = np.maximum(0.0, np.minimum(-slopes[:, :-2] / f_diff[:-1], slopes[:, 2:] / f_diff[1:]))
    fb = fb * (2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels]))[None, :]
    return torch.from_numpy(np.ascontiguousarray(fb.T).astype(np.float32))


class FeatureBatchNormalizer(nn.Module):
    """Per-(batch, feature) masked normalisation, ``div_guard = 1e-5`` added to the std (transform.py:71-92)."""

    def __init__(self):
        super().__init__()
        self.div_guard = 1e-5

    def forward(self, x: Tensor, lengths: Tensor):
        _fused_only("FeatureBatchNormalizer")


class DitherAudio(nn.Module):
    """``x + dither * N(0,1)`` in training, identity in eval (transform.py:95-118)."""

    def __init__(self, dither: float = 1e-5):
        super().__init__()
        self.dither = dither

    def forward(self, x: Tensor) -> Tensor:
        if self.training:
            return x + (self.dither * torch.randn_like(x))
        return x


class PreEmphasisFilter(nn.Module):
    """``y[n] = x[n] - preemph * x[n-1]`` (transform.py:121-144); fused into the feature kernel."""

    def __init__(self, preemph: float = 0.97):
        super().__init__()
        self.preemph = preemph

    def forward(self, x: Tensor):
        _fused_only("PreEmphasisFilter")


class PowerSpectrum(nn.Module):
    """Holder of the STFT configuration and the ``window`` buffer (transform.py:147-208)."""

    def __init__(self, n_window_size: int = 320, n_window_stride: int = 160, n_fft: Optional[int] = None):
        super().__init__()
        if n_window_size <= 0 or n_window_stride <= 0:
            raise ValueError(
                f"{self} got an invalid value for either n_window_size or "
                f"n_window_stride. Both must be positive ints.")
        self.win_length = n_window_size
        self.hop_length = n_window_stride
        self.n_fft = n_fft or 2 ** math.ceil(math.log2(self.win_length))
        self.register_buffer("window", torch.hann_window(self.win_length, periodic=False))

    def get_sequence_length(self, lengths: Tensor) -> Tensor:
        """``floor(len / hop) + 1`` as int64 (transform.py:182-184)."""
        return (torch.floor(lengths / self.hop_length) + 1).to(dtype=torch.long)

    def forward(self, x: Tensor, lengths: Tensor):
        _fused_only("PowerSpectrum")


class MelScale(nn.Module):
    """Holder of the ``fb`` buffer ``[1, nfilt, n_fft//2+1]`` (transform.py:211-255)."""

    def __init__(self, sample_rate: int, n_fft: int, nfilt: int, log_scale: bool = True):
        super().__init__()
        fb = mel_filterbank(1 + n_fft // 2, nfilt, sample_rate, 0.0, sample_rate / 2)
        self.register_buffer("fb", fb.unsqueeze(0))
        self.log_scale = log_scale

    def forward(self, x: Tensor):
        _fused_only("MelScale")


class _FusedFilterbank(nn.Sequential):
    """The object returned by :func:`FilterbankFeatures`.  Children ``0..3`` mirror the reference's
    ``MultiSequential`` so that indexing (``fb[1].window``) and ``state_dict`` keys are unchanged."""

    def __init__(self, dither, preemph, n_window_size, n_window_stride, n_fft, sample_rate, nfilt, augment=()):
        super().__init__(
            Masked(DitherAudio(dither=dither), PreEmphasisFilter(preemph=preemph)),
            PowerSpectrum(n_window_size=n_window_size, n_window_stride=n_window_stride, n_fft=n_fft),
            Masked(MelScale(sample_rate=sample_rate, n_fft=n_fft, nfilt=nfilt)),
            FeatureBatchNormalizer(),
            *[Masked(a) for a in augment],   # children 4.. like the reference factory (transform.py:299-318)
        )
        self._tables = None  # (key, dict of device tensors) derived from the buffers

    # -- derived device tables -------------------------------------------------------------------
    def _device_tables(self, device: torch.device):
        ps: PowerSpectrum = self[1]
        mel: MelScale = self[2].layer[0]
        key = (str(device), ps.window._version, mel.fb._version, ps.window.data_ptr(), mel.fb.data_ptr())
        if self._tables is not None and self._tables[0] == key:
            return self._tables[1]
        n_fft = ps.n_fft
        if not mel.log_scale:
            raise NotImplementedError("MelScale(log_scale=False) is not implemented by the fused kernel")
        win = ps.window.detach().float().cpu().numpy()
        left = (n_fft - ps.win_length) // 2
        wfull = np.zeros(n_fft, np.float32)
        wfull[left:left + ps.win_length] = win
        nz = np.nonzero(wfull)[0]
        win_lo, win_hi = (int(nz[0]), int(nz[-1]) + 1) if nz.size else (0, 1)
        e = np.arange(n_fft, dtype=np.float64)
        tw = np.stack([np.cos(2 * np.pi * e / n_fft), -np.sin(2 * np.pi * e / n_fft)], axis=1).astype(np.float32)
        fb = mel.fb.detach().float().cpu().numpy()[0]  # [nfilt, nbins]
        starts, counts, offs, weights = [], [], [], []
        for row in fb:
            idx = np.nonzero(row)[0]
            if idx.size == 0:
                s, c = 0, 1
            else:
                s, c = int(idx[0]), int(idx[-1] - idx[0] + 1)
            starts.append(s)
            counts.append(c)
            offs.append(len(weights))
            weights.extend(row[s:s + c].tolist())
        t = dict(
            window_full=torch.from_numpy(wfull).to(device),
            twiddle=torch.from_numpy(tw).to(device),
            mel_start=torch.tensor(starts, dtype=torch.int32, device=device),
            mel_count=torch.tensor(counts, dtype=torch.int32, device=device),
            mel_off=torch.tensor(offs, dtype=torch.int32, device=device),
            mel_w=torch.tensor(weights, dtype=torch.float32, device=device),
            win_lo=win_lo, win_hi=win_hi,
        )
        t.update(self._dft_tables(wfull, win_lo, win_hi, fb, device))
        self._tables = (key, t)
        return t

    @staticmethod
    def _dft_tables(wfull: np.ndarray, win_lo: int, win_hi: int, fb: np.ndarray, device) -> dict:
        """Operands of the tensor-core DFT front-end (``ts_logmel_dft``): folded window taps, the split cosine / sine basis
        in the kernel's shared-memory layout, and the filter bank as a sliding two-filter window.  ``dft_ok`` is False when
        the configuration is outside what that kernel implements (then the FFT kernel runs)."""
        n_fft, KM, KS = wfull.shape[0], 160, 10
        ok = n_fft == 512 and win_lo >= 96 and win_hi <= 416 and fb.shape[1] == 257 and fb.shape[0] <= 128
        # sliding window: bin k may only touch filters {j_k, j_k + 1}, j_k non-decreasing
        nf = fb.shape[0]
        w2 = np.zeros((257, 2), np.float32)
        adv = np.zeros(257, np.int32)
        j = 0
        for k in range(257 if ok else 0):
            nzf = np.nonzero(fb[:, k])[0]
            if nzf.size:
                jk = int(nzf[0])
                if jk < j:   # an earlier filter is touched again (its accumulator is gone): not a sliding bank
                    if int(nzf[-1]) > j + 1 or jk < j:
                        ok = False
                        break
                if int(nzf[-1]) > max(jk, j) + 1:
                    ok = False
                    break
                jk = max(jk, j)
                adv[k] = jk - j
                j = jk
                w2[k, 0] = fb[j, k]
                if j + 1 < nf:
                    w2[k, 1] = fb[j + 1, k]
        if not ok:
            return dict(dft_ok=False)
        m = np.arange(KM, dtype=np.float64)
        k = np.arange(256, dtype=np.float64)
        th = 2.0 * np.pi * np.outer(k, m + 0.5) / n_fft           # [bin, m]
        C, S = np.cos(th), np.sin(th)
        S[0, :] = (-1.0) ** np.arange(KM)                          # sine column 0 carries bin 256

        def split(a):
            hi = a.astype(np.float16)
            lo = (a - hi.astype(np.float64)).astype(np.float16)
            return hi, lo

        mats = [*split(C), *split(S)]                              # C_hi, C_lo, S_hi, S_lo
        basis = np.zeros((2, 4, KS, 128 * 16), np.float16)
        n = np.arange(128)
        kk = np.arange(16)
        off = ((n[:, None] // 8) * 256 + (kk[None, :] // 8) * 128 + (n[:, None] % 8) * 16 + (kk[None, :] % 8) * 2) // 2
        for r in range(2):
            for mi, a in enumerate(mats):
                for ks in range(KS):
                    basis[r, mi, ks, off.reshape(-1)] = a[128 * r:128 * r + 128, 16 * ks:16 * ks + 16].reshape(-1)
        return dict(
            dft_ok=True,
            wplus=torch.from_numpy(np.ascontiguousarray(wfull[256:256 + KM])).to(device),
            wminus=torch.from_numpy(np.ascontiguousarray(wfull[255 - KM + 1:256][::-1])).to(device),
            basis=torch.from_numpy(basis.view(np.uint8).reshape(-1).copy()).to(device),
            mel_w2=torch.from_numpy(w2).to(device),
            mel_adv=torch.from_numpy(adv).to(device),
        )

    def features(self, audio: Tensor, lengths: Tensor, bf16_pitch: int = 0, f16: bool = False) -> Tuple[Tensor, Tensor]:
        """Run the fused front-end.  ``bf16_pitch > 0`` emits 16-bit padded rows for the encoder kernels (bf16, or fp16
        with ``f16``)."""
        from .. import ops  # registers torch.ops.thunder_b200.*

        dither: DitherAudio = self[0].layer[0]
        pre: PreEmphasisFilter = self[0].layer[1]
        ps: PowerSpectrum = self[1]
        norm: FeatureBatchNormalizer = self[3]
        with torch.no_grad():
            # train(): DitherAudio's noise is drawn INSIDE the feature kernel (counter-based normals keyed by torch's seed
            # and a device-resident step counter that advances on the stream, so captured graphs draw fresh noise per replay)
            dth = float(dither.dither) if self.training else 0.0
            state = None
            if dth != 0.0:
                state = self.__dict__.get("_dither_state")
                if state is None or state.device != audio.device:
                    state = self.__dict__["_dither_state"] = torch.zeros((1,), dtype=torch.int64, device=audio.device)
                state.add_(0x9E3779B97F4A7C15 - (1 << 64))      # odd increment (golden ratio), wraps mod 2^64
            t = self._device_tables(audio.device)
            from .. import get_stft_kernel

            if dth == 0.0 and get_stft_kernel() == "dft" and t["dft_ok"] and audio.shape[-1] > ps.n_fft // 2:
                # STFT as a DFT-matrix contraction on the tensor cores + normaliser fed by that kernel's partial sums
                feats, feat_len = torch.ops.thunder_b200.filterbank_dft(
                    audio, lengths, t["wplus"], t["wminus"], t["basis"], t["mel_w2"], t["mel_adv"],
                    t["mel_start"].numel(), ps.hop_length, float(pre.preemph), float(norm.div_guard), int(bf16_pitch),
                    bool(f16))
            else:   # shared-memory radix-8 FFT (any window inside n_fft = 512, any filter bank)
                feats, feat_len = torch.ops.thunder_b200.filterbank(
                    audio, lengths, t["window_full"], t["twiddle"], t["mel_start"], t["mel_count"], t["mel_off"],
                    t["mel_w"], ps.hop_length, float(pre.preemph), t["win_lo"], t["win_hi"], float(norm.div_guard),
                    int(bf16_pitch), bool(f16), dth, (int(torch.initial_seed()) & 0x7FFFFFFFFFFFFFFF) if dth != 0.0 else 0, state)
            if self.training and len(self) > 4:   # SpecCutout / SpecAugment: in place on the fresh feature tensor
                from .spec_augment import apply_rects

                F = 1 + audio.shape[-1] // ps.hop_length
                for child in list(self)[4:]:
                    aug = child.layer[0]
                    apply_rects(feats, F, aug.rects(feats.shape[1], F, feats.device))
            return feats, feat_len

    def forward(self, audio: Tensor, lengths: Tensor) -> Tuple[Tensor, Tensor]:
        return self.features(audio, lengths, 0)


def FilterbankFeatures(
    sample_rate: int = 16000,
    n_window_size: int = 320,
    n_window_stride: int = 160,
    n_fft: int = 512,
    preemph: float = 0.97,
    nfilt: int = 64,
    dither: float = 1e-5,
    num_cutout_masks: int = 0,
    num_time_masks: int = 0,
    num_freq_masks: int = 0,
    mask_time_width: int = 50,
    mask_freq_width: int = 20,
) -> nn.Module:
    """Same signature and error behaviour as the reference factory (transform.py:258-321), including the training-time
    SpecCutout / SpecAugment stages appended as children 4.. (identity in eval())."""
    if num_cutout_masks > 0 and (num_freq_masks + num_time_masks > 0):
        raise ValueError("Cutout and SpecAugment can't be used at the same time.")
    from .spec_augment import SpecAugment, SpecCutout

    augment = []
    if num_cutout_masks > 0:
        augment.append(SpecCutout(rect_masks=num_cutout_masks, time_width=mask_time_width, freq_width=mask_freq_width))
    if num_freq_masks + num_time_masks > 0:
        augment.append(SpecAugment(time_masks=num_time_masks, freq_masks=num_freq_masks, time_width=mask_time_width,
                                   freq_width=mask_freq_width))
    return _FusedFilterbank(dither, preemph, n_window_size, n_window_stride, n_fft, sample_rate, nfilt, augment)


def patch_stft(filterbank: nn.Module) -> nn.Module:
    """API counterpart of the reference's ``patch_stft`` (transform.py:324-336), which swaps ``torch.stft`` for a
    convolution-based DFT so that the front-end can be exported to ONNX / run on mobile CPUs.  Here the spectrum never goes
    through ``torch.stft`` -- the fused kernel computes it -- so there is nothing to patch: the filterbank is returned
    unchanged (its outputs are, trivially, identical before and after, which is what the reference's test
    ``tests/quartznet/test_transform_qn.py:286-295`` asserts at atol 1e-3).  Export goes through
    ``CTCModule.to_torchscript`` (custom ops), not ONNX."""
    return filterbank
