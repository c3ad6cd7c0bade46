"""GPU parity tests of the fused feature front-end (csrc/features.cu) through the C ABI.

Tolerance (BASELINE.json north_star): log-mel features within 1e-4 relative (fp32), lengths exact."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import ref_numpy as R
from thunder_speech_b200 import synth
from thunder_speech_b200.quartznet.transform import FilterbankFeatures

pytestmark = pytest.mark.gpu
TOL = 1e-4

FEATURE_CASES = [("qn_noise", "noise"), ("qn_tones", "tones"), ("cn_noise", "noise"), ("short", "noise")]


def run_cuda(x, lens, nfilt=64, **kw):
    fb = FilterbankFeatures(nfilt=nfilt, **kw).eval().cuda()
    f, fl = fb(torch.from_numpy(x).cuda(), torch.as_tensor(lens).cuda())
    torch.cuda.synchronize()
    return f.cpu().numpy(), fl.cpu().numpy()


@pytest.mark.parametrize("name,kind", FEATURE_CASES)
def test_features_vs_golden_and_oracle(golden_features, name, kind):
    g = golden_features
    nfilt, B, N, seed = [int(v) for v in g[f"{name}.meta"]]
    x = synth.audio(B, N, seed, kind)
    lens = synth.ragged_lengths(B, N, seed + 100)
    f, fl = run_cuda(x, lens, nfilt)
    assert f.dtype == np.float32 and fl.dtype == np.int64
    assert np.array_equal(fl, g[f"{name}.lengths"])
    emax, el2 = rel_err(f, g[f"{name}.features"])       # the reference's own output
    assert emax < TOL and el2 < TOL, (emax, el2)
    rf, rl = R.filterbank_features(x, lens, nfilt=nfilt)  # the oracle
    emax, el2 = rel_err(f, rf)
    assert emax < TOL and el2 < TOL, (emax, el2)
    # float lengths (asr_collate contract) behave the same
    f2, fl2 = run_cuda(x, lens.astype(np.float32), nfilt)
    assert np.array_equal(fl2, fl) and np.array_equal(f2, f)


@pytest.mark.parametrize("N", [257, 320, 1599, 1600, 1601, 5121, 31 * 160, 32 * 160, 33 * 160 + 7])
def test_features_ragged_shapes(N):
    """Frame-tile edges (32 frames per CTA), reflect padding near both ends, odd N."""
    x = synth.audio(3, N, N, "noise")
    lens = np.array([N, max(1, N // 2), 0], np.int64)
    f, fl = run_cuda(x, lens)
    rf, rl = R.filterbank_features(x, lens)
    assert f.shape == rf.shape == (3, 64, 1 + N // 160)
    assert np.array_equal(fl, rl)
    assert np.isfinite(f).all()
    emax, el2 = rel_err(f, rf)
    assert emax < TOL and el2 < TOL, (N, emax, el2)


def test_features_other_hop_and_window():
    """Generic (non 320-centred) window path and a different hop."""
    x = synth.audio(2, 6000, 5, "tones")
    lens = np.array([6000, 4100], np.int64)
    for win, hop in ((400, 160), (512, 128), (320, 200)):
        f, fl = run_cuda(x, lens, 64, n_window_size=win, n_window_stride=hop)
        rf, rl = R.filterbank_features(x, lens, n_window_size=win, n_window_stride=hop)
        assert np.array_equal(fl, rl)
        emax, el2 = rel_err(f, rf)
        assert emax < TOL and el2 < TOL, (win, hop, emax, el2)


def test_features_errors():
    with pytest.raises(ValueError):
        FilterbankFeatures(n_window_size=0)
    with pytest.raises(ValueError):
        FilterbankFeatures(num_cutout_masks=1, num_time_masks=1)
    fb = FilterbankFeatures().eval().cuda()
    with pytest.raises(ValueError):  # reflect padding needs N > n_fft/2, like torch.stft
        fb(torch.zeros(1, 200).cuda(), torch.tensor([200]).cuda())
    with pytest.raises(RuntimeError):  # no CPU fallback
        FilterbankFeatures().eval()(torch.zeros(1, 2000), torch.tensor([2000]))


def test_features_config2_full_size_properties():
    """BASELINE config 2: B=64 x 20 s.  Size-independent properties + oracle on a slice of the batch."""
    B, N = 64, 320000
    x = synth.audio(B, N, 1234, "noise")
    lens = np.full((B,), N, np.int64)
    f, fl = run_cuda(x, lens)
    assert f.shape == (B, 64, 2001) and (fl == 2001).all()
    assert np.isfinite(f).all()
    # normalised per (b, feature): mean 0, biased std 1 (tests/quartznet/test_transform_qn.py:43-51 at atol 0.1)
    assert np.abs(f.mean(-1)).max() < 1e-3
    assert np.abs(f.std(-1) - 1).max() < 1e-3
    # batch independence: utterance b only depends on row b
    sel = [0, 17, 63]
    rf, _ = R.filterbank_features(x[sel], lens[sel])
    emax, el2 = rel_err(f[sel], rf)
    assert emax < TOL and el2 < TOL, (emax, el2)


def test_features_bf16_padded_rows():
    x = synth.audio(2, 16000, 9, "tones")
    lens = np.array([16000, 12345], np.int64)
    fb = FilterbankFeatures().eval().cuda()
    pitch = 128
    f, fl = fb.features(torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda(), bf16_pitch=pitch)
    assert f.dtype == torch.bfloat16 and f.shape == (2, 64, pitch)
    rf, rl = R.filterbank_features(x, lens)
    got = f.float().cpu().numpy()
    assert np.array_equal(fl.cpu().numpy(), rl)
    emax, el2 = rel_err(got[:, :, :101], rf)
    assert emax < 2e-2 and el2 < 2e-2
    assert (got[:, :, 101:] == 0).all()
