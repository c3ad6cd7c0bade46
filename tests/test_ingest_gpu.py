"""GPU parity of the audio-ingest kernels through the C ABI: goldens produced by the reference's
AudioFileLoader.preprocess_audio, and the numpy oracle on ragged batches."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import ref_numpy as R
from oracle.make_golden_ingest import CASES, clip
from thunder_speech_b200.data import AudioFileLoader, pcm_ingest, resample

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_preprocess_audio_vs_reference_goldens(case):
    name, ch, sr, secs = case
    g = np.load("tests/golden/ingest.npz")
    pcm = clip(name, ch, sr, secs)                                       # int16 [channels, time]
    loader = AudioFileLoader(force_mono=True, sample_rate=16000)
    ref = g[f"{name}.out"]
    # int16 PCM in (the device scales by 1/32768 like torchaudio.load) and the float path
    y16 = loader.preprocess_audio(torch.from_numpy(pcm).cuda(), sr).cpu().numpy()
    yf = loader.preprocess_audio(torch.from_numpy(pcm.astype(np.float32) / 32768.0).cuda(), sr).cpu().numpy()
    assert y16.shape == ref.shape and yf.shape == ref.shape
    assert rel_err(y16, ref)[0] < 1e-4 and rel_err(yf, ref)[0] < 1e-4, (name, rel_err(y16, ref), rel_err(yf, ref))


@pytest.mark.parametrize("orig,new", [(44100, 16000), (8000, 16000), (48000, 16000), (22050, 16000), (32000, 16000), (16000, 8000)])
def test_ragged_batch_ingest_and_resample_vs_oracle(orig, new):
    """A padded batch with ragged lengths, interleaved int16 stereo: every utterance must come out exactly as if it had been
    processed alone (the reference processes files one by one before asr_collate pads them)."""
    rng = np.random.default_rng(orig // 100 + new // 100)
    B, C, N = 5, 2, int(0.21 * orig)
    lens = np.array([N, N - 17, N // 2 + 3, 1000, 1], np.int64)
    pcm = np.clip(np.round((0.2 * rng.standard_normal((B, N, C)) + 0.03) * 32768), -32768, 32767).astype(np.int16)
    mono = pcm_ingest(torch.from_numpy(pcm).cuda(), torch.from_numpy(lens).cuda(), interleaved=True)
    y, nl = resample(mono, orig, new, torch.from_numpy(lens).cuda())
    y, nl, mono = y.cpu().numpy(), nl.cpu().numpy(), mono.cpu().numpy()
    for b in range(B):
        one = pcm[b, : lens[b]].T.astype(np.float32) / 32768.0          # [channels, len]
        ref = R.preprocess_audio(one, orig, True, new)[0]
        assert nl[b] == ref.shape[0], (b, nl[b], ref.shape)
        assert np.all(y[b, nl[b]:] == 0) and np.all(mono[b, lens[b]:] == 0)
        if ref.size:
            assert np.abs(y[b, : nl[b]] - ref).max() <= 1e-4 * max(np.abs(ref).max(), 1e-3), (b, orig, new)


def test_ingest_error_behaviour_and_planar_layout():
    loader = AudioFileLoader(force_mono=False)
    with pytest.raises(RuntimeError):                                   # reference: audio - audio.mean(1) cannot broadcast
        loader.preprocess_audio(torch.zeros((2, 100), device="cuda"), 16000)
    with pytest.raises(RuntimeError):
        pcm_ingest(torch.zeros((1, 1, 8)))                              # CPU tensor: no fallback
    with pytest.raises(TypeError):
        resample(torch.zeros((1, 8), device="cuda", dtype=torch.int32), 8000, 16000)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((3, 2, 70001)).astype(np.float32)           # planar float, > one 65536-sample chunk
    got = pcm_ingest(torch.from_numpy(x).cuda()).cpu().numpy()
    ref = x.mean(1)
    ref = ref - ref.mean(1, keepdims=True)
    assert np.abs(got - ref).max() < 1e-5
