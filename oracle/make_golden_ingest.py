"""tests/golden/ingest.npz: outputs of the REFERENCE's AudioFileLoader.preprocess_audio (src/thunder/data/dataset.py:50-77,
torchaudio resample) on synthetic clips.  Run in the build container (needs /root/reference):  python -m oracle.make_golden_ingest"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [  # name, channels, sample_rate, seconds
    ("stereo44k", 2, 44100, 0.40), ("mono8k", 1, 8000, 0.50), ("mono48k", 1, 48000, 0.30), ("mono22k", 1, 22050, 0.25),
    ("mono16k", 1, 16000, 0.20), ("stereo32k", 2, 32000, 0.30),
]


def clip(name, ch, sr, secs):
    """int16 PCM [channels, time] with a DC offset (what torchaudio.load would have normalised by 1/32768)."""
    rng = np.random.Generator(np.random.PCG64(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name))))
    n = int(sr * secs)
    t = np.arange(n) / sr
    x = 0.3 * np.sin(2 * np.pi * 440.0 * t)[None, :] + 0.05 * rng.standard_normal((ch, n)) + 0.02 * (1 + np.arange(ch))[:, None]
    return np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)


def main():
    pkg = types.ModuleType("thunder"); pkg.__path__ = ["/root/reference/src/thunder"]; sys.modules["thunder"] = pkg
    spec = importlib.util.spec_from_file_location("thunder.data.dataset", "/root/reference/src/thunder/data/dataset.py")
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    loader = mod.AudioFileLoader(force_mono=True, sample_rate=16000)
    out = {}
    for name, ch, sr, secs in CASES:
        pcm = clip(name, ch, sr, secs)
        audio = torch.from_numpy(pcm.astype(np.float32) / 32768.0)      # torchaudio.load(normalize=True) convention
        y = loader.preprocess_audio(audio, sr)
        out[f"{name}.out"] = y.numpy()
        print(name, pcm.shape, "->", tuple(y.shape))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ingest.npz"), **out)


if __name__ == "__main__":
    main()
