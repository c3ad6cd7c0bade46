"""Is the copy-to-copy difference of Citrinet-1024 logits (identical utterances at different batch positions) a bug or
the amplification of fp32-atomic summation order in the SqueezeExcite pool?  Runs the encoder block by block and prints the
relative difference between copy 0 and copy 1 (and between two runs of the same batch) after every block."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from thunder_speech_b200 import synth  # noqa: E402
from thunder_speech_b200.runner import build_model  # noqa: E402

B, secs = 128, 20
m = build_model("citrinet1024", torch.device("cuda"), seed=3)
N = secs * 16000
base = synth.audio(8, N, 77, "tones")
lens8 = np.array([N, N - 1, N // 2 + 123, N // 3, 16000, N - 4000, 3 * N // 4, 4321], np.int64)
if len(sys.argv) > 1 and sys.argv[1] == "full":
    lens8[:] = N
for b in range(8):
    base[b, lens8[b]:] = 0.0
x = torch.from_numpy(np.tile(base, (B // 8, 1))).cuda()
lens = torch.from_numpy(np.tile(lens8, B // 8)).cuda()


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def run():
    outs = []
    with torch.no_grad():
        f, l = m.audio_transform(x, lens)
        outs.append(f)
        h = f
        for blk in m.encoder:
            h, l = blk(h, l)
            outs.append(h)
    return outs


a, b = run(), run()
for i, (u, v) in enumerate(zip(a, b)):
    g = u.view(B // 8, 8, *u.shape[1:])
    per = [rel(g[1, k], g[0, k]) for k in range(8)]
    print(f"block {i - 1:2d} T={u.shape[-1]:5d} copy1-vs-copy0 max {max(per):.3e} (per utt {' '.join(f'{p:.1e}' for p in per)})  run-to-run {rel(u, v):.3e}")
