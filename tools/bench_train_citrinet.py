"""Citrinet-1024 training step timing (not a BASELINE config; data point for DESIGN.md): python tools/bench_train_citrinet.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from thunder_speech_b200 import synth
from thunder_speech_b200.runner import build_model
from thunder_speech_b200.train import CTCTrainStep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N, L = 20 * 16000, 60
dev = torch.device("cuda")
m = build_model("citrinet1024", dev); m.train()
step = CTCTrainStep(m, lr=1e-4)
rng = np.random.default_rng(0)
x = torch.from_numpy(synth.audio(B, N, 1, "noise")).to(dev)
lens = torch.from_numpy(synth.ragged_lengths(B, N, 2)).to(dev)
y = torch.from_numpy(rng.integers(0, 1024, (B, L)).astype(np.int64)).to(dev)
yl = torch.from_numpy(rng.integers(L // 2, L + 1, B).astype(np.int64)).to(dev)
for _ in range(3): loss = step.step(x, lens, y, yl)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): loss = step.step(x, lens, y, yl)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"citrinet1024 train B={B} x 20 s: {ms:.2f} ms/step  {B * 20 / ms * 1e3:.0f} audio-s/s  loss {float(loss):.4f}  mem {torch.cuda.max_memory_allocated() / 1e9:.1f} GB  params {sum(p.numel() for p in step.params) / 1e6:.1f} M")
