"""Unusual shapes through the public API (no parity here, only: runs, finite, shapes / lengths right)."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from thunder_speech_b200 import synth
from thunder_speech_b200.runner import build_model
from thunder_speech_b200.train import CTCTrainStep
dev = torch.device("cuda")
fails = 0
def check(name, fn):
    global fails
    try:
        fn(); torch.cuda.synchronize(); print("ok  ", name)
    except Exception as e:
        fails += 1; print("FAIL", name, "->", type(e).__name__, str(e).splitlines()[0][:160])
for model in ("quartznet5x5", "citrinet1024"):
    m = build_model(model, dev, seed=1)
    for (B, N) in [(1, 400), (1, 16000), (3, 1599), (1, 16000 * 60), (7, 12345), (130, 8000)]:   # N <= 256 cannot be reflect-padded: ValueError, like torch.stft
        def f():
            x = torch.from_numpy(synth.audio(B, N, 5, "noise")).to(dev)
            t = m.predict(x)
            assert len(t) == B
            lens = torch.randint(1, N + 1, (B,), device=dev); lens[0] = N
            lg, ol = m(x, lens)
            assert torch.isfinite(lg).all() and lg.shape[0] == B and (ol > 0).all()
        check(f"{model} predict/forward B={B} N={N}", f)
    def f0():
        x = torch.from_numpy(synth.audio(2, 8000, 5, "noise")).to(dev)
        lg, ol = m(x, torch.tensor([8000, 0], device=dev))      # a zero-length utterance in the batch
        assert torch.isfinite(lg).all()
    check(f"{model} zero-length utterance", f0)
m = build_model("quartznet5x5", dev, seed=2); m.encoder.train(); m.decoder.train()
for (B, N, L) in [(1, 16000, 5), (3, 24000, 7), (5, 8000, 1), (2, 3200, 3), (9, 40000, 40)]:
    def f():
        step = CTCTrainStep(m, lr=1e-4)
        x = torch.from_numpy(synth.audio(B, N, 6, "noise")).to(dev)
        lens = torch.randint(N // 2, N + 1, (B,), device=dev); lens[0] = N
        y = torch.randint(0, 28, (B, L), device=dev); yl = torch.randint(1, L + 1, (B,), device=dev)
        for _ in range(2):
            loss = step.step(x, lens, y, yl)
        assert torch.isfinite(loss), loss
        assert torch.isfinite(step.flat).all()
    check(f"train step B={B} N={N} L={L}", f)
print("failures:", fails)
sys.exit(1 if fails else 0)
