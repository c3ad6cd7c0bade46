"""A/B of the serving pipeline depth (predict_stream(depth=...)) on the default bench workload, plus the raw
pinned-host -> device bandwidth of one audio batch, to tell a PCIe-bound e2e from a pipeline bubble."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from thunder_speech_b200.runner import make_bench_workload  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
wl = make_bench_workload("quartznet15x5", 256, 15 * 16000, 64, dev, 0)
out = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    wl.stage.copy_(wl.host_audio, non_blocking=True)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    wl.stage.copy_(wl.host_audio, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
out["h2d_ms"] = ms
out["h2d_GBps"] = wl.h2d_bytes / ms / 1e6
steps = 40
e0.record()
for i in range(steps):
    wl.step_device(i)
e1.record()
torch.cuda.synchronize()
out["device_ms"] = e0.elapsed_time(e1) / steps
for rnd in range(2):
    for depth in (2, 3, 4):
        list(wl.model.predict_stream((wl.host_audio for _ in range(4)), depth=depth))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 0
        for texts in wl.model.predict_stream((wl.host_audio for _ in range(steps)), depth=depth):
            n += len(texts)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        out[f"e2e_ms_depth{depth}_r{rnd}"] = dt * 1e3
        out[f"e2e_audio_s_per_s_depth{depth}_r{rnd}"] = wl.B * wl.N / 16000 / dt
print(json.dumps(out, indent=1))
