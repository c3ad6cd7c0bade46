"""Where does predict_stream lose time against back-to-back graph replays?  Records CUDA events around every graph replay
of the serving loop and prints the mean replay duration and the mean idle gap between consecutive replays."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from thunder_speech_b200 import module as M  # noqa: E402
from thunder_speech_b200.runner import make_bench_workload  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
wl = make_bench_workload("quartznet15x5", 256, 15 * 16000, 64, dev, 0)
EV = []
orig = M._PredictGraph.replay


def replay(self, x):
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    self.static_in.copy_(x, non_blocking=True)
    b.record()
    self.graph.replay()
    c.record()
    self.replays += 1
    EV.append((a, b, c))
    return self.static_out


M._PredictGraph.replay = replay
out = {}
steps = 40
for mode in ("device", "stream2", "stream3", "stream3_nodecode"):
    EV.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if mode == "device":
        for i in range(steps):
            wl.step_device(i)
    else:
        if mode.endswith("nodecode"):
            keep = wl.model.text_transform.decode_collapsed
            wl.model.text_transform.decode_collapsed = lambda col, cnt: []
        for _ in wl.model.predict_stream((wl.host_audio for _ in range(steps)), depth=int(mode[6])):
            pass
        if mode.endswith("nodecode"):
            wl.model.text_transform.decode_collapsed = keep
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / steps * 1e3
    ev = EV[2:]
    copy = sum(a.elapsed_time(b) for a, b, c in ev) / len(ev)
    graph = sum(b.elapsed_time(c) for a, b, c in ev) / len(ev)
    gaps = [ev[i][2].elapsed_time(ev[i + 1][0]) for i in range(len(ev) - 1)]
    span = ev[0][0].elapsed_time(ev[-1][2]) / (len(ev) - 1 + 1e-9)
    out[mode] = dict(wall_ms=wall, d2d_copy_ms=copy, graph_ms=graph, gap_ms=sum(gaps) / len(gaps), max_gap_ms=max(gaps),
                     steady_ms_per_batch=ev[0][0].elapsed_time(ev[-1][0]) / (len(ev) - 1))
print(json.dumps(out, indent=1))
