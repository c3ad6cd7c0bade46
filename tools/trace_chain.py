#!/usr/bin/env python
"""Per-kernel phase timeline of the captured inference graph (ts_trace): where do the ~13 us per kernel go at small batch?

    python tools/trace_chain.py [model] [batch] [first] [count]

For every traced launch (pair GEMM = 1, persistent Toeplitz = 2) prints, for CTA 0 and the last CTA, the times in us
relative to the END of the previous traced kernel's CTA 0:
    entry  pro(logue done)  dep (predecessor complete)  ops (first operands in smem)  mma (last MMA issued)
    acc (first accumulator ready)  sto (last store issued)  drn (stores drained)  end
"""
import os
import sys

os.environ.setdefault("THUNDER_B200_TRACE_BUILD", "1")   # the library build with the trace hooks (make -C thunder_speech_b200/csrc trace)
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from thunder_speech_b200 import _lib, runner, synth  # noqa: E402

NAMES = ["entry", "pro", "dep", "ops", "mma", "acc", "sto", "drn", "end"]


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "quartznet15x5"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    first = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    count = int(sys.argv[4]) if len(sys.argv) > 4 else 12
    secs = 15 if name.startswith("quartznet") else 20
    dev = torch.device("cuda", 0)
    model = runner.build_model(name, dev)
    x = torch.from_numpy(synth.audio(B, secs * 16000, 1234, "noise")).to(dev)
    model.predict_ids(x)           # plans, kernel attributes
    torch.cuda.synchronize()
    slots = 512
    buf = torch.zeros((slots, 32), dtype=torch.int64, device=dev)
    L = _lib.lib()
    # the graph's warm-up passes take slots too: install the buffer only for the capture itself
    g = None
    from thunder_speech_b200.module import _PredictGraph

    class G(_PredictGraph):
        pass

    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(2):
            model.predict_ids(x)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    _lib.check(L.ts_trace(buf.data_ptr(), slots), "ts_trace")
    with torch.cuda.graph(graph):
        out = model.predict_ids(x)
    _lib.check(L.ts_trace(None, 0), "ts_trace")
    for _ in range(5):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name} B={B}: {e0.elapsed_time(e1) / 10:.3f} ms per replay")
    t = buf.cpu().numpy()
    used = [i for i in range(slots) if t[i, 0] != 0]
    print(f"{len(used)} traced launches")
    prev_end = None
    tot = {}
    for n, i in enumerate(used):
        kid, grid = int(t[i, 0]) & 0xFF, int(t[i, 0]) >> 8
        a = t[i, 1:10].astype("float64")
        b = t[i, 17:26].astype("float64")
        if prev_end is None:
            prev_end = a[0]
        if first <= n < first + count:
            ra = " ".join(f"{NAMES[k]}={(a[k] - prev_end) / 1e3:6.2f}" for k in range(9))
            rb = " ".join(f"{NAMES[k]}={(b[k] - prev_end) / 1e3:6.2f}" for k in range(9)) if b[0] > 0 else "-"
            print(f"#{n:3d} {'pw' if kid == 1 else 'dw'} grid={grid:3d}\n     first CTA: {ra}\n     last  CTA: {rb}")
        if n > 0:
            d = tot.setdefault(kid, [0.0] * 10 + [0])
            end = max(a[8], b[8])
            for k in range(9):
                d[k] += (a[k] - prev_end) / 1e3
            d[9] += (end - prev_end) / 1e3
            d[10] += 1
        prev_end = max(a[8], b[8])
    for kid, d in tot.items():
        n = d[10]
        print(("pw" if kid == 1 else "dw") + f" mean over {n} launches (us after predecessor's end, first CTA): "
              + " ".join(f"{NAMES[k]}={d[k] / n:6.2f}" for k in range(9)) + f" | kernel end (max of both CTAs)={d[9] / n:6.2f}")


if __name__ == "__main__":
    main()
