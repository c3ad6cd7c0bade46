"""Per-shape timing of the pair GEMM (QuartzNet shapes at B x 751 frames) with optional experiment knobs:
python tools/bench_gemm_shapes.py [B] [dbg values...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from thunder_speech_b200 import ops, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dbgs = [int(v) for v in sys.argv[2:]] or [0]
T = 751; P = ops.row_pitch(T)
dev = torch.device("cuda")
lens = torch.full((B,), T, dtype=torch.int32, device=dev)

def timeit(fn, n=30):
    for _ in range(5): fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for cin, cout, res in ((256, 256, 0), (256, 256, 256), (512, 512, 0), (512, 512, 512), (512, 1024, 0)):
    xs = [torch.randn(B, cin, P, device=dev).bfloat16() for _ in range(3)]
    w = (torch.randn(cout, cin, device=dev) / cin ** 0.5).bfloat16()
    sh = torch.randn(cout, device=dev)
    w1 = (torch.randn(cout, res, device=dev) / max(res, 1) ** 0.5).bfloat16() if res else None
    x1 = [torch.randn(B, res, P, device=dev).bfloat16() for _ in range(3)] if res else None
    fl = 2 * B * T * (cin + res) * cout
    out = []
    for d in dbgs:
        _lib.set_option("dbg", d)
        us = timeit(lambda i: ops.pw_gemm(w, xs[i % 3], w1, x1[i % 3] if res else None, T, sh, lens, False, True, None, None, None))
        out.append(f"dbg={d}: {us:7.1f} us {fl / us / 1e6:6.0f} TFLOP/s")
    print(f"B={B} {cin}+{res}->{cout}: " + " | ".join(out), flush=True)
_lib.set_option("dbg", 0)
