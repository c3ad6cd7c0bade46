"""dw_conv per shape under graph replay, with/without shared halo rows: python tools/bench_dwconv.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from thunder_speech_b200 import _lib, ops
def gtime(f, iters=10):
    for _ in range(2): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
shapes = [(128, 1024, 251, 25), (128, 1024, 501, 19), (128, 1024, 1001, 13), (256, 512, 751, 63), (32, 512, 751, 63)]
for (B, C, T, K) in shapes:
    pitch = ops.row_pitch(T)
    # 4 distinct inputs so that the working set exceeds L2
    xs = [torch.randn(B, C, pitch, device="cuda").to(torch.bfloat16) for _ in range(4)]
    for x in xs: x[:, :, T:] = 0
    w = torch.randn(C, K, device="cuda")
    res = []
    PRO = int(os.environ.get("PRO", "10"))
    _lib.set_option("dw_pro", PRO)
    for o in (0, 1):
        _lib.set_option("dw_share_halo", o)
        i = [0]
        def f():
            i[0] += 1
            return ops.dw_conv(xs[i[0] % 4], T, w, 1, 1, K // 2, None, True)
        us = gtime(f, 12)
        res.append(us)
    by = 2 * B * C * T * 2
    print(f"B {B} C {C} T {T} K {K}: share0 {res[0]:7.1f} us ({by/res[0]/1e3:5.0f} GB/s)  share1 {res[1]:7.1f} us ({by/res[1]/1e3:5.0f} GB/s)")
