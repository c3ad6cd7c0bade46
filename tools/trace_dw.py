#!/usr/bin/env python
"""Stall accounting of the Toeplitz depthwise kernel (ts_trace, CTA 0): who waits for whom?

    python tools/trace_dw.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("THUNDER_B200_TRACE_BUILD", "1")   # the library build with the trace hooks (make -C thunder_speech_b200/csrc trace)
import torch
from thunder_speech_b200 import ops, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = 751; P = ops.row_pitch(T)
dev = torch.device("cuda")
lens = torch.full((B,), T, dtype=torch.int32, device=dev)
L = _lib.lib()
DBG = int(os.environ.get("DBG", "0"))
_lib.set_option("dbg", DBG)
print(f"dbg={DBG}")
for C, K in ((256, 33), (512, 51), (512, 75)):
    xs = [torch.randn(B, C, P, device=dev).bfloat16() for _ in range(3)]
    for x in xs:
        x[:, :, T:] = 0
    w = torch.randn(C, K, device=dev) / K ** 0.5
    run = lambda i: ops.dw_conv(xs[i % 3], T, w, 1, 1, (K - 1) // 2, lens, True)
    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us_k = e0.elapsed_time(e1) / 20 * 1e3
    buf = torch.zeros((32 + 2 * 8192,), dtype=torch.int64, device=dev)
    _lib.set_option("dbg", DBG | 128)
    _lib.check(L.ts_trace(buf.data_ptr(), 1), "ts_trace")
    run(0)
    _lib.check(L.ts_trace(None, 0), "ts_trace")
    _lib.set_option("dbg", DBG)
    torch.cuda.synchronize()
    t = buf.cpu().numpy()
    import numpy as np
    ct = t[32:].reshape(-1, 2).astype("float64")
    ct = ct[ct[:, 0] > 0]
    st, en = (ct[:, 0] - ct[:, 0].min()) / 1e3, (ct[:, 1] - ct[:, 0].min()) / 1e3
    life = en - st
    print(f"   {len(ct)} CTAs: start p50/p90/max {np.percentile(st, 50):5.1f}/{np.percentile(st, 90):5.1f}/{st.max():5.1f} us, "
          f"life min/p50/p90/max {life.min():5.1f}/{np.percentile(life, 50):5.1f}/{np.percentile(life, 90):5.1f}/{life.max():5.1f} us, "
          f"span {en.max():5.1f} us")
    us = (t[9] - t[1]) / 1e3
    loop = max(int(t[15]), 1)
    by = 2 * B * C * T * 2
    print(f"B={B} C={C} K={K}: kernel {us_k:6.1f} us = {by / us_k / 1e3:5.0f} GB/s | CTA0 life {us:6.1f} us: entry->pro {(t[2]-t[1])/1e3:4.1f} "
          f"dep->ops {(t[4]-t[3])/1e3:4.1f}  issuer waits: operands {100 * t[10] / loop:4.1f}% accumulator {100 * t[11] / loop:4.1f}% | "
          f"producer waits for a stage {100 * t[12] / loop:4.1f}% | epilogue: accumulators {100 * t[13] / loop:4.1f}% staging read "
          f"{100 * t[14] / loop:4.1f}%  kind={int(t[0]) & 255} grid={int(t[0]) >> 8}", flush=True)
