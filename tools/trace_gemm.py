#!/usr/bin/env python
"""Stall accounting of the pair GEMM (ts_trace, CTA 0 = leader of pair 0): where does the MMA issuer wait?

    python tools/trace_gemm.py [B]

Per QuartzNet shape at B x 751 frames: kernel time, and the share of the issue loop's life spent waiting for operands
(full barrier: load-starved) and for a free accumulator (tmem_empty: epilogue-bound); producer waiting for a free stage;
epilogue warp 4 waiting for accumulators / for the staging tile to be read by the previous TMA store."""
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
for cin, cout, res in ((256, 256, 0), (256, 256, 256), (512, 512, 0), (512, 512, 512), (512, 1024, 0), (1024, 1024, 0)):
    x = torch.randn(B, cin, P, device=dev).bfloat16()
    w = (torch.randn(cout, cin, device=dev) / cin ** 0.5).bfloat16()
    sh = torch.randn(cout, device=dev)
    w1 = (torch.randn(cout, res, device=dev) / max(res, 1) ** 0.5).bfloat16() if res else None
    x1 = torch.randn(B, res, P, device=dev).bfloat16() if res else None
    run = lambda: ops.pw_gemm(w, x, w1, x1, T, sh, lens, False, True, None, None, None)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    buf = torch.zeros((4, 32), dtype=torch.int64, device=dev)
    _lib.check(L.ts_trace(buf.data_ptr(), 1), "ts_trace")
    run()
    _lib.check(L.ts_trace(None, 0), "ts_trace")
    torch.cuda.synchronize()
    t = buf.cpu().numpy()[0]
    us = (t[9] - t[1]) / 1e3
    loop = max(int(t[15]), 1)
    fl = 2 * B * T * (cin + res) * cout
    print(f"B={B} {cin}+{res}->{cout}: CTA0 life {us:7.1f} us ({fl / us / 1e6:5.0f} TFLOP/s)  issuer: operands {100 * t[10] / loop:4.1f}% "
          f"accumulator {100 * t[11] / loop:4.1f}%  | producer waits for a stage {100 * t[12] / loop:4.1f}%  | epilogue w4: "
          f"accumulators {100 * t[13] / loop:4.1f}% staging {100 * t[14] / loop:4.1f}%", flush=True)
