"""Micro-benchmarks of single ops at the QuartzNet15x5 / Citrinet-1024 layer shapes (CUDA events, L2 flushed by
rotating over buffers larger than L2).  usage: python tools/bench_ops.py [dw|pw|all] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from thunder_speech_b200 import ops, _lib

which = sys.argv[1] if len(sys.argv) > 1 else "all"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda")


def timeit(fn, nbuf):
    for i in range(3):
        fn(i % nbuf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % nbuf)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if which in ("dw", "all"):
    for (B, C, T, K) in [(256, 256, 751, 33), (256, 256, 751, 39), (256, 512, 751, 51), (256, 512, 751, 75),
                         (128, 1024, 1001, 13), (128, 1024, 251, 39)]:
        pitch = ops.row_pitch(T)
        nbuf = max(2, int(300e6 // (B * C * pitch * 2)) + 1)
        xs = [torch.randn(B, C, pitch, device=dev).bfloat16() for _ in range(nbuf)]
        w = torch.randn(C, K, device=dev) / K
        for mode in (1,):
            _lib.set_option("dw_mma", mode)
            ms = timeit(lambda i: ops.dw_conv(xs[i], T, w, 1, 1, K // 2, None), nbuf)
            gb = 2 * B * C * T * 2 / 1e9
            print(f"dw  B={B} C={C} T={T} K={K} mma={mode}: {ms*1e3:8.1f} us  {gb/ms*1e3:7.0f} GB/s  ({gb/ms*1e3/6532*100:4.1f}% of measured HBM)")
        _lib.set_option("dw_mma", 1)

if which in ("pw", "all"):
    for (B, Cin, Cout, T, res) in [(256, 256, 256, 751, 0), (256, 256, 256, 751, 256), (256, 512, 512, 751, 0),
                                   (256, 512, 512, 751, 512), (256, 512, 1024, 751, 0), (128, 1024, 1024, 1001, 0),
                                   (128, 1024, 1024, 251, 0)]:
        pitch = ops.row_pitch(T)
        nbuf = max(2, int(300e6 // (B * Cin * pitch * 2)) + 1)
        xs = [torch.randn(B, Cin, pitch, device=dev).bfloat16() for _ in range(nbuf)]
        w = (torch.randn(Cout, Cin, device=dev) / Cin ** 0.5).bfloat16()
        w1 = (torch.randn(Cout, res, device=dev) / max(res, 1) ** 0.5).bfloat16() if res else None
        x1 = torch.randn(B, res, pitch, device=dev).bfloat16() if res else None
        shift = torch.randn(Cout, device=dev)
        ms = timeit(lambda i: ops.pw_gemm(w, xs[i], w1, x1, T, shift, None, False, True, None, None, None), nbuf)
        fl = 2 * B * T * (Cin + res) * Cout / 1e12
        by = (2 * B * T * (Cin + res) + 2 * B * T * Cout + 2 * Cout * (Cin + res)) / 1e9
        print(f"pw  B={B} {Cin}(+{res})->{Cout} T={T}: {ms*1e3:8.1f} us  {fl/ms*1e3:7.0f} TFLOP/s  {by/ms*1e3:7.0f} GB/s")
