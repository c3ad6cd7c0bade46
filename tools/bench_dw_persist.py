"""A/B of the persistent multi-channel Toeplitz kernel (dwmma3.cu) against the per-channel one (dwmma2.cu):
bit-identical outputs, time per launch, algorithmic GB/s."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from thunder_speech_b200 import ops, _lib

def timeit(fn, n=20):
    for _ in range(3): fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

shapes = [(256, 256, 751, 33, 1), (256, 512, 751, 75, 1), (256, 512, 751, 87, 2), (32, 256, 751, 33, 1), (32, 512, 751, 75, 1),
          (32, 512, 751, 87, 2), (128, 1024, 1001, 11, 1), (128, 1024, 251, 39, 1), (16, 1024, 1001, 11, 1), (16, 1024, 251, 39, 1),
          (3, 40, 100, 5, 1), (1, 8, 64, 3, 1), (5, 24, 700, 129, 1)]
for B, C, T, K, D in shapes:
    P = ops.row_pitch(T)
    pad = D * (K - 1) // 2
    w = torch.randn(C, K, device="cuda") * 0.1
    lens = torch.randint(T // 2, T + 1, (B,), device="cuda", dtype=torch.int32)
    lens[0] = T
    nb = 3 if B * C * P * 2 < 200e6 else 2
    for dt in (torch.bfloat16, torch.float16):
        xs = []
        for _ in range(nb):
            x = torch.randn(B, C, P, device="cuda")
            x = x * (torch.arange(P, device="cuda")[None, None, :] < lens[:, None, None])
            xs.append(x.to(dt))
        res = {}
        for opt in (0, 2):
            _lib.set_option("dw_persist", opt)
            _lib.set_option("serpentine", 0)
            y = ops.dw_conv(xs[0], T, w, 1, D, pad, lens, True)
            us = timeit(lambda i: ops.dw_conv(xs[i % nb], T, w, 1, D, pad, lens, True))
            res[opt] = (y, us)
        same = torch.equal(res[0][0], res[2][0])
        gb = 2 * B * C * T * 2 / 1e9
        print(f"B={B:3d} C={C:4d} T={T:4d} K={K:3d} D={D} {str(dt)[6:]:8s} per-channel {res[0][1]:7.1f} us ({gb/res[0][1]*1e6:5.0f} GB/s)"
              f"  persistent {res[2][1]:7.1f} us ({gb/res[2][1]*1e6:5.0f} GB/s)  identical={same}", flush=True)
        assert same
