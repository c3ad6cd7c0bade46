"""A/B of the two row formats on the hot kernels (CUDA events, graph-free, inputs > L2 by rotation)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from thunder_speech_b200 import ops

def timeit(fn, n=20):
    for _ in range(5): fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

B, T = 256, 751
P = ops.row_pitch(T)
lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
for C, K in ((256, 33), (512, 75)):
    w = torch.randn(C, K, device="cuda") * 0.1
    for dt in (torch.bfloat16, torch.float16):
        xs = [(torch.randn(B, C, P, device="cuda") * (lens[:, None, None] > -1)).to(dt) for _ in range(3)]
        for x in xs: x[:, :, T:] = 0
        us = timeit(lambda i: ops.dw_conv(xs[i % 3], T, w, 1, 1, K // 2, lens, True))
        print(f"dw_conv C={C} K={K} {dt}: {us:.1f} us  {2*B*C*T*2*2/us/1e3:.0f} GB/s")
        wt = (torch.randn(C, C, device="cuda") * 0.05).to(dt)
        sh = torch.randn(C, device="cuda")
        us = timeit(lambda i: ops.pw_gemm(wt, xs[i % 3], None, None, T, sh, lens, False, True, None, None, None))
        print(f"pw_gemm C={C} {dt}: {us:.1f} us  {2*B*T*C*C/us/1e6:.0f} TFLOP/s")

print("# fp16 kernels on bf16-representable values (data-dependent power check)")
for C, K in ((512, 75),):
    w = (torch.randn(C, K, device="cuda") * 0.1).bfloat16().float()
    xs = [torch.randn(B, C, P, device="cuda").bfloat16().to(torch.float16) for _ in range(3)]
    for x in xs: x[:, :, T:] = 0
    us = timeit(lambda i: ops.dw_conv(xs[i % 3], T, w, 1, 1, K // 2, lens, True))
    print(f"dw_conv C={C} K={K} fp16 rows, bf16-representable data: {us:.1f} us")
    wt = (torch.randn(C, C, device="cuda") * 0.05).bfloat16().to(torch.float16)
    sh = torch.randn(C, device="cuda")
    us = timeit(lambda i: ops.pw_gemm(wt, xs[i % 3], None, None, T, sh, lens, False, True, None, None, None))
    print(f"pw_gemm C={C} fp16 rows, bf16-representable data: {us:.1f} us")
    xz = [torch.zeros(B, C, P, device="cuda", dtype=torch.float16) for _ in range(3)]
    us = timeit(lambda i: ops.dw_conv(xz[i % 3], T, w, 1, 1, K // 2, lens, True))
    print(f"dw_conv fp16 rows, all-zero input: {us:.1f} us")
    xz = [torch.zeros(B, C, P, device="cuda", dtype=torch.bfloat16) for _ in range(3)]
    us = timeit(lambda i: ops.dw_conv(xz[i % 3], T, w, 1, 1, K // 2, lens, True))
    print(f"dw_conv bf16 rows, all-zero input: {us:.1f} us")
