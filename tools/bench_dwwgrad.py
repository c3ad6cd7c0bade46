import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from thunder_speech_b200 import _lib, ops
B, T = 32, 751
def run(C, K, D, bchunk, iters=50, flags=0):
    P = D * (K - 1) // 2
    pitch = ops.row_pitch(T)
    da = torch.randn(B, C, pitch, device="cuda").to(torch.bfloat16); x = torch.randn(B, C, pitch, device="cuda").to(torch.bfloat16)
    da[:, :, T:] = 0; x[:, :, T:] = 0
    nchunk = (B + bchunk - 1) // bchunk
    part = torch.empty(nchunk, C, K, device="cuda")
    f = lambda: _lib.check(_lib.lib().ts_dw_wgrad(da.data_ptr(), T, pitch, x.data_ptr(), T, pitch, None, B, C, K, 1, D, P, bchunk, flags, part.data_ptr(), torch.cuda.current_stream().cuda_stream), "x")
    for _ in range(5): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(f"C {C} K {K} D {D} R {bchunk}: {us:7.1f} us  {B*C*T*K/us/1e6:6.2f} TFMA/s  {2*B*C*pitch*2/us/1e3:6.0f} GB/s")
for R in (8,):
    run(256, 33, 1, R); run(512, 75, 1, R)
run(512, 87, 2, 8)
for (C, K, D) in [(256, 33, 1), (512, 75, 1), (512, 87, 2)]:
    run(C, K, D, B, flags=1)
