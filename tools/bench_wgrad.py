import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from thunder_speech_b200 import _lib, ops
from thunder_speech_b200.train import pw_wgrad
B, T = 32, 751
def timeit(f, iters=20):
    """GPU time per call under CUDA-graph replay (no CPU launch overhead); inputs > L2 are the caller's business."""
    for _ in range(3): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for (Cout, Cin) in [(256, 256), (512, 256), (512, 512), (1024, 512), (64, 1024)]:
    pitch = ops.row_pitch(T)
    dz = torch.randn(B, Cout, pitch, device="cuda").to(torch.bfloat16); a = torch.randn(B, Cin, pitch, device="cuda").to(torch.bfloat16)
    out = torch.empty(Cout, Cin, device="cuda")
    us = timeit(lambda: pw_wgrad(dz, a, T, out=out))
    fl = 2 * B * T * Cout * Cin
    by = B * T * (Cout + Cin) * 2
    print(f"wgrad Cout {Cout} Cin {Cin}: {us:7.1f} us  {fl/us/1e6:7.1f} TF/s  min-bytes {by/us/1e3:6.0f} GB/s   bounds: mem {by/6.5e6:5.1f} us  mma {fl/1.4e9:5.1f} us")
    w = torch.randn(Cout, Cin, device="cuda").to(torch.bfloat16)
    us = timeit(lambda: ops.pw_gemm(w, a, None, None, T, None, None, False, False, None, None, None))
    print(f"  fwd gemm              : {us:7.1f} us  {fl/us/1e6:7.1f} TF/s")
