"""Launch the two hot kernels once per variant for ncu (--set full):  python tools/prof_hot.py [B]
order of profiled launches: pair GEMM 512x512 streaming, weight-stationary; Toeplitz conv C=512 K=75 per-channel, persistent."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from thunder_speech_b200 import ops, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
Cin = Cout = 512; T = 751; K = 75
dev = torch.device("cuda"); pitch = ops.row_pitch(T)
xs = [torch.randn(B, Cin, pitch, device=dev).bfloat16() for _ in range(2)]
for x in xs: x[:, :, T:] = 0
w = (torch.randn(Cout, Cin, device=dev) / Cin ** 0.5).bfloat16()
shift = torch.randn(Cout, device=dev)
wd = torch.randn(Cin, K, device=dev) * 0.1
lens = torch.full((B,), T, dtype=torch.int32, device=dev)
_lib.set_option("serpentine", 0)
def warm():
    for i in range(2):
        ops.pw_gemm(w, xs[i], None, None, T, shift, lens, False, True, None, None, None)
        ops.dw_conv(xs[i], T, wd, 1, 1, K // 2, lens, True)
    torch.cuda.synchronize()
for res in (0, 1):
    _lib.set_option("pw_resident", res); warm()
    torch.cuda.profiler.start()
    ops.pw_gemm(w, xs[0], None, None, T, shift, lens, False, True, None, None, None)
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
for per in (0, 1):
    _lib.set_option("dw_persist", per); warm()
    torch.cuda.profiler.start()
    ops.dw_conv(xs[0], T, wd, 1, 1, K // 2, lens, True)
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
