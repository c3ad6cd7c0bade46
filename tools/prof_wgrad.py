import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from thunder_speech_b200 import ops
from thunder_speech_b200.train import pw_wgrad, dw_wgrad
B, T = 32, 751
pitch = ops.row_pitch(T)
for (Cout, Cin) in [(256, 256), (512, 512)]:
    dz = torch.randn(B, Cout, pitch, device="cuda").to(torch.bfloat16); a = torch.randn(B, Cin, pitch, device="cuda").to(torch.bfloat16)
    out = torch.empty(Cout, Cin, device="cuda"); w = torch.randn(Cout, Cin, device="cuda").to(torch.bfloat16)
    for _ in range(3):
        pw_wgrad(dz, a, T, out=out)
        ops.pw_gemm(w, a, None, None, T, None, None, False, False, None, None, None)
        dw_wgrad(dz, T, dz, T, None, 33, 1, 1, 16)
        ops.dw_conv(a, T, torch.randn(Cin, 33, device="cuda"), 1, 1, 16, None, True)
torch.cuda.synchronize()
