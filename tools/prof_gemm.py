import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from thunder_speech_b200 import ops, _lib
bn = int(sys.argv[1]); B, Cin, Cout, T = 256, 512, 512, 751
_lib.set_option("pw_bn", bn)
dev = torch.device("cuda"); pitch = ops.row_pitch(T)
x = torch.randn(B, Cin, pitch, device=dev).bfloat16(); w = (torch.randn(Cout, Cin, device=dev) / Cin ** 0.5).bfloat16()
shift = torch.randn(Cout, device=dev)
for _ in range(3):
    ops.pw_gemm(w, x, None, None, T, shift, None, False, True, None, None, None)
torch.cuda.synchronize()
