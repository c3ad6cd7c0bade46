import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from thunder_speech_b200 import ops
from thunder_speech_b200.train import dw_wgrad
B, T, C, K = 32, 751, 512, 75
pitch = ops.row_pitch(T)
da = torch.randn(B, C, pitch, device="cuda").to(torch.bfloat16); x = torch.randn(B, C, pitch, device="cuda").to(torch.bfloat16)
da[:, :, T:] = 0; x[:, :, T:] = 0
for _ in range(3):
    dw_wgrad(da, T, x, T, None, K, 1, 1, K // 2, premasked=True)
torch.cuda.synchronize()
