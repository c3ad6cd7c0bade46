"""Correctness + speed of the CTA-pair GEMM against the single-CTA persistent kernel (same inputs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from thunder_speech_b200 import ops, _lib
dev = torch.device("cuda")
torch.manual_seed(0)
def run(B, Cin, Cout, T, res, iters=5):
    pitch = ops.row_pitch(T)
    x = torch.randn(B, Cin, pitch, device=dev).bfloat16(); x[:, :, T:] = 0
    w = (torch.randn(Cout, Cin, device=dev) / Cin ** 0.5).bfloat16()
    w1 = (torch.randn(Cout, res, device=dev) / max(res, 1) ** 0.5).bfloat16() if res else None
    x1 = torch.randn(B, res, pitch, device=dev).bfloat16() if res else None
    shift = torch.randn(Cout, device=dev)
    lens = torch.randint(T // 2, T + 1, (B,), device=dev, dtype=torch.int32)
    outs = []
    for mode in (0, 2):
        _lib.set_option("pw_pair", mode)
        y = ops.pw_gemm(w, x, w1, x1, T, shift, lens, False, True, None, None, None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(int(2e6)); e0.record()
        for _ in range(iters):
            ops.pw_gemm(w, x, w1, x1, T, shift, lens, False, True, None, None, None)
        e1.record(); torch.cuda.synchronize()
        outs.append((y.float(), e0.elapsed_time(e1) / iters))
    d = (outs[0][0] - outs[1][0]).abs().max().item()
    fl = 2 * B * T * (Cin + res) * Cout / 1e12
    print(f"B={B} {Cin}(+{res})->{Cout} T={T}: maxdiff {d:.3e}  single {outs[0][1]*1e3:7.1f} us ({fl/outs[0][1]*1e3:5.0f} TF/s)  pair {outs[1][1]*1e3:7.1f} us ({fl/outs[1][1]*1e3:5.0f} TF/s)", flush=True)
for cfg in [(2, 64, 256, 100, 0), (4, 256, 256, 751, 0), (16, 512, 512, 751, 512), (128, 1024, 1024, 1001, 0), (256, 512, 512, 751, 0), (128, 1024, 640, 251, 0)]:
    run(*cfg)
