import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.make_golden_train import TRAIN_BLOCK_CASES, block_case
from thunder_speech_b200 import ops
from thunder_speech_b200.quartznet.blocks import QuartznetBlock
from thunder_speech_b200.train import BlockTrainer
g = np.load("tests/golden/train.npz")
def l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.sqrt(((a-b)**2).sum()) / max(np.sqrt((b*b).sum()), 1e-30))
for ci in range(len(TRAIN_BLOCK_CASES)):
    name, cfg, st, x, lens, Rm = block_case(ci)
    blk = QuartznetBlock(cfg["in_channels"], cfg["out_channels"], repeat=cfg["repeat"], kernel_size=(cfg["kernel_size"],),
                         stride=(cfg["stride"],), dilation=(cfg["dilation"],), residual=cfg["residual"], separable=cfg["separable"])
    blk.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}, strict=True)
    blk = blk.cuda().train(); bt = BlockTrainer(blk)
    l32 = torch.from_numpy(lens.astype(np.int32)).cuda()
    rows = ops.pack_rows(torch.from_numpy(x).cuda(), l32)
    y, T_out, lo, tape = bt.forward(rows, x.shape[-1], l32, zero_tail=False)
    print(name, "out", round(l2(ops.unpack_rows(y, T_out).cpu().numpy(), g[f"{name}.out"]), 4))
    need_dx = cfg["stride"] == 1
    dx = bt.backward(tape, ops.pack_rows(torch.from_numpy(Rm).cuda()), need_dx=need_dx)
    if need_dx: print("   dx", round(l2(ops.unpack_rows(dx, x.shape[-1]).cpu().numpy(), g[f"{name}.dx"]), 4))
    for k, p in blk.named_parameters():
        print("   ", k, round(l2(p.grad.cpu().numpy(), g[f"{name}.grad.{k}"]), 4))
