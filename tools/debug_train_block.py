import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import ref_numpy as R, ref_torch as RT
from thunder_speech_b200 import ops, synth
from thunder_speech_b200.quartznet.blocks import QuartznetBlock
from thunder_speech_b200.train import BlockTrainer
def l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.sqrt(((a-b)**2).sum()) / max(np.sqrt((b*b).sum()), 1e-30))
def run(cin, cout, K, rep, B, T, res, zero_tail, seed=0, full=False):
    rng = np.random.Generator(np.random.PCG64(seed))
    st = synth.block_state(rng, "", cin, cout, rep, K, res, True)
    x = np.maximum(rng.standard_normal((B, cin, T)), 0).astype(np.float32)
    lens = np.sort(rng.integers(T // 2, T + 1, B))[::-1].astype(np.int64).copy(); lens[0] = T
    if full: lens[:] = T
    Rm = rng.standard_normal((B, cout, T)).astype(np.float32)
    m = (np.arange(T)[None, :] < lens[:, None])[:, None, :]
    if zero_tail: Rm = np.where(m, Rm, 0).astype(np.float32)
    xm = np.where(m, x, 0).astype(np.float32)
    cfg = R.BlockCfg(cin, cout, repeat=rep, kernel_size=K, residual=res, separable=True)
    stt = {k: torch.from_numpy(np.asarray(v)).clone() for k, v in st.items()}
    for k, v in stt.items():
        if v.dtype.is_floating_point and "running" not in k: v.requires_grad_(True)
    xt = torch.from_numpy(xm).requires_grad_(True)
    y, yl = RT.block(xt, torch.from_numpy(lens), cfg, stt, "", train=True, store=RT.bf16_store if os.environ.get("SIM") else None)
    (y * torch.from_numpy(Rm)).sum().backward()
    blk = QuartznetBlock(cin, cout, repeat=rep, kernel_size=(K,), residual=res, separable=True)
    blk.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}, strict=True)
    blk = blk.cuda().train(); bt = BlockTrainer(blk)
    l32 = torch.from_numpy(lens.astype(np.int32)).cuda()
    rows = ops.pack_rows(torch.from_numpy(x).cuda(), l32)
    yy, T_out, lo, tape = bt.forward(rows, T, l32, zero_tail=zero_tail)
    out = ops.unpack_rows(yy, T_out).cpu().numpy()
    yref = y.detach().numpy()
    if zero_tail: yref = np.where(m, yref, 0)
    dx = bt.backward(tape, ops.pack_rows(torch.from_numpy(Rm).cuda()), need_dx=True)
    errs = {k: l2(p.grad.cpu().numpy(), stt[k].grad.numpy()) for k, p in blk.named_parameters()}
    print(f"cin {cin} cout {cout} K {K} rep {rep} B {B} T {T} res {res} zt {zero_tail} full {full}: out {l2(out, yref):.4f} dx {l2(ops.unpack_rows(dx, T).cpu().numpy(), xt.grad.numpy()):.4f} max param {max(errs.values()):.4f} ({max(errs, key=errs.get)}) mean {np.mean(list(errs.values())):.4f}")
run(16, 24, 5, 3, 3, 50, True, False)
run(32, 32, 5, 5, 8, 101, True, False)
run(32, 32, 5, 5, 8, 101, True, True)
run(32, 32, 5, 5, 8, 101, True, True, full=True)
run(256, 32, 5, 5, 8, 101, True, True)
run(32, 32, 13, 5, 8, 101, True, True)
run(32, 32, 5, 1, 8, 101, False, True)
run(32, 32, 5, 2, 8, 101, False, True)
run(256, 256, 33, 5, 8, 751, True, True)
