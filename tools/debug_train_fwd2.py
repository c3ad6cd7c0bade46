import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from oracle import ref_numpy as R, ref_torch as RT
from thunder_speech_b200 import ops, synth
from thunder_speech_b200.quartznet.blocks import QuartznetBlock
from thunder_speech_b200.train import BlockTrainer
q = RT.bf16_store
def l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.sqrt(((a-b)**2).sum()) / max(np.sqrt((b*b).sum()), 1e-30))
cin, cout, K, rep, B, T, res = 32, 32, 5, 5, 8, 101, True
if len(sys.argv) > 1: cin, cout, K, rep, B, T = map(int, sys.argv[1:7])
rng = np.random.Generator(np.random.PCG64(0))
st = synth.block_state(rng, "", cin, cout, rep, K, res, True)
x = np.maximum(rng.standard_normal((B, cin, T)), 0).astype(np.float32)
lens = np.sort(rng.integers(T // 2, T + 1, B))[::-1].astype(np.int64).copy(); lens[0] = T
m = torch.from_numpy((np.arange(T)[None, :] < lens[:, None])[:, None, :])
blk = QuartznetBlock(cin, cout, repeat=rep, kernel_size=(K,), residual=res, separable=True)
blk.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}, strict=True)
blk = blk.cuda().train(); bt = BlockTrainer(blk)
l32 = torch.from_numpy(lens.astype(np.int32)).cuda()
rows = ops.pack_rows(torch.from_numpy(x).cuda(), l32)
yy, T_out, lo, tape = bt.forward(rows, T, l32, zero_tail=True)
S = {k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}
cur = q(torch.from_numpy(x)) * m
x0 = cur
for r in range(rep):
    i = 5 * r
    a = q(F.conv1d(cur, q(S[f"mconv.{i}.conv.weight"]), None, 1, K // 2, 1, groups=cur.shape[1])) * m
    z = q(F.conv1d(a, q(S[f"mconv.{i+1}.conv.weight"])))
    p = f"mconv.{i+2}.layer.0"
    mean = z.mean((0, 2)); var = z.var((0, 2), unbiased=False)
    yb = F.batch_norm(z, None, None, S[p + ".weight"], S[p + ".bias"], True, 0.1, 1e-3)
    rec = tape["subs"][r]
    print(r, "a", l2(ops.unpack_rows(rec["a"], T).cpu(), a), "z", l2(ops.unpack_rows(rec["z"], T).cpu(), z),
          "mean", l2(rec["mean"].cpu(), mean), "inv", l2(rec["inv"].cpu(), 1 / torch.sqrt(var + 1e-3)))
    if r == rep - 1 and res:
        zr = q(F.conv1d(x0, q(S["res.0.conv.weight"])))
        yb = yb + F.batch_norm(zr, None, None, S["res.1.layer.0.weight"], S["res.1.layer.0.bias"], True, 0.1, 1e-3)
        print("  zr", l2(ops.unpack_rows(rec["zr"], T).cpu(), zr))
    y = q(F.relu(yb)) * m
    d = (ops.unpack_rows(rec["y"], T).cpu() - y).abs()
    print("   y", l2(ops.unpack_rows(rec["y"], T).cpu(), y), "n diff", int((d > 0).sum()), "max", float(d.max()), "at", np.unravel_index(int(d.argmax()), d.shape), "yref there", float(y.flatten()[int(d.argmax())]))
    cur = y
