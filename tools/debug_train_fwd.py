import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from oracle import ref_torch as RT
from thunder_speech_b200 import ops
from thunder_speech_b200.train import row_stats
q = RT.bf16_store
def l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.sqrt(((a-b)**2).sum()) / max(np.sqrt((b*b).sum()), 1e-30))
for (C, K, B, T) in [(32, 13, 8, 101), (256, 33, 8, 751), (64, 33, 4, 200)]:
    rng = np.random.Generator(np.random.PCG64(1))
    x = np.maximum(rng.standard_normal((B, C, T)), 0).astype(np.float32)
    lens = np.sort(rng.integers(T // 2, T + 1, B))[::-1].astype(np.int64).copy(); lens[0] = T
    w = (rng.standard_normal((C, K)) / np.sqrt(K)).astype(np.float32)
    wp = (rng.standard_normal((C, C)) / np.sqrt(C)).astype(np.float32)
    m = (np.arange(T)[None, :] < lens[:, None])[:, None, :]
    xt = q(torch.from_numpy(np.where(m, x, 0).astype(np.float32)))
    a_ref = q(F.conv1d(xt, q(torch.from_numpy(w))[:, None, :], None, 1, K // 2, 1, groups=C))
    a_ref = a_ref * torch.from_numpy(m)
    z_ref = q(F.conv1d(a_ref, q(torch.from_numpy(wp))[:, :, None]))
    l32 = torch.from_numpy(lens.astype(np.int32)).cuda()
    rows = ops.pack_rows(torch.from_numpy(x).cuda(), l32)
    a = ops.dw_conv(rows, T, torch.from_numpy(w).cuda(), 1, 1, K // 2, l32, True)
    z = ops.pw_gemm(torch.from_numpy(wp).cuda().to(torch.bfloat16), a, None, None, T, None, None, False, False, None, None, None)
    au = ops.unpack_rows(a, T).cpu().numpy(); zu = ops.unpack_rows(z, T).cpu().numpy()
    print(C, K, "a", l2(au, a_ref.numpy()), "z", l2(zu, z_ref.numpy()), "a tail max", np.abs(np.where(m, 0, au)).max(), "z tail max", np.abs(np.where(m, 0, zu)).max())
    d = np.abs(au - a_ref.numpy()); idx = np.argwhere(d > 0.02)
    print(" big a diffs", len(idx), idx[:8].tolist(), "lens", lens.tolist())
    st = row_stats(z, T).cpu().numpy()
    print(" stats", l2(st[:, 0], z_ref.numpy().sum((0, 2))), l2(st[:, 1], (z_ref.numpy().astype(np.float64) ** 2).sum((0, 2))))
