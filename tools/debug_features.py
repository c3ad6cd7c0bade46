import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import ref_numpy as R
from thunder_speech_b200 import synth, _lib
from thunder_speech_b200.quartznet.transform import FilterbankFeatures

def logmel_cuda(x, generic=False):
    fb = FilterbankFeatures().eval().cuda()
    t = fb._device_tables(torch.device("cuda"))
    B, N = x.shape
    F = 1 + N // 160
    a = torch.from_numpy(x).cuda()
    out = torch.full((B, 64, F), -777.0, device="cuda")
    lo, hi = (0, 512) if generic else (t["win_lo"], t["win_hi"])
    L = _lib.lib()
    _lib.check(L.ts_logmel(a.data_ptr(), B, N, 512, 160, 0.97, t["window_full"].data_ptr(), lo, hi,
                           t["twiddle"].data_ptr(), t["mel_start"].data_ptr(), t["mel_count"].data_ptr(),
                           t["mel_off"].data_ptr(), t["mel_w"].data_ptr(), 64, t["mel_w"].numel(),
                           out.data_ptr(), torch.cuda.current_stream().cuda_stream), "x")
    torch.cuda.synchronize()
    return out.cpu().numpy()

x = synth.audio(2, 4800, 11, "noise")
_, _, inter = R.filterbank_features(x, np.array([4800, 4800]), return_intermediate=True)
ref = inter["logmel"]
for generic in (False, True):
    got = logmel_cuda(x, generic)
    d = np.abs(got - ref)
    print("generic", generic, "max abs err", d.max(), "ref range", ref.min(), ref.max())
    print(" per-frame max err:", np.round(d[0].max(0), 3))
    print(" per-mel max err:", np.round(d[0].max(1), 3)[:16])
    print(" got[0,:4,:6]", got[0, :4, :6]); print(" ref[0,:4,:6]", ref[0, :4, :6])
