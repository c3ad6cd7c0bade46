import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import ref_numpy as R
from thunder_speech_b200 import synth, _lib
from thunder_speech_b200.quartznet.transform import FilterbankFeatures

x = synth.audio(3, 4800, 11, "noise")
lens = synth.ragged_lengths(3, 4800, 111)
print("lens", lens)
rf, rl, inter = R.filterbank_features(x, lens, return_intermediate=True)
fb = FilterbankFeatures().eval().cuda()
f, fl = fb(torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda())
torch.cuda.synchronize()
f = f.cpu().numpy()
print("fl", fl.cpu().numpy(), rl)
d = np.abs(f - rf)
for b in range(3):
    print("b", b, "max err", d[b].max(), "per-frame", np.round(d[b].max(0), 2))
print("got[1,0,:8]", f[1, 0, :8]); print("ref[1,0,:8]", rf[1, 0, :8])
print("got[0,0,:8]", f[0, 0, :8]); print("ref[0,0,:8]", rf[0, 0, :8])
print("got mean/std b0", f[0].mean(-1)[:4], f[0].std(-1)[:4])
