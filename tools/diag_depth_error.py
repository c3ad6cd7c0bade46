"""Per-block error growth of the bf16 device path against the fp32 CPU oracle on the NAMED architectures
(QuartzNet 15x5, Citrinet-1024).  Also runs the oracle with bf16 STORAGE simulated 
(oracle.ref_torch.block_storage) so the table separates "what bf16 rows cost by construction" from "what the kernels add".

    python tools/diag_depth_error.py [quartznet15x5|citrinet1024|quartznet5x5] [B] [seconds]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_numpy as R  # noqa: E402
from oracle import ref_torch as RT  # noqa: E402
from thunder_speech_b200 import ops, synth  # noqa: E402
from thunder_speech_b200.runner import build_model  # noqa: E402


def rel(a, b):
    a = a.double()
    b = b.double()
    d = a - b
    return float(d.abs().max() / b.abs().max()), float(d.norm() / b.norm())


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "quartznet15x5"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    secs = float(sys.argv[3]) if len(sys.argv) > 3 else 15
    kind = sys.argv[4] if len(sys.argv) > 4 else "tones"
    prec = sys.argv[5] if len(sys.argv) > 5 else "bf16"
    seed = 3
    N = int(secs * 16000)
    dev = torch.device("cuda")
    m = build_model(name, dev, seed=seed).set_precision(prec)
    f16 = prec == "fp16"
    print(f"# {name} B={B} x {secs} s ({kind}), rows {prec}")
    x = synth.audio(B, N, 77, kind)
    lens = np.sort(np.random.default_rng(5).integers(N // 2, N + 1, B))[::-1].astype(np.int64).copy()
    lens[0] = N
    for b in range(B):
        x[b, lens[b]:] = 0

    if name.startswith("quartznet"):
        rep = 1 if name == "quartznet5x5" else 3
        cfgs = R.quartznet_cfgs(repeat_blocks=rep)
        st = RT.to_torch(synth.encoder_state(synth.quartznet_block_list(repeat_blocks=rep), seed=seed))
        dec = RT.to_torch(synth.decoder_state(1024, 29, seed + 1))
        nfilt = 64
    else:
        c = synth.CITRINET_1024
        cfgs = R.citrinet_cfgs(c["filters"], c["kernel_sizes"], c["strides"], feat_in=80)
        st = RT.to_torch(synth.encoder_state(synth.citrinet_block_list(c["filters"], c["kernel_sizes"], c["strides"], 80),
                                             seed=seed, se=True))
        dec = RT.to_torch(synth.decoder_state(640, 1025, seed + 1))
        nfilt = 80

    torch.set_num_threads(os.cpu_count())
    xt, lt = torch.from_numpy(x), torch.from_numpy(lens)
    with torch.no_grad():
        f, fl = RT.features(xt, lt, nfilt=nfilt)
        # device
        hop = 160
        F = 1 + N // hop
        feats, feat_len = m.audio_transform.features(xt.to(dev), lt.to(dev), bf16_pitch=ops.row_pitch(F), f16=f16)
        l32 = feat_len.to(torch.int32)
        rows, T, lens_d = feats, F, l32
        print("features (16-bit rows) vs oracle:", rel(ops.unpack_rows(rows, T).cpu(), f))
        e32, l_32 = f, fl
        e16, l_16 = RT.rounder(prec)(f), fl
        blocks = list(m.encoder.children())
        print(f"{'blk':>3} {'C':>5} {'T':>5}  dev-vs-fp32(max,l2)      16b-store-vs-fp32        dev-vs-16b-store         teacher-forced dev-vs-fp32")
        for i, (blk, cfg) in enumerate(zip(blocks, cfgs)):
            rows, T, lens_d = blk.forward_rows(rows, T, lens_d, zero_tail=(i != len(blocks) - 1))
            prev32 = (e32, l_32)
            e32, l_32 = RT.block(e32, l_32, cfg, st, f"{i}.")
            e16, l_16 = RT.block_storage(e16, l_16, cfg, st, f"{i}.", prec)
            # teacher forced: the device block fed with the ORACLE's input (rounded to bf16 rows)
            tin, tl = (f, fl) if i == 0 else prev32
            trow = ops.pack_rows(tin.to(dev), tl.to(dev).to(torch.int32), f16)
            tr, tT, _ = blk.forward_rows(trow, tin.shape[-1], tl.to(dev).to(torch.int32), zero_tail=True)
            tf = ops.unpack_rows(tr, tT).cpu()
            d = ops.unpack_rows(rows, T).cpu()
            # compare on valid frames only (the tail is zeroed on the device between blocks)
            mask = (torch.arange(d.shape[-1])[None, :] < l_32[:, None]).unsqueeze(1)
            a, b16, r = d * mask, e16 * mask, e32 * mask
            ea, eb, ec, et = rel(a, r), rel(b16, r), rel(a, b16), rel(tf * mask, r)
            print(f"{i:3d} {d.shape[1]:5d} {T:5d}  {ea[0]:.3e} {ea[1]:.3e}   {eb[0]:.3e} {eb[1]:.3e}   {ec[0]:.3e} {ec[1]:.3e}   "
                  f"{et[0]:.3e} {et[1]:.3e}")
        lg32 = torch.nn.functional.conv1d(e32, dec["weight"], dec["bias"])
        lg16 = torch.nn.functional.conv1d(e16, RT.rounder(prec)(dec["weight"]), dec["bias"])
        logits, out_len = m(xt.to(dev), lt.to(dev))
        lg = logits.cpu()
        print("logits dev-vs-fp32:", rel(lg, lg32), " 16b-store-vs-fp32:", rel(lg16, lg32))
        ids_d, ids_r = lg.argmax(1), lg32.argmax(1)
        print("argmax agreement dev vs fp32:", float((ids_d == ids_r).float().mean()),
              " 16b-store vs fp32:", float((lg16.argmax(1) == ids_r).float().mean()))
        assert torch.equal(out_len.cpu(), l_32)


if __name__ == "__main__":
    main()
