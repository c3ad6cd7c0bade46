"""Second half of ``__graft_entry__.smoke()``: one small QuartzNet 5x5 ``predict`` on the GPU, logits checked
against the numpy oracle (imports ``oracle`` -- only ever called from smoke())."""
import numpy as np
import torch


def run(dev):
    from oracle import ref_numpy as R
    from thunder_speech_b200 import synth
    from thunder_speech_b200.runner import build_model

    m = build_model("quartznet5x5", dev, seed=5)
    x = synth.audio(2, 8000, 21, "tones")
    logits, out_len = m(torch.from_numpy(x).to(dev), torch.tensor([8000, 8000], device=dev))
    texts = m.predict(torch.from_numpy(x).to(dev))
    torch.cuda.synchronize()
    cfgs = R.quartznet_cfgs(repeat_blocks=1)
    st = synth.encoder_state(synth.quartznet_block_list(repeat_blocks=1), seed=5)
    dec = synth.decoder_state(1024, 29, seed=6)
    f, fl = R.filterbank_features(x, np.array([8000, 8000]))
    e, el = R.encoder_forward(f, fl, cfgs, st)
    ref = R.decoder_forward(e, dec["weight"], dec["bias"])
    err = np.abs(logits.cpu().numpy() - ref).max() / np.abs(ref).max()
    assert (out_len.cpu().numpy() == el).all()
    assert err < 2e-2, f"logit parity {err}"
    print(f"smoke quartznet5x5 ok: logits max rel err {err:.2e}; transcripts {texts}")
