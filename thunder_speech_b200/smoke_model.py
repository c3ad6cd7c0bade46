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


def run_train(dev):
    """One tiny training step (QuartzNet 5x5, 4 x 1 s): CTC loss checked against the autograd torch port on the CPU."""
    from oracle import ref_numpy as R
    from oracle import ref_torch as RT
    from thunder_speech_b200 import synth
    from thunder_speech_b200.runner import build_model
    from thunder_speech_b200.train import CTCTrainStep

    m = build_model("quartznet5x5", dev, seed=5)
    m.encoder.train()
    m.decoder.train()
    x = synth.audio(4, 16000, 22, "noise")
    lens = np.array([16000, 16000, 12000, 9000], np.int64)
    rng = np.random.default_rng(3)
    y = rng.integers(0, 28, (4, 6)).astype(np.int64)
    yl = np.array([6, 5, 4, 3], np.int64)
    step = CTCTrainStep(m, lr=1e-3, use_graph=False)
    loss = step.step(torch.from_numpy(x).to(dev), torch.from_numpy(lens).to(dev), torch.from_numpy(y).to(dev),
                     torch.from_numpy(yl).to(dev))
    torch.cuda.synchronize()
    cfgs = R.quartznet_cfgs(repeat_blocks=1)
    st = RT.to_torch(synth.encoder_state(synth.quartznet_block_list(repeat_blocks=1), seed=5))
    dec = RT.to_torch(synth.decoder_state(1024, 29, seed=6))
    with torch.no_grad():
        f, fl = RT.features(torch.from_numpy(x), torch.from_numpy(lens))
        e, el = RT.encoder(f, fl, cfgs, st, train=True)
        ref = float(RT.ctc_loss(torch.nn.functional.conv1d(e, dec["weight"], dec["bias"]), torch.from_numpy(y), el,
                                torch.from_numpy(yl), 28))
    err = abs(float(loss) - ref) / abs(ref)
    assert err < 3e-2, f"training loss parity {err} ({float(loss)} vs {ref})"
    gn = float(step.flat.norm())
    assert np.isfinite(gn) and gn > 0
    print(f"smoke training step ok: loss {float(loss):.4f} (oracle {ref:.4f}), |grad| {gn:.3e}")
