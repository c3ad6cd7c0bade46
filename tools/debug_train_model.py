import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import ref_numpy as R, ref_torch as RT
from thunder_speech_b200 import synth
from thunder_speech_b200.blocks import conv1d_decoder
from thunder_speech_b200.module import CTCModule
from thunder_speech_b200.quartznet.blocks import QuartznetEncoder
from thunder_speech_b200.quartznet.transform import FilterbankFeatures
from thunder_speech_b200.text_processing import BatchTextTransformer
from thunder_speech_b200.train import CTCTrainStep
filters, kernels = [32, 32, 32, 32, 32], [5, 7, 9, 11, 13]
st = synth.encoder_state(synth.quartznet_block_list(filters=filters, kernel_sizes=kernels, repeat_blocks=1), seed=31)
dec = synth.decoder_state(1024, 29, seed=32)
x = synth.audio(8, 32000, 33, "tones")
lens = np.array([32000, 32000, 30000, 28000, 25000, 22222, 20000, 16000], np.int64)
rng = np.random.default_rng(34)
y = rng.integers(0, 28, (8, 12)).astype(np.int64); y_len = rng.integers(4, 13, 8).astype(np.int64)
cfgs = R.quartznet_cfgs(filters=filters, kernel_sizes=kernels, repeat_blocks=1)
stt = {k: torch.from_numpy(np.asarray(v)).clone() for k, v in st.items()}
for k, v in stt.items():
    if v.dtype.is_floating_point and "running" not in k: v.requires_grad_(True)
dw_, db_ = torch.from_numpy(dec["weight"]).requires_grad_(True), torch.from_numpy(dec["bias"]).requires_grad_(True)
with torch.no_grad(): f, fl = RT.features(torch.from_numpy(x), torch.from_numpy(lens))
e, el = RT.encoder(f, fl, cfgs, stt, train=True, store=RT.bf16_store if os.environ.get('SIM') else None)
ref_loss = RT.ctc_loss(torch.nn.functional.conv1d(e, dw_, db_), torch.from_numpy(y), el, torch.from_numpy(y_len), 28)
ref_loss.backward()
ref = {k: v.grad.numpy() for k, v in stt.items() if v.requires_grad}
ref["decoder.weight"], ref["decoder.bias"] = dw_.grad.numpy(), db_.grad.numpy()
enc = QuartznetEncoder(filters=filters, kernel_sizes=kernels, repeat_blocks=1)
enc.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}, strict=True)
d = conv1d_decoder(1024, 29); d.load_state_dict({k: torch.from_numpy(v) for k, v in dec.items()})
m = CTCModule(enc, d, FilterbankFeatures(nfilt=64, dither=0.0), BatchTextTransformer(synth.quartznet_vocab())).cuda()
m.encoder.train(); m.decoder.train()
step = CTCTrainStep(m)
loss = step.loss_and_grads(torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(y_len).cuda())
print("loss", loss.item(), ref_loss.item())
got = {k: p.grad.float().cpu().numpy() for k, p in m.encoder.named_parameters()}
got["decoder.weight"], got["decoder.bias"] = m.decoder.weight.grad.cpu().numpy(), m.decoder.bias.grad.cpu().numpy()
for k, g in ref.items():
    a, b = got[k].ravel().astype(np.float64), g.ravel().astype(np.float64)
    print(f"{k:34s} cos {a@b/max(np.linalg.norm(a)*np.linalg.norm(b),1e-30):7.4f}  norm ours {np.linalg.norm(a):10.4f} ref {np.linalg.norm(b):10.4f}")
