"""Drives the UNMODIFIED reference (baseline/_ref/thunder, see install_ref.sh) for ``bench.py --impl reference`` (CPU, all
host threads) and ``--impl reference_cuda`` (the same modules ``.cuda()``: eager fp32 and autocast bf16, the "second
informative baseline" of BASELINE.md).  BENCH INFRASTRUCTURE ONLY: none of this repo's kernels, models or engine are on
this path -- only ``thunder_speech_b200.synth`` (numpy) is used to generate the same synthetic audio and random-init
weights the B200 arm uses, loaded through the reference's own strict ``load_state_dict``.

``thunder.module`` itself needs pytorch_lightning / torchmetrics, which are not in this image, so ``forward`` /
``predict`` are the three statements of ``BaseCTCModule`` (src/thunder/module.py:84-86, 98-100) applied to the
reference's own ``FilterbankFeatures`` / encoder / ``conv1d_decoder`` / ``BatchTextTransformer`` objects.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "thunder"))


def _import():
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    warnings.filterwarnings("ignore")
    import thunder  # noqa: F401  (its __init__ resolves the version from the dist-info next to it)
    from thunder.blocks import conv1d_decoder
    from thunder.citrinet.blocks import CitrinetEncoder
    from thunder.quartznet.blocks import QuartznetEncoder
    from thunder.quartznet.transform import FilterbankFeatures
    from thunder.text_processing.transform import BatchTextTransformer

    return conv1d_decoder, CitrinetEncoder, QuartznetEncoder, FilterbankFeatures, BatchTextTransformer


class ReferenceModel:
    """audio_transform / encoder / decoder / text_transform of ``BaseCTCModule`` (module.py:53-56), eval mode."""

    def __init__(self, name: str, device="cpu", seed: int = 0):
        from thunder_speech_b200 import synth

        conv1d_decoder, CitrinetEncoder, QuartznetEncoder, FilterbankFeatures, BatchTextTransformer = _import()
        if name in ("quartznet5x5", "quartznet15x5"):
            rep = 1 if name == "quartznet5x5" else 3
            enc = QuartznetEncoder(repeat_blocks=rep)
            st = synth.encoder_state(synth.quartznet_block_list(repeat_blocks=rep), seed=seed)
            dec, dst = conv1d_decoder(1024, 29), synth.decoder_state(1024, 29, seed + 1)
            fb, tt = FilterbankFeatures(nfilt=64), BatchTextTransformer(tokens=synth.quartznet_vocab())
        elif name == "citrinet1024":
            c = synth.CITRINET_1024
            enc = CitrinetEncoder(c["filters"], c["kernel_sizes"], c["strides"], feat_in=80)
            st = synth.encoder_state(synth.citrinet_block_list(c["filters"], c["kernel_sizes"], c["strides"], 80),
                                     seed=seed, se=True)
            dec, dst = conv1d_decoder(640, 1025), synth.decoder_state(640, 1025, seed + 1)
            fb, tt = FilterbankFeatures(nfilt=80), BatchTextTransformer(tokens=synth.citrinet_vocab(1024))
        elif name == "features":
            enc = dec = tt = None
            st = dst = {}
            fb = FilterbankFeatures(nfilt=64)
        else:
            raise ValueError(name)
        if enc is not None:
            enc.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}, strict=True)
            dec.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in dst.items()}, strict=True)
            self.encoder, self.decoder = enc.eval().to(device), dec.eval().to(device)
        self.audio_transform = fb.eval().to(device)
        self.text_transform = tt
        self.name = name

    @torch.no_grad()
    def forward(self, x, lengths):
        features, feature_lengths = self.audio_transform(x, lengths)            # module.py:84
        if self.name == "features":
            return features, feature_lengths
        encoded, out_lengths = self.encoder(features, feature_lengths)          # module.py:85
        return self.decoder(encoded), out_lengths                               # module.py:86

    @torch.no_grad()
    def predict(self, x):
        audio_lengths = torch.tensor(x.shape[0] * [x.shape[-1]], device=x.device)   # module.py:98
        pred, _ = self.forward(x, audio_lengths)                                     # module.py:99
        if self.name == "features":
            return pred
        return self.text_transform.decode_prediction(pred.argmax(1))                 # module.py:100
