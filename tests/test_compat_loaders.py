"""CPU tests of the NeMo-checkpoint loaders (reference: src/thunder/quartznet/compatibility.py,
src/thunder/citrinet/compatibility.py, tests/quartznet/test_compatibility_qn.py).  A synthetic `.nemo` archive is built with
NeMo's parameter naming and must load strictly into the module shells."""
import io
import os
import tarfile
from pathlib import Path

import numpy as np
import pytest
import torch
import yaml

from thunder_speech_b200 import synth
from thunder_speech_b200.citrinet.compatibility import fix_vocab, load_citrinet_checkpoint, load_components_from_citrinet_config
from thunder_speech_b200.compat import extract_nemo, fix_encoder_name, load_yaml_config
from thunder_speech_b200.quartznet.blocks import QuartznetEncoder
from thunder_speech_b200.quartznet.compatibility import load_components_from_quartznet_config, load_quartznet_checkpoint


def to_nemo_name(k: str) -> str:
    """Inverse of fix_encoder_name: this package's (= the reference's) key -> NeMo's key."""
    k = k.replace(".layer.0.", ".")
    parts = k.split(".")
    if parts[1] == "res":                       # 1.res.0.conv.weight -> 1.res.0.0.conv.weight
        parts = parts[:2] + ["0"] + parts[2:]
    return "encoder.encoder." + ".".join(parts)


def jasper_cfg(filters, kernels, strides=None, citrinet=False):
    strides = strides or [1] * len(filters)
    body = [{"filters": f, "repeat": 5, "kernel": [k], "stride": [s], "dilation": [1], "dropout": 0.0, "residual": True,
             "separable": True} for f, k, s in zip(filters, kernels, strides)]
    pro = [{"filters": 256, "repeat": 1, "kernel": [33], "stride": [2], "dilation": [1], "dropout": 0.0, "residual": False}]
    epi = [{"filters": 512, "repeat": 1, "kernel": [87], "stride": [1], "dilation": [2], "dropout": 0.0, "residual": False},
           {"filters": 1024, "repeat": 1, "kernel": [1], "stride": [1], "dilation": [1], "dropout": 0.0, "residual": False}]
    return pro + body + (epi[:1] if citrinet else epi)


def write_nemo(path, config, state):
    with tarfile.open(path, "w:gz") as tar:
        for name, data in (("model_config.yaml", yaml.safe_dump(config).encode()), ("model_weights.ckpt", state)):
            if not isinstance(data, bytes):
                buf = io.BytesIO()
                torch.save(data, buf)
                data = buf.getvalue()
            info = tarfile.TarInfo(name)
            info.size = len(data)
            tar.addfile(info, io.BytesIO(data))


def test_fix_encoder_name_rules():
    assert fix_encoder_name("encoder.encoder.0.mconv.0.conv.weight") == "0.mconv.0.conv.weight"
    assert fix_encoder_name("encoder.encoder.0.mconv.2.weight") == "0.mconv.2.layer.0.weight"
    assert fix_encoder_name("encoder.encoder.3.mconv.7.running_var") == "3.mconv.7.layer.0.running_var"
    assert fix_encoder_name("encoder.encoder.1.res.0.0.conv.weight") == "1.res.0.conv.weight"
    assert fix_encoder_name("encoder.encoder.1.res.0.1.num_batches_tracked") == "1.res.1.layer.0.num_batches_tracked"
    assert fix_encoder_name("encoder.encoder.2.mconv.23.fc.0.weight") == "2.mconv.23.layer.0.fc.0.weight"   # Citrinet SE
    for k in ("0.mconv.0.conv.weight", "0.mconv.2.layer.0.weight", "1.res.0.conv.weight", "1.res.1.layer.0.bias",
              "2.mconv.23.layer.0.fc.2.weight"):
        assert fix_encoder_name(to_nemo_name(k)) == k


def test_quartznet_nemo_roundtrip(tmp_path):
    filters, kernels = [32, 32, 64, 64, 64], [5, 7, 9, 11, 13]
    labels = [" "] + [chr(ord("a") + i) for i in range(26)] + ["'"]
    config = {
        "sample_rate": 16000, "labels": labels,
        "preprocessor": {"params": {"sample_rate": 16000, "window_size": 0.02, "window_stride": 0.01, "n_fft": 512,
                                    "features": 64, "dither": 1e-5}},
        "encoder": {"params": {"feat_in": 64, "jasper": jasper_cfg(filters, kernels)}},
        "decoder": {"params": {"feat_in": 1024, "num_classes": 28, "vocabulary": "${labels}"}},
    }
    enc_state = synth.encoder_state(synth.quartznet_block_list(filters=filters, kernel_sizes=kernels, repeat_blocks=1), seed=3)
    dec_state = synth.decoder_state(1024, 29, seed=4)
    nemo_state = {to_nemo_name(k): torch.from_numpy(np.asarray(v)) for k, v in enc_state.items()}
    nemo_state["decoder.decoder_layers.0.weight"] = torch.from_numpy(dec_state["weight"])
    nemo_state["decoder.decoder_layers.0.bias"] = torch.from_numpy(dec_state["bias"])
    path = tmp_path / "tiny_quartznet.nemo"
    write_nemo(path, config, nemo_state)
    m = load_quartznet_checkpoint(str(path))
    assert not m.training and m.encoder_final_dimension == 1024
    assert m.text_transform.vocab.itos == labels + ["<blank>"] and m.text_transform.num_tokens == 29
    got = m.encoder.state_dict()
    assert set(got) == set(enc_state)
    for k, v in enc_state.items():
        assert torch.equal(got[k], torch.from_numpy(np.asarray(v))), k
    assert torch.equal(m.decoder.weight, torch.from_numpy(dec_state["weight"]))
    fb = m.audio_transform
    assert fb[1].hop_length == 160 and fb[1].n_fft == 512
    # vocabulary through the decoder section + `${labels}` interpolation when there is no top-level `labels`
    cfg2 = dict(config)
    cfg2["decoder"] = {"params": {"vocabulary": labels}}
    del cfg2["labels"]
    (tmp_path / "c2.yaml").write_text(yaml.safe_dump(cfg2))
    _, _, tt = load_components_from_quartznet_config(tmp_path / "c2.yaml", {"dropout": 0.1})
    assert tt.num_tokens == 29
    conf = load_yaml_config(tmp_path / "c2.yaml")
    assert conf["decoder"]["params"]["vocabulary"] == labels
    (tmp_path / "c3.yaml").write_text(yaml.safe_dump(config))
    assert load_yaml_config(tmp_path / "c3.yaml")["decoder"]["params"]["vocabulary"] == labels   # ${labels} resolved
    # a mismatching checkpoint fails strictly, like the reference
    bad = dict(nemo_state)
    bad.pop(to_nemo_name("1.res.0.conv.weight"))
    write_nemo(tmp_path / "bad.nemo", config, bad)
    with pytest.raises(RuntimeError):
        load_quartznet_checkpoint(str(tmp_path / "bad.nemo"))
    with pytest.raises(FileNotFoundError):
        load_quartznet_checkpoint(str(tmp_path / "missing.nemo"))


def test_citrinet_nemo_roundtrip(tmp_path):
    filters, kernels, strides = [64, 64, 64], [5, 7, 9], [1, 2, 1]
    vocab = ["the", "##s", "a", "##ing"]
    config = {
        "preprocessor": {"sample_rate": 16000, "window_size": 0.025, "window_stride": 0.01, "n_fft": 512, "features": 80,
                         "dither": 1e-5},
        "encoder": {"jasper": jasper_cfg(filters, kernels, strides, citrinet=True)},
        "decoder": {"vocabulary": vocab},
    }
    assert fix_vocab(vocab) == ["▁the", "s", "▁a", "ing"]
    (tmp_path / "c.yaml").write_text(yaml.safe_dump(config))
    enc, fb, tt = load_components_from_citrinet_config(tmp_path / "c.yaml", None)
    assert tt.vocab.itos[:4] == ["▁the", "s", "▁a", "ing"] and tt.num_tokens == 5
    assert fb[1].win_length == 400
    enc_state = {k: v.clone() for k, v in enc.state_dict().items()}
    for k, v in enc_state.items():
        if v.dtype.is_floating_point:
            v.uniform_(0.1, 1.0)
    dec = torch.nn.Conv1d(640, 5, 1)
    nemo_state = {to_nemo_name(k): v for k, v in enc_state.items()}
    nemo_state["decoder.decoder_layers.0.weight"] = dec.weight.detach()
    nemo_state["decoder.decoder_layers.0.bias"] = dec.bias.detach()
    write_nemo(tmp_path / "tiny_citrinet.nemo", config, nemo_state)
    m = load_citrinet_checkpoint(str(tmp_path / "tiny_citrinet.nemo"))
    for k, v in enc_state.items():
        assert torch.equal(m.encoder.state_dict()[k], v), k
    assert m.encoder_final_dimension == 640


def test_extract_nemo_rejects_path_traversal(tmp_path):
    evil = tmp_path / "evil.nemo"
    with tarfile.open(evil, "w") as tar:
        data = b"x"
        info = tarfile.TarInfo("../escape.txt")
        info.size = len(data)
        tar.addfile(info, io.BytesIO(data))
    with pytest.raises(Exception):
        extract_nemo(evil, tmp_path / "out")
    assert not (tmp_path / "escape.txt").exists()


REF_SAMPLES = Path("/root/reference/tests/nemo_config_samples")


@pytest.mark.skipif(not REF_SAMPLES.is_dir(), reason="the reference's config samples are only present in the build container")
def test_reference_config_samples_build_matching_encoders():
    """tests/quartznet/test_compatibility_qn.py:30-57 on the reference's own NeMo config samples: the encoder built from
    the YAML is state_dict-compatible with QuartznetEncoder() (5x5) / QuartznetEncoder(repeat_blocks=3) (15x5)."""
    n = 0
    for cfg in REF_SAMPLES.glob("*.yaml"):
        encoder, fb, tt = load_components_from_quartznet_config(cfg)
        ref = QuartznetEncoder() if "Net5x5" in cfg.name else QuartznetEncoder(repeat_blocks=3)
        ref.load_state_dict(encoder.state_dict(), strict=True)
        assert tt.num_tokens == 29 and fb[1].n_fft == 512
        n += 1
    assert n == 3


@pytest.mark.gpu
def test_loaded_checkpoint_drives_the_kernels(tmp_path):
    """A `.nemo` archive loaded through load_quartznet_checkpoint predicts exactly what the same weights loaded directly do."""
    from thunder_speech_b200.runner import build_model

    labels = [" "] + [chr(ord("a") + i) for i in range(26)] + ["'"]
    filters, kernels = [256, 256, 512, 512, 512], [33, 39, 51, 63, 75]
    config = {"labels": labels,
              "preprocessor": {"params": {"sample_rate": 16000, "window_size": 0.02, "window_stride": 0.01, "n_fft": 512,
                                          "features": 64, "dither": 1e-5}},
              "encoder": {"params": {"jasper": jasper_cfg(filters, kernels)}}}
    ref = build_model("quartznet5x5", torch.device("cuda"), seed=5)
    nemo_state = {to_nemo_name(k): v.detach().cpu() for k, v in ref.encoder.state_dict().items()}
    nemo_state["decoder.decoder_layers.0.weight"] = ref.decoder.weight.detach().cpu()
    nemo_state["decoder.decoder_layers.0.bias"] = ref.decoder.bias.detach().cpu()
    write_nemo(tmp_path / "qn5x5.nemo", config, nemo_state)
    m = load_quartznet_checkpoint(str(tmp_path / "qn5x5.nemo")).cuda()
    x = torch.from_numpy(synth.audio(3, 24000, 9, "tones")).cuda()
    assert m.predict(x) == ref.predict(x)
    la, _ = m(x, torch.full((3,), 24000, device="cuda"))
    lb, _ = ref(x, torch.full((3,), 24000, device="cuda"))
    assert torch.equal(la, lb)
