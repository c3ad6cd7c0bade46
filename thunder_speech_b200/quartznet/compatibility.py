"""Load QuartzNet from an original NeMo `.nemo` checkpoint into this package's modules
(mirror of src/thunder/quartznet/compatibility.py:62-205; same function names and return values)."""
from __future__ import annotations

from pathlib import Path
from tempfile import TemporaryDirectory
from typing import Dict, Tuple, Union

from torch import nn

from ..blocks import conv1d_decoder
from ..compat import extract_nemo, load_quartznet_weights, load_yaml_config
from ..module import BaseCTCModule
from ..text_processing import BatchTextTransformer
from .blocks import QuartznetEncoder
from .transform import FilterbankFeatures

__all__ = ["load_components_from_quartznet_config", "load_quartznet_weights", "load_quartznet_checkpoint"]


def load_components_from_quartznet_config(config_path: Union[str, Path], augment_params: Dict = None
                                          ) -> Tuple[nn.Module, nn.Module, BatchTextTransformer]:
    """(encoder, audio_transform, text_transform) from `model_config.yaml` (compatibility.py:62-124): the body is
    `encoder.params.jasper[1:-2]`, one block per entry (so 15x5 arrives as 15 single-repeat blocks, state_dict-compatible
    with `QuartznetEncoder(repeat_blocks=3)`)."""
    augment_params = dict(augment_params or {})
    conf = load_yaml_config(config_path)
    body = conf["encoder"]["params"]["jasper"][1:-2]
    encoder_cfg = {
        "filters": [cfg["filters"] for cfg in body],
        "kernel_sizes": [cfg["kernel"][0] for cfg in body],
        "dropout": augment_params.pop("dropout", 0.0),
    }
    pre = conf["preprocessor"]["params"]
    preprocess_cfg = {
        "sample_rate": pre["sample_rate"],
        "n_window_size": int(pre["window_size"] * pre["sample_rate"]),
        "n_window_stride": int(pre["window_stride"] * pre["sample_rate"]),
        "n_fft": pre["n_fft"],
        "nfilt": pre["features"],
        "dither": pre["dither"],
        **augment_params,
    }
    labels = conf["labels"] if "labels" in conf else conf["decoder"]["params"]["vocabulary"]
    audio_transform = FilterbankFeatures(**preprocess_cfg)
    encoder = QuartznetEncoder(**encoder_cfg)
    text_transform = BatchTextTransformer(tokens=list(labels))
    return encoder, audio_transform, text_transform


def load_quartznet_checkpoint(checkpoint: Union[str, Path], save_folder: str = None, augment_params: Dict = None
                              ) -> BaseCTCModule:
    """`.nemo` file -> `BaseCTCModule` in eval mode (compatibility.py:161-205).  `checkpoint` must be a local path: the
    reference's `QuartznetCheckpoint` download enum needs a network and is not provided."""
    with TemporaryDirectory() as extract_folder:
        extract_path = extract_nemo(checkpoint, extract_folder)
        encoder, audio_transform, text_transform = load_components_from_quartznet_config(
            extract_path / "model_config.yaml", augment_params)
        decoder = conv1d_decoder(1024, text_transform.num_tokens)
        load_quartznet_weights(encoder, decoder, str(extract_path / "model_weights.ckpt"))
        module = BaseCTCModule(encoder, decoder, audio_transform, text_transform, encoder_final_dimension=1024)
        return module.eval()
