"""Load Citrinet from an original NeMo `.nemo` checkpoint into this package's modules
(mirror of src/thunder/citrinet/compatibility.py:54-176)."""
from __future__ import annotations

from pathlib import Path
from tempfile import TemporaryDirectory
from typing import Dict, List, Tuple, Union

from torch import nn

from ..blocks import conv1d_decoder
from ..compat import extract_nemo, load_quartznet_weights, load_yaml_config
from ..module import BaseCTCModule
from ..quartznet.transform import FilterbankFeatures
from ..text_processing import BatchTextTransformer
from .blocks import CitrinetEncoder

__all__ = ["load_components_from_citrinet_config", "fix_vocab", "load_citrinet_checkpoint"]


def fix_vocab(vocab_tokens: List[str]) -> List[str]:
    """NeMo word-piece tokens back to sentencepiece style: `##x` -> `x`, anything else gets the `▁` word-boundary prefix
    (compatibility.py:113-130)."""
    return [t[2:] if t.startswith("##") else "▁" + t for t in vocab_tokens]


def load_components_from_citrinet_config(config_path: Union[str, Path], sentencepiece_path: Union[str, Path] = None,
                                         augment_params: Dict = None) -> Tuple[nn.Module, nn.Module, BatchTextTransformer]:
    """(encoder, audio_transform, text_transform) from `model_config.yaml` (compatibility.py:54-110): the body is
    `encoder.jasper[1:-1]` with per-block filters / kernel / stride.  The sentencepiece model (when the file exists) is only used
    to ENCODE training targets; decoding needs the token list alone."""
    augment_params = dict(augment_params or {})
    conf = load_yaml_config(config_path)
    body = conf["encoder"]["jasper"][1:-1]
    encoder_cfg = {
        "filters": [cfg["filters"] for cfg in body],
        "kernel_sizes": [cfg["kernel"][0] for cfg in body],
        "strides": [cfg["stride"][0] for cfg in body],
        "dropout": augment_params.pop("dropout", 0.0),
    }
    pre = conf["preprocessor"]
    preprocess_cfg = {
        "sample_rate": pre["sample_rate"],
        "n_window_size": int(pre["window_size"] * pre["sample_rate"]),
        "n_window_stride": int(pre["window_stride"] * pre["sample_rate"]),
        "n_fft": pre["n_fft"],
        "nfilt": pre["features"],
        "dither": pre["dither"],
        **augment_params,
    }
    labels = conf["labels"] if "labels" in conf else conf["decoder"]["vocabulary"]
    encoder = CitrinetEncoder(**encoder_cfg)
    sp = str(sentencepiece_path) if sentencepiece_path is not None and Path(sentencepiece_path).is_file() else None
    text_transform = BatchTextTransformer(tokens=fix_vocab(list(labels)), sentencepiece_model=sp)
    audio_transform = FilterbankFeatures(**preprocess_cfg)
    return encoder, audio_transform, text_transform


def load_citrinet_checkpoint(checkpoint: Union[str, Path], save_folder: str = None, augment_params: Dict = None
                             ) -> BaseCTCModule:
    """`.nemo` file -> `BaseCTCModule` in eval mode (compatibility.py:133-176); local paths only."""
    with TemporaryDirectory() as extract_folder:
        extract_path = extract_nemo(checkpoint, extract_folder)
        encoder, audio_transform, text_transform = load_components_from_citrinet_config(
            extract_path / "model_config.yaml", extract_path / "tokenizer.model", augment_params)
        decoder = conv1d_decoder(640, num_classes=text_transform.num_tokens)
        load_quartznet_weights(encoder, decoder, str(extract_path / "model_weights.ckpt"))
        module = BaseCTCModule(encoder, decoder, audio_transform, text_transform, encoder_final_dimension=640)
        return module.eval()
