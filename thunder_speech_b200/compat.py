"""Shared pieces of the NeMo-checkpoint loaders: `.nemo` archive extraction, config reading
without OmegaConf, and the reference's weight-name mapping (src/thunder/quartznet/compatibility.py:127-158).

Host-side only: no kernel is involved.  Downloading (`download_checkpoint`, src/thunder/utils.py) is outside the scope of
this package -- there is no network where it runs; pass a local `.nemo` path."""
from __future__ import annotations

import re
import tarfile
from pathlib import Path
from typing import Any, Dict, Union

import torch
import yaml
from torch import nn

__all__ = ["extract_nemo", "load_yaml_config", "fix_encoder_name", "load_quartznet_weights"]


def extract_nemo(nemo_path: Union[str, Path], folder: Union[str, Path]) -> Path:
    """Unpacks a `.nemo` file (a tar archive, optionally gzip-compressed) into `folder`; members are restricted to plain
    files below the folder (tarfile's "data" filter)."""
    nemo_path = Path(nemo_path)
    if not nemo_path.is_file():
        raise FileNotFoundError(f"{nemo_path} is not a local .nemo file (downloading checkpoints is not supported here)")
    with tarfile.open(str(nemo_path), "r:*") as tar:
        tar.extractall(str(folder), filter="data")
    return Path(folder)


_INTERP = re.compile(r"^\$\{([^}]+)\}$")


def _resolve(node: Any, root: Dict[str, Any], depth: int = 0) -> Any:
    """OmegaConf-style whole-value interpolation (`${a.b}`), which NeMo configs use e.g. for `vocabulary: ${labels}`."""
    if isinstance(node, str):
        m = _INTERP.match(node.strip())
        if m and depth < 8:
            cur: Any = root
            for part in m.group(1).split("."):
                cur = cur[int(part)] if isinstance(cur, list) else cur[part]
            return _resolve(cur, root, depth + 1)
        return node
    if isinstance(node, dict):
        return {k: _resolve(v, root, depth) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root, depth) for v in node]
    return node


class _Conf(dict):
    """Dict whose string values of the form `${path}` are resolved against the root on access."""

    def __init__(self, data: Dict[str, Any], root: Dict[str, Any] = None):
        super().__init__(data)
        self._root = root if root is not None else data

    def __getitem__(self, key):
        v = _resolve(super().__getitem__(key), self._root)
        return _Conf(v, self._root) if isinstance(v, dict) else v


def load_yaml_config(config_path: Union[str, Path]) -> _Conf:
    with open(config_path, "r", encoding="utf-8") as f:
        data = yaml.safe_load(f)
    if not isinstance(data, dict):
        raise ValueError(f"{config_path}: expected a mapping at the top level of model_config.yaml")
    return _Conf(data)


def fix_encoder_name(x: str) -> str:
    """NeMo parameter name -> this package's (= the reference's) name: drop the `encoder.` prefixes, `.res.0` -> `.res`,
    and insert the `layer.0` level of the Masked wrapper for everything that is not a masked conv
    (quartznet/compatibility.py:137-145)."""
    x = x.replace("encoder.", "").replace(".res.0", ".res")
    if ".conv" not in x:
        parts = x.split(".")
        x = ".".join(parts[:3] + ["layer", "0"] + parts[3:])
    return x


def load_quartznet_weights(encoder: nn.Module, decoder: nn.Module, weights_path: str) -> None:
    """Loads `model_weights.ckpt` of a `.nemo` file into encoder / decoder, strict (quartznet/compatibility.py:127-158).
    The checkpoint is read with `weights_only=True` (tensors only, no arbitrary unpickling)."""
    weights = torch.load(weights_path, map_location="cpu", weights_only=True)
    encoder_weights = {fix_encoder_name(k): v for k, v in weights.items() if "encoder" in k}
    encoder.load_state_dict(encoder_weights, strict=True)
    decoder_weights = {k.replace("decoder.decoder_layers.0.", ""): v for k, v in weights.items() if "decoder" in k}
    decoder.load_state_dict(decoder_weights, strict=True)
