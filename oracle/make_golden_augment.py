"""tests/golden/augment.npz: outputs of the REFERENCE's SpecAugment / SpecCutout (src/thunder/quartznet/spec_augment.py) in
train() mode under fixed torch seeds (the draws come from torch.rand(1) on the host generator).
Run in the build container:  python -m oracle.make_golden_augment"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [  # name, kind, kwargs, (B, C, T), seed
    ("aug_t2_f1", "augment", dict(time_masks=2, freq_masks=1, time_width=50, freq_width=20), (3, 64, 201), 11),
    ("aug_t0_f3", "augment", dict(time_masks=0, freq_masks=3, time_width=10, freq_width=27), (2, 80, 97), 12),
    ("aug_t4", "augment", dict(time_masks=4, freq_masks=0, time_width=120, freq_width=10), (1, 64, 333), 13),
    ("cut_3", "cutout", dict(rect_masks=3, time_width=5, freq_width=20), (2, 64, 150), 14),
    ("cut_5", "cutout", dict(rect_masks=5, time_width=50, freq_width=40), (4, 80, 64), 15),
]


def case_input(name, shape):
    rng = np.random.Generator(np.random.PCG64(sum(map(ord, name))))
    return (rng.standard_normal(shape) + 3.0).astype(np.float32)      # no exact zeros: every masked element is visible


def main():
    pkg = types.ModuleType("thunder"); pkg.__path__ = ["/root/reference/src/thunder"]; sys.modules["thunder"] = pkg
    spec = importlib.util.spec_from_file_location("thunder.quartznet.spec_augment",
                                                  "/root/reference/src/thunder/quartznet/spec_augment.py")
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    out = {}
    for name, kind, kw, shape, seed in CASES:
        layer = (mod.SpecAugment if kind == "augment" else mod.SpecCutout)(**kw).train()
        x = torch.from_numpy(case_input(name, shape))
        torch.manual_seed(seed)
        y = layer(x)
        out[f"{name}.out"] = y.numpy()
        print(name, "masked fraction", float((y == 0).float().mean()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "augment.npz"), **out)


if __name__ == "__main__":
    main()
