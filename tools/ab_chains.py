#!/usr/bin/env python
"""A/B: one captured inference graph split into N independent utterance chains (CTCModule.graph_chains).

    python tools/ab_chains.py [quartznet15x5|citrinet1024] [batches...]

Prints ms per replay and audio-s/s per (batch, chains) and checks that the token ids equal the single-chain result."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from thunder_speech_b200 import runner, synth  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "quartznet15x5"
    secs = 15 if name.startswith("quartznet") else 20
    batches = [int(v) for v in sys.argv[2:]] or ([256, 128, 64, 32] if secs == 15 else [128, 64, 32, 16])
    chains_list = [int(v) for v in os.environ.get("CHAINS", "1,2,3,4").split(",")]
    dev = torch.device("cuda", 0)
    model = runner.build_model(name, dev)
    N = secs * 16000
    for B in batches:
        x = [torch.from_numpy(synth.audio(B, N, 1234 + i, "noise")).to(dev) for i in range(2)]
        ref = None
        for c in chains_list:
            model.graph_chains = c
            outs = None
            for i in range(4):
                outs = model.predict_ids_graphed(x[i % 2], in_place=True)
            torch.cuda.synchronize()
            ids = model.predict_ids_graphed(x[0], in_place=True)[0].clone()
            if ref is None:
                ref = ids
            same = bool(torch.equal(ids, ref))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20
            e0.record()
            for i in range(n):
                model.predict_ids_graphed(x[i % 2], in_place=True)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            print(f"{name} B={B:4d} chains={c}: {ms:8.3f} ms  {B * secs / ms * 1e3:10.0f} audio-s/s  ids_equal={same}", flush=True)
        model.invalidate_graphs()
        del x
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
