"""Ingest kernels at BASELINE config 3's batch (256 utterances x 15 s): python tools/bench_ingest.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from thunder_speech_b200.data import pcm_ingest, resample
def gtime(f, iters=10):
    for _ in range(2): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
B, secs = 256, 15
for (sr, ch) in [(16000, 1), (16000, 2), (44100, 2), (48000, 1), (8000, 1)]:
    N = sr * secs
    pcm = torch.randint(-3000, 3000, (B, N, ch), dtype=torch.int16, device="cuda")
    lens = torch.full((B,), N, dtype=torch.int32, device="cuda")
    t1 = gtime(lambda: pcm_ingest(pcm, lens, interleaved=True))
    by1 = B * N * ch * 2 + B * N * 4
    mono = pcm_ingest(pcm, lens, interleaved=True)
    msg = f"sr {sr} ch {ch}: ingest {t1:8.1f} us  {by1/t1/1e3:6.0f} GB/s (int16 read once + f32 write)"
    if sr != 16000:
        t2 = gtime(lambda: resample(mono, sr, 16000, lens))
        by2 = B * N * 4 + B * 16000 * secs * 4
        import math
        g = math.gcd(sr, 16000); o, n = sr // g, 16000 // g
        w = math.ceil(6 * o / (min(o, n) * 0.99)); taps = 2 * w + o
        msg += f" | resample {t2:8.1f} us  {by2/t2/1e3:6.0f} GB/s  {B*16000*secs*taps/t2/1e6:5.2f} TFMA/s (taps {taps})"
    print(msg)
