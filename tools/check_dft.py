"""DFT (tensor-core) vs FFT (SIMT) log-mel kernels: agreement, parity against the numpy oracle, time per launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from thunder_speech_b200 import _lib, synth
from thunder_speech_b200.quartznet.transform import FilterbankFeatures

L = _lib.lib()
dev = torch.device("cuda")
st = lambda: torch.cuda.current_stream().cuda_stream

def run_fft(fb, t, a):
    B, N = a.shape; F = 1 + N // 160; nf = t["mel_start"].numel()
    out = torch.empty((B, nf, F), device=dev)
    _lib.check(L.ts_logmel(a.data_ptr(), B, N, 512, 160, 0.97, t["window_full"].data_ptr(), t["win_lo"], t["win_hi"],
                           t["twiddle"].data_ptr(), t["mel_start"].data_ptr(), t["mel_count"].data_ptr(), t["mel_off"].data_ptr(),
                           t["mel_w"].data_ptr(), nf, t["mel_w"].numel(), out.data_ptr(), st()), "ts_logmel")
    return out

def run_dft(fb, t, a, partials=None, lengths=None):
    B, N = a.shape; F = 1 + N // 160; nf = t["mel_start"].numel()
    out = torch.full((B, nf, F), float("nan"), device=dev)
    _lib.check(L.ts_logmel_dft(a.data_ptr(), B, N, 160, 0.97, t["wplus"].data_ptr(), t["wminus"].data_ptr(), t["basis"].data_ptr(),
                               t["mel_w2"].data_ptr(), t["mel_adv"].data_ptr(), nf, out.data_ptr(),
                               partials.data_ptr() if partials is not None else None,
                               lengths.data_ptr() if lengths is not None else None, st()), "ts_logmel_dft")
    return out

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for nf, B, N, kind in ((64, 2, 4800, "noise"), (64, 3, 16000, "tones"), (80, 2, 3333, "noise"), (64, 2, 700, "noise"), (64, 5, 160 * 300 + 7, "tones")):
    fb = FilterbankFeatures(nfilt=nf).eval().to(dev)
    t = fb._device_tables(dev)
    a = torch.from_numpy(synth.audio(B, N, 11, kind)).to(dev)
    x, y = run_fft(fb, t, a), run_dft(fb, t, a)
    torch.cuda.synchronize()
    d = (x - y).abs()
    print(f"nfilt={nf} B={B} N={N} {kind}: max|fft-dft| = {float(d.max()):.3e}  (max|logmel| {float(x.abs().max()):.2f}), nan={int(torch.isnan(y).sum())}", flush=True)
    if float(d.max()) > 1e-2 or torch.isnan(y).any():
        bad = torch.nonzero((d > 1e-2) | torch.isnan(y))
        print("   first bad (b, filter, frame):", bad[:8].tolist(), " frames bad:", sorted(set(bad[:, 2].tolist()))[:20])
for dbg in (7, 15, 23, 31, 0):
    _lib.set_option("dbg", dbg)
    fb = FilterbankFeatures(nfilt=64).eval().to(dev)
    t = fb._device_tables(dev)
    a = torch.from_numpy(synth.audio(64, 20 * 16000, 1234, "noise")).to(dev)
    print(f"dbg={dbg} (1: no mel walk, 2: no builder loads, 4: no MMAs, 8: builders protocol only, 16: no TMEM loads): dft {timeit(lambda: run_dft(fb, t, a)):.1f} us", flush=True)
for B, secs in ((64, 20), (256, 15)):
    fb = FilterbankFeatures(nfilt=64).eval().to(dev)
    t = fb._device_tables(dev)
    a = torch.from_numpy(synth.audio(B, secs * 16000, 1234, "noise")).to(dev)
    x, y = run_fft(fb, t, a), run_dft(fb, t, a)
    print(f"B={B} x {secs}s: max|fft-dft| = {float((x - y).abs().max()):.3e}; fft {timeit(lambda: run_fft(fb, t, a)):.1f} us, dft {timeit(lambda: run_dft(fb, t, a)):.1f} us", flush=True)
