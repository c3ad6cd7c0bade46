#!/usr/bin/env python
"""Is the pair GEMM power / clock bound?  Loops one GEMM shape for ~2 s per variant while sampling nvidia-smi
(SM clock, power): python tools/gemm_power.py [cin cout]   (variants through THUNDER_B200_OPTIONS / dbg bits)"""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from thunder_speech_b200 import ops, _lib

cin = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cout = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
B, T = 256, 751
P = ops.row_pitch(T)
dev = torch.device("cuda")
lens = torch.full((B,), T, dtype=torch.int32, device=dev)
xs = [torch.randn(B, cin, P, device=dev).bfloat16() for _ in range(3)]
w = (torch.randn(cout, cin, device=dev) / cin ** 0.5).bfloat16()
sh = torch.randn(cout, device=dev)
fl = 2 * B * T * cin * cout


def sample(stop, rows):
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                         stdout=subprocess.PIPE, text=True)
    while not stop.is_set():
        line = p.stdout.readline()
        if line:
            rows.append([float(v) for v in line.split(",")])
    p.terminate()


for name, opts in (("streaming", {"pw_ws": 0, "dbg": 0}), ("weights in TMEM", {"pw_ws": 1, "dbg": 0}),
                   ("streaming, no epilogue", {"pw_ws": 0, "dbg": 32}), ("streaming, no loads", {"pw_ws": 0, "dbg": 24}),
                   ("streaming, MMA + B only", {"pw_ws": 0, "dbg": 48})):
    for k, v in opts.items():
        _lib.set_option(k, v)
    run = lambda i: ops.pw_gemm(w, xs[i % 3], None, None, T, sh, lens, False, True, None, None, None)
    for i in range(5):
        run(i)
    torch.cuda.synchronize()
    stop, rows = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, rows))
    th.start()
    time.sleep(0.3)
    n = 12000
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    stop.set()
    th.join()
    us = e0.elapsed_time(e1) / n * 1e3
    r = np.array(rows[3:]) if len(rows) > 4 else np.zeros((1, 2))
    mhz = float(np.median(r[:, 0]))
    print(f"{cin}->{cout} {name:26s}: {us:7.1f} us  {fl / us / 1e6:5.0f} TFLOP/s  SM clock {mhz:6.0f} MHz  power {np.median(r[:, 1]):5.0f} W"
          f"  -> {fl / us / 1e6 / max(mhz, 1) * 1e3:6.1f} TFLOP/s per GHz", flush=True)
_lib.set_option("dbg", 0)
_lib.set_option("pw_ws", 0)
