"""Run exactly ONE eager forward of a bench workload between cudaProfilerStart/Stop (for ncu
--profile-from-start off).  usage: python tools/profile_forward.py quartznet15x5 [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from thunder_speech_b200 import synth
from thunder_speech_b200.runner import build_model

name = sys.argv[1] if len(sys.argv) > 1 else "quartznet15x5"
B = int(sys.argv[2]) if len(sys.argv) > 2 else {"quartznet15x5": 256, "citrinet1024": 128, "quartznet5x5": 4}[name]
secs = {"quartznet15x5": 15, "citrinet1024": 20, "quartznet5x5": 10}[name]
dev = torch.device("cuda:0")
m = build_model(name, dev)
x = torch.from_numpy(synth.audio(B, secs * 16000, 1234, "noise")).to(dev)
for _ in range(2):
    m.predict_ids(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
m.predict_ids(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
