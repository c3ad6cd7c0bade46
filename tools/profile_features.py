import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from thunder_speech_b200 import synth
from thunder_speech_b200.quartznet.transform import FilterbankFeatures
dev = torch.device("cuda:0")
fb = FilterbankFeatures().eval().to(dev)
x = torch.from_numpy(synth.audio(64, 320000, 1234, "noise")).to(dev)
l = torch.full((64,), 320000, device=dev)
for _ in range(2): fb(x, l)
torch.cuda.synchronize(); torch.cuda.profiler.start(); fb(x, l); torch.cuda.synchronize(); torch.cuda.profiler.stop()
