"""One training step (graph replay) under the profiler: ncu --profile-from-start off ... python tools/profile_train.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from thunder_speech_b200 import runner
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
wl = runner.TrainWorkload("quartznet15x5_train", B, 15 * 16000, 64, torch.device("cuda", 0), 0)
for i in range(3): wl.step_device(i)
torch.cuda.synchronize()
torch.cuda.profiler.start()
wl.step_device(0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
