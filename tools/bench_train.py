"""Per-kernel breakdown of the training step (bench workload quartznet15x5_train): python tools/bench_train.py [B]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from thunder_speech_b200 import runner
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
wl = runner.TrainWorkload("quartznet15x5_train", B, 15 * 16000, 64, torch.device("cuda", 0), 0)
for i in range(3): wl.step_device(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(5): wl.step_device(i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("step ms", ms, "audio-s/s", B * 15 / ms * 1e3, "loss", float(wl.last_loss), "mem GB", torch.cuda.max_memory_allocated() / 1e9)
r = wl.roofline(2)
pk = r.pop("per_kernel")
print(json.dumps(r))
for k, v in sorted(pk.items(), key=lambda kv: -kv[1]["ms_per_step"]):
    print(f"{k:16s} {v['ms_per_step']:8.3f} ms  share {v['share_of_step']:.3f}  launches {v['launches_per_step']:4d}  {v['GBps']:8.0f} GB/s  {v['TFLOPs']:7.1f} TF/s")
