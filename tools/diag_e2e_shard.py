#!/usr/bin/env python
"""Where does the end-to-end rate of a small per-GPU shard go?  torchrun --nproc-per-node N tools/diag_e2e_shard.py [per_gpu_batch] [steps]
Per rank: graph replay ms/step (device only), predict_stream loop ms/step, final all_gather_object ms, host-side cost of one
submit() and one collect()."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from thunder_speech_b200 import runner, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
m = runner.build_model("quartznet15x5", dev)
host = torch.from_numpy(synth.audio(B, 15 * 16000, 1234 + rank, "noise")).pin_memory()
x = host.to(dev)
for _ in range(3):
    m.predict_ids_graphed(x, in_place=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    m.predict_ids_graphed(x, in_place=True)
e1.record()
torch.cuda.synchronize()
dev_ms = e0.elapsed_time(e1) / steps
list(m.predict_stream(host for _ in range(3)))        # builds the pipe
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
mine = [t for t in m.predict_stream(host for _ in range(steps))]
t1 = time.perf_counter()
if world > 1:
    parts = [None] * world
    dist.all_gather_object(parts, mine)
t2 = time.perf_counter()
# the product path: stream + overlapped transcript gather
from thunder_speech_b200.parallel import sharded_predict_stream
sharded_predict_stream(m, (host for _ in range(3)), presharded=True)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t3 = time.perf_counter()
out = sharded_predict_stream(m, (host for _ in range(steps)), presharded=True)
t4 = time.perf_counter()
assert len(out) == steps and len(out[0]) == B * world and out[0][rank * B:(rank + 1) * B] == mine[0]
print(f"rank {rank}: sharded_predict_stream {(t4 - t3) / steps * 1e3:.3f} ms/step (stream + all_gather_object was "
      f"{(t2 - t0) / steps * 1e3:.3f})", flush=True)
# host cost of the two halves of a step, GPU idle
pipe = next(iter(m._pipes.values()))
pipe.reset()
ts, tc = [], []
for i in range(6):
    a = time.perf_counter(); s = pipe.submit(i, host); b = time.perf_counter()
    torch.cuda.synchronize()
    c = time.perf_counter(); pipe.collect(s); d = time.perf_counter()
    ts.append(b - a); tc.append(d - c)
print(f"rank {rank}/{world} B={B}: device {dev_ms:.3f} ms/step | stream loop {(t1 - t0) / steps * 1e3:.3f} ms/step | "
      f"all_gather_object {(t2 - t1) * 1e3:.2f} ms total = {(t2 - t1) / steps * 1e3:.3f} ms/step | host submit {min(ts) * 1e3:.3f} ms "
      f"collect (detokenise) {min(tc) * 1e3:.3f} ms", flush=True)
if world > 1:
    dist.destroy_process_group()
