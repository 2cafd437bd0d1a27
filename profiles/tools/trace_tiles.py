import sys, json, torch
sys.path.insert(0, '.')
from vexautonomousplanner_b200 import synth
from vexautonomousplanner_b200.engine import Engine
from torch.profiler import profile, ProfilerActivity
tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 2
eng = Engine("cuda:0")
db = eng.upload(synth.random_paths(4096, 8, seed=0))
for _ in range(4):
    eng.profile(db, reuse_plan=True, tiles=tiles)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    eng.profile(db, reuse_plan=True, tiles=tiles)
    torch.cuda.synchronize()
print("wall ms", (time.perf_counter() - t0) * 1e3)
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t00 = ev[0].time_range.start
for e in ev:
    if e.time_range.end - e.time_range.start > 30:
        print(f"{(e.time_range.start - t00)/1e3:8.3f} {(e.time_range.end - t00)/1e3:8.3f} dur={(e.time_range.end-e.time_range.start)/1e3:7.3f}  {e.name[:40]}")
