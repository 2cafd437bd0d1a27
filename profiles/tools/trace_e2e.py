import sys, torch, time
sys.path.insert(0, '.')
from vexautonomousplanner_b200 import synth
from vexautonomousplanner_b200.engine import Engine
from torch.profiler import profile, ProfilerActivity
tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 4
eng = Engine("cuda:0")
packed = synth.random_paths(4096, 8, seed=0)
g = eng.capture(eng.upload(packed), tiles=tiles, to_host=True)
for _ in range(3):
    g.run_host(packed)
t0 = time.perf_counter(); g.run_host(packed); print("run_host ms", (time.perf_counter() - t0) * 1e3)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    g.run_host(packed)
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t00 = ev[0].time_range.start
for e in ev:
    d = (e.time_range.end - e.time_range.start) / 1e3
    if d > 0.2 or 'pack' in e.name or 'Memcpy' in e.name and d > 0.05:
        print(f"{(e.time_range.start - t00)/1e3:8.3f} {(e.time_range.end - t00)/1e3:8.3f} dur={d:7.3f}  {e.name[:36]}")
print("last end", max((e.time_range.end - t00) / 1e3 for e in ev))
