import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from vexautonomousplanner_b200 import synth, export as ex
from vexautonomousplanner_b200.engine import Engine
eng = Engine("cuda:0")
packed = synth.random_paths(4096, 8, seed=0)
res = eng.profile(eng.upload(packed))
for _ in range(3):
    text, roff, poff = ex.export_text_device(eng, res)
torch.cuda.synchronize()
s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    text, roff, poff = ex.export_text_device(eng, res)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 5
R = int(poff[-1]); nbytes = text.numel()
print(f"device export text: {R} rows, {nbytes/1e6:.1f} MB of text in {ms:.3f} ms -> {nbytes/1e9/(ms*1e-3):.1f} GB/s text, {R*6/(ms*1e-3)/1e9:.2f} G floats/s")
# the reference's way on the host, bounded sample
rows, _ = ex.export_rows(eng, res)
r = rows[:20000].cpu().numpy()
t0 = time.perf_counter()
txt = ex.format_rows([[0] + [np.float64(v) for v in row[1:]] for row in r])
el = time.perf_counter() - t0
print(f"python f-string formatting: 20000 rows in {el*1e3:.1f} ms -> {20000/el:.0f} rows/s; whole batch would take {R/20000*el:.1f} s")
assert bytes(text[: int(roff[20000])].cpu().numpy()).decode().replace("0 0 ", "0 0.0 ", 0) is not None
