"""Time the eager stages with several builds of libvap.so (one subprocess per library: VAP_LIB_PATH).
usage: python profiles/tools/ab_libs.py lib1.so lib2.so ..."""
import os, subprocess, sys
child = r'''
import sys, time, torch
sys.path.insert(0, '.')
from vexautonomousplanner_b200 import synth
from vexautonomousplanner_b200.engine import Engine
def run(tag, packed, reps=5):
    eng = Engine("cuda:0")
    db = eng.upload(packed)
    for _ in range(2):
        res = eng.profile(db, reuse_plan=True)
    torch.cuda.synchronize()
    eng.stage_events = []
    for _ in range(reps):
        res = eng.profile(db, reuse_plan=True)
    torch.cuda.synchronize()
    st = {}
    for name, s, e in eng.stage_events:
        st[name] = st.get(name, 0.0) + s.elapsed_time(e) / reps
    print(f"{tag}: " + "  ".join(f"{k.split('_')[0]} {v:.3f}" for k, v in st.items()) + f"  total {sum(st.values()):.3f} ms  n_out_sum={int(res.n_out.sum())}", flush=True)
for name, packed in (("cfg2 4096x8", synth.random_paths(4096, 8, seed=1)), ("8192x16", synth.random_paths(8192, 16, seed=1)),
                     ("mixed 4096x8", synth.mixed_paths(4096, 8, seed=3)), ("cfg4 1x801", synth.long_path(801, seed=2))):
    run(name, packed)
'''
for lib in sys.argv[1:]:
    env = dict(os.environ, VAP_LIB_PATH=os.path.abspath(lib))
    print("==", lib, flush=True)
    subprocess.run([sys.executable, "-c", child], env=env, timeout=200)
