"""A/B of k_time_state variants in an A/B build (-DVAP_TS_V1): per-stage CUDA-event times of the eager step.
usage: python profiles/tools/ab_time_state.py   (VAP_TS_V1=1 in the environment selects the round-1 kernel per launch)"""
import os, sys, time, torch
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
    t0 = time.perf_counter()
    for _ in range(reps):
        res = eng.profile(db, reuse_plan=True)
    torch.cuda.synchronize()
    el = (time.perf_counter() - t0) / reps
    st = {}
    for name, s, e in eng.stage_events:
        st[name] = st.get(name, 0.0) + s.elapsed_time(e) / reps
    print(f"{tag}: {el*1e3:.3f} ms/step", {k: round(v, 3) for k, v in st.items()}, flush=True)
    del eng, db, res
    torch.cuda.empty_cache()


cases = [("cfg2 4096x8", synth.random_paths(4096, 8, seed=1)), ("mixed 4096x8", synth.mixed_paths(4096, 8, seed=3)),
         ("8192x16", synth.random_paths(8192, 16, seed=1)), ("cfg4 1x801", synth.long_path(801, seed=2))]
for name, packed in cases:
    for v1, lanes in ((False, None), (True, None), (False, 3), (False, 4), (False, 5), (False, 10)):
        if v1:
            os.environ["VAP_TS_V1"] = "1"
        else:
            os.environ.pop("VAP_TS_V1", None)
        if lanes:
            os.environ["VAP_STATE_LANES"] = str(lanes)
        else:
            os.environ.pop("VAP_STATE_LANES", None)
        if lanes and name.startswith("cfg4"):
            continue
        run(f"{name} {'v1' if v1 else 'new'} lanes={lanes}", packed)
