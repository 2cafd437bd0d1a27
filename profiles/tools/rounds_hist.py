import sys, torch, numpy as np
sys.path.insert(0, '.')
from vexautonomousplanner_b200 import synth
from vexautonomousplanner_b200.engine import Engine
eng = Engine("cuda:0")
db = eng.upload(synth.random_paths(4096, 8, seed=0))
res = eng.profile(db, keep=True)
r = res.extra["rounds"].cpu().numpy(); D = res.n_samples.cpu().numpy(); T = res.n_out.cpu().numpy()
for name, col in (("fwd", 0), ("bwd", 1)):
    print(name, "rounds hist:", np.bincount(r[:, col]).tolist())
print("D min/mean/max", D.min(), D.mean(), D.max(), " T min/mean/max", T.min(), T.mean(), T.max())
work = (D / 32) * (1 + r[:, 0])
print("fwd upper-bound steps per path (Lc*(1+rounds)): mean %.0f p90 %.0f p99 %.0f max %.0f" % (work.mean(), np.percentile(work, 90), np.percentile(work, 99), work.max()))
