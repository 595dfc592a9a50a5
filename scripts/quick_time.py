"""Quick device timing of K1 passes on the LLZO-shaped workload (developer tool, not the bench)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from sitator_b200 import synthetic as syn
from tests import _util as U

name = sys.argv[1] if len(sys.argv) > 1 else "llzo"
F = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
system, cfg = syn.make_config(name)
frames = system.trajectory(F)
eng = U.engine_for(system, dynamic_lattice_mapping=cfg["dynamic"])
eng.set_frames(frames)
L, M = system.n_landmarks, system.n_mobile
# realistic centres: a landmark belongs to the nearest true site if its centre is within 2 A of it
d = system.lm_centers[:, None, :] - system.site_pos[None, :, :]
d -= system.lengths * np.round(d / system.lengths)
dist = np.sqrt((d ** 2).sum(-1))
cid = np.where(dist.min(1) < 2.0, dist.argmin(1), -1).astype(np.int32); w = np.ones(L, dtype=np.float32)
NC = len(system.site_pos)
eng.set_centers(cid, w, NC)
N = F * M
labels = torch.empty(N, dtype=torch.int64, device="cuda"); confs = torch.empty(N, dtype=torch.float64, device="cuda")
counts = torch.zeros(NC, dtype=torch.int64, device="cuda")
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)
t, tavg = timeit(lambda: eng.pass_assign(0.7, labels=labels, confs=confs, counts=counts))
print("%s F=%d assign pass: best %.3f ms avg %.3f ms -> %.3e frames/s, %.3e frame*atoms/s" % (name, F, t, tavg, F / t * 1e3, F * system.n_total / t * 1e3))
seen = torch.zeros(L, dtype=torch.int64, device="cuda"); gram = torch.zeros((L, L), dtype=torch.float64, device="cuda")
t, tavg = timeit(lambda: eng.pass_stats(seen=seen, gram=gram), reps=3)
print("%s F=%d stats pass (sparse FP64 gram): best %.3f ms -> %.3e frames/s" % (name, F, t, F / t * 1e3))
print("status:", vars(eng.status()))
