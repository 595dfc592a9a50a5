"""Kernel time of the assign passes over the cached compressed rows (developer tool): pass B (best row per cluster, no
outputs per row) and pass C (labels, confidences, counts, representative vectors, per-site best row) of the mcl plugin,
with the centres of a real run.
    python scripts/time_assign_sparse.py [frames]"""
import json, sys
sys.path.insert(0, ".")
import numpy as np, torch
from sitator_b200 import synthetic as syn
from sitator_b200.landmark import LandmarkAnalysis
from sitator_b200.landmark.source import LandmarkVectorSource
from sitator_b200.engine import new_best_table

F = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
system, cfg = syn.make_config("llzo")
frames = np.concatenate([system.trajectory(min(20000, F - f0), seed=f0 // 20000 + system.seed) for f0 in range(0, F, 20000)])
la = LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, max_mobile_per_site=4, check_for_zero_landmarks=True)
sn = syn.site_network_for(system)
la.run(sn, frames)
eng = la._engine
cid, w = la.cluster_centers_ if getattr(la, "cluster_centers_", None) is not None else (None, None)
src = LandmarkVectorSource(eng, None)
seen, gram, rows = eng.pass_stats_cached(gram_words=True)
# centres: one cluster per 16 consecutive landmarks unless the run kept its own
L = eng.L
if cid is None:
    cid = (np.arange(L) // 16).astype(np.int32); w = np.full(L, 0.25)
C = int(cid.max()) + 1
eng.set_centers(cid, w, C)
N = rows.n_rows


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        fn()
    t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps


best = new_best_table(C, eng.device)
labels = torch.empty((N,), dtype=torch.int64, device=eng.device); confs = torch.empty((N,), dtype=torch.float64, device=eng.device)
counts = torch.zeros((C,), dtype=torch.int64, device=eng.device)
rep = torch.zeros((C, L), dtype=torch.float64, device=eng.device); rep_w = torch.zeros((C,), dtype=torch.float64, device=eng.device)
site_best = new_best_table(C, eng.device)
out = {"frames": F, "rows": N, "clusters": C, "entries": int(rows.used),
       "pass_B_ms": timed(lambda: eng.assign_sparse(rows, float('nan'), best=best)),
       "pass_C_ms": timed(lambda: eng.assign_sparse(rows, 0.7, labels=labels, confs=confs, counts=counts, rep=rep, rep_w=rep_w,
                                                    site_best=site_best)),
       "labels_only_ms": timed(lambda: eng.assign_sparse(rows, 0.7, labels=labels, confs=confs))}
out["bytes_per_pass"] = int(rows.used) * 10 + N * 8
out["pass_B_GBps"] = out["bytes_per_pass"] / out["pass_B_ms"] / 1e6
print(json.dumps(out, indent=1))
