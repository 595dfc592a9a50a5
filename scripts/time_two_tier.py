"""Two-tier fused fill + assign pass against the exact pass on a BASELINE shape (developer tool): identical labels,
confidence error, rows left to the exact kernel, device time of both."""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from sitator_b200 import synthetic as syn
from sitator_b200.landmark import LandmarkAnalysis
from tests import _util as U

name = sys.argv[1] if len(sys.argv) > 1 else "llzo"
F = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
system, cfg = syn.make_config(name)
frames = system.trajectory(F)
kw = U.analysis_kwargs(cfg)
kw["max_mobile_per_site"] = max(4, kw["max_mobile_per_site"])
la = LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, **kw)
st = la.run(syn.site_network_for(system), frames)
eng = la._engine
M = system.n_mobile
N = F * M
out = {}
res = {}
for mode in ("exact", "two_tier"):
    eng.set_assign_mode(mode)
    labels = torch.empty(N, dtype=torch.int64, device="cuda"); confs = torch.empty(N, dtype=torch.float64, device="cuda")
    counts = torch.zeros(eng.n_clusters, dtype=torch.int64, device="cuda")
    eng.two_tier_info(reset=True)
    eng.pass_assign(0.7, labels=labels, confs=confs, counts=counts)
    torch.cuda.synchronize()
    info = eng.two_tier_info(reset=True)
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); eng.pass_assign(0.7, labels=labels, confs=confs); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    res[mode] = (labels.cpu().numpy(), confs.cpu().numpy(), counts.cpu().numpy())
    out[mode] = {"ms_best": min(ts), "ms_mean": float(np.mean(ts)), "frame_atoms_per_s": F * system.n_total / min(ts) * 1e3, **info}
le, ce, ne = res["exact"]; lf, cf, nf = res["two_tier"]
out["labels_equal"] = bool(np.array_equal(le, lf)); out["n_label_diff"] = int((le != lf).sum())
out["counts_equal"] = bool(np.array_equal(ne, nf))
out["labels_equal_run"] = bool(np.array_equal(le.reshape(F, M), st.traj))
out["conf_max_abs_err"] = float(np.max(np.abs(ce - cf))); out["rows"] = N
out["shape"] = dict(name=name, F=F, S=system.n_static, M=M, L=system.n_landmarks, n_clusters=int(eng.n_clusters))
print(json.dumps(out, indent=1))
