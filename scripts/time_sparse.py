"""Developer tool: time the sparse assign kernel with each optional output on its own."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from sitator_b200 import synthetic as syn
from sitator_b200.engine import new_best_table
from tests import _util as U
F = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
system, cfg = syn.make_config("llzo")
frames = np.concatenate([system.trajectory(20000, seed=i + 1) for i in range(F // 20000)])
eng = U.engine_for(system); eng.set_frames(frames)
L, M = system.n_landmarks, system.n_mobile
d = system.lm_centers[:, None, :] - system.site_pos[None, :, :]; d -= system.lengths * np.round(d / system.lengths)
dist = np.sqrt((d ** 2).sum(-1)); cid = np.where(dist.min(1) < 2.0, dist.argmin(1), -1).astype(np.int32)
NC = len(system.site_pos); eng.set_centers(cid, np.ones(L), NC)
seen, gram, sp = eng.pass_stats_cached()
N = F * M
labels = torch.empty(N, dtype=torch.int64, device="cuda"); confs = torch.empty(N, dtype=torch.float64, device="cuda")
def t(name, **kw):
    torch.cuda.synchronize(); ts = []
    for _ in range(4):
        kw2 = {k: (v() if callable(v) else v) for k, v in kw.items()}
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); eng.assign_sparse(sp, 0.7, **kw2); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print("%-28s %.3f ms" % (name, min(ts)))
Z = lambda *s: (lambda: torch.zeros(s, dtype=torch.int64, device="cuda"))
Zf = lambda *s: (lambda: torch.zeros(s, dtype=torch.float64, device="cuda"))
t("labels+confs", labels=labels, confs=confs)
t("counts", counts=Z(NC))
t("best", best=lambda: new_best_table(NC, "cuda"))
t("site_best (+labels)", labels=labels, site_best=lambda: new_best_table(NC, "cuda"))
t("rep+rep_w", rep=Zf(NC, L), rep_w=Zf(NC))
t("all of pass D", labels=labels, confs=confs, rep=Zf(NC, L), rep_w=Zf(NC), site_best=lambda: new_best_table(NC, "cuda"))
print("pool entries", sp.used, "bytes/row", (sp.used * 10 + 8 * N) / N)
