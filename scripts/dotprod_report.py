"""Developer tool: the constructor-default 'dotprod' clustering on the LLZO-shaped workload -> gpurun_out/dotprod.json
(fit / predict timings, the fit kernel's per-phase cycle counters, whole run())."""
import json, sys, time, logging
import numpy as np, torch
sys.path.insert(0, ".")
from sitator_b200 import synthetic as syn
from sitator_b200.landmark import LandmarkAnalysis
from sitator_b200.landmark.source import LandmarkVectorSource
from sitator_b200.landmark.cluster import dotprod
from tests import _util as U

class Grab(logging.Handler):
    def __init__(self): super().__init__(); self.lines = []
    def emit(self, rec): self.lines.append(rec.getMessage())
grab = Grab(); lg = logging.getLogger("sitator_b200.landmark.cluster.dotprod"); lg.setLevel(logging.DEBUG); lg.addHandler(grab)
F = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
system, cfg = syn.make_config("llzo")
frames = system.trajectory(F)
eng = U.engine_for(system); eng.set_frames(frames)
src = LandmarkVectorSource(eng)
rec = {"config": "llzo", "frames": F, "rows": F * system.n_mobile}
torch.cuda.synchronize(); t = time.perf_counter(); dotprod.first_pass(src); torch.cuda.synchronize()
rec["first_pass_ms"] = (time.perf_counter() - t) * 1e3
t = time.perf_counter(); c, n = dotprod.fit_centers(src, 0.45); torch.cuda.synchronize()
rec["fit_centers_ms"] = (time.perf_counter() - t) * 1e3
rec["fit_us_per_row"] = rec["fit_centers_ms"] * 1e3 / rec["rows"]
rec["centres"] = int(len(c))
t = time.perf_counter(); l, cf, cnt = dotprod._predict(src, c, 0.8, True); torch.cuda.synchronize()
rec["predict_ms"] = (time.perf_counter() - t) * 1e3
rec["fit_log"] = [x for x in grab.lines if x.startswith("dotprod fit")]
eng.close()
la = LandmarkAnalysis(verbose=False, **U.analysis_kwargs(cfg))
t = time.perf_counter(); st = la.run(syn.site_network_for(system), frames)
rec["run_ms"] = (time.perf_counter() - t) * 1e3
rec["n_sites"] = int(st.site_network.n_sites); rec["unassigned_frac"] = float(np.mean(st.traj < 0))
print(json.dumps(rec)); json.dump(rec, open("gpurun_out/dotprod.json", "w"), indent=1)
