"""Developer tool: time the 'dotprod' clustering (sequential fit + predict passes) on the LLZO-shaped workload."""
import sys, time, logging
logging.basicConfig(level=logging.DEBUG)
logging.getLogger("sitator_b200.landmark.LandmarkAnalysis").setLevel(logging.WARNING)
import numpy as np, torch
sys.path.insert(0, ".")
from sitator_b200 import synthetic as syn
from sitator_b200.landmark import LandmarkAnalysis
from sitator_b200.landmark.source import LandmarkVectorSource
from sitator_b200.landmark.cluster import dotprod
from tests import _util as U
F = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
system, cfg = syn.make_config("llzo")
frames = system.trajectory(F)
eng = U.engine_for(system); eng.set_frames(frames)
src = LandmarkVectorSource(eng)
torch.cuda.synchronize(); t = time.perf_counter(); dotprod.first_pass(src); torch.cuda.synchronize()
print("first pass %.1f ms" % ((time.perf_counter() - t) * 1e3))
t = time.perf_counter(); c, n = dotprod.fit_centers(src, 0.45); torch.cuda.synchronize()
dt = time.perf_counter() - t
print("fit_centers: %d rows -> %d centres in %.1f ms (%.3f us/row)" % (src.n_local, len(c), dt * 1e3, dt / src.n_local * 1e6))
t = time.perf_counter(); l, cf, cnt = dotprod._predict(src, c, 0.8, True); torch.cuda.synchronize()
print("predict: %.2f ms, assigned %d" % ((time.perf_counter() - t) * 1e3, int(cnt.sum())))
eng.close()
for rep in range(0):
    la = LandmarkAnalysis(verbose=False, **U.analysis_kwargs(cfg))
    t = time.perf_counter(); st = la.run(syn.site_network_for(system), frames)
    print("run(dotprod) %.1f ms, %d sites, unassigned %.3f" % ((time.perf_counter() - t) * 1e3, st.site_network.n_sites, float(np.mean(st.traj < 0))))
