"""Developer tool: cProfile of a warm LandmarkAnalysis.run on the bench workload."""
import cProfile, pstats, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from sitator_b200 import synthetic as syn
from sitator_b200.landmark import LandmarkAnalysis

F = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
system, cfg = syn.make_config("llzo")
pinned = torch.empty((F, system.n_total, 3), dtype=torch.float64, pin_memory=True)
frames = pinned.numpy()
for f0 in range(0, F, 20000):
    n = min(20000, F - f0); frames[f0:f0 + n] = system.trajectory(n, seed=f0 // 20000 + 1)
sn = syn.site_network_for(system)
kw = dict(cfg.get("analysis", {}))
def go():
    la = LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, dynamic_lattice_mapping=cfg["dynamic"],
                          max_mobile_per_site=cfg.get("max_mobile_per_site", 1),
                          check_for_zero_landmarks=cfg.get("check_for_zero_landmarks", True))
    t = time.perf_counter(); st = la.run(sn, frames); torch.cuda.synchronize()
    return (time.perf_counter() - t) * 1e3
for i in range(3): print("warm run %d: %.1f ms" % (i, go()))
pr = cProfile.Profile(); pr.enable(); ms = go(); pr.disable()
print("profiled run: %.1f ms" % ms)
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
