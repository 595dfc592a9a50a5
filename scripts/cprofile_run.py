"""Developer tool: host-side profile (cProfile) of one warmed-up LandmarkAnalysis.run on the bench workload."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from sitator_b200 import synthetic as syn
from sitator_b200.landmark import LandmarkAnalysis


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    algo = sys.argv[2] if len(sys.argv) > 2 else "mcl"
    system, cfg = syn.make_config("llzo")
    pinned = torch.empty((F, system.n_total, 3), dtype=torch.float64, pin_memory=True)
    frames = pinned.numpy()
    for f0 in range(0, F, 20000):
        n = min(20000, F - f0)
        frames[f0:f0 + n] = system.trajectory(n, seed=f0 // 20000 + system.seed)
    sn = syn.site_network_for(system)
    kw = dict(max_mobile_per_site=max(4, cfg.get("max_mobile_per_site", 1)),
              check_for_zero_landmarks=cfg.get("check_for_zero_landmarks", True))

    def once():
        la = LandmarkAnalysis(clustering_algorithm=algo, verbose=False, **kw)
        return la.run(sn, frames)

    for _ in range(3):
        once()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    once()
    torch.cuda.synchronize()
    print("plain run: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
    pr = cProfile.Profile()
    pr.enable()
    once()
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(45)
    st.sort_stats("tottime").print_stats(25)


if __name__ == "__main__":
    main()
