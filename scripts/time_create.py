"""Developer tool: time engine creation (landmark tables + candidate grid) on the bench system."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from sitator_b200 import synthetic as syn
from sitator_b200.engine import LandmarkEngine


def main():
    for name in ("llzo", "lgps_dynamic"):
        system, cfg = syn.make_config(name)
        for it in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            eng = LandmarkEngine(system.cell, system.static_idx, system.mobile_idx, system.n_total, system.static_pos,
                                 system.lm_centers, system.lm_vertices)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            print("%s create %.2f ms  grid %s" % (name, (t1 - t0) * 1e3, eng.candidate_grid_info()))
            del eng


if __name__ == "__main__":
    main()
