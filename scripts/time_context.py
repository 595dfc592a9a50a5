"""Host wall time of building the analysis context (tables + candidate grid) for one BASELINE shape (developer tool).
    SITB_TIMING=1 python scripts/time_context.py [llzo]
Prints the Python-side split (site-network read-out, vertex table, native create) and, with SITB_TIMING=1, the native stages."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from sitator_b200 import synthetic as syn
from sitator_b200.engine import LandmarkEngine, vertex_table

name = sys.argv[1] if len(sys.argv) > 1 else "llzo"
system, cfg = syn.make_config(name)
sn = syn.site_network_for(system)
torch.cuda.init()
for it in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cell = np.asarray(sn.structure.cell)
    static_idx = np.where(sn.static_mask)[0]; mobile_idx = np.where(sn.mobile_mask)[0]
    pos = sn.static_structure.get_positions(); centers = np.asarray(sn.centers); verts = sn.vertices
    t1 = time.perf_counter()
    vt = vertex_table(verts)
    t2 = time.perf_counter()
    if it == 3:
        print("--- native stages of the last build ---", file=sys.stderr, flush=True)
    eng = LandmarkEngine(cell, static_idx, mobile_idx, sn.n_total, pos, centers, verts)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print("build %d: site network read-out %.3f ms, vertex table %.3f ms, LandmarkEngine() %.3f ms" %
          (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
    eng.close()
