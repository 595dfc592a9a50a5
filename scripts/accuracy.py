"""Developer diagnostic: accuracy of the GPU landmark vectors / confidences against a golden case."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from tests import _util as U
from sitator_b200 import synthetic as syn
from sitator_b200.landmark import LandmarkAnalysis
for name in U.GOLDEN_CASES:
    g, system, cfg, frames = U.load_golden(name)
    la = LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, **U.analysis_kwargs(cfg))
    st = la.run(syn.site_network_for(system), frames)
    lv = np.asarray(la.landmark_vectors); want = g["landmark_vectors"]; nz = want != 0
    print(name, "support equal", np.array_equal(lv != 0, nz), "lv max rel", np.max(np.abs(lv[nz] - want[nz]) / want[nz]),
          "sites", st.site_network.n_sites, len(g["site_centers"]), "labels differ", int(np.sum(st.traj != g["labels"])),
          "conf max abs", np.max(np.abs(st.confidences - g["confs"])) if st.traj.shape == g["labels"].shape else None,
          "centers max", np.max(np.abs(np.asarray(st.site_network.centers) - g["site_centers"])) if st.site_network.n_sites == len(g["site_centers"]) else None,
          "ms", la.stats["run_ms"])
