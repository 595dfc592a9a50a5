#!/usr/bin/env python
"""Golden fixture for NON-DEFAULT parameters of the constructor-default 'dotprod' clustering
(`clustering_threshold`, `assignment_threshold`, cluster/dotprod.py:6-9; minimum_site_occupancy): labels, confidences
and the number of sites of the UNMODIFIED compiled reference (oracle/_ref) on the toy and LLZO-shaped golden inputs.

Run in the build container (needs oracle/_ref):  python tests/golden/make_dotprod_params_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader                      # noqa: E402
from sitator_b200 import synthetic as syn          # noqa: E402
from tests import _util as U                       # noqa: E402


def main():
    if not ref_loader.available():
        sys.exit("needs oracle/_ref (python oracle/build_ref.py in a container with /root/reference)")
    ref = ref_loader.load()
    out = {}
    for case, (name, params, min_occ) in U.dotprod_param_cases().items():
        g, system, cfg, frames = U.load_dotprod_golden(name)
        sn = syn.site_network_for(system, ref.SiteNetwork, ref.Atoms)
        la = ref.LandmarkAnalysis(clustering_params=dict(params), verbose=False, force_no_memmap=True,
                                  minimum_site_occupancy=min_occ, **U.analysis_kwargs(cfg))
        st = la.run(sn, frames)
        lv = np.asarray(la.landmark_vectors)
        confs = st.confidences.copy()
        confs.reshape(-1)[~lv.any(axis=1)] = 0.0          # uninitialised in the reference for all-zero rows
        out[case + "/labels"] = st.traj
        out[case + "/confs"] = confs
        out[case + "/site_centers"] = np.asarray(st.site_network.centers)
        print("%s: %d sites (default parameters: %d), %d unassigned"
              % (case, st.site_network.n_sites, len(g["site_centers"]), int(np.sum(st.traj < 0))))
    np.savez_compressed(os.path.join(HERE, "dotprod_params.npz"), **out)


if __name__ == "__main__":
    main()
