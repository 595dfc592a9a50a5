#!/usr/bin/env python
"""Golden fixture for the ERROR behaviour of the path: which exception the UNMODIFIED compiled reference (oracle/_ref)
raises, with which attributes, on seeded inputs that break the static lattice, produce an all-zero landmark vector or
put two atoms on one site.  The oracle's failure classes are pinned to these (tests/test_oracle_pinning.py); the GPU
path is then checked against the oracle.

Run in the build container (needs oracle/_ref):  python tests/golden/make_error_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader                      # noqa: E402
from sitator_b200 import synthetic as syn          # noqa: E402
from tests import _util as U                       # noqa: E402


def _intlist(x):
    return None if x is None else [int(v) for v in np.asarray(x).reshape(-1)]


def main():
    if not ref_loader.available():
        sys.exit("needs oracle/_ref (python oracle/build_ref.py in a container with /root/reference)")
    ref = ref_loader.load()
    out = {}
    for name, (system, frames, kw) in U.error_cases().items():
        sn = syn.site_network_for(system, ref.SiteNetwork, ref.Atoms)
        la = ref.LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, force_no_memmap=True, **kw)
        try:
            la.run(sn, frames)
            out[name] = {"error": None}
        except ref.landmark_errors.StaticLatticeError as e:
            out[name] = {"error": "StaticLatticeError", "frame": int(e.frame), "lattice_atoms": _intlist(e.lattice_atoms)}
        except ref.landmark_errors.ZeroLandmarkError as e:
            out[name] = {"error": "ZeroLandmarkError", "frame": int(e.frame), "mobile_index": int(e.mobile_index)}
    # SiteTrajectory.check_multiple_occupancy on a random assignment table
    traj, n_sites = U.occupancy_error_table()
    system, _ = syn.make_config("toy_bcc")
    full = syn.site_network_for(system, ref.SiteNetwork, ref.Atoms)
    sn = ref.SiteNetwork(full.structure, full.static_mask, full.mobile_mask)
    sn.centers = np.zeros((n_sites, 3))
    st = ref.SiteTrajectory(sn, traj)
    try:
        st.check_multiple_occupancy(max_mobile_per_site=1)
        out["multiple_occupancy"] = {"error": None}
    except ref.errors.MultipleOccupancyError as e:
        out["multiple_occupancy"] = {"error": "MultipleOccupancyError", "frame": int(e.frame), "site": int(e.site),
                                     "mobile_particles": _intlist(e.mobile_particles)}
    n_more, avg = st.check_multiple_occupancy(max_mobile_per_site=traj.shape[1])
    out["multiple_occupancy_stats"] = {"n_more_than_one": int(n_more), "avg_mobile_per_site": float(avg)}
    with open(os.path.join(HERE, "errors.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
