#!/usr/bin/env python
"""Golden fixture for the non-default `site_centers_method`s (LandmarkAnalysis.py:286-296): site centres of the
UNMODIFIED compiled reference (oracle/_ref) with 'real-unweighted' and 'representative-landmark' on the toy and the
LLZO-shaped golden inputs (the run() goldens hold the default 'real-weighted').

Run in the build container (needs oracle/_ref):  python tests/golden/make_site_center_golden.py
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

CASES = ["toy_bcc_300", "llzo_60"]
METHODS = ["real-unweighted", "representative-landmark"]


def main():
    if not ref_loader.available():
        sys.exit("needs oracle/_ref (python oracle/build_ref.py in a container with /root/reference)")
    ref = ref_loader.load()
    out = {}
    for name in CASES:
        g, system, cfg, frames = U.load_golden(name)
        for method in METHODS:
            sn = syn.site_network_for(system, ref.SiteNetwork, ref.Atoms)
            la = ref.LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, force_no_memmap=True,
                                      site_centers_method=method, **U.analysis_kwargs(cfg))
            st = la.run(sn, frames)
            assert np.array_equal(st.traj, g["labels"])            # the method changes the centres only
            out["%s/%s" % (name, method)] = np.asarray(st.site_network.centers)
            print(name, method, out["%s/%s" % (name, method)].shape)
    np.savez_compressed(os.path.join(HERE, "site_center_methods.npz"), **out)


if __name__ == "__main__":
    main()
