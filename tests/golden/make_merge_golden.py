#!/usr/bin/env python
"""Golden outputs of the compiled reference's site merging:
``MergeSitesByDynamics._get_sites_to_merge`` (dynamics/MergeSitesByDynamics.py:110-153) and ``MergeSites.run``
(network/merging.py:44-133) on the SiteTrajectory of the toy and LLZO goldens.

The reference class cannot be constructed (``self.iterlimit = iterlimit`` is a NameError,
MergeSitesByDynamics.py:54), so the instance is made with ``__new__`` and given the attributes its ``__init__`` would set.
Run in the build container:  python tests/golden/make_merge_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import build_ref, ref_loader          # noqa: E402
from sitator_b200 import synthetic as syn          # noqa: E402
from tests import _util as U                       # noqa: E402

CASES = U.MERGE_CASES


def reference_merger(ref, distance_threshold, factor, markov_parameters, weighted):
    m = ref.MergeSitesByDynamics.__new__(ref.MergeSitesByDynamics)
    # MergeSites.__init__ (network/merging.py:33-41)
    m.check_types = False
    m.maximum_merge_distance = factor * distance_threshold
    m.set_merged_into = True
    m.weighted_spatial_average = weighted
    # MergeSitesByDynamics.__init__ (dynamics/MergeSitesByDynamics.py:36-56)
    m.connectivity_matrix_generator = ref.MergeSitesByDynamics.connectivity_n_ij
    m.distance_threshold = distance_threshold
    m.post_check_thresh_factor = factor
    m.markov_parameters = dict(markov_parameters)
    return m


def main():
    if not build_ref.build(verbose=False):
        sys.exit("needs /root/reference to build oracle/_ref")
    ref = ref_loader.load()
    out = {}
    for case, (gname, thr, factor, mp, weighted) in CASES.items():
        system, frames, labels, confs, centers, verts = U.merge_case_inputs(gname)
        sn = syn.site_network_for(system, ref.SiteNetwork, ref.Atoms)
        sn.centers = centers
        ov = np.empty(len(verts), dtype=object)          # object rows: ragged-safe in the reference's sn.copy() (SiteNetwork.py:110)
        for i, v in enumerate(verts):
            ov[i] = list(v)
        sn.vertices = ov
        st = ref.SiteTrajectory(sn, labels.copy(), confidences=confs.copy())
        if not weighted:
            st.compute_site_occupancies()
        merger = reference_merger(ref, thr, factor, mp, weighted)
        clusters = merger._get_sites_to_merge(st)
        merger2 = reference_merger(ref, thr, factor, mp, weighted)
        new = merger2.run(st)
        out[case + "/clusters"] = np.concatenate([np.asarray(sorted(c), dtype=np.int64) for c in clusters])
        out[case + "/cluster_len"] = np.array([len(c) for c in clusters])
        out[case + "/traj"] = np.asarray(new.traj)
        out[case + "/centers"] = np.asarray(new.site_network.centers)
        out[case + "/verts"] = np.concatenate([np.asarray(sorted(v), dtype=np.int64) for v in new.site_network.vertices])
        out[case + "/verts_len"] = np.array([len(v) for v in new.site_network.vertices])
        out[case + "/merged_into"] = np.asarray(st.site_network.merged_into)
        print("%s: %d -> %d sites" % (case, sn.n_sites, new.site_network.n_sites))
    np.savez_compressed(os.path.join(HERE, "merge_sites.npz"), **out)


if __name__ == "__main__":
    main()
