"""Site merging: sitator_b200.dynamics.MergeSitesByDynamics / network.MergeSites against golden
outputs of the compiled reference (tests/golden/merge_sites.npz, made by make_merge_golden.py from
dynamics/MergeSitesByDynamics.py:110-153 and network/merging.py:44-133 with the unconstructible __init__ bypassed)."""
import os

import numpy as np
import pytest

from sitator_b200 import synthetic as syn
from . import _util as U

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", sorted(U.MERGE_CASES))
def test_merge_sites_by_dynamics_matches_reference(case):
    from sitator_b200 import SiteTrajectory
    from sitator_b200.dynamics import MergeSitesByDynamics
    g = np.load(os.path.join(U.GOLDEN_DIR, "merge_sites.npz"))
    gname, thr, factor, mp, weighted = U.MERGE_CASES[case]
    system, frames, labels, confs, centers, verts = U.merge_case_inputs(gname)
    sn = syn.site_network_for(system)
    sn.centers = centers
    sn.vertices = verts
    st = SiteTrajectory(sn, labels.copy(), confs.copy())
    if not weighted:
        st.compute_site_occupancies()
    merger = MergeSitesByDynamics(distance_threshold=thr, post_check_thresh_factor=factor, check_types=False,
                                  markov_parameters=mp, set_merged_into=True, weighted_spatial_average=weighted)
    clusters = merger._get_sites_to_merge(st)
    lens = g[case + "/cluster_len"]
    ends = np.cumsum(lens)
    want_clusters = [tuple(int(x) for x in g[case + "/clusters"][e - n:e]) for e, n in zip(ends, lens)]
    assert [tuple(sorted(int(x) for x in c)) for c in clusters] == want_clusters      # same groups in the same order
    new = MergeSitesByDynamics(distance_threshold=thr, post_check_thresh_factor=factor, check_types=False,
                               markov_parameters=mp, set_merged_into=True, weighted_spatial_average=weighted).run(st)
    assert np.array_equal(new.traj, g[case + "/traj"])
    assert new.confidences is None or not np.any(new.confidences)                   # merging.py:127-129 drops them
    assert np.max(np.abs(np.asarray(new.site_network.centers) - g[case + "/centers"])) < 1e-10
    vl = g[case + "/verts_len"]
    ve = np.cumsum(vl)
    want_verts = [set(int(x) for x in g[case + "/verts"][e - n:e]) for e, n in zip(ve, vl)]
    assert [set(v) for v in new.site_network.vertices] == want_verts
    assert np.array_equal(np.asarray(st.site_network.merged_into), g[case + "/merged_into"])


def test_merged_sites_too_distant_and_constructor():
    from sitator_b200 import SiteTrajectory
    from sitator_b200.dynamics import MergeSitesByDynamics
    from sitator_b200.network import MergedSitesTooDistantError
    m = MergeSitesByDynamics(iterlimit=50)                 # the reference's constructor raises NameError here
    assert m.iterlimit == 50 and m.maximum_merge_distance == 1.5
    system, frames, labels, confs, centers, verts = U.merge_case_inputs("toy_bcc_2000+flicker")
    sn = syn.site_network_for(system)
    sn.centers = centers
    sn.vertices = verts
    st = SiteTrajectory(sn, labels.copy(), confs.copy())
    with pytest.raises(MergedSitesTooDistantError):
        MergeSitesByDynamics(distance_threshold=3.0, post_check_thresh_factor=0.3, check_types=False).run(st)
