"""The drop-in boundary, executed: the COMPILED REFERENCE's own LandmarkAnalysis.run (oracle/_ref) with its Cython fill
replaced by the C-ABI binding INTEGRATION.md quotes (sitator_b200/integration/reference_binding.py) must return the
golden SiteTrajectory; and a clustering plugin written for the reference's (N, L)-ndarray contract
(LandmarkAnalysis.py:234-256) -- the reference's own mcl plugin -- must run behind sitator_b200's LandmarkAnalysis."""
import numpy as np
import pytest

from sitator_b200 import synthetic as syn
from . import _util as U

pytestmark = pytest.mark.gpu


def _ref():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref (the compiled reference) is not built")
    return ref_loader.load()


@pytest.mark.parametrize("name", ["toy_bcc_300", "lgps_dynamic_40"])
def test_reference_run_with_the_c_abi_fill_reproduces_the_golden(name):
    ref = _ref()
    from sitator_b200.integration import reference_binding as rb
    g, system, cfg, frames = U.load_golden(name)
    sn = syn.site_network_for(system, ref.SiteNetwork, ref.Atoms)
    la = ref.LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, force_no_memmap=True, **U.analysis_kwargs(cfg))
    orig = ref.helpers._fill_landmark_vectors
    ref.helpers._fill_landmark_vectors = rb._fill_landmark_vectors
    try:
        st = la.run(sn, frames)
    finally:
        ref.helpers._fill_landmark_vectors = orig
    lv = np.asarray(la.landmark_vectors)
    want = g["landmark_vectors"]
    assert np.array_equal(lv != 0, want != 0)
    nz = want != 0
    assert np.max(np.abs(lv[nz] - want[nz]) / want[nz]) < U.LV_RTOL
    assert la.n_all_zero_lvecs == int(g["n_all_zero_lvecs"])
    # the reference's clustering ran on those vectors: same sites, labels, confidences
    assert np.array_equal(np.asarray(st.traj), g["labels"])
    known = g["labels"] >= 0
    assert np.max(np.abs(np.asarray(st.confidences)[known] - g["confs"][known])) < 1e-10
    assert np.max(np.abs(np.asarray(st.site_network.centers) - g["site_centers"])) < 1e-8


def test_reference_error_types_through_the_binding():
    ref = _ref()
    from sitator_b200.integration import reference_binding as rb
    system, frames, kw = U.error_cases()["static_moved"]
    sn = syn.site_network_for(system, ref.SiteNetwork, ref.Atoms)
    la = ref.LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, force_no_memmap=True, **kw)
    orig = ref.helpers._fill_landmark_vectors
    ref.helpers._fill_landmark_vectors = rb._fill_landmark_vectors
    try:
        with pytest.raises(ref.landmark_errors.StaticLatticeError) as ei:
            la.run(sn, frames)
    finally:
        ref.helpers._fill_landmark_vectors = orig
    assert ei.value.frame == 9 and list(ei.value.lattice_atoms) == [2]


def test_reference_contract_plugin_runs_behind_our_landmark_analysis():
    """clustering_algorithm given as a dotted module path: a plugin with the reference's ndarray contract."""
    ref = _ref()
    from sitator_b200.landmark import LandmarkAnalysis
    g, system, cfg, frames = U.load_golden("toy_bcc_300")
    la = LandmarkAnalysis(clustering_algorithm='sitator.landmark.cluster.mcl', verbose=False, **U.analysis_kwargs(cfg))
    st = la.run(syn.site_network_for(system), frames)
    assert la.stats["plugin_contract"] == "ndarray"
    assert np.array_equal(st.traj, g["labels"])
    known = g["labels"] >= 0
    assert np.max(np.abs(st.confidences[known] - g["confs"][known])) < 1e-10
    assert np.max(np.abs(np.asarray(st.site_network.centers) - g["site_centers"])) < 1e-8
