"""CPU: the Zeo++-free VoronoiSiteGenerator (sitator_b200/voronoi.py; reference voronoi.py:23-41 + util/zeo.py:46-108).
No parity fixture exists (Zeo++ is not installable here); the defining property of a Voronoi node is checked instead."""
import numpy as np

from sitator_b200 import synthetic as syn
from sitator_b200.voronoi import VoronoiSiteGenerator


def _min_image_dists(cell, p, pts):
    inv = np.linalg.inv(cell)
    d = (pts - p) @ inv
    d -= np.round(d)
    best = np.full(len(pts), np.inf)
    for s in np.stack(np.meshgrid(*[np.array([-1, 0, 1])] * 3, indexing="ij"), -1).reshape(-1, 3):
        best = np.minimum(best, np.linalg.norm((d + s) @ cell, axis=1))
    return best


def test_nodes_are_equidistant_from_their_vertex_atoms():
    system, cfg = syn.make_config("llzo_v4")
    sn = syn.site_network_for(system)
    out = VoronoiSiteGenerator().run(sn)
    assert out.n_sites > 3 * system.n_static
    assert np.array_equal(out.static_mask, sn.static_mask)
    static = np.asarray(sn.static_structure.get_positions())
    rng = np.random.default_rng(0)
    for i in rng.permutation(out.n_sites)[:200]:
        d = _min_image_dists(system.cell, out.centers[i], static)
        vd = d[list(out.vertices[i])]
        assert len(out.vertices[i]) >= 4
        assert np.ptp(vd) < 1e-7                      # equidistant from its generating atoms
        assert d.min() > vd.min() - 1e-7              # and no atom is closer
    # the synthetic LLZO basis was built the same way: its 4-generator nodes are among the generator's
    ours = {tuple(sorted(v)) for v in out.vertices if len(v) == 4}
    assert {tuple(sorted(v)) for v in system.lm_vertices if len(v) == 4} <= ours


def test_triclinic_cell_and_options():
    import pytest
    from sitator_b200 import SiteNetwork, Atoms
    rng = np.random.default_rng(4)
    cell = np.array([[8.0, 0.4, 0.0], [1.1, 7.5, 0.3], [0.2, -0.6, 9.0]])
    pos = rng.random((40, 3)) @ cell
    sn = SiteNetwork(Atoms(positions=pos, cell=cell, numbers=np.full(40, 8)), np.arange(40) < 34, np.arange(40) >= 34)
    out = VoronoiSiteGenerator(zeopp_path="ignored").run(sn)
    static = pos[:34]
    for i in range(0, out.n_sites, 7):
        d = _min_image_dists(cell, out.centers[i], static)
        assert np.ptp(d[list(out.vertices[i])]) < 1e-7 and d.min() > d[list(out.vertices[i])].min() - 1e-7
    with pytest.raises(NotImplementedError):
        VoronoiSiteGenerator(radial=True)
