"""``MergeSites``: base class of the site-merging post-processors (mirrors reference
``sitator/network/merging.py:19-145``).

A subclass names groups of sites (``_get_sites_to_merge``); ``run`` builds the merged ``SiteNetwork`` (periodic
average of the member centres, union of their vertices) and the relabelled ``SiteTrajectory``.  The arithmetic over
sites and over the (frames x mobile) assignment stream runs on the device: pairwise periodic distances and averages
(``sitb_pbc_distances`` / ``sitb_pbc_weighted_average``, the reference's ``PBCCalculator.distances`` / ``.average``)
and the relabelling of the stream (``sitb_relabel_sites``).

Kept as in the reference, on purpose: with ``weighted_spatial_average=True`` (the default) the merged centre is the
UNWEIGHTED average, with ``False`` it is weighted by ``occupancies`` (``merging.py:101-105`` has the two branches
that way round); confidences are dropped (``:127-129``).
"""
import abc
import ctypes as C
import logging

import numpy as np

from .. import _native
from ..SiteTrajectory import SiteTrajectory
from ..errors import InsufficientSitesError

logger = logging.getLogger(__name__)


class MergeSitesError(Exception):
    pass


class MergedSitesTooDistantError(MergeSitesError):
    pass


def _cell_matrices(cell):
    cellmat = np.ascontiguousarray(np.asarray(cell, dtype=np.float64).reshape(3, 3).T)     # PBCCalculator.pyx:33-34
    return cellmat, np.ascontiguousarray(np.linalg.inv(cellmat))


def pbc_distances(cell, a, b):
    """``PBCCalculator(cell).distances(a_i, b)`` for every row of ``a``: (len(a), len(b)) float64, on the device."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("sitator_b200 needs a CUDA device; there is no CPU path")
    lib = _native.load()
    cellmat, cellinv = _cell_matrices(cell)
    da = torch.as_tensor(np.array(a, dtype=np.float64, order="C").reshape(-1, 3), device="cuda")
    db = torch.as_tensor(np.array(b, dtype=np.float64, order="C").reshape(-1, 3), device="cuda")
    out = torch.empty((da.shape[0], db.shape[0]), dtype=torch.float64, device="cuda")
    if out.numel():
        _native.check(lib.sitb_pbc_distances(torch.cuda.current_device(), cellmat.ctypes.data, cellinv.ctypes.data,
                                             C.c_void_p(da.data_ptr()), C.c_void_p(db.data_ptr()), da.shape[0], db.shape[0],
                                             C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out.cpu().numpy()


def pbc_weighted_averages(cell, points, weights):
    """Per row of ``weights`` (n_sets, n_points): ``PBCCalculator(cell).average(points[w > 0], weights=w[w > 0])``."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("sitator_b200 needs a CUDA device; there is no CPU path")
    lib = _native.load()
    cellmat, cellinv = _cell_matrices(cell)
    dp = torch.as_tensor(np.array(points, dtype=np.float64, order="C").reshape(-1, 3), device="cuda")
    dw = torch.as_tensor(np.array(weights, dtype=np.float64, order="C"), device="cuda")
    out = torch.empty((dw.shape[0], 3), dtype=torch.float64, device="cuda")
    _native.check(lib.sitb_pbc_weighted_average(torch.cuda.current_device(), cellmat.ctypes.data, cellinv.ctypes.data,
                                                C.c_void_p(dp.data_ptr()), C.c_void_p(dw.data_ptr()), dw.shape[0], dw.shape[1],
                                                C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out.cpu().numpy()


class MergeSites(abc.ABC):
    """Abstract base class for merging sites (parameters as in the reference, ``merging.py:19-41``)."""

    def __init__(self, check_types=True, maximum_merge_distance=None, set_merged_into=False,
                 weighted_spatial_average=True):
        self.check_types = check_types
        self.maximum_merge_distance = maximum_merge_distance
        self.set_merged_into = set_merged_into
        self.weighted_spatial_average = weighted_spatial_average

    def run(self, st, **kwargs):
        """Takes a ``SiteTrajectory`` and returns a new ``SiteTrajectory`` (``merging.py:44-133``)."""
        import torch
        sn = st.site_network
        if self.check_types and sn.site_types is None:
            raise ValueError("Cannot run a check_types=True MergeSites on a SiteTrajectory without type information.")
        if getattr(st, "_comm", None) is not None:
            raise NotImplementedError("site merging of a frame-sharded SiteTrajectory: gather the shards first")
        cell = np.asarray(sn.structure.cell)
        site_centers = np.asarray(sn.centers)
        site_types = sn.site_types if self.check_types else None

        clusters = self._get_sites_to_merge(st, **kwargs)
        old_n_sites, new_n_sites = sn.n_sites, len(clusters)
        logger.info("After merging %i sites there will be %i sites for %i mobile particles"
                    % (len(site_centers), new_n_sites, sn.n_mobile))
        if new_n_sites < sn.n_mobile:
            raise InsufficientSitesError(verb="Merging", n_sites=new_n_sites, n_mobile=sn.n_mobile)

        translation = np.full(old_n_sites, -1, dtype=np.int64)
        weights = np.zeros((new_n_sites, old_n_sites), dtype=np.float64)
        new_types = np.empty(new_n_sites, dtype=np.int64) if self.check_types else None
        merge_verts = sn.vertices is not None
        new_verts = [] if merge_verts else None
        # (the reference's branches: weighted_spatial_average=True averages WITHOUT weights, merging.py:101-105)
        occs = None if self.weighted_spatial_average else np.asarray(sn.occupancies, dtype=np.float64)
        for newsite, cluster in enumerate(clusters):
            mask = list(cluster)
            if np.any(translation[mask] != -1):
                raise ValueError("Site merging tried to merge site(s) into more than one new site. This shouldn't happen.")
            translation[mask] = newsite
            weights[newsite, mask] = 1.0 if occs is None else occs[mask]
            if self.check_types:
                assert np.all(site_types[mask] == site_types[mask][0])
                new_types[newsite] = site_types[mask][0]
            if merge_verts:
                new_verts.append(set.union(*[set(sn.vertices[i]) for i in mask]))

        # distance check of every group against its first member (merging.py:93-97), one device call for all groups
        if self.maximum_merge_distance is not None and new_n_sites:
            firsts = np.array([list(c)[0] for c in clusters], dtype=np.int64)
            d = pbc_distances(cell, site_centers[firsts], site_centers)             # (new, old)
            member = np.zeros((new_n_sites, old_n_sites), dtype=bool)
            for newsite, cluster in enumerate(clusters):
                member[newsite, list(cluster)[1:]] = True
            if np.any(d[member] > self.maximum_merge_distance):
                raise MergedSitesTooDistantError("Markov clustering tried to merge sites more than %.2f apart. "
                                                 "Lower your distance_threshold?" % self.maximum_merge_distance)

        # new centres: PBCCalculator.average of each group's member centres (merging.py:99-105).  The kernel centres the
        # average on the first maximum weight, as PBCCalculator.average does (:120-122); without weights the reference
        # centres on its first point, which is the first maximum of equal weights when the group is listed ascending --
        # the device rows are built in the group's own order through `order`.
        order = np.concatenate([np.asarray(list(c), dtype=np.int64) for c in clusters]) if new_n_sites else np.zeros(0, np.int64)
        w_ord = weights[:, order] if len(order) else weights
        new_centers = pbc_weighted_averages(cell, site_centers[order], w_ord) if new_n_sites else np.zeros((0, 3))

        newsn = sn.copy()
        newsn.centers = new_centers
        if self.check_types:
            newsn.site_types = new_types
        if merge_verts:
            newsn.vertices = new_verts

        # relabel the assignment stream on the device (translation[traj], unknown stays unknown; merging.py:120-121)
        lib = _native.load()
        traj = torch.as_tensor(np.ascontiguousarray(st.traj, dtype=np.int64), device="cuda").clone()
        trans = torch.as_tensor(translation, device="cuda")
        _native.check(lib.sitb_relabel_sites(torch.cuda.current_device(), C.c_void_p(traj.data_ptr()), traj.numel(), old_n_sites,
                                             C.c_void_p(trans.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        newst = SiteTrajectory(newsn, traj.cpu().numpy(), confidences=None)
        if st.real_trajectory is not None:
            newst.set_real_traj(st.real_trajectory)
        if self.set_merged_into:
            if sn.has_attribute("merged_into"):
                sn.remove_attribute("merged_into")
            sn.add_site_attribute("merged_into", translation)
        return newst

    @abc.abstractmethod
    def _get_sites_to_merge(self, st, **kwargs):
        """Groups of sites to merge: a list of lists/tuples of site numbers, no overlap; a site in no group disappears."""
