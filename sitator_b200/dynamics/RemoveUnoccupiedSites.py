"""Remove sites no mobile atom ever occupies (reference ``sitator/dynamics/RemoveUnoccupiedSites.py:10-71``).

The seen mask and the relabelling of the (n_frames, n_mobile) assignment stream run on the GPU
(``csrc/sitb_post.cu``: ``sitb_seen_sites``, ``sitb_relabel_sites``)."""
import ctypes as C
import logging

import numpy as np

from .. import _native
from ..SiteTrajectory import SiteTrajectory, _device_index
from ..errors import InsufficientSitesError

logger = logging.getLogger(__name__)


class RemoveUnoccupiedSites(object):
    """Remove unoccupied sites."""

    def run(self, st, return_kept_sites=False):
        """
        Args:
            return_kept_sites (bool): if True, the sites of ``st`` that were kept are returned as well.
        Returns:
            A ``SiteTrajectory``, or ``st`` itself if it has no unoccupied sites.
        """
        import torch
        assert isinstance(st, SiteTrajectory)
        lib = _native.load()
        dev = _device_index()
        old_sn = st.site_network
        n_sites = old_sn.n_sites
        traj = st._device_traj()
        stream = torch.cuda.current_stream().cuda_stream
        seen = torch.zeros(n_sites, dtype=torch.int32, device="cuda")
        _native.check(lib.sitb_seen_sites(dev, C.c_void_p(traj.data_ptr()), traj.numel(), n_sites,
                                          C.c_void_p(seen.data_ptr()), C.c_void_p(stream)))
        if st._comm is not None:
            st._comm.allreduce_sum_(seen)
        seen_mask = seen.cpu().numpy() > 0
        if np.all(seen_mask):                                              # ref :33-35
            return st
        logger.info("Removing unoccupied sites %s" % np.where(~seen_mask)[0])
        n_new_sites = int(np.sum(seen_mask))
        if n_new_sites < old_sn.n_mobile:                                   # ref :43-48
            raise InsufficientSitesError(verb="Removing unoccupied sites", n_sites=n_new_sites, n_mobile=old_sn.n_mobile)
        translation = np.full(n_sites, -4321, dtype=np.int64)               # ref :50-53
        translation[seen_mask] = np.arange(n_new_sites)
        d_tr = torch.as_tensor(translation, device="cuda")
        _native.check(lib.sitb_relabel_sites(dev, C.c_void_p(traj.data_ptr()), traj.numel(), n_sites,
                                             C.c_void_p(d_tr.data_ptr()), C.c_void_p(stream)))
        newtraj = traj.cpu().numpy().reshape(st.traj.shape)
        assert -4321 not in newtraj
        newsn = old_sn[seen_mask]                                           # computed attributes are kept (ref :60-61)
        new_st = SiteTrajectory(site_network=newsn, particle_assignments=newtraj, _copy=False)
        new_st.frame0, new_st._comm = st.frame0, st._comm
        if st.real_trajectory is not None:
            new_st.set_real_traj(st.real_trajectory)
        if return_kept_sites:
            return new_st, np.where(seen_mask)
        return new_st
