""""Smooth" a SiteTrajectory with a rolling mode (reference ``sitator/dynamics/SmoothSiteTrajectory.pyx:12-111``).

For each mobile particle the assignment at each frame is replaced by the mode of its site assignments over a
window of frames around it -- a discrete low-pass filter.  The window pass is ``sitb_windowed_mode``
(``csrc/sitb_post.cu``)."""
import ctypes as C
import logging

import numpy as np

from .. import _native
from ..SiteTrajectory import _device_index
from .RemoveUnoccupiedSites import RemoveUnoccupiedSites

logger = logging.getLogger(__name__)


class SmoothSiteTrajectory(object):
    """
    Args:
        window_threshold_factor (float): the total width of the rolling window, in terms of the threshold.
        remove_unoccupied_sites (bool): if True, sites that are unoccupied after the smoothing are removed.
        set_unassigned_under_threshold (bool): if True, a particle whose mode occurs fewer than ``threshold`` times
            in the window is marked unassigned at that frame; if False its assignment is not modified.
    """

    def __init__(self, window_threshold_factor=2.1, remove_unoccupied_sites=True, set_unassigned_under_threshold=True):
        self.window_threshold_factor = window_threshold_factor
        self.remove_unoccupied_sites = remove_unoccupied_sites
        self.set_unassigned_under_threshold = set_unassigned_under_threshold

    def run(self, st, threshold):
        import torch
        lib = _native.load()
        dev = _device_index()
        n_mobile = st.site_network.n_mobile
        window = self.window_threshold_factor * threshold                    # ref :55-56
        wleft, wright = int(np.floor(window / 2)), int(np.ceil(window / 2))
        traj = st._device_traj()
        out = torch.empty_like(traj)
        before = after = None
        if st._comm is not None and st._comm.world > 1:
            # the window reaches into the neighbouring shards: their last wleft / first wright frames
            if st.n_frames < max(wleft, wright):
                raise ValueError("every shard must hold at least one window (%d frames)" % max(wleft, wright))
            tails = st._comm.allgather_numpy(st.traj[st.n_frames - wleft:] if wleft else np.zeros((0, n_mobile), np.int64))
            heads = st._comm.allgather_numpy(st.traj[:wright] if wright else np.zeros((0, n_mobile), np.int64))
            r = st._comm.rank
            if r > 0 and wleft:
                before = torch.as_tensor(np.ascontiguousarray(tails[r - 1], dtype=np.int64), device="cuda")
            if r + 1 < st._comm.world and wright:
                after = torch.as_tensor(np.ascontiguousarray(heads[r + 1], dtype=np.int64), device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        _native.check(lib.sitb_windowed_mode(
            dev, C.c_void_p(traj.data_ptr()), C.c_void_p(out.data_ptr()), st.n_frames, n_mobile, wleft, wright,
            int(threshold), int(bool(self.set_unassigned_under_threshold)),
            0 if before is None else wleft, None if before is None else C.c_void_p(before.data_ptr()),
            0 if after is None else wright, None if after is None else C.c_void_p(after.data_ptr()), C.c_void_p(stream)))
        st = st.copy(with_computed=False)                                     # ref :70-71
        st._traj = out.cpu().numpy()
        if self.remove_unoccupied_sites:
            # removing short jumps could have made some sites completely unoccupied
            st = RemoveUnoccupiedSites().run(st)
        st.site_network.clear_attributes()
        return st
