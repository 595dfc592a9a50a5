"""``JumpAnalysis``: jump statistics of a ``SiteTrajectory`` (mirrors reference
``sitator/dynamics/JumpAnalysis.py:11-135``), computed by integer scan kernels on the GPU.

Adds these edge attributes to the SiteTrajectory's SiteNetwork:
 - ``n_ij``: total number of jumps from i to j (diagonal: frames spent without jumping).
 - ``p_ij``: being at i, the probability of jumping to j.
 - ``jump_lag``: average number of frames a particle spends at i before jumping to j (+inf if never).
And these site attributes:
 - ``residence_times``, ``occupancy_freqs``, ``total_corrected_residences``.

The reference accumulates with NumPy fancy-index ``+=`` (``JumpAnalysis.py:75,79,87-88``), which counts
a duplicated (i, j) pair inside one frame once and lets the last duplicate win for the lag sum; the
kernel reproduces exactly that (``csrc/sitb_traj.cu: k_ja_accumulate``).
``jump_lag_by_type`` / ``plot_jump_lag`` are out of scope (site types / plotting).
"""
import ctypes as C
import logging

import numpy as np

from .. import _native
from ..SiteTrajectory import SiteTrajectory

logger = logging.getLogger(__name__)


class JumpAnalysis(object):
    def __init__(self):
        pass

    def run(self, st):
        """Adds edge/site attributes to ``st``'s ``SiteNetwork``; returns ``st``."""
        import torch
        assert isinstance(st, SiteTrajectory)
        comm = getattr(st, "_comm", None)
        logger.info("Running JumpAnalysis...")
        lib = _native.load()
        n_mobile = st.site_network.n_mobile
        n_frames = st.n_frames
        n_sites = st.site_network.n_sites
        dev = torch.cuda.current_device()
        traj = torch.as_tensor(np.ascontiguousarray(st.traj, dtype=np.int64), device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        first_is_start, carry_label, carry_jump = 1, None, None
        if comm is not None:
            # frame-sharded: the scan's sequential state (last known site, frame of the last jump; JumpAnalysis.py:46-49,
            # :91-96) enters each shard as a carry chained over the lower ranks' summaries
            summ = torch.empty((n_mobile, 4), dtype=torch.int64, device="cuda")
            _native.check(lib.sitb_jump_analysis_summary(dev, C.c_void_p(traj.data_ptr()), n_frames, n_mobile,
                                                         C.c_void_p(summ.data_ptr()), C.c_void_p(stream)))
            frame0 = int(st.frame0)
            mine = summ.cpu().numpy()
            for col in (0, 3):                                   # local -> global frame indices
                mine[:, col] = np.where(mine[:, col] >= 0, mine[:, col] + frame0, -1)
            alls = comm.allgather_numpy(mine)                    # (world, M, 4), rank order = frame order
            lab = np.full(n_mobile, -1, dtype=np.int64)
            jmp = np.full(n_mobile, -1, dtype=np.int64)          # global frame of the last jump; -1 = never
            for r in range(comm.rank):
                s = alls[r]
                has = s[:, 0] >= 0
                jmp = np.where(has & (lab >= 0) & (s[:, 1] != lab), s[:, 0], jmp)
                jmp = np.where(has & (s[:, 3] >= 0), s[:, 3], jmp)
                lab = np.where(has, s[:, 2], lab)
            first_is_start = int(comm.rank == 0)
            carry_label = torch.as_tensor(lab, device="cuda")
            carry_jump = torch.as_tensor(jmp - frame0, device="cuda")      # local index (negative: an earlier shard)
            n_frames_total = comm.allreduce_sum_scalar(n_frames)
        else:
            n_frames_total = n_frames
        n_ij = torch.zeros((n_sites, n_sites), dtype=torch.float64, device="cuda")
        lag_sum = torch.zeros((n_sites, n_sites), dtype=torch.float64, device="cuda")
        lag_n = torch.zeros((n_sites, n_sites), dtype=torch.int64, device="cuda")
        total_time = torch.zeros((n_sites,), dtype=torch.int64, device="cuda")
        n_problems = torch.zeros((1,), dtype=torch.int64, device="cuda")
        P = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
        _native.check(lib.sitb_jump_analysis(dev, P(traj), n_frames, n_mobile, n_sites, first_is_start, P(carry_label),
                                             P(carry_jump), P(n_ij), P(total_time), P(lag_sum), P(lag_n), P(n_problems),
                                             C.c_void_p(stream)))
        if comm is not None:
            for t_ in (n_ij, lag_sum, lag_n, total_time, n_problems):
                comm.allreduce_sum_(t_)
        n_ij = n_ij.cpu().numpy()
        avg_time_before_jump = lag_sum.cpu().numpy()
        avg_time_before_jump_n = lag_n.cpu().numpy()
        total_time_spent_at_site = total_time.cpu().numpy()
        n_problems = int(n_problems.item())

        assert not np.any(np.nonzero(avg_time_before_jump.diagonal()))
        if n_problems != 0:
            logger.warning("Came across %i times where assignment and last known assignment were unassigned." % n_problems)

        msk = avg_time_before_jump_n > 0
        avg_time_before_jump[~msk] = np.inf                       # JumpAnalysis.py:104-108
        avg_time_before_jump[msk] /= avg_time_before_jump_n[msk]

        sn = st.site_network
        if sn.has_attribute('n_ij'):
            for name in ('n_ij', 'p_ij', 'jump_lag', 'residence_times', 'occupancy_freqs', 'total_corrected_residences'):
                sn.remove_attribute(name)
        sn.add_edge_attribute('jump_lag', avg_time_before_jump)
        sn.add_edge_attribute('n_ij', n_ij)
        with np.errstate(divide='ignore', invalid='ignore'):
            sn.add_edge_attribute('p_ij', n_ij / total_time_spent_at_site)

        res_times = np.empty(shape=n_sites, dtype=np.float64)
        for site in range(n_sites):                                # JumpAnalysis.py:122-129
            times = avg_time_before_jump[site]
            noninf = times < np.inf
            res_times[site] = np.mean(times[noninf]) if np.any(noninf) else n_frames_total
        sn.add_site_attribute('residence_times', res_times)
        sn.add_site_attribute('occupancy_freqs', np.sum(n_ij, axis=0) / n_frames_total)
        sn.add_site_attribute('total_corrected_residences', total_time_spent_at_site)
        return st
