"""``SiteTrajectory``: which site every mobile atom occupies in every frame.

Data contract of the landmark-analysis path (mirrors reference ``sitator/SiteTrajectory.py:10-389``:
same constructor, properties, method names, argument meaning and errors).  The per-frame scans --
the multiple-occupancy check and the jump list -- run as integer kernels on the GPU
(``csrc/sitb_traj.cu``) instead of Python loops over frames.  Plotting is out of scope.
"""
import ctypes as C
import logging

import numpy as np

from . import _native
from .errors import MultipleOccupancyError

logger = logging.getLogger(__name__)


def _device_index():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("sitator_b200 needs a CUDA device; there is no CPU path")
    return torch.cuda.current_device()


class SiteTrajectory(object):
    """A trajectory capturing the dynamics of particles through a SiteNetwork."""

    SITE_UNKNOWN = -1

    def __init__(self, site_network, particle_assignments, confidences=None, _copy=True):
        """
        :param SiteNetwork site_network:
        :param ndarray (n_frames, n_mobile) particle_assignments:
        :param ndarray (n_frames, n_mobile) confidences (optional)
        """
        if particle_assignments.ndim != 2:
            raise ValueError("particle_assignments must be 2D")
        if particle_assignments.shape[1] != site_network.n_mobile:
            raise ValueError("particle_assignments has wrong shape %s" % (particle_assignments.shape,))
        self._sn = site_network
        self._traj = particle_assignments.copy() if _copy else particle_assignments
        if confidences is not None:
            if confidences.shape != particle_assignments.shape:
                raise ValueError("confidences has wrong shape %s; should be %s" %
                                 (confidences.shape, particle_assignments.shape))
            self._confs = confidences
        else:
            self._confs = None
        self._real_traj = None
        self._default_plotter = None
        # frame-sharded runs: global index of this shard's first frame and the carry from earlier shards
        self.frame0 = 0
        self._comm = None

    def __len__(self):
        return self.n_frames

    def __getitem__(self, key):
        st = type(self)(self._sn, self._traj[key], confidences=None if self._confs is None else self._confs[key])
        if self._real_traj is not None:
            st.set_real_traj(self._real_traj[key])
        return st

    def __getstate__(self):
        state = self.__dict__.copy()
        state['_real_traj'] = None            # do not pickle giant trajectories (ref :55-63)
        state['_default_plotter'] = None
        state['_comm'] = None
        return state

    # ---- plain properties (ref :65-108) ------------------------------------------------------
    @property
    def traj(self):
        """The site assignments over time."""
        return self._traj

    @property
    def confidences(self):
        return self._confs

    @property
    def n_frames(self):
        return len(self._traj)

    @property
    def n_unassigned(self):
        return np.sum(self._traj < 0)

    @property
    def n_assigned(self):
        return self._sn.n_mobile * self.n_frames - self.n_unassigned

    @property
    def percent_unassigned(self):
        return float(self.n_unassigned) / (self._sn.n_mobile * self.n_frames)

    @property
    def site_network(self):
        return self._sn

    @site_network.setter
    def site_network(self, value):
        assert np.all(value.mobile_mask == self._sn.mobile_mask)
        assert np.all(value.static_mask == self._sn.static_mask)
        self._sn = value

    @property
    def real_trajectory(self):
        return self._real_traj

    def copy(self, with_computed=True):
        st = self[:]
        st.site_network = st.site_network.copy(with_computed=with_computed)
        st.frame0, st._comm = self.frame0, self._comm
        return st

    def set_real_traj(self, real_traj):
        expected_shape = (self.n_frames, self._sn.n_total, 3)
        if not real_traj.shape == expected_shape:
            raise ValueError("real_traj of shape %s does not have expected shape %s" % (real_traj.shape, expected_shape))
        self._real_traj = real_traj

    def remove_real_traj(self):
        self._real_traj = None

    def trajectory_for_particle(self, i, return_confidences=False):
        if return_confidences and self._confs is None:
            raise ValueError("This SiteTrajectory has no confidences")
        if return_confidences:
            return self._traj[:, i], self._confs[:, i]
        return self._traj[:, i]

    def real_positions_for_site(self, site, return_confidences=False):
        if self._real_traj is None:
            raise ValueError("This SiteTrajectory has no real trajectory")
        if return_confidences and self._confs is None:
            raise ValueError("This SiteTrajectory has no confidences")
        assert site < self._sn.n_sites
        msk = self._traj == site
        pts = self._real_traj[:, self._sn.mobile_mask][msk]
        if return_confidences:
            return pts, self._confs[msk].flatten()
        return pts

    def compute_site_occupancies(self):
        """Adds site attribute ``occupancies`` (ref :187-202)."""
        counts = np.bincount(self._traj[self._traj >= 0], minlength=self._sn.n_sites)
        n_frames = self.n_frames
        if self._comm is not None:                # frame-sharded: occupancies over the whole trajectory
            counts = self._comm.allreduce_sum_numpy(counts.astype(np.int64))
            n_frames = self._comm.allreduce_sum_scalar(n_frames)
        occ = np.true_divide(counts, n_frames)
        if self.site_network.has_attribute('occupancies'):
            self.site_network.remove_attribute('occupancies')
        self.site_network.add_site_attribute('occupancies', occ)
        return occ

    # ---- device helpers ------------------------------------------------------------------------
    def _device_traj(self):
        import torch
        t = torch.as_tensor(np.ascontiguousarray(self._traj, dtype=np.int64), device="cuda")
        return t

    # ---- SiteTrajectory.check_multiple_occupancy (ref :205-232) -----------------------------------
    def check_multiple_occupancy(self, max_mobile_per_site=1, _dev_traj=None):
        """Count cases where more than one mobile atom shares a site in a frame.

        Returns:
            int: total number of multiple-assignment incidents; float: average number of mobile atoms
            at any occupied site at any one time.
        Raises ``MultipleOccupancyError`` for the first frame (and lowest site) where a site holds
        more than ``max_mobile_per_site`` atoms.
        """
        import torch
        lib = _native.load()
        dev = _device_index()
        traj = self._device_traj() if _dev_traj is None else _dev_traj
        out = torch.zeros(3, dtype=torch.int64, device="cuda")
        bad = torch.full((1,), -1, dtype=torch.int64, device="cuda")       # all ones
        stream = torch.cuda.current_stream().cuda_stream
        _native.check(lib.sitb_check_multiple_occupancy(
            dev, C.c_void_p(traj.data_ptr()), self.n_frames, self._sn.n_mobile, int(self.frame0),
            int(max_mobile_per_site), C.c_void_p(out.data_ptr()), C.c_void_p(bad.data_ptr()), C.c_void_p(stream)))
        if self._comm is not None:
            self._comm.allreduce_sum_(out)
            self._comm.allreduce_min_u64_(bad)
        key = int(bad.cpu().numpy().view(np.uint64)[0])
        if key != 0xFFFFFFFFFFFFFFFF:
            frame, site = key >> 32, key & 0xFFFFFFFF
            local = frame - self.frame0
            mobile = np.where(self._traj[local] == site)[0] if 0 <= local < self.n_frames else np.array([], dtype=int)
            if self._comm is not None:      # only the shard that owns the frame knows the atoms: one-hot sum over ranks
                mask = np.zeros(self._sn.n_mobile, dtype=np.int64)
                mask[mobile] = 1
                mobile = np.where(self._comm.allreduce_sum_numpy(mask) > 0)[0]
            raise MultipleOccupancyError(mobile=mobile, site=site, frame=frame)
        n_more, n_assigned, n_distinct = (int(x) for x in out.cpu().numpy())
        return n_more, n_assigned / n_distinct

    # ---- SiteTrajectory.assign_to_last_known_site (ref :235-304) ---------------------------------
    def assign_to_last_known_site(self, frame_threshold=1):
        """Assign unassigned mobile particles to their last known site (in place).

        Args:
            frame_threshold (int): the maximum number of frames between the last known site and the present
                frame up to which the last known site can be used.
        Returns:
            dict with ``max_time_unknown``, ``avg_time_unknown``, ``total_reassigned`` (the reference's diagnostics).
        """
        import torch
        lib = _native.load()
        dev = _device_index()
        total_unknown = self.n_unassigned
        logger.info("%i unassigned positions (%i%%); assigning unassigned mobile particles to last known positions "
                    "within %s frames..." % (total_unknown, 100.0 * self.percent_unassigned, frame_threshold))
        traj = self._device_traj()
        M = self._sn.n_mobile
        stream = torch.cuda.current_stream().cuda_stream
        stats = torch.zeros(4, dtype=torch.int64, device="cuda")
        carry_l = carry_t = None
        if self._comm is not None:
            # state at the end of every shard taken alone, chained over the earlier shards on the host
            end_l = torch.empty(M, dtype=torch.int64, device="cuda")
            end_t = torch.empty(M, dtype=torch.int64, device="cuda")
            _native.check(lib.sitb_assign_last_known(dev, C.c_void_p(traj.data_ptr()), self.n_frames, M, int(self.frame0),
                                                     int(frame_threshold), None, None, C.c_void_p(end_l.data_ptr()),
                                                     C.c_void_p(end_t.data_ptr()), None, 0, C.c_void_p(stream)))
            all_l = self._comm.allgather_numpy(end_l.cpu().numpy())
            all_t = self._comm.allgather_numpy(end_t.cpu().numpy())
            lk = np.full(M, -1, dtype=np.int64)
            tu = np.zeros(M, dtype=np.int64)
            for r in range(self._comm.rank):
                has = all_l[r] != -1
                lk = np.where(has, all_l[r], lk)
                tu = np.where(has, all_t[r], tu + all_t[r])
            carry_l = torch.as_tensor(lk, device="cuda")
            carry_t = torch.as_tensor(tu, device="cuda")
        _native.check(lib.sitb_assign_last_known(
            dev, C.c_void_p(traj.data_ptr()), self.n_frames, M, int(self.frame0), int(frame_threshold),
            None if carry_l is None else C.c_void_p(carry_l.data_ptr()),
            None if carry_t is None else C.c_void_p(carry_t.data_ptr()), None, None,
            C.c_void_p(stats.data_ptr()), 1, C.c_void_p(stream)))
        if self._comm is not None:
            self._comm.allreduce_sum_(stats[:3])
            self._comm.allreduce_max_u64_(stats[3:])
        self._traj[...] = traj.cpu().numpy()
        reassigned, sum_times, n_times, key = (int(x) for x in stats.cpu().numpy().view(np.uint64))
        if n_times > 0:
            res = {'max_time_unknown': key & 0xFFFFFF, 'avg_time_unknown': float(sum_times) / n_times,
                   'total_reassigned': reassigned}
            logger.info("  Maximum # of frames any mobile particle spent unassigned: %i" % res['max_time_unknown'])
            logger.info("  Avg. # of frames spent unassigned: %f" % res['avg_time_unknown'])
            logger.info("  Assigned %i/%i unassigned positions, leaving %i (%i%%) unknown"
                        % (reassigned, total_unknown, self.n_unassigned, self.percent_unassigned))
        else:
            logger.info("  None to correct.")
            res = {'max_time_unknown': 0, 'avg_time_unknown': 0, 'total_reassigned': 0}
        return res

    # ---- jumps (ref :307-373) ------------------------------------------------------------------
    def jump_array(self, unknown_as_jump=False):
        """All jumps as an (n_jumps, 4) int64 array of (frame, mobile atom, from site, to site), in the
        order ``jumps()`` yields them (frame-major, atom-minor)."""
        import torch
        lib = _native.load()
        dev = _device_index()
        traj = self._device_traj()
        F, M = self.n_frames, self._sn.n_mobile
        frm = torch.empty((F, M), dtype=torch.int32, device="cuda")
        total = torch.zeros(1, dtype=torch.int64, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        carry = None
        first_is_start = 1
        if self._comm is not None:
            carry, first_is_start = self._comm.jump_carry(traj, unknown_as_jump)
        _native.check(lib.sitb_jump_scan(dev, C.c_void_p(traj.data_ptr()), F, M, int(bool(unknown_as_jump)),
                                         int(first_is_start), C.c_void_p(0 if carry is None else carry.data_ptr()),
                                         C.c_void_p(frm.data_ptr()), C.c_void_p(total.data_ptr()), C.c_void_p(stream)))
        n = int(total.item())
        out = torch.empty((max(n, 1), 4), dtype=torch.int64, device="cuda")
        _native.check(lib.sitb_jump_compact(dev, C.c_void_p(traj.data_ptr()), C.c_void_p(frm.data_ptr()), F, M,
                                            int(self.frame0), C.c_void_p(out.data_ptr()), n, C.c_void_p(stream)))
        return out[:n].cpu().numpy()

    def jumps(self, **kwargs):
        """Iterate over all jumps, jump by jump: (frame_number, mobile_atom_number, from_site, to_site)."""
        for f, a, s0, s1 in self.jump_array(**kwargs):
            yield int(f), int(a), int(s0), int(s1)

    def jumps_by_frame(self, **kwargs):
        """Iterate frame by frame (every frame after the first, as the reference does, ref :331-351):
        (frame_number, mob_that_jumped, from_sites, to_sites)."""
        ja = self.jump_array(**kwargs)
        bounds = np.searchsorted(ja[:, 0], np.arange(self.frame0, self.frame0 + self.n_frames + 1))
        for i in range(1 if self.frame0 == 0 else 0, self.n_frames):
            sl = slice(bounds[i], bounds[i + 1])
            yield self.frame0 + i, ja[sl, 1], ja[sl, 2], ja[sl, 3]

    def plot_frame(self, *args, **kwargs):
        raise NotImplementedError("plotting is outside the scope of sitator_b200 (SURVEY.md section 2, #21)")

    plot_site = plot_frame
    plot_particle_trajectory = plot_frame
