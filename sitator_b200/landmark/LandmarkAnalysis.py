"""``LandmarkAnalysis``: the drop-in, B200-native replacement of sitator's landmark analysis.

Same constructor, ``run(sn, frames) -> SiteTrajectory``, properties, attributes, constants and
exceptions as the reference's ``sitator/landmark/LandmarkAnalysis.py:29-318``.  ``run`` drives
device passes (``sitator_b200.engine``, ``csrc/``) instead of the Cython fill + NumPy clustering:

  reference step                                   here
  ------------------------------------------------ ------------------------------------------------
  wrap a copy of frames (:181-189)                 fused into K1 (frames are only read)
  site-vertex distances (:191-202)                 context creation (sitb_tables.cu)
  helpers._fill_landmark_vectors (:220)            K1, once per clustering pass, never stored dense
  cluster plugin (:234-242)                        ``landmark.cluster.<name>.do_landmark_clustering``
  site centres (:276-299)                          K6 (``sitb_site_*``)
  check_multiple_occupancy (:311)                  K5 (``sitb_check_multiple_occupancy``)

Frame-sharded runs: when ``torch.distributed`` is initialised (one process per GPU) ``frames`` is
this rank's contiguous block of the trajectory, rank order = frame order; the returned
``SiteTrajectory`` covers the same block with globally consistent sites.
"""
import importlib
import logging
import math
import time

import numpy as np

from ..SiteNetwork import SiteNetwork
from ..SiteTrajectory import SiteTrajectory
from ..errors import MultipleOccupancyError, InsufficientSitesError  # noqa: F401
from .errors import StaticLatticeError, ZeroLandmarkError

logger = logging.getLogger(__name__)

from functools import wraps


def analysis_result(func):
    @property
    @wraps(func)
    def wrapper(self, *args, **kwargs):
        if not self._has_run:
            raise ValueError("This LandmarkAnalysis hasn't been run yet.")
        return func(self, *args, **kwargs)
    return wrapper


class LandmarkAnalysis(object):
    """Site analysis of mobile atoms in a static lattice with landmark analysis.

    Parameters are those of the reference (``LandmarkAnalysis.py:95-108``); ``cutoff_center`` is
    accepted as an alias of ``cutoff_midpoint`` (the reference's docstring uses the former name).
    ``force_no_memmap`` is accepted and ignored: landmark vectors are never stored.
    """

    SITE_CENTERS_REAL_UNWEIGHTED = 'real-unweighted'
    SITE_CENTERS_REAL_WEIGHTED = 'real-weighted'
    SITE_CENTERS_REPRESENTATIVE_LANDMARK = 'representative-landmark'

    CLUSTERING_CLUSTER_SIZE = 'cluster-size'
    CLUSTERING_LABELS = 'cluster-labels'
    CLUSTERING_CONFIDENCES = 'cluster-confs'
    CLUSTERING_LANDMARK_GROUPINGS = 'cluster-landmark-groupings'
    CLUSTERING_REPRESENTATIVE_LANDMARKS = 'cluster-representative-lvecs'

    def __init__(self,
                 clustering_algorithm='dotprod',
                 clustering_params={},
                 cutoff_midpoint=1.5,
                 cutoff_steepness=30,
                 minimum_site_occupancy=0.01,
                 site_centers_method=SITE_CENTERS_REAL_WEIGHTED,
                 check_for_zero_landmarks=True,
                 static_movement_threshold=1.0,
                 dynamic_lattice_mapping=False,
                 relaxed_lattice_checks=False,
                 max_mobile_per_site=1,
                 force_no_memmap=False,
                 verbose=True,
                 cutoff_center=None,
                 device=None):
        self._cutoff_midpoint = cutoff_midpoint if cutoff_center is None else cutoff_center
        self._cutoff_steepness = cutoff_steepness
        self._minimum_site_occupancy = minimum_site_occupancy

        self._cluster_algo = clustering_algorithm
        self._clustering_params = clustering_params

        self.verbose = verbose
        self.check_for_zero_landmarks = check_for_zero_landmarks
        self.site_centers_method = site_centers_method
        self.dynamic_lattice_mapping = dynamic_lattice_mapping
        self.relaxed_lattice_checks = relaxed_lattice_checks

        self._landmark_vectors = None
        self._landmark_dimension = None

        self.static_movement_threshold = static_movement_threshold
        self.max_mobile_per_site = max_mobile_per_site

        self.force_no_memmap = force_no_memmap
        self._device = device
        self._engine = None
        self._has_run = False
        self.stats = {}

    @property
    def cutoff(self):
        return (self._cutoff_midpoint, self._cutoff_steepness)

    @analysis_result
    def landmark_vectors(self):
        """Landmark vectors from the last ``run()``: materialised on demand as a read-only
        (n_frames * n_mobile, landmark_dimension) float64 array."""
        if self._landmark_vectors is None:
            self._landmark_vectors = self._dense_landmark_vectors(self._engine)
        view = self._landmark_vectors[:]
        view.flags.writeable = False
        return view

    @staticmethod
    def _dense_landmark_vectors(eng):
        """The reference's (n_frames * n_mobile, L) float64 matrix (LandmarkAnalysis.py:209-220) in host memory, filled
        chunk by chunk from the device."""
        import torch
        out = np.empty((eng.n_frames * eng.M, eng.L), dtype=np.float64)
        step = max(1, (256 << 20) // (eng.M * eng.L * 8))
        for f0 in range(0, eng.n_frames, step):
            n = min(step, eng.n_frames - f0)
            out[f0 * eng.M:(f0 + n) * eng.M] = eng.fill_dense(f0, n, dtype=torch.float64).cpu().numpy()
        return out

    @analysis_result
    def landmark_dimension(self):
        """Number of components in a single landmark vector."""
        return self._landmark_dimension

    # ---------------------------------------------------------------------------------------------
    def _raise_first_error(self, status, comm, engine):
        first = status.first_error(self.check_for_zero_landmarks)
        key = (1 << 62) if first is None else ((first[1] << 8) | first[0])     # (frame, code) ordering
        if comm is not None:
            gkey = comm.allreduce_min_int(key)
            if gkey == (1 << 62):
                return
            if gkey != key:      # another rank holds the first error; raise the same type with its frame
                code, frame = gkey & 0xFF, gkey >> 8
                first = (code, frame, -1)
        if first is None:
            return
        code, frame, index = first
        if code == 1:
            raise StaticLatticeError(
                "No static atom position within %f A threshold of static lattice position %i"
                % (self.static_movement_threshold, index),
                lattice_atoms=[index], frame=frame, try_recentering=True)
        if code == 2:
            not_assigned = engine.unassigned_lattice_atoms(frame) if hasattr(engine, "unassigned_lattice_atoms") else None
            raise StaticLatticeError(
                "At frame %i, static positions of atoms %s not assigned to lattice positions" % (frame, not_assigned),
                lattice_atoms=not_assigned, frame=frame, try_recentering=True)
        if code == 3:
            raise ZeroLandmarkError(mobile_index=index, frame=frame)

    def run(self, sn, frames):
        """Run the landmark analysis (see :meth:`_run`); every device allocation and launch happens on ``device``."""
        import torch
        if self._device is None:
            return self._run(sn, frames)
        with torch.cuda.device(int(self._device)):      # helpers that allocate on "cuda" follow the analysis' device
            return self._run(sn, frames)

    def _run(self, sn, frames):
        """Run the landmark analysis.

        Args:
            sn (SiteNetwork): the landmark basis; each site is a landmark defined by its vertex
                static atoms (``sn.vertices``).
            frames (ndarray n_frames x n_atoms x 3, float64): a trajectory, may be unwrapped.  It is
                only read (the reference wraps a copy).  Also accepted: float32 / ``np.memmap`` arrays, and a
                :class:`sitator_b200.landmark.ChunkedFrames` for trajectories that should not be held in host
                memory as a whole.
        """
        import torch
        from ..engine import LandmarkEngine
        from .source import LandmarkVectorSource
        from .frames import ChunkedFrames
        from . import parallel

        if not (isinstance(sn, SiteNetwork) or all(hasattr(sn, a) for a in ("static_mask", "mobile_mask", "centers", "vertices"))):
            raise TypeError("sn must be a SiteNetwork")
        if self._has_run:
            raise ValueError("Cannot rerun LandmarkAnalysis!")
        if frames.shape[1:] != (sn.n_total, 3):
            raise ValueError("Wrong shape %s for frames." % (frames.shape,))
        if sn.vertices is None:
            raise ValueError("Input SiteNetwork must have vertices")
        if self.site_centers_method not in (self.SITE_CENTERS_REAL_WEIGHTED, self.SITE_CENTERS_REAL_UNWEIGHTED,
                                            self.SITE_CENTERS_REPRESENTATIVE_LANDMARK):
            raise ValueError("Invalid site centers method '%s'" % self.site_centers_method)

        from ..util.phases import PhaseTimer
        n_frames = len(frames)
        logger.info("--- Running Landmark Analysis ---")
        timer = PhaseTimer()
        t_wall0 = time.perf_counter()
        comm = parallel.default_comm()
        with timer.phase("frame offsets (collective)"):
            frame0 = 0 if comm is None else comm.exclusive_scan_int(n_frames)

        # -- Steps 0/1: context (cell, tables, site-vertex distances); frames resident on the device
        self._landmark_dimension = sn.n_sites
        t_start = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
        with timer.phase("context: tables + candidate grid"):
            engine = LandmarkEngine.from_site_network(
                sn, cutoff_midpoint=self._cutoff_midpoint, cutoff_steepness=self._cutoff_steepness,
                static_movement_threshold=self.static_movement_threshold,
                dynamic_lattice_mapping=self.dynamic_lattice_mapping,
                relaxed_lattice_checks=self.relaxed_lattice_checks, device=self._device)
        self._engine = engine
        t_start.record()
        with timer.phase("upload frames (enqueue)"):
            if isinstance(frames, ChunkedFrames):
                # a trajectory that arrives in blocks is assembled in HBM; the engine borrows the device array
                engine.set_frames(frames.to_device(engine.device), frame0=frame0)
            else:
                engine.set_frames(frames, frame0=frame0)
            engine.reset_status()

        # -- Steps 2/3: landmark vectors + clustering
        logger.info("  - computing landmark vectors / clustering -")
        source = LandmarkVectorSource(engine, comm)
        source.timer = timer
        source.defer_d2h = True           # the plugin may leave the labels' device -> host copy in flight (see below)
        try:
            clustermod = importlib.import_module("sitator_b200.landmark.cluster." + self._cluster_algo)
        except ImportError:
            # a dotted module path: a plugin written for the reference's contract (LandmarkAnalysis.py:234-242),
            # do_landmark_clustering(landmark_vectors (N, L) ndarray, clustering_params, min_samples, verbose) -> dict
            clustermod = importlib.import_module(self._cluster_algo)
        native_plugin = hasattr(clustermod, "landmark_graph") or hasattr(clustermod, "first_pass")
        cluster_input = source
        with timer.phase("pass A: fill + seen + Gram (+ H2D wait, all-reduce)"):
            if hasattr(clustermod, "landmark_graph"):
                clustermod.landmark_graph(source, self._clustering_params.get('gram_method', 'sparse'))   # pass A runs the lattice / zero-vector checks
            elif hasattr(clustermod, "first_pass"):
                clustermod.first_pass(source)
            else:
                if comm is not None:
                    raise NotImplementedError("a clustering plugin with the reference's dense-matrix contract cannot run "
                                              "frame-sharded; use 'mcl' or 'dotprod'")
                # the reference's contract: the whole (N, L) matrix, here materialised from the device (this is the
                # reference's memory footprint -- 12 KB per landmark vector at L = 1500 -- and meant for small runs)
                self._landmark_vectors = self._dense_landmark_vectors(engine)
                cluster_input = self._landmark_vectors
        with timer.phase("status + first error (collective)"):
            status = engine.status()
            self._raise_first_error(status, comm, engine)
        n_overflow = status.n_list_overflow if comm is None else comm.allreduce_sum_scalar(status.n_list_overflow)
        if n_overflow:          # (reduced over the ranks first: every rank raises, none is left waiting in a collective)
            raise RuntimeError("%d landmark vectors have more than %d candidate components after the FP32 screen; the "
                               "per-warp lists of the fill kernel would truncate them (a very wide or soft cut-off on a "
                               "dense lattice: lower cutoff_midpoint or raise cutoff_steepness)" % (n_overflow, 255))
        self.n_all_zero_lvecs = status.n_zero_rows if comm is None else comm.allreduce_sum_scalar(status.n_zero_rows)
        if status.n_duplicate_nearest:
            logger.warning("%i times a static atom was the closest to more than one static lattice position"
                           % status.n_duplicate_nearest)
        if not self.check_for_zero_landmarks and self.n_all_zero_lvecs > 0:
            logger.warning("     Had %i all-zero landmark vectors; no error because `check_for_zero_landmarks = False`."
                           % self.n_all_zero_lvecs)

        logger.info("  - clustering landmark vectors -")
        with timer.phase("clustering"):
            clustering = clustermod.do_landmark_clustering(
                cluster_input, clustering_params=self._clustering_params,
                min_samples=self._minimum_site_occupancy / float(sn.n_mobile), verbose=self.verbose)

        cluster_counts = clustering[LandmarkAnalysis.CLUSTERING_CLUSTER_SIZE]
        lmk_lbls = clustering[LandmarkAnalysis.CLUSTERING_LABELS]
        lmk_confs = clustering[LandmarkAnalysis.CLUSTERING_CONFIDENCES]
        landmark_clusters = clustering.get(LandmarkAnalysis.CLUSTERING_LANDMARK_GROUPINGS)
        if landmark_clusters is not None:
            assert len(cluster_counts) == len(landmark_clusters)
        rep_lvecs = clustering.get(LandmarkAnalysis.CLUSTERING_REPRESENTATIVE_LANDMARKS)
        if rep_lvecs is not None:
            rep_lvecs = np.asarray(rep_lvecs)
            assert rep_lvecs.shape == (len(cluster_counts), engine.L)

        with timer.phase("unassigned count (collective)"):
            if clustering.get('_n_unassigned') is not None:
                n_unassigned = int(clustering['_n_unassigned'])                   # already reduced over the ranks
                n_rows = source.n_total
            else:
                if clustering.get('_dev_labels') is not None:
                    n_unassigned = int((clustering['_dev_labels'] < 0).sum().item())      # counted where the labels already are
                else:
                    n_unassigned = int(np.sum(lmk_lbls < 0))
                n_rows = len(lmk_lbls)
                if comm is not None:
                    n_unassigned, n_rows = comm.allreduce_sum_scalars([n_unassigned, n_rows])
        logger.info("    Failed to assign %i%% of mobile particle positions to sites." % (100.0 * n_unassigned / float(n_rows)))

        lmk_lbls = lmk_lbls.reshape(n_frames, sn.n_mobile)
        lmk_confs = lmk_confs.reshape(n_frames, sn.n_mobile)
        n_sites = len(cluster_counts)
        if n_sites < (sn.n_mobile / self.max_mobile_per_site):
            raise InsufficientSitesError(verb="Landmark analysis", n_sites=n_sites, n_mobile=sn.n_mobile)
        logger.info("    Identified %i sites with assignment counts %s" % (n_sites, cluster_counts))

        # -- output network: site centres (LandmarkAnalysis.py:276-299)
        out_sn = sn.copy()
        dev_labels = clustering.get('_dev_labels')
        dev_confs = clustering.get('_dev_confs')
        if dev_labels is None:
            dev_labels = torch.as_tensor(np.ascontiguousarray(lmk_lbls.reshape(-1)), device=engine.device)
            dev_confs = torch.as_tensor(np.ascontiguousarray(lmk_confs.reshape(-1)), device=engine.device)
        with timer.phase("site centres (collective)"):
            if self.site_centers_method in (self.SITE_CENTERS_REAL_WEIGHTED, self.SITE_CENTERS_REAL_UNWEIGHTED):
                weighted = self.site_centers_method == self.SITE_CENTERS_REAL_WEIGHTED
                site_best = clustering.get('_dev_site_best') if weighted else None
                if weighted and site_best is None:      # a plugin with the reference's contract returns labels and confidences only
                    from .cluster.dotprod import _site_best_table
                    site_best = _site_best_table(dev_labels, dev_confs, n_sites, frame0 * sn.n_mobile)
                site_centers = engine.site_centers(dev_labels, dev_confs, n_sites, weighted, site_best, comm)
            else:
                if rep_lvecs is None:
                    raise ValueError("Chosen clustering method (with current parameters) didn't return representative "
                                     "landmark vectors; can't use SITE_CENTERS_REPRESENTATIVE_LANDMARK.")
                site_centers = engine.weighted_point_averages(np.asarray(sn.centers), rep_lvecs)
        out_sn.centers = site_centers
        if landmark_clusters is not None:
            out_sn.vertices = [set.union(*[set(sn.vertices[l]) for l in lclust]) for lclust in landmark_clusters]

        out_st = SiteTrajectory(out_sn, lmk_lbls, lmk_confs, _copy=False)     # the arrays are this run's own
        out_st.frame0 = frame0
        out_st._comm = comm

        if clustering.get('_d2h_done') is not None:
            with timer.phase("labels + confidences D2H (wait)"):
                clustering['_d2h_done'].synchronize()       # st.traj / st.confidences are complete from here on

        # Check that multiple particles are never assigned to one site at the same time
        with timer.phase("occupancy check (collective)"):
            self.n_multiple_assignments, self.avg_mobile_per_site = out_st.check_multiple_occupancy(
                max_mobile_per_site=self.max_mobile_per_site, _dev_traj=dev_labels.view(n_frames, sn.n_mobile))

        out_st.set_real_traj(frames)
        t_end.record()
        torch.cuda.synchronize(engine.device)
        self.stats = {"run_ms": t_start.elapsed_time(t_end), "wall_ms": (time.perf_counter() - t_wall0) * 1e3,
                      "n_screen_rejects": status.n_screen_rejects,
                      "nnz": status.nnz, "mcl_iterations": clustering.get('_mcl_iterations'),
                      "gram_method": getattr(source, "gram_method", None),
                      "plugin_contract": "source" if native_plugin else "ndarray",
                      "phases_ms": timer.as_dict(), "phases_synchronised": timer.sync}
        # (landmark -> cluster map, landmark weight) of the final site centres, in the caller's landmark numbering: the
        # rows of the reference's `centers` matrix (cluster/mcl.py:70-96) in sparse form; None for plugins that do not say
        self.cluster_centers_ = clustering.get('_centers')
        self._has_run = True
        return out_st
