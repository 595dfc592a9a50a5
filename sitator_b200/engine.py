"""Host-side driver of the device passes (thin: argument marshalling only).

``LandmarkEngine`` owns one native context (``include/sitator_b200.h``) on one GPU: the landmark
basis tables and a resident shard of frames.  Device buffers that cross the C ABI are torch CUDA
tensors (torch is plumbing here: allocation, streams, and -- in ``landmark/`` -- the NCCL
collectives); all arithmetic happens in ``csrc/``.
"""
import ctypes as C
import math

import numpy as np

from . import _native


def _torch():
    import torch
    return torch


class EngineStatus(object):
    """Python view of ``sitb_status``."""

    def __init__(self, s):
        self.error_code = int(s.error_code)
        self.index = int(s.index)
        self.frame = int(s.frame)
        self.zero_error = bool(s.zero_error)
        self.zero_index = int(s.zero_index)
        self.zero_frame = int(s.zero_frame)
        self.n_zero_rows = int(s.n_zero_rows)
        self.n_duplicate_nearest = int(s.n_duplicate_nearest)
        self.n_list_overflow = int(s.n_list_overflow)
        self.nnz = int(s.nnz)
        self.n_screen_rejects = int(s.n_screen_rejects)
        self.n_full_walk_frames = int(s.n_full_walk_frames)
        self.n_loose_grid_frames = int(s.n_loose_grid_frames)

    def first_error(self, check_for_zeros):
        """(code, frame, index) of the first error in the reference's iteration order, or None."""
        cands = []
        if self.error_code:
            cands.append((self.frame, self.error_code, self.index))
        if check_for_zeros and self.zero_error:
            cands.append((self.zero_frame, 3, self.zero_index))
        if not cands:
            return None
        frame, code, index = min(cands)
        return code, frame, index


def vertex_table(vertices):
    """-1 padded (L, Vmax) int32 table of the landmark vertex lists (LandmarkAnalysis.py:194-195)."""
    import itertools
    n = len(vertices)
    lens = np.fromiter((len(v) for v in vertices), dtype=np.int64, count=n)
    vmax = int(lens.max())
    flat = np.fromiter(itertools.chain.from_iterable(vertices), dtype=np.int32, count=int(lens.sum()))
    out = np.full((n, vmax), -1, dtype=np.int32)
    out[np.arange(vmax)[None, :] < lens[:, None]] = flat       # row-major fill: each list in its own order
    return out


class LandmarkEngine(object):
    def __init__(self, cell, static_idx, mobile_idx, n_atoms, ideal_static, centers, vertices,
                 cutoff_midpoint=1.5, cutoff_steepness=30.0, static_movement_threshold=1.0,
                 dynamic_lattice_mapping=False, relaxed_lattice_checks=False, device=None, candidate_grid_margin=None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("sitator_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        self._lib = _native.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        cell = np.ascontiguousarray(cell, dtype=np.float64).reshape(3, 3)
        # PBCCalculator.__init__ (PBCCalculator.pyx:27-35): cellmat = cell^T, LAPACK inverse
        self.cellmat = np.ascontiguousarray(cell.T)
        self.cellmat_inv = np.ascontiguousarray(np.linalg.inv(self.cellmat))
        self.cell_centroid = np.sum(0.5 * cell, axis=0)
        self.static_idx = np.ascontiguousarray(static_idx, dtype=np.int32)
        self.mobile_idx = np.ascontiguousarray(mobile_idx, dtype=np.int32)
        self.ideal_static = np.ascontiguousarray(ideal_static, dtype=np.float64).reshape(-1, 3)
        self.centers = np.ascontiguousarray(centers, dtype=np.float64).reshape(-1, 3)
        self.verts = np.ascontiguousarray(vertex_table(vertices))
        self.n_atoms = int(n_atoms)
        self.S, self.M, self.L, self.V = len(self.static_idx), len(self.mobile_idx), len(self.centers), self.verts.shape[1]
        if len(self.ideal_static) != self.S:
            raise ValueError("ideal_static has %d rows for %d static atoms" % (len(self.ideal_static), self.S))
        if len(vertices) != self.L:
            raise ValueError("vertices/centers length mismatch")
        # helpers.pyx:41-43,127-131 with libm log, exactly as the reference evaluates it
        self.cutoff_round_to_zero = cutoff_midpoint + math.log((1 / 0.0001) - 1.) / cutoff_steepness
        d = _native.NetworkDesc()
        d.n_atoms, d.n_static, d.n_mobile, d.n_landmarks, d.max_verts = self.n_atoms, self.S, self.M, self.L, self.V
        d.host_cellmat = self.cellmat.ctypes.data
        d.host_cellmat_inv = self.cellmat_inv.ctypes.data
        d.host_static_idx = self.static_idx.ctypes.data
        d.host_mobile_idx = self.mobile_idx.ctypes.data
        d.host_ideal_static = self.ideal_static.ctypes.data
        d.host_centers = self.centers.ctypes.data
        d.host_verts = self.verts.ctypes.data
        d.cutoff_midpoint = float(cutoff_midpoint)
        d.cutoff_steepness = float(cutoff_steepness)
        d.cutoff_round_to_zero = float(self.cutoff_round_to_zero)
        d.static_movement_threshold = float(static_movement_threshold)
        d.dynamic_lattice_mapping = int(bool(dynamic_lattice_mapping))
        d.relaxed_lattice_checks = int(bool(relaxed_lattice_checks))
        self._ctx = C.c_void_p()
        _native.check(self._lib.sitb_create(C.byref(d), self.device.index, C.byref(self._ctx)))
        self._frames_keepalive = None
        self.frames_bytes = 0          # device bytes the native context holds for uploaded frames
        self.n_frames = 0
        self.frame0 = 0
        self.n_clusters = 0
        n_sms, maj, mnr = C.c_int32(), C.c_int32(), C.c_int32()
        _native.check(self._lib.sitb_device_info(self._ctx, C.byref(n_sms), C.byref(maj), C.byref(mnr)))
        self.n_sms, self.compute_capability = n_sms.value, (maj.value, mnr.value)
        self.use_current_stream()
        if candidate_grid_margin is not None:
            self.set_candidate_grid(candidate_grid_margin)

    def set_candidate_grid(self, static_margin):
        """Rebuild the fill kernel's candidate grid for static atoms within ``static_margin`` Angstrom of their
        ideal positions (<= 0: none, every landmark is walked).  Results do not depend on it, speed does."""
        _native.check(self._lib.sitb_set_candidate_grid(self._ctx, float(static_margin)))

    def candidate_grid_info(self):
        dims = (C.c_int32 * 3)()
        margin, n = C.c_double(), C.c_uint64()
        _native.check(self._lib.sitb_candidate_grid_info(self._ctx, dims, C.byref(margin), C.byref(n)))
        return {"dims": tuple(dims), "static_margin": margin.value, "entries": n.value}

    @classmethod
    def from_site_network(cls, sn, **kw):
        cell = np.asarray(sn.structure.cell)
        static_idx = np.where(sn.static_mask)[0]
        mobile_idx = np.where(sn.mobile_mask)[0]
        return cls(cell, static_idx, mobile_idx, sn.n_total, sn.static_structure.get_positions(),
                   np.asarray(sn.centers), sn.vertices, **kw)

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._lib.sitb_destroy(self._ctx)      # synchronises the compute and copy streams first
            self._ctx = C.c_void_p()
            self._frames_keepalive = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing -------------------------------------------------------------------------
    def use_current_stream(self):
        torch = _torch()
        s = torch.cuda.current_stream(self.device).cuda_stream
        _native.check(self._lib.sitb_set_stream(self._ctx, C.c_void_p(s)))

    def _empty(self, shape, dtype):
        return _torch().empty(shape, dtype=dtype, device=self.device)

    def _zeros(self, shape, dtype):
        return _torch().zeros(shape, dtype=dtype, device=self.device)

    @staticmethod
    def _ptr(t):
        return C.c_void_p(0 if t is None else t.data_ptr())

    def tables(self):
        svd = np.empty((self.L, self.V), dtype=np.float64)
        q = np.empty((self.L, self.V), dtype=np.float64)
        _native.check(self._lib.sitb_get_tables(self._ctx, svd.ctypes.data, q.ctypes.data))
        return svd, q

    def set_frames(self, frames, frame0=0):
        """Make a shard of frames resident.  numpy float64 (F, A, 3) is copied to the device once;
        a torch CUDA float64 tensor is borrowed in place."""
        torch = _torch()
        if isinstance(frames, torch.Tensor):
            if frames.device != self.device or frames.dtype != torch.float64 or not frames.is_contiguous():
                raise ValueError("device frames must be a contiguous float64 tensor on %s" % self.device)
            if frames.shape[1:] != (self.n_atoms, 3):
                raise ValueError("Wrong shape %s for frames." % (tuple(frames.shape),))
            _native.check(self._lib.sitb_borrow_frames(self._ctx, self._ptr(frames), frames.shape[0], frame0))
            self._frames_keepalive = frames
            self.frames_bytes = 0
        else:
            frames = np.asarray(frames)
            if frames.dtype not in (np.float64, np.float32):
                raise ValueError("frames must be float64 (as in the reference) or float32")
            if frames.shape[1:] != (self.n_atoms, 3):
                raise ValueError("Wrong shape %s for frames." % (frames.shape,))
            frames = np.ascontiguousarray(frames)
            # asynchronous when the host array is page-locked: the passes wait for the chunks they read, and the
            # array is kept alive here until the next set_frames()/close().  float32 sources (np.memmap'ed MD dumps
            # included) cross PCIe at half the size and are widened on the device.
            if frames.dtype == np.float32:
                _native.check(self._lib.sitb_upload_frames_f32(self._ctx, frames.ctypes.data, frames.shape[0], frame0))
                self.frames_bytes = max(self.frames_bytes, int(frames.nbytes) * 3)
            else:
                _native.check(self._lib.sitb_upload_frames(self._ctx, frames.ctypes.data, frames.shape[0], frame0))
                self.frames_bytes = max(self.frames_bytes, int(frames.nbytes))
            self._frames_keepalive = frames
        self.n_frames = int(frames.shape[0])
        self.frame0 = int(frame0)

    def unassigned_lattice_atoms(self, global_frame):
        """Error reporting only (helpers.pyx:87-92): the static atoms of one frame that no static-lattice position
        picked as its nearest atom, for ``StaticLatticeError.lattice_atoms``.  One frame, evaluated on the host with
        the reference's own steps (PBCCalculator.pyx:64-103); None if the frame is not resident on this rank."""
        local = int(global_frame) - self.frame0
        fr = self._frames_keepalive
        if fr is None or not (0 <= local < self.n_frames):
            return None
        frame = fr[local].cpu().numpy() if hasattr(fr, "cpu") else np.asarray(fr[local], dtype=np.float64)
        statics = np.asarray(frame, dtype=np.float64)[self.static_idx]
        seen = np.zeros(self.S, dtype=bool)
        for li in range(self.S):
            buf = statics + (self.cell_centroid - self.ideal_static[li])           # shift the lattice point to the centroid
            f = buf @ self.cellmat_inv.T
            f -= np.floor(f)
            buf = f @ self.cellmat.T
            d = np.sqrt(np.sum((buf - self.cell_centroid) ** 2, axis=1))
            seen[int(np.argmin(d))] = True
        return np.where(~seen)[0]

    def upload_chunk_frames(self):
        """Frames per chunk of the last set_frames() upload (0: frames borrowed from a device tensor)."""
        n = C.c_int64()
        _native.check(self._lib.sitb_upload_chunk_frames(self._ctx, C.byref(n)))
        return int(n.value)

    def reset_status(self):
        _native.check(self._lib.sitb_reset_status(self._ctx))

    def status(self):
        s = _native.Status()
        _native.check(self._lib.sitb_get_status(self._ctx, C.byref(s)))
        return EngineStatus(s)

    # ---- passes ---------------------------------------------------------------------------
    def fill_dense(self, begin=0, n=None, dtype=None, out=None):
        torch = _torch()
        n = self.n_frames - begin if n is None else n
        dtype = torch.float32 if dtype is None else dtype
        if out is None:
            out = self._empty((n * self.M, self.L), dtype)
        _native.check(self._lib.sitb_fill_dense(self._ctx, begin, n, self._ptr(out), int(dtype == torch.float64)))
        return out

    def fill_frames(self, frame_indices, dtype=None):
        """Landmark vectors of selected resident frames: (len(frame_indices)*M, L)."""
        torch = _torch()
        dtype = torch.float32 if dtype is None else dtype
        idx = torch.as_tensor(np.asarray(frame_indices, dtype=np.int64), device=self.device)
        out = self._empty((len(idx) * self.M, self.L), dtype)
        _native.check(self._lib.sitb_fill_dense_frames(self._ctx, self._ptr(idx), len(idx), self._ptr(out),
                                                       int(dtype == torch.float64)))
        return out

    def pass_stats(self, begin=0, n=None, seen=None, gram=None):
        torch = _torch()
        n = self.n_frames - begin if n is None else n
        if seen is None:
            seen = self._zeros((self.L,), torch.int64)
        if gram is None:
            gram = self._zeros((self.L, self.L), torch.float64)
        _native.check(self._lib.sitb_pass_stats(self._ctx, begin, n, self._ptr(seen), self._ptr(gram)))
        return seen, gram

    def pass_stats_tc(self, begin=0, n=None, seen=None, gram=None, block_frames=16384):
        """Pass A with the Gram on tcgen05 tensor cores: K1 stages each block of frames as transposed fp16
        hi/lo tiles, sitb_gram_syrk_tc accumulates them into the FP64 upper triangle.  Same outputs as
        :meth:`pass_stats`; the Gram agrees to ~1e-7 relative (fp16 hi+lo operands drop the lo.lo term)."""
        torch = _torch()
        n = self.n_frames - begin if n is None else n
        if seen is None:
            seen = self._zeros((self.L,), torch.int64)
        if gram is None:
            gram = self._zeros((self.L, self.L), torch.float64)
        block_frames = max(1, min(block_frames, n))
        ld = -(-(block_frames * self.M) // 64) * 64
        lpad = -(-self.L // 128) * 128
        hi = self._empty((lpad, ld), torch.float16)
        lo = self._empty((lpad, ld), torch.float16)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        for b in range(begin, begin + n, block_frames):
            nb = min(block_frames, begin + n - b)
            hi.zero_()
            lo.zero_()
            _native.check(self._lib.sitb_pass_stage(self._ctx, b, nb, self._ptr(seen), self._ptr(hi), self._ptr(lo), ld))
            _native.check(self._lib.sitb_gram_syrk_tc(self.device.index, self._ptr(hi), self._ptr(lo), self.L,
                                                      lpad, ld, nb * self.M, self._ptr(gram),
                                                      C.c_void_p(stream)))
        return seen, gram

    def gram_words_finish(self, words):
        """Deterministic Gram: the integer word matrix (2 (L + 1), L) -> float64 (L, L) upper triangle."""
        torch = _torch()
        out = self._empty((self.L, self.L), torch.float64)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _native.check(self._lib.sitb_gram_words_finish(self.device.index, self._ptr(words), self.L, self._ptr(out),
                                                       C.c_void_p(stream)))
        return out

    def pass_stats_cached(self, seen=None, gram=None, entries_per_row=40, gram_from_rows=None, want_gram=True,
                          gram_words=False):
        """Pass A that also caches every landmark vector compressed (SparseRows); grows the pool on overflow.
        ``gram_from_rows`` (default: whenever the shared-memory tables fit, L <= 8192): build the Gram from the
        cached rows per (atom, window of frames) instead of with one atomic per pair product inside K1."""
        if not want_gram:                       # rows and seen counts only (the dotprod plugin)
            gram_from_rows = True
        elif gram_from_rows is None:
            gram_from_rows = self.L <= 8192
        torch = _torch()
        n_rows = self.n_frames * self.M
        if seen is None:
            seen = self._zeros((self.L,), torch.int64)
        gram_words = bool(gram_words and gram_from_rows and want_gram)
        if gram is None and want_gram:
            # gram_words: the deterministic integer form (sitb_gram_words_from_cached), finished by gram_words_finish
            gram = self._zeros((2 * (self.L + 1), self.L), torch.int64) if gram_words else self._zeros((self.L, self.L), torch.float64)
        # Rows of up to ROW_SLOT entries live in fixed, row-ordered slots (the passes over the cached rows then stream
        # them; scattered over a pool they cost a DRAM page per row); longer rows take space behind the slots.
        slot = self.ROW_SLOT
        overflow_per_row = 0.02 * entries_per_row
        while True:
            # (warps reserve the overflow space in slices of 256 entries and leave the tail of a slice unused)
            step = self.upload_chunk_frames() or self.n_frames
            n_launches = (self.n_frames + step - 1) // step
            base = n_rows * slot
            cap = base + int(n_rows * overflow_per_row) + 1024 + 256 * 32 * self.n_sms * n_launches
            cursor = self._empty((1,), torch.int64)
            cursor.fill_(base)
            rows = SparseRows(self._empty((n_rows,), torch.int64), self._empty((cap,), torch.int16),
                              self._empty((cap,), torch.float64), cursor, cap, n_rows,
                              self.frame0 * self.M)
            seen_try = seen.clone()
            gram_try = None if gram_from_rows else gram.clone()
            # launched per upload chunk: each launch waits only for its own chunk of the host -> device copy
            for b in range(0, self.n_frames, step):
                nb = min(step, self.n_frames - b)
                _native.check(self._lib.sitb_pass_stats_slotted(
                    self._ctx, b, nb, self._ptr(seen_try), self._ptr(gram_try),
                    C.c_void_p(rows.ptr.data_ptr() + 8 * b * self.M), self._ptr(rows.k), self._ptr(rows.v),
                    self._ptr(rows.cursor), cap, slot))
            used = int(rows.cursor.item())
            if used <= cap:
                seen.copy_(seen_try)
                if not want_gram:
                    pass
                elif gram_from_rows:
                    fn = self._lib.sitb_gram_words_from_cached if gram_words else self._lib.sitb_gram_from_cached
                    _native.check(fn(self._ctx, self._ptr(rows.ptr), self._ptr(rows.k), self._ptr(rows.v), self.n_frames,
                                     self._ptr(gram)))
                else:
                    gram.copy_(gram_try)
                rows.used = used
                return seen, gram, rows
            overflow_per_row = (used - base) / float(n_rows) * 1.05 + 1     # exact requirement is known now
            self.reset_status()

    def assign_sparse(self, rows, threshold, labels=None, confs=None, counts=None, best=None, rep=None, rep_w=None,
                      site_best=None):
        _native.check(self._lib.sitb_assign_sparse(
            self._ctx, self._ptr(rows.ptr), self._ptr(rows.k), self._ptr(rows.v), rows.n_rows, rows.row0,
            float(threshold), self._ptr(labels), self._ptr(confs), self._ptr(counts), self._ptr(best), self._ptr(rep),
            self._ptr(rep_w), self._ptr(site_best)))

    def repredict_removed(self, rows, threshold, remap, labels, confs, rep=None, rep_w=None, site_best=None):
        """After the min_samples filter (DotProdClassifier.pyx:105-118): renumber ``labels`` (device int64) by ``remap``
        (new id per old cluster, -1 = removed) and predict only the rows of removed clusters again, with the current
        centres.  Returns the device counter of such rows."""
        torch = _torch()
        remap_d = torch.as_tensor(np.ascontiguousarray(remap, dtype=np.int32), device=self.device)
        row_list = self._empty((rows.n_rows,), torch.int64)
        n_list = self._zeros((1,), torch.int64)
        _native.check(self._lib.sitb_relabel_select(self._ctx, self._ptr(labels), rows.n_rows, self._ptr(remap_d),
                                                    self._ptr(row_list), self._ptr(n_list)))
        _native.check(self._lib.sitb_assign_sparse_rows(
            self._ctx, self._ptr(rows.ptr), self._ptr(rows.k), self._ptr(rows.v), self._ptr(row_list), self._ptr(n_list),
            rows.n_rows, rows.row0, float(threshold), self._ptr(labels), self._ptr(confs), self._ptr(None), self._ptr(None),
            self._ptr(rep), self._ptr(rep_w), self._ptr(site_best)))
        return n_list

    def set_centers(self, cluster_of_landmark, weight, n_clusters):
        cid = np.ascontiguousarray(cluster_of_landmark, dtype=np.int32)
        w = np.ascontiguousarray(weight, dtype=np.float64)
        assert cid.shape == (self.L,) and w.shape == (self.L,)
        _native.check(self._lib.sitb_set_centers(self._ctx, cid.ctypes.data, w.ctypes.data, int(n_clusters)))
        self.n_clusters = int(n_clusters)

    ROW_SLOT = 32          # entries per fixed slot of a cached row (pass_stats_cached)

    TWO_TIER_REASONS = ("frame", "support", "margin", "threshold", "long", "rows")

    def set_assign_mode(self, mode):
        """'exact': every component in float64 (default).  'two_tier': FP32 first tier with a proven error bound, the
        exact kernel only for rows whose decisions lie inside the bound -- identical labels, confidences to ~1e-5."""
        _native.check(self._lib.sitb_set_assign_mode(self._ctx, {"exact": 0, "two_tier": 1}[mode]))
        self.assign_mode = mode

    def two_tier_info(self, reset=False):
        avail, tau, kappa = C.c_int32(), C.c_double(), C.c_double()
        counts = (C.c_uint64 * 6)()
        _native.check(self._lib.sitb_two_tier_info(self._ctx, C.byref(avail), C.byref(tau), C.byref(kappa), counts, int(reset)))
        out = {"available": bool(avail.value), "tau": tau.value, "kappa": kappa.value}
        out.update({"recheck_" + k: int(v) for k, v in zip(self.TWO_TIER_REASONS, counts)})
        return out

    def pass_assign(self, threshold, begin=0, n=None, labels=None, confs=None, counts=None, best=None,
                    rep=None, rep_w=None, site_best=None):
        n = self.n_frames - begin if n is None else n
        _native.check(self._lib.sitb_pass_assign(
            self._ctx, begin, n, float(threshold), self._ptr(labels), self._ptr(confs), self._ptr(counts),
            self._ptr(best), self._ptr(rep), self._ptr(rep_w), self._ptr(site_best)))

    # ---- site centres (LandmarkAnalysis.py:276-299) -------------------------------------------------
    def wrapped_mobile_rows(self, global_rows):
        """(n, 3) wrapped positions of the given global (frame * M + mobile) rows; zeros where not resident."""
        torch = _torch()
        rows = torch.as_tensor(np.asarray(global_rows, dtype=np.int64), device=self.device)
        out = self._empty((len(rows), 3), torch.float64)
        _native.check(self._lib.sitb_wrapped_mobile_rows(self._ctx, self._ptr(rows), len(rows), self._ptr(out)))
        return out

    def site_centers(self, labels, confs, n_sites, weighted, site_best=None, comm=None):
        """PBCCalculator.average (PBCCalculator.pyx:106-139) of the wrapped mobile positions assigned to
        each site: (n_sites, 3) float64 numpy."""
        torch = _torch()
        if weighted:
            # centring point = the max-confidence row, first maximum (np.argmax, PBCCalculator.pyx:120-122)
            assert site_best is not None
            _, anchor_rows = read_best_table(site_best, comm)
        else:
            first = torch.full((n_sites,), -1, dtype=torch.int64, device=self.device)     # all ones
            _native.check(self._lib.sitb_site_first_rows(self._ctx, self._ptr(labels), n_sites, self._ptr(first)))
            if comm is not None:
                comm.allreduce_min_u64_(first)
            anchor_rows = first.cpu().numpy().view(np.uint64).astype(np.int64)
        anchors = self.wrapped_mobile_rows(anchor_rows)
        if comm is not None:
            comm.allreduce_sum_(anchors)
        centroid = torch.as_tensor(self.cell_centroid, device=self.device)
        offsets = (centroid[None, :] - anchors).contiguous()
        sums = self._zeros((n_sites, 4), torch.float64)
        _native.check(self._lib.sitb_site_accumulate(self._ctx, self._ptr(labels), self._ptr(confs),
                                                     self._ptr(offsets), n_sites, int(bool(weighted)), self._ptr(sums)))
        if comm is not None:
            comm.allreduce_sum_(sums)
        out = self._empty((n_sites, 3), torch.float64)
        _native.check(self._lib.sitb_site_finish(self._ctx, self._ptr(sums), self._ptr(offsets), n_sites, self._ptr(out)))
        return out.cpu().numpy()

    def weighted_point_averages(self, points, weights):
        """Per row of ``weights`` (n_sites, n_points): periodic weighted average of ``points`` with weight > 0."""
        torch = _torch()
        pts = torch.as_tensor(np.array(points, dtype=np.float64, order="C", copy=True), device=self.device)
        w = torch.as_tensor(np.ascontiguousarray(weights, dtype=np.float64), device=self.device)
        out = self._empty((w.shape[0], 3), torch.float64)
        _native.check(self._lib.sitb_weighted_point_average(self._ctx, self._ptr(pts), self._ptr(w), w.shape[0],
                                                            w.shape[1], self._ptr(out)))
        return out.cpu().numpy()

    def fill_landmark_vectors_host(self, frames):
        """Host in, host out: the drop-in for ``helpers._fill_landmark_vectors`` (helpers.pyx:12)."""
        frames = np.ascontiguousarray(frames, dtype=np.float64)
        if frames.shape[1:] != (self.n_atoms, 3):
            raise ValueError("Wrong shape %s for frames." % (frames.shape,))
        out = np.empty((frames.shape[0] * self.M, self.L), dtype=np.float64)
        s = _native.Status()
        _native.check(self._lib.sitb_fill_landmark_vectors_host(
            self._ctx, frames.ctypes.data, frames.shape[0], out.ctypes.data, C.byref(s)))
        return out, EngineStatus(s)


class SparseRows(object):
    """Compressed landmark vectors of the resident frames (device tensors)."""

    def __init__(self, ptr, k, v, cursor, capacity, n_rows, row0):
        self.ptr, self.k, self.v, self.cursor = ptr, k, v, cursor
        self.capacity, self.n_rows, self.row0 = capacity, n_rows, row0
        self.used = 0


def new_best_table(n, device):
    """Zeroed (3n,) int64 table for the assign pass's lexicographic max of (value, first row)."""
    return _torch().zeros((3 * n,), dtype=_torch().int64, device=device)


def read_best_table(tab, comm=None):
    """(float64 values, int64 rows) of a best table; with ``comm`` the lexicographic max over ranks."""
    n = tab.shape[0] // 3
    h = tab[:2 * n].cpu().numpy()
    vals = h[:n].view(np.float64).copy()
    rows = h[n:].copy()
    if comm is not None:
        allv = comm.allgather_numpy(vals)
        allr = comm.allgather_numpy(rows)
        order = np.lexsort((allr, -allv), axis=0)[0]          # max value, then lowest row
        vals = np.take_along_axis(allv, order[None, :], 0)[0]
        rows = np.take_along_axis(allr, order[None, :], 0)[0]
    return vals, rows
