"""Drop-in for the reference's Cython fill, over the C ABI only (ctypes; no torch, no sitator_b200 Python).

``_fill_landmark_vectors`` has the signature and the side effects of
``sitator.landmark.helpers._fill_landmark_vectors`` (``sitator/landmark/helpers.pyx:12-124``): it fills
``self._landmark_vectors`` in place, sets ``self.n_all_zero_lvecs`` and raises the reference's own exception types.
A maintainer binds it with one line in ``sitator/landmark/LandmarkAnalysis.py``::

    from sitator_b200.integration.reference_binding import _fill_landmark_vectors   # instead of helpers._fill_...

or, without touching sitator, ``sitator.landmark.helpers._fill_landmark_vectors = _fill_landmark_vectors``
(the call site ``LandmarkAnalysis.py:220`` looks the function up on the module).  The struct layouts below are those of
``include/sitator_b200.h``; ``sitb_abi_sizes`` guards them at import.
"""
import ctypes as C
import math
import os

import numpy as np

_LIB_PATH = os.environ.get("SITB_LIB") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        "lib", "libsitator_b200.so")
lib = C.CDLL(_LIB_PATH)
lib.sitb_last_error.restype = C.c_char_p


class Desc(C.Structure):          # sitb_network_desc
    _fields_ = [(n, C.c_int32) for n in ("n_atoms", "n_static", "n_mobile", "n_landmarks", "max_verts")] + \
               [(n, C.c_void_p) for n in ("cellmat", "cellmat_inv", "static_idx", "mobile_idx", "ideal", "centers", "verts")] + \
               [(n, C.c_double) for n in ("mid", "steep", "cutoff", "static_thr")] + \
               [("dynamic", C.c_int32), ("relaxed", C.c_int32)]


class Status(C.Structure):        # sitb_status
    _fields_ = [("error_code", C.c_int32), ("index", C.c_int32), ("frame", C.c_int64),
                ("zero_error", C.c_int32), ("zero_index", C.c_int32), ("zero_frame", C.c_int64)] + \
               [(n, C.c_uint64) for n in ("n_zero_rows", "n_dup", "n_overflow", "nnz", "n_screen_rejects",
                                          "n_full_walk_frames", "n_loose_grid_frames")]


_sizes = (C.c_uint64 * 2)()
lib.sitb_abi_sizes(_sizes)
if (int(_sizes[0]), int(_sizes[1])) != (C.sizeof(Desc), C.sizeof(Status)):
    raise ImportError("libsitator_b200.so has sitb_network_desc / sitb_status of %d / %d bytes, this binding %d / %d"
                      % (_sizes[0], _sizes[1], C.sizeof(Desc), C.sizeof(Status)))


def _fill_landmark_vectors(self, sn, verts_np, site_vert_dists, frames, check_for_zeros=True, tqdm=None, logger=None):
    from sitator.landmark import StaticLatticeError, ZeroLandmarkError          # the reference's own exception types
    cellmat = np.ascontiguousarray(np.asarray(sn.structure.cell, dtype=np.float64).T)   # PBCCalculator.pyx:33-34
    cellinv = np.ascontiguousarray(np.linalg.inv(cellmat))
    verts = np.ascontiguousarray(verts_np, dtype=np.int32)
    keep = dict(sidx=np.where(sn.static_mask)[0].astype(np.int32), midx=np.where(sn.mobile_mask)[0].astype(np.int32),
                ideal=np.ascontiguousarray(sn.static_structure.get_positions(), dtype=np.float64),
                cen=np.ascontiguousarray(sn.centers, dtype=np.float64))
    d = Desc(sn.n_total, sn.n_static, sn.n_mobile, len(verts), verts.shape[1],
             cellmat.ctypes.data, cellinv.ctypes.data, keep["sidx"].ctypes.data, keep["midx"].ctypes.data,
             keep["ideal"].ctypes.data, keep["cen"].ctypes.data, verts.ctypes.data,
             self._cutoff_midpoint, self._cutoff_steepness,
             self._cutoff_midpoint + math.log((1 / 0.0001) - 1.) / self._cutoff_steepness,   # helpers.pyx:127-131
             self.static_movement_threshold, int(self.dynamic_lattice_mapping), int(self.relaxed_lattice_checks))
    ctx, st = C.c_void_p(), Status()
    if lib.sitb_create(C.byref(d), 0, C.byref(ctx)):
        raise RuntimeError(lib.sitb_last_error().decode())
    try:                           # frames (F, A, 3) float64 in host memory -> self._landmark_vectors (F*M, L) float64
        frames = np.ascontiguousarray(frames, dtype=np.float64)
        lv = self._landmark_vectors
        out = lv if (isinstance(lv, np.ndarray) and type(lv) is np.ndarray and lv.flags.c_contiguous) else np.empty(lv.shape, dtype=np.float64)
        if lib.sitb_fill_landmark_vectors_host(ctx, C.c_void_p(frames.ctypes.data), C.c_int64(len(frames)),
                                               C.c_void_p(out.ctypes.data), C.byref(st)):
            raise RuntimeError(lib.sitb_last_error().decode())
        if out is not lv:
            lv[:] = out            # e.g. the reference's np.memmap (LandmarkAnalysis.py:211-218)
    finally:
        lib.sitb_destroy(ctx)
    cands = []
    if st.error_code:
        cands.append((st.frame, st.error_code, st.index))
    if check_for_zeros and st.zero_error:
        cands.append((st.zero_frame, 3, st.zero_index))
    if cands:
        frame, code, index = min(cands)                    # first error in the reference's iteration order
        if code in (1, 2):
            raise StaticLatticeError("No static atom position within %f A threshold of static lattice position %i"
                                     % (self.static_movement_threshold, index) if code == 1 else
                                     "At frame %i, static positions of some atoms not assigned to lattice positions" % frame,
                                     lattice_atoms=[index], frame=frame, try_recentering=True)
        raise ZeroLandmarkError(mobile_index=index, frame=frame)
    if st.n_dup and logger is not None:
        logger.warning("%i times a static atom was the closest to more than one static lattice position" % st.n_dup)
    self.n_all_zero_lvecs = int(st.n_zero_rows)
