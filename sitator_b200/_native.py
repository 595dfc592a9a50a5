"""ctypes binding of ``lib/libsitator_b200.so`` (declared in ``include/sitator_b200.h``).

There is no fallback: if the library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

# SITB_LIB: developer override, to time an experimental build of the same ABI against the in-tree one
_LIB_PATH = os.environ.get("SITB_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libsitator_b200.so")
_lib = None


class NativeError(RuntimeError):
    """A sitb_* call failed (code + sitb_last_error())."""

    def __init__(self, code, message):
        super().__init__("sitator_b200 native call failed (%d): %s" % (code, message))
        self.code = code


class NetworkDesc(C.Structure):
    _fields_ = [
        ("n_atoms", C.c_int32), ("n_static", C.c_int32), ("n_mobile", C.c_int32),
        ("n_landmarks", C.c_int32), ("max_verts", C.c_int32),
        ("host_cellmat", C.c_void_p), ("host_cellmat_inv", C.c_void_p),
        ("host_static_idx", C.c_void_p), ("host_mobile_idx", C.c_void_p),
        ("host_ideal_static", C.c_void_p), ("host_centers", C.c_void_p), ("host_verts", C.c_void_p),
        ("cutoff_midpoint", C.c_double), ("cutoff_steepness", C.c_double),
        ("cutoff_round_to_zero", C.c_double), ("static_movement_threshold", C.c_double),
        ("dynamic_lattice_mapping", C.c_int32), ("relaxed_lattice_checks", C.c_int32),
    ]


class Status(C.Structure):
    _fields_ = [
        ("error_code", C.c_int32), ("index", C.c_int32), ("frame", C.c_int64),
        ("zero_error", C.c_int32), ("zero_index", C.c_int32), ("zero_frame", C.c_int64),
        ("n_zero_rows", C.c_uint64), ("n_duplicate_nearest", C.c_uint64),
        ("n_list_overflow", C.c_uint64), ("nnz", C.c_uint64), ("n_screen_rejects", C.c_uint64),
        ("n_full_walk_frames", C.c_uint64), ("n_loose_grid_frames", C.c_uint64),
    ]


_P = C.c_void_p
# name -> (restype, argtypes); every symbol include/sitator_b200.h declares
SIGNATURES = {
    "sitb_last_error": (C.c_char_p, []),
    "sitb_version": (C.c_int, []),
    "sitb_abi_sizes": (C.c_int, [C.POINTER(C.c_uint64)]),
    "sitb_create": (C.c_int, [C.POINTER(NetworkDesc), C.c_int, C.POINTER(_P)]),
    "sitb_destroy": (None, [_P]),
    "sitb_set_stream": (C.c_int, [_P, _P]),
    "sitb_set_candidate_grid": (C.c_int, [_P, C.c_double]),
    "sitb_candidate_grid_info": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "sitb_device_info": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "sitb_get_tables": (C.c_int, [_P, _P, _P]),
    "sitb_upload_frames": (C.c_int, [_P, _P, C.c_int64, C.c_int64]),
    "sitb_upload_frames_f32": (C.c_int, [_P, _P, C.c_int64, C.c_int64]),
    "sitb_borrow_frames": (C.c_int, [_P, _P, C.c_int64, C.c_int64]),
    "sitb_reset_status": (C.c_int, [_P]),
    "sitb_get_status": (C.c_int, [_P, C.POINTER(Status)]),
    "sitb_fill_dense": (C.c_int, [_P, C.c_int64, C.c_int64, _P, C.c_int32]),
    "sitb_fill_dense_frames": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int32]),
    "sitb_pass_stats": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P]),
    "sitb_pass_stats_cached": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, _P, _P, _P, _P, C.c_uint64]),
    "sitb_pass_stats_slotted": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, _P, _P, _P, _P, C.c_uint64, C.c_int32]),
    "sitb_gram_from_cached": (C.c_int, [_P, _P, _P, _P, C.c_int64, _P]),
    "sitb_gram_words_from_cached": (C.c_int, [_P, _P, _P, _P, C.c_int64, _P]),
    "sitb_gram_words_finish": (C.c_int, [C.c_int, _P, C.c_int32, _P, _P]),
    "sitb_assign_sparse": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int64, C.c_double] + [_P] * 7),
    "sitb_sparse_row_norm2": (C.c_int, [_P, _P, _P, C.c_int64, _P, C.c_int32, _P]),
    "sitb_relabel_select": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P]),
    "sitb_assign_sparse_rows": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_double] + [_P] * 7),
    "sitb_set_centers": (C.c_int, [_P, _P, _P, C.c_int32]),
    "sitb_pass_assign": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_double] + [_P] * 7),
    "sitb_set_assign_mode": (C.c_int, [_P, C.c_int32]),
    "sitb_two_tier_info": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                     C.POINTER(C.c_uint64), C.c_int32]),
    "sitb_dotprod_limits": (C.c_int, [C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "sitb_dotprod_fit": (C.c_int, [C.c_int, _P, _P, _P, C.c_int64, C.c_int32, C.c_double, C.c_int32, C.c_int32,
                                   _P, _P, _P, _P, _P, _P, _P]),
    "sitb_dotprod_predict": (C.c_int, [C.c_int, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, C.c_double,
                                       _P, _P, _P, _P]),
    "sitb_landmark_graph": (C.c_int, [C.c_int, _P, C.c_int32, C.c_double, _P, _P, _P]),
    "sitb_principal_vectors": (C.c_int, [C.c_int, _P, C.c_int32, _P, _P, C.c_int32, _P, _P, _P]),
    "sitb_markov_clustering": (C.c_int, [C.c_int, _P, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32, _P,
                                         C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P]),
    "sitb_wrapped_mobile_rows": (C.c_int, [_P, _P, C.c_int32, _P]),
    "sitb_site_first_rows": (C.c_int, [_P, _P, C.c_int32, _P]),
    "sitb_site_accumulate": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, _P]),
    "sitb_site_finish": (C.c_int, [_P, _P, _P, C.c_int32, _P]),
    "sitb_weighted_point_average": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P]),
    "sitb_check_multiple_occupancy": (C.c_int, [C.c_int, _P, C.c_int64, C.c_int32, C.c_int64, C.c_int32, _P, _P, _P]),
    "sitb_jump_scan": (C.c_int, [C.c_int, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "sitb_jump_compact": (C.c_int, [C.c_int, _P, _P, C.c_int64, C.c_int32, C.c_int64, _P, C.c_uint64, _P]),
    "sitb_jump_analysis": (C.c_int, [C.c_int, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32] + [_P] * 8),
    "sitb_jump_analysis_summary": (C.c_int, [C.c_int, _P, C.c_int64, C.c_int32, _P, _P]),
    "sitb_recenter": (C.c_int, [C.c_int, _P, C.c_int64, C.c_int32, _P, _P, _P]),
    "sitb_assign_last_known": (C.c_int, [C.c_int, _P, C.c_int64, C.c_int32, C.c_int64, C.c_int64, _P, _P, _P, _P, _P,
                                         C.c_int32, _P]),
    "sitb_windowed_mode": (C.c_int, [C.c_int, _P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int32,
                                     C.c_int64, _P, C.c_int64, _P, _P]),
    "sitb_seen_sites": (C.c_int, [C.c_int, _P, C.c_int64, C.c_int32, _P, _P]),
    "sitb_relabel_sites": (C.c_int, [C.c_int, _P, C.c_int64, C.c_int32, _P, _P]),
    "sitb_pbc_distances": (C.c_int, [C.c_int, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P]),
    "sitb_pbc_weighted_average": (C.c_int, [C.c_int, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P]),
    "sitb_upload_chunk_frames": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "sitb_pass_stage": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, _P, C.c_int64]),
    "sitb_gram_syrk_tc": (C.c_int, [C.c_int, _P, _P, C.c_int32, C.c_int32, C.c_int64, C.c_int64, _P, _P]),
    "sitb_microbench": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "sitb_fill_landmark_vectors_host": (C.c_int, [_P, _P, C.c_int64, _P, C.POINTER(Status)]),
}


def library_path():
    return _LIB_PATH


def load():
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(_LIB_PATH):
        raise ImportError(
            "sitator_b200: %s is missing. Build it with `python -m sitator_b200.build` "
            "(needs nvcc; there is no CPU fallback)." % _LIB_PATH)
    lib = C.CDLL(_LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    sizes = (C.c_uint64 * 2)()
    lib.sitb_abi_sizes(sizes)
    if (int(sizes[0]), int(sizes[1])) != (C.sizeof(NetworkDesc), C.sizeof(Status)):
        raise ImportError("sitator_b200: %s was built from a different include/sitator_b200.h (struct sizes %d / %d, "
                          "binding %d / %d); rebuild with `python -m sitator_b200.build --force`"
                          % (_LIB_PATH, sizes[0], sizes[1], C.sizeof(NetworkDesc), C.sizeof(Status)))
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise NativeError(code, load().sitb_last_error().decode("utf-8", "replace"))
