"""CPU: the C-ABI library builds, loads, exports every symbol include/sitator_b200.h declares, and fails
loudly without a CUDA device (no compute calls here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "sitator_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sitb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from sitator_b200.build import build_native
    build_native()
    from sitator_b200 import _native
    lib = _native.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
        assert n in _native.SIGNATURES, "no ctypes prototype for %s" % n
    assert lib.sitb_version() >= 100


def test_every_declared_entry_point_is_documented_in_the_integration_guide():
    """INTEGRATION.md says which reference lines each ABI entry replaces; a new entry point must get its row."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in _declared() if n not in doc]
    assert not missing, "not mentioned in INTEGRATION.md: %s" % missing


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sitator_b200 import _native
    lib = _native.load()
    d = _native.NetworkDesc()
    d.n_atoms = d.n_static = d.n_mobile = d.n_landmarks = d.max_verts = 1
    a = np.eye(3).ravel().copy()
    i = np.zeros(4, dtype=np.int32)
    for f in ("host_cellmat", "host_ideal_static", "host_centers"):
        setattr(d, f, a.ctypes.data)
    for f in ("host_static_idx", "host_mobile_idx", "host_verts"):
        setattr(d, f, i.ctypes.data)
    out = C.c_void_p()
    rc = lib.sitb_create(C.byref(d), 0, C.byref(out))
    assert rc == -2 and b"no CPU path" in lib.sitb_last_error()
    from sitator_b200.engine import LandmarkEngine
    with pytest.raises(RuntimeError, match="no CPU path"):
        LandmarkEngine(np.eye(3), [0], [1], 2, np.zeros((1, 3)), np.zeros((1, 3)), [[0]])
    from sitator_b200.util.mcl import markov_clustering
    with pytest.raises(RuntimeError, match="no CPU path"):
        markov_clustering(np.eye(3))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sitator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, f)


def test_struct_sizes_of_every_binding_match_the_library():
    """sitb_get_status writes the whole struct: a binding with a shorter Status would be overrun (round-1 advice).  The
    in-package binding, the reference-side binding INTEGRATION.md quotes, and the header must agree."""
    from sitator_b200.build import build_native
    build_native()
    from sitator_b200 import _native
    lib = _native.load()
    sizes = (C.c_uint64 * 2)()
    assert lib.sitb_abi_sizes(sizes) == 0
    assert (sizes[0], sizes[1]) == (C.sizeof(_native.NetworkDesc), C.sizeof(_native.Status))
    from sitator_b200.integration import reference_binding as rb      # asserts the same at import
    assert (C.sizeof(rb.Desc), C.sizeof(rb.Status)) == (sizes[0], sizes[1])
    # INTEGRATION.md shows that file's struct definitions verbatim
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    src = open(os.path.join(ROOT, "sitator_b200", "integration", "reference_binding.py")).read()
    block = src[src.index("class Status(C.Structure):"):src.index("_sizes = ")].strip()
    assert block in doc, "INTEGRATION.md no longer quotes the Status struct of reference_binding.py"
