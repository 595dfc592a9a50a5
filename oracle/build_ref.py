#!/usr/bin/env python
"""Build recipe for ``oracle/_ref`` -- the UNMODIFIED reference, compiled to binaries.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sitator_b200/`` imports this.

The reference (sitator v2.0.0) is Python + Cython.  Its landmark-analysis path
is compiled here *from the sources where they lie* (``/root/reference``, or
``$SITATOR_REFERENCE``) into CPython extension modules under
``oracle/_ref/sitator/...``: the ``.pyx`` files (as the reference's own
``setup.py:18-26`` does) and also the pure-Python modules on the path, so that
``oracle/_ref`` holds binaries only and no reference source is ever copied
into this repository.  ``oracle/_ref`` is git-ignored but travels to the GPU
box with the snapshot, where ``/root/reference`` does not exist.

We do not run the reference's ``setup.py``: ``sitator/misc/GenerateClampedTrajectory.pyx``
does not compile under Cython 3 (``:117``) and is not on the path.  The list
below is the import closure of ``LandmarkAnalysis.run`` + ``JumpAnalysis.run``.

Usage:  python oracle/build_ref.py [--force]
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
BUILD = os.path.join(OUT, "_build")

# (module path relative to the reference root)  -- import closure of the hot path
MODULES = [
    "sitator/__init__.py",
    "sitator/errors.py",
    "sitator/SiteNetwork.py",
    "sitator/SiteTrajectory.py",
    "sitator/util/__init__.py",
    "sitator/util/progress.py",
    "sitator/util/mcl.py",
    "sitator/util/zeo.py",                     # imported by util/__init__.py:7 (needs only the ase stub)
    "sitator/util/PBCCalculator.pyx",
    "sitator/util/DotProdClassifier.pyx",
    "sitator/util/RecenterTrajectory.pyx",
    "sitator/landmark/__init__.py",
    "sitator/landmark/errors.py",
    "sitator/landmark/LandmarkAnalysis.py",
    "sitator/landmark/helpers.pyx",
    "sitator/landmark/cluster/__init__.py",
    "sitator/landmark/cluster/mcl.py",
    "sitator/landmark/cluster/dotprod.py",
    # package __init__ of sitator.dynamics pulls in merging/network/ase calculators;
    # ref_loader installs a namespace stub for the package and loads this one module.
    "sitator/dynamics/JumpAnalysis.py",
    "sitator/dynamics/RemoveUnoccupiedSites.py",
    "sitator/dynamics/SmoothSiteTrajectory.pyx",
    # site merging by Markov clustering of the jump statistics (pure Python; the package
    # __init__ of sitator.network pulls in ase calculators, so ref_loader installs a namespace stub for it too)
    "sitator/dynamics/MergeSitesByDynamics.py",
    "sitator/network/merging.py",
]


def reference_root():
    return os.environ.get("SITATOR_REFERENCE", "/root/reference")


def have_reference():
    return os.path.isfile(os.path.join(reference_root(), "sitator", "landmark", "helpers.pyx"))


def is_built():
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    for rel in MODULES:
        stem = os.path.splitext(rel)[0]
        if not os.path.isfile(os.path.join(OUT, stem + suffix)):
            return False
    return True


def build(force=False, verbose=True):
    """Compile the reference modules into oracle/_ref.  Returns True if built/present."""
    if is_built() and not force:
        return True
    if not have_reference():
        return False
    import numpy as np
    from Cython.Compiler import Options
    from Cython.Compiler.Main import compile as cython_compile, CompilationOptions

    Options.cimport_from_pyx = True  # as the reference's setup.py:7
    # LandmarkAnalysis.py:267 names InsufficientSitesError without importing it (a NameError at
    # run time in CPython); keep that behaviour instead of failing the compile.
    Options.error_on_unknown_names = False
    root = reference_root()
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    pyinc = sysconfig.get_paths()["include"]
    npinc = np.get_include()
    os.makedirs(BUILD, exist_ok=True)
    for rel in MODULES:
        src = os.path.join(root, rel)
        stem = os.path.splitext(rel)[0]
        c_file = os.path.join(BUILD, stem.replace("/", "__") + ".c")
        so_file = os.path.join(OUT, stem + suffix)
        os.makedirs(os.path.dirname(so_file), exist_ok=True)
        opts = CompilationOptions(
            language_level=3,
            include_path=[root],
            output_file=c_file,
            compiler_directives={"binding": True},
        )
        res = cython_compile(src, options=opts)
        if res.num_errors:
            raise RuntimeError("cython failed on %s" % rel)
        cmd = [
            "gcc", "-O2", "-fPIC", "-shared", "-fno-strict-aliasing", "-w",
            "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
            "-I", pyinc, "-I", npinc, c_file, "-o", so_file, "-lm",
        ]
        if verbose:
            print("[build_ref] %s -> %s" % (rel, os.path.relpath(so_file, HERE)))
        subprocess.check_call(cmd)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "built" if ok else "reference sources not found; nothing built")
    sys.exit(0 if ok else 1)
