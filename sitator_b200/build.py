"""In-tree build of the CUDA library (sm_100a only): ``python -m sitator_b200.build``.

nvcc cross-compiles without a GPU; the resulting ``sitator_b200/lib/libsitator_b200.so`` is
git-ignored but travels to the GPU box with the repository snapshot.
"""
import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "lib", "libsitator_b200.so")
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(PKG, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force=False, verbose=False, extra_flags=()):
    """Compile csrc/*.cu into lib/libsitator_b200.so.  Returns the library path."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    extra_flags = list(extra_flags) + os.environ.get("SITB_NVCC_FLAGS", "").split()      # developer experiments (-DSITB_...)
    cmd = [nvcc] + NVCC_FLAGS + extra_flags + ["-o", LIB] + sources()
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build_native(force="--force" in sys.argv, verbose=True,
                 extra_flags=["-Xptxas", "-v"] if "--ptxas" in sys.argv else [])
    print(LIB)
