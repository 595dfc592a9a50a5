"""Per-phase timeline of an analysis: host wall-clock per named phase + an NVTX range around each.

``LandmarkAnalysis.run`` records one entry per phase into ``la.stats['phases_ms']``.  By default nothing is
synchronised (the numbers are the host time spent issuing each phase and waiting where the phase itself waits);
with ``SITB_PHASE_SYNC=1`` the device is synchronised at the end of every phase, so that each entry is that
phase's full cost (what ``profiles/r02_e2e_*_phases.json`` hold).  NVTX ranges make the same phases visible to
Nsight tools.  No reference counterpart (the reference has tqdm bars only).
"""
import contextlib
import os
import time


class PhaseTimer(object):
    def __init__(self, device=None):
        self.ms = {}
        self.order = []
        self.sync = os.environ.get("SITB_PHASE_SYNC", "0") not in ("0", "", "false")
        self.device = device
        try:
            import torch
            self._nvtx = torch.cuda.nvtx if torch.cuda.is_available() else None
            self._torch = torch
        except Exception:                       # pragma: no cover
            self._nvtx = None
            self._torch = None

    @contextlib.contextmanager
    def phase(self, name):
        if self._nvtx is not None:
            self._nvtx.range_push("sitator_b200:" + name)
        t0 = time.perf_counter()
        try:
            yield
        finally:
            if self.sync and self._torch is not None:
                self._torch.cuda.synchronize(self.device)
            dt = (time.perf_counter() - t0) * 1e3
            if name not in self.ms:
                self.order.append(name)
                self.ms[name] = 0.0
            self.ms[name] += dt
            if self._nvtx is not None:
                self._nvtx.range_pop()

    def as_dict(self):
        return {k: self.ms[k] for k in self.order}


_NULL = None


def null_timer():
    """A timer that only keeps the dictionary (used when a plugin is called outside ``run``)."""
    global _NULL
    if _NULL is None:
        _NULL = PhaseTimer()
    return PhaseTimer()
