"""Recenter a trajectory on the centre of mass of the static atoms (reference
``sitator/util/RecenterTrajectory.pyx:9-100``), IN PLACE like the reference.  The per-frame pass is
``sitb_recenter`` (``csrc/sitb_post.cu``); host arrays are streamed through the device in chunks."""
import ctypes as C

import numpy as np

from .. import _native


class RecenterTrajectory(object):

    def run(self, structure, static_mask, positions, velocities=None, masses=None):
        """Recenters ``positions`` (and ``velocities``) on the centre of mass of the atoms indicated by
        ``static_mask``, in place, and brings the positions to the cell centre.

        Args:
            structure: the simulation's structure (``ase.Atoms`` or anything with ``cell``, ``get_atomic_numbers()``
                and, if ``masses`` is None, ``get_masses()``).
            static_mask (ndarray): boolean mask of the atoms to recenter on.
            positions (ndarray): (n_frames, n_atoms, 3) float64, modified in place.
            velocities (ndarray, optional): same; modified in place if provided.
            masses (None, dict or ndarray): None: ``structure.get_masses()``; dict: chemical symbol -> mass
                (needs ``structure.get_chemical_symbols()``); ndarray: one mass per atom.
        """
        static_mask = np.asarray(static_mask)
        assert np.any(static_mask), "Static mask all false; there must be static atoms to recenter on."
        factors = static_mask.astype(np.float64)
        if masses is None:
            mass_arr = np.asarray(structure.get_masses(), dtype=np.float64)
        elif isinstance(masses, dict):
            symbols = structure.get_chemical_symbols()
            mass_arr = np.array([masses[s] for s in symbols], dtype=np.float64)
        elif isinstance(masses, np.ndarray):
            mass_arr = np.asarray(masses, dtype=np.float64)
        else:
            raise TypeError("Don't know how to interpret masses `%s`; must be None, dict, or ndarray" % masses)
        cell = np.asarray(structure.cell, dtype=np.float64).reshape(3, 3)
        centroid = np.sum(0.5 * cell, axis=0)                               # PBCCalculator.pyx:35
        _recenter_array(positions, mass_arr, factors, centroid)
        if velocities is not None:
            _recenter_array(velocities, mass_arr, factors, None)
        return None


def _recenter_array(array, masses, factors, shift):
    """recenter_traj_array (RecenterTrajectory.pyx:61-100) + the centroid shift of :56-57, in place."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("sitator_b200 needs a CUDA device; there is no CPU path")
    if array.dtype != np.float64 or array.ndim != 3 or array.shape[2] != 3:
        raise ValueError("array must be (n_frames, n_atoms, 3) float64")
    n_frames, n_atoms = array.shape[0], array.shape[1]
    assert len(masses) == n_atoms and len(factors) == n_atoms
    total = 0.0
    for j in range(n_atoms):                                                # :80-83, sequential like the reference
        total += factors[j] * masses[j]
    tmi = 1.0 / total
    weights = (tmi * factors) * masses                                      # :93: tmi * factors[j] * masses[j], left to right
    lib = _native.load()
    dev = torch.cuda.current_device()
    d_w = torch.as_tensor(np.ascontiguousarray(weights), device="cuda")
    shift_arr = None if shift is None else np.ascontiguousarray(shift, dtype=np.float64)
    stream = torch.cuda.current_stream().cuda_stream
    step = max(1, (256 << 20) // (n_atoms * 24))
    for f0 in range(0, n_frames, step):
        n = min(step, n_frames - f0)
        chunk = torch.as_tensor(np.ascontiguousarray(array[f0:f0 + n]), device="cuda")
        _native.check(lib.sitb_recenter(dev, C.c_void_p(chunk.data_ptr()), n, n_atoms, C.c_void_p(d_w.data_ptr()),
                                        None if shift_arr is None else shift_arr.ctypes.data, C.c_void_p(stream)))
        array[f0:f0 + n] = chunk.cpu().numpy()
