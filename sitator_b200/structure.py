"""A minimal stand-in for ``ase.Atoms`` so the package runs where ASE is absent.

``SiteNetwork`` only needs ``positions``, ``cell``, ``numbers``, ``len``, ``copy`` and
mask deletion from its structure (reference ``SiteNetwork.py:58-67``); a real
``ase.Atoms`` satisfies the same duck type and is accepted everywhere.
"""
import numpy as np


class Atoms(object):
    def __init__(self, positions, cell, numbers=None, pbc=True):
        self.positions = np.array(positions, dtype=np.float64).reshape(-1, 3)
        self.cell = np.array(cell, dtype=np.float64).reshape(3, 3)
        n = len(self.positions)
        self.numbers = (np.zeros(n, dtype=np.int64) if numbers is None
                        else np.array(numbers, dtype=np.int64))
        if len(self.numbers) != n:
            raise ValueError("numbers and positions differ in length")
        self.pbc = pbc

    def __len__(self):
        return len(self.positions)

    def copy(self):
        return Atoms(self.positions.copy(), self.cell.copy(), self.numbers.copy(), self.pbc)

    def __delitem__(self, key):
        key = np.asarray(key)
        if key.dtype == bool:
            keep = ~key
        else:
            keep = np.ones(len(self), dtype=bool)
            keep[key] = False
        self.positions = self.positions[keep]
        self.numbers = self.numbers[keep]

    def get_positions(self):
        return self.positions.copy()

    def get_atomic_numbers(self):
        return self.numbers.copy()

    def get_cell(self):
        return self.cell.copy()
