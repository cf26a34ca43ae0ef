"""Minimal stand-in for the parts of ``ase`` that rlsim/actions/action.py and rlsim/utils/sro.py touch.

TEST INFRASTRUCTURE (see tests/ref_stub/README.md): lets the reference's own enumerators run unmodified in an image
without ASE.  Not a general ASE replacement."""
from __future__ import annotations

import numpy as np

from .geometry import get_distances as _get_distances

_SYMBOLS = {24: "Cr", 27: "Co", 28: "Ni", 29: "Cu", 79: "Au", 26: "Fe", 13: "Al"}
_NUMBERS = {s: z for z, s in _SYMBOLS.items()}


class Atoms:
    def __init__(self, symbols=None, positions=None, numbers=None, cell=None, pbc=None):
        if numbers is None:
            if isinstance(symbols, str):
                raise ValueError("the stub takes a list of symbols or numbers")
            numbers = [_NUMBERS[s] for s in symbols]
        self.arrays = {"numbers": np.asarray(numbers, dtype=int).copy(),
                       "positions": np.asarray(positions, dtype=float).reshape(-1, 3).copy()}
        self.cell = np.zeros((3, 3)) if cell is None else np.asarray(cell, dtype=float).reshape(3, 3).copy()
        self.pbc = np.zeros(3, bool) if pbc is None else np.broadcast_to(np.asarray(pbc, dtype=bool), (3,)).copy()

    def __len__(self):
        return len(self.arrays["numbers"])

    def copy(self):
        return Atoms(numbers=self.arrays["numbers"], positions=self.arrays["positions"], cell=self.cell, pbc=self.pbc)

    def get_positions(self):
        return self.arrays["positions"].copy()

    def set_positions(self, pos):
        self.arrays["positions"][:] = np.asarray(pos, dtype=float).reshape(-1, 3)

    positions = property(get_positions, set_positions)

    def get_cell(self):
        return self.cell.copy()

    def get_pbc(self):
        return self.pbc.copy()

    def get_atomic_numbers(self):
        return self.arrays["numbers"].copy()

    @property
    def numbers(self):
        return self.get_atomic_numbers()

    def get_chemical_symbols(self):
        return [_SYMBOLS[int(z)] for z in self.arrays["numbers"]]

    def get_all_distances(self, mic=False, vector=False):
        R = self.arrays["positions"]
        D, D_len = _get_distances(R, cell=self.cell if mic else None, pbc=self.pbc if mic else None)
        return D if vector else D_len

    def get_distances(self, a, indices, mic=False, vector=False):
        R = self.arrays["positions"]
        D, D_len = _get_distances([R[a]], R[list(indices)], cell=self.cell if mic else None,
                                  pbc=self.pbc if mic else None)
        return D[0] if vector else D_len.reshape(-1)

    def repeat(self, rep):
        if isinstance(rep, (int, np.integer)):
            rep = (rep,) * 3
        m = [int(x) for x in rep]
        n = len(self)
        pos = np.tile(self.arrays["positions"], (m[0] * m[1] * m[2], 1))
        i0 = 0
        for m0 in range(m[0]):
            for m1 in range(m[1]):
                for m2 in range(m[2]):
                    pos[i0:i0 + n] += np.dot((m0, m1, m2), self.cell)
                    i0 += n
        return Atoms(numbers=np.tile(self.arrays["numbers"], m[0] * m[1] * m[2]), positions=pos,
                     cell=np.array([m[c] * self.cell[c] for c in range(3)]), pbc=self.pbc)
