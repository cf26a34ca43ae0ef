"""Lattice-replica containers and file formats used either side of the hot path.

The reference keeps its state in ``ase.Atoms`` (``rlsim/environment.py:22-53``) and
reads/writes VASP POSCAR / XDATCAR through ``ase.io`` (``rlsim/drl/simulator.py:175,215``).
ASE is not part of this build, so :class:`Atoms` is a small stand-in that exposes exactly
the accessors the hot path and its callers touch (``get_positions``, ``set_positions``,
``get_cell``, ``get_pbc``, ``get_atomic_numbers``, ``get_chemical_symbols``, ``copy``,
``__len__``).  A real ``ase.Atoms`` duck-types to the same interface, so everything in
this package accepts either.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import numpy as np

SYMBOL_TO_Z = {"Cr": 24, "Co": 27, "Ni": 28, "Cu": 29, "Au": 79, "Fe": 26, "Al": 13}
Z_TO_SYMBOL = {z: s for s, z in SYMBOL_TO_Z.items()}


class Atoms:
    """Minimal ``ase.Atoms`` look-alike: species, Cartesian positions (Å), row-vector cell."""

    def __init__(self, numbers: Sequence[int], positions, cell, pbc=(True, True, True)):
        self.numbers = np.asarray(numbers, dtype=np.int64).copy()
        self.positions = np.asarray(positions, dtype=np.float64).reshape(-1, 3).copy()
        self.cell = np.asarray(cell, dtype=np.float64).reshape(3, 3).copy()
        self.pbc = np.asarray(pbc, dtype=bool).copy()
        if len(self.numbers) != len(self.positions):
            raise ValueError("numbers and positions differ in length")

    def __len__(self) -> int:
        return len(self.numbers)

    def copy(self) -> "Atoms":
        return Atoms(self.numbers, self.positions, self.cell, self.pbc)

    def get_positions(self) -> np.ndarray:
        return self.positions.copy()

    def set_positions(self, positions) -> None:
        positions = np.asarray(positions, dtype=np.float64).reshape(-1, 3)
        if positions.shape != self.positions.shape:
            raise ValueError("positions have the wrong shape")
        self.positions = positions.copy()

    def get_cell(self) -> np.ndarray:
        return self.cell.copy()

    def get_pbc(self) -> np.ndarray:
        return self.pbc.copy()

    def get_atomic_numbers(self) -> np.ndarray:
        return self.numbers.copy()

    def get_chemical_symbols(self) -> List[str]:
        return [Z_TO_SYMBOL.get(int(z), f"Z{int(z)}") for z in self.numbers]

    def get_scaled_positions(self, wrap: bool = True) -> np.ndarray:
        frac = np.linalg.solve(self.cell.T, self.positions.T).T
        if wrap:
            frac = frac - np.floor(frac)
        return frac

    def wrap(self) -> None:
        """Fold every atom back into the unit cell (``Environment.normalize_positions``)."""
        self.positions = self.get_scaled_positions(wrap=True) @ self.cell

    def get_volume(self) -> float:
        return float(abs(np.linalg.det(self.cell)))


def cell_heights(cell) -> np.ndarray:
    """Perpendicular widths of a row-vector cell (distance between opposite faces)."""
    cell = np.asarray(cell, dtype=np.float64)
    vol = abs(np.linalg.det(cell))
    cross = np.stack([np.cross(cell[1], cell[2]), np.cross(cell[2], cell[0]), np.cross(cell[0], cell[1])])
    return vol / np.linalg.norm(cross, axis=1)


# ----------------------------------------------------------------------------- POSCAR / XDATCAR

def read_poscar(path: str) -> Atoms:
    """VASP5 POSCAR reader (Cartesian or Direct, optional Selective dynamics line)."""
    with open(path) as fh:
        lines = [ln.rstrip("\n") for ln in fh]
    scale = float(lines[1].split()[0])
    cell = np.array([[float(x) for x in lines[i].split()[:3]] for i in (2, 3, 4)])
    if scale < 0:  # negative scale means "target volume"
        scale = (-scale / abs(np.linalg.det(cell))) ** (1.0 / 3.0)
    cell = cell * scale
    symbols = lines[5].split()
    try:  # VASP4: counts directly on line 5
        counts = [int(x) for x in symbols]
        symbols = lines[0].split()[: len(counts)]
        cursor = 6
    except ValueError:
        counts = [int(x) for x in lines[6].split()]
        cursor = 7
    if lines[cursor].strip()[:1].lower() == "s":
        cursor += 1
    mode = lines[cursor].strip()[:1].lower()
    cursor += 1
    n_atoms = sum(counts)
    coords = np.array([[float(x) for x in lines[cursor + i].split()[:3]] for i in range(n_atoms)])
    if mode in ("c", "k"):
        positions = coords * scale
    else:
        positions = coords @ cell
    numbers: List[int] = []
    for sym, cnt in zip(symbols, counts):
        numbers += [SYMBOL_TO_Z[sym]] * cnt
    return Atoms(numbers, positions, cell)


def _species_blocks(atoms) -> tuple[list[str], list[int], np.ndarray]:
    symbols = list(atoms.get_chemical_symbols())
    order, counts, index = [], [], []
    for sym in symbols:
        if sym not in order:
            order.append(sym)
    for sym in order:
        ids = [i for i, s in enumerate(symbols) if s == sym]
        counts.append(len(ids))
        index += ids
    return order, counts, np.asarray(index, dtype=np.int64)


def write_poscar(path: str, atoms, direct: bool = False) -> None:
    order, counts, index = _species_blocks(atoms)
    cell = np.asarray(atoms.get_cell(), dtype=np.float64)
    pos = np.asarray(atoms.get_positions(), dtype=np.float64)[index]
    with open(path, "w") as fh:
        fh.write(" ".join(order) + "\n 1.0000000000000000\n")
        for row in cell:
            fh.write("  " + " ".join(f"{x:21.16f}" for x in row) + "\n")
        fh.write(" " + " ".join(f"{s:>3s}" for s in order) + "\n")
        fh.write(" " + " ".join(f"{c:3d}" for c in counts) + "\n")
        if direct:
            fh.write("Direct\n")
            pos = np.linalg.solve(cell.T, pos.T).T
        else:
            fh.write("Cartesian\n")
        for row in pos:
            fh.write(" " + " ".join(f"{x:19.16f}" for x in row) + "\n")


def append_xdatcar(path: str, atoms, frame: int, append: bool = True) -> None:
    """One XDATCAR frame (``io.write(..., format='vasp-xdatcar', append=True)``,
    ``rlsim/drl/simulator.py:175,215``)."""
    order, counts, index = _species_blocks(atoms)
    cell = np.asarray(atoms.get_cell(), dtype=np.float64)
    frac = np.linalg.solve(cell.T, np.asarray(atoms.get_positions(), dtype=np.float64)[index].T).T
    with open(path, "a" if append else "w") as fh:
        if not append or frame == 0:
            fh.write(" ".join(order) + "\n 1.0\n")
            for row in cell:
                fh.write("  " + " ".join(f"{x:12.6f}" for x in row) + "\n")
            fh.write(" " + " ".join(f"{s:>3s}" for s in order) + "\n")
            fh.write(" " + " ".join(f"{c:3d}" for c in counts) + "\n")
        fh.write(f"Direct configuration={frame + 1:6d}\n")
        for row in frac:
            fh.write(" " + " ".join(f"{x:11.8f}" for x in row) + "\n")


# ----------------------------------------------------------------------------- synthetic CrCoNi

_FCC_BASIS = np.array([[0.0, 0.0, 0.0], [0.0, 0.5, 0.5], [0.5, 0.0, 0.5], [0.5, 0.5, 0.0]])


def fcc_sites(n_rep: Sequence[int] | int, a0: float) -> np.ndarray:
    """Perfect conventional-cubic FCC sites in the order ``ase.build.bulk('Cu','fcc',a,cubic=True)
    .repeat(n)`` produces them (outer loop over the first repeat axis, 4-atom basis innermost);
    the live enumerator indexes sites in this order (``rlsim/actions/action.py:22-23``)."""
    if isinstance(n_rep, (int, np.integer)):
        n_rep = (int(n_rep),) * 3
    n0, n1, n2 = (int(x) for x in n_rep)
    m = np.stack(np.meshgrid(np.arange(n0), np.arange(n1), np.arange(n2), indexing="ij"), axis=-1).reshape(-1, 3)
    sites = (m[:, None, :] + _FCC_BASIS[None, :, :]).reshape(-1, 3) * a0
    return sites


def make_crconi_supercell(n_rep: int = 4, a0: float = 3.528, n_vac: int = 1, seed: int = 0,
                          jitter: float = 0.0, non_adjacent: bool = True) -> Atoms:
    """Synthetic equiatomic CrCoNi replica.

    ``4*n_rep**3`` FCC sites; species are a seeded permutation of near-equal Cr/Co/Ni counts;
    ``n_vac`` seeded sites are deleted (mutually non-adjacent so every vacancy keeps 12 candidate
    hops); optional Gaussian jitter (Å) exercises non-degenerate distances.
    """
    rng = np.random.default_rng(seed)
    sites = fcc_sites(n_rep, a0)
    n_sites = len(sites)
    base = n_sites // 3
    counts = [base, base, n_sites - 2 * base]
    numbers = np.concatenate([np.full(c, z) for c, z in zip(counts, (24, 27, 28))])
    numbers = numbers[rng.permutation(n_sites)]
    box = n_rep * a0
    vac: List[int] = []
    nn = a0 / np.sqrt(2.0) * 1.1
    while len(vac) < n_vac:
        cand = int(rng.integers(n_sites))
        if cand in vac:
            continue
        if non_adjacent and vac:
            d = sites[vac] - sites[cand]
            d -= box * np.round(d / box)
            if np.any(np.linalg.norm(d, axis=1) < nn):
                continue
        vac.append(cand)
    keep = np.setdiff1d(np.arange(n_sites), np.asarray(vac, dtype=np.int64))
    pos = sites[keep]
    if jitter > 0.0:
        pos = pos + rng.normal(0.0, jitter, size=pos.shape)
    return Atoms(numbers[keep], pos, np.eye(3) * box)


def apply_action(atoms, action: Iterable[float]):
    """Product state of one hop: ``pos[act[0]] += act[1:]`` (``rlsim/drl/simulator.py:446-451``)."""
    out = atoms.copy()
    pos = np.asarray(out.get_positions(), dtype=np.float64)
    action = list(action)
    pos[int(action[0])] += np.asarray(action[1:4], dtype=np.float64)
    out.set_positions(pos)
    return out
