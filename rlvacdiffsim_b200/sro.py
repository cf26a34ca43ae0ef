"""Warren–Cowley short-range order.

:func:`get_sro_from_atoms` mirrors ``rlsim/utils/sro.py:49-83`` in numpy (no ASE): all minimum-image pair
distances, first-shell cutoff ``(1 + sqrt 2) / 2 * NN_dist``, pair counts per species pair,
``SRO[a, b] = 1 - (n_ab / n_pairs) / c_a / c_b``.  :func:`sro_device` gets the integer pair counts of many structures
from one kernel launch (``rlvd_sro_counts``) and applies the same fp64 formula on the host, so both agree exactly
whenever the pair sets agree.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from .actions import _atoms_of, mic_vectors


def _sro_from_counts(pair_counts: np.ndarray, species_counts: np.ndarray) -> np.ndarray:
    """``pair_counts[a, b]`` = number of ordered first-shell pairs (i of species a, j of species b)."""
    ns = len(species_counts)
    total_pairs = pair_counts.sum() / 2
    conc = species_counts / species_counts.sum()
    sro = np.zeros((ns, ns))
    for n1 in range(ns):
        for n2 in range(n1 + 1):
            number_of_pairs = pair_counts[n2, n1] / 2                       # sro.py:70-73
            sro[n1, n2] = 1 - (number_of_pairs / total_pairs) / conc[n1] / conc[n2]
            sro[n2, n1] = sro[n1, n2]
    return sro


def get_sro_from_atoms(atoms) -> np.ndarray:
    atoms = _atoms_of(atoms)
    Z = np.asarray(atoms.get_atomic_numbers())
    species = sorted(set(Z.tolist()))
    pos = np.asarray(atoms.get_positions(), dtype=np.float64)
    dist_all = np.linalg.norm(mic_vectors(pos, pos, np.asarray(atoms.get_cell(), dtype=np.float64)), axis=2)
    nn = dist_all[dist_all > 0.1].min()                                     # sro.py:56-58
    r_cut = (1 + np.sqrt(2)) / 2 * nn
    pairs = (dist_all > 0.1) & (dist_all < r_cut)
    idx = np.searchsorted(species, Z)
    counts = np.zeros((len(species), len(species)))
    np.add.at(counts, (idx[:, None].repeat(len(Z), 1)[pairs], idx[None, :].repeat(len(Z), 0)[pairs]), 1)
    return _sro_from_counts(counts, np.bincount(idx, minlength=len(species)).astype(np.float64))


def sro_device(engine, structures: Sequence, species: Sequence[int]) -> List[np.ndarray]:
    """SRO matrices of many structures with orthorhombic cells from one launch; ``species`` = sorted atomic numbers."""
    import torch

    atoms_l = [_atoms_of(s) for s in structures]
    node_ptr = np.zeros(len(atoms_l) + 1, dtype=np.int32)
    np.cumsum([len(a) for a in atoms_l], out=node_ptr[1:])
    pos = np.concatenate([np.asarray(a.get_positions(), dtype=np.float64) for a in atoms_l])
    cells = np.stack([np.asarray(a.get_cell(), dtype=np.float64).reshape(3, 3) for a in atoms_l])
    species = sorted(int(z) for z in species)
    sid = np.concatenate([np.searchsorted(species, np.asarray(a.get_atomic_numbers())) for a in atoms_l]).astype(np.int32)
    dev = engine.device
    counts, status = engine.sro_counts(torch.from_numpy(pos).to(dev), torch.from_numpy(cells).to(dev),
                                       torch.from_numpy(node_ptr).to(dev), torch.from_numpy(sid).to(dev), len(species))
    counts, status = counts.cpu().numpy(), status.cpu().numpy()
    out = []
    for g, atoms in enumerate(atoms_l):
        if status[g] != 0:                                                   # non-orthorhombic cell: host path
            out.append(get_sro_from_atoms(atoms))
            continue
        spc = np.bincount(sid[node_ptr[g]:node_ptr[g + 1]], minlength=len(species)).astype(np.float64)
        out.append(_sro_from_counts(counts[g].astype(np.float64), spc))
    return out
