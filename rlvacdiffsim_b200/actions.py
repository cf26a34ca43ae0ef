"""Candidate vacancy-hop enumeration.

Two enumerators, both host-side numpy like the reference's:

* :func:`get_action_space` follows the live ``rlsim/actions/action.py:12-61``: map atoms to
  perfect FCC sites by minimum-image argmin, the unoccupied sites are the vacancies, every
  occupied nearest-neighbour site of a vacancy yields ``[atom_idx, dx, dy, dz]`` with the
  minimum-image displacement to the vacant site.  The reference asserts exactly one vacancy
  (``action.py:35``) and a cubic cell (``action.py:20``); here ``max_vacancies`` lifts the first
  restriction (BASELINE config 3) and non-cubic cells get their sites from the same FCC lattice
  clipped to the cell instead of ``cell.diagonal()``.  On a cubic single-vacancy cell the site
  order, the action order and the numbers are the reference's.
* :func:`get_action_space_legacy` follows the commented-out ``action.py:162-198`` ("Previous
  version that worked perfectly for CrCoNi"): atoms whose nearest-neighbour count differs from
  13 try the twelve <110>*a/2*0.8 hops and keep those that land on an empty spot.  The bundled
  checkpoint was trained under this 0.8-scaled convention (``simulator.py:66,70``).

Either accepts an ``Environment``-like object (anything with ``.atoms``) or the atoms themselves.
"""
from __future__ import annotations

from typing import List

import numpy as np

from .structures import fcc_sites

_LEGACY_DIRS = np.array([[1, 1, 0], [1, -1, 0], [-1, 1, 0], [-1, -1, 0],
                         [1, 0, 1], [1, 0, -1], [-1, 0, 1], [-1, 0, -1],
                         [0, 1, 1], [0, 1, -1], [0, -1, 1], [0, -1, -1]], dtype=np.float64)


def _atoms_of(config):
    return config.atoms if hasattr(config, "atoms") else config


def mic_vectors(p1: np.ndarray, p2: np.ndarray, cell: np.ndarray) -> np.ndarray:
    """Minimum-image vectors p2[j]-p1[i] → [n1,n2,3] (``ase.geometry.get_distances`` semantics).

    Fractional rounding followed by a scan of the 27 surrounding images, which is exact for any
    cell whose shortest lattice vector is not much shorter than its heights (all cells here)."""
    cell = np.asarray(cell, dtype=np.float64)
    d = p2[None, :, :] - p1[:, None, :]
    frac = np.linalg.solve(cell.T, d.reshape(-1, 3).T).T
    frac -= np.round(frac)
    base = (frac @ cell).reshape(d.shape)
    if np.allclose(cell, np.diag(np.diag(cell))):
        return base
    best = base.copy()
    best_n = np.einsum("ijk,ijk->ij", base, base)
    for s in np.ndindex(3, 3, 3):
        shift = (np.array(s) - 1.0) @ cell
        if not shift.any():
            continue
        cand = base + shift
        n = np.einsum("ijk,ijk->ij", cand, cand)
        better = n < best_n - 1e-12
        best[better] = cand[better]
        best_n[better] = n[better]
    return best


def lattice_sites_for_cell(cell: np.ndarray, lattice_parameter: float) -> np.ndarray:
    """Perfect FCC sites filling ``cell``.  Cubic/orthorhombic cells: the reference's
    ``bulk(...).repeat(round(diag/a))`` order.  Other cells: the same lattice clipped to
    fractional coordinates in [0,1), ordered by conventional-cell index then basis."""
    cell = np.asarray(cell, dtype=np.float64)
    if np.allclose(cell, np.diag(np.diag(cell))):
        n_rep = np.round(cell.diagonal() / lattice_parameter).astype(int)
        return fcc_sites(n_rep, lattice_parameter)
    corners = np.array([[i, j, k] for i in (0, 1) for j in (0, 1) for k in (0, 1)], dtype=np.float64) @ cell
    lo = np.floor(corners.min(0) / lattice_parameter).astype(int) - 1
    hi = np.ceil(corners.max(0) / lattice_parameter).astype(int) + 1
    sites = fcc_sites(hi - lo, lattice_parameter) + lo * lattice_parameter
    frac = np.linalg.solve(cell.T, sites.T).T
    frac = np.where(np.abs(frac - np.round(frac)) < 1e-9, np.round(frac), frac)
    inside = np.all((frac >= 0.0) & (frac < 1.0), axis=1)
    return sites[inside]


def get_action_space(config, lattice_parameter: float = 3.615, max_vacancies: int = 1) -> List[List[float]]:
    atoms = _atoms_of(config)
    cell = np.array(atoms.get_cell(), dtype=np.float64)
    sites = lattice_sites_for_cell(cell, lattice_parameter)
    atom_pos = np.asarray(atoms.get_positions(), dtype=np.float64)

    dists = np.linalg.norm(mic_vectors(atom_pos, sites, cell), axis=2)        # action.py:27-28
    atom_to_site = np.argmin(dists, axis=1)                                    # action.py:30
    occupied = set(atom_to_site.tolist())
    vacant = sorted(set(range(len(sites))) - occupied)                         # action.py:33-34
    if max_vacancies == 1:
        assert len(vacant) == 1, f"Expected 1 vacancy, found {len(vacant)}"    # action.py:35
    elif not 1 <= len(vacant) <= max_vacancies:
        raise AssertionError(f"Expected 1..{max_vacancies} vacancies, found {len(vacant)}")
    site_to_atom = {int(site): int(atom) for atom, site in enumerate(atom_to_site)}

    nn_dist = lattice_parameter / np.sqrt(2.0)
    actions: List[List[float]] = []
    for vac_site in vacant:
        vac_pos = sites[vac_site]
        vac_to_sites = np.linalg.norm(mic_vectors(vac_pos[None, :], sites, cell), axis=2)[0]
        nn_sites = np.where((vac_to_sites > 0.1) & (vac_to_sites < nn_dist * 1.1))[0]   # action.py:41-44
        for site in nn_sites:
            if int(site) not in site_to_atom:   # neighbouring vacancy: nothing to move
                continue
            atom_idx = site_to_atom[int(site)]
            diff = vac_pos - atom_pos[atom_idx]                                # action.py:53-57
            frac = np.linalg.solve(cell.T, diff)
            frac -= np.round(frac)
            disp = cell.T @ frac
            actions.append([atom_idx] + disp.tolist())
    return actions


def get_action_space_legacy(config, lattice_parameter: float = 3.528) -> List[List[float]]:
    atoms = _atoms_of(config)
    a = lattice_parameter
    cell = np.array(atoms.get_cell(), dtype=np.float64)
    pos = np.asarray(atoms.get_positions(), dtype=np.float64)
    dist_mat = np.linalg.norm(mic_vectors(pos, pos, cell), axis=2)             # action.py:165
    crit = np.sum(dist_mat < a / np.sqrt(2.0) * 1.2, axis=1)                   # action.py:167
    vacancy_l = np.argwhere(crit != 13).T[0]                                   # action.py:168
    acts = _LEGACY_DIRS * a / 2 * 0.8                                          # action.py:178-189
    actions: List[List[float]] = []
    for i in vacancy_l:
        for vec in acts:
            trial = pos.copy()
            trial[i] += vec
            d = np.linalg.norm(mic_vectors(trial[i][None, :], trial, cell), axis=2)[0]
            if np.sum(d < 0.8) == 1:                                           # action.py:176
                actions.append([int(i)] + vec.tolist())
    return actions


def get_action_space_device(engine, replicas, lattice_parameter: float = 3.615, max_vacancies: int = 1) -> List[List[List[float]]]:
    """:func:`get_action_space` for many replicas in ONE device launch (``rlvd_action_space``).

    Same rows in the same order as the host enumerator; orthorhombic cells take ``rlvd_action_space``, every other cell
    ``rlvd_action_space_general`` (only a cell with more than 8192 lattice sites falls through to the host enumerator),
    and the reference's vacancy-count assert
    (``action.py:35``) is raised for status 2.  ``engine`` is an :class:`rlvacdiffsim_b200.engine.Engine`."""
    import torch

    atoms_l = [_atoms_of(r) for r in replicas]
    counts = np.array([len(a) for a in atoms_l], dtype=np.int64)
    node_ptr = np.zeros(len(atoms_l) + 1, dtype=np.int32)
    np.cumsum(counts, out=node_ptr[1:])
    pos = np.concatenate([np.asarray(a.get_positions(), dtype=np.float64) for a in atoms_l])
    cells = np.stack([np.array(a.get_cell(), dtype=np.float64).reshape(3, 3) for a in atoms_l])
    dev = engine.device
    cnt, act, status = engine.action_space(torch.from_numpy(pos).to(dev), torch.from_numpy(cells).to(dev),
                                           torch.from_numpy(node_ptr).to(dev), lattice_parameter, max_vacancies,
                                           max_actions=12 * max_vacancies)
    cnt, act, status = cnt.cpu().numpy(), act.cpu().numpy(), status.cpu().numpy()
    general = [g for g in range(len(atoms_l)) if status[g] == 1]
    if general:
        # non-orthorhombic cells (e.g. the reference's rhombohedral demo POSCAR): the general minimum-image kernel with the
        # cells' perfect-lattice sites computed once on the host
        site_list = [lattice_sites_for_cell(cells[g], lattice_parameter) for g in general]
        sp = np.zeros(len(general) + 1, dtype=np.int32)
        np.cumsum([len(x) for x in site_list], out=sp[1:])
        np2 = np.zeros(len(general) + 1, dtype=np.int32)
        np.cumsum([counts[g] for g in general], out=np2[1:])
        pos_g = np.concatenate([pos[node_ptr[g]:node_ptr[g + 1]] for g in general])
        c2, a2, s2 = engine.action_space_general(
            torch.from_numpy(pos_g).to(dev), torch.from_numpy(cells[general]).to(dev),
            torch.from_numpy(np.linalg.inv(cells[general])).to(dev), torch.from_numpy(np2).to(dev), torch.from_numpy(sp).to(dev),
            torch.from_numpy(np.concatenate(site_list) if site_list else np.zeros((0, 3))).to(dev), lattice_parameter, max_vacancies,
            max_actions=12 * max_vacancies)
        c2, a2, s2 = c2.cpu().numpy(), a2.cpu().numpy(), s2.cpu().numpy()
        for k, g in enumerate(general):
            cnt[g], act[g], status[g] = c2[k], a2[k], s2[k]
    out: List[List[List[float]]] = []
    for g, atoms in enumerate(atoms_l):
        if status[g] == 1:
            out.append(get_action_space(atoms, lattice_parameter, max_vacancies))
            continue
        if status[g] == 2:
            raise AssertionError("Expected 1 vacancy" if max_vacancies == 1 else f"Expected 1..{max_vacancies} vacancies")
        if status[g] != 0:
            raise RuntimeError(f"rlvd_action_space: status {int(status[g])} for replica {g}")
        out.append([[int(r[0])] + r[1:].tolist() for r in act[g, : cnt[g]]])
    return out
