"""``ase.geometry.get_distances`` / ``find_mic`` semantics (minimum-image convention), numpy only."""
import itertools

import numpy as np


def _reduce_cell(cell):
    """Reduced basis of the same lattice: greedy pair and triple reduction until no shorter vector is found
    (what ASE obtains from its Minkowski reduction; for such a basis the minimum image lies among the 27 neighbours)."""
    b = np.array(cell, dtype=float)
    for _ in range(100):
        changed = False
        b = b[np.argsort(np.linalg.norm(b, axis=1), kind="stable")]
        for i in (1, 2):
            for j in range(i):
                x = np.round(b[i] @ b[j] / (b[j] @ b[j]))
                if x != 0:
                    b[i] = b[i] - x * b[j]
                    changed = True
        best, best_n = None, b[2] @ b[2]
        for x1, x2 in itertools.product((-1, 0, 1), repeat=2):
            cand = b[2] + x1 * b[0] + x2 * b[1]
            if cand @ cand < best_n - 1e-12:
                best, best_n = cand, cand @ cand
        if best is not None:
            b[2] = best
            changed = True
        if not changed:
            break
    return b


def naive_find_mic(v, cell):
    f = np.linalg.solve(cell.T, v.T).T
    f -= np.floor(f + 0.5)
    vmin = f @ cell
    return vmin, np.linalg.norm(vmin, axis=1)


def general_find_mic(v, cell):
    rcell = _reduce_cell(cell)
    f = np.linalg.solve(rcell.T, v.T).T
    f -= np.floor(f)                                   # wrap into the reduced cell
    positions = f @ rcell
    hkls = np.array([(0, 0, 0)] + list(itertools.product((-1, 0, 1), repeat=3)))
    x = positions + (hkls @ rcell)[:, None]
    lengths = np.linalg.norm(x, axis=2)
    idx = np.argmin(lengths, axis=0)
    ar = np.arange(len(positions))
    return x[idx, ar, :], lengths[idx, ar]


def find_mic(v, cell, pbc=True):
    cell = np.asarray(cell, dtype=float).reshape(3, 3)
    pbc = np.broadcast_to(np.asarray(pbc, dtype=bool), (3,))
    v = np.atleast_2d(np.asarray(v, dtype=float))
    if not pbc.all():
        raise NotImplementedError("the stub handles fully periodic cells only")
    vmin, vlen = naive_find_mic(v, cell)
    if not (vlen < 0.5 * np.linalg.norm(cell, axis=1).min()).all():
        vmin, vlen = general_find_mic(v, cell)
    return vmin, vlen


def get_distances(p1, p2=None, cell=None, pbc=None):
    p1 = np.atleast_2d(np.asarray(p1, dtype=float))
    if p2 is None:
        n = len(p1)
        i1, i2 = np.triu_indices(n, 1)
        D = p1[i2] - p1[i1]
    else:
        p2 = np.atleast_2d(np.asarray(p2, dtype=float))
        D = (p2[np.newaxis, :, :] - p1[:, np.newaxis, :]).reshape(-1, 3)
    if cell is not None and pbc is not None and np.any(pbc):
        D, D_len = find_mic(D, cell, pbc)
    else:
        D_len = np.linalg.norm(D, axis=1)
    if p2 is None:
        Dout = np.zeros((n, n, 3))
        Dout[(i1, i2)] = D
        Dout -= np.transpose(Dout, axes=(1, 0, 2))
        Lout = np.zeros((n, n))
        Lout[(i1, i2)] = D_len
        Lout += Lout.T
        return Dout, Lout
    return D.reshape(len(p1), len(p2), 3), D_len.reshape(len(p1), len(p2))
