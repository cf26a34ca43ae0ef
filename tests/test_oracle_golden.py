"""CPU: the oracle against the committed golden vectors and the facts SURVEY.md Appendix D records."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLD, load_gold
from oracle.painn_oracle import (PaiNNOracle, action_probabilities, batched_radius_graph, build_reaction_batch,
                                 image_range, radius_graph_pbc, sample_action, tks_time_step)
from rlvacdiffsim_b200.structures import Atoms, cell_heights, make_crconi_supercell, read_poscar


def _digest(ri, sj, S):
    arr = np.concatenate([ri.astype(np.int64)[:, None], sj.astype(np.int64)[:, None], S.astype(np.int64)], axis=1)
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def m32():
    return PaiNNOracle.load(os.path.join(GOLD, "model_trained"))


def test_demo_structure_facts():
    at = read_poscar(os.path.join(GOLD, "POSCAR_0"))                    # demo/poscars/POSCAR_0
    assert len(at) == 215
    assert sorted(np.bincount(at.numbers)[[24, 27, 28]].tolist()) == [71, 72, 72]
    np.testing.assert_allclose(cell_heights(at.cell), 12.2213505, atol=1e-6)
    assert image_range(at.cell, 5.0).tolist() == [0, 0, 0]


@pytest.mark.parametrize("name,n_edges", [("demo", 11556), ("syn255_s0_live", 13716), ("syn107_s0", 5724),
                                           ("syn255_s2_jitter", 12378), ("syn31_multiimage", 1620)])
def test_edge_sets_match_golden(name, n_edges):
    g = load_gold(name)
    ri, sj, S = radius_graph_pbc(g["positions"], g["cell"], 5.0)
    assert len(ri) == n_edges == int(g["n_edges"])
    assert _digest(ri, sj, S) == bytes(g["edge_sha256"]).decode()
    if "edge_recv" in g:
        assert np.array_equal(ri, g["edge_recv"]) and np.array_equal(sj, g["edge_send"])
        assert np.array_equal(S, g["edge_shift"])
    # receiver-sorted, sender ascending inside a row
    assert np.all(np.diff(ri) >= 0)
    same = np.diff(ri) == 0
    assert np.all(np.diff(sj)[same] >= 0)


def test_edge_symmetry_and_counts():
    g = load_gold("syn255_s0_live")
    ri, sj, S = g["edge_recv"], g["edge_send"], g["edge_shift"]
    fwd = set(zip(ri.tolist(), sj.tolist(), map(tuple, S.tolist())))
    assert all((j, i, (-s[0], -s[1], -s[2])) in fwd for i, j, s in fwd)   # (i<-j,S) <=> (j<-i,-S)
    deg = np.bincount(ri, minlength=255)
    assert deg.max() == 54 and deg.min() >= 42                            # 54 FCC neighbours inside 5.0 Å


def test_multi_image_cell_has_repeated_pairs():
    g = load_gold("syn31_multiimage")
    assert image_range(g["cell"], 5.0).tolist() == [1, 1, 1]
    pairs = list(zip(g["edge_recv"].tolist(), g["edge_send"].tolist())) if "edge_recv" in g else None
    ri, sj, S = radius_graph_pbc(g["positions"], g["cell"], 5.0)
    pairs = list(zip(ri.tolist(), sj.tolist()))
    assert len(pairs) > len(set(pairs))                                    # same (i,j), different images
    assert not np.any(ri == sj)                                            # box 7.056 Å > rc: no self-images
    tiny = make_crconi_supercell(1, n_vac=0, seed=0)                       # box 3.528 Å < rc: self-images appear
    ti, tj, tS = radius_graph_pbc(tiny.positions, tiny.cell, 5.0)
    assert np.any(ti == tj) and not np.any((ti == tj) & (np.abs(tS).sum(1) == 0))


def test_brute_force_images_agree_with_rounding():
    """Rounding rule vs exhaustive 27-image search in fp64 on a skewed cell (demo is rhombohedral)."""
    at = read_poscar(os.path.join(GOLD, "POSCAR_0"))
    ri, sj, S = radius_graph_pbc(at.positions, at.cell, 5.0)
    pos = at.positions.astype(np.float32).astype(np.float64)
    cell = at.cell.astype(np.float32).astype(np.float64)
    count = 0
    for s in np.ndindex(3, 3, 3):
        sh = (np.array(s) - 1.0) @ cell
        d = pos[None, :, :] - pos[:, None, :] + sh
        d2 = np.einsum("ijk,ijk->ij", d, d)
        m = d2 < 25.0
        if not sh.any():
            m &= ~np.eye(len(pos), dtype=bool)
        count += int(m.sum())
    assert count == len(ri)


@pytest.mark.parametrize("name", ["demo", "syn255_s0_live", "syn31_multiimage"])
def test_oracle_outputs_match_golden(name, m32):
    g = load_gold(name)
    tag = "lss700"
    qp = dict(zip(("alpha", "beta", "dqn", "temperature"), g[f"{tag}.q_params"].tolist()))
    qp["dqn"] = bool(qp["dqn"])
    Z, pR, pP, cells, ptr = build_reaction_batch(g["numbers"], g["positions"], g["cell"], g["actions"].tolist())
    with torch.no_grad():
        out = m32.reaction_forward(Z, pR, pP, cells, ptr, qp)
    for k in ("rl_q", "q0", "q1", "delta_e", "E_R", "E_P"):
        ref64 = g[f"{tag}.f64.{k}"]
        got = out[k].numpy().astype(np.float64)
        scale = np.maximum(np.abs(ref64), 1.0)
        assert np.max(np.abs(got - ref64) / scale) < 2e-4, k      # fp32 eager vs fp64: its own rounding level
        np.testing.assert_allclose(got, g[f"{tag}.f32.{k}"], rtol=2e-4, atol=2e-4)
    assert int(np.argmax(out["rl_q"].numpy())) == int(g[f"{tag}.argmax"])


def test_selection_and_clock_helpers():
    g = load_gold("demo")
    Q = torch.from_numpy(g["tks900.f64.rl_q"])
    p = action_probabilities(Q, 900.0)
    assert abs(float(p.sum()) - 1.0) < 1e-12
    assert sample_action(torch.from_numpy(g["tks900.f32.probs"]), seed=0) == int(g["tks900.sampled_seed0"])
    assert tks_time_step(Q, 900.0) == pytest.approx(float(g["tks900.tks_dt"]), rel=1e-12)


def test_energy_head_reproduces_readme_scale():
    """README.md:53-56 logs MACE energies of about -1595 eV for the demo cell; the restated energy head gives
    the same scale, the one number of the reference this oracle can be compared to at all."""
    g = load_gold("demo")
    assert -1596.5 < float(g["lss700.f64.E_R"][0]) < -1593.5
