"""CPU: candidate-hop enumeration against rlsim/actions/action.py semantics."""
import os

import numpy as np
import pytest

from conftest import GOLD, load_gold
from rlvacdiffsim_b200.actions import get_action_space, get_action_space_legacy, lattice_sites_for_cell
from rlvacdiffsim_b200.structures import fcc_sites, make_crconi_supercell, read_poscar

A0 = 3.528


def test_legacy_on_demo_matches_survey():
    at = read_poscar(os.path.join(GOLD, "POSCAR_0"))
    acts = get_action_space_legacy(at, A0)
    assert len(acts) == 12
    assert sorted({a[0] for a in acts}) == [41, 43, 50, 51, 52, 53, 62, 117, 155, 168, 173, 200]
    np.testing.assert_allclose(acts[0], [41, -1.4112, 1.4112, 0.0], atol=1e-12)
    for a in acts:
        assert np.linalg.norm(a[1:]) == pytest.approx(0.8 * A0 / np.sqrt(2))


def test_live_on_cubic_single_vacancy():
    s = make_crconi_supercell(4, seed=0)
    acts = get_action_space(s, A0)
    assert len(acts) == 12 and len({a[0] for a in acts}) == 12
    sites = fcc_sites(4, A0)
    for a in acts:
        assert np.linalg.norm(a[1:]) == pytest.approx(A0 / np.sqrt(2))
        # the moved atom lands on a perfect, unoccupied lattice site
        new = s.positions[a[0]] + np.array(a[1:])
        d = sites - new
        d -= 14.112 * np.round(d / 14.112)
        assert np.min(np.linalg.norm(d, axis=1)) < 1e-9
    g = load_gold("syn255_s0_live")
    np.testing.assert_allclose(np.asarray(acts), g["actions"], atol=1e-12)


def test_live_asserts_one_vacancy_like_reference():
    s = make_crconi_supercell(4, n_vac=2, seed=0)
    with pytest.raises(AssertionError, match="Expected 1 vacancy"):
        get_action_space(s, A0)                                           # action.py:35
    assert len(get_action_space(s, A0, max_vacancies=2)) == 24


def test_legacy_is_live_scaled_by_0p8():
    s = make_crconi_supercell(4, seed=3)
    live = {a[0]: np.array(a[1:]) for a in get_action_space(s, A0)}
    legacy = {a[0]: np.array(a[1:]) for a in get_action_space_legacy(s, A0)}
    assert live.keys() == legacy.keys()
    for k in live:
        np.testing.assert_allclose(legacy[k], 0.8 * live[k], atol=1e-9)


def test_generalised_sites_on_rhombohedral_demo():
    at = read_poscar(os.path.join(GOLD, "POSCAR_0"))
    sites = lattice_sites_for_cell(at.cell, A0)
    assert len(sites) == 216
    acts = get_action_space(at, A0)                 # the reference's live version trips its assert here
    assert sorted(a[0] for a in acts) == [41, 43, 50, 51, 52, 53, 62, 117, 155, 168, 173, 200]


def test_config3_eight_vacancies():
    g = load_gold("syn2040_v8")
    s = make_crconi_supercell(8, n_vac=8, seed=0)
    acts = get_action_space(s, A0, max_vacancies=8)
    assert len(acts) == int(g["n_actions_total"]) == 96
    np.testing.assert_allclose(np.asarray(acts), g["all_actions"], atol=1e-12)
