"""CPU: host mirrors of the reference's own enumerators / SRO / selection against fixtures minted by RUNNING the
reference's code (scripts/make_ref_fixtures.py: rlsim/actions/action.py and rlsim/utils/sro.py imported unmodified,
action.py:163-198 de-commented, simulator.py:94-98 exec'd).  These pin rows A1, A5, f3 of SURVEY.md §8 to the
reference rather than to a restatement."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLD
from rlvacdiffsim_b200.actions import get_action_space, get_action_space_legacy
from rlvacdiffsim_b200.simulator import sample_from_q
from rlvacdiffsim_b200.sro import get_sro_from_atoms
from rlvacdiffsim_b200.structures import Atoms, apply_action

A0 = 3.528
REF = np.load(os.path.join(GOLD, "ref_actions_sro.npz"))
SEL = np.load(os.path.join(GOLD, "ref_selection.npz"))
NAMES = [str(n) for n in REF["names"]]


def _atoms(name):
    return Atoms(REF[f"{name}.numbers"], REF[f"{name}.positions"], REF[f"{name}.cell"])


def _check_actions(got, ref):
    got = np.asarray(got, dtype=np.float64)
    assert got.shape == ref.shape
    assert np.array_equal(got[:, 0], ref[:, 0])                       # atom indices and row order: exact
    np.testing.assert_allclose(got[:, 1:], ref[:, 1:], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", [n for n in NAMES if f"{n}.live" in REF.files])
def test_live_enumerator_equals_reference(name):
    _check_actions(get_action_space(_atoms(name), A0), REF[f"{name}.live"])                     # action.py:12-61


@pytest.mark.parametrize("name", [n for n in NAMES if f"{n}.legacy" in REF.files] + ["cubic4_v2"])
def test_legacy_enumerator_equals_reference(name):
    _check_actions(get_action_space_legacy(_atoms(name), A0), REF[f"{name}.legacy"])           # action.py:163-198


def test_reference_asserts_are_mirrored():
    assert int(REF["demo.live_asserts"]) == 1 and int(REF["cubic4_v2.live_asserts"]) == 1
    msg = bytes(REF["cubic4_v2.live_assert_message"]).decode()
    with pytest.raises(AssertionError) as exc:
        get_action_space(_atoms("cubic4_v2"), A0)                                               # action.py:35
    assert str(exc.value) == msg


@pytest.mark.parametrize("name", NAMES)
def test_sro_equals_reference(name):
    at = _atoms(name)
    assert np.array_equal(get_sro_from_atoms(at), REF[f"{name}.sro"])                           # sro.py:49-83, exact
    key = f"{name}.sro_products"
    if key in REF.files:
        acts = REF[f"{name}.live"] if f"{name}.live" in REF.files else REF[f"{name}.legacy"]
        for act, ref in zip(acts, REF[key]):
            assert np.array_equal(get_sro_from_atoms(apply_action(at, act)), ref)


@pytest.mark.parametrize("name", [str(n) for n in SEL["names"]])
def test_selection_equals_reference(name):
    """simulator.py:94-98: softmax(Q / (T kb)) in fp32, np.random.choice on the global RNG."""
    Q, T = torch.from_numpy(SEL[f"{name}.Q"]), float(SEL[f"{name}.T"])
    for seed, want in enumerate(SEL[f"{name}.picks"]):
        np.random.seed(seed)
        act, probs = sample_from_q(Q, T)
        assert int(act) == int(want)
    assert np.array_equal(probs.numpy(), SEL[f"{name}.probs"])
    assert float(SEL["kb"]) == 8.617 * 10 ** -5
