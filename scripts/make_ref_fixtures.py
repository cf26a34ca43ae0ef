"""Mint ``tests/golden/ref_*.npz`` by RUNNING THE REFERENCE'S OWN CODE (build container only: needs /root/reference).

    python scripts/make_ref_fixtures.py

What is executed, unmodified, straight from ``/root/reference``:

* ``rlsim/actions/action.py`` — imported as a module (file path import, so ``rlsim/__init__`` chains are not needed):
  the live ``get_action_space`` (``action.py:12-61``);
* the commented-out "Previous version that worked perfectly for CrCoNi" (``action.py:162-198``): those lines are read
  from the file, the leading ``# `` is stripped, and the resulting source is exec'd in the module's namespace — the
  checkpoint was trained under this enumerator, and nothing else pins it;
* ``rlsim/utils/sro.py`` — imported as a module: ``get_sro_from_atoms`` (``sro.py:49-83``);
* ``rlsim/drl/simulator.py:40`` (``self.kb``) and ``:94-98`` (softmax + ``np.random.choice``): the file cannot be imported
  (rgnn / torch_geometric are absent), so exactly those source lines are read, dedented and exec'd with ``Q``,
  ``temperature`` and a seeded global numpy RNG.

ASE is absent from this image; ``tests/ref_stub/ase`` provides the few entry points the two modules call (see its
README).  The fixtures pin the host mirrors (``rlvacdiffsim_b200/actions.py``, ``sro.py``, ``simulator.py``) AND the
device kernels (``rlvd_action_space``, ``rlvd_sro_counts``, ``rlvd_select_apply``) to the reference's outputs:
``tests/test_ref_fixtures.py`` (CPU) and ``tests/test_gpu_parity.py`` (GPU).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import textwrap
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "tests", "ref_stub"))
sys.path.insert(0, ROOT)

import ase  # noqa: E402  (the stub)

from rlvacdiffsim_b200.structures import make_crconi_supercell, read_poscar  # noqa: E402  (input generators only)


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _lines(path, first, last):
    with open(path) as fh:
        src = fh.readlines()
    return src[first - 1:last]


def load_reference():
    action = _load(os.path.join(REF, "rlsim", "actions", "action.py"), "ref_action")
    sro = _load(os.path.join(REF, "rlsim", "utils", "sro.py"), "ref_sro")
    # action.py:162-198, de-commented
    legacy_src = "".join(ln[2:] if ln.startswith("# ") else ln.lstrip("#") for ln in
                         _lines(os.path.join(REF, "rlsim", "actions", "action.py"), 163, 198))
    ns = dict(action.__dict__)
    exec(compile(legacy_src, "action.py:163-198", "exec"), ns)
    legacy = ns["get_action_space"]
    assert legacy is not action.get_action_space
    # simulator.py:40 and :94-98
    sim = os.path.join(REF, "rlsim", "drl", "simulator.py")
    kb_src = textwrap.dedent("".join(_lines(sim, 40, 40)))
    sel_src = textwrap.dedent("".join(_lines(sim, 94, 98)))
    assert "self.kb" in kb_src and "np.random.choice" in sel_src and "Softmax" in sel_src

    def select(Q, temperature, seed):
        self = types.SimpleNamespace()
        scope = {"self": self, "np": np, "nn": torch.nn, "torch": torch, "Q": Q, "temperature": temperature}
        exec(kb_src, scope)
        np.random.seed(seed)
        exec(sel_src, scope)
        return int(scope["action"]), scope["action_probs"].numpy(), float(self.kb)

    return action, legacy, sro, select


def to_ase(at):
    return ase.Atoms(numbers=at.numbers, positions=at.positions, cell=at.cell, pbc=True)


class Env:  # the enumerators only touch ``config.atoms``
    def __init__(self, atoms):
        self.atoms = atoms


def main():
    action, legacy, sro, select = load_reference()
    a0 = 3.528
    rec = {}
    names = []

    def add(name, at, live=True, leg=True, hop_sro=True):
        env = Env(to_ase(at))
        rec[f"{name}.numbers"], rec[f"{name}.positions"], rec[f"{name}.cell"] = at.numbers, at.positions, at.cell
        acts_for_sro = None
        if live:
            acts = action.get_action_space(env, a0)
            rec[f"{name}.live"] = np.asarray(acts, dtype=np.float64)
            acts_for_sro = acts
        if leg:
            acts = legacy(env, a0)
            rec[f"{name}.legacy"] = np.asarray([[float(x) for x in a] for a in acts], dtype=np.float64)
            acts_for_sro = acts_for_sro or acts
        rec[f"{name}.sro"] = sro.get_sro_from_atoms(env.atoms)
        if hop_sro and acts_for_sro:
            # the candidate filter of select_action (simulator.py:60-70): SRO of every product state
            out = []
            for act in acts_for_sro:
                prod = env.atoms.copy()
                pos = prod.get_positions()
                pos[int(act[0])] += np.array(act[1:])
                prod.set_positions(pos)
                out.append(sro.get_sro_from_atoms(prod))
            rec[f"{name}.sro_products"] = np.stack(out)
        names.append(name)
        print(name, "live", len(rec.get(f"{name}.live", [])), "legacy", len(rec.get(f"{name}.legacy", [])))

    for seed in range(8):
        add(f"cubic4_s{seed}", make_crconi_supercell(4, seed=seed), hop_sro=(seed < 2))
    for seed in (2, 5):
        add(f"cubic4_s{seed}_jit", make_crconi_supercell(4, seed=seed, jitter=0.05), hop_sro=False)
    for seed in (0, 1):
        add(f"cubic3_s{seed}", make_crconi_supercell(3, seed=seed), hop_sro=(seed == 0))
    add("cubic5_s0", make_crconi_supercell(5, seed=0), hop_sro=False)
    # the reference's own demo input: rhombohedral cell — the live enumerator trips its assert, so
    # only the legacy one is recorded, and the assertion is recorded as a fact
    demo = read_poscar(os.path.join(REF, "demo", "poscars", "POSCAR_0"))
    add("demo", demo, live=False, hop_sro=True)
    try:
        action.get_action_space(Env(to_ase(demo)), a0)
        rec["demo.live_asserts"] = np.int64(0)
    except AssertionError as exc:
        rec["demo.live_asserts"] = np.int64(1)
        rec["demo.live_assert_message"] = np.frombuffer(str(exc).encode(), dtype=np.uint8)
    # two vacancies: the live version asserts (action.py:35)
    two = make_crconi_supercell(4, n_vac=2, seed=0)
    try:
        action.get_action_space(Env(to_ase(two)), a0)
        rec["cubic4_v2.live_asserts"] = np.int64(0)
    except AssertionError as exc:
        rec["cubic4_v2.live_asserts"] = np.int64(1)
        rec["cubic4_v2.live_assert_message"] = np.frombuffer(str(exc).encode(), dtype=np.uint8)
    rec["cubic4_v2.legacy"] = np.asarray([[float(x) for x in a] for a in legacy(Env(to_ase(two)), a0)])
    rec["cubic4_v2.numbers"], rec["cubic4_v2.positions"], rec["cubic4_v2.cell"] = two.numbers, two.positions, two.cell
    rec["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "ref_actions_sro.npz"), **rec)

    # selection: simulator.py:94-98 on the golden Q vectors (fp32, like the reference's device tensors)
    sel = {}
    g = np.load(os.path.join(GOLD, "golden_demo.npz"))
    cases = {"demo_lss1200": (g["lss1200.f32.rl_q"], 1200.0), "demo_lss700": (g["lss700.f32.rl_q"], 700.0),
             "demo_tks900": (g["tks900.f32.rl_q"], 900.0)}
    rng = np.random.default_rng(0)
    cases["random12"] = (rng.normal(0.0, 0.05, 12).astype(np.float32), 700.0)
    cases["random96"] = (rng.normal(-0.8, 0.1, 96).astype(np.float32), 900.0)
    for name, (Q, T) in cases.items():
        Qt = torch.from_numpy(np.asarray(Q, dtype=np.float32))
        picks = []
        for seed in range(16):
            a, p, kb = select(Qt, T, seed)
            picks.append(a)
        sel[f"{name}.Q"], sel[f"{name}.T"], sel[f"{name}.probs"] = Qt.numpy(), np.float64(T), p
        sel[f"{name}.picks"] = np.asarray(picks, dtype=np.int64)
        sel["kb"] = np.float64(kb)
    sel["names"] = np.array(list(cases))
    np.savez_compressed(os.path.join(GOLD, "ref_selection.npz"), **sel)
    print("wrote ref_actions_sro.npz, ref_selection.npz")


if __name__ == "__main__":
    main()
