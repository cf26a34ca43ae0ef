"""Mint the golden fixtures under tests/golden/ from the CPU oracle (SURVEY.md §8c "golden vectors the new
repo must mint itself").  Run in the build container, where /root/reference exists:

    python scripts/make_golden.py

It copies the two DATA files of the reference the path needs at test time (the bundled checkpoint
``data/model_trained`` and ``demo/poscars/POSCAR_0`` — the GPU box has no /root/reference) and writes
``golden_*.npz``: inputs, canonical edge lists, and fp64 + fp32 oracle outputs.  No reference source
code is copied.  The oracle is parity-unpinned (see oracle/painn_oracle.py); these vectors pin the
CUDA path to the oracle, and the oracle to itself across refactors.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.painn_oracle import (PaiNNOracle, action_probabilities, batched_radius_graph,  # noqa: E402
                                 build_reaction_batch, radius_graph_pbc, sample_action, tks_time_step)
from rlvacdiffsim_b200.actions import get_action_space, get_action_space_legacy  # noqa: E402
from rlvacdiffsim_b200.structures import Atoms, make_crconi_supercell, read_poscar  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def edge_digest(recv, send, shift) -> str:
    arr = np.concatenate([np.asarray(recv, np.int64)[:, None], np.asarray(send, np.int64)[:, None],
                          np.asarray(shift, np.int64)], axis=1)
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


def score(model, atoms, actions, q_params):
    Z, pR, pP, cells, ptr = build_reaction_batch(atoms.numbers, atoms.positions, atoms.cell, actions)
    with torch.no_grad():
        out = model.reaction_forward(Z, pR, pP, cells, ptr, q_params)
    return {k: v.numpy() for k, v in out.items()}


def case(name, atoms, actions, q_param_sets, m32, m64, store_edges=True, store_q=False, extra=None):
    rec = {"numbers": atoms.numbers, "positions": atoms.positions, "cell": atoms.cell,
           "actions": np.asarray(actions, dtype=np.float64)}
    ri, sj, S = radius_graph_pbc(atoms.positions, atoms.cell, 5.0)
    rec["n_edges"] = np.int64(len(ri))
    rec["edge_sha256"] = np.frombuffer(edge_digest(ri, sj, S).encode(), dtype=np.uint8)
    if store_edges:
        rec["edge_recv"], rec["edge_send"], rec["edge_shift"] = ri.astype(np.int32), sj.astype(np.int32), S
    # product graph of candidate 0 (moved atom) as a second edge-set check
    Z, pR, pP, cells, ptr = build_reaction_batch(atoms.numbers, atoms.positions, atoms.cell, actions[:1])
    rp, sp, Sp = radius_graph_pbc(pP, atoms.cell, 5.0)
    rec["n_edges_product0"] = np.int64(len(rp))
    rec["edge_sha256_product0"] = np.frombuffer(edge_digest(rp, sp, Sp).encode(), dtype=np.uint8)
    for tag, qp in q_param_sets.items():
        o64, o32 = score(m64, atoms, actions, qp), score(m32, atoms, actions, qp)
        for k in ("rl_q", "q0", "q1", "delta_e", "E_R", "E_P"):
            rec[f"{tag}.f64.{k}"] = o64[k]
            rec[f"{tag}.f32.{k}"] = o32[k]
        T = qp["temperature"]
        p64 = action_probabilities(torch.from_numpy(o64["rl_q"]), T)
        p32 = action_probabilities(torch.from_numpy(o32["rl_q"]), T)
        rec[f"{tag}.f64.probs"], rec[f"{tag}.f32.probs"] = p64.numpy(), p32.numpy()
        rec[f"{tag}.argmax"] = np.int64(int(np.argmax(o64["rl_q"])))
        rec[f"{tag}.sampled_seed0"] = np.int64(sample_action(p32, seed=0))
        rec[f"{tag}.tks_dt"] = np.float64(tks_time_step(torch.from_numpy(o64["rl_q"]), T))
        rec[f"{tag}.q_params"] = np.array([qp["alpha"], qp["beta"], float(qp["dqn"]), qp["temperature"]])
    if store_q:
        with torch.no_grad():
            q, mu = m64.representation(Z[: len(atoms)], pR[: len(atoms)], cells[:1], ptr[:2], return_mu=True)
        rec["f64.q_reactant"] = q.numpy()
        rec["f64.mu_reactant"] = mu.numpy()
    if extra:
        rec.update(extra)
    np.savez_compressed(os.path.join(GOLD, f"golden_{name}.npz"), **rec)
    print(f"{name}: N={len(atoms)} E={int(rec['n_edges'])} actions={len(actions)}")


def main():
    os.makedirs(GOLD, exist_ok=True)
    shutil.copyfile(os.path.join(REF, "data", "model_trained"), os.path.join(GOLD, "model_trained"))
    shutil.copyfile(os.path.join(REF, "demo", "poscars", "POSCAR_0"), os.path.join(GOLD, "POSCAR_0"))
    m32 = PaiNNOracle.load(os.path.join(GOLD, "model_trained"), dtype=torch.float32)
    m64 = PaiNNOracle.load(os.path.join(GOLD, "model_trained"), dtype=torch.float64)
    LSS = lambda T: {"alpha": 0.0, "beta": 1.0, "dqn": True, "temperature": float(T)}     # demo/config.toml:9-12
    TKS = {"alpha": 1.0, "beta": 0.0, "dqn": False, "temperature": 900.0}                  # configs/deploy_drl_tks.toml

    demo = read_poscar(os.path.join(GOLD, "POSCAR_0"))
    acts = get_action_space_legacy(demo, 3.528)
    case("demo", demo, acts, {"lss1200": LSS(1200), "lss700": LSS(700), "tks900": TKS}, m32, m64, store_q=True)

    for seed in (0, 1):
        s = make_crconi_supercell(4, seed=seed)
        case(f"syn255_s{seed}_live", s, get_action_space(s, 3.528), {"lss700": LSS(700), "tks900": TKS}, m32, m64,
             store_edges=(seed == 0))
    s = make_crconi_supercell(4, seed=0)
    case("syn255_s0_legacy", s, get_action_space_legacy(s, 3.528), {"lss700": LSS(700)}, m32, m64, store_edges=False)
    s = make_crconi_supercell(4, seed=2, jitter=0.05)
    case("syn255_s2_jitter", s, get_action_space(s, 3.528), {"lss700": LSS(700)}, m32, m64)
    s = make_crconi_supercell(3, seed=0)
    case("syn107_s0", s, get_action_space(s, 3.528), {"lss700": LSS(700)}, m32, m64)
    # cell smaller than 2*cutoff: several periodic images per pair (image_range = 1)
    s = make_crconi_supercell(2, seed=0)
    case("syn31_multiimage", s, get_action_space(s, 3.528), {"lss700": LSS(700)}, m32, m64)
    # BASELINE config 3: 8x8x8, 8 vacancies, generalised enumerator; score the first 6 candidates
    s = make_crconi_supercell(8, n_vac=8, seed=0)
    acts = get_action_space(s, 3.528, max_vacancies=8)
    case("syn2040_v8", s, acts[:6], {"lss700": LSS(700)}, m32, m64, store_edges=False,
         extra={"n_actions_total": np.int64(len(acts)), "all_actions": np.asarray(acts, dtype=np.float64)})


if __name__ == "__main__":
    main()
