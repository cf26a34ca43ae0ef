"""The callers of the hot path: ``RLSimulator.select_action`` and a replica-batched variant.

:class:`RLSimulator` keeps the reference's signature and return values (``rlsim/drl/simulator.py:26-103``):
candidates → ``convert_to_graph_list`` → ``ReactionDataset`` → ``DataLoader(batch_size=16)`` →
``batch_to`` → ``model(batch, q_params)["rl_q"]`` → softmax(Q / kT) → ``np.random.choice``.

:class:`ReplicaScorer` is the B200-first entry the reference lacks: the reference loops replicas
serially (``rlsim/drl/deploy.py:75-94``); here all candidates of many independent replicas go through
ONE call of the C ABI from pinned host buffers, and replicas shard across ranks with no data-path
collective.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from .actions import get_action_space, get_action_space_legacy
from .engine import Engine
from .graph import DataLoader, ReactionDataset, ReactionGraph, batch_to
from .structures import Atoms, apply_action

KB_EV = 8.617 * 10 ** -5   # simulator.py:40


def sample_from_q(Q: torch.Tensor, temperature: float):
    """``simulator.py:94-98``: fp32 ``softmax(Q / (T kb))`` then ``np.random.choice`` on the global numpy RNG.
    Returns ``(index, action_probs)``; pinned against the reference's own lines by tests/test_ref_fixtures.py."""
    action_probs = nn.Softmax(dim=0)(Q / (temperature * KB_EV))
    p = action_probs.detach().cpu().numpy()
    return np.random.choice(len(p), p=p), action_probs


def convert_to_graph_list(atoms, actions: Sequence[Sequence[float]]) -> List[ReactionGraph]:
    """``simulator.py:439-460`` / ``trainer.py:185-205``: one reaction graph per candidate hop."""
    return [ReactionGraph.from_ase(atoms, apply_action(atoms, act)) for act in actions]


class LatticeEnv:
    """The slice of ``rlsim.environment.Environment`` the scoring path touches: ``atoms``,
    ``calc_params["device"]``, ``step`` (``environment.py:220-241``) and ``normalize_positions``.
    Relaxation / NEB (ASE + MACE) are outside this build's scope."""

    def __init__(self, atoms, calc_params: Optional[dict] = None):
        self.atoms = atoms
        self.calc_params = dict(calc_params or {"device": "cuda"})
        self.n_atom = len(atoms)

    def get_calculator(self, **_kw):
        return None

    def step(self, action: Sequence[float]) -> None:
        self.atoms = apply_action(self.atoms, action)

    def normalize_positions(self) -> None:
        if hasattr(self.atoms, "wrap"):
            self.atoms.wrap()


class RLSimulator:
    def __init__(self, environment, model=None,
                 q_params: Dict[str, float | bool] | None = {"alpha": 0.0, "beta": 0.5, "dqn": True},
                 sro_pixel=None, save_sro: bool = False, **kwargs):
        self.env = environment
        self.calculator = self.env.get_calculator(**self.env.calc_params) if hasattr(self.env, "get_calculator") else None
        self.q_params = q_params        # the caller's dict itself, like simulator.py:33: a Trainer sharing it sees the temperature
        self.model = model
        self.device = self.env.calc_params["device"]
        self.kb = KB_EV
        self.sro_pixel = sro_pixel
        self.save_sro = save_sro
        self.kwargs = kwargs

    def update_q_params(self, **new_q_params) -> None:
        self.q_params.update(**new_q_params)

    def select_action(self, action_space, temperature):
        self.update_q_params(**{"temperature": temperature})
        self.model.to(self.device)
        self.model.eval()
        dataset = ReactionDataset(convert_to_graph_list(self.env.atoms, action_space))
        total_q_list = []
        dataloader = DataLoader(dataset, batch_size=16)
        with torch.no_grad():
            for batch in dataloader:
                batch = batch_to(batch, self.device)
                rl_q = self.model(batch, q_params=self.q_params)["rl_q"]
                total_q_list.append(rl_q.detach())
        Q = torch.concat(total_q_list, dim=-1)
        sro_list = None
        valid = None
        if self.sro_pixel is not None:                         # simulator.py:60-92: keep candidates whose product SRO is in the pixel
            from .sro import get_sro_from_atoms
            valid, sros = [], []
            atoms = self.env.atoms.copy()
            pos = atoms.get_positions()
            for i, action in enumerate(action_space):
                pos[int(action[0])] += np.array(action[1:]) * 1 / 0.8
                atoms.set_positions(pos)
                sro = get_sro_from_atoms(atoms)
                sros.append(sro)
                pos[int(action[0])] -= np.array(action[1:]) * 1 / 0.8
                d = np.diag(sro)
                if len(self.sro_pixel) == 4:
                    x0, y0, z0, L = self.sro_pixel
                    if x0 - L / 2 < d[0] < x0 + L / 2 and y0 - L / 2 < d[1] < y0 + L / 2 and z0 - L / 2 < d[2] < z0 + L / 2:
                        valid.append(i)
                elif len(self.sro_pixel) == 2:
                    x0, L = self.sro_pixel
                    scalar = np.sqrt(np.sum(sro[0, :] ** 2) + np.sum(sro[1, 1:] ** 2) + sro[2, 2] ** 2)
                    if x0 - L / 2 < scalar < x0 + L / 2:
                        valid.append(i)
            vt = torch.tensor(valid, device=Q.device, dtype=torch.long)
            Q = Q[vt]
            sro_list = torch.tensor(np.asarray(sros), device=Q.device)[vt]
        action, action_probs = sample_from_q(Q, temperature)
        if valid is not None:
            action = valid[action]                             # simulator.py:99-101
        return action, action_probs, Q, sro_list

    def _action_space(self):
        enumerator = self.kwargs.get("enumerator", "live")
        a = self.kwargs.get("lattice_parameter", None)
        if enumerator == "legacy":
            return get_action_space_legacy(self.env, a if a is not None else 3.528)
        return get_action_space(self.env, lattice_parameter=a) if a is not None else get_action_space(self.env)

    def run_LSS(self, horizon: int, temperature_schedule) -> Tuple[list, list]:
        """Scoring-only LSS loop (``simulator.py:187-236`` without the MACE relaxations / file writes)."""
        Qlist, action_idx_list = [[]], [[]]
        for tstep in range(horizon):
            T = temperature_schedule(tstep) if callable(temperature_schedule) else temperature_schedule
            action_space = self._action_space()
            act_id, _, Q, _ = self.select_action(action_space, T)
            self.env.step(action_space[act_id])
            Qlist.append(Q.tolist())
            action_idx_list.append(int(act_id))
        return Qlist, action_idx_list

    def run_TKS(self, horizon: int, temperature: float):
        """``simulator.py:238-287``: residence-time clock ``dt = 1e-6 / sum(exp(Q/kT))``."""
        assert not self.q_params["dqn"], "TKS is only available for dqn==False."
        kT = temperature * 8.617 * 10 ** -5
        tlist, Qlist, action_idx_list = [0], [[]], [None]
        for _ in range(horizon):
            action_space = self._action_space()
            act_id, _, Q, _ = self.select_action(action_space, temperature)
            self.env.step(action_space[act_id])
            Gamma = float(torch.sum(torch.exp(Q / kT)))
            tlist.append(tlist[-1] + 1 / Gamma * 10 ** -6)
            Qlist.append(Q.tolist())
            action_idx_list.append(int(act_id))
        return Qlist, action_idx_list, tlist


class ThermalAnnealing:
    """``simulator.py:418-436``."""

    def __init__(self, total_horizon, T_start, T_end, annealing_time=None, T_speed=None):
        if T_speed is not None and annealing_time is None:
            self.annealing_time = int((T_start - T_end) / T_speed)
        elif annealing_time is not None and T_speed is None:
            self.annealing_time = annealing_time
        else:
            raise ValueError("Annealing time is not given")
        assert self.annealing_time <= total_horizon
        self.T_start, self.T_end = T_start, T_end

    def get_temperature(self, tstep):
        if tstep <= self.annealing_time:
            return self.T_start - (self.T_start - self.T_end) * tstep / self.annealing_time
        return self.T_end


# ----------------------------------------------------------------------------- replica-batched scoring

def pack_replicas(replicas: Sequence[Atoms], action_spaces: Sequence[Sequence[Sequence[float]]],
                  dedup_reactants: bool = False):
    """Flatten many replicas × candidates into the structure/pair arrays of ``rlvd_score_reactions_host``.

    Like-for-like with the reference (``dedup_reactants=False``) every candidate carries its own reactant
    structure (2 passes per Q-eval); with ``dedup_reactants=True`` the candidates of a replica share one."""
    Zs, poss, cells, counts, rs, ps, owner = [], [], [], [], [], [], []
    n_struct = 0
    for r, (atoms, acts) in enumerate(zip(replicas, action_spaces)):
        Z = np.asarray(atoms.get_atomic_numbers(), dtype=np.int32)
        P = np.asarray(atoms.get_positions(), dtype=np.float64)
        cell = np.asarray(atoms.get_cell(), dtype=np.float64)
        if dedup_reactants:
            Zs.append(Z); poss.append(P.astype(np.float32)); cells.append(cell); counts.append(len(Z))
            shared = n_struct
            n_struct += 1
        for act in acts:
            if not dedup_reactants:
                Zs.append(Z); poss.append(P.astype(np.float32)); cells.append(cell); counts.append(len(Z))
                rs.append(n_struct)
                n_struct += 1
            else:
                rs.append(shared)
            Pp = P.copy()
            Pp[int(act[0])] += np.asarray(act[1:4], dtype=np.float64)
            Zs.append(Z); poss.append(Pp.astype(np.float32)); cells.append(cell); counts.append(len(Z))
            ps.append(n_struct)
            n_struct += 1
            owner.append(r)
    node_ptr = np.zeros(n_struct + 1, dtype=np.int32)
    np.cumsum(np.asarray(counts, dtype=np.int64), out=node_ptr[1:])
    return {"Z": np.concatenate(Zs), "pos": np.concatenate(poss), "cells": np.stack(cells), "node_ptr": node_ptr,
            "react_struct": np.asarray(rs, dtype=np.int32), "prod_struct": np.asarray(ps, dtype=np.int32),
            "owner": np.asarray(owner, dtype=np.int64)}


class ReplicaScorer:
    """Scores every candidate of many independent replicas in one device pass and selects actions."""

    def __init__(self, model, device: int | str = 0, dedup_reactants: bool = False, max_structs_per_call: int = 4096):
        self.model = model
        self.device = torch.device(device) if not isinstance(device, int) else torch.device("cuda", device)
        self.dedup = dedup_reactants
        self.max_structs = max_structs_per_call
        self.engine: Engine = model._engine_for(self.device)

    def score(self, replicas: Sequence[Atoms], action_spaces, q_params: dict,
              temperatures: Optional[Sequence[float]] = None) -> List[np.ndarray]:
        """→ per-replica arrays of rl_q (one entry per candidate)."""
        qp = self.engine.q_params(q_params, self.model.head_spec)
        out: List[np.ndarray] = []
        per = 2 if not self.dedup else 1
        start = 0
        while start < len(replicas):
            n_structs, stop = 0, start
            while stop < len(replicas):
                need = len(action_spaces[stop]) * per + (1 if self.dedup else 0)
                if stop > start and n_structs + need > self.max_structs:
                    break
                n_structs += need
                stop += 1
            pk = pack_replicas(replicas[start:stop], action_spaces[start:stop], self.dedup)
            temp = None
            if temperatures is not None:
                temp = np.asarray(temperatures, dtype=np.float32)[start:stop][pk["owner"]]
            res = self.engine.score_host(pk["Z"], pk["pos"], pk["cells"], pk["node_ptr"], pk["react_struct"],
                                         pk["prod_struct"], qp, temp)
            q = res["rl_q"]
            for r in range(stop - start):
                out.append(q[pk["owner"] == r])
            start = stop
        return out

    @staticmethod
    def select(Q: np.ndarray, temperature: float, rng: Optional[np.random.Generator] = None) -> Tuple[int, np.ndarray]:
        """softmax(Q / kT) in fp32 like ``simulator.py:94`` then a draw (global ``np.random`` if no rng)."""
        z = torch.from_numpy(np.asarray(Q, dtype=np.float32)) / (temperature * KB_EV)
        p = torch.softmax(z, dim=0).numpy()
        idx = int(np.random.choice(len(p), p=p)) if rng is None else int(rng.choice(len(p), p=p / p.sum()))
        return idx, p


class DeviceLSS:
    """Whole LSS steps for many replicas without a host round trip.

    The reference's loop (``rlsim/drl/simulator.py:187-236`` per replica, ``deploy.py:75-94`` over replicas) does, per
    step and per replica on the host: ``get_action_space`` (ASE + numpy), ``convert_to_graph_list``, collate, H2D, the
    model call, softmax + ``np.random.choice``, ``env.step``.  Here every stage is a kernel over all replicas:
    ``rlvd_action_space`` → ``rlvd_candidate_offsets`` → ``rlvd_expand_candidates`` → K1..K5 → ``rlvd_select_apply``;
    the replicas' fp64 positions never leave the device.  One 4-byte D2H per step (the candidate count, to size the
    batch).  Replicas must share one atom count and have orthorhombic cells with heights >= 2 x cutoff (the BASELINE
    supercells); relaxation / NEB between steps is outside this build's scope."""

    def __init__(self, model, replicas: Sequence[Atoms], device: int | str = 0, lattice_parameter: float = 3.528,
                 max_vacancies: int = 1, q_params: Optional[dict] = None):
        self.model = model
        self.device = torch.device(device) if not isinstance(device, int) else torch.device("cuda", device)
        self.engine: Engine = model._engine_for(self.device)
        self.a = float(lattice_parameter)
        self.max_vac = int(max_vacancies)
        self.q_params = dict(q_params or {"alpha": 0.0, "beta": 1.0, "dqn": True})
        n = {len(r) for r in replicas}
        if len(n) != 1:
            raise ValueError("DeviceLSS needs replicas of one size")
        self.n_atoms = n.pop()
        self.R = len(replicas)
        for r in replicas:
            c = np.asarray(r.get_cell(), dtype=np.float64)
            if np.abs(c - np.diag(np.diag(c))).max() != 0.0 or np.diag(c).min() < 2 * self.engine.cutoff:
                raise ValueError("DeviceLSS needs orthorhombic cells with heights >= 2 x cutoff")
        dev = self.device
        self.pos = torch.from_numpy(np.concatenate([np.asarray(r.get_positions(), dtype=np.float64) for r in replicas])).to(dev)
        self.Z = torch.from_numpy(np.concatenate([np.asarray(r.get_atomic_numbers(), dtype=np.int32) for r in replicas])).to(dev)
        self.cell = torch.from_numpy(np.stack([np.asarray(r.get_cell(), dtype=np.float64) for r in replicas])).to(dev)
        self.rep_ptr = torch.arange(0, (self.R + 1) * self.n_atoms, self.n_atoms, dtype=torch.int32, device=dev)

    def positions(self) -> np.ndarray:
        return self.pos.cpu().numpy().reshape(self.R, self.n_atoms, 3)

    def step(self, temperature: float, uniforms: Optional[np.ndarray] = None, apply: bool = True) -> Dict[str, torch.Tensor]:
        """One LSS step of every replica.  ``uniforms[R]`` are the inverse-cdf draws (``np.random.random_sample(R)``
        reproduces the reference's ``np.random.choice`` stream); default: a fresh draw from the global numpy RNG."""
        eng, dev = self.engine, self.device
        count, actions, status = eng.action_space(self.pos, self.cell, self.rep_ptr, self.a, self.max_vac,
                                                  max_actions=12 * self.max_vac)
        first, total = eng.candidate_offsets(count)
        head = torch.cat([total, status.max().reshape(1).to(torch.int32)]).cpu().numpy()     # the step's one D2H
        if head[1] != 0:
            raise AssertionError("Expected 1 vacancy" if head[1] == 2 and self.max_vac == 1 else
                                 f"rlvd_action_space status {int(head[1])}")
        G = int(head[0])
        b = eng.expand_candidates(self.rep_ptr, first, G, self.n_atoms, self.Z, self.pos, self.cell, actions)
        qp = eng.q_params(dict(self.q_params, temperature=float(temperature)), self.model.head_spec)
        out = eng.score_device(b["Z"], b["pos"], b["cell"], b["inv"], b["img"], b["ptr"], b["rs"], b["ps"], qp,
                               max_struct_nodes=self.n_atoms)
        u = np.random.random_sample(self.R) if uniforms is None else np.asarray(uniforms, dtype=np.float64)
        T = torch.full((self.R,), float(temperature), dtype=torch.float32, device=dev)
        chosen, amax, probs, gamma = eng.select_apply(self.rep_ptr, first, out["rl_q"], T, torch.from_numpy(u).to(dev),
                                                      actions, self.pos, apply)
        return {"chosen": chosen, "argmax": amax, "probs": probs, "rl_q": out["rl_q"], "first": first, "actions": actions,
                "count": count, "gamma": gamma}

    def run(self, horizon: int, temperature: float, mode: str = "lss", writer=None,
            uniforms: Optional[np.ndarray] = None) -> Dict[str, list]:
        """``horizon`` steps of every replica: the loop of ``RLSimulator.run_LSS`` / ``run_TKS``
        (``rlsim/drl/simulator.py:187-236, 239-283``) without relaxation or energy logging.  ``mode="tks"`` also keeps
        the residence-time clock ``t += 1e-6 / sum(exp(Q / kT))`` (simulator.py:262-265) and the tracked atom's
        coordinates (the LAST atom, simulator.py:241,271).  ``writer``: a ``TrajectoryWriter`` fed once per step.
        ``uniforms[horizon, R]`` pins the draws.  Returns per-replica lists shaped like the reference's outputs."""
        if mode not in ("lss", "tks"):
            raise ValueError("DeviceLSS.run supports mode 'lss' and 'tks'")
        if mode == "tks":
            assert not self.q_params["dqn"], "TKS needs rates, not DQN Q-values (simulator.py:181)"
        R = self.R
        q_lists = [[] for _ in range(R)]
        act_lists = [[] for _ in range(R)]
        t_lists = [[0.0] for _ in range(R)]
        c_lists = [[self.positions()[r, -1].tolist()] for r in range(R)] if mode == "tks" else None
        for t in range(horizon):
            out = self.step(temperature, None if uniforms is None else uniforms[t])
            first = out["first"].cpu().numpy()
            q = out["rl_q"].cpu().numpy()
            chosen = out["chosen"].cpu().numpy()
            gamma = out["gamma"].cpu().numpy()
            pos = None
            if writer is not None or mode == "tks":
                pos = self.positions()
            for r in range(R):
                q_lists[r].append(q[first[r]:first[r + 1]].tolist())
                act_lists[r].append(int(chosen[r]))
                if mode == "tks":
                    t_lists[r].append(t_lists[r][-1] + 1.0 / float(gamma[r]) * 10 ** -6)
                    c_lists[r].append(pos[r, -1].tolist())
            if writer is not None:
                writer.append(pos, q_values=[q[first[r]:first[r + 1]] for r in range(R)], chosen=chosen,
                              times=[t_lists[r][-1] for r in range(R)] if mode == "tks" else None)
        res = {"Q": q_lists, "actions": act_lists}
        if mode == "tks":
            res["time"] = t_lists
            res["coords"] = c_lists
        return res
