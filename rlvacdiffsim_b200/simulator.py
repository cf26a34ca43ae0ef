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
    """Whole LSS / TKS steps for many replicas without a host round trip.

    The reference's loop (``rlsim/drl/simulator.py:187-236`` per replica, ``deploy.py:75-94`` over replicas) does, per
    step and per replica on the host: ``get_action_space`` (ASE + numpy), ``convert_to_graph_list``, collate, H2D, the
    model call, softmax + ``np.random.choice``, ``env.step``.  Here every stage is a kernel over all replicas:
    ``rlvd_action_space`` (orthorhombic cells) / ``rlvd_action_space_general`` (any cell, e.g. the rhombohedral demo
    POSCAR) → ``rlvd_candidate_offsets`` → ``rlvd_expand_candidates_ex`` → K1..K5 → ``rlvd_select_apply``; the replicas'
    fp64 positions never leave the device.  :meth:`step` reads the candidate count back (one 8-byte D2H) and is the
    ``np.random``-compatible form; :meth:`run` with ``sync_free=True`` queues whole steps with NO host synchronisation
    (counter-based Philox draws on the device, worst-case buffer sizes, per-step results appended to device logs that are
    read once at the end).  Replicas must share one atom count; relaxation / NEB between steps is outside this build's
    scope."""

    def __init__(self, model, replicas: Sequence[Atoms], device: int | str = 0, lattice_parameter: float = 3.528,
                 max_vacancies: int = 1, q_params: Optional[dict] = None, seed: int = 0, replica_id0: int = 0):
        from .actions import lattice_sites_for_cell
        from .engine import cell_geometry
        self.model = model
        self.device = torch.device(device) if not isinstance(device, int) else torch.device("cuda", device)
        self.engine: Engine = model._engine_for(self.device)
        self.a = float(lattice_parameter)
        self.max_vac = int(max_vacancies)
        self.q_params = dict(q_params or {"alpha": 0.0, "beta": 1.0, "dqn": True})
        self.seed, self.replica_id0, self.step_index = int(seed), int(replica_id0), 0
        n = {len(r) for r in replicas}
        if len(n) != 1:
            raise ValueError("DeviceLSS needs replicas of one size")
        self.n_atoms = n.pop()
        self.R = len(replicas)
        cells = np.stack([np.asarray(r.get_cell(), dtype=np.float64) for r in replicas])
        self.general = bool(np.abs(cells - cells * np.eye(3)[None]).max() != 0.0)
        inv32, img = cell_geometry(cells, self.engine.cutoff)
        if not self.general and int(img.max()) > 0:
            raise ValueError("DeviceLSS: orthorhombic cells need heights >= 2 x cutoff (single minimum image)")
        dev = self.device
        self.pos = torch.from_numpy(np.concatenate([np.asarray(r.get_positions(), dtype=np.float64) for r in replicas])).to(dev)
        self.Z = torch.from_numpy(np.concatenate([np.asarray(r.get_atomic_numbers(), dtype=np.int32) for r in replicas])).to(dev)
        self.cell = torch.from_numpy(cells).to(dev)
        self.rep_ptr = torch.arange(0, (self.R + 1) * self.n_atoms, self.n_atoms, dtype=torch.int32, device=dev)
        self.inv32 = self.img = self.inv64 = self.sites = self.site_ptr = None
        if self.general:
            # the perfect-lattice sites and inverse cells are properties of the (fixed) cells: computed once, on the host
            cache: Dict[bytes, np.ndarray] = {}
            site_list = []
            for c in cells:
                key = c.tobytes()
                if key not in cache:
                    cache[key] = lattice_sites_for_cell(c, self.a)
                site_list.append(cache[key])
            sp = np.zeros(self.R + 1, dtype=np.int32)
            np.cumsum([len(x) for x in site_list], out=sp[1:])
            self.sites = torch.from_numpy(np.concatenate(site_list)).to(dev)
            self.site_ptr = torch.from_numpy(sp).to(dev)
            self.inv64 = torch.from_numpy(np.linalg.inv(cells)).to(dev)
            self.inv32 = torch.from_numpy(inv32).to(dev)
            self.img = torch.from_numpy(img).to(dev)

    def positions(self) -> np.ndarray:
        return self.pos.cpu().numpy().reshape(self.R, self.n_atoms, 3)

    def _enumerate(self):
        eng = self.engine
        if self.general:
            return eng.action_space_general(self.pos, self.cell, self.inv64, self.rep_ptr, self.site_ptr, self.sites, self.a,
                                            self.max_vac, max_actions=12 * self.max_vac)
        return eng.action_space(self.pos, self.cell, self.rep_ptr, self.a, self.max_vac, max_actions=12 * self.max_vac)

    def step(self, temperature: float, uniforms: Optional[np.ndarray] = None, apply: bool = True,
             rng: str = "numpy") -> Dict[str, torch.Tensor]:
        """One LSS step of every replica.  ``uniforms[R]`` are the inverse-cdf draws (``np.random.random_sample(R)``
        reproduces the reference's ``np.random.choice`` stream); default: a fresh draw from the global numpy RNG, or with
        ``rng="philox"`` the counter-based device generator keyed by (seed, replica id, step index)."""
        eng, dev = self.engine, self.device
        count, actions, status = self._enumerate()
        first, total = eng.candidate_offsets(count)
        head = torch.cat([total, status.max().reshape(1).to(torch.int32)]).cpu().numpy()     # the step's one D2H
        if head[1] != 0:
            raise AssertionError("Expected 1 vacancy" if head[1] == 2 and self.max_vac == 1 else
                                 f"rlvd_action_space status {int(head[1])}")
        G = int(head[0])
        b = eng.expand_candidates(self.rep_ptr, first, G, self.n_atoms, self.Z, self.pos, self.cell, actions, self.inv32, self.img)
        qp = eng.q_params(dict(self.q_params, temperature=float(temperature)), self.model.head_spec)
        out = eng.score_device(b["Z"], b["pos"], b["cell"], b["inv"], b["img"], b["ptr"], b["rs"], b["ps"], qp,
                               max_struct_nodes=self.n_atoms)
        T = torch.full((self.R,), float(temperature), dtype=torch.float32, device=dev)
        if uniforms is None and rng == "philox":
            chosen, amax, probs, gamma = eng.select_apply(self.rep_ptr, first, out["rl_q"], T, None, actions, self.pos, apply,
                                                          philox=(self.seed, self.step_index, self.replica_id0))
        else:
            u = np.random.random_sample(self.R) if uniforms is None else np.asarray(uniforms, dtype=np.float64)
            chosen, amax, probs, gamma = eng.select_apply(self.rep_ptr, first, out["rl_q"], T, torch.from_numpy(u).to(dev),
                                                          actions, self.pos, apply)
        self.step_index += 1
        return {"chosen": chosen, "argmax": amax, "probs": probs, "rl_q": out["rl_q"], "first": first, "actions": actions,
                "count": count, "gamma": gamma, "E_R": out["E_R"]}

    # ---- sync-free stepping: nothing below reads a device value on the host
    def step_async(self, temperature: float, n_cand: int, edge_capacity: int, logs: Optional[dict] = None, k: int = 0) -> dict:
        """One step queued on the stream with the host-side sizes fixed in advance: ``n_cand`` candidates in total (every
        vacancy of a perfect FCC lattice has 12 occupied neighbours, so it is ``R * 12 * n_vacancies`` unless vacancies
        touch) and at most ``edge_capacity`` edges.  A mismatch raises the device flag ``self.fault`` (checked by the
        caller at its next synchronisation point).  Philox draws; results are appended to ``logs`` at row ``k``."""
        eng, dev = self.engine, self.device
        count, actions, status = self._enumerate()
        first, total = eng.candidate_offsets(count)
        if getattr(self, "fault", None) is None:
            self.fault = torch.zeros(1, dtype=torch.int32, device=dev)
        self.fault += ((total != n_cand) | (status.max() != 0)).to(torch.int32)
        b = eng.expand_candidates(self.rep_ptr, first, n_cand, self.n_atoms, self.Z, self.pos, self.cell, actions, self.inv32, self.img)
        qp = eng.q_params(dict(self.q_params, temperature=float(temperature)), self.model.head_spec)
        out = eng.score_device(b["Z"], b["pos"], b["cell"], b["inv"], b["img"], b["ptr"], b["rs"], b["ps"], qp,
                               max_struct_nodes=self.n_atoms, edge_capacity=edge_capacity)
        T = self._T if getattr(self, "_T", None) is not None and self._T_value == float(temperature) else None
        if T is None:
            T = self._T = torch.full((self.R,), float(temperature), dtype=torch.float32, device=dev)
            self._T_value = float(temperature)
        chosen, amax, probs, gamma = eng.select_apply(self.rep_ptr, first, out["rl_q"], T, None, actions, self.pos, True,
                                                      philox=(self.seed, self.step_index, self.replica_id0))
        self.step_index += 1
        if logs is not None:
            logs["Q"][k, :n_cand] = out["rl_q"]
            logs["first"][k] = first
            logs["chosen"][k] = chosen
            logs["gamma"][k] = gamma
            if "coords" in logs:
                logs["coords"][k] = self.pos.view(self.R, self.n_atoms, 3)[:, -1, :]
            if "frames" in logs:
                logs["frames"][k] = self.pos.view(self.R, self.n_atoms, 3)
        return {"rl_q": out["rl_q"], "first": first, "chosen": chosen, "gamma": gamma, "E_R": out["E_R"], "actions": actions}

    def run(self, horizon: int, temperature: float, mode: str = "lss", writer=None,
            uniforms: Optional[np.ndarray] = None, sync_free: bool = False, flush_every: int = 64,
            on_frames=None) -> Dict[str, list]:
        """``horizon`` steps of every replica: the loop of ``RLSimulator.run_LSS`` / ``run_TKS``
        (``rlsim/drl/simulator.py:187-236, 239-283``) without relaxation or energy logging.  ``mode="tks"`` also keeps
        the residence-time clock ``t += 1e-6 / sum(exp(Q / kT))`` (simulator.py:262-265) and the tracked atom's
        coordinates (the LAST atom, simulator.py:241,271).  ``writer``: a ``TrajectoryWriter`` fed once per step.
        ``uniforms[horizon, R]`` pins the draws (``np.random``-compatible replay).  ``sync_free=True``: Philox draws, steps
        queued without any host synchronisation, device logs read back every ``flush_every`` steps (one D2H per flush
        instead of four per step, no per-step Python loop over replicas); ``on_frames(frames[k, R, N, 3] device tensor,
        first_step)`` is called at each flush (e.g. the t_net evaluation of config 4).
        Returns per-replica lists shaped like the reference's outputs."""
        if mode not in ("lss", "tks"):
            raise ValueError("DeviceLSS.run supports mode 'lss' and 'tks'")
        if mode == "tks":
            assert not self.q_params["dqn"], "TKS needs rates, not DQN Q-values (simulator.py:181)"
        if sync_free:
            if uniforms is not None:
                raise ValueError("sync_free runs draw with the device generator; pass uniforms only to the synchronous form")
            return self._run_sync_free(horizon, temperature, mode, writer, flush_every, on_frames)
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

    def _run_sync_free(self, horizon, temperature, mode, writer, flush_every, on_frames):
        import time
        R, dev = self.R, self.device
        # host-known sizes: one synchronous probe step's worth of information (candidate count, edge count)
        count, _, status = self._enumerate()
        probe = torch.cat([count.sum().reshape(1), status.max().reshape(1)]).cpu().numpy()
        if probe[1] != 0:
            raise AssertionError("Expected 1 vacancy" if probe[1] == 2 and self.max_vac == 1 else f"rlvd_action_space status {int(probe[1])}")
        n_cand = int(probe[0])
        edge_capacity = int(2 * n_cand * self.n_atoms * 64)              # 64 neighbours per atom: FCC has 54 inside 5.0 A
        A = 12 * self.max_vac
        K = max(1, min(int(flush_every), horizon))
        need_frames = writer is not None or on_frames is not None
        logs = {"Q": torch.full((K, R * A), float("nan"), dtype=torch.float32, device=dev),
                "first": torch.zeros((K, R + 1), dtype=torch.int32, device=dev),
                "chosen": torch.zeros((K, R), dtype=torch.int32, device=dev),
                "gamma": torch.zeros((K, R), dtype=torch.float32, device=dev)}
        if mode == "tks":
            logs["coords"] = torch.zeros((K, R, 3), dtype=torch.float64, device=dev)
        if need_frames:
            logs["frames"] = torch.zeros((K, R, self.n_atoms, 3), dtype=torch.float64, device=dev)
        q_lists = [[] for _ in range(R)]
        act_lists = [[] for _ in range(R)]
        t_arr = np.zeros((horizon + 1, R))
        c_lists = [[self.positions()[r, -1].tolist()] for r in range(R)] if mode == "tks" else None
        host_s = 0.0
        done = 0
        t_wall = time.perf_counter()
        while done < horizon:
            k_n = min(K, horizon - done)
            t0 = time.perf_counter()
            for k in range(k_n):
                self.step_async(temperature, n_cand, edge_capacity, logs, k)
            host_s += time.perf_counter() - t0                          # queueing only: no device value is read here
            # ---- one flush: the block's logs in a handful of D2H copies
            if on_frames is not None:
                on_frames(logs["frames"][:k_n], done)
            fault = int(self.fault.item()) + self.engine.edge_overflow() + (self.engine.check_status() & 3)
            if fault:
                raise RuntimeError("DeviceLSS sync-free run: the candidate count / edge capacity / operand range assumed for the "
                                   "queued steps did not hold (touching vacancies or an unphysical state); use the synchronous form")
            t0 = time.perf_counter()
            Q = logs["Q"][:k_n].cpu().numpy()
            first = logs["first"][:k_n].cpu().numpy()
            chosen = logs["chosen"][:k_n].cpu().numpy()
            gamma = logs["gamma"][:k_n].cpu().numpy().astype(np.float64)
            coords = logs["coords"][:k_n].cpu().numpy() if mode == "tks" else None
            frames = logs["frames"][:k_n].cpu().numpy() if writer is not None else None
            for k in range(k_n):
                f = first[k]
                for r in range(R):
                    q_lists[r].append(Q[k, f[r]:f[r + 1]].tolist())
                    act_lists[r].append(int(chosen[k, r]))
                if mode == "tks":
                    t_arr[done + k + 1] = t_arr[done + k] + 1.0 / gamma[k] * 10 ** -6
                    for r in range(R):
                        c_lists[r].append(coords[k, r].tolist())
                if writer is not None:
                    writer.append(frames[k], q_values=[Q[k, f[r]:f[r + 1]] for r in range(R)], chosen=chosen[k],
                                  times=t_arr[done + k + 1].tolist() if mode == "tks" else None)
            host_s += time.perf_counter() - t0
            done += k_n
        res = {"Q": q_lists, "actions": act_lists, "host_seconds": host_s, "wall_seconds": time.perf_counter() - t_wall}
        if mode == "tks":
            res["time"] = [t_arr[:, r].tolist() for r in range(R)]
            res["coords"] = c_lists
        return res
