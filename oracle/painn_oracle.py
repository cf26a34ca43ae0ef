"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  **PARITY UNPINNED** for the model arithmetic restated here.

(The parts of the path the reference DOES ship as source — `rlsim/actions/action.py`, `rlsim/utils/sro.py`, the softmax +
`np.random.choice` lines of `rlsim/drl/simulator.py:94-98` — are not restated in this file: they are pinned by fixtures minted
from the reference's own code, `scripts/make_ref_fixtures.py` -> `tests/golden/ref_*.npz`.)


This file restates, in eager PyTorch / numpy on the CPU, the arithmetic the reference reaches at
``rlsim/drl/simulator.py:56`` (``model(batch, q_params)["rl_q"]``).  It is the checker for the CUDA
path and the ``cpu_baseline`` leg of ``bench.py``; nothing under ``rlvacdiffsim_b200/`` imports it and
the product path never falls back to it.

Why "unpinned": the arithmetic is not in ``/root/reference``.  It lives in the un-vendored, un-pinned
dependency ``rgnn @ git+https://github.com/learningmatter-mit/ReactionGraphNeuralNetwork.git@main``
(``pyproject.toml:18``) plus ``torch_cluster`` / ``torch_scatter`` / ``torch_geometric``
(``pyproject.toml:19,22,23``), none of which is importable here, and the reference has no tests or
golden vectors.  What *is* pinned: every tensor name/shape/value of the bundled
checkpoint ``data/model_trained`` and the call-site conventions (``q0=-barrier``, ``q1=freq``
``trainer.py:81-82``; ``kb=8.617e-5`` ``simulator.py:40``; cutoff 5.0).  The representation follows the
published PaiNN equations (Schütt et al. 2021), whose parameter names are exactly the checkpoint's.
Every choice that cannot be checked is a named switch in :class:`HeadSpec` / ``ASSUMPTIONS``.

Canonical edge arithmetic (shared verbatim by the CUDA neighbor kernel so that edge sets are
bit-identical): see :func:`radius_graph_pbc`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

KB_EV = 8.617 * 10 ** -5  # rlsim/drl/simulator.py:40

ASSUMPTIONS = {
    "A1_edge_inclusion": "strict d^2 < rc^2, no neighbor cap (torch_cluster default cap of 32 would truncate 54 real neighbors)",
    "A2_edge_direction": "r_ij = pos_j - pos_i + S.cell, message j -> i (sign invisible to scalar outputs)",
    "A3_bessel_prefactor": "phi_k = sin(freqs_k d)/d with no sqrt(2/rc) prefactor",
    "A4_filter_split": "[dq | dmuR | dmumu] per layer slice of 384",
    "A5_heads": "see HeadSpec (HEAD_SPEC_V0)",
}


@dataclass(frozen=True)
class HeadSpec:
    """Named, swappable assumptions for the parts of rgnn's heads the checkpoint does not pin
    (SURVEY.md Appendix C.7).  The CUDA readout receives the same switches as integers."""
    node_diff: str = "p_minus_r"          # input of reaction_representation: q_P - q_R
    delta_e_input: str = "raw"            # 33rd input of reaction_output / Q_emb: delta_e in eV ("raw") or standardised ("std")
    painn_head_activation: str = "silu"   # energy_output / reaction_representation / reaction_output
    dqn_head_activation: str = "silu"     # Q_emb / Q_output
    q_extra: Tuple[str, str] = ("kT_x10", "delta_e")  # the two extra Q_output inputs (18 = 16 + 2)
    name: str = "HEAD_SPEC_V0"


HEAD_SPEC_V0 = HeadSpec()


@dataclass(frozen=True)
class TimeHeadSpec:
    """Named assumptions for rgnn's ``t_net`` head — nothing in the reference pins them: no
    t_net checkpoint is bundled.  Call sites fix only the interface: ``model(batch, inference=True)["time"]`` with
    per-frame ``T`` and ``defect`` (``rlsim/time/misc.py:115-141``), ``t_real = -tau * log(1 - t / tau)``
    (``estimate_time.py:15-35``), constructor keywords ``n_feat, pooling, tau0, D_scaler, T_scaler_m``."""
    pooling: str = "mean"                 # mean of the node scalars q over the frame's atoms
    T_input: str = "T / T_scaler_m"
    D_input: str = "defect / D_scaler"
    activation: str = "silu"
    squash: str = "tau * sigmoid"         # time in (0, tau)
    # t_net_binary (rlsim/time/train.py:128, misc.py:205: binary_cross_entropy_with_logits(prediction["goal"], goal_labels)):
    goal_head: str = "a second MLP of the same shape on the same pooled input; 'goal' is its raw output (a logit)"
    name: str = "TIME_HEAD_SPEC_V0"


TIME_HEAD_SPEC_V0 = TimeHeadSpec()

_ACT = {"silu": F.silu, "relu": F.relu, "tanh": torch.tanh,
        "leaky_relu": lambda x: F.leaky_relu(x, 0.01)}


# ----------------------------------------------------------------------------- K1: radius graph

def inverse_cell_f32(cell: np.ndarray) -> np.ndarray:
    """fp32 image of the fp64 inverse; both the oracle and the CUDA path receive THIS matrix."""
    return np.linalg.inv(np.asarray(cell, dtype=np.float64)).astype(np.float32)


def image_range(cell: np.ndarray, cutoff: float) -> np.ndarray:
    """Per-axis number of extra periodic images to scan around the rounded one.

    After ``S = -rint(f)`` the residual fractional coordinate is within 0.5 of zero; any image closer
    than ``rc`` has ``|f_k| <= rc/h_k`` (h_k = cell height), so offsets up to ``floor(rc/h_k + 0.5)``
    are enough.  All BASELINE cells have ``h_k >= 2 rc`` → 0 → single minimum image."""
    cell = np.asarray(cell, dtype=np.float64)
    vol = abs(np.linalg.det(cell))
    cross = np.stack([np.cross(cell[1], cell[2]), np.cross(cell[2], cell[0]), np.cross(cell[0], cell[1])])
    h = vol / np.linalg.norm(cross, axis=1)
    return np.floor(cutoff / h + 0.5 + 1e-9).astype(np.int32)


def radius_graph_pbc(pos: np.ndarray, cell: np.ndarray, cutoff: float = 5.0
                     ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Periodic radius graph of ONE structure in the canonical fp32 arithmetic.

    For every ordered pair (i, j)::

        d   = pos[j] - pos[i]                                   (fp32)
        f_k = (d.x*inv[0,k] + d.y*inv[1,k]) + d.z*inv[2,k]      (each op rounded, no FMA)
        S   = -rint(f) + o,   o in [-R..R]^3                     (R = image_range, 0 for all BASELINE cells)
        r_c = ((d_c + S0*cell[0,c]) + S1*cell[1,c]) + S2*cell[2,c]
        keep iff (r.x*r.x + r.y*r.y) + r.z*r.z < fp32(rc)*fp32(rc)  and not (i == j and S == 0)

    Returns ``(recv i, send j, shift S)`` sorted by (i, j, S0, S1, S2): receiver-sorted CSR order."""
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    cell32 = np.ascontiguousarray(cell, dtype=np.float32)
    inv = inverse_cell_f32(cell)
    R = image_range(cell, cutoff)
    rc = np.float32(cutoff)
    rc2 = np.float32(rc * rc)
    n = len(pos)
    d = pos[None, :, :] - pos[:, None, :]                                  # [i, j, 3]
    dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
    base = []
    for k in range(3):
        f = (dx * inv[0, k] + dy * inv[1, k]) + dz * inv[2, k]
        base.append(-np.rint(f))
    recv, send, shifts = [], [], []
    eye = np.eye(n, dtype=bool)
    for o0 in range(-R[0], R[0] + 1):
        for o1 in range(-R[1], R[1] + 1):
            for o2 in range(-R[2], R[2] + 1):
                S0 = (base[0] + np.float32(o0)).astype(np.float32)
                S1 = (base[1] + np.float32(o1)).astype(np.float32)
                S2 = (base[2] + np.float32(o2)).astype(np.float32)
                r = []
                for c, dc in enumerate((dx, dy, dz)):
                    r.append(((dc + S0 * cell32[0, c]) + S1 * cell32[1, c]) + S2 * cell32[2, c])
                d2 = (r[0] * r[0] + r[1] * r[1]) + r[2] * r[2]
                keep = d2 < rc2
                keep &= ~(eye & (S0 == 0) & (S1 == 0) & (S2 == 0))
                ii, jj = np.nonzero(keep)
                recv.append(ii)
                send.append(jj)
                shifts.append(np.stack([S0[ii, jj], S1[ii, jj], S2[ii, jj]], axis=1).astype(np.int32))
    recv = np.concatenate(recv).astype(np.int64)
    send = np.concatenate(send).astype(np.int64)
    shifts = np.concatenate(shifts).reshape(-1, 3)
    order = np.lexsort((shifts[:, 2], shifts[:, 1], shifts[:, 0], send, recv))
    return recv[order], send[order], shifts[order].astype(np.int8)


def batched_radius_graph(pos: np.ndarray, cells: np.ndarray, node_ptr: np.ndarray, cutoff: float = 5.0):
    """Disjoint-union radius graph → (row_ptr[M+1] i64, col[E] i64 global sender index, shift[E,3] i8)."""
    M = int(node_ptr[-1])
    deg = np.zeros(M, dtype=np.int64)
    cols, shifts = [], []
    for g in range(len(node_ptr) - 1):
        a, b = int(node_ptr[g]), int(node_ptr[g + 1])
        ri, sj, S = radius_graph_pbc(pos[a:b], cells[g], cutoff)
        np.add.at(deg, ri + a, 1)
        cols.append(sj + a)
        shifts.append(S)
    row_ptr = np.zeros(M + 1, dtype=np.int64)
    np.cumsum(deg, out=row_ptr[1:])
    col = np.concatenate(cols) if cols else np.zeros(0, dtype=np.int64)
    shift = np.concatenate(shifts) if shifts else np.zeros((0, 3), dtype=np.int8)
    return row_ptr, col, shift


# ----------------------------------------------------------------------------- the model

class PaiNNOracle:
    """dqn_v2 / painn restated from the checkpoint tensor map."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], hyper_parameters: dict,
                 dtype: torch.dtype = torch.float32, head_spec: HeadSpec = HEAD_SPEC_V0):
        self.dtype = dtype
        self.spec = head_spec
        self.hp = hyper_parameters
        rm = hyper_parameters["reaction_model"]
        self.cutoff = float(rm["cutoff"])
        self.F = int(rm["hidden_channels"])
        self.L = int(rm["n_interactions"])
        self.eps = float(rm["epsilon"])
        self.means = {k: float(v) for k, v in rm["means"].items()}
        self.stddevs = {k: float(v) for k, v in rm["stddevs"].items()}
        self.w = {k: (v.to(dtype) if v.is_floating_point() else v.clone()) for k, v in state_dict.items()}

    @classmethod
    def load(cls, path: str, **kw) -> "PaiNNOracle":
        ck = torch.load(path, weights_only=True, map_location="cpu")
        return cls(ck["state_dict"], ck["hyper_parameters"], **kw)

    # -- small helpers
    def _lin(self, x, prefix):
        return F.linear(x, self.w[prefix + ".weight"], self.w.get(prefix + ".bias"))

    def _mlp(self, x, prefix, idx: Sequence[int], act):
        for n, i in enumerate(idx):
            x = self._lin(x, f"{prefix}.layers.{i}")
            if n + 1 < len(idx):
                x = act(x)
        return x

    # -- K2..K4, R2..R6
    def representation(self, Z: np.ndarray, pos: np.ndarray, cells: np.ndarray, node_ptr: np.ndarray,
                       edges=None, return_mu: bool = False):
        """PaiNN representation on a disjoint-union batch → q[M, F] (and mu[M, 3, F])."""
        dt = self.dtype
        if edges is None:
            edges = batched_radius_graph(pos, cells, node_ptr, self.cutoff)
        row_ptr, col, shift = edges
        M = int(node_ptr[-1])
        deg = np.diff(row_ptr)
        idx_i = torch.from_numpy(np.repeat(np.arange(M, dtype=np.int64), deg))
        idx_j = torch.from_numpy(np.asarray(col, dtype=np.int64))
        graph_of_node = np.repeat(np.arange(len(node_ptr) - 1), np.diff(node_ptr))
        P = torch.from_numpy(np.asarray(pos, dtype=np.float32)).to(dt)
        C = torch.from_numpy(np.asarray(cells, dtype=np.float32)).to(dt)          # [G,3,3]
        S = torch.from_numpy(np.asarray(shift, dtype=np.float32)).to(dt)          # [E,3]
        g_of_edge = torch.from_numpy(graph_of_node)[idx_i]
        r = P[idx_j] - P[idx_i] + torch.einsum("ek,ekc->ec", S, C[g_of_edge])      # C.1
        d = r.norm(dim=1, keepdim=True)                                            # C.2
        dirs = r / d
        freqs = self.w["reaction_model.representation.radial_basis.freqs"]
        phi = torch.sin(freqs[None, :] * d) / d
        fcut = 0.5 * (torch.cos(d * (math.pi / self.cutoff)) + 1.0) * (d < self.cutoff).to(dt)
        Wf = self._lin(phi, "reaction_model.representation.filter_net") * fcut     # C.3  [E, L*3F]
        Fc = self.F
        q = self.w["reaction_model.representation.embedding.weight"][torch.from_numpy(np.asarray(Z, dtype=np.int64))]
        mu = torch.zeros(M, 3, Fc, dtype=dt)                                       # C.4
        for l in range(self.L):
            pre = f"reaction_model.representation.interactions.{l}.interatomic_context_net"
            x = self._mlp(q, pre, (0, 3), F.silu)                                  # C.5 [M,3F]
            m = Wf[:, 3 * Fc * l: 3 * Fc * (l + 1)] * x[idx_j]
            dq, dmuR, dmumu = m[:, :Fc], m[:, Fc:2 * Fc], m[:, 2 * Fc:]
            dmu = dmuR[:, None, :] * dirs[:, :, None] + dmumu[:, None, :] * mu[idx_j]
            q = q.index_add(0, idx_i, dq)
            mu = mu.index_add(0, idx_i, dmu)
            pre = f"reaction_model.representation.mixing.{l}"
            mix = F.linear(mu, self.w[pre + ".mu_channel_mix.weight"])              # C.6 [M,3,2F]
            mu_V, mu_W = mix[..., :Fc], mix[..., Fc:]
            nrm = torch.sqrt((mu_V ** 2).sum(dim=1) + self.eps)
            ctx = self._mlp(torch.cat([q, nrm], dim=1), pre + ".intraatomic_context_net", (0, 3), F.silu)
            a, b, c = ctx[:, :Fc], ctx[:, Fc:2 * Fc], ctx[:, 2 * Fc:]
            q = q + a + c * (mu_V * mu_W).sum(dim=1)
            mu = mu + b[:, None, :] * mu_W
        return (q, mu) if return_mu else q

    # -- T1: t_net head (TimeHeadSpec)
    def time_forward(self, Z, pos, cells, node_ptr, T, defect, head: Dict[str, np.ndarray], tau: float = 1.0,
                     T_scaler_m: float = 7918.951990135471, D_scaler: float = 1 / 256, pooling: str = "mean", logit: bool = False):
        """→ time[G]; ``head`` = {"w0","b0","w1","b1","w2","b2"} in nn.Linear layout.  ``logit=True`` returns the head's raw
        output instead (t_net_binary's ``"goal"``, TimeHeadSpec.goal_head)."""
        dt = self.dtype
        q = self.representation(Z, pos, cells, node_ptr)
        out = []
        W = {k: torch.from_numpy(np.asarray(v, dtype=np.float64)).to(dt) for k, v in head.items()}
        for g in range(len(node_ptr) - 1):
            qs = q[int(node_ptr[g]):int(node_ptr[g + 1])]
            pooled = qs.mean(dim=0) if pooling == "mean" else qs.sum(dim=0)
            z = torch.cat([pooled, torch.tensor([float(T[g]) / T_scaler_m, float(defect[g]) / D_scaler], dtype=dt)])
            h = F.silu(W["w0"] @ z + W["b0"])
            h = F.silu(W["w1"] @ h + W["b1"])
            o = W["w2"] @ h + W["b2"]
            out.append(o if logit else tau * torch.sigmoid(o))
        return torch.cat(out)

    # -- H1..H4
    def atom_energies(self, q, Z):
        act = _ACT[self.spec.painn_head_activation]
        out = self._mlp(q, "reaction_model.energy_output", (0, 3), act)[:, 0]
        look = self.w["reaction_model.species_energy_scale.elem_lookup"][torch.from_numpy(np.asarray(Z, dtype=np.int64))]
        return out * self.w["reaction_model.species_energy_scale.scales"][look] + \
            self.w["reaction_model.species_energy_scale.shifts"][look]

    def reaction_forward(self, Z, pos_R, pos_P, cells, node_ptr, q_params: dict,
                         edges_R=None, edges_P=None) -> Dict[str, torch.Tensor]:
        """One disjoint-union batch of reaction graphs → rl_q[G], q0, q1, delta_e, E_R, E_P."""
        dt = self.dtype
        G = len(node_ptr) - 1
        q_R = self.representation(Z, pos_R, cells, node_ptr, edges_R)
        q_P = self.representation(Z, pos_P, cells, node_ptr, edges_P)
        return self.heads(q_R, q_P, Z, node_ptr, q_params)

    def heads(self, q_R, q_P, Z, node_ptr, q_params: dict) -> Dict[str, torch.Tensor]:
        dt = self.dtype
        G = len(node_ptr) - 1
        batch = torch.from_numpy(np.repeat(np.arange(G, dtype=np.int64), np.diff(node_ptr)))
        e_R, e_P = self.atom_energies(q_R, Z), self.atom_energies(q_P, Z)
        E_R = torch.zeros(G, dtype=dt).index_add(0, batch, e_R)
        E_P = torch.zeros(G, dtype=dt).index_add(0, batch, e_P)
        delta_e = torch.zeros(G, dtype=dt).index_add(0, batch, e_P - e_R)           # per-atom difference, then pool
        p_act = _ACT[self.spec.painn_head_activation]
        if self.spec.node_diff != "p_minus_r":
            raise NotImplementedError(self.spec.node_diff)
        h_atom = self._mlp(q_P - q_R, "reaction_model.reaction_representation", (0, 3), p_act)
        h = torch.zeros(G, h_atom.shape[1], dtype=dt).index_add(0, batch, h_atom)  # pool_method='sum'
        if self.spec.delta_e_input == "raw":
            de_in = delta_e
        else:
            de_in = (delta_e - self.means["delta_e"]) / self.stddevs["delta_e"]
        z = torch.cat([h, de_in[:, None]], dim=1)                                   # [G,33]
        bf = self._mlp(z, "reaction_model.reaction_output", (0, 3, 6), p_act)
        barrier = bf[:, 0] * self.stddevs["barrier"] + self.means["barrier"]
        freq = bf[:, 1] * self.stddevs["freq"] + self.means["freq"]
        q0, q1 = -barrier, freq                                                     # trainer.py:81-82
        alpha, beta = float(q_params["alpha"]), float(q_params["beta"])
        kT = float(q_params.get("temperature", 0.0)) * KB_EV
        rl_q = alpha * (q0 + kT * q1) + beta * (-delta_e)
        q_corr = torch.zeros(G, dtype=dt)
        if bool(q_params.get("dqn", False)):
            d_act = _ACT[self.spec.dqn_head_activation]
            emb = self._mlp(z, "Q_emb", (0, 3), d_act)
            extras = []
            for name in self.spec.q_extra:
                if name == "kT_x10":
                    extras.append(torch.full((G, 1), kT * 10.0, dtype=dt))
                elif name == "delta_e":
                    extras.append(delta_e[:, None])
                else:
                    raise NotImplementedError(name)
            q_corr = self._mlp(torch.cat([emb] + extras, dim=1), "Q_output", (0, 3, 6), d_act)[:, 0]
            rl_q = rl_q + q_corr
        return {"rl_q": rl_q, "q0": q0, "q1": q1, "barrier": barrier, "freq": freq,
                "delta_e": delta_e, "E_R": E_R, "E_P": E_P, "q_corr": q_corr}


# ----------------------------------------------------------------------------- A2/A3/A5/A6 around the model

def build_reaction_batch(numbers, positions, cell, actions):
    """``convert_to_graph_list`` + PyG collate (``simulator.py:439-460, 50-55``) as flat arrays."""
    numbers = np.asarray(numbers, dtype=np.int64)
    positions = np.asarray(positions, dtype=np.float64)
    n, A = len(numbers), len(actions)
    Z = np.tile(numbers, A)
    pos_R = np.tile(positions, (A, 1))
    pos_P = pos_R.copy()
    for a, act in enumerate(actions):
        pos_P[a * n + int(act[0])] += np.asarray(act[1:4], dtype=np.float64)
    cells = np.tile(np.asarray(cell, dtype=np.float64)[None], (A, 1, 1))
    node_ptr = np.arange(A + 1, dtype=np.int64) * n
    return Z, pos_R.astype(np.float32), pos_P.astype(np.float32), cells.astype(np.float32), node_ptr


def score_actions(model: PaiNNOracle, numbers, positions, cell, actions, q_params: dict,
                  batch_size: int = 16) -> Dict[str, torch.Tensor]:
    """``RLSimulator.select_action`` lines 49-58: candidates in DataLoader batches of 16."""
    outs: List[Dict[str, torch.Tensor]] = []
    for s in range(0, len(actions), batch_size):
        Z, pR, pP, cells, ptr = build_reaction_batch(numbers, positions, cell, actions[s:s + batch_size])
        with torch.no_grad():
            outs.append(model.reaction_forward(Z, pR, pP, cells, ptr, q_params))
    return {k: torch.cat([o[k] for o in outs], dim=-1) for k in outs[0]}


def action_probabilities(Q: torch.Tensor, temperature: float) -> torch.Tensor:
    """``simulator.py:94``: softmax(Q / (T * kb)) in the dtype of Q."""
    return torch.softmax(Q / (temperature * KB_EV), dim=0)


def sample_action(probs: torch.Tensor, seed: Optional[int] = None) -> int:
    """``simulator.py:95-98`` with the global numpy RNG optionally seeded for parity runs."""
    if seed is not None:
        np.random.seed(seed)
    p = probs.detach().cpu().numpy()
    return int(np.random.choice(len(p), p=p))


def tks_time_step(Q: torch.Tensor, temperature: float) -> float:
    """``simulator.py:262-263``: dt = 1e-6 / sum(exp(Q/kT))."""
    kT = temperature * 8.617 * 10 ** -5
    gamma = float(torch.sum(torch.exp(Q / kT)))
    return 1 / gamma * 10 ** -6
