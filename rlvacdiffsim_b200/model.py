"""Host-side mirror of rgnn's ``dqn_v2`` / ``painn`` / ``t_net`` model classes.

The classes are ``nn.Module`` parameter holders whose ``state_dict`` keys are exactly those of the
bundled checkpoint ``data/model_trained``, so ``.load(path)`` / ``.save(path)`` /
``.to`` / ``.eval`` / ``.state_dict`` / ``.named_parameters`` behave as the reference's callers expect
(``rlsim/drl/deploy.py:18``, ``train.py:48-58``, ``trainer.py:36-39,120``).  ``forward`` does no torch
arithmetic: it hands device pointers to the CUDA engine (``include/rlvd_b200.h``).  There is no eager
fallback — on a machine without the built library or without a CUDA device ``forward`` raises.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from .engine import HEAD_SPEC_V0, Engine, HeadSpec, cell_geometry

_ACTS = {"silu": nn.SiLU, "relu": nn.ReLU, "tanh": nn.Tanh, "leaky_relu": nn.LeakyReLU}


class MLP(nn.Module):
    """``layers = [Linear, act, Dropout] * n_hidden + [Linear]`` → checkpoint keys ``layers.{0,3,6}``."""

    def __init__(self, n_input: int, n_output: int, hidden_layers, activation: str = "silu",
                 dropout_rate: float = 0.0):
        super().__init__()
        mods: List[nn.Module] = []
        d = n_input
        for h in hidden_layers:
            mods += [nn.Linear(d, h), _ACTS[activation](), nn.Dropout(dropout_rate)]
            d = h
        mods.append(nn.Linear(d, n_output))
        self.layers = nn.Sequential(*mods)


class _Buffer(nn.Module):
    def __init__(self, name: str, value: torch.Tensor):
        super().__init__()
        self.register_buffer(name, value)


class _Interaction(nn.Module):
    def __init__(self, F: int):
        super().__init__()
        self.interatomic_context_net = MLP(F, 3 * F, (F,), "silu")


class _Mixing(nn.Module):
    def __init__(self, F: int):
        super().__init__()
        self.intraatomic_context_net = MLP(2 * F, 3 * F, (F,), "silu")
        self.mu_channel_mix = nn.Linear(F, 2 * F, bias=False)


class PaiNNRepresentation(nn.Module):
    def __init__(self, hidden_channels=128, n_interactions=3, n_rbf=20, cutoff=5.0, max_z=100):
        super().__init__()
        F = hidden_channels
        self.cutoff_fn = _Buffer("cutoff", torch.tensor(float(cutoff)))
        self.radial_basis = _Buffer("freqs", torch.arange(1, n_rbf + 1, dtype=torch.float32) * (np.pi / cutoff))
        self.embedding = nn.Embedding(max_z, F)
        self.filter_net = nn.Linear(n_rbf, n_interactions * 3 * F)
        self.interactions = nn.ModuleList([_Interaction(F) for _ in range(n_interactions)])
        self.mixing = nn.ModuleList([_Mixing(F) for _ in range(n_interactions)])


class SpeciesEnergyScale(nn.Module):
    def __init__(self, species):
        super().__init__()
        from .structures import SYMBOL_TO_Z
        self.scales = nn.Parameter(torch.ones(len(species)))
        self.shifts = nn.Parameter(torch.zeros(len(species)))
        look = torch.zeros(100, dtype=torch.int64)
        for i, s in enumerate(species):
            look[SYMBOL_TO_Z[s]] = i
        self.register_buffer("elem_lookup", look)


class PaiNN(nn.Module):
    """``registry.get_reaction_model_class("painn")`` — the reaction model inside ``dqn_v2``."""

    def __init__(self, species=("Cr", "Co", "Ni"), cutoff=5.0, hidden_channels=128, reaction_feat=32,
                 n_interactions=3, rbf_type="bessel", n_rbf=20, trainable_rbf=False, activation="silu",
                 shared_interactions=False, shared_filters=False, epsilon=1e-8, means=None, stddevs=None,
                 pool_method="sum", **_ignored):
        super().__init__()
        if (hidden_channels, n_rbf, n_interactions) != (128, 20, 3) or rbf_type != "bessel" or pool_method != "sum":
            raise NotImplementedError("the sm_100a kernels are specialised for painn F=128, n_rbf=20, L=3, bessel, sum")
        self.hyper = dict(species=list(species), cutoff=cutoff, hidden_channels=hidden_channels,
                          reaction_feat=reaction_feat, n_interactions=n_interactions, rbf_type=rbf_type, n_rbf=n_rbf,
                          trainable_rbf=trainable_rbf, activation=activation, shared_interactions=shared_interactions,
                          shared_filters=shared_filters, epsilon=epsilon,
                          means=means or {"barrier": 0.0, "freq": 0.0, "delta_e": 0.0},
                          stddevs=stddevs or {"barrier": 1.0, "freq": 1.0, "delta_e": 1.0}, pool_method=pool_method)
        self.hyper["@name"] = "painn"
        self.cutoff = float(cutoff)
        self.reaction_feat = reaction_feat
        F = hidden_channels
        self.species_energy_scale = SpeciesEnergyScale(species)
        self.representation = PaiNNRepresentation(F, n_interactions, n_rbf, cutoff)
        self.energy_output = MLP(F, 1, (F // 2,), activation)
        self.reaction_representation = MLP(F, reaction_feat, (F,), activation)
        self.reaction_output = MLP(reaction_feat + 1, 2, (reaction_feat, reaction_feat), activation)

    @classmethod
    def load(cls, path: str) -> "PaiNN":
        ck = torch.load(path, weights_only=True, map_location="cpu")
        hp = dict(ck["hyper_parameters"])
        hp.pop("@name", None)
        model = cls(**hp)
        model.load_state_dict(ck["state_dict"])
        return model

    def save(self, path: str) -> None:
        torch.save({"state_dict": self.state_dict(), "hyper_parameters": _plain(self.hyper)}, path)


def _plain(obj):
    if isinstance(obj, dict):
        return {k: _plain(v) for k, v in obj.items()}
    if torch.is_tensor(obj):
        return obj.detach().cpu()
    return obj


class _EngineMixin:
    """Keeps one CUDA :class:`Engine` per device in sync with the module's parameters."""

    head_spec: HeadSpec = HEAD_SPEC_V0

    def __getstate__(self):
        """``copy.deepcopy(model)`` / ``torch.save(model)`` after a forward: the per-device engines hold ctypes handles
        (not picklable, and one ``rlvd_engine*`` must not get two owners), so they are dropped and rebuilt lazily."""
        state = self.__dict__.copy()
        state.pop("_engines", None)
        return state

    def _engine_for(self, device: torch.device) -> Engine:
        engines = self.__dict__.setdefault("_engines", {})
        key = (device.type, device.index)
        eng = engines.get(key)
        sig = tuple((p.data_ptr(), p._version) for p in self._engine_tensors())
        if eng is None:
            eng = Engine(device)
            engines[key] = eng
            eng._sig = None
        if eng._sig != sig:
            sd, hp = self._engine_state()
            eng.load_state_dict(sd, hp)
            eng._sig = sig
        return eng


class DQNv2(nn.Module, _EngineMixin):
    """``registry.get_model_class("dqn_v2")`` — ``model(batch, q_params)["rl_q"]`` (``simulator.py:56``)."""

    def __init__(self, reaction_model: PaiNN, N_emb: int = 16, N_feat: int = 32, dropout_rate: float = 0.15,
                 canonical: bool = True, **_ignored):
        super().__init__()
        self.reaction_model = reaction_model
        self.hyper = {"@name": "dqn_v2", "N_emb": N_emb, "N_feat": N_feat, "dropout_rate": dropout_rate,
                      "canonical": canonical}
        rf = reaction_model.reaction_feat
        self.Q_emb = MLP(rf + 1, N_emb, (N_emb,), "silu", dropout_rate)
        self.Q_output = MLP(N_emb + 2, 1, (N_feat, N_feat), "silu", dropout_rate)
        if (N_emb, N_feat, rf) != (16, 32, 32):
            raise NotImplementedError("the readout kernel is specialised for N_emb=16, N_feat=32, reaction_feat=32")
        self.kb = 8.617 * 10 ** -5
        self.dedup_reactants = False   # share one representation pass between candidates of the same state

    # ---- checkpoint I/O: {"state_dict", "hyper_parameters"} (SURVEY.md §5 checkpoint/resume)
    @classmethod
    def load(cls, path: str) -> "DQNv2":
        ck = torch.load(path, weights_only=True, map_location="cpu")
        hp = dict(ck["hyper_parameters"])
        hp.pop("@name", None)
        rm_hp = dict(hp.pop("reaction_model"))
        rm_hp.pop("@name", None)
        model = cls(PaiNN(**rm_hp), **hp)
        model.load_state_dict(ck["state_dict"])
        return model

    def save(self, path: str) -> None:
        hp = dict(self.hyper)
        hp["reaction_model"] = _plain(self.reaction_model.hyper)
        torch.save({"state_dict": self.state_dict(), "hyper_parameters": hp}, path)

    # ---- engine plumbing
    def _engine_tensors(self):
        return list(self.parameters()) + list(self.buffers())

    def _engine_state(self):
        hp = dict(self.hyper)
        hp["reaction_model"] = self.reaction_model.hyper
        return self.state_dict(), hp

    @property
    def cutoff(self) -> float:
        return self.reaction_model.cutoff

    def forward(self, batch, q_params: Dict[str, float | bool]) -> Dict[str, torch.Tensor]:
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError(
                "dqn_v2.forward builds no autograd graph.  Replay-minibatch training (SURVEY.md §8a row B1) exists for the "
                "train_all=False form of rlsim/drl/trainer.py (reaction model frozen): use rlvacdiffsim_b200.trainer.Trainer "
                "or DQNv2.q_head_update; backward through the PaiNN representation has no sm_100a kernels yet.  Call under "
                "torch.no_grad() / .eval() for scoring")
        return self._score(batch, q_params)

    # ---- B1, train_all = False: one minibatch update of Q_emb / Q_output on the device
    Q_HEAD_PREFIXES = ("Q_emb.", "Q_output.")

    def q_head_update(self, batch, q_params: Dict[str, float | bool], target: torch.Tensor, state: Optional[dict], lr: float,
                      max_norm: float = 0.8, dropout_uniform: Optional[torch.Tensor] = None, apply: bool = True) -> dict:
        """``q_pred = self(batch, q_params)["rl_q"]; loss = mean((target - q_pred)**2); loss.backward();
        clip_grad_norm_(parameters, max_norm); Adam.step()`` of ``trainer.py:166-175`` for a model whose reaction model is
        frozen (``trainer.py:36-39``).  ``state``: the Adam state from ``new_q_head_state()`` (kept on the batch's device).
        ``dropout_uniform[G, 80]`` drives the heads' train-mode dropout (p = ``hyper['dropout_rate']``); None = identity.
        Returns ``{"loss", "pred", "grad_norm", "grad", "state"}``; the module's Q-head parameters are refreshed from the
        device when ``apply``."""
        if not q_params.get("dqn", False):
            raise ValueError("q_head_update needs q_params['dqn'] = True (the DQN heads do not enter rl_q otherwise)")
        dev = batch["pos"].device
        eng = self._engine_for(dev)
        if state is None:
            state = eng.new_q_head_state()
        out = self._score(batch, q_params, head_inputs=True)
        spec = self.head_spec
        from .engine import ACT_CODES
        p_drop = float(self.hyper.get("dropout_rate", 0.0)) if dropout_uniform is not None else 0.0
        res = eng.q_head_train_step(out["head_in"], target.to(dev, torch.float32), state, lr, ACT_CODES[spec.dqn_head_activation],
                                    max_norm=max_norm, dropout_uniform=dropout_uniform, dropout_p=p_drop, apply=apply)
        if apply:
            self._pull_q_heads(eng)
        res["state"] = state
        res["frozen_rl_q"] = out["rl_q"]          # eval-mode prediction with the pre-update weights
        return res

    # ---- B1, train_all = True: forward + backward of the whole model on a flat parameter vector (csrc/train_full.cu)
    def new_train_state(self, device) -> dict:
        """Flat fp32 parameters (checkpoint order), gradient bucket, Adam moments and the trainable mask
        (``requires_grad`` of the module's parameters: ``Trainer(train_all=False)`` freezes ``reaction_model.*``)."""
        dev = torch.device(device) if not isinstance(device, torch.device) else device
        if dev.type == "cuda" and dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        eng = self._engine_for(dev)
        params = eng.flatten_state(self.state_dict())
        req = {n: p.requires_grad for n, p in self.named_parameters()}
        mask = torch.zeros(params.numel(), dtype=torch.uint8)
        for name, off, n, trainable in eng.train_layout():
            if trainable and req.get(name, False):
                mask[off:off + n] = 1
        z = lambda: torch.zeros_like(params)
        return {"params": params, "grads": z(), "exp_avg": z(), "exp_avg_sq": z(), "step": 0, "mask": mask.to(dev),
                "device": dev}

    def train_forward_backward(self, batch, q_params: Dict[str, float | bool], state: dict, target0: torch.Tensor,
                               target1: Optional[torch.Tensor] = None, mode: str = "dqn",
                               dropout_uniform: Optional[torch.Tensor] = None, return_q: bool = False) -> dict:
        """``output = self(batch, q_params); loss = ...; loss.backward()`` of ``trainer.py:89-92,166-172`` with the
        parameters of ``state`` (not the module's): fills ``state["grads"]`` (unclipped) and returns loss / predictions.
        ``mode="dqn"``: ``mean((target0 - rl_q)**2)``; ``mode="bandit"``: ``mse(q0, target0) + mse(q1, target1)``."""
        prep = self.prepare_train_batch(batch, state["device"])
        return self.train_forward_backward_prepared(prep, q_params, state, target0, target1, mode, dropout_uniform, return_q)

    def prepare_train_batch(self, batch, device) -> dict:
        """Host side of a training minibatch: the structure arrays in the layout ``rlvd_train_forward_backward`` expects
        (reactants, then products), cell geometry, pointers — everything that needs a host value.  What remains
        (:meth:`train_forward_backward_prepared`) only queues device work, so a fixed-shape minibatch can be captured in a
        CUDA graph."""
        dev = torch.device(device) if not isinstance(device, torch.device) else device
        eng = self._engine_for(dev)
        G = batch.num_graphs
        pos = batch["pos"].to(dev, torch.float32)
        n_atoms = np.diff(batch["ptr"].cpu().numpy())
        cells64 = np.asarray(batch["cell64"], dtype=np.float64)
        counts = np.concatenate([n_atoms, n_atoms])
        node_ptr_host = np.zeros(len(counts) + 1, dtype=np.int32)
        np.cumsum(counts, out=node_ptr_host[1:])
        inv, img = cell_geometry(np.concatenate([cells64, cells64]), eng.cutoff)
        return {"G": G, "pos": torch.cat([pos, batch["pos_product"].to(dev, torch.float32)], dim=0).contiguous(),
                "Z": torch.cat([batch["elems"], batch["elems"]]).to(dev, torch.int32).contiguous(),
                "cell": torch.cat([batch["cell"], batch["cell"]], dim=0).to(dev, torch.float32).contiguous(),
                "ptr": torch.from_numpy(node_ptr_host).to(dev), "inv": torch.from_numpy(inv).to(dev),
                "img": torch.from_numpy(img).to(dev), "n_nodes": int(node_ptr_host[-1])}

    def train_forward_backward_prepared(self, prep: dict, q_params, state: dict, target0: torch.Tensor,
                                        target1: Optional[torch.Tensor] = None, mode: str = "dqn",
                                        dropout_uniform: Optional[torch.Tensor] = None, return_q: bool = False,
                                        edge_capacity: Optional[int] = None) -> dict:
        dev = state["device"]
        eng = self._engine_for(dev)
        row_ptr, col, shift, node_struct = eng.radius_graph(prep["pos"], prep["cell"], prep["inv"], prep["img"], prep["ptr"],
                                                            edge_capacity=edge_capacity)
        qp = eng.q_params(q_params, self.head_spec)
        p_drop = float(self.hyper.get("dropout_rate", 0.0)) if dropout_uniform is not None else 0.0
        t0 = target0.to(dev, torch.float32).reshape(-1).contiguous()
        t1 = None if target1 is None else target1.to(dev, torch.float32).reshape(-1).contiguous()
        return eng.train_forward_backward(prep["Z"], prep["pos"], prep["cell"], prep["ptr"], node_struct, row_ptr, col, shift,
                                          prep["G"], qp, state["params"], state["grads"], t0, t1, mode=mode,
                                          dropout_uniform=dropout_uniform, dropout_p=p_drop, return_q=return_q)

    def train_graph(self, prep: dict, q_params, state: dict, mode: str, dropout: bool) -> "TrainGraph":
        """The forward + backward of a minibatch SHAPE (graph count, node count) frozen into a CUDA graph — a replay
        minibatch is ~160 small launches, i.e. launch-bound; clip + Adam stay outside (their bias corrections change every
        step, and a multi-GPU run all-reduces the bucket in between)."""
        return TrainGraph(self, prep, q_params, state, mode, dropout)

    def optimizer_step(self, state: dict, lr: float, max_norm: float) -> torch.Tensor:
        """``clip_grad_norm_(parameters, max_norm); optimizer.step()`` (``trainer.py:93-96,173-176``) on ``state``."""
        return self._engine_for(state["device"]).adam_step(state["params"], state["grads"], state, lr, max_norm, state["mask"])

    def load_train_state(self, state: dict) -> None:
        """Copy the flat parameters back into the module (``state_dict()`` / ``save`` / the next scoring call see them)."""
        flat = state["params"].detach().cpu()
        own = dict(self.named_parameters())
        with torch.no_grad():
            for name, off, n, trainable in self._engine_for(state["device"]).train_layout():
                if trainable and name in own:
                    own[name].copy_(flat[off:off + n].reshape(own[name].shape))

    def _pull_q_heads(self, eng: Engine) -> None:
        """Copy the updated Q_emb / Q_output tensors from the engine into this module's parameters (state_dict, save)."""
        sd = dict(self.named_parameters())
        with torch.no_grad():
            for name, n in eng.Q_HEAD_TENSORS:
                p = sd[name]
                p.copy_(torch.from_numpy(eng.get_weight(name, n)).reshape(p.shape))
        eng._sig = tuple((p.data_ptr(), p._version) for p in self._engine_tensors())   # the engine already holds these values

    def _score(self, batch, q_params: Dict[str, float | bool], head_inputs: bool = False) -> Dict[str, torch.Tensor]:
        pos = batch["pos"]
        if pos.device.type != "cuda":
            raise RuntimeError("dqn_v2.forward needs the batch on a CUDA device (no CPU fallback); use batch_to")
        dev = pos.device
        eng = self._engine_for(dev)
        G = batch.num_graphs
        n_atoms_host = np.diff(batch["ptr"].cpu().numpy()) if batch["ptr"].device.type != "cpu" else \
            np.diff(batch["ptr"].numpy())
        cells64 = np.asarray(batch["cell64"], dtype=np.float64)
        if self.dedup_reactants:
            keys = batch["reactant_key"]
            first: Dict[int, int] = {}
            r_of_graph = []
            uniq: List[int] = []
            for g, k in enumerate(keys):
                if k not in first:
                    first[k] = len(uniq)
                    uniq.append(g)
                r_of_graph.append(first[k])
        else:
            uniq = list(range(G))
            r_of_graph = list(range(G))
        nR = len(uniq)
        ptr = batch["ptr"].to(dev)
        if nR == G:
            pos_all = torch.cat([pos, batch["pos_product"]], dim=0)
            Z_all = torch.cat([batch["elems"], batch["elems"]]).to(torch.int32)
            cell_all = torch.cat([batch["cell"], batch["cell"]], dim=0)
            counts = np.concatenate([n_atoms_host, n_atoms_host])
            cells_host = np.concatenate([cells64, cells64])
        else:
            sel = torch.cat([torch.arange(int(ptr[g]), int(ptr[g + 1]), device=dev) for g in uniq])
            pos_all = torch.cat([pos[sel], batch["pos_product"]], dim=0)
            Z_all = torch.cat([batch["elems"][sel], batch["elems"]]).to(torch.int32)
            cell_all = torch.cat([batch["cell"][uniq], batch["cell"]], dim=0)
            counts = np.concatenate([n_atoms_host[uniq], n_atoms_host])
            cells_host = np.concatenate([cells64[uniq], cells64])
        node_ptr_host = np.zeros(len(counts) + 1, dtype=np.int32)
        np.cumsum(counts, out=node_ptr_host[1:])
        inv, img = cell_geometry(cells_host, eng.cutoff)
        node_ptr = torch.from_numpy(node_ptr_host).to(dev)
        inv_d = torch.from_numpy(inv).to(dev)
        img_d = torch.from_numpy(img).to(dev)
        rs = torch.tensor(r_of_graph, dtype=torch.int32, device=dev)
        ps = torch.arange(nR, nR + G, dtype=torch.int32, device=dev)
        qp = eng.q_params(q_params, self.head_spec)
        out = eng.score_device(Z_all.contiguous(), pos_all.contiguous(), cell_all.contiguous(), inv_d, img_d,
                               node_ptr, rs, ps, qp, max_struct_nodes=int(counts.max()), head_inputs=head_inputs)
        out["barrier"] = -out["q0"]
        out["freq"] = out["q1"]
        return out


class TrainGraph:
    """See :meth:`DQNv2.train_graph`.  Inputs live in static device buffers; ``__call__`` copies a prepared minibatch of the
    same shape in, replays, and returns the static outputs (loss, predictions; ``state["grads"]`` holds the gradient)."""

    def __init__(self, model: "DQNv2", prep: dict, q_params, state: dict, mode: str, dropout: bool):
        dev = state["device"]
        self.G, self.n_nodes, self.mode = prep["G"], prep["n_nodes"], mode
        self.static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in prep.items()}
        self.t0 = torch.zeros(self.G, dtype=torch.float32, device=dev)
        self.t1 = torch.zeros(self.G, dtype=torch.float32, device=dev) if mode != "dqn" else None
        self.u = torch.ones((self.G, 80), dtype=torch.float32, device=dev) if dropout else None
        cap = self.n_nodes * 64
        run = lambda: model.train_forward_backward_prepared(self.static, q_params, state, self.t0, self.t1, mode, self.u,
                                                            edge_capacity=cap)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                     # warm-up: sizes the engine's training workspace before the capture
            for _ in range(2):
                run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = run()

    def __call__(self, prep: dict, target0, target1=None, dropout_uniform=None) -> dict:
        for k in ("pos", "Z", "cell", "ptr", "inv", "img"):
            self.static[k].copy_(prep[k], non_blocking=True)
        self.t0.copy_(target0.reshape(-1), non_blocking=True)
        if self.t1 is not None:
            self.t1.copy_(target1.reshape(-1), non_blocking=True)
        if self.u is not None:
            self.u.copy_(dropout_uniform, non_blocking=True)
        self.graph.replay()
        return {k: v.clone() for k, v in self.out.items()}


class TNet(nn.Module, _EngineMixin):
    """``registry.get_model_class("t_net")`` — ``model(batch, inference=True)["time"]``
    (``rlsim/cli/scripts/estimate_time.py:58``, ``rlsim/time/train.py:44-48,108,114``).

    The PaiNN representation is shared with the DQN's reaction model and runs on the same CUDA kernels
    (one graph per frame); the head is ``rlvd_time_readout``.  No t_net checkpoint ships with the reference and
    rgnn's head is not in the repo, so its wiring is a named assumption (``oracle.painn_oracle.TimeHeadSpec``):
    mean-pooled node scalars ‖ T / T_scaler_m ‖ defect / D_scaler → MLP(130 → n_feat → n_feat → 1, SiLU) →
    ``time = tau * sigmoid(.)`` ∈ (0, tau), matching ``timer_converter``'s ``-tau * log(1 - t / tau)``
    (``estimate_time.py:15-35``).  :class:`TNetBinary` (``t_net_binary``) adds the ``"goal"`` logit."""
    HAS_GOAL = False

    def __init__(self, reaction_model: PaiNN, n_feat: int = 64, pooling: str = "mean", dropout_rate: float = 0.1,
                 tau0: float = 1.0, D_scaler: float = 1 / 256, T_scaler_m: float = 7918.951990135471, **_ignored):
        super().__init__()
        if n_feat > 128 or pooling not in ("mean", "sum"):
            raise NotImplementedError("t_net head kernel: n_feat <= 128, pooling in {mean, sum}")
        self.representation = reaction_model.representation
        self.__dict__["_reaction_model"] = reaction_model          # not a registered submodule: no duplicate parameters
        self.cutoff = reaction_model.cutoff
        self.tau0 = tau0
        self.tau = tau0
        self.hyper = {"@name": "t_net", "n_feat": n_feat, "pooling": pooling, "dropout_rate": dropout_rate,
                      "tau0": tau0, "D_scaler": D_scaler, "T_scaler_m": T_scaler_m}
        self.time_output = MLP(reaction_model.hyper["hidden_channels"] + 2, 1, (n_feat, n_feat), "silu", dropout_rate)

    @classmethod
    def load_representation(cls, reaction_model: PaiNN, **cfg) -> "TNet":
        return cls(reaction_model, **cfg)

    # ---- checkpoint I/O in the reference's format, {"state_dict", "hyper_parameters"} (time/train.py:221-226, estimate_time.py:88)
    def save(self, path: str) -> None:
        hp = dict(self.hyper)
        hp["reaction_model"] = _plain(self._reaction_model.hyper)
        hp["tau"] = float(self.tau)
        torch.save({"state_dict": self.state_dict(), "hyper_parameters": hp}, path)

    @classmethod
    def load(cls, path: str) -> "TNet":
        ck = torch.load(path, weights_only=True, map_location="cpu")
        hp = dict(ck["hyper_parameters"])
        hp.pop("@name", None)
        tau = hp.pop("tau", None)
        rm_hp = dict(hp.pop("reaction_model"))
        rm_hp.pop("@name", None)
        model = cls(PaiNN(**rm_hp), **hp)
        model.load_state_dict(ck["state_dict"])
        if tau is not None:
            model.tau = float(tau)
        return model

    # ---- engine plumbing: the engine wants the dqn_v2 checkpoint layout; the Q heads are irrelevant here (zeros)
    def _engine_tensors(self):
        return list(self._reaction_model.parameters()) + list(self._reaction_model.buffers())

    def _engine_state(self):
        sd = {"reaction_model." + k: v for k, v in self._reaction_model.state_dict().items()}
        rf = self._reaction_model.reaction_feat
        for name, shape in (("Q_emb.layers.0.weight", (16, rf + 1)), ("Q_emb.layers.0.bias", (16,)),
                            ("Q_emb.layers.3.weight", (16, 16)), ("Q_emb.layers.3.bias", (16,)),
                            ("Q_output.layers.0.weight", (32, 18)), ("Q_output.layers.0.bias", (32,)),
                            ("Q_output.layers.3.weight", (32, 32)), ("Q_output.layers.3.bias", (32,)),
                            ("Q_output.layers.6.weight", (1, 32)), ("Q_output.layers.6.bias", (1,))):
            sd[name] = torch.zeros(shape)
        return sd, {"@name": "dqn_v2", "reaction_model": self._reaction_model.hyper}

    def _head(self, dev: torch.device, which: str = "time"):
        from ._lib import ACT_CODES, TimeHead
        L = (self.time_output if which == "time" else self.goal_output).layers
        ws = [L[0].weight, L[0].bias, L[3].weight, L[3].bias, L[6].weight, L[6].bias]
        keep = []
        for w in ws:
            t = w.detach()
            if t.device != dev or t.dtype != torch.float32 or not t.is_contiguous():
                t = t.to(device=dev, dtype=torch.float32).contiguous()
            keep.append(t)
        h = TimeHead()
        h.d_w0, h.d_b0, h.d_w1, h.d_b1, h.d_w2, h.d_b2 = [t.data_ptr() for t in keep]
        h.n_feat = int(self.hyper["n_feat"])
        h.pooling = 0 if self.hyper["pooling"] == "mean" else 1
        h.act = ACT_CODES["silu"]
        h.tau = float(self.tau)
        h.T_scale = 1.0 / float(self.hyper["T_scaler_m"])
        h.D_scale = 1.0 / float(self.hyper["D_scaler"])
        h.output = 0 if which == "time" else 1
        return h, keep

    def time_device(self, Z, pos, cell, inv, img, node_ptr, max_struct_nodes: int, T, defect, edge_capacity=None,
                    with_goal: bool = False):
        """The forward on device-resident frames (fp32 ``pos[M,3]``, ``cell / inv [G,3,3]``, ``img[G,3]``, ``node_ptr``):
        K1 → K2-K4 → ``rlvd_time_readout``.  ``edge_capacity``: build the graph without reading the edge count back (frames
        of a sync-free TKS run, config 4)."""
        eng = self._engine_for(pos.device)
        G = node_ptr.numel() - 1
        row_ptr, col, shift, node_struct = eng.radius_graph(pos, cell, inv, img, node_ptr, edge_capacity=edge_capacity)

        def rep():
            return eng.representation(Z, pos, cell, node_struct, row_ptr, col, shift, G, node_ptr=node_ptr,
                                      max_struct_nodes=max_struct_nodes)
        q = rep()
        if edge_capacity is None and eng.check_status() & 1:   # operand outside the tensor-core split's range: fp32 SIMT kernels
            eng.set_option("tensor_cores", 0)
            try:
                q = rep()
            finally:
                eng.set_option("tensor_cores", 1)
        head, keep = self._head(pos.device)
        time = eng.time_readout(q, node_ptr, T, defect, head)
        del keep
        if with_goal:                                           # the same pooled input through the second head (raw logit)
            head, keep = self._head(pos.device, "goal")
            goal = eng.time_readout(q, node_ptr, T, defect, head)
            del keep
            return time, goal
        return time

    def forward(self, batch, inference: bool = False, training: bool = False):
        if torch.is_grad_enabled() and (training or self.training) and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("t_net backward has no sm_100a kernels yet; call under torch.no_grad() / .eval()")
        pos = batch["pos"]
        if pos.device.type != "cuda":
            raise RuntimeError("t_net.forward needs the batch on a CUDA device (no CPU fallback); use batch_to")
        dev = pos.device
        eng = self._engine_for(dev)
        counts = np.diff(batch["ptr"].cpu().numpy())
        node_ptr_host = np.zeros(len(counts) + 1, dtype=np.int32)
        np.cumsum(counts, out=node_ptr_host[1:])
        inv, img = cell_geometry(np.asarray(batch["cell64"], dtype=np.float64), eng.cutoff)
        node_ptr = torch.from_numpy(node_ptr_host).to(dev)
        cell = batch["cell"].to(torch.float32).contiguous()
        posc = pos.to(torch.float32).contiguous()
        Zc = batch["elems"].to(torch.int32).contiguous()
        G = len(counts)
        T = torch.as_tensor(batch["T"], dtype=torch.float32, device=dev).reshape(G).contiguous()
        defect = torch.as_tensor(batch["defect"], dtype=torch.float32, device=dev).reshape(G).contiguous()
        out = self.time_device(Zc, posc, cell, torch.from_numpy(inv).to(dev), torch.from_numpy(img).to(dev), node_ptr,
                               int(counts.max()), T, defect, with_goal=self.HAS_GOAL)
        return {"time": out[0], "goal": out[1]} if self.HAS_GOAL else {"time": out}


class TNetBinary(TNet):
    """``registry.get_model_class("t_net_binary")`` (``rlsim/time/train.py:42-48,128,196``): ``forward`` also returns
    ``"goal"``, the logit ``combined_loss_binary`` feeds to ``binary_cross_entropy_with_logits`` (``time/misc.py:205``).
    rgnn's head is not in the reference; the wiring is the named assumption ``TimeHeadSpec.goal_head``: a second MLP
    (130 -> n_feat -> n_feat -> 1, SiLU) on the same pooled input, raw output.  Same kernels: ``rlvd_time_readout`` with
    ``head.output = 1``."""
    HAS_GOAL = True

    def __init__(self, reaction_model: PaiNN, **cfg):
        super().__init__(reaction_model, **cfg)
        self.hyper["@name"] = "t_net_binary"
        self.goal_output = MLP(reaction_model.hyper["hidden_channels"] + 2, 1, (self.hyper["n_feat"],) * 2, "silu",
                               self.hyper["dropout_rate"])
