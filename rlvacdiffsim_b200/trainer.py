"""Host mirror of ``rlsim/drl/trainer.py`` (``Trainer``) and ``rlsim/memory.py`` (``Memory``) for the replay-minibatch
path, SURVEY.md §8a row B1.

What runs on the device: ``get_max_Q`` (the target network's scoring of every next state, ``trainer.py:102-117`` — the
same forward as ``RLSimulator.select_action``) and, for ``train_all=False`` (``trainer.py:36-39``: every parameter
whose name contains ``reaction_model`` is frozen, only ``Q_emb`` / ``Q_output`` learn), the whole minibatch update of
``dqn_update`` (``trainer.py:163-177``): forward, MSE against ``next_Q * gamma + reward``, backward, ``clip_grad_norm_``
at 0.8 and the Adam step, in one kernel launch behind ``DQNv2.q_head_update`` (``csrc/train.cu``).

``train_all=True`` (the value in both reference configs, ``configs/train_dqn.toml:27``) and ``context_bandit_update``
(``trainer.py:54-100``, fits ``q0`` / ``q1``) back-propagate through the whole PaiNN model: ``csrc/train_full.cu`` —
training forward with saved activations, edge-message / mixing / head backward, the flat 614,447-float gradient bucket,
``clip_grad_norm_`` + Adam on the device.  With ``torch.distributed`` initialised (rl-train on 2/4/8 GPUs)
every rank runs its own minibatches and the bucket is averaged with ONE ``all_reduce`` per optimizer step.
"""
from __future__ import annotations

import copy
import json
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .graph import DataLoader, ReactionDataset, batch_to
from .simulator import convert_to_graph_list
from .structures import Atoms


class Memory:
    """One episode of experience (``rlsim/memory.py:16-62``): the lists ``Trainer`` reads plus ``add`` / ``save`` /
    ``load`` in the reference's JSON layout (``memory.py:88-133``)."""

    def __init__(self, alpha: float, beta: float, T: float = 0.0):
        self.alpha, self.beta, self.T = alpha, beta, T
        self.states: List[Atoms] = []
        self.next_states: List[Atoms] = []
        self.actions: List[int] = []
        self.act_space: List[list] = []
        self.action_probs: List[list] = []
        self.rewards: List[float] = []
        self.freq: List[float] = []
        self.E_min: List[float] = []
        self.barrier: List[float] = []
        self.E_s: List[float] = []
        self.E_next: List[float] = []
        self.fail: List[bool] = []
        self.fail_panelty = -0.5
        self.kb = 1.380649 / 1.602 * 10 ** -4

    def add(self, info: dict) -> None:
        self.states.append(info["state"])
        self.actions.append(info["act"])
        self.act_space.append(info["act_space"])
        self.action_probs.append(info["act_probs"])
        self.fail.append(info["fail"])
        self.next_states.append(info["next"])
        self.freq.append(info["log_freq"])
        self.E_s.append(info["E_s"])
        self.E_min.append(info["E_min"])
        self.E_next.append(info["E_next"])
        if not info["fail"]:
            self.barrier.append(info["E_s"] - info["E_min"])
            self.rewards.append(self.alpha * (-self.barrier[-1] + self.kb * self.T * self.freq[-1])
                                + self.beta * (self.E_min[-1] - self.E_next[-1]))
        else:
            self.barrier.append(0.0)
            self.rewards.append(self.fail_panelty)

    @staticmethod
    def _atoms_dict(a: Atoms) -> dict:
        return {"numbers": np.asarray(a.get_atomic_numbers()).tolist(), "positions": np.asarray(a.get_positions()).tolist(),
                "cell": np.asarray(a.get_cell()).tolist(), "pbc": np.asarray(a.get_pbc()).tolist()}

    def save(self, filename: str) -> None:
        to_list = [self.alpha, self.beta, [self._atoms_dict(u) for u in self.states], self.E_min,
                   [self._atoms_dict(u) for u in self.next_states], self.actions,
                   [[[int(v[0])] + list(v[1:]) for v in u] for u in self.act_space], self.action_probs, self.rewards,
                   self.E_min, self.barrier, self.E_s, self.E_next, self.fail, self.freq]
        with open(filename + ".json", "w") as fh:
            json.dump(to_list, fh)

    def load(self, filename: str) -> None:
        with open(filename) as fh:
            data = json.load(fh)
        (self.alpha, self.beta, states, self.E_min, next_states, self.actions, self.act_space, self.action_probs,
         self.rewards, self.E_min, self.barrier, self.E_s, self.E_next, self.fail, self.freq) = data
        for i in range(len(self.actions)):
            self.states.append(Atoms(states[i]["numbers"], states[i]["positions"], states[i]["cell"]))
            self.next_states.append(Atoms(next_states[i]["numbers"], next_states[i]["positions"], next_states[i]["cell"]))


def convert(atoms, actions: Sequence[Sequence[float]]):
    """``rlsim/drl/trainer.py:185-205`` — the twin of ``convert_to_graph_list``."""
    return convert_to_graph_list(atoms, actions)


class AverageMeter:
    def __init__(self):
        self.sum, self.count = 0.0, 0

    def update(self, v: float, n: int = 1) -> None:
        self.sum += v * n
        self.count += n

    @property
    def avg(self) -> float:
        return self.sum / max(self.count, 1)


class Trainer:
    """``Trainer(model, q_params, offline_model, lr, train_all)`` (``trainer.py:24-46``)."""

    def __init__(self, model, q_params: Dict, offline_model=None, lr: float = 10 ** -3, train_all: bool = True):
        self.policy_value_net = model
        self.target_net = offline_model
        self.train_all = bool(train_all)
        if not train_all:
            for name, param in self.policy_value_net.named_parameters():
                if "reaction_model" in name:
                    param.requires_grad = False
        self.lr = float(lr)
        self.q_params = q_params
        self._adam: Dict[tuple, dict] = {}          # Adam state of the DQN heads, per device (optim.Adam, trainer.py:43)
        self._full: Dict[tuple, dict] = {}          # flat parameters / gradient bucket / Adam moments of the whole model
        self.cuda_graphs = True                     # whole-model minibatches: replay forward + backward as one CUDA graph per shape
        self._graphs: Dict[tuple, object] = {}
        self.heads_only_kernel = True               # train_all=False: use the fused one-launch head update (csrc/train.cu)
        self.allreduce_ms = 0.0                     # accumulated device time of the gradient all-reduce (bench.py --config 5)
        # trainer.py:123 calls policy_value_net.train(): the DQN heads' Dropout(p = hyper['dropout_rate'], 0.15 in the
        # bundled checkpoint) is ON during the update.  Same here by default (uniforms from the global numpy RNG — a
        # different stream from torch's Philox, so masks differ from a torch run draw by draw, not in distribution);
        # set ``trainer.dropout = False`` for deterministic runs / parity tests.
        self.dropout = float(getattr(model, "hyper", {}).get("dropout_rate", 0.0)) > 0.0

    def update(self, memory_l, mode: str = "dqn", **kwargs) -> float:
        if mode == "dqn":
            return self.dqn_update(memory_l, **kwargs)
        return self.context_bandit_update(memory_l, **kwargs)

    # ---- whole-model updates (train_all = True, context bandit): csrc/train_full.cu
    @staticmethod
    def _device(device) -> torch.device:
        dev = torch.device(device) if not isinstance(device, torch.device) else device
        if dev.type == "cuda" and dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        return dev

    def _full_state(self, dev: torch.device) -> dict:
        key = (dev.type, dev.index)
        st = self._full.get(key)
        if st is None:
            st = self.policy_value_net.new_train_state(dev)
            self._full[key] = st
        else:      # the module may have been modified outside (load_state_dict): refresh the flat copy, keep the moments
            st["params"].copy_(self.policy_value_net._engine_for(dev).flatten_state(self.policy_value_net.state_dict()))
        return st

    def _sync_gradients(self, state: dict) -> None:
        """rl-train on several GPUs: average the flat gradient bucket over the ranks (2.46 MB, one NCCL all-reduce)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dist.all_reduce(state["grads"], op=dist.ReduceOp.SUM)
        state["grads"].mul_(1.0 / dist.get_world_size())
        b.record()
        self._ar_events = getattr(self, "_ar_events", []) + [(a, b)]

    def _forward_backward(self, batch, st, dev, mode, u):
        """One minibatch's loss + gradient bucket: eagerly (~160 launches) or, with ``cuda_graphs``, as the replay of a CUDA
        graph captured for this minibatch shape and these q_params (same kernels, same order: bitwise the same numbers)."""
        net = self.policy_value_net
        t0 = batch["rl_q"] if mode == "dqn" else batch["q0"]
        t1 = None if mode == "dqn" else batch["q1"]
        if not self.cuda_graphs:
            return net.train_forward_backward(batch, self.q_params, st, t0.reshape(-1), None if t1 is None else t1.reshape(-1),
                                              mode=mode, dropout_uniform=u)
        prep = net.prepare_train_batch(batch, dev)
        qkey = tuple(sorted((k, float(v)) for k, v in self.q_params.items() if isinstance(v, (int, float, bool))))
        key = (dev.index, mode, prep["G"], prep["n_nodes"], u is not None, qkey, id(st))
        g = self._graphs.get(key)
        if g is None:
            g = net.train_graph(prep, self.q_params, st, mode, u is not None)
            self._graphs[key] = g
        return g(prep, t0.to(dev, torch.float32), None if t1 is None else t1.to(dev, torch.float32), u)

    def _run_minibatches(self, q_dataloader, num_epoch, dev, mode, max_norm) -> float:
        st = self._full_state(dev)
        losses = AverageMeter()
        pending = []
        for _ in range(num_epoch):
            for batch in q_dataloader:
                batch = batch_to(batch, dev)
                G = batch.num_graphs
                u = torch.from_numpy(np.random.random_sample((G, 80)).astype(np.float32)).to(dev) if self.dropout else None
                res = self._forward_backward(batch, st, dev, mode, u)
                self._sync_gradients(st)
                self.policy_value_net.optimizer_step(st, self.lr, max_norm)
                pending.append(res["loss"])
        for l in pending:                                   # one host sync at the end instead of loss.item() per batch
            losses.update(float(l.item()))
        if self.cuda_graphs and self.policy_value_net._engine_for(dev).edge_overflow():
            raise RuntimeError("Trainer: a minibatch held more than 64 neighbours per atom, beyond the captured graphs' edge "
                               "buffer; rerun with trainer.cuda_graphs = False")
        for a, b in getattr(self, "_ar_events", []):
            self.allreduce_ms += a.elapsed_time(b)
        self._ar_events = []
        self.policy_value_net.load_train_state(st)
        return losses.avg

    def context_bandit_update(self, memory_l, episode_size, num_epoch, batch_size=8, device="cuda") -> float:
        """``trainer.py:54-100``: fit (q0, q1) = (-barrier, freq) of the taken actions; clip at 1.0."""
        dev = self._device(device)
        self.policy_value_net.to(dev)
        self.policy_value_net.train()
        age_w = np.asarray([0.99 ** (len(memory_l) - i) for i in range(len(memory_l))])
        episode_size = min(episode_size, len(memory_l))
        picks = np.random.choice(range(len(memory_l)), size=episode_size, p=age_w / np.sum(age_w), replace=False)
        graphs = []
        for u in picks:
            mem = memory_l[u]
            for i, state in enumerate(mem.states):
                for data in convert(state, [mem.act_space[i][mem.actions[i]]]):
                    data.q1 = torch.tensor([float(mem.freq[i])], dtype=torch.float)
                    data.q0 = torch.tensor([-1.0 * float(mem.barrier[i])], dtype=torch.float)
                    graphs.append(data)
        q_dataloader = DataLoader(ReactionDataset(graphs), batch_size=batch_size)
        return self._run_minibatches(q_dataloader, num_epoch, dev, "bandit", 1.0)

    def get_max_Q(self, model, state, action_space, device="cuda") -> torch.Tensor:
        dataloader = DataLoader(ReactionDataset(convert(state, action_space)), batch_size=16, shuffle=False)
        q_values = []
        with torch.no_grad():
            model.eval()
            model.to(device)
            for batch in dataloader:
                batch = batch_to(batch, device)
                q_values.append(model(batch, q_params=self.q_params)["rl_q"].detach())
        return torch.max(torch.cat(q_values, dim=0))

    # ---- the three stages of trainer.py:119-160, split out: target-net sync, replay sampling, TD targets
    def _sync_target(self, device) -> None:
        """``target_net.load_state_dict(policy.state_dict())``; eval / train modes and device placement (trainer.py:120-124)."""
        if self.target_net is None:
            self.target_net = copy.deepcopy(self.policy_value_net)
        self.target_net.load_state_dict(self.policy_value_net.state_dict())
        self.target_net.eval()
        self.target_net.to(device)
        self.policy_value_net.train()
        self.policy_value_net.to(device)

    @staticmethod
    def _sample_transitions(memory_l, episode_size):
        """Episodes drawn with p ∝ 0.99**age (with replacement, trainer.py:126-127); every step but an episode's last
        becomes (state, taken action row, next_state, reward, next action space) (trainer.py:135-146)."""
        age_w = np.asarray([0.99 ** (len(memory_l) - i) for i in range(len(memory_l))])
        picks = np.random.choice(range(len(memory_l)), size=episode_size, p=age_w / np.sum(age_w))
        out = []
        for u in picks:
            mem = memory_l[u]
            n = len(mem.act_space)
            for i in range(n - 1):
                out.append((mem.states[i], mem.act_space[i][mem.actions[i]], mem.next_states[i], mem.rewards[i],
                            mem.act_space[i + 1]))
        return out

    def _td_targets(self, transitions, gamma, device):
        """``next_Q * gamma + reward`` with ``next_Q = max_a Q_target(next_state, a)`` (trainer.py:147-157), fp32 tensors."""
        rewards = torch.tensor([tr[3] for tr in transitions], dtype=torch.float)
        next_q = torch.zeros(len(transitions))
        for i, (_, _, nxt, _, nxt_space) in enumerate(transitions):
            next_q[i] = self.get_max_Q(self.target_net, nxt, nxt_space, device=device)
        return [next_q[i] * gamma + rewards[i] for i in range(len(transitions))]

    def dqn_update(self, memory_l, gamma, episode_size, num_epoch, batch_size=8, device="cuda") -> float:
        self._sync_target(device)
        transitions = self._sample_transitions(memory_l, episode_size)
        targets = self._td_targets(transitions, gamma, device)
        graphs = []
        for (state, taken, _, _, _), y in zip(transitions, targets):
            for data in convert(state, [taken]):
                data.rl_q = y
                graphs.append(data)
        q_dataloader = DataLoader(ReactionDataset(graphs), batch_size=batch_size, shuffle=False)
        dev = self._device(device)
        if self.train_all or not self.heads_only_kernel:
            return self._run_minibatches(q_dataloader, num_epoch, dev, "dqn", 0.8)
        # train_all = False: only Q_emb / Q_output learn -> the frozen forward + the one-launch head update (csrc/train.cu)
        losses = AverageMeter()
        key = (dev.type, dev.index)
        for _ in range(num_epoch):
            for batch in q_dataloader:
                batch = batch_to(batch, dev)
                G = batch.num_graphs
                u = torch.from_numpy(np.random.random_sample((G, 80)).astype(np.float32)).to(dev) if self.dropout else None
                res = self.policy_value_net.q_head_update(batch, self.q_params, batch["rl_q"].reshape(-1), self._adam.get(key),
                                                          self.lr, max_norm=0.8, dropout_uniform=u)
                self._adam[key] = res["state"]
                losses.update(float(res["loss"].item()))
        return losses.avg
