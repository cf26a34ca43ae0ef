"""Python handle on one ``rlvd_engine`` (one per GPU / process).

PyTorch is used here only for device memory, streams and (elsewhere) ``torch.distributed``; every
kernel is launched through the C ABI of ``include/rlvd_b200.h``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import ACT_CODES, QX_CODES, QParams, TimeHead, check


@dataclass(frozen=True)
class HeadSpec:
    """Swappable head assumptions — mirrors ``oracle.painn_oracle.HeadSpec``."""
    delta_e_input: str = "raw"
    painn_head_activation: str = "silu"
    dqn_head_activation: str = "silu"
    q_extra: Tuple[str, str] = ("kT_x10", "delta_e")
    name: str = "HEAD_SPEC_V0"


HEAD_SPEC_V0 = HeadSpec()


def cell_geometry(cells: np.ndarray, cutoff: float):
    """fp32 inverse cells (fp32 image of the fp64 inverse) and per-axis periodic image ranges.

    Same definitions as ``oracle.painn_oracle.inverse_cell_f32`` / ``image_range``: the two sides must
    receive identical matrices for the edge sets to be bit-identical."""
    cells64 = np.asarray(cells, dtype=np.float64).reshape(-1, 3, 3)
    if len(cells64) > 8:
        # consecutive structures usually share a cell (a replica's reactant and its candidate products): solve each
        # run once — the batched LAPACK inverse is the largest host cost of an end-to-end scoring call otherwise
        first = np.concatenate([[True], np.any(cells64[1:] != cells64[:-1], axis=(1, 2))])
        if int(first.sum()) * 2 <= len(cells64):
            inv_u, img_u = _cell_geometry_dense(cells64[first], cutoff)
            run = np.cumsum(first) - 1
            return inv_u[run], img_u[run]
    return _cell_geometry_dense(cells64, cutoff)


def _cell_geometry_dense(cells64: np.ndarray, cutoff: float):
    inv = np.linalg.inv(cells64).astype(np.float32)
    vol = np.abs(np.linalg.det(cells64))
    cross = np.stack([np.cross(cells64[:, 1], cells64[:, 2]), np.cross(cells64[:, 2], cells64[:, 0]),
                      np.cross(cells64[:, 0], cells64[:, 1])], axis=1)
    h = vol[:, None] / np.linalg.norm(cross, axis=2)
    img = np.floor(cutoff / h + 0.5 + 1e-9).astype(np.int32)
    return inv, img


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class Engine:
    """Owns the device-side weights and workspaces of one GPU."""

    def __init__(self, device: int | str | torch.device = 0):
        self.lib = _lib.load()
        dev = torch.device(device) if not isinstance(device, int) else torch.device("cuda", device)
        if dev.type != "cuda":
            raise RuntimeError("rlvacdiffsim_b200 runs on CUDA devices only (no CPU fallback)")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        handle = C.c_void_p()
        check(self.lib.rlvd_engine_create(self.device.index, C.byref(handle)))
        self._h = handle
        self.cutoff = 5.0
        self.means: Dict[str, float] = {}
        self.stddevs: Dict[str, float] = {}
        self._pinned: Dict[str, torch.Tensor] = {}
        self._geom_cache = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self.lib.rlvd_engine_destroy(h)
            self._h = None

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], hyper_parameters: dict) -> None:
        for name, t in state_dict.items():
            t = t.detach().cpu()
            if t.is_floating_point():
                a = np.ascontiguousarray(t.to(torch.float32).numpy().reshape(-1))
                check(self.lib.rlvd_set_weight(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), a.size, 0))
            else:
                a = np.ascontiguousarray(t.to(torch.int32).numpy().reshape(-1))
                check(self.lib.rlvd_set_weight(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), a.size, 1))
        rm = hyper_parameters["reaction_model"]
        species = rm.get("species")
        if species:
            from .structures import SYMBOL_TO_Z
            zs = np.asarray([SYMBOL_TO_Z[s] if isinstance(s, str) else int(s) for s in species], dtype=np.int32)
            check(self.lib.rlvd_set_species(self._h, zs.ctypes.data_as(C.c_void_p), int(zs.size)))
        check(self.lib.rlvd_set_hyper(self._h, b"epsilon", float(rm.get("epsilon", 1e-8))))
        check(self.lib.rlvd_finalize_weights(self._h))
        self.cutoff = float(rm["cutoff"])
        self.means = {k: float(v) for k, v in rm["means"].items()}
        self.stddevs = {k: float(v) for k, v in rm["stddevs"].items()}

    def q_params(self, q_params: dict, spec: HeadSpec = HEAD_SPEC_V0) -> QParams:
        qp = QParams()
        qp.alpha = float(q_params["alpha"])
        qp.beta = float(q_params["beta"])
        qp.dqn = int(bool(q_params.get("dqn", False)))
        qp.temperature = float(q_params.get("temperature", 0.0))
        qp.delta_e_standardised = int(spec.delta_e_input != "raw")
        qp.painn_head_act = ACT_CODES[spec.painn_head_activation]
        qp.dqn_head_act = ACT_CODES[spec.dqn_head_activation]
        qp.q_extra[0] = QX_CODES[spec.q_extra[0]]
        qp.q_extra[1] = QX_CODES[spec.q_extra[1]]
        qp.mean_barrier, qp.std_barrier = self.means["barrier"], self.stddevs["barrier"]
        qp.mean_freq, qp.std_freq = self.means["freq"], self.stddevs["freq"]
        qp.mean_delta_e, qp.std_delta_e = self.means["delta_e"], self.stddevs["delta_e"]
        return qp

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ K1
    def radius_graph(self, pos: torch.Tensor, cell: torch.Tensor, inv_cell: torch.Tensor, img_range: torch.Tensor,
                     node_ptr: torch.Tensor, cutoff: Optional[float] = None, edge_capacity: Optional[int] = None):
        """→ (row_ptr i32[M+1], col i32[E], shift i8[E,3], node_struct i32[M]); receiver-sorted CSR.

        ``edge_capacity``: size ``col`` / ``shift`` for at most this many edges WITHOUT reading the edge count back (no host
        synchronisation: whole LSS / TKS steps queue up on the stream); the returned ``col`` then has ``edge_capacity``
        rows, the tail unused, and ``self.overflow`` (a device flag, read it with :meth:`edge_overflow`) is raised if the
        true count did not fit."""
        cutoff = self.cutoff if cutoff is None else cutoff
        n_struct, M = node_ptr.numel() - 1, pos.shape[0]
        dev = self.device
        row_ptr = torch.empty(M + 1, dtype=torch.int32, device=dev)
        node_struct = torch.empty(max(M, 1), dtype=torch.int32, device=dev)
        n_edges = torch.zeros(1, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            check(self.lib.rlvd_radius_graph_count(self._h, self._stream(), n_struct, M, _ptr(node_ptr), _ptr(pos),
                                                   _ptr(cell), _ptr(inv_cell), _ptr(img_range), cutoff,
                                                   _ptr(row_ptr), _ptr(node_struct), _ptr(n_edges)))
            if edge_capacity is not None:
                E = int(edge_capacity)
                if getattr(self, "overflow", None) is None:
                    self.overflow = torch.zeros(1, dtype=torch.int64, device=dev)
                self.overflow += (n_edges > E).to(torch.int64)
                row_ptr.mul_((n_edges <= E).to(torch.int32))      # a truncated CSR must not be walked: empty every row
                col = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
                shift = torch.empty((max(E, 1), 3), dtype=torch.int8, device=dev)
                check(self.lib.rlvd_radius_graph_fill(self._h, self._stream(), n_struct, M, _ptr(node_ptr), _ptr(pos),
                                                      _ptr(cell), _ptr(inv_cell), _ptr(img_range), cutoff,
                                                      _ptr(row_ptr), E, _ptr(col), _ptr(shift)))
                return row_ptr, col, shift, node_struct[:M]
            E = int(n_edges.item())
            if E >= 2 ** 31 - 1:
                raise RuntimeError(f"{E} edges overflow int32 CSR offsets; split the batch")
            col = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
            shift = torch.empty((max(E, 1), 3), dtype=torch.int8, device=dev)
            check(self.lib.rlvd_radius_graph_fill(self._h, self._stream(), n_struct, M, _ptr(node_ptr), _ptr(pos),
                                                  _ptr(cell), _ptr(inv_cell), _ptr(img_range), cutoff,
                                                  _ptr(row_ptr), E, _ptr(col), _ptr(shift)))
        return row_ptr, col[:E], shift[:E], node_struct[:M]

    # ------------------------------------------------------------------ K2..K4
    def representation(self, Z: torch.Tensor, pos: torch.Tensor, cell: torch.Tensor, node_struct: torch.Tensor,
                       row_ptr: torch.Tensor, col: torch.Tensor, shift: torch.Tensor, n_struct: int,
                       return_mu: bool = False, node_ptr: Optional[torch.Tensor] = None, max_struct_nodes: int = 0):
        M = pos.shape[0]
        if node_ptr is None:
            max_struct_nodes = 0
        q = torch.empty((M, 128), dtype=torch.float32, device=self.device)
        mu = torch.empty((M, 3, 128), dtype=torch.float32, device=self.device) if return_mu else None
        with torch.cuda.device(self.device):
            check(self.lib.rlvd_painn_representation(self._h, self._stream(), n_struct, M, col.shape[0],
                                                     int(max_struct_nodes), _ptr(node_ptr), _ptr(Z),
                                                     _ptr(pos), _ptr(cell), _ptr(node_struct), _ptr(row_ptr),
                                                     _ptr(col), _ptr(shift), _ptr(q), _ptr(mu)))
        return (q, mu) if return_mu else q

    # ------------------------------------------------------------------ K5
    def readout(self, q: torch.Tensor, Z: torch.Tensor, node_ptr: torch.Tensor, react_struct: torch.Tensor,
                prod_struct: torch.Tensor, qp: QParams, temperature: Optional[torch.Tensor] = None,
                head_inputs: bool = False) -> Dict[str, torch.Tensor]:
        """``head_inputs``: also return ``head_in[G, 36]``, what the trainable DQN heads see (``q_head_train_step``)."""
        G = react_struct.numel()
        out = torch.empty((6, max(G, 1)), dtype=torch.float32, device=self.device)
        head_in = torch.empty((max(G, 1), 36), dtype=torch.float32, device=self.device) if head_inputs else None
        with torch.cuda.device(self.device):
            check(self.lib.rlvd_reaction_readout_ex(self._h, self._stream(), G, q.shape[0], _ptr(node_ptr),
                                                    _ptr(react_struct), _ptr(prod_struct), _ptr(Z), _ptr(q),
                                                    C.byref(qp), _ptr(temperature), _ptr(out[0]), _ptr(out[1]),
                                                    _ptr(out[2]), _ptr(out[3]), _ptr(out[4]), _ptr(out[5]), _ptr(head_in)))
        names = ("rl_q", "q0", "q1", "delta_e", "E_R", "E_P")
        res = {k: out[i, :G] for i, k in enumerate(names)}
        if head_inputs:
            res["head_in"] = head_in[:G]
        return res

    # ------------------------------------------------------------------ B1: replay-minibatch update of the DQN heads
    Q_HEAD_TENSORS = (("Q_emb.layers.0.weight", 16 * 33), ("Q_emb.layers.0.bias", 16), ("Q_emb.layers.3.weight", 16 * 16),
                      ("Q_emb.layers.3.bias", 16), ("Q_output.layers.0.weight", 32 * 18), ("Q_output.layers.0.bias", 32),
                      ("Q_output.layers.3.weight", 32 * 32), ("Q_output.layers.3.bias", 32),
                      ("Q_output.layers.6.weight", 32), ("Q_output.layers.6.bias", 1))
    N_Q_HEAD_PARAMS = sum(n for _, n in Q_HEAD_TENSORS)

    def q_head_train_step(self, head_in: torch.Tensor, target: torch.Tensor, state: Dict[str, torch.Tensor], lr: float,
                          head_act: int, max_norm: float = 0.8, betas=(0.9, 0.999), eps: float = 1e-8,
                          dropout_uniform: Optional[torch.Tensor] = None, dropout_p: float = 0.0, apply: bool = True
                          ) -> Dict[str, torch.Tensor]:
        """``rlvd_q_head_train_step`` (``trainer.py:166-175`` with the reaction model frozen): forward of the heads, MSE
        against ``target``, backward, ``clip_grad_norm_(max_norm)``, Adam — in place on the engine's weights.  ``state``
        holds ``exp_avg`` / ``exp_avg_sq`` (f32[2513]) and the integer ``step``; it is advanced when ``apply``."""
        G = head_in.shape[0]
        dev = self.device
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        gnorm = torch.empty(1, dtype=torch.float32, device=dev)
        pred = torch.empty(G, dtype=torch.float32, device=dev)
        grad = torch.empty(self.N_Q_HEAD_PARAMS, dtype=torch.float32, device=dev)
        step = int(state["step"]) + 1
        with torch.cuda.device(dev):
            check(self.lib.rlvd_q_head_train_step(
                self._h, self._stream(), G, _ptr(head_in.contiguous()), _ptr(target.contiguous()), _ptr(dropout_uniform),
                float(dropout_p), int(head_act), float(lr), float(betas[0]), float(betas[1]), float(eps), float(max_norm),
                step, int(apply), _ptr(state["exp_avg"]), _ptr(state["exp_avg_sq"]), _ptr(loss), _ptr(pred), _ptr(gnorm),
                _ptr(grad)))
        if apply:
            state["step"] = step
        return {"loss": loss, "pred": pred, "grad_norm": gnorm, "grad": grad}

    def new_q_head_state(self) -> Dict[str, torch.Tensor]:
        z = lambda: torch.zeros(self.N_Q_HEAD_PARAMS, dtype=torch.float32, device=self.device)
        return {"exp_avg": z(), "exp_avg_sq": z(), "step": 0}

    def get_weight(self, name: str, numel: int) -> np.ndarray:
        out = np.empty(numel, dtype=np.float32)
        with torch.cuda.device(self.device):
            check(self.lib.rlvd_get_weight(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), numel))
        return out

    # ------------------------------------------------------------------ B1, train_all = True: the whole model
    _TRAIN_LAYOUT = None

    @classmethod
    def train_layout(cls):
        """``[(checkpoint name, offset, numel, trainable)]`` of the flat parameter / gradient vector (``rlvd_train_layout``)."""
        if cls._TRAIN_LAYOUT is None:
            lib = _lib.load()
            tab, i = [], 0
            name, off, n, tr = C.c_char_p(), C.c_int64(), C.c_int64(), C.c_int32()
            while lib.rlvd_train_layout(i, C.byref(name), C.byref(off), C.byref(n), C.byref(tr)) == 0:
                tab.append((name.value.decode(), int(off.value), int(n.value), bool(tr.value)))
                i += 1
            assert tab[-1][1] + tab[-1][2] == lib.rlvd_train_param_count()
            cls._TRAIN_LAYOUT = tab
        return cls._TRAIN_LAYOUT

    def flatten_state(self, state_dict: Dict[str, torch.Tensor]) -> torch.Tensor:
        """Every floating-point checkpoint tensor, in ``train_layout`` order, as one fp32 device vector."""
        flat = torch.empty(self.lib.rlvd_train_param_count(), dtype=torch.float32)
        for name, off, n, _ in self.train_layout():
            t = state_dict[name].detach().to("cpu", torch.float32).reshape(-1)
            if t.numel() != n:
                raise ValueError(f"{name}: {t.numel()} elements, the kernels expect {n}")
            flat[off:off + n] = t
        return flat.to(self.device)

    def train_forward_backward(self, Z, pos, cell, node_ptr, node_struct, row_ptr, col, shift, n_react: int, qp: QParams,
                               params: torch.Tensor, grads: torch.Tensor, target0: torch.Tensor,
                               target1: Optional[torch.Tensor] = None, mode: str = "dqn",
                               temperature: Optional[torch.Tensor] = None, dropout_uniform: Optional[torch.Tensor] = None,
                               dropout_p: float = 0.0, return_q: bool = False) -> Dict[str, torch.Tensor]:
        """``rlvd_train_forward_backward``: loss, predictions and the UNCLIPPED gradient of every parameter for one replay
        minibatch (``trainer.py:166-172`` / ``:89-92``).  ``grads`` is overwritten."""
        M, dev = pos.shape[0], self.device
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        pred = torch.empty((n_react, 3), dtype=torch.float32, device=dev)
        q = torch.empty((M, 128), dtype=torch.float32, device=dev) if return_q else None
        with torch.cuda.device(dev):
            check(self.lib.rlvd_train_forward_backward(
                self._h, self._stream(), node_ptr.numel() - 1, M, col.shape[0], _ptr(node_ptr), _ptr(Z), _ptr(pos), _ptr(cell),
                _ptr(node_struct), _ptr(row_ptr), _ptr(col), _ptr(shift), n_react, C.byref(qp), _ptr(temperature),
                0 if mode == "dqn" else 1, _ptr(target0), _ptr(target1), _ptr(dropout_uniform), float(dropout_p), _ptr(params),
                _ptr(grads), _ptr(loss), _ptr(pred), _ptr(q)))
        out = {"loss": loss, "rl_q": pred[:, 0], "q0": pred[:, 1], "q1": pred[:, 2]}
        if return_q:
            out["q"] = q
        return out

    def adam_step(self, params: torch.Tensor, grads: torch.Tensor, state: Dict[str, torch.Tensor], lr: float,
                  max_norm: float, trainable: Optional[torch.Tensor] = None, betas=(0.9, 0.999), eps: float = 1e-8) -> torch.Tensor:
        """``clip_grad_norm_(max_norm)`` + ``torch.optim.Adam.step`` in place on ``params``; → [total norm, clip coefficient]."""
        info = torch.empty(2, dtype=torch.float32, device=self.device)
        state["step"] = int(state["step"]) + 1
        with torch.cuda.device(self.device):
            check(self.lib.rlvd_adam_step(self._h, self._stream(), params.numel(), _ptr(params), _ptr(grads), _ptr(trainable),
                                          _ptr(state["exp_avg"]), _ptr(state["exp_avg_sq"]), float(lr), float(betas[0]),
                                          float(betas[1]), float(eps), float(max_norm), int(state["step"]), _ptr(info)))
        return info

    # ------------------------------------------------------------------ A1: candidate hops on the device
    def action_space(self, pos: torch.Tensor, cells: torch.Tensor, node_ptr: torch.Tensor, lattice_parameter: float,
                     max_vacancies: int = 1, max_actions: int = 16):
        """``rlvd_action_space``: fp64 ``pos[M,3]``, ``cells[G,3,3]`` → (count i32[G], actions f64[G,max_actions,4],
        status i32[G]) with the reference's row order (``rlsim/actions/action.py:12-61``)."""
        G = node_ptr.numel() - 1
        count = torch.empty(max(G, 1), dtype=torch.int32, device=self.device)
        status = torch.empty(max(G, 1), dtype=torch.int32, device=self.device)
        actions = torch.zeros((max(G, 1), max_actions, 4), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.rlvd_action_space(self._h, self._stream(), G, _ptr(node_ptr), _ptr(pos), _ptr(cells),
                                             float(lattice_parameter), int(max_vacancies), int(max_actions),
                                             _ptr(count), _ptr(actions), _ptr(status)))
        return count[:G], actions[:G], status[:G]

    def action_space_general(self, pos: torch.Tensor, cells: torch.Tensor, inv_cells: torch.Tensor, node_ptr: torch.Tensor,
                             site_ptr: torch.Tensor, sites: torch.Tensor, lattice_parameter: float, max_vacancies: int = 1,
                             max_actions: int = 16):
        """``rlvd_action_space_general``: any periodic cell; ``sites`` (fp64) / ``site_ptr`` = the perfect-lattice sites of
        every structure's cell (``actions.lattice_sites_for_cell``), ``inv_cells`` the fp64 inverse cells."""
        G = node_ptr.numel() - 1
        count = torch.empty(max(G, 1), dtype=torch.int32, device=self.device)
        status = torch.empty(max(G, 1), dtype=torch.int32, device=self.device)
        actions = torch.zeros((max(G, 1), max_actions, 4), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.rlvd_action_space_general(self._h, self._stream(), G, _ptr(node_ptr), _ptr(pos), _ptr(cells), _ptr(inv_cells),
                                                     _ptr(site_ptr), _ptr(sites), float(lattice_parameter), int(max_vacancies),
                                                     int(max_actions), _ptr(count), _ptr(actions), _ptr(status)))
        return count[:G], actions[:G], status[:G]

    # ------------------------------------------------------------------ A2/A3/A5 on the device
    def candidate_offsets(self, count: torch.Tensor):
        R = count.numel()
        first = torch.empty(R + 1, dtype=torch.int32, device=self.device)
        total = torch.empty(1, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.rlvd_candidate_offsets(self._h, self._stream(), R, _ptr(count), _ptr(first), _ptr(total)))
        return first, total

    def expand_candidates(self, rep_ptr, first, n_cand: int, n_atoms_per: int, Z, pos, cell, actions, inv_in=None, img_in=None):
        """→ dict of the scoring-batch tensors (Z, pos, cell, inv, img, ptr, rs, ps, owner) for ``score_device``.
        ``inv_in`` f32[R,3,3] / ``img_in`` i32[R,3]: the replicas' inverse cells / image ranges (general cells)."""
        R, S = rep_ptr.numel() - 1, 2 * n_cand
        dev = self.device
        o = {"Z": torch.empty(S * n_atoms_per, dtype=torch.int32, device=dev),
             "pos": torch.empty((S * n_atoms_per, 3), dtype=torch.float32, device=dev),
             "cell": torch.empty((S, 3, 3), dtype=torch.float32, device=dev),
             "inv": torch.empty((S, 3, 3), dtype=torch.float32, device=dev),
             "img": torch.empty((S, 3), dtype=torch.int32, device=dev),
             "ptr": torch.empty(S + 1, dtype=torch.int32, device=dev),
             "rs": torch.empty(n_cand, dtype=torch.int32, device=dev),
             "ps": torch.empty(n_cand, dtype=torch.int32, device=dev),
             "owner": torch.empty(n_cand, dtype=torch.int32, device=dev)}
        with torch.cuda.device(dev):
            check(self.lib.rlvd_expand_candidates_ex(self._h, self._stream(), R, n_cand, n_atoms_per, _ptr(rep_ptr), _ptr(first),
                                                     _ptr(Z), _ptr(pos), _ptr(cell), _ptr(actions), actions.shape[1], _ptr(inv_in),
                                                     _ptr(img_in), _ptr(o["ptr"]), _ptr(o["Z"]), _ptr(o["pos"]), _ptr(o["cell"]),
                                                     _ptr(o["inv"]), _ptr(o["img"]), _ptr(o["rs"]), _ptr(o["ps"]), _ptr(o["owner"])))
        return o

    def select_apply(self, rep_ptr, first, rl_q, temperature, uniform, actions, pos, apply: bool = True,
                     philox: Optional[Tuple[int, int, int]] = None):
        """→ (chosen i32[R], argmax i32[R], probs f32[R, max_actions], gamma f32[R]); ``pos`` (fp64) is updated in place
        when ``apply``; gamma = sum(exp(Q / kT)) is the TKS total rate (simulator.py:262).  ``uniform``: the caller's
        draws (``np.random``-compatible replay); ``uniform=None`` with ``philox=(seed, step, replica0)``: counter-based
        draws on the device (``rlvd_select_apply_philox``)."""
        R = rep_ptr.numel() - 1
        if uniform is None:
            if philox is None:
                raise ValueError("select_apply needs uniforms or a (seed, step, replica0) Philox counter")
            chosen = torch.empty(R, dtype=torch.int32, device=self.device)
            amax = torch.empty(R, dtype=torch.int32, device=self.device)
            probs = torch.zeros((R, actions.shape[1]), dtype=torch.float32, device=self.device)
            gamma = torch.empty(R, dtype=torch.float32, device=self.device)
            with torch.cuda.device(self.device):
                check(self.lib.rlvd_select_apply_philox(self._h, self._stream(), R, _ptr(rep_ptr), _ptr(first), _ptr(rl_q),
                                                        _ptr(temperature), int(philox[0]) & (2 ** 64 - 1), int(philox[1]),
                                                        int(philox[2]), _ptr(actions), actions.shape[1], int(apply), _ptr(pos),
                                                        _ptr(chosen), _ptr(amax), _ptr(probs), _ptr(gamma)))
            return chosen, amax, probs, gamma
        chosen = torch.empty(R, dtype=torch.int32, device=self.device)
        amax = torch.empty(R, dtype=torch.int32, device=self.device)
        probs = torch.zeros((R, actions.shape[1]), dtype=torch.float32, device=self.device)
        gamma = torch.empty(R, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.rlvd_select_apply(self._h, self._stream(), R, _ptr(rep_ptr), _ptr(first), _ptr(rl_q),
                                             _ptr(temperature), _ptr(uniform), _ptr(actions), actions.shape[1], int(apply),
                                             _ptr(pos), _ptr(chosen), _ptr(amax), _ptr(probs), _ptr(gamma)))
        return chosen, amax, probs, gamma

    def pack_step_records(self, first, rl_q, chosen, E_R, replica_id0: int, replica_id_stride: int, max_actions: int) -> torch.Tensor:
        """``rlvd_pack_step_records`` → f32[R, 4 + max_actions] rows {replica id, n_actions, chosen, E_R, Q...}."""
        R = first.numel() - 1
        rec = torch.empty((R, 4 + max_actions), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.rlvd_pack_step_records(self._h, self._stream(), R, _ptr(first), _ptr(rl_q), _ptr(chosen), _ptr(E_R),
                                                  int(replica_id0), int(replica_id_stride), int(max_actions), _ptr(rec)))
        return rec

    # ------------------------------------------------------------------ Warren-Cowley pair counts
    def sro_counts(self, pos: torch.Tensor, cells: torch.Tensor, node_ptr: torch.Tensor, species_index: torch.Tensor,
                   n_species: int):
        G = node_ptr.numel() - 1
        counts = torch.zeros((max(G, 1), n_species, n_species), dtype=torch.int64, device=self.device)
        status = torch.zeros(max(G, 1), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.rlvd_sro_counts(self._h, self._stream(), G, _ptr(node_ptr), _ptr(pos), _ptr(cells),
                                           _ptr(species_index), n_species, _ptr(counts), _ptr(status)))
        return counts[:G], status[:G]

    # ------------------------------------------------------------------ T1: t_net head
    def time_readout(self, q: torch.Tensor, node_ptr: torch.Tensor, T: torch.Tensor, defect: torch.Tensor,
                     head: TimeHead) -> torch.Tensor:
        """``rlvd_time_readout``: pooled node scalars + (T, defect) → MLP → tau * sigmoid, one value per structure."""
        G = node_ptr.numel() - 1
        out = torch.empty(max(G, 1), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.rlvd_time_readout(self._h, self._stream(), G, _ptr(node_ptr), _ptr(q), _ptr(T), _ptr(defect),
                                             C.byref(head), _ptr(out)))
        return out[:G]

    # ------------------------------------------------------------------ whole path, device-resident inputs
    def edge_overflow(self) -> int:
        """Number of capacity-bounded radius graphs whose true edge count did not fit (synchronises; clears the flag)."""
        if getattr(self, "overflow", None) is None:
            return 0
        n = int(self.overflow.item())
        self.overflow.zero_()
        return n

    def score_device(self, Z, pos, cell, inv_cell, img_range, node_ptr, react_struct, prod_struct, qp: QParams,
                     temperature=None, max_struct_nodes: int = 0, head_inputs: bool = False,
                     edge_capacity: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """K1..K5 on device-resident inputs.  With ``edge_capacity`` the call never synchronises with the host: the graph is
        built into a buffer of that many edges and the fp16-range status word is NOT polled (``check_status`` /
        ``edge_overflow`` at the caller's next synchronisation point)."""
        n_struct = node_ptr.numel() - 1
        row_ptr, col, shift, node_struct = self.radius_graph(pos, cell, inv_cell, img_range, node_ptr, edge_capacity=edge_capacity)
        if edge_capacity is not None:
            q = self.representation(Z, pos, cell, node_struct, row_ptr, col, shift, n_struct, node_ptr=node_ptr,
                                    max_struct_nodes=max_struct_nodes)
            out = self.readout(q, Z, node_ptr, react_struct, prod_struct, qp, temperature, head_inputs=head_inputs)
            out["n_edges"] = int(edge_capacity)
            return out

        def run():
            q = self.representation(Z, pos, cell, node_struct, row_ptr, col, shift, n_struct, node_ptr=node_ptr,
                                    max_struct_nodes=max_struct_nodes)
            return self.readout(q, Z, node_ptr, react_struct, prod_struct, qp, temperature, head_inputs=head_inputs)

        out = run()
        status = self.check_status()
        if status & 2:
            raise RuntimeError("rlvd_b200: the edge-tile pool overflowed its capacity bound; results discarded")
        if status & 1:
            # an activation left the fp16-split range of the tensor-core GEMMs (the network diverging on an unphysical
            # structure, |x| >= 65504): repeat on the fp32 SIMT kernels, which carry any finite fp32 value
            self.set_option("tensor_cores", 0)
            try:
                out = run()
            finally:
                self.set_option("tensor_cores", 1)
        out["n_edges"] = col.shape[0]
        return out

    def capture_score(self, Z, pos, cell, inv_cell, img_range, node_ptr, react_struct, prod_struct, qp: QParams,
                      temperature: torch.Tensor, max_struct_nodes: int, edge_capacity: Optional[int] = None) -> "GraphedScore":
        """Capture ONE scoring call (K1..K5, ~30 launches) of this batch SHAPE as a CUDA graph: the reference's deployment
        shape is a single replica stepping serially (deploy.py:75-94: 12 candidates per call), where launch latency, not
        kernel time, sets the pace.  The returned object replays the graph on new positions / species / temperatures of the
        same shape (``GraphedScore.__call__``)."""
        return GraphedScore(self, Z, pos, cell, inv_cell, img_range, node_ptr, react_struct, prod_struct, qp, temperature,
                            max_struct_nodes, edge_capacity)

    def check_status(self) -> int:
        """``rlvd_check_status``: synchronises the current stream, returns and clears the device status word."""
        flags = C.c_int32(0)
        with torch.cuda.device(self.device):
            check(self.lib.rlvd_check_status(self._h, self._stream(), C.byref(flags)))
        return int(flags.value)

    # ------------------------------------------------------------------ whole path, host buffers (e2e)
    def _pin(self, key: str, arr: np.ndarray) -> torch.Tensor:
        """Copy ``arr`` into a cached pinned staging tensor and return the (sliced) tensor."""
        t = torch.from_numpy(np.ascontiguousarray(arr))
        if t.is_pinned():                  # the caller's array already lives in page-locked memory (pin_arrays): no staging copy
            return t.reshape(-1)
        buf = self._pinned.get(key)
        if buf is None or buf.numel() < t.numel() or buf.dtype != t.dtype:
            buf = torch.empty(max(t.numel(), 16), dtype=t.dtype).pin_memory()
            self._pinned[key] = buf
        view = buf[: t.numel()]
        view.copy_(t.reshape(-1))
        return view

    @staticmethod
    def pin_arrays(arrays: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
        """Copies of ``arrays`` that live in page-locked host memory (numpy views of pinned torch tensors): ``score_host``
        then DMA-copies them straight to the device instead of staging them through its own pinned buffers first."""
        out = {}
        for k, a in arrays.items():
            if not isinstance(a, np.ndarray):
                out[k] = a
                continue
            t = torch.empty(a.shape, dtype=torch.from_numpy(np.empty(0, dtype=a.dtype)).dtype).pin_memory()
            v = t.numpy()
            v[...] = a
            out[k] = v
        return out

    def score_host(self, Z: np.ndarray, pos: np.ndarray, cells: np.ndarray, node_ptr: np.ndarray,
                   react_struct: np.ndarray, prod_struct: np.ndarray, qp: QParams,
                   temperature: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
        """H2D + K1..K5 + D2H in one C call (``rlvd_score_reactions_host``); inputs are numpy arrays."""
        n_struct, M, G = len(node_ptr) - 1, int(node_ptr[-1]), len(react_struct)
        # cells rarely change between calls (LSS / TKS keep the box fixed): reuse the inverse cells / image ranges when the
        # bytes are the same as last time
        c64 = np.ascontiguousarray(cells, dtype=np.float64)
        key = (c64.shape, c64.tobytes())
        if self._geom_cache is not None and self._geom_cache[0] == key:
            inv, img = self._geom_cache[1]
        else:
            inv, img = cell_geometry(c64, self.cutoff)
            self._geom_cache = (key, (inv, img))
        h = {
            "node_ptr": self._pin("node_ptr", np.asarray(node_ptr, dtype=np.int32)),
            "Z": self._pin("Z", np.asarray(Z, dtype=np.int32)),
            "pos": self._pin("pos", np.asarray(pos, dtype=np.float32)),
            "cell": self._pin("cell", np.asarray(cells, dtype=np.float32)),
            "inv": self._pin("inv", inv),
            "img": self._pin("img", img),
            "rs": self._pin("rs", np.asarray(react_struct, dtype=np.int32)),
            "ps": self._pin("ps", np.asarray(prod_struct, dtype=np.int32)),
        }
        t = None if temperature is None else self._pin("temp", np.asarray(temperature, dtype=np.float32))
        out = self._pinned.get("out")
        if out is None or out.numel() < 6 * G:
            out = torch.empty(max(6 * G, 16), dtype=torch.float32).pin_memory()
            self._pinned["out"] = out
        o = [out[i * G:(i + 1) * G] for i in range(6)]
        n_edges = C.c_int64(0)
        with torch.cuda.device(self.device):
            check(self.lib.rlvd_score_reactions_host(
                self._h, self._stream(), n_struct, M, _ptr(h["node_ptr"]), _ptr(h["Z"]), _ptr(h["pos"]),
                _ptr(h["cell"]), _ptr(h["inv"]), _ptr(h["img"]), self.cutoff, G, _ptr(h["rs"]), _ptr(h["ps"]),
                C.byref(qp), _ptr(t), _ptr(o[0]), _ptr(o[1]), _ptr(o[2]), _ptr(o[3]), _ptr(o[4]), _ptr(o[5]),
                C.byref(n_edges)))
        names = ("rl_q", "q0", "q1", "delta_e", "E_R", "E_P")
        res = {k: o[i].numpy().copy() for i, k in enumerate(names)}
        res["n_edges"] = int(n_edges.value)
        return res

    # ------------------------------------------------------------------ standalone linear + options
    def set_option(self, name: str, value: int) -> None:
        check(self.lib.rlvd_set_option(self._h, name.encode(), int(value)))

    def linear(self, A: torch.Tensor, W: np.ndarray, bias: Optional[np.ndarray] = None, act: int = 0,
               A2: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Y = act([A | A2] @ W.T + bias) through ``rlvd_linear`` (W, bias: host fp32 arrays)."""
        M, K1 = A.shape
        K = K1 + (A2.shape[1] if A2 is not None else 0)
        W = np.ascontiguousarray(W, dtype=np.float32)
        N = W.shape[0]
        assert W.shape[1] == K
        Y = torch.empty((M, N), dtype=torch.float32, device=self.device)
        b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
        with torch.cuda.device(self.device):
            check(self.lib.rlvd_linear(self._h, self._stream(), M, K, N, _ptr(A), _ptr(A2), K1,
                                       W.ctypes.data_as(C.c_void_p),
                                       None if b is None else b.ctypes.data_as(C.c_void_p), act, _ptr(Y)))
        return Y

    # ------------------------------------------------------------------ introspection
    def launch_count(self) -> int:
        return int(self.lib.rlvd_launch_count(self._h))

    def profile(self, on: bool) -> None:
        check(self.lib.rlvd_profile_enable(self._h, int(on)))

    def profile_reset(self) -> None:
        check(self.lib.rlvd_profile_reset(self._h))

    def profile_read(self, family: str) -> Tuple[float, int]:
        ms, n = C.c_double(0.0), C.c_int64(0)
        check(self.lib.rlvd_profile_read(self._h, family.encode(), C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)


class GraphedScore:
    """A scoring call frozen into a CUDA graph (see :meth:`Engine.capture_score`).  Inputs live in static device buffers;
    ``__call__`` copies the new state in, replays, and returns the (static) output tensors.  The sync-free form of
    ``score_device`` is what gets captured: worst-case edge capacity, no status polling inside — ``Engine.edge_overflow`` /
    ``check_status`` at the caller's synchronisation points."""

    def __init__(self, engine: Engine, Z, pos, cell, inv_cell, img_range, node_ptr, react_struct, prod_struct, qp, temperature,
                 max_struct_nodes: int, edge_capacity: Optional[int]):
        self.engine = engine
        dev = engine.device
        self.static = {k: v.clone() for k, v in (("Z", Z), ("pos", pos), ("cell", cell), ("inv", inv_cell), ("img", img_range),
                                                  ("ptr", node_ptr), ("rs", react_struct), ("ps", prod_struct), ("T", temperature))}
        self.qp = qp
        self.max_nodes = int(max_struct_nodes)
        self.cap = int(edge_capacity) if edge_capacity is not None else int(pos.shape[0]) * 64
        s = self.static
        run = lambda: engine.score_device(s["Z"], s["pos"], s["cell"], s["inv"], s["img"], s["ptr"], s["rs"], s["ps"], self.qp,
                                          temperature=s["T"], max_struct_nodes=self.max_nodes, edge_capacity=self.cap)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                     # warm-up on the side stream: sizes every engine workspace
            for _ in range(3):
                run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = run()
        self.launches_per_replay = None

    def __call__(self, pos: Optional[torch.Tensor] = None, Z: Optional[torch.Tensor] = None,
                 temperature: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        if pos is not None:
            self.static["pos"].copy_(pos, non_blocking=True)
        if Z is not None:
            self.static["Z"].copy_(Z, non_blocking=True)
        if temperature is not None:
            self.static["T"].copy_(temperature, non_blocking=True)
        self.graph.replay()
        return self.out
