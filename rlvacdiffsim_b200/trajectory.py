"""Buffered, batched writers for the per-replica outputs of a deployment run.

The reference appends ONE XDATCAR frame per replica per step with ``ase.io.write(atoms_traj, atoms,
format="vasp-xdatcar", append=True)`` (``rlsim/drl/simulator.py:175,215``) and rewrites the whole
``q_values.json`` / ``diffuse.json`` after every episode (``rlsim/drl/deploy.py:96-131``).  With 1024 replicas
stepping together on the device that is 1024 file opens per step.  ``TrajectoryWriter`` takes the positions of ALL
replicas as one array per step (one D2H copy from ``DeviceLSS``), keeps frames in memory and writes each replica's
frames in one ``open`` per flush.  The bytes on disk are exactly what per-step ``append_xdatcar`` calls produce
(``tests/test_host_api.py::test_trajectory_writer_matches_per_step_appends``); file names follow ``deploy.py:89``
(``<dir>/XDATCAR<u>``) and the JSON layouts follow ``deploy.py:96-103`` (a list over replicas/episodes of the per-step
lists ``run`` returns).
"""
from __future__ import annotations

import json
import os
from typing import List, Optional, Sequence

import numpy as np

from .structures import Z_TO_SYMBOL


class TrajectoryWriter:
    """Frames of ``n_replicas`` fixed-composition structures, flushed every ``flush_every`` steps.

    numbers: [n_atoms] or [n_replicas, n_atoms] atomic numbers; cells: [3,3] or [n_replicas,3,3] (fixed during the
    run, as in the reference's LSS / TKS modes without relaxation)."""

    def __init__(self, out_dir: str, numbers, cells, n_replicas: Optional[int] = None, flush_every: int = 64,
                 first_replica: int = 0):
        numbers = np.asarray(numbers, dtype=np.int64)
        cells = np.asarray(cells, dtype=np.float64)
        if numbers.ndim == 1:
            if n_replicas is None:
                n_replicas = cells.shape[0] if cells.ndim == 3 else 1
            numbers = np.broadcast_to(numbers, (n_replicas, numbers.shape[0]))
        self.n_rep, self.n_atoms = numbers.shape
        if cells.ndim == 2:
            cells = np.broadcast_to(cells, (self.n_rep, 3, 3))
        if cells.shape != (self.n_rep, 3, 3):
            raise ValueError(f"cells must be [3,3] or [{self.n_rep},3,3], got {cells.shape}")
        self.out_dir = out_dir
        os.makedirs(out_dir, exist_ok=True)
        self.flush_every = int(flush_every)
        self.first_replica = int(first_replica)
        self._headers: List[str] = []
        self._order: List[np.ndarray] = []      # species-block permutation of each replica (VASP groups by species)
        self._cell_T: List[np.ndarray] = []
        for r in range(self.n_rep):
            syms = [Z_TO_SYMBOL[int(z)] for z in numbers[r]]
            order: List[str] = []
            for s in syms:
                if s not in order:
                    order.append(s)
            index = np.concatenate([[i for i, s in enumerate(syms) if s == sym] for sym in order]).astype(np.int64)
            counts = [syms.count(sym) for sym in order]
            head = " ".join(order) + "\n 1.0\n"
            for row in cells[r]:
                head += "  " + " ".join(f"{x:12.6f}" for x in row) + "\n"
            head += " " + " ".join(f"{s:>3s}" for s in order) + "\n"
            head += " " + " ".join(f"{c:3d}" for c in counts) + "\n"
            self._headers.append(head)
            self._order.append(index)
            self._cell_T.append(np.asarray(cells[r], dtype=np.float64).T.copy())
        self._frames: List[np.ndarray] = []     # each [n_rep, n_atoms, 3] fp64 Cartesian
        self._n_written = 0                     # frames already on disk
        self._q: List[List[list]] = [[] for _ in range(self.n_rep)]
        self._actions: List[List[int]] = [[] for _ in range(self.n_rep)]
        self._times: List[List[float]] = [[] for _ in range(self.n_rep)]

    # ------------------------------------------------------------------ per-step records
    def append(self, positions, q_values: Optional[Sequence] = None, chosen: Optional[Sequence[int]] = None,
               times: Optional[Sequence[float]] = None) -> None:
        """positions: [n_replicas, n_atoms, 3] (numpy or a CPU/GPU torch tensor); q_values: per-replica 1-D arrays of
        this step's candidate Q-values (``Qlist.append(Q.tolist())``, simulator.py:217); chosen: per-replica action
        index (simulator.py:223-226); times: per-replica residence time (TKS clock, simulator.py:238)."""
        if hasattr(positions, "detach"):
            positions = positions.detach().cpu().numpy()
        pos = np.asarray(positions, dtype=np.float64).reshape(self.n_rep, self.n_atoms, 3)
        self._frames.append(pos.copy())
        if q_values is not None:
            for r in range(self.n_rep):
                self._q[r].append(np.asarray(q_values[r], dtype=np.float64).tolist())
        if chosen is not None:
            for r in range(self.n_rep):
                self._actions[r].append(int(chosen[r]))
        if times is not None:
            for r in range(self.n_rep):
                self._times[r].append(float(times[r]))
        if len(self._frames) >= self.flush_every:
            self.flush()

    def xdatcar_path(self, replica: int) -> str:
        return os.path.join(self.out_dir, f"XDATCAR{self.first_replica + replica}")

    # ------------------------------------------------------------------ disk
    def flush(self) -> None:
        """Write the buffered frames: one ``open`` per replica, frames numbered as the per-step appends would."""
        if not self._frames:
            return
        frames = np.stack(self._frames)                       # [T, n_rep, n_atoms, 3]
        for r in range(self.n_rep):
            cart = frames[:, r][:, self._order[r]]            # [T, n_atoms, 3], species-grouped
            frac = np.linalg.solve(self._cell_T[r], cart.reshape(-1, 3).T).T.reshape(cart.shape)
            chunks: List[str] = []
            for t in range(frac.shape[0]):
                k = self._n_written + t
                if k == 0:
                    chunks.append(self._headers[r])
                chunks.append(f"Direct configuration={k + 1:6d}\n")
                chunks.append("".join(" " + " ".join(f"{x:11.8f}" for x in row) + "\n" for row in frac[t]))
            with open(self.xdatcar_path(r), "a" if self._n_written > 0 else "w") as fh:
                fh.write("".join(chunks))
        self._n_written += len(self._frames)
        self._frames.clear()

    def close(self) -> None:
        """Flush the frames and write ``q_values.json`` / ``actions.json`` / ``times.json`` (one entry per replica,
        each the per-step list the reference's ``run`` returns)."""
        self.flush()
        if any(self._q):
            with open(os.path.join(self.out_dir, "q_values.json"), "w") as fh:
                json.dump(self._q, fh)
        if any(self._actions):
            with open(os.path.join(self.out_dir, "actions.json"), "w") as fh:
                json.dump(self._actions, fh)
        if any(self._times):
            with open(os.path.join(self.out_dir, "times.json"), "w") as fh:
                json.dump(self._times, fh)
