"""The callers of the t_net path: ``rlsim/cli/scripts/estimate_time.py``.

``timer_converter`` and ``estimate_time`` keep the reference's signatures and file output (a JSON list of per-trajectory
time lists); frames go through ``AtomsGraph.from_ase`` → ``AtomsDataset`` → ``DataLoader(batch_size)`` →
``model(batch, inference=True)["time"]`` exactly as at ``estimate_time.py:38-66``.  :func:`estimate_time_device` is the
B200-first form of the same computation: every frame of every trajectory in a few large device batches
(``TNet.time_device``) instead of ``batch_size = 8`` at a time.  :func:`read_xdatcar` replaces ``ase.io.read(file,
index='::skip')`` (``estimate_time.py:97``) for the XDATCAR files ``TrajectoryWriter`` / the reference write.
"""
from __future__ import annotations

import json
from typing import List, Sequence

import numpy as np
import torch

from .graph import AtomsDataset, AtomsGraph, DataLoader, batch_to
from .structures import SYMBOL_TO_Z, Atoms


def timer_converter(t: torch.Tensor, tau) -> torch.Tensor:
    """``estimate_time.py:15-35``: ``-tau * log(1 - t / tau)``; values above ``tau`` raise like the reference."""
    mask = t > tau
    if torch.any(mask):
        raise ValueError(f"Time value {t[mask]} exceeds the tau {tau}.")
    return -tau * torch.log(1 - t / tau)


def read_xdatcar(path: str, skip: int = 1) -> List[Atoms]:
    """Frames ``::skip`` of a VASP XDATCAR (fixed cell, 'Direct configuration=' blocks)."""
    with open(path) as fh:
        lines = [ln.rstrip("\n") for ln in fh]
    scale = float(lines[1].split()[0])
    cell = np.array([[float(x) for x in lines[i].split()[:3]] for i in (2, 3, 4)]) * scale
    symbols = lines[5].split()
    counts = [int(x) for x in lines[6].split()]
    numbers: List[int] = []
    for sym, cnt in zip(symbols, counts):
        numbers += [SYMBOL_TO_Z[sym]] * cnt
    n = sum(counts)
    frames, i = [], 7
    while i < len(lines):
        if lines[i].strip().lower().startswith("direct"):
            frac = np.array([[float(x) for x in lines[i + 1 + a].split()[:3]] for a in range(n)])
            frames.append(Atoms(numbers, frac @ cell, cell))
            i += n + 1
        else:
            i += 1
    return frames[::max(int(skip), 1)]


def estimate_time(model, temperature, defect, traj_l: Sequence[Sequence[Atoms]], time_file, batch_size, device):
    """``estimate_time.py:38-66``, unchanged flow; returns the list it writes."""
    out = []
    for traj in traj_l:
        datas = []
        for atoms in traj:
            data = AtomsGraph.from_ase(atoms, model.cutoff, read_properties=False, neighborlist_backend="ase")
            data.T = torch.as_tensor(temperature, dtype=torch.get_default_dtype())
            data.defect = torch.as_tensor(defect, dtype=torch.get_default_dtype())
            datas.append(data)
        loader = DataLoader(AtomsDataset(datas), batch_size=batch_size, shuffle=False)
        times = []
        for batch in loader:
            batch = batch_to(batch, device)
            pred = model(batch, inference=True)["time"]
            times.append(timer_converter(pred, model.tau).detach().cpu())
        out.append(torch.cat(times, dim=0).tolist())
    if time_file is not None:
        with open(time_file, "w") as fh:
            json.dump(out, fh)
    return out


def estimate_time_device(model, temperature, defect, traj_l: Sequence[Sequence[Atoms]], time_file=None, device="cuda",
                         frames_per_call: int = 2048):
    """The same numbers with every frame scored in device batches of ``frames_per_call`` (one K1 + representation +
    ``rlvd_time_readout`` pass per batch; frames may differ in size and cell)."""
    from .engine import cell_geometry
    dev = torch.device(device) if not isinstance(device, torch.device) else device
    if dev.type == "cuda" and dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    model.to(dev)
    eng = model._engine_for(dev)
    flat = [(ti, atoms) for ti, traj in enumerate(traj_l) for atoms in traj]
    out: List[List[float]] = [[] for _ in traj_l]
    for s in range(0, len(flat), frames_per_call):
        chunk = flat[s:s + frames_per_call]
        counts = np.array([len(a) for _, a in chunk], dtype=np.int64)
        ptr = np.zeros(len(chunk) + 1, dtype=np.int32)
        np.cumsum(counts, out=ptr[1:])
        pos = np.concatenate([np.asarray(a.get_positions(), dtype=np.float32) for _, a in chunk])
        Z = np.concatenate([np.asarray(a.get_atomic_numbers(), dtype=np.int32) for _, a in chunk])
        cells = np.stack([np.asarray(a.get_cell(), dtype=np.float64) for _, a in chunk])
        inv, img = cell_geometry(cells, eng.cutoff)
        B = len(chunk)
        t = model.time_device(torch.from_numpy(Z).to(dev), torch.from_numpy(pos).to(dev),
                              torch.from_numpy(cells.astype(np.float32)).to(dev), torch.from_numpy(inv).to(dev),
                              torch.from_numpy(img).to(dev), torch.from_numpy(ptr).to(dev), int(counts.max()),
                              torch.full((B,), float(temperature), dtype=torch.float32, device=dev),
                              torch.full((B,), float(defect), dtype=torch.float32, device=dev))
        real = timer_converter(t, model.tau).cpu().tolist()
        for (ti, _), v in zip(chunk, real):
            out[ti].append(v)
    if time_file is not None:
        with open(time_file, "w") as fh:
            json.dump(out, fh)
    return out
