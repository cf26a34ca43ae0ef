"""Graph containers, datasets and the collating loader of the hot path.

Mirrors the surface the reference uses from ``rgnn.graph`` and ``torch_geometric.loader``:
``ReactionGraph.from_ase(reactant, product)`` (``rlsim/drl/simulator.py:456``, ``trainer.py:202``),
``AtomsGraph.from_ase(atoms, cutoff, read_properties=False, neighborlist_backend="ase", add_batch=True)``
(``rlsim/time/misc.py:109-113``, ``estimate_time.py:46``), mapping access ``data["elems"]``, attribute
assignment ``data.T = ...`` (``misc.py:115-141``, ``trainer.py:81-82,157``), ``ReactionDataset(list)``,
``AtomsDataset(list)``, ``DataLoader(dataset, batch_size, shuffle, sampler)`` collating to one
disjoint-union batch, and ``batch_to(batch, device)``.

No neighbor list is built on the host: like the reference's ``ReactionGraph.from_ase`` (called without
a cutoff) the edges are found on the device at the model's cutoff by the K1 kernel.  ``AtomsGraph``
accepts the ``cutoff`` / ``neighborlist_backend`` arguments for signature parity and ignores the
backend for the same reason.
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, Iterator, List, Optional, Sequence

import numpy as np
import torch

_NODE_KEYS = ("elems", "pos", "pos_product")


class _Graph:
    """Attribute + mapping container of tensors describing one structure (or reactant/product pair)."""

    def __init__(self, **fields):
        object.__setattr__(self, "_store", {})
        for k, v in fields.items():
            self._store[k] = v

    def __getitem__(self, key: str):
        return self._store[key]

    def __setitem__(self, key: str, value) -> None:
        self._store[key] = value

    def __contains__(self, key: str) -> bool:
        return key in self._store

    def __getattr__(self, key: str):
        store = object.__getattribute__(self, "_store")
        if key in store:
            return store[key]
        raise AttributeError(key)

    def __setattr__(self, key: str, value) -> None:
        self._store[key] = value

    def keys(self):
        return self._store.keys()

    @property
    def num_nodes(self) -> int:
        return int(self._store["elems"].shape[0])

    def to(self, device):
        out = self.__class__.__new__(self.__class__)
        object.__setattr__(out, "_store", {k: (v.to(device) if torch.is_tensor(v) else v)
                                           for k, v in self._store.items()})
        return out


def _structure_fields(atoms) -> Dict[str, Any]:
    numbers = np.asarray(atoms.get_atomic_numbers(), dtype=np.int64)
    pos = np.asarray(atoms.get_positions(), dtype=np.float64)
    cell = np.asarray(atoms.get_cell(), dtype=np.float64).reshape(3, 3)
    return {
        "elems": torch.from_numpy(numbers),
        "pos": torch.from_numpy(pos.astype(np.float32)),
        "cell": torch.from_numpy(cell.astype(np.float32)).reshape(1, 3, 3),
        "n_atoms": torch.tensor([len(numbers)], dtype=torch.int64),
        "cell64": cell,                      # host copy for the exact inverse / image ranges
        "reactant_key": hash(pos.tobytes()) ^ hash(numbers.tobytes()) ^ hash(cell.tobytes()),
    }


class ReactionGraph(_Graph):
    """Reactant/product pair with identical species and cell; differs by one displaced atom."""

    @classmethod
    def from_ase(cls, reactant, product) -> "ReactionGraph":
        f = _structure_fields(reactant)
        p = np.asarray(product.get_positions(), dtype=np.float64)
        if p.shape[0] != f["pos"].shape[0]:
            raise ValueError("reactant and product differ in atom count")
        f["pos_product"] = torch.from_numpy(p.astype(np.float32))
        return cls(**f)


class AtomsGraph(_Graph):
    @classmethod
    def from_ase(cls, atoms, cutoff: Optional[float] = None, read_properties: bool = False,
                 neighborlist_backend: str = "ase", add_batch: bool = True) -> "AtomsGraph":
        f = _structure_fields(atoms)
        f["cutoff"] = cutoff
        return cls(**f)


class _ListDataset(torch.utils.data.Dataset):
    def __init__(self, graphs: Sequence[_Graph]):
        self.graphs = list(graphs)

    def __len__(self) -> int:
        return len(self.graphs)

    def __getitem__(self, idx):
        return self.graphs[idx]


class ReactionDataset(_ListDataset):
    pass


class AtomsDataset(_ListDataset):
    pass


class Batch(_Graph):
    """Disjoint union of graphs: node tensors concatenated, ``batch`` / ``ptr`` index vectors, per-graph
    tensors stacked.  ``cell64`` stays a host array [G,3,3]."""

    @property
    def num_graphs(self) -> int:
        return int(self._store["n_atoms"].shape[0])

    @classmethod
    def from_data_list(cls, graphs: Sequence[_Graph]) -> "Batch":
        out: Dict[str, Any] = {}
        keys = list(graphs[0].keys())
        for k in keys:
            vals = [g[k] for g in graphs]
            if k == "cell64":
                out[k] = np.stack(vals)
            elif k == "reactant_key":
                out[k] = list(vals)
            elif k == "cutoff":
                out[k] = vals[0]
            elif torch.is_tensor(vals[0]):
                out[k] = torch.stack(vals) if vals[0].dim() == 0 else torch.cat(vals, dim=0)
            else:
                out[k] = torch.as_tensor(vals)
        n = out["n_atoms"]
        out["batch"] = torch.repeat_interleave(torch.arange(len(graphs), dtype=torch.int64), n)
        ptr = torch.zeros(len(graphs) + 1, dtype=torch.int64)
        ptr[1:] = torch.cumsum(n, 0)
        out["ptr"] = ptr
        return cls(**out)


class DataLoader:
    """``torch_geometric.loader.DataLoader`` stand-in: batches of ``batch_size`` graphs → :class:`Batch`."""

    def __init__(self, dataset, batch_size: int = 1, shuffle: bool = False, sampler: Optional[Iterable[int]] = None,
                 **_ignored):
        self.dataset, self.batch_size, self.shuffle, self.sampler = dataset, int(batch_size), shuffle, sampler

    def __len__(self) -> int:
        n = len(self.sampler) if self.sampler is not None else len(self.dataset)
        return (n + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[Batch]:
        if self.sampler is not None:
            order = list(iter(self.sampler))
        elif self.shuffle:
            order = np.random.permutation(len(self.dataset)).tolist()
        else:
            order = list(range(len(self.dataset)))
        for s in range(0, len(order), self.batch_size):
            yield Batch.from_data_list([self.dataset[int(i)] for i in order[s:s + self.batch_size]])


def batch_to(batch: _Graph, device) -> _Graph:
    """``rgnn.graph.utils.batch_to`` (``simulator.py:55``): move every tensor of the batch to ``device``."""
    return batch.to(device)
