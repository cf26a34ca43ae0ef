"""rlvacdiffsim_b200 — B200-native (sm_100a) candidate vacancy-hop scoring path of RLVacDiffSim.

Only the hot path named in BASELINE.json lives here: periodic radius graph → edge embedding → PaiNN
message passing → deterministic segmented scatter → per-action Q readout, behind the reference's own
Python surface (``registry``, ``ReactionGraph``, ``DataLoader``, ``batch_to``, ``model(batch, q_params)``,
``RLSimulator.select_action``).  See DESIGN.md and INTEGRATION.md.
"""
from .actions import get_action_space, get_action_space_device, get_action_space_legacy  # noqa: F401
from .graph import (AtomsDataset, AtomsGraph, Batch, DataLoader, ReactionDataset, ReactionGraph,  # noqa: F401
                    batch_to)
from .registry import registry  # noqa: F401
from .structures import Atoms, make_crconi_supercell, read_poscar, write_poscar  # noqa: F401
from .timeest import estimate_time, estimate_time_device, read_xdatcar, timer_converter  # noqa: F401
from .trainer import Memory, Trainer  # noqa: F401
from .trajectory import TrajectoryWriter  # noqa: F401

__version__ = "0.1.0"
