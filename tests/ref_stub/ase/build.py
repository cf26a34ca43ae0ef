"""``ase.build.bulk`` for the one call the reference makes: ``bulk("Cu", 'fcc', a=a, cubic=True)`` (action.py:22)."""
import numpy as np

from . import Atoms, _NUMBERS


def bulk(name, crystalstructure=None, a=None, cubic=False):
    if crystalstructure != "fcc" or not cubic or a is None:
        raise NotImplementedError("the stub only builds the conventional cubic fcc cell")
    basis = np.array([[0.0, 0.0, 0.0], [0.0, 0.5, 0.5], [0.5, 0.0, 0.5], [0.5, 0.5, 0.0]])
    return Atoms(numbers=[_NUMBERS[name]] * 4, positions=basis * a, cell=np.eye(3) * a, pbc=True)
