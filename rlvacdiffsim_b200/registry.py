"""``rgnn.common.registry`` look-alike (``rlsim/drl/deploy.py:18``, ``train.py:50-52``, ``time/train.py:40-48``)."""
from __future__ import annotations

from .model import DQNv2, PaiNN, TNet, TNetBinary


class _Registry:
    _models = {"dqn_v2": DQNv2, "t_net": TNet, "t_net_binary": TNetBinary}
    _reaction_models = {"painn": PaiNN}

    def get_model_class(self, name: str):
        try:
            return self._models[name]
        except KeyError:
            raise KeyError(f"model '{name}' is not available in rlvacdiffsim_b200 (have {sorted(self._models)})")

    def get_reaction_model_class(self, name: str):
        try:
            return self._reaction_models[name]
        except KeyError:
            raise KeyError(f"reaction model '{name}' is not available (have {sorted(self._reaction_models)})")


registry = _Registry()
