"""Abstract policy contract the unchanged trainers rely on (reference: policy/base_policy.py:8-26)."""
from typing import Dict

import numpy as np
import torch.nn as nn


class BasePolicy(nn.Module):
    def train(self) -> None:                                     # noqa: D401  (signature as used by the trainers)
        raise NotImplementedError

    def eval(self) -> None:
        raise NotImplementedError

    def select_action(self, obs: np.ndarray, deterministic: bool = False) -> np.ndarray:
        raise NotImplementedError

    def learn(self, batch: Dict) -> Dict[str, float]:
        raise NotImplementedError
