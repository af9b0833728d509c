"""Exploration noise for evaluation-time ``select_action`` (reference: utils/noise.py).  Not on the gradient-step path."""
from typing import Optional, Sequence, Union

import numpy as np


class GaussianNoise:
    def __init__(self, mu: float = 0.0, sigma: float = 1.0) -> None:
        self._mu, self._sigma = mu, sigma
        assert sigma >= 0, "noise std must be non-negative"

    def __call__(self, size: Sequence[int]) -> np.ndarray:
        return np.random.normal(self._mu, self._sigma, size)

    def reset(self) -> None:
        pass
