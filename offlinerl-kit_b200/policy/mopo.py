"""MOPOPolicy facade (reference: policy/model_based/mopo.py:13-84).

``learn`` is SAC on the concatenated real + fake batch; ``rollout`` keeps the whole imagination loop on the device
(actor forward, scaler, ensemble forward, elite pick, sampling, termination, penalty, stable survivor compaction)
and copies the concatenated transitions to the host once, in the reference's output format."""
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .. import _lib as L
from .sac import SACPolicy


class MOPOPolicy(SACPolicy):
    def __init__(self, dynamics, actor, critic1, critic2, actor_optim, critic1_optim, critic2_optim, tau: float = 0.005,
                 gamma: float = 0.99, alpha=0.2) -> None:
        super().__init__(actor, critic1, critic2, actor_optim, critic1_optim, critic2_optim, tau=tau, gamma=gamma, alpha=alpha)
        self.dynamics = dynamics
        self._roll = None

    device_rollouts = True      # rollout(..., device_out=True) returns CUDA tensors

    def rollout(self, init_obss: np.ndarray, rollout_length: int, noise: Optional[Dict[str, np.ndarray]] = None,
                device_out: bool = False) -> Tuple[Dict[str, np.ndarray], Dict]:
        """noise (parity tests): per-step lists ``eps`` [S_t, A], ``normal`` [E, S_t, D] float64, ``midx`` [S_t]."""
        from ..engine.rollout import RolloutEngine
        if self._roll is None:
            self._roll = RolloutEngine(self)
        return self._roll.run(np.asarray(init_obss, np.float32), int(rollout_length), noise, device_out)

    def learn(self, batch: Dict, noise=None) -> Dict[str, float]:
        return self._learn_mixed(batch["real"], batch["fake"], noise)
