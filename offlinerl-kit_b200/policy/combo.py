"""COMBOPolicy facade (reference: policy/model_based/combo.py:16-243).

``learn`` is the CQL step over the concatenated real + fake batch with the conservative term drawn from the mix or from
the model rows alone (``rho_s``) and its data term taken over the real rows (engine/sac_family.py:CQLLearner with
``n_real`` / ``cons_rows``); ``rollout`` is MOPO's device-resident imagination loop, with uniform actions instead of
the actor's when ``uniform_rollout`` is set (combo.py:81-88)."""
from typing import Dict, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from .base_policy import engine_for
from .cql import CQLPolicy


class COMBOPolicy(CQLPolicy):
    def __init__(self, dynamics, actor: nn.Module, critic1: nn.Module, critic2: nn.Module,
                 actor_optim: torch.optim.Optimizer, critic1_optim: torch.optim.Optimizer,
                 critic2_optim: torch.optim.Optimizer, action_space, tau: float = 0.005, gamma: float = 0.99,
                 alpha: Union[float, Tuple[float, torch.Tensor, torch.optim.Optimizer]] = 0.2,
                 cql_weight: float = 1.0, temperature: float = 1.0, max_q_backup: bool = False,
                 deterministic_backup: bool = True, with_lagrange: bool = True, lagrange_threshold: float = 10.0,
                 cql_alpha_lr: float = 1e-4, num_repeart_actions: int = 10, uniform_rollout: bool = False,
                 rho_s: str = "mix") -> None:
        super().__init__(actor, critic1, critic2, actor_optim, critic1_optim, critic2_optim, action_space, tau=tau,
                         gamma=gamma, alpha=alpha, cql_weight=cql_weight, temperature=temperature,
                         max_q_backup=max_q_backup, deterministic_backup=deterministic_backup,
                         with_lagrange=with_lagrange, lagrange_threshold=lagrange_threshold, cql_alpha_lr=cql_alpha_lr,
                         num_repeart_actions=num_repeart_actions)
        self.dynamics = dynamics
        self._uniform_rollout = uniform_rollout
        self._rho_s = rho_s
        self._roll = None
        self._split = None      # (n_real, n_fake) the step graph was built for

    device_rollouts = True      # rollout(..., device_out=True) returns CUDA tensors

    def rollout(self, init_obss: np.ndarray, rollout_length: int, noise: Optional[Dict[str, np.ndarray]] = None,
                device_out: bool = False) -> Tuple[Dict[str, np.ndarray], Dict]:
        """noise (parity tests): per-step lists ``eps`` [S_t, A] (or ``actions`` [S_t, A] with uniform_rollout),
        ``normal`` [E, S_t, D] float64, ``midx`` [S_t]."""
        from ..engine.rollout import RolloutEngine
        if self._roll is None:
            uniform = None
            if self._uniform_rollout:
                uniform = (float(self.action_space.low[0]), float(self.action_space.high[0]))
            self._roll = RolloutEngine(self, uniform=uniform)
        return self._roll.run(np.asarray(init_obss, np.float32), int(rollout_length), noise, device_out)

    def _rows(self):
        n_real, n_fake = self._split
        return dict(n_real=n_real, cons_rows=(n_real, n_real + n_fake) if self._rho_s == "model" else None)

    def _make_engine(self, batch_size: int):
        from ..engine.sac_family import CQLLearner
        return CQLLearner(self, batch_size, **self._rows())

    def engine(self, batch_size: int):
        # keyed by the (real, fake) split: another split is another step graph over the same parameters
        return engine_for(self, tuple(self._split), lambda: self._make_engine(batch_size), **self._rows())

    def learn(self, batch: Dict, noise=None) -> Dict[str, float]:
        real, fake = batch["real"], batch["fake"]
        size = lambda b: getattr(b, "batch_size", None) or int(b["observations"].shape[0])   # no gather forced
        split = (size(real), size(fake))
        if split[0] == 0 or (self._rho_s == "model" and split[1] == 0):
            raise ValueError("COMBO needs real rows (and model rows with rho_s='model') in every batch")
        self._split = split
        return self._learn_mixed(real, fake, noise)
