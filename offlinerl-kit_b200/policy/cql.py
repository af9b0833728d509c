"""CQLPolicy facade (reference: policy/model_free/cql.py:16-207); ``learn`` runs engine/sac_family.py:CQLLearner."""
from typing import Dict, Tuple, Union

import torch
import torch.nn as nn

from .sac import SACPolicy


class CQLPolicy(SACPolicy):
    def __init__(self, actor: nn.Module, critic1: nn.Module, critic2: nn.Module,
                 actor_optim: torch.optim.Optimizer, critic1_optim: torch.optim.Optimizer,
                 critic2_optim: torch.optim.Optimizer, action_space, tau: float = 0.005, gamma: float = 0.99,
                 alpha: Union[float, Tuple[float, torch.Tensor, torch.optim.Optimizer]] = 0.2,
                 cql_weight: float = 1.0, temperature: float = 1.0, max_q_backup: bool = False,
                 deterministic_backup: bool = True, with_lagrange: bool = True, lagrange_threshold: float = 10.0,
                 cql_alpha_lr: float = 1e-4, num_repeart_actions: int = 10) -> None:
        super().__init__(actor, critic1, critic2, actor_optim, critic1_optim, critic2_optim, tau=tau, gamma=gamma,
                         alpha=alpha)
        self.action_space = action_space
        self._cql_weight, self._temperature = cql_weight, temperature
        self._max_q_backup, self._deterministic_backup = max_q_backup, deterministic_backup
        self._with_lagrange, self._lagrange_threshold = with_lagrange, lagrange_threshold
        # plain tensors, not registered parameters -- exactly as cql.py:57-58 (they are not in state_dict())
        self.cql_log_alpha = torch.zeros(1, requires_grad=True, device=self.actor.device)
        self.cql_alpha_optim = torch.optim.Adam([self.cql_log_alpha], lr=cql_alpha_lr)
        self._num_repeat_actions = num_repeart_actions      # (sic) keyword spelled as in the reference

    def _make_engine(self, batch_size: int):
        from ..engine.sac_family import CQLLearner
        return CQLLearner(self, batch_size)
