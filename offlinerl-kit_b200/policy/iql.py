"""IQLPolicy facade (reference: policy/model_free/iql.py:13-139); ``learn`` runs engine/td3_iql.py:IQLLearner."""
from copy import deepcopy
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from .base_policy import BasePolicy, engine_for, learn_many as _learn_many


class IQLPolicy(BasePolicy):
    def __init__(self, actor: nn.Module, critic_q1: nn.Module, critic_q2: nn.Module, critic_v: nn.Module,
                 actor_optim: torch.optim.Optimizer, critic_q1_optim: torch.optim.Optimizer,
                 critic_q2_optim: torch.optim.Optimizer, critic_v_optim: torch.optim.Optimizer, action_space,
                 tau: float = 0.005, gamma: float = 0.99, expectile: float = 0.8, temperature: float = 0.1) -> None:
        super().__init__()
        self.actor = actor
        self.critic_q1, self.critic_q1_old = critic_q1, deepcopy(critic_q1)
        self.critic_q2, self.critic_q2_old = critic_q2, deepcopy(critic_q2)
        self.critic_q1_old.eval()
        self.critic_q2_old.eval()
        self.critic_v = critic_v
        self.actor_optim, self.critic_v_optim = actor_optim, critic_v_optim
        self.critic_q1_optim, self.critic_q2_optim = critic_q1_optim, critic_q2_optim
        self.action_space = action_space
        self._tau, self._gamma, self._expectile, self._temperature = tau, gamma, expectile, temperature
        self._engine = None

    def train(self) -> None:
        for m in (self.actor, self.critic_q1, self.critic_q2, self.critic_v):
            m.train()

    def eval(self) -> None:
        for m in (self.actor, self.critic_q1, self.critic_q2, self.critic_v):
            m.eval()

    def select_action(self, obs: np.ndarray, deterministic: bool = False) -> np.ndarray:
        if len(obs.shape) == 1:
            obs = obs.reshape(1, -1)
        with torch.no_grad():
            dist = self.actor(obs)
            action = (dist.mode() if deterministic else dist.sample()).cpu().numpy()
        return np.clip(action, self.action_space.low[0], self.action_space.high[0])

    def engine(self, batch_size: int):
        from ..engine.td3_iql import IQLLearner
        return engine_for(self, int(batch_size), lambda: IQLLearner(self, batch_size))

    def learn_many(self, buffer, n_steps: int, batch_size: int):
        """``n_steps`` x ``learn(buffer.sample(batch_size))`` behind one host synchronisation (base_policy.learn_many)."""
        return _learn_many(self, buffer, n_steps, batch_size)

    def learn(self, batch: Dict, noise=None) -> Dict[str, float]:
        return self.engine((getattr(batch, "batch_size", None) or int(batch["observations"].shape[0]))).step(batch, noise)
