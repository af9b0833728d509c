"""TD3Policy / TD3BCPolicy facades (reference: policy/model_free/td3.py:11-127, td3bc.py:13-124)."""
from copy import deepcopy
from typing import Callable, Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from .base_policy import BasePolicy, engine_for, learn_many as _learn_many
from ..utils.noise import GaussianNoise


class TD3BCPolicy(BasePolicy):
    def __init__(self, actor: nn.Module, critic1: nn.Module, critic2: nn.Module, actor_optim: torch.optim.Optimizer,
                 critic1_optim: torch.optim.Optimizer, critic2_optim: torch.optim.Optimizer, tau: float = 0.005,
                 gamma: float = 0.99, max_action: float = 1.0, exploration_noise: Callable = GaussianNoise,
                 policy_noise: float = 0.2, noise_clip: float = 0.5, update_actor_freq: int = 2, alpha: float = 2.5,
                 scaler=None) -> None:
        super().__init__()
        self.actor, self.actor_old = actor, deepcopy(actor)
        self.critic1, self.critic1_old = critic1, deepcopy(critic1)
        self.critic2, self.critic2_old = critic2, deepcopy(critic2)
        for m in (self.actor_old, self.critic1_old, self.critic2_old):
            m.eval()
        self.actor_optim, self.critic1_optim, self.critic2_optim = actor_optim, critic1_optim, critic2_optim
        self._tau, self._gamma, self._max_action = tau, gamma, max_action
        self.exploration_noise = exploration_noise
        self._policy_noise, self._noise_clip, self._freq = policy_noise, noise_clip, update_actor_freq
        self._cnt, self._last_actor_loss = 0, 0
        self._alpha, self.scaler = alpha, scaler
        self._engine = None

    def train(self) -> None:
        for m in (self.actor, self.critic1, self.critic2):
            m.train()

    def eval(self) -> None:
        for m in (self.actor, self.critic1, self.critic2):
            m.eval()

    def select_action(self, obs: np.ndarray, deterministic: bool = False) -> np.ndarray:
        if self.scaler is not None:
            obs = self.scaler.transform(obs)
        with torch.no_grad():
            action = self.actor(obs).cpu().numpy()
        if not deterministic:
            action = np.clip(action + self.exploration_noise(action.shape), -self._max_action, self._max_action)
        return action

    def engine(self, batch_size: int):
        from ..engine.td3_iql import TD3BCLearner
        return engine_for(self, int(batch_size), lambda: TD3BCLearner(self, batch_size))

    def learn_many(self, buffer, n_steps: int, batch_size: int):
        """``n_steps`` x ``learn(buffer.sample(batch_size))`` behind one host synchronisation (base_policy.learn_many)."""
        return _learn_many(self, buffer, n_steps, batch_size)

    def learn(self, batch: Dict, noise: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, float]:
        return self.engine((getattr(batch, "batch_size", None) or int(batch["observations"].shape[0]))).step(batch, noise)
