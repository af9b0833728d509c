"""SACPolicy facade (reference: policy/model_free/sac.py:10-140).

Holds the same attributes as the reference class (the run scripts and the trainers only touch those) and
delegates ``learn`` to the CUDA engine (engine/sac_family.py).  Used directly by MOPO.
"""
from copy import deepcopy
from typing import Dict, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from .base_policy import BasePolicy, engine_for, learn_many as _learn_many


class SACPolicy(BasePolicy):
    _engine_cls = None      # set below (import cycle-free)

    def __init__(self, actor: nn.Module, critic1: nn.Module, critic2: nn.Module,
                 actor_optim: torch.optim.Optimizer, critic1_optim: torch.optim.Optimizer,
                 critic2_optim: torch.optim.Optimizer, tau: float = 0.005, gamma: float = 0.99,
                 alpha: Union[float, Tuple[float, torch.Tensor, torch.optim.Optimizer]] = 0.2) -> None:
        super().__init__()
        self.actor = actor
        self.critic1, self.critic1_old = critic1, deepcopy(critic1)
        self.critic2, self.critic2_old = critic2, deepcopy(critic2)
        self.critic1_old.eval()
        self.critic2_old.eval()
        self.actor_optim, self.critic1_optim, self.critic2_optim = actor_optim, critic1_optim, critic2_optim
        self._tau, self._gamma = tau, gamma
        self._is_auto_alpha = isinstance(alpha, tuple)
        if self._is_auto_alpha:
            self._target_entropy, self._log_alpha, self.alpha_optim = alpha
            self._alpha = self._log_alpha.detach().exp()
        else:
            self._alpha = alpha
        self._engine = None

    def train(self) -> None:
        for m in (self.actor, self.critic1, self.critic2):
            m.train()

    def eval(self) -> None:
        for m in (self.actor, self.critic1, self.critic2):
            m.eval()

    # ---- inference (evaluation / rollouts)
    def actforward(self, obs: torch.Tensor, deterministic: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        dist = self.actor(obs)
        squashed, raw = dist.mode() if deterministic else dist.rsample()
        return squashed, dist.log_prob(squashed, raw)

    def select_action(self, obs: np.ndarray, deterministic: bool = False) -> np.ndarray:
        with torch.no_grad():
            action, _ = self.actforward(obs, deterministic)
        return action.cpu().numpy()

    # ---- the gradient step
    def _make_engine(self, batch_size: int):
        from ..engine.sac_family import SACLearner
        return SACLearner(self, batch_size)

    def engine(self, batch_size: int):
        return engine_for(self, int(batch_size), lambda: self._make_engine(batch_size))

    def _after_step(self, out: Dict[str, float]) -> None:
        if self._is_auto_alpha and "alpha" in out:
            self._alpha_value = out["alpha"]        # the tensor form is rebuilt on demand (see the _alpha property)
            self._alpha_tensor = None

    # ``_alpha`` is a tensor in the reference (sac.py:43-49); building a device tensor per step costs more host time
    # than the rest of ``learn``, so the engine keeps the float and the tensor is materialised when somebody reads it.
    @property
    def _alpha(self):
        if self._alpha_tensor is None:
            self._alpha_tensor = torch.tensor([self._alpha_value], device=self.actor.device)
        return self._alpha_tensor

    @_alpha.setter
    def _alpha(self, value) -> None:
        if torch.is_tensor(value):
            self._alpha_tensor, self._alpha_value = value, None
        else:
            self._alpha_tensor, self._alpha_value = None, float(value)
            if not self._is_auto_alpha:
                self._alpha_tensor = value      # a plain float alpha stays a float, as in the reference

    def _learn_mixed(self, real, fake, noise=None) -> Dict[str, float]:
        """The model-based policies' step on real + model-buffer rows (mopo.py:81-84, combo.py:110-112).  Draws of two
        device-mirrored buffers are gathered straight into the step's batch inside the step graph; anything else is
        concatenated as the reference does."""
        tr, tf = getattr(real, "token", None), getattr(fake, "token", None)
        if tr is not None and tf is not None and tr is not tf:
            eng = self.engine(real.batch_size + fake.batch_size)
            out = eng.step([real, fake], noise)
            self._after_step(out)
            return out
        mix = {k: torch.cat([real[k], fake[k]], 0) for k in real.keys()}
        return SACPolicy.learn(self, mix, noise)

    def learn_many(self, buffer, n_steps: int, batch_size: int):
        """``n_steps`` x ``learn(buffer.sample(batch_size))`` behind one host synchronisation (base_policy.learn_many)."""
        return _learn_many(self, buffer, n_steps, batch_size)

    def learn(self, batch: Dict, noise: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, float]:
        eng = self.engine((getattr(batch, "batch_size", None) or int(batch["observations"].shape[0])))
        out = eng.step(batch, noise)
        self._after_step(out)
        return out
