"""EDACPolicy facade (reference: policy/model_free/edac.py:11-166); ``learn`` runs engine/edac.py:EDACLearner."""
from copy import deepcopy
from typing import Dict, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from .base_policy import BasePolicy, engine_for, learn_many as _learn_many


class EDACPolicy(BasePolicy):
    def __init__(self, actor: nn.Module, critics: nn.Module, actor_optim: torch.optim.Optimizer,
                 critics_optim: torch.optim.Optimizer, tau: float = 0.005, gamma: float = 0.99,
                 alpha: Union[float, Tuple[float, torch.Tensor, torch.optim.Optimizer]] = 0.2, max_q_backup: bool = False,
                 deterministic_backup: bool = True, eta: float = 1.0) -> None:
        super().__init__()
        self.actor = actor
        self.critics, self.critics_old = critics, deepcopy(critics)
        self.critics_old.eval()
        self.actor_optim, self.critics_optim = actor_optim, critics_optim
        self._tau, self._gamma = tau, gamma
        self._is_auto_alpha = isinstance(alpha, tuple)
        if self._is_auto_alpha:
            self._target_entropy, self._log_alpha, self.alpha_optim = alpha
            self._alpha = self._log_alpha.detach().exp()
        else:
            self._alpha = alpha
        self._max_q_backup, self._deterministic_backup, self._eta = max_q_backup, deterministic_backup, eta
        self._num_critics = self.critics._num_ensemble
        self._engine = None

    def train(self) -> None:
        self.actor.train()
        self.critics.train()

    def eval(self) -> None:
        self.actor.eval()
        self.critics.eval()

    def actforward(self, obs: torch.Tensor, deterministic: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        dist = self.actor(obs)
        squashed, raw = dist.mode() if deterministic else dist.rsample()
        return squashed, dist.log_prob(squashed, raw)

    def select_action(self, obs: np.ndarray, deterministic: bool = False) -> np.ndarray:
        with torch.no_grad():
            action, _ = self.actforward(obs, deterministic)
        return action.cpu().numpy()

    def shard_critics(self, rank: int, world: int, comm=None) -> None:
        """Train only members partition_members(E, world)[rank] of the critic ensemble on this rank (BASELINE.json
        configs[2]: "members sharded across GPUs"); actor, alpha, batch and noise must be replicated (same seeds).  ``comm``
        = engine.edac_sharded.NcclComm() between processes; None when the ranks are driven by EmulatedShardGroup.  Call
        before the first ``learn``; ``gather_critics()`` makes ``state_dict()`` whole again."""
        if self._engine is not None:
            raise RuntimeError("shard_critics must be called before the first learn()")
        self._shard = (int(rank), int(world), comm)

    def gather_critics(self) -> None:
        if self._engine is not None and hasattr(self._engine, "gather_all"):
            self._engine.gather_all()

    def engine(self, batch_size: int):
        shard = getattr(self, "_shard", None)
        if shard is not None:
            from ..engine.edac_sharded import EDACShardedLearner
            return engine_for(self, int(batch_size), lambda: EDACShardedLearner(self, batch_size, *shard))
        from ..engine.edac import EDACLearner
        return engine_for(self, int(batch_size), lambda: EDACLearner(self, batch_size))

    def learn_many(self, buffer, n_steps: int, batch_size: int):
        """``n_steps`` x ``learn(buffer.sample(batch_size))`` behind one host synchronisation (base_policy.learn_many)."""
        return _learn_many(self, buffer, n_steps, batch_size)

    def learn(self, batch: Dict, noise: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, float]:
        out = self.engine((getattr(batch, "batch_size", None) or int(batch["observations"].shape[0]))).step(batch, noise)
        if self._is_auto_alpha:
            self._alpha = torch.tensor([out["alpha"]], device=self.actor.device)
        return out
