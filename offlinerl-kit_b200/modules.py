"""Actor / critic / distribution-head / dynamics-model containers (reference: offlinerlkit/modules/*).

Same class names, constructor arguments, attribute names (``backbone``, ``last``, ``dist_net.mu`` / ``.sigma`` /
``.sigma_param``, ``model``, ``backbones``, ``output_layer``, ``max_logvar`` ...) and ``state_dict`` keys as the
reference, because the run scripts reach into them (run_edac.py:100-103, run_iql.py:121-125) and checkpoints are
``state_dict`` dumps.  ``forward`` is a plain-torch inference path for evaluation; the gradient step and model
rollouts run in the CUDA engine.
"""
import math
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .nets import EnsembleLinear

ArrayLike = Union[np.ndarray, torch.Tensor]


def _as_f32(x: ArrayLike, device: torch.device) -> torch.Tensor:
    return torch.as_tensor(x, device=device, dtype=torch.float32)


# ------------------------------------------------------------------------------------------- distributions
class _DiagNormal:
    """Diagonal Gaussian with the reference's summed log-prob (dist_module.py:6-14)."""

    squash = False

    def __init__(self, mu: torch.Tensor, sigma: torch.Tensor):
        self.mean, self.stddev = mu, sigma

    def _raw_log_prob(self, x: torch.Tensor) -> torch.Tensor:
        var = self.stddev ** 2
        return -((x - self.mean) ** 2) / (2 * var) - self.stddev.log() - 0.5 * math.log(2 * math.pi)

    def log_prob(self, actions: torch.Tensor) -> torch.Tensor:
        return self._raw_log_prob(actions).sum(-1, keepdim=True)

    def mode(self):
        return self.mean

    def rsample(self):
        return self.mean + self.stddev * torch.randn_like(self.mean)

    def sample(self):
        with torch.no_grad():
            return self.rsample()

    def entropy(self):
        return (0.5 + 0.5 * math.log(2 * math.pi) + self.stddev.log()).sum(-1)


class _TanhDiagNormal(_DiagNormal):
    """tanh-squashed Gaussian: rsample / mode return (action, pre-tanh action) (dist_module.py:17-42)."""

    squash = True

    def log_prob(self, action: torch.Tensor, raw_action: Optional[torch.Tensor] = None) -> torch.Tensor:
        if raw_action is None:
            hi, lo = (1 + action).clamp(min=1e-6), (1 - action).clamp(min=1e-6)
            raw_action = 0.5 * torch.log(hi / lo)
        lp = self._raw_log_prob(raw_action).sum(-1, keepdim=True)
        return lp - torch.log((1 - action.pow(2)) + 1e-6).sum(-1, keepdim=True)

    def mode(self):
        return torch.tanh(self.mean), self.mean

    def rsample(self):
        raw = super().rsample()
        return torch.tanh(raw), raw


class DiagGaussian(nn.Module):
    _dist = _DiagNormal

    def __init__(self, latent_dim, output_dim, unbounded=False, conditioned_sigma=False, max_mu=1.0,
                 sigma_min=-5.0, sigma_max=2.0):
        super().__init__()
        self.mu = nn.Linear(latent_dim, output_dim)
        self._c_sigma = conditioned_sigma
        if conditioned_sigma:
            self.sigma = nn.Linear(latent_dim, output_dim)
        else:
            self.sigma_param = nn.Parameter(torch.zeros(output_dim, 1))
        self._unbounded, self._max = unbounded, max_mu
        self._sigma_min, self._sigma_max = sigma_min, sigma_max

    def get_dist_params(self, logits: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        mu = self.mu(logits)
        if not self._unbounded:
            mu = self._max * torch.tanh(mu)
        if self._c_sigma:
            log_sigma = self.sigma(logits).clamp(self._sigma_min, self._sigma_max)
        else:
            log_sigma = self.sigma_param.view(1, -1).expand_as(mu)
        return mu, log_sigma

    def forward(self, logits: torch.Tensor):
        mu, log_sigma = self.get_dist_params(logits)
        return self._dist(mu, log_sigma.exp())


class TanhDiagGaussian(DiagGaussian):
    _dist = _TanhDiagNormal


# ------------------------------------------------------------------------------------------- actors / critics
class ActorProb(nn.Module):
    def __init__(self, backbone: nn.Module, dist_net: nn.Module, device: str = "cpu") -> None:
        super().__init__()
        self.device = torch.device(device)
        self.backbone = backbone.to(device)
        self.dist_net = dist_net.to(device)

    def forward(self, obs: ArrayLike):
        return self.dist_net(self.backbone(_as_f32(obs, self.device)))


class Actor(nn.Module):
    def __init__(self, backbone: nn.Module, action_dim: int, max_action: float = 1.0, device: str = "cpu") -> None:
        super().__init__()
        self.device = torch.device(device)
        self.backbone = backbone.to(device)
        self.last = nn.Linear(getattr(backbone, "output_dim"), action_dim).to(device)
        self._max = max_action

    def forward(self, obs: ArrayLike) -> torch.Tensor:
        return self._max * torch.tanh(self.last(self.backbone(_as_f32(obs, self.device))))


class Critic(nn.Module):
    def __init__(self, backbone: nn.Module, device: str = "cpu") -> None:
        super().__init__()
        self.device = torch.device(device)
        self.backbone = backbone.to(device)
        self.last = nn.Linear(getattr(backbone, "output_dim"), 1).to(device)

    def forward(self, obs: ArrayLike, actions: Optional[ArrayLike] = None) -> torch.Tensor:
        x = _as_f32(obs, self.device)
        if actions is not None:
            x = torch.cat([x, _as_f32(actions, self.device).flatten(1)], dim=1)
        return self.last(self.backbone(x))


class EnsembleCritic(nn.Module):
    def __init__(self, obs_dim: int, action_dim: int, hidden_dims: Sequence[int], activation: type = nn.ReLU,
                 num_ensemble: int = 10, device: str = "cpu") -> None:
        super().__init__()
        widths = [int(obs_dim) + int(action_dim)] + [int(h) for h in hidden_dims]
        stack: List[nn.Module] = []
        for fan_in, fan_out in zip(widths, widths[1:]):
            stack += [EnsembleLinear(fan_in, fan_out, num_ensemble), activation()]
        stack.append(EnsembleLinear(widths[-1], 1, num_ensemble))
        self.device = torch.device(device)
        self.model = nn.Sequential(*stack).to(device)
        self._num_ensemble = num_ensemble
        self.activation_type = activation

    def forward(self, obs: ArrayLike, actions: Optional[ArrayLike] = None) -> torch.Tensor:
        x = _as_f32(obs, self.device)
        if actions is not None:
            x = torch.cat([x, _as_f32(actions, self.device)], dim=-1)
        return self.model(x)


# ------------------------------------------------------------------------------------------- dynamics model
class Swish(nn.Module):
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x * torch.sigmoid(x)


def soft_clamp(x: torch.Tensor, _min: Optional[torch.Tensor] = None, _max: Optional[torch.Tensor] = None):
    if _max is not None:
        x = _max - F.softplus(_max - x)
    if _min is not None:
        x = _min + F.softplus(x - _min)
    return x


class EnsembleDynamicsModel(nn.Module):
    def __init__(self, obs_dim: int, action_dim: int, hidden_dims: Sequence[int], num_ensemble: int = 7,
                 num_elites: int = 5, activation: type = Swish, weight_decays: Optional[Sequence[float]] = None,
                 with_reward: bool = True, device: str = "cpu") -> None:
        super().__init__()
        self.num_ensemble, self.num_elites = num_ensemble, num_elites
        self._with_reward = with_reward
        self.device = torch.device(device)
        self.activation = activation()
        if weight_decays is None:
            weight_decays = [0.0] * (len(hidden_dims) + 1)
        assert len(weight_decays) == len(hidden_dims) + 1
        widths = [int(obs_dim) + int(action_dim)] + [int(h) for h in hidden_dims]
        self.backbones = nn.ModuleList(
            EnsembleLinear(i, o, num_ensemble, wd) for i, o, wd in zip(widths, widths[1:], weight_decays))
        out_dim = int(obs_dim) + int(with_reward)
        self.output_layer = EnsembleLinear(widths[-1], 2 * out_dim, num_ensemble, weight_decays[-1])
        self.max_logvar = nn.Parameter(torch.ones(out_dim) * 0.5, requires_grad=True)
        self.min_logvar = nn.Parameter(torch.ones(out_dim) * -10, requires_grad=True)
        self.elites = nn.Parameter(torch.tensor(list(range(0, num_elites))), requires_grad=False)
        self.to(self.device)

    def forward(self, obs_action: ArrayLike) -> Tuple[torch.Tensor, torch.Tensor]:
        h = torch.as_tensor(obs_action, dtype=torch.float32).to(self.device)
        for layer in self.backbones:
            h = self.activation(layer(h))
        mean, logvar = torch.chunk(self.output_layer(h), 2, dim=-1)
        return mean, soft_clamp(logvar, self.min_logvar, self.max_logvar)

    def _all_layers(self):
        return list(self.backbones) + [self.output_layer]

    def load_save(self) -> None:
        for layer in self._all_layers():
            layer.load_save()

    def update_save(self, indexes: List[int]) -> None:
        for layer in self._all_layers():
            layer.update_save(indexes)

    def get_decay_loss(self) -> torch.Tensor:
        return sum(layer.get_decay_loss() for layer in self._all_layers())

    def set_elites(self, indexes: List[int]) -> None:
        assert len(indexes) <= self.num_ensemble and max(indexes) < self.num_ensemble
        self.register_parameter("elites", nn.Parameter(torch.tensor(indexes), requires_grad=False))

    def random_elite_idxs(self, batch_size: int) -> np.ndarray:
        return np.random.choice(self.elites.data.cpu().numpy(), size=batch_size)
