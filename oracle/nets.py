"""Functional restatement of the reference networks on plain parameter dicts.

Parameter dicts use the reference's ``state_dict`` key names, so a golden
``state_dict`` can be fed in unchanged.  Test infrastructure only (see
oracle/__init__.py).
"""
import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

P = Dict[str, torch.Tensor]

LOG_SIG_MIN, LOG_SIG_MAX = -5.0, 2.0      # modules/dist_module.py:53-54
HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


# Every ReLU of the oracle goes through ``relu`` so that the GPU parity tests can swap in an implementation that records
# pre-activations within rounding of zero and forces chosen mask bits (tests/kinks.py): a ReLU network's gradient is
# only defined up to those kink decisions, and one flipped bit moves a 7936-row critic gradient by ~5e-4 (rel. L2).
_RELU_IMPL = torch.relu


def set_relu_impl(fn=None) -> None:
    global _RELU_IMPL
    _RELU_IMPL = torch.relu if fn is None else fn


def relu(x: torch.Tensor) -> torch.Tensor:
    return _RELU_IMPL(x)


def count_hidden(p: P, prefix: str) -> int:
    """Number of Linear layers in ``<prefix>.model`` (nets/mlp.py:21-24: Linear at even slots)."""
    n = 0
    while f"{prefix}.model.{2 * n}.weight" in p:
        n += 1
    return n


def mlp_relu(p: P, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """nets/mlp.py:9-33 with ReLU activations, no dropout, no output layer."""
    for i in range(count_hidden(p, prefix)):
        x = relu(F.linear(x, p[f"{prefix}.model.{2 * i}.weight"], p[f"{prefix}.model.{2 * i}.bias"]))
    return x


def critic(p: P, name: str, obs: torch.Tensor, act: torch.Tensor = None) -> torch.Tensor:
    """modules/critic_module.py:17-27: Q(s,a) (or V(s) when ``act`` is None) -> [M,1]."""
    x = obs if act is None else torch.cat([obs, act.flatten(1)], dim=1)
    h = mlp_relu(p, f"{name}.backbone", x)
    return F.linear(h, p[f"{name}.last.weight"], p[f"{name}.last.bias"])


def det_actor(p: P, name: str, obs: torch.Tensor, max_action: float = 1.0) -> torch.Tensor:
    """modules/actor_module.py:46-50: deterministic TD3 actor, max * tanh(Linear(backbone))."""
    h = mlp_relu(p, f"{name}.backbone", obs)
    return max_action * torch.tanh(F.linear(h, p[f"{name}.last.weight"], p[f"{name}.last.bias"]))


def gauss_head(p: P, name: str, obs: torch.Tensor, unbounded: bool, max_mu: float = 1.0
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """modules/actor_module.py:22-26 + dist_module.py:65-76 / 117-127 -> (mu, sigma).

    conditioned sigma  : sigma = exp(clamp(Linear(h), -5, 2))
    state-independent  : sigma = exp(sigma_param[A,1] broadcast as [1,A])
    """
    h = mlp_relu(p, f"{name}.backbone", obs)
    mu = F.linear(h, p[f"{name}.dist_net.mu.weight"], p[f"{name}.dist_net.mu.bias"])
    if not unbounded:
        mu = max_mu * torch.tanh(mu)
    if f"{name}.dist_net.sigma.weight" in p:
        raw = F.linear(h, p[f"{name}.dist_net.sigma.weight"], p[f"{name}.dist_net.sigma.bias"])
        sigma = torch.clamp(raw, min=LOG_SIG_MIN, max=LOG_SIG_MAX).exp()
    else:
        sigma = (p[f"{name}.dist_net.sigma_param"].view(1, -1) + torch.zeros_like(mu)).exp()
    return mu, sigma


def normal_logp(x: torch.Tensor, mu: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
    """torch.distributions.Normal.log_prob, summed over the action axis (dist_module.py:7-8)."""
    var = sigma ** 2
    lp = -((x - mu) ** 2) / (2 * var) - sigma.log() - HALF_LOG_2PI
    return lp.sum(-1, keepdim=True)


def tanh_gauss_sample(mu: torch.Tensor, sigma: torch.Tensor, eps: torch.Tensor = None
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
    """dist_module.py:39-42 (rsample) + :21-27 (log_prob) ; eps=None -> mode (:29-32).

    returns (squashed action [M,A], log-prob [M,1]).
    """
    raw = mu if eps is None else mu + sigma * eps
    act = torch.tanh(raw)
    logp = normal_logp(raw, mu, sigma) - torch.log((1 - act.pow(2)) + 1e-6).sum(-1, keepdim=True)
    return act, logp


def actforward(p: P, name: str, obs: torch.Tensor, eps: torch.Tensor = None):
    """policy/model_free/sac.py:66-77 for a TanhDiagGaussian(unbounded, conditioned sigma) actor."""
    mu, sigma = gauss_head(p, name, obs, unbounded=True)
    return tanh_gauss_sample(mu, sigma, eps)


def ensemble_linear(w: torch.Tensor, b: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """nets/ensemble_linear.py:30-41: shared 2-D input or per-member 3-D input."""
    if x.dim() == 2:
        y = torch.einsum("ij,bjk->bik", x, w)
    else:
        y = torch.einsum("bij,bjk->bik", x, w)
    return y + b


def ensemble_critic(p: P, name: str, obs: torch.Tensor, act: torch.Tensor) -> torch.Tensor:
    """modules/ensemble_critic_module.py:33-43 -> [E, M, 1]."""
    x = torch.cat([obs, act], dim=-1)
    n = 0
    while f"{name}.model.{2 * n}.weight" in p:
        n += 1
    for i in range(n):
        x = ensemble_linear(p[f"{name}.model.{2 * i}.weight"], p[f"{name}.model.{2 * i}.bias"], x)
        if i < n - 1:
            x = relu(x)
    return x


def soft_clamp(x, lo, hi):
    """modules/dynamics_module.py:19-29."""
    x = hi - F.softplus(hi - x)
    x = lo + F.softplus(x - lo)
    return x


def dynamics_forward(p: P, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """modules/dynamics_module.py:87-94 -> (mean, logvar), each [E, M, obs+1]."""
    n = 0
    while f"backbones.{n}.weight" in p:
        n += 1
    h = x
    for i in range(n):
        z = ensemble_linear(p[f"backbones.{i}.weight"], p[f"backbones.{i}.bias"], h)
        h = z * torch.sigmoid(z)                      # Swish, dynamics_module.py:14-16
    out = ensemble_linear(p["output_layer.weight"], p["output_layer.bias"], h)
    mean, raw = torch.chunk(out, 2, dim=-1)
    return mean, soft_clamp(raw, p["min_logvar"], p["max_logvar"])
