"""CPU oracle for the MOPO ensemble dynamics path.  TEST INFRASTRUCTURE ONLY.

Follows dynamics/ensemble_dynamics.py (learn :178-208, validate :210-217,
step :28-79), modules/dynamics_module.py, utils/scaler.py and
utils/termination_fns.py of the reference.
"""
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import nets

Tensors = Dict[str, torch.Tensor]


class DynamicsOracle:
    """Gaussian-NLL training step(s) and one-step imagination of ``EnsembleDynamics``."""

    LAYERS = ("backbones.0", "backbones.1", "backbones.2", "backbones.3", "output_layer")

    def __init__(self, state: Tensors, weight_decays, lr: float = 1e-3):
        self.p: Tensors = {}
        for k, v in state.items():
            t = torch.as_tensor(v).detach().clone()
            train = t.is_floating_point() and "saved_" not in k and k != "elites"
            self.p[k] = (t.float() if t.is_floating_point() else t).requires_grad_(train)
        self.layers = [n for n in self.LAYERS if f"{n}.weight" in self.p]
        self.wd = list(weight_decays)
        assert len(self.wd) == len(self.layers)
        self.names = [k for k, v in self.p.items() if v.requires_grad]
        self.optim = torch.optim.Adam([self.p[n] for n in self.names], lr=lr)
        self.grads: Tensors = {}

    def decay_loss(self) -> torch.Tensor:
        """dynamics_module.py:106-111 + ensemble_linear.py:51-53."""
        return sum(wd * (0.5 * (self.p[f"{n}.weight"] ** 2).sum()) for n, wd in zip(self.layers, self.wd))

    def loss(self, x: torch.Tensor, y: torch.Tensor, logvar_loss_coef: float = 0.01) -> torch.Tensor:
        """ensemble_dynamics.py:193-201 for one [E, b, in] / [E, b, out] mini-batch."""
        mean, logvar = nets.dynamics_forward(self.p, x)
        inv_var = torch.exp(-logvar)
        mse_inv = (torch.pow(mean - y, 2) * inv_var).mean(dim=(1, 2))
        var_loss = logvar.mean(dim=(1, 2))
        loss = mse_inv.sum() + var_loss.sum()
        loss = loss + self.decay_loss()
        return loss + logvar_loss_coef * self.p["max_logvar"].sum() - logvar_loss_coef * self.p["min_logvar"].sum()

    def learn_batch(self, x, y, logvar_loss_coef: float = 0.01) -> float:
        loss = self.loss(torch.as_tensor(x, dtype=torch.float32), torch.as_tensor(y, dtype=torch.float32),
                         logvar_loss_coef)
        self.optim.zero_grad()
        loss.backward()
        self.grads = {n: self.p[n].grad.detach().clone() for n in self.names if self.p[n].grad is not None}
        self.optim.step()
        return loss.item()

    def learn(self, inputs: np.ndarray, targets: np.ndarray, batch_size: int = 256,
              logvar_loss_coef: float = 0.01) -> float:
        """ensemble_dynamics.py:178-208: one pass over [E, n, .] in slices of ``batch_size``."""
        n = inputs.shape[1]
        losses = []
        for b in range(int(np.ceil(n / batch_size))):
            sl = slice(b * batch_size, (b + 1) * batch_size)
            losses.append(self.learn_batch(inputs[:, sl], targets[:, sl], logvar_loss_coef))
        return float(np.mean(losses))

    @torch.no_grad()
    def validate(self, inputs: np.ndarray, targets: np.ndarray) -> List[float]:
        """ensemble_dynamics.py:210-217: per-member holdout MSE of the mean head."""
        mean, _ = nets.dynamics_forward(self.p, torch.as_tensor(inputs, dtype=torch.float32))
        return list(((mean - torch.as_tensor(targets, dtype=torch.float32)) ** 2).mean(dim=(1, 2)).numpy())

    @torch.no_grad()
    def step(self, obs: np.ndarray, action: np.ndarray, mu: np.ndarray, std: np.ndarray,
             terminal_fn: Callable, penalty_coef: float, normal_noise: np.ndarray,
             model_idxs: np.ndarray, uncertainty_mode: str = "aleatoric") -> Tuple[np.ndarray, np.ndarray, np.ndarray, Dict]:
        """ensemble_dynamics.py:28-79 with the two NumPy draws passed in.

        ``normal_noise`` = np.random.normal(size=[E,B,D]) (float64, :48);
        ``model_idxs``   = np.random.choice(elites, B) (dynamics_module.py:118).
        """
        obs_act = np.concatenate([obs, action], axis=-1)
        obs_act = (obs_act - mu) / std                                   # scaler.py:31
        mean, logvar = nets.dynamics_forward(self.p, torch.as_tensor(obs_act, dtype=torch.float32))
        mean, logvar = mean.numpy(), logvar.numpy()
        mean[..., :-1] += obs
        sd = np.sqrt(np.exp(logvar))
        samples_all = (mean + normal_noise * sd).astype(np.float32)
        B = samples_all.shape[1]
        samples = samples_all[model_idxs, np.arange(B)]
        next_obs, reward = samples[..., :-1], samples[..., -1:]
        terminal = terminal_fn(obs, action, next_obs)
        info = {"raw_reward": reward}
        if penalty_coef:
            if uncertainty_mode == "aleatoric":                          # :60-62
                penalty = np.amax(np.linalg.norm(sd, axis=2), axis=0)
            elif uncertainty_mode == "pairwise-diff":                    # :63-67
                m = mean[..., :-1]
                penalty = np.amax(np.linalg.norm(m - np.mean(m, axis=0), axis=2), axis=0)
            elif uncertainty_mode == "ensemble_std":                     # :68-70
                penalty = np.sqrt(mean[..., :-1].var(0).mean(1))
            else:
                raise ValueError(uncertainty_mode)
            penalty = np.expand_dims(penalty, 1).astype(np.float32)
            reward = reward - penalty_coef * penalty
            info["penalty"] = penalty
        return next_obs, reward, terminal, info


@torch.no_grad()
def sample_next_obss(ora: "DynamicsOracle", obs: np.ndarray, action: np.ndarray, mu: np.ndarray, std: np.ndarray,
                     elites: np.ndarray, noise: np.ndarray) -> np.ndarray:
    """dynamics/ensemble_dynamics.py:81-99 (MOBILE's uncertainty samples): every elite member's Gaussian prediction of the
    next observation, ``num_samples`` draws each.  ``noise`` [num_samples, n_elites, B, D] = the reference's
    ``torch.randn_like(std)`` draws in order (one per sample).  Returns [num_samples, n_elites, B, obs_dim]."""
    obs_act = np.concatenate([obs, action], axis=-1)
    obs_act = (obs_act - mu) / std                                       # scaler.transform_tensor (scaler.py:34-38)
    mean, logvar = nets.dynamics_forward(ora.p, torch.as_tensor(obs_act, dtype=torch.float32))
    mean = mean.clone()
    mean[..., :-1] += torch.as_tensor(obs, dtype=torch.float32)          # :90
    sd = torch.sqrt(torch.exp(logvar))                                   # :91
    mean, sd = mean[elites], sd[elites]                                  # :93-94
    samples = mean[None] + torch.as_tensor(noise, dtype=torch.float32) * sd[None]        # :96
    return samples[..., :-1].numpy()                                     # :97


def train(ora: "DynamicsOracle", inputs: np.ndarray, targets: np.ndarray, num_elites: int, max_epochs: Optional[int] = None,
          max_epochs_since_update: int = 5, batch_size: int = 256, holdout_ratio: float = 0.2,
          logvar_loss_coef: float = 0.01) -> Dict:
    """dynamics/ensemble_dynamics.py:111-176: holdout split (torch ``random_split``), scaler fit on the training rows,
    per-member bootstrap index matrix (``np.random.randint``), one ``learn`` pass + ``validate`` per epoch, per-member
    ``update_save`` when the holdout loss improved by more than 1 %, row shuffle of the index matrix
    (``np.random.uniform`` + argsort), early stop, elite selection and ``load_save``.  Consumes the torch CPU generator
    and the NumPy global generator in the reference's order.  Returns the per-epoch log and the final state."""
    data_size = inputs.shape[0]
    holdout_size = min(int(data_size * holdout_ratio), 1000)
    train_size = data_size - holdout_size
    tr, ho = torch.utils.data.random_split(range(data_size), (train_size, holdout_size))      # :125
    tx, ty = inputs[tr.indices], targets[tr.indices]
    hx, hy = inputs[ho.indices], targets[ho.indices]
    mu, std = scaler_fit(tx)                                                                  # :129
    tx, hx = (tx - mu) / std, (hx - mu) / std
    E = ora.p["backbones.0.weight"].shape[0]
    holdout_losses = [1e10] * E
    idxes = np.random.randint(train_size, size=[E, train_size])                               # :134
    saved = {k: ora.p[k].detach().clone() for k in ora.p if "saved_" in k}
    log, epoch, cnt = [], 0, 0
    while True:
        epoch += 1
        train_loss = ora.learn(tx[idxes], ty[idxes], batch_size, logvar_loss_coef)            # :144
        new = ora.validate(hx, hy)
        log.append((train_loss, float(np.sort(new)[:num_elites].mean()), [float(v) for v in new]))
        order = np.argsort(np.random.uniform(size=idxes.shape), axis=-1)                      # :135-137,153
        idxes = idxes[np.arange(E)[:, None], order]
        improved = []
        for i, (n_, o_) in enumerate(zip(new, holdout_losses)):
            if (o_ - n_) / o_ > 0.01:
                improved.append(i)
                holdout_losses[i] = n_
        if improved:                                                                          # ensemble_linear.py:46-49
            for k in saved:
                saved[k][improved] = ora.p[k.replace("saved_", "")].detach()[improved]
            cnt = 0
        else:
            cnt += 1
        if cnt >= max_epochs_since_update or (max_epochs and epoch >= max_epochs):
            break
    elites = [i for _, i in sorted(zip(holdout_losses, range(E)), key=lambda x: x[0])][:num_elites]      # :219-223
    with torch.no_grad():                                                                     # load_save, ensemble_linear.py:43-45
        for k, v in saved.items():
            ora.p[k].copy_(v)
            ora.p[k.replace("saved_", "")].copy_(v)
    return {"log": log, "elites": elites, "holdout_losses": [float(v) for v in holdout_losses], "mu": mu, "std": std,
            "epochs": epoch}


def scaler_fit(data: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """utils/scaler.py:11-23."""
    mu = np.mean(data, axis=0, keepdims=True)
    std = np.std(data, axis=0, keepdims=True)
    std[std < 1e-12] = 1.0
    return mu, std


def term_halfcheetah(obs, act, next_obs):
    """utils/termination_fns.py:10-16."""
    ok = np.logical_and(np.all(next_obs > -100, axis=-1), np.all(next_obs < 100, axis=-1))
    return (~ok)[:, None]


def term_hopper(obs, act, next_obs):
    """utils/termination_fns.py:18-30 -- including ``np.abs(bool_array)``: only the upper bound is enforced."""
    height, angle = next_obs[:, 0], next_obs[:, 1]
    ok = np.isfinite(next_obs).all(axis=-1) * np.abs(next_obs[:, 1:] < 100).all(axis=-1) \
        * (height > .7) * (np.abs(angle) < .2)
    return (~ok)[:, None]


def term_walker2d(obs, act, next_obs):
    """utils/termination_fns.py:63-75."""
    height, angle = next_obs[:, 0], next_obs[:, 1]
    ok = np.logical_and(np.all(next_obs > -100, axis=-1), np.all(next_obs < 100, axis=-1)) \
        * (height > 0.8) * (height < 2.0) * (angle > -1.0) * (angle < 1.0)
    return (~ok)[:, None]


def rollout(select_action: Callable, step: Callable, init_obss: np.ndarray, length: int):
    """policy/model_based/mopo.py:45-79: h-step imagination with stable survivor compaction."""
    out = {k: [] for k in ("obss", "next_obss", "actions", "rewards", "terminals")}
    obs, n, rews = init_obss, 0, np.array([])
    for _ in range(length):
        act = select_action(obs)
        nobs, rew, term, _ = step(obs, act)
        for k, v in zip(out, (obs, nobs, act, rew, term)):
            out[k].append(v)
        n += len(obs)
        rews = np.append(rews, rew.flatten())
        keep = (~term).flatten()
        if keep.sum() == 0:
            break
        obs = nobs[keep]
    return {k: np.concatenate(v, axis=0) for k, v in out.items()}, \
        {"num_transitions": n, "reward_mean": rews.mean()}
