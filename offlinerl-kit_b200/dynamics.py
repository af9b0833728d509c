"""BaseDynamics / EnsembleDynamics facades (reference: dynamics/base_dynamics.py:8-23, dynamics/ensemble_dynamics.py).

Same constructor and methods as the reference (``step``, ``train``, ``learn``, ``validate``, ``select_elites``, ``save``,
``load``, ``format_samples_for_training``); all model arithmetic runs in engine/dynamics.py on the device.  ``train``
keeps the (scaled) training set, the targets and the per-member bootstrap index matrix resident in HBM and gathers
each mini-batch with a kernel, instead of the reference's > 1 GB host fancy-index per epoch.
"""
import os
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from .utils.scaler import StandardScaler


UNCERTAINTY_MODES = {"aleatoric": 0, "pairwise-diff": 1, "ensemble_std": 2}      # ensemble_dynamics.py:60-70


class BaseDynamics:
    def __init__(self, model: nn.Module, optim: torch.optim.Optimizer) -> None:
        self.model, self.optim = model, optim

    def step(self, obs: np.ndarray, action: np.ndarray):
        raise NotImplementedError


class EnsembleDynamics(BaseDynamics):
    def __init__(self, model: nn.Module, optim: torch.optim.Optimizer, scaler: StandardScaler,
                 terminal_fn: Callable[[np.ndarray, np.ndarray, np.ndarray], np.ndarray], penalty_coef: float = 0.0,
                 uncertainty_mode: str = "aleatoric") -> None:
        super().__init__(model, optim)
        self.scaler, self.terminal_fn = scaler, terminal_fn
        self._penalty_coef, self._uncertainty_mode = penalty_coef, uncertainty_mode
        if uncertainty_mode not in UNCERTAINTY_MODES:
            raise ValueError(f"unknown uncertainty_mode {uncertainty_mode!r}")       # ensemble_dynamics.py:71-72
        self._engine = None
        self._scaler_dev = None
        self._shard = None
        # Noise of the imagination step when the caller injects none: "device" = Philox on the GPU (default);
        # "numpy" = the reference's two host draws per step, np.random.normal(size=[E, S, D]) float64 and
        # np.random.choice(elites, S) (ensemble_dynamics.py:48, dynamics_module.py:118) -- the same stream as the
        # reference under the same np.random.seed (parity tests), but 220 ms of host time per 50 000-state step.
        self.rng = os.environ.get("ORLK_DYN_RNG", "device")

    def shard_members(self, rank: int, world: int, comm=None) -> None:
        """Train only members partition_members(E, world)[rank] on this rank (BASELINE.json configs[4]: "members sharded over
        8 x B200").  Data, seeds and the model's initial state must be replicated on every rank.  ``train`` then exchanges
        the shared log-variance bounds' gradients per mini-batch and the holdout losses per epoch, and ends with every rank
        holding the whole trained ensemble (rollouts need all members).  ``comm``: engine.edac_sharded.NcclComm()."""
        if self._engine is not None:
            raise RuntimeError("shard_members must be called before the engine is first used")
        self._shard = (int(rank), int(world), comm)

    # ------------------------------------------------------------------ engine plumbing
    @property
    def engine(self):
        if self._engine is None:
            from .engine.dynamics import DynamicsEngine
            self._engine = DynamicsEngine(self.model, self.optim, shard=self._shard)
        return self._engine

    def _dev(self, a, dtype=torch.float32) -> torch.Tensor:
        return torch.as_tensor(np.ascontiguousarray(a)).to(device=self.engine.dev, dtype=dtype)

    def _scaler_tensors(self):
        """Device copies of the scaler's mu / std.  The cache holds the host arrays it was made from and compares their
        CONTENT (36-element arrays): a re-fit, a load or an in-place edit of scaler.mu / std is always seen."""
        mu, std = self.scaler.mu, self.scaler.std
        c = self._scaler_dev
        if c is None or not (np.array_equal(c[0], mu) and np.array_equal(c[1], std)):
            self._scaler_dev = c = (np.array(mu, copy=True), np.array(std, copy=True),
                                    self._dev(np.asarray(mu, np.float32).reshape(-1)),
                                    self._dev(np.asarray(std, np.float32).reshape(-1)))
        return c[2], c[3]

    # ------------------------------------------------------------------ imagination
    def step_device(self, obs: torch.Tensor, act: torch.Tensor, noise64=None, midx=None, out=None, noise_buf=None):
        """One imagined step on device tensors -> (next_obs, reward, terminal(uint8), raw_reward, penalty) on the device.
        ``out`` / ``noise_buf`` [S*D + S]: pre-allocated results / Philox scratch (engine/rollout.py's sync-free loop)."""
        eng = self.engine
        S = obs.shape[0]
        mu, sd = self._scaler_tensors()
        kind = getattr(self.terminal_fn, "device_kind", None)
        n32 = pick = elites = None
        if noise64 is None:
            if self.rng == "numpy":           # ensemble_dynamics.py:48 and dynamics_module.py:118, in that order
                noise64 = self._dev(np.random.normal(size=(eng.E, S, eng.D)), torch.float64)
                midx = self._dev(self.model.random_elite_idxs(S), torch.int32)
            else:
                buf = noise_buf if noise_buf is not None else torch.empty(S * eng.D + S, dtype=torch.float32, device=eng.dev)
                L.call("orlk_philox_fill", buf.data_ptr(), S * eng.D, S, 0.0, 1.0, 0x5eed, eng.philox_counter.data_ptr(), None,
                       eng.rt.cur)
                L.call("orlk_step_end", eng.groups_ptr, 0, eng.philox_counter.data_ptr(), eng.rt.cur)
                n32, pick = buf[:S * eng.D], buf[S * eng.D:]
                elites = self._elites_dev()
        return eng.imagine(obs, act, mu, sd, 3 if kind is None else kind, self._penalty_coef, noise64, midx, n32, pick, elites,
                           UNCERTAINTY_MODES[self._uncertainty_mode], out=out)

    def _elites_dev(self) -> torch.Tensor:
        """int32 device copy of model.elites (re-made when set_elites replaced the parameter)."""
        el = self.model.elites
        c = getattr(self, "_elites_cache", None)
        if c is None or c[0] is not el or c[1] != el._version:
            self._elites_cache = c = (el, el._version, el.data.to(device=self.engine.dev, dtype=torch.int32))
        return c[2]

    @torch.no_grad()
    def step(self, obs: np.ndarray, action: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray, Dict]:
        obs_d, act_d = self._dev(obs), self._dev(action)
        nobs, rew, term, raw, pen = self.step_device(obs_d, act_d)
        self.engine.rt.sync()
        next_obs, reward = nobs.cpu().numpy(), rew.cpu().numpy()
        if getattr(self.terminal_fn, "device_kind", None) is None:
            terminal = self.terminal_fn(obs, action, next_obs)       # unknown predicate: evaluate it on the host
        else:
            terminal = term.cpu().numpy().astype(bool)
        info = {"raw_reward": raw.cpu().numpy()}
        if self._penalty_coef:
            info["penalty"] = pen.cpu().numpy()
        return next_obs, reward, terminal, info

    @torch.no_grad()
    def sample_next_obss(self, obs: torch.Tensor, action: torch.Tensor, num_samples: int, noise: Optional[torch.Tensor] = None
                         ) -> torch.Tensor:
        """ensemble_dynamics.py:81-99 (the uncertainty samples of MOBILE's penalty): ``num_samples`` draws of every ELITE
        member's Gaussian next-observation prediction, [num_samples, n_elites, B, obs_dim] on the device.  The ensemble
        forward pass runs on the engine (tensor cores from 2048 rows up); the elementwise tail follows the reference line
        by line.  ``noise`` [num_samples, n_elites, B, obs_dim + 1]: parity hook - the reference draws
        ``torch.randn_like(std)`` once per sample."""
        eng = self.engine
        obs_d = torch.as_tensor(obs, dtype=torch.float32, device=eng.dev).contiguous()
        act_d = torch.as_tensor(action, dtype=torch.float32, device=eng.dev).contiguous()
        S, O, A = obs_d.shape[0], obs_d.shape[1], act_d.shape[1]
        mu, sd = self._scaler_tensors()
        xbuf = eng.input_buffer(S)
        L.call("orlk_dyn_input", obs_d.data_ptr(), obs_d.stride(0), act_d.data_ptr(), act_d.stride(0), mu.data_ptr(), sd.data_ptr(),
               S, O, A, xbuf.data_ptr(), xbuf.stride(0), eng.rt.cur)
        out = eng._forward(xbuf).OUT                                    # [E, S, 2 D]: mean | raw log-variance
        D = out.shape[-1] // 2
        mean, logvar = out[..., :D].clone(), out[..., D:]
        max_lv, min_lv = self.model.max_logvar, self.model.min_logvar   # soft_clamp, dynamics_module.py:8-16
        logvar = max_lv - torch.nn.functional.softplus(max_lv - logvar)
        logvar = min_lv + torch.nn.functional.softplus(logvar - min_lv)
        mean[..., :-1] += obs_d
        std = torch.sqrt(torch.exp(logvar))
        el = self._elites_dev().long()
        mean, std = mean[el], std[el]
        if noise is None:
            noise = torch.randn((num_samples,) + tuple(std.shape), dtype=torch.float32, device=eng.dev)
        else:
            noise = torch.as_tensor(noise, dtype=torch.float32, device=eng.dev)
        samples = mean[None] + noise * std[None]
        return samples[..., :-1]

    # ------------------------------------------------------------------ training
    def format_samples_for_training(self, data: Dict) -> Tuple[np.ndarray, np.ndarray]:
        inputs = np.concatenate((data["observations"], data["actions"]), axis=-1)
        targets = np.concatenate((data["next_observations"] - data["observations"], data["rewards"]), axis=-1)
        return inputs, targets

    def train(self, data: Dict, logger, max_epochs: Optional[float] = None, max_epochs_since_update: int = 5,
              batch_size: int = 256, holdout_ratio: float = 0.2, logvar_loss_coef: float = 0.01) -> None:
        inputs, targets = self.format_samples_for_training(data)
        data_size = inputs.shape[0]
        holdout_size = min(int(data_size * holdout_ratio), 1000)
        train_size = data_size - holdout_size
        train_split, holdout_split = torch.utils.data.random_split(range(data_size), (train_size, holdout_size))
        train_inputs, train_targets = inputs[train_split.indices], targets[train_split.indices]
        holdout_inputs, holdout_targets = inputs[holdout_split.indices], targets[holdout_split.indices]
        self.scaler.fit(train_inputs)
        E = self.model.num_ensemble
        x_dev, y_dev = self._dev(self.scaler.transform(train_inputs)), self._dev(train_targets)
        hx_dev, hy_dev = self._dev(self.scaler.transform(holdout_inputs)), self._dev(holdout_targets)
        holdout_losses = [1e10 for _ in range(E)]
        data_idxes = np.random.randint(train_size, size=[E, train_size])

        def shuffle_rows(arr):
            order = np.argsort(np.random.uniform(size=arr.shape), axis=-1)
            return arr[np.arange(arr.shape[0])[:, None], order]

        epoch = cnt = 0
        logger.log("Training dynamics:")
        # ensemble_dynamics.py:141-144 reshuffles the bootstrap index matrix on the host after every epoch (an argsort of
        # E x N uniforms: ~0.8 s for 7 x 1M rows, as long as the device epoch itself).  The shuffle is the only consumer
        # of np.random inside this loop and does not depend on the epoch's results, so it runs on a worker thread while
        # the device trains; the random stream and every index matrix are exactly the reference's.
        overlap = os.environ.get("ORLK_DYN_SHUFFLE_OVERLAP", "1") != "0"
        pool = ThreadPoolExecutor(max_workers=1) if overlap else None
        while True:
            epoch += 1
            idx_dev = torch.as_tensor(data_idxes, dtype=torch.int64).to(self.engine.dev)
            pending = pool.submit(shuffle_rows, data_idxes) if overlap else None
            train_loss = self.engine.learn(x_dev, y_dev, idx_dev, batch_size, logvar_loss_coef)
            new_holdout_losses = self.engine.validate(hx_dev, hy_dev)
            holdout_loss = (np.sort(new_holdout_losses)[:self.model.num_elites]).mean()
            logger.logkv("loss/dynamics_train_loss", train_loss)
            logger.logkv("loss/dynamics_holdout_loss", holdout_loss)
            logger.set_timestep(epoch)
            logger.dumpkvs(exclude=["policy_training_progress"])
            data_idxes = pending.result() if overlap else shuffle_rows(data_idxes)
            improved = []
            for i, (new, old) in enumerate(zip(new_holdout_losses, holdout_losses)):
                if (old - new) / old > 0.01:
                    improved.append(i)
                    holdout_losses[i] = new
            if improved:
                if self._shard is not None:
                    self.engine.write_back()        # the own members' trained rows -> the full model, for update_save
                self.model.update_save(improved)
                cnt = 0
            else:
                cnt += 1
            if cnt >= max_epochs_since_update or (max_epochs and epoch >= max_epochs):
                break
        if pool is not None:
            pool.shutdown(wait=True)
        elites = self.select_elites(holdout_losses)
        self.model.set_elites(elites)
        self.model.load_save()
        if self._shard is not None:
            # every rank contributes the best snapshot of ITS members; afterwards all ranks hold the whole ensemble and the
            # next engine (validation, rollouts) is an ordinary unsharded one
            eng = self.engine
            eng.read_back()
            eng.gather_all()
            eng.rt.sync()
            self._engine, self._shard = None, None
        self.save(logger.model_dir)
        self.model.eval()
        logger.log("elites:{} , holdout loss: {}".format(elites, (np.sort(holdout_losses)[:self.model.num_elites]).mean()))

    def learn(self, inputs: np.ndarray, targets: np.ndarray, batch_size: int = 256, logvar_loss_coef: float = 0.01) -> float:
        """Reference signature: inputs [E, n, in], targets [E, n, out] already gathered per member."""
        self.model.train()
        E, n = inputs.shape[0], inputs.shape[1]
        x = self._dev(np.asarray(inputs, np.float32).reshape(E * n, -1))
        y = self._dev(np.asarray(targets, np.float32).reshape(E * n, -1))
        idx = (torch.arange(E, device=x.device).view(E, 1) * n + torch.arange(n, device=x.device).view(1, n)).contiguous()
        return self.engine.learn(x, y, idx, batch_size, logvar_loss_coef)

    @torch.no_grad()
    def validate(self, inputs: np.ndarray, targets: np.ndarray) -> List[float]:
        self.model.eval()
        return self.engine.validate(self._dev(inputs), self._dev(targets))

    def select_elites(self, metrics: List) -> List[int]:
        order = sorted(range(len(metrics)), key=lambda i: metrics[i])
        return order[:self.model.num_elites]

    def save(self, save_path: str) -> None:
        torch.save(self.model.state_dict(), os.path.join(save_path, "dynamics.pth"))
        self.scaler.save_scaler(save_path)

    def load(self, load_path: str) -> None:
        self.model.load_state_dict(torch.load(os.path.join(load_path, "dynamics.pth"), map_location=self.model.device))
        self.scaler.load_scaler(load_path)
