"""MFPolicyTrainer / MBPolicyTrainer: the callers of the hot path (reference: policy_trainer/mf_policy_trainer.py:41-118,
policy_trainer/mb_policy_trainer.py:57-207).  Same constructor arguments and the same loop structure -- sample, learn,
log, (periodic model rollouts into the fake buffer), evaluate, checkpoint -- so that the reference's run scripts drive
this package unchanged.  The per-step tqdm redraw of the reference is replaced by one progress line per epoch: it cost
~55 us per step, more than a quarter of a fused gradient step."""
import os
import time
from collections import deque
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch


class _TrainerBase:
    def _evaluate(self) -> Dict[str, List[float]]:
        self.policy.eval()
        obs = self.eval_env.reset()
        episodes, ep_reward, ep_len = [], 0.0, 0
        while len(episodes) < self._eval_episodes:
            action = self.policy.select_action(np.asarray(obs).reshape(1, -1), deterministic=True)
            obs, reward, terminal, _ = self.eval_env.step(action.flatten())
            ep_reward += reward
            ep_len += 1
            if terminal:
                episodes.append((ep_reward, ep_len))
                ep_reward, ep_len = 0.0, 0
                obs = self.eval_env.reset()
        return {"eval/episode_reward": [r for r, _ in episodes], "eval/episode_length": [n for _, n in episodes]}

    def _log_eval(self, last_10: deque) -> None:
        info = self._evaluate()
        rew, length = info["eval/episode_reward"], info["eval/episode_length"]
        if hasattr(self.eval_env, "get_normalized_score"):
            norm_mean = self.eval_env.get_normalized_score(np.mean(rew)) * 100
            last_10.append(norm_mean)
            self.logger.logkv("eval/normalized_episode_reward", norm_mean)
            self.logger.logkv("eval/normalized_episode_reward_std", self.eval_env.get_normalized_score(np.std(rew)) * 100)
        self.logger.logkv("eval/episode_reward", np.mean(rew))
        self.logger.logkv("eval/episode_reward_std", np.std(rew))
        self.logger.logkv("eval/episode_length", np.mean(length))
        self.logger.logkv("eval/episode_length_std", np.std(length))


class MFPolicyTrainer(_TrainerBase):
    def __init__(self, policy, eval_env, buffer, logger, epoch: int = 1000, step_per_epoch: int = 1000, batch_size: int = 256,
                 eval_episodes: int = 10, lr_scheduler: Optional[torch.optim.lr_scheduler._LRScheduler] = None) -> None:
        self.policy, self.eval_env, self.buffer, self.logger = policy, eval_env, buffer, logger
        self._epoch, self._step_per_epoch, self._batch_size = epoch, step_per_epoch, batch_size
        self._eval_episodes, self.lr_scheduler = eval_episodes, lr_scheduler

    def train(self) -> Dict[str, float]:
        start = time.time()
        num_timesteps, last_10 = 0, deque(maxlen=10)
        for e in range(1, self._epoch + 1):
            self.policy.train()
            t0 = time.time()
            many = getattr(self.policy, "learn_many", None) if getattr(self, "use_learn_many", True) else None
            left = self._step_per_epoch
            while left > 0:
                if many is not None and hasattr(self.buffer, "gather_device"):
                    # K steps behind one host synchronisation: same np.random index stream, same losses, same logged means
                    # (mf_policy_trainer.py:52-60 spends ~55 us per step on the host between two ~300 us device steps)
                    losses = many(self.buffer, min(left, 250), self._batch_size)
                else:
                    losses = [self.policy.learn(self.buffer.sample(self._batch_size))]
                for loss in losses:
                    for k, v in loss.items():
                        self.logger.logkv_mean(k, v)
                num_timesteps += len(losses)
                left -= len(losses)
            self.logger.logkv("train/steps_per_second", self._step_per_epoch / max(time.time() - t0, 1e-9))
            if self.lr_scheduler is not None:
                self.lr_scheduler.step()
            self._log_eval(last_10)
            self.logger.set_timestep(num_timesteps)
            self.logger.dumpkvs()
            torch.save(self.policy.state_dict(), os.path.join(self.logger.checkpoint_dir, "policy.pth"))
        self.logger.log("total time: {:.2f}s".format(time.time() - start))
        torch.save(self.policy.state_dict(), os.path.join(self.logger.model_dir, "policy.pth"))
        self.logger.close()
        return {"last_10_performance": np.mean(last_10) if last_10 else float("nan")}


class MBPolicyTrainer(_TrainerBase):
    def __init__(self, policy, eval_env, real_buffer, fake_buffer, logger, rollout_setting: Tuple[int, int, int],
                 epoch: int = 1000, step_per_epoch: int = 1000, batch_size: int = 256, real_ratio: float = 0.05,
                 eval_episodes: int = 10, lr_scheduler=None, dynamics_update_freq: int = 0, horizon: Optional[int] = None) -> None:
        self.policy, self.eval_env, self.real_buffer, self.fake_buffer, self.logger = policy, eval_env, real_buffer, fake_buffer, logger
        self._rollout_freq, self._rollout_batch_size, self._rollout_length = rollout_setting
        self._dynamics_update_freq = dynamics_update_freq
        self._epoch, self._step_per_epoch, self._batch_size, self._real_ratio = epoch, step_per_epoch, batch_size, real_ratio
        self._eval_episodes, self.lr_scheduler, self.horizon = eval_episodes, lr_scheduler, horizon

    def train(self) -> Dict[str, float]:
        start = time.time()
        num_timesteps, last_10 = 0, deque(maxlen=10)
        for e in range(1, self._epoch + 1):
            self.policy.train()
            for _ in range(self._step_per_epoch):
                if num_timesteps % self._rollout_freq == 0:
                    init = self.real_buffer.sample(self._rollout_batch_size)["observations"].cpu().numpy()
                    if getattr(self.policy, "device_rollouts", False) and getattr(self.fake_buffer, "accepts_device_batches", False):
                        # same hand-off as mb_policy_trainer.py:71-73, with the transitions staying on the device
                        transitions, info = self.policy.rollout(init, self._rollout_length, device_out=True)
                    else:
                        transitions, info = self.policy.rollout(init, self._rollout_length)
                    self.fake_buffer.add_batch(**transitions)
                    self.logger.log("num rollout transitions: {}, reward mean: {:.4f}".format(info["num_transitions"],
                                                                                             info["reward_mean"]))
                    for k, v in info.items():
                        self.logger.logkv_mean("rollout_info/" + k, v)
                real_n = int(self._batch_size * self._real_ratio)
                batch = {"real": self.real_buffer.sample(real_n), "fake": self.fake_buffer.sample(self._batch_size - real_n)}
                loss = self.policy.learn(batch)
                for k, v in loss.items():
                    self.logger.logkv_mean(k, v)
                if 0 < self._dynamics_update_freq and (num_timesteps + 1) % self._dynamics_update_freq == 0:
                    for k, v in self.policy.update_dynamics(self.real_buffer).items():
                        self.logger.logkv_mean(k, v)
                num_timesteps += 1
            if self.lr_scheduler is not None:
                self.lr_scheduler.step()
            self._log_eval(last_10)
            self.logger.set_timestep(num_timesteps)
            self.logger.dumpkvs(exclude=["dynamics_training_progress"])
            torch.save(self.policy.state_dict(), os.path.join(self.logger.checkpoint_dir, "policy.pth"))
        self.logger.log("total time: {:.2f}s".format(time.time() - start))
        torch.save(self.policy.state_dict(), os.path.join(self.logger.model_dir, "policy.pth"))
        self.policy.dynamics.save(self.logger.model_dir)
        self.logger.close()
        return {"last_10_performance": np.mean(last_10) if last_10 else float("nan")}
