"""Termination predicates for model rollouts (reference: utils/termination_fns.py).

Each predicate exists twice: a NumPy version with the reference's signature ``fn(obs, act, next_obs) -> bool [B,1]``
(kept for API compatibility and as the host-side check) and a device version inside ``orlk_dyn_step`` selected by
``fn.device_kind`` (0 halfcheetah, 1 hopper, 2 walker2d, 3 never).  Predicates without a ``device_kind`` are
evaluated on the host by ``EnsembleDynamics.step`` exactly as the reference does.
"""
import numpy as np


def _check(obs, act, next_obs):
    assert len(obs.shape) == len(next_obs.shape) == len(act.shape) == 2


def _within(next_obs, lo=-100, hi=100):
    return np.logical_and(np.all(next_obs > lo, axis=-1), np.all(next_obs < hi, axis=-1))


def termination_fn_halfcheetah(obs, act, next_obs):
    _check(obs, act, next_obs)
    return (~_within(next_obs))[:, None]


def termination_fn_hopper(obs, act, next_obs):
    _check(obs, act, next_obs)
    height, angle = next_obs[:, 0], next_obs[:, 1]
    # NB: the reference takes np.abs() of a BOOLEAN array (termination_fns.py:24), so only the upper bound on the
    # non-height coordinates is enforced; reproduced on purpose.
    alive = np.isfinite(next_obs).all(axis=-1) * np.abs(next_obs[:, 1:] < 100).all(axis=-1) \
        * (height > .7) * (np.abs(angle) < .2)
    return (~alive)[:, None]


def termination_fn_walker2d(obs, act, next_obs):
    _check(obs, act, next_obs)
    height, angle = next_obs[:, 0], next_obs[:, 1]
    alive = _within(next_obs) * (height > 0.8) * (height < 2.0) * (angle > -1.0) * (angle < 1.0)
    return (~alive)[:, None]


def termination_fn_never(obs, act, next_obs):
    _check(obs, act, next_obs)
    return np.zeros((len(obs), 1), dtype=bool)


termination_fn_halfcheetah.device_kind = 0
termination_fn_hopper.device_kind = 1
termination_fn_walker2d.device_kind = 2
termination_fn_never.device_kind = 3
termination_fn_halfcheetahveljump = termination_fn_never
termination_fn_point2denv = termination_fn_never
termination_fn_point2dwallenv = termination_fn_never


def obs_unnormalization(termination_fn, obs_mean, obs_std):
    def thunk(obs, act, next_obs):
        return termination_fn(obs * obs_std + obs_mean, act, next_obs * obs_std + obs_mean)
    return thunk


_BY_TASK = (("halfcheetahvel", termination_fn_never), ("halfcheetah", termination_fn_halfcheetah),
            ("hopper", termination_fn_hopper), ("walker2d", termination_fn_walker2d),
            ("point2dwallenv", termination_fn_never), ("point2denv", termination_fn_never), ("maze", termination_fn_never))


def get_termination_fn(task: str):
    for key, fn in _BY_TASK:
        if key in task:
            return fn
    raise NotImplementedError(f"no termination function for task {task!r}")
