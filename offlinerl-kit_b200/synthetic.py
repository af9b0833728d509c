"""Synthetic D4RL-shaped data and deterministic parameter recipes.

There is no network and no d4rl/mujoco in this environment, so benchmarks,
tests and golden vectors all use the recipes below.
Everything here is NumPy ``default_rng`` (PCG64) based and therefore identical
on every machine; nothing here touches the CUDA library.
"""
from typing import Dict, Mapping

import numpy as np

SHAPES = {"hopper": (11, 3), "halfcheetah": (17, 6), "walker2d": (17, 6)}


def make_dataset(n: int, obs_dim: int, act_dim: int, seed: int = 0) -> Dict[str, np.ndarray]:
    """The exact recipe of SURVEY.md section 8(d) / BASELINE.md section 3."""
    rng = np.random.default_rng(seed)
    observations = rng.standard_normal((n, obs_dim), dtype=np.float32)
    next_observations = rng.standard_normal((n, obs_dim), dtype=np.float32)
    actions = rng.uniform(-1, 1, (n, act_dim)).astype(np.float32)
    rewards = rng.standard_normal(n).astype(np.float32)
    terminals = (rng.random(n) < 0.01).astype(np.float32)
    return {"observations": observations, "next_observations": next_observations, "actions": actions,
            "rewards": rewards, "terminals": terminals}


def param_recipe(shapes: Mapping[str, tuple], seed: int) -> Dict[str, np.ndarray]:
    """Deterministic fp32 parameter values for a ``name -> shape`` mapping, in mapping order.

    Weights and biases are drawn U(-b, b) with b = 1/sqrt(fan_in) (the scale of
    ``nn.Linear``'s default init); fan_in is the last axis for 2-D ``[out,in]``
    tensors and the second axis for 3-D ensemble ``[E,in,out]`` tensors; 1-D
    tensors use the fan_in of the tensor drawn just before them.  Golden files
    store only ``seed``; both the reference run and the tests rebuild the
    values from here.
    """
    rng = np.random.default_rng(seed)
    out, last_fan = {}, 1
    for name, shape in shapes.items():
        shape = tuple(int(s) for s in shape)
        if len(shape) == 2:
            last_fan = shape[1]
        elif len(shape) == 3 and shape[1] > 1:
            last_fan = shape[1]
        b = 1.0 / np.sqrt(max(last_fan, 1))
        out[name] = rng.uniform(-b, b, size=shape).astype(np.float32)
    return out
