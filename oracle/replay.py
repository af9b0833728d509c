"""CPU oracle for ``ReplayBuffer`` (buffer/buffer.py).  TEST INFRASTRUCTURE ONLY.

Index draw and gather are integer / byte work: plain NumPy, bit-exact.
"""
from typing import Dict, Tuple

import numpy as np

FIELDS = ("observations", "actions", "next_observations", "terminals", "rewards")


def draw_indices(size: int, batch_size: int) -> np.ndarray:
    """buffer.py:98 -- consumes the NumPy *legacy global* generator."""
    return np.random.randint(0, size, size=batch_size)


def gather(data: Dict[str, np.ndarray], idx: np.ndarray) -> Dict[str, np.ndarray]:
    """buffer.py:100-106 without the torch wrapping: fancy-index each of the five arrays."""
    return {k: data[k][idx] for k in FIELDS}


def normalize_obs(obs: np.ndarray, next_obs: np.ndarray, eps: float = 1e-3
                  ) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """buffer.py:88-94 -> (obs', next_obs', mean, std)."""
    mean = obs.mean(0, keepdims=True)
    std = obs.std(0, keepdims=True) + eps
    return (obs - mean) / std, (next_obs - mean) / std, mean, std


def ring_write(arrays: Dict[str, np.ndarray], ptr: int, size: int, max_size: int,
               new: Dict[str, np.ndarray]) -> Tuple[int, int]:
    """buffer.py:52-70 (add_batch): modular ring write; returns the new (ptr, size)."""
    n = len(new["observations"])
    at = np.arange(ptr, ptr + n) % max_size
    for k in FIELDS:
        arrays[k][at] = np.array(new[k]).copy()
    return (ptr + n) % max_size, min(size + n, max_size)
