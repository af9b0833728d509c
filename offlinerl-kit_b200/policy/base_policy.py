"""Abstract policy contract the unchanged trainers rely on (reference: policy/base_policy.py:8-26)."""
from typing import Dict

import numpy as np
import torch.nn as nn


class BasePolicy(nn.Module):
    def train(self) -> None:                                     # noqa: D401  (signature as used by the trainers)
        raise NotImplementedError

    def eval(self) -> None:
        raise NotImplementedError

    def select_action(self, obs: np.ndarray, deterministic: bool = False) -> np.ndarray:
        raise NotImplementedError

    def learn(self, batch: Dict) -> Dict[str, float]:
        raise NotImplementedError


def engine_for(policy, key, make, **kw):
    """The policy's CUDA engine for the batch ``key`` (a batch size, or COMBO's (real, fake) split).  The first call
    builds the engine (``make()``); a different key gets a sibling that shares parameters, Adam state and scalars and
    owns its own staging / step graphs (engine/learner.py:Learner.for_batch) -- the reference's ``learn`` accepts any
    batch size per call (policy/base_policy.py:25-26).  ``policy._engine`` is always the one used last."""
    engines = policy.__dict__.setdefault("_engines", {})
    cur = policy._engine
    if cur is None:
        cur = policy._engine = engines[key] = make()
        return cur
    if engines.get(key) is cur:
        return cur
    if not engines:                 # an engine installed by hand (tests): treat it as the one for this key
        engines[key] = cur
        return cur
    eng = engines.get(key)
    if eng is None:
        B = key if isinstance(key, int) else sum(key)
        eng = engines[key] = cur.for_batch(B, **kw)
    eng.invalidate()                # the other sibling's updates may have by-passed this one's derived weight copies
    policy._engine = eng
    return eng


def learn_many(policy, buffer, n_steps: int, batch_size: int):
    """K-step form of ``for _ in range(K): policy.learn(buffer.sample(batch_size))`` with one host synchronisation
    (engine/learner.py:Learner.learn_many); same index stream, same noise, same parameters, the K loss dicts in order."""
    outs = policy.engine(int(batch_size)).learn_many(buffer, int(n_steps))
    after = getattr(policy, "_after_step", None)
    if outs and after is not None:
        after(outs[-1])
    return outs
