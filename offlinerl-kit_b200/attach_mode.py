"""Attach mode (mode i): put the CUDA engine behind objects built by the UNMODIFIED reference.

    import offlinerlkit                      # the reference, e.g. from baseline/_ref
    policy = offlinerlkit.policy.CQLPolicy(actor, critic1, critic2, ...)       # exactly as run_example/run_cql.py:80-128
    buffer = offlinerlkit.buffer.ReplayBuffer(...); buffer.load_dataset(dataset)
    import offlinerlkit_b200 as orlk
    orlk.attach(policy, buffer)              # from here on policy.learn / buffer.sample run on the engine

``attach`` does not copy or rebuild anything the caller can see: the reference's ``nn.Module`` parameters stay the
canonical storage (their ``.data`` is re-pointed at views of the engine's arenas, so ``state_dict`` / ``torch.save`` /
``select_action`` keep working), the reference's ``torch.optim.Adam`` objects are read for their hyper-parameters
(lr schedulers included), and the reference's trainers (``MFPolicyTrainer`` / ``MBPolicyTrainer``) call the same two
methods as before.  What changes is the class of the two objects: each becomes an instance of a dynamically created
subclass of ITS OWN reference class that overrides ``learn`` (``sample`` / ``add_batch`` / ... for the buffer), so
``isinstance`` checks against the reference's classes still hold.

The engine reads the reference objects through the attribute names the reference itself uses (``policy.actor``,
``policy._tau``, ``policy.critic1_old``, ``actor.dist_net._c_sigma`` ...: policy/model_free/{sac,cql,edac,iql,td3bc}.py);
the stand-alone facades of this package carry the same names, which is why one set of engines serves both modes."""
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib as L
from .policy.base_policy import engine_for

# reference class name (searched along the MRO, most derived first) -> (engine module, learner class)
_LEARNERS = {
    "CQLPolicy": ("sac_family", "CQLLearner"),
    "EDACPolicy": ("edac", "EDACLearner"),
    "IQLPolicy": ("td3_iql", "IQLLearner"),
    "TD3BCPolicy": ("td3_iql", "TD3BCLearner"),
    "SACPolicy": ("sac_family", "SACLearner"),
}
_UNSUPPORTED = ("COMBOPolicy", "MOPOPolicy", "RAMBOPolicy", "MOBILEPolicy", "MCQPolicy", "TD3Policy")


def _learner_for(policy):
    names = [c.__name__ for c in type(policy).__mro__]
    for n in names:
        if n in _UNSUPPORTED:
            raise L.OrlkError(f"attach: {n} is not attachable (model-based policies: use the stand-alone facades "
                              "offlinerlkit_b200.policy.MOPOPolicy / COMBOPolicy, which take the same constructor arguments)")
        if n in _LEARNERS:
            import importlib
            mod, cls = _LEARNERS[n]
            return getattr(importlib.import_module(f".engine.{mod}", __package__), cls)
    raise L.OrlkError(f"attach: no CUDA engine for {type(policy).__name__} (supported: {sorted(_LEARNERS)})")


def attach_policy(policy):
    """Swap ``policy.learn`` for the engine-backed step; returns the same object (now of an ``Attached<Class>`` subclass)."""
    if getattr(policy, "_orlk_attached", False):
        return policy
    learner = _learner_for(policy)
    base = type(policy)

    def engine(self, batch_size: int):
        return engine_for(self, int(batch_size), lambda: learner(self, int(batch_size)))

    def learn(self, batch: Dict, noise: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, float]:
        """The reference's contract, policy/base_policy.py:25-26: Dict[str, Tensor] -> Dict[str, float]."""
        B = getattr(batch, "batch_size", None) or int(batch["observations"].shape[0])
        out = self.engine(B).step(batch, noise)
        if "alpha" in out and getattr(self, "_is_auto_alpha", False):
            self.__dict__["_alpha_value"], self.__dict__["_alpha_tensor"] = out["alpha"], None
        return out

    ns = {"learn": learn, "engine": engine, "_orlk_attached": True}
    if getattr(policy, "_is_auto_alpha", False):
        # ``_alpha`` is a tensor attribute of the reference (sac.py:43-49) that its ``learn`` re-assigns every step; the
        # engine keeps the value on the device and the tensor is materialised when somebody reads the attribute
        def get_alpha(self):
            t = self.__dict__.get("_alpha_tensor")
            if t is None:
                t = torch.tensor([self.__dict__["_alpha_value"]], device=self.actor.device)
                self.__dict__["_alpha_tensor"] = t
            return t

        def set_alpha(self, value):
            self.__dict__["_alpha_tensor"] = value if torch.is_tensor(value) else None
            self.__dict__["_alpha_value"] = float(value.detach().reshape(-1)[0]) if torch.is_tensor(value) else float(value)

        cur = policy.__dict__.pop("_alpha")
        policy.__dict__["_alpha_tensor"] = cur
        policy.__dict__["_alpha_value"] = float(cur.detach().reshape(-1)[0])
        ns["_alpha"] = property(get_alpha, set_alpha)
    policy.__dict__["_engine"] = None
    policy.__class__ = type("Attached" + base.__name__, (base,), ns)
    return policy


def attach_buffer(buffer):
    """Give a reference ``ReplayBuffer`` (buffer/buffer.py:8-115) the device row table and the lazy in-graph gather.

    The object's NumPy arrays stay where they are and stay the source of truth (``sample_all``, ``normalize_obs`` and
    user code that reads ``buffer.observations`` see the same arrays); ``sample`` draws the same ``np.random.randint``
    index stream and returns the batch from the device mirror."""
    if getattr(buffer, "_orlk_attached", False):
        return buffer
    from .buffer import ReplayBuffer as Mirror
    base = type(buffer)
    m = Mirror.__new__(Mirror)
    # adopt the reference object's state: same arrays, same ring position
    m._max_size = int(buffer._max_size)
    m.obs_shape, m.obs_dtype = tuple(buffer.obs_shape), buffer.obs_dtype
    m.action_dim, m.action_dtype = int(buffer.action_dim), buffer.action_dtype
    m._ptr, m._size = int(buffer._ptr), int(buffer._size)
    m._host_pending, m._host_pending_rows = [], 0
    m._h_obs, m._h_nobs, m._h_act = buffer.observations, buffer.next_observations, buffer.actions
    m._h_rew, m._h_term = buffer.rewards, buffer.terminals
    m.device = torch.device(buffer.device)
    m._rt, m._table, m._stages = None, None, {}
    m._dirty = [(0, max(m._size, 0))]
    buffer.__dict__["_orlk_mirror"] = m

    def _sync_from(self):
        """host-side writes through the reference's own methods re-bind the arrays / move the ring: follow them"""
        mm = self._orlk_mirror
        changed = False
        for name, attr in (("observations", "_h_obs"), ("next_observations", "_h_nobs"), ("actions", "_h_act"),
                           ("rewards", "_h_rew"), ("terminals", "_h_term")):
            arr = self.__dict__.get(name)
            if arr is not None and arr is not getattr(mm, attr):
                setattr(mm, attr, arr)                      # load_dataset / normalize_obs re-bind the arrays
                changed = True
        if changed:
            mm._max_size = len(mm._h_obs)
            mm._dirty = [(0, len(mm._h_obs))]               # (a table of another capacity is re-allocated by _sync_mirror)
        mm._ptr, mm._size = int(self._ptr), int(self._size)

    def sample(self, batch_size: int):
        _sync_from(self)
        return self._orlk_mirror.sample(batch_size)

    def add(self, *a, **k):
        i = self._ptr
        base.add(self, *a, **k)
        self._orlk_mirror._mark(i, i + 1)

    def add_batch(self, obss, next_obss, actions, rewards, terminals):
        n, p, cap = len(obss), self._ptr, self._max_size
        base.add_batch(self, obss, next_obss, actions, rewards, terminals)
        mm = self._orlk_mirror
        if n >= cap or p + n > cap:
            mm._mark(0, cap)
        else:
            mm._mark(p, p + n)

    def load_dataset(self, dataset):
        base.load_dataset(self, dataset)
        _sync_from(self)

    def normalize_obs(self, *a, **k):
        out = base.normalize_obs(self, *a, **k)
        _sync_from(self)
        self._orlk_mirror._dirty = [(0, len(self.observations))]
        return out

    ns = {"sample": sample, "add": add, "add_batch": add_batch, "load_dataset": load_dataset, "normalize_obs": normalize_obs,
          "_orlk_attached": True}
    buffer.__class__ = type("Attached" + base.__name__, (base,), ns)
    return buffer


def attach(policy, buffer=None):
    """``attach(policy, buffer)``: see the module docstring.  Raises ``OrlkError`` for configurations the engine does not
    implement (at the first ``learn``, when the step plan is built) -- there is no silent fallback to the reference path."""
    from .engine.core import require_cuda
    require_cuda(policy.actor.device)
    attach_policy(policy)
    if buffer is not None:
        attach_buffer(buffer)
    return policy, buffer
