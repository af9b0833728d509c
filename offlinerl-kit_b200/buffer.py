"""ReplayBuffer with a device-resident mirror and a CUDA gather (reference: offlinerlkit/buffer/buffer.py).

Same constructor, public NumPy attributes and methods as the reference (``add``, ``add_batch``,
``load_dataset``, ``normalize_obs``, ``sample``, ``sample_all``).  The host arrays stay the source of truth; they are
mirrored into one row-major fp32 table in HBM, and ``sample`` draws the indices with the SAME
``np.random.randint`` call as the reference (buffer.py:98, so the index stream is bit-identical), uploads the
2 KB of int64 indices and launches ``orlk_replay_gather``.  There is no host-side gather and no CPU fallback.
"""
import ctypes as C
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib as L

FIELDS = ("observations", "actions", "next_observations", "terminals", "rewards")


class Batch(dict):
    """The dict returned by ``sample``.  The tensors are views of persistent staging memory owned by the buffer
    (overwritten by the next ``sample`` of the same size); ``stable`` tells the policy engines that they may bind
    their CUDA graphs to these addresses.

    The gather is LAZY: ``sample`` only draws the indices into pinned host memory.  A policy engine whose step graph is
    bound to this staging memory runs the index upload and the gather as the first two nodes of that graph (one launch
    for the whole step); any other access to the tensors materialises them first, so the batch always reads as the
    reference's dict of tensors."""
    stable = True

    def _materialise(self) -> None:
        st = self.__dict__.get("token")
        if st is not None and st.pending:
            st.materialise()

    def __getitem__(self, k):
        self._materialise()
        return dict.__getitem__(self, k)

    def get(self, k, default=None):
        self._materialise()
        return dict.get(self, k, default)

    def items(self):
        self._materialise()
        return dict.items(self)

    def values(self):
        self._materialise()
        return dict.values(self)

    # ``{**batch}`` / ``dict(batch)`` go through ``keys()`` + ``__getitem__`` for dict SUBCLASSES that override ``keys``
    # (CPython only takes the raw-table fast path otherwise), ``copy`` / ``__iter__`` / ``__or__`` are overridden too:
    # every way of reading the rows materialises them first.
    def keys(self):
        self._materialise()
        return dict.keys(self)

    def __iter__(self):
        self._materialise()
        return dict.__iter__(self)

    def copy(self):
        self._materialise()
        return dict(dict.items(self))

    def __or__(self, other):
        self._materialise()
        return dict(dict.items(self)) | other

    def __ror__(self, other):
        self._materialise()
        return other | dict(dict.items(self))

    @property
    def obs2(self) -> torch.Tensor:
        """[2B, O]: observations then next_observations (one GEMM operand for the actor)."""
        self._materialise()
        return self.__dict__["_obs2"]

    @property
    def indices(self) -> torch.Tensor:
        """The sampled row indices on the device (int64 [B])."""
        self._materialise()
        return self.__dict__["_indices"]

    @property
    def batch_size(self) -> int:
        return int(self.__dict__["_obs2"].shape[0]) // 2


class _Stage:
    N_SLOTS = 4

    def __init__(self, rt, B: int, O: int, A: int):
        dev = rt.device
        self.idx_host = [torch.empty(B, dtype=torch.int64).pin_memory() for _ in range(self.N_SLOTS)]
        self.idx_np = [t.numpy() for t in self.idx_host]
        self.events: List[C.c_void_p] = []
        for _ in range(self.N_SLOTS):
            ev = C.c_void_p()
            L.call("orlk_event_create", C.byref(ev))
            self.events.append(ev)
        self.armed = [0] * self.N_SLOTS            # the slot's event has been recorded at least once
        self.slot = 0
        self.idx_dev = torch.zeros(B, dtype=torch.int64, device=dev)
        self.obs2 = torch.zeros(2 * B, O, dtype=torch.float32, device=dev)
        self.act = torch.zeros(B, A, dtype=torch.float32, device=dev)
        self.rew = torch.zeros(B, 1, dtype=torch.float32, device=dev)
        self.term = torch.zeros(B, 1, dtype=torch.float32, device=dev)
        self.batch = Batch(observations=self.obs2[:B], actions=self.act, next_observations=self.obs2[B:],
                           terminals=self.term, rewards=self.rew)
        self.batch.__dict__["_obs2"] = self.obs2
        self.batch.__dict__["_indices"] = self.idx_dev
        self.batch.token = self                    # identity of the staging memory (engines bind their graphs to it)
        # lazy gather: the indices of the latest sample wait in one pinned buffer until somebody needs the rows
        self.rt = rt
        self.pending = False
        self.idx_pin = torch.empty(B, dtype=torch.int64).pin_memory()
        self.idx_pin_np = self.idx_pin.numpy()
        self.idx_pin_ptr = self.idx_pin.data_ptr()
        self.pin_event = C.c_void_p()
        L.call("orlk_event_create", C.byref(self.pin_event))
        self.pin_armed = False
        self.gather_args = None                    # (table, rows, row_w, O, A) of the owning buffer, set by gather()
        # constant tail of the orlk_replay_sample argument list
        self.out_args = (self.obs2.data_ptr(), self.act.data_ptr(), self.rew.data_ptr(), self.term.data_ptr())
        self.pinned_ptrs = [t.data_ptr() for t in self.idx_host]
        self.idx_dev_ptr = self.idx_dev.data_ptr()


    # ---- lazy gather
    def upload_op(self):
        """Index upload out of the pinned buffer (a graph node of the engines' fused step, or eager)."""
        L.call("orlk_memcpy_h2d_async", self.idx_dev_ptr, self.idx_pin_ptr, 8 * self.idx_dev.shape[0], self.rt.cur)

    def gather_op(self):
        tp, rows, row_w, O, A = self.gather_args
        L.call("orlk_replay_gather", tp, rows, row_w, O, A, self.idx_dev_ptr, self.idx_dev.shape[0], *self.out_args, self.rt.cur)

    def materialise(self) -> None:
        self.pending = False
        self.upload_op()
        L.call("orlk_event_record", self.pin_event, self.rt.cur)
        self.pin_armed = True
        self.gather_op()


class ReplayBuffer:
    def __init__(self, buffer_size: int, obs_shape: Tuple, obs_dtype: np.dtype, action_dim: int,
                 action_dtype: np.dtype, device: str = "cpu") -> None:
        self._max_size = int(buffer_size)
        self.obs_shape, self.obs_dtype = tuple(obs_shape), obs_dtype
        self.action_dim, self.action_dtype = int(action_dim), action_dtype
        self._ptr = 0
        self._size = 0
        # rows that so far exist only in the device table (add_batch of CUDA tensors): (segments, device tensors)
        self._host_pending: List[Tuple[list, list]] = []
        self._host_pending_rows = 0
        self.observations = np.zeros((self._max_size,) + self.obs_shape, dtype=obs_dtype)
        self.next_observations = np.zeros((self._max_size,) + self.obs_shape, dtype=obs_dtype)
        self.actions = np.zeros((self._max_size, self.action_dim), dtype=action_dtype)
        self.rewards = np.zeros((self._max_size, 1), dtype=np.float32)
        self.terminals = np.zeros((self._max_size, 1), dtype=np.float32)
        self.device = torch.device(device)
        # device mirror state
        self._rt = None
        self._table: Optional[torch.Tensor] = None
        self._dirty: List[Tuple[int, int]] = []       # host row ranges [lo, hi) not yet mirrored
        self._stages: Dict[int, _Stage] = {}

    # ------------------------------------------------------------------ host arrays (public attributes of the reference)
    # buffer.py:26-30 exposes the five NumPy arrays; here they are properties so that rows which arrived as CUDA tensors
    # are copied to the host only when somebody looks (sample_all, normalize_obs, a user reading .observations, ...).
    def _flush_host(self) -> None:
        if not self._host_pending:
            return
        pend, self._host_pending, self._host_pending_rows = self._host_pending, [], 0
        dst = (self._h_obs, self._h_nobs, self._h_act, self._h_rew, self._h_term)
        for segs, tensors in pend:
            for d, t in zip(dst, tensors):
                h = t.cpu().numpy().reshape((t.shape[0],) + d.shape[1:])
                for lo, s0, cnt in segs:
                    d[lo:lo + cnt] = h[s0:s0 + cnt]

    def _host_prop(name):
        def get(self):
            self._flush_host()
            return getattr(self, name)

        def put(self, value):
            self._flush_host()
            setattr(self, name, value)
        return property(get, put)

    observations = _host_prop("_h_obs")
    next_observations = _host_prop("_h_nobs")
    actions = _host_prop("_h_act")
    rewards = _host_prop("_h_rew")
    terminals = _host_prop("_h_term")
    del _host_prop

    # ------------------------------------------------------------------ host-side API (as the reference)
    def add(self, obs, next_obs, action, reward, terminal) -> None:
        i = self._ptr
        self.observations[i] = np.array(obs).copy()
        self.next_observations[i] = np.array(next_obs).copy()
        self.actions[i] = np.array(action).copy()
        self.rewards[i] = np.array(reward).copy()
        self.terminals[i] = np.array(terminal).copy()
        self._mark(i, i + 1)
        self._ptr = (self._ptr + 1) % self._max_size
        self._size = min(self._size + 1, self._max_size)

    accepts_device_batches = True      # add_batch takes CUDA tensors (MBPolicyTrainer hands rollouts over on the device)

    def add_batch(self, obss, next_obss, actions, rewards, terminals) -> None:
        """buffer.py:58-72.  The rows go to [ptr, ptr + n) modulo the capacity; a batch that fits is written as one or
        two contiguous slices (the reference's fancy-index assignment of 250 000 rollout rows costs 50 ms, the slices 7).
        CUDA tensors (``policy.rollout(..., device_out=True)``) are additionally packed straight into the device row
        table, so the rows never travel host -> device again."""
        fields = (obss, next_obss, actions, rewards, terminals)
        dev_in = all(torch.is_tensor(f) and f.is_cuda for f in fields)
        n = len(fields[0])
        cap = self._max_size
        direct = (dev_in and 0 < n < cap and self._table is not None and self._table.shape[0] == len(self._h_obs) == cap
                  and self.obs_dtype == np.float32 and self.action_dtype == np.float32)
        first = min(n, cap - self._ptr)
        segs = [(self._ptr, 0, first)] + ([(0, first, n - first)] if n > first else [])
        if direct:
            O, A = self._obs_dim, self.action_dim
            t = [f.detach().reshape(n, -1).to(torch.float32).contiguous() for f in fields]
            t = [x.clone() if x.data_ptr() == f.data_ptr() else x for x, f in zip(t, fields)]    # ours until flushed
            rt = self._runtime()
            for lo, s0, cnt in segs:
                L.call("orlk_replay_pack", t[0].data_ptr() + 4 * s0 * O, t[1].data_ptr() + 4 * s0 * O,
                       t[2].data_ptr() + 4 * s0 * A, t[3].data_ptr() + 4 * s0, t[4].data_ptr() + 4 * s0, cnt, O, A,
                       self._table.data_ptr(), self.row_width, lo, rt.cur)
            rt.sync()
            # the host arrays catch up lazily (``t`` is kept until then).  The backlog is a run of consecutive ring
            # segments: once the newer ones cover a whole ring, the oldest has been overwritten everywhere and is dropped
            self._host_pending.append((segs, t))
            self._host_pending_rows += n
            while self._host_pending_rows - self._host_pending[0][1][0].shape[0] >= cap:
                self._host_pending_rows -= self._host_pending.pop(0)[1][0].shape[0]
        else:
            host = [f.detach().cpu().numpy() if torch.is_tensor(f) else np.array(f) for f in fields]
            dst = (self.observations, self.next_observations, self.actions, self.rewards, self.terminals)
            if n >= cap:               # later rows overwrite earlier ones: keep the reference's index arithmetic
                at = np.arange(self._ptr, self._ptr + n) % cap
                for d, h in zip(dst, host):
                    d[at] = h.reshape((n,) + d.shape[1:])
                self._mark(0, cap)
            else:
                for d, h in zip(dst, host):
                    h = h.reshape((n,) + d.shape[1:])
                    for lo, s0, cnt in segs:
                        d[lo:lo + cnt] = h[s0:s0 + cnt]
                for lo, _, cnt in segs:
                    self._mark(lo, lo + cnt)
        self._ptr = (self._ptr + n) % self._max_size
        self._size = min(self._size + n, self._max_size)

    def load_dataset(self, dataset: Dict[str, np.ndarray]) -> None:
        self.observations = np.array(dataset["observations"], dtype=self.obs_dtype)
        self.next_observations = np.array(dataset["next_observations"], dtype=self.obs_dtype)
        self.actions = np.array(dataset["actions"], dtype=self.action_dtype)
        self.rewards = np.array(dataset["rewards"], dtype=np.float32).reshape(-1, 1)
        self.terminals = np.array(dataset["terminals"], dtype=np.float32).reshape(-1, 1)
        self._ptr = self._size = len(self.observations)
        self._table = None                              # capacity may have changed: rebuild the mirror lazily
        self._dirty = [(0, self._size)]

    def normalize_obs(self, eps: float = 1e-3) -> Tuple[np.ndarray, np.ndarray]:
        mean = self.observations.mean(0, keepdims=True)
        std = self.observations.std(0, keepdims=True) + eps
        self.observations = (self.observations - mean) / std
        self.next_observations = (self.next_observations - mean) / std
        self._dirty = [(0, len(self.observations))]
        return mean, std

    def sample_all(self) -> Dict[str, np.ndarray]:
        n = self._size
        return {"observations": self.observations[:n].copy(), "actions": self.actions[:n].copy(),
                "next_observations": self.next_observations[:n].copy(), "terminals": self.terminals[:n].copy(),
                "rewards": self.rewards[:n].copy()}

    # ------------------------------------------------------------------ device mirror
    def _mark(self, lo: int, hi: int) -> None:
        if self._dirty and self._dirty[-1][1] == lo:
            self._dirty[-1] = (self._dirty[-1][0], hi)
        else:
            self._dirty.append((lo, hi))

    @property
    def _obs_dim(self) -> int:
        return int(np.prod(self.obs_shape))

    @property
    def row_width(self) -> int:
        """floats per table row: [obs | next_obs | act | rew | term], padded to a multiple of 4 (16-byte rows)."""
        return (2 * self._obs_dim + self.action_dim + 2 + 3) // 4 * 4

    def _runtime(self):
        if self._rt is None:
            from .engine.core import get_runtime
            self._rt = get_runtime(self.device)
        return self._rt

    def _sync_mirror(self) -> None:
        rt = self._runtime()
        cap = len(self._h_obs)
        if self._table is None or self._table.shape[0] != cap:
            self._table = torch.zeros(cap, self.row_width, dtype=torch.float32, device=rt.device)
            self._dirty = [(0, max(self._size, 0))]
        O, A = self._obs_dim, self.action_dim
        for lo, hi in self._dirty:
            if hi <= lo:
                continue
            up = lambda a, w: torch.from_numpy(np.ascontiguousarray(a[lo:hi], dtype=np.float32).reshape(hi - lo, w)).to(rt.device)
            o, no, ac = up(self.observations, O), up(self.next_observations, O), up(self.actions, A)
            rw, tm = up(self.rewards, 1), up(self.terminals, 1)
            L.call("orlk_replay_pack", o.data_ptr(), no.data_ptr(), ac.data_ptr(), rw.data_ptr(), tm.data_ptr(),
                   hi - lo, O, A, self._table.data_ptr(), self.row_width, lo, rt.cur)
            rt.sync()      # the temporaries above are released after this
        self._dirty = []
        # cached scalars of the per-sample host call
        self._table_ptr, self._table_rows, self._row_w, self._O = self._table.data_ptr(), cap, self.row_width, O
        self._sample_fn = L.load().orlk_replay_sample

    def _stage(self, batch_size: int) -> _Stage:
        st = self._stages.get(batch_size)
        if st is None:
            st = _Stage(self._runtime(), batch_size, self._obs_dim, self.action_dim)
            self._stages[batch_size] = st
        return st

    def draw_indices(self, batch_size: int) -> np.ndarray:
        """buffer.py:98 -- the legacy NumPy global generator, exactly as the reference."""
        return np.random.randint(0, self._size, size=batch_size)

    def gather(self, indices: np.ndarray) -> Batch:
        """Device gather of the given host indices into the staging batch of that size."""
        if self.obs_dtype != np.float32 or self.action_dtype != np.float32:
            raise L.OrlkError("the device mirror stores fp32 rows; obs/action dtype must be float32")
        rt = self._rt or self._runtime()
        if self._dirty or self._table is None:
            self._sync_mirror()
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        B = idx.shape[0]
        st = self._stages.get(B) or self._stage(B)
        if st.pin_armed:            # an eager upload out of the pinned buffer may still be queued
            L.call("orlk_event_sync", st.pin_event)
            st.pin_armed = False
        st.idx_pin_np[:] = idx
        st.gather_args = (self._table_ptr, self._table_rows, self._row_w, self._O, self.action_dim)
        st.pending = True           # the rows are gathered by the first consumer (Batch docstring)
        return st.batch

    def gather_device(self, idx_dev: torch.Tensor) -> Batch:
        """Gather with indices that already live on the device (int64 [B]); no host traffic at all."""
        rt = self._runtime()
        if self._dirty or self._table is None:
            self._sync_mirror()
        B = int(idx_dev.shape[0])
        st = self._stage(B)
        st.pending = False
        L.call("orlk_replay_gather", self._table.data_ptr(), len(self._h_obs), self.row_width, self._obs_dim,
               self.action_dim, idx_dev.data_ptr(), B, st.obs2.data_ptr(), st.act.data_ptr(), st.rew.data_ptr(),
               st.term.data_ptr(), rt.cur)
        return st.batch

    def sample(self, batch_size: int) -> Dict[str, torch.Tensor]:
        return self.gather(self.draw_indices(batch_size))
