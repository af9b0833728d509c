"""Shared machinery of the per-algorithm gradient-step engines.

A ``Learner`` owns: the Adam group table (lr / betas / eps / tau / step counters, device resident), the loss block
(device + pinned host), the noise block (filled by the Philox kernel, or by the caller in parity mode) and the
staging batch.  Sub-classes build a ``Plan`` (ordered launches) per step variant; ``step()`` replays it as one
CUDA graph and returns the loss scalars after a single stream synchronisation -- the ``Dict[str, float]``
contract of ``policy.learn`` (policy/base_policy.py:25-26).
"""
import copy
import ctypes as C
import os
import struct
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from .. import _lib as L
from .core import GP, AdamT, Mat, Plan, Runtime, get_runtime
from .nets import (GradBuf, Layer, ParamSet, TC_MIN_ROWS, TC_MIN_ROWS_FWD, ens_n_tile, ens_tc_ok, tc_n_tile, tc_ok_dgrad_io,
                   tc_ok_fwd_io, adam_descs, dgrad_problem, fwd_problem, pick_cfg,
                   tc_ok_dgrad, tc_ok_fwd, tc_ok_wgrad, wgrad_problem, wgrad_splits)

MAX_GROUPS = 16
N_LOSS = 32
PRECISIONS = {"fp32": 0, "tf32x3": 3, "tf32": 1}


def linears_of(module: nn.Module) -> List[nn.Linear]:
    return [m for m in module.modules() if isinstance(m, nn.Linear)]


def check_plain_mlp(backbone: nn.Module, what: str) -> None:
    """The engine implements Linear+ReLU stacks without dropout (what the five run scripts build)."""
    for m in backbone.modules():
        if isinstance(m, nn.Dropout):
            raise L.OrlkError(f"{what}: dropout is not supported by the CUDA engine")
        if isinstance(m, (nn.Tanh, nn.Sigmoid, nn.ELU, nn.LeakyReLU, nn.GELU)):
            raise L.OrlkError(f"{what}: only ReLU hidden activations are supported, found {type(m).__name__}")


def adam_hyper(optim: torch.optim.Optimizer) -> Dict[str, float]:
    if not isinstance(optim, torch.optim.Adam):
        raise L.OrlkError(f"the CUDA engine implements torch.optim.Adam only, got {type(optim).__name__}")
    g = optim.param_groups[0]
    if g.get("amsgrad", False) or g.get("weight_decay", 0) != 0 or g.get("maximize", False):
        raise L.OrlkError("Adam options amsgrad / weight_decay / maximize are not supported")
    b1, b2 = g["betas"]
    return dict(lr=float(g["lr"]), beta1=float(b1), beta2=float(b2), eps=float(g["eps"]))


class Learner:
    def __init__(self, device):
        self.rt: Runtime = get_runtime(device)
        self.dev = self.rt.device
        self._groups = (L.AdamGroup * MAX_GROUPS)()
        self._n_groups = 0
        self._group_optim: Dict[int, torch.optim.Optimizer] = {}
        self._group_lr: Dict[int, float] = {}
        self.groups_dev = torch.zeros(MAX_GROUPS * C.sizeof(L.AdamGroup), dtype=torch.uint8, device=self.dev)
        self._lr_host = torch.zeros(MAX_GROUPS, dtype=torch.float32).pin_memory()
        self.loss_dev = self.rt.zeros(N_LOSS)
        self.loss_host = torch.zeros(N_LOSS, dtype=torch.float32).pin_memory()
        self.loss_np = self.loss_host.numpy()
        self.scalars = self.rt.zeros(L.SC_COUNT)
        self.philox_counter = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.noise_enable = torch.ones(1, dtype=torch.int32, device=self.dev)
        self._noise_flag = [True]       # host mirror of noise_enable (a list: shared by the per-batch-size siblings)
        self.plans: Dict[str, Plan] = {}
        self.use_graph = True
        self.steps_done = 0
        # GEMM precision of the wide layers: "fp32" = SIMT FFMA, "tf32x3" = tensor cores with hi/lo operand split
        # (fp32-grade, default), "tf32" = single-pass TF32 tensor cores (fast mode, looser tolerance)
        self.precision = os.environ.get("ORLK_PRECISION", "tf32x3")
        if self.precision not in PRECISIONS:
            raise L.OrlkError(f"unknown precision {self.precision!r}; choose one of {sorted(PRECISIONS)}")
        self.param_sets: List[ParamSet] = []
        self._launch_sync = L.load().orlk_graph_launch_sync
        self._launch_wait = L.load().orlk_graph_launch_wait_event
        # recorded by a node of the step graph right behind the loss block's device-to-host copy: ``learn`` returns on it
        self._loss_ev = C.c_void_p()
        L.call("orlk_event_create_notiming", C.byref(self._loss_ev))
        self.early_return = os.environ.get("ORLK_EARLY_RETURN", "1") != "0"

    # ------------------------------------------------------------------ Adam groups
    def add_group(self, optim: Optional[torch.optim.Optimizer], tau: float = 0.0, **hyper) -> int:
        g = self._n_groups
        assert g < MAX_GROUPS
        h = adam_hyper(optim) if optim is not None else hyper
        e = self._groups[g]
        e.lr, e.beta1, e.beta2, e.eps, e.tau, e.step = h["lr"], h["beta1"], h["beta2"], h["eps"], tau, 0
        if optim is not None:
            self._group_optim[g] = optim
            optim._opt_called = True      # the engine applies this optimiser's steps; keeps lr_scheduler's order check quiet
        self._group_lr[g] = h["lr"]
        self._n_groups += 1
        return g

    def push_groups(self) -> None:
        for g in range(self._n_groups):
            self._groups[g].refresh()
        raw = torch.frombuffer(bytearray(bytes(self._groups)), dtype=torch.uint8)
        self.groups_dev.copy_(raw.to(self.dev))

    def sync_lr(self) -> None:
        """lr schedulers mutate ``optim.param_groups`` between epochs (run_iql.py:132-135): re-read and upload."""
        for g, opt in self._group_optim.items():
            lr = opt.param_groups[0]["lr"]
            if lr != self._group_lr[g]:
                lr = float(lr)
                self._group_lr[g] = lr
                self._lr_host[g] = lr
                L.call("orlk_memcpy_h2d_async", self.groups_dev.data_ptr() + g * C.sizeof(L.AdamGroup),
                       self._lr_host.data_ptr() + 4 * g, 4, self.rt.cur)

    def group_steps(self) -> List[int]:
        raw = self.groups_dev.cpu().numpy().tobytes()
        sz = C.sizeof(L.AdamGroup)
        return [struct.unpack_from("i", raw, g * sz + 20)[0] for g in range(self._n_groups)]

    @property
    def groups_ptr(self) -> int:
        return self.groups_dev.data_ptr()

    # ------------------------------------------------------------------ noise
    def set_noise_enabled(self, on: bool) -> None:
        if on != self._noise_flag[0]:
            self.noise_enable.fill_(1 if on else 0)
            self._noise_flag[0] = on

    # ------------------------------------------------------------------ other batch sizes
    def _rebatch(self, B: int, **kw) -> None:
        """Everything that depends on the batch size (staging, noise block, lazily the activations and the step graphs).
        Called by ``__init__`` and by ``for_batch``."""
        raise NotImplementedError

    def for_batch(self, B: int, **kw) -> "Learner":
        """A sibling engine for another batch size (the reference's ``learn`` takes any batch, policy/base_policy.py:25).
        It SHARES the parameter arenas, Adam moments and step counters, the scalar block (alpha), the loss block and the
        Philox counter with this engine and owns its own staging memory, activations and captured step graphs."""
        sib = copy.copy(self)
        sib.plans = {}
        sib.__dict__.pop("_gather_plans", None)
        sib._rebatch(int(B), **kw)
        return sib

    @property
    def tc_passes(self) -> int:
        return PRECISIONS[self.precision]

    def mlp_run(self, ps: ParamSet, M: int, n_hidden: int, need_grad: bool, **kw) -> "MlpRun":
        """MlpRun with this learner's precision mode (tensor-core passes for the eligible wide layers)."""
        kw.setdefault("tc_passes", self.tc_passes)
        return MlpRun(self.rt, ps, M, n_hidden, need_grad, **kw)

    def refresh(self) -> None:
        """Host-side writes to the parameters (load_state_dict, custom init) invalidate derived copies."""
        for ps in self.param_sets:
            ps.refresh_wt()

    def invalidate(self) -> None:
        """After writing parameters through ``p.data`` (which no version counter sees): re-derive every derived copy."""
        for ps in self.param_sets:
            ps.invalidate()

    # ------------------------------------------------------------------ plan execution
    def run(self, key: str) -> List[float]:
        """Replay the step and wait for it; returns the loss block as Python floats."""
        plan = self.plans[key]
        tok = getattr(self, "_bound_token", None)
        srcs = getattr(self, "_sources", None)
        if srcs:
            # several buffers feed one batch (real + model rows): their uploads and gathers ride in front of the step when
            # every draw is still pending, else the rows are gathered eagerly from the indices already on the device
            if self.use_graph and all(t.pending for t, _ in srcs):
                plan = self._with_gather_parts(key, plan, srcs)
            else:
                for t, off in srcs:
                    if t.pending:
                        t.upload_op()
                        L.call("orlk_event_record", t.pin_event, self.rt.cur)
                        t.pin_armed = True
                    self.gather_part_op(t, off)()
        elif tok is not None and getattr(tok, "pending", False):
            if self.use_graph:      # index upload + gather ride in front of the step, inside the same graph launch
                plan = self._with_gather(key, plan, tok)
                tok.pending = False
            else:
                tok.materialise()
        if self.use_graph:
            if plan.graph is None:
                plan.capture()
            if self.early_return and getattr(plan, "loss_event", False):
                # launch + wait for the loss block only (one host call): the losses are final long before the step is, and
                # everything the caller can do next - draw a batch, launch the next step, read parameters through torch -
                # is ordered behind this graph on the same stream
                rc = self._launch_wait(plan.graph, self.rt.cur, self._loss_ev)
            else:
                rc = self._launch_sync(plan.graph, self.rt.cur)      # launch + stream sync in one host call
            if rc:
                L.check(rc, "orlk_graph_launch_sync")
        else:
            plan.run_eager()
            self.rt.sync()
        self.steps_done += 1
        return self.loss_np.tolist()

    def _with_gather(self, key: str, plan: Plan, tok) -> Plan:
        """The step plan with the replay buffer's pending index upload and row gather as its first two launches."""
        cache = self.__dict__.setdefault("_gather_plans", {})
        k = (key, id(tok), tok.gather_args)
        p2 = cache.get(k)
        if p2 is None:
            p2 = Plan(self.rt, getattr(plan, "name", key) + "+gather")
            head = [("idx_h2d", tok.upload_op), ("gather", tok.gather_op)]
            p2.ops = head + list(plan.ops)
            p2.flat_ops = head + list(plan.flat_ops)
            p2.keep = list(plan.keep) + [tok, plan]
            p2.loss_event = getattr(plan, "loss_event", False)
            cache[k] = p2
        return p2

    def _with_gather_parts(self, key: str, plan: Plan, srcs) -> Plan:
        """The step plan behind the pending index uploads and row gathers of all its source buffers."""
        cache = self.__dict__.setdefault("_gather_plans", {})
        k = (key,) + tuple((id(t), t.gather_args, off) for t, off in srcs)
        p2 = cache.get(k)
        if p2 is None:
            p2 = Plan(self.rt, getattr(plan, "name", key) + "+gather")
            head = []
            for i, (t, off) in enumerate(srcs):
                head += [(f"idx_h2d{i}", t.upload_op), (f"gather{i}", self.gather_part_op(t, off))]
            p2.ops = head + list(plan.ops)
            p2.flat_ops = head + list(plan.flat_ops)
            p2.keep = list(plan.keep) + [t for t, _ in srcs] + [plan]
            p2.loss_event = getattr(plan, "loss_event", False)
            cache[k] = p2
        return p2

    def gather_part_op(self, tok, row_off: int):
        """Launch closure: rows of ``tok``'s buffer at its uploaded indices -> rows [row_off, row_off + n) of this
        engine's batch staging.  Engines that accept multi-buffer batches implement it."""
        raise NotImplementedError

    def enqueue(self, key: str) -> None:
        """Launch a step without waiting for it (used by the device-resident benchmark loop)."""
        tok = getattr(self, "_bound_token", None)
        if tok is not None and getattr(tok, "pending", False):
            tok.materialise()
        plan = self.plans[key]
        if self.use_graph:
            plan.launch()
        else:
            plan.run_eager()
        self.steps_done += 1

    # ------------------------------------------------------------------ K steps behind one host synchronisation
    def next_key(self) -> str:
        """Plan the next step replays (engines with step variants override; called once per step, in order)."""
        return "step"

    def result_of(self, out: Sequence[float]) -> Dict[str, float]:
        """The reference's loss dict from one step's loss block (engines override)."""
        raise NotImplementedError

    def learn_many(self, buffer, n_steps: int) -> List[Dict[str, float]]:
        """``n_steps`` times ``learn(buffer.sample(B))`` with ONE host synchronisation at the end instead of one per step.

        Valid whenever nothing else consumes ``np.random`` between the draws (true inside a trainer epoch):
        ``np.random.randint(0, n, size=(K, B))`` is bit-for-bit K successive ``randint(0, n, B)`` calls (SURVEY.md section
        8c, fact i), so the batches, the Philox noise and therefore every parameter are identical to K single calls; the
        K index rows are uploaded once, each step is one gather launch + one graph replay, and the K loss blocks are read
        back together.  Returns the K loss dicts in order."""
        B = self.B
        outs: List[Dict[str, float]] = []
        if n_steps <= 0:
            return outs
        if not getattr(self, "_built", False) or getattr(self, "_bound_token", None) is None:
            outs.append(self.step(buffer.sample(B)))        # builds the plans and binds them to the buffer's staging rows
            n_steps -= 1
            if n_steps == 0:
                return outs
        stage = buffer._stages.get(B)
        bound = getattr(self, "_bound_ptrs", None)
        if stage is None or bound is None or bound[0] != stage.obs2.data_ptr() or not self.use_graph:
            return outs + [self.step(buffer.sample(B)) for _ in range(n_steps)]     # foreign staging: the plain path
        idx = np.random.randint(0, buffer._size, size=(n_steps, B))
        pin = torch.from_numpy(idx).pin_memory()
        idx_dev = pin.to(self.dev, non_blocking=True)
        self.set_noise(None)
        self.sync_lr()
        self.refresh()
        losses = torch.empty(n_steps, N_LOSS, dtype=torch.float32, device=self.dev)
        keys = []
        src = self.loss_dev.data_ptr()
        for t in range(n_steps):
            buffer.gather_device(idx_dev[t])
            key = self.next_key()
            keys.append(key)
            plan = self.plans[key]
            if plan.graph is None:
                plan.capture()
            L.call("orlk_graph_launch", plan.graph, self.rt.cur)
            L.call("orlk_memcpy_d2d_async", losses.data_ptr() + 4 * N_LOSS * t, src, 4 * N_LOSS, self.rt.cur)
            self.steps_done += 1
        host = losses.cpu().tolist()                        # the one synchronisation
        self._many_keys = keys
        for t in range(n_steps):
            self._many_t = t
            outs.append(self.result_of(host[t]))
        return outs

    def emit_loss_readback(self, plan: Plan) -> None:
        """Copy the loss block to pinned host memory on a detached branch (call once no later launch writes it)."""
        ld, lh = C.c_void_p(self.loss_dev.data_ptr()), C.c_void_p(self.loss_host.data_ptr())

        def op():
            L.call("orlk_memcpy_d2h_async", lh, ld, 4 * N_LOSS, self.rt.cur)
            L.call("orlk_event_record_external", self._loss_ev, self.rt.cur)
        plan.detach("losses_d2h", op)
        plan.loss_event = True

    def finish_ops(self, plan: Plan, group_mask: int) -> None:
        gp, cp = C.c_void_p(self.groups_ptr), C.c_void_p(self.philox_counter.data_ptr())
        plan.add("step_end", lambda: L.call("orlk_step_end", gp, group_mask, cp, self.rt.cur))
        if plan.has_detached:
            plan.join_detached()
        else:
            self.emit_loss_readback(plan)
            plan.join_detached()


# --------------------------------------------------------------------------------------------------------------
# Reusable schedule fragments for ReLU MLP member groups
# --------------------------------------------------------------------------------------------------------------
class MlpRun:
    """Activation / gradient buffers of one forward(+backward) pass of a ParamSet over M rows.

    ``tc_passes`` = 0 runs every GEMM on the SIMT fp32 kernel; 1 / 3 routes eligible wide layers to the tcgen05
    kernel (TF32, or 3xTF32 fp32-grade).  The tensor-core weight gradient wants K-major operands, i.e. the
    TRANSPOSED activations / gradients; those are emitted by the producing kernels' epilogues into HT / dZT.
    """

    def __init__(self, rt: Runtime, ps: ParamSet, M: int, n_hidden: int, need_grad: bool, store: str = "P",
                 tc_passes: int = 0, members: Optional[int] = None, share_forward: Optional["MlpRun"] = None,
                 fused_min_rows: int = TC_MIN_ROWS):
        self.ps, self.M, self.nh, self.store = ps, M, n_hidden, store
        self.tc = tc_passes if M >= TC_MIN_ROWS_FWD else 0
        # arithmetic of the small-row kernel: 0 = fp32 FFMA (default: exact fp32, and as fast in the step because those
        # launches are bound by launch / prologue latency, not by the k loop), 3 = 3xTF32 MMAs, 1 = TF32 (ORLK_TINY_MMA=1)
        self.passes = tc_passes if os.environ.get("ORLK_TINY_MMA", "0") == "1" else 0
        if CHAIN_ON:
            self.passes = tc_passes
        self.Mt = (M + 3) // 4 * 4
        self.G = G = members if members is not None else ps.G      # the first `members` members of the ParamSet
        lays = ps.layers
        # ``share_forward``: a second gradient chain over the SAME activations (EDAC's input-gradient pass)
        self.H = share_forward.H if share_forward is not None else [rt.zeros(G, M, lays[l].out_dim) for l in range(n_hidden)]
        self.dZ = [rt.zeros(G, M, lays[l].out_dim) for l in range(n_hidden)] if need_grad else None
        self.has_head = len(lays) > n_hidden
        self.NS = lays[n_hidden].out_dim if self.has_head else 0
        self.out = (share_forward.out if share_forward is not None else rt.zeros(G, M, self.NS)) if self.has_head else None
        self.dOut = rt.zeros(G, M, self.NS) if (self.has_head and need_grad) else None
        # per-layer kernel choice
        self.tc_fwd = [bool(self.tc) and l >= 1 and tc_ok_fwd(lays[l], M) for l in range(n_hidden)]
        self.tc_dgrad = [bool(self.tc) and need_grad and l >= 1 and tc_ok_dgrad(lays[l], M) for l in range(n_hidden)]
        self.tc_wgrad = [bool(self.tc) and need_grad and l >= 1 and tc_ok_wgrad(lays[l], M) for l in range(n_hidden)]
        # ensembles of short members ('io' weights): forward and input gradients of the hidden layers as ONE n-tiled
        # tensor-core launch over all members (the weights are an MN-major B operand forward and a K-major one backward,
        # so no transposed copies exist); their weight gradients stay on the grouped small-row kernel
        self.ens_tc = bool(tc_passes) and ens_tc_ok(lays[:n_hidden], G, M)
        if self.ens_tc:
            self.tc = tc_passes
            self.tc_fwd = [l >= 1 and tc_ok_fwd_io(lays[l]) and ens_n_tile(G, M, lays[l].out_dim) > 0 for l in range(n_hidden)]
            self.tc_dgrad = [need_grad and l >= 1 and tc_ok_dgrad_io(lays[l]) and ens_n_tile(G, M, lays[l].in_dim) > 0
                             for l in range(n_hidden)]
        # The tensor-core weight gradient dW[o][i] = sum_m dZ[m][o] H[m][i] reads the ROW-MAJOR gradients / activations
        # as MN-major operands (orlk_tc_gemm a_mn / b_mn); only layers whose input width is not a multiple of 32 still
        # need the transposed copies HT / dZT written by the producing kernels' epilogues.
        self.wgrad_mn = [self.tc_wgrad[l] and lays[l].in_dim % 32 == 0 and lays[l].out_dim % 4 == 0
                         and os.environ.get("ORLK_WGRAD_MN", "1") != "0" for l in range(n_hidden)]
        self.HT = [None] * n_hidden
        self.dZT = [None] * n_hidden
        for l in range(n_hidden):
            if self.tc_wgrad[l] and not self.wgrad_mn[l]:
                self.dZT[l] = rt.zeros(G, lays[l].out_dim, self.Mt)
                self.HT[l - 1] = rt.zeros(G, lays[l - 1].out_dim, self.Mt)
        # Scalar head over a tensor-core last hidden layer: dZ[last] = dOut (x) w_head * relu'(H[last]) is rank-1 times a
        # mask, so the dgrad / wgrad GEMMs that consume it build it in shared memory from H / H^T (orlk_tc_gemm's
        # operand generator) and the head_dgrad launch with its two 16 MB outputs disappears.
        last = n_hidden - 1
        self.fuse_head_bwd = (need_grad and self.has_head and self.NS == 1 and n_hidden >= 2 and share_forward is None
                              and self.tc_dgrad[last] and self.tc_wgrad[last] and lays[n_hidden].layout == "oi"
                              and lays[last].out_dim % 4 == 0 and self.Mt == M
                              and lays[n_hidden].w_off % 4 == 0 and lays[n_hidden].w_gs % 4 == 0
                              and os.environ.get("ORLK_FUSE_HEAD_BWD", "1") != "0")
        if self.fuse_head_bwd:
            if not self.wgrad_mn[last]:
                self.HT[last] = rt.zeros(G, lays[last].out_dim, self.Mt)
            self.dZT[last] = None           # never materialised
        if any(self.tc_dgrad) and not self.ens_tc:
            ps.enable_wt([l for l in range(n_hidden) if self.tc_dgrad[l]])
        # the whole forward pass (all hidden layers + scalar head) of a long-row Linear+ReLU stack as ONE tensor-core launch
        # that keeps the activations on the SM between layers (csrc/orlk_fused.cu); fp32-grade (3xTF32) mode only
        widths = {lays[l].out_dim for l in range(n_hidden)}
        self.keep_h = True              # False: the fused forward does not store the activations (set by the owner)
        self.fused_fwd = (FUSED_FWD and tc_passes == 3 and M >= fused_min_rows and not self.ens_tc and share_forward is None
                          and 2 <= n_hidden <= L.FUSED_MAX_LAYERS and self.has_head and self.NS == 1
                          and all(lay.layout == "oi" for lay in lays[:n_hidden + 1])
                          and len(widths) == 1 and (lays[0].out_dim == 32 or lays[0].out_dim % 64 == 0) and 32 <= lays[0].out_dim <= 256
                          and lays[0].in_dim <= 32 and all(lays[l].in_dim == lays[0].out_dim for l in range(1, n_hidden + 1))
                          and all(h is None for h in self.HT) and ps.block % 4 == 0
                          and all(lays[l].w_off % 4 == 0 and lays[l].w_gs == ps.block and lays[l].b_gs == ps.block
                                  for l in range(n_hidden + 1)))
        # ... and its input-gradient chain (scalar head folded in, relu' from the decision bits the forward pass leaves)
        self.fused_bwd = (self.fused_fwd and need_grad and FUSED_BWD and self.fuse_head_bwd
                          and all(self.tc_dgrad[l] for l in range(1, n_hidden)) and all(t is None for t in self.dZT))
        self.relu_bits = rt.zeros(n_hidden, G, 8, M, dtype=torch.int32) if self.fused_bwd else None
        self.bits_written = False
        # streaming kernels for the narrow first layer / narrow head at large row counts (any precision mode)
        # streaming kernels for the narrow first layer (K <= 32) and the narrow head's weight gradient
        self.narrow0 = lays[0].layout == "oi" and lays[0].in_dim <= 32
        self.narrow_head = self.has_head and lays[n_hidden].layout == "oi" and self.NS <= 32 and M >= 128
        self.chunks = L.load().orlk_narrow_wgrad_chunks(M)
        self.w0_part = self.b0_part = self.hw_part = self.hb_part = None
        big = M >= TC_MIN_ROWS          # the chunked weight-gradient kernels only pay off for long reductions
        if need_grad and self.narrow0 and big:
            self.w0_part = rt.zeros(self.chunks, G, lays[0].out_dim, lays[0].in_dim)
            self.b0_part = rt.zeros(self.chunks, G, lays[0].out_dim)
        if need_grad and self.narrow_head and big:
            self.hw_part = rt.zeros(self.chunks, G, self.NS, lays[n_hidden].in_dim)
            self.hb_part = rt.zeros(self.chunks, G, self.NS)

    def h(self, l: int, g: int) -> Mat:
        return Mat.of(self.H[l][g])

    def dz(self, l: int, g: int) -> Mat:
        return Mat.of(self.dZ[l][g])

    def ht(self, l: int, g: int) -> Optional[Mat]:
        return None if self.HT[l] is None else Mat(self.HT[l][g].data_ptr(), self.HT[l].shape[1], self.M, self.Mt)

    def dzt(self, l: int, g: int) -> Optional[Mat]:
        return None if self.dZT[l] is None else Mat(self.dZT[l][g].data_ptr(), self.dZT[l].shape[1], self.M, self.Mt)


def _grouped(t: torch.Tensor, rows: int, cols: int, ld: int) -> Mat:
    """Member 0 of a [G, rows, ld] tensor as a Mat (the tensor-core launcher strides over the members itself)."""
    return Mat(t.data_ptr(), rows, cols, ld, t)


# Small-row passes (actor / target / policy-improvement chains) as ONE cluster launch per pass (csrc/orlk_chain.cu): 8 CTAs
# per 16-row strip, a hardware cluster barrier + an L2 re-read of the 16 KB strip between stages instead of a kernel
# boundary.  Round 2's rewrite (round 1: 4 CTAs per 32-row strip, scalar DSMEM pushes, 6.6 us per stage) runs a stage in
# 3.0 us, a four-stage pass in 17-18 us against ~22 us for four PDL-chained launches -- but the step loses the fused
# head+sampler / backward-entry launches and the twin-critic passes run two CTAs per SM, so CQL measures 299.6 us with it
# and 291.8 us without (profiles/chain_trace_r02.txt): OFF by default, ORLK_CHAIN=1 enables it (tensor-core modes only).
CHAIN_ON = os.environ.get("ORLK_CHAIN", "0") == "1"
CHAIN_MAX_DESC = 24


def chainable(run: "MlpRun", with_head: bool) -> bool:
    """Small-row pass whose layers all fit the fused chain kernel (csrc/orlk_chain.cu)."""
    lays = run.ps.layers[:run.nh + (1 if with_head else 0)]
    return (CHAIN_ON and run.passes != 0 and run.M < TC_MIN_ROWS and len(lays) <= 8
            and all(lay.layout == "oi" and lay.in_dim <= 256 and lay.out_dim <= 256 for lay in lays))


FUSED_FWD = os.environ.get("ORLK_FUSED_FWD", "1") != "0"
FUSED_BWD = os.environ.get("ORLK_FUSED_BWD", "1") != "0"
WGRAD_CHAIN = os.environ.get("ORLK_WGRAD_CHAIN", "0") == "1"      # measured slower (262 vs 255 us): off


def emit_lo_refresh(rt: Runtime, plan: Plan, ps: ParamSet, store: str) -> None:
    """Derived operand copies of a parameter arena for the fused passes (orlk_fused_prep: lo words + padded first layer);
    once per plan and store, placed by the caller after the last update of that arena and before its first fused use."""
    done = plan.__dict__.setdefault("_lo_fresh", set())
    key = (id(ps), store)
    if key in done:
        return
    done.add(key)
    plan.keep.append(ps)
    lay0 = ps.layers[0]
    if store == "WT":       # the transposed copies feed the backward chain only: no first layer there
        plan.add(f"{ps.name}.WT.fused_prep", rt.fused_prep(ps.WT, ps.lo_arena("WT")))
        return
    plan.add(f"{ps.name}.{store}.fused_prep", rt.fused_prep(getattr(ps, store), ps.lo_arena(store), W0=ps.w(0, 0, store),
                                                            gs=lay0.w_gs, N=lay0.out_dim, K0=lay0.in_dim, G=ps.G,
                                                            w0pad=ps.w0_pad(store)))


def emit_lo_refresh_many(rt: Runtime, plan: Plan, items: Sequence[Tuple[ParamSet, str]]) -> None:
    """``emit_lo_refresh`` for several (ParamSet, store) pairs in ONE launch (orlk_fused_prep_multi)."""
    done = plan.__dict__.setdefault("_lo_fresh", set())
    jobs = []
    for ps, store in items:
        key = (id(ps), store)
        if key in done:
            continue
        done.add(key)
        plan.keep.append(ps)
        lay0 = ps.layers[0]
        if store == "WT":
            jobs.append(dict(src=ps.WT, dst_lo=ps.lo_arena("WT")))
        else:
            jobs.append(dict(src=getattr(ps, store), dst_lo=ps.lo_arena(store), W0=ps.w(0, 0, store), gs=lay0.w_gs,
                             N=lay0.out_dim, K0=lay0.in_dim, G=ps.G, w0pad=ps.w0_pad(store)))
    if jobs:
        plan.add("fused_prep", rt.fused_prep_multi(jobs))


def fused_fwd_job(rt: Runtime, run: "MlpRun", X: Mat) -> dict:
    """``run``'s whole forward pass (hidden layers + scalar head, all members) as one job of a fused launch."""
    ps, nh, M, G = run.ps, run.nh, run.M, run.G
    N, K0 = ps.layers[0].out_dim, ps.layers[0].in_dim
    pad = ps.w0_pad(run.store)
    return rt.fused_fwd_job(
        X=Mat(X.ptr, M, K0, X.ld), W0pad=pad[0].data_ptr(), W0pad_lo=pad[1].data_ptr(),
        W=[0] + [ps.w(l, 0, run.store) for l in range(1, nh)], Wlo=[0] + [ps.w_lo(l, 0, run.store) for l in range(1, nh)],
        bias=[ps.b(l, 0, run.store) for l in range(nh)], H=[run.H[l].data_ptr() for l in range(nh)] if run.keep_h else None,
        gs=ps.block, h_gs=M * N, head_w=ps.w(nh, 0, run.store), head_b=ps.b(nh, 0, run.store), out=run.out.data_ptr(),
        out_gs=M * run.NS, M=M, N=N, K0=K0, G=G, relu_bits=run.relu_bits.data_ptr() if run.fused_bwd else 0)


def fusable_x(run: "MlpRun", X: Sequence[Mat]) -> bool:
    return (run.fused_fwd and all(x.ptr == X[0].ptr and x.ld == X[0].ld for x in X) and X[0].ld % 4 == 0
            and X[0].ptr % 16 == 0)


def emit_forward_pair(rt: Runtime, plan: Plan, run_a: "MlpRun", Xa: Sequence[Mat], tag_a: str, run_b: "MlpRun",
                      Xb: Sequence[Mat], tag_b: str) -> bool:
    """Two independent forward passes (e.g. the online critics on the big batch and the target critics on the next-state
    rows) as the two jobs of ONE fused launch; False (nothing emitted) when either pass is not fusable."""
    if not (fusable_x(run_a, Xa) and fusable_x(run_b, Xb)):
        return False
    plan.keep += [run_a, run_b, [x.keep for x in Xa], [x.keep for x in Xb]]
    emit_lo_refresh(rt, plan, run_a.ps, run_a.store)
    emit_lo_refresh(rt, plan, run_b.ps, run_b.store)
    plan.add(f"{tag_a}+{tag_b}.fwd_fused.tc", rt.critic_fwd_fused([fused_fwd_job(rt, run_a, Xa[0]), fused_fwd_job(rt, run_b, Xb[0])]))
    run_a.bits_written = run_b.bits_written = True
    return True


# the streaming first-layer kernel pays off for long row counts; short passes take the small-row GEMM (one k pass)
NARROW_MIN_ROWS = int(os.environ.get("ORLK_NARROW_MIN_ROWS", "1024"))


def emit_forward(rt: Runtime, plan: Plan, run: MlpRun, X: Sequence[Mat], tag: str, skip_head: bool = False) -> None:
    """Hidden layers (+bias+ReLU fused) as tcgen05 or grouped SIMT GEMMs, then the narrow head (warp per row).
    ``skip_head``: the caller evaluates the head itself (fused with the sampler, orlk_head_sample)."""
    ps, G, M = run.ps, run.G, run.M
    plan.keep += [run, [x.keep for x in X]]
    if not skip_head and chainable(run, with_head=run.has_head) and all(h is None for h in run.HT):
        # the whole pass (hidden layers + head) as fused chain launches: one cluster per 32-row strip and member
        n_st = run.nh + (1 if run.has_head else 0)
        chains = []
        for g in range(G):
            st = [fwd_problem(ps, l, g, X[g] if l == 0 else run.h(l - 1, g), run.h(l, g), L.EPI_RELU, run.store)
                  for l in range(run.nh)]
            if run.has_head:
                st.append(fwd_problem(ps, run.nh, g, run.h(run.nh - 1, g), Mat.of(run.out[g]), L.EPI_NONE, run.store))
            chains.append(st)
        per = max(1, CHAIN_MAX_DESC // n_st)
        for c0 in range(0, G, per):
            plan.add(f"{tag}.fwd_chain" + (f"{c0}" if c0 else ""), rt.gemm_chain(chains[c0:c0 + per], run.passes, passes0=3))
        return
    if not skip_head and fusable_x(run, X):
        emit_lo_refresh(rt, plan, ps, run.store)        # (a no-op when the caller has placed it earlier in the step)
        plan.add(f"{tag}.fwd_fused.tc", rt.critic_fwd_fused([fused_fwd_job(rt, run, X[0])]))
        run.bits_written = True
        return
    for l in range(run.nh):
        lay = ps.layers[l]
        if run.tc_fwd[l] and run.ens_tc:
            K, N = lay.in_dim, lay.out_dim
            plan.add(f"{tag}.fwd{l}.tc", rt.tc_gemm(
                A=_grouped(run.H[l - 1], M, K, K), a_gs=M * K, B=Mat(ps.w(l, 0, run.store), K, N, N), b_gs=lay.w_gs, b_mn=True,
                G=G, passes=run.tc, n_tile=ens_n_tile(G, M, N), epi=L.EPI_RELU, C=_grouped(run.H[l], M, N, N), c_gs=M * N,
                bias=ps.b(l, 0, run.store), bias_gs=lay.b_gs))
            continue
        if run.tc_fwd[l]:
            K, N = lay.in_dim, lay.out_dim
            plan.add(f"{tag}.fwd{l}.tc", rt.tc_gemm(
                A=_grouped(run.H[l - 1], M, K, K), a_gs=M * K, B=Mat(ps.w(l, 0, run.store), N, K, K), b_gs=lay.w_gs, G=G,
                passes=run.tc, n_tile=tc_n_tile(M, N), epi=L.EPI_RELU, C=_grouped(run.H[l], M, N, N), c_gs=M * N,
                CT=_grouped(run.HT[l], N, M, run.Mt) if run.HT[l] is not None else None, ct_gs=N * run.Mt,
                bias=ps.b(l, 0, run.store), bias_gs=lay.b_gs))
            continue
        same_x = all(x.ptr == X[0].ptr and x.ld == X[0].ld for x in X)
        if (l == 0 and run.ens_tc and same_x and X[0].ld % 4 == 0 and X[0].ptr % 16 == 0 and lay.out_dim % 32 == 0
                and lay.out_dim <= 256 and lay.w_gs % 4 == 0 and ens_n_tile(G, M, lay.out_dim) > 0):
            # first layer of an ensemble on a shared input: one zero-padded k-slab, weights MN-major as stored
            K, N = lay.in_dim, lay.out_dim
            plan.add(f"{tag}.fwd0.tc", rt.tc_gemm(
                A=Mat(X[0].ptr, M, K, X[0].ld), a_gs=0, B=Mat(ps.w(0, 0, run.store), K, N, N), b_gs=lay.w_gs, b_mn=True, G=G,
                passes=3, n_tile=ens_n_tile(G, M, N), epi=L.EPI_RELU, C=_grouped(run.H[0], M, N, N), c_gs=M * N,
                bias=ps.b(0, 0, run.store), bias_gs=lay.b_gs))
            continue
        if (l == 0 and run.tc and run.narrow0 and M >= TC_MIN_ROWS and same_x and X[0].ld % 4 == 0 and X[0].ptr % 16 == 0
                and lay.out_dim % 16 == 0 and lay.out_dim <= 256 and os.environ.get("ORLK_TC_FWD0", "1") != "0"):
            # obs+act wide first layer on the tensor cores: one zero-padded k-slab, the kernel is all epilogue
            K, N = lay.in_dim, lay.out_dim
            plan.add(f"{tag}.fwd0.tc", rt.tc_gemm(
                A=Mat(X[0].ptr, M, K, X[0].ld), a_gs=0, B=Mat(ps.w(0, 0, run.store), N, K, K), b_gs=lay.w_gs, G=G,
                passes=3,       # raw observations: always fp32-grade (a single k-slab, the extra MMAs are free)
                epi=L.EPI_RELU, C=_grouped(run.H[0], M, N, N), c_gs=M * N,
                CT=_grouped(run.HT[0], N, M, run.Mt) if run.HT[0] is not None else None, ct_gs=N * run.Mt,
                bias=ps.b(0, 0, run.store), bias_gs=lay.b_gs))
            continue
        if l == 0 and run.narrow0 and M >= NARROW_MIN_ROWS and same_x:
            ht = run.HT[0]
            args = (X[0].ptr, X[0].ld, 0, ps.w(0, 0, run.store), lay.in_dim, lay.w_gs, ps.b(0, 0, run.store), lay.b_gs,
                    run.H[0].data_ptr(), lay.out_dim, M * lay.out_dim, ht.data_ptr() if ht is not None else None, run.Mt,
                    lay.out_dim * run.Mt, M, lay.out_dim, lay.in_dim, G, 1)
            plan.add(f"{tag}.fwd0.narrow", lambda args=args: L.call("orlk_narrow_fwd", *args, rt.cur))
            continue
        probs = [fwd_problem(ps, l, g, X[g] if l == 0 else run.h(l - 1, g), run.h(l, g), L.EPI_RELU, run.store,
                             YT=run.ht(l, g)) for g in range(G)]
        # (the first layer sees raw observations: fp32-grade even in the single-pass mode)
        plan.add(f"{tag}.fwd{l}", rt.gemm(probs, pick_cfg(M * G, lay.out_dim, rows_per_problem=M), passes=(3 if (l == 0 and run.passes) else run.passes)))
    if run.has_head and not skip_head:
        emit_head_forward(rt, plan, run, tag)


def emit_head_forward(rt: Runtime, plan: Plan, run: MlpRun, tag: str) -> None:
    ps, G, l = run.ps, run.G, run.nh
    lay = ps.layers[l]
    hin = run.H[l - 1]
    K = lay.in_dim
    if lay.layout == "io":
        assert lay.out_dim == 1, "ensemble heads wider than 1 go through the GEMM path"
    args = (hin.data_ptr(), K, run.M * K, ps.w(l, 0, run.store), K, 1, lay.w_gs, ps.b(l, 0, run.store), lay.b_gs,
            run.out.data_ptr(), run.NS, run.M * run.NS, run.M, K, run.NS, G)
    plan.add(f"{tag}.head", lambda: L.call("orlk_skinny_fwd", *args, rt.cur))


def emit_head_dgrad(rt: Runtime, plan: Plan, run: MlpRun, tag: str) -> None:
    """dZ[last hidden] = (dOut W_head) * relu'(H[last hidden])  (+ its transpose for a tensor-core wgrad)."""
    ps, G, l = run.ps, run.G, run.nh
    if run.fuse_head_bwd:
        return                  # folded into the consumers' operand generator
    if chainable(run, with_head=True) and all(t is None for t in run.dZT):
        run.pending_head_dgrad = True       # becomes the first stage of the chain emitted by emit_hidden_dgrad
        return
    lay = ps.layers[l]
    K = lay.in_dim
    hmask = run.H[l - 1]
    dzt = run.dZT[l - 1]
    args = (run.dOut.data_ptr(), run.NS, run.M * run.NS, ps.w(l, 0), K, lay.w_gs, hmask.data_ptr(), K, run.M * K,
            run.dZ[l - 1].data_ptr(), K, run.M * K, dzt.data_ptr() if dzt is not None else None, run.Mt, K * run.Mt,
            run.M, K, run.NS, G)
    plan.add(f"{tag}.head_dgrad", lambda: L.call("orlk_skinny_dgrad", *args, rt.cur))


def emit_hidden_dgrad(rt: Runtime, plan: Plan, run: MlpRun, tag: str, down_to: int = 1, dact=None,
                      from_layer: Optional[int] = None) -> None:
    """dZ[l-1] = (dZ[l] W_l) * relu'(H[l-1]) for l = nh-1 .. down_to.  ``dact = (dA, col0, ncols)`` appends
    dA[g] = dZ[0][g] . W0[:, col0:col0+ncols] (the gradient w.r.t. some input columns, emit_dact) to the pass."""
    ps, G, M = run.ps, run.G, run.M
    if getattr(run, "pending_head_dgrad", False):
        run.pending_head_dgrad = False
        chains = []
        for g in range(G):
            st = [dgrad_problem(ps, run.nh, g, Mat.of(run.dOut[g]), run.dz(run.nh - 1, g), L.EPI_RELU_MASK, run.h(run.nh - 1, g))]
            st += [dgrad_problem(ps, l, g, run.dz(l, g), run.dz(l - 1, g), L.EPI_RELU_MASK, run.h(l - 1, g))
                   for l in range(run.nh - 1, down_to - 1, -1)]
            if dact is not None:
                assert down_to == 1
                dA, col0, ncols = dact
                st.append(dgrad_problem(ps, 0, g, run.dz(0, g), Mat.of(dA[g]), L.EPI_NONE, None, col0=col0, ncols=ncols))
            chains.append(st)
        if dact is not None:
            plan.keep.append(dact[0])
        per = max(1, CHAIN_MAX_DESC // len(chains[0]))
        for c0 in range(0, G, per):
            plan.add(f"{tag}.bwd_chain" + (f"{c0}" if c0 else ""), rt.gemm_chain(chains[c0:c0 + per], run.passes))
        return
    if dact is not None:
        raise L.OrlkError("emit_hidden_dgrad(dact=...) needs the fused chain path; call emit_dact instead")
    if run.fused_bwd and run.bits_written and down_to == 1:
        emit_lo_refresh(rt, plan, ps, "WT")         # (a no-op when the caller has placed it earlier in the step)
        nh, N = run.nh, ps.layers[0].out_dim
        head = ps.layers[nh]
        plan.add(f"{tag}.dgrad_fused.tc", rt.critic_bwd_fused(
            dq=run.dOut.data_ptr(), dq_gs=M, head_w=ps.w(nh, 0), relu_bits=run.relu_bits.data_ptr(),
            WT=[0] + [ps.wt(l, 0) for l in range(1, nh)], WTlo=[0] + [ps._ptr(ps.lo_arena("WT"), ps.layers[l].w_off) for l in range(1, nh)],
            dZ=[run.dZ[l].data_ptr() for l in range(nh - 1)], gs=ps.block, dz_gs=M * N, M=M, N=N, G=G))
        return
    for l in range(run.nh - 1 if from_layer is None else from_layer, down_to - 1, -1):
        lay = ps.layers[l]
        if run.tc_dgrad[l] and run.ens_tc:
            # dX[m][i] = sum_o dZ[m][o] W[i][o]: the 'io' weight is the K-major B operand as it is stored
            K, N = lay.out_dim, lay.in_dim
            plan.add(f"{tag}.dgrad{l}.tc", rt.tc_gemm(
                A=_grouped(run.dZ[l], M, K, K), a_gs=M * K, B=Mat(ps.w(l, 0), N, K, K), b_gs=lay.w_gs, G=G, passes=run.tc,
                n_tile=ens_n_tile(G, M, N), epi=L.EPI_RELU_MASK, C=_grouped(run.dZ[l - 1], M, N, N), c_gs=M * N,
                aux=_grouped(run.H[l - 1], M, N, N), aux_gs=M * N))
            continue
        if run.tc_dgrad[l]:
            K, N = lay.out_dim, lay.in_dim
            gen = {}
            a_src = run.dZ[l]
            if run.fuse_head_bwd and l == run.nh - 1:
                head = ps.layers[run.nh]
                a_src = run.H[l]            # the generator turns relu(H) into dOut[m] * w_head[k] * (H > 0)
                gen = dict(gen_row=run.dOut.data_ptr(), gen_row_gs=M, gen_col=ps.w(run.nh, 0), gen_col_gs=head.w_gs)
            plan.add(f"{tag}.dgrad{l}.tc", rt.tc_gemm(
                A=_grouped(a_src, M, K, K), a_gs=M * K, B=Mat(ps.wt(l, 0), N, K, K), b_gs=lay.w_gs, G=G, **gen,
                passes=run.tc, n_tile=tc_n_tile(M, N), epi=L.EPI_RELU_MASK, C=_grouped(run.dZ[l - 1], M, N, N), c_gs=M * N,
                CT=_grouped(run.dZT[l - 1], N, M, run.Mt) if run.dZT[l - 1] is not None else None, ct_gs=N * run.Mt,
                aux=_grouped(run.H[l - 1], M, N, N), aux_gs=M * N))
            continue
        probs = [dgrad_problem(ps, l, g, run.dz(l, g), run.dz(l - 1, g), L.EPI_RELU_MASK, run.h(l - 1, g),
                               dXT=run.dzt(l - 1, g)) for g in range(G)]
        plan.add(f"{tag}.dgrad{l}", rt.gemm(probs, pick_cfg(M * G, lay.in_dim, rows_per_problem=M), passes=run.passes))


def emit_dact(rt: Runtime, plan: Plan, run: MlpRun, dA: torch.Tensor, col0: int, ncols: int, tag: str) -> None:
    """dL/d(input columns col0 .. col0+ncols) of the first layer: dA[g] = dZ0[g] . W0[:, cols]  (a narrow product)."""
    ps, G, M = run.ps, run.G, run.M
    lay = ps.layers[0]
    K = lay.out_dim
    if lay.layout == "oi":      # W0[o][i]: element (n=a, k=o) at W0 + (col0+a) + o*in
        w, ldw, w_sk = ps.w(0, 0) + 4 * col0, 1, lay.in_dim
    else:                       # W0[i][o]: element (n=a, k=o) at W0 + (col0+a)*out + o
        w, ldw, w_sk = ps.w(0, 0) + 4 * col0 * lay.out_dim, lay.out_dim, 1
    args = (run.dZ[0].data_ptr(), K, M * K, w, ldw, w_sk, lay.w_gs, None, 0, dA.data_ptr(), ncols, M * ncols, M, K, ncols, G)
    plan.add(f"{tag}.dact", lambda: L.call("orlk_skinny_fwd", *args, rt.cur))
    plan.keep.append(dA)


TC_WGRAD_SPLITS = 31


def wgrad_layout(ps: ParamSet, n_layers: int, M: int, tc_layers: Sequence[bool] = ()) -> List[Tuple[int, int]]:
    """(tile config or -1 for the tensor-core kernel, split-K factor) per layer for a reduction over M rows."""
    out = []
    for l in range(n_layers):
        lay = ps.layers[l]
        if l < len(tc_layers) and tc_layers[l]:
            out.append((-1, L.load().orlk_tc_effective_splits(M, TC_WGRAD_SPLITS)))
            continue
        out_r, out_c = (lay.out_dim, lay.in_dim) if lay.layout == "oi" else (lay.in_dim, lay.out_dim)
        if M < TC_MIN_ROWS:        # short reductions: whole-k tiles, no split-K partials for Adam to re-read
            out.append((L.CFG_TINY, 1))
            continue
        cfg = L.CFG_BIG if (out_r >= 128 and out_c >= 128 and M >= 2048) else L.CFG_SMALL
        BM, BN, _ = L.CFG_TILES[cfg]
        tiles = (-(-out_r // BM)) * (-(-out_c // BN)) * ps.G
        out.append((cfg, wgrad_splits(tiles, M, cfg)))
    return out


def polyak_descs(ps: ParamSet) -> List[AdamT]:
    """Target-only update tgt <- (1-tau) tgt + tau p over the whole ParamSet (one descriptor per member block)."""
    return [AdamT(p=ps._ptr(ps.P, 0), n=ps.total, group=ps.group_ids[0], tgt=ps._ptr(ps.T, 0), flags=L.OPT_POLYAK)]


def make_gradbuf(rt: Runtime, ps: ParamSet, runs: Sequence["MlpRun"]) -> GradBuf:
    """GradBuf with enough split slots for the weight-gradient reductions of the given passes."""
    slots = 1
    for run in runs:
        n_l = run.nh + (1 if run.has_head else 0)
        slots = max([slots] + [s for _, s in wgrad_layout(ps, n_l, run.M, run.tc_wgrad)])
    return GradBuf(rt, ps, slots)


def emit_wgrad_adam(rt: Runtime, plan: Plan, run: MlpRun, X: Sequence[Mat], gb: GradBuf, groups_ptr: int, tag: str,
                    polyak: bool, only_layers: Optional[Sequence[int]] = None, do_wgrad: bool = True, do_adam: bool = True,
                    tail_ops: Optional[Dict[int, Callable[[], None]]] = None) -> None:
    """All weight / bias gradients of the pass (split-K partials) and the fused Adam(+polyak) update.
    ``only_layers``: just these layers (the caller schedules the others elsewhere, e.g. as soon as their dZ exists);
    ``do_wgrad`` / ``do_adam``: only the gradient launches / only the update (which must not run before the layer's own
    input-gradient launch has read the old weights).  ``tail_ops[l]()`` emits further launches on the branch of layer l's
    gradient launch, behind its update."""
    ps, G, M = run.ps, run.G, run.M
    plan.keep += [run, gb, [x.keep for x in X]]
    big, small, tiny = [], [], []
    n_l = run.nh + (1 if run.has_head else 0)
    layout = wgrad_layout(ps, n_l, M, run.tc_wgrad)
    splits = [s for _, s in layout] + [1] * (len(ps.layers) - n_l)
    grad_src = {}
    launches: List[Tuple[str, Callable[[], None]]] = []       # mutually independent: run on parallel graph branches
    launch_layers: List[List[int]] = []                       # the layers whose gradients each launch produces
    grouped_layers = {L.CFG_BIG: [], L.CFG_SMALL: [], L.CFG_TINY: []}
    for l in range(n_l):
        if only_layers is not None and l not in only_layers:
            continue
        cfg, s = layout[l]
        lay = ps.layers[l]
        same_x = all(x.ptr == X[0].ptr and x.ld == X[0].ld for x in X)
        if (l == 0 and run.tc and run.narrow0 and M >= TC_MIN_ROWS and same_x and X[0].ld % 4 == 0 and X[0].ptr % 16 == 0
                and lay.out_dim % 4 == 0 and os.environ.get("ORLK_TC_WGRAD0", "1") != "0"):
            # dW0[o][i] = sum_m dZ0[m][o] X[m][i] on the tensor cores: both operands row-major as they are (MN-major
            # tiles), one X shared by the members, 32-column n-tile with the columns past in_dim zero
            s0 = L.load().orlk_tc_effective_splits(M, TC_WGRAD_SPLITS)
            assert s0 <= gb.n_slots, (s0, gb.n_slots)
            splits[0] = s0
            launches.append((f"{tag}.wgrad0.tc", rt.tc_gemm(
                A=_grouped(run.dZ[0], M, lay.out_dim, lay.out_dim), a_gs=M * lay.out_dim, a_mn=True,
                B=Mat(X[0].ptr, M, lay.in_dim, X[0].ld), b_gs=0, b_mn=True, n_tile=32, G=G, passes=3,
                C=Mat(gb.ptr(lay.w_off), lay.out_dim, lay.in_dim, lay.in_dim), c_gs=lay.w_gs, c_split_stride=gb.stride,
                rowsum=gb.ptr(lay.b_off), rowsum_gs=lay.b_gs, rowsum_split_stride=gb.stride, k_splits=s0)))
            launch_layers.append([0])
            continue
        if l == 0 and run.w0_part is not None and same_x:
            o, i = lay.out_dim, lay.in_dim
            args = (run.dZ[0].data_ptr(), o, M * o, X[0].ptr, X[0].ld, 0, run.w0_part.data_ptr(), 1, i, o * i, G * o * i,
                    run.b0_part.data_ptr(), o, G * o, None, 0, 0, M, o, i, G)
            launches.append((f"{tag}.wgrad0.narrow", lambda args=args: L.call("orlk_narrow_wgrad", *args, rt.cur)))
            launch_layers.append([0])
            grad_src[(0, "w")] = (run.w0_part.data_ptr(), o * i, G * o * i, run.chunks)
            grad_src[(0, "b")] = (run.b0_part.data_ptr(), o, G * o, run.chunks)
            continue
        if l == run.nh and run.hw_part is not None:
            K, NS = lay.in_dim, run.NS
            args = (run.H[l - 1].data_ptr(), K, M * K, run.dOut.data_ptr(), NS, M * NS, run.hw_part.data_ptr(), K, 1, NS * K,
                    G * NS * K, None, 0, 0, run.hb_part.data_ptr(), NS, G * NS, M, K, NS, G)
            launches.append((f"{tag}.wgrad_head.narrow", lambda args=args: L.call("orlk_narrow_wgrad", *args, rt.cur)))
            launch_layers.append([l])
            grad_src[(l, "w")] = (run.hw_part.data_ptr(), NS * K, G * NS * K, run.chunks)
            grad_src[(l, "b")] = (run.hb_part.data_ptr(), NS, G * NS, run.chunks)
            continue
        assert s <= gb.n_slots, (s, gb.n_slots)
        if cfg == -1:
            # dW[o,i] = sum_m dZ^T[o,m] * H^T[i,m]; bias gradient = row sums of dZ^T (a ones-tile MMA)
            gen = {}
            fused = run.fuse_head_bwd and l == run.nh - 1
            if fused:       # dZ^T[o][m] = w_head[o] * dOut[m] * (H[m][o] > 0), built inside the kernel
                head = ps.layers[run.nh]
                gen = dict(gen_row=ps.w(run.nh, 0), gen_row_gs=head.w_gs, gen_col=run.dOut.data_ptr(), gen_col_gs=M)
            if run.wgrad_mn[l]:
                a_src = run.H[l] if fused else run.dZ[l]
                operands = dict(A=_grouped(a_src, M, lay.out_dim, lay.out_dim), a_gs=M * lay.out_dim, a_mn=True,
                                B=_grouped(run.H[l - 1], M, lay.in_dim, lay.in_dim), b_gs=M * lay.in_dim, b_mn=True)
            else:
                a_src = run.HT[l] if fused else run.dZT[l]
                operands = dict(A=_grouped(a_src, lay.out_dim, M, run.Mt), a_gs=lay.out_dim * run.Mt,
                                B=_grouped(run.HT[l - 1], lay.in_dim, M, run.Mt), b_gs=lay.in_dim * run.Mt)
            launches.append((f"{tag}.wgrad{l}.tc", rt.tc_gemm(
                **operands, **gen, G=G, passes=run.tc,
                C=Mat(gb.ptr(lay.w_off), lay.out_dim, lay.in_dim, lay.in_dim), c_gs=lay.w_gs, c_split_stride=gb.stride,
                rowsum=gb.ptr(lay.b_off), rowsum_gs=lay.b_gs, rowsum_split_stride=gb.stride, k_splits=s)))
            launch_layers.append([l])
            continue
        for g in range(G):
            xin = X[g] if l == 0 else run.h(l - 1, g)
            dy = run.dz(l, g) if l < run.nh else Mat.of(run.dOut[g])
            {L.CFG_BIG: big, L.CFG_SMALL: small, L.CFG_TINY: tiny}[cfg].append(wgrad_problem(ps, gb, l, g, xin, dy, s))
        grouped_layers[cfg].append(l)
    if big:
        launches.append((f"{tag}.wgrad_big", rt.gemm(big, L.CFG_BIG)))
        launch_layers.append(grouped_layers[L.CFG_BIG])
    if small:
        launches.append((f"{tag}.wgrad_small", rt.gemm(small, L.CFG_SMALL)))
        launch_layers.append(grouped_layers[L.CFG_SMALL])
    if tiny:
        launches.append((f"{tag}.wgrad_tiny", rt.gemm(tiny, L.CFG_TINY)))
        launch_layers.append(grouped_layers[L.CFG_TINY])

    def adam_for(layers):
        return rt.adam(adam_descs(ps, gb, splits, polyak, layers=layers, grad_src=grad_src, members=range(G)), groups_ptr)

    if not (do_wgrad and do_adam):
        assert only_layers is not None
        if do_wgrad:
            for label, op in launches:
                plan.add(label, op)
        if do_adam:
            plan.add(f"{tag}.adam", adam_for(sorted(set(only_layers))))
        return
    n_tc = sum(1 for label, _ in launches if label.endswith(".tc"))
    if n_tc >= 2 and WGRAD_CHAIN:
        # Experiment (ORLK_WGRAD_CHAIN=1, off): the optimiser launches of the first layers start late because a ready grid that
        # still waits for SMs (the next weight gradient on a parallel branch) is dispatched before any later grid.  Here the
        # GEMMs form ONE chain on the main stream, last layer first, each followed by its Adam(+polyak) launch on a side
        # stream.  Measured: the updates do start earlier, but the GEMMs lose the overlap of their tails (262 vs 255 us).
        tc = sorted((i for i in range(len(launches)) if launches[i][0].endswith(".tc")), key=lambda i: -max(launch_layers[i]))
        simt = [i for i in range(len(launches)) if not launches[i][0].endswith(".tc")]
        side = 1
        for n, i in enumerate(tc):
            plan.add(*launches[i])
            if n == len(tc) - 1:        # the last GEMM's (small) update stays behind it on the main stream
                plan.add(f"{tag}.adam{i}", adam_for(sorted(set(launch_layers[i]))))
                break
            plan.fork()                 # side streams wait for this GEMM
            if n == 0:                  # SIMT weight gradients (scalar head): beside the second GEMM
                for j in simt:
                    plan.branch(side)
                    plan.add(*launches[j])
                    plan.add(f"{tag}.adam{j}", adam_for(sorted(set(launch_layers[j]))))
                    side = side % Plan.N_SIDE + 1
            plan.branch(side)
            plan.add(f"{tag}.adam{i}", adam_for(sorted(set(launch_layers[i]))))
            side = side % Plan.N_SIDE + 1
            plan.branch(0)
        plan.join()
    elif len(launches) > 1:
        # each branch: one weight-gradient launch and right behind it the Adam(+polyak) update of exactly those layers, so
        # the bandwidth-bound optimiser work of one layer overlaps the tensor-core work of the others
        # (a scalar head folded into the operand generator: the last hidden layer's weight gradient READS the head weights,
        # so the head's own update waits for that launch - an event between the two branches, no cost: the GEMM runs first)
        plan.fork()
        ev_gen = None
        for i, (label, op) in enumerate(launches):
            plan.branch(i % (Plan.N_SIDE + 1))
            plan.add(label, op)
            if run.fuse_head_bwd and run.nh - 1 in launch_layers[i]:
                ev_gen = plan.record_event()
            if run.fuse_head_bwd and launch_layers[i] == [run.nh] and ev_gen is not None:
                plan.wait_event(ev_gen)
            plan.add(f"{tag}.adam{i}", adam_for(sorted(set(launch_layers[i]))))
            for l in launch_layers[i]:
                if tail_ops and l in tail_ops:
                    tail_ops[l]()
        plan.join()
    else:
        for label, op in launches:
            plan.add(label, op)
        plan.add(f"{tag}.adam", adam_for(range(n_l) if only_layers is None else sorted(set(only_layers))))
