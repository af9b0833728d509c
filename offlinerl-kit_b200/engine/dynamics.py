"""Device engine for the MOPO ensemble dynamics model: training step, holdout validation and one-step imagination.

Reference: dynamics/ensemble_dynamics.py (learn :178-208, validate :210-217, step :28-79) on
modules/dynamics_module.py:32-119 (E x [in -> hidden x n Swish -> 2*(obs+1)], learned log-variance bounds).
The E members live in one ParamSet ('io' layout, members contiguous per tensor); every layer of every member is
one problem of a grouped GEMM launch; Swish and its derivative are GEMM epilogues.
"""
import ctypes as C
import os
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _lib as L
from .core import AdamT, GP, Mat, Plan, Runtime, get_runtime
from .learner import Learner, N_LOSS
from .nets import (TC_MIN_ROWS, GradBuf, ParamSet, adam_descs, dgrad_problem, fwd_problem, pick_cfg, wgrad_problem,
                   wgrad_splits)


class DynRun:
    """Buffers of one forward (+backward) pass of the ensemble over M rows per member."""

    def __init__(self, rt: Runtime, ps: ParamSet, M: int, need_grad: bool):
        self.ps, self.M = ps, M
        E, nl = ps.G, len(ps.layers)
        self.nh = nl - 1
        self.H = [rt.zeros(E, M, ps.layers[l].out_dim) for l in range(self.nh)]
        self.Z = [rt.zeros(E, M, ps.layers[l].out_dim) for l in range(self.nh)] if need_grad else None
        self.OUT = rt.zeros(E, M, ps.layers[-1].out_dim)
        self.dZ = [rt.zeros(E, M, ps.layers[l].out_dim) for l in range(self.nh)] if need_grad else None
        self.dOUT = rt.zeros(E, M, ps.layers[-1].out_dim) if need_grad else None


def _tc_layer_ok(ps: ParamSet, l: int, x0: Optional[Mat]) -> bool:
    """Can layer l of the ensemble ('io' weights [member][in][out]: an MN-major B operand) run on the tensor-core kernel?"""
    lay = ps.layers[l]
    ok = (lay.layout == "io" and lay.out_dim % 4 == 0 and lay.out_dim <= 256 and ps.w(l, 0) % 16 == 0 and lay.w_gs % 4 == 0
          and ps.b(l, 0) % 16 == 0)
    if l == 0:
        return ok and x0 is not None and x0.ld % 4 == 0 and x0.ptr % 16 == 0
    return ok and lay.in_dim % 4 == 0


def emit_dyn_forward(rt: Runtime, plan: Plan, run: DynRun, X: Callable[[int], Mat], tag: str, tc_passes: int = 0) -> None:
    """tc_passes (1 or 3; 0 = off): inference passes over many rows (rollouts: 50 000 states x 7 members) go through the
    tcgen05 kernel, one launch per layer with the members as groups and the Swish in its epilogue."""
    ps, E, M = run.ps, run.ps.G, run.M
    shared_x = all(X(e).ptr == X(0).ptr and X(e).ld == X(0).ld for e in range(E))
    use_tc = (tc_passes in (1, 3) and run.Z is None and M >= TC_MIN_ROWS and shared_x
              and os.environ.get("ORLK_DYN_TC", "1") != "0"
              and all(_tc_layer_ok(ps, l, X(0)) for l in range(run.nh + 1)))
    for l in range(run.nh + 1):
        last = l == run.nh
        if use_tc:
            lay = ps.layers[l]
            K, N = lay.in_dim, lay.out_dim
            if l == 0:
                a, a_gs = Mat(X(0).ptr, M, K, X(0).ld), 0
            else:
                a, a_gs = Mat(run.H[l - 1].data_ptr(), M, K, K), M * K
            out = run.OUT if last else run.H[l]
            plan.add(f"{tag}.fwd{l}.tc", rt.tc_gemm(
                A=a, a_gs=a_gs, B=Mat(ps.w(l, 0), K, N, N), b_gs=lay.w_gs, b_mn=True, G=E,
                passes=3 if l == 0 else tc_passes,      # raw (scaled) inputs: always fp32-grade, one k-slab
                n_tile=(N + 31) // 32 * 32, epi=L.EPI_NONE if last else L.EPI_SWISH,
                C=Mat(out.data_ptr(), M, N, N), c_gs=M * N, bias=ps.b(l, 0), bias_gs=lay.b_gs))
            continue
        probs = []
        for e in range(E):
            xin = X(e) if l == 0 else Mat.of(run.H[l - 1][e])
            yout = Mat.of(run.OUT[e]) if last else Mat.of(run.H[l][e])
            z = Mat.of(run.Z[l][e]) if (not last and run.Z is not None) else None
            probs.append(fwd_problem(ps, l, e, xin, yout, L.EPI_NONE if last else L.EPI_SWISH, Z=z))
        plan.add(f"{tag}.fwd{l}", rt.gemm(probs, pick_cfg(M * E, ps.layers[l].out_dim, rows_per_problem=M)))


class DynamicsEngine(Learner):
    def __init__(self, model, optim, shard=None):
        """shard = (rank, world, comm): train only members partition_members(E, world)[rank] (BASELINE.json configs[4],
        "members sharded over 8 x B200").  The loss is a sum over members
        (ensemble_dynamics.py:197-199), so the only coupling is the shared max_logvar / min_logvar: every mini-batch the
        ranks all-gather their partial gradients of the two bounds (2 x (obs+1) floats, + the partial loss) and apply the
        same Adam step to their replicas; ``validate`` all-gathers the per-member holdout losses (:145-168).  ``comm`` has
        ``all_gather(send, recv)`` (engine/edac_sharded.py:NcclComm), or is None when a test drives the ranks in lockstep."""
        super().__init__(model.device)
        rt = self.rt
        self.model = model
        layers = list(model.backbones) + [model.output_layer]
        self.E_all = int(layers[0].weight.shape[0])
        self.shard = shard
        self.rank, self.world, self.comm = (0, 1, None) if shard is None else shard
        self.e0, self.e1, self.counts = 0, self.E_all, [self.E_all]
        self._slices = []
        if shard is not None:
            from ..parallel import partition_members
            from .edac_sharded import MemberSlice
            parts = partition_members(self.E_all, self.world)
            if min(len(p) for p in parts) == 0 or self.world > 8:
                raise L.OrlkError(f"cannot shard {self.E_all} members over {self.world} ranks")
            self.counts = [len(p) for p in parts]
            self.e0, self.e1 = parts[self.rank][0], parts[self.rank][-1] + 1
            self._slices = [MemberSlice(lay, self.e0, self.e1) for lay in layers]
            for sl, lay in zip(self._slices, layers):
                sl.weight_decay = float(getattr(lay, "weight_decay", 0.0))
            layers = self._slices
        self.ps = ParamSet.from_ensemble(rt, "dynamics", layers,
                                         extra={"max_logvar": model.max_logvar, "min_logvar": model.min_logvar})
        self.E = self.ps.G
        self.in_dim = self.ps.layers[0].in_dim
        self.D = self.ps.layers[-1].out_dim // 2
        self.wd = [float(getattr(lay, "weight_decay", 0.0)) for lay in layers]
        self.g = self.add_group(optim)
        self.ps.group_ids = [self.g]
        self.push_groups()
        self._learn_plans: Dict[int, Tuple[Plan, dict]] = {}
        self._fwd_runs: Dict[int, Tuple[DynRun, Plan, torch.Tensor]] = {}
        n_dec = sum(self.rt.lib.orlk_sumsq_chunks(self.E * lay.w_numel) for lay in self.ps.layers)
        self.decay_partials = rt.zeros(max(n_dec, 1))
        self.n_decay = n_dec
        self.dmax, self.dmin = rt.zeros(self.D), rt.zeros(self.D)

    # ------------------------------------------------------------------ training step
    def _learn_plan(self, Bn: int):
        if Bn in self._learn_plans:
            return self._learn_plans[Bn]
        rt, ps, E, D = self.rt, self.ps, self.E, self.D
        run = DynRun(rt, ps, Bn, need_grad=True)
        X = rt.zeros(E, Bn, self.in_dim)
        Y = rt.zeros(E, Bn, D)
        nl = len(ps.layers)
        cfgs, splits = [], []
        # short reductions (the reference's 256-row mini-batches): whole-k 32 x 16 tiles, no split-K partials for Adam to
        # re-read -- 13 us per layer against 29 us for the split-K configuration
        wg_tiny = Bn <= L.CFG_TILES[L.CFG_TINY][2] and os.environ.get("ORLK_DYN_WGRAD_TINY", "1") != "0"
        for l in range(nl):
            lay = ps.layers[l]
            cfg = L.CFG_TINY if wg_tiny else L.CFG_SMALL
            BM, BN, _ = L.CFG_TILES[cfg]
            tiles = (-(-lay.in_dim // BM)) * (-(-lay.out_dim // BN)) * E
            splits.append(1 if wg_tiny else wgrad_splits(tiles, Bn, cfg))
            cfgs.append(cfg)
        gb = GradBuf(rt, ps, max(splits))
        plan = Plan(rt, f"dyn.learn{Bn}")
        par = os.environ.get("ORLK_DYN_BRANCHES", "1") != "0"
        # weight-decay term of the reported loss: sum_l wd_l * 0.5 * sum W_l^2 -- reads the weights only, so it runs on a
        # branch beside the forward pass
        if par:
            plan.fork()
            plan.branch(1)
        off = 0
        for l, lay in enumerate(ps.layers):
            n = E * lay.w_numel
            args = (ps._ptr(ps.P, lay.w_off), n, 0.5 * self.wd[l], self.decay_partials.data_ptr() + 4 * off)
            plan.add(f"D.decay{l}", lambda args=args: L.call("orlk_sumsq", *args, rt.cur))
            off += rt.lib.orlk_sumsq_chunks(n)
        if par:
            plan.branch(0)
        emit_dyn_forward(rt, plan, run, lambda e: Mat.of(X[e]), "D")
        if par:
            plan.join()
        nll_scratch = rt.zeros(rt.lib.orlk_dyn_nll_scratch_floats(E, Bn, D))
        largs = [run.OUT.data_ptr(), Y.data_ptr(), E, Bn, D, ps.extra_ptr("max_logvar"), ps.extra_ptr("min_logvar"), 0.01,
                 self.decay_partials.data_ptr(), self.n_decay, run.dOUT.data_ptr(), self.dmax.data_ptr(), self.dmin.data_ptr(),
                 self.loss_dev.data_ptr(), nll_scratch.data_ptr()]
        state = {"coef": 0.01}

        def loss_op():
            largs[7] = state["coef"]
            L.call("orlk_dyn_nll", *largs, rt.cur)
        plan.add("D.loss", loss_op)

        def wgrad_of(l):
            probs = []
            for e in range(E):
                xin = Mat.of(X[e]) if l == 0 else Mat.of(run.H[l - 1][e])
                dy = Mat.of(run.dOUT[e]) if l == run.nh else Mat.of(run.dZ[l][e])
                probs.append(wgrad_problem(ps, gb, l, e, xin, dy, splits[l]))
            return probs

        # backward: dH_last = (dOUT W_out^T) * swish'(Z_last), then down the stack.  The weight gradient of layer l needs
        # only that layer's upstream gradient, so it runs on a branch beside the input gradient of the same layer.
        wprobs = []
        for l in range(run.nh, 0, -1):
            probs = []
            for e in range(E):
                dy = Mat.of(run.dOUT[e]) if l == run.nh else Mat.of(run.dZ[l][e])
                probs.append(dgrad_problem(ps, l, e, dy, Mat.of(run.dZ[l - 1][e]), L.EPI_DSWISH, Mat.of(run.Z[l - 1][e])))
            if par:
                plan.fork()
                plan.branch(1)
                plan.add(f"D.wgrad{l}", rt.gemm(wgrad_of(l), cfgs[l]))
                plan.branch(0)
            else:
                wprobs += wgrad_of(l)
            plan.add(f"D.dgrad{l}", rt.gemm(probs, pick_cfg(Bn * E, ps.layers[l].in_dim, rows_per_problem=Bn)))
            if par:
                plan.join()
        if par:
            plan.add("D.wgrad0", rt.gemm(wgrad_of(0), cfgs[0]))
        else:
            plan.add("D.wgrad", rt.gemm(wprobs + wgrad_of(0), cfgs[0]))
        descs = adam_descs(ps, gb, splits, polyak=False)
        # weight decay enters the gradient as wd_l * W_l (d/dW of wd * 0.5 * sum W^2); biases are not decayed
        k = 0
        for l in range(nl):
            descs[k].wd = self.wd[l]
            k += 2
        gp = C.c_void_p(self.groups_ptr)
        plan2 = send = recv = None
        if self.shard is None:
            for name, gbuf in (("max_logvar", self.dmax), ("min_logvar", self.dmin)):
                o = ps.extra[name][0]
                descs.append(AdamT(p=ps._ptr(ps.P, o), n=D, group=self.g, m=ps._ptr(ps.Mo, o), v=ps._ptr(ps.Vo, o),
                                   grad=gbuf.data_ptr(), g_splits=1, g_split_stride=D))
            plan.add("D.adam", rt.adam(descs, self.groups_ptr))
            plan.add("step_end", lambda: L.call("orlk_step_end", gp, 1 << self.g, None, rt.cur))
        else:
            # members: Adam on the own slice.  Shared bounds: partial gradients (and the partial loss) go to the exchange;
            # after the all-gather the second segment sums the `world` partials in rank order inside the Adam launch
            # (g_splits = world: the same fixed-order reduction as split-K partials) -- identical on every rank.
            blk = 2 * D + 4
            send, recv = rt.zeros(blk), rt.zeros(self.world * blk)
            plan.add("D.adam", rt.adam(descs, self.groups_ptr))
            for src, off, n in ((self.dmax, 0, D), (self.dmin, D, D), (self.loss_dev, 2 * D, 1)):
                args = (send.data_ptr() + 4 * off, src.data_ptr(), 4 * n)
                plan.add("X.stage", lambda args=args: L.call("orlk_memcpy_d2d_async", *args, rt.cur))
            plan2 = Plan(rt, f"dyn.learn{Bn}.bounds")
            bdescs = []
            for name, off in (("max_logvar", 0), ("min_logvar", D)):
                o = ps.extra[name][0]
                bdescs.append(AdamT(p=ps._ptr(ps.P, o), n=D, group=self.g, m=ps._ptr(ps.Mo, o), v=ps._ptr(ps.Vo, o),
                                    grad=recv.data_ptr() + 4 * off, g_splits=self.world, g_split_stride=blk))
            plan2.add("D.adam_bounds", rt.adam(bdescs, self.groups_ptr))
            plan2.add("step_end", lambda: L.call("orlk_step_end", gp, 1 << self.g, None, rt.cur))
            plan2.keep += [send, recv]
        plan.keep += [gb, run, X, Y, nll_scratch]
        self._learn_plans[Bn] = (plan, dict(X=X, Y=Y, run=run, state=state, plan2=plan2, send=send, recv=recv, blk=2 * D + 4))
        return self._learn_plans[Bn]

    def learn_batch_begin(self, src_x, src_y, idx, r0: int, Bn: int, coef: float):
        """Gather mini-batch [r0, r0+Bn) of the own members and run the step up to the exchange (all of it when the
        ensemble is not sharded).  idx is the FULL [E_all, n] bootstrap index matrix."""
        rt, E = self.rt, self.E
        plan, st = self._learn_plan(Bn)
        # the direct coef * (sum max_logvar - sum min_logvar) term belongs to the ensemble, not to a member: rank 0 carries it
        eff = float(coef) if self.rank == 0 else 0.0
        X, Y = st["X"], st["Y"]
        iptr = idx.data_ptr() + idx.element_size() * self.e0 * idx.stride(0)
        L.call("orlk_gather_rows", src_x.data_ptr(), src_x.stride(0), self.in_dim, iptr, idx.stride(0), r0, E, Bn,
               X.data_ptr(), self.in_dim, Bn * self.in_dim, rt.cur)
        L.call("orlk_gather_rows", src_y.data_ptr(), src_y.stride(0), self.D, iptr, idx.stride(0), r0, E, Bn,
               Y.data_ptr(), self.D, Bn * self.D, rt.cur)
        if self.use_graph and st.get("graph_coef", eff) == eff:
            st["state"]["coef"] = eff
            st["graph_coef"] = eff      # (the logvar coefficient is baked into the captured launch arguments)
            plan.launch()
        else:
            st["state"]["coef"] = eff
            plan.run_eager()
        return st

    def learn_batch_end(self, st, losses: torch.Tensor, b: int) -> None:
        """Behind the exchange: the shared bounds' Adam step from the gathered partial gradients; record the loss(es)."""
        rt = self.rt
        if self.shard is None:
            L.call("orlk_memcpy_d2d_async", losses.data_ptr() + 4 * b, self.loss_dev.data_ptr(), 4, rt.cur)
            return
        plan2 = st["plan2"]
        plan2.launch() if self.use_graph else plan2.run_eager()
        recv, blk, D2 = st["recv"], st["blk"], 2 * self.D
        for r in range(self.world):     # partial losses of all ranks (summed on the host at the end of the pass)
            L.call("orlk_memcpy_d2d_async", losses.data_ptr() + 4 * (b * self.world + r), recv.data_ptr() + 4 * (r * blk + D2), 4,
                   rt.cur)

    def learn(self, src_x: torch.Tensor, src_y: torch.Tensor, idx: torch.Tensor, batch_size: int, coef: float) -> float:
        """One pass over idx [E, n] (row ids into src_x [N,in] / src_y [N,D]) in mini-batches of ``batch_size``."""
        n = idx.shape[1]
        nb = -(-n // batch_size)
        losses = torch.zeros(nb * self.world, dtype=torch.float32, device=self.dev)
        self.sync_lr()
        if self.shard is not None and self.comm is None:
            raise L.OrlkError("a sharded DynamicsEngine needs a communicator (or a lockstep driver calling learn_batch_*)")
        for b in range(nb):
            r0 = b * batch_size
            st = self.learn_batch_begin(src_x, src_y, idx, r0, min(batch_size, n - r0), coef)
            if self.shard is not None:
                self.comm.all_gather(st["send"], st["recv"])
            self.learn_batch_end(st, losses, b)
        return self.pass_loss(losses, nb)

    def pass_loss(self, losses: torch.Tensor, nb: int) -> float:
        return float(losses.cpu().numpy().astype(np.float64).reshape(nb, self.world).sum(1).mean())

    # ------------------------------------------------------------------ member-sharded bookkeeping
    def write_back(self) -> None:
        for s in self._slices:
            s.write_back()

    @torch.no_grad()
    def read_back(self) -> None:
        """The full model's rows of the own members -> the slices (after ``model.load_save()``)."""
        for s in self._slices:
            lay = s.src[0]
            s.weight.data.copy_(lay.weight.data[s.e0:s.e1])
            s.bias.data.copy_(lay.bias.data[s.e0:s.e1])

    def gather_members(self, local: torch.Tensor) -> torch.Tensor:
        """[E_r, ...] per rank -> [E_all, ...] on every rank (one padded all-gather)."""
        if self.shard is None:
            return local
        e_max = max(self.counts)
        per = local[0].numel()
        send = torch.zeros(e_max * per, dtype=local.dtype, device=local.device)
        send[:local.numel()].copy_(local.reshape(-1))
        recv = torch.empty(self.world * e_max * per, dtype=local.dtype, device=local.device)
        self.comm.all_gather(send, recv)
        parts = [recv[r * e_max * per:r * e_max * per + c * per] for r, c in enumerate(self.counts)]
        return torch.cat(parts).view((self.E_all,) + tuple(local.shape[1:]))

    @torch.no_grad()
    def gather_all(self) -> None:
        """Every rank's full model gets every member's current weights (and saved_* copies) from its owner."""
        self.write_back()
        if self.shard is None or self.comm is None:
            return
        self.rt.sync()
        for s in self._slices:
            lay = s.src[0]
            for name in ("weight", "bias"):
                full = self.gather_members(getattr(s, name).data)
                getattr(lay, name).data.copy_(full)
                saved = getattr(lay, "saved_" + name, None)
                if saved is not None:
                    saved.data.copy_(full)

    # ------------------------------------------------------------------ inference
    def _forward(self, x: torch.Tensor) -> DynRun:
        """Ensemble forward on a shared [S, in] device input (already scaled)."""
        S = x.shape[0]
        self._forward_alloc(S)
        run, plan, xbuf = self._fwd_runs[S]
        if x.data_ptr() != xbuf.data_ptr():
            xbuf.copy_(x)
        plan.run_eager()
        return run

    def input_buffer(self, S: int) -> torch.Tensor:
        self._forward_alloc(S)
        return self._fwd_runs[S][2]

    def _forward_alloc(self, S: int) -> None:
        if S not in self._fwd_runs:
            if len(self._fwd_runs) > 8:
                self._fwd_runs.clear()
            run = DynRun(self.rt, self.ps, S, need_grad=False)
            xbuf = self.rt.zeros(S, (self.in_dim + 3) // 4 * 4)[:, :self.in_dim]   # 16-byte rows: a TMA operand
            plan = Plan(self.rt, f"dyn.fwd{S}")
            xm = Mat.of(xbuf)
            emit_dyn_forward(self.rt, plan, run, lambda e: xm, "F", tc_passes=self.tc_passes)
            self._fwd_runs[S] = (run, plan, xbuf)

    def validate_local(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        run = self._forward(x)
        mse = torch.zeros(self.E, dtype=torch.float32, device=self.dev)
        L.call("orlk_dyn_val_mse", run.OUT.data_ptr(), y.data_ptr(), self.E, x.shape[0], self.D, mse.data_ptr(), self.rt.cur)
        return mse

    def validate(self, x: torch.Tensor, y: torch.Tensor) -> List[float]:
        """Per-member holdout MSE of ALL members (sharded: the own members' losses all-gathered, ensemble_dynamics.py:145)."""
        mse = self.validate_local(x, y)
        if self.shard is not None:
            if self.comm is None:
                raise L.OrlkError("a sharded DynamicsEngine needs a communicator (or a lockstep driver calling validate_local)")
            mse = self.gather_members(mse)
        return list(mse.cpu().numpy())

    def imagine(self, obs: torch.Tensor, act: torch.Tensor, mu: torch.Tensor, sd: torch.Tensor, term_kind: int,
                penalty_coef: float, noise64: Optional[torch.Tensor], midx: Optional[torch.Tensor],
                noise32: Optional[torch.Tensor], pick_u: Optional[torch.Tensor], elites: Optional[torch.Tensor],
                uncertainty_mode: int = 0, out=None):
        """One imagined transition for S device-resident states; returns device tensors.
        ``out`` = (next_obs [S,O], reward [S,1], terminal uint8 [S,1], raw_reward [S,1], penalty [S,1]): write there instead
        of allocating (the sync-free rollout loop pre-allocates the whole horizon)."""
        rt, S, O = self.rt, obs.shape[0], obs.shape[1]
        A = act.shape[1]
        self._forward_alloc(S)
        xbuf = self._fwd_runs[S][2]
        L.call("orlk_dyn_input", obs.data_ptr(), obs.stride(0), act.data_ptr(), act.stride(0), mu.data_ptr(), sd.data_ptr(), S,
               O, A, xbuf.data_ptr(), xbuf.stride(0), rt.cur)
        run = self._forward(xbuf)
        if out is not None:
            nobs, rew, term, raw, pen = out
        else:
            nobs = torch.empty(S, O, dtype=torch.float32, device=self.dev)
            rew, raw, pen = (torch.empty(S, 1, dtype=torch.float32, device=self.dev) for _ in range(3))
            term = torch.empty(S, 1, dtype=torch.uint8, device=self.dev)
        ptr = lambda t: t.data_ptr() if t is not None else None
        L.call("orlk_dyn_step", run.OUT.data_ptr(), self.E, S, self.D, self.ps.extra_ptr("max_logvar"),
               self.ps.extra_ptr("min_logvar"), obs.data_ptr(), obs.stride(0), ptr(noise64), ptr(midx), ptr(noise32), ptr(pick_u),
               ptr(elites), int(elites.numel()) if elites is not None else 0, term_kind, float(penalty_coef), int(uncertainty_mode),
               nobs.data_ptr(), rew.data_ptr(), raw.data_ptr(), pen.data_ptr(), term.data_ptr(), rt.cur)
        return nobs, rew, term, raw, pen
