"""Device-resident model rollouts for MOPO and COMBO (reference: policy/model_based/mopo.py:45-79, combo.py:67-107 +
sac.py:79-86 + dynamics/ensemble_dynamics.py:28-79).  The reference ping-pongs every imagined step through the host (50 MB of model
outputs per step, 6.3 M float64 normals on one core); here only the survivor count (4 bytes) crosses per step and the
transitions are copied out once at the end."""
import ctypes as C
import time
from typing import Dict, List, Optional

import numpy as np
import torch

from .. import _lib as L
from .core import Mat, Plan, get_runtime
from .learner import MlpRun, emit_forward, linears_of
from .nets import ParamSet


class RolloutEngine:
    def __init__(self, policy, uniform=None):
        """uniform: (low, high) -- COMBO's uniform_rollout: actions are uniform draws instead of the actor's samples."""
        self.policy = policy
        self.uniform = uniform
        self.dyn = policy.dynamics
        self.rt = get_runtime(policy.actor.device)
        self.dev = self.rt.device
        eng = getattr(policy, "_engine", None)
        self._own_ps = None
        if eng is not None:
            self.actor_ps = eng.actor_ps            # share the learner's arena: the rollout sees the trained weights
        else:
            self._own_ps = self.actor_ps = ParamSet.from_linear_members(self.rt, "actor", [linears_of(policy.actor)], fuse_last=2)
        self.nh = len(self.actor_ps.layers) - 1
        self.O = self.actor_ps.layers[0].in_dim
        self.A = self.actor_ps.layers[-1].out_dim // 2
        self._plans: Dict[int, tuple] = {}
        self.count_dev = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.philox_counter = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.tc_passes = getattr(eng, "tc_passes", 3) if eng is not None else 3

    def _actor_plan(self, S: int):
        if S not in self._plans:
            if len(self._plans) > 8:
                self._plans.clear()
            run = MlpRun(self.rt, self.actor_ps, S, self.nh, need_grad=False, tc_passes=self.tc_passes)
            obs = torch.zeros(S, self.O, dtype=torch.float32, device=self.dev)
            plan = Plan(self.rt, f"rollout.actor{S}")
            emit_forward(self.rt, plan, run, [Mat.of(obs)], "R.actor")
            self._plans[S] = (run, plan, obs)
        return self._plans[S]

    # ------------------------------------------------------------------ the default path: no host round trip per step
    def _workspace(self, S: int, h: int):
        key = (S, h)
        ws = self.__dict__.setdefault("_ws", {})
        if key not in ws:
            if len(ws) > 4:
                ws.clear()
            dev, O, A = self.dev, self.O, self.A
            D = self.dyn.engine.D
            f = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
            ws[key] = dict(OBS=f(h + 1, S, O), ACT=f(h, S, A), REW=f(h, S, 1), RAW=f(S, 1), PEN=f(S, 1),
                           TERM=torch.empty(h, S, 1, dtype=torch.uint8, device=dev),
                           DEAD=torch.zeros(h + 1, S, dtype=torch.uint8, device=dev), EPS=f(S, A), NZ=f(S * D + S))
        return ws[key]

    def _run_async(self, init_obss: np.ndarray, length: int, device_out: bool):
        """mopo.py:45-79 without a device->host round trip per imagined step.  The reference compacts the survivors after
        every step, which needs the survivor count on the host (a stream sync + a 4-byte read per step, and step buffers of
        a new size each time).  Here every step runs on ALL start rows into buffers pre-allocated for the whole horizon,
        with a per-row dead flag (dead[t+1] = dead[t] | terminal[t]); rows that are dead at step t are dropped when the
        step's transitions are assembled at the end -- a stable selection, so the rows of step t+1 are exactly the
        surviving rows of step t, in order, as in the reference.  While nothing terminates the result is bit-identical to the
        per-step loop (same Philox draws); once rows die the two consume different draws for the later steps (this loop
        also draws for dead rows), i.e. they are equal in distribution (tests/test_gpu_dynamics.py).  The host waits once,
        for the dead counts.  (Dead rows still cost arithmetic: 0 % for halfcheetah, ~15 % for walker2d.)"""
        rt, O, A = self.rt, self.O, self.A
        dyn = self.dyn
        t_start = time.perf_counter()
        S, h = int(len(init_obss)), int(length)
        ws = self._workspace(S, h)
        OBS, ACT, REW, TERM, DEAD = ws["OBS"], ws["ACT"], ws["REW"], ws["TERM"], ws["DEAD"]
        OBS[0].copy_(torch.as_tensor(init_obss, dtype=torch.float32), non_blocking=False)
        DEAD[0].zero_()
        for t in range(h):
            cur = OBS[t]
            if self.uniform is not None:        # combo.py:81-86: a ~ U(low, high), no actor pass
                L.call("orlk_philox_fill", ACT[t].data_ptr(), 0, S * A, self.uniform[0], self.uniform[1], 0xac7,
                       self.philox_counter.data_ptr(), None, rt.cur)
                L.call("orlk_step_end", dyn.engine.groups_ptr, 0, self.philox_counter.data_ptr(), rt.cur)
            else:
                run, plan, obs_buf = self._actor_plan(S)
                obs_buf.copy_(cur)
                plan.run_eager()
                eps = ws["EPS"]
                L.call("orlk_philox_fill", eps.data_ptr(), S * A, 0, 0.0, 1.0, 0xac7, self.philox_counter.data_ptr(), None, rt.cur)
                L.call("orlk_step_end", dyn.engine.groups_ptr, 0, self.philox_counter.data_ptr(), rt.cur)
                L.call("orlk_tanh_gauss_sample", run.out.data_ptr(), 2 * A, 0, 1, eps.data_ptr(), S, A, ACT[t].data_ptr(), A, None,
                       None, 0, 0, None, 0, rt.cur)
            dyn.step_device(cur, ACT[t], out=(OBS[t + 1], REW[t], TERM[t], ws["RAW"], ws["PEN"]), noise_buf=ws["NZ"])
            torch.bitwise_or(DEAD[t], TERM[t].view(-1), out=DEAD[t + 1])
        # ---- assemble: the only host wait of the loop
        n_dead = DEAD[:h].sum(dim=1, dtype=torch.int64).cpu().tolist()
        t_loop = time.perf_counter()
        if sum(n_dead) == 0:
            res = {"obss": OBS[:h].reshape(h * S, O), "next_obss": OBS[1:h + 1].reshape(h * S, O), "actions": ACT.reshape(h * S, A),
                   "rewards": REW.reshape(h * S, 1), "terminals": TERM.reshape(h * S, 1)}
            if device_out:
                res = {k: v.clone() for k, v in res.items()}          # the workspace is reused by the next rollout
        else:
            parts = {k: [] for k in ("obss", "next_obss", "actions", "rewards", "terminals")}
            for t in range(h):
                if n_dead[t] == S:
                    break                       # nobody left: the reference stops here (mopo.py:71-72)
                srcs = (OBS[t], OBS[t + 1], ACT[t], REW[t], TERM[t])
                if n_dead[t] == 0:
                    for k, v in zip(parts, srcs):
                        parts[k].append(v)
                else:
                    keep = (DEAD[t] == 0).nonzero().squeeze(1)
                    for k, v in zip(parts, srcs):
                        parts[k].append(v.index_select(0, keep))
            res = {k: torch.cat(v, 0) for k, v in parts.items()}
        n_total = int(res["obss"].shape[0])
        r_mean = float(res["rewards"].double().mean().item()) if n_total else 0.0
        if device_out:
            self.last_timing = {"loop_ms": 1e3 * (t_loop - t_start), "export_ms": 0.0}
            return res, {"num_transitions": n_total, "reward_mean": r_mean}
        # export through pinned host memory (torch's caching host allocator recycles the blocks between rollouts; the
        # arrays handed out keep their block alive, so they are never overwritten by a later call)
        out = {}
        for k, v in res.items():
            hbuf = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
            hbuf.copy_(v, non_blocking=True)
            out[k] = hbuf
        rt.sync()
        torch.cuda.current_stream(self.dev).synchronize()
        out = {k: v.numpy() for k, v in out.items()}
        out["terminals"] = out["terminals"].astype(bool)
        self.last_timing = {"loop_ms": 1e3 * (t_loop - t_start), "export_ms": 1e3 * (time.perf_counter() - t_loop)}
        return out, {"num_transitions": n_total, "reward_mean": r_mean}

    def run(self, init_obss: np.ndarray, length: int, noise: Optional[Dict[str, List[np.ndarray]]] = None,
            device_out: bool = False):
        """device_out: return the transitions as device tensors (terminals uint8) instead of the reference's NumPy
        arrays -- for callers that keep working on the device (parallel.rollout_state_sharded)."""
        rt, O, A = self.rt, self.O, self.A
        dyn = self.dyn
        E, D = dyn.engine.E, dyn.engine.D
        if getattr(dyn.terminal_fn, "device_kind", None) is None:
            raise L.OrlkError("device rollouts need a termination function with a device_kind (utils/termination_fns.py)")
        eng = getattr(self.policy, "_engine", None)
        if eng is not None and eng.actor_ps is not self.actor_ps:
            # the learner was created after this engine and has re-homed the actor's parameters in its own arena
            self.actor_ps = eng.actor_ps
            self.tc_passes = eng.tc_passes
            self._plans.clear()
        self.actor_ps.refresh_wt()
        if noise is None and dyn.rng == "device" and not getattr(self, "sync_loop", False):
            return self._run_async(init_obss, length, device_out)
        t_start = time.perf_counter()
        cur = torch.as_tensor(init_obss, dtype=torch.float32).to(self.dev).contiguous()
        outs = {k: [] for k in ("obss", "next_obss", "actions", "rewards", "terminals")}
        n_total, t = 0, 0
        for t in range(length):
            S = cur.shape[0]
            if self.uniform is not None:
                # combo.py:81-86: a ~ U(low, high), no actor pass
                obs_buf = cur.clone()
                if noise is not None:
                    act = torch.as_tensor(noise["actions"][t], dtype=torch.float32).to(self.dev).contiguous()
                else:
                    act = torch.empty(S, A, dtype=torch.float32, device=self.dev)
                    L.call("orlk_philox_fill", act.data_ptr(), 0, S * A, self.uniform[0], self.uniform[1], 0xac7,
                           self.philox_counter.data_ptr(), None, rt.cur)
                    L.call("orlk_step_end", dyn.engine.groups_ptr, 0, self.philox_counter.data_ptr(), rt.cur)
            else:
                run, plan, obs_buf = self._actor_plan(S)
                obs_buf.copy_(cur)
                plan.run_eager()
                if noise is not None:
                    eps = torch.as_tensor(noise["eps"][t], dtype=torch.float32).to(self.dev).contiguous()
                else:
                    eps = torch.empty(S, A, dtype=torch.float32, device=self.dev)
                    L.call("orlk_philox_fill", eps.data_ptr(), S * A, 0, 0.0, 1.0, 0xac7, self.philox_counter.data_ptr(), None,
                           rt.cur)
                    L.call("orlk_step_end", dyn.engine.groups_ptr, 0, self.philox_counter.data_ptr(), rt.cur)
                act = torch.empty(S, A, dtype=torch.float32, device=self.dev)
                L.call("orlk_tanh_gauss_sample", run.out.data_ptr(), 2 * A, 0, 1, eps.data_ptr(), S, A, act.data_ptr(), A, None,
                       None, 0, 0, None, 0, rt.cur)
            n64 = midx = None
            if noise is not None:
                n64 = torch.as_tensor(noise["normal"][t], dtype=torch.float64).to(self.dev).contiguous()
                midx = torch.as_tensor(noise["midx"][t], dtype=torch.int32).to(self.dev).contiguous()
            nobs, rew, term, raw, pen = dyn.step_device(obs_buf, act, n64, midx)
            for k, v in zip(outs, (obs_buf.clone(), nobs, act, rew, term)):
                outs[k].append(v)
            n_total += S
            nxt = torch.empty_like(nobs)
            L.call("orlk_compact_rows", term.data_ptr(), S, nobs.data_ptr(), O, O, nxt.data_ptr(), O, self.count_dev.data_ptr(),
                   rt.cur)
            rt.sync()
            alive = int(self.count_dev.item())
            if alive == 0:
                break
            cur = nxt[:alive]
        t_loop = time.perf_counter()
        if device_out:
            res = {k: torch.cat(v, 0) for k, v in outs.items()}
            self.last_timing = {"loop_ms": 1e3 * (t_loop - t_start), "export_ms": 0.0}
            return res, {"num_transitions": n_total, "reward_mean": float(res["rewards"].double().mean().item())}
        res = {k: torch.cat(v, 0).cpu().numpy() for k, v in outs.items()}
        res["terminals"] = res["terminals"].astype(bool)
        # where the wall time of the last call went (the loop ends with a device sync every step)
        self.last_timing = {"loop_ms": 1e3 * (t_loop - t_start), "export_ms": 1e3 * (time.perf_counter() - t_loop)}
        return res, {"num_transitions": n_total, "reward_mean": float(res["rewards"].astype(np.float64).mean())}
