#!/usr/bin/env python
"""Headline benchmark: CQL gradient steps/s, halfcheetah-shaped (obs 17, act 6), batch 256, 10 repeat actions,
hidden 256x3 (BASELINE.json `metric`, `configs[1]`).

    python bench.py --gpus N --steps K --warmup W            # the CUDA engine (this repo)
    python bench.py --impl reference --steps K --warmup W    # the unmodified reference (baseline/_ref) on the host CPU

One "step" = ReplayBuffer.sample(256) + CQLPolicy.learn(batch) with everything the reference's learn does
(actor, alpha, both critics with the conservative term, polyak, loss dict).  N > 1 runs N independent seeds, one
process per GPU, no communication on the data path ("replicas only"); the rates add up.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

import csv
import glob
import re

METRIC = "CQL gradient steps/s (hc-shaped, bs256)"
O_DIM, A_DIM, HIDDEN, BATCH, N_REPEAT, N_DATA = 17, 6, [256, 256, 256], 256, 10, 1_000_000
HYPER = dict(actor_lr=1e-4, critic_lr=3e-4, tau=0.005, gamma=0.99, cql_weight=5.0, temperature=1.0, max_q_backup=False,
             deterministic_backup=True, with_lagrange=False, lagrange_threshold=10.0, cql_alpha_lr=3e-4,
             num_repeat_actions=N_REPEAT)      # run_example/run_cql.py:26-56 defaults
ALPHA_LR = 1e-4
# algorithmic FLOP of one CQL step, SURVEY.md section 8(d): 13.84 GFLOP for the halfcheetah shape
FLOP_PER_STEP = 13.84e9
CONFIG = {"workload": "cql_halfcheetah_shaped obs17 act6 hidden256x3 batch256 repeat10 auto-alpha no-lagrange "
                      "buffer1M (configs[1])",
          "batch": BATCH, "buffer_rows": N_DATA,
          "precision_mode": "tf32x3 = fp32-grade: 3 TF32 tensor-core MMAs per product on hi/lo split operands, fp32 "
                            "accumulate (parity-tested at 1e-4); narrow layers in fp32 FFMA",
          "l2": "replay table (176 MB, random rows) is larger than L2; the 6 MB of model state is L2-resident by "
                "design of the workload; no explicit flush",
          "parallelism": "seed-parallel replicas, one per GPU, no data-path collective"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def ncu_tc_summary(stem: str = "ncu_tc_gemm"):
    """(csv path, mean DRAM bytes per launch, mean tensor-pipe-active %) of the launches of one tensor-core kernel, parsed at
    run time from the newest committed `profiles/<stem>_r*.csv` (one `ncu --set full` capture of this step)."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", stem + "_r*.csv")),
                   key=lambda f: int(re.search(r"_r(\d+)", os.path.basename(f)).group(1)))
    if not files:
        return None, None, None
    path = files[-1]
    rows = [r for r in csv.reader(ln for ln in open(path) if not ln.startswith("#"))]
    head, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(head)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, pipe = [], []
    for r in rows[2:]:
        if len(r) < len(head) or re.search(r"K=\d|head|wgrad0|fwd0", r[0]):
            continue            # the narrow first layer (one k-slab) is not one of the six 2.08 GFLOP launches
        rd = float(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]]
        tot.append(rd + wr)
        pipe.append(float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]))
    if not tot:
        return path, None, None
    return os.path.relpath(path, ROOT), float(np.mean(tot)), float(np.mean(pipe))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, period_ms: int = 10):
        self.index, self.proc, self.lines, self.period_ms = index, None, [], period_ms
        self.window = None          # (t0, t1) of the timed region, time.time() clock

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, inside = [], None, set(), 0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            if self.window is not None and self.window[0] <= ts <= self.window[1]:
                inside += 1
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        # the sampler runs from before the warm-up steps to after the timed loops: every sample is taken while this
        # process keeps the GPU busy with the same step; `samples_in_timed_region` counts those inside the two timed loops
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "samples_in_timed_region": inside, "period_ms": self.period_ms, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ builders
class _Box:
    def __init__(self, low, high, shape):
        self.low, self.high, self.shape = np.full(shape, low, np.float32), np.full(shape, high, np.float32), shape


def build_engine(device: str, seed: int, n_data: int):
    """Policy + buffer exactly as run_example/run_cql.py:72-139 builds them, on the synthetic halfcheetah-shaped data."""
    import random
    from offlinerlkit_b200.nets import MLP
    from offlinerlkit_b200.modules import ActorProb, Critic, TanhDiagGaussian
    from offlinerlkit_b200.policy import CQLPolicy
    from offlinerlkit_b200.buffer import ReplayBuffer
    from offlinerlkit_b200.synthetic import make_dataset
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    actor_backbone = MLP(input_dim=O_DIM, hidden_dims=HIDDEN)
    c1b, c2b = MLP(input_dim=O_DIM + A_DIM, hidden_dims=HIDDEN), MLP(input_dim=O_DIM + A_DIM, hidden_dims=HIDDEN)
    dist = TanhDiagGaussian(latent_dim=actor_backbone.output_dim, output_dim=A_DIM, unbounded=True, conditioned_sigma=True)
    actor, critic1, critic2 = ActorProb(actor_backbone, dist, device), Critic(c1b, device), Critic(c2b, device)
    log_alpha = torch.zeros(1, requires_grad=True, device=device)
    alpha = (-A_DIM, log_alpha, torch.optim.Adam([log_alpha], lr=ALPHA_LR))
    policy = CQLPolicy(actor, critic1, critic2,
                       torch.optim.Adam(actor.parameters(), lr=HYPER["actor_lr"]),
                       torch.optim.Adam(critic1.parameters(), lr=HYPER["critic_lr"]),
                       torch.optim.Adam(critic2.parameters(), lr=HYPER["critic_lr"]),
                       action_space=_Box(-1, 1, (A_DIM,)), tau=HYPER["tau"], gamma=HYPER["gamma"], alpha=alpha,
                       cql_weight=HYPER["cql_weight"], temperature=HYPER["temperature"], max_q_backup=False,
                       deterministic_backup=True, with_lagrange=HYPER["with_lagrange"],
                       lagrange_threshold=HYPER["lagrange_threshold"], cql_alpha_lr=HYPER["cql_alpha_lr"],
                       num_repeart_actions=N_REPEAT)
    policy.train()
    buf = ReplayBuffer(buffer_size=n_data, obs_shape=(O_DIM,), obs_dtype=np.float32, action_dim=A_DIM,
                       action_dtype=np.float32, device=device)
    buf.load_dataset(make_dataset(n_data, O_DIM, A_DIM, seed=0))
    return policy, buf


def build_oracle(device: str, seed: int, n_data: int):
    """The same workload for the CPU oracle port (oracle/algos.py:CQLOracle): identical shapes, init scale and data."""
    from oracle import algos
    from tests.helpers import actorprob_shapes, critic_shapes, recipe_state
    from offlinerlkit_b200.synthetic import make_dataset
    st = {}
    st.update(recipe_state(actorprob_shapes(O_DIM, A_DIM, HIDDEN), 100 + seed, "actor"))
    for i, c in enumerate(("critic1", "critic2")):
        cs = recipe_state(critic_shapes(O_DIM + A_DIM, HIDDEN), 101 + i + seed, c)
        st.update(cs)
        st.update({k.replace(c + ".", c + "_old.", 1): v.clone() for k, v in cs.items()})
    algos._Learner.device = device
    try:
        ora = algos.CQLOracle(st, alpha=(-A_DIM, 0.0, ALPHA_LR), **HYPER)
    finally:
        algos._Learner.device = "cpu"
    data = make_dataset(n_data, O_DIM, A_DIM, seed=0)
    data["rewards"] = data["rewards"].reshape(-1, 1)
    data["terminals"] = data["terminals"].reshape(-1, 1)
    return ora, data


def oracle_step_fn(ora, data, device: str):
    """One reference-style step: host index draw + NumPy fancy-index gather + 5 H2D copies (buffer.py:96-106),
    noise drawn as the reference does (device generator for eps, CPU generator for the uniform actions)."""
    from oracle import replay
    n, R = len(data["observations"]), BATCH * N_REPEAT

    def step():
        idx = replay.draw_indices(n, BATCH)
        batch = {k: torch.tensor(v).to(device) for k, v in replay.gather(data, idx).items()}
        noise = {"eps_actor": torch.randn(BATCH, A_DIM, device=device), "eps_next": torch.randn(BATCH, A_DIM, device=device),
                 "rand_act": torch.FloatTensor(R, A_DIM).uniform_(-1.0, 1.0).to(device),
                 "eps_pi": torch.randn(R, A_DIM, device=device), "eps_pi_next": torch.randn(R, A_DIM, device=device)}
        return ora.step(batch, noise)
    return step


def time_oracle(device: str, steps: int, warmup: int, budget_s: float, n_data: int):
    ora, data = build_oracle(device, 0, n_data)
    step = oracle_step_fn(ora, data, device)
    np.random.seed(0)
    torch.manual_seed(0)
    for _ in range(warmup):
        step()
    if device != "cpu":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    done = 0
    while done < steps:
        step()
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    if device != "cpu":
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return done / dt, done, dt


# ------------------------------------------------------------------------------------------------ reference arm
REF_BUILD = dict(obs_dim=O_DIM, act_dim=A_DIM, hidden=HIDDEN, hyper=HYPER, alpha_lr=ALPHA_LR)


def time_reference(device: str, steps: int, warmup: int, budget_s: float, n_data: int):
    """steps/s of the reference's CQL step on `device`: the UNMODIFIED reference from baseline/_ref (stock
    ReplayBuffer.sample + CQLPolicy.learn, kind "reference") when it is installed, else the oracle port of the same
    op stream (kind "port").  Returns (rate, steps done, seconds, kind)."""
    from baseline import reference_runner as rr
    ok, why = rr.available()
    if ok:
        rate, done, dt, _ = rr.time_cql(device, steps, warmup, budget_s, n_data, BATCH, **REF_BUILD)
        return rate, done, dt, "reference"
    print(f"[bench] reference not importable ({why}); timing the oracle port instead", file=sys.stderr, flush=True)
    rate, done, dt = time_oracle(device, steps, warmup, budget_s, n_data)
    return rate, done, dt, "port"


def run_reference(args, rank: int):
    """The reference's own CQL step on the box's host cores, all threads: `baseline/_ref` (the unmodified reference,
    installed by the recipe in baseline/reference_runner.py) through its public API; rank 0 only."""
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1; the reference arm is entitled to every host core
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    threads = torch.get_num_threads()
    w = min(args.warmup, 5)
    rate, done, dt, kind = time_reference("cpu", args.steps, w, budget_s=150.0, n_data=args.rows)
    sample = (f"{done} consecutive gradient steps (buffer.sample + policy.learn) of the same workload on the same "
              f"{args.rows}-row buffer (requested {args.steps}; capped at 150 s of CPU time), {w} warm-up steps")
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "steps/s", "n_gpus": args.gpus, "steps": done,
            "warmup": w, "ms_per_step": 1e3 / rate, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": CONFIG,
            "cpu_baseline": {"value": rate, "unit": "steps/s", "cores": threads, "kind": kind, "sample": sample,
                             "host_cpus": os.cpu_count()},
            "e2e": {"value": rate, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ engine arm
def per_launch_breakdown(eng, reps: int = 20):
    """Device time of every launch of the step: each launch is captured alone into a CUDA graph, repeated `reps`
    times back to back, and that graph is timed with a CUDA-event pair on the launching stream (warm caches, no
    host launch cost in the measurement)."""
    from offlinerlkit_b200 import _lib as L
    import ctypes as C
    plan = eng.plans["step"]
    rt = eng.rt
    e0, e1 = C.c_void_p(), C.c_void_p()
    L.call("orlk_event_create", C.byref(e0))
    L.call("orlk_event_create", C.byref(e1))
    out = []
    for label, op in plan.flat_ops:
        g = C.c_void_p()
        torch.cuda.synchronize()
        rt.cur = C.c_void_p(rt.capture_stream.cuda_stream)
        try:
            L.call("orlk_graph_begin", rt.cur)
            try:
                for _ in range(reps):
                    op()
            finally:
                L.call("orlk_graph_end", rt.cur, C.byref(g))
        finally:
            rt.cur = rt.exec_ptr
        L.call("orlk_graph_launch", g, rt.cur)
        L.call("orlk_event_record", e0, rt.cur)
        L.call("orlk_graph_launch", g, rt.cur)
        L.call("orlk_event_record", e1, rt.cur)
        ms = C.c_float()
        L.call("orlk_event_elapsed_ms", e0, e1, C.byref(ms))
        L.call("orlk_graph_destroy", g)
        out.append((label, 1e3 * ms.value / reps))
    L.call("orlk_event_destroy", e0)
    L.call("orlk_event_destroy", e1)
    return out      # microseconds per launch


def run_engine(args, rank: int, world: int, local_rank: int):
    import ctypes as C
    from offlinerlkit_b200 import _lib as L
    device = f"cuda:{local_rank}"
    torch.cuda.set_device(local_rank)
    from offlinerlkit_b200 import parallel
    dist_on = parallel.init("nccl", torch.device(device))
    if dist_on:
        import torch.distributed as dist
    policy, buf = build_engine(device, seed=parallel.seed_for_rank(0, rank), n_data=args.rows)
    K, W = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    sampler.start()                 # before the warm-up: short timed regions (--steps 20 = 6 ms) still get samples

    # ---- warm-up through the public API (uploads the table, builds and captures the step graph)
    for _ in range(W):
        loss = policy.learn(buf.sample(BATCH))
    eng = policy._engine
    rt = eng.rt
    n_kernels = sum(1 for lbl, _ in eng.plans["step"].flat_ops if lbl != "losses_d2h") + 1      # + the gather kernel

    def barrier():
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize()

    def ev():
        e = C.c_void_p()
        L.call("orlk_event_create", C.byref(e))
        return e

    e0, e1 = ev(), ev()

    # ---- (1) device-resident: indices for all K steps already in HBM, no host sync inside the timed region
    idx_all = torch.from_numpy(np.random.randint(0, buf._size, size=(K, BATCH))).to(device)
    barrier()
    t_win0 = time.time()
    L.call("orlk_event_record", e0, rt.cur)
    for t in range(K):
        buf.gather_device(idx_all[t])
        eng.enqueue("step")
    L.call("orlk_event_record", e1, rt.cur)
    barrier()
    ms = C.c_float()
    L.call("orlk_event_elapsed_ms", e0, e1, C.byref(ms))
    dev_ms = ms.value

    # ---- (2) end to end through the public API: host index draw, pinned H2D of the indices, D2H of the loss block
    barrier()
    t0 = time.perf_counter()
    L.call("orlk_event_record", e0, rt.cur)
    for t in range(K):
        loss = policy.learn(buf.sample(BATCH))
    L.call("orlk_event_record", e1, rt.cur)
    barrier()
    e2e_wall = time.perf_counter() - t0
    L.call("orlk_event_elapsed_ms", e0, e1, C.byref(ms))
    e2e_ms = max(ms.value, 1e3 * e2e_wall)
    # ---- (3) informational: the K-step call (policy.learn_many: one host synchronisation for all K steps)
    barrier()
    t0 = time.perf_counter()
    many = policy.learn_many(buf, K, BATCH)
    torch.cuda.synchronize()
    many_ms = 1e3 * (time.perf_counter() - t0)
    sampler.window = (t_win0, time.time())
    # keep the same step running until the sampler has seen >= 0.25 s of it (not timed; clocks only)
    t_probe = time.perf_counter()
    while time.perf_counter() - t_probe < 0.25:
        for _ in range(50):
            eng.enqueue("step")
        torch.cuda.synchronize()
    clocks = sampler.stop()

    dev_ms, e2e_ms, many_ms = parallel.reduce_scalars([dev_ms, e2e_ms, many_ms], "max", device)      # the slowest rank bounds the job
    value = world * K / (dev_ms * 1e-3)
    e2e = world * K / (e2e_ms * 1e-3)

    if rank == 0:
        peaks = load_peaks()
        # ---- roofline of the dominant kernel: per-launch device time from CUDA events (eager replay of the same list)
        br = per_launch_breakdown(eng)
        Mc = BATCH + 3 * BATCH * N_REPEAT
        D0 = 17 + 6                     # obs + act columns of the first critic layer
        gemm = 2 * 2 * Mc * 256 * 256   # one hidden layer of both critics over the critic batch: 2.08 GFLOP
        flop = {}
        for l in (1, 2):
            for kind in ("fwd", "dgrad", "wgrad"):
                flop[f"C.critic.{kind}{l}"] = flop[f"C.critic.{kind}{l}.tc"] = gemm
        flop["C.critic.wgrad_big"] = 2 * gemm
        # fused passes: first layer + two hidden layers + scalar head of the online critics (+ the target critics' 256 rows)
        fwd_flop = gemm * 2 + 2 * 2 * Mc * D0 * 256 + 2 * 2 * Mc * 256
        fwd_flop += (gemm * 2 + 2 * 2 * Mc * D0 * 256 + 2 * 2 * Mc * 256) * BATCH // Mc
        flop["C.critic+C.target.fwd_fused.tc"] = fwd_flop
        flop["C.critic.fwd_fused.tc"] = gemm * 2 + 2 * 2 * Mc * D0 * 256 + 2 * 2 * Mc * 256
        flop["C.critic.dgrad_fused.tc"] = gemm * 2
        big = [(lbl, us) for lbl, us in br if lbl in flop]
        on_tc = any(lbl.endswith(".tc") for lbl, _ in big)
        big_us = sum(us for _, us in big)
        big_flop = sum(flop[lbl] for lbl, _ in big)
        step_us = sum(us for _, us in br)
        peak = peaks["bf16_tflops_sustained"]
        fused = [(lbl, us) for lbl, us in big if "fwd_fused" in lbl]
        if fused:
            # the dominant kernel of the step: the fused forward pass of the critics (one launch)
            dom_lbl, dom_us = fused[0]
            achieved = flop[dom_lbl] / (dom_us * 1e-6) / 1e12
            kname = ("k_critic_fwd (tcgen05.mma kind::tf32, 3 MMA passes per product; first layer + two hidden layers + scalar "
                     "head of both critics over 7936 rows and of both target critics over 256 rows in one launch: TMEM "
                     "ping-pong accumulators, the epilogue writes the next layer's operand tiles, weights by TMA)")
            ncu_csv, ncu_traffic, ncu_pipe = ncu_tc_summary("ncu_critic_fwd")
            # X in, three activation tensors + decision bits + q out, weights (+ lo words) once per member
            alg_bytes = Mc * 24 * 4 + 3 * 2 * Mc * 256 * 4 + 3 * 2 * Mc * 32 + 2 * Mc * 4 + 2 * 2 * 2 * (256 * 32 + 2 * 65536) * 4
            n_dom, dom_total_us = 1, dom_us
        else:
            achieved = big_flop / (big_us * 1e-6) / 1e12
            kname = (f"k_tc_gemm<{eng.tc_passes}> (tcgen05.mma kind::tf32, {eng.tc_passes} MMA pass(es) per product, TMEM "
                     "accumulators, TMA operand ring; hidden-layer fwd/dgrad/wgrad of both critics over 7936 rows)") if on_tc \
                else "k_gemm_grouped<128,128,16,8,8> (fp32 FFMA; hidden-layer fwd/dgrad/wgrad of both critics over 7936 rows)"
            ncu_csv, ncu_traffic, ncu_pipe = ncu_tc_summary()
            alg_bytes = 2 * Mc * 256 * 4 * 2 + 2 * 256 * 256 * 4
            n_dom, dom_total_us = len(big), big_us
        roof = {"bound": "tensor", "kernel": kname,
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_source": f"{peaks['source']} bf16 dense (sustained)",
                # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of this
                # step (parsed at run time; ncu flushes the caches before each pass, in the running step the operands are L2 hits)
                "traffic": ncu_traffic if on_tc else None,
                "traffic_source": (f"{ncu_csv}, parsed at run time (ncu --set full, B200, same step)") if on_tc else None,
                "algorithmic_bytes_per_launch": alg_bytes,
                # each product is three TF32 MMAs: what the tensor pipe really executes, against the TF32 (= bf16 / 2) peak
                "tf32_mma_frac_of_tf32_peak": (3 if eng.tc_passes == 3 else 1) * achieved / (peak / 2),
                "ncu_tensor_pipe_active_pct": ncu_pipe if on_tc else None,
                "launches_per_step": n_dom, "us_per_step": dom_total_us, "share_of_step": dom_total_us / step_us,
                # all tensor-core launches of the critic update together (fused forward, fused dgrad chain, weight gradients)
                "all_tensor_core_launches": {"launches": len(big), "us_per_step": big_us, "tflops": big_flop / (big_us * 1e-6) / 1e12,
                                             "frac": big_flop / (big_us * 1e-6) / 1e12 / peak, "share_of_step": big_us / step_us},
                "step_frac": FLOP_PER_STEP * (value / world) / 1e12 / peak,
                "fp32_simt_peak_tflops": 148 * 128 * 2 * 1.965e9 / 1e12,
                "frac_of_fp32_simt_peak": achieved / (148 * 128 * 2 * 1.965e9 / 1e12)}
        top = sorted(br, key=lambda x: -x[1])[:8]
        if args.full_breakdown:
            # the captured step graph replayed back to back with nothing in between (no gather, no host work)
            ev0, ev1, rms = C.c_void_p(), C.c_void_p(), C.c_float()
            L.call("orlk_event_create", C.byref(ev0))
            L.call("orlk_event_create", C.byref(ev1))
            plan = eng.plans["step"]
            for _ in range(5):
                plan.launch()
            L.call("orlk_event_record", ev0, rt.cur)
            for _ in range(100):
                plan.launch()
            L.call("orlk_event_record", ev1, rt.cur)
            L.call("orlk_event_elapsed_ms", ev0, ev1, C.byref(rms))
            with open(args.full_breakdown, "w") as f:
                json.dump({"precision": eng.precision, "launch_us": br, "graph_step_us": 1e3 * dev_ms / K,
                           "graph_replay_only_us": 1e3 * rms.value / 100}, f, indent=1)
        line = {"metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": dev_ms / K, "us_per_update": 1e3 * dev_ms / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": CONFIG, "clocks": clocks,
                "e2e": {"value": e2e, "unit": "steps/s", "h2d_bytes_per_step": 8 * BATCH, "d2h_bytes_per_step": 4 * 32,
                        "ms_per_step": e2e_ms / K},
                "learn_many": {"value": world * K / (many_ms * 1e-3), "unit": "steps/s", "ms_per_step": many_ms / K,
                               "note": "policy.learn_many(buffer, K, 256): the same K steps (same index stream, same results) "
                                       "behind one host synchronisation; informational, not the e2e headline"},
                "gpu_launches": n_kernels * K, "launches_per_step": n_kernels, "roofline": roof, "precision": eng.precision,
                "launch_breakdown_us": {lbl: round(us, 2) for lbl, us in top}, "eager_step_us": step_us,
                "last_loss": {k: float(v) for k, v in loss.items()}}
        if world == 1 and not args.no_cpu_baseline:
            torch.set_num_threads(max(1, os.cpu_count() or 1))
            threads = torch.get_num_threads()
            # the reference's own step (baseline/_ref, unmodified) on the host cores, same 1 M-row buffer: a bounded sample
            rate, done, dt, kind = time_reference("cpu", 400, 3, budget_s=20.0, n_data=args.rows)
            cb = {"value": rate, "unit": "steps/s", "cores": threads, "kind": kind, "host_cpus": os.cpu_count(),
                  "sample": f"{done} gradient steps (buffer.sample + policy.learn) of the same workload on the same "
                            f"{args.rows}-row buffer in {dt:.1f} s, 3 warm-up steps"}
            try:
                # the north_star's denominator: the same reference code with device="cuda" on this B200 (BASELINE.md
                # section 3, row R-GPU): 200 warm-up steps, >= 1000 timed, synchronised on both sides
                grate, gdone, gdt, gkind = time_reference(device, 1500, 200, budget_s=40.0, n_data=args.rows)
                cb.update({"gpu_eager_steps_s": grate, "gpu_eager_kind": gkind, "gpu_eager_steps": gdone,
                           "speedup_vs_gpu_eager_device": value / grate, "speedup_vs_gpu_eager_e2e": e2e / grate,
                           "target_50x_met": bool(e2e / grate >= 50.0),
                           "gpu_eager_note": "stock CQLPolicy.learn + ReplayBuffer.sample of the reference with "
                                             "device='cuda' on the same B200; the north_star's 50x target is against "
                                             "this rate"})
            except Exception as ex:      # never let the informational leg break the bench line
                cb["gpu_eager_error"] = repr(ex)
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if dist_on:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ member-sharded EDAC
def run_edac_sharded(args, rank: int, world: int, local_rank: int):
    """BASELINE.json configs[2]: EDAC, halfcheetah-shaped, 10 critics, eta = 1, the critics sharded over the ranks (3/3/2/2 on
    four GPUs), three NCCL all-gathers per step (engine/edac_sharded.py).  ONE training run on N GPUs: `value` is that
    run's gradient steps/s (scaling: "strong")."""
    import ctypes as C
    import random
    from offlinerlkit_b200 import _lib as L
    from offlinerlkit_b200 import parallel
    from offlinerlkit_b200.nets import MLP
    from offlinerlkit_b200.modules import ActorProb, EnsembleCritic, TanhDiagGaussian
    from offlinerlkit_b200.policy import EDACPolicy
    from offlinerlkit_b200.buffer import ReplayBuffer
    from offlinerlkit_b200.synthetic import make_dataset
    device = f"cuda:{local_rank}"
    torch.cuda.set_device(local_rank)
    dist_on = parallel.init("nccl", torch.device(device))
    E = 10
    random.seed(0), np.random.seed(0), torch.manual_seed(0), torch.cuda.manual_seed_all(0)      # replicated on every rank
    ab = MLP(input_dim=O_DIM, hidden_dims=HIDDEN)
    actor = ActorProb(ab, TanhDiagGaussian(ab.output_dim, A_DIM, unbounded=True, conditioned_sigma=True), device)
    critics = EnsembleCritic(O_DIM, A_DIM, HIDDEN, num_ensemble=E, device=device)
    log_alpha = torch.zeros(1, requires_grad=True, device=device)
    policy = EDACPolicy(actor, critics, torch.optim.Adam(actor.parameters(), lr=1e-4),
                        torch.optim.Adam(critics.parameters(), lr=3e-4), tau=0.005, gamma=0.99,
                        alpha=(-A_DIM, log_alpha, torch.optim.Adam([log_alpha], lr=1e-4)), max_q_backup=False,
                        deterministic_backup=False, eta=1.0)       # run_example/run_edac.py:26-58
    policy.train()
    if world > 1:
        from offlinerlkit_b200.engine.edac_sharded import NcclComm
        policy.shard_critics(rank, world, NcclComm())
    buf = ReplayBuffer(buffer_size=args.rows, obs_shape=(O_DIM,), obs_dtype=np.float32, action_dim=A_DIM,
                       action_dtype=np.float32, device=device)
    buf.load_dataset(make_dataset(args.rows, O_DIM, A_DIM, seed=0))
    K, W = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(W):
        loss = policy.learn(buf.sample(BATCH))
    eng = policy._engine

    def barrier():
        if dist_on:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    e0, e1 = C.c_void_p(), C.c_void_p()
    L.call("orlk_event_create", C.byref(e0))
    L.call("orlk_event_create", C.byref(e1))
    barrier()
    t0w = time.time()
    t0 = time.perf_counter()
    L.call("orlk_event_record", e0, eng.rt.cur)
    for _ in range(K):
        loss = policy.learn(buf.sample(BATCH))
    L.call("orlk_event_record", e1, eng.rt.cur)
    barrier()
    wall = time.perf_counter() - t0
    ms = C.c_float()
    L.call("orlk_event_elapsed_ms", e0, e1, C.byref(ms))
    sampler.window = (t0w, time.time())
    clocks = sampler.stop()
    (t_ms,) = parallel.reduce_scalars([max(ms.value, 1e3 * wall)], "max", device)
    if rank == 0:
        n_launch = sum(p.n_launches for p in eng.plans.values())
        counts = getattr(eng, "counts", [E])
        line = {"metric": "EDAC gradient steps/s (hc-shaped, bs256, 10 critics, members sharded)", "value": K / (t_ms * 1e-3),
                "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_ms / K, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
                "config": {"workload": "edac_halfcheetah_shaped obs17 act6 hidden256x3 batch256 E10 eta1 buffer1M (configs[2])",
                           "members_per_rank": counts, "collectives_per_step": 0 if world == 1 else 3,
                           "parallelism": "critic members sharded over the ranks; actor / alpha / batch / noise replicated"},
                "clocks": clocks,
                "e2e": {"value": K / (t_ms * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": 8 * BATCH, "d2h_bytes_per_step": 128},
                "gpu_launches": n_launch * K, "launches_per_step": n_launch, "precision": eng.precision,
                "last_loss": {k: float(v) for k, v in loss.items()}}
        print(json.dumps(line), flush=True)
    if dist_on:
        import torch.distributed as dist
        # a captured graph that holds NCCL kernels must die before the communicator: destroy_process_group() hung for
        # minutes with the step graph alive (round 2, 4 x B200).  Drop the graph, synchronise, and leave without the
        # communicator's own teardown.
        if hasattr(eng, "_whole"):
            eng._whole = None
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


# ------------------------------------------------------------------------------------------------ member-sharded dynamics
def run_dyn_sharded(args, rank: int, world: int, local_rank: int):
    """BASELINE.json configs[4], training half: EnsembleDynamicsModel (7 members, hidden 200 x 4, halfcheetah-shaped), one
    `learn` pass of --steps mini-batches of 256 rows per member, the members sharded over the ranks (one all-gather of the
    shared log-variance bounds' partial gradients per mini-batch).  ONE training run on N GPUs (scaling: "strong")."""
    import ctypes as C
    from offlinerlkit_b200 import _lib as L
    from offlinerlkit_b200 import parallel
    from offlinerlkit_b200.modules import EnsembleDynamicsModel
    from offlinerlkit_b200.dynamics import EnsembleDynamics
    from offlinerlkit_b200.utils.scaler import StandardScaler
    from offlinerlkit_b200.utils.termination_fns import termination_fn_halfcheetah
    from offlinerlkit_b200.synthetic import make_dataset
    device = f"cuda:{local_rank}"
    torch.cuda.set_device(local_rank)
    dist_on = parallel.init("nccl", torch.device(device))
    E, Bm = 7, 256
    np.random.seed(0), torch.manual_seed(0), torch.cuda.manual_seed_all(0)
    model = EnsembleDynamicsModel(O_DIM, A_DIM, [200, 200, 200, 200], num_ensemble=E, num_elites=5,
                                  weight_decays=[2.5e-5, 5e-5, 7.5e-5, 7.5e-5, 1e-4], device=device)     # run_mopo.py:60-72
    dyn = EnsembleDynamics(model, torch.optim.Adam(model.parameters(), lr=1e-3), StandardScaler(), termination_fn_halfcheetah)
    if world > 1:
        from offlinerlkit_b200.engine.edac_sharded import NcclComm
        dyn.shard_members(rank, world, NcclComm())
    n = min(args.rows, 200_000)
    d = make_dataset(n, O_DIM, A_DIM, seed=0)
    x = np.concatenate([d["observations"], d["actions"]], 1)
    y = np.concatenate([d["next_observations"] - d["observations"], d["rewards"].reshape(-1, 1)], 1)
    dyn.scaler.fit(x)
    eng = dyn.engine
    src_x = torch.from_numpy(dyn.scaler.transform(x).astype(np.float32)).to(device)
    src_y = torch.from_numpy(y.astype(np.float32)).to(device)
    K, W = args.steps, max(args.warmup, 3)
    idx = torch.from_numpy(np.random.randint(0, n, size=(E, Bm * max(K, W)))).to(device)
    sampler = ClockSampler(local_rank)
    sampler.start()
    eng.learn(src_x, src_y, idx[:, :Bm * W].contiguous(), Bm, 0.01)

    def barrier():
        if dist_on:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    t0w, t0 = time.time(), time.perf_counter()
    loss = eng.learn(src_x, src_y, idx[:, :Bm * K].contiguous(), Bm, 0.01)
    barrier()
    wall = time.perf_counter() - t0
    sampler.window = (t0w, time.time())
    clocks = sampler.stop()
    (t_ms,) = parallel.reduce_scalars([1e3 * wall], "max", device)
    if rank == 0:
        line = {"metric": "dynamics mini-batches/s (7 members x 256 rows, hidden 200x4, members sharded)", "value": K / (t_ms * 1e-3),
                "unit": "mini-batches/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_ms / K,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
                "config": {"workload": "dynamics_training obs17 act6 E7 hidden200x4 batch256/member (configs[4])",
                           "members_per_rank": eng.counts, "collectives_per_step": 0 if world == 1 else 1,
                           "parallelism": "ensemble members sharded over the ranks; shared logvar bounds replicated"},
                "clocks": clocks, "e2e": {"value": K / (t_ms * 1e-3), "unit": "mini-batches/s", "h2d_bytes_per_step": 0,
                                          "d2h_bytes_per_step": 0},
                "gpu_launches": K * sum(p.n_launches for p, _ in eng._learn_plans.values()), "last_loss": float(loss)}
        print(json.dumps(line), flush=True)
    if dist_on:
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--rows", type=int, default=N_DATA)
    ap.add_argument("--workload", default="cql", choices=["cql", "edac_sharded", "dyn_sharded"],
                    help="cql = the headline (seed-parallel replicas); edac_sharded = BASELINE.json configs[2], ONE run with "
                         "the 10 critics sharded over the --gpus ranks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default=None, choices=["fp32", "tf32x3", "tf32"],
                    help="GEMM mode of the wide layers (default tf32x3, the fp32-parity tensor-core mode)")
    ap.add_argument("--full-breakdown", default=None, help="write the per-launch device times (JSON) to this file")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.precision:
        os.environ["ORLK_PRECISION"] = args.precision
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback (use --impl reference for the CPU arm)")
    if args.workload == "edac_sharded":
        run_edac_sharded(args, rank, world, local_rank)
        return
    if args.workload == "dyn_sharded":
        run_dyn_sharded(args, rank, world, local_rank)
        return
    run_engine(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
