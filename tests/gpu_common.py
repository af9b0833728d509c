"""Helpers for the GPU parity tests: build facade policies from golden metadata and run golden steps."""
from typing import Dict

import numpy as np
import torch

from tests.helpers import (Golden, initial_state, assert_stats_close, assert_grad_stats_close, assert_grads_close, rel_err)


class Box:
    def __init__(self, low, high, shape):
        self.low = np.full(shape, low, dtype=np.float32)
        self.high = np.full(shape, high, dtype=np.float32)
        self.shape = shape


def build_policy(meta, device="cuda:0"):
    from offlinerlkit_b200.nets import MLP
    from offlinerlkit_b200.modules import ActorProb, Actor, Critic, EnsembleCritic, TanhDiagGaussian, DiagGaussian
    import offlinerlkit_b200.policy as P
    algo, O, A, hid, hy = meta["algo"], meta["O"], meta["A"], meta["hidden"], meta.get("hyper", {})
    adam = lambda m, lr: torch.optim.Adam(m.parameters(), lr=lr)

    def alpha_tuple():
        la = torch.zeros(1, requires_grad=True, device=device)
        return (meta["target_entropy"], la, torch.optim.Adam([la], lr=meta["alpha_lr"]))

    def tanh_actor():
        bb = MLP(O, hid)
        return ActorProb(bb, TanhDiagGaussian(bb.output_dim, A, unbounded=True, conditioned_sigma=True), device)

    if algo in ("cql", "sac", "combo"):
        actor, c1, c2 = tanh_actor(), Critic(MLP(O + A, hid), device), Critic(MLP(O + A, hid), device)
        opt = (adam(actor, hy["actor_lr"]), adam(c1, hy["critic_lr"]), adam(c2, hy["critic_lr"]))
        if algo == "sac":
            return P.SACPolicy(actor, c1, c2, *opt, tau=hy["tau"], gamma=hy["gamma"], alpha=alpha_tuple())
        cql_kw = dict(action_space=Box(-1, 1, (A,)), tau=hy["tau"], gamma=hy["gamma"],
                      alpha=alpha_tuple(), cql_weight=hy["cql_weight"], temperature=hy["temperature"],
                      max_q_backup=hy["max_q_backup"], deterministic_backup=hy["deterministic_backup"],
                      with_lagrange=hy["with_lagrange"], lagrange_threshold=hy["lagrange_threshold"],
                      cql_alpha_lr=hy["cql_alpha_lr"], num_repeart_actions=hy["num_repeat_actions"])
        if algo == "combo":
            return P.COMBOPolicy(meta.get("dynamics"), actor, c1, c2, *opt, uniform_rollout=False, rho_s=meta["rho_s"],
                                 **cql_kw)
        return P.CQLPolicy(actor, c1, c2, *opt, **cql_kw)
    if algo == "edac":
        actor = tanh_actor()
        critics = EnsembleCritic(O, A, hid, num_ensemble=meta["E"], device=device)
        return P.EDACPolicy(actor, critics, adam(actor, hy["actor_lr"]), adam(critics, hy["critic_lr"]), tau=hy["tau"],
                            gamma=hy["gamma"], alpha=alpha_tuple(), max_q_backup=hy.get("max_q_backup", False),
                            deterministic_backup=hy["deterministic_backup"], eta=hy["eta"])
    if algo == "iql":
        bb = MLP(O, hid, dropout_rate=None)
        actor = ActorProb(bb, DiagGaussian(bb.output_dim, A, unbounded=False, conditioned_sigma=False), device)
        q1, q2, v = Critic(MLP(O + A, hid), device), Critic(MLP(O + A, hid), device), Critic(MLP(O, hid), device)
        return P.IQLPolicy(actor, q1, q2, v, adam(actor, hy["actor_lr"]), adam(q1, hy["critic_q_lr"]),
                           adam(q2, hy["critic_q_lr"]), adam(v, hy["critic_v_lr"]), action_space=Box(-1, 1, (A,)),
                           tau=hy["tau"], gamma=hy["gamma"], expectile=hy["expectile"], temperature=hy["temperature"])
    if algo == "td3bc":
        actor = Actor(MLP(O, hid), A, device=device)
        c1, c2 = Critic(MLP(O + A, hid), device), Critic(MLP(O + A, hid), device)
        return P.TD3BCPolicy(actor, c1, c2, adam(actor, hy["actor_lr"]), adam(c1, hy["critic_lr"]),
                             adam(c2, hy["critic_lr"]), tau=hy["tau"], gamma=hy["gamma"], max_action=hy["max_action"],
                             policy_noise=hy["policy_noise"], noise_clip=hy["noise_clip"],
                             update_actor_freq=hy["update_actor_freq"], alpha=hy["alpha"])
    raise KeyError(algo)


class EngineGrads:
    """The gradients the engine's fused Adam launches consumed, recovered from the first moments:
    m_t = beta1 m_{t-1} + (1 - beta1) g_t  =>  g_t = (m_t - beta1 m_{t-1}) / (1 - beta1).  ``snapshot()`` before a step,
    ``after()`` behind it.  Names follow the policy's ``state_dict`` (+ ``log_alpha`` / ``cql_log_alpha``)."""

    def __init__(self, policy, eng, param_sets=None):
        self.eng = eng
        self.where = {}
        for name, p in policy.named_parameters():
            for ps in (param_sets or eng.param_sets):
                off = (p.data_ptr() - ps.P.data_ptr()) // 4
                if 0 <= off < ps.total and p.device == ps.P.device:
                    self.where[name] = (ps, off, p.shape, p.numel())
        self.beta1 = float(eng._groups[0].beta1)
        self.prev = None

    def _moments(self):
        out = {n: ps.Mo[off:off + cnt].clone() for n, (ps, off, shp, cnt) in self.where.items()}
        for n, attr in (("log_alpha", "alpha_mv"), ("cql_log_alpha", "cql_mv")):
            t = getattr(self.eng, attr, None)
            if t is not None:
                out[n] = t[:1].clone()
        return out

    def snapshot(self):
        self.prev = self._moments()

    def after(self, names):
        cur, b1 = self._moments(), self.beta1
        out = {}
        for n in names:
            shape = self.where[n][2] if n in self.where else (1,)
            out[n] = ((cur[n].double() - b1 * self.prev[n].double()) / (1.0 - b1)).view(shape).cpu()
        return out


def make_oracle(meta):
    """The CPU oracle (pinned to the reference by tests/test_oracle_golden.py) on the fixture's initial state: supplies
    the FULL gradient tensors of config-size runs, of which the fixture only stores fingerprints."""
    from oracle import algos
    algo, hy = meta["algo"], meta.get("hyper", {})
    st = initial_state(meta)
    alpha = (meta["target_entropy"], 0.0, meta["alpha_lr"]) if "alpha_lr" in meta else None
    if algo == "cql":
        return algos.CQLOracle(st, alpha=alpha, **hy)
    if algo == "combo":
        return algos.COMBOOracle(st, rho_s=meta["rho_s"], alpha=alpha, **hy)
    if algo == "sac":
        return algos.SACOracle(st, alpha=alpha, **hy)
    if algo == "edac":
        return algos.EDACOracle(st, alpha=alpha, **hy)
    if algo == "iql":
        return algos.IQLOracle(st, **hy)
    if algo == "td3bc":
        return algos.TD3BCOracle(st, **hy)
    raise KeyError(algo)


def sync_oracle_to_engine(ora, policy) -> None:
    """Put the oracle at the ENGINE's current parameters (and entropy / Lagrange multipliers).  From the second step on
    the two differ by Adam's sign-level noise (an element whose first gradient is a rounding-level cancellation moves by
    +lr in one and -lr in the other); the gradient check of step t is a statement about step t alone."""
    sd = policy.state_dict()
    with torch.no_grad():
        for k, v in ora.p.items():
            if k in sd and v.is_floating_point():
                v.copy_(sd[k].detach().cpu())
        la = getattr(policy, "_log_alpha", None)
        if la is not None and getattr(ora, "auto_alpha", False):
            ora.log_alpha.copy_(la.detach().cpu().reshape(ora.log_alpha.shape))
            a = ora.log_alpha.detach().exp()
            ora.alpha = torch.clamp(a, 0.0, 1.0) if ora._clamp01 else a
        cla = getattr(policy, "cql_log_alpha", None)
        if cla is not None and hasattr(ora, "cql_log_alpha"):
            ora.cql_log_alpha.copy_(cla.detach().cpu().reshape(ora.cql_log_alpha.shape))


def check_step_grads(g: Golden, t: int, tap: "EngineGrads", ora, ref_batch, noise, tol: float):
    """Engine gradients of step t vs the oracle's FULL tensors (the oracle is pinned to the reference's own gradients by
    tests/test_oracle_golden.py): per tensor relative L2 <= tol and max |diff| <= 2 tol max|g| (north_star: 1e-4), up to
    the ReLU decisions at pre-activations within rounding of zero (tests/kinks.py).  At step 0, where the engine and the
    reference start from identical parameters, also vs the reference's fingerprints stored in the fixture."""
    from tests.kinks import assert_grads_close_up_to_kinks
    stats = g.group(f"gradstats{t}")
    got = tap.after(stats.keys())
    run = (lambda o: o.step(ref_batch, noise)) if noise is not None else (lambda o: o.step(ref_batch))
    flips = assert_grads_close_up_to_kinks(got, ora, run, tol, what=f"{g.meta['algo']} step {t}", eng=tap.eng)
    if t == 0:
        # one flipped ReLU bit moves a cancelling 7936-row gradient sum by ~5e-4 (profiles/mask_flip_r02.txt)
        # (a flipped unit's own bias-gradient element changes by a whole row's contribution: with flips only the norms are
        # compared with the reference's raw fingerprints; the full tensors were compared above under the engine's bits)
        assert_grad_stats_close(got, stats, tol=tol if flips == 0 else 50 * tol, what=f"step {t}", elements=flips == 0)
    run(ora)            # advance the oracle (optimiser counters, TD3+BC's update counter)


def load_state(policy, state: Dict[str, torch.Tensor]) -> None:
    missing, unexpected = policy.load_state_dict(state, strict=False)
    assert not unexpected, unexpected
    assert all("saved_" in k for k in missing), missing


def make_buffer(g: Golden, device="cuda:0"):
    from offlinerlkit_b200.buffer import ReplayBuffer
    m = g.meta
    data = g.dataset()
    buf = ReplayBuffer(m["n_data"], (m["O"],), np.float32, m["A"], np.float32, device=device)
    buf.load_dataset(data)
    return buf, data


def run_golden_steps(g: Golden, n_steps=None, tol=1e-4, verbose=False, use_graph=True, device="cuda:0", precision=None,
                     elementwise=True, grads=True):
    """Engine vs golden (= the real reference): index draw + gather bit-exact, then losses and parameters."""
    m = g.meta
    policy = build_policy(m, device)
    load_state(policy, initial_state(m))
    policy.train()
    buf, data = make_buffer(g, device)
    np.random.seed(m["np_seed"])
    n_steps = n_steps or m["n_steps"]
    lr_atol = 2.5 * max(v for k, v in m["hyper"].items() if k.endswith("_lr"))
    tap, ora = None, (make_oracle(m) if grads else None)
    for t in range(n_steps):
        batch = buf.sample(m["B"])
        torch.cuda.synchronize()
        assert np.array_equal(batch.indices.cpu().numpy(), g["idx"][t]), "index stream differs from the reference"
        ref_b = g.batch(t, data)
        for k, v in ref_b.items():
            assert torch.equal(batch[k].cpu().reshape(v.shape), v), f"gather not bit-exact: {k}"
        eng = policy.engine(m["B"])
        eng.use_graph = use_graph
        if precision is not None and t == 0:
            eng.precision = precision
        noise = g.noise(t) if any(k.startswith(f"noise{t}|") for k in g.z.files) else None
        if grads and tap is None and eng.param_sets:
            tap = EngineGrads(policy, eng)
        if tap is not None:
            tap.snapshot()
            if t > 0:
                sync_oracle_to_engine(ora, policy)
        out = policy.learn(batch, noise=noise) if noise is not None else policy.learn(batch)
        if tap is not None:
            check_step_grads(g, t, tap, ora, ref_b, noise, tol)
        ref = g.losses(t)
        if verbose:
            print(f"step {t}: engine {out}\n        golden {ref}", flush=True)
        assert out.keys() == ref.keys(), (out.keys(), ref.keys())
        for k in ref:
            assert abs(out[k] - ref[k]) <= tol * max(1.0, abs(ref[k])), (t, k, out[k], ref[k])
        sd = {k: v.detach().cpu() for k, v in policy.state_dict().items()}
        # an element whose gradient is a rounding-level cancellation can take Adam's +-lr step the other way in every
        # step: the element-wise slack grows by 2 lr per step (2.5 lr after the first)
        step_atol = lr_atol * (1 + 0.8 * t)
        assert_stats_close(sd, g.group(f"stats{t}"), tol=tol, lr_atol=step_atol if elementwise else 10 * lr_atol)
    post = g.group("post")
    if post and n_steps == m["n_steps"]:
        sd = {k: v.detach().cpu() for k, v in policy.state_dict().items()}
        for k, v in post.items():
            if v.dtype.kind == "f" and "saved_" not in k:
                err = np.abs(sd[k].numpy() - v).max()
                assert err <= tol * np.abs(v).max() + lr_atol, (k, err)
    return policy


def run_combo_golden_steps(g: Golden, tol=1e-4, verbose=False, device="cuda:0", precision=None, mode="checked"):
    """COMBOPolicy.learn on {"real", "fake"} batches sampled from two facade buffers (as MBPolicyTrainer does) vs the
    golden run of the real reference: both index streams and gathers bit-exact, then losses and parameters.
    mode: "checked" reads the sampled rows first (the draws are materialised, the engine re-gathers them eagerly);
    "lazy" hands the untouched draws over, so both index uploads and gathers run inside the step graph;
    "concat" passes plain tensor dicts, which the policy concatenates as the reference does."""
    from offlinerlkit_b200.buffer import ReplayBuffer
    m = g.meta
    policy = build_policy(m, device)
    load_state(policy, initial_state(m))
    policy.train()
    datasets = (g.dataset(), g.fake_dataset())
    bufs = []
    for d in datasets:
        b = ReplayBuffer(m["n_data"], (m["O"],), np.float32, m["A"], np.float32, device=device)
        b.load_dataset(d)
        bufs.append(b)
    states = []
    for off in (0, 1):          # make_golden.gen_combo: the real stream runs under np_seed, the fake one under np_seed + 1
        np.random.seed(m["np_seed"] + off)
        states.append(np.random.get_state())
    lr_atol = 2.5 * max(v for k, v in m["hyper"].items() if k.endswith("_lr"))
    sizes, keys = (m["n_real"], m["n_fake"]), ("idx", "fake_idx")
    tap, ora = None, make_oracle(m)
    for t in range(m["n_steps"]):
        parts = []
        for i in range(2):
            np.random.set_state(states[i])
            parts.append(bufs[i].sample(sizes[i]))
            states[i] = np.random.get_state()
            if mode != "lazy":
                torch.cuda.synchronize()
                assert np.array_equal(parts[i].indices.cpu().numpy(), g[keys[i]][t]), "index stream differs from the reference"
        if mode != "lazy":
            ref_b = g.batch(t, datasets[0])
            for part, name in zip(parts, ("real", "fake")):
                for k, v in ref_b[name].items():
                    assert torch.equal(part[k].cpu().reshape(v.shape), v), f"gather not bit-exact: {name}.{k}"
        if mode == "concat":
            parts = [{k: v.clone() for k, v in part.items()} for part in parts]
        if precision is not None and t == 0:
            policy._split = sizes
            policy.engine(m["B"]).precision = precision
        if tap is not None:
            tap.snapshot()
            sync_oracle_to_engine(ora, policy)
        out = policy.learn({"real": parts[0], "fake": parts[1]}, noise=g.noise(t))
        if tap is None:         # the engine exists after the first learn: its moments started at zero
            tap = EngineGrads(policy, policy._engine)
            tap.prev = {k: torch.zeros_like(v) for k, v in tap._moments().items()}
        check_step_grads(g, t, tap, ora, g.batch(t, datasets[0]), g.noise(t), tol)
        ref = g.losses(t)
        if verbose:
            print(f"step {t}: engine {out}\n        golden {ref}", flush=True)
        assert out.keys() == ref.keys(), (out.keys(), ref.keys())
        for k in ref:
            assert abs(out[k] - ref[k]) <= tol * max(1.0, abs(ref[k])), (t, k, out[k], ref[k])
        sd = {k: v.detach().cpu() for k, v in policy.state_dict().items()}
        assert_stats_close(sd, g.group(f"stats{t}"), tol=tol, lr_atol=lr_atol * (1 + 0.8 * t))
    post = g.group("post")
    if post:
        sd = {k: v.detach().cpu() for k, v in policy.state_dict().items()}
        for k, v in post.items():
            if v.dtype.kind == "f" and "saved_" not in k:
                err = np.abs(sd[k].numpy() - v).max()
                assert err <= tol * np.abs(v).max() + lr_atol, (k, err)
    return policy
