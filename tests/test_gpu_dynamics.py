"""GPU: MOPO ensemble dynamics (learn / validate / step) and MOPOPolicy.rollout vs golden vectors from the reference."""
import numpy as np
import pytest
import torch

from tests.helpers import (Golden, initial_state, assert_stats_close, assert_grad_stats_close, assert_grads_close, rel_err,
                           cfg5_setup, array_stats)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


def _build_dynamics(m, state, mu, std, term):
    from offlinerlkit_b200.modules import EnsembleDynamicsModel
    from offlinerlkit_b200.dynamics import EnsembleDynamics
    from offlinerlkit_b200.utils.scaler import StandardScaler
    from offlinerlkit_b200.utils import termination_fns as T
    model = EnsembleDynamicsModel(m["O"], m["A"], m["hidden"] if "hidden" in m and m["algo"] == "dynamics" else m["dyn_hidden"],
                                  num_ensemble=m["E"], num_elites=m["n_elites"], weight_decays=m["weight_decays"], device=DEV)
    missing, unexpected = model.load_state_dict(state, strict=False)
    assert not unexpected
    optim = torch.optim.Adam(model.parameters(), lr=m.get("lr", 1e-3))
    fn = {"halfcheetah": T.termination_fn_halfcheetah, "hopper": T.termination_fn_hopper, "walker2d": T.termination_fn_walker2d}[m["term"]]
    return EnsembleDynamics(model, optim, StandardScaler(mu, std), fn, penalty_coef=m.get("penalty_coef", 0.5))


@pytest.mark.parametrize("name", ["dynamics_small", "dynamics_hc"])
def test_dynamics_learn_validate_step(name):
    from offlinerlkit_b200.synthetic import make_dataset
    g = Golden(name)
    m = g.meta
    d = make_dataset(m["n_data"], m["O"], m["A"], seed=m["data_seed"])
    x = np.concatenate([d["observations"], d["actions"]], axis=-1)
    y = np.concatenate([d["next_observations"] - d["observations"], d["rewards"].reshape(-1, 1)], axis=-1)
    mu, std = g["scaler_mu"], g["scaler_std"]
    xs = (x - mu) / std
    dyn = _build_dynamics(m, initial_state(m), mu, std, m["term"])
    boot = g["boot"]
    # all mini-batches but the last, then the last one alone: its gradients are what the fixture's fingerprints describe
    from tests.gpu_common import EngineGrads
    from oracle import dynamics as odyn
    nb, B = m["n_batches"], m["B"]
    head = slice(0, (nb - 1) * B)
    loss_a = dyn.learn(xs[boot][:, head], y[boot][:, head], batch_size=B)
    tap = EngineGrads(dyn.model, dyn.engine, param_sets=[dyn.engine.ps])
    tap.snapshot()
    loss_b = dyn.learn(xs[boot][:, (nb - 1) * B:], y[boot][:, (nb - 1) * B:], batch_size=B)
    loss = (loss_a * (nb - 1) + loss_b) / nb
    assert loss == pytest.approx(float(g["learn_loss"]), rel=TOL)
    stats = g.group("gradstats_last")
    got = tap.after(stats.keys())
    assert_grad_stats_close(got, stats, tol=TOL, what=name)
    ora = odyn.DynamicsOracle(initial_state(m), m["weight_decays"], lr=m["lr"])      # full tensors from the pinned oracle
    ora.learn(xs[boot], y[boot], batch_size=B)
    assert_grads_close(got, {k: ora.grads[k] for k in stats}, tol=TOL, what=name)
    sd = {k: v.detach().cpu() for k, v in dyn.model.state_dict().items()}
    assert_stats_close(sd, g.group("stats"), tol=TOL, lr_atol=2.5 * m["lr"])
    val = dyn.validate(xs[:m["holdout"]], y[:m["holdout"]])
    assert rel_err(val, g["val"]) < TOL
    # imagination step: in "numpy" mode the facade consumes np.random exactly like the reference (normal then choice)
    assert dyn.rng == "device"          # the default draws on the GPU
    dyn.rng = "numpy"
    np.random.seed(11)
    nobs, rew, term, info = dyn.step(g["step_obs"], g["step_act"])
    assert rel_err(nobs, g["step_next_obs"]) < TOL and rel_err(rew, g["step_reward"]) < TOL
    assert np.array_equal(term, g["step_terminal"])
    assert rel_err(info["penalty"], g["step_penalty"]) < TOL and rel_err(info["raw_reward"], g["step_raw_reward"]) < TOL
    # the two disagreement penalties (ensemble_dynamics.py:63-70) on the same draws
    for mode in ("pairwise-diff", "ensemble_std"):
        dyn._uncertainty_mode = mode
        np.random.seed(11)
        _, rew_m, _, info_m = dyn.step(g["step_obs"], g["step_act"])
        assert rel_err(info_m["penalty"], g["step_penalty_" + mode]) < 5 * TOL, mode
        assert rel_err(rew_m, g["step_reward_" + mode]) < TOL, mode
    dyn._uncertainty_mode = "aleatoric"
    # device-side noise: same distribution, different stream -> only sanity-check shapes / finiteness / penalty
    dyn.rng = "device"
    nobs2, rew2, term2, info2 = dyn.step(g["step_obs"], g["step_act"])
    assert nobs2.shape == nobs.shape and np.isfinite(nobs2).all()
    assert rel_err(info2["penalty"], g["step_penalty"]) < TOL


@pytest.mark.parametrize("name", ["rollout_small", "combo_rollout_uniform"])
def test_mopo_rollout_matches_reference(name):
    """MOPOPolicy.rollout, and COMBOPolicy.rollout with uniform_rollout=True (combo.py:67-107)."""
    from offlinerlkit_b200.nets import MLP
    from offlinerlkit_b200.modules import ActorProb, Critic, TanhDiagGaussian
    from offlinerlkit_b200.policy import MOPOPolicy, COMBOPolicy
    from tests.gpu_common import Box
    g = Golden(name)
    m = g.meta
    uniform = bool(m.get("uniform"))
    O, A, hid = m["O"], m["A"], m["hidden"]
    dyn_state = {k: torch.from_numpy(v) for k, v in g.group("dyn").items()}
    dyn = _build_dynamics(m, dyn_state, g["scaler_mu"], g["scaler_std"], m["term"])
    bb = MLP(O, hid)
    actor = ActorProb(bb, TanhDiagGaussian(bb.output_dim, A, unbounded=True, conditioned_sigma=True), DEV)
    actor.load_state_dict({k: torch.from_numpy(v) for k, v in g.group("actor").items()})
    c1, c2 = Critic(MLP(O + A, hid), DEV), Critic(MLP(O + A, hid), DEV)
    adam = lambda mod: torch.optim.Adam(mod.parameters(), lr=1e-4)
    if uniform:
        pol = COMBOPolicy(dyn, actor, c1, c2, adam(actor), adam(c1), adam(c2), action_space=Box(-1, 1, (A,)), alpha=0.2,
                          uniform_rollout=True)
    else:
        pol = MOPOPolicy(dyn, actor, c1, c2, adam(actor), adam(c1), adam(c2), alpha=0.2)
    counts, E, D = g["counts"], m["E"], O + 1
    eps, nrm, mid, r = [], [], [], 0
    for c in counts:
        eps.append(g["uniform_actions" if uniform else "eps"][r:r + c])
        nrm.append(g["normal"][:, r * D:(r + c) * D].reshape(E, c, D))
        mid.append(g["midx"][r:r + c])
        r += c
    out, info = pol.rollout(g["init"], m["horizon"], noise={("actions" if uniform else "eps"): eps, "normal": nrm, "midx": mid})
    assert info["num_transitions"] == int(counts.sum())
    for k in ("obss", "next_obss", "actions", "rewards"):
        assert out[k].shape == g["out|" + k].shape, k
        assert rel_err(out[k], g["out|" + k]) < 2e-4, k
    assert np.array_equal(out["terminals"], g["out|terminals"])
    assert info["reward_mean"] == pytest.approx(float(g["reward_mean"]), rel=1e-4)
    # performance mode (device noise): runs, keeps the output contract
    out2, info2 = pol.rollout(g["init"], m["horizon"])
    assert set(out2) == {"obss", "next_obss", "actions", "rewards", "terminals"} and out2["terminals"].dtype == bool
    assert len(out2["obss"]) == info2["num_transitions"]
    if uniform:
        assert np.abs(out2["actions"]).max() <= 1.0 and abs(float(out2["actions"].mean())) < 0.2


@pytest.mark.parametrize("precision,tol", [("tf32x3", 2e-5), ("tf32", 5e-3)])
def test_ensemble_forward_on_tensor_cores(precision, tol):
    """Rollout-sized ensemble inference (dynamics_module.py:83-96) goes through the tcgen05 kernel: members as groups,
    'io' weights as an MN-major B operand, hidden width 200 inside 224-column tiles, K = 200 with a zero-filled tail
    slab, Swish in the epilogue.  Checked against a float64 evaluation of the same layers."""
    from offlinerlkit_b200.modules import EnsembleDynamicsModel
    from offlinerlkit_b200.dynamics import EnsembleDynamics
    from offlinerlkit_b200.utils.scaler import StandardScaler
    from offlinerlkit_b200.utils.termination_fns import termination_fn_halfcheetah
    O, A, E, S = 17, 6, 3, 5000
    torch.manual_seed(3)
    model = EnsembleDynamicsModel(O, A, [200, 200, 200], num_ensemble=E, num_elites=2,
                                  weight_decays=[2.5e-5, 5e-5, 7.5e-5, 1e-4], device=DEV)
    with torch.no_grad():
        for lay in list(model.backbones) + [model.output_layer]:
            lay.bias.normal_(0.0, 0.1)
    dyn = EnsembleDynamics(model, torch.optim.Adam(model.parameters(), lr=1e-3), StandardScaler(), termination_fn_halfcheetah)
    eng = dyn.engine
    eng.precision = precision
    x = torch.randn(S, O + A, device=DEV)
    run = eng._forward(x)
    torch.cuda.synchronize()
    _, plan, _ = eng._fwd_runs[S]
    assert all(lbl.endswith(".tc") for lbl, _ in plan.flat_ops), [lbl for lbl, _ in plan.flat_ops]
    h = x.double().unsqueeze(0).repeat(E, 1, 1)
    layers = list(model.backbones) + [model.output_layer]
    with torch.no_grad():
        for i, lay in enumerate(layers):
            h = torch.bmm(h, lay.weight.double()) + lay.bias.double()
            if i < len(layers) - 1:
                h = h * torch.sigmoid(h)
    assert rel_err(run.OUT.cpu().numpy(), h.cpu().numpy()) < tol


def test_dynamics_train_shuffle_overlap_is_invisible(tmp_path, monkeypatch):
    """EnsembleDynamics.train (ensemble_dynamics.py:111-176) reshuffles the bootstrap indices on a worker thread while
    the device runs the epoch.  With and without the overlap the run must consume np.random identically and end with
    bit-identical parameters, elites and scaler."""
    from offlinerlkit_b200.modules import EnsembleDynamicsModel
    from offlinerlkit_b200.dynamics import EnsembleDynamics
    from offlinerlkit_b200.utils.logger import Logger
    from offlinerlkit_b200.utils.scaler import StandardScaler
    from offlinerlkit_b200.utils.termination_fns import termination_fn_halfcheetah
    from offlinerlkit_b200.synthetic import make_dataset
    O, A = 5, 3
    data = make_dataset(3000, O, A, seed=2)
    data["rewards"] = data["rewards"].reshape(-1, 1)
    results = []
    for flag in ("1", "0"):
        monkeypatch.setenv("ORLK_DYN_SHUFFLE_OVERLAP", flag)
        torch.manual_seed(4)
        np.random.seed(4)
        model = EnsembleDynamicsModel(O, A, [24, 24], num_ensemble=3, num_elites=2, weight_decays=[2.5e-5, 5e-5, 7.5e-5], device=DEV)
        dyn = EnsembleDynamics(model, torch.optim.Adam(model.parameters(), lr=1e-3), StandardScaler(), termination_fn_halfcheetah)
        logger = Logger(str(tmp_path / flag))
        logger.quiet = True
        dyn.train(data, logger, max_epochs=3, max_epochs_since_update=5)
        results.append(({k: v.detach().clone() for k, v in model.state_dict().items()}, np.random.get_state()[1].copy(),
                        dyn.scaler.mu.copy()))
    (sd1, st1, mu1), (sd0, st0, mu0) = results
    assert np.array_equal(st1, st0) and np.array_equal(mu1, mu0)
    for k in sd1:
        assert torch.equal(sd1[k], sd0[k]), k


@pytest.mark.parametrize("name", ["dynamics_sample_next_small", "dynamics_sample_next_hc"])
def test_sample_next_obss_matches_reference(name):
    """EnsembleDynamics.sample_next_obss (ensemble_dynamics.py:81-99) with the reference's draws injected; elites as set by
    ``set_elites`` (not the first n members)."""
    g = Golden(name)
    m = dict(g.meta, algo="dynamics", term="halfcheetah")
    dyn = _build_dynamics(m, initial_state(g.meta), g["scaler_mu"], g["scaler_std"], "halfcheetah")
    dyn.model.set_elites([int(e) for e in g["elites"]])
    got = dyn.sample_next_obss(torch.as_tensor(g["obs"]), torch.as_tensor(g["act"]), m["num_samples"], noise=g["noise"])
    assert tuple(got.shape) == (m["num_samples"], m["n_elites"], m["S"], m["O"]) and got.is_cuda
    assert rel_err(got.cpu().numpy(), g["next_obss"]) < 2e-4
    free = dyn.sample_next_obss(torch.as_tensor(g["obs"]), torch.as_tensor(g["act"]), 3)        # own draws: shape and sanity
    assert tuple(free.shape) == (3, m["n_elites"], m["S"], m["O"]) and bool(torch.isfinite(free).all())


def test_dynamics_learn_eager_equals_graph():
    """The branched training step (decay sums beside the forward pass, weight gradients beside the input gradients)
    gives bit-identical parameters whether it is replayed as a CUDA graph or launched eagerly on side streams."""
    from offlinerlkit_b200.synthetic import make_dataset
    g = Golden("dynamics_small")
    m = g.meta
    d = make_dataset(m["n_data"], m["O"], m["A"], seed=m["data_seed"])
    x = np.concatenate([d["observations"], d["actions"]], axis=-1)
    y = np.concatenate([d["next_observations"] - d["observations"], d["rewards"].reshape(-1, 1)], axis=-1)
    xs = (x - g["scaler_mu"]) / g["scaler_std"]
    boot = g["boot"]
    states = []
    for use_graph in (True, False):
        dyn = _build_dynamics(m, initial_state(m), g["scaler_mu"], g["scaler_std"], m["term"])
        dyn.engine.use_graph = use_graph
        loss = dyn.learn(xs[boot], y[boot], batch_size=m["B"])
        states.append((loss, {k: v.detach().clone() for k, v in dyn.model.state_dict().items()}))
    assert states[0][0] == states[1][0]
    for k in states[0][1]:
        assert torch.equal(states[0][1][k], states[1][1][k]), k


class _ListLogger:
    """Records what EnsembleDynamics.train logs per epoch (the Logger API subset it uses)."""

    def __init__(self, model_dir):
        self.model_dir, self.rows, self._kv = str(model_dir), [], {}

    def log(self, *a, **k):
        pass

    def logkv(self, k, v):
        self._kv[k] = float(v)

    def set_timestep(self, t):
        self._kv["timestep"] = t

    def dumpkvs(self, exclude=None):
        self.rows.append(dict(self._kv))
        self._kv = {}


def test_dynamics_train_matches_reference(tmp_path):
    """EnsembleDynamics.train end to end vs the reference's own run (ensemble_dynamics.py:111-176): the torch
    ``random_split`` holdout, the scaler, the NumPy bootstrap matrix and its per-epoch row shuffles are consumed in the
    reference's order; per-epoch train / holdout losses, the elites, the scaler and the final (load_save'd) parameters."""
    from tests.test_oracle_golden import dynamics_train_data
    g = Golden("dynamics_train_small")
    m = g.meta
    x, y = dynamics_train_data(g)
    O = m["O"]
    data = {"observations": x[:, :O], "actions": x[:, O:], "next_observations": x[:, :O] + y[:, :O], "rewards": y[:, O:]}
    from offlinerlkit_b200.utils.scaler import StandardScaler
    dyn = _build_dynamics(m, initial_state(m), None, None, m["term"])
    dyn.scaler = StandardScaler()
    logger = _ListLogger(tmp_path)
    torch.manual_seed(m["torch_seed"])
    np.random.seed(m["np_seed"])
    dyn.train(data, logger, **m["train_kw"])
    assert len(logger.rows) == m["epochs"]
    assert rel_err([r["loss/dynamics_train_loss"] for r in logger.rows], g["train_loss"]) < TOL
    assert rel_err([r["loss/dynamics_holdout_loss"] for r in logger.rows], g["holdout_loss"]) < TOL
    assert dyn.model.elites.data.tolist() == g["elites"].tolist()
    assert rel_err(dyn.scaler.mu, g["scaler_mu"]) < 1e-6 and rel_err(dyn.scaler.std, g["scaler_std"]) < 1e-6
    sd = {k: v.detach().cpu().numpy() for k, v in dyn.model.state_dict().items()}
    lr_atol = 2.5 * m["lr"]
    for k, v in g.group("post").items():
        assert np.abs(sd[k] - v).max() <= TOL * np.abs(v).max() + lr_atol, k
        assert abs(np.abs(sd[k]).sum() - np.abs(v).sum()) <= 10 * TOL * np.abs(v).sum() + 3 * lr_atol, k


@pytest.mark.parametrize("name", ["rollout_cfg5_hc", "rollout_cfg5_walker"])
def test_rollout_config5_size_matches_reference(name):
    """MOPOPolicy.rollout at BASELINE.json configs[4] size -- 50 000 start states x horizon 5, 7 members of 200 x 4 (the
    tcgen05 ensemble path and the two-launch block-scan compaction are both live) -- vs the reference's own run: survivor
    counts per step and the number of terminals exact, per-array fingerprints (norms + 64 strided rows) at 2e-4.  The
    draws are re-made from the fixture's seeds in the reference's consumption order."""
    from offlinerlkit_b200.nets import MLP
    from offlinerlkit_b200.modules import ActorProb, Critic, TanhDiagGaussian
    from offlinerlkit_b200.policy import MOPOPolicy
    g = Golden(name)
    m = g.meta
    O, A, hid, E, D = m["O"], m["A"], m["hidden"], m["E"], m["O"] + 1
    dyn_state, actor_state, mu, std, init = cfg5_setup(m)
    dyn = _build_dynamics(m, dyn_state, mu, std, m["term"])
    dyn.model.set_elites(m["elites"])
    bb = MLP(O, hid)
    actor = ActorProb(bb, TanhDiagGaussian(bb.output_dim, A, unbounded=True, conditioned_sigma=True), DEV)
    actor.load_state_dict({k[len("actor."):]: v for k, v in actor_state.items()})
    c1, c2 = Critic(MLP(O + A, hid), DEV), Critic(MLP(O + A, hid), DEV)
    adam = lambda mod: torch.optim.Adam(mod.parameters(), lr=1e-4)
    pol = MOPOPolicy(dyn, actor, c1, c2, adam(actor), adam(c1), adam(c2), alpha=0.2)
    counts = g["counts"].tolist()
    torch.manual_seed(m["torch_seed"])
    np.random.seed(m["np_seed"])
    elites = np.asarray(m["elites"])
    eps, nrm, mid = [], [], []
    for c in counts:            # per imagined step: eps (torch CPU generator), then normal and choice (NumPy global generator)
        eps.append(torch.randn(c, A).numpy())
        nrm.append(np.random.normal(size=(E, c, D)))
        mid.append(np.random.choice(elites, size=c))
    out, info = pol.rollout(init, m["horizon"], noise={"eps": eps, "normal": nrm, "midx": mid})
    assert info["num_transitions"] == m["num_transitions"] == sum(counts)
    assert int(out["terminals"].sum()) == int(g["terminal_count"])
    for k, v in out.items():
        got, ref = array_stats(v.astype(np.float64) if v.dtype == bool else v), g["outstats|" + k]
        assert got.shape == ref.shape, k
        assert rel_err(got[:3], ref[:3]) < 2e-4, (k, got[:3], ref[:3])
        assert rel_err(got[3:], ref[3:]) < 2e-4, k
    assert info["reward_mean"] == pytest.approx(float(g["reward_mean"]), rel=2e-4)


def _cfg5_policy(name):
    from offlinerlkit_b200.nets import MLP
    from offlinerlkit_b200.modules import ActorProb, Critic, TanhDiagGaussian
    from offlinerlkit_b200.policy import MOPOPolicy
    g = Golden(name)
    m = g.meta
    O, A, hid = m["O"], m["A"], m["hidden"]
    dyn_state, actor_state, mu, std, init = cfg5_setup(m)
    dyn = _build_dynamics(m, dyn_state, mu, std, m["term"])
    dyn.model.set_elites(m["elites"])
    bb = MLP(O, hid)
    actor = ActorProb(bb, TanhDiagGaussian(bb.output_dim, A, unbounded=True, conditioned_sigma=True), DEV)
    actor.load_state_dict({k[len("actor."):]: v for k, v in actor_state.items()})
    c1, c2 = Critic(MLP(O + A, hid), DEV), Critic(MLP(O + A, hid), DEV)
    adam = lambda mod: torch.optim.Adam(mod.parameters(), lr=1e-4)
    return MOPOPolicy(dyn, actor, c1, c2, adam(actor), adam(c1), adam(c2), alpha=0.2), init, m


def test_default_rollout_has_no_per_step_sync_and_equals_the_per_step_loop():
    """The default rollout (device noise, nothing injected) runs the horizon without a host round trip per step.  With no
    terminations (halfcheetah) it must equal the per-step loop bit for bit -- both draw the same Philox numbers."""
    outs = []
    for sync in (False, True):
        pol, init, m = _cfg5_policy("rollout_cfg5_hc")
        assert pol.dynamics.rng == "device"
        out, info = pol.rollout(init[:6000], m["horizon"])          # the first call creates the engine
        pol._roll.sync_loop = sync
        pol._roll.philox_counter.zero_()
        pol.dynamics.engine.philox_counter.zero_()
        out, info = pol.rollout(init[:6000], m["horizon"])
        assert info["num_transitions"] == 6000 * m["horizon"] and out["terminals"].dtype == bool
        outs.append(out)
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k]), k


def test_default_rollout_drops_dead_rows_like_the_reference():
    """walker2d-style terminations: rows that terminate at step t must not appear at step t+1, the survivors keep their
    order (mopo.py:69-73), and every array has one row per transition."""
    pol, init, m = _cfg5_policy("rollout_cfg5_walker")
    S, h = 20000, m["horizon"]
    out, info = pol.rollout(init[:S], h)
    n = info["num_transitions"]
    assert all(len(v) == n for v in out.values()) and n < S * h
    r0, cnt, total = 0, S, 0
    for t in range(h):
        obs_t, nobs_t, term_t = out["obss"][r0:r0 + cnt], out["next_obss"][r0:r0 + cnt], out["terminals"][r0:r0 + cnt, 0]
        if t == 0:
            assert np.array_equal(obs_t, init[:S])
        total += cnt
        alive = ~term_t
        nxt = int(alive.sum())
        r0 += cnt
        if t + 1 < h and nxt > 0:
            assert np.array_equal(out["obss"][r0:r0 + nxt], nobs_t[alive]), t
        cnt = nxt
        if cnt == 0:
            break
    assert total == n
    assert 0.005 < out["terminals"].mean() < 0.5
    # device tensors on request (what MBPolicyTrainer hands to fake_buffer.add_batch)
    dout, dinfo = pol.rollout(init[:S], h, device_out=True)
    assert all(v.is_cuda for v in dout.values()) and dout["obss"].shape[0] == dinfo["num_transitions"]
