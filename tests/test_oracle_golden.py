"""CPU: pin the oracle (oracle/) against the golden vectors produced by the real reference."""
import numpy as np
import pytest
import torch

from oracle import algos, dynamics as odyn, replay as oreplay
from tests.helpers import (Golden, initial_state, assert_stats_close, assert_grad_stats_close, rel_err, FIELDS, cfg5_setup,
                           array_stats)

TOL = 2e-5


def _alpha(meta):
    return (meta["target_entropy"], 0.0, meta["alpha_lr"])


def _run(g, ora, with_noise=True):
    data = g.dataset()
    for t in range(g.meta["n_steps"]):
        out = ora.step(g.batch(t, data), g.noise(t) if with_noise else None)
        ref = g.losses(t)
        assert out.keys() == ref.keys()
        for k in ref:
            assert out[k] == pytest.approx(ref[k], rel=TOL, abs=TOL), (t, k)
        assert_stats_close(ora.state_dict(), g.group(f"stats{t}"), tol=TOL)
        # the gradients every Adam step of the reference consumed (step pre-hooks on the reference's own optimisers)
        assert_grad_stats_close(ora.grads, g.group(f"gradstats{t}"), tol=TOL, what=f"step {t}")
        for k, v in g.group(f"grads{t}").items():
            assert rel_err(ora.grads[k].numpy(), v) < TOL, (t, k)
    post = g.group("post")
    if post:
        sd = ora.state_dict()
        for k, v in post.items():
            if v.dtype.kind == "f":
                assert rel_err(sd[k].numpy(), v) < TOL, k


@pytest.mark.parametrize("name", ["cql_small", "cql_small_lagrange", "cql_hc", "cql_hc_lagrange", "cql_hopper", "cql_small_maxq",
                                  "cql_hc_maxq", "cql_hc_stochastic_backup"])
def test_cql(name):
    g = Golden(name)
    _run(g, algos.CQLOracle(initial_state(g.meta), alpha=_alpha(g.meta), **g.meta["hyper"]))


@pytest.mark.parametrize("name", ["combo_small_mix", "combo_small_model", "combo_hc", "combo_hc_model"])
def test_combo(name):
    """combo.py:109-243 on real+fake batches, both rho_s settings, with and without the Lagrange multiplier."""
    g = Golden(name)
    _run(g, algos.COMBOOracle(initial_state(g.meta), rho_s=g.meta["rho_s"], alpha=_alpha(g.meta), **g.meta["hyper"]))


@pytest.mark.parametrize("name", ["sac_small", "sac_hc"])
def test_sac(name):
    g = Golden(name)
    _run(g, algos.SACOracle(initial_state(g.meta), alpha=_alpha(g.meta), **g.meta["hyper"]))


@pytest.mark.parametrize("name", ["edac_small", "edac_hc", "edac_small_maxq", "edac_hopper_e50"])
def test_edac(name):
    g = Golden(name)
    _run(g, algos.EDACOracle(initial_state(g.meta), alpha=_alpha(g.meta), **g.meta["hyper"]))


@pytest.mark.parametrize("name", ["iql_small", "iql_walker", "iql_walker_b1024"])
def test_iql(name):
    g = Golden(name)
    _run(g, algos.IQLOracle(initial_state(g.meta), **g.meta["hyper"]), with_noise=False)


@pytest.mark.parametrize("name", ["td3bc_small", "td3bc_walker", "td3bc_walker_b1024"])
def test_td3bc(name):
    g = Golden(name)
    _run(g, algos.TD3BCOracle(initial_state(g.meta), **g.meta["hyper"]))


def test_replay_indices_and_gather():
    """buffer.py:96-106: the legacy NumPy global generator gives the golden indices; gather is bit-exact."""
    g = Golden("cql_small")
    data = g.dataset()
    np.random.seed(g.meta["np_seed"])
    for t in range(g.meta["n_steps"]):
        idx = oreplay.draw_indices(g.meta["n_data"], g.meta["B"])
        assert np.array_equal(idx, g["idx"][t])
        got = oreplay.gather(data, idx)
        for k in FIELDS:
            assert np.array_equal(got[k], data[k][idx])


def _dyn_inputs(g):
    m = g.meta
    from offlinerlkit_b200.synthetic import make_dataset
    d = make_dataset(m["n_data"], m["O"], m["A"], seed=m["data_seed"])
    x = np.concatenate([d["observations"], d["actions"]], axis=-1)
    y = np.concatenate([d["next_observations"] - d["observations"], d["rewards"].reshape(-1, 1)], axis=-1)
    mu, std = odyn.scaler_fit(x)
    assert np.array_equal(mu, g["scaler_mu"]) and np.array_equal(std, g["scaler_std"])
    return (x - mu) / std, y, mu, std


@pytest.mark.parametrize("name", ["dynamics_small", "dynamics_hc"])
def test_dynamics(name):
    g = Golden(name)
    m = g.meta
    x, y, mu, std = _dyn_inputs(g)
    ora = odyn.DynamicsOracle(initial_state(m), m["weight_decays"], lr=m["lr"])
    boot = g["boot"]
    loss = ora.learn(x[boot], y[boot], batch_size=m["B"])
    assert loss == pytest.approx(float(g["learn_loss"]), rel=TOL)
    assert_stats_close({k: v.detach() for k, v in ora.p.items()}, g.group("stats"), tol=TOL)
    assert_grad_stats_close(ora.grads, g.group("gradstats_last"), tol=TOL, what="last mini-batch")
    val = ora.validate(x[:m["holdout"]], y[:m["holdout"]])
    assert rel_err(val, g["val"]) < TOL
    fn = {"halfcheetah": odyn.term_halfcheetah, "hopper": odyn.term_hopper, "walker2d": odyn.term_walker2d}[m["term"]]
    nobs, rew, term, info = ora.step(g["step_obs"], g["step_act"], mu, std, fn, m["penalty_coef"], g["step_noise"],
                                     g["step_midx"])
    assert rel_err(nobs, g["step_next_obs"]) < TOL and rel_err(rew, g["step_reward"]) < TOL
    assert np.array_equal(term, g["step_terminal"])
    assert rel_err(info["penalty"], g["step_penalty"]) < TOL
    for mode in ("pairwise-diff", "ensemble_std"):      # ensemble_dynamics.py:63-70
        _, rew_m, _, info_m = ora.step(g["step_obs"], g["step_act"], mu, std, fn, m["penalty_coef"], g["step_noise"],
                                       g["step_midx"], uncertainty_mode=mode)
        assert rel_err(info_m["penalty"], g["step_penalty_" + mode]) < TOL, mode
        assert rel_err(rew_m, g["step_reward_" + mode]) < TOL, mode


@pytest.mark.parametrize("name", ["dynamics_sample_next_small", "dynamics_sample_next_hc"])
def test_dynamics_sample_next_obss(name):
    """ensemble_dynamics.py:81-99 (MOBILE's uncertainty samples) against the reference's output under a seeded torch
    generator, replayed from the stored draws."""
    g = Golden(name)
    m = g.meta
    ora = odyn.DynamicsOracle(initial_state(m), m["weight_decays"], lr=1e-3)
    got = odyn.sample_next_obss(ora, g["obs"], g["act"], g["scaler_mu"], g["scaler_std"], g["elites"], g["noise"])
    assert got.shape == (m["num_samples"], m["n_elites"], m["S"], m["O"])
    assert rel_err(got, g["next_obss"]) < TOL


@pytest.mark.parametrize("name", ["rollout_small", "combo_rollout_uniform"])
def test_rollout_compaction(name):
    """mopo.py:45-79 / combo.py:67-107: stable survivor compaction and per-step draw order, replayed with stored noise."""
    from oracle import nets
    g = Golden(name)
    m = g.meta
    dyn_state = {k: torch.from_numpy(v) for k, v in g.group("dyn").items()}
    actor = {"actor." + k: torch.from_numpy(v) for k, v in g.group("actor").items()}
    ora = odyn.DynamicsOracle(dyn_state, m["weight_decays"])
    E, D = m["E"], m["O"] + 1
    cur = {"row": 0}

    def select_action(obs):
        n = len(obs)
        if m.get("uniform"):        # combo.py:82-86: uniform actions, no actor pass
            return g["uniform_actions"][cur["row"]:cur["row"] + n]
        eps = torch.from_numpy(g["eps"][cur["row"]:cur["row"] + n])
        with torch.no_grad():
            a, _ = nets.actforward(actor, "actor", torch.from_numpy(obs), eps)
        return a.numpy()

    def step(obs, act):
        n = len(obs)
        r0 = cur["row"]
        noise = g["normal"][:, r0 * D:(r0 + n) * D].reshape(E, n, D)
        out = ora.step(obs, act, g["scaler_mu"], g["scaler_std"], odyn.term_hopper, m["penalty_coef"], noise,
                       g["midx"][r0:r0 + n])
        cur["row"] += n
        return out

    out, info = odyn.rollout(select_action, step, g["init"], m["horizon"])
    assert info["num_transitions"] == int(g["counts"].sum())
    for k in ("obss", "next_obss", "actions", "rewards"):
        assert rel_err(out[k], g["out|" + k]) < 1e-4, k
    assert np.array_equal(out["terminals"], g["out|terminals"])


def test_dynamics_train_loop():
    """ensemble_dynamics.py:111-176: holdout split, bootstrap matrix, per-epoch losses, update_save, elites, load_save."""
    from offlinerlkit_b200.synthetic import make_dataset
    g = Golden("dynamics_train_small")
    m = g.meta
    x, y = dynamics_train_data(g)
    ora = odyn.DynamicsOracle(initial_state(m), m["weight_decays"], lr=m["lr"])
    torch.manual_seed(m["torch_seed"])
    np.random.seed(m["np_seed"])
    res = odyn.train(ora, x, y, m["n_elites"], **m["train_kw"])
    assert res["epochs"] == m["epochs"]
    assert rel_err([r[0] for r in res["log"]], g["train_loss"]) < TOL
    assert rel_err([r[1] for r in res["log"]], g["holdout_loss"]) < TOL
    assert rel_err([r[2] for r in res["log"]], g["member_holdout"]) < TOL
    assert res["elites"] == g["elites"].tolist()
    assert np.array_equal(res["mu"], g["scaler_mu"]) and np.array_equal(res["std"], g["scaler_std"])
    for k, v in g.group("post").items():
        assert rel_err(ora.p[k].detach().numpy(), v) < TOL, k


def dynamics_train_data(g):
    """The learnable synthetic transition set of make_golden.gen_dynamics_train."""
    from offlinerlkit_b200.synthetic import make_dataset
    m = g.meta
    data = make_dataset(m["n_data"], m["O"], m["A"], seed=m["data_seed"])
    xin = np.concatenate([data["observations"], data["actions"]], 1)
    nobs = (data["observations"] + np.tanh(xin @ g["Wd"]) + 0.05 * data["next_observations"]).astype(np.float32)
    rew = (np.sin(xin.sum(1, keepdims=True)) + 0.05 * data["rewards"].reshape(-1, 1)).astype(np.float32)
    return xin, np.concatenate([nobs - data["observations"], rew], axis=-1)


@pytest.mark.parametrize("name", ["rollout_cfg5_walker"])
def test_rollout_config5_size(name):
    """mopo.py:45-79 at BASELINE.json configs[4] size (50 000 starts x horizon 5): survivor counts exact, array
    fingerprints; the draws are re-made from the fixture's seeds in the reference's consumption order."""
    from oracle import nets
    g = Golden(name)
    m = g.meta
    dyn_state, actor, mu, std, init = cfg5_setup(m)
    ora = odyn.DynamicsOracle(dyn_state, m["weight_decays"])
    E, D, A = m["E"], m["O"] + 1, m["A"]
    elites = np.asarray(m["elites"])
    fn = {"halfcheetah": odyn.term_halfcheetah, "walker2d": odyn.term_walker2d}[m["term"]]
    torch.manual_seed(m["torch_seed"])
    np.random.seed(m["np_seed"])
    counts = []

    def select_action(obs):
        counts.append(len(obs))
        with torch.no_grad():
            a, _ = nets.actforward(actor, "actor", torch.from_numpy(obs), torch.randn(len(obs), A))
        return a.numpy()

    def step(obs, act):
        n = len(obs)
        noise = np.random.normal(size=(E, n, D))
        return ora.step(obs, act, mu, std, fn, m["penalty_coef"], noise, np.random.choice(elites, size=n))

    out, info = odyn.rollout(select_action, step, init, m["horizon"])
    assert counts == g["counts"].tolist()
    assert info["num_transitions"] == m["num_transitions"]
    assert int(out["terminals"].sum()) == int(g["terminal_count"])
    for k, v in out.items():
        got, ref = array_stats(v.astype(np.float64) if v.dtype == bool else v), g["outstats|" + k]
        assert rel_err(got, ref) < 1e-4, k
