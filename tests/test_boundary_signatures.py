"""CPU: the facade classes keep the reference's call surface -- constructor and method signatures
(parameter names, order, defaults) and ``state_dict()`` key sets, compared with the reference itself (baseline/_ref, the
pip-installed unmodified reference; skipped where it is absent).  Extra facade parameters are allowed only at the end and
only with defaults (``noise=None`` for parity tests, ``device_out=False`` for device-resident rollouts)."""
import inspect

import numpy as np
import pytest
import torch


def _ref():
    from baseline import reference_runner as rr
    ok, why = rr.available()
    if not ok:
        pytest.skip(f"reference not installed: {why}")
    import offlinerlkit
    return offlinerlkit


def _params(fn):
    return [(p.name, p.default if p.default is not inspect._empty else "<required>", p.kind)
            for p in inspect.signature(fn).parameters.values() if p.name != "self"]


def _same_default(a, b):
    if isinstance(a, float) and isinstance(b, float):
        return a == b
    if inspect.isclass(a) and inspect.isclass(b):
        return a.__name__ == b.__name__            # e.g. activation=Swish: each package's own class of that name
    return a is b or a == b or (type(a).__name__ == type(b).__name__ and repr(a) == repr(b))


def _check(ref_fn, our_fn, what):
    r, o = _params(ref_fn), _params(our_fn)
    assert len(o) >= len(r), f"{what}: facade has fewer parameters {o} vs {r}"
    for (rn, rd, rk), (on, od, ok) in zip(r, o):
        assert rn == on, f"{what}: parameter {on!r} should be {rn!r}"
        assert (rd == "<required>") == (od == "<required>") and (rd == "<required>" or _same_default(rd, od)), \
            f"{what}: default of {rn!r}: {od!r} vs reference {rd!r}"
    for on, od, ok in o[len(r):]:
        assert od != "<required>" or ok in (inspect.Parameter.VAR_KEYWORD, inspect.Parameter.VAR_POSITIONAL), \
            f"{what}: extra facade parameter {on!r} must have a default"


CLASSES = [
    ("buffer", "ReplayBuffer", ["__init__", "add", "add_batch", "load_dataset", "normalize_obs", "sample", "sample_all"]),
    ("nets", "MLP", ["__init__", "forward"]),
    ("nets", "EnsembleLinear", ["__init__", "forward", "load_save", "update_save", "get_decay_loss"]),
    ("modules", "ActorProb", ["__init__", "forward"]),
    ("modules", "Actor", ["__init__", "forward"]),
    ("modules", "Critic", ["__init__", "forward"]),
    ("modules", "EnsembleCritic", ["__init__", "forward"]),
    ("modules", "TanhDiagGaussian", ["__init__", "forward"]),
    ("modules", "DiagGaussian", ["__init__", "forward"]),
    ("modules", "EnsembleDynamicsModel", ["__init__", "forward", "load_save", "update_save", "get_decay_loss", "set_elites",
                                          "random_elite_idxs"]),
    ("dynamics", "EnsembleDynamics", ["__init__", "step", "sample_next_obss", "format_samples_for_training", "train", "learn",
                                      "validate", "select_elites", "save", "load"]),
    ("policy", "SACPolicy", ["__init__", "train", "eval", "actforward", "select_action", "learn"]),
    ("policy", "CQLPolicy", ["__init__", "learn", "select_action"]),
    ("policy", "EDACPolicy", ["__init__", "train", "eval", "actforward", "select_action", "learn"]),
    ("policy", "IQLPolicy", ["__init__", "train", "eval", "select_action", "learn"]),
    ("policy", "TD3BCPolicy", ["__init__", "train", "eval", "select_action", "learn"]),
    ("policy", "MOPOPolicy", ["__init__", "rollout", "learn"]),
    ("policy", "COMBOPolicy", ["__init__", "rollout", "learn"]),
    ("policy_trainer", "MFPolicyTrainer", ["__init__", "train"]),
    ("policy_trainer", "MBPolicyTrainer", ["__init__", "train"]),
    ("utils.scaler", "StandardScaler", ["__init__", "fit", "transform", "inverse_transform", "save_scaler", "load_scaler"]),
]


@pytest.mark.parametrize("mod,cls,methods", CLASSES, ids=[c[1] for c in CLASSES])
def test_signatures_match_the_reference(mod, cls, methods):
    import importlib
    _ref()
    rcls = getattr(importlib.import_module(f"offlinerlkit.{mod}"), cls)
    ocls = getattr(importlib.import_module(f"offlinerlkit_b200.{mod}"), cls)
    for m in methods:
        assert hasattr(ocls, m), f"{cls}.{m} missing in the facade"
        _check(getattr(rcls, m), getattr(ocls, m), f"{cls}.{m}")


def test_termination_functions_and_logger_surface():
    import importlib
    _ref()
    rt = importlib.import_module("offlinerlkit.utils.termination_fns")
    ot = importlib.import_module("offlinerlkit_b200.utils.termination_fns")
    for name in ("termination_fn_halfcheetah", "termination_fn_hopper", "termination_fn_walker2d", "get_termination_fn"):
        _check(getattr(rt, name), getattr(ot, name), name)
    rl = importlib.import_module("offlinerlkit.utils.logger")
    ol = importlib.import_module("offlinerlkit_b200.utils.logger")
    for m in ("logkv", "logkv_mean", "dumpkvs", "log", "set_timestep", "log_hyperparameters", "close"):
        _check(getattr(rl.Logger, m), getattr(ol.Logger, m), f"Logger.{m}")
    _check(rl.make_log_dirs, ol.make_log_dirs, "make_log_dirs")


def _build(pkg, algo):
    """(policy) of a small configuration, built on the CPU from package ``pkg`` ('offlinerlkit' or 'offlinerlkit_b200')."""
    import importlib
    nets, mods, pol = (importlib.import_module(f"{pkg}.{m}") for m in ("nets", "modules", "policy"))
    O, A, hid = 5, 3, [16, 16]
    adam = lambda m: torch.optim.Adam(m.parameters(), lr=1e-3)

    class Box:
        low, high, shape = np.full(A, -1.0, np.float32), np.full(A, 1.0, np.float32), (A,)

    def tanh_actor():
        bb = nets.MLP(input_dim=O, hidden_dims=hid)
        return mods.ActorProb(bb, mods.TanhDiagGaussian(latent_dim=bb.output_dim, output_dim=A, unbounded=True,
                                                        conditioned_sigma=True), "cpu")

    critic = lambda i: mods.Critic(nets.MLP(input_dim=i, hidden_dims=hid), "cpu")
    if algo in ("sac", "cql"):
        a, c1, c2 = tanh_actor(), critic(O + A), critic(O + A)
        la = torch.zeros(1, requires_grad=True)
        alpha = (-A, la, torch.optim.Adam([la], lr=1e-4))
        if algo == "sac":
            return pol.SACPolicy(a, c1, c2, adam(a), adam(c1), adam(c2), alpha=alpha)
        return pol.CQLPolicy(a, c1, c2, adam(a), adam(c1), adam(c2), action_space=Box(), alpha=alpha)
    if algo == "edac":
        a = tanh_actor()
        cs = mods.EnsembleCritic(O, A, hid, num_ensemble=3, device="cpu")
        return pol.EDACPolicy(a, cs, adam(a), adam(cs), alpha=0.2)
    if algo == "iql":
        bb = nets.MLP(input_dim=O, hidden_dims=hid, dropout_rate=None)
        a = mods.ActorProb(bb, mods.DiagGaussian(latent_dim=bb.output_dim, output_dim=A, unbounded=False, conditioned_sigma=False),
                           "cpu")
        q1, q2, v = critic(O + A), critic(O + A), critic(O)
        return pol.IQLPolicy(a, q1, q2, v, adam(a), adam(q1), adam(q2), adam(v), action_space=Box())
    if algo == "td3bc":
        a = mods.Actor(nets.MLP(input_dim=O, hidden_dims=hid), A, device="cpu")
        c1, c2 = critic(O + A), critic(O + A)
        return pol.TD3BCPolicy(a, c1, c2, adam(a), adam(c1), adam(c2))
    raise KeyError(algo)


@pytest.mark.parametrize("algo", ["sac", "cql", "edac", "iql", "td3bc"])
def test_state_dict_keys_and_shapes_match_the_reference(algo):
    _ref()
    r, o = _build("offlinerlkit", algo).state_dict(), _build("offlinerlkit_b200", algo).state_dict()
    assert list(r.keys()) == list(o.keys())
    for k in r:
        assert tuple(r[k].shape) == tuple(o[k].shape) and r[k].dtype == o[k].dtype, k


def test_dynamics_model_state_dict_matches_the_reference():
    import importlib
    _ref()
    kw = dict(obs_dim=5, action_dim=3, hidden_dims=[8, 8], num_ensemble=3, num_elites=2, weight_decays=[1e-5, 2e-5, 3e-5],
              device="cpu")
    r = importlib.import_module("offlinerlkit.modules").EnsembleDynamicsModel(**kw).state_dict()
    o = importlib.import_module("offlinerlkit_b200.modules").EnsembleDynamicsModel(**kw).state_dict()
    assert list(r.keys()) == list(o.keys())
    for k in r:
        assert tuple(r[k].shape) == tuple(o[k].shape), k
