"""Shared helpers for the tests: golden loading and deterministic model/batch reconstruction."""
import json
import os
from typing import Dict

import numpy as np
import torch

from offlinerlkit_b200.synthetic import make_dataset, param_recipe

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = ("observations", "actions", "next_observations", "terminals", "rewards")


class Golden:
    def __init__(self, name: str):
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
        self.meta = json.loads(str(self.z["meta"]))

    def group(self, prefix: str) -> Dict[str, np.ndarray]:
        pre = prefix + "|"
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}

    def __getitem__(self, k):
        return self.z[k]

    def dataset(self):
        m = self.meta
        d = make_dataset(m["n_data"], m["O"], m["A"], seed=m["data_seed"])
        d["rewards"] = d["rewards"].reshape(-1, 1)
        d["terminals"] = d["terminals"].reshape(-1, 1)
        return d

    def fake_dataset(self):
        """COMBO goldens: the second synthetic dataset the model-buffer rows are drawn from."""
        m = self.meta
        d = make_dataset(m["n_data"], m["O"], m["A"], seed=m["fake_data_seed"])
        d["rewards"] = d["rewards"].reshape(-1, 1)
        d["terminals"] = d["terminals"].reshape(-1, 1)
        return d

    def batch(self, t: int, data=None) -> Dict[str, torch.Tensor]:
        data = data or self.dataset()
        idx = self.z["idx"][t]
        real = {k: torch.from_numpy(data[k][idx]) for k in FIELDS}
        if self.meta["algo"] != "combo":
            return real
        fdata, fidx = self.fake_dataset(), self.z["fake_idx"][t]
        return {"real": real, "fake": {k: torch.from_numpy(fdata[k][fidx]) for k in FIELDS}}

    def noise(self, t: int) -> Dict[str, torch.Tensor]:
        return {k: torch.from_numpy(v) for k, v in self.group(f"noise{t}").items()}

    def losses(self, t: int) -> Dict[str, float]:
        return {k: float(v) for k, v in self.group(f"loss{t}").items()}


def mlp_shapes(prefix, in_dim, hidden):
    out, d = {}, in_dim
    for i, h in enumerate(hidden):
        out[f"{prefix}.model.{2 * i}.weight"] = (h, d)
        out[f"{prefix}.model.{2 * i}.bias"] = (h,)
        d = h
    return out


def critic_shapes(in_dim, hidden):
    s = mlp_shapes("backbone", in_dim, hidden)
    s["last.weight"] = (1, hidden[-1])
    s["last.bias"] = (1,)
    return s


def actorprob_shapes(O, A, hidden, conditioned_sigma=True):
    s = mlp_shapes("backbone", O, hidden)
    if not conditioned_sigma:
        s["dist_net.sigma_param"] = (A, 1)
    s["dist_net.mu.weight"] = (A, hidden[-1])
    s["dist_net.mu.bias"] = (A,)
    if conditioned_sigma:
        s["dist_net.sigma.weight"] = (A, hidden[-1])
        s["dist_net.sigma.bias"] = (A,)
    return s


def det_actor_shapes(O, A, hidden):
    s = mlp_shapes("backbone", O, hidden)
    s["last.weight"] = (A, hidden[-1])
    s["last.bias"] = (A,)
    return s


def ensemble_critic_shapes(O, A, hidden, E):
    out, d = {}, O + A
    for i, h in enumerate(list(hidden) + [1]):
        for nm, shp in (("weight", (E, d, h)), ("bias", (E, 1, h)), ("saved_weight", (E, d, h)), ("saved_bias", (E, 1, h))):
            out[f"model.{2 * i}.{nm}"] = shp
        d = h
    return out


def dynamics_shapes(O, A, hidden, E):
    out, d = {}, O + A
    names = [f"backbones.{i}" for i in range(len(hidden))] + ["output_layer"]
    dims = list(hidden) + [2 * (O + 1)]
    head = {"max_logvar": (O + 1,), "min_logvar": (O + 1,)}
    for n, h in zip(names, dims):
        for nm, shp in (("weight", (E, d, h)), ("bias", (E, 1, h)), ("saved_weight", (E, d, h)), ("saved_bias", (E, 1, h))):
            out[f"{n}.{nm}"] = shp
        d = h
    return {**head, **out}


def recipe_state(shapes: Dict[str, tuple], seed: int, prefix: str) -> Dict[str, torch.Tensor]:
    return {f"{prefix}.{k}": torch.from_numpy(v) for k, v in param_recipe(shapes, seed).items()}


def initial_state(meta) -> Dict[str, torch.Tensor]:
    """Rebuild the pre-step state_dict of a golden run from its parameter seeds (see make_golden.overwrite_params)."""
    algo, O, A, hid = meta["algo"], meta["O"], meta["A"], meta["hidden"]
    ps = meta.get("param_seeds", {})
    st = {}
    if algo in ("cql", "sac", "combo"):
        st.update(recipe_state(actorprob_shapes(O, A, hid), ps["actor"], "actor"))
        for c in ("critic1", "critic2"):
            cs = recipe_state(critic_shapes(O + A, hid), ps[c], c)
            st.update(cs)
            st.update({k.replace(c + ".", c + "_old.", 1): v.clone() for k, v in cs.items()})
    elif algo == "edac":
        st.update(recipe_state(actorprob_shapes(O, A, hid), ps["actor"], "actor"))
        cs = recipe_state(ensemble_critic_shapes(O, A, hid, meta["E"]), ps["critics"], "critics")
        st.update(cs)
        st.update({k.replace("critics.", "critics_old.", 1): v.clone() for k, v in cs.items()})
    elif algo == "iql":
        st.update(recipe_state(actorprob_shapes(O, A, hid, conditioned_sigma=False), ps["actor"], "actor"))
        for c in ("critic_q1", "critic_q2"):
            cs = recipe_state(critic_shapes(O + A, hid), ps[c], c)
            st.update(cs)
            st.update({k.replace(c + ".", c + "_old.", 1): v.clone() for k, v in cs.items()})
        st.update(recipe_state(critic_shapes(O, hid), ps["critic_v"], "critic_v"))
    elif algo == "td3bc":
        for c, shp in (("actor", det_actor_shapes(O, A, hid)), ("critic1", critic_shapes(O + A, hid)),
                       ("critic2", critic_shapes(O + A, hid))):
            cs = recipe_state(shp, ps[c], c)
            st.update(cs)
            st.update({k.replace(c + ".", c + "_old.", 1): v.clone() for k, v in cs.items()})
    elif algo in ("dynamics", "dynamics_sample_next"):
        raw = param_recipe(dynamics_shapes(O, A, hid, meta["E"]), meta["param_seed"])
        st = {k: torch.from_numpy(v) for k, v in raw.items()}
        st["max_logvar"] = torch.full((O + 1,), 0.5)
        st["min_logvar"] = torch.full((O + 1,), -10.0)
        st["elites"] = torch.arange(meta["n_elites"])
    else:
        raise KeyError(algo)
    return st


def tensor_stats(t: torch.Tensor) -> np.ndarray:
    x = t.detach().double().flatten().cpu().numpy()
    stride = max(1, x.size // 64)
    return np.concatenate([[x.sum(), np.abs(x).sum(), np.sqrt((x * x).sum())], x[::stride][:64]])


def rel_err(a, b) -> float:
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


def assert_stats_close(state: Dict[str, torch.Tensor], stats: Dict[str, np.ndarray], tol: float, lr_atol: float = 0.0,
                       skip=("saved_",)):
    """Compare a state dict with the golden per-tensor fingerprints.

    The three norms are compared relatively; the 64 sampled elements use
    ``rtol=tol`` plus ``atol=lr_atol`` (Adam's sign sensitivity, SURVEY.md section 7 "hard parts": the first Adam
    steps move every element by about +-lr whatever the size of its gradient, so an element whose gradient is a
    rounding-level cancellation may legitimately move the other way).  The norms get room for three such elements.
    """
    for k, ref in stats.items():
        if any(s in k for s in skip):
            continue
        got = tensor_stats(state[k])
        assert got.shape == ref.shape, k
        np.testing.assert_allclose(got[1:3], ref[1:3], rtol=tol, atol=3 * lr_atol, err_msg=k)
        np.testing.assert_allclose(got[3:], ref[3:], rtol=tol, atol=lr_atol + 1e-7, err_msg=k)


# ---------------------------------------------------------------------------------------------- gradients
def grad_stats(t) -> np.ndarray:
    """[sum, abs-sum, L2, max-abs, 64 strided elements] -- same fingerprint as make_golden.grad_stats."""
    x = (t.detach().double().flatten().cpu().numpy() if torch.is_tensor(t) else np.asarray(t, np.float64).ravel())
    stride = max(1, x.size // 64)
    return np.concatenate([[x.sum(), np.abs(x).sum(), np.sqrt((x * x).sum()), np.abs(x).max()], x[::stride][:64]])


def assert_grad_stats_close(grads: Dict[str, torch.Tensor], stats: Dict[str, np.ndarray], tol: float, what: str = "",
                            elements: bool = True):
    """Gradients vs the reference's fingerprints (``gradstats{t}`` of a golden file): the same tensors must be present,
    abs-sum / L2 / max-abs relative to ``tol``, the sampled elements with ``atol = tol * max|g|``."""
    assert set(grads) == set(stats), (what, sorted(set(grads) ^ set(stats)))
    for k, ref in stats.items():
        got = grad_stats(grads[k])
        assert got.shape == ref.shape, (what, k)
        scale = max(ref[3], 1e-30)
        t = 4 * tol if ref.shape[0] == 5 else tol       # one-element tensors: a heavily cancelling scalar sum (tests/kinks.py)
        np.testing.assert_allclose(got[1:4], ref[1:4], rtol=t, atol=t * scale, err_msg=f"{what} {k} norms")
        if elements:
            np.testing.assert_allclose(got[4:], ref[4:], rtol=t, atol=t * scale, err_msg=f"{what} {k} elements")


def assert_grads_close(got: Dict[str, torch.Tensor], ref: Dict[str, torch.Tensor], tol: float, what: str = ""):
    """Full gradient tensors: relative L2 error <= tol per tensor and element-wise rtol = tol, atol = tol * max|g|."""
    assert set(got) == set(ref), (what, sorted(set(got) ^ set(ref)))
    for k, r in ref.items():
        a = got[k].detach().double().cpu().reshape(-1).numpy()
        b = r.detach().double().cpu().reshape(-1).numpy()
        assert a.shape == b.shape, (what, k, a.shape, b.shape)
        l2 = np.sqrt(((a - b) ** 2).sum()) / max(np.sqrt((b * b).sum()), 1e-30)
        assert l2 <= tol, (what, k, "rel-L2", l2)
        np.testing.assert_allclose(a, b, rtol=tol, atol=tol * max(np.abs(b).max(), 1e-30), err_msg=f"{what} {k}")


# ---------------------------------------------------------------------------------------------- config-5 rollout fixtures
def cfg5_setup(meta):
    """Inputs of a ``rollout_cfg5_*`` fixture rebuilt from its recipes (make_golden.gen_rollout_cfg5):
    (dynamics state dict, actor state dict, scaler mu, scaler std, start states)."""
    O, A = meta["O"], meta["A"]
    raw = param_recipe(dynamics_shapes(O, A, meta["dyn_hidden"], meta["E"]), meta["dyn_seed"])
    dyn = {k: torch.from_numpy(v) for k, v in raw.items()}
    dyn["max_logvar"] = torch.full((O + 1,), 0.5)
    dyn["min_logvar"] = torch.full((O + 1,), -10.0)
    dyn["output_layer.weight"].mul_(0.1)
    dyn["output_layer.bias"][..., O + 1:] = -6.0
    dyn["elites"] = torch.tensor(meta["elites"])
    actor = recipe_state(actorprob_shapes(O, A, meta["hidden"]), meta["param_seeds"]["actor"], "actor")
    data = make_dataset(meta["S"], O, A, seed=meta["data_seed"])
    x = np.concatenate([data["observations"], data["actions"]], axis=-1)
    mu = np.mean(x, axis=0, keepdims=True)
    std = np.std(x, axis=0, keepdims=True)
    std[std < 1e-12] = 1.0
    init = data["observations"].copy()
    if meta["term"] == "walker2d":
        init[:, 0] = 1.4 + 0.35 * init[:, 0]
        init[:, 1] = 0.3 * init[:, 1]
    return dyn, actor, mu, std, init


def array_stats(x, rows: int = 64) -> np.ndarray:
    x = np.asarray(x)
    x2 = x.reshape(len(x), -1).astype(np.float64)
    stride = max(1, len(x2) // rows)
    return np.concatenate([[x2.sum(), np.abs(x2).sum(), np.sqrt((x2 * x2).sum())], x2[::stride][:rows].ravel()])
