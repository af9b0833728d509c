"""GPU: drop-in hazards at the ``policy.learn`` boundary -- host-side parameter writes after the engine exists, changing
batch sizes between calls, foreign batches, dict copies of a lazily gathered batch."""
import numpy as np
import pytest
import torch

from tests.helpers import Golden, initial_state

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cql_noise(B, N, A, seed):
    g = torch.Generator().manual_seed(seed)
    R = B * N
    return {"eps_actor": torch.randn(B, A, generator=g), "eps_next": torch.randn(B, A, generator=g),
            "rand_act": torch.rand(R, A, generator=g) * 2 - 1, "eps_pi": torch.randn(R, A, generator=g),
            "eps_pi_next": torch.randn(R, A, generator=g)}


def _batch(data, idx):
    from tests.helpers import FIELDS
    return {k: torch.from_numpy(data[k][idx]) for k in FIELDS}


def test_load_state_dict_after_first_learn_refreshes_derived_weights():
    """``ParamSet.refresh_wt`` must see ``load_state_dict`` (in-place ``param.copy_``) although the adopted parameters
    carry their own version counters.  cql_hc runs the 7936-row critic pass whose tensor-core dgrad reads the transposed
    weight copies WT.  Three engines with identical history (one step) get the same new parameters: (a) through
    load_state_dict alone, (b) through load_state_dict + an explicit invalidate(), (c) through ``p.data.copy_`` WITHOUT
    invalidate (no counter moves: WT stays stale -- the control that shows the test can see a stale copy).  After one
    more step (a) must equal (b) bit for bit and (c) must not."""
    from tests.gpu_common import build_policy, load_state
    g = Golden("cql_hc")
    m = g.meta
    data = g.dataset()
    st0 = initial_state(m)
    st1 = {k: (v * 1.25 + 0.01 if v.is_floating_point() else v) for k, v in st0.items()}
    rng = np.random.default_rng(0)
    idx = rng.integers(0, m["n_data"], size=(2, m["B"]))
    noise = [_cql_noise(m["B"], m["N"], m["A"], s) for s in (1, 2)]

    def run(how):
        pol = build_policy(m, DEV)
        load_state(pol, st0)
        pol.train()
        pol.learn(_batch(data, idx[0]), noise=noise[0])
        if how == "data":
            with torch.no_grad():
                for k, p in pol.state_dict().items():
                    if k in st1 and p.is_floating_point():
                        p.data.copy_(st1[k].to(DEV))
        else:
            load_state(pol, st1)                    # host-side write AFTER the engine adopted the parameters
            if how == "invalidate":
                pol._engine.invalidate()
        pol.learn(_batch(data, idx[1]), noise=noise[1])
        return {k: v.detach().clone() for k, v in pol.state_dict().items()}

    a, b, c = run("load"), run("invalidate"), run("data")
    for k in a:
        assert torch.equal(a[k], b[k]), f"stale derived weights after load_state_dict: {k}"
    assert any(not torch.equal(a[k], c[k]) for k in a if k.startswith("critic")), \
        "control failed: a stale transposed copy should have changed the critic update"


@pytest.mark.parametrize("name", ["cql_small", "sac_small", "iql_small", "td3bc_small", "edac_small"])
def test_batch_size_may_change_between_calls(name):
    """The reference's ``learn`` takes any batch size per call; the engine builds a sibling step graph over the SAME
    parameters / Adam state for every new size.  Checked against the CPU oracle fed the same batches and noise."""
    from tests.gpu_common import build_policy, load_state, make_oracle
    g = Golden(name)
    m = g.meta
    data = g.dataset()
    A, N = m["A"], m["hyper"].get("num_repeat_actions", 1)
    pol = build_policy(m, DEV)
    load_state(pol, initial_state(m))
    pol.train()
    ora = make_oracle(m)
    rng = np.random.default_rng(3)
    sizes = [m["B"], 24, m["B"], 8, 24]
    gen = torch.Generator().manual_seed(9)
    for t, B in enumerate(sizes):
        idx = rng.integers(0, m["n_data"], size=B)
        batch = _batch(data, idx)
        if m["algo"] == "cql":
            noise = _cql_noise(B, N, A, 100 + t)
        elif m["algo"] == "sac":
            noise = {"eps_next": torch.randn(B, A, generator=gen), "eps_actor": torch.randn(B, A, generator=gen)}
        elif m["algo"] == "edac":
            noise = {"eps_actor": torch.randn(B, A, generator=gen), "eps_next": torch.randn(B, A, generator=gen)}
        elif m["algo"] == "td3bc":
            noise = {"eps_target": torch.randn(B, A, generator=gen)}
        else:
            noise = None
        out = pol.learn({k: v.clone() for k, v in batch.items()}, noise=noise) if noise is not None else pol.learn(batch)
        ref = ora.step(batch, noise) if noise is not None else ora.step(batch)
        assert out.keys() == ref.keys()
        for k in ref:
            assert out[k] == pytest.approx(ref[k], rel=1e-4, abs=1e-4), (t, B, k, out[k], ref[k])
    assert len(pol._engines) == 3
    sd = pol.state_dict()
    for k, v in ora.state_dict().items():
        if v.is_floating_point() and "saved_" not in k:
            lr = 3e-4
            assert (sd[k].cpu() - v).abs().max().item() <= 1e-4 * v.abs().max().item() + 2.5 * lr * len(sizes), k


def test_foreign_batch_is_not_overwritten_by_a_pending_draw():
    """A sampled-but-unused draw of the bound buffer must not gather over a foreign batch at the head of the graph."""
    from tests.gpu_common import build_policy, load_state, make_buffer
    g = Golden("sac_small")
    m = g.meta
    data = g.dataset()
    noise = g.noise(0)

    def run(with_pending):
        pol = build_policy(m, DEV)
        load_state(pol, initial_state(m))
        pol.train()
        buf, _ = make_buffer(g, DEV)
        np.random.seed(0)
        pol.learn(buf.sample(m["B"]), noise=noise)          # binds the engine to the buffer's staging memory
        if with_pending:
            buf.sample(m["B"])                              # drawn, never read: its gather is still pending
        foreign = {k: v.to(DEV) for k, v in g.batch(1, data).items()}
        return pol.learn(foreign, noise=g.noise(1))

    a, b = run(True), run(False)
    assert a == b, (a, b)


def test_batch_dict_copies_materialise_the_rows():
    """``{**batch}``, ``dict(batch)`` and ``batch.copy()`` must hand out the SAMPLED rows, not stale staging memory."""
    from tests.gpu_common import make_buffer
    g = Golden("sac_small")
    m = g.meta
    buf, data = make_buffer(g, DEV)
    np.random.seed(4)
    for mode in ("star", "dict", "copy", "or"):
        b = buf.sample(m["B"])
        assert b.token.pending
        c = {"star": lambda: {**b}, "dict": lambda: dict(b), "copy": lambda: b.copy(), "or": lambda: b | {}}[mode]()
        idx = b.indices.cpu().numpy()
        assert np.array_equal(c["observations"].cpu().numpy(), data["observations"][idx]), mode
        assert np.array_equal(c["rewards"].cpu().numpy().reshape(-1), data["rewards"][idx].reshape(-1)), mode


@pytest.mark.parametrize("name", ["cql_small", "sac_small", "iql_small", "td3bc_small", "edac_small", "cql_hc"])
def test_learn_many_equals_single_steps(name):
    """``policy.learn_many(buffer, K, B)`` (K steps behind ONE host synchronisation) == K calls of
    ``policy.learn(buffer.sample(B))``: the same np.random index stream, bit-identical loss dicts and parameters."""
    from tests.gpu_common import build_policy, load_state, make_buffer
    g = Golden(name)
    m = g.meta
    K = 7

    def make():
        pol = build_policy(m, DEV)
        load_state(pol, initial_state(m))
        pol.train()
        buf, _ = make_buffer(g, DEV)
        pol.engine(m["B"]).seed = 11
        np.random.seed(21)
        return pol, buf

    pa, ba = make()
    single = [pa.learn(ba.sample(m["B"])) for _ in range(K)]
    end_a = np.random.randint(0, 1 << 30)
    pb, bb = make()
    many = pb.learn_many(bb, 3, m["B"]) + pb.learn_many(bb, K - 3, m["B"])
    end_b = np.random.randint(0, 1 << 30)
    assert end_a == end_b, "learn_many must consume np.random exactly like K sample() calls"
    assert many == single
    for (k, x), (_, y) in zip(pa.state_dict().items(), pb.state_dict().items()):
        assert torch.equal(x, y), k
