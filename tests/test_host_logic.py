"""CPU: the C-ABI library loads and exports every symbol of include/orlk_b200.h; host-side logic that needs no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "orlk_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(orlk_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from offlinerlkit_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _header_symbols()
    assert len(declared) >= 50
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/orlk_b200.h but not exported"
    bound = set(_lib.EXPORTS)
    assert set(declared) <= bound, sorted(set(declared) - bound)
    assert _lib.load().orlk_abi_version() == _lib.ABI_VERSION


def test_compute_entry_points_fail_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from offlinerlkit_b200 import _lib
    from offlinerlkit_b200.buffer import ReplayBuffer
    buf = ReplayBuffer(16, (3,), np.float32, 2, np.float32, device="cpu")
    buf.add_batch(np.zeros((4, 3), np.float32), np.zeros((4, 3), np.float32), np.zeros((4, 2), np.float32),
                  np.zeros((4, 1), np.float32), np.zeros((4, 1), np.float32))
    with pytest.raises(_lib.OrlkError):
        buf.sample(2)                     # no CPU gather fallback


def test_replay_buffer_host_api_matches_oracle():
    """Ring writes, load_dataset, normalize_obs, sample_all (buffer/buffer.py:52-115) on the host arrays."""
    from offlinerlkit_b200.buffer import ReplayBuffer
    from oracle import replay as oreplay
    rng = np.random.default_rng(0)
    buf = ReplayBuffer(50, (4,), np.float32, 2, np.float32, device="cpu")
    ref = {k: np.zeros_like(getattr(buf, k)) for k in oreplay.FIELDS}
    ptr = size = 0
    for _ in range(7):
        n = int(rng.integers(1, 30))
        new = {"observations": rng.standard_normal((n, 4), dtype=np.float32),
               "next_observations": rng.standard_normal((n, 4), dtype=np.float32),
               "actions": rng.standard_normal((n, 2), dtype=np.float32),
               "rewards": rng.standard_normal((n, 1), dtype=np.float32),
               "terminals": (rng.random((n, 1)) < 0.3).astype(np.float32)}
        buf.add_batch(new["observations"], new["next_observations"], new["actions"], new["rewards"], new["terminals"])
        ptr, size = oreplay.ring_write(ref, ptr, size, 50, new)
        assert (buf._ptr, buf._size) == (ptr, size)
        for k in oreplay.FIELDS:
            assert np.array_equal(getattr(buf, k), ref[k]), k
        covered = np.zeros(50, bool)          # the dirty ranges cover every row written since the last mirror sync
        for lo, hi in buf._dirty:
            covered[lo:hi] = True
        assert covered[:size].all() or size < 50
    o2, n2, mean, std = oreplay.normalize_obs(buf.observations.copy(), buf.next_observations.copy())
    m, s = buf.normalize_obs()
    assert np.array_equal(m, mean) and np.array_equal(s, std) and np.array_equal(buf.observations, o2)
    allb = buf.sample_all()
    assert allb["observations"].shape == (size, 4) and allb["rewards"].shape == (size, 1)


def test_gemm_descriptor_builder_tiles_and_splits():
    from offlinerlkit_b200 import _lib as L
    from offlinerlkit_b200.engine.core import Runtime
    for K, want, cfg in [(7936, 18, L.CFG_SMALL), (7936, 16, L.CFG_BIG), (256, 4, L.CFG_SMALL), (23, 3, L.CFG_BIG), (1, 1, 0)]:
        s = Runtime.effective_splits(K, want, cfg)
        BK = L.CFG_TILES[cfg][2]
        chunk = -(-(-(-K // max(1, want))) // BK) * BK
        assert s == -(-K // chunk) and 1 <= s <= max(1, want)
        assert (s - 1) * chunk < K <= s * chunk          # every split is non-empty
    lib = L.load()
    for K, want in [(7936, 16), (256, 1), (33, 5)]:
        s = lib.orlk_tc_effective_splits(K, want)
        slabs = -(-K // 32)
        per = -(-slabs // max(want, 1))
        assert s == -(-slabs // per)


def test_param_recipe_and_dataset_are_deterministic():
    from offlinerlkit_b200.synthetic import make_dataset, param_recipe
    a, b = make_dataset(100, 17, 6), make_dataset(100, 17, 6)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert a["terminals"].dtype == np.float32 and set(np.unique(a["terminals"])) <= {0.0, 1.0}
    p1 = param_recipe({"w": (4, 3), "b": (4,)}, 5)
    p2 = param_recipe({"w": (4, 3), "b": (4,)}, 5)
    assert np.array_equal(p1["w"], p2["w"]) and np.abs(p1["w"]).max() <= 1 / np.sqrt(3)


def test_facade_state_dict_layout():
    """state_dict keys of the facade policies are the reference's: 42 tensors for CQL."""
    from tests.gpu_common import build_policy
    from tests.helpers import Golden, initial_state
    for name in ["cql_small", "sac_small", "td3bc_small", "iql_small", "edac_small"]:
        m = Golden(name).meta
        pol = build_policy(m, "cpu")
        sd = pol.state_dict()
        ref = initial_state(m)            # key set and shapes of the reference's state_dict (tests/helpers.py)
        assert set(ref) == set(sd), (sorted(set(ref) - set(sd))[:5], sorted(set(sd) - set(ref))[:5])
        if name == "cql_small":
            assert len(sd) == 42
        for k, v in ref.items():
            assert tuple(sd[k].shape) == tuple(v.shape), k


def test_adam_group_bias_corrections_follow_torch():
    """The group table carries the bias corrections of the NEXT step (the kernels no longer evaluate pow() themselves):
    AdamGroup.refresh() must give what torch.optim.Adam computes for step t = step + 1 (torch/optim/adam.py, single-tensor
    path: bias_correction1 = 1 - beta1 ** step, bias_correction2_sqrt = sqrt(1 - beta2 ** step))."""
    import math
    from offlinerlkit_b200 import _lib as L
    g = L.AdamGroup()
    g.lr, g.beta1, g.beta2, g.eps = 3e-4, 0.9, 0.999, 1e-8
    for step in (0, 1, 7, 1000, 123456):
        g.step = step
        g.refresh()
        t = step + 1
        b1, b2 = float(g.beta1), float(g.beta2)          # the float32 values the kernels see
        assert g.bc1 == 1.0 - b1 ** t
        assert abs(g.bc2_sqrt - math.sqrt(1.0 - b2 ** t)) <= 1e-7
        assert ctypes.sizeof(L.AdamGroup) == 40
