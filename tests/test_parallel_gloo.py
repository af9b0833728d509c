"""CPU, world_size 2 over gloo: the replica-mode bookkeeping (timing reduction, seeds, member partition)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank: int, world: int, port: int, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from offlinerlkit_b200 import parallel
    assert parallel.init("gloo")
    # rank 1 is slower: the whole-job rate is bounded by it
    elapsed = 100.0 if rank == 0 else 250.0
    rate = parallel.aggregate_rate(steps_per_rank=500, elapsed_ms=elapsed)
    sums = parallel.reduce_scalars([float(rank + 1), 2.0], "sum")
    import torch.distributed as dist
    # replica determinism contract: every rank derives its own seed; same seed -> same index stream on any rank
    import numpy as np
    np.random.seed(parallel.seed_for_rank(7, rank))
    idx = torch.from_numpy(np.random.randint(0, 1000, size=8))
    gathered = [torch.zeros_like(idx) for _ in range(world)]
    dist.all_gather(gathered, idx)
    q.put((rank, rate, sums, [g.tolist() for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


def test_replica_bookkeeping_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import numpy as np
    for rank, rate, sums, gathered in res:
        assert rate == pytest.approx(2 * 500 / 0.250)            # all ranks agree: total steps / slowest rank
        assert sums == [3.0, 4.0]
        for r in range(2):                                        # rank r's stream == a single-process run with seed 7+r
            np.random.seed(7 + r)
            assert gathered[r] == np.random.randint(0, 1000, size=8).tolist()
    assert res[0][3][0] != res[0][3][1]


def test_member_partition():
    from offlinerlkit_b200.parallel import partition_members
    assert [len(p) for p in partition_members(10, 4)] == [3, 3, 2, 2]
    assert [len(p) for p in partition_members(10, 8)] == [2, 2, 1, 1, 1, 1, 1, 1]
    assert sum(partition_members(7, 8), []) == list(range(7)) and partition_members(7, 8)[7] == []


def _fake_rollout(obs, length):
    """Stand-in for MOPOPolicy.rollout with its ragged output: every step keeps the states whose first coordinate is
    below a threshold that shrinks with the step (deterministic, so the sharded result can be checked exactly)."""
    import numpy as np
    rows = {"obss": [], "next_obss": [], "actions": [], "rewards": [], "terminals": []}
    cur = obs
    for t in range(length):
        nxt = cur * 0.5
        rows["obss"].append(cur)
        rows["next_obss"].append(nxt)
        rows["actions"].append(cur[:, :2] + t)
        rows["rewards"].append(cur[:, :1] * 2.0)
        rows["terminals"].append((cur[:, :1] > 0.5).astype(np.float32))
        cur = nxt[nxt[:, 0] <= 0.5 / (t + 1)]
        if len(cur) == 0:
            break
    out = {k: np.concatenate(v, 0) for k, v in rows.items()}
    return out, {"num_transitions": len(out["obss"]), "reward_mean": float(out["rewards"].mean())}


def _fake_rollout_dev(obs, length: int, device_out: bool = False):
    """The same stand-in with the ``device_out`` contract of MOPOPolicy.rollout: torch tensors, uint8 terminals."""
    import numpy as np
    import torch
    out, info = _fake_rollout(obs, length)
    if not device_out:
        return out, info
    dev = {k: torch.from_numpy(v) for k, v in out.items()}
    dev["terminals"] = dev["terminals"].to(torch.uint8)
    return dev, info


def _rollout_worker(rank: int, world: int, port: int, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import numpy as np
    from offlinerlkit_b200 import parallel
    import torch.distributed as dist
    assert parallel.init("gloo")
    obs = np.random.default_rng(3).random((101, 4), dtype=np.float32)       # 101 rows: uneven 51 / 50 split
    out, info = parallel.rollout_state_sharded(_fake_rollout, obs, 3)
    # the packed exchange of device tensors (one all-gather for the five arrays) must give the same batch
    out2, info2 = parallel.rollout_state_sharded(_fake_rollout_dev, obs, 3, device_out=True)
    assert out2["terminals"].dtype == np.bool_ and info2 == info
    for k, v in out.items():
        assert np.array_equal(np.asarray(out2[k], dtype=np.float32).reshape(v.shape), v), k
    q.put((rank, {k: v.tolist() for k, v in out.items()}, info))
    dist.barrier()
    dist.destroy_process_group()


def test_state_sharded_rollout_world2():
    """Two ranks imagine disjoint shares of the start states and all-gather ragged transition arrays: both ranks end up
    with rank 0's transitions followed by rank 1's, i.e. what the two single-process rollouts of the shares produce."""
    import numpy as np
    from offlinerlkit_b200.parallel import shard_rows
    assert [shard_rows(50000, r, 8) for r in (0, 7)] == [(0, 6250), (43750, 50000)]
    assert [shard_rows(101, r, 2) for r in (0, 1)] == [(0, 51), (51, 101)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rollout_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    obs = np.random.default_rng(3).random((101, 4), dtype=np.float32)
    parts = [_fake_rollout(obs[lo:hi], 3) for lo, hi in ((0, 51), (51, 101))]
    expect = {k: np.concatenate([p[0][k] for p in parts], 0) for k in parts[0][0]}
    n = sum(p[1]["num_transitions"] for p in parts)
    rmean = sum(p[1]["reward_mean"] * p[1]["num_transitions"] for p in parts) / n
    for rank, out, info in res:
        for k, v in expect.items():
            assert np.array_equal(np.asarray(out[k], dtype=np.float32), v), (rank, k)
        assert info["num_transitions"] == n and info["reward_mean"] == pytest.approx(rmean, rel=1e-6)


def _shard_worker(rank: int, world: int, port: int, q):
    """Member-sharded EDAC exchanges (engine/edac_sharded.py) at the host level: every rank holds a slice of the critics'
    outputs; after the padded all-gather the ensemble-wide reductions must equal the unsharded ones."""
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from offlinerlkit_b200 import parallel
    assert parallel.init("gloo")
    import torch.distributed as dist
    E, B, A = 5, 64, 3                              # 5 critics over 2 ranks: 3 / 2 (an uneven split)
    g = torch.Generator().manual_seed(0)
    q_full = torch.randn(E, B, generator=g)
    tq_full = torch.randn(E, B, generator=g)
    gin_full = torch.randn(E, B, A, generator=g)
    parts = parallel.partition_members(E, world)
    counts = [len(p) for p in parts]
    e0, e1 = parts[rank][0], parts[rank][-1] + 1
    q_all = parallel.gather_member_blocks(q_full[e0:e1].contiguous(), counts)
    tq_all = parallel.gather_member_blocks(tq_full[e0:e1].contiguous(), counts)
    gin_all = parallel.gather_member_blocks(gin_full[e0:e1].contiguous(), counts)
    ghat = gin_all / (gin_all.norm(dim=2, keepdim=True) + 1e-10)
    q.put((rank, torch.equal(q_all, q_full), q_all.argmin(0).tolist(), tq_all.min(0).values.tolist(),
           ghat.sum(0).tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_member_sharded_exchange_world2():
    """X1 / X3 of the sharded EDAC step: gathered == unsharded bit for bit, hence argmin over critics (actor loss,
    edac.py:96-102) and min over target critics (:124-131) exact, S = sum_e ghat_e (:136-149) to 1e-6."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    E, B, A = 5, 64, 3
    g = torch.Generator().manual_seed(0)
    q_full, tq_full, gin_full = torch.randn(E, B, generator=g), torch.randn(E, B, generator=g), torch.randn(E, B, A, generator=g)
    S = (gin_full / (gin_full.norm(dim=2, keepdim=True) + 1e-10)).sum(0)
    for rank, same, argmin, tmin, s in res:
        assert same
        assert argmin == q_full.argmin(0).tolist()
        assert tmin == tq_full.min(0).values.tolist()
        assert torch.allclose(torch.tensor(s), S, atol=1e-6, rtol=0)
