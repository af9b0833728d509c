"""Multi-GPU plumbing for the replica modes.

The CQL / IQL / TD3+BC / SAC gradient step does not shard: N GPUs run N independent seeds (or sweep members), one
process per GPU, with NO collective on the data path.  The only cross-rank traffic is bookkeeping: a barrier around
timed regions and a MAX / SUM reduction of a few scalars.  ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests) is
used for exactly that.
"""
import os
from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


def env_rank() -> tuple:
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


class _StdoutToStderr:
    """NCCL announces its version on stdout when the first communicator is created; callers that promise ONE JSON line
    on stdout route the file descriptor to stderr for the duration of the rendezvous."""

    def __enter__(self):
        import sys
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        import sys
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)


def init(backend: str, device: torch.device = None) -> bool:
    """Join the process group described by the torchrun environment; returns False for a single process."""
    rank, world, _ = env_rank()
    if world <= 1:
        return False
    if not dist.is_initialized():
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        with _StdoutToStderr():
            dist.init_process_group(backend, **kw)
            t = torch.zeros(1, device=device if backend == "nccl" else "cpu")
            dist.all_reduce(t)                  # creates the communicator now, inside the redirected region
            if backend == "nccl":
                torch.cuda.synchronize(device)
    return True


def seed_for_rank(base_seed: int, rank: int) -> int:
    """Replica r trains seed base+r: its result must equal a single-GPU run with that seed."""
    return base_seed + rank


def reduce_scalars(values: Sequence[float], op: str, device="cpu") -> List[float]:
    """MAX or SUM of a few python floats over all ranks (identity for a single process)."""
    if not (dist.is_available() and dist.is_initialized()):
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return t.tolist()


def aggregate_rate(steps_per_rank: int, elapsed_ms: float, device="cpu") -> float:
    """Whole-job steps/s of N independent replicas timed together: all steps / the slowest rank's time."""
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    (t_max,) = reduce_scalars([elapsed_ms], "max", device)
    return world * steps_per_rank / (t_max * 1e-3)


def partition_members(n_members: int, world: int) -> List[List[int]]:
    """Contiguous, as-even-as-possible split of ensemble members over ranks (E=10 over 4 ranks -> 3/3/2/2)."""
    base, extra = divmod(n_members, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append(list(range(start, start + n)))
        start += n
    return out


def gather_member_blocks(local: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """Host-level form of the member-sharded exchange (engine/edac_sharded.py does the same with ``orlk_compact_blocks``
    inside its step graphs): every rank contributes ``local`` [counts[rank], ...]; all ranks get the dense
    [sum(counts), ...] tensor in rank order.  One equal-block all-gather, blocks padded to the largest slice."""
    world = dist.get_world_size(group)
    e_max = max(counts)
    per = local[0].numel() if local.shape[0] else int(torch.tensor(local.shape[1:]).prod())
    send = torch.zeros(e_max * per, dtype=local.dtype, device=local.device)
    send[:local.numel()].copy_(local.reshape(-1))
    recv = torch.empty(world * e_max * per, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    parts = [recv[r * e_max * per:r * e_max * per + c * per] for r, c in enumerate(counts)]
    return torch.cat(parts).view((sum(counts),) + tuple(local.shape[1:]))


# --------------------------------------------------------------------------------------------------------------
# State-sharded model rollouts (MOPO rollouts, variant B)
# --------------------------------------------------------------------------------------------------------------
def shard_rows(n_rows: int, rank: int, world: int) -> tuple:
    """[lo, hi) of the contiguous, as-even-as-possible share of ``n_rows`` for ``rank`` (50 000 over 8 -> 6 250 each)."""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _gather_ragged(arr, device) -> "np.ndarray":
    """All-gather of per-rank arrays that differ in their first dimension (rollouts drop terminated states), in rank
    order.  One size exchange + one padded all-gather per array."""
    import numpy as np
    world = dist.get_world_size()
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(device)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    n_max = max(sizes + [1])
    pad = torch.zeros((n_max,) + tuple(t.shape[1:]), dtype=t.dtype, device=device)
    pad[:t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], 0).cpu().numpy()


def _gather_packed(local, device, to_host: bool = True):
    """The same gather for a dict of DEVICE tensors with a common first dimension: the columns are packed into one
    [n, W] fp32 matrix, so the whole exchange is one size all-gather, one padded all-gather and one device-to-host copy per array
    (uint8 / bool columns survive the round trip through fp32 exactly)."""
    import numpy as np
    world = dist.get_world_size()
    keys = list(local)
    n_loc = int(local[keys[0]].shape[0])
    cols = [local[k].reshape(n_loc, -1) for k in keys]
    widths = [int(c.shape[1]) for c in cols]
    pack = torch.cat([c.to(torch.float32) for c in cols], 1).contiguous()
    n = torch.tensor([n_loc], dtype=torch.int64, device=pack.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(v.item()) for v in sizes]
    pad = torch.zeros((max(sizes + [1]), pack.shape[1]), dtype=torch.float32, device=pack.device)
    pad[:n_loc] = pack
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    full = torch.cat([p[:v] for p, v in zip(parts, sizes)], 0)
    if not to_host:             # device tensors: what ReplayBuffer.add_batch packs straight into its device row table
        out, c0 = {}, 0
        for k, w in zip(keys, widths):
            a = full[:, c0:c0 + w]
            out[k] = (a != 0).to(local[k].dtype) if local[k].dtype in (torch.uint8, torch.bool) else a.contiguous()
            c0 += w
        return out
    # split on the device and copy each array out on its own: host blocks of <= 32 MB are recycled by the allocator,
    # one 42 MB block would be mmap-ed and page-faulted afresh on every call
    out, c0 = {}, 0
    for k, w in zip(keys, widths):
        a = full[:, c0:c0 + w]
        if local[k].dtype in (torch.uint8, torch.bool):
            a = a != 0
        out[k] = a.contiguous().cpu().numpy()
        c0 += w
    return out


def rollout_state_sharded(rollout_fn, init_obss, rollout_length: int, device="cpu", device_out: bool = False,
                          device_result: bool = False):
    """``MOPOPolicy.rollout`` (policy/model_based/mopo.py:45-79) with the start states split over the ranks.

    Every rank holds the full dynamics ensemble and the actor (replicated after training), imagines the whole horizon
    for its own contiguous share of ``init_obss`` with no communication, and the resulting transitions are all-gathered
    in rank order, so every rank ends up with the same fake-buffer batch.  ``rollout_fn(obs, length) -> (dict, info)`` is
    the single-GPU rollout.  The random streams differ per rank, so the result is distributed like - not bit-equal to -
    a single-GPU rollout of all states.  Single process: plain call.
    ``device_out``: ``rollout_fn`` takes ``device_out=True`` and then returns device tensors (MOPOPolicy / COMBOPolicy
    do); the transitions are exchanged straight from device memory and reach the host once.
    ``device_result`` (with ``device_out``): return the gathered transitions as DEVICE tensors -- nothing goes through the
    host; ``ReplayBuffer.add_batch`` packs them straight into the fake buffer's device row table."""
    import numpy as np
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        if device_out and device_result:
            return rollout_fn(init_obss, rollout_length, device_out=True)
        return rollout_fn(init_obss, rollout_length)
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_rows(len(init_obss), rank, world)
    if device_out:
        local, info = rollout_fn(init_obss[lo:hi], rollout_length, device_out=True)
        out = _gather_packed(local, device, to_host=not device_result)
    else:
        local, info = rollout_fn(init_obss[lo:hi], rollout_length)
        out = {k: _gather_ragged(v, device) for k, v in local.items()}
    n_local = float(len(next(iter(local.values())))) if local else 0.0
    n_tot, r_sum = reduce_scalars([n_local, float(info.get("reward_mean", 0.0)) * n_local], "sum", device)
    merged = dict(info)
    merged["num_transitions"] = int(n_tot)
    merged["reward_mean"] = r_sum / max(n_tot, 1.0)
    return out, merged
