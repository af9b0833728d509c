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
