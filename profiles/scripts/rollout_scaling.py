"""MOPO model rollouts, state-sharded over the ranks (SURVEY section 8e, config 5 shape: E=7 members, hidden 200x4,
50 000 start states, horizon 5, SAC actor 256x2, halfcheetah termination).  One process per GPU:
    python profiles/scripts/rollout_scaling.py                               # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        profiles/scripts/rollout_scaling.py
Prints one JSON line on rank 0: imagined transitions per second for the whole job (device time, max over ranks)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from offlinerlkit_b200 import parallel


def main():
    rank, world, local = parallel.env_rank()
    dev = f"cuda:{local}"
    torch.cuda.set_device(local)
    dist_on = parallel.init("nccl", torch.device(dev))
    from offlinerlkit_b200.dynamics import EnsembleDynamics
    from offlinerlkit_b200.modules import ActorProb, Critic, EnsembleDynamicsModel, TanhDiagGaussian
    from offlinerlkit_b200.nets import MLP
    from offlinerlkit_b200.policy import MOPOPolicy
    from offlinerlkit_b200.utils import termination_fns as T
    from offlinerlkit_b200.utils.scaler import StandardScaler
    O, A, S, H = 17, 6, 50_000, 5
    torch.manual_seed(0)                                  # same (replicated) models on every rank
    np.random.seed(0)
    model = EnsembleDynamicsModel(O, A, [200, 200, 200, 200], num_ensemble=7, num_elites=5,
                                  weight_decays=[2.5e-5, 5e-5, 7.5e-5, 7.5e-5, 1e-4], device=dev)
    with torch.no_grad():                                 # tame the random model so that states stay bounded
        sd = model.state_dict()
        for k, v in sd.items():
            if "backbones.3" in k or "output" in k:
                v.mul_(0.1)
    dyn = EnsembleDynamics(model, torch.optim.Adam(model.parameters(), lr=1e-3),
                           StandardScaler(np.zeros((1, O + A), np.float32), np.ones((1, O + A), np.float32)),
                           T.termination_fn_halfcheetah, penalty_coef=0.5)
    assert dyn.rng == "device"                            # the default since round 2
    bb = MLP(O, [256, 256])
    actor = ActorProb(bb, TanhDiagGaussian(bb.output_dim, A, unbounded=True, conditioned_sigma=True), dev)
    c1, c2 = Critic(MLP(O + A, [256, 256]), dev), Critic(MLP(O + A, [256, 256]), dev)
    adam = lambda m: torch.optim.Adam(m.parameters(), lr=1e-4)
    pol = MOPOPolicy(dyn, actor, c1, c2, adam(actor), adam(c1), adam(c2), alpha=0.2)
    init = np.random.default_rng(1).standard_normal((S, O), dtype=np.float32)

    to_host = os.environ.get("ROLLOUT_TO_HOST", "0") == "1"

    def run():
        # default: the gathered transitions stay on the device (what ReplayBuffer.add_batch takes); ROLLOUT_TO_HOST=1
        # adds the export to NumPy arrays in the reference's format
        if to_host and not dist_on:
            return pol.rollout(init, H)
        return parallel.rollout_state_sharded(pol.rollout, init, H, device=dev, device_out=True, device_result=not to_host)

    for _ in range(3):
        out, info = run()
    if dist_on:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    K = 10
    t0 = time.perf_counter()
    for _ in range(K):
        out, info = run()
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / K
    (ms_max,) = parallel.reduce_scalars([ms], "max", dev)
    if rank == 0:
        print(json.dumps({"metric": "MOPO imagined transitions/s (state-sharded rollout, all-gather included)",
                          "value": info["num_transitions"] / (ms_max * 1e-3), "unit": "transitions/s", "n_gpus": world,
                          "ms_per_rollout": ms_max, "transitions_per_rollout": info["num_transitions"], "scaling": "strong",
                          "last_call_split_ms": getattr(pol._roll, "last_timing", None), "result": "host numpy" if to_host else "device tensors",
                          "config": {"workload": "mopo_rollout E7 hidden200x4 S50000 H5 hc"}}))
    if dist_on:
        torch.distributed.barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
