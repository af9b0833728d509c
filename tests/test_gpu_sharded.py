"""GPU: member-sharded EDAC (BASELINE.json configs[2]) with the ranks emulated on one device -- every rank engine owns a
slice of the critics, the three exchanges of engine/edac_sharded.py are device copies -- vs the golden run of the
reference (which is the unsharded computation) and vs the unsharded engine."""
import numpy as np
import pytest
import torch

from tests.helpers import Golden, initial_state, assert_stats_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


@pytest.mark.parametrize("name,world", [("edac_small", 2), ("edac_small", 3), ("edac_hc", 2), ("edac_hc", 4), ("edac_small_maxq", 2)])
def test_edac_member_sharded_matches_reference(name, world):
    from offlinerlkit_b200.engine.edac_sharded import EmulatedShardGroup
    from offlinerlkit_b200.parallel import partition_members
    from tests.gpu_common import build_policy, load_state, make_buffer
    g = Golden(name)
    m = g.meta
    pols = []
    for r in range(world):
        p = build_policy(m, DEV)
        load_state(p, initial_state(m))
        p.train()
        p.shard_critics(r, world, None)
        pols.append(p)
    buf, data = make_buffer(g, DEV)
    np.random.seed(m["np_seed"])
    group = None
    lr_atol = 2.5 * max(v for k, v in m["hyper"].items() if k.endswith("_lr"))
    for t in range(m["n_steps"]):
        batch = buf.sample(m["B"])
        if group is None:
            group = EmulatedShardGroup([p.engine(m["B"]) for p in pols])
        outs = group.step([batch] * world, noise=g.noise(t))
        ref = g.losses(t)
        for r, out in enumerate(outs):
            assert out.keys() == ref.keys()
            for k in ref:
                assert abs(out[k] - ref[k]) <= TOL * max(1.0, abs(ref[k])), (t, r, k, out[k], ref[k])
        assert all(o == outs[0] for o in outs), "the replicated quantities must be bit-identical on every rank"
        # every rank's actor is the same; the critics are assembled from the ranks' slices
        parts = partition_members(m["E"], world)
        for p in pols:
            p._engine.write_back()
        sd = {k: v.detach().cpu().clone() for k, v in pols[0].state_dict().items()}
        for r in range(1, world):
            sdr = pols[r].state_dict()
            for k, v in sdr.items():
                if k.startswith("actor."):
                    assert torch.equal(v.cpu(), sd[k]), (r, k)
                elif k.startswith("critics") and "saved_" not in k:
                    sd[k][parts[r][0]:parts[r][-1] + 1] = v.detach().cpu()[parts[r][0]:parts[r][-1] + 1]
        assert_stats_close(sd, g.group(f"stats{t}"), tol=TOL, lr_atol=lr_atol * (1 + 0.8 * t))


@pytest.mark.parametrize("name,world", [("dynamics_small", 2), ("dynamics_small", 3), ("dynamics_hc", 2), ("dynamics_hc", 4)])
def test_dynamics_member_sharded_matches_reference(name, world):
    """EnsembleDynamics.learn with the members sharded (BASELINE.json configs[4]), ranks emulated on one device in
    lockstep: per mini-batch the partial gradients of the shared log-variance bounds are exchanged and every rank applies
    the same Adam step to its replica.  Loss, assembled parameters and holdout losses vs the reference's unsharded run."""
    from offlinerlkit_b200.modules import EnsembleDynamicsModel
    from offlinerlkit_b200.dynamics import EnsembleDynamics
    from offlinerlkit_b200.parallel import partition_members
    from offlinerlkit_b200.synthetic import make_dataset
    from offlinerlkit_b200.utils.scaler import StandardScaler
    from offlinerlkit_b200.utils.termination_fns import termination_fn_halfcheetah
    from tests.helpers import rel_err
    g = Golden(name)
    m = g.meta
    d = make_dataset(m["n_data"], m["O"], m["A"], seed=m["data_seed"])
    x = np.concatenate([d["observations"], d["actions"]], axis=-1)
    y = np.concatenate([d["next_observations"] - d["observations"], d["rewards"].reshape(-1, 1)], axis=-1)
    xs = ((x - g["scaler_mu"]) / g["scaler_std"]).astype(np.float32)
    boot = g["boot"]
    dyns = []
    for r in range(world):
        model = EnsembleDynamicsModel(m["O"], m["A"], m["hidden"], num_ensemble=m["E"], num_elites=m["n_elites"],
                                      weight_decays=m["weight_decays"], device=DEV)
        model.load_state_dict(initial_state(m), strict=False)
        dyn = EnsembleDynamics(model, torch.optim.Adam(model.parameters(), lr=m["lr"]), StandardScaler(g["scaler_mu"], g["scaler_std"]),
                               termination_fn_halfcheetah, penalty_coef=0.5)
        dyn.shard_members(r, world, None)
        dyns.append(dyn)
    engs = [dy.engine for dy in dyns]
    src_x = torch.from_numpy(xs).to(DEV)
    src_y = torch.from_numpy(y.astype(np.float32)).to(DEV)
    idx = torch.from_numpy(boot).to(DEV)
    B, nb = m["B"], m["n_batches"]
    losses = [torch.zeros(nb * world, dtype=torch.float32, device=DEV) for _ in engs]
    for b in range(nb):
        sts = [e.learn_batch_begin(src_x, src_y, idx, b * B, B, 0.01) for e in engs]
        for e, st in zip(engs, sts):            # the exchange: every rank receives every rank's block
            n = st["recv"].numel() // world
            for q, sq in zip(engs, sts):
                st["recv"][q.rank * n:(q.rank + 1) * n].copy_(sq["send"])
        for e, st, ls in zip(engs, sts, losses):
            e.learn_batch_end(st, ls, b)
    vals = [e.pass_loss(ls, nb) for e, ls in zip(engs, losses)]
    assert all(v == vals[0] for v in vals)
    assert vals[0] == pytest.approx(float(g["learn_loss"]), rel=TOL)
    parts = partition_members(m["E"], world)
    for e in engs:
        e.write_back()
    sd = {k: v.detach().cpu().clone() for k, v in dyns[0].model.state_dict().items()}
    for r in range(1, world):
        sdr = dyns[r].model.state_dict()
        for k, v in sdr.items():
            if k in ("max_logvar", "min_logvar"):
                assert torch.equal(v.cpu(), sd[k]), (r, k)          # replicas of the shared bounds stay bit-identical
            elif v.dim() == 3 and "saved_" not in k:
                sd[k][parts[r][0]:parts[r][-1] + 1] = v.detach().cpu()[parts[r][0]:parts[r][-1] + 1]
    assert_stats_close(sd, g.group("stats"), tol=TOL, lr_atol=2.5 * m["lr"])
    hold = m["holdout"]
    hx, hy = src_x[:hold].contiguous(), src_y[:hold].contiguous()
    val = torch.cat([e.validate_local(hx, hy) for e in engs]).cpu().numpy()
    assert rel_err(val, g["val"]) < TOL
