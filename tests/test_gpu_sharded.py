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
