"""Are the critic-gradient differences between the tensor-core (tf32x3) step and the reference ReLU-mask flips or arithmetic?
After one engine CQL step on the cql_hc fixture: recompute the critic pass in float64 from the engine's own inputs, count
the ReLU masks that differ, and compare the engine's weight gradients with float64 gradients (a) under the engine's masks
and (b) under the float64 masks.  Run on the GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from tests.helpers import Golden, initial_state
from tests.gpu_common import EngineGrads, build_policy, load_state, make_buffer

name = sys.argv[1] if len(sys.argv) > 1 else "cql_hc"
for prec in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["fp32", "tf32x3"]):
    g = Golden(name)
    m = g.meta
    policy = build_policy(m, "cuda:0")
    load_state(policy, initial_state(m))
    policy.train()
    buf, data = make_buffer(g, "cuda:0")
    np.random.seed(m["np_seed"])
    batch = buf.sample(m["B"])
    eng = policy.engine(m["B"])
    eng.precision = prec
    W = {k: v.detach().clone().double() for k, v in policy.state_dict().items()}
    tap = EngineGrads(policy, eng)
    tap.snapshot()
    policy.learn(batch, noise=g.noise(0))
    torch.cuda.synchronize()
    run = eng.run_critic
    X = eng.Xc.double()                       # [Mc, O+A]
    dq = run.dOut.double()                    # [2, Mc, 1]
    names = [k for k in g.group("gradstats0") if k.startswith("critic")]
    got = tap.after(names)
    print(f"== {name} {prec}")
    for c in (0, 1):
        p = f"critic{c + 1}"
        Ws = [W[f"{p}.backbone.model.{2 * l}.weight"] for l in range(3)]
        bs = [W[f"{p}.backbone.model.{2 * l}.bias"] for l in range(3)]
        wh = W[f"{p}.last.weight"]            # [1, 256]
        h, Z, Hs = X, [], []
        for l in range(3):
            z = h @ Ws[l].t() + bs[l]
            Z.append(z)
            h = torch.relu(z)
            Hs.append(h)
        eng_masks = [(run.H[l][c] > 0) for l in range(3)]
        for l in range(3):
            flips = (eng_masks[l] != (Z[l] > 0))
            herr = (run.H[l][c].double() - Hs[l]).abs().max().item()
            print(f"   {p} layer {l}: mask flips {int(flips.sum())} of {flips.numel()}  (|z| at flips <= "
                  f"{Z[l][flips].abs().max().item() if flips.any() else 0:.2e}); max |H_engine - H_fp64| {herr:.2e} (scale {Hs[l].abs().max().item():.2f})")
        for tag, masks in (("engine masks", eng_masks), ("fp64 masks", [z > 0 for z in Z])):
            dZ2 = (dq[c] * wh) * masks[2]
            dZ1 = (dZ2 @ Ws[2]) * masks[1]
            dZ0 = (dZ1 @ Ws[1]) * masks[0]
            ref = {f"{p}.backbone.model.4.weight": dZ2.t() @ Hs[1], f"{p}.backbone.model.4.bias": dZ2.sum(0),
                   f"{p}.backbone.model.2.weight": dZ1.t() @ Hs[0], f"{p}.backbone.model.2.bias": dZ1.sum(0),
                   f"{p}.backbone.model.0.weight": dZ0.t() @ X, f"{p}.backbone.model.0.bias": dZ0.sum(0),
                   f"{p}.last.weight": (dq[c] * Hs[2]).sum(0, keepdim=True), f"{p}.last.bias": dq[c].sum(0)}
            line = []
            for k, r in ref.items():
                a, b = got[k].double().reshape(-1), r.cpu().reshape(-1)
                l2 = ((a - b).norm() / b.norm()).item()
                line.append(f"{k.split('.', 1)[1].replace('backbone.model.', 'L')}: {l2:.1e}")
            print(f"   {p} rel-L2 of the engine's gradients vs float64 under the {tag}: " + "  ".join(line))
