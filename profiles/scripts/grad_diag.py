"""Per-tensor gradient errors of the engine vs the pinned oracle on golden fixtures (diagnostic; run on the GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from tests.helpers import Golden, initial_state
from tests.gpu_common import EngineGrads, build_policy, load_state, make_buffer, make_oracle

names = sys.argv[1].split(",")
precisions = sys.argv[2].split(",") if len(sys.argv) > 2 else ["tf32x3"]
for name in names:
    for prec in precisions:
        g = Golden(name)
        m = g.meta
        if m["algo"] == "combo":
            continue
        policy = build_policy(m, "cuda:0")
        load_state(policy, initial_state(m))
        policy.train()
        buf, data = make_buffer(g, "cuda:0")
        np.random.seed(m["np_seed"])
        ora = make_oracle(m)
        tap = None
        for t in range(m["n_steps"]):
            batch = buf.sample(m["B"])
            eng = policy.engine(m["B"])
            if t == 0:
                eng.precision = prec
            noise = g.noise(t) if any(k.startswith(f"noise{t}|") for k in g.z.files) else None
            if tap is None:
                tap = EngineGrads(policy, eng)
            tap.snapshot()
            out = policy.learn(batch, noise=noise) if noise is not None else policy.learn(batch)
            ref_b = g.batch(t, data)
            ora.step(ref_b, noise) if noise is not None else ora.step(ref_b)
            stats = g.group(f"gradstats{t}")
            got = tap.after(stats.keys())
            print(f"== {name} {prec} step {t}: losses {out}")
            for k in stats:
                a = got[k].double().reshape(-1).numpy()
                b = ora.grads[k].double().reshape(-1).numpy()
                l2 = np.sqrt(((a - b) ** 2).sum()) / max(np.sqrt((b * b).sum()), 1e-30)
                mx = np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
                # parameter difference after the step
                pe = (policy.state_dict()[k].detach().cpu().double().reshape(-1).numpy() - ora.p[k].detach().double().reshape(-1).numpy()) if k in ora.p else np.zeros(1)
                print(f"   {k:40s} relL2 {l2:.2e}  max|d|/max|g| {mx:.2e}  |g|max {np.abs(b).max():.3e}  param maxdiff {np.abs(pe).max():.2e}")
