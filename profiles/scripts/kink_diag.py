"""Diagnostic (GPU box): for golden fixtures, how many near-zero ReLU pre-activations the oracle sees, how many of their
rows are found among the engine's activation buffers, which bits differ and what gradient error is left."""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from tests.helpers import Golden, initial_state
from tests.gpu_common import EngineGrads, build_policy, load_state, make_buffer, make_oracle, sync_oracle_to_engine
from tests.kinks import KinkRecorder, engine_activation_rows, _per_tensor_ok

for name in sys.argv[1].split(","):
    prec = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
    g = Golden(name)
    m = g.meta
    policy = build_policy(m, "cuda:0")
    load_state(policy, initial_state(m))
    policy.train()
    buf, data = make_buffer(g, "cuda:0")
    np.random.seed(m["np_seed"])
    ora = make_oracle(m)
    tap = None
    for t in range(m["n_steps"]):
        batch = buf.sample(m["B"])
        if m["algo"] == "combo":
            raise SystemExit("combo fixtures: use tests/test_gpu_cql.py (two buffers)")
        eng = policy.engine(m["B"])
        if t == 0:
            eng.precision = prec
        noise = g.noise(t) if any(k.startswith(f"noise{t}|") for k in g.z.files) else None
        if tap is None:
            tap = EngineGrads(policy, eng)
        tap.snapshot()
        if t > 0:
            sync_oracle_to_engine(ora, policy)
            sd = policy.state_dict()
            worst = max(((sd[k].detach().cpu() - v.detach()).abs().max().item(), k) for k, v in ora.p.items() if k in sd and v.is_floating_point())
            print("   after sync: max param diff", worst)
        policy.learn(batch, noise=noise) if noise is not None else policy.learn(batch)
        ref_b = g.batch(t, data)
        run = (lambda o: o.step(ref_b, noise)) if noise is not None else (lambda o: o.step(ref_b))
        stats = g.group(f"gradstats{t}")
        got = tap.after(stats.keys())
        gotn = {k: got[k].double().reshape(-1).numpy() for k in got}
        o = copy.deepcopy(ora)
        with KinkRecorder(2e-5) as rec:
            run(o)
        base = {k: o.grads[k].double().reshape(-1).numpy() for k in got}
        bad = _per_tensor_ok(gotn, base, 1e-4)
        acts = engine_activation_rows(eng)
        print(f"== {name} {prec} step {t}: {len(rec.found)} candidates; engine activation rows: { {w: tuple(v.shape) for w, v in acts.items()} }; "
              f"tensors off before adjustment: {len(bad)}")
        flips, dists = [], []
        for (c, i, z), row in zip(rec.found, rec.rows):
            w = row.numel()
            if w not in acts:
                continue
            E = acts[w]
            d = (E - row.to(E.device)).abs().amax(dim=1)
            j = int(d.argmin())
            dists.append(float(d[j]))
            if dists[-1] <= 3e-5 and bool(E[j, i % w] > 0) != (z > 0):
                flips.append((c, i, bool(E[j, i % w] > 0), z))
        dists = np.asarray(dists)
        print(f"   matched rows: {int((dists <= 3e-5).sum())} of {len(dists)} (distance quantiles {np.quantile(dists, [0.5, 0.9, 1.0]) if len(dists) else None}); "
              f"differing bits: {[(c, f'{z:.1e}') for c, i, b, z in flips]}")
        if flips:
            o2 = copy.deepcopy(ora)
            with KinkRecorder(0.0, [(c, i, b) for c, i, b, z in flips]):
                run(o2)
            adj = {k: o2.grads[k].double().reshape(-1).numpy() for k in got}
            bad = _per_tensor_ok(gotn, adj, 1e-4)
        print("   still off:", [(k, f"{l2:.1e}", f"{mx:.1e}") for k, l2, mx in bad])
        run(ora)
