"""Gradient steps/s of every algorithm of the hot path at its golden-fixture shape (run on the GPU box):
public API loop (buffer.sample + policy.learn, host indices in, loss dict out) and device-resident graph replays.
Uses the metadata of tests/golden/*.npz (shapes / hyper-parameters), synthetic data, random-init weights.
Usage: python profiles/scripts/algo_bench.py [out.json]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from tests.gpu_common import build_policy, make_buffer
from tests.helpers import Golden

out = {}
for name in ("cql_hc", "combo_hc", "combo_hc_model", "sac_hc", "edac_hc", "iql_walker", "iql_walker_b1024", "td3bc_walker"):
    g = Golden(name)
    m = g.meta
    torch.manual_seed(0)
    np.random.seed(0)
    pol = build_policy(m)
    pol.train()
    buf, _ = make_buffer(g)
    B = m["B"]
    if m["algo"] == "combo":        # MBPolicyTrainer's loop: one draw from the real buffer, one from the model buffer
        fbuf, _ = make_buffer(g)
        draw = lambda: {"real": buf.sample(m["n_real"]), "fake": fbuf.sample(m["n_fake"])}
    else:
        draw = lambda: buf.sample(B)
    for _ in range(30):
        loss = pol.learn(draw())
    torch.cuda.synchronize()
    N = 1000
    t0 = time.perf_counter()
    for _ in range(N):
        loss = pol.learn(draw())
    e2e = N / (time.perf_counter() - t0)
    eng = pol._engine
    key = "step" if "step" in eng.plans else sorted(eng.plans)[0]
    plans = [eng.plans[k] for k in sorted(eng.plans)]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s = torch.cuda.ExternalStream(eng.rt.cur.value) if hasattr(eng.rt.cur, "value") and eng.rt.cur.value else torch.cuda.current_stream()
    with torch.cuda.stream(s):
        for p in plans:
            p.launch()
        ev0.record(s)
        for i in range(N):
            plans[i % len(plans)].launch()
        ev1.record(s)
    torch.cuda.synchronize()
    dev = N / (ev0.elapsed_time(ev1) * 1e-3)
    out[name] = {"algo": m["algo"], "batch": B, "obs": m["O"], "act": m["A"], "hidden": m["hidden"],
                 "steps_per_s_public_api": round(e2e, 1), "steps_per_s_graph_replay": round(dev, 1),
                 "us_per_step_graph_replay": round(1e6 / dev, 1), "launches": [p.n_launches for p in plans],
                 "finite_loss": bool(all(np.isfinite(v) for v in loss.values()))}
    print(name, out[name], flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
