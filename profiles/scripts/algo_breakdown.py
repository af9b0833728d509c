"""Per-launch device time of one algorithm's captured step (run on the GPU box).  Usage: python profiles/scripts/algo_breakdown.py edac_hc"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from tests.gpu_common import build_policy, make_buffer
from tests.helpers import Golden
name = sys.argv[1] if len(sys.argv) > 1 else "edac_hc"
g = Golden(name); m = g.meta
torch.manual_seed(0); np.random.seed(0)
pol = build_policy(m); pol.train()
buf, _ = make_buffer(g)
for _ in range(10):
    pol.learn(buf.sample(m["B"]))
eng = pol._engine
for key in sorted(eng.plans):
    class E: pass
    e = E(); e.plans = {"step": eng.plans[key]}; e.rt = eng.rt
    br = bench.per_launch_breakdown(e)
    print(f"== plan {key}: {len(br)} launches, sum {sum(u for _, u in br):.1f} us")
    for i, (l, u) in enumerate(br):
        print(f"{l:34s} {u:6.1f}", end="  ")
        if i % 3 == 2: print()
    print()
