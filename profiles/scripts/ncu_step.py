"""A few CQL steps (headline shape) and nothing else: the command line to put under ncu.
Usage: ncu ... python profiles/scripts/ncu_step.py [n_steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
policy, buf = bench.build_engine("cuda:0", seed=0, n_data=200_000)
for _ in range(n):
    out = policy.learn(buf.sample(bench.BATCH))
torch.cuda.synchronize()
print("ok", {k: round(v, 4) for k, v in out.items()})
