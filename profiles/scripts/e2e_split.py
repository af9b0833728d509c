"""Where the host time of one public-API step goes (run on the GPU box): buf.sample vs policy.learn, and inside learn
the part spent waiting for the device.  Usage: python profiles/scripts/e2e_split.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from offlinerlkit_b200 import _lib as L

policy, buf = bench.build_engine("cuda:0", seed=0, n_data=200_000)
for _ in range(50):
    policy.learn(buf.sample(bench.BATCH))
eng = policy._engine
rt = eng.rt
orig_sync = rt.sync
wait = [0.0]
def timed_sync():
    t = time.perf_counter(); orig_sync(); wait[0] += time.perf_counter() - t
rt.sync = timed_sync
N = 3000
ts = tl = 0.0
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(N):
    a = time.perf_counter()
    b = buf.sample(bench.BATCH)
    c = time.perf_counter()
    policy.learn(b)
    d = time.perf_counter()
    ts += c - a; tl += d - c
tot = time.perf_counter() - t0
print(f"per step: total {1e6*tot/N:.1f} us | sample {1e6*ts/N:.1f} | learn {1e6*tl/N:.1f} (of which device wait {1e6*wait[0]/N:.1f})")
