"""Kernel timeline of one captured CQL step (run on the GPU box): replays the step graph under torch.profiler (CUPTI
activity records, which see the kernels of liborlk_b200.so as well) and prints every kernel of the last replay with
its start offset, duration and stream.  Profiling aid only - numbers taken under the profiler are not bench values.
Usage: python profiles/timeline.py [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench

policy, buf = bench.build_engine("cuda:0", seed=0, n_data=200_000)
for _ in range(10):
    policy.learn(buf.sample(bench.BATCH))
eng = policy._engine
plan = eng.plans["step"]
torch.cuda.synchronize()
REPS = 4
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(REPS):
        plan.launch()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
n = len(ev) // REPS
last = ev[-n:]
t0 = last[0].time_range.start
rows = []
for e in last:
    rows.append({"name": e.name[:60], "start_us": round(e.time_range.start - t0, 2),
                 "dur_us": round(e.time_range.end - e.time_range.start, 2)})
labels = [lbl for lbl, _ in plan.flat_ops]
end = max(r["start_us"] + r["dur_us"] for r in rows)
busy = sum(r["dur_us"] for r in rows)
print(f"{len(rows)} device activities in the last replay; span {end:.1f} us; sum of durations {busy:.1f} us")
prev_end = 0.0
for r in rows:
    gap = r["start_us"] - prev_end
    print(f"{r['start_us']:8.2f} +{r['dur_us']:7.2f} end {r['start_us'] + r['dur_us']:8.2f}  gap {gap:6.2f}  {r['name'][:48]}")
    prev_end = max(prev_end, r["start_us"] + r["dur_us"])
if len(sys.argv) > 1:
    json.dump({"labels": labels, "rows": rows, "span_us": end, "busy_us": busy}, open(sys.argv[1], "w"), indent=1)
