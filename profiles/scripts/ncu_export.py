"""Turn ncu outputs (read in the authoring container) into the committed summaries under profiles/.
  python profiles/scripts/ncu_export.py launches <ncu --csv launch list> <out.csv> <out_summary.csv> "<command>"
  python profiles/scripts/ncu_export.py full <report.ncu-rep> <out.csv> "<command>" [kernel-regex]"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", name)
    return name.split("(")[0]


def launches(src, out, out_sum, cmd):
    rows = list(csv.reader(ln for ln in open(src) if ln.startswith('"')))
    head = rows[0]
    ci = {n: i for i, n in enumerate(head)}
    data = [r for r in rows[1:] if len(r) == len(head) and r[ci["Metric Name"]] == "gpu__time_duration.sum"]
    with open(out, "w") as f:
        f.write(f"# {cmd}\n# one line per kernel launch, in launch order (cold caches, serialised: shares, not absolute times, "
                "are comparable with the captured step)\nid,kernel,grid,block,duration_us\n")
        for r in data:
            v = float(r[ci["Metric Value"]].replace(",", ""))
            unit = r[ci["Metric Unit"]]
            us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
            f.write(f'{r[ci["ID"]]},"{short(r[ci["Kernel Name"]])}","{r[ci["Grid Size"]]}","{r[ci["Block Size"]]}",{us:.3f}\n')
    tot = OrderedDict()
    for r in data:
        v = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        k = short(r[ci["Kernel Name"]])
        n, t = tot.get(k, (0, 0.0))
        tot[k] = (n + 1, t + us)
    total = sum(t for _, t in tot.values())
    with open(out_sum, "w") as f:
        f.write(f"# per-kernel totals of {out} ({cmd}); share = of the summed kernel time\nkernel,launches,total_us,share\n")
        for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f'"{k}",{n},{t:.2f},{t / total:.4f}\n')
    print(f"{len(data)} launches, {total:.0f} us summed -> {out}, {out_sum}")


def full(rep, out, cmd, rx=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    ci = {n: i for i, n in enumerate(head)}
    want = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_bytes.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max"]
    want = [w for w in want if w in ci]
    with open(out, "w") as f:
        f.write(f"# {cmd}\n")
        w = csv.writer(f)
        w.writerow(want)
        w.writerow([units[ci[n]] for n in want])
        for r in rows[2:]:
            if len(r) < len(head) or (rx and not re.search(rx, r[ci["Kernel Name"]])):
                continue
            w.writerow([short(r[ci[n]]) if n == "Kernel Name" else r[ci[n]] for n in want])
    print("wrote", out)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(*sys.argv[2:6])
    else:
        full(*sys.argv[2:])
