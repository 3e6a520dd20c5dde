"""Summarises `ncu --metrics <list> --csv` of one training step of bench.py into one line per (kernel, grid size):
launches, average duration, DRAM bytes and GB/s, L2->SM bytes, tensor-pipe and issue activity.

    python tools/step_metrics.py gpurun_out/step_metrics.csv profiles/<name>.csv "<command>"
Per-launch times under ncu are cold-cache and serialised: shares and per-kernel ratios are the evidence, not absolutes."""
import csv
import sys
from collections import defaultdict

raw, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
hdr = None
per = defaultdict(dict)
meta = {}
for r in csv.reader(open(raw, errors="replace")):
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    meta[d["ID"]] = (name[:70], d["Grid Size"].replace(" ", ""))
    try:
        per[d["ID"]][d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        pass
units = {}
agg = defaultdict(lambda: defaultdict(float))
for i, m in per.items():
    a = agg[meta[i]]
    a["n"] += 1
    for k, v in m.items():
        a[k] += v
tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
T, DR, DW, X = "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"
TP, IS = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active"
with open(out, "w") as f:
    f.write(f"# {cmd}\n# {len(per)} launches, {tot / 1e6:.2f} ms of kernel time in all (ns durations; cold-cache, serialised under ncu)\n")
    f.write("share_pct,launches,avg_us,dram_MB_per_launch,dram_GBps,l2_to_sm_MB_per_launch,tensor_pipe_pct,issue_active_pct,grid,kernel\n")
    for (name, grid), a in sorted(agg.items(), key=lambda kv: -kv[1][T]):
        n = a["n"]
        us = a[T] / n / 1e3
        dram = (a[DR] + a[DW]) / n
        f.write(f"{100 * a[T] / tot:.2f},{int(n)},{us:.1f},{dram / 1e6:.2f},{dram / (us * 1e3):.0f},{a[X] / n / 1e6:.2f},"
                f"{a[TP] / n:.1f},{a[IS] / n:.1f},\"{grid}\",{name}\n")
print(open(out).read()[:3000])
