"""Top stall instructions of an ncu --import-source report + headline counters.   python tools/ncu_top.py <rep> [n]"""
import csv, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
d = dict(zip(raw[0], raw[-1]))
print(d.get("Kernel Name"))
for k in ("gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
          "smsp__warps_active.avg.per_cycle_active", "launch__grid_size", "launch__registers_per_thread"):
    if k in d:
        print(f"  {k}: {d[k]} {raw[1][raw[0].index(k)]}")
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h, rows = src[1], src[2:]
isrc, iex, ismp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
ci = {c: h.index(c) for c in cols}
tot = sum(int(r[iex]) for r in rows); tots = sum(int(r[ismp]) for r in rows)
print(f"instructions executed {tot}, samples {tots}")
for i in sorted(sorted(range(len(rows)), key=lambda i: -int(rows[i][ismp]))[:topn]):
    r = rows[i]
    st = {c[6:]: int(r[ci[c]]) for c in cols if int(r[ci[c]]) > 0}
    print(f"{i:5d} ex {int(r[iex]):9d} smp {int(r[ismp]):5d} ({100 * int(r[ismp]) / max(tots, 1):4.1f}%) {r[isrc].strip()[:58]:58s} {st}")
