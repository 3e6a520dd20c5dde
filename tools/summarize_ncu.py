"""Turns the raw ncu outputs brought back in gpurun_out/ into the small summaries committed under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/r01b_launches_raw.csv profiles/r01b_launch_summary.csv "<command>"
    python tools/summarize_ncu.py kernel   gpurun_out/r01b_conv_tc_full.ncu-rep profiles/r01b_conv_tc_ncu_summary.json "<command>" \
                                           <flops_per_launch> <algorithmic_bytes_per_launch>
"""
import csv
import json
import subprocess
import sys
from collections import defaultdict


def launches(raw, out, command):
    rows = [r for r in csv.reader(open(raw)) if len(r) > 14 and r[0].isdigit()]
    agg = defaultdict(lambda: [0.0, 0])
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "")
        agg[name][0] += float(r[14]) / 1e3
        agg[name][1] += 1
    total = sum(v[0] for v in agg.values())

    def fam(n):
        if "k_conv_tc" in n or "k_conv_mma" in n or "k_conv_generic" in n or "k_conv_cin1" in n:
            return "conv fwd/dgrad"
        if "k_wgrad" in n:
            return "conv wgrad"
        if "k_bn" in n or "k_col_reduce" in n or "k_acc_to_float" in n:
            return "batch norm / column sums"
        if "k_subm" in n or "k_insert" in n or "k_strided" in n or "RadixSort" in n or "DeviceScan" in n or "k_input" in n \
                or "k_pack" in n or "k_mirror" in n or "k_head" in n:
            return "rulebooks (hash/sort/scan)"
        if "at::" in n or "cutlass" in n or "cublas" in n or "nccl" in n:
            return "torch (heads, loss, Adam, copies)"
        return "other scn kernels"
    fams = defaultdict(float)
    for n, (us, _) in agg.items():
        fams[fam(n)] += us
    with open(out, "w") as f:
        f.write(f"# ncu launch list of `{command}`: {len(rows)} launches, {total:.1f} us in all\n")
        f.write("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write("# family shares: " + "; ".join(f"{k} {100 * v / total:.1f}%" for k, v in sorted(fams.items(), key=lambda kv: -kv[1])) + "\n")
        f.write("share_pct,total_us,launches,kernel\n")
        for n, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
            f.write(f"{100 * us / total:.2f},{us:.1f},{c},{n[:110]}\n")
    print(open(out).read()[:1500])


def kernel(rep, out, command, flops, abytes):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, vals = rows[0], rows[-1]
    d = dict(zip(hdr, vals))

    def g(k, default=None):
        v = d.get(k, default)
        try:
            return float(str(v).replace(",", ""))
        except (TypeError, ValueError):
            return v
    rd, wr = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
    units = dict(zip(hdr, rows[1])) if len(rows) > 2 else {}

    def to_bytes(v, k):
        u = units.get(k, "")
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        return v * mult if isinstance(v, float) else v
    rd, wr = to_bytes(rd, "dram__bytes_read.sum"), to_bytes(wr, "dram__bytes_write.sum")
    dur = g("gpu__time_duration.sum")
    dur_us = dur * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units.get("gpu__time_duration.sum", "ns"), 1e-3)
    summary = {
        "what": f"ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 2 -c 1 of `{command}`",
        "kernel": d.get("Kernel Name"), "report": rep + " (scratch, not committed)",
        "gpu_time_us": dur_us, "dram_bytes_read": rd, "dram_bytes_write": wr,
        "traffic_bytes_per_launch": (rd or 0) + (wr or 0),
        "algorithmic_bytes_per_launch": abytes, "flops_per_launch": flops,
        "dram_throughput_pct_of_peak": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "tensor_pipe_active_pct": g("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
                                    g("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active")),
        "l2_hit_rate_pct": g("lts__t_sector_hit_rate.pct"),
        "registers_per_thread": g("launch__registers_per_thread"), "grid": g("launch__grid_size"),
        "block": g("launch__block_size"), "dynamic_smem_kb": g("launch__shared_mem_per_block_dynamic"),
        "sm_cycles": g("sm__cycles_elapsed.max"),
    }
    tens = [k for k in hdr if "tensor" in k and "pct" in k]
    summary["tensor_metrics"] = {k: g(k) for k in tens[:8]}
    json.dump(summary, open(out, "w"), indent=1)
    print(json.dumps(summary, indent=1)[:2500])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        kernel(sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5]), float(sys.argv[6]))
