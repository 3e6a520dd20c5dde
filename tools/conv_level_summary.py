"""Per-level `ncu --set full` summaries of tc::k_conv_tc (the six level shapes of the bench workload) from the reports a
GPU call brought back in gpurun_out/, tied to the build by the sha256 of the kernel's sources (bench.kernel_source_hash).

    python tools/conv_level_summary.py profiles/<name>.json
Also writes profiles/r02_conv_tc_ncu_summary.json (the 317 k x 64 level), which bench.py reads for `roofline.traffic`."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

SHAPES = [(495518, 27, 32, 32), (317485, 27, 64, 64), (154605, 27, 96, 96), (59700, 27, 128, 128), (20727, 27, 160, 160),
          (7332, 27, 192, 192)]
KEYS = {"gpu__time_duration.sum": "gpu_time", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm", "lts__t_sectors_srcunit_tex_op_read.sum": "lts_tex_read_sectors",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "lsu_wavefronts_pct",
        "launch__registers_per_thread": "registers", "launch__grid_size": "grid",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
        "smsp__inst_executed.sum": "warp_instructions"}
MULT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
out = {"what": "ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 2 -c 1 python tools/tc_profile.py <shape> 3",
       "kernel_source_sha256": bench.kernel_source_hash(), "levels": []}
for n, K, cin, cout in SHAPES:
    rep = os.path.join(ROOT, "gpurun_out", f"conv_tc_{n}_{K}_{cin}_{cout}.ncu-rep")
    if not os.path.exists(rep):
        continue
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))
    rec = {"shape": f"{n} rows, {K} offsets, {cin}->{cout}", "kernel": d.get("Kernel Name")}
    for k, name in KEYS.items():
        if k in d:
            try:
                rec[name] = float(d[k].replace(",", "")) * MULT.get(u.get(k, ""), 1.0)
            except ValueError:
                pass
    plain = os.path.join(ROOT, "gpurun_out", f"plain_{n}_{K}_{cin}_{cout}.log")
    pairs = None
    if os.path.exists(plain):
        txt = open(plain).read()
        if "pairs=" in txt:
            pairs = int(txt.split("pairs=")[1].split(",")[0])
            rec["plain_run"] = txt.strip().splitlines()[0]
    if pairs:
        rec["pairs"] = pairs
        rec["flops"] = 2.0 * pairs * cin * cout
        rec["algorithmic_bytes"] = n * cin * 2.0 + n * cout * 2.0 + K * cin * cout * 2.0 + 8.0 * pairs
        if "dram_read" in rec:
            rec["dram_over_algorithmic"] = (rec["dram_read"] + rec["dram_write"]) / rec["algorithmic_bytes"]
            rec["l2_to_sm_over_algorithmic"] = rec.get("l2_to_sm", 0.0) / rec["algorithmic_bytes"]
    out["levels"].append(rec)
json.dump(out, open(sys.argv[1], "w"), indent=1)
for r in out["levels"]:
    print(r["shape"], {k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items() if k in
                       ("gpu_time", "tensor_pipe_active_pct", "issue_active_pct", "dram_over_algorithmic", "l2_to_sm_over_algorithmic")})
lvl = [r for r in out["levels"] if r["shape"].startswith("317485")]
if lvl and "dram_read" in lvl[0]:
    r = lvl[0]
    json.dump({"what": out["what"], "kernel_source_sha256": out["kernel_source_sha256"], "shape": r["shape"],
               "traffic_bytes_per_launch": r["dram_read"] + r["dram_write"], "algorithmic_bytes_per_launch": r.get("algorithmic_bytes"),
               "gpu_time_us": r.get("gpu_time"), "tensor_pipe_active_pct": r.get("tensor_pipe_active_pct"),
               "l2_to_sm_bytes": r.get("l2_to_sm")}, open(os.path.join(ROOT, "profiles", "r02_conv_tc_ncu_summary.json"), "w"), indent=1)
