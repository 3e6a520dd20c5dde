"""Instruction accounting of one k_conv_tc launch from an ncu --set full --import-source on report (source page):
which role issued how many warp instructions, how many of them were mbarrier retry loops, and the headline counters.

    python tools/ncu_regions.py gpurun_out/conv_tc_<shape>.ncu-rep [stages]        (stages = tiles x stages per tile)
The kernel is issue-bound (smsp__issue_active ~60%), so instructions per stage is the number to drive down."""
import csv
import subprocess
import sys

rep = sys.argv[1]
stages = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, vals = raw[0], raw[-1]
d = dict(zip(hdr, vals))
name = d.get("Kernel Name", "?")
print(f"kernel: {name}")
for k in ("gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
          "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
          "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"):
    if k in d:
        print(f"  {k}: {d[k]} {raw[1][hdr.index(k)]}")
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h = src[1]
rows = src[2:]
isrc, iex, ismp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
ins = [r[isrc].strip() for r in rows]
ex = [int(r[iex]) for r in rows]
sm = [int(r[ismp]) for r in rows]
tot, tots = sum(ex), max(sum(sm), 1)


def first(pat, start=0):
    for i in range(start, len(ins)):
        if pat in ins[i]:
            return i
    return -1


i_ballot = first("VOTE.ANY")
i_prod0 = max(i for i in range(i_ballot) if "LDG.E.NA" in ins[i]) - 40 if i_ballot > 0 else 0
i_arr = first("ARRIVES.LDGSTSBAR")
i_blk = first("UBLKCP")
i_mma = first("UTCHMMA")
i_ldtm = first("LDTM")
last_utcbar = max(i for i, s in enumerate(ins) if "UTCBAR" in s)
bounds = [("setup", 0, i_prod0), ("gathering warps (16)", i_prod0, i_arr + 20), ("weight loader", i_arr + 20, i_blk + 25),
          ("issuing warps (<= 4)", i_blk + 25, last_utcbar + 3), ("epilogue warps (4)", last_utcbar + 3, len(ins))]
print(f"warp instructions executed: {tot}" + (f" = {tot / stages:.0f} per stage ({stages} stages)" if stages else ""))
for nm, a, b in bounds:
    e, s = sum(ex[a:b]), sum(sm[a:b])
    print(f"  {nm:24s} {100 * e / tot:5.1f}% of instructions, {100 * s / tots:5.1f}% of stall samples" +
          (f", {e / stages:7.1f} per stage" if stages else ""))
# mbarrier retry loops: instructions between a TRYWAIT and the BPT.TRAP that bounds its loop, beyond the first pass
spin = 0
for i, s in enumerate(ins):
    if "TRYWAIT" in s:
        j = first("BPT.TRAP", i)
        if 0 < j - i < 12:
            body = ex[i + 1:j]
            retries = min(body) if body else 0
            spin += sum(min(v, retries) for v in ex[i - 1:j])
            print(f"  try_wait @{i}: executed {ex[i]}, retried {retries} times")
print(f"mbarrier retry loops: ~{spin} instructions = {100 * spin / tot:.1f}% of all issued")
