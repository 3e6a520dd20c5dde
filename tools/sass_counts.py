"""SASS evidence for the tcgen05 kernels: per kernel of libscn_b200.so, how many UTCHMMA (tcgen05.mma), UTCBAR
(tcgen05.commit), LDTM (tcgen05.ld), UBLKCP (cp.async.bulk), LDGSTS (cp.async), SYNCS (mbarrier) instructions the
sm_100a cubin holds, plus registers / shared memory from the resource usage.   python tools/sass_counts.py > profiles/<name>.txt"""
import collections
import os
import re
import subprocess

lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sparseeventid_b200", "lib", "libscn_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
ops = ["UTCHMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "UBLKCP", "LDGSTS", "SYNCS", "HMMA", "REDG", "ATOMG"]
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = counts.setdefault(m.group(1), collections.Counter())
        continue
    if cur is None:
        continue
    mm = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if mm:
        cur["_n"] += 1
        for o in ops:
            if mm.group(1).startswith(o):
                cur[o] += 1
usage = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
    usage[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)))
demangle = subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.basename(lib)}: cubin architectures {arch}; {len(counts)} kernels")
print(f"# {'kernel':78s} {'instr':>6s} {'regs':>4s} " + " ".join(f"{o:>8s}" for o in ops))
tot = collections.Counter()
for (name, c), dn in zip(counts.items(), demangle):
    short = re.sub(r"\(.*", "", dn.replace("(anonymous namespace)::", "").replace("void ", ""))
    if not any(c[o] for o in ops[:6]):          # only the kernels that use the Blackwell async / tensor paths
        continue
    regs = usage.get(name, (0, 0, 0))[0]
    print(f"{short[:80]:80s} {c['_n']:6d} {regs:4d} " + " ".join(f"{c[o]:8d}" for o in ops))
    tot.update(c)
print(f"{'TOTAL (listed kernels)':80s} {tot['_n']:6d}      " + " ".join(f"{tot[o]:8d}" for o in ops))
