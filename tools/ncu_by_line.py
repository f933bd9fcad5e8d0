#!/usr/bin/env python3
"""Dynamic instruction accounting of one kernel from an `ncu --set full --import-source on` capture: executed warp-instructions
per source line of the kernel body (inlined helpers are charged to the line that calls them), per pipe, plus the SASS opcode
histogram weighted by execution count.  The SASS <-> source-line map comes from nvdisasm's line info of the object that was
profiled (built with -lineinfo).

    python tools/ncu_by_line.py REPORT.ncu-rep OBJECT.o KERNEL_SUBSTRING [--windows N] [--launch I] [--depth D] > profiles/...txt
(--launch I: the I-th launch of a report that holds several)"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, key = sys.argv[1:4]
nwin = float(sys.argv[sys.argv.index("--windows") + 1]) if "--windows" in sys.argv else 524288.0
depth = int(sys.argv[sys.argv.index("--depth") + 1]) if "--depth" in sys.argv else 0  # 1: the kernel body is itself an inlined function
ALU = {"LOP3", "SHF", "PRMT", "ISETP", "SEL", "IADD3", "VIADD", "LEA", "PLOP3", "POPC", "FLO", "BREV", "VIMNMX", "SGXT", "BMSK", "MOV", "IADD", "VOTE"}

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]  # one section per captured launch
which = int(sys.argv[sys.argv.index("--launch") + 1]) if "--launch" in sys.argv else 0
h0 = heads[which]
h1 = heads[which + 1] if which + 1 < len(heads) else len(rows)
hdr, data = rows[h0 + 1], rows[h0 + 2:h1]
iS, iE = hdr.index("Source"), hdr.index("Instructions Executed")
execs = [(r[iS].strip(), int(r[iE] or 0)) for r in data if len(r) > max(iS, iE)]

with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, capture_output=True)
    cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and key in l)
end = next((i for i in range(start + 1, len(dis)) if dis[i].startswith("//-----") and ".text." in dis[i]), len(dis))
ins, cur, pend = [], None, []
for l in dis[start:end]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        pend.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        if pend:
            cur, pend = pend[max(0, len(pend) - 1 - depth)], []
        ins.append((m.group(2).strip(), cur))
if len(ins) != len(execs):
    sys.exit("instruction count mismatch: report %d vs object %d (not the profiled binary?)" % (len(execs), len(ins)))

by, alu, fma, ops = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
for (s, e), (t, line) in zip(execs, ins):
    o = re.match(r"(@!?U?P\w+\s+)?([\w\.]+)", t).group(2)
    base = o.split(".")[0]
    by[line] += e
    ops[o if base == "IMAD" else base] += e
    if base in ALU:
        alu[line] += e
    elif base in ("IMAD", "FFMA", "FMUL", "FADD"):
        fma[line] += e
tot = sum(by.values())
print("kernel %s: %d warp-instructions executed = %.1f per 2 KiB of chars (%d windows); ALU pipe %.1f, FMA pipe %.1f per window"
      % (key, tot, tot / nwin, nwin, sum(alu.values()) / nwin, sum(fma.values()) / nwin))
print("\n-- per source line of the kernel body (>= 0.5 per window)        total     ALU     FMA")
for k in sorted(by, key=lambda k: (k[0], k[1])):
    if by[k] / nwin >= 0.5:
        print("%-28s %5d  %31.1f %7.1f %7.1f" % (k[0], k[1], by[k] / nwin, alu[k] / nwin, fma[k] / nwin))
print("\n-- SASS opcode histogram (executed warp-instructions per window)")
for o, v in ops.most_common(40):
    print("%-18s %12d %8.1f" % (o, v, v / nwin))
