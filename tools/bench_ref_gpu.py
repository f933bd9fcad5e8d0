#!/usr/bin/env python3
"""Times the reference custrings CUDA build (baseline/_ref/libref_gpu.so) on the B200 for the BASELINE.json workloads:
    python tools/bench_ref_gpu.py [--rows R --bytes B --reps K] [--ops contains,count,replace_re,tokenize,split_record,category]
create_from_offsets(devmem=true) is excluded from the op timings (reported separately)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from baseline import ref_gpu  # noqa: E402
from custrings_b200.workloads import c2_corpus  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--bytes", type=int, default=1 << 30)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--ops", default="contains,count,replace_re,tokenize")
a = ap.parse_args()
PAT = r"\b\w{4,}\b"


def timed(f, reps):
    f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = f()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        del r
    ts.sort()
    return ts[len(ts) // 2], ts[0]


chars, offsets, validity, nulls = c2_corpus(a.rows, a.bytes)
dc, do, dv = (torch.from_numpy(x).cuda() for x in (chars, offsets, validity))
# the reference uses the legacy default stream; torch's current stream is the legacy default stream too
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
col = ref_gpu.RefGpuStrings.from_device(dc, do, a.rows, dv, nulls)
e1.record()
torch.cuda.synchronize()
print(json.dumps({"op": "create_from_offsets(devmem)", "ms": e0.elapsed_time(e1), "rows": a.rows, "memsize": col.memsize()}))
res = torch.empty(a.rows, dtype=torch.uint8, device="cuda")
res32 = torch.empty(a.rows, dtype=torch.int32, device="cuda")
for op in a.ops.split(","):
    if op == "contains":
        m = col.contains_re(PAT, res)
        med, best = timed(lambda: col.contains_re(PAT, res), a.reps)
        print(json.dumps({"op": "contains_re", "pattern": PAT, "ms_median": med, "ms_best": best, "matches": m, "strings_per_s": a.rows / med * 1e3}))
    elif op == "count":
        med, best = timed(lambda: col.count_re(PAT, res32), a.reps)
        print(json.dumps({"op": "count_re", "ms_median": med, "ms_best": best}))
    elif op == "replace_re":
        med, best = timed(lambda: col.replace_re(PAT, "#"), max(1, a.reps // 2))
        print(json.dumps({"op": "replace_re", "ms_median": med, "ms_best": best}))
    elif op == "tokenize":
        med, best = timed(lambda: col.tokenize(), max(1, a.reps // 2))
        print(json.dumps({"op": "tokenize", "ms_median": med, "ms_best": best}))
    elif op == "split_record":
        med, best = timed(lambda: col.split_record_total(" "), 1)
        print(json.dumps({"op": "split_record", "ms_median": med, "ms_best": best}))
    elif op == "category":
        def f():
            c = col.category()
            ref_gpu.lib().refgpu_category_destroy(c)
        med, best = timed(f, max(1, a.reps // 2))
        print(json.dumps({"op": "category", "ms_median": med, "ms_best": best}))
    sys.stdout.flush()
