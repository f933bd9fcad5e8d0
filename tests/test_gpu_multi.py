"""Multi-GPU (needs >= 2 devices, skipped otherwise): row-sharded NVCategory build with the NCCL key all-gather."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

WORKER = r'''
import json, os, sys, random
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from custrings_b200 import nvstrings, nvcategory
from custrings_b200._lib import lib
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); lib().custr_set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
rng = random.Random(1234)
keys = ["k%%03d" %% i for i in range(50)] + ["é", "", "zz" * 9]
rows = [rng.choice(keys + [None]) for _ in range(4000)]
lo, hi = len(rows) * rank // world, len(rows) * (rank + 1) // world
cat = nvcategory.from_strings_sharded(nvstrings.to_device(rows[lo:hi]))
out = {"keys": cat.keys().to_host(), "values": cat.values(), "rows": rows[lo:hi]}
json.dump(out, open(os.path.join(sys.argv[1], "rank%%d.json" %% rank), "w"))
dist.barrier(); dist.destroy_process_group()
'''


def test_category_sharded_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29633", str(script), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    outs = [json.load(open(tmp_path / ("rank%d.json" % k))) for k in range(2)]
    all_rows = outs[0]["rows"] + outs[1]["rows"]
    want_keys = sorted({x for x in all_rows if x is not None}, key=lambda s: s.encode())
    if any(x is None for x in all_rows):
        want_keys = [None] + want_keys
    for o in outs:
        assert o["keys"] == want_keys
        assert [want_keys[v] for v in o["values"]] == o["rows"]
