"""Host-side logic that needs no GPU: packing, row sharding, the C2 generator's invariants, and the world_size-2
(gloo) run of the NVCategory key exchange used for multi-GPU builds."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pack_roundtrip():
    from custrings_b200.nvstrings import _pack
    strs = ["abc", None, "", "é日", "x" * 100]
    chars, offsets, validity, nulls = _pack(strs)
    assert nulls == 1 and offsets.tolist() == [0, 3, 3, 3, 8, 108] and validity.tolist() == [0b11101]
    raw = chars.tobytes()
    assert raw[3:8].decode() == "é日"


def test_c2_generator_invariants():
    from custrings_b200.workloads import c2_corpus, slice_rows
    chars, offsets, validity, nulls = c2_corpus(20000, 20000 * 107, seed=5)
    assert offsets[-1] == 20000 * 107 == chars.size and offsets.dtype == np.int32
    valid = np.unpackbits(validity, bitorder="little")[:20000].astype(bool)
    lens = np.diff(offsets)
    assert nulls == (~valid).sum() and (lens[~valid] == 0).all() and (lens[valid] >= 20).all()
    assert 100 < nulls < 320
    assert (chars == 0xC3).sum() > 200 and (chars == 95).sum() > 40 and ((chars >= 48) & (chars <= 57)).sum() > 1000
    c, o, v, nn = slice_rows(chars, offsets, validity, 5000, 9000)
    assert o[0] == 0 and o[-1] == c.size and nn == (~valid[5000:9000]).sum()
    assert bytes(c[o[7]:o[8]]) == bytes(chars[offsets[5007]:offsets[5008]])


def test_c2_match_rate_against_oracle(oracle):
    from custrings_b200.workloads import c2_corpus
    chars, offsets, validity, nulls = c2_corpus(20000, 20000 * 107, seed=6)
    res, cnt = oracle.RefStrings.from_arrays(chars, offsets, validity, nulls).contains_re(r"\b\w{4,}\b")
    assert 0.45 < cnt / 20000 < 0.55


def test_category_key_exchange_gloo_world2(tmp_path):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(ROOT, "tests", "dist_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    out = [json.load(open(tmp_path / ("rank%d.json" % k))) for k in range(2)]
    want = ["aaa", "ccc", None, "é", "b", "ccc", "zz" * 10]
    assert out[0]["merged"] == want and out[1]["merged"] == want
    assert out[0]["distinct"] == [None, "aaa", "b", "ccc", "zz" * 10, "é"] == out[1]["distinct"]
    assert out[0]["remap"] == [1, 3, 0, 5] and out[1]["remap"] == [2, 3, 4]
