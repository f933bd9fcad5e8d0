"""Worker for the world_size-2 gloo test of the multi-GPU host logic (run by torchrun from test_host_logic.py)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch.distributed as dist
    from custrings_b200.nvcategory import exchange_key_arrays
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    keys = [["aaa", "ccc", None, "é"], ["b", "ccc", "zz" * 10]][rank]
    enc = [b"" if k is None else k.encode() for k in keys]
    chars = np.frombuffer(b"".join(enc), np.uint8)
    offsets = np.zeros(len(keys) + 1, np.int32)
    np.cumsum([len(e) for e in enc], out=offsets[1:])
    valid = np.array([k is not None for k in keys])
    c, o, v = exchange_key_arrays(chars, offsets, valid)
    raw = c.tobytes()
    merged = [raw[o[i]:o[i + 1]].decode() if v[i] else None for i in range(len(o) - 1)]
    # global dictionary every rank would build + remap of its local value indices
    distinct = sorted({m for m in merged if m is not None}, key=lambda s: s.encode())
    if any(m is None for m in merged):
        distinct = [None] + distinct
    remap = [distinct.index(k) for k in keys]
    with open(os.path.join(sys.argv[1], "rank%d.json" % rank), "w") as f:
        json.dump({"merged": merged, "distinct": distinct, "remap": remap, "world": world}, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
