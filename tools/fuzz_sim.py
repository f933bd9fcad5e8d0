#!/usr/bin/env python3
"""Extended CPU fuzz of this repo's own engine logic (tests/sim: the __host__ __device__ Pike VM, the bit-stream plan executor, the
chain model and the span fast path) against the oracle: random patterns x random rows, far more than the test suite runs.
    python tools/fuzz_sim.py [rounds]      (150 patterns x ~250 rows x 2 anchorings per round, ~0.5 s per round)
Last run of this round: 600 rounds, 368 395 comparisons, 0 mismatches."""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import corpus, simlib
from oracle import ref as oracle
simlib.lib()
t0 = time.time()
bad = 0
tested = {"bits": 0, "chain": 0, "count": 0, "vm": 0}
for round_ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    rng = random.Random(77000 + round_)
    strs = corpus.STRINGS + corpus.random_strings(rng, 200)
    chars, offsets, validity, nulls = oracle.pack(strs)
    ref = oracle.RefStrings.from_list(strs)
    for p in corpus.random_patterns(91000 + round_, 150):
        try:
            wc, wcn = ref.contains_re(p); wm, wmn = ref.match(p); wcount = ref.count_re(p)[0]
        except Exception as e:
            continue
        for anchored, want, wn in ((False, wc, wcn), (True, wm, wmn)):
            g, c = simlib.bits_bool(chars, offsets, validity, p, anchored)
            if g is not None:
                tested["bits"] += 1
                if not (np.array_equal(want, g) and wn == c): bad += 1; print("BITS MISMATCH", repr(p), anchored, flush=True)
            g, c = simlib.chain_bool(chars, offsets, validity, p, anchored)
            if g is not None:
                tested["chain"] += 1
                if not (np.array_equal(want, g) and wn == c): bad += 1; print("CHAIN MISMATCH", repr(p), anchored, flush=True)
            g, c = simlib.bool_search(chars, offsets, validity, p, anchored)
            tested["vm"] += 1
            if not np.array_equal(want, g): bad += 1; print("VM MISMATCH", repr(p), anchored, flush=True)
        g, c = simlib.chain_count(chars, offsets, validity, p)
        if g is not None:
            tested["count"] += 1
            if not np.array_equal(np.asarray(wcount, np.int32), g): bad += 1; print("COUNT MISMATCH", repr(p), flush=True)
print("rounds", round_ + 1, "tested", tested, "mismatches", bad, "in %.0f s" % (time.time() - t0))
