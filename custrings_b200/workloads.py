"""Synthetic workloads of BASELINE.json, generated with vectorised numpy so that the 1 GiB C2
column is built in well under a minute on the host."""
import numpy as np


def c2_corpus(n_rows=10_000_000, total_bytes=1 << 30, seed=20240917):
    """C2: `n_rows` rows, exactly `total_bytes` chars.  Rows are lower-case words separated by single spaces; a fair
    coin picks short rows (word lengths U{1..3}) or long rows (U{1..8}) so that ~50% of rows match \\b\\w{4,}\\b;
    2% of rows carry one 2-byte 'é', ~1% of words are digits, 0.5% of rows get a '_' in place of a space, 1% of
    rows are null (validity bit 0, zero length).  Returns (chars uint8[total], offsets int32[n+1],
    validity uint8[(n+7)//8], nulls)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    valid = rng.random(n_rows) >= 0.01
    nn = int(valid.sum())
    lens = np.zeros(n_rows, np.int64)
    base = rng.integers(20, 196, size=nn)
    diff = total_bytes - int(base.sum())
    base += diff // nn
    base[: diff % nn] += 1
    if base.min() < 4:
        raise ValueError("total_bytes too small for n_rows")
    lens[valid] = base
    offsets = np.zeros(n_rows + 1, np.int64)
    np.cumsum(lens, out=offsets[1:])
    assert offsets[-1] == total_bytes
    is_long = rng.random(n_rows) < 0.5
    chars = np.empty(total_bytes, np.uint8)
    byte_long = np.repeat(is_long, lens)
    for kind, hi in ((False, 4), (True, 9)):
        sel = byte_long == kind
        t = int(sel.sum())
        stream = rng.integers(97, 123, size=t, dtype=np.uint8)
        avg = (1 + hi) / 2.0
        wl = rng.integers(1, hi, size=int(t / avg) + 16)
        sp = np.cumsum(wl + 1) - 1
        stream[sp[sp < t]] = 32
        chars[sel] = stream
        del stream, sel
    del byte_long
    starts = offsets[:-1]
    # ~1% of words become digits: pick word starts (byte after a space) and overwrite up to 8 letters
    cand = np.flatnonzero(chars[:-9] == 32)
    pick = cand[rng.random(cand.size) < 0.01] + 1
    alive = np.ones(pick.size, bool)
    digits = rng.integers(48, 58, size=(pick.size, 8), dtype=np.uint8)
    for j in range(8):
        alive &= chars[pick + j] != 32
        chars[(pick + j)[alive]] = digits[alive, j]
    # the digit pass must not run across a row boundary: re-plant nothing, rows are only byte ranges of the stream
    rows = np.flatnonzero(valid)
    # 2% of rows: one 'é' (0xC3 0xA9) at a random in-row position
    r = rows[rng.random(rows.size) < 0.02]
    pos = starts[r] + (rng.random(r.size) * (lens[r] - 1)).astype(np.int64)
    chars[pos] = 0xC3
    chars[pos + 1] = 0xA9
    # 0.5% of rows: '_' at a random in-row position (normally replacing a letter or a space)
    r = rows[rng.random(rows.size) < 0.005]
    pos = starts[r] + (rng.random(r.size) * lens[r]).astype(np.int64)
    ok = (chars[pos] < 0x80) & (chars[np.maximum(pos - 1, 0)] != 0xC3)
    chars[pos[ok]] = 95
    validity = np.packbits(valid, bitorder="little")
    return chars, offsets.astype(np.int32), validity, int(n_rows - nn)


def slice_rows(chars, offsets, validity, lo, hi):
    """Contiguous row range [lo, hi) as an independent (chars, offsets, validity, nulls) column (row sharding)."""
    off = offsets[lo:hi + 1].astype(np.int64)
    c = chars[off[0]:off[-1]]
    n = hi - lo
    v = np.unpackbits(validity, bitorder="little")[lo:hi]
    return c, (off - off[0]).astype(np.int32), np.packbits(v, bitorder="little"), int(n - v.sum())
