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


# ---- the other BASELINE.json configs -------------------------------------------------------------------
import gzip as _gzip
import io as _io
import os as _os

_DATA = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tests", "golden", "data")


def _fixture(name):
    """bytes of one of the reference's sample data files (tests/golden/data/*.gz, made by tools/make_fixtures.py)"""
    with _gzip.open(_os.path.join(_DATA, name + ".gz"), "rb") as f:
        return f.read()


def pack_rows(rows):
    """list of bytes / None -> (chars uint8[], offsets int32[n+1], validity uint8[(n+7)//8], nulls)"""
    n = len(rows)
    lens = np.fromiter((0 if r is None else len(r) for r in rows), np.int64, n)
    offsets = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=offsets[1:])
    chars = np.frombuffer(b"".join(r for r in rows if r), np.uint8).copy() if offsets[-1] else np.zeros(0, np.uint8)
    valid = np.fromiter((r is not None for r in rows), bool, n)
    return chars, offsets.astype(np.int32), np.packbits(valid, bitorder="little"), int(n - valid.sum())


def c1_lines():
    """C1: data/985-rows.csv cut at '\\r' (header + 985 rows, every line holds exactly 11 commas), empties dropped"""
    return [l for l in _fixture("985-rows.csv").split(b"\r") if l]


def tile_rows(rows, total_bytes):
    """Round-robin tiling of `rows` (list of bytes) up to `total_bytes` chars -> (chars, offsets, n).  Vectorised: the
    pool is packed once and whole copies of it are repeated; the last copy is cut at a row boundary."""
    chars0, off0, _, _ = pack_rows(rows)
    per = int(off0[-1])
    reps = total_bytes // per
    tail_rows = int(np.searchsorted(off0, total_bytes - reps * per, side="right")) - 1
    n = reps * len(rows) + tail_rows
    chars = np.empty(reps * per + int(off0[tail_rows]), np.uint8)
    if reps:
        chars[: reps * per].reshape(reps, per)[:] = chars0
    chars[reps * per:] = chars0[: int(off0[tail_rows])]
    lens0 = np.diff(off0.astype(np.int64))
    lens = np.concatenate([np.tile(lens0, reps), lens0[:tail_rows]])
    offsets = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=offsets[1:])
    return chars, offsets.astype(np.int32), n


def c5_rows():
    """C5 row pool: the `text` column of data/tweets.csv (pandas parses the embedded quotes / newlines) with one line of
    data/utf8.csv (cut at '\\r') after every 8 tweets, so that multi-byte UTF-8 is actually exercised (tweets.csv holds
    only 163 non-ASCII bytes)."""
    import pandas as pd
    tweets = [t.encode("utf-8") for t in pd.read_csv(_io.BytesIO(_fixture("tweets.csv")))["text"].astype(str)]
    utf8 = [l for l in _fixture("utf8.csv").split(b"\r") if l.strip()]
    rows, u = [], 0
    for i, t in enumerate(tweets):
        rows.append(t)
        if i % 8 == 7:
            rows.append(utf8[u % len(utf8)])
            u += 1
    return rows


def c5_corpus(total_bytes=512 << 20):
    """C5 shard: c5_rows() tiled to `total_bytes` chars (512 MiB per GPU = 4 GiB over 8 GPUs)"""
    return tile_rows(c5_rows(), total_bytes)


DAYS = ("Sun", "Mon", "Tues", "Wed", "Thur", "Fri", "Sat")


def c3_readme_rows(n=10_000_000, seed=7):
    """C3 input A (README): n rows "%.2f,%.2f,{Female|Male},{Yes|No},{day},{Lunch|Dinner},%d" -> (chars, offsets)"""
    rng = np.random.Generator(np.random.PCG64(seed))
    pools = [
        np.array([b"%.2f" % v for v in np.arange(3.07, 50.81, 0.01)], object),
        np.array([b"%.2f" % v for v in np.arange(1.0, 10.0, 0.01)], object),
        np.array([b"Female", b"Male"], object), np.array([b"Yes", b"No"], object),
        np.array([d.encode() for d in DAYS], object), np.array([b"Lunch", b"Dinner"], object),
        np.array([b"%d" % v for v in range(1, 7)], object),
    ]
    # a vocabulary of complete rows would be too small; build rows from per-field byte tables instead
    fields = []
    for p in pools:
        idx = rng.integers(0, len(p), size=n)
        width = max(len(x) for x in p)
        tab = np.zeros((len(p), width), np.uint8)
        ln = np.zeros(len(p), np.int64)
        for i, x in enumerate(p):
            tab[i, : len(x)] = np.frombuffer(x, np.uint8)
            ln[i] = len(x)
        fields.append((tab, ln, idx))
    row_len = sum(ln[idx] for _, ln, idx in fields) + (len(fields) - 1)
    offsets = np.zeros(n + 1, np.int64)
    np.cumsum(row_len, out=offsets[1:])
    chars = np.full(int(offsets[-1]), ord(","), np.uint8)
    pos = offsets[:-1].copy()
    for k, (tab, ln, idx) in enumerate(fields):
        l = ln[idx]
        for j in range(tab.shape[1]):
            sel = l > j
            chars[pos[sel] + j] = tab[idx[sel], j]
        pos += l + 1
    return chars, offsets.astype(np.int32)


def c4_category_rows(n=12_500_000, nkeys=1000, seed=11, rank=0):
    """C4 shard: n rows drawn uniformly from `nkeys` distinct keys ([a-z0-9], length U{8..24}); the key set depends only on
    `seed`, the draw also on `rank` -> (chars, offsets, keys list)"""
    rng = np.random.Generator(np.random.PCG64(seed))
    alphabet = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz0123456789", np.uint8)
    keys = set()
    while len(keys) < nkeys:
        keys.add(alphabet[rng.integers(0, 36, size=int(rng.integers(8, 25)))].tobytes())
    keys = sorted(keys)
    rng2 = np.random.Generator(np.random.PCG64(seed * 1000 + 1 + rank))
    idx = rng2.integers(0, nkeys, size=n)
    width = 24
    tab = np.zeros((nkeys, width), np.uint8)
    ln = np.zeros(nkeys, np.int64)
    for i, k in enumerate(keys):
        tab[i, : len(k)] = np.frombuffer(k, np.uint8)
        ln[i] = len(k)
    l = ln[idx]
    offsets = np.zeros(n + 1, np.int64)
    np.cumsum(l, out=offsets[1:])
    chars = np.empty(int(offsets[-1]), np.uint8)
    for j in range(width):
        sel = l > j
        chars[offsets[:-1][sel] + j] = tab[idx[sel], j]
    return chars, offsets.astype(np.int32), keys, idx
