#!/usr/bin/env python3
"""Copies the reference's public sample DATA files that BASELINE.json's configs name (C1: data/985-rows.csv, C5:
data/tweets.csv + data/utf8.csv) into tests/golden/data/ (gzip, deterministic mtime) so that bench.py and the -m gpu
tests can use them on the GPU box, where /root/reference does not exist.  Data only — no reference source is copied.
    python tools/make_fixtures.py [/root/reference]"""
import gzip
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "data")
os.makedirs(OUT, exist_ok=True)
for name in ("985-rows.csv", "tweets.csv", "utf8.csv"):
    raw = open(os.path.join(REF, "data", name), "rb").read()
    with open(os.path.join(OUT, name + ".gz"), "wb") as f:
        with gzip.GzipFile(filename="", mode="wb", fileobj=f, mtime=0, compresslevel=9) as g:
            g.write(raw)
    print(name, len(raw), "->", os.path.getsize(os.path.join(OUT, name + ".gz")))
