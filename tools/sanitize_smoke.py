#!/usr/bin/env python3
"""Small end-to-end exercise of every kernel family, meant to run under compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from custrings_b200 import nvstrings, nvcategory, nvtext  # noqa: E402

rng = random.Random(4)
words = ["abcd", "x", "héllo", "12345", "a_b", "日本語", "wörld9", "_", "zz😀zz", "the", "fox", "a,b", "", "colour", "color"]
strs = []
for i in range(6000):
    r = rng.random()
    if r < 0.03:
        strs.append(None)
    elif r < 0.06:
        strs.append("")
    elif r < 0.063:
        strs.append(" ".join(rng.choice(words) for _ in range(rng.randrange(500, 3000))))
    else:
        strs.append(rng.choice([" ", ",", "  "]).join(rng.choice(words) for _ in range(rng.randrange(0, 25))))
strs[17] = "nul\x00byte row abcd"
col = nvstrings.to_device(strs)
n = col.size()
for pat in [r"\b\w{4,}\b", r"\d+", r"\w+", r"\bthe\b|\bfox\b", r"colou?r", r"\d{1,3}", r"(a|b)c", r"[a-z]+@", r"é+", r"^\w+ \w+"]:
    a = col.contains(pat)
    b = col.match(pat)
    c = col.count(pat)
    d = col.replace(pat, "<#>")
    e = col.replace(pat, "", 1)
    assert len(a) == len(b) == len(c) == n and d.size() == n and e.size() == n
col.replace_with_backrefs(r"(\w)(\w)", r"\2\1")
col.findall(r"\d+")
col.extract(r"(\w+) (\w+)")
for lit in ["abcd", "é", " ", "zz"]:
    col.contains(lit, regex=False)
    col.find(lit)
    col.rfind(lit)
    col.replace(lit, "Q", regex=False)
for d in (None, " ", ","):
    col.split(d, 3)
    col.rsplit(d, 3)
    col.split_record_flat(d)
    col.split_record_flat(d, 2, right=True)
col.partition_flat(" ")
col.partition_flat(",", right=True)
for d in (None, " ,", "a"):
    nvtext.tokenize(col, d)
    nvtext.token_count(col, d)
cat = nvcategory.from_strings(col)
cat2 = nvcategory.from_strings(nvstrings.to_device([s for s in strs[:500]]))
cat.merge_and_remap(cat2)
cat.merge_category(cat2)
cat.to_strings()
col.hash()
col.len()
print("sanitize smoke ok:", n, "rows,", cat.keys_size(), "keys")
