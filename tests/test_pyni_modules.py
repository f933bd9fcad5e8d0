"""The CPython extension modules pyniNVStrings / pyniNVCategory / pyniNVText (custrings_b200/pyni): they load without a GPU,
export the hot-path n_* functions of the reference's method tables (python/cpp/pystrings.cpp:3860-3973, pycategory.cpp:900-937,
pytext.cpp:653-674), and — where the reference checkout exists — the reference's OWN python shims import against them
unmodified.  On a GPU (-m gpu) the n_* entry points are driven with the reference shims' argument conventions and checked
against the oracle."""
import importlib
import os
import random
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PYNI = os.path.join(ROOT, "custrings_b200", "pyni")

STRINGS_API = ["n_createFromHostStrings", "n_destroyStrings", "n_createHostStrings", "n_createFromOffsets", "n_create_offsets", "n_size", "n_len",
               "n_byte_count", "n_set_null_bitmask", "n_null_count", "n_hash", "n_copy", "n_gather", "n_contains", "n_match", "n_count", "n_replace",
               "n_replace_multi", "n_replace_with_backrefs", "n_find", "n_rfind", "n_find_from", "n_startswith", "n_endswith", "n_match_strings",
               "n_find_multiple", "n_split", "n_rsplit", "n_split_record", "n_rsplit_record", "n_partition", "n_rpartition", "n_findall",
               "n_findall_record", "n_extract", "n_extract_record"]
CATEGORY_API = ["n_createCategoryFromNVStrings", "n_createCategoryFromHostStrings", "n_destroyCategory", "n_size", "n_keys_size", "n_keys_type",
                "n_get_keys", "n_get_values", "n_get_values_cpointer", "n_to_strings", "n_merge_category", "n_merge_and_remap"]
TEXT_API = ["n_tokenize", "n_token_count"]


def _mods():
    if PYNI not in sys.path:
        sys.path.insert(0, PYNI)
    return [importlib.import_module(m) for m in ("pyniNVStrings", "pyniNVCategory", "pyniNVText")]


def test_modules_load_and_export():
    ps, pc, pt = _mods()
    for mod, names in ((ps, STRINGS_API), (pc, CATEGORY_API), (pt, TEXT_API)):
        missing = [n for n in names if not callable(getattr(mod, n, None))]
        assert not missing, (mod.__name__, missing)


@pytest.mark.skipif(not os.path.isdir("/root/reference/python"), reason="reference checkout not present")
def test_reference_python_shims_import_unmodified():
    code = ("import nvstrings, nvcategory, nvtext, pyniNVStrings\n"
            "assert nvstrings.__file__.startswith('/root/reference/python/')\n"
            "assert pyniNVStrings.__file__.startswith(%r)\n"
            "print('ok', all(hasattr(nvstrings, n) for n in ('to_device', 'from_offsets', 'free', 'bind_cpointer')))\n" % PYNI)
    env = dict(os.environ, PYTHONPATH=PYNI + ":/root/reference/python")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and "ok True" in r.stdout, r.stderr[-2000:]


@pytest.mark.gpu
def test_pyni_entry_points_against_oracle(oracle):
    """every call below is written the way the reference's python/nvstrings.py / nvcategory.py / nvtext.py write it"""
    ps, pc, pt = _mods()
    from custrings_b200 import nvstrings as shim
    from tests import corpus
    rng = random.Random(3)
    strs = corpus.STRINGS + corpus.random_strings(rng, 300)
    ref = oracle.RefStrings.from_list(strs)
    h = ps.n_createFromHostStrings(strs)
    try:
        assert ps.n_size(h) == len(strs) and ps.n_createHostStrings(h) == strs
        nn = lambda vals: [None if s is None else v for s, v in zip(strs, vals)]  # noqa: E731
        assert ps.n_null_count(h, False) == sum(s is None for s in strs)
        assert ps.n_len(h, 0) == nn([len(s or "") for s in strs])
        for pat in (r"\b\w{4,}\b", r"\d+", "é"):
            assert ps.n_contains(h, pat, True, 0) == nn([bool(x) for x in ref.contains_re(pat)[0]])
            assert ps.n_match(h, pat, 0) == nn([bool(x) for x in ref.match(pat)[0]])
            assert ps.n_count(h, pat, 0) == nn([int(x) for x in ref.count_re(pat)[0]])
        assert ps.n_contains(h, "fox", False, 0) == nn([bool(x) for x in ref.contains("fox")[0]])
        assert ps.n_find(h, "o", 0, None, 0) == nn([int(x) for x in ref.find("o")[0]])
        assert ps.n_rfind(h, "o", 0, None, 0) == nn([int(x) for x in ref.rfind("o")[0]])
        assert ps.n_startswith(h, "the", 0) == nn([bool(x) for x in ref.startswith("the")[0]])
        assert ps.n_hash(h, 0) == [int(x) for x in ref.hash()[0]]

        def host(hh):
            try:
                return ps.n_createHostStrings(hh)
            finally:
                ps.n_destroyStrings(hh)

        dec = lambda r: [None if x is None else x.decode() for x in r.to_list()]  # noqa: E731
        assert host(ps.n_replace(h, r"\d+", "#", -1, True)) == dec(ref.replace_re(r"\d+", "#"))
        assert host(ps.n_replace(h, "the", "THE", -1, False)) == dec(ref.replace("the", "THE"))
        assert host(ps.n_replace_with_backrefs(h, r"(\w)(\d)", r"\2\1")) == dec(ref.replace_with_backrefs(r"(\w)(\d)", r"\2\1"))
        assert [host(c) for c in ps.n_split(h, " ", -1)] == [dec(c) for c in ref.split(" ")]
        assert [host(c) for c in ps.n_rsplit(h, None, 2)] == [dec(c) for c in ref.split(None, 2, right=True)]
        rec = ps.n_split_record(h, ",", -1)
        want = ref.split_record(",")[0]
        assert [None if c == 0 else host(c) for c in rec] == [None if w is None else dec(w) for w in want]
        part = ps.n_partition(h, " ")
        wantp = ref.partition(" ")[0]
        assert [None if c == 0 else host(c) for c in part] == [None if w is None else dec(w) for w in wantp]
        assert [host(c) for c in ps.n_findall(h, r"\d+")] == [dec(c) for c in ref.findall(r"\d+")]
        assert [host(c) for c in ps.n_extract(h, r"(\w)(\d)")] == [dec(c) for c in ref.extract(r"(\w)(\d)")]
        # category / text take the nvstrings OBJECT and read its m_cptr (type name "nvstrings")
        obj = shim.bind_cpointer(h, own=False)
        cat = pc.n_createCategoryFromNVStrings(obj)
        rcat = oracle.RefCategory(ref)
        assert pc.n_keys_size(cat) == rcat.keys_size() and pc.n_size(cat) == len(strs)
        assert host(pc.n_get_keys(cat)) == dec(rcat.keys())
        assert pc.n_get_values(cat, 0) == rcat.values().tolist()
        assert host(pc.n_to_strings(cat)) == strs
        pc.n_destroyCategory(cat)
        assert host(pt.n_tokenize(obj, None)) == dec(ref.tokenize())
        assert pt.n_token_count(obj, None, 0) == [int(x) for x in ref.token_count()[0]]
    finally:
        ps.n_destroyStrings(h)
