"""Pins the oracle: the reference compiled for the CPU (oracle/_ref) must reproduce every known-answer vector of the
reference's own test-suite for the hot path.  Runs without a GPU."""
import pytest

from tests.adapters import OracleAPI
from tests.golden.reference_kats import KATS


@pytest.mark.parametrize("idx", range(len(KATS)))
def test_reference_kat(oracle, idx):
    op, strs, args, want = KATS[idx]
    api = OracleAPI(oracle)
    got = getattr(api, op)(api.column(strs), *args)
    if isinstance(want, tuple):
        want = tuple(want)
        got = tuple(got)
    assert got == want, (op, args)
