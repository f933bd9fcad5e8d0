"""The reference's own known-answer vectors replayed through the C-ABI on the GPU."""
import pytest

from tests.golden.reference_kats import KATS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("idx", range(len(KATS)))
def test_product_kat(idx):
    from tests.adapters import ProductAPI
    op, strs, args, want = KATS[idx]
    api = ProductAPI()
    got = getattr(api, op)(api.column(strs), *args)
    if isinstance(want, tuple):
        want, got = tuple(want), tuple(got)
    assert got == want, (op, args)
