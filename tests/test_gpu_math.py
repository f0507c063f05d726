"""Device arithmetic shortcuts must reproduce the plain IEEE results bit for bit (they sit on the
path that decides which blocks get allocated)."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("which", [0, 1])
def test_selftest_has_no_mismatch(gpu_ctx, which):
    from coxgraph_b200 import capi
    bad = C.c_uint64(123)
    capi.check(capi.load().cg_debug_selftest(gpu_ctx._h, which, 2_000_000_000, C.byref(bad)))
    assert bad.value == 0
