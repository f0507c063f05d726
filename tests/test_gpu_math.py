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


def test_direct_raycaster_state_matches_the_walk(gpu_ctx):
    """Groundwork for a walk without the serial DDA: the RayCaster state right
    after every step that crosses a block face, computed directly (closed-form float accumulation
    + a binary search over the other axes), equals the sequential walk bit for bit."""
    from coxgraph_b200 import capi
    lib = capi.load()
    bad, seen = C.c_uint64(123), C.c_uint64(0)
    capi.check(lib.cg_debug_selftest(gpu_ctx._h, 2, 2_000_000, C.byref(bad)))
    capi.check(lib.cg_debug_selftest(gpu_ctx._h, 3, 2_000_000, C.byref(seen)))
    assert seen.value > 5_000_000, "the self test compared too few block entries"
    assert bad.value == 0
