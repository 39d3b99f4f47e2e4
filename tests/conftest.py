import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_case(npz, name):
    return {k.split(".", 1)[1]: npz[k] for k in npz.files if k.startswith(name + ".")}


# ------------------------------------------------------------------------------------- tie-aware index comparison
def neighbour_d2(queries, supports, idx):
    """f32 squared distances of every entry of an index matrix in the reference's arithmetic ((dx*dx + dy*dy) + dz*dz,
    d = query - support, products rounded individually: nanoflann.hpp:432-440); shadow entries (== Ns) get +inf."""
    import numpy as np
    q = np.asarray(queries, np.float32)
    s = np.asarray(supports, np.float32)
    idx = np.asarray(idx, np.int64)
    shadow = idx >= len(s)
    j = np.where(shadow, 0, idx)
    d = q[:, None, :] - s[j]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    assert d2.dtype == np.float32
    return np.where(shadow, np.float32(np.inf), d2)


def canonical_rows(queries, supports, idx):
    """Rows re-sorted by the stated total order (d2 ascending, support index ascending): identical for any two
    matrices that differ only by permutations inside groups of exactly equal d2."""
    import numpy as np
    d2 = neighbour_d2(queries, supports, idx)
    key = (d2.view(np.uint32).astype(np.uint64) << np.uint64(32)) | np.asarray(idx, np.int64).astype(np.uint64)
    order = np.argsort(key, axis=1, kind="stable")
    return np.take_along_axis(np.asarray(idx, np.int64), order, 1), np.take_along_axis(d2, order, 1)


def assert_same_up_to_ties(queries, supports, got, ref, what=""):
    """``got`` (our (d2, index)-ordered matrix) against ``ref`` (the reference's: std::sort on d2 alone, unstable):
    after canonicalising exact-d2 tie groups the matrices must be IDENTICAL, except where a tie group straddles the last
    column of a cropped, completely filled row (big_neighborhood_filter keeps an arbitrary member of that group): there
    the differing members must all sit at exactly the row's largest kept d2. Anything else is a real mismatch."""
    import numpy as np
    got, ref = np.asarray(got, np.int64), np.asarray(ref, np.int64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    if got.size == 0:
        return 0, 0
    cg, dg = canonical_rows(queries, supports, got)
    cr, dr = canonical_rows(queries, supports, ref)
    assert np.array_equal(cg, got), f"{what}: our rows are not in (d2, index) order"
    assert np.array_equal(dg, dr), f"{what}: sorted distance profiles differ (membership mismatch beyond ties)"
    bad = np.nonzero((cg != cr).any(1))[0]
    for i in bad:
        last = dg[i, -1]
        assert np.isfinite(last), f"{what} row {i}: differs although the row is not full"
        diff = cg[i] != cr[i]
        assert (dg[i][diff] == last).all() and (dr[i][diff] == last).all(), f"{what} row {i}: differs outside the crop tie group"
    n_perm = int(((got != ref).any(1)).sum())
    return n_perm, len(bad)
