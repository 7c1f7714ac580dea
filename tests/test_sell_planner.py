"""The host planner of the sliced-ELL trajectory kernel (csrc/traj_sell.cu), run without a GPU through
oqb_sell_plan_probe: every entry of the operator placed exactly once on the lane of its row, σ a permutation, and no
LDS.128 phase (8 lanes × one slot) in which two lanes gather from the same 16-byte bank group — the property the kernel's
shared-memory throughput rests on (profiles/r2_traj_sell.md)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sps

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def probe(mats, unit=2):
    import oqb_b200  # noqa: F401
    from oqb_b200 import _lib
    lib = _lib.load()
    n = mats[0].shape[0]
    keep, cps, rvs = [], (C.POINTER(C.c_int64) * len(mats))(), (C.POINTER(C.c_int32) * len(mats))()
    for t, m in enumerate(mats):
        m = sps.csc_matrix(m)
        cp = np.ascontiguousarray(m.indptr, dtype=np.int64)
        rv = np.ascontiguousarray(m.indices, dtype=np.int32)
        keep += [cp, rv]
        cps[t] = cp.ctypes.data_as(C.POINTER(C.c_int64))
        rvs[t] = rv.ctypes.data_as(C.POINTER(C.c_int32))
    pad, ph, cf, ident, bad = C.c_double(), C.c_int64(), C.c_int64(), C.c_int(), C.c_int()
    st = lib.oqb_sell_plan_probe(n, len(mats), cps, rvs, unit, C.byref(pad), C.byref(ph), C.byref(cf), C.byref(ident), C.byref(bad))
    assert st == 0
    return {"padding": pad.value, "phases": ph.value, "conflicting": cf.value, "identity": bool(ident.value), "misplaced": bad.value}


def random_pattern(n, per_row, seed, ragged=False):
    rng = np.random.default_rng(seed)
    rows = np.repeat(np.arange(n), per_row)
    cols = rng.integers(0, n, size=n * per_row)
    keep = rows != cols
    if ragged:
        keep &= rng.random(n * per_row) < (0.1 + 0.9 * rows / n)
    A = sps.csc_matrix((np.ones(keep.sum()), (rows[keep], cols[keep])), shape=(n, n))
    return (A + A.T).tocsc()


@pytest.mark.parametrize("n", [32, 100, 257, 1000, 4096])
@pytest.mark.parametrize("unit", [1, 2])
def test_random_patterns_are_placed_exactly_and_conflict_free(n, unit):
    r = probe([random_pattern(n, 6, n + unit)], unit)
    assert r["misplaced"] == 0 and r["conflicting"] == 0 and r["phases"] > 0
    assert 1.0 <= r["padding"] < 1.6


def test_c4_sized_random_pattern_padding():
    """The benchmark operator (profiles/r2_traj_sell.md): 4096 rows, ≈13 entries per row — stored ÷ real entries."""
    r1, r2 = probe([random_pattern(4096, 6, 4100)], 1), probe([random_pattern(4096, 6, 4100)], 2)
    assert r1["misplaced"] == r2["misplaced"] == 0 and r1["conflicting"] == r2["conflicting"] == 0
    assert r1["padding"] < 1.25 and r2["padding"] < 1.30 and not r2["identity"]


@pytest.mark.parametrize("nq", [5, 8, 12])
def test_tfim_driver_keeps_its_row_order_without_padding(nq):
    n = 1 << nq
    idx = np.arange(n)
    rows = np.concatenate([idx ^ (1 << q) for q in range(nq)])
    H = sps.csc_matrix((np.ones(n * nq), (rows, np.tile(idx, nq))), shape=(n, n))
    r = probe([H])
    assert r == {"padding": 1.0, "phases": r["phases"], "conflicting": 0, "identity": True, "misplaced": 0}


def test_ragged_rows_two_terms_and_a_dense_row():
    n = 512
    A, B = random_pattern(n, 7, 1, ragged=True), random_pattern(n, 3, 2)
    M = sps.lil_matrix((n, n))
    M[11, :] = 1.0
    M[:, 5] = 1.0
    r = probe([A, B, sps.csc_matrix(M)])
    assert r["misplaced"] == 0 and r["conflicting"] == 0


def test_rejects_bad_arguments():
    from oqb_b200 import _lib
    lib = _lib.load()
    pad, ph, cf, ident, bad = C.c_double(), C.c_int64(), C.c_int64(), C.c_int(), C.c_int()
    assert lib.oqb_sell_plan_probe(16, 1, None, None, 2, C.byref(pad), C.byref(ph), C.byref(cf), C.byref(ident), C.byref(bad)) != 0
