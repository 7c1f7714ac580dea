"""CPU tests of the host-side logic (no GPU, no compute calls into the CUDA library):
 * the C-ABI library loads and exports every symbol include/oqb.h declares;
 * the product's gap preprocessing (gaps.py) reproduces the oracle's literal GapIndices grouping,
   and the closed-form Davies tables it feeds to the device (emulated here in numpy, operation for
   operation as csrc/elementwise.cu does them) equal the oracle's per-gap-group loop;
 * the product's host quadrature equals the oracle's QuadGK mirror;
 * creating a context without a GPU fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import oqb_b200
from oqb_b200 import _lib
from oqb_b200.gaps import GapGroups, round_sigdigits
from oqb_b200.quadrature import quadgk_matrix
from oracle import oqb_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "oqb.h")).read()
    declared = set(re.findall(r"\b(oqb_[a-z0-9_]+)\s*\(", hdr))
    lib = _lib.load()
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/oqb.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.oqb_version() == 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    assert _lib.load().oqb_create(0, ctypes.byref(h)) != 0
    with pytest.raises(_lib.OQBError):
        oqb_b200.Context(0)


def test_round_sigdigits_matches_oracle():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(2000) * 10.0 ** rng.integers(-6, 6, 2000),
                        [8.05669612345, 12.9873017464, 1e-7, 123456789.0, 0.5, 9.99999995, 99999999.5]])
    ref = np.array([O.round_sigdigits(float(v), 8) for v in x])
    assert np.array_equal(round_sigdigits(x, 8), ref)


def _mock_gamma(x):
    return x + 1 if x >= 0 else (1 - x) * np.exp(x)


def _mock_S(x):
    return x + 0.1


def closed_form_numpy(w, A_list, groups, gamma, S, rho):
    """numpy emulation of the device Davies path (davies_tables + hadamard + davies_diag + davies_cross +
    davies_lamb_build in csrc/elementwise.cu) — eigenbasis, includes −i[diag w, ρ]."""
    l = len(w)
    gam, Sm, crate, lrs = groups.tables(gamma, S)
    m2 = sum(np.abs(A) ** 2 for A in A_list)
    Gam = (gam * m2).sum(axis=0)
    hls = (Sm * m2).sum(axis=0)
    W = gam * m2
    np.fill_diagonal(W, 0.0)
    dA = sum(np.outer(np.diag(A), np.diag(A).conj()) for A in A_list)
    Cm = -0.5 * (Gam[:, None] + Gam[None, :]) - 1j * (hls[:, None] - hls[None, :]) \
        - 1j * (w[:, None] - w[None, :]) + gam[0, 0] * dA
    X = Cm * rho
    X[np.diag_indices(l)] += W @ np.diag(rho)
    for (a, b, a2, b2), r in zip(groups.cross_idx, crate):
        X[a, a2] += r * sum(A[a, b] * np.conj(A[a2, b2]) for A in A_list) * rho[b, b2]
    if len(groups.lamb_idx):
        M = np.zeros((l, l), complex)
        for (a, b, b2), (r, s) in zip(groups.lamb_idx, lrs):
            M[b, b2] += (-0.5 * r - 1j * s) * sum(np.conj(A[a, b]) * A[a, b2] for A in A_list)
        X += M @ rho + rho @ M.conj().T
    return X


@pytest.mark.parametrize("spectrum", ["generic", "equal_spaced", "degenerate", "unsorted", "ref_util_test"])
def test_gap_groups_closed_form_equals_literal_loop(spectrum):
    rng = np.random.default_rng(11)
    w = {"generic": np.sort(rng.standard_normal(8) * 3),
         "equal_spaced": np.arange(7) * 0.75 - 2.0,
         "degenerate": np.array([-2.0, -2.0, -0.5, 0.3, 0.3, 0.3, 1.9]),
         "unsorted": np.array([0.4, -1.0, 2.0, 1.9, 1.0, 0.5, 0.0]),
         "ref_util_test": np.array([-3, 1, 3, 3, 4.5, 5.5, 8.0])}[spectrum]  # reference test/opensys/util.jl:3
    l = len(w)
    A_list = [(lambda g: g + g.conj().T)(rng.standard_normal((l, l)) + 1j * rng.standard_normal((l, l))) for _ in range(2)]
    rho = rng.standard_normal((l, l)) + 1j * rng.standard_normal((l, l))
    groups = GapGroups(w, 8, 8)
    gi = O.build_gap_indices(w, 8, 8)
    # the signed rounded gap matrix agrees with the oracle's literal grouping
    om = np.zeros((l, l))
    for key, aa, bb in gi.positive_gap_indices():
        for a, b in zip(aa, bb):
            om[a - 1, b - 1] = key
            om[b - 1, a - 1] = -key
    assert np.array_equal(groups.omega, om)
    # closed form == literal loop (davies.jl:21-47) + Hamiltonian commutator
    dav = O.DaviesGenerator(A_list, _mock_gamma, _mock_S)
    ref = -1j * (np.diag(w) @ rho - rho @ np.diag(w))
    ref = dav.rhs_acc(ref.astype(complex), rho, gi, None, 0.0)
    got = closed_form_numpy(w, A_list, groups, _mock_gamma, _mock_S, rho)
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-13


def test_host_quadrature_matches_oracle_mirror():
    import scipy.linalg as sla
    sx = np.array([[0, 1], [1, 0]], complex)
    sz = np.array([[1, 0], [0, -1]], complex)
    U = lambda t: sla.expm(-1j * t * sx)
    f = lambda x: np.cos(3 * x) * (U(5.0) @ U(x).conj().T @ sz @ U(x) @ U(5.0).conj().T)
    a = quadgk_matrix(f, 0.0, 5.0, 1e-6, 1e-8)
    b, _ = O.quadgk(f, 0.0, 5.0, 1e-6, 1e-8)
    assert np.linalg.norm(a - b) < 1e-14


def test_gk15_node_tables_and_sampled_unitary():
    """redfield_device's 15-node tables reproduce the oracle's evalrule on a scalar integrand, and the sampled
    unitary closure interpolates linearly (QuadGK.kronrod(7))."""
    from oqb_b200.redfield_device import NODE_WG, NODE_WK, NODE_XI, SampledUnitary
    assert abs(NODE_WK.sum() - 2.0) < 1e-15 and abs(NODE_WG.sum() - 2.0) < 1e-15
    assert np.count_nonzero(NODE_WG) == 7 and np.all(NODE_WG[:7] > 0) and len(set(NODE_XI)) == 15  # Gauss nodes first
    f = lambda x: np.exp(-x) * np.cos(3 * x)
    a, b = 0.3, 1.7
    h = 0.5 * (b - a)
    x = a + (1 + NODE_XI) * h
    Ik, Ig = (NODE_WK * f(x)).sum() * h, (NODE_WG * f(x)).sum() * h
    I_or, E_or = O._evalrule(f, a, b)
    assert abs(Ik - I_or) < 1e-15 and abs(abs(Ik - Ig) - E_or) < 1e-15
    U = SampledUnitary(np.stack([np.eye(2) * g for g in range(5)]), 0.5)
    assert np.allclose(U(0.75), 1.5 * np.eye(2)) and np.allclose(U(2.0), 4 * np.eye(2))
    g, th = SampledUnitary.locate(np.array([0.0, 0.74, 2.0, 2.5]), 0.5, 5)
    assert list(g) == [0, 1, 4, 4] and th[2] == 0.0 and th[3] == 0.0


def test_jump_slots_follow_the_reference_order():
    """gaps.jump_slots (sort-based) lists ame_jump's probability slots exactly as the literal GapIndices loop does
    (src/opensys/opensys_util.jl:25-61, src/opensys/trajectory_jump.jl:22-44), including degenerate groups."""
    from oqb_b200.gaps import jump_slots
    rng = np.random.default_rng(3)
    for w in (np.sort(rng.standard_normal(7)), np.array([0.0, 1.0, 1.0, 2.0, 3.0, 3.0 + 1e-12, 5.0]),
              np.array([-1.0, -0.5, 0.0, 0.5, 1.0])):
        K = 2
        G = O.build_gap_indices(w, 8, 8)
        ptr, terms, rkey, tags = jump_slots(w, K, 8, 8)
        exp = []
        for (gw, a, b) in G.positive_gap_indices():
            for i in range(K):
                exp.append((i, list(np.array(a) - 1), list(np.array(b) - 1), gw))
                exp.append((i, list(np.array(b) - 1), list(np.array(a) - 1), -gw))
        a0, b0 = G.zero_gap_indices()
        for i in range(K):
            exp.append((i, list(np.array(a0) - 1), list(np.array(b0) - 1), 0.0))
        assert len(exp) == len(rkey) == len(tags) == len(ptr) - 1
        for s, (i, rows, cols, kv) in enumerate(exp):
            ti, tr, tc = tags[s]
            assert ti == i and list(tr) == rows and list(tc) == cols and rkey[s] == kv
            seg = terms[ptr[s]:ptr[s + 1]]
            assert sorted(zip(rows, cols)) == [(int(r), int(c)) for _, r, c in seg] and all(t[0] == i for t in seg)


def test_c_abi_header_is_plain_c_and_links(c_demo_exe):
    """include/oqb.h is a C header (no C++/torch types) and every symbol the demo uses resolves at link time; without a GPU
    the program stops at oqb_create with a clear message and a non-zero exit code (no CPU fallback)."""
    import subprocess
    import torch
    exe = c_demo_exe
    if not torch.cuda.is_available():
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 3 and "no CPU fallback" in out.stderr


def test_c_only_examples_compile(c_liouv_exe, c_comm_exe):
    """The C-only AME and NCCL demos build against include/oqb.h with -Wall -Werror; without a GPU they stop at oqb_create."""
    import subprocess
    import torch
    if not torch.cuda.is_available():
        for exe in (c_liouv_exe, c_comm_exe):
            out = subprocess.run([exe], capture_output=True, text=True)
            assert out.returncode == 3 and "no CPU fallback" in out.stderr


def test_gridded_linear_lambshift_table():
    """build_lambshift on a NON-uniform ω grid (src/bath/bath_util.jl:26-38 → construct_interpolations for a non-range grid,
    src/interpolation.jl:36-65: Gridded(Linear()), fill value 0): host mirror against the oracle's scalar restatement, and
    the reference's linear-data known answers (test/interpolations.jl:3-20: exact on linear data, 0 outside the grid)."""
    from oqb_b200 import host as H
    x = np.array([0.0, 0.4, 1.0, 2.5, 2.6, 4.0, 7.5, 10.0])
    g, o = H.GriddedLinear(x, 10 - x), O.GriddedLinear(x, 10 - x)
    tx = np.linspace(0, 10, 77)
    assert np.allclose(g.vectorized(tx), 10 - tx, rtol=0, atol=1e-14)
    assert g(-1) == 0 and g(11) == 0 and o(-1) == 0 and o(11) == 0
    rng = np.random.default_rng(8)
    y = rng.standard_normal(len(x))
    g, o = H.GriddedLinear(x, y), O.GriddedLinear(x, y)
    pts = np.r_[x, rng.uniform(-1, 11, 200)]
    assert np.allclose(g.vectorized(pts), [o(p) for p in pts], rtol=0, atol=1e-15)

    class Piecewise:  # the spectrum of test/coupling_bath_interaction/bath.jl:72-80
        def spectrum(self, w):
            return np.exp(-abs(w)) if w >= 0 else np.exp(-abs(w)) * 0.5
    S_o = O.build_lambshift(x - 3.0, True, Piecewise())
    assert isinstance(S_o, O.GriddedLinear)


def test_quadratic_bspline_host_coefficients():
    """Host mirror of construct_interpolations (interpolation.jl:22-34): banded prefilter solve and the vectorised
    evaluation agree with the oracle's dense restatement; linear data is reproduced (test/interpolations.jl:3-17)."""
    from oqb_b200 import host as H
    x = np.linspace(0, 10, 100)
    sp, so = H.QuadraticBSpline(0.0, 10 / 99, 10 - x), O.QuadraticBSpline(0.0, 10 / 99, 10 - x)
    assert np.allclose(sp.coef, so.c, rtol=0, atol=1e-13)
    tx = np.linspace(0, 10, 50)
    assert np.allclose(sp.vectorized(tx), 10 - tx) and sp(-1) == 0 and sp(11) == 0
    rng = np.random.default_rng(5)
    y = rng.standard_normal(37)
    sp, so = H.QuadraticBSpline(-2.0, 0.25, y), O.QuadraticBSpline(-2.0, 0.25, y)
    xs = rng.uniform(-3, 8, 200)
    assert np.max(np.abs(sp.vectorized(xs) - np.array([so(v) for v in xs]))) < 1e-13


def test_lambshift_host_closure_known_answer():
    """lambshift of a spectrum given as a host closure (ext_util.jl:23-28) and its tabulated spline
    (bath_util.jl:26-38): test/coupling_bath_interaction/bath.jl:58-80 expects 0.03551 (atol 1e-4) from both."""
    from oqb_b200 import host as H
    gfun = lambda w: np.exp(-w) if w >= 0 else np.exp(0.8 * w)
    assert abs(H.lambshift(0.0, gfun) - 0.03551) < 1e-4
    assert abs(H.lambshift(0.3, gfun) - O.lambshift(0.3, gfun)) < 1e-12
    assert H.build_lambshift([0.0], False, gfun)(0.5) == 0
    S_tab = H.build_lambshift(np.linspace(-5, 5, 200), True, gfun)
    assert isinstance(S_tab, H.QuadraticBSpline) and abs(S_tab(0.0) - 0.03551) < 2e-3  # coarser table than the reference's
    assert abs(H.build_lambshift([], True, gfun)(0.0) - 0.03551) < 1e-4


def test_vectorised_cross_lists_equal_the_loop():
    """gaps._group_cross_lists / _zero_group_lists (numpy) produce exactly the lists — entries AND order — of the plain
    double loop GapGroups._pairs that states the semantics (davies.jl:25-44 restricted to multi-pair groups)."""
    from oqb_b200 import gaps as G
    rng = np.random.default_rng(11)
    for trial in range(6):
        l = [5, 8, 12, 16, 9, 7][trial]
        if trial % 2 == 0:
            w = np.sort(rng.integers(-3, 4, l) * 0.5 + (rng.integers(0, 2, l) * 0.25))  # many equal gaps and equal levels
        else:
            w = np.arange(l) * 0.75 + (rng.random(l) < 0.3) * 0.75
        gg = G.GapGroups(w, 8, 8)
        # the loop, group by group
        iu, ju = np.triu_indices(l, 1)
        gap = w[ju] - w[iu]
        zero = np.abs(gap) <= 1e-8
        key = np.zeros_like(gap)
        key[~zero] = G.round_sigdigits(gap[~zero], 8)
        cross, ck, lamb, lk = [], [], [], []
        nzi, nzj, nzk = iu[~zero], ju[~zero], key[~zero]
        order = np.argsort(nzk, kind="stable")
        ks = nzk[order]
        starts = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]]) if ks.size else np.zeros(0, dtype=int)
        ends = np.r_[starts[1:], ks.size] if ks.size else np.zeros(0, dtype=int)
        for a, b in zip(starts, ends):
            if b - a < 2:
                continue
            mem = order[a:b]
            for (aa, bb, kv) in ((nzi[mem], nzj[mem], float(ks[a])), (nzj[mem], nzi[mem], -float(ks[a]))):
                G.GapGroups._pairs(aa, bb, kv, cross, ck, lamb, lk)
        zi, zj = iu[zero], ju[zero]
        if zi.size:
            G.GapGroups._pairs(np.r_[zi, zj, np.arange(l)], np.r_[zj, zi, np.arange(l)], 0.0, cross, ck, lamb, lk, skip_diag_diag=True)
        assert np.array_equal(gg.cross_idx, np.array(cross, dtype=np.int32).reshape(-1, 4))
        assert np.array_equal(gg.cross_key, np.array(ck, dtype=float))
        assert np.array_equal(gg.lamb_idx, np.array(lamb, dtype=np.int32).reshape(-1, 3))
        assert np.array_equal(gg.lamb_key, np.array(lk, dtype=float))
        assert len(cross) > 0
