"""GPU parity of the Davies generator on DEGENERATE spectra and with several generators per DiffEqLiouvillian:
the dense per-group path of csrc/ame_terms.cu (L_x = A∘[ω̃ == x] applied group by group, src/opensys/davies.jl:26-45),
the library-side GapIndices (src/opensys/opensys_util.jl:25-61) and the term list
(src/opensys/diffeq_liouvillian.jl:101-103, src/coupling/interaction.jl:101-117).

Checker: the oracle's LITERAL gap loop where it can finish (≤ 6 qubits), the independent group oracle
(oracle/davies_groups.py, validated against the literal loop in the CPU suite) above that.  Tolerance 1e-12 relative
Frobenius (BASELINE.json north_star)."""
import ctypes as C

import numpy as np
import pytest

from oracle import davies_groups as DG
from oracle import fixtures as F
from oracle import oqb_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-12


def gamma(x):  # mock bath of reference test/opensys/davies.jl:7-8
    return x + 1 if x >= 0 else (1 - x) * np.exp(x)


def lamb(x):
    return x + 0.1


@pytest.fixture(scope="module")
def Q():
    import oqb_b200
    oqb_b200.get_context()
    return oqb_b200


@pytest.fixture(scope="module")
def host_mod(Q):
    import importlib
    return importlib.import_module("oqb_b200.host")


def relerr(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0)


def rand_c(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def z_couplings(q):
    out = []
    for k in range(q):
        out.append(F.q_translate("".join("Z" if j == k else "I" for j in range(q))))
    return out


def oracle_rhs(u, w, v, coup, terms, literal):
    if literal:
        gi = O.build_gap_indices(w, 8, 8)
        out = []
        for x in u:
            rho = v.conj().T @ x @ v
            X = -1j * (w[:, None] * rho - rho * w[None, :])
            for g, S, idx in terms:
                O.DaviesGenerator([coup[k] for k in idx], g, S).rhs_acc(X, rho, gi, v, 0.0)
            out.append(v @ X @ v.conj().T)
        return np.stack(out)
    return np.stack([DG.ame_rhs(x, w, v, coup, terms, 8, 8) for x in u])


@pytest.mark.parametrize("q", [4, 6, 8])
@pytest.mark.parametrize("mode", ["auto", "all_dense", "all_list"])
def test_davies_transverse_field_start(Q, host_mod, q, mode):
    """H(0) = −ΣX: (q+1) levels with binomial degeneracies — the starting point of every anneal.  The zero group holds
    C(2q, q) − 2^q … level pairs; every gap group is large."""
    if mode == "all_list" and q > 4:
        pytest.skip("the pair list of a fully degenerate spectrum is O(m²): only the 4-qubit case is run that way")
    rng = np.random.default_rng(100 + q)
    n = 2 ** q
    Hd = -np.asarray(F.standard_driver(q), dtype=complex)
    coup = z_couplings(q)
    w, v = np.linalg.eigh(Hd)
    ctx = Q.get_context()
    one = lambda s: 1.0
    Hg = Q.DenseHamiltonian([one], [Hd], unit="ħ")
    opg = Q.DiffEqLiouvillian(Hg, [Q.DaviesGenerator(coup, gamma, lamb)], [], n)
    u = np.stack([F.random_density_matrix(n, rng) for _ in range(2)] + [rand_c(rng, n, n)])
    thr = {"auto": 0, "all_dense": 2, "all_list": 10 ** 9}[mode]
    with ctx.option("davies_dense_min_pairs", thr), ctx.option("davies_max_cross", 10 ** 12 if mode == "all_list" else 50_000_000), \
            ctx.option("debug_checks", 1):  # the term builder's own device-side bounds / consistency checks are on
        opg.clear_plans()
        du = opg.rhs_with_eig(np.zeros_like(u), u, Q.ODEParams(None, 1.0), 0.0, w, v)
        stats = host_mod.davies_stats(ctx, opg._ame)
        cache = opg.update_cache(np.zeros((n, n), complex), Q.ODEParams(None, 1.0), 0.0, w, v)
    ref = oracle_rhs(u, w, v, coup, [(gamma, lamb, list(range(q)))], literal=q <= 6)
    assert relerr(du, ref) < TOL
    if mode == "all_dense":
        assert stats[1] == 0 and stats[3] > 0 and stats[4] > 0
    if mode == "all_list":
        assert stats[3] == 0 and stats[1] > 0
    if mode == "auto":
        assert stats[3] > 0  # the zero group and the ±2 groups are large
    if q <= 4:  # update_cache! (src/opensys/davies.jl:76-99) through the same tables
        gi = O.build_gap_indices(w, 8, 8)
        refc = np.diag(-1j * w).astype(complex)
        O.DaviesGenerator(coup, gamma, lamb).update_cache_acc(refc, gi, v, 0.0)
        assert relerr(cache, v @ refc @ v.conj().T) < TOL


@pytest.mark.parametrize("q", [4, 6])
@pytest.mark.parametrize("s", [0.0, 1e-9, 1.0, 0.37])
def test_davies_c3_family_at_symmetric_points(Q, q, s):
    """s = 0 (exactly degenerate), s = 1e-9 (splittings below the 10^-8 zero-gap threshold and gaps that coincide after
    rounding to 8 digits), s = 1 (diagonal Ising: every level at least doubly degenerate), and a generic s."""
    rng = np.random.default_rng(7 * q)
    n = 2 ** q
    Hd = -np.asarray(F.standard_driver(q), dtype=complex)
    Hp = np.asarray(F.random_ising(q, rng), dtype=complex)
    H = (1 - s) * Hd + s * Hp
    w, v = np.linalg.eigh(H)
    coup = z_couplings(q) if q == 4 else [sum(z_couplings(q))]
    bath = Q.Ohmic(1e-4, 4, 16)
    ob = O.Ohmic(1e-4, 4, 16)
    ctx = Q.get_context()
    Hg = Q.DenseHamiltonian([lambda x: 1.0], [H], unit="ħ")
    opg = Q.DiffEqLiouvillian(Hg, [Q.DaviesGenerator(coup, bath, Q.ZERO_LAMBSHIFT)], [], n)
    u = np.stack([F.random_density_matrix(n, rng) for _ in range(2)])
    du = opg.rhs_with_eig(np.zeros_like(u), u, Q.ODEParams(None, 1.0), 0.0, w, v)
    ref = oracle_rhs(u, w, v, coup, [(ob.spectrum, lambda x: 0.0, list(range(len(coup))))], literal=q <= 6)
    assert relerr(du, ref) < TOL
    # host-evaluated closures (the generic path a Julia caller takes) give the same tables
    opc = Q.DiffEqLiouvillian(Hg, [Q.DaviesGenerator(coup, ob.spectrum, lambda x: 0.0)], [], n)
    du2 = opc.rhs_with_eig(np.zeros_like(u), u, Q.ODEParams(None, 1.0), 0.0, w, v)
    assert relerr(du2, ref) < TOL


def test_two_davies_generators_and_onesided(Q):
    """opensys_eig with two DaviesGenerators (one per interaction) — and a OneSidedAMELiouvillian next to them."""
    rng = np.random.default_rng(21)
    n = 24
    Hm = rand_c(rng, n, n); Hm = (Hm + Hm.conj().T) / 2
    w = np.sort(np.r_[np.repeat([-1.5, 0.25], 3), np.arange(6) * 0.5 + 1.0, rng.standard_normal(n - 12) * 2])  # degenerate + equally spaced + generic
    v = np.linalg.qr(rand_c(rng, n, n))[0]
    Hm = v @ np.diag(w) @ v.conj().T
    cs = [(lambda m: m + m.conj().T)(rand_c(rng, n, n)) for _ in range(4)]
    g2 = lambda x: 0.3 * gamma(x)
    s2 = lambda x: 0.5 * x
    Hg = Q.DenseHamiltonian([lambda x: 1.0], [Hm], unit="ħ")
    Ho = O.DenseHamiltonian([lambda x: 1.0], [Hm], unit="ħ")
    terms_g = [Q.DaviesGenerator(cs[:2], gamma, lamb), Q.DaviesGenerator(cs[2:3], g2, s2),
               Q.OneSidedAMELiouvillian(cs[3:], gamma, lamb, [(1, 1)])]
    terms_o = [O.DaviesGenerator(cs[:2], gamma, lamb), O.DaviesGenerator(cs[2:3], g2, s2),
               O.OneSidedAMELiouvillian(cs[3:], gamma, lamb, [(1, 1)])]
    opg = Q.DiffEqLiouvillian(Hg, terms_g, [], n)
    opo = O.DiffEqLiouvillian(Ho, terms_o, [], n)
    u = rand_c(rng, 3, n, n)
    ctx = Q.get_context()
    for thr in (0, 2, 10 ** 9):
        with ctx.option("davies_dense_min_pairs", thr):
            opg.clear_plans()
            du = opg.rhs_with_eig(np.zeros_like(u), u, Q.ODEParams(None, 1.0), 0.5, w, v)
        ref = np.stack([opo.rhs_with_eig(x, O.ODEParams(None, 1.0), 0.5, w, v) for x in u])
        assert relerr(du, ref) < TOL, thr
    # cross terms only (two generators, equally spaced levels, no degenerate level ⇒ no L'L list): the fused-epilogue form
    # (cross-term sources recovered as X/C, also on the diagonal) and the separate-pass form
    w2 = np.sort(np.r_[np.arange(8) * 0.5, rng.standard_normal(n - 8) * 2 + 7.0])
    H2 = v @ np.diag(w2) @ v.conj().T
    Hg2, Ho2 = Q.DenseHamiltonian([lambda x: 1.0], [H2], unit="ħ"), O.DenseHamiltonian([lambda x: 1.0], [H2], unit="ħ")
    og2 = Q.DiffEqLiouvillian(Hg2, terms_g[:2], [], n)
    oo2 = O.DiffEqLiouvillian(Ho2, terms_o[:2], [], n)
    ref2 = np.stack([oo2.rhs_with_eig(x, O.ODEParams(None, 1.0), 0.5, w2, v) for x in u])
    for fuse in (1, 0):
        with ctx.option("fuse_cross", fuse), ctx.option("davies_dense_min_pairs", 10 ** 9):
            og2.clear_plans()
            assert relerr(og2.rhs_with_eig(np.zeros_like(u), u, Q.ODEParams(None, 1.0), 0.5, w2, v), ref2) < TOL, fuse
    # update_cache! accumulates both generators (src/opensys/diffeq_liouvillian.jl:118-120)
    opg2 = Q.DiffEqLiouvillian(Hg, terms_g[:2], [], n)
    opo2 = O.DiffEqLiouvillian(Ho, terms_o[:2], [], n)
    cache = opg2.update_cache(np.zeros((n, n), complex), Q.ODEParams(None, 1.0), 0.5, w, v)
    assert relerr(cache, opo2.update_cache(O.ODEParams(None, 1.0), 0.5, w, v)) < TOL


@pytest.mark.parametrize("kind", ["generic", "degenerate", "unsorted"])
def test_library_gap_indices_match_reference_lists(Q, kind):
    """oqb_ame_gaps_build / _keys / _fetch return GapIndices in the reference's own form: uniq_w ascending, uniq_a/uniq_b in
    the (i<j) loop order, a0/b0 as (i,j),(j,i) per degenerate pair then the diagonal (src/opensys/opensys_util.jl:25-61)."""
    from oqb_b200 import _lib
    rng = np.random.default_rng(3)
    l = 37
    w = {"generic": np.sort(rng.standard_normal(l)),
         "degenerate": np.sort(np.round(rng.standard_normal(l) * 2) * 0.5),
         "unsorted": np.round(rng.standard_normal(l) * 3) * 0.25 + rng.standard_normal(l) * (rng.random(l) < 0.3)}[kind]
    ctx = Q.get_context()
    lib = ctx.lib
    h = C.c_void_p()
    coup = np.ascontiguousarray(np.eye(l, dtype=complex))
    ctx.check(lib.oqb_ame_create(ctx.h, l, l, 1, coup.ctypes.data, C.byref(h)))
    wbuf = np.ascontiguousarray(w, dtype=np.float64)
    ctx.check(lib.oqb_ame_set_eigen(h, _lib.dptr(wbuf), None))
    nk, nz = C.c_int64(), C.c_int64()
    ctx.check(lib.oqb_ame_gaps_build(h, 8, 8, C.byref(nk), C.byref(nz)))
    gi = O.build_gap_indices(w, 8, 8)
    assert nk.value == len(gi.uniq_w)
    keys = np.empty(nk.value)
    ctx.check(lib.oqb_ame_gaps_keys(h, _lib.dptr(keys)))
    assert np.array_equal(keys, np.asarray(gi.uniq_w, dtype=float))
    npairs = l * (l - 1) // 2 - nz.value
    ptr = np.empty(nk.value + 1, dtype=np.int64)
    a = np.empty(npairs, dtype=np.int32); b = np.empty(npairs, dtype=np.int32)
    a0 = np.empty(2 * nz.value + l, dtype=np.int32); b0 = np.empty(2 * nz.value + l, dtype=np.int32)
    ctx.check(lib.oqb_ame_gaps_fetch(h, ptr.ctypes.data_as(_lib.c_i64p), a.ctypes.data_as(_lib.c_i32p), b.ctypes.data_as(_lib.c_i32p),
                                     a0.ctypes.data_as(_lib.c_i32p), b0.ctypes.data_as(_lib.c_i32p)))
    for g in range(nk.value):
        assert list(a[ptr[g]:ptr[g + 1]]) == list(gi.uniq_a[g])
        assert list(b[ptr[g]:ptr[g + 1]]) == list(gi.uniq_b[g])
    assert list(a0) == list(gi.a0) and list(b0) == list(gi.b0)
    lib.oqb_ame_destroy(h)


def test_small_batches_beyond_grid_limit(Q):
    """More than 65 535 tiny density matrices through the AME path (gridDim.y slicing of the elementwise kernels)."""
    rng = np.random.default_rng(9)
    n, B = 2, 70_000
    H = np.array([[0.3, 0.2 - 0.1j], [0.2 + 0.1j, -0.4]])
    w, v = np.linalg.eigh(H)
    coup = [np.array([[1.0, 0.5], [0.5, -1.0]], dtype=complex)]
    Hg = Q.DenseHamiltonian([lambda x: 1.0], [H], unit="ħ")
    opg = Q.DiffEqLiouvillian(Hg, [Q.DaviesGenerator(coup, gamma, lamb)], [], 2)
    u = rand_c(rng, B, n, n)
    du = opg.rhs_with_eig(np.zeros_like(u), u, Q.ODEParams(None, 1.0), 0.5, w, v)
    Ho = O.DenseHamiltonian([lambda x: 1.0], [H], unit="ħ")
    opo = O.DiffEqLiouvillian(Ho, [O.DaviesGenerator(coup, gamma, lamb)], [], 2)
    pick = [0, 1, 65534, 65535, 65536, B - 1]
    ref = np.stack([opo.rhs_with_eig(u[i], O.ODEParams(None, 1.0), 0.5, w, v) for i in pick])
    assert relerr(du[pick], ref) < TOL


# ---- the whole call inside the library: oqb_liouv (csrc/liouvillian.cu) ---------------------------------------------
def _lib_case(Q, q, rng, closures=False, onesided=False):
    n = 2 ** q
    Hd = -np.asarray(F.standard_driver(q), dtype=complex)
    Hp = np.asarray(F.random_ising(q, rng), dtype=complex)
    coup = z_couplings(q)
    fs = [lambda s: 1 - s, lambda s: s]
    bath, ob = Q.Ohmic(1e-4, 4, 16), O.Ohmic(1e-4, 4, 16)
    g_g, S_g = (ob.spectrum, lambda x: 0.05 * x) if closures else (bath, Q.ZERO_LAMBSHIFT)
    g_o, S_o = (ob.spectrum, (lambda x: 0.05 * x) if closures else (lambda x: 0.0))
    Hg, Ho = Q.DenseHamiltonian(fs, [Hd, Hp]), O.DenseHamiltonian(fs, [Hd, Hp])
    cg = Q.ConstantCouplings(coup, unit="h")
    co = [2 * np.pi * c for c in coup]
    if onesided:
        tg = [Q.DaviesGenerator(Q.ConstantCouplings(coup[:2], unit="h"), g_g, S_g),
              Q.OneSidedAMELiouvillian(Q.ConstantCouplings(coup[2:4], unit="h"), g_g, S_g, [(1, 1), (2, 2)])]
        to = [O.DaviesGenerator(co[:2], g_o, S_o), O.OneSidedAMELiouvillian(co[2:4], g_o, S_o, [(1, 1), (2, 2)])]
    else:
        tg = [Q.DaviesGenerator(cg, g_g, S_g)]
        to = [O.DaviesGenerator(co, g_o, S_o)]
    opg = Q.DiffEqLiouvillian(Hg, tg, [], n)
    opg.library = True
    opo = O.DiffEqLiouvillian(Ho, to, [], n)
    return opg, opo, n


@pytest.mark.parametrize("closures", [False, True])
@pytest.mark.parametrize("onesided", [False, True])
def test_library_liouvillian_one_call(Q, closures, onesided):
    """(Op::DiffEqLiouvillian{true,false})(du,u,p,t) entirely behind oqb_liouv_rhs: H(s) assembly, device eigensolver,
    GapIndices, term tables (device-evaluated Ohmic or C callbacks into host closures), rotations.  The eigen-system comes
    from another solver than the oracle's LAPACK, so the bound is the gauge-invariant 1e-10 (as test_ame_device_eigensolver)."""
    rng = np.random.default_rng(41)
    opg, opo, n = _lib_case(Q, 5, rng, closures, onesided)
    pg, po = Q.ODEParams(None, 10.0), O.ODEParams(None, 10.0)
    u = np.stack([F.random_density_matrix(n, rng) for _ in range(5)])
    # per-member s, interleaved: members 0,2,4 share s = 0.3; 1,3 share s = 0.7 (gather/scatter inside the library)
    t = np.array([3.0, 7.0, 3.0, 7.0, 3.0])
    du = opg(np.zeros_like(u), u, pg, t)
    ref = np.stack([opo.rhs(x, po, tt) for x, tt in zip(u, t)])
    assert relerr(du, ref) < 1e-10
    built, hits, ms, eig_ms = opg.liouv_stats()
    assert built == 2 and hits == 0
    du2 = opg(np.zeros_like(u), u, pg, t)          # both eigen-systems come from the cache now
    assert np.array_equal(du2, du)
    assert opg.liouv_stats()[:2] == (2, 2)
    # shared s, host buffers, update_cache!
    du3 = opg(np.zeros_like(u), u, pg, 3.0)
    assert relerr(du3, np.stack([opo.rhs(x, po, 3.0) for x in u])) < 1e-10
    uh = np.ascontiguousarray(u.transpose(0, 2, 1)); duh = np.zeros_like(uh)
    cf = np.ascontiguousarray([[0.7, 0.3]] * 5)
    import oqb_b200._lib as _l
    ctx = Q.get_context()
    ctx.check(ctx.lib.oqb_liouv_rhs_host(opg._liouv(), 5, _l.dptr(cf), 0, uh.ctypes.data, duh.ctypes.data))
    assert relerr(duh.transpose(0, 2, 1), du3) < 1e-14
    if not onesided:
        cache = opg.update_cache(np.zeros((n, n), complex), pg, 3.0)
        assert relerr(cache, opo.update_cache(po, 3.0)) < 1e-10


def test_library_liouvillian_degenerate_start(Q):
    """The anneal's starting point through the one-call path: s = 0 (H = −ΣX, binomially degenerate, dense gap groups) and
    the first stage times after it."""
    rng = np.random.default_rng(43)
    opg, opo, n = _lib_case(Q, 6, rng)
    u = np.stack([F.random_density_matrix(n, rng) for _ in range(2)])
    pg, po = Q.ODEParams(None, 10.0), O.ODEParams(None, 10.0)
    for t in (0.0, 1e-8, 0.05):
        du = opg(np.zeros_like(u), u, pg, t)
        ref = np.stack([opo.rhs(x, po, t) for x in u])
        assert relerr(du, ref) < 1e-10, t


# ---- CorrelatedDaviesGenerator / ConstHCorrelatedDaviesGenerator on DEGENERATE gap structures -------------------------
@pytest.mark.parametrize("thr", [0, 2, 10 ** 9])
@pytest.mark.parametrize("with_v", [False, True])
def test_correlated_davies_degenerate_groups(Q, thr, with_v):
    """src/opensys/davies.jl:231-296 with multi-pair gap groups and a non-trivial zero group (degenerate, equally spaced and
    generic levels mixed), a NON-symmetric pair set (so −½G ∓ iH_LS are two different matrices), complex couplings — against
    the oracle's literal loop, with every gap group forced onto the dense path, onto the list, and by the cost model."""
    from oqb_b200.extra_liouvillians import CorrelatedDaviesGenerator
    rng = np.random.default_rng(55)
    n, K, nb = 14, 3, 3
    w = np.sort(np.r_[np.repeat([-1.5, 0.25], 3), np.arange(4) * 0.5 + 1.0, rng.standard_normal(n - 10) * 2])
    v = np.linalg.qr(rand_c(rng, n, n))[0]
    coup = [rand_c(rng, n, n) for _ in range(K)]
    inds = [(1, 2), (2, 1), (1, 3), (2, 2)]  # (3,1) missing on purpose
    gd = {pq: (lambda x, c=0.3 * (pq[0] + 2 * pq[1]): gamma(x) * c) for pq in inds}
    Sd = {pq: (lambda x, c=0.1 * pq[0]: lamb(x) * c) for pq in inds}
    ctx = Q.get_context()
    Dg = CorrelatedDaviesGenerator(coup, gd, Sd, inds)
    Do = O.CorrelatedDaviesGenerator(coup, gd, Sd, inds)
    G = O.build_gap_indices(w, 8, 8)
    rho = rand_c(rng, nb, n, n)
    du0 = rand_c(rng, nb, n, n)
    with ctx.option("davies_dense_min_pairs", thr), ctx.option("debug_checks", 1):
        got = Dg(du0.copy(), rho, w, v, 0.4) if with_v else Dg(du0.copy(), rho, w, 0.4)
    ref = np.stack([Do.rhs_acc(du0[b].copy(), rho[b], G, v if with_v else None, 0.4) for b in range(nb)])
    assert relerr(got, ref) < TOL


def test_const_h_correlated_davies(Q):
    """ConstHCorrelatedDaviesGenerator (src/opensys/davies.jl:297-345): fixed GapIndices of a constant Hamiltonian with
    UNSORTED, partly degenerate energies, time-dependent couplings given in the eigenbasis."""
    from oqb_b200.extra_liouvillians import ConstHCorrelatedDaviesGenerator
    rng = np.random.default_rng(56)
    l, nb = 9, 2
    w = np.array([0.4, -1.0, 2.0, 0.4, 1.0, -1.0, 0.0, 1.9, 0.4])
    c1, c2 = rand_c(rng, l, l), rand_c(rng, l, l)
    coup = [lambda s: (1 - s) * c1 + s * c2, lambda s: c2 * (1 + s)]
    inds = [(1, 2), (2, 1), (1, 1)]
    gd = {pq: (lambda x, c=0.2 * (pq[0] + pq[1]): gamma(x) * c) for pq in inds}
    Sd = {pq: (lambda x, c=0.05 * pq[1]: lamb(x) * c) for pq in inds}
    Dg = ConstHCorrelatedDaviesGenerator(w, coup, gd, Sd, inds)
    Do = O.CorrelatedDaviesGenerator(coup, gd, Sd, inds)  # the same loop on a fixed GapIndices (oracle docstring)
    G = O.build_gap_indices(w, 8, 8)
    rho = rand_c(rng, nb, l, l)
    for s in (0.0, 0.3):
        got = Dg(np.zeros_like(rho), rho, None, s)
        ref = np.stack([Do.rhs_acc(np.zeros((l, l), complex), rho[b], G, None, s) for b in range(nb)])
        assert relerr(got, ref) < TOL
