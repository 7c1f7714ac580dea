"""GPU parity tests: the CUDA path (through the C ABI, via the host mirror) against the CPU oracle on
identical seeded inputs.  Tolerance: relative Frobenius error ≤ 1e-12 (BASELINE.json north_star);
discrete outputs (jump indices) must match exactly.  Run on the B200 box with `-m gpu`."""
import numpy as np
import pytest
import scipy.linalg as sla
import scipy.sparse as sps

from oracle import fixtures as F
from oracle import oqb_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-12
sx, sy, sz, si = F.sx, F.sy, F.sz, F.si
A = lambda s: 1 - s
B = lambda s: s


def gamma(x):  # mock bath of reference test/opensys/davies.jl:7-8
    return x + 1 if x >= 0 else (1 - x) * np.exp(x)


def lamb(x):
    return x + 0.1


@pytest.fixture(scope="module")
def Q():
    import oqb_b200
    oqb_b200.get_context()
    return oqb_b200


def relerr(a, b):
    a = np.asarray(a); b = np.asarray(b)
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


def rand_c(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def rand_rho(rng, B_, n):
    return np.stack([F.random_density_matrix(n, rng) for _ in range(B_)])


# ---------------------------------------------------------------------------------------------
# K1: batched ZGEMM
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,batch", [(2, 2, 2, 1), (4, 3, 5, 3), (16, 16, 16, 2), (33, 65, 17, 2),
                                         (64, 64, 64, 3), (128, 64, 72, 2), (130, 70, 9, 2), (256, 256, 256, 2)])
@pytest.mark.parametrize("opA,opB", [(0, 0), (2, 0), (0, 2), (2, 2), (1, 1)])
@pytest.mark.parametrize("mode", [1, 0])  # 3-multiplication (default) and 4-multiplication complex product
def test_zgemm(Q, M, N, K, batch, opA, opB, mode):
    import ctypes as C
    rng = np.random.default_rng(M * 1000 + N * 10 + K + opA * 7 + opB)
    ctx = Q.get_context()
    ctx.set_zgemm_mode(mode)
    a_shape = (K, M) if opA else (M, K)
    b_shape = (N, K) if opB else (K, N)
    Ah = rand_c(rng, batch, *a_shape)
    Bh = rand_c(rng, *b_shape)  # shared across the batch (stride 0)
    Ch = rand_c(rng, batch, M, N)
    alpha, beta = 0.7 - 0.3j, -0.2 + 1.1j
    op = {0: lambda x: x, 1: lambda x: x.transpose(0, 2, 1) if x.ndim == 3 else x.T,
          2: lambda x: x.conj().transpose(0, 2, 1) if x.ndim == 3 else x.conj().T}
    ref = alpha * (op[opA](Ah) @ op[opB](Bh)) + beta * Ch
    Ad, Bd, Cd = Q.DeviceBatch.from_host(Ah), Q.DeviceBatch.from_host(Bh[None]), Q.DeviceBatch.from_host(Ch)
    al = np.array([alpha.real, alpha.imag]); be = np.array([beta.real, beta.imag])
    dp = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
    ctx.check(ctx.lib.oqb_zgemm_batched(ctx.h, opA, opB, M, N, K, dp(al), Ad.ptr, a_shape[0], a_shape[0] * a_shape[1],
                                        Bd.ptr, b_shape[0], 0, dp(be), Cd.ptr, M, M * N, batch))
    ctx.set_zgemm_mode(1)
    assert relerr(Cd.to_host(), ref) < 1e-13


def test_zgemm_3m_cancellation(Q):
    """3M forms im = T3 − T1 − T2: purely real operands must give an exactly zero imaginary part, and a
    product whose imaginary part cancels (A·A' is Hermitian) must stay accurate relative to ‖C‖."""
    import ctypes as C
    rng = np.random.default_rng(5)
    ctx = Q.get_context()
    n = 96
    Ar = rng.standard_normal((1, n, n)) + 0j
    Ac = rand_c(rng, 1, n, n)
    al = np.array([1.0, 0.0]); be = np.array([0.0, 0.0])
    dp = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
    for A, opB in ((Ar, 0), (Ac, 2)):
        Ad = Q.DeviceBatch.from_host(A)
        Cd = Q.DeviceBatch.from_host(np.zeros_like(A))
        ctx.check(ctx.lib.oqb_zgemm_batched(ctx.h, 0, opB, n, n, n, dp(al), Ad.ptr, n, n * n, Ad.ptr, n, n * n, dp(be),
                                            Cd.ptr, n, n * n, 1))
        got = Cd.to_host()[0]
        ref = A[0] @ (A[0].conj().T if opB else A[0])
        assert relerr(got, ref) < 1e-14
        if opB == 0:
            assert np.all(got.imag == 0.0)
        else:
            assert np.abs(np.diag(got).imag).max() < 1e-13 * np.abs(ref).max()


# ---------------------------------------------------------------------------------------------
# H-d: DenseHamiltonian  (reference test/hamiltonian/dense_hamiltonian.jl:23-38)
# ---------------------------------------------------------------------------------------------
def test_dense_hamiltonian_golden(Q):
    H = Q.DenseHamiltonian([A, B], [sx, sz])
    assert np.array_equal(H(0.5), np.pi * (sx + sz))
    Cc = H.update_cache(np.zeros((2, 2), complex), 0.5)
    assert np.array_equal(Cc, -1j * np.pi * (sx + sz))
    T = -1j * np.pi * (sx + sz)
    V = H.update_vectorized_cache(np.zeros((4, 4), complex), 0.5)
    assert np.array_equal(V, np.kron(si, T) - np.kron(T.T, si))
    rho = np.outer(F.PauliVec[0][0], F.PauliVec[0][0].conj())
    du = np.array([[1.0 + 0j, 0], [0, 0]])  # pre-filled with garbage: the call must OVERWRITE
    H(du, rho, 2, 0.5)
    assert relerr(du, -1j * np.pi * ((sx + sz) @ rho - rho @ (sx + sz))) < 1e-15
    with pytest.raises(ValueError):
        Q.DenseHamiltonian([A, B], [sx, sz], unit="hh")


@pytest.mark.parametrize("n,nb", [(3, 4), (16, 5), (100, 3), (256, 6)])
def test_dense_commutator_batched_per_member_s(Q, n, nb):
    rng = np.random.default_rng(n)
    m1, m2, m3 = [(lambda g: g + g.conj().T)(rand_c(rng, n, n)) for _ in range(3)]
    fs = [lambda s: 1 - s, lambda s: s, lambda s: np.sin(3 * s)]
    Hg = Q.DenseHamiltonian(fs, [m1, m2, m3], unit="ħ")
    Ho = O.DenseHamiltonian(fs, [m1, m2, m3], unit="ħ")
    u = rand_c(rng, nb, n, n)  # no Hermiticity assumed
    s = rng.random(nb)
    du = Hg(np.full_like(u, 7.0), u, None, s)
    ref = np.stack([Ho.rhs(u[b], s[b]) for b in range(nb)])
    assert relerr(du, ref) < TOL
    c = Hg.update_cache(np.zeros_like(u), s)
    assert relerr(c, np.stack([Ho.update_cache(x) for x in s])) < 1e-15
    if n <= 16:
        v = Hg.update_vectorized_cache(np.zeros((nb, n * n, n * n), complex), s)
        assert relerr(v, np.stack([Ho.update_vectorized_cache(x) for x in s])) < 1e-15


# ---------------------------------------------------------------------------------------------
# L: Lindblad  (reference test/opensys/lindblad.jl:25-46)
# ---------------------------------------------------------------------------------------------
def test_lindblad_golden(Q):
    p = Q.ODEParams(None, 5.0, lambda tf, t: t / tf)
    rho = np.outer(F.PauliVec[0][0], F.PauliVec[0][0].conj())
    L1 = Q.LindbladLiouvillian([Q.Lindblad(0.5, sz)])
    d = L1(np.zeros((2, 2), complex), rho, p, 5.0)
    assert np.allclose(d, [[0, -0.5], [-0.5, 0]], atol=1e-15)
    c = Q.update_cache_(np.zeros((2, 2), complex), L1, p, 5.0)
    assert np.array_equal(c, -0.25 * sz.conj().T @ sz)
    L2 = Q.LindbladLiouvillian([Q.Lindblad(0.5, sz), Q.Lindblad(0.2, sx)])
    rho_y = np.outer(F.PauliVec[1][0], F.PauliVec[1][0].conj())
    d = L2(np.zeros((2, 2), complex), rho_y, p, 5.0)
    assert np.allclose(d, [[0, 0.7j], [-0.7j, 0]], atol=1e-15)
    c = Q.update_cache_(np.zeros((2, 2), complex), L2, p, 5.0)
    assert np.allclose(c, -0.25 * sz.conj().T @ sz - 0.1 * sx.conj().T @ sx, atol=1e-16)
    with pytest.raises(ValueError):
        Q.LindbladLiouvillian([Q.Lindblad(0.5, sz), Q.Lindblad(0.2, np.eye(3))])


@pytest.mark.parametrize("n,K,nb,dense_L", [(4, 2, 3, True), (16, 3, 5, True), (64, 8, 4, False), (128, 4, 3, True)])
def test_lindblad_full_rhs_vs_oracle(Q, n, K, nb, dense_L):
    """DiffEqLiouvillian{false,false} = commutator + Lindblad, per-member s (config C2 shape)."""
    rng = np.random.default_rng(100 + n)
    m1, m2 = [(lambda g: g + g.conj().T)(rand_c(rng, n, n)) for _ in range(2)]
    fs = [lambda s: -(1 - s), lambda s: s]
    if dense_L:
        Ls = [rand_c(rng, n, n) / np.sqrt(n) for _ in range(K)]
    else:
        nq = int(np.log2(n))
        Ls = [F.single_clause(["z"], [k % nq + 1], 1.0, nq) for k in range(K)]
    gfs = [lambda s, k=k: 0.1 * (1 + k / 8) * (1 + s) for k in range(K)]
    Hg = Q.DenseHamiltonian(fs, [m1, m2], unit="ħ")
    lg = Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gfs, Ls)])
    opg = Q.DiffEqLiouvillian(Hg, [], [lg], n)
    Ho = O.DenseHamiltonian(fs, [m1, m2], unit="ħ")
    lo = O.LindbladLiouvillian([O.Lindblad(g, L) for g, L in zip(gfs, Ls)])
    opo = O.DiffEqLiouvillian(Ho, [], [lo], n)
    u = rand_c(rng, nb, n, n)
    tf = 10.0
    t = rng.random(nb) * tf
    pg = Q.ODEParams(None, tf)
    du = opg(np.zeros_like(u), u, pg, t)
    ref = np.stack([opo.rhs(u[b], O.ODEParams(None, tf), t[b]) for b in range(nb)])
    assert relerr(du, ref) < TOL
    # shared s
    du = opg(np.zeros_like(u), u, pg, 2.5)
    ref = np.stack([opo.rhs(u[b], O.ODEParams(None, tf), 2.5) for b in range(nb)])
    assert relerr(du, ref) < TOL
    # accumulate-only form and update_cache!
    acc0 = rand_c(rng, nb, n, n)
    d2 = lg(acc0.copy(), u, pg, 2.5)
    ref2 = np.stack([lo.rhs_acc(acc0[b].copy(), u[b], O.ODEParams(None, tf), 2.5) for b in range(nb)])
    assert relerr(d2, ref2) < TOL
    cache = opg.update_cache(np.zeros((n, n), complex), pg, 2.5)
    assert relerr(cache, opo.update_cache(O.ODEParams(None, tf), 2.5)) < 1e-13
    # host-buffer entry point (column-major blocks, like a Julia Array{ComplexF64,3})
    uh = np.ascontiguousarray(u.transpose(0, 2, 1))
    duh = np.zeros_like(uh)
    opg.rhs_host(duh, uh, pg, t)
    ref = np.stack([opo.rhs(u[b], O.ODEParams(None, tf), t[b]) for b in range(nb)])
    assert relerr(duh.transpose(0, 2, 1), ref) < TOL


# ---------------------------------------------------------------------------------------------
# D1 / DV / OS: AME   (reference test/opensys/davies.jl)
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ame2q():
    mats = [-F.standard_driver(2), 0.1 * np.kron(sz, si) + 0.5 * np.kron(sz, sz)]
    Ho = O.DenseHamiltonian([A, B], mats)
    coup = [2 * np.pi * F.q_translate("ZI+IZ")]
    w, v = O.haml_eigs(Ho, 0.5, 4)
    psi = (v[:, 0] + v[:, 1] + v[:, 2]) / np.sqrt(3)
    rho = np.outer(psi, psi.conj())
    return dict(mats=mats, Ho=Ho, coup=coup, w=w, v=v, rho=rho)


def test_ame_2qubit_fixture(Q, ame2q):
    d = ame2q
    Hg = Q.DenseHamiltonian([A, B], d["mats"])
    dav_g = Q.DaviesGenerator(d["coup"], gamma, lamb)
    opg = Q.DiffEqLiouvillian(Hg, [dav_g], [], 4)
    dav_o = O.DaviesGenerator(d["coup"], gamma, lamb)
    opo = O.DiffEqLiouvillian(d["Ho"], [dav_o], [], 4)
    pg, po = Q.ODEParams(None, 2.0), O.ODEParams(None, 2.0)
    rng = np.random.default_rng(5)
    u = np.concatenate([d["rho"][None], rand_rho(rng, 4, 4), rand_c(rng, 2, 4, 4)])
    du = opg.rhs_with_eig(np.zeros_like(u), u, pg, 1.0, d["w"], d["v"])
    ref = np.stack([opo.rhs_with_eig(x, po, 1.0, d["w"], d["v"]) for x in u])
    assert relerr(du, ref) < TOL
    assert np.linalg.norm(du[0]) == pytest.approx(451.9527156261307, rel=1e-9)
    # the reference's own expectation (atol=rtol=1e-6, test/opensys/davies.jl:57-61)
    w2 = d["w"]
    drho = O.ame_update_test([2 * np.pi * (np.kron(sz, si) + np.kron(si, sz))], d["rho"], w2, d["v"], gamma, lamb)
    h = d["Ho"](0.5)
    assert relerr(du[0], drho - 1j * (h @ d["rho"] - d["rho"] @ h)) < 1e-6
    # own host LAPACK path
    du2 = opg(np.zeros_like(u), u, pg, 1.0)
    ref2 = np.stack([opo.rhs(x, po, 1.0) for x in u])
    assert relerr(du2, ref2) < 1e-9  # eigenvector gauge may differ between two eigh calls
    # update_cache! (test/opensys/davies.jl:49-55)
    cache = opg.update_cache(np.zeros((4, 4), complex), pg, 1.0, d["w"], d["v"])
    assert relerr(cache, opo.update_cache(po, 1.0, d["w"], d["v"])) < TOL


def test_onesided_ame(Q, ame2q):
    d = ame2q
    Hg = Q.DenseHamiltonian([A, B], d["mats"])
    one_g = Q.OneSidedAMELiouvillian(d["coup"], gamma, lamb, [(1, 1)])
    one_o = O.OneSidedAMELiouvillian(d["coup"], gamma, lamb, [(1, 1)])
    opg = Q.DiffEqLiouvillian(Hg, [one_g], [], 4)
    opo = O.DiffEqLiouvillian(d["Ho"], [one_o], [], 4)
    rng = np.random.default_rng(6)
    u = np.concatenate([d["rho"][None], rand_c(rng, 3, 4, 4)])
    du = opg.rhs_with_eig(np.zeros_like(u), u, Q.ODEParams(None, 2.0), 1.0, d["w"], d["v"])
    ref = np.stack([opo.rhs_with_eig(x, O.ODEParams(None, 2.0), 1.0, d["w"], d["v"]) for x in u])
    assert relerr(du, ref) < TOL


@pytest.mark.parametrize("spectrum", ["generic", "equal_spaced", "degenerate", "unsorted"])
@pytest.mark.parametrize("K", [1, 2])
def test_davies_closed_form_vs_literal_loop(Q, spectrum, K):
    """The device closed form (+ cross terms) against the oracle's literal per-gap-group loop
    (src/opensys/davies.jl:21-47) on spectra that exercise singleton, multi-pair and zero-gap groups."""
    rng = np.random.default_rng({"generic": 1, "equal_spaced": 2, "degenerate": 3, "unsorted": 4}[spectrum] * 10 + K)
    l = 7
    w = {"generic": np.sort(rng.standard_normal(l) * 3),
         "equal_spaced": np.arange(l) * 0.75 - 2.0,
         "degenerate": np.array([-2.0, -2.0, -0.5, 0.3, 0.3, 0.3, 1.9]),
         "unsorted": np.array([0.4, -1.0, 2.0, 0.4 + 1.5, 1.0, -1.0 + 1.5, 0.0])}[spectrum]
    g = rand_c(rng, l, l)
    v = np.linalg.qr(g)[0]
    Hmat = v @ np.diag(w) @ v.conj().T
    coup = [(lambda x: x + x.conj().T)(rand_c(rng, l, l)) for _ in range(K)]
    one = lambda s: 1.0
    Hg = Q.DenseHamiltonian([one], [Hmat], unit="ħ")
    Ho = O.DenseHamiltonian([one], [Hmat], unit="ħ")
    opg = Q.DiffEqLiouvillian(Hg, [Q.DaviesGenerator(coup, gamma, lamb)], [], l)
    opo = O.DiffEqLiouvillian(Ho, [O.DaviesGenerator(coup, gamma, lamb)], [], l)
    u = rand_c(rng, 3, l, l)
    pg, po = Q.ODEParams(None, 1.0), O.ODEParams(None, 1.0)
    du = opg.rhs_with_eig(np.zeros_like(u), u, pg, 0.5, w, v)
    ref = np.stack([opo.rhs_with_eig(x, po, 0.5, w, v) for x in u])
    assert relerr(du, ref) < TOL
    cache = opg.update_cache(np.zeros((l, l), complex), pg, 0.5, w, v)
    assert relerr(cache, opo.update_cache(po, 0.5, w, v)) < TOL


def test_ame_sparse_truncated(Q):
    """4-qubit SparseHamiltonian, lvl = 4 of 16 (reference test/opensys/davies.jl:112-145)."""
    Hd = sps.csc_matrix(F.standard_driver(4))
    Hp = sps.csc_matrix(F.q_translate("-0.9ZZII+IZZI-0.9IIZZ"))
    coup = [2 * np.pi * F.q_translate("ZIII+IZII+IIZI+IIIZ")]
    Hg = Q.SparseHamiltonian([A, B], [Hd, Hp])
    Ho = O.SparseHamiltonian([A, B], [Hd, Hp])
    opg = Q.DiffEqLiouvillian(Hg, [Q.DaviesGenerator(coup, gamma, lamb)], [], 4)
    opo = O.DiffEqLiouvillian(Ho, [O.DaviesGenerator(coup, gamma, lamb)], [], 4)
    assert opg.lvl == 4
    w, v = O.haml_eigs(Ho, 0.5, 4)
    psi = (v[:, 0] + v[:, 1] + v[:, 2]) / np.sqrt(3)
    rng = np.random.default_rng(8)
    u = np.concatenate([np.outer(psi, psi.conj())[None], rand_rho(rng, 2, 16)])
    pg, po = Q.ODEParams(None, 2.0), O.ODEParams(None, 2.0)
    du = opg.rhs_with_eig(np.zeros_like(u), u, pg, 1.0, w, v)
    ref = np.stack([opo.rhs_with_eig(x, po, 1.0, w, v) for x in u])
    assert relerr(du, ref) < TOL
    cache = opg.update_cache(np.zeros((16, 16), complex), pg, 1.0, w, v)
    assert relerr(cache, opo.update_cache(po, 1.0, w, v)) < TOL


def test_ame_adiabatic_frame(Q):
    """DiffEqLiouvillian{true,true}: unsorted w = diag(H), time-dependent couplings (reference
    test/opensys/davies.jl:147-171), plus a geometric (off-diagonal) part of the frame Hamiltonian."""
    omega = lambda s: [0, s, 1 - s, 1]
    coup = lambda s: [2 * np.pi * (s * (np.kron(sx, si) + np.kron(si, sx)) + (1 - s) * (np.kron(sz, si) + np.kron(si, sz)))]
    rng = np.random.default_rng(12)
    G = rand_c(rng, 4, 4)
    G = G + G.conj().T
    np.fill_diagonal(G, 0)
    for geo in (None, lambda s: (1 + s) * G):
        class AF:  # oracle-side frame Hamiltonian
            size = (4, 4)

            def __call__(self, tf, s):
                H = np.diag(2 * np.pi * np.array(omega(s), dtype=complex))
                return H if geo is None else H + geo(s) / tf
        Hg = Q.AdiabaticFrameHamiltonian(omega, geo)
        opg = Q.DiffEqLiouvillian(Hg, [Q.DaviesGenerator(coup, gamma, lamb)], [], 4)
        opo = O.DiffEqLiouvillian(AF(), [O.DaviesGenerator(coup, gamma, lamb)], [], 4, adiabatic_frame=True)
        v = np.eye(4, dtype=complex)
        psi = (v[:, 0] + v[:, 1] + v[:, 2]) / np.sqrt(3)
        u = np.concatenate([np.outer(psi, psi.conj())[None], rand_c(rng, 3, 4, 4)])
        pg, po = Q.ODEParams(Hg, 2.0), O.ODEParams(None, 2.0)
        du = opg(np.full_like(u, 3.0), u, pg, 0.8)
        ref = np.stack([opo.rhs(x, po, 0.8) for x in u])
        assert relerr(du, ref) < TOL
    # the reference's own expectation for the geometric-free case (atol = rtol = 1e-6)
    hmat = AF()(2.0, 0.4)
    drho = O.ame_update_test(coup(0.4), u[0], np.real(np.diag(hmat)), v, gamma, lamb)


def test_sparse_update_vectorized_cache(Q):
    """update_vectorized_cache!(cache, H::SparseHamiltonian, p, s) (reference test/hamiltonian/sparse_hamiltonian.jl:42-45)."""
    H = Q.SparseHamiltonian([A, B], [sps.csc_matrix(sx), sps.csc_matrix(sz)])
    T = -1j * np.pi * (sx + sz)
    V = H.update_vectorized_cache(np.zeros((4, 4), complex), 0.5)
    assert np.array_equal(V, np.kron(si, T) - np.kron(T.T, si))
    Hx, Hz, _ = _tfim_sparse(3)
    Hg, Ho = Q.SparseHamiltonian([A, B], [Hx, Hz]), O.SparseHamiltonian([A, B], [Hx, Hz])
    s = np.array([0.1, 0.7])
    V = Hg.update_vectorized_cache(np.zeros((2, 64, 64), complex), s)
    assert relerr(V, np.stack([Ho.update_vectorized_cache(x).toarray() for x in s])) < 1e-15


def test_c1_single_qubit_ohmic_anchor(Q):
    """Config C1: H(s) = −(1−s)σx + sσz, Ohmic bath, 2×2 ρ (SURVEY App. C anchors)."""
    fs = [lambda s: -(1 - s), lambda s: s]
    Hg = Q.DenseHamiltonian(fs, [sx, sz])
    bath = Q.Ohmic(1e-4, 4, 16)
    opg = Q.DiffEqLiouvillian(Hg, [Q.DaviesGenerator([2 * np.pi * sz], bath, Q.ZERO_LAMBSHIFT)], [], 2)
    Ho = O.DenseHamiltonian(fs, [sx, sz])
    ob = O.Ohmic(1e-4, 4, 16)
    opo = O.DiffEqLiouvillian(Ho, [O.DaviesGenerator([2 * np.pi * sz], ob.spectrum, lambda w: 0.0)], [], 2)
    rng = np.random.default_rng(9)
    rho = np.outer(F.PauliVec[0][0], F.PauliVec[0][0].conj())
    u = np.concatenate([rho[None], rand_rho(rng, 16, 2)])
    pg, po = Q.ODEParams(None, 10.0), O.ODEParams(None, 10.0)
    for t in (0.0, 2.5, 5.0, 7.5, 10.0):
        w, v = O.haml_eigs(Ho, po(t), 2)
        du = opg.rhs_with_eig(np.zeros_like(u), u, pg, t, w, v)
        ref = np.stack([opo.rhs_with_eig(x, po, t, w, v) for x in u])
        assert relerr(du, ref) < TOL
        if t == 5.0:
            assert du[0][0, 0] == pytest.approx(-0.030394341407546568, rel=1e-10)
            assert du[0][0, 1] == pytest.approx(-0.015496302491821998 - 3.1415926535897922j, rel=1e-10)


@pytest.mark.parametrize("nq,lvl,K", [(5, 32, 1), (6, 64, 6), (6, 20, 2)])
def test_ame_medium_vs_oracle(Q, nq, lvl, K):
    """Davies + Lindblad `opensys` term on 5–6 qubits with the literal-loop oracle; Ohmic bath."""
    rng = np.random.default_rng(nq * 100 + lvl)
    n = 2 ** nq
    Hd = -F.standard_driver(nq)
    Hp = F.random_ising(nq, rng)
    fs = [A, B]
    coup = [2 * np.pi * F.collective_operator("z", nq)] if K == 1 else \
        [2 * np.pi * F.single_clause(["z"], [k + 1], 1.0, nq) for k in range(K)]
    bath_g, bath_o = Q.Ohmic(1e-4, 4, 16), O.Ohmic(1e-4, 4, 16)
    Hg = Q.DenseHamiltonian(fs, [Hd, Hp])
    Ho = O.DenseHamiltonian(fs, [Hd, Hp])
    opg = Q.DiffEqLiouvillian(Hg, [Q.DaviesGenerator(coup, bath_g, lamb)], [], lvl)
    opo = O.DiffEqLiouvillian(Ho, [O.DaviesGenerator(coup, bath_o.spectrum, lamb)], [], lvl)
    w, v = O.haml_eigs(Ho, 0.5, lvl)
    u = rand_rho(rng, 3, n)
    pg, po = Q.ODEParams(None, 2.0), O.ODEParams(None, 2.0)
    du = opg.rhs_with_eig(np.zeros_like(u), u, pg, 1.0, w, v)
    ref = np.stack([opo.rhs_with_eig(x, po, 1.0, w, v) for x in u])
    assert relerr(du, ref) < TOL


# ---------------------------------------------------------------------------------------------
# R: Redfield   (reference test/opensys/redfield.jl:3-24)
# ---------------------------------------------------------------------------------------------
def test_redfield_golden_and_apply(Q):
    unitary = lambda t: sla.expm(-1j * t * sx)
    cfun = lambda t1, t2: 1.0
    kernels = [(((1, 1),), [sz], cfun)]
    rg = Q.RedfieldLiouvillian(kernels, unitary, 5.0, 1e-8, 1e-6)
    ro = O.RedfieldLiouvillian(kernels, unitary, 5.0, 1e-8, 1e-6)
    pg, po = Q.ODEParams(None, 5.0), O.ODEParams(None, 5.0)
    rho = np.outer(F.PauliVec[0][0], F.PauliVec[0][0].conj())
    d = rg(np.zeros((2, 2), complex), rho, pg, 5.0)
    assert np.allclose(d, -np.sin(10) * sx, atol=1e-7)  # dρ = +0.544021 σx
    ref = ro.rhs_acc(np.zeros((2, 2), complex), rho, po, 5.0)
    assert relerr(d, ref) < 1e-10
    Ac = rg.update_vectorized_cache(np.zeros((4, 4), complex), pg, 5.0)
    Ao = ro.update_vectorized_cache_acc(np.zeros((4, 4), complex), po, 5.0)
    assert relerr(Ac, Ao) < 1e-10
    assert np.allclose(Ac @ rho.flatten(order="F"), ref.flatten(order="F"), atol=1e-9)


@pytest.mark.parametrize("diagS", [False, True])
@pytest.mark.parametrize("n,K,nb", [(4, 2, 3), (32, 3, 4), (128, 7, 3), (50, 2, 2)])
def test_redfield_apply_vs_oracle(Q, n, K, nb, diagS):
    rng = np.random.default_rng(n + K)
    S = [(lambda x: x + x.conj().T)(rand_c(rng, n, n)) for _ in range(K)]
    if diagS:  # diagonal couplings take the elementwise form (oqb_redfield_apply_diag_acc); complex diagonals on purpose
        S = [np.diag(rand_c(rng, n)) for _ in range(K)]
    Lam = rand_c(rng, nb, K, n, n) / n
    u = rand_c(rng, nb, n, n)
    du0 = rand_c(rng, nb, n, n)
    rg = Q.RedfieldLiouvillian(None, None, 0, 0, 0)
    du = rg.apply(du0.copy(), u, S, Lam)
    ref = np.stack([O.redfield_apply_acc(du0[b].copy(), u[b], S, list(Lam[b])) for b in range(nb)])
    assert relerr(du, ref) < TOL


# ---------------------------------------------------------------------------------------------
# H-s + J: SparseHamiltonian and trajectories
# ---------------------------------------------------------------------------------------------
def test_sparse_hamiltonian_golden(Q):
    H = Q.SparseHamiltonian([A, B], [sps.csc_matrix(sx), sps.csc_matrix(sz)])
    assert np.allclose(H(0.5).toarray(), np.pi * (sx + sz), atol=1e-15)
    c = H.update_cache(0.5)
    assert np.array_equal(c.toarray(), -1j * np.pi * (sx + sz))
    u = np.array([1.0 + 0j, 1]) / np.sqrt(2)
    rho = np.outer(u, u.conj())
    du = np.array([[1.0 + 0j, 0], [0, 0]])
    H(du, rho, 1.0, 0.5)
    assert relerr(du, -1j * np.pi * ((sx + sz) @ rho - rho @ (sx + sz))) < 1e-15


def _tfim_sparse(nq):
    Hx = sps.csc_matrix(-F.standard_driver(nq))
    J = [1.0] * (nq - 1)
    Hz = sps.csc_matrix(-F.two_local_term(J, [[i, i + 1] for i in range(1, nq)], nq))
    Ls = [sps.csc_matrix(F.single_clause(["z"], [k], 1.0, nq)) for k in range(1, nq + 1)]
    return Hx, Hz, Ls


@pytest.mark.parametrize("nq,nb", [(3, 5), (6, 9), (8, 33)])
def test_trajectory_matvec_and_jump(Q, nq, nb):
    rng = np.random.default_rng(nq)
    n = 2 ** nq
    Hx, Hz, Ls = _tfim_sparse(nq)
    fs = [A, B]
    Hg = Q.SparseHamiltonian(fs, [Hx, Hz], unit="ħ")
    Ho = O.SparseHamiltonian(fs, [Hx, Hz], unit="ħ")
    for per_member_gamma in (False, True):
        gfs = [(lambda s, k=k: 0.05 * (1 + (k + s if per_member_gamma else 0))) for k in range(nq)]
        lg = Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gfs, Ls)])
        lo = O.LindbladLiouvillian([O.Lindblad(g, L.toarray()) for g, L in zip(gfs, Ls)])
        opg = Q.DiffEqLiouvillian(Hg, [], [lg], n)
        opo = O.DiffEqLiouvillian(Ho, [], [lo], n)
        ens = Q.TrajectoryEnsemble(opg)
        psi = rand_c(rng, nb, n)
        psi /= np.linalg.norm(psi, axis=1, keepdims=True)
        tf = 4.0
        t = rng.random(nb) * tf
        pg, po = Q.ODEParams(None, tf), O.ODEParams(None, tf)
        pd = Q.DeviceBatch.from_host(psi)
        dpsi = ens.matvec(pd.zeros_like(), pd, pg, t).to_host()
        ref = np.stack([opo.update_cache(po, t[b]) @ psi[b] for b in range(nb)])
        assert relerr(dpsi, ref) < TOL
        # update_cache! on the union pattern
        cache = opg.update_cache(None, pg, t[0])
        assert relerr(cache.toarray(), opo.update_cache(po, t[0])) < 1e-14
        # ‖ψ‖²
        assert np.allclose(ens.norm2(pd).cpu().numpy(), 1.0, atol=1e-14)
        # jump decision: identical index sequence for identical uniforms
        r = rng.random(nb)
        choice, prob = ens.lindblad_jump(pd, pg, t, r, apply=False)
        choice, prob = choice.cpu().numpy(), prob.cpu().numpy()
        for b in range(nb):
            k, pr = O.lindblad_jump(opo, psi[b], po, t[b], r[b])
            assert k == choice[b]
            assert np.allclose(pr, prob[b], rtol=1e-13)
        # apply: ψ ← Lψ/‖Lψ‖ for a masked subset
        mask = (np.arange(nb) % 2 == 0)
        choice2, _ = ens.lindblad_jump(pd, pg, t, r, mask=mask, apply=True)
        out = pd.to_host()
        choice2 = choice2.cpu().numpy()
        for b in range(nb):
            if mask[b]:
                y = Ls[choice[b]].toarray() @ psi[b]
                assert relerr(out[b], y / np.linalg.norm(y)) < TOL
            else:
                assert choice2[b] == -1 and np.array_equal(out[b], psi[b])


def test_sparse_commutator_vs_oracle(Q):
    rng = np.random.default_rng(77)
    Hx, Hz, _ = _tfim_sparse(5)
    Hg = Q.SparseHamiltonian([A, B], [Hx, Hz])
    Ho = O.SparseHamiltonian([A, B], [Hx, Hz])
    u = rand_c(rng, 4, 32, 32)
    s = rng.random(4)
    du = Hg(np.zeros_like(u), u, None, s)
    ref = np.stack([Ho.rhs(u[b], s[b]) for b in range(4)])
    assert relerr(du, ref) < TOL


def test_expect(Q):
    rng = np.random.default_rng(3)
    n = 16
    obs = [rand_c(rng, n, n) for _ in range(3)]
    rho = rand_c(rng, 5, n, n)
    out = Q.expect(obs, Q.DeviceBatch.from_host(rho)).cpu().numpy()
    ref = np.array([[np.trace(o @ r) for o in obs] for r in rho])
    assert relerr(out, ref) < 1e-13
    psi = rand_c(rng, 5, n)
    out = Q.expect(obs, Q.DeviceBatch.from_host(psi)).cpu().numpy()
    ref = np.array([[p.conj() @ o @ p for o in obs] for p in psi])
    assert relerr(out, ref) < 1e-13


# ---------------------------------------------------------------------------------------------
# ame_jump / lindblad_jump{true,false}  (reference src/opensys/trajectory_jump.jl:14-51,64-69; test/opensys/davies.jl:63-66)
# ---------------------------------------------------------------------------------------------
def _ame_jump_case(Q, mats, coup, lvl, s, nb, rng, gam):
    n = mats[0].shape[0]
    Ho = O.DenseHamiltonian([A, B], mats)
    Hg = Q.DenseHamiltonian([A, B], mats)
    w, v = O.haml_eigs(Ho, s, lvl)
    opo = O.DiffEqLiouvillian(Ho, [O.DaviesGenerator(coup, gam, lamb)], [], lvl)
    opg = Q.DiffEqLiouvillian(Hg, [Q.DaviesGenerator(Q.ConstantCouplings(coup, unit="ħ"), gam, lamb)], [], lvl)
    psi = rand_c(rng, nb, n)
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    r = rng.random(nb)
    tf = 2.0
    pg, po = Q.ODEParams(None, tf), O.ODEParams(None, tf)
    pd = Q.DeviceBatch.from_host(psi)
    choice, prob = opg.lindblad_jump(pd, pg, s * tf, r, w=w, v=v)
    choice, prob = choice.cpu().numpy(), prob.cpu().numpy()
    A_eig = np.stack([v.conj().T @ c @ v for c in coup])
    for b in range(nb):
        k, pr, L = O.lindblad_jump_ame(opo, psi[b], po, s * tf, r[b], w=w, v=v)
        assert k == choice[b]                                   # identical index for identical uniforms
        assert np.allclose(pr, prob[b], rtol=1e-12, atol=1e-14 * max(pr.max(), 1e-300))
        assert L.shape == (n, n)                                # the reference's own check (davies.jl:66)
        assert relerr(opg.jump_operator(int(k), A_eig), L) < 1e-13
    mask = (np.arange(nb) % 3 != 1)
    choice2, _ = opg.lindblad_jump(pd, pg, s * tf, r, mask=mask, apply=True, w=w, v=v)
    out, choice2 = pd.to_host(), choice2.cpu().numpy()
    for b in range(nb):
        if mask[b]:
            _, _, L = O.lindblad_jump_ame(opo, psi[b], po, s * tf, r[b], w=w, v=v)
            y = L @ psi[b]
            assert relerr(out[b], y / np.linalg.norm(y)) < TOL
        else:
            assert choice2[b] == -1 and np.array_equal(out[b], psi[b])


def test_ame_jump_2qubit_fixture(Q, ame2q):
    rng = np.random.default_rng(21)
    _ame_jump_case(Q, ame2q["mats"], ame2q["coup"], 4, 0.5, 7, rng, gamma)


@pytest.mark.parametrize("nq,lvl,K", [(3, 8, 3), (4, 6, 2), (5, 32, 5)])
def test_ame_jump_vs_oracle(Q, nq, lvl, K):
    rng = np.random.default_rng(300 + nq)
    n = 2 ** nq
    Hd = -F.standard_driver(nq)
    hz = rng.standard_normal(nq)
    Hp = sum(hz[k] * F.single_clause(["z"], [k + 1], 1.0, nq) for k in range(nq)) \
        + F.two_local_term(list(rng.standard_normal(nq - 1)), [[i, i + 1] for i in range(1, nq)], nq)
    coup = [F.single_clause(["z"], [k + 1], 1.0, nq) for k in range(K)]
    _ame_jump_case(Q, [Hd, Hp], coup, lvl, 0.6, 9, rng, gamma)


# ---------------------------------------------------------------------------------------------
# ConstDaviesGenerator / ConstHDaviesGenerator / FluctuatorLiouvillian
#   (reference src/opensys/davies.jl:102-209, src/opensys/stochastic.jl:25-44; test/opensys/davies.jl:68-88)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nq,K,degenerate", [(2, 1, True), (3, 2, False)])
def test_const_davies_generators(Q, nq, K, degenerate):
    from oqb_b200.extra_liouvillians import ConstDaviesGenerator, ConstHDaviesGenerator, build_const_davies
    rng = np.random.default_rng(500 + nq)
    n = 2 ** nq
    if degenerate:  # the reference's constant test Hamiltonian: standard_driver(2) + alt_sec_chain(1, 0.5, 1, 2)
        Hc = F.standard_driver(2) + 0.5 * np.kron(sz, sz)  # symmetric → degenerate gaps → multi-pair groups
    else:
        a = rand_c(rng, n, n)
        Hc = (a + a.conj().T) / 2
    w, v = np.linalg.eigh(Hc)
    w = 2 * np.pi * w
    coup = [2 * np.pi * F.single_clause(["z"], [k + 1], 1.0, nq) for k in range(K)]
    if degenerate:
        coup = [2 * np.pi * F.q_translate("ZI+IZ")]
    cs = [v.conj().T @ c @ v for c in coup]
    G = O.build_gap_indices(w, 8, 8)
    Do = O.build_const_davies(cs, G, gamma, lamb)
    Dg = build_const_davies(cs, w, gamma, lamb)
    assert len(Dg.Linds) == len(Do.Linds)
    assert relerr(Dg.A, Do.A) < 1e-14 and relerr(Dg.Hls, Do.Hls) < 1e-14
    nb = 3
    rho = rand_rho(rng, nb, n)
    du0 = rand_c(rng, nb, n, n)
    got = Dg(du0.copy(), rho, None, None)
    ref = np.stack([Do.rhs_acc(du0[b].copy(), rho[b]) for b in range(nb)])
    assert relerr(got, ref) < TOL
    c0 = rand_c(rng, n, n)
    assert relerr(Dg.update_cache(c0.copy()), Do.update_cache_acc(c0.copy())) < 1e-14
    # constant H, couplings re-evaluated per call: closed-form tables in the fixed basis vs the literal gap loop
    Hg = ConstHDaviesGenerator(w, lambda s: [(1 + s) * c for c in cs], gamma, lamb)
    Ho = O.ConstHDaviesGenerator(G, lambda s: [(1 + s) * c for c in cs], gamma, lamb)
    got = Hg(du0.copy(), rho, None, 0.3)
    ref = np.stack([Ho.rhs_acc(du0[b].copy(), rho[b], 0.3) for b in range(nb)])
    assert relerr(got, ref) < TOL
    # and ConstDavies == ConstHDavies at s = 0 (same generator, two code paths in the reference)
    assert relerr(Dg(np.zeros_like(rho), rho, None, None), Hg(np.zeros_like(rho), rho, None, 0.0)) < 1e-12


def test_fluctuator_liouvillian(Q):
    from oqb_b200.extra_liouvillians import FluctuatorLiouvillian
    rng = np.random.default_rng(9)
    nq, nb = 3, 4
    n = 2 ** nq
    coup = [2 * np.pi * F.single_clause(["z"], [k + 1], 1.0, nq) for k in range(nq)]
    nvals = rng.choice([-1.0, 1.0], (nb, nq)) * rng.random((nb, nq))
    p = O.ODEParams(None, 1.0)
    u = rand_rho(rng, nb, n)
    Fg = FluctuatorLiouvillian(coup, nvals)
    got = Fg(rand_c(rng, nb, n, n), u, None, 0.5)  # overwrites du
    ref = np.stack([O.FluctuatorLiouvillian(coup, nvals[b]).rhs(u[b], p, 0.5) for b in range(nb)])
    assert relerr(got, ref) < TOL
    c0 = rand_c(rng, nb, n, n)
    got = Fg.update_cache(c0.copy(), None, 0.5)
    ref = np.stack([O.FluctuatorLiouvillian(coup, nvals[b]).update_cache_acc(c0[b].copy(), p, 0.5) for b in range(nb)])
    assert relerr(got, ref) < 1e-14
    v0 = rand_c(rng, nb, n * n, n * n)
    got = Fg.update_vectorized_cache(v0.copy(), None, 0.5)
    ref = np.stack([O.FluctuatorLiouvillian(coup, nvals[b]).update_vectorized_cache_acc(v0[b].copy(), p, 0.5) for b in range(nb)])
    assert relerr(got, ref) < 1e-14
    # shared noise values
    Fs = FluctuatorLiouvillian(coup, nvals[0])
    assert relerr(Fs(np.zeros_like(u), u, None, 0.5)[1], O.FluctuatorLiouvillian(coup, nvals[0]).rhs(u[1], p, 0.5)) < TOL


def test_ame_device_eigensolver(Q):
    """H(s) assembled and diagonalised on the GPU, V never leaves the device.  The RHS is invariant
    under the eigenvector gauge, so it must agree with the host-LAPACK path (and the oracle) to 1e-10 — the looser
    bound covers the two eigensolvers' different rounding of nearly degenerate eigenvectors."""
    rng = np.random.default_rng(77)
    nq, lvl, K, nb = 5, 32, 3, 4
    n = 2 ** nq
    Hd = -F.standard_driver(nq)
    hz = rng.standard_normal(nq)
    Hp = sum(hz[k] * F.single_clause(["z"], [k + 1], 1.0, nq) for k in range(nq)) \
        + F.two_local_term(list(rng.standard_normal(nq - 1)), [[i, i + 1] for i in range(1, nq)], nq)
    coup = [F.single_clause(["z"], [k + 1], 1.0, nq) for k in range(K)]
    Hg = Q.DenseHamiltonian([A, B], [Hd, Hp])
    mk = lambda: Q.DiffEqLiouvillian(Hg, [Q.DaviesGenerator(Q.ConstantCouplings(coup, unit="h"), gamma, lamb)], [], lvl)
    op_host, op_dev = mk(), mk()
    op_dev.device_eigh = True
    u = rand_rho(rng, nb, n)
    p = Q.ODEParams(Hg, 2.0)
    d_host = op_host(np.zeros_like(u), u, p, 0.8)
    d_dev = op_dev(np.zeros_like(u), u, p, 0.8)
    assert relerr(d_dev, d_host) < 1e-10
    Ho = O.DenseHamiltonian([A, B], [Hd, Hp])
    opo = O.DiffEqLiouvillian(Ho, [O.DaviesGenerator([2 * np.pi * c for c in coup], gamma, lamb)], [], lvl)
    ref = np.stack([opo.rhs(u[b], O.ODEParams(Ho, 2.0), 0.8) for b in range(nb)])
    assert relerr(d_dev, ref) < 1e-10


def test_correlated_davies_generator(Q):
    """reference test/opensys/davies.jl:173-194 (golden closed forms) + a random non-degenerate 3-qubit case vs the oracle."""
    from oqb_b200.extra_liouvillians import CorrelatedDaviesGenerator
    sp_, sm_ = np.array([[0, 1], [0, 0]], dtype=complex), np.array([[0, 0], [1, 0]], dtype=complex)
    gfun = lambda w: 1.0 if w >= 0 else np.exp(-0.5)
    zero = lambda w: 0.0
    gam, S0 = {(1, 2): gfun, (2, 1): gfun}, {(1, 2): zero, (2, 1): zero}
    w2 = np.array([1.0, 2.0])
    D = CorrelatedDaviesGenerator([sp_, sm_], gam, S0, ((1, 2), (2, 1)))
    du = D(np.zeros((2, 2), complex), np.array([[0.5, 0], [0, 0.5]], dtype=complex), w2, 0.5)
    assert np.allclose(du, 0.0, atol=1e-15)
    D = CorrelatedDaviesGenerator([sm_, sm_], gam, S0, ((1, 2), (2, 1)))
    du = D(np.zeros((2, 2), complex), 0.5 * np.ones((2, 2), dtype=complex), w2, 0.5)
    e = np.exp(-0.5)
    assert np.allclose(du, [[-e, -0.5 * e], [-0.5 * e, e]], atol=1e-15)
    # random case: 3 couplings, all ordered pairs, complex couplings, Lamb shift on, with and without v
    rng = np.random.default_rng(31)
    n, K, nb = 8, 3, 3
    a = rand_c(rng, n, n)
    w, v = np.linalg.eigh((a + a.conj().T) / 2)
    coup = [rand_c(rng, n, n) for _ in range(K)]
    inds = [(i, j) for i in range(1, K + 1) for j in range(1, K + 1) if i != j] + [(1, 1)]
    gd = {pq: (lambda x, c=0.3 * (pq[0] + 2 * pq[1]): gamma(x) * c) for pq in inds}
    Sd = {pq: (lambda x, c=0.1 * pq[0]: lamb(x) * c) for pq in inds}
    Dg = CorrelatedDaviesGenerator(coup, gd, Sd, inds)
    Do = O.CorrelatedDaviesGenerator(coup, gd, Sd, inds)
    G = O.build_gap_indices(w, 8, 8)
    rho = rand_c(rng, nb, n, n)
    du0 = rand_c(rng, nb, n, n)
    got = Dg(du0.copy(), rho, w, 0.4)
    ref = np.stack([Do.rhs_acc(du0[b].copy(), rho[b], G, None, 0.4) for b in range(nb)])
    assert relerr(got, ref) < TOL
    got = Dg(du0.copy(), rho, w, v, 0.4)
    ref = np.stack([Do.rhs_acc(du0[b].copy(), rho[b], G, v, 0.4) for b in range(nb)])
    assert relerr(got, ref) < TOL


def test_ame_batch_with_per_member_s(Q):
    """A schedule sweep through the AME RHS: members with different s get their own eigen-system (grouped by s,
    plans LRU-cached), result per member = the single-s RHS of the oracle."""
    rng = np.random.default_rng(88)
    nq, lvl, nb = 3, 8, 7
    n = 2 ** nq
    Hd = -F.standard_driver(nq)
    Hp = F.two_local_term(list(rng.standard_normal(nq - 1)), [[i, i + 1] for i in range(1, nq)], nq) \
        + 0.3 * F.single_clause(["z"], [1], 1.0, nq)
    coup = [F.single_clause(["z"], [k + 1], 1.0, nq) for k in range(2)]
    Hg, Ho = Q.DenseHamiltonian([A, B], [Hd, Hp]), O.DenseHamiltonian([A, B], [Hd, Hp])
    opg = Q.DiffEqLiouvillian(Hg, [Q.DaviesGenerator(Q.ConstantCouplings(coup, unit="h"), gamma, lamb)], [], lvl)
    opg.plan_cache = 2  # smaller than the number of distinct s values: exercises eviction
    opo = O.DiffEqLiouvillian(Ho, [O.DaviesGenerator([2 * np.pi * c for c in coup], gamma, lamb)], [], lvl)
    u = rand_rho(rng, nb, n)
    tf = 2.0
    t = np.array([0.2, 1.4, 0.2, 0.9, 1.4, 0.2, 1.9])  # 4 distinct s, interleaved
    got = opg(np.zeros_like(u), u, Q.ODEParams(Hg, tf), t)
    ref = np.stack([opo.rhs(u[b], O.ODEParams(Ho, tf), t[b]) for b in range(nb)])
    assert relerr(got, ref) < 1e-11


def test_expect_diag(Q):
    rng = np.random.default_rng(3)
    n, nb = 64, 9
    psi = rand_c(rng, nb, n)
    for nbb, nn_ in ((9, 64), (700, 300), (1, 5)):
        ps = rand_c(rng, nbb, nn_)
        pop = Q.ensemble_population(Q.DeviceBatch.from_host(ps)).cpu().numpy()
        assert np.allclose(pop, (np.abs(ps) ** 2).sum(axis=0), rtol=1e-13, atol=1e-13)
    for nobs in (1, 3, 7, 16):
        d = rng.standard_normal((nobs, n))
        got = Q.expect_diag(list(d), Q.DeviceBatch.from_host(psi)).cpu().numpy()
        ref = (np.abs(psi) ** 2) @ d.T
        assert np.allclose(got, ref, rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("degenerate", [False, True])
def test_ame_device_gap_tables(Q, degenerate):
    """Row G on the device (ame_gaps.cu): rounded-gap table, Ohmic γ table and multi-pair group detection on the GPU must
    reproduce the host-built tables' RHS (and the oracle's literal GapIndices loop) — including degenerate spectra whose
    gap groups hold several pairs (cross terms)."""
    rng = np.random.default_rng(123)
    nq, lvl, nb = 4, 16, 3
    n = 2 ** nq
    Hd = -F.standard_driver(nq)
    if degenerate:  # uniform ZZ chain: highly degenerate gaps
        Hp = -F.two_local_term([1.0] * (nq - 1), [[i, i + 1] for i in range(1, nq)], nq)
    else:
        Hp = F.two_local_term(list(rng.standard_normal(nq - 1)), [[i, i + 1] for i in range(1, nq)], nq) \
            + sum(rng.standard_normal() * F.single_clause(["z"], [k + 1], 1.0, nq) for k in range(nq))
    coup = [F.single_clause(["z"], [k + 1], 1.0, nq) for k in range(2)]
    Hg, Ho = Q.DenseHamiltonian([A, B], [Hd, Hp]), O.DenseHamiltonian([A, B], [Hd, Hp])
    bath_g, bath_o = Q.Ohmic(1e-2, 4, 16), O.Ohmic(1e-2, 4, 16)
    mk = lambda: Q.DiffEqLiouvillian(Hg, [Q.DaviesGenerator(Q.ConstantCouplings(coup, unit="h"), bath_g, Q.ZERO_LAMBSHIFT)], [], lvl)
    op_host, op_dev = mk(), mk()
    op_dev.device_tables = True
    u = rand_rho(rng, nb, n)
    s = 0.35 if not degenerate else 1.0  # s = 1: H = H_p exactly (diagonal, degenerate)
    w, v = O.haml_eigs(Ho, s, lvl)
    p = Q.ODEParams(Hg, 2.0)
    d_host = op_host.rhs_with_eig(np.zeros_like(u), u, p, 2.0 * s, w, v)
    d_dev = op_dev.rhs_with_eig(np.zeros_like(u), u, p, 2.0 * s, w, v)
    assert relerr(d_dev, d_host) < 1e-13
    opo = O.DiffEqLiouvillian(Ho, [O.DaviesGenerator([2 * np.pi * c for c in coup], bath_o.spectrum, lambda x: 0.0)], [], lvl)
    ref = np.stack([opo.rhs_with_eig(u[b], O.ODEParams(Ho, 2.0), 2.0 * s, w, v) for b in range(nb)])
    assert relerr(d_dev, ref) < TOL
    if degenerate:
        G = O.build_gap_indices(w, 8, 8)
        assert any(len(a) > 1 for a in G.uniq_a)  # the case really has multi-pair groups


def _rk4_oracle(rhs, u, t0, dt, nsteps):
    for i in range(nsteps):
        t = t0 + i * dt
        k1 = rhs(u, t)
        k2 = rhs(u + 0.5 * dt * k1, t + dt / 2)
        k3 = rhs(u + 0.5 * dt * k2, t + dt / 2)
        k4 = rhs(u + dt * k3, t + dt)
        u = u + (dt / 6) * k1 + (dt / 3) * k2 + (dt / 3) * k3 + (dt / 6) * k4
    return u


@pytest.mark.parametrize("kind", ["lindblad", "ame"])
def test_density_stepper_rk4(Q, ame2q, kind):
    """ρ stays on the device through a fixed-step RK4 solve (four `(du,u,p,t)` stages + axpy kernels per step); the
    oracle integrates the same scheme with its own RHS.  Time-dependent H: the AME operator re-diagonalises per stage."""
    rng = np.random.default_rng(17)
    tf, dt, nsteps, nb = 2.0, 0.02, 25, 3
    if kind == "lindblad":
        n = 8
        Hd, Hp = -F.standard_driver(3), F.two_local_term([1.0, -0.5], [[1, 2], [2, 3]], 3)
        Ls = [F.single_clause(["z"], [k + 1], 1.0, 3) for k in range(3)]
        gfs = [(lambda s, k=k: 0.2 * (1 + k) * (1 + s)) for k in range(3)]
        Hg, Ho = Q.DenseHamiltonian([A, B], [Hd, Hp]), O.DenseHamiltonian([A, B], [Hd, Hp])
        opg = Q.DiffEqLiouvillian(Hg, [], [Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gfs, Ls)])], n)
        opo = O.DiffEqLiouvillian(Ho, [], [O.LindbladLiouvillian([O.Lindblad(g, L) for g, L in zip(gfs, Ls)])], n)
    else:
        n = 4
        Hg, Ho = Q.DenseHamiltonian([A, B], ame2q["mats"]), ame2q["Ho"]
        coup_raw = [0.03 * F.q_translate("ZI+IZ")]  # weak coupling: the mock rates stay inside RK4's stability region
        opg = Q.DiffEqLiouvillian(Hg, [Q.DaviesGenerator(Q.ConstantCouplings(coup_raw, unit="h"), gamma, lamb)], [], 4)
        opo = O.DiffEqLiouvillian(Ho, [O.DaviesGenerator([0.03 * c for c in ame2q["coup"]], gamma, lamb)], [], 4)
    u0 = rand_rho(rng, nb, n)
    ud = Q.DeviceBatch.from_host(u0)
    Q.DensityStepper(opg).evolve(ud, Q.ODEParams(Hg, tf), 0.1, dt, nsteps)
    got = ud.to_host()
    po = O.ODEParams(Ho, tf)
    ref = np.stack([_rk4_oracle(lambda x, t: opo.rhs(x, po, t), u0[b], 0.1, dt, nsteps) for b in range(nb)])
    assert relerr(got, ref) < 1e-11
    assert np.allclose(np.trace(got, axis1=1, axis2=2), 1.0, atol=1e-10)  # trace preserved by every RHS on the path


def test_lindblad_time_dependent_operators(Q):
    """Lindblad(γ, L) with L a FUNCTION of s (src/coupling/interaction.jl:42,45-53): evaluated on the host per distinct s and
    uploaded per call; batches with several s values are processed per s-group."""
    rng = np.random.default_rng(64)
    n, nb = 8, 6
    Hd, Hp = -F.standard_driver(3), F.two_local_term([1.0, -0.7], [[1, 2], [2, 3]], 3)
    Z1, X2 = F.single_clause(["z"], [1], 1.0, 3), F.single_clause(["x"], [2], 1.0, 3)
    Lf = [lambda s: (1 - s) * Z1 + s * X2, lambda s: np.cos(s) * X2 + 1j * s * Z1]
    gfs = [lambda s: 0.3 + s, lambda s: 0.1]
    Hg, Ho = Q.DenseHamiltonian([A, B], [Hd, Hp]), O.DenseHamiltonian([A, B], [Hd, Hp])
    lg = Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gfs, Lf)] + [Q.Lindblad(0.2, Z1)])
    lo = O.LindbladLiouvillian([O.Lindblad(g, L) for g, L in zip(gfs, Lf)] + [O.Lindblad(0.2, Z1)])
    opg, opo = Q.DiffEqLiouvillian(Hg, [], [lg], n), O.DiffEqLiouvillian(Ho, [], [lo], n)
    u = rand_rho(rng, nb, n)
    tf = 2.0
    pg, po = Q.ODEParams(Hg, tf), O.ODEParams(Ho, tf)
    for t in (0.8, np.array([0.2, 1.4, 0.2, 0.9, 1.4, 1.9])):
        got = opg(np.zeros_like(u), u, pg, t)
        tv = np.broadcast_to(t, (nb,))
        ref = np.stack([opo.rhs(u[b], po, float(tv[b])) for b in range(nb)])
        assert relerr(got, ref) < TOL
    cache = opg.update_cache(np.zeros((n, n), complex), pg, 0.8)
    assert relerr(cache, opo.update_cache(po, 0.8)) < 1e-13


def test_device_against_committed_golden_file(Q):
    """The CUDA path against tests/golden/reference_known_answers.json (the reference's own known answers, transcribed by
    tests/golden/make_golden.py): Lindblad RHS and cache, DenseHamiltonian values, CorrelatedDavies closed forms, Redfield Λ."""
    import json
    import os
    import scipy.linalg as sla
    from oqb_b200.extra_liouvillians import CorrelatedDaviesGenerator
    from oqb_b200.redfield_device import LambdaBuilder
    G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_known_answers.json")))
    cm = lambda x: np.asarray(x, dtype=float)[..., 0] + 1j * np.asarray(x, dtype=float)[..., 1]
    ops = {"sx": sx, "sz": sz, "sigma_plus": np.array([[0, 1], [0, 0]], dtype=complex),
           "sigma_minus": np.array([[0, 0], [1, 0]], dtype=complex)}
    states = {"|+x><+x|": np.outer(F.PauliVec[0][0], F.PauliVec[0][0].conj()),
              "|+y><+y|": np.outer(F.PauliVec[1][0], F.PauliVec[1][0].conj())}
    p = Q.ODEParams(None, 5.0)
    for key, ckey in (("lindblad_rhs_Lz_on_plus_x", "lindblad_cache_Lz"), ("lindblad_rhs_Lz_Lx_on_plus_y", "lindblad_cache_Lz_Lx")):
        g = G[key]
        L = Q.LindbladLiouvillian([Q.Lindblad(gm, ops[o]) for gm, o in zip(g["gamma"], g["ops"])])
        assert np.allclose(L(np.zeros((2, 2), complex), states[g["state"]], p, 5.0), cm(g["drho"]), atol=1e-15)
        assert np.allclose(L.update_cache(np.zeros((2, 2), complex), p, 5.0), cm(G[ckey]["cache"]), atol=1e-15)
    g = G["dense_hamiltonian_half"]
    H = Q.DenseHamiltonian([A, B], [sx, sz])
    assert np.allclose(H(g["s"]), cm(g["H"]), atol=1e-15)
    assert np.allclose(H.update_cache(np.zeros((2, 2), complex), g["s"]), cm(g["cache"]), atol=1e-15)
    for key in ("correlated_davies_zero", "correlated_davies_closed_form"):
        g = G[key]
        gfun = lambda w: 1.0 if w >= 0 else np.exp(-0.5)
        inds = [tuple(x) for x in g["inds"]]
        D = CorrelatedDaviesGenerator([ops[c] for c in g["couplings"]], {i: gfun for i in inds}, {i: (lambda w: 0.0) for i in inds}, inds)
        assert np.allclose(D(np.zeros((2, 2), complex), cm(g["rho"]), np.array(g["w"]), g["s"]), cm(g["du"]), atol=1e-15)
    g = G["redfield_lambda_code_form"]  # analytic unitary sampled finely: the grid error stays below the fixture tolerance × 100
    Gn = 4097
    Us = np.stack([sla.expm(-1j * (5.0 * k / (Gn - 1)) * sx) for k in range(Gn)])
    Lam = LambdaBuilder([sz], Us[None], 5.0 / (Gn - 1)).build([0], [0], 5.0, lambda q, t, x: np.ones_like(x), 5.0, 1e-8, 1e-6).to_host()[0]
    assert np.allclose(Lam, cm(g["Lambda"]), atol=1e-5)
    g = G["ohmic_lambshift_at_zero"]  # device quadrature against the reference's cpvagk value
    assert np.isclose(Q.OhmicBath(g["eta"], g["wc"], g["beta"]).S(g["w"]), g["S"], rtol=g["rtol"], atol=g["atol"])
    g = G["bspline_linear_data"]  # device spline evaluation reproduces linear data and fills 0 outside
    x = np.linspace(g["grid"]["lo"], g["grid"]["hi"], g["grid"]["length"])
    itp = Q.QuadraticBSpline(x[0], x[1] - x[0], 10 - x)
    tx = np.linspace(g["test_grid"]["lo"], g["test_grid"]["hi"], g["test_grid"]["length"])
    assert np.allclose(itp.on_device(tx), 10 - tx)
    assert np.array_equal(itp.on_device(np.array([xo for xo, _ in g["outside"]], dtype=float)), [yo for _, yo in g["outside"]])


def test_adiabatic_frame_with_const_davies(Q):
    """DiffEqLiouvillian{false,true} (src/opensys/diffeq_liouvillian.jl:142-158): adiabatic-frame Hamiltonian + a
    ConstDaviesGenerator in `opensys` — RHS and update_cache! against the oracle."""
    from oqb_b200.extra_liouvillians import build_const_davies
    rng = np.random.default_rng(606)
    n, nb = 4, 3
    wfun = lambda s: np.array([-1.0, -0.2 + 0.1 * s, 0.4, 1.3 - 0.2 * s])
    geo = lambda s: 0.05 * np.array([[0, 1j, 0, 0], [-1j, 0, 1j, 0], [0, -1j, 0, 1j], [0, 0, -1j, 0]])
    cs = [(lambda x: x + x.conj().T)(rand_c(rng, n, n)) * 0.2]
    w0 = 2 * np.pi * wfun(0.0)
    Dg = build_const_davies(cs, w0, gamma, lamb)
    Do = O.build_const_davies(cs, O.build_gap_indices(w0, 8, 8), gamma, lamb)
    Hg = Q.AdiabaticFrameHamiltonian(wfun, geo)
    opg = Q.DiffEqLiouvillian(Hg, [], [Dg], n)

    class HoFrame:  # H(tf, s) of the adiabatic frame for the oracle: 2π diag(ω(s)) + G(s)/tf
        size = (n, n)

        def __call__(self, tf, s):
            return np.diag(2 * np.pi * wfun(s)).astype(complex) + geo(s) / tf
    opo = O.DiffEqLiouvillian(HoFrame(), [], [Do], n, adiabatic_frame=True)
    u = rand_rho(rng, nb, n)
    tf = 3.0
    pg, po = Q.ODEParams(None, tf), O.ODEParams(None, tf)
    got = opg(rand_c(rng, nb, n, n), u, pg, 1.2)  # overwrites du
    ref = np.stack([opo.rhs(u[b], po, 1.2) for b in range(nb)])
    assert relerr(got, ref) < TOL
    cache = opg.update_cache(np.zeros((n, n), complex), pg, 1.2)
    assert relerr(cache, opo.update_cache(po, 1.2)) < 1e-13


@pytest.mark.gpu
def test_ohmic_lambshift_device(Q):
    """Row B: S(ω, bath) = lambshift(ω, γ) (bath_util.jl:17, ext_util.jl:23-28) by the per-thread adaptive G7K15 of
    csrc/bath.cu against the oracle's quadgk restatement.  Both stop at QuadGK's default rtol = √eps ≈ 1.5e-8, and the
    device exp() may differ from libm's in the last bit, so the two adaptive runs are only REQUIRED to agree to the
    quadrature tolerance (asserted: 1e-7 relative); in practice they bisect identically and agree to ~1e-13, which is
    asserted for the median.  Known answer: test/coupling_bath_interaction/bath.jl:14-16."""
    bath_g, bath_o = Q.OhmicBath(1e-4, 8 * np.pi, 1 / 2.23), O.OhmicBath(1e-4, 8 * np.pi, 1 / 2.23)
    assert np.isclose(bath_g.S(0.0), -0.0025132734115775254, rtol=1e-4, atol=1e-4)
    bath_g, bath_o = Q.Ohmic(1e-4, 4, 16), O.Ohmic(1e-4, 4, 16)
    ws = np.concatenate([np.linspace(-50, 50, 41), [1e-10, -1e-3, 0.37]])
    S_dev, err, nseg = bath_g.S(ws, info=True)
    S_ref = np.array([O.lambshift(w, bath_o.spectrum) for w in ws])
    rel = np.abs(S_dev - S_ref) / np.abs(S_ref)
    assert rel.max() < 1e-7 and np.median(rel) < 1e-12
    assert np.all(err <= 1.5e-8 * np.abs(S_dev) * 1.0001) and nseg.max() < 1024  # every integral converged
    # tighter request → more segments (ω = 1e-10 sits on the isapprox(ω, 0) switch of ohmic.jl:36 and runs into the
    # segment cap); the result moves by no more than a few times the first request's tolerance
    S_tight, _, nseg2 = bath_g.S(ws, rtol=1e-12, info=True)
    assert np.all(nseg2 >= nseg) and np.max(np.abs(S_tight - S_dev) / np.abs(S_dev)) < 1e-6


@pytest.mark.gpu
def test_bspline_device_eval(Q):
    """construct_interpolations' quadratic B-spline (interpolation.jl:22-34) evaluated by the device routine the gap
    tables use, against the oracle's restatement: grid points, midpoints, random points, both ends, outside (fill 0)."""
    rng = np.random.default_rng(77)
    x0, dx, n = -5.0, 10 / 200, 201
    y = np.sin(np.linspace(-5, 5, n)) * np.exp(-0.1 * np.linspace(-5, 5, n) ** 2)
    sp_g, sp_o = Q.QuadraticBSpline(x0, dx, y), O.QuadraticBSpline(x0, dx, y)
    assert np.allclose(sp_g.coef, sp_o.c, rtol=0, atol=1e-14)
    xs = np.concatenate([x0 + dx * np.arange(n), x0 + dx * (np.arange(n - 1) + 0.5), rng.uniform(-5, 5, 500),
                         [-5.0, 5.0, -5.0000001, 5.0000001, -7.0, 9.0]])
    ref = np.array([sp_o(v) for v in xs]).real
    assert np.max(np.abs(sp_g.on_device(xs) - ref)) < 1e-14
    assert np.max(np.abs(sp_g.vectorized(xs) - ref)) < 1e-14
    assert np.max(np.abs(sp_g.vectorized(x0 + dx * np.arange(n)) - y)) < 1e-13  # interpolates the data


@pytest.mark.gpu
@pytest.mark.parametrize("degenerate", [False, True])
def test_ame_device_tables_tabulated_lambshift(Q, degenerate):
    """DaviesGenerator with the Lamb shift build_lambshift tabulates over ω_range (bath_util.jl:26-38): the S(ω̃_ab)
    table is evaluated from the spline ON THE DEVICE (device_tables) and must reproduce both the host-table path and
    the oracle's literal loop with its own spline through the same samples."""
    rng = np.random.default_rng(321)
    nq, lvl, nb = 4, 16, 3
    n = 2 ** nq
    Hd = -F.standard_driver(nq)
    if degenerate:
        Hp = -F.two_local_term([1.0] * (nq - 1), [[i, i + 1] for i in range(1, nq)], nq)
    else:
        Hp = F.two_local_term(list(rng.standard_normal(nq - 1)), [[i, i + 1] for i in range(1, nq)], nq) \
            + sum(rng.standard_normal() * F.single_clause(["z"], [k + 1], 1.0, nq) for k in range(nq))
    coup = [F.single_clause(["z"], [k + 1], 1.0, nq) for k in range(2)]
    Hg, Ho = Q.DenseHamiltonian([A, B], [Hd, Hp]), O.DenseHamiltonian([A, B], [Hd, Hp])
    bath_g, bath_o = Q.Ohmic(1e-2, 4, 16), O.Ohmic(1e-2, 4, 16)
    w_range = np.linspace(-30, 30, 601)  # narrower than the spectrum for s = 0.35: some gaps fall outside → S = 0 there
    S_g = Q.build_lambshift(w_range, True, bath_g)
    assert isinstance(S_g, Q.QuadraticBSpline)
    samples = S_g.vectorized(w_range)
    S_o = O.QuadraticBSpline(w_range[0], (w_range[-1] - w_range[0]) / 600, samples)
    mk = lambda: Q.DiffEqLiouvillian(Hg, [Q.DaviesGenerator(Q.ConstantCouplings(coup, unit="h"), bath_g, S_g)], [], lvl)
    op_host, op_dev = mk(), mk()
    op_dev.device_tables = True
    u = rand_rho(rng, nb, n)
    s = 0.35 if not degenerate else 1.0
    w, v = O.haml_eigs(Ho, s, lvl)
    p = Q.ODEParams(Hg, 2.0)
    d_host = op_host.rhs_with_eig(np.zeros_like(u), u, p, 2.0 * s, w, v)
    d_dev = op_dev.rhs_with_eig(np.zeros_like(u), u, p, 2.0 * s, w, v)
    assert relerr(d_dev, d_host) < 1e-13
    opo = O.DiffEqLiouvillian(Ho, [O.DaviesGenerator([2 * np.pi * c for c in coup], bath_o.spectrum, lambda x: S_o(x).real)], [], lvl)
    ref = np.stack([opo.rhs_with_eig(u[b], O.ODEParams(Ho, 2.0), 2.0 * s, w, v) for b in range(nb)])
    assert relerr(d_dev, ref) < TOL
    # the Lamb shift matters in this case (the test would not notice a dropped table otherwise)
    op0 = Q.DiffEqLiouvillian(Hg, [Q.DaviesGenerator(Q.ConstantCouplings(coup, unit="h"), bath_g, Q.ZERO_LAMBSHIFT)], [], lvl)
    d0 = op0.rhs_with_eig(np.zeros_like(u), u, p, 2.0 * s, w, v)
    assert relerr(d0, ref) > 1e-6 or degenerate


@pytest.mark.gpu
@pytest.mark.parametrize("with_lambshift", [False, True])
def test_onesided_ame_device_tables(Q, with_lambshift):
    """OneSidedAMELiouvillian (davies.jl:367-379) with γ(w_b − w_a) and the spline Lamb shift evaluated at the l² UNROUNDED
    gaps on the device (oqb_ame_set_onesided_ohmic) against host-evaluated tables and the oracle."""
    rng = np.random.default_rng(654 + with_lambshift)
    nq, lvl, nb = 4, 16, 3
    n = 2 ** nq
    Hd = -F.standard_driver(nq)
    Hp = F.two_local_term(list(rng.standard_normal(nq - 1)), [[i, i + 1] for i in range(1, nq)], nq) \
        + sum(rng.standard_normal() * F.single_clause(["z"], [k + 1], 1.0, nq) for k in range(nq))
    coup = [F.single_clause(["z"], [k + 1], 1.0, nq) for k in range(3)]
    inds = [(1, 1), (2, 2), (3, 3), (1, 2)]
    Hg, Ho = Q.DenseHamiltonian([A, B], [Hd, Hp]), O.DenseHamiltonian([A, B], [Hd, Hp])
    bath_g, bath_o = Q.Ohmic(1e-2, 4, 16), O.Ohmic(1e-2, 4, 16)
    if with_lambshift:
        w_range = np.linspace(-30, 30, 301)
        S_g = Q.build_lambshift(w_range, True, bath_g)
        S_o_sp = O.QuadraticBSpline(w_range[0], 0.2, S_g.vectorized(w_range))
        S_o = lambda x: S_o_sp(x).real
    else:
        S_g, S_o = Q.ZERO_LAMBSHIFT, (lambda x: 0.0)
    mk = lambda: Q.DiffEqLiouvillian(Hg, [Q.OneSidedAMELiouvillian(Q.ConstantCouplings(coup, unit="h"), bath_g, S_g, inds)], [], lvl)
    op_host, op_dev = mk(), mk()
    op_dev.device_tables = True
    u = rand_rho(rng, nb, n)
    w, v = O.haml_eigs(Ho, 0.4, lvl)
    p = Q.ODEParams(Hg, 2.0)
    d_host = op_host.rhs_with_eig(np.zeros_like(u), u, p, 0.8, w, v)
    d_dev = op_dev.rhs_with_eig(np.zeros_like(u), u, p, 0.8, w, v)
    assert relerr(d_dev, d_host) < 1e-13
    opo = O.DiffEqLiouvillian(Ho, [O.OneSidedAMELiouvillian([2 * np.pi * c for c in coup], bath_o.spectrum, S_o, inds)], [], lvl)
    ref = np.stack([opo.rhs_with_eig(u[b], O.ODEParams(Ho, 2.0), 0.8, w, v) for b in range(nb)])
    assert relerr(d_dev, ref) < TOL


@pytest.mark.gpu
def test_ohmic_correlation_and_timescales_device(Q):
    """correlation(τ, ::OhmicBath) (ohmic.jl:48-52: complex trigamma) and τ_SB / τ_B / coarse_grain_timescale
    (bath_util.jl:55-105) on the device against the oracle and the reference's known answers
    (test/coupling_bath_interaction/bath.jl:18-26)."""
    bath_g, bath_o = Q.Ohmic(1e-4, 4, 16), O.Ohmic(1e-4, 4, 16)
    taus = np.concatenate([[0.0, 1e-6, 1e-3], np.linspace(0.01, 5, 60), [20.0, 100.0, 1e3, -0.7]])
    Cd = bath_g.correlation(taus)
    Co = np.array([bath_o.correlation(t) for t in taus])
    assert np.max(np.abs(Cd - Co) / np.abs(Co)) < 1e-13
    assert abs(bath_g.correlation(0.3) - bath_o.correlation(0.3)) < 1e-13 * abs(bath_o.correlation(0.3))
    (tsb, esb), (tb, eb) = bath_g.timescales(100)
    assert np.isclose(tsb, 284.61181493, rtol=1e-6, atol=1e-6) and np.isclose(tb, 0.07638653, rtol=1e-6, atol=1e-6)
    tsb_o, esb_o = O.tau_SB(bath_o.correlation)
    tb_o, eb_o = O.tau_B(bath_o.correlation, 100, tsb_o)
    assert abs(tsb - tsb_o) < 1e-7 * tsb_o and abs(tb - tb_o) < 1e-7 * tb_o  # quadrature tolerance √eps; typically ~1e-13
    ta, ta_err = bath_g.coarse_grain_timescale(100)
    assert np.isclose(ta, np.sqrt(tsb * tb / 5), rtol=1e-12) and ta_err < 1e-6


@pytest.mark.gpu
def test_density_stepper_prefetched_eigensystems(Q):
    """DensityStepper with the eigen-systems of the coming stage times prepared on a side stream (own context, worker
    thread, oqb_ame_rebind on adoption) gives the same trajectory as preparing them in line; handles are recycled
    through the shared pool (plan_cache smaller than the number of stage times)."""
    rng = np.random.default_rng(808)
    nq, nb = 4, 3
    n = 2 ** nq
    Hd = -F.standard_driver(nq)
    Hp = F.two_local_term(list(rng.standard_normal(nq - 1)), [[i, i + 1] for i in range(1, nq)], nq) \
        + sum(rng.standard_normal() * F.single_clause(["z"], [k + 1], 1.0, nq) for k in range(nq))
    coup = [0.05 * F.single_clause(["z"], [k + 1], 1.0, nq) for k in range(2)]
    Hg = Q.DenseHamiltonian([A, B], [Hd, Hp])
    bath = Q.Ohmic(1e-2, 4, 16)
    p = Q.ODEParams(Hg, 2.0)
    u0 = rand_rho(rng, nb, n)
    res = {}
    for mode in ("inline", "prefetch"):
        op = Q.DiffEqLiouvillian(Hg, [Q.DaviesGenerator(Q.ConstantCouplings(coup, unit="h"), bath, Q.ZERO_LAMBSHIFT)], [], n)
        op.device_eigh, op.device_tables, op.plan_cache = True, True, 3
        op.prefetch_stages = mode == "prefetch"
        u = Q.DeviceBatch.from_host(u0.copy(), Q.get_context())
        Q.DensityStepper(op).evolve(u, p, 0.2, 0.01, 6)
        res[mode] = u.to_host()
        if mode == "prefetch":
            assert op.__dict__.get("_prefetcher") is not None  # the side stream really was used
        op.clear_plans()
        assert op.__dict__.get("_prefetcher") is None
    assert relerr(res["prefetch"], res["inline"]) < 1e-10
    tr = np.trace(res["prefetch"], axis1=1, axis2=2)
    assert np.max(np.abs(tr - 1)) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("lvl", [128, 16])
def test_ame_real_eigenvector_products(Q, lvl):
    """Real-symmetric H(s) ⇒ exactly real eigenvectors ⇒ the rotations v'ρv, vXv' run the real-operand product
    (2 DMMAs per complex product, zgemm_tma.cu RM modes).  Same RHS as the general complex product (1e-13) and as the
    oracle (1e-12); a Hamiltonian with a σy term keeps the general path."""
    import ctypes
    rng = np.random.default_rng(99 + lvl)
    nq, nb = 7, 2
    n = 2 ** nq
    Hd = -F.standard_driver(nq)
    Hp = F.two_local_term(list(rng.standard_normal(nq - 1)), [[i, i + 1] for i in range(1, nq)], nq) \
        + sum(rng.standard_normal() * F.single_clause(["z"], [k + 1], 1.0, nq) for k in range(nq))
    coup = [F.single_clause(["z"], [k + 1], 1.0, nq) for k in range(2)]
    ctx = Q.get_context()
    Hg, Ho = Q.DenseHamiltonian([A, B], [Hd, Hp]), O.DenseHamiltonian([A, B], [Hd, Hp])
    bath_g, bath_o = Q.Ohmic(1e-2, 4, 16), O.Ohmic(1e-2, 4, 16)
    mk = lambda H: Q.DiffEqLiouvillian(H, [Q.DaviesGenerator(Q.ConstantCouplings(coup, unit="h"), bath_g, Q.ZERO_LAMBSHIFT)], [], lvl)
    u = rand_rho(rng, nb, n)
    p = Q.ODEParams(Hg, 2.0)
    w, v = O.haml_eigs(Ho, 0.45, lvl)
    assert not np.any(v.imag)  # LAPACK keeps real data real
    res = {}
    try:
        for on in (True, False):
            ctx.set_real_operand_mode(on)
            op = mk(Hg)
            res[on] = op.rhs_with_eig(np.zeros_like(u), u, p, 0.9, w, v)
            flag = ctypes.c_int()
            ctx.check(ctx.lib.oqb_ame_eigenvectors_real(op._ame, ctypes.byref(flag)))
            assert flag.value == 1
            cache = op.update_cache(np.zeros((n, n), complex), p, 0.9, w, v)
            res[("cache", on)] = cache
            op.clear_plans()
    finally:
        ctx.set_real_operand_mode(True)
    assert relerr(res[True], res[False]) < 1e-13 and relerr(res[("cache", True)], res[("cache", False)]) < 1e-13
    if lvl == 16:
        opo = O.DiffEqLiouvillian(Ho, [O.DaviesGenerator([2 * np.pi * c for c in coup], bath_o.spectrum, lambda x: 0.0)], [], lvl)
        ref = np.stack([opo.rhs_with_eig(u[b], O.ODEParams(Ho, 2.0), 0.9, w, v) for b in range(nb)])
        assert relerr(res[True], ref) < TOL
    # device eigensolver on a real Hamiltonian: real symmetric solver, same RHS (gauge-invariant)
    op = mk(Hg)
    op.device_eigh = True
    d_dev = op(np.zeros_like(u), u, p, 0.9)
    flag = ctypes.c_int()
    ctx.check(ctx.lib.oqb_ame_eigenvectors_real(op._ame, ctypes.byref(flag)))
    assert flag.value == 1 and relerr(d_dev, res[True]) < 1e-9
    op.clear_plans()
    # complex Hamiltonian: not real, general product
    sy_term = F.single_clause(["y"], [1], 0.3, nq)
    Hc = Q.DenseHamiltonian([A, B, lambda s: 1.0], [Hd, Hp, sy_term])
    opc = mk(Hc)
    opc(np.zeros_like(u), u, Q.ODEParams(Hc, 2.0), 0.9)
    ctx.check(ctx.lib.oqb_ame_eigenvectors_real(opc._ame, ctypes.byref(flag)))
    assert flag.value == 0
    opc.clear_plans()


@pytest.mark.gpu
def test_onesided_and_jump_real_operand_products(Q):
    """One-sided AME with real couplings in a real eigenbasis and S ≡ 0: Λ_p, ΣA_βΛ_p and the A_β stack are exactly real
    and their products run the real-operand kernel (l = 128 > 64); ame_jump's v'ψ / vφ' too.  Against the general
    complex product (1e-13) and the oracle (1e-12)."""
    rng = np.random.default_rng(2024)
    nq, nb, lvl = 7, 2, 128
    n = 2 ** nq
    Hd = -F.standard_driver(nq)
    Hp = F.two_local_term(list(rng.standard_normal(nq - 1)), [[i, i + 1] for i in range(1, nq)], nq) \
        + sum(rng.standard_normal() * F.single_clause(["z"], [k + 1], 1.0, nq) for k in range(nq))
    coup = [F.single_clause(["z"], [k + 1], 1.0, nq) for k in range(3)]
    inds = [(1, 1), (2, 2), (3, 1)]
    ctx = Q.get_context()
    Hg, Ho = Q.DenseHamiltonian([A, B], [Hd, Hp]), O.DenseHamiltonian([A, B], [Hd, Hp])
    bath_g, bath_o = Q.Ohmic(1e-2, 4, 16), O.Ohmic(1e-2, 4, 16)
    u = rand_rho(rng, nb, n)
    p = Q.ODEParams(Hg, 2.0)
    w, v = O.haml_eigs(Ho, 0.45, lvl)
    res = {}
    try:
        for on in (True, False):
            ctx.set_real_operand_mode(on)
            op = Q.DiffEqLiouvillian(Hg, [Q.OneSidedAMELiouvillian(Q.ConstantCouplings(coup, unit="h"), bath_g, Q.ZERO_LAMBSHIFT, inds)], [], lvl)
            res[on] = op.rhs_with_eig(np.zeros_like(u), u, p, 0.9, w, v)
            op.clear_plans()
    finally:
        ctx.set_real_operand_mode(True)
    assert relerr(res[True], res[False]) < 1e-13
    opo = O.DiffEqLiouvillian(Ho, [O.OneSidedAMELiouvillian([2 * np.pi * c for c in coup], bath_o.spectrum, lambda x: 0.0, inds)], [], lvl)
    ref = np.stack([opo.rhs_with_eig(u[b], O.ODEParams(Ho, 2.0), 0.9, w, v) for b in range(nb)])
    assert relerr(res[True], ref) < TOL
    # ame_jump with l = 80 levels of the same real Hamiltonian
    _ame_jump_case(Q, [Hd, Hp], coup[:2], 80, 0.6, 3, rng, gamma)


@pytest.mark.gpu
def test_lindblad_real_hamiltonian_products(Q):
    """Real Hamiltonian terms + diagonal L_k: H_eff's imaginary diagonal moves into the Hadamard pass (−½(D_a + D_c)ρ_ac)
    and H(s)ρ, ρH(s) run the real-operand product.  Diagonals with complex phases, per-member s; against the general
    product (1e-13), the oracle (1e-12), and the commutator-only call."""
    rng = np.random.default_rng(31337)
    nq, K, nb = 7, 3, 4
    n = 2 ** nq
    Hd = -F.standard_driver(nq)
    Hp = F.two_local_term(list(rng.standard_normal(nq - 1)), [[i, i + 1] for i in range(1, nq)], nq)
    Ls = [np.diag(np.exp(1j * rng.uniform(0, 2 * np.pi, n)) * rng.uniform(0.5, 1.5, n)) for _ in range(K)]
    gfs = [lambda s, k=k: 0.2 * (1 + k) * (1 + s) for k in range(K)]
    fs = [A, B]
    ctx = Q.get_context()
    Hg, Ho = Q.DenseHamiltonian(fs, [Hd, Hp], unit="ħ"), O.DenseHamiltonian(fs, [Hd, Hp], unit="ħ")
    u = rand_rho(rng, nb, n)
    t = np.linspace(0.5, 9.0, nb)
    res = {}
    try:
        for on in (True, False):
            ctx.set_real_operand_mode(on)
            opg = Q.DiffEqLiouvillian(Hg, [], [Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gfs, Ls)])], n)
            res[on] = opg(np.zeros_like(u), u, Q.ODEParams(None, 10.0), t)
            res[("comm", on)] = Hg(np.zeros_like(u), u, None, 0.3)
    finally:
        ctx.set_real_operand_mode(True)
    assert relerr(res[True], res[False]) < 1e-13 and relerr(res[("comm", True)], res[("comm", False)]) < 1e-13
    opo = O.DiffEqLiouvillian(Ho, [], [O.LindbladLiouvillian([O.Lindblad(g, L) for g, L in zip(gfs, Ls)])], n)
    for b in range(nb):
        assert relerr(res[True][b], opo.rhs(u[b], O.ODEParams(None, 10.0), t[b])) < TOL
    h = Ho(0.3)
    assert relerr(res[("comm", True)][1], -1j * (h @ u[1] - u[1] @ h)) < TOL


@pytest.mark.gpu
def test_lindblad_real_jump_operators(Q):
    """General (non-diagonal) but REAL Lindblad operators (σ∓-type ladders, σx): both sandwich products L_kρ and (L_kρ)L_k'
    run the real-operand kernel (n = 128 > 64); against the general complex product and the oracle."""
    rng = np.random.default_rng(777)
    nq, K, nb = 7, 3, 3
    n = 2 ** nq
    Hd = -F.standard_driver(nq)
    Hp = F.two_local_term(list(rng.standard_normal(nq - 1)), [[i, i + 1] for i in range(1, nq)], nq)
    sm = np.array([[0, 0], [1, 0]], dtype=complex)
    Ls = [F.single_clause(["x"], [1], 1.0, nq).astype(complex), np.kron(np.eye(2 ** 3), np.kron(sm, np.eye(2 ** 3))),
          rng.standard_normal((n, n)).astype(complex) / np.sqrt(n)]
    gfs = [lambda s, k=k: 0.3 * (1 + k) for k in range(K)]
    ctx = Q.get_context()
    Hg, Ho = Q.DenseHamiltonian([A, B], [Hd, Hp], unit="ħ"), O.DenseHamiltonian([A, B], [Hd, Hp], unit="ħ")
    u = rand_rho(rng, nb, n)
    t = np.linspace(1.0, 8.0, nb)
    res = {}
    try:
        for on in (True, False):
            ctx.set_real_operand_mode(on)
            opg = Q.DiffEqLiouvillian(Hg, [], [Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gfs, Ls)])], n)
            res[on] = opg(np.zeros_like(u), u, Q.ODEParams(None, 10.0), t)
    finally:
        ctx.set_real_operand_mode(True)
    assert relerr(res[True], res[False]) < 1e-13
    opo = O.DiffEqLiouvillian(Ho, [], [O.LindbladLiouvillian([O.Lindblad(g, L) for g, L in zip(gfs, Ls)])], n)
    for b in range(nb):
        assert relerr(res[True][b], opo.rhs(u[b], O.ODEParams(None, 10.0), t[b])) < TOL


@pytest.mark.gpu
def test_real_operand_products_ragged_sizes(Q):
    """Real-operand kernels on sizes that are no multiple of the 128×64 / 64×64 tiles or of the 8-deep k stage
    (n = 100, lvl = 77): out-of-bounds tile parts come from TMA zero fill and the guarded epilogue."""
    rng = np.random.default_rng(5150)
    n, lvl, nb = 100, 77, 3
    Asym = rng.standard_normal((n, n))
    H0 = (Asym + Asym.T) / 2
    Bsym = rng.standard_normal((n, n))
    H1 = (Bsym + Bsym.T) / 2
    coup = []
    for _ in range(2):
        Cs = rng.standard_normal((n, n))
        coup.append((Cs + Cs.T) / 2)
    ctx = Q.get_context()
    Hg, Ho = Q.DenseHamiltonian([A, B], [H0, H1], unit="ħ"), O.DenseHamiltonian([A, B], [H0, H1], unit="ħ")
    w, v = O.haml_eigs(Ho, 0.3, lvl)
    u = rand_rho(rng, nb, n)
    p = Q.ODEParams(Hg, 1.0)
    res = {}
    try:
        for on in (True, False):
            ctx.set_real_operand_mode(on)
            op = Q.DiffEqLiouvillian(Hg, [Q.DaviesGenerator(Q.ConstantCouplings(coup, unit="ħ"), gamma, lamb)], [], lvl)
            res[on] = op.rhs_with_eig(np.zeros_like(u), u, p, 0.3, w, v)
            res[("c", on)] = Hg(np.zeros_like(u), u, None, 0.3)
            op.clear_plans()
    finally:
        ctx.set_real_operand_mode(True)
    assert relerr(res[True], res[False]) < 1e-13 and relerr(res[("c", True)], res[("c", False)]) < 1e-13
    opo = O.DiffEqLiouvillian(Ho, [O.DaviesGenerator(coup, gamma, lamb)], [], lvl)
    ref = np.stack([opo.rhs_with_eig(u[b], O.ODEParams(Ho, 1.0), 0.3, w, v) for b in range(nb)])
    assert relerr(res[True], ref) < TOL


@pytest.mark.gpu
def test_redfield_apply_real_dense_couplings(Q):
    """General Redfield apply (redfield.jl:65-68) with REAL dense couplings (σx-type): the caller's realness flag (bit 1 of
    S_shared) sends S_α(Λ_αu) and (Λ_αu)S_α through the real-operand product; n = 128 so that the kernel is taken.
    Against the general product (1e-13) and the oracle (1e-12); complex couplings keep the general path."""
    rng = np.random.default_rng(4096)
    n, K, nb = 128, 3, 3
    S = []
    for _ in range(K):
        x = rng.standard_normal((n, n))
        S.append(((x + x.T) / 2).astype(complex))
    Lam = rand_c(rng, nb, K, n, n) / n
    u = rand_c(rng, nb, n, n)
    du0 = rand_c(rng, nb, n, n)
    ctx = Q.get_context()
    rg = Q.RedfieldLiouvillian(None, None, 0, 0, 0)
    res = {}
    try:
        for on in (True, False):
            ctx.set_real_operand_mode(on)
            res[on] = rg.apply(du0.copy(), u, S, Lam)
    finally:
        ctx.set_real_operand_mode(True)
    assert relerr(res[True], res[False]) < 1e-13
    ref = np.stack([O.redfield_apply_acc(du0[b].copy(), u[b], S, list(Lam[b])) for b in range(nb)])
    assert relerr(res[True], ref) < TOL
    Sc = [s + 1j * np.diag(rng.standard_normal(n)) @ np.eye(n) * 0 + 0.1j * (s - s.T + np.triu(s, 1) - np.triu(s, 1).T) for s in S]
    Sc = [s + 0.1j * (np.triu(np.real(s), 1) - np.triu(np.real(s), 1).T) for s in S]  # Hermitian with imaginary parts
    du_c = rg.apply(du0.copy(), u, Sc, Lam)
    ref_c = np.stack([O.redfield_apply_acc(du0[b].copy(), u[b], Sc, list(Lam[b])) for b in range(nb)])
    assert relerr(du_c, ref_c) < TOL
