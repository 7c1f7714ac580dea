"""Pins the CPU oracle against every golden vector / known-answer test the reference's own
test-suite holds for the master-equation RHS path.  Each test cites the
reference test it reproduces (paths relative to /root/reference).  CPU only."""
import numpy as np
import pytest
import scipy.linalg as sla
import scipy.sparse as sps

from oracle import oqb_oracle as O
from oracle import fixtures as F

sx, sy, sz, si = F.sx, F.sy, F.sz, F.si
A = lambda s: 1 - s
B = lambda s: s


def gamma(x):  # test/opensys/davies.jl:7
    return x + 1 if x >= 0 else (1 - x) * np.exp(x)


def lamb(x):  # test/opensys/davies.jl:8
    return x + 0.1


def close(a, b, atol=1e-6, rtol=1e-6):
    """Julia isapprox for arrays: ‖a−b‖ ≤ max(atol, rtol·max(‖a‖,‖b‖))."""
    a = np.asarray(a); b = np.asarray(b)
    return np.linalg.norm(a - b) <= max(atol, rtol * max(np.linalg.norm(a), np.linalg.norm(b)))


# ---- test/hamiltonian/dense_hamiltonian.jl -------------------------------------------------
def test_dense_hamiltonian_values():
    H = O.DenseHamiltonian([A, B], [sx, sz])
    assert np.array_equal(H(0.5), np.pi * (sx + sz))                                  # :12
    assert np.array_equal(H.update_cache(0.5), -1j * np.pi * (sx + sz))               # :23-25
    T = -1j * np.pi * (sx + sz)
    assert np.array_equal(H.update_vectorized_cache(0.5), np.kron(si, T) - np.kron(T.T, si))  # :28-32
    rho = np.outer(F.PauliVec[0][0], F.PauliVec[0][0].conj())
    du = H.rhs(rho, 0.5)                                                              # :35-38
    assert np.allclose(du, -1j * np.pi * ((sx + sz) @ rho - rho @ (sx + sz)), rtol=1e-14, atol=0)
    w, v = O.eigen_decomp(H, 0.5)                                                     # :41-44
    assert np.allclose(w, np.array([-1, 1]) / np.sqrt(2))
    assert np.isclose(abs(v[:, 0].conj() @ np.array([1 - np.sqrt(2), 1]) / np.sqrt(4 - 2 * np.sqrt(2))), 1)
    assert np.isclose(abs(v[:, 1].conj() @ np.array([1 + np.sqrt(2), 1]) / np.sqrt(4 + 2 * np.sqrt(2))), 1)
    with pytest.raises(ValueError):                                                   # :50
        O.DenseHamiltonian([A, B], [sx, sz], unit="hh")
    Hst = O.DenseHamiltonian([A, B], [sx, sz], unit="ħ")                              # :59-77 (static path: same arithmetic)
    assert np.array_equal(Hst(0.5), (sx + sz) / 2)
    assert np.array_equal(Hst.update_cache(0.5), -0.5j * (sx + sz))


# ---- test/hamiltonian/sparse_hamiltonian.jl ------------------------------------------------
def test_sparse_hamiltonian_values():
    H = O.SparseHamiltonian([A, B], [sps.csc_matrix(sx), sps.csc_matrix(sz)])
    assert np.allclose(H(0).toarray(), 2 * np.pi * sx)                                # :26
    assert np.allclose(H(0.5).toarray(), np.pi * (sx + sz))                           # :28
    u = np.array([1.0 + 0j, 1]) / np.sqrt(2)
    rho = np.outer(u, u.conj())
    du = H.rhs(rho, 0.5)                                                              # :32-34
    assert np.allclose(du, -1j * np.pi * ((sx + sz) @ rho - rho @ (sx + sz)), rtol=1e-14)
    assert np.array_equal(H.update_cache(0.5).toarray(), -1j * np.pi * (sx + sz))     # :37-39
    T = -1j * np.pi * (sx + sz)
    assert np.array_equal(H.update_vectorized_cache(0.5).toarray(), np.kron(si, T) - np.kron(T.T, si))  # :42-45
    H2 = O.SparseHamiltonian([A, B], [sps.csc_matrix((np.kron(sx, si) + np.kron(si, sx)).real),
                                      sps.csc_matrix((0.1 * np.kron(sz, si) - np.kron(sz, sz)).real)])
    w, v = O.eigen_decomp(H2, 1.0)                                                    # :47-53
    assert np.allclose(w, [-1.1, -0.9])
    assert np.isclose(abs(v[-1, 0]), 1) and np.isclose(abs(v[0, 1]), 1)


# ---- test/hamiltonian/constant_hamiltonian.jl ----------------------------------------------
def test_constant_hamiltonian_values():
    one = lambda s: 1.0
    H1 = O.DenseHamiltonian([one], [sx], unit="ħ")
    H2 = O.DenseHamiltonian([one], [sx])
    H3 = O.SparseHamiltonian([one], [sps.csc_matrix(np.kron(sx, sx))])
    assert np.array_equal(H1.update_cache(0.2), -1j * sx)                             # :26
    assert np.array_equal(H2.update_cache(0.1), -2 * np.pi * 1j * sx)                 # :27
    assert np.array_equal(H3.update_cache(0.5).toarray(), -2 * np.pi * 1j * np.kron(sx, sx))  # :28
    assert np.array_equal(H1.update_vectorized_cache(0.2), -1j * O.vectorized_commutator(H1(0.1)))  # :36
    assert np.array_equal(H2.update_vectorized_cache(0.1), -1j * O.vectorized_commutator(H2(1)))    # :37
    assert np.allclose(H3.update_vectorized_cache(0.5).toarray(),
                       -1j * O.vectorized_commutator(H3(2).toarray()))                # :38
    r1 = np.ones((2, 2), dtype=complex) / 2
    r3 = np.ones((4, 4), dtype=complex) / 4
    assert np.allclose(H1.rhs(r1, 0), -1j * (H1(0) @ r1 - r1 @ H1(0)), atol=1e-15)    # :44
    assert np.allclose(H2.rhs(r1, 0.5), -1j * (H2(0.5) @ r1 - r1 @ H2(0.5)), atol=1e-15)  # :45
    h3 = H3(1).toarray()
    assert np.allclose(H3.rhs(r3, 1), -1j * (h3 @ r3 - r3 @ h3), atol=1e-15)          # :46


def test_constant_hamiltonian_interface():
    """test/hamiltonian/constant_hamiltonian.jl through the oracle's ConstantHamiltonian / Hamiltonian interface
    (src/hamiltonian/interface.jl:1-35, dense_hamiltonian.jl:134-168, sparse_hamiltonian.jl:166-203)."""
    H1 = O.ConstantHamiltonian(sx, unit="ħ")
    H2 = O.ConstantHamiltonian(sx)
    H3 = O.ConstantHamiltonian(sps.csc_matrix(np.kron(sx, sx)))
    Hs1, Hs3 = O.Hamiltonian(sx, unit="ħ"), O.Hamiltonian(sps.csc_matrix(np.kron(sx, sx)))
    assert isinstance(H1, O.ConstantDenseHamiltonian) and isinstance(H3, O.ConstantSparseHamiltonian) and H1.isconstant and H3.isconstant
    assert np.array_equal(H1(2), sx) and np.array_equal(Hs1(2), sx)                                 # :17
    assert np.array_equal(H2(1), 2 * np.pi * sx)                                                     # :18
    assert np.array_equal(H3(0.2).toarray(), 2 * np.pi * np.kron(sx, sx))                            # :19
    assert np.array_equal(Hs3(0.2).toarray(), H3(0.2).toarray())
    assert np.array_equal(H1.update_cache(0.2), -1j * sx)                                            # :26
    assert np.array_equal(H2.update_cache(0.1), -2 * np.pi * 1j * sx)                                # :27
    assert np.array_equal(H3.update_cache(0.5).toarray(), -2 * np.pi * 1j * np.kron(sx, sx))         # :28
    assert np.array_equal(H1.update_vectorized_cache(0.2), -1j * O.vectorized_commutator(H1(0.1)))   # :36
    assert np.allclose(H3.update_vectorized_cache(0.5).toarray(), -1j * O.vectorized_commutator(H3(2).toarray()))  # :38
    r1 = np.ones((2, 2), dtype=complex) / 2
    r3 = np.ones((4, 4), dtype=complex) / 4
    assert np.allclose(H1.rhs(r1, 0), -1j * (H1(0) @ r1 - r1 @ H1(0)), atol=1e-15)                   # :44
    h3 = H3(1).toarray()
    assert np.allclose(H3.rhs(r3, 1), -1j * (h3 @ r3 - r3 @ h3), atol=1e-15)                         # :46
    w1, _ = O.eigen_decomp(H1, 0)
    w2, _ = O.eigen_decomp(H2, 0.5)
    w3, _ = O.eigen_decomp(H3, 1, lvl=4)
    assert np.allclose(w1, np.array([-1, 1]) / 2 / np.pi) and np.allclose(w2, [-1, 1]) and np.allclose(w3, [-1, -1, 1, 1])  # :48-57
    # the time-dependent interface picks the storage by the matrix type (interface.jl:18-33)
    assert isinstance(O.Hamiltonian([lambda s: s], [sx]), O.DenseHamiltonian)
    assert isinstance(O.Hamiltonian([lambda s: s], [sps.csc_matrix(sx)]), O.SparseHamiltonian)


# ---- test/annealing.jl:19-22 ; test/coupling_bath_interaction/bath.jl:3-14 -----------------
def test_odeparams_and_ohmic():
    p = O.ODEParams(None, 10.0, lambda tf, t: t / tf)
    assert p(5) == 0.5
    eta, fc, T = 1e-4, 4, 16
    bath = O.Ohmic(eta, fc, T)
    assert bath.spectrum(0.0) == 2 * np.pi * eta / bath.beta
    # SURVEY App. C anchors
    assert bath.beta == pytest.approx(0.4773896944364562, rel=1e-15)
    assert bath.spectrum(0.0) == pytest.approx(0.0013161543662136847, rel=1e-14)
    assert bath.spectrum(8.8857659) == pytest.approx(0.003977576906265448, rel=1e-12)


# ---- test/opensys/lindblad.jl:25-50 --------------------------------------------------------
def test_lindblad_golden():
    p = O.ODEParams(None, 5.0, lambda tf, t: t / tf)
    rho = np.outer(F.PauliVec[0][0], F.PauliVec[0][0].conj())
    Lz = O.Lindblad(0.5, sz)
    Lx = O.Lindblad(0.2, sx)
    L1 = O.LindbladLiouvillian([Lz])
    d = L1.rhs_acc(np.zeros((2, 2), complex), rho, p, 5.0)
    assert np.allclose(d, [[0, -0.5], [-0.5, 0]], atol=1e-15)                          # :31
    c = L1.update_cache_acc(np.zeros((2, 2), complex), p, 5.0)
    assert np.array_equal(c, -0.25 * sz.conj().T @ sz)                                 # :33-35
    L2 = O.LindbladLiouvillian([Lz, Lx])
    rho_y = np.outer(F.PauliVec[1][0], F.PauliVec[1][0].conj())
    d = L2.rhs_acc(np.zeros((2, 2), complex), rho_y, p, 5.0)
    assert np.allclose(d, [[0, 0.7j], [-0.7j, 0]], atol=1e-15)                         # :42
    c = L2.update_cache_acc(np.zeros((2, 2), complex), p, 5.0)
    assert np.array_equal(c, -0.25 * sz.conj().T @ sz - 0.1 * sx.conj().T @ sx)        # :44-46
    rng = np.random.default_rng(1234)
    for _ in range(4):                                                                 # :48-50 membership
        k, prob = O.lind_jump(L2, F.PauliVec[0][0], p, 0.5, rng.random())
        assert k in (0, 1)
    assert np.allclose(prob, [0.5, 0.2])
    # inverse-CDF boundaries
    assert O.sample_weights([0.5, 0.2], 0.0) == 0
    assert O.sample_weights([0.5, 0.2], 0.5 / 0.7 - 1e-12) == 0
    assert O.sample_weights([0.5, 0.2], 0.5 / 0.7 + 1e-12) == 1
    assert O.sample_weights([0.5, 0.2], 1.0 - 1e-16) == 1


# ---- test/opensys/util.jl:3-9 --------------------------------------------------------------
def test_gap_indices_vs_find_unique_gap():
    w = [-3, 1, 3, 3, 4.5, 5.5, 8]
    g = O.build_gap_indices(w, 8, 8)
    uniq, indices, indices0 = O.find_unique_gap(w)
    assert g.uniq_w == uniq
    assert g.uniq_a == [[x[0] for x in i] for i in indices]
    assert g.uniq_b == [[x[1] for x in i] for i in indices]
    assert set(zip(g.a0, g.b0)) == set(indices0)


def test_round_sigdigits():
    assert O.round_sigdigits(8.05669612345, 8) == 8.0566961
    assert O.round_sigdigits(12.9873017464, 8) == 12.987302
    assert O.round_sigdigits(-0.00123456789012, 8) == -0.0012345679
    assert O.round_sigdigits(123456789.0, 8) == 123456790.0
    assert O.round_sigdigits(0.5, 8) == 0.5


# ---- test/opensys/davies.jl:11-61 (2-qubit dense AME) --------------------------------------
@pytest.fixture(scope="module")
def ame2q():
    H = O.DenseHamiltonian([A, B], [-F.standard_driver(2), 0.1 * np.kron(sz, si) + 0.5 * np.kron(sz, sz)])
    coupling = [2 * np.pi * F.q_translate("ZI+IZ")]          # ConstantCouplings(["ZI+IZ"]) unit :h
    ops = [2 * np.pi * (np.kron(sz, si) + np.kron(si, sz))]
    w, v = O.eigen_decomp(H, 0.5, lvl=4)
    w = 2 * np.pi * w
    psi = (v[:, 0] + v[:, 1] + v[:, 2]) / np.sqrt(3)
    rho = np.outer(psi, psi.conj())
    u = v.conj().T @ rho @ v
    g = O.build_gap_indices(w, 8, 8)
    return dict(H=H, coupling=coupling, ops=ops, w=w, v=v, psi=psi, rho=rho, u=u, g=g)


def test_ame2q_spectrum_and_groups(ame2q):
    # SURVEY App. C anchors
    assert np.allclose(np.sort(ame2q["w"]), [-6.49365087321, -1.56304517842, 1.56304517842, 6.49365087321], atol=1e-10)
    g = ame2q["g"]
    assert g.uniq_w == [3.1260904, 4.9306057, 8.0566961, 12.987302]
    assert g.uniq_a == [[2], [1, 3], [1, 2], [1]]
    assert g.uniq_b == [[3], [2, 4], [3, 4], [4]]


def test_davies_rhs_vs_ame_update_test(ame2q):
    d = ame2q
    davies = O.DaviesGenerator(d["coupling"], gamma, lamb)
    drho = O.ame_update_test(d["ops"], d["rho"], d["w"], d["v"], gamma, lamb)
    du = davies.rhs_acc(np.zeros((4, 4), complex), d["u"], d["g"], d["v"], 0.5)
    assert close(d["v"] @ du @ d["v"].conj().T, drho)                                  # :29-32
    # the two gap-grouping implementations differ at the 1e-8 level (SURVEY §4 note)
    rel = np.linalg.norm(d["v"] @ du @ d["v"].conj().T - drho) / np.linalg.norm(drho)
    assert rel < 1e-7


def test_onesided_vs_dev_tools(ame2q):
    d = ame2q
    one = O.OneSidedAMELiouvillian(d["coupling"], gamma, lamb, [(1, 1)])
    exp = O.onesided_ame_update_test(d["ops"], d["rho"], d["w"], d["v"], gamma, lamb)
    du = one.rhs_acc(np.zeros((4, 4), complex), d["u"], d["g"], d["v"], 0.5)
    assert close(d["v"] @ du @ d["v"].conj().T, exp)                                   # :34-37
    v = d["v"]
    exp2 = O.onesided_ame_update_test([v @ op @ v.conj().T for op in d["ops"]], d["rho"], d["w"], v, gamma, lamb)
    du2 = one.rhs_acc(np.zeros((4, 4), complex), d["u"], d["g"], None, 0.5)
    assert close(du2, v.conj().T @ exp2 @ v)                                           # :39-42


def test_davies_update_cache(ame2q):
    d = ame2q
    davies = O.DaviesGenerator(d["coupling"], gamma, lamb)
    expH = O.ame_trajectory_Heff_test(d["ops"], d["w"], d["v"], gamma, lamb)
    cache = davies.update_cache_acc(np.zeros((4, 4), complex), d["g"], d["v"], 0.5)
    assert close(d["v"] @ cache @ d["v"].conj().T, -1j * expH)                         # :44-47
    op = O.DiffEqLiouvillian(d["H"], [davies], [], 4)
    p = O.ODEParams(d["H"], 2.0, lambda tf, t: t / tf)
    cache = op.update_cache(p, 1)
    assert close(cache, -1j * (expH + d["H"](0.5)))                                    # :49-55


def test_full_ame_rhs_dense(ame2q):
    d = ame2q
    davies = O.DaviesGenerator(d["coupling"], gamma, lamb)
    op = O.DiffEqLiouvillian(d["H"], [davies], [], 4)
    p = O.ODEParams(d["H"], 2.0, lambda tf, t: t / tf)
    drho = O.ame_update_test(d["ops"], d["rho"], d["w"], d["v"], gamma, lamb)
    h = d["H"](0.5)
    du = op.rhs(d["rho"], p, 1)
    assert close(du, drho - 1j * (h @ d["rho"] - d["rho"] @ h))                        # :57-61
    # SURVEY App. C anchor
    assert np.linalg.norm(du) == pytest.approx(451.9527156261307, rel=1e-9)
    assert abs(np.trace(du)) < 1e-11


# ---- test/opensys/davies.jl:112-145 (4-qubit sparse H, lvl = 4 of 16) ----------------------
def test_full_ame_rhs_sparse_truncated():
    Hd = sps.csc_matrix(F.standard_driver(4))
    Hp = sps.csc_matrix(F.q_translate("-0.9ZZII+IZZI-0.9IIZZ"))
    H = O.SparseHamiltonian([A, B], [Hd, Hp])
    ops = [2 * np.pi * F.q_translate("ZIII+IZII+IIZI+IIIZ")]
    davies = O.DaviesGenerator(ops, gamma, lamb)
    w, v = O.eigen_decomp(H, 0.5, lvl=4)
    w = 2 * np.pi * np.real(w)
    assert np.allclose(w, [-14.651797, -12.298275, -8.368612, -6.015089], atol=2e-6)
    psi = (v[:, 0] + v[:, 1] + v[:, 2]) / np.sqrt(3)
    rho = np.outer(psi, psi.conj())
    drho = O.ame_update_test(ops, rho, w, v, gamma, lamb)
    expH = O.ame_trajectory_Heff_test(ops, w, v, gamma, lamb) + v @ np.diag(w) @ v.conj().T
    op = O.DiffEqLiouvillian(H, [davies], [], 4)
    p = O.ODEParams(H, 2.0, lambda tf, t: t / tf)
    du = op.rhs(rho, p, 1)
    h = H(0.5).toarray()
    assert close(du, drho - 1j * (h @ rho - rho @ h))                                  # :131-140
    cache = op.update_cache(p, 1)
    assert close(cache, -1j * expH)                                                    # :142-145
    assert np.linalg.matrix_rank(cache, tol=1e-9) == 4


# ---- test/opensys/davies.jl:147-171 (adiabatic frame, unsorted w) --------------------------
def test_ame_adiabatic_frame():
    class AF:
        size = (4, 4)

        def __call__(self, tf, s):  # AdiabaticFrameHamiltonian((s)->[0,s,1-s,1], nothing)
            return np.diag(2 * np.pi * np.array([0, s, 1 - s, 1], dtype=complex))
    H = AF()
    hmat = H(2.0, 0.4)
    w = np.real(np.diag(hmat))
    v = np.eye(4, dtype=complex)
    coup = lambda s: [2 * np.pi * (s * (np.kron(sx, si) + np.kron(si, sx)) + (1 - s) * (np.kron(sz, si) + np.kron(si, sz)))]
    davies = O.DaviesGenerator(coup, gamma, lamb)
    psi = (v[:, 0] + v[:, 1] + v[:, 2]) / np.sqrt(3)
    rho = np.outer(psi, psi.conj())
    drho = O.ame_update_test(coup(0.4), rho, w, v, gamma, lamb)
    p = O.ODEParams(H, 2.0, lambda tf, t: t / tf)
    op = O.DiffEqLiouvillian(H, [davies], [], 4, adiabatic_frame=True)
    du = op.rhs(rho, p, 0.4 * 2)
    assert close(du, drho - 1j * (hmat @ rho - rho @ hmat))


# ---- test/opensys/redfield.jl:3-24, 37-46 --------------------------------------------------
def test_redfield_golden():
    unitary = lambda t: sla.expm(-1j * t * sx)
    cfun = lambda t1, t2: 1.0
    rho = np.outer(F.PauliVec[0][0], F.PauliVec[0][0].conj())
    kernels = [(((1, 1),), [sz], cfun)]
    red = O.RedfieldLiouvillian(kernels, unitary, 5.0, 1e-8, 1e-6)
    p = O.ODEParams(None, 5.0, lambda tf, t: t / tf)
    Lam_test, _ = O.quadgk(lambda x: unitary(x).conj().T @ sz @ unitary(x), 0, 5, 1.5e-8, 0.0)
    exp = -(sz @ (Lam_test @ rho - rho @ Lam_test.conj().T) - (Lam_test @ rho - rho @ Lam_test.conj().T) @ sz)
    d = red.rhs_acc(np.zeros((2, 2), complex), rho, p, 5.0)
    assert close(d, exp)                                                               # :18-21
    Acache = red.update_vectorized_cache_acc(np.zeros((4, 4), complex), p, 5.0)
    assert np.allclose(Acache @ rho.flatten(order="F"), exp.flatten(order="F"), atol=1e-6)  # :22-24
    # Λ as the CODE computes it, and dρ = 0.544021 σx
    Lam = red.Lambda(p, 5.0, [sz], cfun, 1)
    a = np.sin(10) / 2
    b = (1 - np.cos(10)) / 2
    assert np.allclose(Lam, [[a, 1j * b], [-1j * b, -a]], atol=1e-7)
    assert np.allclose(Lam, [[-0.272011, 0.919536j], [-0.919536j, 0.272011]], atol=1e-6)
    assert np.allclose(d, -2 * a * sx, atol=1e-7)
    # apply-form == superoperator form on a generic ρ (column-major vec)
    rng = np.random.default_rng(7)
    r = F.random_density_matrix(2, rng)
    d1 = red.rhs_acc(np.zeros((2, 2), complex), r, p, 5.0)
    assert np.allclose(Acache @ r.flatten(order="F"), d1.flatten(order="F"), atol=1e-12)
    # half-interval variant (:37-46)
    A2 = red.update_vectorized_cache_acc(np.zeros((4, 4), complex), p, 2.5)
    L2, _ = O.quadgk(lambda x: unitary(x).conj().T @ sz @ unitary(x), 0, 2.5, 1.5e-8, 0.0)
    exp2 = -(sz @ (L2 @ rho - rho @ L2.conj().T) - (L2 @ rho - rho @ L2.conj().T) @ sz)
    assert np.allclose(A2 @ rho.flatten(order="F"), exp2.flatten(order="F"), atol=1e-6)


def test_quadgk_rule_exactness():
    # Kronrod-15 integrates degree ≤ 22 exactly, embedded Gauss-7 degree ≤ 13.
    for deg in range(0, 23):
        I, E = O.quadgk(lambda x, d=deg: np.array([x ** d]), -1.0, 2.0, 0.0, 1e300)
        exact = (2.0 ** (deg + 1) - (-1.0) ** (deg + 1)) / (deg + 1)
        assert I[0] == pytest.approx(exact, rel=2e-13)
    I, E = O.quadgk(lambda x: np.array([np.exp(-x * x) * np.cos(30 * x)]), 0.0, 3.0, 1e-12, 0.0)
    import scipy.integrate as si_
    ref = si_.quad(lambda x: np.exp(-x * x) * np.cos(30 * x), 0, 3, epsabs=1e-14, epsrel=1e-14, limit=500)[0]
    assert I[0] == pytest.approx(ref, abs=1e-13)


# ---- C1 anchor: 1-qubit annealing AME ---------------------------------------
def test_c1_single_qubit_ame_anchor():
    H = O.DenseHamiltonian([lambda s: -(1 - s), lambda s: s], [sx, sz])
    bath = O.Ohmic(1e-4, 4, 16)
    davies = O.DaviesGenerator([2 * np.pi * sz], bath.spectrum, lambda w: 0.0)
    op = O.DiffEqLiouvillian(H, [davies], [], 2)
    p = O.ODEParams(H, 10.0, lambda tf, t: t / tf)
    rho = np.outer(F.PauliVec[0][0], F.PauliVec[0][0].conj())
    du = op.rhs(rho, p, 5.0)
    assert du[0, 0] == pytest.approx(-0.030394341407546568, rel=1e-10)
    assert du[0, 1] == pytest.approx(-0.015496302491821998 - 3.1415926535897922j, rel=1e-10)
    assert du[1, 1] == pytest.approx(+0.030394341407546568, rel=1e-10)
    du1 = op.rhs(rho, p, 10.0)
    assert du1[0, 1] == pytest.approx(-0.05195969170118221 - 6.2831853071795845j, rel=1e-10)


# ---- test/opensys/davies.jl:173-194 — CorrelatedDaviesGenerator closed forms -----------------------------
def test_correlated_davies_golden():
    sp_, sm_ = np.array([[0, 1], [0, 0]], dtype=complex), np.array([[0, 0], [1, 0]], dtype=complex)  # σ₊, σ₋ (math_util.jl:7-8)
    gfun = lambda w: 1.0 if w >= 0 else np.exp(-0.5)
    zero = lambda w: 0.0
    gam = {(1, 2): gfun, (2, 1): gfun}
    S = {(1, 2): zero, (2, 1): zero}
    G = O.build_gap_indices(np.array([1.0, 2.0]), 8, 8)
    D = O.CorrelatedDaviesGenerator([sp_, sm_], gam, S, ((1, 2), (2, 1)))
    du = D.rhs_acc(np.zeros((2, 2), complex), np.array([[0.5, 0], [0, 0.5]], dtype=complex), G, None, 0.5)
    assert np.allclose(du, 0.0, atol=1e-15)                                                     # :184
    D = O.CorrelatedDaviesGenerator([sm_, sm_], gam, S, ((1, 2), (2, 1)))
    du = D.rhs_acc(np.zeros((2, 2), complex), 0.5 * np.ones((2, 2), dtype=complex), G, None, 0.5)
    e = np.exp(-0.5)
    assert np.allclose(du, [[-e, -0.5 * e], [-0.5 * e, e]], atol=1e-15)                         # :194


# ---- test/opensys/lindblad.jl:3-19 — ULELiouvillian --------------------------------------------------------
def test_ule_golden():
    unitary = lambda t: sla.expm(-1j * t * sx)
    rho = np.outer(F.PauliVec[0][0], F.PauliVec[0][0].conj())
    kernels = [(((1, 1),), [sz], lambda t1, t2: 1.0)]
    ule = O.ULELiouvillian(kernels, unitary, 5.0, 1e-8, 1e-6)
    p = O.ODEParams(None, 5.0, lambda tf, t: t / tf)
    L, _ = O.quadgk(lambda x: unitary(x).conj().T @ sz @ unitary(x), 0, 5, 1.5e-8, 0.0)   # :14
    d = ule.rhs_acc(np.zeros((2, 2), complex), rho, p, 5.0)
    exp = L @ rho @ L.conj().T - 0.5 * (L.conj().T @ L @ rho + rho @ L.conj().T @ L)
    assert close(d, exp)                                                                  # :19


# ---- tests/golden/reference_known_answers.json (transcribed from the reference's tests by tests/golden/make_golden.py) ----
def _gold():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_known_answers.json")) as f:
        return json.load(f)


def _cm(x):
    a = np.asarray(x, dtype=float)
    return a[..., 0] + 1j * a[..., 1]


def test_oracle_against_committed_golden_file():
    G = _gold()
    ops = {"sx": sx, "sz": sz, "sigma_plus": np.array([[0, 1], [0, 0]], dtype=complex),
           "sigma_minus": np.array([[0, 0], [1, 0]], dtype=complex)}
    states = {"|+x><+x|": np.outer(F.PauliVec[0][0], F.PauliVec[0][0].conj()),
              "|+y><+y|": np.outer(F.PauliVec[1][0], F.PauliVec[1][0].conj())}
    p = O.ODEParams(None, 5.0, lambda tf, t: t / tf)
    for key in ("lindblad_rhs_Lz_on_plus_x", "lindblad_rhs_Lz_Lx_on_plus_y"):
        g = G[key]
        L = O.LindbladLiouvillian([O.Lindblad(gm, ops[o]) for gm, o in zip(g["gamma"], g["ops"])])
        assert np.allclose(L.rhs_acc(np.zeros((2, 2), complex), states[g["state"]], p, 5.0), _cm(g["drho"]), atol=1e-15)
    g = G["dense_hamiltonian_half"]
    H = O.DenseHamiltonian([A, B], [sx, sz])
    assert np.array_equal(H(g["s"]), _cm(g["H"])) and np.array_equal(H.update_cache(g["s"]), _cm(g["cache"]))
    for key in ("correlated_davies_zero", "correlated_davies_closed_form"):
        g = G[key]
        gfun = lambda w: 1.0 if w >= 0 else np.exp(-0.5)
        inds = [tuple(x) for x in g["inds"]]
        D = O.CorrelatedDaviesGenerator([ops[c] for c in g["couplings"]], {i: gfun for i in inds}, {i: (lambda w: 0.0) for i in inds}, inds)
        du = D.rhs_acc(np.zeros((2, 2), complex), _cm(g["rho"]), O.build_gap_indices(np.array(g["w"]), 8, 8), None, g["s"])
        assert np.allclose(du, _cm(g["du"]), atol=1e-15)
    g = G["redfield_lambda_code_form"]
    red = O.RedfieldLiouvillian([(((1, 1),), [sz], lambda t1, t2: 1.0)], lambda t: sla.expm(-1j * t * sx), 5.0, 1e-8, 1e-6)
    assert np.allclose(red.Lambda(p, 5.0, [sz], lambda t1, t2: 1.0, 1), _cm(g["Lambda"]), atol=g["tolerance"])
    g = G["odeparams"]
    assert O.ODEParams(None, g["tf"], lambda tf, t: t / tf)(g["t"]) == g["s"]
    g = G["ohmic_gamma0"]
    bath = O.Ohmic(g["eta"], g["fc"], g["T"])
    assert bath.spectrum(0.0) == 2 * np.pi * g["eta"] / bath.beta
    g = G["ohmic_lambshift_at_zero"]
    assert np.isclose(O.lambshift(g["w"], O.OhmicBath(g["eta"], g["wc"], g["beta"]).spectrum), g["S"], rtol=g["rtol"], atol=g["atol"])
    g = G["lambshift_piecewise_exponential_spectrum"]
    gfun = lambda w: np.exp(-w) if w >= 0 else np.exp(0.8 * w)
    assert abs(O.lambshift(g["w"], gfun) - g["S"]) < g["atol"]
    tb = g["table"]
    assert abs(O.build_lambshift(np.linspace(tb["lo"], tb["hi"], tb["length"]), True, gfun)(g["w"]) - g["S"]) < g["atol"]
    g = G["bspline_linear_data"]
    x = np.linspace(g["grid"]["lo"], g["grid"]["hi"], g["grid"]["length"])
    itp = O.QuadraticBSpline(x[0], x[1] - x[0], 10 - x)
    tx = np.linspace(g["test_grid"]["lo"], g["test_grid"]["hi"], g["test_grid"]["length"])
    assert np.allclose([itp(v) for v in tx], 10 - tx) and all(itp(xo) == yo for xo, yo in g["outside"])
    assert np.isclose(itp.gradient(2.3), g["gradient"])


def test_lambshift_known_answers():
    """test/coupling_bath_interaction/bath.jl:14-16 (S(0) of the Ohmic bath against the cpvagk value, rtol 1e-4) and
    :72-80 (correlated-bath spectrum: quadgk per call and the 1000-point spline both give 0.03551, atol 1e-4)."""
    bath = O.OhmicBath(1e-4, 8 * np.pi, 1 / 2.23)
    assert np.isclose(O.lambshift(0.0, bath.spectrum), -0.0025132734115775254, rtol=1e-4, atol=1e-4)
    gfun = lambda w: np.exp(-w) if w >= 0 else np.exp(0.8 * w)
    assert abs(O.lambshift(0.0, gfun) - 0.03551) < 1e-4
    assert O.build_lambshift([0.0], False, gfun)(0.5) == 0
    S_tab = O.build_lambshift(np.linspace(-5, 5, 1000), True, gfun)
    assert abs(S_tab(0.0) - 0.03551) < 1e-4


def test_quadratic_bspline_known_answers():
    """test/interpolations.jl:3-22: linear data on a 100-point range is reproduced on a 50-point range, the value
    outside the grid is the fill value 0, the gradient of 10 − x is −1."""
    x = np.linspace(0, 10, 100)
    itp = O.QuadraticBSpline(0.0, 10 / 99, 10 - x)
    itc = O.QuadraticBSpline(0.0, 10 / 99, x + 1j * x)
    tx = np.linspace(0, 10, 50)
    assert np.allclose([itp(v) for v in tx], 10 - tx) and np.allclose([itc(v) for v in tx], tx + 1j * tx)
    assert itc(-1) == 0 and isinstance(itc(-1), complex) or np.iscomplexobj(itc(-1))
    assert itp(11) == 0
    assert np.isclose(itp.gradient(2.3), -1) and np.allclose([itp.gradient(v) for v in (0.13, 0.21)], [-1, -1])


def test_ohmic_correlation_and_timescales_known_answers():
    """test/coupling_bath_interaction/bath.jl:10,18-26: correlation(t1, t2) = correlation(t1 − t2); for Ohmic(1e-4, 4, 16)
    τ_SB ≈ 284.61181493, τ_B(bath, 100, τ_SB) ≈ 0.07638653 (atol = rtol = 1e-6), T_c = √(τ_SB τ_B/5).  These pin the complex
    trigamma restatement; it is also checked against mpmath where that is importable."""
    bath = O.Ohmic(1e-4, 4, 16)
    tsb, _ = O.tau_SB(bath.correlation)
    tb, _ = O.tau_B(bath.correlation, 100, tsb)
    assert np.isclose(tsb, 284.61181493, rtol=1e-6, atol=1e-6) and np.isclose(tb, 0.07638653, rtol=1e-6, atol=1e-6)
    tc, _ = O.coarse_grain_timescale(bath.correlation, 100)
    assert np.isclose(tc, np.sqrt(tsb * tb / 5), rtol=1e-6, atol=1e-6)
    b2 = O.OhmicBath(1e-4, 8 * np.pi, 1 / 2.23)
    assert b2.correlation(0.02 - 0.01) == b2.correlation(0.01)
    mpmath = pytest.importorskip("mpmath")
    for z in (0.084 + 0.3j, 1.084 - 2j, 0.5, 3 + 40j, 0.084 + 1e-3j, 12.5 + 0.1j, -0.3 + 0.2j):
        ref = complex(mpmath.psi(1, z))
        assert abs(O.trigamma(z) - ref) <= 1e-15 * abs(ref) * 4


# ---- oracle/davies_groups.py: the scalable Davies statement against the literal loop --------------------------------
def _spectra(l, rng):
    """generic / equally spaced / degenerate / 8-digit coincidences forced / several of those mixed"""
    w_gen = np.sort(rng.standard_normal(l) * 3.0)
    w_eq = np.arange(l, dtype=float) * 0.5
    w_deg = np.sort(np.repeat(np.arange(l // 4 + 1, dtype=float), 4)[:l] * 1.25)
    w_coin = w_gen.copy()
    w_coin[l // 2] = w_coin[l // 2 - 1] + (w_coin[3] - w_coin[1])            # an exact coincidence of two gaps
    w_coin[l - 1] = w_coin[l - 2] + (w_coin[2] - w_coin[0]) * (1 + 3e-10)    # a coincidence after rounding to 8 digits
    w_coin = np.sort(w_coin)
    w_tiny = w_gen.copy()
    w_tiny[1] = w_tiny[0] + 3e-9                                              # inside the 10^-8 zero-gap threshold
    w_tiny[5] = w_tiny[4] + 1.2e-8                                            # just outside it
    return {"generic": w_gen, "equal": w_eq, "degenerate": w_deg, "coincidence": w_coin, "tiny": np.sort(w_tiny)}


@pytest.mark.parametrize("l", [6, 16, 40])
@pytest.mark.parametrize("kind", ["generic", "equal", "degenerate", "coincidence", "tiny"])
def test_davies_groups_oracle_matches_literal_loop(l, kind):
    from oracle import davies_groups as DG
    rng = np.random.default_rng(l * 31 + len(kind))
    w = _spectra(l, rng)[kind]
    As = [(lambda m: m + m.conj().T)(rng.standard_normal((l, l)) + 1j * rng.standard_normal((l, l))) for _ in range(2)]
    rho = rng.standard_normal((l, l)) + 1j * rng.standard_normal((l, l))  # no Hermiticity assumed (App. A.3)
    gi = O.build_gap_indices(w, 8, 8)
    D = O.DaviesGenerator(As, gamma, lamb)
    ref = D.rhs_acc(np.zeros((l, l), dtype=complex), rho, gi, None, 0.0)   # literal loop, couplings already in the eigenbasis
    got = DG.davies_eig_acc(np.zeros((l, l), dtype=complex), rho, As, w, gamma, lamb, 8, 8)
    assert np.linalg.norm(got - ref) <= 2e-14 * np.linalg.norm(ref)
    # the grouping itself: same keys, same member lists in loop order, same zero group
    grp, zero = DG.gap_groups(w, 8, 8)
    assert sorted(grp) == list(gi.uniq_w)
    for key, a, b in gi.positive_gap_indices():
        assert [(i + 1, j + 1) for i, j in grp[key]] == list(zip(a, b))
    assert [(i + 1, j + 1) for i, j in zero] == list(zip(gi.a0[0:2 * len(zero):2], gi.b0[0:2 * len(zero):2]))


def test_davies_groups_oracle_two_generators_full_rhs():
    """ame_rhs with two DaviesGenerators (davies_from_interactions: one per interaction, src/coupling/interaction.jl:101-117)
    against the literal DiffEqLiouvillian{true,false} restatement."""
    from oracle import davies_groups as DG
    rng = np.random.default_rng(5)
    n = 16
    H = F.random_hermitian(n, rng) if hasattr(F, "random_hermitian") else (lambda m: (m + m.conj().T) / 2)(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    w, v = np.linalg.eigh(H)
    cs = [(lambda m: m + m.conj().T)(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) for _ in range(3)]
    g2 = lambda x: 0.3 * gamma(x)
    s2 = lambda x: 0.5 * x
    u = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    got = DG.ame_rhs(u, w, v, cs, [(gamma, lamb, [0, 1]), (g2, s2, [2])], 8, 8)
    gi = O.build_gap_indices(w, 8, 8)
    rho = v.conj().T @ u @ v
    X = -1j * (np.diag(w) @ rho - rho @ np.diag(w))
    O.DaviesGenerator(cs[:2], gamma, lamb).rhs_acc(X, rho, gi, v, 0.0)
    O.DaviesGenerator(cs[2:], g2, s2).rhs_acc(X, rho, gi, v, 0.0)
    ref = v @ X @ v.conj().T
    assert np.linalg.norm(got - ref) <= 1e-13 * np.linalg.norm(ref)


# ---- test/utilities/multi_qubits_construction.jl:3-15 — the builders every fixture of the parity tests goes through ----
def test_fixture_builders_known_answers():
    k = F.kron
    assert np.array_equal(F.q_translate("ZZ+0.5ZI-XZ"), k(sz, sz) + 0.5 * k(sz, si) - k(sx, sz))           # :3
    assert np.array_equal(F.single_clause(["x"], [2], 0.5, 2), 0.5 * k(si, sx))                             # :4
    assert np.array_equal(F.single_clause(["z", "z"], [2, 3], -2, 4), -2 * k(si, sz, sz, si))               # :6
    assert np.array_equal(F.single_clause([F.sp], [3], 0.1, 3), 0.1 * k(si, si, F.sp))                      # :7
    assert np.array_equal(F.standard_driver(2), k(sx, si) + k(si, sx))                                      # :9
    assert np.allclose(F.collective_operator("z", 3), k(sz, si, si) + k(si, sz, si) + k(si, si, sz))        # :11
    assert np.allclose(F.local_field_term([-1.0, 0.5], [1, 3], 3), -1.0 * k(sz, si, si) + 0.5 * k(si, si, sz))   # :13
    assert np.allclose(F.two_local_term([-1.0, 0.5], [[1, 3], [1, 2]], 3), -1.0 * k(sz, si, sz) + 0.5 * k(sz, sz, si))  # :15
    # the diagonal short-cut used by the full-size configurations is the same operator
    J, idx = [-1.0, 0.5, 0.25], [[1, 3], [1, 2], [2, 4]]
    assert np.allclose(np.diag(F.ising_diag(4, J, idx)), F.two_local_term(J, idx, 4))
    # σ₊ / σ₋ of src/math_util.jl:7-8
    assert np.array_equal(F.sp, np.array([[0, 1], [0, 0]])) and np.array_equal(F.sm, F.sp.T)


# ---- test/opensys/stochastic.jl:3-17 -------------------------------------------------------------
def test_fluctuator_cache_known_answers():
    """FluctuatorLiouvillian with ConstantCouplings(["Z"], unit=:ħ) and two fluctuators of amplitude b = [1, 1]: the noise
    values are n = b .* (±1), so update_cache! gives −i Σ_k n_k σz and update_vectorized_cache! its superoperator form."""
    p = O.ODEParams(None, 5.0, lambda tf, t: t / tf)
    for signs in ([1, 1], [1, -1], [-1, -1]):
        b0 = np.array([1.0, 1.0]) * np.array(signs)
        fl = O.FluctuatorLiouvillian([sz], [b0.sum()])  # one coupling, n = sum(b0, dims=1) (src/opensys/stochastic.jl:55-57)
        cache_exp = -1j * sum(b * sz for b in b0)                                                   # :8
        cache = fl.update_cache_acc(np.zeros((2, 2), dtype=complex), p, 2.5)
        assert np.allclose(cache, cache_exp)                                                         # :13
        v = fl.update_vectorized_cache_acc(np.zeros((4, 4), dtype=complex), p, 2.5)
        one = np.eye(2)
        assert np.allclose(v, np.kron(one, cache_exp) - np.kron(cache_exp.T, one))                  # :17
