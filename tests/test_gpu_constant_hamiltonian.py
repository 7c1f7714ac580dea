"""ConstantDenseHamiltonian / ConstantSparseHamiltonian (src/hamiltonian/dense_hamiltonian.jl:134-168,
sparse_hamiltonian.jl:166-203, interface.jl:1-35) on the device: the known answers of
test/hamiltonian/constant_hamiltonian.jl, parity with the oracle at larger sizes, and the stencil route of a single
Pauli-structured matrix (a TFIM at fixed s stored as ONE matrix)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sps

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oqb_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-12
sx = np.array([[0, 1], [1, 0]], dtype=complex)
sz = np.array([[1, 0], [0, -1]], dtype=complex)


@pytest.fixture(scope="module")
def Q():
    import oqb_b200
    oqb_b200.get_context()
    return oqb_b200


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


def rand_c(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def tfim_matrix(nq, s):
    n = 1 << nq
    idx = np.arange(n)
    H = np.zeros((n, n), dtype=complex)
    for q in range(nq):
        H[idx ^ (1 << q), idx] += -(1 - s)
    z = [1.0 - 2.0 * ((idx >> (nq - 1 - q)) & 1) for q in range(nq)]
    H += np.diag(-s * sum(z[i] * z[i + 1] for i in range(nq - 1)))
    return H


def test_reference_known_answers(Q):
    """test/hamiltonian/constant_hamiltonian.jl:5-46 through the device entry points."""
    H1 = Q.ConstantHamiltonian(sx, unit="ħ")
    H2 = Q.ConstantHamiltonian(sx, static=False)
    H3 = Q.ConstantHamiltonian(sps.csc_matrix(np.kron(sx, sx)))
    Hs1 = Q.Hamiltonian(sx, unit="ħ")
    assert isinstance(H1, Q.ConstantDenseHamiltonian) and isinstance(H3, Q.ConstantSparseHamiltonian) and H1.isconstant and H3.isconstant
    assert np.array_equal(H1(2), sx) and np.array_equal(Hs1(2), sx)                              # :17
    assert np.array_equal(H2(1), 2 * np.pi * sx)                                                  # :18
    assert np.array_equal(H3(0.2).toarray(), 2 * np.pi * np.kron(sx, sx))                         # :19
    assert np.array_equal(H1.evaluate_device(2), sx)                                              # the same through oqb_dense_hamiltonian
    assert np.array_equal(H1.update_cache(np.zeros((2, 2), complex), 0.2), -1j * sx)              # :26
    assert np.array_equal(H2.update_cache(np.zeros((2, 2), complex), 0.1), -2 * np.pi * 1j * sx)  # :27
    assert np.array_equal(H3.update_cache(0.5).toarray(), -2 * np.pi * 1j * np.kron(sx, sx))      # :28
    v1 = H1.update_vectorized_cache(np.zeros((4, 4), complex), 0.2)
    assert np.array_equal(v1, -1j * O.vectorized_commutator(H1(0.1)))                             # :36
    v3 = H3.update_vectorized_cache(np.zeros((16, 16), complex), 0.5)
    assert np.allclose(v3, -1j * O.vectorized_commutator(H3(2).toarray()), atol=1e-14)            # :38
    r1 = np.ones((2, 2), dtype=complex) / 2
    r3 = np.ones((4, 4), dtype=complex) / 4
    assert np.allclose(H1(np.zeros_like(r1), r1, None, 0), -1j * (sx @ r1 - r1 @ sx), atol=1e-15)  # :44
    h2 = 2 * np.pi * sx
    assert np.allclose(H2(np.zeros_like(r1), r1, None, 0.5), -1j * (h2 @ r1 - r1 @ h2), atol=1e-15)  # :45
    h3 = 2 * np.pi * np.kron(sx, sx)
    assert np.allclose(H3(np.zeros_like(r3), r3, None, 1), -1j * (h3 @ r3 - r3 @ h3), atol=1e-15)    # :46
    w1, _ = Q.haml_eigs(H1, 0, 2)
    assert np.allclose(w1 / 2 / np.pi, np.array([-1, 1]) / 2 / np.pi)                             # :48-55 (eigen_decomp divides by 2π)


@pytest.mark.parametrize("n,nb", [(3, 4), (16, 5), (100, 3), (256, 4)])
def test_constant_dense_vs_oracle(Q, n, nb):
    rng = np.random.default_rng(n + nb)
    m = (lambda g: g + g.conj().T)(rand_c(rng, n, n))
    Hg, Ho = Q.ConstantDenseHamiltonian(m), O.ConstantDenseHamiltonian(m)
    u = rand_c(rng, nb, n, n)
    du = Hg(np.full_like(u, 7.0), u, None, rng.random(nb))  # the s argument is ignored
    assert relerr(du, np.stack([Ho.rhs(x) for x in u])) < TOL
    assert relerr(Hg.update_cache(np.zeros((1, n, n), complex), 0.3)[0], Ho.update_cache()) < 1e-15
    if n <= 16:
        v = Hg.update_vectorized_cache(np.zeros((1, n * n, n * n), complex), 0.3)[0]
        assert relerr(v, Ho.update_vectorized_cache()) < 1e-15


def test_constant_sparse_vs_oracle(Q):
    rng = np.random.default_rng(8)
    n = 64
    m = sps.random(n, n, density=0.1, random_state=3, format="csc") + 0j
    m = sps.csc_matrix(m + m.conj().T)
    Hg, Ho = Q.ConstantSparseHamiltonian(m), O.ConstantSparseHamiltonian(m)
    u = rand_c(rng, 3, n, n)
    du = Hg(np.zeros_like(u), u, None, 0.4)
    assert relerr(du, np.stack([Ho.rhs(x) for x in u])) < TOL
    assert relerr(Hg.update_cache(0.1).toarray(), Ho.update_cache().toarray()) < 1e-15


@pytest.mark.parametrize("nq,nb", [(4, 3), (6, 4), (8, 3)])
@pytest.mark.parametrize("with_lindblad", [False, True])
def test_constant_pauli_structured_matrix_takes_the_stencil_route(Q, nq, nb, with_lindblad):
    """A TFIM at fixed s stored as ONE matrix: X part on one XOR diagonal per qubit, ZZ part on XOR diagonal 0 — the
    commutator (and, with dephasing operators, the whole Lindblad RHS) runs as the stencil pass: no product.  Against the
    oracle and against the product routes (options lindblad_stencil = 0, and lindblad_split = 0 on top)."""
    n = 1 << nq
    rng = np.random.default_rng(nq + nb + with_lindblad)
    m = tfim_matrix(nq, 0.37)
    u = rand_c(rng, nb, n, n)
    ctx = Q.get_context()
    if with_lindblad:
        idx = np.arange(n)
        Ls = [np.diag(1.0 - 2.0 * ((idx >> (nq - 1 - q)) & 1)).astype(complex) for q in range(nq)]
        gam = [0.05 * (1 + k) for k in range(nq)]
        opg = Q.DiffEqLiouvillian(Q.ConstantDenseHamiltonian(m, unit="ħ"), [], [Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gam, Ls)])], n)
        opo = O.DiffEqLiouvillian(O.ConstantDenseHamiltonian(m, unit="ħ"), [], [O.LindbladLiouvillian([O.Lindblad(g, L) for g, L in zip(gam, Ls)])], n)
        run = lambda: opg(np.zeros_like(u), u, Q.ODEParams(None, 10.0), 4.2)
        ref = np.stack([opo.rhs(x, O.ODEParams(None, 10.0), 4.2) for x in u])
    else:
        Hg, Ho = Q.ConstantDenseHamiltonian(m, unit="ħ"), O.ConstantDenseHamiltonian(m, unit="ħ")
        run = lambda: Hg(np.full_like(u, 7.0), u, None, 0.0)
        ref = np.stack([Ho.rhs(x) for x in u])
    run()
    ctx.check(ctx.lib.oqb_profile_begin(ctx.h))
    got = run()
    ms, nl, fl = C.c_double(), C.c_uint64(), C.c_double()
    ctx.check(ctx.lib.oqb_profile_end(ctx.h, C.byref(ms), C.byref(nl), C.byref(fl)))
    assert relerr(got, ref) < TOL
    assert fl.value == 0.0  # no product was launched
    with ctx.option("lindblad_stencil", 0):
        plain = run()
        with ctx.option("lindblad_split", 0):
            plain2 = run()
    assert relerr(plain, ref) < TOL and relerr(plain2, ref) < TOL and relerr(got, plain) < 1e-12
