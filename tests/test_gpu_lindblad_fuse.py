"""Diagonal Lindblad operators with rates shared by the batch (config C2: per-qubit dephasing, per-member s): the sandwich
Σ_k γ_k L_k ρ L_k' = G∘ρ uses ONE n×n factor G for the whole batch (built once per call) and a pure streaming pass, instead of a
K-term sum per element and member.  Against the oracle's literal `γ(LρL' − ½{L'L, ρ})` loop (src/opensys/lindblad.jl:94-101) at 1e-12, and against the
unfused route (option lindblad_fuse = 0)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oqb_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def Q():
    import oqb_b200
    return oqb_b200


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


def rand_c(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


@pytest.mark.parametrize("n,K,nb", [(8, 3, 5), (50, 4, 7), (64, 8, 6), (100, 2, 3), (256, 8, 9)])
@pytest.mark.parametrize("real_h", [True, False])
@pytest.mark.parametrize("per_member_s", [True, False])
def test_fused_sandwich_vs_oracle(Q, n, K, nb, real_h, per_member_s):
    rng = np.random.default_rng(n * 13 + K + nb + 2 * real_h + per_member_s)
    mk = (lambda: (lambda g: g + g.T)(rng.standard_normal((n, n))).astype(complex)) if real_h else (lambda: (lambda g: g + g.conj().T)(rand_c(rng, n, n)))
    m1, m2 = mk(), mk()
    fs = [lambda s: -(1 - s), lambda s: s]
    Ls = [np.diag(np.exp(1j * rng.uniform(0, 2 * np.pi, n)) * rng.uniform(0.5, 1.5, n)) for _ in range(K)]  # complex diagonals
    gam = [0.1 * (1 + k / 8) for k in range(K)]  # constant rates: shared by the batch whatever s is
    Hg, Ho = Q.DenseHamiltonian(fs, [m1, m2], unit="ħ"), O.DenseHamiltonian(fs, [m1, m2], unit="ħ")
    opg = Q.DiffEqLiouvillian(Hg, [], [Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gam, Ls)])], n)
    opo = O.DiffEqLiouvillian(Ho, [], [O.LindbladLiouvillian([O.Lindblad(g, L) for g, L in zip(gam, Ls)])], n)
    u = rand_c(rng, nb, n, n)  # NOT Hermitian: the reference assumes nothing
    tf = 10.0
    t = rng.random(nb) * tf if per_member_s else 3.7
    tl = t if per_member_s else [t] * nb
    ctx = Q.get_context()
    ref = np.stack([opo.rhs(u[b], O.ODEParams(None, tf), tl[b]) for b in range(nb)])
    fused = opg(np.zeros_like(u), u, Q.ODEParams(None, tf), t)
    with ctx.option("lindblad_fuse", 0):
        plain = opg(np.zeros_like(u), u, Q.ODEParams(None, tf), t)
    assert relerr(fused, ref) < TOL and relerr(plain, ref) < TOL
    assert relerr(fused, plain) < 1e-13


def test_per_member_rates_keep_the_separate_pass(Q):
    """γ_k(s) with per-member s: G differs per member, the Hadamard pass stays (and stays correct)."""
    n, K, nb = 64, 3, 6
    rng = np.random.default_rng(9)
    m1, m2 = [(lambda g: g + g.T)(rng.standard_normal((n, n))).astype(complex) for _ in range(2)]
    fs = [lambda s: -(1 - s), lambda s: s]
    Ls = [np.diag(rng.choice([-1.0, 1.0], n)).astype(complex) for _ in range(K)]
    gfs = [lambda s, k=k: 0.1 * (1 + k) * (1 + s) for k in range(K)]
    opg = Q.DiffEqLiouvillian(Q.DenseHamiltonian(fs, [m1, m2], unit="ħ"), [], [Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gfs, Ls)])], n)
    opo = O.DiffEqLiouvillian(O.DenseHamiltonian(fs, [m1, m2], unit="ħ"), [], [O.LindbladLiouvillian([O.Lindblad(g, L) for g, L in zip(gfs, Ls)])], n)
    u = rand_c(rng, nb, n, n)
    t = rng.random(nb) * 10.0
    got = opg(np.zeros_like(u), u, Q.ODEParams(None, 10.0), t)
    ref = np.stack([opo.rhs(u[b], O.ODEParams(None, 10.0), t[b]) for b in range(nb)])
    assert relerr(got, ref) < TOL


def test_shared_factor_with_the_fallback_product(Q):
    """The same route with the cp.async fallback product (option zgemm_tma = 0)."""
    n, K, nb = 40, 2, 4
    rng = np.random.default_rng(10)
    m1, m2 = [(lambda g: g + g.conj().T)(rand_c(rng, n, n)) for _ in range(2)]
    fs = [lambda s: -(1 - s), lambda s: s]
    Ls = [np.diag(rand_c(rng, n)) for _ in range(K)]
    opg = Q.DiffEqLiouvillian(Q.DenseHamiltonian(fs, [m1, m2], unit="ħ"), [], [Q.LindbladLiouvillian([Q.Lindblad(0.3, L) for L in Ls])], n)
    opo = O.DiffEqLiouvillian(O.DenseHamiltonian(fs, [m1, m2], unit="ħ"), [], [O.LindbladLiouvillian([O.Lindblad(0.3, L) for L in Ls])], n)
    u = rand_c(rng, nb, n, n)
    ctx = Q.get_context()
    with ctx.option("zgemm_tma", 0):
        got = opg(np.zeros_like(u), u, Q.ODEParams(None, 10.0), 4.0)
    ref = np.stack([opo.rhs(u[b], O.ODEParams(None, 10.0), 4.0) for b in range(nb)])
    assert relerr(got, ref) < TOL


@pytest.mark.parametrize("nq,nb", [(3, 5), (6, 4), (8, 6)])
@pytest.mark.parametrize("per_member_s", [True, False])
@pytest.mark.parametrize("complex_driver", [False, True])
def test_split_hamiltonian_vs_oracle(Q, nq, nb, per_member_s, complex_driver):
    """H(s) = A(s)·H_driver + B(s)·H_problem + C(s)·H_bias with ONE off-diagonal term (the driver) and diagonal problem terms,
    dephasing operators, constant rates — config C2's structure.  The library never assembles H(s_b): both commutator products
    use the shared driver matrix scaled by A(s_b) and the diagonal terms join the Hadamard pass (option lindblad_split = 0: the
    assembled route).  Oracle: the literal commutator + Lindblad loop."""
    from oracle import fixtures as F
    n = 2 ** nq
    rng = np.random.default_rng(nq * 5 + nb + per_member_s)
    Hd = -F.standard_driver(nq).astype(complex)
    if complex_driver:  # a Hermitian driver with imaginary entries (σy components): general complex product
        Y = (lambda g: 1j * (g - g.T))(rng.standard_normal((n, n)))
        Hd = Hd + 0.3 * Y
    Hp = -F.two_local_term([1.0] * (nq - 1), [[i, i + 1] for i in range(1, nq)], nq).astype(complex)
    Hb = np.diag(rng.standard_normal(n)).astype(complex)
    fs = [lambda s: 1 - s, lambda s: s, lambda s: 0.3 * s * (1 - s)]
    Ls = [F.single_clause(["z"], [k + 1], 1.0, nq).astype(complex) for k in range(nq)]
    gam = [0.05 * (1 + k) for k in range(nq)]
    opg = Q.DiffEqLiouvillian(Q.DenseHamiltonian(fs, [Hd, Hp, Hb], unit="ħ"), [], [Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gam, Ls)])], n)
    opo = O.DiffEqLiouvillian(O.DenseHamiltonian(fs, [Hd, Hp, Hb], unit="ħ"), [], [O.LindbladLiouvillian([O.Lindblad(g, L) for g, L in zip(gam, Ls)])], n)
    u = rand_c(rng, nb, n, n)
    tf = 10.0
    t = rng.random(nb) * tf if per_member_s else 6.1
    tl = t if per_member_s else [t] * nb
    ref = np.stack([opo.rhs(u[b], O.ODEParams(None, tf), tl[b]) for b in range(nb)])
    ctx = Q.get_context()
    got = opg(np.zeros_like(u), u, Q.ODEParams(None, tf), t)
    with ctx.option("lindblad_split", 0):
        plain = opg(np.zeros_like(u), u, Q.ODEParams(None, tf), t)
    assert relerr(got, ref) < TOL and relerr(plain, ref) < TOL
    assert relerr(got, plain) < 1e-13
    # host-buffer entry point takes the same route
    uh = np.ascontiguousarray(u.transpose(0, 2, 1))
    duh = opg.rhs_host(np.zeros_like(uh), uh, Q.ODEParams(None, tf), t) if hasattr(opg, "rhs_host") else None
    if duh is not None:
        assert relerr(duh.transpose(0, 2, 1), ref) < TOL


@pytest.mark.parametrize("K,rates", [(10, "const"), (10, "per_member"), (14, "per_member"), (14, "const")])
def test_ten_qubit_dephasing_takes_the_diagonal_route(Q, K, rates):
    """10 qubits, per-qubit dephasing (north_star's batched 10-qubit Lindblad RHS): the K diagonals (K·n·16 B = 164 KB at
    K = 10) must not push the RHS onto the 2+2K-product route.  Shared rates need no shared memory at all; per-member rates
    keep the diagonals in shared memory up to 208 KB and take the general products beyond.  All against the oracle."""
    nq, n, nb = 10, 1024, 2
    rng = np.random.default_rng(K + len(rates))
    idx = np.arange(n)
    Hd = np.zeros((n, n), dtype=complex)
    for q in range(nq):
        Hd[idx ^ (1 << q), idx] = -1.0
    z = [1.0 - 2.0 * ((idx >> (nq - 1 - q)) & 1) for q in range(nq)]
    Hp = np.diag(-sum(z[i] * z[i + 1] for i in range(nq - 1))).astype(complex)
    Ls = [np.diag(z[k % nq] * (1.0 + 0.1 * (k // nq))).astype(complex) for k in range(K)]
    fs = [lambda s: 1 - s, lambda s: s]
    gfs = [(lambda s, k=k: 0.05 * (1 + k)) if rates == "const" else (lambda s, k=k: 0.05 * (1 + k) * (1 + s)) for k in range(K)]
    opg = Q.DiffEqLiouvillian(Q.DenseHamiltonian(fs, [Hd, Hp], unit="ħ"), [], [Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gfs, Ls)])], n)
    opo = O.DiffEqLiouvillian(O.DenseHamiltonian(fs, [Hd, Hp], unit="ħ"), [], [O.LindbladLiouvillian([O.Lindblad(g, L) for g, L in zip(gfs, Ls)])], n)
    u = rand_c(rng, nb, n, n) / n
    t = np.array([2.0, 7.5])
    ctx = Q.get_context()
    opg(np.zeros_like(u), u, Q.ODEParams(None, 10.0), t)  # first call builds the device operator (its L_k'L_k products are not the RHS)
    ctx.check(ctx.lib.oqb_profile_begin(ctx.h))
    got = opg(np.zeros_like(u), u, Q.ODEParams(None, 10.0), t)
    import ctypes as C
    ms, nl, fl = C.c_double(), C.c_uint64(), C.c_double()
    ctx.check(ctx.lib.oqb_profile_end(ctx.h, C.byref(ms), C.byref(nl), C.byref(fl)))
    ref = np.stack([opo.rhs(u[b], O.ODEParams(None, 10.0), t[b]) for b in range(nb)])
    assert relerr(got, ref) < TOL
    gemm_flops_per_rho = fl.value / nb / (8.0 * n ** 3)  # products per ρ as the profile hooks counted them
    if rates == "const" or K <= 12:
        assert gemm_flops_per_rho <= 2.01, gemm_flops_per_rho     # 2 products per ρ
    else:
        assert gemm_flops_per_rho >= 2 * K, gemm_flops_per_rho    # the general route: 2 + 2K products per ρ


def _sigma_minus(nq, q):
    """σ⁻ on qubit q (1-based, qubit 1 = most significant bit) as a dense 2^nq matrix: |0⟩⟨1| on that qubit."""
    n = 1 << nq
    idx = np.arange(n)
    bit = 1 << (nq - q)
    L = np.zeros((n, n), dtype=complex)
    src = idx[(idx & bit) != 0]
    L[src ^ bit, src] = 1.0
    return L


@pytest.mark.parametrize("kind", ["amplitude_damping", "pauli_channel", "random_partial_permutation", "mixed_with_dephasing"])
@pytest.mark.parametrize("nq,nb", [(3, 5), (6, 4), (8, 3)])
@pytest.mark.parametrize("real_h", [True, False])
def test_monomial_operators_take_the_gather_route(Q, kind, nq, nb, real_h):
    """L_k with at most one entry per row and per column (σ⁻: amplitude damping; Pauli X/Y strings; any partial permutation
    with phases): L_k ρ L_k' is a gather of ρ and L_k'L_k is diagonal — 2 products per ρ instead of 2 + 2K.  Against the
    oracle's literal loop (src/opensys/lindblad.jl:94-101) and against the general route (option lindblad_mono = 0)."""
    import ctypes as C
    n = 1 << nq
    rng = np.random.default_rng(nq * 11 + nb + len(kind) + real_h)
    if kind == "amplitude_damping":
        Ls = [_sigma_minus(nq, q) for q in range(1, nq + 1)]
    elif kind == "pauli_channel":
        X = np.array([[0, 1], [1, 0]], dtype=complex); Y = np.array([[0, -1j], [1j, 0]]); I2 = np.eye(2, dtype=complex)
        def on(P, q):
            m = np.eye(1, dtype=complex)
            for j in range(1, nq + 1):
                m = np.kron(m, P if j == q else I2)
            return m
        Ls = [on(X, 1), on(Y, nq), on(X, 1) @ on(Y, 2)]
    elif kind == "random_partial_permutation":
        Ls = []
        for _ in range(3):
            L = np.zeros((n, n), dtype=complex)
            rows = rng.permutation(n)[: n - n // 4]
            cols = rng.permutation(n)[: n - n // 4]
            L[rows, cols] = rand_c(rng, n - n // 4)
            Ls.append(L)
    else:
        idx = np.arange(n)
        Ls = [_sigma_minus(nq, 1), np.diag(1.0 - 2.0 * ((idx >> (nq - 2)) & 1)).astype(complex), _sigma_minus(nq, nq).conj().T]
    K = len(Ls)
    herm = (lambda g: (g + g.T).astype(complex)) if real_h else (lambda g: g + g.conj().T)
    m1 = herm(rng.standard_normal((n, n)) if real_h else rand_c(rng, n, n))
    m2 = np.diag(rng.standard_normal(n)).astype(complex)
    fs = [lambda s: 1 - s, lambda s: s]
    gfs = [(lambda s, k=k: 0.1 * (1 + k) * (1 + 0.5 * s)) for k in range(K)]
    opg = Q.DiffEqLiouvillian(Q.DenseHamiltonian(fs, [m1, m2], unit="ħ"), [], [Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gfs, Ls)])], n)
    opo = O.DiffEqLiouvillian(O.DenseHamiltonian(fs, [m1, m2], unit="ħ"), [], [O.LindbladLiouvillian([O.Lindblad(g, L) for g, L in zip(gfs, Ls)])], n)
    u = rand_c(rng, nb, n, n)
    t = rng.random(nb) * 10.0
    p = Q.ODEParams(None, 10.0)
    ref = np.stack([opo.rhs(u[b], O.ODEParams(None, 10.0), t[b]) for b in range(nb)])
    ctx = Q.get_context()
    opg(np.zeros_like(u), u, p, t)  # builds the device operator
    ctx.check(ctx.lib.oqb_profile_begin(ctx.h))
    got = opg(np.zeros_like(u), u, p, t)
    ms, nl, fl = C.c_double(), C.c_uint64(), C.c_double()
    ctx.check(ctx.lib.oqb_profile_end(ctx.h, C.byref(ms), C.byref(nl), C.byref(fl)))
    assert relerr(got, ref) < TOL
    assert fl.value / nb / (8.0 * n ** 3) <= 2.01  # two products per ρ
    with ctx.option("lindblad_mono", 0):
        plain = opg(np.zeros_like(u), u, p, t)
    assert relerr(plain, ref) < TOL and relerr(got, plain) < 1e-12
    # the accumulate-only Lindblad form and shared s
    acc0 = rand_c(rng, nb, n, n)
    lg, lo = opg._lind[0], opo.opensys[0] if hasattr(opo, "opensys") else None
    d2 = lg(acc0.copy(), u, p, 2.5)
    if lo is not None:
        ref2 = np.stack([lo.rhs_acc(acc0[b].copy(), u[b], O.ODEParams(None, 10.0), 2.5) for b in range(nb)])
        assert relerr(d2, ref2) < TOL


@pytest.mark.parametrize("kind", ["sigma_x", "sigma_minus", "pauli_y_strings", "random_partial_permutation", "mixed_with_diagonal"])
@pytest.mark.parametrize("nq,nb", [(2, 3), (5, 4), (7, 3)])
def test_redfield_apply_with_monomial_couplings(Q, kind, nq, nb):
    """Redfield apply (src/opensys/redfield.jl:57-64: K₂ = Σ_α S_αΛ_αu − Λ_αuS_α, du −= K₂ + K₂') with couplings that have at
    most one entry per row and column — σx (the usual transverse coupling), σ⁻, Pauli-Y strings, partial permutations with
    phases: S M and M S are gathers, K products per member instead of 3K.  Against the oracle and the general route."""
    n = 1 << nq
    rng = np.random.default_rng(nq * 17 + nb + len(kind))
    X = np.array([[0, 1], [1, 0]], dtype=complex); Y = np.array([[0, -1j], [1j, 0]]); I2 = np.eye(2, dtype=complex)

    def on(P, q):
        m = np.eye(1, dtype=complex)
        for j in range(1, nq + 1):
            m = np.kron(m, P if j == q else I2)
        return m
    if kind == "sigma_x":
        S = [on(X, q) for q in range(1, nq + 1)]
    elif kind == "sigma_minus":
        S = [_sigma_minus(nq, q) for q in range(1, nq + 1)]
    elif kind == "pauli_y_strings":
        S = [on(Y, 1), on(Y, nq) @ on(X, 1)] if nq > 1 else [on(Y, 1)]
    elif kind == "random_partial_permutation":
        S = []
        for _ in range(3):
            L = np.zeros((n, n), dtype=complex)
            keep = max(1, n - n // 4)
            L[rng.permutation(n)[:keep], rng.permutation(n)[:keep]] = rand_c(rng, keep)
            S.append(L)
    else:
        S = [on(X, 1), np.diag(rand_c(rng, n)), _sigma_minus(nq, nq)]
    K = len(S)
    Lam = rand_c(rng, nb, K, n, n) / n
    u = rand_c(rng, nb, n, n)
    du0 = rand_c(rng, nb, n, n)
    rg = Q.RedfieldLiouvillian(None, None, 0, 0, 0)
    ctx = Q.get_context()
    ref = np.stack([O.redfield_apply_acc(du0[b].copy(), u[b], S, list(Lam[b])) for b in range(nb)])
    import ctypes as C
    ctx.check(ctx.lib.oqb_profile_begin(ctx.h))
    got = rg.apply(du0.copy(), u, S, Lam)
    ms, nl, fl = C.c_double(), C.c_uint64(), C.c_double()
    ctx.check(ctx.lib.oqb_profile_end(ctx.h, C.byref(ms), C.byref(nl), C.byref(fl)))
    assert relerr(got, ref) < TOL
    assert fl.value / nb / (8.0 * n ** 3) <= K + 0.01  # K products per member
    with ctx.option("coupling_mono", 0):
        plain = rg.apply(du0.copy(), u, S, Lam)
    assert relerr(plain, ref) < TOL and relerr(got, plain) < 1e-12


def _pauli_sum_dense(nq, strings):
    P = {"X": np.array([[0, 1], [1, 0]], dtype=complex), "Y": np.array([[0, -1j], [1j, 0]]), "Z": np.array([[1, 0], [0, -1]], dtype=complex),
         "I": np.eye(2, dtype=complex)}
    n = 1 << nq
    H = np.zeros((n, n), dtype=complex)
    for cf, ops in strings:
        m = np.eye(1, dtype=complex)
        for q in range(1, nq + 1):
            m = np.kron(m, P[ops.get(q, "I")])
        H = H + cf * m
    return H


@pytest.mark.parametrize("driver", ["transverse_field", "xx_catalyst", "y_field", "xz_products", "x_plus_z"])
@pytest.mark.parametrize("nq,nb", [(2, 3), (4, 5), (6, 4), (7, 3), (8, 5), (9, 2)])
@pytest.mark.parametrize("per_member_s", [True, False])
@pytest.mark.parametrize("mode", [2, 1])
def test_pauli_structured_driver_takes_the_stencil_route(Q, driver, nq, nb, per_member_s, mode):
    """DenseHamiltonian whose one off-diagonal term is a sum of Pauli strings (−ΣX; + XX catalysts: two-bit masks; Y fields:
    imaginary, row-signed entries; X·Z products: real, row-signed) + diagonal problem terms + dephasing: −i[H,ρ] is a
    stencil over ≤ 16 XOR diagonals fused with the Hadamard factor — no product launched at all.  Against the oracle's
    literal commutator + Lindblad loop and against the shared-operand product route (option lindblad_stencil = 0)."""
    import ctypes as C
    from oracle import fixtures as F
    n = 1 << nq
    rng = np.random.default_rng(nq * 19 + nb + len(driver) + per_member_s)
    if driver == "transverse_field":
        Hd = _pauli_sum_dense(nq, [(-1.0 - 0.1 * q, {q: "X"}) for q in range(1, nq + 1)])
    elif driver == "xx_catalyst":
        Hd = _pauli_sum_dense(nq, [(-1.0, {q: "X"}) for q in range(1, nq + 1)] + [(0.4, {q: "X", q + 1: "X"}) for q in range(1, nq)])
    elif driver == "y_field":
        Hd = _pauli_sum_dense(nq, [(0.7 + 0.1 * q, {q: "Y"}) for q in range(1, nq + 1)])
    elif driver == "x_plus_z":  # the off-diagonal term carries a diagonal part of its own (XOR mask 0)
        Hd = _pauli_sum_dense(nq, [(-1.0, {q: "X"}) for q in range(1, nq + 1)] + [(0.3 + 0.05 * q, {q: "Z"}) for q in range(1, nq + 1)])
    else:
        Hd = _pauli_sum_dense(nq, [(-1.0, {q: "X"}) for q in range(1, nq + 1)] + [(0.3, {q: "X", (q % nq) + 1: "Z"}) for q in range(1, nq + 1) if nq > 1])
    Hp = np.diag(rng.standard_normal(n)).astype(complex)
    Hb = np.diag(rng.standard_normal(n)).astype(complex)
    fs = [lambda s: 1 - s, lambda s: s, lambda s: 0.5 * s * s]
    idx = np.arange(n)
    Ls = [np.diag(1.0 - 2.0 * ((idx >> (nq - 1 - q)) & 1)).astype(complex) for q in range(nq)]
    gam = [0.05 * (1 + k) for k in range(nq)]
    opg = Q.DiffEqLiouvillian(Q.DenseHamiltonian(fs, [Hd, Hp, Hb], unit="ħ"), [], [Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gam, Ls)])], n)
    opo = O.DiffEqLiouvillian(O.DenseHamiltonian(fs, [Hd, Hp, Hb], unit="ħ"), [], [O.LindbladLiouvillian([O.Lindblad(g, L) for g, L in zip(gam, Ls)])], n)
    u = rand_c(rng, nb, n, n)
    tf = 10.0
    t = rng.random(nb) * tf if per_member_s else 4.2
    tl = t if per_member_s else [t] * nb
    ref = np.stack([opo.rhs(u[b], O.ODEParams(None, tf), tl[b]) for b in range(nb)])
    ctx = Q.get_context()
    with ctx.option("lindblad_stencil", mode):  # 2: register micro-tiles (n ≥ 64); 1: the plain kernel
        opg(np.zeros_like(u), u, Q.ODEParams(None, tf), t)
        ctx.check(ctx.lib.oqb_profile_begin(ctx.h))
        got = opg(np.zeros_like(u), u, Q.ODEParams(None, tf), t)
        ms, nl, fl = C.c_double(), C.c_uint64(), C.c_double()
        ctx.check(ctx.lib.oqb_profile_end(ctx.h, C.byref(ms), C.byref(nl), C.byref(fl)))
    assert relerr(got, ref) < TOL
    if not (driver == "xx_catalyst" and nq > 8):  # 9 single + 8 pair masks exceed the 16 XOR diagonals of the stencil: product route
        assert fl.value == 0.0  # no product was launched
    with ctx.option("lindblad_stencil", 0):
        plain = opg(np.zeros_like(u), u, Q.ODEParams(None, tf), t)
    assert relerr(plain, ref) < TOL and relerr(got, plain) < 1e-12


@pytest.mark.parametrize("driver", ["transverse_field", "xx_catalyst", "y_field", "x_plus_z"])
@pytest.mark.parametrize("nq,nb", [(3, 4), (6, 5), (7, 3), (8, 4), (10, 2)])
@pytest.mark.parametrize("unit", ["h", "ħ"])
def test_plain_commutator_of_a_pauli_structured_hamiltonian_is_a_stencil(Q, driver, nq, nb, unit):
    """(h::DenseHamiltonian)(du, u, p, s) — src/hamiltonian/dense_hamiltonian.jl:84-89 — with no Lindblad term at all: for a
    Pauli-structured driver + diagonal problem terms the commutator is the same stencil pass without the Hadamard factor
    (no product launched; OVERWRITES du).  Against the oracle's −i(Hu − uH) and against the two-product route."""
    import ctypes as C
    n = 1 << nq
    rng = np.random.default_rng(nq * 23 + nb + len(driver))
    if driver == "transverse_field":
        Hd = _pauli_sum_dense(nq, [(-1.0 - 0.1 * q, {q: "X"}) for q in range(1, nq + 1)])
    elif driver == "xx_catalyst":
        Hd = _pauli_sum_dense(nq, [(-1.0, {q: "X"}) for q in range(1, min(nq, 8) + 1)] + [(0.4, {q: "X", q + 1: "X"}) for q in range(1, min(nq, 8))])
    elif driver == "y_field":
        Hd = _pauli_sum_dense(nq, [(0.7 + 0.1 * q, {q: "Y"}) for q in range(1, nq + 1)])
    else:
        Hd = _pauli_sum_dense(nq, [(-1.0, {q: "X"}) for q in range(1, nq + 1)] + [(0.3 + 0.05 * q, {q: "Z"}) for q in range(1, nq + 1)])
    Hp = np.diag(rng.standard_normal(n)).astype(complex)
    fs = [lambda s: 1 - s, lambda s: s * s]
    Hg = Q.DenseHamiltonian(fs, [Hd, Hp], unit=unit)
    Ho = O.DenseHamiltonian(fs, [Hd, Hp], unit=unit)
    u = rand_c(rng, nb, n, n)
    s = rng.random(nb)
    ref = np.stack([Ho.rhs(u[b], s[b]) for b in range(nb)])
    ctx = Q.get_context()
    ctx.check(ctx.lib.oqb_profile_begin(ctx.h))
    got = Hg(np.full_like(u, 7.0), u, None, s)
    ms, nl, fl = C.c_double(), C.c_uint64(), C.c_double()
    ctx.check(ctx.lib.oqb_profile_end(ctx.h, C.byref(ms), C.byref(nl), C.byref(fl)))
    assert relerr(got, ref) < TOL
    assert fl.value == 0.0  # no product was launched
    shared = Hg(np.full_like(u, 7.0), u, None, 0.3)  # one s for the whole batch
    assert relerr(shared, np.stack([Ho.rhs(u[b], 0.3) for b in range(nb)])) < TOL
    with ctx.option("lindblad_stencil", 0):
        plain = Hg(np.full_like(u, 7.0), u, None, s)
    assert relerr(plain, ref) < TOL and relerr(got, plain) < 1e-12


@pytest.mark.parametrize("kind", ["amplitude_damping", "pauli_channel", "dephasing_per_member_rates", "mixed_with_dephasing"])
@pytest.mark.parametrize("nq,nb", [(3, 5), (6, 4), (8, 3)])
@pytest.mark.parametrize("driver", ["transverse_field", "y_field"])
def test_pauli_structured_hamiltonian_with_monomial_or_rate_dependent_operators(Q, kind, nq, nb, driver):
    """Pauli-structured H(s) with L_k that are monomial (σ⁻, Pauli strings) or diagonal with rates that depend on s_b: the
    commutator is the stencil pass and the dissipator (anticommutator included) the gather / Hadamard pass — no product.
    Against the oracle's literal loop (src/opensys/lindblad.jl:94-101) and the product route (option lindblad_stencil = 0)."""
    import ctypes as C
    n = 1 << nq
    rng = np.random.default_rng(nq * 29 + nb + len(kind) + len(driver))
    idx = np.arange(n)
    zq = lambda q: np.diag(1.0 - 2.0 * ((idx >> (nq - q)) & 1)).astype(complex)
    if kind == "amplitude_damping":
        Ls = [_sigma_minus(nq, q) for q in range(1, nq + 1)]
    elif kind == "pauli_channel":
        Ls = [_pauli_sum_dense(nq, [(1.0, {1: "X"})]), _pauli_sum_dense(nq, [(1.0, {nq: "Y"})]), _pauli_sum_dense(nq, [(1.0, {1: "X", 2: "Y"})])]
    elif kind == "dephasing_per_member_rates":
        Ls = [zq(q) for q in range(1, nq + 1)]
    else:
        Ls = [_sigma_minus(nq, 1), zq(2), _sigma_minus(nq, nq).conj().T]
    K = len(Ls)
    if driver == "transverse_field":
        Hd = _pauli_sum_dense(nq, [(-1.0 - 0.1 * q, {q: "X"}) for q in range(1, nq + 1)])
    else:
        Hd = _pauli_sum_dense(nq, [(0.7 + 0.1 * q, {q: "Y"}) for q in range(1, nq + 1)])
    Hp = np.diag(rng.standard_normal(n)).astype(complex)
    fs = [lambda s: 1 - s, lambda s: s]
    gfs = [(lambda s, k=k: 0.1 * (1 + k) * (1 + 0.5 * s)) for k in range(K)]
    opg = Q.DiffEqLiouvillian(Q.DenseHamiltonian(fs, [Hd, Hp], unit="ħ"), [], [Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gfs, Ls)])], n)
    opo = O.DiffEqLiouvillian(O.DenseHamiltonian(fs, [Hd, Hp], unit="ħ"), [], [O.LindbladLiouvillian([O.Lindblad(g, L) for g, L in zip(gfs, Ls)])], n)
    u = rand_c(rng, nb, n, n)
    t = rng.random(nb) * 10.0
    p = Q.ODEParams(None, 10.0)
    ref = np.stack([opo.rhs(u[b], O.ODEParams(None, 10.0), t[b]) for b in range(nb)])
    ctx = Q.get_context()
    opg(np.zeros_like(u), u, p, t)  # builds the device operator
    ctx.check(ctx.lib.oqb_profile_begin(ctx.h))
    got = opg(np.zeros_like(u), u, p, t)
    ms, nl, fl = C.c_double(), C.c_uint64(), C.c_double()
    ctx.check(ctx.lib.oqb_profile_end(ctx.h, C.byref(ms), C.byref(nl), C.byref(fl)))
    assert relerr(got, ref) < TOL
    assert fl.value == 0.0  # no product was launched
    shared = opg(np.zeros_like(u), u, p, 4.2)  # one time for the whole batch
    assert relerr(shared, np.stack([opo.rhs(u[b], O.ODEParams(None, 10.0), 4.2) for b in range(nb)])) < TOL
    with ctx.option("lindblad_stencil", 0):
        plain = opg(np.zeros_like(u), u, p, t)
    assert relerr(plain, ref) < TOL and relerr(got, plain) < 1e-12
