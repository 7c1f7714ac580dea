"""The host-only structure analysis of oqb_dense_create (csrc/dense.cu analyze_terms), run without a GPU through
oqb_dense_structure_probe: which Hamiltonian term carries the off-diagonal entries, the XOR diagonals of a Pauli-structured
term, and — with the tables the probe returns — the stencil form of the commutator the kernels evaluate,
    du[r,c] = −i f_N Σ_j (x_j[r] u[r^m_j, c] − u[r, c^m_j] x_j[c^m_j]) − i Σ_i f_i (h_i[r] − h_i[c]) u[r,c],
restated in numpy and checked against the oracle's −i(Hu − uH) (src/hamiltonian/dense_hamiltonian.jl:84-89)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import oqb_oracle as O  # noqa: E402

PAULI = {"X": np.array([[0, 1], [1, 0]], dtype=complex), "Y": np.array([[0, -1j], [1j, 0]]), "Z": np.array([[1, 0], [0, -1]], dtype=complex),
         "I": np.eye(2, dtype=complex)}


def pauli_sum(nq, strings):
    n = 1 << nq
    H = np.zeros((n, n), dtype=complex)
    for cf, ops in strings:
        m = np.eye(1, dtype=complex)
        for q in range(1, nq + 1):
            m = np.kron(m, PAULI[ops.get(q, "I")])
        H = H + cf * m
    return H


def probe(mats):
    import oqb_b200  # noqa: F401
    from oqb_b200 import _lib
    lib = _lib.load()
    n = mats[0].shape[0]
    # column-major n×n matrices, as DenseHamiltonian stores them
    hm = np.ascontiguousarray(np.stack([np.asarray(m, dtype=complex).T for m in mats]))
    off, nm, real = C.c_int(), C.c_int(), C.c_int()
    masks, cst = (C.c_int * 16)(), (C.c_int * 16)()
    xv = np.zeros((16, n), dtype=complex)
    st = lib.oqb_dense_structure_probe(n, len(mats), hm.ctypes.data_as(C.c_void_p), C.byref(off), C.byref(nm), masks, cst, C.byref(real),
                                       xv.ctypes.data_as(C.c_void_p))
    assert st == 0
    k = nm.value
    return {"offdiag_term": off.value, "masks": list(masks[:k]), "isconst": list(cst[:k]), "all_real": bool(real.value), "xval": xv[:k]}


def stencil_commutator(u, f, mats, st):
    """The formula of lindblad_stencil64_kernel without the Hadamard factor, element by element in numpy."""
    n = u.shape[0]
    idx = np.arange(n)
    N = st["offdiag_term"]
    acc = np.zeros_like(u)
    for m, x in zip(st["masks"], st["xval"]):
        acc += x[:, None] * u[idx ^ m, :] - u[:, idx ^ m] * x[idx ^ m][None, :]
    du = -1j * f[N] * acc
    for i, h in enumerate(mats):
        if i == N:
            continue
        d = np.diag(h)
        du += -1j * f[i] * (d[:, None] - d[None, :]) * u
    return du


def test_transverse_field_ising_has_one_mask_per_qubit():
    nq, n = 6, 64
    Hd = pauli_sum(nq, [(-1.0, {q: "X"}) for q in range(1, nq + 1)])
    Hp = pauli_sum(nq, [(-1.0, {q: "Z", q + 1: "Z"}) for q in range(1, nq)])
    st = probe([Hd, Hp])
    assert st["offdiag_term"] == 0
    assert st["masks"] == [1 << b for b in range(nq)]
    assert st["isconst"] == [1] * nq and st["all_real"]
    assert np.all(st["xval"] == -1.0)
    st2 = probe([Hp, Hd])  # order of the terms does not matter
    assert st2["offdiag_term"] == 1 and st2["masks"] == st["masks"]


@pytest.mark.parametrize("driver", ["xx_catalyst", "y_field", "xz_products", "x_plus_z"])
@pytest.mark.parametrize("nq", [2, 4, 6])
def test_stencil_tables_reproduce_the_commutator(driver, nq):
    n = 1 << nq
    rng = np.random.default_rng(nq + len(driver))
    if driver == "xx_catalyst":
        Hd = pauli_sum(nq, [(-1.0, {q: "X"}) for q in range(1, nq + 1)] + [(0.4, {q: "X", q + 1: "X"}) for q in range(1, nq)])
    elif driver == "y_field":
        Hd = pauli_sum(nq, [(0.7 + 0.1 * q, {q: "Y"}) for q in range(1, nq + 1)])
    elif driver == "xz_products":
        Hd = pauli_sum(nq, [(-1.0, {q: "X"}) for q in range(1, nq + 1)] + [(0.3, {q: "X", (q % nq) + 1: "Z"}) for q in range(1, nq + 1)])
    else:
        Hd = pauli_sum(nq, [(-1.0, {q: "X"}) for q in range(1, nq + 1)] + [(0.3 + 0.05 * q, {q: "Z"}) for q in range(1, nq + 1)])
    Hp = np.diag(rng.standard_normal(n)).astype(complex)
    Hb = np.diag(rng.standard_normal(n)).astype(complex)
    st = probe([Hp, Hd, Hb])
    assert st["offdiag_term"] == 1 and 0 < len(st["masks"]) <= 16
    assert st["masks"] == sorted(st["masks"])
    assert st["all_real"] == (driver != "y_field")
    if driver == "x_plus_z":
        assert st["masks"][0] == 0  # the term's own diagonal is XOR diagonal 0
    if driver in ("xz_products", "x_plus_z", "y_field"):
        assert 0 in st["isconst"]   # row-signed values
    # the tables rebuild the term …
    idx = np.arange(n)
    R = np.zeros((n, n), dtype=complex)
    for m, x in zip(st["masks"], st["xval"]):
        R[idx, idx ^ m] += x
    assert np.array_equal(R, Hd)
    # … and the stencil formula is the oracle's commutator
    fs = [lambda s: s, lambda s: 1 - s, lambda s: 0.5 * s * s]
    Ho = O.DenseHamiltonian(fs, [Hp, Hd, Hb], unit="ħ")
    u = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    s = 0.37
    got = stencil_commutator(u, [f(s) for f in fs], [Hp, Hd, Hb], st)
    ref = Ho.rhs(u, s)
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-13


def test_no_structure_cases():
    rng = np.random.default_rng(5)
    n = 16
    herm = lambda g: g + g.conj().T
    A, B = herm(rng.standard_normal((n, n)) + 0j), herm(rng.standard_normal((n, n)) + 0j)
    D = np.diag(rng.standard_normal(n)).astype(complex)
    assert probe([A, B])["offdiag_term"] == -1                     # two terms with off-diagonal entries
    A32 = herm(rng.standard_normal((32, 32)) + 0j)
    assert probe([A32])["offdiag_term"] == -1                      # a single term without Pauli structure: nothing to split
    st = probe([A, D])                                             # split form, but a dense random term has n XOR diagonals ≤ 16 here
    assert st["offdiag_term"] == 0 and len(st["masks"]) == 16
    n = 32
    A = herm(rng.standard_normal((n, n)) + 0j)
    D = np.diag(rng.standard_normal(n)).astype(complex)
    st = probe([A, D])
    assert st["offdiag_term"] == 0 and st["masks"] == []           # 32 XOR diagonals: product route
    n = 12                                                         # not a power of two: no XOR structure
    A = herm(rng.standard_normal((n, n)) + 0j)
    D = np.diag(rng.standard_normal(n)).astype(complex)
    st = probe([A, D])
    assert st["offdiag_term"] == 0 and st["masks"] == []


def test_probe_rejects_bad_arguments():
    import oqb_b200  # noqa: F401
    from oqb_b200 import _lib
    lib = _lib.load()
    assert lib.oqb_dense_structure_probe(0, 1, None, None, None, None, None, None, None) != 0
    assert lib.oqb_dense_structure_probe(4, 1, None, None, None, None, None, None, None) != 0


def test_single_term_pauli_hamiltonian_is_recorded():
    """ConstantDenseHamiltonian-shaped input (src/hamiltonian/dense_hamiltonian.jl:149-168): the whole TFIM at fixed s as ONE
    matrix — X part on one XOR diagonal per qubit, ZZ part on XOR diagonal 0."""
    nq, n = 5, 32
    H = pauli_sum(nq, [(-0.6, {q: "X"}) for q in range(1, nq + 1)] + [(-0.4, {q: "Z", q + 1: "Z"}) for q in range(1, nq)])
    st = probe([H])
    assert st["offdiag_term"] == 0 and st["masks"] == [0] + [1 << b for b in range(nq)]
    assert st["isconst"][0] == 0 and st["isconst"][1:] == [1] * nq and st["all_real"]
    rng = np.random.default_rng(9)
    u = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    got = stencil_commutator(u, [1.0], [H], st)
    ref = -1j * (H @ u - u @ H)
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-13


@pytest.mark.parametrize("nq", [3, 5])
def test_fused_stencil_and_hadamard_formula_is_the_lindblad_rhs(nq):
    """The whole one-pass RHS the stencil kernels evaluate for diagonal L_k with shared rates,
        du = stencil commutator + G∘u,  G[r,c] = Σ_k γ_k (d_k[r] conj(d_k[c]) − ½(|d_k[r]|² + |d_k[c]|²)),
    restated in numpy from the probe's tables against the oracle's literal loop (src/opensys/lindblad.jl:94-101 after
    src/opensys/diffeq_liouvillian.jl:70-76)."""
    n = 1 << nq
    rng = np.random.default_rng(40 + nq)
    Hd = pauli_sum(nq, [(-1.0 - 0.1 * q, {q: "X"}) for q in range(1, nq + 1)] + [(0.25, {q: "Y", q + 1: "Y"}) for q in range(1, nq)])
    Hp = np.diag(rng.standard_normal(n)).astype(complex)
    st = probe([Hd, Hp])
    assert st["offdiag_term"] == 0 and len(st["masks"]) > 0
    ds = [np.exp(1j * rng.standard_normal(n)) * rng.standard_normal(n) for _ in range(nq)]  # complex diagonal L_k
    gam = [0.05 * (1 + k) for k in range(nq)]
    fs = [lambda s: 1 - s, lambda s: s]
    opo = O.DiffEqLiouvillian(O.DenseHamiltonian(fs, [Hd, Hp], unit="ħ"), [],
                              [O.LindbladLiouvillian([O.Lindblad(g, np.diag(d)) for g, d in zip(gam, ds)])], n)
    u = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    tf, t = 10.0, 3.7
    s = t / tf
    G = np.zeros((n, n), dtype=complex)
    for g, d in zip(gam, ds):
        G += g * (np.outer(d, d.conj()) - 0.5 * (np.abs(d)[:, None] ** 2 + np.abs(d)[None, :] ** 2))
    got = stencil_commutator(u, [f(s) for f in fs], [Hd, Hp], st) + G * u
    ref = opo.rhs(u, O.ODEParams(None, tf), t)
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-13
