"""Sliced-ELL trajectory kernel (csrc/traj_sell.cu) — the path of SparseHamiltonians without Pauli/XOR structure
(src/hamiltonian/sparse_hamiltonian.jl:70-75, −½γL'L of src/opensys/lindblad.jl:103-109).  The plan permutes rows, pads
slices and reorders slots; none of that may show in the result: every case is compared with scipy's assembled
−iH(s) − ½ΣγL'L applied to each ψ at 1e-13, and with the plain ELL kernels (option traj_ell = 0)."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sps

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

pytestmark = pytest.mark.gpu
TOL = 1e-13


@pytest.fixture(scope="module")
def Q():
    import oqb_b200
    return oqb_b200


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


def random_h(n, per_row, rng, kind="complex", ragged=True):
    """Hermitian (or, kind="imag", anti-symmetric imaginary) random-pattern matrix; `ragged` gives row lengths from 0 up."""
    rows = np.repeat(np.arange(n), per_row)
    cols = rng.integers(0, n, size=n * per_row)
    keep = rows != cols
    if ragged:
        keep &= rng.random(n * per_row) < (0.2 + 0.8 * rows / n)  # short rows at the top, long at the bottom
    vals = rng.standard_normal(n * per_row).astype(complex)
    if kind == "complex":
        vals = vals + 1j * rng.standard_normal(n * per_row)
    A = sps.csc_matrix((vals[keep], (rows[keep], cols[keep])), shape=(n, n))
    if kind == "imag":
        return (1j * (A - A.T)).tocsc()
    return (A + A.conj().T).tocsc()


def reference(H1, H2, Ls, gam, s, psi):
    out = []
    for b in range(psi.shape[0]):
        M = -1j * ((1 - s[b]) * H1 + s[b] * H2)
        for k, L in enumerate(Ls):
            M = M - 0.5 * gam[b if gam.shape[0] > 1 else 0, k] * (L.conj().T @ L)
        out.append(M @ psi[b])
    return np.stack(out)


def run(Q, H1, H2, Ls, psi, s, gam, **opts):
    import ctypes as C
    n = H1.shape[0]
    H = Q.SparseHamiltonian([lambda x: 1 - x, lambda x: x], [H1, H2], unit="ħ")
    op = Q.DiffEqLiouvillian(H, [], [Q.LindbladLiouvillian([Q.Lindblad(0.1, L) for L in Ls])] if Ls else [], n)
    ens = Q.TrajectoryEnsemble(op)
    ctx = Q.get_context()
    pd = Q.DeviceBatch.from_host(psi)
    out = pd.zeros_like()
    cf = np.ascontiguousarray(np.stack([1 - s, s], axis=1))
    g = np.ascontiguousarray(gam, dtype=np.float64)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    import contextlib
    with contextlib.ExitStack() as st:
        for k, v in opts.items():
            st.enter_context(ctx.option(k, v))
        ctx.check(ctx.lib.oqb_traj_matvec(ens.sp, psi.shape[0], dp(cf), 0, dp(g) if Ls else None, 1 if g.shape[0] == 1 else 0, pd.ptr, out.ptr))
        # second call on the same handle: the cached plan (and, for non-diagonal L'L, the refreshed values)
        ctx.check(ctx.lib.oqb_traj_matvec(ens.sp, psi.shape[0], dp(cf), 0, dp(g) if Ls else None, 1 if g.shape[0] == 1 else 0, pd.ptr, out.ptr))
    return out.to_host().reshape(psi.shape[0], -1)


@pytest.mark.parametrize("n", [32, 100, 257, 1024, 4096])
@pytest.mark.parametrize("nb", [1, 5, 301])
@pytest.mark.parametrize("r", [1, 2, 3])
def test_random_pattern_matches_scipy(Q, n, nb, r):
    rng = np.random.default_rng(n * 31 + nb * 7 + r)
    H1 = random_h(n, 6, rng)
    H2 = sps.diags(rng.standard_normal(n)).tocsc().astype(complex)
    Ls = [sps.diags(np.where(rng.random(n) < 0.5, 1.0, -1.0)).tocsc().astype(complex) for _ in range(3)]
    psi = rng.standard_normal((nb, n)) + 1j * rng.standard_normal((nb, n))
    s = rng.random(nb)
    gam = np.full((1, 3), 0.05)
    got = run(Q, H1, H2, Ls, psi, s, gam, traj_kernel=1, sell_r=r)
    assert relerr(got, reference(H1, H2, Ls, gam, s, psi)) < TOL
    plain = run(Q, H1, H2, Ls, psi, s, gam, traj_kernel=1, traj_ell=0)
    assert relerr(got, plain) < TOL


@pytest.mark.parametrize("kind", ["real", "imag", "complex"])
@pytest.mark.parametrize("n", [64, 1000])
def test_value_kinds_and_two_offdiagonal_terms(Q, kind, n):
    """Real terms store 8-byte values, purely imaginary ones their imaginary parts (factor i on the coefficient); two
    off-diagonal terms share the row permutation but are coloured separately."""
    rng = np.random.default_rng(n + len(kind))
    H1 = random_h(n, 5, rng, kind)
    H2 = random_h(n, 3, rng, "complex") + sps.diags(rng.standard_normal(n)).tocsc()
    nb = 9
    psi = rng.standard_normal((nb, n)) + 1j * rng.standard_normal((nb, n))
    s = rng.random(nb)
    got = run(Q, H1, sps.csc_matrix(H2), [], psi, s, np.zeros((1, 0)), traj_kernel=1)
    assert relerr(got, reference(H1, sps.csc_matrix(H2), [], np.zeros((1, 0)), s, psi)) < TOL


@pytest.mark.parametrize("per_member_gamma", [False, True])
def test_nondiagonal_jump_operators(Q, per_member_gamma):
    """L'L with off-diagonal entries: shared γ folds −½Σγ_k L_k'L_k into one term whose VALUES are rebuilt on the device every
    call (the plan keeps the structure and re-gathers them); per-member γ keeps one term per L_k."""
    n, nb, K = 256, 12, 3
    rng = np.random.default_rng(77 + per_member_gamma)
    H1 = random_h(n, 6, rng)
    H2 = sps.diags(rng.standard_normal(n)).tocsc().astype(complex)
    Ls = []
    for _ in range(K):
        rows = rng.integers(0, n, size=2 * n)
        cols = rng.integers(0, n, size=2 * n)
        Ls.append(sps.csc_matrix((rng.standard_normal(2 * n) + 1j * rng.standard_normal(2 * n), (rows, cols)), shape=(n, n)))
    psi = rng.standard_normal((nb, n)) + 1j * rng.standard_normal((nb, n))
    s = rng.random(nb)
    gam = rng.random((nb if per_member_gamma else 1, K))
    got = run(Q, H1, H2, Ls, psi, s, gam, traj_kernel=1)
    assert relerr(got, reference(H1, H2, Ls, gam, s, psi)) < TOL
    # other rates on a fresh handle with the same structure
    gam2 = gam * 0.3 + 0.1
    got2 = run(Q, H1, H2, Ls, psi, s, gam2, traj_kernel=1)
    assert relerr(got2, reference(H1, H2, Ls, gam2, s, psi)) < TOL


def test_empty_rows_and_dense_row(Q):
    """Rows without entries, one row touching every column, and a matrix whose only off-diagonal entries sit in one column."""
    n, nb = 128, 4
    rng = np.random.default_rng(5)
    M = sps.lil_matrix((n, n), dtype=complex)
    M[7, :] = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    M[:, 3] = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).reshape(n, 1)
    H1 = sps.csc_matrix(M + M.conj().T)
    H2 = sps.diags(rng.standard_normal(n)).tocsc().astype(complex)
    psi = rng.standard_normal((nb, n)) + 1j * rng.standard_normal((nb, n))
    s = rng.random(nb)
    got = run(Q, H1, H2, [], psi, s, np.zeros((1, 0)), traj_kernel=1)
    assert relerr(got, reference(H1, H2, [], np.zeros((1, 0)), s, psi)) < TOL


def test_default_dispatch_uses_it_for_unstructured_operators(Q):
    """Without any option an unstructured operator takes traj_sell_kernel (kernel list of the profile hooks)."""
    n, nb = 512, 8
    rng = np.random.default_rng(11)
    H1 = random_h(n, 6, rng, ragged=False)
    H2 = sps.diags(rng.standard_normal(n)).tocsc().astype(complex)
    psi = rng.standard_normal((nb, n)) + 1j * rng.standard_normal((nb, n))
    s = rng.random(nb)
    got = run(Q, H1, H2, [], psi, s, np.zeros((1, 0)))
    assert relerr(got, reference(H1, H2, [], np.zeros((1, 0)), s, psi)) < TOL
