"""Randomised differential tests: random small problems × random combinations of the per-context options that select
between kernels/routes.  Every combination must reproduce the oracle's literal restatement of the reference
(src/hamiltonian/dense_hamiltonian.jl:84-89, src/opensys/lindblad.jl:94-101, src/hamiltonian/sparse_hamiltonian.jl:70-75)
at 1e-12 — no option may change the answer beyond rounding."""
import contextlib
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


@pytest.fixture(scope="module")
def Q():
    import oqb_b200
    return oqb_b200


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


def rand_c(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


@contextlib.contextmanager
def options(ctx, opts):
    with contextlib.ExitStack() as st:
        for k, v in opts.items():
            st.enter_context(ctx.option(k, v))
        yield


@pytest.mark.parametrize("seed", range(40))
def test_dense_lindblad_rhs_under_random_options(Q, seed):
    rng = np.random.default_rng(9000 + seed)
    n = int(rng.choice([3, 8, 17, 32, 50, 64, 96, 130]))
    nterms = int(rng.integers(1, 4))
    K = int(rng.integers(0, 5))
    nb = int(rng.integers(1, 7))
    real_terms = bool(rng.integers(0, 2))
    herm = (lambda g: g + g.T) if real_terms else (lambda g: g + g.conj().T)
    mats = []
    n_off = 0
    for i in range(nterms):
        if rng.random() < 0.5 or (i == nterms - 1 and n_off == 0 and rng.random() < 0.7):
            mats.append(herm(rng.standard_normal((n, n)) if real_terms else rand_c(rng, n, n)).astype(complex))
            n_off += 1
        else:
            mats.append(np.diag(rng.standard_normal(n)).astype(complex))  # a diagonal (σz-type) term
    fs = [(lambda s, a=a, b=b: a * (1 - s) + b * s * s) for a, b in rng.standard_normal((nterms, 2))]
    diagL = bool(rng.integers(0, 2))
    Ls = [np.diag(rand_c(rng, n)) if diagL else rand_c(rng, n, n) / np.sqrt(n) for _ in range(K)]
    const_rates = bool(rng.integers(0, 2))
    gfs = [(lambda s, g=g: g) if const_rates else (lambda s, g=g: g * (1 + s)) for g in rng.uniform(0.05, 0.5, K)]
    per_member = bool(rng.integers(0, 2))
    Hg, Ho = Q.DenseHamiltonian(fs, mats, unit="ħ"), O.DenseHamiltonian(fs, mats, unit="ħ")
    lg = [Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gfs, Ls)])] if K else []
    lo = [O.LindbladLiouvillian([O.Lindblad(g, L) for g, L in zip(gfs, Ls)])] if K else []
    opg, opo = Q.DiffEqLiouvillian(Hg, [], lg, n), O.DiffEqLiouvillian(Ho, [], lo, n)
    u = rand_c(rng, nb, n, n)
    tf = 4.0
    t = rng.random(nb) * tf if per_member else float(rng.random() * tf)
    tl = t if per_member else [t] * nb
    ref = np.stack([opo.rhs(u[b], O.ODEParams(None, tf), tl[b]) for b in range(nb)])
    ctx = Q.get_context()
    for _ in range(3):
        opts = {k: int(rng.integers(0, 2)) for k in ("lindblad_split", "lindblad_fuse", "lindblad_diag", "real_operand", "zgemm_3m", "zgemm_tma")}
        with options(ctx, opts):
            got = opg(np.zeros_like(u), u, Q.ODEParams(None, tf), t)
        assert relerr(got, ref) < TOL, (opts, n, nterms, K, diagL, const_rates, per_member, real_terms)


def _rand_sparse_term(rng, n, kind):
    if kind == "diag":
        return sps.diags(rng.standard_normal(n)).tocsc().astype(complex)
    if kind == "xor":  # Pauli-X-type strings: M[r, r ^ m] = v
        nq = int(np.log2(n))
        idx = np.arange(n)
        rows, cols, vals = [], [], []
        for m in rng.choice(np.arange(1, n), size=min(nq, 5), replace=False):
            rows.append(idx ^ int(m)); cols.append(idx); vals.append(np.full(n, rng.standard_normal()))
        A = sps.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n), dtype=complex)
        return ((A + A.T) * 0.5).tocsc()
    per_row = int(rng.integers(1, 6))
    rows = np.repeat(np.arange(n), per_row)
    cols = rng.integers(0, n, size=n * per_row)
    keep = rng.random(n * per_row) < 0.8
    vals = rng.standard_normal(n * per_row).astype(complex)
    if kind == "complex":
        vals = vals + 1j * rng.standard_normal(n * per_row)
    A = sps.csc_matrix((vals[keep], (rows[keep], cols[keep])), shape=(n, n))
    return (A + A.conj().T).tocsc()


@pytest.mark.parametrize("seed", range(40))
def test_trajectory_matvec_under_random_options(Q, seed):
    rng = np.random.default_rng(7000 + seed)
    n = int(rng.choice([32, 64, 100, 128, 257, 512]))
    pow2 = (n & (n - 1)) == 0
    nterms = int(rng.integers(1, 4))
    kinds = [str(rng.choice(["diag", "real", "complex"] + (["xor"] if pow2 else []))) for _ in range(nterms)]
    mats = [_rand_sparse_term(rng, n, k) for k in kinds]
    K = int(rng.integers(0, 4))
    Ls = []
    for _ in range(K):
        if rng.random() < 0.6:
            Ls.append(sps.diags(rng.choice([-1.0, 1.0], n)).tocsc().astype(complex))
        else:
            r, c_ = rng.integers(0, n, size=n), rng.integers(0, n, size=n)
            Ls.append(sps.csc_matrix((rand_c(rng, n), (r, c_)), shape=(n, n)))
    fs = [(lambda s, a=a, b=b: a * (1 - s) + b * s) for a, b in rng.standard_normal((nterms, 2))]
    nb = int(rng.integers(1, 40))
    H = Q.SparseHamiltonian(fs, mats, unit="ħ")
    op = Q.DiffEqLiouvillian(H, [], [Q.LindbladLiouvillian([Q.Lindblad(0.1, L) for L in Ls])] if K else [], n)
    ens = Q.TrajectoryEnsemble(op)
    ctx = Q.get_context()
    psi = rand_c(rng, nb, n)
    s = rng.random(nb)
    cf = np.ascontiguousarray(np.stack([[f(x) for f in fs] for x in s]))
    per_member_gamma = bool(rng.integers(0, 2)) and K > 0
    gam = rng.uniform(0.05, 0.5, (nb if per_member_gamma else 1, max(K, 1)))[:, :K] if K else np.zeros((1, 0))
    gam = np.ascontiguousarray(gam)
    ref = []
    for b in range(nb):
        M = sum(-1j * cf[b, i] * mats[i] for i in range(nterms))
        for k, L in enumerate(Ls):
            M = M - 0.5 * gam[b if per_member_gamma else 0, k] * (L.conj().T @ L)
        ref.append(M @ psi[b])
    ref = np.stack(ref)
    pd = Q.DeviceBatch.from_host(psi)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    for _ in range(3):
        opts = {"traj_kernel": int(rng.integers(0, 3)), "traj_ell": int(rng.integers(0, 2)), "sell_r": int(rng.integers(0, 4)),
                "sell_unit": int(rng.integers(1, 3))}
        # a fresh handle per combination: plans are cached per handle and sell_unit is read when a plan is built
        ens = Q.TrajectoryEnsemble(Q.DiffEqLiouvillian(Q.SparseHamiltonian(fs, mats, unit="ħ"), [],
                                                       [Q.LindbladLiouvillian([Q.Lindblad(0.1, L) for L in Ls])] if K else [], n))
        out = pd.zeros_like()
        with options(ctx, opts):
            ctx.check(ctx.lib.oqb_traj_matvec(ens.sp, nb, dp(cf), 0, dp(gam) if K else None, 0 if per_member_gamma else 1, pd.ptr, out.ptr))
        assert relerr(out.to_host().reshape(nb, n), ref) < TOL, (opts, n, kinds, K, per_member_gamma)


def _spectrum(rng, n, kind):
    """Eigenvalues with the structures GapIndices cares about: generic, exactly repeated levels, equally spaced ladders
    (many pairs per gap), splittings below the 1e-8 zero threshold, and gaps that coincide only after 8-digit rounding."""
    if kind == "generic":
        w = np.sort(rng.standard_normal(n) * 3)
    elif kind == "degenerate":
        w = np.sort(rng.choice(rng.standard_normal(max(2, n // 3)) * 3, size=n))
    elif kind == "ladder":
        w = np.sort(rng.integers(-4, 5, size=n).astype(float) * 0.75)
    elif kind == "tiny_splittings":
        base = np.sort(rng.choice(rng.standard_normal(max(2, n // 2)) * 3, size=n))
        w = np.sort(base + rng.uniform(-4e-9, 4e-9, n))
    else:  # "rounded": gaps equal to 8 significant digits but not exactly
        w = np.sort(np.cumsum(rng.choice([0.5, 1.25, 0.5 * (1 + 3e-10), 1.25 * (1 - 2e-10)], size=n)))
    return w


@pytest.mark.parametrize("seed", range(36))
def test_davies_ame_rhs_under_random_options(Q, seed):
    from oracle import davies_groups as DG
    rng = np.random.default_rng(5000 + seed)
    n = int(rng.choice([4, 7, 12, 16, 24, 33, 48]))
    kind = str(rng.choice(["generic", "degenerate", "ladder", "tiny_splittings", "rounded"]))
    w = _spectrum(rng, n, kind)
    real_v = bool(rng.integers(0, 2))
    g = rng.standard_normal((n, n)) if real_v else rand_c(rng, n, n)
    v, _ = np.linalg.qr(g)
    v = v.astype(complex)
    Hm = (v * w) @ v.conj().T
    Hm = 0.5 * (Hm + Hm.conj().T)
    ncoup = int(rng.integers(1, 4))
    herm = lambda a: a + a.conj().T
    coup = [herm(rng.standard_normal((n, n)).astype(complex) if real_v else rand_c(rng, n, n)) for _ in range(ncoup)]
    ngen = int(rng.integers(1, 3))
    gam = [(lambda x, a=a: (x + a if x >= 0 else (a - x) * np.exp(x))) for a in rng.uniform(0.5, 1.5, ngen)]  # mock bath of test/opensys/davies.jl:7-8
    lam = [(lambda x, b=b: b * x + 0.1) for b in rng.uniform(0.5, 1.5, ngen)]
    idxs = [sorted(rng.choice(ncoup, size=int(rng.integers(1, ncoup + 1)), replace=False).tolist()) for _ in range(ngen)]
    Hg = Q.DenseHamiltonian([lambda s: 1.0], [Hm], unit="ħ")
    gens = [Q.DaviesGenerator([coup[k] for k in idx], gm, lm) for idx, gm, lm in zip(idxs, gam, lam)]
    opg = Q.DiffEqLiouvillian(Hg, gens, [], n)
    nb = int(rng.integers(1, 4))
    u = rand_c(rng, nb, n, n)
    ref = np.stack([DG.ame_rhs(x, w, v, coup, [(gm, lm, idx) for idx, gm, lm in zip(idxs, gam, lam)], 8, 8) for x in u])
    ctx = Q.get_context()
    for _ in range(3):
        opts = {"davies_dense_min_pairs": int(rng.choice([0, 2, 6, 10 ** 9])), "fuse_cross": int(rng.integers(0, 2)),
                "real_operand": int(rng.integers(0, 2)), "zgemm_3m": int(rng.integers(0, 2)), "zgemm_tma": int(rng.integers(0, 2)),
                "debug_checks": 1}
        with options(ctx, opts):
            opg.clear_plans()
            du = opg.rhs_with_eig(np.zeros_like(u), u, Q.ODEParams(None, 1.0), 0.0, w, v)
        assert relerr(du, ref) < TOL, (opts, n, kind, real_v, ncoup, idxs)
