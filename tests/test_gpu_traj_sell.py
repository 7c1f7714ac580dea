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


def circulant_h(n, rng):
    """Hermitian band/circulant matrix with random complex values: every row has 6 entries, no Pauli structure — the plan
    keeps the natural row order (σ = identity), which is when the RK4 stages are fused into the kernel's store."""
    rows, cols, vals = [], [], []
    for k in (1, 5, 17):
        v = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        rows.append(np.arange(n)); cols.append((np.arange(n) + k) % n); vals.append(v)
    A = sps.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    return (A + A.conj().T).tocsc()


@pytest.mark.parametrize("n,nb,fuse,kind", [(48, 7, 1, "random"), (200, 5, 1, "random"), (256, 6, 1, "random"), (256, 6, 0, "random"),
                                            (64, 5, 1, "circulant"), (100, 6, 1, "circulant"), (256, 5, 1, "circulant"), (256, 5, 0, "circulant")])
def test_evolve_unstructured_hamiltonian_matches_oracle(Q, n, nb, fuse, kind):
    """oqb_traj_evolve on a SparseHamiltonian without Pauli structure.  With the natural row order kept (circulant) the RK4
    stage arithmetic rides in the sliced-ELL kernel's store (option rk_fuse = 0: separate stage kernels); with permuted rows
    (random pattern) the stage kernels stay separate.  Jump steps, operator indices and RNG consumption identical to the
    oracle's loop, ψ within 1e-12."""
    from oracle import oqb_oracle as O
    rng = np.random.default_rng(n + nb)
    H1 = random_h(n, 4, rng) if kind == "random" else circulant_h(n, rng)
    H2 = sps.diags(rng.standard_normal(n)).tocsc().astype(complex)
    Ls = [sps.diags(np.where(rng.random(n) < 0.5, 1.0, -0.5)).tocsc().astype(complex) for _ in range(3)]
    fa, fb = (lambda s: 1 - s), (lambda s: s)
    gfs = [(lambda s, k=k: 0.5 * (1 + 0.3 * k + 0.5 * s)) for k in range(3)]
    Hg, Ho = Q.SparseHamiltonian([fa, fb], [H1, H2], unit="ħ"), O.SparseHamiltonian([fa, fb], [H1, H2], unit="ħ")
    opg = Q.DiffEqLiouvillian(Hg, [], [Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gfs, Ls)])], n)
    opo = O.DiffEqLiouvillian(Ho, [], [O.LindbladLiouvillian([O.Lindblad(g, L.toarray()) for g, L in zip(gfs, Ls)])], n)
    ens = Q.TrajectoryEnsemble(opg)
    tf, dt, nsteps, seed = 2.0, 0.02, 50, 91
    pg, po = Q.ODEParams(None, tf), O.ODEParams(None, tf)
    psi0 = rng.standard_normal((nb, n)) + 1j * rng.standard_normal((nb, n))
    psi0 /= np.linalg.norm(psi0, axis=1, keepdims=True)
    pd = Q.DeviceBatch.from_host(psi0)
    st = Q.TrajectoryStepper(ens, nb, seed=seed, max_log=64)
    ctx = Q.get_context()
    with ctx.option("rk_fuse", fuse):
        st.evolve(pd, pg, 0.0, dt, 20)
        st.evolve(pd, pg, 20 * dt, dt, nsteps - 20)
    got = pd.to_host().reshape(nb, n)
    nj, log, thr = st.njumps.cpu().numpy(), st.log.cpu().numpy(), st.thr.cpu().numpy()
    Ld = [L.toarray() for L in Ls]
    total = 0
    for b in range(nb):
        r = O.Xoshiro256pp.from_seed(seed, b)
        thr0 = r.rand()
        ref, thr_end, lg_ref = O.traj_evolve_rk4(lambda t: opo.update_cache(po, t), lambda t: Ld, lambda t: [g(po(t)) for g in gfs],
                                                 psi0[b], r, thr0, 0.0, dt, nsteps)
        assert nj[b] == len(lg_ref)
        assert [tuple(x) for x in log[b, :nj[b]]] == lg_ref
        assert thr[b] == thr_end
        assert relerr(got[b], ref) < 1e-12
        total += len(lg_ref)
    assert total > 0


@pytest.mark.parametrize("kind", ["circulant", "random"])
def test_rk4_stage_fusion_follows_the_row_order(Q, kind):
    """Launch counts of oqb_traj_evolve: with σ = identity the 4 stage kernels and the norm pass of a step disappear into the
    matvec launches; with permuted rows the option changes nothing."""
    n, nb, steps = 256, 4, 10
    rng = np.random.default_rng(3)
    H1 = random_h(n, 4, rng) if kind == "random" else circulant_h(n, rng)
    H2 = sps.diags(rng.standard_normal(n)).tocsc().astype(complex)
    Ls = [sps.diags(np.where(rng.random(n) < 0.5, 1.0, -0.5)).tocsc().astype(complex) for _ in range(2)]
    H = Q.SparseHamiltonian([lambda s: 1 - s, lambda s: s], [H1, H2], unit="ħ")
    op = Q.DiffEqLiouvillian(H, [], [Q.LindbladLiouvillian([Q.Lindblad(1e-9, L) for L in Ls])], n)  # no jumps within the run
    ens = Q.TrajectoryEnsemble(op)
    ctx = Q.get_context()
    psi0 = rng.standard_normal((nb, n)) + 1j * rng.standard_normal((nb, n))
    psi0 /= np.linalg.norm(psi0, axis=1, keepdims=True)
    counts, outs = {}, {}
    for fuse in (1, 0):
        pd = Q.DeviceBatch.from_host(psi0)
        st = Q.TrajectoryStepper(ens, nb, seed=5)
        with ctx.option("rk_fuse", fuse):
            st.evolve(pd, Q.ODEParams(None, 1.0), 0.0, 0.01, 2)
            l0 = ctx.launch_count()
            st.evolve(pd, Q.ODEParams(None, 1.0), 0.02, 0.01, steps)
            counts[fuse] = (ctx.launch_count() - l0) / steps
        outs[fuse] = pd.to_host()
    assert relerr(outs[1], outs[0]) < 1e-13
    if kind == "circulant":
        assert counts[1] <= counts[0] - 4, counts
    else:
        assert abs(counts[1] - counts[0]) <= 1, counts


@pytest.mark.parametrize("n,nb", [(5000, 5), (8192, 3)])
def test_large_dimensions_with_one_or_two_ring_slots(Q, n, nb):
    """n = 5000: two 80 KB slots (both members of a pair are loaded at the top of their own iteration); n = 8192 (13 qubits):
    a single 131 KB slot, one trajectory per operator fetch."""
    rng = np.random.default_rng(n)
    H1 = random_h(n, 4, rng)
    H2 = sps.diags(rng.standard_normal(n)).tocsc().astype(complex)
    psi = rng.standard_normal((nb, n)) + 1j * rng.standard_normal((nb, n))
    s = rng.random(nb)
    got = run(Q, H1, H2, [], psi, s, np.zeros((1, 0)), traj_kernel=1)
    assert relerr(got, reference(H1, H2, [], np.zeros((1, 0)), s, psi)) < TOL


def test_more_terms_than_the_plan_holds_fall_back(Q):
    """7 off-diagonal terms exceed the sliced-ELL plan's 6: the plain kernels take over, same answer."""
    import ctypes as C
    n, nb, nt = 64, 5, 7
    rng = np.random.default_rng(21)
    mats = [random_h(n, 3, rng) for _ in range(nt)]
    fs = [(lambda s, a=a: a + s) for a in rng.standard_normal(nt)]
    H = Q.SparseHamiltonian(fs, mats, unit="ħ")
    ens = Q.TrajectoryEnsemble(Q.DiffEqLiouvillian(H, [], [], n))
    ctx = Q.get_context()
    psi = rng.standard_normal((nb, n)) + 1j * rng.standard_normal((nb, n))
    s = rng.random(nb)
    cf = np.ascontiguousarray(np.stack([[f(x) for f in fs] for x in s]))
    pd = Q.DeviceBatch.from_host(psi)
    out = pd.zeros_like()
    with ctx.option("traj_kernel", 1):
        ctx.check(ctx.lib.oqb_traj_matvec(ens.sp, nb, cf.ctypes.data_as(C.POINTER(C.c_double)), 0, None, 1, pd.ptr, out.ptr))
    pl, pad = C.c_int(), C.c_double()
    ctx.check(ctx.lib.oqb_traj_sell_stats(ens.sp, C.byref(pl), C.byref(pad)))
    assert pl.value == 0
    ref = np.stack([sum(-1j * cf[b, i] * mats[i] for i in range(nt)) @ psi[b] for b in range(nb)])
    assert relerr(out.to_host().reshape(nb, n), ref) < TOL
