"""GPU tests at BASELINE.json's full sizes.  The oracle is too slow to check whole batches there, so
these use (a) the oracle on a few members and (b) size-independent properties of the path:
linearity, trace preservation, Hermiticity preservation, the norm-decay identity
d‖ψ‖²/dt = −Σ_k γ_k‖L_kψ‖² that ties the trajectory matvec to the jump probabilities, and
agreement between the device-resident and host-buffer entry points."""
import numpy as np
import pytest
import scipy.sparse as sps

from oracle import fixtures as F
from oracle import oqb_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def Q():
    import oqb_b200
    oqb_b200.get_context()
    return oqb_b200


def relerr(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0)


def rand_c(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def lowrank_rho(rng, n, r=8):
    g = rand_c(rng, n, r)
    rho = g @ g.conj().T
    return rho / np.trace(rho).real


def tfim_dense(nq, rng=None):
    """H_d = −ΣX_i (dense), H_p = Σ J Z_iZ_{i+1} or random Ising (diagonal), Z_k list."""
    n = 1 << nq
    idx = np.arange(n)
    Hd = np.zeros((n, n), dtype=complex)
    for q in range(nq):
        Hd[idx ^ (1 << q), idx] = -1.0
    z = [1.0 - 2.0 * ((idx >> (nq - 1 - q)) & 1) for q in range(nq)]
    if rng is None:
        diag = -sum(z[i] * z[i + 1] for i in range(nq - 1))
    else:
        diag = sum((2 * rng.random() - 1) * z[i] * z[j] for i in range(nq) for j in range(i + 1, nq))
    return Hd, np.diag(diag).astype(complex), z


# ---- C3: 10-qubit Davies / AME, n = lvl = 1024 -----------------------------------------------
def test_c3_ame_1024_properties(Q):
    rng = np.random.default_rng(3000)
    nq, n = 10, 1024
    Hd, Hp, z = tfim_dense(nq, rng)
    coup = [np.diag(zq).astype(complex) for zq in z]
    H = Q.DenseHamiltonian([lambda s: 1 - s, lambda s: s], [Hd, Hp])
    dav = Q.DaviesGenerator(Q.ConstantCouplings(coup, unit="h"), Q.Ohmic(1e-4, 4, 16), Q.ZERO_LAMBSHIFT)
    op = Q.DiffEqLiouvillian(H, [dav], [], n)
    p = Q.ODEParams(H, 10.0)
    u1, u2 = lowrank_rho(rng, n), lowrank_rho(rng, n)
    a, b = 0.3 - 0.2j, 1.1 + 0.4j
    u = np.stack([u1, u2, a * u1 + b * u2, rand_c(rng, n, n)])
    du = op(np.zeros_like(u), u, p, 5.0)
    assert relerr(du[2], a * du[0] + b * du[1]) < 1e-12                      # linearity
    for k in range(4):
        assert abs(np.trace(du[k])) < 1e-9 * np.linalg.norm(du[k])            # trace preserving (lvl = n)
    assert relerr(du[0], du[0].conj().T) < 1e-12                               # Hermiticity preserving
    # host-buffer entry point == device-resident entry point (bit-identical kernels)
    uh = np.ascontiguousarray(u.transpose(0, 2, 1))
    duh = np.zeros_like(uh)
    op.rhs_host(duh, uh, p, 5.0)
    assert np.array_equal(duh.transpose(0, 2, 1), du)
    # steady-state check: the Gibbs state of H(s) is a fixed point of the Davies generator
    # (KMS condition γ(−ω) = e^{−βω}γ(ω) of the Ohmic spectrum) — a physics-level known answer.
    w, v = op._w, op._v
    beta = Q.Ohmic(1e-4, 4, 16).beta
    pop = np.exp(-beta * (w - w.min()))
    pop /= pop.sum()
    gibbs = (v * pop) @ v.conj().T
    dg = op(np.zeros((1, n, n), complex), gibbs[None], p, 5.0)[0]
    rate_scale = np.linalg.norm(du[0]) / np.linalg.norm(u[0])
    assert np.linalg.norm(dg) < 1e-9 * rate_scale * np.linalg.norm(gibbs)


def test_c3_ame_1024_vs_group_oracle(Q):
    """The headline configuration — 10 qubits, n = lvl = 1024, K = 10 per-qubit Z couplings, Ohmic bath — at 1e-12 against
    the independent group oracle (oracle/davies_groups.py: dict grouping by the reference's scalar rounding rule,
    sub-matrix algebra for the multi-pair groups; validated against the literal loop at l ≤ 40 in the CPU suite).  The
    C3 spectrum has thousands of gap pairs that coincide after rounding to 8 digits, so the cross-term lists, their int32
    index tables and the general (non-fused) branch of the RHS are all exercised.  Covered: device-evaluated Ohmic
    tables, host-evaluated closure tables, real-operand products on and off, the 3M and 4M complex products, and the
    host-buffer entry point."""
    import scipy.linalg as sla
    from oracle import davies_groups as DG
    rng = np.random.default_rng(3000)
    nq, n = 10, 1024
    Hd, Hp, z = tfim_dense(nq, rng)
    coup = [2 * np.pi * np.diag(zq).astype(complex) for zq in z]
    Hm = 2 * np.pi * (0.5 * Hd + 0.5 * Hp)
    w, v = sla.eigh(Hm.real)  # real symmetric: exactly real eigenvectors (the real-operand products are eligible)
    v = v.astype(complex)
    bath, ob = Q.Ohmic(1e-4, 4, 16), O.Ohmic(1e-4, 4, 16)
    H = Q.DenseHamiltonian([lambda s: 1 - s, lambda s: s], [Hd, Hp])
    p = Q.ODEParams(H, 10.0)
    u = np.stack([lowrank_rho(rng, n), rand_c(rng, n, n) / n])
    groups = DG.gap_groups(w, 8, 8)
    assert sum(len(m) > 1 for m in groups[0].values()) > 100  # the coincidences this test is about are present
    rot = [v.conj().T @ c @ v for c in coup]
    ref = []
    for x in u:
        rho = v.conj().T @ x @ v
        X = -1j * (w[:, None] * rho - rho * w[None, :])
        DG.davies_eig_acc(X, rho, rot, w, ob.spectrum, lambda x_: 0.0, 8, 8, groups)
        ref.append(v @ X @ v.conj().T)
    ref = np.stack(ref)
    ctx = Q.get_context()
    op_dev = Q.DiffEqLiouvillian(H, [Q.DaviesGenerator(Q.ConstantCouplings(coup, unit="ħ"), bath, Q.ZERO_LAMBSHIFT)], [], n)
    op_host = Q.DiffEqLiouvillian(H, [Q.DaviesGenerator(Q.ConstantCouplings(coup, unit="ħ"), ob.spectrum, lambda x_: 0.0)], [], n)
    for real_op in (1, 0):
        for mode3m in (1, 0):
            with ctx.option("real_operand", real_op), ctx.option("zgemm_3m", mode3m):
                for op in (op_dev, op_host):
                    op.clear_plans()
                    du = op.rhs_with_eig(np.zeros_like(u), u, p, 5.0, w, v)
                    assert relerr(du, ref) < TOL, (real_op, mode3m, op is op_dev)
    # the Hadamard kept in the rotation product's epilogue with the cross-term sources recovered as X/C (default) against
    # the separate-pass form: both against the oracle
    for fuse in (1, 0):
        with ctx.option("fuse_cross", fuse):
            op_dev.clear_plans()
            du = op_dev.rhs_with_eig(np.zeros_like(u), u, p, 5.0, w, v)
            assert relerr(du, ref) < TOL, fuse
    # host-buffer entry point (Julia layout: column-major blocks)
    op_dev.clear_plans()
    op_dev._prepare_ame(0.5, w, v)
    uh = np.ascontiguousarray(u.transpose(0, 2, 1))
    duh = np.zeros_like(uh)
    ctx.check(ctx.lib.oqb_ame_rhs_host(op_dev._ame, 2, uh.ctypes.data, duh.ctypes.data))
    assert relerr(duh.transpose(0, 2, 1), ref) < TOL


def test_c3_degenerate_start_1024_vs_group_oracle(Q):
    """The anneal's starting point at full size: H(0) = −ΣX on 10 qubits (11 levels with binomial degeneracies; the zero-gap
    group holds 91 866 level pairs, every other group tens of thousands) with the K = 10 per-qubit Z couplings — the dense
    per-group path (masked couplings as stacked products) against the independent group oracle at 1e-12 with the SAME (w, v),
    and the library's one-call route (its own device eigensolver: another basis inside every degenerate level, the RHS is
    invariant) at 1e-9."""
    import scipy.linalg as sla
    from oqb_b200 import host as Hm
    from oracle import davies_groups as DG
    rng = np.random.default_rng(3100)
    nq, n = 10, 1024
    Hd, Hp, z = tfim_dense(nq, rng)
    coup = [2 * np.pi * np.diag(zq).astype(complex) for zq in z]
    w, v = sla.eigh((2 * np.pi * Hd).real)
    v = v.astype(complex)
    bath, ob = Q.Ohmic(1e-4, 4, 16), O.Ohmic(1e-4, 4, 16)
    H = Q.DenseHamiltonian([lambda s: 1 - s, lambda s: s], [Hd, Hp])
    p = Q.ODEParams(H, 10.0)
    u = np.stack([lowrank_rho(rng, n)])
    groups = DG.gap_groups(w, 8, 8)
    assert len(groups[1]) == 91866 and len(groups[0]) == 10
    rot = [v.conj().T @ c @ v for c in coup]
    rho = v.conj().T @ u[0] @ v
    X = -1j * (w[:, None] * rho - rho * w[None, :])
    DG.davies_eig_acc(X, rho, rot, w, ob.spectrum, lambda x_: 0.0, 8, 8, groups)
    ref = (v @ X @ v.conj().T)[None]
    ctx = Q.get_context()
    op = Q.DiffEqLiouvillian(H, [Q.DaviesGenerator(Q.ConstantCouplings(coup, unit="ħ"), bath, Q.ZERO_LAMBSHIFT)], [], n)
    with ctx.option("debug_checks", 1):
        du = op.rhs_with_eig(np.zeros_like(u), u, p, 0.0, w, v)
        stats = Hm.davies_stats(ctx, op._ame)
    assert relerr(du, ref) < TOL
    assert stats[3] >= 3 and 0 < stats[4] <= 40   # zero group and the ±2·(2π) groups dense; only the non-vanishing blocks are kept
    op.clear_plans()
    du2 = op(np.zeros_like(u), u, p, 0.0)          # library route: device eigensolver, tables, dense groups behind one call
    assert relerr(du2, ref) < 1e-9
    op.clear_plans()


def test_long_k_real_operand_kernel_vs_numpy(Q):
    """The kernel the headline number is timed on — zgemm_tma_kernel real-operand variant, K > 512 (128×64 CTA) — checked
    directly: with no eigenbasis term the AME RHS is v(C_H∘(v'uv))v', C_H[a,b] = −i(w_a − w_b), i.e. four long-K products
    with a real operand, against numpy at n = 1024."""
    import scipy.linalg as sla
    rng = np.random.default_rng(77)
    n = 1024
    Hm = rng.standard_normal((n, n)); Hm = (Hm + Hm.T) / 2
    w, v = sla.eigh(Hm)
    H = Q.DenseHamiltonian([lambda s: 1.0], [Hm.astype(complex)], unit="ħ")
    op = Q.DiffEqLiouvillian(H, [Q.DaviesGenerator([np.eye(n, dtype=complex)], lambda x: 0.0, lambda x: 0.0)], [], n)
    u = np.stack([rand_c(rng, n, n), lowrank_rho(rng, n)])
    ctx = Q.get_context()
    rho = v.T @ u @ v  # v' u v (v real), broadcast over the batch
    X = -1j * (w[None, :, None] - w[None, None, :]) * rho
    ref = v @ X @ v.T
    for real_op in (1, 0):
        with ctx.option("real_operand", real_op):
            op.clear_plans()
            du = op.rhs_with_eig(np.zeros_like(u), u, Q.ODEParams(H, 1.0), 0.5, w, v.astype(complex))
            import ctypes as C
            flag = C.c_int()
            ctx.check(ctx.lib.oqb_ame_eigenvectors_real(op._ame, C.byref(flag)))
            assert flag.value == 1
            assert relerr(du, ref) < 1e-13, real_op


# ---- C2: 8-qubit dense TFIM + 8 dephasing Lindblad operators, n = 256 ------------------------------
def test_c2_lindblad_256_vs_oracle_and_properties(Q):
    rng = np.random.default_rng(2000)
    nq, n, K, nb = 8, 256, 8, 12
    Hd, Hp, z = tfim_dense(nq)
    Ls = [np.diag(zq).astype(complex) for zq in z]
    fs = [lambda s: 1 - s, lambda s: s]
    gfs = [lambda s, k=k: 0.1 * (1 + k / 8) for k in range(K)]
    Hg, Ho = Q.DenseHamiltonian(fs, [Hd, Hp], unit="ħ"), O.DenseHamiltonian(fs, [Hd, Hp], unit="ħ")
    opg = Q.DiffEqLiouvillian(Hg, [], [Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gfs, Ls)])], n)
    opo = O.DiffEqLiouvillian(Ho, [], [O.LindbladLiouvillian([O.Lindblad(g, L) for g, L in zip(gfs, Ls)])], n)
    u = np.stack([lowrank_rho(rng, n) for _ in range(nb)])
    t = np.arange(nb) / nb * 10.0  # s_b = b/B: per-member coefficients
    du = opg(np.zeros_like(u), u, Q.ODEParams(None, 10.0), t)
    for b in (0, 5, 11):
        assert relerr(du[b], opo.rhs(u[b], O.ODEParams(None, 10.0), t[b])) < TOL
    for b in range(nb):
        assert abs(np.trace(du[b])) < 1e-12 * np.linalg.norm(du[b]) * n
        assert relerr(du[b], du[b].conj().T) < 1e-12
    # general dense L_k (structure cannot be exploited)
    Ld = [rand_c(rng, n, n) / np.sqrt(n) for _ in range(K)]
    opg2 = Q.DiffEqLiouvillian(Hg, [], [Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gfs, Ld)])], n)
    opo2 = O.DiffEqLiouvillian(Ho, [], [O.LindbladLiouvillian([O.Lindblad(g, L) for g, L in zip(gfs, Ld)])], n)
    du2 = opg2(np.zeros_like(u[:3]), u[:3], Q.ODEParams(None, 10.0), t[:3])
    for b in range(3):
        assert relerr(du2[b], opo2.rhs(u[b], O.ODEParams(None, 10.0), t[b])) < TOL


# ---- C4: 12-qubit sparse TFIM trajectories, n = 4096 ----------------------------------------------
def tfim_sparse(nq):
    n = 1 << nq
    idx = np.arange(n)
    rows = np.concatenate([idx ^ (1 << q) for q in range(nq)])
    cols = np.tile(idx, nq)
    Hx = sps.csc_matrix((-np.ones(n * nq), (rows, cols)), shape=(n, n), dtype=complex)
    z = [1.0 - 2.0 * ((idx >> (nq - 1 - q)) & 1) for q in range(nq)]
    Hz = sps.diags(-sum(z[i] * z[i + 1] for i in range(nq - 1))).tocsc().astype(complex)
    Ls = [sps.diags(zq).tocsc().astype(complex) for zq in z]
    return Hx, Hz, Ls


def test_c4_trajectories_4096(Q):
    rng = np.random.default_rng(4000)
    nq, n, nb = 12, 4096, 512
    Hx, Hz, Ls = tfim_sparse(nq)
    fs = [lambda s: 1 - s, lambda s: s]
    Hg, Ho = Q.SparseHamiltonian(fs, [Hx, Hz], unit="ħ"), O.SparseHamiltonian(fs, [Hx, Hz], unit="ħ")
    lg = Q.LindbladLiouvillian([Q.Lindblad(0.05, L) for L in Ls])
    opg = Q.DiffEqLiouvillian(Hg, [], [lg], n)
    ens = Q.TrajectoryEnsemble(opg)
    psi = rand_c(rng, nb, n)
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    tf = 10.0
    t = rng.random(nb) * tf  # per-trajectory s_b
    pg = Q.ODEParams(None, tf)
    pd = Q.DeviceBatch.from_host(psi)
    dpsi = ens.matvec(pd.zeros_like(), pd, pg, t).to_host()
    # oracle on a few trajectories: cache = −iH(s) − ½ΣγL'L  (sparse on the CPU)
    for b in (0, 17, 511):
        s = t[b] / tf
        cache = Ho.update_cache(s) - 0.5 * sum(0.05 * (L.conj().T @ L) for L in Ls)
        assert relerr(dpsi[b], cache @ psi[b]) < TOL
    # norm-decay identity: 2 Re⟨ψ|dψ⟩ = −Σ_k γ_k‖L_kψ‖²  links K7 (matvec) to K8 (jump probabilities)
    r = rng.random(nb)
    choice, prob = ens.lindblad_jump(pd, pg, t, r)
    lhs = 2 * np.real(np.sum(psi.conj() * dpsi, axis=1))
    assert np.allclose(lhs, -prob.cpu().numpy().sum(axis=1), rtol=1e-11, atol=1e-13)
    # jump indices reproduce the inverse-CDF scan on the probabilities
    pr = prob.cpu().numpy()
    exp_choice = np.array([O.sample_weights(list(pr[b]), r[b]) for b in range(nb)])
    assert np.array_equal(choice.cpu().numpy(), exp_choice)
    # host-buffer entry point
    ph = np.ascontiguousarray(psi)
    dh = np.zeros_like(ph)
    ens.matvec_host(dh, ph, pg, t)
    assert np.array_equal(dh, dpsi)


# ---- C5: 7-qubit Redfield apply with supplied Λ, n = 128 -------------------------------------------
def test_c5_redfield_128(Q):
    rng = np.random.default_rng(5000)
    nq, n, K, nb = 7, 128, 7, 16
    Hd, Hp, z = tfim_dense(nq)
    S = [np.diag(zq).astype(complex) for zq in z]
    fs = [lambda s: 1 - s, lambda s: s]
    Hg, Ho = Q.DenseHamiltonian(fs, [Hd, Hp], unit="ħ"), O.DenseHamiltonian(fs, [Hd, Hp], unit="ħ")
    Lam = rand_c(rng, nb, K, n, n)
    Lam /= np.linalg.norm(Lam, axis=(2, 3), keepdims=True)
    u = np.stack([lowrank_rho(rng, n) for _ in range(nb)])
    pw = 0.5 + 1.5 * np.arange(nb) / nb  # per-schedule s_b(t) = (t/tf)^{p_b}
    s = 0.37 ** pw
    du = Hg(np.zeros_like(u), u, None, s)  # commutator, per-schedule s
    rg = Q.RedfieldLiouvillian(None, None, 0, 0, 0)
    du = rg.apply(du, u, S, Lam)
    for b in (0, 7, 15):
        ref = O.redfield_apply_acc(Ho.rhs(u[b], s[b]), u[b], S, list(Lam[b]))
        assert relerr(du[b], ref) < TOL
    for b in range(nb):
        assert abs(np.trace(du[b])) < 1e-12 * np.linalg.norm(du[b]) * n


# ---- full-size cross-route checks of the round-2 fast paths --------------------------------------------------------
def test_c4_unstructured_4096_sell_vs_plain_and_scipy(Q):
    """An operator WITHOUT Pauli structure at the C4 dimension (n = 4096, ≈13 random entries per row, complex values) over
    8192 trajectories with per-trajectory s: the sliced-ELL kernel against the plain ELL kernel on the whole ensemble (1e-13),
    against scipy on sampled members (1e-12), and linear in ψ."""
    import ctypes as C
    import torch
    nq, n, nb = 12, 4096, 8192
    rng = np.random.default_rng(4100)
    rows = np.repeat(np.arange(n), 6)
    cols = rng.integers(0, n, size=n * 6)
    keep = rows != cols
    vals = rng.standard_normal(n * 6) + 1j * rng.standard_normal(n * 6)
    A = sps.csc_matrix((vals[keep], (rows[keep], cols[keep])), shape=(n, n))
    H1 = (A + A.conj().T).tocsc()
    H2 = sps.diags(rng.standard_normal(n)).tocsc().astype(complex)
    _, _, Ls = tfim_sparse(nq)
    H = Q.SparseHamiltonian([lambda s: 1 - s, lambda s: s], [H1, H2], unit="ħ")
    ens = Q.TrajectoryEnsemble(Q.DiffEqLiouvillian(H, [], [Q.LindbladLiouvillian([Q.Lindblad(0.05, L) for L in Ls])], n))
    ctx = Q.get_context()
    gen = torch.Generator(device="cuda").manual_seed(41)
    psi_t = torch.randn((nb, 1, n), dtype=torch.complex128, device="cuda", generator=gen)
    psi = Q.DeviceBatch(nb, n, 1, ctx, psi_t.contiguous())
    s = rng.random(nb)
    cf = np.ascontiguousarray(np.stack([1 - s, s], axis=1))
    gam = np.full((1, nq), 0.05)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))

    def matvec(x):
        out = x.zeros_like()
        ctx.check(ctx.lib.oqb_traj_matvec(ens.sp, nb, dp(cf), 0, dp(gam), 1, x.ptr, out.ptr))
        return out

    sell = matvec(psi)
    pl, pad = C.c_int(), C.c_double()
    ctx.check(ctx.lib.oqb_traj_sell_stats(ens.sp, C.byref(pl), C.byref(pad)))
    assert pl.value == 1 and 1.0 < pad.value < 1.3  # the sliced-ELL plan was built and used
    with ctx.option("traj_ell", 0):
        plain = matvec(psi)
    assert float(torch.linalg.vector_norm(sell.t - plain.t) / torch.linalg.vector_norm(plain.t)) < 1e-13
    hs = sell.t.reshape(nb, n)
    for b in (0, 4097, nb - 1):
        M = -1j * ((1 - s[b]) * H1 + s[b] * H2) - 0.5 * 0.05 * nq * sps.identity(n)
        assert relerr(hs[b].cpu().numpy(), M @ psi_t[b, 0].cpu().numpy()) < TOL
    # linearity: A(2ψ − iψ') = 2Aψ − iAψ'
    psi2_t = torch.randn((nb, 1, n), dtype=torch.complex128, device="cuda", generator=gen)
    comb = Q.DeviceBatch(nb, n, 1, ctx, (2 * psi_t - 1j * psi2_t).contiguous())
    lhs = matvec(comb).t
    rhs = 2 * sell.t - 1j * matvec(Q.DeviceBatch(nb, n, 1, ctx, psi2_t.contiguous())).t
    assert float(torch.linalg.vector_norm(lhs - rhs) / torch.linalg.vector_norm(rhs)) < 1e-13


def test_c2_full_batch_split_route_vs_assembled(Q):
    """C2 at its full batch (4096 ρ of 256×256, per-member s, dephasing): the split-Hamiltonian route (shared driver operand
    scaled per member, diagonal terms and the Lindblad factor in one Hadamard pass) against the assembled-H_eff route on
    every member (1e-13), the oracle on sampled members (1e-12), trace and Hermiticity on all."""
    import torch
    rng = np.random.default_rng(2001)
    nq, n, K, nb = 8, 256, 8, 4096
    Hd, Hp, z = tfim_dense(nq)
    Ls = [np.diag(zq).astype(complex) for zq in z]
    fs = [lambda s: 1 - s, lambda s: s]
    gam = [0.1 * (1 + k / 8) for k in range(K)]
    opg = Q.DiffEqLiouvillian(Q.DenseHamiltonian(fs, [Hd, Hp], unit="ħ"), [], [Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gam, Ls)])], n)
    opo = O.DiffEqLiouvillian(O.DenseHamiltonian(fs, [Hd, Hp], unit="ħ"), [], [O.LindbladLiouvillian([O.Lindblad(g, L) for g, L in zip(gam, Ls)])], n)
    ctx = Q.get_context()
    gen = torch.Generator(device="cuda").manual_seed(7)
    g = torch.randn((nb, 16, n), dtype=torch.complex128, device="cuda", generator=gen)
    rho = g.transpose(1, 2) @ g.conj()
    rho = rho / torch.diagonal(rho, dim1=1, dim2=2).sum(-1).real[:, None, None]
    u = Q.DeviceBatch(nb, n, n, ctx, rho.contiguous())
    t = np.arange(nb) / nb * 10.0
    p = Q.ODEParams(None, 10.0)
    du = opg(u.zeros_like(), u, p, t)
    with ctx.option("lindblad_split", 0), ctx.option("lindblad_fuse", 0):
        du0 = opg(u.zeros_like(), u, p, t)
    assert float(torch.linalg.vector_norm(du.t - du0.t) / torch.linalg.vector_norm(du0.t)) < 1e-13
    # DeviceBatch blocks are column-major (Julia layout): block[b] read row-major is the transpose
    for b in (0, 1234, nb - 1):
        ub = u.t[b].cpu().numpy().T
        assert relerr(du.t[b].cpu().numpy().T, opo.rhs(ub, O.ODEParams(None, 10.0), t[b])) < TOL
    tr = torch.diagonal(du.t, dim1=1, dim2=2).sum(-1)
    nrm = torch.linalg.matrix_norm(du.t)
    assert float((tr.abs() / (nrm * n)).max()) < 1e-12
    assert float(torch.linalg.vector_norm(du.t - du.t.transpose(1, 2).conj()) / torch.linalg.vector_norm(du.t)) < 1e-12
