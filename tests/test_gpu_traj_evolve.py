"""On-device trajectory stepping (oqb_traj_evolve) against the oracle's restatement of the external
solver's loop: RK4 of dψ/dt = cache(t)ψ (cache of src/opensys/diffeq_liouvillian.jl:78-83) + the jump callback
(lindblad_jump, src/opensys/trajectory_jump.jl:2-12,71-74).  Same xoshiro256++ streams on both sides: jump steps and
operator indices must be IDENTICAL, ψ within 1e-12 (relative, Frobenius)."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sps

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures as F  # noqa: E402
from oracle import oqb_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
A = lambda s: 1 - s
B = lambda s: s


@pytest.fixture(scope="module")
def Q():
    import oqb_b200
    return oqb_b200


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


def tfim_sparse(nq):
    Hx = sps.csc_matrix(-F.standard_driver(nq))
    Hz = sps.csc_matrix(-F.two_local_term([1.0] * (nq - 1), [[i, i + 1] for i in range(1, nq)], nq))
    Ls = [sps.csc_matrix(F.single_clause(["z"], [k], 1.0, nq)) for k in range(1, nq + 1)]
    return Hx, Hz, Ls


def test_rng_streams_match_oracle(Q):
    import torch
    Hx, Hz, Ls = tfim_sparse(2)
    H = Q.SparseHamiltonian([A, B], [Hx, Hz], unit="ħ")
    lg = Q.LindbladLiouvillian([Q.Lindblad(lambda s: 0.1, L) for L in Ls])
    ens = Q.TrajectoryEnsemble(Q.DiffEqLiouvillian(H, [], [lg], 4))
    nb, seed = 37, 12345
    st = Q.TrajectoryStepper(ens, nb, seed=seed)
    ens.ctx.sync()
    first = st.thr.cpu().numpy()
    out = torch.zeros(nb, dtype=torch.float64, device=st.thr.device)
    import ctypes as C
    ens.ctx.check(ens.ctx.lib.oqb_traj_rng_uniform(ens.sp, nb, C.c_void_p(st.rng.data_ptr()), C.c_void_p(out.data_ptr())))
    ens.ctx.sync()
    second = out.cpu().numpy()
    for b in range(nb):
        r = O.Xoshiro256pp.from_seed(seed, b)
        assert r.rand() == first[b] and r.rand() == second[b]
        assert [int(x) for x in st.rng[b].cpu().numpy().view(np.uint64)] == r.s
    # known-answer: xoshiro256++ reference vector from state (1,2,3,4): first outputs 41943041, 58720359, 3588806011781223
    r = O.Xoshiro256pp([1, 2, 3, 4])
    assert [r.next_u64() for _ in range(3)] == [41943041, 58720359, 3588806011781223]


# n ≥ 256 (nq ≥ 8) runs the XOR-structured matvec with the RK4 stage arithmetic fused into its store
@pytest.mark.parametrize("nq,nb,per_member", [(3, 6, False), (5, 8, True), (7, 5, False), (8, 5, False), (9, 4, True)])
def test_evolve_matches_oracle(Q, nq, nb, per_member):
    rng = np.random.default_rng(40 + nq)
    n = 2 ** nq
    Hx, Hz, Ls = tfim_sparse(nq)
    gfs = [(lambda s, k=k: 0.6 * (1 + 0.3 * k + 0.5 * s)) for k in range(nq)]  # strong damping: several jumps per run
    Hg = Q.SparseHamiltonian([A, B], [Hx, Hz], unit="ħ")
    Ho = O.SparseHamiltonian([A, B], [Hx, Hz], unit="ħ")
    lg = Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gfs, Ls)])
    lo = O.LindbladLiouvillian([O.Lindblad(g, L.toarray()) for g, L in zip(gfs, Ls)])
    opg = Q.DiffEqLiouvillian(Hg, [], [lg], n)
    opo = O.DiffEqLiouvillian(Ho, [], [lo], n)
    ens = Q.TrajectoryEnsemble(opg)
    tf, dt, nsteps, seed = 4.0, 0.02, (120 if nq < 8 else 60), 77  # the dense oracle is slow at n ≥ 256
    if per_member:
        pws = 0.5 + 1.5 * rng.random(nb)
        pg = [Q.ODEParams(None, tf, lambda tf_, t, q=q: (t / tf_) ** q) for q in pws]
        po = [O.ODEParams(None, tf, lambda tf_, t, q=q: (t / tf_) ** q) for q in pws]
    else:
        pg, po = Q.ODEParams(None, tf), [O.ODEParams(None, tf)] * nb
    psi0 = rng.standard_normal((nb, n)) + 1j * rng.standard_normal((nb, n))
    psi0 /= np.linalg.norm(psi0, axis=1, keepdims=True)
    pd = Q.DeviceBatch.from_host(psi0)
    st = Q.TrajectoryStepper(ens, nb, seed=seed, max_log=64)
    # two calls: the stepper carries RNG state, thresholds, counters and the step index across calls
    st.evolve(pd, pg, 0.0, dt, 25)
    st.evolve(pd, pg, 25 * dt, dt, nsteps - 25)
    got = pd.to_host()
    nj = st.njumps.cpu().numpy()
    log = st.log.cpu().numpy()
    thr = st.thr.cpu().numpy()
    total_jumps = 0
    Ld = [L.toarray() for L in Ls]
    for b in range(nb):
        r = O.Xoshiro256pp.from_seed(seed, b)
        thr0 = r.rand()
        ref, thr_end, lg_ref = O.traj_evolve_rk4(lambda t: opo.update_cache(po[b], t), lambda t: Ld,
                                                 lambda t: [g(po[b](t)) for g in gfs], psi0[b], r, thr0, 0.0, dt, nsteps)
        assert nj[b] == len(lg_ref)
        assert [tuple(x) for x in log[b, :nj[b]]] == lg_ref      # identical jump steps and operator indices
        assert thr[b] == thr_end                                   # identical RNG consumption
        assert relerr(got[b], ref) < 1e-12
        total_jumps += len(lg_ref)
    assert total_jumps > 0


# ---- general Pauli-string SparseHamiltonians through the implicit-column (signed XOR) trajectory kernel ----------------
def _pauli_sum(nq, strings):
    """Σ c·P for Pauli strings given as (coefficient, {qubit: 'X'|'Y'|'Z'}) — qubit 1 is the most significant bit."""
    import scipy.sparse as sps
    P = {"X": np.array([[0, 1], [1, 0]], dtype=complex), "Y": np.array([[0, -1j], [1j, 0]]), "Z": np.array([[1, 0], [0, -1]], dtype=complex),
         "I": np.eye(2, dtype=complex)}
    n = 2 ** nq
    H = sps.csc_matrix((n, n), dtype=complex)
    for c, ops in strings:
        m = sps.identity(1, dtype=complex, format="csc")
        for q in range(1, nq + 1):
            m = sps.kron(m, sps.csc_matrix(P[ops.get(q, "I")]), format="csc")
        H = H + c * m
    H = sps.csc_matrix(H)
    H.eliminate_zeros()
    return H


@pytest.mark.parametrize("nq", [8, 10, 12])
@pytest.mark.parametrize("model", ["heisenberg_xyz", "y_field_and_xz", "generic_kernel"])
def test_pauli_string_hamiltonians_matvec(Q, nq, model):
    """SparseHamiltonians that are sums of general Pauli strings — Y fields (imaginary, row-signed entries), XX + YY + ZZ
    bonds (two strings sharing their X-type factors: half-empty patterns), X·Z products — take the implicit-column kernel
    with per-slot sign masks M[r, r^m] = v·(−1)^popcount(r & z); `generic_kernel` runs the same operators through the generic
    ELL kernels (option traj_kernel = 1).  Both against the oracle's assembled −iH(s) − ½ΣγL'L."""
    import scipy.sparse as sps
    rng = np.random.default_rng(nq * 7 + len(model))
    n, nb = 2 ** nq, 6
    if model == "y_field_and_xz":
        H1 = _pauli_sum(nq, [(-1.0 - 0.1 * q, {q: "Y"}) for q in range(1, nq + 1)])
        H2 = _pauli_sum(nq, [(0.7, {q: "X", q + 1: "Z"}) for q in range(1, nq)] + [(0.3 * (-1) ** q, {q: "Z", q + 1: "Z"}) for q in range(1, nq)])
    else:
        H1 = _pauli_sum(nq, [(0.5, {q: "X", q + 1: "X"}) for q in range(1, nq)] + [(0.5, {q: "Y", q + 1: "Y"}) for q in range(1, nq)])
        H2 = _pauli_sum(nq, [(0.8, {q: "Z", q + 1: "Z"}) for q in range(1, nq)] + [(0.2, {q: "Z"}) for q in range(1, nq + 1)])
    Ls = [_pauli_sum(nq, [(1.0, {q: "Z"})]) for q in range(1, min(nq, 6) + 1)]
    fs = [lambda s: 1 - s, lambda s: s]
    Hg, Ho = Q.SparseHamiltonian(fs, [H1, H2], unit="ħ"), O.SparseHamiltonian(fs, [H1, H2], unit="ħ")
    gfs = [lambda s, k=k: 0.05 * (1 + k) for k in range(len(Ls))]
    opg = Q.DiffEqLiouvillian(Hg, [], [Q.LindbladLiouvillian([Q.Lindblad(g, L) for g, L in zip(gfs, Ls)])], n)
    ens = Q.TrajectoryEnsemble(opg)
    psi = rng.standard_normal((nb, n)) + 1j * rng.standard_normal((nb, n))
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    t = rng.random(nb) * 3.0
    pd = Q.DeviceBatch.from_host(psi)
    ctx = Q.get_context()
    with ctx.option("traj_kernel", 1 if model == "generic_kernel" else 0):
        l0 = ctx.launch_count()
        dpsi = ens.matvec(pd.zeros_like(), pd, Q.ODEParams(None, 3.0), t).to_host()
    ref = []
    for b in range(nb):
        s = t[b] / 3.0
        M = -1j * ((1 - s) * H1 + s * H2)
        for g, L in zip(gfs, Ls):
            M = M - 0.5 * g(s) * (L.conj().T @ L)
        ref.append(M @ psi[b])
    assert relerr(dpsi, np.stack(ref)) < 1e-12
