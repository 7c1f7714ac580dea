"""Error behaviour and edge cases of the C ABI (include/oqb.h): every entry point returns a status and leaves a message
in oqb_last_error — nothing throws, nothing crashes; empty batches are no-ops.  Mirrors the reference's argument
checks (ArgumentError at src/opensys/lindblad.jl:88-90, src/coupling/interaction.jl:45-47)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Q():
    import oqb_b200
    return oqb_b200


def test_status_codes_and_messages(Q):
    ctx = Q.get_context()
    lib = ctx.lib
    n = 4
    m = np.ascontiguousarray(np.eye(n, dtype=complex)[None])
    h = C.c_void_p()
    assert lib.oqb_dense_create(ctx.h, n, 1, m.ctypes.data, 0, None, C.byref(h)) == 0
    u = Q.DeviceBatch(2, n, n, ctx)
    one = np.ones((1, 1))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    # empty batch: no-op, success
    assert lib.oqb_dense_rhs(h, 0, dp(one), 1, None, 1, u.ptr, u.ptr) == 0
    # null state pointer → EINVAL with a message, context stays usable
    st = lib.oqb_dense_rhs(h, 2, dp(one), 1, None, 1, None, u.ptr)
    assert st == 1 and b"null" in lib.oqb_last_error(ctx.h)
    assert lib.oqb_dense_rhs(h, 2, dp(one), 1, None, 1, u.ptr, u.zeros_like().ptr) == 0
    lib.oqb_dense_destroy(h)
    # AME: RHS before the eigen-system was set
    a = C.c_void_p()
    assert lib.oqb_ame_create(ctx.h, n, n, 1, m.ctypes.data, C.byref(a)) == 0
    assert lib.oqb_ame_rhs(a, 2, u.ptr, u.ptr) == 1 and b"set_eigen" in lib.oqb_last_error(ctx.h)
    w = np.arange(n, dtype=float)
    assert lib.oqb_ame_jump(a, 2, u.ptr, u.ptr, None, None, None, 0) == 1
    assert lib.oqb_ame_set_eigen(a, dp(w), None) == 0
    lib.oqb_ame_destroy(a)
    # zgemm mode
    assert lib.oqb_set_zgemm_mode(ctx.h, 7) == 1
    assert lib.oqb_set_zgemm_mode(ctx.h, 1) == 0
    # Λ builder: evaluation before the unitary was set, node outside the sampled grid
    lam = C.c_void_p()
    S = np.ascontiguousarray(np.diag([1.0, -1, 1, -1]).astype(complex)[None])
    assert lib.oqb_lambda_create(ctx.h, n, 1, S.ctypes.data, C.byref(lam)) == 0
    i32 = lambda v: v.ctypes.data_as(C.POINTER(C.c_int32))
    mem, cp = np.zeros(1, np.int32), np.zeros(1, np.int32)
    gi, th = np.zeros((1, 16), np.int32), np.zeros((1, 16))
    ck, cg = np.zeros((1, 15, 2)), np.zeros((1, 7, 2))
    slot, E = np.zeros(1, np.int64), np.zeros(1)
    call = lambda: lib.oqb_lambda_eval(lam, 1, i32(mem), i32(cp), i32(gi), dp(th), dp(ck), dp(cg),
                                       slot.ctypes.data_as(C.POINTER(C.c_int64)), dp(E))
    assert call() == 1 and b"set_unitary" in lib.oqb_last_error(ctx.h)
    U = Q.DeviceBatch.from_host(np.stack([np.eye(n, dtype=complex)] * 3), ctx)
    assert lib.oqb_lambda_set_unitary(lam, 1, 3, U.ptr) == 0
    assert lib.oqb_lambda_reserve_slots(lam, 4) == 0
    gi[0, 3] = 5
    assert call() == 1 and b"outside" in lib.oqb_last_error(ctx.h)
    gi[0, 3] = 0
    assert call() == 0 and E[0] == 0.0
    lib.oqb_lambda_destroy(lam)


def test_host_mirror_argument_errors(Q):
    sz = np.diag([1.0, -1]).astype(complex)
    with pytest.raises(ValueError):                      # src/coupling/interaction.jl:45-47
        Q.Lindblad(lambda s: np.eye(2), sz)
    with pytest.raises(ValueError):                      # src/opensys/lindblad.jl:88-90
        Q.LindbladLiouvillian([Q.Lindblad(0.1, sz), Q.Lindblad(0.1, np.eye(4, dtype=complex))])
    H = Q.DenseHamiltonian([lambda s: 1.0], [np.kron(sz, sz)])
    op = Q.DiffEqLiouvillian(H, [Q.DaviesGenerator([np.kron(sz, np.eye(2))], lambda w: 1.0, lambda w: 0.0)], [], 4)
    u = np.stack([np.eye(4, dtype=complex) / 4] * 2)
    w4, v4 = np.linalg.eigh(np.kron(sz, sz))
    with pytest.raises(ValueError):                      # a supplied (w, v) belongs to ONE s
        op.rhs_with_eig(np.zeros_like(u), u, Q.ODEParams(H, 1.0), np.array([0.1, 0.9]), w4, v4)


def test_plain_c_host_program(c_demo_exe):
    """examples/c_abi_demo.c — a host program in C calling the ABI directly — reproduces the reference's Lindblad and
    commutator known answers (test/opensys/lindblad.jl:25-31, test/hamiltonian/dense_hamiltonian.jl:35-38)."""
    import subprocess
    out = subprocess.run([c_demo_exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "kernels launched" in out.stdout


def test_plain_c_ame_end_to_end(c_liouv_exe):
    """examples/c_liouv_demo.c — a 6-qubit C3-shaped AME (gap coincidences, degenerate start) through C only: H(s) assembly,
    device eigensolver, GapIndices, Davies tables and rotations behind oqb_liouv_rhs_host, against a golden file written
    from the oracle's literal gap loop (tests/golden/make_liouv_golden.py)."""
    import os
    import subprocess
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "liouv_c_demo.bin")
    out = subprocess.run([c_liouv_exe, golden], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "c_liouv_demo OK" in out.stdout


def test_plain_c_nccl_reduce(c_comm_exe):
    """examples/c_comm_demo.c — oqb_comm_init_all + oqb_expect_sum + oqb_allreduce_obs from one C process with one context
    (and one host thread) per visible GPU; with a single GPU the communicator has one rank and the reduce is the identity."""
    import subprocess
    out = subprocess.run([c_comm_exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "c_comm_demo OK" in out.stdout


def test_round2_entry_points_reject_bad_arguments(Q):
    """Argument and ordering checks of the round-2 entry points: options, GapIndices / term builder, the Liouvillian object,
    resample, the Λ integrator and the communicator calls return a status and a message instead of crashing."""
    ctx = Q.get_context()
    lib = ctx.lib
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    assert lib.oqb_set_option(ctx.h, b"no_such_option", 1) == 1 and b"unknown option" in lib.oqb_last_error(ctx.h)
    v = C.c_int64()
    assert lib.oqb_set_option(ctx.h, b"davies_dense_min_pairs", 5) == 0 and lib.oqb_get_option(ctx.h, b"davies_dense_min_pairs", C.byref(v)) == 0 and v.value == 5
    assert lib.oqb_set_option(ctx.h, b"davies_dense_min_pairs", 0) == 0
    n = 4
    m = np.ascontiguousarray(np.diag([1.0, -1, 1, -1]).astype(complex)[None])
    a = C.c_void_p()
    assert lib.oqb_ame_create(ctx.h, n, n, 1, m.ctypes.data, C.byref(a)) == 0
    nk, nz = C.c_int64(), C.c_int64()
    assert lib.oqb_ame_gaps_build(a, 8, 8, C.byref(nk), C.byref(nz)) == 1 and b"set_eigen" in lib.oqb_last_error(ctx.h)
    assert lib.oqb_ame_add_davies_ohmic(a, 0, 1, 1e-4, 25.0, 0.5, 0, 0.0, 1.0, None) == 1 and b"gaps_build" in lib.oqb_last_error(ctx.h)
    w = np.array([0.0, 1.0, 1.0, 2.5])
    assert lib.oqb_ame_set_eigen(a, dp(w), None) == 0
    assert lib.oqb_ame_gaps_build(a, 8, 8, C.byref(nk), C.byref(nz)) == 0 and nk.value == 3 and nz.value == 1   # gaps 1, 1.5, 2.5; one degenerate pair
    assert lib.oqb_ame_add_davies_ohmic(a, 0, 2, 1e-4, 25.0, 0.5, 0, 0.0, 1.0, None) == 1 and b"coupling slice" in lib.oqb_last_error(ctx.h)
    assert lib.oqb_ame_add_davies(a, 0, 1, None, None, None, None, 0.0, 0.0) == 1
    assert lib.oqb_ame_add_correlated(a, 0, 3, None, None, None, None, 0.0, 0.0) == 1 and b"out of range" in lib.oqb_last_error(ctx.h)
    assert lib.oqb_ame_add_davies_ohmic(a, 0, 1, 1e-4, 25.0, 0.5, 0, 0.0, 1.0, None) == 0
    st = [C.c_int64() for _ in range(5)]
    assert lib.oqb_ame_davies_stats(a, *[C.byref(x) for x in st]) == 0 and st[0].value == 1
    cache = Q.DeviceBatch(1, n, n)
    assert lib.oqb_ame_update_cache(a, None) == 1
    assert lib.oqb_ame_update_cache(a, cache.ptr) == 0
    lib.oqb_ame_destroy(a)
    # the Liouvillian object
    L = C.c_void_p()
    assert lib.oqb_liouv_create(ctx.h, None, n, 8, 8, 1, m.ctypes.data, C.byref(L)) == 1 and b"null Hamiltonian" in lib.oqb_last_error(ctx.h)
    hterms = np.ascontiguousarray(np.stack([np.diag([0.0, 1, 1, 2.5]).astype(complex)]))
    h = C.c_void_p()
    assert lib.oqb_dense_create(ctx.h, n, 1, hterms.ctypes.data, 0, None, C.byref(h)) == 0
    assert lib.oqb_liouv_create(ctx.h, h, n, 8, 8, 1, m.ctypes.data, C.byref(L)) == 0
    assert lib.oqb_liouv_add_davies_ohmic(L, 1, 1, 1e-4, 25.0, 0.5, 0, 0.0, 1.0, None) == 1 and b"coupling slice" in lib.oqb_last_error(ctx.h)
    assert lib.oqb_liouv_add_davies_fn(L, 0, 1, None, None, None, None) == 1
    assert lib.oqb_liouv_set_plan_cache(L, 0) == 1
    assert lib.oqb_liouv_add_davies_ohmic(L, 0, 1, 1e-4, 25.0, 0.5, 0, 0.0, 1.0, None) == 0
    u = Q.DeviceBatch(2, n, n)
    one = np.ones((1, 1))
    assert lib.oqb_liouv_rhs(L, 2, None, 1, u.ptr, u.ptr) == 1
    assert lib.oqb_liouv_rhs(L, 2, dp(one), 1, u.ptr, u.zeros_like().ptr) == 0
    pb, ch = C.c_uint64(), C.c_uint64()
    assert lib.oqb_liouv_stats(L, C.byref(pb), C.byref(ch), None, None) == 0 and pb.value == 1
    assert lib.oqb_liouv_destroy(L) == 0 and lib.oqb_dense_destroy(h) == 0
    # resample / communicator
    assert lib.oqb_resample(ctx.h, 4, 0, None, None, None, None) == 1
    assert lib.oqb_allreduce_obs(None, None, 1) == 1 and lib.oqb_comm_init_rank(ctx.h, 2, 5, None, C.byref(L)) != 0


def test_set_davies_with_host_tables_and_lists(Q):
    """oqb_ame_set_davies — l×l γ/S tables at the signed rounded gaps and explicit cross-term / L'L lists, everything evaluated by
    the host (here: the test-side `gaps.GapGroups` statement of GapIndices) — against the oracle's literal gap loop on a
    spectrum with degenerate levels and equal spacings."""
    from oqb_b200 import _lib
    from oqb_b200.gaps import GapGroups
    from oracle import oqb_oracle as O
    rng = np.random.default_rng(17)
    l, nb = 7, 3
    w = np.array([-2.0, -2.0, -0.5, 0.25, 1.0, 1.75, 1.75])
    gam_f = lambda x: x + 1 if x >= 0 else (1 - x) * np.exp(x)
    lam_f = lambda x: x + 0.1
    cs = [(lambda z: z + z.conj().T)(rng.standard_normal((l, l)) + 1j * rng.standard_normal((l, l))) for _ in range(2)]
    ctx = Q.get_context()
    lib = ctx.lib
    a = C.c_void_p()
    cbuf = np.ascontiguousarray(np.stack([c.T for c in cs]))
    ctx.check(lib.oqb_ame_create(ctx.h, l, l, 2, cbuf.ctypes.data, C.byref(a)))
    ctx.check(lib.oqb_ame_set_eigen(a, _lib.dptr(np.ascontiguousarray(w)), None))
    G = GapGroups(w, 8, 8)
    gam, Sm, crate, lrs = G.tables(gam_f, lam_f)
    gam, Sm = np.ascontiguousarray(gam.T), np.ascontiguousarray(Sm.T)
    ci, li = np.ascontiguousarray(G.cross_idx), np.ascontiguousarray(G.lamb_idx)
    ctx.check(lib.oqb_ame_set_davies(a, _lib.dptr(gam), _lib.dptr(Sm), len(ci), ci.ctypes.data_as(_lib.c_i32p), _lib.dptr(crate),
                                     len(li), li.ctypes.data_as(_lib.c_i32p), _lib.dptr(np.ascontiguousarray(lrs))))
    rho = rng.standard_normal((nb, l, l)) + 1j * rng.standard_normal((nb, l, l))
    u, du = Q.DeviceBatch.from_host(rho), Q.DeviceBatch.from_host(np.zeros_like(rho))
    ctx.check(lib.oqb_ame_eig_acc(a, nb, u.ptr, du.ptr, 0))
    ctx.sync()
    gi = O.build_gap_indices(w, 8, 8)
    ref = np.stack([O.DaviesGenerator(cs, gam_f, lam_f).rhs_acc(np.zeros((l, l), complex), rho[b], gi, None, 0.0) for b in range(nb)])
    got = du.to_host()
    assert np.linalg.norm(got - ref) < 1e-12 * np.linalg.norm(ref)
    lib.oqb_ame_destroy(a)


def test_bath_and_option_entry_points_reject_bad_arguments(Q):
    """Argument checks of the bath routines, the spline table and the handle/option calls added for rows B and (f)-3."""
    ctx = Q.get_context()
    lib = ctx.lib
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    w = np.array([0.0, 1.0])
    S = np.zeros(2)
    # empty request: no-op
    assert lib.oqb_ohmic_lambshift(ctx.h, 1e-4, 25.0, 0.5, 0, w.ctypes.data, 1e-8, 0.0, 64, S.ctypes.data, None, None) == 0
    # null output / non-positive cutoff / one-segment heap
    assert lib.oqb_ohmic_lambshift(ctx.h, 1e-4, 25.0, 0.5, 2, w.ctypes.data, 1e-8, 0.0, 64, None, None, None) == 1
    assert b"null" in lib.oqb_last_error(ctx.h)
    assert lib.oqb_ohmic_lambshift(ctx.h, 1e-4, -1.0, 0.5, 2, w.ctypes.data, 1e-8, 0.0, 64, S.ctypes.data, None, None) == 1
    assert lib.oqb_ohmic_lambshift(ctx.h, 1e-4, 25.0, 0.5, 2, w.ctypes.data, 1e-8, 0.0, 1, S.ctypes.data, None, None) == 1
    # a segment cap that is too small for the tolerance still returns (the error estimate says it did not converge)
    err = np.zeros(2)
    nseg = np.zeros(2, dtype=np.int32)
    assert lib.oqb_ohmic_lambshift(ctx.h, 1e-4, 25.0, 0.5, 2, w.ctypes.data, 1e-14, 0.0, 4, S.ctypes.data, err.ctypes.data,
                                   nseg.ctypes.data) == 0
    assert np.all(nseg <= 4) and np.all(np.isfinite(S)) and np.all(err > 1e-14 * np.abs(S))
    Cc = np.zeros(2, dtype=complex)
    assert lib.oqb_ohmic_correlation(ctx.h, 1e-4, 25.0, 0.5, 2, None, Cc.ctypes.data) == 1
    assert lib.oqb_ohmic_correlation(ctx.h, 1e-4, 25.0, 0.5, 0, None, None) == 0
    out = [C.c_double() for _ in range(4)]
    assert lib.oqb_ohmic_timescales(ctx.h, 1e-4, 25.0, 0.5, float("inf"), 1e-8, 0.0, 256, *[C.byref(x) for x in out]) == 1
    assert lib.oqb_ohmic_timescales(ctx.h, 1e-4, 25.0, 0.5, 10.0, 1e-8, 0.0, 256, None, None, None, None) == 1
    coef = np.zeros(5)
    x = np.zeros(3)
    assert lib.oqb_bspline_eval(ctx.h, 3, 0.0, -1.0, coef.ctypes.data, 3, x.ctypes.data, x.ctypes.data) == 1
    assert lib.oqb_bspline_eval(ctx.h, 3, 0.0, 1.0, coef.ctypes.data, 0, None, None) == 0
    # handle-level calls
    n = 4
    m = np.ascontiguousarray(np.eye(n, dtype=complex)[None])
    a = C.c_void_p()
    assert lib.oqb_ame_create(ctx.h, n, n, 1, m.ctypes.data, C.byref(a)) == 0
    assert lib.oqb_ame_rebind(a, None) == 1 and lib.oqb_ame_rebind(None, ctx.h) == 1
    assert lib.oqb_ame_rebind(a, ctx.h) == 0
    assert lib.oqb_ame_set_lambshift_spline(a, 3, 0.0, 0.0, coef.ctypes.data) == 1
    assert lib.oqb_ame_set_lambshift_spline(a, 3, 0.0, 1.0, None) == 1
    assert lib.oqb_ame_set_lambshift_spline(a, 0, 0.0, 1.0, None) == 0
    i32 = np.zeros(1, dtype=np.int32)
    ip = i32.ctypes.data_as(C.POINTER(C.c_int32))
    assert lib.oqb_ame_set_onesided_ohmic(a, 1, ip, ip, 1e-4, 25.0, 0.5) == 1  # no eigen-system yet
    assert b"set_eigen" in lib.oqb_last_error(ctx.h)
    flag = C.c_int(7)
    assert lib.oqb_ame_eigenvectors_real(a, C.byref(flag)) == 0 and flag.value == 0
    assert lib.oqb_ame_eigenvectors_real(a, None) == 1
    wv = np.arange(n, dtype=float)
    V = np.ascontiguousarray(np.eye(n, dtype=complex))
    assert lib.oqb_ame_set_eigen(a, dp(wv), V.ctypes.data) == 0
    assert lib.oqb_ame_eigenvectors_real(a, C.byref(flag)) == 0 and flag.value == 1
    V[0, 0] = 1j
    assert lib.oqb_ame_set_eigen(a, dp(wv), V.ctypes.data) == 0
    assert lib.oqb_ame_eigenvectors_real(a, C.byref(flag)) == 0 and flag.value == 0
    bad = np.array([5], dtype=np.int32)
    assert lib.oqb_ame_set_onesided_ohmic(a, 1, bad.ctypes.data_as(C.POINTER(C.c_int32)), ip, 1e-4, 25.0, 0.5) == 1
    lib.oqb_ame_destroy(a)
    on = C.c_int()
    assert lib.oqb_set_real_operand_mode(ctx.h, 0) == 0 and lib.oqb_get_real_operand_mode(ctx.h, C.byref(on)) == 0 and on.value == 0
    assert lib.oqb_set_real_operand_mode(ctx.h, 1) == 0
