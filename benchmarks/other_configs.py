#!/usr/bin/env python
"""Measurements of the other BASELINE.json configurations (C2 Lindblad, C4 trajectories, C5 Redfield)
on one GPU, same timing hygiene as bench.py (≥3 warm-ups, CUDA events, inputs larger than L2).
One JSON line per configuration.  bench.py (the driver contract) stays on the headline C3 workload.

  python benchmarks/other_configs.py [--configs c2,c2dense,c4,c5] [--steps K]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _exec_ratio():
    """Executed / algorithmic flops of the ZGEMM: 0.75 for the default 3-multiplication complex product (6·M·N·K of 8·M·N·K)."""
    import oqb_b200 as Q
    return 1.0 if Q.get_context().zgemm_mode() == 0 else 0.75


def _ame_real(ctx, op):
    """True when the handle's eigenvectors are exactly real and the real-operand products are on (every GEMM of the
    step then executes 4·M·N·K of its 8·M·N·K algorithmic flops)."""
    f = C.c_int()
    ctx.check(ctx.lib.oqb_ame_eigenvectors_real(op._ame, C.byref(f)))
    return bool(f.value) and ctx.real_operand_mode()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = 6650.0
    src = "fallback 6.65 TB/s (B200_PROFILING.md)"
    if os.path.exists(p):
        hbm = json.load(open(p))["hbm_gbs"]
        src = "MEASURED_PEAKS.json hbm_gbs"
    return hbm, src


def tfim_dense(nq):
    n = 1 << nq
    idx = np.arange(n)
    Hd = np.zeros((n, n), dtype=complex)
    for q in range(nq):
        Hd[idx ^ (1 << q), idx] = -1.0
    z = [1.0 - 2.0 * ((idx >> (nq - 1 - q)) & 1) for q in range(nq)]
    diag = -sum(z[i] * z[i + 1] for i in range(nq - 1))
    return Hd, np.diag(diag).astype(complex), z


def timed(fn, steps, warmup, torch):
    for _ in range(max(warmup, 3)):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def rand_states(torch, B, n, cols, seed):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn((B, cols, n), dtype=torch.complex128, device="cuda", generator=gen)


def sigma_minus(nq, q):
    """σ⁻ on qubit q (0-based, qubit 0 = most significant bit): |0><1| on that qubit — amplitude damping."""
    n = 1 << nq
    idx = np.arange(n)
    bit = 1 << (nq - 1 - q)
    L = np.zeros((n, n), dtype=complex)
    src = idx[(idx & bit) != 0]
    L[src ^ bit, src] = 1.0
    return L


def run_c2(Q, torch, steps, dense_L, nq=8, B=4096, damping=False):
    """C2: 8-qubit dense TFIM + 8 Lindblad operators, batch 4096 of 256×256 ρ, per-member s_b = b/B.
    C2' (SURVEY §8(d)): the same at 10 qubits — n = 1024, K = 10, B = 64."""
    n, K = 2 ** nq, nq
    rng = np.random.default_rng(2000)
    Hd, Hp, z = tfim_dense(nq)
    Ls = [rng.standard_normal((n, n)) / np.sqrt(n) + 1j * rng.standard_normal((n, n)) / np.sqrt(n) for _ in range(K)] if dense_L \
        else ([sigma_minus(nq, q) for q in range(nq)] if damping else [np.diag(zq).astype(complex) for zq in z])
    H = Q.DenseHamiltonian([lambda s: 1 - s, lambda s: s], [Hd, Hp], unit="ħ")
    lind = Q.LindbladLiouvillian([Q.Lindblad(0.1 * (1 + k / nq), L) for k, L in enumerate(Ls)])
    op = Q.DiffEqLiouvillian(H, [], [lind], n)
    p = Q.ODEParams(None, 1.0)
    t = np.arange(B) / B
    g = rand_states(torch, B, n, 16, 2000)
    rho = g.transpose(1, 2) @ g.conj()
    rho = rho / torch.diagonal(rho, dim1=1, dim2=2).sum(-1).real[:, None, None]
    u = Q.DeviceBatch(B, n, n, Q.get_context(), rho.contiguous())
    du = u.zeros_like()
    ctx = Q.get_context()
    fn = lambda: op(du, u, p, t)
    fn()
    ctx.check(ctx.lib.oqb_profile_begin(ctx.h))
    ms = timed(fn, steps, 3, torch)
    zms, zn, zfl = C.c_double(), C.c_uint64(), C.c_double()
    ctx.check(ctx.lib.oqb_profile_end(ctx.h, C.byref(zms), C.byref(zn), C.byref(zfl)))
    diag_path = (not dense_L) and ctx.get_option("lindblad_diag") != 0 and (not damping or ctx.get_option("lindblad_mono") != 0)
    # general L_k: 2 + 2K GEMMs per rho; diagonal L_k exploited (its own row): 2 GEMMs + one Hadamard pass
    flops = 8.0 * n ** 3 * ((2 if diag_path else 2 + 2 * K)) * B
    dmma, _ = ctx.probe_fp64()
    if zfl.value == 0.0:
        # no product launched: the Pauli-structured driver's commutator ran as a hypercube stencil fused with the Hadamard factor
        # (lindblad_stencil_kernel) — an HBM-bound pass: read rho, write drho = 32 B per element
        hbm, src = peaks()
        gbs = 32.0 * n * n * B / 1e9 / (ms * 1e-3)
        note = ("Pauli-structured driver (XOR diagonals) + diagonal problem terms + diagonal L_k with shared rates: the whole RHS is ONE pass "
                "(lindblad_stencil64_kernel: commutator as a stencil over the driver's XOR diagonals, fused with the Hadamard factor) — no GEMM; "
                "option lindblad_stencil = 0 gives the 2-GEMM row")
        if damping:
            note = ("Pauli-structured driver + diagonal problem terms + monomial L_k (sigma-minus per qubit): TWO passes, no GEMM — the commutator as a stencil "
                    "(lindblad_stencil64_kernel) and the dissipator incl. its anticommutator as a gather (lindblad_mono_sandwich_kernel, reads rho K times from L2); "
                    "frac counts the RHS's minimal 32 B per element against both passes' time; option lindblad_stencil = 0 gives the 2-GEMM + gather row")
        return {"note": note,
                "config": ("C2" if nq == 8 else "C2prime") + ("-damping" if damping else "-Zk") + "-stencil",
                "workload": f"{nq}-qubit dense TFIM + {K} dense-stored {'sigma-minus_k' if damping else 'Z_k'} Lindblad ops, batch {B} of {n}x{n} rho, per-member s",
                "metric": "rho_rhs_evals_per_sec", "value": B / (ms * 1e-3), "ms_per_step": ms,
                "algorithmic_tflop_per_step": flops / 1e12, "step_tflops": flops / (ms * 1e-3) / 1e12,
                "roofline": {"bound": "hbm", "kernel": "lindblad_stencil64_kernel" + (" + lindblad_mono_sandwich_kernel" if damping else ""), "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                             "algorithmic_bytes_per_step": 32.0 * n * n * B, "peak_source": src}}
    ach = zfl.value / (zms.value * 1e-3) / 1e12
    extra = {}
    if diag_path:
        extra = {"note": "diagonal L_k exploited: sum_k gamma_k L_k rho L_k' is a Hadamard product with G_b[a,c] = sum_k gamma_bk d_k[a] conj(d_k[c]) "
                         "(lindblad_diag_sandwich_kernel, 48 B per element) — 2 GEMMs per rho instead of 2+2K; OQB_LINDBLAD_DIAG=0 gives the general row",
                 "hadamard_bytes_per_step_GB": 48.0 * n * n * B / 1e9}
    if damping and diag_path:
        extra = {"note": "monomial L_k (sigma-minus per qubit: one entry per row and column) exploited: L_k rho L_k' is a gather of rho (lindblad_mono_sandwich_kernel), "
                         "L_k'L_k diagonal — 2 GEMMs per rho instead of 2+2K; option lindblad_mono = 0 gives the general row"}
    return {**extra, "config": ("C2" if nq == 8 else "C2prime") + ("-denseL" if dense_L else (("-damping-monoL" if damping else "-Zk-diagL") if diag_path else ("-damping" if damping else "-Zk"))),
            "workload": f"{nq}-qubit dense TFIM + {K} {'random dense' if dense_L else ('dense-stored sigma-minus_k (amplitude damping)' if damping else 'dense-stored Z_k')} Lindblad ops, batch {B} of {n}x{n} rho, per-member s",
            "metric": "rho_rhs_evals_per_sec", "value": B / (ms * 1e-3), "ms_per_step": ms,
            "algorithmic_tflop_per_step": flops / 1e12, "step_tflops": flops / (ms * 1e-3) / 1e12,
            "roofline": {"bound": "tensor", "kernel": "zgemm_tma_kernel", "achieved": ach, "peak": dmma, "unit": "TFLOP/s",
                         "frac": ach / dmma,
                         "frac_executed": ach * (0.5 if ((not dense_L) and ctx.real_operand_mode())  # real H terms and real Z_k: every product is a real-operand product
                                                 else _exec_ratio()) / dmma,
                         "kernel_share_of_step": zms.value / ((max(3, 3) + steps) * ms),
                         "peak_source": "DMMA issue probe, this run"}}


def run_c3_onesided(Q, torch, steps, K=10, B=64):
    """C3, one-sided AME: 10-qubit OneSidedAMELiouvillian, pairs (α,α) for K couplings Z_α, Ohmic γ, S ≡ 0."""
    import bench
    n = bench.N
    Hd, Hp, coup = bench.build_c3()
    coup = coup[:K]
    H = Q.DenseHamiltonian([bench.fA, bench.fB], [Hd, Hp], unit="h")
    osd = Q.OneSidedAMELiouvillian(Q.ConstantCouplings(coup, unit="h"), Q.Ohmic(bench.ETA, bench.FC, bench.TMK), Q.ZERO_LAMBSHIFT,
                                   [(a + 1, a + 1) for a in range(K)])
    op = Q.DiffEqLiouvillian(H, [osd], [], n)
    ctx = Q.get_context()
    import time
    prep = {}
    op.plan_cache = 1
    for mode in ("host_tables", "device_tables"):  # a NEW eigen-system per call: eigh on the device, γ tables host vs device
        op.device_eigh, op.device_tables = True, mode == "device_tables"
        best = 1e9
        for k in range(3):
            t0 = time.perf_counter()
            op._prepare_ame(bench.S_POINT + 0.01 * (k + 1) + (0.004 if mode == "device_tables" else 0.0))
            ctx.sync()
            best = min(best, time.perf_counter() - t0)
        prep[mode + "_s"] = best
    op.device_eigh = op.device_tables = False
    op.plan_cache = 8
    op.clear_plans()
    op._prepare_ame(bench.S_POINT)
    g = rand_states(torch, B, n, 32, 3100)
    rho = g.transpose(1, 2) @ g.conj()
    rho = rho / torch.diagonal(rho, dim1=1, dim2=2).sum(-1).real[:, None, None]
    u = Q.DeviceBatch(B, n, n, ctx, rho.contiguous())
    du = u.zeros_like()
    fn = lambda: ctx.check(ctx.lib.oqb_ame_rhs(op._ame, B, u.ptr, du.ptr))
    fn()
    ctx.check(ctx.lib.oqb_profile_begin(ctx.h))
    ms = timed(fn, steps, 3, torch)
    zms, zn, zfl = C.c_double(), C.c_uint64(), C.c_double()
    ctx.check(ctx.lib.oqb_profile_end(ctx.h, C.byref(zms), C.byref(zn), C.byref(zfl)))
    flops = 8.0 * n ** 3 * (5 + 2 * K) * B  # 2 forward + P ρ̃ + K·(Λρ̃, ·A) + 2 back
    dmma, _ = ctx.probe_fp64()
    ach = zfl.value / (zms.value * 1e-3) / 1e12
    real = _ame_real(ctx, op)
    return {"config": f"C3-onesided-K{K}", "workload": f"10-qubit OneSidedAMELiouvillian, {K} pairs, batch {B} of {n}x{n} rho",
            "real_operand_products": real,
            "metric": "rho_rhs_evals_per_sec", "value": B / (ms * 1e-3), "ms_per_step": ms,
            "algorithmic_tflop_per_step": flops / 1e12, "step_tflops": flops / (ms * 1e-3) / 1e12,
            "per_call_eigensystem": dict(prep, evals_per_sec_incl_host_tables=B / (ms * 1e-3 + prep["host_tables_s"]),
                                         evals_per_sec_incl_device_tables=B / (ms * 1e-3 + prep["device_tables_s"])),
            "roofline": {"bound": "tensor", "kernel": "zgemm_tma_kernel", "achieved": ach, "peak": dmma, "unit": "TFLOP/s",
                         "frac": ach / dmma, "frac_executed": ach * (0.5 if real else _exec_ratio()) / dmma,
                         "kernel_share_of_step": zms.value / ((3 + steps) * ms), "peak_source": "DMMA issue probe, this run"}}


def run_c3_lambshift(Q, torch, steps, B=64):
    """C3 with the Lamb shift switched on ("S from a 1001-point table on [−50,50]"):
    (1) build_lambshift's tabulation — 1001 adaptive quadratures — on the device vs the CPU port on a sample,
    (2) the per-call eigen-system cost with the S(ω̃_ab) table evaluated from the spline on the device vs on the host,
    (3) the resident RHS rate (the tables are precomputed, so it must equal the S ≡ 0 number)."""
    import time
    import bench
    from oracle import oqb_oracle as O
    n = bench.N
    Hd, Hp, coup = bench.build_c3()
    H = Q.DenseHamiltonian([bench.fA, bench.fB], [Hd, Hp], unit="h")
    bath = Q.Ohmic(bench.ETA, bench.FC, bench.TMK)
    ctx = Q.get_context()
    w_range = np.linspace(-50, 50, 1001)
    bath.S(w_range[:8])  # warm-up (workspace growth, module load)
    t0 = time.perf_counter()
    S_tab, err, nseg = bath.S(w_range, info=True)
    t_dev = time.perf_counter() - t0
    bo = O.Ohmic(bench.ETA, bench.FC, bench.TMK)
    sample = np.arange(0, 1001, 25)
    t0 = time.perf_counter()
    S_cpu = np.array([O.lambshift(w_range[i], bo.spectrum) for i in sample])
    t_cpu = (time.perf_counter() - t0) / len(sample) * len(w_range)
    S_fn = Q.build_lambshift(w_range, True, bath)
    op = Q.DiffEqLiouvillian(H, [Q.DaviesGenerator(Q.ConstantCouplings(coup, unit="h"), bath, S_fn)], [], n)
    op.plan_cache = 1
    prep = {}
    for mode in ("host_tables", "device_tables"):
        op.device_eigh, op.device_tables = True, mode == "device_tables"
        best = 1e9
        for k in range(3):
            t0 = time.perf_counter()
            op._prepare_ame(bench.S_POINT + 0.01 * (k + 1) + (0.004 if mode == "device_tables" else 0.0))
            ctx.sync()
            best = min(best, time.perf_counter() - t0)
        prep[mode + "_s"] = best
    g = rand_states(torch, B, n, 32, 3300)
    rho = g.transpose(1, 2) @ g.conj()
    rho = rho / torch.diagonal(rho, dim1=1, dim2=2).sum(-1).real[:, None, None]
    u = Q.DeviceBatch(B, n, n, ctx, rho.contiguous())
    du = u.zeros_like()
    w_chk, v_chk = Q.haml_eigs(H, bench.S_POINT, n)
    p = Q.ODEParams(H, bench.TF)
    chk = []
    for flag in (False, True):
        op.device_eigh, op.device_tables = False, flag
        op.clear_plans()
        chk.append(op.rhs_with_eig(du.zeros_like(), u, p, bench.S_POINT * bench.TF, w_chk, v_chk).t.clone())
    rel = float((chk[0] - chk[1]).norm().item() / chk[0].norm().item())
    op0 = Q.DiffEqLiouvillian(H, [Q.DaviesGenerator(Q.ConstantCouplings(coup, unit="h"), bath, Q.ZERO_LAMBSHIFT)], [], n)
    d0 = op0.rhs_with_eig(du.zeros_like(), u, p, bench.S_POINT * bench.TF, w_chk, v_chk).t
    lamb_effect = float((chk[1] - d0).norm().item() / chk[1].norm().item())
    del chk, d0
    fn = lambda: ctx.check(ctx.lib.oqb_ame_rhs(op._ame, B, u.ptr, du.ptr))
    fn()
    ms = timed(fn, steps, 3, torch)
    return {"config": "C3-lambshift", "workload": f"10-qubit DaviesGenerator with tabulated Lamb shift (1001-point spline on [-50,50]), batch {B}",
            "metric": "rho_rhs_evals_per_sec", "value": B / (ms * 1e-3), "ms_per_step": ms,
            "lambshift_table": {"points": 1001, "device_s": t_dev, "cpu_port_s_extrapolated": t_cpu, "cpu_sample_points": len(sample),
                                "speedup": t_cpu / t_dev, "segments_median": float(np.median(nseg)), "segments_max": int(nseg.max()),
                                "max_rel_diff_vs_cpu_port_on_sample": float(np.max(np.abs(S_tab[sample] - S_cpu) / np.abs(S_cpu))),
                                "rtol": 1.4901161193847656e-08},
            "per_call_eigensystem_with_lambshift": dict(prep, evals_per_sec_incl_host_tables=B / (ms * 1e-3 + prep["host_tables_s"]),
                                                        evals_per_sec_incl_device_tables=B / (ms * 1e-3 + prep["device_tables_s"])),
            "device_tables_rel_diff_vs_host_tables": rel, "rhs_rel_change_due_to_lambshift": lamb_effect}


def tfim_sparse_c4(nq):
    """C4 operators: X-part (nq entries per row) and ZZ-part (diagonal) as CSC terms, L_k = Z_k sparse diagonal."""
    import scipy.sparse as sps
    n = 1 << nq
    idx = np.arange(n)
    rows = np.concatenate([idx ^ (1 << q) for q in range(nq)])
    cols = np.tile(idx, nq)
    Hx = sps.csc_matrix((-np.ones(n * nq), (rows, cols)), shape=(n, n), dtype=complex)
    z = [1.0 - 2.0 * ((idx >> (nq - 1 - q)) & 1) for q in range(nq)]
    Hz = sps.diags(-sum(z[i] * z[i + 1] for i in range(nq - 1))).tocsc().astype(complex)
    Ls = [sps.diags(zq).tocsc().astype(complex) for zq in z]
    return Hx, Hz, Ls


def run_c4(Q, torch, steps):
    """C4: 12-qubit sparse TFIM trajectories, 65536 ψ of dimension 4096, per-trajectory s_b, γ_k = 0.05."""
    nq, n, B = 12, 4096, 65536
    Hx, Hz, Ls = tfim_sparse_c4(nq)
    H = Q.SparseHamiltonian([lambda s: 1 - s, lambda s: s], [Hx, Hz], unit="ħ")
    op = Q.DiffEqLiouvillian(H, [], [Q.LindbladLiouvillian([Q.Lindblad(0.05, L) for L in Ls])], n)
    ens = Q.TrajectoryEnsemble(op)
    ctx = Q.get_context()
    psi_t = rand_states(torch, B, n, 1, 4000)
    psi_t = psi_t / torch.linalg.vector_norm(psi_t, dim=2, keepdim=True)
    psi = Q.DeviceBatch(B, n, 1, ctx, psi_t.contiguous())
    dpsi = psi.zeros_like()
    rng = np.random.default_rng(4000)
    s = rng.random(B)
    cf = np.ascontiguousarray(np.stack([1 - s, s], axis=1))
    gam = np.full((1, nq), 0.05)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    mv = lambda: ctx.check(ctx.lib.oqb_traj_matvec(ens.sp, B, dp(cf), 0, dp(gam), 1, psi.ptr, dpsi.ptr))
    ms_mv = timed(mv, steps, 3, torch)
    def kernel_only(fn):
        ctx.check(ctx.lib.oqb_profile_begin(ctx.h))
        for _ in range(steps):
            fn()
        a, b_, w = C.c_double(), C.c_uint64(), C.c_double()
        ctx.check(ctx.lib.oqb_profile_end(ctx.h, C.byref(a), C.byref(b_), C.byref(w)))
        return a.value / max(b_.value, 1), w.value / max(b_.value, 1)
    kms_mv, kb_mv = kernel_only(mv)
    uni = torch.as_tensor(rng.random(B), device="cuda")
    choice = torch.zeros(B, dtype=torch.int32, device="cuda")
    prob = torch.zeros((B, nq), dtype=torch.float64, device="cuda")
    jp = lambda ap: (lambda: ctx.check(ctx.lib.oqb_traj_jump(ens.sp, B, dp(gam), 1, psi.ptr, C.c_void_p(uni.data_ptr()), None,
                                                             C.c_void_p(choice.data_ptr()), C.c_void_p(prob.data_ptr()), ap)))
    ms_j0 = timed(jp(0), steps, 3, torch)
    ms_j1 = timed(jp(1), steps, 3, torch)  # Z_k are unitary: repeated in-place application keeps ψ normalised
    kms_j0, kb_j0 = kernel_only(jp(0))
    kms_j1, kb_j1 = kernel_only(jp(1))
    # library bar: cuSPARSE SpMM, CSR (4096×4096, the assembled −iH_eff at one s) × dense (4096 × B) through
    # torch.sparse — one shared operator for the whole ensemble (it cannot express per-trajectory coefficients)
    lib_ms = None
    try:
        Hs = (-1j * (0.5 * Hx + 0.5 * Hz)).tocsr()
        A_sp = torch.sparse_csr_tensor(torch.as_tensor(Hs.indptr, dtype=torch.int64), torch.as_tensor(Hs.indices, dtype=torch.int64),
                                       torch.as_tensor(Hs.data), size=(n, n), dtype=torch.complex128, device="cuda")
        Xd = psi.t.reshape(B, n).T.contiguous()  # n × B dense
        lib_ms = timed(lambda: torch.sparse.mm(A_sp, Xd), max(2, steps // 2), 2, torch)
        del Xd, A_sp
    except Exception as e:  # noqa: BLE001
        lib_ms = None
    hbm, src = peaks()
    gb_mv, gb_j0, gb_j1 = 32.0 * n * B / 1e9, 16.0 * n * B / 1e9, 32.0 * n * B / 1e9
    # on-device stepping (oqb_traj_evolve): RK4 + jump loop, one shared schedule, γ = 0.05, dt = 0.01
    st = Q.TrajectoryStepper(ens, B, seed=1)
    pO = Q.ODEParams(None, 10.0)
    nst = 5
    st.evolve(psi, pO, 0.0, 0.01, 2)
    ms_ev = timed(lambda: st.evolve(psi, pO, 0.0, 0.01, nst), max(2, steps // 5), 1, torch) / nst
    njump = int(st.njumps.sum().item())
    fusedrk = ctx.get_option("rk_fuse") != 0
    # vector passes per RK4 step: fused stage arithmetic 3+5+5+3 (the norm rides the last stage); unfused 4 matvecs (2 each)
    # + stage kernels 4+5+5+3 + 1 norm pass
    passes = 16 if fusedrk else 26
    moved = passes * 16.0 * n * B / 1e9
    return {"config": "C4", "workload": f"12-qubit SparseHamiltonian TFIM trajectories, {B} psi of dim {n}, per-trajectory s, K=12 diagonal L_k",
            "metric": "trajectory_matvecs_per_sec", "value": B / (ms_mv * 1e-3), "ms_per_step": ms_mv,
            "call_GBs_incl_coef_upload": gb_mv / (ms_mv * 1e-3),
            "roofline": {"bound": "hbm", "kernel": "trajectory matvec kernel (CUDA-event pair around the launch)",
                         "achieved": kb_mv / 1e9 / (kms_mv * 1e-3), "peak": hbm, "unit": "GB/s",
                         "frac": kb_mv / 1e9 / (kms_mv * 1e-3) / hbm, "frac_of_nominal_8TBs": kb_mv / 1e9 / (kms_mv * 1e-3) / 8000.0,
                         "peak_source": src, "kernel_ms": kms_mv, "algorithmic_bytes_per_launch": kb_mv},
            "cusparse_spmm_ms": lib_ms, "cusparse_note": "torch.sparse.mm CSR x dense, ONE operator for all trajectories (no per-trajectory s)",
            "jump_kernel": {"decide_only_ms": kms_j0, "decide_only_frac": kb_j0 / 1e9 / (kms_j0 * 1e-3) / hbm,
                            "decide_and_apply_ms": kms_j1, "decide_and_apply_frac": kb_j1 / 1e9 / (kms_j1 * 1e-3) / hbm},
            "jump": {"decide_only_ms": ms_j0, "decide_only_GBs": gb_j0 / (ms_j0 * 1e-3), "decide_only_frac": gb_j0 / (ms_j0 * 1e-3) / hbm,
                     "decide_and_apply_ms": ms_j1, "decide_and_apply_GBs": gb_j1 / (ms_j1 * 1e-3),
                     "decide_and_apply_frac": gb_j1 / (ms_j1 * 1e-3) / hbm},
            "evolve": {"what": "oqb_traj_evolve: fixed-step RK4 + jump loop on the device, xoshiro256++ per trajectory, shared schedule",
                       "ms_per_rk4_step": ms_ev, "trajectory_steps_per_sec": B / (ms_ev * 1e-3), "jumps_so_far": njump,
                       "vector_passes_per_step": passes, "bytes_moved_model_GB_per_step": moved, "GBs_of_moved_bytes": moved / (ms_ev * 1e-3),
                       "frac_of_measured_hbm": moved / (ms_ev * 1e-3) / hbm}}


def random_sparse_c4(nq, per_row=6, seed=4100):
    """An UNSTRUCTURED 12-qubit SparseHamiltonian: `per_row` random off-diagonal entries per row with random complex values,
    symmetrised (≈ 2·per_row + 1 = 13 entries per row like the TFIM, but no Pauli/XOR structure), plus a random real diagonal term."""
    import scipy.sparse as sps
    n = 1 << nq
    rng = np.random.default_rng(seed)
    rows = np.repeat(np.arange(n), per_row)
    cols = rng.integers(0, n, size=n * per_row)
    keep = rows != cols
    vals = rng.standard_normal(n * per_row) + 1j * rng.standard_normal(n * per_row)
    A = sps.csc_matrix((vals[keep], (rows[keep], cols[keep])), shape=(n, n))
    Hoff = (A + A.conj().T).tocsc()
    Hdiag = sps.diags(rng.standard_normal(n)).tocsc().astype(complex)
    return Hoff, Hdiag


def pauli_sparse_c4(nq, model):
    """12-qubit Pauli-string Hamiltonians WITHOUT pure-X structure: "yfield" = −Σ Y_i driver (imaginary, row-signed entries) +
    Σ Z_iZ_{i+1}; "heisenberg" = ½Σ(X_iX_{i+1} + Y_iY_{i+1}) (two strings per bond sharing their flip mask) + Σ Z_iZ_{i+1}."""
    import scipy.sparse as sps
    n = 1 << nq
    idx = np.arange(n)
    bit = lambda q: (idx >> (nq - 1 - q)) & 1
    z = [1.0 - 2.0 * bit(q) for q in range(nq)]
    Hz = sps.diags(-sum(z[i] * z[i + 1] for i in range(nq - 1))).tocsc().astype(complex)
    rows, cols, vals = [], [], []
    if model == "yfield":
        for q in range(nq):  # Y = [[0, −i], [i, 0]]: entry (r ^ m, r) = i·(−1)^bit_q(r)
            m = 1 << (nq - 1 - q)
            rows.append(idx ^ m); cols.append(idx); vals.append(-1.0 * 1j * z[q])
    else:
        for q in range(nq - 1):  # XX + YY on a bond flips both bits and vanishes when they are equal: (1 − z_q z_{q+1})
            m = (1 << (nq - 1 - q)) | (1 << (nq - 2 - q))
            v = 0.5 * (1.0 - z[q] * z[q + 1])
            keep = v != 0
            rows.append((idx ^ m)[keep]); cols.append(idx[keep]); vals.append(v[keep].astype(complex))
    Hx = sps.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n), dtype=complex)
    Ls = [sps.diags(zq).tocsc().astype(complex) for zq in z]
    return Hx, Hz, Ls


def run_c4_matvec_only(Q, torch, steps, generic=False, random_pattern=False, pauli=None, plain_ell=False):
    """The C4 ensemble matvec alone: `generic=True` forces the kernels for operators WITHOUT Pauli structure on the TFIM operators
    (option "traj_kernel" = 1), `random_pattern=True` swaps in an unstructured Hamiltonian (no Pauli specialisation applies); both take
    the sliced-ELL kernel (csrc/traj_sell.cu) unless `plain_ell` (option "traj_ell" = 0: the round-1 ELL kernel).  HBM roofline
    fraction of the kernel."""
    nq, n, B = 12, 4096, 65536
    ctx = Q.get_context()
    if pauli:
        Ha, Hb, Ls = pauli_sparse_c4(nq, pauli)
    elif random_pattern:
        Ha, Hb = random_sparse_c4(nq)
        _, _, Ls = tfim_sparse_c4(nq)
    else:
        Ha, Hb, Ls = tfim_sparse_c4(nq)
    H = Q.SparseHamiltonian([lambda s: 1 - s, lambda s: s], [Ha, Hb], unit="ħ")
    op = Q.DiffEqLiouvillian(H, [], [Q.LindbladLiouvillian([Q.Lindblad(0.05, L) for L in Ls])], n)
    ens = Q.TrajectoryEnsemble(op)
    psi_t = rand_states(torch, B, n, 1, 4000)
    psi_t = psi_t / torch.linalg.vector_norm(psi_t, dim=2, keepdim=True)
    psi = Q.DeviceBatch(B, n, 1, ctx, psi_t.contiguous())
    dpsi = psi.zeros_like()
    rng = np.random.default_rng(4000)
    s = rng.random(B)
    cf = np.ascontiguousarray(np.stack([1 - s, s], axis=1))
    gam = np.full((1, nq), 0.05)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    mv = lambda: ctx.check(ctx.lib.oqb_traj_matvec(ens.sp, B, dp(cf), 0, dp(gam), 1, psi.ptr, dpsi.ptr))
    with ctx.option("traj_kernel", 1 if generic else 0), ctx.option("traj_ell", 0 if plain_ell else 1):
        ms = timed(mv, steps, 3, torch)
        ctx.check(ctx.lib.oqb_profile_begin(ctx.h))
        for _ in range(steps):
            mv()
        a, b_, w = C.c_double(), C.c_uint64(), C.c_double()
        ctx.check(ctx.lib.oqb_profile_end(ctx.h, C.byref(a), C.byref(b_), C.byref(w)))
    kms, kb = a.value / max(b_.value, 1), w.value / max(b_.value, 1)
    hbm, src = peaks()
    nnz_row = (Ha.nnz + Hb.nnz) / n
    pl, pad = C.c_int(), C.c_double()
    ctx.check(ctx.lib.oqb_traj_sell_stats(ens.sp, C.byref(pl), C.byref(pad)))
    kernel = "traj_xor_kernel" if not (generic or random_pattern) else ("traj_matvec_kernel (plain ELL)" if plain_ell or not pl.value else "traj_sell_kernel")
    ceil = None
    if random_pattern or generic:
        # what the shared-memory gathers alone allow (measured on this device): gathers/s ÷ off-diagonal entries per row × 32 B.
        # Conflict-free order is what traj_sell_kernel's slot colouring buys (at `sell_padding` × the gathers); the random
        # order is what a kernel that takes the columns as they come is capped at.
        per_row = Ha.nnz / n
        ceil = {"gathers_per_row": per_row,
                "what": "oqb_probe_smem_gather: LDS.128 gathers from a staged 4096-entry state, nothing else in the kernel — an upper bound for any kernel that gathers psi from shared memory"}
        for pat, label in ((0, "conflict_free"), (1, "random_order")):
            if pat == 1 and not random_pattern:
                continue
            g = C.c_double()
            ctx.check(ctx.lib.oqb_probe_smem_gather(ctx.h, pat, C.byref(g)))
            ceil[label] = {"gathers_per_sec": g.value, "ceiling_GBs": g.value / per_row * 32.0 / 1e9, "ceiling_frac_of_hbm": g.value / per_row * 32.0 / 1e9 / hbm}
    return {"kernel": kernel, "sell_padding": pad.value if pl.value else None, "smem_gather_ceiling": ceil, "workload": f"12-qubit {('Pauli-string (' + pauli + ')') if pauli else ('UNSTRUCTURED random-pattern' if random_pattern else 'TFIM')} SparseHamiltonian, {B} psi of dim {n}, "
                        f"{nnz_row:.1f} entries per row, per-trajectory s" + (", generic kernels forced" if generic else ""),
            "value": B / (ms * 1e-3), "ms_per_step": ms,
            "roofline": {"bound": "hbm", "achieved": kb / 1e9 / (kms * 1e-3), "peak": hbm, "unit": "GB/s",
                         "frac": kb / 1e9 / (kms * 1e-3) / hbm, "kernel_ms": kms, "algorithmic_bytes_per_launch": kb, "peak_source": src}}


def run_c5(Q, torch, steps):
    """C5 (per-GPU share): 1024 schedules × 7-qubit Redfield RHS (commutator + apply with supplied Λ), n = 128, K = 7."""
    nq, n, K, B = 7, 128, 7, 1024
    Hd, Hp, z = tfim_dense(nq)
    S = [np.diag(zq).astype(complex) for zq in z]
    H = Q.DenseHamiltonian([lambda s: 1 - s, lambda s: s], [Hd, Hp], unit="ħ")
    ctx = Q.get_context()
    g = rand_states(torch, B, n, 16, 5000)
    rho = g.transpose(1, 2) @ g.conj()
    rho = rho / torch.diagonal(rho, dim1=1, dim2=2).sum(-1).real[:, None, None]
    u = Q.DeviceBatch(B, n, n, ctx, rho.contiguous())
    du = u.zeros_like()
    Lam = Q.DeviceBatch(B * K, n, n, ctx, rand_states(torch, B * K, n, n, 5001) / n)
    Sd = Q.DeviceBatch.from_host(np.stack(S), ctx)
    pw = 0.5 + 1.5 * np.arange(B) / B
    s = 0.37 ** pw

    diag_path = os.environ.get("OQB_REDFIELD_DIAG", "1") != "0"  # S_α = Z_α are diagonal: elementwise form of S M − M S
    Sdiag = Q.DeviceBatch.from_host(np.stack([np.diag(x) for x in S]), ctx)

    def fn():
        H.commutator(du, u, s)
        if diag_path:
            ctx.check(ctx.lib.oqb_redfield_apply_diag_acc(ctx.h, n, K, Sdiag.ptr, Lam.ptr, B, u.ptr, du.ptr))
        else:
            ctx.check(ctx.lib.oqb_redfield_apply_acc(ctx.h, n, K, Sd.ptr, 1 | 2, Lam.ptr, B, u.ptr, du.ptr))  # bit 1: the Z_α couplings are real
    fn()
    ctx.check(ctx.lib.oqb_profile_begin(ctx.h))
    ms = timed(fn, steps, 3, torch)
    zms, zn, zfl = C.c_double(), C.c_uint64(), C.c_double()
    ctx.check(ctx.lib.oqb_profile_end(ctx.h, C.byref(zms), C.byref(zn), C.byref(zfl)))
    flops = 8.0 * n ** 3 * (2 + (K if diag_path else 3 * K)) * B
    dmma, _ = ctx.probe_fp64()
    ach = zfl.value / (zms.value * 1e-3) / 1e12
    return {"config": "C5-per-GPU" + ("-diagS" if diag_path else ""), "workload": f"{B} schedules x 7-qubit Redfield RHS (commutator + apply, supplied Lambda), n={n}, K={K}, per-schedule s"
            + ("; diagonal couplings Z_α exploited: K GEMMs + one elementwise pass per schedule instead of 3K GEMMs (OQB_REDFIELD_DIAG=0: general row)" if diag_path else ""),
            "metric": "rho_rhs_evals_per_sec", "value": B / (ms * 1e-3), "ms_per_step": ms,
            "algorithmic_tflop_per_step": flops / 1e12, "step_tflops": flops / (ms * 1e-3) / 1e12,
            "roofline": {"bound": "tensor", "kernel": "zgemm_tma_kernel", "achieved": ach, "peak": dmma, "unit": "TFLOP/s",
                         "frac": ach / dmma, "frac_executed": ach * _exec_ratio() / dmma, "kernel_share_of_step": zms.value / ((3 + steps) * ms), "peak_source": "DMMA issue probe, this run"}}


def run_c5_sigmax(Q, torch, steps, mono=True):
    """C5 with TRANSVERSE couplings S_α = X_α (monomial: one entry per row and column) instead of Z_α: the apply with the
    gather form of S M − M S (oqb_redfield_apply_mono_acc: K products per schedule) or, mono=False, the general 3K products."""
    nq, n, K, B = 7, 128, 7, 1024
    Hd, Hp, z = tfim_dense(nq)
    idx = np.arange(n)
    S = []
    for q in range(nq):
        m = np.zeros((n, n), dtype=complex)
        m[idx ^ (1 << (nq - 1 - q)), idx] = 1.0
        S.append(m)
    H = Q.DenseHamiltonian([lambda s: 1 - s, lambda s: s], [Hd, Hp], unit="ħ")
    ctx = Q.get_context()
    g = rand_states(torch, B, n, 16, 5000)
    rho = g.transpose(1, 2) @ g.conj()
    rho = rho / torch.diagonal(rho, dim1=1, dim2=2).sum(-1).real[:, None, None]
    u = Q.DeviceBatch(B, n, n, ctx, rho.contiguous())
    du = u.zeros_like()
    Lam = Q.DeviceBatch(B * K, n, n, ctx, rand_states(torch, B * K, n, n, 5001) / n)
    Sd = Q.DeviceBatch.from_host(np.stack(S), ctx)
    perm = np.stack([np.nonzero(x)[1][np.argsort(np.nonzero(x)[0])] for x in S]).astype(np.int32)   # row a -> its column
    pinv = np.stack([np.nonzero(x)[0][np.argsort(np.nonzero(x)[1])] for x in S]).astype(np.int32)   # column c -> its row
    dperm, dpinv = torch.as_tensor(perm, device="cuda"), torch.as_tensor(pinv, device="cuda")
    dval = Q.DeviceBatch.from_host(np.ones((K, n), dtype=complex), ctx)
    pw = 0.5 + 1.5 * np.arange(B) / B
    s = 0.37 ** pw

    def fn():
        H.commutator(du, u, s)
        if mono:
            ctx.check(ctx.lib.oqb_redfield_apply_mono_acc(ctx.h, n, K, C.c_void_p(dperm.data_ptr()), C.c_void_p(dpinv.data_ptr()), dval.ptr,
                                                          Lam.ptr, B, u.ptr, du.ptr))
        else:
            ctx.check(ctx.lib.oqb_redfield_apply_acc(ctx.h, n, K, Sd.ptr, 1 | 2, Lam.ptr, B, u.ptr, du.ptr))
    fn()
    ms = timed(fn, steps, 3, torch)
    return {"config": "C5-per-GPU-sigmax" + ("-mono" if mono else "-general"),
            "workload": f"{B} schedules x 7-qubit Redfield RHS with transverse couplings X_α (commutator + apply, supplied Lambda), n={n}, K={K}: "
                        + ("monomial couplings as gathers, K GEMMs per schedule" if mono else "general route, 3K GEMMs per schedule"),
            "metric": "rho_rhs_evals_per_sec", "value": B / (ms * 1e-3), "ms_per_step": ms}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c2,c2general,c2dense,c2prime,c2prime_diag,c3onesided,c4,c5,c5general")
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import oqb_b200 as Q
    Q.get_context()
    for c in args.configs.split(","):
        if c == "c5general":
            import subprocess
            env = dict(os.environ, OQB_REDFIELD_DIAG="0")
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--configs", "c5", "--steps", str(args.steps)],
                                 env=env, capture_output=True, text=True)
            sys.stderr.write(out.stderr[-2000:])
            print(out.stdout.strip().splitlines()[-1], flush=True)
            continue
        if c in ("c5sigmax", "c5sigmax_general"):
            print(json.dumps(run_c5_sigmax(Q, torch, args.steps, mono=c == "c5sigmax")), flush=True)
            torch.cuda.empty_cache()
            continue
        if c == "c2general":  # same operators, diagonal structure NOT exploited
            with Q.get_context().option("lindblad_diag", 0):
                r = run_c2(Q, torch, args.steps, False)
            print(json.dumps(r), flush=True)
            torch.cuda.empty_cache()
            continue
        if c in ("c4generic", "c4random", "c4yfield", "c4heisenberg", "c4generic_plain", "c4random_plain"):
            r = run_c4_matvec_only(Q, torch, args.steps, generic=c.startswith("c4generic"), random_pattern=c.startswith("c4random"), plain_ell=c.endswith("_plain"),
                                   pauli={"c4yfield": "yfield", "c4heisenberg": "heisenberg"}.get(c))
            print(json.dumps(r), flush=True)
            torch.cuda.empty_cache()
            continue
        if c in ("c2_tiles", "c2prime_diag_tiles"):  # stencil route with the plain kernel (no register micro-tiles)
            with Q.get_context().option("lindblad_stencil", 1):
                r = run_c2(Q, torch, args.steps, False) if c == "c2_tiles" else run_c2(Q, torch, args.steps, False, nq=10, B=64)
            r["config"] += "-plain-kernel"
            print(json.dumps(r), flush=True)
            torch.cuda.empty_cache()
            continue
        if c == "c2":
            r = run_c2(Q, torch, args.steps, False)
        elif c == "c2dense":
            r = run_c2(Q, torch, args.steps, True)
        elif c == "c2prime":
            r = run_c2(Q, torch, args.steps, True, nq=10, B=64)
        elif c == "c2prime_damping":  # 10-qubit amplitude damping: monomial operators
            r = run_c2(Q, torch, args.steps, False, nq=10, B=64, damping=True)
        elif c == "c2_damping":
            r = run_c2(Q, torch, args.steps, False, damping=True)
        elif c == "c2prime_diag":  # the 10-qubit Lindblad RHS with per-qubit dephasing (north_star's "batched 10-qubit Lindblad RHS")
            r = run_c2(Q, torch, args.steps, False, nq=10, B=64)
        elif c == "c3onesided":
            r = run_c3_onesided(Q, torch, args.steps)
        elif c == "c3lamb":
            r = run_c3_lambshift(Q, torch, args.steps)
        elif c == "c4":
            r = run_c4(Q, torch, args.steps)
        elif c == "c5":
            r = run_c5(Q, torch, args.steps)
        else:
            continue
        print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
