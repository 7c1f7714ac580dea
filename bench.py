#!/usr/bin/env python
"""bench.py — ρ·RHS evaluations per second on BASELINE.json's headline configuration.

Workload (config C3 of SURVEY §8d, BASELINE.json configs[2]): 10-qubit (1024×1024) DaviesGenerator/AME
RHS with an Ohmic bath — eigenbasis transform, γ(ω_ab) Hadamard, back-transform — on a batch of 64
ComplexF64 density matrices per GPU.  One "step" = one batched RHS call (64 evaluations per GPU).

  value     device-resident states, CUDA-event timed (whole job: all ranks' evaluations ÷ max time)
  e2e       same call through the host-buffer entry point (pinned host u/du, H2D+D2H inside)
  roofline  the ZGEMM kernel (FP64 DMMA) timed per launch with CUDA events inside the timed region
  cpu_baseline  reference-structured CPU restatement on a bounded sample (oracle/cpu_baseline.py)

`--impl reference` times only the CPU restatement of the reference path (Julia is not installed here or
on the GPU box, so the reference package itself cannot run).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NQ, N, BATCH, KCOUP = 10, 1024, 64, 10
S_POINT, TF = 0.5, 10.0
ETA, FC, TMK = 1e-4, 4.0, 16.0
METRIC = "rho_rhs_evals_per_sec_10q_ame_batch64"
UNIT = "evals/s"


def build_c3(seed=3000):
    """Synthetic C3 operators (host): H(s) = −(1−s)ΣX_i + s·Σ J_ij Z_i Z_j (J ~ U[−1,1)), couplings Z_k,
    all multiplied by 2π (unit :h).  Qubit 1 is the most significant bit (kron ordering)."""
    rng = np.random.default_rng(seed)
    idx = np.arange(N)
    Hd = np.zeros((N, N), dtype=np.complex128)
    for q in range(NQ):
        Hd[idx ^ (1 << q), idx] = -1.0
    z = [1.0 - 2.0 * ((idx >> (NQ - 1 - q)) & 1) for q in range(NQ)]
    diag = np.zeros(N)
    for i in range(NQ):
        for j in range(i + 1, NQ):
            diag += (2 * rng.random() - 1) * z[i] * z[j]
    Hp = np.diag(diag).astype(np.complex128)
    coup = [np.diag(zq).astype(np.complex128) for zq in z]
    return Hd, Hp, coup


def fA(s):
    return 1 - s


def fB(s):
    return s


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {getattr(nv, k): k for k in dir(nv) if k.startswith("nvmlClocksEventReason") or k.startswith("nvmlClocksThrottleReason")}
            while not self._halt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if isinstance(bit, int) and bit and (r & bit) == bit:
                        self.reasons.add(name.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", ""))
                time.sleep(0.05)
        except Exception as e:  # NVML unavailable: report that instead of inventing numbers
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        reasons = sorted(x for x in self.reasons if x not in ("GpuIdle", "None", "ApplicationsClocksSetting"))
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": reasons}


def cpu_reference_arm(steps, warmup, evals_per_step=1):
    """The reference's CPU path (restatement, see oracle/cpu_baseline.py) on a bounded sample."""
    from oracle import cpu_baseline as CB
    cores = CB.use_all_host_threads()  # torchrun exports OMP_NUM_THREADS=1: the CPU arm must still use every host core
    Hd, Hp, coup = build_c3()
    terms = [2 * np.pi * Hd, 2 * np.pi * Hp]
    coef = [fA(S_POINT), fB(S_POINT)]
    cps = [2 * np.pi * c for c in coup]
    beta = 5 * 6.62607004 / 1.38064852 / np.pi / TMK
    bath = (ETA, 2 * np.pi * FC, beta)
    rng = np.random.default_rng(1)
    g = rng.standard_normal((N, 16)) + 1j * rng.standard_normal((N, 16))
    rho = g @ g.conj().T
    rho /= np.trace(rho).real
    states = [rho] * evals_per_step
    for _ in range(warmup):
        CB.time_ame_reference(terms, coef, cps, bath, states[:1])
    t = 0.0
    for _ in range(steps):
        dt, _n = CB.time_ame_reference(terms, coef, cps, bath, states)
        t += dt
    return steps * evals_per_step / t, t / steps * 1e3, cores


def cpu_lambda_baseline(Uh, dt, couplings, cf, tf, atol, rtol, nint):
    """CPU-baseline leg for benchmarks/lambda_bench.py (kept here: bench.py is the one measurement script allowed to
    execute oracle/): `nint` Redfield Λ integrals of one schedule through the oracle's QuadGK restatement."""
    from oracle import oqb_oracle as O
    from oqb_b200.redfield_device import SampledUnitary
    Uf = SampledUnitary(Uh, dt)
    p_ = O.ODEParams(None, tf, lambda tf_, t: t / tf_)
    red = O.RedfieldLiouvillian([(((1, 1),), couplings, cf)], Uf, tf, atol, rtol)
    t0 = time.perf_counter()
    ref = None
    for j in range(nint):
        ref = red.Lambda(p_, tf, couplings, cf, j % len(couplings) + 1)
    return time.perf_counter() - t0, ref


def run_secondary(Q, torch, ctx):
    """C2 (8-qubit Lindblad, batch 4096), C4 (12-qubit trajectory ensemble: the XOR-specialised kernel AND the generic kernel
    on the same operators, plus an unstructured random-pattern Hamiltonian) and C5 (7-qubit Redfield apply, 1024 schedules)
    in a few steps each — so that these numbers are driver-run, not builder-run.  Details: benchmarks/other_configs.py."""
    sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
    import other_configs as OC
    out = {}

    def short(r, keys):
        return {k: r[k] for k in keys if k in r}

    def clocked(fn):  # SM clock / throttle reasons while this configuration runs
        smp = ClockSampler(ctx.device)
        smp.start()
        try:
            r = fn()
        finally:
            ck = smp.stop()
        r["clocks"] = ck
        return r
    try:
        # C2 as BASELINE words it.  The driver −ΣX is Pauli-structured, the problem term and the dephasing operators are diagonal:
        # the library runs the whole RHS as one stencil + Hadamard pass (HBM row).  The same operators with that route off:
        # 2 real-operand ZGEMMs + one Hadamard pass per rho (FP64 tensor row).
        r = clocked(lambda: OC.run_c2(Q, torch, 5, False))
        out["C2_dephasing_diagL"] = short(r, ("note", "config", "workload", "value", "ms_per_step", "roofline", "clocks"))
        torch.cuda.empty_cache()
        with ctx.option("lindblad_stencil", 0):
            r = OC.run_c2(Q, torch, 3, False)
        out["C2_dephasing_2gemm_route"] = short(r, ("config", "value", "ms_per_step", "roofline"))
        torch.cuda.empty_cache()
        with ctx.option("lindblad_diag", 0):
            r = OC.run_c2(Q, torch, 2, False)
        out["C2_general_2plus2K_gemms"] = short(r, ("config", "value", "ms_per_step", "roofline"))
        torch.cuda.empty_cache()
        # north_star's "batched 10-qubit Lindblad RHS": the same operators at 10 qubits, 64 rho of 1024 x 1024
        r = OC.run_c2(Q, torch, 3, False, nq=10, B=64)
        out["C2prime_10q_dephasing"] = short(r, ("config", "workload", "value", "ms_per_step", "roofline"))
        torch.cuda.empty_cache()
        with ctx.option("lindblad_stencil", 0):
            r = OC.run_c2(Q, torch, 3, False, nq=10, B=64)
        out["C2prime_10q_dephasing_2gemm_route"] = short(r, ("config", "value", "ms_per_step", "roofline"))
        torch.cuda.empty_cache()
        # … with amplitude damping (sigma-minus_k: monomial operators) instead of dephasing: stencil pass for the commutator + gather pass for the dissipator
        r = OC.run_c2(Q, torch, 3, False, nq=10, B=64, damping=True)
        out["C2prime_10q_amplitude_damping"] = short(r, ("config", "workload", "value", "ms_per_step", "roofline"))
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out["C2_error"] = f"{type(e).__name__}: {e}"
    try:
        r = clocked(lambda: OC.run_c4(Q, torch, 5))
        out["C4_traj_xor"] = short(r, ("config", "workload", "value", "ms_per_step", "roofline", "jump_kernel", "evolve", "clocks"))
        torch.cuda.empty_cache()
        out["C4_generic_kernel_same_operators"] = OC.run_c4_matvec_only(Q, torch, 3, generic=True)
        torch.cuda.empty_cache()
        out["C4_random_pattern"] = OC.run_c4_matvec_only(Q, torch, 3, random_pattern=True)
        torch.cuda.empty_cache()
        out["C4_pauli_yfield_signed_slots"] = OC.run_c4_matvec_only(Q, torch, 3, pauli="yfield")
        torch.cuda.empty_cache()
        out["C4_pauli_heisenberg_signed_slots"] = OC.run_c4_matvec_only(Q, torch, 3, pauli="heisenberg")
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out["C4_error"] = f"{type(e).__name__}: {e}"
    try:
        r = clocked(lambda: OC.run_c5(Q, torch, 5))
        out["C5_per_gpu"] = short(r, ("config", "workload", "value", "ms_per_step", "roofline", "clocks"))
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out["C5_error"] = f"{type(e).__name__}: {e}"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C2/C4/C5 block (N = 1 only)")
    args = ap.parse_args()
    # stdout must carry exactly ONE JSON line: libraries (NCCL's version banner, torch warnings) write to fd 1 at the C
    # level, so the real stdout is kept aside and fd 1 points at stderr for the rest of the run
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "C3: 10-qubit (1024x1024) DaviesGenerator/AME RHS, Ohmic bath, 64 rho per GPU, K=10 Z_k couplings, s=0.5, lvl=1024",
              "batch_per_gpu": BATCH, "n": N, "inputs_exceed_l2": True,
              "l2_note": "u and du are 1 GiB each per GPU (> 126 MB L2); no flush needed"}
    try:  # the metric key is a slug of BASELINE.json's wording; carry the wording itself too
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "BASELINE.json")) as f:
            config["baseline_metric"] = json.load(f)["metric"]
    except Exception:
        pass

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = max(1, min(args.steps, 5))
        warm = max(0, min(args.warmup, 1))
        val, ms, cores = cpu_reference_arm(steps, warm)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
                "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64 (complex128)", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"{steps} steps x 1 rho: reference-structured AME RHS per call (H(s) assembly, LAPACK eigh, gap grouping, "
                                           "10 coupling rotations, closed-form Davies, back-transform); numpy/OpenBLAS; Julia unavailable"},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "steps/warmup clamped so the CPU arm finishes in minutes"}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    # NUMA: run this rank on the CPUs next to its GPU so that the pinned host buffers of the end-to-end leg are
    # first-touched in local memory (8 ranks move 16 GiB per step through host DRAM)
    numa = None
    try:
        import pynvml as nv
        nv.nvmlInit()
        hnd = nv.nvmlDeviceGetHandleByIndex(local_rank)
        nv.nvmlDeviceSetCpuAffinity(hnd)
        numa = len(os.sched_getaffinity(0))
    except Exception:
        numa = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    import oqb_b200 as Q
    ctx = Q.get_context(local_rank)
    lib = ctx.lib

    # ---- operators (host closures evaluated on the host, everything else on the device) ----
    Hd, Hp, coup = build_c3()
    H = Q.DenseHamiltonian([fA, fB], [Hd, Hp], unit="h", ctx=ctx)
    dav = Q.DaviesGenerator(Q.ConstantCouplings(coup, unit="h"), Q.Ohmic(ETA, FC, TMK), Q.ZERO_LAMBSHIFT)
    op = Q.DiffEqLiouvillian(H, [dav], [], N)
    p = Q.ODEParams(H, TF)
    t_setup0 = time.perf_counter()
    op._prepare_ame(S_POINT)  # eigen-system (host LAPACK), gap groups, γ tables → device (shared by the batch)
    ctx.sync()
    setup_s = time.perf_counter() - t_setup0

    # ---- synthetic states: 64 random density matrices per rank, generated on the device ----
    gen = torch.Generator(device=f"cuda:{local_rank}").manual_seed(3000 + rank)
    g = torch.randn((BATCH, N, 32), dtype=torch.complex128, device=f"cuda:{local_rank}", generator=gen)
    rho = g @ g.conj().transpose(1, 2)
    rho = rho / torch.diagonal(rho, dim1=1, dim2=2).sum(-1).real[:, None, None]
    u = Q.DeviceBatch(BATCH, N, N, ctx, rho.transpose(1, 2).contiguous())  # column-major blocks
    du = u.zeros_like()
    del g, rho

    def step():
        ctx.check(lib.oqb_ame_rhs(op._ame, BATCH, u.ptr, du.ptr))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    import ctypes as C
    l0 = ctx.launch_count()
    ctx.check(lib.oqb_profile_begin(ctx.h))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    zg_ms, zg_n, zg_fl = C.c_double(), C.c_uint64(), C.c_double()
    ctx.check(lib.oqb_profile_end(ctx.h, C.byref(zg_ms), C.byref(zg_n), C.byref(zg_fl)))
    launches = ctx.launch_count() - l0
    clocks = sampler.stop()
    tmax = torch.tensor([ms_total], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total_max = float(tmax.item())
    value = world * BATCH * args.steps / (ms_total_max * 1e-3)

    # ---- end to end: pinned host buffers, H2D + D2H inside the call ----
    uh = torch.empty((BATCH, N, N), dtype=torch.complex128, pin_memory=True)
    duh = torch.empty((BATCH, N, N), dtype=torch.complex128, pin_memory=True)
    uh.copy_(u.t)
    torch.cuda.synchronize()
    e2e_steps = max(2, min(args.steps, 5))
    ctx.check(lib.oqb_ame_rhs_host(op._ame, BATCH, C.c_void_p(uh.data_ptr()), C.c_void_p(duh.data_ptr())))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.check(lib.oqb_ame_rhs_host(op._ame, BATCH, C.c_void_p(uh.data_ptr()), C.c_void_p(duh.data_ptr())))
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * BATCH * e2e_steps / float(te.item())
    e2e_check = float((duh.to(f"cuda:{local_rank}") - du.t).abs().max().item())  # same result as the resident path

    # ---- the same step with the GENERAL complex product (eigenvectors treated as complex), rank-local, untimed by the
    #      contract: C3's H(s) is real symmetric, so its eigenvectors are exactly real and the default path issues
    #      2 real products per complex product in the four rotation GEMMs; a Hamiltonian with complex entries runs this ----
    v_real = C.c_int()
    ctx.check(lib.oqb_ame_eigenvectors_real(op._ame, C.byref(v_real)))
    real_path = bool(v_real.value) and ctx.real_operand_mode()
    general = None
    if real_path:
        du_real = du.t.clone()
        ctx.set_real_operand_mode(False)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gsteps = max(2, min(args.steps, 5))
        g0.record()
        for _ in range(gsteps):
            step()
        g1.record()
        torch.cuda.synchronize()
        gms = g0.elapsed_time(g1) / gsteps
        general = {"value_per_gpu": BATCH / (gms * 1e-3), "ms_per_step": gms,
                   "algorithmic_tflops": 32.0 * N ** 3 * BATCH / (gms * 1e-3) / 1e12,
                   "rel_diff_real_vs_general_product": float((du_real - du.t).norm().item() / du.t.norm().item()),
                   "what": "same workload with oqb_set_real_operand_mode(0): 3-multiplication complex product in all four rotations"}
        ctx.set_real_operand_mode(True)
        step()
        torch.cuda.synchronize()
        del du_real

    # ---- ensemble-average observable reduce (the only collective of the path), timed separately: per-GPU partial sums
    #      (oqb_expect_sum) and ncclAllReduce through the library's own communicator (oqb_comm_*, NCCL bound with dlopen), both on
    #      the context stream; median of 30 calls after 10 warm-up calls.  torch.distributed's all_reduce on the same
    #      buffer is timed beside it. ----
    obs = None
    if world > 1:
        idbuf = [None]
        if rank == 0:
            raw = (C.c_char * 128)()
            ctx.check(lib.oqb_comm_unique_id(raw))
            idbuf = [bytes(raw)]
        dist.broadcast_object_list(idbuf, src=0)
        comm = C.c_void_p()
        ctx.check(lib.oqb_comm_init_rank(ctx.h, world, rank, idbuf[0], C.byref(comm)))
        zs = np.stack([np.diag(np.diag(c)) for c in coup[:3]] + [np.eye(N)]).astype(np.complex128)
        d_obs = torch.as_tensor(zs, device=f"cuda:{local_rank}")  # diagonal: row- and column-major coincide
        d_out = torch.zeros(8, dtype=torch.float64, device=f"cuda:{local_rank}")

        def reduce_once(with_sum):
            if with_sum:
                ctx.check(lib.oqb_expect_sum(ctx.h, N, 4, C.c_void_p(d_obs.data_ptr()), BATCH, u.ptr, 0, C.c_void_p(d_out.data_ptr())))
            ctx.check(lib.oqb_allreduce_obs(comm, C.c_void_p(d_out.data_ptr()), 8))

        def med_ms(fn, reps=30, warm=10):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                dist.barrier()
                a0.record(); fn(); a1.record()
                torch.cuda.synchronize()
                ts.append(a0.elapsed_time(a1))
            return float(np.median(ts)), float(np.min(ts)), float(np.max(ts))
        m_all = med_ms(lambda: reduce_once(False))
        m_sum = med_ms(lambda: reduce_once(True))
        m_torch = med_ms(lambda: dist.all_reduce(d_out))
        reduce_once(True)
        torch.cuda.synchronize()
        ref = Q.expect(list(zs), u).sum(dim=0)  # per-rank partial sums by the per-member kernel
        dist.all_reduce(ref)
        got = torch.view_as_complex(d_out.reshape(4, 2))
        obs = {"what": "8 doubles (tr(Z_1 rho), tr(Z_2 rho), tr(Z_3 rho), tr rho summed over the ensemble)",
               "allreduce_ms_median": m_all[0], "allreduce_ms_min_max": m_all[1:],
               "expect_sum_plus_allreduce_ms_median": m_sum[0],
               "torch_distributed_allreduce_ms_median": m_torch[0], "calls": 30, "warmup_calls": 10,
               "rel_diff_vs_torch_path": float((got - ref).abs().max().item() / ref.abs().max().item())}
        ctx.check(lib.oqb_comm_destroy(comm))

    # ---- what the host can feed: every rank copies pinned host <-> device in both directions AT THE SAME TIME (the
    #      end-to-end leg moves 1 GiB each way per rank per step); e2e is then readable as a fraction of this link ----
    link = None
    try:
        nbytes = 1 << 28  # 256 MiB per direction, 8 back-to-back copies: keeps the pinned footprint small at 8 ranks
        hb_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        hb_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        db_in = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{local_rank}")
        db_out = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{local_rank}")
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
        reps = 8
        for it in range(2):  # first pass warms up
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                with torch.cuda.stream(s_up):
                    db_in.copy_(hb_in, non_blocking=True)
                with torch.cuda.stream(s_dn):
                    hb_out.copy_(db_out, non_blocking=True)
            torch.cuda.synchronize()
            dt_link = time.perf_counter() - t0
        per_dir = reps * nbytes / dt_link / 1e9
        agg = torch.tensor([per_dir, dt_link], dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        link = {"what": "concurrent pinned H2D + D2H (8 x 256 MiB per direction) on every rank at once (torch copy_ on two streams), GB/s per direction",
                "this_rank_gbs_per_direction": per_dir, "all_ranks_sum_gbs_per_direction": float(agg[0].item()),
                "e2e_bytes_per_direction_per_s_GB": e2e_val * N * N * 16 / 1e9,
                "e2e_fraction_of_link": e2e_val * N * N * 16 / 1e9 / float(agg[0].item())}
        del hb_in, hb_out, db_in, db_out
    except Exception as e:  # noqa: BLE001
        link = {"error": f"{type(e).__name__}: {e}"}

    # ---- like for like with the CPU arm: ONE C call per step does what the reference does per call — H(s) assembly, device
    #      eigensolver, GapIndices, Davies tables, rotations — for a NEW s every step, with HOST buffers (oqb_liouv_rhs_host) ----
    lfl = None
    try:
        opL = Q.DiffEqLiouvillian(H, [dav], [], N)
        opL.library = True
        opL.plan_cache = 2
        Lh = opL._liouv()
        cf0 = np.ascontiguousarray([[fA(S_POINT), fB(S_POINT)]])
        ctx.check(lib.oqb_liouv_rhs_host(Lh, BATCH, cf0.ctypes.data_as(C.POINTER(C.c_double)), 1, C.c_void_p(uh.data_ptr()), C.c_void_p(duh.data_ptr())))
        lfl_diff = float((duh.to(f"cuda:{local_rank}") - du.t).norm().item() / du.t.norm().item())
        for k in range(2):  # untimed: both slots of the plan cache get their handle (couplings uploaded once per handle)
            cfw = np.ascontiguousarray([[fA(S_POINT - 0.02 * (k + 1)), fB(S_POINT - 0.02 * (k + 1))]])
            ctx.check(lib.oqb_liouv_rhs_host(Lh, BATCH, cfw.ctypes.data_as(C.POINTER(C.c_double)), 1, C.c_void_p(uh.data_ptr()), C.c_void_p(duh.data_ptr())))
        lsteps = max(2, min(args.steps, 5))
        lsteps = max(lsteps, 8)
        barrier()
        t0 = time.perf_counter()
        per_step = []
        for k in range(lsteps):
            sk = S_POINT + 0.013 * (k + 1)
            cfk = np.ascontiguousarray([[fA(sk), fB(sk)]])
            t1 = time.perf_counter()
            ctx.check(lib.oqb_liouv_rhs_host(Lh, BATCH, cfk.ctypes.data_as(C.POINTER(C.c_double)), 1, C.c_void_p(uh.data_ptr()), C.c_void_p(duh.data_ptr())))
            per_step.append((time.perf_counter() - t1) * 1e3)  # the call synchronises itself (host buffers)
        torch.cuda.synchronize()
        tl = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            dist.all_reduce(tl, op=dist.ReduceOp.MAX)
        built, hits, prep_ms, eig_ms = opL.liouv_stats()
        lfl = {"what": "e2e with a NEW eigen-system every step, everything inside one C call (oqb_liouv_rhs_host): H(s) assembly, device "
                       "eigensolver (cuSOLVER Dsyevd via dlopen), GapIndices + Ohmic tables on the device, rotations, host buffers — "
                       "the work the reference's CPU path does per call",
               "value": world * BATCH * lsteps / float(tl.item()), "unit": UNIT, "steps": lsteps,
               "ms_per_step_this_rank": [round(x, 2) for x in per_step], "ms_per_step_median_this_rank": float(np.median(per_step)),
               "plan_prepare_ms_last": prep_ms, "eigensolver_ms_last": eig_ms, "plans_built": built,
               "rel_diff_vs_host_lapack_plan_at_s0": lfl_diff}
        opL.clear_plans()
    except Exception as e:  # noqa: BLE001
        lfl = {"error": f"{type(e).__name__}: {e}"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- FP64 ceilings measured in this run (MEASURED_PEAKS.json only has bf16 and HBM) ----
    a = torch.randn((8192, 8192), dtype=torch.float64, device=f"cuda:{local_rank}")
    b = torch.randn((8192, 8192), dtype=torch.float64, device=f"cuda:{local_rank}")
    best = 1e9
    for _ in range(6):
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record(); torch.matmul(a, b); x1.record(); torch.cuda.synchronize()
        best = min(best, x0.elapsed_time(x1))
    dgemm_tf = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
    del a, b
    za = torch.randn((4096, 4096), dtype=torch.complex128, device=f"cuda:{local_rank}")
    best = 1e9
    for _ in range(6):
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record(); torch.matmul(za, za); x1.record(); torch.cuda.synchronize()
        best = min(best, x0.elapsed_time(x1))
    zgemm_tf = 8 * 4096 ** 3 / (best * 1e-3) / 1e12
    del za
    dmma_tf, dfma_tf = ctx.probe_fp64()

    alg_tflops = zg_fl.value / (zg_ms.value * 1e-3) / 1e12 if zg_ms.value > 0 else None  # 8·M·N·K per complex product
    peak = max(dgemm_tf, dmma_tf)
    mode = C.c_int()
    ctx.check(lib.oqb_get_zgemm_mode(ctx.h, C.byref(mode)))
    config["eigenvectors"] = ("exactly real (H(s) of this config is real symmetric): the four rotation products skip the zero "
                              "imaginary part of V — the structure row; the contract row (general complex product) is "
                              "roofline.contract_row" if real_path else "complex (general product)")
    common = {"bound": "tensor", "peak": peak, "unit": "TFLOP/s",
              "peak_source": "measured in this run: max(cuBLAS DGEMM 8192^3 best-of-6, DMMA issue probe); MEASURED_PEAKS.json has no FP64 figure",
              "cublas_dgemm_tflops": dgemm_tf, "cublas_zgemm_4096_tflops": zgemm_tf, "dmma_probe_tflops": dmma_tf,
              "dfma_probe_tflops": dfma_tf}
    traffic_profile = {"value_bytes_per_launch": 2.157e9 if real_path else (2.138e9 if mode.value == 1 else 2.163e9),
                       "from_profile_not_this_run": ("profiles/r1_zgemm_real_ncu.md" if real_path else
                                                     ("profiles/r1_zgemm_3m_ncu.md" if mode.value == 1 else "profiles/r1_zgemm_tma_ncu.md")),
                       "algorithmic_bytes_per_launch": 2 * BATCH * N * N * 16 + N * N * 16}
    kern = {"kernel_ms_per_launch": zg_ms.value / max(zg_n.value, 1), "kernel_launches_timed": zg_n.value,
            "kernel_share_of_step": zg_ms.value / ms_total if ms_total > 0 else None}
    if real_path:
        # STRUCTURE ROW (an exploited structure is its own row, never divided into the dense flop count): with an
        # exactly real eigenvector factor a complex product is 2 real products, so this row's algorithmic work is 4·M·N·K
        # flops per product (16 n³ per ρ) and frac is the FP64 tensor pipe's utilisation.
        work = 0.5 * zg_fl.value
        ach = work / (zg_ms.value * 1e-3) / 1e12 if zg_ms.value > 0 else None
        roofline = {**common, "row": "structure: real eigenvectors (real-symmetric H), 16 n^3 flops per rho",
                    "kernel": "zgemm_tma_kernel real-operand variant (FP64 DMMA.8x8x4, TMA-staged, 128x64 CTA, 2 DMMAs per complex 8x8x4; "
                              "4 launches of 64 products per step)",
                    "achieved": ach, "frac": ach / peak if ach else None,
                    "algorithmic_flops_per_launch": work / max(zg_n.value, 1),
                    "speedup_vs_contract_row": (general["ms_per_step"] / (ms_total / args.steps)) if general else None,
                    "tflops_if_counted_as_8MNK": alg_tflops, "traffic": None, "traffic_profile": traffic_profile, **kern,
                    "contract_row": None if general is None else {
                        "row": "contract: general complex product (any complex H), 32 n^3 algorithmic flops per rho",
                        "kernel": "zgemm_tma_kernel, 3-multiplication complex product" if mode.value == 1 else "zgemm_tma_kernel, 4-multiplication complex product",
                        "value_per_gpu": general["value_per_gpu"], "ms_per_step": general["ms_per_step"],
                        "achieved": general["algorithmic_tflops"], "peak": peak, "unit": "TFLOP/s",
                        "frac": general["algorithmic_tflops"] / peak,
                        "executed_over_algorithmic_flops": 0.75 if mode.value == 1 else 1.0,
                        "frac_executed": general["algorithmic_tflops"] * (0.75 if mode.value == 1 else 1.0) / peak,
                        "frac_note": "achieved counts the ALGORITHMIC 8·M·N·K flops per complex product; the 3-multiplication kernel issues "
                                     "6·M·N·K on the tensor pipe, so frac may exceed 1 — frac_executed is the pipe utilisation",
                        "rel_diff_vs_structure_row_result": general["rel_diff_real_vs_general_product"]}}
    else:
        exec_ratio = 0.75 if mode.value == 1 else 1.0
        roofline = {**common, "row": "contract: general complex product, 32 n^3 algorithmic flops per rho",
                    "kernel": "zgemm_tma_kernel (FP64 DMMA.8x8x4, TMA-staged, %s; 4 launches of 64 products per step)"
                              % ("3-multiplication complex product" if mode.value == 1 else "4-multiplication complex product"),
                    "achieved": alg_tflops, "frac": alg_tflops / peak if alg_tflops else None,
                    "executed_over_algorithmic_flops": exec_ratio, "frac_executed": alg_tflops * exec_ratio / peak if alg_tflops else None,
                    "algorithmic_flops_per_launch": zg_fl.value / max(zg_n.value, 1), "traffic": None, "traffic_profile": traffic_profile, **kern}

    # ---- cost of a NEW eigen-system per call (time-dependent H: the reference diagonalises in every RHS call,
    #      src/opensys/diffeq_liouvillian.jl:94-96); shared by the whole batch, reported next to the resident number ----
    prep = {}
    op.plan_cache = 1  # steady state of a time-dependent solve: every call brings a new s, the plan (and its handle) is recycled
    for mode in ("host_lapack", "device_eigh", "device_eigh_and_tables", "library_one_call"):
        op.device_eigh = mode != "host_lapack"
        op.device_tables = mode == "device_eigh_and_tables"
        op.library = mode == "library_one_call"
        op.clear_plans()
        best = 1e9
        for k in range(3):
            s_k = S_POINT + 0.01 * (k + 1) + {"host_lapack": 0.0, "device_eigh": 0.003, "device_eigh_and_tables": 0.006, "library_one_call": 0.008}[mode]
            t0 = time.perf_counter()
            op._prepare_ame(s_k)
            ctx.sync()
            best = min(best, time.perf_counter() - t0)
        prep[mode + "_s"] = best
        prep["evals_per_sec_incl_" + mode] = BATCH / (ms_total / args.steps * 1e-3 + best)
    # the device-built tables give the same RHS as the host-built ones (same s, same device eigen-system is not
    # guaranteed across two eigh calls, so compare through host LAPACK's (w, v) for both)
    w_chk, v_chk = Q.haml_eigs(H, S_POINT, N)
    chk = []
    op.library = False
    for flag in (False, True):
        op.device_eigh, op.device_tables = False, flag
        op.clear_plans()
        chk.append(op.rhs_with_eig(du.zeros_like(), u, p, S_POINT * TF, w_chk, v_chk).t.clone())
    op.library = False
    prep["device_tables_rel_diff_vs_host_tables"] = float((chk[0] - chk[1]).norm().item() / chk[0].norm().item())
    del chk
    op.device_tables = False
    op.plan_cache = 8
    op.clear_plans()
    op._prepare_ame(S_POINT)
    prep["what"] = ("haml_eigs (H(s) assembly + Hermitian eigensolver) + GapIndices grouping + γ/S tables + coupling rotation for one "
                    "new s, shared by the 64-ρ batch; host_lapack = scipy zheevd + numpy grouping/tables on the host cores; "
                    "device_eigh = oqb_dense_hamiltonian + cuSOLVER (torch.linalg.eigh) + oqb_ame_set_eigen_device, tables on the host; "
                    "device_eigh_and_tables = additionally the rounded-gap/Ohmic tables and the gap grouping on the device "
                    "(oqb_ame_gaps_build + oqb_ame_add_davies_ohmic: grouping, cross-term lists and tables all on the device); "
                    "library_one_call = oqb_liouv_prepare: the same with the eigensolver bound inside the library (dlopen cuSOLVER) and "
                    "no host-language step at all")
    op.device_eigh = False

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cval, cms, cores = cpu_reference_arm(2, 1)
        from oracle import cpu_baseline as CB
        cpu = {"value": cval, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "2 rho through the reference-structured per-call AME RHS (eigh + rotations + closed-form Davies), numpy/OpenBLAS; see oracle/cpu_baseline.py",
               "literal_davies_loop": CB.time_literal_davies_loop(),
               "julia_script": "baseline/julia_ref.jl runs the real package for anyone with Julia >= 1.9 (not installed here)"}

    # ---- the other BASELINE.json configurations, measured in the same driver run (N = 1 only; short) ----
    secondary = None
    if world == 1 and not args.no_secondary:
        secondary = run_secondary(Q, torch, ctx)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 (complex128)", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": BATCH * N * N * 16, "d2h_bytes_per_step": BATCH * N * N * 16,
                    "steps": e2e_steps, "max_abs_diff_vs_resident": e2e_check},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "setup_s_not_timed": setup_s, "cpus_bound_to_gpu_numa_node": numa, "per_call_eigensystem": prep, "like_for_like_e2e": lfl, "host_link": link,
            "obs_allreduce": obs, "secondary": secondary}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
