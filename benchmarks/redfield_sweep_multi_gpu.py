#!/usr/bin/env python
"""C5 as BASELINE.json states it: an annealing-schedule sweep — 8192 schedules × 7-qubit Redfield RHS — sharded over the
GPUs of one box (1024 schedules per GPU, contiguous blocks, no data-path collective), with the per-schedule observables
⟨Z_α⟩ and tr ρ gathered by ONE NCCL all_gather per output point.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/redfield_sweep_multi_gpu.py

Per step and schedule: du = −i[H(s_b), ρ_b] − Σ_α (K₂ + K₂†), K₂ = S_αΛ_{bα}ρ_b − Λ_{bα}ρ_bS_α with the schedule's own Λ
(supplied; the device Λ build is measured by benchmarks/lambda_bench.py).  Rank 0 prints one JSON line.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-gpu", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    import oqb_b200 as Q
    from oqb_b200.sharding import shard_bounds
    from other_configs import tfim_dense
    nq, n, K, B = 7, 128, 7, args.per_gpu
    total = B * world
    lo, hi = shard_bounds(total, rank, world)
    ctx = Q.get_context(local)
    Hd, Hp, z = tfim_dense(nq)
    S = [np.diag(zq).astype(complex) for zq in z]
    H = Q.DenseHamiltonian([lambda s: 1 - s, lambda s: s], [Hd, Hp], unit="ħ", ctx=ctx)
    gen = torch.Generator(device=dev).manual_seed(5000 + rank)
    g = torch.randn((B, 16, n), dtype=torch.complex128, device=dev, generator=gen)
    rho = g.transpose(1, 2) @ g.conj()
    rho = rho / torch.diagonal(rho, dim1=1, dim2=2).sum(-1).real[:, None, None]
    u = Q.DeviceBatch(B, n, n, ctx, rho.contiguous())
    du = u.zeros_like()
    Lam = Q.DeviceBatch(B * K, n, n, ctx, torch.randn((B * K, n, n), dtype=torch.complex128, device=dev, generator=gen) / n)
    Sd = Q.DeviceBatch.from_host(np.stack(S), ctx)
    pw = 0.5 + 1.5 * np.arange(lo, hi) / total  # schedule b: s_b(t) = (t/tf)^{p_b}
    s = 0.37 ** pw
    obs = S + [np.eye(n, dtype=complex)]

    diag_path = os.environ.get("OQB_REDFIELD_DIAG", "1") != "0"  # S_α = Z_α: elementwise form of S_αM − MS_α
    Sdiag = Q.DeviceBatch.from_host(np.stack([np.diag(x) for x in S]), ctx)

    def step():
        H.commutator(du, u, s)
        if diag_path:
            ctx.check(ctx.lib.oqb_redfield_apply_diag_acc(ctx.h, n, K, Sdiag.ptr, Lam.ptr, B, u.ptr, du.ptr))
        else:
            ctx.check(ctx.lib.oqb_redfield_apply_acc(ctx.h, n, K, Sd.ptr, 1 | 2, Lam.ptr, B, u.ptr, du.ptr))  # bit 1: the Z_α couplings are real

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(3):
        step()
    Q.expect(obs, u)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    tms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    # per-schedule observables of this rank's block, then the one collective of the path: an all-gather into the full sweep
    # through the LIBRARY's communicator (oqb_allgather_obs: ncclAllGather on the context stream, csrc/comm.cu);
    # torch.distributed on the same buffers is timed beside it.  Median of 20 warmed calls.
    import ctypes as C
    ev = Q.expect(obs, u)  # (B, K+1) complex
    loc = torch.view_as_real(ev).contiguous()
    full = torch.empty((world,) + tuple(loc.shape), dtype=loc.dtype, device=dev)
    full_t = torch.empty_like(full)
    lib_ms = torch_ms = 0.0
    if world > 1:
        idbuf = [None]
        if rank == 0:
            raw = (C.c_char * 128)()
            ctx.check(ctx.lib.oqb_comm_unique_id(raw))
            idbuf = [bytes(raw)]
        dist.broadcast_object_list(idbuf, src=0)
        comm = C.c_void_p()
        ctx.check(ctx.lib.oqb_comm_init_rank(ctx.h, world, rank, idbuf[0], C.byref(comm)))

        def med(fn, reps=20, warm=5):
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
            return float(np.median(ts))
        lib_ms = med(lambda: ctx.check(ctx.lib.oqb_allgather_obs(comm, C.c_void_p(loc.data_ptr()), C.c_void_p(full.data_ptr()), loc.numel())))
        torch_ms = med(lambda: dist.all_gather_into_tensor(full_t, loc))
        assert torch.equal(full, full_t)
        ctx.check(ctx.lib.oqb_comm_destroy(comm))
    else:
        full[0].copy_(loc)
    if rank == 0:
        ms = float(tms.item()) / args.steps
        flops = 8.0 * n ** 3 * (2 + (K if diag_path else 3 * K)) * total
        tr = torch.view_as_complex(full)[..., -1].real
        real_stdout.write(json.dumps({
            "config": "C5" + ("-diagS" if diag_path else ""), "diagonal_couplings_exploited": diag_path, "workload": f"{total} schedules x 7-qubit Redfield RHS (commutator + apply, supplied Lambda), n={n}, K={K}, {B} per GPU x {world} GPUs",
            "metric": "rho_rhs_evals_per_sec", "value": total / (ms * 1e-3), "n_gpus": world, "scaling": "weak", "ms_per_step": ms,
            "algorithmic_tflop_per_step": flops / 1e12, "aggregate_tflops_algorithmic": flops / (ms * 1e-3) / 1e12,
            "observables": {"per_schedule": K + 1, "allgather_ms": lib_ms, "torch_distributed_allgather_ms": torch_ms, "gathered_shape": list(full.shape),
                            "collective": "one oqb_allgather_obs (ncclAllGather through the library's communicator) of (K+1) complex numbers per schedule, median of 20 calls" if world > 1 else "none (1 GPU)",
                            "trace_min": float(tr.min().item()), "trace_max": float(tr.max().item())}}) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
