#!/usr/bin/env python
"""C4 across GPUs: the trajectory ensemble is sharded into contiguous blocks, one per rank (one process per
GPU, no data-path collective); every rank steps its block on the device (oqb_traj_evolve: RK4 + jump loop), reduces
the diagonal observables ⟨Z_k⟩ and the norm of its members (oqb_ensemble_population: one pass over ψ) and ONE NCCL all-reduce of
nobs + 1 numbers forms the ensemble average.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/ensemble_multi_gpu.py
  python benchmarks/ensemble_multi_gpu.py --per-gpu 65536          # single GPU

Weak scaling: --per-gpu trajectories (default 65536 = C4) on every rank.  Rank 0 prints one JSON line.
"""
import argparse
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
    ap.add_argument("--per-gpu", type=int, default=65536)
    ap.add_argument("--qubits", type=int, default=12)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--gamma", type=float, default=0.05)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    import oqb_b200 as Q
    from oqb_b200.sharding import ensemble_average, shard_bounds
    from other_configs import tfim_sparse_c4
    nq, B = args.qubits, args.per_gpu
    n = 1 << nq
    lo, hi = shard_bounds(B * world, rank, world)  # this rank's block of the global ensemble
    ctx = Q.get_context(local)
    Hx, Hz, Ls = tfim_sparse_c4(nq)
    H = Q.SparseHamiltonian([lambda s: 1 - s, lambda s: s], [Hx, Hz], unit="ħ", ctx=ctx)
    op = Q.DiffEqLiouvillian(H, [], [Q.LindbladLiouvillian([Q.Lindblad(args.gamma, L) for L in Ls])], n)
    ens = Q.TrajectoryEnsemble(op)
    gen = torch.Generator(device=f"cuda:{local}").manual_seed(1000 + rank)
    g = torch.randn((hi - lo, 1, n), dtype=torch.complex128, device=f"cuda:{local}", generator=gen)
    psi = Q.DeviceBatch(hi - lo, n, 1, ctx, (g / torch.linalg.vector_norm(g, dim=2, keepdim=True)).contiguous())
    del g
    st = Q.TrajectoryStepper(ens, hi - lo, seed=7 + 1000003 * rank)  # distinct xoshiro streams per rank
    p = Q.ODEParams(None, 10.0)
    idx = np.arange(n)
    zdiag = [1.0 - 2.0 * ((idx >> (nq - 1 - q)) & 1) for q in range(min(nq, 15))] + [np.ones(n)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    st.evolve(psi, p, 0.0, 0.01, 2)  # warm-up
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st.evolve(psi, p, 0.02, 0.01, args.steps)
    e1.record()
    barrier()
    tms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    # observables: per-GPU partial sums, then the one collective of the path
    # diagonal ensemble averages: ONE pass over ψ gives the populations P[r] = Σ_b |ψ_b[r]|², every ⟨Z_k⟩ is a dot with P
    zd = torch.as_tensor(np.stack(zdiag), device=f"cuda:{local}")
    ensemble_average(zd @ Q.ensemble_population(psi), hi - lo)  # warm-up (allocator, NCCL channel set-up)
    torch.cuda.synchronize()
    a0, a1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
    a0.record()
    part = zd @ Q.ensemble_population(psi)
    a1.record()
    torch.cuda.synchronize()
    # the collective itself through the LIBRARY's communicator (oqb_allreduce_obs: ncclAllReduce on the context stream,
    # csrc/comm.cu), torch.distributed on the same numbers beside it; median of 20 warmed calls
    import ctypes as C
    buf = torch.cat([part.to(torch.float64), torch.tensor([float(hi - lo)], dtype=torch.float64, device=part.device)]).contiguous()
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
                b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                dist.barrier()
                b0.record(); fn(); b1.record()
                torch.cuda.synchronize()
                ts.append(b0.elapsed_time(b1))
            return float(np.median(ts))
        scratch = buf.clone()
        lib_ms = med(lambda: ctx.check(ctx.lib.oqb_allreduce_obs(comm, C.c_void_p(scratch.data_ptr()), scratch.numel())))
        torch_ms = med(lambda: dist.all_reduce(scratch))
        red = buf.clone()
        ctx.check(ctx.lib.oqb_allreduce_obs(comm, C.c_void_p(red.data_ptr()), red.numel()))
        torch.cuda.synchronize()
        ctx.check(ctx.lib.oqb_comm_destroy(comm))
    else:
        red = buf
    avg = red[:-1] / red[-1]
    jumps = st.njumps.sum().to(torch.float64)
    if world > 1:
        dist.all_reduce(jumps)
    if rank == 0:
        ms = float(tms.item())
        hbm = 6453.1
        try:
            hbm = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            pass
        gb = 16 * 16.0 * n * B / 1e9  # 16 vector passes of one fused RK4 step, per GPU
        real_stdout.write(json.dumps({
            "config": "C4-multi-gpu", "workload": f"{nq}-qubit sparse TFIM trajectories, {B} psi per GPU x {world} GPUs, RK4 + jump loop on the device",
            "metric": "trajectory_steps_per_sec", "value": world * B * args.steps / (ms * 1e-3), "n_gpus": world, "scaling": "weak",
            "ms_per_rk4_step": ms / args.steps, "per_gpu_GBs": gb / (ms / args.steps * 1e-3),
            "per_gpu_frac_of_measured_hbm": gb / (ms / args.steps * 1e-3) / hbm,
            "observables": {"n": len(zdiag), "local_reduce_ms": a0.elapsed_time(a1), "allreduce_ms": lib_ms, "torch_distributed_allreduce_ms": torch_ms,
                            "collective": "one oqb_allreduce_obs (ncclAllReduce through the library's communicator) of nobs+1 doubles, median of 20 calls" if world > 1 else "none (1 GPU)",
                            "mean_norm2": float(avg[-1]), "mean_Z1": float(avg[0])},
            "jumps_total": float(jumps.item())}) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
