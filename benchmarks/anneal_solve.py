#!/usr/bin/env python
"""A whole solve on the device: fixed-step RK4 of the 10-qubit AME master equation with a TIME-DEPENDENT Hamiltonian
(s = t/tf changes at every stage, so every stage time needs its own eigen-system — what the reference does in every
RHS call, src/opensys/diffeq_liouvillian.jl:94-96) for a batch of 64 ρ that never leave HBM.

  python benchmarks/anneal_solve.py [--steps 5]

Per RK4 step: 4 RHS evaluations (4 batched ZGEMM launches each) + 3 new eigen-systems (device eigensolver + device
gap grouping/tables).  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402  (operator builders of the headline config)
import oqb_b200 as Q  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    ctx = Q.get_context()
    N, B = bench.N, args.batch
    Hd, Hp, coup = bench.build_c3()
    H = Q.DenseHamiltonian([bench.fA, bench.fB], [Hd, Hp], unit="h", ctx=ctx)
    dav = Q.DaviesGenerator(Q.ConstantCouplings(coup, unit="h"), Q.Ohmic(bench.ETA, bench.FC, bench.TMK), Q.ZERO_LAMBSHIFT)
    op = Q.DiffEqLiouvillian(H, [dav], [], N)
    op.device_eigh, op.device_tables, op.plan_cache = True, True, 3
    p = Q.ODEParams(H, bench.TF)
    gen = torch.Generator(device="cuda").manual_seed(11)
    g = torch.randn((B, N, 32), dtype=torch.complex128, device="cuda", generator=gen)
    rho = g @ g.conj().transpose(1, 2)
    rho = rho / torch.diagonal(rho, dim1=1, dim2=2).sum(-1).real[:, None, None]
    u = Q.DeviceBatch(B, N, N, ctx, rho.transpose(1, 2).contiguous())
    st = Q.DensityStepper(op)
    dt = 0.01
    u0 = u.t.clone()
    out, finals = {}, {}
    for mode in ("sequential", "prefetch"):
        # prefetch: the eigen-systems of the coming stage times are prepared on a high-priority side stream (own context,
        # worker thread) while the current stages' GEMMs run
        op.prefetch_stages = mode == "prefetch"
        op.plan_cache = 3 if mode == "sequential" else 6
        op.clear_plans()
        u.t.copy_(u0)
        st.evolve(u, p, 1.0, dt, 1)  # warm-up
        ctx.sync()
        l0, b0 = ctx.launch_count(), op.plans_built
        t0 = time.perf_counter()
        st.evolve(u, p, 1.0 + dt, dt, args.steps)
        ctx.sync()
        wall = time.perf_counter() - t0
        finals[mode] = u.t.clone()
        out[mode] = {"ms_per_rk4_step": wall / args.steps * 1e3, "rk4_steps_per_sec": args.steps / wall,
                     "rho_rhs_evals_per_sec_incl_eigensystems": 4 * B * args.steps / wall,
                     "gpu_launches_per_step_main_context": (ctx.launch_count() - l0) / args.steps,
                     "new_eigensystems_per_step": (op.plans_built - b0) / args.steps}
    op.clear_plans()
    tr = torch.diagonal(u.t, dim1=1, dim2=2).sum(-1)
    herm = float((u.t - u.t.conj().transpose(1, 2)).norm().item() / u.t.norm().item())
    diff = float((finals["prefetch"] - finals["sequential"]).norm().item() / finals["sequential"].norm().item())
    best = min(out.values(), key=lambda m: m["ms_per_rk4_step"])
    print(json.dumps({"config": "C3-solve", "workload": f"10-qubit AME, time-dependent H(s=t/tf), {B} rho, fixed-step RK4 on the device",
                      "metric": "rk4_steps_per_sec", "value": best["rk4_steps_per_sec"], "ms_per_rk4_step": best["ms_per_rk4_step"],
                      "rho_rhs_evals_per_sec_incl_eigensystems": best["rho_rhs_evals_per_sec_incl_eigensystems"],
                      "modes": out, "prefetch_vs_sequential_rel_diff": diff,
                      "trace_drift_max": float((tr - 1).abs().max().item()), "hermiticity_defect": herm}))


if __name__ == "__main__":
    main()
