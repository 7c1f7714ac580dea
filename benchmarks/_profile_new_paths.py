#!/usr/bin/env python
"""Small drivers of the §8(f) paths for `ncu --metrics gpu__time_duration.sum` launch lists (profiles/):
  python benchmarks/_profile_new_paths.py lambda   # 64 schedules x 7 couplings, n=128
  python benchmarks/_profile_new_paths.py evolve   # 8192 trajectories x 4096, 3 RK4 steps with jumps
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import oqb_b200 as Q  # noqa: E402


def lam():
    sys.argv = ["lambda_bench.py", "--schedules", "64", "--cpu-integrals", "0"]
    import lambda_bench
    lambda_bench.main()


def evolve():
    from other_configs import tfim_sparse_c4
    nq, B = 12, 8192
    n = 1 << nq
    Hx, Hz, Ls = tfim_sparse_c4(nq)
    H = Q.SparseHamiltonian([lambda s: 1 - s, lambda s: s], [Hx, Hz], unit="ħ")
    op = Q.DiffEqLiouvillian(H, [], [Q.LindbladLiouvillian([Q.Lindblad(2.0, L) for L in Ls])], n)
    ens = Q.TrajectoryEnsemble(op)
    ctx = Q.get_context()
    g = torch.randn((B, 1, n), dtype=torch.complex128, device="cuda")
    psi = Q.DeviceBatch(B, n, 1, ctx, (g / torch.linalg.vector_norm(g, dim=2, keepdim=True)).contiguous())
    st = Q.TrajectoryStepper(ens, B, seed=3)
    st.evolve(psi, Q.ODEParams(None, 10.0), 0.0, 0.05, 3)
    ctx.sync()
    print("jumps", int(st.njumps.sum().item()))


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    {"lambda": lam, "evolve": evolve}[sys.argv[1]]()
