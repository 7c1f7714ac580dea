"""Scratch: on-device RK4 stepping (oqb_traj_evolve) of the C4-shaped ensemble with an unstructured SparseHamiltonian:
fused (stage arithmetic in traj_sell_kernel's store) against separate stage kernels."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oqb_b200 as Q
from benchmarks.other_configs import random_sparse_c4, tfim_sparse_c4, rand_states, timed
nq, n, B = 12, 4096, 65536
ctx = Q.get_context()
psi_t = rand_states(torch, B, n, 1, 4000)
psi_t = psi_t / torch.linalg.vector_norm(psi_t, dim=2, keepdim=True)
for name in ("random", "tfim_forced"):
    if name == "random":
        Ha, Hb = random_sparse_c4(nq); _, _, Ls = tfim_sparse_c4(nq)
    else:
        Ha, Hb, Ls = tfim_sparse_c4(nq)
    for fuse in (1, 0):
        H = Q.SparseHamiltonian([lambda x: 1 - x, lambda x: x], [Ha, Hb], unit="ħ")
        op = Q.DiffEqLiouvillian(H, [], [Q.LindbladLiouvillian([Q.Lindblad(0.05, L) for L in Ls])], n)
        ens = Q.TrajectoryEnsemble(op)
        psi = Q.DeviceBatch(B, n, 1, ctx, psi_t.clone().contiguous())
        st = Q.TrajectoryStepper(ens, B, seed=1)
        pO = Q.ODEParams(None, 10.0)
        with ctx.option("traj_kernel", 1), ctx.option("rk_fuse", fuse):
            st.evolve(psi, pO, 0.0, 0.01, 2)
            l0 = ctx.launch_count()
            ms = timed(lambda: st.evolve(psi, pO, 0.0, 0.01, 3), 2, 1, torch) / 3
            l1 = ctx.launch_count()
        print(json.dumps({"operator": name, "rk_fuse": fuse, "ms_per_rk4_step": round(ms, 3), "traj_steps_per_sec": B / (ms * 1e-3),
                          "launches_per_step": (l1 - l0) / 9.0}), flush=True)
        del st, psi, ens
        torch.cuda.empty_cache()
