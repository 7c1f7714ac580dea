"""Scratch: one configuration of the sliced-ELL kernel a few times (ncu target)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oqb_b200 as Q
from benchmarks.other_configs import random_sparse_c4, tfim_sparse_c4, rand_states
nq, n, B = 12, 4096, 65536
ctx = Q.get_context()
psi = Q.DeviceBatch(B, n, 1, ctx, rand_states(torch, B, n, 1, 4000).contiguous())
dpsi = psi.zeros_like()
s = np.random.default_rng(4000).random(B)
cf = np.ascontiguousarray(np.stack([1 - s, s], axis=1))
gam = np.full((1, nq), 0.05)
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
name, r = sys.argv[1], int(sys.argv[2])
if name == "random":
    Ha, Hb = random_sparse_c4(nq); _, _, Ls = tfim_sparse_c4(nq)
else:
    Ha, Hb, Ls = tfim_sparse_c4(nq)
H = Q.SparseHamiltonian([lambda x: 1 - x, lambda x: x], [Ha, Hb], unit="ħ")
op = Q.DiffEqLiouvillian(H, [], [Q.LindbladLiouvillian([Q.Lindblad(0.05, L) for L in Ls])], n)
ens = Q.TrajectoryEnsemble(op)
ctx.set_option("traj_kernel", 1); ctx.set_option("sell_r", r)
for _ in range(3):
    ctx.check(ctx.lib.oqb_traj_matvec(ens.sp, B, dp(cf), 0, dp(gam), 1, psi.ptr, dpsi.ptr))
torch.cuda.synchronize()
