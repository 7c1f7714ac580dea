"""Scratch: the C4-shaped ensemble matvec through the sliced-ELL kernel, random pattern and TFIM-forced, per option value."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import oqb_b200 as Q  # noqa: E402
from benchmarks.other_configs import random_sparse_c4, tfim_sparse_c4, rand_states, peaks  # noqa: E402

nq, n, B = 12, 4096, 65536
ctx = Q.get_context()
hbm, _ = peaks()
psi_t = rand_states(torch, B, n, 1, 4000)
psi = Q.DeviceBatch(B, n, 1, ctx, psi_t.contiguous())
dpsi = psi.zeros_like()
rng = np.random.default_rng(4000)
s = rng.random(B)
cf = np.ascontiguousarray(np.stack([1 - s, s], axis=1))
gam = np.full((1, nq), 0.05)
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
for name in sys.argv[1].split(","):
    if name == "random":
        Ha, Hb = random_sparse_c4(nq)
        _, _, Ls = tfim_sparse_c4(nq)
    else:
        Ha, Hb, Ls = tfim_sparse_c4(nq)
    for opts in [dict(), dict(sell_r=3)]:
        import contextlib
        H = Q.SparseHamiltonian([lambda x: 1 - x, lambda x: x], [Ha, Hb], unit="ħ")
        op = Q.DiffEqLiouvillian(H, [], [Q.LindbladLiouvillian([Q.Lindblad(0.05, L) for L in Ls])], n)
        ens = Q.TrajectoryEnsemble(op)
        mv = lambda: ctx.check(ctx.lib.oqb_traj_matvec(ens.sp, B, dp(cf), 0, dp(gam), 1, psi.ptr, dpsi.ptr))
        with contextlib.ExitStack() as st:
            st.enter_context(ctx.option("traj_kernel", 1))
            for k, v in opts.items():
                st.enter_context(ctx.option(k, v))
            for _ in range(3):
                mv()
            torch.cuda.synchronize()
            ctx.check(ctx.lib.oqb_profile_begin(ctx.h))
            for _ in range(5):
                mv()
            a, b_, w = C.c_double(), C.c_uint64(), C.c_double()
            ctx.check(ctx.lib.oqb_profile_end(ctx.h, C.byref(a), C.byref(b_), C.byref(w)))
        kms = a.value / max(b_.value, 1)
        pl, pad = C.c_int(), C.c_double()
        ctx.check(ctx.lib.oqb_traj_sell_stats(ens.sp, C.byref(pl), C.byref(pad)))
        print(json.dumps({"pattern": name, **opts, "kernel_ms": round(kms, 3), "frac": round(32.0 * n * B / 1e9 / (kms * 1e-3) / hbm, 4),
                          "plans": pl.value, "padding": round(pad.value, 3)}), flush=True)
