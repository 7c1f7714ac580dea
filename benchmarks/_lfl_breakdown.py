"""Where the like-for-like step goes: wall time of oqb_liouv_prepare (new s), oqb_liouv_rhs_host with a cached plan, and both in one call."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import oqb_b200 as Q  # noqa: E402

ctx = Q.get_context()
lib = ctx.lib
Hd, Hp, coup = bench.build_c3()
H = Q.DenseHamiltonian([bench.fA, bench.fB], [Hd, Hp], unit="h", ctx=ctx)
dav = Q.DaviesGenerator(Q.ConstantCouplings(coup, unit="h"), Q.Ohmic(bench.ETA, bench.FC, bench.TMK), Q.ZERO_LAMBSHIFT)
op = Q.DiffEqLiouvillian(H, [dav], [], bench.N)
op.plan_cache = 2
L = op._liouv()
B, N = bench.BATCH, bench.N
uh = torch.randn((B, N, N), dtype=torch.complex128).pin_memory()
duh = torch.empty((B, N, N), dtype=torch.complex128, pin_memory=True)
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))


def row(s):
    return np.ascontiguousarray([[bench.fA(s), bench.fB(s)]])


def wall(fn, n=5):
    ts = []
    for k in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn(k)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return [round(t, 2) for t in ts]


plan = C.c_void_p()
cf0 = row(0.5)
lib.oqb_liouv_rhs_host(L, B, dp(cf0), 1, C.c_void_p(uh.data_ptr()), C.c_void_p(duh.data_ptr()))
print("prepare only, new s each call (ms):", wall(lambda k: ctx.check(lib.oqb_liouv_prepare(L, dp(row(0.31 + 0.01 * k)), C.byref(plan)))))
print("rhs_host, cached plan (ms):       ", wall(lambda k: ctx.check(lib.oqb_liouv_rhs_host(L, B, dp(row(0.31 + 0.01 * 4)), 1, C.c_void_p(uh.data_ptr()), C.c_void_p(duh.data_ptr())))))
print("rhs_host, new s each call (ms):   ", wall(lambda k: ctx.check(lib.oqb_liouv_rhs_host(L, B, dp(row(0.41 + 0.01 * k)), 1, C.c_void_p(uh.data_ptr()), C.c_void_p(duh.data_ptr())))))
print("stats", op.liouv_stats())
