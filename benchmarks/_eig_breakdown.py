"""Phase timing of one new eigen-system at C3 (scratch diagnostic): H(s) assembly, eigensolver, rotation, tables."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import oqb_b200 as Q
from oqb_b200 import host as Hm

ctx = Q.get_context()
Hd, Hp, coup = bench.build_c3()
H = Q.DenseHamiltonian([bench.fA, bench.fB], [Hd, Hp], unit="h")
bath = Q.Ohmic(bench.ETA, bench.FC, bench.TMK)
op = Q.DiffEqLiouvillian(H, [Q.DaviesGenerator(Q.ConstantCouplings(coup, unit="h"), bath, Q.ZERO_LAMBSHIFT)], [], bench.N)
op.plan_cache, op.device_eigh, op.device_tables = 1, True, True
for k in range(3):
    op._prepare_ame(0.3 + 0.01 * k)
ctx.sync()
res = {}
def T(name, fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    res[name] = best * 1e3
    return r
cnt = [0]
def eigs():
    cnt[0] += 1
    return Hm.haml_eigs_device(H, 0.4 + 0.001 * cnt[0], bench.N)
T("haml_eigs_device_ms", eigs)
A = torch.randn(1024, 1024, dtype=torch.complex128, device="cuda"); A = A + A.conj().T
T("torch_eigh_only_ms", lambda: torch.linalg.eigh(A))
T("torch_eigvalsh_only_ms", lambda: torch.linalg.eigvalsh(A))
def full():
    cnt[0] += 1
    op._prepare_ame(0.6 + 0.001 * cnt[0]); ctx.sync()
T("prepare_total_ms", full)
import cProfile, pstats, io
pr = cProfile.Profile(); pr.enable(); full(); pr.disable()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(18)
print(json.dumps(res)); print(st.getvalue()[:3500])

# per-call wall time of every C-ABI function used by one _prepare_ame
import collections
acc = collections.OrderedDict()
lib = ctx.lib
names = ["oqb_dense_hamiltonian", "oqb_ame_set_eigen_device", "oqb_ame_set_lambshift_spline", "oqb_ame_gaps_build", "oqb_ame_add_davies_ohmic", "oqb_ame_create", "oqb_ame_destroy"]
class W:
    def __init__(s, f, n): s.f, s.n = f, n
    def __call__(s, *a):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = s.f(*a); torch.cuda.synchronize()
        acc[s.n] = acc.get(s.n, 0.0) + (time.perf_counter() - t0) * 1e3
        return r
class LibProxy:
    def __init__(s, lib): s._lib = lib
    def __getattr__(s, n):
        f = getattr(s._lib, n)
        return W(f, n) if n in names else f
ctx.lib = LibProxy(lib)
H.ctx.lib = ctx.lib
for _ in range(4):
    acc.clear(); full(); print(json.dumps(acc))
