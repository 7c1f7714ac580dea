#!/usr/bin/env python
"""C5 variant (ii), SURVEY §8(d): Redfield Λ(t) built on the device for a sweep of schedules.

  python benchmarks/lambda_bench.py [--schedules 256] [--grid 33]

n = 128 (7 qubits), K = 7 diagonal couplings Z_α, one unitary grid per schedule (U_b(t) = exp(−i H_b t) sampled on
`grid` points, generated on the device), Gaussian-decaying complex bath correlation, atol 1e-8 / rtol 1e-6 (the
reference test's tolerances, test/opensys/redfield.jl:12).  Prints one JSON line: wall time of the adaptive build,
time inside the device calls, the ZGEMM share (CUDA-event bracketed) and algorithmic TFLOP/s
(segments × 8n³ × 26: 15 Kronrod + 7 Gauss node products + 4 for the U(t) rotation).
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import oqb_b200 as Q  # noqa: E402
from oqb_b200.redfield_device import LambdaBuilder  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--schedules", type=int, default=256)
    ap.add_argument("--grid", type=int, default=33)
    ap.add_argument("--qubits", type=int, default=7)
    ap.add_argument("--rtol", type=float, default=1e-6)
    ap.add_argument("--cpu-integrals", type=int, default=2, help="integrals timed through the oracle (0 = skip)")
    args = ap.parse_args()
    nq, nb, G = args.qubits, args.schedules, args.grid
    n, K, tf = 2 ** nq, nq, 10.0
    ctx = Q.get_context()
    dev = f"cuda:{ctx.device}"
    gen = torch.Generator(device=dev).manual_seed(5)
    A = torch.randn((nb, n, n), dtype=torch.complex128, device=dev, generator=gen)
    H = (A + A.conj().transpose(1, 2)) * (0.25 / np.sqrt(n))
    w, v = torch.linalg.eigh(H)
    dt = tf / (G - 1)
    U = torch.empty((nb, G, n, n), dtype=torch.complex128, device=dev)
    for g in range(G):
        ug = (v * torch.exp(-1j * w * (g * dt))[:, None, :]) @ v.conj().transpose(1, 2)
        U[:, g] = ug.transpose(1, 2)  # column-major blocks
    Ud = Q.DeviceBatch(nb * G, n, n, ctx, U.reshape(nb * G, n, n))
    z = []
    for a in range(K):
        d = np.array([1.0 - 2.0 * ((r >> (nq - 1 - a)) & 1) for r in range(n)])
        z.append(np.diag(d).astype(complex))
    lb = LambdaBuilder(z, Ud, dt, G=G, ctx=ctx)
    member = np.repeat(np.arange(nb), K)
    coupling = np.tile(np.arange(K), nb)
    cf = lambda q, t, x: np.exp(-0.5 * (t - x) ** 2) * np.exp(-0.8j * (t - x))

    # time spent inside the device entry points
    dev_s = [0.0]
    orig_eval, orig_tot = lb._eval, lb._totals

    def timed(fn):
        def w_(*a, **k):
            t0 = time.perf_counter()
            r = fn(*a, **k)
            ctx.sync()
            dev_s[0] += time.perf_counter() - t0
            return r
        return w_
    lb._eval, lb._totals = timed(orig_eval), timed(orig_tot)

    out = Q.DeviceBatch(nb * K, n, n, ctx)
    lb.build(member[:K], coupling[:K], tf, cf, tf, 1e-8, args.rtol)  # warm-up (workspace, tensor maps)
    dev_s[0] = 0.0
    ctx.check(ctx.lib.oqb_profile_begin(ctx.h))
    t0 = time.perf_counter()
    lb.build(member, coupling, tf, cf, tf, 1e-8, args.rtol, out=out)
    wall = time.perf_counter() - t0
    zg_ms, zg_n, zg_fl = C.c_double(), C.c_uint64(), C.c_double()
    ctx.check(ctx.lib.oqb_profile_end(ctx.h, C.byref(zg_ms), C.byref(zg_n), C.byref(zg_fl)))
    st = lb.stats
    cpu = None
    if args.cpu_integrals > 0:  # the reference-structured CPU path (oracle: quadgk + 3 GEMMs per node) on the same integrals
        from bench import cpu_lambda_baseline  # bench.py's CPU-baseline leg is the only place that executes oracle/
        Uh = U[0].cpu().numpy().transpose(0, 2, 1)
        cfs = lambda t, x: complex(cf(None, np.float64(t), np.float64(x)))
        cs, ref = cpu_lambda_baseline(Uh, dt, z, cfs, tf, 1e-8, args.rtol, args.cpu_integrals)
        got = out.to_host()[(args.cpu_integrals - 1) % K]
        cpu = {"value": args.cpu_integrals / cs, "unit": "lambda_matrices_per_sec", "kind": "port", "cores": os.cpu_count(),
               "sample": f"{args.cpu_integrals} integrals of schedule 0 through oracle RedfieldLiouvillian.Lambda (numpy/OpenBLAS)",
               "relerr_device_vs_oracle": float(np.linalg.norm(got - ref) / np.linalg.norm(ref))}
    segs = st["segments_evaluated"]
    flops = segs * 8.0 * n ** 3 * 26
    dmma, _ = ctx.probe_fp64()
    print(json.dumps({
        "config": "C5-lambda-build", "workload": f"{nb} schedules x {K} couplings, n={n}, unitary grid {G} points, atol 1e-8 rtol {args.rtol}",
        "metric": "lambda_matrices_per_sec", "value": nb * K / wall, "wall_s": wall, "device_call_s": dev_s[0],
        "host_heap_and_closure_s": wall - dev_s[0], "rounds": st["rounds"], "segments_evaluated": segs,
        "max_integrand_evals_per_integral": st["max_evals"], "algorithmic_tflop": flops / 1e12,
        "tflops_over_device_time": flops / dev_s[0] / 1e12,
        "roofline": {"bound": "tensor", "kernel": "zgemm_tma_kernel", "achieved": zg_fl.value / (zg_ms.value * 1e-3) / 1e12,
                     "peak": dmma, "unit": "TFLOP/s", "frac": zg_fl.value / (zg_ms.value * 1e-3) / 1e12 / dmma,
                     "kernel_share_of_device_time": zg_ms.value * 1e-3 / dev_s[0], "zgemm_launches": zg_n.value,
                     "peak_source": "DMMA issue probe, this run"}, "cpu_baseline": cpu}))


if __name__ == "__main__":
    main()
