#!/usr/bin/env python
"""Micro-benchmark of oqb_zgemm_batched (C ABI) on the shapes of the dense RHS paths.

  python benchmarks/zgemm_bench.py            # C3 / C2 / C5 shapes, all op combinations used by the RHS
  OQB_ZGEMM_3M=1 python benchmarks/zgemm_bench.py

Prints one JSON line per shape: TFLOP/s by ALGORITHMIC flops (8·M·N·K·batch) and the relative
Frobenius error of member 0 against numpy (complex128 matmul).
"""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import oqb_b200 as Q  # noqa: E402


def main():
    ctx = Q.get_context()
    rng = np.random.default_rng(0)
    shapes = [("C3", 1024, 1024, 1024, 64), ("C2", 256, 256, 256, 4096), ("C5", 128, 128, 128, 8192),
              ("C2-wide", 256, 256, 2048, 512)]
    ops = [(2, 0), (0, 0), (0, 2)]
    if len(sys.argv) > 1:  # e.g. `zgemm_bench.py C3 20` → one shape, one op pair (for ncu captures)
        shapes = [s for s in shapes if s[0] == sys.argv[1]]
    if len(sys.argv) > 2:
        ops = [(int(sys.argv[2][0]), int(sys.argv[2][1]))]
    dp = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
    al = np.array([1.0, 0.0]); be = np.array([0.0, 0.0])
    for name, M, N, K, batch in shapes:
        for opA, opB in ops:
            a_shape = (K, M) if opA else (M, K)
            b_shape = (N, K) if opB else (K, N)
            # Julia/column-major n×n×B batches == numpy (B, cols, rows); fill on the device
            A = torch.randn(batch, a_shape[1], a_shape[0], dtype=torch.complex128, device="cuda")
            B = torch.randn(1, b_shape[1], b_shape[0], dtype=torch.complex128, device="cuda")
            Cm = torch.zeros(batch, N, M, dtype=torch.complex128, device="cuda")
            call = lambda: ctx.check(ctx.lib.oqb_zgemm_batched(
                ctx.h, opA, opB, M, N, K, dp(al), A.data_ptr(), a_shape[0], a_shape[0] * a_shape[1], B.data_ptr(),
                b_shape[0], 0, dp(be), Cm.data_ptr(), M, M * N, batch))
            for _ in range(2):
                call()
            ctx.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()  # the context runs on torch's current stream (host.Context.__init__)
            for _ in range(reps):
                call()
            e1.record()
            ctx.sync()
            ms = e0.elapsed_time(e1) / reps
            # accuracy of member 0 (column-major storage: numpy sees the transposes)
            a0 = A[0].cpu().numpy().T
            b0 = B[0].cpu().numpy().T
            op = {0: lambda x: x, 1: lambda x: x.T, 2: lambda x: x.conj().T}
            ref = op[opA](a0) @ op[opB](b0)
            got = Cm[0].cpu().numpy().T
            err = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
            # library bar: cublasZgemmStridedBatched on the same operands through torch.bmm
            lib_ms = None
            if opA == 0 and opB == 0:
                # column-major A (M×K) stored as torch (batch, K, M); the product in torch's row-major view: C^T = B^T A^T
                Bt = B.expand(batch, -1, -1)
                for _ in range(2):
                    torch.bmm(Bt, A)
                torch.cuda.synchronize()
                l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                l0.record()
                for _ in range(reps):
                    torch.bmm(Bt, A)
                l1.record()
                torch.cuda.synchronize()
                lib_ms = l0.elapsed_time(l1) / reps
            print(json.dumps({"shape": name, "M": M, "N": N, "K": K, "batch": batch, "opA": opA, "opB": opB,
                              "ms": ms, "tflops_algorithmic": 8.0 * M * N * K * batch / ms / 1e9, "relerr": err,
                              "cublas_zgemm_strided_batched_ms": lib_ms,
                              "cublas_tflops": None if lib_ms is None else 8.0 * M * N * K * batch / lib_ms / 1e9,
                              "mode3m": os.environ.get("OQB_ZGEMM_3M", "1")}), flush=True)
            del A, B, Cm


if __name__ == "__main__":
    main()
