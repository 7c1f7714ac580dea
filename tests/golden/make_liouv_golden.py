"""Writes tests/golden/liouv_c_demo.bin: the expected output of examples/c_liouv_demo.c, computed by the CPU oracle's LITERAL
Davies gap loop (oracle/oqb_oracle.py: build_gap_indices + DaviesGenerator.rhs_acc, src/opensys/davies.jl:21-47) on the
operators and states the C program constructs for itself.  NS = 3 values of s × B = 2 states × 64×64 complex128, each block
column-major.  Run from the repo root:  python tests/golden/make_liouv_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oqb_oracle as O  # noqa: E402

Q, N = 6, 64


def zval(state, qubit):
    return -1.0 if (state >> (Q - 1 - qubit)) & 1 else 1.0


def build():
    Hd = np.zeros((N, N), dtype=complex)
    dz = np.zeros(N)
    for c in range(N):
        for q in range(Q):
            Hd[c ^ (1 << q), c] = -1.0
        for i in range(Q):
            for j in range(i + 1, Q):
                dz[c] += (((i + 1) * 7 + (j + 1) * 3) % 5 - 2) * 0.5 * zval(c, i) * zval(c, j)
    Hp = np.diag(dz).astype(complex)
    coup = [2 * np.pi * np.diag([zval(c, q) for c in range(N)]).astype(complex) for q in range(Q)]
    i, j = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    u = np.stack([(np.sin(0.37 * i + 0.11 * j + b) + 1j * np.cos(0.23 * i - 0.19 * j - b)) / N for b in range(2)])
    return 2 * np.pi * Hd, 2 * np.pi * Hp, coup, u


def main():
    Hd, Hp, coup, u = build()
    bath = O.Ohmic(1e-4, 4, 16)
    out = []
    for s in (0.0, 0.5, 1.0):
        w, v = np.linalg.eigh((1 - s) * Hd + s * Hp)
        gi = O.build_gap_indices(w, 8, 8)
        D = O.DaviesGenerator(coup, bath.spectrum, lambda x: 0.0)
        for x in u:
            rho = v.conj().T @ x @ v
            X = -1j * (w[:, None] * rho - rho * w[None, :])
            D.rhs_acc(X, rho, gi, v, s)
            out.append((v @ X @ v.conj().T).T.copy())  # column-major block
    arr = np.ascontiguousarray(np.stack(out))
    path = os.path.join(ROOT, "tests", "golden", "liouv_c_demo.bin")
    arr.tofile(path)
    print(path, arr.shape, arr.nbytes, "bytes")


if __name__ == "__main__":
    main()
