"""Fixture builders for the oracle (TEST INFRASTRUCTURE ONLY — never imported by the product).

Restates the operator-construction helpers the reference's own tests use to build their
inputs, so that the golden vectors of /root/reference/test can be regenerated without Julia.
Citations are relative to /root/reference.

  Pauli constants              src/math_util.jl:3-17
  kron (the ``⊗`` operator)    src/math_util.jl:47
  q_translate                  src/matrix_util.jl:16-33   (string → Σ coefficient · Pauli string)
  single_clause                src/matrix_util.jl:54-106
  collective_operator          src/matrix_util.jl:123-130
  standard_driver              src/matrix_util.jl:137-143
  local_field_term             src/matrix_util.jl:177-183
  two_local_term               src/matrix_util.jl:197-203
  alt_sec_chain / random_ising src/dev_tools.jl:7-32
"""
import re
from functools import reduce

import numpy as np

sx = np.array([[0, 1], [1, 0]], dtype=np.complex128)
sy = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
sz = np.array([[1, 0], [0, -1]], dtype=np.complex128)
si = np.eye(2, dtype=np.complex128)
sm = np.array([[0, 0], [1, 0]], dtype=np.complex128)  # σ₋  math_util.jl:7
sp = np.array([[0, 1], [0, 0]], dtype=np.complex128)  # σ₊  math_util.jl:8
_PAULI = {"X": sx, "Y": sy, "Z": sz, "I": si}

# PauliVec (math_util.jl:10-30): eigenvectors of σx, σy, σz; 0-based here.
PauliVec = [
    [np.array([1, 1], dtype=np.complex128) / np.sqrt(2), np.array([1, -1], dtype=np.complex128) / np.sqrt(2)],
    [np.array([1, 1j], dtype=np.complex128) / np.sqrt(2), np.array([1, -1j], dtype=np.complex128) / np.sqrt(2)],
    [np.array([1, 0], dtype=np.complex128), np.array([0, 1], dtype=np.complex128)],
]


def kron(*ops):
    return reduce(np.kron, ops)


def q_translate(h: str) -> np.ndarray:
    """"-0.9ZZII+IZZI" → dense matrix.  The reference evals the string as Julia code
    (matrix_util.jl:31-32); for the grammar its tests use that is a signed, optionally
    weighted sum of Pauli strings."""
    h = h.replace(" ", "")
    res = None
    for sign, coef, word in re.findall(r"([+-]?)([0-9.]*)\*?([XYZI]+)", h):
        c = float(coef) if coef else 1.0
        if sign == "-":
            c = -c
        term = c * kron(*[_PAULI[ch] for ch in word])
        res = term if res is None else res + term
    if res is None:
        raise ValueError(f"cannot parse Pauli expression {h!r}")
    return res


def single_clause(ops, q_ind, weight, num_qubit):
    """matrix_util.jl:54-106; q_ind is 1-based like the reference."""
    mats = []
    for i in range(1, num_qubit + 1):
        if i in q_ind:
            op = ops[list(q_ind).index(i)]
            mats.append(_PAULI[op.upper()] if isinstance(op, str) else op)
        else:
            mats.append(si)
    return weight * kron(*mats)


def collective_operator(op: str, num_qubit: int):
    return sum(single_clause([op], [i], 1.0, num_qubit) for i in range(1, num_qubit + 1))


def standard_driver(num_qubit: int):
    return collective_operator("x", num_qubit)


def local_field_term(h, idx, num_qubit):
    return sum(single_clause(["z"], [i], w, num_qubit) for w, i in zip(h, idx))


def two_local_term(j, idx, num_qubit):
    return sum(single_clause(["z", "z"], list(ij), w, num_qubit) for w, ij in zip(j, idx))


def alt_sec_chain(w1, w2, n, num_qubits):
    """dev_tools.jl:24-32."""
    J, idx = [], []
    for i in range(1, num_qubits):
        J.append(w1 if (int(np.ceil(i / n)) % 2 == 1) else w2)
        idx.append([i, i + 1])
    return two_local_term(J, idx, num_qubits)


def random_ising(num_qubits, rng):
    """dev_tools.jl:7-17 with an explicit numpy Generator in place of Julia's global RNG."""
    J, idx = [], []
    for i in range(1, num_qubits + 1):
        for j in range(i + 1, num_qubits + 1):
            J.append(2 * rng.random() - 1)
            idx.append([i, j])
    return two_local_term(J, idx, num_qubits)


def ising_diag(n, J, idx):
    """Diagonal of Σ J_ij Z_i Z_j as a real vector (qubit 1 is the most significant bit,
    matching the kron ordering above); used for the large synthetic configs where building
    dense krons would be wasteful."""
    dim = 1 << n
    k = np.arange(dim)
    z = [1.0 - 2.0 * ((k >> (n - q)) & 1) for q in range(1, n + 1)]
    d = np.zeros(dim)
    for w, (a, b) in zip(J, idx):
        d += w * z[a - 1] * z[b - 1]
    return d


def random_density_matrix(n, rng):
    g = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    rho = g @ g.conj().T
    return rho / np.trace(rho).real
