"""TEST INFRASTRUCTURE — an independent CPU statement of the Davies generator that scales to lvl = 1024.

The literal restatement of the reference's loop (`oqb_oracle.DaviesGenerator.rhs_acc`, src/opensys/davies.jl:21-47 with
`build_gap_indices`, src/opensys/opensys_util.jl:25-61) is O(#groups · l²) and its GapIndices builder inserts into Python
lists — it cannot finish at the headline size (10 qubits, 523 776 level pairs).  This module evaluates the SAME sum

    du += Σ_x γ(x)·(L_x ρ L_x' − ½{L_x'L_x, ρ}) − i[Σ_x S(x) L_x'L_x, ρ],      L_x = A∘[ω̃ == x]      

by a different route than the CUDA path (no pair lists, no Hadamard tables):
  * gap keys by the scalar rule of the reference for every pair i<j (|gap| ≤ 10^-digits → zero group, else
    `round(gap, sigdigits)` — the oracle's own scalar `round_sigdigits`), grouped in a dict by exact float equality;
  * single-pair groups vectorised as rank-one terms;
  * every multi-pair group (and the zero group) by SUB-MATRIX algebra: L_x restricted to the rows/columns it touches,
    `du[R,R] += γ·Lsub ρ[C,C] Lsub'`, `G[C,C] += γ·Lsub'Lsub`.
It is validated against the literal loop at l ≤ 64 (generic, equally spaced, degenerate and forced 8-digit-coincidence
spectra) in tests/test_oracle_golden.py, and is then the checker for the full-size C3 parity tests.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this package.
"""
import numpy as np

from .oqb_oracle import C128, round_sigdigits


def gap_groups(w, digits=8, sigdigits=8):
    """{key: [(i, j), …]} for the non-zero groups (pairs in the reference's loop order) and the zero-group pairs."""
    w = np.asarray(w, dtype=np.float64)
    l = len(w)
    thr = 10.0 ** (-digits)
    groups, zero = {}, []
    for i in range(l - 1):
        wi = w[i]
        for j in range(i + 1, l):
            gap = float(w[j] - wi)
            if abs(gap) <= thr:
                zero.append((i, j))
            else:
                groups.setdefault(round_sigdigits(gap, sigdigits), []).append((i, j))
    return groups, zero


def _submatrix_terms(As, rows, cols, g, s, rho, du, G, Hls):
    rows, cols = np.asarray(rows), np.asarray(cols)
    ur, pr = np.unique(rows, return_inverse=True)
    uc, pc = np.unique(cols, return_inverse=True)
    rsub = rho[np.ix_(uc, uc)]
    for A in As:
        Ls = np.zeros((len(ur), len(uc)), dtype=C128)
        Ls[pr, pc] = A[rows, cols]
        if g != 0.0:
            du[np.ix_(ur, ur)] += g * (Ls @ rsub @ Ls.conj().T)
        LL = Ls.conj().T @ Ls
        G[np.ix_(uc, uc)] += g * LL
        Hls[np.ix_(uc, uc)] += s * LL


def davies_eig_acc(du, rho, As, w, gamma, S, digits=8, sigdigits=8, groups=None):
    """du += Davies terms of ρ given in the eigenbasis; As: rotated couplings v'A_αv (l×l each); gamma, S scalar callables."""
    l = len(w)
    As = [np.asarray(A, dtype=C128) for A in As]
    grp, zero = groups if groups is not None else gap_groups(w, digits, sigdigits)
    m2 = sum(np.abs(A) ** 2 for A in As)  # Σ_α |A_α[a,b]|²
    G = np.zeros((l, l), dtype=C128)
    Hls = np.zeros((l, l), dtype=C128)
    pop = np.diag(rho).copy()
    # ---- single-pair groups: L = A_ij|i><j| (rate γ(+x)) and A_ji|j><i| (rate γ(−x)) -------------------------------
    si, sj, sk = [], [], []
    for key, pairs in grp.items():
        if len(pairs) == 1:
            si.append(pairs[0][0]); sj.append(pairs[0][1]); sk.append(key)
    if si:
        si, sj, sk = np.asarray(si), np.asarray(sj), np.asarray(sk)
        gp = np.array([gamma(float(x)) for x in sk]); gm = np.array([gamma(-float(x)) for x in sk])
        sp = np.array([S(float(x)) for x in sk]); sm = np.array([S(-float(x)) for x in sk])
        dd = np.zeros(l, dtype=C128)
        np.add.at(dd, si, gp * m2[si, sj] * pop[sj])
        np.add.at(dd, sj, gm * m2[sj, si] * pop[si])
        du[np.arange(l), np.arange(l)] += dd
        gd = np.zeros(l); hd = np.zeros(l)
        np.add.at(gd, sj, gp * m2[si, sj]); np.add.at(gd, si, gm * m2[sj, si])
        np.add.at(hd, sj, sp * m2[si, sj]); np.add.at(hd, si, sm * m2[sj, si])
        G[np.arange(l), np.arange(l)] += gd
        Hls[np.arange(l), np.arange(l)] += hd
    # ---- multi-pair groups ------------------------------------------------------------------------------------------
    for key, pairs in grp.items():
        if len(pairs) == 1:
            continue
        ii = [p[0] for p in pairs]; jj = [p[1] for p in pairs]
        _submatrix_terms(As, ii, jj, gamma(float(key)), S(float(key)), rho, du, G, Hls)
        _submatrix_terms(As, jj, ii, gamma(-float(key)), S(-float(key)), rho, du, G, Hls)
    # ---- zero group: degenerate pairs in both orders + the diagonal -------------------------------------------------
    zi = [p[0] for p in zero]; zj = [p[1] for p in zero]
    rows = zi + zj + list(range(l)); cols = zj + zi + list(range(l))
    _submatrix_terms(As, rows, cols, gamma(0.0), S(0.0), rho, du, G, Hls)
    du += -0.5 * (G @ rho + rho @ G) - 1.0j * (Hls @ rho - rho @ Hls)
    return du


def ame_rhs(u, w, v, couplings, terms, digits=8, sigdigits=8):
    """(Op::DiffEqLiouvillian{true,false})(du,u,p,t), src/opensys/diffeq_liouvillian.jl:92-105, with the Davies terms
    above: terms = [(γ, S, coupling indices)] — one entry per DaviesGenerator of opensys_eig.  v: n×lvl."""
    w = np.asarray(w, dtype=np.float64)
    v = np.asarray(v, dtype=C128)
    rho = v.conj().T @ u @ v
    X = -1.0j * (w[:, None] * rho - rho * w[None, :])
    groups = gap_groups(w, digits, sigdigits)
    rot = [v.conj().T @ np.asarray(c, dtype=C128) @ v for c in couplings]
    for gamma, S, idx in terms:
        davies_eig_acc(X, rho, [rot[k] for k in idx], w, gamma, S, digits, sigdigits, groups)
    return v @ X @ v.conj().T
