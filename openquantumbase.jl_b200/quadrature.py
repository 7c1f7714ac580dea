"""Adaptive Gauss–Kronrod (7,15) quadrature of matrix-valued host closures.

Used only for the closure-driven, ρ-independent Λ(t) of RedfieldLiouvillian (reference
src/opensys/redfield.jl:46-64 calls QuadGK.quadgk!).  Same scheme as QuadGK.jl 2.x: evaluate the
rule on [a,b]; while E > atol and E > rtol·‖I‖, bisect the segment with the largest error estimate
(max-heap), E = Frobenius norm of (Kronrod − Gauss).
"""
import heapq

import numpy as np

_X = np.array([0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
               0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
               0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
               0.207784955007898467600689403773245])
_WK = np.array([0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
                0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
                0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                0.204432940075298892414161999234649])
_WK0 = 0.209482141084727828012999174891714
_WG = np.array([0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
                0.381830050505118944950369775488975])
_WG0 = 0.417959183673469387755102040816327


def _rule(f, a, b):
    h = 0.5 * (b - a)
    Ig = Ik = 0.0
    for i in range(7):
        fs = f(a + (1 - _X[i]) * h) + f(a + (1 + _X[i]) * h)
        Ik = Ik + fs * _WK[i]
        if i % 2 == 1:
            Ig = Ig + fs * _WG[i // 2]
    f0 = f(a + h)
    Ik = (Ik + f0 * _WK0) * h
    Ig = (Ig + f0 * _WG0) * h
    return Ik, float(np.linalg.norm(Ik - Ig))


def quadgk_matrix(f, a, b, rtol, atol, maxevals=10 ** 7):
    if a == b:
        return 0.0 * f(a)
    I, E = _rule(f, a, b)
    heap = [(-E, 0, a, b, I, E)]
    evals, tick = 15, 1
    while E > atol and E > rtol * float(np.linalg.norm(I)) and evals < maxevals:
        _, _, sa, sb, sI, sE = heapq.heappop(heap)
        mid = 0.5 * (sa + sb)
        I1, E1 = _rule(f, sa, mid)
        I2, E2 = _rule(f, mid, sb)
        evals += 30
        I = (I - sI) + I1 + I2
        E = (E - sE) + E1 + E2
        heapq.heappush(heap, (-E1, tick, sa, mid, I1, E1)); tick += 1
        heapq.heappush(heap, (-E2, tick, mid, sb, I2, E2)); tick += 1
    tot = None
    for seg in heap:
        tot = seg[4] if tot is None else tot + seg[4]
    return tot
