"""Host driver of the device Λ(t) builder for RedfieldLiouvillian.

Mirrors what `quadgk!` does inside (R::RedfieldLiouvillian)(du,u,p,t), reference src/opensys/redfield.jl:46-64:
adaptive Gauss–Kronrod (7,15), one max-heap of segments per integral, bisect the segment with the largest error
until E ≤ atol or E ≤ rtol‖I‖ (QuadGK.jl 2.x `adapt`).  The heap logic and the scalar closures
C_ij(t,x) / f_j(p(x)) run here on the host; every refinement round hands ALL new segments of ALL integrals to
one `oqb_lambda_eval` call (csrc/redfield_lambda.cu), so a sweep over thousands of schedules refines in lock step.

The closed-system unitary of each schedule is given as samples on a uniform grid (what `solve_unitary`'s
interpolated solution provides in practice); `SampledUnitary` is the matching host closure, so the oracle's
`RedfieldLiouvillian(kernels, unitary=SampledUnitary(...), ...)` integrates exactly the same function.
"""
import ctypes as C
import heapq
from typing import Callable, Optional, Sequence

import numpy as np

from . import _lib
from .host import C128, Context, DeviceBatch, get_context

# Kronrod abscissae (non-negative half, descending), Kronrod weights, Gauss weights — QuadGK.kronrod(7)
_X = np.array([0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
               0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
               0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
               0.207784955007898467600689403773245, 0.0])
_WK = np.array([0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
                0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
                0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                0.204432940075298892414161999234649, 0.209482141084727828012999174891714])
_WG = np.array([0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
                0.381830050505118944950369775488975, 0.417959183673469387755102040816327])
# the 15 nodes on [-1, 1] with their Kronrod / Gauss weights: ascending order first (Gauss nodes = odd positions) …
_xi = np.concatenate([-_X[:7], [0.0], _X[:7][::-1]])
_wk = np.concatenate([_WK[:7], [_WK[7]], _WK[:7][::-1]])
_wg_half = np.zeros(7)
_wg_half[1::2] = _WG[:3]
_wg = np.concatenate([_wg_half, [_WG[3]], _wg_half[::-1]])
# … then reordered GAUSS-FIRST (7 Gauss nodes, 8 Kronrod-only nodes): the device evaluates the Gauss rule as the first
# 7 block rows of the node stack (oqb_lambda_eval)
_order = np.concatenate([np.arange(1, 15, 2), np.arange(0, 15, 2)])
NODE_XI, NODE_WK, NODE_WG = _xi[_order], _wk[_order], _wg[_order]
N_GAUSS = 7
_KERNEL_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double),
                         C.POINTER(C.c_double))  # oqb_kernel_fn


class SampledUnitary:
    """U(x) = (1−θ)·U_g + θ·U_{g+1}, g = ⌊x/dt⌋, θ = x/dt − g, from G samples U_g = U(g·dt)."""

    def __init__(self, samples: np.ndarray, dt: float):
        self.samples = np.asarray(samples, dtype=C128)
        self.dt = float(dt)

    @staticmethod
    def locate(x, dt, G):
        pos = np.asarray(x, dtype=np.float64) / dt
        g = np.floor(pos).astype(np.int64)
        th = pos - g
        last = g >= G - 1
        g = np.where(last, G - 1, g)
        th = np.where(last, 0.0, th)
        return g.astype(np.int32), th

    def __call__(self, x: float) -> np.ndarray:
        g, th = self.locate(x, self.dt, self.samples.shape[0])
        g, th = int(g), float(th)
        if th == 0.0:
            return self.samples[g].copy()
        return (1.0 - th) * self.samples[g] + th * self.samples[g + 1]


class LambdaBuilder:
    """Batched Λ(t) for nb schedules × K couplings on one GPU.

    couplings : K constant n×n matrices S_j (already unit-scaled, as ConstantCouplings stores them)
    Ugrid     : (nb, G, n, n) samples of each schedule's unitary (numpy or a DeviceBatch of nb·G matrices)
    dt        : grid spacing per schedule (scalar or (nb,))
    """

    def __init__(self, couplings: Sequence[np.ndarray], Ugrid, dt, G: Optional[int] = None, ctx: Optional[Context] = None):
        self.ctx = ctx or get_context()
        S = np.stack([np.asarray(m, dtype=C128) for m in couplings])
        self.K, self.n = S.shape[0], S.shape[1]
        if isinstance(Ugrid, DeviceBatch):
            if G is None:
                raise ValueError("G is required when Ugrid is a DeviceBatch")
            self.Ud, self.G = Ugrid, int(G)
            self.nb = Ugrid.B // self.G
        else:
            Ugrid = np.asarray(Ugrid, dtype=C128)
            self.nb, self.G = Ugrid.shape[0], Ugrid.shape[1]
            self.Ud = DeviceBatch.from_host(Ugrid.reshape(self.nb * self.G, self.n, self.n), self.ctx)
        self.dt = np.broadcast_to(np.asarray(dt, dtype=np.float64), (self.nb,)).copy()
        h = C.c_void_p()
        Scm = np.ascontiguousarray(S.transpose(0, 2, 1))  # column-major blocks
        self.ctx.check(self.ctx.lib.oqb_lambda_create(self.ctx.h, self.n, self.K, Scm.ctypes.data_as(C.c_void_p), C.byref(h)))
        self.h = h
        self.ctx.check(self.ctx.lib.oqb_lambda_set_unitary(self.h, self.nb, self.G, self.Ud.ptr))
        self.stats = {}

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.ctx.lib.oqb_lambda_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # -- one device round ------------------------------------------------------------------------
    def _eval(self, member, coupling, a, b, t, cvals_fn, slot):
        """GK15 on segments [a_s, b_s]; returns E_s.  cvals_fn(idx, x) → complex weights C(t,x)·f_j(p(x))."""
        nseg = len(member)
        h = 0.5 * (b - a)
        x = a[:, None] + (1.0 + NODE_XI[None, :]) * h[:, None]  # (nseg, 15)
        cv = cvals_fn(x)  # (nseg, 15) complex
        ck = (NODE_WK[None, :] * h[:, None]) * cv
        cg = (NODE_WG[None, :N_GAUSS] * h[:, None]) * cv[:, :N_GAUSS]
        pts = np.concatenate([x, t[:, None]], axis=1)
        dtm = self.dt[member][:, None]
        gi, th = SampledUnitary.locate(pts, dtm, self.G)
        gi = np.ascontiguousarray(gi, dtype=np.int32)
        th = np.ascontiguousarray(th, dtype=np.float64)
        ckf = np.ascontiguousarray(ck.astype(C128)).view(np.float64)
        cgf = np.ascontiguousarray(cg.astype(C128)).view(np.float64)
        member = np.ascontiguousarray(member, dtype=np.int32)
        coupling = np.ascontiguousarray(coupling, dtype=np.int32)
        slot = np.ascontiguousarray(slot, dtype=np.int64)
        E = np.empty(nseg, dtype=np.float64)
        i32 = lambda v: v.ctypes.data_as(_lib.c_i32p)
        self.ctx.check(self.ctx.lib.oqb_lambda_eval(self.h, nseg, i32(member), i32(coupling), i32(gi), _lib.dptr(th),
                                                    _lib.dptr(ckf), _lib.dptr(cgf), slot.ctypes.data_as(_lib.c_i64p),
                                                    _lib.dptr(E)))
        return E

    def _totals(self, qs, heaps, out: DeviceBatch, want_norm=True):
        ptr = np.zeros(len(qs) + 1, dtype=np.int64)
        idx = []
        for k, q in enumerate(qs):
            idx.extend(seg[4] for seg in heaps[q])  # heap array order == QuadGK's re-summation order
            ptr[k + 1] = len(idx)
        idx = np.asarray(idx, dtype=np.int64)
        nrm = np.empty(len(qs), dtype=np.float64) if want_norm else None
        self.ctx.check(self.ctx.lib.oqb_lambda_total(self.h, len(qs), ptr.ctypes.data_as(_lib.c_i64p),
                                                     idx.ctypes.data_as(_lib.c_i64p), out.ptr,
                                                     _lib.dptr(nrm) if want_norm else None))
        return nrm

    # -- the adaptive loop -----------------------------------------------------------------------
    def build(self, member, coupling, t, cfun, Ta: float, atol: float, rtol: float, maxevals: int = 10 ** 7,
              out: Optional[DeviceBatch] = None, t_hi=None, host_heap: bool = False) -> DeviceBatch:
        """Λ for integrals q = 0..nint-1 (see `build_host_heap` for the arguments).  By default the adaptive control flow —
        QuadGK's per-integral segment heap — runs INSIDE the library (oqb_lambda_integrate, csrc/lambda_integrate.cu); `cfun`
        is handed over as a C callback, or, when it is an OhmicBath, evaluated by the library itself (C(t − x)).
        host_heap=True runs the same loop in Python (the round-1 path, kept for A/B checks: identical segment sequence)."""
        from .host import OhmicBath
        if host_heap and not isinstance(cfun, OhmicBath):
            return self.build_host_heap(member, coupling, t, cfun, Ta, atol, rtol, maxevals, out, t_hi)
        member = np.ascontiguousarray(member, dtype=np.int32)
        coupling = np.ascontiguousarray(coupling, dtype=np.int32)
        nint = len(member)
        t = np.ascontiguousarray(np.broadcast_to(np.asarray(t, dtype=np.float64), (nint,)))
        lo = np.ascontiguousarray(np.maximum(0.0, t - Ta))
        hi = t if t_hi is None else np.ascontiguousarray(np.broadcast_to(np.asarray(t_hi, dtype=np.float64), (nint,)))
        Lam = out if out is not None else DeviceBatch(nint, self.n, self.n, self.ctx)
        rounds, segs = C.c_int64(), C.c_int64()
        evals = np.zeros(nint, dtype=np.int64)
        ev_p = evals.ctypes.data_as(_lib.c_i64p)
        i32 = lambda v: v.ctypes.data_as(_lib.c_i32p)
        lib, ctx = self.ctx.lib, self.ctx
        dt = np.ascontiguousarray(self.dt)
        if isinstance(cfun, OhmicBath):
            ctx.check(lib.oqb_lambda_integrate_ohmic(self.h, ctx.h, nint, i32(member), i32(coupling), _lib.dptr(t), _lib.dptr(lo),
                                                     _lib.dptr(hi), _lib.dptr(dt), self.G, cfun.eta, cfun.wc, cfun.beta, atol, rtol,
                                                     maxevals, Lam.ptr, C.byref(rounds), C.byref(segs), ev_p))
        else:
            err = []

            def cb(_user, nseg, q_ptr, t_ptr, x_ptr, out_ptr):
                try:
                    q = np.ctypeslib.as_array(q_ptr, shape=(nseg,)).astype(np.int64)
                    tt = np.ctypeslib.as_array(t_ptr, shape=(nseg,))
                    x = np.ctypeslib.as_array(x_ptr, shape=(nseg, 15))
                    cv = np.asarray(cfun(q[:, None], tt[:, None], x), dtype=C128)
                    np.ctypeslib.as_array(out_ptr, shape=(nseg, 30))[:] = np.ascontiguousarray(np.broadcast_to(cv, (nseg, 15))).view(np.float64)
                except BaseException as e:  # an exception must not cross the C frame
                    err.append(e)
                    np.ctypeslib.as_array(out_ptr, shape=(nseg, 30))[:] = np.nan
            fn = _KERNEL_FN(cb)
            ctx.check(lib.oqb_lambda_integrate(self.h, ctx.h, nint, i32(member), i32(coupling), _lib.dptr(t), _lib.dptr(lo), _lib.dptr(hi),
                                               _lib.dptr(dt), self.G, C.cast(fn, C.c_void_p), None, atol, rtol, maxevals, Lam.ptr,
                                               C.byref(rounds), C.byref(segs), ev_p))
            if err:
                raise err[0]
        self.stats = {"rounds": rounds.value, "segments_evaluated": segs.value, "integrals": nint, "max_evals": int(evals.max())}
        return Lam

    def build_host_heap(self, member, coupling, t, cfun: Callable, Ta: float, atol: float, rtol: float, maxevals: int = 10 ** 7,
                        out: Optional[DeviceBatch] = None, t_hi=None) -> DeviceBatch:
        """Λ for integrals q = 0..nint-1: schedule member[q], coupling index coupling[q] (0-based j of S_j),
        end time t[q].  cfun(q_idx, t, x) → C(t,x)·f_j(p(x)) evaluated on arrays: q_idx (m,1) integral numbers,
        t (m,1), x (m,15).  Returns a DeviceBatch of nint n×n matrices (written into `out` when given).
        The integration window is [max(0, t−Ta), t] (Redfield, redfield.jl:60-61) or, with t_hi given,
        [max(0, t−Ta), t_hi] with U(t) still taken at t (ULE: t_hi = min(t+Ta, tf), lindblad.jl:60-61)."""
        member = np.asarray(member, dtype=np.int32)
        coupling = np.asarray(coupling, dtype=np.int32)
        nint = len(member)
        t = np.broadcast_to(np.asarray(t, dtype=np.float64), (nint,)).copy()
        a0 = np.maximum(0.0, t - Ta)
        hi = t if t_hi is None else np.broadcast_to(np.asarray(t_hi, dtype=np.float64), (nint,)).copy()
        Lam = out if out is not None else DeviceBatch(nint, self.n, self.n, self.ctx)
        import torch
        scratch = DeviceBatch(nint, self.n, self.n, self.ctx)  # totals of the active integrals during refinement
        live = np.nonzero(a0 < hi)[0]  # a == b ⇒ Λ = 0 (quadgk returns zero(f(a)))
        if len(live) < nint:
            Lam.t.zero_()
        if len(live) == 0:
            return Lam
        self.ctx.check(self.ctx.lib.oqb_lambda_reserve_slots(self.h, 2 * nint))
        next_slot = nint
        cv = lambda qs: (lambda x: np.asarray(cfun(qs[:, None], t[qs][:, None], x), dtype=C128))
        E0 = self._eval(member[live], coupling[live], a0[live], hi[live], t[live], cv(live), live.astype(np.int64))
        heaps = {int(q): [(-float(e), 0, float(a0[q]), float(hi[q]), int(q), float(e))] for q, e in zip(live, E0)}
        Etot = {int(q): float(e) for q, e in zip(live, E0)}
        evals = {int(q): 15 for q in live}
        tick = {int(q): 1 for q in live}
        order = [int(q) for q in live]
        nrm = dict(zip(order, self._totals(order, heaps, scratch)))
        active = [q for q in order if Etot[q] > atol and Etot[q] > rtol * nrm[q] and evals[q] < maxevals]
        rounds, nseg_total = 1, len(live)
        while active:
            qs, sa, sb, slots, popped = [], [], [], [], {}
            need = next_slot + len(active)
            self.ctx.check(self.ctx.lib.oqb_lambda_reserve_slots(self.h, need))
            for q in active:
                _, _, a, b, slot, sE = heapq.heappop(heaps[q])
                mid = (a + b) * 0.5
                popped[q] = (a, mid, b, slot, next_slot, sE)
                qs += [q, q]; sa += [a, mid]; sb += [mid, b]; slots += [slot, next_slot]
                next_slot += 1
            qa = np.asarray(qs, dtype=np.int64)
            o = np.argsort(coupling[qa], kind="stable")  # runs of equal coupling → fewer launches for dense S_j
            E = np.empty(len(qa))
            E[o] = self._eval(member[qa][o], coupling[qa][o], np.asarray(sa)[o], np.asarray(sb)[o], t[qa][o], cv(qa[o]),
                              np.asarray(slots, dtype=np.int64)[o])
            for k, q in enumerate(active):
                a, mid, b, s1, s2, sE = popped[q]
                E1, E2 = float(E[2 * k]), float(E[2 * k + 1])
                evals[q] += 30
                Etot[q] = (Etot[q] - sE) + E1 + E2
                heapq.heappush(heaps[q], (-E1, tick[q], a, mid, s1, E1)); tick[q] += 1
                heapq.heappush(heaps[q], (-E2, tick[q], mid, b, s2, E2)); tick[q] += 1
            nrm.update(zip(active, self._totals(active, heaps, scratch)))
            rounds += 1
            nseg_total += len(qa)
            active = [q for q in active if Etot[q] > atol and Etot[q] > rtol * nrm[q] and evals[q] < maxevals]
        # final re-summation of every integral in heap order, written in integral order
        if len(live) == nint:
            self._totals(order, heaps, Lam, want_norm=False)
        else:
            self._totals(order, heaps, scratch, want_norm=False)
            Lam.t[torch.as_tensor(live, device=Lam.t.device)] = scratch.t[: len(live)]
        self.ctx.sync()
        self.stats = {"rounds": rounds, "segments_evaluated": nseg_total, "integrals": nint,
                      "max_evals": max(evals.values()), "E_max": max(Etot.values())}
        return Lam


class ULELiouvillianDevice:
    """(L::ULELiouvillian)(du, u, p, t), reference src/opensys/lindblad.jl:41-69, for nb schedules at once:
    LO_b = Σ_(i,j) Λ_ij (built on the device: window [max(0,t−Ta), min(t+Ta, tf)], U(t) at t), then
    du_b += LO_b u_b LO_b† − ½(LO_b†LO_b u_b + u_b LO_b†LO_b) as five batched ZGEMMs.
    pairs: 1-based (i,j) index pairs; cfun(q, t, x) as in LambdaBuilder.build with q numbering (member, pair)."""

    def __init__(self, builder: LambdaBuilder, pairs, Ta: float, atol: float, rtol: float):
        self.lb, self.pairs, self.Ta, self.atol, self.rtol = builder, [tuple(x) for x in pairs], Ta, atol, rtol

    def LO(self, t, tf, cfun) -> DeviceBatch:
        lb, P = self.lb, len(self.pairs)
        member = np.repeat(np.arange(lb.nb), P)
        coupling = np.tile(np.array([j - 1 for (_, j) in self.pairs]), lb.nb)
        tt = np.broadcast_to(np.asarray(t, dtype=np.float64), (lb.nb,))[member]
        tfv = np.broadcast_to(np.asarray(tf, dtype=np.float64), (lb.nb,))[member]
        Lam = lb.build(member, coupling, tt, cfun, self.Ta, self.atol, self.rtol, t_hi=np.minimum(tt + self.Ta, tfv))
        LO = DeviceBatch(lb.nb, lb.n, lb.n, lb.ctx)
        nn, one = lb.n * lb.n, np.array([1.0, 0.0])
        for k in range(P):  # axpy!(1.0, Λ, LO), lindblad.jl:65
            lb.ctx.check(lb.ctx.lib.oqb_axpy_batched(lb.ctx.h, nn, lb.nb, _lib.dptr(one), C.c_void_p(Lam.t.data_ptr() + 16 * nn * k),
                                                     P * nn, LO.ptr))
        return LO

    def __call__(self, du: DeviceBatch, u: DeviceBatch, t, tf, cfun):
        """ACCUMULATES into du (device batches of nb n×n matrices)."""
        lb = self.lb
        ctx, n, nb = lb.ctx, lb.n, lb.nb
        LO = self.LO(t, tf, cfun)
        T1, M = u.zeros_like(), u.zeros_like()
        nn = n * n
        one, zero, mh = np.array([1.0, 0.0]), np.array([0.0, 0.0]), np.array([-0.5, 0.0])
        g = lambda opA, opB, al, A, B_, be, C_: ctx.check(ctx.lib.oqb_zgemm_batched(
            ctx.h, opA, opB, n, n, n, _lib.dptr(al), A.ptr, n, nn, B_.ptr, n, nn, _lib.dptr(be), C_.ptr, n, nn, nb))
        g(0, 0, one, LO, u, zero, T1)     # T1 = LO u
        g(0, 2, one, T1, LO, one, du)     # du += T1 LO†
        g(2, 0, one, LO, LO, zero, M)     # M = LO† LO
        g(0, 0, mh, M, u, one, du)        # du −= ½ M u
        g(0, 0, mh, u, M, one, du)        # du −= ½ u M
        return du
