"""Host mirrors of the reference's remaining Liouvillian types on the RHS path, all evaluated by the
same device entry points as the main types:

  ConstDaviesGenerator   src/opensys/davies.jl:102-119   → oqb_dense_rhs (H = H_LS, L_k = Linds, γ ≡ 1) + oqb_axpy_batched
  ConstHDaviesGenerator  src/opensys/davies.jl:121-163   → the closed-form Davies tables in a fixed basis (oqb_ame_* with v = I)
  FluctuatorLiouvillian  src/opensys/stochastic.jl:11-44 → oqb_dense_rhs / update_cache / update_vectorized_cache with the
                                                           noise values n_i as per-member coefficients

`build_const_davies` (src/opensys/davies.jl:173-209) is construction-time host work (index bookkeeping on l×l
matrices), like the reference's.
"""
import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .gaps import GapGroups, _eval, jump_slots
from .host import C128, Context, DeviceBatch, _HostIO, _colmajor_stack, _svec, add_davies_terms, get_context

_ONE = np.array([1.0, 0.0])


def _axpy(ctx, x: DeviceBatch, y: DeviceBatch, shared_x=False):
    n_el = y.rows * y.cols
    ctx.check(ctx.lib.oqb_axpy_batched(ctx.h, n_el, y.B, _lib.dptr(_ONE), x.ptr, 0 if shared_x else n_el, y.ptr))


class ConstDaviesGenerator:
    def __init__(self, Linds: Sequence[np.ndarray], Hls: np.ndarray, A: np.ndarray, ctx: Optional[Context] = None):
        self.Linds = [np.asarray(L, dtype=C128) for L in Linds]
        self.Hls, self.A = np.asarray(Hls, dtype=C128), np.asarray(A, dtype=C128)
        self.ctx = ctx or get_context()
        self._op = None

    def _dense(self):
        if self._op is None:
            h = C.c_void_p()
            n = self.Hls.shape[0]
            hb, lb = _colmajor_stack([self.Hls]), _colmajor_stack(self.Linds)
            self.ctx.check(self.ctx.lib.oqb_dense_create(self.ctx.h, n, 1, hb.ctypes.data, len(self.Linds), lb.ctypes.data, C.byref(h)))
            self._op = h
        return self._op

    def __call__(self, du, rho, *_):
        """(D::ConstDaviesGenerator)(du, ρ, _, _) :111-117 — ACCUMULATES."""
        io = _HostIO(rho, du)
        B, K = io.u.B, len(self.Linds)
        tmp = io.u.zeros_like()
        one, g = np.ones((1, 1)), np.ones((1, K))
        self.ctx.check(self.ctx.lib.oqb_dense_rhs(self._dense(), B, _lib.dptr(one), 1, _lib.dptr(g), 1, io.u.ptr, tmp.ptr))
        _axpy(self.ctx, tmp, io.du)
        return io.finish()

    def ame_jump(self, psi: DeviceBatch, uniforms, mask=None, apply=False):
        """ame_jump(D::ConstDaviesGenerator, u, _, _), src/opensys/trajectory_jump.jl:53-56: prob_k = real(u'L_k'L_ku) for the
        precomputed operators (√γ already inside), sample(D.Linds, Weights(prob)) with the supplied uniforms; returns
        (choice[B], prob[B, K]) device tensors.  apply=True performs ψ ← L_kψ/‖L_kψ‖.  The operators are sparse in one fixed
        basis, so they go to the AME jump kernel as probability slots (one per operator: its non-zero entries, rate 1) of a
        handle in the adiabatic frame (v = I) — any number of operators."""
        import torch
        lib, ctx = self.ctx.lib, self.ctx
        l, K = self.Hls.shape[0], len(self.Linds)
        if getattr(self, "_jump", None) is None:
            h = C.c_void_p()
            cbuf = _colmajor_stack(self.Linds)
            ctx.check(lib.oqb_ame_create(ctx.h, l, l, K, cbuf.ctypes.data, C.byref(h)))
            ctx.check(lib.oqb_ame_set_eigen(h, _lib.dptr(np.zeros(l)), None))
            ptr, terms = [0], []
            for k, L in enumerate(self.Linds):
                r, c_ = np.nonzero(L)
                o = np.lexsort((c_, r))
                terms += [(k, int(a), int(b)) for a, b in zip(r[o], c_[o])]
                ptr.append(len(terms))
            ptr = np.asarray(ptr, dtype=np.int64)
            terms = np.ascontiguousarray(np.asarray(terms, dtype=np.int32).reshape(-1, 3))
            rate = np.ones(K)
            ctx.check(lib.oqb_ame_set_jump(h, K, ptr.ctypes.data_as(_lib.c_i64p), terms.ctypes.data_as(_lib.c_i32p), _lib.dptr(rate)))
            self._jump = h
        B, dev = psi.B, psi.t.device
        u_d = torch.as_tensor(np.asarray(uniforms, dtype=np.float64), device=dev)
        m_d = None if mask is None else torch.as_tensor(np.asarray(mask, dtype=np.uint8), device=dev)
        choice = torch.zeros(B, dtype=torch.int32, device=dev)
        prob = torch.zeros((B, K), dtype=torch.float64, device=dev)
        ctx.check(lib.oqb_ame_jump(self._jump, B, psi.ptr, C.c_void_p(u_d.data_ptr()), None if m_d is None else C.c_void_p(m_d.data_ptr()),
                                   C.c_void_p(choice.data_ptr()), C.c_void_p(prob.data_ptr()), int(apply)))
        ctx.sync()
        return choice, prob

    def update_cache(self, cache, *_):
        """update_cache!(cache, D, _, _) = cache .+= D.A  (:119)."""
        io = _HostIO(cache, cache)
        Ad = DeviceBatch.from_host(self.A[None], self.ctx)
        _axpy(self.ctx, Ad, io.du, shared_x=True)
        return io.finish()


def build_const_davies(couplings: Sequence[np.ndarray], w, gamma, S, digits=8, sigdigits=8, ctx=None) -> ConstDaviesGenerator:
    """src/opensys/davies.jl:173-209 for constant couplings given in the eigenbasis (rotate(couplings, v))."""
    cs = [np.asarray(c, dtype=C128) for c in couplings]
    l = len(w)
    _, _, rkey, tags = jump_slots(w, len(cs), digits, sigdigits)
    uq, inv = np.unique(rkey, return_inverse=True)
    gv, sv = _eval(gamma, uq)[inv], _eval(S, uq)[inv]
    res = []
    Hls = np.zeros((l, l), dtype=C128)
    Am = np.zeros((l, l), dtype=C128)
    for (i, rows, cols), g, sv_ in zip(tags, gv, sv):
        L = np.zeros((l, l), dtype=C128)
        L[rows, cols] = cs[i][rows, cols]
        LL = L.conj().T @ L
        res.append(np.sqrt(g) * L)
        Hls += sv_ * LL
        Am -= 0.5 * np.sqrt(g) ** 2 * LL
    Am -= 1.0j * Hls
    return ConstDaviesGenerator(res, Hls, Am, ctx)


class ConstHDaviesGenerator:
    """Constant Hamiltonian with eigen-energies w (GapIndices fixed at construction), couplings given in the
    eigenbasis (possibly s-dependent).  (D)(du, ρ, _, s) ACCUMULATES the Davies terms of ρ in that basis."""

    def __init__(self, w, coupling, gamma, S, digits=8, sigdigits=8, ctx: Optional[Context] = None):
        self.w = np.ascontiguousarray(np.real(w), dtype=np.float64)
        self.coupling = coupling if callable(coupling) else (lambda s, c=[np.asarray(m, dtype=C128) for m in coupling]: c)
        self.gamma, self.S = gamma, S
        self.digits, self.sigdigits = digits, sigdigits
        self.ctx = ctx or get_context()
        self._ame, self._key = None, None

    def _plan(self, s):
        if self._ame is not None and self._key == s:
            return self._ame
        lib, ctx = self.ctx.lib, self.ctx
        if self._ame is not None:
            lib.oqb_ame_destroy(self._ame)
        coup = [np.asarray(m, dtype=C128) for m in self.coupling(s)]
        l = len(self.w)
        h = C.c_void_p()
        cbuf = _colmajor_stack(coup)  # keep the buffer alive across the call
        ctx.check(lib.oqb_ame_create(ctx.h, l, l, len(coup), cbuf.ctypes.data, C.byref(h)))
        ctx.check(lib.oqb_ame_set_eigen(h, _lib.dptr(self.w), None))
        add_davies_terms(ctx, h, [(self.gamma, self.S, 0, len(coup))], self.digits, self.sigdigits)
        self._ame, self._key = h, s
        return h

    def __call__(self, du, rho, _unused, s):
        io = _HostIO(rho, du)
        self.ctx.check(self.ctx.lib.oqb_ame_eig_acc(self._plan(float(s)), io.u.B, io.u.ptr, io.du.ptr, 0))
        return io.finish()


class CorrelatedDaviesGenerator:
    """src/opensys/davies.jl:220-296 (also ConstHCorrelatedDaviesGenerator :298-345).  coupling: K matrices (or
    callables of s); gamma, S: dicts keyed by the 1-based (α,β) pairs of `inds` (array-wise closures welcome).
    Evaluated through the general closed-form tables (`oqb_ame_set_tables`): for a gap structure whose non-zero
    groups are single pairs,  X = Cm∘ρ̃ + diag(W·diag ρ̃)  with
        W[a,b]   = Σ_(α,β) γ_αβ(ω_ab)·A_β[a,b]·conj(A_α[a,b])          (a ≠ b)
        Cm[a,a'] = −i(w_a−w_a') + Σ γ_αβ(0)A_β[a,a]conj(A_α[a',a']) − ½(G_a+G_a') − i(h_a−h_a'),
        G_b = Σ_a W[a,b] + Σ γ_αβ(0)conj(A_α[b,b])A_β[b,b],   h_b likewise with S.
    Degenerate gap groups (several pairs per group) are not on the device path."""

    def __init__(self, coupling, gamma, S, inds, digits=8, sigdigits=8, ctx: Optional[Context] = None):
        self.coupling, self.gamma, self.S = list(coupling), gamma, S
        self.inds = [tuple(x) for x in inds]
        self.digits, self.sigdigits = digits, sigdigits
        self.ctx = ctx or get_context()
        self._ame, self._key = None, None

    def _plan(self, w, v, s):
        lib, ctx = self.ctx.lib, self.ctx
        key = (float(s), tuple(np.asarray(w).tolist()), None if v is None else id(v))
        if self._ame is not None and self._key == key:
            return self._ame
        if self._ame is not None:
            lib.oqb_ame_destroy(self._ame)
            self._ame = None
        coup = [np.asarray(c(s) if callable(c) else c, dtype=C128) for c in self.coupling]
        w = np.ascontiguousarray(np.real(w), dtype=np.float64)
        l, n, K = len(w), coup[0].shape[0], len(coup)
        groups = GapGroups(w, self.digits, self.sigdigits)
        if len(groups.cross_idx):
            raise NotImplementedError("CorrelatedDaviesGenerator with degenerate gap groups is not on the device path")
        h = C.c_void_p()
        cbuf = _colmajor_stack(coup)
        ctx.check(lib.oqb_ame_create(ctx.h, n, l, K, cbuf.ctypes.data, C.byref(h)))
        if v is None:
            ctx.check(lib.oqb_ame_set_eigen(h, _lib.dptr(w), None))
        else:
            vbuf = np.ascontiguousarray(np.asarray(v, dtype=C128).T)
            ctx.check(lib.oqb_ame_set_eigen(h, _lib.dptr(w), vbuf.ctypes.data))
        Abuf = np.empty((K, l, l), dtype=C128)
        ctx.check(lib.oqb_ame_get_rotated_couplings(h, Abuf.ctypes.data))
        A = Abuf.transpose(0, 2, 1)  # (K, row, col)
        om = groups.omega
        off = ~np.eye(l, dtype=bool)
        W = np.zeros((l, l), dtype=C128)
        Hoff = np.zeros((l, l), dtype=C128)
        Z = np.zeros((l, l), dtype=C128)
        g0d = np.zeros(l, dtype=C128)
        h0d = np.zeros(l, dtype=C128)
        zero = np.zeros(1)
        for (al, be) in self.inds:
            Aa, Ab = A[al - 1], A[be - 1]
            prod = Ab * Aa.conj()
            W += np.where(off, _eval(self.gamma[(al, be)], om.ravel()).reshape(l, l), 0.0) * prod
            Hoff += np.where(off, _eval(self.S[(al, be)], om.ravel()).reshape(l, l), 0.0) * prod
            g0 = float(_eval(self.gamma[(al, be)], zero)[0])
            s0 = float(_eval(self.S[(al, be)], zero)[0])
            da, db = np.diag(Aa), np.diag(Ab)
            Z += g0 * np.outer(db, da.conj())
            g0d += g0 * da.conj() * db
            h0d += s0 * da.conj() * db
        G = W.sum(axis=0) + g0d
        hl = Hoff.sum(axis=0) + h0d
        Cm = -1j * (w[:, None] - w[None, :]) + Z - 0.5 * (G[:, None] + G[None, :]) - 1j * (hl[:, None] - hl[None, :])
        cmbuf = np.ascontiguousarray(Cm.T)
        wre, wim = np.ascontiguousarray(W.real), np.ascontiguousarray(W.imag)
        ctx.check(lib.oqb_ame_set_tables(h, cmbuf.ctypes.data, _lib.dptr(wre), _lib.dptr(wim)))
        self._ame, self._key = h, key
        return h

    def __call__(self, du, rho, w, *args):
        """(D)(du, ρ, gap_idx, s) — adiabatic frame, ρ in the eigenbasis — or (D)(du, ρ̃, gap_idx, v, s); `w` stands in
        for the GapIndices argument (its energies).  ACCUMULATES."""
        v, s = (None, args[0]) if len(args) == 1 else args
        io = _HostIO(rho, du)
        self.ctx.check(self.ctx.lib.oqb_ame_eig_acc(self._plan(w, v, s), io.u.B, io.u.ptr, io.du.ptr, 0))
        return io.finish()


class FluctuatorLiouvillian:
    """coupling: K constant matrices (unit-scaled); n: noise values, (K,) shared or (B, K) one realisation per member
    (each trajectory of a stochastic ensemble carries its own fluctuator state)."""

    def __init__(self, coupling: Sequence[np.ndarray], n, ctx: Optional[Context] = None):
        self.mats = [np.asarray(m, dtype=C128) for m in coupling]
        self.n = np.atleast_2d(np.asarray(n, dtype=np.float64))
        self.ctx = ctx or get_context()
        h = C.c_void_p()
        mbuf = _colmajor_stack(self.mats)  # keep the buffer alive across the call
        self.ctx.check(self.ctx.lib.oqb_dense_create(self.ctx.h, self.mats[0].shape[0], len(self.mats),
                                                     mbuf.ctypes.data, 0, None, C.byref(h)))
        self._op = h

    def _coef(self, B):
        n = np.ascontiguousarray(self.n)
        if n.shape[0] not in (1, B):
            raise ValueError("noise values must be (K,) or (B, K)")
        return n, int(n.shape[0] == 1)

    def __call__(self, du, u, p, t):
        """(F::FluctuatorLiouvillian)(du,u,p,t) :25-31 — OVERWRITES du."""
        io = _HostIO(u, du)
        cf, shared = self._coef(io.u.B)
        self.ctx.check(self.ctx.lib.oqb_dense_rhs(self._op, io.u.B, _lib.dptr(cf), shared, None, 1, io.u.ptr, io.du.ptr))
        return io.finish()

    def update_cache(self, cache, p, t):
        """update_cache!(cache, F, p, t) :34-37 — ACCUMULATES −iΣ n_i c_i."""
        io = _HostIO(cache, cache)
        cf, shared = self._coef(io.u.B)
        tmp = io.u.zeros_like()
        self.ctx.check(self.ctx.lib.oqb_dense_update_cache(self._op, io.u.B, _lib.dptr(cf), shared, None, 1, tmp.ptr))
        _axpy(self.ctx, tmp, io.du)
        return io.finish()

    def update_vectorized_cache(self, cache, p, t):
        """update_vectorized_cache!(cache, F, p, t) :39-44 — ACCUMULATES i(Hᵀ⊗I − I⊗H)."""
        io = _HostIO(cache, cache)
        cf, shared = self._coef(io.u.B)
        tmp = io.u.zeros_like()
        self.ctx.check(self.ctx.lib.oqb_dense_update_vectorized_cache(self._op, io.u.B, _lib.dptr(cf), shared, tmp.ptr))
        _axpy(self.ctx, tmp, io.du)
        return io.finish()
