"""Host-side gap preprocessing for the Davies generator (what the Julia host does with
`GapIndices`, reference src/opensys/opensys_util.jl:25-61), in the sort-based O(l² log l) form the
device tables need.

Produces, for an energy vector w:
  * omega[a,b]: the SIGNED ROUNDED gap the reference evaluates γ and S at for the pair (row a, col b):
    for i<j the key is round(w[j]-w[i], sigdigits) and the pair (j,i) inherits −key; pairs with
    |gap| ≤ 10^-digits (and the diagonal) form the zero group with omega = 0;
  * the cross terms of gap groups that hold more than one pair (accidental 8-digit coincidences,
    degeneracies): `cross` rows (a,b,a2,b2) and `lamb` rows (a,b,b2), see include/oqb.h.
This is index bookkeeping on l² scalars; no state arithmetic happens on the host.
"""
import math

import numpy as np

_POW10_TABLE = np.array([float(f"1e{k}") for k in range(0, 31)])


def round_sigdigits(x: np.ndarray, sigdigits: int) -> np.ndarray:
    """Vectorised Julia `round(x, sigdigits=n)` for Float64 (Base floatfuncs.jl):
    h = 1+floor(log10|x|); d = n−h; d ≥ 0 ? rint(x·10^d)/10^d : rint(x/10^-d)·10^-d."""
    x = np.asarray(x, dtype=np.float64)
    out = np.array(x, copy=True)
    nz = (x != 0) & np.isfinite(x)
    if not nz.any():
        return out
    xv = x[nz]
    h = 1 + np.floor(np.log10(np.abs(xv))).astype(np.int64)
    d = sigdigits - h
    scale = _POW10_TABLE[np.minimum(np.abs(d), 30)]  # exact powers of ten (10^k is exact for k ≤ 22)
    pos = d >= 0
    r = np.empty_like(xv)
    r[pos] = np.rint(xv[pos] * scale[pos]) / scale[pos]
    r[~pos] = np.rint(xv[~pos] / scale[~pos]) * scale[~pos]
    bad = ~np.isfinite(r)
    r[bad] = xv[bad]
    out[nz] = r
    return out


def _group_cross_lists(mi, mj, starts, ends, keys):
    """Cross-term and Lamb-shift lists of all multi-pair gap groups at once (the vectorised form of
    `GapGroups._pairs`): groups are slices [starts[g], ends[g]) of the pair arrays (mi, mj) with key keys[g]; per group
    the L₊ entries (rows mi, cols mj, +key) come first, then the L₋ entries (rows mj, cols mi, −key); inside each the
    ordered pairs (p, q), p ≠ q, with p outermost — the order the loop produces."""
    sizes = (ends - starts).astype(np.int64)
    cr, ck, cg, lm, lk, lg = [], [], [], [], [], []
    for m in np.unique(sizes):
        if m < 2:
            continue
        gsel = np.flatnonzero(sizes == m)
        idx = starts[gsel][:, None] + np.arange(m)[None, :]
        I, J, kv = mi[idx], mj[idx], keys[gsel]
        P, Q = np.nonzero(~np.eye(int(m), dtype=bool))
        for sign, (aa, bb) in enumerate(((I, J), (J, I))):
            ent = np.stack([aa[:, P], bb[:, P], aa[:, Q], bb[:, Q]], axis=-1)  # (G, m(m−1), 4)
            kk = np.repeat((kv if sign == 0 else -kv)[:, None], len(P), axis=1)
            gid = np.repeat((2 * gsel + sign)[:, None], len(P), axis=1)
            cr.append(ent.reshape(-1, 4)); ck.append(kk.ravel()); cg.append(gid.ravel())
            same = (aa[:, P] == aa[:, Q])
            lm.append(ent[same][:, [0, 1, 3]]); lk.append(kk[same]); lg.append(gid[same])
    if not cr:
        return (np.zeros((0, 4), dtype=np.int32), np.zeros(0), np.zeros((0, 3), dtype=np.int32), np.zeros(0))
    cr, ck, cg = np.concatenate(cr), np.concatenate(ck), np.concatenate(cg)
    lm, lk, lg = np.concatenate(lm), np.concatenate(lk), np.concatenate(lg)
    oc, ol = np.argsort(cg, kind="stable"), np.argsort(lg, kind="stable")
    return cr[oc].astype(np.int32), ck[oc], lm[ol].astype(np.int32).reshape(-1, 3), lk[ol]


def _zero_group_lists(aa, bb):
    """The same for the zero-gap group (degenerate pairs in both orders + the diagonal): ordered pairs p ≠ q, minus the
    (k,k)×(k',k') combinations, which are the γ(0)·A_kk·conj(A_k'k') Hadamard term."""
    aa, bb = np.asarray(aa, dtype=np.int64), np.asarray(bb, dtype=np.int64)
    m = len(aa)
    diag = aa == bb
    keep = ~np.eye(m, dtype=bool) & ~(diag[:, None] & diag[None, :])
    P, Q = np.nonzero(keep)
    ent = np.stack([aa[P], bb[P], aa[Q], bb[Q]], axis=-1)
    same = aa[P] == aa[Q]
    return (ent.astype(np.int32).reshape(-1, 4), np.zeros(len(P)), ent[same][:, [0, 1, 3]].astype(np.int32).reshape(-1, 3),
            np.zeros(int(same.sum())))


class GapGroups:
    def __init__(self, w, digits=8, sigdigits=8):
        w = np.asarray(w, dtype=np.float64)
        l = len(w)
        self.w, self.lvl = w, l
        iu, ju = np.triu_indices(l, 1)  # (i,j), i<j, lexicographic — the reference's loop order
        gap = w[ju] - w[iu]
        zero = np.abs(gap) <= 10.0 ** (-digits)
        key = np.zeros_like(gap)
        key[~zero] = round_sigdigits(gap[~zero], sigdigits)
        omega = np.zeros((l, l))
        omega[iu, ju] = key
        omega[ju, iu] = -key
        self.omega = omega
        self.gap_matrix = w[None, :] - w[:, None]  # UNROUNDED (opensys_util.jl:65), one-sided AME only
        self._iu, self._ju, self._zero, self._key = iu, ju, zero, key
        self._lists = None

    def _build_lists(self):
        """Cross-term / L'L lists of the multi-pair groups — host-side statement of what csrc/ame_terms.cu builds on the
        device (used by the CPU tests; O(Σ m²) entries, so only meant for spectra whose groups are small)."""
        iu, ju, zero, key, l = self._iu, self._ju, self._zero, self._key, self.lvl
        parts = []
        nzi, nzj, nzk = iu[~zero], ju[~zero], key[~zero]
        if nzk.size:
            order = np.argsort(nzk, kind="stable")
            ks = nzk[order]
            starts = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]])
            ends = np.r_[starts[1:], ks.size]
            multi = np.flatnonzero(ends - starts > 1)
            parts.append(_group_cross_lists(nzi[order], nzj[order], starts[multi], ends[multi], ks[starts[multi]]))
        # ---- zero group: degenerate pairs (both orders) + the diagonal ---------------------------
        zi, zj = iu[zero], ju[zero]
        if zi.size:
            aa = np.r_[zi, zj, np.arange(l)]
            bb = np.r_[zj, zi, np.arange(l)]
            parts.append(_zero_group_lists(aa, bb))
        self._lists = (
            np.concatenate([q[0] for q in parts]) if parts else np.zeros((0, 4), dtype=np.int32),
            np.concatenate([q[1] for q in parts]) if parts else np.zeros(0),
            np.concatenate([q[2] for q in parts]) if parts else np.zeros((0, 3), dtype=np.int32),
            np.concatenate([q[3] for q in parts]) if parts else np.zeros(0))

    def _list(self, k):
        if self._lists is None:
            self._build_lists()
        return self._lists[k]

    cross_idx = property(lambda self: self._list(0))
    cross_key = property(lambda self: self._list(1))
    lamb_idx = property(lambda self: self._list(2))
    lamb_key = property(lambda self: self._list(3))

    @staticmethod
    def _pairs(aa, bb, kval, cross, cross_key, lamb, lamb_key, skip_diag_diag=False):
        m = len(aa)
        for p in range(m):
            for q in range(m):
                if p == q:
                    continue
                if skip_diag_diag and aa[p] == bb[p] and aa[q] == bb[q]:
                    continue  # (k,k)×(k',k') is the γ(0)·A_kk·conj(A_k'k') Hadamard term
                cross.append((aa[p], bb[p], aa[q], bb[q]))
                cross_key.append(kval)
                if aa[p] == aa[q]:
                    lamb.append((aa[p], bb[p], bb[q]))
                    lamb_key.append(kval)

    def tables(self, gamma, S):
        """γ and S evaluated at omega (l×l, column-major flattening is done by the caller) and at the
        cross-term keys."""
        n2 = self.lvl * self.lvl
        nc = len(self.cross_key)
        if hasattr(gamma, "vectorized") and hasattr(S, "vectorized"):
            # array-wise closures: evaluating all l² keys directly is cheaper than sorting them for np.unique
            allk = np.r_[self.omega.ravel(), self.cross_key, self.lamb_key]
            gv, sv, inv = _eval(gamma, allk), _eval(S, allk), np.arange(allk.size)
        else:
            uniq, inv = np.unique(np.r_[self.omega.ravel(), self.cross_key, self.lamb_key], return_inverse=True)
            gv = _eval(gamma, uniq)
            sv = _eval(S, uniq)
        gam = gv[inv[:n2]].reshape(self.lvl, self.lvl)
        Sm = sv[inv[:n2]].reshape(self.lvl, self.lvl)
        cross_rate = gv[inv[n2:n2 + nc]]
        lamb_rate_S = np.stack([gv[inv[n2 + nc:]], sv[inv[n2 + nc:]]], axis=1) if len(self.lamb_key) else np.zeros((0, 2))
        return gam, Sm, np.ascontiguousarray(cross_rate), np.ascontiguousarray(lamb_rate_S)


def jump_slots(w, K, digits=8, sigdigits=8):
    """Probability slots of `ame_jump` (reference src/opensys/trajectory_jump.jl:14-51) in the reference's order,
    from the sort-based grouping: unique non-zero gap keys ascending (GapIndices keeps `uniq_w` sorted,
    src/opensys/opensys_util.jl:41-54), pairs inside a group in the (i<j) lexicographic loop order; per group and
    coupling the slots L₊ (rows a, cols b, rate key +ω) and L₋ (rows b, cols a, rate key −ω); then per coupling the
    zero-gap slot (degenerate pairs in both orders + the diagonal).  Returns (slot_ptr int64, term int32 (T,3) =
    (coupling,row,col) 0-based sorted by (row,col) inside a slot, rate_key float64 per slot, tags) where tags[s] =
    (coupling, rows, cols) 0-based in list order (what the reference's `tag` holds, 1-based)."""
    w = np.asarray(w, dtype=np.float64)
    l = len(w)
    iu, ju = np.triu_indices(l, 1)
    gap = w[ju] - w[iu]
    zero = np.abs(gap) <= 10.0 ** (-digits)
    key = np.zeros_like(gap)
    key[~zero] = round_sigdigits(gap[~zero], sigdigits)
    nzi, nzj, nzk = iu[~zero], ju[~zero], key[~zero]
    order = np.argsort(nzk, kind="stable")  # stable: keeps the loop order inside a group
    ks = nzk[order]
    starts = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]]) if ks.size else np.zeros(0, dtype=np.int64)
    ends = np.r_[starts[1:], ks.size] if ks.size else np.zeros(0, dtype=np.int64)
    slot_ptr, terms, rate_key, tags = [0], [], [], []

    def add(i, rows, cols, kval):
        o = np.lexsort((cols, rows))
        for r, c in zip(rows[o], cols[o]):
            terms.append((i, int(r), int(c)))
        slot_ptr.append(len(terms))
        rate_key.append(kval)
        tags.append((i, rows, cols))
    for g in range(len(starts)):
        mem = order[starts[g]:ends[g]]
        a, b, kval = nzi[mem], nzj[mem], float(ks[starts[g]])
        for i in range(K):
            add(i, a, b, kval)
            add(i, b, a, -kval)
    zi, zj = iu[zero], ju[zero]
    a0 = np.r_[np.stack([zi, zj], axis=1).ravel(), np.arange(l)].astype(np.int64)  # push i, push j per pair (:34-37)
    b0 = np.r_[np.stack([zj, zi], axis=1).ravel(), np.arange(l)].astype(np.int64)
    for i in range(K):
        add(i, a0, b0, 0.0)
    return (np.asarray(slot_ptr, dtype=np.int64), np.asarray(terms, dtype=np.int32).reshape(-1, 3),
            np.asarray(rate_key, dtype=np.float64), tags)


def _eval(f, x):
    """Evaluate a scalar host closure on an array (vectorised when the object offers it)."""
    if hasattr(f, "vectorized"):
        return np.asarray(f.vectorized(x), dtype=np.float64)
    return np.array([float(f(float(v))) for v in x], dtype=np.float64)


def ohmic_spectrum_vec(eta, wc, beta, w):
    """Vectorised src/bath/ohmic.jl:35-41."""
    w = np.asarray(w, dtype=np.float64)
    out = np.full(w.shape, 2 * math.pi * eta / beta)
    nz = np.abs(w) > 1e-9
    wv = w[nz]
    out[nz] = 2 * math.pi * eta * wv * np.exp(-np.abs(wv) / wc) / (1 - np.exp(-beta * wv))
    return out
