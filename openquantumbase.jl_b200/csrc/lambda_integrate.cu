// The adaptive control flow of the Redfield / ULE kernel integrals inside the library:
// what `quadgk!` does in (R::RedfieldLiouvillian)(du,u,p,t), reference src/opensys/redfield.jl:46-64 (and
// src/opensys/lindblad.jl:41-69 for ULE) — Gauss–Kronrod (7,15), one max-heap of segments per integral, bisect the segment
// with the largest error until E ≤ atol or E ≤ rtol‖I‖ (QuadGK.jl 2.x `adapt`), final re-summation over the
// heap — for ALL integrals in lock step: every refinement round is one oqb_lambda_eval launch set.  The correlation function
// C_ij(t, x)·f_j(p(x)) is a C callback evaluated on the calling thread for all new nodes of a round at once (Julia passes an
// @cfunction pointer), or the Ohmic bath's C(t − x) evaluated by the library (src/bath/ohmic.jl:48-52).
//
// The heap is CPython-heapq-shaped on purpose (same sift rules, ties by insertion tick): the final re-summation runs over
// the heap ARRAY order, so this driver and the Python host mirror it replaces add the segments in the same order.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/oqb.h"
#include "common.cuh"

using namespace oqb;

namespace {

// Kronrod abscissae (non-negative half, descending), Kronrod weights, Gauss weights — QuadGK.kronrod(7) / QUADPACK
const double kX[8] = {0.991455371120812639206854697526329, 0.949107912342758524526189684047851, 0.864864423359769072789712788640926,
                      0.741531185599394439863864773280788, 0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                      0.207784955007898467600689403773245, 0.0};
const double kWK[8] = {0.022935322010529224963732008058970, 0.063092092629978553290700663189204, 0.104790010322250183839876322541518,
                       0.140653259715525918745189590510238, 0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                       0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
const double kWG[4] = {0.129484966168869693270611432679082, 0.279705391489276667901467771423780, 0.381830050505118944950369775488975,
                       0.417959183673469387755102040816327};

struct Nodes {
    double xi[15], wk[15], wg[15];
    Nodes() {
        // ascending order on [−1, 1] (the Gauss nodes are the odd positions) …
        double x[15], k[15], g[15];
        for (int i = 0; i < 7; ++i) { x[i] = -kX[i]; k[i] = kWK[i]; x[14 - i] = kX[i]; k[14 - i] = kWK[i]; }
        x[7] = 0.0; k[7] = kWK[7];
        for (int i = 0; i < 15; ++i) g[i] = 0.0;
        g[1] = g[13] = kWG[0]; g[3] = g[11] = kWG[1]; g[5] = g[9] = kWG[2]; g[7] = kWG[3];
        // … then GAUSS-FIRST (7 Gauss nodes, 8 Kronrod-only nodes): the layout oqb_lambda_eval takes
        int o = 0;
        for (int i = 1; i < 15; i += 2, ++o) { xi[o] = x[i]; wk[o] = k[i]; wg[o] = g[i]; }
        for (int i = 0; i < 15; i += 2, ++o) { xi[o] = x[i]; wk[o] = k[i]; wg[o] = g[i]; }
    }
};

struct Seg {
    double negE;
    long long tick;
    double a, b;
    long long slot;
    double E;
};
inline bool seg_less(const Seg& x, const Seg& y) { return x.negE < y.negE || (x.negE == y.negE && x.tick < y.tick); }

// CPython heapq
void sift_down(std::vector<Seg>& h, size_t start, size_t pos) {
    Seg item = h[pos];
    while (pos > start) {
        const size_t parent = (pos - 1) >> 1;
        if (seg_less(item, h[parent])) { h[pos] = h[parent]; pos = parent; } else break;
    }
    h[pos] = item;
}
void sift_up(std::vector<Seg>& h, size_t pos) {
    const size_t end = h.size(), start = pos;
    Seg item = h[pos];
    size_t child = 2 * pos + 1;
    while (child < end) {
        const size_t right = child + 1;
        if (right < end && !seg_less(h[child], h[right])) child = right;
        h[pos] = h[child];
        pos = child;
        child = 2 * pos + 1;
    }
    h[pos] = item;
    sift_down(h, start, pos);
}
void heap_push(std::vector<Seg>& h, const Seg& s) { h.push_back(s); sift_down(h, 0, h.size() - 1); }
Seg heap_pop(std::vector<Seg>& h) {
    Seg last = h.back();
    h.pop_back();
    if (h.empty()) return last;
    Seg top = h[0];
    h[0] = last;
    sift_up(h, 0);
    return top;
}

struct OhmicUser { oqb_ctx* c; double eta, wc, beta; };

}  // namespace

extern "C" {

int oqb_lambda_integrate(oqb_lambda* lam, oqb_ctx* c, int nint, const int32_t* h_member, const int32_t* h_coupling, const double* h_t,
                         const double* h_lo, const double* h_hi, const double* h_dt, int G, oqb_kernel_fn cfun, void* user, double atol,
                         double rtol, int64_t maxevals, void* d_Lambda, int64_t* rounds_out, int64_t* segments_out, int64_t* h_evals) {
    if (!lam || !c) return OQB_EINVAL;
    OQB_DEVICE(c, c);
    if (nint <= 0) return 0;
    if (!h_member || !h_coupling || !h_t || !h_lo || !h_hi || !h_dt || !cfun || !d_Lambda || G < 1)
        return set_error(c, OQB_EINVAL, "oqb_lambda_integrate: null argument");
    static const Nodes nodes;
    if (maxevals <= 0) maxevals = 10000000;
    const long long nn = 0;  // (matrix size is the handle's business)
    (void)nn;
    std::vector<int> live;
    for (int q = 0; q < nint; ++q)
        if (h_lo[q] < h_hi[q]) live.push_back(q);  // a == b ⇒ Λ = 0 (quadgk returns zero(f(a)))
    OQB_TRY(oqb_lambda_reserve_slots(lam, 2ll * nint));
    long long next_slot = nint;
    std::vector<std::vector<Seg>> heaps(nint);
    std::vector<double> Etot(nint, 0.0), nrm(nint, 0.0);
    std::vector<long long> evals(nint, 0), tick(nint, 0);

    // one Gauss–Kronrod evaluation of a set of segments (q, a, b, slot) → E
    std::vector<int32_t> mem, cpl, gi, qidx;
    std::vector<double> th, ck, cg, x, tt, cv, E;
    std::vector<int64_t> slot;
    auto eval = [&](const std::vector<int>& qs, const std::vector<double>& sa, const std::vector<double>& sb,
                    const std::vector<long long>& sl) -> int {
        const size_t ns = qs.size();
        mem.resize(ns); cpl.resize(ns); qidx.resize(ns); tt.resize(ns); slot.resize(ns); E.resize(ns);
        gi.resize(ns * 16); th.resize(ns * 16); x.resize(ns * 15); cv.resize(ns * 30); ck.resize(ns * 30); cg.resize(ns * 14);
        for (size_t s = 0; s < ns; ++s) {
            const int q = qs[s];
            mem[s] = h_member[q]; cpl[s] = h_coupling[q]; qidx[s] = q; tt[s] = h_t[q]; slot[s] = sl[s];
            const double h = 0.5 * (sb[s] - sa[s]);
            for (int i = 0; i < 15; ++i) x[s * 15 + i] = sa[s] + (1.0 + nodes.xi[i]) * h;
        }
        cfun(user, (int64_t)ns, qidx.data(), tt.data(), x.data(), cv.data());
        for (size_t s = 0; s < ns; ++s) {
            const double h = 0.5 * (sb[s] - sa[s]);
            const double dt = h_dt[mem[s]];
            for (int i = 0; i < 16; ++i) {  // the 15 nodes, then the end time t (U(t) is applied once per segment)
                const double pos = (i < 15 ? x[s * 15 + i] : tt[s]) / dt;
                double g = std::floor(pos);
                double frac = pos - g;
                if (g >= (double)(G - 1)) { g = (double)(G - 1); frac = 0.0; }
                gi[s * 16 + i] = (int32_t)g;
                th[s * 16 + i] = frac;
            }
            for (int i = 0; i < 15; ++i) {
                const double wk = nodes.wk[i] * h;
                ck[s * 30 + 2 * i] = wk * cv[s * 30 + 2 * i];
                ck[s * 30 + 2 * i + 1] = wk * cv[s * 30 + 2 * i + 1];
            }
            for (int i = 0; i < 7; ++i) {
                const double wg = nodes.wg[i] * h;
                cg[s * 14 + 2 * i] = wg * cv[s * 30 + 2 * i];
                cg[s * 14 + 2 * i + 1] = wg * cv[s * 30 + 2 * i + 1];
            }
        }
        return oqb_lambda_eval(lam, (int)ns, mem.data(), cpl.data(), gi.data(), th.data(), ck.data(), cg.data(), slot.data(), E.data());
    };
    // Λ_q = Σ over the heap array (QuadGK's final re-summation order); optionally ‖Λ_q‖_F
    auto totals = [&](const std::vector<int>& qs, void* d_out, std::vector<double>* norms) -> int {
        std::vector<int64_t> ptr(qs.size() + 1, 0), idx;
        for (size_t k = 0; k < qs.size(); ++k) {
            for (const Seg& s : heaps[qs[k]]) idx.push_back(s.slot);
            ptr[k + 1] = (int64_t)idx.size();
        }
        if (norms) norms->assign(qs.size(), 0.0);
        return oqb_lambda_total(lam, (int)qs.size(), ptr.data(), idx.data(), d_out, norms ? norms->data() : nullptr);
    };

    int n_mat = 0, K_dummy = 0;
    OQB_TRY(oqb_lambda_dims(lam, &n_mat, &K_dummy));
    const size_t mat_bytes = (size_t)n_mat * n_mat * sizeof(c128);
    if (live.size() < (size_t)nint) OQB_CUDA(c, cudaMemsetAsync(d_Lambda, 0, mat_bytes * nint, c->stream));
    long long rounds = 0, nseg_total = 0;
    if (!live.empty()) {
        c128* d_scratch = nullptr;  // totals of the active integrals during refinement
        OQB_CUDA(c, cudaMallocAsync((void**)&d_scratch, mat_bytes * live.size(), c->stream));
        int st = 0;
        do {
            std::vector<double> sa, sb, nv;
            std::vector<long long> sl;
            for (int q : live) { sa.push_back(h_lo[q]); sb.push_back(h_hi[q]); sl.push_back(q); }
            if ((st = eval(live, sa, sb, sl))) break;
            for (size_t k = 0; k < live.size(); ++k) {
                const int q = live[k];
                heaps[q].push_back(Seg{-E[k], 0, h_lo[q], h_hi[q], (long long)q, E[k]});
                Etot[q] = E[k];
                evals[q] = 15;
                tick[q] = 1;
            }
            if ((st = totals(live, d_scratch, &nv))) break;
            for (size_t k = 0; k < live.size(); ++k) nrm[live[k]] = nv[k];
            rounds = 1;
            nseg_total = (long long)live.size();
            std::vector<int> active;
            auto still = [&](int q) { return Etot[q] > atol && Etot[q] > rtol * nrm[q] && evals[q] < maxevals; };
            for (int q : live)
                if (still(q)) active.push_back(q);
            while (!active.empty()) {
                if ((st = oqb_lambda_reserve_slots(lam, next_slot + (long long)active.size()))) break;
                std::vector<int> qs;
                std::vector<Seg> popped;
                std::vector<long long> second;
                sa.clear(); sb.clear(); sl.clear();
                for (int q : active) {
                    const Seg s = heap_pop(heaps[q]);
                    const double mid = (s.a + s.b) * 0.5;
                    popped.push_back(s);
                    second.push_back(next_slot);
                    qs.push_back(q); qs.push_back(q);
                    sa.push_back(s.a); sa.push_back(mid);
                    sb.push_back(mid); sb.push_back(s.b);
                    sl.push_back(s.slot); sl.push_back(next_slot);
                    ++next_slot;
                }
                // runs of equal coupling → fewer launches for dense S_j (stable: keeps the pairing of the two halves findable)
                std::vector<size_t> order(qs.size());
                for (size_t i = 0; i < order.size(); ++i) order[i] = i;
                std::stable_sort(order.begin(), order.end(), [&](size_t i, size_t j) { return h_coupling[qs[i]] < h_coupling[qs[j]]; });
                std::vector<int> qs2(qs.size());
                std::vector<double> sa2(qs.size()), sb2(qs.size());
                std::vector<long long> sl2(qs.size());
                for (size_t i = 0; i < order.size(); ++i) { qs2[i] = qs[order[i]]; sa2[i] = sa[order[i]]; sb2[i] = sb[order[i]]; sl2[i] = sl[order[i]]; }
                if ((st = eval(qs2, sa2, sb2, sl2))) break;
                std::vector<double> Eo(qs.size());
                for (size_t i = 0; i < order.size(); ++i) Eo[order[i]] = E[i];
                for (size_t k = 0; k < active.size(); ++k) {
                    const int q = active[k];
                    const Seg& s = popped[k];
                    const double mid = (s.a + s.b) * 0.5;
                    const double E1 = Eo[2 * k], E2 = Eo[2 * k + 1];
                    evals[q] += 30;
                    Etot[q] = (Etot[q] - s.E) + E1 + E2;
                    heap_push(heaps[q], Seg{-E1, tick[q]++, s.a, mid, s.slot, E1});
                    heap_push(heaps[q], Seg{-E2, tick[q]++, mid, s.b, second[k], E2});
                }
                if ((st = totals(active, d_scratch, &nv))) break;
                for (size_t k = 0; k < active.size(); ++k) nrm[active[k]] = nv[k];
                ++rounds;
                nseg_total += (long long)qs.size();
                std::vector<int> next;
                for (int q : active)
                    if (still(q)) next.push_back(q);
                active.swap(next);
            }
            if (st) break;
            // final re-summation of every integral in heap order, written in integral order
            if (live.size() == (size_t)nint) {
                st = totals(live, d_Lambda, nullptr);
            } else {
                if ((st = totals(live, d_scratch, nullptr))) break;
                for (size_t k = 0; k < live.size() && !st; ++k)
                    if (cudaMemcpyAsync((char*)d_Lambda + mat_bytes * live[k], (char*)d_scratch + mat_bytes * k, mat_bytes, cudaMemcpyDeviceToDevice,
                                        c->stream) != cudaSuccess)
                        st = set_error(c, OQB_ECUDA, "oqb_lambda_integrate: copy failed");
            }
        } while (0);
        cudaStreamSynchronize(c->stream);
        cudaFreeAsync(d_scratch, c->stream);
        if (st) return st;
    }
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (rounds_out) *rounds_out = rounds;
    if (segments_out) *segments_out = nseg_total;
    if (h_evals)
        for (int q = 0; q < nint; ++q) h_evals[q] = evals[q];
    return 0;
}

// cfun for an Ohmic bath: C(t − x) (src/bath/ohmic.jl:48-52), the cfun of the kernels InteractionSet builds for a single
// Ohmic bath (src/coupling/interaction.jl:178-186).  user → {ctx, η, ωc, β}
static void ohmic_kernel_fn(void* user, int64_t nseg, const int32_t*, const double* t, const double* x, double* out) {
    OhmicUser* u = (OhmicUser*)user;
    std::vector<double> tau((size_t)nseg * 15);
    for (int64_t s = 0; s < nseg; ++s)
        for (int i = 0; i < 15; ++i) tau[s * 15 + i] = t[s] - x[s * 15 + i];
    if (oqb_ohmic_correlation(u->c, u->eta, u->wc, u->beta, (int)tau.size(), tau.data(), out) != 0)
        for (size_t i = 0; i < tau.size() * 2; ++i) out[i] = NAN;  // surfaces as a failed convergence / NaN result
}

int oqb_lambda_integrate_ohmic(oqb_lambda* lam, oqb_ctx* c, int nint, const int32_t* h_member, const int32_t* h_coupling, const double* h_t,
                               const double* h_lo, const double* h_hi, const double* h_dt, int G, double eta, double wc, double beta,
                               double atol, double rtol, int64_t maxevals, void* d_Lambda, int64_t* rounds_out, int64_t* segments_out,
                               int64_t* h_evals) {
    OhmicUser u{c, eta, wc, beta};
    return oqb_lambda_integrate(lam, c, nint, h_member, h_coupling, h_t, h_lo, h_hi, h_dt, G, ohmic_kernel_fn, &u, atol, rtol, maxevals, d_Lambda,
                                rounds_out, segments_out, h_evals);
}

}  // extern "C"
