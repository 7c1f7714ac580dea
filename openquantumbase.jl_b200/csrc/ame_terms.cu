// GapIndices and the Davies term list of a DiffEqLiouvillian{true,false}, built in the library (no host-language
// bookkeeping): SURVEY §8(a) rows G and DV.
//
//   oqb_ame_gaps_build      GapIndices (src/opensys/opensys_util.jl:25-61) of the current eigen-system on the device: the
//                           l(l−1)/2 gaps are thresholded / rounded (Julia's round(x, sigdigits)), radix-sorted (stable, so a
//                           group keeps the reference's (i<j) loop order), numbered, and scattered into an l×l table of signed
//                           group numbers.  oqb_ame_gaps_keys / _fetch hand the result back in the reference's own form
//                           (uniq_w, uniq_a, uniq_b, a0, b0).
//   oqb_ame_add_davies*     one DaviesGenerator (src/opensys/davies.jl:21-47) over a slice of the handle's couplings, with
//                           γ(±x), S(±x) given PER GAP KEY (what the reference's loop evaluates: one closure call per group)
//                           or evaluated on the device for an Ohmic bath.  Several generators accumulate
//                           (src/opensys/diffeq_liouvillian.jl:101-103, src/coupling/interaction.jl:101-117).
//
// Per gap group x the reference applies L_x = A∘[ω̃ == x]:  γ(x)(L_xρL_x' − ½{L_x'L_x, ρ}) − i[S(x)L_x'L_x, ρ].
//   * a group that is a single level pair contributes rank-one terms: the closed form (Hadamard matrix C, rate matrix W);
//   * a group with a few pairs adds the cross terms between its pairs: the cross-term list (a,b,a2,b2) and the
//     off-diagonal part of L_x'L_x (Moff);
//   * a LARGE group (degenerate spectra — H(0) = −ΣX of every anneal has C(2q, q) level pairs in its zero group) is
//     applied the way the reference writes it, as dense products with the masked coupling, stacked over groups.
// The choice per group is a cost model (list entries vs. two l³ products), overridable per context
// ("davies_dense_min_pairs", "davies_max_cross").
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "../../include/oqb.h"
#include "ame.cuh"
#include "bath.cuh"
#include "common.cuh"

using namespace oqb;

namespace {

struct Pow10 { double v[31]; };

Pow10 make_pow10() {
    Pow10 t;
    for (int k = 0; k <= 30; ++k) {  // exactly the decimal literals "1e<k>" (correctly rounded)
        char buf[8];
        snprintf(buf, sizeof(buf), "1e%d", k);
        t.v[k] = strtod(buf, nullptr);
    }
    return t;
}

// Julia Base.round(x, sigdigits=n) for Float64 (floatfuncs.jl: hidigit + _round_digits); x != 0, finite.
__device__ __forceinline__ double round_sigdigits(double x, int sig, const Pow10& t) {
    const double ax = fabs(x);
    int h;
    if (ax >= 1.0) {
        h = 1;
        while (h < 30 && ax >= t.v[h]) ++h;  // 10^(h-1) ≤ |x| < 10^h
    } else {
        int k = 1;
        while (k < 30 && ax * t.v[k] < 1.0) ++k;
        if (k > 1 && ax >= 1.0 / t.v[k - 1]) --k;
        h = 1 - k;
    }
    const int d = sig - h;
    const int ad = d < 0 ? -d : d;
    const double sc = ad <= 30 ? t.v[ad] : pow(10.0, (double)ad);
    const double r = d >= 0 ? rint(x * sc) / sc : rint(x / sc) * sc;
    return isfinite(r) ? r : x;
}

// lexicographic pair number p(i<j) = i·l − i(i+1)/2 + (j−i−1)  →  (i, j)
__device__ __forceinline__ void pair_ij(long long p, int l, int& i, int& j) {
    const double b = 2.0 * l - 1.0;
    long long ii = (long long)floor((b - sqrt(b * b - 8.0 * (double)p)) * 0.5);
    if (ii < 0) ii = 0;
    if (ii > l - 2) ii = l - 2;
    while (ii > 0 && ii * l - ii * (ii + 1) / 2 > p) --ii;
    while ((ii + 1) * l - (ii + 1) * (ii + 2) / 2 <= p) ++ii;
    i = (int)ii;
    j = (int)(p - (ii * l - ii * (ii + 1) / 2)) + i + 1;
}

__global__ void gap_keys_kernel(int l, const double* __restrict__ w, double thr, int sig, Pow10 t, double* __restrict__ keys,
                                int32_t* __restrict__ pairs, unsigned long long* __restrict__ nzero) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)l * l) return;
    const int i = (int)(e % l), j = (int)(e / l);
    if (i >= j) return;
    const double gap = w[j] - w[i];  // the reference's loop variable (i<j), opensys_util.jl:31
    const bool zero = fabs(gap) <= thr;
    const long long p = (long long)i * l - (long long)i * (i + 1) / 2 + (j - i - 1);
    keys[p] = zero ? INFINITY : round_sigdigits(gap, sig, t);
    pairs[p] = (int32_t)p;
    if (zero) atomicAdd(nzero, 1ull);
}

__global__ void head_flag_kernel(long long m, const double* __restrict__ keys, int32_t* __restrict__ head) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const double v = keys[k];
    head[k] = (isfinite(v) && (k == 0 || keys[k - 1] != v)) ? 1 : 0;
}

// gid must be zero-filled (zero group and diagonal)
__global__ void scatter_gid_kernel(int l, long long m, const double* __restrict__ keys, const int32_t* __restrict__ pairs,
                                   const int32_t* __restrict__ head, const int32_t* __restrict__ grp, int32_t* __restrict__ gid,
                                   double* __restrict__ ukeys, long long* __restrict__ gptr) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const double v = keys[k];
    if (!isfinite(v)) return;
    int i, j;
    pair_ij(pairs[k], l, i, j);
    const int g = grp[k];  // 1-based group number in ascending key order
    gid[i + (long long)j * l] = g;
    gid[j + (long long)i * l] = -g;
    if (head[k]) {
        ukeys[g - 1] = v;
        gptr[g - 1] = k;
    }
}

// ---- cross-term list of the groups that stay on the list ------------------------------------------------------------
// group t of the list: pairs spair[start[t] .. start[t]+m[t]), entries off[t] .. off[t] + 2m(m−1): first the L₊ entries
// (rows i, cols j), then the L₋ entries (rows j, cols i); inside each the ordered pairs (p, q), p ≠ q.
__global__ void cross_gen_kernel(int l, long long nent, int nlg, const long long* __restrict__ start, const int32_t* __restrict__ msz,
                                 const long long* __restrict__ off, const int32_t* __restrict__ gnum, const int32_t* __restrict__ spair,
                                 int32_t* __restrict__ idx, int32_t* __restrict__ cg) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nent) return;
    int lo = 0, hi = nlg - 1;  // last t with off[t] ≤ e
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (off[mid] <= e) lo = mid; else hi = mid - 1;
    }
    const long long m = msz[lo];
    long long r = e - off[lo];
    const int sign = r >= m * (m - 1) ? 1 : 0;
    r -= sign ? m * (m - 1) : 0;
    const long long p = r / (m - 1);
    long long q = r % (m - 1);
    q += q >= p ? 1 : 0;
    int ip, jp, iq, jq;
    pair_ij(spair[start[lo] + p], l, ip, jp);
    pair_ij(spair[start[lo] + q], l, iq, jq);
    int32_t* o = idx + 4 * e;
    if (!sign) { o[0] = ip; o[1] = jp; o[2] = iq; o[3] = jq; }
    else { o[0] = jp; o[1] = ip; o[2] = jq; o[3] = iq; }
    cg[e] = sign ? -gnum[lo] : gnum[lo];
}

// zero group on the list: members = the nz degenerate pairs in both orders (2nz off-diagonal entries), then the l
// diagonal entries; ordered pairs (p, q), p ≠ q, except diagonal × diagonal (the γ(0)·A_kk·conj(A_k'k') Hadamard term).
__global__ void cross_zero_kernel(int l, long long nz, const int32_t* __restrict__ zpair, long long nent, int32_t* __restrict__ idx,
                                  int32_t* __restrict__ cg) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nent) return;
    const long long mo = 2 * nz, mz = mo + l;
    long long p, q;
    if (e < mo * (mz - 1)) {
        p = e / (mz - 1);
        q = e % (mz - 1);
        q += q >= p ? 1 : 0;
    } else {
        const long long r = e - mo * (mz - 1);
        p = mo + r / mo;
        q = r % mo;
    }
    auto member = [&](long long t, int& a, int& b) {
        if (t >= mo) { a = b = (int)(t - mo); return; }
        int i, j;
        pair_ij(zpair[t >> 1], l, i, j);
        if (t & 1) { a = j; b = i; } else { a = i; b = j; }
    };
    int a, b, a2, b2;
    member(p, a, b);
    member(q, a2, b2);
    int32_t* o = idx + 4 * e;
    o[0] = a; o[1] = b; o[2] = a2; o[3] = b2;
    cg[e] = 0;
}

__global__ void lamb_flag_kernel(long long n, const int32_t* __restrict__ idx, uint8_t* __restrict__ flag) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) flag[e] = idx[4 * e] == idx[4 * e + 2] ? 1 : 0;  // same row a: contributes to (L'L)[b, b2]
}

__global__ void lamb_fill_kernel(long long n, const long long* __restrict__ src, const int32_t* __restrict__ idx, const int32_t* __restrict__ cg,
                                 int32_t* __restrict__ lidx, int32_t* __restrict__ lg) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const long long e = src[k];
    lidx[3 * k] = idx[4 * e];
    lidx[3 * k + 1] = idx[4 * e + 1];
    lidx[3 * k + 2] = idx[4 * e + 3];
    lg[k] = cg[e];
}

// ---- per-term values --------------------------------------------------------------------------------------------------
// vals layout: [γ(+x_g) | γ(−x_g) | S(+x_g) | S(−x_g)] (ng each), then γ(0), S(0)
__global__ void term_vals_ohmic_kernel(long long ng, const double* __restrict__ ukeys, double eta, double wc, double beta, int nS,
                                       double x0, double dx, const double* __restrict__ coef, double* __restrict__ vals) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < ng) {
        const double x = ukeys[g];
        vals[g] = ohmic_spectrum(x, eta, wc, beta);
        vals[ng + g] = ohmic_spectrum(-x, eta, wc, beta);
        vals[2 * ng + g] = nS ? bspline2_eval(x, nS, x0, dx, coef) : 0.0;
        vals[3 * ng + g] = nS ? bspline2_eval(-x, nS, x0, dx, coef) : 0.0;
    }
    if (g == 0) {
        vals[4 * ng] = ohmic_spectrum(0.0, eta, wc, beta);
        vals[4 * ng + 1] = nS ? bspline2_eval(0.0, nS, x0, dx, coef) : 0.0;
    }
}

__device__ __forceinline__ void lookup(int sg, long long ng, const double* __restrict__ vals, double& g, double& S) {
    if (sg == 0) { g = vals[4 * ng]; S = vals[4 * ng + 1]; }
    else if (sg > 0) { g = vals[sg - 1]; S = vals[2 * ng + sg - 1]; }
    else { g = vals[ng - sg - 1]; S = vals[3 * ng - sg - 1]; }
}

// γ, S tables at the l×l signed rounded gaps; entries of groups on the dense path are left out (0) — they are applied
// as dense products instead
__global__ void term_table_kernel(int l, const int32_t* __restrict__ gid, const int32_t* __restrict__ gslot, int zero_dense, long long ng,
                                  const double* __restrict__ vals, double* __restrict__ gam, double* __restrict__ S) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)l * l) return;
    const int sg = gid[e];
    double g = 0.0, s = 0.0;
    const bool dense = sg == 0 ? zero_dense != 0 : gslot[(sg > 0 ? sg : -sg) - 1] >= 0;
    if (!dense) lookup(sg, ng, vals, g, s);
    gam[e] = g;
    S[e] = s;
}

__global__ void term_rate_kernel(long long n, const int32_t* __restrict__ sg, long long ng, const double* __restrict__ vals,
                                 double* __restrict__ rate, double* __restrict__ rate_S) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    double g, s;
    lookup(sg[e], ng, vals, g, s);
    if (rate) rate[e] = g;
    if (rate_S) { rate_S[2 * e] = g; rate_S[2 * e + 1] = s; }
}

// one block per column b of the term's couplings: Γ_t[b] = Σ_a γ_ab Σ_k|A_k[a,b]|², h_t[b] likewise with S; Wt += (transposed)
__global__ void term_colsum_kernel(int l, int K, const c128* __restrict__ A, const double* __restrict__ gam, const double* __restrict__ S,
                                   double* __restrict__ Gam_t, double* __restrict__ hls_t, double* __restrict__ Wt) {
    const int b = blockIdx.x;
    double sg = 0.0, sh = 0.0;
    for (int a = threadIdx.x; a < l; a += blockDim.x) {
        double m2 = 0.0;
        for (int k = 0; k < K; ++k) {
            const c128 v = A[(long long)k * l * l + a + (long long)b * l];
            m2 += v.x * v.x + v.y * v.y;
        }
        const double g = gam[a + (long long)b * l] * m2;
        sg += g;
        sh += S[a + (long long)b * l] * m2;
        if (a != b) Wt[b + (long long)a * l] += g;  // W[a,b] stored transposed; several generators accumulate
    }
    __shared__ double red[2][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sg += __shfl_xor_sync(0xffffffffu, sg, o);
        sh += __shfl_xor_sync(0xffffffffu, sh, o);
    }
    if (lane == 0) { red[0][warp] = sg; red[1][warp] = sh; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a0 = 0.0, a1 = 0.0;
        for (int i = 0; i < (blockDim.x >> 5); ++i) { a0 += red[0][i]; a1 += red[1][i]; }
        Gam_t[b] = a0;
        hls_t[b] = a1;
    }
}

// Cm[a,b] += −½(Γ_t[a]+Γ_t[b]) − i(h_t[a]−h_t[b]) + γ_t(0)·Σ_k A_k[a,a]·conj(A_k[b,b])   (the last term only while the
// zero group is on the list: it is the diagonal×diagonal part of L_0 ρ L_0')
__global__ void term_cmat_acc_kernel(int l, int K, const c128* __restrict__ A, const double* __restrict__ Gam_t, const double* __restrict__ hls_t,
                                     const double* __restrict__ g0p, int zero_dense, c128* __restrict__ Cm) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)l * l) return;
    const int a = (int)(e % l), b = (int)(e / l);
    c128 v = Cm[e];
    v.x += -0.5 * (Gam_t[a] + Gam_t[b]);
    v.y += -(hls_t[a] - hls_t[b]);
    if (!zero_dense) {
        const double g0 = *g0p;
        double sx = 0.0, sy = 0.0;
        for (int k = 0; k < K; ++k) {
            const c128* Ak = A + (long long)k * l * l;
            const c128 p = Ak[a + (long long)a * l], q = Ak[b + (long long)b * l];
            sx += p.x * q.x + p.y * q.y;  // p·conj(q)
            sy += p.y * q.x - p.x * q.y;
        }
        v.x += g0 * sx;
        v.y += g0 * sy;
    }
    Cm[e] = v;
}

// ---- correlated pair terms (CorrelatedDaviesGenerator, src/opensys/davies.jl:231-296): for one pair (α,β) the operators are
// L = A_β∘[ω̃ == x] on the left and L_d' = (A_α∘[ω̃ == x])' on the right: γ(L ρ L_d' − ½{L_d'L, ρ}) − i[S·L_d'L, ρ].  The rank-one
// products A_β[a,b]·conj(A_α[a,b]) are complex, so Γ, h and the rate matrix W are complex, and −½G ∓ iH_LS are two
// different matrices (M1 on the left of ρ, M2 on the right) unless the pair set is symmetric.
// one block per column b:  Γ_t[b] = Σ_a γ_ab conj(A_α[a,b])A_β[a,b],  h_t[b] with S;  Wt(+i·Wt_im)[b + a·l] += γ_ab A_β[a,b]conj(A_α[a,b])
__global__ void corr_colsum_kernel(int l, const c128* __restrict__ Aa, const c128* __restrict__ Ab, const double* __restrict__ gam,
                                   const double* __restrict__ S, c128* __restrict__ Gam_t, c128* __restrict__ hls_t, double* __restrict__ Wt,
                                   double* __restrict__ Wt_im) {
    const int b = blockIdx.x;
    double gr = 0.0, gi = 0.0, hr = 0.0, hi = 0.0;
    for (int a = threadIdx.x; a < l; a += blockDim.x) {
        const c128 x = Aa[a + (long long)b * l], y = Ab[a + (long long)b * l];
        const double pr = x.x * y.x + x.y * y.y, pi = x.x * y.y - x.y * y.x;  // conj(A_α)·A_β
        const double g = gam[a + (long long)b * l], sv = S[a + (long long)b * l];
        gr += g * pr; gi += g * pi;
        hr += sv * pr; hi += sv * pi;
        if (a != b) {  // W[a,b] = γ·A_β·conj(A_α) = the same product (conj(x)·y)
            Wt[b + (long long)a * l] += g * pr;
            Wt_im[b + (long long)a * l] += g * pi;
        }
    }
    __shared__ double red[4][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        gr += __shfl_xor_sync(0xffffffffu, gr, o); gi += __shfl_xor_sync(0xffffffffu, gi, o);
        hr += __shfl_xor_sync(0xffffffffu, hr, o); hi += __shfl_xor_sync(0xffffffffu, hi, o);
    }
    if (lane == 0) { red[0][warp] = gr; red[1][warp] = gi; red[2][warp] = hr; red[3][warp] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = 0; i < (blockDim.x >> 5); ++i)
            for (int k = 0; k < 4; ++k) t[k] += red[k][i];
        Gam_t[b] = make_double2(t[0], t[1]);
        hls_t[b] = make_double2(t[2], t[3]);
    }
}
// Cm[a,b] += (−½Γ_a − i h_a) + (−½Γ_b + i h_b) + γ(0)·A_β[a,a]·conj(A_α[b,b])   (complex Γ, h; last term while the zero group is on the list)
__global__ void corr_cmat_acc_kernel(int l, const c128* __restrict__ Aa, const c128* __restrict__ Ab, const c128* __restrict__ Gam_t,
                                     const c128* __restrict__ hls_t, const double* __restrict__ g0p, int zero_dense, c128* __restrict__ Cm) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)l * l) return;
    const int a = (int)(e % l), b = (int)(e / l);
    c128 v = Cm[e];
    const c128 Ga = Gam_t[a], Gb = Gam_t[b], ha = hls_t[a], hb = hls_t[b];
    v.x += -0.5 * Ga.x + ha.y - 0.5 * Gb.x - hb.y;   // −i·h = (h.y, −h.x);  +i·h = (−h.y, h.x)
    v.y += -0.5 * Ga.y - ha.x - 0.5 * Gb.y + hb.x;
    if (!zero_dense) {
        const double g0 = *g0p;
        const c128 p = Ab[a + (long long)a * l], q = Aa[b + (long long)b * l];
        v.x += g0 * (p.x * q.x + p.y * q.y);
        v.y += g0 * (p.y * q.x - p.x * q.y);
    }
    Cm[e] = v;
}
// X[b][a + a2·l] += rate·A_β[a,b]·conj(A_α[a2,b2])·R[b][b + b2·l]
__global__ void corr_cross_kernel(int l, const c128* __restrict__ Aa, const c128* __restrict__ Ab, long long ncross, const int32_t* __restrict__ idx,
                                  const double* __restrict__ rate, const c128* __restrict__ R, c128* __restrict__ X) {
    const int bb = blockIdx.y;
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= ncross) return;
    const int a = idx[4 * e], b = idx[4 * e + 1], a2 = idx[4 * e + 2], b2 = idx[4 * e + 3];
    const c128 y = Ab[a + (long long)b * l], x = Aa[a2 + (long long)b2 * l];
    const double r = rate[e];
    const double sr = r * (y.x * x.x + y.y * x.y), si = r * (y.y * x.x - y.x * x.y);  // r·y·conj(x)
    const c128 q = R[(long long)bb * l * l + b + (long long)b2 * l];
    c128* d = X + (long long)bb * l * l + a + (long long)a2 * l;
    atomicAdd(&d->x, sr * q.x - si * q.y);
    atomicAdd(&d->y, sr * q.y + si * q.x);
}
// M1[b + b2·l] += (−½γ − iS)·conj(A_α[a,b])·A_β[a,b2],  M2[…] += (−½γ + iS)·(same)
__global__ void corr_lamb_kernel(int l, const c128* __restrict__ Aa, const c128* __restrict__ Ab, long long nlamb, const int32_t* __restrict__ idx,
                                 const double* __restrict__ rs, c128* __restrict__ M1, c128* __restrict__ M2) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= nlamb) return;
    const int a = idx[3 * e], b = idx[3 * e + 1], b2 = idx[3 * e + 2];
    const c128 x = Aa[a + (long long)b * l], y = Ab[a + (long long)b2 * l];
    const double pr = x.x * y.x + x.y * y.y, pi = x.x * y.y - x.y * y.x;  // conj(x)·y
    const double g = -0.5 * rs[2 * e], sv = rs[2 * e + 1];
    atomicAdd(&M1[b + (long long)b2 * l].x, g * pr + sv * pi);   // (g − i·sv)(pr + i·pi)
    atomicAdd(&M1[b + (long long)b2 * l].y, g * pi - sv * pr);
    atomicAdd(&M2[b + (long long)b2 * l].x, g * pr - sv * pi);   // (g + i·sv)(pr + i·pi)
    atomicAdd(&M2[b + (long long)b2 * l].y, g * pi + sv * pr);
}

// out = in' (l×l)
__global__ void adjoint_kernel(int l, const c128* __restrict__ in, c128* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)l * l) return;
    const int a = (int)(e % l), b = (int)(e / l);
    const c128 v = in[b + (long long)a * l];
    out[e] = make_double2(v.x, -v.y);
}

__global__ void vec_add_kernel(int n, const double* __restrict__ x, double* __restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += x[i];
}

// ---- dense per-group blocks -------------------------------------------------------------------------------------------
// signed slot of a table entry: 2d (entries +g), 2d+1 (entries −g) for the d-th dense group, 2·nd_groups for a dense zero group
__device__ __forceinline__ int entry_slot(int sg, const int32_t* __restrict__ gslot, int zero_slot) {
    if (sg == 0) return zero_slot;
    const int d = gslot[(sg > 0 ? sg : -sg) - 1];
    return d < 0 ? -1 : 2 * d + (sg < 0 ? 1 : 0);
}

// norms[slot·K + k] += Σ |A_k[e]|² over the entries of that slot; total[k] += Σ_e |A_k[e]|²
__global__ void dense_norm_kernel(int l, int K, const c128* __restrict__ A, const int32_t* __restrict__ gid, const int32_t* __restrict__ gslot,
                                  int zero_slot, double* __restrict__ norms, double* __restrict__ total) {
    const long long ll = (long long)l * l;
    const int k = blockIdx.y;
    double tot = 0.0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < ll; e += (long long)gridDim.x * blockDim.x) {
        const c128 v = A[(long long)k * ll + e];
        const double m2 = v.x * v.x + v.y * v.y;
        tot += m2;
        if (m2 != 0.0) {
            const int s = entry_slot(gid[e], gslot, zero_slot);
            if (s >= 0) atomicAdd(&norms[(long long)s * K + k], m2);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if ((threadIdx.x & 31) == 0 && tot != 0.0) atomicAdd(&total[k], tot);
}

// out[j] = A_{bk[j]} ∘ [slot(entry) == bslot[j]]      (grid.y = block j)
__global__ void dense_mask_kernel(int l, const c128* __restrict__ A, const int32_t* __restrict__ gid, const int32_t* __restrict__ gslot,
                                  int zero_slot, const int32_t* __restrict__ bslot, const int32_t* __restrict__ bk, c128* __restrict__ out) {
    const long long ll = (long long)l * l;
    const int j = blockIdx.y;
    const int want = bslot[j];
    const c128* Ak = A + (long long)bk[j] * ll;
    c128* o = out + (long long)j * ll;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < ll; e += (long long)gridDim.x * blockDim.x)
        o[e] = entry_slot(gid[e], gslot, zero_slot) == want ? Ak[e] : make_double2(0.0, 0.0);
}

// "debug_checks": every index the term builder produced is range-checked on the device before it is used (the pool this
// library is developed on does not allow compute-sanitizer, so the builder carries its own bounds checks)
__global__ void validate_lists_kernel(int l, long long ng, long long ncross, const int32_t* __restrict__ cidx, const int32_t* __restrict__ cg,
                                      long long nlamb, const int32_t* __restrict__ lidx, const int32_t* __restrict__ lg,
                                      const int32_t* __restrict__ gid, const int32_t* __restrict__ gslot, int ndense, unsigned long long* __restrict__ bad) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long b = 0;
    if (e < ncross) {
        for (int k = 0; k < 4; ++k) b += (cidx[4 * e + k] < 0 || cidx[4 * e + k] >= l) ? 1 : 0;
        const int g = cg[e];
        b += ((g > 0 ? g : -g) > ng) ? 1 : 0;
        if (!b) {  // the two entries of a cross term belong to the group the term claims
            b += gid[cidx[4 * e] + (long long)cidx[4 * e + 1] * l] != g ? 1 : 0;
            b += gid[cidx[4 * e + 2] + (long long)cidx[4 * e + 3] * l] != g ? 1 : 0;
        }
    }
    if (e < nlamb) {
        for (int k = 0; k < 3; ++k) b += (lidx[3 * e + k] < 0 || lidx[3 * e + k] >= l) ? 1 : 0;
        b += ((lg[e] > 0 ? lg[e] : -lg[e]) > ng) ? 1 : 0;
    }
    if (e < (long long)l * l) {
        const int g = gid[e];
        const long long ag = g > 0 ? g : -g;
        b += ag > ng ? 1 : 0;
        if (ag >= 1 && ag <= ng) b += (gslot[ag - 1] < -1 || gslot[ag - 1] >= ndense) ? 1 : 0;
        const int a = (int)(e % l), c = (int)(e / l);
        b += (a == c && g != 0) ? 1 : 0;                                   // the diagonal is in the zero group
        b += (a != c && gid[c + (long long)a * l] != -g) ? 1 : 0;          // (j,i) carries −key of (i,j)
    }
    if (b) atomicAdd(bad, b);
}

inline unsigned cdivu(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// grow a device array preserving its content (stream-ordered)
template <class T>
int grow(oqb_ctx* c, T*& p, size_t old_n, size_t new_n) {
    T* q = nullptr;
    OQB_TRY(ame_alloc(c, (void**)&q, new_n * sizeof(T)));
    if (p && old_n) OQB_CUDA(c, cudaMemcpyAsync(q, p, old_n * sizeof(T), cudaMemcpyDeviceToDevice, c->stream));
    ame_free(c, p);
    p = q;
    return 0;
}

// Which gap groups take the dense path, then the shared cross-term / Lamb-shift lists of the others.
int build_lists(oqb_ame* a) {
    oqb_ctx* c = a->ctx;
    const int l = a->lvl;
    const long long ng = a->ng, nz = a->nzp;
    const long long m_all = (long long)l * (l - 1) / 2;
    // ---- per-group decision -----------------------------------------------------------------------------------------
    // list cost ≈ 2m² entries × (K complex MACs + 2 atomics) per member; dense cost = 2 signs × 2 products of 8l³ flops on the
    // tensor pipe.  Break-even m² ≈ 4.6e-4·l³ (35 TFLOP/s against ≈1 ns per list entry and coupling).
    long long thr = c->opt.davies_dense_min_pairs;
    if (thr <= 0) {
        thr = (long long)std::sqrt(4.6e-4 * (double)l * l * l);
        if (thr < 24) thr = 24;
    }
    std::vector<long long> sizes(ng);
    for (long long g = 0; g < ng; ++g) sizes[g] = a->h_gptr[g + 1] - a->h_gptr[g];
    const long long mz = 2 * nz + l;
    auto zero_entries = [&]() { return nz > 0 ? 2 * nz * (mz - 1) + 2 * nz * l : 0ll; };
    // the cap on the list: lower the threshold until the entries fit (largest groups go dense first)
    for (;;) {
        long long tot = (nz > 0 && 2 * nz < thr) ? zero_entries() : 0;
        for (long long g = 0; g < ng; ++g)
            if (sizes[g] >= 2 && sizes[g] < thr) tot += 2 * sizes[g] * (sizes[g] - 1);
        if (tot <= c->opt.davies_max_cross || thr <= 2) break;
        thr = thr * 3 / 4 > 2 ? thr * 3 / 4 : 2;
    }
    a->h_dense_groups.clear();
    std::vector<int32_t> gslot(ng > 0 ? ng : 1, -1);
    for (long long g = 0; g < ng; ++g)
        if (sizes[g] >= 2 && sizes[g] >= thr) {
            gslot[g] = (int32_t)a->h_dense_groups.size();
            a->h_dense_groups.push_back(g);
        }
    a->zero_dense = nz > 0 && 2 * nz >= thr;
    OQB_TRY(ame_alloc(c, (void**)&a->d_gslot, gslot.size() * sizeof(int32_t)));
    OQB_CUDA(c, cudaMemcpyAsync(a->d_gslot, gslot.data(), gslot.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    // ---- cross-term list --------------------------------------------------------------------------------------------
    std::vector<long long> start, off;
    std::vector<int32_t> msz, gnum;
    long long nent = 0;
    for (long long g = 0; g < ng; ++g)
        if (sizes[g] >= 2 && gslot[g] < 0) {
            start.push_back(a->h_gptr[g]);
            msz.push_back((int32_t)sizes[g]);
            gnum.push_back((int32_t)(g + 1));
            off.push_back(nent);
            nent += 2 * sizes[g] * (sizes[g] - 1);
        }
    const long long nzent = a->zero_dense ? 0 : zero_entries();
    const long long ncross = nent + nzent;
    a->ncross = ncross;
    a->nlamb = 0;
    if (ncross > 0) {
        OQB_TRY(ame_alloc(c, (void**)&a->d_cross_idx, (size_t)ncross * 4 * sizeof(int32_t)));
        OQB_TRY(ame_alloc(c, (void**)&a->d_cross_g, (size_t)ncross * sizeof(int32_t)));
        if (nent > 0) {
            const size_t nl = start.size();
            const size_t bytes = nl * (8 + 8 + 4 + 4);
            OQB_TRY(ws_reserve(c, bytes + 64));
            char* base = (char*)c->ws;
            long long* d_start = (long long*)base;
            long long* d_off = d_start + nl;
            int32_t* d_m = (int32_t*)(d_off + nl);
            int32_t* d_g = d_m + nl;
            OQB_CUDA(c, cudaMemcpyAsync(d_start, start.data(), nl * 8, cudaMemcpyHostToDevice, c->stream));
            OQB_CUDA(c, cudaMemcpyAsync(d_off, off.data(), nl * 8, cudaMemcpyHostToDevice, c->stream));
            OQB_CUDA(c, cudaMemcpyAsync(d_m, msz.data(), nl * 4, cudaMemcpyHostToDevice, c->stream));
            OQB_CUDA(c, cudaMemcpyAsync(d_g, gnum.data(), nl * 4, cudaMemcpyHostToDevice, c->stream));
            cross_gen_kernel<<<cdivu(nent, 256), 256, 0, c->stream>>>(l, nent, (int)nl, d_start, d_m, d_off, d_g, a->d_spair, a->d_cross_idx,
                                                                       a->d_cross_g);
            OQB_LAUNCH_CHECK(c);
            OQB_CUDA(c, cudaStreamSynchronize(c->stream));  // the host vectors go out of scope / the workspace is reused below
        }
        if (nzent > 0) {
            cross_zero_kernel<<<cdivu(nzent, 256), 256, 0, c->stream>>>(l, nz, a->d_spair + (m_all - nz), nzent, a->d_cross_idx + 4 * nent,
                                                                         a->d_cross_g + nent);
            OQB_LAUNCH_CHECK(c);
        }
        // entries with a == a2 also feed L'L: compact their numbers (order preserved)
        size_t tmp = 0;
        thrust::counting_iterator<long long> cnt_it(0);
        cub::DeviceSelect::Flagged(nullptr, tmp, cnt_it, (const uint8_t*)nullptr, (long long*)nullptr, (unsigned long long*)nullptr, ncross,
                                   c->stream);
        const size_t bytes = (size_t)ncross * (1 + 8) + tmp + 1024;
        OQB_TRY(ws_reserve(c, bytes));
        char* base = (char*)c->ws;
        long long* d_src = (long long*)base;
        unsigned long long* d_n = (unsigned long long*)(d_src + ncross);
        void* d_tmp = (void*)(((uintptr_t)(d_n + 2) + 255) & ~(uintptr_t)255);
        uint8_t* d_flag = (uint8_t*)d_tmp + ((tmp + 255) & ~(size_t)255);
        lamb_flag_kernel<<<cdivu(ncross, 256), 256, 0, c->stream>>>(ncross, a->d_cross_idx, d_flag);
        OQB_LAUNCH_CHECK(c);
        OQB_CUDA(c, cub::DeviceSelect::Flagged(d_tmp, tmp, cnt_it, d_flag, d_src, d_n, ncross, c->stream));
        unsigned long long h_n = 0;
        OQB_CUDA(c, cudaMemcpyAsync(&h_n, d_n, sizeof(h_n), cudaMemcpyDeviceToHost, c->stream));
        OQB_CUDA(c, cudaStreamSynchronize(c->stream));
        if (h_n > 0) {
            OQB_TRY(ame_alloc(c, (void**)&a->d_lamb_idx, (size_t)h_n * 3 * sizeof(int32_t)));
            OQB_TRY(ame_alloc(c, (void**)&a->d_lamb_g, (size_t)h_n * sizeof(int32_t)));
            lamb_fill_kernel<<<cdivu((long long)h_n, 256), 256, 0, c->stream>>>((long long)h_n, d_src, a->d_cross_idx, a->d_cross_g, a->d_lamb_idx,
                                                                                 a->d_lamb_g);
            OQB_LAUNCH_CHECK(c);
            OQB_CUDA(c, cudaStreamSynchronize(c->stream));  // d_src lives in the workspace
        }
        a->nlamb = (long long)h_n;
    }
    if (c->opt.debug_checks) {
        unsigned long long* d_bad = nullptr;
        OQB_TRY(ame_alloc(c, (void**)&d_bad, sizeof(unsigned long long)));
        OQB_CUDA(c, cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), c->stream));
        const long long work = std::max<long long>(std::max<long long>(a->ncross, a->nlamb), (long long)l * l);
        validate_lists_kernel<<<cdivu(work, 256), 256, 0, c->stream>>>(l, ng, a->ncross, a->d_cross_idx, a->d_cross_g, a->nlamb, a->d_lamb_idx,
                                                                        a->d_lamb_g, a->d_gid, a->d_gslot, (int)a->h_dense_groups.size(), d_bad);
        c->launches++;
        unsigned long long h_bad = 1;
        OQB_CUDA(c, cudaMemcpyAsync(&h_bad, d_bad, sizeof(h_bad), cudaMemcpyDeviceToHost, c->stream));
        OQB_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFreeAsync(d_bad, c->stream);
        if (h_bad) return set_error(c, OQB_ECUDA, "davies term builder: %llu index/group consistency violations (debug_checks)", h_bad);
    }
    a->lists_built = true;
    return 0;
}

// Dense per-group blocks of one term: left operators A_left[k]∘[ω̃ == x], right operators A_right[k]∘[…] (nullptr: the same —
// a DaviesGenerator; a correlated pair has A_β on the left and A_α on the right).  Appends to the handle's stacks and adds
// (−½γ ∓ iS)·L_d'L to the off-diagonal matrices.
int add_dense_blocks(oqb_ame* a, const c128* A, const c128* Ar, int nk, const double* d_vals) {
    oqb_ctx* c = a->ctx;
    const int l = a->lvl;
    const long long ll = (long long)l * l, ng = a->ng;
    const int ndg = (int)a->h_dense_groups.size();
    const int nslots = 2 * ndg + (a->zero_dense ? 1 : 0);
    if (nslots == 0) return 0;
    const bool pair = Ar != nullptr;
    const int zero_slot = a->zero_dense ? 2 * ndg : -1;
    // squared norms of every candidate block and of the couplings themselves; γ, S of the slots
    const size_t nnorm = (size_t)nslots * nk + nk;
    double* d_norm = nullptr;
    OQB_TRY(ame_alloc(c, (void**)&d_norm, 2 * nnorm * sizeof(double)));
    OQB_CUDA(c, cudaMemsetAsync(d_norm, 0, 2 * nnorm * sizeof(double), c->stream));
    dim3 gridn(std::min<unsigned>(cdivu(ll, 256), 1184u), (unsigned)nk);
    dense_norm_kernel<<<gridn, 256, 0, c->stream>>>(l, nk, A, a->d_gid, a->d_gslot, zero_slot, d_norm, d_norm + (size_t)nslots * nk);
    if (pair) dense_norm_kernel<<<gridn, 256, 0, c->stream>>>(l, nk, Ar, a->d_gid, a->d_gslot, zero_slot, d_norm + nnorm, d_norm + nnorm + (size_t)nslots * nk);
    c->launches += pair ? 2 : 1;
    std::vector<double> h_norm(2 * nnorm);
    cudaError_t e = cudaMemcpyAsync(h_norm.data(), d_norm, h_norm.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    std::vector<double> h_g(nslots), h_S(nslots);
    auto fetch = [&](double* dst, long long index) {
        if (e == cudaSuccess) e = cudaMemcpyAsync(dst, d_vals + index, sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    };
    for (int d = 0; d < ndg; ++d) {
        const long long g = a->h_dense_groups[d];
        fetch(&h_g[2 * d], g); fetch(&h_g[2 * d + 1], ng + g);
        fetch(&h_S[2 * d], 2 * ng + g); fetch(&h_S[2 * d + 1], 3 * ng + g);
    }
    if (a->zero_dense) { fetch(&h_g[2 * ndg], 4 * ng); fetch(&h_S[2 * ndg], 4 * ng + 1); }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFreeAsync(d_norm, c->stream);
    if (e != cudaSuccess) return set_error(c, OQB_ECUDA, "oqb_ame_add_davies: %s", cudaGetErrorString(e));
    // keep the blocks that matter: ‖L‖² ≤ 1e-26·‖A_k‖² contributes ≤ 1e-26 relative (eigenvectors of a degenerate level
    // leave "forbidden" blocks at rounding level, ≈1e-32)
    std::vector<int32_t> bslot, bk;
    std::vector<double> brate;
    std::vector<c128> bcoef;
    for (int sl = 0; sl < nslots; ++sl)
        for (int k = 0; k < nk; ++k) {
            if (!(h_norm[(size_t)sl * nk + k] > 1e-26 * h_norm[(size_t)nslots * nk + k])) continue;
            if (pair && !(h_norm[nnorm + (size_t)sl * nk + k] > 1e-26 * h_norm[nnorm + (size_t)nslots * nk + k])) continue;
            if (h_g[sl] == 0.0 && h_S[sl] == 0.0) continue;
            bslot.push_back(sl);
            bk.push_back(k);
            brate.push_back(h_g[sl]);
            bcoef.push_back(make_double2(-0.5 * h_g[sl], -h_S[sl]));  // −½γ − iS
        }
    const int nb_new = (int)bslot.size();
    if (nb_new == 0) return 0;
    const size_t blk_bytes = (size_t)ll * sizeof(c128);
    const bool need_right = pair || a->d_Rstack != nullptr;
    if (((size_t)a->nd + nb_new) * blk_bytes * (need_right ? 2 : 1) > ((size_t)24 << 30))
        return set_error(c, OQB_ENOMEM, "oqb_ame_add_davies: %d dense gap-group blocks of %d×%d exceed the 24 GiB budget", a->nd + nb_new, l, l);
    if (pair && !a->d_Rstack && a->nd > 0) {  // earlier DaviesGenerator blocks: right = left
        OQB_TRY(ame_alloc(c, (void**)&a->d_Rstack, (size_t)a->nd * blk_bytes));
        OQB_CUDA(c, cudaMemcpyAsync(a->d_Rstack, a->d_Lstack, (size_t)a->nd * blk_bytes, cudaMemcpyDeviceToDevice, c->stream));
    }
    OQB_TRY(grow(c, a->d_Lstack, (size_t)a->nd * ll, ((size_t)a->nd + nb_new) * ll));
    if (need_right) OQB_TRY(grow(c, a->d_Rstack, (size_t)a->nd * ll, ((size_t)a->nd + nb_new) * ll));
    OQB_TRY(grow(c, a->d_drate, (size_t)a->nd, (size_t)a->nd + nb_new));
    int32_t* d_b = nullptr;
    OQB_TRY(ame_alloc(c, (void**)&d_b, (size_t)2 * nb_new * sizeof(int32_t)));
    OQB_CUDA(c, cudaMemcpyAsync(d_b, bslot.data(), nb_new * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    OQB_CUDA(c, cudaMemcpyAsync(d_b + nb_new, bk.data(), nb_new * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    OQB_CUDA(c, cudaMemcpyAsync(a->d_drate + a->nd, brate.data(), nb_new * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    c128* Lnew = a->d_Lstack + (long long)a->nd * ll;
    c128* Rnew = need_right ? a->d_Rstack + (long long)a->nd * ll : Lnew;
    for (int j0 = 0; j0 < nb_new; j0 += 65535) {
        const int cnt = nb_new - j0 < 65535 ? nb_new - j0 : 65535;
        dim3 gridm(std::min<unsigned>(cdivu(ll, 256), 1184u), (unsigned)cnt);
        dense_mask_kernel<<<gridm, 256, 0, c->stream>>>(l, A, a->d_gid, a->d_gslot, zero_slot, d_b + j0, d_b + nb_new + j0, Lnew + (long long)j0 * ll);
        OQB_LAUNCH_CHECK(c);
        if (need_right) {
            dense_mask_kernel<<<gridm, 256, 0, c->stream>>>(l, pair ? Ar : A, a->d_gid, a->d_gslot, zero_slot, d_b + j0, d_b + nb_new + j0,
                                                            Rnew + (long long)j0 * ll);
            OQB_LAUNCH_CHECK(c);
        }
    }
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));  // host vectors were the copy sources
    cudaFreeAsync(d_b, c->stream);
    // M1 += (−½γ − iS) L_d'L per block (and M2 += (−½γ + iS) L_d'L when the two differ)
    for (int j = 0; j < nb_new; ++j) {
        ZgemmParams p;
        p.M = p.N = p.K = l;
        p.opA = OP_C;
        p.A = Rnew + (long long)j * ll; p.lda = l;
        p.B = Lnew + (long long)j * ll; p.ldb = l;
        p.C = a->d_Moff; p.ldc = l;
        p.alpha = bcoef[j];
        p.beta = make_double2(1.0, 0.0);
        OQB_TRY(zgemm(c, p));
        if (a->d_Moff2) {
            p.C = a->d_Moff2;
            p.alpha = make_double2(bcoef[j].x, -bcoef[j].y);
            OQB_TRY(zgemm(c, p));
        }
    }
    a->nd += nb_new;
    OQB_TRY(ame_detect_real(a, a->d_Lstack, (long long)a->nd * ll, &a->dense_real));
    if (a->d_Rstack && a->dense_real) OQB_TRY(ame_detect_real(a, a->d_Rstack, (long long)a->nd * ll, &a->dense_real));
    return 0;
}

// Adds one DaviesGenerator whose per-key values are in d_vals (device, layout of term_vals_ohmic_kernel).
int add_davies_impl(oqb_ame* a, int k0, int nk, const double* d_vals) {
    oqb_ctx* c = a->ctx;
    const int l = a->lvl;
    const long long ll = (long long)l * l, ng = a->ng;
    if (!a->lists_built) OQB_TRY(build_lists(a));
    const c128* A = a->d_A + (long long)k0 * ll;
    // first generator: totals start from the Hamiltonian-only coefficients that set_eigen left in d_Cm
    if (!a->d_Wt) {
        OQB_TRY(ame_alloc(c, (void**)&a->d_Gam, l * sizeof(double)));
        OQB_TRY(ame_alloc(c, (void**)&a->d_hls, l * sizeof(double)));
        OQB_TRY(ame_alloc(c, (void**)&a->d_Wt, (size_t)ll * sizeof(double)));
        OQB_CUDA(c, cudaMemsetAsync(a->d_Gam, 0, l * sizeof(double), c->stream));
        OQB_CUDA(c, cudaMemsetAsync(a->d_hls, 0, l * sizeof(double), c->stream));
        OQB_CUDA(c, cudaMemsetAsync(a->d_Wt, 0, (size_t)ll * sizeof(double), c->stream));
    }
    // γ/S tables of this generator (scratch), its column sums, and the totals
    double* d_tab = nullptr;
    OQB_TRY(ame_alloc(c, (void**)&d_tab, ((size_t)2 * ll + 2 * l) * sizeof(double)));
    double* d_gam = d_tab;
    double* d_S = d_tab + ll;
    double* d_Gt = d_S + ll;
    double* d_ht = d_Gt + l;
    int st = 0;
    do {
        term_table_kernel<<<cdivu(ll, 256), 256, 0, c->stream>>>(l, a->d_gid, a->d_gslot, a->zero_dense ? 1 : 0, ng, d_vals, d_gam, d_S);
        term_colsum_kernel<<<l, 256, 0, c->stream>>>(l, nk, A, d_gam, d_S, d_Gt, d_ht, a->d_Wt);
        term_cmat_acc_kernel<<<cdivu(ll, 256), 256, 0, c->stream>>>(l, nk, A, d_Gt, d_ht, d_vals + 4 * ng, a->zero_dense ? 1 : 0, a->d_Cm);
        vec_add_kernel<<<cdivu(l, 256), 256, 0, c->stream>>>(l, d_Gt, a->d_Gam);
        vec_add_kernel<<<cdivu(l, 256), 256, 0, c->stream>>>(l, d_ht, a->d_hls);
        c->launches += 5;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { st = set_error(c, OQB_ECUDA, "oqb_ame_add_davies: kernel launch failed: %s", cudaGetErrorString(e)); break; }
    } while (0);
    cudaFreeAsync(d_tab, c->stream);
    if (st) return st;
    // cross-term rates of this generator
    DaviesTerm t;
    t.k0 = k0; t.nk = nk;
    if (a->ncross > 0) {
        OQB_TRY(ame_alloc(c, (void**)&t.d_rate, (size_t)a->ncross * sizeof(double)));
        term_rate_kernel<<<cdivu(a->ncross, 256), 256, 0, c->stream>>>(a->ncross, a->d_cross_g, ng, d_vals, t.d_rate, nullptr);
        OQB_LAUNCH_CHECK(c);
    }
    a->terms.push_back(t);
    // off-diagonal part of −½G − iH_LS from the list entries that share a row
    const bool need_moff = a->nlamb > 0 || !a->h_dense_groups.empty() || a->zero_dense;
    if (need_moff && !a->d_Moff) {
        OQB_TRY(ame_alloc(c, (void**)&a->d_Moff, (size_t)ll * sizeof(c128)));
        OQB_CUDA(c, cudaMemsetAsync(a->d_Moff, 0, (size_t)ll * sizeof(c128), c->stream));
    }
    if (a->nlamb > 0) {
        double* d_rs = nullptr;
        OQB_TRY(ame_alloc(c, (void**)&d_rs, (size_t)a->nlamb * 2 * sizeof(double)));
        term_rate_kernel<<<cdivu(a->nlamb, 256), 256, 0, c->stream>>>(a->nlamb, a->d_lamb_g, ng, d_vals, nullptr, d_rs);
        c->launches++;
        st = davies_lamb_build(c, l, nk, A, a->nlamb, a->d_lamb_idx, d_rs, a->d_Moff);
        if (!st && a->d_Moff2) {  // a correlated term is present: the right-hand matrix is kept separately (= M1' for this term)
            for (int k = 0; k < nk && !st; ++k) {
                corr_lamb_kernel<<<cdivu(a->nlamb, 256), 256, 0, c->stream>>>(l, A + (long long)k * ll, A + (long long)k * ll, a->nlamb, a->d_lamb_idx, d_rs,
                                                                             a->d_Mscratch, a->d_Moff2);
                c->launches++;
            }
        }
        cudaFreeAsync(d_rs, c->stream);
        if (st) return st;
    }
    OQB_TRY(add_dense_blocks(a, A, nullptr, nk, d_vals));
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    a->davies = true;
    a->cross_fuse_state = -1;  // the Hadamard coefficients changed
    return 0;
}

// Adds one pair (α,β) of a CorrelatedDaviesGenerator; d_vals as for add_davies_impl (γ_αβ, S_αβ at the keys).
int add_correlated_impl(oqb_ame* a, int ka, int kb, const double* d_vals) {
    oqb_ctx* c = a->ctx;
    const int l = a->lvl;
    const long long ll = (long long)l * l, ng = a->ng;
    if (!a->lists_built) OQB_TRY(build_lists(a));
    const c128* Aa = a->d_A + (long long)ka * ll;
    const c128* Ab = a->d_A + (long long)kb * ll;
    if (!a->d_Wt) {
        OQB_TRY(ame_alloc(c, (void**)&a->d_Gam, l * sizeof(double)));
        OQB_TRY(ame_alloc(c, (void**)&a->d_hls, l * sizeof(double)));
        OQB_TRY(ame_alloc(c, (void**)&a->d_Wt, (size_t)ll * sizeof(double)));
        OQB_CUDA(c, cudaMemsetAsync(a->d_Gam, 0, l * sizeof(double), c->stream));
        OQB_CUDA(c, cudaMemsetAsync(a->d_hls, 0, l * sizeof(double), c->stream));
        OQB_CUDA(c, cudaMemsetAsync(a->d_Wt, 0, (size_t)ll * sizeof(double), c->stream));
    }
    if (!a->d_Wt_im) {
        OQB_TRY(ame_alloc(c, (void**)&a->d_Wt_im, (size_t)ll * sizeof(double)));
        OQB_CUDA(c, cudaMemsetAsync(a->d_Wt_im, 0, (size_t)ll * sizeof(double), c->stream));
    }
    // M1 (left of ρ) and M2 (right of ρ) are kept separately from now on; M2 starts as M1' of what is there already
    if (!a->d_Moff) {
        OQB_TRY(ame_alloc(c, (void**)&a->d_Moff, (size_t)ll * sizeof(c128)));
        OQB_CUDA(c, cudaMemsetAsync(a->d_Moff, 0, (size_t)ll * sizeof(c128), c->stream));
    }
    if (!a->d_Moff2) {
        OQB_TRY(ame_alloc(c, (void**)&a->d_Moff2, (size_t)ll * sizeof(c128)));
        OQB_TRY(ame_alloc(c, (void**)&a->d_Mscratch, (size_t)ll * sizeof(c128)));
        adjoint_kernel<<<cdivu(ll, 256), 256, 0, c->stream>>>(l, a->d_Moff, a->d_Moff2);
        OQB_LAUNCH_CHECK(c);
    }
    a->has_corr = true;
    double* d_tab = nullptr;
    OQB_TRY(ame_alloc(c, (void**)&d_tab, ((size_t)2 * ll + 4 * l) * sizeof(double)));
    double* d_gam = d_tab;
    double* d_S = d_tab + ll;
    c128* d_Gt = (c128*)(d_S + ll);
    c128* d_ht = d_Gt + l;
    term_table_kernel<<<cdivu(ll, 256), 256, 0, c->stream>>>(l, a->d_gid, a->d_gslot, a->zero_dense ? 1 : 0, ng, d_vals, d_gam, d_S);
    corr_colsum_kernel<<<l, 256, 0, c->stream>>>(l, Aa, Ab, d_gam, d_S, d_Gt, d_ht, a->d_Wt, a->d_Wt_im);
    corr_cmat_acc_kernel<<<cdivu(ll, 256), 256, 0, c->stream>>>(l, Aa, Ab, d_Gt, d_ht, d_vals + 4 * ng, a->zero_dense ? 1 : 0, a->d_Cm);
    c->launches += 3;
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(d_tab, c->stream);
    if (e != cudaSuccess) return set_error(c, OQB_ECUDA, "oqb_ame_add_correlated: kernel launch failed: %s", cudaGetErrorString(e));
    CorrTerm t;
    t.ka = ka; t.kb = kb;
    if (a->ncross > 0) {
        OQB_TRY(ame_alloc(c, (void**)&t.d_rate, (size_t)a->ncross * sizeof(double)));
        term_rate_kernel<<<cdivu(a->ncross, 256), 256, 0, c->stream>>>(a->ncross, a->d_cross_g, ng, d_vals, t.d_rate, nullptr);
        OQB_LAUNCH_CHECK(c);
    }
    a->corr_terms.push_back(t);
    if (a->nlamb > 0) {
        double* d_rs = nullptr;
        OQB_TRY(ame_alloc(c, (void**)&d_rs, (size_t)a->nlamb * 2 * sizeof(double)));
        term_rate_kernel<<<cdivu(a->nlamb, 256), 256, 0, c->stream>>>(a->nlamb, a->d_lamb_g, ng, d_vals, nullptr, d_rs);
        corr_lamb_kernel<<<cdivu(a->nlamb, 256), 256, 0, c->stream>>>(l, Aa, Ab, a->nlamb, a->d_lamb_idx, d_rs, a->d_Moff, a->d_Moff2);
        c->launches += 2;
        cudaFreeAsync(d_rs, c->stream);
    }
    OQB_TRY(add_dense_blocks(a, Ab, Aa, 1, d_vals));
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    a->davies = true;
    return 0;
}

}  // namespace

extern "C" {

int ame_clear_gaps(oqb_ame* a) {
    oqb_ctx* c = a->ctx;
    ame_free(c, a->d_gid);
    ame_free(c, a->d_ukeys);
    ame_free(c, a->d_spair);
    a->h_gptr.clear();
    a->h_ukeys.clear();
    a->have_gaps = false;
    a->ng = a->nzp = 0;
    return 0;
}

int ame_clear_terms(oqb_ame* a) {
    oqb_ctx* c = a->ctx;
    for (DaviesTerm& t : a->terms) ame_free(c, t.d_rate);
    a->terms.clear();
    ame_free(c, a->d_cross_idx);
    ame_free(c, a->d_cross_g);
    ame_free(c, a->d_lamb_idx);
    ame_free(c, a->d_lamb_g);
    ame_free(c, a->d_gslot);
    ame_free(c, a->d_Lstack);
    ame_free(c, a->d_Rstack);
    ame_free(c, a->d_drate);
    for (CorrTerm& t : a->corr_terms) ame_free(c, t.d_rate);
    a->corr_terms.clear();
    ame_free(c, a->d_Moff2);
    ame_free(c, a->d_Mscratch);
    a->has_corr = false;
    a->h_dense_groups.clear();
    a->zero_dense = false;
    a->lists_built = false;
    a->cross_fuse_state = -1;
    a->dense_real = false;
    a->nd = 0;
    a->ncross = a->nlamb = 0;
    return 0;
}

int oqb_ame_gaps_build(oqb_ame* a, int digits, int sigdigits, int64_t* n_keys, int64_t* n_zero_pairs) {
    if (!a) return OQB_EINVAL;
    OQB_DEVICE(a, a->ctx);
    oqb_ctx* c = a->ctx;
    if (!a->have_eigen) return set_error(c, OQB_EINVAL, "oqb_ame_gaps_build: call oqb_ame_set_eigen first");
    OQB_TRY(oqb_ame_clear_davies(a));  // tables of another grouping are stale
    OQB_TRY(ame_clear_gaps(a));
    const int l = a->lvl;
    const long long ll = (long long)l * l, m = (long long)l * (l - 1) / 2;
    a->gap_digits = digits;
    a->gap_sigdigits = sigdigits;
    OQB_TRY(ame_alloc(c, (void**)&a->d_gid, (size_t)ll * sizeof(int32_t)));
    OQB_CUDA(c, cudaMemsetAsync(a->d_gid, 0, (size_t)ll * sizeof(int32_t), c->stream));
    a->h_gptr.assign(1, 0);
    if (m > 0) {
        size_t tmp_sort = 0, tmp_scan = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, (const double*)nullptr, (double*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr,
                                        (int)m, 0, 64, c->stream);
        cub::DeviceScan::InclusiveSum(nullptr, tmp_scan, (const int32_t*)nullptr, (int32_t*)nullptr, (int)m, c->stream);
        const size_t tmp = std::max(tmp_sort, tmp_scan);
        const size_t bytes = (size_t)m * (8 + 8 + 4 + 4 + 4 + 8) + 1024 + tmp;
        OQB_TRY(ws_reserve(c, bytes));
        OQB_TRY(ame_alloc(c, (void**)&a->d_spair, (size_t)m * sizeof(int32_t)));
        char* base = (char*)c->ws;
        double* k_in = (double*)base;
        double* k_out = k_in + m;
        long long* gptr = (long long*)(k_out + m);
        int32_t* p_in = (int32_t*)(gptr + m);
        int32_t* head = p_in + m;
        int32_t* grp = head + m;
        unsigned long long* cnt = (unsigned long long*)(((uintptr_t)(grp + m) + 255) & ~(uintptr_t)255);
        void* cub_tmp = (void*)(((uintptr_t)(cnt + 4) + 255) & ~(uintptr_t)255);
        OQB_CUDA(c, cudaMemsetAsync(cnt, 0, 4 * sizeof(unsigned long long), c->stream));
        static const Pow10 p10 = make_pow10();
        gap_keys_kernel<<<cdivu(ll, 256), 256, 0, c->stream>>>(l, a->d_w, pow(10.0, -(double)digits), sigdigits, p10, k_in, p_in, cnt);
        OQB_LAUNCH_CHECK(c);
        size_t t1 = tmp;
        OQB_CUDA(c, cub::DeviceRadixSort::SortPairs(cub_tmp, t1, k_in, k_out, p_in, a->d_spair, (int)m, 0, 64, c->stream));
        head_flag_kernel<<<cdivu(m, 256), 256, 0, c->stream>>>(m, k_out, head);
        OQB_LAUNCH_CHECK(c);
        size_t t2 = tmp;
        OQB_CUDA(c, cub::DeviceScan::InclusiveSum(cub_tmp, t2, head, grp, (int)m, c->stream));
        int32_t h_ng = 0;
        unsigned long long h_nz = 0;
        OQB_CUDA(c, cudaMemcpyAsync(&h_ng, grp + (m - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        OQB_CUDA(c, cudaMemcpyAsync(&h_nz, cnt, sizeof(h_nz), cudaMemcpyDeviceToHost, c->stream));
        OQB_CUDA(c, cudaStreamSynchronize(c->stream));
        a->ng = h_ng;
        a->nzp = (long long)h_nz;
        OQB_TRY(ame_alloc(c, (void**)&a->d_ukeys, (size_t)(a->ng ? a->ng : 1) * sizeof(double)));
        scatter_gid_kernel<<<cdivu(m, 256), 256, 0, c->stream>>>(l, m, k_out, a->d_spair, head, grp, a->d_gid, a->d_ukeys, gptr);
        OQB_LAUNCH_CHECK(c);
        a->h_gptr.assign((size_t)a->ng + 1, 0);
        a->h_ukeys.assign((size_t)a->ng, 0.0);
        if (a->ng > 0) {
            OQB_CUDA(c, cudaMemcpyAsync(a->h_gptr.data(), gptr, (size_t)a->ng * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
            OQB_CUDA(c, cudaMemcpyAsync(a->h_ukeys.data(), a->d_ukeys, (size_t)a->ng * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        }
        OQB_CUDA(c, cudaStreamSynchronize(c->stream));
        a->h_gptr[a->ng] = m - a->nzp;
    }
    a->have_gaps = true;
    if (n_keys) *n_keys = a->ng;
    if (n_zero_pairs) *n_zero_pairs = a->nzp;
    return 0;
}

int oqb_ame_gaps_keys(oqb_ame* a, double* h_keys) {
    if (!a) return OQB_EINVAL;
    oqb_ctx* c = a->ctx;
    if (!a->have_gaps) return set_error(c, OQB_EINVAL, "oqb_ame_gaps_keys: call oqb_ame_gaps_build first");
    if (a->ng > 0 && !h_keys) return set_error(c, OQB_EINVAL, "oqb_ame_gaps_keys: null output");
    for (long long g = 0; g < a->ng; ++g) h_keys[g] = a->h_ukeys[g];
    return 0;
}

int oqb_ame_gaps_fetch(oqb_ame* a, int64_t* h_group_ptr, int32_t* h_a, int32_t* h_b, int32_t* h_a0, int32_t* h_b0) {
    if (!a) return OQB_EINVAL;
    OQB_DEVICE(a, a->ctx);
    oqb_ctx* c = a->ctx;
    if (!a->have_gaps) return set_error(c, OQB_EINVAL, "oqb_ame_gaps_fetch: call oqb_ame_gaps_build first");
    const int l = a->lvl;
    const long long m = (long long)l * (l - 1) / 2, np = m - a->nzp;
    if (h_group_ptr)
        for (long long g = 0; g <= a->ng; ++g) h_group_ptr[g] = a->h_gptr[g];
    std::vector<int32_t> sp((size_t)m);
    if (m > 0) {
        OQB_CUDA(c, cudaMemcpyAsync(sp.data(), a->d_spair, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    auto ij = [l](long long p, int32_t& i, int32_t& j) {
        long long ii = 0;
        const double b = 2.0 * l - 1.0;
        ii = (long long)std::floor((b - std::sqrt(b * b - 8.0 * (double)p)) * 0.5);
        if (ii < 0) ii = 0;
        if (ii > l - 2) ii = l - 2;
        while (ii > 0 && ii * l - ii * (ii + 1) / 2 > p) --ii;
        while ((ii + 1) * l - (ii + 1) * (ii + 2) / 2 <= p) ++ii;
        i = (int32_t)ii;
        j = (int32_t)(p - (ii * l - ii * (ii + 1) / 2)) + i + 1;
    };
    if (h_a && h_b)
        for (long long k = 0; k < np; ++k) {
            int32_t i, j;
            ij(sp[k], i, j);
            h_a[k] = i + 1;  // 1-based like GapIndices.uniq_a / uniq_b
            h_b[k] = j + 1;
        }
    if (h_a0 && h_b0) {
        // a0/b0 in the reference's order: (i, j) then (j, i) per degenerate pair in loop order, then the diagonal (:34-37, :58-59)
        std::vector<int32_t> z(sp.begin() + np, sp.end());
        std::sort(z.begin(), z.end());
        long long o = 0;
        for (int32_t p : z) {
            int32_t i, j;
            ij(p, i, j);
            h_a0[o] = i + 1; h_b0[o++] = j + 1;
            h_a0[o] = j + 1; h_b0[o++] = i + 1;
        }
        for (int k = 0; k < l; ++k) { h_a0[o] = k + 1; h_b0[o++] = k + 1; }
    }
    return 0;
}

int oqb_ame_add_davies(oqb_ame* a, int k0, int nk, const double* h_gamma_plus, const double* h_gamma_minus, const double* h_S_plus,
                       const double* h_S_minus, double gamma0, double S0) {
    if (!a) return OQB_EINVAL;
    OQB_DEVICE(a, a->ctx);
    oqb_ctx* c = a->ctx;
    if (!a->have_gaps) return set_error(c, OQB_EINVAL, "oqb_ame_add_davies: call oqb_ame_gaps_build first");
    if (k0 < 0 || nk <= 0 || k0 + nk > a->K) return set_error(c, OQB_EINVAL, "oqb_ame_add_davies: coupling slice [%d, %d) outside 0..%d", k0, k0 + nk, a->K);
    const long long ng = a->ng;
    if (ng > 0 && (!h_gamma_plus || !h_gamma_minus)) return set_error(c, OQB_EINVAL, "oqb_ame_add_davies: null rate array");
    std::vector<double> vals((size_t)4 * ng + 2, 0.0);
    for (long long g = 0; g < ng; ++g) {
        vals[g] = h_gamma_plus[g];
        vals[ng + g] = h_gamma_minus[g];
        vals[2 * ng + g] = h_S_plus ? h_S_plus[g] : 0.0;
        vals[3 * ng + g] = h_S_minus ? h_S_minus[g] : 0.0;
    }
    vals[4 * ng] = gamma0;
    vals[4 * ng + 1] = S0;
    double* d_vals = nullptr;
    OQB_TRY(ame_alloc(c, (void**)&d_vals, vals.size() * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(d_vals, vals.data(), vals.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    int st = e == cudaSuccess ? add_davies_impl(a, k0, nk, d_vals) : set_error(c, OQB_ECUDA, "oqb_ame_add_davies: %s", cudaGetErrorString(e));
    cudaStreamSynchronize(c->stream);
    cudaFreeAsync(d_vals, c->stream);
    return st;
}

int oqb_ame_add_davies_ohmic(oqb_ame* a, int k0, int nk, double eta, double wc, double beta, int nS, double S_x0, double S_dx,
                             const double* h_S_coef) {
    if (!a) return OQB_EINVAL;
    OQB_DEVICE(a, a->ctx);
    oqb_ctx* c = a->ctx;
    if (!a->have_gaps) return set_error(c, OQB_EINVAL, "oqb_ame_add_davies_ohmic: call oqb_ame_gaps_build first");
    if (k0 < 0 || nk <= 0 || k0 + nk > a->K) return set_error(c, OQB_EINVAL, "oqb_ame_add_davies_ohmic: coupling slice [%d, %d) outside 0..%d", k0, k0 + nk, a->K);
    if (nS < 0 || (nS > 0 && (!h_S_coef || !(S_dx > 0.0)))) return set_error(c, OQB_EINVAL, "oqb_ame_add_davies_ohmic: bad Lamb-shift spline");
    const long long ng = a->ng;
    double* d_vals = nullptr;
    OQB_TRY(ame_alloc(c, (void**)&d_vals, ((size_t)4 * ng + 2 + (nS ? nS + 2 : 0)) * sizeof(double)));
    double* d_coef = d_vals + 4 * ng + 2;
    cudaError_t e = cudaSuccess;
    if (nS) e = cudaMemcpyAsync(d_coef, h_S_coef, (size_t)(nS + 2) * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        term_vals_ohmic_kernel<<<cdivu(ng > 0 ? ng : 1, 256), 256, 0, c->stream>>>(ng, a->d_ukeys, eta, wc, beta, nS, S_x0, S_dx, d_coef, d_vals);
        c->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);  // h_S_coef may be reused by the caller
    int st = e == cudaSuccess ? add_davies_impl(a, k0, nk, d_vals) : set_error(c, OQB_ECUDA, "oqb_ame_add_davies_ohmic: %s", cudaGetErrorString(e));
    cudaStreamSynchronize(c->stream);
    cudaFreeAsync(d_vals, c->stream);
    return st;
}

int oqb_ame_add_correlated(oqb_ame* a, int alpha, int beta, const double* h_gamma_plus, const double* h_gamma_minus, const double* h_S_plus,
                           const double* h_S_minus, double gamma0, double S0) {
    if (!a) return OQB_EINVAL;
    OQB_DEVICE(a, a->ctx);
    oqb_ctx* c = a->ctx;
    if (!a->have_gaps) return set_error(c, OQB_EINVAL, "oqb_ame_add_correlated: call oqb_ame_gaps_build first");
    if (alpha < 0 || alpha >= a->K || beta < 0 || beta >= a->K) return set_error(c, OQB_EINVAL, "oqb_ame_add_correlated: coupling index out of range");
    const long long ng = a->ng;
    if (ng > 0 && (!h_gamma_plus || !h_gamma_minus)) return set_error(c, OQB_EINVAL, "oqb_ame_add_correlated: null rate array");
    std::vector<double> vals((size_t)4 * ng + 2, 0.0);
    for (long long g = 0; g < ng; ++g) {
        vals[g] = h_gamma_plus[g];
        vals[ng + g] = h_gamma_minus[g];
        vals[2 * ng + g] = h_S_plus ? h_S_plus[g] : 0.0;
        vals[3 * ng + g] = h_S_minus ? h_S_minus[g] : 0.0;
    }
    vals[4 * ng] = gamma0;
    vals[4 * ng + 1] = S0;
    double* d_vals = nullptr;
    OQB_TRY(ame_alloc(c, (void**)&d_vals, vals.size() * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(d_vals, vals.data(), vals.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    int st = e == cudaSuccess ? add_correlated_impl(a, alpha, beta, d_vals) : set_error(c, OQB_ECUDA, "oqb_ame_add_correlated: %s", cudaGetErrorString(e));
    cudaStreamSynchronize(c->stream);
    cudaFreeAsync(d_vals, c->stream);
    return st;
}

// cross terms of the correlated pairs (called by eig_apply in ame.cu)
int ame_corr_cross(oqb_ame* a, int nb, const c128* R, c128* X) {
    oqb_ctx* c = a->ctx;
    const int l = a->lvl;
    const long long ll = (long long)l * l;
    if (a->ncross <= 0) return 0;
    for (const CorrTerm& t : a->corr_terms)
        for (int b0 = 0; b0 < nb; b0 += 65535) {
            const int cnt = nb - b0 < 65535 ? nb - b0 : 65535;
            dim3 grid(cdivu(a->ncross, 256), cnt);
            corr_cross_kernel<<<grid, 256, 0, c->stream>>>(l, a->d_A + (long long)t.ka * ll, a->d_A + (long long)t.kb * ll, a->ncross, a->d_cross_idx,
                                                           t.d_rate, R + (long long)b0 * ll, X + (long long)b0 * ll);
            OQB_LAUNCH_CHECK(c);
        }
    return 0;
}

int oqb_ame_davies_stats(const oqb_ame* a, int64_t* n_terms, int64_t* n_cross, int64_t* n_lamb, int64_t* n_dense_groups, int64_t* n_dense_blocks) {
    if (!a) return OQB_EINVAL;
    if (n_terms) *n_terms = (int64_t)a->terms.size();
    if (n_cross) *n_cross = a->ncross;
    if (n_lamb) *n_lamb = a->nlamb;
    if (n_dense_groups) *n_dense_groups = (int64_t)a->h_dense_groups.size() + (a->zero_dense ? 1 : 0);
    if (n_dense_blocks) *n_dense_blocks = a->nd;
    return 0;
}

}  // extern "C"
