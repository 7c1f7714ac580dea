// Device-side GapIndices + Davies tables for closed-form spectra.
//
// The reference builds GapIndices (src/opensys/opensys_util.jl:25-61) with a double loop over the l(l−1)/2 level pairs
// and sorted insertion — O(l⁴) in the worst case, and it does so in EVERY RHS call of a time-dependent Hamiltonian
// (src/opensys/diffeq_liouvillian.jl:96).  Here the per-call part runs on the GPU:
//   1. gap_table_kernel    for every (a,b): the signed ROUNDED gap ω̃_ab (Julia's round(x, sigdigits): hidigit from an exact
//                          power-of-ten table, rint, same scaling rule), γ(ω̃_ab) of the Ohmic bath (src/bath/ohmic.jl:35-41)
//                          and S ≡ 0 straight into the l×l tables; for a<b also the pair's key into a sort buffer
//                          (zero-group pairs get +inf and are counted)
//   2. cub radix sort      (key, pair) — stable, so pairs inside a group keep the reference's (i<j) loop order
//   3. flag + compact      pairs whose key equals a neighbour's are the members of MULTI-pair groups; only they (and the
//                          zero-group pairs) go back to the host, which turns them into the cross-term lists exactly as
//                          gaps.GapGroups does (a few thousand entries at most)
//   4. ame_davies_finish   column sums / Hadamard coefficients / cross lists, as for host-supplied tables.
// Spectra given as arbitrary host closures keep using oqb_ame_set_davies with host-evaluated tables.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include <cmath>

#include "../../include/oqb.h"
#include "ame.cuh"
#include "bath.cuh"
#include "common.cuh"

using namespace oqb;

namespace {

__constant__ double c_pow10[31];

// Julia Base.round(x, sigdigits=n) for Float64 (floatfuncs.jl: hidigit + _round_digits); x != 0, finite.
__device__ __forceinline__ double round_sigdigits(double x, int sig) {
    const double ax = fabs(x);
    // hidigit = 1 + floor(log10|x|), from exact powers of ten (10^k is exact in binary64 for k ≤ 22)
    int h;
    if (ax >= 1.0) {
        h = 1;
        while (h < 30 && ax >= c_pow10[h]) ++h;  // 10^(h-1) ≤ |x| < 10^h
    } else {
        int k = 1;
        while (k < 30 && ax * c_pow10[k] < 1.0) ++k;  // |x|·10^k ≥ 1 first at k = −floor(log10|x|)
        // guard the scaled comparison against the rounding of ax·10^k near an exact power: compare by division too
        if (k > 1 && ax >= 1.0 / c_pow10[k - 1]) --k;
        h = 1 - k;
    }
    const int d = sig - h;
    const int ad = d < 0 ? -d : d;
    const double sc = ad <= 30 ? c_pow10[ad] : pow(10.0, (double)ad);
    const double r = d >= 0 ? rint(x * sc) / sc : rint(x / sc) * sc;
    return isfinite(r) ? r : x;
}

__global__ void gap_table_kernel(int l, const double* __restrict__ w, double thr, int sig, double eta, double wc, double beta,
                                 int nS, double S_x0, double S_dx, const double* __restrict__ S_coef,
                                 double* __restrict__ gam, double* __restrict__ S, double* __restrict__ keys,
                                 int32_t* __restrict__ pairs, unsigned long long* __restrict__ nzero) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)l * l) return;
    const int a = (int)(e % l), b = (int)(e / l);  // table entry (row a, col b) at a + b·l
    double key = 0.0;
    if (a != b) {
        const int i = a < b ? a : b, j = a < b ? b : a;
        const double gap = w[j] - w[i];  // the reference's loop variable (i<j)
        const bool zero = fabs(gap) <= thr;
        const double k = zero ? 0.0 : round_sigdigits(gap, sig);
        key = a < b ? k : -k;
        if (a < b) {
            const long long p = (long long)i * l - (long long)i * (i + 1) / 2 + (j - i - 1);  // lexicographic pair index
            keys[p] = zero ? INFINITY : k;
            pairs[p] = (int32_t)p;
            if (zero) atomicAdd(nzero, 1ull);
        }
    }
    gam[e] = ohmic_spectrum(key, eta, wc, beta);
    S[e] = nS ? bspline2_eval(key, nS, S_x0, S_dx, S_coef) : 0.0;  // tabulated Lamb shift (build_lambshift) or S ≡ 0
}

// one-sided AME tables at the UNROUNDED gap matrix ω_ba[a,b] = w[b] − w[a] (src/opensys/opensys_util.jl:65, davies.jl:368)
__global__ void onesided_table_kernel(int l, const double* __restrict__ w, double eta, double wc, double beta, int nS, double S_x0,
                                      double S_dx, const double* __restrict__ S_coef, double* __restrict__ gam,
                                      double* __restrict__ S) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)l * l) return;
    const int a = (int)(e % l), b = (int)(e / l);
    const double g = w[b] - w[a];
    gam[e] = ohmic_spectrum(g, eta, wc, beta);
    S[e] = nS ? bspline2_eval(g, nS, S_x0, S_dx, S_coef) : 0.0;
}

__global__ void flag_multi_kernel(long long m, const double* __restrict__ keys, uint8_t* __restrict__ flag) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const double v = keys[k];
    bool f = false;
    if (isfinite(v)) f = (k > 0 && keys[k - 1] == v) || (k + 1 < m && keys[k + 1] == v);
    flag[k] = f ? 1 : 0;
}

}  // namespace

extern "C" {

int oqb_ame_set_lambshift_spline(oqb_ame* a, int n, double x0, double dx, const double* h_coef) {
    OQB_DEVICE(a, a->ctx);
    if (!a) return OQB_EINVAL;
    oqb_ctx* c = a->ctx;
    if (n < 0 || (n > 0 && (!h_coef || !(dx > 0.0)))) return set_error(c, OQB_EINVAL, "oqb_ame_set_lambshift_spline: bad argument");
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    ame_free(c, a->d_Scoef);
    a->n_Stab = 0;
    if (n == 0) return 0;
    OQB_TRY(ame_alloc(c, (void**)&a->d_Scoef, (size_t)(n + 2) * 8));
    OQB_CUDA(c, cudaMemcpyAsync(a->d_Scoef, h_coef, (size_t)(n + 2) * 8, cudaMemcpyHostToDevice, c->stream));
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    a->n_Stab = n;
    a->S_x0 = x0;
    a->S_dx = dx;
    return 0;
}

int oqb_ame_set_onesided_ohmic(oqb_ame* a, int npairs, const int32_t* h_alpha, const int32_t* h_beta, double eta, double wc,
                               double beta) {
    OQB_DEVICE(a, a->ctx);
    if (!a) return OQB_EINVAL;
    oqb_ctx* c = a->ctx;
    if (!a->have_eigen) return set_error(c, OQB_EINVAL, "oqb_ame_set_onesided_ohmic: call oqb_ame_set_eigen first");
    if (npairs <= 0 || !h_alpha || !h_beta) return set_error(c, OQB_EINVAL, "oqb_ame_set_onesided_ohmic: bad arguments");
    const long long ll = (long long)a->lvl * a->lvl;
    double* d_g = nullptr;
    OQB_TRY(ame_alloc(c, (void**)&d_g, (size_t)ll * 2 * sizeof(double)));
    double* d_s = d_g + ll;
    onesided_table_kernel<<<(unsigned)((ll + 255) / 256), 256, 0, c->stream>>>(a->lvl, a->d_w, eta, wc, beta, a->n_Stab, a->S_x0,
                                                                               a->S_dx, a->d_Scoef, d_g, d_s);
    int s = 0;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) s = set_error(c, OQB_ECUDA, "onesided_table_kernel launch failed: %s", cudaGetErrorString(e));
    if (!s) s = ame_onesided_install(a, npairs, h_alpha, h_beta, d_g, d_s, 1);
    cudaStreamSynchronize(c->stream);
    cudaFreeAsync(d_g, c->stream);
    return s;
}

int oqb_ame_davies_ohmic_begin(oqb_ame* a, double eta, double wc, double beta, int digits, int sigdigits, int64_t* n_multi,
                               int64_t* n_zero) {
    OQB_DEVICE(a, a->ctx);
    if (!a) return OQB_EINVAL;
    oqb_ctx* c = a->ctx;
    if (!a->have_eigen) return set_error(c, OQB_EINVAL, "oqb_ame_davies_ohmic_begin: call oqb_ame_set_eigen first");
    if (!n_multi || !n_zero) return set_error(c, OQB_EINVAL, "oqb_ame_davies_ohmic_begin: null output");
    static DevOnce table_set;
    if (!table_set.done(c->device)) {
        double p10[31];
        for (int k = 0; k <= 30; ++k) {  // exactly the decimal literals "1e<k>" (correctly rounded), like gaps._POW10_TABLE
            char buf[8];
            snprintf(buf, sizeof(buf), "1e%d", k);
            p10[k] = strtod(buf, nullptr);
        }
        OQB_CUDA(c, cudaMemcpyToSymbol(c_pow10, p10, sizeof(p10)));
        table_set.set(c->device);
    }
    OQB_TRY(ame_davies_alloc(a));
    const int l = a->lvl;
    const long long ll = (long long)l * l, m = (long long)l * (l - 1) / 2;
    ame_free(c, a->d_grp_key);
    ame_free(c, a->d_grp_pair);
    a->n_multi = a->n_zero = 0;
    *n_multi = *n_zero = 0;
    if (m == 0) return 0;
    // scratch: keys/pairs (in + out), flags, counters, cub temp storage
    size_t tmp_sort = 0, tmp_sel = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, (const double*)nullptr, (double*)nullptr, (const int32_t*)nullptr,
                                    (int32_t*)nullptr, (int)m, 0, 64, c->stream);
    cub::DeviceSelect::Flagged(nullptr, tmp_sel, (const int32_t*)nullptr, (const uint8_t*)nullptr, (int32_t*)nullptr,
                               (unsigned long long*)nullptr, (int)m, c->stream);
    const size_t tmp = tmp_sort > tmp_sel ? tmp_sort : tmp_sel;
    const size_t bytes = (size_t)m * (8 + 8 + 4 + 4 + 1) + 256 + tmp + 2048;
    OQB_TRY(ws_reserve(c, bytes));
    char* base = (char*)c->ws;
    double* k_in = (double*)base;
    double* k_out = k_in + m;
    int32_t* p_in = (int32_t*)(k_out + m);
    int32_t* p_out = p_in + m;
    uint8_t* flag = (uint8_t*)(p_out + m);
    unsigned long long* cnt = (unsigned long long*)(((uintptr_t)(flag + m) + 255) & ~(uintptr_t)255);
    void* cub_tmp = (void*)(((uintptr_t)(cnt + 4) + 255) & ~(uintptr_t)255);  // cub wants a 256-byte aligned blob
    OQB_CUDA(c, cudaMemsetAsync(cnt, 0, 4 * sizeof(unsigned long long), c->stream));
    gap_table_kernel<<<(unsigned)((ll + 255) / 256), 256, 0, c->stream>>>(l, a->d_w, pow(10.0, -(double)digits), sigdigits, eta, wc,
                                                                          beta, a->n_Stab, a->S_x0, a->S_dx, a->d_Scoef, a->d_gam, a->d_S,
                                                                          k_in, p_in, cnt);
    OQB_LAUNCH_CHECK(c);
    size_t t1 = tmp;
    OQB_CUDA(c, cub::DeviceRadixSort::SortPairs(cub_tmp, t1, k_in, k_out, p_in, p_out, (int)m, 0, 64, c->stream));
    flag_multi_kernel<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(m, k_out, flag);
    OQB_LAUNCH_CHECK(c);
    // compact the flagged pairs (order preserved) — pair ids into p_in, keys into k_in (both free again)
    size_t t2 = tmp;
    OQB_CUDA(c, cub::DeviceSelect::Flagged(cub_tmp, t2, p_out, flag, p_in, cnt + 1, (int)m, c->stream));
    size_t t3 = tmp;
    OQB_CUDA(c, cub::DeviceSelect::Flagged(cub_tmp, t3, k_out, flag, k_in, cnt + 2, (int)m, c->stream));
    unsigned long long h_cnt[4];
    OQB_CUDA(c, cudaMemcpyAsync(h_cnt, cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, c->stream));
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    const long long nz = (long long)h_cnt[0], nm = (long long)h_cnt[1];
    // keep: [multi pairs (nm) | zero pairs (nz)] and the multi keys; zero-group pairs are the +inf tail of the sorted list
    const long long keep = nm + nz;
    if (keep > 0) {
        OQB_TRY(ame_alloc(c, (void**)&a->d_grp_pair, (size_t)keep * sizeof(int32_t)));
        OQB_TRY(ame_alloc(c, (void**)&a->d_grp_key, (size_t)(nm ? nm : 1) * sizeof(double)));
        if (nm) {
            OQB_CUDA(c, cudaMemcpyAsync(a->d_grp_pair, p_in, (size_t)nm * 4, cudaMemcpyDeviceToDevice, c->stream));
            OQB_CUDA(c, cudaMemcpyAsync(a->d_grp_key, k_in, (size_t)nm * 8, cudaMemcpyDeviceToDevice, c->stream));
        }
        if (nz) OQB_CUDA(c, cudaMemcpyAsync(a->d_grp_pair + nm, p_out + (m - nz), (size_t)nz * 4, cudaMemcpyDeviceToDevice, c->stream));
        OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    a->n_multi = nm;
    a->n_zero = nz;
    *n_multi = nm;
    *n_zero = nz;
    return 0;
}

int oqb_ame_davies_fetch_groups(oqb_ame* a, int32_t* h_multi_pair, double* h_multi_key, int32_t* h_zero_pair) {
    OQB_DEVICE(a, a->ctx);
    if (!a) return OQB_EINVAL;
    oqb_ctx* c = a->ctx;
    if (a->n_multi > 0) {
        if (!h_multi_pair || !h_multi_key) return set_error(c, OQB_EINVAL, "oqb_ame_davies_fetch_groups: null output");
        OQB_CUDA(c, cudaMemcpyAsync(h_multi_pair, a->d_grp_pair, (size_t)a->n_multi * 4, cudaMemcpyDeviceToHost, c->stream));
        OQB_CUDA(c, cudaMemcpyAsync(h_multi_key, a->d_grp_key, (size_t)a->n_multi * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    if (a->n_zero > 0) {
        if (!h_zero_pair) return set_error(c, OQB_EINVAL, "oqb_ame_davies_fetch_groups: null output");
        OQB_CUDA(c, cudaMemcpyAsync(h_zero_pair, a->d_grp_pair + a->n_multi, (size_t)a->n_zero * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int oqb_ame_davies_ohmic_finish(oqb_ame* a, int64_t ncross, const int32_t* h_cross_idx, const double* h_cross_rate, int64_t nlamb,
                                const int32_t* h_lamb_idx, const double* h_lamb_rate_S) {
    OQB_DEVICE(a, a->ctx);
    if (!a) return OQB_EINVAL;
    oqb_ctx* c = a->ctx;
    if (!a->d_gam) return set_error(c, OQB_EINVAL, "oqb_ame_davies_ohmic_finish: call oqb_ame_davies_ohmic_begin first");
    if ((ncross > 0 && (!h_cross_idx || !h_cross_rate)) || (nlamb > 0 && (!h_lamb_idx || !h_lamb_rate_S)))
        return set_error(c, OQB_EINVAL, "oqb_ame_davies_ohmic_finish: null cross-term arrays");
    ame_free(c, a->d_grp_key);
    ame_free(c, a->d_grp_pair);
    return ame_davies_finish(a, ncross, h_cross_idx, h_cross_rate, nlamb, h_lamb_idx, h_lamb_rate_S);
}

}  // extern "C"
