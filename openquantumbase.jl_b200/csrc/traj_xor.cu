// K7 for Pauli-structured operators: dψ_b = Σ_t c_bt M_t ψ_b where every non-diagonal term is in XOR
// form, M[r, r ^ mask_j] = v_j (X-type Pauli strings: transverse fields, XX couplings, …) and the
// diagonal terms (Z-type strings, −½Σγ_k L_k'L_k) are real vectors.
//
// The operator then needs NO per-row storage: columns are implicit (r ^ mask_j), values are per-slot
// scalars in the kernel parameters, real diagonals are cached in registers.  Memory traffic per
// trajectory is exactly the algorithmic 32n bytes (ψ in, dψ out).  ψ is staged by cp.async.bulk (UBLKCP)
// into a 3-deep shared-memory ring guarded by full/empty mbarriers; warps run free of any CTA-wide
// barrier and may be up to two trajectories apart, so HBM reads, shared-memory gathers and the
// coalesced dψ stores of neighbouring trajectories overlap.
#include "sparse_terms.cuh"

namespace oqb {

namespace {

constexpr int kXSlots = 16, kXTerms = 4, kXDiag = 6;

struct XorDesc {
    int n_x, n_diag;
    int x_begin[kXTerms + 1];
    int x_nrem[kXTerms];  // slots [begin, begin+nrem) need a shared-memory gather,
    int x_nlane[kXTerms]; // the next nlane are warp-local (mask < 32: shuffles), the rest thread-local (registers)
    int x_coef[kXTerms];
    int mask[kXSlots];
    double vre[kXSlots], vim[kXSlots];
    const double* diag_r[kXDiag];
    int diag_coef[kXDiag];
    double fconst_re, fconst_im;  // constant diagonal contribution (e.g. −½Σγ_k when every L_k'L_k = c_k·I)
};

__device__ __forceinline__ unsigned s_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void xb_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void xb_expect(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void xb_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void xb_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "X_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra X_DONE;\n\t"
        "bra X_LOOP;\n\t"
        "X_DONE:\n\t"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void xb_bulk(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// acc[kk] += v · xo[(c·CH + kk) ^ HK]  — partner rows that live in this thread's own registers
template <int R, int CH, int HK, bool RV>
__device__ __forceinline__ void local_acc(const c128 (&xo)[R], int c, double (&ar)[CH], double (&ai)[CH], double vr, double vi) {
#pragma unroll
    for (int cc = 0; cc < R / CH; ++cc) {
        if (cc != c) continue;
#pragma unroll
        for (int kk = 0; kk < CH; ++kk) {
            const c128 v = xo[(cc * CH + kk) ^ HK];
            if (RV) {
                ar[kk] = fma(vr, v.x, ar[kk]);
                ai[kk] = fma(vr, v.y, ai[kk]);
            } else {
                ar[kk] = fma(vr, v.x, ar[kk]); ar[kk] = fma(-vi, v.y, ar[kk]);
                ai[kk] = fma(vr, v.y, ai[kk]); ai[kk] = fma(vi, v.x, ai[kk]);
            }
        }
    }
}
template <int R, int CH, bool RV>
__device__ __forceinline__ void local_dispatch(int hk, const c128 (&xo)[R], int c, double (&ar)[CH], double (&ai)[CH], double vr,
                                               double vi) {
#define OQB_LC(H) case H: if (H < R) local_acc<R, CH, (H < R ? H : 0), RV>(xo, c, ar, ai, vr, vi); break;
    switch (hk) {
        OQB_LC(1) OQB_LC(2) OQB_LC(3) OQB_LC(4) OQB_LC(5) OQB_LC(6) OQB_LC(7) OQB_LC(8)
        OQB_LC(9) OQB_LC(10) OQB_LC(11) OQB_LC(12) OQB_LC(13) OQB_LC(14) OQB_LC(15)
        default: break;
    }
#undef OQB_LC
}

// R rows per thread (r = tid + k·NT), S ring stages, RV: all slot values real, ND: real diagonal terms whose
// values are cached in registers.  A thread owns the rows that differ only in the top log2(R) index bits, so
// XOR partners across those bits are its own registers (no shared-memory gather).
template <int R, int NT, int S, bool RV, int ND>
__global__ void __launch_bounds__(NT, 1)
    traj_xor_kernel(int nb, XorDesc d, const c128* __restrict__ coef, long long coef_ld, const c128* __restrict__ psi,
                    c128* __restrict__ dpsi) {
    constexpr int n = R * NT, NW = NT / 32, CH = R < 8 ? R : 8;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    c128* ring = reinterpret_cast<c128*>(smem_raw);
    __shared__ __align__(8) unsigned long long bars[2 * S];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned bar_base = s_addr(bars);
    const unsigned bytes = n * (unsigned)sizeof(c128);

    double dreg[R][ND > 0 ? ND : 1];
#pragma unroll
    for (int k = 0; k < R; ++k)
#pragma unroll
        for (int t = 0; t < ND; ++t) dreg[k][t] = d.diag_r[t][tid + k * NT];

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            xb_init(bar_base + 8 * s, 1);         // full
            xb_init(bar_base + 8 * (S + s), NW);  // empty
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const int stride = gridDim.x;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < S - 1; ++i) {
            const long long b = blockIdx.x + (long long)i * stride;
            if (b < nb) {
                xb_expect(bar_base + 8 * i, bytes);
                xb_bulk(s_addr(ring + (size_t)i * n), psi + (size_t)b * n, bytes, bar_base + 8 * i);
            }
        }
    }
    int i = 0;
    for (long long b = blockIdx.x; b < nb; b += stride, ++i) {
        if (warp == 0) {
            if (lane == 0) {  // keep S-1 trajectories in flight
                const int li = i + S - 1;
                const long long lb = b + (long long)(S - 1) * stride;
                if (lb < nb) {
                    const int ls = li % S;
                    xb_wait(bar_base + 8 * (S + ls), (unsigned)(((li / S) & 1) ^ 1));
                    xb_expect(bar_base + 8 * ls, bytes);
                    xb_bulk(s_addr(ring + (size_t)ls * n), psi + (size_t)lb * n, bytes, bar_base + 8 * ls);
                }
            }
            __syncwarp();
        }
        const int s = i % S;
        const c128* __restrict__ cf = coef + b * coef_ld;
        c128 cdiag[ND > 0 ? ND : 1];
#pragma unroll
        for (int t = 0; t < ND; ++t) cdiag[t] = __ldg(cf + d.diag_coef[t]);
        xb_wait(bar_base + 8 * s, (unsigned)((i / S) & 1));
        const c128* __restrict__ x = ring + (size_t)s * n;
        c128 xo[R];
#pragma unroll
        for (int k = 0; k < R; ++k) xo[k] = x[tid + k * NT];
        c128* __restrict__ out = dpsi + (size_t)b * n;
#pragma unroll
        for (int c = 0; c < R / CH; ++c) {
            double outr[CH], outi[CH];
#pragma unroll
            for (int kk = 0; kk < CH; ++kk) {  // diagonal part: (const + Σ_t c_bt d_t[r]) ψ[r]
                const int k = c * CH + kk;
                double fr = d.fconst_re, fi = d.fconst_im;
#pragma unroll
                for (int t = 0; t < ND; ++t) {
                    fr = fma(cdiag[t].x, dreg[k][t], fr);
                    fi = fma(cdiag[t].y, dreg[k][t], fi);
                }
                outr[kk] = fr * xo[k].x - fi * xo[k].y;
                outi[kk] = fr * xo[k].y + fi * xo[k].x;
            }
            for (int e = 0; e < d.n_x; ++e) {
                double ar[CH], ai[CH];
#pragma unroll
                for (int kk = 0; kk < CH; ++kk) ar[kk] = ai[kk] = 0.0;
                const int jb = d.x_begin[e], nrem = d.x_nrem[e];
                // remote partners: fully unrolled so that the gathers of slot j+1 are in flight while slot j is consumed
#pragma unroll
                for (int j = 0; j < kXSlots; ++j) {
                    if (j < nrem) {
                        const int m = d.mask[jb + j];
                        const double vr = d.vre[jb + j], vi = d.vim[jb + j];
#pragma unroll
                        for (int kk = 0; kk < CH; ++kk) {
                            const c128 v = x[(tid + (c * CH + kk) * NT) ^ m];
                            if (RV) {
                                ar[kk] = fma(vr, v.x, ar[kk]);
                                ai[kk] = fma(vr, v.y, ai[kk]);
                            } else {
                                ar[kk] = fma(vr, v.x, ar[kk]); ar[kk] = fma(-vi, v.y, ar[kk]);
                                ai[kk] = fma(vr, v.y, ai[kk]); ai[kk] = fma(vi, v.x, ai[kk]);
                            }
                        }
                    }
                }
                // partners inside the warp (mask < 32): register shuffles instead of shared-memory gathers
                const int nlane = d.x_nlane[e];
                for (int j = jb + nrem; j < jb + nrem + nlane; ++j) {
                    const int m = d.mask[j];
                    const double vr = d.vre[j], vi = d.vim[j];
#pragma unroll
                    for (int kk = 0; kk < CH; ++kk) {
                        c128 v;
                        v.x = __shfl_xor_sync(0xffffffffu, xo[c * CH + kk].x, m);
                        v.y = __shfl_xor_sync(0xffffffffu, xo[c * CH + kk].y, m);
                        if (RV) {
                            ar[kk] = fma(vr, v.x, ar[kk]);
                            ai[kk] = fma(vr, v.y, ai[kk]);
                        } else {
                            ar[kk] = fma(vr, v.x, ar[kk]); ar[kk] = fma(-vi, v.y, ar[kk]);
                            ai[kk] = fma(vr, v.y, ai[kk]); ai[kk] = fma(vi, v.x, ai[kk]);
                        }
                    }
                }
                // partners that differ only in thread-local index bits: own registers
                for (int j = jb + nrem + nlane; j < d.x_begin[e + 1]; ++j)
                    local_dispatch<R, CH, RV>(d.mask[j] / NT, xo, c, ar, ai, d.vre[j], d.vim[j]);
                const c128 cc = __ldg(cf + d.x_coef[e]);
#pragma unroll
                for (int kk = 0; kk < CH; ++kk) {
                    outr[kk] = fma(cc.x, ar[kk], outr[kk]); outr[kk] = fma(-cc.y, ai[kk], outr[kk]);
                    outi[kk] = fma(cc.x, ai[kk], outi[kk]); outi[kk] = fma(cc.y, ar[kk], outi[kk]);
                }
            }
#pragma unroll
            for (int kk = 0; kk < CH; ++kk) out[tid + (c * CH + kk) * NT] = make_double2(outr[kk], outi[kk]);
        }
        __syncwarp();
        if (lane == 0) xb_arrive(bar_base + 8 * (S + s));
    }
}

template <int R, int NT, int S, bool RV, int ND>
int launch(Ctx* c, int nb, const XorDesc& d_in, const c128* coef, long long coef_ld, const c128* psi, c128* dpsi) {
    XorDesc d = d_in;
    for (int e = 0; e < d.n_x; ++e) {  // remote slots first, thread-local ones (mask % NT == 0) last
        static const bool use_shfl = [] { const char* e = getenv("OQB_XOR_SHFL"); return e && e[0] == '1'; }();  // measured slower than LDS gathers: off by default
        auto cls = [&](int m) { return (m & (NT - 1)) == 0 ? 2 : ((use_shfl && m > 0 && m < 32) ? 1 : 0); };
        int w = d.x_begin[e], cnt[3] = {0, 0, 0};
        for (int pass = 0; pass < 3; ++pass)
            for (int j = d_in.x_begin[e]; j < d_in.x_begin[e + 1]; ++j)
                if (cls(d_in.mask[j]) == pass) {
                    d.mask[w] = d_in.mask[j]; d.vre[w] = d_in.vre[j]; d.vim[w] = d_in.vim[j];
                    ++w; ++cnt[pass];
                }
        d.x_nrem[e] = cnt[0];
        d.x_nlane[e] = cnt[1];
    }
    {   // timing experiments only (results are wrong when set): cap the number of slots per term
        static const int cap = [] { const char* e = getenv("OQB_XOR_DEBUG_MAXSLOTS"); return e ? atoi(e) : -1; }();
        if (cap >= 0)
            for (int e = 0; e < d.n_x; ++e) {
                if (d.x_nrem[e] > cap) d.x_nrem[e] = cap;
                if (d.x_nrem[e] + d.x_nlane[e] > cap) d.x_nlane[e] = cap - d.x_nrem[e];
                if (d.x_begin[e + 1] - d.x_begin[e] > cap) d.x_begin[e + 1] = d.x_begin[e] + cap;
            }
    }
    const size_t smem = (size_t)S * R * NT * sizeof(c128);
    static bool cfg = false;
    if (!cfg) {
        OQB_CUDA(c, cudaFuncSetAttribute(traj_xor_kernel<R, NT, S, RV, ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cfg = true;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    int per_sm = (int)((224 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2048 / NT) per_sm = 2048 / NT;
    const int grid = nb < sms * per_sm ? nb : sms * per_sm;
    {
        ProfScope prof(c, 32.0 * R * NT * (double)nb);  // algorithmic bytes: read ψ + write dψ
        traj_xor_kernel<R, NT, S, RV, ND><<<grid, NT, smem, c->stream>>>(nb, d, coef, coef_ld, psi, dpsi);
    }
    OQB_LAUNCH_CHECK(c);
    return 0;
}

template <bool RV, int ND>
int by_size(Ctx* c, int n, int nb, const XorDesc& d, const c128* coef, long long coef_ld, const c128* psi, c128* dpsi,
            bool* used) {
    *used = true;
    switch (n) {
        case 4096: {
            static const int rsel = [] { const char* e = getenv("OQB_XOR_R"); return e ? atoi(e) : 4; }();
            if (rsel == 16) return launch<16, 256, 3, RV, ND>(c, nb, d, coef, coef_ld, psi, dpsi);
            if (rsel == 4) return launch<4, 1024, 3, RV, ND>(c, nb, d, coef, coef_ld, psi, dpsi);
            return launch<8, 512, 3, RV, ND>(c, nb, d, coef, coef_ld, psi, dpsi);
        }
        case 2048: return launch<8, 256, 3, RV, ND>(c, nb, d, coef, coef_ld, psi, dpsi);
        case 1024: return launch<4, 256, 3, RV, ND>(c, nb, d, coef, coef_ld, psi, dpsi);
        case 512: return launch<4, 128, 3, RV, ND>(c, nb, d, coef, coef_ld, psi, dpsi);
        case 256: return launch<2, 128, 3, RV, ND>(c, nb, d, coef, coef_ld, psi, dpsi);
        default: *used = false; return 0;
    }
}

}  // namespace

int traj_matvec_xor(Ctx* c, int n, int nb, const std::vector<const TermDev*>& ts, const c128* d_coef, long long coef_ld,
                    const c128* psi, c128* dpsi, bool* used) {
    *used = false;
    static const int mode = [] { const char* e = getenv("OQB_TRAJ"); return e ? (int)e[0] : 0; }();
    if (mode == 'g' || mode == 'f') return 0;  // "generic" / "fast": skip this path (A/B measurements)
    XorDesc d;
    d.n_x = d.n_diag = 0;
    d.fconst_re = d.fconst_im = 0.0;
    d.x_begin[0] = 0;
    int slots = 0;
    bool all_real = true;
    for (size_t t = 0; t < ts.size(); ++t) {
        const TermDev* td = ts[t];
        if (td->is_diag) {
            if (td->diag_const) {  // c_bt must be shared for a constant term to fold into a scalar
                if (coef_ld != 0 && !td->const_coef_one) return 0;
                d.fconst_re += td->const_val.x;
                d.fconst_im += td->const_val.y;
                continue;
            }
            if (!(td->is_real && td->rvals) || d.n_diag >= 2) return 0;
            d.diag_r[d.n_diag] = td->rvals;
            d.diag_coef[d.n_diag] = (int)t;
            d.n_diag++;
        } else {
            if (!td->xor_form || d.n_x >= kXTerms || slots + (int)td->xor_masks.size() > kXSlots) return 0;
            for (size_t j = 0; j < td->xor_masks.size(); ++j) {
                d.mask[slots] = td->xor_masks[j];
                d.vre[slots] = td->xor_vals[j].x;
                d.vim[slots] = td->xor_vals[j].y;
                all_real = all_real && td->xor_vals[j].y == 0.0;
                ++slots;
            }
            d.x_coef[d.n_x] = (int)t;
            d.n_x++;
            d.x_begin[d.n_x] = slots;
        }
    }
    for (int j = slots; j < kXSlots; ++j) { d.mask[j] = 0; d.vre[j] = d.vim[j] = 0.0; }
    for (int t = d.n_diag; t < kXDiag; ++t) { d.diag_r[t] = nullptr; d.diag_coef[t] = 0; }
    if (d.n_diag == 0) return all_real ? by_size<true, 0>(c, n, nb, d, d_coef, coef_ld, psi, dpsi, used)
                                       : by_size<false, 0>(c, n, nb, d, d_coef, coef_ld, psi, dpsi, used);
    if (d.n_diag == 1) return all_real ? by_size<true, 1>(c, n, nb, d, d_coef, coef_ld, psi, dpsi, used)
                                       : by_size<false, 1>(c, n, nb, d, d_coef, coef_ld, psi, dpsi, used);
    return all_real ? by_size<true, 2>(c, n, nb, d, d_coef, coef_ld, psi, dpsi, used)
                    : by_size<false, 2>(c, n, nb, d, d_coef, coef_ld, psi, dpsi, used);
}

}  // namespace oqb
