// K7 for Pauli-structured operators: dψ_b = Σ_t c_bt M_t ψ_b where every non-diagonal term is in XOR
// form, M[r, r ^ mask_j] = v_j (X-type Pauli strings: transverse fields, XX couplings, …) and the
// diagonal terms (Z-type strings, −½Σγ_k L_k'L_k) are real vectors or constants.
//
// The operator then needs NO per-row storage: columns are implicit (r ^ mask_j), values are per-slot
// scalars, real diagonals are cached in registers.  Memory traffic per trajectory is exactly the
// algorithmic 32n bytes (ψ in, dψ out).  ψ and the trajectory's coefficient row are staged by
// cp.async.bulk (UBLKCP, mbarrier complete_tx) into a 3-deep shared-memory ring guarded by full/empty
// mbarriers; warps run free of any CTA-wide barrier and may be up to two trajectories apart, so HBM
// reads, shared-memory gathers and the coalesced dψ stores of neighbouring trajectories overlap.
// A thread owns the R rows that differ only in the top log2(R) index bits: XOR partners across those
// bits are its own registers; the remaining partners are gathered four slots at a time (4·8 independent
// LDS.128 in flight per thread).  When the operator has the transverse-field structure (ONE X-type term whose
// thread-local slots are exactly the single-bit flips NT·{1,2,4,…}) the slot loops are compiled with fixed bounds
// (template parameter NG): no runtime loop bounds, no dispatch switch, no parameter loads inside the row loop — on C4
// this alone took the kernel from 71 % to 98 % of the measured HBM bandwidth (1.865 → 1.359 ms).
#include <cstring>
#include <utility>

#include "sparse_terms.cuh"

namespace oqb {

namespace {

#ifndef OQB_XOR_SPR
#define OQB_XOR_SPR 4
#endif
constexpr int kSPR = OQB_XOR_SPR;  // gather slots issued back to back per round
constexpr int kXSlots = 28, kXTerms = 4, kXDiag = 2, kCoefPad = 16;  // coefficient area: 16 complex per stage

struct XorDesc {
    int n_x, n_diag, ncoef, coef_shared;
    int x_begin[kXTerms + 1];
    int x_nrem[kXTerms];  // slots [begin, begin+nrem) are shared-memory gathers (count padded to even), rest thread-local
    int x_coef[kXTerms];
    int mask[kXSlots + 4];
    // two-valued slot: M[r, r^mask] = (popcount(r & zmask) odd ? w : v) — one Pauli string (w = −v: Y / Z factors next to its X
    // factors) or two strings sharing their flip mask (XX + YY: v, w ∈ {0, 2J}); zmask = 0 ⇒ one value
    int zmask[kXSlots + 4];
    unsigned ktab[kXSlots + 4];  // bit k = parity((k·NT) & zmask): the part of the row parity that depends on the thread's k-th row
    int any_signed;
    int x_imag[kXTerms];         // the term's slot values are purely imaginary: stored as real numbers, the factor i goes onto the coefficient
    double vre[kXSlots + 4], vim[kXSlots + 4], wre[kXSlots + 4], wim[kXSlots + 4];
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

template <bool RV>
__device__ __forceinline__ void cacc(double& ar, double& ai, double vr, double vi, const c128 v) {
    if (RV) {
        ar = fma(vr, v.x, ar);
        ai = fma(vr, v.y, ai);
    } else {
        ar = fma(vr, v.x, ar); ar = fma(-vi, v.y, ar);
        ai = fma(vr, v.y, ai); ai = fma(vi, v.x, ai);
    }
}

// acc[kk] += v · xo[(C0 + kk) ^ HK]  — partner rows that live in this thread's own registers
template <int R, int CH, int C0, int HK, bool RV>
__device__ __forceinline__ void local_acc(const c128 (&xo)[R], double (&ar)[CH], double (&ai)[CH], double vr, double vi, unsigned par = 0,
                                          double wr = 0.0, double wi = 0.0) {
    // par: bit k set ⇒ row k of this thread takes the slot's second value (w) instead of the first (v)
#pragma unroll
    for (int kk = 0; kk < CH; ++kk) {
        const bool odd = (par >> (C0 + kk)) & 1u;
        cacc<RV>(ar[kk], ai[kk], odd ? wr : vr, odd ? wi : vi, xo[((C0 + kk) ^ HK) % R]);
    }
}
template <int R, int CH, int C0, bool RV>
__device__ __forceinline__ void local_dispatch(int hk, const c128 (&xo)[R], double (&ar)[CH], double (&ai)[CH], double vr, double vi,
                                               unsigned par = 0, double wr = 0.0, double wi = 0.0) {
#define OQB_LC(H) case H: local_acc<R, CH, C0, H, RV>(xo, ar, ai, vr, vi, par, wr, wi); break;
    if (hk < 4) {
        switch (hk) { OQB_LC(1) OQB_LC(2) OQB_LC(3) default: break; }
    } else if (R > 4 && hk < 8) {
        switch (hk) { OQB_LC(4) OQB_LC(5) OQB_LC(6) OQB_LC(7) default: break; }
    } else if (R > 8) {
        switch (hk) { OQB_LC(8) OQB_LC(9) OQB_LC(10) OQB_LC(11) OQB_LC(12) OQB_LC(13) OQB_LC(14) OQB_LC(15) default: break; }
    }
#undef OQB_LC
}

// one chunk of CH rows (rows C0 … C0+CH-1 of this thread)
template <int R, int NT, int CH, int C0, bool RV, int ND, int EPI, int NG, bool SG>
__device__ __forceinline__ void do_chunk(const XorDesc& d, const int* s_mask, const unsigned* s_kt, unsigned pmask, const double* s_vre,
                                         const double* s_vim, const double* s_wre, const double* s_wim,
                                         const c128* __restrict__ x, const c128* cfs, const c128 (&xo)[R],
                                         const double (&dreg)[R][ND > 0 ? ND : 1], const c128 (&cdiag)[ND > 0 ? ND : 1],
                                         c128* out, int tid, double wa, double wt, const c128* __restrict__ rk_psi0,
                                         c128* __restrict__ rk_acc, double& nrm) {
    // fused RK4 stage (EPI != 0): the extra operands are requested before the gathers so that their HBM latency
    // hides behind the shared-memory work of the chunk
    c128 e0[EPI == 2 ? CH : 1], e1[EPI >= 2 ? CH : 1];
    if (EPI == 2) {
#pragma unroll
        for (int kk = 0; kk < CH; ++kk) e0[kk] = rk_psi0[tid + (C0 + kk) * NT];
    }
    if (EPI >= 2) {
#pragma unroll
        for (int kk = 0; kk < CH; ++kk) e1[kk] = rk_acc[tid + (C0 + kk) * NT];
    }
    double outr[CH], outi[CH];
#pragma unroll
    for (int kk = 0; kk < CH; ++kk) {  // diagonal part: (const + Σ_t c_bt d_t[r]) ψ[r]
        const int k = C0 + kk;
        double fr = d.fconst_re, fi = d.fconst_im;
#pragma unroll
        for (int t = 0; t < ND; ++t) {
            fr = fma(cdiag[t].x, dreg[k][t], fr);
            fi = fma(cdiag[t].y, dreg[k][t], fi);
        }
        outr[kk] = fr * xo[k].x - fi * xo[k].y;
        outi[kk] = fr * xo[k].y + fi * xo[k].x;
    }
    if constexpr (NG > 0) {
        // Fixed structure (one X-type term: NG gathered slots, then the single-bit thread-local slots NT·{1,2,4,…} in
        // ascending order): no runtime loop bounds, no dispatch switch — the whole chunk is straight-line code.
        double ar[CH], ai[CH];
#pragma unroll
        for (int kk = 0; kk < CH; ++kk) ar[kk] = ai[kk] = 0.0;
#pragma unroll
        for (int j = 0; j < NG; j += kSPR) {
            c128 vv[kSPR][CH];
#pragma unroll
            for (int q = 0; q < kSPR; ++q) {
                const int m = s_mask[j + q];
#pragma unroll
                for (int kk = 0; kk < CH; ++kk) vv[q][kk] = x[(tid + (C0 + kk) * NT) ^ m];
            }
#pragma unroll
            for (int q = 0; q < kSPR; ++q) {
                const double vr = s_vre[j + q], vi = s_vim[j + q];
                if (SG) {
                    const unsigned par = s_kt[j + q] ^ (0u - ((pmask >> (j + q)) & 1u));
                    const double wr = s_wre[j + q], wi = s_wim[j + q];
#pragma unroll
                    for (int kk = 0; kk < CH; ++kk) {
                        const bool odd = (par >> (C0 + kk)) & 1u;
                        cacc<RV>(ar[kk], ai[kk], odd ? wr : vr, odd ? wi : vi, vv[q][kk]);
                    }
                } else {
#pragma unroll
                    for (int kk = 0; kk < CH; ++kk) cacc<RV>(ar[kk], ai[kk], vr, vi, vv[q][kk]);
                }
            }
        }
        if (SG) {
#define OQB_PAR(J) (s_kt[J] ^ (0u - ((pmask >> (J)) & 1u)))
            local_acc<R, CH, C0, 1, RV>(xo, ar, ai, s_vre[NG], s_vim[NG], OQB_PAR(NG), s_wre[NG], s_wim[NG]);
            if (R > 2) local_acc<R, CH, C0, 2 % R, RV>(xo, ar, ai, s_vre[NG + 1], s_vim[NG + 1], OQB_PAR(NG + 1), s_wre[NG + 1], s_wim[NG + 1]);
            if (R > 4) local_acc<R, CH, C0, 4 % R, RV>(xo, ar, ai, s_vre[NG + 2], s_vim[NG + 2], OQB_PAR(NG + 2), s_wre[NG + 2], s_wim[NG + 2]);
            if (R > 8) local_acc<R, CH, C0, 8 % R, RV>(xo, ar, ai, s_vre[NG + 3], s_vim[NG + 3], OQB_PAR(NG + 3), s_wre[NG + 3], s_wim[NG + 3]);
#undef OQB_PAR
        } else {
            local_acc<R, CH, C0, 1, RV>(xo, ar, ai, s_vre[NG], s_vim[NG]);
            if (R > 2) local_acc<R, CH, C0, 2 % R, RV>(xo, ar, ai, s_vre[NG + 1], s_vim[NG + 1]);
            if (R > 4) local_acc<R, CH, C0, 4 % R, RV>(xo, ar, ai, s_vre[NG + 2], s_vim[NG + 2]);
            if (R > 8) local_acc<R, CH, C0, 8 % R, RV>(xo, ar, ai, s_vre[NG + 3], s_vim[NG + 3]);
        }
        c128 cc = cfs[d.x_coef[0]];
        if (SG && d.x_imag[0]) cc = make_double2(-cc.y, cc.x);  // the term's slot values were stored without their factor i
#pragma unroll
        for (int kk = 0; kk < CH; ++kk) {
            outr[kk] = fma(cc.x, ar[kk], outr[kk]); outr[kk] = fma(-cc.y, ai[kk], outr[kk]);
            outi[kk] = fma(cc.x, ai[kk], outi[kk]); outi[kk] = fma(cc.y, ar[kk], outi[kk]);
        }
    } else {
        for (int e = 0; e < d.n_x; ++e) {
            double ar[CH], ai[CH];
    #pragma unroll
            for (int kk = 0; kk < CH; ++kk) ar[kk] = ai[kk] = 0.0;
            const int jb = d.x_begin[e], nrem = d.x_nrem[e];
            for (int j = jb; j < jb + nrem; j += kSPR) {  // kSPR slots per round: kSPR·CH independent gathers in flight
                c128 vv[kSPR][CH];
    #pragma unroll
                for (int q = 0; q < kSPR; ++q) {
                    const int m = s_mask[j + q];
    #pragma unroll
                    for (int kk = 0; kk < CH; ++kk) vv[q][kk] = x[(tid + (C0 + kk) * NT) ^ m];
                }
    #pragma unroll
                for (int q = 0; q < kSPR; ++q) {
                    const double vr = s_vre[j + q], vi = s_vim[j + q];
                    if (SG) {
                        // row parity = parity(tid & z) ^ parity(k·NT & z): bit k of `par` says which of the slot's two values row k takes
                        const unsigned par = s_kt[j + q] ^ (0u - ((pmask >> (j + q)) & 1u));
                        const double wr = s_wre[j + q], wi = s_wim[j + q];
    #pragma unroll
                        for (int kk = 0; kk < CH; ++kk) {
                            const bool odd = (par >> (C0 + kk)) & 1u;
                            cacc<RV>(ar[kk], ai[kk], odd ? wr : vr, odd ? wi : vi, vv[q][kk]);
                        }
                    } else {
    #pragma unroll
                        for (int kk = 0; kk < CH; ++kk) cacc<RV>(ar[kk], ai[kk], vr, vi, vv[q][kk]);
                    }
                }
            }
            for (int j = jb + nrem; j < d.x_begin[e + 1]; ++j) {  // partners in this thread's own registers
                if (SG) local_dispatch<R, CH, C0, RV>(s_mask[j] / NT, xo, ar, ai, s_vre[j], s_vim[j], s_kt[j] ^ (0u - ((pmask >> j) & 1u)), s_wre[j], s_wim[j]);
                else local_dispatch<R, CH, C0, RV>(s_mask[j] / NT, xo, ar, ai, s_vre[j], s_vim[j]);
            }
            c128 cc = cfs[d.x_coef[e]];
            if (SG && d.x_imag[e]) cc = make_double2(-cc.y, cc.x);  // the term's slot values were stored without their factor i
    #pragma unroll
            for (int kk = 0; kk < CH; ++kk) {
                outr[kk] = fma(cc.x, ar[kk], outr[kk]); outr[kk] = fma(-cc.y, ai[kk], outr[kk]);
                outi[kk] = fma(cc.x, ai[kk], outi[kk]); outi[kk] = fma(cc.y, ar[kk], outi[kk]);
            }
        }
    }
#pragma unroll
    for (int kk = 0; kk < CH; ++kk) {
        const int idx = tid + (C0 + kk) * NT;
        if (EPI == 0) {
            out[idx] = make_double2(outr[kk], outi[kk]);
        } else if (EPI == 1) {
            const c128 p0 = xo[C0 + kk];
            rk_acc[idx] = make_double2(p0.x + wa * outr[kk], p0.y + wa * outi[kk]);
            out[idx] = make_double2(p0.x + wt * outr[kk], p0.y + wt * outi[kk]);
        } else if (EPI == 2) {
            rk_acc[idx] = make_double2(e1[kk].x + wa * outr[kk], e1[kk].y + wa * outi[kk]);
            out[idx] = make_double2(e0[kk].x + wt * outr[kk], e0[kk].y + wt * outi[kk]);
        } else {
            const c128 v = make_double2(e1[kk].x + wa * outr[kk], e1[kk].y + wa * outi[kk]);
            out[idx] = v;
            nrm = fma(v.x, v.x, nrm);  // ‖ψ_new‖² partial of this thread (last RK stage): saves the separate norm pass
            nrm = fma(v.y, v.y, nrm);
        }
    }
}

// R rows per thread (r = tid + k·NT), S ring stages, RV: all slot values real, ND: real diagonal terms in registers
template <int R, int NT, int S, bool RV, int ND, int EPI, int NG, bool SG>
__global__ void __launch_bounds__(NT, 1)
    traj_xor_kernel(int nb, const __grid_constant__ XorDesc d, const c128* __restrict__ coef, long long coef_ld,
                    const c128* psi, c128* dpsi, double wa, double wt, const c128* __restrict__ rk_psi0,
                    c128* __restrict__ rk_acc, double* __restrict__ norm_part) {
    constexpr int n = R * NT, NW = NT / 32, CH = R < 8 ? R : 8, STAGE = n + kCoefPad;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    c128* ring = reinterpret_cast<c128*>(smem_raw);
    __shared__ __align__(8) unsigned long long bars[2 * S];
    __shared__ int s_mask[kXSlots + 4];
    __shared__ unsigned s_kt[kXSlots + 4];
    __shared__ double s_vre[kXSlots + 4], s_vim[kXSlots + 4], s_wre[kXSlots + 4], s_wim[kXSlots + 4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned bar_base = s_addr(bars);
    const unsigned psi_bytes = n * (unsigned)sizeof(c128);
    const unsigned coef_bytes = d.coef_shared ? 0u : (unsigned)(d.ncoef * sizeof(c128));

    double dreg[R][ND > 0 ? ND : 1];
#pragma unroll
    for (int k = 0; k < R; ++k)
#pragma unroll
        for (int t = 0; t < ND; ++t) dreg[k][t] = d.diag_r[t][tid + k * NT];
    if (tid < kXSlots + 4) {
        s_mask[tid] = d.mask[tid];
        s_kt[tid] = d.ktab[tid];
        s_vre[tid] = d.vre[tid];
        s_vim[tid] = d.vim[tid];
        s_wre[tid] = d.wre[tid];
        s_wim[tid] = d.wim[tid];
    }
    unsigned pmask = 0;  // bit j = parity(tid & zmask_j): this thread's share of the row parity of slot j
    if (SG) {
#pragma unroll 4
        for (int j = 0; j < kXSlots + 4; ++j) pmask |= ((unsigned)__popc(tid & d.zmask[j]) & 1u) << j;
    }
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            xb_init(bar_base + 8 * s, 1);         // full
            xb_init(bar_base + 8 * (S + s), NW);  // empty
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (d.coef_shared && tid < d.ncoef)  // one coefficient row for everybody: park it in every stage
        for (int s = 0; s < S; ++s) ring[(size_t)s * STAGE + n + tid] = coef[tid];
    __syncthreads();
    const int stride = gridDim.x;
    auto issue = [&](int slot, long long b) {
        const unsigned full = bar_base + 8 * slot;
        xb_expect(full, psi_bytes + coef_bytes);
        xb_bulk(s_addr(ring + (size_t)slot * STAGE), psi + (size_t)b * n, psi_bytes, full);
        if (coef_bytes) xb_bulk(s_addr(ring + (size_t)slot * STAGE + n), coef + b * coef_ld, coef_bytes, full);
    };
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < S - 1; ++i) {
            const long long b = blockIdx.x + (long long)i * stride;
            if (b < nb) issue(i, b);
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
                    issue(ls, lb);
                }
            }
            __syncwarp();
        }
        const int s = i % S;
        xb_wait(bar_base + 8 * s, (unsigned)((i / S) & 1));
        const c128* __restrict__ x = ring + (size_t)s * STAGE;
        const c128* cfs = x + n;
        c128 cdiag[ND > 0 ? ND : 1];
#pragma unroll
        for (int t = 0; t < ND; ++t) cdiag[t] = cfs[d.diag_coef[t]];
        c128 xo[R];
#pragma unroll
        for (int k = 0; k < R; ++k) xo[k] = x[tid + k * NT];
        c128* out = dpsi + (size_t)b * n;
        const c128* p0 = EPI == 2 ? rk_psi0 + (size_t)b * n : nullptr;
        c128* ac = EPI != 0 ? rk_acc + (size_t)b * n : nullptr;
        double nrm = 0.0;
        do_chunk<R, NT, CH, 0, RV, ND, EPI, NG, SG>(d, s_mask, s_kt, pmask, s_vre, s_vim, s_wre, s_wim, x, cfs, xo, dreg, cdiag, out, tid, wa, wt, p0, ac, nrm);
        if (R > 8)
            do_chunk<R, NT, CH, (R > 8 ? 8 : 0), RV, ND, EPI, NG, SG>(d, s_mask, s_kt, pmask, s_vre, s_vim, s_wre, s_wim, x, cfs, xo, dreg, cdiag, out, tid, wa, wt, p0, ac, nrm);
        if (EPI == 3 && norm_part) {  // per-warp partial of ‖ψ_b‖² in a fixed order: 16 slots per trajectory, unused ones zero
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
            if (lane == 0) norm_part[(size_t)b * 16 + warp] = nrm;
            if (warp == 0 && lane >= NW && lane < 16) norm_part[(size_t)b * 16 + lane] = 0.0;
        }
        __syncwarp();
        if (lane == 0) xb_arrive(bar_base + 8 * (S + s));
    }
}

inline void swap_slots(XorDesc& d, int a, int b) {
    std::swap(d.mask[a], d.mask[b]); std::swap(d.zmask[a], d.zmask[b]);
    std::swap(d.vre[a], d.vre[b]); std::swap(d.vim[a], d.vim[b]); std::swap(d.wre[a], d.wre[b]); std::swap(d.wim[a], d.wim[b]);
}

template <int R, int NT, int S, bool RV, int ND, int EPI, int NG = 0, bool SG = false>
int launch(Ctx* c, int nb, const XorDesc& d_in, const c128* coef, long long coef_ld, const c128* psi, c128* dpsi,
           const RkEpilogue* rk) {
    XorDesc d = d_in;
    int w = 0;
    for (int e = 0; e < d_in.n_x; ++e) {  // gathers first (padded to an even count), thread-local slots (mask % NT == 0) last
        d.x_begin[e] = w;
        int nrem = 0;
        for (int pass = 0; pass < 2; ++pass) {
            for (int j = d_in.x_begin[e]; j < d_in.x_begin[e + 1]; ++j) {
                const bool local = (d_in.mask[j] & (NT - 1)) == 0;
                if (local == (pass == 1)) {
                    if (w >= kXSlots + 4) return -1;
                    d.mask[w] = d_in.mask[j]; d.vre[w] = d_in.vre[j]; d.vim[w] = d_in.vim[j]; d.zmask[w] = d_in.zmask[j];
                    d.wre[w] = d_in.wre[j]; d.wim[w] = d_in.wim[j];
                    ++w;
                    if (!local) ++nrem;
                }
            }
            while (pass == 0 && (nrem % kSPR)) {  // dummy gathers with zero weight keep the rounds uniform
                if (w >= kXSlots + 4) return -1;
                d.mask[w] = 1; d.vre[w] = 0.0; d.vim[w] = 0.0; d.zmask[w] = 0; d.wre[w] = 0.0; d.wim[w] = 0.0;
                ++w; ++nrem;
            }
        }
        d.x_nrem[e] = nrem;
    }
    d.x_begin[d_in.n_x] = w;
    for (int j = w; j < kXSlots + 4; ++j) { d.mask[j] = 1; d.vre[j] = d.vim[j] = d.wre[j] = d.wim[j] = 0.0; d.zmask[j] = 0; }
    if constexpr (NG == 0 && RV && ND <= 1 && (!SG || EPI == 0)) {
        // One X-type term whose thread-local slots are exactly the single-bit flips NT·{1,2,4,…} (transverse-field
        // drivers): run the fully unrolled variant.  The local slots are put in ascending mask order first.
        constexpr int LOG2R = R == 16 ? 4 : (R == 8 ? 3 : (R == 4 ? 2 : 1));
        const bool no_fixed = !c->opt.xor_fixed;
        if (!no_fixed && d.n_x == 1) {
            const int ng = d.x_nrem[0], nl = d.x_begin[1] - ng;
            bool ok = nl == LOG2R && (ng == 8 || (R == 16 && (ng == 4 || ng == 12)));
            if (ok) {
                for (int a = ng; a < ng + nl; ++a)  // sort the local slots by mask
                    for (int b = a + 1; b < ng + nl; ++b)
                        if (d.mask[b] < d.mask[a]) swap_slots(d, a, b);
                for (int q = 0; q < LOG2R; ++q) ok = ok && d.mask[ng + q] == (NT << q);
            }
            if (ok) {
                if (ng == 8) return launch<R, NT, S, RV, ND, EPI, 8, SG>(c, nb, d_in, coef, coef_ld, psi, dpsi, rk);
                if constexpr (R == 16) {
                    if (ng == 4) return launch<R, NT, S, RV, ND, EPI, 4, SG>(c, nb, d_in, coef, coef_ld, psi, dpsi, rk);
                    return launch<R, NT, S, RV, ND, EPI, 12, SG>(c, nb, d_in, coef, coef_ld, psi, dpsi, rk);
                }
            }
        }
    }
    if constexpr (NG > 0) {  // same reordering as the detection above: local slots ascending
        const int ng = d.x_nrem[0], nl = d.x_begin[1] - ng;
        for (int a = ng; a < ng + nl; ++a)
            for (int b = a + 1; b < ng + nl; ++b)
                if (d.mask[b] < d.mask[a]) swap_slots(d, a, b);
    }
    for (int j = 0; j < kXSlots + 4; ++j) {  // after every reordering: the k-dependent half of each slot's row parity
        unsigned kt = 0;
        for (int k = 0; k < R; ++k) kt |= ((unsigned)__builtin_popcount((k * NT) & d.zmask[j]) & 1u) << k;
        d.ktab[j] = kt;
    }
    d.coef_shared = coef_ld == 0 ? 1 : 0;
    const size_t smem = (size_t)S * (R * NT + kCoefPad) * sizeof(c128);
    static DevOnce cfg;
    if (!cfg.done(c->device)) {
        OQB_CUDA(c, cudaFuncSetAttribute(traj_xor_kernel<R, NT, S, RV, ND, EPI, NG, SG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cfg.set(c->device);
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    int per_sm = (int)((224 * 1024) / (smem + 2048));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2048 / NT) per_sm = 2048 / NT;
    const int grid = nb < sms * per_sm ? nb : sms * per_sm;
    {
        // algorithmic bytes: read ψ + write dψ (+ the RK operands: 1 extra vector for modes 1/3, 3 for mode 2)
        ProfScope prof(c, (EPI == 0 ? 32.0 : (EPI == 2 ? 80.0 : 48.0)) * R * NT * (double)nb);
        traj_xor_kernel<R, NT, S, RV, ND, EPI, NG, SG><<<grid, NT, smem, c->stream>>>(nb, d, coef, coef_ld, psi, dpsi, rk ? rk->wa : 0.0,
                                                                             rk ? rk->wt : 0.0, rk ? rk->psi0 : nullptr,
                                                                             rk ? rk->acc : nullptr, rk ? rk->norm_part : nullptr);
    }
    OQB_LAUNCH_CHECK(c);
    return 0;
}

template <bool RV, int ND, int EPI>
int by_size_epi(Ctx* c, int n, int nb, const XorDesc& d, const c128* coef, long long coef_ld, const c128* psi, c128* dpsi,
                bool* used, const RkEpilogue* rk) {
    *used = true;
    int st = 0;
    if (d.any_signed) {
        // signed slots (general Pauli strings): the plain matvec only — the fused RK stages fall back to the unfused sequence
        if constexpr (EPI != 0) {
            *used = false;
            return 0;
        } else {
            switch (n) {
                case 4096: st = launch<16, 256, 3, RV, ND, 0, 0, true>(c, nb, d, coef, coef_ld, psi, dpsi, rk); break;
                case 2048: st = launch<8, 256, 3, RV, ND, 0, 0, true>(c, nb, d, coef, coef_ld, psi, dpsi, rk); break;
                case 1024: st = launch<4, 256, 3, RV, ND, 0, 0, true>(c, nb, d, coef, coef_ld, psi, dpsi, rk); break;
                case 512: st = launch<4, 128, 3, RV, ND, 0, 0, true>(c, nb, d, coef, coef_ld, psi, dpsi, rk); break;
                case 256: st = launch<2, 128, 3, RV, ND, 0, 0, true>(c, nb, d, coef, coef_ld, psi, dpsi, rk); break;
                default: *used = false; return 0;
            }
            if (st < 0) { *used = false; return 0; }
            return st;
        }
    }
    switch (n) {
        case 4096: {
            const int rsel = c->opt.xor_rows;
            if (rsel == 8) st = launch<8, 512, 3, RV, ND, EPI>(c, nb, d, coef, coef_ld, psi, dpsi, rk);
            else st = launch<16, 256, 3, RV, ND, EPI>(c, nb, d, coef, coef_ld, psi, dpsi, rk);
            break;
        }
        case 2048: st = launch<8, 256, 3, RV, ND, EPI>(c, nb, d, coef, coef_ld, psi, dpsi, rk); break;
        case 1024: st = launch<4, 256, 3, RV, ND, EPI>(c, nb, d, coef, coef_ld, psi, dpsi, rk); break;
        case 512: st = launch<4, 128, 3, RV, ND, EPI>(c, nb, d, coef, coef_ld, psi, dpsi, rk); break;
        case 256: st = launch<2, 128, 3, RV, ND, EPI>(c, nb, d, coef, coef_ld, psi, dpsi, rk); break;
        default: *used = false; return 0;
    }
    if (st < 0) {  // slot table overflow after padding: let the caller fall back
        *used = false;
        return 0;
    }
    return st;
}

template <bool RV, int ND>
int by_size(Ctx* c, int n, int nb, const XorDesc& d, const c128* coef, long long coef_ld, const c128* psi, c128* dpsi,
            bool* used, const RkEpilogue* rk) {
    switch (rk ? rk->mode : 0) {
        case 0: return by_size_epi<RV, ND, 0>(c, n, nb, d, coef, coef_ld, psi, dpsi, used, rk);
        case 1: return by_size_epi<RV, ND, 1>(c, n, nb, d, coef, coef_ld, psi, dpsi, used, rk);
        case 2: return by_size_epi<RV, ND, 2>(c, n, nb, d, coef, coef_ld, psi, dpsi, used, rk);
        default: return by_size_epi<RV, ND, 3>(c, n, nb, d, coef, coef_ld, psi, dpsi, used, rk);
    }
}

}  // namespace

int traj_matvec_xor(Ctx* c, int n, int nb, const std::vector<const TermDev*>& ts, const c128* d_coef, long long coef_ld,
                    const c128* psi, c128* dpsi, bool* used, const RkEpilogue* rk) {
    *used = false;
    if (c->opt.traj_kernel != 0) return 0;  // generic / no-XOR modes skip this path (A/B measurements)
    if ((int)ts.size() > kCoefPad) return 0;
    XorDesc d;
    memset(&d, 0, sizeof(d));
    d.ncoef = (int)ts.size();
    int slots = 0;
    bool all_real = true;
    for (size_t t = 0; t < ts.size(); ++t) {
        const TermDev* td = ts[t];
        if (td->is_diag) {
            if (td->diag_const) {  // c_bt must be 1 for every member for a constant term to fold into a scalar
                if (!td->const_coef_one) return 0;
                d.fconst_re += td->const_val.x;
                d.fconst_im += td->const_val.y;
                continue;
            }
            if (!(td->is_real && td->rvals) || d.n_diag >= kXDiag) return 0;
            d.diag_r[d.n_diag] = td->rvals;
            d.diag_coef[d.n_diag] = (int)t;
            d.n_diag++;
        } else {
            if (!td->xor_form || d.n_x >= kXTerms || slots + (int)td->xor_masks.size() > kXSlots) return 0;
            d.x_begin[d.n_x] = slots;
            // a term whose slot values are all purely imaginary (a Y-field driver) is stored as real numbers: 2 FMAs per gather
            // instead of 4, the factor i goes onto the term's coefficient inside the kernel
            bool term_imag = true, term_real = true;
            for (size_t j = 0; j < td->xor_masks.size(); ++j) {
                const c128 v = td->xor_vals[j], w = j < td->xor_vals1.size() ? td->xor_vals1[j] : v;
                term_imag = term_imag && v.x == 0.0 && w.x == 0.0;
                term_real = term_real && v.y == 0.0 && w.y == 0.0;
            }
            const bool rot = term_imag && !term_real;
            d.x_imag[d.n_x] = rot ? 1 : 0;
            if (rot) d.any_signed = 1;  // the rotation lives in the two-valued code path
            for (size_t j = 0; j < td->xor_masks.size(); ++j) {
                const c128 v = td->xor_vals[j], w = j < td->xor_vals1.size() ? td->xor_vals1[j] : v;
                d.mask[slots] = td->xor_masks[j];
                d.zmask[slots] = j < td->xor_zmasks.size() ? td->xor_zmasks[j] : 0;
                if (d.zmask[slots]) d.any_signed = 1;
                d.vre[slots] = rot ? v.y : v.x;  d.vim[slots] = rot ? 0.0 : v.y;
                d.wre[slots] = rot ? w.y : w.x;  d.wim[slots] = rot ? 0.0 : w.y;
                all_real = all_real && d.vim[slots] == 0.0 && d.wim[slots] == 0.0;
                ++slots;
            }
            d.x_coef[d.n_x] = (int)t;
            d.n_x++;
            d.x_begin[d.n_x] = slots;
        }
    }
    if (d.n_diag == 0) return all_real ? by_size<true, 0>(c, n, nb, d, d_coef, coef_ld, psi, dpsi, used, rk)
                                       : by_size<false, 0>(c, n, nb, d, d_coef, coef_ld, psi, dpsi, used, rk);
    if (d.n_diag == 1) return all_real ? by_size<true, 1>(c, n, nb, d, d_coef, coef_ld, psi, dpsi, used, rk)
                                       : by_size<false, 1>(c, n, nb, d, d_coef, coef_ld, psi, dpsi, used, rk);
    return all_real ? by_size<true, 2>(c, n, nb, d, d_coef, coef_ld, psi, dpsi, used, rk)
                    : by_size<false, 2>(c, n, nb, d, d_coef, coef_ld, psi, dpsi, used, rk);
}

}  // namespace oqb
