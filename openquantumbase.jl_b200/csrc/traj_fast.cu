// K7 fast path: ensemble matvec dψ_b = Σ_t c_bt M_t ψ_b with the operator resident in registers.
//
// The generic kernel (sparse.cu) re-reads every operator entry (value + column) from L2 for every
// trajectory: ≈1.1 MB of L2 traffic per 128 KB of HBM traffic at the C4 configuration, i.e. L2-bound
// at ≈24 % of the HBM roofline (profiles/r1_other_configs.md).  For Hamiltonians built from Pauli
// strings the structure allows much better:
//   * every ELL slot of an X-type term carries ONE value for all rows (slot_const), so only the column
//     index is per-row data — 16-bit, and each thread keeps the indices of the R rows it owns in
//     registers for the whole kernel (persistent CTAs);
//   * diagonal terms (Z-type terms, the combined −½Σγ_k L_k'L_k) are real vectors read through L1.
// Per trajectory the only memory traffic left is the algorithmic one: ψ in (16n bytes) and dψ out.
// ψ is staged with cp.async.bulk (UBLKCP, mbarrier complete_tx) into a double-buffered shared-memory
// ring so that the HBM read of trajectory i+1 overlaps the shared-memory gathers of trajectory i.
#include "sparse_terms.cuh"

namespace oqb {

namespace {

constexpr int kMaxSlots = 16;   // total ELL width over all slot-constant terms
constexpr int kMaxEll = 4;      // number of ELL terms
constexpr int kMaxDiag = 6;     // number of diagonal terms

struct FastDesc {
    int n_ell, n_diag, wtot;
    int ell_begin[kMaxEll + 1];       // slots [begin, end) of ELL term e
    int ell_coef[kMaxEll];            // index of its coefficient in the per-trajectory row
    const uint16_t* slot_cols[kMaxSlots];
    double slot_re[kMaxSlots], slot_im[kMaxSlots];
    const double* diag_r[kMaxDiag];   // real diagonal (nullptr ⇒ use diag_c)
    const c128* diag_c[kMaxDiag];
    int diag_coef[kMaxDiag];
};

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void bar_expect(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "W_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra W_DONE;\n\t"
        "bra W_LOOP;\n\t"
        "W_DONE:\n\t"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// R rows per thread, NT threads: n == R*NT.  WT = total number of ELL slots (compile-time so that the
// 16-bit column indices live in registers).
template <int R, int NT, int WT>
__global__ void __launch_bounds__(NT, 1)
    traj_fast_kernel(int nb, FastDesc fd, const c128* __restrict__ coef, long long coef_ld, const c128* __restrict__ psi,
                     c128* __restrict__ dpsi) {
    constexpr int n = R * NT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    c128* buf0 = reinterpret_cast<c128*>(smem_raw);
    c128* buf1 = buf0 + n;
    __shared__ __align__(8) unsigned long long bars[2];
    __shared__ c128 s_coef[kMaxEll + kMaxDiag + 2];
    const int tid = threadIdx.x;
    const unsigned bar0 = smem_addr(&bars[0]), bar1 = smem_addr(&bars[1]);

    // ---- operator rows owned by this thread: r = tid + k*NT ----
    unsigned colpk[R][(WT + 1) / 2];  // two 16-bit column indices per register
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int r = tid + k * NT;
#pragma unroll
        for (int j = 0; j < WT; j += 2) {
            unsigned lo = j < fd.wtot ? fd.slot_cols[j][r] : 0u;
            unsigned hi = (j + 1 < WT && j + 1 < fd.wtot) ? fd.slot_cols[j + 1][r] : 0u;
            colpk[k][j / 2] = lo | (hi << 16);
        }
    }
    if (tid == 0) {
        bar_init(bar0, 1);
        bar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const unsigned bytes = n * (unsigned)sizeof(c128);
    int b = blockIdx.x;
    if (tid == 0 && b < nb) {
        bar_expect(bar0, bytes);
        bulk_load(smem_addr(buf0), psi + (size_t)b * n, bytes, bar0);
    }
    const int ncoef = fd.n_ell + fd.n_diag;
    for (int it = 0; b < nb; b += gridDim.x, ++it) {
        const int cur = it & 1;
        const int nxt_b = b + gridDim.x;
        if (tid == 0 && nxt_b < nb) {  // the other buffer was released by the __syncthreads of the previous iteration
            const unsigned nb_bar = cur ? bar0 : bar1;
            bar_expect(nb_bar, bytes);
            bulk_load(smem_addr(cur ? buf0 : buf1), psi + (size_t)nxt_b * n, bytes, nb_bar);
        }
        if (tid < ncoef) s_coef[tid] = coef[b * coef_ld + tid];
        bar_wait(cur ? bar1 : bar0, (unsigned)((it >> 1) & 1));
        __syncthreads();  // s_coef visible
        const c128* __restrict__ x = cur ? buf1 : buf0;
        c128* __restrict__ out = dpsi + (size_t)b * n;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int r = tid + k * NT;
            double or_ = 0.0, oi_ = 0.0;
            // X-type terms: acc_e = Σ_j v_j ψ[col_j]; out += c_be · acc_e
            for (int e = 0; e < fd.n_ell; ++e) {
                double ar = 0.0, ai = 0.0;
#pragma unroll
                for (int j = 0; j < WT; ++j) {
                    if (j >= fd.ell_begin[e] && j < fd.ell_begin[e + 1]) {
                        const unsigned cidx = (colpk[k][j / 2] >> (16 * (j & 1))) & 0xffffu;
                        const c128 v = x[cidx];
                        const double sr = fd.slot_re[j], si = fd.slot_im[j];
                        ar = fma(sr, v.x, ar); ar = fma(-si, v.y, ar);
                        ai = fma(sr, v.y, ai); ai = fma(si, v.x, ai);
                    }
                }
                const c128 cc = s_coef[fd.ell_coef[e]];
                or_ = fma(cc.x, ar, or_); or_ = fma(-cc.y, ai, or_);
                oi_ = fma(cc.x, ai, oi_); oi_ = fma(cc.y, ar, oi_);
            }
            // diagonal terms: out += (Σ_t c_bt d_t[r]) ψ[r]
            double dr = 0.0, di = 0.0;
            for (int d = 0; d < fd.n_diag; ++d) {
                const c128 cc = s_coef[fd.diag_coef[d]];
                if (fd.diag_r[d]) {
                    const double v = __ldg(fd.diag_r[d] + r);
                    dr = fma(cc.x, v, dr);
                    di = fma(cc.y, v, di);
                } else {
                    const c128 v = __ldg(fd.diag_c[d] + r);
                    dr = fma(cc.x, v.x, dr); dr = fma(-cc.y, v.y, dr);
                    di = fma(cc.x, v.y, di); di = fma(cc.y, v.x, di);
                }
            }
            const c128 xr = x[r];
            or_ = fma(dr, xr.x, or_); or_ = fma(-di, xr.y, or_);
            oi_ = fma(dr, xr.y, oi_); oi_ = fma(di, xr.x, oi_);
            out[r] = make_double2(or_, oi_);
        }
        __syncthreads();  // everyone is done with buffer `cur` and s_coef
    }
}

template <int R, int NT, int WT>
int launch(Ctx* c, int nb, const FastDesc& fd, const c128* coef, long long coef_ld, const c128* psi, c128* dpsi) {
    const size_t smem = 2 * (size_t)R * NT * sizeof(c128);
    static DevOnce cfg;
    if (!cfg.done(c->device)) {
        OQB_CUDA(c, cudaFuncSetAttribute(traj_fast_kernel<R, NT, WT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cfg.set(c->device);
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    const int per_sm = (int)((220 * 1024) / (smem + 2048)) < 1 ? 1 : (int)((220 * 1024) / (smem + 2048));
    const int grid = nb < sms * per_sm ? nb : sms * per_sm;
    {
        ProfScope prof(c, 32.0 * R * NT * (double)nb);
        traj_fast_kernel<R, NT, WT><<<grid, NT, smem, c->stream>>>(nb, fd, coef, coef_ld, psi, dpsi);
    }
    OQB_LAUNCH_CHECK(c);
    return 0;
}

template <int WT>
int dispatch_n(Ctx* c, int n, int nb, const FastDesc& fd, const c128* coef, long long coef_ld, const c128* psi, c128* dpsi,
               bool* used) {
    *used = true;
    switch (n) {
        case 4096: return launch<8, 512, WT>(c, nb, fd, coef, coef_ld, psi, dpsi);
        case 2048: return launch<8, 256, WT>(c, nb, fd, coef, coef_ld, psi, dpsi);
        case 1024: return launch<4, 256, WT>(c, nb, fd, coef, coef_ld, psi, dpsi);
        case 512: return launch<4, 128, WT>(c, nb, fd, coef, coef_ld, psi, dpsi);
        case 256: return launch<2, 128, WT>(c, nb, fd, coef, coef_ld, psi, dpsi);
        default: *used = false; return 0;
    }
}

}  // namespace

int traj_matvec_fast(Ctx* c, int n, int nb, const std::vector<const TermDev*>& ts, const c128* d_coef, long long coef_ld,
                     const c128* psi, c128* dpsi, bool* used) {
    *used = false;
    const bool disabled = c->opt.traj_kernel == 1;  // generic kernels only
    if (disabled) return 0;
    FastDesc fd;
    fd.n_ell = fd.n_diag = fd.wtot = 0;
    fd.ell_begin[0] = 0;
    for (size_t t = 0; t < ts.size(); ++t) {
        const TermDev* td = ts[t];
        if (td->is_diag) {
            if (fd.n_diag >= kMaxDiag) return 0;
            fd.diag_r[fd.n_diag] = (td->is_real && td->rvals) ? td->rvals : nullptr;
            fd.diag_c[fd.n_diag] = td->vals;
            fd.diag_coef[fd.n_diag] = (int)t;
            fd.n_diag++;
        } else {
            if (!td->slot_const || !td->cols16 || fd.n_ell >= kMaxEll || fd.wtot + td->width > kMaxSlots) return 0;
            for (int j = 0; j < td->width; ++j) {
                fd.slot_cols[fd.wtot + j] = td->cols16 + (size_t)j * n;
                fd.slot_re[fd.wtot + j] = td->slot_vals[j].x;
                fd.slot_im[fd.wtot + j] = td->slot_vals[j].y;
            }
            fd.ell_coef[fd.n_ell] = (int)t;
            fd.wtot += td->width;
            fd.n_ell++;
            fd.ell_begin[fd.n_ell] = fd.wtot;
        }
    }
    for (int j = fd.wtot; j < kMaxSlots; ++j) { fd.slot_cols[j] = nullptr; fd.slot_re[j] = fd.slot_im[j] = 0.0; }
    if (fd.wtot <= 8) return dispatch_n<8>(c, n, nb, fd, d_coef, coef_ld, psi, dpsi, used);
    if (fd.wtot <= 12) return dispatch_n<12>(c, n, nb, fd, d_coef, coef_ld, psi, dpsi, used);
    return dispatch_n<16>(c, n, nb, fd, d_coef, coef_ld, psi, dpsi, used);
}

}  // namespace oqb
