// Batched complex (ComplexF64) GEMM on the FP64 tensor pipe of sm_100a.
//
// This is the dense work-horse behind every ρ-RHS path (replaces the OpenBLAS zgemm calls at
// reference src/hamiltonian/dense_hamiltonian.jl:87-88, src/opensys/lindblad.jl:99,
// src/opensys/diffeq_liouvillian.jl:98,105, src/opensys/davies.jl:24,375, src/opensys/redfield.jl:66).
//
// Design:
//  * FP64 has no tcgen05/UMMA kind; the only FP64 tensor instruction on sm_100a is
//    mma.sync.m8n8k4 (SASS DMMA.8x8x4; ptxas lowers m16n8k* to sequences of it).
//  * One complex 8x8x4 product = 4 real DMMAs on interleaved operands:
//        re += ar*br ; re += (-ai)*bi ; im += ar*bi ; im += ai*br
//    conj(A)/conj(B) are sign-bit flips of the imaginary fragment (integer XOR, no FP64 op).
//  * Each lane fetches its complex fragment element with one 16-byte LDS; tiles are staged by
//    a STAGES-deep cp.async ring.  Shared layouts are padded so every quarter-warp LDS.128
//    touches all 32 banks exactly once (pitch ≡ 2 mod 8 for m/n-major, ≡ 4 mod 8 for k-major).
//  * Warp tile 32x32 complex (16 DMMA tiles x {re,im} = 128 accumulator registers),
//    CTA tile 128x64 (8 warps) or 64x64 (4 warps).
#include "common.cuh"

namespace oqb {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = pred ? 16 : 0;  // src-size 0 => 16 bytes of zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}
__device__ __forceinline__ double flip(double x, unsigned mask) {
    return __hiloint2double(__double2hiint(x) ^ (int)mask, __double2loint(x));
}

template <int BM_, int BN_, int BK_, int WM_, int WN_, int STAGES_, bool TA_, bool TB_>
struct Cfg {
    static constexpr int BM = BM_, BN = BN_, BK = BK_, WM = WM_, WN = WN_, STAGES = STAGES_;
    static constexpr bool TA = TA_, TB = TB_;
    static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
    static constexpr int NTHREADS = WARPS_M * WARPS_N * 32;
    static constexpr int PM = BM + 2, PN = BN + 2, PK = BK + 4;
    static constexpr int A_STAGE = TA ? BM * PK : BK * PM;  // c128 elements
    static constexpr int B_STAGE = TB ? BK * PN : BN * PK;
    static constexpr size_t SMEM = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(c128);
    static_assert(BM % 8 == 0 && BN % 8 == 0 && BK % 4 == 0, "tile shape");
    static_assert(PM % 8 == 2 && PN % 8 == 2 && PK % 8 == 4, "bank-conflict-free pitches");
};

template <class CF>
__global__ void __launch_bounds__(CF::NTHREADS) zgemm_kernel(ZgemmParams p, int z0) {
    constexpr int BM = CF::BM, BN = CF::BN, BK = CF::BK, WM = CF::WM, WN = CF::WN;
    constexpr int STAGES = CF::STAGES, NT = CF::NTHREADS;
    constexpr int PM = CF::PM, PN = CF::PN, PK = CF::PK;
    constexpr int MT = WM / 8, NTL = WN / 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c128* As = reinterpret_cast<c128*>(smem_raw);
    c128* Bs = As + (size_t)STAGES * CF::A_STAGE;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp % CF::WARPS_M) * WM;
    const int wn0 = (warp / CF::WARPS_M) * WN;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int z = blockIdx.z + z0;
    const int zb = z / p.batch2, zk = z - zb * p.batch2;

    const c128* __restrict__ A = p.A + zb * p.strideA + zk * p.strideA2;
    const c128* __restrict__ B = p.B + zb * p.strideB + zk * p.strideB2;
    const int M = p.M, N = p.N, K = p.K;
    const long long lda = p.lda, ldb = p.ldb;

    auto load_tiles = [&](int stage, int kt) {
        const int k0 = kt * BK;
        c128* as = As + (size_t)stage * CF::A_STAGE;
        c128* bs = Bs + (size_t)stage * CF::B_STAGE;
        if (!CF::TA) {
#pragma unroll
            for (int idx = tid; idx < BM * BK; idx += NT) {
                int m = idx % BM, k = idx / BM;
                bool ok = (m0 + m < M) && (k0 + k < K);
                const c128* src = ok ? A + (m0 + m) + (long long)(k0 + k) * lda : A;
                cp_async16(as + k * PM + m, src, ok);
            }
        } else {
#pragma unroll
            for (int idx = tid; idx < BM * BK; idx += NT) {
                int k = idx % BK, m = idx / BK;
                bool ok = (m0 + m < M) && (k0 + k < K);
                const c128* src = ok ? A + (k0 + k) + (long long)(m0 + m) * lda : A;
                cp_async16(as + m * PK + k, src, ok);
            }
        }
        if (!CF::TB) {
#pragma unroll
            for (int idx = tid; idx < BN * BK; idx += NT) {
                int k = idx % BK, n = idx / BK;
                bool ok = (n0 + n < N) && (k0 + k < K);
                const c128* src = ok ? B + (k0 + k) + (long long)(n0 + n) * ldb : B;
                cp_async16(bs + n * PK + k, src, ok);
            }
        } else {
#pragma unroll
            for (int idx = tid; idx < BN * BK; idx += NT) {
                int n = idx % BN, k = idx / BN;
                bool ok = (n0 + n < N) && (k0 + k < K);
                const c128* src = ok ? B + (n0 + n) + (long long)(k0 + k) * ldb : B;
                cp_async16(bs + k * PN + n, src, ok);
            }
        }
    };

    double acc_re[MT][NTL][2], acc_im[MT][NTL][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NTL; ++j) {
            acc_re[i][j][0] = acc_re[i][j][1] = 0.0;
            acc_im[i][j][0] = acc_im[i][j][1] = 0.0;
        }

    const unsigned sgnA = (p.opA == OP_C) ? 0x80000000u : 0u;
    const unsigned sgnB = (p.opB == OP_C) ? 0x80000000u : 0u;
    const int KT = (K + BK - 1) / BK;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_tiles(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + STAGES - 1;
            if (nk < KT) load_tiles(nk % STAGES, nk);
            cp_async_commit();
        }
        const c128* as = As + (size_t)(kt % STAGES) * CF::A_STAGE;
        const c128* bs = Bs + (size_t)(kt % STAGES) * CF::B_STAGE;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double ar[MT], ai[MT], nai[MT], br[NTL], bi[NTL];
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                c128 a = CF::TA ? as[(wm0 + i * 8 + g) * PK + kk + t] : as[(kk + t) * PM + wm0 + i * 8 + g];
                ar[i] = a.x;
                ai[i] = flip(a.y, sgnA);
                nai[i] = flip(ai[i], 0x80000000u);
            }
#pragma unroll
            for (int j = 0; j < NTL; ++j) {
                c128 b = CF::TB ? bs[(kk + t) * PN + wn0 + j * 8 + g] : bs[(wn0 + j * 8 + g) * PK + kk + t];
                br[j] = b.x;
                bi[j] = flip(b.y, sgnB);
            }
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NTL; ++j) {
                    dmma884(acc_re[i][j][0], acc_re[i][j][1], ar[i], br[j]);
                    dmma884(acc_im[i][j][0], acc_im[i][j][1], ar[i], bi[j]);
                }
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NTL; ++j) {
                    dmma884(acc_re[i][j][0], acc_re[i][j][1], nai[i], bi[j]);
                    dmma884(acc_im[i][j][0], acc_im[i][j][1], ai[i], br[j]);
                }
        }
    }
    cp_async_wait<0>();

    // ---- epilogue ----
    c128 alpha = p.alpha;
    if (p.alpha_scale) {
        double sc = p.alpha_scale[p.alpha_scale_mod > 0 ? z % p.alpha_scale_mod : z];
        alpha.x *= sc;
        alpha.y *= sc;
    }
    const c128 beta = p.beta;
    const bool has_beta = (beta.x != 0.0 || beta.y != 0.0);
    c128* __restrict__ C = p.C + zb * p.strideC + zk * p.strideC2;
    const c128* __restrict__ E = p.E ? p.E + zb * p.strideE : nullptr;
    c128* __restrict__ D = p.diag_out ? p.diag_out + zb * p.strideDiag : nullptr;
#pragma unroll
    for (int j = 0; j < NTL; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int col = n0 + wn0 + j * 8 + 2 * t + e;
            if (col >= N) continue;
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const int row = m0 + wm0 + i * 8 + g;
                if (row >= M) continue;
                double xr = acc_re[i][j][e], xi = acc_im[i][j][e];
                c128 v;
                v.x = alpha.x * xr - alpha.y * xi;
                v.y = alpha.x * xi + alpha.y * xr;
                if (D && row == col) D[row] = v;
                if (E) {
                    c128 f = E[row + (long long)col * p.lde];
                    double r2 = v.x * f.x - v.y * f.y;
                    double i2 = v.x * f.y + v.y * f.x;
                    v.x = r2;
                    v.y = i2;
                }
                c128* dst = C + row + (long long)col * p.ldc;
                if (has_beta) {
                    c128 o = *dst;
                    v.x += beta.x * o.x - beta.y * o.y;
                    v.y += beta.x * o.y + beta.y * o.x;
                }
                *dst = v;
            }
        }
}

template <class CF>
static int launch_cfg(Ctx* c, const ZgemmParams& p) {
    static DevOnce configured;
    if (!configured.done(c->device)) {
        OQB_CUDA(c, cudaFuncSetAttribute(zgemm_kernel<CF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)CF::SMEM));
        configured.set(c->device);
    }
    const long long nz = (long long)p.batch * p.batch2;
    dim3 block(CF::NTHREADS);
    for (long long z0 = 0; z0 < nz; z0 += 65535) {
        long long cnt = nz - z0 < 65535 ? nz - z0 : 65535;
        dim3 grid((p.M + CF::BM - 1) / CF::BM, (p.N + CF::BN - 1) / CF::BN, (unsigned)cnt);
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (c->profiling) {
            OQB_CUDA(c, cudaEventCreate(&e0));
            OQB_CUDA(c, cudaEventCreate(&e1));
            OQB_CUDA(c, cudaEventRecord(e0, c->stream));
        }
        zgemm_kernel<CF><<<grid, block, CF::SMEM, c->stream>>>(p, (int)z0);
        OQB_LAUNCH_CHECK(c);
        if (c->profiling) {
            OQB_CUDA(c, cudaEventRecord(e1, c->stream));
            c->prof_events.push_back(e0);
            c->prof_events.push_back(e1);
            c->prof_flops += 8.0 * p.M * p.N * p.K * (double)cnt;
        }
    }
    return 0;
}

template <bool TA, bool TB>
static int dispatch_tile(Ctx* c, const ZgemmParams& p) {
    if (p.M > 64)
        return launch_cfg<Cfg<128, 64, 8, 32, 32, 4, TA, TB>>(c, p);
    return launch_cfg<Cfg<64, 64, 8, 32, 32, 4, TA, TB>>(c, p);
}

int zgemm(Ctx* c, const ZgemmParams& p) {
    if (p.M <= 0 || p.N <= 0 || p.batch <= 0 || p.batch2 <= 0) return 0;
    if (p.K < 0) return set_error(c, 1, "zgemm: negative K");
    if (!p.A || !p.B || !p.C) return set_error(c, 1, "zgemm: null operand");
    {
        bool used = false;  // TMA-staged warp-specialised kernel first; cp.async kernel is the fallback
        OQB_TRY(zgemm_tma(c, p, &used));
        if (used) return 0;
    }
    const bool ta = p.opA != OP_N, tb = p.opB != OP_N;
    if (!ta && !tb) return dispatch_tile<false, false>(c, p);
    if (ta && !tb) return dispatch_tile<true, false>(c, p);
    if (!ta && tb) return dispatch_tile<false, true>(c, p);
    return dispatch_tile<true, true>(c, p);
}

}  // namespace oqb
