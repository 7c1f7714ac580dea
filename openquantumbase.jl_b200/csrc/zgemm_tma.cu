// Batched complex GEMM on the FP64 tensor pipe — TMA-staged, warp-specialised variant (sm_100a).
//
// Same arithmetic as zgemm.cu (4 real DMMA.8x8x4 per complex 8x8x4 product, conj by sign-bit XOR)
// but the operand tiles are brought in by the Tensor Memory Accelerator:
//   * a single elected lane issues cp.async.bulk.tensor for a whole tile (4-D tensor maps over
//     the interleaved complex matrices viewed as FP64 arrays: {2·inner, outer, batch2, batch}),
//     completion is signalled on per-stage `full` mbarriers (expect_tx = tile bytes); the producer
//     duty rotates among the warps (a 9th warp would cap registers at 168 per thread);
//   * all warps (32x32 complex warp tiles) wait on `full`, run 128 DMMAs per stage and release
//     the slot through the `empty` mbarrier — there is no CTA-wide barrier in the main loop and no
//     per-thread address arithmetic for the loads;
//   * tiles are 8 complex (= 128 bytes) deep in k, stored with the 128-byte TMA swizzle.  Fragment
//     rows are assigned to lanes through the permutation π(g) = (g>>1) + 4·(g&1) so that every
//     quarter-warp LDS.128 hits 8 distinct 16-byte bank groups:
//         chunk(lane) = π(g) ^ (kk + t)      for both k-major and m/n-major tiles.
//   * out-of-range rows/columns/k are zero-filled by TMA, so any M, N, K ≥ 1 is handled.
//   * default arithmetic is the 3-multiplication complex product (3 DMMAs per complex 8x8x4: T1 = ArBr,
//     T2 = AiBi, T3 = (Ar+Ai)(Br+Bi); re = T1 − T2, im = T3 − T1 − T2) with a software-pipelined fragment
//     fetch: 43.1 TFLOP/s algorithmic (8·M·N·K) on 1024³×64 vs 35.6 for the 4-multiplication kernel and a DMMA
//     issue peak of 37.1 — the FP64 tensor pipe executes 6·M·N·K flops.  Error vs numpy complex128 matmul is the
//     same 1.6e-15 (relative Frobenius) for both.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace oqb {

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(unsigned dst, const CUtensorMap* map, unsigned bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n" ::
            "r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ c128 lds128(unsigned addr) {
    c128 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void dmma884t(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}
__device__ __forceinline__ double flipt(double x, unsigned mask) {
    return __hiloint2double(__double2hiint(x) ^ (int)mask, __double2loint(x));
}

// M3_ selects the 3-multiplication complex product (T1 = ArBr, T2 = AiBi, T3 = (Ar+Ai)(Br+Bi);
// re = T1 − T2, im = T3 − T1 − T2): 3 DMMAs per complex 8x8x4 instead of 4, at the price of a third
// accumulator set — the warp tile shrinks to 32x16 complex (96 accumulator registers).
// RM_ (4-multiplication layout only): 1 = A is real, 2 = B is real — two DMMAs per complex 8x8x4
// (re += Ar·Br and im += Ar·Bi, resp. im += Ai·Br); the unused imaginary fragments are never touched.
template <int BM_, int BN_, int STAGES_, bool TA_, bool TB_, bool M3_ = false, int MINB_ = (BM_ == 64 ? 2 : 1), int RM_ = 0>
struct TCfg {
    static constexpr int BM = BM_, BN = BN_, STAGES = STAGES_;  // k depth of a stage: 8 complex = 128 bytes
    static constexpr bool TA = TA_, TB = TB_, M3 = M3_;
    static constexpr int RM = RM_;
    static_assert(!(M3_ && RM_ != 0), "the real-operand product uses the two-accumulator layout");
    static constexpr int WTN = M3 ? 16 : 32, NJ = WTN / 8, MINB = MINB_;
    static constexpr int WARPS_M = BM / 32, WARPS_N = BN / WTN, NCONS = WARPS_M * WARPS_N;
    static constexpr int NTHREADS = NCONS * 32;
    static constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 2 * STAGES * 8;
};

template <class CF>
__global__ void __launch_bounds__(CF::NTHREADS, CF::MINB)
    zgemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, ZgemmParams p,
                     int z0, int zmulA2, int zmulA, int zmulB2, int zmulB, int oneA, int oneB) {
    constexpr int BM = CF::BM, BN = CF::BN, STAGES = CF::STAGES;
    extern __shared__ unsigned char smem_raw[];
    const unsigned base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const unsigned bars = base + STAGES * CF::STAGE_BYTES;  // full[0..S), empty[0..S)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int z = blockIdx.z + z0;
    const int zb = z / p.batch2, zk = z - zb * p.batch2;
    const int KT = (p.K + 7) >> 3;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bars + s * 8, 1);
            mbar_init(bars + (STAGES + s) * 8, CF::NCONS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    // ------------------------------ TMA producer duty ------------------------------
    // A dedicated 9th warp would cap every thread at 168 registers (three warps on one SM
    // sub-partition), so the producer role rotates among the consumer warps instead: the elected
    // lane of warp (tile % NCONS) waits for the slot to drain and issues the tile's TMA boxes.
    const int a2 = zk * zmulA2, a3 = zb * zmulA, b2 = zk * zmulB2, b3 = zb * zmulB;
    auto produce = [&](int lt) {
        const int s = lt % STAGES;
        const unsigned ph = (lt / STAGES) & 1;
        const unsigned full = bars + s * 8, empty = bars + (STAGES + s) * 8;
        mbar_wait(empty, ph ^ 1u);
        mbar_expect_tx(full, CF::STAGE_BYTES);
        const unsigned as = base + s * CF::STAGE_BYTES, bs = as + CF::A_BYTES;
        const int k0 = lt * 8;
        if (CF::TA) {
            tma_load_5d(as, &mapA, full, 2 * k0, m0, 0, a2, a3);  // box {16, BM, 1}: rows m, 8 k each
        } else if (oneA) {
            tma_load_5d(as, &mapA, full, 0, k0, m0 >> 3, a2, a3);  // box {16, 8, BM/8}: whole tile, [mo][k][8 m]
        } else {
#pragma unroll 4
            for (int mo = 0; mo < BM / 8; ++mo)  // box {16, 8, 1}: slab mo = 8 k-rows of 8 m
                tma_load_5d(as + mo * 1024, &mapA, full, 2 * (m0 + 8 * mo), k0, 0, a2, a3);
        }
        if (!CF::TB) {
            tma_load_5d(bs, &mapB, full, 2 * k0, n0, 0, b2, b3);  // box {16, BN, 1}
        } else if (oneB) {
            tma_load_5d(bs, &mapB, full, 0, k0, n0 >> 3, b2, b3);
        } else {
#pragma unroll 4
            for (int no = 0; no < BN / 8; ++no)
                tma_load_5d(bs + no * 1024, &mapB, full, 2 * (n0 + 8 * no), k0, 0, b2, b3);
        }
    };
    // tiles kept in flight: STAGES-1, or STAGES-2 for the 3M loop (one slot of slack, so that the warp on
    // producer duty does not stall on a slot whose last reader is only a few DMMAs behind)
    constexpr int AHEAD = (CF::M3 || CF::RM != 0) ? STAGES - 2 : STAGES - 1;
    if (warp == 0 && lane == 0) {
        const int pre = KT < AHEAD ? KT : AHEAD;
        for (int lt = 0; lt < pre; ++lt) produce(lt);
    }

    // ---------------------------------- consumers ----------------------------------
    const int g = lane >> 2, t = lane & 3;
    const int pg = (g >> 1) + 4 * (g & 1);
    constexpr int NJ = CF::NJ;
    constexpr bool M3 = CF::M3;
    const int wm0 = (warp % CF::WARPS_M) * 32, wn0 = (warp / CF::WARPS_M) * CF::WTN;
    // byte offsets of this lane's fragment element inside a stage, for kk = 0 and kk = 4 (tile i/j adds 1024·i)
    unsigned aoff[2], boff[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k = 4 * h + t;
        const unsigned chunk = (unsigned)((pg ^ k) * 16);
        aoff[h] = (CF::TA ? (unsigned)((wm0 + pg) * 128) : (unsigned)((wm0 / 8) * 1024 + k * 128)) + chunk;
        boff[h] = (unsigned)CF::A_BYTES + (!CF::TB ? (unsigned)((wn0 + pg) * 128) : (unsigned)((wn0 / 8) * 1024 + k * 128)) + chunk;
    }

    // standard product: acc_re/acc_im; 3M product: acc_re = T1, acc_im = T2, acc_t3 = T3
    double acc_re[4][NJ][2], acc_im[4][NJ][2], acc_t3[M3 ? 4 : 1][NJ][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            acc_re[i][j][0] = acc_re[i][j][1] = 0.0;
            acc_im[i][j][0] = acc_im[i][j][1] = 0.0;
            if constexpr (M3) acc_t3[i][j][0] = acc_t3[i][j][1] = 0.0;
        }
    const unsigned sgnA = (p.opA == OP_C) ? 0x80000000u : 0u;
    const unsigned sgnB = (p.opB == OP_C) ? 0x80000000u : 0u;

    if constexpr (M3) {
        // 3M main loop, software-pipelined over half-stages (4 k each): the fragments of half-stage q+1 are
        // fetched (and their re+im sums formed) while the 24 DMMAs of half-stage q issue, so a warp never waits
        // on shared-memory latency between DMMA bursts even with two warps per SM sub-partition.
        struct Frag { double ar[4], ai[4], br[NJ], bi[NJ]; };
        auto fetch = [&](Frag& f, int q) {  // q = 2*kt + h; shared-memory loads and sign flips only
            const int kt = q >> 1, h = q & 1;
            const int s_ = kt % STAGES;
            if (h == 0) mbar_wait(bars + s_ * 8, (kt / STAGES) & 1);
            const unsigned st = base + s_ * CF::STAGE_BYTES;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                c128 a = lds128(st + aoff[h] + i * 1024);
                f.ar[i] = a.x;
                f.ai[i] = flipt(a.y, sgnA);
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                c128 b = lds128(st + boff[h] + j * 1024);
                f.br[j] = b.x;
                f.bi[j] = flipt(b.y, sgnB);
            }
        };
        // T1 first; the re+im sums (DADD, on values fetched a whole half-stage earlier) sit between the T1 and T2
        // bursts so that they never wait on shared-memory latency, then T2 and T3.
        auto mma = [&](const Frag& f) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884t(acc_re[i][j][0], acc_re[i][j][1], f.ar[i], f.br[j]);
            double as[4], bs[NJ];
#pragma unroll
            for (int i = 0; i < 4; ++i) as[i] = f.ar[i] + f.ai[i];
#pragma unroll
            for (int j = 0; j < NJ; ++j) bs[j] = f.br[j] + f.bi[j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884t(acc_im[i][j][0], acc_im[i][j][1], f.ai[i], f.bi[j]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884t(acc_t3[i][j][0], acc_t3[i][j][1], as[i], bs[j]);
        };
        auto duty = [&](int kt) {
            const int lt = kt + AHEAD;
            if (lt < KT && warp == lt % CF::NCONS) {
                if (lane == 0) produce(lt);
                __syncwarp();
            }
        };
        Frag f0, f1;
        duty(0);
        fetch(f0, 0);
        for (int kt = 0; kt < KT; ++kt) {
            fetch(f1, 2 * kt + 1);
            mma(f0);
            if (kt + 1 < KT) {
                duty(kt + 1);
                fetch(f0, 2 * kt + 2);
            }
            mma(f1);
            // the DMMAs above consumed every fragment register of stage kt, so its shared-memory reads are complete
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + (STAGES + kt % STAGES) * 8);
        }
    } else if constexpr (CF::RM != 0) {
        // real-operand main loop, software-pipelined like the 3M loop: 32 DMMAs per half-stage leave less time to hide
        // the fragment loads than the 64 of the 4-multiplication product, so the next half-stage's fragments are
        // fetched while the current one's DMMAs issue.
        struct Frag { double ar[4], ax[4], br[NJ], bx[NJ]; };  // ax = Ai (RM 2 only), bx = Bi (RM 1 only)
        auto fetch = [&](Frag& f, int q) {
            const int kt = q >> 1, h = q & 1;
            const int s_ = kt % STAGES;
            if (h == 0) mbar_wait(bars + s_ * 8, (kt / STAGES) & 1);
            const unsigned st = base + s_ * CF::STAGE_BYTES;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                c128 a = lds128(st + aoff[h] + i * 1024);
                f.ar[i] = a.x;
                f.ax[i] = flipt(a.y, sgnA);
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                c128 b = lds128(st + boff[h] + j * 1024);
                f.br[j] = b.x;
                f.bx[j] = flipt(b.y, sgnB);
            }
        };
        auto mma = [&](const Frag& f) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    dmma884t(acc_re[i][j][0], acc_re[i][j][1], f.ar[i], f.br[j]);
                    if constexpr (CF::RM == 2)
                        dmma884t(acc_im[i][j][0], acc_im[i][j][1], f.ax[i], f.br[j]);
                    else
                        dmma884t(acc_im[i][j][0], acc_im[i][j][1], f.ar[i], f.bx[j]);
                }
        };
        auto duty = [&](int kt) {
            const int lt = kt + AHEAD;
            if (lt < KT && warp == lt % CF::NCONS) {
                if (lane == 0) produce(lt);
                __syncwarp();
            }
        };
        Frag f0, f1;
        duty(0);
        fetch(f0, 0);
        for (int kt = 0; kt < KT; ++kt) {
            fetch(f1, 2 * kt + 1);
            mma(f0);
            if (kt + 1 < KT) {
                duty(kt + 1);
                fetch(f0, 2 * kt + 2);
            }
            mma(f1);
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + (STAGES + kt % STAGES) * 8);
        }
    } else {
        for (int kt = 0; kt < KT; ++kt) {
            {   // keep STAGES-1 tiles in flight
                const int lt = kt + AHEAD;
                if (lt < KT && warp == lt % CF::NCONS) {
                    if (lane == 0) produce(lt);
                    __syncwarp();
                }
            }
            const int s = kt % STAGES;
            const unsigned ph = (kt / STAGES) & 1;
            mbar_wait(bars + s * 8, ph);
            const unsigned st = base + s * CF::STAGE_BYTES;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double ar[4], ai[4], nai[4], br[NJ], bi[NJ];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    c128 a = lds128(st + aoff[h] + i * 1024);
                    ar[i] = a.x;
                    ai[i] = flipt(a.y, sgnA);
                    nai[i] = flipt(ai[i], 0x80000000u);
                }
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    c128 b = lds128(st + boff[h] + j * 1024);
                    br[j] = b.x;
                    bi[j] = flipt(b.y, sgnB);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        dmma884t(acc_re[i][j][0], acc_re[i][j][1], ar[i], br[j]);
                        if constexpr (CF::RM == 2)
                            dmma884t(acc_im[i][j][0], acc_im[i][j][1], ai[i], br[j]);
                        else
                            dmma884t(acc_im[i][j][0], acc_im[i][j][1], ar[i], bi[j]);
                    }
                if constexpr (CF::RM == 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            dmma884t(acc_re[i][j][0], acc_re[i][j][1], nai[i], bi[j]);
                            dmma884t(acc_im[i][j][0], acc_im[i][j][1], ai[i], br[j]);
                        }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + (STAGES + s) * 8);
        }
    }

    // ---- epilogue: C fragment (g, 2t+e) of tile (i,j) is matrix element (π(g), t + 4e) ----
    c128 alpha = p.alpha;
    if (p.alpha_scale) {
        const double sc = p.alpha_scale[p.alpha_scale_mod > 0 ? z % p.alpha_scale_mod : z];
        alpha.x *= sc;
        alpha.y *= sc;
    }
    const c128 beta = p.beta;
    const bool has_beta = (beta.x != 0.0 || beta.y != 0.0);
    c128* __restrict__ C = p.C + zb * p.strideC + zk * p.strideC2;
    const c128* __restrict__ E = p.E ? p.E + zb * p.strideE : nullptr;
    c128* __restrict__ D = p.diag_out ? p.diag_out + zb * p.strideDiag : nullptr;
    // Loads (old C for beta != 0, Hadamard factors) are issued for a whole j-slab (8 elements) before any
    // dependent arithmetic or store, so their L2 latencies overlap instead of serialising element by element.
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        c128 oldc[2][4], fac[2][4];
        bool ok[2][4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int col = n0 + wn0 + j * 8 + t + 4 * e;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = m0 + wm0 + i * 8 + pg;
                ok[e][i] = (col < p.N) && (row < p.M);
                oldc[e][i] = make_double2(0.0, 0.0);
                fac[e][i] = make_double2(1.0, 0.0);
                if (ok[e][i]) {
                    if (has_beta) oldc[e][i] = C[row + (long long)col * p.ldc];
                    if (E) fac[e][i] = E[row + (long long)col * p.lde];
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int col = n0 + wn0 + j * 8 + t + 4 * e;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (!ok[e][i]) continue;
                const int row = m0 + wm0 + i * 8 + pg;
                double xr = acc_re[i][j][e], xi = acc_im[i][j][e];
                if constexpr (M3) {
                    const double t1 = xr, t2 = xi;
                    xr = t1 - t2;
                    xi = (acc_t3[i][j][e] - t1) - t2;
                }
                c128 v;
                v.x = alpha.x * xr - alpha.y * xi;
                v.y = alpha.x * xi + alpha.y * xr;
                if (D && row == col) D[row] = v;
                if (E) {
                    const c128 f = fac[e][i];
                    const double r2 = v.x * f.x - v.y * f.y, i2 = v.x * f.y + v.y * f.x;
                    v.x = r2;
                    v.y = i2;
                }
                if (has_beta) {
                    const c128 o = oldc[e][i];
                    v.x += beta.x * o.x - beta.y * o.y;
                    v.y += beta.x * o.y + beta.y * o.x;
                }
                C[row + (long long)col * p.ldc] = v;
            }
        }
    }
}

// 5-D FP64 views of an interleaved complex operand (inner = the contiguous index, `outer` the strided one):
//   rows form (k-major tiles, or the per-slab fallback of m/n-major tiles):
//       dims {2·inner, outer, 1, batch2, batch}, box {16, box_outer, 1, 1, 1}
//   one-shot form for m/n-major tiles (inner % 8 == 0): the inner index is split into (8-element group,
//   group number) and the group number becomes a third, 128-byte-strided dimension, so that one TMA
//   instruction delivers the whole [group][k][8] tile:
//       dims {16, outer, inner/8, batch2, batch}, strides {ld·16, 128, …}, box {16, 8, tile/8, 1, 1}
bool make_map(CUtensorMap* map, const c128* ptr, long long inner, long long outer, long long ld, int batch2,
              long long stride2, int batch, long long stride, unsigned box_outer, bool one_shot, unsigned tile_groups,
              int* zmul2, int* zmul) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    const long long natural = ld * outer;  // used when a batch level is shared (extent 1, stride unused)
    *zmul2 = stride2 != 0 ? 1 : 0;
    *zmul = stride != 0 ? 1 : 0;
    const cuuint64_t nb2 = (cuuint64_t)(stride2 != 0 ? batch2 : 1), nb1 = (cuuint64_t)(stride != 0 ? batch : 1);
    const cuuint64_t sb2 = (cuuint64_t)(stride2 != 0 ? stride2 : natural) * 16;
    const cuuint64_t sb1 = (cuuint64_t)(stride != 0 ? stride : natural * (long long)nb2) * 16;
    cuuint64_t dims[5], strides[4];
    cuuint32_t box[5] = {16, box_outer, 1, 1, 1};
    if (!one_shot) {
        dims[0] = (cuuint64_t)(2 * inner); dims[1] = (cuuint64_t)outer; dims[2] = 1;
        strides[0] = (cuuint64_t)ld * 16; strides[1] = (cuuint64_t)natural * 16;
    } else {
        dims[0] = 16; dims[1] = (cuuint64_t)outer; dims[2] = (cuuint64_t)(inner / 8);
        strides[0] = (cuuint64_t)ld * 16; strides[1] = 128;
        box[1] = 8; box[2] = tile_groups;
    }
    dims[3] = nb2; dims[4] = nb1;
    strides[2] = sb2; strides[3] = sb1;
    for (int i = 0; i < 4; ++i)
        if (strides[i] == 0 || (strides[i] & 15) || strides[i] >= (1ull << 40)) return false;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, (void*)ptr, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <class CF>
int launch_tma(Ctx* c, const ZgemmParams& p, bool* used) {
    *used = false;
    alignas(64) CUtensorMap mapA, mapB;
    int zA2, zA, zB2, zB;
    // A: k-major when op != N (A stored K×M, k contiguous), else m-major
    const bool no_one_shot = !c->opt.tma_oneshot;
    int oneA = 0, oneB = 0;
    bool okA, okB;
    if (CF::TA) {
        okA = make_map(&mapA, p.A, p.K, p.M, p.lda, p.batch2, p.strideA2, p.batch, p.strideA, CF::BM, false, 0, &zA2, &zA);
    } else {
        okA = !no_one_shot && p.M % 8 == 0 &&
              make_map(&mapA, p.A, p.M, p.K, p.lda, p.batch2, p.strideA2, p.batch, p.strideA, 8, true, CF::BM / 8, &zA2, &zA);
        if (okA) oneA = 1;
        else okA = make_map(&mapA, p.A, p.M, p.K, p.lda, p.batch2, p.strideA2, p.batch, p.strideA, 8, false, 0, &zA2, &zA);
    }
    if (!CF::TB) {
        okB = make_map(&mapB, p.B, p.K, p.N, p.ldb, p.batch2, p.strideB2, p.batch, p.strideB, CF::BN, false, 0, &zB2, &zB);
    } else {
        okB = !no_one_shot && p.N % 8 == 0 &&
              make_map(&mapB, p.B, p.N, p.K, p.ldb, p.batch2, p.strideB2, p.batch, p.strideB, 8, true, CF::BN / 8, &zB2, &zB);
        if (okB) oneB = 1;
        else okB = make_map(&mapB, p.B, p.N, p.K, p.ldb, p.batch2, p.strideB2, p.batch, p.strideB, 8, false, 0, &zB2, &zB);
    }
    if (!okA || !okB) return 0;
    static DevOnce configured;
    if (!configured.done(c->device)) {
        OQB_CUDA(c, cudaFuncSetAttribute(zgemm_tma_kernel<CF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CF::SMEM));
        configured.set(c->device);
    }
    const long long nz = (long long)p.batch * p.batch2;
    for (long long z0 = 0; z0 < nz; z0 += 65535) {
        const long long cnt = nz - z0 < 65535 ? nz - z0 : 65535;
        dim3 grid((p.M + CF::BM - 1) / CF::BM, (p.N + CF::BN - 1) / CF::BN, (unsigned)cnt);
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (c->profiling) {
            OQB_CUDA(c, cudaEventCreate(&e0));
            OQB_CUDA(c, cudaEventCreate(&e1));
            OQB_CUDA(c, cudaEventRecord(e0, c->stream));
        }
        zgemm_tma_kernel<CF><<<grid, CF::NTHREADS, CF::SMEM, c->stream>>>(mapA, mapB, p, (int)z0, zA2, zA, zB2, zB, oneA, oneB);
        OQB_LAUNCH_CHECK(c);
        if (c->profiling) {
            OQB_CUDA(c, cudaEventRecord(e1, c->stream));
            c->prof_events.push_back(e0);
            c->prof_events.push_back(e1);
            c->prof_flops += 8.0 * p.M * p.N * p.K * (double)cnt;
        }
    }
    *used = true;
    return 0;
}

template <bool TA, bool TB>
int dispatch(Ctx* c, const ZgemmParams& p, bool* used) {
    // Long-K products amortise the prologue/epilogue of a 128x64 CTA (1 per SM); for short K two co-resident
    // 64x64 CTAs per SM overlap one CTA's epilogue with the other's main loop (measured: C2 +1.3 %, C5 +3.7 %).
    const int bm_force = c->opt.zgemm_bm;
    const bool big = bm_force ? bm_force == 128 : p.K > 512;
    // 3M product (Ctx::zgemm_3m, default on; oqb_set_zgemm_mode / OQB_ZGEMM_3M=0 select the 4-multiplication kernel):
    // 64x64 CTA of 8 warps (32x16 warp tiles), 8 stages, one CTA per SM.
    // real operand (eigenvectors of a real-symmetric Hamiltonian): half the DMMAs of the 4-multiplication product
    if (c->real_operand && p.real_operand == 1 && p.M > 64)
        return big ? launch_tma<TCfg<128, 64, 8, TA, TB, false, 1, 1>>(c, p, used) : launch_tma<TCfg<64, 64, 6, TA, TB, false, 2, 1>>(c, p, used);
    if (c->real_operand && p.real_operand == 2 && p.M > 64)
        return big ? launch_tma<TCfg<128, 64, 8, TA, TB, false, 1, 2>>(c, p, used) : launch_tma<TCfg<64, 64, 6, TA, TB, false, 2, 2>>(c, p, used);
    if (c->zgemm_3m) return launch_tma<TCfg<64, 64, 8, TA, TB, true, 1>>(c, p, used);
    if (p.M > 64 && big) return launch_tma<TCfg<128, 64, 6, TA, TB>>(c, p, used);
    return launch_tma<TCfg<64, 64, 6, TA, TB>>(c, p, used);
}

}  // namespace

// Returns 0 and sets *used when the TMA kernel handled the product; *used == false ⇒ caller falls back.
int zgemm_tma(Ctx* c, const ZgemmParams& p, bool* used) {
    *used = false;
    const bool disabled = !c->opt.zgemm_tma;
    if (disabled || p.K < 1 || p.M < 1 || p.N < 1) return 0;
    if (((uintptr_t)p.A & 15) || ((uintptr_t)p.B & 15)) return 0;
    const bool ta = p.opA != OP_N, tb = p.opB != OP_N;
    if (!ta && !tb) return dispatch<false, false>(c, p, used);
    if (ta && !tb) return dispatch<true, false>(c, p, used);
    if (!ta && tb) return dispatch<false, true>(c, p, used);
    return dispatch<true, true>(c, p, used);
}

}  // namespace oqb
