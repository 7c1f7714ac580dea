// DenseHamiltonian + LindbladLiouvillian + Redfield-apply entry points (include/oqb.h).
// Everything dense is composed from the DMMA ZGEMM (zgemm.cu) and the streaming kernels of
// elementwise.cu; this file is host orchestration only.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/oqb.h"
#include "common.cuh"

using namespace oqb;

struct oqb_dense {
    oqb_ctx* ctx = nullptr;
    int n = 0, nterms = 0, K = 0;
    c128* d_mats = nullptr;  // (nterms + K) n×n : m_1..m_nterms, L_1'L_1..L_K'L_K
    c128* d_lops = nullptr;  // K n×n == the n × (K n) matrix [L_1 | … | L_K]
    c128* d_ldiag = nullptr; // K × n diagonals when EVERY L_k is diagonal (dephasing Z_k, …), else nullptr
    // every L_k MONOMIAL (at most one entry per row and per column: σ±, Pauli strings, their products — amplitude damping,
    // Pauli channels) but not all diagonal: L_k[a, π_k(a)] = v_k[a]; L_k'L_k is then diagonal with entries lld_k
    int32_t* d_mperm = nullptr;  // K × n: π_k(a), −1 for an empty row
    c128* d_mval = nullptr;      // K × n: v_k[a]
    double* d_mlld = nullptr;    // K × n: (L_k'L_k)[j, j]
    bool lops_real = false;   // every L_k has zero imaginary part (σx, σ±, σz, …): both sandwich products are real-operand products
    bool terms_real = false;  // every Hamiltonian term matrix has zero imaginary part: H(s)ρ and ρH(s) take the real-operand product
    // split form of H(s) = f_N(s)·m_N + Σ_{i diagonal} f_i(s)·diag(h_i): exactly one term has off-diagonal entries
    int offdiag_term = -1;       // its index, −1 when the split does not apply
    bool offdiag_real = false;   // … and it has no imaginary part
    c128* d_hdiag = nullptr;     // nterms × n: the diagonals h_i (zero row for m_N)
    // … and that term is a sum of Pauli strings stored densely (−ΣX drivers, XX catalysts, σy terms): its non-zeros sit on
    // at most 16 XOR diagonals, m_N[r, r ^ mask_j] = xval_j[r] — the commutator with it is a stencil on the hypercube
    int nxmask = 0;
    int xmask[16] = {0};
    int xconst[16] = {0};        // xval_j is one value for every row
    c128 xcval[16] = {};
    c128* d_xval = nullptr;      // nxmask × n
    bool xreal = false;          // every x_j is real
};

namespace {

const size_t kWorkspaceBudget = (size_t)3 << 30;  // bytes of scratch per batch chunk

// Complex coefficient rows for lincomb: row b = [hs * coefH[b][i] …, gs * gamma[b][k] …]
// Returns the number of rows (1 when everything is shared across the batch).
int build_coef(const oqb_dense* op, int nb, const double* coefH, int coef_shared, c128 hs, const double* gamma,
               int gamma_shared, c128 gs, bool with_h, std::vector<c128>& out) {
    const int nt = with_h ? op->nterms : 0, K = gamma ? op->K : 0;
    const bool shared = (!with_h || coef_shared) && (!gamma || gamma_shared);
    const int rows = shared ? 1 : nb;
    out.resize((size_t)rows * (nt + K));
    for (int b = 0; b < rows; ++b) {
        c128* r = out.data() + (size_t)b * (nt + K);
        for (int i = 0; i < nt; ++i) {
            double f = coefH[(coef_shared ? 0 : (size_t)b * nt) + i];
            r[i] = make_double2(hs.x * f, hs.y * f);
        }
        for (int k = 0; k < K; ++k) {
            double g = gamma[(gamma_shared ? 0 : (size_t)b * op->K) + k];
            r[nt + k] = make_double2(gs.x * g, gs.y * g);
        }
    }
    return rows;
}

// du[b] += Σ_k γ_bk L_k u_b L_k'   for members [0,cnt) of the given chunk; T is scratch cnt×K n×n
int sandwich(oqb_dense* op, int cnt, const double* d_gamma, int gamma_mod, const c128* u, c128* du, c128* T) {
    oqb_ctx* c = op->ctx;
    const long long n = op->n, nn = n * n;
    const int K = op->K;
    ZgemmParams p;  // T[b,k] = γ_bk L_k u_b
    p.M = p.N = p.K = (int)n;
    p.batch = cnt; p.batch2 = K;
    p.A = op->d_lops; p.lda = n; p.strideA = 0; p.strideA2 = nn;
    p.real_operand = op->lops_real ? 1 : 0;
    p.B = u; p.ldb = n; p.strideB = nn; p.strideB2 = 0;
    p.C = T; p.ldc = n; p.strideC = K * nn; p.strideC2 = nn;
    p.alpha_scale = d_gamma; p.alpha_scale_mod = gamma_mod;
    OQB_TRY(zgemm(c, p));
    ZgemmParams q;  // du[b] += [T_b1 | … | T_bK] · [L_1 | … | L_K]'
    q.M = q.N = (int)n; q.K = (int)(K * n);
    q.batch = cnt;
    q.A = T; q.lda = n; q.strideA = K * nn;
    q.B = op->d_lops; q.ldb = n; q.strideB = 0; q.opB = OP_C;
    q.real_operand = op->lops_real ? 2 : 0;
    q.C = du; q.ldc = n; q.strideC = nn;
    q.beta = make_double2(1.0, 0.0);
    return zgemm(c, q);
}

// Diagonal Lindblad operators L_k = diag(d_k): Σ_k γ_bk L_k u L_k' is the Hadamard product of u with
// G_b[a,c] = Σ_k γ_bk d_k[a] conj(d_k[c]) — one elementwise pass (48 B per element) instead of 2K GEMMs.
// grid (tiles of 256 elements, members); the K diagonals sit in shared memory.
// with_anti: also the anticommutator part −½(D_a + D_c)·u[a,c], D_a = Σ_k γ_bk |d_k[a]|² — used when H_eff's diagonal
// imaginary part is kept out of the GEMM operand so that the operand stays real.
__global__ void lindblad_diag_sandwich_kernel(int n, int K, const c128* __restrict__ ldiag, const double* __restrict__ gamma,
                                              int gamma_ld, const c128* __restrict__ u, c128* __restrict__ du, int with_anti) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c128* sd = reinterpret_cast<c128*>(smem_raw);  // K × n
    double* sg = reinterpret_cast<double*>(sd + (size_t)K * n);
    const int b = blockIdx.y;
    for (int e = threadIdx.x; e < K * n; e += blockDim.x) sd[e] = ldiag[e];
    if (threadIdx.x < K) sg[threadIdx.x] = gamma[(size_t)b * gamma_ld + threadIdx.x];
    __syncthreads();
    const long long nn = (long long)n * n;
    const c128* ub = u + (long long)b * nn;
    c128* dub = du + (long long)b * nn;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nn; e += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(e % n), cc = (int)(e / n);
        double gr = 0.0, gi = 0.0;
        for (int k = 0; k < K; ++k) {
            const c128 da = sd[k * n + a], dc = sd[k * n + cc];
            const double g = sg[k];
            // g · da · conj(dc)
            gr = fma(g, da.x * dc.x + da.y * dc.y, gr);
            gi = fma(g, da.y * dc.x - da.x * dc.y, gi);
            if (with_anti) gr = fma(-0.5 * g, (da.x * da.x + da.y * da.y) + (dc.x * dc.x + dc.y * dc.y), gr);
        }
        const c128 x = ub[e];
        c128 o = dub[e];
        o.x += gr * x.x - gi * x.y;
        o.y += gr * x.y + gi * x.x;
        dub[e] = o;
    }
}

// Shared rates: G[a,c] = Σ_k γ_k d_k[a] conj(d_k[c]) (− ½γ_k(|d_k[a]|² + |d_k[c]|²) with_anti) once per call; the sandwich
// Σ_k γ_k L_k u L_k' = G∘u is then a pure streaming pass (hadamard_acc_kernel: G stays in L2, no K-loop per element).
// Folding G∘u into the epilogue of the second commutator product instead was measured SLOWER (25.1 vs 22.8 ms per C2 step):
// two more operand arrays in the epilogue push every zgemm_tma variant into register spills.
__global__ void lindblad_diag_gmat_kernel(int n, int K, const c128* __restrict__ ldiag, const double* __restrict__ gamma, int with_anti,
                                          c128* __restrict__ G) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n * n) return;
    const int a = (int)(e % n), cc = (int)(e / n);
    double gr = 0.0, gi = 0.0;
    for (int k = 0; k < K; ++k) {  // same operation order as lindblad_diag_sandwich_kernel
        const c128 da = ldiag[(size_t)k * n + a], dc = ldiag[(size_t)k * n + cc];
        const double g = gamma[k];
        gr = fma(g, da.x * dc.x + da.y * dc.y, gr);
        gi = fma(g, da.y * dc.x - da.x * dc.y, gi);
        if (with_anti) gr = fma(-0.5 * g, (da.x * da.x + da.y * da.y) + (dc.x * dc.x + dc.y * dc.y), gr);
    }
    G[e] = make_double2(gr, gi);
}

// du[b] += G ∘ u[b]  (G shared by the batch): 2 elements per thread, grid (tiles, members)
__global__ void hadamard_acc_kernel(long long nn, const c128* __restrict__ G, const c128* __restrict__ u, c128* __restrict__ du) {
    const long long b = blockIdx.y;
    const c128* ub = u + b * nn;
    c128* dub = du + b * nn;
    for (long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2; e < nn; e += (long long)gridDim.x * blockDim.x * 2) {
        if (e + 1 < nn) {
            const c128 g0 = G[e], g1 = G[e + 1], x0 = ub[e], x1 = ub[e + 1];
            c128 o0 = dub[e], o1 = dub[e + 1];
            o0.x += g0.x * x0.x - g0.y * x0.y; o0.y += g0.x * x0.y + g0.y * x0.x;
            o1.x += g1.x * x1.x - g1.y * x1.y; o1.y += g1.x * x1.y + g1.y * x1.x;
            dub[e] = o0; dub[e + 1] = o1;
        } else {
            const c128 g0 = G[e], x0 = ub[e];
            c128 o0 = dub[e];
            o0.x += g0.x * x0.x - g0.y * x0.y; o0.y += g0.x * x0.y + g0.y * x0.x;
            dub[e] = o0;
        }
    }
}

// du[b][a,c] += (G[a,c] − i Σ_i f_bi (h_i[a] − h_i[c])) · u[b][a,c]: the shared Lindblad factor plus the commutator with the
// DIAGONAL Hamiltonian terms (−i(Hu − uH) for H = diag h is −i(h_a − h_c)u_ac, src/hamiltonian/dense_hamiltonian.jl:84-89);
// coef row b holds (f_bi, 0); the nt diagonals sit in shared memory (zero row for the off-diagonal term).
__global__ void hadamard_acc_split_kernel(int n, int nt, const c128* __restrict__ G, const c128* __restrict__ hdiag,
                                          const c128* __restrict__ coef, long long coef_ld, const c128* __restrict__ u,
                                          c128* __restrict__ du) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c128* sh = reinterpret_cast<c128*>(smem_raw);  // nt × n
    double* sf = reinterpret_cast<double*>(sh + (size_t)nt * n);
    const long long b = blockIdx.y, nn = (long long)n * n;
    for (int e = threadIdx.x; e < nt * n; e += blockDim.x) sh[e] = hdiag[e];
    if (threadIdx.x < nt) sf[threadIdx.x] = coef[b * coef_ld + threadIdx.x].x;
    __syncthreads();
    const c128* ub = u + b * nn;
    c128* dub = du + b * nn;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nn; e += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(e % n), cc = (int)(e / n);
        c128 g = G ? G[e] : make_double2(0.0, 0.0);
        for (int i = 0; i < nt; ++i) {
            const c128 ha = sh[i * n + a], hc = sh[i * n + cc];
            const double f = sf[i];
            // −i f (ha − hc) = f·((ha − hc).y, −(ha − hc).x)
            g.x = fma(f, ha.y - hc.y, g.x);
            g.y = fma(-f, ha.x - hc.x, g.y);
        }
        const c128 x = ub[e];
        c128 o = dub[e];
        o.x += g.x * x.x - g.y * x.y;
        o.y += g.x * x.y + g.y * x.x;
        dub[e] = o;
    }
}

// Monomial L_k: (L_k u L_k')[a,c] = v_k[a] conj(v_k[c]) u[π_k(a), π_k(c)] — a gather of u instead of two n³ products.
// One block per MONO_CB columns c, threads over the rows: everything that depends on the column only (π_k(c), γ_k conj(v_k[c]),
// −½Σ_k γ_k lld_k[c]) is staged in shared memory once per block, everything that depends on the row only (π_k(a), v_k[a], the
// row half of the anticommutator) is fetched once per thread and serves all MONO_CB columns; u_b is gathered K times from L2.
// with_anti: also −½ Σ_k γ_bk (lld_k[a] + lld_k[c]) u[a,c] (the anticommutator, when it was kept out of the commutator's operand).
constexpr int MONO_CB = 8;
__global__ void __launch_bounds__(256) lindblad_mono_sandwich_kernel(int n, int K, const int32_t* __restrict__ perm, const c128* __restrict__ val,
                                                                     const double* __restrict__ lld, const double* __restrict__ gamma, int gamma_ld,
                                                                     const c128* __restrict__ u, c128* __restrict__ du, int with_anti) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c128* s_w = reinterpret_cast<c128*>(smem_raw);                 // [K][CB]  γ_k conj(v_k[c])
    double* s_anti = reinterpret_cast<double*>(s_w + (size_t)K * MONO_CB);  // [CB]  −½ Σ_k γ_k lld_k[c]
    double* s_g = s_anti + MONO_CB;                                 // [K]
    int* s_pc = reinterpret_cast<int*>(s_g + K);                    // [K][CB]  π_k(c), −1 = empty column (or c ≥ n)
    const int c0 = blockIdx.x * MONO_CB;
    const long long b = blockIdx.y, nn = (long long)n * n;
    const c128* ub = u + b * nn;
    c128* dub = du + b * nn;
    const double* g = gamma + b * gamma_ld;
    for (int t = threadIdx.x; t < K * MONO_CB; t += blockDim.x) {
        const int k = t / MONO_CB, c = c0 + t % MONO_CB;
        int pc = -1;
        c128 w = make_double2(0.0, 0.0);
        if (c < n) {
            pc = perm[(size_t)k * n + c];
            const c128 vc = val[(size_t)k * n + c];
            w = make_double2(g[k] * vc.x, -g[k] * vc.y);
        }
        s_pc[t] = pc;
        s_w[t] = w;
    }
    for (int k = threadIdx.x; k < K; k += blockDim.x) s_g[k] = g[k];
    if (threadIdx.x < MONO_CB) {
        const int c = c0 + threadIdx.x;
        double acc = 0.0;
        if (with_anti && c < n)
            for (int k = 0; k < K; ++k) acc = fma(-0.5 * g[k], lld[(size_t)k * n + c], acc);
        s_anti[threadIdx.x] = acc;
    }
    __syncthreads();
    for (int a = threadIdx.x; a < n; a += blockDim.x) {
        double sr[MONO_CB], si[MONO_CB];
#pragma unroll
        for (int j = 0; j < MONO_CB; ++j) sr[j] = si[j] = 0.0;
        double anti_a = 0.0;
        for (int k = 0; k < K; ++k) {
            if (with_anti) anti_a = fma(-0.5 * s_g[k], lld[(size_t)k * n + a], anti_a);
            const int pa = perm[(size_t)k * n + a];
            if (pa < 0) continue;
            const c128 va = val[(size_t)k * n + a];
#pragma unroll
            for (int j = 0; j < MONO_CB; ++j) {
                const int pc = s_pc[k * MONO_CB + j];
                if (pc < 0) continue;
                const c128 wc = s_w[k * MONO_CB + j];
                const double wr = va.x * wc.x - va.y * wc.y, wi = va.x * wc.y + va.y * wc.x;  // γ v_a conj(v_c)
                const c128 x = ub[pa + (long long)pc * n];
                sr[j] = fma(wr, x.x, sr[j]); sr[j] = fma(-wi, x.y, sr[j]);
                si[j] = fma(wr, x.y, si[j]); si[j] = fma(wi, x.x, si[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < MONO_CB; ++j) {
            const int c = c0 + j;
            if (c >= n) break;
            const long long e = a + (long long)c * n;
            c128 o = dub[e];
            double r = sr[j], i = si[j];
            if (with_anti) {
                const c128 x = ub[e];
                const double f = anti_a + s_anti[j];
                r = fma(f, x.x, r);
                i = fma(f, x.y, i);
            }
            o.x += r;
            o.y += i;
            dub[e] = o;
        }
    }
}

// The whole split-form RHS in ONE pass when the off-diagonal Hamiltonian term is Pauli-structured (m_N[r, r^mask_j] = x_j[r]):
//   du[r,c] = −i f_N(s_b) Σ_j (x_j[r]·u[r^mask_j, c] − u[r, c^mask_j]·x_j[c^mask_j]) + (G[r,c] − i Σ_i f_bi (h_i[r] − h_i[c]))·u[r,c]
// — −i(Hu − uH) of src/hamiltonian/dense_hamiltonian.jl:84-89 with H = f_N m_N + Σ f_i diag(h_i), plus the diagonal-L_k
// dissipator as the factor G.  One CTA per TS×TS tile of one ρ (TS = 64): partners whose mask lies inside the tile come
// from shared memory, the others from global memory (coalesced: r^mask moves a warp's rows as a block); no product at all.
struct StencilDesc {
    int nmask, nt;
    int mask[16], isconst[16];
    double cre[16], cim[16];
};
__global__ void __launch_bounds__(256) lindblad_stencil_kernel(int n, int TS, const __grid_constant__ StencilDesc d,
                                                               const c128* __restrict__ xval, const double* __restrict__ fn, int fn_shared,
                                                               const c128* __restrict__ G, const c128* __restrict__ hdiag,
                                                               const c128* __restrict__ coef, long long coef_ld, const c128* __restrict__ u,
                                                               c128* __restrict__ du) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c128* tile = reinterpret_cast<c128*>(smem_raw);  // [TS cols][TS rows]
    __shared__ double sf[16];
    const long long b = blockIdx.z, nn = (long long)n * n;
    const int r0 = blockIdx.x * TS, c0 = blockIdx.y * TS;
    const c128* ub = u + b * nn;
    c128* dub = du + b * nn;
    if (threadIdx.x < d.nt) sf[threadIdx.x] = coef[b * coef_ld + threadIdx.x].x;
    // a thread keeps ONE row of the tile (TS ≤ 64 divides 256) and walks over columns: everything that depends on the row only
    // — its diagonal-term values, its non-constant stencil values — is fetched once
    const int rl = threadIdx.x & (TS - 1), cl0 = threadIdx.x / TS, cstep = blockDim.x / TS;
    const int r = r0 + rl;
    for (int cl = cl0; cl < TS; cl += cstep) tile[cl * TS + rl] = ub[r + (long long)(c0 + cl) * n];
    __syncthreads();
    const double f = fn[fn_shared ? 0 : b];
    // −i Σ_i f_bi h_i[r]  (row part of the diagonal terms' commutator) as one complex number per thread
    double hrr = 0.0, hri = 0.0;
    for (int i = 0; i < d.nt; ++i) {
        const c128 ha = hdiag[(size_t)i * n + r];
        hrr = fma(sf[i], ha.y, hrr);
        hri = fma(-sf[i], ha.x, hri);
    }
    for (int cl = cl0; cl < TS; cl += cstep) {
        const int c = c0 + cl;
        double ar = 0.0, ai = 0.0;
#pragma unroll 4
        for (int j = 0; j < d.nmask; ++j) {
            const int m = d.mask[j];
            c128 x1, x2;
            if (m < TS) {
                x1 = tile[cl * TS + (rl ^ m)];
                x2 = tile[(cl ^ m) * TS + rl];
            } else {
                x1 = ub[(r ^ m) + (long long)c * n];
                x2 = ub[r + (long long)(c ^ m) * n];
            }
            if (d.isconst[j]) {
                const double vr_ = d.cre[j], vi_ = d.cim[j];
                const double dx = x1.x - x2.x, dy = x1.y - x2.y;  // v·(x1 − x2)
                ar += vr_ * dx - vi_ * dy;
                ai += vr_ * dy + vi_ * dx;
            } else {
                const c128 vr = xval[(size_t)j * n + r], vc = xval[(size_t)j * n + (c ^ m)];
                ar += (vr.x * x1.x - vr.y * x1.y) - (x2.x * vc.x - x2.y * vc.y);
                ai += (vr.x * x1.y + vr.y * x1.x) - (x2.x * vc.y + x2.y * vc.x);
            }
        }
        // −i f (ar + i ai) = f (ai − i ar)
        double outr = f * ai, outi = -f * ar;
        c128 g = G ? G[(long long)r + (long long)c * n] : make_double2(0.0, 0.0);
        double hcr = 0.0, hci = 0.0;  // −i Σ_i f_bi h_i[c]
        for (int i = 0; i < d.nt; ++i) {
            const c128 hc = hdiag[(size_t)i * n + c];
            hcr = fma(sf[i], hc.y, hcr);
            hci = fma(-sf[i], hc.x, hci);
        }
        g.x += hrr - hcr;
        g.y += hri - hci;
        const c128 x0 = tile[cl * TS + rl];
        outr += g.x * x0.x - g.y * x0.y;
        outi += g.x * x0.y + g.y * x0.x;
        dub[(long long)r + (long long)c * n] = make_double2(outr, outi);
    }
}

// The same pass for n ≥ 64 with register micro-tiles (the version above issues ≈ 445 instructions per element and is issue-bound
// at 0.24 of the HBM bandwidth).  256 threads own a tile of 64 rows × 32 columns of one ρ; thread (a, q) keeps the 4 × 2
// elements (rows a + 16 i, columns q + 16 j) in registers, so that
//   – masks 16 / 32 / 48 on the row side and 16 on the column side pair elements of the SAME thread: no memory traffic at all,
//   – the other in-tile masks read shared memory with addresses that differ by compile-time offsets over (i, j):
//     2 LDS.128 + 4 (real values) or 8 DFMA per element and mask, ≈ 0.5 integer instructions,
//   – masks that leave the tile read the partner tile from global memory (L2),
//   – everything that depends on the row or the column only (stencil values, the diagonal terms' −i f h) is fetched once per
//     thread for its 4 rows / 2 columns.
// REALV: every stencil value is real (−ΣX, XX, X·Z products; not σy fields): half the multiplications.
template <bool REALV>
struct StencilAcc {
    double ar[4][2], ai[4][2];
    double vrr[4], vri[4], vcr[2], vci[2];  // value at the thread's rows (times u[r^m, c]) and at its partner columns (times u[r, c^m])
    __device__ __forceinline__ void row(int i, int j, const c128 x1) {  // += v[r_i] · u[r_i ^ m, c_j]
        if (REALV) {
            ar[i][j] = fma(vrr[i], x1.x, ar[i][j]);
            ai[i][j] = fma(vrr[i], x1.y, ai[i][j]);
        } else {
            ar[i][j] = fma(vrr[i], x1.x, fma(-vri[i], x1.y, ar[i][j]));
            ai[i][j] = fma(vrr[i], x1.y, fma(vri[i], x1.x, ai[i][j]));
        }
    }
    __device__ __forceinline__ void col(int i, int j, const c128 x2) {  // −= u[r_i, c_j ^ m] · v[c_j ^ m]
        if (REALV) {
            ar[i][j] = fma(-vcr[j], x2.x, ar[i][j]);
            ai[i][j] = fma(-vcr[j], x2.y, ai[i][j]);
        } else {
            ar[i][j] = fma(-vcr[j], x2.x, fma(vci[j], x2.y, ar[i][j]));
            ai[i][j] = fma(-vcr[j], x2.y, fma(-vci[j], x2.x, ai[i][j]));
        }
    }
    template <int MH>
    __device__ __forceinline__ void own_rows(const c128 (&x0)[4][2]) {  // row partners inside the thread's own elements
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) row(i, j, x0[i ^ MH][j]);
    }
};

template <bool REALV>
__global__ void __launch_bounds__(256, 2) lindblad_stencil64_kernel(int n, const __grid_constant__ StencilDesc d, const c128* __restrict__ xval,
                                                                    const double* __restrict__ fn, int fn_shared, const c128* __restrict__ G,
                                                                    const c128* __restrict__ hdiag, const c128* __restrict__ coef, long long coef_ld,
                                                                    const c128* __restrict__ u, c128* __restrict__ du) {
    __shared__ __align__(16) c128 tile[32 * 64];  // [32 cols][64 rows]
    const long long b = blockIdx.z, nn = (long long)n * n;
    const int a = threadIdx.x & 15, q = threadIdx.x >> 4;
    const int r0 = blockIdx.x * 64 + a, c0 = blockIdx.y * 32 + q;  // the thread's first row / column; the others are + 16 i / + 16 j
    const c128* ub = u + b * nn;
    const long long cs = 16LL * n;  // element distance of the thread's two columns
    c128 x0[4][2];
    {
        const c128* up = ub + r0 + (long long)c0 * n;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) x0[i][j] = up[16 * i + j * cs];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) tile[(q + 16 * j) * 64 + a + 16 * i] = x0[i][j];
    }
    __syncthreads();
    StencilAcc<REALV> A;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) A.ar[i][j] = A.ai[i][j] = 0.0;
    for (int jm = 0; jm < d.nmask; ++jm) {
        const int m = d.mask[jm];
        if (d.isconst[jm]) {
#pragma unroll
            for (int i = 0; i < 4; ++i) A.vrr[i] = d.cre[jm], A.vri[i] = d.cim[jm];
#pragma unroll
            for (int j = 0; j < 2; ++j) A.vcr[j] = d.cre[jm], A.vci[j] = d.cim[jm];
        } else {
            const c128* xv = xval + (size_t)jm * n;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const c128 v1 = xv[r0 + 16 * i];
                A.vrr[i] = v1.x, A.vri[i] = v1.y;
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const c128 v2 = xv[(c0 + 16 * j) ^ m];
                A.vcr[j] = v2.x, A.vci[j] = v2.y;
            }
        }
        const int ml = m & 15, mh = m >> 4;
        // row side: u[r_i ^ m, c_j]
        if (m >= 64) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const c128* t1 = ub + ((r0 + 16 * i) ^ m) + (long long)c0 * n;
#pragma unroll
                for (int j = 0; j < 2; ++j) A.row(i, j, t1[j * cs]);
            }
        } else if (ml == 0) {
            if (mh == 0) A.template own_rows<0>(x0);  // mask 0: the term's own diagonal
            else if (mh == 1) A.template own_rows<1>(x0);
            else if (mh == 2) A.template own_rows<2>(x0);
            else A.template own_rows<3>(x0);
        } else {
            const c128* p1 = tile + (a ^ ml) + q * 64;  // u[r_i ^ m, c_j] = p1[((i ^ mh) << 4) + j·1024]
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const c128* t1 = p1 + ((i ^ mh) << 4);
#pragma unroll
                for (int j = 0; j < 2; ++j) A.row(i, j, t1[j * 1024]);
            }
        }
        // column side: u[r_i, c_j ^ m]
        if (m >= 32) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const c128* t2 = ub + r0 + (long long)((c0 + 16 * j) ^ m) * n;
#pragma unroll
                for (int i = 0; i < 4; ++i) A.col(i, j, t2[16 * i]);
            }
        } else if (m == 16) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) A.col(i, j, x0[i][j ^ 1]);
        } else {
            const c128* p2 = tile + a + (q ^ ml) * 64;  // u[r_i, c_j ^ m] = p2[16 i + ((j ^ mh) << 10)]
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const c128* t2 = p2 + ((j ^ mh) << 10);
#pragma unroll
                for (int i = 0; i < 4; ++i) A.col(i, j, t2[16 * i]);
            }
        }
    }
    // −i Σ_t f_bt h_t[·] at the thread's rows and columns
    double hrr[4], hri[4], hcr[2], hci[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) hrr[i] = hri[i] = 0.0;
#pragma unroll
    for (int j = 0; j < 2; ++j) hcr[j] = hci[j] = 0.0;
    for (int t = 0; t < d.nt; ++t) {
        const double sf = coef[b * coef_ld + t].x;
        const c128* ht = hdiag + (size_t)t * n;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const c128 ha = ht[r0 + 16 * i];
            hrr[i] = fma(sf, ha.y, hrr[i]);
            hri[i] = fma(-sf, ha.x, hri[i]);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const c128 hc = ht[c0 + 16 * j];
            hcr[j] = fma(sf, hc.y, hcr[j]);
            hci[j] = fma(-sf, hc.x, hci[j]);
        }
    }
    const double f = fn[fn_shared ? 0 : b];
    c128* dp = du + b * nn + r0 + (long long)c0 * n;
    const c128* gp = G ? G + r0 + (long long)c0 * n : nullptr;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            c128 g = gp ? gp[16 * i + j * cs] : make_double2(0.0, 0.0);
            g.x += hrr[i] - hcr[j];
            g.y += hri[i] - hci[j];
            // −i f (ar + i ai) = f (ai − i ar)
            const double outr = fma(g.x, x0[i][j].x, fma(-g.y, x0[i][j].y, f * A.ai[i][j]));
            const double outi = fma(g.x, x0[i][j].y, fma(g.y, x0[i][j].x, -f * A.ar[i][j]));
            dp[16 * i + j * cs] = make_double2(outr, outi);
        }
}

// He[b][a,a] += Σ_k coef[b][k]·lld_k[a]   (monomial L_k: L_k'L_k is diagonal)
__global__ void heff_lld_add_kernel(int n, int K, const double* __restrict__ lld, const c128* __restrict__ coef, long long coef_ld,
                                    c128* __restrict__ He, long long strideHe) {
    const int b = blockIdx.y;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    double sr = 0.0, si = 0.0;
    for (int k = 0; k < K; ++k) {
        const double m2 = lld[(size_t)k * n + a];
        const c128 cf = coef[(size_t)b * coef_ld + k];
        sr = fma(cf.x, m2, sr);
        si = fma(cf.y, m2, si);
    }
    c128* h = He + (long long)b * strideHe + a + (long long)a * n;
    c128 v = *h;
    v.x += sr;
    v.y += si;
    *h = v;
}

// He[b][a,a] += Σ_k coef[b][k]·|d_k[a]|²  — the L_k'L_k part of H_eff for diagonal L_k (coef already carries −(i/2)γ_bk or −½γ_bk)
__global__ void heff_diag_add_kernel(int n, int K, const c128* __restrict__ ldiag, const c128* __restrict__ coef, long long coef_ld,
                                     c128* __restrict__ He, long long strideHe) {
    const int b = blockIdx.y;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    double sr = 0.0, si = 0.0;
    for (int k = 0; k < K; ++k) {
        const c128 d = ldiag[(size_t)k * n + a];
        const double m2 = d.x * d.x + d.y * d.y;
        const c128 cf = coef[(size_t)b * coef_ld + k];
        sr = fma(cf.x, m2, sr);
        si = fma(cf.y, m2, si);
    }
    c128* h = He + (long long)b * strideHe + a + (long long)a * n;
    c128 v = *h;
    v.x += sr;
    v.y += si;
    *h = v;
}

// shared implementation of oqb_dense_rhs (with_h) and oqb_lindblad_acc (!with_h)
int dense_rhs_impl(oqb_dense* op, int nb, const double* coefH, int coef_shared, const double* gamma,
                   int gamma_shared, const c128* u, c128* du, bool with_h) {
    oqb_ctx* c = op->ctx;
    if (nb <= 0) return 0;
    if (!u || !du) return set_error(c, OQB_EINVAL, "dense rhs: null state pointer");
    if (with_h && !coefH) return set_error(c, OQB_EINVAL, "dense rhs: null coefH");
    const long long n = op->n, nn = n * n;
    const int K = (gamma && op->K > 0) ? op->K : 0;
    if (!with_h && K == 0) return 0;
    if (K && !gamma_shared) {  // constant rates handed over per member (per-member s): one row serves everybody
        bool same = true;
        for (int b = 1; b < nb && same; ++b) same = memcmp(gamma, gamma + (size_t)b * K, (size_t)K * sizeof(double)) == 0;
        if (same) gamma_shared = 1;
    }

    // He_b = H(s_b) − (i/2) Σ γ_bk L_k'L_k    (or G_b = −½ Σ γ L'L for the accumulate form)
    std::vector<c128> coef;
    const c128 gs = with_h ? make_double2(0.0, -0.5) : make_double2(-0.5, 0.0);
    const int rows = build_coef(op, nb, coefH, coef_shared, make_double2(1.0, 0.0), K ? gamma : nullptr,
                                gamma_shared, gs, with_h, coef);
    const int nmat = (with_h ? op->nterms : 0) + K;
    const c128* mats = with_h ? op->d_mats : op->d_mats + (long long)op->nterms * nn;
    // lincomb needs [m…, LdL…] contiguous: true when with_h (terms then LdL) and when !with_h (LdL only)
    // one staged upload: [complex coefficient rows | raw γ rows]
    const size_t coef_bytes = coef.size() * sizeof(c128);
    const size_t gam_elems = K ? (gamma_shared ? (size_t)K : (size_t)nb * K) : 0;
    // split form (one off-diagonal Hamiltonian term): its coefficient f_N(s_b) per member, as the real scale of alpha
    // … and the plain commutator −i[H(s_b), u] (no Lindblad term: the H-d row itself, Redfield's Hamiltonian part) of a
    // Pauli-structured Hamiltonian needs no product either: the same stencil pass without the Hadamard factor
    const bool stencil_h = with_h && op->offdiag_term >= 0 && op->nxmask > 0 && op->nterms <= 16 && c->opt.lindblad_stencil;
    const bool h_only_stencil = stencil_h && K == 0;
    const bool may_split = (with_h && op->offdiag_term >= 0 && K > 0 && gamma_shared && c->opt.lindblad_split) || h_only_stencil;
    const size_t fn_elems = (may_split || stencil_h) ? (size_t)(coef_shared ? 1 : nb) : 0;
    std::vector<char> stage(coef_bytes + (gam_elems + fn_elems) * sizeof(double));
    memcpy(stage.data(), coef.data(), coef_bytes);
    if (gam_elems) memcpy(stage.data() + coef_bytes, gamma, gam_elems * sizeof(double));
    if (fn_elems) {
        double* fn = (double*)(stage.data() + coef_bytes + gam_elems * sizeof(double));
        for (size_t b = 0; b < fn_elems; ++b) fn[b] = coefH[b * op->nterms + op->offdiag_term];
    }
    void* d_stage = nullptr;
    OQB_TRY(coef_upload(c, stage.data(), stage.size(), &d_stage));
    void* d_coef = d_stage;
    void* d_gamma = (char*)d_stage + coef_bytes;
    const double* d_fn = (const double*)((char*)d_stage + coef_bytes + gam_elems * sizeof(double));

    const bool no_diag = !c->opt.lindblad_diag;
    const size_t diag_sm = (size_t)K * n * sizeof(c128) + (size_t)K * sizeof(double);
    // rates shared by the batch: the Hadamard factor is one n×n matrix built from global memory — no shared-memory limit;
    // per-member rates keep the K diagonals in shared memory (K·n ≤ 13 300: K = 12 at 10 qubits)
    const bool can_fuse = K > 0 && op->d_ldiag && !no_diag && gamma_shared && c->opt.lindblad_fuse;
    const bool use_diag = K > 0 && op->d_ldiag && !no_diag && (can_fuse || diag_sm <= 208 * 1024);
    const bool use_mono = K > 0 && !use_diag && op->d_mperm && c->opt.lindblad_mono;
    const int Kws = (use_diag || use_mono) ? 0 : K;  // the Hadamard / gather forms of the sandwich need no L_k u scratch
    const size_t per = (size_t)(1 + Kws) * nn * sizeof(c128);
    int chunk = (int)(kWorkspaceBudget / per);
    if (chunk < 1) chunk = 1;
    if (chunk > nb) chunk = nb;
    chunk = (nb + (nb + chunk - 1) / chunk - 1) / ((nb + chunk - 1) / chunk);  // balance the chunks
    const size_t he_elems = (size_t)(rows == 1 ? 1 : chunk) * nn;
    // shared rates + diagonal L_k: the sandwich is one n×n Hadamard factor for every member, folded into the second product
    const bool fuse_sandwich = can_fuse;
    OQB_TRY(ws_reserve(c, (he_elems + (size_t)chunk * Kws * nn + (fuse_sandwich ? nn : 0)) * sizeof(c128)));
    c128* He = (c128*)c->ws;
    c128* T = He + he_elems;
    c128* Gmat = T + (size_t)chunk * Kws * nn;

    // −i[H(s_b), u_b] (+ G∘u_b) of members [m_first, m_first + cnt) as one stencil pass (Pauli-structured off-diagonal term), OVERWRITES du
    auto stencil_pass = [&](const c128* G2, int m_first, int cnt) -> int {
        StencilDesc sd;
        memset(&sd, 0, sizeof(sd));
        sd.nmask = op->nxmask;
        sd.nt = op->nterms;
        for (int j = 0; j < op->nxmask; ++j) {
            sd.mask[j] = op->xmask[j];
            sd.isconst[j] = op->xconst[j];
            sd.cre[j] = op->xcval[j].x;
            sd.cim[j] = op->xcval[j].y;
        }
        const int TS = n >= 64 ? 64 : (int)n;
        static DevOnce cfg;
        if (!cfg.done(c->device)) {
            OQB_CUDA(c, cudaFuncSetAttribute(lindblad_stencil_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            cfg.set(c->device);
        }
        const bool micro = n >= 64 && c->opt.lindblad_stencil >= 2;  // 2 (default): register micro-tiles; 1: the plain one-element-at-a-time kernel
        for (int m0 = m_first; m0 < m_first + cnt; m0 += 65535) {
            const int mc = m_first + cnt - m0 < 65535 ? m_first + cnt - m0 : 65535;
            dim3 grid((unsigned)(n / TS), (unsigned)(n / TS), mc), grid64((unsigned)(n / 64), (unsigned)(n / 32), mc);
            const double* fnp = d_fn + (coef_shared ? 0 : m0);
            const c128* cfp = (const c128*)d_coef + (rows == 1 ? 0 : (long long)m0 * nmat);
            const long long cld = rows == 1 ? 0 : nmat;
            if (micro && op->xreal)
                lindblad_stencil64_kernel<true><<<grid64, 256, 0, c->stream>>>((int)n, sd, op->d_xval, fnp, coef_shared ? 1 : 0, G2, op->d_hdiag, cfp, cld,
                                                                               u + (long long)m0 * nn, du + (long long)m0 * nn);
            else if (micro)
                lindblad_stencil64_kernel<false><<<grid64, 256, 0, c->stream>>>((int)n, sd, op->d_xval, fnp, coef_shared ? 1 : 0, G2, op->d_hdiag, cfp, cld,
                                                                                u + (long long)m0 * nn, du + (long long)m0 * nn);
            else
                lindblad_stencil_kernel<<<grid, 256, (size_t)TS * TS * sizeof(c128), c->stream>>>((int)n, TS, sd, op->d_xval, fnp, coef_shared ? 1 : 0, G2, op->d_hdiag,
                                                                                                  cfp, cld, u + (long long)m0 * nn, du + (long long)m0 * nn);
            OQB_LAUNCH_CHECK(c);
        }
        return 0;
    };

    // Split form: H(s_b) = f_N(s_b)·m_N + diagonal terms, diagonal L_k, shared rates (config C2: −ΣX driver, ZZ problem term,
    // dephasing).  No per-member H_eff is assembled at all: both products use the SHARED operand m_N with alpha scaled by
    // f_N(s_b), and the diagonal terms' commutator joins the Lindblad factor in the one Hadamard pass.
    const size_t split_sm = (size_t)op->nterms * n * sizeof(c128) + (size_t)op->nterms * sizeof(double);
    if (may_split && (fuse_sandwich || h_only_stencil) && split_sm <= 160 * 1024) {
        c128* G2 = nullptr;
        if (K) {
            G2 = (c128*)c->ws;  // the workspace was reserved for at least nn elements (he_elems ≥ nn)
            lindblad_diag_gmat_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, c->stream>>>((int)n, K, op->d_ldiag, (const double*)d_gamma, 1, G2);
            OQB_LAUNCH_CHECK(c);
        }
        if (stencil_h) {
            OQB_TRY(stencil_pass(G2, 0, nb));
            return 0;
        }
        const c128* mN = op->d_mats + (long long)op->offdiag_term * nn;
        const bool rN = op->offdiag_real && c->real_operand;
        ZgemmParams p;  // du = −i f_N(s_b) m_N u
        p.M = p.N = p.K = (int)n; p.batch = nb;
        p.A = mN; p.lda = n; p.strideA = 0;
        p.real_operand = rN ? 1 : 0;
        p.B = u; p.ldb = n; p.strideB = nn;
        p.C = du; p.ldc = n; p.strideC = nn;
        p.alpha = make_double2(0.0, -1.0);
        p.alpha_scale = d_fn; p.alpha_scale_mod = coef_shared ? 1 : 0;
        OQB_TRY(zgemm(c, p));
        ZgemmParams q;  // du += +i f_N(s_b) u m_N'
        q.M = q.N = q.K = (int)n; q.batch = nb;
        q.A = u; q.lda = n; q.strideA = nn;
        q.B = mN; q.ldb = n; q.strideB = 0; q.opB = OP_C;
        q.real_operand = rN ? 2 : 0;
        q.C = du; q.ldc = n; q.strideC = nn;
        q.alpha = make_double2(0.0, 1.0);
        q.alpha_scale = d_fn; q.alpha_scale_mod = coef_shared ? 1 : 0;
        q.beta = make_double2(1.0, 0.0);
        OQB_TRY(zgemm(c, q));
        static DevOnce cfg;
        if (!cfg.done(c->device)) { OQB_CUDA(c, cudaFuncSetAttribute(hadamard_acc_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); cfg.set(c->device); }
        for (int m0 = 0; m0 < nb; m0 += 65535) {
            const int mc = nb - m0 < 65535 ? nb - m0 : 65535;
            dim3 grid((unsigned)std::min<long long>((nn + 255) / 256, 64), mc);
            hadamard_acc_split_kernel<<<grid, 256, split_sm, c->stream>>>((int)n, op->nterms, G2, op->d_hdiag,
                                                                          (const c128*)d_coef + (rows == 1 ? 0 : (long long)m0 * nmat), rows == 1 ? 0 : nmat,
                                                                          u + (long long)m0 * nn, du + (long long)m0 * nn);
            OQB_LAUNCH_CHECK(c);
        }
        return 0;
    }

    // diagonal L_k: L_k'L_k = diag|d_k|² — assemble H_eff from the Hamiltonian terms only and add the diagonal afterwards
    // (the per-member lincomb then reads nterms matrices instead of nterms + K)
    const bool diag_heff = (use_diag || use_mono) && with_h;
    const int nlin = diag_heff ? op->nterms : nmat;
    // real Hamiltonian terms: keep the (imaginary) L'L diagonal out of the operand — it moves into the Hadamard pass —
    // and run both products with the real-operand kernel; without Lindblad operators the operand is H(s) itself
    const bool real_h = with_h && op->terms_real && c->real_operand && (diag_heff || K == 0);
    auto assemble = [&](const c128* cf, long long cf_ld, int cnt) -> int {
        OQB_TRY(lincomb(c, nlin, nn, mats, cf, cf_ld, He, cnt, false));
        if (diag_heff && !real_h) {
            for (int m0 = 0; m0 < cnt; m0 += 65535) {
                const int mc = cnt - m0 < 65535 ? cnt - m0 : 65535;
                dim3 grid((unsigned)((n + 255) / 256), mc);
                if (use_mono)
                    heff_lld_add_kernel<<<grid, 256, 0, c->stream>>>((int)n, K, op->d_mlld, cf + op->nterms + (long long)m0 * cf_ld, cf_ld,
                                                                     He + (long long)m0 * nn, nn);
                else
                    heff_diag_add_kernel<<<grid, 256, 0, c->stream>>>((int)n, K, op->d_ldiag, cf + op->nterms + (long long)m0 * cf_ld, cf_ld,
                                                                      He + (long long)m0 * nn, nn);
                OQB_LAUNCH_CHECK(c);
            }
        }
        return 0;
    };
    // Pauli-structured Hamiltonian + diagonal or monomial L_k (per-member rates, amplitude damping, Pauli channels): the
    // commutator is the stencil pass, the dissipator — anticommutator included — the Hadamard / gather pass: no product at all
    const bool stencil_comm = stencil_h && (use_diag || use_mono) && !fuse_sandwich;
    if (rows == 1 && !stencil_comm) OQB_TRY(assemble((const c128*)d_coef, 0, 1));
    if (fuse_sandwich) {
        lindblad_diag_gmat_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, c->stream>>>((int)n, K, op->d_ldiag, (const double*)d_gamma,
                                                                                       real_h ? 1 : 0, Gmat);
        OQB_LAUNCH_CHECK(c);
    }
    for (int b0 = 0; b0 < nb; b0 += chunk) {
        const int cnt = nb - b0 < chunk ? nb - b0 : chunk;
        const c128* ub = u + (long long)b0 * nn;
        c128* dub = du + (long long)b0 * nn;
        if (stencil_comm) OQB_TRY(stencil_pass(nullptr, b0, cnt));
        else {
        if (rows != 1) OQB_TRY(assemble((const c128*)d_coef + (long long)b0 * nmat, nmat, cnt));
        ZgemmParams p;  // du = (−i) He u      [accumulate form: du += G u]
        p.M = p.N = p.K = (int)n; p.batch = cnt;
        p.A = He; p.lda = n; p.strideA = rows == 1 ? 0 : nn;
        p.real_operand = real_h ? 1 : 0;
        p.B = ub; p.ldb = n; p.strideB = nn;
        p.C = dub; p.ldc = n; p.strideC = nn;
        p.alpha = with_h ? make_double2(0.0, -1.0) : make_double2(1.0, 0.0);
        p.beta = with_h ? make_double2(0.0, 0.0) : make_double2(1.0, 0.0);
        OQB_TRY(zgemm(c, p));
        ZgemmParams q;  // du += (+i) u He'    [accumulate form: du += u G]
        q.M = q.N = q.K = (int)n; q.batch = cnt;
        q.A = ub; q.lda = n; q.strideA = nn;
        q.B = He; q.ldb = n; q.strideB = rows == 1 ? 0 : nn; q.opB = with_h ? OP_C : OP_N;
        q.real_operand = real_h ? 2 : 0;
        q.C = dub; q.ldc = n; q.strideC = nn;
        q.alpha = with_h ? make_double2(0.0, 1.0) : make_double2(1.0, 0.0);
        q.beta = make_double2(1.0, 0.0);
        OQB_TRY(zgemm(c, q));
        }
        const int anti = (real_h || stencil_comm) ? 1 : 0;  // the anticommutator −½{Σγ L'L, u} was kept out of the commutator's operand
        if (fuse_sandwich) {
            for (int m0 = 0; m0 < cnt; m0 += 65535) {
                const int mc = cnt - m0 < 65535 ? cnt - m0 : 65535;
                dim3 grid((unsigned)std::min<long long>((nn / 2 + 255) / 256, 64), mc);
                hadamard_acc_kernel<<<grid, 256, 0, c->stream>>>(nn, Gmat, ub + (long long)m0 * nn, dub + (long long)m0 * nn);
                OQB_LAUNCH_CHECK(c);
            }
        } else if (K) {
            const double* dg = (const double*)d_gamma + (gamma_shared ? 0 : (long long)b0 * K);
            const size_t sm = diag_sm;
            if (use_diag) {
                static DevOnce cfg;
                if (!cfg.done(c->device)) { OQB_CUDA(c, cudaFuncSetAttribute(lindblad_diag_sandwich_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024)); cfg.set(c->device); }
                for (int m0 = 0; m0 < cnt; m0 += 65535) {
                    const int mc = cnt - m0 < 65535 ? cnt - m0 : 65535;
                    const int tiles = (int)std::min<long long>((nn + 255) / 256, 64);
                    dim3 grid(tiles, mc);
                    lindblad_diag_sandwich_kernel<<<grid, 256, sm, c->stream>>>(
                        (int)n, K, op->d_ldiag, dg + (gamma_shared ? 0 : (long long)m0 * K), gamma_shared ? 0 : K,
                        ub + (long long)m0 * nn, dub + (long long)m0 * nn, anti);
                    OQB_LAUNCH_CHECK(c);
                }
            } else if (use_mono) {
                for (int m0 = 0; m0 < cnt; m0 += 65535) {
                    const int mc = cnt - m0 < 65535 ? cnt - m0 : 65535;
                    dim3 grid((unsigned)((n + MONO_CB - 1) / MONO_CB), mc);
                    const size_t msm = (size_t)K * MONO_CB * (sizeof(c128) + sizeof(int)) + (size_t)(MONO_CB + K) * sizeof(double);
                    lindblad_mono_sandwich_kernel<<<grid, 256, msm, c->stream>>>(
                        (int)n, K, op->d_mperm, op->d_mval, op->d_mlld, dg + (gamma_shared ? 0 : (long long)m0 * K), gamma_shared ? 0 : K,
                        ub + (long long)m0 * nn, dub + (long long)m0 * nn, anti);
                    OQB_LAUNCH_CHECK(c);
                }
            } else {
                OQB_TRY(sandwich(op, cnt, dg, gamma_shared ? K : 0, ub, dub, T));
            }
        }
    }
    return 0;
}

}  // namespace

// Host-only analysis of the Hamiltonian terms (no GPU needed; also behind oqb_dense_structure_probe):
//   – exactly one term with off-diagonal entries among ≥ 2 terms (or a single Pauli-structured term) → the split form
//     H(s) = f_N(s)·m_N + Σ f_i(s)·diag(h_i);
//   – … and that term's non-zeros on ≤ 16 XOR diagonals (n a power of two): m_N[r, r ^ mask_j] = x_j[r] — a sum of Pauli strings.
struct TermStructure {
    int offdiag_term = -1;
    bool offdiag_real = false;
    std::vector<c128> hdiag;    // nterms × n diagonals (zero row for the off-diagonal term)
    std::vector<int> masks;     // sorted XOR masks, empty when the term has no such structure
    std::vector<int> isconst;   // x_j is one value for every row
    std::vector<c128> xval;     // masks.size() × n
    bool xreal = false;         // every x_j is real
};
static void analyze_terms(const c128* Hm, int n, int nterms, TermStructure& ts) {
    const size_t nn = (size_t)n * n;
    int noff = 0, last = -1;
    ts.hdiag.assign((size_t)std::max(nterms, 1) * n, make_double2(0.0, 0.0));
    for (int i = 0; i < nterms; ++i) {
        bool diag = true;
        for (int j = 0; j < n && diag; ++j) {
            const c128* col = Hm + i * nn + (size_t)j * n;
            for (int r = 0; r < n; ++r)
                if (r != j && (col[r].x != 0.0 || col[r].y != 0.0)) { diag = false; break; }
        }
        if (diag) {
            for (int r = 0; r < n; ++r) ts.hdiag[(size_t)i * n + r] = Hm[i * nn + (size_t)r * n + r];
        } else {
            ++noff;
            last = i;
        }
    }
    if (noff != 1) return;
    // a single-term Hamiltonian (ConstantDenseHamiltonian, src/hamiltonian/dense_hamiltonian.jl:149-168) has nothing to split;
    // it is recorded only when it has Pauli structure (its diagonal part is XOR diagonal 0)
    const bool single = nterms < 2;
    bool r2 = true;
    for (size_t e = 0; r2 && e < nn; ++e) r2 = Hm[last * nn + e].y == 0.0;
    if (!single) {
        ts.offdiag_term = last;
        ts.offdiag_real = r2;
    }
    if ((n & (n - 1)) != 0 || n < 2) return;
    std::vector<int> masks;
    for (int j = 0; j < n; ++j)
        for (int r = 0; r < n; ++r) {
            const c128 v = Hm[last * nn + (size_t)j * n + r];
            if (v.x == 0.0 && v.y == 0.0) continue;
            const int m = r ^ j;
            if (std::find(masks.begin(), masks.end(), m) == masks.end()) {
                if (masks.size() >= 16) return;  // more than 16 XOR diagonals: no stencil
                masks.push_back(m);
            }
        }
    if (masks.empty()) return;
    std::sort(masks.begin(), masks.end());
    ts.xval.resize(masks.size() * (size_t)n);
    ts.isconst.resize(masks.size());
    ts.xreal = true;
    for (size_t q = 0; q < masks.size(); ++q) {
        bool cst = true;
        for (int r = 0; r < n; ++r) {
            const c128 v = Hm[last * nn + (size_t)(r ^ masks[q]) * n + r];  // element (row r, col r ^ mask)
            ts.xval[q * n + r] = v;
            if (v.x != ts.xval[q * n].x || v.y != ts.xval[q * n].y) cst = false;
            if (v.y != 0.0) ts.xreal = false;
        }
        ts.isconst[q] = cst ? 1 : 0;
    }
    ts.masks = masks;
    ts.offdiag_term = last;
    ts.offdiag_real = r2;
}

extern "C" {

int oqb_dense_structure_probe(int n, int nterms, const void* h_mats, int* offdiag_term, int* nmask, int* masks16, int* isconst16,
                               int* all_real, void* xval) {
    if (n <= 0 || nterms <= 0 || !h_mats) return OQB_EINVAL;
    TermStructure ts;
    analyze_terms((const c128*)h_mats, n, nterms, ts);
    if (offdiag_term) *offdiag_term = ts.offdiag_term;
    if (nmask) *nmask = (int)ts.masks.size();
    for (size_t q = 0; q < ts.masks.size(); ++q) {
        if (masks16) masks16[q] = ts.masks[q];
        if (isconst16) isconst16[q] = ts.isconst[q];
    }
    if (all_real) *all_real = ts.xreal ? 1 : 0;
    if (xval && !ts.xval.empty()) memcpy(xval, ts.xval.data(), ts.xval.size() * sizeof(c128));
    return 0;
}

int oqb_dense_create(oqb_ctx* c, int n, int nterms, const void* h_mats, int K, const void* h_lops, oqb_dense** out) {
    OQB_DEVICE(c, c);
    if (!c || !out) return OQB_EINVAL;
    *out = nullptr;
    if (n <= 0 || nterms < 0 || K < 0 || (nterms > 0 && !h_mats) || (K > 0 && !h_lops))
        return set_error(c, OQB_EINVAL, "oqb_dense_create: bad arguments");
    oqb_dense* op = new oqb_dense();
    op->ctx = c; op->n = n; op->nterms = nterms; op->K = K;
    const size_t nn = (size_t)n * n;
    cudaError_t e = cudaMalloc((void**)&op->d_mats, (nterms + K + 1) * nn * sizeof(c128));
    if (e == cudaSuccess) e = cudaMalloc((void**)&op->d_lops, (K + 1) * nn * sizeof(c128));
    if (e != cudaSuccess) { oqb_dense_destroy(op); return set_error(c, OQB_ENOMEM, "oqb_dense_create: %s", cudaGetErrorString(e)); }
    if (nterms) OQB_CUDA(c, cudaMemcpyAsync(op->d_mats, h_mats, nterms * nn * sizeof(c128), cudaMemcpyHostToDevice, c->stream));
    {
        const c128* Hm = (const c128*)h_mats;
        bool re = nterms > 0;
        for (size_t e = 0; re && e < (size_t)nterms * nn; ++e) re = Hm[e].y == 0.0;
        op->terms_real = re;
        // which terms are diagonal matrices (σz / σzσz problem Hamiltonians stored densely)?  Does the remaining one have Pauli structure?
        TermStructure ts;
        analyze_terms(Hm, n, nterms, ts);
        if (ts.offdiag_term >= 0) {
            op->offdiag_term = ts.offdiag_term;
            op->offdiag_real = ts.offdiag_real;
            if (cudaMalloc((void**)&op->d_hdiag, ts.hdiag.size() * sizeof(c128)) != cudaSuccess) { oqb_dense_destroy(op); return set_error(c, OQB_ENOMEM, "oqb_dense_create: cudaMalloc failed"); }
            OQB_CUDA(c, cudaMemcpyAsync(op->d_hdiag, ts.hdiag.data(), ts.hdiag.size() * sizeof(c128), cudaMemcpyHostToDevice, c->stream));
            OQB_CUDA(c, cudaStreamSynchronize(c->stream));
            if (!ts.masks.empty()) {
                for (size_t q = 0; q < ts.masks.size(); ++q) {
                    op->xmask[q] = ts.masks[q];
                    op->xconst[q] = ts.isconst[q];
                    op->xcval[q] = ts.xval[q * n];
                }
                if (cudaMalloc((void**)&op->d_xval, ts.xval.size() * sizeof(c128)) != cudaSuccess) { oqb_dense_destroy(op); return set_error(c, OQB_ENOMEM, "oqb_dense_create: cudaMalloc failed"); }
                OQB_CUDA(c, cudaMemcpyAsync(op->d_xval, ts.xval.data(), ts.xval.size() * sizeof(c128), cudaMemcpyHostToDevice, c->stream));
                OQB_CUDA(c, cudaStreamSynchronize(c->stream));
                op->nxmask = (int)ts.masks.size();
                op->xreal = ts.xreal;
            }
        }
    }
    if (K) {
        OQB_CUDA(c, cudaMemcpyAsync(op->d_lops, h_lops, K * nn * sizeof(c128), cudaMemcpyHostToDevice, c->stream));
        ZgemmParams p;  // L_k' L_k   (src/opensys/lindblad.jl:99,107)
        p.M = p.N = p.K = n; p.batch = K;
        p.opA = OP_C;
        p.A = op->d_lops; p.lda = n; p.strideA = (long long)nn;
        p.B = op->d_lops; p.ldb = n; p.strideB = (long long)nn;
        p.C = op->d_mats + (size_t)nterms * nn; p.ldc = n; p.strideC = (long long)nn;
        int s = zgemm(c, p);
        if (s) { oqb_dense_destroy(op); return s; }
        // every L_k diagonal (dephasing operators stored as dense matrices)? then keep the K diagonals for the
        // Hadamard form of the sandwich term
        const c128* L = (const c128*)h_lops;
        {
            bool re = true;
            for (size_t e = 0; re && e < (size_t)K * nn; ++e) re = L[e].y == 0.0;
            op->lops_real = re;
        }
        bool diag = true;
        for (long long k = 0; k < K && diag; ++k)
            for (size_t e = 0; e < nn && diag; ++e)
                if (e % n != e / n && (L[k * nn + e].x != 0.0 || L[k * nn + e].y != 0.0)) diag = false;
        if (diag) {
            std::vector<c128> d((size_t)K * n);
            for (long long k = 0; k < K; ++k)
                for (int r = 0; r < n; ++r) d[k * n + r] = L[k * nn + (size_t)r * n + r];
            if (cudaMalloc((void**)&op->d_ldiag, d.size() * sizeof(c128)) != cudaSuccess) { oqb_dense_destroy(op); return set_error(c, OQB_ENOMEM, "oqb_dense_create: cudaMalloc failed"); }
            OQB_CUDA(c, cudaMemcpyAsync(op->d_ldiag, d.data(), d.size() * sizeof(c128), cudaMemcpyHostToDevice, c->stream));
            OQB_CUDA(c, cudaStreamSynchronize(c->stream));
        } else {
            // monomial operators: at most one entry per row and per column
            std::vector<int32_t> perm((size_t)K * n, -1);
            std::vector<c128> val((size_t)K * n, make_double2(0.0, 0.0));
            std::vector<double> lld((size_t)K * n, 0.0);
            bool mono = true;
            for (long long k = 0; k < K && mono; ++k) {
                std::vector<char> col_used(n, 0);
                for (int j = 0; j < n && mono; ++j)          // column j of the column-major block
                    for (int r = 0; r < n && mono; ++r) {
                        const c128 v = L[k * nn + (size_t)j * n + r];
                        if (v.x == 0.0 && v.y == 0.0) continue;
                        if (perm[k * n + r] >= 0 || col_used[j]) { mono = false; break; }
                        perm[k * n + r] = j;
                        val[k * n + r] = v;
                        col_used[j] = 1;
                        lld[k * n + j] = v.x * v.x + v.y * v.y;
                    }
            }
            if (mono) {
                if (cudaMalloc((void**)&op->d_mperm, perm.size() * sizeof(int32_t)) != cudaSuccess ||
                    cudaMalloc((void**)&op->d_mval, val.size() * sizeof(c128)) != cudaSuccess ||
                    cudaMalloc((void**)&op->d_mlld, lld.size() * sizeof(double)) != cudaSuccess) {
                    oqb_dense_destroy(op);
                    return set_error(c, OQB_ENOMEM, "oqb_dense_create: cudaMalloc failed");
                }
                OQB_CUDA(c, cudaMemcpyAsync(op->d_mperm, perm.data(), perm.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
                OQB_CUDA(c, cudaMemcpyAsync(op->d_mval, val.data(), val.size() * sizeof(c128), cudaMemcpyHostToDevice, c->stream));
                OQB_CUDA(c, cudaMemcpyAsync(op->d_mlld, lld.data(), lld.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
                OQB_CUDA(c, cudaStreamSynchronize(c->stream));
            }
        }
    }
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    *out = op;
    return OQB_OK;
}

int oqb_dense_destroy(oqb_dense* op) {
    OQB_DEVICE(op, op->ctx);
    if (!op) return OQB_OK;
    cudaStreamSynchronize(op->ctx->stream);
    if (op->d_mats) cudaFree(op->d_mats);
    if (op->d_lops) cudaFree(op->d_lops);
    if (op->d_ldiag) cudaFree(op->d_ldiag);
    if (op->d_xval) cudaFree(op->d_xval);
    if (op->d_mperm) cudaFree(op->d_mperm);
    if (op->d_mval) cudaFree(op->d_mval);
    if (op->d_mlld) cudaFree(op->d_mlld);
    if (op->d_hdiag) cudaFree(op->d_hdiag);
    delete op;
    return OQB_OK;
}

oqb_ctx* oqb_dense_ctx(const oqb_dense* op) { return op ? op->ctx : nullptr; }

int oqb_dense_info(const oqb_dense* op, int* n, int* nterms, int* K, int* terms_real) {
    if (!op) return OQB_EINVAL;
    if (n) *n = op->n;
    if (nterms) *nterms = op->nterms;
    if (K) *K = op->K;
    if (terms_real) *terms_real = op->terms_real ? 1 : 0;
    return 0;
}

int oqb_dense_hamiltonian(oqb_dense* op, int nb, const double* h_coefH, int coef_shared, void* d_out) {
    OQB_DEVICE(op, op->ctx);
    oqb_ctx* c = op->ctx;
    if (nb <= 0) return 0;
    if (!h_coefH || !d_out) return set_error(c, OQB_EINVAL, "oqb_dense_hamiltonian: null argument");
    std::vector<c128> coef;
    const int rows = build_coef(op, nb, h_coefH, coef_shared, make_double2(1.0, 0.0), nullptr, 1,
                                make_double2(0.0, 0.0), true, coef);
    void* d_coef = nullptr;
    OQB_TRY(coef_upload(c, coef.data(), coef.size() * sizeof(c128), &d_coef));
    return lincomb(c, op->nterms, (long long)op->n * op->n, op->d_mats, (const c128*)d_coef,
                   rows == 1 ? 0 : op->nterms, (c128*)d_out, nb, false);
}

int oqb_dense_update_cache(oqb_dense* op, int nb, const double* h_coefH, int coef_shared, const double* h_gamma,
                           int gamma_shared, void* d_cache) {
    OQB_DEVICE(op, op->ctx);
    oqb_ctx* c = op->ctx;
    if (nb <= 0) return 0;
    if (!h_coefH || !d_cache) return set_error(c, OQB_EINVAL, "oqb_dense_update_cache: null argument");
    const bool lind = h_gamma && op->K > 0;
    std::vector<c128> coef;
    const int rows = build_coef(op, nb, h_coefH, coef_shared, make_double2(0.0, -1.0), lind ? h_gamma : nullptr,
                                gamma_shared, make_double2(-0.5, 0.0), true, coef);
    const int nmat = op->nterms + (lind ? op->K : 0);
    void* d_coef = nullptr;
    OQB_TRY(coef_upload(c, coef.data(), coef.size() * sizeof(c128), &d_coef));
    return lincomb(c, nmat, (long long)op->n * op->n, op->d_mats, (const c128*)d_coef, rows == 1 ? 0 : nmat,
                   (c128*)d_cache, nb, false);
}

int oqb_dense_update_vectorized_cache(oqb_dense* op, int nb, const double* h_coefH, int coef_shared, void* d_cache) {
    OQB_DEVICE(op, op->ctx);
    oqb_ctx* c = op->ctx;
    if (nb <= 0) return 0;
    const long long nn = (long long)op->n * op->n;
    const int rows = coef_shared ? 1 : nb;
    OQB_TRY(ws_reserve(c, (size_t)rows * nn * sizeof(c128)));
    OQB_TRY(oqb_dense_hamiltonian(op, rows, h_coefH, coef_shared, c->ws));
    return vectorized_commutator(c, op->n, (const c128*)c->ws, rows == 1 ? 0 : nn, (c128*)d_cache, nb);
}

int oqb_dense_rhs(oqb_dense* op, int nb, const double* h_coefH, int coef_shared, const double* h_gamma,
                  int gamma_shared, const void* d_u, void* d_du) {
    OQB_DEVICE(op, op->ctx);
    return dense_rhs_impl(op, nb, h_coefH, coef_shared, h_gamma, gamma_shared, (const c128*)d_u, (c128*)d_du, true);
}

int oqb_lindblad_acc(oqb_dense* op, int nb, const double* h_gamma, int gamma_shared, const void* d_u, void* d_du) {
    OQB_DEVICE(op, op->ctx);
    return dense_rhs_impl(op, nb, nullptr, 1, h_gamma, gamma_shared, (const c128*)d_u, (c128*)d_du, false);
}

int oqb_dense_rhs_host(oqb_dense* op, int nb, const double* h_coefH, int coef_shared, const double* h_gamma,
                       int gamma_shared, const void* h_u, void* h_du) {
    OQB_DEVICE(op, op->ctx);
    oqb_ctx* c = op->ctx;
    const size_t bytes = (size_t)op->n * op->n * sizeof(c128);
    int chunk = (int)(((size_t)256 << 20) / bytes);
    if (chunk < 1) chunk = 1;
    const int nt = op->nterms, K = op->K;
    return host_pipeline(c, nb, bytes, bytes, h_u, h_du, chunk, [&](int b0, int cnt, const void* din, void* dout) {
        const double* cf = coef_shared ? h_coefH : h_coefH + (size_t)b0 * nt;
        const double* gm = (!h_gamma || gamma_shared) ? h_gamma : h_gamma + (size_t)b0 * K;
        return dense_rhs_impl(op, cnt, cf, coef_shared, gm, gamma_shared, (const c128*)din, (c128*)dout, true);
    });
}

int oqb_redfield_apply_acc(oqb_ctx* c, int n, int K, const void* d_S, int S_shared, const void* d_Lambda, int nb,
                           const void* d_u, void* d_du) {
    OQB_DEVICE(c, c);
    if (nb <= 0 || K <= 0) return 0;
    if (!d_S || !d_Lambda || !d_u || !d_du) return set_error(c, OQB_EINVAL, "oqb_redfield_apply_acc: null argument");
    const long long nn = (long long)n * n;
    const bool s_real = (S_shared & 2) != 0;  // caller asserts: no coupling entry has an imaginary part
    S_shared &= 1;
    const size_t per = (size_t)2 * K * nn * sizeof(c128);
    int chunk = (int)(kWorkspaceBudget / per);
    if (chunk < 1) chunk = 1;
    if (chunk > nb) chunk = nb;
    chunk = (nb + (nb + chunk - 1) / chunk - 1) / ((nb + chunk - 1) / chunk);  // balance the chunks
    OQB_TRY(ws_reserve(c, (size_t)chunk * per));
    c128* M1 = (c128*)c->ws;
    c128* Kb = M1 + (size_t)chunk * K * nn;
    for (int b0 = 0; b0 < nb; b0 += chunk) {
        const int cnt = nb - b0 < chunk ? nb - b0 : chunk;
        const c128* u = (const c128*)d_u + (long long)b0 * nn;
        const c128* Lam = (const c128*)d_Lambda + (long long)b0 * K * nn;
        const c128* S = (const c128*)d_S + (S_shared ? 0 : (long long)b0 * K * nn);
        ZgemmParams p;  // M1[b,α] = Λ[b,α] u_b
        p.M = p.N = p.K = n; p.batch = cnt; p.batch2 = K;
        p.A = Lam; p.lda = n; p.strideA = K * nn; p.strideA2 = nn;
        p.B = u; p.ldb = n; p.strideB = nn; p.strideB2 = 0;
        p.C = M1; p.ldc = n; p.strideC = K * nn; p.strideC2 = nn;
        OQB_TRY(zgemm(c, p));
        ZgemmParams q = p;  // Kb[b,α] = S_α M1[b,α]
        q.A = S; q.strideA = S_shared ? 0 : K * nn; q.strideA2 = nn;
        q.real_operand = s_real ? 1 : 0;
        q.B = M1; q.strideB = K * nn; q.strideB2 = nn;
        q.C = Kb;
        OQB_TRY(zgemm(c, q));
        ZgemmParams r = p;  // Kb[b,α] −= M1[b,α] S_α
        r.A = M1; r.strideA = K * nn; r.strideA2 = nn;
        r.B = S; r.strideB = S_shared ? 0 : K * nn; r.strideB2 = nn;
        r.real_operand = s_real ? 2 : 0;
        r.C = Kb;
        r.alpha = make_double2(-1.0, 0.0);
        r.beta = make_double2(1.0, 0.0);
        OQB_TRY(zgemm(c, r));
        OQB_TRY(add_adjoint(c, n, K, -1.0, Kb, (c128*)d_du + (long long)b0 * nn, cnt));  // du −= K₂ + K₂'
    }
    return 0;
}

int oqb_redfield_apply_diag_acc(oqb_ctx* c, int n, int K, const void* d_Sdiag, const void* d_Lambda, int nb, const void* d_u,
                                void* d_du) {
    OQB_DEVICE(c, c);
    if (nb <= 0 || K <= 0) return 0;
    if (!d_Sdiag || !d_Lambda || !d_u || !d_du) return set_error(c, OQB_EINVAL, "oqb_redfield_apply_diag_acc: null argument");
    const long long nn = (long long)n * n;
    const size_t per = (size_t)K * nn * sizeof(c128);
    int chunk = (int)(kWorkspaceBudget / per);
    if (chunk < 1) chunk = 1;
    if (chunk > nb) chunk = nb;
    chunk = (nb + (nb + chunk - 1) / chunk - 1) / ((nb + chunk - 1) / chunk);  // balance the chunks
    OQB_TRY(ws_reserve(c, (size_t)chunk * per));
    c128* M1 = (c128*)c->ws;
    for (int b0 = 0; b0 < nb; b0 += chunk) {
        const int cnt = nb - b0 < chunk ? nb - b0 : chunk;
        ZgemmParams p;  // M1[b,α] = Λ[b,α] u_b
        p.M = p.N = p.K = n; p.batch = cnt; p.batch2 = K;
        p.A = (const c128*)d_Lambda + (long long)b0 * K * nn; p.lda = n; p.strideA = K * nn; p.strideA2 = nn;
        p.B = (const c128*)d_u + (long long)b0 * nn; p.ldb = n; p.strideB = nn; p.strideB2 = 0;
        p.C = M1; p.ldc = n; p.strideC = K * nn; p.strideC2 = nn;
        OQB_TRY(zgemm(c, p));
        OQB_TRY(redfield_diag_combine(c, n, K, (const c128*)d_Sdiag, M1, (c128*)d_du + (long long)b0 * nn, cnt));
    }
    return 0;
}

int oqb_redfield_apply_mono_acc(oqb_ctx* c, int n, int K, const int32_t* d_perm, const int32_t* d_pinv, const void* d_val,
                                const void* d_Lambda, int nb, const void* d_u, void* d_du) {
    OQB_DEVICE(c, c);
    if (nb <= 0 || K <= 0) return 0;
    if (!d_perm || !d_pinv || !d_val || !d_Lambda || !d_u || !d_du) return set_error(c, OQB_EINVAL, "oqb_redfield_apply_mono_acc: null argument");
    const long long nn = (long long)n * n;
    const size_t per = (size_t)K * nn * sizeof(c128);
    int chunk = (int)(kWorkspaceBudget / per);
    if (chunk < 1) chunk = 1;
    if (chunk > nb) chunk = nb;
    chunk = (nb + (nb + chunk - 1) / chunk - 1) / ((nb + chunk - 1) / chunk);  // balance the chunks
    OQB_TRY(ws_reserve(c, (size_t)chunk * per));
    c128* M1 = (c128*)c->ws;
    for (int b0 = 0; b0 < nb; b0 += chunk) {
        const int cnt = nb - b0 < chunk ? nb - b0 : chunk;
        ZgemmParams p;  // M1[b,α] = Λ[b,α] u_b
        p.M = p.N = p.K = n; p.batch = cnt; p.batch2 = K;
        p.A = (const c128*)d_Lambda + (long long)b0 * K * nn; p.lda = n; p.strideA = K * nn; p.strideA2 = nn;
        p.B = (const c128*)d_u + (long long)b0 * nn; p.ldb = n; p.strideB = nn; p.strideB2 = 0;
        p.C = M1; p.ldc = n; p.strideC = K * nn; p.strideC2 = nn;
        OQB_TRY(zgemm(c, p));
        OQB_TRY(redfield_mono_combine(c, n, K, d_perm, d_pinv, (const c128*)d_val, M1, (c128*)d_du + (long long)b0 * nn, cnt));
    }
    return 0;
}

int oqb_redfield_vectorized_acc(oqb_ctx* c, int n, int K, const void* d_S, int S_shared, const void* d_Lambda,
                                int nb, void* d_cache) {
    OQB_DEVICE(c, c);
    if (nb <= 0 || K <= 0) return 0;
    const long long nn = (long long)n * n;
    OQB_TRY(ws_reserve(c, (size_t)nb * nn * sizeof(c128)));
    c128* SL = (c128*)c->ws;
    for (int a = 0; a < K; ++a) {
        const c128* S = (const c128*)d_S + a * nn;
        const c128* Lam = (const c128*)d_Lambda + a * nn;
        ZgemmParams p;  // SΛ
        p.M = p.N = p.K = n; p.batch = nb;
        p.A = S; p.lda = n; p.strideA = S_shared ? 0 : K * nn;
        p.B = Lam; p.ldb = n; p.strideB = K * nn;
        p.C = SL; p.ldc = n; p.strideC = nn;
        OQB_TRY(zgemm(c, p));
        OQB_TRY(redfield_superop_acc(c, n, S, S_shared ? 0 : K * nn, Lam, K * nn, SL, nn, (c128*)d_cache, nb));
    }
    return 0;
}

int oqb_ensemble_population(oqb_ctx* c, int n, int nb, const void* d_psi, double* d_out) {
    OQB_DEVICE(c, c);
    if (!d_psi || !d_out) return set_error(c, OQB_EINVAL, "oqb_ensemble_population: null argument");
    return ensemble_population(c, n, nb, (const c128*)d_psi, d_out);
}

int oqb_expect_diag(oqb_ctx* c, int n, int nobs, const double* d_diag, int nb, const void* d_psi, double* d_out) {
    OQB_DEVICE(c, c);
    if (!d_diag || !d_psi || !d_out) return set_error(c, OQB_EINVAL, "oqb_expect_diag: null argument");
    return expect_diag(c, n, nobs, d_diag, nb, (const c128*)d_psi, d_out);
}

int oqb_expect(oqb_ctx* c, int n, int nobs, const void* d_obs, int nb, const void* d_state, int is_vector,
               double* d_out) {
    OQB_DEVICE(c, c);
    if (!d_obs || !d_state || !d_out) return set_error(c, OQB_EINVAL, "oqb_expect: null argument");
    return expect(c, n, nobs, (const c128*)d_obs, nb, (const c128*)d_state, is_vector, (c128*)d_out);
}

}  // extern "C"
