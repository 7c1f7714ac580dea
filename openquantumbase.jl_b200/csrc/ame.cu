// Eigenbasis Liouvillian: DiffEqLiouvillian{true,false} with DaviesGenerator and/or
// OneSidedAMELiouvillian terms (reference src/opensys/diffeq_liouvillian.jl:92-128,
// src/opensys/davies.jl:21-99, :356-393).  Host orchestration of zgemm.cu + elementwise.cu.
//
// Davies uses the closed form of SURVEY App. A.5 (algebraically identical to the reference's
// per-gap-group loop):   X = C∘ρ̃ + diag(W·diag ρ̃) + cross terms of non-singleton gap groups.
#include <cstring>
#include <vector>

#include "../../include/oqb.h"
#include "common.cuh"
#include "ame.cuh"

using namespace oqb;


namespace {

const size_t kAmeBudget = (size_t)8 << 30;  // bytes of scratch per batch chunk (180 GB HBM per GPU)

// counts entries of V with a non-zero imaginary part (any NaN counts too)
__global__ void imag_nonzero_kernel(long long n, const c128* __restrict__ v, unsigned int* __restrict__ cnt) {
    bool nz = false;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
        nz |= !(v[e].y == 0.0);
    if (__syncthreads_or(nz) && threadIdx.x == 0) atomicAdd(cnt, 1u);
}

// T[α][i + j·n] = d_α[i] · V[i + j·n]  — c_α·V for diagonal couplings (grid.y = α)
__global__ void rowscale_kernel(int n, int l, const c128* __restrict__ diag, const c128* __restrict__ V, c128* __restrict__ T) {
    const long long ln = (long long)n * l;
    const c128* d = diag + (size_t)blockIdx.y * n;
    c128* t = T + (long long)blockIdx.y * ln;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < ln; e += (long long)gridDim.x * blockDim.x) {
        const c128 s = d[e % n], v = V[e];
        t[e] = make_double2(s.x * v.x - s.y * v.y, s.x * v.y + s.y * v.x);
    }
}

}  // namespace

// Cross terms without a separate copy of ρ̃ (the Hadamard stays in the GEMM epilogue): the ρ̃ entries the cross terms read are
// recovered from X = C∘ρ̃ as X/C — exact to rounding whatever |C| is, as long as C ≠ 0 at those positions (checked once per plan).
// All sources are gathered BEFORE anything is scattered, because a cross term's target may be another term's source.
__global__ void cross_gather_kernel(int l, long long ncross, const int32_t* __restrict__ idx, const c128* __restrict__ Cm,
                                    const c128* __restrict__ X, c128* __restrict__ val) {
    const int bb = blockIdx.y;
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= ncross) return;
    const long long pos = idx[4 * e + 1] + (long long)idx[4 * e + 3] * l;  // (b, b2)
    const c128 x = X[(long long)bb * l * l + pos], cv = Cm[pos];
    const double den = cv.x * cv.x + cv.y * cv.y;
    val[(long long)bb * ncross + e] = make_double2((x.x * cv.x + x.y * cv.y) / den, (x.y * cv.x - x.x * cv.y) / den);
}
__global__ void cross_scatter_kernel(int l, int K, const c128* __restrict__ A, long long ncross, const int32_t* __restrict__ idx,
                                     const double* __restrict__ rate, const c128* __restrict__ val, c128* __restrict__ X) {
    const int bb = blockIdx.y;
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= ncross) return;
    const int a = idx[4 * e], b = idx[4 * e + 1], a2 = idx[4 * e + 2], b2 = idx[4 * e + 3];
    double sx = 0.0, sy = 0.0;
    for (int k = 0; k < K; ++k) {
        const c128* Ak = A + (long long)k * l * l;
        const c128 p = Ak[a + (long long)b * l], q = Ak[a2 + (long long)b2 * l];
        sx += p.x * q.x + p.y * q.y;  // p·conj(q)
        sy += p.y * q.x - p.x * q.y;
    }
    const double r = rate[e];
    sx *= r; sy *= r;
    const c128 v = val[(long long)bb * ncross + e];
    c128* d = X + (long long)bb * l * l + a + (long long)a2 * l;
    atomicAdd(&d->x, sx * v.x - sy * v.y);
    atomicAdd(&d->y, sx * v.y + sy * v.x);
}
__global__ void cross_source_zero_kernel(int l, long long ncross, const int32_t* __restrict__ idx, const c128* __restrict__ Cm,
                                         unsigned int* __restrict__ cnt) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= ncross) return;
    const c128 cv = Cm[idx[4 * e + 1] + (long long)idx[4 * e + 3] * l];
    const double den = cv.x * cv.x + cv.y * cv.y;
    if (!(den > 1e-280) || !isfinite(den)) atomicAdd(cnt, 1u);
}

// *out = true when no entry of v[0..n) has a non-zero imaginary part (synchronises the stream)
int ame_detect_real(oqb_ame* a, const c128* v, long long n, bool* out) {
    oqb_ctx* c = a->ctx;
    *out = false;
    if (!a->d_flag) OQB_TRY(ame_alloc(c, (void**)&a->d_flag, 256));
    unsigned int h_cnt = 1;
    OQB_CUDA(c, cudaMemsetAsync(a->d_flag, 0, sizeof(unsigned int), c->stream));
    imag_nonzero_kernel<<<296, 256, 0, c->stream>>>(n, v, a->d_flag);
    OQB_LAUNCH_CHECK(c);
    OQB_CUDA(c, cudaMemcpyAsync(&h_cnt, a->d_flag, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    *out = h_cnt == 0;
    return 0;
}

namespace {

int detect_real(oqb_ame* a, const c128* v, long long n, bool* out) { return ame_detect_real(a, v, n, out); }

bool needs_general_path(const oqb_ame* a) { return a->ncross > 0 || a->d_Moff || a->npairs > 0 || a->nd > 0; }

const int kDenseSlab = 16;  // dense per-group blocks processed per pair of products (scratch: slab·l² per member)

// Cross terms only (no off-diagonal L'L matrix, no dense groups, no one-sided / correlated terms, a list small enough to
// keep one recovered value per entry and member): the Hadamard can stay in GEMM 2's epilogue.  Decided once per plan.
int cross_fusable(oqb_ame* a, bool* out) {
    oqb_ctx* c = a->ctx;
    *out = false;
    if (!c->opt.fuse_cross || !(a->ncross > 0 && a->ncross <= 2000000) || a->d_Moff || a->npairs > 0 || a->nd > 0 || !a->corr_terms.empty() || !a->davies)
        return 0;
    if (a->cross_fuse_state < 0) {
        if (!a->d_flag) OQB_TRY(ame_alloc(c, (void**)&a->d_flag, 256));
        unsigned int h_cnt = 1;
        OQB_CUDA(c, cudaMemsetAsync(a->d_flag, 0, sizeof(unsigned int), c->stream));
        cross_source_zero_kernel<<<(unsigned)((a->ncross + 255) / 256), 256, 0, c->stream>>>(a->lvl, a->ncross, a->d_cross_idx, a->d_Cm, a->d_flag);
        OQB_LAUNCH_CHECK(c);
        OQB_CUDA(c, cudaMemcpyAsync(&h_cnt, a->d_flag, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
        OQB_CUDA(c, cudaStreamSynchronize(c->stream));
        a->cross_fuse_state = h_cnt == 0 ? 1 : 0;
    }
    *out = a->cross_fuse_state == 1;
    return 0;
}

// Eigenbasis terms on ρ̃ = R (lvl×lvl per member).  overwrite: X = …, else X += ….
// with_h: include −i[diag w, ρ̃].  scratch: nb·l + (npairs > 0 ? (npairs + 1)·nb·l² : 0) complex.
int eig_apply(oqb_ame* a, int nb, const c128* R, c128* X, bool overwrite, bool with_h, c128* scratch) {
    oqb_ctx* c = a->ctx;
    const int l = a->lvl;
    const long long ll = (long long)l * l;
    OQB_TRY(hadamard(c, l, a->d_Cm, with_h ? nullptr : a->d_w, R, X, nb, !overwrite));
    c128* pop = scratch;
    c128* M1 = scratch + (size_t)nb * l;
    c128* Kb = M1 + (size_t)nb * ll * (a->npairs > 0 ? a->npairs : 0);
    c128* Md = Kb + (a->npairs > 0 ? (size_t)nb * ll : 0);  // dense per-group scratch: nb × slab l×l blocks
    if (a->davies) {
        OQB_TRY(extract_diag(c, l, R, pop, nb));
        OQB_TRY(davies_diag(c, l, a->d_Wt, pop, X, nb));
        if (a->d_Wt_im) OQB_TRY(davies_diag(c, l, a->d_Wt_im, pop, X, nb, true));
        for (const DaviesTerm& t : a->terms)  // every DaviesGenerator of opensys_eig: its couplings, its γ
            OQB_TRY(davies_cross(c, l, t.nk, a->d_A + (long long)t.k0 * ll, a->ncross, a->d_cross_idx, t.d_rate, R, X, nb));
        if (!a->corr_terms.empty()) OQB_TRY(ame_corr_cross(a, nb, R, X));  // pairs (α,β) of a CorrelatedDaviesGenerator
        if (a->d_Moff) {
            ZgemmParams p;  // X += Moff R
            p.M = p.N = p.K = l; p.batch = nb;
            p.A = a->d_Moff; p.lda = l; p.strideA = 0;
            p.B = R; p.ldb = l; p.strideB = ll;
            p.C = X; p.ldc = l; p.strideC = ll;
            p.beta = make_double2(1.0, 0.0);
            OQB_TRY(zgemm(c, p));
            ZgemmParams q;  // X += R Moff'
            q.M = q.N = q.K = l; q.batch = nb;
            q.A = R; q.lda = l; q.strideA = ll;
            if (a->d_Moff2) { q.B = a->d_Moff2; q.ldb = l; q.strideB = 0; }  // correlated pairs: the right-hand matrix is not Moff'
            else { q.B = a->d_Moff; q.ldb = l; q.strideB = 0; q.opB = OP_C; }
            q.C = X; q.ldc = l; q.strideC = ll;
            q.beta = make_double2(1.0, 0.0);
            OQB_TRY(zgemm(c, q));
        }
        // Large gap groups (degenerate spectra: the start of every anneal) as dense products, group by group like the
        // reference's loop (src/opensys/davies.jl:26-45):  X += Σ_j γ_j L_j ρ̃ L_j',  L_j = A_k∘[ω̃ == x] — the blocks of a slab
        // side by side, so the sum over j rides the second product's k loop.
        for (int j0 = 0; j0 < a->nd; j0 += kDenseSlab) {
            const int jc = a->nd - j0 < kDenseSlab ? a->nd - j0 : kDenseSlab;
            ZgemmParams p;  // Md[b, j] = γ_j L_j ρ̃_b
            p.M = p.N = p.K = l; p.batch = nb; p.batch2 = jc;
            p.A = a->d_Lstack + (long long)j0 * ll; p.lda = l; p.strideA = 0; p.strideA2 = ll;
            p.real_operand = a->dense_real ? 1 : 0;
            p.B = R; p.ldb = l; p.strideB = ll; p.strideB2 = 0;
            p.C = Md; p.ldc = l; p.strideC = (long long)jc * ll; p.strideC2 = ll;
            p.alpha_scale = a->d_drate + j0; p.alpha_scale_mod = jc;
            OQB_TRY(zgemm(c, p));
            ZgemmParams q;  // X_b += [Md_b1 | … | Md_bJ]·[L_1 | … | L_J]'   (depth J·l)
            q.M = q.N = l; q.K = jc * l; q.batch = nb;
            q.A = Md; q.lda = l; q.strideA = (long long)jc * ll;
            q.B = (a->d_Rstack ? a->d_Rstack : a->d_Lstack) + (long long)j0 * ll; q.ldb = l; q.strideB = 0; q.opB = OP_C;
            q.real_operand = a->dense_real ? 2 : 0;
            q.C = X; q.ldc = l; q.strideC = ll;
            q.beta = make_double2(1.0, 0.0);
            OQB_TRY(zgemm(c, q));
        }
    }
    if (a->npairs > 0) {
        // src/opensys/davies.jl:369-378:  K₂ = Σ_p (A_βΛ_p ρ̃ − Λ_p ρ̃ A_β) = P ρ̃ − [Λ_1ρ̃ | … | Λ_Pρ̃]·[A_β1; …; A_βP]
        // with P = Σ_p A_βΛ_p precomputed per eigen-system: npairs + 1 products of depth l and ONE of depth npairs·l per
        // member instead of 3·npairs products of depth l.
        const int P = a->npairs;
        ZgemmParams p;  // M1[b,p] = Λ_p ρ̃_b, the blocks of member b side by side (l × P·l)
        p.M = p.N = p.K = l; p.batch = nb; p.batch2 = P;
        p.A = a->d_Lam; p.lda = l; p.strideA = 0; p.strideA2 = ll;
        p.real_operand = a->lam_real ? 1 : 0;
        p.B = R; p.ldb = l; p.strideB = ll; p.strideB2 = 0;
        p.C = M1; p.ldc = l; p.strideC = (long long)P * ll; p.strideC2 = ll;
        OQB_TRY(zgemm(c, p));
        ZgemmParams q;  // Kb = P ρ̃
        q.M = q.N = q.K = l; q.batch = nb;
        q.A = a->d_Psum; q.lda = l; q.strideA = 0;
        q.real_operand = a->psum_real ? 1 : 0;
        q.B = R; q.ldb = l; q.strideB = ll;
        q.C = Kb; q.ldc = l; q.strideC = ll;
        OQB_TRY(zgemm(c, q));
        ZgemmParams r;  // Kb −= M1stack · Abstack   (depth P·l)
        r.M = r.N = l; r.K = P * l; r.batch = nb;
        r.A = M1; r.lda = l; r.strideA = (long long)P * ll;
        r.B = a->d_Abstack; r.ldb = (long long)P * l; r.strideB = 0;
        r.real_operand = a->ab_real ? 2 : 0;
        r.C = Kb; r.ldc = l; r.strideC = ll;
        r.alpha = make_double2(-1.0, 0.0);
        r.beta = make_double2(1.0, 0.0);
        OQB_TRY(zgemm(c, r));
        OQB_TRY(add_adjoint(c, l, 1, -1.0, Kb, X, nb));  // dρ −= K₂ + K₂'
    }
    return 0;
}

size_t eig_scratch_elems(const oqb_ame* a, int nb) {
    const size_t ll = (size_t)a->lvl * a->lvl;
    const size_t slab = (size_t)(a->nd < kDenseSlab ? a->nd : kDenseSlab);
    return (size_t)nb * a->lvl + (a->npairs > 0 ? ((size_t)a->npairs + 1) * nb * ll : 0) + slab * nb * ll;
}

int ame_rhs_impl(oqb_ame* a, int nb, const c128* u, c128* du) {
    oqb_ctx* c = a->ctx;
    if (nb <= 0) return 0;
    if (!a->have_eigen) return set_error(c, OQB_EINVAL, "oqb_ame_rhs: call oqb_ame_set_eigen first");
    if (!u || !du) return set_error(c, OQB_EINVAL, "oqb_ame_rhs: null state pointer");
    const int n = a->n, l = a->lvl;
    const long long nn = (long long)n * n, ll = (long long)l * l, ln = (long long)l * n;
    bool fuse_cross = false;
    OQB_TRY(cross_fusable(a, &fuse_cross));
    const bool general = needs_general_path(a) && !fuse_cross;
    // per-member scratch: T (l×n) + X (l×l) [+ R (l×l)] + eig scratch (fused cross terms: one recovered value per entry)
    const size_t per = ((size_t)ln + ll + (general ? ll : 0) + eig_scratch_elems(a, 1) + (fuse_cross ? (size_t)a->ncross : 0)) * sizeof(c128);
    int chunk = (int)(kAmeBudget / per);
    if (chunk < 1) chunk = 1;
    if (chunk > nb) chunk = nb;
    chunk = (nb + (nb + chunk - 1) / chunk - 1) / ((nb + chunk - 1) / chunk);  // balance the chunks
    OQB_TRY(ws_reserve(c, (size_t)chunk * per));
    c128* T = (c128*)c->ws;
    c128* X = T + (size_t)chunk * ln;
    c128* R = X + (size_t)chunk * ll;
    c128* scratch = general ? R + (size_t)chunk * ll : R;

    for (int b0 = 0; b0 < nb; b0 += chunk) {
        const int cnt = nb - b0 < chunk ? nb - b0 : chunk;
        const c128* ub = u + (long long)b0 * nn;
        c128* dub = du + (long long)b0 * nn;
        if (!a->has_v) {
            // adiabatic frame: v = I, ρ̃ = u   (src/opensys/diffeq_liouvillian.jl:130-140 with diagonal H)
            OQB_TRY(eig_apply(a, cnt, ub, dub, true, true, scratch));
            continue;
        }
        ZgemmParams p;  // T = v' u            (l × n)
        p.M = l; p.N = n; p.K = n; p.batch = cnt;
        p.opA = OP_C;
        p.A = a->d_V; p.lda = n; p.strideA = 0;
        p.real_operand = a->v_real ? 1 : 0;
        p.B = ub; p.ldb = n; p.strideB = nn;
        p.C = T; p.ldc = l; p.strideC = ln;
        OQB_TRY(zgemm(c, p));
        ZgemmParams q;  // ρ̃ = T v            (l × l)
        q.M = l; q.N = l; q.K = n; q.batch = cnt;
        q.A = T; q.lda = l; q.strideA = ln;
        q.B = a->d_V; q.ldb = n; q.strideB = 0;
        q.real_operand = a->v_real ? 2 : 0;
        q.ldc = l; q.strideC = ll;
        if (!general) {
            // fused: X = C∘ρ̃ in the GEMM epilogue, diagonal of ρ̃ captured for the W·pop term
            q.C = X;
            q.E = a->d_Cm; q.lde = l; q.strideE = 0;
            if (a->davies) { q.diag_out = scratch; q.strideDiag = l; }
            OQB_TRY(zgemm(c, q));
            c128* val = scratch + (size_t)cnt * l;
            if (fuse_cross) {
                // ρ̃ at the cross sources = X/C, gathered while X is still exactly C∘ρ̃ (before the diagonal rate term and before
                // any scatter: a cross term's target may be another term's source)
                for (int b0 = 0; b0 < cnt; b0 += 65535) {
                    const int bc = cnt - b0 < 65535 ? cnt - b0 : 65535;
                    dim3 grid((unsigned)((a->ncross + 255) / 256), (unsigned)bc);
                    cross_gather_kernel<<<grid, 256, 0, c->stream>>>(l, a->ncross, a->d_cross_idx, a->d_Cm, X + (long long)b0 * ll, val + (long long)b0 * a->ncross);
                    OQB_LAUNCH_CHECK(c);
                }
            }
            if (a->davies) OQB_TRY(davies_diag(c, l, a->d_Wt, scratch, X, cnt));
            if (a->davies && a->d_Wt_im) OQB_TRY(davies_diag(c, l, a->d_Wt_im, scratch, X, cnt, true));
            if (fuse_cross) {
                for (const DaviesTerm& t : a->terms)
                    for (int b0 = 0; b0 < cnt; b0 += 65535) {
                        const int bc = cnt - b0 < 65535 ? cnt - b0 : 65535;
                        dim3 grid((unsigned)((a->ncross + 255) / 256), (unsigned)bc);
                        cross_scatter_kernel<<<grid, 256, 0, c->stream>>>(l, t.nk, a->d_A + (long long)t.k0 * ll, a->ncross, a->d_cross_idx, t.d_rate,
                                                                          val + (long long)b0 * a->ncross, X + (long long)b0 * ll);
                        OQB_LAUNCH_CHECK(c);
                    }
            }
        } else {
            q.C = R;
            OQB_TRY(zgemm(c, q));
            OQB_TRY(eig_apply(a, cnt, R, X, true, true, scratch));
        }
        ZgemmParams r;  // T = X v'           (l × n)
        r.M = l; r.N = n; r.K = l; r.batch = cnt;
        r.A = X; r.lda = l; r.strideA = ll;
        r.B = a->d_V; r.ldb = n; r.strideB = 0; r.opB = OP_C;
        r.real_operand = a->v_real ? 2 : 0;
        r.C = T; r.ldc = l; r.strideC = ln;
        OQB_TRY(zgemm(c, r));
        ZgemmParams s;  // du = v T           (n × n), OVERWRITES  (diffeq_liouvillian.jl:105)
        s.M = n; s.N = n; s.K = l; s.batch = cnt;
        s.A = a->d_V; s.lda = n; s.strideA = 0;
        s.real_operand = a->v_real ? 1 : 0;
        s.B = T; s.ldb = l; s.strideB = ln;
        s.C = dub; s.ldc = n; s.strideC = nn;
        OQB_TRY(zgemm(c, s));
    }
    return 0;
}

}  // namespace

extern "C" {

int oqb_ame_create(oqb_ctx* c, int n, int lvl, int K, const void* h_couplings, oqb_ame** out) {
    OQB_DEVICE(c, c);
    if (!c || !out) return OQB_EINVAL;
    *out = nullptr;
    if (n <= 0 || lvl <= 0 || lvl > n || K < 0 || (K > 0 && !h_couplings))
        return set_error(c, OQB_EINVAL, "oqb_ame_create: bad arguments (n=%d lvl=%d K=%d)", n, lvl, K);
    oqb_ame* a = new oqb_ame();
    a->ctx = c; a->n = n; a->lvl = lvl; a->K = K;
    const size_t nn = (size_t)n * n, ll = (size_t)lvl * lvl;
    int s = ame_alloc(c, (void**)&a->d_coup, (K + 1) * nn * sizeof(c128));
    if (!s) s = ame_alloc(c, (void**)&a->d_V, (size_t)n * lvl * sizeof(c128));
    if (!s) s = ame_alloc(c, (void**)&a->d_w, lvl * sizeof(double));
    if (!s) s = ame_alloc(c, (void**)&a->d_A, (K + 1) * ll * sizeof(c128));
    if (!s) s = ame_alloc(c, (void**)&a->d_Cm, ll * sizeof(c128));
    if (s) { oqb_ame_destroy(a); return s; }
    if (K) {
        OQB_CUDA(c, cudaMemcpyAsync(a->d_coup, h_couplings, K * nn * sizeof(c128), cudaMemcpyHostToDevice, c->stream));
        // every coupling diagonal (σz-type system operators stored as dense matrices)?  Then c_α·v is a row scaling and
        // the rotation v'c_αv of every new eigen-system needs one GEMM per coupling instead of two.
        const c128* Cm = (const c128*)h_couplings;
        bool diag = true;
        for (int k = 0; diag && k < K; ++k)
            for (size_t col = 0; diag && col < (size_t)n; ++col)
                for (size_t row = 0; row < (size_t)n; ++row) {
                    if (row == col) continue;
                    const c128 x = Cm[(size_t)k * nn + row + col * n];
                    if (x.x != 0.0 || x.y != 0.0) { diag = false; break; }
                }
        std::vector<c128> dg;
        if (diag) {
            dg.resize((size_t)K * n);
            for (int k = 0; k < K; ++k)
                for (int i = 0; i < n; ++i) dg[(size_t)k * n + i] = Cm[(size_t)k * nn + i + (size_t)i * n];
            s = ame_alloc(c, (void**)&a->d_cdiag, dg.size() * sizeof(c128));
            if (s) { oqb_ame_destroy(a); return s; }
            OQB_CUDA(c, cudaMemcpyAsync(a->d_cdiag, dg.data(), dg.size() * sizeof(c128), cudaMemcpyHostToDevice, c->stream));
        }
        OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    *out = a;
    return OQB_OK;
}

int oqb_ame_clear_davies(oqb_ame* a) {
    OQB_DEVICE(a, a->ctx);
    cudaStreamSynchronize(a->ctx->stream);
    ame_free(a->ctx, a->d_gam); ame_free(a->ctx, a->d_S); ame_free(a->ctx, a->d_Gam); ame_free(a->ctx, a->d_hls); ame_free(a->ctx, a->d_Wt); ame_free(a->ctx, a->d_Wt_im);
    ame_free(a->ctx, a->d_Moff);
    ame_clear_terms(a);
    a->davies = false;
    if (a->have_eigen)  // Hamiltonian-only Hadamard coefficients
        return davies_tables(a->ctx, a->lvl, a->K, a->d_A, nullptr, nullptr, a->d_w, 0, nullptr, nullptr, nullptr, a->d_Cm);
    return 0;
}

int oqb_ame_rebind(oqb_ame* a, oqb_ctx* ctx) {
    OQB_DEVICE(a, a->ctx);
    if (!a || !ctx) return OQB_EINVAL;
    if (ctx->device != a->ctx->device) return set_error(ctx, OQB_EINVAL, "oqb_ame_rebind: contexts live on different devices");
    a->ctx = ctx;
    return 0;
}

int oqb_ame_clear_onesided(oqb_ame* a) {
    OQB_DEVICE(a, a->ctx);
    cudaStreamSynchronize(a->ctx->stream);
    ame_free(a->ctx, a->d_Lam);
    ame_free(a->ctx, a->d_Psum);
    ame_free(a->ctx, a->d_Abstack);
    a->npairs = 0;
    a->lam_real = a->psum_real = a->ab_real = false;
    a->h_beta.clear();
    return 0;
}

int oqb_ame_destroy(oqb_ame* a) {
    OQB_DEVICE(a, a->ctx);
    if (!a) return OQB_OK;
    a->have_eigen = false;
    oqb_ame_clear_davies(a);
    oqb_ame_clear_onesided(a);
    oqb_ame_clear_jump(a);
    ame_clear_gaps(a);
    ame_free(a->ctx, a->d_flag);
    ame_free(a->ctx, a->d_cdiag);
    ame_free(a->ctx, a->d_Scoef);
    ame_free(a->ctx, a->d_coup); ame_free(a->ctx, a->d_V); ame_free(a->ctx, a->d_w); ame_free(a->ctx, a->d_A); ame_free(a->ctx, a->d_Cm);
    delete a;
    return OQB_OK;
}

static int set_eigen_impl(oqb_ame* a, const double* w, const void* v, bool on_device, const char* who) {
    oqb_ctx* c = a->ctx;
    if (!w) return set_error(c, OQB_EINVAL, "%s: null w", who);
    const int n = a->n, l = a->lvl, K = a->K;
    const long long nn = (long long)n * n, ll = (long long)l * l, ln = (long long)l * n;
    if (!v && l != n) return set_error(c, OQB_EINVAL, "%s: adiabatic frame needs lvl == n", who);
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    // tables of a previous eigen-system are stale
    a->have_eigen = false;
    OQB_TRY(oqb_ame_clear_davies(a));
    OQB_TRY(oqb_ame_clear_onesided(a));
    OQB_TRY(oqb_ame_clear_jump(a));
    OQB_TRY(ame_clear_gaps(a));
    OQB_CUDA(c, cudaMemcpyAsync(a->d_w, w, l * sizeof(double), kind, c->stream));
    a->has_v = v != nullptr;
    a->v_real = false;
    if (a->has_v) {
        OQB_CUDA(c, cudaMemcpyAsync(a->d_V, v, (size_t)ln * sizeof(c128), kind, c->stream));
        // real eigenvectors (real-symmetric H): the rotations below and in every RHS use the real-operand product
        OQB_TRY(detect_real(a, a->d_V, ln, &a->v_real));
        if (K) {  // A_α = v' c_α v   (src/opensys/davies.jl:24)
            OQB_TRY(ws_reserve(c, (size_t)K * ln * sizeof(c128)));
            c128* T = (c128*)c->ws;
            if (a->d_cdiag && c->opt.coupling_diag) {  // T_α = d_α ∘ v (rows scaled)
                dim3 grid((unsigned)std::min<long long>((ln + 255) / 256, 592), (unsigned)K);
                rowscale_kernel<<<grid, 256, 0, c->stream>>>(n, l, a->d_cdiag, a->d_V, T);
                OQB_LAUNCH_CHECK(c);
            } else {
                ZgemmParams p;  // T_α = c_α v   (n × l)
                p.M = n; p.N = l; p.K = n; p.batch = K;
                p.A = a->d_coup; p.lda = n; p.strideA = nn;
                p.B = a->d_V; p.ldb = n; p.strideB = 0;
                p.real_operand = a->v_real ? 2 : 0;
                p.C = T; p.ldc = n; p.strideC = ln;
                OQB_TRY(zgemm(c, p));
            }
            ZgemmParams q;  // A_α = v' T_α  (l × l)
            q.M = l; q.N = l; q.K = n; q.batch = K;
            q.opA = OP_C;
            q.A = a->d_V; q.lda = n; q.strideA = 0;
            q.real_operand = a->v_real ? 1 : 0;
            q.B = T; q.ldb = n; q.strideB = ln;
            q.C = a->d_A; q.ldc = l; q.strideC = ll;
            OQB_TRY(zgemm(c, q));
        }
    } else if (K) {
        OQB_CUDA(c, cudaMemcpyAsync(a->d_A, a->d_coup, (size_t)K * nn * sizeof(c128), cudaMemcpyDeviceToDevice, c->stream));
    }
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));  // w / v may be reused by the caller
    a->have_eigen = true;
    return davies_tables(c, l, K, a->d_A, nullptr, nullptr, a->d_w, 0, nullptr, nullptr, nullptr, a->d_Cm);
}

int oqb_ame_eigenvectors_real(const oqb_ame* a, int* out) {
    OQB_DEVICE(a, a->ctx);
    if (!a || !out) return OQB_EINVAL;
    *out = (a->have_eigen && a->has_v && a->v_real) ? 1 : 0;
    return 0;
}

int oqb_ame_set_eigen(oqb_ame* a, const double* h_w, const void* h_v) {
    OQB_DEVICE(a, a->ctx);
    return set_eigen_impl(a, h_w, h_v, false, "oqb_ame_set_eigen");
}

int oqb_ame_set_eigen_device(oqb_ame* a, const double* d_w, const void* d_v) {
    OQB_DEVICE(a, a->ctx);
    return set_eigen_impl(a, d_w, d_v, true, "oqb_ame_set_eigen_device");
}

int oqb_ame_get_levels(oqb_ame* a, double* h_w) {
    OQB_DEVICE(a, a->ctx);
    oqb_ctx* c = a->ctx;
    if (!a->have_eigen) return set_error(c, OQB_EINVAL, "oqb_ame_get_levels: call oqb_ame_set_eigen first");
    if (!h_w) return set_error(c, OQB_EINVAL, "oqb_ame_get_levels: null output");
    OQB_CUDA(c, cudaMemcpyAsync(h_w, a->d_w, (size_t)a->lvl * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int oqb_ame_get_eigenvectors(oqb_ame* a, void* h_V) {
    OQB_DEVICE(a, a->ctx);
    oqb_ctx* c = a->ctx;
    if (!a->have_eigen || !a->has_v) return set_error(c, OQB_EINVAL, "oqb_ame_get_eigenvectors: no eigenvector matrix is set");
    if (!h_V) return set_error(c, OQB_EINVAL, "oqb_ame_get_eigenvectors: null output");
    OQB_CUDA(c, cudaMemcpyAsync(h_V, a->d_V, (size_t)a->n * a->lvl * sizeof(c128), cudaMemcpyDeviceToHost, c->stream));
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int oqb_ame_get_rotated_couplings(oqb_ame* a, void* h_A) {
    OQB_DEVICE(a, a->ctx);
    oqb_ctx* c = a->ctx;
    if (!a->have_eigen) return set_error(c, OQB_EINVAL, "oqb_ame_get_rotated_couplings: call oqb_ame_set_eigen first");
    if (!h_A) return set_error(c, OQB_EINVAL, "oqb_ame_get_rotated_couplings: null output");
    OQB_CUDA(c, cudaMemcpyAsync(h_A, a->d_A, (size_t)a->K * a->lvl * a->lvl * sizeof(c128), cudaMemcpyDeviceToHost, c->stream));
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// Davies tables: γ/S at the signed rounded gaps are in a->d_gam / a->d_S (device); builds the derived tables and
// installs the cross-term lists of multi-pair gap groups.
int ame_davies_alloc(oqb_ame* a) {
    oqb_ctx* c = a->ctx;
    OQB_TRY(oqb_ame_clear_davies(a));
    const int l = a->lvl;
    const size_t ll = (size_t)l * l;
    OQB_TRY(ame_alloc(c, (void**)&a->d_gam, ll * sizeof(double)));
    OQB_TRY(ame_alloc(c, (void**)&a->d_S, ll * sizeof(double)));
    OQB_TRY(ame_alloc(c, (void**)&a->d_Gam, l * sizeof(double)));
    OQB_TRY(ame_alloc(c, (void**)&a->d_hls, l * sizeof(double)));
    OQB_TRY(ame_alloc(c, (void**)&a->d_Wt, ll * sizeof(double)));
    return 0;
}

int ame_davies_finish(oqb_ame* a, int64_t ncross, const int32_t* h_cross_idx, const double* h_cross_rate, int64_t nlamb,
                      const int32_t* h_lamb_idx, const double* h_lamb_rate_S) {
    oqb_ctx* c = a->ctx;
    const int l = a->lvl;
    const size_t ll = (size_t)l * l;
    OQB_TRY(davies_tables(c, l, a->K, a->d_A, a->d_gam, a->d_S, a->d_w, 1, a->d_Gam, a->d_hls, a->d_Wt, a->d_Cm));
    if (ncross > 0) {
        DaviesTerm t;  // host-supplied lists: one generator over all K couplings
        t.k0 = 0; t.nk = a->K;
        OQB_TRY(ame_alloc(c, (void**)&a->d_cross_idx, (size_t)ncross * 4 * sizeof(int32_t)));
        OQB_TRY(ame_alloc(c, (void**)&t.d_rate, (size_t)ncross * sizeof(double)));
        OQB_CUDA(c, cudaMemcpyAsync(a->d_cross_idx, h_cross_idx, (size_t)ncross * 4 * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        OQB_CUDA(c, cudaMemcpyAsync(t.d_rate, h_cross_rate, (size_t)ncross * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        a->terms.push_back(t);
        a->ncross = ncross;
    }
    if (nlamb > 0) {
        OQB_TRY(ame_alloc(c, (void**)&a->d_Moff, ll * sizeof(c128)));
        OQB_CUDA(c, cudaMemsetAsync(a->d_Moff, 0, ll * sizeof(c128), c->stream));
        int32_t* d_idx = nullptr;
        double* d_rs = nullptr;
        OQB_TRY(ame_alloc(c, (void**)&d_idx, (size_t)nlamb * 3 * sizeof(int32_t)));
        OQB_TRY(ame_alloc(c, (void**)&d_rs, (size_t)nlamb * 2 * sizeof(double)));
        OQB_CUDA(c, cudaMemcpyAsync(d_idx, h_lamb_idx, (size_t)nlamb * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        OQB_CUDA(c, cudaMemcpyAsync(d_rs, h_lamb_rate_S, (size_t)nlamb * 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        int s = davies_lamb_build(c, l, a->K, a->d_A, nlamb, d_idx, d_rs, a->d_Moff);
        cudaStreamSynchronize(c->stream);
        cudaFreeAsync(d_idx, c->stream);
        cudaFreeAsync(d_rs, c->stream);
        if (s) return s;
        a->nlamb = nlamb;
    }
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    a->davies = true;
    a->cross_fuse_state = -1;
    return 0;
}

int oqb_ame_set_davies(oqb_ame* a, const double* h_gam, const double* h_S, int64_t ncross, const int32_t* h_cross_idx,
                       const double* h_cross_rate, int64_t nlamb, const int32_t* h_lamb_idx,
                       const double* h_lamb_rate_S) {
    OQB_DEVICE(a, a->ctx);
    oqb_ctx* c = a->ctx;
    if (!a->have_eigen) return set_error(c, OQB_EINVAL, "oqb_ame_set_davies: call oqb_ame_set_eigen first");
    if (!h_gam || !h_S) return set_error(c, OQB_EINVAL, "oqb_ame_set_davies: null table");
    if ((ncross > 0 && (!h_cross_idx || !h_cross_rate)) || (nlamb > 0 && (!h_lamb_idx || !h_lamb_rate_S)))
        return set_error(c, OQB_EINVAL, "oqb_ame_set_davies: null cross-term arrays");
    OQB_TRY(ame_davies_alloc(a));
    const size_t ll = (size_t)a->lvl * a->lvl;
    OQB_CUDA(c, cudaMemcpyAsync(a->d_gam, h_gam, ll * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    OQB_CUDA(c, cudaMemcpyAsync(a->d_S, h_S, ll * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    return ame_davies_finish(a, ncross, h_cross_idx, h_cross_rate, nlamb, h_lamb_idx, h_lamb_rate_S);
}

// Λ_p = (½γ + iS)∘A_αp from DEVICE tables, then the per-eigen-system products of the one-sided generator
int ame_onesided_install(oqb_ame* a, int npairs, const int32_t* h_alpha, const int32_t* h_beta, const double* d_g,
                         const double* d_s, int tables_shared) {
    oqb_ctx* c = a->ctx;
    for (int p = 0; p < npairs; ++p)
        if (h_alpha[p] < 0 || h_alpha[p] >= a->K || h_beta[p] < 0 || h_beta[p] >= a->K)
            return set_error(c, OQB_EINVAL, "oqb_ame_set_onesided: coupling index out of range");
    OQB_TRY(oqb_ame_clear_onesided(a));
    const int l = a->lvl;
    const size_t ll = (size_t)l * l;
    int32_t* d_alpha = nullptr;
    OQB_TRY(ame_alloc(c, (void**)&a->d_Lam, (size_t)npairs * ll * sizeof(c128)));
    OQB_TRY(ame_alloc(c, (void**)&d_alpha, npairs * sizeof(int32_t)));
    OQB_CUDA(c, cudaMemcpyAsync(d_alpha, h_alpha, npairs * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    int s = onesided_lambda(c, l, npairs, d_alpha, a->d_A, d_g, d_s, tables_shared ? 0 : (long long)ll, a->d_Lam);
    cudaStreamSynchronize(c->stream);
    cudaFreeAsync(d_alpha, c->stream);
    if (s) return s;
    a->npairs = npairs;
    a->h_beta.assign(h_beta, h_beta + npairs);
    // per eigen-system: P = Σ_p A_βp Λ_p and the vertical stack [A_β1; …; A_βP] (npairs·l × l)
    OQB_TRY(ame_alloc(c, (void**)&a->d_Psum, ll * sizeof(c128)));
    OQB_TRY(ame_alloc(c, (void**)&a->d_Abstack, (size_t)npairs * ll * sizeof(c128)));
    for (int p = 0; p < npairs; ++p) {
        const c128* Ab = a->d_A + (long long)h_beta[p] * ll;
        ZgemmParams g;
        g.M = g.N = g.K = l; g.batch = 1;
        g.A = Ab; g.lda = l;
        g.B = a->d_Lam + (long long)p * ll; g.ldb = l;
        g.C = a->d_Psum; g.ldc = l;
        g.beta = p == 0 ? make_double2(0.0, 0.0) : make_double2(1.0, 0.0);
        OQB_TRY(zgemm(c, g));
        OQB_CUDA(c, cudaMemcpy2DAsync(a->d_Abstack + (long long)p * l, (size_t)npairs * l * sizeof(c128), Ab, (size_t)l * sizeof(c128),
                                      (size_t)l * sizeof(c128), l, cudaMemcpyDeviceToDevice, c->stream));
    }
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    // real couplings in a real eigenbasis with S ≡ 0 make Λ_p, ΣA_βΛ_p and the A_β stack real: real-operand products
    OQB_TRY(detect_real(a, a->d_Lam, (long long)npairs * ll, &a->lam_real));
    OQB_TRY(detect_real(a, a->d_Psum, (long long)ll, &a->psum_real));
    OQB_TRY(detect_real(a, a->d_Abstack, (long long)npairs * ll, &a->ab_real));
    return 0;
}

int oqb_ame_set_onesided(oqb_ame* a, int npairs, const int32_t* h_alpha, const int32_t* h_beta, const double* h_gam,
                         const double* h_S, int tables_shared) {
    OQB_DEVICE(a, a->ctx);
    oqb_ctx* c = a->ctx;
    if (!a->have_eigen) return set_error(c, OQB_EINVAL, "oqb_ame_set_onesided: call oqb_ame_set_eigen first");
    if (npairs <= 0 || !h_alpha || !h_beta || !h_gam || !h_S) return set_error(c, OQB_EINVAL, "oqb_ame_set_onesided: bad arguments");
    const size_t ll = (size_t)a->lvl * a->lvl;
    const size_t nt = tables_shared ? 1 : (size_t)npairs;
    double *d_g = nullptr, *d_s = nullptr;
    OQB_TRY(ame_alloc(c, (void**)&d_g, nt * ll * sizeof(double)));
    OQB_TRY(ame_alloc(c, (void**)&d_s, nt * ll * sizeof(double)));
    OQB_CUDA(c, cudaMemcpyAsync(d_g, h_gam, nt * ll * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    OQB_CUDA(c, cudaMemcpyAsync(d_s, h_S, nt * ll * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const int s = ame_onesided_install(a, npairs, h_alpha, h_beta, d_g, d_s, tables_shared);
    cudaStreamSynchronize(c->stream);
    cudaFreeAsync(d_g, c->stream); cudaFreeAsync(d_s, c->stream);
    return s;
}

int oqb_ame_rhs(oqb_ame* a, int nb, const void* d_u, void* d_du) {
    OQB_DEVICE(a, a->ctx);
    return ame_rhs_impl(a, nb, (const c128*)d_u, (c128*)d_du);
}

int oqb_ame_rhs_host(oqb_ame* a, int nb, const void* h_u, void* h_du) {
    OQB_DEVICE(a, a->ctx);
    oqb_ctx* c = a->ctx;
    const size_t bytes = (size_t)a->n * a->n * sizeof(c128);
    const int chunk_mib = c->opt.ame_host_chunk_mib > 0 ? c->opt.ame_host_chunk_mib : 128;
    int chunk = (int)(((size_t)chunk_mib << 20) / bytes);
    if (chunk < 1) chunk = 1;
    return host_pipeline(c, nb, bytes, bytes, h_u, h_du, chunk, [&](int, int cnt, const void* din, void* dout) {
        return ame_rhs_impl(a, cnt, (const c128*)din, (c128*)dout);
    });
}

int oqb_ame_eig_acc(oqb_ame* a, int nb, const void* d_rho_eig, void* d_du_eig, int with_hamiltonian) {
    OQB_DEVICE(a, a->ctx);
    oqb_ctx* c = a->ctx;
    if (nb <= 0) return 0;
    if (!a->have_eigen) return set_error(c, OQB_EINVAL, "oqb_ame_eig_acc: call oqb_ame_set_eigen first");
    const size_t ll = (size_t)a->lvl * a->lvl;
    const size_t per = (ll * 0 + eig_scratch_elems(a, 1)) * sizeof(c128);
    int chunk = (int)(kAmeBudget / (per ? per : 1));
    if (chunk < 1) chunk = 1;
    if (chunk > nb) chunk = nb;
    chunk = (nb + (nb + chunk - 1) / chunk - 1) / ((nb + chunk - 1) / chunk);  // balance the chunks
    OQB_TRY(ws_reserve(c, (size_t)chunk * per + 16));
    for (int b0 = 0; b0 < nb; b0 += chunk) {
        const int cnt = nb - b0 < chunk ? nb - b0 : chunk;
        OQB_TRY(eig_apply(a, cnt, (const c128*)d_rho_eig + (size_t)b0 * ll, (c128*)d_du_eig + (size_t)b0 * ll, false,
                          with_hamiltonian != 0, (c128*)c->ws));
    }
    return 0;
}

int oqb_ame_update_cache(oqb_ame* a, void* d_cache) {
    OQB_DEVICE(a, a->ctx);
    oqb_ctx* c = a->ctx;
    if (!a->have_eigen) return set_error(c, OQB_EINVAL, "oqb_ame_update_cache: call oqb_ame_set_eigen first");
    if (!d_cache) return set_error(c, OQB_EINVAL, "oqb_ame_update_cache: null cache");
    if (a->has_corr) return set_error(c, OQB_EINVAL, "oqb_ame_update_cache: the reference defines no update_cache! for CorrelatedDaviesGenerator");
    const int n = a->n, l = a->lvl;
    const long long ll = (long long)l * l, ln = (long long)l * n;
    if (!a->has_v)
        return davies_heff_diag(c, l, a->d_w, a->davies ? a->d_Gam : nullptr, a->davies ? a->d_hls : nullptr,
                                a->d_Moff, (c128*)d_cache);
    OQB_TRY(ws_reserve(c, (size_t)(ll + ln) * sizeof(c128)));
    c128* D = (c128*)c->ws;
    c128* T = D + ll;
    OQB_TRY(davies_heff_diag(c, l, a->d_w, a->davies ? a->d_Gam : nullptr, a->davies ? a->d_hls : nullptr, a->d_Moff, D));
    ZgemmParams p;  // T = D v'   (l × n)
    p.M = l; p.N = n; p.K = l;
    p.A = D; p.lda = l;
    p.B = a->d_V; p.ldb = n; p.opB = OP_C;
    p.real_operand = a->v_real ? 2 : 0;
    p.C = T; p.ldc = l;
    OQB_TRY(zgemm(c, p));
    ZgemmParams q;  // cache = v T   (src/opensys/diffeq_liouvillian.jl:124)
    q.M = n; q.N = n; q.K = l;
    q.A = a->d_V; q.lda = n;
    q.real_operand = a->v_real ? 1 : 0;
    q.B = T; q.ldb = l;
    q.C = (c128*)d_cache; q.ldc = n;
    return zgemm(c, q);
}

}  // extern "C"
