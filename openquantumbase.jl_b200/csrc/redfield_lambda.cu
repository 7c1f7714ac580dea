// Device builder of the Redfield kernel matrices Λ(t).
//
// Replaces the `quadgk!` call inside (R::RedfieldLiouvillian)(du,u,p,t), reference
// src/opensys/redfield.jl:46-64 (and the identical one in update_vectorized_cache!, :76-96):
//
//     Λ_ij(t) = ∫_{max(0,t−Ta)}^{t} C_ij(t,x) · U(t)U(x)† S_j(p(x)) U(x)U(t)† dx
//
// for MANY integrals at once (schedules × coupling pairs).  The adaptive Gauss–Kronrod (7,15) control flow of
// QuadGK (one max-heap of segments per integral, bisect the worst one, stop when E ≤ atol or E ≤ rtol‖I‖) stays
// on the host, which also evaluates the scalar closures C_ij(t,x) at the nodes; every refinement round is ONE call
// of oqb_lambda_eval that applies the 15-point rule to a list of segments.  Per segment the device does
//
//   1. lambda_stack_kernel   U(x_i) for the 15 nodes (piecewise-linear interpolant of the member's sampled unitary)
//                            written as ONE tall matrix  A = [U(x_1); …; U(x_15)]  (15n × n), plus U(t)
//   2. (dense S_j only)      Y_i = S_j U(x_i), a batched ZGEMM over the nodes
//      lambda_weight_kernel  (diagonal S_j: fused into step 1, lambda_stack_weight_kernel)
//                            Bk = [ck_i·Y_i]_{i<15} (15n × n),  Bg = [cg_i·Y_i]_{i<7} (7n × n),  ck_i = wk_i·h·C(t,x_i),
//                            cg_i = wg_i·h·C(t,x_i); the nodes come Gauss-first, so the Gauss rule is the first
//                            7 block rows of the same stack (diagonal S_j: Y_i = diag(s_j)·U(x_i) fused here)
//   3. ZGEMM  Ik = A† Bk  (K-dimension 15n),  ZGEMM  Ig = A[:7n]† Bg  (K-dimension 7n): the weighted node sums ride
//                            the GEMMs' k loops — U(x)† S U(x) is never materialised per node
//   4. ZGEMM  T = U(t)[Ik | Ig],  ZGEMM  R = T_half U(t)†   → Kronrod integral R[:, :n],  Gauss integral R[:, n:]
//   5. lambda_finish_kernel  slot[s] = R[:, :n],  E_seg = ‖R[:, :n] − R[:, n:]‖_F   (fixed-order block reduction)
//
// U(t) is pulled out of the node loop (U(t)·M·U(t)† is linear in M), which turns the reference's 3 GEMMs per node
// (redfield.jl:49-54) into 15 + 7 node products + 4 per segment: 8n³·26 flops per segment (41 when S_j is dense)
// instead of 8n³·45.  oqb_lambda_total then sums each integral's leaf segments in heap order (QuadGK's final re-summation)
// and returns ‖Λ‖_F for the convergence test.
#include <cmath>
#include <cstring>

#include "../../include/oqb.h"
#include "common.cuh"

using namespace oqb;

struct oqb_lambda {
    Ctx* c = nullptr;
    int n = 0, K = 0;
    bool diag = false;
    c128* dS = nullptr;      // K n×n couplings
    c128* dSdiag = nullptr;  // K×n diagonals when every coupling is diagonal
    const c128* dU = nullptr;  // borrowed: nb × G sampled unitaries (n×n each)
    int nb = 0, G = 0;
    c128* slots = nullptr;  // leaf-segment integrals, n×n each
    long long nslots = 0;
    // per-call staging (device)
    void* dmeta = nullptr;
    size_t dmeta_bytes = 0;
};

namespace {

constexpr int NODES = 15, GNODES = 7;  // Kronrod nodes; the first GNODES of them are the Gauss nodes
constexpr size_t kLambdaWorkspace = (size_t)3 << 30;

// One block per (segment, node); node 15 is the end time t of the segment's integral.
__global__ void lambda_stack_kernel(int n, int G, const c128* __restrict__ U, const int32_t* __restrict__ member,
                                    const int32_t* __restrict__ gi, const double* __restrict__ theta, c128* __restrict__ A,
                                    c128* __restrict__ Ut) {
    const int seg = blockIdx.x / (NODES + 1), node = blockIdx.x % (NODES + 1);
    const long long nn = (long long)n * n;
    const int g = gi[seg * (NODES + 1) + node];
    const double th = theta[seg * (NODES + 1) + node];
    const c128* u0 = U + ((long long)member[seg] * G + g) * nn;
    const c128* u1 = u0 + nn;
    const bool two = th != 0.0;  // θ == 0 never touches the next sample (the last grid point has none)
    const double om = 1.0 - th;
    for (long long e = threadIdx.x; e < nn; e += blockDim.x) {
        c128 a = u0[e];
        if (two) {
            const c128 b = u1[e];
            a.x = om * a.x + th * b.x;
            a.y = om * a.y + th * b.y;
        }
        if (node < NODES) {
            const int r = (int)(e % n), col = (int)(e / n);
            A[(long long)seg * NODES * nn + (long long)col * NODES * n + node * n + r] = a;
        } else {
            Ut[(long long)seg * nn + e] = a;
        }
    }
}

// Diagonal couplings: the same interpolation, fused with the weighting — writes A, Bk = ck_i·diag(s_j)·U(x_i) and (Gauss
// nodes) Bg = cg_i·diag(s_j)·U(x_i) in one pass, so the stack is never re-read by an elementwise kernel.
__global__ void lambda_stack_weight_kernel(int n, int G, const c128* __restrict__ U, const int32_t* __restrict__ member,
                                           const int32_t* __restrict__ gi, const double* __restrict__ theta,
                                           const c128* __restrict__ sdiag, const int32_t* __restrict__ coupling,
                                           const c128* __restrict__ ck, const c128* __restrict__ cg, c128* __restrict__ A,
                                           c128* __restrict__ Bk, c128* __restrict__ Bg, c128* __restrict__ Ut) {
    const int seg = blockIdx.x / (NODES + 1), node = blockIdx.x % (NODES + 1);
    const long long nn = (long long)n * n;
    const int g = gi[seg * (NODES + 1) + node];
    const double th = theta[seg * (NODES + 1) + node];
    const c128* u0 = U + ((long long)member[seg] * G + g) * nn;
    const c128* u1 = u0 + nn;
    const bool two = th != 0.0;
    const double om = 1.0 - th;
    const bool gauss = node < GNODES;
    const c128 wk = node < NODES ? ck[seg * NODES + node] : make_double2(0.0, 0.0);
    const c128 wg = gauss ? cg[seg * GNODES + node] : make_double2(0.0, 0.0);
    const c128* sd = sdiag + (long long)coupling[seg] * n;
    for (long long e = threadIdx.x; e < nn; e += blockDim.x) {
        c128 a = u0[e];
        if (two) {
            const c128 b = u1[e];
            a.x = om * a.x + th * b.x;
            a.y = om * a.y + th * b.y;
        }
        if (node < NODES) {
            const int r = (int)(e % n), col = (int)(e / n);
            A[(long long)seg * NODES * nn + (long long)col * NODES * n + node * n + r] = a;
            const c128 s = sd[r];
            const c128 v = make_double2(s.x * a.x - s.y * a.y, s.x * a.y + s.y * a.x);
            Bk[(long long)seg * NODES * nn + (long long)col * NODES * n + node * n + r] =
                make_double2(wk.x * v.x - wk.y * v.y, wk.x * v.y + wk.y * v.x);
            if (gauss)
                Bg[(long long)seg * GNODES * nn + (long long)col * GNODES * n + node * n + r] =
                    make_double2(wg.x * v.x - wg.y * v.y, wg.x * v.y + wg.y * v.x);
        } else {
            Ut[(long long)seg * nn + e] = a;
        }
    }
}

// Bk[seg] = [ck_i·Y_i]_{i<15}, Bg[seg] = [cg_i·Y_i]_{i<7} with Y from the Y-stack (dense couplings) or diag(s_j)·A-stack.
__global__ void lambda_weight_kernel(int n, const c128* __restrict__ Y, const c128* __restrict__ sdiag,
                                     const int32_t* __restrict__ coupling, const c128* __restrict__ ck,
                                     const c128* __restrict__ cg, c128* __restrict__ Bk, c128* __restrict__ Bg) {
    const int seg = blockIdx.x / NODES, node = blockIdx.x % NODES;
    const long long nn = (long long)n * n;
    const c128 wk = ck[seg * NODES + node];
    const bool gauss = node < GNODES;
    const c128 wg = gauss ? cg[seg * GNODES + node] : make_double2(0.0, 0.0);
    const c128* y = Y + (long long)seg * NODES * nn + node * n;
    c128* bk = Bk + (long long)seg * NODES * nn + node * n;
    c128* bg = Bg + (long long)seg * GNODES * nn + node * n;
    const c128* sd = sdiag ? sdiag + (long long)coupling[seg] * n : nullptr;
    for (long long e = threadIdx.x; e < nn; e += blockDim.x) {
        const int r = (int)(e % n), col = (int)(e / n);
        c128 v = y[(long long)col * NODES * n + r];
        if (sd) {
            const c128 s = sd[r];
            v = make_double2(s.x * v.x - s.y * v.y, s.x * v.y + s.y * v.x);
        }
        bk[(long long)col * NODES * n + r] = make_double2(wk.x * v.x - wk.y * v.y, wk.x * v.y + wk.y * v.x);
        if (gauss) bg[(long long)col * GNODES * n + r] = make_double2(wg.x * v.x - wg.y * v.y, wg.x * v.y + wg.y * v.x);
    }
}

// slot[slot_id[seg]] = R[seg][:, :n] (Kronrod);  E[seg] = ‖R[seg][:, :n] − R[seg][:, n:]‖_F (Kronrod − Gauss)  — one block
// per segment, fixed reduction order.
__global__ void lambda_finish_kernel(int n, const c128* __restrict__ R, const long long* __restrict__ slot_id,
                                     c128* __restrict__ slots, double* __restrict__ E) {
    const int seg = blockIdx.x;
    const long long nn = (long long)n * n;
    const c128* r = R + (long long)seg * 2 * nn;
    c128* dst = slots + slot_id[seg] * nn;
    double acc = 0.0;
    for (long long e = threadIdx.x; e < nn; e += blockDim.x) {
        const c128 k = r[e], g = r[nn + e];
        dst[e] = k;
        const double dx = k.x - g.x, dy = k.y - g.y;
        acc += dx * dx + dy * dy;
    }
    __shared__ double sh[256];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) E[seg] = sqrt(sh[0]);
}

// Λ_q = Σ_k slot[idx[k]], k ∈ [ptr[q], ptr[q+1]) in order;  nrm[q] = ‖Λ_q‖_F
__global__ void lambda_total_kernel(int n, const long long* __restrict__ ptr, const long long* __restrict__ idx,
                                    const c128* __restrict__ slots, c128* __restrict__ Lam, double* __restrict__ nrm) {
    const int q = blockIdx.x;
    const long long nn = (long long)n * n;
    const long long k0 = ptr[q], k1 = ptr[q + 1];
    double acc = 0.0;
    for (long long e = threadIdx.x; e < nn; e += blockDim.x) {
        c128 s = make_double2(0.0, 0.0);
        for (long long k = k0; k < k1; ++k) {
            const c128 v = slots[idx[k] * nn + e];
            if (k == k0) s = v;
            else { s.x += v.x; s.y += v.y; }
        }
        Lam[(long long)q * nn + e] = s;
        acc += s.x * s.x + s.y * s.y;
    }
    __shared__ double sh[256];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) nrm[q] = sqrt(sh[0]);
}

int meta_reserve(oqb_lambda* L, size_t bytes) {
    if (bytes <= L->dmeta_bytes) return 0;
    Ctx* c = L->c;
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (L->dmeta) cudaFree(L->dmeta);
    L->dmeta = nullptr;
    L->dmeta_bytes = 0;
    const size_t want = bytes + bytes / 2 + 4096;
    OQB_CUDA(c, cudaMalloc(&L->dmeta, want));
    L->dmeta_bytes = want;
    return 0;
}

size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

}  // namespace

extern "C" {

int oqb_lambda_create(oqb_ctx* c, int n, int K, const void* h_S, oqb_lambda** out) {
    OQB_DEVICE(c, c);
    if (!c || !out || !h_S || n < 1 || K < 1) return c ? set_error(c, OQB_EINVAL, "oqb_lambda_create: bad argument") : OQB_EINVAL;
    oqb_lambda* L = new oqb_lambda();
    L->c = c; L->n = n; L->K = K;
    const long long nn = (long long)n * n;
    const c128* S = (const c128*)h_S;
    bool diag = true;
    for (long long a = 0; a < K && diag; ++a)
        for (long long e = 0; e < nn && diag; ++e)
            if (e % n != e / n && (S[a * nn + e].x != 0.0 || S[a * nn + e].y != 0.0)) diag = false;
    L->diag = diag;
    if (cudaMalloc(&L->dS, (size_t)K * nn * sizeof(c128)) != cudaSuccess) { delete L; return set_error(c, OQB_ENOMEM, "oqb_lambda_create: cudaMalloc failed"); }
    cudaMemcpy(L->dS, S, (size_t)K * nn * sizeof(c128), cudaMemcpyHostToDevice);
    if (diag) {
        std::vector<c128> d((size_t)K * n);
        for (int a = 0; a < K; ++a)
            for (int r = 0; r < n; ++r) d[(size_t)a * n + r] = S[a * nn + (long long)r * n + r];
        if (cudaMalloc(&L->dSdiag, d.size() * sizeof(c128)) != cudaSuccess) { cudaFree(L->dS); delete L; return set_error(c, OQB_ENOMEM, "oqb_lambda_create: cudaMalloc failed"); }
        cudaMemcpy(L->dSdiag, d.data(), d.size() * sizeof(c128), cudaMemcpyHostToDevice);
    }
    *out = L;
    return 0;
}

int oqb_lambda_dims(const oqb_lambda* L, int* n, int* K) {
    if (!L) return OQB_EINVAL;
    if (n) *n = L->n;
    if (K) *K = L->K;
    return 0;
}

int oqb_lambda_destroy(oqb_lambda* L) {
    OQB_DEVICE(L, L->c);
    if (!L) return 0;
    cudaStreamSynchronize(L->c->stream);
    if (L->dS) cudaFree(L->dS);
    if (L->dSdiag) cudaFree(L->dSdiag);
    if (L->slots) cudaFree(L->slots);
    if (L->dmeta) cudaFree(L->dmeta);
    delete L;
    return 0;
}

int oqb_lambda_set_unitary(oqb_lambda* L, int nb, int G, const void* d_Ugrid) {
    OQB_DEVICE(L, L->c);
    if (!L) return OQB_EINVAL;
    if (nb < 1 || G < 1 || !d_Ugrid) return set_error(L->c, OQB_EINVAL, "oqb_lambda_set_unitary: bad argument");
    L->nb = nb; L->G = G; L->dU = (const c128*)d_Ugrid;
    return 0;
}

int oqb_lambda_reserve_slots(oqb_lambda* L, int64_t nslots) {
    OQB_DEVICE(L, L->c);
    if (!L) return OQB_EINVAL;
    Ctx* c = L->c;
    if (nslots <= L->nslots) return 0;
    const long long nn = (long long)L->n * L->n;
    long long want = nslots + nslots / 2;
    c128* fresh = nullptr;
    cudaError_t e = cudaMalloc(&fresh, (size_t)want * nn * sizeof(c128));
    if (e != cudaSuccess) {  // retry without head-room
        want = nslots;
        e = cudaMalloc(&fresh, (size_t)want * nn * sizeof(c128));
    }
    if (e != cudaSuccess) return set_error(c, OQB_ENOMEM, "oqb_lambda_reserve_slots: cudaMalloc(%lld slots) failed", want);
    if (L->slots) {
        OQB_CUDA(c, cudaMemcpyAsync(fresh, L->slots, (size_t)L->nslots * nn * sizeof(c128), cudaMemcpyDeviceToDevice, c->stream));
        OQB_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(L->slots);
    }
    L->slots = fresh;
    L->nslots = want;
    return 0;
}

int oqb_lambda_eval(oqb_lambda* L, int nseg, const int32_t* h_member, const int32_t* h_coupling, const int32_t* h_gi,
                    const double* h_theta, const double* h_ck, const double* h_cg, const int64_t* h_slot, double* h_E) {
    OQB_DEVICE(L, L->c);
    if (!L) return OQB_EINVAL;
    Ctx* c = L->c;
    if (nseg <= 0) return 0;
    if (!L->dU) return set_error(c, OQB_EINVAL, "oqb_lambda_eval: call oqb_lambda_set_unitary first");
    if (!h_member || !h_coupling || !h_gi || !h_theta || !h_ck || !h_cg || !h_slot || !h_E)
        return set_error(c, OQB_EINVAL, "oqb_lambda_eval: null argument");
    const int n = L->n;
    const long long nn = (long long)n * n;
    for (int s = 0; s < nseg; ++s) {
        if (h_member[s] < 0 || h_member[s] >= L->nb || h_coupling[s] < 0 || h_coupling[s] >= L->K || h_slot[s] < 0 ||
            h_slot[s] >= L->nslots)
            return set_error(c, OQB_EINVAL, "oqb_lambda_eval: segment %d has member/coupling/slot out of range", s);
        for (int i = 0; i <= NODES; ++i) {
            const int g = h_gi[s * (NODES + 1) + i];
            const double th = h_theta[s * (NODES + 1) + i];
            if (g < 0 || g >= L->G || (th != 0.0 && g + 1 >= L->G) || th < 0.0 || th >= 1.0)
                return set_error(c, OQB_EINVAL, "oqb_lambda_eval: segment %d node %d lies outside the sampled unitary", s, i);
        }
    }
    // ---- stage the per-segment metadata ----
    const size_t o_member = 0;
    const size_t o_coupling = align16(o_member + (size_t)nseg * 4);
    const size_t o_gi = align16(o_coupling + (size_t)nseg * 4);
    const size_t o_theta = align16(o_gi + (size_t)nseg * (NODES + 1) * 4);
    const size_t o_ck = align16(o_theta + (size_t)nseg * (NODES + 1) * 8);
    const size_t o_cg = align16(o_ck + (size_t)nseg * NODES * 16);
    const size_t o_slot = align16(o_cg + (size_t)nseg * GNODES * 16);
    const size_t o_E = align16(o_slot + (size_t)nseg * 8);
    const size_t total = align16(o_E + (size_t)nseg * 8);
    OQB_TRY(meta_reserve(L, total));
    char* dm = (char*)L->dmeta;
    auto up = [&](size_t off, const void* src, size_t bytes) {
        return cudaMemcpyAsync(dm + off, src, bytes, cudaMemcpyHostToDevice, c->stream);
    };
    OQB_CUDA(c, up(o_member, h_member, (size_t)nseg * 4));
    OQB_CUDA(c, up(o_coupling, h_coupling, (size_t)nseg * 4));
    OQB_CUDA(c, up(o_gi, h_gi, (size_t)nseg * (NODES + 1) * 4));
    OQB_CUDA(c, up(o_theta, h_theta, (size_t)nseg * (NODES + 1) * 8));
    OQB_CUDA(c, up(o_ck, h_ck, (size_t)nseg * NODES * 16));
    OQB_CUDA(c, up(o_cg, h_cg, (size_t)nseg * GNODES * 16));
    OQB_CUDA(c, up(o_slot, h_slot, (size_t)nseg * 8));
    const int32_t* d_member = (const int32_t*)(dm + o_member);
    const int32_t* d_coupling = (const int32_t*)(dm + o_coupling);
    const int32_t* d_gi = (const int32_t*)(dm + o_gi);
    const double* d_theta = (const double*)(dm + o_theta);
    const c128* d_ck = (const c128*)(dm + o_ck);
    const c128* d_cg = (const c128*)(dm + o_cg);
    const long long* d_slot = (const long long*)(dm + o_slot);
    double* d_E = (double*)(dm + o_E);

    // ---- workspace per segment: A (15nn) [+ Y (15nn)] + Bk (15nn) + Bg (7nn) + [Ik|Ig] (2nn) + T (2nn) + Ut (nn) ----
    const size_t per = (size_t)((L->diag ? 15 : 30) + 15 + 7 + 2 + 2 + 1) * nn * sizeof(c128);
    int chunk = (int)(kLambdaWorkspace / per);
    if (chunk < 1) chunk = 1;
    if (chunk > nseg) chunk = nseg;
    OQB_TRY(ws_reserve(c, (size_t)chunk * per));
    c128* A = (c128*)c->ws;
    c128* Y = L->diag ? A : A + (size_t)chunk * NODES * nn;
    c128* Bk = Y + (size_t)chunk * NODES * nn;
    c128* Bg = Bk + (size_t)chunk * NODES * nn;
    c128* IkD = Bg + (size_t)chunk * GNODES * nn;
    c128* T = IkD + (size_t)chunk * 2 * nn;
    c128* Ut = T + (size_t)chunk * 2 * nn;
    const int threads = nn >= 256 ? 256 : 64;
    for (int s0 = 0; s0 < nseg; s0 += chunk) {
        const int cnt = nseg - s0 < chunk ? nseg - s0 : chunk;
        if (L->diag) {
            lambda_stack_weight_kernel<<<cnt * (NODES + 1), threads, 0, c->stream>>>(
                n, L->G, L->dU, d_member + s0, d_gi + (size_t)s0 * (NODES + 1), d_theta + (size_t)s0 * (NODES + 1), L->dSdiag,
                d_coupling + s0, d_ck + (size_t)s0 * NODES, d_cg + (size_t)s0 * GNODES, A, Bk, Bg, Ut);
            OQB_LAUNCH_CHECK(c);
        } else {
            lambda_stack_kernel<<<cnt * (NODES + 1), threads, 0, c->stream>>>(n, L->G, L->dU, d_member + s0,
                                                                             d_gi + (size_t)s0 * (NODES + 1),
                                                                             d_theta + (size_t)s0 * (NODES + 1), A, Ut);
            OQB_LAUNCH_CHECK(c);
            // Y_i = S_j U(x_i): one batched product per run of equal coupling index
            int r0 = 0;
            while (r0 < cnt) {
                int r1 = r0 + 1;
                while (r1 < cnt && h_coupling[s0 + r1] == h_coupling[s0 + r0]) ++r1;
                ZgemmParams p;
                p.M = p.N = p.K = n; p.batch = r1 - r0; p.batch2 = NODES;
                p.A = L->dS + (long long)h_coupling[s0 + r0] * nn; p.lda = n; p.strideA = 0; p.strideA2 = 0;
                p.B = A + (long long)r0 * NODES * nn; p.ldb = (long long)NODES * n; p.strideB = NODES * nn; p.strideB2 = n;
                p.C = Y + (long long)r0 * NODES * nn; p.ldc = (long long)NODES * n; p.strideC = NODES * nn; p.strideC2 = n;
                OQB_TRY(zgemm(c, p));
                r0 = r1;
            }
            lambda_weight_kernel<<<cnt * NODES, threads, 0, c->stream>>>(n, Y, nullptr, d_coupling + s0, d_ck + (size_t)s0 * NODES,
                                                                        d_cg + (size_t)s0 * GNODES, Bk, Bg);
            OQB_LAUNCH_CHECK(c);
        }
        {   // Ik = A† Bk  (all 15 nodes)
            ZgemmParams p;
            p.M = n; p.N = n; p.K = NODES * n; p.batch = cnt; p.opA = OP_C;
            p.A = A; p.lda = (long long)NODES * n; p.strideA = NODES * nn;
            p.B = Bk; p.ldb = (long long)NODES * n; p.strideB = NODES * nn;
            p.C = IkD; p.ldc = n; p.strideC = 2 * nn;
            OQB_TRY(zgemm(c, p));
        }
        {   // Ig = A[:7n]† Bg  (the 7 Gauss nodes = the first 7 block rows of the stack)
            ZgemmParams p;
            p.M = n; p.N = n; p.K = GNODES * n; p.batch = cnt; p.opA = OP_C;
            p.A = A; p.lda = (long long)NODES * n; p.strideA = NODES * nn;
            p.B = Bg; p.ldb = (long long)GNODES * n; p.strideB = GNODES * nn;
            p.C = IkD + nn; p.ldc = n; p.strideC = 2 * nn;
            OQB_TRY(zgemm(c, p));
        }
        {   // T = U(t) [Ik | Ig]
            ZgemmParams p;
            p.M = n; p.N = 2 * n; p.K = n; p.batch = cnt;
            p.A = Ut; p.lda = n; p.strideA = nn;
            p.B = IkD; p.ldb = n; p.strideB = 2 * nn;
            p.C = T; p.ldc = n; p.strideC = 2 * nn;
            OQB_TRY(zgemm(c, p));
        }
        {   // R_half = T_half U(t)†  (written back over IkD)
            ZgemmParams p;
            p.M = p.N = p.K = n; p.batch = cnt; p.batch2 = 2; p.opB = OP_C;
            p.A = T; p.lda = n; p.strideA = 2 * nn; p.strideA2 = nn;
            p.B = Ut; p.ldb = n; p.strideB = nn; p.strideB2 = 0;
            p.C = IkD; p.ldc = n; p.strideC = 2 * nn; p.strideC2 = nn;
            OQB_TRY(zgemm(c, p));
        }
        lambda_finish_kernel<<<cnt, 256, 0, c->stream>>>(n, IkD, d_slot + s0, L->slots, d_E + s0);
        OQB_LAUNCH_CHECK(c);
    }
    OQB_CUDA(c, cudaMemcpyAsync(h_E, d_E, (size_t)nseg * 8, cudaMemcpyDeviceToHost, c->stream));
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int oqb_lambda_total(oqb_lambda* L, int nint, const int64_t* h_ptr, const int64_t* h_idx, void* d_Lambda, double* h_norm) {
    OQB_DEVICE(L, L->c);
    if (!L) return OQB_EINVAL;
    Ctx* c = L->c;
    if (nint <= 0) return 0;
    if (!h_ptr || !h_idx || !d_Lambda) return set_error(c, OQB_EINVAL, "oqb_lambda_total: null argument");
    const long long nidx = h_ptr[nint];
    for (long long k = 0; k < nidx; ++k)
        if (h_idx[k] < 0 || h_idx[k] >= L->nslots) return set_error(c, OQB_EINVAL, "oqb_lambda_total: slot index out of range");
    const size_t o_ptr = 0;
    const size_t o_idx = align16((size_t)(nint + 1) * 8);
    const size_t o_nrm = align16(o_idx + (size_t)nidx * 8);
    const size_t total = align16(o_nrm + (size_t)nint * 8);
    OQB_TRY(meta_reserve(L, total));
    char* dm = (char*)L->dmeta;
    OQB_CUDA(c, cudaMemcpyAsync(dm + o_ptr, h_ptr, (size_t)(nint + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    OQB_CUDA(c, cudaMemcpyAsync(dm + o_idx, h_idx, (size_t)nidx * 8, cudaMemcpyHostToDevice, c->stream));
    lambda_total_kernel<<<nint, 256, 0, c->stream>>>(L->n, (const long long*)(dm + o_ptr), (const long long*)(dm + o_idx),
                                                    L->slots, (c128*)d_Lambda, (double*)(dm + o_nrm));
    OQB_LAUNCH_CHECK(c);
    if (h_norm) {
        OQB_CUDA(c, cudaMemcpyAsync(h_norm, dm + o_nrm, (size_t)nint * 8, cudaMemcpyDeviceToHost, c->stream));
        OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return 0;
}

}  // extern "C"
