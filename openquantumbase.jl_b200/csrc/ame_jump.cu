// ame_jump for DaviesGenerator trajectories — lindblad_jump(Op::DiffEqLiouvillian{true,false}, u, p, t),
// reference src/opensys/trajectory_jump.jl:14-51 (ame_jump) and :64-69, on a batch of state vectors.
//
// The reference builds, per call and per trajectory, one sparse operator per (gap group, coupling, sign):
//   L₊ = sparse(a, b, σ_i[a,b]),  L₋ = sparse(b, a, σ_i[b,a]),  σ_i = v'A_i v,  prob = γ(±ω)·‖L ϕ‖²,  ϕ = v'ψ,
// then the zero-gap group, samples one by StatsBase's inverse-CDF scan and returns √γ·v L v'.  Here the index
// structure (the "probability slots", in the reference's order) is uploaded once per eigen-system as a flat term
// list; per call the device does
//   1. ZGEMM   Φ = V'Ψ                                   (lvl × nb)
//   2. ame_jump_prob_kernel    prob[b][slot] = γ_slot · Σ_rows |Σ_terms σ_i[row,col] ϕ_b[col]|²
//   3. ame_jump_choose_kernel  left-to-right total, inverse-CDF scan with the supplied uniform (one thread per member,
//                              the order of the scan IS the semantics), and Φ'[b] = L_choice ϕ_b
//   4. ZGEMM   Ψ' = V Φ',  ame_jump_apply_kernel  ψ_b ← ψ'_b/‖ψ'_b‖ for the jumping members (√γ cancels)
#include <vector>

#include "../../include/oqb.h"
#include "ame.cuh"
#include "common.cuh"

using namespace oqb;

namespace {

template <class T>
void dfree(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

__device__ __forceinline__ c128 cmul(c128 a, c128 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// grid (nb, ceil(nslots/128)): one thread per (member, slot)
__global__ void ame_jump_prob_kernel(int l, int nslots, const long long* __restrict__ slot_ptr, const int32_t* __restrict__ term,
                                     const double* __restrict__ rate, const c128* __restrict__ A, const c128* __restrict__ Phi,
                                     const uint8_t* __restrict__ mask, double* __restrict__ prob, int s0) {
    const int b = blockIdx.x;
    const int s = s0 + blockIdx.y * blockDim.x + threadIdx.x;
    if (s >= nslots || (mask && !mask[b])) return;
    const c128* phi = Phi + (size_t)b * l;
    const long long ll = (long long)l * l;
    double tot = 0.0;
    c128 acc = make_double2(0.0, 0.0);
    int cur = -1;
    for (long long k = slot_ptr[s]; k < slot_ptr[s + 1]; ++k) {
        const int i = term[3 * k], r = term[3 * k + 1], c = term[3 * k + 2];
        if (r != cur) {
            tot += acc.x * acc.x + acc.y * acc.y;
            acc = make_double2(0.0, 0.0);
            cur = r;
        }
        const c128 v = cmul(A[i * ll + r + (long long)c * l], phi[c]);
        acc.x += v.x;
        acc.y += v.y;
    }
    tot += acc.x * acc.x + acc.y * acc.y;
    prob[(size_t)b * nslots + s] = rate[s] * tot;
}

// one thread per member: StatsBase.sample(Weights(prob)) with the supplied uniform, then Φ'[b] = L_choice ϕ_b
__global__ void ame_jump_choose_kernel(int l, int nb, int nslots, const long long* __restrict__ slot_ptr,
                                       const int32_t* __restrict__ term, const c128* __restrict__ A, const c128* __restrict__ Phi,
                                       const double* __restrict__ prob, const double* __restrict__ uniform,
                                       const uint8_t* __restrict__ mask, int32_t* __restrict__ choice, c128* __restrict__ PhiNew) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    if (mask && !mask[b]) {
        if (choice) choice[b] = -1;
        return;
    }
    const double* p = prob + (size_t)b * nslots;
    double total = 0.0;
    for (int s = 0; s < nslots; ++s) total += p[s];
    const double t = uniform[b] * total;
    int i = 0;
    double cw = p[0];
    while (cw < t && i < nslots - 1) {
        ++i;
        cw += p[i];
    }
    if (choice) choice[b] = i;
    if (!PhiNew) return;
    c128* out = PhiNew + (size_t)b * l;
    const c128* phi = Phi + (size_t)b * l;
    for (int r = 0; r < l; ++r) out[r] = make_double2(0.0, 0.0);
    const long long ll = (long long)l * l;
    for (long long k = slot_ptr[i]; k < slot_ptr[i + 1]; ++k) {
        const int ci = term[3 * k], r = term[3 * k + 1], c = term[3 * k + 2];
        const c128 v = cmul(A[ci * ll + r + (long long)c * l], phi[c]);
        out[r].x += v.x;
        out[r].y += v.y;
    }
}

// one block per member: ψ_b ← ψ'_b/‖ψ'_b‖ where mask[b] (fixed-order block reduction)
__global__ void ame_jump_apply_kernel(int n, const c128* __restrict__ PsiNew, const uint8_t* __restrict__ mask, c128* __restrict__ Psi) {
    const int b = blockIdx.x;
    if (mask && !mask[b]) return;
    const c128* src = PsiNew + (size_t)b * n;
    c128* dst = Psi + (size_t)b * n;
    __shared__ double sh[256];
    double acc = 0.0;
    for (int e = threadIdx.x; e < n; e += blockDim.x) acc += src[e].x * src[e].x + src[e].y * src[e].y;
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    const double inv = 1.0 / sqrt(sh[0]);
    for (int e = threadIdx.x; e < n; e += blockDim.x) dst[e] = make_double2(src[e].x * inv, src[e].y * inv);
}

}  // namespace

extern "C" {

int oqb_ame_clear_jump(oqb_ame* a) {
    OQB_DEVICE(a, a->ctx);
    if (!a) return 0;
    cudaStreamSynchronize(a->ctx->stream);
    dfree(a->d_slot_ptr); dfree(a->d_term); dfree(a->d_slot_rate);
    a->nslots = 0; a->nterms = 0;
    return 0;
}

int oqb_ame_set_jump(oqb_ame* a, int nslots, const int64_t* h_slot_ptr, const int32_t* h_term, const double* h_rate) {
    OQB_DEVICE(a, a->ctx);
    if (!a) return OQB_EINVAL;
    oqb_ctx* c = a->ctx;
    if (nslots < 1 || !h_slot_ptr || !h_term || !h_rate) return set_error(c, OQB_EINVAL, "oqb_ame_set_jump: bad argument");
    const long long nterms = h_slot_ptr[nslots];
    for (long long k = 0; k < nterms; ++k)
        if (h_term[3 * k] < 0 || h_term[3 * k] >= a->K || h_term[3 * k + 1] < 0 || h_term[3 * k + 1] >= a->lvl ||
            h_term[3 * k + 2] < 0 || h_term[3 * k + 2] >= a->lvl)
            return set_error(c, OQB_EINVAL, "oqb_ame_set_jump: term %lld out of range", k);
    oqb_ame_clear_jump(a);
    OQB_CUDA(c, cudaMalloc((void**)&a->d_slot_ptr, (size_t)(nslots + 1) * 8));
    OQB_CUDA(c, cudaMalloc((void**)&a->d_term, (size_t)(nterms ? nterms : 1) * 12));
    OQB_CUDA(c, cudaMalloc((void**)&a->d_slot_rate, (size_t)nslots * 8));
    OQB_CUDA(c, cudaMemcpy(a->d_slot_ptr, h_slot_ptr, (size_t)(nslots + 1) * 8, cudaMemcpyHostToDevice));
    if (nterms) OQB_CUDA(c, cudaMemcpy(a->d_term, h_term, (size_t)nterms * 12, cudaMemcpyHostToDevice));
    OQB_CUDA(c, cudaMemcpy(a->d_slot_rate, h_rate, (size_t)nslots * 8, cudaMemcpyHostToDevice));
    a->nslots = nslots;
    a->nterms = nterms;
    return 0;
}

int oqb_ame_jump(oqb_ame* a, int nb, void* d_psi, const double* d_uniform, const uint8_t* d_mask, int32_t* d_choice,
                 double* d_prob, int apply) {
    OQB_DEVICE(a, a->ctx);
    if (!a) return OQB_EINVAL;
    oqb_ctx* c = a->ctx;
    if (nb <= 0) return 0;
    if (!a->have_eigen || a->nslots == 0) return set_error(c, OQB_EINVAL, "oqb_ame_jump: call oqb_ame_set_eigen and oqb_ame_set_jump first");
    if (!d_psi || !d_uniform) return set_error(c, OQB_EINVAL, "oqb_ame_jump: null argument");
    const int n = a->n, l = a->lvl, ns = a->nslots;
    // scratch: Φ (l×nb), Φ' (l×nb), Ψ' (n×nb), prob (nb×nslots unless supplied)
    const size_t e_phi = (size_t)l * nb, e_psi = (size_t)n * nb;
    const size_t bytes = (2 * e_phi + e_psi) * sizeof(c128) + (d_prob ? 0 : (size_t)nb * ns * sizeof(double));
    OQB_TRY(ws_reserve(c, bytes));
    c128* Phi = (c128*)c->ws;
    c128* PhiNew = Phi + e_phi;
    c128* PsiNew = PhiNew + e_phi;
    double* prob = d_prob ? d_prob : (double*)(PsiNew + e_psi);
    c128* Psi = (c128*)d_psi;
    if (a->has_v) {  // Φ = V'Ψ
        ZgemmParams p;
        p.M = l; p.N = nb; p.K = n; p.opA = OP_C;
        p.A = a->d_V; p.lda = n;
        p.real_operand = a->v_real ? 1 : 0;
        p.B = Psi; p.ldb = n;
        p.C = Phi; p.ldc = l;
        OQB_TRY(zgemm(c, p));
    } else {
        OQB_CUDA(c, cudaMemcpyAsync(Phi, Psi, e_phi * sizeof(c128), cudaMemcpyDeviceToDevice, c->stream));
    }
    if (apply) OQB_CUDA(c, cudaMemsetAsync(PhiNew, 0, e_phi * sizeof(c128), c->stream));
    for (int s0 = 0; s0 < ns; s0 += 65535 * 128) {  // gridDim.y ≤ 65535
        const int cnt = ns - s0 < 65535 * 128 ? ns - s0 : 65535 * 128;
        dim3 grid(nb, (cnt + 127) / 128);
        ame_jump_prob_kernel<<<grid, 128, 0, c->stream>>>(l, ns, a->d_slot_ptr, a->d_term, a->d_slot_rate, a->d_A, Phi, d_mask, prob, s0);
        OQB_LAUNCH_CHECK(c);
    }
    ame_jump_choose_kernel<<<(nb + 63) / 64, 64, 0, c->stream>>>(l, nb, ns, a->d_slot_ptr, a->d_term, a->d_A, Phi, prob, d_uniform,
                                                              d_mask, d_choice, apply ? PhiNew : nullptr);
    OQB_LAUNCH_CHECK(c);
    if (apply) {
        const c128* src = PhiNew;
        if (a->has_v) {  // Ψ' = V Φ'   (columns of members that do not jump are zero and never read)
            ZgemmParams p;
            p.M = n; p.N = nb; p.K = l;
            p.A = a->d_V; p.lda = n;
            p.real_operand = a->v_real ? 1 : 0;
            p.B = PhiNew; p.ldb = l;
            p.C = PsiNew; p.ldc = n;
            OQB_TRY(zgemm(c, p));
            src = PsiNew;
        }
        ame_jump_apply_kernel<<<nb, 256, 0, c->stream>>>(n, src, d_mask, Psi);
        OQB_LAUNCH_CHECK(c);
    }
    return 0;
}

}  // extern "C"
