// Private layout of the eigenbasis-Liouvillian handle (shared by ame.cu and ame_jump.cu).
#pragma once
#include <vector>

#include "common.cuh"

struct oqb_ame {
    oqb_ctx* ctx = nullptr;
    int n = 0, lvl = 0, K = 0;
    bool has_v = false, have_eigen = false;
    bool v_real = false;  // every entry of V has zero imaginary part: rotations use the real-operand product
    bool lam_real = false, psum_real = false, ab_real = false;  // same for Λ_p, Σ A_βΛ_p and the A_β stack (one-sided AME)
    unsigned int* d_flag = nullptr;  // counter for the exact-realness checks
    oqb::c128* d_coup = nullptr;  // K n×n couplings as given
    oqb::c128* d_cdiag = nullptr; // K × n diagonals when EVERY coupling is a diagonal matrix (σz-type), else nullptr
    oqb::c128* d_V = nullptr;     // n × lvl
    double* d_w = nullptr;   // lvl
    oqb::c128* d_A = nullptr;     // K lvl×lvl rotated couplings v' A v
    oqb::c128* d_Cm = nullptr;    // lvl×lvl Hadamard coefficients (always includes −i(w_a−w_b))
    // Davies
    bool davies = false;
    double *d_gam = nullptr, *d_S = nullptr, *d_Gam = nullptr, *d_hls = nullptr, *d_Wt = nullptr;
    double* d_Wt_im = nullptr;  // imaginary part of the rate matrix (general tables: CorrelatedDaviesGenerator)
    long long ncross = 0, nlamb = 0;
    int32_t* d_cross_idx = nullptr;
    double* d_cross_rate = nullptr;
    oqb::c128* d_Moff = nullptr;  // dense lvl×lvl off-diagonal part of −½G − iH_LS (nullptr when empty)
    // one-sided AME
    int npairs = 0;
    std::vector<int32_t> h_beta;
    oqb::c128* d_Lam = nullptr;  // npairs lvl×lvl
    oqb::c128* d_Psum = nullptr;     // Σ_p A_βp Λ_p (lvl×lvl)
    oqb::c128* d_Abstack = nullptr;  // [A_β1; …; A_βP]  (npairs·lvl × lvl)
    // ame_jump tables (ame_jump.cu): probability slots in the reference's order
    int nslots = 0;
    long long nterms = 0;
    long long* d_slot_ptr = nullptr;  // nslots+1 offsets into the term list
    int32_t* d_term = nullptr;        // (coupling, row, col) per term, sorted by (row, col) within a slot
    double* d_slot_rate = nullptr;    // γ at the slot's gap
    // device-side GapIndices (ame_gaps.cu): pairs flagged as members of multi-pair groups / of the zero group
    double* d_grp_key = nullptr;
    int32_t* d_grp_pair = nullptr;
    long long n_multi = 0, n_zero = 0;
    // tabulated Lamb shift for the device-built tables: quadratic B-spline coefficients (n_Stab+2) on x0 + k·dx
    int n_Stab = 0;
    double S_x0 = 0.0, S_dx = 1.0;
    double* d_Scoef = nullptr;
};

// Per-handle device buffers come from the stream-ordered allocator: cudaMalloc/cudaFree would synchronise the whole
// device, which stalls a side stream that prepares the next eigen-system while the current one is in use.
namespace oqb {
template <class T>
inline void ame_free(Ctx* c, T*& p) {
    if (p) cudaFreeAsync((void*)p, c->stream);
    p = nullptr;
}
inline int ame_alloc(Ctx* c, void** p, size_t bytes) {
    cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 16, c->stream);
    if (e != cudaSuccess) return set_error(c, OQB_ENOMEM, "cudaMallocAsync(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return 0;
}
}  // namespace oqb

// ame.cu internals shared with ame_gaps.cu (defined inside ame.cu's extern "C" block; not part of the public ABI)
extern "C" int ame_davies_alloc(oqb_ame* a);
extern "C" int ame_davies_finish(oqb_ame* a, int64_t ncross, const int32_t* h_cross_idx, const double* h_cross_rate, int64_t nlamb,
                      const int32_t* h_lamb_idx, const double* h_lamb_rate_S);

extern "C" int ame_onesided_install(oqb_ame* a, int npairs, const int32_t* h_alpha, const int32_t* h_beta, const double* d_g,
                                    const double* d_s, int tables_shared);
