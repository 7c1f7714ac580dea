// Private layout of the eigenbasis-Liouvillian handle (shared by ame.cu and ame_jump.cu).
#pragma once
#include <vector>

#include "common.cuh"

// One DaviesGenerator of the term list `opensys_eig` (src/opensys/diffeq_liouvillian.jl:101-103): its coupling slice
// [k0, k0+nk) of the handle's K couplings and γ at the keys of the shared cross-term list.
struct DaviesTerm {
    int k0 = 0, nk = 0;
    double* d_rate = nullptr;  // ncross rates (owned)
};

// One (α,β) pair of a CorrelatedDaviesGenerator (src/opensys/davies.jl:231-296): A_β on the left, A_α' on the right.
struct CorrTerm {
    int ka = 0, kb = 0;
    double* d_rate = nullptr;  // ncross rates (owned)
};

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
    int32_t* d_cross_idx = nullptr;  // (a,b,a2,b2) per cross entry — shared by all Davies terms (it depends on the gaps only)
    std::vector<DaviesTerm> terms;
    int cross_fuse_state = -1;  // −1 unknown, 1: C ≠ 0 at every cross-term source (the Hadamard may stay in the GEMM epilogue), 0: not
    oqb::c128* d_Moff = nullptr;  // dense lvl×lvl off-diagonal part of −½G − iH_LS (nullptr when empty)
    // gap structure of the current eigen-system (ame_terms.cu, GapIndices of src/opensys/opensys_util.jl:25-61)
    bool have_gaps = false, lists_built = false;
    int gap_digits = 8, gap_sigdigits = 8;
    long long ng = 0;              // distinct non-zero gap keys (= GapIndices.uniq_w, ascending)
    long long nzp = 0;             // level pairs (i<j) of the zero group
    int32_t* d_gid = nullptr;      // lvl×lvl signed group number: +(g+1) at (i,j), −(g+1) at (j,i), i<j; 0 inside the zero group
    double* d_ukeys = nullptr;     // ng keys
    int32_t* d_spair = nullptr;    // the l(l−1)/2 pair numbers sorted by key (stable); the zero-group pairs are the tail
    std::vector<long long> h_gptr; // ng+1 offsets of the groups in d_spair
    std::vector<double> h_ukeys;
    int32_t* d_gslot = nullptr;    // per group: number of its dense slot, −1 when it is on the cross-term list
    std::vector<long long> h_dense_groups;  // groups that take the dense per-group products
    bool zero_dense = false;
    int32_t* d_cross_g = nullptr;  // signed group number per cross entry (0: zero group)
    int32_t* d_lamb_idx = nullptr; // (a,b,b2) per entry that also contributes to L'L
    int32_t* d_lamb_g = nullptr;
    // dense per-group blocks L = A_k∘[ω̃ == x] of all terms (src/opensys/davies.jl:26-45 applied group by group)
    int nd = 0;
    oqb::c128* d_Lstack = nullptr;  // nd lvl×lvl blocks side by side (an lvl × nd·lvl matrix)
    double* d_drate = nullptr;      // γ(x) per block
    bool dense_real = false;
    // correlated pairs: right-hand operators differ from the left-hand ones
    std::vector<CorrTerm> corr_terms;
    bool has_corr = false;
    oqb::c128* d_Rstack = nullptr;   // right-hand blocks (nullptr: same as d_Lstack)
    oqb::c128* d_Moff2 = nullptr;    // −½G + iH_LS off-diagonal part applied on the RIGHT of ρ (nullptr: d_Moff')
    oqb::c128* d_Mscratch = nullptr; // l×l scratch
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
extern "C" int ame_clear_gaps(oqb_ame* a);      // ame_terms.cu
extern "C" int ame_clear_terms(oqb_ame* a);     // ame_terms.cu: cross lists, dense blocks, per-term rates
int ame_detect_real(oqb_ame* a, const oqb::c128* v, long long n, bool* out);  // ame.cu
extern "C" int ame_corr_cross(oqb_ame* a, int nb, const oqb::c128* R, oqb::c128* X);  // ame_terms.cu
extern "C" int ame_davies_finish(oqb_ame* a, int64_t ncross, const int32_t* h_cross_idx, const double* h_cross_rate, int64_t nlamb,
                      const int32_t* h_lamb_idx, const double* h_lamb_rate_S);

extern "C" int ame_onesided_install(oqb_ame* a, int npairs, const int32_t* h_alpha, const int32_t* h_beta, const double* d_g,
                                    const double* d_s, int tables_shared);
