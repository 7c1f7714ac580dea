// Device representation of one sparse operator term (shared by sparse.cu and traj_fast.cu).
#pragma once
#include <vector>

#include "common.cuh"

namespace oqb {

struct TermDev {
    int is_diag = 0, is_real = 0, width = 0;
    c128* vals = nullptr;     // diag: n ; ELL: width*n (column-major: [j*n + r])
    int32_t* cols = nullptr;  // ELL only
    // compressed forms used by the fast trajectory kernel (traj_fast.cu)
    double* rvals = nullptr;         // diag && is_real: n real values
    double* absq = nullptr;          // diag: |d_r|² (jump probabilities)
    bool absq_const = false;         // … identical for every row (unit-modulus diagonals: Pauli Z strings)
    double absq_val = 0.0;
    uint16_t* cols16 = nullptr;      // ELL && n ≤ 65536 && every row full: 16-bit column indices
    bool slot_const = false;         // ELL: every slot j holds one value for all rows (Pauli-string structure)
    std::vector<c128> slot_vals;     // … those values
    // constant diagonal (value·I) whose coefficient is 1 for every member: folds into one scalar per call
    bool diag_const = false, const_coef_one = false;
    c128 const_val = {0.0, 0.0};
    // XOR form: M[r, r ^ mask_j] = (popcount(r & zmask_j) even ? val_j : val1_j) for every row r — Pauli strings: the X-type
    // factors give the mask, Y/Z factors the row-parity pattern (val1 = −val for one string; {0, 2J} for XX + YY): no per-row data
    bool xor_form = false;
    std::vector<int> xor_masks;
    std::vector<c128> xor_vals;
    std::vector<c128> xor_vals1;  // per slot: the value on rows with popcount(r & zmask) odd (xor_vals: even)
    std::vector<int> xor_zmasks;  // per slot: which row bits decide between the two values (0 = one value)
    // host copy of the ELL structure for the SELL planner (traj_sell.cu): entry j < h_cnt[r] of row r sits at [j*n + r]
    std::vector<int32_t> h_cols;
    std::vector<int> h_cnt;
    bool ell_real = false, ell_imag = false;  // every stored value real / purely imaginary
    bool dynamic_vals = false;                // values rewritten on the device between calls (−½Σγ_k L_k'L_k)
};

// dψ_b = Σ_t coef[b][t] · M_t ψ_b with the operator held in registers (see traj_fast.cu).
// *used == false ⇒ the term list does not fit the fast path; the caller runs the generic kernel.
int traj_matvec_fast(Ctx* c, int n, int nb, const std::vector<const TermDev*>& ts, const c128* d_coef,
                     long long coef_ld, const c128* psi, c128* dpsi, bool* used);

// Optional RK4 stage arithmetic fused into the store of k = Aψ_in (used by oqb_traj_evolve, traj_evolve.cu):
//   mode 1 (first stage, ψ_in = ψ0):  acc = ψ0 + wa·k,  out = ψ0 + wt·k
//   mode 2 (middle stages):           acc += wa·k,      out = ψ0 + wt·k     (out may alias ψ_in: a trajectory is
//                                                                            staged completely before it is computed)
//   mode 3 (last stage):              out = acc + wa·k;  norm_part[b·16 + w] = the warps' partial sums of ‖out_b‖² (w < 16,
//                                     unused slots zero) when norm_part != nullptr
struct RkEpilogue {
    int mode = 0;
    double wa = 0.0, wt = 0.0;
    const c128* psi0 = nullptr;
    c128* acc = nullptr;
    double* norm_part = nullptr;
};

// Same contract for terms in XOR form + real diagonal terms (implicit columns, 3-deep bulk-async ψ ring,
// no CTA-wide barrier): traj_xor.cu.
int traj_matvec_xor(Ctx* c, int n, int nb, const std::vector<const TermDev*>& ts, const c128* d_coef,
                    long long coef_ld, const c128* psi, c128* dpsi, bool* used, const RkEpilogue* rk = nullptr);

// Unstructured operators: sliced-ELL (32-row slices, rows sorted by length, slots ordered so that the 8 threads of a
// shared-memory phase gather from 8 different bank groups), ψ staged by bulk-async copies, R trajectories per operator
// read: traj_sell.cu.  `cache` is owned by the caller's oqb_sparse (sell_cache_free releases it).
struct SellCache;
int traj_matvec_sell(Ctx* c, int n, int nb, const std::vector<const TermDev*>& ts, const c128* d_coef, long long coef_ld,
                     const c128* psi, c128* dpsi, bool* used, SellCache** cache, const RkEpilogue* rk = nullptr);
void sell_cache_free(SellCache* cache);
void sell_cache_stats(const SellCache* cache, int* plans, double* padding);
int sell_plan_probe(int n, int nterms, const int64_t* const* colptr, const int32_t* const* rowval, int unit, double* padding,
                    long long* phases, long long* conflicting, int* identity, int* misplaced);  // padding = stored ÷ real entries of the newest plan

}  // namespace oqb
