/* oqb.h — C ABI of the B200-native master-equation RHS evaluator for OpenQuantumBase.jl.
 *
 * The reference (/root/reference, pure Julia, v0.7.9) has no FFI today: its "operator API" is
 * the callable-struct protocol `op(du,u,p,t)`, `update_cache!`, `update_vectorized_cache!`.
 * Each entry point below names the reference method(s) it evaluates (file:line relative to
 * /root/reference); INTEGRATION.md shows the Julia `ccall` methods a maintainer adds so that the
 * unchanged Julia types dispatch here for device-resident batches.
 *
 * Conventions
 *  - Every function returns 0 on success, non-zero on failure; oqb_last_error(ctx) gives the
 *    message.  Nothing throws, nothing calls back into the host language.
 *  - Layouts are Julia-native: column-major, interleaved ComplexF64 (re,im doubles).  A batch
 *    of nb matrices is nb consecutive n×n blocks (a Julia Array{ComplexF64,3} of size n×n×nb);
 *    a batch of state vectors is n×nb.
 *  - Pointers prefixed d_ are DEVICE pointers (from oqb_malloc or any CUDA allocator), h_ are
 *    HOST pointers.  Per-call scalar coefficient arrays (closure values f_i(s_b), γ_k(s_b) that
 *    the host language evaluated) are small HOST arrays, row-major [nb][count]; passing
 *    nb_coef == 1 rows with the *_shared flag set shares one row across the whole batch.
 *  - All work is enqueued on the context's CUDA stream; call oqb_sync before reading results
 *    on the host (the *_host entry points synchronise themselves).
 *  - A context is not thread-safe; use one per (host thread, GPU) — the reference's functors are
 *    not re-entrant either (src/hamiltonian/dense_hamiltonian.jl:61 mutates an internal cache).
 *  - Process model: one context per (host thread, GPU).  A process may hold contexts on several GPUs (one host
 *    thread per GPU) or run one rank per GPU under torchrun / MPI / Distributed.jl: every entry
 *    point selects the device of the context it was handed and per-device kernel configuration is tracked per
 *    device.  The library reads no environment variables; switches are per context (oqb_set_option).
 *  - There is no CPU fallback anywhere in this library.
 */
#ifndef OQB_H
#define OQB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OQB_VERSION 100

typedef struct oqb_ctx oqb_ctx;       /* stream + workspace                                   */
typedef struct oqb_dense oqb_dense;   /* DenseHamiltonian terms + Lindblad operators          */
typedef struct oqb_sparse oqb_sparse; /* SparseHamiltonian terms + sparse Lindblad operators  */
typedef struct oqb_ame oqb_ame;       /* eigenbasis (Davies / one-sided AME) Liouvillian      */
typedef struct oqb_lambda oqb_lambda; /* batched builder of the Redfield kernel matrices Λ(t) */
typedef struct oqb_liouv oqb_liouv;   /* DiffEqLiouvillian{true,false}: H terms + eigenbasis terms + cached eigen-systems */
typedef struct oqb_comm oqb_comm;     /* NCCL communicator for the ensemble-average reduce      */

enum { OQB_OP_N = 0, OQB_OP_T = 1, OQB_OP_C = 2 };
enum { OQB_OK = 0, OQB_EINVAL = 1, OQB_ECUDA = 2, OQB_ENOMEM = 3 };

/* ------------------------------------------------------------------ context / memory ------ */
int oqb_version(void);
int oqb_create(int device, oqb_ctx** out);
int oqb_destroy(oqb_ctx* ctx);
const char* oqb_last_error(const oqb_ctx* ctx);
int oqb_sync(oqb_ctx* ctx);
int oqb_set_stream(oqb_ctx* ctx, void* cuda_stream); /* borrow an existing cudaStream_t        */
int oqb_get_stream(oqb_ctx* ctx, void** cuda_stream);
int oqb_launch_count(const oqb_ctx* ctx, uint64_t* n); /* kernels launched through this ctx    */
/* Per-context switches by name (A/B measurements and structure exploitation; defaults in parentheses):
 *   "zgemm_3m" (1), "real_operand" (1)        — as oqb_set_zgemm_mode / oqb_set_real_operand_mode below
 *   "zgemm_tma" (1)  0 = cp.async ZGEMM only;  "zgemm_bm" (0) 64/128 force the CTA tile height;  "tma_oneshot" (1)
 *   "lindblad_diag" (1), "coupling_diag" (1)   — treat diagonal L_k / couplings as general dense matrices when 0
 *   "coupling_mono" (1)  consulted by host bindings: route monomial Redfield couplings to oqb_redfield_apply_mono_acc
 *   "lindblad_mono" (1)  monomial L_k (one entry per row and column: sigma+-, Pauli strings): gather of rho instead of 2K products
 *   "lindblad_fuse" (1)  diagonal L_k with rates shared by the batch: one n×n factor G per call and du += G∘u as a streaming pass
 *   "lindblad_split" (1) … and one off-diagonal Hamiltonian term: shared-operand products scaled by f_N(s_b), diagonal terms in that pass
 *   "traj_kernel" (0)  0 = pick by operator structure, 1 = no Pauli/XOR specialisation at all, 2 = no XOR specialisation
 *   "traj_ell" (1)     operators without Pauli structure take the sliced-ELL kernel (csrc/traj_sell.cu); 0 = the plain ELL kernels
 *   "sell_r" (0)       trajectories per operator fetch in that kernel (0 = pick: 2);  "sell_unit" (2) rows permuted in pairs or singly
 *   "rk_fuse" (1), "xor_fixed" (1), "xor_rows" (16)
 *   "ame_host_chunk_mib" (128)                  — chunk size of oqb_ame_rhs_host's H2D/D2H pipeline
 *   "davies_dense_min_pairs" (0 = cost model)   — gap groups with at least this many level pairs use the dense per-group
 *                                                  products instead of the cross-term list;  "davies_max_cross" (5e7) caps the list.
 *   "fuse_cross" (1)                            — keep the Hadamard in the rotation product's epilogue when the only extra Davies terms
 *                                                  are list cross terms (their ρ̃ sources are recovered as X/C); 0: separate passes
 *   "debug_checks" (0)                          — 1: every index list the Davies term builder produces is range- and
 *                                                  consistency-checked on the device before use (an error is returned) */
int oqb_set_option(oqb_ctx* ctx, const char* name, int64_t value);
int oqb_get_option(const oqb_ctx* ctx, const char* name, int64_t* value);
/* Complex product used by the tensor-pipe ZGEMM behind every dense path: 1 (default) = 3-multiplication
 * form (3 real DMMA products per complex product, ≈1.2× faster, same ≈1e-15 normwise error), 0 = the classic
 * 4-multiplication form whose rounding matches BLAS zgemm element by element (src/opensys/lindblad.jl:99). */
int oqb_set_zgemm_mode(oqb_ctx* ctx, int mode);
int oqb_get_zgemm_mode(const oqb_ctx* ctx, int* mode);
/* Real eigenvectors: when the eigenvector matrix handed to oqb_ame_set_eigen(_device) has no imaginary part (H(s) real
 * symmetric — every X/Z/ZZ annealing Hamiltonian), the rotations v'ρv and vXv' need 2 real products per complex product
 * instead of 3 (the other factor's imaginary part times zero is skipped).  1 (default) = detect and exploit, 0 = always
 * run the general complex product.  Results agree to rounding. */
int oqb_set_real_operand_mode(oqb_ctx* ctx, int on);
int oqb_get_real_operand_mode(const oqb_ctx* ctx, int* on);

int oqb_malloc(oqb_ctx* ctx, size_t nbytes, void** d_ptr);
int oqb_free(oqb_ctx* ctx, void* d_ptr);
int oqb_malloc_host(oqb_ctx* ctx, size_t nbytes, void** h_ptr); /* pinned                      */
int oqb_free_host(oqb_ctx* ctx, void* h_ptr);
int oqb_upload(oqb_ctx* ctx, void* d_dst, const void* h_src, size_t nbytes);   /* synchronous */
int oqb_download(oqb_ctx* ctx, void* h_dst, const void* d_src, size_t nbytes); /* synchronous */

/* y[b] += alpha·x[b] for nb blocks of len complex elements (stride_x = 0 shares one x): the `cache .+= D.A` /
 * `du .+= …` accumulations of the reference's constant-operator Liouvillians (src/opensys/davies.jl:119,
 * src/opensys/stochastic.jl:34-44) when the term was produced by an overwriting entry point. */
int oqb_axpy_batched(oqb_ctx* ctx, int64_t len, int nb, const double* alpha, const void* d_x, int64_t stride_x, void* d_y);

/* --------------------------------------------------------------- K1: batched ZGEMM -------- */
/* C[b] = alpha*opA(A[b])*opB(B[b]) + beta*C[b] on the FP64 tensor pipe.  Replaces LinearAlgebra
 * `gemm!`/`mul!` on batches (src/hamiltonian/dense_hamiltonian.jl:87-88).  Strides in complex
 * elements; stride 0 shares the operand.  alpha/beta point to (re,im). */
int oqb_zgemm_batched(oqb_ctx* ctx, int opA, int opB, int M, int N, int K, const double* alpha,
                      const void* d_A, int64_t lda, int64_t strideA, const void* d_B, int64_t ldb,
                      int64_t strideB, const double* beta, void* d_C, int64_t ldc, int64_t strideC,
                      int batch);

/* ------------------------------------ DenseHamiltonian + LindbladLiouvillian (+ Redfield) -- */
/* h_mats: nterms n×n matrices m_i exactly as DenseHamiltonian stores them (already unit-scaled,
 * src/hamiltonian/dense_hamiltonian.jl:38).  h_lops: K Lindblad operators L_k (may be NULL/K=0),
 * src/opensys/lindblad.jl:76-92.  Uploads once and precomputes L_k'L_k on the device. */
int oqb_dense_create(oqb_ctx* ctx, int n, int nterms, const void* h_mats, int K, const void* h_lops,
                     oqb_dense** out);
/* The host-only structure analysis oqb_dense_create runs on the Hamiltonian terms — needs NO GPU and no context.
 *   *offdiag_term = index of the ONE term with off-diagonal entries when every other term is a diagonal matrix (and
 *   nterms ≥ 2), else −1: H(s) = f_N(s)·m_N + Σ f_i(s)·diag(h_i), src/hamiltonian/dense_hamiltonian.jl:60-66, is then never
 *   assembled.  *nmask = number of XOR diagonals that hold the non-zeros of m_N (m_N[r, r ^ masks16[j]] = x_j[r]; 0 when there
 *   are more than 16 or n is not a power of two): with nmask > 0 the commutator :84-89 runs as a stencil pass instead of
 *   two products.  isconst16[j] = x_j is one value for every row; *all_real = every x_j is real; xval (may be NULL):
 *   nmask × n complex values x_j[r]. */
int oqb_dense_structure_probe(int n, int nterms, const void* h_mats, int* offdiag_term, int* nmask, int* masks16, int* isconst16,
                              int* all_real, void* xval);
int oqb_dense_destroy(oqb_dense* op);

/* out[b] = Σ_i coefH[b][i] m_i                      — (h::DenseHamiltonian)(s), :60-66          */
int oqb_dense_hamiltonian(oqb_dense* op, int nb, const double* h_coefH, int coef_shared, void* d_out);
/* cache[b] = Σ_i (-i coefH[b][i]) m_i − ½ Σ_k gamma[b][k] L_k'L_k   (OVERWRITES)
 *   — update_cache!(cache, H::DenseHamiltonian, p, s) :71-76 followed by
 *     update_cache!(cache, lind::LindbladLiouvillian, p, t) src/opensys/lindblad.jl:103-109,
 *     i.e. update_cache!(cache, Op::DiffEqLiouvillian{false,false}, p, t)
 *     src/opensys/diffeq_liouvillian.jl:78-83.  h_gamma may be NULL (Hamiltonian only). */
int oqb_dense_update_cache(oqb_dense* op, int nb, const double* h_coefH, int coef_shared,
                           const double* h_gamma, int gamma_shared, void* d_cache);
/* cache[b] (n²×n²) = i(H'ᵀ⊗I − I⊗H)                — update_vectorized_cache! :78-82          */
int oqb_dense_update_vectorized_cache(oqb_dense* op, int nb, const double* h_coefH, int coef_shared,
                                      void* d_cache);
/* du[b] = −i[H(s_b),u[b]] + Σ_k γ_k (L_k u L_k' − ½{L_k'L_k,u})     (OVERWRITES du)
 *   — (Op::DiffEqLiouvillian{false,false})(du,u,p,t) src/opensys/diffeq_liouvillian.jl:70-76 =
 *     (h::DenseHamiltonian)(du,u,p,s) :84-89 + (Lind::LindbladLiouvillian)(du,u,p,t)
 *     src/opensys/lindblad.jl:94-101.  h_gamma NULL ⇒ commutator only. */
int oqb_dense_rhs(oqb_dense* op, int nb, const double* h_coefH, int coef_shared,
                  const double* h_gamma, int gamma_shared, const void* d_u, void* d_du);
/* same, with HOST state buffers (pinned or pageable); chunks are pipelined H2D→RHS→D2H. */
int oqb_dense_rhs_host(oqb_dense* op, int nb, const double* h_coefH, int coef_shared,
                       const double* h_gamma, int gamma_shared, const void* h_u, void* h_du);
/* du[b] += Σ_k γ_k (L_k u L_k' − ½{L_k'L_k,u})      (ACCUMULATES) — src/opensys/lindblad.jl:94 */
int oqb_lindblad_acc(oqb_dense* op, int nb, const double* h_gamma, int gamma_shared,
                     const void* d_u, void* d_du);
/* du[b] −= K₂+K₂', K₂ = S_α Λ_α u − Λ_α u S_α summed over α<K  (ACCUMULATES)
 *   — the ρ-dependent half of (R::RedfieldLiouvillian)(du,u,p,t) src/opensys/redfield.jl:65-68.
 *   d_S: K n×n coupling matrices (shared, or nb×K when bit 0 of S_shared is clear); d_Lambda: nb×K n×n.
 *   S_shared bit 1 (value 2): the caller asserts that no entry of d_S has an imaginary part (σx, σ±, σz … couplings) —
 *   the two products with S_α then run as real-operand products (see oqb_set_real_operand_mode). */
int oqb_redfield_apply_acc(oqb_ctx* ctx, int n, int K, const void* d_S, int S_shared,
                           const void* d_Lambda, int nb, const void* d_u, void* d_du);
/* The same for DIAGONAL couplings S_α = diag(s_α) (σz-type system–bath couplings stored as dense matrices): K₂ has the
 * entries Σ_α (s_α[i] − s_α[j])·(Λ_α u)[i,j], so the 2K products with S_α become one elementwise pass — K GEMMs per
 * member instead of 3K.  d_Sdiag: K×n complex diagonals (device).  Its own roofline row. */
int oqb_redfield_apply_diag_acc(oqb_ctx* ctx, int n, int K, const void* d_Sdiag, const void* d_Lambda, int nb,
                                const void* d_u, void* d_du);
/* The same for MONOMIAL couplings (at most one entry per row and per column: σx, σy, σ±, Pauli strings):
 * S_α[a, perm_α[a]] = val_α[a] (perm −1 = empty row), pinv_α[c] = the row holding column c's entry (−1 = none); all K×n,
 * on the device.  (S_αM − MS_α)[a,c] = val_α[a]·M[perm_α[a], c] − M[a, pinv_α[c]]·val_α[pinv_α[c]] with M = Λ_α u: two
 * gathers instead of two products — K GEMMs per member instead of 3K (src/opensys/redfield.jl:57-64). */
int oqb_redfield_apply_mono_acc(oqb_ctx* ctx, int n, int K, const int32_t* d_perm, const int32_t* d_pinv, const void* d_val,
                                const void* d_Lambda, int nb, const void* d_u, void* d_du);
/* cache[b] (n²×n²) −= I⊗SΛ + conj(SΛ)⊗I − Sᵀ⊗Λ − conj(Λ)⊗S   (ACCUMULATES)
 *   — update_vectorized_cache!(cache, R::RedfieldLiouvillian, p, t) src/opensys/redfield.jl:97-102 */
int oqb_redfield_vectorized_acc(oqb_ctx* ctx, int n, int K, const void* d_S, int S_shared,
                                const void* d_Lambda, int nb, void* d_cache);

/* ------------------------------------------------------- Redfield Λ(t) builder */
/* Replaces the `quadgk!` call of (R::RedfieldLiouvillian)(du,u,p,t), src/opensys/redfield.jl:46-64 (same call
 * in update_vectorized_cache!, :76-96), for many integrals at once:
 *     Λ_ij(t) = ∫_{max(0,t−Ta)}^{t} C_ij(t,x) U(t)U(x)† S_j(p(x)) U(x)U(t)† dx.
 * The adaptive G7K15 control flow (QuadGK: per-integral max-heap, bisect the worst segment) and the scalar
 * closures C_ij(t,x) stay in the host language; each refinement round is one oqb_lambda_eval call.
 * h_S: K coupling matrices S_j (n×n, unit-scaled, src/coupling/coupling.jl:42,66); a scalar time dependence
 * f_j(p(x)) is folded into the node weights by the host. */
int oqb_lambda_create(oqb_ctx* ctx, int n, int K, const void* h_S, oqb_lambda** out);
int oqb_lambda_destroy(oqb_lambda* lam);
/* The closed-system unitaries `R.unitary` (src/opensys/redfield.jl:16, :38) of nb schedules, each sampled on G
 * uniformly spaced points starting at 0: d_Ugrid = nb×G n×n matrices in device memory (borrowed, must stay
 * valid); U(x) is the piecewise-linear interpolant (1−θ)U_g + θU_{g+1}. */
int oqb_lambda_set_unitary(oqb_lambda* lam, int nb, int G, const void* d_Ugrid);
/* storage for the integrals of the live (leaf) segments, n×n each; grows preserving content */
int oqb_lambda_reserve_slots(oqb_lambda* lam, int64_t nslots);
/* One Gauss–Kronrod(7,15) evaluation of nseg segments.  For segment s: h_member[s] (which schedule),
 * h_coupling[s] (which S_j), 16 interpolation points h_gi/h_theta[s][0..14] = the 15 nodes x_i — the 7 GAUSS nodes
 * first, then the 8 Kronrod-only nodes — and [15] = the end time t (grid index g and fraction θ ∈ [0,1)); complex
 * node weights as (re,im) pairs: h_ck[s][i] = wk_i·h·C(t,x_i) for the 15 nodes and h_cg[s][i] = wg_i·h·C(t,x_i) for
 * the first 7 (h = half width); h_slot[s] the slot that receives the segment's Kronrod integral.  h_E[s] returns
 * ‖Kronrod − Gauss‖_F (QuadGK's segment error).  Synchronises. */
int oqb_lambda_eval(oqb_lambda* lam, int nseg, const int32_t* h_member, const int32_t* h_coupling,
                    const int32_t* h_gi, const double* h_theta, const double* h_ck, const double* h_cg,
                    const int64_t* h_slot, double* h_E);
/* Λ_q = Σ_k slot[h_idx[k]], k ∈ [h_ptr[q], h_ptr[q+1]) summed in that order (QuadGK's final re-summation over
 * its heap) into d_Lambda[q] (n×n, device); h_norm[q] = ‖Λ_q‖_F for the rtol test (NULL ⇒ not returned, no
 * synchronisation). */
int oqb_lambda_total(oqb_lambda* lam, int nint, const int64_t* h_ptr, const int64_t* h_idx, void* d_Lambda,
                     double* h_norm);

/* The whole adaptive quadrature inside the library: quadgk!(…, max(0, t−Ta), t; rtol, atol) of src/opensys/redfield.jl:57-64
 * (window [lo, hi] per integral: ULE passes hi = min(t+Ta, tf), src/opensys/lindblad.jl:60-61) for nint integrals in lock
 * step — per-integral max-heap, bisect the worst segment until E ≤ atol or E ≤ rtol·‖I‖_F, final re-summation over the heap.
 * Integral q: schedule h_member[q], coupling h_coupling[q], end time h_t[q] (where U(t) is taken), window h_lo/h_hi[q];
 * h_dt[m]: grid spacing of schedule m's sampled unitary (G points, oqb_lambda_set_unitary).  The correlation function is a C
 * callback evaluated on the calling thread once per refinement round for all new nodes:
 *   cfun(user, nseg, integral[nseg], t[nseg], x[nseg×15], out[nseg×15 (re,im)]) with out = C(t, x)·f_j(p(x)).
 * d_Lambda: nint n×n matrices (device).  rounds / segments report the work done (may be NULL). */
typedef void (*oqb_kernel_fn)(void* user, int64_t nseg, const int32_t* integral, const double* t, const double* x, double* out);
int oqb_lambda_integrate(oqb_lambda* lam, oqb_ctx* ctx, int nint, const int32_t* h_member, const int32_t* h_coupling,
                         const double* h_t, const double* h_lo, const double* h_hi, const double* h_dt, int G,
                         oqb_kernel_fn cfun, void* user, double atol, double rtol, int64_t maxevals, void* d_Lambda,
                         int64_t* rounds, int64_t* segments, int64_t* h_evals /* integrand evaluations per integral, may be NULL */);
/* The same with the Ohmic bath's C(t − x) (src/bath/ohmic.jl:48-52) evaluated by the library — no callback at all. */
int oqb_lambda_integrate_ohmic(oqb_lambda* lam, oqb_ctx* ctx, int nint, const int32_t* h_member, const int32_t* h_coupling,
                               const double* h_t, const double* h_lo, const double* h_hi, const double* h_dt, int G, double eta,
                               double wc, double beta, double atol, double rtol, int64_t maxevals, void* d_Lambda,
                               int64_t* rounds, int64_t* segments, int64_t* h_evals);
int oqb_lambda_dims(const oqb_lambda* lam, int* n, int* K);

/* ------------------------------------------ eigenbasis Liouvillian: DiffEqLiouvillian{true,·} */
/* n: Hilbert dimension; lvl: kept levels (src/opensys/diffeq_liouvillian.jl:45-54);
 * h_couplings: K coupling matrices A_α (n×n, unit-scaled, src/coupling/coupling.jl:42,66). */
int oqb_ame_create(oqb_ctx* ctx, int n, int lvl, int K, const void* h_couplings, oqb_ame** out);
int oqb_ame_destroy(oqb_ame* ame);
/* Moves a handle to another context of the same device: later calls launch on that context's stream and use its
 * workspace.  The caller orders the two streams (all work issued through the old context must have completed).  Lets
 * a side stream prepare the next eigen-system while the current one is in use. */
int oqb_ame_rebind(oqb_ame* ame, oqb_ctx* ctx);
/* (w, v) exactly as haml_eigs returned them on the host (src/hamiltonian/hamiltonian_base.jl:155);
 * h_v NULL ⇒ adiabatic frame (v = I, n == lvl).  Rotates the couplings on the device
 * (src/opensys/davies.jl:24). */
int oqb_ame_set_eigen(oqb_ame* ame, const double* h_w, const void* h_v);
/* same with (w, v) already in device memory — e.g. straight out of a device eigensolver (cuSOLVER Zheevd), so that
 * the n×lvl eigenvector matrix never crosses PCIe; only w (lvl doubles) goes back for the host's gap grouping. */
int oqb_ame_set_eigen_device(oqb_ame* ame, const double* d_w, const void* d_v);
/* *out = 1 when the current eigenvector matrix is exactly real (the real-operand products are then used). */
int oqb_ame_eigenvectors_real(const oqb_ame* ame, int* out);
/* Davies generator tables (src/opensys/davies.jl:21-47, closed form of SURVEY App. A.5):
 *  h_gam[a+b*lvl] = γ(ω̃_ab), h_S[a+b*lvl] = S(ω̃_ab) at the signed ROUNDED gaps of GapIndices
 *  (src/opensys/opensys_util.jl:25-61; 0 inside the zero group), evaluated by the host closures.
 *  Cross terms of non-singleton gap groups: ncross entries (a,b,a2,b2) 0-based int32 quadruples
 *  with rate h_cross_rate, meaning du[a,a2] += rate·Σ_α A_α[a,b]·conj(A_α[a2,b2])·ρ[b,b2];
 *  nlamb entries (a,b,b2) with (rate,S) meaning (L'L)[b,b2] gets conj(A[a,b])A[a,b2]. */
int oqb_ame_set_davies(oqb_ame* ame, const double* h_gam, const double* h_S, int64_t ncross,
                       const int32_t* h_cross_idx, const double* h_cross_rate, int64_t nlamb,
                       const int32_t* h_lamb_idx, const double* h_lamb_rate_S);
/* Tabulated Lamb shift for the device-built tables: the n+2 prefiltered coefficients of the quadratic B-spline that
 * build_lambshift / construct_interpolations return for a range grid x0 + k·dx, k = 0…n−1 (src/bath/bath_util.jl:26-38,
 * src/interpolation.jl:22-34; Interpolations.jl BSpline(Quadratic(Line(OnGrid()))), value 0 outside the grid).
 * oqb_ame_set_onesided_ohmic then fills S at the unrounded gaps from it on the device (the Davies tables take their spline as an
 * argument of oqb_ame_add_davies_ohmic); n = 0 restores S ≡ 0. */
int oqb_ame_set_lambshift_spline(oqb_ame* ame, int n, double x0, double dx, const double* h_coef);
int oqb_ame_clear_davies(oqb_ame* ame);

/* ---- GapIndices and the Davies term list, built inside the library ------------------------------------------------
 * GapIndices(w, digits, sigdigits) of the current eigen-system (src/opensys/opensys_util.jl:25-61), on the device:
 * |w_j − w_i| ≤ 10^-digits → zero group, else round(gap, sigdigits) (Julia's rule), grouped by exact equality.
 * *n_keys = number of distinct non-zero keys (= length(uniq_w)), *n_zero_pairs = level pairs i<j of the zero group.
 * Invalidates Davies tables built for another grouping. */
int oqb_ame_gaps_build(oqb_ame* ame, int digits, int sigdigits, int64_t* n_keys, int64_t* n_zero_pairs);
/* h_keys[n_keys] = uniq_w, ascending — the arguments the reference's loop hands to γ(w), γ(−w), S(w), S(−w)
 * (src/opensys/davies.jl:26-28): the host language evaluates its closures at these and passes the values back. */
int oqb_ame_gaps_keys(oqb_ame* ame, double* h_keys);
/* The index lists in the reference's own form, 1-based: group g owns h_a/h_b[h_group_ptr[g] .. h_group_ptr[g+1]) (= uniq_a[g],
 * uniq_b[g] in the (i<j) loop order); h_a0/h_b0 (2·n_zero_pairs + lvl entries) = a0, b0 (:34-37, :58-59).  Any pointer may be
 * NULL.  For callers that want the GapIndices object itself; the RHS does not need it. */
int oqb_ame_gaps_fetch(oqb_ame* ame, int64_t* h_group_ptr, int32_t* h_a, int32_t* h_b, int32_t* h_a0, int32_t* h_b0);
/* Adds one DaviesGenerator (src/opensys/davies.jl:21-47) over the couplings [k0, k0+nk) of the handle:
 * h_gamma_plus[g] = γ(uniq_w[g]), h_gamma_minus[g] = γ(−uniq_w[g]), likewise S (NULL ⇒ S ≡ 0), gamma0 = γ(0), S0 = S(0).
 * Calls accumulate: one per generator of `opensys_eig` (src/opensys/diffeq_liouvillian.jl:101-103; davies_from_interactions
 * builds one per interaction, src/coupling/interaction.jl:101-117).  Single-pair groups go into the closed-form
 * tables, small groups onto the cross-term list, large groups (degenerate spectra) are applied as dense products
 * L_x ρ̃ L_x' with L_x = A∘[ω̃ == x], group by group like the reference's loop. */
int oqb_ame_add_davies(oqb_ame* ame, int k0, int nk, const double* h_gamma_plus, const double* h_gamma_minus,
                       const double* h_S_plus, const double* h_S_minus, double gamma0, double S0);
/* The same with γ of an Ohmic bath (src/bath/ohmic.jl:35-41) and S from a quadratic B-spline table (nS = 0 ⇒ S ≡ 0;
 * arguments as oqb_ame_set_lambshift_spline) evaluated at the keys on the device. */
int oqb_ame_add_davies_ohmic(oqb_ame* ame, int k0, int nk, double eta, double wc, double beta, int nS, double S_x0,
                             double S_dx, const double* h_S_coef);
/* One pair (α,β) of a CorrelatedDaviesGenerator / ConstHCorrelatedDaviesGenerator (src/opensys/davies.jl:231-345):
 *   du += Σ_x γ_αβ(x)(L^β_x ρ L^α_x' − ½{L^α_x'L^β_x, ρ}) − i[S_αβ(x) L^α_x'L^β_x, ρ],  L^k_x = A_k∘[ω̃ == x]
 * with γ_αβ(±uniq_w[g]), S_αβ(±uniq_w[g]) per gap key and at 0 (real-valued spectra); one call per pair of `inds`, on top of
 * oqb_ame_gaps_build; accumulates with any other term.  Degenerate gap groups are handled like for oqb_ame_add_davies. */
int oqb_ame_add_correlated(oqb_ame* ame, int alpha, int beta, const double* h_gamma_plus, const double* h_gamma_minus,
                           const double* h_S_plus, const double* h_S_minus, double gamma0, double S0);
/* What the gap structure turned into: generators added, cross-term and L'L list entries, gap groups on the dense path,
 * dense blocks kept (blocks whose masked coupling vanishes to rounding are dropped).  Any pointer may be NULL. */
int oqb_ame_davies_stats(const oqb_ame* ame, int64_t* n_terms, int64_t* n_cross, int64_t* n_lamb, int64_t* n_dense_groups,
                         int64_t* n_dense_blocks);
/* K lvl×lvl matrices v'A_αv as the device computed them (column-major, host) */
int oqb_ame_get_rotated_couplings(oqb_ame* ame, void* h_A);
int oqb_ame_get_levels(oqb_ame* ame, double* h_w); /* the lvl eigen-energies of the current eigen-system */
int oqb_ame_get_eigenvectors(oqb_ame* ame, void* h_V); /* its n×lvl eigenvector matrix (column-major, host) */
/* One-sided AME (src/opensys/davies.jl:367-379): npairs (α,β) 0-based; tables at the UNROUNDED
 * gap matrix ω_ba[i,j] = w[j]−w[i] (src/opensys/opensys_util.jl:65): h_gam/h_S are lvl×lvl per pair,
 * or one shared table when tables_shared != 0 (SingleFunctionMatrix). */
int oqb_ame_set_onesided(oqb_ame* ame, int npairs, const int32_t* h_alpha, const int32_t* h_beta,
                         const double* h_gam, const double* h_S, int tables_shared);
/* The same with the shared tables built ON THE DEVICE for an Ohmic bath: γ(w[b] − w[a]) in closed form and S from the
 * spline of oqb_ame_set_lambshift_spline (S ≡ 0 when none is set). */
int oqb_ame_set_onesided_ohmic(oqb_ame* ame, int npairs, const int32_t* h_alpha, const int32_t* h_beta, double eta,
                               double wc, double beta);
int oqb_ame_clear_onesided(oqb_ame* ame);
/* du[b] = v·( −i[diag w, ρ̃] + Σ eigenbasis terms(ρ̃) )·v',  ρ̃ = v' u[b] v   (OVERWRITES du)
 *   — (Op::DiffEqLiouvillian{true,false})(du,u,p,t) src/opensys/diffeq_liouvillian.jl:92-105.
 *   Non-eigenbasis terms (:106-108) are added by the caller with oqb_lindblad_acc. */
int oqb_ame_rhs(oqb_ame* ame, int nb, const void* d_u, void* d_du);
int oqb_ame_rhs_host(oqb_ame* ame, int nb, const void* h_u, void* h_du);
/* eigenbasis terms only, on ρ̃ given in the eigenbasis (ACCUMULATES into du):
 *   (D::DaviesGenerator)(du,ρ,gap_idx,v,s) src/opensys/davies.jl:21 /
 *   (A::OneSidedAMELiouvillian)(dρ,ρ,g_idx,v,s) :367; with_hamiltonian adds −i[diag w,ρ̃]. */
int oqb_ame_eig_acc(oqb_ame* ame, int nb, const void* d_rho_eig, void* d_du_eig, int with_hamiltonian);
/* cache (n×n) = v·(diag(−i w) − Σ_x (iS(x)+½γ(x)) L_x'L_x)·v'   (OVERWRITES)
 *   — update_cache!(cache, Op::DiffEqLiouvillian{true,false}, p, t) :111-125 with
 *     update_cache!(cache, D::DaviesGenerator, gap_idx, v, s) src/opensys/davies.jl:76-99. */
int oqb_ame_update_cache(oqb_ame* ame, void* d_cache);

/* ame_jump for a DaviesGenerator on a batch of state vectors — lindblad_jump(Op::DiffEqLiouvillian{true,false}, u, p, t)
 * src/opensys/trajectory_jump.jl:64-69 with ame_jump :14-51.  The probability slots are listed in the reference's
 * order: for every unique positive gap (ascending), for every coupling i: L₊ = sparse(a, b, σ_i[a,b]) then
 * L₋ = sparse(b, a, σ_i[b,a]); then, for every coupling, the zero-gap group.  Slot s owns the terms
 * [h_slot_ptr[s], h_slot_ptr[s+1]) of h_term = (coupling, row, col) 0-based int32 triples sorted by (row, col),
 * and h_rate[s] = γ(±ω) / γ(0).  prob[b][s] = rate_s·‖L_s v'ψ_b‖²; choice[b] by StatsBase's inverse-CDF scan with
 * d_uniform[b] (−1 where d_mask[b] == 0); apply != 0 sets ψ_b ← v L_s v'ψ_b / ‖·‖ (the √γ factor of :47 cancels).
 * d_psi: n×nb; d_prob (may be NULL): nb×nslots.  The tables depend on (w, v): set them after oqb_ame_set_eigen. */
int oqb_ame_set_jump(oqb_ame* ame, int nslots, const int64_t* h_slot_ptr, const int32_t* h_term, const double* h_rate);
int oqb_ame_clear_jump(oqb_ame* ame);
int oqb_ame_jump(oqb_ame* ame, int nb, void* d_psi, const double* d_uniform, const uint8_t* d_mask, int32_t* d_choice,
                 double* d_prob, int apply);

/* ---------------------------------------------- DiffEqLiouvillian{true,false} as one library object ---------- */
/* Everything (Op::DiffEqLiouvillian{true,false})(du,u,p,t) does, src/opensys/diffeq_liouvillian.jl:92-105, behind one call:
 * H(s) = Σ f_i(s) m_i from the DenseHamiltonian terms of `H` (borrowed; must outlive the object), haml_eigs on the device
 * (cuSOLVER Zheevd, or Dsyevd when every term matrix is real — bound with dlopen at first use), GapIndices(w, digits,
 * sigdigits), the tables of every eigenbasis term, the rotations.  Prepared eigen-systems are cached (LRU) by the
 * coefficient row f_i(s), so RK stages and schedule sweeps that revisit an s pay the O(n³) setup once.
 * lvl: kept levels (≤ 0 or > n ⇒ n).  h_couplings: the K coupling matrices of ALL eigenbasis terms, concatenated
 * (n×n each, unit-scaled); terms refer to them by index.  Constant couplings only. */
int oqb_liouv_create(oqb_ctx* ctx, oqb_dense* H, int lvl, int digits, int sigdigits, int K, const void* h_couplings,
                     oqb_liouv** out);
int oqb_liouv_destroy(oqb_liouv* L);
/* A spectrum γ(ω) or S(ω) as a C callback, evaluated on the CALLING thread during plan preparation at the distinct gap
 * keys (vectorised: out[i] = f(w[i]), i < n) — Julia passes an @cfunction pointer wrapping the user's closure. */
typedef void (*oqb_spectrum_fn)(void* user, int64_t n, const double* w, double* out);
/* push!(opensys_eig, DaviesGenerator(couplings[k0:k0+nk), γ, S)) — src/opensys/davies.jl:12-19; one call per interaction
 * (src/coupling/interaction.jl:101-117).  Ohmic: γ in closed form, S ≡ 0 (nS = 0) or the quadratic B-spline of
 * build_lambshift (arguments as oqb_ame_set_lambshift_spline), all evaluated on the device. */
int oqb_liouv_add_davies_ohmic(oqb_liouv* L, int k0, int nk, double eta, double wc, double beta, int nS, double S_x0,
                               double S_dx, const double* h_S_coef);
int oqb_liouv_add_davies_fn(oqb_liouv* L, int k0, int nk, oqb_spectrum_fn gamma, void* gamma_user, oqb_spectrum_fn S,
                            void* S_user); /* S may be NULL (≡ 0) */
/* push!(opensys_eig, OneSidedAMELiouvillian(…, inds)) — src/opensys/davies.jl:356-365; (α,β) 0-based into the couplings. */
int oqb_liouv_add_onesided_ohmic(oqb_liouv* L, int npairs, const int32_t* h_alpha, const int32_t* h_beta, double eta,
                                 double wc, double beta, int nS, double S_x0, double S_dx, const double* h_S_coef);
int oqb_liouv_add_onesided_fn(oqb_liouv* L, int npairs, const int32_t* h_alpha, const int32_t* h_beta,
                              oqb_spectrum_fn gamma, void* gamma_user, oqb_spectrum_fn S, void* S_user);
int oqb_liouv_set_plan_cache(oqb_liouv* L, int capacity); /* default 8 eigen-systems */
/* du[b] = full eigenbasis RHS (OVERWRITES).  h_coefH: the closure values f_i(s_b), [nb][nterms] (one row when
 * coef_shared): members are grouped by row, one eigen-system per distinct row. */
int oqb_liouv_rhs(oqb_liouv* L, int nb, const double* h_coefH, int coef_shared, const void* d_u, void* d_du);
int oqb_liouv_rhs_host(oqb_liouv* L, int nb, const double* h_coefH, int coef_shared, const void* h_u, void* h_du);
/* update_cache!(cache, Op::DiffEqLiouvillian{true,false}, p, t) :111-125 for one s (h_coefH: one row)  (OVERWRITES) */
int oqb_liouv_update_cache(oqb_liouv* L, const double* h_coefH, void* d_cache);
/* The prepared eigen-system for one coefficient row (built if not cached), e.g. for oqb_ame_jump / oqb_ame_gaps_fetch.
 * Owned by the object; valid until evicted. */
int oqb_liouv_prepare(oqb_liouv* L, const double* h_coefH, oqb_ame** plan);
int oqb_liouv_stats(const oqb_liouv* L, uint64_t* plans_built, uint64_t* cache_hits, double* last_prepare_ms,
                    double* last_eig_ms);
/* haml_eigs(H, s, lvl) on the device (src/hamiltonian/hamiltonian_base.jl:155-160): H(s) from one coefficient row,
 * ascending eigenvalues into d_w (n doubles) and eigenvectors into d_V (n×n column-major; the first lvl columns are the
 * truncated basis).  Real term matrices take the real symmetric solver and return exactly real eigenvectors. */
int oqb_dense_eigen(oqb_dense* op, const double* h_coefH, double* d_w, void* d_V);
int oqb_dense_info(const oqb_dense* op, int* n, int* nterms, int* K, int* terms_real);

/* ------------------------------------------- SparseHamiltonian + trajectories (ψ ensembles) */
/* Terms and Lindblad operators as Julia SparseMatrixCSC{ComplexF64,Int64}: 1-based colptr
 * (n+1) / rowval / nzval per matrix, concatenated; *_nnz_off[i] is the offset of matrix i in the
 * concatenated rowval/nzval arrays (length count+1).  src/hamiltonian/sparse_hamiltonian.jl:10-37. */
int oqb_sparse_create(oqb_ctx* ctx, int n, int nterms, const int64_t* h_colptr, const int64_t* h_rowval,
                      const void* h_nzval, const int64_t* h_nnz_off, int K, const int64_t* h_l_colptr,
                      const int64_t* h_l_rowval, const void* h_l_nzval, const int64_t* h_l_nnz_off,
                      oqb_sparse** out);
int oqb_sparse_destroy(oqb_sparse* sp);
/* union sparsity pattern of the cache (src/hamiltonian/sparse_hamiltonian.jl:33-34) as CSC,
 * 1-based, so Julia can wrap the value array in a SparseMatrixCSC: query nnz, then fetch. */
int oqb_sparse_pattern_nnz(oqb_sparse* sp, int64_t* nnz);
int oqb_sparse_pattern(oqb_sparse* sp, int64_t* h_colptr, int64_t* h_rowval);
/* nzval[b] on the union pattern = Σ_i (-i coefH[b][i]) m_i − ½ Σ_k γ_k L_k'L_k   (OVERWRITES)
 *   — update_cache!(cache, H::SparseHamiltonian, p, s) :70-75 + src/opensys/lindblad.jl:103-109 */
int oqb_sparse_update_cache(oqb_sparse* sp, int nb, const double* h_coefH, int coef_shared,
                            const double* h_gamma, int gamma_shared, void* d_nzval);
/* cache[b] (dense n²×n²) = i(Hᵀ⊗I − I⊗H)  (OVERWRITES) — update_vectorized_cache!(cache, H::SparseHamiltonian, p, s)
 *   :77-81; the reference returns a sparse matrix with the same entries (small n only). */
int oqb_sparse_update_vectorized_cache(oqb_sparse* sp, int nb, const double* h_coefH, int coef_shared,
                                       void* d_cache);
/* du[b] = −i(H u[b] − u[b] H)  (OVERWRITES)  — (h::SparseHamiltonian)(du,u,p,s) :83-91 */
int oqb_sparse_commutator(oqb_sparse* sp, int nb, const double* h_coefH, int coef_shared,
                          const void* d_u, void* d_du);
/* dψ[b] = (−iH(s_b) − ½Σ_k γ_k L_k'L_k) ψ[b]  — the `cache*u` matvec the external solver performs
 *   with the cache of src/opensys/diffeq_liouvillian.jl:78-83. */
int oqb_traj_matvec(oqb_sparse* sp, int nb, const double* h_coefH, int coef_shared,
                    const double* h_gamma, int gamma_shared, const void* d_psi, void* d_dpsi);
int oqb_traj_matvec_host(oqb_sparse* sp, int nb, const double* h_coefH, int coef_shared,
                         const double* h_gamma, int gamma_shared, const void* h_psi, void* h_dpsi);
/* Diagnostics of the sliced-ELL plans oqb_traj_matvec builds for operators without Pauli structure (csrc/traj_sell.cu):
 *   *plans = plans held by this handle, *padding = stored ÷ real entries of the newest one (1.0 = no padding). */
int oqb_traj_sell_stats(oqb_sparse* sp, int* plans, double* padding);
/* The host planner of that kernel run on its own — needs NO GPU and no context.  Input: the off-diagonal terms as 0-based
 * CSC patterns (colptr[t][n+1], rowval[t][nnz]); unit = 2 (default: rows permuted in pairs) or 1.  Output: *padding as above;
 * *phases = LDS.128 phases (8 lanes × one slot) of one matvec, *conflicting_phases = those in which two lanes gather from the
 * same 16-byte bank group (0 by construction); *identity = 1 when the natural row order was kept; *misplaced = entries
 * lost, duplicated or attached to the wrong row/column (0 for a valid plan). */
int oqb_sell_plan_probe(int n, int nterms, const int64_t* const* colptr, const int32_t* const* rowval, int unit,
                        double* padding, int64_t* phases, int64_t* conflicting_phases, int* identity, int* misplaced);
/* out[b] = ‖ψ[b]‖²  (jump condition of the external callback) */
int oqb_traj_norm2(oqb_sparse* sp, int nb, const void* d_psi, double* d_out);
/* lind_jump / lindblad_jump{false,false}: prob[b][k] = γ_k‖L_k ψ[b]‖² (src/opensys/trajectory_jump.jl:2-12),
 * choice[b] by StatsBase's inverse-CDF scan with the supplied uniform d_uniform[b];
 * d_mask (may be NULL) selects which members jump; apply != 0 additionally does
 * ψ ← L_kψ/‖L_kψ‖ in place (what the external affect! does). */
int oqb_traj_jump(oqb_sparse* sp, int nb, const double* h_gamma, int gamma_shared, void* d_psi,
                  const double* d_uniform, const uint8_t* d_mask, int32_t* d_choice, double* d_prob,
                  int apply);

/* resample (src/opensys/trajectory_jump.jl:81-88): when `opensys` / `opensys_eig` holds several Liouvillians, each proposes
 * one jump operator (its own oqb_traj_jump / oqb_ame_jump with its own uniform) and the winner is drawn with the weights
 * ‖L_m ψ‖ (first power).  oqb_traj_apply_choice: norm[b] = ‖L_{choice[b]} ψ_b‖ for the operators of `sp` (d_norm may be
 * NULL) and, when apply != 0, ψ_b ← L_kψ_b/‖L_kψ_b‖ where d_mask[b] != 0.  oqb_resample: pick[b] = sample(1:M,
 * Weights(weight[:, b])) with the supplied uniform (StatsBase's inverse-CDF scan); d_weight is M×nb, row m = the
 * candidates of Liouvillian m; pick = −1 where d_mask[b] == 0. */
int oqb_traj_apply_choice(oqb_sparse* sp, int nb, void* d_psi, const int32_t* d_choice, const uint8_t* d_mask,
                          double* d_norm, int apply);
int oqb_resample(oqb_ctx* ctx, int nb, int M, const double* d_weight, const double* d_uniform, const uint8_t* d_mask,
                 int32_t* d_pick);

/* --------------------------------------------- on-device trajectory stepping */
/* One xoshiro256++ stream per trajectory (the algorithm of Julia's default RNG; Float64 = (x >>> 11)·2⁻⁵³):
 * d_state is nb×4 uint64 in device memory, state-in/state-out.  oqb_traj_rng_init fills it from one seed
 * (splitmix64 of seed ^ b·0xD1342543DE82EF95); a caller that wants Julia's own streams uploads the task-local
 * RNG words with oqb_upload instead.  oqb_traj_rng_uniform draws one Float64 per stream (e.g. the initial
 * jump thresholds). */
int oqb_traj_rng_init(oqb_sparse* sp, int nb, uint64_t seed, uint64_t* d_state);
int oqb_traj_rng_uniform(oqb_sparse* sp, int nb, uint64_t* d_state, double* d_out);
/* nsteps fixed RK4 steps of dψ/dt = (−iH(s) − ½Σ_k γ_k L_k'L_k)ψ (the `cache*u` of the external solver with the
 * cache of src/opensys/diffeq_liouvillian.jl:78-83) with the jump loop of its callback after every step:
 * where ‖ψ_b‖² ≤ d_threshold[b], draw u (choice) and the next threshold from the member's stream, pick
 * k = lindblad_jump/lind_jump (src/opensys/trajectory_jump.jl:2-12,71-74; prob_k = γ_k‖L_kψ‖², inverse-CDF scan with u)
 * and set ψ ← L_kψ/‖L_kψ‖.  Coefficients: the host closures evaluated at the 2·nsteps+1 stage times t0 + j·dt/2:
 * h_coefH[j][rows][nterms], h_gamma[j][rows][K] (rows = 1 when shared, else nb); the jump uses γ at the step's end.
 * d_njumps[b] is incremented per jump; d_log (may be NULL) records the first max_log events of each member as
 * (step0 + step index, k) int32 pairs.  Nothing synchronises; ψ stays in HBM throughout. */
int oqb_traj_evolve(oqb_sparse* sp, int nb, int nsteps, double dt, const double* h_coefH, int coef_shared,
                    const double* h_gamma, int gamma_shared, void* d_psi, uint64_t* d_rng, double* d_threshold,
                    int32_t* d_njumps, int32_t* d_log, int max_log, int step0);

/* ---------------------------------------------------------------------------------- bath ----- */
/* S(ω) of an Ohmic bath for n frequencies (src/bath/bath_util.jl:17 → lambshift, src/integration/ext_util.jl:23-28:
 * S = −(1/2π)∫₀^∞ (γ(ω+x) − γ(ω−x))/x dx by adaptive G7K15 on QuadGK's t/(1−t) map, stop at E ≤ atol or E ≤ rtol|I|;
 * QuadGK's defaults are rtol = √eps, atol = 0): one device thread per frequency with its own segment heap
 * (at most max_segments segments).  h_err (error estimates) and h_nseg (segments used) may be NULL. */
int oqb_ohmic_lambshift(oqb_ctx* ctx, double eta, double wc, double beta, int n, const double* h_w, double rtol,
                        double atol, int max_segments, double* h_S, double* h_err, int32_t* h_nseg);
/* C(τ) of the Ohmic bath (src/bath/ohmic.jl:48-52: η(ψ₁(1 + x₂ − x₁) + ψ₁(x₁ + x₂))/β², x₁ = iτ/β, x₂ = 1/(βωc)) for n lags;
 * h_C: n interleaved complex128.  This is the cfun of the Redfield / ULE kernels (src/coupling/interaction.jl:178-186). */
int oqb_ohmic_correlation(oqb_ctx* ctx, double eta, double wc, double beta, int n, const double* h_tau, void* h_C);
/* Bath time scales (src/bath/bath_util.jl:55-89): 1/τ_SB = ∫₀^∞|C(τ)|dτ, τ_B = τ_SB·∫₀^lim τ|C(τ)|dτ with the error
 * estimates the reference returns (err/res², err·τ_SB; may be NULL); coarse_grain_timescale is √(τ_SB τ_B/5). */
int oqb_ohmic_timescales(oqb_ctx* ctx, double eta, double wc, double beta, double lim, double rtol, double atol,
                         int max_segments, double* tau_sb, double* err_sb, double* tau_b, double* err_b);
/* y[k] = spline(x[k]) for the quadratic B-spline of oqb_ame_set_lambshift_spline (host arrays; m points). */
int oqb_bspline_eval(oqb_ctx* ctx, int n, double x0, double dx, const double* h_coef, int m, const double* h_x,
                     double* h_y);

/* ------------------------------------------------------------------------- observables ----- */
/* out[b][o] = tr(O_o ρ[b]) (is_vector==0, state n×n) or ψ[b]'O_oψ[b] (is_vector!=0); complex
 * (re,im) pairs; d_obs: nobs dense n×n matrices.  Partial sums for the NCCL ensemble average. */
int oqb_expect(oqb_ctx* ctx, int n, int nobs, const void* d_obs, int nb, const void* d_state,
               int is_vector, double* d_out);

/* Per-GPU partial sums for an ensemble average: out[o] = Σ_b tr(O_o ρ[b]) (or Σ_b ψ[b]'O_oψ[b]), nobs complex (re,im) pairs in
 * device memory, summed in a fixed order (deterministic) on the context stream — the input of oqb_allreduce_obs. */
int oqb_expect_sum(oqb_ctx* ctx, int n, int nobs, const void* d_obs, int nb, const void* d_state, int is_vector,
                   double* d_out);

/* ------------------------------------------------------------- multi-GPU: the ensemble reduce ----- */
/* The only collective of the path: every ρ / ψ / schedule is independent inside an RHS call; ensemble
 * averages (C4) are one ncclAllReduce(sum, double) and per-schedule observables (C5) one ncclAllGather per output point.
 * NCCL is bound with dlopen("libnccl.so.2") when the first communicator is created.
 *  - one process or thread per GPU: rank 0 calls oqb_comm_unique_id (128 bytes), the host ships them to the other ranks
 *    (MPI, torch.distributed, Distributed.jl, a file …), every rank calls oqb_comm_init_rank with its own context;
 *  - one process driving all GPUs (a Julia process with Threads.@threads over devices): oqb_comm_init_all over one
 *    context per GPU (ncclCommInitAll); collectives issued from ONE thread for several communicators go between
 *    oqb_comm_group_start / oqb_comm_group_end.
 * Collectives are enqueued on the communicator's context stream (no host synchronisation) and work in place. */
int oqb_nccl_version(int* version);
int oqb_comm_unique_id(void* id128);
int oqb_comm_init_rank(oqb_ctx* ctx, int nranks, int rank, const void* id128, oqb_comm** out);
int oqb_comm_init_all(int ndev, oqb_ctx* const* ctxs, oqb_comm** out /* ndev handles */);
int oqb_comm_destroy(oqb_comm* comm);
int oqb_comm_rank(const oqb_comm* comm, int* rank, int* nranks);
int oqb_comm_group_start(void);
int oqb_comm_group_end(void);
int oqb_allreduce_obs(oqb_comm* comm, double* d_buf, int64_t count);                                   /* sum, in place */
int oqb_allgather_obs(oqb_comm* comm, const double* d_send, double* d_recv, int64_t count_per_rank);  /* rank-major */

/* Diagonal observables of a ψ ensemble (populations, ⟨Z_k⟩, energies of a diagonal problem Hamiltonian — what
 * inst_population src/hamiltonian/hamiltonian_base.jl:216-237 reduces to in the computational basis):
 * out[b][o] = Σ_r diag_o[r]·|ψ_b[r]|², nobs ≤ 16 real diagonals (nobs×n, device) in ONE pass over ψ. */
/* Ensemble populations P[r] = Σ_b |ψ_b[r]|² in one pass over ψ (inst_population of the ensemble, src/hamiltonian/
 * hamiltonian_base.jl:216-237, in the computational basis): every diagonal ensemble average is then d_O·P / B — the n doubles of
 * P (or the few dot products) are what the NCCL all-reduce carries.  Fixed summation order (two-stage), d_out: n doubles. */
int oqb_ensemble_population(oqb_ctx* ctx, int n, int nb, const void* d_psi, double* d_out);
int oqb_expect_diag(oqb_ctx* ctx, int n, int nobs, const double* d_diag, int nb, const void* d_psi, double* d_out);

/* ------------------------------------------------------------------------------ probes ----- */
/* Measures FP64 DMMA / DFMA issue throughput of this device (TFLOP/s) — used by bench.py to state
 * the FP64 tensor ceiling next to the cuBLAS DGEMM figure. */
int oqb_probe_fp64(oqb_ctx* ctx, double* dmma_tflops, double* dfma_tflops);
/* Measures the shared-memory gather ceiling of this device for the trajectory matvec: 16-byte gathers per second from a staged
 * 4096-entry state, with conflict-free XOR-permuted indices (pattern 0) or pseudo-random indices (pattern 1).  An operator with
 * m entries per row cannot run faster than (gathers/s ÷ m)·32 algorithmic bytes per second, whatever the rest of the kernel does
 * (why an unstructured sparse operator cannot reach the HBM roofline). */
int oqb_probe_smem_gather(oqb_ctx* ctx, int pattern, double* gathers_per_sec);
/* Per-launch CUDA-event timing of the ZGEMM kernel (the dominant kernel of every dense path) on the
 * context stream: between begin and end every ZGEMM launch is bracketed by an event pair;
 * end returns the summed kernel time, the launch count and the algorithmic flops (8·M·N·K·batch). */
int oqb_profile_begin(oqb_ctx* ctx);
int oqb_profile_end(oqb_ctx* ctx, double* zgemm_ms, uint64_t* zgemm_launches, double* zgemm_flops);

#ifdef __cplusplus
}
#endif
#endif /* OQB_H */
