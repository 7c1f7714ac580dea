// Shared declarations for the B200-native OpenQuantumBase RHS library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace oqb {

typedef double2 c128;  // interleaved (re, im) == Julia ComplexF64 == numpy complex128

enum Op : int { OP_N = 0, OP_T = 1, OP_C = 2 };

// Per-context switches (oqb_set_option / oqb_get_option).  The library reads no environment variables.
struct Options {
    int zgemm_tma = 1;        // 0: cp.async fallback ZGEMM only
    int zgemm_bm = 0;         // 0 = by K, 64 / 128 force the CTA tile height of the TMA ZGEMM
    int tma_oneshot = 1;      // one-instruction 5-D operand tiles
    int lindblad_fuse = 1;    // diagonal L_k with shared rates: one n×n factor G for the whole batch, du += G∘u as a pure streaming pass
    int lindblad_stencil = 2; // … and that term Pauli-structured (≤ 16 XOR diagonals): the commutator as a hypercube stencil fused with the Hadamard factor — no product at all (2: register micro-tiles for n ≥ 64, 1: the plain kernel)
    int lindblad_split = 1;   // … and one off-diagonal Hamiltonian term: shared-operand products scaled by f_N(s_b), diagonal terms in the Hadamard pass
    int coupling_mono = 1;    // Redfield: monomial couplings (σx, σ±, Pauli strings) as gathers (host mirrors consult it when choosing the entry point)
    int lindblad_mono = 1;    // exploit monomial L_k (one entry per row and column: σ±, Pauli strings): gather instead of 2K products
    int lindblad_diag = 1;    // exploit diagonal L_k
    int coupling_diag = 1;    // exploit diagonal couplings in the rotation v'c_αv
    int traj_kernel = 0;      // 0 = pick by operator structure, 1 = generic kernels only, 2 = no XOR specialisation
    int traj_ell = 1;         // sliced-ELL trajectory kernel (traj_sell.cu) for operators without XOR structure; 0 = the plain ELL kernels
    int sell_r = 0;           // trajectories per operator read in that kernel (0 = pick)
    int sell_unit = 2;        // rows permuted in units of 2 neighbours (full 32-byte sectors per store) or 1
    int rk_fuse = 1;          // RK4 stage arithmetic fused into the trajectory matvec
    int xor_fixed = 1;        // fully unrolled transverse-field variant of the XOR kernel
    int xor_rows = 16;        // rows per thread of the XOR kernel at n = 4096
    int host_slots = 2;       // device buffers per direction in the host-buffer pipelines (2–4; more than 2 measured no faster)
    int ame_host_chunk_mib = 128;
    int davies_dense_min_pairs = 0;        // gap groups with at least this many level pairs take the dense per-group path (0 = cost model)
    long long davies_max_cross = 50000000; // cap on the cross-term list; larger groups go dense first
    int fuse_cross = 1;       // keep the Hadamard in GEMM 2's epilogue when the only extra Davies terms are list cross terms
    int debug_checks = 0;     // device-side range / consistency checks of every index list the term builder produces
};

// "configured once per DEVICE" flag for cudaFuncSetAttribute and constant-memory uploads (several contexts / devices per
// process; setting an attribute twice is harmless, so the benign race between two contexts of one device is fine).
struct DevOnce {
    std::atomic<unsigned char> f[64];
    DevOnce() { for (auto& x : f) x.store(0); }
    bool done(int dev) const { return f[dev & 63].load(std::memory_order_acquire) != 0; }
    void set(int dev) { f[dev & 63].store(1, std::memory_order_release); }
};

struct Ctx {
    int device = 0;
    Options opt;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream[2] = {nullptr, nullptr};
    cudaEvent_t ev[16] = {};  // host_pipeline: 4 upload, 4 compute, 4 download events + 1 ordering event
    std::string err;
    uint64_t launches = 0;
    // complex product of the TMA ZGEMM: 1 = 3-multiplication (default), 0 = classic 4-multiplication
    int zgemm_3m = 1;
    // exploit ZgemmParams::real_operand hints (oqb_set_real_operand_mode; OQB_REAL_OPERAND=0 disables)
    int real_operand = 1;
    // grow-only device workspace
    void* ws = nullptr;
    size_t ws_bytes = 0;
    // pinned staging for small per-call coefficient arrays
    void* pin = nullptr;
    size_t pin_bytes = 0;
    void* dcoef = nullptr;
    size_t dcoef_bytes = 0;
    size_t coef_head = 0;
    // optional per-launch timing of the ZGEMM kernel (oqb_profile_begin / oqb_profile_end)
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;  // pairs (start, stop)
    double prof_flops = 0.0;
    // grow-only device staging for the *_host entry points (double-buffered u / du chunks)
    void* io = nullptr;
    size_t io_bytes = 0;
    // cuSOLVER handle of the device eigensolver (liouvillian.cu; the library is bound with dlopen on first use)
    void* cusolver = nullptr;
    void (*cusolver_destroy)(void*) = nullptr;
};

int set_error(Ctx* c, int code, const char* fmt, ...);

struct EntryScope {
    EntryScope(const Ctx* c, const char* name) {
        if (c) cudaSetDevice(c->device);
        nvtxRangePushA(name);
    }
    ~EntryScope() { nvtxRangePop(); }
};

// Brackets one launch of a path's dominant kernel with a CUDA-event pair while profiling is on
// (oqb_profile_begin/end); `work` is the launch's algorithmic flops or bytes.
struct ProfScope {
    Ctx* c;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    double work;
    ProfScope(Ctx* ctx, double w) : c(ctx), work(w) {
        if (c->profiling && cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess)
            cudaEventRecord(e0, c->stream);
    }
    ~ProfScope() {
        if (c->profiling && e0 && e1) {
            cudaEventRecord(e1, c->stream);
            c->prof_events.push_back(e0);
            c->prof_events.push_back(e1);
            c->prof_flops += work;
        }
    }
};
int ws_reserve(Ctx* c, size_t bytes);
int io_reserve(Ctx* c, size_t bytes);
int coef_upload(Ctx* c, const void* host, size_t bytes, void** dptr);  // async on c->stream

#define OQB_CUDA(c, call)                                                                      \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return oqb::set_error((c), 2, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                                  __FILE__, __LINE__);                                         \
    } while (0)

#define OQB_LAUNCH_CHECK(c)                                                                    \
    do {                                                                                       \
        (c)->launches++;                                                                       \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess)                                                                 \
            return oqb::set_error((c), 2, "kernel launch failed: %s (%s:%d)",                  \
                                  cudaGetErrorString(_e), __FILE__, __LINE__);                 \
    } while (0)

// First statement of every extern "C" entry point: selects the GPU of the context it was handed (several devices per
// process) and brackets the call with an NVTX range named after the entry point (visible in Nsight Systems / ncu --nvtx).
#define OQB_DEVICE(handle, ctxexpr) oqb::EntryScope _oqb_entry((handle) ? (const oqb::Ctx*)(ctxexpr) : nullptr, __func__)

#define OQB_TRY(expr)            \
    do {                         \
        int _s = (expr);         \
        if (_s != 0) return _s;  \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Batched complex GEMM on the FP64 tensor pipe (DMMA.8x8x4).  zgemm.cu
//   C[b] = epi( alpha * opA(A[b]) * opB(B[b]) ) + beta * C[b]
// All matrices column-major, interleaved complex128; strides in complex elements; a batch
// stride of 0 shares the operand across the batch.
// ---------------------------------------------------------------------------------------------
struct ZgemmParams {
    int M = 0, N = 0, K = 0, batch = 1;
    int opA = OP_N, opB = OP_N;
    // batch index z = b * batch2 + k (b < batch, k < batch2); offset = b*stride + k*stride2
    int batch2 = 1;
    const c128* A = nullptr; long long lda = 0, strideA = 0, strideA2 = 0;
    const c128* B = nullptr; long long ldb = 0, strideB = 0, strideB2 = 0;
    c128* C = nullptr;       long long ldc = 0, strideC = 0, strideC2 = 0;
    c128 alpha = {1.0, 0.0}, beta = {0.0, 0.0};
    // optional Hadamard epilogue: result(m,n) *= E(m,n)   (E col-major, lde, batch stride strideE)
    const c128* E = nullptr; long long lde = 0, strideE = 0;
    // optional: capture the diagonal of alpha*opA(A)opB(B) (before the Hadamard factor)
    c128* diag_out = nullptr; long long strideDiag = 0;
    // optional per-(b,k) real scale of alpha: alpha_z = alpha * alpha_scale[z]   (device pointer)
    //   (index z % alpha_scale_mod when alpha_scale_mod > 0)
    const double* alpha_scale = nullptr;
    int alpha_scale_mod = 0;
    // hint: 1 = every entry of A has zero imaginary part, 2 = same for B (e.g. the eigenvectors of a real-symmetric
    // Hamiltonian).  The TMA kernel then issues 2 real products per complex one (re += Ar·Br, im += Ar·Bi) instead of
    // 3 or 4; kernels that ignore the hint still give the right answer.  Only honoured when Ctx::real_operand is on.
    int real_operand = 0;
};
int zgemm(Ctx* c, const ZgemmParams& p);
int zgemm_tma(Ctx* c, const ZgemmParams& p, bool* used);  // zgemm_tma.cu

// Double-buffered H2D → compute → D2H pipeline over batch chunks (the *_host entry points).
// compute(b0, cnt, d_in, d_out) enqueues work for members [b0, b0+cnt) on c->stream.
template <class F>
int host_pipeline(Ctx* c, int nb, size_t in_bytes, size_t out_bytes, const void* h_in, void* h_out, int chunk,
                  F compute) {
    if (nb <= 0) return 0;
    if (chunk > nb) chunk = nb;
    // NBUF device slots per direction (option host_slots, default 2).  A third or fourth slot — taking the upload of chunk
    // i+2 off the download of chunk i−2 — was measured and changes nothing: C3 e2e 1 794 / 1 803 / 1 803 evals/s with
    // 2 / 3 / 4 slots on one GPU, 3 905 / 3 880 / 3 896 on eight (0.84 of the host link either way).
    const int NBUF = c->opt.host_slots >= 2 && c->opt.host_slots <= 4 ? c->opt.host_slots : 2;
    const size_t inb = (size_t)chunk * in_bytes, outb = (size_t)chunk * out_bytes;
    OQB_TRY(io_reserve(c, NBUF * (inb + outb)));
    char* base = (char*)c->io;
    char* d_in[4];
    char* d_out[4];
    for (int s = 0; s < NBUF; ++s) {
        d_in[s] = base + (size_t)s * inb;
        d_out[s] = base + (size_t)NBUF * inb + (size_t)s * outb;
    }
    cudaStream_t up = c->copy_stream[0], down = c->copy_stream[1];
    cudaEvent_t* ev_up = &c->ev[0];
    cudaEvent_t* ev_comp = &c->ev[4];
    cudaEvent_t* ev_down = &c->ev[8];
    // Chunk sizes ramp up at the start and down at the end (chunk/4, chunk/2, chunk, …, chunk/2, chunk/4): only the
    // first upload and the last download are not hidden behind compute, so they are made small.
    std::vector<int> sizes;
    {
        int left = nb;
        std::vector<int> head, tail;
        if (nb >= 4 * chunk) {
            // ramp chunk/4, chunk/2 — then the mirror image at the end (measured on C3, e2e evals/s: chunks of 8 ρ with
            // this ramp 1 775; a finer ramp starting at 1 ρ 1 738; chunks of 4 / 16 / 32 ρ 1 630 / 1 605 / 1 235)
            const int first = chunk / 4 > 0 ? chunk / 4 : 1;
            for (int sz = first; sz < chunk; sz *= 2) head.push_back(sz);
            tail.assign(head.rbegin(), head.rend());
            for (int sz : head) left -= 2 * sz;
        }
        sizes = head;
        while (left > 0) {
            const int cnt = left < chunk ? left : chunk;
            sizes.push_back(cnt);
            left -= cnt;
        }
        sizes.insert(sizes.end(), tail.begin(), tail.end());
    }
    // order the pipeline after whatever is already queued on the compute stream
    OQB_CUDA(c, cudaEventRecord(c->ev[12], c->stream));
    OQB_CUDA(c, cudaStreamWaitEvent(up, c->ev[12], 0));
    int b0 = 0;
    for (int i = 0; i < (int)sizes.size(); ++i) {
        const int cnt = sizes[i];
        const int s = i % NBUF;
        if (i >= NBUF) OQB_CUDA(c, cudaStreamWaitEvent(up, ev_comp[s], 0));
        OQB_CUDA(c, cudaMemcpyAsync(d_in[s], (const char*)h_in + (size_t)b0 * in_bytes, (size_t)cnt * in_bytes,
                                    cudaMemcpyHostToDevice, up));
        OQB_CUDA(c, cudaEventRecord(ev_up[s], up));
        OQB_CUDA(c, cudaStreamWaitEvent(c->stream, ev_up[s], 0));
        if (i >= NBUF) OQB_CUDA(c, cudaStreamWaitEvent(c->stream, ev_down[s], 0));
        OQB_TRY(compute(b0, cnt, (const void*)d_in[s], (void*)d_out[s]));
        OQB_CUDA(c, cudaEventRecord(ev_comp[s], c->stream));
        OQB_CUDA(c, cudaStreamWaitEvent(down, ev_comp[s], 0));
        OQB_CUDA(c, cudaMemcpyAsync((char*)h_out + (size_t)b0 * out_bytes, d_out[s], (size_t)cnt * out_bytes,
                                    cudaMemcpyDeviceToHost, down));
        OQB_CUDA(c, cudaEventRecord(ev_down[s], down));
        b0 += cnt;
    }
    OQB_CUDA(c, cudaStreamSynchronize(down));
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

}  // namespace oqb

// ---------------------------------------------------------------------------------------------
// elementwise / reduction kernels (elementwise.cu) — launched on c->stream
// ---------------------------------------------------------------------------------------------
namespace oqb {
// out[b][e] (+)= Σ_j coef[b*coef_ld + j] * mats[j*len + e],  e < len, b < nb   (coef complex, device)
int lincomb(Ctx* c, int nmat, long long len, const c128* mats, const c128* coef, long long coef_ld,
            c128* out, int nb, bool accumulate);
// y[b][e] += alpha · x[b*stride_x + e]
int axpy_batched(Ctx* c, long long len, int nb, c128 alpha, const c128* x, long long stride_x, c128* y);
// X[b][a+b'*l] (+)= (Cm[a+b'*l] + (w_sub ? i(w_sub[a]-w_sub[b']) : 0)) * R[b][a+b'*l]
int hadamard(Ctx* c, int l, const c128* Cm, const double* w_sub, const c128* R, c128* X, int nb,
             bool accumulate);
// X[b][a + a*l] += Σ_b' Wt[b' + a*l] * pop[b*l + b']      (Wt real)
int davies_diag(Ctx* c, int l, const double* Wt, const c128* pop, c128* X, int nb, bool times_i = false);
// pop[b*l + a] = R[b][a + a*l]
int extract_diag(Ctx* c, int l, const c128* R, c128* pop, int nb);
// X[b][i+j*l] += alpha * Σ_{k<nk} (Kb[(b*nk+k)][i+j*l] + conj(Kb[(b*nk+k)][j+i*l]))
int add_adjoint(Ctx* c, int l, int nk, double alpha, const c128* Kb, c128* X, int nb);
// X[b] −= K₂ + K₂',  K₂[i,j] = Σ_k (s_k[i] − s_k[j])·Mb[(b*nk+k)][i,j]   (Redfield apply, diagonal couplings)
int redfield_diag_combine(Ctx* c, int l, int nk, const c128* sdiag, const c128* Mb, c128* X, int nb);
// the same for monomial couplings S_k[a, perm_k[a]] = val_k[a] (pinv_k[c] = row of column c's entry, −1 = none)
int redfield_mono_combine(Ctx* c, int l, int nk, const int32_t* perm, const int32_t* pinv, const c128* val, const c128* Mb, c128* X, int nb);
// Davies tables from rotated couplings A (K l×l), gam/S (l×l real), w (l):
//   Gam[b], hls[b] (column sums), Wt (l×l, transposed, zero diagonal), Cm (l×l complex)
int davies_tables(Ctx* c, int l, int K, const c128* A, const double* gam, const double* S,
                  const double* w, int with_davies, double* Gam, double* hls, double* Wt, c128* Cm);
// cross terms: X[b][a + a2*l] += rate * Σ_α A[a,b] conj(A[a2,b2]) * R[b][b + b2*l]
int davies_cross(Ctx* c, int l, int K, const c128* A, long long ncross, const int32_t* idx,
                 const double* rate, const c128* R, c128* X, int nb);
// Moff[b + b2*l] += (-rate/2 - i S) Σ_α conj(A[a,b]) A[a,b2]   for entries (a,b,b2)
int davies_lamb_build(Ctx* c, int l, int K, const c128* A, long long nlamb, const int32_t* idx,
                      const double* rate_S, c128* Moff);
// D (l×l) = diag(-i w - i hls - Gam/2) + (Moff ? Moff : 0)
int davies_heff_diag(Ctx* c, int l, const double* w, const double* Gam, const double* hls,
                     const c128* Moff, c128* D);
// Lam[p][i+j*l] = (0.5*gam[p][i+j*l] + i S[p][i+j*l]) * A[alpha[p]][i+j*l]
int onesided_lambda(Ctx* c, int l, int npairs, const int32_t* alpha, const c128* A, const double* gam,
                    const double* S, long long table_stride, c128* Lam);
// out[b][(p*n+q) + (r*n+s)*n*n] = i(H[b][r,p] δ(q,s) − δ(p,r) H[b][q,s])
int vectorized_commutator(Ctx* c, int n, const c128* H, long long strideH, c128* out, int nb);
// cache[b] −= I⊗SL + conj(SL)⊗I − Sᵀ⊗Λ − conj(Λ)⊗S   for one coupling
int redfield_superop_acc(Ctx* c, int n, const c128* S, long long strideS, const c128* Lam,
                         long long strideLam, const c128* SL, long long strideSL, c128* cache, int nb);
// out[(b*nobs+o)] = tr(O_o ρ_b) or ψ_b' O_o ψ_b
int expect(Ctx* c, int n, int nobs, const c128* obs, int nb, const c128* state, int is_vector, c128* out);
// out[r] = Σ_b |ψ_b[r]|²  (deterministic two-stage reduction)
int ensemble_population(Ctx* c, int n, int nb, const c128* psi, double* out);
// out[b*nobs+o] = Σ_r diag[o*n+r]·|ψ_b[r]|²  (nobs ≤ 16)
int expect_diag(Ctx* c, int n, int nobs, const double* diag, int nb, const c128* psi, double* out);
}  // namespace oqb

struct oqb_ctx : oqb::Ctx {};
