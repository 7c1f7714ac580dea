// DiffEqLiouvillian{true,false} as ONE library object: everything the reference does inside
// (Op::DiffEqLiouvillian{true,false})(du, u, p, t), src/opensys/diffeq_liouvillian.jl:92-109 —
//     w, v = haml_eigs(H, s, lvl);  gap_ind = GapIndices(w, digits, sigdigits);  ρ̃ = v'uv;  −i[diag w, ρ̃];
//     every term of opensys_eig;  du = v·X·v'
// — happens behind one C call: H(s) assembled from the DenseHamiltonian terms, diagonalised on the device, GapIndices and
// the term tables built by ame_terms.cu, the batch split into groups of members that share s.  An LRU cache of prepared
// eigen-systems ("plans"), keyed by the coefficient row f_i(s), serves RK stages and schedule sweeps that revisit an s.
//
// The eigensolver is cuSOLVER's Zheevd / Dsyevd (SURVEY §8(f)-3 allows the library solver), bound at run time with
// dlopen — liboqb_b200.so itself links against nothing but the CUDA runtime.
#include <cusolverDn.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <limits>
#include <map>
#include <vector>

#include "../../include/oqb.h"
#include "ame.cuh"
#include "common.cuh"

using namespace oqb;

// ---------------------------------------------------------------------------------------------------------------------
// cuSOLVER binding
// ---------------------------------------------------------------------------------------------------------------------
namespace {

struct CusolverApi {
    void* lib = nullptr;
    decltype(&cusolverDnCreate) Create = nullptr;
    decltype(&cusolverDnDestroy) Destroy = nullptr;
    decltype(&cusolverDnSetStream) SetStream = nullptr;
    decltype(&cusolverDnZheevd_bufferSize) ZheevdBuf = nullptr;
    decltype(&cusolverDnZheevd) Zheevd = nullptr;
    decltype(&cusolverDnDsyevd_bufferSize) DsyevdBuf = nullptr;
    decltype(&cusolverDnDsyevd) Dsyevd = nullptr;
    bool ok = false;
};

CusolverApi& cusolver_api() {
    static CusolverApi api = [] {
        CusolverApi a;
        const char* names[] = {"libcusolver.so.11", "libcusolver.so.12", "libcusolver.so", "/usr/local/cuda/lib64/libcusolver.so.11"};
        for (const char* nm : names) {
            a.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (a.lib) break;
        }
        if (!a.lib) return a;
        a.Create = (decltype(a.Create))dlsym(a.lib, "cusolverDnCreate");
        a.Destroy = (decltype(a.Destroy))dlsym(a.lib, "cusolverDnDestroy");
        a.SetStream = (decltype(a.SetStream))dlsym(a.lib, "cusolverDnSetStream");
        a.ZheevdBuf = (decltype(a.ZheevdBuf))dlsym(a.lib, "cusolverDnZheevd_bufferSize");
        a.Zheevd = (decltype(a.Zheevd))dlsym(a.lib, "cusolverDnZheevd");
        a.DsyevdBuf = (decltype(a.DsyevdBuf))dlsym(a.lib, "cusolverDnDsyevd_bufferSize");
        a.Dsyevd = (decltype(a.Dsyevd))dlsym(a.lib, "cusolverDnDsyevd");
        a.ok = a.Create && a.Destroy && a.SetStream && a.ZheevdBuf && a.Zheevd && a.DsyevdBuf && a.Dsyevd;
        return a;
    }();
    return api;
}

__global__ void real_part_kernel(long long n, const c128* __restrict__ in, double* __restrict__ out) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) out[e] = in[e].x;
}
// V (n × lvl complex) from the first lvl columns of a real n × n eigenvector matrix
__global__ void real_to_complex_kernel(long long n, const double* __restrict__ in, c128* __restrict__ out) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
        out[e] = make_double2(in[e], 0.0);
}
// dst[k] = src[idx[k]] for blocks of len complex elements (grid.y = k), or the inverse scatter
__global__ void gather_blocks_kernel(long long len, const c128* __restrict__ src, c128* __restrict__ dst, const int32_t* __restrict__ idx,
                                     int scatter) {
    const long long k = blockIdx.y;
    const c128* s = src + (scatter ? k : (long long)idx[k]) * len;
    c128* d = dst + (scatter ? (long long)idx[k] : k) * len;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < len; e += (long long)gridDim.x * blockDim.x) d[e] = s[e];
}

int gather_blocks(Ctx* c, long long len, const c128* src, c128* dst, const int32_t* d_idx, int cnt, bool scatter) {
    for (int k0 = 0; k0 < cnt; k0 += 65535) {
        const int kc = cnt - k0 < 65535 ? cnt - k0 : 65535;
        dim3 grid((unsigned)std::min<long long>((len + 255) / 256, 2048), (unsigned)kc);
        gather_blocks_kernel<<<grid, 256, 0, c->stream>>>(len, scatter ? src + (long long)k0 * len : src, scatter ? dst : dst + (long long)k0 * len,
                                                          d_idx + k0, scatter ? 1 : 0);
        OQB_LAUNCH_CHECK(c);
    }
    return 0;
}

}  // namespace

namespace oqb {

// Eigen-decomposition of the Hermitian n×n matrix in d_H (column-major, destroyed) on the context stream:
// ascending eigenvalues into d_w (n), eigenvectors into d_V (n×n complex).  real_sym: no entry of H has an imaginary part
// — the real symmetric solver is used and V comes out exactly real (the RHS then takes the real-operand products).
int device_eigh(Ctx* c, int n, c128* d_H, bool real_sym, double* d_w, c128* d_V) {
    CusolverApi& api = cusolver_api();
    if (!api.ok) return set_error(c, OQB_ECUDA, "device eigensolver: libcusolver.so.11 could not be loaded (%s)", dlerror());
    if (!c->cusolver) {
        cusolverDnHandle_t h = nullptr;
        if (api.Create(&h) != CUSOLVER_STATUS_SUCCESS) return set_error(c, OQB_ECUDA, "cusolverDnCreate failed");
        c->cusolver = (void*)h;
        c->cusolver_destroy = (void (*)(void*))(void*)api.Destroy;
    }
    cusolverDnHandle_t h = (cusolverDnHandle_t)c->cusolver;
    if (api.SetStream(h, c->stream) != CUSOLVER_STATUS_SUCCESS) return set_error(c, OQB_ECUDA, "cusolverDnSetStream failed");
    const long long nn = (long long)n * n;
    int lwork = 0;
    int* d_info = nullptr;
    OQB_CUDA(c, cudaMallocAsync((void**)&d_info, sizeof(int), c->stream));
    int st = 0;
    void* d_work = nullptr;
    double* d_R = nullptr;
    do {
        cusolverStatus_t cs;
        if (real_sym) {
            if (cudaMallocAsync((void**)&d_R, (size_t)nn * sizeof(double), c->stream) != cudaSuccess) { st = set_error(c, OQB_ENOMEM, "device eigensolver: out of memory"); break; }
            real_part_kernel<<<(unsigned)std::min<long long>((nn + 255) / 256, 4096), 256, 0, c->stream>>>(nn, d_H, d_R);
            c->launches++;
            cs = api.DsyevdBuf(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, d_R, n, d_w, &lwork);
            if (cs != CUSOLVER_STATUS_SUCCESS) { st = set_error(c, OQB_ECUDA, "cusolverDnDsyevd_bufferSize failed (%d)", (int)cs); break; }
            if (cudaMallocAsync(&d_work, (size_t)lwork * sizeof(double), c->stream) != cudaSuccess) { st = set_error(c, OQB_ENOMEM, "device eigensolver: out of memory"); break; }
            cs = api.Dsyevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, d_R, n, d_w, (double*)d_work, lwork, d_info);
            if (cs != CUSOLVER_STATUS_SUCCESS) { st = set_error(c, OQB_ECUDA, "cusolverDnDsyevd failed (%d)", (int)cs); break; }
            real_to_complex_kernel<<<(unsigned)std::min<long long>((nn + 255) / 256, 4096), 256, 0, c->stream>>>(nn, d_R, d_V);
            c->launches++;
        } else {
            if (d_V != d_H) {
                if (cudaMemcpyAsync(d_V, d_H, (size_t)nn * sizeof(c128), cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess) { st = set_error(c, OQB_ECUDA, "device eigensolver: copy failed"); break; }
            }
            cs = api.ZheevdBuf(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, (cuDoubleComplex*)d_V, n, d_w, &lwork);
            if (cs != CUSOLVER_STATUS_SUCCESS) { st = set_error(c, OQB_ECUDA, "cusolverDnZheevd_bufferSize failed (%d)", (int)cs); break; }
            if (cudaMallocAsync(&d_work, (size_t)lwork * sizeof(cuDoubleComplex), c->stream) != cudaSuccess) { st = set_error(c, OQB_ENOMEM, "device eigensolver: out of memory"); break; }
            cs = api.Zheevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, (cuDoubleComplex*)d_V, n, d_w, (cuDoubleComplex*)d_work, lwork, d_info);
            if (cs != CUSOLVER_STATUS_SUCCESS) { st = set_error(c, OQB_ECUDA, "cusolverDnZheevd failed (%d)", (int)cs); break; }
        }
        int h_info = -1;
        if (cudaMemcpyAsync(&h_info, d_info, sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) { st = set_error(c, OQB_ECUDA, "device eigensolver: %s", cudaGetErrorString(cudaGetLastError())); break; }
        if (h_info != 0) st = set_error(c, OQB_ECUDA, "device eigensolver did not converge (info = %d)", h_info);
    } while (0);
    if (d_work) cudaFreeAsync(d_work, c->stream);
    if (d_R) cudaFreeAsync(d_R, c->stream);
    cudaFreeAsync(d_info, c->stream);
    return st;
}

}  // namespace oqb

// ---------------------------------------------------------------------------------------------------------------------
// the Liouvillian object
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int oqb_dense_info(const oqb_dense* op, int* n, int* nterms, int* K, int* terms_real);
extern "C" oqb_ctx* oqb_dense_ctx(const oqb_dense* op);

namespace {

enum TermKind { DAVIES_OHMIC = 0, DAVIES_FN = 1, ONESIDED_OHMIC = 2, ONESIDED_FN = 3 };

struct LTerm {
    int kind = 0;
    int k0 = 0, nk = 0;                       // Davies: coupling slice
    std::vector<int32_t> alpha, beta;         // one-sided: (α,β) pairs, 0-based into the handle's couplings
    double eta = 0, wc = 0, beta_T = 0;       // Ohmic
    int nS = 0; double S_x0 = 0, S_dx = 1;    // tabulated Lamb shift
    std::vector<double> S_coef;
    oqb_spectrum_fn gamma = nullptr, S = nullptr;
    void *gamma_user = nullptr, *S_user = nullptr;
};

struct Plan {
    std::vector<double> key;  // the coefficient row f_i(s) this eigen-system belongs to
    oqb_ame* h = nullptr;
    uint64_t stamp = 0;
};

}  // namespace

struct oqb_liouv {
    oqb_ctx* ctx = nullptr;
    oqb_dense* H = nullptr;  // borrowed
    int n = 0, nterms = 0, lvl = 0, digits = 8, sigdigits = 8, K = 0;
    bool terms_real = false;
    std::vector<c128> h_coup;  // host copy of the K couplings: every new plan handle uploads them once
    std::vector<LTerm> terms;
    std::vector<Plan> plans;
    int capacity = 8;
    uint64_t clock = 0;
    uint64_t plans_built = 0, cache_hits = 0;
    double last_prepare_ms = 0.0, last_eig_ms = 0.0;
};

namespace {

int eval_fn(oqb_spectrum_fn f, void* user, const std::vector<double>& x, std::vector<double>& out) {
    out.assign(x.size(), 0.0);
    if (f && !x.empty()) f(user, (int64_t)x.size(), x.data(), out.data());
    return 0;
}

int build_plan(oqb_liouv* L, const double* row, Plan& plan) {
    oqb_ctx* c = L->ctx;
    const int n = L->n, lvl = L->lvl;
    const long long nn = (long long)n * n;
    const auto t0 = std::chrono::steady_clock::now();
    if (!plan.h) OQB_TRY(oqb_ame_create(c, n, lvl, L->K, L->h_coup.data(), &plan.h));
    // H(s) = Σ f_i(s) m_i on the device, its eigen-system on the device (haml_eigs, src/hamiltonian/hamiltonian_base.jl:155-160)
    c128* d_H = nullptr;
    double* d_w = nullptr;
    OQB_CUDA(c, cudaMallocAsync((void**)&d_H, (size_t)2 * nn * sizeof(c128) + (size_t)n * sizeof(double), c->stream));
    c128* d_V = d_H + nn;
    d_w = (double*)(d_V + nn);
    int st = oqb_dense_hamiltonian(L->H, 1, row, 1, d_H);
    const auto e0 = std::chrono::steady_clock::now();
    if (!st) st = device_eigh(c, n, d_H, L->terms_real, d_w, d_V);
    L->last_eig_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - e0).count();
    // the first lvl eigenpairs (:157-158): V is column-major n × n, so they are its leading n·lvl elements
    if (!st) st = oqb_ame_set_eigen_device(plan.h, d_w, d_V);
    cudaStreamSynchronize(c->stream);
    cudaFreeAsync(d_H, c->stream);
    if (st) return st;
    // GapIndices + the eigenbasis terms (:96, :101-103)
    bool have_davies = false;
    for (const LTerm& t : L->terms) have_davies |= t.kind == DAVIES_OHMIC || t.kind == DAVIES_FN;
    std::vector<double> keys, mkeys, zero1(1, 0.0);
    if (have_davies) {
        int64_t ng = 0, nz = 0;
        OQB_TRY(oqb_ame_gaps_build(plan.h, L->digits, L->sigdigits, &ng, &nz));
        bool need_keys = false;
        for (const LTerm& t : L->terms) need_keys |= t.kind == DAVIES_FN;
        if (need_keys) {
            keys.resize((size_t)ng);
            OQB_TRY(oqb_ame_gaps_keys(plan.h, keys.data()));
            mkeys.resize(keys.size());
            for (size_t i = 0; i < keys.size(); ++i) mkeys[i] = -keys[i];
        }
    }
    std::vector<int32_t> os_alpha, os_beta;
    std::vector<double> os_gam, os_S;
    int os_ohmic = -1, n_os_terms = 0;
    for (size_t ti = 0; ti < L->terms.size(); ++ti) {
        const LTerm& t = L->terms[ti];
        if (t.kind == DAVIES_OHMIC) {
            OQB_TRY(oqb_ame_add_davies_ohmic(plan.h, t.k0, t.nk, t.eta, t.wc, t.beta_T, t.nS, t.S_x0, t.S_dx, t.nS ? t.S_coef.data() : nullptr));
        } else if (t.kind == DAVIES_FN) {
            std::vector<double> gp, gm, sp, sm, g0, s0;
            eval_fn(t.gamma, t.gamma_user, keys, gp);
            eval_fn(t.gamma, t.gamma_user, mkeys, gm);
            eval_fn(t.S, t.S_user, keys, sp);
            eval_fn(t.S, t.S_user, mkeys, sm);
            eval_fn(t.gamma, t.gamma_user, zero1, g0);
            eval_fn(t.S, t.S_user, zero1, s0);
            OQB_TRY(oqb_ame_add_davies(plan.h, t.k0, t.nk, gp.data(), gm.data(), sp.data(), sm.data(), g0[0], s0[0]));
        } else {
            ++n_os_terms;
            if (t.kind == ONESIDED_OHMIC) os_ohmic = (int)ti;
        }
    }
    if (n_os_terms == 1 && os_ohmic >= 0) {
        const LTerm& t = L->terms[os_ohmic];
        OQB_TRY(oqb_ame_set_lambshift_spline(plan.h, t.nS, t.S_x0, t.S_dx, t.nS ? t.S_coef.data() : nullptr));
        OQB_TRY(oqb_ame_set_onesided_ohmic(plan.h, (int)t.alpha.size(), t.alpha.data(), t.beta.data(), t.eta, t.wc, t.beta_T));
    } else if (n_os_terms > 0) {
        // several one-sided terms, or host closures: per-pair tables at the l² UNROUNDED gaps w_b − w_a (src/opensys/davies.jl:368)
        std::vector<double> w((size_t)lvl), gaps((size_t)lvl * lvl);
        OQB_TRY(oqb_ame_get_levels(plan.h, w.data()));
        for (int b = 0; b < lvl; ++b)
            for (int a = 0; a < lvl; ++a) gaps[a + (size_t)b * lvl] = w[b] - w[a];
        for (const LTerm& t : L->terms) {
            if (t.kind != ONESIDED_OHMIC && t.kind != ONESIDED_FN) continue;
            std::vector<double> g, s;
            if (t.kind == ONESIDED_FN) {
                eval_fn(t.gamma, t.gamma_user, gaps, g);
                eval_fn(t.S, t.S_user, gaps, s);
            } else {
                return set_error(c, OQB_EINVAL, "oqb_liouv: an Ohmic one-sided term cannot be combined with other one-sided terms; pass its γ as a callback");
            }
            for (size_t p = 0; p < t.alpha.size(); ++p) {
                os_alpha.push_back(t.alpha[p]);
                os_beta.push_back(t.beta[p]);
                os_gam.insert(os_gam.end(), g.begin(), g.end());
                os_S.insert(os_S.end(), s.begin(), s.end());
            }
        }
        OQB_TRY(oqb_ame_set_onesided(plan.h, (int)os_alpha.size(), os_alpha.data(), os_beta.data(), os_gam.data(), os_S.data(), 0));
    }
    L->plans_built++;
    L->last_prepare_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

// the plan for one coefficient row: LRU lookup, else build (recycling the evicted plan's handle — its couplings are already on the device)
int get_plan(oqb_liouv* L, const double* row, oqb_ame** out) {
    const size_t nt = (size_t)L->nterms;
    for (Plan& p : L->plans)
        if (p.h && !memcmp(p.key.data(), row, nt * sizeof(double))) {
            p.stamp = ++L->clock;
            L->cache_hits++;
            *out = p.h;
            return 0;
        }
    Plan* slot = nullptr;
    if ((int)L->plans.size() < L->capacity) {
        L->plans.emplace_back();
        slot = &L->plans.back();
    } else {
        slot = &L->plans[0];
        for (Plan& p : L->plans)
            if (p.stamp < slot->stamp) slot = &p;
    }
    slot->key.assign(row, row + nt);
    slot->stamp = ++L->clock;
    const int st = build_plan(L, row, *slot);
    if (st) {  // leave no half-built plan behind
        if (slot->h) oqb_ame_destroy(slot->h);
        slot->h = nullptr;
        slot->key.assign(nt, std::numeric_limits<double>::quiet_NaN());
        return st;
    }
    *out = slot->h;
    return 0;
}

// members grouped by coefficient row (the reference diagonalises once per call; a batch with several s values is a
// set of such calls).  Returns for each distinct row the member list, in first-appearance order.
void group_rows(int nb, int nt, const double* coef, std::vector<std::vector<int32_t>>& groups) {
    auto less = [nt](const double* a, const double* b) { return memcmp(a, b, (size_t)nt * sizeof(double)) < 0; };
    std::map<const double*, int, decltype(less)> seen(less);
    for (int b = 0; b < nb; ++b) {
        const double* row = coef + (size_t)b * nt;
        auto it = seen.find(row);
        if (it == seen.end()) {
            seen.emplace(row, (int)groups.size());
            groups.emplace_back(1, (int32_t)b);
        } else {
            groups[it->second].push_back((int32_t)b);
        }
    }
}

int liouv_rhs_impl(oqb_liouv* L, int nb, const double* h_coefH, int coef_shared, const c128* d_u, c128* d_du) {
    oqb_ctx* c = L->ctx;
    const long long nn = (long long)L->n * L->n;
    std::vector<std::vector<int32_t>> groups;
    if (coef_shared) {
        groups.emplace_back((size_t)nb);
        for (int b = 0; b < nb; ++b) groups[0][b] = b;
    } else {
        group_rows(nb, L->nterms, h_coefH, groups);
    }
    if ((int)groups.size() > L->capacity && L->capacity < 64) {
        // a sweep with more distinct s than cached plans would thrash the cache inside ONE call only if members of a group
        // were interleaved with other groups' — they are not (each group is served completely before the next), so no action.
    }
    for (const auto& g : groups) {
        oqb_ame* plan = nullptr;
        OQB_TRY(get_plan(L, h_coefH + (coef_shared ? 0 : (size_t)g[0] * L->nterms), &plan));
        const int cnt = (int)g.size();
        bool contiguous = true;
        for (int k = 1; k < cnt && contiguous; ++k) contiguous = g[k] == g[0] + k;
        if (contiguous) {
            OQB_TRY(oqb_ame_rhs(plan, cnt, d_u + (long long)g[0] * nn, d_du + (long long)g[0] * nn));
            continue;
        }
        // gather the group's members into a contiguous sub-batch, run, scatter back
        c128* d_sub = nullptr;
        int32_t* d_idx = nullptr;
        OQB_CUDA(c, cudaMallocAsync((void**)&d_sub, (size_t)2 * cnt * nn * sizeof(c128) + (size_t)cnt * sizeof(int32_t), c->stream));
        d_idx = (int32_t*)(d_sub + (size_t)2 * cnt * nn);
        int st = 0;
        if (cudaMemcpyAsync(d_idx, g.data(), (size_t)cnt * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
            st = set_error(c, OQB_ECUDA, "oqb_liouv_rhs: index upload failed");
        if (!st) st = gather_blocks(c, nn, d_u, d_sub, d_idx, cnt, false);
        if (!st) st = oqb_ame_rhs(plan, cnt, d_sub, d_sub + (size_t)cnt * nn);
        if (!st) st = gather_blocks(c, nn, d_sub + (size_t)cnt * nn, d_du, d_idx, cnt, true);
        cudaStreamSynchronize(c->stream);  // g (the copy source) lives until the end of this iteration
        cudaFreeAsync(d_sub, c->stream);
        if (st) return st;
    }
    return 0;
}

}  // namespace

extern "C" {

int oqb_dense_eigen(oqb_dense* op, const double* h_coefH, double* d_w, void* d_V) {
    if (!op) return OQB_EINVAL;
    int n = 0, nt = 0, K = 0, real = 0;
    OQB_TRY(oqb_dense_info(op, &n, &nt, &K, &real));
    oqb_ctx* c = oqb_dense_ctx(op);
    OQB_DEVICE(op, c);
    if (!h_coefH || !d_w || !d_V) return set_error(c, OQB_EINVAL, "oqb_dense_eigen: null argument");
    c128* d_H = nullptr;
    OQB_CUDA(c, cudaMallocAsync((void**)&d_H, (size_t)n * n * sizeof(c128), c->stream));
    int st = oqb_dense_hamiltonian(op, 1, h_coefH, 1, d_H);
    if (!st) st = device_eigh(c, n, d_H, real != 0, d_w, (c128*)d_V);
    cudaFreeAsync(d_H, c->stream);
    return st;
}

int oqb_liouv_create(oqb_ctx* c, oqb_dense* H, int lvl, int digits, int sigdigits, int K, const void* h_couplings, oqb_liouv** out) {
    if (!c || !out) return OQB_EINVAL;
    OQB_DEVICE(c, c);
    *out = nullptr;
    if (!H) return set_error(c, OQB_EINVAL, "oqb_liouv_create: null Hamiltonian");
    int n = 0, nt = 0, Kl = 0, real = 0;
    OQB_TRY(oqb_dense_info(H, &n, &nt, &Kl, &real));
    if (K < 0 || (K > 0 && !h_couplings)) return set_error(c, OQB_EINVAL, "oqb_liouv_create: bad couplings");
    oqb_liouv* L = new oqb_liouv();
    L->ctx = c; L->H = H; L->n = n; L->nterms = nt; L->K = K;
    L->lvl = lvl <= 0 || lvl > n ? n : lvl;  // lvl = min(size, lvl), src/opensys/diffeq_liouvillian.jl:50
    L->digits = digits; L->sigdigits = sigdigits;
    L->terms_real = real != 0;
    L->h_coup.assign((const c128*)h_couplings, (const c128*)h_couplings + (size_t)K * n * n);
    *out = L;
    return 0;
}

int oqb_liouv_destroy(oqb_liouv* L) {
    if (!L) return 0;
    OQB_DEVICE(L, L->ctx);
    cudaStreamSynchronize(L->ctx->stream);
    for (Plan& p : L->plans)
        if (p.h) oqb_ame_destroy(p.h);
    delete L;
    return 0;
}

static int check_slice(oqb_liouv* L, int k0, int nk, const char* who) {
    if (k0 < 0 || nk <= 0 || k0 + nk > L->K) return set_error(L->ctx, OQB_EINVAL, "%s: coupling slice [%d, %d) outside 0..%d", who, k0, k0 + nk, L->K);
    return 0;
}

static void drop_plans(oqb_liouv* L) {  // the term list changed: every prepared eigen-system is stale
    cudaStreamSynchronize(L->ctx->stream);
    for (Plan& p : L->plans)
        if (p.h) oqb_ame_destroy(p.h);
    L->plans.clear();
}

int oqb_liouv_add_davies_ohmic(oqb_liouv* L, int k0, int nk, double eta, double wc, double beta, int nS, double S_x0, double S_dx,
                               const double* h_S_coef) {
    if (!L) return OQB_EINVAL;
    OQB_DEVICE(L, L->ctx);
    OQB_TRY(check_slice(L, k0, nk, "oqb_liouv_add_davies_ohmic"));
    if (nS < 0 || (nS > 0 && (!h_S_coef || !(S_dx > 0.0)))) return set_error(L->ctx, OQB_EINVAL, "oqb_liouv_add_davies_ohmic: bad Lamb-shift spline");
    LTerm t;
    t.kind = DAVIES_OHMIC; t.k0 = k0; t.nk = nk; t.eta = eta; t.wc = wc; t.beta_T = beta;
    t.nS = nS; t.S_x0 = S_x0; t.S_dx = S_dx;
    if (nS) t.S_coef.assign(h_S_coef, h_S_coef + nS + 2);
    drop_plans(L);
    L->terms.push_back(t);
    return 0;
}

int oqb_liouv_add_davies_fn(oqb_liouv* L, int k0, int nk, oqb_spectrum_fn gamma, void* gamma_user, oqb_spectrum_fn S, void* S_user) {
    if (!L) return OQB_EINVAL;
    OQB_DEVICE(L, L->ctx);
    OQB_TRY(check_slice(L, k0, nk, "oqb_liouv_add_davies_fn"));
    if (!gamma) return set_error(L->ctx, OQB_EINVAL, "oqb_liouv_add_davies_fn: null spectrum callback");
    LTerm t;
    t.kind = DAVIES_FN; t.k0 = k0; t.nk = nk;
    t.gamma = gamma; t.gamma_user = gamma_user; t.S = S; t.S_user = S_user;
    drop_plans(L);
    L->terms.push_back(t);
    return 0;
}

static int check_pairs(oqb_liouv* L, int npairs, const int32_t* a, const int32_t* b, const char* who) {
    if (npairs <= 0 || !a || !b) return set_error(L->ctx, OQB_EINVAL, "%s: bad pair list", who);
    for (int p = 0; p < npairs; ++p)
        if (a[p] < 0 || a[p] >= L->K || b[p] < 0 || b[p] >= L->K) return set_error(L->ctx, OQB_EINVAL, "%s: coupling index out of range", who);
    return 0;
}

int oqb_liouv_add_onesided_ohmic(oqb_liouv* L, int npairs, const int32_t* h_alpha, const int32_t* h_beta, double eta, double wc,
                                 double beta, int nS, double S_x0, double S_dx, const double* h_S_coef) {
    if (!L) return OQB_EINVAL;
    OQB_DEVICE(L, L->ctx);
    OQB_TRY(check_pairs(L, npairs, h_alpha, h_beta, "oqb_liouv_add_onesided_ohmic"));
    LTerm t;
    t.kind = ONESIDED_OHMIC; t.eta = eta; t.wc = wc; t.beta_T = beta;
    t.alpha.assign(h_alpha, h_alpha + npairs); t.beta.assign(h_beta, h_beta + npairs);
    t.nS = nS; t.S_x0 = S_x0; t.S_dx = S_dx;
    if (nS) t.S_coef.assign(h_S_coef, h_S_coef + nS + 2);
    drop_plans(L);
    L->terms.push_back(t);
    return 0;
}

int oqb_liouv_add_onesided_fn(oqb_liouv* L, int npairs, const int32_t* h_alpha, const int32_t* h_beta, oqb_spectrum_fn gamma,
                              void* gamma_user, oqb_spectrum_fn S, void* S_user) {
    if (!L) return OQB_EINVAL;
    OQB_DEVICE(L, L->ctx);
    OQB_TRY(check_pairs(L, npairs, h_alpha, h_beta, "oqb_liouv_add_onesided_fn"));
    if (!gamma) return set_error(L->ctx, OQB_EINVAL, "oqb_liouv_add_onesided_fn: null spectrum callback");
    LTerm t;
    t.kind = ONESIDED_FN;
    t.alpha.assign(h_alpha, h_alpha + npairs); t.beta.assign(h_beta, h_beta + npairs);
    t.gamma = gamma; t.gamma_user = gamma_user; t.S = S; t.S_user = S_user;
    drop_plans(L);
    L->terms.push_back(t);
    return 0;
}

int oqb_liouv_set_plan_cache(oqb_liouv* L, int capacity) {
    if (!L) return OQB_EINVAL;
    OQB_DEVICE(L, L->ctx);
    if (capacity < 1) return set_error(L->ctx, OQB_EINVAL, "oqb_liouv_set_plan_cache: capacity must be ≥ 1");
    if (capacity < (int)L->plans.size()) drop_plans(L);
    L->capacity = capacity;
    return 0;
}

int oqb_liouv_prepare(oqb_liouv* L, const double* h_coefH, oqb_ame** plan) {
    if (!L) return OQB_EINVAL;
    OQB_DEVICE(L, L->ctx);
    if (!h_coefH) return set_error(L->ctx, OQB_EINVAL, "oqb_liouv_prepare: null coefficient row");
    oqb_ame* p = nullptr;
    OQB_TRY(get_plan(L, h_coefH, &p));
    if (plan) *plan = p;
    return 0;
}

int oqb_liouv_rhs(oqb_liouv* L, int nb, const double* h_coefH, int coef_shared, const void* d_u, void* d_du) {
    if (!L) return OQB_EINVAL;
    OQB_DEVICE(L, L->ctx);
    if (nb <= 0) return 0;
    if (!h_coefH || !d_u || !d_du) return set_error(L->ctx, OQB_EINVAL, "oqb_liouv_rhs: null argument");
    return liouv_rhs_impl(L, nb, h_coefH, coef_shared, (const c128*)d_u, (c128*)d_du);
}

int oqb_liouv_rhs_host(oqb_liouv* L, int nb, const double* h_coefH, int coef_shared, const void* h_u, void* h_du) {
    if (!L) return OQB_EINVAL;
    OQB_DEVICE(L, L->ctx);
    if (nb <= 0) return 0;
    if (!h_coefH || !h_u || !h_du) return set_error(L->ctx, OQB_EINVAL, "oqb_liouv_rhs_host: null argument");
    const size_t bytes = (size_t)L->n * L->n * sizeof(c128);
    std::vector<std::vector<int32_t>> groups;
    if (coef_shared) {
        groups.emplace_back((size_t)nb);
        for (int b = 0; b < nb; ++b) groups[0][b] = b;
    } else {
        group_rows(nb, L->nterms, h_coefH, groups);
    }
    for (const auto& g : groups) {
        oqb_ame* plan = nullptr;
        OQB_TRY(get_plan(L, h_coefH + (coef_shared ? 0 : (size_t)g[0] * L->nterms), &plan));
        // runs of consecutive members go through the pipelined host entry point in one piece
        size_t k = 0;
        while (k < g.size()) {
            size_t e = k + 1;
            while (e < g.size() && g[e] == g[e - 1] + 1) ++e;
            OQB_TRY(oqb_ame_rhs_host(plan, (int)(e - k), (const char*)h_u + (size_t)g[k] * bytes, (char*)h_du + (size_t)g[k] * bytes));
            k = e;
        }
    }
    return 0;
}

int oqb_liouv_update_cache(oqb_liouv* L, const double* h_coefH, void* d_cache) {
    if (!L) return OQB_EINVAL;
    OQB_DEVICE(L, L->ctx);
    if (!h_coefH || !d_cache) return set_error(L->ctx, OQB_EINVAL, "oqb_liouv_update_cache: null argument");
    oqb_ame* plan = nullptr;
    OQB_TRY(get_plan(L, h_coefH, &plan));
    return oqb_ame_update_cache(plan, d_cache);
}

int oqb_liouv_stats(const oqb_liouv* L, uint64_t* plans_built, uint64_t* cache_hits, double* last_prepare_ms, double* last_eig_ms) {
    if (!L) return OQB_EINVAL;
    if (plans_built) *plans_built = L->plans_built;
    if (cache_hits) *cache_hits = L->cache_hits;
    if (last_prepare_ms) *last_prepare_ms = L->last_prepare_ms;
    if (last_eig_ms) *last_eig_ms = L->last_eig_ms;
    return 0;
}

}  // extern "C"
