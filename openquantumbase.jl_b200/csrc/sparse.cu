// SparseHamiltonian + quantum-trajectory kernels (include/oqb.h, "SparseHamiltonian + trajectories").
//
// Reference arithmetic: src/hamiltonian/sparse_hamiltonian.jl:59-91 (H(s), update_cache!, commutator),
// src/opensys/lindblad.jl:103-109 (cache −= ½γL'L), src/opensys/trajectory_jump.jl:2-12,71-74,81-88
// (lind_jump / lindblad_jump).  The `cache*u` matvec and the `u ← Lu/‖Lu‖` update are done by the
// external solver in the reference; here they are the ensemble kernels K7/K8.
//
// Layout in HBM: ψ ensemble n×nb column-major (one trajectory = 16n contiguous bytes).  Operator
// terms are tiny (L2-resident) and shared by the whole ensemble; each is kept either as a diagonal
// vector or in column-major ELL (entry j of every row contiguous ⇒ coalesced for thread-per-row).
// Kernel shape: one CTA owns one trajectory at a time, stages ψ in shared memory with coalesced
// 16-byte loads, every thread produces rows r, r+T, … from shared memory and writes dψ coalesced:
// DRAM traffic is exactly read-ψ + write-dψ = 32n bytes per trajectory.
#include <algorithm>
#include <cstring>
#include <map>
#include <vector>

#include "../../include/oqb.h"
#include "common.cuh"
#include "sparse_terms.cuh"

using namespace oqb;

namespace {

constexpr int kMaxTerms = 24;

struct HostCsc {  // 0-based, rows sorted within a column
    int n = 0;
    std::vector<int64_t> colptr;
    std::vector<int32_t> row;
    std::vector<c128> val;
};

struct TermDesc {  // POD passed to kernels
    const c128* vals;
    const int32_t* cols;
    int is_diag, width;
};
struct TermList {
    TermDesc t[kMaxTerms];
    int count;
};

HostCsc parse_csc(int n, const int64_t* colptr, const int64_t* rowval, const c128* nzval) {
    HostCsc m;
    m.n = n;
    m.colptr.resize(n + 1);
    const int64_t base = colptr[0];  // 1 for Julia
    for (int j = 0; j <= n; ++j) m.colptr[j] = colptr[j] - base;
    const int64_t nnz = m.colptr[n];
    m.row.resize(nnz);
    m.val.resize(nnz);
    for (int j = 0; j < n; ++j) {
        std::vector<std::pair<int32_t, c128>> col;
        for (int64_t p = m.colptr[j]; p < m.colptr[j + 1]; ++p) col.push_back({(int32_t)(rowval[p] - 1), nzval[p]});
        std::sort(col.begin(), col.end(), [](const std::pair<int32_t, c128>& a, const std::pair<int32_t, c128>& b) { return a.first < b.first; });
        for (size_t k = 0; k < col.size(); ++k) {
            m.row[m.colptr[j] + k] = col[k].first;
            m.val[m.colptr[j] + k] = col[k].second;
        }
    }
    return m;
}

// L' L for a sparse L (setup-time only)
HostCsc ldagl(const HostCsc& L) {
    const int n = L.n;
    HostCsc r;
    r.n = n;
    r.colptr.assign(n + 1, 0);
    // (L'L)[i,j] = Σ_k conj(L[k,i]) L[k,j]; build row-lists of L first
    std::vector<std::vector<std::pair<int32_t, c128>>> rows(n);
    for (int j = 0; j < n; ++j)
        for (int64_t p = L.colptr[j]; p < L.colptr[j + 1]; ++p) rows[L.row[p]].push_back({j, L.val[p]});
    for (int j = 0; j < n; ++j) {
        std::map<int32_t, c128> acc;
        for (int64_t p = L.colptr[j]; p < L.colptr[j + 1]; ++p) {
            const int k = L.row[p];
            const c128 lkj = L.val[p];
            for (auto& e : rows[k]) {  // L[k,i]
                c128& a = acc[e.first];
                a.x += e.second.x * lkj.x + e.second.y * lkj.y;
                a.y += e.second.x * lkj.y - e.second.y * lkj.x;
            }
        }
        for (auto& e : acc) { r.row.push_back(e.first); r.val.push_back(e.second); }
        r.colptr[j + 1] = (int64_t)r.row.size();
    }
    return r;
}

}  // namespace

struct oqb_sparse {
    oqb_ctx* ctx = nullptr;
    int n = 0, nterms = 0, K = 0;
    // union pattern (CSC order = Julia nzval order) + CSR view
    int64_t nnzU = 0;
    std::vector<int64_t> u_colptr;  // 0-based
    std::vector<int32_t> u_row;
    int32_t *d_u_colptr = nullptr, *d_u_row = nullptr;            // CSC (int32)
    int32_t *d_r_rowptr = nullptr, *d_r_col = nullptr, *d_r_perm = nullptr;  // CSR + index into CSC order
    c128* d_termvals = nullptr;  // (nterms + K) × nnzU on the union pattern: m_i …, L_k'L_k …
    // matvec representation
    std::vector<TermDev> h_terms;    // nterms
    std::vector<TermDev> ldl_terms;  // K
    std::vector<TermDev> l_terms;    // K (the L_k themselves, for the jump kernel)
    TermDev ldl_comb;                // −½Σγ_k L_k'L_k on the L'L-union pattern (shared-γ fast path)
    c128* d_ldl_src = nullptr;       // K × (comb entries): L_k'L_k expanded on the comb pattern
    long long comb_len = 0;
    bool ldl_all_const = false;      // every L_k'L_k is c_k·I (e.g. Pauli jump operators)
    std::vector<double> ldl_const;   // the c_k
    SellCache* sell = nullptr;       // sliced-ELL plans of the term lists seen so far (traj_sell.cu)
};

namespace {

void free_term(TermDev& t) {
    if (t.vals) cudaFree(t.vals);
    if (t.cols) cudaFree(t.cols);
    if (t.rvals) cudaFree(t.rvals);
    if (t.absq) cudaFree(t.absq);
    t.absq = nullptr;
    if (t.cols16) cudaFree(t.cols16);
    t.vals = nullptr;
    t.cols = nullptr;
    t.rvals = nullptr;
    t.cols16 = nullptr;
}

// Build the device representation (diag or ELL) for a CSC matrix; `pattern_of` (optional) forces
// the structure of another matrix list's union so that several matrices share one layout.
int make_term(oqb_ctx* c, const HostCsc& m, TermDev& out) {
    const int n = m.n;
    bool diag = true, real = true;
    std::vector<int> cnt(n, 0);
    for (int j = 0; j < n; ++j)
        for (int64_t p = m.colptr[j]; p < m.colptr[j + 1]; ++p) {
            if (m.row[p] != j) diag = false;
            if (m.val[p].y != 0.0) real = false;
            cnt[m.row[p]]++;
        }
    out.is_real = real ? 1 : 0;
    if (diag) {
        std::vector<c128> d(n, make_double2(0.0, 0.0));
        for (int j = 0; j < n; ++j)
            for (int64_t p = m.colptr[j]; p < m.colptr[j + 1]; ++p) d[j] = m.val[p];
        out.is_diag = 1;
        out.width = 1;
        OQB_CUDA(c, cudaMalloc((void**)&out.vals, (size_t)n * sizeof(c128)));
        OQB_CUDA(c, cudaMemcpy(out.vals, d.data(), (size_t)n * sizeof(c128), cudaMemcpyHostToDevice));
        {
            std::vector<double> aq(n);
            out.absq_const = true;
            for (int j = 0; j < n; ++j) {
                aq[j] = d[j].x * d[j].x + d[j].y * d[j].y;
                if (aq[j] != aq[0]) out.absq_const = false;
            }
            out.absq_val = n ? aq[0] : 0.0;
            OQB_CUDA(c, cudaMalloc((void**)&out.absq, (size_t)n * sizeof(double)));
            OQB_CUDA(c, cudaMemcpy(out.absq, aq.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
        }
        if (real) {
            std::vector<double> dr(n);
            for (int j = 0; j < n; ++j) dr[j] = d[j].x;
            OQB_CUDA(c, cudaMalloc((void**)&out.rvals, (size_t)n * sizeof(double)));
            OQB_CUDA(c, cudaMemcpy(out.rvals, dr.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
        }
        return 0;
    }
    int width = 0;
    for (int r = 0; r < n; ++r) width = std::max(width, cnt[r]);
    std::vector<c128> v((size_t)width * n, make_double2(0.0, 0.0));
    std::vector<int32_t> ci((size_t)width * n);
    for (int r = 0; r < n; ++r)
        for (int j = 0; j < width; ++j) ci[(size_t)j * n + r] = r;  // padding reads ψ[r] * 0
    std::vector<int> fill(n, 0);
    for (int j = 0; j < n; ++j)
        for (int64_t p = m.colptr[j]; p < m.colptr[j + 1]; ++p) {
            const int r = m.row[p];
            const int slot = fill[r]++;
            v[(size_t)slot * n + r] = m.val[p];
            ci[(size_t)slot * n + r] = j;
        }
    out.is_diag = 0;
    out.width = width;
    out.h_cols = ci;
    out.h_cnt = cnt;
    out.ell_real = out.ell_imag = true;
    for (const c128& e : m.val) {
        if (e.y != 0.0) out.ell_real = false;
        if (e.x != 0.0) out.ell_imag = false;
    }
    OQB_CUDA(c, cudaMalloc((void**)&out.vals, v.size() * sizeof(c128)));
    OQB_CUDA(c, cudaMalloc((void**)&out.cols, ci.size() * sizeof(int32_t)));
    OQB_CUDA(c, cudaMemcpy(out.vals, v.data(), v.size() * sizeof(c128), cudaMemcpyHostToDevice));
    OQB_CUDA(c, cudaMemcpy(out.cols, ci.data(), ci.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    // XOR form: group the entries by mask = row ^ col.  For every mask the row-dependent value v(r) = M[r, r^mask] (0 where the
    // entry is absent) must take at most TWO values, chosen by the parity of popcount(r & z) for some bit set z — one Pauli
    // string (the two values are ±c: X/Y/Z factors), or two strings sharing their X-type factors (XX + YY, X + XZ …).
    if ((n & (n - 1)) == 0) {
        std::map<int, std::vector<c128>> byMask;
        bool ok = true;
        for (int j = 0; j < n && ok; ++j)
            for (int64_t p = m.colptr[j]; p < m.colptr[j + 1] && ok; ++p) {
                const int mask = m.row[p] ^ j;
                auto it = byMask.find(mask);
                if (it == byMask.end()) {
                    if ((int)byMask.size() >= 28) { ok = false; break; }
                    it = byMask.emplace(mask, std::vector<c128>(n, make_double2(0.0, 0.0))).first;
                }
                it->second[m.row[p]] = m.val[p];
            }
        std::vector<int> masks, zms;
        std::vector<c128> vals, vals1;
        auto same = [](c128 x, c128 y) { return x.x == y.x && x.y == y.y; };
        for (auto& e : byMask) {
            if (!ok) break;
            const std::vector<c128>& v = e.second;
            const c128 p0 = v[0];
            int z = 0;
            for (int bit = 1; bit < n; bit <<= 1)
                if (!same(v[bit], p0)) z |= bit;
            const c128 p1 = z ? v[z & -z] : p0;
            for (int r = 0; r < n && ok; ++r) ok = same(v[r], (__builtin_popcount(r & z) & 1) ? p1 : p0);
            if (!ok) break;
            masks.push_back(e.first); vals.push_back(p0); vals1.push_back(p1); zms.push_back(z);
        }
        if (ok && !masks.empty()) {
            out.xor_form = true;
            out.xor_masks = masks;
            out.xor_vals = vals;
            out.xor_vals1 = vals1;
            out.xor_zmasks = zms;
        }
    }
    // compressed forms for the fast trajectory kernel
    bool full = true;
    for (int r = 0; r < n; ++r) full = full && cnt[r] == width;
    if (full && n <= 65536) {
        std::vector<uint16_t> c16(ci.size());
        for (size_t i = 0; i < ci.size(); ++i) c16[i] = (uint16_t)ci[i];
        OQB_CUDA(c, cudaMalloc((void**)&out.cols16, c16.size() * sizeof(uint16_t)));
        OQB_CUDA(c, cudaMemcpy(out.cols16, c16.data(), c16.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        out.slot_const = true;
        out.slot_vals.resize(width);
        for (int j = 0; j < width && out.slot_const; ++j) {
            out.slot_vals[j] = v[(size_t)j * n];
            for (int r = 1; r < n; ++r)
                if (v[(size_t)j * n + r].x != out.slot_vals[j].x || v[(size_t)j * n + r].y != out.slot_vals[j].y) {
                    out.slot_const = false;
                    break;
                }
        }
    }
    return 0;
}

HostCsc union_pattern(const std::vector<const HostCsc*>& ms, int n) {
    HostCsc u;
    u.n = n;
    u.colptr.assign(n + 1, 0);
    for (int j = 0; j < n; ++j) {
        std::vector<int32_t> rows;
        for (auto* m : ms) rows.insert(rows.end(), m->row.begin() + m->colptr[j], m->row.begin() + m->colptr[j + 1]);
        std::sort(rows.begin(), rows.end());
        rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
        u.row.insert(u.row.end(), rows.begin(), rows.end());
        u.colptr[j + 1] = (int64_t)u.row.size();
    }
    u.val.assign(u.row.size(), make_double2(0.0, 0.0));
    return u;
}

// values of m expanded on the pattern u (zeros where absent)
std::vector<c128> expand_on(const HostCsc& u, const HostCsc& m) {
    std::vector<c128> v(u.row.size(), make_double2(0.0, 0.0));
    for (int j = 0; j < u.n; ++j) {
        int64_t q = u.colptr[j];
        for (int64_t p = m.colptr[j]; p < m.colptr[j + 1]; ++p) {
            while (u.row[q] != m.row[p]) ++q;
            v[q] = m.val[p];
        }
    }
    return v;
}

// ------------------------------------------------------------------------------------------
// K7: ensemble matvec  dψ_b = Σ_t c_bt M_t ψ_b
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ c128 term_row(const TermDesc& td, const c128* __restrict__ psi, int r, int n) {
    if (td.is_diag) {
        const c128 d = td.vals[r], x = psi[r];
        return make_double2(d.x * x.x - d.y * x.y, d.x * x.y + d.y * x.x);
    }
    double sr = 0.0, si = 0.0;
    for (int j = 0; j < td.width; ++j) {
        const c128 v = td.vals[(size_t)j * n + r];
        const c128 x = psi[td.cols[(size_t)j * n + r]];
        sr = fma(v.x, x.x, sr); sr = fma(-v.y, x.y, sr);
        si = fma(v.x, x.y, si); si = fma(v.y, x.x, si);
    }
    return make_double2(sr, si);
}

template <bool SMEM>
__global__ void __launch_bounds__(512) traj_matvec_kernel(int n, int nb, TermList tl, const c128* __restrict__ coef,
                                                         long long coef_ld, const c128* __restrict__ psi,
                                                         c128* __restrict__ dpsi) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c128* sp = reinterpret_cast<c128*>(smem_raw);
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        const c128* pb = psi + (size_t)b * n;
        const c128* x = pb;
        if (SMEM) {
            __syncthreads();  // previous trajectory fully consumed
            for (int r = threadIdx.x; r < n; r += blockDim.x) sp[r] = pb[r];
            __syncthreads();
            x = sp;
        }
        const c128* cf = coef + b * coef_ld;
        for (int r = threadIdx.x; r < n; r += blockDim.x) {
            double ar = 0.0, ai = 0.0;
            for (int t = 0; t < tl.count; ++t) {
                const c128 y = term_row(tl.t[t], x, r, n);
                const c128 cc = cf[t];
                ar = fma(cc.x, y.x, ar); ar = fma(-cc.y, y.y, ar);
                ai = fma(cc.x, y.y, ai); ai = fma(cc.y, y.x, ai);
            }
            dpsi[(size_t)b * n + r] = make_double2(ar, ai);
        }
    }
}

// combined dissipator values: out[e] = Σ_k (−½γ_k) src[k*len + e]
__global__ void ldl_combine_kernel(int K, long long len, const double* __restrict__ gamma, const c128* __restrict__ src,
                                   c128* __restrict__ out, double* __restrict__ out_real) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= len) return;
    double sr = 0.0, si = 0.0;
    for (int k = 0; k < K; ++k) {
        const double g = -0.5 * gamma[k];
        const c128 v = src[(size_t)k * len + e];
        sr = fma(g, v.x, sr);
        si = fma(g, v.y, si);
    }
    out[e] = make_double2(sr, si);
    if (out_real) out_real[e] = sr;  // diagonal of a Hermitian L'L is real: compressed copy for the fast kernel
}

__global__ void norm2_kernel(int n, const c128* __restrict__ psi, double* __restrict__ out) {
    const int b = blockIdx.x;
    const c128* p = psi + (size_t)b * n;
    double s = 0.0;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        const c128 v = p[r];
        s = fma(v.x, v.x, s);
        s = fma(v.y, v.y, s);
    }
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int i = 0; i < (blockDim.x >> 5); ++i) a += red[i];
        out[b] = a;
    }
}

// ------------------------------------------------------------------------------------------
// K8: jump decision  (src/opensys/trajectory_jump.jl:2-12 + StatsBase inverse-CDF scan)
// ------------------------------------------------------------------------------------------
template <bool SMEM>
__global__ void __launch_bounds__(512) traj_jump_kernel(int n, int nb, TermList tl, const double* __restrict__ gamma,
                                                       long long gamma_ld, c128* __restrict__ psi,
                                                       const double* __restrict__ uniform,
                                                       const unsigned char* __restrict__ mask, int32_t* __restrict__ choice,
                                                       double* __restrict__ prob, int apply) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c128* sp = reinterpret_cast<c128*>(smem_raw);
    __shared__ double red[kMaxTerms][32];
    __shared__ double s_norm2[kMaxTerms];
    __shared__ int s_choice;
    const int K = tl.count;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        if (mask && !mask[b]) {
            if (threadIdx.x == 0 && choice) choice[b] = -1;
            continue;
        }
        c128* pb = psi + (size_t)b * n;
        const c128* x = pb;
        __syncthreads();
        if (SMEM) {
            for (int r = threadIdx.x; r < n; r += blockDim.x) sp[r] = pb[r];
            __syncthreads();
            x = sp;
        }
        // ‖L_k ψ‖² for every k: fixed-order tree reduction ⇒ run-to-run reproducible
        for (int k = 0; k < K; ++k) {
            double s = 0.0;
            for (int r = threadIdx.x; r < n; r += blockDim.x) {
                const c128 y = term_row(tl.t[k], x, r, n);
                s = fma(y.x, y.x, s);
                s = fma(y.y, y.y, s);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) red[k][warp] = s;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const double* g = gamma + b * gamma_ld;
            double total = 0.0;
            for (int k = 0; k < K; ++k) {
                double a = 0.0;
                for (int i = 0; i < nwarp; ++i) a += red[k][i];
                s_norm2[k] = a;
                const double pk = g[k] * a;  // γ_k · norm(L_k u)^2
                red[k][0] = pk;
                if (prob) prob[(size_t)b * K + k] = pk;
                total += pk;  // Weights(prob).sum, left to right
            }
            // StatsBase.sample(rng, wv): t = rand()*sum; i=1; cw=w[1]; while cw < t && i < n …
            const double tthr = uniform[b] * total;
            int i = 0;
            double cw = red[0][0];
            while (cw < tthr && i < K - 1) { ++i; cw += red[i][0]; }
            s_choice = i;
            if (choice) choice[b] = i;
        }
        __syncthreads();
        if (apply) {
            const int k = s_choice;
            const double nrm = s_norm2[k] > 0.0 ? 1.0 / sqrt(s_norm2[k]) : 0.0;
            if (SMEM) {
                for (int r = threadIdx.x; r < n; r += blockDim.x) {
                    const c128 y = term_row(tl.t[k], x, r, n);
                    pb[r] = make_double2(y.x * nrm, y.y * nrm);
                }
            } else {
                // ψ is read from global while being overwritten: stage the result in a second pass is not
                // possible without scratch, so the non-smem path requires diag-only operators (checked on host)
                for (int r = threadIdx.x; r < n; r += blockDim.x) {
                    const c128 y = term_row(tl.t[k], x, r, n);
                    pb[r] = make_double2(y.x * nrm, y.y * nrm);
                }
            }
        }
    }
}

// resample (src/opensys/trajectory_jump.jl:81-88) needs ‖L_k ψ‖ of the operator each Liouvillian proposed, and the winner is
// applied afterwards: norm[b] = ‖L_{choice[b]} ψ_b‖ and, when apply != 0, ψ_b ← L_kψ_b/‖L_kψ_b‖ (members with mask 0 or a
// negative choice are skipped).  ψ staged in shared memory, fixed-order reduction.
__global__ void __launch_bounds__(512) traj_apply_choice_kernel(int n, int nb, TermList tl, c128* __restrict__ psi,
                                                                const int32_t* __restrict__ choice,
                                                                const unsigned char* __restrict__ mask, double* __restrict__ norm,
                                                                int apply) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c128* sp = reinterpret_cast<c128*>(smem_raw);
    __shared__ double red[32];
    __shared__ double s_n2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        const int k = choice[b];
        if ((mask && !mask[b]) || k < 0 || k >= tl.count) continue;
        c128* pb = psi + (size_t)b * n;
        __syncthreads();
        for (int r = threadIdx.x; r < n; r += blockDim.x) sp[r] = pb[r];
        __syncthreads();
        double s = 0.0;
        for (int r = threadIdx.x; r < n; r += blockDim.x) {
            const c128 y = term_row(tl.t[k], sp, r, n);
            s = fma(y.x, y.x, s);
            s = fma(y.y, y.y, s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[warp] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0.0;
            for (int i = 0; i < nwarp; ++i) a += red[i];
            s_n2 = a;
            if (norm) norm[b] = sqrt(a);
        }
        __syncthreads();
        if (apply) {
            const double nrm = s_n2 > 0.0 ? 1.0 / sqrt(s_n2) : 0.0;
            for (int r = threadIdx.x; r < n; r += blockDim.x) {
                const c128 y = term_row(tl.t[k], sp, r, n);
                pb[r] = make_double2(y.x * nrm, y.y * nrm);
            }
        }
    }
}

// pick[b] = StatsBase.sample(1:M, Weights(w[:, b])) with the supplied uniform: t = u·Σw; scan left to right
__global__ void resample_kernel(int nb, int M, const double* __restrict__ w, const double* __restrict__ uniform,
                                const unsigned char* __restrict__ mask, int32_t* __restrict__ pick) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    if (mask && !mask[b]) { pick[b] = -1; return; }
    double total = 0.0;
    for (int m = 0; m < M; ++m) total += w[(size_t)m * nb + b];
    const double t = uniform[b] * total;
    int i = 0;
    double cw = w[b];
    while (cw < t && i < M - 1) { ++i; cw += w[(size_t)i * nb + b]; }
    pick[b] = i;
}

// K8 for diagonal jump operators (e.g. Z_k dephasing): no gathers, so ψ never touches shared memory — every
// thread keeps its R rows in registers, the K norms are reduced by a fixed-order shuffle tree, and the
// selected operator is applied from registers.  DRAM traffic: 16n read (+16n written when applied).
struct DiagJumpDesc {
    int K, all_const;
    const double* absq[kMaxTerms];
    const c128* d[kMaxTerms];
    const double* dreal[kMaxTerms];
    double cabs[kMaxTerms];
};

template <int R, int NT>
__global__ void __launch_bounds__(NT) traj_jump_diag_kernel(int n, int nb, DiagJumpDesc dj, const double* __restrict__ gamma,
                                                            long long gamma_ld, c128* __restrict__ psi,
                                                            const double* __restrict__ uniform,
                                                            const unsigned char* __restrict__ mask, int32_t* __restrict__ choice,
                                                            double* __restrict__ prob, int apply) {
    __shared__ double red[kMaxTerms][NT / 32];
    __shared__ double s_norm2;
    __shared__ int s_choice;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = dj.K;
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        if (mask && !mask[b]) {
            if (tid == 0 && choice) choice[b] = -1;
            continue;
        }
        c128* pb = psi + (size_t)b * n;
        c128 x[R];
        double a2[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int r = tid + k * NT;
            x[k] = r < n ? pb[r] : make_double2(0.0, 0.0);
            a2[k] = fma(x[k].x, x[k].x, x[k].y * x[k].y);
        }
        __syncthreads();  // red / s_* of the previous trajectory fully consumed
        const int nacc = dj.all_const ? 1 : K;
        for (int k = 0; k < nacc; ++k) {
            double s = 0.0;
            if (dj.all_const) {
#pragma unroll
                for (int q = 0; q < R; ++q) s += a2[q];
            } else {
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const int r = tid + q * NT;
                    if (r < n) s = fma(__ldg(dj.absq[k] + r), a2[q], s);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) red[k][warp] = s;
        }
        __syncthreads();
        if (tid == 0) {
            const double* g = gamma + b * gamma_ld;
            double nrm[kMaxTerms];
            for (int k = 0; k < nacc; ++k) {
                double a = 0.0;
                for (int i = 0; i < NT / 32; ++i) a += red[k][i];
                nrm[k] = a;
            }
            double total = 0.0, pk[kMaxTerms];
            for (int k = 0; k < K; ++k) {
                const double n2 = dj.all_const ? dj.cabs[k] * nrm[0] : nrm[k];  // ‖L_kψ‖²
                nrm[k == 0 ? 0 : k] = dj.all_const ? nrm[0] : nrm[k];
                pk[k] = g[k] * n2;  // γ_k · norm(L_k u)^2   (trajectory_jump.jl:8)
                if (prob) prob[(size_t)b * K + k] = pk[k];
                total += pk[k];
            }
            const double tthr = uniform[b] * total;  // StatsBase.sample(rng, wv)
            int i = 0;
            double cw = pk[0];
            while (cw < tthr && i < K - 1) { ++i; cw += pk[i]; }
            s_choice = i;
            s_norm2 = dj.all_const ? dj.cabs[i] * nrm[0] : nrm[i];
            if (choice) choice[b] = i;
        }
        if (apply) {
            __syncthreads();
            const int kc = s_choice;
            const double inv = s_norm2 > 0.0 ? 1.0 / sqrt(s_norm2) : 0.0;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int r = tid + k * NT;
                if (r < n) {
                    c128 dv;
                    if (dj.dreal[kc]) dv = make_double2(__ldg(dj.dreal[kc] + r), 0.0);
                    else dv = __ldg(dj.d[kc] + r);
                    const double yr = dv.x * x[k].x - dv.y * x[k].y, yi = dv.x * x[k].y + dv.y * x[k].x;
                    pb[r] = make_double2(yr * inv, yi * inv);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// sparse commutator on density matrices: du = −i(H u − u H), H given by values on the union pattern
// ------------------------------------------------------------------------------------------
__global__ void sparse_commutator_kernel(int n, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ rcol,
                                         const int32_t* __restrict__ rperm, const int32_t* __restrict__ colptr,
                                         const int32_t* __restrict__ crow, const c128* __restrict__ hvals,
                                         long long hv_stride, const c128* __restrict__ u, c128* __restrict__ du) {
    const int b = blockIdx.z;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int cidx = blockIdx.y;
    if (r >= n) return;
    const c128* hv = hvals + b * hv_stride;
    const c128* ub = u + (size_t)b * n * n;
    double sr = 0.0, si = 0.0;
    for (int p = rowptr[r]; p < rowptr[r + 1]; ++p) {  // (H u)[r,c] = Σ_j H[r,j] u[j,c]
        const c128 h = hv[rperm[p]];
        const c128 x = ub[rcol[p] + (size_t)cidx * n];
        sr += h.x * x.x - h.y * x.y;
        si += h.x * x.y + h.y * x.x;
    }
    for (int p = colptr[cidx]; p < colptr[cidx + 1]; ++p) {  // (u H)[r,c] = Σ_j u[r,j] H[j,c]
        const c128 h = hv[p];
        const c128 x = ub[r + (size_t)crow[p] * n];
        sr -= h.x * x.x - h.y * x.y;
        si -= h.x * x.y + h.y * x.x;
    }
    du[(size_t)b * n * n + r + (size_t)cidx * n] = make_double2(si, -sr);  // −i (sr + i si)
}

// dense[b][row + col*n] = vals[b][p] for the CSC entry p (column found by binary search in colptr)
__global__ void csc_densify_kernel(int n, long long nnz, const int32_t* __restrict__ colptr, const int32_t* __restrict__ rowidx,
                                   const c128* __restrict__ vals, c128* __restrict__ dense) {
    const int b = blockIdx.y;
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    int lo = 0, hi = n;  // largest col with colptr[col] <= p
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (colptr[mid] <= p) lo = mid; else hi = mid;
    }
    dense[(size_t)b * n * n + rowidx[p] + (size_t)lo * n] = vals[(size_t)b * nnz + p];
}

int term_list(const std::vector<const TermDev*>& ts, TermList& tl, oqb_ctx* c) {
    if ((int)ts.size() > kMaxTerms) return set_error(c, OQB_EINVAL, "too many operator terms (%zu > %d)", ts.size(), kMaxTerms);
    tl.count = (int)ts.size();
    for (int i = 0; i < tl.count; ++i) {
        tl.t[i].vals = ts[i]->vals;
        tl.t[i].cols = ts[i]->cols;
        tl.t[i].is_diag = ts[i]->is_diag;
        tl.t[i].width = ts[i]->width;
    }
    return 0;
}

int sm_count(oqb_ctx* c) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    return sms;
}

// rk != nullptr: fuse the RK4 stage arithmetic into the store (XOR-structured operators only); *fused reports whether
// that happened — when it did not, NOTHING was computed and the caller runs the unfused sequence.
int traj_matvec_impl(oqb_sparse* sp, int nb, const double* coefH, int coef_shared, const double* gamma, int gamma_shared,
                     const c128* psi, c128* dpsi, const RkEpilogue* rk = nullptr, int* fused = nullptr) {
    if (fused) *fused = 0;
    oqb_ctx* c = sp->ctx;
    if (nb <= 0) return 0;
    if (!coefH || !psi || !dpsi) return set_error(c, OQB_EINVAL, "oqb_traj_matvec: null argument");
    const int n = sp->n, nt = sp->nterms, K = gamma ? sp->K : 0;
    std::vector<const TermDev*> ts;
    for (int i = 0; i < nt; ++i) ts.push_back(&sp->h_terms[i]);
    const bool comb = K > 0 && gamma_shared;
    if (comb) {
        sp->ldl_comb.diag_const = sp->ldl_all_const;
        sp->ldl_comb.const_coef_one = true;
        double cv = 0.0;
        if (sp->ldl_all_const)
            for (int k = 0; k < K; ++k) cv = fma(-0.5 * gamma[k], sp->ldl_const[k], cv);
        sp->ldl_comb.const_val = make_double2(cv, 0.0);
        ts.push_back(&sp->ldl_comb);
    }
    else for (int k = 0; k < K; ++k) ts.push_back(&sp->ldl_terms[k]);
    TermList tl;
    OQB_TRY(term_list(ts, tl, c));
    const int ncoef = tl.count;
    const bool shared = coef_shared && (K == 0 || gamma_shared);
    const int rows = shared ? 1 : nb;
    // stage: complex coefficient rows, then (comb) the K shared γ
    std::vector<char> stage((size_t)rows * ncoef * sizeof(c128) + (comb ? K * sizeof(double) : 0));
    c128* cf = (c128*)stage.data();
    for (int b = 0; b < rows; ++b) {
        c128* r = cf + (size_t)b * ncoef;
        for (int i = 0; i < nt; ++i) r[i] = make_double2(0.0, -coefH[(coef_shared ? 0 : (size_t)b * nt) + i]);  // −i f_i(s)
        if (comb) r[nt] = make_double2(1.0, 0.0);
        else for (int k = 0; k < K; ++k) r[nt + k] = make_double2(-0.5 * gamma[(size_t)b * sp->K + k], 0.0);
    }
    if (comb) memcpy(stage.data() + (size_t)rows * ncoef * sizeof(c128), gamma, K * sizeof(double));
    void* d_stage = nullptr;
    OQB_TRY(coef_upload(c, stage.data(), stage.size(), &d_stage));
    if (comb) {
        const double* dg = (const double*)((char*)d_stage + (size_t)rows * ncoef * sizeof(c128));
        ldl_combine_kernel<<<(unsigned)((sp->comb_len + 255) / 256), 256, 0, c->stream>>>(K, sp->comb_len, dg, sp->d_ldl_src,
                                                                                       sp->ldl_comb.vals,
                                                                                       sp->ldl_comb.is_diag ? sp->ldl_comb.rvals : nullptr);
        OQB_LAUNCH_CHECK(c);
    }
    {
        bool used = false;  // implicit-column (XOR-structured) operator, 3-deep ψ ring, no CTA barrier
        OQB_TRY(traj_matvec_xor(c, n, nb, ts, (const c128*)d_stage, shared ? 0 : ncoef, psi, dpsi, &used, rk));
        if (used) {
            if (fused) *fused = 1;
            return 0;
        }
    }
    if (c->opt.traj_ell) {
        bool used = false;  // unstructured operator: sliced ELL with conflict-free slot order, ψ ring, several ψ per operator read
        OQB_TRY(traj_matvec_sell(c, n, nb, ts, (const c128*)d_stage, shared ? 0 : ncoef, psi, dpsi, &used, &sp->sell, rk));
        if (used) {
            if (fused) *fused = rk ? 1 : 0;
            return 0;
        }
    }
    if (rk) return 0;
    {
        bool used = false;  // register-resident operator + bulk-async ψ staging when the term list allows it
        OQB_TRY(traj_matvec_fast(c, n, nb, ts, (const c128*)d_stage, shared ? 0 : ncoef, psi, dpsi, &used));
        if (used) return 0;
    }
    const size_t smem = (size_t)n * sizeof(c128);
    const int threads = n >= 2048 ? 512 : (n >= 256 ? 256 : 64);
    const int sms = sm_count(c);
    if (smem <= 200 * 1024) {
        static DevOnce cfg;
        if (!cfg.done(c->device)) { OQB_CUDA(c, cudaFuncSetAttribute(traj_matvec_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); cfg.set(c->device); }
        int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / (smem + 1024)));
        int grid = std::min(nb, sms * per_sm);
        ProfScope prof(c, 32.0 * n * (double)nb);
        traj_matvec_kernel<true><<<grid, threads, smem, c->stream>>>(n, nb, tl, (const c128*)d_stage, shared ? 0 : ncoef, psi, dpsi);
    } else {
        int grid = std::min(nb, sms * 4);
        traj_matvec_kernel<false><<<grid, threads, 0, c->stream>>>(n, nb, tl, (const c128*)d_stage, shared ? 0 : ncoef, psi, dpsi);
    }
    OQB_LAUNCH_CHECK(c);
    return 0;
}

}  // namespace

int oqb_traj_matvec_rk(oqb_sparse* sp, int nb, const double* h_coefH, int coef_shared, const double* h_gamma, int gamma_shared,
                       const c128* psi_in, c128* out, const RkEpilogue* rk, int* fused) {
    OQB_DEVICE(sp, sp->ctx);
    return traj_matvec_impl(sp, nb, h_coefH, coef_shared, h_gamma, gamma_shared, psi_in, out, rk, fused);
}

// accessors for traj_evolve.cu (the struct layout stays private to this file)
oqb_ctx* oqb_sparse_ctx(oqb_sparse* sp) { return sp->ctx; }
int oqb_sparse_dims(oqb_sparse* sp, int* n, int* nterms, int* K) {
    OQB_DEVICE(sp, sp->ctx);
    *n = sp->n; *nterms = sp->nterms; *K = sp->K;
    return 0;
}

extern "C" {

int oqb_sparse_destroy(oqb_sparse* sp) {
    OQB_DEVICE(sp, sp->ctx);
    if (!sp) return OQB_OK;
    cudaStreamSynchronize(sp->ctx->stream);
    for (auto& t : sp->h_terms) free_term(t);
    for (auto& t : sp->ldl_terms) free_term(t);
    for (auto& t : sp->l_terms) free_term(t);
    free_term(sp->ldl_comb);
    sell_cache_free(sp->sell);
    sp->sell = nullptr;
    if (sp->d_ldl_src) cudaFree(sp->d_ldl_src);
    if (sp->d_termvals) cudaFree(sp->d_termvals);
    if (sp->d_u_colptr) cudaFree(sp->d_u_colptr);
    if (sp->d_u_row) cudaFree(sp->d_u_row);
    if (sp->d_r_rowptr) cudaFree(sp->d_r_rowptr);
    if (sp->d_r_col) cudaFree(sp->d_r_col);
    if (sp->d_r_perm) cudaFree(sp->d_r_perm);
    delete sp;
    return OQB_OK;
}

int oqb_sparse_create(oqb_ctx* c, int n, int nterms, const int64_t* h_colptr, const int64_t* h_rowval, const void* h_nzval,
                      const int64_t* h_nnz_off, int K, const int64_t* h_l_colptr, const int64_t* h_l_rowval,
                      const void* h_l_nzval, const int64_t* h_l_nnz_off, oqb_sparse** out) {
    OQB_DEVICE(c, c);
    if (!c || !out) return OQB_EINVAL;
    *out = nullptr;
    if (n <= 0 || nterms < 0 || K < 0 || nterms + K > kMaxTerms - 1 || (nterms > 0 && (!h_colptr || !h_rowval || !h_nzval || !h_nnz_off)) ||
        (K > 0 && (!h_l_colptr || !h_l_rowval || !h_l_nzval || !h_l_nnz_off)))
        return set_error(c, OQB_EINVAL, "oqb_sparse_create: bad arguments");
    OQB_CUDA(c, cudaSetDevice(c->device));
    oqb_sparse* sp = new oqb_sparse();
    sp->ctx = c; sp->n = n; sp->nterms = nterms; sp->K = K;
    std::vector<HostCsc> hm(nterms), lm(K), ldl(K);
    for (int i = 0; i < nterms; ++i)
        hm[i] = parse_csc(n, h_colptr + (size_t)i * (n + 1), h_rowval + h_nnz_off[i], (const c128*)h_nzval + h_nnz_off[i]);
    for (int k = 0; k < K; ++k) {
        lm[k] = parse_csc(n, h_l_colptr + (size_t)k * (n + 1), h_l_rowval + h_l_nnz_off[k], (const c128*)h_l_nzval + h_l_nnz_off[k]);
        ldl[k] = ldagl(lm[k]);
    }
    // union pattern of the cache (sparse_hamiltonian.jl:33-34 + the pattern `cache .-= ½γL'L` creates)
    std::vector<const HostCsc*> all;
    for (auto& m : hm) all.push_back(&m);
    for (auto& m : ldl) all.push_back(&m);
    HostCsc U = union_pattern(all, n);
    sp->nnzU = (int64_t)U.row.size();
    sp->u_colptr = U.colptr;
    sp->u_row = U.row;
    int s = 0;
    auto fail = [&](int code) { oqb_sparse_destroy(sp); return code; };
    {
        std::vector<c128> tv((size_t)(nterms + K) * std::max<int64_t>(sp->nnzU, 1));
        for (int i = 0; i < nterms + K; ++i) {
            std::vector<c128> e = expand_on(U, i < nterms ? hm[i] : ldl[i - nterms]);
            std::copy(e.begin(), e.end(), tv.begin() + (size_t)i * sp->nnzU);
        }
        if (cudaMalloc((void**)&sp->d_termvals, tv.size() * sizeof(c128) + 16) != cudaSuccess) return fail(set_error(c, OQB_ENOMEM, "oqb_sparse_create: out of memory"));
        cudaMemcpy(sp->d_termvals, tv.data(), tv.size() * sizeof(c128), cudaMemcpyHostToDevice);
        // CSC (int32) and CSR view with permutation
        std::vector<int32_t> cp(n + 1), rp(n + 1, 0), rc(sp->nnzU), perm(sp->nnzU);
        for (int j = 0; j <= n; ++j) cp[j] = (int32_t)U.colptr[j];
        for (int64_t p = 0; p < sp->nnzU; ++p) rp[U.row[p] + 1]++;
        for (int r = 0; r < n; ++r) rp[r + 1] += rp[r];
        std::vector<int32_t> fill(rp.begin(), rp.end() - 1);
        for (int j = 0; j < n; ++j)
            for (int64_t p = U.colptr[j]; p < U.colptr[j + 1]; ++p) {
                int q = fill[U.row[p]]++;
                rc[q] = j;
                perm[q] = (int32_t)p;
            }
        auto up = [&](int32_t** d, const std::vector<int32_t>& h) {
            if (cudaMalloc((void**)d, h.size() * sizeof(int32_t) + 16) != cudaSuccess) return 1;
            cudaMemcpy(*d, h.data(), h.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
            return 0;
        };
        if (up(&sp->d_u_colptr, cp) || up(&sp->d_u_row, U.row) || up(&sp->d_r_rowptr, rp) || up(&sp->d_r_col, rc) || up(&sp->d_r_perm, perm))
            return fail(set_error(c, OQB_ENOMEM, "oqb_sparse_create: out of memory"));
    }
    sp->h_terms.resize(nterms);
    sp->ldl_terms.resize(K);
    sp->l_terms.resize(K);
    for (int i = 0; i < nterms && !s; ++i) s = make_term(c, hm[i], sp->h_terms[i]);
    for (int k = 0; k < K && !s; ++k) s = make_term(c, ldl[k], sp->ldl_terms[k]);
    for (int k = 0; k < K && !s; ++k) s = make_term(c, lm[k], sp->l_terms[k]);
    if (!s && K > 0) {
        sp->ldl_all_const = true;
        sp->ldl_const.assign(K, 0.0);
        for (int k = 0; k < K && sp->ldl_all_const; ++k) {
            const HostCsc& m = ldl[k];
            if ((int64_t)m.row.size() != n) { sp->ldl_all_const = false; break; }
            for (int j = 0; j < n; ++j) {
                const int64_t p0 = m.colptr[j];
                if (m.colptr[j + 1] - p0 != 1 || m.row[p0] != j || m.val[p0].y != 0.0 || m.val[p0].x != m.val[0].x) {
                    sp->ldl_all_const = false;
                    break;
                }
            }
            sp->ldl_const[k] = m.val.empty() ? 0.0 : m.val[0].x;
        }
        // shared-γ fast path: every L_k'L_k expanded on their common pattern so that −½Σγ_k L_k'L_k is one term
        std::vector<const HostCsc*> ls;
        for (auto& m : ldl) ls.push_back(&m);
        HostCsc LU = union_pattern(ls, n);
        for (size_t p = 0; p < LU.val.size(); ++p) LU.val[p] = make_double2(1.0, 0.0);  // structure only
        s = make_term(c, LU, sp->ldl_comb);
        sp->ldl_comb.dynamic_vals = true;
        sp->ldl_comb.ell_real = sp->ldl_comb.ell_imag = false;
        if (!s) {
            const int width = sp->ldl_comb.width;
            sp->comb_len = sp->ldl_comb.is_diag ? n : (long long)width * n;
            std::vector<c128> src((size_t)K * sp->comb_len, make_double2(0.0, 0.0));
            for (int k = 0; k < K; ++k) {
                // place L_k'L_k values at the slots make_term assigned for LU (same fill order: by column, then row)
                std::vector<c128> e = expand_on(LU, ldl[k]);
                if (sp->ldl_comb.is_diag) {
                    for (int j = 0; j < n; ++j)
                        for (int64_t p = LU.colptr[j]; p < LU.colptr[j + 1]; ++p) src[(size_t)k * sp->comb_len + j] = e[p];
                } else {
                    std::vector<int> fill(n, 0);
                    for (int j = 0; j < n; ++j)
                        for (int64_t p = LU.colptr[j]; p < LU.colptr[j + 1]; ++p) {
                            const int r = LU.row[p];
                            const int slot = fill[r]++;
                            src[(size_t)k * sp->comb_len + (size_t)slot * n + r] = e[p];
                        }
                }
            }
            if (cudaMalloc((void**)&sp->d_ldl_src, src.size() * sizeof(c128) + 16) != cudaSuccess) return fail(set_error(c, OQB_ENOMEM, "oqb_sparse_create: out of memory"));
            cudaMemcpy(sp->d_ldl_src, src.data(), src.size() * sizeof(c128), cudaMemcpyHostToDevice);
        }
    }
    if (s) return fail(s);
    *out = sp;
    return OQB_OK;
}

int oqb_sell_plan_probe(int n, int nterms, const int64_t* const* colptr, const int32_t* const* rowval, int unit, double* padding,
                        int64_t* phases, int64_t* conflicting_phases, int* identity, int* misplaced) {
    if (!colptr || !rowval || !padding || !phases || !conflicting_phases || !identity || !misplaced) return OQB_EINVAL;
    long long ph = 0, cf = 0;
    const int s = sell_plan_probe(n, nterms, colptr, rowval, unit, padding, &ph, &cf, identity, misplaced);
    *phases = ph;
    *conflicting_phases = cf;
    return s;
}

int oqb_traj_sell_stats(oqb_sparse* sp, int* plans, double* padding) {
    if (!sp || !plans || !padding) return OQB_EINVAL;
    sell_cache_stats(sp->sell, plans, padding);
    return 0;
}

int oqb_sparse_pattern_nnz(oqb_sparse* sp, int64_t* nnz) { *nnz = sp->nnzU; return 0; }
int oqb_sparse_pattern(oqb_sparse* sp, int64_t* h_colptr, int64_t* h_rowval) {
    OQB_DEVICE(sp, sp->ctx);
    for (int j = 0; j <= sp->n; ++j) h_colptr[j] = sp->u_colptr[j] + 1;
    for (int64_t p = 0; p < sp->nnzU; ++p) h_rowval[p] = sp->u_row[p] + 1;
    return 0;
}

int oqb_sparse_update_cache(oqb_sparse* sp, int nb, const double* h_coefH, int coef_shared, const double* h_gamma,
                            int gamma_shared, void* d_nzval) {
    OQB_DEVICE(sp, sp->ctx);
    oqb_ctx* c = sp->ctx;
    if (nb <= 0 || sp->nnzU == 0) return 0;
    if (!h_coefH || !d_nzval) return set_error(c, OQB_EINVAL, "oqb_sparse_update_cache: null argument");
    const int nt = sp->nterms, K = h_gamma ? sp->K : 0, nmat = nt + K;
    const bool shared = coef_shared && (K == 0 || gamma_shared);
    const int rows = shared ? 1 : nb;
    std::vector<c128> coef((size_t)rows * nmat);
    for (int b = 0; b < rows; ++b) {
        for (int i = 0; i < nt; ++i) coef[(size_t)b * nmat + i] = make_double2(0.0, -h_coefH[(coef_shared ? 0 : (size_t)b * nt) + i]);
        for (int k = 0; k < K; ++k) coef[(size_t)b * nmat + nt + k] = make_double2(-0.5 * h_gamma[(gamma_shared ? 0 : (size_t)b * sp->K) + k], 0.0);
    }
    void* d_coef = nullptr;
    OQB_TRY(coef_upload(c, coef.data(), coef.size() * sizeof(c128), &d_coef));
    return lincomb(c, nmat, sp->nnzU, sp->d_termvals, (const c128*)d_coef, shared ? 0 : nmat, (c128*)d_nzval, nb, false);
}

int oqb_sparse_update_vectorized_cache(oqb_sparse* sp, int nb, const double* h_coefH, int coef_shared, void* d_cache) {
    OQB_DEVICE(sp, sp->ctx);
    oqb_ctx* c = sp->ctx;
    if (nb <= 0) return 0;
    if (!h_coefH || !d_cache) return set_error(c, OQB_EINVAL, "oqb_sparse_update_vectorized_cache: null argument");
    const int n = sp->n, nt = sp->nterms;
    const int rows = coef_shared ? 1 : nb;
    const long long nn = (long long)n * n;
    std::vector<c128> coef((size_t)rows * nt);
    for (size_t i = 0; i < coef.size(); ++i) coef[i] = make_double2(h_coefH[i], 0.0);  // H(s) = Σ f_i m_i
    void* d_coef = nullptr;
    OQB_TRY(coef_upload(c, coef.data(), coef.size() * sizeof(c128), &d_coef));
    const size_t nnzb = (size_t)std::max<int64_t>(sp->nnzU, 1);
    OQB_TRY(ws_reserve(c, (size_t)rows * (nnzb + nn) * sizeof(c128)));
    c128* vals = (c128*)c->ws;
    c128* dense = vals + (size_t)rows * nnzb;
    OQB_TRY(lincomb(c, nt, sp->nnzU, sp->d_termvals, (const c128*)d_coef, coef_shared ? 0 : nt, vals, rows, false));
    OQB_CUDA(c, cudaMemsetAsync(dense, 0, (size_t)rows * nn * sizeof(c128), c->stream));
    if (sp->nnzU > 0) {
        dim3 grid((unsigned)((sp->nnzU + 255) / 256), rows);
        csc_densify_kernel<<<grid, 256, 0, c->stream>>>(n, sp->nnzU, sp->d_u_colptr, sp->d_u_row, vals, dense);
        OQB_LAUNCH_CHECK(c);
    }
    return vectorized_commutator(c, n, dense, coef_shared ? 0 : nn, (c128*)d_cache, nb);
}

int oqb_sparse_commutator(oqb_sparse* sp, int nb, const double* h_coefH, int coef_shared, const void* d_u, void* d_du) {
    OQB_DEVICE(sp, sp->ctx);
    oqb_ctx* c = sp->ctx;
    if (nb <= 0) return 0;
    if (!h_coefH || !d_u || !d_du) return set_error(c, OQB_EINVAL, "oqb_sparse_commutator: null argument");
    const int n = sp->n, nt = sp->nterms;
    const int rows = coef_shared ? 1 : nb;
    std::vector<c128> coef((size_t)rows * nt);
    for (size_t i = 0; i < coef.size(); ++i) coef[i] = make_double2(h_coefH[i], 0.0);  // H(s) = Σ f_i m_i
    void* d_coef = nullptr;
    OQB_TRY(coef_upload(c, coef.data(), coef.size() * sizeof(c128), &d_coef));
    OQB_TRY(ws_reserve(c, (size_t)rows * std::max<int64_t>(sp->nnzU, 1) * sizeof(c128)));
    OQB_TRY(lincomb(c, nt, sp->nnzU, sp->d_termvals, (const c128*)d_coef, coef_shared ? 0 : nt, (c128*)c->ws, rows, false));
    for (int b0 = 0; b0 < nb; b0 += 65535) {
        const int cnt = nb - b0 < 65535 ? nb - b0 : 65535;
        dim3 grid((n + 127) / 128, n, cnt);
        sparse_commutator_kernel<<<grid, 128, 0, c->stream>>>(n, sp->d_r_rowptr, sp->d_r_col, sp->d_r_perm, sp->d_u_colptr,
                                                              sp->d_u_row, (const c128*)c->ws + (coef_shared ? 0 : (size_t)b0 * sp->nnzU),
                                                              coef_shared ? 0 : sp->nnzU, (const c128*)d_u + (size_t)b0 * n * n,
                                                              (c128*)d_du + (size_t)b0 * n * n);
        OQB_LAUNCH_CHECK(c);
    }
    return 0;
}

int oqb_traj_matvec(oqb_sparse* sp, int nb, const double* h_coefH, int coef_shared, const double* h_gamma, int gamma_shared,
                    const void* d_psi, void* d_dpsi) {
    OQB_DEVICE(sp, sp->ctx);
    return traj_matvec_impl(sp, nb, h_coefH, coef_shared, h_gamma, gamma_shared, (const c128*)d_psi, (c128*)d_dpsi);
}

int oqb_traj_matvec_host(oqb_sparse* sp, int nb, const double* h_coefH, int coef_shared, const double* h_gamma,
                         int gamma_shared, const void* h_psi, void* h_dpsi) {
    OQB_DEVICE(sp, sp->ctx);
    oqb_ctx* c = sp->ctx;
    const size_t bytes = (size_t)sp->n * sizeof(c128);
    int chunk = (int)(((size_t)256 << 20) / bytes);
    if (chunk < 1) chunk = 1;
    const int nt = sp->nterms, K = sp->K;
    return host_pipeline(c, nb, bytes, bytes, h_psi, h_dpsi, chunk, [&](int b0, int cnt, const void* din, void* dout) {
        const double* cf = coef_shared ? h_coefH : h_coefH + (size_t)b0 * nt;
        const double* gm = (!h_gamma || gamma_shared) ? h_gamma : h_gamma + (size_t)b0 * K;
        return traj_matvec_impl(sp, cnt, cf, coef_shared, gm, gamma_shared, (const c128*)din, (c128*)dout);
    });
}

int oqb_traj_norm2(oqb_sparse* sp, int nb, const void* d_psi, double* d_out) {
    OQB_DEVICE(sp, sp->ctx);
    oqb_ctx* c = sp->ctx;
    if (nb <= 0) return 0;
    norm2_kernel<<<nb, 256, 0, c->stream>>>(sp->n, (const c128*)d_psi, d_out);
    OQB_LAUNCH_CHECK(c);
    return 0;
}

int oqb_traj_jump(oqb_sparse* sp, int nb, const double* h_gamma, int gamma_shared, void* d_psi, const double* d_uniform,
                  const uint8_t* d_mask, int32_t* d_choice, double* d_prob, int apply) {
    OQB_DEVICE(sp, sp->ctx);
    oqb_ctx* c = sp->ctx;
    if (nb <= 0) return 0;
    if (sp->K <= 0) return set_error(c, OQB_EINVAL, "oqb_traj_jump: no Lindblad operators");
    if (!h_gamma || !d_psi || !d_uniform) return set_error(c, OQB_EINVAL, "oqb_traj_jump: null argument");
    const int n = sp->n, K = sp->K;
    std::vector<const TermDev*> ts;
    for (int k = 0; k < K; ++k) ts.push_back(&sp->l_terms[k]);
    TermList tl;
    OQB_TRY(term_list(ts, tl, c));
    void* d_gamma = nullptr;
    OQB_TRY(coef_upload(c, h_gamma, sizeof(double) * (gamma_shared ? K : (size_t)nb * K), &d_gamma));
    {
        bool all_diag = true, all_const = true;
        for (int k = 0; k < K; ++k) {
            all_diag = all_diag && sp->l_terms[k].is_diag && sp->l_terms[k].absq;
            all_const = all_const && sp->l_terms[k].absq_const;
        }
        const bool no_fast = c->opt.traj_kernel == 1;
        if (all_diag && !no_fast && n <= 512 * 16) {
            DiagJumpDesc dj;
            dj.K = K;
            dj.all_const = all_const ? 1 : 0;
            for (int k = 0; k < K; ++k) {
                dj.absq[k] = sp->l_terms[k].absq;
                dj.d[k] = sp->l_terms[k].vals;
                dj.dreal[k] = sp->l_terms[k].is_real ? sp->l_terms[k].rvals : nullptr;
                dj.cabs[k] = sp->l_terms[k].absq_val;
            }
            const int sms2 = sm_count(c);
            const int grid = std::min(nb, sms2 * 16);
            ProfScope prof(c, (apply ? 32.0 : 16.0) * n * (double)nb);
#define OQB_DJ(RR, TT)                                                                                          \
    traj_jump_diag_kernel<RR, TT><<<grid, TT, 0, c->stream>>>(n, nb, dj, (const double*)d_gamma, gamma_shared ? 0 : K, \
                                                              (c128*)d_psi, d_uniform, d_mask, d_choice, d_prob, apply)
            if (n <= 256) OQB_DJ(2, 128);
            else if (n <= 1024) OQB_DJ(4, 256);
            else if (n <= 2048) OQB_DJ(8, 256);
            else if (n <= 4096) OQB_DJ(8, 512);
            else OQB_DJ(16, 512);
#undef OQB_DJ
            OQB_LAUNCH_CHECK(c);
            return 0;
        }
    }
    const size_t smem = (size_t)n * sizeof(c128);
    const int threads = n >= 2048 ? 512 : (n >= 256 ? 256 : 64);
    const int sms = sm_count(c);
    if (smem <= 200 * 1024) {
        static DevOnce cfg;
        if (!cfg.done(c->device)) { OQB_CUDA(c, cudaFuncSetAttribute(traj_jump_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); cfg.set(c->device); }
        int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / (smem + 8 * 1024)));
        int grid = std::min(nb, sms * per_sm);
        ProfScope prof(c, (apply ? 32.0 : 16.0) * n * (double)nb);
        traj_jump_kernel<true><<<grid, threads, smem, c->stream>>>(n, nb, tl, (const double*)d_gamma, gamma_shared ? 0 : K,
                                                                   (c128*)d_psi, d_uniform, d_mask, d_choice, d_prob, apply);
    } else {
        if (apply)
            for (int k = 0; k < K; ++k)
                if (!sp->l_terms[k].is_diag)
                    return set_error(c, OQB_EINVAL, "oqb_traj_jump: in-place apply for n=%d needs diagonal L_k (state exceeds shared memory)", n);
        int grid = std::min(nb, sms * 4);
        traj_jump_kernel<false><<<grid, threads, 0, c->stream>>>(n, nb, tl, (const double*)d_gamma, gamma_shared ? 0 : K,
                                                                 (c128*)d_psi, d_uniform, d_mask, d_choice, d_prob, apply);
    }
    OQB_LAUNCH_CHECK(c);
    return 0;
}

int oqb_traj_apply_choice(oqb_sparse* sp, int nb, void* d_psi, const int32_t* d_choice, const uint8_t* d_mask, double* d_norm,
                          int apply) {
    OQB_DEVICE(sp, sp->ctx);
    oqb_ctx* c = sp->ctx;
    if (nb <= 0) return 0;
    if (sp->K <= 0) return set_error(c, OQB_EINVAL, "oqb_traj_apply_choice: no Lindblad operators");
    if (!d_psi || !d_choice) return set_error(c, OQB_EINVAL, "oqb_traj_apply_choice: null argument");
    const int n = sp->n, K = sp->K;
    std::vector<const TermDev*> ts;
    for (int k = 0; k < K; ++k) ts.push_back(&sp->l_terms[k]);
    TermList tl;
    OQB_TRY(term_list(ts, tl, c));
    const size_t smem = (size_t)n * sizeof(c128);
    if (smem > 200 * 1024) return set_error(c, OQB_EINVAL, "oqb_traj_apply_choice: n = %d exceeds the shared-memory staging (12 800 levels)", n);
    static DevOnce cfg;
    if (!cfg.done(c->device)) { OQB_CUDA(c, cudaFuncSetAttribute(traj_apply_choice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); cfg.set(c->device); }
    const int threads = n >= 2048 ? 512 : (n >= 256 ? 256 : 64);
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / (smem + 8 * 1024)));
    const int grid = std::min(nb, sm_count(c) * per_sm);
    traj_apply_choice_kernel<<<grid, threads, smem, c->stream>>>(n, nb, tl, (c128*)d_psi, d_choice, d_mask, d_norm, apply);
    OQB_LAUNCH_CHECK(c);
    return 0;
}

int oqb_resample(oqb_ctx* c, int nb, int M, const double* d_weight, const double* d_uniform, const uint8_t* d_mask, int32_t* d_pick) {
    if (!c) return OQB_EINVAL;
    OQB_DEVICE(c, c);
    if (nb <= 0) return 0;
    if (M <= 0 || !d_weight || !d_uniform || !d_pick) return set_error(c, OQB_EINVAL, "oqb_resample: bad arguments");
    resample_kernel<<<(nb + 255) / 256, 256, 0, c->stream>>>(nb, M, d_weight, d_uniform, d_mask, d_pick);
    OQB_LAUNCH_CHECK(c);
    return 0;
}

}  // extern "C"
