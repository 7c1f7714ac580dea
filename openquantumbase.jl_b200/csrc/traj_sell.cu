// Ensemble matvec dψ_b = Σ_t c_bt M_t ψ_b for operators WITHOUT Pauli/XOR structure (the reference's
// SparseHamiltonian × vector, src/hamiltonian/sparse_hamiltonian.jl:70-75, plus the −½γL'L terms of
// src/opensys/lindblad.jl:103-109) — the path every operator that traj_xor.cu cannot describe takes.
//
// What bounds such a kernel on B200 (measured, oqb_probe_smem_gather): 13 gathers of 16 bytes per 32 bytes of HBM traffic
// put the shared-memory pipe, not HBM, on the critical path, and random columns make the 8 threads of an LDS.128 phase
// collide on bank groups (2.3× fewer gathers per second than a conflict-free pattern).  Three things are therefore fixed
// at plan time, on the host, once per term list:
//   * rows are sorted by length and cut into 32-row slices (SELL-32): a warp pays for its own longest row, not for the
//     longest row of the matrix; the permutation σ is applied when the result is stored;
//   * inside each 8-row group (one LDS.128 phase) the slots are an edge colouring of the bipartite graph rows × bank
//     groups (König: max-degree colours suffice), so the 8 gathers of a phase hit 8 different bank groups; rows are
//     picked into groups greedily so that the bank-group degrees stay close to the row degrees (little padding);
//   * column indices are 16-bit, values real when the term is real (or purely imaginary: factor i on the coefficient).
// At run time ψ is staged by cp.async.bulk into a ring of single-ψ slots (full/empty mbarriers, no CTA-wide barrier) and
// every operator entry fetched from L2 is applied to R consecutive trajectories.
#include <algorithm>
#include <cstring>
#include <map>
#include <numeric>

#include "../../include/oqb.h"
#include "sparse_terms.cuh"

namespace oqb {

namespace {

constexpr int kSellOff = 6;    // off-diagonal terms per list
constexpr int kSellDiag = 10;  // diagonal terms per list
constexpr uint16_t kPad = 0xFFFF;

struct SellOffDev {
    const uint32_t* off;   // [nslices + 1], in 32-entry columns
    const uint16_t* cols;  // [off[nslices] * 32]
    const c128* vals;      // complex values (is_real == 0)
    const double* rvals;   // real values (is_real == 1; purely imaginary terms store the imaginary parts)
    int is_real, x_imag, coef;
};
struct SellDiagDev {
    const c128* vals;     // indexed by the ORIGINAL row (read through σ)
    const double* rvals;  // real diagonal
    int is_const, coef;
    double cre, cim;  // constant diagonal value·I
};
struct SellDesc {
    int noff, ndiag, npos;
    const uint16_t* sigma;  // [npos]: original row of position p (kPad: none)
    SellOffDev o[kSellOff];
    SellDiagDev g[kSellDiag];
};

__device__ __forceinline__ unsigned s_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sb_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void sb_expect(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sb_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void sb_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "S_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra S_DONE;\n\t"
        "bra S_LOOP;\n\t"
        "S_DONE:\n\t"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void sb_bulk(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

constexpr int kMaxSlots = 12;
constexpr int kCoefPad = 16;   // a slot = ψ (n entries) + the trajectory's coefficient row (≤ 16 entries)
constexpr int kDiagAhead = 4;  // diagonal terms whose values are fetched before the off-diagonal loop of a row

// One off-diagonal term, one row, R trajectories: the operator entries of the row's slice column are fetched U at a time
// (one L2 round trip per batch instead of one per entry — with 16 warps per SM the kernel is otherwise latency-bound).
template <int R, int U, typename VT>
__device__ __forceinline__ void sell_row(const uint16_t* __restrict__ cp, const VT* __restrict__ vp, unsigned o0, unsigned o1,
                                         const c128* const (&x)[R], double (&ar)[R], double (&ai)[R]) {
    const unsigned lane = threadIdx.x & 31;
    const uint16_t* __restrict__ cq = cp + (size_t)o0 * 32;  // walking pointers: the U loads of a batch use immediate offsets
    const VT* __restrict__ vq = vp + (size_t)o0 * 32;
    for (int left = (int)(o1 - o0); left > 0; left -= U, cq += U * 32, vq += U * 32) {
        unsigned cj[U];
        VT v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            // past the slice: gather ψ[lane] (conflict-free) with a zero value — no branch anywhere in the batch
            cj[u] = lane;
            memset(&v[u], 0, sizeof(VT));
            if (u < left) {
                cj[u] = (unsigned)cq[u * 32];
                v[u] = vq[u * 32];
            }
        }
#pragma unroll
        for (int u0 = 0; u0 < U; u0 += 4) {
            if (u0 && u0 >= left) break;  // warp-uniform (a slice has one width): whole groups of 4 past the end are skipped
#pragma unroll
            for (int u = u0; u < u0 + 4; ++u) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const c128 xv = x[r][cj[u]];
                    if constexpr (sizeof(VT) == sizeof(double)) {
                        ar[r] = fma(v[u], xv.x, ar[r]);
                        ai[r] = fma(v[u], xv.y, ai[r]);
                    } else {
                        ar[r] = fma(v[u].x, xv.x, ar[r]);
                        ar[r] = fma(-v[u].y, xv.y, ar[r]);
                        ai[r] = fma(v[u].x, xv.y, ai[r]);
                        ai[r] = fma(v[u].y, xv.x, ai[r]);
                    }
                }
            }
        }
    }
}

// R trajectories per operator read, NT threads, U entries per fetch batch; position p = k·NT + tid, slice = p / 32.
// EPI: RK4 stage arithmetic fused into the store (RkEpilogue, sparse_terms.cuh) — dpsi may then alias psi: a trajectory
// is staged completely before any of its rows is written.
template <int R, int NT, int U, int EPI>
__global__ void __launch_bounds__(NT, 1)
    traj_sell_kernel(int n, int ns, int nb, const __grid_constant__ SellDesc d, const c128* __restrict__ coef, long long coef_ld,
                     int ncoef, const c128* psi, c128* dpsi, double wa, double wt, const c128* __restrict__ rk_psi0, c128* rk_acc,
                     double* __restrict__ norm_part) {
    constexpr int NW = NT / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    c128* ring = reinterpret_cast<c128*>(smem_raw);
    const int stage = n + kCoefPad;
    const int nsl = d.npos >> 5;
    uint16_t* s_sigma = reinterpret_cast<uint16_t*>(ring + (size_t)ns * stage);
    unsigned* s_off = reinterpret_cast<unsigned*>(s_sigma + ((d.npos + 7) & ~7));  // [noff][nsl + 1]
    __shared__ __align__(8) unsigned long long bars[2 * kMaxSlots];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned bar_base = s_addr(bars);
    const unsigned psi_bytes = (unsigned)n * (unsigned)sizeof(c128);
    const unsigned coef_bytes = coef_ld ? (unsigned)(ncoef * sizeof(c128)) : 0u;
    if (tid == 0) {
        for (int s = 0; s < ns; ++s) {
            sb_init(bar_base + 8 * s, 1);                // full
            sb_init(bar_base + 8 * (kMaxSlots + s), NW);  // empty
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int p = tid; p < d.npos; p += NT) s_sigma[p] = d.sigma[p];
    for (int t = 0; t < d.noff; ++t)
        for (int q = tid; q <= nsl; q += NT) s_off[t * (nsl + 1) + q] = d.o[t].off[q];
    if (!coef_ld && tid < ncoef)  // one coefficient row for everybody: park it in every slot
        for (int s = 0; s < ns; ++s) ring[(size_t)s * stage + n + tid] = coef[tid];
    __syncthreads();
    const long long stride = gridDim.x;
    // load sequence of this CTA: q = i·R + r  ↔  trajectory (blockIdx.x + i·stride)·R + r, slot q % ns
    int issued = 0;
    auto pump = [&](int upto) {
        while (issued < upto) {
            const long long b = (blockIdx.x + (long long)(issued / R) * stride) * R + issued % R;
            if (b >= nb) {
                issued = 0x3fffffff;
                break;
            }
            const int slot = issued % ns;
            sb_wait(bar_base + 8 * (kMaxSlots + slot), (unsigned)(((issued / ns) & 1) ^ 1));
            const unsigned full = bar_base + 8 * slot;
            sb_expect(full, psi_bytes + coef_bytes);
            sb_bulk(s_addr(ring + (size_t)slot * stage), psi + (size_t)b * n, psi_bytes, full);
            if (coef_bytes) sb_bulk(s_addr(ring + (size_t)slot * stage + n), coef + b * coef_ld, coef_bytes, full);
            ++issued;
        }
    };
    const long long ngroups = ((long long)nb + R - 1) / R;
    int i = 0;
    for (long long g = blockIdx.x; g < ngroups; g += stride, ++i) {
        if (tid == 0) pump(i * R + ns);  // slots whose previous occupants belong to iterations < i
        __syncwarp();
        const long long b0 = g * R;
        const c128* x[R];
        bool live[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int q = i * R + r;
            live[r] = b0 + r < nb;
            x[r] = ring + (size_t)((live[r] ? q : i * R) % ns) * stage;  // a dead member of the last group recomputes member 0 (not stored)
            if (live[r]) sb_wait(bar_base + 8 * (q % ns), (unsigned)((q / ns) & 1));
        }
        double nrm[R];
#pragma unroll
        for (int r = 0; r < R; ++r) nrm[r] = 0.0;
        for (int p = tid; p < d.npos; p += NT) {
            const int sl = p >> 5;
            const unsigned sg = s_sigma[p];
            // diagonal values first: their loads travel together with the first batch of operator entries
            c128 dv[kDiagAhead];
#pragma unroll
            for (int t = 0; t < kDiagAhead; ++t) {
                dv[t] = make_double2(0.0, 0.0);
                if (t < d.ndiag && sg != kPad) {
                    const SellDiagDev& gd = d.g[t];
                    if (gd.is_const) dv[t] = make_double2(gd.cre, gd.cim);
                    else if (gd.rvals) dv[t].x = gd.rvals[sg];
                    else dv[t] = gd.vals[sg];
                }
            }
            // fused RK4 stage: the extra operands are requested before the row's gathers so that their latency hides behind them
            c128 e0[EPI == 2 ? R : 1], e1[EPI >= 2 ? R : 1];
            if (EPI >= 2 && sg != kPad) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const size_t at = (size_t)(live[r] ? b0 + r : b0) * n + sg;
                    if (EPI == 2) e0[r] = rk_psi0[at];
                    e1[r] = rk_acc[at];
                }
            }
            double yr[R], yi[R];
#pragma unroll
            for (int r = 0; r < R; ++r) yr[r] = yi[r] = 0.0;
            for (int t = 0; t < d.noff; ++t) {
                const SellOffDev& o = d.o[t];
                const unsigned o0 = s_off[t * (nsl + 1) + sl], o1 = s_off[t * (nsl + 1) + sl + 1];
                double ar[R], ai[R];
#pragma unroll
                for (int r = 0; r < R; ++r) ar[r] = ai[r] = 0.0;
                if (o.is_real) sell_row<R, 2 * U, double>(o.cols + lane, o.rvals + lane, o0, o1, x, ar, ai);
                else sell_row<R, U, c128>(o.cols + lane, o.vals + lane, o0, o1, x, ar, ai);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    c128 cc = x[r][n + o.coef];
                    if (o.x_imag) cc = make_double2(-cc.y, cc.x);  // stored values are Im: the factor i goes onto the coefficient
                    yr[r] = fma(cc.x, ar[r], yr[r]);
                    yr[r] = fma(-cc.y, ai[r], yr[r]);
                    yi[r] = fma(cc.x, ai[r], yi[r]);
                    yi[r] = fma(cc.y, ar[r], yi[r]);
                }
            }
            if (sg != kPad) {
                c128 xo[R];
#pragma unroll
                for (int r = 0; r < R; ++r) xo[r] = x[r][sg];
                auto diag_acc = [&](const c128 v, int ci) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const c128 cc = x[r][n + ci];
                        const double zr = v.x * xo[r].x - v.y * xo[r].y, zi = v.x * xo[r].y + v.y * xo[r].x;
                        yr[r] = fma(cc.x, zr, yr[r]);
                        yr[r] = fma(-cc.y, zi, yr[r]);
                        yi[r] = fma(cc.x, zi, yi[r]);
                        yi[r] = fma(cc.y, zr, yi[r]);
                    }
                };
#pragma unroll
                for (int t = 0; t < kDiagAhead; ++t)
                    if (t < d.ndiag) diag_acc(dv[t], d.g[t].coef);
                for (int t = kDiagAhead; t < d.ndiag; ++t) {
                    const SellDiagDev& gd = d.g[t];
                    c128 v;
                    if (gd.is_const) v = make_double2(gd.cre, gd.cim);
                    else if (gd.rvals) v = make_double2(gd.rvals[sg], 0.0);
                    else v = gd.vals[sg];
                    diag_acc(v, gd.coef);
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (!live[r]) continue;
                    const size_t at = (size_t)(b0 + r) * n + sg;
                    if (EPI == 0) {
                        dpsi[at] = make_double2(yr[r], yi[r]);
                    } else if (EPI == 1) {
                        rk_acc[at] = make_double2(xo[r].x + wa * yr[r], xo[r].y + wa * yi[r]);
                        dpsi[at] = make_double2(xo[r].x + wt * yr[r], xo[r].y + wt * yi[r]);
                    } else if (EPI == 2) {
                        rk_acc[at] = make_double2(e1[r].x + wa * yr[r], e1[r].y + wa * yi[r]);
                        dpsi[at] = make_double2(e0[r].x + wt * yr[r], e0[r].y + wt * yi[r]);
                    } else {
                        const c128 v = make_double2(e1[r].x + wa * yr[r], e1[r].y + wa * yi[r]);
                        dpsi[at] = v;
                        nrm[r] = fma(v.x, v.x, nrm[r]);
                        nrm[r] = fma(v.y, v.y, nrm[r]);
                    }
                }
            }
        }
        if (EPI == 3 && norm_part) {  // per-warp partials of ‖ψ_b‖² in a fixed order: 16 slots per trajectory, unused ones zero
#pragma unroll
            for (int r = 0; r < R; ++r) {
                double v = nrm[r];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (live[r]) {
                    if (lane == 0) norm_part[(size_t)(b0 + r) * 16 + warp] = v;
                    if (warp == 0 && lane >= NW && lane < 16) norm_part[(size_t)(b0 + r) * 16 + lane] = 0.0;
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (live[r]) sb_arrive(bar_base + 8 * (kMaxSlots + (i * R + r) % ns));
        }
    }
}

// svals[e] = src[e] ≥ 0 ? ell[src[e]] : 0   (complex / real part / imaginary part)
__global__ void sell_fill_kernel(long long len, const int32_t* __restrict__ src, const c128* __restrict__ ell, c128* __restrict__ out,
                                 double* __restrict__ out_real, int take_imag) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= len) return;
    const int32_t s = src[e];
    const c128 v = s >= 0 ? ell[s] : make_double2(0.0, 0.0);
    if (out) out[e] = v;
    if (out_real) out_real[e] = take_imag ? v.y : v.x;
}

}  // namespace

struct SellOffPlan {
    const TermDev* term = nullptr;
    int list_idx = 0;
    uint32_t* d_off = nullptr;
    uint16_t* d_cols = nullptr;
    int32_t* d_src = nullptr;
    c128* d_vals = nullptr;
    double* d_rvals = nullptr;
    long long len = 0;  // entries (32 per column)
    bool real = false, imag = false;
};
struct SellPlan {
    std::vector<const TermDev*> key;
    bool usable = false;
    int npos = 0;
    uint16_t* d_sigma = nullptr;
    std::vector<SellOffPlan> off;
    double padding = 1.0;  // stored entries ÷ real entries
    bool identity = true;  // σ(p) = p: stores (and the fused RK4 operands) are coalesced
};
struct SellCache {
    std::vector<SellPlan*> plans;
};

namespace {

void free_plan(SellPlan* p) {
    if (!p) return;
    if (p->d_sigma) cudaFree(p->d_sigma);
    for (auto& o : p->off) {
        if (o.d_off) cudaFree(o.d_off);
        if (o.d_cols) cudaFree(o.d_cols);
        if (o.d_src) cudaFree(o.d_src);
        if (o.d_vals) cudaFree(o.d_vals);
        if (o.d_rvals) cudaFree(o.d_rvals);
    }
    delete p;
}

// Proper edge colouring of a bipartite multigraph rows(8) × bank groups(8) with Δ = max degree colours (alternating-path
// recolouring).  colour[e] ∈ [0, Δ).
int colour_group(const std::vector<std::pair<int, int>>& edges, std::vector<int>& colour) {
    int dr[8] = {0}, db[8] = {0};
    for (auto& e : edges) { dr[e.first]++; db[e.second]++; }
    int delta = 0;
    for (int i = 0; i < 8; ++i) delta = std::max(delta, std::max(dr[i], db[i]));
    colour.assign(edges.size(), -1);
    if (delta == 0) return 0;
    std::vector<int> rc(8 * delta, -1), bc(8 * delta, -1);  // edge holding colour c at row / bank vertex
    for (int e = 0; e < (int)edges.size(); ++e) {
        const int u = edges[e].first, v = edges[e].second;
        int a = 0, b = 0;
        while (rc[u * delta + a] >= 0) ++a;
        while (bc[v * delta + b] >= 0) ++b;
        if (bc[v * delta + a] >= 0) {
            // a is taken at v: flip the a/b alternating path that starts at v
            std::vector<int> path;
            int vert = v, side = 1, want = a;
            while (true) {
                const int e1 = side ? bc[vert * delta + want] : rc[vert * delta + want];
                if (e1 < 0) break;
                path.push_back(e1);
                vert = side ? edges[e1].first : edges[e1].second;
                side ^= 1;
                want = want == a ? b : a;
            }
            for (int e1 : path) {
                rc[edges[e1].first * delta + colour[e1]] = -1;
                bc[edges[e1].second * delta + colour[e1]] = -1;
            }
            for (int e1 : path) {
                colour[e1] = colour[e1] == a ? b : a;
                rc[edges[e1].first * delta + colour[e1]] = e1;
                bc[edges[e1].second * delta + colour[e1]] = e1;
            }
        }
        colour[e] = a;
        rc[u * delta + a] = e;
        bc[v * delta + a] = e;
    }
    return delta;
}

// ---- host planner (no CUDA calls: oqb_sell_plan_probe runs it without a GPU) ---------------------------------------------
struct EllPattern {  // one off-diagonal term: entry j < cnt[r] of row r has column cols[j*n + r]
    const int32_t* cols;
    const int* cnt;
};
struct HostTermPlan {
    std::vector<uint32_t> off;   // [nslices + 1]
    std::vector<uint16_t> cols;  // [off.back() * 32] (padded entries: a filler column of an unused bank group)
    std::vector<int32_t> src;    // ELL index j*n + r of the entry, −1 = padding
};
struct HostPlan {
    int npos = 0;
    std::vector<uint16_t> sigma;
    std::vector<HostTermPlan> terms;
    long long real_entries = 0, stored_entries = 0;
    bool identity = true;
};

void plan_host(int n, const std::vector<EllPattern>& pats, int unit, HostPlan& hp) {
    const int npos = (n + 31) / 32 * 32;
    hp.npos = npos;
    // rows by total length; the term with most entries decides the bank balance of the groups
    std::vector<int> deg(n, 0);
    size_t primary = 0;
    long long best_nnz = -1;
    for (size_t i = 0; i < pats.size(); ++i) {
        long long nnz = 0;
        for (int r = 0; r < n; ++r) { deg[r] += pats[i].cnt[r]; nnz += pats[i].cnt[r]; }
        if (nnz > best_nnz) { best_nnz = nnz; primary = i; }
    }
    // Rows travel in units of 2 neighbours (2k, 2k+1): their 16-byte results share a 32-byte sector, and a sector
    // written half by one warp and half by another costs a fill read and a second write in L2/HBM (measured: 12.5 GB of
    // DRAM traffic per launch instead of 8.6 GB with single-row units).  Units by length, longest first (stable: equal
    // lengths keep their order, so regular operators stay unpermuted); a trailing single row (odd n) goes last.
    const int nunits = (n + unit - 1) / unit;
    std::vector<int> udeg(nunits, 0);
    for (int u = 0; u < nunits; ++u)
        for (int r = u * unit; r < std::min(n, (u + 1) * unit); ++r) udeg[u] = std::max(udeg[u], deg[r]);
    if (n % unit) udeg[nunits - 1] = -1;
    std::vector<int> sorted(nunits);
    std::iota(sorted.begin(), sorted.end(), 0);
    std::stable_sort(sorted.begin(), sorted.end(), [&](int a, int b) { return udeg[a] > udeg[b]; });
    // 8-row groups: first unused unit, then more out of a window of similar units, each minimising the largest bank-group
    // degree of the group (ties: the earliest unit)
    const EllPattern& P = pats[primary];
    std::vector<uint16_t>& sigma = hp.sigma;
    sigma.assign(npos, kPad);
    std::vector<char> used(nunits, 0);
    constexpr int kWindow = 1024;  // candidates per pick: 64 → 1024 takes the random-pattern padding from 1.23 to 1.18
    int head = 0, pos = 0;
    auto add_unit = [&](int u, int* bk) {
        for (int r = u * unit; r < std::min(n, (u + 1) * unit); ++r)
            for (int j = 0; j < P.cnt[r]; ++j) bk[P.cols[(size_t)j * n + r] & 7]++;
    };
    auto place = [&](int u) {
        used[u] = 1;
        for (int r = u * unit; r < std::min(n, (u + 1) * unit); ++r) sigma[pos++] = (uint16_t)r;
    };
    while (pos < n) {
        while (head < nunits && used[sorted[head]]) ++head;
        int bank[8] = {0};
        const int u0 = sorted[head];
        place(u0);
        add_unit(u0, bank);
        while (pos % 8 != 0 && pos < n) {
            int best = -1, best_max = 1 << 30, seen = 0;
            for (int q = head + 1; q < nunits && seen < kWindow; ++q) {
                const int u = sorted[q];
                if (used[u]) continue;
                ++seen;
                int bk[8];
                memcpy(bk, bank, sizeof(bk));
                add_unit(u, bk);
                int mx = 0;
                for (int i = 0; i < 8; ++i) mx = std::max(mx, bk[i]);
                mx = std::max(mx, udeg[u0]);  // below the row degree nothing is gained
                if (mx < best_max) { best_max = mx; best = u; }
            }
            if (best < 0) break;
            place(best);
            add_unit(best, bank);
        }
    }
    for (int p = 0; p < n; ++p) hp.identity = hp.identity && sigma[p] == p;
    // per term: colour every 8-row group, slice widths = max over the 4 groups of a slice
    const int nslices = npos / 32, ngroups = npos / 8;
    for (const EllPattern& T : pats) {
        HostTermPlan tp;
        tp.off.assign(nslices + 1, 0);
        std::vector<std::vector<int>> gcol(ngroups);
        std::vector<int> gdelta(ngroups, 0);
        std::vector<std::vector<std::pair<int, int>>> gedges(ngroups);
        std::vector<std::vector<std::pair<int, int32_t>>> gent(ngroups);  // (col, ELL index) per edge
        for (int g = 0; g < ngroups; ++g) {
            for (int i = 0; i < 8; ++i) {
                const unsigned r = sigma[g * 8 + i];
                if (r == kPad) continue;
                for (int j = 0; j < T.cnt[r]; ++j) {
                    const int col = T.cols[(size_t)j * n + r];
                    gedges[g].push_back({i, col & 7});
                    gent[g].push_back({col, (int32_t)((size_t)j * n + r)});
                }
            }
            gdelta[g] = colour_group(gedges[g], gcol[g]);
            hp.real_entries += (long long)gedges[g].size();
        }
        for (int s = 0; s < nslices; ++s) {
            int w = 0;
            for (int q = 0; q < 4; ++q) w = std::max(w, gdelta[s * 4 + q]);
            tp.off[s + 1] = tp.off[s] + (uint32_t)w;
        }
        const long long len = (long long)tp.off[nslices] * 32;
        hp.stored_entries += len;
        tp.cols.assign((size_t)std::max<long long>(len, 1), kPad);
        tp.src.assign((size_t)std::max<long long>(len, 1), -1);
        for (int g = 0; g < ngroups; ++g) {
            const int s = g / 4, lane0 = (g % 4) * 8;
            for (size_t e = 0; e < gedges[g].size(); ++e) {
                const size_t at = ((size_t)tp.off[s] + gcol[g][e]) * 32 + lane0 + gedges[g][e].first;
                tp.cols[at] = (uint16_t)gent[g][e].first;
                tp.src[at] = gent[g][e].second;
            }
            // padded entries carry the value 0 and gather ψ[b] for a bank group b that no real entry of the phase uses
            for (uint32_t j = tp.off[s]; j < tp.off[s + 1]; ++j) {
                bool taken[8] = {false};
                for (int i = 0; i < 8; ++i) {
                    const size_t at = (size_t)j * 32 + lane0 + i;
                    if (tp.cols[at] != kPad) taken[tp.cols[at] & 7] = true;
                }
                int b = 0;
                for (int i = 0; i < 8; ++i) {
                    const size_t at = (size_t)j * 32 + lane0 + i;
                    if (tp.cols[at] != kPad) continue;
                    while (taken[b]) ++b;
                    taken[b] = true;
                    tp.cols[at] = (uint16_t)b;
                }
            }
        }
        hp.terms.push_back(std::move(tp));
    }
}

int build_plan(Ctx* c, int n, const std::vector<const TermDev*>& ts, SellPlan* plan) {
    plan->key = ts;
    plan->usable = false;
    if (n > 8192 || n < 32) return 0;
    std::vector<int> offs;
    std::vector<EllPattern> pats;
    int ndiag = 0;
    for (int i = 0; i < (int)ts.size(); ++i) {
        if (ts[i]->is_diag) ++ndiag;
        else {
            if ((int)ts[i]->h_cnt.size() != n) return 0;
            offs.push_back(i);
            pats.push_back({ts[i]->h_cols.data(), ts[i]->h_cnt.data()});
        }
    }
    if (offs.empty() || (int)offs.size() > kSellOff || ndiag > kSellDiag) return 0;
    HostPlan hp;
    plan_host(n, pats, c->opt.sell_unit >= 2 ? 2 : 1, hp);
    for (size_t k = 0; k < offs.size(); ++k) {
        const TermDev& T = *ts[offs[k]];
        const HostTermPlan& tp = hp.terms[k];
        SellOffPlan op;
        op.term = &T;
        op.list_idx = offs[k];
        op.real = T.ell_real && !T.dynamic_vals;
        op.imag = T.ell_imag && !T.dynamic_vals && !op.real;
        op.len = (long long)tp.off.back() * 32;
        OQB_CUDA(c, cudaMalloc((void**)&op.d_off, tp.off.size() * sizeof(uint32_t)));
        OQB_CUDA(c, cudaMalloc((void**)&op.d_cols, tp.cols.size() * sizeof(uint16_t)));
        OQB_CUDA(c, cudaMalloc((void**)&op.d_src, tp.src.size() * sizeof(int32_t)));
        if (op.real || op.imag) OQB_CUDA(c, cudaMalloc((void**)&op.d_rvals, tp.cols.size() * sizeof(double)));
        else OQB_CUDA(c, cudaMalloc((void**)&op.d_vals, tp.cols.size() * sizeof(c128)));
        OQB_CUDA(c, cudaMemcpyAsync(op.d_off, tp.off.data(), tp.off.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        OQB_CUDA(c, cudaMemcpyAsync(op.d_cols, tp.cols.data(), tp.cols.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream));
        OQB_CUDA(c, cudaMemcpyAsync(op.d_src, tp.src.data(), tp.src.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        plan->off.push_back(op);
    }
    OQB_CUDA(c, cudaMalloc((void**)&plan->d_sigma, hp.sigma.size() * sizeof(uint16_t)));
    OQB_CUDA(c, cudaMemcpyAsync(plan->d_sigma, hp.sigma.data(), hp.sigma.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream));
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));  // the host vectors die with this scope
    plan->npos = hp.npos;
    plan->identity = hp.identity;
    plan->padding = hp.real_entries ? (double)hp.stored_entries / (double)hp.real_entries : 1.0;
    plan->usable = true;
    return 0;
}

int fill_values(Ctx* c, SellOffPlan& o) {
    if (o.len <= 0) return 0;
    sell_fill_kernel<<<(unsigned)((o.len + 255) / 256), 256, 0, c->stream>>>(o.len, o.d_src, o.term->vals, o.d_vals, o.d_rvals,
                                                                             o.imag ? 1 : 0);
    OQB_LAUNCH_CHECK(c);
    return 0;
}

size_t sell_smem(int n, int ns, int npos, int noff) {
    return (size_t)ns * (n + kCoefPad) * sizeof(c128) + (size_t)((npos + 7) & ~7) * sizeof(uint16_t) +
           (size_t)noff * (npos / 32 + 1) * sizeof(unsigned);
}

template <int R, int NT, int U, int EPI>
int launch(Ctx* c, int n, int ns, int nb, const SellDesc& d, const c128* coef, long long coef_ld, int ncoef, const c128* psi,
           c128* dpsi, const RkEpilogue* rk) {
    static DevOnce cfg;
    if (!cfg.done(c->device)) {
        OQB_CUDA(c, cudaFuncSetAttribute(traj_sell_kernel<R, NT, U, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        cfg.set(c->device);
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    const long long groups = ((long long)nb + R - 1) / R;
    const int grid = (int)std::min<long long>(groups, sms);
    ProfScope prof(c, (EPI == 0 ? 32.0 : (EPI == 2 ? 80.0 : 48.0)) * n * (double)nb);
    traj_sell_kernel<R, NT, U, EPI><<<grid, NT, sell_smem(n, ns, d.npos, d.noff), c->stream>>>(
        n, ns, nb, d, coef, coef_ld, ncoef, psi, dpsi, rk ? rk->wa : 0.0, rk ? rk->wt : 0.0, rk ? rk->psi0 : nullptr,
        rk ? rk->acc : nullptr, rk ? rk->norm_part : nullptr);
    OQB_LAUNCH_CHECK(c);
    return 0;
}

template <int NT, int U>
int launch_r(Ctx* c, int R, int n, int ns, int nb, const SellDesc& d, const c128* coef, long long coef_ld, int ncoef,
             const c128* psi, c128* dpsi, const RkEpilogue* rk) {
    if (rk) {  // fused RK4 stages: two trajectories per fetch, or one when a single slot fits
        if (R >= 2) {
            switch (rk->mode) {
                case 1: return launch<2, NT, U, 1>(c, n, ns, nb, d, coef, coef_ld, ncoef, psi, dpsi, rk);
                case 2: return launch<2, NT, U, 2>(c, n, ns, nb, d, coef, coef_ld, ncoef, psi, dpsi, rk);
                default: return launch<2, NT, U, 3>(c, n, ns, nb, d, coef, coef_ld, ncoef, psi, dpsi, rk);
            }
        }
        switch (rk->mode) {
            case 1: return launch<1, NT, U, 1>(c, n, ns, nb, d, coef, coef_ld, ncoef, psi, dpsi, rk);
            case 2: return launch<1, NT, U, 2>(c, n, ns, nb, d, coef, coef_ld, ncoef, psi, dpsi, rk);
            default: return launch<1, NT, U, 3>(c, n, ns, nb, d, coef, coef_ld, ncoef, psi, dpsi, rk);
        }
    }
    if (R >= 3) return launch<3, NT, U, 0>(c, n, ns, nb, d, coef, coef_ld, ncoef, psi, dpsi, nullptr);
    if (R == 2) return launch<2, NT, U, 0>(c, n, ns, nb, d, coef, coef_ld, ncoef, psi, dpsi, nullptr);
    return launch<1, NT, U, 0>(c, n, ns, nb, d, coef, coef_ld, ncoef, psi, dpsi, nullptr);
}

}  // namespace

// Host-only check of the planner (no GPU): see oqb_sell_plan_probe in include/oqb.h
int sell_plan_probe(int n, int nterms, const int64_t* const* colptr, const int32_t* const* rowval, int unit, double* padding,
                    long long* phases, long long* conflicting, int* identity, int* misplaced) {
    if (n < 32 || n > 8192 || nterms <= 0 || nterms > kSellOff) return OQB_EINVAL;
    std::vector<std::vector<int32_t>> cols(nterms);
    std::vector<std::vector<int>> cnt(nterms);
    std::vector<EllPattern> pats;
    for (int t = 0; t < nterms; ++t) {
        cnt[t].assign(n, 0);
        for (int j = 0; j < n; ++j)
            for (int64_t p = colptr[t][j]; p < colptr[t][j + 1]; ++p) {
                if (rowval[t][p] < 0 || rowval[t][p] >= n) return OQB_EINVAL;
                cnt[t][rowval[t][p]]++;
            }
        int width = 0;
        for (int r = 0; r < n; ++r) width = std::max(width, cnt[t][r]);
        cols[t].assign((size_t)std::max(width, 1) * n, 0);
        std::vector<int> fill(n, 0);
        for (int j = 0; j < n; ++j)
            for (int64_t p = colptr[t][j]; p < colptr[t][j + 1]; ++p) {
                const int r = rowval[t][p];
                cols[t][(size_t)(fill[r]++) * n + r] = j;
            }
        pats.push_back({cols[t].data(), cnt[t].data()});
    }
    HostPlan hp;
    plan_host(n, pats, unit >= 2 ? 2 : 1, hp);
    *padding = hp.real_entries ? (double)hp.stored_entries / (double)hp.real_entries : 1.0;
    *identity = hp.identity ? 1 : 0;
    *phases = *conflicting = 0;
    *misplaced = 0;
    // σ is a permutation of the rows
    std::vector<int> seen_row(n, 0);
    for (int p = 0; p < hp.npos; ++p)
        if (hp.sigma[p] != kPad) {
            if (hp.sigma[p] >= n || seen_row[hp.sigma[p]]++) ++*misplaced;
        }
    for (int r = 0; r < n; ++r)
        if (!seen_row[r]) ++*misplaced;
    for (int t = 0; t < nterms; ++t) {
        const HostTermPlan& tp = hp.terms[t];
        std::vector<char> hit((size_t)cols[t].size(), 0);
        const int nslices = hp.npos / 32;
        for (int s = 0; s < nslices; ++s)
            for (uint32_t j = tp.off[s]; j < tp.off[s + 1]; ++j)
                for (int q = 0; q < 4; ++q) {  // one LDS.128 phase = 8 consecutive lanes
                    int banks[8] = {0};
                    bool clash = false;
                    for (int i = 0; i < 8; ++i) {
                        const size_t at = (size_t)j * 32 + q * 8 + i;
                        const int32_t e = tp.src[at];
                        if (++banks[tp.cols[at] & 7] > 1) clash = true;  // fillers count: they are gathered too
                        if (e < 0) continue;
                        const unsigned r = hp.sigma[s * 32 + q * 8 + i];
                        // the entry must belong to the row this lane computes and carry that entry's column
                        if (r == kPad || (size_t)e >= hit.size() || e % n != (int)r || e / n >= cnt[t][r] || hit[e]++ ||
                            cols[t][e] != tp.cols[at])
                            ++*misplaced;
                    }
                    ++*phases;
                    if (clash) ++*conflicting;
                }
        for (int r = 0; r < n; ++r)
            for (int j = 0; j < cnt[t][r]; ++j)
                if (!hit[(size_t)j * n + r]) ++*misplaced;  // an entry the plan lost
    }
    return 0;
}

void sell_cache_free(SellCache* cache) {
    if (!cache) return;
    for (SellPlan* p : cache->plans) free_plan(p);
    delete cache;
}

void sell_cache_stats(const SellCache* cache, int* plans, double* padding) {
    *plans = 0;
    *padding = 0.0;
    if (!cache) return;
    for (const SellPlan* p : cache->plans)
        if (p->usable) { ++*plans; *padding = p->padding; }
}

int traj_matvec_sell(Ctx* c, int n, int nb, const std::vector<const TermDev*>& ts, const c128* d_coef, long long coef_ld,
                     const c128* psi, c128* dpsi, bool* used, SellCache** cache, const RkEpilogue* rk) {
    *used = false;
    if (nb <= 0 || n > 8192 || n < 32) return 0;
    if (!*cache) *cache = new SellCache();
    SellPlan* plan = nullptr;
    for (SellPlan* p : (*cache)->plans)
        if (p->key == ts) { plan = p; break; }
    if (!plan) {
        plan = new SellPlan();
        const int s = build_plan(c, n, ts, plan);
        if (s) { free_plan(plan); return s; }
        if (plan->usable)
            for (auto& o : plan->off)
                if (!o.term->dynamic_vals) OQB_TRY(fill_values(c, o));
        if ((*cache)->plans.size() >= 8) {  // a handle sees a handful of term lists (shared/per-member γ, with/without jumps)
            free_plan((*cache)->plans.front());
            (*cache)->plans.erase((*cache)->plans.begin());
        }
        (*cache)->plans.push_back(plan);
    }
    if (!plan->usable) return 0;
    // Fused RK4 stages read ψ0/acc and write acc/out at σ(p): with a permuted σ those are 32-byte accesses scattered over the
    // trajectory — measured SLOWER than the separate stage kernels (35.9 vs 34.4 ms per step on the random pattern, 20.9 vs
    // 28.1 ms with σ = identity): the caller then runs the unfused sequence, whose matvecs come back here without `rk`.
    if (rk && !plan->identity) return 0;
    SellDesc d;
    memset(&d, 0, sizeof(d));
    d.npos = plan->npos;
    d.sigma = plan->d_sigma;
    for (auto& o : plan->off) {
        if (o.term->dynamic_vals) OQB_TRY(fill_values(c, o));
        SellOffDev& od = d.o[d.noff++];
        od.off = o.d_off; od.cols = o.d_cols; od.vals = o.d_vals; od.rvals = o.d_rvals;
        od.is_real = (o.real || o.imag) ? 1 : 0;
        od.x_imag = o.imag ? 1 : 0;
        od.coef = o.list_idx;
    }
    for (int i = 0; i < (int)ts.size(); ++i) {
        if (!ts[i]->is_diag) continue;
        SellDiagDev& g = d.g[d.ndiag++];
        g.coef = i;
        g.is_const = ts[i]->diag_const ? 1 : 0;
        g.cre = ts[i]->const_val.x; g.cim = ts[i]->const_val.y;
        g.vals = ts[i]->vals;
        g.rvals = (ts[i]->is_real && ts[i]->rvals) ? ts[i]->rvals : nullptr;
    }
    const int ncoef = (int)ts.size();
    if (ncoef > kCoefPad) return 0;
    const size_t fixed = sell_smem(n, 0, plan->npos, d.noff);
    const size_t slot = (size_t)(n + kCoefPad) * sizeof(c128);
    if (fixed + slot > 224 * 1024) return 0;
    const int ns = (int)std::min<size_t>(kMaxSlots, (224 * 1024 - fixed) / slot);
    // R: every operator entry fetched from L2 serves R trajectories.  Measured on 65 536 × 4096 (profiles/r2_traj_sell.md):
    // R = 1 is twice as slow as R = 2 (the fetch round trips dominate); R = 3 needs a third accumulator set (spills at the
    // 128-register cap of 512 threads) and, with 3 slots of 64 KB, leaves no slot to prefetch into — no faster than R = 2.
    const int rmode = c->opt.sell_r;  // 0 = pick
    int R = rmode > 0 ? rmode : std::min(ns, 2);
    if (R > ns) R = ns;
    if (n >= 2048) OQB_TRY((launch_r<512, 8>(c, R, n, ns, nb, d, d_coef, coef_ld, ncoef, psi, dpsi, rk)));
    else OQB_TRY((launch_r<256, 8>(c, R, n, ns, nb, d, d_coef, coef_ld, ncoef, psi, dpsi, rk)));
    *used = true;
    return 0;
}

}  // namespace oqb
