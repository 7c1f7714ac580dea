// On-device stepping of a quantum-trajectory ensemble: fixed-step RK4 of dψ/dt = cache(t)·ψ with
// the jump event loop of the external solver's callback, so that ψ never leaves HBM between RHS evaluations.
//
// What the reference leaves to its caller ([external] OpenQuantumTools): the matvec `cache*u` with
// cache = −iH(s) − ½Σγ_k L_k'L_k (src/opensys/diffeq_liouvillian.jl:78-83), the jump condition ‖ψ‖² ≤ r, the
// operator choice lindblad_jump → lind_jump (src/opensys/trajectory_jump.jl:2-12,71-74; StatsBase inverse-CDF scan),
// ψ ← L_kψ/‖L_kψ‖ and the redraw of r.  Random numbers come from one xoshiro256++ stream per trajectory (the
// algorithm of Julia's default RNG; Float64 = (x >>> 11)·2⁻⁵³), state-in/state-out in device memory, consumed in a
// fixed order (first draw: the initial threshold; per jump: the choice uniform, then the new threshold), so a
// run is reproducible bit for bit from the stream states.
#include <vector>

#include "../../include/oqb.h"
#include "common.cuh"
#include "sparse_terms.cuh"

using namespace oqb;

// internal entry point of sparse.cu: matvec with the RK4 stage arithmetic fused into the store (see RkEpilogue)
int oqb_traj_matvec_rk(oqb_sparse* sp, int nb, const double* h_coefH, int coef_shared, const double* h_gamma, int gamma_shared,
                       const c128* psi_in, c128* out, const RkEpilogue* rk, int* fused);

// accessors implemented in sparse.cu (the oqb_sparse layout is private to that file)
oqb_ctx* oqb_sparse_ctx(oqb_sparse* sp);
int oqb_sparse_dims(oqb_sparse* sp, int* n, int* nterms, int* K);

namespace {

__device__ __forceinline__ unsigned long long rotl64(unsigned long long x, int k) { return (x << k) | (x >> (64 - k)); }

__device__ __forceinline__ double xoshiro_next_f64(unsigned long long* s) {
    const unsigned long long s0 = s[0], s3 = s[3];
    const unsigned long long result = rotl64(s0 + s3, 23) + s0;
    const unsigned long long t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl64(s[3], 45);
    return (double)(result >> 11) * 0x1.0p-53;
}

__global__ void rng_init_kernel(int nb, unsigned long long seed, unsigned long long* state) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    unsigned long long x = seed ^ ((unsigned long long)b * 0xD1342543DE82EF95ull);
    for (int i = 0; i < 4; ++i) {  // splitmix64
        x += 0x9E3779B97F4A7C15ull;
        unsigned long long z = x;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        state[(size_t)b * 4 + i] = z ^ (z >> 31);
    }
}

__global__ void rng_draw_kernel(int nb, unsigned long long* state, double* out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    unsigned long long s[4];
    for (int i = 0; i < 4; ++i) s[i] = state[(size_t)b * 4 + i];
    out[b] = xoshiro_next_f64(s);
    for (int i = 0; i < 4; ++i) state[(size_t)b * 4 + i] = s[i];
}

// RK4 stage bookkeeping, 16-byte vectorised, grid-stride:
//   first: acc = ψ + wa·k,  tmp = ψ + wt·k        mid: acc += wa·k,  tmp = ψ + wt·k        last: ψ = acc + wa·k
template <int MODE>
__global__ void rk_stage_kernel(long long len, double wa, double wt, const c128* __restrict__ k, c128* __restrict__ psi,
                                c128* __restrict__ acc, c128* __restrict__ tmp) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < len; e += (long long)gridDim.x * blockDim.x) {
        const c128 kv = k[e];
        if (MODE == 0) {
            const c128 p = psi[e];
            acc[e] = make_double2(p.x + wa * kv.x, p.y + wa * kv.y);
            tmp[e] = make_double2(p.x + wt * kv.x, p.y + wt * kv.y);
        } else if (MODE == 1) {
            const c128 p = psi[e], a = acc[e];
            acc[e] = make_double2(a.x + wa * kv.x, a.y + wa * kv.y);
            tmp[e] = make_double2(p.x + wt * kv.x, p.y + wt * kv.y);
        } else {
            const c128 a = acc[e];
            psi[e] = make_double2(a.x + wa * kv.x, a.y + wa * kv.y);
        }
    }
}

// mask[b] = ‖ψ_b‖² ≤ r_b; a jumping member draws its choice uniform and its next threshold, counts and logs the event.
__global__ void jump_prepare_kernel(int nb, int step, const double* __restrict__ norm2, double* __restrict__ thr,
                                    unsigned long long* __restrict__ state, uint8_t* __restrict__ mask,
                                    double* __restrict__ uni, int32_t* __restrict__ njumps, int32_t* __restrict__ log,
                                    int max_log) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const bool jump = norm2[b] <= thr[b];
    mask[b] = jump ? 1 : 0;
    if (!jump) return;
    unsigned long long s[4];
    for (int i = 0; i < 4; ++i) s[i] = state[(size_t)b * 4 + i];
    uni[b] = xoshiro_next_f64(s);
    thr[b] = xoshiro_next_f64(s);
    for (int i = 0; i < 4; ++i) state[(size_t)b * 4 + i] = s[i];
    const int j = njumps[b];
    njumps[b] = j + 1;
    if (log && j < max_log) log[((size_t)b * max_log + j) * 2] = step;  // the choice is filled in by jump_log_kernel
}

// ‖ψ_b‖² from the 16 per-warp partials the last fused RK stage left behind (summed in slot order: deterministic)
__global__ void norm_from_parts_kernel(int nb, const double* __restrict__ part, double* __restrict__ norm2) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 16; ++w) s += part[(size_t)b * 16 + w];
    norm2[b] = s;
}

__global__ void jump_log_kernel(int nb, const uint8_t* __restrict__ mask, const int32_t* __restrict__ choice,
                                const int32_t* __restrict__ njumps, int32_t* __restrict__ log, int max_log) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb || !mask[b]) return;
    const int j = njumps[b] - 1;
    if (j < max_log) log[((size_t)b * max_log + j) * 2 + 1] = choice[b];
}

}  // namespace

extern "C" {

int oqb_traj_rng_init(oqb_sparse* sp, int nb, uint64_t seed, uint64_t* d_state) {
    OQB_DEVICE(sp, oqb_sparse_ctx(sp));
    oqb_ctx* c = oqb_sparse_ctx(sp);
    if (nb <= 0) return 0;
    if (!d_state) return set_error(c, OQB_EINVAL, "oqb_traj_rng_init: null state");
    rng_init_kernel<<<(nb + 255) / 256, 256, 0, c->stream>>>(nb, (unsigned long long)seed, (unsigned long long*)d_state);
    OQB_LAUNCH_CHECK(c);
    return 0;
}

int oqb_traj_rng_uniform(oqb_sparse* sp, int nb, uint64_t* d_state, double* d_out) {
    OQB_DEVICE(sp, oqb_sparse_ctx(sp));
    oqb_ctx* c = oqb_sparse_ctx(sp);
    if (nb <= 0) return 0;
    if (!d_state || !d_out) return set_error(c, OQB_EINVAL, "oqb_traj_rng_uniform: null argument");
    rng_draw_kernel<<<(nb + 255) / 256, 256, 0, c->stream>>>(nb, (unsigned long long*)d_state, d_out);
    OQB_LAUNCH_CHECK(c);
    return 0;
}

int oqb_traj_evolve(oqb_sparse* sp, int nb, int nsteps, double dt, const double* h_coefH, int coef_shared,
                    const double* h_gamma, int gamma_shared, void* d_psi, uint64_t* d_rng, double* d_threshold,
                    int32_t* d_njumps, int32_t* d_log, int max_log, int step0) {
    OQB_DEVICE(sp, oqb_sparse_ctx(sp));
    oqb_ctx* c = oqb_sparse_ctx(sp);
    if (nb <= 0 || nsteps <= 0) return 0;
    int n = 0, nt = 0, K = 0;
    OQB_TRY(oqb_sparse_dims(sp, &n, &nt, &K));
    if (!h_coefH || !d_psi) return set_error(c, OQB_EINVAL, "oqb_traj_evolve: null argument");
    const bool jumps = K > 0 && h_gamma != nullptr;
    if (jumps && (!d_rng || !d_threshold || !d_njumps)) return set_error(c, OQB_EINVAL, "oqb_traj_evolve: rng/threshold/njumps required with Lindblad operators");
    const long long len = (long long)n * nb;
    // scratch: k, tmp, acc (n×nb each) + norm2, uniform (nb doubles) + choice (nb int32) + mask (nb bytes)
    const size_t vec = (size_t)len * sizeof(c128);
    const size_t small = ((size_t)nb * (8 + 8 + 4 + 1 + 16 * 8) + 511) / 256 * 256;
    OQB_TRY(ws_reserve(c, 3 * vec + small));
    c128* kbuf = (c128*)c->ws;
    c128* tmp = kbuf + len;
    c128* acc = tmp + len;
    double* norm2 = (double*)(acc + len);
    double* uni = norm2 + nb;
    double* nparts = uni + nb;  // 16 per trajectory
    int32_t* choice = (int32_t*)(nparts + (size_t)16 * nb);
    uint8_t* mask = (uint8_t*)(choice + nb);
    c128* psi = (c128*)d_psi;
    const size_t rowH = (size_t)(coef_shared ? 1 : nb) * nt, rowG = (size_t)(gamma_shared ? 1 : nb) * K;
    const int grid = 148 * 8, thr = 256;
    const int gb = (nb + 255) / 256;
    const bool no_fuse = !c->opt.rk_fuse;
    bool try_fused = !no_fuse;
    for (int i = 0; i < nsteps; ++i) {
        const double* cf0 = h_coefH + (size_t)(2 * i) * rowH;
        const double* cf1 = cf0 + rowH;
        const double* cf2 = cf1 + rowH;
        const double* g0 = jumps ? h_gamma + (size_t)(2 * i) * rowG : nullptr;
        const double* g1 = jumps ? g0 + rowG : nullptr;
        const double* g2 = jumps ? g1 + rowG : nullptr;
        bool done = false;
        if (try_fused) {
            // XOR-structured operators: the stage arithmetic rides the matvec's store — 3+5+5+3 vector passes instead of 25
            RkEpilogue rk;
            rk.psi0 = psi; rk.acc = acc;
            int fused = 0;
            rk.mode = 1; rk.wa = dt / 6.0; rk.wt = 0.5 * dt;
            OQB_TRY(oqb_traj_matvec_rk(sp, nb, cf0, coef_shared, g0, gamma_shared, psi, tmp, &rk, &fused));
            if (!fused) {
                try_fused = false;
            } else {
                rk.mode = 2; rk.wa = dt / 3.0; rk.wt = 0.5 * dt;
                OQB_TRY(oqb_traj_matvec_rk(sp, nb, cf1, coef_shared, g1, gamma_shared, tmp, tmp, &rk, &fused));
                rk.mode = 2; rk.wa = dt / 3.0; rk.wt = dt;
                OQB_TRY(oqb_traj_matvec_rk(sp, nb, cf1, coef_shared, g1, gamma_shared, tmp, tmp, &rk, &fused));
                rk.mode = 3; rk.wa = dt / 6.0; rk.wt = 0.0;
                rk.norm_part = jumps ? nparts : nullptr;
                OQB_TRY(oqb_traj_matvec_rk(sp, nb, cf2, coef_shared, g2, gamma_shared, tmp, psi, &rk, &fused));
                done = true;
            }
        }
        if (!done) {
        OQB_TRY(oqb_traj_matvec(sp, nb, cf0, coef_shared, g0, gamma_shared, psi, kbuf));
        rk_stage_kernel<0><<<grid, thr, 0, c->stream>>>(len, dt / 6.0, 0.5 * dt, kbuf, psi, acc, tmp);
        OQB_LAUNCH_CHECK(c);
        OQB_TRY(oqb_traj_matvec(sp, nb, cf1, coef_shared, g1, gamma_shared, tmp, kbuf));
        rk_stage_kernel<1><<<grid, thr, 0, c->stream>>>(len, dt / 3.0, 0.5 * dt, kbuf, psi, acc, tmp);
        OQB_LAUNCH_CHECK(c);
        OQB_TRY(oqb_traj_matvec(sp, nb, cf1, coef_shared, g1, gamma_shared, tmp, kbuf));
        rk_stage_kernel<1><<<grid, thr, 0, c->stream>>>(len, dt / 3.0, dt, kbuf, psi, acc, tmp);
        OQB_LAUNCH_CHECK(c);
        OQB_TRY(oqb_traj_matvec(sp, nb, cf2, coef_shared, g2, gamma_shared, tmp, kbuf));
        rk_stage_kernel<2><<<grid, thr, 0, c->stream>>>(len, dt / 6.0, 0.0, kbuf, psi, acc, tmp);
        OQB_LAUNCH_CHECK(c);
        }
        if (jumps) {
            if (done) {
                norm_from_parts_kernel<<<gb, 256, 0, c->stream>>>(nb, nparts, norm2);
                OQB_LAUNCH_CHECK(c);
            } else {
                OQB_TRY(oqb_traj_norm2(sp, nb, psi, norm2));
            }
            jump_prepare_kernel<<<gb, 256, 0, c->stream>>>(nb, step0 + i, norm2, d_threshold, (unsigned long long*)d_rng, mask,
                                                           uni, d_njumps, d_log, max_log);
            OQB_LAUNCH_CHECK(c);
            OQB_TRY(oqb_traj_jump(sp, nb, g2, gamma_shared, psi, uni, mask, choice, nullptr, 1));
            if (d_log && max_log > 0) {
                jump_log_kernel<<<gb, 256, 0, c->stream>>>(nb, mask, choice, d_njumps, d_log, max_log);
                OQB_LAUNCH_CHECK(c);
            }
        }
    }
    return 0;
}

}  // extern "C"
