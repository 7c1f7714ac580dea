// Bandwidth-bound elementwise / small-reduction kernels of the ρ-RHS paths (sm_100a).
// All of them stream ComplexF64 data with 16-byte accesses, grid-sized over (elements, batch).
#include "common.cuh"

namespace oqb {

static inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

__device__ __forceinline__ c128 cmul(c128 a, c128 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ c128 cmulc(c128 a, c128 b) {  // a * conj(b)
    return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ void cfma(c128& acc, c128 a, c128 b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}

// ---------------------------------------------------------------------------------------------
// H(s_b) = Σ f_i(s_b) m_i   (src/hamiltonian/dense_hamiltonian.jl:60-66, :71-76 with complex coef)
// ---------------------------------------------------------------------------------------------
__global__ void lincomb_kernel(int nmat, long long len, const c128* __restrict__ mats,
                               const c128* __restrict__ coef, long long coef_ld, c128* __restrict__ out,
                               int accumulate) {
    const int b = blockIdx.y;
    const c128* cf = coef + b * coef_ld;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < len;
         e += (long long)gridDim.x * blockDim.x) {
        c128 acc = accumulate ? out[b * len + e] : make_double2(0.0, 0.0);
        for (int j = 0; j < nmat; ++j) cfma(acc, cf[j], mats[j * len + e]);
        out[b * len + e] = acc;
    }
}
int lincomb(Ctx* c, int nmat, long long len, const c128* mats, const c128* coef, long long coef_ld,
            c128* out, int nb, bool accumulate) {
    if (nb <= 0 || len <= 0) return 0;
    for (int b0 = 0; b0 < nb; b0 += 65535) {
        int cnt = nb - b0 < 65535 ? nb - b0 : 65535;
        dim3 grid(cdiv(len, 256) < 4096 ? cdiv(len, 256) : 4096, cnt);
        lincomb_kernel<<<grid, 256, 0, c->stream>>>(nmat, len, mats, coef + (long long)b0 * coef_ld, coef_ld,
                                                    out + (long long)b0 * len, accumulate ? 1 : 0);
        OQB_LAUNCH_CHECK(c);
    }
    return 0;
}

// y[b][e] += alpha · x[b·stride_x + e]   (stride_x = 0 shares x across the batch)
__global__ void axpy_batched_kernel(long long len, c128 alpha, const c128* __restrict__ x, long long stride_x,
                                    c128* __restrict__ y) {
    const c128* xb = x + (long long)blockIdx.y * stride_x;
    c128* yb = y + (long long)blockIdx.y * len;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < len; e += (long long)gridDim.x * blockDim.x) {
        const c128 v = cmul(alpha, xb[e]);
        c128 o = yb[e];
        o.x += v.x;
        o.y += v.y;
        yb[e] = o;
    }
}
int axpy_batched(Ctx* c, long long len, int nb, c128 alpha, const c128* x, long long stride_x, c128* y) {
    if (nb <= 0 || len <= 0) return 0;
    for (int b0 = 0; b0 < nb; b0 += 65535) {
        int cnt = nb - b0 < 65535 ? nb - b0 : 65535;
        dim3 grid(cdiv(len, 256) < 4096 ? cdiv(len, 256) : 4096, cnt);
        axpy_batched_kernel<<<grid, 256, 0, c->stream>>>(len, alpha, x + (long long)b0 * stride_x, stride_x, y + (long long)b0 * len);
        OQB_LAUNCH_CHECK(c);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Davies closed form, per-ρ part:  X = C∘ρ̃ ;  X_aa += Σ_b W_ab ρ̃_bb   
// ---------------------------------------------------------------------------------------------
__global__ void hadamard_kernel(int l, const c128* __restrict__ Cm, const double* __restrict__ w_sub,
                                const c128* __restrict__ R, c128* __restrict__ X, int accumulate) {
    const long long len = (long long)l * l;
    const long long off = (long long)blockIdx.y * len;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < len;
         e += (long long)gridDim.x * blockDim.x) {
        c128 f = Cm[e];
        if (w_sub) f.y += w_sub[e % l] - w_sub[e / l];  // remove the −i(w_a − w_b) Hamiltonian part
        c128 v = cmul(f, R[off + e]);
        if (accumulate) {
            c128 o = X[off + e];
            v.x += o.x;
            v.y += o.y;
        }
        X[off + e] = v;
    }
}
int hadamard(Ctx* c, int l, const c128* Cm, const double* w_sub, const c128* R, c128* X, int nb, bool accumulate) {
    if (nb <= 0) return 0;
    const long long len = (long long)l * l;
    for (int b0 = 0; b0 < nb; b0 += 65535) {
        int cnt = nb - b0 < 65535 ? nb - b0 : 65535;
        dim3 grid(cdiv(len, 256) < 8192 ? cdiv(len, 256) : 8192, cnt);
        hadamard_kernel<<<grid, 256, 0, c->stream>>>(l, Cm, w_sub, R + b0 * len, X + b0 * len, accumulate ? 1 : 0);
        OQB_LAUNCH_CHECK(c);
    }
    return 0;
}

__global__ void extract_diag_kernel(int l, const c128* __restrict__ R, c128* __restrict__ pop) {
    const int b = blockIdx.y;
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < l) pop[(long long)b * l + a] = R[(long long)b * l * l + a + (long long)a * l];
}
int extract_diag(Ctx* c, int l, const c128* R, c128* pop, int nb) {
    for (int b0 = 0; b0 < nb; b0 += 65535) {  // gridDim.y ≤ 65535
        const int cnt = nb - b0 < 65535 ? nb - b0 : 65535;
        dim3 grid(cdiv(l, 256), cnt);
        extract_diag_kernel<<<grid, 256, 0, c->stream>>>(l, R + (long long)b0 * l * l, pop + (long long)b0 * l);
        OQB_LAUNCH_CHECK(c);
    }
    return 0;
}

// one warp per (row a, member b): X_aa += Σ_b' Wt[b' + a*l] pop[b']
__global__ void davies_diag_kernel(int l, const double* __restrict__ Wt, const c128* __restrict__ pop,
                                   c128* __restrict__ X, int times_i) {
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a = blockIdx.x * (blockDim.x >> 5) + warp;
    if (a >= l) return;
    const double* wrow = Wt + (long long)a * l;
    const c128* pp = pop + (long long)b * l;
    double sr = 0.0, si = 0.0;
    for (int k = lane; k < l; k += 32) {
        double wv = wrow[k];
        c128 pv = pp[k];
        sr = fma(wv, pv.x, sr);
        si = fma(wv, pv.y, si);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
        si += __shfl_xor_sync(0xffffffffu, si, o);
    }
    if (lane == 0) {
        c128* d = X + (long long)b * l * l + a + (long long)a * l;
        c128 v = *d;
        if (times_i) {  // the imaginary part of a complex rate matrix: i·(sr + i·si)
            v.x -= si;
            v.y += sr;
        } else {
            v.x += sr;
            v.y += si;
        }
        *d = v;
    }
}
int davies_diag(Ctx* c, int l, const double* Wt, const c128* pop, c128* X, int nb, bool times_i) {
    for (int b0 = 0; b0 < nb; b0 += 65535) {
        const int cnt = nb - b0 < 65535 ? nb - b0 : 65535;
        dim3 grid(cdiv(l, 8), cnt);
        davies_diag_kernel<<<grid, 256, 0, c->stream>>>(l, Wt, pop + (long long)b0 * l, X + (long long)b0 * l * l, times_i ? 1 : 0);
        OQB_LAUNCH_CHECK(c);
    }
    return 0;
}

// X += alpha * Σ_k (K_k + K_k')  — 32x32 tiles, the adjoint read goes through shared memory
__global__ void add_adjoint_kernel(int l, int nk, double alpha, const c128* __restrict__ Kb,
                                   c128* __restrict__ X) {
    __shared__ c128 tile[32][33];
    const int b = blockIdx.z;
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    c128 acc[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[r] = make_double2(0.0, 0.0);
    for (int k = 0; k < nk; ++k) {
        const c128* Km = Kb + ((long long)b * nk + k) * l * l;
        // transposed block: rows j0.., cols i0..
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int jj = j0 + tx, ii = i0 + ty + r * 8;
            tile[ty + r * 8][tx] = (jj < l && ii < l) ? Km[jj + (long long)ii * l] : make_double2(0.0, 0.0);
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int i = i0 + tx, j = j0 + ty + r * 8;
            if (i < l && j < l) {
                c128 d = Km[i + (long long)j * l];
                c128 tt = tile[tx][ty + r * 8];  // = K[j, i]
                acc[r].x += d.x + tt.x;
                acc[r].y += d.y - tt.y;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int i = i0 + tx, j = j0 + ty + r * 8;
        if (i < l && j < l) {
            c128* d = X + (long long)b * l * l + i + (long long)j * l;
            c128 v = *d;
            v.x += alpha * acc[r].x;
            v.y += alpha * acc[r].y;
            *d = v;
        }
    }
}
int add_adjoint(Ctx* c, int l, int nk, double alpha, const c128* Kb, c128* X, int nb) {
    if (nb <= 0 || nk <= 0) return 0;
    for (int b0 = 0; b0 < nb; b0 += 65535) {
        int cnt = nb - b0 < 65535 ? nb - b0 : 65535;
        dim3 grid(cdiv(l, 32), cdiv(l, 32), cnt), block(32, 8);
        add_adjoint_kernel<<<grid, block, 0, c->stream>>>(l, nk, alpha, Kb + (long long)b0 * nk * l * l,
                                                          X + (long long)b0 * l * l);
        OQB_LAUNCH_CHECK(c);
    }
    return 0;
}

// Redfield apply with DIAGONAL couplings S_k = diag(s_k):  K₂ = Σ_k (S_k M_k − M_k S_k) has entries (s_k[i] − s_k[j])·M_k[i,j],
// so  X −= K₂ + K₂'  is one pass over the K blocks M_k = Λ_k u (32x32 tiles, the adjoint read goes through shared memory).
__global__ void redfield_diag_combine_kernel(int l, int nk, const c128* __restrict__ sdiag, const c128* __restrict__ Mb,
                                             c128* __restrict__ X) {
    __shared__ c128 tile[32][33];
    const int b = blockIdx.z;
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    c128 acc[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[r] = make_double2(0.0, 0.0);
    for (int k = 0; k < nk; ++k) {
        const c128* Km = Mb + ((long long)b * nk + k) * l * l;
        const c128* sk = sdiag + (long long)k * l;
#pragma unroll
        for (int r = 0; r < 4; ++r) {  // transposed block: element (jj, ii) weighted by (s[jj] − s[ii])
            const int jj = j0 + tx, ii = i0 + ty + r * 8;
            c128 v = make_double2(0.0, 0.0);
            if (jj < l && ii < l) {
                const c128 w = make_double2(sk[jj].x - sk[ii].x, sk[jj].y - sk[ii].y);
                v = cmul(w, Km[jj + (long long)ii * l]);
            }
            tile[ty + r * 8][tx] = v;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = i0 + tx, j = j0 + ty + r * 8;
            if (i < l && j < l) {
                const c128 w = make_double2(sk[i].x - sk[j].x, sk[i].y - sk[j].y);
                const c128 d = cmul(w, Km[i + (long long)j * l]);
                const c128 tt = tile[tx][ty + r * 8];  // = K₂_k[j, i]
                acc[r].x += d.x + tt.x;
                acc[r].y += d.y - tt.y;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = i0 + tx, j = j0 + ty + r * 8;
        if (i < l && j < l) {
            c128* d = X + (long long)b * l * l + i + (long long)j * l;
            c128 v = *d;
            v.x -= acc[r].x;
            v.y -= acc[r].y;
            *d = v;
        }
    }
}
int redfield_diag_combine(Ctx* c, int l, int nk, const c128* sdiag, const c128* Mb, c128* X, int nb) {
    if (nb <= 0) return 0;
    for (int b0 = 0; b0 < nb; b0 += 65535) {
        const int cnt = nb - b0 < 65535 ? nb - b0 : 65535;
        dim3 grid(cdiv(l, 32), cdiv(l, 32), cnt), block(32, 8);
        redfield_diag_combine_kernel<<<grid, block, 0, c->stream>>>(l, nk, sdiag, Mb + (long long)b0 * nk * l * l, X + (long long)b0 * l * l);
        OQB_LAUNCH_CHECK(c);
    }
    return 0;
}

// Redfield apply with MONOMIAL couplings (at most one entry per row and per column: σx, σy, σ±, Pauli strings):
// S_k[a, π_k(a)] = v_k[a], r_k(c) = the row holding column c's entry.  (S_k M_k − M_k S_k)[a,c] = v_k[a]·M_k[π_k(a), c] − M_k[a, r_k(c)]·v_k[r_k(c)]
// — two gathers per block instead of two n³ products; X −= K₂ + K₂' as in the diagonal kernel (32x32 tiles, adjoint through
// shared memory).
__global__ void redfield_mono_combine_kernel(int l, int nk, const int32_t* __restrict__ perm, const int32_t* __restrict__ pinv,
                                             const c128* __restrict__ val, const c128* __restrict__ Mb, c128* __restrict__ X) {
    __shared__ c128 tile[32][33];
    const int b = blockIdx.z;
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    c128 acc[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[r] = make_double2(0.0, 0.0);
    for (int k = 0; k < nk; ++k) {
        const c128* Km = Mb + ((long long)b * nk + k) * l * l;
        const int32_t* pk = perm + (long long)k * l;
        const int32_t* rk = pinv + (long long)k * l;
        const c128* vk = val + (long long)k * l;
        auto k2 = [&](int a, int cc) -> c128 {  // (S_k M_k − M_k S_k)[a, cc]
            c128 o = make_double2(0.0, 0.0);
            const int pa = pk[a], rc = rk[cc];
            if (pa >= 0) o = cmul(vk[a], Km[pa + (long long)cc * l]);
            if (rc >= 0) {
                const c128 t = cmul(Km[a + (long long)rc * l], vk[rc]);
                o.x -= t.x;
                o.y -= t.y;
            }
            return o;
        };
#pragma unroll
        for (int r = 0; r < 4; ++r) {  // transposed block: element (jj, ii)
            const int jj = j0 + tx, ii = i0 + ty + r * 8;
            tile[ty + r * 8][tx] = (jj < l && ii < l) ? k2(jj, ii) : make_double2(0.0, 0.0);
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = i0 + tx, j = j0 + ty + r * 8;
            if (i < l && j < l) {
                const c128 d = k2(i, j);
                const c128 tt = tile[tx][ty + r * 8];  // = K₂_k[j, i]
                acc[r].x += d.x + tt.x;
                acc[r].y += d.y - tt.y;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = i0 + tx, j = j0 + ty + r * 8;
        if (i < l && j < l) {
            c128* d = X + (long long)b * l * l + i + (long long)j * l;
            c128 v = *d;
            v.x -= acc[r].x;
            v.y -= acc[r].y;
            *d = v;
        }
    }
}
int redfield_mono_combine(Ctx* c, int l, int nk, const int32_t* perm, const int32_t* pinv, const c128* val, const c128* Mb, c128* X, int nb) {
    if (nb <= 0) return 0;
    for (int b0 = 0; b0 < nb; b0 += 65535) {
        const int cnt = nb - b0 < 65535 ? nb - b0 : 65535;
        dim3 grid(cdiv(l, 32), cdiv(l, 32), cnt), block(32, 8);
        redfield_mono_combine_kernel<<<grid, block, 0, c->stream>>>(l, nk, perm, pinv, val, Mb + (long long)b0 * nk * l * l, X + (long long)b0 * l * l);
        OQB_LAUNCH_CHECK(c);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Davies tables (per distinct s, shared by the batch).  One block per column b for the sums.
// ---------------------------------------------------------------------------------------------
__global__ void davies_colsum_kernel(int l, int K, const c128* __restrict__ A, const double* __restrict__ gam,
                                     const double* __restrict__ S, double* __restrict__ Gam,
                                     double* __restrict__ hls, double* __restrict__ Wt) {
    const int b = blockIdx.x;
    double sg = 0.0, sh = 0.0;
    for (int a = threadIdx.x; a < l; a += blockDim.x) {
        double m2 = 0.0;
        for (int k = 0; k < K; ++k) {
            c128 v = A[(long long)k * l * l + a + (long long)b * l];
            m2 += v.x * v.x + v.y * v.y;
        }
        double g = gam[a + (long long)b * l] * m2;
        sg += g;
        sh += S[a + (long long)b * l] * m2;
        Wt[b + (long long)a * l] = (a == b) ? 0.0 : g;  // W[a,b] stored transposed
    }
    __shared__ double red[2][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sg += __shfl_xor_sync(0xffffffffu, sg, o);
        sh += __shfl_xor_sync(0xffffffffu, sh, o);
    }
    if (lane == 0) { red[0][warp] = sg; red[1][warp] = sh; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a0 = 0.0, a1 = 0.0;
        for (int i = 0; i < (blockDim.x >> 5); ++i) { a0 += red[0][i]; a1 += red[1][i]; }
        Gam[b] = a0;
        hls[b] = a1;
    }
}
__global__ void davies_cmat_kernel(int l, int K, const c128* __restrict__ A, const double* __restrict__ gam,
                                   const double* __restrict__ w, const double* __restrict__ Gam,
                                   const double* __restrict__ hls, int with_davies, c128* __restrict__ Cm) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= (long long)l * l) return;
    int a = (int)(e % l), b = (int)(e / l);
    c128 v = make_double2(0.0, -(w[a] - w[b]));  // −i(w_a − w_b): diffeq_liouvillian.jl:99-100
    if (with_davies) {
        v.x += -0.5 * (Gam[a] + Gam[b]);
        v.y += -(hls[a] - hls[b]);
        double g0 = gam[0];
        c128 s = make_double2(0.0, 0.0);
        for (int k = 0; k < K; ++k) {
            const c128* Ak = A + (long long)k * l * l;
            c128 t = cmulc(Ak[a + (long long)a * l], Ak[b + (long long)b * l]);
            s.x += t.x;
            s.y += t.y;
        }
        v.x += g0 * s.x;
        v.y += g0 * s.y;
    }
    Cm[e] = v;
}
int davies_tables(Ctx* c, int l, int K, const c128* A, const double* gam, const double* S, const double* w,
                  int with_davies, double* Gam, double* hls, double* Wt, c128* Cm) {
    if (with_davies) {
        davies_colsum_kernel<<<l, 256, 0, c->stream>>>(l, K, A, gam, S, Gam, hls, Wt);
        OQB_LAUNCH_CHECK(c);
    }
    davies_cmat_kernel<<<cdiv((long long)l * l, 256), 256, 0, c->stream>>>(l, K, A, gam, w, Gam, hls, with_davies, Cm);
    OQB_LAUNCH_CHECK(c);
    return 0;
}

__global__ void davies_cross_kernel(int l, int K, const c128* __restrict__ A, long long ncross,
                                    const int32_t* __restrict__ idx, const double* __restrict__ rate,
                                    const c128* __restrict__ R, c128* __restrict__ X) {
    const int bb = blockIdx.y;
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= ncross) return;
    int a = idx[4 * e], b = idx[4 * e + 1], a2 = idx[4 * e + 2], b2 = idx[4 * e + 3];
    c128 s = make_double2(0.0, 0.0);
    for (int k = 0; k < K; ++k) {
        const c128* Ak = A + (long long)k * l * l;
        c128 t = cmulc(Ak[a + (long long)b * l], Ak[a2 + (long long)b2 * l]);
        s.x += t.x;
        s.y += t.y;
    }
    double r = rate[e];
    s.x *= r;
    s.y *= r;
    c128 v = cmul(s, R[(long long)bb * l * l + b + (long long)b2 * l]);
    c128* d = X + (long long)bb * l * l + a + (long long)a2 * l;
    atomicAdd(&d->x, v.x);
    atomicAdd(&d->y, v.y);
}
int davies_cross(Ctx* c, int l, int K, const c128* A, long long ncross, const int32_t* idx, const double* rate,
                 const c128* R, c128* X, int nb) {
    if (ncross <= 0 || nb <= 0) return 0;
    for (int b0 = 0; b0 < nb; b0 += 65535) {
        const int cnt = nb - b0 < 65535 ? nb - b0 : 65535;
        dim3 grid(cdiv(ncross, 256), cnt);
        davies_cross_kernel<<<grid, 256, 0, c->stream>>>(l, K, A, ncross, idx, rate, R + (long long)b0 * l * l, X + (long long)b0 * l * l);
        OQB_LAUNCH_CHECK(c);
    }
    return 0;
}

__global__ void davies_lamb_kernel(int l, int K, const c128* __restrict__ A, long long nlamb,
                                   const int32_t* __restrict__ idx, const double* __restrict__ rs,
                                   c128* __restrict__ Moff) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= nlamb) return;
    int a = idx[3 * e], b = idx[3 * e + 1], b2 = idx[3 * e + 2];
    c128 s = make_double2(0.0, 0.0);
    for (int k = 0; k < K; ++k) {
        const c128* Ak = A + (long long)k * l * l;
        c128 x = Ak[a + (long long)b * l], y = Ak[a + (long long)b2 * l];
        // conj(x) * y
        s.x += x.x * y.x + x.y * y.y;
        s.y += x.x * y.y - x.y * y.x;
    }
    c128 f = make_double2(-0.5 * rs[2 * e], -rs[2 * e + 1]);  // −½γ − iS
    c128 v = cmul(f, s);
    atomicAdd(&Moff[b + (long long)b2 * l].x, v.x);
    atomicAdd(&Moff[b + (long long)b2 * l].y, v.y);
}
int davies_lamb_build(Ctx* c, int l, int K, const c128* A, long long nlamb, const int32_t* idx,
                      const double* rate_S, c128* Moff) {
    if (nlamb <= 0) return 0;
    davies_lamb_kernel<<<cdiv(nlamb, 256), 256, 0, c->stream>>>(l, K, A, nlamb, idx, rate_S, Moff);
    OQB_LAUNCH_CHECK(c);
    return 0;
}

__global__ void davies_heff_kernel(int l, const double* __restrict__ w, const double* __restrict__ Gam,
                                   const double* __restrict__ hls, const c128* __restrict__ Moff,
                                   c128* __restrict__ D) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= (long long)l * l) return;
    int a = (int)(e % l), b = (int)(e / l);
    c128 v = Moff ? Moff[e] : make_double2(0.0, 0.0);
    if (a == b) {
        v.x += Gam ? -0.5 * Gam[a] : 0.0;
        v.y += -w[a] - (hls ? hls[a] : 0.0);
    }
    D[e] = v;
}
int davies_heff_diag(Ctx* c, int l, const double* w, const double* Gam, const double* hls, const c128* Moff,
                     c128* D) {
    davies_heff_kernel<<<cdiv((long long)l * l, 256), 256, 0, c->stream>>>(l, w, Gam, hls, Moff, D);
    OQB_LAUNCH_CHECK(c);
    return 0;
}

// Λ_p = (½γ + iS) ∘ A_α(p)      src/opensys/davies.jl:373
__global__ void onesided_lambda_kernel(int l, const int32_t* __restrict__ alpha, const c128* __restrict__ A,
                                       const double* __restrict__ gam, const double* __restrict__ S,
                                       long long table_stride, c128* __restrict__ Lam) {
    const int p = blockIdx.y;
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= (long long)l * l) return;
    c128 f = make_double2(0.5 * gam[p * table_stride + e], S[p * table_stride + e]);
    Lam[(long long)p * l * l + e] = cmul(f, A[(long long)alpha[p] * l * l + e]);
}
int onesided_lambda(Ctx* c, int l, int npairs, const int32_t* alpha, const c128* A, const double* gam,
                    const double* S, long long table_stride, c128* Lam) {
    if (npairs <= 0) return 0;
    dim3 grid(cdiv((long long)l * l, 256), npairs);
    onesided_lambda_kernel<<<grid, 256, 0, c->stream>>>(l, alpha, A, gam, S, table_stride, Lam);
    OQB_LAUNCH_CHECK(c);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// superoperator assembly (small n only)
// ---------------------------------------------------------------------------------------------
__global__ void vec_commutator_kernel(int n, const c128* __restrict__ H, long long strideH, c128* __restrict__ out) {
    const int bb = blockIdx.y;
    const long long n2 = (long long)n * n, tot = n2 * n2;
    const c128* Hb = H + bb * strideH;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot;
         e += (long long)gridDim.x * blockDim.x) {
        long long row = e % n2, col = e / n2;
        int p = (int)(row / n), q = (int)(row % n), r = (int)(col / n), s = (int)(col % n);
        // i ( H[r,p] δ(q,s) − δ(p,r) H[q,s] )
        c128 v = make_double2(0.0, 0.0);
        if (q == s) { c128 h = Hb[r + (long long)p * n]; v.x += h.x; v.y += h.y; }
        if (p == r) { c128 h = Hb[q + (long long)s * n]; v.x -= h.x; v.y -= h.y; }
        out[bb * tot + e] = make_double2(-v.y, v.x);
    }
}
int vectorized_commutator(Ctx* c, int n, const c128* H, long long strideH, c128* out, int nb) {
    if (nb <= 0) return 0;
    long long tot = (long long)n * n * n * n;
    dim3 grid(cdiv(tot, 256) < 4096 ? cdiv(tot, 256) : 4096, nb);
    vec_commutator_kernel<<<grid, 256, 0, c->stream>>>(n, H, strideH, out);
    OQB_LAUNCH_CHECK(c);
    return 0;
}

__global__ void redfield_superop_kernel(int n, const c128* __restrict__ S, long long strideS,
                                        const c128* __restrict__ Lam, long long strideLam,
                                        const c128* __restrict__ SL, long long strideSL, c128* __restrict__ cache) {
    const int bb = blockIdx.y;
    const long long n2 = (long long)n * n, tot = n2 * n2;
    const c128* Sb = S + bb * strideS;
    const c128* Lb = Lam + bb * strideLam;
    const c128* SLb = SL + bb * strideSL;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot;
         e += (long long)gridDim.x * blockDim.x) {
        long long row = e % n2, col = e / n2;
        int p = (int)(row / n), q = (int)(row % n), r = (int)(col / n), s = (int)(col % n);
        c128 v = make_double2(0.0, 0.0);
        if (p == r) { c128 x = SLb[q + (long long)s * n]; v.x += x.x; v.y += x.y; }           // I⊗SΛ
        if (q == s) { c128 x = SLb[p + (long long)r * n]; v.x += x.x; v.y -= x.y; }           // conj(SΛ)⊗I
        { c128 x = cmul(Sb[r + (long long)p * n], Lb[q + (long long)s * n]); v.x -= x.x; v.y -= x.y; }   // Sᵀ⊗Λ
        { c128 x = cmulc(Sb[q + (long long)s * n], Lb[p + (long long)r * n]); v.x -= x.x; v.y -= x.y; }  // conj(Λ)⊗S
        c128 o = cache[bb * tot + e];
        o.x -= v.x;
        o.y -= v.y;
        cache[bb * tot + e] = o;
    }
}
int redfield_superop_acc(Ctx* c, int n, const c128* S, long long strideS, const c128* Lam, long long strideLam,
                         const c128* SL, long long strideSL, c128* cache, int nb) {
    if (nb <= 0) return 0;
    long long tot = (long long)n * n * n * n;
    dim3 grid(cdiv(tot, 256) < 4096 ? cdiv(tot, 256) : 4096, nb);
    redfield_superop_kernel<<<grid, 256, 0, c->stream>>>(n, S, strideS, Lam, strideLam, SL, strideSL, cache);
    OQB_LAUNCH_CHECK(c);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// observables: one block per (observable, member)
// ---------------------------------------------------------------------------------------------
__global__ void expect_kernel(int n, const c128* __restrict__ obs, const c128* __restrict__ state, int is_vector,
                              c128* __restrict__ out) {
    const int o = blockIdx.x, b = blockIdx.y, nobs = gridDim.x;
    const c128* O = obs + (long long)o * n * n;
    double sr = 0.0, si = 0.0;
    if (!is_vector) {
        const c128* rho = state + (long long)b * n * n;
        for (long long e = threadIdx.x; e < (long long)n * n; e += blockDim.x) {
            int i = (int)(e % n), j = (int)(e / n);
            c128 v = cmul(O[e], rho[j + (long long)i * n]);  // O_ij ρ_ji
            sr += v.x;
            si += v.y;
        }
    } else {
        const c128* psi = state + (long long)b * n;
        for (long long e = threadIdx.x; e < (long long)n * n; e += blockDim.x) {
            int i = (int)(e % n), j = (int)(e / n);
            c128 v = cmul(O[e], psi[j]);
            c128 u = make_double2(psi[i].x * v.x + psi[i].y * v.y, psi[i].x * v.y - psi[i].y * v.x);  // conj(ψ_i) v
            sr += u.x;
            si += u.y;
        }
    }
    __shared__ double red[2][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, of);
        si += __shfl_xor_sync(0xffffffffu, si, of);
    }
    if (lane == 0) { red[0][warp] = sr; red[1][warp] = si; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a0 = 0.0, a1 = 0.0;
        for (int i = 0; i < (blockDim.x >> 5); ++i) { a0 += red[0][i]; a1 += red[1][i]; }
        out[(long long)b * nobs + o] = make_double2(a0, a1);
    }
}
// Diagonal observables of a ψ ensemble: out[b][o] = Σ_r d_o[r]·|ψ_b[r]|², all observables in ONE pass over ψ
// (one block per member; the NOBS accumulators live in registers).
template <int NOBS>
__global__ void expect_diag_kernel(int n, int nobs, const double* __restrict__ diag, const c128* __restrict__ psi,
                                   double* __restrict__ out) {
    const int b = blockIdx.x;
    const c128* x = psi + (long long)b * n;
    double acc[NOBS];
#pragma unroll
    for (int o = 0; o < NOBS; ++o) acc[o] = 0.0;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        const c128 v = x[r];
        const double p = v.x * v.x + v.y * v.y;
#pragma unroll
        for (int o = 0; o < NOBS; ++o)
            if (o < nobs) acc[o] = fma(diag[(long long)o * n + r], p, acc[o]);
    }
    __shared__ double red[NOBS][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 0; o < NOBS; ++o) {
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], of);
        if (lane == 0) red[o][warp] = acc[o];
    }
    __syncthreads();
    if (threadIdx.x < nobs) {
        double a = 0.0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) a += red[threadIdx.x][w];
        out[(long long)b * nobs + threadIdx.x] = a;
    }
}
int expect_diag(Ctx* c, int n, int nobs, const double* diag, int nb, const c128* psi, double* out) {
    if (nb <= 0 || nobs <= 0) return 0;
    if (nobs > 16) return set_error(c, 1, "oqb_expect_diag: at most 16 observables per call");
    if (nobs <= 4) expect_diag_kernel<4><<<nb, 256, 0, c->stream>>>(n, nobs, diag, psi, out);
    else expect_diag_kernel<16><<<nb, 256, 0, c->stream>>>(n, nobs, diag, psi, out);
    OQB_LAUNCH_CHECK(c);
    return 0;
}

// Ensemble populations P[r] = Σ_b |ψ_b[r]|² (what every diagonal ensemble average reduces to: ⟨O⟩ = d_O·P / B).
// Two deterministic stages: NG groups of trajectories → partial[g][r], then the groups summed in order.
__global__ void population_partial_kernel(int n, int nb, int per_group, const c128* __restrict__ psi, double* __restrict__ partial) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (r >= n) return;
    const int b0 = g * per_group, b1 = min(nb, b0 + per_group);
    double s = 0.0;
    for (int b = b0; b < b1; ++b) {
        const c128 v = psi[(long long)b * n + r];
        s = fma(v.x, v.x, s);
        s = fma(v.y, v.y, s);
    }
    partial[(long long)g * n + r] = s;
}
__global__ void population_final_kernel(int n, int ng, const double* __restrict__ partial, double* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double s = 0.0;
    for (int g = 0; g < ng; ++g) s += partial[(long long)g * n + r];
    out[r] = s;
}
int ensemble_population(Ctx* c, int n, int nb, const c128* psi, double* out) {
    if (nb <= 0 || n <= 0) return 0;
    int ng = 148 * 8 / ((n + 255) / 256);  // enough blocks to fill the machine
    if (ng < 1) ng = 1;
    if (ng > nb) ng = nb;
    if (ng > 4096) ng = 4096;
    const int per_group = (nb + ng - 1) / ng;
    ng = (nb + per_group - 1) / per_group;
    OQB_TRY(ws_reserve(c, (size_t)ng * n * sizeof(double)));
    double* partial = (double*)c->ws;
    dim3 grid((n + 255) / 256, ng);
    population_partial_kernel<<<grid, 256, 0, c->stream>>>(n, nb, per_group, psi, partial);
    OQB_LAUNCH_CHECK(c);
    population_final_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(n, ng, partial, out);
    OQB_LAUNCH_CHECK(c);
    return 0;
}

int expect(Ctx* c, int n, int nobs, const c128* obs, int nb, const c128* state, int is_vector, c128* out) {
    if (nb <= 0 || nobs <= 0) return 0;
    for (int b0 = 0; b0 < nb; b0 += 65535) {
        int cnt = nb - b0 < 65535 ? nb - b0 : 65535;
        dim3 grid(nobs, cnt);
        const c128* st = state + (long long)b0 * (is_vector ? n : (long long)n * n);
        expect_kernel<<<grid, 256, 0, c->stream>>>(n, obs, st, is_vector, out + (long long)b0 * nobs);
        OQB_LAUNCH_CHECK(c);
    }
    return 0;
}

}  // namespace oqb
