// Lamb shift of the Ohmic bath on the device — row B of the hot-path table:
//   S(ω, bath) = lambshift(ω, γ)          reference src/bath/bath_util.jl:17
//   lambshift(ω, γ) = −(1/2π)·quadgk(x → (γ(ω+x) − γ(ω−x))/x, 0, Inf)     src/integration/ext_util.jl:23-28
// and the tabulation [S(ω) for ω in ω_range] that build_lambshift (src/bath/bath_util.jl:26-38) hands to
// construct_interpolations.  The reference spends milliseconds per point (adaptive G7K15, one closure call per node);
// here one thread owns one ω and runs QuadGK's control flow itself:
//   * the semi-infinite range is mapped to t∈[0,1): x = t·den, den = 1/(1−t), integrand·den²   (QuadGK handle_infinities)
//   * evalrule on a segment = 15 nodes, Kronrod sum and embedded Gauss sum, error |Ik − Ig|·h
//   * the worst segment (binary max-heap on E, in this thread's slice of a global scratch) is bisected until
//     E ≤ atol or E ≤ rtol·|I|; the result is re-summed over the final segments as QuadGK does.
// A table of 1001 frequencies is one launch of 1001 threads; the per-thread work is a few hundred exp() calls.
//
// The same adaptive routine, instantiated on other integrands, gives the bath time scales of src/bath/bath_util.jl:55-89
//   1/τ_SB = ∫₀^∞ |C(τ)|dτ,   τ_B/τ_SB = ∫₀^T τ|C(τ)|dτ,   T_a = √(τ_SB τ_B/5)
// with the Ohmic correlation function C(τ) = η(ψ₁(1 + x₂ − x₁) + ψ₁(x₁ + x₂))/β², x₁ = iτ/β, x₂ = 1/(βωc)
// (src/bath/ohmic.jl:48-52); ψ₁ of a complex argument follows SpecialFunctions.jl: recurrence up to Re z ≥ 10, then the
// asymptotic Bernoulli series.  oqb_ohmic_correlation evaluates C on arrays of lags — the cfun of the Redfield kernels.
#include "../../include/oqb.h"
#include "bath.cuh"
#include "common.cuh"

using namespace oqb;

namespace {

// G7K15 on [-1,1]: the 8 non-negative Kronrod nodes (descending … 0), their weights, the 4 Gauss weights (QUADPACK)
__constant__ double c_xgk[8] = {0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
                                0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
                                0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                                0.207784955007898467600689403773245, 0.0};
__constant__ double c_wgk[8] = {0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
                                0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
                                0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                                0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
__constant__ double c_wg[4] = {0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
                               0.381830050505118944950369775488975, 0.417959183673469387755102040816327};

struct Seg {
    double a, b, I, E;
};

// ψ₁(z), Re z > 0 (the two arguments of the Ohmic correlation have real parts x₂ and 1 + x₂)
__device__ double2 ctrigamma(double2 z) {
    double2 psi = make_double2(0.0, 0.0);
    auto inv = [](double2 a) {
        const double d = a.x * a.x + a.y * a.y;
        return make_double2(a.x / d, -a.y / d);
    };
    auto mul = [](double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); };
    const int N = 10;
    if (z.x < (double)N) {
        const int n = N - (int)floor(z.x);
        for (int nu = 0; nu < n; ++nu) {
            const double2 t = inv(make_double2(z.x + nu, z.y));
            const double2 t2 = mul(t, t);
            psi.x += t2.x;
            psi.y += t2.y;
        }
        z.x += n;
    }
    const double2 t = inv(z);
    const double2 w = mul(t, t);
    psi.x += t.x + 0.5 * w.x;
    psi.y += t.y + 0.5 * w.y;
    const double coef[8] = {0.16666666666666666, -0.03333333333333333, 0.023809523809523808, -0.03333333333333333,
                            0.07575757575757576, -0.2531135531135531,  1.1666666666666667,   -7.092156862745098};
    double2 poly = make_double2(0.0, 0.0);
#pragma unroll
    for (int k = 7; k >= 0; --k) {
        poly = mul(poly, w);
        poly.x += coef[k];
    }
    const double2 tw = mul(mul(t, w), poly);
    return make_double2(psi.x + tw.x, psi.y + tw.y);
}

__device__ double2 ohmic_correlation(double tau, double eta, double wc, double beta) {
    const double x2 = 1.0 / beta / wc, x1 = tau / beta;  // x₁ = i·τ/β
    const double2 a = ctrigamma(make_double2(1.0 + x2, -x1)), b = ctrigamma(make_double2(x2, x1));
    const double sc = eta / (beta * beta);
    return make_double2(sc * (a.x + b.x), sc * (a.y + b.y));
}

// integrands: value at the ORIGINAL variable; MAP = 1 means QuadGK's t/(1−t) map of [0,∞)
struct LambF {  // (γ(ω+x) − γ(ω−x))/x
    double w, eta, wc, beta;
    __device__ double operator()(double x) const {
        return (ohmic_spectrum(w + x, eta, wc, beta) - ohmic_spectrum(w - x, eta, wc, beta)) / x;
    }
};
struct AbsCorrF {  // |C(τ)|  or  τ|C(τ)|
    double eta, wc, beta;
    int times_x;
    __device__ double operator()(double x) const {
        const double2 c = ohmic_correlation(x, eta, wc, beta);
        const double a = hypot(c.x, c.y);
        return times_x ? x * a : a;
    }
};

template <int MAP, class F>
__device__ __forceinline__ double integrand(const F& f, double t) {
    if (MAP) {
        const double den = 1.0 / (1.0 - t);
        return f(t * den) * den * den;
    }
    return f(t);
}

// QuadGK evalrule, n = 7 (Gauss nodes at the odd Kronrod positions, centre node last)
template <int MAP, class F>
__device__ Seg evalrule(const F& p, double a, double b) {
    const double s = 0.5 * (b - a);
    double Ig = 0.0, Ik = 0.0;
#pragma unroll 1
    for (int i = 0; i < 3; ++i) {
        const double xg = -c_xgk[2 * i + 1], xk = -c_xgk[2 * i];
        const double fg = integrand<MAP>(p, a + (1.0 + xg) * s) + integrand<MAP>(p, a + (1.0 - xg) * s);
        const double fk = integrand<MAP>(p, a + (1.0 + xk) * s) + integrand<MAP>(p, a + (1.0 - xk) * s);
        if (i == 0) {
            Ig = fg * c_wg[0];
            Ik = fg * c_wgk[1] + fk * c_wgk[0];
        } else {
            Ig = Ig + fg * c_wg[i];
            Ik = Ik + (fg * c_wgk[2 * i + 1] + fk * c_wgk[2 * i]);
        }
    }
    const double f0 = integrand<MAP>(p, a + s);
    const double x6 = -c_xgk[6];
    Ig = Ig + f0 * c_wg[3];
    Ik = Ik + (f0 * c_wgk[7] + (integrand<MAP>(p, a + (1.0 + x6) * s) + integrand<MAP>(p, a + (1.0 - x6) * s)) * c_wgk[6]);
    Seg r;
    r.a = a;
    r.b = b;
    r.I = Ik * s;
    r.E = fabs(Ik * s - Ig * s);
    return r;
}

__device__ void heap_push(Seg* h, int& m, const Seg& s) {
    int i = m++;
    while (i > 0) {
        const int par = (i - 1) >> 1;
        if (!(h[par].E < s.E)) break;
        h[i] = h[par];
        i = par;
    }
    h[i] = s;
}

__device__ Seg heap_pop(Seg* h, int& m) {
    const Seg top = h[0];
    const Seg last = h[--m];
    int i = 0;
    for (;;) {
        int c = 2 * i + 1;
        if (c >= m) break;
        if (c + 1 < m && h[c].E < h[c + 1].E) ++c;
        if (!(last.E < h[c].E)) break;
        h[i] = h[c];
        i = c;
    }
    if (m > 0) h[i] = last;
    return top;
}

// QuadGK do_quadgk/adapt on [a,b] (MAP: the mapped [0,1)): returns I, E, segment count
template <int MAP, class F>
__device__ void adaptive_gk(const F& p, double a, double b, double rtol, double atol, int maxseg, Seg* heap, double& I_out,
                            double& E_out, int& m_out) {
    int m = 0;
    Seg s0 = evalrule<MAP>(p, a, b);
    double I = s0.I, E = s0.E;
    heap_push(heap, m, s0);
    while (E > atol && E > rtol * fabs(I) && m + 1 <= maxseg) {
        const Seg s = heap_pop(heap, m);
        const double mid = (s.a + s.b) * 0.5;
        const Seg s1 = evalrule<MAP>(p, s.a, mid);
        const Seg s2 = evalrule<MAP>(p, mid, s.b);
        I = (I - s.I) + s1.I + s2.I;
        E = (E - s.E) + s1.E + s2.E;
        heap_push(heap, m, s1);
        heap_push(heap, m, s2);
    }
    I = heap[0].I;
    E = heap[0].E;
    for (int i = 1; i < m; ++i) {
        I += heap[i].I;
        E += heap[i].E;
    }
    I_out = I;
    E_out = E;
    m_out = m;
}

__global__ void ohmic_lambshift_kernel(int n, const double* __restrict__ w, double eta, double wc, double beta, double rtol,
                                       double atol, int maxseg, Seg* __restrict__ scratch, double* __restrict__ S,
                                       double* __restrict__ err, int32_t* __restrict__ nseg) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    LambF p{w[k], eta, wc, beta};
    double I, E;
    int m;
    adaptive_gk<1>(p, 0.0, 1.0, rtol, atol, maxseg, scratch + (size_t)k * maxseg, I, E, m);
    S[k] = -I / 2.0 / 3.141592653589793;
    if (err) err[k] = E / 2.0 / 3.141592653589793;
    if (nseg) nseg[k] = m;
}

// thread 0: ∫₀^∞|C|, thread 1: ∫₀^lim τ|C|  → out[0..1] = I, out[2..3] = E, nseg[0..1]
__global__ void ohmic_timescale_kernel(double eta, double wc, double beta, double lim, double rtol, double atol, int maxseg,
                                       Seg* __restrict__ scratch, double* __restrict__ out, int32_t* __restrict__ nseg) {
    const int k = threadIdx.x;
    if (k > 1) return;
    AbsCorrF p{eta, wc, beta, k};
    double I, E;
    int m;
    if (k == 0)
        adaptive_gk<1>(p, 0.0, 1.0, rtol, atol, maxseg, scratch, I, E, m);
    else
        adaptive_gk<0>(p, 0.0, lim, rtol, atol, maxseg, scratch + maxseg, I, E, m);
    out[k] = I;
    out[2 + k] = E;
    nseg[k] = m;
}

__global__ void ohmic_correlation_kernel(int n, const double* __restrict__ tau, double eta, double wc, double beta,
                                         double2* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = ohmic_correlation(tau[k], eta, wc, beta);
}

__global__ void bspline2_eval_kernel(int n, double x0, double dx, const double* __restrict__ coef, int m,
                                     const double* __restrict__ x, double* __restrict__ y) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < m) y[k] = bspline2_eval(x[k], n, x0, dx, coef);
}

}  // namespace

extern "C" {

int oqb_ohmic_lambshift(oqb_ctx* c, double eta, double wc, double beta, int n, const double* h_w, double rtol, double atol,
                        int max_segments, double* h_S, double* h_err, int32_t* h_nseg) {
    OQB_DEVICE(c, c);
    if (!c) return OQB_EINVAL;
    if (n <= 0) return 0;
    if (!h_w || !h_S) return set_error(c, OQB_EINVAL, "oqb_ohmic_lambshift: null argument");
    if (max_segments < 2 || !(wc > 0.0) || !(beta > 0.0) || !(rtol >= 0.0) || !(atol >= 0.0))
        return set_error(c, OQB_EINVAL, "oqb_ohmic_lambshift: bad parameter");
    // workspace: [w (n) | S (n) | err (n) | nseg (n, padded to 8 B) | heaps (n × max_segments segments)]
    const size_t head = (size_t)n * 8 * 4;
    const size_t bytes = head + (size_t)n * max_segments * sizeof(Seg);
    OQB_TRY(ws_reserve(c, bytes));
    double* d_w = (double*)c->ws;
    double* d_S = d_w + n;
    double* d_err = d_S + n;
    int32_t* d_ns = (int32_t*)(d_err + n);
    Seg* d_heap = (Seg*)((char*)c->ws + head);
    OQB_CUDA(c, cudaMemcpyAsync(d_w, h_w, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    ohmic_lambshift_kernel<<<(n + 63) / 64, 64, 0, c->stream>>>(n, d_w, eta, wc, beta, rtol, atol, max_segments, d_heap, d_S,
                                                               d_err, d_ns);
    OQB_LAUNCH_CHECK(c);
    OQB_CUDA(c, cudaMemcpyAsync(h_S, d_S, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    if (h_err) OQB_CUDA(c, cudaMemcpyAsync(h_err, d_err, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    if (h_nseg) OQB_CUDA(c, cudaMemcpyAsync(h_nseg, d_ns, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int oqb_ohmic_correlation(oqb_ctx* c, double eta, double wc, double beta, int n, const double* h_tau, void* h_C) {
    OQB_DEVICE(c, c);
    if (!c) return OQB_EINVAL;
    if (n <= 0) return 0;
    if (!h_tau || !h_C || !(wc > 0.0) || !(beta > 0.0)) return set_error(c, OQB_EINVAL, "oqb_ohmic_correlation: bad argument");
    OQB_TRY(ws_reserve(c, (size_t)n * 24));
    double2* d_C = (double2*)c->ws;
    double* d_t = (double*)(d_C + n);
    OQB_CUDA(c, cudaMemcpyAsync(d_t, h_tau, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    ohmic_correlation_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(n, d_t, eta, wc, beta, d_C);
    OQB_LAUNCH_CHECK(c);
    OQB_CUDA(c, cudaMemcpyAsync(h_C, d_C, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream));
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int oqb_ohmic_timescales(oqb_ctx* c, double eta, double wc, double beta, double lim, double rtol, double atol, int max_segments,
                         double* tau_sb, double* err_sb, double* tau_b, double* err_b) {
    OQB_DEVICE(c, c);
    if (!c) return OQB_EINVAL;
    if (!tau_sb || !tau_b || max_segments < 2 || !(wc > 0.0) || !(beta > 0.0) || !(lim > 0.0) || !isfinite(lim) || !(rtol >= 0.0) ||
        !(atol >= 0.0))
        return set_error(c, OQB_EINVAL, "oqb_ohmic_timescales: bad argument");
    const size_t head = 64;
    OQB_TRY(ws_reserve(c, head + 2 * (size_t)max_segments * sizeof(Seg)));
    double* d_out = (double*)c->ws;
    int32_t* d_ns = (int32_t*)(d_out + 4);
    Seg* d_heap = (Seg*)((char*)c->ws + head);
    ohmic_timescale_kernel<<<1, 32, 0, c->stream>>>(eta, wc, beta, lim, rtol, atol, max_segments, d_heap, d_out, d_ns);
    OQB_LAUNCH_CHECK(c);
    double h[4];
    OQB_CUDA(c, cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    // τ_SB = 1/res, err/res² (bath_util.jl:55-58);  τ_B = res·τ_SB, err·τ_SB (:73-76)
    *tau_sb = 1.0 / h[0];
    if (err_sb) *err_sb = h[2] / (h[0] * h[0]);
    *tau_b = h[1] * *tau_sb;
    if (err_b) *err_b = h[3] * *tau_sb;
    return 0;
}

int oqb_bspline_eval(oqb_ctx* c, int n, double x0, double dx, const double* h_coef, int m, const double* h_x, double* h_y) {
    OQB_DEVICE(c, c);
    if (!c) return OQB_EINVAL;
    if (m <= 0) return 0;
    if (n < 1 || !(dx > 0.0) || !h_coef || !h_x || !h_y) return set_error(c, OQB_EINVAL, "oqb_bspline_eval: bad argument");
    const size_t bytes = ((size_t)(n + 2) + 2 * (size_t)m) * 8;
    OQB_TRY(ws_reserve(c, bytes));
    double* d_c = (double*)c->ws;
    double* d_x = d_c + (n + 2);
    double* d_y = d_x + m;
    OQB_CUDA(c, cudaMemcpyAsync(d_c, h_coef, (size_t)(n + 2) * 8, cudaMemcpyHostToDevice, c->stream));
    OQB_CUDA(c, cudaMemcpyAsync(d_x, h_x, (size_t)m * 8, cudaMemcpyHostToDevice, c->stream));
    bspline2_eval_kernel<<<(m + 255) / 256, 256, 0, c->stream>>>(n, x0, dx, d_c, m, d_x, d_y);
    OQB_LAUNCH_CHECK(c);
    OQB_CUDA(c, cudaMemcpyAsync(h_y, d_y, (size_t)m * 8, cudaMemcpyDeviceToHost, c->stream));
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

}  // extern "C"
