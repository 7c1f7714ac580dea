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

struct Bath {
    double w, eta, wc, beta;
};

// integrand of lambshift after QuadGK's change of variable
__device__ __forceinline__ double integrand(const Bath& p, double t) {
    const double den = 1.0 / (1.0 - t);
    const double x = t * den;
    const double f = (ohmic_spectrum(p.w + x, p.eta, p.wc, p.beta) - ohmic_spectrum(p.w - x, p.eta, p.wc, p.beta)) / x;
    return f * den * den;
}

// QuadGK evalrule, n = 7 (Gauss nodes at the odd Kronrod positions, centre node last)
__device__ Seg evalrule(const Bath& p, double a, double b) {
    const double s = 0.5 * (b - a);
    double Ig = 0.0, Ik = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double xg = -c_xgk[2 * i + 1], xk = -c_xgk[2 * i];
        const double fg = integrand(p, a + (1.0 + xg) * s) + integrand(p, a + (1.0 - xg) * s);
        const double fk = integrand(p, a + (1.0 + xk) * s) + integrand(p, a + (1.0 - xk) * s);
        if (i == 0) {
            Ig = fg * c_wg[0];
            Ik = fg * c_wgk[1] + fk * c_wgk[0];
        } else {
            Ig = Ig + fg * c_wg[i];
            Ik = Ik + (fg * c_wgk[2 * i + 1] + fk * c_wgk[2 * i]);
        }
    }
    const double f0 = integrand(p, a + s);
    const double x6 = -c_xgk[6];
    Ig = Ig + f0 * c_wg[3];
    Ik = Ik + (f0 * c_wgk[7] + (integrand(p, a + (1.0 + x6) * s) + integrand(p, a + (1.0 - x6) * s)) * c_wgk[6]);
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

__global__ void ohmic_lambshift_kernel(int n, const double* __restrict__ w, double eta, double wc, double beta, double rtol,
                                       double atol, int maxseg, Seg* __restrict__ scratch, double* __restrict__ S,
                                       double* __restrict__ err, int32_t* __restrict__ nseg) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    Bath p{w[k], eta, wc, beta};
    Seg* heap = scratch + (size_t)k * maxseg;
    int m = 0;
    Seg s0 = evalrule(p, 0.0, 1.0);
    double I = s0.I, E = s0.E;
    heap_push(heap, m, s0);
    while (E > atol && E > rtol * fabs(I) && m + 1 <= maxseg) {
        const Seg s = heap_pop(heap, m);
        const double mid = (s.a + s.b) * 0.5;
        const Seg s1 = evalrule(p, s.a, mid);
        const Seg s2 = evalrule(p, mid, s.b);
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
    S[k] = -I / 2.0 / 3.141592653589793;
    if (err) err[k] = E / 2.0 / 3.141592653589793;
    if (nseg) nseg[k] = m;
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

int oqb_bspline_eval(oqb_ctx* c, int n, double x0, double dx, const double* h_coef, int m, const double* h_x, double* h_y) {
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
