// Bath functions evaluated on the device (shared by ame_gaps.cu and bath.cu):
//   γ(ω) of the Ohmic bath, reference src/bath/ohmic.jl:35-41
//   the quadratic B-spline that construct_interpolations returns for a range grid, src/interpolation.jl:22-34
#pragma once
#include <cmath>

namespace oqb {

// γ(ω) = 2πη·ω·exp(−|ω|/ωc)/(1 − exp(−βω)),  γ(0) = 2πη/β  (isapprox(ω, 0, atol=1e-9))
__device__ __forceinline__ double ohmic_spectrum(double w, double eta, double wc, double beta) {
    const double two_pi = 6.283185307179586;
    if (fabs(w) <= 1e-9) return __ddiv_rn(__dmul_rn(two_pi, eta), beta);
    const double num = __dmul_rn(__dmul_rn(__dmul_rn(two_pi, eta), w), exp(-fabs(w) / wc));
    return __ddiv_rn(num, 1.0 - exp(-beta * w));
}

// BSpline(Quadratic(Line(OnGrid()))) scaled to the range x0 + k·dx (k = 0…n−1), fill value 0 outside:
// coef holds the n+2 prefiltered coefficients c₀…c_{n+1};  i = round(u), δ = u − i,
// value = c_i·½(δ−½)² + c_{i+1}·(¾−δ²) + c_{i+2}·½(δ+½)²
__device__ __forceinline__ double bspline2_eval(double x, int n, double x0, double dx, const double* __restrict__ coef) {
    const double u = (x - x0) / dx;
    if (!(u >= 0.0) || u > (double)(n - 1)) return 0.0;
    int i = (int)floor(u + 0.5);
    i = i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
    const double d = u - (double)i;
    const double dm = d - 0.5, dp = d + 0.5;
    return coef[i] * (0.5 * (dm * dm)) + coef[i + 1] * (0.75 - d * d) + coef[i + 2] * (0.5 * (dp * dp));
}

}  // namespace oqb
