// Device-built tables of the one-sided AME generator and the tabulated Lamb shift (SURVEY §8(a) rows OS and B):
//   oqb_ame_set_lambshift_spline   quadratic B-spline coefficients of build_lambshift's table (src/bath/bath_util.jl:26-38)
//   oqb_ame_set_onesided_ohmic     γ(w_b − w_a) in closed form and S from that spline at the l² UNROUNDED gaps
//                                  (src/opensys/davies.jl:368, src/opensys/opensys_util.jl:65), then Λ_p, ΣA_βΛ_p, the A_β stack
// GapIndices and the Davies tables themselves live in ame_terms.cu (oqb_ame_gaps_build, oqb_ame_add_davies*).
#include <cmath>

#include "../../include/oqb.h"
#include "ame.cuh"
#include "bath.cuh"
#include "common.cuh"

using namespace oqb;

namespace {

// one-sided AME tables at the UNROUNDED gap matrix ω_ba[a,b] = w[b] − w[a] (src/opensys/opensys_util.jl:65, davies.jl:368)
__global__ void onesided_table_kernel(int l, const double* __restrict__ w, double eta, double wc, double beta, int nS, double S_x0,
                                      double S_dx, const double* __restrict__ S_coef, double* __restrict__ gam,
                                      double* __restrict__ S) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)l * l) return;
    const int a = (int)(e % l), b = (int)(e / l);
    const double g = w[b] - w[a];
    gam[e] = ohmic_spectrum(g, eta, wc, beta);
    S[e] = nS ? bspline2_eval(g, nS, S_x0, S_dx, S_coef) : 0.0;
}

}  // namespace

extern "C" {

int oqb_ame_set_lambshift_spline(oqb_ame* a, int n, double x0, double dx, const double* h_coef) {
    OQB_DEVICE(a, a->ctx);
    if (!a) return OQB_EINVAL;
    oqb_ctx* c = a->ctx;
    if (n < 0 || (n > 0 && (!h_coef || !(dx > 0.0)))) return set_error(c, OQB_EINVAL, "oqb_ame_set_lambshift_spline: bad argument");
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    ame_free(c, a->d_Scoef);
    a->n_Stab = 0;
    if (n == 0) return 0;
    OQB_TRY(ame_alloc(c, (void**)&a->d_Scoef, (size_t)(n + 2) * 8));
    OQB_CUDA(c, cudaMemcpyAsync(a->d_Scoef, h_coef, (size_t)(n + 2) * 8, cudaMemcpyHostToDevice, c->stream));
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    a->n_Stab = n;
    a->S_x0 = x0;
    a->S_dx = dx;
    return 0;
}

int oqb_ame_set_onesided_ohmic(oqb_ame* a, int npairs, const int32_t* h_alpha, const int32_t* h_beta, double eta, double wc,
                               double beta) {
    OQB_DEVICE(a, a->ctx);
    if (!a) return OQB_EINVAL;
    oqb_ctx* c = a->ctx;
    if (!a->have_eigen) return set_error(c, OQB_EINVAL, "oqb_ame_set_onesided_ohmic: call oqb_ame_set_eigen first");
    if (npairs <= 0 || !h_alpha || !h_beta) return set_error(c, OQB_EINVAL, "oqb_ame_set_onesided_ohmic: bad arguments");
    const long long ll = (long long)a->lvl * a->lvl;
    double* d_g = nullptr;
    OQB_TRY(ame_alloc(c, (void**)&d_g, (size_t)ll * 2 * sizeof(double)));
    double* d_s = d_g + ll;
    onesided_table_kernel<<<(unsigned)((ll + 255) / 256), 256, 0, c->stream>>>(a->lvl, a->d_w, eta, wc, beta, a->n_Stab, a->S_x0,
                                                                               a->S_dx, a->d_Scoef, d_g, d_s);
    int s = 0;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) s = set_error(c, OQB_ECUDA, "onesided_table_kernel launch failed: %s", cudaGetErrorString(e));
    if (!s) s = ame_onesided_install(a, npairs, h_alpha, h_beta, d_g, d_s, 1);
    cudaStreamSynchronize(c->stream);
    cudaFreeAsync(d_g, c->stream);
    return s;
}

}  // extern "C"
