/* c_abi_demo.c — the C ABI of include/oqb.h used from plain C (no Python, no torch): what a Julia `ccall`, a cgo or a JNI
 * binding does.  Evaluates the reference's Lindblad known answer (test/opensys/lindblad.jl:25-31):
 *     H = 0,  L = σz with γ = 0.5,  ρ = |+x><+x|   →   dρ = [0 −0.5; −0.5 0]
 * and the Hamiltonian commutator known answer (test/hamiltonian/dense_hamiltonian.jl:35-38) on a batch of 3 copies.
 *
 *   gcc -O2 -I include examples/c_abi_demo.c -L openquantumbase.jl_b200 -loqb_b200 -Wl,-rpath,$PWD/openquantumbase.jl_b200 -lm -o c_abi_demo
 *   ./c_abi_demo        # exit code 0 when both results match
 */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "oqb.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        int st_ = (call);                                                             \
        if (st_ != 0) {                                                               \
            fprintf(stderr, "%s failed (%d): %s\n", #call, st_, oqb_last_error(ctx)); \
            return 2;                                                                 \
        }                                                                             \
    } while (0)

int main(void) {
    oqb_ctx* ctx = NULL;
    if (oqb_create(0, &ctx) != 0 || !ctx) {
        fprintf(stderr, "oqb_create failed: a B200 (sm_100) is required, there is no CPU fallback\n");
        return 3;
    }
    const double PI = 3.14159265358979323846;
    /* column-major interleaved complex 2x2 matrices: (re,im) pairs, columns first */
    const double sx[8] = {0, 0, 1, 0, 1, 0, 0, 0};
    const double sz[8] = {1, 0, 0, 0, 0, 0, -1, 0};
    double terms[16]; /* m_1 = 2π σx, m_2 = 2π σz  (DenseHamiltonian with unit :h) */
    for (int i = 0; i < 8; ++i) {
        terms[i] = 2 * PI * sx[i];
        terms[8 + i] = 2 * PI * sz[i];
    }
    oqb_dense* op = NULL;
    CHECK(oqb_dense_create(ctx, 2, 2, terms, 1, sz, &op));

    enum { B = 3 };
    double rho[B * 8], out[B * 8];
    for (int b = 0; b < B; ++b) /* |+x><+x| = [[.5,.5],[.5,.5]] */
        for (int e = 0; e < 4; ++e) {
            rho[b * 8 + 2 * e] = 0.5;
            rho[b * 8 + 2 * e + 1] = 0.0;
        }
    void *d_u = NULL, *d_du = NULL;
    CHECK(oqb_malloc(ctx, sizeof(rho), &d_u));
    CHECK(oqb_malloc(ctx, sizeof(rho), &d_du));
    CHECK(oqb_upload(ctx, d_u, rho, sizeof(rho)));

    /* 1. Lindblad only: H coefficients 0, γ = 0.5 (shared by the batch) */
    double coefH[2] = {0.0, 0.0}, gamma[1] = {0.5};
    CHECK(oqb_dense_rhs(op, B, coefH, 1, gamma, 1, d_u, d_du));
    CHECK(oqb_sync(ctx));
    CHECK(oqb_download(ctx, out, d_du, sizeof(out)));
    const double want1[8] = {0, 0, -0.5, 0, -0.5, 0, 0, 0};
    double err1 = 0;
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < 8; ++i) err1 = fmax(err1, fabs(out[b * 8 + i] - want1[i]));

    /* 2. commutator only at s = 0.5: H = π(σx+σz); dρ = −iπ[(σx+σz), ρ] = −iπ [[0,-1],[1,0]]·… computed here in C */
    double coef2[2] = {0.5, 0.5};
    CHECK(oqb_dense_rhs(op, B, coef2, 1, NULL, 1, d_u, d_du));
    CHECK(oqb_sync(ctx));
    CHECK(oqb_download(ctx, out, d_du, sizeof(out)));
    /* (σx+σz)ρ − ρ(σx+σz) with ρ = ½[[1,1],[1,1]] is [[0,1],[−1,0]]... times −iπ: du = [[0,−iπ],[iπ,0]] (column-major below) */
    const double want2[8] = {0, 0, 0, PI, 0, -PI, 0, 0};
    double err2 = 0;
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < 8; ++i) err2 = fmax(err2, fabs(out[b * 8 + i] - want2[i]));

    uint64_t launches = 0;
    oqb_launch_count(ctx, &launches);
    printf("c_abi_demo: lindblad max|err| = %.3e, commutator max|err| = %.3e, kernels launched = %llu\n", err1, err2,
           (unsigned long long)launches);
    oqb_free(ctx, d_u);
    oqb_free(ctx, d_du);
    oqb_dense_destroy(op);
    oqb_destroy(ctx);
    return (err1 < 1e-14 && err2 < 1e-13) ? 0 : 1;
}
