/* c_liouv_demo.c — a C3-shaped AME right-hand side end to end through the C ABI only (no Python, no torch): what a Julia
 * `ccall` user gets.  6 qubits (n = 64):
 *     H(s) = −(1−s)ΣX_i + s·Σ_{i<j} J_ij Z_iZ_j   with J_ij ∈ {−1, −½, 0, ½, 1}  → many gaps coincide after rounding to 8
 *                                                    digits (multi-pair gap groups, cross terms) and, at s = 0, the levels
 *                                                    are binomially degenerate (dense gap groups)
 *     DiffEqLiouvillian(H, [DaviesGenerator(Z_1..Z_6, Ohmic(1e-4, 4, 16), S ≡ 0)], [], 64)
 * Everything of (Op::DiffEqLiouvillian{true,false})(du,u,p,t), src/opensys/diffeq_liouvillian.jl:92-105, happens inside
 * oqb_liouv_rhs_host: H(s) assembly, device eigensolver, GapIndices, Davies tables, rotations.  The result is compared
 * with a golden file written from the CPU oracle's literal gap loop (tests/golden/make_liouv_golden.py).
 *
 *   gcc -O2 -I include examples/c_liouv_demo.c -L openquantumbase.jl_b200 -loqb_b200 -Wl,-rpath,$PWD/openquantumbase.jl_b200 -lm -o c_liouv_demo
 *   ./c_liouv_demo tests/golden/liouv_c_demo.bin      # exit code 0 when every case matches to 1e-10
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oqb.h"

#define Q 6
#define N 64

#define CHECK(call)                                                                   \
    do {                                                                              \
        int st_ = (call);                                                             \
        if (st_ != 0) {                                                               \
            fprintf(stderr, "%s failed (%d): %s\n", #call, st_, oqb_last_error(ctx)); \
            return 2;                                                                 \
        }                                                                             \
    } while (0)

/* qubit 1 is the most significant bit (kron ordering of the reference's q_translate, src/matrix_util.jl:16-33) */
static double zval(int state, int qubit) { return ((state >> (Q - 1 - qubit)) & 1) ? -1.0 : 1.0; }

int main(int argc, char** argv) {
    const char* golden = argc > 1 ? argv[1] : "tests/golden/liouv_c_demo.bin";
    const double PI = 3.14159265358979323846;
    const double svals[3] = {0.0, 0.5, 1.0};
    enum { NS = 3, B = 2 };
    oqb_ctx* ctx = NULL;
    if (oqb_create(0, &ctx) != 0 || !ctx) {
        fprintf(stderr, "oqb_create failed: a B200 (sm_100) is required, there is no CPU fallback\n");
        return 3;
    }
    /* operators, column-major interleaved complex, unit :h (×2π, src/base_util.jl:14-22) */
    double* terms = calloc((size_t)2 * N * N * 2, sizeof(double));
    double* coup = calloc((size_t)Q * N * N * 2, sizeof(double));
    double* u = calloc((size_t)B * N * N * 2, sizeof(double));
    double* du = calloc((size_t)NS * B * N * N * 2, sizeof(double));
    double* want = calloc((size_t)NS * B * N * N * 2, sizeof(double));
    if (!terms || !coup || !u || !du || !want) return 4;
    for (int c = 0; c < N; ++c) {
        for (int q = 0; q < Q; ++q) { /* −ΣX: flips bit q */
            const int r = c ^ (1 << q);
            terms[2 * ((size_t)r + (size_t)c * N)] = -2 * PI;
        }
        double d = 0.0;
        for (int i = 0; i < Q; ++i)
            for (int j = i + 1; j < Q; ++j) d += (((i + 1) * 7 + (j + 1) * 3) % 5 - 2) * 0.5 * zval(c, i) * zval(c, j);
        terms[2 * ((size_t)N * N + c + (size_t)c * N)] = 2 * PI * d;
        for (int q = 0; q < Q; ++q) coup[2 * ((size_t)q * N * N + c + (size_t)c * N)] = 2 * PI * zval(c, q);
    }
    for (int b = 0; b < B; ++b)
        for (int j = 0; j < N; ++j)
            for (int i = 0; i < N; ++i) {
                double* e = u + 2 * ((size_t)b * N * N + i + (size_t)j * N);
                e[0] = sin(0.37 * i + 0.11 * j + b) / N;
                e[1] = cos(0.23 * i - 0.19 * j - b) / N;
            }
    FILE* f = fopen(golden, "rb");
    if (!f || fread(want, sizeof(double), (size_t)NS * B * N * N * 2, f) != (size_t)NS * B * N * N * 2) {
        fprintf(stderr, "cannot read golden file %s\n", golden);
        return 5;
    }
    fclose(f);

    oqb_dense* H = NULL;
    oqb_liouv* L = NULL;
    CHECK(oqb_dense_create(ctx, N, 2, terms, 0, NULL, &H));
    CHECK(oqb_liouv_create(ctx, H, N, 8, 8, Q, coup, &L));
    /* Ohmic(1e-4, 4, 16): ωc = 2π·4, β = 5·6.62607004/1.38064852/π/16  (src/bath/ohmic.jl:24-28, src/unit_util.jl:13-15) */
    const double beta = 5.0 * 6.62607004 / 1.38064852 / PI / 16.0;
    CHECK(oqb_liouv_add_davies_ohmic(L, 0, Q, 1e-4, 2 * PI * 4.0, beta, 0, 0.0, 1.0, NULL));
    double worst = 0.0;
    for (int k = 0; k < NS; ++k) {
        double coefH[2] = {1.0 - svals[k], svals[k]};
        double* out = du + (size_t)k * B * N * N * 2;
        CHECK(oqb_liouv_rhs_host(L, B, coefH, 1, u, out));
        const double* w = want + (size_t)k * B * N * N * 2;
        double num = 0.0, den = 0.0;
        for (size_t e = 0; e < (size_t)B * N * N * 2; ++e) {
            num += (out[e] - w[e]) * (out[e] - w[e]);
            den += w[e] * w[e];
        }
        const double rel = sqrt(num / den);
        printf("s = %.1f: relative Frobenius error vs the oracle's literal loop %.3e\n", svals[k], rel);
        if (!(rel <= worst)) worst = rel;
    }
    uint64_t built = 0, hits = 0;
    double ms = 0, eig = 0;
    CHECK(oqb_liouv_stats(L, &built, &hits, &ms, &eig));
    printf("eigen-systems prepared: %llu (cache hits %llu), last preparation %.2f ms of which eigensolver %.2f ms\n",
           (unsigned long long)built, (unsigned long long)hits, ms, eig);
    CHECK(oqb_liouv_destroy(L));
    CHECK(oqb_dense_destroy(H));
    oqb_destroy(ctx);
    free(terms); free(coup); free(u); free(du); free(want);
    if (!(worst < 1e-10)) {
        fprintf(stderr, "MISMATCH: worst relative error %.3e\n", worst);
        return 1;
    }
    printf("c_liouv_demo OK\n");
    return 0;
}
