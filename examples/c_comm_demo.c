/* c_comm_demo.c — the ensemble-average reduce of observables through the C ABI only (SURVEY §8(e) process model: ONE
 * process, one host thread + one oqb context per GPU, ncclCommInitAll).  Every GPU holds its own block of a ψ ensemble,
 * computes its partial sums Σ_b ψ_b'O_oψ_b with oqb_expect_sum and the library reduces them with ncclAllReduce on the
 * same streams; the result is checked against the sum computed on the host.
 *
 *   gcc -O2 -I include examples/c_comm_demo.c -L openquantumbase.jl_b200 -loqb_b200 -Wl,-rpath,$PWD/openquantumbase.jl_b200 -lm -lpthread -o c_comm_demo
 *   ./c_comm_demo [ngpus]        # default: all visible GPUs (at least 1); exit code 0 when the reduced observables match
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>

#include "oqb.h"

#define N 32
#define NOBS 3
#define PER 64 /* ψ per GPU */
#define MAXG 8

#define AQ 6
#define AN 64 /* 6-qubit AME on every GPU of the process: exercises the per-device kernel configuration */

typedef struct {
    int rank, ngpu;
    oqb_ctx* ctx;
    oqb_comm* comm;
    double result[2 * NOBS];
    double ame_sum[2]; /* Σ du and Σ |du|² of a small AME right-hand side computed on this GPU */
    int status;
} worker_t;

static double zval6(int state, int qubit) { return ((state >> (AQ - 1 - qubit)) & 1) ? -1.0 : 1.0; }

/* the operators of examples/c_liouv_demo.c at s = 0.5, one rho; identical on every GPU, so the results must agree bit for bit */
static int ame_on_this_gpu(oqb_ctx* ctx, double* sums) {
    const double PI = 3.14159265358979323846;
    double* terms = calloc((size_t)2 * AN * AN * 2, sizeof(double));
    double* coup = calloc((size_t)AQ * AN * AN * 2, sizeof(double));
    double* u = calloc((size_t)AN * AN * 2, sizeof(double));
    double* du = calloc((size_t)AN * AN * 2, sizeof(double));
    int st = 1;
    oqb_dense* H = NULL;
    oqb_liouv* L = NULL;
    if (!terms || !coup || !u || !du) goto done;
    for (int c = 0; c < AN; ++c) {
        for (int q = 0; q < AQ; ++q) terms[2 * ((size_t)(c ^ (1 << q)) + (size_t)c * AN)] = -2 * PI;
        double d = 0.0;
        for (int i = 0; i < AQ; ++i)
            for (int j = i + 1; j < AQ; ++j) d += (((i + 1) * 7 + (j + 1) * 3) % 5 - 2) * 0.5 * zval6(c, i) * zval6(c, j);
        terms[2 * ((size_t)AN * AN + c + (size_t)c * AN)] = 2 * PI * d;
        for (int q = 0; q < AQ; ++q) coup[2 * ((size_t)q * AN * AN + c + (size_t)c * AN)] = 2 * PI * zval6(c, q);
    }
    for (int j = 0; j < AN; ++j)
        for (int i = 0; i < AN; ++i) {
            u[2 * (i + (size_t)j * AN)] = sin(0.37 * i + 0.11 * j) / AN;
            u[2 * (i + (size_t)j * AN) + 1] = cos(0.23 * i - 0.19 * j) / AN;
        }
    if (oqb_dense_create(ctx, AN, 2, terms, 0, NULL, &H)) goto done;
    if (oqb_liouv_create(ctx, H, AN, 8, 8, AQ, coup, &L)) goto done;
    if (oqb_liouv_add_davies_ohmic(L, 0, AQ, 1e-4, 2 * PI * 4.0, 5.0 * 6.62607004 / 1.38064852 / PI / 16.0, 0, 0.0, 1.0, NULL)) goto done;
    {
        double coefH[2] = {0.5, 0.5};
        if (oqb_liouv_rhs_host(L, 1, coefH, 1, u, du)) goto done;
    }
    sums[0] = sums[1] = 0.0;
    for (int e = 0; e < 2 * AN * AN; ++e) { sums[0] += du[e]; sums[1] += du[e] * du[e]; }
    st = 0;
done:
    if (L) oqb_liouv_destroy(L);
    if (H) oqb_dense_destroy(H);
    free(terms); free(coup); free(u); free(du);
    return st;
}

static void fill_state(int rank, int b, double* psi) { /* deterministic, not normalised: linearity is all that matters */
    for (int r = 0; r < N; ++r) {
        psi[2 * r] = sin(0.1 * (r + 1) * (b + 1) + rank);
        psi[2 * r + 1] = cos(0.07 * (r + 2) * (b + 3) - rank);
    }
}
static void fill_obs(int o, double* m) { /* Hermitian: diagonal r·(o+1), nearest-neighbour coupling (o+1)(1 ± i/2) */
    for (int e = 0; e < 2 * N * N; ++e) m[e] = 0.0;
    for (int r = 0; r < N; ++r) {
        m[2 * (r + r * N)] = (double)(r * (o + 1)) / N;
        if (r + 1 < N) {
            m[2 * (r + (r + 1) * N)] = o + 1.0;     m[2 * (r + (r + 1) * N) + 1] = 0.5 * (o + 1);
            m[2 * (r + 1 + r * N)] = o + 1.0;       m[2 * (r + 1 + r * N) + 1] = -0.5 * (o + 1);
        }
    }
}

static void* work(void* arg) {
    worker_t* w = (worker_t*)arg;
    oqb_ctx* ctx = w->ctx;
    double* psi = malloc(sizeof(double) * 2 * N * PER);
    double* obs = malloc(sizeof(double) * 2 * N * N * NOBS);
    void *d_psi = NULL, *d_obs = NULL, *d_out = NULL;
    w->status = 1;
    for (int b = 0; b < PER; ++b) fill_state(w->rank, b, psi + 2 * N * b);
    for (int o = 0; o < NOBS; ++o) fill_obs(o, obs + 2 * N * N * o);
    if (oqb_malloc(ctx, sizeof(double) * 2 * N * PER, &d_psi) || oqb_malloc(ctx, sizeof(double) * 2 * N * N * NOBS, &d_obs) ||
        oqb_malloc(ctx, sizeof(double) * 2 * NOBS, &d_out))
        goto done;
    if (oqb_upload(ctx, d_psi, psi, sizeof(double) * 2 * N * PER) || oqb_upload(ctx, d_obs, obs, sizeof(double) * 2 * N * N * NOBS)) goto done;
    /* partial sums and the collective are enqueued back to back on this context's stream */
    if (oqb_expect_sum(ctx, N, NOBS, d_obs, PER, d_psi, 1, (double*)d_out)) goto done;
    if (oqb_allreduce_obs(w->comm, (double*)d_out, 2 * NOBS)) goto done;
    if (oqb_sync(ctx) || oqb_download(ctx, w->result, d_out, sizeof(double) * 2 * NOBS)) goto done;
    if (ame_on_this_gpu(ctx, w->ame_sum)) goto done; /* tensor-pipe ZGEMMs, device eigensolver, term builder on THIS device */
    w->status = 0;
done:
    if (w->status) fprintf(stderr, "rank %d: %s\n", w->rank, oqb_last_error(ctx));
    if (d_psi) oqb_free(ctx, d_psi);
    if (d_obs) oqb_free(ctx, d_obs);
    if (d_out) oqb_free(ctx, d_out);
    free(psi); free(obs);
    return NULL;
}

int main(int argc, char** argv) {
    int ngpu = argc > 1 ? atoi(argv[1]) : MAXG;
    if (ngpu > MAXG) ngpu = MAXG;
    oqb_ctx* ctxs[MAXG];
    oqb_comm* comms[MAXG];
    int have = 0;
    for (; have < ngpu; ++have)
        if (oqb_create(have, &ctxs[have]) != 0) break;
    if (have == 0) {
        fprintf(stderr, "oqb_create failed: a B200 (sm_100) is required, there is no CPU fallback\n");
        return 3;
    }
    ngpu = have;
    if (oqb_comm_init_all(ngpu, ctxs, comms) != 0) {
        fprintf(stderr, "oqb_comm_init_all: %s\n", oqb_last_error(ctxs[0]));
        return 2;
    }
    int ver = 0;
    oqb_nccl_version(&ver);
    worker_t w[MAXG];
    pthread_t th[MAXG];
    for (int r = 0; r < ngpu; ++r) {
        w[r].rank = r; w[r].ngpu = ngpu; w[r].ctx = ctxs[r]; w[r].comm = comms[r]; w[r].status = 1;
        pthread_create(&th[r], NULL, work, &w[r]);
    }
    for (int r = 0; r < ngpu; ++r) pthread_join(th[r], NULL);
    /* host reference: Σ over every GPU's block */
    double want[2 * NOBS] = {0};
    double psi[2 * N], obs[2 * N * N];
    for (int o = 0; o < NOBS; ++o) {
        fill_obs(o, obs);
        for (int r = 0; r < ngpu; ++r)
            for (int b = 0; b < PER; ++b) {
                fill_state(r, b, psi);
                for (int i = 0; i < N; ++i)
                    for (int j = 0; j < N; ++j) { /* conj(ψ_i) O_ij ψ_j */
                        const double orr = obs[2 * (i + j * N)], oi = obs[2 * (i + j * N) + 1];
                        const double tr = orr * psi[2 * j] - oi * psi[2 * j + 1], ti = orr * psi[2 * j + 1] + oi * psi[2 * j];
                        want[2 * o] += psi[2 * i] * tr + psi[2 * i + 1] * ti;
                        want[2 * o + 1] += psi[2 * i] * ti - psi[2 * i + 1] * tr;
                    }
            }
    }
    int bad = 0;
    for (int r = 0; r < ngpu; ++r) {
        if (w[r].status) { bad = 1; continue; }
        for (int k = 0; k < 2 * NOBS; ++k)
            if (fabs(w[r].result[k] - want[k]) > 1e-10 * (1.0 + fabs(want[k]))) bad = 1;
    }
    for (int r = 1; r < ngpu; ++r) /* the same AME right-hand side on every GPU of this process: identical results */
        if (!w[r].status && !w[0].status && (fabs(w[r].ame_sum[1] - w[0].ame_sum[1]) > 1e-9 * fabs(w[0].ame_sum[1]) || !(w[r].ame_sum[1] > 0.0))) bad = 1;
    printf("NCCL %d, %d GPU(s): <O_0> = %.12f (host %.12f), <O_2> = %.12f (host %.12f); AME |du|^2 on GPU 0 / last: %.12e / %.12e\n", ver, ngpu,
           w[0].result[0], want[0], w[0].result[4], want[4], w[0].ame_sum[1], w[ngpu - 1].ame_sum[1]);
    for (int r = 0; r < ngpu; ++r) {
        oqb_comm_destroy(comms[r]);
        oqb_destroy(ctxs[r]);
    }
    if (bad) {
        fprintf(stderr, "MISMATCH\n");
        return 1;
    }
    printf("c_comm_demo OK\n");
    return 0;
}
