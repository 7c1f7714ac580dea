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

typedef struct {
    int rank, ngpu;
    oqb_ctx* ctx;
    oqb_comm* comm;
    double result[2 * NOBS];
    int status;
} worker_t;

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
    printf("NCCL %d, %d GPU(s): <O_0> = %.12f (host %.12f), <O_2> = %.12f (host %.12f)\n", ver, ngpu, w[0].result[0], want[0], w[0].result[4], want[4]);
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
