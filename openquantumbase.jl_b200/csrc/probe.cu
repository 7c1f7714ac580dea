// FP64 issue-throughput probes: what the DMMA (tensor) and DFMA (vector) pipes of this device
// sustain when nothing else is in the way.  bench.py reports them next to the cuBLAS DGEMM
// figure so that the FP64 roofline denominator is a measured number (MEASURED_PEAKS.json only
// carries bf16 and HBM).
#include "../../include/oqb.h"
#include "common.cuh"

using namespace oqb;

namespace {

constexpr int kChains = 16;

__global__ void __launch_bounds__(256) probe_dmma_kernel(double* out, int iters) {
    double c0[kChains], c1[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) { c0[i] = 0.0; c1[i] = 0.0; }
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[i]), "+d"(c1[i])
                         : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;  // keep the chain alive
}

__global__ void __launch_bounds__(256) probe_dfma_kernel(double* out, int iters) {
    double c[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) c[i] = 1e-3 * i;
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

// Shared-memory gather ceiling: every thread of a 512-thread CTA (3 CTAs per SM, 64 KB of staged ψ each — the shape of the
// generic trajectory kernel) issues LDS.128 gathers from its CTA's 4096-entry vector, 8 independent loads in flight, and
// nothing else of substance.  pattern 0: XOR-permuted indices (tid ^ mask: conflict-free, what Pauli-structured operators
// give); pattern 1: pseudo-random indices (what an unstructured sparse operator gives: 8-thread phase conflicts).
__global__ void __launch_bounds__(512) probe_gather_kernel(double* out, int iters, int pattern) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* x = reinterpret_cast<double2*>(smem_raw);
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) x[i] = make_double2(1e-3 * i, -1e-3 * i);
    __syncthreads();
    unsigned st = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    double ar = 0.0, ai = 0.0;
    for (int it = 0; it < iters; ++it) {
        double2 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            unsigned idx;
            if (pattern == 0) {
                idx = (threadIdx.x + 512u * (unsigned)(q & 7)) ^ (1u << ((it + q) % 12));
            } else {
                st = st * 1664525u + 1013904223u;
                idx = st >> 20;
            }
            v[q] = x[idx & 4095u];
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) { ar += v[q].x; ai += v[q].y; }
    }
    if (ar == 123.456 && ai == 654.321) out[0] = ar;
}

}  // namespace

extern "C" int oqb_probe_smem_gather(oqb_ctx* c, int pattern, double* gathers_per_sec) {
    OQB_DEVICE(c, c);
    if (!c || !gathers_per_sec) return OQB_EINVAL;
    int sms = 0;
    OQB_CUDA(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    static DevOnce cfg;
    if (!cfg.done(c->device)) {
        OQB_CUDA(c, cudaFuncSetAttribute(probe_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        cfg.set(c->device);
    }
    double* d_out = nullptr;
    OQB_CUDA(c, cudaMalloc((void**)&d_out, 64));
    cudaEvent_t e0, e1;
    OQB_CUDA(c, cudaEventCreate(&e0));
    OQB_CUDA(c, cudaEventCreate(&e1));
    const int grid = sms * 3, block = 512, iters = 2048;
    float ms = 0.f;
    probe_gather_kernel<<<grid, block, 64 * 1024, c->stream>>>(d_out, 16, pattern);
    c->launches++;
    OQB_CUDA(c, cudaEventRecord(e0, c->stream));
    probe_gather_kernel<<<grid, block, 64 * 1024, c->stream>>>(d_out, iters, pattern);
    OQB_LAUNCH_CHECK(c);
    OQB_CUDA(c, cudaEventRecord(e1, c->stream));
    OQB_CUDA(c, cudaEventSynchronize(e1));
    OQB_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
    *gathers_per_sec = (double)grid * block * iters * 8.0 / (ms * 1e-3);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return 0;
}

extern "C" int oqb_probe_fp64(oqb_ctx* c, double* dmma_tflops, double* dfma_tflops) {
    OQB_DEVICE(c, c);
    int sms = 0;
    OQB_CUDA(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    double* d_out = nullptr;
    OQB_CUDA(c, cudaMalloc((void**)&d_out, 64));
    cudaEvent_t e0, e1;
    OQB_CUDA(c, cudaEventCreate(&e0));
    OQB_CUDA(c, cudaEventCreate(&e1));
    const int grid = sms * 8, block = 256, iters = 4096;
    float ms = 0.f;
    // warm-up + timed, DMMA
    probe_dmma_kernel<<<grid, block, 0, c->stream>>>(d_out, 64);
    c->launches++;
    OQB_CUDA(c, cudaEventRecord(e0, c->stream));
    probe_dmma_kernel<<<grid, block, 0, c->stream>>>(d_out, iters);
    OQB_LAUNCH_CHECK(c);
    OQB_CUDA(c, cudaEventRecord(e1, c->stream));
    OQB_CUDA(c, cudaEventSynchronize(e1));
    OQB_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
    if (dmma_tflops) *dmma_tflops = (double)grid * (block / 32) * iters * kChains * 512.0 / (ms * 1e-3) / 1e12;
    probe_dfma_kernel<<<grid, block, 0, c->stream>>>(d_out, 64);
    c->launches++;
    OQB_CUDA(c, cudaEventRecord(e0, c->stream));
    probe_dfma_kernel<<<grid, block, 0, c->stream>>>(d_out, iters);
    OQB_LAUNCH_CHECK(c);
    OQB_CUDA(c, cudaEventRecord(e1, c->stream));
    OQB_CUDA(c, cudaEventSynchronize(e1));
    OQB_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
    if (dfma_tflops) *dfma_tflops = (double)grid * block * iters * kChains * 2.0 / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return 0;
}
