// Context, memory and the exposed ZGEMM entry point of the C ABI (include/oqb.h).
#include <cstdarg>
#include <cstring>

#include "../../include/oqb.h"
#include <cstdlib>

#include "common.cuh"

namespace oqb {

int set_error(Ctx* c, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

int ws_reserve(Ctx* c, size_t bytes) {
    if (bytes <= c->ws_bytes) return 0;
    // the old block may still be in use by queued kernels
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->ws) cudaFree(c->ws);
    c->ws = nullptr;
    c->ws_bytes = 0;
    size_t want = bytes + (bytes >> 3);
    cudaError_t e = cudaMalloc(&c->ws, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = cudaMalloc(&c->ws, want);
    }
    if (e != cudaSuccess) return set_error(c, OQB_ENOMEM, "workspace allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    c->ws_bytes = want;
    return 0;
}

int io_reserve(Ctx* c, size_t bytes) {
    if (bytes <= c->io_bytes) return 0;
    OQB_CUDA(c, cudaDeviceSynchronize());
    if (c->io) cudaFree(c->io);
    c->io = nullptr;
    c->io_bytes = 0;
    cudaError_t e = cudaMalloc(&c->io, bytes);
    if (e != cudaSuccess) return set_error(c, OQB_ENOMEM, "staging allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    c->io_bytes = bytes;
    return 0;
}

// Stage a small host array through pinned memory and copy it to the device asynchronously on the
// context stream.  The staging area is a ring of slots so that back-to-back calls do not overwrite
// data a still-queued copy reads.
int coef_upload(Ctx* c, const void* host, size_t bytes, void** dptr) {
    const size_t align = 256;
    size_t need = (bytes + align - 1) / align * align;
    if (need * 4 > c->pin_bytes) {
        OQB_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->pin) cudaFreeHost(c->pin);
        if (c->dcoef) cudaFree(c->dcoef);
        c->pin = c->dcoef = nullptr;
        size_t cap = need * 8 < (1u << 20) ? (1u << 20) : need * 8;
        OQB_CUDA(c, cudaMallocHost(&c->pin, cap));
        OQB_CUDA(c, cudaMalloc(&c->dcoef, cap));
        c->pin_bytes = c->dcoef_bytes = cap;
        c->launches += 0;
        c->coef_head = 0;
    }
    if (c->coef_head + need > c->pin_bytes) {
        // wrap: make sure everything queued so far has been consumed
        OQB_CUDA(c, cudaStreamSynchronize(c->stream));
        c->coef_head = 0;
    }
    char* hp = (char*)c->pin + c->coef_head;
    char* dp = (char*)c->dcoef + c->coef_head;
    memcpy(hp, host, bytes);
    OQB_CUDA(c, cudaMemcpyAsync(dp, hp, bytes, cudaMemcpyHostToDevice, c->stream));
    c->coef_head += need;
    *dptr = dp;
    return 0;
}

}  // namespace oqb

using namespace oqb;

extern "C" {

int oqb_version(void) { return OQB_VERSION; }

int oqb_create(int device, oqb_ctx** out) {
    if (!out) return OQB_EINVAL;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return OQB_ECUDA;  // no CPU fallback: fail loudly
    if (device < 0 || device >= ndev) return OQB_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return OQB_ECUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return OQB_ECUDA;
    if (prop.major != 10) return OQB_ECUDA;  // sm_100a code only
    {  // the stream-ordered allocator keeps freed blocks instead of returning them to the driver at every synchronisation
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    oqb_ctx* c = new oqb_ctx();
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return OQB_ECUDA; }
    c->own_stream = true;
    for (int i = 0; i < 2; ++i) cudaStreamCreateWithFlags(&c->copy_stream[i], cudaStreamNonBlocking);
    for (int i = 0; i < 16; ++i) cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming);
    *out = c;
    return OQB_OK;
}

int oqb_destroy(oqb_ctx* c) {
    if (!c) return OQB_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->cusolver && c->cusolver_destroy) c->cusolver_destroy(c->cusolver);
    if (c->ws) cudaFree(c->ws);
    if (c->io) cudaFree(c->io);
    if (c->pin) cudaFreeHost(c->pin);
    if (c->dcoef) cudaFree(c->dcoef);
    for (int i = 0; i < 2; ++i) if (c->copy_stream[i]) cudaStreamDestroy(c->copy_stream[i]);
    for (int i = 0; i < 16; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return OQB_OK;
}

const char* oqb_last_error(const oqb_ctx* c) { return c ? c->err.c_str() : "null context"; }

int oqb_sync(oqb_ctx* c) {
    OQB_DEVICE(c, c);
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int oqb_set_stream(oqb_ctx* c, void* s) {
    OQB_DEVICE(c, c);
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->own_stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)s;
    c->own_stream = false;
    return 0;
}
int oqb_get_stream(oqb_ctx* c, void** s) { *s = (void*)c->stream; return 0; }
int oqb_launch_count(const oqb_ctx* c, uint64_t* n) { *n = c->launches; return 0; }
int oqb_axpy_batched(oqb_ctx* c, int64_t len, int nb, const double* alpha, const void* d_x, int64_t stride_x, void* d_y) {
    OQB_DEVICE(c, c);
    if (!alpha || !d_x || !d_y) return set_error(c, OQB_EINVAL, "oqb_axpy_batched: null argument");
    return axpy_batched(c, len, nb, make_double2(alpha[0], alpha[1]), (const c128*)d_x, stride_x, (c128*)d_y);
}
namespace {
struct OptEntry { const char* name; int Options::* field; };
const OptEntry kOpts[] = {
    {"zgemm_tma", &Options::zgemm_tma}, {"zgemm_bm", &Options::zgemm_bm}, {"tma_oneshot", &Options::tma_oneshot},
    {"host_slots", &Options::host_slots}, {"lindblad_diag", &Options::lindblad_diag}, {"lindblad_fuse", &Options::lindblad_fuse}, {"lindblad_mono", &Options::lindblad_mono}, {"coupling_mono", &Options::coupling_mono}, {"lindblad_split", &Options::lindblad_split}, {"lindblad_stencil", &Options::lindblad_stencil}, {"coupling_diag", &Options::coupling_diag}, {"traj_kernel", &Options::traj_kernel},
    {"traj_ell", &Options::traj_ell}, {"sell_r", &Options::sell_r}, {"sell_unit", &Options::sell_unit}, {"rk_fuse", &Options::rk_fuse}, {"xor_fixed", &Options::xor_fixed}, {"xor_rows", &Options::xor_rows},
    {"ame_host_chunk_mib", &Options::ame_host_chunk_mib}, {"davies_dense_min_pairs", &Options::davies_dense_min_pairs},
    {"debug_checks", &Options::debug_checks}, {"fuse_cross", &Options::fuse_cross},
};
}  // namespace

int oqb_set_option(oqb_ctx* c, const char* name, int64_t value) {
    if (!c || !name) return OQB_EINVAL;
    if (!strcmp(name, "zgemm_3m")) return oqb_set_zgemm_mode(c, (int)value);
    if (!strcmp(name, "real_operand")) return oqb_set_real_operand_mode(c, (int)value);
    if (!strcmp(name, "davies_max_cross")) { c->opt.davies_max_cross = value; return 0; }
    for (const OptEntry& e : kOpts)
        if (!strcmp(name, e.name)) { c->opt.*(e.field) = (int)value; return 0; }
    return set_error(c, OQB_EINVAL, "oqb_set_option: unknown option '%s'", name);
}

int oqb_get_option(const oqb_ctx* c, const char* name, int64_t* value) {
    if (!c || !name || !value) return OQB_EINVAL;
    if (!strcmp(name, "zgemm_3m")) { *value = c->zgemm_3m; return 0; }
    if (!strcmp(name, "real_operand")) { *value = c->real_operand; return 0; }
    if (!strcmp(name, "davies_max_cross")) { *value = c->opt.davies_max_cross; return 0; }
    for (const OptEntry& e : kOpts)
        if (!strcmp(name, e.name)) { *value = c->opt.*(e.field); return 0; }
    return OQB_EINVAL;
}

int oqb_set_zgemm_mode(oqb_ctx* c, int mode) {
    OQB_DEVICE(c, c);
    if (mode != 0 && mode != 1) return set_error(c, OQB_EINVAL, "oqb_set_zgemm_mode: mode must be 0 (4M) or 1 (3M)");
    c->zgemm_3m = mode;
    return 0;
}
int oqb_get_zgemm_mode(const oqb_ctx* c, int* mode) { *mode = c->zgemm_3m; return 0; }
int oqb_set_real_operand_mode(oqb_ctx* c, int on) {
    OQB_DEVICE(c, c);
    c->real_operand = on != 0;
    return 0;
}
int oqb_get_real_operand_mode(const oqb_ctx* c, int* on) { *on = c->real_operand; return 0; }

int oqb_malloc(oqb_ctx* c, size_t nbytes, void** d) {
    OQB_DEVICE(c, c);
    cudaError_t e = cudaMalloc(d, nbytes ? nbytes : 16);
    if (e != cudaSuccess) return set_error(c, OQB_ENOMEM, "cudaMalloc(%zu) failed: %s", nbytes, cudaGetErrorString(e));
    return 0;
}
int oqb_free(oqb_ctx* c, void* d) {
    OQB_DEVICE(c, c);
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    OQB_CUDA(c, cudaFree(d));
    return 0;
}
int oqb_malloc_host(oqb_ctx* c, size_t nbytes, void** h) {
    OQB_DEVICE(c, c);
    cudaError_t e = cudaMallocHost(h, nbytes ? nbytes : 16);
    if (e != cudaSuccess) return set_error(c, OQB_ENOMEM, "cudaMallocHost(%zu) failed: %s", nbytes, cudaGetErrorString(e));
    return 0;
}
int oqb_free_host(oqb_ctx* c, void* h) { OQB_CUDA(c, cudaFreeHost(h)); return 0; }
int oqb_upload(oqb_ctx* c, void* d, const void* h, size_t n) {
    OQB_DEVICE(c, c);
    OQB_CUDA(c, cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, c->stream));
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int oqb_download(oqb_ctx* c, void* h, const void* d, size_t n) {
    OQB_DEVICE(c, c);
    OQB_CUDA(c, cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, c->stream));
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int oqb_zgemm_batched(oqb_ctx* c, int opA, int opB, int M, int N, int K, const double* alpha, const void* dA,
                      int64_t lda, int64_t strideA, const void* dB, int64_t ldb, int64_t strideB,
                      const double* beta, void* dC, int64_t ldc, int64_t strideC, int batch) {
    OQB_DEVICE(c, c);
    if (opA < 0 || opA > 2 || opB < 0 || opB > 2) return set_error(c, OQB_EINVAL, "zgemm: bad op");
    ZgemmParams p;
    p.M = M; p.N = N; p.K = K; p.batch = batch;
    p.opA = opA; p.opB = opB;
    p.A = (const c128*)dA; p.lda = lda; p.strideA = strideA;
    p.B = (const c128*)dB; p.ldb = ldb; p.strideB = strideB;
    p.C = (c128*)dC; p.ldc = ldc; p.strideC = strideC;
    p.alpha = make_double2(alpha[0], alpha[1]);
    p.beta = make_double2(beta[0], beta[1]);
    return zgemm(c, p);
}

int oqb_profile_begin(oqb_ctx* c) {
    OQB_DEVICE(c, c);
    for (cudaEvent_t e : c->prof_events) cudaEventDestroy(e);
    c->prof_events.clear();
    c->prof_flops = 0.0;
    c->profiling = true;
    return 0;
}

int oqb_profile_end(oqb_ctx* c, double* zgemm_ms, uint64_t* zgemm_launches, double* zgemm_flops) {
    OQB_DEVICE(c, c);
    c->profiling = false;
    OQB_CUDA(c, cudaStreamSynchronize(c->stream));
    double total = 0.0;
    for (size_t i = 0; i + 1 < c->prof_events.size(); i += 2) {
        float ms = 0.f;
        OQB_CUDA(c, cudaEventElapsedTime(&ms, c->prof_events[i], c->prof_events[i + 1]));
        total += ms;
    }
    if (zgemm_ms) *zgemm_ms = total;
    if (zgemm_launches) *zgemm_launches = c->prof_events.size() / 2;
    if (zgemm_flops) *zgemm_flops = c->prof_flops;
    for (cudaEvent_t e : c->prof_events) cudaEventDestroy(e);
    c->prof_events.clear();
    return 0;
}

}  // extern "C"
