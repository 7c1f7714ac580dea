// The one collective of the path (BASELINE.json north_star): the ensemble-average reduce of observables over
// the GPUs of a box.  Every ρ / ψ / schedule is independent inside an RHS call, so nothing else ever crosses NVLink.
//
//   oqb_expect_sum            per-GPU partial sums Σ_b tr(O_o ρ_b) / Σ_b ψ_b'O_oψ_b in a fixed order (deterministic), on
//                             the context stream
//   oqb_allreduce_obs         ncclAllReduce(sum, double) of those few numbers, enqueued on the SAME stream — no host
//                             synchronisation between the observable kernel and the collective
//   oqb_allgather_obs         ncclAllGather for per-schedule observables (C5)
//
// NCCL is bound at run time with dlopen("libnccl.so.2"): liboqb_b200.so links against nothing but the CUDA runtime, and a
// process that never creates a communicator never loads NCCL.  Process models: one process (or thread) per GPU with
// oqb_comm_unique_id + oqb_comm_init_rank (the id travels by whatever the host uses: MPI, torch.distributed, a file), or
// one process driving all GPUs with oqb_comm_init_all (ncclCommInitAll) and group calls.
#include <dlfcn.h>

#include <cstring>
#include <vector>

#include "../../include/oqb.h"
#include "common.cuh"

using namespace oqb;

namespace {

// minimal NCCL 2.x ABI (stable since 2.0): opaque communicator, 128-byte unique id, enum values of nccl.h
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSum = 0 };
enum { ncclFloat64 = 8 };

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    bool ok = false;
};

NcclApi& nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            a.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (a.lib) break;
        }
        if (!a.lib) return a;
#define OQB_SYM(field, name) a.field = (decltype(a.field))dlsym(a.lib, name)
        OQB_SYM(GetUniqueId, "ncclGetUniqueId");
        OQB_SYM(CommInitRank, "ncclCommInitRank");
        OQB_SYM(CommInitAll, "ncclCommInitAll");
        OQB_SYM(CommDestroy, "ncclCommDestroy");
        OQB_SYM(AllReduce, "ncclAllReduce");
        OQB_SYM(AllGather, "ncclAllGather");
        OQB_SYM(GroupStart, "ncclGroupStart");
        OQB_SYM(GroupEnd, "ncclGroupEnd");
        OQB_SYM(GetErrorString, "ncclGetErrorString");
        OQB_SYM(GetVersion, "ncclGetVersion");
#undef OQB_SYM
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommInitAll && a.CommDestroy && a.AllReduce && a.AllGather && a.GroupStart &&
               a.GroupEnd && a.GetErrorString;
        return a;
    }();
    return api;
}

// out[o] = Σ_b in[b·nobs + o]  (complex) — one block per observable, fixed summation order
__global__ void sum_members_kernel(int nb, int nobs, const c128* __restrict__ in, c128* __restrict__ out) {
    const int o = blockIdx.x;
    double sr = 0.0, si = 0.0;
    // thread t owns the contiguous member range [t·per, (t+1)·per): the order does not depend on the launch
    const int per = (nb + blockDim.x - 1) / blockDim.x;
    const int b0 = threadIdx.x * per, b1 = min(nb, b0 + per);
    for (int b = b0; b < b1; ++b) {
        const c128 v = in[(size_t)b * nobs + o];
        sr += v.x;
        si += v.y;
    }
    __shared__ double red[2][256];
    red[0][threadIdx.x] = sr;
    red[1][threadIdx.x] = si;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            red[0][threadIdx.x] += red[0][threadIdx.x + s];
            red[1][threadIdx.x] += red[1][threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[o] = make_double2(red[0][0], red[1][0]);
}

}  // namespace

struct oqb_comm {
    oqb_ctx* ctx = nullptr;
    ncclComm_t comm = nullptr;
    int nranks = 0, rank = 0;
};

extern "C" {

int oqb_expect_sum(oqb_ctx* c, int n, int nobs, const void* d_obs, int nb, const void* d_state, int is_vector, double* d_out) {
    if (!c) return OQB_EINVAL;
    OQB_DEVICE(c, c);
    if (nobs <= 0 || nb <= 0) return 0;
    if (!d_obs || !d_state || !d_out) return set_error(c, OQB_EINVAL, "oqb_expect_sum: null argument");
    OQB_TRY(ws_reserve(c, (size_t)nb * nobs * sizeof(c128)));
    c128* per = (c128*)c->ws;
    OQB_TRY(expect(c, n, nobs, (const c128*)d_obs, nb, (const c128*)d_state, is_vector, per));
    sum_members_kernel<<<nobs, 256, 0, c->stream>>>(nb, nobs, per, (c128*)d_out);
    OQB_LAUNCH_CHECK(c);
    return 0;
}

int oqb_nccl_version(int* version) {
    NcclApi& api = nccl_api();
    if (!api.ok || !api.GetVersion || !version) return OQB_ECUDA;
    return api.GetVersion(version) == 0 ? 0 : OQB_ECUDA;
}

int oqb_comm_unique_id(void* id128) {
    NcclApi& api = nccl_api();
    if (!api.ok || !id128) return OQB_ECUDA;
    ncclUniqueId id;
    if (api.GetUniqueId(&id) != 0) return OQB_ECUDA;
    memcpy(id128, &id, sizeof(id));
    return 0;
}

int oqb_comm_init_rank(oqb_ctx* c, int nranks, int rank, const void* id128, oqb_comm** out) {
    if (!c || !out) return OQB_EINVAL;
    OQB_DEVICE(c, c);
    *out = nullptr;
    NcclApi& api = nccl_api();
    if (!api.ok) return set_error(c, OQB_ECUDA, "oqb_comm_init_rank: libnccl.so.2 could not be loaded (%s)", dlerror());
    if (!id128 || nranks <= 0 || rank < 0 || rank >= nranks) return set_error(c, OQB_EINVAL, "oqb_comm_init_rank: bad arguments");
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    const ncclResult_t r = api.CommInitRank(&comm, nranks, id, rank);
    if (r != 0) return set_error(c, OQB_ECUDA, "ncclCommInitRank failed: %s", api.GetErrorString(r));
    oqb_comm* m = new oqb_comm();
    m->ctx = c; m->comm = comm; m->nranks = nranks; m->rank = rank;
    *out = m;
    return 0;
}

int oqb_comm_init_all(int ndev, oqb_ctx* const* ctxs, oqb_comm** out) {
    if (ndev <= 0 || !ctxs || !out) return OQB_EINVAL;
    NcclApi& api = nccl_api();
    oqb_ctx* c0 = ctxs[0];
    if (!api.ok) return set_error(c0, OQB_ECUDA, "oqb_comm_init_all: libnccl.so.2 could not be loaded (%s)", dlerror());
    std::vector<int> devs(ndev);
    for (int i = 0; i < ndev; ++i) {
        if (!ctxs[i]) return set_error(c0, OQB_EINVAL, "oqb_comm_init_all: null context");
        devs[i] = ctxs[i]->device;
    }
    std::vector<ncclComm_t> comms(ndev, nullptr);
    const ncclResult_t r = api.CommInitAll(comms.data(), ndev, devs.data());
    if (r != 0) return set_error(c0, OQB_ECUDA, "ncclCommInitAll failed: %s", api.GetErrorString(r));
    for (int i = 0; i < ndev; ++i) {
        oqb_comm* m = new oqb_comm();
        m->ctx = ctxs[i]; m->comm = comms[i]; m->nranks = ndev; m->rank = i;
        out[i] = m;
    }
    return 0;
}

int oqb_comm_destroy(oqb_comm* m) {
    if (!m) return 0;
    OQB_DEVICE(m, m->ctx);
    cudaStreamSynchronize(m->ctx->stream);
    NcclApi& api = nccl_api();
    if (api.ok && m->comm) api.CommDestroy(m->comm);
    delete m;
    return 0;
}

int oqb_comm_rank(const oqb_comm* m, int* rank, int* nranks) {
    if (!m) return OQB_EINVAL;
    if (rank) *rank = m->rank;
    if (nranks) *nranks = m->nranks;
    return 0;
}

int oqb_comm_group_start(void) { return nccl_api().ok && nccl_api().GroupStart() == 0 ? 0 : OQB_ECUDA; }
int oqb_comm_group_end(void) { return nccl_api().ok && nccl_api().GroupEnd() == 0 ? 0 : OQB_ECUDA; }

int oqb_allreduce_obs(oqb_comm* m, double* d_buf, int64_t count) {
    if (!m) return OQB_EINVAL;
    OQB_DEVICE(m, m->ctx);
    oqb_ctx* c = m->ctx;
    if (count <= 0) return 0;
    if (!d_buf) return set_error(c, OQB_EINVAL, "oqb_allreduce_obs: null buffer");
    NcclApi& api = nccl_api();
    const ncclResult_t r = api.AllReduce(d_buf, d_buf, (size_t)count, ncclFloat64, ncclSum, m->comm, c->stream);
    if (r != 0) return set_error(c, OQB_ECUDA, "ncclAllReduce failed: %s", api.GetErrorString(r));
    return 0;
}

int oqb_allgather_obs(oqb_comm* m, const double* d_send, double* d_recv, int64_t count_per_rank) {
    if (!m) return OQB_EINVAL;
    OQB_DEVICE(m, m->ctx);
    oqb_ctx* c = m->ctx;
    if (count_per_rank <= 0) return 0;
    if (!d_send || !d_recv) return set_error(c, OQB_EINVAL, "oqb_allgather_obs: null buffer");
    NcclApi& api = nccl_api();
    const ncclResult_t r = api.AllGather(d_send, d_recv, (size_t)count_per_rank, ncclFloat64, m->comm, c->stream);
    if (r != 0) return set_error(c, OQB_ECUDA, "ncclAllGather failed: %s", api.GetErrorString(r));
    return 0;
}

}  // extern "C"
