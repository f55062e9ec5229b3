// Data-parallel exchange of the LoRA-gradient buckets: NCCL all-reduce(mean) over NVLink 5 / NVSwitch behind the C ABI
// (b2q_comm_*, include/b2q.h).  Stands in for the DDP reducer that `accel.prepare` installs in the reference
// (/root/reference/trainer/utils_trainer.py:32-37) for callers that do not go through torch.distributed.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2): libb2q.so has no link-time dependency on it, a single-GPU user
// never needs it, and inside a torch process the SONAME lookup returns the copy torch has already loaded, so both share
// one NCCL.  Only the handful of entry points below is used; their types are restated here (plain C ABI: opaque comm
// pointer, 128-byte unique id passed by value, int-sized enums) so that no NCCL header is needed to build.
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <mutex>

#include "b2q_internal.h"

namespace {

typedef struct ncclComm* nccl_comm_t;
struct nccl_unique_id { char internal[128]; };                 // NCCL_UNIQUE_ID_BYTES
constexpr int kNcclSuccess = 0;
constexpr int kNcclFloat32 = 7, kNcclBfloat16 = 9;              // ncclDataType_t
constexpr int kNcclAvg = 4;                                     // ncclRedOp_t

struct NcclApi {
    void* handle = nullptr;
    int (*get_version)(int*) = nullptr;
    int (*get_unique_id)(nccl_unique_id*) = nullptr;
    int (*comm_init_rank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
    int (*all_reduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*comm_destroy)(nccl_comm_t) = nullptr;
    const char* (*get_error_string)(int) = nullptr;
    bool ok = false;
};

NcclApi g_api;
std::once_flag g_api_once;

template <class Fn>
bool resolve(void* h, const char* name, Fn& fn) {
    fn = reinterpret_cast<Fn>(dlsym(h, name));
    return fn != nullptr;
}

const NcclApi& nccl() {
    std::call_once(g_api_once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            g_api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (g_api.handle != nullptr) break;
        }
        if (g_api.handle == nullptr) {
            b2q::set_error_detail("dlopen(libnccl.so.2) failed");
            return;
        }
        void* h = g_api.handle;
        g_api.ok = resolve(h, "ncclGetVersion", g_api.get_version) && resolve(h, "ncclGetUniqueId", g_api.get_unique_id) &&
                   resolve(h, "ncclCommInitRank", g_api.comm_init_rank) && resolve(h, "ncclAllReduce", g_api.all_reduce) &&
                   resolve(h, "ncclCommDestroy", g_api.comm_destroy) &&
                   resolve(h, "ncclGetErrorString", g_api.get_error_string);
        if (!g_api.ok) b2q::set_error_detail("libnccl.so.2 lacks one of the required entry points");
    });
    return g_api;
}

int nccl_fail(const char* what, int res) {
    char msg[256];
    const NcclApi& api = nccl();
    snprintf(msg, sizeof(msg), "%s -> ncclResult %d (%s)", what, res,
             api.get_error_string != nullptr ? api.get_error_string(res) : "?");
    b2q::set_error_detail(msg);
    return B2Q_ERR_COMM;
}

}  // namespace

struct b2q_comm {
    nccl_comm_t comm;
    cudaStream_t stream;   // the communicator's own stream (overlapped mode)
    cudaEvent_t ready;     // producer stream -> comm stream
    cudaEvent_t done;      // comm stream -> consumer stream
    int nranks, rank, device;
    bool pending;          // an overlapped all-reduce has been issued since the last b2q_comm_wait
};

extern "C" int b2q_comm_nccl_version(void) {
    const NcclApi& api = nccl();
    if (!api.ok) return B2Q_ERR_COMM;
    int v = 0;
    const int r = api.get_version(&v);
    return r == kNcclSuccess ? v : nccl_fail("ncclGetVersion", r);
}

extern "C" int b2q_comm_unique_id(void* id_out, size_t id_bytes) {
    if (id_out == nullptr || id_bytes < B2Q_COMM_ID_BYTES) return B2Q_ERR_ARG;
    const NcclApi& api = nccl();
    if (!api.ok) return B2Q_ERR_COMM;
    nccl_unique_id id;
    const int r = api.get_unique_id(&id);
    if (r != kNcclSuccess) return nccl_fail("ncclGetUniqueId", r);
    memcpy(id_out, id.internal, sizeof(id.internal));
    return 0;
}

extern "C" int b2q_comm_init(b2q_comm** comm_out, const void* id, size_t id_bytes, int nranks, int rank) {
    if (comm_out == nullptr || id == nullptr || id_bytes < B2Q_COMM_ID_BYTES) return B2Q_ERR_ARG;
    if (nranks < 1 || rank < 0 || rank >= nranks) return B2Q_ERR_ARG;
    const NcclApi& api = nccl();
    if (!api.ok) return B2Q_ERR_COMM;
    b2q_comm* c = new b2q_comm();
    memset(c, 0, sizeof(*c));
    c->nranks = nranks;
    c->rank = rank;
    cudaError_t e = cudaGetDevice(&c->device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming);
    if (e != cudaSuccess) {   // release whatever was created before the failing call
        if (c->done != nullptr) cudaEventDestroy(c->done);
        if (c->ready != nullptr) cudaEventDestroy(c->ready);
        if (c->stream != nullptr) cudaStreamDestroy(c->stream);
        delete c;
        return static_cast<int>(e);
    }
    nccl_unique_id uid;
    memcpy(uid.internal, id, sizeof(uid.internal));
    const int r = api.comm_init_rank(&c->comm, nranks, uid, rank);   // collective over all ranks
    if (r != kNcclSuccess) {
        cudaEventDestroy(c->ready);
        cudaEventDestroy(c->done);
        cudaStreamDestroy(c->stream);
        delete c;
        return nccl_fail("ncclCommInitRank", r);
    }
    *comm_out = c;
    return 0;
}

extern "C" int b2q_comm_allreduce_bucket(b2q_comm* c, void* bucket, int64_t count, int dtype, int overlap,
                                         cudaStream_t stream) {
    if (c == nullptr || bucket == nullptr || count < 0 || (dtype != B2Q_DTYPE_BF16 && dtype != B2Q_DTYPE_F32))
        return B2Q_ERR_ARG;
    if (count == 0) return 0;
    const NcclApi& api = nccl();
    if (!api.ok) return B2Q_ERR_COMM;
    const int dt = dtype == B2Q_DTYPE_BF16 ? kNcclBfloat16 : kNcclFloat32;
    cudaStream_t on = stream;
    if (overlap) {
        // after everything enqueued on `stream` so far (the kernels that wrote this bucket), on the communicator's stream
        cudaError_t e = cudaEventRecord(c->ready, stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c->stream, c->ready, 0);
        if (e != cudaSuccess) return static_cast<int>(e);
        on = c->stream;
    }
    const int r = api.all_reduce(bucket, bucket, static_cast<size_t>(count), dt, kNcclAvg, c->comm, on);
    if (r != kNcclSuccess) return nccl_fail("ncclAllReduce", r);
    if (overlap) c->pending = true;
    return 0;
}

extern "C" int b2q_comm_wait(b2q_comm* c, cudaStream_t stream) {
    if (c == nullptr) return B2Q_ERR_ARG;
    if (!c->pending) return 0;
    cudaError_t e = cudaEventRecord(c->done, c->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, c->done, 0);
    if (e != cudaSuccess) return static_cast<int>(e);
    c->pending = false;
    return 0;
}

extern "C" int b2q_comm_destroy(b2q_comm* c) {
    if (c == nullptr) return 0;
    const NcclApi& api = nccl();
    int rc = 0;
    if (api.ok && c->comm != nullptr) {
        cudaStreamSynchronize(c->stream);
        const int r = api.comm_destroy(c->comm);
        if (r != kNcclSuccess) rc = nccl_fail("ncclCommDestroy", r);
    }
    cudaEventDestroy(c->ready);
    cudaEventDestroy(c->done);
    cudaStreamDestroy(c->stream);
    delete c;
    return rc;
}
