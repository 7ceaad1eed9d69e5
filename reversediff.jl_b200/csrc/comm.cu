// Multi-GPU result assembly: one process per GPU, every rank replays its own tape copy over its shard of
// jacobian! seed rows / hessian! columns (no data-path collective), and the ONLY exchange is the gather of the result blocks into
// the reference's result layout - a column-major (length(output) x length(input)) matrix, `result_matrix = reshape(result, ...)`
// in src/api/utils.jl:44-58, so a rank's row block is strided by m in the destination.  NCCL over NVLink / NVSwitch moves the
// blocks (grouped broadcasts, one per rank, so uneven partitions need no padding); row blocks land in a staging buffer and a 2-D
// copy kernel places them (both sides coalesced along the rows); column blocks land in place.
//
// NCCL is reached through dlopen("libnccl.so.2") - in a torch process that is the copy torch already loaded, in a Julia process
// the system library - so librdb200.so itself links only libcudart.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "kernels.h"
#include "rdb200.h"

using namespace rdb;

extern "C" const char *rdb_last_error(void);
int rdb_fail(int code, const char *fmt, ...);                         // engine.cu: sets the thread-local message
cudaStream_t rdb_ctx_stream(rdb_ctx *);                               // engine.cu
int rdb_ctx_device(rdb_ctx *);

namespace {
struct ncclUniqueId_ { char internal[128]; };
typedef struct ncclComm *ncclComm_t_;
struct Nccl {
    int (*GetUniqueId)(ncclUniqueId_ *);
    int (*CommInitRank)(ncclComm_t_ *, int, ncclUniqueId_, int);
    int (*CommDestroy)(ncclComm_t_);
    int (*GroupStart)();
    int (*GroupEnd)();
    int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t_, cudaStream_t);
    int (*Send)(const void *, size_t, int, int, ncclComm_t_, cudaStream_t);
    int (*Recv)(void *, size_t, int, int, ncclComm_t_, cudaStream_t);
    const char *(*GetErrorString)(int);
    bool ok = false;
};
Nccl &nccl() {
    static Nccl n;
    static bool tried = false;
    if (tried) return n;
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) { h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) return n;
#define RDB_NCCL(field, sym) n.field = (decltype(n.field))dlsym(h, sym); if (!n.field) return n;
    RDB_NCCL(GetUniqueId, "ncclGetUniqueId")
    RDB_NCCL(CommInitRank, "ncclCommInitRank")
    RDB_NCCL(CommDestroy, "ncclCommDestroy")
    RDB_NCCL(GroupStart, "ncclGroupStart")
    RDB_NCCL(GroupEnd, "ncclGroupEnd")
    RDB_NCCL(Broadcast, "ncclBroadcast")
    RDB_NCCL(Send, "ncclSend")
    RDB_NCCL(Recv, "ncclRecv")
    RDB_NCCL(GetErrorString, "ncclGetErrorString")
#undef RDB_NCCL
    n.ok = true;
    return n;
}
constexpr int kNcclUint8 = 1;
}  // namespace

struct rdb_comm {
    rdb_ctx *ctx = nullptr;
    ncclComm_t_ comm = nullptr;
    int world = 1, rank = 0;
    void *stage = nullptr;
    size_t stage_bytes = 0;
};

#define NC(call)                                                                                          \
    do {                                                                                                  \
        int r__ = (call);                                                                                 \
        if (r__ != 0) return rdb_fail(RDB_ECUDA, "%s failed: %s", #call, nccl().GetErrorString(r__));     \
    } while (0)
#define CUC(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) { cudaGetLastError(); return rdb_fail(RDB_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); } \
    } while (0)

// result (+)= block placed at its partition offset; host memory, plain loops.  Shared by the gloo (CPU) test of the N > 1 host
// logic and by callers that gathered the blocks with a transport of their own (MPI.jl, torch.distributed).
template <typename T>
static void place_host(int kind, const T *block, int64_t b, int64_t e, int64_t m, int64_t n, T *result) {
    if (kind == 0) {                       // rows [b, e) of the column-major m x n result; block is (e-b) x n, ld = e-b
        const int64_t rows = e - b;
        for (int64_t j = 0; j < n; ++j) memcpy(result + b + j * m, block + j * rows, (size_t)rows * sizeof(T));
    } else {                               // columns [b, e): contiguous
        memcpy(result + b * m, block, (size_t)((e - b) * m) * sizeof(T));
    }
}

extern "C" {

int rdb_comm_unique_id(void *id) {
    if (!id) return rdb_fail(RDB_EINVAL, "null id");
    if (!nccl().ok) return rdb_fail(RDB_EUNSUPPORTED, "libnccl.so.2 could not be loaded: %s", dlerror() ? dlerror() : "missing symbol");
    ncclUniqueId_ u;
    NC(nccl().GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return RDB_OK;
}

rdb_comm *rdb_comm_create(rdb_ctx *ctx, int world, int rank, const void *id) {
    if (!ctx || !id || world < 1 || rank < 0 || rank >= world) { rdb_fail(RDB_EINVAL, "bad communicator arguments"); return nullptr; }
    if (!nccl().ok) { rdb_fail(RDB_EUNSUPPORTED, "libnccl.so.2 could not be loaded"); return nullptr; }
    cudaSetDevice(rdb_ctx_device(ctx));
    auto *c = new rdb_comm();
    c->ctx = ctx; c->world = world; c->rank = rank;
    ncclUniqueId_ u;
    memcpy(&u, id, sizeof(u));
    int r = nccl().CommInitRank(&c->comm, world, u, rank);
    if (r != 0) { rdb_fail(RDB_ECUDA, "ncclCommInitRank failed: %s", nccl().GetErrorString(r)); delete c; return nullptr; }
    return c;
}

void rdb_comm_destroy(rdb_comm *c) {
    if (!c) return;
    cudaSetDevice(rdb_ctx_device(c->ctx));
    cudaStreamSynchronize(rdb_ctx_stream(c->ctx));
    if (c->comm) nccl().CommDestroy(c->comm);
    if (c->stage) cudaFree(c->stage);
    delete c;
}

int rdb_gather_result(rdb_comm *c, int dtype, int kind, const void *block, const int64_t *bounds, int64_t other_dim, void *result, int root) {
    if (!c || !bounds || (kind != 0 && kind != 1) || (dtype != RDB_F64 && dtype != RDB_F32)) return rdb_fail(RDB_EINVAL, "bad gather arguments");
    const int W = c->world, me = c->rank;
    if (root >= W) return rdb_fail(RDB_EINVAL, "root %d outside the communicator", root);
    for (int r = 0; r < W; ++r) if (bounds[r] > bounds[r + 1] || bounds[0] != 0) return rdb_fail(RDB_EINVAL, "bounds must be a partition of [0, m)");
    const bool receiver = root < 0 || root == me;
    if (receiver && !result) return rdb_fail(RDB_EINVAL, "null result on a receiving rank");
    if (bounds[me + 1] > bounds[me] && !block) return rdb_fail(RDB_EINVAL, "null block");
    const size_t es = dtype == RDB_F64 ? 8 : 4;
    const int64_t total = bounds[W];                               // m (row blocks) or number of columns (column blocks)
    cudaSetDevice(rdb_ctx_device(c->ctx));
    cudaStream_t st = rdb_ctx_stream(c->ctx);
    // where rank r's block lands on a receiver: in place (column blocks: result + bounds[r] * ld, ld = other_dim) or in the
    // staging buffer (row blocks)
    char *land = (char *)result;
    if (kind == 0 && receiver) {
        const size_t need = (size_t)total * other_dim * es;
        if (need > c->stage_bytes) {
            if (c->stage) cudaFree(c->stage);
            c->stage = nullptr; c->stage_bytes = 0;
            CUC(cudaMalloc(&c->stage, need));
            c->stage_bytes = need;
        }
        land = (char *)c->stage;
    }
    auto offset = [&](int r) { return (size_t)bounds[r] * other_dim * es; };
    auto bytes = [&](int r) { return (size_t)(bounds[r + 1] - bounds[r]) * other_dim * es; };
    NC(nccl().GroupStart());
    if (root < 0) {
        for (int r = 0; r < W; ++r)
            if (bytes(r)) NC(nccl().Broadcast(r == me ? block : nullptr, land + offset(r), bytes(r), kNcclUint8, r, c->comm, st));
    } else if (me == root) {
        for (int r = 0; r < W; ++r)
            if (r != me && bytes(r)) NC(nccl().Recv(land + offset(r), bytes(r), kNcclUint8, r, c->comm, st));
    } else if (bytes(me)) {
        NC(nccl().Send(block, bytes(me), kNcclUint8, root, c->comm, st));
    }
    NC(nccl().GroupEnd());
    if (receiver && root >= 0 && bytes(me) && (const char *)block != land + offset(me))
        CUC(cudaMemcpyAsync(land + offset(me), block, bytes(me), cudaMemcpyDeviceToDevice, st));
    if (kind == 0 && receiver) {
        // result[bounds[r] + i + j*m] = stage_r[i + j*rows_r]
        for (int r = 0; r < W; ++r) {
            const int64_t rows = bounds[r + 1] - bounds[r];
            if (!rows) continue;
            cudaError_t e = dtype == RDB_F64
                ? launch_copy2d<double>(st, (double *)result + bounds[r], total, (const double *)(land + offset(r)), rows, rows, other_dim)
                : launch_copy2d<float>(st, (float *)result + bounds[r], total, (const float *)(land + offset(r)), rows, rows, other_dim);
            if (e != cudaSuccess) return rdb_fail(RDB_ECUDA, "block placement kernel failed: %s", cudaGetErrorString(e));
        }
    }
    CUC(cudaStreamSynchronize(st));
    return RDB_OK;
}

int rdb_place_block(int dtype, int kind, const void *block, int64_t begin, int64_t end, int64_t m, int64_t n, void *result) {
    if (!block || !result || begin < 0 || end < begin || (kind != 0 && kind != 1)) return rdb_fail(RDB_EINVAL, "bad block placement arguments");
    if ((kind == 0 && end > m) || (kind == 1 && end > n)) return rdb_fail(RDB_EINVAL, "block outside the result");
    if (dtype == RDB_F64) place_host<double>(kind, (const double *)block, begin, end, m, n, (double *)result);
    else if (dtype == RDB_F32) place_host<float>(kind, (const float *)block, begin, end, m, n, (float *)result);
    else return rdb_fail(RDB_EINVAL, "unknown dtype");
    return RDB_OK;
}

}  // extern "C"
