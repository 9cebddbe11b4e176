// w2_dist.cu -- row-slab decomposition across the GPUs of one node (one process per GPU).
//
// Rank g owns the unknown pressure rows [J0, J1] (a contiguous block of j = 2..ny) and holds every array
// on rows [A0, A1] = owned rows +- HG halo rows (clipped to 0..ny+1).  Arrays keep GLOBAL j indexing: the
// pointers stored in the context are shifted by -pitch*A0, so kernels address f[i + pitch*j] exactly as
// on one GPU and only their row loops are clipped.  Because both momentum split steps run along i
//  no tridiagonal line crosses a slab; what crosses are
//   * halo rows of us, vs (per QL iteration and after Project) and of p (per fused SOR pass),
//   * the per-iteration max-norms (all-reduce, bitwise max as on one GPU),
//   * the 10-number records of the level-0 tridiagonal segments (summed into a global table) and the
//     few increments a straddling segment computes in the neighbour's first row.
// All exchanges are NCCL calls enqueued on the context's stream (no host synchronisation).
// NCCL is loaded with dlopen so that the library still loads on machines without it.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "w2.cuh"

namespace {
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
} N;

int load_nccl() {
    if (N.h) return W2_OK;
    N.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!N.h) { w2_set_error("multi-GPU run needs NCCL: dlopen(libnccl.so.2) failed: %s", dlerror()); return W2_ERR_UNSUPPORTED; }
#define LOAD(sym) *(void **)(&N.sym) = dlsym(N.h, "nccl" #sym); if (!N.sym) { w2_set_error("NCCL symbol nccl" #sym " missing"); return W2_ERR_UNSUPPORTED; }
    LOAD(GetUniqueId) LOAD(CommInitRank) LOAD(CommDestroy) LOAD(AllReduce) LOAD(Send) LOAD(Recv) LOAD(GroupStart) LOAD(GroupEnd)
    LOAD(GetErrorString)
#undef LOAD
    return W2_OK;
}
}  // namespace

#define W2_NCCL(call)                                                                              \
    do {                                                                                           \
        ncclResult_t r__ = (call);                                                                 \
        if (r__ != ncclSuccess) {                                                                  \
            w2_set_error("NCCL error %s at %s:%d: %s", #call, __FILE__, __LINE__, N.GetErrorString(r__)); \
            return W2_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

static ncclComm_t g_comm = nullptr;
static int g_rank = 0, g_world = 1;

// 128-byte NCCL unique id, created on rank 0 and handed to every rank by the launcher (bench.py
// broadcasts it with torch.distributed).
extern "C" int wolfd2_b200_comm_unique_id(unsigned char id[128]) {
    W2_TRY(load_nccl());
    ncclUniqueId u;
    W2_NCCL(N.GetUniqueId(&u));
    static_assert(sizeof(ncclUniqueId) == 128, "unexpected ncclUniqueId size");
    memcpy(id, &u, 128);
    return W2_OK;
}

extern "C" int wolfd2_b200_comm_init(int32_t rank, int32_t world, const unsigned char id[128]) {
    if (world < 1 || rank < 0 || rank >= world) { w2_set_error("bad rank/world %d/%d", rank, world); return W2_ERR_BAD_ARG; }
    if (world == 1) { g_rank = 0; g_world = 1; return W2_OK; }
    W2_TRY(load_nccl());
    W2_CUDA(cudaSetDevice(g_device));
    ncclUniqueId u;
    memcpy(&u, id, 128);
    W2_NCCL(N.CommInitRank(&g_comm, world, u, rank));
    g_rank = rank; g_world = world;
    return W2_OK;
}

extern "C" int wolfd2_b200_comm_finalize(void) {
    if (g_comm) { N.CommDestroy(g_comm); g_comm = nullptr; }
    g_rank = 0; g_world = 1;
    return W2_OK;
}

int w2_dist_rank() { return g_rank; }
int w2_dist_world() { return g_world; }

// Exchange `depth` halo rows of a (shifted-pointer, global-j) field with both slab neighbours:
// my top owned rows [J1-depth+1, J1] -> rank+1's rows of the same global index, and so on.
int w2_halo_exchange(wolfd2_ctx *c, double *const *fields, int nfields, int depth) {
    if (c->world == 1) return W2_OK;
    const size_t rowlen = (size_t)c->pitch;
    W2_NCCL(N.GroupStart());
    for (int f = 0; f < nfields; ++f) {
        double *p = fields[f];
        if (c->rank + 1 < c->world) {   // north neighbour
            W2_NCCL(N.Send(p + rowlen * (size_t)(c->J1 - depth + 1), rowlen * depth, ncclDouble, c->rank + 1, g_comm, c->stream));
            W2_NCCL(N.Recv(p + rowlen * (size_t)(c->J1 + 1), rowlen * depth, ncclDouble, c->rank + 1, g_comm, c->stream));
        }
        if (c->rank > 0) {              // south neighbour
            W2_NCCL(N.Send(p + rowlen * (size_t)c->J0, rowlen * depth, ncclDouble, c->rank - 1, g_comm, c->stream));
            W2_NCCL(N.Recv(p + rowlen * (size_t)(c->J0 - depth), rowlen * depth, ncclDouble, c->rank - 1, g_comm, c->stream));
        }
    }
    W2_NCCL(N.GroupEnd());
    c->launches[0] += 1;
    return W2_OK;
}

// In-place all-reduce of n 64-bit words: bitwise max (non-negative doubles: same as atomicMax on one GPU)
int w2_allreduce_max_u64(wolfd2_ctx *c, unsigned long long *d, int n) {
    if (c->world == 1) return W2_OK;
    W2_NCCL(N.AllReduce(d, d, (size_t)n, ncclUint64, ncclMax, g_comm, c->stream));
    return W2_OK;
}
// In-place sum of doubles (each entry is non-zero on exactly one rank, so the sum is exact)
int w2_allreduce_sum_f64(wolfd2_ctx *c, double *d, size_t n) {
    if (c->world == 1) return W2_OK;
    W2_NCCL(N.AllReduce(d, d, n, ncclDouble, ncclSum, g_comm, c->stream));
    return W2_OK;
}
// Neighbour transfer of a few short contiguous pieces (the increments a straddling tridiagonal segment
// computed in the next slab's first rows): `up` goes to rank+1, `dn` arrives from rank-1, same order.
int w2_send_recv_pieces(wolfd2_ctx *c, const W2Piece *up, int nup, const W2Piece *dn, int ndn) {
    if (c->world == 1 || (nup == 0 && ndn == 0)) return W2_OK;
    W2_NCCL(N.GroupStart());
    for (int k = 0; k < nup; ++k) W2_NCCL(N.Send(up[k].p, up[k].n, ncclDouble, c->rank + 1, g_comm, c->stream));
    for (int k = 0; k < ndn; ++k) W2_NCCL(N.Recv(dn[k].p, dn[k].n, ncclDouble, c->rank - 1, g_comm, c->stream));
    W2_NCCL(N.GroupEnd());
    c->launches[0] += 1;
    return W2_OK;
}
