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
#include <stdlib.h>
#include <string.h>

#include "w2.cuh"

namespace {
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
    LOAD(GetUniqueId) LOAD(CommInitRank) LOAD(CommDestroy) LOAD(AllReduce) LOAD(AllGather) LOAD(Send) LOAD(Recv) LOAD(GroupStart) LOAD(GroupEnd)
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

// ---- peer memory for the fused SOR loop ----------------------------------------------------------------
// Every rank exports its two colour-split pressure buffers and its mailbox with CUDA IPC; the handles travel
// by one ncclAllGather.  Afterwards the SOR passes need no NCCL call: the kernel stores its edge rows
// straight into the neighbours' halo rows and its max-norms into every rank's mailbox over NVLink.
// Set W2_SOR_P2P=0 to keep the NCCL path (all-reduce + send/recv after every pass).
int w2_peer_setup(wolfd2_ctx *c) {
    W2Peer &P = c->peer;
    if (P.state != 0) return W2_OK;
    P.state = -1;
    const char *e = getenv("W2_SOR_P2P");
    // every rank must take the same branch: the decision depends only on the environment and the world size
    if (c->world == 1 || c->world > W2_MAXRANKS || (e && atoi(e) == 0)) return W2_OK;
    struct Handles { cudaIpcMemHandle_t a, b, m, g; int ok; int pad[15]; };
    static_assert(sizeof(Handles) % 8 == 0, "handle record must be a multiple of 8 bytes");
    Handles mine;
    memset(&mine, 0, sizeof(mine));
    W2Mail *mail = nullptr;
    bool ok = cudaMalloc((void **)&mail, sizeof(W2Mail)) == cudaSuccess && cudaMemset(mail, 0, sizeof(W2Mail)) == cudaSuccess;
    ok = ok && cudaIpcGetMemHandle(&mine.a, c->sorf_buf[0] + c->row_off) == cudaSuccess;
    ok = ok && cudaIpcGetMemHandle(&mine.b, c->sorf_buf[1] + c->row_off) == cudaSuccess;
    ok = ok && cudaIpcGetMemHandle(&mine.m, mail) == cudaSuccess;
    W2BcGather *bcg = nullptr;
    const size_t bcg_bytes = sizeof(W2BcGather) + (size_t)2 * (2 * (size_t)(c->ny + 2) + 2) * sizeof(double);
    ok = ok && cudaMalloc((void **)&bcg, bcg_bytes) == cudaSuccess && cudaMemset(bcg, 0, bcg_bytes) == cudaSuccess;
    ok = ok && cudaIpcGetMemHandle(&mine.g, bcg) == cudaSuccess;
    cudaDeviceSynchronize();   // the zero fills above ran on the default stream (see dalloc, w2_context.cu)
    mine.ok = ok ? 1 : 0;
    cudaGetLastError();
    Handles *d_all = nullptr, *h_all = (Handles *)malloc(sizeof(Handles) * c->world);
    W2_CUDA(cudaMalloc((void **)&d_all, sizeof(Handles) * c->world));
    W2_CUDA(cudaMemcpyAsync(d_all + c->rank, &mine, sizeof(Handles), cudaMemcpyHostToDevice, c->stream));
    W2_NCCL(N.AllGather(d_all + c->rank, d_all, sizeof(Handles), ncclChar, g_comm, c->stream));
    W2_CUDA(cudaMemcpyAsync(h_all, d_all, sizeof(Handles) * c->world, cudaMemcpyDeviceToHost, c->stream));
    W2_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_all);
    int all_ok = 1;
    for (int r = 0; r < c->world; ++r) all_ok &= h_all[r].ok;
    // opening can fail too (no peer access between two devices): agree on the outcome with one more gather
    int opened_ok = all_ok;
    P.nopened = 0;
    if (all_ok) {
        for (int r = 0; r < c->world && opened_ok; ++r) {
            if (r == c->rank) { P.mail[r] = mail; P.bcg[r] = bcg; continue; }
            void *m = nullptr;
            if (cudaIpcOpenMemHandle(&m, h_all[r].m, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { opened_ok = 0; break; }
            P.opened[P.nopened++] = m;
            P.mail[r] = (W2Mail *)m;
            void *gq = nullptr;
            if (cudaIpcOpenMemHandle(&gq, h_all[r].g, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { opened_ok = 0; break; }
            P.opened[P.nopened++] = gq;
            P.bcg[r] = (W2BcGather *)gq;
            if (r == c->rank - 1 || r == c->rank + 1) {
                void *a = nullptr, *b = nullptr;
                if (cudaIpcOpenMemHandle(&a, h_all[r].a, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { opened_ok = 0; break; }
                P.opened[P.nopened++] = a;
                if (cudaIpcOpenMemHandle(&b, h_all[r].b, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { opened_ok = 0; break; }
                P.opened[P.nopened++] = b;
                int J0, J1, A0, A1, HG;
                w2_slab_layout(c->nx, c->ny, c->world, r, &J0, &J1, &A0, &A1, &HG);
                const int side = r == c->rank - 1 ? 0 : 1;
                P.nbrA[side] = (double *)a - (size_t)c->pitch * (size_t)A0;   // peer rows are indexed globally too
                P.nbrB[side] = (double *)b - (size_t)c->pitch * (size_t)A0;
            }
        }
        cudaGetLastError();
    }
    free(h_all);
    unsigned long long *d_flag = nullptr;
    W2_CUDA(cudaMalloc((void **)&d_flag, 8));
    unsigned long long bad = opened_ok ? 0ull : 1ull;
    W2_CUDA(cudaMemcpyAsync(d_flag, &bad, 8, cudaMemcpyHostToDevice, c->stream));
    W2_NCCL(N.AllReduce(d_flag, d_flag, 1, ncclUint64, ncclMax, g_comm, c->stream));
    W2_CUDA(cudaMemcpyAsync(&bad, d_flag, 8, cudaMemcpyDeviceToHost, c->stream));
    W2_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_flag);
    if (bad) {
        fprintf(stderr, "wolfd2_b200: CUDA IPC peer mapping unavailable on rank %d; the SOR loop uses NCCL exchanges\n", c->rank);
        for (int k = 0; k < P.nopened; ++k) cudaIpcCloseMemHandle(P.opened[k]);
        P.nopened = 0;
        cudaFree(mail);
        cudaFree(bcg);
        return W2_OK;
    }
    P.state = 1;
    return W2_OK;
}

void w2_peer_release(wolfd2_ctx *c) {
    W2Peer &P = c->peer;
    for (int k = 0; k < P.nopened; ++k) cudaIpcCloseMemHandle(P.opened[k]);
    if (P.state == 1) { cudaFree(P.mail[c->rank]); cudaFree(P.bcg[c->rank]); }
    memset(&P, 0, sizeof(P));
}

// ---- verification plumbing: a slab run collected into ONE global context on rank 0 ---------------------------
// bench.py and the multi-GPU tests use this to run the same grid on one GPU from the same state and compare the
// fields bit for bit.  Rank r contributes the rows it updates (E0..E1: its unknown rows
// plus the physical ghost rows of the first / last rank), so together the ranks cover rows 0..ny+1 exactly once.
static int gather_array(wolfd2_ctx *s, double *src, double *dst /* rank 0: global array, global row indexing */) {
    const size_t pitch = (size_t)s->pitch;
    if (s->rank == 0) {
        W2_CUDA(cudaMemcpyAsync(dst + pitch * (size_t)s->E0, src + pitch * (size_t)s->E0,
                                pitch * (size_t)(s->E1 - s->E0 + 1) * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
        W2_NCCL(N.GroupStart());
        for (int r = 1; r < s->world; ++r) {
            int J0, J1, A0, A1, HG;
            w2_slab_layout(s->nx, s->ny, s->world, r, &J0, &J1, &A0, &A1, &HG);
            const int e0 = J0, e1 = r == s->world - 1 ? s->ny + 1 : J1;
            W2_NCCL(N.Recv(dst + pitch * (size_t)e0, pitch * (size_t)(e1 - e0 + 1), ncclDouble, r, g_comm, s->stream));
        }
        W2_NCCL(N.GroupEnd());
    } else {
        W2_NCCL(N.Send(src + pitch * (size_t)s->E0, pitch * (size_t)(s->E1 - s->E0 + 1), ncclDouble, 0, g_comm, s->stream));
    }
    return W2_OK;
}

static int gather_check(wolfd2_ctx *s, wolfd2_ctx *g) {
    if (!s || s->world < 2 || s->world != g_world || s->rank != g_rank) { w2_set_error("gather: not a slab context of the current communicator"); return W2_ERR_BAD_ARG; }
    if (s->rank == 0 && (!g || g->world != 1 || g->nx != s->nx || g->ny != s->ny || g->device != s->device)) {
        w2_set_error("gather: rank 0 needs a one-GPU context of the same %dx%d grid on the same device", s->nx, s->ny);
        return W2_ERR_BAD_ARG;
    }
    return W2_OK;
}

extern "C" int wolfd2_b200_gather_global(wolfd2_ctx *s, wolfd2_ctx *g, int32_t what) {
    W2_TRY(gather_check(s, g));
    W2_CUDA(cudaSetDevice(s->device));
    if (g) W2_CUDA(cudaStreamSynchronize(g->stream));
    if (what & 1) {
        double **sp = &s->met.rau, **gp = g ? &g->met.rau : nullptr;
        for (int k = 0; k < 30; ++k) W2_TRY(gather_array(s, sp[k], gp ? gp[k] : nullptr));
        if (g) { g->sorf_met_valid = 0; if (g->cart_state != -2) g->cart_state = 0; }
    }
    if (what & 2)
        for (int k = W2_F_U; k <= W2_F_P; ++k) W2_TRY(gather_array(s, s->fld[k], g ? g->fld[k] : nullptr));
    W2_CUDA(cudaStreamSynchronize(s->stream));
    return W2_OK;
}

__global__ void __launch_bounds__(256) field_compare_kernel(int nx, int ny, int pitch, const double *__restrict__ a,
                                                            const double *__restrict__ b, unsigned long long *out) {
    __shared__ double red[32];
    unsigned long long nd = 0;
    double mx = 0.0;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= nx + 1)
        for (int j = blockIdx.y; j <= ny + 1; j += gridDim.y) {
            const double x = a[IDX(i, j)], y = b[IDX(i, j)];
            if (__double_as_longlong(x) != __double_as_longlong(y)) { ++nd; const double d = fabs(x - y); mx = d == d ? fmax(mx, d) : 1.0e300; }
        }
    mx = w2_block_max(mx, red);
    if (threadIdx.x == 0 && mx > 0.0) atomicMax(out + 1, w2_dbits(mx));
    if (nd) atomicAdd(out, nd);
}

// Field `which` of the slab run against the same field of the one-GPU context: number of cells (0..nx+1, 0..ny+1)
// whose bit patterns differ and the largest |difference|.  Results on rank 0 (zeros elsewhere).
extern "C" int wolfd2_b200_compare_global(wolfd2_ctx *s, wolfd2_ctx *g, int32_t which, uint64_t *ndiff, double *maxabs) {
    W2_TRY(gather_check(s, g));
    if (which < 0 || which >= W2_F_CORE) return W2_ERR_BAD_ARG;
    W2_CUDA(cudaSetDevice(s->device));
    if (g) W2_CUDA(cudaStreamSynchronize(g->stream));
    W2_TRY(gather_array(s, s->fld[which], g ? g->div : nullptr));   // div is rebuilt by every Ppe: free as scratch
    if (ndiff) *ndiff = 0;
    if (maxabs) *maxabs = 0.0;
    if (g) {
        W2_CUDA(cudaMemsetAsync(g->d_norm + 32, 0, 16, s->stream));
        dim3 grid((g->nx + 2 + 255) / 256, g->ny + 2 < 1024 ? g->ny + 2 : 1024);
        field_compare_kernel<<<grid, 256, 0, s->stream>>>(g->nx, g->ny, g->pitch, g->div, g->fld[which], g->d_norm + 32);
        W2_CUDA(cudaGetLastError());
        unsigned long long h[2];
        W2_CUDA(cudaMemcpyAsync(h, g->d_norm + 32, 16, cudaMemcpyDeviceToHost, s->stream));
        W2_CUDA(cudaStreamSynchronize(s->stream));
        if (ndiff) *ndiff = h[0];
        if (maxabs) memcpy(maxabs, &h[1], 8);
    } else {
        W2_CUDA(cudaStreamSynchronize(s->stream));
    }
    return W2_OK;
}
