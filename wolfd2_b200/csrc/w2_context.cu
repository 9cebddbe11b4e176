// w2_context.cu -- library configuration, device context, host<->device field copies.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "w2.cuh"

int g_mnx = 302, g_mny = 302, g_mgri = 20, g_mgrj = 10, g_device = 0;  // include/config.f:28-35
static char g_err[1024] = "";

void w2_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    fprintf(stderr, "wolfd2_b200: %s\n", g_err);
}

extern "C" const char *wolfd2_b200_last_error(void) { return g_err; }
extern "C" const char *wolfd2_b200_version(void) { return "wolfd2_b200 0.1 (sm_100a, fp64, fmad=false)"; }

extern "C" int wolfd2_b200_config(int32_t mnx, int32_t mny, int32_t mgri, int32_t mgrj) {
    if (mnx < 4 || mny < 4 || mgri < 1 || mgrj < 1 || (long long)mgri * mgrj > W2_MAXREG) {
        w2_set_error("wolfd2_b200_config: bad dimensions mnx=%d mny=%d mgri=%d mgrj=%d (mgri*mgrj <= %d)",
                     mnx, mny, mgri, mgrj, W2_MAXREG);
        return W2_ERR_BAD_ARG;
    }
    g_mnx = mnx; g_mny = mny; g_mgri = mgri; g_mgrj = mgrj;
    return W2_OK;
}

extern "C" int wolfd2_b200_set_device(int32_t device) {
    g_device = device;
    return W2_OK;
}

static int check_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        w2_set_error("no usable CUDA device (%s); this library has no CPU fallback",
                     e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return W2_ERR_NO_DEVICE;
    }
    if (g_device >= n) {
        w2_set_error("device %d requested but only %d present", g_device, n);
        return W2_ERR_NO_DEVICE;
    }
    W2_CUDA(cudaSetDevice(g_device));
    return W2_OK;
}

// (cudaMemset runs on the legacy default stream, which the context's non-blocking streams do not wait for: without the
// synchronisation the zero fill of an array allocated on first use could land AFTER the first kernel that writes it --
// seen as a wrong first step of a fresh one-GPU context once YMomentum had a stream of its own.  Allocation is rare.)
static int dalloc(double **p, size_t n) {
    W2_CUDA(cudaMalloc((void **)p, n * sizeof(double)));
    W2_CUDA(cudaMemset(*p, 0, n * sizeof(double)));
    W2_CUDA(cudaDeviceSynchronize());
    return W2_OK;
}
// field-layout array of this rank's rows; the stored pointer is shifted so that f[i + pitch*j] takes GLOBAL j
static int falloc(wolfd2_ctx *c, double **p) {
    W2_TRY(dalloc(p, c->nelem));
    *p -= c->row_off;
    return W2_OK;
}
static int balloc(wolfd2_ctx *c, unsigned char **p, size_t planes) {
    W2_CUDA(cudaMalloc((void **)p, planes * c->nelem));
    W2_CUDA(cudaMemset(*p, 0, planes * c->nelem));
    W2_CUDA(cudaDeviceSynchronize());
    *p -= c->row_off;
    return W2_OK;
}
// chain-layout work arrays of the line solvers (ppe_solver 2-4) and of the alttridlu_ / rhsppe_ shims: allocated on
// first use (five full-size arrays: 10.7 GB at 16384^2 that the point-SOR path never touches)
int w2_ensure_chain(wolfd2_ctx *c) {
    if (c->tx) return W2_OK;
    const long long nmax = ((long long)c->nx * (long long)c->ny / 4096 + 3) * 4096;
    W2_TRY(dalloc(&c->ta, (size_t)nmax));
    W2_TRY(dalloc(&c->td, (size_t)nmax));
    W2_TRY(dalloc(&c->tc, (size_t)nmax));
    W2_TRY(dalloc(&c->tb, (size_t)nmax));
    W2_TRY(dalloc(&c->tx, (size_t)nmax));
    return W2_OK;
}
int w2_alloc_pormap(wolfd2_ctx *c) { return c->pormap ? W2_OK : balloc(c, &c->pormap, 6); }
int w2_alloc_field(wolfd2_ctx *c, double **p) { return falloc(c, p); }
static void ffree(wolfd2_ctx *c, double *p) { if (p) cudaFree(p + c->row_off); }
static void bfree(wolfd2_ctx *c, unsigned char *p) { if (p) cudaFree(p + c->row_off); }

// Row slab of `rank` (same arithmetic as wolfd2_b200/slab.py slab_rows): the unknown pressure rows 2..ny in
// `world` contiguous blocks, earlier ranks take the remainder.  Halo depth: 2T = 4 rows for the fused SOR
// pass and the rows a straddling 2048-unknown momentum segment reaches into, plus the stencil row.
void w2_slab_layout(int nx, int ny, int world, int rank, int *J0, int *J1, int *A0, int *A1, int *HG) {
    const int nrows = ny - 1, base = nrows / world, rem = nrows % world;
    *J0 = 2 + rank * base + (rank < rem ? rank : rem);
    *J1 = *J0 + base + (rank < rem ? 1 : 0) - 1;
    *HG = world == 1 ? 0 : 4 + 2047 / (nx - 1);
    if (world > 1 && *HG < 5) *HG = 5;
    *A0 = rank == 0 ? 0 : *J0 - *HG;
    *A1 = rank == world - 1 ? ny + 1 : *J1 + *HG;
    if (*A0 < 0) *A0 = 0;
    if (*A1 > ny + 1) *A1 = ny + 1;
}
extern "C" int wolfd2_b200_slab_layout(int32_t nx, int32_t ny, int32_t world, int32_t rank, int32_t out[5]) {
    if (world < 1 || rank < 0 || rank >= world || nx < 6 || ny < 6) { w2_set_error("slab_layout: bad arguments"); return W2_ERR_BAD_ARG; }
    int J0, J1, A0, A1, HG;
    w2_slab_layout(nx, ny, world, rank, &J0, &J1, &A0, &A1, &HG);
    if (world > 1 && (ny - 1) / world < 2 * HG) {
        w2_set_error("%d rows cannot be split into %d slabs of >= %d rows (2 x halo depth)", ny - 1, world, 2 * HG);
        return W2_ERR_BAD_ARG;
    }
    out[0] = J0; out[1] = J1; out[2] = A0; out[3] = A1; out[4] = HG;
    return W2_OK;
}

// Translate the Fortran region tables (bound_cond.f SetUpBCs output) to the device struct.
int w2_fill_regions(W2Regions *r, int nx, int ny, const int32_t *nReg, const int32_t *nRegBrd,
                    const int32_t *nRegType, const int32_t *nMomBdTp, const double *dBCVal,
                    const double *poros, const double *c1, const double *c2) {
    memset(r, 0, sizeof(*r));
    const int ni = nReg[0], nj = nReg[1];
    if (ni < 1 || nj < 1 || ni > g_mgri || nj > g_mgrj || ni * nj > W2_MAXREG) {
        w2_set_error("bad nReg = (%d,%d) for mgri=%d mgrj=%d", ni, nj, g_mgri, g_mgrj);
        return W2_ERR_BAD_ARG;
    }
    r->nregI = ni; r->nregJ = nj; r->nreg = ni * nj;
    for (int jr = 1; jr <= nj; ++jr)
        for (int ir = 1; ir <= ni; ++ir) {
            const int q = (ir - 1) + ni * (jr - 1);
            const int base = (ir - 1) + g_mgri * (jr - 1);
            const int plane = g_mgri * g_mgrj;
            r->iW[q] = nRegBrd[base + plane * (W2_WEST - 1)];
            r->iE[q] = nRegBrd[base + plane * (W2_EAST - 1)];
            r->jS[q] = nRegBrd[base + plane * (W2_SOUTH - 1)];
            r->jN[q] = nRegBrd[base + plane * (W2_NORTH - 1)];
            if (r->iW[q] < 1 || r->iE[q] > nx || r->jS[q] < 1 || r->jN[q] > ny ||
                r->iW[q] >= r->iE[q] || r->jS[q] >= r->jN[q]) {
                w2_set_error("region (%d,%d) has bad borders W=%d E=%d S=%d N=%d for grid %dx%d", ir, jr,
                             r->iW[q], r->iE[q], r->jS[q], r->jN[q], nx, ny);
                return W2_ERR_BAD_ARG;
            }
            r->type[q] = nRegType ? nRegType[base] : W2_RM_INTERN;
            if (r->type[q] == W2_RM_BLOCKG) r->has_blockage = 1;
            if (r->type[q] == W2_RM_POROUS) r->has_porous = 1;
            for (int k = 0; k < 4; ++k) {
                r->bd[q][k] = nMomBdTp ? nMomBdTp[base + plane * k] : W2_BM_INTERN;
                if (r->bd[q][k] < W2_BM_INTERN || r->bd[q][k] > W2_BM_OUTLT2) {
                    // the reference prints 'Wrong nBdType? flag in region' and stops (e.g. momentum.f:472-474)
                    w2_set_error("Wrong nBdType flag %d in region %d,%d face %d", r->bd[q][k], ir, jr, k + 1);
                    return W2_ERR_BAD_ARG;
                }
                for (int l = 0; l < 4; ++l)
                    r->val[q][k][l] = dBCVal ? dBCVal[base + plane * (k + 4 * l)] : 0.0;
            }
            r->poros[q] = poros ? poros[base] : 1.0;
            r->porc1[q] = c1 ? c1[base] : 0.0;
            r->porc2[q] = c2 ? c2[base] : 0.0;
        }
    for (int jr = 1; jr <= nj; ++jr)
        for (int ir = 1; ir <= ni; ++ir) {
            const int q = (ir - 1) + ni * (jr - 1);
            r->nbW[q] = (ir > 1 && r->type[q - 1] == W2_RM_BLOCKG);
            r->nbE[q] = (ir < ni && r->type[q + 1] == W2_RM_BLOCKG);
            r->nbS[q] = (jr > 1 && r->type[q - ni] == W2_RM_BLOCKG);
            r->nbN[q] = (jr < nj && r->type[q + ni] == W2_RM_BLOCKG);
        }
    return W2_OK;
}

int w2_ctx_set_regions(wolfd2_ctx *c, const W2Regions *r) {
    if (r != &c->hreg) c->hreg = *r;
    W2_CUDA(cudaMemcpyAsync(c->dreg, &c->hreg, sizeof(W2Regions), cudaMemcpyHostToDevice, c->stream));
    W2_CUDA(cudaStreamSynchronize(c->stream));  // hreg may be overwritten by the next shim call
    W2_TRY(w2_build_pmask(c));
    W2_TRY(w2_build_mom_masks(c));
    return W2_OK;
}

static int ctx_create_body(wolfd2_ctx *c, int nx, int ny, int rank, int world, const int32_t *lay);

int w2_ctx_create_raw(wolfd2_ctx **out, int nx, int ny, int rank, int world) {
    *out = nullptr;
    W2_TRY(check_device());
    int32_t lay[5];
    W2_TRY(wolfd2_b200_slab_layout(nx, ny, world, rank, lay));
    if (nx < 6 || ny < 6) {
        w2_set_error("grid %dx%d too small (need nx,ny >= 6)", nx, ny);
        return W2_ERR_BAD_ARG;
    }
    if (nx + 1 > g_mnx || (world == 1 ? ny + 1 : lay[3] - lay[2]) > g_mny) {  // CheckGridSize, src/grid.f:551
        w2_set_error("grid %dx%d does not fit mnx=%d mny=%d (need mnx>=nx+1, mny>=ny+1); call wolfd2_b200_config",
                     nx, ny, g_mnx, g_mny);
        return W2_ERR_BAD_ARG;
    }
    wolfd2_ctx *c = (wolfd2_ctx *)calloc(1, sizeof(wolfd2_ctx));
    if (!c) return W2_ERR_BAD_ARG;
    // a half-built context (e.g. out of memory at 16384^2) is torn down again: destroy copes with null members
    const int rc = ctx_create_body(c, nx, ny, rank, world, lay);
    if (rc != W2_OK) { wolfd2_b200_destroy(c); return rc; }
    *out = c;
    return W2_OK;
}

static int ctx_create_body(wolfd2_ctx *c, int nx, int ny, int rank, int world, const int32_t *lay) {
    c->device = g_device;
    c->nx = nx; c->ny = ny; c->mnx = g_mnx; c->mny = g_mny;
    c->pitch = ((nx + 2 + 15) / 16) * 16;
    c->rank = rank; c->world = world;
    c->J0 = lay[0]; c->J1 = lay[1]; c->A0 = lay[2]; c->A1 = lay[3]; c->HG = lay[4];
    c->E0 = rank == 0 ? 0 : c->J0;
    c->E1 = rank == world - 1 ? ny + 1 : c->J1;
    c->rows = c->A1 - c->A0 + 1;
    c->row_off = (size_t)c->pitch * (size_t)c->A0;
    c->nelem = (size_t)c->pitch * (size_t)(c->rows + 1) + 512;  // guard: strip loads may overrun a row
    cudaDeviceProp prop;
    W2_CUDA(cudaGetDeviceProperties(&prop, c->device));
    c->num_sms = prop.multiProcessorCount;
    c->coop_ok = prop.cooperativeLaunch;
    if (prop.major < 10) {
        w2_set_error("device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
        return W2_ERR_NO_DEVICE;
    }
    W2_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int k = 0; k < 8; ++k) W2_CUDA(cudaEventCreate(&c->ev[k]));
    W2_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    W2_CUDA(cudaEventCreateWithFlags(&c->ev_p, cudaEventDisableTiming));
    W2_CUDA(cudaMalloc((void **)&c->dreg, sizeof(W2Regions)));
    double **mp = &c->met.rau;
    for (int k = 0; k < 30; ++k) W2_TRY(falloc(c, &mp[k]));
    for (int k = 0; k < W2_F_CORE; ++k) W2_TRY(falloc(c, &c->fld[k]));
    W2_TRY(falloc(c, &c->dus));
    W2_TRY(falloc(c, &c->dvs));
    W2_TRY(falloc(c, &c->div));
    W2_TRY(falloc(c, &c->x1));
    W2_TRY(balloc(c, &c->pmask, 1));
    W2_TRY(balloc(c, &c->xmask, 1));
    W2_TRY(balloc(c, &c->ymask, 1));
    W2_TRY(balloc(c, &c->tmask, 1));
    W2_TRY(falloc(c, &c->heat_s));
    W2_CUDA(cudaMalloc((void **)&c->dth, sizeof(W2Thermal)));
    W2_CUDA(cudaMemset(c->dth, 0, sizeof(W2Thermal)));   // no heat sources, internal faces: TAveraged is then the plain average
    c->th.pe = 1.0;
    // chain arrays are read in whole segments by the tridiagonal solver: pad generously
    const long long nmax = ((long long)nx * (long long)ny / 4096 + 3) * 4096;
    // level-0 spike arrays hold only this rank's segments; the segment table and upper levels are global
    const long long cap0 = world == 1 ? nmax : (long long)(c->rows + 2) * nx + 3 * 2048;
    W2_TRY(w2_tri_prepare(c, nmax, cap0));
    W2_CUDA(cudaMalloc((void **)&c->d_norm, 64 * sizeof(unsigned long long)));
    W2_CUDA(cudaMemset(c->d_norm, 0, 64 * sizeof(unsigned long long)));
    W2_CUDA(cudaMalloc((void **)&c->d_flags, 64 * sizeof(int)));
    W2_CUDA(cudaMemset(c->d_flags, 0, 64 * sizeof(int)));
    W2_CUDA(cudaMallocHost((void **)&c->h_norm, 64 * sizeof(unsigned long long)));
    W2_CUDA(cudaMallocHost((void **)&c->h_flags, 64 * sizeof(int)));
    W2_CUDA(cudaMallocHost((void **)&c->h_sor, 96 * sizeof(int)));
    memset(c->h_sor, 0, 96 * sizeof(int));
    return W2_OK;
}

extern "C" void wolfd2_b200_destroy(wolfd2_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    w2_atd_release(c);
    w2_traj_release(c);
    w2_probes_release(c);
    w2_timeavg_release(c);
    double **mp = &c->met.rau;
    for (int k = 0; k < 30; ++k) ffree(c, mp[k]);
    for (int k = 0; k < W2_F_COUNT; ++k) ffree(c, c->fld[k]);
    ffree(c, c->dus); ffree(c, c->dvs); ffree(c, c->div); ffree(c, c->x1); ffree(c, c->x1b); ffree(c, c->qh);
    bfree(c, c->pmask); bfree(c, c->xmask); bfree(c, c->ymask); bfree(c, c->pormap); bfree(c, c->tmask);
    ffree(c, c->heat_s); cudaFree(c->dth);
    w2_peer_release(c);   // before the buffers the peers have mapped go away
    for (int k = 0; k < 4; ++k) ffree(c, c->sorf_buf[k]);
    if (c->sorf_coef) cudaFree(c->sorf_coef);
    cudaFree(c->ta); cudaFree(c->td); cudaFree(c->tc); cudaFree(c->tb); cudaFree(c->tx);
    w2_tri_release(c);
    cudaFree(c->d_norm); cudaFree(c->d_flags); cudaFree(c->dreg);
    cudaFreeHost(c->h_norm); cudaFreeHost(c->h_flags); cudaFreeHost(c->h_sor);
    for (int k = 0; k < 2; ++k) { if (c->ev_ql[k]) cudaEventDestroy(c->ev_ql[k]); if (c->ev_sor[k]) cudaEventDestroy(c->ev_sor[k]); }
    if (c->h_stage) cudaFreeHost(c->h_stage);
    for (int k = 0; k < 8; ++k) if (c->ev[k]) cudaEventDestroy(c->ev[k]);
    if (c->ev_p) cudaEventDestroy(c->ev_p);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->stream) cudaStreamDestroy(c->stream);
    free(c);
}

// Host (0:mnx,0:mny) <-> device pitched copies of the (0..nx+1, 0..ny+1) window.  In a slab context the
// host array holds this rank's rows only: host row 0 is global row A0 (wolfd2_b200_slab_layout).
int w2_upload2d(wolfd2_ctx *c, double *dev, const double *host, cudaStream_t stream) {
    if (dev == c->met.rau || dev == c->met.rgv) c->sorf_met_valid = 0;   // the fused SOR keeps colour-split copies
    for (int k = 0; k < 30; ++k)
        if (dev == (&c->met.rau)[k] && c->cart_state != -2) c->cart_state = 0;   // metrics changed: re-verify the Cartesian classes
    W2_CUDA(cudaMemcpy2DAsync(dev + c->row_off, (size_t)c->pitch * 8, host, (size_t)(c->mnx + 1) * 8,
                              (size_t)(c->nx + 2) * 8, (size_t)c->rows, cudaMemcpyHostToDevice,
                              stream ? stream : c->stream));
    return W2_OK;
}
int w2_download2d(wolfd2_ctx *c, double *host, const double *dev) {
    W2_CUDA(cudaMemcpy2DAsync(host, (size_t)(c->mnx + 1) * 8, dev + c->row_off, (size_t)c->pitch * 8,
                              (size_t)(c->nx + 2) * 8, (size_t)c->rows, cudaMemcpyDeviceToHost,
                              c->stream));
    return W2_OK;
}

extern "C" int wolfd2_b200_set_params(wolfd2_ctx *c, const wolfd2_params *par) {
    if (!c || !par) return W2_ERR_BAD_ARG;
    if (par->nx != c->nx || par->ny != c->ny) {
        w2_set_error("set_params: grid size %dx%d differs from the context's %dx%d", par->nx, par->ny, c->nx, c->ny);
        return W2_ERR_BAD_ARG;
    }
    if (par->nPpeSolver < 1 || par->nPpeSolver > 6) {
        w2_set_error("Wrong nPpeSolver flag passed to Ppe: %d", par->nPpeSolver);  // pressure.f:238-239
        return W2_ERR_BAD_ARG;
    }
    c->par = *par;
    return W2_OK;
}

extern "C" int wolfd2_b200_create(wolfd2_ctx **out, const wolfd2_params *par, const wolfd2_regions *reg,
                                  const wolfd2_metrics *met) {
    return wolfd2_b200_create_slab(out, par, reg, met, 0, 1);
}

// One context per rank of a multi-GPU run (after wolfd2_b200_comm_init).  par and reg describe the GLOBAL
// grid; the metric arrays (and every field passed to upload/download/step_host) hold this rank's rows
// A0..A1 of wolfd2_b200_slab_layout, host row 0 = global row A0.
extern "C" int wolfd2_b200_create_slab(wolfd2_ctx **out, const wolfd2_params *par, const wolfd2_regions *reg,
                                       const wolfd2_metrics *met, int32_t rank, int32_t world) {
    if (!out || !par || !reg || !met) return W2_ERR_BAD_ARG;
    if (world > 1 && (world != w2_dist_world() || rank != w2_dist_rank())) {
        w2_set_error("create_slab(rank %d of %d) does not match wolfd2_b200_comm_init (rank %d of %d)", rank, world,
                     w2_dist_rank(), w2_dist_world());
        return W2_ERR_BAD_ARG;
    }
    if (world > 1) {
        if (par->nPpeSolver != W2_PPE_RB_SOR && par->nPpeSolver != W2_PPE_PAR_RB_SOR) {
            w2_set_error("multi-GPU runs support ppe_solver 5/6 (rb_sor, par_rb_sor) only, got %d", par->nPpeSolver);
            return W2_ERR_UNSUPPORTED;
        }
        if (!par->lCartesGrid || par->nx < 254) {
            w2_set_error("multi-GPU runs need a Cartesian grid with nx >= 254 (fused SOR pipeline)");
            return W2_ERR_UNSUPPORTED;
        }
    }
    wolfd2_ctx *c = nullptr;
    W2_TRY(w2_ctx_create_raw(&c, par->nx, par->ny, rank, world));
    int rc = wolfd2_b200_set_params(c, par);
    if (rc == W2_OK) {
        W2Regions r;
        rc = w2_fill_regions(&r, par->nx, par->ny, reg->nReg, reg->nRegBrd, reg->nRegType, reg->nMomBdTp,
                             reg->dBCVal, reg->dPRporos, reg->dPRporc1, reg->dPRporc2);
        if (rc == W2_OK && world > 1) {   // OUTLT2 west / east faces: recurrences along the face, across the slabs (w2_bc.cu)
            bool o2 = false;
            for (int q = 0; q < r.nreg; ++q) o2 |= r.bd[q][W2_WEST - 1] == W2_BM_OUTLT2 || r.bd[q][W2_EAST - 1] == W2_BM_OUTLT2;
            if (o2) {
                for (int k = 0; k < 2 && rc == W2_OK; ++k)
                    if (!c->sorf_buf[k]) rc = w2_alloc_field(c, &c->sorf_buf[k]);   // the peer set-up exports them
                if (rc == W2_OK) rc = w2_peer_setup(c);
                if (rc == W2_OK && c->peer.state != 1) {
                    w2_set_error("multi-GPU runs need CUDA IPC peer mapping for OUTLT2 (mass_cons) faces on west / east borders");
                    rc = W2_ERR_UNSUPPORTED;
                }
            }
        }
        if (rc == W2_OK) rc = w2_ctx_set_regions(c, &r);
    }
    if (rc == W2_OK) {
        const double *const *hp = &met->rau;
        double **dp = &c->met.rau;
        for (int k = 0; k < 30 && rc == W2_OK; ++k)
            if (hp[k]) rc = w2_upload2d(c, dp[k], hp[k]);
        if (rc == W2_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = W2_ERR_CUDA;
    }
    if (rc != W2_OK) { wolfd2_b200_destroy(c); return rc; }
    *out = c;
    return W2_OK;
}

extern "C" int wolfd2_b200_upload_field(wolfd2_ctx *c, int32_t which, const double *host) {
    if (!c || !host || which < 0 || which >= W2_F_COUNT) return W2_ERR_BAD_ARG;
    if (!c->fld[which]) { w2_set_error("field %d does not exist in this context (small-scale model off)", which); return W2_ERR_BAD_ARG; }
    W2_CUDA(cudaSetDevice(c->device));
    W2_TRY(w2_upload2d(c, c->fld[which], host));
    if (which == W2_F_D || which == W2_F_DN) { c->dn_valid = 0; c->d_nonzero = 1; }
    W2_CUDA(cudaStreamSynchronize(c->stream));
    return W2_OK;
}
// Rows jfirst .. jfirst+nrows-1 (GLOBAL row indices, inside the rows this context holds) of metric array `which`
// (0-based position in wolfd2_metrics) from a host block of nrows x (mnx+1) doubles.  Lets a caller build the
// metrics of a large grid window by window instead of holding 30 full arrays on the host (16384^2: 64 GB).
extern "C" int wolfd2_b200_upload_metric_rows(wolfd2_ctx *c, int32_t which, int32_t jfirst, int32_t nrows, const double *host) {
    if (!c || !host || which < 0 || which >= 30 || nrows < 0) return W2_ERR_BAD_ARG;
    if (jfirst < c->A0 || jfirst + nrows - 1 > c->A1) {
        w2_set_error("upload_metric_rows: rows %d..%d outside the rows %d..%d held by this context", jfirst, jfirst + nrows - 1, c->A0, c->A1);
        return W2_ERR_BAD_ARG;
    }
    if (nrows == 0) return W2_OK;
    W2_CUDA(cudaSetDevice(c->device));
    double *dev = (&c->met.rau)[which];
    if (dev == c->met.rau || dev == c->met.rgv) c->sorf_met_valid = 0;
    if (c->cart_state != -2) c->cart_state = 0;
    W2_CUDA(cudaMemcpy2DAsync(dev + (size_t)c->pitch * (size_t)jfirst, (size_t)c->pitch * 8, host, (size_t)(c->mnx + 1) * 8,
                              (size_t)(c->nx + 2) * 8, (size_t)nrows, cudaMemcpyHostToDevice, c->stream));
    W2_CUDA(cudaStreamSynchronize(c->stream));   // the caller reuses the host block
    return W2_OK;
}
extern "C" int wolfd2_b200_download_field(wolfd2_ctx *c, int32_t which, double *host) {
    if (!c || !host || which < 0 || which >= W2_F_COUNT) return W2_ERR_BAD_ARG;
    if (!c->fld[which]) { w2_set_error("field %d does not exist in this context (small-scale model off)", which); return W2_ERR_BAD_ARG; }
    W2_CUDA(cudaSetDevice(c->device));
    W2_TRY(w2_download2d(c, host, c->fld[which]));
    W2_CUDA(cudaStreamSynchronize(c->stream));
    return W2_OK;
}
extern "C" int wolfd2_b200_sync(wolfd2_ctx *c) {
    if (!c) return W2_ERR_BAD_ARG;
    W2_CUDA(cudaStreamSynchronize(c->stream));
    return W2_OK;
}

extern "C" void *wolfd2_b200_host_alloc(uint64_t bytes) {
    void *p = nullptr;
    if (cudaSetDevice(g_device) != cudaSuccess || cudaMallocHost(&p, bytes) != cudaSuccess) {
        w2_set_error("host_alloc(%llu) failed: %s", (unsigned long long)bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    memset(p, 0, bytes);
    return p;
}
extern "C" void wolfd2_b200_host_free(void *p) { if (p) cudaFreeHost(p); }

extern int g_sor_T;
extern int g_mom_np_cache;
extern int g_mom_cart;
extern int g_sor_resident;
extern int g_sor_slab_inpass;
extern int g_mom_two_streams;
extern "C" int wolfd2_b200_set_option(const char *name, int32_t value) {
    if (name && !strcmp(name, "mom_np_cache")) { g_mom_np_cache = value != 0; return W2_OK; }
    if (name && !strcmp(name, "mom_cart")) { g_mom_cart = value != 0; return W2_OK; }
    if (name && !strcmp(name, "sor_resident")) { g_sor_resident = value != 0; return W2_OK; }
    if (name && !strcmp(name, "mom_two_streams")) { g_mom_two_streams = value != 0; return W2_OK; }
    if (name && !strcmp(name, "sor_slab_inpass")) { g_sor_slab_inpass = value != 0; return W2_OK; }
    if (name && !strcmp(name, "sor_fused_T")) {
        if (value < -1 || value > 2) { w2_set_error("sor_fused_T must be -1 (default), 0, 1 or 2"); return W2_ERR_BAD_ARG; }
        g_sor_T = value;
        return W2_OK;
    }
    w2_set_error("unknown option %s", name ? name : "(null)");
    return W2_ERR_BAD_ARG;
}
