// w2_ppe.cu -- pressure Poisson equation: Ppe (src/pressure.f:30-249), Divergence (:265-323),
// RhsPpe (:329-378) and the red/black point SOR of SorRB / SorRBP (:457-656).
//
// Data: p, b are pitched 2-D arrays (i fastest).  The 5-point matrix a(mn,5) of the reference
// (:97-106) is never stored: a1=rgv(i,j-1), a2=rau(i-1,j), a4=rau(i,j), a5=rgv(i,j),
// a3=-rau(i,j)-rau(i-1,j)-rgv(i,j)-rgv(i,j-1) are re-formed from the two metric arrays inside the
// sweep (16 B/cell instead of 40).  Blockage rows (:109-196) are a 1-byte/cell mask.
//
// Arithmetic is ordered exactly as in the reference and the library is compiled with
// -fmad=false, so the iterate path, the max-norm and therefore the iteration count are
// bit-identical to a non-FMA CPU build.
//
// Loop control lives on the device (no per-iteration host sync): every sweep kernel first reads
// a `done` word; the last CTA of each red sweep closes the iteration, tests
// `m > 1 .and. dif < sortol` (:534-537) and publishes the result for the launches already queued
// behind it, which then return immediately.
#include <stdlib.h>

#include "w2.cuh"

#define P(i, j) p[IDX(i, j)]

// ------------------------------------------------------------------ Ppe row mask
__global__ void pmask_kernel(const W2Regions *__restrict__ R, int nx, int jlo, int jhi, int pitch,
                             unsigned char *__restrict__ mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = jlo + blockIdx.y;
    if (i > nx + 1 || j > jhi) return;
    unsigned char m = 0;
    for (int q = 0; q < R->nreg; ++q) {
        if (R->type[q] != W2_RM_BLOCKG) continue;
        const int iW = R->iW[q], iE = R->iE[q], jS = R->jS[q], jN = R->jN[q];
        if (j >= jS + 2 && j <= jN - 1 && i >= iW + 2 && i <= iE - 1) m = 1;           // :123-137
        if (R->nbW[q] && i == iW + 1 && j >= jS + 1 && j <= jN) m = 1;                 // :143-153
        if (R->nbE[q] && i == iE && j >= jS + 1 && j <= jN) m = 1;                     // :155-166
        if (R->nbS[q] && j == jS + 1 && i >= iW + 1 && i <= iE) m = 1;                 // :168-178
        if (R->nbN[q] && j == jN && i >= iW + 1 && i <= iE) m = 1;                     // :180-191
    }
    mask[IDX(i, j)] = m;
}

int w2_build_pmask(wolfd2_ctx *c) {
    if (!c->hreg.has_blockage) {
        W2_CUDA(cudaMemsetAsync(c->pmask + c->row_off, 0, c->nelem, c->stream));
        return W2_OK;
    }
    dim3 grid((c->nx + 2 + 255) / 256, c->rows);
    pmask_kernel<<<grid, 256, 0, c->stream>>>(c->dreg, c->nx, c->A0, c->A1, c->pitch, c->pmask);
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}

// ------------------------------------------------------------------ Divergence
// nloc = 1: centre of pressure C.V. (:287-299); nloc = 2: natural grid points (:302-314).
template <int NLOC>
__device__ __forceinline__ double div_point(int i, int j, int pitch, const double *__restrict__ xet,
                                            const double *__restrict__ yet, const double *__restrict__ xzi,
                                            const double *__restrict__ yzi, const double *__restrict__ u,
                                            const double *__restrict__ v) {
#define F(a, ii, jj) a[IDX(ii, jj)]
    if (NLOC == 1) {
        const double ucij = F(yet, i, j) * F(u, i, j)
                            - F(xet, i, j) * (F(v, i + 1, j) + F(v, i, j) + F(v, i + 1, j - 1) + F(v, i, j - 1)) / 4.0;
        const double uci1j = F(yet, i - 1, j) * F(u, i - 1, j)
                             - F(xet, i - 1, j) * (F(v, i, j) + F(v, i - 1, j) + F(v, i, j - 1) + F(v, i - 1, j - 1)) / 4.0;
        const double vcij = F(xzi, i, j) * F(v, i, j)
                            - F(yzi, i, j) * (F(u, i, j + 1) + F(u, i - 1, j + 1) + F(u, i, j) + F(u, i - 1, j)) / 4.0;
        const double vcij1 = F(xzi, i, j - 1) * F(v, i, j - 1)
                             - F(yzi, i, j - 1) * (F(u, i, j) + F(u, i - 1, j) + F(u, i, j - 1) + F(u, i - 1, j - 1)) / 4.0;
        return ucij - uci1j + vcij - vcij1;
    } else {
        const double uci1j = F(yet, i + 1, j) * (F(u, i, j + 1) + F(u, i + 1, j + 1) + F(u, i, j) + F(u, i + 1, j)) / 4.0
                             - F(xet, i + 1, j) * F(v, i + 1, j);
        const double ucij = F(yet, i, j) * (F(u, i - 1, j + 1) + F(u, i, j + 1) + F(u, i - 1, j) + F(u, i, j)) / 4.0
                            - F(xet, i, j) * F(v, i, j);
        const double vcij1 = F(xzi, i, j + 1) * (F(v, i, j + 1) + F(v, i + 1, j + 1) + F(v, i, j) + F(v, i + 1, j)) / 4.0
                             - F(yzi, i, j + 1) * F(u, i, j + 1);
        const double vcij = F(xzi, i, j) * (F(v, i, j) + F(v, i + 1, j) + F(v, i, j - 1) + F(v, i + 1, j - 1)) / 4.0
                            - F(yzi, i, j) * F(u, i, j);
        return uci1j - ucij + vcij1 - vcij;
    }
#undef F
}

template <int NLOC>
__global__ void __launch_bounds__(256) divergence_kernel(int nx, int ny, int pitch, const double *__restrict__ xet,
                                                         const double *__restrict__ yet, const double *__restrict__ xzi,
                                                         const double *__restrict__ yzi, const double *__restrict__ u,
                                                         const double *__restrict__ v, double *__restrict__ div) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    for (int j = 1 + blockIdx.y; j <= ny; j += gridDim.y)
        if (i <= nx) div[IDX(i, j)] = div_point<NLOC>(i, j, pitch, xet, yet, xzi, yzi, u, v);
}

int w2_divergence(wolfd2_ctx *c, const double *u, const double *v, double *div, int nloc, const double *xet,
                  const double *yet, const double *xzi, const double *yzi) {
    dim3 grid((c->nx + 255) / 256, c->ny < 4096 ? c->ny : 4096);
    if (nloc == 1)
        divergence_kernel<1><<<grid, 256, 0, c->stream>>>(c->nx, c->ny, c->pitch, xet, yet, xzi, yzi, u, v, div);
    else if (nloc == 2)
        divergence_kernel<2><<<grid, 256, 0, c->stream>>>(c->nx, c->ny, c->pitch, xet, yet, xzi, yzi, u, v, div);
    else {
        w2_set_error("Error: Wrong location flag passed to Divergence: %d", nloc);  // :316-319
        return W2_ERR_BAD_ARG;
    }
    c->launches[2]++;
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}

// Fused Divergence(nloc=1) + blockage zeroing + RhsPpe (Cartesian part): b = div/dk on 2..nx,2..ny.
// If div_out != nullptr the masked divergence is also stored (needed when the grid is not
// Cartesian and b is rebuilt every sweep, :425-427).
__global__ void __launch_bounds__(256) div_rhs_kernel(int nx, int jlo, int jhi, int pitch, double dk,
                                                      const double *__restrict__ xeu, const double *__restrict__ yeu,
                                                      const double *__restrict__ xzv, const double *__restrict__ yzv,
                                                      const double *__restrict__ u, const double *__restrict__ v,
                                                      const unsigned char *__restrict__ mask, int has_mask,
                                                      int sentinel, double *__restrict__ b,
                                                      double *__restrict__ div_out) {
    const int i = 2 + blockIdx.x * blockDim.x + threadIdx.x;
    for (int j = jlo + blockIdx.y; j <= jhi; j += gridDim.y) {
        if (i > nx) continue;
        double d = div_point<1>(i, j, pitch, xeu, yeu, xzv, yzv, u, v);
        const bool masked = has_mask && mask[IDX(i, j)];
        if (masked) d = 0.0;
        if (div_out) div_out[IDX(i, j)] = d;
        // the fused SOR kernel recognises identity rows by a NaN in b (it then uses b = 0, a = identity)
        // (sentinel mode also means: b is consumed by the fused kernel, which wants colour-split rows)
        const size_t bo = sentinel ? (size_t)pitch * j + (size_t)(i & 1) * (pitch >> 1) + (i >> 1) : IDX(i, j);
        b[bo] = (masked && sentinel) ? __longlong_as_double(0x7ff8000000000000LL) : d / dk;
    }
}

// RhsPpe for a non-Cartesian grid (:354-375): b = div/dk - cross-derivative terms of p.
__global__ void __launch_bounds__(256) rhs_cross_kernel(int nx, int ny, int pitch, double dk,
                                                        const double *__restrict__ rbu, const double *__restrict__ rbv,
                                                        const double *__restrict__ div, const double *__restrict__ p,
                                                        double *__restrict__ b, const int *__restrict__ done) {
    if (done && *done) return;
    const int i = 2 + blockIdx.x * blockDim.x + threadIdx.x;
    for (int j = 2 + blockIdx.y; j <= ny; j += gridDim.y) {
        if (i > nx) continue;
        double bb = div[IDX(i, j)] / dk;
        bb = bb - (rbu[IDX(i, j)] * (P(i + 1, j + 1) + P(i, j + 1) - P(i + 1, j - 1) - P(i, j - 1))
                   - rbu[IDX(i - 1, j)] * (P(i, j + 1) + P(i - 1, j + 1) - P(i, j - 1) - P(i - 1, j - 1))
                   + rbv[IDX(i, j)] * (P(i + 1, j + 1) + P(i + 1, j) - P(i - 1, j + 1) - P(i - 1, j))
                   - rbv[IDX(i, j - 1)] * (P(i + 1, j) + P(i + 1, j - 1) - P(i - 1, j) - P(i - 1, j - 1)));
        b[IDX(i, j)] = bb;
    }
}

// RhsPpe as a unit (rhsppe_ shim): b(ind), ind = (j-2)*(nx-1)+i-1 (:347-352), from a given div.  The non-Cartesian
// branch is the production rhs_cross_kernel; this kernel only gathers its field-layout result into the
// reference's vector ordering (Cartesian grids: b = div/dk, the statement div_rhs_kernel fuses with Divergence).
__global__ void __launch_bounds__(256) unit_rhs_gather_kernel(int nx, int ny, int pitch, int cartes, double dk,
                                                              const double *__restrict__ div, const double *__restrict__ bf,
                                                              double *__restrict__ bvec) {
    const int i = 2 + blockIdx.x * blockDim.x + threadIdx.x, j = 2 + blockIdx.y;
    if (i > nx || j > ny) return;
    const size_t ind = (size_t)(j - 2) * (size_t)(nx - 1) + (size_t)(i - 2);
    bvec[ind] = cartes ? div[IDX(i, j)] / dk : bf[IDX(i, j)];
}
int w2_unit_rhsppe(wolfd2_ctx *c, int cartes, double dk, const double *rbu, const double *rbv, const double *div, const double *p,
                   double *bfield, double *bvec) {
    dim3 g((c->nx - 1 + 255) / 256, c->ny - 1);
    if (!cartes) rhs_cross_kernel<<<g, 256, 0, c->stream>>>(c->nx, c->ny, c->pitch, dk, rbu, rbv, div, p, bfield, nullptr);
    unit_rhs_gather_kernel<<<g, 256, 0, c->stream>>>(c->nx, c->ny, c->pitch, cartes, dk, div, bfield, bvec);
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}

// ------------------------------------------------------------------ SOR control block
// ctl[0]=done, ctl[1]=iterations completed (m), ctl[2]=nConv, ctl[3]=CTA ticket counter
// slot: running max |sum| of the current iteration (bit pattern of a non-negative double).
struct SorCtl {
    int done, m, nconv, ticket;
    unsigned long long slot;
    unsigned long long last_dif;
};

__global__ void sor_ctl_reset(SorCtl *ctl) {
    ctl->done = 0; ctl->m = 0; ctl->nconv = 0; ctl->ticket = 0; ctl->slot = 0ull; ctl->last_dif = 0ull;
}

// One colour half-sweep of SorRB (:505-517 black = i+j even, :520-532 red), in place.
// Each warp owns rows; lanes stride over the cells of the active colour.
template <bool HAS_MASK>
__global__ void __launch_bounds__(256) sor_rb_sweep(int nx, int ny, int pitch, int colour /*0 black, 1 red*/,
                                                    double sorrel, double sortol, int msorit,
                                                    const double *__restrict__ rau, const double *__restrict__ rgv,
                                                    const double *__restrict__ b, const unsigned char *__restrict__ mask,
                                                    double *p, SorCtl *ctl) {
    if (ctl->done) return;
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    double lmax = 0.0;
    for (int j = 2 + blockIdx.x * nwb + wib; j <= ny; j += gridDim.x * nwb) {
        // black: i = 2+mod(j,2), step 2; red: i = 2+mod(j+1,2)
        const int istart = 2 + ((j + colour) & 1);
        const size_t row = (size_t)pitch * (size_t)j;
        for (int i = istart + 2 * lane; i <= nx; i += 64) {
            const size_t c0 = row + i;
            double a1 = rgv[c0 - pitch], a2 = rau[c0 - 1], a4 = rau[c0], a5 = rgv[c0];
            double a3 = -a4 - a2 - a5 - a1;
            if (HAS_MASK && mask[c0]) { a1 = 0.0; a2 = 0.0; a3 = 1.0; a4 = 0.0; a5 = 0.0; }
            const double pc = p[c0];
            double sum = b[c0] - a1 * p[c0 - pitch] - a2 * p[c0 - 1] - a4 * p[c0 + 1] - a5 * p[c0 + pitch];
            sum = w2_div_exact(sum, a3) - pc;
            p[c0] = pc + sorrel * sum;
            lmax = fmax(lmax, fabs(sum));
        }
    }
    lmax = w2_block_max(lmax, red);
    if (threadIdx.x == 0) {
        atomicMax(&ctl->slot, w2_dbits(lmax));
        if (colour == 1) {  // close the iteration when the last CTA of the red sweep arrives
            __threadfence();
            const int t = atomicAdd(&ctl->ticket, 1);
            if (t == (int)gridDim.x - 1) {
                __threadfence();
                const unsigned long long bits = atomicExch(&ctl->slot, 0ull);
                const double dif = __longlong_as_double((long long)bits);
                const int m = ctl->m + 1;
                ctl->m = m;
                ctl->last_dif = bits;
                ctl->ticket = 0;
                if (m > 1 && dif < sortol) { ctl->nconv = m; ctl->done = 1; }
                else if (m >= msorit) { ctl->done = 1; }
                __threadfence();
            }
        }
    }
}

static int sor_chunk(const wolfd2_ctx *c) {
    const double cells = (double)(c->nx - 1) * (double)(c->ny - 1);
    const double t_iter = cells * 80.0 / 5.0e12 + 8.0e-6;  // rough: HBM time + 2 launches
    int chunk = (int)(2.0e-3 / t_iter);
    if (chunk < 8) chunk = 8;
    if (chunk > 256) chunk = 256;
    return chunk;
}

// Ppe (:30-249) with nPpeSolver 5 or 6 (SorRB / SorRBP: same update, same max-norm).
int w2_ppe_line_sor(wolfd2_ctx *c, double *p, int *nSorConv, int *converged, int *iters_done);
int w2_ppe_lex_sor(wolfd2_ctx *c, double *p, int *nSorConv, int *converged, int *iters_done);

// RhsPpe on a non-Cartesian grid, for the solvers in w2_ppe_other.cu
void rhs_cross_launch(wolfd2_ctx *c, double *p) {
    dim3 g2((c->nx - 1 + 255) / 256, (c->ny - 1) < 2048 ? (c->ny - 1) : 2048);
    rhs_cross_kernel<<<g2, 256, 0, c->stream>>>(c->nx, c->ny, c->pitch, c->par.dk, c->met.rbu, c->met.rbv, c->div, p,
                                                c->fld[W2_F_B], nullptr);
    c->launches[2]++;
}

int w2_sor_fused(wolfd2_ctx *c, double *p, double *scratch, int T, int *nSorConv, int *converged, double **p_final,
                 int *iters_done);
int w2_sor_fused_result(wolfd2_ctx *c, int *nSorConv, int *converged, int *iters_done);
bool w2_sor_resident_ok(const wolfd2_ctx *c);
int w2_sor_resident(wolfd2_ctx *c, double *p);

// Outcome and device time of the last fused solve (the stream has been synchronised since it was enqueued).
int w2_sor_collect(wolfd2_ctx *c, int *nSorConv, int *converged) {
    if (!c->sor_pending) {
        if (c->sor_saved[2]) {   // collected early (another Ppe of the same step needed the events)
            if (nSorConv) *nSorConv = c->sor_saved[0];
            if (converged) *converged = c->sor_saved[1];
            c->sor_saved[2] = 0;
        }
        return W2_OK;
    }
    c->sor_pending = 0;
    int iters = 0;
    W2_TRY(w2_sor_fused_result(c, nSorConv, converged, &iters));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]);
    c->sor_ms += ms;
    c->sor_iters += iters;
    return W2_OK;
}

int g_sor_T = -1;   // set by wolfd2_b200_set_option("sor_fused_T", t); -1: take W2_SOR_T or the default
static int fused_T() {   // 0 selects the plain half-sweep kernels, 1 or 2 the fused pipeline
    if (g_sor_T >= 0) return g_sor_T;
    const char *e = getenv("W2_SOR_T");
    int t = e ? atoi(e) : 2;
    if (t < 0 || t > 2) t = 2;
    return t;
}

static int ppe_impl(wolfd2_ctx *c, const double *u, const double *v, double *p, int *nSorConv, int *converged, bool deferred);

int w2_ppe(wolfd2_ctx *c, const double *u, const double *v, double *p, int *nSorConv, int *converged) {
    // deferred outcome (both pointers null): the fused path leaves it pending for w2_sor_collect; the other solvers
    // know it at once and park it where w2_sor_collect looks
    const bool deferred = nSorConv == nullptr && converged == nullptr;
    int n = 0, cv = 0;
    W2_TRY(ppe_impl(c, u, v, p, deferred ? &n : nSorConv, deferred ? &cv : converged, deferred));
    if (deferred && !c->sor_pending) { c->sor_saved[0] = n; c->sor_saved[1] = cv; c->sor_saved[2] = 1; }
    return W2_OK;
}

static int ppe_impl(wolfd2_ctx *c, const double *u, const double *v, double *p, int *nSorConv, int *converged, bool deferred) {
    const wolfd2_params &par = c->par;
    if (par.nPpeSolver < 1 || par.nPpeSolver > 6) {
        w2_set_error("Wrong nPpeSolver flag passed to Ppe: %d", par.nPpeSolver);   // :238-239
        return W2_ERR_BAD_ARG;
    }
    const bool rb_point = par.nPpeSolver == W2_PPE_RB_SOR || par.nPpeSolver == W2_PPE_PAR_RB_SOR;
    const int nx = c->nx, ny = c->ny, pitch = c->pitch;
    const int cart = par.lCartesGrid != 0;
    const int has_mask = c->hreg.has_blockage;
    double *b = c->fld[W2_F_B];
    SorCtl *ctl = (SorCtl *)c->d_flags;
    static_assert(sizeof(SorCtl) <= 64 * sizeof(int), "ctl block too large");

    int T = (rb_point && cart && nx >= 254 && ny >= 8) ? fused_T() : 0;
    // small grids (1024^2 and below): the whole solve in one cooperative launch, p resident in shared memory
    const bool resident = rb_point && cart && fused_T() > 0 && w2_sor_resident_ok(c);
    if (resident) T = 0;
    if (c->world > 1) {
        if (!(rb_point && cart && nx >= 254)) {
            w2_set_error("multi-GPU runs support ppe_solver 5/6 on a Cartesian grid with nx >= 254 only");
            return W2_ERR_UNSUPPORTED;
        }
        if (T == 0) T = 2;   // the plain half-sweep kernels have no slab path
    }
    dim3 g2((nx - 1 + 255) / 256, (ny - 1) < 2048 ? (ny - 1) : 2048);
    // every held row whose stencil is held too: on several GPUs this covers the halo rows the fused SOR
    // pass reads (their us, vs are valid), so b needs no exchange of its own
    const int bj0 = c->A0 + 1 > 2 ? c->A0 + 1 : 2, bj1 = c->A1 - 1 < ny ? c->A1 - 1 : ny;
    div_rhs_kernel<<<g2, 256, 0, c->stream>>>(nx, bj0, bj1, pitch, par.dk, c->met.xeu, c->met.yeu, c->met.xzv, c->met.yzv, u, v,
                                              c->pmask, has_mask, T > 0, b, cart ? nullptr : c->div);
    c->launches[2]++;
    if (resident) {
        if (c->sor_pending) {
            W2_CUDA(cudaStreamSynchronize(c->stream));
            c->host_syncs++;
            W2_TRY(w2_sor_collect(c, &c->sor_saved[0], &c->sor_saved[1]));
            c->sor_saved[2] = 1;
        }
        cudaEventRecord(c->ev[4], c->stream);
        W2_TRY(w2_sor_resident(c, p));
        cudaEventRecord(c->ev[5], c->stream);
        c->sor_pending = 1;
        if (!deferred) {
            W2_CUDA(cudaStreamSynchronize(c->stream));
            c->host_syncs++;
            W2_TRY(w2_sor_collect(c, nSorConv, converged));
        }
        return W2_OK;
    }
    if (T > 0) {
        // fused red/black pipeline (w2_sor_fused.cu); c->div is free on Cartesian grids and serves as
        // the second pressure buffer
        double *pf = nullptr;
        int iters = 0;
        if (c->sor_pending) {   // an earlier solve of this step (the small-scale model's own Ppe) still owns the events
            W2_CUDA(cudaStreamSynchronize(c->stream));
            c->host_syncs++;
            W2_TRY(w2_sor_collect(c, &c->sor_saved[0], &c->sor_saved[1]));
            c->sor_saved[2] = 1;
        }
        cudaEventRecord(c->ev[4], c->stream);
        W2_TRY(w2_sor_fused(c, p, c->div, T, nSorConv, converged, &pf, &iters));
        if (pf != p) W2_TRY(w2_copy_field(c, p, pf));
        cudaEventRecord(c->ev[5], c->stream);
        c->sor_pending = 1;
        if (!deferred) {   // the caller wants the outcome now
            W2_CUDA(cudaStreamSynchronize(c->stream));
            c->host_syncs++;
            W2_TRY(w2_sor_collect(c, nSorConv, converged));
        }
        return W2_OK;
    }
    if (!rb_point) {   // ids 1-4: w2_ppe_other.cu
        int iters = 0, conv = 0, nconv = par.msorit;
        cudaEventRecord(c->ev[4], c->stream);
        if (par.nPpeSolver == W2_PPE_SOR) W2_TRY(w2_ppe_lex_sor(c, p, &nconv, &conv, &iters));
        else W2_TRY(w2_ppe_line_sor(c, p, &nconv, &conv, &iters));
        cudaEventRecord(c->ev[5], c->stream);
        W2_CUDA(cudaStreamSynchronize(c->stream));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]);
        c->sor_ms += ms;
        c->sor_iters += iters;
        if (converged) *converged = conv;
        if (nSorConv) *nSorConv = nconv;
        return W2_OK;
    }
    sor_ctl_reset<<<1, 1, 0, c->stream>>>(ctl);
    W2_CUDA(cudaGetLastError());

    // grid: enough CTAs to fill the machine, one warp per row
    const int nwb = 8;
    int blocks = (ny - 1 + nwb - 1) / nwb;
    const int maxb = c->num_sms * 8;
    if (blocks > maxb) blocks = maxb;
    const int chunk = sor_chunk(c);
    SorCtl h;
    memset(&h, 0, sizeof(h));
    cudaEventRecord(c->ev[4], c->stream);
    int queued = 0;
    while (queued < par.msorit) {
        const int n = (par.msorit - queued) < chunk ? (par.msorit - queued) : chunk;
        for (int k = 0; k < n; ++k) {
            if (!cart) {
                rhs_cross_kernel<<<g2, 256, 0, c->stream>>>(nx, ny, pitch, par.dk, c->met.rbu, c->met.rbv, c->div, p, b,
                                                            &ctl->done);
                c->launches[2]++;
            }
            for (int colour = 0; colour < 2; ++colour) {
                if (has_mask)
                    sor_rb_sweep<true><<<blocks, nwb * 32, 0, c->stream>>>(nx, ny, pitch, colour, par.sorrel, par.sortol,
                                                                           par.msorit, c->met.rau, c->met.rgv, b, c->pmask, p, ctl);
                else
                    sor_rb_sweep<false><<<blocks, nwb * 32, 0, c->stream>>>(nx, ny, pitch, colour, par.sorrel, par.sortol,
                                                                            par.msorit, c->met.rau, c->met.rgv, b, c->pmask, p, ctl);
                c->launches[2]++;
            }
        }
        queued += n;
        W2_CUDA(cudaGetLastError());
        W2_CUDA(cudaMemcpyAsync(c->h_flags, ctl, sizeof(SorCtl), cudaMemcpyDeviceToHost, c->stream));
        W2_CUDA(cudaStreamSynchronize(c->stream));
        memcpy(&h, c->h_flags, sizeof(SorCtl));
        if (h.done) break;
    }
    cudaEventRecord(c->ev[5], c->stream);
    W2_CUDA(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]);
    c->sor_ms += ms;
    c->sor_iters += h.m;
    // :242-246: if not converged, nSorConv = msorit (and a warning on stdout in the reference)
    if (converged) *converged = h.nconv > 0;
    if (nSorConv) *nSorConv = h.nconv > 0 ? h.nconv : par.msorit;
    return W2_OK;
}
