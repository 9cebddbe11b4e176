// w2_project.cu -- Project (src/utility.f:253-440), Filter (:33-247), DiffMaxNorm (:446-473),
// DMaxNorm (:479-507) and whole-field copies.
#include "w2.cuh"

#define P(i, j) p[IDX(i, j)]

// One launch per non-blockage region.  The reference updates the region interior and, depending
// on the face type, the W (OUTLT1 only) and E (INTERN or OUTLT1) columns for u (:314-363) and the
// S / N rows for v (:384-432); the face loops use the same formula as the interior, so they are
// folded into the index range [ulo..uhi] x [jS+1..jN] and [iW+1..iE] x [vlo..vhi].
__global__ void __launch_bounds__(256) project_region_kernel(int pitch, double dk, int ulo, int uhi, int ujlo, int ujhi,
                                                             int vilo, int vihi, int vlo, int vhi,
                                                             const double *__restrict__ dju, const double *__restrict__ djv,
                                                             const double *__restrict__ yeu, const double *__restrict__ xzv,
                                                             const double *__restrict__ yzu, const double *__restrict__ xev,
                                                             const double *__restrict__ p, double *__restrict__ u,
                                                             double *__restrict__ v, int jbase) {
    const double dFour = 4.0;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // absolute i, starts at 0
    for (int j = jbase + blockIdx.y; j <= (ujhi > vhi ? ujhi : vhi); j += gridDim.y) {
        if (i >= ulo && i <= uhi && j >= ujlo && j <= ujhi) {
            const double pzi = P(i + 1, j) - P(i, j);
            const double pet = (P(i + 1, j + 1) + P(i, j + 1) - P(i + 1, j - 1) - P(i, j - 1)) / dFour;
            const double djk = dju[IDX(i, j)] * dk;
            u[IDX(i, j)] = u[IDX(i, j)] - djk * (yeu[IDX(i, j)] * pzi - yzu[IDX(i, j)] * pet);
        }
        if (i >= vilo && i <= vihi && j >= vlo && j <= vhi) {
            const double pzi = (P(i + 1, j + 1) + P(i + 1, j) - P(i - 1, j + 1) - P(i - 1, j)) / dFour;
            const double pet = P(i, j + 1) - P(i, j);
            const double djk = djv[IDX(i, j)] * dk;
            v[IDX(i, j)] = v[IDX(i, j)] - djk * (-xev[IDX(i, j)] * pzi + xzv[IDX(i, j)] * pet);
        }
    }
}

int w2_project(wolfd2_ctx *c, const double *p, double *u, double *v) {
    const W2Regions &R = c->hreg;
    for (int q = 0; q < R.nreg; ++q) {
        if (R.type[q] == W2_RM_BLOCKG) continue;  // :305, :375
        const int iW = R.iW[q], iE = R.iE[q], jS = R.jS[q], jN = R.jN[q];
        const int bW = R.bd[q][W2_WEST - 1], bE = R.bd[q][W2_EAST - 1];
        const int bS = R.bd[q][W2_SOUTH - 1], bN = R.bd[q][W2_NORTH - 1];
        const int ulo = (bW == W2_BM_OUTLT1) ? iW : iW + 1;
        const int uhi = (bE == W2_BM_INTERN || bE == W2_BM_OUTLT1) ? iE : iE - 1;
        const int vlo = (bS == W2_BM_OUTLT1) ? jS : jS + 1;
        const int vhi = (bN == W2_BM_INTERN || bN == W2_BM_OUTLT1) ? jN : jN - 1;
        // rows of this rank only (all rows on one GPU)
        int ujlo = jS + 1, ujhi = jN, vjlo = vlo, vjhi = vhi;
        w2_clip(c, ujlo, ujhi);
        w2_clip(c, vjlo, vjhi);
        if (ujhi < ujlo && vjhi < vjlo) continue;
        const int jbase = (vjhi < vjlo || ujlo < vjlo) && ujhi >= ujlo ? ujlo : vjlo;
        const int jtop = ujhi > vjhi ? ujhi : vjhi;
        int gy = jtop - jbase + 1;
        if (gy > 4096) gy = 4096;
        if (gy < 1) gy = 1;
        dim3 grid((c->nx + 2 + 255) / 256, gy);
        project_region_kernel<<<grid, 256, 0, c->stream>>>(c->pitch, c->par.dk, ulo, uhi, ujlo, ujhi, iW + 1, iE, vjlo, vjhi,
                                                           c->met.dju, c->met.djv, c->met.yeu, c->met.xzv, c->met.yzu,
                                                           c->met.xev, p, u, v, jbase);
        c->launches[3]++;
    }
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}

// ---------------------------------------------------------------------------- Filter
// Shuman filter of one region into qh (which starts as a copy of qu): interior plus the face
// columns/rows the reference filters (:94-139 for u, :157-203 for v), again one formula.
__global__ void __launch_bounds__(256) filter_region_kernel(int pitch, double fp, int ilo, int ihi, int jlo, int jhi,
                                                            const double *__restrict__ qu, double *__restrict__ qh) {
    const double dFour = 4.0;
    const int i = ilo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > ihi) return;
    for (int j = jlo + blockIdx.y; j <= jhi; j += gridDim.y)
        qh[IDX(i, j)] = (qu[IDX(i, j - 1)] + qu[IDX(i - 1, j)] + qu[IDX(i, j + 1)] + qu[IDX(i + 1, j)]
                         + fp * qu[IDX(i, j)]) / (fp + dFour);
}

int w2_filter(wolfd2_ctx *c, int ncomp, double fp, double *qu) {
    if (ncomp != W2_U && ncomp != W2_V && ncomp != W2_T) {
        w2_set_error("Wrong ncomp flag passed to Filter: %d", ncomp);  // :232-234
        return W2_ERR_BAD_ARG;
    }
    if (!c->qh) W2_TRY(w2_alloc_field(c, &c->qh));
    W2_TRY(w2_copy_field(c, c->qh, qu));  // :72-76
    const W2Regions &R = c->hreg;
    for (int q = 0; q < R.nreg; ++q) {
        if (ncomp == W2_T ? c->hth.ttype[q] == W2_BT_TEMPER /* sic: a BT_ constant against nTRgType, :212 */
                          : R.type[q] == W2_RM_BLOCKG) continue;
        const int iW = R.iW[q], iE = R.iE[q], jS = R.jS[q], jN = R.jN[q];
        int ilo, ihi, jlo, jhi;
        if (ncomp == W2_T) {   // :221-226
            ilo = iW + 1; ihi = iE; jlo = jS + 1; jhi = jN;
        } else if (ncomp == W2_U) {
            const int bW = R.bd[q][W2_WEST - 1], bE = R.bd[q][W2_EAST - 1];
            ilo = (bW == W2_BM_OUTLT1) ? iW : iW + 1;
            ihi = (bE == W2_BM_INTERN || bE == W2_BM_OUTLT1) ? iE : iE - 1;
            jlo = jS + 1; jhi = jN;
        } else {
            const int bS = R.bd[q][W2_SOUTH - 1], bN = R.bd[q][W2_NORTH - 1];
            ilo = iW + 1; ihi = iE;
            jlo = (bS == W2_BM_OUTLT1) ? jS : jS + 1;
            jhi = (bN == W2_BM_INTERN || bN == W2_BM_OUTLT1) ? jN : jN - 1;
        }
        w2_clip(c, jlo, jhi);
        if (ihi < ilo || jhi < jlo) continue;
        int gy = jhi - jlo + 1;
        if (gy > 4096) gy = 4096;
        dim3 grid((ihi - ilo + 1 + 255) / 256, gy);
        filter_region_kernel<<<grid, 256, 0, c->stream>>>(c->pitch, fp, ilo, ihi, jlo, jhi, qu, c->qh);
        c->launches[3]++;
    }
    W2_TRY(w2_copy_field(c, qu, c->qh));  // :238-243
    W2_CUDA(cudaGetLastError());
    if (c->world > 1) {   // the filtered field's halo rows come from their owners
        double *f[1] = {qu};
        W2_TRY(w2_halo_exchange(c, f, 1, c->HG));
    }
    return W2_OK;
}

int w2_copy_field(wolfd2_ctx *c, double *dst, const double *src) {
    W2_CUDA(cudaMemcpyAsync(dst + c->row_off, src + c->row_off, c->nelem * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    return W2_OK;
}

// ---------------------------------------------------------------------------- max-norms
// DIFF: max |a-b|, seed (2,2) (:459); else max |a|, seed (5,5) (:493); both scan 2..nx-1,2..ny-1.
template <bool DIFF>
__global__ void __launch_bounds__(256) maxnorm_kernel(int nx, int jlo, int jhi, int seed, int pitch, const double *__restrict__ a,
                                                      const double *__restrict__ b, unsigned long long *slot) {
    __shared__ double red[32];
    double m = 0.0;
    const int i = 2 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= nx - 1)
        for (int j = jlo + blockIdx.y; j <= jhi; j += gridDim.y) {
            const double x = DIFF ? fabs(a[IDX(i, j)] - b[IDX(i, j)]) : fabs(a[IDX(i, j)]);
            m = fmax(m, x);
        }
    if (seed && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        const double s = DIFF ? fabs(a[IDX(2, 2)] - b[IDX(2, 2)]) : fabs(a[IDX(5, 5)]);
        m = fmax(m, s);
    }
    m = w2_block_max(m, red);
    if (threadIdx.x == 0) atomicMax(slot, w2_dbits(m));
}

int w2_norm_reset(wolfd2_ctx *c) {
    W2_CUDA(cudaMemsetAsync(c->d_norm, 0, 16 * sizeof(unsigned long long), c->stream));
    return W2_OK;
}
int w2_diffmaxnorm_async(wolfd2_ctx *c, const double *a, const double *b, int slot) {
    int jlo = 2, jhi = c->ny - 1;
    w2_clip(c, jlo, jhi);
    int gy = jhi - jlo + 1;
    if (gy > 1024) gy = 1024;
    if (gy < 1) gy = 1;
    dim3 grid((c->nx - 2 + 255) / 256, gy);
    maxnorm_kernel<true><<<grid, 256, 0, c->stream>>>(c->nx, jlo, jhi, c->rank == 0, c->pitch, a, b, c->d_norm + slot);
    c->launches[3]++;
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}
int w2_dmaxnorm_async(wolfd2_ctx *c, const double *a, int slot) {
    int jlo = 2, jhi = c->ny - 1;
    w2_clip(c, jlo, jhi);
    int gy = jhi - jlo + 1;
    if (gy > 1024) gy = 1024;
    if (gy < 1) gy = 1;
    dim3 grid((c->nx - 2 + 255) / 256, gy);
    maxnorm_kernel<false><<<grid, 256, 0, c->stream>>>(c->nx, jlo, jhi, c->rank == 0, c->pitch, a, a, c->d_norm + slot);
    c->launches[3]++;
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}
int w2_norm_fetch(wolfd2_ctx *c, int nslots, double *out) {
    W2_TRY(w2_allreduce_max_u64(c, c->d_norm, nslots));
    W2_CUDA(cudaMemcpyAsync(c->h_norm, c->d_norm, nslots * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    W2_CUDA(cudaStreamSynchronize(c->stream));
    c->host_syncs++;
    for (int k = 0; k < nslots; ++k) memcpy(&out[k], &c->h_norm[k], 8);
    return W2_OK;
}
