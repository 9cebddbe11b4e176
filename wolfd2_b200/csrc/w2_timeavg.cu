// w2_timeavg.cu -- the time-averaging mode of the reference (-D_TIMEAVG_) on the resident fields:
//   begin       src/main.f:510-541    nineteen arrays zeroed on 0..nx+1, 0..ny+1
//   accumulate  src/main.f:1107-1208  at the end of every time step of pass 1 (means of the node averages of
//                                     u, v, t, p) or pass 2 (fluctuations about those means, their products, the
//                                     squared gradients of the velocity fluctuations, dT/dy)
//   finish      src/main.f:1239-1297  division by the number of steps, turbulence kinetic energy, dissipation rate
// The reference runs the simulation twice (nTimeAvg = 1, 2); the caller does the same (restart, second run).
// Output: wolfd2_b200_timeavg_get, in the plane order of SaveTmAvgP3D (wolfd2_b200/plot3d.py writes the file).
#include <stdlib.h>
#include <string.h>

#include "w2.cuh"

enum { TA_UBAR, TA_VBAR, TA_TBAR, TA_PBAR, TA_UPB, TA_VPB, TA_TPB, TA_UPUPB, TA_VPVPB, TA_UPVPB, TA_UPTPB, TA_VPTPB,
       TA_UPXSB, TA_UPYSB, TA_VPXSB, TA_VPYSB, TA_TRBKE, TA_DSSRT, TA_DTDYB, TA_COUNT };

struct W2TimeAvg {
    int pass;                 // 0: not accumulating, 1 / 2
    double *a[TA_COUNT];      // field layout, zero outside what the loops write
};

struct TaPtrs { double *a[TA_COUNT]; };

// :1126-1133
__global__ void __launch_bounds__(256) tavg_pass1_kernel(int nx, int ny, int pitch, const double *__restrict__ us,
                                                         const double *__restrict__ vs, const double *__restrict__ ts,
                                                         const double *__restrict__ pn, TaPtrs A) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx) return;
    for (int j = 1 + blockIdx.y; j <= ny; j += gridDim.y) {
        const size_t o = IDX(i, j);
        A.a[TA_UBAR][o] = A.a[TA_UBAR][o] + us[o];
        A.a[TA_VBAR][o] = A.a[TA_VBAR][o] + vs[o];
        A.a[TA_TBAR][o] = A.a[TA_TBAR][o] + ts[o];
        A.a[TA_PBAR][o] = A.a[TA_PBAR][o] + pn[o];
    }
}

// :1146-1204
__global__ void __launch_bounds__(256) tavg_pass2_kernel(int nx, int ny, int pitch, const double *__restrict__ u,
                                                         const double *__restrict__ v, const double *__restrict__ t,
                                                         const double *__restrict__ us, const double *__restrict__ vs,
                                                         const double *__restrict__ ts, const double *__restrict__ djn,
                                                         const double *__restrict__ xen, const double *__restrict__ yen,
                                                         const double *__restrict__ xzn, const double *__restrict__ yzn,
                                                         TaPtrs A) {
    const double dQrtr = 0.25, dHalf = 0.5;
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx) return;
    const double *ubar = A.a[TA_UBAR], *vbar = A.a[TA_VBAR], *tbar = A.a[TA_TBAR];
#define AT(f, ii, jj) f[IDX(ii, jj)]
    for (int j = 1 + blockIdx.y; j <= ny; j += gridDim.y) {
        const size_t o = IDX(i, j);
        const double dupij = us[o] - ubar[o], dvpij = vs[o] - vbar[o], dtpij = ts[o] - tbar[o];
        A.a[TA_UPB][o] = A.a[TA_UPB][o] + dupij;
        A.a[TA_VPB][o] = A.a[TA_VPB][o] + dvpij;
        A.a[TA_TPB][o] = A.a[TA_TPB][o] + dtpij;
        A.a[TA_UPUPB][o] = A.a[TA_UPUPB][o] + dupij * dupij;
        A.a[TA_VPVPB][o] = A.a[TA_VPVPB][o] + dvpij * dvpij;
        A.a[TA_UPVPB][o] = A.a[TA_UPVPB][o] + dupij * dvpij;
        A.a[TA_UPTPB][o] = A.a[TA_UPTPB][o] + dupij * dtpij;
        A.a[TA_VPTPB][o] = A.a[TA_VPTPB][o] + dvpij * dtpij;
        const double dupipj = (AT(u, i + 1, j + 1) + AT(u, i + 1, j) + AT(u, i, j + 1) + AT(u, i, j)
                               - AT(ubar, i + 1, j + 1) - AT(ubar, i + 1, j) - AT(ubar, i, j + 1) - AT(ubar, i, j)) * dQrtr;
        const double dupimj = (AT(u, i, j + 1) + AT(u, i, j) + AT(u, i - 1, j + 1) + AT(u, i - 1, j)
                               - AT(ubar, i, j + 1) - AT(ubar, i, j) - AT(ubar, i - 1, j + 1) - AT(ubar, i - 1, j)) * dQrtr;
        const double dupijp = AT(u, i, j + 1) - AT(ubar, i, j + 1);
        const double dupijm = AT(u, i, j) - AT(ubar, i, j);
        const double dvpipj = AT(v, i + 1, j) - AT(vbar, i + 1, j);
        const double dvpimj = AT(v, i, j) - AT(vbar, i, j);
        const double dvpijp = (AT(v, i + 1, j + 1) + AT(v, i + 1, j) + AT(v, i, j + 1) + AT(v, i, j)
                               - AT(vbar, i + 1, j + 1) - AT(vbar, i + 1, j) - AT(vbar, i, j + 1) - AT(vbar, i, j)) * dQrtr;
        const double dvpijm = (AT(v, i + 1, j) + AT(v, i, j) + AT(v, i + 1, j - 1) + AT(v, i, j - 1)
                               - AT(vbar, i + 1, j) - AT(vbar, i, j) - AT(vbar, i + 1, j - 1) - AT(vbar, i, j - 1)) * dQrtr;
        const double upzi = dupipj - dupimj, upet = dupijp - dupijm, vpzi = dvpipj - dvpimj, vpet = dvpijp - dvpijm;
        const double dj = djn[o], xe = xen[o], ye = yen[o], xz = xzn[o], yz = yzn[o];
        const double dupx = dj * (ye * upzi - yz * upet), dupy = dj * (-xe * upzi + xz * upet);
        const double dvpx = dj * (ye * vpzi - yz * vpet), dvpy = dj * (-xe * vpzi + xz * vpet);
        A.a[TA_UPXSB][o] = A.a[TA_UPXSB][o] + dupx * dupx;
        A.a[TA_UPYSB][o] = A.a[TA_UPYSB][o] + dupy * dupy;
        A.a[TA_VPXSB][o] = A.a[TA_VPXSB][o] + dvpx * dvpx;
        A.a[TA_VPYSB][o] = A.a[TA_VPYSB][o] + dvpy * dvpy;
        const double tzi = dHalf * (AT(t, i + 1, j + 1) + AT(t, i + 1, j) - AT(t, i, j + 1) - AT(t, i, j));
        const double tet = dHalf * (AT(t, i + 1, j + 1) + AT(t, i, j + 1) - AT(t, i + 1, j) - AT(t, i, j));
        A.a[TA_DTDYB][o] = A.a[TA_DTDYB][o] + dj * (-xe * tzi + xz * tet);
    }
#undef AT
}

// :1246-1297
__global__ void __launch_bounds__(256) tavg_finish_kernel(int nx, int ny, int pitch, int pass, double dnts, double uref, double dlref,
                                                          double re, TaPtrs A) {
    const double dHalf = 0.5;
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx) return;
    for (int j = 1 + blockIdx.y; j <= ny; j += gridDim.y) {
        const size_t o = IDX(i, j);
        if (pass == 1) {
            for (int k = TA_UBAR; k <= TA_PBAR; ++k) A.a[k][o] = A.a[k][o] / dnts;
        } else {
            for (int k = TA_UPB; k <= TA_VPYSB; ++k) A.a[k][o] = A.a[k][o] / dnts;
            A.a[TA_DSSRT][o] = (A.a[TA_UPXSB][o] + A.a[TA_UPYSB][o] + A.a[TA_VPXSB][o] + A.a[TA_VPYSB][o]) * uref * dlref / re;   // :1285
            A.a[TA_TRBKE][o] = (A.a[TA_UPUPB][o] + A.a[TA_VPVPB][o]) * dHalf;
            A.a[TA_DTDYB][o] = A.a[TA_DTDYB][o] / dnts;
        }
    }
}

void w2_timeavg_release(wolfd2_ctx *c) {
    W2TimeAvg *q = (W2TimeAvg *)c->tavg;
    if (!q) return;
    for (int k = 0; k < TA_COUNT; ++k) if (q->a[k]) cudaFree(q->a[k] + c->row_off);
    free(q);
    c->tavg = nullptr;
}

static TaPtrs ptrs(const W2TimeAvg *q) {
    TaPtrs A;
    for (int k = 0; k < TA_COUNT; ++k) A.a[k] = q->a[k];
    return A;
}

// op 0: begin (allocate and zero the nineteen arrays); 1 / 2: accumulate pass 1 / 2 at the end of every following step;
// 3: stop accumulating; 4: release
extern "C" int wolfd2_b200_timeavg(wolfd2_ctx *c, int32_t op) {
    if (!c) return W2_ERR_BAD_ARG;
    W2_CUDA(cudaSetDevice(c->device));
    if (c->world > 1) { w2_set_error("time averaging is not supported on several GPUs"); return W2_ERR_UNSUPPORTED; }
    W2TimeAvg *q = (W2TimeAvg *)c->tavg;
    if (op == 4) { w2_timeavg_release(c); return W2_OK; }
    if (op == 0) {
        if (!q) {
            q = (W2TimeAvg *)calloc(1, sizeof(W2TimeAvg));
            if (!q) return W2_ERR_BAD_ARG;
            c->tavg = q;
            for (int k = 0; k < TA_COUNT; ++k) W2_TRY(w2_alloc_field(c, &q->a[k]));
        } else {
            for (int k = 0; k < TA_COUNT; ++k)
                W2_CUDA(cudaMemsetAsync(q->a[k] + c->row_off, 0, c->nelem * sizeof(double), c->stream));
        }
        q->pass = 0;
        return W2_OK;
    }
    if (!q) { w2_set_error("timeavg: call op 0 (begin) first"); return W2_ERR_BAD_ARG; }
    if (op < 1 || op > 3) { w2_set_error("timeavg: unknown op %d", op); return W2_ERR_BAD_ARG; }
    q->pass = op == 3 ? 0 : op;
    return W2_OK;
}

// end of a time step (main.f:1107-1208).  As in the reference the node averages land in the starred arrays us, vs,
// ts (and, in pass 1, pn for the pressure).
int w2_timeavg_step(wolfd2_ctx *c) {
    W2TimeAvg *q = (W2TimeAvg *)c->tavg;
    if (!q || !q->pass) return W2_OK;
    double *us = c->fld[W2_F_US], *vs = c->fld[W2_F_VS], *ts = c->fld[W2_F_TS], *pn = c->fld[W2_F_PN];
    W2_TRY(w2_velavg(c, c->fld[W2_F_U], c->fld[W2_F_V], us, vs));
    W2_TRY(w2_taveraged(c, 0, c->fld[W2_F_T], ts));
    dim3 g((c->nx + 255) / 256, c->ny < 2048 ? c->ny : 2048);
    if (q->pass == 1) {
        W2_TRY(w2_ptdavg(c, c->fld[W2_F_P], pn));
        tavg_pass1_kernel<<<g, 256, 0, c->stream>>>(c->nx, c->ny, c->pitch, us, vs, ts, pn, ptrs(q));
    } else {
        tavg_pass2_kernel<<<g, 256, 0, c->stream>>>(c->nx, c->ny, c->pitch, c->fld[W2_F_U], c->fld[W2_F_V], c->fld[W2_F_T], us,
                                                    vs, ts, c->met.djn, c->met.xen, c->met.yen, c->met.xzn, c->met.yzn, ptrs(q));
    }
    c->launches[3]++;
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}

// end of a pass (main.f:1239-1297): nts time steps were accumulated; dssrt = (...) * uref * dlref / re
extern "C" int wolfd2_b200_timeavg_finish(wolfd2_ctx *c, int32_t pass, int32_t nts, double uref, double dlref) {
    if (!c || !c->tavg || (pass != 1 && pass != 2) || nts < 1) { w2_set_error("timeavg_finish: bad arguments"); return W2_ERR_BAD_ARG; }
    W2_CUDA(cudaSetDevice(c->device));
    W2TimeAvg *q = (W2TimeAvg *)c->tavg;
    q->pass = 0;
    dim3 g((c->nx + 255) / 256, c->ny < 2048 ? c->ny : 2048);
    tavg_finish_kernel<<<g, 256, 0, c->stream>>>(c->nx, c->ny, c->pitch, pass, (double)nts, uref, dlref, c->par.re, ptrs(q));
    W2_CUDA(cudaGetLastError());
    W2_CUDA(cudaStreamSynchronize(c->stream));
    return W2_OK;
}

// which = 0..18: ubar vbar tbar pbar upb vpb tpb upupb vpvpb upvpb uptpb vptpb upxsb upysb vpxsb vpysb trbke dssrt dtdyb
extern "C" int wolfd2_b200_timeavg_get(wolfd2_ctx *c, int32_t which, double *host) {
    if (!c || !c->tavg || !host || which < 0 || which >= TA_COUNT) return W2_ERR_BAD_ARG;
    W2_CUDA(cudaSetDevice(c->device));
    W2_TRY(w2_download2d(c, host, ((W2TimeAvg *)c->tavg)->a[which]));
    W2_CUDA(cudaStreamSynchronize(c->stream));
    return W2_OK;
}
