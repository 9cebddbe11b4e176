// w2_thermal.cu -- thermal energy equation: TempBoundCond (src/bound_cond.f:1030-1204),
// ThermEnergy (src/thermal.f:24-272; its two split steps are COMP 2 of w2_momentum.cu), EqState
// (src/thermal.f:283-327), and the region maps they need.  One GPU only.
#include <string.h>

#include "w2.cuh"

#define BC_THREADS 1024
#define T(i, j) t[IDX(i, j)]
#define PFOR(var, lo, hi) for (int var = (lo) + (int)threadIdx.x; var <= (hi); var += BC_THREADS)

// Same walking order as the reference (regions jreg outer / ireg inner, faces W,E,S,N), one CTA, a barrier
// after every loop: later loops read what earlier ones wrote at region corners (see w2_bc.cu).
__global__ void __launch_bounds__(BC_THREADS) temp_bc_kernel(const W2Regions *__restrict__ R, const W2Thermal *__restrict__ H,
                                                            int pitch, double *t) {
    const double dTwo = 2.0;
    const int nreg = R->nreg;
    for (int q = 0; q < nreg; ++q) {
        const int iW = R->iW[q], iE = R->iE[q], jS = R->jS[q], jN = R->jN[q];
        const double vW = R->val[q][W2_WEST - 1][W2_T - 1], vE = R->val[q][W2_EAST - 1][W2_T - 1];
        const double vS = R->val[q][W2_SOUTH - 1][W2_T - 1], vN = R->val[q][W2_NORTH - 1][W2_T - 1];
        if (H->ttype[q] == W2_RT_TEMPER) {   // :1083-1113
            const int w = iE - iW, h = jN - jS;
            const double val = H->trgval[q];
            for (int k = threadIdx.x; k < w * h; k += BC_THREADS) T(iW + 1 + k % w, jS + 1 + k / w) = val;
            __syncthreads();
            PFOR(j, jS + 1, jN) T(iW + 1, j) = dTwo * vW - T(iW, j);
            __syncthreads();
            PFOR(j, jS + 1, jN) T(iE, j) = dTwo * vE - T(iE + 1, j);
            __syncthreads();
            PFOR(i, iW + 1, iE) T(i, jS + 1) = dTwo * vS - T(i, jS);
            __syncthreads();
            PFOR(i, iW + 1, iE) T(i, jN) = dTwo * vN - T(i, jN + 1);
            __syncthreads();
            continue;
        }
        int bt = H->tbd[q][W2_WEST - 1];
        if (bt == W2_BT_TEMPER) { PFOR(j, jS + 1, jN) T(iW, j) = dTwo * vW - T(iW + 1, j); }
        else if (bt == W2_BT_HTFLUX) { PFOR(j, jS + 1, jN) T(iW, j) = vW + T(iW + 1, j); }
        __syncthreads();
        bt = H->tbd[q][W2_EAST - 1];
        if (bt == W2_BT_TEMPER) { PFOR(j, jS + 1, jN) T(iE + 1, j) = dTwo * vE - T(iE, j); }
        else if (bt == W2_BT_HTFLUX) { PFOR(j, jS + 1, jN) T(iE + 1, j) = vE + T(iE, j); }
        __syncthreads();
        bt = H->tbd[q][W2_SOUTH - 1];
        if (bt == W2_BT_TEMPER) { PFOR(i, iW + 1, iE) T(i, jS) = dTwo * vS - T(i, jS + 1); }
        else if (bt == W2_BT_HTFLUX) { PFOR(i, iW + 1, iE) T(i, jS) = vS + T(i, jS + 1); }
        __syncthreads();
        bt = H->tbd[q][W2_NORTH - 1];
        if (bt == W2_BT_TEMPER) { PFOR(i, iW + 1, iE) T(i, jN + 1) = dTwo * vN - T(i, jN); }
        else if (bt == W2_BT_HTFLUX) { PFOR(i, iW + 1, iE) T(i, jN + 1) = vN + T(i, jN); }
        __syncthreads();
    }
}

// s(i,j) (thermal.f:123-147) and the fixed-temperature mask (:242-266): the regions tile (iW+1..iE, jS+1..jN)
__global__ void thermal_map_kernel(const W2Regions *__restrict__ R, const W2Thermal *__restrict__ H, int nx, int ny, int pitch,
                                   double *__restrict__ heat, unsigned char *__restrict__ tmask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i > nx + 1 || j > ny + 1) return;
    double s = 0.0;
    unsigned char m = 0;
    for (int q = 0; q < R->nreg; ++q) {
        if (i >= R->iW[q] + 1 && i <= R->iE[q] && j >= R->jS[q] + 1 && j <= R->jN[q]) {
            s = H->ttype[q] == W2_RT_HEATGN ? H->hgst[q] : 0.0;
            if (H->ttype[q] == W2_RT_TEMPER) m = 1;
        }
    }
    heat[IDX(i, j)] = s;
    tmask[IDX(i, j)] = m;
}

int w2_set_thermal_tables(wolfd2_ctx *c, const int32_t *nTRgType, const int32_t *nTemBdTp, const double *dTRgVal,
                          const double *dHGSTval) {
    W2Thermal &H = c->hth;
    memset(&H, 0, sizeof(H));
    const W2Regions &R = c->hreg;
    const int plane = g_mgri * g_mgrj;
    for (int jr = 1; jr <= R.nregJ; ++jr)
        for (int ir = 1; ir <= R.nregI; ++ir) {
            const int q = (ir - 1) + R.nregI * (jr - 1);
            const int base = (ir - 1) + g_mgri * (jr - 1);
            H.ttype[q] = nTRgType ? nTRgType[base] : W2_RT_NOSRCE;
            for (int k = 0; k < 4; ++k) {
                H.tbd[q][k] = nTemBdTp ? nTemBdTp[base + plane * k] : W2_BT_INTERN;
                if (H.tbd[q][k] < W2_BT_INTERN || H.tbd[q][k] > W2_BT_HTFLUX) {   // bound_cond.f:1133-1135: message + stop
                    w2_set_error("Wrong nTemBdTp flag %d in region %d,%d face %d", H.tbd[q][k], ir, jr, k + 1);
                    return W2_ERR_BAD_ARG;
                }
            }
            H.trgval[q] = dTRgVal ? dTRgVal[base] : 0.0;
            H.hgst[q] = dHGSTval ? dHGSTval[base] : 0.0;
        }
    c->th_tables = nTRgType && nTemBdTp;
    W2_CUDA(cudaMemcpyAsync(c->dth, &c->hth, sizeof(W2Thermal), cudaMemcpyHostToDevice, c->stream));
    W2_CUDA(cudaStreamSynchronize(c->stream));
    dim3 grid((c->nx + 2 + 255) / 256, c->ny + 2);
    thermal_map_kernel<<<grid, 256, 0, c->stream>>>(c->dreg, c->dth, c->nx, c->ny, c->pitch, c->heat_s, c->tmask);
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}

extern "C" int wolfd2_b200_set_thermal(wolfd2_ctx *c, const wolfd2_thermal *th) {
    if (!c || !th) return W2_ERR_BAD_ARG;
    W2_CUDA(cudaSetDevice(c->device));
    if (th->nthermen == 1) {
        if (c->world > 1) { w2_set_error("the thermal energy equation is not supported on several GPUs"); return W2_ERR_UNSUPPORTED; }
        if (!(th->pe > 0.0)) { w2_set_error("set_thermal: Peclet number must be positive"); return W2_ERR_BAD_ARG; }
        if (!th->nTRgType || !th->nTemBdTp || !th->dTRgVal || !th->dHGSTval) {
            w2_set_error("set_thermal: region tables missing");
            return W2_ERR_BAD_ARG;
        }
        W2_TRY(w2_set_thermal_tables(c, th->nTRgType, th->nTemBdTp, th->dTRgVal, th->dHGSTval));
    } else if (th->nTRgType && th->nTemBdTp && th->dTRgVal) {
        // cold flow, tables given all the same: the ATD model calls TempBoundCond and Filter(_T_) in any case
        W2_TRY(w2_set_thermal_tables(c, th->nTRgType, th->nTemBdTp, th->dTRgVal, th->dHGSTval));
    }
    c->th = *th;
    c->th.nTRgType = nullptr; c->th.nTemBdTp = nullptr; c->th.dTRgVal = nullptr; c->th.dHGSTval = nullptr;
    return W2_OK;
}

int w2_temp_bc(wolfd2_ctx *c, double *t) {
    temp_bc_kernel<<<1, BC_THREADS, 0, c->stream>>>(c->dreg, c->dth, c->pitch, t);
    c->launches[3]++;
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}

__global__ void __launch_bounds__(256) t_update_kernel(int nx, int ny, int pitch, const double *__restrict__ dts,
                                                       double *__restrict__ t) {
    const int i = 2 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx) return;
    for (int j = 2 + blockIdx.y; j <= ny; j += gridDim.y) t[IDX(i, j)] = t[IDX(i, j)] + dts[IDX(i, j)];   // :274-279
}

// ThermEnergy (thermal.f:24-272) on t; un, vn, us, vs, tn are the context's fields (main.f:840-850)
int w2_thermenergy(wolfd2_ctx *c, double *t) {
    W2_TRY(w2_temp_bc(c, t));                       // :104
    W2_TRY(w2_thermal_solve(c, c->dus));            // :153-270; dus is free outside nAuxMomentum
    dim3 g((c->nx - 1 + 255) / 256, (c->ny - 1) < 2048 ? (c->ny - 1) : 2048);
    t_update_kernel<<<g, 256, 0, c->stream>>>(c->nx, c->ny, c->pitch, c->dus, t);
    c->launches[3]++;
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}

__global__ void __launch_bounds__(256) eqstate_kernel(int nx, int ny, int pitch, double uref, double densref, double tmax,
                                                      double tref, double rconst, const double *__restrict__ p,
                                                      const double *__restrict__ t, double *__restrict__ den) {
    const double pref = densref * rconst * tref;
    const int i = 2 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx) return;
    for (int j = 2 + blockIdx.y; j <= ny; j += gridDim.y) {
        const double c1 = p[IDX(i, j)] * densref * (uref * uref) + pref;
        const double c2 = densref * rconst * (t[IDX(i, j)] * (tmax - tref) + tref);
        double d = c1 / c2 - 1.0;
        if (fabs(d) < 1.e-10) d = 0.0;
        den[IDX(i, j)] = d;
    }
}

int w2_eqstate(wolfd2_ctx *c, const double *p, const double *t, double *den) {
    c->d_nonzero = 1;
    dim3 g((c->nx - 1 + 255) / 256, (c->ny - 1) < 2048 ? (c->ny - 1) : 2048);
    eqstate_kernel<<<g, 256, 0, c->stream>>>(c->nx, c->ny, c->pitch, c->th.uref, c->th.densref, c->th.tmax, c->th.tref,
                                             c->th.rconst, p, t, den);
    c->launches[3]++;
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}
