// w2_traject.cu -- Lagrangian particle trajectories on the device and the node averages
// they read (N4): Traject, FwdEuler, HeunTrap, TrajFunc, TrajJac, iFindPos, BiLinInterp, Gauss
// (src/traject.f:154-719); VelAvg, PTDAvg (src/utility.f:512-647).  One GPU only.
//
// Particles are independent and the fields are frozen during Traject, so one thread carries one particle
// through all sub-steps.  iFindPos (:507-586) is an O(nx*ny) scan per particle per sub-step in the reference:
// "first (i,j), j outer / i inner, with x(i,j) > xp and y(i,j) > yp".  On a rectilinear grid (x depends on i only,
// y on j only, both increasing -- checked on the host at set-up) that is two binary searches with the same
// result; any other grid takes the literal scan (correct, slow).
//
// Two points are defined rather than reproduced: (a) Traject passes the INTEGER sub-step counter
// where HeunTrap expects the REAL step (traject.f:281 vs :336) -- h is the sub-step size here; (b) when the
// first iFindPos of a particle is non-zero the reference interpolates with out-of-range indices, integrates a
// local copy and throws it away at :296 (the second iFindPos sees the same, not yet updated position): the
// particle is only flagged here.
#include <stdlib.h>
#include <string.h>

#include "w2.cuh"

// ---------------------------------------------------------------------------------- node averages
// Regions are visited in the reference's order and later regions overwrite shared border nodes: each node
// takes the value of the LAST region whose loop covers it; nodes no loop covers are left alone.
__global__ void __launch_bounds__(256) velavg_kernel(const W2Regions *__restrict__ R, int nx, int ny, int pitch,
                                                     const double *__restrict__ u, const double *__restrict__ v,
                                                     double *__restrict__ util, double *__restrict__ vbar) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx + 1) return;
    for (int j = blockIdx.y; j <= ny + 1; j += gridDim.y) {
        int hit = 0;   // 1: average, 2: zero
        for (int q = 0; q < R->nreg; ++q)
            if (i >= R->iW[q] && i <= R->iE[q] && j >= R->jS[q] && j <= R->jN[q]) hit = R->type[q] == W2_RM_BLOCKG ? 2 : 1;
        if (hit == 1) {
            util[IDX(i, j)] = (u[IDX(i, j)] + u[IDX(i, j + 1)]) / 2.0;
            vbar[IDX(i, j)] = (v[IDX(i, j)] + v[IDX(i + 1, j)]) / 2.0;
        } else if (hit == 2) {
            util[IDX(i, j)] = 0.0;
            vbar[IDX(i, j)] = 0.0;
        }
    }
}
__global__ void __launch_bounds__(256) ptdavg_kernel(const W2Regions *__restrict__ R, int nx, int ny, int pitch,
                                                     const double *__restrict__ p, double *__restrict__ pav) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx + 1) return;
    for (int j = blockIdx.y; j <= ny + 1; j += gridDim.y) {
        int hit = 0;
        for (int q = 0; q < R->nreg; ++q) {
            if (R->type[q] == W2_RM_BLOCKG) {   // interior only (:548-552)
                if (i >= R->iW[q] + 1 && i <= R->iE[q] - 1 && j >= R->jS[q] + 1 && j <= R->jN[q] - 1) hit = 2;
            } else if (i >= R->iW[q] && i <= R->iE[q] && j >= R->jS[q] && j <= R->jN[q]) hit = 1;
        }
        if (hit == 1) {
            double sum = (p[IDX(i, j)] + p[IDX(i, j + 1)] + p[IDX(i + 1, j + 1)] + p[IDX(i + 1, j)]) / 4.00;
            if (fabs(sum) < 1.e-20) sum = 0.0;
            pav[IDX(i, j)] = sum;
        } else if (hit == 2) {
            pav[IDX(i, j)] = 0.0;
        }
    }
}
// TAveraged (src/utility.f:668-740): RT_TEMPER regions take dTRgVal (nScale 0) or zero (small scales)
__global__ void __launch_bounds__(256) taveraged_kernel(const W2Regions *__restrict__ R, const W2Thermal *__restrict__ H, int nscale,
                                                        int nx, int ny, int pitch, const double *__restrict__ t,
                                                        double *__restrict__ tav) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx + 1) return;
    for (int j = blockIdx.y; j <= ny + 1; j += gridDim.y) {
        int hit = -1;
        for (int q = 0; q < R->nreg; ++q)
            if (i >= R->iW[q] && i <= R->iE[q] && j >= R->jS[q] && j <= R->jN[q]) hit = q;
        if (hit < 0) continue;
        if (H->ttype[hit] == W2_RT_TEMPER) tav[IDX(i, j)] = nscale == 0 ? H->trgval[hit] : 0.0;
        else tav[IDX(i, j)] = (t[IDX(i, j)] + t[IDX(i, j + 1)] + t[IDX(i + 1, j + 1)] + t[IDX(i + 1, j)]) / 4.0;
    }
}
int w2_taveraged(wolfd2_ctx *c, int nscale, const double *t, double *tav) {
    dim3 g((c->nx + 2 + 255) / 256, c->ny + 2 < 2048 ? c->ny + 2 : 2048);
    taveraged_kernel<<<g, 256, 0, c->stream>>>(c->dreg, c->dth, nscale, c->nx, c->ny, c->pitch, t, tav);
    c->launches[3]++;
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}
int w2_velavg(wolfd2_ctx *c, const double *u, const double *v, double *util, double *vbar) {
    dim3 g((c->nx + 2 + 255) / 256, c->ny + 2 < 2048 ? c->ny + 2 : 2048);
    velavg_kernel<<<g, 256, 0, c->stream>>>(c->dreg, c->nx, c->ny, c->pitch, u, v, util, vbar);
    c->launches[3]++;
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}
int w2_ptdavg(wolfd2_ctx *c, const double *p, double *pav) {
    dim3 g((c->nx + 2 + 255) / 256, c->ny + 2 < 2048 ? c->ny + 2 : 2048);
    ptdavg_kernel<<<g, 256, 0, c->stream>>>(c->dreg, c->nx, c->ny, c->pitch, p, pav);
    c->launches[3]++;
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}

// ---------------------------------------------------------------------------------- particle kernel
struct TrFields {
    const double *x, *y, *u, *v, *un, *vn, *dens, *densn;   // node arrays, field layout
    const double *xs, *ys;                                  // rectilinear grids: x(i,1), y(1,j), 1-based
};

// iFindPos (:507-586): returns the status, ip/jp as the reference leaves them
__device__ __forceinline__ int tr_findpos(int nx, int ny, int pitch, int rectilinear, const TrFields &F, double xp, double yp,
                                          int &ip, int &jp) {
    ip = -1; jp = -1;
    double xrel = 0.0, yrel = 0.0;
    if (rectilinear) {
        // first index with xs > xp (xs increasing): lower bound by bisection; none -> nx+1
        int lo = 1, hi = nx + 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (F.xs[mid] > xp) hi = mid; else lo = mid + 1; }
        const int i = lo;
        lo = 1; hi = ny + 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (F.ys[mid] > yp) hi = mid; else lo = mid + 1; }
        const int j = lo;
        if (i <= nx && j <= ny) { ip = i; jp = j; xrel = F.xs[i] - xp; yrel = F.ys[j] - yp; }
    } else {
        for (int j = 1; j <= ny && ip < 0; ++j)
            for (int i = 1; i <= nx; ++i)
                if (F.x[IDX(i, j)] > xp && F.y[IDX(i, j)] > yp) {
                    ip = i; jp = j;
                    xrel = F.x[IDX(i, j)] - xp; yrel = F.y[IDX(i, j)] - yp;
                    break;
                }
    }
    int r = 0;
    if (ip <= 1 && xrel >= 0.0) r = 1;
    if (jp <= 1 && yrel >= 0.0) r = 3;
    if (ip < 0 || jp < 0) r = 2;
    return r;
}

// BiLinInterp (:592-660); the geometry terms are shared by the six fields of one particle
struct TrGeom { double x1, x2, x3, x4, y1, y2, y3, y4, xs, ys, ya, xb, yc, xd; };
__device__ __forceinline__ void tr_geom(int pitch, const TrFields &F, int i, int j, double xs, double ys, TrGeom &G) {
    G.x1 = F.x[IDX(i - 1, j - 1)]; G.x2 = F.x[IDX(i, j - 1)]; G.x3 = F.x[IDX(i, j)]; G.x4 = F.x[IDX(i - 1, j)];
    G.y1 = F.y[IDX(i - 1, j - 1)]; G.y2 = F.y[IDX(i, j - 1)]; G.y3 = F.y[IDX(i, j)]; G.y4 = F.y[IDX(i - 1, j)];
    G.xs = xs; G.ys = ys;
    G.ya = G.y1 + (xs - G.x1) * (G.y2 - G.y1) / (G.x2 - G.x1);
    G.xb = G.x2 + (ys - G.y2) * (G.x3 - G.x2) / (G.y3 - G.y2);
    G.yc = G.y4 + (xs - G.x4) * (G.y3 - G.y4) / (G.x3 - G.x4);
    G.xd = G.x1 + (ys - G.y1) * (G.x4 - G.x1) / (G.y4 - G.y1);
}
__device__ __forceinline__ double tr_interp(int pitch, const TrGeom &G, int i, int j, const double *__restrict__ f) {
    const double f1 = f[IDX(i - 1, j - 1)], f2 = f[IDX(i, j - 1)], f3 = f[IDX(i, j)], f4 = f[IDX(i - 1, j)];
    const double fa = f1 + (G.xs - G.x1) * (f2 - f1) / (G.x2 - G.x1);
    const double fb = f2 + (G.ys - G.y2) * (f3 - f2) / (G.y3 - G.y2);
    const double fc = f4 + (G.xs - G.x4) * (f3 - f4) / (G.x3 - G.x4);
    const double fd = f1 + (G.ys - G.y1) * (f4 - f1) / (G.y4 - G.y1);
    const double fsx = fd + (G.xs - G.xd) * (fb - fd) / (G.xb - G.xd);
    const double fsy = fa + (G.ys - G.ya) * (fc - fa) / (G.yc - G.ya);
    return (fsx + fsy) / 2.00;
}

__device__ __forceinline__ double tr_func(int i, double fr, double uf, double vf, double cpx, double cpy, const double *w) {
    switch (i) {   // TrajFunc (:439-469)
    case 1: return w[1];
    case 2: return cpx * fabs(uf - w[1]) * (uf - w[1]);
    case 3: return w[3];
    default: return cpy * fabs(vf - w[3]) * (vf - w[3]) - 1.00 / fr;
    }
}
#define GA(i, j) a[((i) - 1) + 4 * ((j) - 1)]
// Gauss (:666-719), n = 4
__device__ void tr_gauss(double *a, double *b, double *x) {
    const int n = 4;
    for (int k = 1; k <= n - 1; ++k) {
        double amax = fabs(GA(k, k));
        int imax = k;
        for (int i = k + 1; i <= n; ++i) {
            const double da = fabs(GA(i, k));
            if (da > amax) { amax = da; imax = i; }
        }
        if (imax != k) {
            for (int j = k; j <= n; ++j) { const double t = GA(k, j); GA(k, j) = GA(imax, j); GA(imax, j) = t; }
            const double t = b[k - 1]; b[k - 1] = b[imax - 1]; b[imax - 1] = t;
        }
        for (int i = k + 1; i <= n; ++i) {
            const double dm = GA(i, k) / GA(k, k);
            b[i - 1] = b[i - 1] - dm * b[k - 1];
            for (int j = k + 1; j <= n; ++j) GA(i, j) = GA(i, j) - dm * GA(k, j);
        }
    }
    x[n - 1] = b[n - 1] / GA(n, n);
    for (int i = n - 1; i >= 1; --i) {
        x[i - 1] = 0.00;
        for (int j = i + 1; j <= n; ++j) x[i - 1] = x[i - 1] + GA(i, j) * x[j - 1];
        x[i - 1] = (b[i - 1] - x[i - 1]) / GA(i, i);
    }
}
// HeunTrap (:336-433) with h = the sub-step size (see header)
__device__ void tr_heuntrap(int maxit, double h, double toler, double delta, double fr, double uf1, double vf1, double ufn,
                            double vfn, double cpx1, double cpxn, double cpy1, double cpyn, double *u) {
    double g[4], udel[4], f[4], us[4], a[16];
    const double h2 = h / 2.0;
    for (int m = 1; m <= maxit; ++m) {
        if (m == 1) {
            for (int i = 1; i <= 4; ++i) {
                g[i - 1] = tr_func(i, fr, ufn, vfn, cpxn, cpyn, u);
                us[i - 1] = u[i - 1] + h * g[i - 1];
            }
            if (maxit == 1) { for (int i = 0; i < 4; ++i) u[i] = us[i]; return; }
            for (int i = 1; i <= 4; ++i) us[i - 1] = u[i - 1] + h2 * (g[i - 1] + tr_func(i, fr, uf1, vf1, cpx1, cpy1, us));
        }
        for (int j = 0; j < 16; ++j) a[j] = 0.00;   // TrajJac (:475-501)
        GA(1, 2) = 1.00;
        GA(2, 2) = cpx1 * ((uf1 - us[1]) - fabs(uf1 - us[1]));
        GA(3, 4) = 1.00;
        GA(4, 4) = cpy1 * ((vf1 - us[3]) - fabs(vf1 - us[3]));
        for (int i = 1; i <= 4; ++i) {
            f[i - 1] = us[i - 1] - h2 * tr_func(i, fr, uf1, vf1, cpx1, cpy1, us) - (u[i - 1] + h2 * g[i - 1]);
            for (int j = 1; j <= 4; ++j) GA(i, j) = (i == j ? 1.0 : 0.0) - h2 * GA(i, j);
        }
        for (int i = 0; i < 4; ++i) f[i] = -f[i];
        tr_gauss(a, f, udel);
        double dumax = 0.0;
        for (int i = 0; i < 4; ++i) {
            const double du = fabs(udel[i]);
            if (du > dumax) dumax = du;
            us[i] = us[i] + delta * udel[i];
        }
        if (dumax < toler) { for (int i = 0; i < 4; ++i) u[i] = us[i]; return; }
    }
}
#undef GA

struct TrPar { int ntr, ntsubstp, method, cdeq, maxit; double dk, densref, fr, toler, delta; };

__global__ void __launch_bounds__(128) traject_kernel(int nx, int ny, int pitch, int rectilinear, TrPar P, TrFields F,
                                                      const double *__restrict__ cpartx, const double *__restrict__ cparty,
                                                      const double *__restrict__ repc, int *__restrict__ nTOutBnd,
                                                      double *__restrict__ xp, double *__restrict__ yp, double *__restrict__ up,
                                                      double *__restrict__ vp) {
    const double dOne = 1.0, dTwo = 2.0, dThree = 3.0, dSix = 6.0;
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < P.ntr; l += gridDim.x * blockDim.x) {
        int out = nTOutBnd[l];
        double w[4] = {xp[l], up[l], yp[l], vp[l]};
        const double cx = cpartx[l], cy = cparty[l], rc = repc[l];
        for (int k = 1; k <= P.ntsubstp && out <= 0; ++k) {
            int ip, jp;
            out = tr_findpos(nx, ny, pitch, rectilinear, F, w[0], w[2], ip, jp);   // :226
            if (out > 0) break;
            TrGeom G;
            tr_geom(pitch, F, ip, jp, w[0], w[2], G);
            const double uf1 = tr_interp(pitch, G, ip, jp, F.u), vf1 = tr_interp(pitch, G, ip, jp, F.v);   // :229-234
            const double ufn = tr_interp(pitch, G, ip, jp, F.un), vfn = tr_interp(pitch, G, ip, jp, F.vn);
            double df1 = tr_interp(pitch, G, ip, jp, F.dens), dfn = tr_interp(pitch, G, ip, jp, F.densn);
            df1 = P.densref * (df1 + dOne);
            dfn = P.densref * (dfn + dOne);
            const double du = uf1 - w[1], dv = vf1 - w[3];
            const double rep = rc * sqrt(du * du + dv * dv);       // :241
            const double dstokes = 24.00 / rep;
            double cd;
            if (P.cdeq == 1) cd = dstokes;                         // :247-264
            else if (P.cdeq == 2) cd = dstokes * (dOne + pow(rep, dTwo / dThree) / dSix);
            else if (P.cdeq == 3) cd = 0.40 + dstokes + dSix / (dOne + sqrt(rep));
            else cd = dstokes * (dOne + 0.1970 * pow(rep, 0.63) + (0.26e-3) * pow(rep, 1.38));
            const double cpx1 = cd * cx * df1, cpxn = cd * cx * dfn;
            const double cpy1 = cd * cy * df1, cpyn = cd * cy * dfn;
            if (P.method == 1) {
                tr_heuntrap(P.maxit, P.dk, P.toler, P.delta, P.fr, uf1, vf1, ufn, vfn, cpx1, cpxn, cpy1, cpyn, w);
            } else {                                               // FwdEuler (:315-330)
                w[1] = w[1] + P.dk * (cpxn * fabs(ufn - w[1]) * (ufn - w[1]));
                w[0] = w[0] + P.dk * w[1];
                w[3] = w[3] + P.dk * (cpyn * fabs(vfn - w[3]) * (vfn - w[3]) - (1.0 / P.fr));
                w[2] = w[2] + P.dk * w[3];
            }
        }
        nTOutBnd[l] = out;
        xp[l] = w[0]; up[l] = w[1]; yp[l] = w[2]; vp[l] = w[3];
    }
}

// ---------------------------------------------------------------------------------- host side
static int tr_alloc_n(void **p, size_t bytes) {
    W2_CUDA(cudaMalloc(p, bytes));
    W2_CUDA(cudaMemset(*p, 0, bytes));
    W2_CUDA(cudaDeviceSynchronize());   // the fill is on the default stream: see dalloc (w2_context.cu)
    return W2_OK;
}
void w2_traj_release(wolfd2_ctx *c) {
    W2Traj *t = c->traj;
    if (!t) return;
    double *f[] = {t->gx, t->gy, t->un_av, t->vn_av, t->dn_av};
    for (int k = 0; k < 5; ++k) if (f[k]) cudaFree(f[k] + c->row_off);
    void *v[] = {t->xs, t->ys, t->cpartx, t->cparty, t->repc, t->xp, t->yp, t->up, t->vp, t->out};
    for (int k = 0; k < 10; ++k) cudaFree(v[k]);
    free(t);
    c->traj = nullptr;
}

// x, y: the grid nodes as main.f holds them, REAL*8 (0:mnx,0:mny)
int w2_traj_set_grid(wolfd2_ctx *c, const double *x, const double *y) {
    if (!c->traj) {
        c->traj = (W2Traj *)calloc(1, sizeof(W2Traj));
        if (!c->traj) return W2_ERR_BAD_ARG;
    }
    W2Traj *t = c->traj;
    double **f[] = {&t->gx, &t->gy, &t->un_av, &t->vn_av, &t->dn_av};
    for (int k = 0; k < 5; ++k) if (!*f[k]) W2_TRY(w2_alloc_field(c, f[k]));
    W2_TRY(w2_upload2d(c, t->gx, x));
    W2_TRY(w2_upload2d(c, t->gy, y));
    // rectilinear?  x(i,j) == x(i,1), y(i,j) == y(1,j), both strictly increasing (host check, set-up only)
    const int nx = c->nx, ny = c->ny;
    const size_t ld = (size_t)c->mnx + 1;
    int rect = 1;
    for (int j = 1; j <= ny && rect; ++j)
        for (int i = 1; i <= nx; ++i)
            if (x[i + ld * j] != x[i + ld * 1] || y[i + ld * j] != y[1 + ld * j]) { rect = 0; break; }
    for (int i = 2; i <= nx && rect; ++i) if (!(x[i + ld] > x[i - 1 + ld])) rect = 0;
    for (int j = 2; j <= ny && rect; ++j) if (!(y[1 + ld * j] > y[1 + ld * (j - 1)])) rect = 0;
    t->rectilinear = rect;
    if (!t->xs) W2_TRY(tr_alloc_n((void **)&t->xs, (size_t)(nx + 2) * 8));
    if (!t->ys) W2_TRY(tr_alloc_n((void **)&t->ys, (size_t)(ny + 2) * 8));
    double *hx = (double *)calloc(nx + 2, 8), *hy = (double *)calloc(ny + 2, 8);
    if (!hx || !hy) { free(hx); free(hy); return W2_ERR_BAD_ARG; }
    for (int i = 1; i <= nx; ++i) hx[i] = x[i + ld];
    for (int j = 1; j <= ny; ++j) hy[j] = y[1 + ld * j];
    cudaMemcpyAsync(t->xs, hx, (size_t)(nx + 2) * 8, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(t->ys, hy, (size_t)(ny + 2) * 8, cudaMemcpyHostToDevice, c->stream);
    W2_CUDA(cudaStreamSynchronize(c->stream));
    free(hx); free(hy);
    return W2_OK;
}

int w2_traj_set_particles(wolfd2_ctx *c, const wolfd2_traject *tr, const double *cpartx, const double *cparty, const double *repc,
                          const double *xp, const double *yp, const double *up, const double *vp, const int32_t *nTOutBnd) {
    W2Traj *t = c->traj;
    if (!t) { w2_set_error("trajectories: the grid nodes were not given"); return W2_ERR_BAD_ARG; }
    if (tr->ntr < 0 || tr->ntsubstp < 1) { w2_set_error("trajectories: ntr >= 0 and ntsubstp >= 1 required"); return W2_ERR_BAD_ARG; }
    if (tr->nTrMethod != 1 && tr->nTrMethod != 2) { w2_set_error("Error: Wrong nTrMethod flag passed to Traject"); return W2_ERR_BAD_ARG; }   // :287-289
    if (tr->nTrCdEq < 1 || tr->nTrCdEq > 4) { w2_set_error("Error: Wrong nTrCdEq flag passed to Traject"); return W2_ERR_BAD_ARG; }         // :261-263
    if (tr->ntr > t->cap) {
        void **v[] = {(void **)&t->cpartx, (void **)&t->cparty, (void **)&t->repc, (void **)&t->xp, (void **)&t->yp,
                      (void **)&t->up, (void **)&t->vp};
        for (int k = 0; k < 7; ++k) { cudaFree(*v[k]); *v[k] = nullptr; W2_TRY(tr_alloc_n(v[k], (size_t)tr->ntr * 8)); }
        cudaFree(t->out); t->out = nullptr;
        W2_TRY(tr_alloc_n((void **)&t->out, (size_t)tr->ntr * 4));
        t->cap = tr->ntr;
    }
    t->tr = *tr;
    const size_t nb = (size_t)tr->ntr * 8;
    const double *src[] = {cpartx, cparty, repc, xp, yp, up, vp};
    double *dst[] = {t->cpartx, t->cparty, t->repc, t->xp, t->yp, t->up, t->vp};
    for (int k = 0; k < 7; ++k) if (src[k] && nb) W2_CUDA(cudaMemcpyAsync(dst[k], src[k], nb, cudaMemcpyHostToDevice, c->stream));
    if (nTOutBnd && nb) W2_CUDA(cudaMemcpyAsync(t->out, nTOutBnd, (size_t)tr->ntr * 4, cudaMemcpyHostToDevice, c->stream));
    W2_CUDA(cudaStreamSynchronize(c->stream));
    t->active = 1;
    return W2_OK;
}

// Traject on node-averaged fields that are already on the device
int w2_traject(wolfd2_ctx *c, double dkflow, double fr, const double *u, const double *v, const double *un, const double *vn,
               const double *dens, const double *densn) {
    W2Traj *t = c->traj;
    if (!t || !t->active) { w2_set_error("Traject: no particles"); return W2_ERR_BAD_ARG; }
    if (t->tr.ntr == 0) return W2_OK;
    TrPar P;
    P.ntr = t->tr.ntr; P.ntsubstp = t->tr.ntsubstp; P.method = t->tr.nTrMethod; P.cdeq = t->tr.nTrCdEq; P.maxit = t->tr.mTrHTmit;
    P.dk = t->tr.ntsubstp != 1 ? dkflow / (double)t->tr.ntsubstp : dkflow;   // :210-214
    P.densref = t->tr.densref; P.fr = fr; P.toler = t->tr.dTrHTtol; P.delta = t->tr.dTrHTdel;
    TrFields F;
    F.x = t->gx; F.y = t->gy; F.u = u; F.v = v; F.un = un; F.vn = vn; F.dens = dens; F.densn = densn; F.xs = t->xs; F.ys = t->ys;
    int blocks = (P.ntr + 127) / 128;
    const int cap = c->num_sms * 16;
    if (blocks > cap) blocks = cap;
    traject_kernel<<<blocks, 128, 0, c->stream>>>(c->nx, c->ny, c->pitch, t->rectilinear, P, F, t->cpartx, t->cparty, t->repc, t->out,
                                                  t->xp, t->yp, t->up, t->vp);
    c->launches[3]++;
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}

// src/main.f:1000-1024: node averages of (u,v), (un,vn), d, dn, then Traject.  As in the reference the averages
// of the new time level go to the starred arrays us, vs, ts.
int w2_traject_step(wolfd2_ctx *c) {
    W2Traj *t = c->traj;
    double *us = c->fld[W2_F_US], *vs = c->fld[W2_F_VS], *ts = c->fld[W2_F_TS];
    W2_TRY(w2_velavg(c, c->fld[W2_F_U], c->fld[W2_F_V], us, vs));
    W2_TRY(w2_velavg(c, c->fld[W2_F_UN], c->fld[W2_F_VN], t->un_av, t->vn_av));
    W2_TRY(w2_ptdavg(c, c->fld[W2_F_D], ts));
    W2_TRY(w2_ptdavg(c, c->fld[W2_F_DN], t->dn_av));
    return w2_traject(c, c->par.dk, c->par.fr, us, vs, t->un_av, t->vn_av, ts, t->dn_av);
}

extern "C" int wolfd2_b200_set_trajectories(wolfd2_ctx *c, const wolfd2_traject *tr, const double *x, const double *y,
                                            const double *cpartx, const double *cparty, const double *repc, const double *xp,
                                            const double *yp, const double *up, const double *vp, const int32_t *nTOutBnd) {
    if (!c || !tr) return W2_ERR_BAD_ARG;
    W2_CUDA(cudaSetDevice(c->device));
    if (tr->ntr <= 0) { if (c->traj) c->traj->active = 0; return W2_OK; }
    if (c->world > 1) { w2_set_error("trajectories are not supported on several GPUs"); return W2_ERR_UNSUPPORTED; }
    if (!x || !y || !cpartx || !cparty || !repc || !xp || !yp || !up || !vp) { w2_set_error("set_trajectories: missing array"); return W2_ERR_BAD_ARG; }
    W2_TRY(w2_traj_set_grid(c, x, y));
    return w2_traj_set_particles(c, tr, cpartx, cparty, repc, xp, yp, up, vp, nTOutBnd);
}
extern "C" int wolfd2_b200_get_particles(wolfd2_ctx *c, double *xp, double *yp, double *up, double *vp, int32_t *nTOutBnd) {
    if (!c || !c->traj || !c->traj->active) { w2_set_error("get_particles: no trajectories in this context"); return W2_ERR_BAD_ARG; }
    W2_CUDA(cudaSetDevice(c->device));
    W2Traj *t = c->traj;
    const size_t nb = (size_t)t->tr.ntr * 8;
    double *dst[] = {xp, yp, up, vp};
    const double *src[] = {t->xp, t->yp, t->up, t->vp};
    for (int k = 0; k < 4; ++k) if (dst[k] && nb) W2_CUDA(cudaMemcpyAsync(dst[k], src[k], nb, cudaMemcpyDeviceToHost, c->stream));
    if (nTOutBnd && nb) W2_CUDA(cudaMemcpyAsync(nTOutBnd, t->out, (size_t)t->tr.ntr * 4, cudaMemcpyDeviceToHost, c->stream));
    W2_CUDA(cudaStreamSynchronize(c->stream));
    return W2_OK;
}
