// w2_momentum.cu -- auxiliary-velocity solve: nAuxMomentum (src/momentum.f:33-193), XMomentum
// (:199-514), YMomentum (:520-838) with ConvCoef (:864-981), DConvU/DDiffU/DConvV/DDiffV
// (:987-1109) inlined.
//
// The reference materialises 13 full-size work arrays per component (cj1..difn) and an AoS
// matrix a(3,mn).  Here every coefficient is re-formed per unknown from the primary fields and
// metrics inside the assembly kernel, with the reference's operation order, so the assembled
// rows are bit-identical to a non-FMA CPU build; only the tridiagonal solve (w2_tridiag.cu) uses
// a different (parallel) elimination order.
//
// Cells the reference never writes are zeros there (static storage): cj1(nx+1,j),
// cj2(i,ny+1) below are the load-bearing cases and are written out as literal 0.0.
#include "w2.cuh"

#define F(a, i, j) a[IDX(i, j)]

struct MomArgs {
    int nx, ny, pitch;
    double dk, re, fr;
    const double *us, *vs, *un, *vn, *d, *dn;
    // x-momentum metrics
    const double *rbn, *rgn, *rac, *rbc, *dju, *xec, *yec, *xzn, *yzn, *xeu, *yeu, *xzu, *yzu;
    // y-momentum metrics
    const double *ran, *rgc, *djv, *xen, *yen, *xzc, *yzc, *xev, *yev, *xzv, *yzv;
    const unsigned char *xmask, *ymask;
    double *ta, *td, *tc, *tb;
    const double *x1;  // step-1 solution (chain order)
};

// ---- ConvCoef pieces ------------------------------------------------------------------------
// case 1 (x-momentum rhs, djac = 1): cc1 on 1..nx+1,1..ny+1 ; cc2 on 1..nx,1..ny   (:901-913)
__device__ __forceinline__ double x_c1(const MomArgs &m, const double *u, const double *v, int i, int j) {
    const int pitch = m.pitch;
    return (F(m.yec, i, j) * (F(u, i, j) + F(u, i - 1, j)) - F(m.xec, i, j) * (F(v, i, j) + F(v, i, j - 1))) * 0.5;
}
__device__ __forceinline__ double x_c2(const MomArgs &m, const double *u, const double *v, int i, int j) {
    const int pitch = m.pitch;
    return (F(m.xzn, i, j) * (F(v, i + 1, j) + F(v, i, j)) - F(m.yzn, i, j) * (F(u, i, j + 1) + F(u, i, j))) * 0.5;
}
// case 4 (x-momentum upwind Jacobian, djac = 2): written on i=0..nx, j=1..ny   (:942-950)
__device__ __forceinline__ double x_cj1(const MomArgs &m, int i, int j) {
    const int pitch = m.pitch;
    if (i > m.nx) return 0.0;  // never written by the reference -> static zero
    const double *u = m.us, *v = m.vs;
    return 2.0 * F(m.yeu, i, j) * F(u, i, j)
           - F(m.xeu, i, j) * (F(v, i + 1, j) + F(v, i, j) + F(v, i + 1, j - 1) + F(v, i, j - 1)) / 4.0;
}
__device__ __forceinline__ double x_cj2(const MomArgs &m, int i, int j) {
    const int pitch = m.pitch;
    if (j > m.ny) return 0.0;  // cj2(i,ny+1): never written
    const double *u = m.us, *v = m.vs;
    return F(m.xzu, i, j) * (F(v, i + 1, j) + F(v, i, j) + F(v, i + 1, j - 1) + F(v, i, j - 1)) / 4.0
           - 2.0 * F(m.yzu, i, j) * F(u, i, j);
}
// case 2 (y-momentum rhs): cc1 on 1..nx,1..ny ; cc2 on 1..nx+1,1..ny+1   (:916-928)
__device__ __forceinline__ double y_c1(const MomArgs &m, const double *u, const double *v, int i, int j) {
    const int pitch = m.pitch;
    return (F(m.yen, i, j) * (F(u, i, j + 1) + F(u, i, j)) - F(m.xen, i, j) * (F(v, i + 1, j) + F(v, i, j))) * 0.5;
}
__device__ __forceinline__ double y_c2(const MomArgs &m, const double *u, const double *v, int i, int j) {
    const int pitch = m.pitch;
    return (F(m.xzc, i, j) * (F(v, i, j) + F(v, i, j - 1)) - F(m.yzc, i, j) * (F(u, i, j) + F(u, i - 1, j))) * 0.5;
}
// case 5 (y-momentum upwind Jacobian, djac = 2): written on i=1..nx, j=0..ny   (:953-961)
__device__ __forceinline__ double y_cj1(const MomArgs &m, int i, int j) {
    const int pitch = m.pitch;
    if (i > m.nx) return 0.0;  // cj1(nx+1,j): never written
    const double *u = m.us, *v = m.vs;
    return F(m.yev, i, j) * (F(u, i, j + 1) + F(u, i - 1, j + 1) + F(u, i, j) + F(u, i - 1, j)) / 4.0
           - 2.0 * F(m.xev, i, j) * F(v, i, j);
}
__device__ __forceinline__ double y_cj2(const MomArgs &m, int i, int j) {
    const int pitch = m.pitch;
    if (j > m.ny) return 0.0;  // cj2(i,ny+1): never written
    const double *u = m.us, *v = m.vs;
    return 2.0 * F(m.xzv, i, j) * F(v, i, j)
           - F(m.yzv, i, j) * (F(u, i, j + 1) + F(u, i - 1, j + 1) + F(u, i, j) + F(u, i - 1, j)) / 4.0;
}

// DConvU (:1000-1006) and DDiffU (:1030-1042) at one point
__device__ __forceinline__ double x_conv(const MomArgs &m, const double *u, const double *v, int i, int j) {
    const int pitch = m.pitch;
    const double c1 = x_c1(m, u, v, i, j), c1e = x_c1(m, u, v, i + 1, j);
    const double c2 = x_c2(m, u, v, i, j), c2s = x_c2(m, u, v, i, j - 1);
    return -c2s * F(u, i, j - 1) - c1 * F(u, i - 1, j) + (c1e - c1 + c2 - c2s) * F(u, i, j)
           + c1e * F(u, i + 1, j) + c2 * F(u, i, j + 1);
}
__device__ __forceinline__ double x_diff(const MomArgs &m, const double *u, int i, int j) {
    const int pitch = m.pitch;
    const double *ac = m.rac, *bc = m.rbc, *bn = m.rbn, *gn = m.rgn;
    const double s1 = F(ac, i + 1, j) * (F(u, i + 1, j) - F(u, i, j)) - F(ac, i, j) * (F(u, i, j) - F(u, i - 1, j))
                      + F(bc, i + 1, j) * (F(u, i + 1, j + 1) + F(u, i, j + 1) - F(u, i + 1, j - 1) - F(u, i, j - 1))
                      - F(bc, i, j) * (F(u, i, j + 1) + F(u, i - 1, j + 1) - F(u, i, j - 1) - F(u, i - 1, j - 1));
    const double s2 = F(bn, i, j) * (F(u, i + 1, j + 1) + F(u, i + 1, j) - F(u, i - 1, j + 1) - F(u, i - 1, j))
                      - F(bn, i, j - 1) * (F(u, i + 1, j) + F(u, i + 1, j - 1) - F(u, i - 1, j) - F(u, i - 1, j - 1))
                      + F(gn, i, j) * (F(u, i, j + 1) - F(u, i, j)) - F(gn, i, j - 1) * (F(u, i, j) - F(u, i, j - 1));
    return s1 + s2;
}
// DConvV (:1064-1070) and DDiffV (:1094-1106)
__device__ __forceinline__ double y_conv(const MomArgs &m, const double *u, const double *v, int i, int j) {
    const int pitch = m.pitch;
    const double c1 = y_c1(m, u, v, i, j), c1w = y_c1(m, u, v, i - 1, j);
    const double c2 = y_c2(m, u, v, i, j), c2n = y_c2(m, u, v, i, j + 1);
    return -c2 * F(v, i, j - 1) - c1w * F(v, i - 1, j) + (c1 - c1w + c2n - c2) * F(v, i, j)
           + c1 * F(v, i + 1, j) + c2n * F(v, i, j + 1);
}
__device__ __forceinline__ double y_diff(const MomArgs &m, const double *v, int i, int j) {
    const int pitch = m.pitch;
    const double *an = m.ran, *bc = m.rbc, *bn = m.rbn, *gc = m.rgc;
    const double s1 = F(an, i, j) * (F(v, i + 1, j) - F(v, i, j)) - F(an, i - 1, j) * (F(v, i, j) - F(v, i - 1, j))
                      + F(bn, i, j) * (F(v, i + 1, j + 1) + F(v, i, j + 1) - F(v, i + 1, j - 1) - F(v, i, j - 1))
                      - F(bn, i - 1, j) * (F(v, i, j + 1) + F(v, i - 1, j + 1) - F(v, i, j - 1) - F(v, i - 1, j - 1));
    const double s2 = F(bc, i, j + 1) * (F(v, i + 1, j + 1) + F(v, i + 1, j) - F(v, i - 1, j + 1) - F(v, i - 1, j))
                      - F(bc, i, j) * (F(v, i + 1, j) + F(v, i + 1, j - 1) - F(v, i - 1, j) - F(v, i - 1, j - 1))
                      + F(gc, i, j + 1) * (F(v, i, j + 1) - F(v, i, j)) - F(gc, i, j) * (F(v, i, j) - F(v, i, j - 1));
    return s1 + s2;
}

// ---- assembly kernels -------------------------------------------------------------------------
// X, first split step (:350-384): unknown ind = (j-2)*nx + i, i=1..nx, j=2..ny
__global__ void __launch_bounds__(256) xmom_step1_kernel(MomArgs m) {
    const int pitch = m.pitch, nx = m.nx;
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx) return;
    const double re1 = 1.0 / m.re, dk2 = m.dk * 0.5;
    for (int j = 2 + blockIdx.y; j <= m.ny; j += gridDim.y) {
        const size_t ind = (size_t)(j - 2) * nx + (i - 1);
        const double rkj = dk2 * F(m.dju, i, j);
        const double cj = x_cj1(m, i, j);
        const double rac0 = F(m.rac, i, j), rac1 = F(m.rac, i + 1, j);
        double a1, a2, a3;
        if (cj >= 0.0) {
            a1 = rkj * (-x_cj1(m, i - 1, j) - re1 * rac0);
            a2 = 1.0 + rkj * (cj + re1 * (rac1 + rac0));
            a3 = rkj * (-re1 * rac1);
        } else {
            a1 = rkj * (-re1 * rac0);
            a2 = 1.0 + rkj * (-cj + re1 * (rac1 + rac0));
            a3 = rkj * (x_cj1(m, i + 1, j) - re1 * rac1);
        }
        const double cnvs = x_conv(m, m.us, m.vs, i, j), cnvn = x_conv(m, m.un, m.vn, i, j);
        const double difs = x_diff(m, m.us, i, j), difn = x_diff(m, m.un, i, j);
        const double b = F(m.un, i, j) - F(m.us, i, j) + rkj * (-cnvs - cnvn) + rkj * re1 * (difs + difn);
        m.ta[ind] = a1; m.td[ind] = a2; m.tc[ind] = a3; m.tb[ind] = b;
    }
}
// X, second split step LHS (:396-428) + identity rows (:434-496); rhs = first-step solution
__global__ void __launch_bounds__(256) xmom_step2_kernel(MomArgs m) {
    const int pitch = m.pitch, nx = m.nx;
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx) return;
    const double re1 = 1.0 / m.re, dk2 = m.dk * 0.5;
    for (int j = 2 + blockIdx.y; j <= m.ny; j += gridDim.y) {
        const size_t ind = (size_t)(j - 2) * nx + (i - 1);
        double a1, a2, a3, b;
        if (m.xmask[IDX(i, j)]) { a1 = 0.0; a2 = 1.0; a3 = 0.0; b = 0.0; }
        else {
            const double rkj = dk2 * F(m.dju, i, j);
            const double cj = x_cj2(m, i, j);
            const double g0 = F(m.rgn, i, j), gm = F(m.rgn, i, j - 1);
            if (cj >= 0.0) {
                a1 = rkj * (-x_cj2(m, i, j - 1) - re1 * gm);
                a2 = 1.0 + rkj * (cj + re1 * (g0 + gm));
                a3 = rkj * (-re1 * g0);
            } else {
                a1 = rkj * (-re1 * gm);
                a2 = 1.0 + rkj * (-cj + re1 * (g0 + gm));
                a3 = rkj * (x_cj2(m, i, j + 1) - re1 * g0);
            }
            b = m.x1[ind];
        }
        m.ta[ind] = a1; m.td[ind] = a2; m.tc[ind] = a3; m.tb[ind] = b;
    }
}
// Y, first split step (:675-711): unknown ind = (j-1)*(nx-1) + i-1, i=2..nx, j=1..ny
__global__ void __launch_bounds__(256) ymom_step1_kernel(MomArgs m) {
    const int pitch = m.pitch, nx = m.nx;
    const int i = 2 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx) return;
    const double re1 = 1.0 / m.re, dk2 = m.dk * 0.5;
    for (int j = 1 + blockIdx.y; j <= m.ny; j += gridDim.y) {
        const size_t ind = (size_t)(j - 1) * (nx - 1) + (i - 2);
        const double rkj = dk2 * F(m.djv, i, j);
        const double cj = y_cj1(m, i, j);
        const double an0 = F(m.ran, i, j), anm = F(m.ran, i - 1, j);
        double a1, a2, a3;
        if (cj >= 0.0) {
            a1 = rkj * (-y_cj1(m, i - 1, j) - re1 * anm);
            a2 = rkj * (cj + re1 * (an0 + anm)) + 1.0;
            a3 = rkj * (-re1 * an0);
        } else {
            a1 = rkj * (-re1 * anm);
            a2 = rkj * (-cj + re1 * (an0 + anm)) + 1.0;
            a3 = rkj * (y_cj1(m, i + 1, j) - re1 * an0);
        }
        const double buoy = m.dk * (F(m.d, i, j + 1) + F(m.d, i, j) + F(m.dn, i, j + 1) + F(m.dn, i, j)) / (4.0 * m.fr);
        const double cnvs = y_conv(m, m.us, m.vs, i, j), cnvn = y_conv(m, m.un, m.vn, i, j);
        const double difs = y_diff(m, m.vs, i, j), difn = y_diff(m, m.vn, i, j);
        const double b = F(m.vn, i, j) - F(m.vs, i, j) + rkj * (-cnvs - cnvn) + rkj * re1 * (difs + difn) - buoy;
        m.ta[ind] = a1; m.td[ind] = a2; m.tc[ind] = a3; m.tb[ind] = b;
    }
}
// Y, second split step LHS (:723-754) + identity rows (:760-821)
__global__ void __launch_bounds__(256) ymom_step2_kernel(MomArgs m) {
    const int pitch = m.pitch, nx = m.nx;
    const int i = 2 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx) return;
    const double re1 = 1.0 / m.re, dk2 = m.dk * 0.5;
    for (int j = 1 + blockIdx.y; j <= m.ny; j += gridDim.y) {
        const size_t ind = (size_t)(j - 1) * (nx - 1) + (i - 2);
        double a1, a2, a3, b;
        if (m.ymask[IDX(i, j)]) { a1 = 0.0; a2 = 1.0; a3 = 0.0; b = 0.0; }
        else {
            const double rkj = dk2 * F(m.djv, i, j);
            const double cj = y_cj2(m, i, j);
            const double g0 = F(m.rgc, i, j), gp = F(m.rgc, i, j + 1);
            if (cj >= 0.0) {
                a1 = rkj * (-y_cj2(m, i, j - 1) - re1 * g0);
                a2 = rkj * (cj + re1 * (gp + g0)) + 1.0;
                a3 = rkj * (-re1 * gp);
            } else {
                a1 = rkj * (-re1 * g0);
                a2 = rkj * (-cj + re1 * (gp + g0)) + 1.0;
                a3 = rkj * (y_cj2(m, i, j + 1) - re1 * gp);
            }
            b = m.x1[ind];
        }
        m.ta[ind] = a1; m.td[ind] = a2; m.tc[ind] = a3; m.tb[ind] = b;
    }
}

// identity-row masks of the second split step (momentum.f:434-496 for u, :760-821 for v)
__global__ void mom_mask_kernel(const W2Regions *__restrict__ R, int nx, int ny, int pitch,
                                unsigned char *__restrict__ xmask, unsigned char *__restrict__ ymask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i > nx + 1 || j > ny + 1) return;
    unsigned char mx = 0, my = 0;
    for (int q = 0; q < R->nreg; ++q) {
        const int iW = R->iW[q], iE = R->iE[q], jS = R->jS[q], jN = R->jN[q];
        const bool blk = R->type[q] == W2_RM_BLOCKG;
        const int bW = R->bd[q][0], bE = R->bd[q][1], bS = R->bd[q][2], bN = R->bd[q][3];
        const bool wW = bW == W2_BM_WALL1 || bW == W2_BM_WALL2 || bW == W2_BM_INLET;
        const bool wE = bE == W2_BM_WALL1 || bE == W2_BM_WALL2 || bE == W2_BM_INLET;
        const bool wS = bS == W2_BM_WALL1 || bS == W2_BM_WALL2 || bS == W2_BM_INLET;
        const bool wN = bN == W2_BM_WALL1 || bN == W2_BM_WALL2 || bN == W2_BM_INLET;
        if (j >= jS + 1 && j <= jN) {
            if (blk && i >= iW && i <= iE) mx = 1;
            if (wW && i == iW) mx = 1;
            if (wE && i == iE) mx = 1;
        }
        if (i >= iW + 1 && i <= iE) {
            if (blk && j >= jS && j <= jN) my = 1;
            if (wS && j == jS) my = 1;
            if (wN && j == jN) my = 1;
        }
    }
    xmask[IDX(i, j)] = mx;
    ymask[IDX(i, j)] = my;
}

// chain -> field scatter (:505-510, :829-834)
__global__ void __launch_bounds__(256) scatter_x_kernel(int nx, int ny, int pitch, const double *__restrict__ x,
                                                        double *__restrict__ dus) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx) return;
    for (int j = 2 + blockIdx.y; j <= ny; j += gridDim.y) dus[IDX(i, j)] = x[(size_t)(j - 2) * nx + (i - 1)];
}
__global__ void __launch_bounds__(256) scatter_y_kernel(int nx, int ny, int pitch, const double *__restrict__ x,
                                                        double *__restrict__ dvs) {
    const int i = 2 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx) return;
    for (int j = 1 + blockIdx.y; j <= ny; j += gridDim.y) dvs[IDX(i, j)] = x[(size_t)(j - 1) * (nx - 1) + (i - 2)];
}

// us,vs <- un,vn on 1..nx+1, 1..ny+1 (:114-119)
__global__ void __launch_bounds__(256) ql_init_kernel(int nx, int ny, int pitch, const double *__restrict__ un,
                                                      const double *__restrict__ vn, double *__restrict__ us,
                                                      double *__restrict__ vs) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx + 1) return;
    for (int j = 1 + blockIdx.y; j <= ny + 1; j += gridDim.y) {
        us[IDX(i, j)] = un[IDX(i, j)];
        vs[IDX(i, j)] = vn[IDX(i, j)];
    }
}

// us += dus, vs += dvs on 1..nx,1..ny (:171-176) fused with the two DMaxNorm scans (:179-180)
__global__ void __launch_bounds__(256) ql_update_kernel(int nx, int ny, int pitch, const double *__restrict__ dus,
                                                        const double *__restrict__ dvs, double *__restrict__ us,
                                                        double *__restrict__ vs, unsigned long long *slots) {
    __shared__ double red[32];
    double mu = 0.0, mv = 0.0;
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= nx)
        for (int j = 1 + blockIdx.y; j <= ny; j += gridDim.y) {
            const double du = dus[IDX(i, j)], dv = dvs[IDX(i, j)];
            us[IDX(i, j)] = us[IDX(i, j)] + du;
            vs[IDX(i, j)] = vs[IDX(i, j)] + dv;
            if (i >= 2 && i <= nx - 1 && j >= 2 && j <= ny - 1) { mu = fmax(mu, fabs(du)); mv = fmax(mv, fabs(dv)); }
        }
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {  // DMaxNorm seed |u(5,5)| (utility.f:493)
        mu = fmax(mu, fabs(dus[IDX(5, 5)]));
        mv = fmax(mv, fabs(dvs[IDX(5, 5)]));
    }
    mu = w2_block_max(mu, red);
    mv = w2_block_max(mv, red);
    if (threadIdx.x == 0) { atomicMax(slots + 0, w2_dbits(mu)); atomicMax(slots + 1, w2_dbits(mv)); }
}

// ---- host side -----------------------------------------------------------------------------------
int w2_build_mom_masks(wolfd2_ctx *c) {
    dim3 grid((c->nx + 2 + 255) / 256, c->ny + 2);
    mom_mask_kernel<<<grid, 256, 0, c->stream>>>(c->dreg, c->nx, c->ny, c->pitch, c->xmask, c->ymask);
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}

static void fill_args(wolfd2_ctx *c, MomArgs &m) {
    m.nx = c->nx; m.ny = c->ny; m.pitch = c->pitch;
    m.dk = c->par.dk; m.re = c->par.re; m.fr = c->par.fr;
    m.us = c->fld[W2_F_US]; m.vs = c->fld[W2_F_VS]; m.un = c->fld[W2_F_UN]; m.vn = c->fld[W2_F_VN];
    m.d = c->fld[W2_F_D]; m.dn = c->fld[W2_F_DN];
    const W2Metrics &t = c->met;
    m.rbn = t.rbn; m.rgn = t.rgn; m.rac = t.rac; m.rbc = t.rbc; m.dju = t.dju;
    m.xec = t.xec; m.yec = t.yec; m.xzn = t.xzn; m.yzn = t.yzn;
    m.xeu = t.xeu; m.yeu = t.yeu; m.xzu = t.xzu; m.yzu = t.yzu;
    m.ran = t.ran; m.rgc = t.rgc; m.djv = t.djv;
    m.xen = t.xen; m.yen = t.yen; m.xzc = t.xzc; m.yzc = t.yzc;
    m.xev = t.xev; m.yev = t.yev; m.xzv = t.xzv; m.yzv = t.yzv;
    m.xmask = c->xmask; m.ymask = c->ymask;
    m.ta = c->ta; m.td = c->td; m.tc = c->tc; m.tb = c->tb; m.x1 = c->tx;
}

static int check_porous(wolfd2_ctx *c) {
    if (c->hreg.has_porous) {
        w2_set_error("porous regions (RM_POROUS, PorosCoef momentum.f:1115-1226) are not implemented on the device yet");
        return W2_ERR_UNSUPPORTED;
    }
    return W2_OK;
}

int w2_xmomentum(wolfd2_ctx *c, double *dus) {
    W2_TRY(check_porous(c));
    MomArgs m;
    fill_args(c, m);
    const int nx = c->nx, ny = c->ny;
    const long long n = (long long)nx * (ny - 1);
    dim3 grid((nx + 255) / 256, (ny - 1) < 2048 ? (ny - 1) : 2048);
    xmom_step1_kernel<<<grid, 256, 0, c->stream>>>(m);
    c->launches[1]++;
    W2_TRY(w2_tri_solve(c, n, c->ta, c->td, c->tc, c->tb, c->tx, 1));  // :389
    xmom_step2_kernel<<<grid, 256, 0, c->stream>>>(m);
    c->launches[1]++;
    W2_TRY(w2_tri_solve(c, n, c->ta, c->td, c->tc, c->tb, c->tx, 1));  // :501
    scatter_x_kernel<<<grid, 256, 0, c->stream>>>(nx, ny, c->pitch, c->tx, dus);
    c->launches[1]++;
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}

int w2_ymomentum(wolfd2_ctx *c, double *dvs) {
    W2_TRY(check_porous(c));
    MomArgs m;
    fill_args(c, m);
    const int nx = c->nx, ny = c->ny;
    const long long n = (long long)(nx - 1) * ny;
    dim3 grid((nx - 1 + 255) / 256, ny < 2048 ? ny : 2048);
    ymom_step1_kernel<<<grid, 256, 0, c->stream>>>(m);
    c->launches[1]++;
    W2_TRY(w2_tri_solve(c, n, c->ta, c->td, c->tc, c->tb, c->tx, 1));  // :716
    ymom_step2_kernel<<<grid, 256, 0, c->stream>>>(m);
    c->launches[1]++;
    W2_TRY(w2_tri_solve(c, n, c->ta, c->td, c->tc, c->tb, c->tx, 1));  // :826
    scatter_y_kernel<<<grid, 256, 0, c->stream>>>(nx, ny, c->pitch, c->tx, dvs);
    c->launches[1]++;
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}

// nAuxMomentum (:33-193).  One host read-back (16 bytes) per QL iteration decides convergence.
int w2_nauxmomentum(wolfd2_ctx *c, int init_star, int *nQLiter) {
    const int nx = c->nx, ny = c->ny;
    double *us = c->fld[W2_F_US], *vs = c->fld[W2_F_VS];
    *nQLiter = -1;  // :111
    if (init_star) {
        dim3 g((nx + 1 + 255) / 256, (ny + 1) < 2048 ? (ny + 1) : 2048);
        ql_init_kernel<<<g, 256, 0, c->stream>>>(nx, ny, c->pitch, c->fld[W2_F_UN], c->fld[W2_F_VN], us, vs);
        c->launches[1]++;
    }
    for (int m = 1; m <= c->par.mqiter; ++m) {
        W2_TRY(w2_outflow_bc(c, us, vs));  // :133
        // dus, dvs are zero outside the ranges XMomentum/YMomentum write (:139-144 re-zeroes the same cells)
        W2_TRY(w2_xmomentum(c, c->dus));   // :147
        W2_TRY(w2_ymomentum(c, c->dvs));   // :158
        W2_CUDA(cudaMemsetAsync(c->d_norm + 8, 0, 2 * sizeof(unsigned long long), c->stream));
        dim3 g((nx + 255) / 256, ny < 2048 ? ny : 2048);
        ql_update_kernel<<<g, 256, 0, c->stream>>>(nx, ny, c->pitch, c->dus, c->dvs, us, vs, c->d_norm + 8);
        c->launches[1]++;
        W2_CUDA(cudaGetLastError());
        W2_CUDA(cudaMemcpyAsync(c->h_norm + 8, c->d_norm + 8, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        W2_CUDA(cudaStreamSynchronize(c->stream));
        double dif[2];
        memcpy(dif, c->h_norm + 8, 16);
        const double difmax = dif[0] > dif[1] ? dif[0] : dif[1];  // :182
        if (difmax <= c->par.qtol) { *nQLiter = m; return W2_OK; }  // :185-188
    }
    return W2_OK;
}
