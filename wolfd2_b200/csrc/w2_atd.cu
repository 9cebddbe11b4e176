// w2_atd.cu -- ATD small-scale model on the device: SmallScale (src/small_scale.f:31-581)
// and SmlSclBC (src/bound_cond.f:1209-1649).  One GPU only.
//
// SmallScale is per-cell independent once its `save`d state exists: the three chaotic maps per variable
// (umap, vmap, tmap, seeded by ONE sequential recurrence over the whole grid at :203-220), the cell areas and
// their sum tArea (:223-236, accumulated i-outer / j-inner).  Both are set-up work and order-dependent, so they
// are formed on the host exactly as written and uploaded once (w2_smallscale_seed); everything per step runs
// here: unscale + contravariant components (:252-279), TempBoundCond + three Shuman filters (:284-305), the
// high-pass (:309-315), the per-cell map iteration and amplitude model (:319-524), SmlSclBC, and the model's
// own Ppe / Project (:527-577), which reuse the large-scale kernels with the atd_* solver parameters.
//
// Arithmetic: +,-,*,/ and sqrt are IEEE and in the reference's order (the library is built -fmad=false).
// x**y with a REAL exponent, dtanh and datanh are CUDA's pow/tanh (<= 2 ulp) instead of glibc's (< 1 ulp):
// the chaotic map amplifies that difference by up to |4.83| per iterate, so parity of uss, vss, tss against
// a CPU build is a tolerance statement that weakens with nmap x steps; cell-independent
// constants (pi/sqrt 2, datanh(rc/rmax), sqrt 15, cuT**4 ...) are computed on the host with glibc.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "w2.cuh"

#define BC_THREADS 1024
#define U(i, j) u[IDX(i, j)]
#define V(i, j) v[IDX(i, j)]
#define T(i, j) t[IDX(i, j)]
#define PFOR(var, lo, hi) for (int var = (lo) + (int)threadIdx.x; var <= (hi); var += BC_THREADS)
#define SEQ if (threadIdx.x == 0)

// ---------------------------------------------------------------------------------- SmlSclBC
// velocity part (:1257-1512) and temperature part (:1527-1644); PresBoundCond(p) runs between them (:1519)
// as its own launch.  Same walking order and barrier discipline as vel_bc_kernel (w2_bc.cu).
__global__ void __launch_bounds__(BC_THREADS) smlscl_vel_bc_kernel(const W2Regions *__restrict__ R, int pitch, double *u, double *v) {
    const double dZero = 0.0, dThree = 3.0, dFour = 4.0, dFive = 5.0, dEight = 8.0;
    const int nreg = R->nreg;
    for (int q = 0; q < nreg; ++q) {
        const int iW = R->iW[q], iE = R->iE[q], jS = R->jS[q], jN = R->jN[q];
        int bt = R->bd[q][W2_WEST - 1];   // :1270-1319
        if (bt == W2_BM_WALL1 || bt == W2_BM_INLET || bt == W2_BM_WALL2) {
            PFOR(j, jS, jN) U(iW, j) = dZero;
            __syncthreads();
            if (bt == W2_BM_WALL2) { PFOR(j, jS + 1, jN) V(iW, j) = V(iW + 1, j); }
            else { PFOR(j, jS + 1, jN) V(iW, j) = -V(iW + 1, j); }
        } else if (bt == W2_BM_OUTLT1) {
            PFOR(j, jS, jN) U(iW - 1, j) = +U(iW, j);
            __syncthreads();
            PFOR(j, jS + 1, jN) V(iW, j) = -V(iW + 1, j);
        } else if (bt == W2_BM_OUTLT2) {
            PFOR(j, jS, jN) U(iW, j) = U(iW + 1, j) - V(iW, j) + V(iW, j - 1);
            __syncthreads();
            SEQ {
                double prev = V(iW, jS);
                for (int j = jS + 1; j <= jN; ++j) {
                    prev = -prev + dFive * (V(iW + 1, j) - V(iW + 1, j - 1)) + dEight * (U(iW + 1, j) - U(iW, j));
                    V(iW, j) = prev;
                }
            }
        }
        __syncthreads();
        bt = R->bd[q][W2_EAST - 1];       // :1323-1388
        if (bt == W2_BM_WALL1 || bt == W2_BM_INLET || bt == W2_BM_WALL2) {
            PFOR(j, jS, jN) U(iE, j) = dZero;
            __syncthreads();
            if (bt == W2_BM_WALL2) { PFOR(j, jS + 1, jN) V(iE + 1, j) = V(iE, j); }
            else { PFOR(j, jS + 1, jN) V(iE + 1, j) = -V(iE, j); }
        } else if (bt == W2_BM_OUTLT1) {
            // u(iE+1,j) = +u(iE+1,j): a self-assignment as written (:1350), nothing to do
            PFOR(j, jS + 1, jN) V(iE + 1, j) = -V(iE, j);
        } else if (bt == W2_BM_OUTLT2) {
            PFOR(j, jS + 1, jN) U(iE, j) = U(iE - 1, j) - (V(iE, j) - V(iE, j - 1));
            __syncthreads();
            SEQ {
                double prev = V(iE + 1, jS);
                for (int j = jS + 1; j <= jN - 1; ++j) {
                    prev = prev + dThree * (V(iE, j - 1) - V(iE, j)) - dFour * (U(iE, j) - U(iE - 1, j));
                    V(iE + 1, j) = prev;
                }
            }
        }
        __syncthreads();
        bt = R->bd[q][W2_SOUTH - 1];      // :1392-1441
        if (bt == W2_BM_WALL1 || bt == W2_BM_INLET || bt == W2_BM_WALL2) {
            if (bt == W2_BM_WALL2) { PFOR(i, iW + 1, iE) U(i, jS) = U(i, jS + 1); }
            else { PFOR(i, iW + 1, iE) U(i, jS) = -U(i, jS + 1); }
            __syncthreads();
            PFOR(i, iW, iE) V(i, jS) = dZero;
        } else if (bt == W2_BM_OUTLT1) {
            PFOR(i, iW + 1, iE) U(i, jS) = -U(i, jS + 1);
            __syncthreads();
            PFOR(i, iW, iE) V(i, jS - 1) = +V(i, jS);
        } else if (bt == W2_BM_OUTLT2) {
            SEQ {   // reads v(i-1,jS-1) before the v loop below rewrites that row
                double prev = U(iW, jS - 1);
                for (int i = iW + 1; i <= iE; ++i) {
                    prev = prev + dFive * (U(i, jS) - U(i - 1, jS)) + dEight * (V(i, jS) - V(i - 1, jS - 1));
                    U(i, jS - 1) = prev;
                }
            }
            __syncthreads();
            PFOR(i, iW, iE) V(i, jS - 1) = V(i, jS) - U(i, jS) - U(i - 1, jS);
        }
        __syncthreads();
        bt = R->bd[q][W2_NORTH - 1];      // :1445-1510
        if (bt == W2_BM_WALL1 || bt == W2_BM_INLET || bt == W2_BM_WALL2) {
            if (bt == W2_BM_WALL2) { PFOR(i, iW + 1, iE) U(i, jN + 1) = U(i, jN); }
            else { PFOR(i, iW + 1, iE) U(i, jN + 1) = -U(i, jN); }
            __syncthreads();
            PFOR(i, iW, iE) V(i, jN) = dZero;
        } else if (bt == W2_BM_OUTLT1) {
            PFOR(i, iW + 1, iE) U(i, jN + 1) = -U(i, jN);
            __syncthreads();
            PFOR(i, iW, iE) V(i, jN + 1) = +V(i, jN);
        } else if (bt == W2_BM_OUTLT2) {
            PFOR(i, iW, iE) V(i, jN) = V(i, jN - 1) - (U(i, jN) - U(i - 1, jN));
            __syncthreads();
            SEQ {
                double prev = U(iW, jN + 1);
                for (int i = iW + 1; i <= iE - 1; ++i) {
                    prev = prev + dThree * (U(i - 1, jN) - U(i, jN)) - dFour * (V(i, jN) - V(i, jN - 1));
                    U(i, jN + 1) = prev;
                }
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(BC_THREADS) smlscl_temp_bc_kernel(const W2Regions *__restrict__ R, const W2Thermal *__restrict__ H,
                                                                   int pitch, double *t) {
    const double dZero = 0.0;
    const int nreg = R->nreg;
    for (int q = 0; q < nreg; ++q) {
        const int iW = R->iW[q], iE = R->iE[q], jS = R->jS[q], jN = R->jN[q];
        if (H->ttype[q] == W2_RT_TEMPER) {   // :1540-1570
            const int w = iE - iW, h = jN - jS;
            for (int k = threadIdx.x; k < w * h; k += BC_THREADS) T(iW + 1 + k % w, jS + 1 + k / w) = dZero;
            __syncthreads();
            PFOR(j, jS + 1, jN) T(iW + 1, j) = -T(iW, j);
            __syncthreads();
            PFOR(j, jS + 1, jN) T(iE, j) = -T(iE + 1, j);
            __syncthreads();
            PFOR(i, iW + 1, iE) T(i, jS + 1) = -T(i, jS);
            __syncthreads();
            PFOR(i, iW + 1, iE) T(i, jN) = -T(i, jN + 1);
            __syncthreads();
            continue;
        }
        int bt = H->tbd[q][W2_WEST - 1];
        if (bt == W2_BT_TEMPER) { PFOR(j, jS + 1, jN) T(iW, j) = -T(iW + 1, j); }
        else if (bt == W2_BT_HTFLUX) { PFOR(j, jS + 1, jN) T(iW, j) = T(iW + 1, j); }
        __syncthreads();
        bt = H->tbd[q][W2_EAST - 1];
        if (bt == W2_BT_TEMPER) { PFOR(j, jS + 1, jN) T(iE + 1, j) = -T(iE, j); }
        else if (bt == W2_BT_HTFLUX) { PFOR(j, jS + 1, jN) T(iE + 1, j) = T(iE, j); }
        __syncthreads();
        bt = H->tbd[q][W2_SOUTH - 1];
        if (bt == W2_BT_TEMPER) { PFOR(i, iW + 1, iE) T(i, jS) = -T(i, jS + 1); }
        else if (bt == W2_BT_HTFLUX) { PFOR(i, iW + 1, iE) T(i, jS) = T(i, jS + 1); }
        __syncthreads();
        bt = H->tbd[q][W2_NORTH - 1];
        if (bt == W2_BT_TEMPER) { PFOR(i, iW + 1, iE) T(i, jN + 1) = -T(i, jN); }
        else if (bt == W2_BT_HTFLUX) { PFOR(i, iW + 1, iE) T(i, jN + 1) = T(i, jN); }
        __syncthreads();
    }
}

int w2_smlscl_bc(wolfd2_ctx *c, double *u, double *v, double *p, double *t) {
    smlscl_vel_bc_kernel<<<1, BC_THREADS, 0, c->stream>>>(c->dreg, c->pitch, u, v);
    W2_TRY(w2_pres_bc(c, p));   // :1519 "same boundary conditions as large-scale variables"
    smlscl_temp_bc_kernel<<<1, BC_THREADS, 0, c->stream>>>(c->dreg, c->dth, c->pitch, t);
    c->launches[3] += 2;
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}

// ---------------------------------------------------------------------------------- SmallScale
struct SsConst {   // cell-independent scalars, formed on the host (glibc) in the reference's order
    double dlref, uref, tref, tmax;
    double rdl2;          // dlref*dlref
    double rnu, pr, dk;   // :188-195
    double pehmin;
    double cu0, TemCoef;
    int iterate;          // cu0 > 1.e-10 (:371)
    double ts_num;        // TsCoef*piosr2 (:377)
    double sqrt15, hs2;   // :390
    double bnumc, rlc, rmax, atanh_rc;   // :393
    double sqrt2;
    double cuT4_3;        // dThree*cuT**dFour (:496)
    double sq_kappa;      // dsqrt(dkappa)
    double tArea, dtspan; // tmax - tref
};

// :252-279
__global__ void __launch_bounds__(256) ss_prepare_kernel(int nx, int ny, int pitch, double dlref, double uref, double tref, double tmax,
                                                         const double *__restrict__ u1, const double *__restrict__ v1,
                                                         const double *__restrict__ t1, const double *__restrict__ xec,
                                                         const double *__restrict__ yec, const double *__restrict__ xzc,
                                                         const double *__restrict__ yzc, double *__restrict__ ul,
                                                         double *__restrict__ vl, double *__restrict__ tl, double *__restrict__ uf,
                                                         double *__restrict__ vf, double *__restrict__ tf) {
    const double dHalf = 0.5;
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx) return;
    for (int j = 1 + blockIdx.y; j <= ny; j += gridDim.y) {
        const size_t o = IDX(i, j);
        const double xz = dlref * xzc[o], xe = dlref * xec[o], yz = dlref * yzc[o], ye = dlref * yec[o];
        const double vlv = uref * (v1[o] + v1[IDX(i, j - 1)]) * dHalf;
        const double ulv = uref * (u1[o] + u1[IDX(i - 1, j)]) * dHalf;
        const double tlv = (tmax - tref) * t1[o] + tref;
        vl[o] = vlv; ul[o] = ulv; tl[o] = tlv;
        uf[o] = ye * ulv - xe * vlv;
        vf[o] = xz * vlv - yz * ulv;
        tf[o] = tlv;
    }
}

// :309-315: uf <- uc - filtered(uc); uc, vc are re-formed from ul, vl with the operations of :270-271
__global__ void __launch_bounds__(256) ss_highpass_kernel(int nx, int ny, int pitch, double dlref, const double *__restrict__ xec,
                                                          const double *__restrict__ yec, const double *__restrict__ xzc,
                                                          const double *__restrict__ yzc, const double *__restrict__ ul,
                                                          const double *__restrict__ vl, const double *__restrict__ tl,
                                                          double *__restrict__ uf, double *__restrict__ vf, double *__restrict__ tf) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx) return;
    for (int j = 1 + blockIdx.y; j <= ny; j += gridDim.y) {
        const size_t o = IDX(i, j);
        const double xz = dlref * xzc[o], xe = dlref * xec[o], yz = dlref * yzc[o], ye = dlref * yec[o];
        const double uc = ye * ul[o] - xe * vl[o];
        const double vc = xz * vl[o] - yz * ul[o];
        uf[o] = uc - uf[o];
        vf[o] = vc - vf[o];
        tf[o] = tl[o] - tf[o];
    }
}

struct SsFields {
    const double *xec, *yec, *xzc, *yzc, *djc, *ul, *vl, *tl, *uf, *vf, *tf;
    double *map[9];   // umap(.,.,1..3), vmap, tmap
    double *uss, *vss, *tss;
};

// :329-521 for the cells (ilo..ihi, jlo..jhi) of one region that is neither a blockage nor RT_TEMPER
__global__ void __launch_bounds__(128) ss_cell_kernel(int pitch, int ilo, int ihi, int jlo, int jhi, SsConst K, SsFields F) {
    const double dZero = 0.0, dOne = 1.0, dThree = 3.0, dSix = 6.0, dHalf = 0.5;
    const double dAr = 4.82842712474e+00, dAm = 1.47839783948e+00;
    const int i = ilo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > ihi) return;
    for (int j = jlo + blockIdx.y; j <= jhi; j += gridDim.y) {
        const size_t o = IDX(i, j);
        const double XZ = K.dlref * F.xzc[o], XE = K.dlref * F.xec[o], YZ = K.dlref * F.yzc[o], YE = K.dlref * F.yec[o];
        const double djc = F.djc[o];
        const double RJ = djc / K.rdl2;
        const double hxy = sqrt(XZ * XZ + YZ * YZ + XE * XE + YE * YE);
        double uz = (F.ul[IDX(i + 1, j)] - F.ul[IDX(i - 1, j)]) * dHalf;
        double ue = (F.ul[IDX(i, j + 1)] - F.ul[IDX(i, j - 1)]) * dHalf;
        double vz = (F.vl[IDX(i + 1, j)] - F.vl[IDX(i - 1, j)]) * dHalf;
        double ve = (F.vl[IDX(i, j + 1)] - F.vl[IDX(i, j - 1)]) * dHalf;
        const double uxs = RJ * (YE * uz - YZ * ue), uys = RJ * (XZ * ue - XE * uz);
        const double vxs = RJ * (YE * vz - YZ * ve), vys = RJ * (XZ * ve - XE * vz);
        const double delu2n = sqrt(uxs * uxs + uys * uys + vxs * vxs + vys * vys);
        double tz = (F.tl[IDX(i + 1, j)] - F.tl[IDX(i - 1, j)]) * dHalf;
        double te = (F.tl[IDX(i, j + 1)] - F.tl[IDX(i, j - 1)]) * dHalf;
        const double txs = RJ * (YE * tz - YZ * te), tys = RJ * (XZ * te - XE * tz);
        const double delt2n = sqrt(txs * txs + tys * tys);
        const double reh = delu2n * (hxy * hxy) / K.rnu;
        const double peh = K.pr * reh;
        if (!(peh > K.pehmin)) continue;
        int nmap = 0;
        if (K.iterate) {
            const double ts = K.ts_num * pow(reh, dOne / dThree) / (K.cu0 * delu2n);
            const double q = dOne + K.dk / ts;
            nmap = q >= 2147483647.0 ? 2147483647 : (int)q;
        }
        if (nmap > 50) nmap = 50;
        const double bnum = K.sqrt15 * pow(K.hs2 * delu2n / K.rnu, dOne / dSix);
        const double rmap = K.rmax * tanh(pow(bnum / K.bnumc, K.rlc) * K.atanh_rc);
        uz = (F.uf[IDX(i + 1, j)] - F.uf[IDX(i - 1, j)]) * dHalf;
        ue = (F.uf[IDX(i, j + 1)] - F.uf[IDX(i, j - 1)]) * dHalf;
        vz = (F.vf[IDX(i + 1, j)] - F.vf[IDX(i - 1, j)]) * dHalf;
        ve = (F.vf[IDX(i, j + 1)] - F.vf[IDX(i, j - 1)]) * dHalf;
        tz = (F.tf[IDX(i + 1, j)] - F.tf[IDX(i - 1, j)]) * dHalf;
        te = (F.tf[IDX(i, j + 1)] - F.tf[IDX(i, j - 1)]) * dHalf;
        const double grduf1 = RJ * (YE * uz - YZ * ue), grduf2 = RJ * (XZ * ue - XE * uz);
        const double grdvf1 = RJ * (YE * vz - YZ * ve), grdvf2 = RJ * (XZ * ve - XE * vz);
        const double grdtf1 = RJ * (YE * tz - YZ * te), grdtf2 = RJ * (XZ * te - XE * tz);
        const double grduf = sqrt(grduf1 * grduf1 + grduf2 * grduf2);
        const double grdvf = sqrt(grdvf1 * grdvf1 + grdvf2 * grdvf2);
        const double grdtf = sqrt(grdtf1 * grdtf1 + grdtf2 * grdtf2);
        const double grd = sqrt(grduf * grduf + grdvf * grdvf + grdtf * grdtf);
        const double s1 = grduf / grd, s2 = grdvf / grd;
        const double ra = XZ * s1 + XE * s2, rb = YZ * s1 + YE * s2;
        const double rnrmjs = sqrt(ra * ra + rb * rb);
        const double zeta1 = K.sqrt2 * s1 / rnrmjs, zeta2 = K.sqrt2 * s2 / rnrmjs;
        double alf11 = dZero, alf12 = dZero, alf21 = dZero, alf22 = dZero, alf31 = dZero, alf32 = dZero, alf33 = dZero;
        if (fabs(grduf) > dZero) { alf11 = grduf1 / grduf; alf12 = grduf2 / grduf; }
        if (fabs(grdvf) > dZero) { alf21 = grdvf1 / grdvf; alf22 = grdvf2 / grdvf; }
        if (fabs(grdtf) > dZero) {
            alf31 = grduf1 / grdtf; alf32 = grduf2 / grdtf;
            alf33 = sqrt(grdtf1 * grdtf1 + grdtf2 * grdtf2) / grdtf;
        }
        double m[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) m[q] = F.map[q][o];
        for (int k = 1; k <= nmap; ++k) {
#pragma unroll
            for (int q = 0; q < 9; ++q) m[q] = rmap * dAr * m[q] * (dOne - dAm * fabs(m[q]));
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) F.map[q][o] = m[q];
        const double um = alf11 * m[0] + alf12 * m[1];
        const double vm = alf21 * m[0] + alf22 * m[1];   // sic: um1, um2 (:490)
        const double tm = alf31 * m[6] + alf32 * m[7] + alf33 * m[8];
        double av = K.cu0 * pow(reh, dOne / dSix) * sqrt(K.rnu * delu2n);
        double at = pow(K.cuT4_3 * peh / K.pr, dOne / dSix) * K.sq_kappa;
        at = at * delt2n / sqrt(delu2n) * K.TemCoef;
        const double sc = sqrt((dOne / djc) / K.tArea);
        const double h13 = pow(hxy, dOne / dThree);
        av = av * sc * h13;
        at = at * sc * h13;
        const double uscon = av * zeta1 * um;
        const double vscon = av * zeta2 * vm;
        double us = RJ * (XZ * uscon + YE * vscon) / K.uref;
        double vs = RJ * (YE * vscon + YZ * uscon) / K.uref;
        double tsv = (at * tm * dOne) / K.dtspan;
        if (fabs(us) < 1.e-14) us = dZero;
        if (fabs(vs) < 1.e-14) vs = dZero;
        if (fabs(tsv) < 1.e-14) tsv = dZero;
        F.uss[o] = us; F.vss[o] = vs; F.tss[o] = tsv;
    }
}

// pss(i,j) = 0 on i = 0..nx, j = 0..ny (:538-542)
__global__ void __launch_bounds__(256) ss_zero_p_kernel(int nx, int ny, int pitch, double *__restrict__ p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx) return;
    for (int j = blockIdx.y; j <= ny; j += gridDim.y) p[IDX(i, j)] = 0.0;
}

// a += s*b on 0..nx+1, 0..ny+1 (the whole-array loops of main.f:711-727, :901-907, :934-940); s = +-1
__global__ void __launch_bounds__(256) ss_axpy3_kernel(int nx, int ny, int pitch, double s, double *__restrict__ a0, const double *__restrict__ b0,
                                                       double *__restrict__ a1, const double *__restrict__ b1, double *__restrict__ a2,
                                                       const double *__restrict__ b2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx + 1) return;
    for (int j = blockIdx.y; j <= ny + 1; j += gridDim.y) {
        const size_t o = IDX(i, j);
        if (s > 0.0) { a0[o] = a0[o] + b0[o]; a1[o] = a1[o] + b1[o]; a2[o] = a2[o] + b2[o]; }
        else { a0[o] = a0[o] - b0[o]; a1[o] = a1[o] - b1[o]; a2[o] = a2[o] - b2[o]; }
    }
}
int w2_axpy3(wolfd2_ctx *c, double s, double *a0, const double *b0, double *a1, const double *b1, double *a2, const double *b2) {
    dim3 g((c->nx + 2 + 255) / 256, c->ny + 2 < 2048 ? c->ny + 2 : 2048);
    ss_axpy3_kernel<<<g, 256, 0, c->stream>>>(c->nx, c->ny, c->pitch, s, a0, b0, a1, b1, a2, b2);
    c->launches[3]++;
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}

static int atd_alloc(wolfd2_ctx *c) {
    if (c->atd) return W2_OK;
    W2Atd *a = (W2Atd *)calloc(1, sizeof(W2Atd));
    if (!a) return W2_ERR_BAD_ARG;
    c->atd = a;
    double **f[] = {&c->fld[W2_F_USS], &c->fld[W2_F_VSS], &c->fld[W2_F_PSS], &c->fld[W2_F_TSS], &a->usn, &a->vsn, &a->tsn,
                    &a->ul, &a->vl, &a->tl, &a->uf, &a->vf, &a->tf};
    for (size_t k = 0; k < sizeof(f) / sizeof(f[0]); ++k)
        if (!*f[k]) W2_TRY(w2_alloc_field(c, f[k]));
    for (int k = 0; k < 9; ++k) W2_TRY(w2_alloc_field(c, &a->map[k]));
    return W2_OK;
}
void w2_atd_release(wolfd2_ctx *c) {
    if (!c->atd) return;
    W2Atd *a = c->atd;
    double *f[] = {a->usn, a->vsn, a->tsn, a->ul, a->vl, a->tl, a->uf, a->vf, a->tf};
    for (size_t k = 0; k < sizeof(f) / sizeof(f[0]); ++k) if (f[k]) cudaFree(f[k] + c->row_off);
    for (int k = 0; k < 9; ++k) if (a->map[k]) cudaFree(a->map[k] + c->row_off);
    free(a);
    c->atd = nullptr;
}

// :198-236: map seeds and cell areas.  Sequential by construction (one recurrence runs through all cells,
// i outer / j inner), set-up only: done on the host as written, then uploaded.
static int w2_smallscale_seed(wolfd2_ctx *c) {
    W2Atd *a = c->atd;
    const int nx = c->nx, ny = c->ny;
    const size_t ld = (size_t)c->mnx + 1, plane = ld * ((size_t)c->mny + 1);
    const double dOne = 1.0, dAr = 4.82842712474e+00, dAm = 1.47839783948e+00, rc = 0.20710678119e+00;
    const double seed[3] = {0.92, 0.31, 0.50};   // umpsd, vmpsd, tmpsd (:178-180)
    double *h = (double *)calloc(3 * plane, sizeof(double));
    if (!h) { w2_set_error("smallscale: out of host memory for the map seeds"); return W2_ERR_BAD_ARG; }
    for (int fam = 0; fam < 3; ++fam) {   // the u, v, t recurrences do not interact
        double mp = seed[fam];
        for (int i = 0; i <= nx + 1; ++i)
            for (int j = 0; j <= ny + 1; ++j)
                for (int l = 0; l < 3; ++l) {
                    mp = rc * dAr * mp * (dOne - dAm * fabs(mp));
                    h[(size_t)l * plane + (size_t)i + ld * (size_t)j] = mp;
                }
        for (int l = 0; l < 3; ++l) W2_TRY(w2_upload2d(c, a->map[3 * fam + l], h + (size_t)l * plane));
        W2_CUDA(cudaStreamSynchronize(c->stream));
    }
    // tArea = sum of 1/djc in the reference's order
    W2_TRY(w2_download2d(c, h, c->met.djc));
    W2_CUDA(cudaStreamSynchronize(c->stream));
    double tArea = 0.0;
    for (int i = 1; i <= nx; ++i)
        for (int j = 1; j <= ny; ++j) tArea = tArea + dOne / h[(size_t)i + ld * (size_t)j];
    a->tArea = tArea;
    free(h);
    W2_CUDA(cudaMemsetAsync(c->fld[W2_F_USS] + c->row_off, 0, c->nelem * sizeof(double), c->stream));   // :215-217
    W2_CUDA(cudaMemsetAsync(c->fld[W2_F_VSS] + c->row_off, 0, c->nelem * sizeof(double), c->stream));
    W2_CUDA(cudaMemsetAsync(c->fld[W2_F_TSS] + c->row_off, 0, c->nelem * sizeof(double), c->stream));
    a->seeded = 1;
    return W2_OK;
}

// SmallScale(initflg, ...) on u1, v1, t1 -> uss, vss, pss, tss (the context's W2_F_USS.. fields)
int w2_smallscale(wolfd2_ctx *c, int initflg, const double *u1, const double *v1, const double *t1) {
    W2Atd *a = c->atd;
    if (!a) { w2_set_error("SmallScale: wolfd2_b200_set_smallscale has not been called"); return W2_ERR_BAD_ARG; }
    const wolfd2_smallscale &S = a->ss;
    double *uss = c->fld[W2_F_USS], *vss = c->fld[W2_F_VSS], *pss = c->fld[W2_F_PSS], *tss = c->fld[W2_F_TSS];
    if (initflg <= 0) {
        W2_TRY(w2_smallscale_seed(c));
        if (initflg < 0) return W2_OK;   // :243
    } else if (!a->seeded) {
        w2_set_error("SmallScale: initflg > 0 before the maps were initialised (initflg <= 0)");
        return W2_ERR_BAD_ARG;
    }
    const int nx = c->nx, ny = c->ny;
    const W2Metrics &M = c->met;
    dim3 g((nx + 255) / 256, ny < 2048 ? ny : 2048);
    ss_prepare_kernel<<<g, 256, 0, c->stream>>>(nx, ny, c->pitch, S.dlref, S.uref, S.tref, S.tmax, u1, v1, t1, M.xec, M.yec, M.xzc,
                                                M.yzc, a->ul, a->vl, a->tl, a->uf, a->vf, a->tf);
    c->launches[3]++;
    W2_TRY(w2_temp_bc(c, a->tf));                               // :284
    W2_TRY(w2_filter(c, W2_U, S.ssFiltPar[W2_U - 1], a->uf));   // :292-305
    W2_TRY(w2_filter(c, W2_V, S.ssFiltPar[W2_V - 1], a->vf));
    W2_TRY(w2_filter(c, W2_T, S.ssFiltPar[W2_T - 1], a->tf));
    ss_highpass_kernel<<<g, 256, 0, c->stream>>>(nx, ny, c->pitch, S.dlref, M.xec, M.yec, M.xzc, M.yzc, a->ul, a->vl, a->tl, a->uf,
                                                 a->vf, a->tf);
    c->launches[3]++;
    // cell-independent scalars with the host's libm, in the reference's order (:175-195, :366-393, :494-497)
    SsConst K;
    const double dOne = 1.0, dTwo = 2.0, dThree = 3.0, dFour = 4.0;
    const double rc = 0.20710678119e+00;
    K.dlref = S.dlref; K.uref = S.uref; K.tref = S.tref; K.tmax = S.tmax;
    K.rdl2 = S.dlref * S.dlref;
    const double piosr2 = acos(-dOne) / sqrt(dTwo);
    const double hs = S.dlref * S.ssHsCoef;
    K.dk = c->par.dk * S.dlref / S.uref;
    K.pr = S.pe / c->par.re;
    K.rnu = S.uref * S.dlref / c->par.re;
    K.sq_kappa = sqrt(K.rnu / K.pr);
    K.pehmin = 3.0;
    K.cu0 = S.ssCu0; K.TemCoef = S.ssTemCoef;
    K.iterate = S.ssCu0 > (double)1.e-10f;
    K.ts_num = S.ssTsCoef * piosr2;
    K.sqrt15 = sqrt(15.0); K.hs2 = hs * hs;
    K.bnumc = S.ssBnCrit; K.rlc = S.ssRMpExp; K.rmax = S.ssRMpMax; K.atanh_rc = atanh(rc / S.ssRMpMax);
    K.sqrt2 = sqrt(dTwo);
    K.cuT4_3 = dThree * pow(S.ssCu0, dFour);
    K.tArea = a->tArea; K.dtspan = S.tmax - S.tref;
    SsFields F;
    F.xec = M.xec; F.yec = M.yec; F.xzc = M.xzc; F.yzc = M.yzc; F.djc = M.djc;
    F.ul = a->ul; F.vl = a->vl; F.tl = a->tl; F.uf = a->uf; F.vf = a->vf; F.tf = a->tf;
    for (int k = 0; k < 9; ++k) F.map[k] = a->map[k];
    F.uss = uss; F.vss = vss; F.tss = tss;
    const W2Regions &R = c->hreg;
    for (int q = 0; q < R.nreg; ++q) {   // :319-333
        if (R.type[q] == W2_RM_BLOCKG || c->hth.ttype[q] == W2_RT_TEMPER) continue;
        const int ilo = R.iW[q] + 1, ihi = R.iE[q], jlo = R.jS[q] + 1, jhi = R.jN[q];
        if (ihi < ilo || jhi < jlo) continue;
        dim3 gc((ihi - ilo + 1 + 127) / 128, jhi - jlo + 1 < 4096 ? jhi - jlo + 1 : 4096);
        ss_cell_kernel<<<gc, 128, 0, c->stream>>>(c->pitch, ilo, ihi, jlo, jhi, K, F);
        c->launches[3]++;
    }
    W2_CUDA(cudaGetLastError());
    W2_TRY(w2_smlscl_bc(c, uss, vss, pss, tss));                // :527
    dim3 gz((nx + 1 + 255) / 256, ny + 1 < 2048 ? ny + 1 : 2048);
    ss_zero_p_kernel<<<gz, 256, 0, c->stream>>>(nx, ny, c->pitch, pss);   // :538-542
    c->launches[3]++;
    W2_TRY(w2_smlscl_bc(c, uss, vss, pss, tss));                // :545
    const wolfd2_params keep = c->par;                          // the model's own solver settings (:552-560)
    c->par.nPpeSolver = S.nssPpeSlvr; c->par.msorit = S.mssSorIt; c->par.sortol = S.ssSorTol; c->par.sorrel = S.ssSorRel;
    int nconv = 0, conv = 0;
    const int rc_ppe = w2_ppe(c, uss, vss, pss, &nconv, &conv);
    c->par = keep;
    W2_TRY(rc_ppe);
    a->nSorConv = nconv;
    W2_TRY(w2_smlscl_bc(c, uss, vss, pss, tss));                // :563
    W2_TRY(w2_project(c, pss, uss, vss));                       // :571
    return W2_OK;
}

extern "C" int wolfd2_b200_set_smallscale(wolfd2_ctx *c, const wolfd2_smallscale *ss) {
    if (!c || !ss) return W2_ERR_BAD_ARG;
    W2_CUDA(cudaSetDevice(c->device));
    if (ss->nsmallscl != 1) { if (c->atd) c->atd->ss.nsmallscl = 0; return W2_OK; }
    if (c->world > 1) { w2_set_error("the ATD small-scale model is not supported on several GPUs"); return W2_ERR_UNSUPPORTED; }
    if (!c->th_tables) {
        w2_set_error("set_smallscale: give the thermal region tables first (wolfd2_b200_set_thermal; nthermen may be 0)");
        return W2_ERR_BAD_ARG;
    }
    if (ss->nssPpeSlvr < 1 || ss->nssPpeSlvr > 6) { w2_set_error("Wrong nPpeSolver flag passed to Ppe: %d", ss->nssPpeSlvr); return W2_ERR_BAD_ARG; }
    if (!(ss->dlref > 0.0) || !(ss->uref > 0.0) || !(ss->pe > 0.0) || !(ss->ssRMpMax > 0.0)) {
        w2_set_error("set_smallscale: dlref, uref, pe and ssRMpMax must be positive");
        return W2_ERR_BAD_ARG;
    }
    W2_TRY(atd_alloc(c));
    c->atd->ss = *ss;
    return W2_OK;
}

// src/main.f:643-665: SmallScale(initflg = 0) on the current u, v, t before the time loop
extern "C" int wolfd2_b200_smallscale_init(wolfd2_ctx *c) {
    if (!c) return W2_ERR_BAD_ARG;
    W2_CUDA(cudaSetDevice(c->device));
    W2_TRY(w2_smallscale(c, 0, c->fld[W2_F_U], c->fld[W2_F_V], c->fld[W2_F_T]));
    W2_CUDA(cudaStreamSynchronize(c->stream));
    return W2_OK;
}
// test / restart hook: the saved map iterates, family 0..2 (u, v, t), plane 1..3, host layout (0:mnx,0:mny)
extern "C" int wolfd2_b200_smallscale_map(wolfd2_ctx *c, int32_t family, int32_t plane, double *host, int32_t upload) {
    if (!c || !c->atd || !host || family < 0 || family > 2 || plane < 1 || plane > 3) return W2_ERR_BAD_ARG;
    W2_CUDA(cudaSetDevice(c->device));
    double *d = c->atd->map[3 * family + plane - 1];
    if (upload) { W2_TRY(w2_upload2d(c, d, host)); c->atd->seeded = 1; }
    else W2_TRY(w2_download2d(c, host, d));
    W2_CUDA(cudaStreamSynchronize(c->stream));
    return W2_OK;
}
