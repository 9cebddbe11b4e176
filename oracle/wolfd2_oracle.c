/*
 * wolfd2_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT.
 *
 * A serial CPU restatement, in plain C, of the per-time-step hot path of
 * vgarzon/wolfd2 (Fortran 77).  It exists so that the CUDA path can be checked for
 * parity.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (wolfd2_b200/) never does.
 *
 * PARITY UNPINNED: the reference ships no golden vectors, fixtures, sample decks or
 * tests, and no Fortran compiler exists in this image or on the GPU
 * box, so this restatement could not be checked against output of the
 * reference itself.  Mitigation: every routine is transcribed statement by statement
 * with the same 1-based index expressions, the same operation order, the same
 * (0:mnx,0:mny) pitch and zero-initialised static storage, and is compiled with
 * -O2 -ffp-contract=off (gfortran -O2 on baseline x86-64 emits no FMA).  A second
 * restatement, written independently in numpy / plain Python from the same Fortran
 * text (tests/test_oracle_numpy_crosscheck.py), reproduces every routine of the
 * cold-flow path, the thermal row, SmallScale and Traject, whole time steps included,
 * bit for bit;
 * physical pins (Ghia's cavity at two Reynolds numbers, plane Poiseuille flow, de Vahl
 * Davis' natural convection, analytic conduction) check the physics
 * (tests/test_oracle_cpu.py).
 *
 * Each function cites the reference file:line it follows (paths relative to the
 * reference tree).  Fortran `stop` becomes: set orc_errflag and return.
 *
 * Calling convention: gfortran-style (all arguments by pointer, names lower-case +
 * trailing underscore) with an `orc_` prefix, so that the same Python signature table
 * binds the oracle and the product's literal shims.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/wolfd2_b200.h"

/* ---- include/config.f:19-26: compile-time in the reference, run-time here -------- */
static int MNX = 302, MNY = 302, MGRI = 20, MGRJ = 10;
static size_t LD = 303;        /* pitch of (0:mnx,0:mny) arrays                      */
static size_t NFULL = 303 * 303;
static int MN = 302 * 302;     /* mn = mnx*mny  (momentum.f:217)                     */

int orc_errflag = 0;

/* f(i,j) of a REAL f(0:mnx,0:mny) array */
#ifdef ORC_BOUNDS /* debug build: every subscript checked against (0:mnx,0:mny), like -fcheck=all */
static size_t orc_chk(long i, long j, int line) {
    if (i < 0 || i > MNX || j < 0 || j > MNY) { fprintf(stderr, "oracle: subscript (%ld,%ld) out of bounds at line %d\n", i, j, line); abort(); }
    return (size_t)i + LD * (size_t)j;
}
#define A(f, i, j) (f)[orc_chk((i), (j), __LINE__)]
#else
#define A(f, i, j) (f)[(size_t)(i) + LD * (size_t)(j)]
#endif
/* region tables, Fortran (mgri,mgrj,*) layout */
#define RB(ir, jr, k) nRegBrd[((ir) - 1) + MGRI * (((jr) - 1) + MGRJ * ((k) - 1))]
#define RT(ir, jr) nRegType[((ir) - 1) + MGRI * ((jr) - 1)]
#define MB(ir, jr, k) nMomBdTp[((ir) - 1) + MGRI * (((jr) - 1) + MGRJ * ((k) - 1))]
#define BV(ir, jr, k, l) \
    dBCVal[((ir) - 1) + MGRI * (((jr) - 1) + MGRJ * (((k) - 1) + 4 * ((l) - 1)))]
#define PR(a, ir, jr) (a)[((ir) - 1) + MGRI * ((jr) - 1)]

enum { WEST = 1, EAST = 2, SOUTH = 3, NORTH = 4 };
enum { _U_ = 1, _V_ = 2, _P_ = 3, _T_ = 4 };
enum { _I_ = 0, _J_ = 1 }; /* 0-based positions of nReg(_I_), nReg(_J_) */
enum { RM_BLOCKG = 0, RM_INTERN = 1, RM_POROUS = 2 };
enum { BM_INTERN = 0, BM_WALL1 = 1, BM_WALL2 = 2, BM_INLET = 3, BM_OUTLT1 = 4, BM_OUTLT2 = 5 };

/* ---- static (zero-initialised) local arrays of the reference subroutines ---------- */
typedef struct {
    double *cj1, *cj2, *c1s, *c2s, *c1n, *c2n, *cps, *cpn, *cpj, *cnvs, *difs, *cnvn, *difn;
    double *a, *b; /* a(3,mn), b(mn) */
} mom_work_t;
static mom_work_t WX, WY;              /* XMomentum / YMomentum locals (momentum.f:258-266, 581-589) */
static double *W_dus, *W_dvs;          /* nAuxMomentum locals (momentum.f:101) */
static double *W_div, *W_pa, *W_pb;    /* Ppe locals div, a(mn,5), b(mn) (pressure.f:75-76) */
static double *W_qh;                   /* Filter local (utility.f:65) */
static double *W_pn;                   /* SlorRBP local pn (pressure.f:1008) */
static double *W_aline, *W_bline, *W_plold; /* Slor* locals (pressure.f:700) */
static double *W_dif;                  /* SorRBP automatic array dif (pressure.f:574) */
static double *W_cu1, *W_cun, *W_cv1, *W_cvn, *W_s, *W_ta, *W_tb; /* ThermEnergy locals (thermal.f:87-91) */

static double *zalloc(size_t n) {
    double *p = (double *)calloc(n ? n : 1, sizeof(double));
    if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    return p;
}
static void mom_work_alloc(mom_work_t *w) {
    double **f = &w->cj1;
    for (int k = 0; k < 13; ++k) { free(f[k]); f[k] = zalloc(NFULL); }
    free(w->a); free(w->b);
    w->a = zalloc(3 * (size_t)MN);
    w->b = zalloc((size_t)MN);
}

void orc_config(int32_t mnx, int32_t mny, int32_t mgri, int32_t mgrj) {
    MNX = mnx; MNY = mny; MGRI = mgri; MGRJ = mgrj;
    LD = (size_t)mnx + 1;
    NFULL = LD * ((size_t)mny + 1);
    MN = mnx * mny;
    mom_work_alloc(&WX);
    mom_work_alloc(&WY);
    free(W_dus); free(W_dvs); free(W_div); free(W_pa); free(W_pb); free(W_qh); free(W_pn);
    free(W_aline); free(W_bline); free(W_plold); free(W_dif);
    W_dus = zalloc(NFULL); W_dvs = zalloc(NFULL);
    W_div = zalloc(NFULL); W_pa = zalloc(5 * (size_t)MN); W_pb = zalloc((size_t)MN);
    W_qh = zalloc(NFULL); W_pn = zalloc(NFULL);
    size_t mnl = (size_t)(mnx > mny ? mnx : mny);
    W_aline = zalloc(3 * mnl); W_bline = zalloc(mnl); W_plold = zalloc(mnl);
    W_dif = zalloc(NFULL);
    free(W_cu1); free(W_cun); free(W_cv1); free(W_cvn); free(W_s); free(W_ta); free(W_tb);
    W_cu1 = zalloc(NFULL); W_cun = zalloc(NFULL); W_cv1 = zalloc(NFULL); W_cvn = zalloc(NFULL); W_s = zalloc(NFULL);
    W_ta = zalloc(3 * (size_t)MN); W_tb = zalloc((size_t)MN);
    orc_errflag = 0;
}
int32_t orc_get_errflag(void) { return orc_errflag; }

/* =============================== momentum.f ===================================== */

/* AltTridLU, src/momentum.f:1307-1339.  a(3,n) AoS, 1-based a(k,i) -> a[(k-1)+3*(i-1)] */
void orc_alttridlu_(const int32_t *n_, double *a, double *b) {
    const int n = *n_;
#define AA(k, i) a[((k) - 1) + 3 * ((size_t)(i) - 1)]
#define BB(i) b[(size_t)(i) - 1]
    AA(3, 1) = AA(3, 1) / AA(2, 2);          /* :1319  (sic: divides by row 2's diagonal) */
    BB(1) = BB(1) / AA(2, 1);                /* :1320 */
    for (int i = 2; i <= n - 1; ++i) {       /* :1322-1328 */
        AA(2, i) = AA(2, i) - (AA(1, i) * AA(3, i - 1));
        AA(3, i) = AA(3, i) / AA(2, i);
        BB(i) = (BB(i) - AA(1, i) * BB(i - 1)) / AA(2, i);
    }
    AA(2, n) = AA(2, n) - (AA(1, n) * AA(3, n - 1));     /* :1330 */
    BB(n) = (BB(n) - AA(1, n) * BB(n - 1)) / AA(2, n);   /* :1331 */
    for (int i = n - 1; i >= 1; --i)                     /* :1334-1336 */
        BB(i) = BB(i) - (AA(3, i) * BB(i + 1));
#undef AA
#undef BB
}

/* ConvCoef, src/momentum.f:864-981 */
void orc_convcoef_(const int32_t *nx_, const int32_t *ny_, const int32_t *ncomp_,
                   const int32_t *njacob_, const double *xzi, const double *xet,
                   const double *yzi, const double *yet, const double *u, const double *v,
                   double *cc1, double *cc2) {
    const int nx = *nx_, ny = *ny_, ncomp = *ncomp_, njacob = *njacob_;
    const double dOne = 1.0, dTwo = 2.0, dFour = 4.0, dHalf = 0.5;
    double djac = dOne;                 /* :895 */
    if (njacob == 1) djac = dTwo;       /* :896 */
    int i, j;
    switch (ncomp) {
    case 1: /* :901-913 */
        for (j = 1; j <= ny + 1; ++j)
            for (i = 1; i <= nx + 1; ++i)
                A(cc1, i, j) = (djac * A(yet, i, j) * (A(u, i, j) + A(u, i - 1, j))
                                - A(xet, i, j) * (A(v, i, j) + A(v, i, j - 1))) * dHalf;
        for (j = 1; j <= ny; ++j)
            for (i = 1; i <= nx; ++i)
                A(cc2, i, j) = (A(xzi, i, j) * (A(v, i + 1, j) + A(v, i, j))
                                - djac * A(yzi, i, j) * (A(u, i, j + 1) + A(u, i, j))) * dHalf;
        break;
    case 2: /* :916-928 */
        for (j = 1; j <= ny; ++j)
            for (i = 1; i <= nx; ++i)
                A(cc1, i, j) = (A(yet, i, j) * (A(u, i, j + 1) + A(u, i, j))
                                - djac * A(xet, i, j) * (A(v, i + 1, j) + A(v, i, j))) * dHalf;
        for (j = 1; j <= ny + 1; ++j)
            for (i = 1; i <= nx + 1; ++i)
                A(cc2, i, j) = (djac * A(xzi, i, j) * (A(v, i, j) + A(v, i, j - 1))
                                - A(yzi, i, j) * (A(u, i, j) + A(u, i - 1, j))) * dHalf;
        break;
    case 3: /* :931-939 */
        for (j = 1; j <= ny; ++j)
            for (i = 1; i <= nx; ++i) {
                A(cc1, i, j) = (A(yet, i, j) * A(u, i, j)
                                - A(xet, i, j) * (A(v, i + 1, j) + A(v, i, j)
                                                  + A(v, i + 1, j - 1) + A(v, i, j - 1)) / dFour) * dHalf;
                A(cc2, i, j) = (A(xzi, i, j) * A(v, i, j)
                                - A(yzi, i, j) * (A(u, i, j + 1) + A(u, i - 1, j + 1)
                                                  + A(u, i, j) + A(u, i - 1, j)) / dFour) * dHalf;
            }
        break;
    case 4: /* :942-950 */
        for (j = 1; j <= ny; ++j)
            for (i = 0; i <= nx; ++i) {
                A(cc1, i, j) = djac * A(yet, i, j) * A(u, i, j)
                               - A(xet, i, j) * (A(v, i + 1, j) + A(v, i, j)
                                                 + A(v, i + 1, j - 1) + A(v, i, j - 1)) / dFour;
                A(cc2, i, j) = A(xzi, i, j) * (A(v, i + 1, j) + A(v, i, j) + A(v, i + 1, j - 1)
                                               + A(v, i, j - 1)) / dFour
                               - djac * A(yzi, i, j) * A(u, i, j);
            }
        break;
    case 5: /* :953-961 */
        for (j = 0; j <= ny; ++j)
            for (i = 1; i <= nx; ++i) {
                A(cc1, i, j) = A(yet, i, j) * (A(u, i, j + 1) + A(u, i - 1, j + 1) + A(u, i, j)
                                               + A(u, i - 1, j)) / dFour
                               - djac * A(xet, i, j) * A(v, i, j);
                A(cc2, i, j) = djac * A(xzi, i, j) * A(v, i, j)
                               - A(yzi, i, j) * (A(u, i, j + 1) + A(u, i - 1, j + 1) + A(u, i, j)
                                                 + A(u, i - 1, j)) / dFour;
            }
        break;
    case 6: /* :964-972 */
        for (j = 1; j <= ny + 1; ++j)
            for (i = 1; i <= nx + 1; ++i) {
                A(cc1, i, j) = (A(yet, i, j) * (A(u, i, j) + A(u, i - 1, j))
                                - A(xet, i, j) * (A(v, i, j) + A(v, i, j - 1))) * dHalf;
                A(cc2, i, j) = (A(xzi, i, j) * (A(v, i, j) + A(v, i, j - 1))
                                - A(yzi, i, j) * (A(u, i, j) + A(u, i - 1, j))) * dHalf;
            }
        break;
    default:
        fprintf(stderr, "Error: Wrong component flag passed to ConvCoef: %d\n", ncomp);
        orc_errflag = 1;
    }
}

/* DConvU, src/momentum.f:987-1009 */
void orc_dconvu_(const int32_t *nx_, const int32_t *ny_, const double *c1, const double *c2,
                 const double *u, double *c) {
    const int nx = *nx_, ny = *ny_;
    for (int j = 2; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i)
            A(c, i, j) = -A(c2, i, j - 1) * A(u, i, j - 1) - A(c1, i, j) * A(u, i - 1, j)
                         + (A(c1, i + 1, j) - A(c1, i, j) + A(c2, i, j) - A(c2, i, j - 1)) * A(u, i, j)
                         + A(c1, i + 1, j) * A(u, i + 1, j) + A(c2, i, j) * A(u, i, j + 1);
}

/* DDiffU, src/momentum.f:1015-1045 */
void orc_ddiffu_(const int32_t *nx_, const int32_t *ny_, const double *ac, const double *bc,
                 const double *bn, const double *gn, const double *u, double *d) {
    const int nx = *nx_, ny = *ny_;
    for (int j = 2; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) {
            double s1 = A(ac, i + 1, j) * (A(u, i + 1, j) - A(u, i, j))
                        - A(ac, i, j) * (A(u, i, j) - A(u, i - 1, j))
                        + A(bc, i + 1, j) * (A(u, i + 1, j + 1) + A(u, i, j + 1) - A(u, i + 1, j - 1) - A(u, i, j - 1))
                        - A(bc, i, j) * (A(u, i, j + 1) + A(u, i - 1, j + 1) - A(u, i, j - 1) - A(u, i - 1, j - 1));
            double s2 = A(bn, i, j) * (A(u, i + 1, j + 1) + A(u, i + 1, j) - A(u, i - 1, j + 1) - A(u, i - 1, j))
                        - A(bn, i, j - 1) * (A(u, i + 1, j) + A(u, i + 1, j - 1) - A(u, i - 1, j) - A(u, i - 1, j - 1))
                        + A(gn, i, j) * (A(u, i, j + 1) - A(u, i, j))
                        - A(gn, i, j - 1) * (A(u, i, j) - A(u, i, j - 1));
            A(d, i, j) = s1 + s2;
        }
}

/* DConvV, src/momentum.f:1051-1073 */
void orc_dconvv_(const int32_t *nx_, const int32_t *ny_, const double *c1, const double *c2,
                 const double *v, double *c) {
    const int nx = *nx_, ny = *ny_;
    for (int j = 1; j <= ny; ++j)
        for (int i = 2; i <= nx; ++i)
            A(c, i, j) = -A(c2, i, j) * A(v, i, j - 1) - A(c1, i - 1, j) * A(v, i - 1, j)
                         + (A(c1, i, j) - A(c1, i - 1, j) + A(c2, i, j + 1) - A(c2, i, j)) * A(v, i, j)
                         + A(c1, i, j) * A(v, i + 1, j) + A(c2, i, j + 1) * A(v, i, j + 1);
}

/* DDiffV, src/momentum.f:1079-1109 */
void orc_ddiffv_(const int32_t *nx_, const int32_t *ny_, const double *an, const double *bc,
                 const double *bn, const double *gc, const double *v, double *d) {
    const int nx = *nx_, ny = *ny_;
    for (int j = 1; j <= ny; ++j)
        for (int i = 2; i <= nx; ++i) {
            double s1 = A(an, i, j) * (A(v, i + 1, j) - A(v, i, j))
                        - A(an, i - 1, j) * (A(v, i, j) - A(v, i - 1, j))
                        + A(bn, i, j) * (A(v, i + 1, j + 1) + A(v, i, j + 1) - A(v, i + 1, j - 1) - A(v, i, j - 1))
                        - A(bn, i - 1, j) * (A(v, i, j + 1) + A(v, i - 1, j + 1) - A(v, i, j - 1) - A(v, i - 1, j - 1));
            double s2 = A(bc, i, j + 1) * (A(v, i + 1, j + 1) + A(v, i + 1, j) - A(v, i - 1, j + 1) - A(v, i - 1, j))
                        - A(bc, i, j) * (A(v, i + 1, j) + A(v, i + 1, j - 1) - A(v, i - 1, j) - A(v, i - 1, j - 1))
                        + A(gc, i, j + 1) * (A(v, i, j + 1) - A(v, i, j))
                        - A(gc, i, j) * (A(v, i, j) - A(v, i, j - 1));
            A(d, i, j) = s1 + s2;
        }
}

/* PorosCoef, src/momentum.f:1115-1226 */
void orc_poroscoef_(const int32_t *nx_, const int32_t *ny_, const int32_t *ncomp_,
                    const int32_t *njacob_, const int32_t *nReg, const int32_t *nRegBrd,
                    const int32_t *nRegType, const double *dPRporos, const double *dPRporc1,
                    const double *dPRporc2, const double *u, const double *v, double *cp) {
    (void)nx_; (void)ny_; (void)dPRporos;
    const int ncomp = *ncomp_, njacob = *njacob_;
    const double dZero = 0.0, dFour = 4.0;
    for (int jreg = 1; jreg <= nReg[_J_]; ++jreg)
        for (int ireg = 1; ireg <= nReg[_I_]; ++ireg) {
            int iW = RB(ireg, jreg, WEST), iE = RB(ireg, jreg, EAST);
            int jS = RB(ireg, jreg, SOUTH), jN = RB(ireg, jreg, NORTH);
            int i, j;
            if (RT(ireg, jreg) != RM_POROUS) { /* :1166-1177 */
                for (j = jS; j <= jN; ++j)
                    for (i = iW; i <= iE; ++i) A(cp, i, j) = dZero;
                continue;
            }
            double porc1 = PR(dPRporc1, ireg, jreg), porc2 = PR(dPRporc2, ireg, jreg);
            double unorm, unrm1;
            switch (ncomp) {
            case 1: /* :1188-1201; note the mis-parenthesised /dFour (sic) */
                for (j = jS + 1; j <= jN; ++j)
                    for (i = iW; i <= iE; ++i) {
                        double t = (A(v, i, j) + A(v, i + 1, j) + A(v, i, j - 1)
                                    + A(v, i + 1, j - 1) / dFour);
                        unorm = sqrt(A(u, i, j) * A(u, i, j) + t * t);
                        unrm1 = dZero;
                        if (unorm > 1.e-8) unrm1 = (A(u, i, j) * A(u, i, j)) / unorm;
                        if (njacob == 1) unorm = unrm1 + unorm;
                        A(cp, i, j) = porc1 + porc2 * unorm;
                    }
                break;
            case 2: /* :1203-1214 */
                for (j = jS; j <= jN; ++j)
                    for (i = iW + 1; i <= iE; ++i) {
                        double t = ((A(u, i - 1, j + 1) + A(u, i, j + 1) + A(u, i - 1, j)
                                     + A(u, i, j)) / dFour);
                        unorm = sqrt(t * t + A(v, i, j) * A(v, i, j));
                        unrm1 = dZero;
                        if (unorm > 1.e-8) unrm1 = (A(v, i, j) * A(v, i, j)) / unorm;
                        if (njacob == 1) unorm = unrm1 + unorm;
                        A(cp, i, j) = porc1 + porc2 * unorm;
                    }
                break;
            default:
                fprintf(stderr, "Error: Wrong ncomp flag passed to PorosCoef: %d\n", ncomp);
                orc_errflag = 1;
                return;
            }
        }
}

/* identity row helper used by the blockage / wall loops of X/YMomentum */
#define IDROW(a, b, ind) do { (a)[0 + 3 * ((size_t)(ind) - 1)] = 0.0; \
    (a)[1 + 3 * ((size_t)(ind) - 1)] = 1.0; (a)[2 + 3 * ((size_t)(ind) - 1)] = 0.0; \
    (b)[(size_t)(ind) - 1] = 0.0; } while (0)

/* XMomentum, src/momentum.f:199-514 */
void orc_xmomentum_(const int32_t *nx_, const int32_t *ny_,
    const int32_t *nReg, const int32_t *nRegBrd, const int32_t *nRegType,
    const int32_t *nMomBdTp, const double *dk_, const double *re_,
    const double *dPRporos, const double *dPRporc1, const double *dPRporc2,
    const double *rbn, const double *rgn, const double *rac, const double *rbc,
    const double *dju,
    const double *xec, const double *yec, const double *xzn, const double *yzn,
    const double *xeu, const double *yeu, const double *xzu, const double *yzu,
    const double *us, const double *vs, const double *un, const double *vn, double *dus) {
    const int nx = *nx_, ny = *ny_;
    const double dk = *dk_, re = *re_;
    const double dZero = 0.0, dOne = 1.0, dHalf = 0.5;
    mom_work_t *w = &WX;
    double *cj1 = w->cj1, *cj2 = w->cj2, *c1s = w->c1s, *c2s = w->c2s, *c1n = w->c1n, *c2n = w->c2n;
    double *cps = w->cps, *cpn = w->cpn, *cpj = w->cpj;
    double *cnvs = w->cnvs, *difs = w->difs, *cnvn = w->cnvn, *difn = w->difn;
    double *a = w->a, *b = w->b;
#define AA(k, ind) a[((k) - 1) + 3 * ((size_t)(ind) - 1)]
    int i, j, ind, ireg, jreg;
    const int32_t c4 = 4, c1 = 1, c0 = 0;

    const double re1 = dOne / re;   /* :276 */
    const double dk2 = dk * dHalf;  /* :277 */

    orc_convcoef_(nx_, ny_, &c4, &c1, xzu, xeu, yzu, yeu, us, vs, cj1, cj2);  /* :282 */
    orc_convcoef_(nx_, ny_, &c1, &c0, xzn, xec, yzn, yec, us, vs, c1s, c2s);  /* :285 */
    orc_convcoef_(nx_, ny_, &c1, &c0, xzn, xec, yzn, yec, un, vn, c1n, c2n);  /* :288 */

    /* porous regions: divide convective terms by porosity, :296-324 */
    for (jreg = 1; jreg <= nReg[_J_]; ++jreg)
        for (ireg = 1; ireg <= nReg[_I_]; ++ireg) {
            int iW = RB(ireg, jreg, WEST), iE = RB(ireg, jreg, EAST);
            int jS = RB(ireg, jreg, SOUTH), jN = RB(ireg, jreg, NORTH);
            if (RT(ireg, jreg) == RM_POROUS) {
                double dLocPoros = PR(dPRporos, ireg, jreg);
                for (j = jS + 1; j <= jN; ++j)
                    for (i = iW; i <= iE; ++i) {
                        A(cj1, i, j) = A(cj1, i, j) / dLocPoros;
                        A(cj2, i, j) = A(cj2, i, j) / dLocPoros;
                        A(c1s, i, j) = A(c1s, i, j) / dLocPoros;
                        A(c2s, i, j) = A(c2s, i, j) / dLocPoros;
                        A(c1n, i, j) = A(c1n, i, j) / dLocPoros;
                        A(c2n, i, j) = A(c2n, i, j) / dLocPoros;
                    }
            }
        }

    orc_dconvu_(nx_, ny_, c1s, c2s, us, cnvs);               /* :327 */
    orc_dconvu_(nx_, ny_, c1n, c2n, un, cnvn);               /* :328 */
    orc_ddiffu_(nx_, ny_, rac, rbc, rbn, rgn, us, difs);     /* :329 */
    orc_ddiffu_(nx_, ny_, rac, rbc, rbn, rgn, un, difn);     /* :330 */

    orc_poroscoef_(nx_, ny_, &c1, &c1, nReg, nRegBrd, nRegType, dPRporos, dPRporc1, dPRporc2, us, vs, cpj);
    orc_poroscoef_(nx_, ny_, &c1, &c0, nReg, nRegBrd, nRegType, dPRporos, dPRporc1, dPRporc2, us, vs, cps);
    orc_poroscoef_(nx_, ny_, &c1, &c0, nReg, nRegBrd, nRegType, dPRporos, dPRporc1, dPRporc2, un, vn, cpn);

    /* first split step, :350-384 */
    for (j = 2; j <= ny; ++j)
        for (i = 1; i <= nx; ++i) {
            ind = (j - 2) * nx + i;
            double rkj = dk2 * A(dju, i, j);
            if (A(cj1, i, j) >= dZero) {
                AA(1, ind) = rkj * (-A(cj1, i - 1, j) - re1 * A(rac, i, j));
                AA(2, ind) = dOne + rkj * (A(cj1, i, j)
                             + re1 * (A(rac, i + 1, j) + A(rac, i, j))) + dk2 * A(cpj, i, j);
                AA(3, ind) = rkj * (-re1 * A(rac, i + 1, j));
            } else {
                AA(1, ind) = rkj * (-re1 * A(rac, i, j));
                AA(2, ind) = dOne + rkj * (-A(cj1, i, j)
                             + re1 * (A(rac, i + 1, j) + A(rac, i, j))) + dk2 * A(cpj, i, j);
                AA(3, ind) = rkj * (A(cj1, i + 1, j) - re1 * A(rac, i + 1, j));
            }
            b[ind - 1] = A(un, i, j) - A(us, i, j) + rkj * (-A(cnvs, i, j) - A(cnvn, i, j))
                         + rkj * re1 * (A(difs, i, j) + A(difn, i, j))
                         - dk2 * (A(cps, i, j) * A(us, i, j) + A(cpn, i, j) * A(un, i, j));
        }
    { int32_t n = nx * (ny - 1); orc_alttridlu_(&n, a, b); }   /* :389 */

    /* second split step LHS, :396-428 */
    for (j = 2; j <= ny; ++j)
        for (i = 1; i <= nx; ++i) {
            ind = (j - 2) * nx + i;
            double rkj = dk2 * A(dju, i, j);
            if (A(cj2, i, j) >= dZero) {
                AA(1, ind) = rkj * (-A(cj2, i, j - 1) - re1 * A(rgn, i, j - 1));
                AA(2, ind) = dOne + rkj * (A(cj2, i, j)
                             + re1 * (A(rgn, i, j) + A(rgn, i, j - 1))) + dk2 * A(cpj, i, j);
                AA(3, ind) = rkj * (-re1 * A(rgn, i, j));
            } else {
                AA(1, ind) = rkj * (-re1 * A(rgn, i, j - 1));
                AA(2, ind) = dOne + rkj * (-A(cj2, i, j)
                             + re1 * (A(rgn, i, j) + A(rgn, i, j - 1))) + dk2 * A(cpj, i, j);
                AA(3, ind) = rkj * (A(cj2, i, j + 1) - re1 * A(rgn, i, j));
            }
        }

    /* blockage and wall/inlet rows, :434-496 */
    for (jreg = 1; jreg <= nReg[_J_]; ++jreg)
        for (ireg = 1; ireg <= nReg[_I_]; ++ireg) {
            int iW = RB(ireg, jreg, WEST), iE = RB(ireg, jreg, EAST);
            int jS = RB(ireg, jreg, SOUTH), jN = RB(ireg, jreg, NORTH);
            if (RT(ireg, jreg) == RM_BLOCKG)
                for (j = jS + 1; j <= jN; ++j)
                    for (i = iW; i <= iE; ++i) { ind = (j - 2) * nx + i; IDROW(a, b, ind); }
            switch (MB(ireg, jreg, WEST)) {
            case BM_INTERN: case BM_OUTLT1: case BM_OUTLT2: break;
            case BM_WALL1: case BM_WALL2: case BM_INLET:
                for (j = jS + 1; j <= jN; ++j) { ind = (j - 2) * nx + iW; IDROW(a, b, ind); }
                break;
            default:
                fprintf(stderr, "Wrong nBdTypeW flag in region %d,%d\n", ireg, jreg);
                orc_errflag = 1; return;
            }
            switch (MB(ireg, jreg, EAST)) {
            case BM_INTERN: case BM_OUTLT1: case BM_OUTLT2: break;
            case BM_WALL1: case BM_WALL2: case BM_INLET:
                for (j = jS + 1; j <= jN; ++j) { ind = (j - 2) * nx + iE; IDROW(a, b, ind); }
                break;
            default:
                fprintf(stderr, "Wrong nBdTypeE flag in region %d,%d\n", ireg, jreg);
                orc_errflag = 1; return;
            }
        }
    { int32_t n = nx * (ny - 1); orc_alttridlu_(&n, a, b); }   /* :501 */

    for (j = 2; j <= ny; ++j)                                   /* :505-510 */
        for (i = 1; i <= nx; ++i) { ind = (j - 2) * nx + i; A(dus, i, j) = b[ind - 1]; }
#undef AA
}

/* YMomentum, src/momentum.f:520-838 */
void orc_ymomentum_(const int32_t *nx_, const int32_t *ny_,
    const int32_t *nReg, const int32_t *nRegBrd, const int32_t *nRegType,
    const int32_t *nMomBdTp, const double *dk_, const double *re_, const double *fr_,
    const double *dPRporos, const double *dPRporc1, const double *dPRporc2,
    const double *ran, const double *rbn, const double *rbc, const double *rgc,
    const double *djv,
    const double *xen, const double *yen, const double *xzc, const double *yzc,
    const double *xev, const double *yev, const double *xzv, const double *yzv,
    const double *d, const double *dn,
    const double *us, const double *vs, const double *un, const double *vn, double *dvs) {
    const int nx = *nx_, ny = *ny_;
    const double dk = *dk_, re = *re_, fr = *fr_;
    const double dZero = 0.0, dOne = 1.0, dFour = 4.0, dHalf = 0.5;
    mom_work_t *w = &WY;
    double *cj1 = w->cj1, *cj2 = w->cj2, *c1s = w->c1s, *c2s = w->c2s, *c1n = w->c1n, *c2n = w->c2n;
    double *cps = w->cps, *cpn = w->cpn, *cpj = w->cpj;
    double *cnvs = w->cnvs, *difs = w->difs, *cnvn = w->cnvn, *difn = w->difn;
    double *a = w->a, *b = w->b;
#define AA(k, ind) a[((k) - 1) + 3 * ((size_t)(ind) - 1)]
    int i, j, ind, ireg, jreg;
    const int32_t c5 = 5, c2 = 2, c1 = 1, c0 = 0;

    const double re1 = dOne / re;   /* :600 */
    const double dk2 = dk * dHalf;  /* :601 */

    orc_convcoef_(nx_, ny_, &c5, &c1, xzv, xev, yzv, yev, us, vs, cj1, cj2);  /* :607 */
    orc_convcoef_(nx_, ny_, &c2, &c0, xzc, xen, yzc, yen, us, vs, c1s, c2s);  /* :610 */
    orc_convcoef_(nx_, ny_, &c2, &c0, xzc, xen, yzc, yen, un, vn, c1n, c2n);  /* :613 */

    for (jreg = 1; jreg <= nReg[_J_]; ++jreg)                                  /* :621-649 */
        for (ireg = 1; ireg <= nReg[_I_]; ++ireg) {
            int iW = RB(ireg, jreg, WEST), iE = RB(ireg, jreg, EAST);
            int jS = RB(ireg, jreg, SOUTH), jN = RB(ireg, jreg, NORTH);
            if (RT(ireg, jreg) == RM_POROUS) {
                double dLocPoros = PR(dPRporos, ireg, jreg);
                for (j = jS; j <= jN; ++j)
                    for (i = iW + 1; i <= iE; ++i) {
                        A(cj1, i, j) = A(cj1, i, j) / dLocPoros;
                        A(cj2, i, j) = A(cj2, i, j) / dLocPoros;
                        A(c1s, i, j) = A(c1s, i, j) / dLocPoros;
                        A(c2s, i, j) = A(c2s, i, j) / dLocPoros;
                        A(c1n, i, j) = A(c1n, i, j) / dLocPoros;
                        A(c2n, i, j) = A(c2n, i, j) / dLocPoros;
                    }
            }
        }

    orc_dconvv_(nx_, ny_, c1s, c2s, vs, cnvs);               /* :652 */
    orc_dconvv_(nx_, ny_, c1n, c2n, vn, cnvn);               /* :653 */
    orc_ddiffv_(nx_, ny_, ran, rbc, rbn, rgc, vs, difs);     /* :654 */
    orc_ddiffv_(nx_, ny_, ran, rbc, rbn, rgc, vn, difn);     /* :655 */

    orc_poroscoef_(nx_, ny_, &c2, &c1, nReg, nRegBrd, nRegType, dPRporos, dPRporc1, dPRporc2, us, vs, cpj);
    orc_poroscoef_(nx_, ny_, &c2, &c0, nReg, nRegBrd, nRegType, dPRporos, dPRporc1, dPRporc2, us, vs, cps);
    orc_poroscoef_(nx_, ny_, &c2, &c0, nReg, nRegBrd, nRegType, dPRporos, dPRporc1, dPRporc2, un, vn, cpn);

    /* first split step, :675-711 */
    for (j = 1; j <= ny; ++j)
        for (i = 2; i <= nx; ++i) {
            ind = (j - 1) * (nx - 1) + i - 1;
            double rkj = dk2 * A(djv, i, j);
            if (A(cj1, i, j) >= dZero) {
                AA(1, ind) = rkj * (-A(cj1, i - 1, j) - re1 * A(ran, i - 1, j));
                AA(2, ind) = rkj * (A(cj1, i, j)
                             + re1 * (A(ran, i, j) + A(ran, i - 1, j))) + dk2 * A(cpj, i, j) + dOne;
                AA(3, ind) = rkj * (-re1 * A(ran, i, j));
            } else {
                AA(1, ind) = rkj * (-re1 * A(ran, i - 1, j));
                AA(2, ind) = rkj * (-A(cj1, i, j)
                             + re1 * (A(ran, i, j) + A(ran, i - 1, j))) + dk2 * A(cpj, i, j) + dOne;
                AA(3, ind) = rkj * (A(cj1, i + 1, j) - re1 * A(ran, i, j));
            }
            double buoy = dk * (A(d, i, j + 1) + A(d, i, j) + A(dn, i, j + 1) + A(dn, i, j)) / (dFour * fr);
            b[ind - 1] = A(vn, i, j) - A(vs, i, j) + rkj * (-A(cnvs, i, j) - A(cnvn, i, j))
                         + rkj * re1 * (A(difs, i, j) + A(difn, i, j))
                         - dk2 * (A(cps, i, j) * A(vs, i, j) + A(cpn, i, j) * A(vn, i, j))
                         - buoy;
        }
    { int32_t n = (nx - 1) * ny; orc_alttridlu_(&n, a, b); }   /* :716 */

    /* second split step LHS, :723-754 */
    for (j = 1; j <= ny; ++j)
        for (i = 2; i <= nx; ++i) {
            ind = (j - 1) * (nx - 1) + i - 1;
            double rkj = dk2 * A(djv, i, j);
            if (A(cj2, i, j) >= dZero) {
                AA(1, ind) = rkj * (-A(cj2, i, j - 1) - re1 * A(rgc, i, j));
                AA(2, ind) = rkj * (A(cj2, i, j)
                             + re1 * (A(rgc, i, j + 1) + A(rgc, i, j))) + dk2 * A(cpj, i, j) + dOne;
                AA(3, ind) = rkj * (-re1 * A(rgc, i, j + 1));
            } else {
                AA(1, ind) = rkj * (-re1 * A(rgc, i, j));
                AA(2, ind) = rkj * (-A(cj2, i, j)
                             + re1 * (A(rgc, i, j + 1) + A(rgc, i, j))) + dk2 * A(cpj, i, j) + dOne;
                AA(3, ind) = rkj * (A(cj2, i, j + 1) - re1 * A(rgc, i, j + 1));
            }
        }

    /* blockage and wall/inlet rows, :760-821 */
    for (jreg = 1; jreg <= nReg[_J_]; ++jreg)
        for (ireg = 1; ireg <= nReg[_I_]; ++ireg) {
            int iW = RB(ireg, jreg, WEST), iE = RB(ireg, jreg, EAST);
            int jS = RB(ireg, jreg, SOUTH), jN = RB(ireg, jreg, NORTH);
            if (RT(ireg, jreg) == RM_BLOCKG)
                for (j = jS; j <= jN; ++j)
                    for (i = iW + 1; i <= iE; ++i) { ind = (j - 1) * (nx - 1) + i - 1; IDROW(a, b, ind); }
            switch (MB(ireg, jreg, SOUTH)) {
            case BM_INTERN: case BM_OUTLT1: case BM_OUTLT2: break;
            case BM_WALL1: case BM_WALL2: case BM_INLET:
                for (i = iW + 1; i <= iE; ++i) { ind = (jS - 1) * (nx - 1) + i - 1; IDROW(a, b, ind); }
                break;
            default:
                fprintf(stderr, "Wrong nBdTypeS flag in region %d,%d\n", ireg, jreg);
                orc_errflag = 1; return;
            }
            switch (MB(ireg, jreg, NORTH)) {
            case BM_INTERN: case BM_OUTLT1: case BM_OUTLT2: break;
            case BM_WALL1: case BM_WALL2: case BM_INLET:
                for (i = iW + 1; i <= iE; ++i) { ind = (jN - 1) * (nx - 1) + i - 1; IDROW(a, b, ind); }
                break;
            default:
                fprintf(stderr, "Wrong nBdTypeN flag in region %d,%d\n", ireg, jreg);
                orc_errflag = 1; return;
            }
        }
    { int32_t n = (nx - 1) * ny; orc_alttridlu_(&n, a, b); }   /* :826 */

    for (j = 1; j <= ny; ++j)                                   /* :829-834 */
        for (i = 2; i <= nx; ++i) { ind = (j - 1) * (nx - 1) + i - 1; A(dvs, i, j) = b[ind - 1]; }
#undef AA
}

/* forward declarations (other files of the reference) */
void orc_veloutflowbcs_(const int32_t *, const int32_t *, const int32_t *, const int32_t *,
                        const int32_t *, const double *, double *, double *);
double orc_dmaxnorm_(const int32_t *, const int32_t *, const double *);

/* Test hook: number of QL iterations executed and their max-norms (not in reference). */
double orc_last_ql_dif[2];

/* nAuxMomentum, src/momentum.f:33-193 */
int32_t orc_nauxmomentum_(const int32_t *nx_, const int32_t *ny_, const int32_t *mqiter_,
    const int32_t *nReg, const int32_t *nRegBrd, const int32_t *nRegType,
    const int32_t *nMomBdTp,
    const double *dk, const double *re, const double *fr, const double *qtol_,
    const double *dPRporos, const double *dPRporc1, const double *dPRporc2,
    const double *dBCVal,
    const double *ran, const double *rbn, const double *rgn,
    const double *rac, const double *rbc, const double *rgc,
    const double *dju, const double *djv,
    const double *xec, const double *yec, const double *xzn, const double *yzn,
    const double *xen, const double *yen, const double *xzc, const double *yzc,
    const double *xeu, const double *yeu, const double *xzu, const double *yzu,
    const double *xev, const double *yev, const double *xzv, const double *yzv,
    const double *d, const double *dn,
    const double *un, const double *vn, double *us, double *vs) {
    const int nx = *nx_, ny = *ny_, mqiter = *mqiter_;
    const double qtol = *qtol_;
    double *dus = W_dus, *dvs = W_dvs;
    int i, j, m;
    int32_t ret = -1;                                   /* :111 */
    for (j = 1; j <= ny + 1; ++j)                       /* :114-119 */
        for (i = 1; i <= nx + 1; ++i) { A(us, i, j) = A(un, i, j); A(vs, i, j) = A(vn, i, j); }
    for (m = 1; m <= mqiter; ++m) {                     /* :122 */
        orc_veloutflowbcs_(nx_, ny_, nReg, nRegBrd, nMomBdTp, dBCVal, us, vs);   /* :133 */
        for (j = 0; j <= ny + 1; ++j)                   /* :139-144 */
            for (i = 0; i <= nx + 1; ++i) { A(dus, i, j) = 0.0; A(dvs, i, j) = 0.0; }
        orc_xmomentum_(nx_, ny_, nReg, nRegBrd, nRegType, nMomBdTp, dk, re,
                       dPRporos, dPRporc1, dPRporc2, rbn, rgn, rac, rbc, dju,
                       xec, yec, xzn, yzn, xeu, yeu, xzu, yzu, us, vs, un, vn, dus);   /* :147 */
        orc_ymomentum_(nx_, ny_, nReg, nRegBrd, nRegType, nMomBdTp, dk, re, fr,
                       dPRporos, dPRporc1, dPRporc2, ran, rbn, rbc, rgc, djv,
                       xen, yen, xzc, yzc, xev, yev, xzv, yzv, d, dn, us, vs, un, vn, dvs); /* :158 */
        if (orc_errflag) return ret;
        for (j = 1; j <= ny; ++j)                       /* :171-176 */
            for (i = 1; i <= nx; ++i) {
                A(us, i, j) = A(us, i, j) + A(dus, i, j);
                A(vs, i, j) = A(vs, i, j) + A(dvs, i, j);
            }
        double dif1 = orc_dmaxnorm_(nx_, ny_, dus);     /* :179-180 */
        double dif2 = orc_dmaxnorm_(nx_, ny_, dvs);
        orc_last_ql_dif[0] = dif1; orc_last_ql_dif[1] = dif2;
        double difmax = dif1 > dif2 ? dif1 : dif2;      /* :182 */
        if (difmax <= qtol) { ret = m; return ret; }    /* :185-188 */
    }
    return ret;
}

/* =============================== pressure.f ===================================== */

/* Divergence, src/pressure.f:265-323 */
void orc_divergence_(const int32_t *nx_, const int32_t *ny_, const int32_t *nloc_,
                     const double *xet, const double *yet, const double *xzi, const double *yzi,
                     const double *u, const double *v, double *div) {
    const int nx = *nx_, ny = *ny_, nloc = *nloc_;
    int i, j;
    double uci1j, ucij, vcij1, vcij;
    switch (nloc) {
    case 1: /* :287-299 */
        for (j = 1; j <= ny; ++j)
            for (i = 1; i <= nx; ++i) {
                ucij = A(yet, i, j) * A(u, i, j) - A(xet, i, j) * (A(v, i + 1, j) + A(v, i, j)
                       + A(v, i + 1, j - 1) + A(v, i, j - 1)) / 4.0;
                uci1j = A(yet, i - 1, j) * A(u, i - 1, j) - A(xet, i - 1, j) * (A(v, i, j)
                        + A(v, i - 1, j) + A(v, i, j - 1) + A(v, i - 1, j - 1)) / 4.0;
                vcij = A(xzi, i, j) * A(v, i, j) - A(yzi, i, j) * (A(u, i, j + 1) + A(u, i - 1, j + 1)
                       + A(u, i, j) + A(u, i - 1, j)) / 4.0;
                vcij1 = A(xzi, i, j - 1) * A(v, i, j - 1) - A(yzi, i, j - 1) * (A(u, i, j)
                        + A(u, i - 1, j) + A(u, i, j - 1) + A(u, i - 1, j - 1)) / 4.0;
                A(div, i, j) = ucij - uci1j + vcij - vcij1;
            }
        break;
    case 2: /* :302-314 */
        for (j = 1; j <= ny; ++j)
            for (i = 1; i <= nx; ++i) {
                uci1j = A(yet, i + 1, j) * (A(u, i, j + 1) + A(u, i + 1, j + 1) + A(u, i, j)
                        + A(u, i + 1, j)) / 4.0 - A(xet, i + 1, j) * A(v, i + 1, j);
                ucij = A(yet, i, j) * (A(u, i - 1, j + 1) + A(u, i, j + 1) + A(u, i - 1, j)
                       + A(u, i, j)) / 4.0 - A(xet, i, j) * A(v, i, j);
                vcij1 = A(xzi, i, j + 1) * (A(v, i, j + 1) + A(v, i + 1, j + 1) + A(v, i, j)
                        + A(v, i + 1, j)) / 4.0 - A(yzi, i, j + 1) * A(u, i, j + 1);
                vcij = A(xzi, i, j) * (A(v, i, j) + A(v, i + 1, j) + A(v, i, j - 1)
                       + A(v, i + 1, j - 1)) / 4.0 - A(yzi, i, j) * A(u, i, j);
                A(div, i, j) = uci1j - ucij + vcij1 - vcij;
            }
        break;
    default:
        fprintf(stderr, "Error: Wrong location flag passed to Divergence: %d\n", nloc);
        orc_errflag = 1;
    }
}

/* RhsPpe, src/pressure.f:329-378 */
void orc_rhsppe_(const int32_t *nx_, const int32_t *ny_, const int32_t *lCartesGrid,
                 const double *dk_, const double *rbu, const double *rbv, const double *div,
                 const double *p, double *b) {
    const int nx = *nx_, ny = *ny_;
    const double dk = *dk_;
    int i, j, ind;
    for (j = 2; j <= ny; ++j)
        for (i = 2; i <= nx; ++i) { ind = (j - 2) * (nx - 1) + i - 1; b[ind - 1] = A(div, i, j) / dk; }
    if (!*lCartesGrid)
        for (j = 2; j <= ny; ++j)
            for (i = 2; i <= nx; ++i) {
                ind = (j - 2) * (nx - 1) + i - 1;
                b[ind - 1] = b[ind - 1]
                    - (A(rbu, i, j) * (A(p, i + 1, j + 1) + A(p, i, j + 1) - A(p, i + 1, j - 1) - A(p, i, j - 1))
                       - A(rbu, i - 1, j) * (A(p, i, j + 1) + A(p, i - 1, j + 1) - A(p, i, j - 1) - A(p, i - 1, j - 1))
                       + A(rbv, i, j) * (A(p, i + 1, j + 1) + A(p, i + 1, j) - A(p, i - 1, j + 1) - A(p, i - 1, j))
                       - A(rbv, i, j - 1) * (A(p, i + 1, j) + A(p, i + 1, j - 1) - A(p, i - 1, j) - A(p, i - 1, j - 1)));
            }
}

/* a(mn,5) SoA: a(ind,k) -> a[(ind-1) + mn*(k-1)] */
#define PA(ind, k) a[((size_t)(ind) - 1) + (size_t)MN * ((k) - 1)]

#define SOR_POINT() do { \
        ind = (j - 2) * (nx - 1) + i - 1; \
        sum = b[ind - 1] - PA(ind, 1) * A(p, i, j - 1) - PA(ind, 2) * A(p, i - 1, j) \
              - PA(ind, 4) * A(p, i + 1, j) - PA(ind, 5) * A(p, i, j + 1); \
        sum = sum / PA(ind, 3) - A(p, i, j); \
        A(p, i, j) = A(p, i, j) + sorrel * sum; \
    } while (0)

/* Sor, src/pressure.f:384-450 */
void orc_sor_(const int32_t *nx_, const int32_t *ny_, const int32_t *lCartesGrid,
              const int32_t *msorit_, int32_t *nConv, const double *dk, const double *sortol_,
              const double *sorrel_, const double *rbu, const double *rbv, const double *a,
              double *b, const double *div, double *p) {
    const int nx = *nx_, ny = *ny_, msorit = *msorit_;
    const double sortol = *sortol_, sorrel = *sorrel_;
    int m, i, j, ind;
    double dif, sum;
    *nConv = 0;
    if (*lCartesGrid) orc_rhsppe_(nx_, ny_, lCartesGrid, dk, rbu, rbv, div, p, b);
    for (m = 1; m <= msorit; ++m) {
        if (!*lCartesGrid) orc_rhsppe_(nx_, ny_, lCartesGrid, dk, rbu, rbv, div, p, b);
        dif = 0.0;
        for (j = 2; j <= ny; ++j)
            for (i = 2; i <= nx; ++i) {
                SOR_POINT();
                sum = fabs(sum);
                dif = dif > sum ? dif : sum;
            }
        if (m > 1 && dif < sortol) { *nConv = m; return; }
    }
}

/* SorRB, src/pressure.f:457-541 */
void orc_sorrb_(const int32_t *nx_, const int32_t *ny_, const int32_t *lCartesGrid,
                const int32_t *msorit_, int32_t *nConv, const double *dk, const double *sortol_,
                const double *sorrel_, const double *rbu, const double *rbv, const double *a,
                double *b, const double *div, double *p) {
    const int nx = *nx_, ny = *ny_, msorit = *msorit_;
    const double sortol = *sortol_, sorrel = *sorrel_;
    int m, i, j, ind;
    double dif, sum;
    *nConv = 0;
    if (*lCartesGrid) orc_rhsppe_(nx_, ny_, lCartesGrid, dk, rbu, rbv, div, p, b);
    for (m = 1; m <= msorit; ++m) {
        if (!*lCartesGrid) orc_rhsppe_(nx_, ny_, lCartesGrid, dk, rbu, rbv, div, p, b);
        dif = 0.0;
        for (j = 2; j <= ny; ++j)                              /* black, :505-517 */
            for (i = 2 + (j % 2); i <= nx; i += 2) {
                SOR_POINT();
                sum = fabs(sum);
                dif = dif > sum ? dif : sum;
            }
        for (j = 2; j <= ny; ++j)                              /* red, :520-532 */
            for (i = 2 + ((j + 1) % 2); i <= nx; i += 2) {
                SOR_POINT();
                sum = fabs(sum);
                dif = dif > sum ? dif : sum;
            }
        if (m > 1 && dif < sortol) { *nConv = m; return; }
    }
}

/* SorRBP, src/pressure.f:548-656 (dif held per point, max taken afterwards) */
void orc_sorrbp_(const int32_t *nx_, const int32_t *ny_, const int32_t *lCartesGrid,
                 const int32_t *msorit_, int32_t *nConv, const double *dk, const double *sortol_,
                 const double *sorrel_, const double *rbu, const double *rbv, const double *a,
                 double *b, const double *div, double *p) {
    const int nx = *nx_, ny = *ny_, msorit = *msorit_;
    const double sortol = *sortol_, sorrel = *sorrel_;
    int m, i, j, ind;
    double difmax, sum;
    double *dif = W_dif; /* dif(0:nx+1,0:ny+1); indexed with the global pitch here */
    *nConv = 0;
    if (*lCartesGrid) orc_rhsppe_(nx_, ny_, lCartesGrid, dk, rbu, rbv, div, p, b);
    for (m = 1; m <= msorit; ++m) {
        if (!*lCartesGrid) orc_rhsppe_(nx_, ny_, lCartesGrid, dk, rbu, rbv, div, p, b);
        for (j = 1; j <= ny; ++j)
            for (i = 1; i <= nx; ++i) A(dif, i, j) = 0.0;
        for (j = 2; j <= ny; ++j)
            for (i = 2 + (j % 2); i <= nx; i += 2) {
                SOR_POINT();
                sum = fabs(sum);
                A(dif, i, j) = A(dif, i, j) > sum ? A(dif, i, j) : sum;
            }
        for (j = 2; j <= ny; ++j)
            for (i = 2 + ((j + 1) % 2); i <= nx; i += 2) {
                SOR_POINT();
                sum = fabs(sum);
                A(dif, i, j) = A(dif, i, j) > sum ? A(dif, i, j) : sum;
            }
        difmax = 0.0;
        for (j = 2; j <= ny; ++j)
            for (i = 2; i <= nx; ++i) difmax = difmax > A(dif, i, j) ? difmax : A(dif, i, j);
        if (m > 1 && difmax < sortol) { *nConv = m; return; }
    }
}

/* common line set-up of Slor / SlorRB for one line nl (pressure.f:750-772, 905-927) */
static void slor_fill_line(int nx, int ndir, int nl, int nlines, int ncomps, int indal, int indar,
                           int indbl, int indbr, const double *a, const double *b,
                           const double *p, double *aline, double *bline, double *plold) {
    (void)nx;
    for (int nc = 2; nc <= ncomps + 1; ++nc) {
        int ind; double plprev, plnext;
        if (ndir == 0) {
            ind = (nc - 2) * nlines + nl - 1;
            plold[nc - 2] = A(p, nl, nc); plprev = A(p, nl - 1, nc); plnext = A(p, nl + 1, nc);
        } else {
            ind = (nl - 2) * ncomps + nc - 1;
            plold[nc - 2] = A(p, nc, nl); plprev = A(p, nc, nl - 1); plnext = A(p, nc, nl + 1);
        }
        aline[0 + 3 * (nc - 2)] = PA(ind, indal);
        aline[1 + 3 * (nc - 2)] = PA(ind, 3);
        aline[2 + 3 * (nc - 2)] = PA(ind, indar);
        bline[nc - 2] = b[ind - 1] - (PA(ind, indbl) * plprev + PA(ind, indbr) * plnext);
    }
}

static int slor_dirs(int ndir, int nx, int ny, int *nlines, int *ncomps, int *indal, int *indar,
                     int *indbl, int *indbr) {
    if (ndir == 0) { *nlines = nx - 1; *ncomps = ny - 1; *indal = 1; *indar = 5; *indbl = 2; *indbr = 4; }
    else if (ndir == 1) { *nlines = ny - 1; *ncomps = nx - 1; *indal = 2; *indar = 4; *indbl = 1; *indbr = 5; }
    else { fprintf(stderr, "Error in direction flag passed to SLor\n"); return 0; }
    return 1;
}

/* Slor, src/pressure.f:673-802 */
void orc_slor_(const int32_t *nx_, const int32_t *ny_, const int32_t *lCartesGrid,
               const int32_t *ndir_, const int32_t *msorit_, int32_t *nConv, const double *dk,
               const double *sortol_, const double *sorrel_, const double *rbu, const double *rbv,
               const double *a, double *b, const double *div, double *p) {
    const int nx = *nx_, ny = *ny_, ndir = *ndir_, msorit = *msorit_;
    const double sortol = *sortol_, sorrel = *sorrel_;
    int nlines, ncomps, indal, indar, indbl, indbr, m, nc, nl;
    double sum, dif;
    double *aline = W_aline, *bline = W_bline, *plold = W_plold;
    *nConv = 0;
    if (!slor_dirs(ndir, nx, ny, &nlines, &ncomps, &indal, &indar, &indbl, &indbr)) return;
    if (*lCartesGrid) orc_rhsppe_(nx_, ny_, lCartesGrid, dk, rbu, rbv, div, p, b);
    for (m = 1; m <= msorit; ++m) {
        dif = 0.0;
        if (!*lCartesGrid) orc_rhsppe_(nx_, ny_, lCartesGrid, dk, rbu, rbv, div, p, b);
        for (nl = 2; nl <= nlines + 1; ++nl) {
            slor_fill_line(nx, ndir, nl, nlines, ncomps, indal, indar, indbl, indbr, a, b, p, aline, bline, plold);
            { int32_t n = ncomps; orc_alttridlu_(&n, aline, bline); }
            for (nc = 1; nc <= ncomps; ++nc) {
                sum = bline[nc - 1] - plold[nc - 1];
                bline[nc - 1] = plold[nc - 1] + sorrel * sum;
                sum = fabs(sum);
                dif = dif > sum ? dif : sum;
            }
            for (nc = 2; nc <= ncomps + 1; ++nc) {
                if (ndir == 0) A(p, nl, nc) = bline[nc - 2];
                if (ndir == 1) A(p, nc, nl) = bline[nc - 2];
            }
        }
        if (m > 1 && dif < sortol) { *nConv = m; return; }
    }
}

/* SlorRB, src/pressure.f:819-959 */
void orc_slorrb_(const int32_t *nx_, const int32_t *ny_, const int32_t *lCartesGrid,
                 const int32_t *ndir_, const int32_t *msorit_, int32_t *nConv, const double *dk,
                 const double *sortol_, const double *sorrel_, const double *rbu, const double *rbv,
                 const double *a, double *b, const double *div, double *p) {
    const int nx = *nx_, ny = *ny_, ndir = *ndir_, msorit = *msorit_;
    const double sortol = *sortol_, sorrel = *sorrel_;
    int nlines, ncomps, indal, indar, indbl, indbr, m, k, nc, nl;
    double sum, dif;
    double *aline = W_aline, *bline = W_bline, *plold = W_plold;
    *nConv = 0;
    if (!slor_dirs(ndir, nx, ny, &nlines, &ncomps, &indal, &indar, &indbl, &indbr)) return;
    if (*lCartesGrid) orc_rhsppe_(nx_, ny_, lCartesGrid, dk, rbu, rbv, div, p, b);
    for (m = 1; m <= msorit; ++m) {
        dif = 0.0;
        for (k = 2; k <= 3; ++k) {
            if (!*lCartesGrid) orc_rhsppe_(nx_, ny_, lCartesGrid, dk, rbu, rbv, div, p, b);
            for (nl = k; nl <= nlines + 1; nl += 2) {
                slor_fill_line(nx, ndir, nl, nlines, ncomps, indal, indar, indbl, indbr, a, b, p, aline, bline, plold);
                { int32_t n = ncomps; orc_alttridlu_(&n, aline, bline); }
                for (nc = 1; nc <= ncomps; ++nc) {
                    sum = bline[nc - 1] - plold[nc - 1];
                    bline[nc - 1] = plold[nc - 1] + sorrel * sum;
                    sum = fabs(sum);
                    dif = dif > sum ? dif : sum;
                }
                for (nc = 2; nc <= ncomps + 1; ++nc) {
                    if (ndir == 0) A(p, nl, nc) = bline[nc - 2];
                    if (ndir == 1) A(p, nc, nl) = bline[nc - 2];
                }
            }
        }
        if (m > 1 && dif < sortol) { *nConv = m; return; }
    }
}

/* SlorRBP, src/pressure.f:976-1138 */
void orc_slorrbp_(const int32_t *nx_, const int32_t *ny_, const int32_t *lCartesGrid,
                  const int32_t *msorit_, int32_t *nConv, const double *dk,
                  const double *sortol_, const double *sorrel_, const double *rbu, const double *rbv,
                  const double *a, double *b, const double *div, double *p) {
    const int nx = *nx_, ny = *ny_, msorit = *msorit_;
    const double sortol = *sortol_, sorrel = *sorrel_;
    const int nlines = ny - 1, ncomps = nx - 1, indal = 2, indar = 4, indbl = 1, indbr = 5;
    int ind, m, nc, nl, pass;
    double sum, dif;
    double *pn = W_pn, *aline = W_aline, *bline = W_bline;
    *nConv = 0;
    if (*lCartesGrid) orc_rhsppe_(nx_, ny_, lCartesGrid, dk, rbu, rbv, div, p, b);
    for (m = 1; m <= msorit; ++m) {
        for (nl = 1; nl <= nlines + 2; ++nl)                   /* :1031-1035 */
            for (nc = 1; nc <= ncomps + 2; ++nc) A(pn, nc, nl) = A(p, nc, nl);
        for (pass = 2; pass <= 3; ++pass) {                    /* red nl=2,4,.. then black nl=3,5,.. */
            if (!*lCartesGrid) orc_rhsppe_(nx_, ny_, lCartesGrid, dk, rbu, rbv, div, p, b);
            for (nl = pass; nl <= nlines + 1; nl += 2) {
                for (nc = 2; nc <= ncomps + 1; ++nc) {
                    ind = (nl - 2) * ncomps + nc - 1;
                    aline[0 + 3 * (nc - 2)] = PA(ind, indal);
                    aline[1 + 3 * (nc - 2)] = PA(ind, 3);
                    aline[2 + 3 * (nc - 2)] = PA(ind, indar);
                    bline[nc - 2] = b[ind - 1] - (PA(ind, indbl) * A(p, nc, nl - 1)
                                                  + PA(ind, indbr) * A(p, nc, nl + 1));
                }
                { int32_t n = ncomps; orc_alttridlu_(&n, aline, bline); }
                for (nc = 2; nc <= ncomps + 1; ++nc) A(p, nc, nl) = bline[nc - 2];
            }
            for (nl = pass; nl <= nlines + 1; nl += 2)         /* over-relaxation :1073-1077 */
                for (nc = 2; nc <= ncomps + 1; ++nc)
                    A(p, nc, nl) = A(pn, nc, nl) + sorrel * (A(p, nc, nl) - A(pn, nc, nl));
        }
        dif = 0.0;                                             /* :1122-1128 */
        for (nl = 2; nl <= nlines + 1; ++nl)
            for (nc = 2; nc <= ncomps + 1; ++nc) {
                sum = fabs(A(pn, nc, nl) - A(p, nc, nl));
                dif = dif > sum ? dif : sum;
            }
        if (m > 1 && dif < sortol) { *nConv = m; return; }
    }
}

/* Matrix + blockage rows of Ppe, src/pressure.f:96-196 (split out so tests can reach it) */
void orc_ppe_matrix(int nx, int ny, const int32_t *nReg, const int32_t *nRegBrd,
                    const int32_t *nRegType, const double *rau, const double *rgv,
                    double *a, double *div) {
    int i, j, ind, ireg, jreg;
    for (j = 2; j <= ny; ++j)
        for (i = 2; i <= nx; ++i) {
            ind = (j - 2) * (nx - 1) + i - 1;
            PA(ind, 1) = A(rgv, i, j - 1);
            PA(ind, 2) = A(rau, i - 1, j);
            PA(ind, 3) = -A(rau, i, j) - A(rau, i - 1, j) - A(rgv, i, j) - A(rgv, i, j - 1);
            PA(ind, 4) = A(rau, i, j);
            PA(ind, 5) = A(rgv, i, j);
        }
#define PID(ind) do { PA(ind, 1) = 0.0; PA(ind, 2) = 0.0; PA(ind, 3) = 1.0; PA(ind, 4) = 0.0; PA(ind, 5) = 0.0; } while (0)
    for (jreg = 1; jreg <= nReg[_J_]; ++jreg)
        for (ireg = 1; ireg <= nReg[_I_]; ++ireg) {
            if (RT(ireg, jreg) != RM_BLOCKG) continue;
            int iW = RB(ireg, jreg, WEST), iE = RB(ireg, jreg, EAST);
            int jS = RB(ireg, jreg, SOUTH), jN = RB(ireg, jreg, NORTH);
            for (j = jS + 2; j <= jN - 1; ++j)
                for (i = iW + 2; i <= iE - 1; ++i) {
                    ind = (j - 2) * (nx - 1) + i - 1; PID(ind); A(div, i, j) = 0.0;
                }
            if (ireg > 1 && RT(ireg - 1, jreg) == RM_BLOCKG)
                for (j = jS + 1; j <= jN; ++j) { ind = (j - 2) * (nx - 1) + iW; PID(ind); A(div, iW + 1, j) = 0.0; }
            if (ireg < nReg[_I_] && RT(ireg + 1, jreg) == RM_BLOCKG)
                for (j = jS + 1; j <= jN; ++j) { ind = (j - 2) * (nx - 1) + iE - 1; PID(ind); A(div, iE, j) = 0.0; }
            if (jreg > 1 && RT(ireg, jreg - 1) == RM_BLOCKG)
                for (i = iW + 1; i <= iE; ++i) { ind = (jS - 1) * (nx - 1) + i - 1; PID(ind); A(div, i, jS + 1) = 0.0; }
            if (jreg < nReg[_J_] && RT(ireg, jreg + 1) == RM_BLOCKG)
                for (i = iW + 1; i <= iE; ++i) { ind = (jN - 2) * (nx - 1) + i - 1; PID(ind); A(div, i, jN) = 0.0; }
        }
#undef PID
}

/* Ppe, src/pressure.f:30-249 */
void orc_ppe_(const int32_t *nx_, const int32_t *ny_,
              const int32_t *nReg, const int32_t *nRegBrd, const int32_t *nRegType,
              const int32_t *lCartesGrid,
              const int32_t *nPpeSolver, const int32_t *msorit, int32_t *nSorConv,
              const double *dk, const double *sortol, const double *sorrel,
              const double *rau, const double *rbu, const double *rbv, const double *rgv,
              const double *xeu, const double *yeu, const double *xzv, const double *yzv,
              const double *u, const double *v, double *p) {
    const int nx = *nx_, ny = *ny_;
    double *div = W_div, *a = W_pa, *b = W_pb;
    const int32_t one = 1;
    orc_divergence_(nx_, ny_, &one, xeu, yeu, xzv, yzv, u, v, div);     /* :92 */
    orc_ppe_matrix(nx, ny, nReg, nRegBrd, nRegType, rau, rgv, a, div);  /* :97-196 */
    *nSorConv = 0;                                                       /* :199 */
    switch (*nPpeSolver) {                                               /* :201-240 */
    case 1: orc_sor_(nx_, ny_, lCartesGrid, msorit, nSorConv, dk, sortol, sorrel, rbu, rbv, a, b, div, p); break;
    case 2: orc_slor_(nx_, ny_, lCartesGrid, &one, msorit, nSorConv, dk, sortol, sorrel, rbu, rbv, a, b, div, p); break;
    case 3: orc_slorrb_(nx_, ny_, lCartesGrid, &one, msorit, nSorConv, dk, sortol, sorrel, rbu, rbv, a, b, div, p); break;
    case 4: orc_slorrbp_(nx_, ny_, lCartesGrid, msorit, nSorConv, dk, sortol, sorrel, rbu, rbv, a, b, div, p); break;
    case 5: orc_sorrb_(nx_, ny_, lCartesGrid, msorit, nSorConv, dk, sortol, sorrel, rbu, rbv, a, b, div, p); break;
    case 6: orc_sorrbp_(nx_, ny_, lCartesGrid, msorit, nSorConv, dk, sortol, sorrel, rbu, rbv, a, b, div, p); break;
    default: fprintf(stderr, "Wrong nPpeSolver flag passed to Ppe\n");
    }
    if (*nSorConv < 1) *nSorConv = *msorit;    /* :242-246 (warning text dropped) */
}
/* Test hook: 1 when the last Ppe hit the msorit cap. Computed by callers from nSorConv. */
#undef PA

/* ============================== bound_cond.f ==================================== */

/* VelBoundCond, src/bound_cond.f:511-847 */
void orc_velboundcond_(const int32_t *nx_, const int32_t *ny_, const int32_t *nReg,
                       const int32_t *nRegBrd, const int32_t *nMomBdTp, const double *dBCVal,
                       double *u, double *v) {
    (void)nx_; (void)ny_;
    const double dZero = 0.0, dTwo = 2.0, dThree = 3.0, dFour = 4.0, dFive = 5.0, dEight = 8.0;
    int i, j;
    for (int jreg = 1; jreg <= nReg[_J_]; ++jreg)
        for (int ireg = 1; ireg <= nReg[_I_]; ++ireg) {
            int iW = RB(ireg, jreg, WEST), iE = RB(ireg, jreg, EAST);
            int jS = RB(ireg, jreg, SOUTH), jN = RB(ireg, jreg, NORTH);
            /* WEST :561-619 */
            switch (MB(ireg, jreg, WEST)) {
            case BM_INTERN: break;
            case BM_WALL1:
                for (j = jS; j <= jN; ++j) A(u, iW, j) = dZero;
                for (j = jS + 1; j <= jN; ++j) A(v, iW, j) = dTwo * BV(ireg, jreg, WEST, _V_) - A(v, iW + 1, j);
                break;
            case BM_WALL2:
                for (j = jS; j <= jN; ++j) A(u, iW, j) = dZero;
                for (j = jS + 1; j <= jN; ++j) A(v, iW, j) = A(v, iW + 1, j);
                break;
            case BM_INLET:
                for (j = jS; j <= jN; ++j) A(u, iW, j) = BV(ireg, jreg, WEST, _U_);
                for (j = jS + 1; j <= jN; ++j) A(v, iW, j) = dTwo * BV(ireg, jreg, WEST, _V_) - A(v, iW + 1, j);
                break;
            case BM_OUTLT1:
                for (j = jS; j <= jN; ++j) A(u, iW - 1, j) = BV(ireg, jreg, WEST, _U_) + A(u, iW, j);
                for (j = jS + 1; j <= jN; ++j) A(v, iW, j) = -A(v, iW + 1, j);
                break;
            case BM_OUTLT2:
                for (j = jS + 1; j <= jN; ++j) A(u, iW, j) = A(u, iW + 1, j) + A(v, iW + 1, j) - A(v, iW + 1, j - 1);
                for (j = jS + 1; j <= jN; ++j)
                    A(v, iW, j) = -A(v, iW, j - 1) + dFive * (A(v, iW + 1, j) - A(v, iW + 1, j - 1))
                                  + dEight * (A(u, iW + 1, j) - A(u, iW, j));
                break;
            default:
                fprintf(stderr, "Wrong nBdTypeW flag in region %d,%d\n", ireg, jreg); orc_errflag = 1; return;
            }
            /* EAST :623-693 */
            switch (MB(ireg, jreg, EAST)) {
            case BM_INTERN: break;
            case BM_WALL1:
                for (j = jS; j <= jN; ++j) A(u, iE, j) = dZero;
                for (j = jS + 1; j <= jN; ++j) A(v, iE + 1, j) = dTwo * BV(ireg, jreg, EAST, _V_) - A(v, iE, j);
                break;
            case BM_WALL2:
                for (j = jS; j <= jN; ++j) A(u, iE, j) = dZero;
                for (j = jS + 1; j <= jN; ++j) A(v, iE + 1, j) = A(v, iE, j);
                break;
            case BM_INLET:
                for (j = jS; j <= jN; ++j) A(u, iE, j) = BV(ireg, jreg, EAST, _U_);
                for (j = jS + 1; j <= jN; ++j) A(v, iE + 1, j) = dTwo * BV(ireg, jreg, EAST, _V_) - A(v, iE, j);
                break;
            case BM_OUTLT1:
                for (j = jS; j <= jN; ++j) A(u, iE + 1, j) = BV(ireg, jreg, EAST, _U_) + A(u, iE, j);
                for (j = jS + 1; j <= jN; ++j) A(v, iE + 1, j) = -A(v, iE, j);
                break;
            case BM_OUTLT2:
                for (j = jS + 1; j <= jN; ++j) A(u, iE, j) = A(u, iE - 1, j) - (A(v, iE, j) - A(v, iE, j - 1));
                for (j = jS + 1; j <= jN - 1; ++j)
                    A(v, iE + 1, j) = A(v, iE + 1, j - 1) + dThree * (A(v, iE, j - 1) - A(v, iE, j))
                                      - dFour * (A(u, iE, j) - A(u, iE - 1, j));
                break;
            default:
                fprintf(stderr, "Wrong nBdTypeE flag in region %d,%d\n", ireg, jreg); orc_errflag = 1; return;
            }
            /* SOUTH :697-766 */
            switch (MB(ireg, jreg, SOUTH)) {
            case BM_INTERN: break;
            case BM_WALL1:
                for (i = iW + 1; i <= iE; ++i) A(u, i, jS) = dTwo * BV(ireg, jreg, SOUTH, _U_) - A(u, i, jS + 1);
                for (i = iW; i <= iE; ++i) A(v, i, jS) = dZero;
                break;
            case BM_WALL2:
                for (i = iW + 1; i <= iE; ++i) A(u, i, jS) = A(u, i, jS + 1);
                for (i = iW; i <= iE; ++i) A(v, i, jS) = dZero;
                break;
            case BM_INLET:
                for (i = iW + 1; i <= iE; ++i) A(u, i, jS) = dTwo * BV(ireg, jreg, SOUTH, _U_) - A(u, i, jS + 1);
                for (i = iW; i <= iE; ++i) A(v, i, jS) = BV(ireg, jreg, SOUTH, _V_);
                break;
            case BM_OUTLT1:
                for (i = iW + 1; i <= iE; ++i) A(u, i, jS) = -A(u, i, jS + 1);
                for (i = iW; i <= iE; ++i) A(v, i, jS) = BV(ireg, jreg, SOUTH, _V_) + A(v, i, jS);  /* sic */
                break;
            case BM_OUTLT2:
                for (i = iW + 1; i <= iE; ++i) A(v, i, jS) = A(v, i, jS + 1) + (A(u, i, jS + 1) - A(u, i - 1, jS + 1));
                for (i = iW + 1; i <= iE - 1; ++i)
                    A(u, i, jS) = A(u, i - 1, jS) + dThree * (A(u, i - 1, jS + 1) - A(u, i, jS + 1))
                                  - dFour * (A(v, i, jS + 1) - A(v, i, jS));
                break;
            default:
                fprintf(stderr, "Wrong nBdTypeS flag in region %d,%d\n", ireg, jreg); orc_errflag = 1; return;
            }
            /* NORTH :770-840 */
            switch (MB(ireg, jreg, NORTH)) {
            case BM_INTERN: break;
            case BM_WALL1:
                for (i = iW + 1; i <= iE; ++i) A(u, i, jN + 1) = dTwo * BV(ireg, jreg, NORTH, _U_) - A(u, i, jN);
                for (i = iW; i <= iE; ++i) A(v, i, jN) = dZero;
                break;
            case BM_WALL2:
                for (i = iW + 1; i <= iE; ++i) A(u, i, jN + 1) = A(u, i, jN);
                for (i = iW; i <= iE; ++i) A(v, i, jN) = dZero;
                break;
            case BM_INLET:
                for (i = iW + 1; i <= iE; ++i) A(u, i, jN + 1) = dTwo * BV(ireg, jreg, NORTH, _U_) - A(u, i, jN);
                for (i = iW; i <= iE; ++i) A(v, i, jN) = BV(ireg, jreg, NORTH, _V_);
                break;
            case BM_OUTLT1:
                for (i = iW + 1; i <= iE; ++i) A(u, i, jN + 1) = -A(u, i, jN);
                for (i = iW; i <= iE; ++i) A(v, i, jN + 1) = BV(ireg, jreg, NORTH, _V_) + A(v, i, jN);
                break;
            case BM_OUTLT2:
                for (i = iW; i <= iE; ++i) A(v, i, jN) = A(v, i, jN - 1) - (A(u, i, jN) - A(u, i - 1, jN));
                for (i = iW + 1; i <= iE - 1; ++i)
                    A(u, i, jN + 1) = A(u, i - 1, jN + 1) + dThree * (A(u, i - 1, jN) - A(u, i, jN))
                                      - dFour * (A(v, i, jN) - A(v, i, jN - 1));
                break;
            default:
                fprintf(stderr, "Wrong nBdTypeN flag in region %d,%d\n", ireg, jreg); orc_errflag = 1; return;
            }
        }
}

/* PresBoundCond, src/bound_cond.f:853-1024 */
void orc_presboundcond_(const int32_t *nx_, const int32_t *ny_, const int32_t *nReg,
                        const int32_t *nRegBrd, const int32_t *nRegType, const int32_t *nMomBdTp,
                        const double *dBCVal, double *p) {
    (void)nx_; (void)ny_;
    const double dZero = 0.0, dTwo = 2.0;
    int i, j;
    for (int jreg = 1; jreg <= nReg[_J_]; ++jreg)
        for (int ireg = 1; ireg <= nReg[_I_]; ++ireg) {
            int iW = RB(ireg, jreg, WEST), iE = RB(ireg, jreg, EAST);
            int jS = RB(ireg, jreg, SOUTH), jN = RB(ireg, jreg, NORTH);
            if (RT(ireg, jreg) == RM_BLOCKG) { /* :903-936 */
                for (j = jS + 1; j <= jN; ++j)
                    for (i = iW + 1; i <= iE; ++i) A(p, i, j) = dZero;
                for (j = jS + 1; j <= jN; ++j) A(p, iW + 1, j) = BV(ireg, jreg, WEST, _P_) + A(p, iW, j);
                for (j = jS + 1; j <= jN; ++j) A(p, iE, j) = BV(ireg, jreg, EAST, _P_) + A(p, iE + 1, j);
                for (i = iW + 1; i <= iE; ++i) A(p, i, jS + 1) = BV(ireg, jreg, SOUTH, _P_) + A(p, i, jS);
                for (i = iW + 1; i <= iE; ++i) A(p, i, jN) = BV(ireg, jreg, NORTH, _P_) + A(p, i, jN + 1);
                continue;
            }
            switch (MB(ireg, jreg, WEST)) {
            case BM_INTERN: break;
            case BM_WALL1: case BM_WALL2: case BM_INLET:
                for (j = jS + 1; j <= jN; ++j) A(p, iW, j) = BV(ireg, jreg, WEST, _P_) + A(p, iW + 1, j);
                break;
            case BM_OUTLT1: case BM_OUTLT2:
                for (j = jS + 1; j <= jN; ++j) A(p, iW, j) = dTwo * BV(ireg, jreg, WEST, _P_) - A(p, iW + 1, j);
                break;
            default:
                fprintf(stderr, "Wrong nBdTypeW flag in region %d,%d\n", ireg, jreg); orc_errflag = 1; return;
            }
            switch (MB(ireg, jreg, EAST)) {
            case BM_INTERN: break;
            case BM_WALL1: case BM_WALL2: case BM_INLET:
                for (j = jS + 1; j <= jN; ++j) A(p, iE + 1, j) = BV(ireg, jreg, EAST, _P_) + A(p, iE, j);
                break;
            case BM_OUTLT1: case BM_OUTLT2:
                for (j = jS + 1; j <= jN; ++j) A(p, iE + 1, j) = dTwo * BV(ireg, jreg, EAST, _P_) - A(p, iE, j);
                break;
            default:
                fprintf(stderr, "Wrong nBdTypeE flag in region %d,%d\n", ireg, jreg); orc_errflag = 1; return;
            }
            switch (MB(ireg, jreg, SOUTH)) {
            case BM_INTERN: break;
            case BM_WALL1: case BM_WALL2: case BM_INLET:
                for (i = iW + 1; i <= iE; ++i) A(p, i, jS) = BV(ireg, jreg, SOUTH, _P_) + A(p, i, jS + 1);
                break;
            case BM_OUTLT1: case BM_OUTLT2:
                for (i = iW + 1; i <= iE; ++i) A(p, i, jS) = dTwo * BV(ireg, jreg, SOUTH, _P_) - A(p, i, jS + 1);
                break;
            default:
                fprintf(stderr, "Wrong nBdTypeS flag in region %d,%d\n", ireg, jreg); orc_errflag = 1; return;
            }
            switch (MB(ireg, jreg, NORTH)) {
            case BM_INTERN: break;
            case BM_WALL1: case BM_WALL2: case BM_INLET:
                for (i = iW + 1; i <= iE; ++i) A(p, i, jN + 1) = BV(ireg, jreg, NORTH, _P_) + A(p, i, jN);
                break;
            case BM_OUTLT1: case BM_OUTLT2:
                for (i = iW + 1; i <= iE; ++i) A(p, i, jN + 1) = dTwo * BV(ireg, jreg, NORTH, _P_) - A(p, i, jN);
                break;
            default:
                fprintf(stderr, "Wrong nBdTypeN flag in region %d,%d\n", ireg, jreg); orc_errflag = 1; return;
            }
        }
}

/* VelOutflowBCs, src/bound_cond.f:1656-1874 */
void orc_veloutflowbcs_(const int32_t *nx_, const int32_t *ny_, const int32_t *nReg,
                        const int32_t *nRegBrd, const int32_t *nMomBdTp, const double *dBCVal,
                        double *u, double *v) {
    (void)nx_; (void)ny_;
    const double dThree = 3.0, dFour = 4.0, dFive = 5.0, dEight = 8.0;
    int i, j;
    for (int jreg = 1; jreg <= nReg[_J_]; ++jreg)
        for (int ireg = 1; ireg <= nReg[_I_]; ++ireg) {
            int iW = RB(ireg, jreg, WEST), iE = RB(ireg, jreg, EAST);
            int jS = RB(ireg, jreg, SOUTH), jN = RB(ireg, jreg, NORTH);
            switch (MB(ireg, jreg, WEST)) { /* :1707-1736 */
            case BM_INTERN: case BM_WALL1: case BM_WALL2: case BM_INLET: break;
            case BM_OUTLT1:
                for (j = jS; j <= jN; ++j) A(u, iW - 1, j) = BV(ireg, jreg, WEST, _U_) + A(u, iW, j);
                for (j = jS + 1; j <= jN; ++j) A(v, iW, j) = -A(v, iW + 1, j);
                break;
            case BM_OUTLT2: /* note the sign pattern differs from VelBoundCond (:1725 vs :608) */
                for (j = jS + 1; j <= jN; ++j) A(u, iW, j) = A(u, iW + 1, j) - A(v, iW + 1, j) + A(v, iW + 1, j - 1);
                for (j = jS + 1; j <= jN; ++j)
                    A(v, iW, j) = -A(v, iW, j - 1) + dFive * (A(v, iW + 1, j) - A(v, iW + 1, j - 1))
                                  + dEight * (A(u, iW + 1, j) - A(u, iW, j));
                break;
            default:
                fprintf(stderr, "Wrong nBdTypeW flag in region %d,%d\n", ireg, jreg); orc_errflag = 1; return;
            }
            switch (MB(ireg, jreg, EAST)) { /* :1740-1780 */
            case BM_INTERN: case BM_WALL1: case BM_WALL2: case BM_INLET: break;
            case BM_OUTLT1:
                for (j = jS; j <= jN; ++j) A(u, iE + 1, j) = BV(ireg, jreg, EAST, _U_) + A(u, iE, j);
                for (j = jS + 1; j <= jN; ++j) A(v, iE + 1, j) = -A(v, iE, j);
                break;
            case BM_OUTLT2:
                for (j = jS + 1; j <= jN; ++j) A(u, iE, j) = A(u, iE - 1, j) - (A(v, iE, j) - A(v, iE, j - 1));
                for (j = jS + 1; j <= jN - 1; ++j)
                    A(v, iE + 1, j) = A(v, iE + 1, j - 1) + dThree * (A(v, iE, j - 1) - A(v, iE, j))
                                      - dFour * (A(u, iE, j) - A(u, iE - 1, j));
                break;
            default:
                fprintf(stderr, "Wrong nBdTypeE flag in region %d,%d\n", ireg, jreg); orc_errflag = 1; return;
            }
            switch (MB(ireg, jreg, SOUTH)) { /* :1784-1824 */
            case BM_INTERN: case BM_WALL1: case BM_WALL2: case BM_INLET: break;
            case BM_OUTLT1:
                for (i = iW + 1; i <= iE; ++i) A(u, i, jS) = -A(u, i, jS + 1);
                for (i = iW; i <= iE; ++i) A(v, i, jS) = BV(ireg, jreg, SOUTH, _V_) + A(v, i, jS);
                break;
            case BM_OUTLT2:
                for (i = iW + 1; i <= iE; ++i) A(v, i, jS) = A(v, i, jS + 1) + (A(u, i, jS + 1) - A(u, i - 1, jS + 1));
                for (i = iW + 1; i <= iE - 1; ++i)
                    A(u, i, jS) = A(u, i - 1, jS) + dThree * (A(u, i - 1, jS + 1) - A(u, i, jS + 1))
                                  - dFour * (A(v, i, jS + 1) - A(v, i, jS));
                break;
            default:
                fprintf(stderr, "Wrong nBdTypeS flag in region %d,%d\n", ireg, jreg); orc_errflag = 1; return;
            }
            switch (MB(ireg, jreg, NORTH)) { /* :1828-1868 */
            case BM_INTERN: case BM_WALL1: case BM_WALL2: case BM_INLET: break;
            case BM_OUTLT1:
                for (i = iW + 1; i <= iE; ++i) A(u, i, jN + 1) = -A(u, i, jN);
                for (i = iW; i <= iE; ++i) A(v, i, jN + 1) = BV(ireg, jreg, NORTH, _V_) + A(v, i, jN);
                break;
            case BM_OUTLT2:
                for (i = iW; i <= iE; ++i) A(v, i, jN) = A(v, i, jN - 1) - (A(u, i, jN) - A(u, i - 1, jN));
                for (i = iW + 1; i <= iE - 1; ++i)
                    A(u, i, jN + 1) = A(u, i - 1, jN + 1) + dThree * (A(u, i - 1, jN) - A(u, i, jN))
                                      - dFour * (A(v, i, jN) - A(v, i, jN - 1));
                break;
            default:
                fprintf(stderr, "Wrong nBdTypeN flag in region %d,%d\n", ireg, jreg); orc_errflag = 1; return;
            }
        }
}

/* ================================ utility.f ====================================== */

/* Filter, src/utility.f:33-247 (ncomp = _U_, _V_, _T_) */
void orc_filter_(const int32_t *nx_, const int32_t *ny_, const int32_t *ncomp_,
                 const int32_t *nReg, const int32_t *nRegBrd, const int32_t *nRegType,
                 const int32_t *nMomBdTp, const int32_t *nTRgType, const double *fp_, double *qu) {
    const int nx = *nx_, ny = *ny_, ncomp = *ncomp_;
    const double fp = *fp_, dFour = 4.0;
    double *qh = W_qh;
    int i, j, ireg, jreg;
#define FILT(i, j) A(qh, i, j) = (A(qu, i, (j) - 1) + A(qu, (i) - 1, j) + A(qu, i, (j) + 1) \
                                  + A(qu, (i) + 1, j) + fp * A(qu, i, j)) / (fp + dFour)
    for (j = 0; j <= ny + 1; ++j)
        for (i = 0; i <= nx + 1; ++i) A(qh, i, j) = A(qu, i, j);
    switch (ncomp) {
    case _U_:
        for (jreg = 1; jreg <= nReg[_J_]; ++jreg)
            for (ireg = 1; ireg <= nReg[_I_]; ++ireg) {
                if (RT(ireg, jreg) == RM_BLOCKG) continue;
                int iW = RB(ireg, jreg, WEST), iE = RB(ireg, jreg, EAST);
                int jS = RB(ireg, jreg, SOUTH), jN = RB(ireg, jreg, NORTH);
                for (j = jS + 1; j <= jN; ++j)
                    for (i = iW + 1; i <= iE - 1; ++i) FILT(i, j);
                switch (MB(ireg, jreg, WEST)) {
                case BM_OUTLT1: for (j = jS + 1; j <= jN; ++j) FILT(iW, j); break;
                case BM_INTERN: case BM_WALL1: case BM_WALL2: case BM_INLET: case BM_OUTLT2: break;
                default: fprintf(stderr, "Wrong nBdTypeW flag in region %d,%d\n", ireg, jreg);
                }
                switch (MB(ireg, jreg, EAST)) {
                case BM_INTERN: case BM_OUTLT1: for (j = jS + 1; j <= jN; ++j) FILT(iE, j); break;
                case BM_WALL1: case BM_WALL2: case BM_INLET: case BM_OUTLT2: break;
                default: fprintf(stderr, "Wrong nBdTypeE flag in region %d,%d\n", ireg, jreg);
                }
            }
        break;
    case _V_:
        for (jreg = 1; jreg <= nReg[_J_]; ++jreg)
            for (ireg = 1; ireg <= nReg[_I_]; ++ireg) {
                if (RT(ireg, jreg) == RM_BLOCKG) continue;
                int iW = RB(ireg, jreg, WEST), iE = RB(ireg, jreg, EAST);
                int jS = RB(ireg, jreg, SOUTH), jN = RB(ireg, jreg, NORTH);
                for (j = jS + 1; j <= jN - 1; ++j)
                    for (i = iW + 1; i <= iE; ++i) FILT(i, j);
                switch (MB(ireg, jreg, SOUTH)) {
                case BM_OUTLT1: for (i = iW + 1; i <= iE; ++i) FILT(i, jS); break;
                case BM_INTERN: case BM_WALL1: case BM_WALL2: case BM_INLET: case BM_OUTLT2: break;
                default: fprintf(stderr, "Wrong nBdTypeS flag in region %d,%d\n", ireg, jreg);
                }
                switch (MB(ireg, jreg, NORTH)) {
                case BM_INTERN: case BM_OUTLT1: for (i = iW + 1; i <= iE; ++i) FILT(i, jN); break;
                case BM_WALL1: case BM_WALL2: case BM_INLET: case BM_OUTLT2: break;
                default: fprintf(stderr, "Wrong nBdTypeN flag in region %d,%d\n", ireg, jreg);
                }
            }
        break;
    case _T_:
        for (jreg = 1; jreg <= nReg[_J_]; ++jreg)
            for (ireg = 1; ireg <= nReg[_I_]; ++ireg) {
                if (nTRgType[(ireg - 1) + MGRI * (jreg - 1)] == 1 /* BT_TEMPER (sic) */) continue;
                int iW = RB(ireg, jreg, WEST), iE = RB(ireg, jreg, EAST);
                int jS = RB(ireg, jreg, SOUTH), jN = RB(ireg, jreg, NORTH);
                for (j = jS + 1; j <= jN; ++j)
                    for (i = iW + 1; i <= iE; ++i) FILT(i, j);
            }
        break;
    default:
        fprintf(stderr, "Wrong ncomp flag passed to Filter\n"); orc_errflag = 1; return;
    }
    for (j = 0; j <= ny + 1; ++j)
        for (i = 0; i <= nx + 1; ++i) A(qu, i, j) = A(qh, i, j);
#undef FILT
}

/* Project, src/utility.f:253-440 */
void orc_project_(const int32_t *nx_, const int32_t *ny_,
                  const int32_t *nReg, const int32_t *nRegBrd, const int32_t *nRegType,
                  const int32_t *nMomBdTp, const double *dk_,
                  const double *dju, const double *djv,
                  const double *yeu, const double *xzv, const double *yzu, const double *xev,
                  const double *p, double *u, double *v) {
    (void)nx_; (void)ny_;
    const double dk = *dk_, dFour = 4.0;
    int i, j, ireg, jreg;
    double pzi, pet, djk;
#define PROJ_U(i, j) do { \
        pzi = A(p, (i) + 1, j) - A(p, i, j); \
        pet = (A(p, (i) + 1, (j) + 1) + A(p, i, (j) + 1) - A(p, (i) + 1, (j) - 1) - A(p, i, (j) - 1)) / dFour; \
        djk = A(dju, i, j) * dk; \
        A(u, i, j) = A(u, i, j) - djk * (A(yeu, i, j) * pzi - A(yzu, i, j) * pet); } while (0)
#define PROJ_V(i, j) do { \
        pzi = (A(p, (i) + 1, (j) + 1) + A(p, (i) + 1, j) - A(p, (i) - 1, (j) + 1) - A(p, (i) - 1, j)) / dFour; \
        pet = A(p, i, (j) + 1) - A(p, i, j); \
        djk = A(djv, i, j) * dk; \
        A(v, i, j) = A(v, i, j) - djk * (-A(xev, i, j) * pzi + A(xzv, i, j) * pet); } while (0)
    for (jreg = 1; jreg <= nReg[_J_]; ++jreg)
        for (ireg = 1; ireg <= nReg[_I_]; ++ireg) {
            if (RT(ireg, jreg) == RM_BLOCKG) continue;
            int iW = RB(ireg, jreg, WEST), iE = RB(ireg, jreg, EAST);
            int jS = RB(ireg, jreg, SOUTH), jN = RB(ireg, jreg, NORTH);
            for (j = jS + 1; j <= jN; ++j)
                for (i = iW + 1; i <= iE - 1; ++i) PROJ_U(i, j);
            switch (MB(ireg, jreg, WEST)) {
            case BM_OUTLT1: for (j = jS + 1; j <= jN; ++j) PROJ_U(iW, j); break;
            case BM_INTERN: case BM_WALL1: case BM_WALL2: case BM_INLET: case BM_OUTLT2: break;
            default: fprintf(stderr, "Wrong nBdTypeW flag in region %d,%d\n", ireg, jreg);
            }
            switch (MB(ireg, jreg, EAST)) {
            case BM_INTERN: case BM_OUTLT1: for (j = jS + 1; j <= jN; ++j) PROJ_U(iE, j); break;
            case BM_WALL1: case BM_WALL2: case BM_INLET: case BM_OUTLT2: break;
            default: fprintf(stderr, "Wrong nBdTypeE flag in region %d,%d\n", ireg, jreg);
            }
        }
    for (jreg = 1; jreg <= nReg[_J_]; ++jreg)
        for (ireg = 1; ireg <= nReg[_I_]; ++ireg) {
            if (RT(ireg, jreg) == RM_BLOCKG) continue;
            int iW = RB(ireg, jreg, WEST), iE = RB(ireg, jreg, EAST);
            int jS = RB(ireg, jreg, SOUTH), jN = RB(ireg, jreg, NORTH);
            for (j = jS + 1; j <= jN - 1; ++j)
                for (i = iW + 1; i <= iE; ++i) PROJ_V(i, j);
            switch (MB(ireg, jreg, SOUTH)) {
            case BM_OUTLT1: for (i = iW + 1; i <= iE; ++i) PROJ_V(i, jS); break;
            case BM_INTERN: case BM_WALL1: case BM_WALL2: case BM_INLET: case BM_OUTLT2: break;
            default: fprintf(stderr, "Wrong nBdTypeS flag in region %d,%d\n", ireg, jreg);
            }
            switch (MB(ireg, jreg, NORTH)) {
            case BM_INTERN: case BM_OUTLT1: for (i = iW + 1; i <= iE; ++i) PROJ_V(i, jN); break;
            case BM_WALL1: case BM_WALL2: case BM_INLET: case BM_OUTLT2: break;
            default: fprintf(stderr, "Wrong nBdTypeN flag in region %d,%d\n", ireg, jreg);
            }
        }
#undef PROJ_U
#undef PROJ_V
}

/* DiffMaxNorm, src/utility.f:446-473 */
double orc_diffmaxnorm_(const int32_t *nx_, const int32_t *ny_, const double *un, const double *u) {
    const int nx = *nx_, ny = *ny_;
    double diff = fabs(A(un, 2, 2) - A(u, 2, 2));
    for (int j = 2; j <= ny - 1; ++j)
        for (int i = 2; i <= nx - 1; ++i) {
            double x = fabs(A(un, i, j) - A(u, i, j));
            diff = diff > x ? diff : x;
        }
    return diff;
}

/* DMaxNorm, src/utility.f:479-507 */
double orc_dmaxnorm_(const int32_t *nx_, const int32_t *ny_, const double *u) {
    const int nx = *nx_, ny = *ny_;
    double diff = fabs(A(u, 5, 5));
    for (int j = 2; j <= ny - 1; ++j)
        for (int i = 2; i <= nx - 1; ++i) {
            double x = fabs(A(u, i, j));
            diff = diff > x ? diff : x;
        }
    return diff;
}

/* ================================== grid.f ====================================== */

/* MirrorPts, src/grid.f:258-306 */
void orc_mirrorpts_(const int32_t *nx_, const int32_t *ny_, double *x, double *y) {
    const int nx = *nx_, ny = *ny_;
    const double dTwo = 2.0;
    int i, j;
    for (i = 1; i <= nx; ++i) {
        A(x, i, 0) = dTwo * A(x, i, 1) - A(x, i, 2);
        A(x, i, ny + 1) = dTwo * A(x, i, ny) - A(x, i, ny - 1);
        A(y, i, 0) = dTwo * A(y, i, 1) - A(y, i, 2);
        A(y, i, ny + 1) = dTwo * A(y, i, ny) - A(y, i, ny - 1);
    }
    for (j = 1; j <= ny; ++j) {
        A(x, 0, j) = dTwo * A(x, 1, j) - A(x, 2, j);
        A(x, nx + 1, j) = dTwo * A(x, nx, j) - A(x, nx - 1, j);
        A(y, 0, j) = dTwo * A(y, 1, j) - A(y, 2, j);
        A(y, nx + 1, j) = dTwo * A(y, nx, j) - A(y, nx - 1, j);
    }
    A(x, 0, 0) = dTwo * A(x, 1, 1) - A(x, 2, 2);
    A(y, 0, 0) = dTwo * A(y, 1, 1) - A(y, 2, 2);
    A(x, 0, ny + 1) = dTwo * A(x, 1, ny) - A(x, 2, ny - 1);
    A(y, 0, ny + 1) = dTwo * A(y, 1, ny) - A(y, 2, ny - 1);
    A(x, nx + 1, 0) = dTwo * A(x, nx, 1) - A(x, nx - 1, 2);
    A(y, nx + 1, 0) = dTwo * A(y, nx, 1) - A(y, nx - 1, 2);
    A(x, nx + 1, ny + 1) = dTwo * A(x, nx, ny) - A(x, nx - 1, ny - 1);
    A(y, nx + 1, ny + 1) = dTwo * A(y, nx, ny) - A(y, nx - 1, ny - 1);
}

/* FullGrid, src/grid.f:312-362 */
void orc_fullgrid_(const int32_t *nx_, const int32_t *ny_, const double *x, const double *y,
                   double *xu, double *yu, double *xv, double *yv, double *xc, double *yc) {
    const int nx = *nx_, ny = *ny_;
    const double dHalf = 0.5;
    int i, j;
    for (j = 1; j <= ny + 1; ++j)
        for (i = 0; i <= nx + 1; ++i) {
            A(xu, i, j) = dHalf * (A(x, i, j) + A(x, i, j - 1));
            A(yu, i, j) = dHalf * (A(y, i, j) + A(y, i, j - 1));
        }
    for (j = 0; j <= ny + 1; ++j)
        for (i = 1; i <= nx + 1; ++i) {
            A(xv, i, j) = dHalf * (A(x, i, j) + A(x, i - 1, j));
            A(yv, i, j) = dHalf * (A(y, i, j) + A(y, i - 1, j));
        }
    for (j = 1; j <= ny + 1; ++j)
        for (i = 1; i <= nx + 1; ++i) {
            A(xc, i, j) = dHalf * (A(xu, i, j) + A(xu, i - 1, j));
            A(yc, i, j) = dHalf * (A(yv, i, j) + A(yv, i, j - 1));
        }
}

/* Metric, src/grid.f:368-535.  m = 30 output arrays in wolfd2_metrics order. */
void orc_metric_(const int32_t *nx_, const int32_t *ny_,
                 const double *xn, const double *yn, const double *xu, const double *yu,
                 const double *xv, const double *yv, const double *xc, const double *yc,
                 double *rau, double *rbu, double *rbv, double *rgv,
                 double *ran, double *rbn, double *rgn,
                 double *rac, double *rbc, double *rgc,
                 double *dju, double *djv, double *djc, double *djn,
                 double *xen, double *yen, double *xzn, double *yzn,
                 double *xec, double *yec, double *xzc, double *yzc,
                 double *xeu, double *yeu, double *xzv, double *yzv,
                 double *xzu, double *yzu, double *xev, double *yev) {
    const int nx = *nx_, ny = *ny_;
    const double dOne = 1.0, dQrtr = 0.25;
    int i, j;
    double g11, g12, g22;
    for (j = 1; j <= ny; ++j)
        for (i = 1; i <= nx; ++i) { /* :425-448 */
            A(xzn, i, j) = A(xv, i + 1, j) - A(xv, i, j);
            A(xen, i, j) = A(xu, i, j + 1) - A(xu, i, j);
            A(yzn, i, j) = A(yv, i + 1, j) - A(yv, i, j);
            A(yen, i, j) = A(yu, i, j + 1) - A(yu, i, j);
            A(djn, i, j) = dOne / (A(xzn, i, j) * A(yen, i, j) - A(xen, i, j) * A(yzn, i, j));
            g11 = A(xzn, i, j) * A(xzn, i, j) + A(yzn, i, j) * A(yzn, i, j);
            g12 = A(xzn, i, j) * A(xen, i, j) + A(yzn, i, j) * A(yen, i, j);
            g22 = A(xen, i, j) * A(xen, i, j) + A(yen, i, j) * A(yen, i, j);
            A(ran, i, j) = A(djn, i, j) * g22;
            A(rbn, i, j) = -A(djn, i, j) * g12 * dQrtr;
            A(rgn, i, j) = A(djn, i, j) * g11;
        }
    for (j = 1; j <= ny; ++j)
        for (i = 1; i <= nx; ++i) { /* :451-477 */
            A(xzu, i, j) = A(xc, i + 1, j) - A(xc, i, j);
            A(xeu, i, j) = A(xn, i, j) - A(xn, i, j - 1);
            A(yzu, i, j) = A(yc, i + 1, j) - A(yc, i, j);
            A(yeu, i, j) = A(yn, i, j) - A(yn, i, j - 1);
            A(dju, i, j) = dOne / (A(xzu, i, j) * A(yeu, i, j) - A(xeu, i, j) * A(yzu, i, j));
            g11 = A(xzu, i, j) * A(xzu, i, j) + A(yzu, i, j) * A(yzu, i, j);
            g12 = A(xzu, i, j) * A(xeu, i, j) + A(yzu, i, j) * A(yeu, i, j);
            g22 = A(xeu, i, j) * A(xeu, i, j) + A(yeu, i, j) * A(yeu, i, j);
            (void)g11;
            A(rau, i, j) = A(dju, i, j) * g22;
            A(rbu, i, j) = -A(dju, i, j) * g12 * dQrtr;
        }
    for (j = 1; j <= ny; ++j)
        for (i = 1; i <= nx; ++i) { /* :480-506 */
            A(xzv, i, j) = A(xn, i, j) - A(xn, i - 1, j);
            A(xev, i, j) = A(xc, i, j + 1) - A(xc, i, j);
            A(yzv, i, j) = A(yn, i, j) - A(yn, i - 1, j);
            A(yev, i, j) = A(yc, i, j + 1) - A(yc, i, j);
            A(djv, i, j) = dOne / (A(xzv, i, j) * A(yev, i, j) - A(xev, i, j) * A(yzv, i, j));
            g11 = A(xzv, i, j) * A(xzv, i, j) + A(yzv, i, j) * A(yzv, i, j);
            g12 = A(xzv, i, j) * A(xev, i, j) + A(yzv, i, j) * A(yev, i, j);
            g22 = A(xev, i, j) * A(xev, i, j) + A(yev, i, j) * A(yev, i, j);
            (void)g22;
            A(rbv, i, j) = -A(djv, i, j) * g12 * dQrtr;
            A(rgv, i, j) = A(djv, i, j) * g11;
        }
    for (j = 1; j <= ny; ++j)
        for (i = 1; i <= nx; ++i) { /* :509-532 */
            A(xzc, i, j) = A(xu, i, j) - A(xu, i - 1, j);
            A(xec, i, j) = A(xv, i, j) - A(xv, i, j - 1);
            A(yzc, i, j) = A(yu, i, j) - A(yu, i - 1, j);
            A(yec, i, j) = A(yv, i, j) - A(yv, i, j - 1);
            A(djc, i, j) = dOne / (A(xzc, i, j) * A(yec, i, j) - A(xec, i, j) * A(yzc, i, j));
            g11 = A(xzc, i, j) * A(xzc, i, j) + A(yzc, i, j) * A(yzc, i, j);
            g12 = A(xzc, i, j) * A(xec, i, j) + A(yzc, i, j) * A(yec, i, j);
            g22 = A(xec, i, j) * A(xec, i, j) + A(yec, i, j) * A(yec, i, j);
            A(rac, i, j) = A(djc, i, j) * g22;
            A(rbc, i, j) = -A(djc, i, j) * g12 * dQrtr;
            A(rgc, i, j) = A(djc, i, j) * g11;
        }
}

/* Grid, src/grid.f:31-127 minus the file reader: gx,gy hold nx*ny node coordinates on entry
 * (1..nx,1..ny), are scaled by dlref, mirrored, and the 30 metric arrays filled.
 * m[] is in wolfd2_metrics order; all arrays must be zero-initialised by the caller (F5). */
void orc_grid(int32_t nx, int32_t ny, double dlref, double *gx, double *gy, double **m) {
    double *xu = zalloc(NFULL), *yu = zalloc(NFULL), *xv = zalloc(NFULL), *yv = zalloc(NFULL),
           *xc = zalloc(NFULL), *yc = zalloc(NFULL);
    for (int j = 1; j <= ny; ++j)                /* :98-103 */
        for (int i = 1; i <= nx; ++i) { A(gx, i, j) = A(gx, i, j) / dlref; A(gy, i, j) = A(gy, i, j) / dlref; }
    orc_mirrorpts_(&nx, &ny, gx, gy);
    orc_fullgrid_(&nx, &ny, gx, gy, xu, yu, xv, yv, xc, yc);
    orc_metric_(&nx, &ny, gx, gy, xu, yu, xv, yv, xc, yc,
                m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], m[12], m[13],
                m[14], m[15], m[16], m[17], m[18], m[19], m[20], m[21], m[22], m[23], m[24], m[25],
                m[26], m[27], m[28], m[29]);
    free(xu); free(yu); free(xv); free(yv); free(xc); free(yc);
}

/* ============================ BC table set-up ==================================== */

/* InitBCFlags, src/parse.f:2257-2379 (momentum part) */
void orc_initbcflags(const int32_t *nReg, int32_t *nRegType, int32_t *nMomBdTp, double *dBCVal) {
    int ireg, jreg, k, l;
    for (l = 1; l <= 4; ++l)
        for (k = 1; k <= 4; ++k)
            for (jreg = 1; jreg <= nReg[_J_]; ++jreg)
                for (ireg = 1; ireg <= nReg[_I_]; ++ireg) BV(ireg, jreg, k, l) = 0.0;
    for (jreg = 1; jreg <= nReg[_J_]; ++jreg)
        for (ireg = 1; ireg <= nReg[_I_]; ++ireg) {
            RT(ireg, jreg) = RM_INTERN;
            for (k = WEST; k <= NORTH; ++k) MB(ireg, jreg, k) = BM_INTERN;
        }
    for (jreg = 1; jreg <= nReg[_J_]; ++jreg) { MB(1, jreg, WEST) = BM_WALL1; MB(nReg[_I_], jreg, EAST) = BM_WALL1; }
    for (ireg = 1; ireg <= nReg[_I_]; ++ireg) { MB(ireg, 1, SOUTH) = BM_WALL1; MB(ireg, nReg[_J_], NORTH) = BM_WALL1; }
}

/* `blockage ir jr` statement, src/parse.f:1317-1344 */
void orc_bc_blockage(const int32_t *nReg, int32_t *nRegType, int32_t *nMomBdTp, int32_t ireg, int32_t jreg) {
    RT(ireg, jreg) = RM_BLOCKG;
    for (int k = WEST; k <= NORTH; ++k) MB(ireg, jreg, k) = BM_WALL1;
    if (ireg > 1) MB(ireg - 1, jreg, EAST) = BM_WALL1;
    if (ireg < nReg[_I_]) MB(ireg + 1, jreg, WEST) = BM_WALL1;
    if (jreg > 1) MB(ireg, jreg - 1, NORTH) = BM_WALL1;
    if (jreg < nReg[_J_]) MB(ireg, jreg + 1, SOUTH) = BM_WALL1;
}

/* Border completion + BC-type mirroring of SetUpBCs, src/bound_cond.f:165-223.
 * On entry nRegBrd(ireg,1,WEST) for ireg>=2 and nRegBrd(1,jreg,SOUTH) for jreg>=2 hold the
 * i_borders / j_borders of the deck. */
void orc_setupbcs_complete(int32_t nx, int32_t ny, const int32_t *nReg, int32_t *nRegBrd,
                           int32_t *nMomBdTp) {
    int ireg, jreg;
    for (jreg = 1; jreg <= nReg[_J_]; ++jreg) { RB(1, jreg, WEST) = 1; RB(nReg[_I_], jreg, EAST) = nx; }
    for (ireg = 1; ireg <= nReg[_I_]; ++ireg) { RB(ireg, 1, SOUTH) = 1; RB(ireg, nReg[_J_], NORTH) = ny; }
    for (jreg = 2; jreg <= nReg[_J_]; ++jreg)
        for (ireg = 2; ireg <= nReg[_I_]; ++ireg) {
            RB(ireg, jreg, WEST) = RB(ireg, 1, WEST);
            RB(ireg, jreg, SOUTH) = RB(1, jreg, SOUTH);
        }
    for (jreg = 1; jreg <= nReg[_J_]; ++jreg)
        for (ireg = 1; ireg <= nReg[_I_] - 1; ++ireg) RB(ireg, jreg, EAST) = RB(ireg + 1, 1, WEST);
    for (jreg = 1; jreg <= nReg[_J_] - 1; ++jreg)
        for (ireg = 1; ireg <= nReg[_I_]; ++ireg) RB(ireg, jreg, NORTH) = RB(1, jreg + 1, SOUTH);
    for (jreg = 1; jreg <= nReg[_J_]; ++jreg)
        for (ireg = 1; ireg <= nReg[_I_] - 1; ++ireg) {
            if (MB(ireg, jreg, EAST) == BM_INLET) MB(ireg + 1, jreg, WEST) = BM_INLET;
            if (MB(ireg + 1, jreg, WEST) == BM_INLET) MB(ireg, jreg, EAST) = BM_INLET;
            if (MB(ireg, jreg, EAST) == BM_WALL1) MB(ireg + 1, jreg, WEST) = BM_WALL1;
            if (MB(ireg + 1, jreg, WEST) == BM_WALL1) MB(ireg, jreg, EAST) = BM_WALL1;
        }
    for (jreg = 1; jreg <= nReg[_J_] - 1; ++jreg)
        for (ireg = 1; ireg <= nReg[_I_]; ++ireg) {
            if (MB(ireg, jreg + 1, SOUTH) == BM_INLET) MB(ireg, jreg, NORTH) = BM_INLET;
            if (MB(ireg, jreg, NORTH) == BM_INLET) MB(ireg, jreg + 1, SOUTH) = BM_INLET;
            if (MB(ireg, jreg + 1, SOUTH) == BM_WALL1) MB(ireg, jreg, NORTH) = BM_WALL1;
            if (MB(ireg, jreg, NORTH) == BM_WALL1) MB(ireg, jreg + 1, SOUTH) = BM_WALL1;
        }
}


/* ================================ thermal.f ====================================== */
#define TRT(ir, jr) nTRgType[((ir) - 1) + MGRI * ((jr) - 1)]
#define TB(ir, jr, k) nTemBdTp[((ir) - 1) + MGRI * (((jr) - 1) + MGRJ * ((k) - 1))]
enum { RT_NOSRCE = 0, RT_HEATGN = 1, RT_TEMPER = 2 };
enum { BT_INTERN = 0, BT_TEMPER = 1, BT_HTFLUX = 2 };

/* TempBoundCond, src/bound_cond.f:1030-1204 */
void orc_tempboundcond_(const int32_t *nx_, const int32_t *ny_, const int32_t *nReg, const int32_t *nRegBrd,
                        const int32_t *nTRgType, const int32_t *nTemBdTp, const double *dTRgVal,
                        const double *dBCVal, double *t) {
    (void)nx_; (void)ny_;
    const double dTwo = 2.0;
    int i, j;
    for (int jreg = 1; jreg <= nReg[_J_]; ++jreg)
        for (int ireg = 1; ireg <= nReg[_I_]; ++ireg) {
            int iW = RB(ireg, jreg, WEST), iE = RB(ireg, jreg, EAST);
            int jS = RB(ireg, jreg, SOUTH), jN = RB(ireg, jreg, NORTH);
            if (TRT(ireg, jreg) == RT_TEMPER) { /* :1083-1113 */
                for (j = jS + 1; j <= jN; ++j)
                    for (i = iW + 1; i <= iE; ++i) A(t, i, j) = PR(dTRgVal, ireg, jreg);
                for (j = jS + 1; j <= jN; ++j) A(t, iW + 1, j) = dTwo * BV(ireg, jreg, WEST, _T_) - A(t, iW, j);
                for (j = jS + 1; j <= jN; ++j) A(t, iE, j) = dTwo * BV(ireg, jreg, EAST, _T_) - A(t, iE + 1, j);
                for (i = iW + 1; i <= iE; ++i) A(t, i, jS + 1) = dTwo * BV(ireg, jreg, SOUTH, _T_) - A(t, i, jS);
                for (i = iW + 1; i <= iE; ++i) A(t, i, jN) = dTwo * BV(ireg, jreg, NORTH, _T_) - A(t, i, jN + 1);
                continue;
            }
            switch (TB(ireg, jreg, WEST)) { /* :1118-1136 */
            case BT_INTERN: break;
            case BT_TEMPER: for (j = jS + 1; j <= jN; ++j) A(t, iW, j) = dTwo * BV(ireg, jreg, WEST, _T_) - A(t, iW + 1, j); break;
            case BT_HTFLUX: for (j = jS + 1; j <= jN; ++j) A(t, iW, j) = BV(ireg, jreg, WEST, _T_) + A(t, iW + 1, j); break;
            default: fprintf(stderr, "Wrong nTemBdTp W flag in region %d,%d\n", ireg, jreg); orc_errflag = 1; return;
            }
            switch (TB(ireg, jreg, EAST)) { /* :1139-1156 */
            case BT_INTERN: break;
            case BT_TEMPER: for (j = jS + 1; j <= jN; ++j) A(t, iE + 1, j) = dTwo * BV(ireg, jreg, EAST, _T_) - A(t, iE, j); break;
            case BT_HTFLUX: for (j = jS + 1; j <= jN; ++j) A(t, iE + 1, j) = BV(ireg, jreg, EAST, _T_) + A(t, iE, j); break;
            default: fprintf(stderr, "Wrong nTemBdTp E flag in region %d,%d\n", ireg, jreg); orc_errflag = 1; return;
            }
            switch (TB(ireg, jreg, SOUTH)) { /* :1160-1177 */
            case BT_INTERN: break;
            case BT_TEMPER: for (i = iW + 1; i <= iE; ++i) A(t, i, jS) = dTwo * BV(ireg, jreg, SOUTH, _T_) - A(t, i, jS + 1); break;
            case BT_HTFLUX: for (i = iW + 1; i <= iE; ++i) A(t, i, jS) = BV(ireg, jreg, SOUTH, _T_) + A(t, i, jS + 1); break;
            default: fprintf(stderr, "Wrong nTemBdTp S flag in region %d,%d\n", ireg, jreg); orc_errflag = 1; return;
            }
            switch (TB(ireg, jreg, NORTH)) { /* :1181-1198 */
            case BT_INTERN: break;
            case BT_TEMPER: for (i = iW + 1; i <= iE; ++i) A(t, i, jN + 1) = dTwo * BV(ireg, jreg, NORTH, _T_) - A(t, i, jN); break;
            case BT_HTFLUX: for (i = iW + 1; i <= iE; ++i) A(t, i, jN + 1) = BV(ireg, jreg, NORTH, _T_) + A(t, i, jN); break;
            default: fprintf(stderr, "Wrong nTemBdTp N flag in region %d,%d\n", ireg, jreg); orc_errflag = 1; return;
            }
        }
}

/* ThermEnergy, src/thermal.f:24-272 */
void orc_thermenergy_(const int32_t *nx_, const int32_t *ny_, const int32_t *nReg, const int32_t *nRegBrd,
                      const int32_t *nTRgType, const int32_t *nTemBdTp, const double *dk_, const double *pe_,
                      const double *dTRgVal, const double *dHGSTval, const double *dBCVal,
                      const double *rau, const double *rbu, const double *rbv, const double *rgv, const double *djc,
                      const double *xeu, const double *yeu, const double *xzv, const double *yzv,
                      const double *xec, const double *yec, const double *xzc, const double *yzc,
                      const double *un, const double *vn, const double *u, const double *v,
                      const double *tn, double *t) {
    const int nx = *nx_, ny = *ny_;
    const double dk = *dk_, pe = *pe_;
    const double dZero = 0.0, dOne = 1.0, dTwo = 2.0, dHalf = 0.5;
    double *cu1 = W_cu1, *cun = W_cun, *cv1 = W_cv1, *cvn = W_cvn, *s = W_s, *a = W_ta, *b = W_tb;
    int i, j, ind;
    double rkj, c, d;
    const double pe1 = dOne / pe;                                   /* :101 */
#define AA(k, ind) a[((k) - 1) + 3 * ((size_t)(ind) - 1)]
    orc_tempboundcond_(nx_, ny_, nReg, nRegBrd, nTRgType, nTemBdTp, dTRgVal, dBCVal, t);   /* :104 */
    if (orc_errflag) return;
    { const int32_t c6 = 6, c3 = 3, z = 0;                          /* :114-119 */
      orc_convcoef_(nx_, ny_, &c6, &z, xzc, xec, yzc, yec, u, v, cu1, cv1);
      orc_convcoef_(nx_, ny_, &c3, &z, xzv, xeu, yzv, yeu, un, vn, cun, cvn); }
    for (int jreg = 1; jreg <= nReg[_J_]; ++jreg)                   /* :123-147 */
        for (int ireg = 1; ireg <= nReg[_I_]; ++ireg) {
            int iW = RB(ireg, jreg, WEST), iE = RB(ireg, jreg, EAST);
            int jS = RB(ireg, jreg, SOUTH), jN = RB(ireg, jreg, NORTH);
            const double val = TRT(ireg, jreg) == RT_HEATGN ? PR(dHGSTval, ireg, jreg) : dZero;
            for (j = jS + 1; j <= jN; ++j)
                for (i = iW + 1; i <= iE; ++i) A(s, i, j) = val;
        }
    for (j = 2; j <= ny; ++j)                                       /* :153-199 */
        for (i = 2; i <= nx; ++i) {
            ind = (j - 2) * (nx - 1) + i - 1;
            rkj = dk * A(djc, i, j) * dHalf;
            if (A(cu1, i, j) >= dZero) {
                AA(1, ind) = rkj * (-A(cu1, i - 1, j) - pe1 * A(rau, i - 1, j));
                AA(2, ind) = dOne + rkj * (A(cu1, i, j) + pe1 * (A(rau, i, j) + A(rau, i - 1, j)));
                AA(3, ind) = rkj * (-pe1 * A(rau, i, j));
            } else {
                AA(1, ind) = rkj * (-pe1 * A(rau, i - 1, j));
                AA(2, ind) = dOne + rkj * (-A(cu1, i, j) + pe1 * (A(rau, i, j) + A(rau, i - 1, j)));
                AA(3, ind) = rkj * (A(cu1, i + 1, j) - pe1 * A(rau, i, j));
            }
            c = -A(cvn, i, j - 1) * A(tn, i, j - 1) - A(cun, i - 1, j) * A(tn, i - 1, j)
                + (A(cun, i, j) - A(cun, i - 1, j) + A(cvn, i, j) - A(cvn, i, j - 1)) * A(tn, i, j)
                + A(cun, i, j) * A(tn, i + 1, j) + A(cvn, i, j) * A(tn, i, j + 1);
            d = A(rau, i, j) * (A(tn, i + 1, j) - A(tn, i, j))
                - A(rau, i - 1, j) * (A(tn, i, j) - A(tn, i - 1, j))
                + A(rbu, i, j) * (A(tn, i + 1, j + 1) + A(tn, i, j + 1) - A(tn, i + 1, j - 1) - A(tn, i, j - 1))
                - A(rbu, i - 1, j) * (A(tn, i, j + 1) + A(tn, i - 1, j + 1) - A(tn, i, j - 1) - A(tn, i - 1, j - 1))
                + A(rbv, i, j) * (A(tn, i + 1, j + 1) + A(tn, i + 1, j) - A(tn, i - 1, j + 1) - A(tn, i - 1, j))
                - A(rbv, i, j - 1) * (A(tn, i + 1, j) + A(tn, i + 1, j - 1) - A(tn, i - 1, j) - A(tn, i - 1, j - 1))
                + A(rgv, i, j) * (A(tn, i, j + 1) - A(tn, i, j))
                - A(rgv, i, j - 1) * (A(tn, i, j) - A(tn, i, j - 1));
            b[ind - 1] = dk * A(s, i, j) * dHalf + dTwo * rkj * (-c + pe1 * d);
        }
    { const int32_t n = (nx - 1) * (ny - 1); orc_alttridlu_(&n, a, b); }   /* :203 */
    for (j = 2; j <= ny; ++j)                                       /* :208-238 (rhs = first-step solution) */
        for (i = 2; i <= nx; ++i) {
            ind = (j - 2) * (nx - 1) + i - 1;
            rkj = dk * A(djc, i, j) * dHalf;
            if (A(cv1, i, j) >= dZero) {
                AA(1, ind) = rkj * (-A(cv1, i, j - 1) - pe1 * A(rgv, i, j - 1));
                AA(2, ind) = dOne + rkj * (A(cv1, i, j) + pe1 * (A(rgv, i, j) + A(rgv, i, j - 1)));
                AA(3, ind) = rkj * (-pe1 * A(rgv, i, j));
            } else {
                AA(1, ind) = rkj * (-pe1 * A(rgv, i, j - 1));
                AA(2, ind) = dOne + rkj * (-A(cv1, i, j) + pe1 * (A(rgv, i, j) + A(rgv, i, j - 1)));
                AA(3, ind) = rkj * (A(cv1, i, j + 1) - pe1 * A(rgv, i, j));
            }
        }
    for (int jreg = 1; jreg <= nReg[_J_]; ++jreg)                   /* :242-266 */
        for (int ireg = 1; ireg <= nReg[_I_]; ++ireg) {
            if (TRT(ireg, jreg) != RT_TEMPER) continue;
            int iW = RB(ireg, jreg, WEST), iE = RB(ireg, jreg, EAST);
            int jS = RB(ireg, jreg, SOUTH), jN = RB(ireg, jreg, NORTH);
            for (j = jS + 1; j <= jN; ++j)
                for (i = iW + 1; i <= iE; ++i) {
                    ind = (j - 2) * (nx - 1) + i - 1;
                    AA(1, ind) = dZero; AA(2, ind) = dOne; AA(3, ind) = dZero; b[ind - 1] = dZero;
                }
        }
    { const int32_t n = (nx - 1) * (ny - 1); orc_alttridlu_(&n, a, b); }   /* :270 */
    for (j = 2; j <= ny; ++j)                                       /* :274-279 */
        for (i = 2; i <= nx; ++i) {
            ind = (j - 2) * (nx - 1) + i - 1;
            A(t, i, j) = A(t, i, j) + b[ind - 1];
        }
#undef AA
}

/* EqState, src/thermal.f:283-327 */
void orc_eqstate_(const int32_t *nx_, const int32_t *ny_, const double *uref_, const double *densref_,
                  const double *tmax_, const double *tref_, const double *rconst_,
                  const double *p, const double *t, double *den) {
    const int nx = *nx_, ny = *ny_;
    const double uref = *uref_, densref = *densref_, tmax = *tmax_, tref = *tref_, rconst = *rconst_;
    const double dZero = 0.0, dOne = 1.0;
    const double pref = densref * rconst * tref;
    for (int j = 2; j <= ny; ++j)
        for (int i = 2; i <= nx; ++i) {
            const double c1 = A(p, i, j) * densref * (uref * uref) + pref;   /* uref**2 */
            const double c2 = densref * rconst * (A(t, i, j) * (tmax - tref) + tref);
            A(den, i, j) = c1 / c2 - dOne;
            if (fabs(A(den, i, j)) < 1.e-10) A(den, i, j) = dZero;
        }
}

#include "wolfd2_oracle_atd.inc"

/* ================================== main.f ======================================= */

typedef struct orc_state {
    double *u, *v, *p, *t, *d;              /* caller-owned state, (0:mnx,0:mny)      */
} orc_state;

static double *S_un, *S_vn, *S_pn, *S_tn, *S_dn, *S_us, *S_vs, *S_ts;
static size_t S_n;
static void step_work(void) {
    if (S_n == NFULL) return;
    double **f[] = {&S_un, &S_vn, &S_pn, &S_tn, &S_dn, &S_us, &S_vs, &S_ts};
    for (int k = 0; k < 8; ++k) { free(*f[k]); *f[k] = zalloc(NFULL); }
    S_n = NFULL;
}

/* Cold-start projection, src/main.f:606-641 */
void orc_coldstart(const wolfd2_params *par, const wolfd2_regions *reg, const wolfd2_metrics *m,
                   double *u, double *v, double *p, int32_t *nSorConv) {
    const int32_t nx = par->nx, ny = par->ny;
    orc_velboundcond_(&nx, &ny, reg->nReg, reg->nRegBrd, reg->nMomBdTp, reg->dBCVal, u, v);
    orc_ppe_(&nx, &ny, reg->nReg, reg->nRegBrd, reg->nRegType, &par->lCartesGrid, &par->nPpeSolver,
             &par->msorit, nSorConv, &par->dk, &par->sortol, &par->sorrel,
             m->rau, m->rbu, m->rbv, m->rgv, m->xeu, m->yeu, m->xzv, m->yzv, u, v, p);
    orc_presboundcond_(&nx, &ny, reg->nReg, reg->nRegBrd, reg->nRegType, reg->nMomBdTp, reg->dBCVal, p);
    orc_project_(&nx, &ny, reg->nReg, reg->nRegBrd, reg->nRegType, reg->nMomBdTp, &par->dk,
                 m->dju, m->djv, m->yeu, m->xzv, m->yzu, m->xev, p, u, v);
    orc_velboundcond_(&nx, &ny, reg->nReg, reg->nRegBrd, reg->nMomBdTp, reg->dBCVal, u, v);
}

/* Step body, src/main.f:690-981 with nsmallscl=0.  th == NULL (or th->nthermen == 0): cold flow;
 * otherwise the momentum-energy iterations with ThermEnergy / EqState (:736-880), Filter(_T_) (:890)
 * and TempBoundCond (:955).  t and d are passed in either case (copies/norms of :696-704, :857-870, :965). */
/* particles of a trajectory run (src/main.f:242-260): what main.f passes to Traject at :1014-1024 */
typedef struct orc_particles {
    const wolfd2_traject *tr;
    const double *gx, *gy;                 /* grid nodes x(0:mnx,0:mny), y */
    const double *cpartx, *cparty, *repc;
    double *xp, *yp, *up, *vp;
    int32_t *nTOutBnd;
} orc_particles;

static void orc_call_smallscale(int32_t initflg, const wolfd2_params *par, const wolfd2_regions *reg, const wolfd2_metrics *m,
                                const wolfd2_thermal *th, const wolfd2_smallscale *ss, const double *u, const double *v,
                                const double *t, double *uss, double *vss, double *pss, double *tss) {
    const int32_t nx = par->nx, ny = par->ny, nthermen = th->nthermen;
    orc_smallscale_(&nx, &ny, &initflg, &nthermen, &par->lCartesGrid, reg->nReg, reg->nRegBrd, reg->nRegType,
                    th->nTRgType, reg->nMomBdTp, th->nTemBdTp, &ss->nssPpeSlvr, &ss->mssSorIt,
                    &ss->dlref, &ss->uref, &ss->tref, &ss->tmax, &par->dk, &par->re, &ss->pe,
                    &ss->ssSorTol, &ss->ssSorRel, ss->ssFiltPar,
                    &ss->ssCu0, &ss->ssTsCoef, &ss->ssHsCoef, &ss->ssTemCoef, &ss->ssBnCrit, &ss->ssRMpMax, &ss->ssRMpExp,
                    th->dTRgVal, reg->dBCVal, m->rau, m->rbu, m->rbv, m->rgv, m->dju, m->djv, m->djc,
                    m->xeu, m->yeu, m->xzv, m->yzv, m->xzu, m->yzu, m->xev, m->yev, m->xec, m->yec, m->xzc, m->yzc,
                    u, v, t, uss, vss, pss, tss);
}
/* src/main.f:643-665: SmallScale with initflg = 0 before the time loop */
void orc_atd_init(const wolfd2_params *par, const wolfd2_regions *reg, const wolfd2_metrics *m, const wolfd2_thermal *th,
                  const wolfd2_smallscale *ss, const double *u, const double *v, const double *t,
                  double *uss, double *vss, double *pss, double *tss) {
    orc_call_smallscale(0, par, reg, m, th, ss, u, v, t, uss, vss, pss, tss);
}

static double *S_usn, *S_vsn, *S_tsn;
static size_t S_usn_n;
/* ss != NULL with nsmallscl == 1: the ATD blocks of :706-727 and :896-940 (th must carry the thermal tables);
 * pt != NULL: the trajectory block of :997-1031 */
int32_t orc_step_full(const wolfd2_params *par, const wolfd2_regions *reg, const wolfd2_metrics *m,
                      const wolfd2_thermal *th, const wolfd2_smallscale *ss, const orc_particles *pt,
                      double *u, double *v, double *p, double *t, double *d,
                      double *uss, double *vss, double *pss, double *tss,
                      int32_t nsteps, wolfd2_step_log *logs) {
    const int nsmallscl = ss ? ss->nsmallscl : 0;
    if (nsmallscl == 1 || pt) {
        if (!S_usn || S_usn_n != NFULL) {
            free(S_usn); free(S_vsn); free(S_tsn);
            S_usn = zalloc(NFULL); S_vsn = zalloc(NFULL); S_tsn = zalloc(NFULL);
            S_usn_n = NFULL;
        }
    }
    double *usn = S_usn, *vsn = S_vsn, *tsn = S_tsn;
    const int32_t nx = par->nx, ny = par->ny;
    const int32_t cU = _U_, cV = _V_, cT = _T_;
    const int nthermen = th ? th->nthermen : 0, neqstate = th ? th->neqstate : 0;
    int i, j;
    step_work();
    double *un = S_un, *vn = S_vn, *pn = S_pn, *tn = S_tn, *dn = S_dn, *us = S_us, *vs = S_vs, *ts = S_ts;
    for (int k = 0; k < nsteps; ++k) {
        for (j = 0; j <= ny + 1; ++j)                                  /* :696-704 */
            for (i = 0; i <= nx + 1; ++i) {
                A(pn, i, j) = A(p, i, j); A(un, i, j) = A(u, i, j); A(vn, i, j) = A(v, i, j);
                A(tn, i, j) = A(t, i, j); A(dn, i, j) = A(d, i, j);
            }
        if (nsmallscl == 1) {                                          /* :706-727 */
            for (j = 0; j <= ny + 1; ++j)
                for (i = 0; i <= nx + 1; ++i) {
                    A(usn, i, j) = A(uss, i, j); A(vsn, i, j) = A(vss, i, j); A(tsn, i, j) = A(tss, i, j);
                }
            for (j = 0; j <= ny + 1; ++j)
                for (i = 0; i <= nx + 1; ++i) {
                    A(un, i, j) = A(un, i, j) + A(uss, i, j); A(vn, i, j) = A(vn, i, j) + A(vss, i, j);
                    A(tn, i, j) = A(tn, i, j) + A(tss, i, j);
                }
        }
        int nmeiter = par->nmeiter;
        if (nthermen != 1 && nmeiter > 0) nmeiter = 1;                 /* :736 */
        int32_t nQLiter = 0, nSorConv = 0;
        for (int l = 1; l <= nmeiter; ++l) {
            for (j = 0; j <= ny + 1; ++j)                              /* :741-747 */
                for (i = 0; i <= nx + 1; ++i) {
                    A(us, i, j) = A(u, i, j); A(vs, i, j) = A(v, i, j); A(ts, i, j) = A(t, i, j);
                }
            nQLiter = orc_nauxmomentum_(&nx, &ny, &par->mqiter, reg->nReg, reg->nRegBrd, reg->nRegType,
                reg->nMomBdTp, &par->dk, &par->re, &par->fr, &par->qtol,
                reg->dPRporos, reg->dPRporc1, reg->dPRporc2, reg->dBCVal,
                m->ran, m->rbn, m->rgn, m->rac, m->rbc, m->rgc, m->dju, m->djv,
                m->xec, m->yec, m->xzn, m->yzn, m->xen, m->yen, m->xzc, m->yzc,
                m->xeu, m->yeu, m->xzu, m->yzu, m->xev, m->yev, m->xzv, m->yzv,
                d, dn, un, vn, us, vs);                                /* :753-768 */
            if (orc_errflag) return 1;
            if (par->nfiltu == 1)                                      /* :783-791 */
                orc_filter_(&nx, &ny, &cU, reg->nReg, reg->nRegBrd, reg->nRegType, reg->nMomBdTp, NULL, &par->fpu, us);
            if (par->nfiltv == 1)
                orc_filter_(&nx, &ny, &cV, reg->nReg, reg->nRegBrd, reg->nRegType, reg->nMomBdTp, NULL, &par->fpv, vs);
            orc_velboundcond_(&nx, &ny, reg->nReg, reg->nRegBrd, reg->nMomBdTp, reg->dBCVal, us, vs);   /* :793 */
            orc_presboundcond_(&nx, &ny, reg->nReg, reg->nRegBrd, reg->nRegType, reg->nMomBdTp, reg->dBCVal, p);
            orc_ppe_(&nx, &ny, reg->nReg, reg->nRegBrd, reg->nRegType, &par->lCartesGrid, &par->nPpeSolver,
                     &par->msorit, &nSorConv, &par->dk, &par->sortol, &par->sorrel,
                     m->rau, m->rbu, m->rbv, m->rgv, m->xeu, m->yeu, m->xzv, m->yzv, us, vs, p);       /* :803 */
            orc_presboundcond_(&nx, &ny, reg->nReg, reg->nRegBrd, reg->nRegType, reg->nMomBdTp, reg->dBCVal, p);
            orc_project_(&nx, &ny, reg->nReg, reg->nRegBrd, reg->nRegType, reg->nMomBdTp, &par->dk,
                         m->dju, m->djv, m->yeu, m->xzv, m->yzu, m->xev, p, us, vs);                   /* :820 */
            orc_velboundcond_(&nx, &ny, reg->nReg, reg->nRegBrd, reg->nMomBdTp, reg->dBCVal, us, vs);   /* :829 */
            orc_presboundcond_(&nx, &ny, reg->nReg, reg->nRegBrd, reg->nRegType, reg->nMomBdTp, reg->dBCVal, p);
            if (nthermen == 1)                                         /* :840-850 */
                orc_thermenergy_(&nx, &ny, reg->nReg, reg->nRegBrd, th->nTRgType, th->nTemBdTp, &par->dk, &th->pe,
                                 th->dTRgVal, th->dHGSTval, reg->dBCVal, m->rau, m->rbu, m->rbv, m->rgv, m->djc,
                                 m->xeu, m->yeu, m->xzv, m->yzv, m->xec, m->yec, m->xzc, m->yzc,
                                 un, vn, us, vs, tn, ts);
            if (orc_errflag) return 1;
            if (neqstate == 1)                                         /* :853-855 */
                orc_eqstate_(&nx, &ny, &th->uref, &th->densref, &th->tmax, &th->tref, &th->rconst, p, ts, d);
            double dme[3];                                             /* :857-859 */
            dme[0] = orc_diffmaxnorm_(&nx, &ny, u, us);
            dme[1] = orc_diffmaxnorm_(&nx, &ny, v, vs);
            dme[2] = orc_diffmaxnorm_(&nx, &ny, t, ts);
            for (j = 0; j <= ny + 1; ++j)                              /* :864-870 */
                for (i = 0; i <= nx + 1; ++i) {
                    A(u, i, j) = A(us, i, j); A(v, i, j) = A(vs, i, j); A(t, i, j) = A(ts, i, j);
                }
            double dmemax = dme[0] > dme[1] ? dme[0] : dme[1];         /* :873-877 */
            dmemax = dmemax > dme[2] ? dmemax : dme[2];
            if (l > 1 && th && dmemax < th->dmeittol) break;
        }
        if (nthermen == 1 && th->nfiltt == 1)                          /* :890-894 */
            orc_filter_(&nx, &ny, &cT, reg->nReg, reg->nRegBrd, reg->nRegType, reg->nMomBdTp, th->nTRgType, &th->fpt, t);
        if (nsmallscl == 1) {                                          /* :896-940 */
            for (j = 0; j <= ny + 1; ++j)
                for (i = 0; i <= nx + 1; ++i) {
                    A(u, i, j) = A(u, i, j) - A(usn, i, j); A(v, i, j) = A(v, i, j) - A(vsn, i, j);
                    A(t, i, j) = A(t, i, j) - A(tsn, i, j);
                }
            orc_call_smallscale(1, par, reg, m, th, ss, u, v, t, uss, vss, pss, tss);
            if (orc_errflag) return 1;
            for (j = 0; j <= ny + 1; ++j)
                for (i = 0; i <= nx + 1; ++i) {
                    A(u, i, j) = A(u, i, j) + A(uss, i, j); A(v, i, j) = A(v, i, j) + A(vss, i, j);
                    A(t, i, j) = A(t, i, j) + A(tss, i, j);
                }
        }
        orc_velboundcond_(&nx, &ny, reg->nReg, reg->nRegBrd, reg->nMomBdTp, reg->dBCVal, u, v);         /* :946 */
        orc_presboundcond_(&nx, &ny, reg->nReg, reg->nRegBrd, reg->nRegType, reg->nMomBdTp, reg->dBCVal, p);
        if (nthermen == 1)                                             /* :955 */
            orc_tempboundcond_(&nx, &ny, reg->nReg, reg->nRegBrd, th->nTRgType, th->nTemBdTp, th->dTRgVal, reg->dBCVal, t);
        double dif[4];
        dif[0] = orc_diffmaxnorm_(&nx, &ny, pn, p);                    /* :962-965 */
        dif[1] = orc_diffmaxnorm_(&nx, &ny, un, u);
        dif[2] = orc_diffmaxnorm_(&nx, &ny, vn, v);
        dif[3] = orc_diffmaxnorm_(&nx, &ny, tn, t);
        double difmax = dif[0];
        for (int q = 1; q < 4; ++q) difmax = difmax > dif[q] ? difmax : dif[q];
        if (logs) {
            logs[k].nQLiter = nQLiter;
            logs[k].nSorConv = nSorConv;
            logs[k].sor_converged = -1; /* not observable through Ppe's interface */
            logs[k].diverged = difmax > 1.e12;
            for (int q = 0; q < 4; ++q) logs[k].dif[q] = dif[q];
        }
        if (difmax > 1.e12) { fprintf(stderr, "* Solution diverged. Please reduce CFL number.\n"); return 2; } /* :969-972 */
        if (pt) {                                                      /* :1000-1024 (averages go to the starred arrays) */
            const wolfd2_traject *tr = pt->tr;
            orc_velavg_(&nx, &ny, reg->nReg, reg->nRegBrd, reg->nRegType, u, v, us, vs);
            orc_velavg_(&nx, &ny, reg->nReg, reg->nRegBrd, reg->nRegType, un, vn, usn, vsn);
            orc_ptdavg_(&nx, &ny, reg->nReg, reg->nRegBrd, reg->nRegType, d, ts);
            orc_ptdavg_(&nx, &ny, reg->nReg, reg->nRegBrd, reg->nRegType, dn, tsn);
            orc_traject_(&nx, &ny, &tr->ntr, &tr->ntsubstp, &tr->nTrMethod, &tr->nTrCdEq, &tr->mTrHTmit, pt->nTOutBnd,
                         &par->dk, &tr->densref, &par->fr, &tr->dTrHTtol, &tr->dTrHTdel, pt->cpartx, pt->cparty, pt->repc,
                         pt->gx, pt->gy, us, vs, usn, vsn, ts, tsn, pt->xp, pt->yp, pt->up, pt->vp);
            if (orc_errflag) return 1;
        }
    }
    return 0;
}

int32_t orc_step_thermal(const wolfd2_params *par, const wolfd2_regions *reg, const wolfd2_metrics *m,
                         const wolfd2_thermal *th, double *u, double *v, double *p, double *t, double *d,
                         int32_t nsteps, wolfd2_step_log *logs) {
    return orc_step_full(par, reg, m, th, NULL, NULL, u, v, p, t, d, NULL, NULL, NULL, NULL, nsteps, logs);
}

int32_t orc_step(const wolfd2_params *par, const wolfd2_regions *reg, const wolfd2_metrics *m,
                 double *u, double *v, double *p, double *t, double *d,
                 int32_t nsteps, wolfd2_step_log *logs) {
    return orc_step_thermal(par, reg, m, NULL, u, v, p, t, d, nsteps, logs);
}
