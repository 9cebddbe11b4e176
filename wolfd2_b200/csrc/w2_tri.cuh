// w2_tri.cuh -- per-CTA core of the partitioned tridiagonal solver (see w2_tridiag.cu for the method).
#pragma once
#include "w2.cuh"

#define TRI_T 128
#define TRI_M 8
#define TRI_S (TRI_T * TRI_M)

// 1/x for the pivots of the elimination: hardware seed (rcp.approx.ftz.f64, 2^-23) + two Newton steps in
// FMA arithmetic, ~1 ulp, 5 instructions instead of the ~25 of the IEEE division sequence with its
// special-case path.  Pivots here are O(1) by diagonal dominance (never 0, inf, NaN or subnormal).  The
// parallel elimination order already differs from AltTridLU's, so nothing here is meant to be bit-exact
// (tolerances: DESIGN.md section 5); the point-wise SOR and stencil kernels keep true divisions.
__device__ __forceinline__ double tri_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = __fma_rn(-x, y, 1.0);
    y = __fma_rn(y, e, y);
    e = __fma_rn(-x, y, 1.0);
    return __fma_rn(y, e, y);
}

// a*b + c of the elimination.  The library is built with -fmad=false so that the stencils keep the reference's rounding;
// the elimination order of this solver differs from AltTridLU's anyway (tolerances: DESIGN.md section 5), so its
// multiply-adds are fused explicitly: one rounding instead of two, and a third fewer fp64 instructions in the core.
// (-DTRI_FMA=0 restores the unfused arithmetic of rounds 1 and 2a.)
#ifndef TRI_FMA
#define TRI_FMA 1
#endif
__device__ __forceinline__ double tri_fma(double a, double b, double c) {
#if TRI_FMA
    return __fma_rn(a, b, c);
#else
    return a * b + c;
#endif
}

// Input: this thread's TRI_M consecutive rows (A,D,C,B).  Output: for each of them the coefficients of
//   x = Ye - Sg[g-1]*Ve - Sg[g]*We
// where Sg[g] is the CTA's separator (its last unknown) and Sg[g-1] that of the previous segment, plus
// the separator row (ar,dr,cr,br) of this thread.  sA..sW are six shared arrays of TRI_T doubles.
// Every thread of the CTA must call this (it contains block barriers).
__device__ __forceinline__ void tri_cta_core(const double (&A)[TRI_M], const double (&D)[TRI_M],
                                             const double (&C)[TRI_M], const double (&B)[TRI_M],
                                             double (&Ye)[TRI_M], double (&Ve)[TRI_M], double (&We)[TRI_M],
                                             double *sA, double *sD, double *sC, double *sY, double *sV, double *sW,
                                             double &ar, double &dr, double &cr, double &br) {
    const int t = threadIdx.x;
    constexpr int L = TRI_M - 2;  // last interior index
    double y[TRI_M - 1], v[TRI_M - 1], w[TRI_M - 1], cp[TRI_M - 1];
    // --- thread-level elimination of the M-1 interior unknowns, 3 right-hand sides
    {
        double inv = tri_rcp(D[0]);
        cp[0] = C[0] * inv; y[0] = B[0] * inv; v[0] = A[0] * inv;
#pragma unroll
        for (int k = 1; k <= L; ++k) {
            inv = tri_rcp(tri_fma(-A[k], cp[k - 1], D[k]));
            cp[k] = C[k] * inv;
            y[k] = tri_fma(-A[k], y[k - 1], B[k]) * inv;
            v[k] = (-A[k] * v[k - 1]) * inv;
        }
        w[L] = cp[L];
#pragma unroll
        for (int k = L - 1; k >= 0; --k) {
            y[k] = tri_fma(-cp[k], y[k + 1], y[k]);
            v[k] = tri_fma(-cp[k], v[k + 1], v[k]);
            w[k] = -cp[k] * w[k + 1];
        }
    }
    // --- exchange first-interior values with the left neighbour thread
    sY[t] = y[0]; sV[t] = v[0]; sW[t] = w[0];
    __syncthreads();
    ar = A[TRI_M - 1]; dr = D[TRI_M - 1]; cr = C[TRI_M - 1]; br = B[TRI_M - 1];
    double rA, rD, rC, rY, rV, rW;  // this thread's separator row, 3 rhs
    if (t < TRI_T - 1) {
        const double yF = sY[t + 1], vF = sV[t + 1], wF = sW[t + 1];
        rA = -ar * v[L];
        rD = tri_fma(-cr, vF, tri_fma(-ar, w[L], dr));
        rC = -cr * wF;
        rY = tri_fma(-ar, y[L], br);
        rY = tri_fma(-cr, yF, rY);
        rV = 0.0; rW = 0.0;
        if (t == 0) { rV = rA; rA = 0.0; }
        if (t == TRI_T - 2) { rW = rC; rC = 0.0; }
    } else {  // the CTA's own separator is not part of the in-CTA system
        rA = 0.0; rD = 1.0; rC = 0.0; rY = 0.0; rV = 0.0; rW = 0.0;
    }
    __syncthreads();
    // --- parallel cyclic reduction over the T thread separators
#pragma unroll 1
    for (int s = 1; s < TRI_T; s <<= 1) {
        sA[t] = rA; sD[t] = rD; sC[t] = rC; sY[t] = rY; sV[t] = rV; sW[t] = rW;
        __syncthreads();
        const int lo = t - s, hi = t + s;
        double aL = 0.0, dL = 1.0, cL = 0.0, yL = 0.0, vL = 0.0, wL = 0.0;
        double aH = 0.0, dH = 1.0, cH = 0.0, yH = 0.0, vH = 0.0, wH = 0.0;
        if (lo >= 0) { aL = sA[lo]; dL = sD[lo]; cL = sC[lo]; yL = sY[lo]; vL = sV[lo]; wL = sW[lo]; }
        if (hi < TRI_T) { aH = sA[hi]; dH = sD[hi]; cH = sC[hi]; yH = sY[hi]; vH = sV[hi]; wH = sW[hi]; }
        const double al = -rA * tri_rcp(dL), ga = -rC * tri_rcp(dH);
        rD = tri_fma(ga, aH, tri_fma(al, cL, rD));
        rY = tri_fma(ga, yH, tri_fma(al, yL, rY));
        rV = tri_fma(ga, vH, tri_fma(al, vL, rV));
        rW = tri_fma(ga, wH, tri_fma(al, wL, rW));
        rA = al * aL;
        rC = ga * cH;
        __syncthreads();
    }
    // separator solution, as coefficients of (1, Sg[g-1], Sg[g])
    double sy, sv, sw;
    if (t < TRI_T - 1) { const double inv = tri_rcp(rD); sy = rY * inv; sv = rV * inv; sw = rW * inv; }
    else { sy = 0.0; sv = 0.0; sw = -1.0; }
    sY[t] = sy; sV[t] = sv; sW[t] = sw;
    __syncthreads();
    double py = 0.0, pv = -1.0, pw = 0.0;  // the separator to the left of this chunk
    if (t > 0) { py = sY[t - 1]; pv = sV[t - 1]; pw = sW[t - 1]; }

    // --- per-element coefficients of the segment-level representation
#pragma unroll
    for (int k = 0; k <= L; ++k) {
        Ye[k] = tri_fma(-sy, w[k], tri_fma(-py, v[k], y[k]));
        Ve[k] = -tri_fma(sv, w[k], pv * v[k]);
        We[k] = -tri_fma(sw, w[k], pw * v[k]);
    }
    Ye[TRI_M - 1] = sy; Ve[TRI_M - 1] = sv; We[TRI_M - 1] = sw;

}
