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
#include "w2_tri.cuh"

#define F(a, i, j) a[IDX(i, j)]

__device__ __forceinline__ double a_sum4_sic(double a, double b, double c, double d) { return a + b + c + d / 4.0; }

enum { NP_COMPUTE = 0, NP_SAME = 1, NP_CACHED = 2 };   // see mom_row (w2_mom_rows.inc)

// The metric arrays the momentum (and thermal) rows read, with the class each has on a Cartesian grid (see MET in
// w2_mom_rows.inc): 0 = two-dimensional, 1 = function of i only, 2 = function of j only, 3 = constant.
#define W2_MET_LIST(X)                                                                                               \
    X(rau, 0) X(rbu, 3) X(rbv, 3) X(rgv, 0) X(ran, 0) X(rbn, 3) X(rgn, 0) X(rac, 0) X(rbc, 3) X(rgc, 0)              \
    X(dju, 0) X(djv, 0) X(djc, 0) X(xen, 3) X(yen, 2) X(xzn, 1) X(yzn, 3) X(xec, 3) X(yec, 2) X(xzc, 1) X(yzc, 3)    \
    X(xeu, 3) X(yeu, 2) X(xzv, 1) X(yzv, 3) X(xzu, 1) X(yzu, 3) X(xev, 3) X(yev, 2)
struct MomConst {   // values of the constant-class arrays (any interior entry), used by the Cartesian variant only
#define W2_MC_FIELD(name, CLS) double name;
    W2_MET_LIST(W2_MC_FIELD)
#undef W2_MC_FIELD
};

struct MomArgs {
    int nx, ny, pitch;
    double dk, re, fr;
    const double *us, *vs, *un, *vn, *d, *dn;
    // x-momentum metrics
    const double *rbn, *rgn, *rac, *rbc, *dju, *xec, *yec, *xzn, *yzn, *xeu, *yeu, *xzu, *yzu;
    // y-momentum metrics
    const double *ran, *rgc, *djv, *xen, *yen, *xzc, *yzc, *xev, *yev, *xzv, *yzv;
    int iref, jref;     // Cartesian variant: the column / row the j-only / i-only arrays are read from (any interior one held)
    MomConst cc;        // Cartesian variant: the constant-class arrays
    const unsigned char *xmask, *ymask;
    const double *x1;  // first-step solution, field layout
    double *np_c, *np_d;   // cache of cnvn, difn of the component being solved (null: not kept)
    // thermal energy equation (COMP 2, thermal.f:24-272): time-level-n temperature, heat source, fixed-T mask
    const double *tn, *heat, *rau, *rbu, *rbv, *rgv, *djc;
    const unsigned char *tmask;
    double pe;
    // porous regions (RM_POROUS): per-cell region maps, see por_map_kernel
    int porous;
    const W2Regions *R;
    const unsigned char *xd1, *xd2, *yd1, *yd2, *xcp, *ycp;
    int buoy;          // 0: d and dn are +0 everywhere (never written since allocation): the buoyancy term is +0 and b - (+0) == b bit for bit
    const int *done;   // device flag of the QL loop (null outside it): set = converged, later launches do nothing
};

namespace mom_np { constexpr bool kPorous = false; constexpr bool kCart = false; constexpr bool kInterior = false;
#include "w2_mom_rows.inc"
}
namespace mom_po { constexpr bool kPorous = true; constexpr bool kCart = false; constexpr bool kInterior = false;
#include "w2_mom_rows.inc"
}
namespace mom_ca { constexpr bool kPorous = false; constexpr bool kCart = true; constexpr bool kInterior = false;   // Cartesian grid: see MET
#include "w2_mom_rows.inc"
}
// the same rows for unknowns away from the boundary (see kInterior in w2_mom_rows.inc): no ring tests, no index clamps
namespace mom_ni { constexpr bool kPorous = false; constexpr bool kCart = false; constexpr bool kInterior = true;
#include "w2_mom_rows.inc"
}
namespace mom_ci { constexpr bool kPorous = false; constexpr bool kCart = true; constexpr bool kInterior = true;
#include "w2_mom_rows.inc"
}

// Chain geometry (0-based index e, reference ordering momentum.f:353 / :677, thermal.f:155):
// COMP 0 = u: i=1..nx, j=2..ny;  COMP 1 = v: i=2..nx, j=1..ny;  COMP 2 = t: i=2..nx, j=2..ny
#define CH_W(COMP, nx) ((COMP) == 0 ? (nx) : (nx) - 1)   /* unknowns per grid row */
#define CH_I0(COMP) ((COMP) == 0 ? 1 : 2)
#define CH_J0(COMP) ((COMP) == 1 ? 1 : 2)
template <int COMP>
__device__ __forceinline__ void chain_ij(const MomArgs &m, long long e, int &i, int &j) {
    const int w = CH_W(COMP, m.nx);
    const int jj = (int)(e / w);
    j = CH_J0(COMP) + jj; i = CH_I0(COMP) + (int)(e - (long long)jj * w);
}

// Walking a segment in the coalesced mapping (element t, t+TRI_T, ...): one 64-bit division per thread for
// the segment base, then add-and-wrap.  (chain_ij per element costs two 64-bit divisions per unknown, which
// was ~25 % of the kernel's instructions.)
template <int COMP>
struct ChainWalk {
    int w, j, pos;   // row length, current row, 0-based position in the row
    __device__ __forceinline__ ChainWalk(const MomArgs &m, long long e) {
        w = CH_W(COMP, m.nx);
        const long long jj = e / w;
        j = CH_J0(COMP) + (int)jj;
        pos = (int)(e - jj * w);
    }
    __device__ __forceinline__ int i() const { return CH_I0(COMP) + pos; }
    __device__ __forceinline__ void advance(int d) {       // d of the order of w or less
        pos += d;
        while (pos >= w) { pos -= w; ++j; }
    }
    // advance by a fixed stride without a loop (a loop is a branch region, and ptxas moves no load across one):
    // stride = sq*w + sr with sr < w, so one conditional wrap, compiled to selects
    __device__ __forceinline__ void advance_split(int sq, int sr) {
        pos += sr; j += sq;
        const bool wrap = pos >= w;
        pos = wrap ? pos - w : pos; j = wrap ? j + 1 : j;
    }
    __device__ __forceinline__ void advance_far(int d) {   // any d >= 0 (one 32-bit division)
        pos += d;
        const int k = pos / w;
        pos -= k * w; j += k;
    }
};

// ---- fused assembly + level-0 reduce --------------------------------------------------------------------
// One CTA per segment of TRI_S chain unknowns.  Rows are assembled in a coalesced mapping (thread <->
// consecutive i), transposed through shared memory to the solver's chunk mapping (thread <-> 8
// consecutive unknowns; padded stride 9 is conflict-free), reduced by tri_cta_core, and the result is
// transposed back: Y goes straight into the FIELD layout of `out` (no chain->field scatter pass), the
// spikes V/W go to chain-layout arrays only over the prefix/suffix where they are not exactly zero.
#define MR_PAD(x) ((x) + ((x) >> 3))
#define MR_LEN (TRI_S + TRI_S / 8)
// Resident CTAs per SM and unroll depth, swept on B200 after phase A became straight-line code (momentum ms/step at
// 4096^2, Q=2): MINB/U1/U2 = 5/2/4 5.64, 4/2/4 5.15, 4/2/2 5.11, 4/1/4 5.33, 4/4/4 5.33, 4/2/8 5.71, 3/2/4 5.66, 6/2/4 7.08
// (80 registers: spills).  128 registers per thread hold more loads in flight than a fifth CTA hides.
// (round 2b, after the interior-row variant below took ~110 instructions and 12 loads per unknown out of phase A:
// MINB 4 (128 registers) 4.13-4.17 ms, 5 (96 registers, no spills on the hot variants) 3.94-3.98, 6 (80 registers) 4.46)
#ifndef MOM_MINB
#define MOM_MINB 5
#endif
// (round 2, after the cached explicit terms and the Cartesian metric variant: U1/U2 = 2/2 4.375 ms, 1/2 4.308, 4/2 4.58,
// 2/4 4.43, 4/4 4.63, MINB 5 (96 registers) 4.45)
#ifndef MOM_U1
#define MOM_U1 1   /* unknowns of the first split step assembled per unrolled loop body */
#endif
#ifndef MOM_U2
#define MOM_U2 2   /* same, second split step */
#endif
static constexpr int kMomU1 = MOM_U1, kMomU2 = MOM_U2;   // (#pragma unroll does not expand macros)

// VAR: 0 = general metrics, 1 = porous deck (general metrics), 2 = Cartesian grid (mom_ca: see MET in w2_mom_rows.inc)
template <int COMP, int STEP, int VAR, int NP>
__global__ void __launch_bounds__(TRI_T, MOM_MINB) mom_reduce_kernel(MomArgs m, long long n, double *__restrict__ out,
                                                              double *__restrict__ Vg, double *__restrict__ Wg,
                                                              double *__restrict__ seg, int *__restrict__ ext,
                                                              long long nseg, int direct, long long seg0) {
    extern __shared__ __align__(16) double sm[];
    if (m.done != nullptr && *m.done) return;
    double *s0 = sm, *s1 = sm + MR_LEN, *s2 = sm + 2 * MR_LEN, *s3 = sm + 3 * MR_LEN;
    __shared__ int s_ext[2];
    __shared__ double sSig;
    const int t = threadIdx.x;
    const long long g = seg0 + blockIdx.x;   // global segment; Vg, Wg, ext are pre-shifted by the caller
    const long long ebase = g * (long long)TRI_S;
    const int pitch = m.pitch;
    constexpr int L = TRI_M - 2;
    if (t == 0) { s_ext[0] = 0; s_ext[1] = 0; }

    // ---- phase A: assemble rows, coalesced (the cheap second-step rows are unrolled for more loads in flight).
    // The loop body is straight-line code: a branch region (chain end, first-row quirk) would stop ptxas from
    // issuing the loads of the next unrolled unknown before the current one is finished.  Elements past the end
    // of the chain (last segment only) assemble the row of the last valid point and are then replaced by the
    // identity; the first-row quirk is patched in after the loop.
    const long long e0 = ebase + t;
    ChainWalk<COMP> wa(m, e0 < n ? e0 : n - 1);
    int ci = wa.i(), cj = wa.j;
    const int sq = TRI_T / wa.w, sr = TRI_T - sq * wa.w;
    constexpr int UB = (STEP == 2 ? kMomU2 : kMomU1);   // unknowns per loop body
    static_assert(TRI_M % UB == 0, "unroll depth must divide the chunk length");
#pragma unroll 1
    for (int q0 = 0; q0 < TRI_M; q0 += UB) {
        int ui[UB], uj[UB];
        bool inner = true;
#pragma unroll
        for (int k = 0; k < UB; ++k) {
            const long long e = ebase + t + TRI_T * (q0 + k);
            const bool live = e < n;
            if (q0 + k > 0) {
                wa.advance_split(sq, sr);
                ci = live ? wa.i() : ci; cj = live ? wa.j : cj;
            }
            ui[k] = ci; uj[k] = cj;
            inner = inner && ci >= 2 && ci <= m.nx - 1 && cj >= 2 && cj <= m.ny - 1;
        }
        // Warps whose unknowns all lie away from the boundary (all but those at the ends of a grid row and in the first /
        // last row) take the rows without ring tests and index clamps (mom_ni / mom_ci: the same numbers).  The branch is
        // warp-uniform and encloses the whole body, so inside either side the loads of the UB unknowns still issue together.
        double a1[UB], a2[UB], a3[UB], b[UB];
        if (VAR != 1 && __all_sync(0xffffffffu, inner)) {
#pragma unroll
            for (int k = 0; k < UB; ++k) {
                if (VAR == 2) mom_ci::mom_row<COMP, STEP, NP>(m, ui[k], uj[k], a1[k], a2[k], a3[k], b[k]);
                else mom_ni::mom_row<COMP, STEP, NP>(m, ui[k], uj[k], a1[k], a2[k], a3[k], b[k]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < UB; ++k) {
                if (VAR == 1) mom_po::mom_row<COMP, STEP, NP>(m, ui[k], uj[k], a1[k], a2[k], a3[k], b[k]);
                else if (VAR == 2) mom_ca::mom_row<COMP, STEP, NP>(m, ui[k], uj[k], a1[k], a2[k], a3[k], b[k]);
                else mom_np::mom_row<COMP, STEP, NP>(m, ui[k], uj[k], a1[k], a2[k], a3[k], b[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < UB; ++k) {
            const int el = t + TRI_T * (q0 + k);
            const long long e = ebase + el;
            const bool live = e < n;
            const double c3 = (e == n - 1) ? 0.0 : a3[k];
            const int p = MR_PAD(el);
            s0[p] = live ? a1[k] : 0.0; s1[p] = live ? a2[k] : 1.0; s2[p] = live ? c3 : 0.0; s3[p] = live ? b[k] : 0.0;
        }
    }
    if (ebase == 0 && t == 0) {   // AltTridLU first row: a(3,1)/a(2,2) (:1319) == plain Thomas with c1*d1/d2
        int i2, j2; double b1, b2, b3, bb;
        chain_ij<COMP>(m, 1, i2, j2);
        if (VAR == 1) mom_po::mom_row<COMP, STEP, NP>(m, i2, j2, b1, b2, b3, bb);
        else if (VAR == 2) mom_ca::mom_row<COMP, STEP, NP>(m, i2, j2, b1, b2, b3, bb);
        else mom_np::mom_row<COMP, STEP, NP>(m, i2, j2, b1, b2, b3, bb);
        s0[0] = 0.0;
        s2[0] = s2[0] * s1[0] / b2;
    }
    __syncthreads();
    // ---- phase B: chunk mapping
    double A[TRI_M], D[TRI_M], C[TRI_M], B[TRI_M];
#pragma unroll
    for (int k = 0; k < TRI_M; ++k) {
        const int p = 9 * t + k;
        A[k] = s0[p]; D[k] = s1[p]; C[k] = s2[p]; B[k] = s3[p];
    }
    __syncthreads();
    double Ye[TRI_M], Ve[TRI_M], We[TRI_M];
    double ar, dr, cr, br;
    // the six TRI_T-long work arrays of the core alias the (now free) staging area
    tri_cta_core(A, D, C, B, Ye, Ve, We, sm, sm + TRI_T, sm + 2 * TRI_T, sm + 3 * TRI_T, sm + 4 * TRI_T, sm + 5 * TRI_T,
                 ar, dr, cr, br);
    if (direct) {
        if (t == TRI_T - 1) sSig = (br - ar * Ye[L]) / (dr - ar * We[L]);
        __syncthreads();
        const double sig = sSig;
#pragma unroll
        for (int k = 0; k < TRI_M; ++k) Ye[k] = Ye[k] - sig * We[k];
    } else {
        // extent of the exactly-non-zero part of the spikes (prefix for V, suffix for W)
        bool nzv = false, nzw = false;
#pragma unroll
        for (int k = 0; k < TRI_M; ++k) { nzv |= (Ve[k] != 0.0); nzw |= (We[k] != 0.0); }
        if (nzv) atomicMax(&s_ext[0], TRI_M * (t + 1));
        if (nzw) atomicMax(&s_ext[1], TRI_S - TRI_M * t);
    }
    __syncthreads();   // core's shared arrays are dead; s_ext final
    // ---- phase C: transpose back and write coalesced (spike entries only where they will be read: el < extV,
    // el >= TRI_S - extW -- a quarter of the segment each at 4096^2)
    const int extV = s_ext[0], extW = s_ext[1];
    const bool putV = !direct && TRI_M * t < extV, putW = !direct && TRI_M * t + TRI_M - 1 >= TRI_S - extW;
#pragma unroll
    for (int k = 0; k < TRI_M; ++k) {
        const int p = 9 * t + k;
        s0[p] = Ye[k];
        if (putV) s1[p] = Ve[k];
        if (putW) s2[p] = We[k];
    }
    __syncthreads();
    ChainWalk<COMP> wc(m, ebase + t);
#pragma unroll 1
    for (int q = 0; q < TRI_M; ++q) {
        const int el = t + TRI_T * q;
        const long long e = ebase + el;
        if (e >= n) break;
        const int i = wc.i(), j = wc.j;
        wc.advance(TRI_T);
        const int p = MR_PAD(el);
        out[IDX(i, j)] = s0[p];
        if (!direct) {
            if (el < extV) Vg[e] = s1[p];
            if (el >= TRI_S - extW) Wg[e] = s2[p];
        }
    }
    if (direct) return;
    if (t == 0) {
        seg[g] = Ye[0]; seg[nseg + g] = Ve[0]; seg[2 * nseg + g] = We[0];
        ext[2 * g] = extV; ext[2 * g + 1] = extW;
    }
    if (t == TRI_T - 1) {
        seg[3 * nseg + g] = Ye[L]; seg[4 * nseg + g] = Ve[L]; seg[5 * nseg + g] = We[L];
        seg[6 * nseg + g] = ar; seg[7 * nseg + g] = dr; seg[8 * nseg + g] = cr; seg[9 * nseg + g] = br;
    }
}

// x = Y - Sg[g-1]*V - Sg[g]*W in field layout, touching only the unknowns whose spike entries are non-zero
// (a few dozen per segment: the spikes of a diagonally dominant system decay fast).  One warp per segment.
template <int COMP>
__global__ void __launch_bounds__(256) mom_finalize_kernel(MomArgs m, long long n, double *__restrict__ out,
                                                           const double *__restrict__ Vg, const double *__restrict__ Wg,
                                                           const double *__restrict__ sig, const int *__restrict__ ext,
                                                           long long seg0, long long seg1) {
    const long long g = seg0 + (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= seg1 || (m.done != nullptr && *m.done)) return;
    const int lane = threadIdx.x & 31;
    const int pitch = m.pitch;
    const int extV = ext[2 * g], extW = ext[2 * g + 1];
    const double sl = g > 0 ? sig[g - 1] : 0.0, sr = sig[g];
    const int lo2 = max(extV, TRI_S - extW);   // start of the W-only part
    const int cnt = extV + (TRI_S - lo2);
    const ChainWalk<COMP> base(m, g * (long long)TRI_S);
    for (int q = lane; q < cnt; q += 32) {
        const int el = q < extV ? q : lo2 + (q - extV);
        const long long e = g * (long long)TRI_S + el;
        if (e >= n) continue;
        ChainWalk<COMP> w = base;
        w.advance_far(el);
        const int i = w.i(), j = w.j;
        const double v = el < extV ? Vg[e] : 0.0;
        const double ww = el >= TRI_S - extW ? Wg[e] : 0.0;
        out[IDX(i, j)] = out[IDX(i, j)] - sl * v - sr * ww;
    }
}

// identity-row masks of the second split step (momentum.f:434-496 for u, :760-821 for v)
__global__ void mom_mask_kernel(const W2Regions *__restrict__ R, int nx, int jlo, int jhi, int pitch,
                                unsigned char *__restrict__ xmask, unsigned char *__restrict__ ymask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = jlo + blockIdx.y;
    if (i > nx + 1 || j > jhi) return;
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

// us,vs <- un,vn on 1..nx+1, 1..ny+1 (:114-119)
__global__ void __launch_bounds__(256) ql_init_kernel(int nx, int jlo, int jhi, int pitch, const double *__restrict__ un,
                                                      const double *__restrict__ vn, double *__restrict__ us,
                                                      double *__restrict__ vs) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx + 1) return;
    for (int j = jlo + blockIdx.y; j <= jhi; j += gridDim.y) {
        us[IDX(i, j)] = un[IDX(i, j)];
        vs[IDX(i, j)] = vn[IDX(i, j)];
    }
}

// us += dus, vs += dvs on 1..nx,1..ny (:171-176) fused with the two DMaxNorm scans (:179-180)
__global__ void __launch_bounds__(256) ql_update_kernel(int nx, int ny, int jlo, int jhi, int seed, int pitch,
                                                        const double *__restrict__ dus,
                                                        const double *__restrict__ dvs, double *__restrict__ us,
                                                        double *__restrict__ vs, unsigned long long *slots, const int *done) {
    __shared__ double red[32];
    if (*done) return;
    double mu = 0.0, mv = 0.0;
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= nx)
        for (int j = jlo + blockIdx.y; j <= jhi; j += gridDim.y) {
            const double du = dus[IDX(i, j)], dv = dvs[IDX(i, j)];
            us[IDX(i, j)] = us[IDX(i, j)] + du;
            vs[IDX(i, j)] = vs[IDX(i, j)] + dv;
            if (i >= 2 && i <= nx - 1 && j >= 2 && j <= ny - 1) { mu = fmax(mu, fabs(du)); mv = fmax(mv, fabs(dv)); }
        }
    if (seed && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {  // DMaxNorm seed |u(5,5)| (utility.f:493)
        mu = fmax(mu, fabs(dus[IDX(5, 5)]));
        mv = fmax(mv, fabs(dvs[IDX(5, 5)]));
    }
    mu = w2_block_max(mu, red);
    mv = w2_block_max(mv, red);
    if (threadIdx.x == 0) { atomicMax(slots + 0, w2_dbits(mu)); atomicMax(slots + 1, w2_dbits(mv)); }
}

// ---- host side -----------------------------------------------------------------------------------
// per-cell porous-region maps (6 planes): xd1,xd2 / yd1,yd2 = porous regions whose division range holds
// the cell (momentum.f:310-319 / :635-644); xcp / ycp = last region whose PorosCoef assignment reaches it
__global__ void por_map_kernel(const W2Regions *__restrict__ R, int nx, int jlo, int jhi, int pitch, size_t plane,
                               unsigned char *__restrict__ maps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = jlo + blockIdx.y;
    if (i > nx + 1 || j > jhi) return;
    unsigned char xd[2] = {0, 0}, yd[2] = {0, 0}, xc = 0, yc = 0;
    int nxd = 0, nyd = 0;
    for (int q = 0; q < R->nreg; ++q) {
        const int iW = R->iW[q], iE = R->iE[q], jS = R->jS[q], jN = R->jN[q];
        const bool por = R->type[q] == W2_RM_POROUS;
        const bool inx = i >= iW && i <= iE && j >= jS + 1 && j <= jN;
        const bool iny = i >= iW + 1 && i <= iE && j >= jS && j <= jN;
        const bool full = i >= iW && i <= iE && j >= jS && j <= jN;
        if (por) {
            if (inx) { if (nxd < 2) xd[nxd] = (unsigned char)(q + 1); ++nxd; xc = (unsigned char)(q + 1); }
            if (iny) { if (nyd < 2) yd[nyd] = (unsigned char)(q + 1); ++nyd; yc = (unsigned char)(q + 1); }
        } else if (full) { xc = (unsigned char)(q + 1); yc = (unsigned char)(q + 1); }
    }
    const size_t o = IDX(i, j);
    maps[o] = xd[0]; maps[plane + o] = xd[1]; maps[2 * plane + o] = yd[0]; maps[3 * plane + o] = yd[1];
    maps[4 * plane + o] = xc; maps[5 * plane + o] = yc;
}

int w2_build_mom_masks(wolfd2_ctx *c) {
    if (c->hreg.has_porous) {
        W2_TRY(w2_alloc_pormap(c));
        dim3 g((c->nx + 2 + 255) / 256, c->rows);
        por_map_kernel<<<g, 256, 0, c->stream>>>(c->dreg, c->nx, c->A0, c->A1, c->pitch, c->nelem, c->pormap);
    }
    dim3 grid((c->nx + 2 + 255) / 256, c->rows);
    mom_mask_kernel<<<grid, 256, 0, c->stream>>>(c->dreg, c->nx, c->A0, c->A1, c->pitch, c->xmask, c->ymask);
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
    m.x1 = c->x1;
    m.np_c = nullptr; m.np_d = nullptr;
    // (+0 needs dk > 0 and 0 < fr < inf; anything else keeps the term)
    m.buoy = c->d_nonzero || !(c->par.dk > 0.0 && c->par.dk < 1.0e300 && c->par.fr > 0.0 && c->par.fr < 1.0e300);
    m.iref = c->cart_iref; m.jref = c->cart_jref;
    if (c->cart_state == 1) memcpy(&m.cc, c->cart_const, sizeof(MomConst)); else memset(&m.cc, 0, sizeof(MomConst));
    m.done = c->ql_active ? (const int *)(c->d_norm + W2_QL_SLOT) : nullptr;
    m.tn = c->fld[W2_F_TN]; m.heat = c->heat_s; m.tmask = c->tmask; m.pe = c->th.pe;
    m.rau = t.rau; m.rbu = t.rbu; m.rbv = t.rbv; m.rgv = t.rgv; m.djc = t.djc;
    m.porous = c->hreg.has_porous; m.R = c->dreg;
    unsigned char *pm = c->pormap;   // planes are nelem bytes apart; the row shift is in the base pointer
    m.xd1 = pm; m.xd2 = pm + c->nelem; m.yd1 = pm + 2 * c->nelem; m.yd2 = pm + 3 * c->nelem;
    m.xcp = pm + 4 * c->nelem; m.ycp = pm + 5 * c->nelem;
}

template <int COMP, int STEP, int VAR, int NP>
static int mom_solve_impl(wolfd2_ctx *c, MomArgs &m, long long n, double *out, const double *sig1, const double **sig_out);

// one split step of component COMP (sig1 / sig_out: separator values, unused hooks)
template <int COMP, int STEP>
static int mom_solve(wolfd2_ctx *c, MomArgs &m, long long n, double *out, int np, const double *sig1, const double **sig_out) {
    if (c->hreg.has_porous) return mom_solve_impl<COMP, STEP, 1, NP_COMPUTE>(c, m, n, out, sig1, sig_out);
    constexpr int S1 = STEP == 1;
    if (c->cart_state == 1) {   // Cartesian grid, verified on the uploaded arrays (w2_mom_cart_prepare)
        if (S1 && np == NP_SAME) return mom_solve_impl<COMP, STEP, 2, (S1 ? NP_SAME : NP_COMPUTE)>(c, m, n, out, sig1, sig_out);
        if (S1 && np == NP_CACHED) return mom_solve_impl<COMP, STEP, 2, (S1 ? NP_CACHED : NP_COMPUTE)>(c, m, n, out, sig1, sig_out);
        return mom_solve_impl<COMP, STEP, 2, NP_COMPUTE>(c, m, n, out, sig1, sig_out);
    }
    if (S1 && np == NP_SAME) return mom_solve_impl<COMP, STEP, 0, (S1 ? NP_SAME : NP_COMPUTE)>(c, m, n, out, sig1, sig_out);
    if (S1 && np == NP_CACHED) return mom_solve_impl<COMP, STEP, 0, (S1 ? NP_CACHED : NP_COMPUTE)>(c, m, n, out, sig1, sig_out);
    return mom_solve_impl<COMP, STEP, 0, NP_COMPUTE>(c, m, n, out, sig1, sig_out);
}

// Chain unknowns [lo, hi) that lie in the rows this rank updates.  The level-0 segments keep their GLOBAL
// boundaries (so every number is the one a single GPU computes); a rank reduces the segments that START in
// its range, and the unknowns of its last segment that fall into the next slab's first rows are shipped
// there after the second split step (mom_tail_exchange).
template <int COMP>
static void chain_range(const wolfd2_ctx *c, long long &lo, long long &hi) {
    const int r0 = c->E0 > CH_J0(COMP) ? c->E0 : CH_J0(COMP);
    const int r1 = c->E1 < c->ny ? c->E1 : c->ny;
    lo = (long long)(r0 - CH_J0(COMP)) * CH_W(COMP, c->nx);
    hi = (long long)(r1 + 1 - CH_J0(COMP)) * CH_W(COMP, c->nx);
}

template <int COMP>
static int chain_pieces(const wolfd2_ctx *c, double *f, long long e0, long long e1, W2Piece *pc, int maxpc) {
    const int w = CH_W(COMP, c->nx);
    int np = 0;
    for (long long e = e0; e < e1;) {
        const long long jj = e / w;
        const int off = (int)(e - jj * w);
        const long long cnt = (e1 - e) < (long long)(w - off) ? (e1 - e) : (long long)(w - off);
        const int j = CH_J0(COMP) + (int)jj, i = CH_I0(COMP) + off;
        if (np >= maxpc) return -1;
        pc[np].p = f + (size_t)i + (size_t)c->pitch * (size_t)j;
        pc[np].n = (size_t)cnt;
        ++np;
        e += cnt;
    }
    return np;
}

template <int COMP>
static int mom_tail_exchange(wolfd2_ctx *c, double *out) {
    if (c->world == 1) return W2_OK;
    long long lo, hi;
    chain_range<COMP>(c, lo, hi);
    const long long s_lo = (lo + TRI_S - 1) / TRI_S, s_hi = (hi + TRI_S - 1) / TRI_S;
    W2Piece up[16], dn[16];
    int nup = 0, ndn = 0;
    if (c->rank + 1 < c->world) nup = chain_pieces<COMP>(c, out, hi, s_hi * TRI_S, up, 16);
    if (c->rank > 0) ndn = chain_pieces<COMP>(c, out, lo, s_lo * TRI_S, dn, 16);
    if (nup < 0 || ndn < 0) { w2_set_error("momentum tail exchange: too many row pieces"); return W2_ERR_BAD_ARG; }
    return w2_send_recv_pieces(c, up, nup, dn, ndn);
}

template <int COMP, int STEP, int VAR, int NP>
static int mom_solve_impl(wolfd2_ctx *c, MomArgs &m, long long n, double *out, const double *sig1, const double **sig_out) {
    static bool attr[W2_MAXDEV] = {};   // the attribute is per device: one flag per device id
    const size_t smem = (size_t)4 * MR_LEN * sizeof(double);
    if (!attr[c->device % W2_MAXDEV]) {
        W2_CUDA(cudaFuncSetAttribute(mom_reduce_kernel<COMP, STEP, VAR, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr[c->device % W2_MAXDEV] = true;
    }
    W2TriWork &w = c->tri;
    const long long nseg = (n + TRI_S - 1) / TRI_S;
    const int direct = nseg == 1;
    long long s_lo = 0, s_hi = nseg;
    if (c->world > 1) {
        long long lo, hi;
        chain_range<COMP>(c, lo, hi);
        s_lo = (lo + TRI_S - 1) / TRI_S;
        if (c->rank + 1 < c->world) s_hi = (hi + TRI_S - 1) / TRI_S;
        if (direct || s_hi <= s_lo || (s_hi - s_lo) * TRI_S > w.cap) {
            w2_set_error("momentum chain cannot be split across %d ranks (segments %lld..%lld)", c->world, s_lo, s_hi);
            return W2_ERR_BAD_ARG;
        }
        W2_CUDA(cudaMemsetAsync(w.lv[0].seg, 0, 10 * (size_t)nseg * sizeof(double), c->stream));
    }
    double *V0 = w.V0 - s_lo * TRI_S, *W0 = w.W0 - s_lo * TRI_S;
    int *ext = w.ext - 2 * s_lo;
    mom_reduce_kernel<COMP, STEP, VAR, NP><<<(unsigned)(s_hi - s_lo), TRI_T, smem, c->stream>>>(
        m, n, out, V0, W0, w.lv[0].seg, ext, nseg, direct, s_lo);
    c->launches[1]++;
    if (sig_out) *sig_out = nullptr;
    if (!direct) {
        // every rank holds the records of its own segments; summing with the zeros of the others is exact
        W2_TRY(w2_allreduce_sum_f64(c, w.lv[0].seg, 10 * (size_t)nseg));
        const double *sigma = nullptr;
        W2_TRY(w2_tri_upper(c, nseg, &sigma, m.done));   // a few thousand unknowns: solved redundantly on every rank
        // (Applying the first step's correction inside the second step's assembly instead was measured: the extra
        // predicated loads cost that kernel 88 us at 4096^2, the finalize launch it saves 50 us.)
        mom_finalize_kernel<COMP><<<(unsigned)((s_hi - s_lo + 7) / 8), 256, 0, c->stream>>>(m, n, out, V0, W0, sigma, ext, s_lo, s_hi);
        c->launches[1]++;
        if (sig_out) *sig_out = sigma;
    }
    W2_CUDA(cudaGetLastError());
    if (STEP == 2) W2_TRY(mom_tail_exchange<COMP>(c, out));
    return W2_OK;
}

// Arrays that are free during the momentum solve hold the cache of cnvn / difn (NP modes of mom_row): the
// divergence work array and the PPE right-hand side for u, the two colour-split pressure buffers of the fused SOR
// for v.  All four are rewritten from scratch by the PPE that follows (w2_ppe.cu, w2_sor_fused.cu).
static int np_cache(wolfd2_ctx *c, int comp, double **pc, double **pd) {
    if (comp == 0) { *pc = c->div; *pd = c->fld[W2_F_B]; return W2_OK; }
    for (int k = 0; k < 2; ++k)
        if (!c->sorf_buf[k]) W2_TRY(w2_alloc_field(c, &c->sorf_buf[k]));
    *pc = c->sorf_buf[0]; *pd = c->sorf_buf[1];
    return W2_OK;
}

// np: NP_COMPUTE / NP_SAME / NP_CACHED (mom_row); keep: fill the cnvn / difn cache for later QL iterations
int w2_xmomentum(wolfd2_ctx *c, double *dus, int np, int keep) {
    MomArgs m;
    fill_args(c, m);
    if (np == NP_CACHED || keep) W2_TRY(np_cache(c, 0, &m.np_c, &m.np_d));
    const long long n = (long long)c->nx * (c->ny - 1);
    const double *sig1 = nullptr;
    W2_TRY((mom_solve<0, 1>(c, m, n, c->x1, np, nullptr, &sig1)));   // :350-389, local part in field layout
    W2_TRY((mom_solve<0, 2>(c, m, n, dus, NP_COMPUTE, sig1, nullptr)));     // :396-510
    return W2_OK;
}

int w2_ymomentum(wolfd2_ctx *c, double *dvs, int np, int keep) {
    MomArgs m;
    fill_args(c, m);
    if (np == NP_CACHED || keep) W2_TRY(np_cache(c, 1, &m.np_c, &m.np_d));
    const long long n = (long long)(c->nx - 1) * c->ny;
    const double *sig1 = nullptr;
    W2_TRY((mom_solve<1, 1>(c, m, n, c->x1, np, nullptr, &sig1)));   // :675-716
    W2_TRY((mom_solve<1, 2>(c, m, n, dvs, NP_COMPUTE, sig1, nullptr)));     // :723-834
    return W2_OK;
}

// ThermEnergy's two split steps (thermal.f:153-270) on the chain of (nx-1)(ny-1) temperature unknowns; the
// increment lands in dts (field layout).  The convective coefficients come from us, vs (first argument pair of
// the reference call, main.f:849) and un, vn; TempBoundCond and the update t += dts are the caller's.
int w2_thermal_solve(wolfd2_ctx *c, double *dts) {
    if (c->world > 1) { w2_set_error("the thermal energy equation is not supported on several GPUs"); return W2_ERR_UNSUPPORTED; }
    MomArgs m;
    fill_args(c, m);
    const long long n = (long long)(c->nx - 1) * (c->ny - 1);
    const double *sig1 = nullptr;
    W2_TRY((mom_solve_impl<2, 1, 0, NP_COMPUTE>(c, m, n, c->x1, nullptr, &sig1)));
    W2_TRY((mom_solve_impl<2, 2, 0, NP_COMPUTE>(c, m, n, dts, sig1, nullptr)));
    return W2_OK;
}

// ---- Cartesian-grid variant: verification of the metric classes (MET in w2_mom_rows.inc) -----------------------------
struct CartCheck { const double *a[32]; int cls[32]; int n; };
// Every array of class 1 / 2 / 3 must be, bit for bit, a function of i only / of j only / one constant on 1..nx x 1..ny
// and +0 on the ring i = 0, nx+1, j = 0, ny+1 (the rows this rank holds).  bad[k] counts the violations of array k.
__global__ void __launch_bounds__(256) mom_cart_check_kernel(CartCheck cc, int nx, int ny, int jlo, int jhi, int iref, int jref,
                                                             int pitch, unsigned int *bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx + 1) return;
    for (int k = 0; k < cc.n; ++k) {
        const int cls = cc.cls[k];
        if (cls == 0) continue;
        const double *a = cc.a[k];
        const long long cref = __double_as_longlong(a[IDX(iref, jref)]);
        unsigned int nb = 0;
        for (int j = jlo + blockIdx.y; j <= jhi; j += gridDim.y) {
            const long long v = __double_as_longlong(a[IDX(i, j)]);
            const bool inner = i >= 1 && i <= nx && j >= 1 && j <= ny;
            long long want;
            if (!inner) want = 0;                                            // +0.0
            else if (cls == 1) want = __double_as_longlong(a[IDX(i, jref)]);
            else if (cls == 2) want = __double_as_longlong(a[IDX(iref, j)]);
            else want = cref;
            nb += v != want;
        }
        if (nb) atomicAdd(bad + k, nb);
    }
}

int g_mom_cart = 1;   // option "mom_cart": 0 = never use the Cartesian-grid variant

// Decide (once per set of uploaded metrics) whether the Cartesian variant may be used: every rank checks the rows it
// holds, the ranks agree on the outcome.  Collective on several GPUs: called at the top of w2_nauxmomentum.
static int mom_cart_prepare(wolfd2_ctx *c) {
    if (c->cart_state != 0) return W2_OK;
    c->cart_state = -1;
    if (!g_mom_cart || c->hreg.has_porous) return W2_OK;
    const int jlo = c->A0, jhi = c->A1;
    const int j1 = jlo > 1 ? jlo : 1, j2 = jhi < c->ny ? jhi : c->ny;
    if (j2 < j1) return W2_OK;
    c->cart_iref = 1; c->cart_jref = j1;
    CartCheck cc;
    memset(&cc, 0, sizeof(cc));
    const W2Metrics &t = c->met;
#define W2_CC_ADD(name, CLS) cc.a[cc.n] = t.name; cc.cls[cc.n] = (CLS); ++cc.n;
    W2_MET_LIST(W2_CC_ADD)
#undef W2_CC_ADD
    unsigned int *bad = (unsigned int *)(c->d_norm + 48);     // 32 counters in the spare norm slots
    W2_CUDA(cudaMemsetAsync(bad, 0, 32 * sizeof(unsigned int), c->stream));
    dim3 g((c->nx + 2 + 255) / 256, (jhi - jlo + 1) < 1024 ? (jhi - jlo + 1) : 1024);
    mom_cart_check_kernel<<<g, 256, 0, c->stream>>>(cc, c->nx, c->ny, jlo, jhi, c->cart_iref, c->cart_jref, c->pitch, bad);
    W2_CUDA(cudaGetLastError());
    unsigned int hbad[32];
    W2_CUDA(cudaMemcpyAsync(hbad, bad, sizeof(hbad), cudaMemcpyDeviceToHost, c->stream));
    W2_CUDA(cudaStreamSynchronize(c->stream));
    unsigned long long any = 0;
    for (int k = 0; k < cc.n; ++k) any |= hbad[k] != 0;
    if (c->world > 1) {   // the same variant on every rank (not needed for the bits, needed for the same kernel sequence timing only)
        unsigned long long *d = c->d_norm + 47;
        W2_CUDA(cudaMemcpyAsync(d, &any, 8, cudaMemcpyHostToDevice, c->stream));
        W2_TRY(w2_allreduce_max_u64(c, d, 1));
        W2_CUDA(cudaMemcpyAsync(&any, d, 8, cudaMemcpyDeviceToHost, c->stream));
        W2_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (any) return W2_OK;
    MomConst mc;
    memset(&mc, 0, sizeof(mc));
#define W2_CC_GET(name, CLS) \
    if ((CLS) == 3) W2_CUDA(cudaMemcpy(&mc.name, t.name + (size_t)c->cart_iref + (size_t)c->pitch * (size_t)c->cart_jref, 8, cudaMemcpyDeviceToHost));
    W2_MET_LIST(W2_CC_GET)
#undef W2_CC_GET
    static_assert(sizeof(MomConst) <= sizeof(c->cart_const), "cart_const too small");
    memcpy(c->cart_const, &mc, sizeof(mc));
    c->cart_state = 1;
    return W2_OK;
}

int g_mom_two_streams = 1;   // option "mom_two_streams": 0 = XMomentum and YMomentum one after the other on one stream

// Second stream and second set of work arrays for YMomentum (one GPU).  If the extra memory (three field-sized arrays) is
// not there, the solve stays on one stream.
static bool mom2_ready(wolfd2_ctx *c) {
    if (!g_mom_two_streams || c->world > 1 || c->mom2_state < 0) return false;
    if (c->mom2_state == 1) return true;
    c->mom2_state = -1;
    bool ok = cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking) == cudaSuccess
              && cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) == cudaSuccess
              && cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && w2_alloc_field(c, &c->x1b) == W2_OK;
    ok = ok && w2_tri_prepare_second(c) == W2_OK;
    if (!ok) { cudaGetLastError(); return false; }   // (whatever was allocated is released with the context)
    c->mom2_state = 1;
    return true;
}
// swap the stream and the work arrays the momentum launchers use (c->stream, c->x1, c->tri) with the second set
static void mom2_swap(wolfd2_ctx *c) {
    cudaStream_t s = c->stream; c->stream = c->stream2; c->stream2 = s;
    double *x = c->x1; c->x1 = c->x1b; c->x1b = x;
    W2TriWork w = c->tri; c->tri = c->tri2; c->tri2 = w;
}

int g_mom_np_cache = 1;   // option "mom_np_cache": 0 = always evaluate cnvn / difn from un, vn (tests compare both ways)

__global__ void ql_ctl_reset(int *ctl) { ctl[0] = 0; ctl[1] = -1; }

// End of QL iteration m: `difmax.le.qtol` (:182-188) on the device.  ctl[0] = converged flag, ctl[1] = nQLiter.
__global__ void ql_decide_kernel(int *ctl, const unsigned long long *slots, double qtol, int m) {
    if (ctl[0]) return;
    const double du = __longlong_as_double((long long)slots[0]), dv = __longlong_as_double((long long)slots[1]);
    const double difmax = du > dv ? du : dv;
    if (difmax <= qtol) { ctl[0] = 1; ctl[1] = m; }
}

// nAuxMomentum (:33-193) with the loop control on the device.  Every kernel of an iteration returns at once when the
// convergence flag is set, so the host enqueues iterations WITHOUT waiting for their outcome: iteration k is enqueued
// as soon as the outcome of iteration k-2 is known (one speculative iteration at most, its launches cost a few
// microseconds each), and the GPU never idles on a host round trip.  With max_ql_iter <= 2 (the fixed-work mode) there
// is no wait at all; nQLiter travels back with the step's norms.  On several GPUs the decision is taken from the
// all-reduced norms, so every rank enqueues the same sequence (the NCCL calls of a speculative iteration move stale
// data and change nothing).
int w2_nauxmomentum(wolfd2_ctx *c, int init_star, int *nQLiter) {
    const int nx = c->nx, ny = c->ny;
    double *us = c->fld[W2_F_US], *vs = c->fld[W2_F_VS];
    int *ctl = (int *)(c->d_norm + W2_QL_SLOT);
    int *hctl = (int *)(c->h_norm + W2_QL_SLOT);     // pinned: [2*k], [2*k+1] = flag, count after iteration k (k = 0: final)
    if (nQLiter) *nQLiter = -1;  // :111
    W2_TRY(mom_cart_prepare(c));
    if (!c->ev_ql[0]) for (int k = 0; k < 2; ++k) W2_CUDA(cudaEventCreateWithFlags(&c->ev_ql[k], cudaEventDisableTiming));
    ql_ctl_reset<<<1, 1, 0, c->stream>>>(ctl);
    if (init_star) {
        // all held rows: un, vn have valid halos, so the copy needs no exchange
        const int jlo = c->A0 > 1 ? c->A0 : 1, jhi = c->A1;
        dim3 g((nx + 1 + 255) / 256, (jhi - jlo + 1) < 2048 ? (jhi - jlo + 1) : 2048);
        ql_init_kernel<<<g, 256, 0, c->stream>>>(nx, jlo, jhi, c->pitch, c->fld[W2_F_UN], c->fld[W2_F_VN], us, vs);
        c->launches[1]++;
    }
    // Time-level-n halves of the explicit terms (mom_row, NP modes).  On the first QL iteration us, vs are bitwise
    // copies of un, vn on every cell (main.f:696-747) unless nAuxMomentum's own initialisation loop ran (it skips row and
    // column 0, :114-119) or VelOutflowBCs (:133) has just rewritten the outlet ghost cells of us, vs.
    bool has_outlet = false;
    for (int q = 0; q < c->hreg.nreg; ++q)
        for (int k = 0; k < 4; ++k) has_outlet |= c->hreg.bd[q][k] == W2_BM_OUTLT1 || c->hreg.bd[q][k] == W2_BM_OUTLT2;
    const bool cache_ok = !c->hreg.has_porous && g_mom_np_cache;
    const int keep = cache_ok && c->par.mqiter > 1;
    c->ql_active = 1;
    int rc = W2_OK;
    for (int m = 1; m <= c->par.mqiter && rc == W2_OK; ++m) {
        if (m >= 3) {   // outcome of iteration m-2 (iteration m-1 is in flight)
            W2_CUDA(cudaEventSynchronize(c->ev_ql[m & 1]));
            c->host_syncs++;
            if (hctl[2 + 2 * (m & 1)]) break;
        }
        rc = w2_outflow_bc(c, us, vs, ctl);  // :133
        const int np = !cache_ok ? NP_COMPUTE : m > 1 ? NP_CACHED : (!init_star && !has_outlet) ? NP_SAME : NP_COMPUTE;
        // dus, dvs are zero outside the ranges XMomentum/YMomentum write (:139-144 re-zeroes the same cells)
        const bool two = rc == W2_OK && mom2_ready(c);
        if (two) {   // fork: the second stream starts where this one stands (us, vs of this iteration are final)
            cudaEventRecord(c->ev_fork, c->stream);
            cudaStreamWaitEvent(c->stream2, c->ev_fork, 0);
        }
        if (rc == W2_OK) rc = w2_xmomentum(c, c->dus, np, keep);   // :147
        if (two) mom2_swap(c);
        if (rc == W2_OK) rc = w2_ymomentum(c, c->dvs, np, keep);   // :158
        if (two) {   // join: the update below needs dus and dvs
            cudaEventRecord(c->ev_join, c->stream);
            mom2_swap(c);
            cudaStreamWaitEvent(c->stream, c->ev_join, 0);
        }
        if (rc != W2_OK) break;
        cudaMemsetAsync(c->d_norm + 8, 0, 2 * sizeof(unsigned long long), c->stream);
        int jlo = 1, jhi = ny;
        w2_clip(c, jlo, jhi);
        dim3 g((nx + 255) / 256, (jhi - jlo + 1) < 2048 ? (jhi - jlo + 1) : 2048);
        ql_update_kernel<<<g, 256, 0, c->stream>>>(nx, ny, jlo, jhi, c->rank == 0, c->pitch, c->dus, c->dvs, us, vs, c->d_norm + 8, ctl);
        c->launches[1]++;
        if (c->world > 1) {
            rc = w2_allreduce_max_u64(c, c->d_norm + 8, 2);
            double *uv[2] = {us, vs};
            if (rc == W2_OK) rc = w2_halo_exchange(c, uv, 2, c->HG);
            if (rc != W2_OK) break;
        }
        ql_decide_kernel<<<1, 1, 0, c->stream>>>(ctl, c->d_norm + 8, c->par.qtol, m);   // :182-188
        c->launches[1]++;
        if (m + 2 <= c->par.mqiter) {   // somebody will ask for this outcome
            cudaMemcpyAsync(hctl + 2 + 2 * (m & 1), ctl, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream);
            cudaEventRecord(c->ev_ql[m & 1], c->stream);
        }
    }
    c->ql_active = 0;
    if (rc != W2_OK) return rc;
    W2_CUDA(cudaGetLastError());
    W2_CUDA(cudaMemcpyAsync(hctl, ctl, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (nQLiter) {
        W2_CUDA(cudaStreamSynchronize(c->stream));
        c->host_syncs++;
        *nQLiter = w2_ql_result(c);
    }
    return W2_OK;
}

// nQLiter of the last w2_nauxmomentum call; valid once the stream has been synchronised after it
int w2_ql_result(wolfd2_ctx *c) {
    const int *hctl = (const int *)(c->h_norm + W2_QL_SLOT);
    return hctl[0] ? hctl[1] : -1;
}

// ---- unit-parity entry points ("internal but worth exporting") -------------------------
// ConvCoef, DConvU/V, DDiffU/V and PorosCoef are never materialised on the production path: mom_row evaluates
// them per unknown.  These kernels run THE SAME device functions (x_c1_raw ... por_coef of w2_mom_rows.inc) one
// operator at a time over the reference's loop ranges and store the result, so that each stencil can be compared
// with the oracle bit for bit (convcoef_, dconvu_, ... in w2_step.cu).  Argument combinations the reference
// never uses (e.g. ncomp 1 with njacob 1, commented out at momentum.f:279) go through the written-out formula.
__global__ void __launch_bounds__(256) unit_convcoef_kernel(MomArgs m, int ncomp, int njacob, const double *__restrict__ xzi,
                                                            const double *__restrict__ xet, const double *__restrict__ yzi,
                                                            const double *__restrict__ yet, const double *__restrict__ u,
                                                            const double *__restrict__ v, double *__restrict__ cc1,
                                                            double *__restrict__ cc2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    const int nx = m.nx, ny = m.ny, pitch = m.pitch;
    if (i > nx + 1 || j > ny + 1) return;
    const double djac = njacob == 1 ? 2.0 : 1.0;   // :895-896
    const bool in1 = i >= 1 && j >= 1;              // 1..nx+1, 1..ny+1
    const bool in0 = in1 && i <= nx && j <= ny;     // 1..nx, 1..ny
    using namespace mom_np;
    switch (ncomp) {
    case 1:   // :901-913
        if (in1) cc1[IDX(i, j)] = njacob == 0 ? x_c1_raw(m, u, v, i, j)
                                              : (djac * F(yet, i, j) * (F(u, i, j) + F(u, i - 1, j)) - F(xet, i, j) * (F(v, i, j) + F(v, i, j - 1))) * 0.5;
        if (in0) cc2[IDX(i, j)] = njacob == 0 ? x_c2_raw(m, u, v, i, j)
                                              : (F(xzi, i, j) * (F(v, i + 1, j) + F(v, i, j)) - djac * F(yzi, i, j) * (F(u, i, j + 1) + F(u, i, j))) * 0.5;
        break;
    case 2:   // :916-928
        if (in0) cc1[IDX(i, j)] = njacob == 0 ? y_c1_raw(m, u, v, i, j)
                                              : (F(yet, i, j) * (F(u, i, j + 1) + F(u, i, j)) - djac * F(xet, i, j) * (F(v, i + 1, j) + F(v, i, j))) * 0.5;
        if (in1) cc2[IDX(i, j)] = njacob == 0 ? y_c2_raw(m, u, v, i, j)
                                              : (djac * F(xzi, i, j) * (F(v, i, j) + F(v, i, j - 1)) - F(yzi, i, j) * (F(u, i, j) + F(u, i - 1, j))) * 0.5;
        break;
    case 3:   // :931-939
        if (in0) { cc1[IDX(i, j)] = t_cun(m, u, v, i, j); cc2[IDX(i, j)] = t_cvn(m, u, v, i, j); }
        break;
    case 4:   // :942-950, i = 0..nx, j = 1..ny
        if (i <= nx && j >= 1 && j <= ny) {
            if (njacob == 1) { cc1[IDX(i, j)] = x_cj1_raw(m, i, j); cc2[IDX(i, j)] = x_cj2_raw(m, i, j); }
            else {
                const double s4 = F(v, i + 1, j) + F(v, i, j) + F(v, i + 1, j - 1) + F(v, i, j - 1);
                cc1[IDX(i, j)] = djac * F(yet, i, j) * F(u, i, j) - F(xet, i, j) * s4 / 4.0;
                cc2[IDX(i, j)] = F(xzi, i, j) * s4 / 4.0 - djac * F(yzi, i, j) * F(u, i, j);
            }
        }
        break;
    case 5:   // :953-961, i = 1..nx, j = 0..ny
        if (i >= 1 && i <= nx && j <= ny) {
            if (njacob == 1) { cc1[IDX(i, j)] = y_cj1_raw(m, i, j); cc2[IDX(i, j)] = y_cj2_raw(m, i, j); }
            else {
                const double s4 = F(u, i, j + 1) + F(u, i - 1, j + 1) + F(u, i, j) + F(u, i - 1, j);
                cc1[IDX(i, j)] = F(yet, i, j) * s4 / 4.0 - djac * F(xet, i, j) * F(v, i, j);
                cc2[IDX(i, j)] = djac * F(xzi, i, j) * F(v, i, j) - F(yzi, i, j) * s4 / 4.0;
            }
        }
        break;
    case 6:   // :964-972
        if (in1) { cc1[IDX(i, j)] = t_cu1(m, u, v, i, j); cc2[IDX(i, j)] = t_cv1(m, u, v, i, j); }
        break;
    }
}

// DConvU (:1000-1006) / DConvV (:1064-1070) on given coefficient arrays; DDiffU (:1028-1042) / DDiffV (:1092-1106)
template <int COMP>
__global__ void __launch_bounds__(256) unit_dconv_kernel(int nx, int ny, int pitch, const double *__restrict__ c1,
                                                         const double *__restrict__ c2, const double *__restrict__ q,
                                                         double *__restrict__ out) {
    const int i = CH_I0(COMP) + blockIdx.x * blockDim.x + threadIdx.x, j = CH_J0(COMP) + blockIdx.y;
    if (i > nx || j > ny) return;
    if (COMP == 0) out[IDX(i, j)] = mom_np::x_conv_sum(F(c1, i, j), F(c1, i + 1, j), F(c2, i, j), F(c2, i, j - 1), q, pitch, i, j);
    else out[IDX(i, j)] = mom_np::y_conv_sum(F(c1, i, j), F(c1, i - 1, j), F(c2, i, j), F(c2, i, j + 1), q, pitch, i, j);
}
template <int COMP>
__global__ void __launch_bounds__(256) unit_ddiff_kernel(MomArgs m, const double *__restrict__ q, double *__restrict__ out) {
    const int i = CH_I0(COMP) + blockIdx.x * blockDim.x + threadIdx.x, j = CH_J0(COMP) + blockIdx.y;
    const int pitch = m.pitch;
    if (i > m.nx || j > m.ny) return;
    out[IDX(i, j)] = COMP == 0 ? mom_np::x_diff(m, q, i, j) : mom_np::y_diff(m, q, i, j);
}
// PorosCoef (:1115-1226): cells no region assignment reaches keep their value
template <int COMP>
__global__ void __launch_bounds__(256) unit_poroscoef_kernel(MomArgs m, int njacob, const double *__restrict__ u,
                                                             const double *__restrict__ v, double *__restrict__ cp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    const int pitch = m.pitch;
    if (i > m.nx + 1 || j > m.ny + 1) return;
    if ((COMP == 0 ? m.xcp : m.ycp)[IDX(i, j)] == 0) return;
    cp[IDX(i, j)] = mom_po::por_coef<COMP>(m, u, v, njacob, i, j);
}

int w2_unit_convcoef(wolfd2_ctx *c, int ncomp, int njacob, const double *xzi, const double *xet, const double *yzi,
                     const double *yet, const double *u, const double *v, double *cc1, double *cc2) {
    if (ncomp < 1 || ncomp > 6) {
        w2_set_error("Error: Wrong component flag passed to ConvCoef: %d", ncomp);   // :975-977
        return W2_ERR_BAD_ARG;
    }
    MomArgs m;
    fill_args(c, m);
    // the metric names mom_row's operators read, bound to this call's arguments (call sites: momentum.f:282-290,
    // :607-615, thermal.f:111-116)
    switch (ncomp) {
    case 1: m.xzn = xzi; m.xec = xet; m.yzn = yzi; m.yec = yet; break;
    case 2: m.xzc = xzi; m.xen = xet; m.yzc = yzi; m.yen = yet; break;
    case 3: m.xzv = xzi; m.xeu = xet; m.yzv = yzi; m.yeu = yet; break;
    case 4: m.xzu = xzi; m.xeu = xet; m.yzu = yzi; m.yeu = yet; m.us = u; m.vs = v; break;
    case 5: m.xzv = xzi; m.xev = xet; m.yzv = yzi; m.yev = yet; m.us = u; m.vs = v; break;
    case 6: m.xzc = xzi; m.xec = xet; m.yzc = yzi; m.yec = yet; break;
    }
    dim3 g((c->nx + 2 + 255) / 256, c->ny + 2);
    unit_convcoef_kernel<<<g, 256, 0, c->stream>>>(m, ncomp, njacob, xzi, xet, yzi, yet, u, v, cc1, cc2);
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}
int w2_unit_dconv(wolfd2_ctx *c, int comp, const double *c1, const double *c2, const double *q, double *out) {
    const int w = comp == 0 ? c->nx : c->nx - 1, h = comp == 0 ? c->ny - 1 : c->ny;
    dim3 g((w + 255) / 256, h);
    if (comp == 0) unit_dconv_kernel<0><<<g, 256, 0, c->stream>>>(c->nx, c->ny, c->pitch, c1, c2, q, out);
    else unit_dconv_kernel<1><<<g, 256, 0, c->stream>>>(c->nx, c->ny, c->pitch, c1, c2, q, out);
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}
// comp 0: (ac, bc, bn, gn) of DDiffU; comp 1: (an, bc, bn, gc) of DDiffV
int w2_unit_ddiff(wolfd2_ctx *c, int comp, const double *a, const double *bc, const double *bn, const double *g_, const double *q,
                  double *out) {
    MomArgs m;
    fill_args(c, m);
    m.rbc = bc; m.rbn = bn;
    if (comp == 0) { m.rac = a; m.rgn = g_; } else { m.ran = a; m.rgc = g_; }
    const int w = comp == 0 ? c->nx : c->nx - 1, h = comp == 0 ? c->ny - 1 : c->ny;
    dim3 g((w + 255) / 256, h);
    if (comp == 0) unit_ddiff_kernel<0><<<g, 256, 0, c->stream>>>(m, q, out);
    else unit_ddiff_kernel<1><<<g, 256, 0, c->stream>>>(m, q, out);
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}
int w2_unit_poroscoef(wolfd2_ctx *c, int ncomp, int njacob, const double *u, const double *v, double *cp) {
    if (ncomp != 1 && ncomp != 2) {
        w2_set_error("Error: Wrong ncomp flag passed to PorosCoef: %d", ncomp);   // :1216-1218
        return W2_ERR_BAD_ARG;
    }
    // the last-assignment maps are built with the momentum masks when the deck has a porous region; without one
    // every region assigns zero on its full index range (:1166-1172)
    if (!c->hreg.has_porous) {
        W2_TRY(w2_alloc_pormap(c));
        dim3 gm((c->nx + 2 + 255) / 256, c->rows);
        por_map_kernel<<<gm, 256, 0, c->stream>>>(c->dreg, c->nx, c->A0, c->A1, c->pitch, c->nelem, c->pormap);
    }
    MomArgs m;
    fill_args(c, m);
    dim3 g((c->nx + 2 + 255) / 256, c->ny + 2);
    if (ncomp == 1) unit_poroscoef_kernel<0><<<g, 256, 0, c->stream>>>(m, njacob, u, v, cp);
    else unit_poroscoef_kernel<1><<<g, 256, 0, c->stream>>>(m, njacob, u, v, cp);
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}
