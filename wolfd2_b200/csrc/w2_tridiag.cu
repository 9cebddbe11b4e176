// w2_tridiag.cu -- multi-level partitioned solver for ONE long tridiagonal chain.
//
// Replaces the serial Thomas recurrence of AltTridLU (src/momentum.f:1307-1339).  The reference
// solves each momentum split step as a single system of nx*(ny-1) or (nx-1)*ny unknowns in which
// consecutive grid lines stay coupled through the first/last unknown of every line
// (momentum.f:353,389,677,716), so a batch of independent line solves would be wrong.
//
// Method (exact to rounding, no truncation): substructuring with scalar separators.
//   level 0: the chain is cut into segments of S = T*M unknowns, one CTA each.  Thread t owns M
//            consecutive unknowns; the last one is its separator, the other M-1 are eliminated
//            in registers by a 3-right-hand-side Thomas pass (rhs, left spike, right spike).
//            The T thread separators of a CTA form a tridiagonal system solved in shared memory
//            by parallel cyclic reduction, again for 3 right-hand sides, which expresses every
//            unknown of the segment as  x = Y - Sg[g-1]*V - Sg[g]*W  where Sg[g] is the segment's
//            own separator (its last unknown) and Sg[g-1] the previous segment's.
//   level l+1: the segment separators obey a scalar tridiagonal system of size ceil(n/S) whose
//            rows are formed from 10 numbers per segment; it is solved by the same kernel.
//   top:     a level that fits one CTA is solved directly.
//   finalize: x = Y - Sg[g-1]*V - Sg[g]*W, level by level, downwards.
// All chains on this path are strictly diagonally dominant (1 + rkj*(|c| + 2 r) on the
// diagonal), for which block elimination and PCR without pivoting are stable -- the reference
// does not pivot either.
//
// AltTridLU's first-row quirk (momentum.f:1319: a(3,1) = a(3,1)/a(2,2)) is algebraically the
// plain Thomas algorithm applied to a system whose first super-diagonal is c1*d1/d2; the row
// provider applies exactly that substitution.
#include "w2.cuh"

#define TRI_T 512
#define TRI_M 8
#define TRI_S (TRI_T * TRI_M)

// ---------------------------------------------------------------------- row providers
struct ProvSoA {  // rows stored as four arrays in HBM
    const double *a, *d, *c, *b;
    long long n;
    int quirk;
    int line_len;  // 0: one chain; >0: independent lines of this length stored back to back
    __device__ __forceinline__ void load(long long e0, double *A, double *D, double *C, double *B) const {
        const double2 *pa = reinterpret_cast<const double2 *>(a + e0);
        const double2 *pd = reinterpret_cast<const double2 *>(d + e0);
        const double2 *pc = reinterpret_cast<const double2 *>(c + e0);
        const double2 *pb = reinterpret_cast<const double2 *>(b + e0);
#pragma unroll
        for (int q = 0; q < TRI_M / 2; ++q) {
            const double2 va = pa[q], vd = pd[q], vc = pc[q], vb = pb[q];
            A[2 * q] = va.x; A[2 * q + 1] = va.y;
            D[2 * q] = vd.x; D[2 * q + 1] = vd.y;
            C[2 * q] = vc.x; C[2 * q + 1] = vc.y;
            B[2 * q] = vb.x; B[2 * q + 1] = vb.y;
        }
#pragma unroll
        for (int k = 0; k < TRI_M; ++k) {
            const long long idx = e0 + k;
            if (idx >= n) { A[k] = 0.0; D[k] = 1.0; C[k] = 0.0; B[k] = 0.0; continue; }
            bool first, last;
            if (line_len > 0) {
                const long long pos = idx % line_len;
                first = pos == 0; last = pos == line_len - 1;
            } else {
                first = idx == 0; last = idx == n - 1;
            }
            if (first) {
                A[k] = 0.0;
                if (quirk && !last) C[k] = C[k] * D[k] / d[idx + 1];
            }
            if (last) C[k] = 0.0;
        }
    }
};

struct ProvSeg {  // rows of the separator system of the level below, formed from its segment records
    const double *seg;  // 10 arrays of length nseg: YF,VF,WF,YL,VL,WL,ar,dr,cr,br
    long long n;        // = nseg of the level below
    __device__ __forceinline__ void load(long long e0, double *A, double *D, double *C, double *B) const {
#pragma unroll
        for (int k = 0; k < TRI_M; ++k) {
            const long long g = e0 + k;
            if (g >= n) { A[k] = 0.0; D[k] = 1.0; C[k] = 0.0; B[k] = 0.0; continue; }
            const double YL = seg[3 * n + g], VL = seg[4 * n + g], WL = seg[5 * n + g];
            const double ar = seg[6 * n + g], dr = seg[7 * n + g], cr = seg[8 * n + g], br = seg[9 * n + g];
            double YF = 0.0, VF = 0.0, WF = 0.0;
            if (g + 1 < n) { YF = seg[g + 1]; VF = seg[n + g + 1]; WF = seg[2 * n + g + 1]; }
            A[k] = (g == 0) ? 0.0 : -ar * VL;
            D[k] = dr - ar * WL - cr * VF;
            C[k] = (g + 1 < n) ? -cr * WF : 0.0;
            B[k] = br - ar * YL - cr * YF;
        }
    }
};

// ---------------------------------------------------------------------- reduce kernel
template <class Prov>
__global__ void __launch_bounds__(TRI_T, 1) tri_reduce_kernel(Prov prov, long long n, double *__restrict__ Yg,
                                                              double *__restrict__ Vg, double *__restrict__ Wg,
                                                              double *__restrict__ seg, long long nseg, int direct,
                                                              double *__restrict__ xout) {
    __shared__ double sA[TRI_T], sD[TRI_T], sC[TRI_T], sY[TRI_T], sV[TRI_T], sW[TRI_T];
    const int t = threadIdx.x;
    const long long g = blockIdx.x;
    const long long e0 = g * (long long)TRI_S + (long long)t * TRI_M;
    constexpr int L = TRI_M - 2;  // last interior index

    double A[TRI_M], D[TRI_M], C[TRI_M], B[TRI_M];
    prov.load(e0, A, D, C, B);

    // --- thread-level elimination of the M-1 interior unknowns, 3 right-hand sides
    double y[TRI_M - 1], v[TRI_M - 1], w[TRI_M - 1], cp[TRI_M - 1];
    {
        double inv = 1.0 / D[0];
        cp[0] = C[0] * inv; y[0] = B[0] * inv; v[0] = A[0] * inv;
#pragma unroll
        for (int k = 1; k <= L; ++k) {
            inv = 1.0 / (D[k] - A[k] * cp[k - 1]);
            cp[k] = C[k] * inv;
            y[k] = (B[k] - A[k] * y[k - 1]) * inv;
            v[k] = (-A[k] * v[k - 1]) * inv;
        }
        w[L] = cp[L];
#pragma unroll
        for (int k = L - 1; k >= 0; --k) {
            y[k] = y[k] - cp[k] * y[k + 1];
            v[k] = v[k] - cp[k] * v[k + 1];
            w[k] = -cp[k] * w[k + 1];
        }
    }
    // --- exchange first-interior values with the left neighbour thread
    sY[t] = y[0]; sV[t] = v[0]; sW[t] = w[0];
    __syncthreads();
    const double ar = A[TRI_M - 1], dr = D[TRI_M - 1], cr = C[TRI_M - 1], br = B[TRI_M - 1];
    double rA, rD, rC, rY, rV, rW;  // this thread's separator row, 3 rhs
    if (t < TRI_T - 1) {
        const double yF = sY[t + 1], vF = sV[t + 1], wF = sW[t + 1];
        rA = -ar * v[L];
        rD = dr - ar * w[L] - cr * vF;
        rC = -cr * wF;
        rY = br - ar * y[L];
        rY = rY - cr * yF;
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
        const double al = -rA / dL, ga = -rC / dH;
        rD = rD + al * cL + ga * aH;
        rY = rY + al * yL + ga * yH;
        rV = rV + al * vL + ga * vH;
        rW = rW + al * wL + ga * wH;
        rA = al * aL;
        rC = ga * cH;
        __syncthreads();
    }
    // separator solution, as coefficients of (1, Sg[g-1], Sg[g])
    double sy, sv, sw;
    if (t < TRI_T - 1) { const double inv = 1.0 / rD; sy = rY * inv; sv = rV * inv; sw = rW * inv; }
    else { sy = 0.0; sv = 0.0; sw = -1.0; }
    sY[t] = sy; sV[t] = sv; sW[t] = sw;
    __syncthreads();
    double py = 0.0, pv = -1.0, pw = 0.0;  // the separator to the left of this chunk
    if (t > 0) { py = sY[t - 1]; pv = sV[t - 1]; pw = sW[t - 1]; }

    // --- per-element coefficients of the segment-level representation
    double Ye[TRI_M], Ve[TRI_M], We[TRI_M];
#pragma unroll
    for (int k = 0; k <= L; ++k) {
        Ye[k] = y[k] - py * v[k] - sy * w[k];
        Ve[k] = -(pv * v[k] + sv * w[k]);
        We[k] = -(pw * v[k] + sw * w[k]);
    }
    Ye[TRI_M - 1] = sy; Ve[TRI_M - 1] = sv; We[TRI_M - 1] = sw;

    if (direct) {
        // single segment: Sg[-1] = 0 and the separator row has no right neighbour
        __shared__ double sSig;
        if (t == TRI_T - 1) sSig = (br - ar * Ye[L]) / (dr - ar * We[L]);
        __syncthreads();
        const double sig = sSig;
#pragma unroll
        for (int k = 0; k < TRI_M; ++k) {
            const long long idx = e0 + k;
            if (idx < n) xout[idx] = Ye[k] - sig * We[k];
        }
        return;
    }
    {
        double2 *py2 = reinterpret_cast<double2 *>(Yg + e0);
        double2 *pv2 = reinterpret_cast<double2 *>(Vg + e0);
        double2 *pw2 = reinterpret_cast<double2 *>(Wg + e0);
#pragma unroll
        for (int q = 0; q < TRI_M / 2; ++q) {
            py2[q] = make_double2(Ye[2 * q], Ye[2 * q + 1]);
            pv2[q] = make_double2(Ve[2 * q], Ve[2 * q + 1]);
            pw2[q] = make_double2(We[2 * q], We[2 * q + 1]);
        }
    }
    if (t == 0) { seg[g] = Ye[0]; seg[nseg + g] = Ve[0]; seg[2 * nseg + g] = We[0]; }
    if (t == TRI_T - 1) {
        seg[3 * nseg + g] = Ye[L]; seg[4 * nseg + g] = Ve[L]; seg[5 * nseg + g] = We[L];
        seg[6 * nseg + g] = ar; seg[7 * nseg + g] = dr; seg[8 * nseg + g] = cr; seg[9 * nseg + g] = br;
    }
}

// x = Y - Sg[g-1]*V - Sg[g]*W
__global__ void __launch_bounds__(256) tri_finalize_kernel(long long n, const double *__restrict__ Y,
                                                           const double *__restrict__ V, const double *__restrict__ W,
                                                           const double *__restrict__ sig, double *__restrict__ x) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const long long g = idx / TRI_S;
    const double sl = g > 0 ? sig[g - 1] : 0.0;
    x[idx] = Y[idx] - sl * V[idx] - sig[g] * W[idx];
}

// ---------------------------------------------------------------------- host driver
static long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

int w2_tri_prepare(wolfd2_ctx *c, long long nmax) {
    W2TriWork &w = c->tri;
    memset(&w, 0, sizeof(w));
    w.cap = round_up(nmax, TRI_S) + TRI_S;
    W2_CUDA(cudaMalloc((void **)&w.Y0, w.cap * sizeof(double)));
    W2_CUDA(cudaMalloc((void **)&w.V0, w.cap * sizeof(double)));
    W2_CUDA(cudaMalloc((void **)&w.W0, w.cap * sizeof(double)));
    long long n = nmax;
    int l = 0;
    while (true) {
        W2TriLevel &lv = w.lv[l];
        lv.n = n; lv.seg_len = TRI_S; lv.nseg = (n + TRI_S - 1) / TRI_S;
        const long long pad = round_up(lv.n, TRI_S) + TRI_S;
        if (l == 0) { lv.Y = w.Y0; lv.V = w.V0; lv.W = w.W0; lv.x = nullptr; }
        else {
            W2_CUDA(cudaMalloc((void **)&lv.Y, pad * sizeof(double)));
            W2_CUDA(cudaMalloc((void **)&lv.V, pad * sizeof(double)));
            W2_CUDA(cudaMalloc((void **)&lv.W, pad * sizeof(double)));
            W2_CUDA(cudaMalloc((void **)&lv.x, pad * sizeof(double)));
        }
        W2_CUDA(cudaMalloc((void **)&lv.seg, 10 * (lv.nseg + 1) * sizeof(double)));
        ++l;
        if (lv.nseg == 1) break;
        n = lv.nseg;
        if (l >= 6) { w2_set_error("tridiagonal hierarchy too deep"); return W2_ERR_BAD_ARG; }
    }
    w.nlevels = l;
    return W2_OK;
}

void w2_tri_release(wolfd2_ctx *c) {
    W2TriWork &w = c->tri;
    cudaFree(w.Y0); cudaFree(w.V0); cudaFree(w.W0);
    for (int l = 0; l < w.nlevels; ++l) {
        if (l > 0) { cudaFree(w.lv[l].Y); cudaFree(w.lv[l].V); cudaFree(w.lv[l].W); cudaFree(w.lv[l].x); }
        cudaFree(w.lv[l].seg);
    }
    memset(&w, 0, sizeof(w));
}

static int tri_solve_impl(wolfd2_ctx *c, const ProvSoA &p0, double *x) {
    W2TriWork &w = c->tri;
    const long long n0 = p0.n;
    if (n0 < 2) { w2_set_error("tridiagonal system too small"); return W2_ERR_BAD_ARG; }
    if (round_up(n0, TRI_S) > w.cap) { w2_set_error("tridiagonal system of %lld exceeds prepared capacity", n0); return W2_ERR_BAD_ARG; }
    // sizes of the hierarchy for this n
    long long ns[6], segs[6];
    int nl = 0;
    for (long long n = n0;; ) {
        ns[nl] = n; segs[nl] = (n + TRI_S - 1) / TRI_S; ++nl;
        if (segs[nl - 1] == 1) break;
        n = segs[nl - 1];
    }
    // upward sweep
    for (int l = 0; l < nl; ++l) {
        W2TriLevel &lv = w.lv[l];
        const int direct = (l == nl - 1);
        double *xo = direct ? (l == 0 ? x : lv.x) : nullptr;
        if (l == 0)
            tri_reduce_kernel<ProvSoA><<<(unsigned)segs[l], TRI_T, 0, c->stream>>>(p0, ns[l], lv.Y, lv.V, lv.W, lv.seg, segs[l], direct, xo);
        else {
            ProvSeg ps{w.lv[l - 1].seg, ns[l]};
            tri_reduce_kernel<ProvSeg><<<(unsigned)segs[l], TRI_T, 0, c->stream>>>(ps, ns[l], lv.Y, lv.V, lv.W, lv.seg, segs[l], direct, xo);
        }
        c->launches[1]++;
    }
    // downward sweep
    for (int l = nl - 2; l >= 0; --l) {
        W2TriLevel &lv = w.lv[l];
        double *xo = (l == 0) ? x : lv.x;
        const unsigned blocks = (unsigned)((ns[l] + 255) / 256);
        tri_finalize_kernel<<<blocks, 256, 0, c->stream>>>(ns[l], lv.Y, lv.V, lv.W, w.lv[l + 1].x, xo);
        c->launches[1]++;
    }
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}

int w2_tri_solve(wolfd2_ctx *c, long long n, const double *a, const double *d, const double *cc, const double *b,
                 double *x, int quirk) {
    ProvSoA p{a, d, cc, b, n, quirk, 0};
    return tri_solve_impl(c, p, x);
}

int w2_tri_solve_lines(wolfd2_ctx *c, long long nlines, int len, const double *a, const double *d, const double *cc,
                       const double *b, double *x, int quirk) {
    ProvSoA p{a, d, cc, b, nlines * (long long)len, quirk, len};
    return tri_solve_impl(c, p, x);
}
