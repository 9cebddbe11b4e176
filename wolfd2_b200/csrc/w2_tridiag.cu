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
#include "w2_tri.cuh"

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
__global__ void __launch_bounds__(TRI_T, 512 / TRI_T) tri_reduce_kernel(Prov prov, long long n, double *__restrict__ Yg,
                                                              double *__restrict__ Vg, double *__restrict__ Wg,
                                                              double *__restrict__ seg, long long nseg, int direct,
                                                              double *__restrict__ xout, const int *__restrict__ done) {
    __shared__ double sA[TRI_T], sD[TRI_T], sC[TRI_T], sY[TRI_T], sV[TRI_T], sW[TRI_T];
    if (done != nullptr && *done) return;   // QL loop already converged: this launch was enqueued speculatively
    const int t = threadIdx.x;
    const long long g = blockIdx.x;
    const long long e0 = g * (long long)TRI_S + (long long)t * TRI_M;
    constexpr int L = TRI_M - 2;  // last interior index

    double A[TRI_M], D[TRI_M], C[TRI_M], B[TRI_M];
    prov.load(e0, A, D, C, B);
    double Ye[TRI_M], Ve[TRI_M], We[TRI_M];
    double ar, dr, cr, br;
    tri_cta_core(A, D, C, B, Ye, Ve, We, sA, sD, sC, sY, sV, sW, ar, dr, cr, br);

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
                                                           const double *__restrict__ sig, double *__restrict__ x,
                                                           const int *__restrict__ done) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n || (done != nullptr && *done)) return;
    const long long g = idx / TRI_S;
    const double sl = g > 0 ? sig[g - 1] : 0.0;
    x[idx] = Y[idx] - sl * V[idx] - sig[g] * W[idx];
}

// ---------------------------------------------------------------------- host driver
static long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

// nmax: unknowns of the (global) chain, sizes the segment tables and the upper levels; cap0: unknowns whose
// level-0 spikes this rank stores (= nmax on one GPU)
static int tri_prepare(W2TriWork &w, long long nmax, long long cap0, bool with_y0) {
    memset(&w, 0, sizeof(w));
    w.cap = round_up(cap0, TRI_S) + TRI_S;
    if (with_y0) W2_CUDA(cudaMalloc((void **)&w.Y0, w.cap * sizeof(double)));
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
        if (l == 0) W2_CUDA(cudaMalloc((void **)&w.ext, 2 * (lv.nseg + 1) * sizeof(int)));
        ++l;
        if (lv.nseg == 1) break;
        n = lv.nseg;
        if (l >= 6) { w2_set_error("tridiagonal hierarchy too deep"); return W2_ERR_BAD_ARG; }
    }
    w.nlevels = l;
    return W2_OK;
}
int w2_tri_prepare(wolfd2_ctx *c, long long nmax, long long cap0) {
    c->tri_nmax = nmax;
    return tri_prepare(c->tri, nmax, cap0, true);
}
// the work arrays of the second momentum stream (one GPU; the momentum kernels write Y in the field layout: no Y0)
int w2_tri_prepare_second(wolfd2_ctx *c) { return tri_prepare(c->tri2, c->tri_nmax, c->tri_nmax, false); }

static void tri_release(W2TriWork &w) {
    cudaFree(w.Y0); cudaFree(w.V0); cudaFree(w.W0); cudaFree(w.ext);
    for (int l = 0; l < 6; ++l) {   // every level: a set-up that ran out of memory half way has nlevels == 0 (cudaFree(0) is a no-op)
        if (l > 0) { cudaFree(w.lv[l].Y); cudaFree(w.lv[l].V); cudaFree(w.lv[l].W); cudaFree(w.lv[l].x); }
        cudaFree(w.lv[l].seg);
    }
    memset(&w, 0, sizeof(w));
}
void w2_tri_release(wolfd2_ctx *c) { tri_release(c->tri); tri_release(c->tri2); }

// Levels >= 1: solve the separator system whose rows are formed from the level-0 segment records in
// tri.lv[0].seg (nseg0 of them).  On return *sigma points at the nseg0 separator values.
int w2_tri_upper(wolfd2_ctx *c, long long nseg0, const double **sigma, const int *done) {
    W2TriWork &w = c->tri;
    long long ns[6], segs[6];
    int nl = 1;
    ns[0] = 0; segs[0] = nseg0;
    for (long long n = nseg0;;) {
        ns[nl] = n; segs[nl] = (n + TRI_S - 1) / TRI_S; ++nl;
        if (segs[nl - 1] == 1) break;
        n = segs[nl - 1];
        if (nl >= 6) { w2_set_error("tridiagonal hierarchy too deep"); return W2_ERR_BAD_ARG; }
    }
    for (int l = 1; l < nl; ++l) {
        W2TriLevel &lv = w.lv[l];
        const int direct = (l == nl - 1);
        ProvSeg ps{w.lv[l - 1].seg, ns[l]};
        tri_reduce_kernel<ProvSeg><<<(unsigned)segs[l], TRI_T, 0, c->stream>>>(ps, ns[l], lv.Y, lv.V, lv.W, lv.seg, segs[l], direct,
                                                                               direct ? lv.x : nullptr, done);
        c->launches[1]++;
    }
    for (int l = nl - 2; l >= 1; --l) {
        W2TriLevel &lv = w.lv[l];
        const unsigned blocks = (unsigned)((ns[l] + 255) / 256);
        tri_finalize_kernel<<<blocks, 256, 0, c->stream>>>(ns[l], lv.Y, lv.V, lv.W, w.lv[l + 1].x, lv.x, done);
        c->launches[1]++;
    }
    W2_CUDA(cudaGetLastError());
    *sigma = w.lv[1].x;
    return W2_OK;
}

static int tri_solve_impl(wolfd2_ctx *c, const ProvSoA &p0, double *x) {
    W2TriWork &w = c->tri;
    const long long n0 = p0.n;
    if (n0 < 2) { w2_set_error("tridiagonal system too small"); return W2_ERR_BAD_ARG; }
    if (round_up(n0, TRI_S) > w.cap) { w2_set_error("tridiagonal system of %lld exceeds prepared capacity", n0); return W2_ERR_BAD_ARG; }
    const long long nseg0 = (n0 + TRI_S - 1) / TRI_S;
    W2TriLevel &lv = w.lv[0];
    const int direct = nseg0 == 1;
    tri_reduce_kernel<ProvSoA><<<(unsigned)nseg0, TRI_T, 0, c->stream>>>(p0, n0, lv.Y, lv.V, lv.W, lv.seg, nseg0, direct,
                                                                         direct ? x : nullptr, nullptr);
    c->launches[1]++;
    if (!direct) {
        const double *sigma = nullptr;
        W2_TRY(w2_tri_upper(c, nseg0, &sigma));
        tri_finalize_kernel<<<(unsigned)((n0 + 255) / 256), 256, 0, c->stream>>>(n0, lv.Y, lv.V, lv.W, sigma, x, nullptr);
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
