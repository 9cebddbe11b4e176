// w2_bc.cu -- ghost-cell fills: VelBoundCond (src/bound_cond.f:511-847), PresBoundCond
// (:853-1024) and VelOutflowBCs (:1656-1874).
//
// The reference visits regions (jreg outer, ireg inner) and faces (W,E,S,N) strictly in order,
// and later loops read what earlier ones wrote at region corners, so the order is part of the
// result.  The work is O(perimeter): one CTA walks the same sequence, running each Fortran
// loop in parallel across its threads with a barrier after every loop.  Loops with a carried
// dependence (the OUTLT2 mass-conservation recurrences, e.g. bound_cond.f:611-614) keep their serial
// order of additions -- a parallel prefix sum would round differently -- but only the two additions
// per element are serial: bc_scan forms the products of a chunk of the face in parallel into shared
// memory, one thread runs the carried recurrence over the chunk from there, and the chunk is written
// back in parallel (a 4096-point face: 50 us instead of 2 ms of dependent global loads).
// HBM traffic is negligible (kernel K3).
#include <string.h>

#include "w2.cuh"

#define BC_THREADS 1024
#define U(i, j) u[IDX(i, j)]
#define V(i, j) v[IDX(i, j)]
#define P(i, j) p[IDX(i, j)]
#define PFOR(var, lo, hi) for (int var = (lo) + (int)threadIdx.x; var <= (hi); var += BC_THREADS)
// row loops are clipped to the rows [jlo, jhi] this rank holds (all rows on one GPU); a south / north face
// is applied only where both rows it touches are held (w2_dist.cu: the outermost halo row is spare)
#define PFORJ(var, lo, hi) for (int var = max((lo), jlo) + (int)threadIdx.x; var <= min((hi), jhi); var += BC_THREADS)
#define ROWS_HELD(a, b) ((a) >= jlo && (b) <= jhi)
#define SEQ if (threadIdx.x == 0)
#define BC_CHUNK 1024

// for k = lo..hi (in order):  prev = (sgn*prev + t1(k)) + t2(k);  put(k, prev)      -- all threads of the CTA call this.
// t1, t2 read values that the recurrence itself does not write; sgn is +1 or -1 (exact).
template <class T1, class T2, class Put>
__device__ __forceinline__ void bc_scan(double *s1, double *s2, double prev0, double sgn, int lo, int hi, T1 t1, T2 t2, Put put) {
    double &s_prev = s2[BC_CHUNK];     // s1: BC_CHUNK doubles, s2: BC_CHUNK + 1
    if (threadIdx.x == 0) s_prev = prev0;
    for (int c0 = lo; c0 <= hi; c0 += BC_CHUNK) {
        const int n = min(BC_CHUNK, hi - c0 + 1);
        __syncthreads();
        for (int k = threadIdx.x; k < n; k += BC_THREADS) { s1[k] = t1(c0 + k); s2[k] = t2(c0 + k); }
        __syncthreads();
        if (threadIdx.x == 0) {
            double prev = s_prev;
            for (int k = 0; k < n; ++k) {
                prev = (sgn * prev + s1[k]) + s2[k];
                s1[k] = prev;
            }
            s_prev = prev;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < n; k += BC_THREADS) put(c0 + k, s1[k]);
    }
    __syncthreads();
}

__device__ __forceinline__ unsigned long long bc_ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// The same recurrence on a row slab (west / east faces: k is the row index): the addends of the rows lo..hi are spread
// over the ranks.  Each rank forms those of the rows it updates (E0..E1) and stores them -- and the start value, if it
// updates row lo-1 -- into EVERY rank's gather buffer (peer stores over NVLink), publishes the scan number in every
// rank's flag word and waits for all the others: a barrier across the GPUs.  Then every rank runs the whole recurrence
// from its own copy, i.e. exactly the additions of bc_scan in the same order, and keeps the rows it holds (jlo..jhi).
// Buffers alternate with the parity of the scan number: nobody can be two scans ahead of a rank that is still reading.
template <class T1, class T2, class P0, class Put>
__device__ __forceinline__ void bc_scan_slab(double *s1, double *s2, const W2BcPeer &bp, unsigned long long seq, double sgn, int lo,
                                             int hi, int jlo, int jhi, T1 t1, T2 t2, P0 prev0, Put put) {
    const int ld = bp.ld, half = (int)(seq & 1ull);
    const size_t boff = (size_t)half * (2 * (size_t)ld + 2);
    for (int k = max(lo, bp.E0) + (int)threadIdx.x; k <= min(hi, bp.E1); k += BC_THREADS) {
        const double a = t1(k), b = t2(k);
        for (int r = 0; r < bp.world; ++r) {
            double *g = reinterpret_cast<double *>(bp.g[r] + 1) + boff;
            g[k] = a; g[ld + k] = b;
        }
    }
    if (threadIdx.x == 0 && lo - 1 >= bp.E0 && lo - 1 <= bp.E1) {
        const double p0 = prev0();
        for (int r = 0; r < bp.world; ++r) (reinterpret_cast<double *>(bp.g[r] + 1) + boff)[2 * ld] = p0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        for (int r = 0; r < bp.world; ++r) {
            volatile unsigned long long *fl = &bp.g[r]->flag[bp.rank];
            *fl = seq;
        }
    }
    if ((int)threadIdx.x < bp.world) {
        const long long t0 = clock64();
        while (bc_ld_acquire_sys(&bp.g[bp.rank]->flag[threadIdx.x]) < seq)
            if (clock64() - t0 > 20000000000ll) { bp.g[bp.rank]->timeout = 1; break; }   // ~10 s: a peer is gone
    }
    __syncthreads();
    const double *g = reinterpret_cast<const double *>(bp.g[bp.rank] + 1) + boff;
    bc_scan(s1, s2, g[2 * ld], sgn, lo, hi,
            [&](int k) { return __ldcg(g + k); }, [&](int k) { return __ldcg(g + ld + k); },
            [&](int k, double x) { if (k >= jlo && k <= jhi) put(k, x); });
}

template <bool kOutflowOnly>
__global__ void __launch_bounds__(BC_THREADS) vel_bc_kernel(const W2Regions *__restrict__ R, int pitch,
                                                            int jlo, int jhi, double *u, double *v, const int *done, W2BcPeer bp) {
    if (done != nullptr && *done) return;   // VelOutflowBCs of a speculatively enqueued QL iteration
    unsigned long long scan = bp.seq0;      // slab runs: one number per OUTLT2 west / east face, in visiting order
    const double dZero = 0.0, dTwo = 2.0, dThree = 3.0, dFour = 4.0, dFive = 5.0, dEight = 8.0;
    __shared__ double sc1[BC_CHUNK], sc2[BC_CHUNK + 1];   // bc_scan staging
    const int nreg = R->nreg;
    for (int q = 0; q < nreg; ++q) {
        const int iW = R->iW[q], iE = R->iE[q], jS = R->jS[q], jN = R->jN[q];
        // ---------------- WEST (:561-619 / :1707-1736)
        {
            const int bt = R->bd[q][W2_WEST - 1];
            const double valU = R->val[q][W2_WEST - 1][W2_U - 1], valV = R->val[q][W2_WEST - 1][W2_V - 1];
            if (!kOutflowOnly && bt == W2_BM_WALL1) {
                PFORJ(j, jS, jN) U(iW, j) = dZero;
                __syncthreads();
                PFORJ(j, jS + 1, jN) V(iW, j) = dTwo * valV - V(iW + 1, j);
            } else if (!kOutflowOnly && bt == W2_BM_WALL2) {
                PFORJ(j, jS, jN) U(iW, j) = dZero;
                __syncthreads();
                PFORJ(j, jS + 1, jN) V(iW, j) = V(iW + 1, j);
            } else if (!kOutflowOnly && bt == W2_BM_INLET) {
                PFORJ(j, jS, jN) U(iW, j) = valU;
                __syncthreads();
                PFORJ(j, jS + 1, jN) V(iW, j) = dTwo * valV - V(iW + 1, j);
            } else if (bt == W2_BM_OUTLT1) {
                PFORJ(j, jS, jN) U(iW - 1, j) = valU + U(iW, j);
                __syncthreads();
                PFORJ(j, jS + 1, jN) V(iW, j) = -V(iW + 1, j);
            } else if (bt == W2_BM_OUTLT2) {
                if (kOutflowOnly) {  // VelOutflowBCs :1725
                    PFORJ(j, jS + 1, jN) U(iW, j) = U(iW + 1, j) - V(iW + 1, j) + V(iW + 1, j - 1);
                } else {             // VelBoundCond :608
                    PFORJ(j, jS + 1, jN) U(iW, j) = U(iW + 1, j) + V(iW + 1, j) - V(iW + 1, j - 1);
                }
                __syncthreads();
                // :612-613  v(iW,j) = -v(iW,j-1) + 5*(...) + 8*(...)
                auto t1 = [&](int j) { return dFive * (V(iW + 1, j) - V(iW + 1, j - 1)); };
                auto t2 = [&](int j) { return dEight * (U(iW + 1, j) - U(iW, j)); };
                auto put = [&](int j, double x) { V(iW, j) = x; };
                if (bp.world > 1) bc_scan_slab(sc1, sc2, bp, scan++, -1.0, jS + 1, jN, jlo, jhi, t1, t2, [&]() { return V(iW, jS); }, put);
                else bc_scan(sc1, sc2, V(iW, jS), -1.0, jS + 1, jN, t1, t2, put);
            }
            __syncthreads();
        }
        // ---------------- EAST (:623-693 / :1740-1780)
        {
            const int bt = R->bd[q][W2_EAST - 1];
            const double valU = R->val[q][W2_EAST - 1][W2_U - 1], valV = R->val[q][W2_EAST - 1][W2_V - 1];
            if (!kOutflowOnly && bt == W2_BM_WALL1) {
                PFORJ(j, jS, jN) U(iE, j) = dZero;
                __syncthreads();
                PFORJ(j, jS + 1, jN) V(iE + 1, j) = dTwo * valV - V(iE, j);
            } else if (!kOutflowOnly && bt == W2_BM_WALL2) {
                PFORJ(j, jS, jN) U(iE, j) = dZero;
                __syncthreads();
                PFORJ(j, jS + 1, jN) V(iE + 1, j) = V(iE, j);
            } else if (!kOutflowOnly && bt == W2_BM_INLET) {
                PFORJ(j, jS, jN) U(iE, j) = valU;
                __syncthreads();
                PFORJ(j, jS + 1, jN) V(iE + 1, j) = dTwo * valV - V(iE, j);
            } else if (bt == W2_BM_OUTLT1) {
                PFORJ(j, jS, jN) U(iE + 1, j) = valU + U(iE, j);
                __syncthreads();
                PFORJ(j, jS + 1, jN) V(iE + 1, j) = -V(iE, j);
            } else if (bt == W2_BM_OUTLT2) {
                PFORJ(j, jS + 1, jN) U(iE, j) = U(iE - 1, j) - (V(iE, j) - V(iE, j - 1));
                __syncthreads();
                // :684-687  x - 4*(...) is x + (-(4*(...))): negation is exact
                auto t1 = [&](int j) { return dThree * (V(iE, j - 1) - V(iE, j)); };
                auto t2 = [&](int j) { return -(dFour * (U(iE, j) - U(iE - 1, j))); };
                auto put = [&](int j, double x) { V(iE + 1, j) = x; };
                if (bp.world > 1) bc_scan_slab(sc1, sc2, bp, scan++, 1.0, jS + 1, jN - 1, jlo, jhi, t1, t2, [&]() { return V(iE + 1, jS); }, put);
                else bc_scan(sc1, sc2, V(iE + 1, jS), 1.0, jS + 1, jN - 1, t1, t2, put);
            }
            __syncthreads();
        }
        // ---------------- SOUTH (:697-766 / :1784-1824)
        {
            const int bt = ROWS_HELD(jS, jS + 1) ? R->bd[q][W2_SOUTH - 1] : W2_BM_INTERN;
            const double valU = R->val[q][W2_SOUTH - 1][W2_U - 1], valV = R->val[q][W2_SOUTH - 1][W2_V - 1];
            if (!kOutflowOnly && bt == W2_BM_WALL1) {
                PFOR(i, iW + 1, iE) U(i, jS) = dTwo * valU - U(i, jS + 1);
                __syncthreads();
                PFOR(i, iW, iE) V(i, jS) = dZero;
            } else if (!kOutflowOnly && bt == W2_BM_WALL2) {
                PFOR(i, iW + 1, iE) U(i, jS) = U(i, jS + 1);
                __syncthreads();
                PFOR(i, iW, iE) V(i, jS) = dZero;
            } else if (!kOutflowOnly && bt == W2_BM_INLET) {
                PFOR(i, iW + 1, iE) U(i, jS) = dTwo * valU - U(i, jS + 1);
                __syncthreads();
                PFOR(i, iW, iE) V(i, jS) = valV;
            } else if (bt == W2_BM_OUTLT1) {
                PFOR(i, iW + 1, iE) U(i, jS) = -U(i, jS + 1);
                __syncthreads();
                PFOR(i, iW, iE) V(i, jS) = valV + V(i, jS);  // sic: self-reference (:738)
            } else if (bt == W2_BM_OUTLT2) {
                PFOR(i, iW + 1, iE) V(i, jS) = V(i, jS + 1) + (U(i, jS + 1) - U(i - 1, jS + 1));
                __syncthreads();
                bc_scan(sc1, sc2, U(iW, jS), 1.0, iW + 1, iE - 1,      // :758-761
                        [&](int i) { return dThree * (U(i - 1, jS + 1) - U(i, jS + 1)); },
                        [&](int i) { return -(dFour * (V(i, jS + 1) - V(i, jS))); },
                        [&](int i, double x) { U(i, jS) = x; });
            }
            __syncthreads();
        }
        // ---------------- NORTH (:770-840 / :1828-1868)
        {
            const int bt = ROWS_HELD(jN - 1, jN + 1) ? R->bd[q][W2_NORTH - 1] : W2_BM_INTERN;
            const double valU = R->val[q][W2_NORTH - 1][W2_U - 1], valV = R->val[q][W2_NORTH - 1][W2_V - 1];
            if (!kOutflowOnly && bt == W2_BM_WALL1) {
                PFOR(i, iW + 1, iE) U(i, jN + 1) = dTwo * valU - U(i, jN);
                __syncthreads();
                PFOR(i, iW, iE) V(i, jN) = dZero;
            } else if (!kOutflowOnly && bt == W2_BM_WALL2) {
                PFOR(i, iW + 1, iE) U(i, jN + 1) = U(i, jN);
                __syncthreads();
                PFOR(i, iW, iE) V(i, jN) = dZero;
            } else if (!kOutflowOnly && bt == W2_BM_INLET) {
                PFOR(i, iW + 1, iE) U(i, jN + 1) = dTwo * valU - U(i, jN);
                __syncthreads();
                PFOR(i, iW, iE) V(i, jN) = valV;
            } else if (bt == W2_BM_OUTLT1) {
                PFOR(i, iW + 1, iE) U(i, jN + 1) = -U(i, jN);
                __syncthreads();
                PFOR(i, iW, iE) V(i, jN + 1) = valV + V(i, jN);
            } else if (bt == W2_BM_OUTLT2) {
                // v(i,jN) for i=iW..iE reads u(i-1,jN) and u(i,jN): no carried dependence
                PFOR(i, iW, iE) V(i, jN) = V(i, jN - 1) - (U(i, jN) - U(i - 1, jN));
                __syncthreads();
                bc_scan(sc1, sc2, U(iW, jN + 1), 1.0, iW + 1, iE - 1,  // :831-834
                        [&](int i) { return dThree * (U(i - 1, jN) - U(i, jN)); },
                        [&](int i) { return -(dFour * (V(i, jN) - V(i, jN - 1))); },
                        [&](int i, double x) { U(i, jN + 1) = x; });
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(BC_THREADS) pres_bc_kernel(const W2Regions *__restrict__ R, int pitch, int jlo, int jhi, double *p) {
    const double dZero = 0.0, dTwo = 2.0;
    const int nreg = R->nreg;
    for (int q = 0; q < nreg; ++q) {
        const int iW = R->iW[q], iE = R->iE[q], jS = R->jS[q], jN = R->jN[q];
        const double vW = R->val[q][W2_WEST - 1][W2_P - 1], vE = R->val[q][W2_EAST - 1][W2_P - 1];
        const double vS = R->val[q][W2_SOUTH - 1][W2_P - 1], vN = R->val[q][W2_NORTH - 1][W2_P - 1];
        if (R->type[q] == W2_RM_BLOCKG) {  // :903-936
            // The zero fill of the interior iW+1..iE x jS+1..jN: its outer layer here, in sequence (neighbouring
            // regions write into it and the four loops below overwrite it); the cells further inside, which no other
            // statement of the routine touches, by blk_zero_kernel on the whole GPU before this kernel starts.
            const int w = iE - iW, h = jN - jS;
            for (int t = threadIdx.x; t < 2 * (w + h); t += BC_THREADS) {
                int i, j;
                if (t < w) { i = iW + 1 + t; j = jS + 1; }
                else if (t < 2 * w) { i = iW + 1 + (t - w); j = jN; }
                else if (t < 2 * w + h) { i = iW + 1; j = jS + 1 + (t - 2 * w); }
                else { i = iE; j = jS + 1 + (t - 2 * w - h); }
                if (j >= jlo && j <= jhi) P(i, j) = dZero;
            }
            __syncthreads();
            PFORJ(j, jS + 1, jN) P(iW + 1, j) = vW + P(iW, j);
            __syncthreads();
            PFORJ(j, jS + 1, jN) P(iE, j) = vE + P(iE + 1, j);
            __syncthreads();
            if (ROWS_HELD(jS, jS + 1)) PFOR(i, iW + 1, iE) P(i, jS + 1) = vS + P(i, jS);
            __syncthreads();
            if (ROWS_HELD(jN, jN + 1)) PFOR(i, iW + 1, iE) P(i, jN) = vN + P(i, jN + 1);
            __syncthreads();
            continue;
        }
        int bt = R->bd[q][W2_WEST - 1];
        if (bt == W2_BM_WALL1 || bt == W2_BM_WALL2 || bt == W2_BM_INLET) {
            PFORJ(j, jS + 1, jN) P(iW, j) = vW + P(iW + 1, j);
        } else if (bt == W2_BM_OUTLT1 || bt == W2_BM_OUTLT2) {
            PFORJ(j, jS + 1, jN) P(iW, j) = dTwo * vW - P(iW + 1, j);
        }
        __syncthreads();
        bt = R->bd[q][W2_EAST - 1];
        if (bt == W2_BM_WALL1 || bt == W2_BM_WALL2 || bt == W2_BM_INLET) {
            PFORJ(j, jS + 1, jN) P(iE + 1, j) = vE + P(iE, j);
        } else if (bt == W2_BM_OUTLT1 || bt == W2_BM_OUTLT2) {
            PFORJ(j, jS + 1, jN) P(iE + 1, j) = dTwo * vE - P(iE, j);
        }
        __syncthreads();
        bt = ROWS_HELD(jS, jS + 1) ? R->bd[q][W2_SOUTH - 1] : W2_BM_INTERN;
        if (bt == W2_BM_WALL1 || bt == W2_BM_WALL2 || bt == W2_BM_INLET) {
            PFOR(i, iW + 1, iE) P(i, jS) = vS + P(i, jS + 1);
        } else if (bt == W2_BM_OUTLT1 || bt == W2_BM_OUTLT2) {
            PFOR(i, iW + 1, iE) P(i, jS) = dTwo * vS - P(i, jS + 1);
        }
        __syncthreads();
        bt = ROWS_HELD(jN, jN + 1) ? R->bd[q][W2_NORTH - 1] : W2_BM_INTERN;
        if (bt == W2_BM_WALL1 || bt == W2_BM_WALL2 || bt == W2_BM_INLET) {
            PFOR(i, iW + 1, iE) P(i, jN + 1) = vN + P(i, jN);
        } else if (bt == W2_BM_OUTLT1 || bt == W2_BM_OUTLT2) {
            PFOR(i, iW + 1, iE) P(i, jN + 1) = dTwo * vN - P(i, jN);
        }
        __syncthreads();
    }
}

// slab runs: the peer plumbing of the OUTLT2 west / east recurrences (bc_scan_slab); every launch of the ghost-fill
// kernels takes as many scan numbers as the deck has such faces, on every rank alike
static int bc_peer(wolfd2_ctx *c, W2BcPeer *bp) {
    memset(bp, 0, sizeof(*bp));
    bp->rank = c->rank; bp->world = 1;
    if (c->world == 1) return W2_OK;
    int nscan = 0;
    for (int q = 0; q < c->hreg.nreg; ++q)
        nscan += (c->hreg.bd[q][W2_WEST - 1] == W2_BM_OUTLT2) + (c->hreg.bd[q][W2_EAST - 1] == W2_BM_OUTLT2);
    if (nscan == 0) return W2_OK;
    if (c->peer.state != 1) {
        w2_set_error("multi-GPU runs need CUDA IPC peer mapping for OUTLT2 (mass_cons) faces on west / east borders");
        return W2_ERR_UNSUPPORTED;
    }
    bp->world = c->world; bp->ld = c->ny + 2; bp->E0 = c->E0; bp->E1 = c->E1;
    bp->seq0 = c->peer.bc_seq + 1;
    c->peer.bc_seq += nscan;
    for (int r = 0; r < c->world; ++r) bp->g[r] = c->peer.bcg[r];
    return W2_OK;
}

int w2_vel_bc(wolfd2_ctx *c, double *u, double *v) {
    W2BcPeer bp;
    W2_TRY(bc_peer(c, &bp));
    vel_bc_kernel<false><<<1, BC_THREADS, 0, c->stream>>>(c->dreg, c->pitch, c->A0, c->A1, u, v, nullptr, bp);
    c->launches[3]++;
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}
int w2_outflow_bc(wolfd2_ctx *c, double *u, double *v, const int *done) {
    // skip the launch when no face is an outlet (VelOutflowBCs is then a no-op, :1709-1710)
    bool any = false;
    for (int q = 0; q < c->hreg.nreg && !any; ++q)
        for (int k = 0; k < 4; ++k)
            if (c->hreg.bd[q][k] == W2_BM_OUTLT1 || c->hreg.bd[q][k] == W2_BM_OUTLT2) any = true;
    if (!any) return W2_OK;
    W2BcPeer bp;
    W2_TRY(bc_peer(c, &bp));
    vel_bc_kernel<true><<<1, BC_THREADS, 0, c->stream>>>(c->dreg, c->pitch, c->A0, c->A1, u, v, done, bp);
    c->launches[1]++;
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}
// deep interior of a blockage region (iW+2..iE-1, jS+2..jN-1): only ever zeroed (:903-908)
__global__ void __launch_bounds__(256) blk_zero_kernel(int iW, int iE, int jS, int jN, int jlo, int jhi, int pitch, double *p) {
    const int i = iW + 2 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > iE - 1) return;
    for (int j = max(jS + 2, jlo) + blockIdx.y; j <= min(jN - 1, jhi); j += gridDim.y) P(i, j) = 0.0;
}

int w2_pres_bc(wolfd2_ctx *c, double *p) {
    if (c->hreg.has_blockage)
        for (int q = 0; q < c->hreg.nreg; ++q) {
            if (c->hreg.type[q] != W2_RM_BLOCKG) continue;
            const int w = c->hreg.iE[q] - c->hreg.iW[q] - 2, h = c->hreg.jN[q] - c->hreg.jS[q] - 2;
            if (w <= 0 || h <= 0) continue;
            dim3 g((w + 255) / 256, h < 1024 ? h : 1024);
            blk_zero_kernel<<<g, 256, 0, c->stream>>>(c->hreg.iW[q], c->hreg.iE[q], c->hreg.jS[q], c->hreg.jN[q], c->A0, c->A1, c->pitch, p);
            c->launches[3]++;
        }
    pres_bc_kernel<<<1, BC_THREADS, 0, c->stream>>>(c->dreg, c->pitch, c->A0, c->A1, p);
    c->launches[3]++;
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}
