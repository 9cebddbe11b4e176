// w2_sor_resident.cu -- red/black point SOR (SorRB / SorRBP, src/pressure.f:457-656) for grids whose pressure
// field fits in the shared memory of the whole GPU (1024^2 and below: BASELINE configs[1]).
//
// At these sizes the working set sits in L2 and the streamed, temporally blocked kernel of w2_sor_fused.cu is bound
// by its own fixed costs: bands of 18 rows carry 8 halo rows and a 13-row pipeline fill, and every two iterations
// cost a launch.  Here the WHOLE solve is one cooperative launch: each CTA (one per SM) keeps its band of rows of p
// and of the two coefficient arrays rau, rgv in shared memory for the duration of the solve, each thread keeps the
// right-hand sides of the cells it owns in registers, and an iteration is two half-sweeps, each followed by an
// exchange of the band's edge rows through L2 and a grid barrier.  The convergence test `m > 1 .and. dif < sortol`
// (:534) is taken by every CTA from the same global max-norm, so the loop ends on the device with no host
// involvement.  Per point the arithmetic is that of the reference (:509-513) with the FMA contraction off and the
// bit-exact division of w2.cuh: iterates, max-norms and iteration counts are those of SorRB.
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>

#include "w2.cuh"

namespace cg = cooperative_groups;

#define SR_T 512          // threads per CTA: one per cell of a colour in a row (nx <= 2*SR_T + 1)

struct SorRArgs {
    int nx, ny, pitch, rows_per_cta, msorit, has_mask;
    double sorrel, sortol;
    const double *rau, *rgv, *b;
    const unsigned char *mask;
    double *p;
    unsigned long long *slots;   // 3 rotating max-norm slots
    int *ctl;                    // SorFCtl-compatible block (w2_sor_fused.cu): done, m, nconv, ticket, cur, redo
};

// Shared rows are colour-split like the streamed kernel's: [cells with even i | cells with odd i], so that the
// threads of a half-sweep touch consecutive words.  hw = words per half row.
__device__ __forceinline__ int sr_col(int i, int hw) { return (i & 1) * hw + (i >> 1); }

template <int RMAX>
__global__ void __launch_bounds__(SR_T, 1) sor_rb_resident_kernel(SorRArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[32];
    const int nx = a.nx, ny = a.ny, pitch = a.pitch, t = threadIdx.x;
    const int hw = (nx + 3) >> 1, W = 2 * hw;              // i = 0..nx+1 -> nx+2 cells, padded to even
    const int jA = 2 + blockIdx.x * a.rows_per_cta, jB = min(ny, jA + a.rows_per_cta - 1);
    const int R = jB - jA + 1;                             // rows of this band (>= 1)
    double *sP = sm;                                       // rows jA-1 .. jB+1
    double *sU = sP + (size_t)(RMAX + 2) * W;              // rau rows jA .. jB
    double *sV = sU + (size_t)RMAX * W;                    // rgv rows jA-1 .. jB
    // ---- load the band (coalesced global reads, split shared writes)
    for (int r = 0; r < R + 2; ++r)
        for (int i = t; i <= nx + 1; i += SR_T) sP[r * W + sr_col(i, hw)] = a.p[IDX(i, jA - 1 + r)];
    for (int r = 0; r < R; ++r)
        for (int i = t; i <= nx + 1; i += SR_T) sU[r * W + sr_col(i, hw)] = a.rau[IDX(i, jA + r)];
    for (int r = 0; r < R + 1; ++r)
        for (int i = t; i <= nx + 1; i += SR_T) sV[r * W + sr_col(i, hw)] = a.rgv[IDX(i, jA - 1 + r)];
    // ---- the cells this thread owns: in band row r and colour c the cell i = 2 + ((c + j) & 1) + 2 t
    double breg[RMAX][2];
    unsigned own = 0, msk = 0;                             // bit 2r+c: cell exists / is an identity row (blockage)
#pragma unroll
    for (int r = 0; r < RMAX; ++r)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int j = jA + r, i = 2 + ((c + j) & 1) + 2 * t;
            const bool ex = r < R && i <= nx;
            breg[r][c] = ex ? a.b[IDX(i, j)] : 0.0;
            own |= (unsigned)ex << (2 * r + c);
            msk |= (unsigned)(ex && a.has_mask && a.mask[IDX(i, j)]) << (2 * r + c);
        }
    const double sorrel = a.sorrel;
    const bool multi = gridDim.x > 1;
    if (multi) grid.sync();                                // nobody writes an edge row before everybody has loaded its halo rows
    __syncthreads();
    int nconv = 0, m = 1;
    for (; m <= a.msorit; ++m) {
        double lmax = 0.0;
#pragma unroll
        for (int c = 0; c < 2; ++c) {                      // black (i+j even) first, then red (:505-532)
#pragma unroll
            for (int r0 = 0; r0 < RMAX; r0 += 4) {         // four rows at a time: four independent updates in flight per thread
                double pcv[4], sumv[4], a3v[4], qdv[4];
                bool okv[4];
                bool all_ok = true;
#pragma unroll
                for (int k = 0; k < 4; ++k) {              // straight-line code: every quotient is formed before anything is branched on
                    const int r = r0 + k, j = jA + r, i = 2 + ((c + j) & 1) + 2 * t;
                    const bool ex = (own >> (2 * r + c)) & 1u;
                    const int ic = ex ? i : 2;             // cells that do not exist compute on cell (2, jA) and store nothing;
                    const int rr = ex ? r : 0;             // their five "pressures" come from rows of rgv, which nobody writes
                    const int q = sr_col(ic, hw), qw = sr_col(ic - 1, hw);   // (racecheck r02: reads of p(2,jA) raced with its owner's store)
                    // (0.8 us per iteration at 1024^2 for the select: 9.65 -> 10.5; reading p there instead is a benign race
                    // -- the value is discarded -- but it is a race, with the cell's owner or with its neighbours' owners)
                    const double *rowP = ex ? sP + (rr + 1) * W : sV + W;
                    const double pc = rowP[q], pW = rowP[qw], pE = rowP[qw + 1], pS = rowP[q - W], pN = rowP[q + W];
                    const double a1 = sV[rr * W + q], a5 = sV[(rr + 1) * W + q], a2 = sU[rr * W + qw], a4 = sU[rr * W + q];
                    a3v[k] = -a4 - a2 - a5 - a1;
                    sumv[k] = breg[r][c] - a1 * pS - a2 * pW - a4 * pE - a5 * pN;
                    qdv[k] = w2_div_fast(sumv[k], a3v[k], okv[k]);
                    pcv[k] = pc;
                    all_ok = all_ok && (okv[k] || !ex);
                }
                if (!all_ok) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (!okv[k] && ((own >> (2 * (r0 + k) + c)) & 1u)) qdv[k] = w2_div_detour(sumv[k], a3v[k]);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int r = r0 + k, j = jA + r, i = 2 + ((c + j) & 1) + 2 * t;
                    const bool ex = (own >> (2 * r + c)) & 1u;
                    double sum = qdv[k] - pcv[k];
                    if ((msk >> (2 * r + c)) & 1u) sum = 0.0 - pcv[k];     // identity row: a = (0,0,1,0,0), b = 0 (:123-137)
                    if (ex) {
                        const double pn = pcv[k] + sorrel * sum;
                        sP[(r + 1) * W + sr_col(i, hw)] = pn;
                        lmax = fmax(lmax, fabs(sum));
                        // the band's first and last rows are the neighbours' halo rows: through L2
                        if (multi && (r == 0 || r == R - 1)) a.p[IDX(i, j)] = pn;
                    }
                }
            }
            if (c == 1) {                                  // max |sum| of the iteration (:534): one atomic per CTA
                lmax = w2_block_max(lmax, red);
                if (t == 0) {
                    atomicMax(a.slots + (m % 3), w2_dbits(lmax));
                    if (blockIdx.x == 0) a.slots[(m + 1) % 3] = 0ull;
                }
            }
            if (multi) {
                grid.sync();                               // edge rows (and, after red, the max-norms) of every band are out
                // cells of colour c in the rows jA-1 and jB+1 (other CTAs' rows; the physical ghost rows never change)
                for (int side = 0; side < 2; ++side) {
                    const int j = side ? jB + 1 : jA - 1;
                    if (j < 2 || j > ny) continue;
                    const int i = 2 + ((c + j) & 1) + 2 * t;
                    if (i <= nx) sP[(side ? R + 1 : 0) * W + sr_col(i, hw)] = __ldcg(a.p + IDX(i, j));
                }
            }
            __syncthreads();
        }
        if (!multi) { __threadfence(); __syncthreads(); }
        const double dif = __longlong_as_double((long long)__ldcg(a.slots + (m % 3)));
        if (m > 1 && dif < a.sortol) { nconv = m; break; }  // :534-537
    }
    // ---- the band goes back to the field
    for (int r = 0; r < R; ++r)
        for (int i = 2 + t; i <= nx; i += SR_T) a.p[IDX(i, jA + r)] = sP[(r + 1) * W + sr_col(i, hw)];
    if (blockIdx.x == 0 && t == 0) {
        a.ctl[0] = 1;                                       // done
        a.ctl[1] = nconv ? nconv : a.msorit;                // m: iterations run
        a.ctl[2] = nconv;                                   // nconv
        a.ctl[3] = 0; a.ctl[4] = 0; a.ctl[5] = 0;
    }
}

int g_sor_resident = 1;   // option "sor_resident": 0 = never use this kernel

// Can this context's pressure solve run resident?  One GPU, a row of one colour fits the CTA, at most 8 rows per SM,
// and the band (p with halo rows, rau, rgv) fits the shared memory of an SM.
static bool resident_plan(const wolfd2_ctx *c, int *rows_per_cta, int *nblocks, int *rmax, size_t *smem) {
    if (!g_sor_resident || c->world > 1 || !c->coop_ok) return false;
    const int nx = c->nx, ny = c->ny;
    if (nx < 4 || (nx - 1 + 1) / 2 > SR_T) return false;
    const int rows = ny - 1;
    int rpc = (rows + c->num_sms - 1) / c->num_sms;
    if (rpc > 8) return false;
    const int rm = rpc <= 4 ? 4 : 8;
    const int hw = (nx + 3) >> 1, W = 2 * hw;
    const size_t bytes = (size_t)((rm + 2) + rm + (rm + 1)) * W * sizeof(double);
    if (bytes > 227 * 1024 - 1024) return false;
    *rows_per_cta = rpc; *nblocks = (rows + rpc - 1) / rpc; *rmax = rm; *smem = bytes;
    return true;
}

bool w2_sor_resident_ok(const wolfd2_ctx *c) {
    int a, b, r; size_t s;
    return resident_plan(c, &a, &b, &r, &s);
}

// The whole SorRB loop on p in one cooperative launch; b (W2_F_B) holds div/dk in the field layout.  The outcome lands
// in the control block of the fused solver (and from there in pinned memory): w2_sor_collect reads it.
int w2_sor_resident(wolfd2_ctx *c, double *p) {
    int rpc = 0, nb = 0, rm = 0;
    size_t smem = 0;
    if (!resident_plan(c, &rpc, &nb, &rm, &smem)) { w2_set_error("resident SOR: grid does not fit"); return W2_ERR_BAD_ARG; }
    const wolfd2_params &par = c->par;
    SorRArgs a;
    a.nx = c->nx; a.ny = c->ny; a.pitch = c->pitch; a.rows_per_cta = rpc; a.msorit = par.msorit;
    a.has_mask = c->hreg.has_blockage; a.sorrel = par.sorrel; a.sortol = par.sortol;
    a.rau = c->met.rau; a.rgv = c->met.rgv; a.b = c->fld[W2_F_B]; a.mask = c->pmask; a.p = p;
    a.slots = c->d_norm + 12; a.ctl = c->d_flags;
    W2_CUDA(cudaMemsetAsync(a.slots, 0, 3 * sizeof(unsigned long long), c->stream));
    W2_CUDA(cudaMemsetAsync(c->d_flags, 0, 32 * sizeof(int), c->stream));
    void *args[] = {&a};
    const void *fn = rm == 4 ? (const void *)sor_rb_resident_kernel<4> : (const void *)sor_rb_resident_kernel<8>;
    static bool attr[W2_MAXDEV][2] = {};
    if (!attr[c->device % W2_MAXDEV][rm == 8]) {
        // (the kernel also holds 256 bytes of static shared memory: the opt-in limit covers both)
        W2_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
        attr[c->device % W2_MAXDEV][rm == 8] = true;
    }
    W2_CUDA(cudaLaunchCooperativeKernel(fn, dim3(nb), dim3(SR_T), args, smem, c->stream));
    c->launches[2]++;
    W2_CUDA(cudaMemcpyAsync(c->h_sor, c->d_flags, 32 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    return W2_OK;
}
