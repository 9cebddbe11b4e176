// w2_sor_fused.cu -- red/black point SOR (SorRB / SorRBP, src/pressure.f:457-656) with both colours
// and T whole iterations fused into ONE pass over HBM.
//
// The plain half-sweep kernel (w2_ppe.cu) is already at the DRAM limit for what it moves
// (ncu: 652 MB per half-sweep at 4096^2 against 537 MB algorithmic), so the only way up is to move
// fewer bytes.  This kernel reads p, b, rau, rgv once and writes p once per T iterations
// (40/T B/cell/iteration instead of 80), out of place (p_src -> p_dst, ping-pong).
//
// Structure: each CTA owns a strip of 256 columns (256-4T of them "owned", the rest halo that is
// recomputed redundantly -- bit-identical, since every cell's arithmetic is the same wherever it
// runs) and streams down a band of rows.  Rows of the four arrays are staged in a shared-memory ring
// by TMA bulk copies (cp.async.bulk + mbarrier complete_tx), several rows ahead of use: three copies per
// row -- the two colour halves of p and one 6 KB tile holding the strip's pieces of b, rau, rgv side by
// side (sorf_tile_kernel; the copy engine's cost is per copy, not per byte).  A software
// pipeline of 2T half-sweep stages runs on the ring: when row r has landed, stage s (s odd = black
// of iteration (s+1)/2, s even = red) relaxes row r-2s+1 in place.  With that two-row lag every
// stage of a time step reads only data written in earlier time steps, so all 2T stages (4 warps
// each) run concurrently with ONE block barrier per streamed row.  After stage 2T a row is final
// and goes back to HBM.
//
// Arithmetic per point is literally that of the reference (:509-513), FMA contraction off, so the
// iterate path and max-norm are bit-identical to SorRB; the per-iteration max|sum| is accumulated
// separately for each of the T fused iterations.  If the convergence test (:534) fires for an
// iteration in the middle of a pass, the next launch redoes the pass from the untouched source
// with exactly that many iterations -- the returned p and iteration count are those of the
// reference.  Blockage (identity) rows arrive as a NaN sentinel in b.
#include <stdlib.h>

#include "w2.cuh"

// (SF_W 128 -- four CTAs of four warps per SM, one warp per stage -- was measured at 0.203 ms per pass against 0.194;
// SF_W 512 -- one CTA of sixteen warps per SM, half as many bulk copies per cell -- with the tiled coefficients at 0.2175
// against 0.1727: nine strips of 504 owned columns waste 10 % at nx = 4096, and one CTA per SM has nobody to hide behind.)
#ifndef SF_W
#define SF_W 256              // strip width held in shared memory (cells)
#endif
#define SF_HALF (SF_W / 2)    // cells of one colour parity per strip row
// Column pairs per thread (threads per stage = SF_HALF / SF_CPT).  2: each thread relaxes two cells of a row as
// straight-line code (w2_div_fast: no branch inside the update), so the two dependent fp64 chains overlap and the
// per-row costs (barrier, mbarrier wait, ring bookkeeping) are paid once per two updates: 0.2175 -> 0.2101 ms per T=2
// pass at 4096^2.  (1 with the same straight-line update: 0.2654 ms -- every lane then pays the full quotient that the
// old zero-numerator branch skipped; 2 with the branchy update was measured slower than 1 in round 1.)
// Also measured and rejected (bit-exact, slower): a split-phase step barrier -- mbarrier arrive after the stores, the
// next row's coefficients, diagonal and reciprocal formed before the wait -- 0.190 -> 0.206 ms per pass: 28 more live
// registers and the arrive / try_wait pair cost more than the barrier skew they hide (ncu r02: 31 % of the stall
// samples sit on the block barrier).
#ifndef SF_CPT
#define SF_CPT 2
#endif
#define SF_TPS (SF_HALF / SF_CPT)
#ifndef SF_DIAG
#define SF_DIAG 0             // timing experiments only (wrong results): 1 no update arithmetic, 2 no store-back, 4 no step barrier, 8 no TMA
#endif
// (Measured and rejected, round 2b: cp.async.bulk.prefetch.L2 of the rows to come -- the ring holds only D + 1 = 4 rows in
// flight and there is no shared memory for more.  Per row and array piece, 8 rows ahead: 0.193 -> 0.243 ms per pass; on the
// tiled coefficients, whose rows are contiguous, one prefetch for 4 / 8 / 16 / 32 rows: 0.1727 -> 0.176 / 0.176 / 0.183 /
// 0.206 ms.  The prefetches compete with the demand copies instead of shortening them.)
#ifndef SF_PWARP
#define SF_PWARP 0            // 1: an extra warp that only issues the bulk copies -- measured 0.1805 against 0.1728 ms per pass
#endif
#define SF_NTHREADS(T) ((T) * 2 * SF_TPS + 32 * SF_PWARP)
#ifndef SF_BAL_STORE
#define SF_BAL_STORE 1        // finished rows are stored by all threads of the CTA (0: by the last stage's threads, round 2a)
#endif
#define SF_PADL 2             // pad cells on each side of a half row (keeps TMA destinations 16-byte aligned)
#define SF_HSTR (SF_HALF + 2 * SF_PADL)   // one half row in shared memory
#define SF_STRIDE (2 * SF_HSTR)           // shared row stride in doubles: [even-i half | odd-i half]

struct SorFCtl {           // device-resident loop control
    int done;              // 1: solve finished, later launches return at once
    int m;                 // iterations completed
    int nconv;             // converged iteration (0: not converged)
    int ticket;            // CTAs finished in the current pass
    int cur;               // 0: current iterate is buffer A, 1: buffer B
    int redo;              // >0: next pass repeats from the same source with `redo` iterations
    int pad0, pad1;
    unsigned long long slot[8];  // max |sum| of each fused iteration of the current pass
    unsigned long long last_dif;
};

struct SorFArgs {
    int nx, ny, pitch;
    int j0, j1;            // unknown rows relaxed by this launch: 2..ny, or this rank's slab
    int ext_decide;        // 0: one GPU, the last CTA closes the pass; 1: NCCL path, sorf_decide_kernel closes it after
                           // the all-reduce; 2: peer-memory path, sorf_decide_p2p closes it from the mailboxes
    int nstrips, nbands, rows_per_band, own_w;
    int msorit;
    double sorrel, sortol;
    // all five arrays are in the COLOUR-SPLIT device layout: row j = [cells with even i | cells with odd i],
    // each half pitch/2 doubles long (see sorf_pack_kernel).  Every cell relaxed in one half-sweep of a row
    // has the same i-parity, so in this layout a warp's shared-memory accesses are unit-stride and free of
    // the 2-way bank conflicts an interleaved row causes (ncu r01: 45 -> 24 wavefronts per warp-update).
    const double *rau, *rgv, *b;
    double *pA, *pB;
    SorFCtl *ctl;
    // b, rau, rgv once more, tiled for the kernel: tile (strip s, row j) = 6 * SF_HALF consecutive doubles at
    // coef + ((size_t)s * crows + j) * SF_CTILE (the pointer is shifted so that j is the global row index)
    const double *coef;
    long long crows;
};

// one rank's slab: rows j0..j0+2T-1 also go to the south neighbour's halo, j1-2T+1..j1 to the north one's.
// A separate argument type keeps the one-GPU kernel's parameter block (and with it its code) unchanged.
struct SorFSlabArgs : SorFArgs {
    int rank, world;
    double *nbrA[2], *nbrB[2];
    W2Mail *mail[W2_MAXRANKS];
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// explicit shared-space accesses with 32-bit addresses: keeps the per-update instruction stream free of the
// generic->shared window arithmetic (S2R SR_CgaCtaId / LEA) the compiler otherwise re-emits per access
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f64(unsigned addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(unsigned bar_addr, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(bar_addr), "r"(parity) : "memory");
}

__global__ void sorf_ctl_reset(SorFCtl *c) {
    c->done = 0; c->m = 0; c->nconv = 0; c->ticket = 0; c->cur = 0; c->redo = 0;
    for (int k = 0; k < 8; ++k) c->slot[k] = 0ull;
    c->last_dif = 0ull;
}

// End of a pass of Tp fused iterations: test `m > 1 .and. dif < sortol` (:534) for each of them in order.
__device__ __forceinline__ void sorf_close_pass(SorFCtl *ctl, int Tp, int cur, double sortol, int msorit) {
    const int m = ctl->m;
    int conv_t = -1;
    unsigned long long last = 0ull;
    for (int t = 0; t < Tp; ++t) {
        const unsigned long long bits = atomicExch(&ctl->slot[t], 0ull);
        last = bits;
        const double dif = __longlong_as_double((long long)bits);
        if (conv_t < 0 && (m + t + 1) > 1 && dif < sortol) conv_t = t;
    }
    ctl->ticket = 0;
    if (ctl->redo > 0) {                       // this was the repeat of a converged prefix
        ctl->m = m + Tp; ctl->nconv = m + Tp; ctl->done = 1; ctl->cur = cur ^ 1; ctl->redo = 0;
    } else if (conv_t == Tp - 1) {             // converged exactly at the end of the pass
        ctl->m = m + Tp; ctl->nconv = m + Tp; ctl->done = 1; ctl->cur = cur ^ 1;
    } else if (conv_t >= 0) {                  // converged mid-pass: repeat conv_t+1 iterations
        ctl->redo = conv_t + 1;
    } else {
        ctl->m = m + Tp; ctl->cur = cur ^ 1;
        if (m + Tp >= msorit) ctl->done = 1;
    }
    ctl->last_dif = last;
}

// One ring slot holds one grid row of the strip: the iterate p in the padded colour-split layout (SF_STRIDE doubles),
// followed by the row's coefficient tile -- b, rau, rgv, each [even-i half | odd-i half] of SF_HALF doubles, dense --
// exactly as it lies in the tiled coefficient array (sorf_tile_kernel), so that ONE bulk copy brings it in.
#define SF_CTILE (6 * SF_HALF)                  // doubles per coefficient tile (one strip, one row): 6 KB
#define SF_SLOT (SF_STRIDE + SF_CTILE)          // doubles per ring slot
template <int T> struct SorFCfg {
    static constexpr int R = (T == 1) ? 8 : 13;      // ring depth in rows (T=2: 105 KB -> two CTAs per SM)
    static constexpr size_t smem = (size_t)R * SF_SLOT * sizeof(double) + R * sizeof(unsigned long long);
};

// The same kernel serves one GPU (rows 2..ny, the last CTA closes the pass) and one rank's row slab (rows
// j0..j1, ext_decide != 0: the pass is closed by a follow-up kernel).  The multi-GPU plumbing lives in that
// follow-up kernel: anything added here, even dead code, was seen to perturb the streaming loop's code.
// Closing a pass on a row slab (peer-memory path): publish this slab's max-norms and the pass number in every rank's
// mailbox, wait for the same record from every rank -- a barrier across the GPUs -- and take the common decision of
// sorf_close_pass.  Run by the first warp of the slab's last CTA (sorf_body<T, true>) or of sorf_edge_kernel.
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p);
template <class ARGS>
__device__ __forceinline__ void sorf_slab_close(const ARGS &a, int T, int cur, int tid);

template <int T, bool SLAB, class ARGS>
__device__ __forceinline__ void sorf_body(const ARGS &a) {
    constexpr int NS = 2 * T;                 // half-sweep stages
    constexpr int R = SorFCfg<T>::R;
    constexpr int LIVE = 4 * T + 1;           // rows between the newest and the one being stored
    constexpr int D = R - LIVE - 1;           // prefetch distance (rows)
    constexpr int H = 2 * T;                  // halo columns/rows on a non-physical side
    constexpr int RS = R * SF_SLOT;
    static_assert(D >= 2, "ring too shallow");
    static_assert((SF_STRIDE * 8) % 16 == 0 && (SF_SLOT * 8) % 16 == 0, "bulk-copy destinations must be 16-byte aligned");

    SorFCtl *ctl = a.ctl;
    if (ctl->done) return;
    // iterations this pass: a redo pass repeats `redo` iterations, else up to T
    const int Tp = ctl->redo > 0 ? ctl->redo : min(T, a.msorit - ctl->m);
    const int cur = ctl->cur;
    const double *__restrict__ psrc = cur ? a.pB : a.pA;
    double *__restrict__ pdst = cur ? a.pA : a.pB;
    double *nbr_lo = nullptr, *nbr_hi = nullptr;   // slab runs: the neighbours' copies of pdst (null at a physical boundary)
    if constexpr (SLAB) { nbr_lo = cur ? a.nbrA[0] : a.nbrB[0]; nbr_hi = cur ? a.nbrA[1] : a.nbrB[1]; }

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sP = reinterpret_cast<double *>(smem_raw);    // ring slots: [p row | coefficient tile]
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sP + RS);
    __shared__ double red[32];

    const int tid = threadIdx.x;
    const int nx = a.nx, ny = a.ny, pitch = a.pitch;
    const int strip = blockIdx.x, band = blockIdx.y;
    // columns: strip covers global i in [i0, i0+256); i0 even
    const int i0 = strip * a.own_w;
    const bool physL = (strip == 0), physR = (i0 + SF_W - 1 >= nx + 1);
    const int own_lo = physL ? 2 : i0 + H;                             // owned columns (global i)
    const int own_hi = physR ? nx : min(nx, i0 + H + a.own_w - 1);
    // rows: band owns [jA, jB]; loads [jL0, jL1]; ring slot of a row = (row - jL0) mod R
    const int jA = a.j0 + band * a.rows_per_band;
    const int jB = min(a.j1, jA + a.rows_per_band - 1);
    const int jL0 = max(1, jA - H), jL1 = min(ny + 1, jB + H);

    if (tid == 0) {
        for (int k = 0; k < R; ++k) mbar_init(&bars[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // zero the pad cells once (read by edge threads, never used by owned cells)
    for (int k = tid; k < R * 8; k += blockDim.x) {
        const int row = k >> 3, q = k & 7;   // pads: 0,1 | 130,131 | 132,133 | 262,263
        const int off = row * SF_SLOT + (q >> 2) * SF_HSTR + ((q & 3) < 2 ? (q & 3) : SF_HALF + (q & 3));
        sP[off] = 0.0;   // (the coefficient tiles have no pads: an edge thread reads a neighbouring word of the tile instead)
    }
    __syncthreads();

    // ---- producer state (lanes 0..2 of warp 0, one bulk copy each): next row to request and its ring slot.
    // Every bulk copy costs the SM's copy engine about the same whatever its size (measured: halving the number of copies
    // of a row at equal bytes took 13 % off the pass), so a row arrives in three copies -- the two colour halves of p (1 KB
    // each; p is written by the kernel itself in the colour-split field layout) and the coefficient tile (6 KB) -- instead
    // of the eight 1 KB copies of round 2a.
    const int hp = pitch >> 1;                       // half pitch: start of the odd-i half of a row
    constexpr int NLD = 3;
    int ld_row = jL0, ld_slot = 0;
    unsigned ld_off = 0;
    const double *ld_src = nullptr;
    double *ld_dst = nullptr;
    unsigned ld_bytes = SF_HALF * 8;
    size_t ld_step = (size_t)pitch;
    // (SF_PWARP: the copies issued by the first lanes of an extra warp that does nothing else and reaches the step barrier
    // at once.  Measured slower: issuing a copy does not hold up the warp that does it; the ninth warp only costs.)
    const int ptid = SF_PWARP ? tid - NS * SF_TPS : tid;   // producer lane number (negative: not a producer)
    if (ptid >= 0 && ptid < 2) {           // p: even-i / odd-i half of the strip's columns
        ld_src = psrc + (size_t)pitch * jL0 + (i0 >> 1) + (ptid ? hp : 0);   // i0 is a multiple of 4: 16-byte aligned
        ld_dst = sP + ptid * SF_HSTR + SF_PADL;
    } else if (ptid == 2) {   // the row's coefficient tile
        ld_src = a.coef + ((size_t)strip * (size_t)a.crows + (size_t)jL0) * SF_CTILE;
        ld_dst = sP + SF_STRIDE;
        ld_bytes = SF_CTILE * 8;
        ld_step = SF_CTILE;
    }
    auto issue_row = [&]() {   // executed by lanes 0..NLD-1 together
        if (ptid == 0) mbar_expect_tx(&bars[ld_slot], (2u * SF_HALF + SF_CTILE) * 8u);
        tma_load_1d(ld_dst + ld_off, ld_src, ld_bytes, &bars[ld_slot]);
        ++ld_row; ld_src += ld_step;
        ld_off += SF_SLOT; ++ld_slot;
        if (ld_slot == R) { ld_slot = 0; ld_off = 0; }
    };
    if (ptid >= 0 && ptid < NLD)
        for (int n = 0; n <= D && ld_row <= jL1 && !(SF_DIAG & 8); ++n) issue_row();

    // ---- consumer state (all shared-memory addresses are 32-bit byte addresses)
    const int stage = tid / SF_TPS + 1;       // 1..NS, SF_TPS threads per stage
    const int kk0 = tid % SF_TPS;             // this thread's column pairs: kk0 + u*SF_TPS, u = 0..SF_CPT-1
    const int colour = (stage - 1) & 1;       // 0 = black (i+j even), 1 = red; i0 is even
    // per pair u, bits 2u / 2u+1: the even / odd column is an unknown (2..nx) resp. owned by this CTA
    unsigned vmask = 0, omask = 0;
#pragma unroll
    for (int u = 0; u < SF_CPT; ++u) {
        const int ig = i0 + 2 * (kk0 + u * SF_TPS);
        vmask |= ((unsigned)(ig >= 2 && ig <= nx) | ((unsigned)(ig + 1 >= 2 && ig + 1 <= nx) << 1)) << (2 * u);
        omask |= ((unsigned)(ig >= own_lo && ig <= own_hi) | ((unsigned)(ig + 1 >= own_lo && ig + 1 <= own_hi) << 1)) << (2 * u);
    }
    // rows this stage may relax: q in [qlo, qlo+qspan] (empty when the stage is switched off for this pass)
    const int qlo = max(2, jL0 + 1);
    const unsigned qspan = (stage <= 2 * Tp) ? (unsigned)(min(ny, jL1 - 1) - qlo) : 0u;
    const bool any_row = (stage <= 2 * Tp) && (min(ny, jL1 - 1) >= qlo);
    const unsigned jspan = (unsigned)(jB - jA);
    const double sorrel = a.sorrel;
    double lmax = 0.0;

    constexpr unsigned ROWB = SF_SLOT * 8, RINGB = RS * 8, HSTRB = SF_HSTR * 8, PAIRB = SF_TPS * 8;
    constexpr unsigned CHALFB = SF_HALF * 8;                       // one colour half of one array inside a tile
    unsigned sbase = smem_u32(smem_raw);
    asm volatile("" : "+r"(sbase));   // opaque: keeps the shared window base in a register instead of re-deriving it per row
    // byte addresses inside a slot: p at 0, then the tile: b, rau, rgv
    const unsigned aP = sbase, aB = sbase + SF_STRIDE * 8, aU = aB + 2 * CHALFB, aV = aB + 4 * CHALFB, aBar = sbase + RINGB;
    const unsigned base8 = (SF_PADL + kk0) * 8;
    const unsigned cbase8 = kk0 * 8;                               // the same pair inside a (pad-free) tile half

    int q = jL0 - 2 * stage + 1;              // row relaxed by this stage at time step r = jL0
    auto ring_off = [&](int row) { int m = (row - jL0) % R; if (m < 0) m += R; return (unsigned)m * ROWB; };
    unsigned off_s = ring_off(q - 1), off_q = ring_off(q), off_n = ring_off(q + 1);
    unsigned par = (unsigned)(colour + q) & 1u;   // column parity of the active cell in row q
    // active cell: half `par`, pair kk; west neighbour: par=0 -> odd[kk-1], par=1 -> even[kk]; east = west + 1
    unsigned ha8 = base8 + par * HSTRB;
    unsigned hw8 = par ? base8 : base8 + HSTRB - 8;
    const unsigned HA_SUM = 2 * base8 + HSTRB, HW_SUM = 2 * base8 + HSTRB - 8;
    // the same two positions in a coefficient tile (no pads; pair 0's west neighbour of an even cell is the word before
    // the odd half -- some finite coefficient, used by the strip's outermost halo cell only, whose result nobody reads)
    unsigned ca8 = cbase8 + par * CHALFB;
    unsigned cw8 = par ? cbase8 : cbase8 + CHALFB - 8;
    const unsigned CA_SUM = 2 * cbase8 + CHALFB, CW_SUM = 2 * cbase8 + CHALFB - 8;
    unsigned w_bar = aBar, w_par = 0;
#if SF_BAL_STORE
    // Store-back of a finished row, shared by ALL threads of the CTA (cell c = tid + k*blockDim of the row's SF_W cells,
    // [even-i half | odd-i half]) instead of by the two warps of the last stage: those warps were the longest path between
    // two block barriers (update + 4 loads and stores per thread), and every other warp waited for them.
    constexpr int NTH = NS * SF_TPS, SPT = SF_W / NTH;   // threads per CTA, cells stored per thread
    static_assert(SF_W % NTH == 0, "store-back mapping");
    int qs = jL0 - 2 * NS + 1;                 // the row that leaves the last stage at time step r
    unsigned off_st = ring_off(qs);
    unsigned st_own = 0, st_sm[SPT];
    double *st_g[SPT];
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
        const int cc = tid + k * NTH, half = cc / SF_HALF, kk = cc % SF_HALF, ig = i0 + 2 * kk + half;
        st_own |= (unsigned)(ig >= own_lo && ig <= own_hi) << k;
        st_sm[k] = (unsigned)(half * SF_HSTR + SF_PADL + kk) * 8u;
        st_g[k] = pdst + (size_t)pitch * qs + (i0 >> 1) + kk + (half ? hp : 0);
    }
#else
    double *gst = pdst + (size_t)pitch * q + (i0 >> 1) + kk0;   // store address of pair kk0's even cell in row q
    const bool last_stage = stage == NS;
#endif
    const int r_end = jB + 4 * T - 1;
#pragma unroll 1
    for (int r = jL0; r <= r_end; ++r) {
        if (r <= jL1 && (!SF_PWARP || ptid < 0)) {
            if (!(SF_DIAG & 8)) mbar_wait_a(w_bar, w_par);
            w_bar += 8;
            if (w_bar == aBar + 8 * R) { w_bar = aBar; w_par ^= 1u; }
        }
        if (any_row && (unsigned)(q - qlo) <= qspan) {
            const bool row_owned = (unsigned)(q - jA) <= jspan;
            // straight-line code for all SF_CPT cells of this thread: every operand is loaded and every quotient formed
            // (w2_div_fast) before anything is branched on, so the independent updates overlap; cells that are not
            // unknowns compute on whatever the ring holds and are simply not stored
            double pcv[SF_CPT], sumv[SF_CPT], a3v[SF_CPT], qdv[SF_CPT], bbv[SF_CPT];
            bool okv[SF_CPT], valid[SF_CPT];
            bool all_ok = true;
#pragma unroll
            for (int u = 0; u < SF_CPT; ++u) {
                valid[u] = (vmask >> (2 * u + par)) & 1u;
                const unsigned iq = off_q + ha8 + u * PAIRB, is = off_s + ha8 + u * PAIRB, in = off_n + ha8 + u * PAIRB,
                               iw = off_q + hw8 + u * PAIRB;
                const unsigned cq = off_q + ca8 + u * PAIRB, cs = off_s + ca8 + u * PAIRB, cw = off_q + cw8 + u * PAIRB;
                bbv[u] = lds_f64(aB + cq);
                pcv[u] = lds_f64(aP + iq);
                const double a1 = lds_f64(aV + cs), a2 = lds_f64(aU + cw), a4 = lds_f64(aU + cq), a5 = lds_f64(aV + cq);
                const double pS = lds_f64(aP + is), pW = lds_f64(aP + iw), pE = lds_f64(aP + iw + 8), pN = lds_f64(aP + in);
                a3v[u] = -a4 - a2 - a5 - a1;
                sumv[u] = bbv[u] - a1 * pS - a2 * pW - a4 * pE - a5 * pN;
                if (SF_DIAG & 1) { qdv[u] = pcv[u]; okv[u] = true; } else qdv[u] = w2_div_fast(sumv[u], a3v[u], okv[u]);
                all_ok = all_ok && (okv[u] || !valid[u]);
            }
            if (!all_ok) {   // operands outside the fast path's range (zero, tiny or huge numerators, ...)
#pragma unroll
                for (int u = 0; u < SF_CPT; ++u)
                    if (!okv[u] && valid[u]) qdv[u] = w2_div_detour(sumv[u], a3v[u]);
            }
#pragma unroll
            for (int u = 0; u < SF_CPT; ++u) {
                const unsigned iq = off_q + ha8 + u * PAIRB;
                double sum = qdv[u] - pcv[u];
                if (bbv[u] != bbv[u]) sum = 0.0 - pcv[u];   // identity row (blockage): a = (0,0,1,0,0), b = 0  (:123-137)
                if (valid[u]) sts_f64(aP + iq, pcv[u] + sorrel * sum);
                if (valid[u] && row_owned && ((omask >> (2 * u + par)) & 1u)) lmax = fmax(lmax, fabs(sum));
            }
        }
        if (!(SF_DIAG & 4)) __syncthreads();
        // the row that has just passed the last stage is final: back to HBM
#if SF_BAL_STORE
        if (!(SF_DIAG & 2) && (unsigned)(qs - jA) <= jspan && (!SF_PWARP || tid < NTH)) {
            double *nd = nullptr;
            if constexpr (SLAB) nd = (qs - a.j0 < H) ? nbr_lo : (a.j1 - qs < H) ? nbr_hi : nullptr;
#pragma unroll
            for (int k = 0; k < SPT; ++k) {
                if ((st_own >> k) & 1u) {
                    const double x = lds_f64(aP + off_st + st_sm[k]);
                    *st_g[k] = x;
                    if (SLAB && nd != nullptr) nd[st_g[k] - pdst] = x;
                }
            }
        }
        ++qs;
        off_st += ROWB;
        if (off_st == RINGB) off_st = 0;
#pragma unroll
        for (int k = 0; k < SPT; ++k) st_g[k] += pitch;
#else
        if (last_stage && (unsigned)(q - jA) <= jspan) {
            // slab runs: the first / last 2T rows of the slab are the neighbours' halo rows -- the same words go straight
            // into the neighbour's destination buffer (peer stores over NVLink), so that no copy kernel follows the pass
            double *nst = nullptr;
            if constexpr (SLAB) {
                double *nd = (q - a.j0 < H) ? nbr_lo : (a.j1 - q < H) ? nbr_hi : nullptr;
                if (nd != nullptr) nst = nd + (gst - pdst);
            }
#pragma unroll
            for (int u = 0; u < SF_CPT; ++u) {
                if ((omask >> (2 * u)) & 1u) {
                    const double x = lds_f64(aP + off_q + base8 + u * PAIRB);
                    gst[u * SF_TPS] = x;
                    if (SLAB && nst != nullptr) nst[u * SF_TPS] = x;
                }
                if ((omask >> (2 * u + 1)) & 1u) {
                    const double x = lds_f64(aP + off_q + HSTRB + base8 + u * PAIRB);
                    gst[hp + u * SF_TPS] = x;
                    if (SLAB && nst != nullptr) nst[hp + u * SF_TPS] = x;
                }
            }
        }
#endif
        // refill: the slot being overwritten held row r-4T-1, dead since the barrier above
        if (ptid >= 0 && ptid < NLD && ld_row <= jL1 && !(SF_DIAG & 8)) issue_row();
        off_s = off_q; off_q = off_n;
        off_n += ROWB;
        if (off_n == RINGB) off_n = 0;
        ha8 = HA_SUM - ha8; hw8 = HW_SUM - hw8;
        ca8 = CA_SUM - ca8; cw8 = CW_SUM - cw8;
        par ^= 1u;
        ++q;
#if !SF_BAL_STORE
        gst += pitch;
#endif
    }

    // ---- per-iteration max-norms: threads of stages 2t+1 and 2t+2 hold iteration t's partial max
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const double v = ((stage - 1) >> 1) == t ? lmax : 0.0;
        const double m = w2_block_max(v, red);
        if (tid == 0 && t < Tp) atomicMax(&ctl->slot[t], w2_dbits(m));
    }
    // ---- the last CTA closes the pass
    if constexpr (SLAB) {
        __shared__ int s_last;
        __syncthreads();
        if (tid == 0) {
            // this CTA's peer stores (bands at the slab's edges only) are visible before its ticket
            if ((jA - a.j0 < H && nbr_lo != nullptr) || (a.j1 - jB < H && nbr_hi != nullptr)) __threadfence_system();
            else __threadfence();
            s_last = atomicAdd(&ctl->ticket, 1) == (int)(gridDim.x * gridDim.y) - 1;
        }
        __syncthreads();
        if (s_last && tid < 32) sorf_slab_close(a, T, cur, tid);
    } else {
        if (tid == 0 && !a.ext_decide) {
            __threadfence();
            const int total = gridDim.x * gridDim.y;
            if (atomicAdd(&ctl->ticket, 1) == total - 1) {
                __threadfence();
                sorf_close_pass(ctl, Tp, cur, a.sortol, a.msorit);
                __threadfence();
            }
        }
    }
}

// The same code serves one GPU (rows 2..ny, the last CTA closes the pass; also the NCCL slab path, closed by a follow-up
// kernel) and one rank's row slab over peer memory (SLAB: edge rows stored to the neighbours by the pass itself, the pass
// closed across the GPUs by the slab's last CTA).  Two kernels, so that the one-GPU kernel's parameter block and code are
// not touched by the slab plumbing.
template <int T>
__global__ void __launch_bounds__(SF_NTHREADS(T), ((T == 1) ? 3 : 2) * (256 / SF_W)) sor_rb_fused_kernel(SorFArgs a) {
    sorf_body<T, false>(a);
}
template <int T>
__global__ void __launch_bounds__(SF_NTHREADS(T), ((T == 1) ? 3 : 2) * (256 / SF_W)) sor_rb_fused_slab_kernel(SorFSlabArgs a) {
    sorf_body<T, true>(a);
}

// Several GPUs: every rank's slots hold the max over its slab; after the all-reduce one thread per rank
// takes the (identical) decision the last CTA takes on one GPU.
__global__ void sorf_decide_kernel(SorFCtl *ctl, int T, double sortol, int msorit) {
    if (ctl->done) return;
    const int Tp = ctl->redo > 0 ? ctl->redo : min(T, msorit - ctl->m);
    sorf_close_pass(ctl, Tp, ctl->cur, sortol, msorit);
}

// Peer-memory path (no NCCL call inside the loop): the slab kernel stores the slab's first / last 2T rows, as they
// become final, into the neighbours' destination buffers as well (sorf_body<T, true>), and the first warp of the LAST CTA
// to finish closes the pass across the GPUs: (1) it publishes this slab's max-norms and the pass number in every rank's
// mailbox (system-scope release), (2) waits for the same record from every rank -- a barrier across the GPUs -- and
// (3) takes the common decision.  (Round 1 did (1)-(3) and the edge copy in a kernel of its own after every pass:
// two more launch gaps per pass.)
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
template <class ARGS>
__device__ __forceinline__ void sorf_slab_close(const ARGS &a, int T, int cur, int tid) {
    SorFCtl *ctl = a.ctl;
    W2Mail *me = a.mail[a.rank];
    const unsigned long long seq = me->my_seq + 1ull;
    const int par = (int)(seq & 1ull);
    if (tid == 0) {
        __threadfence_system();
        for (int r = 0; r < a.world; ++r) {
            volatile unsigned long long *dst = a.mail[r]->slot[par][a.rank];
            for (int t = 0; t < T; ++t) dst[t] = ctl->slot[t];
        }
        __threadfence_system();
        for (int r = 0; r < a.world; ++r) {
            volatile unsigned long long *fl = &a.mail[r]->seq[par][a.rank];
            *fl = seq;
        }
        me->my_seq = seq;
        ctl->ticket = 0;
    }
    __syncwarp();
    bool ok = true;
    if (tid < a.world) {
        const long long t0 = clock64();
        while (ld_acquire_sys(&me->seq[par][tid]) != seq)
            if (clock64() - t0 > 20000000000ll) { ok = false; break; }   // ~10 s: a peer is gone
    }
    ok = __all_sync(0xffffffffu, ok);
    if (tid != 0) return;
    if (!ok) { me->timeout = 1; ctl->done = 1; return; }
    const int Tp = ctl->redo > 0 ? ctl->redo : min(T, a.msorit - ctl->m);
    for (int t = 0; t < Tp; ++t) {
        unsigned long long m = 0ull;
        for (int r = 0; r < a.world; ++r) {
            const unsigned long long v = ld_acquire_sys(&me->slot[par][r][t]);
            m = v > m ? v : m;
        }
        ctl->slot[t] = m;
    }
    sorf_close_pass(ctl, Tp, cur, a.sortol, a.msorit);
}

// The follow-up kernel of the default peer-memory path (option "sor_slab_inpass" 0): after the unchanged one-GPU pass
// kernel, CTA (strip, side) copies its columns of the slab's first / last 2T rows (just written, still in L2) into the
// neighbour's destination buffer with plain stores over NVLink; the last CTA to finish closes the pass across the GPUs.
__global__ void __launch_bounds__(256) sorf_edge_kernel(SorFSlabArgs a, int T) {
    SorFCtl *ctl = a.ctl;
    if (ctl->done) return;
    __shared__ int s_last;
    const int H = 2 * T, tid = threadIdx.x;
    const int cur = ctl->cur;                       // the pass just run wrote the other buffer
    const double *pdst = cur ? a.pA : a.pB;
    const int sd = blockIdx.y;                      // 0: rows j0.. to rank-1, 1: rows ..j1 to rank+1
    double *nd = cur ? a.nbrA[sd] : a.nbrB[sd];
    if (nd != nullptr) {
        const int i0 = blockIdx.x * a.own_w;
        const bool physL = blockIdx.x == 0, physR = (i0 + SF_W - 1 >= a.nx + 1);
        const int own_lo = physL ? 2 : i0 + H;
        const int own_hi = physR ? a.nx : min(a.nx, i0 + H + a.own_w - 1);
        const int ow = own_hi - own_lo + 1, hp = a.pitch >> 1;
        const int row0 = sd ? a.j1 - H + 1 : a.j0;
        for (int k = tid; k < H * ow; k += blockDim.x) {
            const int rr = k / ow, i = own_lo + (k - rr * ow);
            const size_t off = (size_t)a.pitch * (size_t)(row0 + rr) + (size_t)(i & 1) * hp + (size_t)(i >> 1);
            nd[off] = pdst[off];
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();                     // this CTA's peer stores are visible before its ticket
        s_last = atomicAdd(&ctl->ticket, 1) == (int)(gridDim.x * gridDim.y) - 1;
    }
    __syncthreads();
    if (s_last && tid < 32) sorf_slab_close(a, T, cur, tid);
}

// Start-of-solve barrier: tells every rank that this rank's two pressure buffers are initialised (so peer
// stores into their halo rows may begin) and waits for the same word from the others.
__global__ void sorf_ready_p2p(SorFSlabArgs a, unsigned long long solve) {
    const int lane = threadIdx.x;
    W2Mail *me = a.mail[a.rank];
    bool ok = true;
    if (lane < a.world) {
        __threadfence_system();
        volatile unsigned long long *fl = &a.mail[lane]->ready[a.rank];
        *fl = solve;
        const long long t0 = clock64();
        while (ld_acquire_sys(&me->ready[lane]) < solve)
            if (clock64() - t0 > 20000000000ll) { ok = false; break; }
    }
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0 && !ok) { me->timeout = 1; a.ctl->done = 1; }
}

// interleaved (reference order, i fastest) <-> colour-split rows.  TO_SPLIT: dst[j][ (i&1)*hp + i/2 ] = src[j][i]
template <bool TO_SPLIT>
__global__ void __launch_bounds__(256) sorf_pack_kernel(int pitch, int rows, const double *__restrict__ src,
                                                        double *__restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pitch) return;
    const int hp = pitch >> 1;
    const int s = (i & 1) * hp + (i >> 1);
    for (int j = blockIdx.y; j < rows; j += gridDim.y) {
        const size_t r = (size_t)pitch * j;
        if (TO_SPLIT) dst[r + s] = src[r + i];
        else dst[r + i] = src[r + s];
    }
}

// colour-split -> interleaved, the source being the buffer the control block names as the current iterate
__global__ void __launch_bounds__(256) sorf_unpack_cur_kernel(int pitch, int rows, const double *__restrict__ pA,
                                                              const double *__restrict__ pB, const SorFCtl *__restrict__ ctl,
                                                              double *__restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pitch) return;
    const double *__restrict__ src = ctl->cur ? pB : pA;
    const int hp = pitch >> 1;
    const int s = (i & 1) * hp + (i >> 1);
    for (int j = blockIdx.y; j < rows; j += gridDim.y) {
        const size_t r = (size_t)pitch * j;
        dst[r + i] = src[r + s];
    }
}

// Coefficient tiles (SorFArgs::coef) from the colour-split arrays: tile (strip s, row j) holds the strip's SF_HALF column
// pairs of b, rau, rgv, each as [even-i half | odd-i half] -- the six 1 KB pieces the kernel of round 2a fetched with six
// bulk copies, laid side by side so that one copy fetches them.  Strips overlap by their halo columns, so those are stored
// twice (SF_W / own_w - 1 = 3 % more bytes).  which: bit 0 = b (every solve), bits 1, 2 = rau, rgv (after an upload).
__global__ void __launch_bounds__(SF_W) sorf_tile_kernel(int pitch, int jlo, int jhi, int own_w, long long crows, int which,
                                                         const double *__restrict__ b, const double *__restrict__ rau,
                                                         const double *__restrict__ rgv, double *__restrict__ coef) {
    const int s = blockIdx.x, half = threadIdx.x / SF_HALF, k = threadIdx.x % SF_HALF;
    const size_t col = (size_t)((s * own_w) >> 1) + (size_t)k + (half ? (size_t)(pitch >> 1) : 0);
    for (int j = jlo + blockIdx.y; j <= jhi; j += gridDim.y) {
        double *t = coef + ((size_t)s * (size_t)crows + (size_t)j) * SF_CTILE + half * SF_HALF + k;
        const size_t src = (size_t)pitch * (size_t)j + col;
        if (which & 1) t[0] = b[src];
        if (which & 2) t[2 * SF_HALF] = rau[src];
        if (which & 4) t[4 * SF_HALF] = rgv[src];
    }
}

int w2_sorf_pack(wolfd2_ctx *c, const double *src, double *dst, bool to_split) {
    const int rows = c->rows + 1;
    dim3 grid((c->pitch + 255) / 256, rows < 2048 ? rows : 2048);
    src += c->row_off; dst += c->row_off;   // the kernel walks the held rows from 0
    if (to_split) sorf_pack_kernel<true><<<grid, 256, 0, c->stream>>>(c->pitch, rows, src, dst);
    else sorf_pack_kernel<false><<<grid, 256, 0, c->stream>>>(c->pitch, rows, src, dst);
    c->launches[2]++;
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}

// ---------------------------------------------------------------------------------------- host

// option "sor_slab_inpass": 1 = on several GPUs the pass kernel itself stores the edge rows to the neighbours and closes the
// pass across the GPUs (sor_rb_fused_slab_kernel: no follow-up launch); 0 (default) = the one-GPU pass kernel followed by
// sorf_edge_kernel.  Both bit-identical to one GPU (multi_gpu_worker.py with W2_OPTS=sor_slab_inpass=0/1, bench.py's verify
// block).  Two B200s, 4096 x 8192, 50 passes per step, PPE section: round-2a kernel 10.53 ms (edge kernel) / 10.20 ms (in
// pass); with the tiled coefficients 9.53 / 9.75 ms (one GPU on 4096^2: 8.84) -- the slab kernel's 12 extra registers cost
// the streaming loop more than the edge kernel's launch.
int g_sor_slab_inpass = 0;

template <int T>
static int launch_fused(wolfd2_ctx *c, const SorFArgs &a, dim3 grid) {
    const size_t smem = SorFCfg<T>::smem;
    static bool attr_set[W2_MAXDEV] = {};   // per device
    if (!attr_set[c->device % W2_MAXDEV]) {
        W2_CUDA(cudaFuncSetAttribute(sor_rb_fused_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[c->device % W2_MAXDEV] = true;
    }
    sor_rb_fused_kernel<T><<<grid, SF_NTHREADS(T), smem, c->stream>>>(a);
    return W2_OK;
}
template <int T>
static int launch_fused_slab(wolfd2_ctx *c, const SorFSlabArgs &a, dim3 grid) {
    const size_t smem = SorFCfg<T>::smem;
    static bool attr_set[W2_MAXDEV] = {};   // per device
    if (!attr_set[c->device % W2_MAXDEV]) {
        W2_CUDA(cudaFuncSetAttribute(sor_rb_fused_slab_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[c->device % W2_MAXDEV] = true;
    }
    sor_rb_fused_slab_kernel<T><<<grid, SF_NTHREADS(T), smem, c->stream>>>(a);
    return W2_OK;
}
template <int T>
static int fused_occupancy(int *per_sm) {
    const size_t smem = SorFCfg<T>::smem;
    cudaFuncSetAttribute(sor_rb_fused_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, sor_rb_fused_kernel<T>, SF_NTHREADS(T), smem);
    return W2_OK;
}

// Runs the whole SOR loop of SorRB on p (Cartesian grids).  b (W2_F_B) must already hold div/dk in the
// colour-split layout with the NaN sentinel at identity rows.  p is packed into the split layout, iterated
// between two split buffers, and unpacked at the end; the colour-split copies of rau and rgv are kept until
// something is uploaded into those arrays (w2_upload2d clears sorf_met_valid).
int w2_sor_fused(wolfd2_ctx *c, double *p, double *scratch, int T, int *nSorConv, int *converged, double **p_final,
                 int *iters_done) {
    for (int k = 0; k < 4; ++k)
        if (!c->sorf_buf[k]) W2_TRY(w2_alloc_field(c, &c->sorf_buf[k]));
    double *pA = c->sorf_buf[0], *pB = c->sorf_buf[1], *rauS = c->sorf_buf[2], *rgvS = c->sorf_buf[3];
    (void)scratch;
    W2_TRY(w2_sorf_pack(c, p, pA, true));
    if (!c->sorf_met_valid) {   // metric-only: repacked after an upload into rau / rgv (w2_upload2d), not per solve
        W2_TRY(w2_sorf_pack(c, c->met.rau, rauS, true));
        W2_TRY(w2_sorf_pack(c, c->met.rgv, rgvS, true));
        c->sorf_met_valid = 1;
        c->sorf_coef_T = 0;   // the tiles' rau / rgv parts are stale
    }
    const wolfd2_params &par = c->par;
    const int nx = c->nx, ny = c->ny;
    SorFCtl *ctl = (SorFCtl *)c->d_flags;
    static_assert(sizeof(SorFCtl) <= 32 * sizeof(int), "ctl block too large");
    SorFSlabArgs a;
    a.nx = nx; a.ny = ny; a.pitch = c->pitch;
    a.j0 = c->J0; a.j1 = c->J1; a.ext_decide = c->world > 1;
    a.rank = c->rank; a.world = c->world;
    memset(a.nbrA, 0, sizeof(a.nbrA)); memset(a.nbrB, 0, sizeof(a.nbrB)); memset(a.mail, 0, sizeof(a.mail));
    if (c->world > 1) {
        W2_TRY(w2_peer_setup(c));
        if (c->peer.state == 1) {
            a.ext_decide = 2;
            for (int sd = 0; sd < 2; ++sd) { a.nbrA[sd] = c->peer.nbrA[sd]; a.nbrB[sd] = c->peer.nbrB[sd]; }
            for (int r = 0; r < c->world; ++r) a.mail[r] = c->peer.mail[r];
        }
    }
    const int nrows = c->J1 - c->J0 + 1;
    a.own_w = SF_W - 4 * T;
    a.nstrips = 1;
    while ((a.nstrips - 1) * a.own_w + SF_W < nx + 2) a.nstrips++;
    {   // coefficient tiles: rows A0..A1 of every strip (allocated for the narrower strips of T = 2, which are more)
        const long long crows = c->rows + 1;
        int smax = 1;
        while ((smax - 1) * (SF_W - 8) + SF_W < nx + 2) smax++;
        if (!c->sorf_coef) {
            W2_CUDA(cudaMalloc((void **)&c->sorf_coef, (size_t)smax * (size_t)crows * SF_CTILE * sizeof(double)));
            c->sorf_coef_T = 0;
        }
        double *coef = c->sorf_coef - (size_t)c->A0 * SF_CTILE;   // global row index, like every field pointer
        const int which = 1 | (c->sorf_coef_T != T ? 6 : 0);
        dim3 tg(a.nstrips, c->rows < 1024 ? c->rows : 1024);
        sorf_tile_kernel<<<tg, SF_W, 0, c->stream>>>(c->pitch, c->A0, c->A1, a.own_w, crows, which, c->fld[W2_F_B], rauS, rgvS, coef);
        W2_CUDA(cudaGetLastError());
        c->launches[2]++;
        c->sorf_coef_T = T;
        a.coef = coef; a.crows = crows;
    }
    // bands: exactly one resident wave of CTAs (a partial second wave would double the pass time)
    int per_sm = 1;
    if (T == 1) fused_occupancy<1>(&per_sm);
    else fused_occupancy<2>(&per_sm);
    if (per_sm < 1) per_sm = 1;
    int want = (per_sm * c->num_sms) / a.nstrips;
    if (want < 1) want = 1;
    // (Measured and rejected: half as many bands of twice the height on small grids, where the 4T halo rows and the
    // pipeline fill are a large share of an 18-row band -- 1024^2: 31.4 -> 37.5 us per pass, 2048^2: 70.7 -> 95.5 us.
    // Two resident CTAs per SM hide the per-row block barrier of each other; one tall CTA cannot.)
    a.rows_per_band = (nrows + want - 1) / want;
    if (a.rows_per_band < 8 * T) a.rows_per_band = 8 * T;
    a.nbands = (nrows + a.rows_per_band - 1) / a.rows_per_band;
    a.msorit = par.msorit; a.sorrel = par.sorrel; a.sortol = par.sortol;
    a.rau = rauS; a.rgv = rgvS; a.b = c->fld[W2_F_B];
    a.pA = pA; a.pB = pB; a.ctl = ctl;
    dim3 grid(a.nstrips, a.nbands);

    // the ghost ring of p is frozen during the solve (:431-446 never touches it): give the second
    // buffer the same ghosts
    W2_TRY(w2_copy_field(c, pB, pA));
    sorf_ctl_reset<<<1, 1, 0, c->stream>>>(ctl);
    if (a.ext_decide == 2) {
        sorf_ready_p2p<<<1, 32, 0, c->stream>>>(a, ++c->peer.solves);
        c->launches[2]++;
    }
    // Passes are enqueued WITHOUT waiting for their outcome: every pass kernel returns at once when the control block
    // says done, so over-enqueuing costs a few microseconds per launch.  The host only throttles itself: chunk k goes
    // out once the control block as it stood after chunk k-2 has arrived, which also tells it when to stop.  Nothing
    // is read back at the end: the unpack kernel picks the final buffer from the control block on the device, and
    // the outcome (iterations, convergence) is copied to pinned memory for w2_sor_collect.
    const int passes_max = (par.msorit + T - 1) / T + 1;     // + one repeat pass (mid-pass convergence, at most once)
    const double cells = (double)(nx - 1) * (double)nrows;
    int chunk = (int)(2.0e-3 / (cells * 40.0 / 5.0e12 + (c->world > 1 ? 4.0e-5 : 4.0e-6)));
    if (chunk < 32) chunk = 32;   // a pass launched after the solve has finished costs ~2 us: 64 of them are cheaper than a poll per 2 ms
    if (chunk > 128) chunk = 128;
    if (!c->ev_sor[0]) for (int k = 0; k < 2; ++k) W2_CUDA(cudaEventCreateWithFlags(&c->ev_sor[k], cudaEventDisableTiming));
    int queued = 0;
    for (int k = 0; queued < passes_max; ++k) {
        if (k >= 2) {
            W2_CUDA(cudaEventSynchronize(c->ev_sor[k & 1]));
            c->host_syncs++;
            if (((const SorFCtl *)(c->h_sor + 32 + 32 * (k & 1)))->done) break;
        }
        const int n = chunk < passes_max - queued ? chunk : passes_max - queued;
        for (int q = 0; q < n; ++q) {
            if (a.ext_decide == 2 && g_sor_slab_inpass) {
                // slab run over peer memory, in-pass variant: the pass stores its edge rows to the neighbours and closes
                // itself across the GPUs
                if (T == 1) W2_TRY(launch_fused_slab<1>(c, a, grid));
                else W2_TRY(launch_fused_slab<2>(c, a, grid));
            } else {
                if (T == 1) W2_TRY(launch_fused<1>(c, a, grid));
                else W2_TRY(launch_fused<2>(c, a, grid));
            }
            c->launches[2]++;
            if (a.ext_decide == 2 && !g_sor_slab_inpass) {
                // slab run over peer memory: edge rows and max-norms go to the peers, then the common decision
                sorf_edge_kernel<<<dim3(a.nstrips, 2), 256, 0, c->stream>>>(a, T);
                c->launches[2]++;
            }
            if (a.ext_decide == 1) {
                // slab run: global max-norms, the common decision, then the 2T halo rows of the iterate.
                // Which buffer was written is known on the device only; the other one's halos are already
                // right, so both are exchanged (a few hundred KB).
                W2_TRY(w2_allreduce_max_u64(c, ctl->slot, T));
                sorf_decide_kernel<<<1, 1, 0, c->stream>>>(ctl, T, par.sortol, par.msorit);
                c->launches[2]++;
                double *pp[2] = {pA, pB};
                W2_TRY(w2_halo_exchange(c, pp, 2, 2 * T));
            }
        }
        queued += n;
        W2_CUDA(cudaGetLastError());
        if (queued < passes_max) {
            W2_CUDA(cudaMemcpyAsync(c->h_sor + 32 + 32 * (k & 1), ctl, sizeof(SorFCtl), cudaMemcpyDeviceToHost, c->stream));
            W2_CUDA(cudaEventRecord(c->ev_sor[k & 1], c->stream));
        }
    }
    {   // back to the reference order, from whichever buffer holds the final iterate
        const int rows = c->rows + 1;
        dim3 g((c->pitch + 255) / 256, rows < 2048 ? rows : 2048);
        sorf_unpack_cur_kernel<<<g, 256, 0, c->stream>>>(c->pitch, rows, pA + c->row_off, pB + c->row_off, ctl, p + c->row_off);
        c->launches[2]++;
        W2_CUDA(cudaGetLastError());
    }
    W2_CUDA(cudaMemcpyAsync(c->h_sor, ctl, sizeof(SorFCtl), cudaMemcpyDeviceToHost, c->stream));
    (void)nSorConv; (void)converged; (void)iters_done;
    *p_final = p;
    return W2_OK;
}

// Outcome of the last fused solve; the stream must have been synchronised since.
int w2_sor_fused_result(wolfd2_ctx *c, int *nSorConv, int *converged, int *iters_done) {
    SorFCtl h;
    memcpy(&h, c->h_sor, sizeof(SorFCtl));
    if (!h.done) { w2_set_error("fused SOR did not terminate"); return W2_ERR_CUDA; }
    if (c->world > 1 && c->peer.state == 1) {
        int to = 0;
        W2_CUDA(cudaMemcpy(&to, &c->peer.mail[c->rank]->timeout, sizeof(int), cudaMemcpyDeviceToHost));
        if (to) { w2_set_error("fused SOR: timed out waiting for a peer GPU (rank %d of %d)", c->rank, c->world); return W2_ERR_CUDA; }
        W2_CUDA(cudaMemcpy(&to, &c->peer.bcg[c->rank]->timeout, sizeof(int), cudaMemcpyDeviceToHost));
        if (to) { w2_set_error("OUTLT2 ghost fill: timed out waiting for a peer GPU (rank %d of %d)", c->rank, c->world); return W2_ERR_CUDA; }
    }
    if (converged) *converged = h.nconv > 0;
    if (nSorConv) *nSorConv = h.nconv > 0 ? h.nconv : c->par.msorit;
    if (iters_done) *iters_done = h.m;
    return W2_OK;
}
