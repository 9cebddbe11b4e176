// w2_ppe_other.cu -- the four remaining ppe_solver ids of Ppe (src/pressure.f:201-240):
//   1 `sor`          Sor     (:384-450)  lexicographic point SOR
//   2 `lsor`         Slor    (:673-802)  lexicographic line SOR, lines along i (ndir = 1, :209)
//   3 `rb_lsor`      SlorRB  (:819-959)  red/black line SOR (even lines first)
//   4 `par_rb_lsor`  SlorRBP (:976-1138) same sweep, relaxation after the solve, test on |pn - p|
//
// Line SOR: all lines of one colour are independent tridiagonal systems of nx-1 unknowns; they are
// assembled back to back and solved by the partitioned solver in its batched-lines mode
// (w2_tri_solve_lines; AltTridLU's first-row quirk applied per line, :776).  The lexicographic variants
// keep the reference's data dependences exactly: id 2 solves line after line; id 1 walks the
// anti-diagonals i+j = const, on which all points are independent, in one cooperative kernel with a
// grid barrier per diagonal.  They are exact but latency-bound by construction; rb_sor (id 5/6) is the
// solver to use on the device.
#include <cooperative_groups.h>

#include "w2.cuh"

namespace cg = cooperative_groups;

#define P(i, j) p[IDX(i, j)]

struct LineArgs {
    int nx, ny, pitch;
    int j_first, j_step, nlines;   // lines j = j_first + j_step*l, l = 0..nlines-1
    double sorrel;
    const double *rau, *rgv, *b;
    const unsigned char *mask;
    int has_mask;
    double *p;
    double *ta, *td, *tc, *tb;
    const double *x;
    unsigned long long *slot;
    int relaxed_norm;              // 1: dif = |pold - pnew| (SlorRBP :1125), 0: dif = |x - pold| (:780-783)
};

// aline/bline of one line (:766-770): aline = (a2, a3, a4), bline = b - (a1*p(i,j-1) + a5*p(i,j+1))
__global__ void __launch_bounds__(256) line_assemble_kernel(LineArgs a) {
    const int pitch = a.pitch;
    const int len = a.nx - 1;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= len) return;
    const int i = k + 2;
    for (int l = blockIdx.y; l < a.nlines; l += gridDim.y) {
        const int j = a.j_first + a.j_step * l;
        const size_t c0 = IDX(i, j);
        double a1 = a.rgv[c0 - pitch], a2 = a.rau[c0 - 1], a4 = a.rau[c0], a5 = a.rgv[c0];
        double a3 = -a4 - a2 - a5 - a1;
        if (a.has_mask && a.mask[c0]) { a1 = 0.0; a2 = 0.0; a3 = 1.0; a4 = 0.0; a5 = 0.0; }
        const size_t e = (size_t)l * len + k;
        a.ta[e] = a2; a.td[e] = a3; a.tc[e] = a4;
        a.tb[e] = a.b[c0] - (a1 * a.p[c0 - pitch] + a5 * a.p[c0 + pitch]);
    }
}

// over-relaxation and max-norm (:779-790)
__global__ void __launch_bounds__(256) line_relax_kernel(LineArgs a) {
    __shared__ double red[32];
    const int pitch = a.pitch;
    const int len = a.nx - 1;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    double lmax = 0.0;
    if (k < len) {
        const int i = k + 2;
        for (int l = blockIdx.y; l < a.nlines; l += gridDim.y) {
            const int j = a.j_first + a.j_step * l;
            const double pold = a.p[IDX(i, j)];
            const double sum = a.x[(size_t)l * len + k] - pold;
            const double pnew = pold + a.sorrel * sum;
            a.p[IDX(i, j)] = pnew;
            lmax = fmax(lmax, a.relaxed_norm ? fabs(pold - pnew) : fabs(sum));
        }
    }
    lmax = w2_block_max(lmax, red);
    if (threadIdx.x == 0) atomicMax(a.slot, w2_dbits(lmax));
}

void rhs_cross_launch(wolfd2_ctx *c, double *p);   // w2_ppe.cu

static int line_pass(wolfd2_ctx *c, LineArgs &a) {
    if (a.nlines <= 0) return W2_OK;
    const int len = a.nx - 1;
    int gy = a.nlines < 1024 ? a.nlines : 1024;
    dim3 grid((len + 255) / 256, gy);
    line_assemble_kernel<<<grid, 256, 0, c->stream>>>(a);
    c->launches[2]++;
    W2_TRY(w2_tri_solve_lines(c, a.nlines, len, a.ta, a.td, a.tc, a.tb, c->tx, 1));
    line_relax_kernel<<<grid, 256, 0, c->stream>>>(a);
    c->launches[2]++;
    W2_CUDA(cudaGetLastError());
    return W2_OK;
}

// ids 2, 3, 4
int w2_ppe_line_sor(wolfd2_ctx *c, double *p, int *nSorConv, int *converged, int *iters_done) {
    const wolfd2_params &par = c->par;
    const int nx = c->nx, ny = c->ny;
    const int cart = par.lCartesGrid != 0;
    LineArgs a;
    a.nx = nx; a.ny = ny; a.pitch = c->pitch; a.sorrel = par.sorrel;
    a.rau = c->met.rau; a.rgv = c->met.rgv; a.b = c->fld[W2_F_B];
    a.mask = c->pmask; a.has_mask = c->hreg.has_blockage;
    W2_TRY(w2_ensure_chain(c));
    a.p = p; a.ta = c->ta; a.td = c->td; a.tc = c->tc; a.tb = c->tb; a.x = c->tx;
    a.slot = c->d_norm + 12;
    a.relaxed_norm = par.nPpeSolver == W2_PPE_PAR_RB_LSOR;
    *converged = 0; *nSorConv = par.msorit; *iters_done = 0;
    for (int m = 1; m <= par.msorit; ++m) {
        W2_CUDA(cudaMemsetAsync(a.slot, 0, sizeof(unsigned long long), c->stream));
        if (par.nPpeSolver == W2_PPE_LSOR) {
            if (!cart) rhs_cross_launch(c, p);                    // :741-743
            for (int j = 2; j <= ny; ++j) {                       // :747, one line after the other
                a.j_first = j; a.j_step = 1; a.nlines = 1;
                W2_TRY(line_pass(c, a));
            }
        } else {
            for (int k = 2; k <= 3; ++k) {                        // :893 / :1050,1092: even lines, then odd
                if (!cart) rhs_cross_launch(c, p);                // :897-899 / :1039-1041, :1081-1083
                a.j_first = k; a.j_step = 2; a.nlines = (ny - k) / 2 + 1;
                if (k > ny) a.nlines = 0;
                W2_TRY(line_pass(c, a));
            }
        }
        double dif = 0.0;
        W2_CUDA(cudaMemcpyAsync(c->h_norm + 12, a.slot, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        W2_CUDA(cudaStreamSynchronize(c->stream));
        memcpy(&dif, c->h_norm + 12, 8);
        *iters_done = m;
        if (m > 1 && dif < par.sortol) { *converged = 1; *nSorConv = m; break; }   // :794 / :951 / :1130
    }
    return W2_OK;
}

// ------------------------------------------------------------------ id 1: lexicographic point SOR
struct LexArgs {
    int nx, ny, pitch, msorit, cart, has_mask;
    double sorrel, sortol, dk;
    const double *rau, *rgv, *rbu, *rbv, *div;
    double *b;
    const unsigned char *mask;
    double *p;
    unsigned long long *slots;   // 3 rotating max-norm slots
    int *result;                 // [0] nconv, [1] iterations run
};

__global__ void __launch_bounds__(256) sor_lex_kernel(LexArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double red[32];
    const int pitch = a.pitch, nx = a.nx, ny = a.ny;
    double *p = a.p;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    int nconv = 0, m = 1;
    for (; m <= a.msorit; ++m) {
        if (!a.cart) {   // RhsPpe every iteration (:425-427)
            for (long long q = tid; q < (long long)(nx - 1) * (ny - 1); q += nth) {
                const int j = 2 + (int)(q / (nx - 1)), i = 2 + (int)(q % (nx - 1));
                double bb = a.div[IDX(i, j)] / a.dk;
                bb = bb - (a.rbu[IDX(i, j)] * (P(i + 1, j + 1) + P(i, j + 1) - P(i + 1, j - 1) - P(i, j - 1))
                           - a.rbu[IDX(i - 1, j)] * (P(i, j + 1) + P(i - 1, j + 1) - P(i, j - 1) - P(i - 1, j - 1))
                           + a.rbv[IDX(i, j)] * (P(i + 1, j + 1) + P(i + 1, j) - P(i - 1, j + 1) - P(i - 1, j))
                           - a.rbv[IDX(i, j - 1)] * (P(i + 1, j) + P(i + 1, j - 1) - P(i - 1, j) - P(i - 1, j - 1)));
                a.b[IDX(i, j)] = bb;
            }
            grid.sync();
        }
        double lmax = 0.0;
        // the sweep j=2..ny, i=2..nx (:431-441) visits (i,j) after (i-1,j) and (i,j-1) and before
        // (i+1,j) and (i,j+1): points on one anti-diagonal are independent
        for (int s = 4; s <= nx + ny; ++s) {
            const int jlo = max(2, s - nx), jhi = min(ny, s - 2);
            for (long long q = tid; q <= jhi - jlo; q += nth) {
                const int j = jlo + (int)q, i = s - j;
                const size_t c0 = IDX(i, j);
                double a1 = a.rgv[c0 - pitch], a2 = a.rau[c0 - 1], a4 = a.rau[c0], a5 = a.rgv[c0];
                double a3 = -a4 - a2 - a5 - a1;
                if (a.has_mask && a.mask[c0]) { a1 = 0.0; a2 = 0.0; a3 = 1.0; a4 = 0.0; a5 = 0.0; }
                const double pc = p[c0];
                double sum = a.b[c0] - a1 * p[c0 - pitch] - a2 * p[c0 - 1] - a4 * p[c0 + 1] - a5 * p[c0 + pitch];
                sum = w2_div_exact(sum, a3) - pc;
                p[c0] = pc + a.sorrel * sum;
                lmax = fmax(lmax, fabs(sum));
            }
            grid.sync();
        }
        lmax = w2_block_max(lmax, red);
        unsigned long long *slot = a.slots + (m % 3);
        if (threadIdx.x == 0) atomicMax(slot, w2_dbits(lmax));
        if (tid == 0) a.slots[(m + 1) % 3] = 0ull;
        grid.sync();
        const double dif = __longlong_as_double((long long)*((volatile unsigned long long *)slot));
        if (m > 1 && dif < a.sortol) { nconv = m; break; }   // :443-446
    }
    if (tid == 0) { a.result[0] = nconv; a.result[1] = nconv ? nconv : a.msorit; }
}

int w2_ppe_lex_sor(wolfd2_ctx *c, double *p, int *nSorConv, int *converged, int *iters_done) {
    const wolfd2_params &par = c->par;
    if (!c->coop_ok) { w2_set_error("ppe_solver sor needs cooperative launch support"); return W2_ERR_UNSUPPORTED; }
    LexArgs a;
    a.nx = c->nx; a.ny = c->ny; a.pitch = c->pitch; a.msorit = par.msorit;
    a.cart = par.lCartesGrid != 0; a.has_mask = c->hreg.has_blockage;
    a.sorrel = par.sorrel; a.sortol = par.sortol; a.dk = par.dk;
    a.rau = c->met.rau; a.rgv = c->met.rgv; a.rbu = c->met.rbu; a.rbv = c->met.rbv; a.div = c->div;
    a.b = c->fld[W2_F_B]; a.mask = c->pmask; a.p = p;
    a.slots = c->d_norm + 12; a.result = c->d_flags + 32;
    W2_CUDA(cudaMemsetAsync(a.slots, 0, 3 * sizeof(unsigned long long), c->stream));
    int per_sm = 0;
    W2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sor_lex_kernel, 256, 0));
    if (per_sm < 1) per_sm = 1;
    // no more threads than the longest anti-diagonal needs
    int blocks = per_sm * c->num_sms;
    const int need = ((c->nx < c->ny ? c->nx : c->ny) + 255) / 256;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    void *args[] = {&a};
    W2_CUDA(cudaLaunchCooperativeKernel((void *)sor_lex_kernel, dim3(blocks), dim3(256), args, 0, c->stream));
    c->launches[2]++;
    W2_CUDA(cudaMemcpyAsync(c->h_flags + 32, a.result, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    W2_CUDA(cudaStreamSynchronize(c->stream));
    const int nconv = c->h_flags[32];
    *converged = nconv > 0;
    *nSorConv = nconv > 0 ? nconv : par.msorit;
    *iters_done = c->h_flags[33];
    return W2_OK;
}
