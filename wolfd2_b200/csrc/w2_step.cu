// w2_step.cu -- the time-step body of program wolfd2 (src/main.f:690-981, cold flow) and the
// cold-start projection (src/main.f:606-641) on the device, plus the literal gfortran-ABI shims.
#include <stdlib.h>
#include <string.h>

#include "w2.cuh"

// ------------------------------------------------------------------------------ residency API

// slab runs: Project (like the QL update and Filter) changes owned rows only; refresh the halo rows
static int exchange_uv(wolfd2_ctx *c, double *u, double *v) {
    if (c->world == 1) return W2_OK;
    double *f[2] = {u, v};
    return w2_halo_exchange(c, f, 2, c->HG);
}

extern "C" int wolfd2_b200_coldstart(wolfd2_ctx *c, int32_t *nSorConv) {
    if (!c) return W2_ERR_BAD_ARG;
    W2_CUDA(cudaSetDevice(c->device));
    double *u = c->fld[W2_F_U], *v = c->fld[W2_F_V], *p = c->fld[W2_F_P];
    int nconv = 0, conv = 0;
    W2_TRY(w2_vel_bc(c, u, v));                    // :608
    W2_TRY(w2_ppe(c, u, v, p, &nconv, &conv));     // :613
    W2_TRY(w2_pres_bc(c, p));                      // :623
    W2_TRY(w2_project(c, p, u, v));                // :629
    W2_TRY(exchange_uv(c, u, v));
    W2_TRY(w2_vel_bc(c, u, v));                    // :637
    W2_CUDA(cudaStreamSynchronize(c->stream));
    if (nSorConv) *nSorConv = nconv;
    return W2_OK;
}

static int one_step(wolfd2_ctx *c, wolfd2_step_log *log) {
    double *u = c->fld[W2_F_U], *v = c->fld[W2_F_V], *p = c->fld[W2_F_P], *t = c->fld[W2_F_T];
    double *us = c->fld[W2_F_US], *vs = c->fld[W2_F_VS], *ts = c->fld[W2_F_TS];
    double *un = c->fld[W2_F_UN], *vn = c->fld[W2_F_VN], *pn = c->fld[W2_F_PN], *tn = c->fld[W2_F_TN];
    const bool thermal = c->th.nthermen == 1;
    const bool atd = c->atd && c->atd->ss.nsmallscl == 1;
    cudaStream_t s = c->stream;
    cudaEventRecord(c->ev[0], s);
    // Cold flow without the ATD model runs ONE momentum-energy pass, during which nothing reads u, v: un <- u (:696-704) and
    // u <- us (:864-870) are then exchanges of the buffer pointers instead of copies (every user takes c->fld[...] at call
    // time), and the only copy left is us <- un.  Four of the six field copies of a step: 64 of its 96 copied bytes per cell.
    const bool swap_uv = !thermal && !atd && c->par.nmeiter > 0;
    auto swap_fld = [&](int a, int b) { double *t_ = c->fld[a]; c->fld[a] = c->fld[b]; c->fld[b] = t_; };
    // :696-704  time-level n copies (without the thermal energy equation t is identically zero; d only
    // changes through EqState, so dn is refreshed only then)
    if (swap_uv) {
        swap_fld(W2_F_U, W2_F_UN); swap_fld(W2_F_V, W2_F_VN);
        u = c->fld[W2_F_U]; v = c->fld[W2_F_V]; un = c->fld[W2_F_UN]; vn = c->fld[W2_F_VN];
    } else {
        W2_TRY(w2_copy_field(c, un, u));
        W2_TRY(w2_copy_field(c, vn, v));
    }
    if (thermal || atd) W2_TRY(w2_copy_field(c, tn, t));   // with the ATD model t carries tss even in cold flow
    if (atd) {                                             // :706-727
        W2Atd *a = c->atd;
        double *uss = c->fld[W2_F_USS], *vss = c->fld[W2_F_VSS], *tss = c->fld[W2_F_TSS];
        W2_TRY(w2_copy_field(c, a->usn, uss));
        W2_TRY(w2_copy_field(c, a->vsn, vss));
        W2_TRY(w2_copy_field(c, a->tsn, tss));
        W2_TRY(w2_axpy3(c, +1.0, un, uss, vn, vss, tn, tss));
    }
    // d changes through EqState, which runs whenever neqstate == 1 -- with or without the energy equation (:853)
    const bool eqstate = c->th.neqstate == 1;
    if (thermal || eqstate || !c->dn_valid) { W2_TRY(w2_copy_field(c, c->fld[W2_F_DN], c->fld[W2_F_D])); c->dn_valid = 1; }
    int nme = c->par.nmeiter;
    if (!thermal && nme > 0) nme = 1;                        // :736
    int nQL = -1, nSor = 0, conv = 0;
    for (int l = 1; l <= nme; ++l) {                         // momentum-energy iterations, :738-880
        // :741-747  starred quantities
        W2_TRY(w2_copy_field(c, us, swap_uv ? un : u));   // (swap_uv: u's numbers are in un's buffer now)
        W2_TRY(w2_copy_field(c, vs, swap_uv ? vn : v));
        if (thermal || eqstate) W2_TRY(w2_copy_field(c, ts, t));   // EqState reads ts (:853)
        // us == un on the first pass (not with the ATD model, where un carries uss: then the loop of :114-119 is real), so the initialisation loop of nAuxMomentum (:114-119) is a no-op there;
        // on later passes it resets us, vs to un, vn as the reference does
        W2_TRY(w2_nauxmomentum(c, /*init_star=*/l > 1 || atd, nullptr));   // :753; nQLiter comes back with the step's norms
        if (l == 1) {
            // The momentum solve does not read p; step_host uploads p on a second stream meanwhile.  pn <- p
            // (:696) therefore happens here, after that upload has landed.
            if (c->p_pending) { W2_CUDA(cudaStreamWaitEvent(s, c->ev_p, 0)); c->p_pending = 0; }
            W2_TRY(w2_copy_field(c, pn, p));
            cudaEventRecord(c->ev[1], s);
        }
        if (c->par.nfiltu == 1) W2_TRY(w2_filter(c, W2_U, c->par.fpu, us));   // :783
        if (c->par.nfiltv == 1) W2_TRY(w2_filter(c, W2_V, c->par.fpv, vs));   // :788
        W2_TRY(w2_vel_bc(c, us, vs));                   // :793
        W2_TRY(w2_pres_bc(c, p));                       // :797
        if (l == 1) cudaEventRecord(c->ev[2], s);
        W2_TRY(w2_ppe(c, us, vs, p, nullptr, nullptr)); // :803; the outcome comes back with the step's norms (w2_sor_collect)
        if (l == 1) cudaEventRecord(c->ev[3], s);
        W2_TRY(w2_pres_bc(c, p));                       // :813
        W2_TRY(w2_project(c, p, us, vs));               // :820
        W2_TRY(exchange_uv(c, us, vs));
        W2_TRY(w2_vel_bc(c, us, vs));                   // :829
        W2_TRY(w2_pres_bc(c, p));                       // :833
        if (thermal) W2_TRY(w2_thermenergy(c, ts));     // :840-850
        if (eqstate) W2_TRY(w2_eqstate(c, p, ts, c->fld[W2_F_D]));   // :853-855 (not gated on thermal_energy)
        double dme[3] = {0, 0, 0};
        if (nme > 1) {                                  // :857-859 (only the l > 1 test reads them)
            W2_TRY(w2_norm_reset(c));
            W2_TRY(w2_diffmaxnorm_async(c, u, us, 0));
            W2_TRY(w2_diffmaxnorm_async(c, v, vs, 1));
            W2_TRY(w2_diffmaxnorm_async(c, t, ts, 2));
            W2_TRY(w2_norm_fetch(c, 3, dme));
        }
        if (swap_uv) {                                  // :864-870
            swap_fld(W2_F_U, W2_F_US); swap_fld(W2_F_V, W2_F_VS);
            u = c->fld[W2_F_U]; v = c->fld[W2_F_V]; us = c->fld[W2_F_US]; vs = c->fld[W2_F_VS];
        } else {
            W2_TRY(w2_copy_field(c, u, us));
            W2_TRY(w2_copy_field(c, v, vs));
        }
        if (thermal) W2_TRY(w2_copy_field(c, t, ts));
        double dmemax = dme[0] > dme[1] ? dme[0] : dme[1];
        dmemax = dmemax > dme[2] ? dmemax : dme[2];
        if (l > 1 && dmemax < c->th.dmeittol) break;    // :873-877
    }
    if (thermal && c->th.nfiltt == 1) W2_TRY(w2_filter(c, W2_T, c->th.fpt, t));   // :890-894
    if (atd) {                                      // :896-940
        W2Atd *a = c->atd;
        W2_TRY(w2_axpy3(c, -1.0, u, a->usn, v, a->vsn, t, a->tsn));
        W2_TRY(w2_smallscale(c, 1, u, v, t));
        W2_TRY(w2_axpy3(c, +1.0, u, c->fld[W2_F_USS], v, c->fld[W2_F_VSS], t, c->fld[W2_F_TSS]));
    }
    W2_TRY(w2_vel_bc(c, u, v));                     // :946
    W2_TRY(w2_pres_bc(c, p));                       // :950
    if (thermal) W2_TRY(w2_temp_bc(c, t));          // :955
    W2_TRY(w2_norm_reset(c));
    W2_TRY(w2_diffmaxnorm_async(c, pn, p, 0));      // :962-965
    W2_TRY(w2_diffmaxnorm_async(c, un, u, 1));
    W2_TRY(w2_diffmaxnorm_async(c, vn, v, 2));
    if (thermal || atd) W2_TRY(w2_diffmaxnorm_async(c, tn, t, 3));
    // :1000-1024; enqueued before the closing event so that the step time includes the particle integration
    // (the norms above do not depend on it; after a divergence abort the particle state is as undefined as the fields)
    if (c->traj && c->traj->active) W2_TRY(w2_traject_step(c));
    W2_TRY(w2_probes_step(c));                      // :984-995 time-series monitor points
    W2_TRY(w2_timeavg_step(c));                     // :1107-1208 time-average accumulation
    cudaEventRecord(c->ev[6], s);
    double dif[4] = {0, 0, 0, 0};
    W2_TRY(w2_norm_fetch(c, thermal || atd ? 4 : 3, dif)); // syncs the stream
    if (nme > 0) { nQL = w2_ql_result(c); W2_TRY(w2_sor_collect(c, &nSor, &conv)); }
    float t_tot = 0, t_mom = 0, t_bc1 = 0, t_ppe = 0, t_tail = 0;
    if (nme > 0) {
        cudaEventElapsedTime(&t_tot, c->ev[0], c->ev[6]);
        cudaEventElapsedTime(&t_mom, c->ev[0], c->ev[1]);
        cudaEventElapsedTime(&t_bc1, c->ev[1], c->ev[2]);
        cudaEventElapsedTime(&t_ppe, c->ev[2], c->ev[3]);
        cudaEventElapsedTime(&t_tail, c->ev[3], c->ev[6]);
    }
    c->last_ms[0] += t_tot; c->last_ms[1] += t_mom; c->last_ms[2] += t_ppe; c->last_ms[3] += t_bc1 + t_tail;
    double difmax = dif[0];
    for (int q = 1; q < 4; ++q) difmax = difmax > dif[q] ? difmax : dif[q];
    const int diverged = difmax > 1.0e12;          // :969
    if (log) {
        log->nQLiter = nQL; log->nSorConv = nSor; log->sor_converged = conv; log->diverged = diverged;
        for (int q = 0; q < 4; ++q) log->dif[q] = dif[q];
    }
    if (diverged) {
        w2_set_error("* Solution diverged. Please reduce CFL number.");   // :970
        return W2_ERR_DIVERGED;
    }
    return W2_OK;
}

extern "C" int wolfd2_b200_step(wolfd2_ctx *c, int32_t nsteps, wolfd2_step_log *logs) {
    if (!c || nsteps < 0) return W2_ERR_BAD_ARG;
    W2_CUDA(cudaSetDevice(c->device));
    for (int q = 0; q < 4; ++q) c->last_ms[q] = 0.0;
    c->sor_ms = 0.0; c->sor_iters = 0; c->host_syncs = 0;
    for (int k = 0; k < nsteps; ++k) W2_TRY(one_step(c, logs ? &logs[k] : nullptr));
    return W2_OK;
}

extern "C" int wolfd2_b200_step_host(wolfd2_ctx *c, int32_t nsteps, double *u, double *v, double *p,
                                     wolfd2_step_log *logs) {
    if (!c || !u || !v || !p) return W2_ERR_BAD_ARG;
    W2_CUDA(cudaSetDevice(c->device));
    W2_TRY(w2_upload2d(c, c->fld[W2_F_U], u));
    W2_TRY(w2_upload2d(c, c->fld[W2_F_V], v));
    if (nsteps > 0) {   // p is first needed after the momentum solve: its upload overlaps it (one_step waits on ev_p)
        W2_TRY(w2_upload2d(c, c->fld[W2_F_P], p, c->copy_stream));
        W2_CUDA(cudaEventRecord(c->ev_p, c->copy_stream));
        c->p_pending = 1;
    } else {
        W2_TRY(w2_upload2d(c, c->fld[W2_F_P], p));
    }
    int rc = wolfd2_b200_step(c, nsteps, logs);
    if (c->p_pending) { cudaStreamSynchronize(c->copy_stream); c->p_pending = 0; }   // a failed step never waited
    if (rc != W2_OK && rc != W2_ERR_DIVERGED) return rc;
    W2_TRY(w2_download2d(c, u, c->fld[W2_F_U]));
    W2_TRY(w2_download2d(c, v, c->fld[W2_F_V]));
    W2_TRY(w2_download2d(c, p, c->fld[W2_F_P]));
    W2_CUDA(cudaStreamSynchronize(c->stream));
    return rc;
}

extern "C" int wolfd2_b200_last_timing(wolfd2_ctx *c, double ms[4], int64_t launches[4]) {
    if (!c) return W2_ERR_BAD_ARG;
    for (int q = 0; q < 4; ++q) { if (ms) ms[q] = c->last_ms[q]; if (launches) launches[q] = c->launches[q]; }
    return W2_OK;
}
extern "C" int wolfd2_b200_last_host_syncs(wolfd2_ctx *c, int64_t *syncs) {
    if (!c || !syncs) return W2_ERR_BAD_ARG;
    *syncs = c->host_syncs;
    return W2_OK;
}
extern "C" int wolfd2_b200_last_sor_timing(wolfd2_ctx *c, double *ms_total, int64_t *iterations) {
    if (!c) return W2_ERR_BAD_ARG;
    if (ms_total) *ms_total = c->sor_ms;
    if (iterations) *iterations = c->sor_iters;
    return W2_OK;
}

// ------------------------------------------------------------------------------ literal shims
// The reference has no error channel on these routines (it prints and `stop`s, e.g.
// src/momentum.f:472-474); the shims do the same: message on stderr, then abort().

static wolfd2_ctx *g_shim = nullptr;

static void die(const char *who) {
    fprintf(stderr, "wolfd2_b200: %s failed: %s\n", who, wolfd2_b200_last_error());
    abort();
}
#define SHIM_TRY(call, who) do { if ((call) != W2_OK) die(who); } while (0)

static wolfd2_ctx *shim_ctx(int nx, int ny, const char *who) {
    if (g_shim && (g_shim->nx != nx || g_shim->ny != ny || g_shim->mnx != g_mnx || g_shim->mny != g_mny)) {
        wolfd2_b200_destroy(g_shim);
        g_shim = nullptr;
    }
    if (!g_shim) {
        SHIM_TRY(w2_ctx_create_raw(&g_shim, nx, ny), who);
        g_shim->cart_state = -2;   // the literal shims get new metric arrays with every call: always the general variant
        memset(&g_shim->par, 0, sizeof(g_shim->par));
        g_shim->par.nx = nx; g_shim->par.ny = ny;
    }
    if (cudaSetDevice(g_shim->device) != cudaSuccess) die(who);
    return g_shim;
}

static void shim_regions(wolfd2_ctx *c, const int32_t *nReg, const int32_t *nRegBrd, const int32_t *nRegType,
                         const int32_t *nMomBdTp, const double *dBCVal, const double *po, const double *c1,
                         const double *c2, const char *who) {
    W2Regions r;
    SHIM_TRY(w2_fill_regions(&r, c->nx, c->ny, nReg, nRegBrd, nRegType, nMomBdTp, dBCVal, po, c1, c2), who);
    SHIM_TRY(w2_ctx_set_regions(c, &r), who);
}
static void up(wolfd2_ctx *c, double *dev, const double *host, const char *who) { SHIM_TRY(w2_upload2d(c, dev, host), who); }
static void down(wolfd2_ctx *c, double *host, const double *dev, const char *who) { SHIM_TRY(w2_download2d(c, host, dev), who); }
static void sync(wolfd2_ctx *c, const char *who) { if (cudaStreamSynchronize(c->stream) != cudaSuccess) { w2_set_error("stream sync: %s", cudaGetErrorString(cudaGetLastError())); die(who); } }

extern "C" void velboundcond_(const int32_t *nx, const int32_t *ny, const int32_t *nReg, const int32_t *nRegBrd,
                              const int32_t *nMomBdTp, const double *dBCVal, double *u, double *v) {
    const char *who = "velboundcond_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    shim_regions(c, nReg, nRegBrd, nullptr, nMomBdTp, dBCVal, nullptr, nullptr, nullptr, who);
    up(c, c->fld[W2_F_U], u, who); up(c, c->fld[W2_F_V], v, who);
    SHIM_TRY(w2_vel_bc(c, c->fld[W2_F_U], c->fld[W2_F_V]), who);
    down(c, u, c->fld[W2_F_U], who); down(c, v, c->fld[W2_F_V], who);
    sync(c, who);
}

extern "C" void veloutflowbcs_(const int32_t *nx, const int32_t *ny, const int32_t *nReg, const int32_t *nRegBrd,
                               const int32_t *nMomBdTp, const double *dBCVal, double *u, double *v) {
    const char *who = "veloutflowbcs_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    shim_regions(c, nReg, nRegBrd, nullptr, nMomBdTp, dBCVal, nullptr, nullptr, nullptr, who);
    up(c, c->fld[W2_F_U], u, who); up(c, c->fld[W2_F_V], v, who);
    SHIM_TRY(w2_outflow_bc(c, c->fld[W2_F_U], c->fld[W2_F_V]), who);
    down(c, u, c->fld[W2_F_U], who); down(c, v, c->fld[W2_F_V], who);
    sync(c, who);
}

extern "C" void presboundcond_(const int32_t *nx, const int32_t *ny, const int32_t *nReg, const int32_t *nRegBrd,
                               const int32_t *nRegType, const int32_t *nMomBdTp, const double *dBCVal, double *p) {
    const char *who = "presboundcond_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    shim_regions(c, nReg, nRegBrd, nRegType, nMomBdTp, dBCVal, nullptr, nullptr, nullptr, who);
    up(c, c->fld[W2_F_P], p, who);
    SHIM_TRY(w2_pres_bc(c, c->fld[W2_F_P]), who);
    down(c, p, c->fld[W2_F_P], who);
    sync(c, who);
}

extern "C" void divergence_(const int32_t *nx, const int32_t *ny, const int32_t *nloc, const double *xet,
                            const double *yet, const double *xzi, const double *yzi, const double *u, const double *v,
                            double *div) {
    const char *who = "divergence_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    up(c, c->met.xeu, xet, who); up(c, c->met.yeu, yet, who); up(c, c->met.xzv, xzi, who); up(c, c->met.yzv, yzi, who);
    up(c, c->fld[W2_F_U], u, who); up(c, c->fld[W2_F_V], v, who);
    up(c, c->div, div, who);  // cells outside 1..nx,1..ny keep the caller's values
    SHIM_TRY(w2_divergence(c, c->fld[W2_F_U], c->fld[W2_F_V], c->div, *nloc, c->met.xeu, c->met.yeu, c->met.xzv, c->met.yzv), who);
    down(c, div, c->div, who);
    sync(c, who);
}

extern "C" void ppe_(const int32_t *nx, const int32_t *ny, const int32_t *nReg, const int32_t *nRegBrd,
                     const int32_t *nRegType, const int32_t *lCartesGrid, const int32_t *nPpeSolver,
                     const int32_t *msorit, int32_t *nSorConv, const double *dk, const double *sortol,
                     const double *sorrel, const double *rau, const double *rbu, const double *rbv, const double *rgv,
                     const double *xeu, const double *yeu, const double *xzv, const double *yzv, const double *u,
                     const double *v, double *p) {
    const char *who = "ppe_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    shim_regions(c, nReg, nRegBrd, nRegType, nullptr, nullptr, nullptr, nullptr, nullptr, who);
    if (*nPpeSolver < 1 || *nPpeSolver > 6) {   // pressure.f:238-246: message, then "not converged"
        fprintf(stderr, " Wrong nPpeSolver flag passed to Ppe\n");
        *nSorConv = *msorit;
        return;
    }
    c->par.lCartesGrid = *lCartesGrid; c->par.nPpeSolver = *nPpeSolver; c->par.msorit = *msorit;
    c->par.dk = *dk; c->par.sortol = *sortol; c->par.sorrel = *sorrel;
    up(c, c->met.rau, rau, who); up(c, c->met.rbu, rbu, who); up(c, c->met.rbv, rbv, who); up(c, c->met.rgv, rgv, who);
    up(c, c->met.xeu, xeu, who); up(c, c->met.yeu, yeu, who); up(c, c->met.xzv, xzv, who); up(c, c->met.yzv, yzv, who);
    up(c, c->fld[W2_F_U], u, who); up(c, c->fld[W2_F_V], v, who); up(c, c->fld[W2_F_P], p, who);
    int n = 0, conv = 0;
    SHIM_TRY(w2_ppe(c, c->fld[W2_F_U], c->fld[W2_F_V], c->fld[W2_F_P], &n, &conv), who);
    if (!conv) printf(" Warning: SOR iterations did not converge after %d iterations.\n", *msorit);  // :243
    *nSorConv = n;
    down(c, p, c->fld[W2_F_P], who);
    sync(c, who);
}

extern "C" void project_(const int32_t *nx, const int32_t *ny, const int32_t *nReg, const int32_t *nRegBrd,
                         const int32_t *nRegType, const int32_t *nMomBdTp, const double *dk, const double *dju,
                         const double *djv, const double *yeu, const double *xzv, const double *yzu, const double *xev,
                         const double *p, double *u, double *v) {
    const char *who = "project_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    shim_regions(c, nReg, nRegBrd, nRegType, nMomBdTp, nullptr, nullptr, nullptr, nullptr, who);
    c->par.dk = *dk;
    up(c, c->met.dju, dju, who); up(c, c->met.djv, djv, who); up(c, c->met.yeu, yeu, who);
    up(c, c->met.xzv, xzv, who); up(c, c->met.yzu, yzu, who); up(c, c->met.xev, xev, who);
    up(c, c->fld[W2_F_P], p, who); up(c, c->fld[W2_F_U], u, who); up(c, c->fld[W2_F_V], v, who);
    SHIM_TRY(w2_project(c, c->fld[W2_F_P], c->fld[W2_F_U], c->fld[W2_F_V]), who);
    down(c, u, c->fld[W2_F_U], who); down(c, v, c->fld[W2_F_V], who);
    sync(c, who);
}

extern "C" void filter_(const int32_t *nx, const int32_t *ny, const int32_t *ncomp, const int32_t *nReg,
                        const int32_t *nRegBrd, const int32_t *nRegType, const int32_t *nMomBdTp,
                        const int32_t *nTRgType, const double *fp, double *qu) {
    const char *who = "filter_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    shim_regions(c, nReg, nRegBrd, nRegType, nMomBdTp, nullptr, nullptr, nullptr, nullptr, who);
    if (*ncomp == W2_T) SHIM_TRY(w2_set_thermal_tables(c, nTRgType, nullptr, nullptr, nullptr), who);
    up(c, c->fld[W2_F_US], qu, who);
    SHIM_TRY(w2_filter(c, *ncomp, *fp, c->fld[W2_F_US]), who);
    down(c, qu, c->fld[W2_F_US], who);
    sync(c, who);
}

extern "C" void tempboundcond_(const int32_t *nx, const int32_t *ny, const int32_t *nReg, const int32_t *nRegBrd,
                               const int32_t *nTRgType, const int32_t *nTemBdTp, const double *dTRgVal,
                               const double *dBCVal, double *t) {
    const char *who = "tempboundcond_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    shim_regions(c, nReg, nRegBrd, nullptr, nullptr, dBCVal, nullptr, nullptr, nullptr, who);
    SHIM_TRY(w2_set_thermal_tables(c, nTRgType, nTemBdTp, dTRgVal, nullptr), who);
    up(c, c->fld[W2_F_T], t, who);
    SHIM_TRY(w2_temp_bc(c, c->fld[W2_F_T]), who);
    down(c, t, c->fld[W2_F_T], who);
    sync(c, who);
}

extern "C" void thermenergy_(const int32_t *nx, const int32_t *ny, const int32_t *nReg, const int32_t *nRegBrd,
                             const int32_t *nTRgType, const int32_t *nTemBdTp, const double *dk, const double *pe,
                             const double *dTRgVal, const double *dHGSTval, const double *dBCVal,
                             const double *rau, const double *rbu, const double *rbv, const double *rgv, const double *djc,
                             const double *xeu, const double *yeu, const double *xzv, const double *yzv,
                             const double *xec, const double *yec, const double *xzc, const double *yzc,
                             const double *un, const double *vn, const double *u, const double *v,
                             const double *tn, double *t) {
    const char *who = "thermenergy_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    shim_regions(c, nReg, nRegBrd, nullptr, nullptr, dBCVal, nullptr, nullptr, nullptr, who);
    SHIM_TRY(w2_set_thermal_tables(c, nTRgType, nTemBdTp, dTRgVal, dHGSTval), who);
    c->par.dk = *dk; c->th.pe = *pe;
    W2Metrics &m = c->met;
    up(c, m.rau, rau, who); up(c, m.rbu, rbu, who); up(c, m.rbv, rbv, who); up(c, m.rgv, rgv, who); up(c, m.djc, djc, who);
    up(c, m.xeu, xeu, who); up(c, m.yeu, yeu, who); up(c, m.xzv, xzv, who); up(c, m.yzv, yzv, who);
    up(c, m.xec, xec, who); up(c, m.yec, yec, who); up(c, m.xzc, xzc, who); up(c, m.yzc, yzc, who);
    up(c, c->fld[W2_F_UN], un, who); up(c, c->fld[W2_F_VN], vn, who);
    up(c, c->fld[W2_F_US], u, who); up(c, c->fld[W2_F_VS], v, who);
    up(c, c->fld[W2_F_TN], tn, who); up(c, c->fld[W2_F_TS], t, who);
    cudaMemsetAsync(c->dus + c->row_off, 0, c->nelem * sizeof(double), c->stream);
    SHIM_TRY(w2_thermenergy(c, c->fld[W2_F_TS]), who);
    down(c, t, c->fld[W2_F_TS], who);
    sync(c, who);
}

extern "C" void eqstate_(const int32_t *nx, const int32_t *ny, const double *uref, const double *densref,
                         const double *tmax, const double *tref, const double *rconst, const double *p,
                         const double *t, double *den) {
    const char *who = "eqstate_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    c->th.uref = *uref; c->th.densref = *densref; c->th.tmax = *tmax; c->th.tref = *tref; c->th.rconst = *rconst;
    up(c, c->fld[W2_F_P], p, who); up(c, c->fld[W2_F_T], t, who); up(c, c->fld[W2_F_D], den, who);
    SHIM_TRY(w2_eqstate(c, c->fld[W2_F_P], c->fld[W2_F_T], c->fld[W2_F_D]), who);
    down(c, den, c->fld[W2_F_D], who);
    sync(c, who);
}

extern "C" double diffmaxnorm_(const int32_t *nx, const int32_t *ny, const double *un, const double *u) {
    const char *who = "diffmaxnorm_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    up(c, c->fld[W2_F_UN], un, who); up(c, c->fld[W2_F_U], u, who);
    SHIM_TRY(w2_norm_reset(c), who);
    SHIM_TRY(w2_diffmaxnorm_async(c, c->fld[W2_F_UN], c->fld[W2_F_U], 0), who);
    double r = 0.0;
    SHIM_TRY(w2_norm_fetch(c, 1, &r), who);
    return r;
}
extern "C" double dmaxnorm_(const int32_t *nx, const int32_t *ny, const double *u) {
    const char *who = "dmaxnorm_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    up(c, c->fld[W2_F_U], u, who);
    SHIM_TRY(w2_norm_reset(c), who);
    SHIM_TRY(w2_dmaxnorm_async(c, c->fld[W2_F_U], 0), who);
    double r = 0.0;
    SHIM_TRY(w2_norm_fetch(c, 1, &r), who);
    return r;
}

// a(3,n) AoS in, solution in b (momentum.f:1307-1339)
__global__ void aos_to_soa_kernel(long long n, const double *__restrict__ a3, double *__restrict__ a, double *__restrict__ d,
                                  double *__restrict__ c) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { a[i] = a3[3 * i]; d[i] = a3[3 * i + 1]; c[i] = a3[3 * i + 2]; }
}
extern "C" void alttridlu_(const int32_t *n_, double *a, double *b) {
    const char *who = "alttridlu_";
    const long long n = *n_;
    // a context sized for the chain: nx*ny >= n
    int side = 8;
    while ((long long)side * side < n + 8) side += 8;
    if (g_mnx < side + 1 || g_mny < side + 1) { w2_set_error("alttridlu_: n=%lld needs mnx,mny >= %d (wolfd2_b200_config)", n, side + 1); die(who); }
    wolfd2_ctx *c = shim_ctx(side, side, who);
    SHIM_TRY(w2_ensure_chain(c), who);
    double *tmp = nullptr;  // AoS staging
    if (cudaMalloc((void **)&tmp, 3 * n * sizeof(double)) != cudaSuccess) { w2_set_error("alttridlu_: out of device memory"); die(who); }
    cudaMemcpyAsync(tmp, a, 3 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    aos_to_soa_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, tmp, c->ta, c->td, c->tc);
    cudaStreamSynchronize(c->stream);
    cudaFree(tmp);
    cudaMemcpyAsync(c->tb, b, n * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    SHIM_TRY(w2_tri_solve(c, n, c->ta, c->td, c->tc, c->tb, c->tx, 1), who);
    cudaMemcpyAsync(b, c->tx, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    sync(c, who);
}

static void up_xmom(wolfd2_ctx *c, const double *rbn, const double *rgn, const double *rac, const double *rbc,
                    const double *dju, const double *xec, const double *yec, const double *xzn, const double *yzn,
                    const double *xeu, const double *yeu, const double *xzu, const double *yzu, const char *who) {
    W2Metrics &t = c->met;
    up(c, t.rbn, rbn, who); up(c, t.rgn, rgn, who); up(c, t.rac, rac, who); up(c, t.rbc, rbc, who); up(c, t.dju, dju, who);
    up(c, t.xec, xec, who); up(c, t.yec, yec, who); up(c, t.xzn, xzn, who); up(c, t.yzn, yzn, who);
    up(c, t.xeu, xeu, who); up(c, t.yeu, yeu, who); up(c, t.xzu, xzu, who); up(c, t.yzu, yzu, who);
}
static void up_ymom(wolfd2_ctx *c, const double *ran, const double *rbn, const double *rbc, const double *rgc,
                    const double *djv, const double *xen, const double *yen, const double *xzc, const double *yzc,
                    const double *xev, const double *yev, const double *xzv, const double *yzv, const char *who) {
    W2Metrics &t = c->met;
    up(c, t.ran, ran, who); up(c, t.rbn, rbn, who); up(c, t.rbc, rbc, who); up(c, t.rgc, rgc, who); up(c, t.djv, djv, who);
    up(c, t.xen, xen, who); up(c, t.yen, yen, who); up(c, t.xzc, xzc, who); up(c, t.yzc, yzc, who);
    up(c, t.xev, xev, who); up(c, t.yev, yev, who); up(c, t.xzv, xzv, who); up(c, t.yzv, yzv, who);
}

extern "C" void xmomentum_(const int32_t *nx, const int32_t *ny, const int32_t *nReg, const int32_t *nRegBrd,
                           const int32_t *nRegType, const int32_t *nMomBdTp, const double *dk, const double *re,
                           const double *dPRporos, const double *dPRporc1, const double *dPRporc2, const double *rbn,
                           const double *rgn, const double *rac, const double *rbc, const double *dju, const double *xec,
                           const double *yec, const double *xzn, const double *yzn, const double *xeu, const double *yeu,
                           const double *xzu, const double *yzu, const double *us, const double *vs, const double *un,
                           const double *vn, double *dus) {
    const char *who = "xmomentum_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    shim_regions(c, nReg, nRegBrd, nRegType, nMomBdTp, nullptr, dPRporos, dPRporc1, dPRporc2, who);
    c->par.dk = *dk; c->par.re = *re;
    up_xmom(c, rbn, rgn, rac, rbc, dju, xec, yec, xzn, yzn, xeu, yeu, xzu, yzu, who);
    up(c, c->fld[W2_F_US], us, who); up(c, c->fld[W2_F_VS], vs, who);
    up(c, c->fld[W2_F_UN], un, who); up(c, c->fld[W2_F_VN], vn, who);
    up(c, c->dus, dus, who);
    SHIM_TRY(w2_xmomentum(c, c->dus), who);
    down(c, dus, c->dus, who);
    sync(c, who);
}

extern "C" void ymomentum_(const int32_t *nx, const int32_t *ny, const int32_t *nReg, const int32_t *nRegBrd,
                           const int32_t *nRegType, const int32_t *nMomBdTp, const double *dk, const double *re,
                           const double *fr, const double *dPRporos, const double *dPRporc1, const double *dPRporc2,
                           const double *ran, const double *rbn, const double *rbc, const double *rgc, const double *djv,
                           const double *xen, const double *yen, const double *xzc, const double *yzc, const double *xev,
                           const double *yev, const double *xzv, const double *yzv, const double *d, const double *dn,
                           const double *us, const double *vs, const double *un, const double *vn, double *dvs) {
    const char *who = "ymomentum_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    shim_regions(c, nReg, nRegBrd, nRegType, nMomBdTp, nullptr, dPRporos, dPRporc1, dPRporc2, who);
    c->par.dk = *dk; c->par.re = *re; c->par.fr = *fr;
    up_ymom(c, ran, rbn, rbc, rgc, djv, xen, yen, xzc, yzc, xev, yev, xzv, yzv, who);
    up(c, c->fld[W2_F_D], d, who); up(c, c->fld[W2_F_DN], dn, who);
    c->d_nonzero = 1;
    up(c, c->fld[W2_F_US], us, who); up(c, c->fld[W2_F_VS], vs, who);
    up(c, c->fld[W2_F_UN], un, who); up(c, c->fld[W2_F_VN], vn, who);
    up(c, c->dvs, dvs, who);
    SHIM_TRY(w2_ymomentum(c, c->dvs), who);
    down(c, dvs, c->dvs, who);
    sync(c, who);
}

extern "C" int32_t nauxmomentum_(const int32_t *nx, const int32_t *ny, const int32_t *mqiter, const int32_t *nReg,
                                 const int32_t *nRegBrd, const int32_t *nRegType, const int32_t *nMomBdTp,
                                 const double *dk, const double *re, const double *fr, const double *qtol,
                                 const double *dPRporos, const double *dPRporc1, const double *dPRporc2,
                                 const double *dBCVal, const double *ran, const double *rbn, const double *rgn,
                                 const double *rac, const double *rbc, const double *rgc, const double *dju,
                                 const double *djv, const double *xec, const double *yec, const double *xzn,
                                 const double *yzn, const double *xen, const double *yen, const double *xzc,
                                 const double *yzc, const double *xeu, const double *yeu, const double *xzu,
                                 const double *yzu, const double *xev, const double *yev, const double *xzv,
                                 const double *yzv, const double *d, const double *dn, const double *un, const double *vn,
                                 double *us, double *vs) {
    const char *who = "nauxmomentum_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    shim_regions(c, nReg, nRegBrd, nRegType, nMomBdTp, dBCVal, dPRporos, dPRporc1, dPRporc2, who);
    c->par.mqiter = *mqiter; c->par.dk = *dk; c->par.re = *re; c->par.fr = *fr; c->par.qtol = *qtol;
    up_xmom(c, rbn, rgn, rac, rbc, dju, xec, yec, xzn, yzn, xeu, yeu, xzu, yzu, who);
    up_ymom(c, ran, rbn, rbc, rgc, djv, xen, yen, xzc, yzc, xev, yev, xzv, yzv, who);
    up(c, c->fld[W2_F_D], d, who); up(c, c->fld[W2_F_DN], dn, who);
    c->d_nonzero = 1;
    up(c, c->fld[W2_F_UN], un, who); up(c, c->fld[W2_F_VN], vn, who);
    up(c, c->fld[W2_F_US], us, who); up(c, c->fld[W2_F_VS], vs, who);
    // dus/dvs are local to nAuxMomentum and zeroed there (:139-144)
    cudaMemsetAsync(c->dus, 0, c->nelem * sizeof(double), c->stream);
    cudaMemsetAsync(c->dvs, 0, c->nelem * sizeof(double), c->stream);
    int nql = -1;
    SHIM_TRY(w2_nauxmomentum(c, /*init_star=*/1, &nql), who);
    down(c, us, c->fld[W2_F_US], who); down(c, vs, c->fld[W2_F_VS], who);
    sync(c, who);
    return nql;
}

// Node averages of the resident fields for output dumps (src/main.f:1053-1062, :1300-1330: VelAvg and PTDAvg before
// SaveStdVarsP3D / SaveTimeSrs): computed on the device into scratch arrays, only the three averaged arrays cross
// the bus.  set 0: (u, v, p) -> (util, vbar, pav); set 1: the small-scale fields (uss, vss, pss).  Cells outside the
// node range 1..nx, 1..ny come back as zero.  tav: TAveraged of t (set 0) or tss (set 1).
extern "C" int wolfd2_b200_node_averages(wolfd2_ctx *c, int32_t set, double *util, double *vbar, double *pav, double *tav) {
    if (!c || set < 0 || set > 1) return W2_ERR_BAD_ARG;
    if (c->world > 1) { w2_set_error("node averages are not supported on a row slab"); return W2_ERR_UNSUPPORTED; }
    const double *u = c->fld[set ? W2_F_USS : W2_F_U], *v = c->fld[set ? W2_F_VSS : W2_F_V], *p = c->fld[set ? W2_F_PSS : W2_F_P];
    if (!u || !v || !p) { w2_set_error("the small-scale fields do not exist in this context"); return W2_ERR_BAD_ARG; }
    W2_CUDA(cudaSetDevice(c->device));
    double *a = c->x1, *b = c->div, *q = c->fld[W2_F_B];   // scratch: rebuilt by their producers before every use
    const size_t bytes = c->nelem * sizeof(double);
    if (util || vbar) {
        W2_CUDA(cudaMemsetAsync(a, 0, bytes, c->stream));
        W2_CUDA(cudaMemsetAsync(b, 0, bytes, c->stream));
        W2_TRY(w2_velavg(c, u, v, a, b));
        if (util) W2_TRY(w2_download2d(c, util, a));
        if (vbar) W2_TRY(w2_download2d(c, vbar, b));
    }
    if (pav) {
        W2_CUDA(cudaMemsetAsync(q, 0, bytes, c->stream));
        W2_TRY(w2_ptdavg(c, p, q));
        W2_TRY(w2_download2d(c, pav, q));
    }
    if (tav) {   // TAveraged with nScale = set (src/main.f:1057-1066); needs the thermal region tables
        const double *t = c->fld[set ? W2_F_TSS : W2_F_T];
        if (!t || !c->th_tables) { w2_set_error("node average of t: no thermal tables (wolfd2_b200_set_thermal) or no such field"); return W2_ERR_BAD_ARG; }
        W2_CUDA(cudaMemsetAsync(a, 0, bytes, c->stream));
        W2_TRY(w2_taveraged(c, set, t, a));
        W2_TRY(w2_download2d(c, tav, a));
    }
    W2_CUDA(cudaStreamSynchronize(c->stream));
    return W2_OK;
}

// ---- unit-parity shims of the routines below XMomentum / YMomentum / Ppe in the reference's call tree
// ("internal but worth exporting").  Output arrays are in/out as in the reference: cells
// outside the loop ranges keep the caller's values.
extern "C" void convcoef_(const int32_t *nx, const int32_t *ny, const int32_t *ncomp, const int32_t *njacob,
                          const double *xzi, const double *xet, const double *yzi, const double *yet, const double *u,
                          const double *v, double *cc1, double *cc2) {
    const char *who = "convcoef_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    W2Metrics &t = c->met;
    up(c, t.xzn, xzi, who); up(c, t.xec, xet, who); up(c, t.yzn, yzi, who); up(c, t.yec, yet, who);
    up(c, c->fld[W2_F_US], u, who); up(c, c->fld[W2_F_VS], v, who);
    up(c, c->dus, cc1, who); up(c, c->dvs, cc2, who);
    SHIM_TRY(w2_unit_convcoef(c, *ncomp, *njacob, t.xzn, t.xec, t.yzn, t.yec, c->fld[W2_F_US], c->fld[W2_F_VS], c->dus, c->dvs), who);
    down(c, cc1, c->dus, who); down(c, cc2, c->dvs, who);
    sync(c, who);
}
static void shim_dconv(int comp, const int32_t *nx, const int32_t *ny, const double *c1, const double *c2, const double *q,
                       double *out, const char *who) {
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    up(c, c->dus, c1, who); up(c, c->dvs, c2, who); up(c, c->fld[W2_F_US], q, who); up(c, c->x1, out, who);
    SHIM_TRY(w2_unit_dconv(c, comp, c->dus, c->dvs, c->fld[W2_F_US], c->x1), who);
    down(c, out, c->x1, who);
    sync(c, who);
}
extern "C" void dconvu_(const int32_t *nx, const int32_t *ny, const double *c1, const double *c2, const double *u, double *cv) {
    shim_dconv(0, nx, ny, c1, c2, u, cv, "dconvu_");
}
extern "C" void dconvv_(const int32_t *nx, const int32_t *ny, const double *c1, const double *c2, const double *v, double *cv) {
    shim_dconv(1, nx, ny, c1, c2, v, cv, "dconvv_");
}
static void shim_ddiff(int comp, const int32_t *nx, const int32_t *ny, const double *a, const double *bc, const double *bn,
                       const double *g, const double *q, double *out, const char *who) {
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    W2Metrics &t = c->met;
    double *da = comp == 0 ? t.rac : t.ran, *dg = comp == 0 ? t.rgn : t.rgc;
    up(c, da, a, who); up(c, t.rbc, bc, who); up(c, t.rbn, bn, who); up(c, dg, g, who);
    up(c, c->fld[W2_F_US], q, who); up(c, c->x1, out, who);
    SHIM_TRY(w2_unit_ddiff(c, comp, da, t.rbc, t.rbn, dg, c->fld[W2_F_US], c->x1), who);
    down(c, out, c->x1, who);
    sync(c, who);
}
extern "C" void ddiffu_(const int32_t *nx, const int32_t *ny, const double *ac, const double *bc, const double *bn,
                        const double *gn, const double *u, double *d) {
    shim_ddiff(0, nx, ny, ac, bc, bn, gn, u, d, "ddiffu_");
}
extern "C" void ddiffv_(const int32_t *nx, const int32_t *ny, const double *an, const double *bc, const double *bn,
                        const double *gc, const double *v, double *d) {
    shim_ddiff(1, nx, ny, an, bc, bn, gc, v, d, "ddiffv_");
}
extern "C" void poroscoef_(const int32_t *nx, const int32_t *ny, const int32_t *ncomp, const int32_t *njacob,
                           const int32_t *nReg, const int32_t *nRegBrd, const int32_t *nRegType, const double *dPRporos,
                           const double *dPRporc1, const double *dPRporc2, const double *u, const double *v, double *cp) {
    const char *who = "poroscoef_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    shim_regions(c, nReg, nRegBrd, nRegType, nullptr, nullptr, dPRporos, dPRporc1, dPRporc2, who);
    up(c, c->fld[W2_F_US], u, who); up(c, c->fld[W2_F_VS], v, who); up(c, c->x1, cp, who);
    SHIM_TRY(w2_unit_poroscoef(c, *ncomp, *njacob, c->fld[W2_F_US], c->fld[W2_F_VS], c->x1), who);
    down(c, cp, c->x1, who);
    sync(c, who);
}
// b is the reference's vector b(mn): entries 1..(nx-1)(ny-1) are written (pressure.f:347-352)
extern "C" void rhsppe_(const int32_t *nx, const int32_t *ny, const int32_t *lCartesGrid, const double *dk, const double *rbu,
                        const double *rbv, const double *div, const double *p, double *b) {
    const char *who = "rhsppe_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    SHIM_TRY(w2_ensure_chain(c), who);
    up(c, c->met.rbu, rbu, who); up(c, c->met.rbv, rbv, who); up(c, c->div, div, who); up(c, c->fld[W2_F_P], p, who);
    SHIM_TRY(w2_unit_rhsppe(c, *lCartesGrid != 0, *dk, c->met.rbu, c->met.rbv, c->div, c->fld[W2_F_P], c->fld[W2_F_B], c->tb), who);
    const size_t n = (size_t)(*nx - 1) * (size_t)(*ny - 1);
    if (cudaMemcpyAsync(b, c->tb, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) {
        w2_set_error("rhsppe_: copy back failed"); die(who);
    }
    sync(c, who);
}

// ---- optional paths
static void shim_thermal(wolfd2_ctx *c, const int32_t *nTRgType, const int32_t *nTemBdTp, const double *dTRgVal, const char *who) {
    SHIM_TRY(w2_set_thermal_tables(c, nTRgType, nTemBdTp, dTRgVal, nullptr), who);
}

extern "C" void smlsclbc_(const int32_t *nx, const int32_t *ny, const int32_t *nReg, const int32_t *nRegBrd,
                          const int32_t *nRegType, const int32_t *nMomBdTp, const int32_t *nTRgType, const int32_t *nTemBdTp,
                          const double *dBCVal, double *u, double *v, double *p, double *t) {
    const char *who = "smlsclbc_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    shim_regions(c, nReg, nRegBrd, nRegType, nMomBdTp, dBCVal, nullptr, nullptr, nullptr, who);
    shim_thermal(c, nTRgType, nTemBdTp, nullptr, who);
    up(c, c->fld[W2_F_U], u, who); up(c, c->fld[W2_F_V], v, who); up(c, c->fld[W2_F_P], p, who); up(c, c->fld[W2_F_T], t, who);
    SHIM_TRY(w2_smlscl_bc(c, c->fld[W2_F_U], c->fld[W2_F_V], c->fld[W2_F_P], c->fld[W2_F_T]), who);
    down(c, u, c->fld[W2_F_U], who); down(c, v, c->fld[W2_F_V], who); down(c, p, c->fld[W2_F_P], who); down(c, t, c->fld[W2_F_T], who);
    sync(c, who);
}

extern "C" void smallscale_(const int32_t *nx, const int32_t *ny, const int32_t *initflg, const int32_t *nthermen,
    const int32_t *lCartesGrid, const int32_t *nReg, const int32_t *nRegBrd,
    const int32_t *nRegType, const int32_t *nTRgType, const int32_t *nMomBdTp, const int32_t *nTemBdTp,
    const int32_t *nPpeSolver, const int32_t *msorit,
    const double *dlref, const double *uref, const double *tref, const double *tmax,
    const double *dka, const double *re, const double *pe, const double *sortol, const double *sorrel,
    const double *fp, const double *cu0, const double *TsCoef, const double *HsCoef, const double *TemCoef,
    const double *bnumc, const double *rmax, const double *rlc, const double *dTRgVal, const double *dBCVal,
    const double *rau, const double *rbu, const double *rbv, const double *rgv,
    const double *dju, const double *djv, const double *djc,
    const double *xeu, const double *yeu, const double *xzv, const double *yzv,
    const double *xzu, const double *yzu, const double *xev, const double *yev,
    const double *xec, const double *yec, const double *xzc, const double *yzc,
    const double *u1, const double *v1, const double *t1,
    double *uss, double *vss, double *pss, double *tss) {
    const char *who = "smallscale_";
    (void)nthermen; (void)xzu; (void)yev;
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    shim_regions(c, nReg, nRegBrd, nRegType, nMomBdTp, dBCVal, nullptr, nullptr, nullptr, who);
    shim_thermal(c, nTRgType, nTemBdTp, dTRgVal, who);
    wolfd2_smallscale S;
    memset(&S, 0, sizeof(S));
    S.nsmallscl = 1; S.nssPpeSlvr = *nPpeSolver; S.mssSorIt = *msorit;
    S.dlref = *dlref; S.uref = *uref; S.tref = *tref; S.tmax = *tmax; S.pe = *pe;
    S.ssSorTol = *sortol; S.ssSorRel = *sorrel;
    for (int k = 0; k < 4; ++k) S.ssFiltPar[k] = fp[k];
    S.ssCu0 = *cu0; S.ssTsCoef = *TsCoef; S.ssHsCoef = *HsCoef; S.ssTemCoef = *TemCoef;
    S.ssBnCrit = *bnumc; S.ssRMpMax = *rmax; S.ssRMpExp = *rlc;
    c->par.lCartesGrid = *lCartesGrid; c->par.dk = *dka; c->par.re = *re;
    c->par.nPpeSolver = *nPpeSolver; c->par.msorit = *msorit; c->par.sortol = *sortol; c->par.sorrel = *sorrel;
    SHIM_TRY(wolfd2_b200_set_smallscale(c, &S), who);
    W2Metrics &m = c->met;
    up(c, m.rau, rau, who); up(c, m.rbu, rbu, who); up(c, m.rbv, rbv, who); up(c, m.rgv, rgv, who);
    up(c, m.dju, dju, who); up(c, m.djv, djv, who); up(c, m.djc, djc, who);
    up(c, m.xeu, xeu, who); up(c, m.yeu, yeu, who); up(c, m.xzv, xzv, who); up(c, m.yzv, yzv, who);
    up(c, m.yzu, yzu, who); up(c, m.xev, xev, who);
    up(c, m.xec, xec, who); up(c, m.yec, yec, who); up(c, m.xzc, xzc, who); up(c, m.yzc, yzc, who);
    up(c, c->fld[W2_F_U], u1, who); up(c, c->fld[W2_F_V], v1, who); up(c, c->fld[W2_F_T], t1, who);
    up(c, c->fld[W2_F_USS], uss, who); up(c, c->fld[W2_F_VSS], vss, who);
    up(c, c->fld[W2_F_PSS], pss, who); up(c, c->fld[W2_F_TSS], tss, who);
    SHIM_TRY(w2_smallscale(c, *initflg, c->fld[W2_F_U], c->fld[W2_F_V], c->fld[W2_F_T]), who);
    down(c, uss, c->fld[W2_F_USS], who); down(c, vss, c->fld[W2_F_VSS], who);
    down(c, pss, c->fld[W2_F_PSS], who); down(c, tss, c->fld[W2_F_TSS], who);
    sync(c, who);
}

extern "C" void ptdavg_(const int32_t *nx, const int32_t *ny, const int32_t *nReg, const int32_t *nRegBrd,
                        const int32_t *nRegType, const double *p, double *pav) {
    const char *who = "ptdavg_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    shim_regions(c, nReg, nRegBrd, nRegType, nullptr, nullptr, nullptr, nullptr, nullptr, who);
    up(c, c->fld[W2_F_P], p, who); up(c, c->fld[W2_F_PN], pav, who);
    SHIM_TRY(w2_ptdavg(c, c->fld[W2_F_P], c->fld[W2_F_PN]), who);
    down(c, pav, c->fld[W2_F_PN], who);
    sync(c, who);
}
extern "C" void taveraged_(const int32_t *nx, const int32_t *ny, const int32_t *nScale, const int32_t *nReg,
                           const int32_t *nRegBrd, const int32_t *nTRgType, const double *dTRgVal, const double *t, double *tav) {
    const char *who = "taveraged_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    shim_regions(c, nReg, nRegBrd, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, who);
    SHIM_TRY(w2_set_thermal_tables(c, nTRgType, nullptr, dTRgVal, nullptr), who);
    up(c, c->fld[W2_F_T], t, who); up(c, c->fld[W2_F_TN], tav, who);
    SHIM_TRY(w2_taveraged(c, *nScale, c->fld[W2_F_T], c->fld[W2_F_TN]), who);
    down(c, tav, c->fld[W2_F_TN], who);
    sync(c, who);
}
extern "C" void velavg_(const int32_t *nx, const int32_t *ny, const int32_t *nReg, const int32_t *nRegBrd,
                        const int32_t *nRegType, const double *u, const double *v, double *util, double *vbar) {
    const char *who = "velavg_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    shim_regions(c, nReg, nRegBrd, nRegType, nullptr, nullptr, nullptr, nullptr, nullptr, who);
    up(c, c->fld[W2_F_U], u, who); up(c, c->fld[W2_F_V], v, who);
    up(c, c->fld[W2_F_US], util, who); up(c, c->fld[W2_F_VS], vbar, who);
    SHIM_TRY(w2_velavg(c, c->fld[W2_F_U], c->fld[W2_F_V], c->fld[W2_F_US], c->fld[W2_F_VS]), who);
    down(c, util, c->fld[W2_F_US], who); down(c, vbar, c->fld[W2_F_VS], who);
    sync(c, who);
}

extern "C" void traject_(const int32_t *nx, const int32_t *ny, const int32_t *ntr, const int32_t *ntsubstp,
    const int32_t *nTrMethod, const int32_t *nTrCdEq, const int32_t *mTrHTmit, int32_t *nTOutBnd,
    const double *dkflow, const double *densref, const double *fr, const double *dTrHTtol, const double *dTrHTdel,
    const double *cpartx, const double *cparty, const double *repc,
    const double *x, const double *y, const double *u, const double *v, const double *un, const double *vn,
    const double *dens, const double *densn, double *xp, double *yp, double *up_, double *vp) {
    const char *who = "traject_";
    wolfd2_ctx *c = shim_ctx(*nx, *ny, who);
    wolfd2_traject T;
    memset(&T, 0, sizeof(T));
    T.ntr = *ntr; T.ntsubstp = *ntsubstp; T.nTrMethod = *nTrMethod; T.nTrCdEq = *nTrCdEq; T.mTrHTmit = *mTrHTmit;
    T.densref = *densref; T.dTrHTtol = *dTrHTtol; T.dTrHTdel = *dTrHTdel;
    if (T.ntr <= 0) return;
    SHIM_TRY(w2_traj_set_grid(c, x, y), who);
    SHIM_TRY(w2_traj_set_particles(c, &T, cpartx, cparty, repc, xp, yp, up_, vp, nTOutBnd), who);
    up(c, c->fld[W2_F_U], u, who); up(c, c->fld[W2_F_V], v, who); up(c, c->fld[W2_F_UN], un, who); up(c, c->fld[W2_F_VN], vn, who);
    up(c, c->fld[W2_F_D], dens, who); up(c, c->fld[W2_F_DN], densn, who);
    SHIM_TRY(w2_traject(c, *dkflow, *fr, c->fld[W2_F_U], c->fld[W2_F_V], c->fld[W2_F_UN], c->fld[W2_F_VN], c->fld[W2_F_D],
                        c->fld[W2_F_DN]), who);
    SHIM_TRY(wolfd2_b200_get_particles(c, xp, yp, up_, vp, nTOutBnd), who);
    c->traj->active = 0;   // the shim context is shared: a later step through it must not move these particles
}
