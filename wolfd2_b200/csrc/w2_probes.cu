// w2_probes.cu -- time-series monitor points (SaveTimeSrs case 1, src/file_manip.f:806-834; call site
// src/main.f:984-995) sampled from the resident fields, so that a time series costs 8 numbers per point and
// step instead of a download of the fields.  The host keeps the file side (wolfd2_b200/timeseries.py writes
// the reference's `.ts` layout).
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "w2.cuh"

#define W2_PROBE_CAP 1024   // records held on the device before they are moved to the host

struct W2Probes {
    int n, freq;
    int count;              // records on the device
    long long step;         // steps taken since set_probes (k - ks of main.f:987)
    int *d_i, *d_j;
    double *d_rec;          // W2_PROBE_CAP x n x 8
    std::vector<double> host;      // records already moved to the host
    std::vector<int32_t> steps;    // step number of every record, host and device ones
};

// dTSv(nts,1..8) = u, v, p, t, us, vs, ps, ts at (iTS,jTS) (:810-826); fields a run does not carry give 0
__global__ void probe_kernel(int n, const int *__restrict__ pi, const int *__restrict__ pj, int pitch,
                             const double *u, const double *v, const double *p, const double *t, const double *uss,
                             const double *vss, const double *pss, const double *tss, double *__restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n * 8) return;
    const int pt = k >> 3, var = k & 7;
    const double *f = var == 0 ? u : var == 1 ? v : var == 2 ? p : var == 3 ? t : var == 4 ? uss : var == 5 ? vss : var == 6 ? pss : tss;
    out[k] = f ? f[IDX(pi[pt], pj[pt])] : 0.0;
}

void w2_probes_release(wolfd2_ctx *c) {
    W2Probes *q = (W2Probes *)c->probes;
    if (!q) return;
    cudaFree(q->d_i); cudaFree(q->d_j); cudaFree(q->d_rec);
    delete q;
    c->probes = nullptr;
}

static int probes_flush(wolfd2_ctx *c, W2Probes *q) {
    if (q->count == 0) return W2_OK;
    const size_t n = (size_t)q->count * q->n * 8, old = q->host.size();
    q->host.resize(old + n);
    W2_CUDA(cudaMemcpyAsync(q->host.data() + old, q->d_rec, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    W2_CUDA(cudaStreamSynchronize(c->stream));
    q->count = 0;
    return W2_OK;
}

extern "C" int wolfd2_b200_set_probes(wolfd2_ctx *c, int32_t npoints, const int32_t *iTS, const int32_t *jTS, int32_t freq) {
    if (!c) return W2_ERR_BAD_ARG;
    W2_CUDA(cudaSetDevice(c->device));
    w2_probes_release(c);
    if (npoints <= 0) return W2_OK;
    if (c->world > 1) { w2_set_error("time-series probes are not supported on several GPUs"); return W2_ERR_UNSUPPORTED; }
    if (!iTS || !jTS || freq < 1) { w2_set_error("set_probes: need index arrays and a sampling frequency >= 1"); return W2_ERR_BAD_ARG; }
    W2Probes *q = new W2Probes();
    q->n = npoints; q->freq = freq; q->count = 0; q->step = 0;
    std::vector<int> hi(npoints), hj(npoints);
    for (int k = 0; k < npoints; ++k) {
        hi[k] = iTS[k]; hj[k] = jTS[k];
        if (hi[k] < 1 || hi[k] > c->nx || hj[k] < 1 || hj[k] > c->ny) {   // :764-770
            printf(" Warning: Requested t.s. indeces are not in domain. Using (1,1).\n");
            hi[k] = 1; hj[k] = 1;
        }
    }
    c->probes = q;
    W2_CUDA(cudaMalloc((void **)&q->d_i, npoints * sizeof(int)));
    W2_CUDA(cudaMalloc((void **)&q->d_j, npoints * sizeof(int)));
    W2_CUDA(cudaMalloc((void **)&q->d_rec, (size_t)W2_PROBE_CAP * npoints * 8 * sizeof(double)));
    W2_CUDA(cudaMemcpy(q->d_i, hi.data(), npoints * sizeof(int), cudaMemcpyHostToDevice));
    W2_CUDA(cudaMemcpy(q->d_j, hj.data(), npoints * sizeof(int), cudaMemcpyHostToDevice));
    return W2_OK;
}

// called at the end of every time step (main.f:984-995): sample when mod(k - ks, nTSFreq) == 0
int w2_probes_step(wolfd2_ctx *c) {
    W2Probes *q = (W2Probes *)c->probes;
    if (!q) return W2_OK;
    q->step++;
    if (q->step % q->freq != 0) return W2_OK;
    if (q->count == W2_PROBE_CAP) W2_TRY(probes_flush(c, q));
    const bool ss = c->atd && c->atd->ss.nsmallscl == 1;
    probe_kernel<<<(q->n * 8 + 127) / 128, 128, 0, c->stream>>>(
        q->n, q->d_i, q->d_j, c->pitch, c->fld[W2_F_U], c->fld[W2_F_V], c->fld[W2_F_P], c->fld[W2_F_T],
        ss ? c->fld[W2_F_USS] : nullptr, ss ? c->fld[W2_F_VSS] : nullptr, ss ? c->fld[W2_F_PSS] : nullptr,
        ss ? c->fld[W2_F_TSS] : nullptr, q->d_rec + (size_t)q->count * q->n * 8);
    W2_CUDA(cudaGetLastError());
    c->launches[3]++;
    q->count++;
    q->steps.push_back((int32_t)q->step);
    return W2_OK;
}

// Records sampled since the last call, oldest first: out[rec][point][8] (u, v, p, t, uss, vss, pss, tss) and the step
// number k - ks of each.  *nrec returns how many there are; at most maxrec are copied (call again for the rest).
extern "C" int wolfd2_b200_get_probe_records(wolfd2_ctx *c, int32_t maxrec, double *out, int32_t *steps, int32_t *nrec) {
    if (!c || !nrec) return W2_ERR_BAD_ARG;
    W2Probes *q = (W2Probes *)c->probes;
    if (!q) { *nrec = 0; return W2_OK; }
    W2_CUDA(cudaSetDevice(c->device));
    W2_TRY(probes_flush(c, q));
    const size_t per = (size_t)q->n * 8;
    const int have = (int)(q->host.size() / per);
    *nrec = have;
    if (!out || maxrec <= 0) return W2_OK;
    const int take = have < maxrec ? have : maxrec;
    memcpy(out, q->host.data(), take * per * sizeof(double));
    if (steps) memcpy(steps, q->steps.data(), take * sizeof(int32_t));
    q->host.erase(q->host.begin(), q->host.begin() + take * per);
    q->steps.erase(q->steps.begin(), q->steps.begin() + take);
    *nrec = take;
    return W2_OK;
}
