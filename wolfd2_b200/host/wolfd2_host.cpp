// wolfd2_host.cpp -- compiled host driver over the C ABI (include/wolfd2_b200.h), the C++ stand-in for the
// part of `program wolfd2` (src/main.f) that owns the time loop.  The reference's host language is Fortran,
// which this image cannot compile; this file shows -- and tests/test_gpu_host_driver.py checks -- that the
// library is usable from compiled code with nothing but the header: set-up on the host (uniform grid ->
// MirrorPts/FullGrid/Metric as src/grid.f:258-535, default BC flags as src/parse.f:2257-2379 plus a moving
// lid), then cold start (main.f:606-641) and nsteps of the step body (main.f:690-981) on the device, one
// PrintDiff line per step (src/string.f:547-559).
//
//   usage: wolfd2_host <n> <re> <dt> <nsteps> [ppe_solver_id=5] [sorrel=1.0] [msorit=2000]
//   prints: step, time, QlIt, PpeIt, |p1-pn|, |u1-un|, |v1-vn|   and a final checksum line.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/wolfd2_b200.h"

namespace {
struct Field {   // REAL*8 f(0:mnx,0:mny), zero-initialised like the reference's static storage
    int ld;
    std::vector<double> a;
    Field(int mnx, int mny) : ld(mnx + 1), a((size_t)(mnx + 1) * (mny + 1), 0.0) {}
    double &operator()(int i, int j) { return a[(size_t)i + (size_t)ld * j]; }
    double *p() { return a.data(); }
};
}  // namespace

int main(int argc, char **argv) {
    if (argc < 5) { std::fprintf(stderr, "usage: %s n re dt nsteps [solver] [sorrel] [msorit]\n", argv[0]); return 2; }
    const int n = std::atoi(argv[1]);
    const double re = std::atof(argv[2]), dt = std::atof(argv[3]);
    const int nsteps = std::atoi(argv[4]);
    const int solver = argc > 5 ? std::atoi(argv[5]) : W2_PPE_RB_SOR;
    const double sorrel = argc > 6 ? std::atof(argv[6]) : 1.0;
    const int msorit = argc > 7 ? std::atoi(argv[7]) : 2000;
    const int nx = n, ny = n, mnx = nx + 1, mny = ny + 1, mgri = 20, mgrj = 10;
    if (wolfd2_b200_config(mnx, mny, mgri, mgrj) != W2_OK) return 1;

    // ---- Grid: x(i,j) = (i-1)/(nx-1), y(i,j) = (j-1)/(ny-1), then grid.f
    Field x(mnx, mny), y(mnx, mny), xu(mnx, mny), yu(mnx, mny), xv(mnx, mny), yv(mnx, mny), xc(mnx, mny), yc(mnx, mny);
    for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) { x(i, j) = double(i - 1) / double(nx - 1); y(i, j) = double(j - 1) / double(ny - 1); }
    for (Field *g : {&x, &y}) {   // MirrorPts, grid.f:276-303
        Field &f = *g;
        for (int i = 1; i <= nx; ++i) { f(i, 0) = 2.0 * f(i, 1) - f(i, 2); f(i, ny + 1) = 2.0 * f(i, ny) - f(i, ny - 1); }
        for (int j = 1; j <= ny; ++j) { f(0, j) = 2.0 * f(1, j) - f(2, j); f(nx + 1, j) = 2.0 * f(nx, j) - f(nx - 1, j); }
        f(0, 0) = 2.0 * f(1, 1) - f(2, 2); f(0, ny + 1) = 2.0 * f(1, ny) - f(2, ny - 1);
        f(nx + 1, 0) = 2.0 * f(nx, 1) - f(nx - 1, 2); f(nx + 1, ny + 1) = 2.0 * f(nx, ny) - f(nx - 1, ny - 1);
    }
    for (int j = 1; j <= ny + 1; ++j)   // FullGrid, grid.f:338-359
        for (int i = 0; i <= nx + 1; ++i) { xu(i, j) = 0.5 * (x(i, j) + x(i, j - 1)); yu(i, j) = 0.5 * (y(i, j) + y(i, j - 1)); }
    for (int j = 0; j <= ny + 1; ++j)
        for (int i = 1; i <= nx + 1; ++i) { xv(i, j) = 0.5 * (x(i, j) + x(i - 1, j)); yv(i, j) = 0.5 * (y(i, j) + y(i - 1, j)); }
    for (int j = 1; j <= ny + 1; ++j)
        for (int i = 1; i <= nx + 1; ++i) { xc(i, j) = 0.5 * (xu(i, j) + xu(i - 1, j)); yc(i, j) = 0.5 * (yv(i, j) + yv(i, j - 1)); }
    std::vector<Field> M(30, Field(mnx, mny));   // order of wolfd2_metrics
    enum { RAU, RBU, RBV, RGV, RAN, RBN, RGN, RAC, RBC, RGC, DJU, DJV, DJC, DJN, XEN, YEN, XZN, YZN, XEC, YEC, XZC, YZC,
           XEU, YEU, XZV, YZV, XZU, YZU, XEV, YEV };
    auto tensor = [](double xz, double xe, double yz, double ye, double &dj, double &g11, double &g12, double &g22) {
        dj = 1.0 / (xz * ye - xe * yz); g11 = xz * xz + yz * yz; g12 = xz * xe + yz * ye; g22 = xe * xe + ye * ye;
    };
    for (int j = 1; j <= ny; ++j)   // Metric, grid.f:425-532
        for (int i = 1; i <= nx; ++i) {
            double dj, g11, g12, g22;
            M[XZN](i, j) = xv(i + 1, j) - xv(i, j); M[XEN](i, j) = xu(i, j + 1) - xu(i, j);
            M[YZN](i, j) = yv(i + 1, j) - yv(i, j); M[YEN](i, j) = yu(i, j + 1) - yu(i, j);
            tensor(M[XZN](i, j), M[XEN](i, j), M[YZN](i, j), M[YEN](i, j), dj, g11, g12, g22);
            M[DJN](i, j) = dj; M[RAN](i, j) = dj * g22; M[RBN](i, j) = -dj * g12 * 0.25; M[RGN](i, j) = dj * g11;
            M[XZU](i, j) = xc(i + 1, j) - xc(i, j); M[XEU](i, j) = x(i, j) - x(i, j - 1);
            M[YZU](i, j) = yc(i + 1, j) - yc(i, j); M[YEU](i, j) = y(i, j) - y(i, j - 1);
            tensor(M[XZU](i, j), M[XEU](i, j), M[YZU](i, j), M[YEU](i, j), dj, g11, g12, g22);
            M[DJU](i, j) = dj; M[RAU](i, j) = dj * g22; M[RBU](i, j) = -dj * g12 * 0.25;
            M[XZV](i, j) = x(i, j) - x(i - 1, j); M[XEV](i, j) = xc(i, j + 1) - xc(i, j);
            M[YZV](i, j) = y(i, j) - y(i - 1, j); M[YEV](i, j) = yc(i, j + 1) - yc(i, j);
            tensor(M[XZV](i, j), M[XEV](i, j), M[YZV](i, j), M[YEV](i, j), dj, g11, g12, g22);
            M[DJV](i, j) = dj; M[RBV](i, j) = -dj * g12 * 0.25; M[RGV](i, j) = dj * g11;
            M[XZC](i, j) = xu(i, j) - xu(i - 1, j); M[XEC](i, j) = xv(i, j) - xv(i, j - 1);
            M[YZC](i, j) = yu(i, j) - yu(i - 1, j); M[YEC](i, j) = yv(i, j) - yv(i, j - 1);
            tensor(M[XZC](i, j), M[XEC](i, j), M[YZC](i, j), M[YEC](i, j), dj, g11, g12, g22);
            M[DJC](i, j) = dj; M[RAC](i, j) = dj * g22; M[RBC](i, j) = -dj * g12 * 0.25; M[RGC](i, j) = dj * g11;
        }
    // ---- SetUpBCs: one region, no-slip walls, lid `wall 1 1 n tangent_vel 1.0`
    std::vector<int32_t> nReg{1, 1}, brd(mgri * mgrj * 4, 0), typ(mgri * mgrj, 0), mom(mgri * mgrj * 4, 0);
    std::vector<double> val(mgri * mgrj * 16, 0.0), poros(mgri * mgrj, 0.0), c1(mgri * mgrj, 0.0), c2(mgri * mgrj, 0.0);
    const int plane = mgri * mgrj;
    brd[plane * (W2_WEST - 1)] = 1; brd[plane * (W2_EAST - 1)] = nx; brd[plane * (W2_SOUTH - 1)] = 1; brd[plane * (W2_NORTH - 1)] = ny;
    typ[0] = W2_RM_INTERN;
    for (int k = 0; k < 4; ++k) mom[plane * k] = W2_BM_WALL1;
    val[plane * ((W2_NORTH - 1) + 4 * (W2_U - 1))] = 1.0;   // dBCVal(1,1,NORTH,_U_)
    poros[0] = 1.0;

    wolfd2_params par{};
    par.nx = nx; par.ny = ny; par.mqiter = 20; par.nmeiter = 1; par.nPpeSolver = solver; par.msorit = msorit;
    par.lCartesGrid = 1; par.dk = dt; par.re = re; par.fr = 1.0 / 9.81; par.qtol = 1e-4; par.sortol = 1e-8; par.sorrel = sorrel;
    wolfd2_regions reg{nReg.data(), brd.data(), typ.data(), mom.data(), val.data(), poros.data(), c1.data(), c2.data()};
    wolfd2_metrics met{};
    const double **mp = &met.rau;
    for (int k = 0; k < 30; ++k) mp[k] = M[k].p();

    wolfd2_ctx *ctx = nullptr;
    if (wolfd2_b200_create(&ctx, &par, &reg, &met) != W2_OK) { std::fprintf(stderr, "%s\n", wolfd2_b200_last_error()); return 1; }
    Field u(mnx, mny), v(mnx, mny), p(mnx, mny);   // InitialCond cold start, bound_cond.f:484-499
    wolfd2_b200_upload_field(ctx, W2_F_U, u.p()); wolfd2_b200_upload_field(ctx, W2_F_V, v.p()); wolfd2_b200_upload_field(ctx, W2_F_P, p.p());
    int32_t nsor = 0;
    if (wolfd2_b200_coldstart(ctx, &nsor) != W2_OK) { std::fprintf(stderr, "%s\n", wolfd2_b200_last_error()); return 1; }
    double time = 0.0;
    for (int k = 1; k <= nsteps; ++k) {
        wolfd2_step_log lg;
        const int rc = wolfd2_b200_step(ctx, 1, &lg);
        time += dt;
        std::printf(" %6d%13.6E%4d%c%6d%13.6E%13.6E%13.6E\n", k, time, lg.nQLiter, lg.nQLiter < 0 ? '*' : ' ', lg.nSorConv,
                    lg.dif[0], lg.dif[1], lg.dif[2]);
        if (rc != W2_OK) { std::fprintf(stderr, "%s\n", wolfd2_b200_last_error()); return 1; }
    }
    wolfd2_b200_download_field(ctx, W2_F_U, u.p()); wolfd2_b200_download_field(ctx, W2_F_V, v.p()); wolfd2_b200_download_field(ctx, W2_F_P, p.p());
    double su = 0, sv = 0, sp = 0;
    for (size_t q = 0; q < u.a.size(); ++q) { su += u.a[q]; sv += v.a[q]; sp += p.a[q]; }
    std::printf("checksum %.17g %.17g %.17g\n", su, sv, sp);
    wolfd2_b200_destroy(ctx);
    return 0;
}
