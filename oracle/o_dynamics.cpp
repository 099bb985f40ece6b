/* ORACLE (test infrastructure) — restatement of prognostics.f90, geopotential.f90,
 * horizontal_diffusion.f90, implicit.f90, matrix_inversion.f90, tendencies.f90,
 * time_stepping.f90 and diagnostics.f90 of the reference.  Each function cites the lines
 * it follows.  Not re-entrant (large work arrays are function-static, like the
 * reference's fixed-size locals). */
#include "oracle.h"

namespace orc {

#define GLOOP for (int j = 1; j <= il; j++) for (int i = 1; i <= ix; i++)
#define SLOOP for (int n = 1; n <= nx; n++) for (int m = 1; m <= mx; m++)

/* ---------------------------------------------------------------- prognostics.f90:16-24 */
Spec4 vor, div_, t;
FA<cplx, mx, nx, 2> ps;
FA<cplx, mx, nx, kx, 2, ntr> tr;
Spec3 phi;
Spec2 phis;

/* ---------------------------------------------------------------- geopotential.f90 */
double xgeop1[kx + 1], xgeop2[kx + 1];

/* geopotential.f90:19-29 */
void initialize_geopotential() {
    for (int k = 1; k <= kx; k++) {
        xgeop1[k] = rgas * log(geo.hsg[k + 1] / geo.fsg[k]);
        if (k != kx) xgeop2[k + 1] = rgas * log(geo.fsg[k + 1] / geo.hsg[k + 1]);
    }
}

#define S3(a, m, n, k) a[((m)-1) + (size_t)mx * (((n)-1) + (size_t)nx * ((k)-1))]
#define S2(a, m, n) a[((m)-1) + (size_t)mx * ((n)-1)]

/* geopotential.f90:33-57 */
void get_geopotential(const cplx* tt, const cplx* phis_, cplx* phi_) {
    SLOOP S3(phi_, m, n, kx) = S2(phis_, m, n) + xgeop1[kx] * S3(tt, m, n, kx);
    for (int k = kx - 1; k >= 1; k--)
        SLOOP S3(phi_, m, n, k) = S3(phi_, m, n, k + 1) + xgeop2[k + 1] * S3(tt, m, n, k + 1) + xgeop1[k] * S3(tt, m, n, k);
    for (int k = 2; k <= kx - 1; k++) {
        double corf = xgeop1[k] * 0.5 * log(geo.hsg[k + 1] / geo.fsg[k]) / log(geo.fsg[k + 1] / geo.fsg[k - 1]);
        for (int n = 1; n <= nx; n++) S3(phi_, 1, n, k) = S3(phi_, 1, n, k) + corf * (S3(tt, 1, n, k + 1) - S3(tt, 1, n, k - 1));
    }
}

/* ---------------------------------------------------------------- horizontal_diffusion.f90 */
FA<double, mx, nx> dmp, dmpd, dmps, dmp1, dmp1d, dmp1s;
double tcorv[kx + 1], qcorv[kx + 1];
Spec2 tcorh, qcorh;

/* horizontal_diffusion.f90:36-82 */
void initialize_horizontal_diffusion() {
    const int npowhd = 4;
    double hdiff = 1. / (thd * 3600.);
    double hdifd = 1. / (thdd * 3600.);
    double hdifs = 1. / (thds * 3600.);
    double rlap = (double)(1.f / (float)(trunc_ * (trunc_ + 1)));   /* :55 real32 quotient */
    for (int j = 1; j <= nx; j++)
        for (int k = 1; k <= mx; k++) {
            double twn = (double)(float)(k + j - 2);
            double elap = (twn * (twn + 1.) * rlap);
            static_assert(npowhd == 4, "elap**npowhd below is written for npowhd = 4");
            double elapn = (elap * elap) * (elap * elap);   /* elap**4, integer power */
            dmp(k, j) = hdiff * elapn;
            dmpd(k, j) = hdifd * elapn;
            dmps(k, j) = hdifs * elap;
        }
    double rgam = rgas * gamma_ / (1000. * grav);
    double qexp = hscale / hshum;
    tcorv[1] = 0.;
    qcorv[1] = 0.;
    qcorv[2] = 0.;
    for (int k = 2; k <= kx; k++) {
        tcorv[k] = pow(geo.fsg[k], rgam);
        if (k > 2) qcorv[k] = pow(geo.fsg[k], qexp);
    }
}

/* horizontal_diffusion.f90:86-105 (2-D form on one level) */
static inline void do_horizontal_diffusion_2d(const cplx* field, cplx* fdt, const FA<double, mx, nx>& d, const FA<double, mx, nx>& d1) {
    SLOOP S2(fdt, m, n) = (S2(fdt, m, n) - d(m, n) * S2(field, m, n)) * d1(m, n);
}
static void do_horizontal_diffusion_3d(const cplx* field, Spec3& fdt, const FA<double, mx, nx>& d, const FA<double, mx, nx>& d1) {
    for (int k = 1; k <= kx; k++) do_horizontal_diffusion_2d(field + (size_t)mx * nx * (k - 1), fdt.p(1, 1, k), d, d1);
}

/* ---------------------------------------------------------------- matrix_inversion.f90 */
/* :12-73 */
static void ludcmp(double* a, int n, int np, int* indx, double& d) {
#define A(i, j) a[((i)-1) + (size_t)np * ((j)-1)]
    const double tiny = (double)1.0e-20f;
    double vv[101], aamax, dum, sum;
    int imax = 0;
    d = 1.0;
    for (int i = 1; i <= n; i++) {
        aamax = 0.;
        for (int j = 1; j <= n; j++)
            if (fabs(A(i, j)) > aamax) aamax = fabs(A(i, j));
        if (aamax == 0.) { fprintf(stderr, "singular\n"); abort(); }
        vv[i] = 1. / aamax;
    }
    for (int j = 1; j <= n; j++) {
        if (j > 1) {
            for (int i = 1; i <= j - 1; i++) {
                sum = A(i, j);
                if (i > 1) {
                    for (int k = 1; k <= i - 1; k++) sum = sum - A(i, k) * A(k, j);
                    A(i, j) = sum;
                }
            }
        }
        aamax = 0.;
        for (int i = j; i <= n; i++) {
            sum = A(i, j);
            if (j > 1) {
                for (int k = 1; k <= j - 1; k++) sum = sum - A(i, k) * A(k, j);
                A(i, j) = sum;
            }
            dum = vv[i] * fabs(sum);
            if (dum >= aamax) {
                imax = i;
                aamax = dum;
            }
        }
        if (j != imax) {
            for (int k = 1; k <= n; k++) {
                dum = A(imax, k);
                A(imax, k) = A(j, k);
                A(j, k) = dum;
            }
            d = -d;
            vv[imax] = vv[j];
        }
        indx[j] = imax;
        if (j != n) {
            if (A(j, j) == 0) A(j, j) = tiny;
            dum = 1. / A(j, j);
            for (int i = j + 1; i <= n; i++) A(i, j) = A(i, j) * dum;
        }
    }
    if (A(n, n) == 0.) A(n, n) = tiny;
}
/* :75-103 */
static void lubksb(const double* a, int n, int np, const int* indx, double* b /*1-based*/) {
    int ii = 0;
    for (int i = 1; i <= n; i++) {
        int ll = indx[i];
        double sum = b[ll];
        b[ll] = b[i];
        if (ii != 0) {
            for (int j = ii; j <= i - 1; j++) sum = sum - A(i, j) * b[j];
        } else if (sum != 0) {
            ii = i;
        }
        b[i] = sum;
    }
    for (int i = n; i >= 1; i--) {
        double sum = b[i];
        if (i < n)
            for (int j = i + 1; j <= n; j++) sum = sum - A(i, j) * b[j];
        b[i] = sum / A(i, i);
    }
#undef A
}
/* :115-133 */
static void inv(double* a, double* y, int* indx, int n) {
    double d;
    for (int q = 0; q < n * n; q++) y[q] = 0.0;
    for (int i = 1; i <= n; i++) y[(i - 1) + (size_t)n * (i - 1)] = 1.;
    ludcmp(a, n, n, indx, d);
    for (int i = 1; i <= n; i++) lubksb(a, n, n, indx, y + (size_t)n * (i - 1) - 1);
}

/* ---------------------------------------------------------------- implicit.f90 */
double tref[kx + 1], tref1[kx + 1], tref2[kx + 1], tref3[kx + 1], dhsx[kx + 1];
FA<double, kx, kx> xa, xb, xc, xd, xe;
FA<double, kx, kx, mx + nx + 1> xf, xj;
FA<double, mx, nx> elz;

/* implicit.f90:36-165 */
void initialize_implicit(double dt) {
    double dsum[kx + 1];
    FA<double, kx, kx> ya;
    int indx[kx + 1];
    for (int m = 1; m <= mx; m++)
        for (int n = 1; n <= nx; n++) {
            dmp1(m, n) = 1. / (1. + dmp(m, n) * dt);
            dmp1d(m, n) = 1. / (1. + dmpd(m, n) * dt);
            dmp1s(m, n) = 1. / (1. + dmps(m, n) * dt);
        }
    double rgam = rgas * gamma_ / (1000. * grav);
    for (int k = 1; k <= kx; k++) {
        tref[k] = 288. * pow(std::max((double)0.2f, geo.fsg[k]), rgam);   /* :63 max(real32 0.2, real64) */
        tref1[k] = rgas * tref[k];
        tref2[k] = akap * tref[k];
        tref3[k] = geo.fsgr[k] * tref[k];
    }
    double xi = dt * alph;
    double xxi = xi / (rearth * rearth);
    for (int k = 1; k <= kx; k++) dhsx[k] = xi * geo.dhs[k];
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++) elz(m, n) = (double)(float)(m + n - 2) * (double)(float)(m + n - 1) * xxi;
    /* :88 xa(:kx,:kx-1) = 0 ; column kx of xa is never referenced */
    for (int k = 1; k <= kx; k++)
        for (int k1 = 1; k1 <= kx - 1; k1++) xa(k, k1) = 0.0;
    for (int k = 1; k <= kx; k++)
        for (int k1 = 1; k1 <= kx; k1++) ya(k, k1) = -akap * tref[k] * geo.dhs[k1];
    for (int k = 2; k <= kx; k++) xa(k, k - 1) = 0.5 * (akap * tref[k] / geo.fsg[k] - (tref[k] - tref[k - 1]) / geo.dhs[k]);
    for (int k = 1; k <= kx - 1; k++) xa(k, k) = 0.5 * (akap * tref[k] / geo.fsg[k] - (tref[k + 1] - tref[k]) / geo.dhs[k]);
    dsum[1] = geo.dhs[1];
    for (int k = 2; k <= kx; k++) dsum[k] = dsum[k - 1] + geo.dhs[k];
    for (int k = 1; k <= kx - 1; k++)
        for (int k1 = 1; k1 <= kx; k1++) {
            xb(k, k1) = geo.dhs[k1] * dsum[k];
            if (k1 <= k) xb(k, k1) = xb(k, k1) - geo.dhs[k1];
        }
    for (int k = 1; k <= kx; k++)
        for (int k1 = 1; k1 <= kx; k1++) {
            xc(k, k1) = ya(k, k1);
            for (int k2 = 1; k2 <= kx - 1; k2++) xc(k, k1) = xc(k, k1) + xa(k, k2) * xb(k2, k1);
        }
    xd.fill(0.0);
    for (int k = 1; k <= kx; k++)
        for (int k1 = k + 1; k1 <= kx; k1++) xd(k, k1) = rgas * log(geo.hsg[k1 + 1] / geo.hsg[k1]);
    for (int k = 1; k <= kx; k++) xd(k, k) = rgas * log(geo.hsg[k + 1] / geo.fsg[k]);
    for (int k = 1; k <= kx; k++)
        for (int k1 = 1; k1 <= kx; k1++) {
            xe(k, k1) = 0.;
            for (int k2 = 1; k2 <= kx; k2++) xe(k, k1) = xe(k, k1) + xd(k, k2) * xc(k2, k1);
        }
    for (int l = 1; l <= mx + nx + 1; l++) {
        double xxx = ((double)(float)l * (double)(float)(l + 1)) / (rearth * rearth);
        for (int k = 1; k <= kx; k++)
            for (int k1 = 1; k1 <= kx; k1++) xf(k, k1, l) = xi * xi * xxx * (rgas * tref[k] * geo.dhs[k1] - xe(k, k1));
        for (int k = 1; k <= kx; k++) xf(k, k, l) = xf(k, k, l) + 1.;
    }
    for (int l = 1; l <= mx + nx + 1; l++) inv(xf.p(1, 1, l), xj.p(1, 1, l), indx, kx);
    for (int k = 1; k <= kx; k++)
        for (int k1 = 1; k1 <= kx; k1++) xc(k, k1) = xc(k, k1) * xi;
}

/* implicit.f90:168-217 */
void implicit_terms(Spec3& divdt, Spec3& tdt, Spec2& psdt) {
    static Spec3 ye, yf;
    ye.fill(cplx(0.0, 0.0));
    for (int k1 = 1; k1 <= kx; k1++)
        for (int k = 1; k <= kx; k++)
            SLOOP ye(m, n, k) = ye(m, n, k) + xd(k, k1) * tdt(m, n, k1);
    for (int k = 1; k <= kx; k++)
        SLOOP ye(m, n, k) = ye(m, n, k) + tref1[k] * psdt(m, n);
    for (int k = 1; k <= kx; k++)
        for (int m = 1; m <= mx; m++)
            for (int n = 1; n <= nx; n++) yf(m, n, k) = divdt(m, n, k) + elz(m, n) * ye(m, n, k);
    divdt.fill(cplx(0.0, 0.0));
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++)
            if ((m + n - 2) != 0)
                for (int k1 = 1; k1 <= kx; k1++)
                    for (int k = 1; k <= kx; k++) divdt(m, n, k) = divdt(m, n, k) + xj(k, k1, m + n - 2) * yf(m, n, k1);
    for (int k = 1; k <= kx; k++)
        SLOOP psdt(m, n) = psdt(m, n) - divdt(m, n, k) * dhsx[k];
    for (int k = 1; k <= kx; k++)
        for (int k1 = 1; k1 <= kx; k1++)
            SLOOP tdt(m, n, k) = tdt(m, n, k) + xc(k, k1) * divdt(m, n, k1);
}

/* ---------------------------------------------------------------- tendencies.f90 */
/* tendencies.f90:49-235 */
static void get_grid_point_tendencies(Spec3& vordt, Spec3& divdt, Spec3& tdt, Spec2& psdt, FA<cplx, mx, nx, kx, ntr>& trdt, int j1, int j2) {
    static FA<cplx, mx, nx, 2> dumc;
    static Grid3 utend, vtend, ttend;
    static FA<double, ix, il, kx, ntr> trtend, trg;
    static Grid3 ug, vg, tg, vorg, divg, tgg, puv;
    static Grid2 px, py, umean, vmean, dmean, gtmp, gtmp2;
    static FA<double, ix, il, kx + 1> sigdt, temp, sigm;
    static Spec2 stmp, stmp2;

    /* :89-107 */
    for (int k = 1; k <= kx; k++) {
        spec_to_grid(vor.p(1, 1, k, j2), 1, vorg.p(1, 1, k));
        spec_to_grid(div_.p(1, 1, k, j2), 1, divg.p(1, 1, k));
        spec_to_grid(t.p(1, 1, k, j2), 1, tg.p(1, 1, k));
        for (int itr = 1; itr <= ntr; itr++) spec_to_grid(tr.p(1, 1, k, j2, itr), 1, trg.p(1, 1, k, itr));
        uvspec(vor.p(1, 1, k, j2), div_.p(1, 1, k, j2), dumc.p(1, 1, 1), dumc.p(1, 1, 2));
        spec_to_grid(dumc.p(1, 1, 2), 2, vg.p(1, 1, k));
        spec_to_grid(dumc.p(1, 1, 1), 2, ug.p(1, 1, k));
        GLOOP vorg(i, j, k) = vorg(i, j, k) + geo.coriol[j];
    }
    /* :109-117 */
    umean.fill(0.0); vmean.fill(0.0); dmean.fill(0.0);
    for (int k = 1; k <= kx; k++)
        GLOOP {
            umean(i, j) = umean(i, j) + ug(i, j, k) * geo.dhs[k];
            vmean(i, j) = vmean(i, j) + vg(i, j, k) * geo.dhs[k];
            dmean(i, j) = dmean(i, j) + divg(i, j, k) * geo.dhs[k];
        }
    /* :121-126 */
    grad(ps.p(1, 1, j2), dumc.p(1, 1, 1), dumc.p(1, 1, 2));
    spec_to_grid(dumc.p(1, 1, 1), 2, px.p());
    spec_to_grid(dumc.p(1, 1, 2), 2, py.p());
    GLOOP gtmp(i, j) = -umean(i, j) * px(i, j) - vmean(i, j) * py(i, j);
    grid_to_spec(gtmp.p(), psdt.p());
    psdt(1, 1) = cplx(0.0, 0.0);
    /* :129-143 */
    GLOOP { sigdt(i, j, 1) = 0.0; sigdt(i, j, kx + 1) = 0.0; sigm(i, j, 1) = 0.0; sigm(i, j, kx + 1) = 0.0; }
    for (int k = 1; k <= kx; k++)
        GLOOP puv(i, j, k) = (ug(i, j, k) - umean(i, j)) * px(i, j) + (vg(i, j, k) - vmean(i, j)) * py(i, j);
    for (int k = 1; k <= kx; k++)
        GLOOP {
            sigdt(i, j, k + 1) = sigdt(i, j, k) - geo.dhs[k] * (puv(i, j, k) + divg(i, j, k) - dmean(i, j));
            sigm(i, j, k + 1) = sigm(i, j, k) - geo.dhs[k] * puv(i, j, k);
        }
    /* :147-149 */
    for (int k = 1; k <= kx; k++)
        GLOOP tgg(i, j, k) = tg(i, j, k) - tref[k];
    /* :152-163 zonal wind tendency */
    GLOOP { temp(i, j, 1) = 0.0; temp(i, j, kx + 1) = 0.0; }
    for (int k = 2; k <= kx; k++)
        GLOOP temp(i, j, k) = sigdt(i, j, k) * (ug(i, j, k) - ug(i, j, k - 1));
    for (int k = 1; k <= kx; k++)
        GLOOP utend(i, j, k) = vg(i, j, k) * vorg(i, j, k) - tgg(i, j, k) * rgas * px(i, j) - (temp(i, j, k + 1) + temp(i, j, k)) * geo.dhsr[k];
    /* :165-173 meridional wind tendency */
    for (int k = 2; k <= kx; k++)
        GLOOP temp(i, j, k) = sigdt(i, j, k) * (vg(i, j, k) - vg(i, j, k - 1));
    for (int k = 1; k <= kx; k++)
        GLOOP vtend(i, j, k) = -ug(i, j, k) * vorg(i, j, k) - tgg(i, j, k) * rgas * py(i, j) - (temp(i, j, k + 1) + temp(i, j, k)) * geo.dhsr[k];
    /* :175-185 temperature tendency */
    for (int k = 2; k <= kx; k++)
        GLOOP temp(i, j, k) = sigdt(i, j, k) * (tgg(i, j, k) - tgg(i, j, k - 1)) + sigm(i, j, k) * (tref[k] - tref[k - 1]);
    for (int k = 1; k <= kx; k++)
        GLOOP ttend(i, j, k) = tgg(i, j, k) * divg(i, j, k) - (temp(i, j, k + 1) + temp(i, j, k)) * geo.dhsr[k]
                               + geo.fsgr[k] * tgg(i, j, k) * (sigdt(i, j, k + 1) + sigdt(i, j, k))
                               + tref3[k] * (sigm(i, j, k + 1) + sigm(i, j, k))
                               + akap * (tg(i, j, k) * puv(i, j, k) - tgg(i, j, k) * dmean(i, j));
    /* :187-197 tracer tendency */
    for (int itr = 1; itr <= ntr; itr++) {
        for (int k = 2; k <= kx; k++)
            GLOOP temp(i, j, k) = sigdt(i, j, k) * (trg(i, j, k, itr) - trg(i, j, k - 1, itr));
        GLOOP { temp(i, j, 2) = 0.0; temp(i, j, 3) = 0.0; }
        for (int k = 1; k <= kx; k++)
            GLOOP trtend(i, j, k, itr) = trg(i, j, k, itr) * divg(i, j, k) - (temp(i, j, k + 1) + temp(i, j, k)) * geo.dhsr[k];
    }
    /* :203-206 physics */
    get_geopotential(t.p(1, 1, 1, j1), phis.p(), phi.p());
    {
        /* trtend(ix,il,kx,ntr) with ntr = 1 is passed as qtend(ix,il,kx) */
        static Grid3 qtend;
        memcpy(qtend.p(), trtend.p(), sizeof(double) * qtend.size());
        get_physical_tendencies(vor.p(1, 1, 1, j1), div_.p(1, 1, 1, j1), t.p(1, 1, 1, j1), tr.p(1, 1, 1, j1, 1),
                                phi.p(), ps.p(1, 1, j1), utend, vtend, ttend, qtend);
        memcpy(trtend.p(), qtend.p(), sizeof(double) * qtend.size());
    }
    /* :212-234 */
    for (int k = 1; k <= kx; k++) {
        vdspec(utend.p(1, 1, k), vtend.p(1, 1, k), vordt.p(1, 1, k), divdt.p(1, 1, k), 2);
        GLOOP gtmp(i, j) = 0.5 * (ug(i, j, k) * ug(i, j, k) + vg(i, j, k) * vg(i, j, k));
        grid_to_spec(gtmp.p(), stmp.p());
        laplacian(stmp.p(), stmp2.p());
        SLOOP divdt(m, n, k) = divdt(m, n, k) - stmp2(m, n);
        GLOOP { gtmp(i, j) = -ug(i, j, k) * tgg(i, j, k); gtmp2(i, j) = -vg(i, j, k) * tgg(i, j, k); }
        vdspec(gtmp.p(), gtmp2.p(), dumc.p(1, 1, 1), tdt.p(1, 1, k), 2);
        grid_to_spec(ttend.p(1, 1, k), stmp.p());
        SLOOP tdt(m, n, k) = tdt(m, n, k) + stmp(m, n);
        for (int itr = 1; itr <= ntr; itr++) {
            GLOOP { gtmp(i, j) = -ug(i, j, k) * trg(i, j, k, itr); gtmp2(i, j) = -vg(i, j, k) * trg(i, j, k, itr); }
            vdspec(gtmp.p(), gtmp2.p(), dumc.p(1, 1, 1), trdt.p(1, 1, k, itr), 2);
            grid_to_spec(trtend.p(1, 1, k, itr), stmp.p());
            SLOOP trdt(m, n, k, itr) = trdt(m, n, k, itr) + stmp(m, n);
        }
    }
}

/* tendencies.f90:242-293 */
static void get_spectral_tendencies(Spec3& divdt, Spec3& tdt, Spec2& psdt, int j2) {
    static FA<cplx, mx, nx, kx + 1> dumk, sigdtc;
    static Spec2 dmeanc, stmp, stmp2;
    dmeanc.fill(cplx(0.0, 0.0));
    for (int k = 1; k <= kx; k++)
        SLOOP dmeanc(m, n) = dmeanc(m, n) + div_(m, n, k, j2) * geo.dhs[k];
    SLOOP psdt(m, n) = psdt(m, n) - dmeanc(m, n);
    psdt(1, 1) = cplx(0.0, 0.0);
    SLOOP { sigdtc(m, n, 1) = cplx(0.0, 0.0); sigdtc(m, n, kx + 1) = cplx(0.0, 0.0); }
    for (int k = 1; k <= kx - 1; k++)
        SLOOP sigdtc(m, n, k + 1) = sigdtc(m, n, k) - geo.dhs[k] * (div_(m, n, k, j2) - dmeanc(m, n));
    SLOOP { dumk(m, n, 1) = cplx(0.0, 0.0); dumk(m, n, kx + 1) = cplx(0.0, 0.0); }
    for (int k = 2; k <= kx; k++)
        SLOOP dumk(m, n, k) = sigdtc(m, n, k) * (tref[k] - tref[k - 1]);
    for (int k = 1; k <= kx; k++)
        SLOOP tdt(m, n, k) = tdt(m, n, k) - (dumk(m, n, k + 1) + dumk(m, n, k)) * geo.dhsr[k]
                             + tref3[k] * (sigdtc(m, n, k + 1) + sigdtc(m, n, k)) - tref2[k] * dmeanc(m, n);
    get_geopotential(t.p(1, 1, 1, j2), phis.p(), phi.p());
    for (int k = 1; k <= kx; k++) {
        SLOOP stmp(m, n) = phi(m, n, k) + rgas * tref[k] * ps(m, n, j2);
        laplacian(stmp.p(), stmp2.p());
        SLOOP divdt(m, n, k) = divdt(m, n, k) - stmp2(m, n);
    }
}

/* tendencies.f90:11-37 */
void get_tendencies(Spec3& vordt, Spec3& divdt, Spec3& tdt, Spec2& psdt, FA<cplx, mx, nx, kx, ntr>& trdt, int j2) {
    get_grid_point_tendencies(vordt, divdt, tdt, psdt, trdt, 1, j2);
    if (alph < 0.5) {
        get_spectral_tendencies(divdt, tdt, psdt, j2);
    } else {
        get_spectral_tendencies(divdt, tdt, psdt, 1);
        implicit_terms(divdt, tdt, psdt);
    }
}

/* ---------------------------------------------------------------- time_stepping.f90 */
/* time_stepping.f90:141-167; input/output are the two time levels of one 2-D field */
static void step_field_2d(int j1, double dt, double eps, cplx* f1, cplx* f2, cplx* fdt) {
    static Spec2 fnew;
    cplx* lev[3] = {nullptr, f1, f2};
    if (ix == iy * 4) trunct(fdt);
    SLOOP fnew(m, n) = S2(f1, m, n) + dt * S2(fdt, m, n);
    /* :163 — level 1 is updated first; :166 then uses the updated level 1 (and, for j1 == 1, the updated output(:,:,j1)) */
    SLOOP S2(f1, m, n) = S2(lev[j1], m, n) + wil * eps * (S2(f1, m, n) - 2.0 * S2(lev[j1], m, n) + fnew(m, n));
    SLOOP S2(f2, m, n) = fnew(m, n) - (1.0 - wil) * eps * (S2(f1, m, n) - 2.0 * S2(lev[j1], m, n) + fnew(m, n));
}
/* time_stepping.f90:127-139 */
static void step_field_3d(int j1, double dt, double eps, cplx* f /*(mx,nx,kx,2)*/, cplx* fdt /*(mx,nx,kx)*/) {
    const size_t lev = (size_t)mx * nx * kx, sl = (size_t)mx * nx;
    for (int k = 1; k <= kx; k++) step_field_2d(j1, dt, eps, f + sl * (k - 1), f + lev + sl * (k - 1), fdt + sl * (k - 1));
}

/* time_stepping.f90:35-122 */
void step(int j1, int j2, double dt) {
    static Spec3 vordt, divdt, tdt, ctmp;
    static Spec2 psdt;
    static FA<cplx, mx, nx, kx, ntr> trdt;
    get_tendencies(vordt, divdt, tdt, psdt, trdt, j2);
    /* :63-74 */
    do_horizontal_diffusion_3d(vor.p(1, 1, 1, 1), vordt, dmp, dmp1);
    do_horizontal_diffusion_3d(div_.p(1, 1, 1, 1), divdt, dmpd, dmp1d);
    for (int k = 1; k <= kx; k++)
        for (int m = 1; m <= mx; m++)
            for (int n = 1; n <= nx; n++) ctmp(m, n, k) = t(m, n, k, 1) + tcorh(m, n) * tcorv[k];
    do_horizontal_diffusion_3d(ctmp.p(), tdt, dmp, dmp1);
    /* :77-85 */
    double sdrag = 1.0 / (tdrs * 3600.0);
    for (int n = 1; n <= nx; n++) {
        vordt(1, n, 1) = vordt(1, n, 1) - sdrag * vor(1, n, 1, 1);
        divdt(1, n, 1) = divdt(1, n, 1) - sdrag * div_(1, n, 1, 1);
    }
    do_horizontal_diffusion_3d(vor.p(1, 1, 1, 1), vordt, dmps, dmp1s);
    do_horizontal_diffusion_3d(div_.p(1, 1, 1, 1), divdt, dmps, dmp1s);
    do_horizontal_diffusion_3d(ctmp.p(), tdt, dmps, dmp1s);
    /* :88-96 */
    for (int k = 1; k <= kx; k++)
        for (int m = 1; m <= mx; m++)
            for (int n = 1; n <= nx; n++) ctmp(m, n, k) = tr(m, n, k, 1, 1) + qcorh(m, n) * qcorv[k];
    {
        static Spec3 trdt1;
        memcpy(trdt1.p(), trdt.p(1, 1, 1, 1), sizeof(cplx) * trdt1.size());
        do_horizontal_diffusion_3d(ctmp.p(), trdt1, dmpd, dmp1d);
        memcpy(trdt.p(1, 1, 1, 1), trdt1.p(), sizeof(cplx) * trdt1.size());
    }
    /* :108-121 */
    double eps = (j1 == 1) ? 0.0 : rob;
    step_field_2d(j1, dt, eps, ps.p(1, 1, 1), ps.p(1, 1, 2), psdt.p());
    step_field_3d(j1, dt, eps, vor.p(), vordt.p());
    step_field_3d(j1, dt, eps, div_.p(), divdt.p());
    step_field_3d(j1, dt, eps, t.p(), tdt.p());
    for (int itr = 1; itr <= ntr; itr++) step_field_3d(j1, dt, eps, tr.p(1, 1, 1, 1, itr), trdt.p(1, 1, 1, itr));
}

/* time_stepping.f90:12-24 */
void first_step() {
    initialize_implicit(0.5 * delt);
    step(1, 1, 0.5 * delt);
    initialize_implicit(delt);
    step(1, 2, delt);
    initialize_implicit(2 * delt);
}

/* ---------------------------------------------------------------- diagnostics.f90:16-75 */
int check_diagnostics(const cplx* vor_, const cplx* divv, const cplx* tt, int istep, double* diag, bool print) {
#define D(k, c) diag[((k)-1) + kx * ((c)-1)]
    static Spec2 temp;
    for (int k = 1; k <= kx; k++) {
        D(k, 1) = 0.0;
        D(k, 2) = 0.0;
        D(k, 3) = (double)sqrtf(0.5f) * S3(tt, 1, 1, k).real();
        inverse_laplacian(vor_ + (size_t)mx * nx * (k - 1), temp.p());
        for (int m = 2; m <= mx; m++)
            for (int n = 1; n <= nx; n++) D(k, 1) = D(k, 1) - (temp(m, n) * std::conj(S3(vor_, m, n, k))).real();
        inverse_laplacian(divv + (size_t)mx * nx * (k - 1), temp.p());
        for (int m = 2; m <= mx; m++)
            for (int n = 1; n <= nx; n++) D(k, 2) = D(k, 2) - (temp(m, n) * std::conj(S3(divv, m, n, k))).real();
    }
    auto pr = [&]() {
        printf(" step =%6d reke =", istep); for (int k = 1; k <= kx; k++) printf("%8.2f", D(k, 1)); printf("\n");
        printf("             %s", " deke ="); for (int k = 1; k <= kx; k++) printf("%8.2f", D(k, 2)); printf("\n");
        printf("             %s", " temp ="); for (int k = 1; k <= kx; k++) printf("%8.2f", D(k, 3)); printf("\n");
    };
    if (print) pr();
    for (int k = 1; k <= kx; k++)
        if (D(k, 1) > 500.0 || D(k, 2) > 500.0 || D(k, 3) < 180.0 || D(k, 3) > 320.0) {
            if (!print) pr();
            return 1;
        }
    return 0;
#undef D
}

/* ---------------------------------------------------------------- prognostics.f90:34-127 */
void initialize_prognostics() {
    static Spec2 surfs;
    static Grid2 surfg;
    double gam1 = gamma_ / (1000.0 * grav);
    grid_to_spec(phis0.p(), phis.p());
    /* :54-56 (time level 2 is left as the zero-initialised static storage) */
    vor.fill(cplx(0.0, 0.0)); div_.fill(cplx(0.0, 0.0)); tr.fill(cplx(0.0, 0.0)); t.fill(cplx(0.0, 0.0)); ps.fill(cplx(0.0, 0.0));
    double tref_ = 288.0, ttop = 216.0;
    double gam2 = gam1 / tref_;
    double rgam = rgas * gam1;
    double rgamr = 1.0 / rgam;
    SLOOP surfs(m, n) = -gam1 * phis(m, n);
    /* :76-78 sqrt(2.0)*(1.0,0.0) is complex(real32) */
    const double sq2 = (double)sqrtf(2.0f);
    t(1, 1, 1, 1) = cplx(sq2 * ttop, 0.0 * ttop);
    t(1, 1, 2, 1) = cplx(sq2 * ttop, 0.0 * ttop);
    surfs(1, 1) = cplx(sq2 * tref_, 0.0) - gam1 * phis(1, 1);
    for (int k = 3; k <= kx; k++) {
        double f = pow(geo.fsg[k], rgam);
        SLOOP t(m, n, k, 1) = surfs(m, n) * f;
    }
    double rlog0 = (double)logf(1.013f);   /* :87 */
    GLOOP surfg(i, j) = rlog0 + rgamr * log(1.0 - gam2 * phis0(i, j));
    grid_to_spec(surfg.p(), ps.p(1, 1, 1));
    if (ix == iy * 4) trunct(ps.p(1, 1, 1));   /* :96 sequence association: level 1 only */
    double esref = 17.0;
    double qref = refrh1 * (double)0.622f * esref;
    double qexp = hscale / hshum;
    GLOOP surfg(i, j) = qref * exp(qexp * surfg(i, j));
    grid_to_spec(surfg.p(), surfs.p());
    if (ix == iy * 4) trunct(surfs.p());
    for (int k = 3; k <= kx; k++) {
        double f = pow(geo.fsg[k], qexp);
        SLOOP tr(m, n, k, 1, 1) = surfs(m, n) * f;
    }
    double diag[kx * 3];
    check_diagnostics(vor.p(1, 1, 1, 1), div_.p(1, 1, 1, 1), t.p(1, 1, 1, 1), 0, diag, false);
    get_geopotential(t.p(1, 1, 1, 1), phis.p(), phi.p());   /* :123 4-D actual -> time level 1 */
}

}  // namespace orc
