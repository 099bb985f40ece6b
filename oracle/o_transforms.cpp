/* ORACLE (test infrastructure) — restatement of geometry.f90, legendre.f90, fourier.f90,
 * spectral.f90 of the reference.  Each function cites the lines it follows. */
#include "oracle.h"

namespace orc {

Geometry geo;

/* geometry.f90:35-90 */
void initialize_geometry() {
    /* :47 hsg literals are real32 */
    static const float hsg8[9] = {0.000f, 0.050f, 0.140f, 0.260f, 0.420f, 0.600f, 0.770f, 0.900f, 1.000f};
    for (int k = 1; k <= kx + 1; k++) geo.hsg[k] = (double)hsg8[k - 1];
    for (int k = 1; k <= kx; k++) {
        geo.dhs[k] = geo.hsg[k + 1] - geo.hsg[k];
        geo.fsg[k] = 0.5 * (geo.hsg[k + 1] + geo.hsg[k]);
    }
    for (int k = 1; k <= kx; k++) {
        geo.dhsr[k] = 0.5 / geo.dhs[k];
        geo.fsgr[k] = akap / (2. * geo.fsg[k]);
    }
    for (int j = 1; j <= iy; j++) {
        int jj = il + 1 - j;
        /* :68 all-real32 expression incl. cosf */
        float arg = 3.141592654f * ((float)j - 0.25f) / ((float)il + 0.5f);
        geo.sia_half[j] = (double)cosf(arg);
        geo.coa_half[j] = sqrt(1.0 - geo.sia_half[j] * geo.sia_half[j]);
        geo.sia[j] = -geo.sia_half[j];
        geo.sia[jj] = geo.sia_half[j];
        geo.coa[j] = geo.coa_half[j];
        geo.coa[jj] = geo.coa_half[j];
        geo.radang[j] = -asin(geo.sia_half[j]);
        geo.radang[jj] = asin(geo.sia_half[j]);
    }
    for (int j = 1; j <= iy; j++) {
        int jj = il + 1 - j;
        geo.cosg[j] = geo.coa_half[j];
        geo.cosg[jj] = geo.coa_half[j];
        geo.cosgr[j] = 1. / geo.coa_half[j];
        geo.cosgr[jj] = 1. / geo.coa_half[j];
        geo.cosgr2[j] = 1. / (geo.coa_half[j] * geo.coa_half[j]);
        geo.cosgr2[jj] = 1. / (geo.coa_half[j] * geo.coa_half[j]);
    }
    for (int j = 1; j <= il; j++) geo.coriol[j] = 2.0 * omega * geo.sia[j];
}

FA<double, 2 * mx, nx, iy> cpol;
FA<double, mx + 1, nx + 1> epsi, repsi;
int nsh2[nx + 1];
double wt[iy + 1];

/* legendre.f90:158-191 — all real64 */
static void get_weights(double* w /*1-based*/) {
    int n = 2 * iy;
    double z1 = 2.0;
    for (int i = 1; i <= iy; i++) {
        double z = cos(3.141592654 * ((double)i - 0.25) / ((double)n + 0.5));
        double pp = 0.0;
        while (fabs(z - z1) > 2.220446049250313e-16) {
            double p1 = 1.0, p2 = 0.0;
            for (int j = 1; j <= n; j++) {
                double p3 = p2;
                p2 = p1;
                p1 = ((2.0 * (double)j - 1.0) * z * p2 - ((double)j - 1.0) * p3) / j;
            }
            pp = (double)n * (z * p1 - p2) / (z * z - 1.0);
            z1 = z;
            z = z1 - p1 / pp;
        }
        w[i] = 2.0 / ((1.0 - z * z) * (pp * pp));
    }
}

/* legendre.f90:194-237 */
static void get_legendre_poly(int j, FA<double, mx, nx>& poly) {
    const double small = (double)1.e-30f;
    double consq[mx + 1];
    FA<double, mx + 1, nx> alp;
    double y = geo.coa_half[j];
    double x = geo.sia_half[j];
    for (int m = 1; m <= mx; m++)
        consq[m] = (double)sqrtf(0.5f * (2.0f * (float)m + 1.0f) / (float)m); /* :208 real32 */
    alp(1, 1) = (double)sqrtf(0.5f);                                          /* :212 real32 */
    for (int m = 2; m <= mx + 1; m++) alp(m, 1) = consq[m - 1] * y * alp(m - 1, 1);
    for (int m = 1; m <= mx + 1; m++) alp(m, 2) = (x * alp(m, 1)) * repsi(m, 2);
    for (int n = 3; n <= nx; n++)
        for (int m = 1; m <= mx + 1; m++)
            alp(m, n) = (x * alp(m, n - 1) - epsi(m, n - 1) * alp(m, n - 2)) * repsi(m, n);
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx + 1; m++)
            if (fabs(alp(m, n)) <= small) alp(m, n) = 0.0;
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++) poly(m, n) = alp(m, n);
}

/* legendre.f90:23-71 */
void initialize_legendre() {
    get_weights(wt);
    for (int n = 1; n <= nx; n++) {
        nsh2[n] = 0;
        for (int m = 1; m <= mx; m++) {
            int wavenum_tot = (m - 1) + n - 1;
            if (wavenum_tot <= (trunc_ + 1) || ix != 4 * iy) nsh2[n] = nsh2[n] + 2;
        }
    }
    for (int m = 1; m <= mx + 1; m++)
        for (int n = 1; n <= nx + 1; n++) {
            double emm2 = (double)((float)(m - 1) * (float)(m - 1));
            double ell2 = (double)((float)(n + m - 2) * (float)(n + m - 2));
            if (n == nx + 1) epsi(m, n) = 0.0;
            else if (n == 1 && m == 1) epsi(m, n) = 0.0;
            else epsi(m, n) = sqrt((ell2 - emm2) / (4.0 * ell2 - 1.0));
            repsi(m, n) = 0.0;
            if (epsi(m, n) > 0.) repsi(m, n) = 1.0 / epsi(m, n);
        }
    FA<double, mx, nx> poly;
    for (int j = 1; j <= iy; j++) {
        get_legendre_poly(j, poly);
        for (int n = 1; n <= nx; n++)
            for (int m = 1; m <= mx; m++) {
                cpol(2 * m - 1, n, j) = poly(m, n);
                cpol(2 * m, n, j) = poly(m, n);
            }
    }
}

#define IN2(m, n) input[((m)-1) + 2 * mx * ((n)-1)]
#define OUT2(m, j) output[((m)-1) + 2 * mx * ((j)-1)]

/* legendre.f90:74-111 */
void legendre_inv(const double* input, double* output) {
    double even[2 * mx + 1], odd[2 * mx + 1];
    for (int j = 1; j <= iy; j++) {
        int j1 = il + 1 - j;
        for (int m = 1; m <= 2 * mx; m++) { even[m] = 0.0; odd[m] = 0.0; }
        for (int n = 1; n <= nx; n += 2)
            for (int m = 1; m <= nsh2[n]; m++) even[m] = even[m] + IN2(m, n) * cpol(m, n, j);
        for (int n = 2; n <= nx; n += 2)
            for (int m = 1; m <= nsh2[n]; m++) odd[m] = odd[m] + IN2(m, n) * cpol(m, n, j);
        for (int m = 1; m <= 2 * mx; m++) {
            OUT2(m, j1) = even[m] + odd[m];
            OUT2(m, j) = even[m] - odd[m];
        }
    }
}
#undef IN2
#undef OUT2

/* legendre.f90:114-155 */
void legendre_dir(const double* input /*(2mx,il)*/, double* output /*(2mx,nx)*/) {
    FA<double, 2 * mx, iy> even, odd;
    for (int i = 0; i < 2 * mx * nx; i++) output[i] = 0.0;
    for (int j = 1; j <= iy; j++) {
        int j1 = il + 1 - j;
        for (int m = 1; m <= 2 * mx; m++) {
            double a = input[(m - 1) + 2 * mx * (j1 - 1)], b = input[(m - 1) + 2 * mx * (j - 1)];
            even(m, j) = (a + b) * wt[j];
            odd(m, j) = (a - b) * wt[j];
        }
    }
    for (int n = 1; n <= trunc_ + 1; n += 2)
        for (int m = 1; m <= nsh2[n]; m++) {
            double s = 0.0;
            for (int j = 1; j <= iy; j++) s = s + cpol(m, n, j) * even(m, j);
            output[(m - 1) + 2 * mx * (n - 1)] = s;
        }
    for (int n = 2; n <= trunc_ + 1; n += 2)
        for (int m = 1; m <= nsh2[n]; m++) {
            double s = 0.0;
            for (int j = 1; j <= iy; j++) s = s + cpol(m, n, j) * odd(m, j);
            output[(m - 1) + 2 * mx * (n - 1)] = s;
        }
}

double fft_work[ix + 1];
int fft_ifac[16];

/* fourier.f90:17-20 */
void initialize_fourier() { rffti1(ix, fft_work, fft_ifac); }

/* fourier.f90:23-53 */
void fourier_inv(const double* input, int kcos, double* output) {
    double fvar[ix + 1], ch[ix + 1];
    for (int j = 1; j <= il; j++) {
        const double* in = input + (size_t)2 * mx * (j - 1) - 1; /* 1-based */
        fvar[1] = in[1];
        for (int m = 3; m <= 2 * mx; m++) fvar[m - 1] = in[m];
        for (int m = 2 * mx; m <= ix; m++) fvar[m] = 0.0;
        rfftb1(ix, fvar, ch, fft_work, fft_ifac);
        double* out = output + (size_t)ix * (j - 1) - 1;
        if (kcos == 1) {
            for (int i = 1; i <= ix; i++) out[i] = fvar[i];
        } else {
            for (int i = 1; i <= ix; i++) out[i] = fvar[i] * geo.cosgr[j];
        }
    }
}

/* fourier.f90:56-82 */
void fourier_dir(const double* input, double* output) {
    double fvar[ix + 1], ch[ix + 1];
    for (int j = 1; j <= il; j++) {
        const double* in = input + (size_t)ix * (j - 1) - 1;
        for (int i = 1; i <= ix; i++) fvar[i] = in[i];
        rfftf1(ix, fvar, ch, fft_work, fft_ifac);
        double scale = (double)(1.0f / (float)ix); /* :72 real32 quotient */
        double* out = output + (size_t)2 * mx * (j - 1) - 1;
        out[1] = fvar[1] * scale;
        out[2] = 0.0;
        for (int m = 3; m <= 2 * mx; m++) out[m] = fvar[m - 1] * scale;
    }
}

FA<double, mx, nx> el2, elm2, el4, trfilt;
double gradx[mx + 1];
FA<double, mx, nx> gradym, gradyp, uvdx, uvdym, uvdyp, vddym, vddyp;

/* spectral.f90:20-82 */
void initialize_spectral() {
    initialize_fourier();
    initialize_legendre();
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++) {
            int wt_ = (m - 1) + n - 1;
            el2(m, n) = (double)(float)(wt_ * (wt_ + 1)) / (rearth * rearth);
            el4(m, n) = el2(m, n) * el2(m, n);
            trfilt(m, n) = (wt_ <= trunc_) ? 1.0 : 0.0;
        }
    elm2(1, 1) = 0.0;
    for (int n = 1; n <= nx; n++)
        for (int m = 2; m <= mx; m++) elm2(m, n) = 1.0 / el2(m, n);
    for (int n = 2; n <= nx; n++) elm2(1, n) = 1.0 / el2(1, n);
    for (int m = 1; m <= mx; m++)
        for (int n = 1; n <= nx; n++) {
            int m1 = m - 1;
            int m2 = m1 + 1;
            double el1 = (double)(float)((m - 1) + n - 1);
            if (n == 1) {
                gradx[m] = (double)(float)m1 / rearth;
                uvdx(m, 1) = -rearth / (double)(float)(m1 + 1);
                uvdym(m, 1) = 0.0;
                vddym(m, 1) = 0.0;
            } else {
                uvdx(m, n) = -rearth * (double)(float)m1 / (el1 * (el1 + 1));
                gradym(m, n) = (el1 - 1.0) * epsi(m2, n) / rearth;
                uvdym(m, n) = -rearth * epsi(m2, n) / el1;
                vddym(m, n) = (el1 + 1) * epsi(m2, n) / rearth;
            }
            gradyp(m, n) = (el1 + 2.0) * epsi(m2, n + 1) / rearth;
            uvdyp(m, n) = -rearth * epsi(m2, n + 1) / (el1 + 1.0);
            vddyp(m, n) = el1 * epsi(m2, n + 1) / rearth;
        }
}

#define S(a, m, n) a[((m)-1) + mx * ((n)-1)]

/* spectral.f90:84-96 */
void laplacian(const cplx* in, cplx* out) {
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++) S(out, m, n) = -S(in, m, n) * el2(m, n);
}
void inverse_laplacian(const cplx* in, cplx* out) {
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++) S(out, m, n) = -S(in, m, n) * elm2(m, n);
}

/* spectral.f90:98-122 */
void spec_to_grid(const cplx* vorm, int kcos, double* vorg) {
    std::vector<double> four((size_t)2 * mx * il);
    legendre_inv(reinterpret_cast<const double*>(vorm), four.data());
    fourier_inv(four.data(), kcos, vorg);
}
void grid_to_spec(const double* vorg, cplx* vorm) {
    std::vector<double> four((size_t)2 * mx * il);
    fourier_dir(vorg, four.data());
    legendre_dir(four.data(), reinterpret_cast<double*>(vorm));
}

/* x*(0,1) for complex x */
static inline cplx times_i(cplx x) { return cplx(-x.imag(), x.real()); }

/* spectral.f90:124-144 */
void grad(const cplx* psi, cplx* psdx, cplx* psdy) {
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++) S(psdx, m, n) = times_i(gradx[m] * S(psi, m, n));
    for (int m = 1; m <= mx; m++) {
        S(psdy, m, 1) = gradyp(m, 1) * S(psi, m, 2);
        S(psdy, m, nx) = -gradym(m, nx) * S(psi, m, trunc_ + 1);
    }
    for (int n = 2; n <= trunc_ + 1; n++)
        for (int m = 1; m <= mx; m++)
            S(psdy, m, n) = -gradym(m, n) * S(psi, m, n - 1) + gradyp(m, n) * S(psi, m, n + 1);
}

/* spectral.f90:146-171 */
void vds(const cplx* ucosm, const cplx* vcosm, cplx* vorm, cplx* divm) {
    std::vector<cplx> zc((size_t)mx * nx), zp((size_t)mx * nx);
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++) {
            S(zp, m, n) = times_i(gradx[m] * S(ucosm, m, n));
            S(zc, m, n) = times_i(gradx[m] * S(vcosm, m, n));
        }
    for (int m = 1; m <= mx; m++) {
        S(vorm, m, 1) = S(zc, m, 1) - vddyp(m, 1) * S(ucosm, m, 2);
        S(vorm, m, nx) = vddym(m, nx) * S(ucosm, m, trunc_ + 1);
        S(divm, m, 1) = S(zp, m, 1) + vddyp(m, 1) * S(vcosm, m, 2);
        S(divm, m, nx) = -vddym(m, nx) * S(vcosm, m, trunc_ + 1);
    }
    for (int n = 2; n <= trunc_ + 1; n++)
        for (int m = 1; m <= mx; m++) {
            S(vorm, m, n) = vddym(m, n) * S(ucosm, m, n - 1) - vddyp(m, n) * S(ucosm, m, n + 1) + S(zc, m, n);
            S(divm, m, n) = -vddym(m, n) * S(vcosm, m, n - 1) + vddyp(m, n) * S(vcosm, m, n + 1) + S(zp, m, n);
        }
}

/* spectral.f90:173-196 */
void uvspec(const cplx* vorm, const cplx* divm, cplx* ucosm, cplx* vcosm) {
    std::vector<cplx> zc((size_t)mx * nx), zp((size_t)mx * nx);
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++) {
            S(zp, m, n) = times_i(uvdx(m, n) * S(vorm, m, n));
            S(zc, m, n) = times_i(uvdx(m, n) * S(divm, m, n));
        }
    for (int m = 1; m <= mx; m++) {
        S(ucosm, m, 1) = S(zc, m, 1) - uvdyp(m, 1) * S(vorm, m, 2);
        S(ucosm, m, nx) = uvdym(m, nx) * S(vorm, m, trunc_ + 1);
        S(vcosm, m, 1) = S(zp, m, 1) + uvdyp(m, 1) * S(divm, m, 2);
        S(vcosm, m, nx) = -uvdym(m, nx) * S(divm, m, trunc_ + 1);
    }
    for (int n = 2; n <= trunc_ + 1; n++)
        for (int m = 1; m <= mx; m++) {
            S(vcosm, m, n) = -uvdym(m, n) * S(divm, m, n - 1) + uvdyp(m, n) * S(divm, m, n + 1) + S(zp, m, n);
            S(ucosm, m, n) = uvdym(m, n) * S(vorm, m, n - 1) - uvdyp(m, n) * S(vorm, m, n + 1) + S(zc, m, n);
        }
}

/* spectral.f90:198-227 */
void vdspec(const double* ug, const double* vg, cplx* vorm, cplx* divm, int kcos) {
    std::vector<double> ug1((size_t)ix * il), vg1((size_t)ix * il);
    std::vector<cplx> specu((size_t)mx * nx), specv((size_t)mx * nx);
    for (int j = 1; j <= il; j++)
        for (int i = 1; i <= ix; i++) {
            size_t q = (size_t)(i - 1) + (size_t)ix * (j - 1);
            double c = (kcos == 2) ? geo.cosgr[j] : geo.cosgr2[j];
            ug1[q] = ug[q] * c;
            vg1[q] = vg[q] * c;
        }
    grid_to_spec(ug1.data(), specu.data());
    grid_to_spec(vg1.data(), specv.data());
    vds(specu.data(), specv.data(), vorm, divm);
}

/* spectral.f90:229-233 */
void trunct(cplx* vor) {
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++) S(vor, m, n) = S(vor, m, n) * trfilt(m, n);
}
#undef S

}  // namespace orc
