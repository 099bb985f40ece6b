// Product-side host table builders (see tables.h).  Formulas follow the reference's
// start-up routines; citations are file:line under the reference's source/.
#include "tables.h"
#include <cmath>
#include <cstring>
#include <stdexcept>

namespace spd {

Consts make_consts() {
    Consts c;
    c.rearth = 6.371e+6;                 // physical_constants.f90:16
    c.omega = (double)7.292e-05f;        // :17 (real32 literal)
    c.grav = (double)9.81f;              // :18
    c.p0 = 1.e+5;
    c.cp = 1004.0;
    c.akap = (double)(2.0f / 7.0f);      // :23 real32 quotient
    c.rgas = c.akap * c.cp;              // :25
    c.alhc = 2501.0;
    c.alhs = 2801.0;
    c.sbc = (double)5.67e-8f;
    c.gamma = 6.0; c.hscale = 7.5; c.hshum = 2.5;      // dynamical_constants.f90:12-14
    c.refrh1 = (double)0.7f; c.thd = (double)2.4f; c.thdd = (double)2.4f;
    c.thds = 12.0; c.tdrs = 24.0 * 30.0;
    c.rob = (double)0.05f; c.wil = (double)0.53f; c.alph = 0.5;  // params.f90:32-34
    c.nsteps = 36; c.delt = 2400.0; c.nstrad = 3;
    return c;
}

// ------------------------------------------------------------------------------------
// Real FFT with the FFTPACK pass structure (fftpack.f90).  The product only needs it at
// start-up: it pushes unit vectors through it to obtain the dense forward/backward
// operators that the CUDA kernels apply as FP64 tensor-core GEMMs.
// ------------------------------------------------------------------------------------
static void factorize(int n, std::vector<int>& fac) {   // fftpack.f90:13-36
    static const int tries[4] = {4, 2, 3, 5};
    fac.clear();
    int nl = n, t = 0, idx = 0;
    while (nl != 1) {
        t = (idx < 4) ? tries[idx] : t + 2;
        idx++;
        while (nl % t == 0) {
            nl /= t;
            if (t == 2 && !fac.empty()) fac.insert(fac.begin(), 2); else fac.push_back(t);
        }
    }
}

static void make_twiddles(int n, const std::vector<int>& fac, std::vector<double>& wa) {  // :39-66
    wa.assign(n, 0.0);
    const double tpi = (double)(8.f * atanf(1.f));
    const double argh = tpi / n;
    int is = 0, l1 = 1;
    for (size_t f = 0; f + 1 < fac.size(); f++) {
        int ip = fac[f], l2 = l1 * ip, ido = n / l2;
        int ld = 0;
        for (int j = 1; j < ip; j++) {
            ld += l1;
            double argld = ld * argh;
            double fi = 0.;
            int i = is;
            for (int ii = 3; ii <= ido; ii += 2) {
                fi += 1.;
                double arg = fi * argld;
                wa[i] = cos(arg);
                wa[i + 1] = sin(arg);
                i += 2;
            }
            is += ido;
        }
        l1 = l2;
    }
}

namespace {
struct View3 {   // Fortran-order 3-D view, 1-based accessors
    double* p; int n1, n2;
    inline double& operator()(int i, int j, int k) const { return p[(i - 1) + n1 * ((j - 1) + n2 * (k - 1))]; }
};
const double kTaui = (double)(.5f * sqrtf(3.f));     // fftpack.f90:269,787
const double kSqrt2 = (double)sqrtf(2.f);            // :341
const double kHsqt2 = (double)(.5f * sqrtf(2.f));    // :857
}

// backward pass of radix ip: cc(ido,ip,l1) -> ch(ido,l1,ip)   (radb2/3/4 :204-424)
static void pass_backward(int ip, int ido, int l1, double* ccp, double* chp, const double* w) {
    View3 cc{ccp, ido, ip}, ch{chp, ido, l1};
    const double* w1 = w - 1;             // w1[i-2], w1[i-1] with i>=3
    const double* w2 = w + ido - 1;
    const double* w3 = w + 2 * ido - 1;
    for (int k = 1; k <= l1; k++) {
        if (ip == 2) {
            ch(1, k, 1) = cc(1, 1, k) + cc(ido, 2, k);
            ch(1, k, 2) = cc(1, 1, k) - cc(ido, 2, k);
        } else if (ip == 3) {
            double tr2 = cc(ido, 2, k) + cc(ido, 2, k);
            double cr2 = cc(1, 1, k) + (-.5) * tr2;
            ch(1, k, 1) = cc(1, 1, k) + tr2;
            double ci3 = kTaui * (cc(1, 3, k) + cc(1, 3, k));
            ch(1, k, 2) = cr2 - ci3;
            ch(1, k, 3) = cr2 + ci3;
        } else {
            double tr1 = cc(1, 1, k) - cc(ido, 4, k), tr2 = cc(1, 1, k) + cc(ido, 4, k);
            double tr3 = cc(ido, 2, k) + cc(ido, 2, k), tr4 = cc(1, 3, k) + cc(1, 3, k);
            ch(1, k, 1) = tr2 + tr3; ch(1, k, 2) = tr1 - tr4;
            ch(1, k, 3) = tr2 - tr3; ch(1, k, 4) = tr1 + tr4;
        }
        for (int i = 3; i <= ido; i += 2) {
            int ic = ido + 2 - i;
            if (ip == 2) {
                ch(i - 1, k, 1) = cc(i - 1, 1, k) + cc(ic - 1, 2, k);
                double tr2 = cc(i - 1, 1, k) - cc(ic - 1, 2, k);
                ch(i, k, 1) = cc(i, 1, k) - cc(ic, 2, k);
                double ti2 = cc(i, 1, k) + cc(ic, 2, k);
                ch(i - 1, k, 2) = w1[i - 2] * tr2 - w1[i - 1] * ti2;
                ch(i, k, 2) = w1[i - 2] * ti2 + w1[i - 1] * tr2;
            } else if (ip == 3) {
                double tr2 = cc(i - 1, 3, k) + cc(ic - 1, 2, k);
                double cr2 = cc(i - 1, 1, k) + (-.5) * tr2;
                ch(i - 1, k, 1) = cc(i - 1, 1, k) + tr2;
                double ti2 = cc(i, 3, k) - cc(ic, 2, k);
                double ci2 = cc(i, 1, k) + (-.5) * ti2;
                ch(i, k, 1) = cc(i, 1, k) + ti2;
                double cr3 = kTaui * (cc(i - 1, 3, k) - cc(ic - 1, 2, k));
                double ci3 = kTaui * (cc(i, 3, k) + cc(ic, 2, k));
                double dr2 = cr2 - ci3, dr3 = cr2 + ci3, di2 = ci2 + cr3, di3 = ci2 - cr3;
                ch(i - 1, k, 2) = w1[i - 2] * dr2 - w1[i - 1] * di2;
                ch(i, k, 2) = w1[i - 2] * di2 + w1[i - 1] * dr2;
                ch(i - 1, k, 3) = w2[i - 2] * dr3 - w2[i - 1] * di3;
                ch(i, k, 3) = w2[i - 2] * di3 + w2[i - 1] * dr3;
            } else {
                double ti1 = cc(i, 1, k) + cc(ic, 4, k), ti2 = cc(i, 1, k) - cc(ic, 4, k);
                double ti3 = cc(i, 3, k) - cc(ic, 2, k), tr4 = cc(i, 3, k) + cc(ic, 2, k);
                double tr1 = cc(i - 1, 1, k) - cc(ic - 1, 4, k), tr2 = cc(i - 1, 1, k) + cc(ic - 1, 4, k);
                double ti4 = cc(i - 1, 3, k) - cc(ic - 1, 2, k), tr3 = cc(i - 1, 3, k) + cc(ic - 1, 2, k);
                ch(i - 1, k, 1) = tr2 + tr3;
                double cr3 = tr2 - tr3;
                ch(i, k, 1) = ti2 + ti3;
                double ci3 = ti2 - ti3;
                double cr2 = tr1 - tr4, cr4 = tr1 + tr4, ci2 = ti1 + ti4, ci4 = ti1 - ti4;
                ch(i - 1, k, 2) = w1[i - 2] * cr2 - w1[i - 1] * ci2;
                ch(i, k, 2) = w1[i - 2] * ci2 + w1[i - 1] * cr2;
                ch(i - 1, k, 3) = w2[i - 2] * cr3 - w2[i - 1] * ci3;
                ch(i, k, 3) = w2[i - 2] * ci3 + w2[i - 1] * cr3;
                ch(i - 1, k, 4) = w3[i - 2] * cr4 - w3[i - 1] * ci4;
                ch(i, k, 4) = w3[i - 2] * ci4 + w3[i - 1] * cr4;
            }
        }
        if (ido % 2 == 0) {   // middle element (ip==3 never has even ido here: handled like fftpack, no tail)
            if (ip == 2) {
                ch(ido, k, 1) = cc(ido, 1, k) + cc(ido, 1, k);
                ch(ido, k, 2) = -(cc(1, 2, k) + cc(1, 2, k));
            } else if (ip == 4) {
                double ti1 = cc(1, 2, k) + cc(1, 4, k), ti2 = cc(1, 4, k) - cc(1, 2, k);
                double tr1 = cc(ido, 1, k) - cc(ido, 3, k), tr2 = cc(ido, 1, k) + cc(ido, 3, k);
                ch(ido, k, 1) = tr2 + tr2;
                ch(ido, k, 2) = kSqrt2 * (tr1 - ti1);
                ch(ido, k, 3) = ti2 + ti2;
                ch(ido, k, 4) = -kSqrt2 * (tr1 + ti1);
            }
        }
    }
}

// forward pass of radix ip: cc(ido,l1,ip) -> ch(ido,ip,l1)   (radf2/3/4 :722-936)
static void pass_forward(int ip, int ido, int l1, double* ccp, double* chp, const double* w) {
    View3 cc{ccp, ido, l1}, ch{chp, ido, ip};
    const double* w1 = w - 1;
    const double* w2 = w + ido - 1;
    const double* w3 = w + 2 * ido - 1;
    for (int k = 1; k <= l1; k++) {
        if (ip == 2) {
            ch(1, 1, k) = cc(1, k, 1) + cc(1, k, 2);
            ch(ido, 2, k) = cc(1, k, 1) - cc(1, k, 2);
        } else if (ip == 3) {
            double cr2 = cc(1, k, 2) + cc(1, k, 3);
            ch(1, 1, k) = cc(1, k, 1) + cr2;
            ch(1, 3, k) = kTaui * (cc(1, k, 3) - cc(1, k, 2));
            ch(ido, 2, k) = cc(1, k, 1) + (-.5) * cr2;
        } else {
            double tr1 = cc(1, k, 2) + cc(1, k, 4), tr2 = cc(1, k, 1) + cc(1, k, 3);
            ch(1, 1, k) = tr1 + tr2;
            ch(ido, 4, k) = tr2 - tr1;
            ch(ido, 2, k) = cc(1, k, 1) - cc(1, k, 3);
            ch(1, 3, k) = cc(1, k, 4) - cc(1, k, 2);
        }
        for (int i = 3; i <= ido; i += 2) {
            int ic = ido + 2 - i;
            if (ip == 2) {
                double tr2 = w1[i - 2] * cc(i - 1, k, 2) + w1[i - 1] * cc(i, k, 2);
                double ti2 = w1[i - 2] * cc(i, k, 2) - w1[i - 1] * cc(i - 1, k, 2);
                ch(i, 1, k) = cc(i, k, 1) + ti2;
                ch(ic, 2, k) = ti2 - cc(i, k, 1);
                ch(i - 1, 1, k) = cc(i - 1, k, 1) + tr2;
                ch(ic - 1, 2, k) = cc(i - 1, k, 1) - tr2;
            } else if (ip == 3) {
                double dr2 = w1[i - 2] * cc(i - 1, k, 2) + w1[i - 1] * cc(i, k, 2);
                double di2 = w1[i - 2] * cc(i, k, 2) - w1[i - 1] * cc(i - 1, k, 2);
                double dr3 = w2[i - 2] * cc(i - 1, k, 3) + w2[i - 1] * cc(i, k, 3);
                double di3 = w2[i - 2] * cc(i, k, 3) - w2[i - 1] * cc(i - 1, k, 3);
                double cr2 = dr2 + dr3, ci2 = di2 + di3;
                ch(i - 1, 1, k) = cc(i - 1, k, 1) + cr2;
                ch(i, 1, k) = cc(i, k, 1) + ci2;
                double tr2 = cc(i - 1, k, 1) + (-.5) * cr2, ti2 = cc(i, k, 1) + (-.5) * ci2;
                double tr3 = kTaui * (di2 - di3), ti3 = kTaui * (dr3 - dr2);
                ch(i - 1, 3, k) = tr2 + tr3;
                ch(ic - 1, 2, k) = tr2 - tr3;
                ch(i, 3, k) = ti2 + ti3;
                ch(ic, 2, k) = ti3 - ti2;
            } else {
                double cr2 = w1[i - 2] * cc(i - 1, k, 2) + w1[i - 1] * cc(i, k, 2);
                double ci2 = w1[i - 2] * cc(i, k, 2) - w1[i - 1] * cc(i - 1, k, 2);
                double cr3 = w2[i - 2] * cc(i - 1, k, 3) + w2[i - 1] * cc(i, k, 3);
                double ci3 = w2[i - 2] * cc(i, k, 3) - w2[i - 1] * cc(i - 1, k, 3);
                double cr4 = w3[i - 2] * cc(i - 1, k, 4) + w3[i - 1] * cc(i, k, 4);
                double ci4 = w3[i - 2] * cc(i, k, 4) - w3[i - 1] * cc(i - 1, k, 4);
                double tr1 = cr2 + cr4, tr4 = cr4 - cr2, ti1 = ci2 + ci4, ti4 = ci2 - ci4;
                double ti2 = cc(i, k, 1) + ci3, ti3 = cc(i, k, 1) - ci3;
                double tr2 = cc(i - 1, k, 1) + cr3, tr3 = cc(i - 1, k, 1) - cr3;
                ch(i - 1, 1, k) = tr1 + tr2;
                ch(ic - 1, 4, k) = tr2 - tr1;
                ch(i, 1, k) = ti1 + ti2;
                ch(ic, 4, k) = ti1 - ti2;
                ch(i - 1, 3, k) = ti4 + tr3;
                ch(ic - 1, 2, k) = tr3 - ti4;
                ch(i, 3, k) = tr4 + ti3;
                ch(ic, 2, k) = tr4 - ti3;
            }
        }
        if (ido % 2 == 0) {
            if (ip == 2) {
                ch(1, 2, k) = -cc(ido, k, 2);
                ch(ido, 1, k) = cc(ido, k, 1);
            } else if (ip == 4) {
                double ti1 = -kHsqt2 * (cc(ido, k, 2) + cc(ido, k, 4));
                double tr1 = kHsqt2 * (cc(ido, k, 2) - cc(ido, k, 4));
                ch(ido, 1, k) = tr1 + cc(ido, k, 1);
                ch(ido, 3, k) = cc(ido, k, 1) - tr1;
                ch(1, 2, k) = ti1 - cc(ido, k, 3);
                ch(1, 4, k) = ti1 + cc(ido, k, 3);
            }
        }
    }
}

void rfft_backward(const Tables& t, double* x) {   // rfftb1 fftpack.f90:69-134
    const int n = t.d.ix;
    std::vector<double> buf(n);
    double* a = x; double* b = buf.data();
    int l1 = 1, iw = 0;
    for (size_t f = 0; f < t.fft_fac.size(); f++) {
        int ip = t.fft_fac[f], ido = n / (l1 * ip);
        if (ip > 4) throw std::runtime_error("rfft: radix > 4 not supported");
        pass_backward(ip, ido, l1, a, b, t.fft_work.data() + iw);
        std::swap(a, b);
        l1 *= ip;
        iw += (ip - 1) * ido;
    }
    if (a != x) memcpy(x, a, sizeof(double) * n);
}

void rfft_forward(const Tables& t, double* x) {    // rfftf1 fftpack.f90:136-202
    const int n = t.d.ix;
    std::vector<double> buf(n);
    double* a = x; double* b = buf.data();
    int l2 = n, iw = n - 1;
    for (int f = (int)t.fft_fac.size() - 1; f >= 0; f--) {
        int ip = t.fft_fac[f], l1 = l2 / ip, ido = n / l2;
        if (ip > 4) throw std::runtime_error("rfft: radix > 4 not supported");
        iw -= (ip - 1) * ido;
        pass_forward(ip, ido, l1, a, b, t.fft_work.data() + iw);
        std::swap(a, b);
        l2 = l1;
    }
    if (a != x) memcpy(x, a, sizeof(double) * n);
}

// Dense operators.  Backward (fourier_inv, fourier.f90:33-44): half-complex unpack drops
// Im(m=0) and zero-pads m >= mx.  Forward (fourier_dir :65-81): scale by real32 1/ix, keep
// m < mx, Im(m=0) = 0.
static void build_dft(Tables& t) {
    const int ix = t.d.ix, k2 = t.d.k2(), kp = t.d.k2pad();
    t.finv.assign((size_t)ix * kp, 0.0);
    t.ffwd.assign((size_t)kp * ix, 0.0);
    std::vector<double> x(ix);
    for (int c = 0; c < k2; c++) {          // c = Fortran row index-1 of the (2*mx) Fourier array
        if (c == 1) continue;               // Im(m=0) never enters the backward FFT
        std::fill(x.begin(), x.end(), 0.0);
        x[c == 0 ? 0 : c - 1] = 1.0;        // fvar(1)=in(1); fvar(m-1)=in(m), m>=3
        rfft_backward(t, x.data());
        for (int i = 0; i < ix; i++) t.finv[(size_t)i * kp + c] = x[i];
    }
    const double scale = (double)(1.0f / (float)ix);   // fourier.f90:72
    for (int i = 0; i < ix; i++) {
        std::fill(x.begin(), x.end(), 0.0);
        x[i] = 1.0;
        rfft_forward(t, x.data());
        t.ffwd[(size_t)0 * ix + i] = x[0] * scale;
        for (int c = 2; c < k2; c++) t.ffwd[(size_t)c * ix + i] = x[c - 1] * scale;
    }
}

// ------------------------------------------------------------------------------------
static void build_geometry(Tables& t) {   // geometry.f90:35-90
    const int kx = t.d.kx, il = t.d.il, iy = t.d.iy;
    static const float hsg8[9] = {0.000f, 0.050f, 0.140f, 0.260f, 0.420f, 0.600f, 0.770f, 0.900f, 1.000f};
    t.hsg.resize(kx + 1); t.dhs.resize(kx); t.fsg.resize(kx); t.dhsr.resize(kx); t.fsgr.resize(kx);
    for (int k = 0; k <= kx; k++) t.hsg[k] = (double)hsg8[k];
    for (int k = 0; k < kx; k++) {
        t.dhs[k] = t.hsg[k + 1] - t.hsg[k];
        t.fsg[k] = 0.5 * (t.hsg[k + 1] + t.hsg[k]);
        t.dhsr[k] = 0.5 / t.dhs[k];
        t.fsgr[k] = t.c.akap / (2. * t.fsg[k]);
    }
    for (auto* v : {&t.radang, &t.coriol, &t.sia, &t.coa, &t.cosg, &t.cosgr, &t.cosgr2}) v->resize(il);
    t.sia_half.resize(iy); t.coa_half.resize(iy);
    for (int j = 0; j < iy; j++) {
        int jj = il - 1 - j;
        float arg = 3.141592654f * ((float)(j + 1) - 0.25f) / ((float)il + 0.5f);   // :68 real32
        double s = (double)cosf(arg);
        double c = sqrt(1.0 - s * s);
        t.sia_half[j] = s; t.coa_half[j] = c;
        t.sia[j] = -s; t.sia[jj] = s;
        t.coa[j] = c; t.coa[jj] = c;
        t.radang[j] = -asin(s); t.radang[jj] = asin(s);
        t.cosg[j] = t.cosg[jj] = c;
        t.cosgr[j] = t.cosgr[jj] = 1. / c;
        t.cosgr2[j] = t.cosgr2[jj] = 1. / (c * c);
    }
    for (int j = 0; j < il; j++) t.coriol[j] = 2.0 * t.c.omega * t.sia[j];
}

static void build_legendre(Tables& t) {   // legendre.f90:23-71,158-237
    const int mx = t.d.mx, nx = t.d.nx, iy = t.d.iy, trunc = t.d.trunc;
    // Gaussian weights (Newton iteration on P_n, all real64) :158-191
    t.wt.resize(iy);
    {
        const int n = 2 * iy;
        double z1 = 2.0;
        for (int i = 1; i <= iy; i++) {
            double z = cos(3.141592654 * ((double)i - 0.25) / ((double)n + 0.5)), pp = 0.0;
            while (fabs(z - z1) > 2.220446049250313e-16) {
                double p1 = 1.0, p2 = 0.0;
                for (int j = 1; j <= n; j++) {
                    double p3 = p2; p2 = p1;
                    p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
                }
                pp = n * (z * p1 - p2) / (z * z - 1.0);
                z1 = z;
                z = z1 - p1 / pp;
            }
            t.wt[i - 1] = 2.0 / ((1.0 - z * z) * (pp * pp));
        }
    }
    t.nsh2.assign(nx, 0);
    for (int n = 0; n < nx; n++)
        for (int m = 0; m < mx; m++)
            if (m + n <= trunc + 1) t.nsh2[n] += 2;
    const int me = mx + 1, ne = nx + 1;
    t.epsi.assign((size_t)me * ne, 0.0); t.repsi.assign((size_t)me * ne, 0.0);
    auto E = [&](int m, int n) -> double& { return t.epsi[m + (size_t)me * n]; };
    auto R = [&](int m, int n) -> double& { return t.repsi[m + (size_t)me * n]; };
    for (int m = 0; m < me; m++)
        for (int n = 0; n < ne; n++) {
            double emm2 = (double)((float)m * (float)m);
            double ell2 = (double)((float)(n + m) * (float)(n + m));
            if (n == nx || (n == 0 && m == 0)) E(m, n) = 0.0;
            else E(m, n) = sqrt((ell2 - emm2) / (4.0 * ell2 - 1.0));
            R(m, n) = (E(m, n) > 0.) ? 1.0 / E(m, n) : 0.0;
        }
    // P_n^m at the (approximate) latitudes :194-237
    t.poly.assign((size_t)iy * nx * mx, 0.0);
    t.cpol.assign((size_t)2 * mx * nx * iy, 0.0);
    std::vector<double> alp((size_t)me * nx), consq(mx + 1);
    auto A = [&](int m, int n) -> double& { return alp[m + (size_t)me * n]; };
    for (int m = 1; m <= mx; m++) consq[m] = (double)sqrtf(0.5f * (2.0f * (float)m + 1.0f) / (float)m);
    const double small = (double)1.e-30f;
    for (int j = 0; j < iy; j++) {
        double y = t.coa_half[j], x = t.sia_half[j];
        A(0, 0) = (double)sqrtf(0.5f);
        for (int m = 1; m < me; m++) A(m, 0) = consq[m] * y * A(m - 1, 0);
        for (int m = 0; m < me; m++) A(m, 1) = (x * A(m, 0)) * R(m, 1);
        for (int n = 2; n < nx; n++)
            for (int m = 0; m < me; m++) A(m, n) = (x * A(m, n - 1) - E(m, n - 1) * A(m, n - 2)) * R(m, n);
        for (int n = 0; n < nx; n++)
            for (int m = 0; m < mx; m++) {
                double v = A(m, n);
                if (fabs(v) <= small) v = 0.0;
                t.poly[((size_t)j * nx + n) * mx + m] = v;
                t.cpol[(2 * m) + (size_t)2 * mx * (n + (size_t)nx * j)] = v;
                t.cpol[(2 * m + 1) + (size_t)2 * mx * (n + (size_t)nx * j)] = v;
            }
    }
}

static void build_spectral(Tables& t) {   // spectral.f90:20-82
    const int mx = t.d.mx, nx = t.d.nx, trunc = t.d.trunc, me = mx + 1;
    const double a = t.c.rearth;
    size_t N = (size_t)mx * nx;
    for (auto* v : {&t.el2, &t.elm2, &t.el4, &t.trfilt, &t.gradym, &t.gradyp, &t.uvdx, &t.uvdym, &t.uvdyp, &t.vddym, &t.vddyp})
        v->assign(N, 0.0);
    t.gradx.assign(mx, 0.0);
    auto I = [&](int m, int n) { return m + (size_t)mx * n; };
    auto E = [&](int m, int n) { return t.epsi[m + (size_t)me * n]; };
    for (int n = 0; n < nx; n++)
        for (int m = 0; m < mx; m++) {
            int l = m + n;
            t.el2[I(m, n)] = (double)(float)(l * (l + 1)) / (a * a);
            t.el4[I(m, n)] = t.el2[I(m, n)] * t.el2[I(m, n)];
            t.trfilt[I(m, n)] = (l <= trunc) ? 1.0 : 0.0;
            t.elm2[I(m, n)] = (l == 0) ? 0.0 : 1.0 / t.el2[I(m, n)];
        }
    for (int m = 0; m < mx; m++)
        for (int n = 0; n < nx; n++) {
            double el1 = (double)(float)(m + n);
            int m2 = m + 1;     // Fortran index m2 = m1+1 -> 0-based epsi row m
            (void)m2;
            if (n == 0) {
                t.gradx[m] = (double)(float)m / a;
                t.uvdx[I(m, 0)] = -a / (double)(float)(m + 1);
            } else {
                t.uvdx[I(m, n)] = -a * (double)(float)m / (el1 * (el1 + 1));
                t.gradym[I(m, n)] = (el1 - 1.0) * E(m, n) / a;
                t.uvdym[I(m, n)] = -a * E(m, n) / el1;
                t.vddym[I(m, n)] = (el1 + 1) * E(m, n) / a;
            }
            t.gradyp[I(m, n)] = (el1 + 2.0) * E(m, n + 1) / a;
            t.uvdyp[I(m, n)] = -a * E(m, n + 1) / (el1 + 1.0);
            t.vddyp[I(m, n)] = el1 * E(m, n + 1) / a;
        }
}

static void build_dynamics_tables(Tables& t) {
    const int kx = t.d.kx, mx = t.d.mx, nx = t.d.nx, trunc = t.d.trunc;
    const Consts& c = t.c;
    // geopotential.f90:19-29, :51-56
    t.xgeop1.assign(kx, 0.0); t.xgeop2.assign(kx, 0.0); t.geop_corf.assign(kx, 0.0);
    for (int k = 0; k < kx; k++) {
        t.xgeop1[k] = c.rgas * log(t.hsg[k + 1] / t.fsg[k]);
        if (k != kx - 1) t.xgeop2[k + 1] = c.rgas * log(t.fsg[k + 1] / t.hsg[k + 1]);
    }
    for (int k = 1; k < kx - 1; k++)
        t.geop_corf[k] = t.xgeop1[k] * 0.5 * log(t.hsg[k + 1] / t.fsg[k]) / log(t.fsg[k + 1] / t.fsg[k - 1]);
    // horizontal_diffusion.f90:36-82
    size_t N = (size_t)mx * nx;
    t.dmp.assign(N, 0.0); t.dmpd.assign(N, 0.0); t.dmps.assign(N, 0.0);
    double hdiff = 1. / (c.thd * 3600.), hdifd = 1. / (c.thdd * 3600.), hdifs = 1. / (c.thds * 3600.);
    double rlap = (double)(1.f / (float)(trunc * (trunc + 1)));
    for (int n = 0; n < nx; n++)
        for (int m = 0; m < mx; m++) {
            double twn = (double)(float)(m + n);
            double elap = twn * (twn + 1.) * rlap;
            double e2 = elap * elap;
            double elapn = e2 * e2;                      // elap**4 (integer power)
            t.dmp[m + (size_t)mx * n] = hdiff * elapn;
            t.dmpd[m + (size_t)mx * n] = hdifd * elapn;
            t.dmps[m + (size_t)mx * n] = hdifs * elap;
        }
    double rgam = c.rgas * c.gamma / (1000. * c.grav);
    double qexp = c.hscale / c.hshum;
    t.tcorv.assign(kx, 0.0); t.qcorv.assign(kx, 0.0);
    for (int k = 1; k < kx; k++) {
        t.tcorv[k] = pow(t.fsg[k], rgam);
        if (k > 1) t.qcorv[k] = pow(t.fsg[k], qexp);
    }
    // physics.f90:12-39
    t.sigl.assign(kx, 0.0); t.sigh.assign(kx + 1, 0.0); t.grdsig.assign(kx, 0.0); t.grdscp.assign(kx, 0.0);
    t.wvi.assign((size_t)kx * 2, 0.0);
    t.sigh[0] = t.hsg[0];
    for (int k = 0; k < kx; k++) {
        t.sigl[k] = log(t.fsg[k]);
        t.sigh[k + 1] = t.hsg[k + 1];
        t.grdsig[k] = c.grav / (t.dhs[k] * c.p0);
        t.grdscp[k] = t.grdsig[k] / c.cp;
    }
    for (int k = 0; k < kx - 1; k++) {
        t.wvi[k] = 1. / (t.sigl[k + 1] - t.sigl[k]);
        t.wvi[k + kx] = (log(t.sigh[k + 1]) - t.sigl[k]) * t.wvi[k];
    }
    t.wvi[kx - 1] = 0.;
    t.wvi[kx - 1 + kx] = ((double)logf(0.99f) - t.sigl[kx - 1]) * t.wvi[kx - 2];
    // longwave_radiation.f90:197-220 radset, epslw = 0.05 (mod_radcon.f90:26)
    t.fband.assign((size_t)301 * 4, 0.0);
    auto FB = [&](int jt, int jb) -> double& { return t.fband[(jt - 100) + (size_t)301 * (jb - 1)]; };
    double eps1 = 1.0 - (double)0.05f;
    for (int jt = 200; jt <= 320; jt++) {
        float d2 = (float)((jt - 247) * (jt - 247)), d3 = (float)((jt - 282) * (jt - 282)), d4 = (float)((jt - 315) * (jt - 315));
        FB(jt, 2) = (double)(0.148f - 3.0e-6f * d2) * eps1;
        FB(jt, 3) = (double)(0.356f - 5.2e-6f * d3) * eps1;
        FB(jt, 4) = (double)(0.314f + 1.0e-5f * d4) * eps1;
        FB(jt, 1) = eps1 - (FB(jt, 2) + FB(jt, 3) + FB(jt, 4));
    }
    for (int jb = 1; jb <= 4; jb++) {
        for (int jt = 100; jt <= 199; jt++) FB(jt, jb) = FB(200, jb);
        for (int jt = 321; jt <= 400; jt++) FB(jt, jb) = FB(320, jb);
    }
}

// LU inverse with implicit-scaling partial pivoting (matrix_inversion.f90:12-133);
// Fortran order, a is overwritten.
void invert_matrix(double* a, double* y, int n) {
    auto A = [&](int i, int j) -> double& { return a[i + (size_t)n * j]; };
    std::vector<double> vv(n);
    std::vector<int> indx(n);
    for (int i = 0; i < n; i++) {
        double big = 0.;
        for (int j = 0; j < n; j++) big = std::max(big, fabs(A(i, j)));
        if (big == 0.) throw std::runtime_error("singular");
        vv[i] = 1. / big;
    }
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < j; i++) {
            double s = A(i, j);
            for (int k = 0; k < i; k++) s -= A(i, k) * A(k, j);
            A(i, j) = s;
        }
        double big = 0.; int imax = j;
        for (int i = j; i < n; i++) {
            double s = A(i, j);
            for (int k = 0; k < j; k++) s -= A(i, k) * A(k, j);
            A(i, j) = s;
            double dum = vv[i] * fabs(s);
            if (dum >= big) { imax = i; big = dum; }
        }
        if (j != imax) {
            for (int k = 0; k < n; k++) std::swap(A(imax, k), A(j, k));
            vv[imax] = vv[j];
        }
        indx[j] = imax;
        if (j != n - 1) {
            if (A(j, j) == 0.) A(j, j) = 1.0e-20;
            double dum = 1. / A(j, j);
            for (int i = j + 1; i < n; i++) A(i, j) *= dum;
        }
    }
    if (A(n - 1, n - 1) == 0.) A(n - 1, n - 1) = 1.0e-20;
    for (int col = 0; col < n; col++) {
        double* b = y + (size_t)n * col;
        for (int i = 0; i < n; i++) b[i] = (i == col) ? 1. : 0.;
        int ii = -1;
        for (int i = 0; i < n; i++) {
            int ll = indx[i];
            double s = b[ll];
            b[ll] = b[i];
            if (ii >= 0) { for (int j = ii; j < i; j++) s -= A(i, j) * b[j]; }
            else if (s != 0.) ii = i;
            b[i] = s;
        }
        for (int i = n - 1; i >= 0; i--) {
            double s = b[i];
            for (int j = i + 1; j < n; j++) s -= A(i, j) * b[j];
            b[i] = s / A(i, i);
        }
    }
}

void build_implicit(Tables& t, double dt) {   // implicit.f90:36-165
    const int kx = t.d.kx, mx = t.d.mx, nx = t.d.nx;
    const Consts& c = t.c;
    ImplicitTables& p = t.imp;
    p.dt = dt;
    size_t N = (size_t)mx * nx;
    p.dmp1.resize(N); p.dmp1d.resize(N); p.dmp1s.resize(N); p.elz.resize(N);
    for (size_t i = 0; i < N; i++) {
        p.dmp1[i] = 1. / (1. + t.dmp[i] * dt);
        p.dmp1d[i] = 1. / (1. + t.dmpd[i] * dt);
        p.dmp1s[i] = 1. / (1. + t.dmps[i] * dt);
    }
    double rgam = c.rgas * c.gamma / (1000. * c.grav);
    p.tref.resize(kx); p.tref1.resize(kx); p.tref2.resize(kx); p.tref3.resize(kx);
    for (int k = 0; k < kx; k++) {
        p.tref[k] = 288. * pow(std::max((double)0.2f, t.fsg[k]), rgam);
        p.tref1[k] = c.rgas * p.tref[k];
        p.tref2[k] = c.akap * p.tref[k];
        p.tref3[k] = t.fsgr[k] * p.tref[k];
    }
    double xi = dt * c.alph;
    double xxi = xi / (c.rearth * c.rearth);
    p.dhsx.resize(kx);
    for (int k = 0; k < kx; k++) p.dhsx[k] = xi * t.dhs[k];
    for (int n = 0; n < nx; n++)
        for (int m = 0; m < mx; m++) p.elz[m + (size_t)mx * n] = (double)(float)(m + n) * (double)(float)(m + n + 1) * xxi;
    std::vector<double> xa((size_t)kx * kx, 0.0), xb((size_t)kx * kx, 0.0), ya((size_t)kx * kx, 0.0), xe((size_t)kx * kx, 0.0), dsum(kx);
    p.xc.assign((size_t)kx * kx, 0.0); p.xd.assign((size_t)kx * kx, 0.0);
    auto M = [&](std::vector<double>& v, int k, int k1) -> double& { return v[k + (size_t)kx * k1]; };
    for (int k = 0; k < kx; k++)
        for (int k1 = 0; k1 < kx; k1++) M(ya, k, k1) = -c.akap * p.tref[k] * t.dhs[k1];
    for (int k = 1; k < kx; k++) M(xa, k, k - 1) = 0.5 * (c.akap * p.tref[k] / t.fsg[k] - (p.tref[k] - p.tref[k - 1]) / t.dhs[k]);
    for (int k = 0; k < kx - 1; k++) M(xa, k, k) = 0.5 * (c.akap * p.tref[k] / t.fsg[k] - (p.tref[k + 1] - p.tref[k]) / t.dhs[k]);
    dsum[0] = t.dhs[0];
    for (int k = 1; k < kx; k++) dsum[k] = dsum[k - 1] + t.dhs[k];
    for (int k = 0; k < kx - 1; k++)
        for (int k1 = 0; k1 < kx; k1++) {
            M(xb, k, k1) = t.dhs[k1] * dsum[k];
            if (k1 <= k) M(xb, k, k1) = M(xb, k, k1) - t.dhs[k1];
        }
    for (int k = 0; k < kx; k++)
        for (int k1 = 0; k1 < kx; k1++) {
            double s = M(ya, k, k1);
            for (int k2 = 0; k2 < kx - 1; k2++) s = s + M(xa, k, k2) * M(xb, k2, k1);
            M(p.xc, k, k1) = s;
        }
    for (int k = 0; k < kx; k++) {
        for (int k1 = k + 1; k1 < kx; k1++) M(p.xd, k, k1) = c.rgas * log(t.hsg[k1 + 1] / t.hsg[k1]);
        M(p.xd, k, k) = c.rgas * log(t.hsg[k + 1] / t.fsg[k]);
    }
    for (int k = 0; k < kx; k++)
        for (int k1 = 0; k1 < kx; k1++) {
            double s = 0.;
            for (int k2 = 0; k2 < kx; k2++) s = s + M(p.xd, k, k2) * M(p.xc, k2, k1);
            M(xe, k, k1) = s;
        }
    const int nl = mx + nx + 1;
    p.xj.assign((size_t)kx * kx * nl, 0.0);
    std::vector<double> xf((size_t)kx * kx);
    for (int l = 1; l <= nl; l++) {
        double xxx = ((double)(float)l * (double)(float)(l + 1)) / (c.rearth * c.rearth);
        for (int k = 0; k < kx; k++)
            for (int k1 = 0; k1 < kx; k1++) M(xf, k, k1) = xi * xi * xxx * (c.rgas * p.tref[k] * t.dhs[k1] - M(xe, k, k1));
        for (int k = 0; k < kx; k++) M(xf, k, k) = M(xf, k, k) + 1.;
        invert_matrix(xf.data(), p.xj.data() + (size_t)kx * kx * (l - 1), kx);
    }
    for (auto& v : p.xc) v = v * xi;
}

void build_tables(int trunc, Tables& t, int nsteps) {
    if (trunc != 30 && trunc != 47) throw std::runtime_error("only T30 and T47 are supported");
    if (nsteps < 1 || (24 * 60) % nsteps != 0) throw std::runtime_error("nsteps must divide the 1440 minutes of a day (date.f90:113)");
    t.d.trunc = trunc;
    t.d.ix = (trunc == 30) ? 96 : 144;
    t.d.iy = t.d.ix / 4; t.d.il = 2 * t.d.iy; t.d.kx = 8; t.d.nx = trunc + 2; t.d.mx = trunc + 1; t.d.ntr = 1;
    t.c = make_consts();
    t.c.nsteps = nsteps;
    t.c.delt = (double)(86400.0f / (float)nsteps);   // params.f90:31, a real32 quotient (2400 at the reference's 36 steps/day)
    build_geometry(t);
    factorize(t.d.ix, t.fft_fac);
    make_twiddles(t.d.ix, t.fft_fac, t.fft_work);
    build_dft(t);
    build_legendre(t);
    build_spectral(t);
    build_dynamics_tables(t);
    build_implicit(t, 0.5 * t.c.delt);
}

std::map<std::string, std::vector<double>*> Tables::named() {
    return {
        {"hsg", &hsg}, {"dhs", &dhs}, {"fsg", &fsg}, {"dhsr", &dhsr}, {"fsgr", &fsgr},
        {"radang", &radang}, {"coriol", &coriol}, {"sia", &sia}, {"coa", &coa}, {"cosg", &cosg},
        {"cosgr", &cosgr}, {"cosgr2", &cosgr2}, {"sia_half", &sia_half}, {"coa_half", &coa_half},
        {"wt", &wt}, {"epsi", &epsi}, {"repsi", &repsi}, {"poly", &poly}, {"cpol", &cpol},
        {"fft_work", &fft_work}, {"finv", &finv}, {"ffwd", &ffwd},
        {"el2", &el2}, {"elm2", &elm2}, {"el4", &el4}, {"trfilt", &trfilt}, {"gradx", &gradx},
        {"gradym", &gradym}, {"gradyp", &gradyp}, {"uvdx", &uvdx}, {"uvdym", &uvdym}, {"uvdyp", &uvdyp},
        {"vddym", &vddym}, {"vddyp", &vddyp}, {"xgeop1", &xgeop1}, {"xgeop2", &xgeop2}, {"geop_corf", &geop_corf},
        {"dmp", &dmp}, {"dmpd", &dmpd}, {"dmps", &dmps}, {"tcorv", &tcorv}, {"qcorv", &qcorv},
        {"sigl", &sigl}, {"sigh", &sigh}, {"grdsig", &grdsig}, {"grdscp", &grdscp}, {"wvi", &wvi}, {"fband", &fband},
        {"dmp1", &imp.dmp1}, {"dmp1d", &imp.dmp1d}, {"dmp1s", &imp.dmp1s}, {"tref", &imp.tref}, {"tref1", &imp.tref1},
        {"tref2", &imp.tref2}, {"tref3", &imp.tref3}, {"xc", &imp.xc}, {"xd", &imp.xd}, {"xj", &imp.xj},
        {"dhsx", &imp.dhsx}, {"elz", &imp.elz},
    };
}

}  // namespace spd
