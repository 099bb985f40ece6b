// Host-side table builders of the product (B200 library).  These mirror what the
// reference computes once at start-up (initialize_geometry / _legendre / _fourier /
// _spectral / _geopotential / _horizontal_diffusion / _physics / _implicit, radset) so
// that the CUDA kernels can be fed bit-compatible constants.  0-based, C order unless
// noted.  Single-precision literals of the reference are preserved.
#pragma once
#include <vector>
#include <string>
#include <map>

namespace spd {

struct Dims {
    int trunc, ix, iy, il, kx, nx, mx, ntr;
    int nspec() const { return mx * nx; }      // complex coefficients per 2-D spectral field
    int ngrid() const { return ix * il; }      // points per 2-D grid field
    int k2() const { return 2 * mx; }          // Fourier rows (re,im interleaved)
    int k2pad() const { return (2 * mx + 3) / 4 * 4; }
};

struct Consts {
    double rearth, omega, grav, p0, cp, akap, rgas, alhc, alhs, sbc;
    double gamma, hscale, hshum, refrh1, thd, thdd, thds, tdrs;
    double rob, wil, alph, delt;
    int nsteps, nstrad;
};
Consts make_consts();

struct ImplicitTables {
    double dt;
    std::vector<double> dmp1, dmp1d, dmp1s;           // (mx,nx) m fastest
    std::vector<double> tref, tref1, tref2, tref3;    // kx
    std::vector<double> xc, xd;                        // (kx,kx) Fortran order: (k,k1) -> k + kx*k1
    std::vector<double> xj;                            // (kx,kx,mx+nx+1) Fortran order
    std::vector<double> dhsx;                          // kx
    std::vector<double> elz;                           // (mx,nx)
};

struct Tables {
    Dims d;
    Consts c;
    // geometry.f90
    std::vector<double> hsg, dhs, fsg, dhsr, fsgr;                       // kx+1 / kx
    std::vector<double> radang, coriol, sia, coa, cosg, cosgr, cosgr2;   // il
    std::vector<double> sia_half, coa_half;                              // iy
    // legendre.f90
    std::vector<double> wt;                 // iy
    std::vector<double> epsi, repsi;        // (mx+1,nx+1) Fortran order m fastest
    std::vector<int> nsh2;                  // nx
    std::vector<double> poly;               // unique P: [j][n][m] (m fastest), iy*nx*mx
    std::vector<double> cpol;               // reference layout cpol(2*mx,nx,iy) (for get_table)
    // fourier.f90 / fftpack.f90
    std::vector<double> fft_work;           // ix (twiddles as rffti1)
    std::vector<int> fft_fac;               // factor list in rffti1 order
    std::vector<double> finv;               // dense backward operator [ix][k2pad]: grid(i) = sum_c finv[i][c]*four(c)
    std::vector<double> ffwd;               // dense forward operator  [k2pad][ix]: four(c) = sum_i ffwd[c][i]*grid(i)
    // spectral.f90
    std::vector<double> el2, elm2, el4, trfilt;                          // (mx,nx)
    std::vector<double> gradx;                                           // mx
    std::vector<double> gradym, gradyp, uvdx, uvdym, uvdyp, vddym, vddyp;// (mx,nx)
    // geopotential.f90
    std::vector<double> xgeop1, xgeop2;     // kx
    std::vector<double> geop_corf;          // kx (lapse-rate correction factor, 0 outside 2..kx-1)
    // horizontal_diffusion.f90
    std::vector<double> dmp, dmpd, dmps;    // (mx,nx)
    std::vector<double> tcorv, qcorv;       // kx
    // physics.f90 / physical_constants
    std::vector<double> sigl, sigh, grdsig, grdscp, wvi; // kx, kx+1 (sigh(0:kx)), kx, kx, (kx,2) Fortran order
    // longwave radset
    std::vector<double> fband;              // (301,4) Fortran order, first index = T-100
    ImplicitTables imp;

    std::map<std::string, std::vector<double>*> named();
};

void build_tables(int trunc, Tables& t, int nsteps = 36);   // nsteps: time steps per day (params.f90:30)
void build_implicit(Tables& t, double dt);
// product-side real FFT (FFTPACK algorithm, reference constants); x has n elements, in place
void rfft_forward(const Tables& t, double* x);
void rfft_backward(const Tables& t, double* x);
void invert_matrix(double* a, double* y, int n);   // a destroyed (LU in place), y = a^{-1}; Fortran order

}  // namespace spd
