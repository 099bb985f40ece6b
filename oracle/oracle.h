/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the speedy.f90 hot path (reference: the .f90 files under /root/reference/source).
 * It is the checker for the CUDA product in speedy.f90_b200/; nothing in the product may
 * include, link or call it.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it.
 *
 * PARITY UNPINNED: the reference ships no golden vectors and cannot be compiled here
 * (no Fortran compiler, no NetCDF) — see DESIGN.md.  The restatement follows the
 * reference statement by statement, including its single-precision literals
 * (SURVEY.md F8), approximate Gaussian latitudes (F9) and FFTPACK constants (F13).
 *
 * Conventions: Fortran-order, 1-based arrays through FA<>; every un-suffixed Fortran
 * real literal that is not exactly representable is written (double)x.xxf.
 * Build with -ffp-contract=off.
 */
#pragma once
#include <complex>
#include <vector>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>

#ifndef TRUNC
#define TRUNC 30
#endif
#ifndef IXDEF
#define IXDEF 96
#endif

namespace orc {

typedef std::complex<double> cplx;

/* params.f90:19-33 */
constexpr int trunc_ = TRUNC;
constexpr int ix = IXDEF;
constexpr int iy = IXDEF / 4;
constexpr int il = 2 * iy;
constexpr int kx = 8;
constexpr int nx = trunc_ + 2;
constexpr int mx = trunc_ + 1;
constexpr int ntr = 1;
#ifndef NSTEPS
#define NSTEPS 36                          /* params.f90:30; a compile-time parameter of the reference as well */
#endif
constexpr int nsteps = NSTEPS;
constexpr double delt = (double)(86400.0f / (float)NSTEPS);   /* params.f90:31 in real32: 2400 exactly at 36 steps/day */
const double rob = (double)0.05f;
const double wil = (double)0.53f;
constexpr double alph = 0.5;
constexpr int nstrad = 3;
constexpr int issty0 = 1979;

/* physical_constants.f90:17-29 */
constexpr double rearth = 6.371e+6;        /* exact in real32 */
const double omega = (double)7.292e-05f;
const double grav = (double)9.81f;
constexpr double p0 = 1.e+5;
constexpr double cp = 1004.0;
const double akap = (double)(2.0f / 7.0f);
const double rgas = akap * cp;
constexpr double alhc = 2501.0;
constexpr double alhs = 2801.0;
const double sbc = (double)5.67e-8f;

/* dynamical_constants.f90:12-22 */
constexpr double gamma_ = 6.0;
constexpr double hscale = 7.5;
constexpr double hshum = 2.5;
const double refrh1 = (double)0.7f;
const double thd = (double)2.4f;
const double thdd = (double)2.4f;
constexpr double thds = 12.0;
constexpr double tdrs = 24.0 * 30.0;

/* Fortran-order 1-based fixed-extent array */
template <class T, int N1, int N2 = 1, int N3 = 1, int N4 = 1, int N5 = 1>
struct FA {
    std::vector<T> d;
    FA() : d((size_t)N1 * N2 * N3 * N4 * N5, T()) {}
    inline T& operator()(int i, int j = 1, int k = 1, int l = 1, int m = 1) {
        return d[(size_t)(i - 1) + (size_t)N1 * ((j - 1) + (size_t)N2 * ((k - 1) + (size_t)N3 * ((l - 1) + (size_t)N4 * (m - 1))))];
    }
    inline const T& operator()(int i, int j = 1, int k = 1, int l = 1, int m = 1) const {
        return d[(size_t)(i - 1) + (size_t)N1 * ((j - 1) + (size_t)N2 * ((k - 1) + (size_t)N3 * ((l - 1) + (size_t)N4 * (m - 1))))];
    }
    T* p(int i = 1, int j = 1, int k = 1, int l = 1, int m = 1) { return &(*this)(i, j, k, l, m); }
    const T* p(int i = 1, int j = 1, int k = 1, int l = 1, int m = 1) const { return &(*this)(i, j, k, l, m); }
    void fill(T v) { std::fill(d.begin(), d.end(), v); }
    size_t size() const { return d.size(); }
};

typedef FA<double, ix, il> Grid2;
typedef FA<double, ix, il, kx> Grid3;
typedef FA<cplx, mx, nx> Spec2;
typedef FA<cplx, mx, nx, kx> Spec3;

/* ---- fftpack.f90 (o_fftpack.cpp) ---- */
void rffti1(int n, double* wa, int* ifac);
void rfftb1(int n, double* c, double* ch, const double* wa, const int* ifac);
void rfftf1(int n, double* c, double* ch, const double* wa, const int* ifac);

/* ---- geometry.f90 ---- */
struct Geometry {
    double hsg[kx + 2], dhs[kx + 1], fsg[kx + 1], dhsr[kx + 1], fsgr[kx + 1]; /* 1-based */
    double radang[il + 1], coriol[il + 1], sia[il + 1], coa[il + 1], sia_half[iy + 1], coa_half[il + 1];
    double cosg[il + 1], cosgr[il + 1], cosgr2[il + 1];
};
extern Geometry geo;
void initialize_geometry();

/* ---- legendre.f90 ---- */
extern FA<double, 2 * mx, nx, iy> cpol;
extern FA<double, mx + 1, nx + 1> epsi, repsi;
extern int nsh2[nx + 1];
extern double wt[iy + 1];
void initialize_legendre();
void legendre_inv(const double* input /*(2mx,nx)*/, double* output /*(2mx,il)*/);
void legendre_dir(const double* input /*(2mx,il)*/, double* output /*(2mx,nx)*/);

/* ---- fourier.f90 ---- */
extern double fft_work[ix + 1];
extern int fft_ifac[16];
void initialize_fourier();
void fourier_inv(const double* input /*(2mx,il)*/, int kcos, double* output /*(ix,il)*/);
void fourier_dir(const double* input /*(ix,il)*/, double* output /*(2mx,il)*/);

/* ---- spectral.f90 ---- */
extern FA<double, mx, nx> el2, elm2, el4, trfilt;
extern double gradx[mx + 1];
extern FA<double, mx, nx> gradym, gradyp, uvdx, uvdym, uvdyp, vddym, vddyp;
void initialize_spectral();
void laplacian(const cplx* in, cplx* out);
void inverse_laplacian(const cplx* in, cplx* out);
void spec_to_grid(const cplx* vorm, int kcos, double* vorg);
void grid_to_spec(const double* vorg, cplx* vorm);
void grad(const cplx* psi, cplx* psdx, cplx* psdy);
void vds(const cplx* ucosm, const cplx* vcosm, cplx* vorm, cplx* divm);
void uvspec(const cplx* vorm, const cplx* divm, cplx* ucosm, cplx* vcosm);
void vdspec(const double* ug, const double* vg, cplx* vorm, cplx* divm, int kcos);
void trunct(cplx* vor);


/* ======================= model state and procedures (o_dynamics / o_physics / o_env) ======================= */
typedef FA<cplx, mx, nx, kx, 2> Spec4;

/* prognostics.f90:16-24 */
extern Spec4 vor, div_, t;
extern FA<cplx, mx, nx, 2> ps;
extern FA<cplx, mx, nx, kx, 2, ntr> tr;
extern Spec3 phi;
extern Spec2 phis;
void initialize_prognostics();

/* geopotential.f90 */
extern double xgeop1[kx + 1], xgeop2[kx + 1];
void initialize_geopotential();
void get_geopotential(const cplx* t /*(mx,nx,kx)*/, const cplx* phis /*(mx,nx)*/, cplx* phi /*(mx,nx,kx)*/);

/* horizontal_diffusion.f90 */
extern FA<double, mx, nx> dmp, dmpd, dmps, dmp1, dmp1d, dmp1s;
extern double tcorv[kx + 1], qcorv[kx + 1];
extern Spec2 tcorh, qcorh;
void initialize_horizontal_diffusion();

/* implicit.f90 */
extern double tref[kx + 1], tref1[kx + 1], tref2[kx + 1], tref3[kx + 1], dhsx[kx + 1];
extern FA<double, kx, kx> xa, xb, xc, xd, xe;
extern FA<double, kx, kx, mx + nx + 1> xf, xj;
extern FA<double, mx, nx> elz;
void initialize_implicit(double dt);
void implicit_terms(Spec3& divdt, Spec3& tdt, Spec2& psdt);

/* tendencies.f90 / time_stepping.f90 */
void get_tendencies(Spec3& vordt, Spec3& divdt, Spec3& tdt, Spec2& psdt, FA<cplx, mx, nx, kx, ntr>& trdt, int j2);
void step(int j1, int j2, double dt);
void first_step();

/* diagnostics.f90: returns 1 when the reference would `stop` */
int check_diagnostics(const cplx* vor, const cplx* div, const cplx* t, int istep, double* diag /*(kx,3)*/, bool print);

/* physical_constants.f90:31-37 / physics.f90 */
extern double sigl[kx + 1], sigh[kx + 1] /*0..kx*/, grdsig[kx + 1], grdscp[kx + 1];
extern FA<double, kx, 2> wvi;
void initialize_physics();
void get_physical_tendencies(const cplx* vor, const cplx* div, const cplx* t, const cplx* q, const cplx* phi, const cplx* psl,
                             Grid3& utend, Grid3& vtend, Grid3& ttend, Grid3& qtend);

/* auxiliaries.f90 */
extern Grid2 precnv, precls, snowcv, snowls, cbmf, tsr, ssrd, ssr, slrd, slr, olr;
extern FA<double, ix, il, 3> slru, ustr, vstr, shf, evap, hfluxn;

/* mod_radcon.f90 */
extern double albsea, albice, albsn, epslw, emisfc, ablco2_ref;
extern FA<double, 301, 4> fband; /* fband(100:400,4): first index = T-99 */
extern Grid2 alb_l, alb_s, albsfc, snowc;
extern FA<double, ix, il, kx, 4> tau2;
extern FA<double, ix, il, kx, 2> st4a;
extern FA<double, ix, il, 2> stratc;
extern FA<double, ix, il, 4> flux;

/* shortwave_radiation.f90 */
extern double ablco2;
extern Grid2 fsol, ozone, ozupp, zenit, stratz, qcloud;
extern bool compute_shortwave;
void get_zonal_average_fields(double tyear);
void radset();

/* surface_fluxes.f90 */
extern Grid2 forog;
void set_orog_land_sfc_drag(const Grid2& phi0);

/* humidity.f90 */
void get_qsat(const double* ta, const double* ps, double sig, double* qsat); /* (ix,il) */

/* boundaries.f90 */
extern Grid2 fmask, phi0, phis0, alb0;
/* land_model.f90 / sea_model.f90 public state */
extern Grid2 stl_am, snowd_am, soilw_am, fmask_l, stl_lm;
extern Grid2 fmask_s, sstcl_ob, sst_am, sice_am, tice_am, ssti_om, sst_om, tice_om, sice_om;
extern int sea_coupling_flag;

/* date.f90 */
struct DateTime { int year, month, day, hour, minute; };
extern DateTime model_datetime, start_datetime, end_datetime;
extern int imont1, isst0;
extern double tmonth, tyear;

/* sppt.f90 */
extern bool sppt_on;
extern Spec3 sppt_eta;   /* test hook: the Gaussian noise eta(m,n,k) is supplied by the caller (the reference seeds from system_clock) */
void gen_sppt(Grid3& sppt_grid);
void sppt_reset();

/* initialization.f90 / speedy.f90 driver (o_env.cpp) */
int model_initialize(const char* bc_file, int y, int m, int d, int h, int mi);
int model_run_steps(int nsteps_to_run);   /* main-loop body speedy.f90:27-54; returns 1 on diagnostics stop */
extern int model_step;
void test_calendar_init(int y, int m, int d, int h, int mi);   /* date.f90 hooks for the CPU tests */
void test_newdate();

}  // namespace orc
