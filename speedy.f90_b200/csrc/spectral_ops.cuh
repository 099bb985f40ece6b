// Device inline forms of the reference's spectral operators (spectral.f90:84-96,124-233).
// Fields are complex(mx,nx), interleaved (re,im), m fastest.  All operators are
// 3-point stencils in n at fixed m; `f(m,n)` style accessors read global memory.
#pragma once
#include "ctx.h"

namespace spd {

struct cd { double re, im; };
__device__ __forceinline__ cd ld(const double* f, int mx, int m, int n) {
    const double2 v = *reinterpret_cast<const double2*>(f + 2 * (m + (size_t)mx * n));
    return cd{v.x, v.y};
}
__device__ __forceinline__ void st(double* f, int mx, int m, int n, cd v) {
    *reinterpret_cast<double2*>(f + 2 * (m + (size_t)mx * n)) = make_double2(v.re, v.im);
}
__device__ __forceinline__ cd operator*(double a, cd b) { return cd{a * b.re, a * b.im}; }
__device__ __forceinline__ cd operator+(cd a, cd b) { return cd{a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cd operator-(cd a, cd b) { return cd{a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ cd neg(cd a) { return cd{-a.re, -a.im}; }
__device__ __forceinline__ cd times_i(cd a) { return cd{-a.im, a.re}; }

// uvspec  spectral.f90:173-196; the _t forms take the three operator-table entries of (m,n) from the caller
__device__ __forceinline__ void dev_uvspec_t(int mx, int nx, int tr, const double* vor, const double* div, int m, int n,
                                             double dx, double dym, double dyp, cd& uc, cd& vc) {
    const cd zp = times_i(dx * ld(vor, mx, m, n));
    const cd zc = times_i(dx * ld(div, mx, m, n));
    if (n == 0) {
        uc = zc - dyp * ld(vor, mx, m, 1);
        vc = zp + dyp * ld(div, mx, m, 1);
    } else if (n == nx - 1) {
        uc = dym * ld(vor, mx, m, tr);
        vc = neg(dym * ld(div, mx, m, tr));
    } else {
        vc = (neg(dym * ld(div, mx, m, n - 1)) + dyp * ld(div, mx, m, n + 1)) + zp;
        uc = (dym * ld(vor, mx, m, n - 1) - dyp * ld(vor, mx, m, n + 1)) + zc;
    }
}
__device__ __forceinline__ void dev_uvspec(const DevTables& tv, const double* vor, const double* div, int m, int n, cd& uc, cd& vc) {
    const size_t q = m + (size_t)tv.mx * n;
    const double dx = tv.uvdx[q], dym = tv.uvdym[q], dyp = tv.uvdyp[q];   // loaded ahead of the branches: one round trip
    dev_uvspec_t(tv.mx, tv.nx, tv.trunc, vor, div, m, n, dx, dym, dyp, uc, vc);
}

// grad  spectral.f90:124-144
__device__ __forceinline__ void dev_grad_t(int mx, int nx, int tr, const double* psi, int m, int n, double gx, double gym, double gyp, cd& dx, cd& dy) {
    dx = times_i(gx * ld(psi, mx, m, n));
    if (n == 0) dy = gyp * ld(psi, mx, m, 1);
    else if (n == nx - 1) dy = neg(gym * ld(psi, mx, m, tr));
    else dy = neg(gym * ld(psi, mx, m, n - 1)) + gyp * ld(psi, mx, m, n + 1);
}
__device__ __forceinline__ void dev_grad(const DevTables& tv, const double* psi, int m, int n, cd& dx, cd& dy) {
    const size_t q = m + (size_t)tv.mx * n;
    const double gx = tv.gradx[m], gym = tv.gradym[q], gyp = tv.gradyp[q];
    dev_grad_t(tv.mx, tv.nx, tv.trunc, psi, m, n, gx, gym, gyp, dx, dy);
}

// one component of uvspec / grad at (m,n): the K1 input stage builds ONE derived field per transform, so it evaluates (and
// loads the stencil of) that component only.  Same expressions as above.
__device__ __forceinline__ cd dev_ucos_t(int mx, int nx, int tr, const double* vor, const double* div, int m, int n, double dx, double dym, double dyp) {
    const cd zc = times_i(dx * ld(div, mx, m, n));
    if (n == 0) return zc - dyp * ld(vor, mx, m, 1);
    if (n == nx - 1) return dym * ld(vor, mx, m, tr);
    return (dym * ld(vor, mx, m, n - 1) - dyp * ld(vor, mx, m, n + 1)) + zc;
}
__device__ __forceinline__ cd dev_vcos_t(int mx, int nx, int tr, const double* vor, const double* div, int m, int n, double dx, double dym, double dyp) {
    const cd zp = times_i(dx * ld(vor, mx, m, n));
    if (n == 0) return zp + dyp * ld(div, mx, m, 1);
    if (n == nx - 1) return neg(dym * ld(div, mx, m, tr));
    return (neg(dym * ld(div, mx, m, n - 1)) + dyp * ld(div, mx, m, n + 1)) + zp;
}
__device__ __forceinline__ cd dev_gradx_t(int mx, const double* psi, int m, int n, double gx) { return times_i(gx * ld(psi, mx, m, n)); }
__device__ __forceinline__ cd dev_grady_t(int mx, int nx, int tr, const double* psi, int m, int n, double gym, double gyp) {
    if (n == 0) return gyp * ld(psi, mx, m, 1);
    if (n == nx - 1) return neg(gym * ld(psi, mx, m, tr));
    return neg(gym * ld(psi, mx, m, n - 1)) + gyp * ld(psi, mx, m, n + 1);
}

// vds  spectral.f90:146-171
__device__ __forceinline__ void dev_vds(const DevTables& tv, const double* uc, const double* vc, int m, int n, cd& vor, cd& div) {
    const int mx = tv.mx, nx = tv.nx, tr = tv.trunc;
    const size_t q = m + (size_t)mx * n;
    const cd zp = times_i(tv.gradx[m] * ld(uc, mx, m, n));
    const cd zc = times_i(tv.gradx[m] * ld(vc, mx, m, n));
    if (n == 0) {
        vor = zc - tv.vddyp[q] * ld(uc, mx, m, 1);
        div = zp + tv.vddyp[q] * ld(vc, mx, m, 1);
    } else if (n == nx - 1) {
        vor = tv.vddym[q] * ld(uc, mx, m, tr);
        div = neg(tv.vddym[q] * ld(vc, mx, m, tr));
    } else {
        vor = (tv.vddym[q] * ld(uc, mx, m, n - 1) - tv.vddyp[q] * ld(uc, mx, m, n + 1)) + zc;
        div = (neg(tv.vddym[q] * ld(vc, mx, m, n - 1)) + tv.vddyp[q] * ld(vc, mx, m, n + 1)) + zp;
    }
}

// vds on operands already in registers (um/u0/up, vm/v0/vp: ucos, vcos at n-1, n, n+1 with the neighbour index clamped at the
// edges, where it is not used): the caller loads them and the three table entries unconditionally, all at once, so that
// the stencil branches do not serialise the global-memory round trips.  Same expressions as dev_vds.
__device__ __forceinline__ void dev_vds_r(int nx, int n, double gx, double dym, double dyp, cd um, cd u0, cd up, cd vm, cd v0, cd vp, cd& vor, cd& div) {
    const cd zp = times_i(gx * u0);
    const cd zc = times_i(gx * v0);
    if (n == 0) {
        vor = zc - dyp * up;
        div = zp + dyp * vp;
    } else if (n == nx - 1) {
        vor = dym * um;
        div = neg(dym * vm);
    } else {
        vor = (dym * um - dyp * up) + zc;
        div = (neg(dym * vm) + dyp * vp) + zp;
    }
}
// the divergence half alone (spectral.f90:160-168)
__device__ __forceinline__ cd dev_vds_div_r(int nx, int n, double gx, double dym, double dyp, cd u0, cd vm, cd vp) {
    const cd zp = times_i(gx * u0);
    if (n == 0) return zp + dyp * vp;
    if (n == nx - 1) return neg(dym * vm);
    return (neg(dym * vm) + dyp * vp) + zp;
}

}  // namespace spd
