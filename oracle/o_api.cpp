/* ORACLE (test infrastructure) — extern "C" surface used by tests/ via ctypes. */
#include "oracle.h"
#include <string>
using namespace orc;

extern "C" {

int orc_dims(int* out) {
    out[0] = trunc_; out[1] = ix; out[2] = iy; out[3] = il; out[4] = kx; out[5] = nx; out[6] = mx; out[7] = ntr;
    return 0;
}

static bool g_init_transforms = false;
int orc_init_transforms() {
    if (g_init_transforms) return 0;
    initialize_geometry();
    initialize_spectral();
    g_init_transforms = true;
    return 0;
}

/* copy a named table into out (size checked) */
int orc_get_table(const char* name, double* out, int n) {
    auto cpy = [&](const double* src, int cnt) { if (cnt != n) return -2; memcpy(out, src, sizeof(double) * cnt); return 0; };
    std::string s(name);
    if (s == "wt") return cpy(wt + 1, iy);
    if (s == "sia_half") return cpy(geo.sia_half + 1, iy);
    if (s == "sia") return cpy(geo.sia + 1, il);
    if (s == "coa") return cpy(geo.coa + 1, il);
    if (s == "cosgr") return cpy(geo.cosgr + 1, il);
    if (s == "cosgr2") return cpy(geo.cosgr2 + 1, il);
    if (s == "coriol") return cpy(geo.coriol + 1, il);
    if (s == "radang") return cpy(geo.radang + 1, il);
    if (s == "hsg") return cpy(geo.hsg + 1, kx + 1);
    if (s == "fsg") return cpy(geo.fsg + 1, kx);
    if (s == "dhs") return cpy(geo.dhs + 1, kx);
    if (s == "cpol") return cpy(cpol.p(), 2 * mx * nx * iy);
    if (s == "epsi") return cpy(epsi.p(), (mx + 1) * (nx + 1));
    if (s == "fft_work") return cpy(fft_work + 1, ix);
    if (s == "el2") return cpy(el2.p(), mx * nx);
    if (s == "elm2") return cpy(elm2.p(), mx * nx);
    if (s == "trfilt") return cpy(trfilt.p(), mx * nx);
    if (s == "gradx") return cpy(gradx + 1, mx);
    if (s == "gradym") return cpy(gradym.p(), mx * nx);
    if (s == "gradyp") return cpy(gradyp.p(), mx * nx);
    if (s == "uvdx") return cpy(uvdx.p(), mx * nx);
    if (s == "uvdym") return cpy(uvdym.p(), mx * nx);
    if (s == "uvdyp") return cpy(uvdyp.p(), mx * nx);
    if (s == "vddym") return cpy(vddym.p(), mx * nx);
    if (s == "vddyp") return cpy(vddyp.p(), mx * nx);
    return -1;
}
int orc_get_itable(const char* name, int* out, int n) {
    std::string s(name);
    if (s == "nsh2") { if (n != nx) return -2; memcpy(out, nsh2 + 1, sizeof(int) * nx); return 0; }
    if (s == "ifac") { if (n != 15) return -2; memcpy(out, fft_ifac + 1, sizeof(int) * 15); return 0; }
    return -1;
}

int orc_rfftf(double* x /*ix*/) { std::vector<double> ch(ix + 1); rfftf1(ix, x - 1, ch.data(), fft_work, fft_ifac); return 0; }
int orc_rfftb(double* x /*ix*/) { std::vector<double> ch(ix + 1); rfftb1(ix, x - 1, ch.data(), fft_work, fft_ifac); return 0; }

int orc_legendre_inv(const double* in, double* out) { legendre_inv(in, out); return 0; }
int orc_legendre_dir(const double* in, double* out) { legendre_dir(in, out); return 0; }
int orc_fourier_inv(const double* in, int kcos, double* out) { fourier_inv(in, kcos, out); return 0; }
int orc_fourier_dir(const double* in, double* out) { fourier_dir(in, out); return 0; }
int orc_spec_to_grid(const double* spec, int nbatch, const int* kcos, double* grid) {
    for (int b = 0; b < nbatch; b++)
        spec_to_grid(reinterpret_cast<const cplx*>(spec) + (size_t)b * mx * nx, kcos[b], grid + (size_t)b * ix * il);
    return 0;
}
int orc_grid_to_spec(const double* grid, int nbatch, double* spec) {
    for (int b = 0; b < nbatch; b++)
        grid_to_spec(grid + (size_t)b * ix * il, reinterpret_cast<cplx*>(spec) + (size_t)b * mx * nx);
    return 0;
}
int orc_uvspec(const double* vorm, const double* divm, double* ucosm, double* vcosm) {
    uvspec((const cplx*)vorm, (const cplx*)divm, (cplx*)ucosm, (cplx*)vcosm); return 0; }
int orc_grad(const double* psi, double* psdx, double* psdy) { grad((const cplx*)psi, (cplx*)psdx, (cplx*)psdy); return 0; }
int orc_vds(const double* u, const double* v, double* vorm, double* divm) { vds((const cplx*)u, (const cplx*)v, (cplx*)vorm, (cplx*)divm); return 0; }
int orc_vdspec(const double* ug, const double* vg, double* vorm, double* divm, int kcos) { vdspec(ug, vg, (cplx*)vorm, (cplx*)divm, kcos); return 0; }
int orc_laplacian(const double* in, double* out) { laplacian((const cplx*)in, (cplx*)out); return 0; }
int orc_inverse_laplacian(const double* in, double* out) { inverse_laplacian((const cplx*)in, (cplx*)out); return 0; }
int orc_trunct(double* x) { trunct((cplx*)x); return 0; }

}  // extern "C"
