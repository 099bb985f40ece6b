/* ORACLE (test infrastructure) — extern "C" surface used by tests/ via ctypes. */
#include "oracle.h"
#include <string>
#include <map>
namespace orc {
extern FA<int, ix, il> dbg_iptop, dbg_icnv;
extern FA<int, ix, il, 2> dbg_icltop;
extern Grid3 dbg_tt_rsw;
}
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


/* ======================= model-level surface ======================= */
struct FieldRef { double* p; size_t n; };
static std::map<std::string, FieldRef>& registry() {
    static std::map<std::string, FieldRef> r;
    if (r.empty()) {
#define REG(name, obj) r[name] = FieldRef{reinterpret_cast<double*>((obj).p()), (obj).size() * sizeof(*(obj).p()) / sizeof(double)}
        REG("vor", vor); REG("div", div_); REG("t", t); REG("tr", tr); REG("ps", ps); REG("phi", phi); REG("phis", phis);
        REG("tcorh", tcorh); REG("qcorh", qcorh);
        REG("precnv", precnv); REG("precls", precls); REG("cbmf", cbmf); REG("tsr", tsr); REG("ssrd", ssrd); REG("ssr", ssr);
        REG("slrd", slrd); REG("slr", slr); REG("olr", olr); REG("slru", slru); REG("ustr", ustr); REG("vstr", vstr);
        REG("shf", shf); REG("evap", evap); REG("hfluxn", hfluxn);
        REG("alb_l", alb_l); REG("alb_s", alb_s); REG("albsfc", albsfc); REG("snowc", snowc);
        REG("tau2", tau2); REG("st4a", st4a); REG("stratc", stratc); REG("flux", flux);
        REG("fsol", fsol); REG("ozone", ozone); REG("ozupp", ozupp); REG("zenit", zenit); REG("stratz", stratz); REG("qcloud", qcloud);
        REG("forog", forog); REG("fmask", fmask); REG("phi0", phi0); REG("phis0", phis0); REG("alb0", alb0);
        REG("stl_am", stl_am); REG("snowd_am", snowd_am); REG("soilw_am", soilw_am); REG("fmask_l", fmask_l); REG("stl_lm", stl_lm);
        REG("fmask_s", fmask_s); REG("sstcl_ob", sstcl_ob); REG("sst_am", sst_am); REG("sice_am", sice_am); REG("tice_am", tice_am);
        REG("ssti_om", ssti_om); REG("sst_om", sst_om); REG("tice_om", tice_om); REG("sice_om", sice_om);
        REG("tt_rsw", dbg_tt_rsw); REG("fband", fband); REG("sppt_eta", sppt_eta);
        REG("dmp", dmp); REG("dmpd", dmpd); REG("dmps", dmps); REG("dmp1", dmp1); REG("dmp1d", dmp1d); REG("dmp1s", dmp1s);
        REG("xj", xj); REG("xc", xc); REG("xd", xd); REG("elz", elz);
#undef REG
    }
    return r;
}

int orc_get_field(const char* name, double* out, long long n) {
    auto& r = registry();
    auto it = r.find(name);
    if (it == r.end()) return -1;
    if ((long long)it->second.n != n) return -2;
    memcpy(out, it->second.p, sizeof(double) * n);
    return 0;
}
int orc_set_field(const char* name, const double* in, long long n) {
    auto& r = registry();
    auto it = r.find(name);
    if (it == r.end()) return -1;
    if ((long long)it->second.n != n) return -2;
    memcpy(it->second.p, in, sizeof(double) * n);
    return 0;
}
long long orc_field_len(const char* name) {
    auto& r = registry();
    auto it = r.find(name);
    return it == r.end() ? -1 : (long long)it->second.n;
}
int orc_get_ifield(const char* name, int* out, long long n) {
    std::string s(name);
    const int* src = nullptr; size_t cnt = 0;
    if (s == "iptop") { src = dbg_iptop.p(); cnt = dbg_iptop.size(); }
    else if (s == "icnv") { src = dbg_icnv.p(); cnt = dbg_icnv.size(); }
    else if (s == "icltop") { src = dbg_icltop.p(); cnt = (size_t)ix * il; }
    else return -1;
    if ((long long)cnt != n) return -2;
    memcpy(out, src, sizeof(int) * cnt);
    return 0;
}
/* kx-vectors and scalars of the implicit scheme / physics constants */
int orc_get_vec(const char* name, double* out, int n) {
    std::string s(name);
    const double* src = nullptr; int cnt = kx;
    if (s == "tref") src = tref + 1; else if (s == "tref1") src = tref1 + 1; else if (s == "tref2") src = tref2 + 1;
    else if (s == "tref3") src = tref3 + 1; else if (s == "dhsx") src = dhsx + 1; else if (s == "tcorv") src = tcorv + 1;
    else if (s == "qcorv") src = qcorv + 1; else if (s == "xgeop1") src = xgeop1 + 1; else if (s == "xgeop2") src = xgeop2 + 1;
    else if (s == "sigl") src = sigl + 1; else if (s == "grdsig") src = grdsig + 1; else if (s == "grdscp") src = grdscp + 1;
    else if (s == "sigh") { src = sigh; cnt = kx + 1; }
    else if (s == "wvi") { src = wvi.p(); cnt = 2 * kx; }
    else return -1;
    if (cnt != n) return -2;
    memcpy(out, src, sizeof(double) * cnt);
    return 0;
}

int orc_model_init(const char* bc_file, int y, int m, int d, int h, int mi) {
    g_init_transforms = true;
    return model_initialize(bc_file, y, m, d, h, mi);
}
int orc_model_run(int nsteps_to_run) { return model_run_steps(nsteps_to_run); }
int orc_model_date(int* ymdhm, long long* step) {
    ymdhm[0] = model_datetime.year; ymdhm[1] = model_datetime.month; ymdhm[2] = model_datetime.day;
    ymdhm[3] = model_datetime.hour; ymdhm[4] = model_datetime.minute; *step = model_step;
    return 0;
}
int orc_set_date_fractions(double tmonth_, double tyear_, int imont1_) { tmonth = tmonth_; tyear = tyear_; imont1 = imont1_; return 0; }
int orc_initialize_implicit(double dt) { initialize_implicit(dt); return 0; }
/* implicit.f90:168-217 and horizontal_diffusion.f90:86-105 on caller-supplied arrays (in place) */
int orc_implicit_terms(double* divdt, double* tdt, double* psdt) {
    static Spec3 a, b; static Spec2 c;
    memcpy((void*)a.p(), divdt, sizeof(cplx) * a.size()); memcpy((void*)b.p(), tdt, sizeof(cplx) * b.size()); memcpy((void*)c.p(), psdt, sizeof(cplx) * c.size());
    implicit_terms(a, b, c);
    memcpy(divdt, a.p(), sizeof(cplx) * a.size()); memcpy(tdt, b.p(), sizeof(cplx) * b.size()); memcpy(psdt, c.p(), sizeof(cplx) * c.size());
    return 0;
}
int orc_do_horizontal_diffusion(const double* field, double* fdt, const double* d, const double* d1, int nlev) {
    const cplx* f = (const cplx*)field; cplx* t = (cplx*)fdt;
    for (int k = 0; k < nlev; k++)
        for (int q = 0; q < mx * nx; q++) t[(size_t)k * mx * nx + q] = (t[(size_t)k * mx * nx + q] - d[q] * f[(size_t)k * mx * nx + q]) * d1[q];
    return 0;
}
int orc_step(int j1, int j2, double dt, int csw) { compute_shortwave = csw != 0; step(j1, j2, dt); return 0; }
int orc_set_sppt(int on) { sppt_on = on != 0; if (on) sppt_reset(); return 0; }
int orc_get_geopotential(const double* tt, const double* phis_, double* phi_) {
    get_geopotential((const cplx*)tt, (const cplx*)phis_, (cplx*)phi_); return 0; }
/* get_tendencies on the resident state */
int orc_get_tendencies(int j2, int csw, double* vordt, double* divdt, double* tdt, double* psdt, double* trdt) {
    static Spec3 a, b, c; static Spec2 d; static FA<cplx, mx, nx, kx, ntr> e;
    compute_shortwave = csw != 0;
    get_tendencies(a, b, c, d, e, j2);
    memcpy(vordt, a.p(), sizeof(cplx) * a.size()); memcpy(divdt, b.p(), sizeof(cplx) * b.size());
    memcpy(tdt, c.p(), sizeof(cplx) * c.size()); memcpy(psdt, d.p(), sizeof(cplx) * d.size());
    memcpy(trdt, e.p(), sizeof(cplx) * e.size());
    return 0;
}
/* physics.f90:43 on caller-supplied spectral inputs and grid tendencies (in/out) */
int orc_get_physical_tendencies(const double* vor_, const double* divv, const double* tt, const double* q, const double* phi_, const double* psl,
                                double* utend, double* vtend, double* ttend, double* qtend, int csw) {
    static Grid3 u, v, w, x;
    compute_shortwave = csw != 0;
    memcpy(u.p(), utend, sizeof(double) * u.size()); memcpy(v.p(), vtend, sizeof(double) * u.size());
    memcpy(w.p(), ttend, sizeof(double) * u.size()); memcpy(x.p(), qtend, sizeof(double) * u.size());
    get_physical_tendencies((const cplx*)vor_, (const cplx*)divv, (const cplx*)tt, (const cplx*)q, (const cplx*)phi_, (const cplx*)psl, u, v, w, x);
    memcpy(utend, u.p(), sizeof(double) * u.size()); memcpy(vtend, v.p(), sizeof(double) * u.size());
    memcpy(ttend, w.p(), sizeof(double) * u.size()); memcpy(qtend, x.p(), sizeof(double) * u.size());
    return 0;
}
int orc_check_diagnostics(int level, double* diag) {
    return check_diagnostics(vor.p(1, 1, 1, level), div_.p(1, 1, 1, level), t.p(1, 1, 1, level), model_step, diag, false);
}
/* input_output.f90:184-206: float32 u,v,t,q,phi (ix,il,kx), ps (ix,il) from time level 1 */
int orc_output_fields(float* u, float* v, float* tt, float* q, float* ph, float* pso) {
    static Spec2 ucos, vcos; static Grid2 g;
    const size_t N = (size_t)ix * il;
    for (int k = 1; k <= kx; k++) {
        uvspec(vor.p(1, 1, k, 1), div_.p(1, 1, k, 1), ucos.p(), vcos.p());
        spec_to_grid(ucos.p(), 2, g.p()); for (size_t i = 0; i < N; i++) u[N * (k - 1) + i] = (float)g.d[i];
        spec_to_grid(vcos.p(), 2, g.p()); for (size_t i = 0; i < N; i++) v[N * (k - 1) + i] = (float)g.d[i];
        spec_to_grid(t.p(1, 1, k, 1), 1, g.p()); for (size_t i = 0; i < N; i++) tt[N * (k - 1) + i] = (float)g.d[i];
        spec_to_grid(tr.p(1, 1, k, 1, 1), 1, g.p()); for (size_t i = 0; i < N; i++) q[N * (k - 1) + i] = (float)(g.d[i] * (double)1.0e-3f);
        spec_to_grid(phi.p(1, 1, k), 1, g.p()); for (size_t i = 0; i < N; i++) ph[N * (k - 1) + i] = (float)(g.d[i] / grav);
    }
    spec_to_grid(ps.p(1, 1, 1), 1, g.p()); for (size_t i = 0; i < N; i++) pso[i] = (float)(p0 * exp(g.d[i]));
    return 0;
}


/* calendar-only hooks for the CPU tests (date.f90) */
int orc_calendar_init(int y, int m, int d, int h, int mi) { test_calendar_init(y, m, d, h, mi); return 0; }
int orc_newdate() { test_newdate(); return 0; }
int orc_get_date(int* ymdhm, double* tm, double* ty, int* im) {
    ymdhm[0] = model_datetime.year; ymdhm[1] = model_datetime.month; ymdhm[2] = model_datetime.day;
    ymdhm[3] = model_datetime.hour; ymdhm[4] = model_datetime.minute; *tm = tmonth; *ty = tyear; *im = imont1;
    return 0;
}

}  // extern "C"
