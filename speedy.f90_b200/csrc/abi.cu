// C ABI (include/speedy_b200.h): life cycle, tables, transforms and spectral operators.
// The model-level entry points live in model.cu.
#include "../../include/speedy_b200.h"
#include "ctx.h"
#include "model.h"
#include "abi_util.h"
#include <cstring>

using namespace spd;

namespace spd {
std::string& last_error() { static thread_local std::string e; return e; }

static const double* up(speedy_ctx* ctx, const char* name, const std::vector<double>& v) {
    auto& b = ctx->dtab[name];
    b.upload(v);
    return b.p;
}

void upload_implicit(speedy_ctx* ctx) {
    ImplicitTables& p = ctx->tab.imp;
    ctx->dv.dmp1 = up(ctx, "dmp1", p.dmp1);
    ctx->dv.dmp1d = up(ctx, "dmp1d", p.dmp1d);
    ctx->dv.dmp1s = up(ctx, "dmp1s", p.dmp1s);
    ctx->dv.elz = up(ctx, "elz", p.elz);
    ctx->dv.xj = up(ctx, "xj", p.xj);
    {   // xj(k,k1,l) transposed to [k + kx*k1][l]: the spectral-step kernel reads one l per lane (conflict-free shared memory)
        const int nl = (int)(p.xj.size() / 64);
        std::vector<double> t(p.xj.size());
        for (int l = 0; l < nl; l++) for (int e = 0; e < 64; e++) t[(size_t)e * nl + l] = p.xj[(size_t)l * 64 + e];
        ctx->dv.xjt = up(ctx, "xjt", t);
    }
    ctx->dv.xc = up(ctx, "xc", p.xc);
    ctx->dv.xd = up(ctx, "xd", p.xd);
    upload_level_consts(ctx);
}

void upload_tables(speedy_ctx* ctx) {
    Tables& t = ctx->tab;
    DevTables& v = ctx->dv;
    v.trunc = t.d.trunc; v.ix = t.d.ix; v.iy = t.d.iy; v.il = t.d.il; v.kx = t.d.kx; v.nx = t.d.nx; v.mx = t.d.mx;
    v.poly = up(ctx, "poly", t.poly);
    {   // device-only copy of P grouped by zonal wavenumber: one contiguous tile per CTA of the streaming direct transform
        const int ng = polyd_groups(t.d.trunc), mg = polyd_mg(t.d.trunc), iy = t.d.iy, nx = t.d.nx, mx = t.d.mx;
        std::vector<double> pd((size_t)ng * iy * nx * mg, 0.0);
        for (int gq = 0; gq < ng; gq++) for (int j = 0; j < iy; j++) for (int n = 0; n < nx; n++) for (int ml = 0; ml < mg; ml++) {
            const int m = gq * mg + ml;
            if (m < mx) pd[(((size_t)gq * iy + j) * nx + n) * mg + ml] = t.poly[((size_t)j * nx + n) * mx + m];
        }
        v.polyd = up(ctx, "polyd", pd);
        // packed triangle per latitude for the streaming inverse transform: row n holds m = 0..min(mx-1, mx-n)
        const int tr = polyt_row(t.d.trunc);
        std::vector<double> pt((size_t)iy * tr, 0.0);
        for (int j = 0; j < iy; j++) {
            size_t o = (size_t)j * tr;
            for (int n = 0; n < nx; n++) {
                const int cnt = std::min(mx, mx - n + 1);
                for (int m = 0; m < cnt; m++) pt[o + m] = t.poly[((size_t)j * nx + n) * mx + m];
                o += cnt;
            }
        }
        v.polyt = up(ctx, "polyt", pt);
    }
    if (t.d.trunc == 30) {   // quad kernels (transforms_quad.cu)
        std::vector<int> tiles; std::vector<double> pq;
        build_quad_tables(t, tiles, pq);
        v.polyq = up(ctx, "polyq", pq);
        ctx->d_qtile.upload(tiles);
        v.qtile = ctx->d_qtile.p;
        build_quad_inverse_tables(t, tiles, pq);
        v.polyi = up(ctx, "polyi", pq);
        ctx->d_qtile_inv.upload(tiles);
        v.qtile_inv = ctx->d_qtile_inv.p;
    }
    v.finv = up(ctx, "finv", t.finv);
    v.ffwd = up(ctx, "ffwd", t.ffwd);
    v.fftwa = up(ctx, "fftwa", t.fft_work);
    v.wt = up(ctx, "wt", t.wt);
    v.cosgr = up(ctx, "cosgr", t.cosgr);
    v.cosgr2 = up(ctx, "cosgr2", t.cosgr2);
    v.coriol = up(ctx, "coriol", t.coriol);
    v.cosg = up(ctx, "cosg", t.cosg);
    v.sia = up(ctx, "sia", t.sia);
    v.coa = up(ctx, "coa", t.coa);
    v.el2 = up(ctx, "el2", t.el2);
    v.elm2 = up(ctx, "elm2", t.elm2);
    v.trfilt = up(ctx, "trfilt", t.trfilt);
    v.gradx = up(ctx, "gradx", t.gradx);
    v.gradym = up(ctx, "gradym", t.gradym);
    v.gradyp = up(ctx, "gradyp", t.gradyp);
    v.uvdx = up(ctx, "uvdx", t.uvdx);
    v.uvdym = up(ctx, "uvdym", t.uvdym);
    v.uvdyp = up(ctx, "uvdyp", t.uvdyp);
    v.vddym = up(ctx, "vddym", t.vddym);
    v.vddyp = up(ctx, "vddyp", t.vddyp);
    v.dmp = up(ctx, "dmp", t.dmp);
    v.dmpd = up(ctx, "dmpd", t.dmpd);
    v.dmps = up(ctx, "dmps", t.dmps);
    v.fband = up(ctx, "fband", t.fband);
    upload_implicit(ctx);
}

// identity descriptors (field b at offset b*stride), cached per (nbatch, stride, flags)
static const XDesc* make_desc(speedy_ctx* ctx, int nbatch, size_t stride, const int* flags_host, int flag_if_set, bool kcos_semantics) {
    std::vector<XDesc> h(nbatch);
    unsigned long long key = 1469598103934665603ull;
    auto mix = [&](unsigned long long v) { key ^= v; key *= 1099511628211ull; };
    mix((unsigned long long)nbatch); mix((unsigned long long)stride);
    for (int b = 0; b < nbatch; b++) {
        h[b].off = (long long)b * (long long)stride;
        int f = 0;
        if (flags_host) f = kcos_semantics ? ((flags_host[b] != 1) ? 1 : 0) : flags_host[b];
        else f = flag_if_set;
        h[b].flags = f;
        h[b].op = 0;
        h[b].off2 = 0;
        mix((unsigned long long)f + 7);
    }
    auto it = ctx->desc_cache.find(key);
    if (it != ctx->desc_cache.end()) return it->second.p;
    auto& buf = ctx->desc_cache[key];
    buf.alloc(nbatch);
    CUDA_CHECK(cudaMemcpyAsync(buf.p, h.data(), sizeof(XDesc) * nbatch, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return buf.p;
}

static void h2d(speedy_ctx* ctx, double* d, const double* h, size_t n) {
    CUDA_CHECK(cudaMemcpyAsync(d, h, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
}
static void d2h(speedy_ctx* ctx, double* h, const double* d, size_t n) {
    CUDA_CHECK(cudaMemcpyAsync(h, d, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}
}  // namespace spd

extern "C" {

const char* speedy_last_error(void) { return last_error().c_str(); }

int speedy_create(const speedy_cfg* cfg, speedy_ctx** out) {
    API_BEGIN
    if (!cfg || !out) throw std::runtime_error("null argument");
    if (cfg->kx != 8 || cfg->ntr != 1) throw std::runtime_error("kx must be 8 and ntr 1 (params.f90:23,26)");
    if (cfg->nmembers < 1) throw std::runtime_error("nmembers must be >= 1");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw std::runtime_error("no CUDA device: speedy_b200 has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) throw std::runtime_error("bad device ordinal");
    CUDA_CHECK(cudaSetDevice(cfg->device));
    speedy_ctx* ctx = new speedy_ctx();
    try {
        build_tables(cfg->trunc, ctx->tab, cfg->nsteps ? cfg->nsteps : 36);
        ctx->d = ctx->tab.d;
        ctx->nmembers = cfg->nmembers;
        ctx->device = cfg->device;
        ctx->sppt_on = cfg->sppt_on;
        ctx->seed = cfg->seed;
        ctx->member_offset = cfg->member_offset;
        ctx->precision = cfg->precision;
        ctx->trace_pdl = getenv("SPEEDY_TRACE_PDL") != nullptr;
        ctx->fft_inverse = getenv("SPEEDY_DENSE_INVERSE") == nullptr;
        if (const char* v = getenv("SPEEDY_K2_FIELD")) ctx->k2_field = atoi(v) > 0 ? atoi(v) : 1;
        if (const char* v = getenv("SPEEDY_K2_QUAD")) ctx->k2_quad = atoi(v) != 0;
        if (const char* v = getenv("SPEEDY_K1_QUAD")) ctx->k1_quad = atoi(v) != 0;
        if (const char* v = getenv("SPEEDY_MEMBER_READY")) ctx->member_ready = atoi(v) != 0;
        if (const char* v = getenv("SPEEDY_L2_DISCARD")) ctx->l2_discard = atoi(v) != 0;
        if (const char* v = getenv("SPEEDY_TRANSIENT_ALIAS")) ctx->transient_alias = atoi(v) != 0;
        if (const char* v = getenv("SPEEDY_SPPT_FOLD")) ctx->sppt_fold = atoi(v) != 0;
        if (cfg->precision != 0 && cfg->precision != 1) throw std::runtime_error("precision must be 0 (fp64) or 1 (real32 transforms)");
        if (cfg->member_offset < 0 || cfg->member_offset + cfg->nmembers > 65536) throw std::runtime_error("member_offset + nmembers must stay within 65536");
        CUDA_CHECK(cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, cfg->device));
        CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreateWithFlags(&ctx->copy_event, cudaEventDisableTiming));
        setup_transform_kernels();
        setup_quad_kernels();
        setup_f32_kernels();
        setup_column_kernels();
        setup_spec_step_kernels();
        upload_tables(ctx);
        model_create(ctx);
    } catch (...) { delete ctx; throw; }
    *out = ctx;
    API_END
}

int speedy_destroy(speedy_ctx* ctx) {
    API_BEGIN
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    model_destroy(ctx);
    cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->copy_event) cudaEventDestroy(ctx->copy_event);
    delete ctx;
    API_END
}

int speedy_synchronize(speedy_ctx* ctx) {
    API_BEGIN
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    API_END
}

int speedy_dims(const speedy_ctx* ctx, int* dims) {
    API_BEGIN
    const Dims& d = ctx->d;
    dims[0] = d.trunc; dims[1] = d.ix; dims[2] = d.iy; dims[3] = d.il; dims[4] = d.kx; dims[5] = d.nx; dims[6] = d.mx; dims[7] = d.ntr;
    API_END
}

int speedy_run_info(const speedy_ctx* ctx, int* info) {
    API_BEGIN
    if (!ctx || !info) throw std::runtime_error("null argument");
    info[0] = ctx->nmembers; info[1] = ctx->tab.c.nsteps; info[2] = ctx->sppt_on; info[3] = ctx->precision;
    API_END
}

int speedy_get_table(const speedy_ctx* ctx, const char* name, double* out, size_t n) {
    API_BEGIN
    auto m = const_cast<speedy_ctx*>(ctx)->tab.named();
    auto it = m.find(name);
    if (it == m.end()) throw std::runtime_error(std::string("unknown table ") + name);
    if (it->second->size() != n) throw std::runtime_error(std::string("table ") + name + ": size mismatch, have " + std::to_string(it->second->size()));
    memcpy(out, it->second->data(), n * sizeof(double));
    API_END
}

static Tables& host_tables(int trunc) {
    static std::map<int, Tables> cache;
    auto it = cache.find(trunc);
    if (it == cache.end()) { build_tables(trunc, cache[trunc]); build_implicit(cache[trunc], 2.0 * cache[trunc].c.delt); it = cache.find(trunc); }
    return it->second;
}
long long speedy_host_table_len(int trunc, const char* name) {
    try {
        auto m = host_tables(trunc).named();
        auto it = m.find(name);
        return it == m.end() ? -1 : (long long)it->second->size();
    } catch (const std::exception& e_) { spd::last_error() = e_.what(); return -1; }
}
int speedy_host_table(int trunc, const char* name, double* out, size_t n) {
    API_BEGIN
    auto m = host_tables(trunc).named();
    auto it = m.find(name);
    if (it == m.end()) throw std::runtime_error(std::string("unknown table ") + name);
    if (it->second->size() != n) throw std::runtime_error(std::string("table ") + name + ": size mismatch, have " + std::to_string(it->second->size()));
    memcpy(out, it->second->data(), n * sizeof(double));
    API_END
}

// host-only: the start-up boundary fields (boundaries.f90:28-68, land_model.f90:50-181, sea_model.f90:80-250) from either source
long long speedy_host_boundary(const char* bc_path, int trunc, const char* name, double* out, size_t n) {
    try {
        if (!bc_path || !name) throw std::runtime_error("null argument");
        HostEnv env;
        load_host_env(bc_path, host_tables(trunc), env);
        const std::string s = name;
        std::vector<double> ssta_d;
        const std::vector<double>* v = nullptr;
        if (s == "phi0") v = &env.phi0; else if (s == "fmask") v = &env.fmask; else if (s == "alb0") v = &env.alb0;
        else if (s == "fmask_l") v = &env.fmask_l; else if (s == "bmask_l") v = &env.bmask_l; else if (s == "stl12") v = &env.stl12;
        else if (s == "snowd12") v = &env.snowd12; else if (s == "soilw12") v = &env.soilw12; else if (s == "rhcapl") v = &env.rhcapl;
        else if (s == "cdland") v = &env.cdland; else if (s == "fmask_s") v = &env.fmask_s; else if (s == "bmask_s") v = &env.bmask_s;
        else if (s == "sst12") v = &env.sst12; else if (s == "sice12") v = &env.sice12; else if (s == "rhcaps") v = &env.rhcaps;
        else if (s == "rhcapi") v = &env.rhcapi; else if (s == "cdsea") v = &env.cdsea; else if (s == "cdice") v = &env.cdice;
        else if (s == "solar") v = &env.solar;
        else if (s == "ssta") { ssta_d.assign(env.ssta.begin(), env.ssta.end()); v = &ssta_d; }
        else throw std::runtime_error("unknown boundary field " + s);
        if (out) {
            if (n > v->size()) throw std::runtime_error("boundary field " + s + " holds " + std::to_string(v->size()) + " values");
            memcpy(out, v->data(), n * sizeof(double));       // the first n values (ssta: the leading months)
        }
        return (long long)v->size();
    } catch (const std::exception& e_) { spd::last_error() = e_.what(); return -1; }
}

int speedy_set_table(speedy_ctx* ctx, const char* name, const double* in, size_t n) {
    API_BEGIN
    auto m = ctx->tab.named();
    auto it = m.find(name);
    if (it == m.end()) throw std::runtime_error(std::string("unknown table ") + name);
    if (it->second->size() != n) throw std::runtime_error(std::string("table ") + name + ": size mismatch");
    memcpy(it->second->data(), in, n * sizeof(double));
    if (std::string(name) == "cpol") {   // keep the unique-P layout in sync with a caller-supplied cpol
        const Dims& d = ctx->d;
        for (int j = 0; j < d.iy; j++) for (int nn = 0; nn < d.nx; nn++) for (int mm = 0; mm < d.mx; mm++)
            ctx->tab.poly[((size_t)j * d.nx + nn) * d.mx + mm] = ctx->tab.cpol[(2 * mm) + (size_t)2 * d.mx * (nn + (size_t)d.nx * j)];
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    upload_tables(ctx);
    API_END
}

// ---- transforms ----------------------------------------------------------------------
static int xform_host(speedy_ctx* ctx, bool inverse, int mode, const double* in, size_t in_len, int nbatch,
                      const int* kcos, double* out, size_t out_len) {
    API_BEGIN
    if (nbatch < 0) throw std::runtime_error("nbatch < 0");
    if (nbatch == 0) return 0;
    CUDA_CHECK(cudaSetDevice(ctx->device));
    ctx->ensure_scratch(ctx->scratch_a, in_len * nbatch);
    ctx->ensure_scratch(ctx->scratch_b, out_len * nbatch);
    const XDesc* dd = make_desc(ctx, nbatch, in_len, kcos, 0, true);
    h2d(ctx, ctx->scratch_a.p, in, in_len * nbatch);
    if (inverse) launch_spec_to_grid(ctx, ctx->scratch_a.p, 0, dd, nbatch, ctx->scratch_b.p, 0, 1, mode, nullptr, true);
    else launch_grid_to_spec(ctx, ctx->scratch_a.p, 0, dd, nbatch, ctx->scratch_b.p, 0, 1, mode);
    d2h(ctx, out, ctx->scratch_b.p, out_len * nbatch);
    API_END
}

int speedy_spec_to_grid(speedy_ctx* ctx, const double* spec, int nbatch, const int* kcos, double* grid) {
    return xform_host(ctx, true, 0, spec, (size_t)2 * ctx->d.nspec(), nbatch, kcos, grid, ctx->d.ngrid());
}
int speedy_grid_to_spec(speedy_ctx* ctx, const double* grid, int nbatch, double* spec) {
    return xform_host(ctx, false, 0, grid, ctx->d.ngrid(), nbatch, nullptr, spec, (size_t)2 * ctx->d.nspec());
}
int speedy_legendre_inv(speedy_ctx* ctx, const double* in, int nbatch, double* out) {
    return xform_host(ctx, true, 1, in, (size_t)2 * ctx->d.nspec(), nbatch, nullptr, out, (size_t)ctx->d.k2() * ctx->d.il);
}
int speedy_legendre_dir(speedy_ctx* ctx, const double* in, int nbatch, double* out) {
    return xform_host(ctx, false, 2, in, (size_t)ctx->d.k2() * ctx->d.il, nbatch, nullptr, out, (size_t)2 * ctx->d.nspec());
}
int speedy_fourier_inv(speedy_ctx* ctx, const double* in, int nbatch, const int* kcos, double* out) {
    return xform_host(ctx, true, 2, in, (size_t)ctx->d.k2() * ctx->d.il, nbatch, kcos, out, ctx->d.ngrid());
}
int speedy_fourier_dir(speedy_ctx* ctx, const double* in, int nbatch, double* out) {
    return xform_host(ctx, false, 1, in, ctx->d.ngrid(), nbatch, nullptr, out, (size_t)ctx->d.k2() * ctx->d.il);
}

int speedy_spec_to_grid_dev(speedy_ctx* ctx, const double* d_spec, int nbatch, const int* kcos, double* d_grid) {
    API_BEGIN
    const XDesc* dd = make_desc(ctx, nbatch, (size_t)2 * ctx->d.nspec(), kcos, 0, true);
    launch_spec_to_grid(ctx, d_spec, 0, dd, nbatch, d_grid, 0, 1, 0, nullptr, true);
    API_END
}
int speedy_grid_to_spec_dev(speedy_ctx* ctx, const double* d_grid, int nbatch, double* d_spec) {
    API_BEGIN
    const XDesc* dd = make_desc(ctx, nbatch, ctx->d.ngrid(), nullptr, 0, false);
    launch_grid_to_spec(ctx, d_grid, 0, dd, nbatch, d_spec, 0, 1, 0);
    API_END
}

// ---- spectral operators ----------------------------------------------------------------
static int specop_host(speedy_ctx* ctx, int op, const double* a, const double* b, int nbatch, double* o1, double* o2) {
    API_BEGIN
    if (nbatch <= 0) return 0;
    CUDA_CHECK(cudaSetDevice(ctx->device));
    const size_t len = (size_t)2 * ctx->d.nspec() * nbatch;
    ctx->ensure_scratch(ctx->scratch_a, len);
    ctx->ensure_scratch(ctx->scratch_b, len);
    ctx->ensure_scratch(ctx->scratch_c, len);
    ctx->ensure_scratch(ctx->scratch_d, len);
    h2d(ctx, ctx->scratch_a.p, a, len);
    if (b) h2d(ctx, ctx->scratch_b.p, b, len);
    launch_spectral_op(ctx, op, ctx->scratch_a.p, ctx->scratch_b.p, ctx->scratch_c.p, ctx->scratch_d.p, nbatch);
    d2h(ctx, o1, ctx->scratch_c.p, len);
    if (o2) d2h(ctx, o2, ctx->scratch_d.p, len);
    API_END
}
int speedy_laplacian(speedy_ctx* ctx, const double* in, int nbatch, double* out) { return specop_host(ctx, OP_LAPLACIAN, in, nullptr, nbatch, out, nullptr); }
int speedy_inverse_laplacian(speedy_ctx* ctx, const double* in, int nbatch, double* out) { return specop_host(ctx, OP_INVLAPLACIAN, in, nullptr, nbatch, out, nullptr); }
int speedy_grad(speedy_ctx* ctx, const double* psi, int nbatch, double* psdx, double* psdy) { return specop_host(ctx, OP_GRAD, psi, nullptr, nbatch, psdx, psdy); }
int speedy_vds(speedy_ctx* ctx, const double* u, const double* v, int nbatch, double* vorm, double* divm) { return specop_host(ctx, OP_VDS, u, v, nbatch, vorm, divm); }
int speedy_uvspec(speedy_ctx* ctx, const double* vorm, const double* divm, int nbatch, double* u, double* v) { return specop_host(ctx, OP_UVSPEC, vorm, divm, nbatch, u, v); }
int speedy_trunct(speedy_ctx* ctx, double* vor, int nbatch) { return specop_host(ctx, OP_TRUNCT, vor, nullptr, nbatch, vor, nullptr); }

int speedy_vdspec(speedy_ctx* ctx, const double* ug, const double* vg, int nbatch, int kcos, double* vorm, double* divm) {
    API_BEGIN
    if (nbatch <= 0) return 0;
    CUDA_CHECK(cudaSetDevice(ctx->device));
    const size_t glen = (size_t)ctx->d.ngrid() * nbatch, slen = (size_t)2 * ctx->d.nspec() * nbatch;
    ctx->ensure_scratch(ctx->scratch_a, 2 * glen);
    ctx->ensure_scratch(ctx->scratch_b, 2 * slen);
    ctx->ensure_scratch(ctx->scratch_c, slen);
    ctx->ensure_scratch(ctx->scratch_d, slen);
    h2d(ctx, ctx->scratch_a.p, ug, glen);
    h2d(ctx, ctx->scratch_a.p + glen, vg, glen);
    std::vector<int> fl(2 * nbatch, kcos == 2 ? 1 : 2);   // spectral.f90:208-222
    const XDesc* dd = make_desc(ctx, 2 * nbatch, ctx->d.ngrid(), fl.data(), 0, false);
    launch_grid_to_spec(ctx, ctx->scratch_a.p, 0, dd, 2 * nbatch, ctx->scratch_b.p, 0, 1, 0);
    launch_spectral_op(ctx, OP_VDS, ctx->scratch_b.p, ctx->scratch_b.p + slen, ctx->scratch_c.p, ctx->scratch_d.p, nbatch);
    d2h(ctx, vorm, ctx->scratch_c.p, slen);
    d2h(ctx, divm, ctx->scratch_d.p, slen);
    API_END
}

long long speedy_launch_count(const speedy_ctx* ctx) { return ctx ? ctx->launches : 0; }
void* speedy_stream(const speedy_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int speedy_set_graphs(speedy_ctx* ctx, int on) { ctx->use_graphs = on != 0; return 0; }

}  // extern "C"
