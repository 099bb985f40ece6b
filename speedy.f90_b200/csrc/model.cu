#include <algorithm>
// Model-level entry points of the C ABI: device-resident state, start-up
// (initialization.f90:12-82), the time step (time_stepping.f90:12-122), the main-loop body
// (speedy.f90:27-54) replayed as a CUDA graph, the surface slabs, diagnostics and output.
#include "../../include/speedy_b200.h"
#include "model.h"
#include "calendar.h"
#include "close_step.cuh"
#include "abi_util.h"
#include "host/netcdf_classic.h"
#include <cstdio>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <cmath>
#include <cstring>

using namespace spd;

namespace spd {

static const int KXc = 8;

void calendar_init(DevClock& c, int y, int m, int d, int h, int mi, int nssta, int nsteps) {
    memset(&c, 0, sizeof(c));
    c.year = y; c.month = m; c.day = d; c.hour = h; c.minute = mi;
    c.start_year = y;
    c.model_step = 1;
    c.nssta = nssta;
    c.nsteps = nsteps;
    cal_fractions(c);
    cal_step_flags(c);
}
void calendar_advance(DevClock& c) { cal_advance(c); }

void upload_level_consts(speedy_ctx* ctx) {
    if (!ctx->model) return;
    const Tables& t = ctx->tab;
    LevelConsts h;
    memset(&h, 0, sizeof(h));
    for (int k = 0; k < 9; k++) { h.hsg[k] = t.hsg[k]; h.sigh[k] = t.sigh[k]; }
    for (int k = 0; k < 8; k++) {
        h.dhs[k] = t.dhs[k]; h.fsg[k] = t.fsg[k]; h.dhsr[k] = t.dhsr[k]; h.fsgr[k] = t.fsgr[k];
        h.tref[k] = t.imp.tref[k]; h.tref1[k] = t.imp.tref1[k]; h.tref2[k] = t.imp.tref2[k]; h.tref3[k] = t.imp.tref3[k];
        h.dhsx[k] = t.imp.dhsx[k];
        h.xgeop1[k] = t.xgeop1[k]; h.xgeop2[k] = t.xgeop2[k]; h.geop_corf[k] = t.geop_corf[k];
        h.tcorv[k] = t.tcorv[k]; h.qcorv[k] = t.qcorv[k];
        h.sigl[k] = t.sigl[k]; h.grdsig[k] = t.grdsig[k]; h.grdscp[k] = t.grdscp[k];
    }
    for (int k = 0; k < 16; k++) h.wvi[k] = t.wvi[k];
    const Consts& c = t.c;
    h.rgas = c.rgas; h.akap = c.akap; h.cp = c.cp; h.p0 = c.p0; h.grav = c.grav; h.alhc = c.alhc; h.alhs = c.alhs;
    h.sbc = c.sbc; h.rearth = c.rearth; h.refrh1 = c.refrh1; h.gamma = c.gamma;
    h.rob = c.rob; h.wil = c.wil;
    h.sdrag = 1.0 / (c.tdrs * 3600.0);   // time_stepping.f90:77
    {   // convection.f90:118-131 entrainment profile and mass-flux constants; large_scale_condensation.f90:52; shortwave_radiation.f90:227
        const double psmin = (double)0.8f, trcnv = 6.0, entmax = 0.5, epslw = (double)0.05f;
        double sentr = 0.0;
        for (int k = 2; k <= 7; k++) {
            const double d = h.fsg[k - 1] - 0.5;
            const double ee = 0.0 > d ? 0.0 : d;
            h.entr[k - 1] = ee * ee;
            sentr = sentr + h.entr[k - 1];
        }
        sentr = entmax / sentr;
        for (int k = 2; k <= 7; k++) h.entr[k - 1] = h.entr[k - 1] * sentr;
        h.ralhc = 1.0 / c.alhc;
        h.fm0 = c.p0 * h.dhs[7] / (c.grav * trcnv * 3600.0);
        h.rdps = 2.0 / (1.0 - psmin);
        h.prg = c.p0 / c.grav;
        h.tfact = c.alhc / c.cp;
        h.eps1 = epslw / (h.dhs[0] + h.dhs[1]);
        h.rcp = 1.0 / c.cp;
    }
    Model& M = *ctx->model;
    if (!M.lc.p) M.lc.alloc(1);
    CUDA_CHECK(cudaMemcpyAsync(M.lc.p, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

static void push_clock(speedy_ctx* ctx) {
    Model& M = *ctx->model;
    CUDA_CHECK(cudaMemcpyAsync(M.clock.p, &M.hclock, sizeof(DevClock), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}
static void pull_clock(speedy_ctx* ctx) {
    Model& M = *ctx->model;
    CUDA_CHECK(cudaMemcpyAsync(&M.hclock, M.clock.p, sizeof(DevClock), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

static void drop_graph(Model& M) {
    if (M.day_graph) { cudaGraphExecDestroy(M.day_graph); M.day_graph = nullptr; M.day_graph_steps = 0; }
}

void model_create(speedy_ctx* ctx) {
    Model* Mp = new Model();
    ctx->model = Mp;
    Model& M = *Mp;
    const Dims& d = ctx->d;
    const long long NS2 = 2LL * d.nspec(), NG = d.ngrid();
    Layout& L = M.L;
    long long off = 0;
    auto reg = [&](const char* name, long long& slot, long long len) {
        slot = off;
        M.fields[name] = FieldInfo{off, (size_t)len, false};
        off += len;
    };
    // spectral state, Fortran layouts (prognostics.f90:16-24)
    reg("vor", L.vor, 2 * KXc * NS2); reg("div", L.div, 2 * KXc * NS2); reg("t", L.t, 2 * KXc * NS2); reg("tr", L.tr, 2 * KXc * NS2);
    reg("ps", L.ps, 2 * NS2); reg("phi", L.phi, KXc * NS2); reg("phi_next", L.phi_next, KXc * NS2); reg("phis", L.phis, NS2);
    reg("tcorh", L.tcorh, NS2); reg("qcorh", L.qcorh, NS2);
    reg("sout", L.sout, GO_N * NS2);
    reg("sppt_spec", L.sppt_spec, KXc * NS2); reg("sppt_eta", L.sppt_eta, KXc * NS2);
    reg("vordt", L.vordt, KXc * NS2); reg("divdt", L.divdt, KXc * NS2); reg("tdt", L.tdt, KXc * NS2); reg("trdt", L.trdt, KXc * NS2);
    reg("psdt", L.psdt, NS2);
    reg("gin", L.gin, GI_N * NG); reg("gout", L.gout, GO_N * NG);
    reg("phis0", L.phis0, NG); reg("fmask_l", L.fmask_l, NG); reg("fmask_s", L.fmask_s, NG); reg("forog", L.forog, NG); reg("alb0", L.alb0, NG);
    reg("fsol", L.fsol, NG); reg("ozone", L.ozone, NG); reg("ozupp", L.ozupp, NG); reg("zenit", L.zenit, NG); reg("stratz", L.stratz, NG);
    reg("alb_l", L.alb_l, NG); reg("alb_s", L.alb_s, NG); reg("albsfc", L.albsfc, NG); reg("snowc", L.snowc, NG);
    reg("stl_am", L.stl_am, NG); reg("stl_lm", L.stl_lm, NG); reg("snowd_am", L.snowd_am, NG); reg("soilw_am", L.soilw_am, NG);
    reg("sst_am", L.sst_am, NG); reg("sice_am", L.sice_am, NG); reg("tice_am", L.tice_am, NG); reg("ssti_om", L.ssti_om, NG);
    reg("sst_om", L.sst_om, NG); reg("tice_om", L.tice_om, NG); reg("sice_om", L.sice_om, NG);
    reg("sstcl_ob", L.sstcl_ob, NG); reg("sicecl_ob", L.sicecl_ob, NG); reg("ticecl_ob", L.ticecl_ob, NG); reg("stlcl_ob", L.stlcl_ob, NG);
    reg("sstan3", L.sstan3, 3 * NG);
    reg("tau2", L.tau2, 4 * KXc * NG); reg("stratc", L.stratc, 2 * NG); reg("tt_rsw", L.tt_rsw, KXc * NG);
    reg("ssrd", L.ssrd, NG); reg("ssr", L.ssr, NG); reg("tsr", L.tsr, NG);
    reg("precnv", L.precnv, NG); reg("precls", L.precls, NG); reg("cbmf", L.cbmf, NG); reg("slrd", L.slrd, NG); reg("slr", L.slr, NG); reg("olr", L.olr, NG);
    reg("slru", L.slru, 3 * NG); reg("ustr", L.ustr, 3 * NG); reg("vstr", L.vstr, 3 * NG); reg("shf", L.shf, 3 * NG); reg("evap", L.evap, 3 * NG);
    reg("hfluxn", L.hfluxn, 3 * NG);
    reg("ts", L.ts, NG); reg("tskin", L.tskin, NG); reg("u0", L.u0, NG); reg("v0", L.v0, NG); reg("t0", L.t0, NG);
    reg("qcloud", L.qcloud, NG); reg("cloudc", L.cloudc, NG); reg("clstr", L.clstr, NG);
    L.qcorh_g = L.gout + (long long)GO_QCORH * NG;     // the daily humidity-correction field is the 74th K2 input
    M.fields["qcorh_g"] = FieldInfo{L.qcorh_g, (size_t)NG, false};
    L.stride = (off + 15) / 16 * 16;
    long long ioff = 0;
    auto ireg = [&](const char* name, long long& slot, long long len) {
        slot = ioff;
        M.fields[name] = FieldInfo{ioff, (size_t)len, true};
        ioff += len;
    };
    ireg("iptop", L.iptop, NG); ireg("icltop", L.icltop, NG); ireg("icnv", L.icnv, NG);
    L.istride = ioff;
    M.mem.alloc((size_t)L.stride * ctx->nmembers);
    M.imem.alloc((size_t)L.istride * ctx->nmembers);
    M.clock.alloc(1);
    calendar_init(M.hclock, 1982, 1, 1, 0, 0, 0, ctx->tab.c.nsteps);
    CUDA_CHECK(cudaMemcpy(M.clock.p, &M.hclock, sizeof(DevClock), cudaMemcpyHostToDevice));
    upload_level_consts(ctx);
    {
        const size_t nb = (d.nspec() + 31) / 32;
        M.diag_partial.alloc(nb * ctx->nmembers * 2 * KXc + (size_t)ctx->nmembers * KXc);
        M.ready.alloc(ctx->nmembers);
    }

    // transform descriptors --------------------------------------------------------------
    {
        std::vector<XDesc> h(2 * GI_N);
        for (int j2 = 1; j2 <= 2; j2++) {
            XDesc* D = h.data() + (size_t)(j2 - 1) * GI_N;
            const long long lev = (long long)(j2 - 1) * KXc * NS2;
            for (int k = 0; k < KXc; k++) {
                D[GI_VOR + k] = XDesc{L.vor + lev + k * NS2, 0, 0};
                D[GI_DIV + k] = XDesc{L.div + lev + k * NS2, 0, 0};
                D[GI_T + k] = XDesc{L.t + lev + k * NS2, 0, 0};
                D[GI_TR + k] = XDesc{L.tr + lev + k * NS2, 0, 0};
                // winds: K1 builds ucos/vcos from (vor, div) in its input stage (uvspec) and scales by cosgr (spec_to_grid(., 2))
                D[GI_U + k] = XDesc{L.vor + lev + k * NS2, 1, 1, L.div + lev + k * NS2};
                D[GI_V + k] = XDesc{L.vor + lev + k * NS2, 1, 2, L.div + lev + k * NS2};
                D[GI_U1 + k] = XDesc{L.vor + k * NS2, 1, 1, L.div + k * NS2};
                D[GI_V1 + k] = XDesc{L.vor + k * NS2, 1, 2, L.div + k * NS2};
                D[GI_T1 + k] = XDesc{L.t + k * NS2, 0, 0, 0};
                D[GI_Q1 + k] = XDesc{L.tr + k * NS2, 0, 0, 0};
                D[GI_PHI + k] = XDesc{L.phi_next + k * NS2, 0, 0, 0};
                D[GI_SPPT + k] = XDesc{L.sppt_spec + k * NS2, 0, 0, 0};
            }
            D[GI_PX] = XDesc{L.ps + (long long)(j2 - 1) * NS2, 1, 3, 0};     // grad(ps(:,:,j2)) tendencies.f90:121-123
            D[GI_PY] = XDesc{L.ps + (long long)(j2 - 1) * NS2, 1, 4, 0};
            D[GI_PSL] = XDesc{L.ps, 0, 0, 0};
        }
        M.desc_inv.upload(h);
        // per-step list: the physics uses the level-1 wind at the lowest level only (surface fluxes), so the
        // uvspec + transform of u,v at levels 1..kx-1 (physics.f90:95-96) is dead work and is not enqueued
        {
            // order: the derived fields first, as PAIRS sharing their sources — (ucos, vcos) of a level, (d/dx, d/dy) of ps — so that the
            // quad kernel finds a pair in slots (0, 1) or (2, 3) of a quad and evaluates it in place; then the plain fields
            std::vector<XDesc> cs;
            const int nf = ctx->sppt_on ? GI_N : GI_NBASE;
            for (int j2 = 1; j2 <= 2; j2++) {
                std::vector<int> order;
                for (int k = 0; k < KXc; k++) { order.push_back(GI_U + k); order.push_back(GI_V + k); }
                order.push_back(GI_PX); order.push_back(GI_PY);
                order.push_back(GI_U1 + KXc - 1); order.push_back(GI_V1 + KXc - 1);
                std::vector<char> used(GI_N, 0);
                for (int f : order) used[f] = 1;
                for (int f = 0; f < nf; f++) {
                    const bool dead = (f >= GI_U1 && f < GI_U1 + KXc - 1) || (f >= GI_V1 && f < GI_V1 + KXc - 1);
                    if (!dead && !used[f]) order.push_back(f);
                }
                for (int f : order) {
                    XDesc x = h[(size_t)(j2 - 1) * GI_N + f];
                    x.oslot1 = f + 1;
                    cs.push_back(x);
                }
            }
            M.nstep_fields = (int)cs.size() / 2;
            M.desc_step.upload(cs);
        }
        // output(): the 41 level-1 fields with the module variable phi (input_output.f90:184-192)
        std::vector<XDesc> o(h.begin() + GI_U1, h.begin() + GI_U1 + 41);
        for (int k = 0; k < KXc; k++) o[GI_PHI - GI_U1 + k].off = L.phi + k * NS2;
        M.desc_out.upload(o);
        std::vector<XDesc> g(GO_N);
        for (int f = 0; f < GO_N; f++) {
            const int r = f % GO_PER;
            const bool cosgr = f < GO_PSDT && (r == 0 || r == 1 || r == 3 || r == 4 || r == 6 || r == 7);   // vdspec(.,.,2): spectral.f90:208-213
            g[f] = XDesc{f * NG, cosgr ? 1 : 0, 0};      // relative to the gout region (whole rows of the K2 input tensor map)
        }
        g[GO_QCORH].flags = 4;      // gated on the device clock's do_forcing
        M.desc_dir.upload(g);
        std::vector<XDesc> one(1, XDesc{(long long)GO_QCORH * NG, 0, 0});
        M.desc_one_dir.upload(one);
    }
}

void outpipe_destroy(Model& M);
void model_destroy(speedy_ctx* ctx) {
    if (!ctx->model) return;
    outpipe_destroy(*ctx->model);
    drop_graph(*ctx->model);
    free_column_maps(*ctx->model);
    delete ctx->model;
    ctx->model = nullptr;
}

// ---- launch sequences -------------------------------------------------------------------
static void xform_inverse(speedy_ctx* ctx, int j2, int first, int count) {
    Model& M = *ctx->model;
    const long long NG = ctx->d.ngrid();
    launch_spec_to_grid(ctx, M.mem.p, M.L.stride, M.desc_inv.p + (size_t)(j2 - 1) * GI_N + first, count,
                        M.mem.p + M.L.gin + first * NG, M.L.stride, ctx->nmembers, 0);
}
// the inverse transforms of one time step (tendencies.f90:89-123 + physics.f90:95-104), compact list
// with_close: an extra CTA of the kernel closes the previous main-loop step if one is pending (close_step.cuh)
static void xform_step(speedy_ctx* ctx, int j2, bool with_close = false) {
    Model& M = *ctx->model;
    CloseArgs cl{M.clock.p, M.diag_partial.p, (int)((ctx->d.nspec() + 31) / 32), ctx->nmembers, nullptr, nullptr};
    M.ready_target = 0;
    if (with_close && ctx->member_ready && s2g_quad_selected(ctx, M.nstep_fields, ctx->nmembers, true)) {
        cl.ready = M.ready.p;
        M.ready_target = s2g_quad_ready_counts(M.nstep_fields) + 1;     // + the closing CTA
    }
    launch_spec_to_grid(ctx, M.mem.p, M.L.stride, M.desc_step.p + (size_t)(j2 - 1) * M.nstep_fields, M.nstep_fields,
                        M.mem.p + M.L.gin, M.L.stride, ctx->nmembers, 0, with_close ? &cl : nullptr, true);
}
static void xform_output(speedy_ctx* ctx) {
    Model& M = *ctx->model;
    launch_spec_to_grid(ctx, M.mem.p, M.L.stride, M.desc_out.p, 41, M.mem.p + M.L.gin + (long long)GI_U1 * ctx->d.ngrid(), M.L.stride, ctx->nmembers, 0);
}
static void xform_direct(speedy_ctx* ctx, bool with_daily_qcorh = false) {
    Model& M = *ctx->model;
    const bool alias = with_daily_qcorh && M.alias_active;
    G2sStepOpts step;
    step.transient_input = with_daily_qcorh;        // main-loop step: the column kernel's output is read here and nowhere else
    step.out_field_stride = alias ? ctx->d.ngrid() : 0;
    launch_grid_to_spec(ctx, M.mem.p + (alias ? M.L.gin : M.L.gout), M.L.stride, M.desc_dir.p, with_daily_qcorh ? GO_N : GO_QCORH,
                        M.mem.p + (alias ? M.L.gin : M.L.sout), M.L.stride, ctx->nmembers, 0, with_daily_qcorh ? &M.clock.p->do_forcing : nullptr, step);
}
static void xform_qcorh(speedy_ctx* ctx, bool gated) {
    Model& M = *ctx->model;
    (void)gated;
    launch_grid_to_spec(ctx, M.mem.p + M.L.gout, M.L.stride, M.desc_one_dir.p, 1, M.mem.p + M.L.qcorh, M.L.stride, ctx->nmembers, 0, nullptr);
}

// gen_sppt of a get_tendencies call (sppt.f90:45): the pattern is already there when the last spectral step prepared it in its
// prologue (dynamics.cu), else the stand-alone kernel draws it now.  The host flag follows the order of the enqueued work.
static void sppt_next(speedy_ctx* ctx) {
    if (!ctx->sppt_on) return;
    Model& M = *ctx->model;
    if (M.sppt_prepared) M.sppt_prepared = false;
    else launch_sppt_update(ctx);
}
// get_tendencies up to (and including) the direct transforms
static void enqueue_tendency_front(speedy_ctx* ctx, int j2, int csw_override) {
    launch_geopotential(ctx, 3);
    sppt_next(ctx);
    xform_step(ctx, j2);
    launch_grid_columns(ctx, 0, csw_override);
    xform_direct(ctx);
}
// step(j1,j2,dt)  time_stepping.f90:35-122
static void enqueue_step(speedy_ctx* ctx, int j1, int j2, double dt, int csw_override) {
    enqueue_tendency_front(ctx, j2, csw_override);
    launch_spec_step(ctx, j1, j2, dt, 0);
}
// main-loop body speedy.f90:27-54 in 4 launches.  The column kernel first applies the pending
// couple_sea_land of the previous step and, when due, set_forcing(1); the spectral-step kernel
// ends with the check_diagnostics partial sums; the step is closed (final reduction, range guard, calendar) by a spare CTA
// of the next step's spec->grid kernel or by k_close_step.
static const int kLaunchesPerStep = 4;
static void enqueue_main_loop_step(speedy_ctx* ctx) {
    const double delt = ctx->tab.c.delt;
    // trace mode (exact timeline) closes each step with the stand-alone kernel
    const bool tracing = ctx->dv.trace != nullptr;
    sppt_next(ctx);
    ctx->model->alias_active = ctx->transient_alias && ctx->l2_discard && g2s_quad_selected(ctx, GO_N, ctx->nmembers) &&
                               s2g_quad_selected(ctx, ctx->model->nstep_fields, ctx->nmembers, true);
    xform_step(ctx, 2, !tracing);
    launch_grid_columns(ctx, 0, -1, 1);
    xform_direct(ctx, true);
    launch_spec_step(ctx, 2, 2, 2 * delt, 0, 1);
    ctx->model->ready_target = 0;
    ctx->model->alias_active = false;
    if (tracing) launch_close_step(ctx);
}
// the coupler call of the last step (speedy.f90:53) when no further step follows in this call
static void flush_pending_slab(speedy_ctx* ctx) {
    Model& M = *ctx->model;
    launch_close_step(ctx);      // diagnostics + calendar of the last step (no-op if already closed)
    launch_slab(ctx, 0);     // also clears slab_pending
}

static void set_implicit(speedy_ctx* ctx, double dt) {
    Model& M = *ctx->model;
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    build_implicit(ctx->tab, dt);
    upload_implicit(ctx);      // also refreshes the level constants
    M.implicit_dt = dt;
}

// keeps_phi_next: only the device-resident main loop leaves phi_next valid; every other entry point may change the state
static void check_ready(speedy_ctx* ctx, bool keeps_phi_next = false) {
    if (!ctx || !ctx->model) throw std::runtime_error("null context");
    CUDA_CHECK(cudaSetDevice(ctx->device));
    if (!keeps_phi_next) ctx->model->phi_next_valid = false;
}

static void set_all_members(speedy_ctx* ctx, long long off, const double* host, size_t len) {
    Model& M = *ctx->model;
    for (int e = 0; e < ctx->nmembers; e++)
        CUDA_CHECK(cudaMemcpyAsync(M.mem.p + (size_t)e * M.L.stride + off, host, len * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

}  // namespace spd

extern "C" {

size_t speedy_output_len(const speedy_ctx* ctx) { return (size_t)(5 * ctx->d.kx + 1) * ctx->d.ngrid(); }
size_t speedy_state_len(const speedy_ctx* ctx) { return (size_t)2 * ctx->d.nspec() * (4 * 2 * ctx->d.kx + 2); }

const char* speedy_field_names(void) {
    return "vor div t tr ps phi phi_next phis tcorh qcorh vordt divdt tdt trdt psdt gin gout sout sppt_spec sppt_eta "
           "phis0 fmask_l fmask_s forog alb0 fsol ozone ozupp zenit stratz alb_l alb_s albsfc snowc "
           "stl_am stl_lm snowd_am soilw_am sst_am sice_am tice_am ssti_om sst_om tice_om sice_om "
           "sstcl_ob sicecl_ob ticecl_ob stlcl_ob sstan3 tau2 stratc tt_rsw ssrd ssr tsr precnv precls cbmf slrd slr olr "
           "slru ustr vstr shf evap hfluxn ts tskin u0 v0 t0 qcloud cloudc clstr qcorh_g iptop icltop icnv";
}

static const FieldInfo& find_field(speedy_ctx* ctx, const char* name, bool want_int) {
    auto it = ctx->model->fields.find(name);
    if (it == ctx->model->fields.end()) throw std::runtime_error(std::string("unknown field ") + name);
    if (it->second.is_int != want_int) throw std::runtime_error(std::string("field ") + name + (want_int ? " is not an integer field" : " is an integer field"));
    return it->second;
}

// n == len: broadcast to every member; n == nmembers*len: one copy per member
int speedy_set_field(speedy_ctx* ctx, const char* name, const double* host, size_t n) {
    API_BEGIN
    check_ready(ctx);
    const FieldInfo& f = find_field(ctx, name, false);
    Model& M = *ctx->model;
    if (n == f.len) set_all_members(ctx, f.off, host, f.len);
    else if (n == f.len * (size_t)ctx->nmembers) {
        for (int e = 0; e < ctx->nmembers; e++)
            CUDA_CHECK(cudaMemcpyAsync(M.mem.p + (size_t)e * M.L.stride + f.off, host + (size_t)e * f.len, f.len * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    } else throw std::runtime_error(std::string("field ") + name + ": expected " + std::to_string(f.len) + " (or nmembers x that) values");
    API_END
}
// n == len: member 0; n == nmembers*len: every member
int speedy_get_field(speedy_ctx* ctx, const char* name, double* host, size_t n) {
    API_BEGIN
    check_ready(ctx);
    const FieldInfo& f = find_field(ctx, name, false);
    Model& M = *ctx->model;
    if (n != f.len && n != f.len * (size_t)ctx->nmembers)
        throw std::runtime_error(std::string("field ") + name + ": expected " + std::to_string(f.len) + " (or nmembers x that) values");
    const int ne = (int)(n / f.len);
    for (int e = 0; e < ne; e++)
        CUDA_CHECK(cudaMemcpyAsync(host + (size_t)e * f.len, M.mem.p + (size_t)e * M.L.stride + f.off, f.len * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    API_END
}
int speedy_get_ifield(speedy_ctx* ctx, const char* name, int* host, size_t n) {
    API_BEGIN
    check_ready(ctx);
    const FieldInfo& f = find_field(ctx, name, true);
    Model& M = *ctx->model;
    if (n != f.len && n != f.len * (size_t)ctx->nmembers) throw std::runtime_error(std::string("field ") + name + ": size mismatch");
    const int ne = (int)(n / f.len);
    for (int e = 0; e < ne; e++)
        CUDA_CHECK(cudaMemcpyAsync(host + (size_t)e * f.len, M.imem.p + (size_t)e * M.L.istride + f.off, f.len * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    API_END
}

int speedy_initialize_implicit(speedy_ctx* ctx, double dt) {
    API_BEGIN
    check_ready(ctx);
    set_implicit(ctx, dt);
    API_END
}

// implicit.f90:168-217 on the caller's tendencies (host arrays, in place), with the matrices of the last speedy_initialize_implicit
int speedy_implicit_terms(speedy_ctx* ctx, double* divdt, double* tdt, double* psdt) {
    API_BEGIN
    check_ready(ctx);
    if (!divdt || !tdt || !psdt) throw std::runtime_error("speedy_implicit_terms: null argument");
    if (ctx->model->implicit_dt == 0.0) throw std::runtime_error("speedy_implicit_terms: call speedy_initialize_implicit(dt) first (implicit.f90:36)");
    const size_t n2 = (size_t)2 * ctx->d.nspec(), n3 = n2 * KXc;
    ctx->ensure_scratch(ctx->scratch_a, n3);
    ctx->ensure_scratch(ctx->scratch_b, n3);
    ctx->ensure_scratch(ctx->scratch_c, n2);
    CUDA_CHECK(cudaMemcpyAsync(ctx->scratch_a.p, divdt, n3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(ctx->scratch_b.p, tdt, n3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(ctx->scratch_c.p, psdt, n2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    launch_implicit_terms(ctx, ctx->scratch_a.p, ctx->scratch_b.p, ctx->scratch_c.p);
    CUDA_CHECK(cudaMemcpyAsync(divdt, ctx->scratch_a.p, n3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(tdt, ctx->scratch_b.p, n3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(psdt, ctx->scratch_c.p, n2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    API_END
}

// horizontal_diffusion.f90:86-105: fdt = (fdt - dmp*field)*dmp1 on nlev levels (1: the 2-D form, kx: the 3-D form), host arrays
int speedy_do_horizontal_diffusion(speedy_ctx* ctx, const double* field, double* fdt, const double* dmp, const double* dmp1, int nlev) {
    API_BEGIN
    check_ready(ctx);
    if (!field || !fdt || !dmp || !dmp1) throw std::runtime_error("speedy_do_horizontal_diffusion: null argument");
    if (nlev < 1) throw std::runtime_error("speedy_do_horizontal_diffusion: nlev must be positive");
    const size_t ns = (size_t)ctx->d.nspec(), n3 = 2 * ns * nlev;
    ctx->ensure_scratch(ctx->scratch_a, n3);
    ctx->ensure_scratch(ctx->scratch_b, n3);
    ctx->ensure_scratch(ctx->scratch_c, 2 * ns);
    CUDA_CHECK(cudaMemcpyAsync(ctx->scratch_a.p, field, n3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(ctx->scratch_b.p, fdt, n3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(ctx->scratch_c.p, dmp, ns * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(ctx->scratch_c.p + ns, dmp1, ns * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    launch_horizontal_diffusion(ctx, ctx->scratch_a.p, ctx->scratch_b.p, ctx->scratch_c.p, ctx->scratch_c.p + ns, nlev);
    CUDA_CHECK(cudaMemcpyAsync(fdt, ctx->scratch_b.p, n3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    API_END
}

int speedy_get_geopotential(speedy_ctx* ctx, int j) {
    API_BEGIN
    check_ready(ctx);
    if (j != 1) throw std::runtime_error("get_geopotential: the resident phi is tied to time level 1 (tendencies.f90:203,288)");
    launch_geopotential(ctx, 3);
    API_END
}

int speedy_get_tendencies(speedy_ctx* ctx, int j2, int compute_shortwave) {
    API_BEGIN
    check_ready(ctx);
    if (j2 != 1 && j2 != 2) throw std::runtime_error("j2 must be 1 or 2");
    enqueue_tendency_front(ctx, j2, compute_shortwave ? 1 : 0);
    launch_spec_step(ctx, 1, j2, 0.0, 1);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    API_END
}

int speedy_get_physical_tendencies(speedy_ctx* ctx, const double* vor, const double* div, const double* t, const double* q,
                                   const double* phi, const double* psl, double* utend, double* vtend, double* ttend, double* qtend,
                                   int compute_shortwave) {
    API_BEGIN
    check_ready(ctx);
    if (ctx->nmembers != 1) throw std::runtime_error("host-array physics entry point needs a single-member context");
    Model& M = *ctx->model;
    const Layout& L = M.L;
    const size_t NS2 = (size_t)2 * ctx->d.nspec(), NG = ctx->d.ngrid();
    auto up = [&](long long off, const double* h, size_t n) { CUDA_CHECK(cudaMemcpyAsync(M.mem.p + off, h, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream)); };
    // the arguments are the caller's time-level-1 arrays (tendencies.f90:205): they become the resident level 1
    up(L.vor, vor, KXc * NS2); up(L.div, div, KXc * NS2); up(L.t, t, KXc * NS2); up(L.tr, q, KXc * NS2); up(L.phi, phi, KXc * NS2); up(L.ps, psl, NS2);
    up(L.phi_next, phi, KXc * NS2);
    std::vector<double> g((size_t)GO_N * NG, 0.0);
    for (int k = 0; k < KXc; k++) {
        memcpy(&g[(size_t)(GO_PER * k + 0) * NG], utend + k * NG, NG * sizeof(double));
        memcpy(&g[(size_t)(GO_PER * k + 1) * NG], vtend + k * NG, NG * sizeof(double));
        memcpy(&g[(size_t)(GO_PER * k + 5) * NG], ttend + k * NG, NG * sizeof(double));
        memcpy(&g[(size_t)(GO_PER * k + 8) * NG], qtend + k * NG, NG * sizeof(double));
    }
    up(L.gout, g.data(), g.size());
    if (ctx->sppt_on) { sppt_next(ctx); xform_inverse(ctx, 1, GI_SPPT, 8); }
    xform_inverse(ctx, 1, GI_U1, 41);
    launch_grid_columns(ctx, 1, compute_shortwave ? 1 : 0);
    CUDA_CHECK(cudaMemcpyAsync(g.data(), M.mem.p + L.gout, g.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < KXc; k++) {
        memcpy(utend + k * NG, &g[(size_t)(GO_PER * k + 0) * NG], NG * sizeof(double));
        memcpy(vtend + k * NG, &g[(size_t)(GO_PER * k + 1) * NG], NG * sizeof(double));
        memcpy(ttend + k * NG, &g[(size_t)(GO_PER * k + 5) * NG], NG * sizeof(double));
        memcpy(qtend + k * NG, &g[(size_t)(GO_PER * k + 8) * NG], NG * sizeof(double));
    }
    API_END
}

int speedy_step(speedy_ctx* ctx, int j1, int j2, double dt, int compute_shortwave) {
    API_BEGIN
    check_ready(ctx);
    if ((j1 != 1 && j1 != 2) || (j2 != 1 && j2 != 2)) throw std::runtime_error("j1, j2 must be 1 or 2");
    enqueue_step(ctx, j1, j2, dt, compute_shortwave < 0 ? -1 : (compute_shortwave ? 1 : 0));
    API_END
}

int speedy_first_step(speedy_ctx* ctx) {   // time_stepping.f90:12-24; compute_shortwave = .true. (shortwave_radiation.f90:67)
    API_BEGIN
    check_ready(ctx);
    const double delt = ctx->tab.c.delt;
    set_implicit(ctx, 0.5 * delt);
    enqueue_step(ctx, 1, 1, 0.5 * delt, 1);
    set_implicit(ctx, delt);
    enqueue_step(ctx, 1, 2, delt, 1);
    set_implicit(ctx, 2 * delt);
    API_END
}

int speedy_step_host(speedy_ctx* ctx, double* state, size_t n, int j1, int j2, double dt, int compute_shortwave) {
    API_BEGIN
    check_ready(ctx);
    if (ctx->nmembers != 1) throw std::runtime_error("host-array step needs a single-member context");
    if (n != speedy_state_len(ctx)) throw std::runtime_error("state length mismatch (see speedy_state_len)");
    Model& M = *ctx->model;
    const Layout& L = M.L;
    const size_t NS2 = (size_t)2 * ctx->d.nspec(), n4 = 2 * KXc * NS2;
    const long long offs[5] = {L.vor, L.div, L.t, L.tr, L.ps};
    const size_t lens[5] = {n4, n4, n4, n4, 2 * NS2};
    size_t p = 0;
    for (int i = 0; i < 5; i++) { CUDA_CHECK(cudaMemcpyAsync(M.mem.p + offs[i], state + p, lens[i] * sizeof(double), cudaMemcpyHostToDevice, ctx->stream)); p += lens[i]; }
    enqueue_step(ctx, j1, j2, dt, compute_shortwave ? 1 : 0);
    p = 0;
    for (int i = 0; i < 5; i++) { CUDA_CHECK(cudaMemcpyAsync(state + p, M.mem.p + offs[i], lens[i] * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)); p += lens[i]; }
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    API_END
}

int speedy_couple_sea_land(speedy_ctx* ctx, int day) {
    API_BEGIN
    check_ready(ctx);
    launch_slab(ctx, day == 0 ? 1 : 0);
    API_END
}

int speedy_set_forcing(speedy_ctx* ctx, int imode) {
    API_BEGIN
    check_ready(ctx);
    (void)imode;   // the time-independent part of imode 0 (radset, forog, tcorh) is done by speedy_model_init
    launch_daily_forcing(ctx, 1);
    xform_qcorh(ctx, false);
    API_END
}

int speedy_check_diagnostics(speedy_ctx* ctx, int time_level, double* diag) {
    API_BEGIN
    check_ready(ctx);
    Model& M = *ctx->model;
    launch_diagnostics(ctx, time_level);
    pull_clock(ctx);
    if (diag) memcpy(diag, M.hclock.diag, sizeof(double) * 24);
    if (M.hclock.diag_fail) return 1;
    API_END
}

int speedy_range_failure(speedy_ctx* ctx, long long* step, double* diag) {
    API_BEGIN
    check_ready(ctx, true);
    pull_clock(ctx);
    const DevClock& c = ctx->model->hclock;
    if (step) *step = c.diag_fail;
    if (diag) memcpy(diag, c.diag, sizeof(double) * 24);
    if (c.diag_fail) return 1;
    API_END
}

int speedy_model_date(const speedy_ctx* cctx, int* ymdhm, long long* model_step) {
    API_BEGIN
    speedy_ctx* ctx = const_cast<speedy_ctx*>(cctx);
    check_ready(ctx);
    pull_clock(ctx);
    const DevClock& c = ctx->model->hclock;
    if (ymdhm) { ymdhm[0] = c.year; ymdhm[1] = c.month; ymdhm[2] = c.day; ymdhm[3] = c.hour; ymdhm[4] = c.minute; }
    if (model_step) *model_step = c.model_step;
    API_END
}

// initialization.f90:12-82
int speedy_model_init(speedy_ctx* ctx, const char* bc_path, int year, int month, int day, int hour, int minute) {
    API_BEGIN
    check_ready(ctx);
    Model& M = *ctx->model;
    drop_graph(M);
    const Dims& d = ctx->d;
    const Consts& c = ctx->tab.c;
    const int NG = d.ngrid(), NS = d.nspec();
    HostEnv env;
    load_host_env(bc_path, ctx->tab, env);
    // shared device tables
    {
        std::vector<double> sh;
        auto put = [&](const std::vector<double>& v) { size_t o = sh.size(); sh.insert(sh.end(), v.begin(), v.end()); return o; };
        const size_t o_stl = put(env.stl12), o_snd = put(env.snowd12), o_sw = put(env.soilw12), o_sst = put(env.sst12), o_sic = put(env.sice12);
        const size_t o_rl = put(env.rhcapl), o_cl = put(env.cdland), o_rs = put(env.rhcaps), o_ri = put(env.rhcapi), o_cs = put(env.cdsea), o_ci = put(env.cdice);
        const size_t o_bm = put(env.bmask_s), o_sol = put(env.solar);
        M.shared.upload(sh);
        M.ssta.upload(env.ssta);
        const double* b = M.shared.p;
        M.sh = SharedDev{b + o_stl, b + o_snd, b + o_sw, b + o_sst, b + o_sic, b + o_rl, b + o_cl, b + o_rs, b + o_ri, b + o_cs, b + o_ci, b + o_bm, M.ssta.p, b + o_sol};
    }
    // caller-supplied SPPT noise (speedy_set_sppt_draw(ctx, 0) + set_field("sppt_eta")) survives the reset
    std::vector<double> keep_eta;
    const size_t eta_len = (size_t)KXc * 2 * NS;
    if (!M.sppt_draw) {
        keep_eta.resize(eta_len * ctx->nmembers);
        for (int e = 0; e < ctx->nmembers; e++)
            CUDA_CHECK(cudaMemcpyAsync(keep_eta.data() + e * eta_len, M.mem.p + (size_t)e * M.L.stride + M.L.sppt_eta, eta_len * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    CUDA_CHECK(cudaMemsetAsync(M.mem.p, 0, M.mem.n * sizeof(double), ctx->stream));
    CUDA_CHECK(cudaMemsetAsync(M.imem.p, 0, M.imem.n * sizeof(int), ctx->stream));
    for (int e = 0; e < ctx->nmembers && !keep_eta.empty(); e++)
        CUDA_CHECK(cudaMemcpyAsync(M.mem.p + (size_t)e * M.L.stride + M.L.sppt_eta, keep_eta.data() + e * eta_len, eta_len * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    // date.f90:53-105, initialization.f90:37
    calendar_init(M.hclock, year, month, day, hour, minute, env.nssta, ctx->tab.c.nsteps);
    { const int st[5] = {year, month, day, hour, minute}; memcpy(M.start, st, sizeof st); }
    const int isst0 = (year - 1979) * 12 + month;
    if (isst0 - 1 < 1 || isst0 + 1 > env.nssta) throw std::runtime_error("start date outside the resident SST-anomaly window of the boundary file");
    push_clock(ctx);
    // boundaries.f90:28-43: spectrally truncated surface geopotential (through the transform kernels)
    std::vector<double> phis0(NG), spec((size_t)2 * NS), tmp(NG);
    if (speedy_grid_to_spec(ctx, env.phi0.data(), 1, spec.data())) throw std::runtime_error(speedy_last_error());
    for (int n = 0; n < d.nx; n++)
        for (int m = 0; m < d.mx; m++)
            if (m + n > d.trunc) { spec[2 * (m + (size_t)d.mx * n)] = 0.0; spec[2 * (m + (size_t)d.mx * n) + 1] = 0.0; }
    int kcos1 = 1;
    if (speedy_spec_to_grid(ctx, spec.data(), 1, &kcos1, phis0.data())) throw std::runtime_error(speedy_last_error());
    M.h_phis0 = phis0;
    set_all_members(ctx, M.L.phis0, phis0.data(), NG);
    set_all_members(ctx, M.L.fmask_l, env.fmask_l.data(), NG);
    set_all_members(ctx, M.L.fmask_s, env.fmask_s.data(), NG);
    set_all_members(ctx, M.L.alb0, env.alb0.data(), NG);
    {   // set_orog_land_sfc_drag  surface_fluxes.f90:300-309
        const double rhdrag = 1.0 / (c.grav * 2000.0);
        for (int q = 0; q < NG; q++) tmp[q] = 1.0 + rhdrag * (1.0 - exp(-std::max(phis0[q], 0.0) * rhdrag));
        set_all_members(ctx, M.L.forog, tmp.data(), NG);
    }
    // ---- prognostics.f90:34-127 rest state (time level 1; level 2 stays zero until the first step writes it)
    {
        const size_t NS2 = (size_t)2 * NS;
        std::vector<double> phis(NS2), surfs(NS2), st((size_t)2 * KXc * NS2, 0.0), ps2(2 * NS2, 0.0), trs((size_t)2 * KXc * NS2, 0.0);
        if (speedy_grid_to_spec(ctx, phis0.data(), 1, phis.data())) throw std::runtime_error(speedy_last_error());
        set_all_members(ctx, M.L.phis, phis.data(), NS2);
        const double gam1 = c.gamma / (1000.0 * c.grav);
        const double tref = 288.0, ttop = 216.0, gam2 = gam1 / tref, rgam = c.rgas * gam1, rgamr = 1.0 / rgam;
        for (size_t i = 0; i < NS2; i++) surfs[i] = -gam1 * phis[i];
        const double sq2 = (double)sqrtf(2.0f);
        st[0] = sq2 * ttop; st[1] = 0.0 * ttop;
        st[NS2 + 0] = sq2 * ttop; st[NS2 + 1] = 0.0 * ttop;
        surfs[0] = sq2 * tref - gam1 * phis[0];
        surfs[1] = 0.0 - gam1 * phis[1];
        for (int k = 2; k < KXc; k++) {
            const double f = pow(ctx->tab.fsg[k], rgam);
            for (size_t i = 0; i < NS2; i++) st[(size_t)k * NS2 + i] = surfs[i] * f;
        }
        set_all_members(ctx, M.L.t, st.data(), st.size());
        const double rlog0 = (double)logf(1.013f);
        std::vector<double> surfg(NG);
        for (int q = 0; q < NG; q++) surfg[q] = rlog0 + rgamr * log(1.0 - gam2 * phis0[q]);
        if (speedy_grid_to_spec(ctx, surfg.data(), 1, ps2.data())) throw std::runtime_error(speedy_last_error());
        for (int n = 0; n < d.nx; n++)
            for (int m = 0; m < d.mx; m++)
                if (m + n > d.trunc) { ps2[2 * (m + (size_t)d.mx * n)] = 0.0; ps2[2 * (m + (size_t)d.mx * n) + 1] = 0.0; }
        set_all_members(ctx, M.L.ps, ps2.data(), ps2.size());
        const double esref = 17.0, qref = c.refrh1 * (double)0.622f * esref, qexp = c.hscale / c.hshum;
        for (int q = 0; q < NG; q++) surfg[q] = qref * exp(qexp * surfg[q]);
        if (speedy_grid_to_spec(ctx, surfg.data(), 1, surfs.data())) throw std::runtime_error(speedy_last_error());
        for (int n = 0; n < d.nx; n++)
            for (int m = 0; m < d.mx; m++)
                if (m + n > d.trunc) { surfs[2 * (m + (size_t)d.mx * n)] = 0.0; surfs[2 * (m + (size_t)d.mx * n) + 1] = 0.0; }
        for (int k = 2; k < KXc; k++) {
            const double f = pow(ctx->tab.fsg[k], qexp);
            for (size_t i = 0; i < NS2; i++) trs[(size_t)k * NS2 + i] = surfs[i] * f;
        }
        set_all_members(ctx, M.L.tr, trs.data(), trs.size());
        // tcorh = grid_to_spec(gamlat*phis0) (forcing.f90:77-82) does not depend on the date
        for (int q = 0; q < NG; q++) tmp[q] = gam1 * phis0[q];
        if (speedy_grid_to_spec(ctx, tmp.data(), 1, surfs.data())) throw std::runtime_error(speedy_last_error());
        set_all_members(ctx, M.L.tcorh, surfs.data(), NS2);
    }
    // ---- initialize_coupler (coupler.f90:15-30): sstan3 months isst0-1..isst0+1, then the day-0 coupling
    {
        std::vector<double> an((size_t)3 * NG);
        for (int s = 0; s < 3; s++)
            for (int q = 0; q < NG; q++) an[(size_t)s * NG + q] = (double)env.ssta[(size_t)(isst0 - 2 + s) * NG + q];
        set_all_members(ctx, M.L.sstan3, an.data(), an.size());
    }
    launch_slab(ctx, 1);
    // ---- set_forcing(0) (forcing.f90:15-100); fband/forog/tcorh are already resident
    launch_daily_forcing(ctx, 1);
    xform_qcorh(ctx, false);
    // geopotential of the rest state (prognostics.f90:123) so that an immediate output sees it
    launch_geopotential(ctx, 3);
    if (M.sppt_state.n != 2) M.sppt_state.alloc(2);
    CUDA_CHECK(cudaMemsetAsync(M.sppt_state.p, 0, 2 * sizeof(int), ctx->stream));   // gen_sppt's `first` (sppt.f90:52)
    M.sppt_prepared = false;
    // ---- first_step (time_stepping.f90:12-24)
    if (speedy_first_step(ctx)) throw std::runtime_error(speedy_last_error());
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    M.initialized = true;
    API_END
}

// speedy.f90:27-54 repeated nsteps times; whole days are replayed from a CUDA graph.  Enqueue only.
static void run_steps_core(speedy_ctx* ctx, int nsteps) {
    Model& M = *ctx->model;
    if (!M.initialized) throw std::runtime_error("speedy_model_init has not been called");
    if (M.implicit_dt != 2 * ctx->tab.c.delt) set_implicit(ctx, 2 * ctx->tab.c.delt);
    int left = nsteps;
    const int G = ctx->tab.c.nsteps;      // one simulated day per graph
    // phi_next of the current level-1 T: every main-loop step leaves it up to date; it is recomputed only after another
    // entry point of the library ran in between (which may have changed the state)
    if (nsteps > 0 && !M.phi_next_valid) launch_geopotential(ctx, 2);
    // SPPT with device-drawn noise: every spectral step prepares the next step's pattern, so a captured day holds no stand-alone update —
    // which is right only if the day starts with a prepared pattern: the first step after start-up runs outside the graph
    const bool folding = ctx->sppt_on && M.sppt_draw && ctx->sppt_fold;
    if (folding && !M.sppt_prepared && left > 0) { enqueue_main_loop_step(ctx); left--; }
    if (ctx->use_graphs && left >= G) {
        if (!M.day_graph) {
            cudaGraph_t graph;
            const long long before = ctx->launches;
            CUDA_CHECK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
            try {
                for (int s = 0; s < G; s++) enqueue_main_loop_step(ctx);
            } catch (...) { cudaGraph_t g2; cudaStreamEndCapture(ctx->stream, &g2); throw; }
            CUDA_CHECK(cudaStreamEndCapture(ctx->stream, &graph));
            ctx->launches = before;
            CUDA_CHECK(cudaGraphInstantiate(&M.day_graph, graph, 0));
            CUDA_CHECK(cudaGraphDestroy(graph));
            M.day_graph_steps = G;
        }
        while (left >= G) {
            CUDA_CHECK(cudaGraphLaunch(M.day_graph, ctx->stream));
            ctx->launches += (long long)G * (kLaunchesPerStep + ((ctx->sppt_on && !folding) ? 1 : 0));
            left -= G;
        }
    }
    for (int s = 0; s < left; s++) enqueue_main_loop_step(ctx);
    if (nsteps > 0) { flush_pending_slab(ctx); M.phi_next_valid = true; }
}
static int finish_run(speedy_ctx* ctx) {
    Model& M = *ctx->model;
    pull_clock(ctx);
    if (M.hclock.ssta_missing) throw std::runtime_error("run left the SST-anomaly window resident in the boundary file (pack more months)");
    return M.hclock.diag_fail ? 1 : 0;   // 1: 'Model variables out of accepted range' (diagnostics.f90:68)
}

int speedy_run_steps(speedy_ctx* ctx, int nsteps) {
    API_BEGIN
    check_ready(ctx, true);
    run_steps_core(ctx, nsteps);
    if (finish_run(ctx)) return 1;
    API_END
}

// The same loop without the host round trip: speedy_enqueue_steps only enqueues (graph replays + tail steps) and returns;
// speedy_finish drains the stream and reports the range guard like speedy_run_steps.  A caller that integrates many days back
// to back (bench.py, ensemble drivers) enqueues them all and polls once.
int speedy_enqueue_steps(speedy_ctx* ctx, int nsteps) {
    API_BEGIN
    check_ready(ctx, true);
    run_steps_core(ctx, nsteps);
    API_END
}
int speedy_finish(speedy_ctx* ctx) {
    API_BEGIN
    check_ready(ctx, true);
    if (finish_run(ctx)) return 1;
    API_END
}

// enqueue the output() conversions of one member into the context's device buffer
static float* enqueue_output(speedy_ctx* ctx, int member) {
    Model& M = *ctx->model;
    const size_t n = speedy_output_len(ctx);
    if (M.outbuf.n < n) M.outbuf.alloc(n);
    xform_output(ctx);                         // phi stays the one of the last step (input_output.f90:184-192)
    launch_output_convert(ctx, member, M.outbuf.p);
    return M.outbuf.p;
}

// Host-resident drop-in of the main loop: the caller keeps the prognostic arrays (vor, div, t, tr,
// ps as in prognostics.f90:16-20, concatenated, nmembers copies) in HOST memory; they are uploaded,
// advanced nsteps time steps and downloaded, and `out` (optional, speedy_output_len floats) receives
// member 0's output() fields.  One stream synchronisation at the end.
int speedy_run_steps_host(speedy_ctx* ctx, double* state, size_t n, int nsteps, float* out) {
    API_BEGIN
    check_ready(ctx);
    Model& M = *ctx->model;
    const size_t len = speedy_state_len(ctx);
    if (n != len * (size_t)ctx->nmembers) throw std::runtime_error("state length mismatch (nmembers x speedy_state_len)");
    for (int e = 0; e < ctx->nmembers; e++)
        CUDA_CHECK(cudaMemcpyAsync(M.mem.p + (size_t)e * M.L.stride + M.L.vor, state + (size_t)e * len, len * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    run_steps_core(ctx, nsteps);
    // the state goes home on a second stream while the output() transforms and conversions (which only read it) run
    CUDA_CHECK(cudaEventRecord(ctx->copy_event, ctx->stream));
    CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_event, 0));
    for (int e = 0; e < ctx->nmembers; e++)
        CUDA_CHECK(cudaMemcpyAsync(state + (size_t)e * len, M.mem.p + (size_t)e * M.L.stride + M.L.vor, len * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream));
    if (out) {
        float* d = enqueue_output(ctx, 0);
        CUDA_CHECK(cudaMemcpyAsync(out, d, speedy_output_len(ctx) * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    }
    int rc = 0;
    try { rc = finish_run(ctx); } catch (...) { cudaStreamSynchronize(ctx->copy_stream); throw; }   // the caller's buffer must not be in flight when an error is reported
    CUDA_CHECK(cudaStreamSynchronize(ctx->copy_stream));
    if (rc) return 1;
    API_END
}

int speedy_output_fields(speedy_ctx* ctx, int member, float* u, float* v, float* t, float* q, float* phi, float* ps) {
    API_BEGIN
    check_ready(ctx);
    if (member < 0 || member >= ctx->nmembers) throw std::runtime_error("bad member index");
    const size_t NG = ctx->d.ngrid(), n = speedy_output_len(ctx);
    float* d = enqueue_output(ctx, member);
    std::vector<float> h(n);
    CUDA_CHECK(cudaMemcpyAsync(h.data(), d, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    float* dst[5] = {u, v, t, q, phi};
    for (int f = 0; f < 5; f++)
        if (dst[f]) memcpy(dst[f], h.data() + (size_t)f * KXc * NG, KXc * NG * sizeof(float));
    if (ps) memcpy(ps, h.data() + (size_t)5 * KXc * NG, NG * sizeof(float));
    API_END
}

// ---- output files (input_output.f90:95-217) --------------------------------------------------------
// One NetCDF classic file per output time, named yyyymmddhhmm.nc after the model date, holding the float32
// u, v, t, q, phi (lon, lat, lev, time) and ps (lon, lat, time) of speedy_output_fields plus the coordinate
// variables and attributes the reference writes.  Host-only: no device work, callable without a GPU.
int speedy_write_output_file(const char* path, int trunc, int nsteps, const int* start_ymdhm, int timestep,
                             const float* u, const float* v, const float* t, const float* q, const float* phi, const float* ps) {
    API_BEGIN
    if (!path || !start_ymdhm || !u || !v || !t || !q || !phi || !ps) throw std::runtime_error("speedy_write_output_file: null argument");
    // coordinate values of a resolution: built once (the full table set costs ~1 ms, an output file is written every step by default)
    struct Coord { int ix, il, nsteps; std::vector<double> radang, fsg; };
    static std::mutex coord_mu;
    static std::map<std::pair<int, int>, Coord> coord_cache;
    Coord co;
    {
        std::lock_guard<std::mutex> lk(coord_mu);
        const auto key = std::make_pair(trunc, nsteps > 0 ? nsteps : 36);
        auto it = coord_cache.find(key);
        if (it == coord_cache.end()) {
            Tables tab;
            build_tables(trunc, tab, key.second);
            Coord c{tab.d.ix, tab.d.il, tab.c.nsteps, tab.radang, std::vector<double>(tab.fsg.begin(), tab.fsg.begin() + KXc)};
            it = coord_cache.emplace(key, std::move(c)).first;
        }
        co = it->second;
    }
    const int ix = co.ix, il = co.il;
    NcClassicWriter nc;
    char units[64];
    snprintf(units, sizeof units, "hours since %04d-%02d-%02d %02d:%02d:0.0", start_ymdhm[0], start_ymdhm[1], start_ymdhm[2], start_ymdhm[3], start_ymdhm[4]);
    // definition order as in the reference: time, lon, lat, lev, then the fields
    const int dt = nc.def_dim("time", 0);
    const int vt = nc.def_var("time", {dt});
    nc.put_att(vt, "units", units);
    const int dlon = nc.def_dim("lon", ix), dlat = nc.def_dim("lat", il), dlev = nc.def_dim("lev", KXc);
    const int vlon = nc.def_var("lon", {dlon}); nc.put_att(vlon, "long_name", "longitude");
    const int vlat = nc.def_var("lat", {dlat}); nc.put_att(vlat, "long_name", "latitude");
    const int vlev = nc.def_var("lev", {dlev}); nc.put_att(vlev, "long_name", "atmosphere_sigma_coordinate");
    struct F { const char *name, *long_name, *units; const float* data; };
    const F f3[5] = {{"u", "eastward_wind", "m/s", u}, {"v", "northward_wind", "m/s", v}, {"t", "air_temperature", "K", t},
                     {"q", "specific_humidity", "1", q}, {"phi", "geopotential_height", "m", phi}};
    int vf[5];
    for (int i = 0; i < 5; i++) {
        vf[i] = nc.def_var(f3[i].name, {dt, dlev, dlat, dlon});   // Fortran (lon, lat, lev, time)
        nc.put_att(vf[i], "long_name", f3[i].long_name);
        nc.put_att(vf[i], "units", f3[i].units);
    }
    const int vps = nc.def_var("ps", {dt, dlat, dlon});
    nc.put_att(vps, "long_name", "surface_air_pressure");
    nc.put_att(vps, "units", "Pa");
    // coordinate values, in the reference's mixed real32 / real64 arithmetic (input_output.f90:178-181)
    const float hours = (float)timestep * 24.0f / (float)co.nsteps;
    nc.put_var(vt, &hours, 1);
    std::vector<float> lon(ix), lat(il), lev(KXc);
    const float dlon_deg = (float)(360.0 / ix);                    // 3.75 at T30
    for (int k = 0; k < ix; k++) lon[k] = dlon_deg * (float)k;
    const double quarter = (double)asinf(1.0f);
    for (int k = 0; k < il; k++) lat[k] = (float)(co.radang[k] * 90.0 / quarter);
    for (int k = 0; k < KXc; k++) lev[k] = (float)co.fsg[k];
    nc.put_var(vlon, lon.data(), lon.size());
    nc.put_var(vlat, lat.data(), lat.size());
    nc.put_var(vlev, lev.data(), lev.size());
    const size_t NG = (size_t)ix * il;
    for (int i = 0; i < 5; i++) nc.put_var_ref(vf[i], f3[i].data, (size_t)KXc * NG);
    nc.put_var_ref(vps, ps, NG);
    nc.write(path);
    API_END
}

// output() of the resident state of one member: converts on the device (speedy_output_fields), names the file after
// the model date and writes it into `dir`; the path is returned in path_out (optional)
int speedy_write_output(speedy_ctx* ctx, int member, const char* dir, char* path_out, size_t path_cap) {
    API_BEGIN
    check_ready(ctx);
    if (member < 0 || member >= ctx->nmembers) throw std::runtime_error("bad member index");
    Model& M = *ctx->model;
    const size_t NG = ctx->d.ngrid(), n = speedy_output_len(ctx);
    float* d = enqueue_output(ctx, member);
    std::vector<float> h(n);
    CUDA_CHECK(cudaMemcpyAsync(h.data(), d, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    pull_clock(ctx);
    const DevClock& c = M.hclock;
    char name[32];
    snprintf(name, sizeof name, "%04d%02d%02d%02d%02d.nc", c.year, c.month, c.day, c.hour, c.minute);
    std::string path = (dir && *dir) ? std::string(dir) + "/" + name : std::string(name);
    const float* f = h.data();
    const size_t L3 = (size_t)KXc * NG;
    if (speedy_write_output_file(path.c_str(), ctx->d.trunc, ctx->tab.c.nsteps, M.start, c.model_step - 1, f, f + L3, f + 2 * L3, f + 3 * L3, f + 4 * L3, f + 5 * L3))
        throw std::runtime_error(speedy_last_error());
    if (path_out && path_cap) { strncpy(path_out, path.c_str(), path_cap - 1); path_out[path_cap - 1] = 0; }
    API_END
}

// ---- asynchronous output ------------------------------------------------------------
// With the reference's default nsteps_out = 1 a file is written after EVERY step; written synchronously that is ~0.8 ms of host work
// (byte swap + a 741 KB write) against a 26 us model step.  Here the conversions are enqueued on the context's stream like any kernel,
// the float32 fields go to one of NSLOT pinned host buffers, an event marks the copy, and host threads write the files while the
// device runs on.  The caller names the date and the step of the enqueued state (the host knows the calendar: no device round trip).
namespace {
struct OutPipe {
    static constexpr int NSLOT = 16;
    struct Job { int slot; std::string path; int trunc, nsteps, start[5]; long long timestep; size_t ng; };
    int device = 0;
    size_t n = 0;                          // floats per slot
    float* slot[NSLOT] = {};
    cudaEvent_t ev[NSLOT] = {};
    std::vector<int> free_slots;
    std::deque<Job> jobs;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::vector<std::thread> workers;
    int in_flight = 0;
    bool stop = false;
    std::string error;

    void start(int dev, size_t nfloats) {
        device = dev; n = nfloats;
        for (int i = 0; i < NSLOT; i++) {
            CUDA_CHECK(cudaMallocHost(&slot[i], n * sizeof(float)));
            CUDA_CHECK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
            free_slots.push_back(i);
        }
        unsigned hw = std::thread::hardware_concurrency();
        const int nw = (int)std::min(8u, std::max(2u, hw / 2));
        for (int w = 0; w < nw; w++) workers.emplace_back([this] { work(); });
    }
    void work() {
        cudaSetDevice(device);
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_job.wait(lk, [&] { return stop || !jobs.empty(); });
                if (jobs.empty()) return;
                j = std::move(jobs.front());
                jobs.pop_front();
            }
            std::string err;
            if (cudaEventSynchronize(ev[j.slot]) != cudaSuccess) err = "asynchronous output: the device-to-host copy failed";
            else {
                const float* f = slot[j.slot];
                const size_t L3 = (size_t)KXc * j.ng;
                if (speedy_write_output_file(j.path.c_str(), j.trunc, j.nsteps, j.start, (int)j.timestep, f, f + L3, f + 2 * L3, f + 3 * L3, f + 4 * L3, f + 5 * L3))
                    err = speedy_last_error();      // thread-local
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                if (!err.empty() && error.empty()) error = err;
                free_slots.push_back(j.slot);
                in_flight--;
            }
            cv_done.notify_all();
        }
    }
    int acquire() {                         // blocks while every slot is in flight
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return !free_slots.empty(); });
        const int s = free_slots.back();
        free_slots.pop_back();
        return s;
    }
    void release(int s) {                   // a slot that was acquired but never submitted
        { std::lock_guard<std::mutex> lk(mu); free_slots.push_back(s); }
        cv_done.notify_all();
    }
    void submit(Job j) {
        { std::lock_guard<std::mutex> lk(mu); jobs.push_back(std::move(j)); in_flight++; }
        cv_job.notify_one();
    }
    std::string drain() {
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return in_flight == 0; });
        std::string e;
        e.swap(error);
        return e;
    }
    ~OutPipe() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv_job.notify_all();
        for (auto& t : workers) t.join();
        for (int i = 0; i < NSLOT; i++) { if (slot[i]) cudaFreeHost(slot[i]); if (ev[i]) cudaEventDestroy(ev[i]); }
    }
};
}  // namespace
}  // extern "C"
void spd::outpipe_destroy(Model& M) { delete static_cast<OutPipe*>(M.outpipe); M.outpipe = nullptr; }
extern "C" {

int speedy_write_output_async(speedy_ctx* ctx, int member, const char* dir, const int* ymdhm, long long timestep) {
    API_BEGIN
    check_ready(ctx, true);                 // output() only reads the state: a main loop in flight keeps its phi_next
    if (member < 0 || member >= ctx->nmembers) throw std::runtime_error("bad member index");
    if (!ymdhm) throw std::runtime_error("speedy_write_output_async: the date of the enqueued state is required");
    Model& M = *ctx->model;
    const size_t n = speedy_output_len(ctx);
    if (!M.outpipe) {                       // published only when complete: a start that throws (pinned memory) leaves no half-built pipe behind
        std::unique_ptr<OutPipe> p(new OutPipe);
        p->start(ctx->device, n);
        M.outpipe = p.release();
    }
    OutPipe& P = *static_cast<OutPipe*>(M.outpipe);
    const int s = P.acquire();
    try {
        float* d = enqueue_output(ctx, member);
        CUDA_CHECK(cudaMemcpyAsync(P.slot[s], d, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaEventRecord(P.ev[s], ctx->stream));
    } catch (...) { P.release(s); throw; }
    char name[40];
    snprintf(name, sizeof name, "%04d%02d%02d%02d%02d.nc", ymdhm[0], ymdhm[1], ymdhm[2], ymdhm[3], ymdhm[4]);
    OutPipe::Job j;
    j.slot = s;
    j.path = (dir && *dir) ? std::string(dir) + "/" + name : std::string(name);
    j.trunc = ctx->d.trunc; j.nsteps = ctx->tab.c.nsteps; j.timestep = timestep; j.ng = ctx->d.ngrid();
    memcpy(j.start, M.start, sizeof j.start);
    P.submit(std::move(j));
    API_END
}

int speedy_output_drain(speedy_ctx* ctx) {
    API_BEGIN
    if (!ctx || !ctx->model) throw std::runtime_error("null context");
    if (ctx->model->outpipe) {
        const std::string e = static_cast<OutPipe*>(ctx->model->outpipe)->drain();
        if (!e.empty()) throw std::runtime_error(e);
    }
    API_END
}

// ---- restart files -----------------------------------------------------------------------------------
// The reference always starts from rest (prognostics.f90:29-31).  A restart file holds everything a context needs to
// continue a run bit for bit: the members' device-resident state (prognostics at both time levels, tendencies, slab
// models, radiation state, SPPT AR(1) state), the calendar and the SPPT draw counter.  The boundary data and tables
// are not in the file: load into a context created with the same configuration after speedy_model_init.
namespace {
struct RestartHeader {
    char magic[8];
    int version, trunc, nmembers, nsteps, sppt_on, member_offset;
    long long stride, istride;
    unsigned long long seed;
    int start[5];
    int phi_next_valid;
    double implicit_dt;
    DevClock clock;
    int precision, sppt_draw;    // version 2: the arithmetic mode and the noise source continue as they were
    int sppt_prepared, reserved; // version 3: the SPPT pattern of the next step is already drawn (the update count in the file includes it)
};
const char kRestartMagic[8] = {'S', 'P', 'D', 'B', '2', '0', '0', 'R'};
}  // namespace

int speedy_save_restart(speedy_ctx* ctx, const char* path) {
    API_BEGIN
    check_ready(ctx, true);
    Model& M = *ctx->model;
    if (!M.initialized) throw std::runtime_error("speedy_save_restart: model not initialized");
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    pull_clock(ctx);
    RestartHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, kRestartMagic, 8);
    h.version = 3; h.precision = ctx->precision; h.sppt_draw = M.sppt_draw ? 1 : 0; h.sppt_prepared = M.sppt_prepared ? 1 : 0; h.trunc = ctx->d.trunc; h.nmembers = ctx->nmembers; h.nsteps = ctx->tab.c.nsteps; h.sppt_on = ctx->sppt_on;
    h.member_offset = ctx->member_offset; h.stride = (long long)M.L.stride; h.istride = (long long)M.L.istride; h.seed = ctx->seed;
    memcpy(h.start, M.start, sizeof h.start);
    h.phi_next_valid = M.phi_next_valid ? 1 : 0;
    h.implicit_dt = M.implicit_dt;
    h.clock = M.hclock;
    std::vector<double> mem(M.mem.n);
    std::vector<int> imem(M.imem.n), sppt(M.sppt_state.n);
    CUDA_CHECK(cudaMemcpy(mem.data(), M.mem.p, mem.size() * sizeof(double), cudaMemcpyDeviceToHost));
    if (!imem.empty()) CUDA_CHECK(cudaMemcpy(imem.data(), M.imem.p, imem.size() * sizeof(int), cudaMemcpyDeviceToHost));
    if (!sppt.empty()) CUDA_CHECK(cudaMemcpy(sppt.data(), M.sppt_state.p, sppt.size() * sizeof(int), cudaMemcpyDeviceToHost));
    FILE* f = fopen(path, "wb");
    if (!f) throw std::runtime_error(std::string("speedy_save_restart: cannot create ") + path);
    const unsigned long long cnt[3] = {mem.size(), imem.size(), sppt.size()};
    bool ok = fwrite(&h, sizeof h, 1, f) == 1 && fwrite(cnt, sizeof cnt, 1, f) == 1;
    ok = ok && fwrite(mem.data(), sizeof(double), mem.size(), f) == mem.size();
    ok = ok && fwrite(imem.data(), sizeof(int), imem.size(), f) == imem.size();
    ok = ok && fwrite(sppt.data(), sizeof(int), sppt.size(), f) == sppt.size();
    ok = (fclose(f) == 0) && ok;
    if (!ok) throw std::runtime_error(std::string("speedy_save_restart: short write to ") + path);
    API_END
}

int speedy_load_restart(speedy_ctx* ctx, const char* path) {
    API_BEGIN
    check_ready(ctx);
    Model& M = *ctx->model;
    if (!M.initialized) throw std::runtime_error("speedy_load_restart: call speedy_model_init first (boundary data and tables are not in the file)");
    FILE* f = fopen(path, "rb");
    if (!f) throw std::runtime_error(std::string("speedy_load_restart: cannot open ") + path);
    RestartHeader h;
    unsigned long long cnt[3];
    std::vector<double> mem;
    std::vector<int> imem, sppt;
    std::string err;
    if (fread(&h, sizeof h, 1, f) != 1 || fread(cnt, sizeof cnt, 1, f) != 1 || memcmp(h.magic, kRestartMagic, 8) != 0 || h.version != 3) err = "not a restart file of this library (version 3)";
    else if (h.trunc != ctx->d.trunc || h.nmembers != ctx->nmembers || h.nsteps != ctx->tab.c.nsteps || h.sppt_on != ctx->sppt_on ||
             h.stride != (long long)M.L.stride || h.istride != (long long)M.L.istride || cnt[0] != M.mem.n || cnt[1] != M.imem.n || cnt[2] != M.sppt_state.n)
        err = "restart file was written by a context of another configuration (trunc / nmembers / nsteps / sppt_on)";
    // a bit-identical continuation needs the same SPPT stream and arithmetic: refuse a silent change of either
    else if (ctx->sppt_on && (h.seed != ctx->seed || h.member_offset != ctx->member_offset))
        err = "restart file was written with another SPPT seed / member_offset (create the context with the file's values)";
    else if (h.precision != ctx->precision || (ctx->sppt_on && h.sppt_draw != (M.sppt_draw ? 1 : 0)))
        err = "restart file was written with another precision mode / SPPT noise source";
    else {
        mem.resize(cnt[0]); imem.resize(cnt[1]); sppt.resize(cnt[2]);
        if (fread(mem.data(), sizeof(double), mem.size(), f) != mem.size() || fread(imem.data(), sizeof(int), imem.size(), f) != imem.size() ||
            fread(sppt.data(), sizeof(int), sppt.size(), f) != sppt.size())
            err = "truncated restart file";
    }
    fclose(f);
    if (!err.empty()) throw std::runtime_error("speedy_load_restart: " + err + ": " + path);
    drop_graph(M);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    CUDA_CHECK(cudaMemcpy(M.mem.p, mem.data(), mem.size() * sizeof(double), cudaMemcpyHostToDevice));
    if (!imem.empty()) CUDA_CHECK(cudaMemcpy(M.imem.p, imem.data(), imem.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (!sppt.empty()) CUDA_CHECK(cudaMemcpy(M.sppt_state.p, sppt.data(), sppt.size() * sizeof(int), cudaMemcpyHostToDevice));
    memcpy(M.start, h.start, sizeof h.start);
    M.hclock = h.clock;
    push_clock(ctx);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (h.implicit_dt != M.implicit_dt && h.implicit_dt > 0.0 && speedy_initialize_implicit(ctx, h.implicit_dt)) throw std::runtime_error(speedy_last_error());
    M.phi_next_valid = h.phi_next_valid != 0;
    M.sppt_prepared = h.sppt_prepared != 0;
    API_END
}

// SPPT noise source: 1 (default) eta is drawn on the device by the counter generator; 0 eta is read
// from the `sppt_eta` field (caller-supplied noise, e.g. to compare with another implementation)
int speedy_set_sppt_draw(speedy_ctx* ctx, int on) {
    API_BEGIN
    check_ready(ctx);
    ctx->model->sppt_draw = on != 0;      // a pattern the last spectral step has already drawn is still used by the next step
    drop_graph(*ctx->model);   // the flag is a captured kernel argument
    API_END
}

// Kernel-selection switches of a context (A/B measurements and the parity tests of the alternative kernels):
//   "k2_field"      1: grid->spec of ensemble batches through the whole-field FFT kernel (default from SPEEDY_K2_FIELD)
//   "dense_inverse" 1: spec->grid Fourier stage as the dense FFTPACK operator on the FP64 tensor pipe (default from SPEEDY_DENSE_INVERSE)
//   "graphs"        as speedy_set_graphs
int speedy_set_option(speedy_ctx* ctx, const char* name, int value) {
    API_BEGIN
    check_ready(ctx);
    if (!name) throw std::runtime_error("null option name");
    const std::string n(name);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (n == "k2_field") ctx->k2_field = value;
    else if (n == "k2_quad") ctx->k2_quad = value != 0;
    else if (n == "k1_quad") ctx->k1_quad = value != 0;
    else if (n == "member_ready") ctx->member_ready = value != 0;
    else if (n == "l2_discard") ctx->l2_discard = value != 0;
    else if (n == "transient_alias") ctx->transient_alias = value != 0;
    else if (n == "sppt_fold") ctx->sppt_fold = value != 0;
    else if (n == "dense_inverse") ctx->fft_inverse = value == 0;
    else if (n == "graphs") ctx->use_graphs = value != 0;
    else throw std::runtime_error("unknown option " + n);
    drop_graph(*ctx->model);   // the kernel choice is baked into a captured graph
    API_END
}

// In-graph timeline of the main-loop kernels (debug aid): on != 0 makes every kernel stamp the GPU's global
// timer; speedy_trace_read returns, in microseconds per step, the duration of each of the four kernels
// (order of speedy_kernel_names) and the idle gap in front of each, then the number of steps traced.
int speedy_trace(speedy_ctx* ctx, int on) {
    API_BEGIN
    check_ready(ctx);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    drop_graph(*ctx->model);
    if (on) {
        std::vector<unsigned long long> h(64 + 2 * 640 + 16, 0ull);   // + per-CTA timelines of the two quad transforms (transforms_quad.cu CSTAMP)
        for (int i = 0; i < 4; i++) h[i] = ~0ull;
        ctx->trace.upload(h);
        ctx->dv.trace = ctx->trace.p;
    } else {
        ctx->dv.trace = nullptr;
    }
    API_END
}
int speedy_trace_read(speedy_ctx* ctx, double* out9) {
    API_BEGIN
    check_ready(ctx);
    if (!ctx->dv.trace) throw std::runtime_error("tracing is off");
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    unsigned long long h[32];
    CUDA_CHECK(cudaMemcpy(h, ctx->trace.p, sizeof(h), cudaMemcpyDeviceToHost));
    const double n = h[16] ? (double)h[16] : 1.0;
    for (int i = 0; i < 4; i++) { out9[i] = 1e-3 * (double)h[8 + i] / n; out9[4 + i] = 1e-3 * (double)h[12 + i] / n; }
    out9[8] = (double)h[16];
    if (getenv("SPEEDY_TRACE_STAMPS")) {   // column-kernel role stamps of one CTA (physics.cu STAMP), microseconds after the CTA's start
        unsigned long long st[32];
        CUDA_CHECK(cudaMemcpy(st, ctx->trace.p + 32, sizeof(st), cudaMemcpyDeviceToHost));
        fprintf(stderr, "column stamps (us):");
        for (int i = 0; i < 13; i++) fprintf(stderr, " [%d] %.2f", i, 1e-3 * (double)st[i] / n);
        fprintf(stderr, "\nK1 stamps (us) [wait,input,legendre,mma] per field:");
        for (int i = 16; i < 24; i++) fprintf(stderr, " %.2f", 1e-3 * (double)st[i] / n);
        {
            unsigned long long k2[8];
            CUDA_CHECK(cudaMemcpy(k2, ctx->trace.p + 24, sizeof(k2), cudaMemcpyDeviceToHost));
            fprintf(stderr, "\nK2 stamps (us) [wait,dft,fold,legendre] per field:");
            for (int i = 0; i < 8; i++) fprintf(stderr, " %.2f", 1e-3 * (double)k2[i] / n);
        }
        for (int kq = 0; kq < 2; kq++) {     // per-CTA timeline of the last launch of the quad kernels: min / median / max over the CTAs, us after the first start
            std::vector<unsigned long long> c(640);
            CUDA_CHECK(cudaMemcpy(c.data(), ctx->trace.p + 64 + 640 * kq, 640 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            unsigned long long t0 = ~0ull; int nc = 0;
            for (int i = 0; i < 160; i++) if (c[4 * i]) { t0 = std::min(t0, c[4 * i]); nc++; }
            if (!nc) continue;
            fprintf(stderr, "\n%s quad CTAs (%d) [start, first quad, after quads, end] min/med/max us:", kq ? "K2" : "K1", nc);
            for (int sidx = 0; sidx < 4; sidx++) {
                std::vector<double> v;
                for (int i = 0; i < 160; i++) if (c[4 * i] && c[4 * i + sidx]) v.push_back(1e-3 * (double)(c[4 * i + sidx] - t0));
                std::sort(v.begin(), v.end());
                if (!v.empty()) fprintf(stderr, " %.2f/%.2f/%.2f", v.front(), v[v.size() / 2], v.back());
            }
        }
        {
            unsigned long long f2[8];
            CUDA_CHECK(cudaMemcpy(f2, ctx->trace.p + 1344, sizeof(f2), cudaMemcpyDeviceToHost));
            fprintf(stderr, "\nK2 per-field stamps of CTA 0 (kcycles) [band wait, fold, stage A, sync, stage B wait, stage B, sync]:");
            for (int i = 0; i < 7; i++) fprintf(stderr, " %.2f", 1e-3 * (double)f2[i] / n);
        }
        fprintf(stderr, "\nspec_step stamps (us) [operands, spectral tendencies, implicit, leapfrog, geopotential, end]:");
        for (int i = 24; i < 30; i++) fprintf(stderr, " %.2f", 1e-3 * (double)st[i] / n);
        fprintf(stderr, "\n");
    }
    API_END
}

int speedy_ensemble_sums_dev(speedy_ctx* ctx, double* d_sum, double* d_sumsq) {
    API_BEGIN
    check_ready(ctx);
    xform_output(ctx);
    launch_ensemble_sums(ctx, d_sum, d_sumsq);
    API_END
}


// ---- per-kernel timing of the main-loop body (bench.py's roofline leg) -------------------
const char* speedy_kernel_names(void) { return "spec_to_grid grid_columns grid_to_spec spec_step"; }
// Runs nsteps main-loop steps with plain launches, a CUDA-event pair around every launch on
// the context's stream; ms[] receives the mean duration of each kernel (names above).
// flush_l2 != 0 overwrites a 256 MiB buffer before every launch (cold-cache timing).
int speedy_time_kernels(speedy_ctx* ctx, int nsteps, int flush_l2, double* ms) {
    API_BEGIN
    check_ready(ctx);
    Model& M = *ctx->model;
    if (!M.initialized) throw std::runtime_error("speedy_model_init has not been called");
    if (ctx->sppt_on) throw std::runtime_error("speedy_time_kernels: SPPT contexts are not supported");
    if (M.implicit_dt != 2 * ctx->tab.c.delt) set_implicit(ctx, 2 * ctx->tab.c.delt);
    const int NK = 4;
    cudaEvent_t ev[NK][2];
    for (int i = 0; i < NK; i++) { CUDA_CHECK(cudaEventCreate(&ev[i][0])); CUDA_CHECK(cudaEventCreate(&ev[i][1])); }
    DevBuf<double> flush;
    if (flush_l2) flush.alloc((size_t)32 << 20);
    for (int i = 0; i < NK; i++) ms[i] = 0.0;
    const double delt = ctx->tab.c.delt;
    if (nsteps > 0) launch_geopotential(ctx, 2);
    for (int s = 0; s < nsteps; s++) {
        for (int i = 0; i < NK; i++) {
            if (flush_l2) CUDA_CHECK(cudaMemsetAsync(flush.p, s & 1, flush.n * sizeof(double), ctx->stream));
            CUDA_CHECK(cudaEventRecord(ev[i][0], ctx->stream));
            switch (i) {
                case 0: xform_step(ctx, 2, true); break;
                case 1: launch_grid_columns(ctx, 0, -1, 1); break;
                case 2: xform_direct(ctx, true); break;
                case 3: launch_spec_step(ctx, 2, 2, 2 * delt, 0, 1); break;
            }
            CUDA_CHECK(cudaEventRecord(ev[i][1], ctx->stream));
        }
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < NK; i++) {
            float t = 0.f;
            CUDA_CHECK(cudaEventElapsedTime(&t, ev[i][0], ev[i][1]));
            ms[i] += t / nsteps;
        }
    }
    if (nsteps > 0) flush_pending_slab(ctx);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < NK; i++) { cudaEventDestroy(ev[i][0]); cudaEventDestroy(ev[i][1]); }
    API_END
}

// host-only calendar check (no GPU): advance a date by nsteps time steps as newdate does
int speedy_host_calendar(int* ymdhm, int nsteps, double* tmonth, double* tyear, int* imont1) {
    API_BEGIN
    DevClock c;
    calendar_init(c, ymdhm[0], ymdhm[1], ymdhm[2], ymdhm[3], ymdhm[4], 1 << 30);
    for (int s = 0; s < nsteps; s++) cal_advance(c);
    ymdhm[0] = c.year; ymdhm[1] = c.month; ymdhm[2] = c.day; ymdhm[3] = c.hour; ymdhm[4] = c.minute;
    if (tmonth) *tmonth = c.tmonth;
    if (tyear) *tyear = c.tyear;
    if (imont1) *imont1 = c.imont1;
    API_END
}

}  // extern "C"
