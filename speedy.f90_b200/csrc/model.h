// Device-resident model state of one context (a batch of ensemble members on one GPU).
// All per-member arrays live in ONE allocation at fixed offsets inside a member region
// (member e starts at base + e*stride), in the reference's own Fortran layouts, so that
// speedy_set_field/get_field are plain copies of the reference's module arrays.
#pragma once
#include "ctx.h"
#include <map>
#include <string>
#include <vector>

namespace spd {

// device-resident calendar and per-step flags (date.f90:20-38, speedy.f90:21-54)
struct DevClock {
    int model_step;      // 1-based main-loop step counter (speedy.f90:21)
    int csw;             // compute_shortwave for the coming step (speedy.f90:35)
    int year, month, day, hour, minute;
    int start_year;
    int imont1;          // date.f90:33
    int doy;             // 0-based day-of-year index into the daily solar tables
    int do_forcing;      // set_forcing(1) is due before the coming step (speedy.f90:29-32)
    int obs_ssta;        // the next couple_sea_atm call fires obs_ssta (sea_model.f90:273)
    int next_month;      // 1-based record obs_ssta reads (sea_model.f90:377)
    int ssta_missing;    // sticky: a record outside the resident ssta window was requested
    int diag_fail;       // sticky: check_diagnostics range violation (diagnostics.f90:60-70); holds the failing step
    int nssta;           // records in the resident ssta window
    int slab_pending;    // the coupler call of the last completed step has not run yet (it rides in the next column kernel)
    int nsteps;          // time steps per day (params.f90:30)
    int close_pending;   // the step's closing (diagnostics reduction + calendar) has not run yet: it rides in the next spec->grid kernel
    double tmonth, tyear;
    double diag[24];     // (kx,3) of the last check_diagnostics
};

// K1 output fields (grid) ---------------------------------------------------------------
enum {
    GI_VOR = 0, GI_DIV = 8, GI_T = 16, GI_TR = 24, GI_U = 32, GI_V = 40, GI_PX = 48, GI_PY = 49,   // dynamics, time level j2
    GI_U1 = 50, GI_V1 = 58, GI_T1 = 66, GI_Q1 = 74, GI_PHI = 82, GI_PSL = 90,                     // physics, time level 1
    GI_SPPT = 91, GI_N = 99, GI_NBASE = 91
};
// K2 input fields: per level 9 (utend, vtend, KE, -uT', -vT', ttend, -uq, -vq, qtend) + psdt
enum { GO_PER = 9, GO_PSDT = 72, GO_QCORH = 73, GO_N = 74 };   // slot 73: daily humidity-correction field (forcing.f90:98)

// offsets (in doubles) inside a member region
struct Layout {
    long long stride;
    // spectral (complex interleaved)
    long long vor, div, t, tr, ps, phi, phi_next, phis, tcorh, qcorh, sout, sppt_spec, sppt_eta;
    long long vordt, divdt, tdt, trdt, psdt;
    // grid work
    long long gin, gout;
    // grid state (doubles)
    long long phis0, fmask_l, fmask_s, forog, alb0;
    long long fsol, ozone, ozupp, zenit, stratz, alb_l, alb_s, albsfc, snowc;
    long long stl_am, stl_lm, snowd_am, soilw_am, sst_am, sice_am, tice_am, ssti_om, sst_om, tice_om, sice_om;
    long long sstcl_ob, sicecl_ob, ticecl_ob, stlcl_ob, sstan3;
    long long tau2, stratc, tt_rsw, ssrd, ssr, tsr;
    long long precnv, precls, cbmf, slrd, slr, olr, slru, ustr, vstr, shf, evap, hfluxn, ts, tskin, u0, v0, t0;
    long long qcloud, cloudc, clstr;
    long long qcorh_g;   // grid-point humidity correction before its transform (forcing.f90:98)
    // int fields, offsets in ints inside the member's int region
    long long istride, iptop, icltop, icnv;
};

// member-independent device tables of the surface models
struct SharedDev {
    const double *stl12, *snowd12, *soilw12, *sst12, *sice12;       // (ix,il,12)
    const double *rhcapl, *cdland, *rhcaps, *rhcapi, *cdsea, *cdice, *bmask_s;   // (ix,il)
    const float* ssta;      // (ix,il,nssta) already flipped to S->N, forchk-masked
    const double* solar;    // [365][5][il]: fsol, ozone, ozupp, zenit, stratz per latitude
};

struct FieldInfo { long long off; size_t len; bool is_int; };

struct Model {
    Layout L;
    DevBuf<double> mem;       // nmembers * L.stride
    DevBuf<int> imem;         // nmembers * L.istride
    DevBuf<double> shared;    // climatologies, slab constants, solar tables
    DevBuf<float> ssta;
    SharedDev sh;
    DevBuf<DevClock> clock;   // one clock (members share the calendar)
    DevBuf<LevelConsts> lc;
    DevBuf<float> outbuf;          // output() staging (input_output.f90:201-206)
    DevBuf<unsigned> ready;        // [member] completion counts of the step's spec->grid kernel (member_ready.cuh)
    bool alias_active = false;     // while a main-loop step is being enqueued: gout and sout live in the gin rows (speedy_ctx::transient_alias)
    unsigned ready_target = 0;     // while a main-loop step is being enqueued: the count at which a member's grid fields are complete; 0: hand-off by kernel boundary
    DevBuf<double> diag_partial;   // [member][block][kx][2] + [member][kx] partial sums of check_diagnostics
    DevBuf<XDesc> desc_step;       // compact per-step list (two time levels): the fields the column kernel actually reads
    int nstep_fields = 0;
    DevBuf<XDesc> desc_inv, desc_dir, desc_out, desc_one_dir, desc_sppt;
    std::map<std::string, FieldInfo> fields;
    // host-side calendar mirror
    DevClock hclock;
    int start[5] = {1982, 1, 1, 0, 0};   // start_datetime (date.f90:21), for the time axis of the output files
    bool initialized = false;
    bool phi_next_valid = false;   // phi_next matches the resident level-1 temperature (true after main-loop steps)
    // host copies needed by the daily/implicit logic
    std::vector<double> h_phis0;
    // CUDA graph of one day (36 steps) of the main loop
    cudaGraphExec_t day_graph = nullptr;
    int day_graph_steps = 0;
    double implicit_dt = 0.0;
    // SPPT (sppt.f90): AR(1) state is device-resident; eta drawn on device unless supplied
    bool sppt_draw = true;
    bool sppt_prepared = false;    // the SPPT pattern of the next get_tendencies call is already on the device (the last spectral step drew it)
    void* colmaps = nullptr;       // tensor maps of the column kernel's tiles (physics.cu)
    void* outpipe = nullptr;       // asynchronous output: pinned slots + writer threads (model.cu, speedy_write_output_async)
    DevBuf<int> sppt_state;   // [0] AR(1) updates done so far (device-resident: CUDA-graph replays advance it), [1] block ticket
};

// ---- kernels (dynamics.cu / physics.cu) ------------------------------------------------
void launch_geopotential(speedy_ctx* ctx, int which);   // which: bit0 module phi, bit1 phi_next (K1's physics input)
void launch_grid_columns(speedy_ctx* ctx, int mode, int csw_override, int merged = 0);   // mode 0 dyn+phys, 1 physics only on resident tendencies
void launch_spec_step(speedy_ctx* ctx, int j1, int j2, double dt, int store_tend_only, int close_step = 0);
void launch_close_step(speedy_ctx* ctx);
void launch_implicit_terms(speedy_ctx* ctx, double* d_divdt, double* d_tdt, double* d_psdt);   // stand-alone implicit.f90:168-217 on device arrays
void launch_horizontal_diffusion(speedy_ctx* ctx, const double* d_field, double* d_fdt, const double* d_dmp, const double* d_dmp1, int nlev);
void free_column_maps(Model& M);   // stand-alone closing of a pending step
void launch_diagnostics(speedy_ctx* ctx, int level);
void launch_slab(speedy_ctx* ctx, int day0);
void launch_daily_forcing(speedy_ctx* ctx, int force);
void launch_qcorh_finish(speedy_ctx* ctx);
void launch_clock_advance(speedy_ctx* ctx);
void launch_output_convert(speedy_ctx* ctx, int member, float* d_out);
void launch_ensemble_sums(speedy_ctx* ctx, double* d_sum, double* d_sumsq);
void launch_sppt_update(speedy_ctx* ctx);

// ---- host environment (host/env.cpp) ----------------------------------------------------
struct HostEnv {
    int ix, il, nssta;
    std::vector<double> phi0, fmask, alb0;                       // boundaries.f90
    std::vector<double> fmask_l, bmask_l, stl12, snowd12, soilw12, rhcapl, cdland;   // land_model.f90
    std::vector<double> fmask_s, bmask_s, sst12, sice12, rhcaps, rhcapi, cdsea, cdice, deglat_s;   // sea_model.f90
    std::vector<float> ssta;                                     // all resident records, flipped + masked
    std::vector<double> solar;                                   // [365][5][il]
};
void load_host_env(const char* bc_path, const Tables& tab, HostEnv& env);   // throws on error
void calendar_init(DevClock& c, int y, int m, int d, int h, int mi, int nssta, int nsteps = 36);
void calendar_advance(DevClock& c);   // speedy.f90:44-47 + flags for the next step (same arithmetic as the device kernel)

}  // namespace spd
