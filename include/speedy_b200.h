/* speedy_b200.h — C ABI of the B200-native (sm_100a) replacement for the hot path of
 * samhatfield/speedy.f90: spectral transforms, spectral operators, grid-point dynamics,
 * column physics, semi-implicit spectral time step and the per-step surface slabs.
 *
 * Every entry point replaces a Fortran module procedure of the reference; the comment
 * above each one names it (file:line relative to the reference's source/).  The
 * reference-side iso_c_binding interface block is in fortran/speedy_b200_c.f90 and the
 * integration recipe in INTEGRATION.md.
 *
 * Conventions
 *  - All arrays are Fortran (column-major) order, exactly as the reference declares them:
 *    complex(mx,nx) spectral fields are interleaved (re,im) doubles, m fastest;
 *    real(ix,il) grid fields have longitude fastest, latitude j=1 southernmost.
 *  - Batched calls take `nbatch` fields stored back to back.
 *  - Host-pointer entry points copy in/out (synchronously, on the ctx stream);
 *    `_dev` entry points take device pointers and only enqueue work on the ctx stream.
 *  - Every function returns 0 on success, <0 on argument/CUDA errors (text via
 *    speedy_last_error()), >0 for model range errors (check_diagnostics).
 *  - One ctx per (GPU, member batch); calls on one ctx are serialised on its stream and
 *    a ctx is not re-entrant (like the reference's module state).
 *  - There is no CPU fallback: if no CUDA device is usable speedy_create fails.
 */
#ifndef SPEEDY_B200_H
#define SPEEDY_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct speedy_ctx speedy_ctx;

typedef struct speedy_cfg {
    int trunc;     /* params.f90:19  30 (T30: ix=96, iy=24) or 47 (ix=144, iy=36) */
    int kx;        /* params.f90:23  must be 8 */
    int ntr;       /* params.f90:26  must be 1 */
    int nmembers;  /* ensemble members batched in this ctx (outermost batch dim), >=1 */
    int device;    /* CUDA device ordinal */
    int sppt_on;   /* params.f90:42 */
    unsigned long long seed; /* SPPT counter-RNG seed (sppt.f90:119-132 uses system_clock) */
    int member_offset; /* global index of this context's member 0 in a sharded ensemble: the SPPT noise of
                        * member e is keyed by (seed, step, member_offset + e, coefficient), so a member's
                        * trajectory does not depend on how the ensemble is spread over GPUs */
    int nsteps;        /* params.f90:30 time steps per day; 0 = the reference's 36 (delt = 2400 s).  A compile-time
                        * parameter in the reference: T47 needs 72 to stay stable beyond a month */
    int precision;     /* 0: fp64 everywhere (types.f90:12).  1: the spherical-harmonic transforms (Legendre + Fourier, incl.
                        * uvspec/grad) in real32; grid-point columns, semi-implicit solve and time stepping stay fp64
                        * (BASELINE configs[4] tolerance study) */
} speedy_cfg;

/* ---- life cycle -------------------------------------------------------------------- */
/* initialize_geometry/spectral/geopotential/horizontal_diffusion/physics
 * (initialization.f90:47-59): builds every table on the host and uploads it. */
int speedy_create(const speedy_cfg* cfg, speedy_ctx** out);
int speedy_destroy(speedy_ctx* ctx);
const char* speedy_last_error(void);
int speedy_synchronize(speedy_ctx* ctx);
/* dims[8] = trunc, ix, iy, il, kx, nx, mx, ntr */
int speedy_dims(const speedy_ctx* ctx, int* dims);
/* Copy a named host table (e.g. "wt","cpol","epsi","el2","dmp","xj","fband", ...) */
int speedy_get_table(const speedy_ctx* ctx, const char* name, double* out, size_t n);
/* Replace a named table with the caller's own (a Fortran host passes the arrays its own
 * initialize_* computed so that table arithmetic is shared bit for bit). */
int speedy_set_table(speedy_ctx* ctx, const char* name, const double* in, size_t n);

/* Host-only access to the start-up tables for a truncation (no GPU needed; the CPU test
 * suite checks them against the oracle). */
int speedy_host_table(int trunc, const char* name, double* out, size_t n);
long long speedy_host_table_len(int trunc, const char* name);

/* ---- transforms: legendre.f90 / fourier.f90 / spectral.f90 -------------------------- */
/* spec_to_grid  spectral.f90:98-110 ; kcos[b]==1 -> no scaling, else * cosgr(j)
 * (fourier.f90:47-51). spec: nbatch*complex(mx,nx); grid: nbatch*real(ix,il) */
int speedy_spec_to_grid(speedy_ctx* ctx, const double* spec, int nbatch, const int* kcos, double* grid);
/* grid_to_spec  spectral.f90:112-122 */
int speedy_grid_to_spec(speedy_ctx* ctx, const double* grid, int nbatch, double* spec);
/* legendre_inv  legendre.f90:74-111: real(2*mx,nx) -> real(2*mx,il) */
int speedy_legendre_inv(speedy_ctx* ctx, const double* in, int nbatch, double* out);
/* legendre_dir  legendre.f90:114-155: real(2*mx,il) -> real(2*mx,nx) */
int speedy_legendre_dir(speedy_ctx* ctx, const double* in, int nbatch, double* out);
/* fourier_inv   fourier.f90:23-53: real(2*mx,il) -> real(ix,il) */
int speedy_fourier_inv(speedy_ctx* ctx, const double* in, int nbatch, const int* kcos, double* out);
/* fourier_dir   fourier.f90:56-82: real(ix,il) -> real(2*mx,il) */
int speedy_fourier_dir(speedy_ctx* ctx, const double* in, int nbatch, double* out);
/* device-pointer variants (enqueue only; kcos stays a HOST array, NULL = all 1).  The buffers must be 16-byte aligned
 * (cudaMalloc pointers are): the fields are fetched with bulk / tensor-map copies */
int speedy_spec_to_grid_dev(speedy_ctx* ctx, const double* d_spec, int nbatch, const int* kcos, double* d_grid);
int speedy_grid_to_spec_dev(speedy_ctx* ctx, const double* d_grid, int nbatch, double* d_spec);

/* ---- spectral operators: spectral.f90:84-96,124-233 --------------------------------- */
int speedy_laplacian(speedy_ctx* ctx, const double* in, int nbatch, double* out);
int speedy_inverse_laplacian(speedy_ctx* ctx, const double* in, int nbatch, double* out);
int speedy_grad(speedy_ctx* ctx, const double* psi, int nbatch, double* psdx, double* psdy);
int speedy_vds(speedy_ctx* ctx, const double* ucosm, const double* vcosm, int nbatch, double* vorm, double* divm);
int speedy_uvspec(speedy_ctx* ctx, const double* vorm, const double* divm, int nbatch, double* ucosm, double* vcosm);
int speedy_vdspec(speedy_ctx* ctx, const double* ug, const double* vg, int nbatch, int kcos, double* vorm, double* divm);
int speedy_trunct(speedy_ctx* ctx, double* vor, int nbatch);

/* ---- model state (prognostics.f90:16-24, module state of physics/slabs) --------------
 * Named fields, each nmembers copies back to back.  Spectral prognostics:
 *   "vor","div","t" complex(mx,nx,kx,2); "tr" complex(mx,nx,kx,2,ntr); "ps" complex(mx,nx,2);
 *   "phi" complex(mx,nx,kx); "phis" complex(mx,nx);
 * grid/surface/forcing fields: see speedy_field_names(). */
int speedy_set_field(speedy_ctx* ctx, const char* name, const double* host, size_t n);
int speedy_get_field(speedy_ctx* ctx, const char* name, double* host, size_t n);
int speedy_get_ifield(speedy_ctx* ctx, const char* name, int* host, size_t n);
const char* speedy_field_names(void);

/* initialize_implicit(dt)  implicit.f90:36-165 (host LU, uploads xj,xc,xd,elz,dmp1*,tref*) */
int speedy_initialize_implicit(speedy_ctx* ctx, double dt);
/* implicit_terms(divdt, tdt, psdt)  implicit.f90:168-217 on the caller's tendencies (complex(mx,nx,kx) x 2, complex(mx,nx); host
 * arrays, in place) with the matrices of the last speedy_initialize_implicit: the operator-level drop-in (the main loop runs it fused
 * inside the spectral step) */
int speedy_implicit_terms(speedy_ctx* ctx, double* divdt, double* tdt, double* psdt);
/* do_horizontal_diffusion(field, fdt_in, dmp, dmp1)  horizontal_diffusion.f90:86-105, the generic of both forms: nlev = 1 for
 * complex(mx,nx), kx for complex(mx,nx,kx); fdt = (fdt - dmp*field)*dmp1 in place; dmp / dmp1 are real(mx,nx) (the module's
 * coefficient arrays are the tables "dmp", "dmpd", "dmps", "dmp1", "dmp1d", "dmp1s" of speedy_get_table) */
int speedy_do_horizontal_diffusion(speedy_ctx* ctx, const double* field, double* fdt, const double* dmp, const double* dmp1, int nlev);
/* get_geopotential  geopotential.f90:33-57 on the resident state: phi <- T(time level j) */
int speedy_get_geopotential(speedy_ctx* ctx, int j);
/* get_tendencies    tendencies.f90:11-37 (grid-point + spectral + implicit) into the
 * resident tendency arrays ("vordt","divdt","tdt","trdt","psdt") */
int speedy_get_tendencies(speedy_ctx* ctx, int j2, int compute_shortwave);
/* get_physical_tendencies physics.f90:43-223 on host arrays (testing/drop-in form):
 * spectral inputs complex(mx,nx,kx) [psl complex(mx,nx)], grid tendencies real(ix,il,kx) in/out */
int speedy_get_physical_tendencies(speedy_ctx* ctx, const double* vor, const double* div, const double* t,
                                   const double* q, const double* phi, const double* psl,
                                   double* utend, double* vtend, double* ttend, double* qtend,
                                   int compute_shortwave);
/* step(j1,j2,dt)    time_stepping.f90:35-122 on the resident state */
int speedy_step(speedy_ctx* ctx, int j1, int j2, double dt, int compute_shortwave);
/* first_step        time_stepping.f90:12-24 */
int speedy_first_step(speedy_ctx* ctx);
/* couple_sea_land per-step slab update (land_model.f90:184-239, sea_model.f90:253-444);
 * daily climatological inputs must have been set with speedy_set_field */
int speedy_couple_sea_land(speedy_ctx* ctx, int day);
/* set_forcing daily part that depends on resident state (forcing.f90:55-99):
 * albedos, snow cover, tcorh/qcorh */
int speedy_set_forcing(speedy_ctx* ctx, int imode);
/* check_diagnostics diagnostics.f90:16-75: diag[kx*3] (reke, deke, temp per level), member 0..;
 * returns 1 if out of range like the reference's `stop` */
int speedy_check_diagnostics(speedy_ctx* ctx, int time_level, double* diag);
/* main-loop body speedy.f90:27-54 repeated nsteps times with the state resident;
 * `model_step` is the 1-based step counter of the reference (speedy.f90:21).
 * Daily host-side inputs are pulled through the callback-free "env" below. */
int speedy_run_steps(speedy_ctx* ctx, int nsteps);
/* the same loop in two halves: speedy_enqueue_steps only enqueues the work on the ctx stream (no host synchronisation: many
 * days can be queued back to back), speedy_finish drains the stream and returns 1 if check_diagnostics tripped at any step */
int speedy_enqueue_steps(speedy_ctx* ctx, int nsteps);
int speedy_finish(speedy_ctx* ctx);

/* ---- model environment: boundaries.f90, forcing.f90, date.f90, land/sea init -------- */
/* initialize (initialization.f90:12-82) from a boundary-condition source.  `bc_path` is either
 *  - a DIRECTORY holding the reference's own NetCDF-4 files — side by side, as run.sh links them into the run directory, or as the
 *    data/bc/t30 tree with clim/ and anom/ — read without an HDF5 library (contiguous float32 variables located through their HDF5
 *    link / object-header messages, csrc/host/env.cpp; all 420 months of the SST anomaly are resident then), or
 *  - the packed .bin that tools/pack_boundary.py makes from those files (what travels with this repo: float32 variables copied
 *    verbatim, a window of the SST-anomaly record — 72 months from 1979-01 as shipped, `--months` widens it; a run that leaves the
 *    window fails with an error instead of reading past it).
 * The flip / missing-value / forchk logic (input_output.f90:36-41, boundaries.f90:47-72) runs in the loader for both.  Start date as in namelist.nml. */
int speedy_model_init(speedy_ctx* ctx, const char* bc_path, int year, int month, int day, int hour, int minute);
/* host-only (no GPU): the start-up boundary field `name` as boundaries.f90:28-68 / land_model.f90:50-181 / sea_model.f90:80-250 leave it,
 * from either source ("phi0", "fmask", "alb0", "stl12", "snowd12", "soilw12", "sst12", "sice12", "ssta", "fmask_l", ...).  Returns the
 * field's length (out == NULL: just that) or -1; copies its first n values. */
long long speedy_host_boundary(const char* bc_path, int trunc, const char* name, double* out, size_t n);
/* current model date (date.f90:20) and step counter */
int speedy_model_date(const speedy_ctx* ctx, int* ymdhm, long long* model_step);
/* output() conversions input_output.f90:184-214: float32 u,v,t,q,phi (ix,il,kx) and ps (ix,il) */
int speedy_output_fields(speedy_ctx* ctx, int member, float* u, float* v, float* t, float* q, float* phi, float* ps);
/* output() file writer, input_output.f90:95-217: one NetCDF classic (CDF-1) file `yyyymmddhhmm.nc` per output time
 * with time/lon/lat/lev and float32 u,v,t,q,phi(lon,lat,lev,time), ps(lon,lat,time), names, attributes and coordinate
 * arithmetic as in the reference; written without a NetCDF library.  speedy_write_output converts member `member`'s
 * resident state on the device and writes the file into `dir` (path returned in path_out when non-NULL);
 * speedy_write_output_file is the host-only writer underneath (timestep = model_step - 1, speedy.f90:50). */
int speedy_write_output(speedy_ctx* ctx, int member, const char* dir, char* path_out, size_t path_cap);
/* the same without waiting for the device or the file system: the conversions are enqueued on the context's stream, the fields
 * land in a ring of pinned host buffers and host threads write the files while the device runs on (the reference's default
 * writes a file after EVERY step).  ymdhm / timestep: the model date and `model_step - 1` the enqueued state will have (the host knows
 * the calendar: speedy_host_calendar; no device round trip).  speedy_output_drain returns when every file is on disk (<0: a write
 * failed); speedy_destroy drains too. */
int speedy_write_output_async(speedy_ctx* ctx, int member, const char* dir, const int* ymdhm, long long timestep);
int speedy_output_drain(speedy_ctx* ctx);
int speedy_write_output_file(const char* path, int trunc, int nsteps, const int* start_ymdhm, int timestep,
                             const float* u, const float* v, const float* t, const float* q, const float* phi, const float* ps);
/* restart files (absent in the reference, which always starts from rest, prognostics.f90:29-31): the complete
 * device-resident state of all members + calendar + SPPT counters; a run continued from a restart file is bit-identical
 * to the uninterrupted one.  Load into a context of the same configuration after speedy_model_init. */
int speedy_save_restart(speedy_ctx* ctx, const char* path);
int speedy_load_restart(speedy_ctx* ctx, const char* path);
/* ensemble sums for the mean/spread diagnostic: writes sum and sum of squares of the 41
 * output levels over this ctx's members into device buffers (for NCCL all-reduce) */
/* sppt.f90:45-99 noise source: on != 0 (default) draws eta on the device; 0 reads it from the
 * `sppt_eta` field set by the caller */
int speedy_set_sppt_draw(speedy_ctx* ctx, int on);
int speedy_ensemble_sums_dev(speedy_ctx* ctx, double* d_sum, double* d_sumsq);
size_t speedy_output_len(const speedy_ctx* ctx); /* (5*kx+1)*ix*il */

/* host-resident drop-in step: uploads the prognostic state from host arrays, runs
 * step(j1,j2,dt), downloads it (the literal replacement of `call step` with the Fortran
 * module arrays left on the host).  state layout: vor,div,t,tr,ps concatenated. */
int speedy_step_host(speedy_ctx* ctx, double* state, size_t n, int j1, int j2, double dt, int compute_shortwave);
size_t speedy_state_len(const speedy_ctx* ctx);
/* host-resident drop-in of the main loop (speedy.f90:27-54 x nsteps): state = nmembers x
 * [vor,div,t,tr,ps] in HOST memory is uploaded, advanced and downloaded; out (nullable,
 * speedy_output_len floats: u,v,t,q,phi (ix,il,kx) then ps) gets member 0's output() fields */
int speedy_run_steps_host(speedy_ctx* ctx, double* state, size_t n, int nsteps, float* out);

/* number of kernel launches issued by this ctx so far (bench.py's gpu_launches) */
long long speedy_launch_count(const speedy_ctx* ctx);
/* raw CUDA stream handle (cudaStream_t) for event timing from the host language */
void* speedy_stream(const speedy_ctx* ctx);
/* use CUDA graphs for speedy_run_steps (1, default) or plain launches (0) */
int speedy_set_graphs(speedy_ctx* ctx, int on);
/* kernel-selection switches (no counterpart in the reference; used by the A/B tools and by the parity tests of the alternative
 * kernels): "k2_field" (grid->spec batches through the whole-field FFT kernel), "dense_inverse" (spec->grid Fourier stage as the
 * dense FFTPACK operator), "k1_quad" / "k2_quad" (ensemble batches through the four-field FFT + DMMA kernels, default 1),
 * "member_ready" (default 1: in the main-loop step the column tiles of a member start when that member's grid fields are stored,
 * 0: when the whole spec->grid launch is complete), "l2_discard" (default 1: the ensemble step drops its transient grid fields
 * from L2 after their only read), "transient_alias" (default 1: the ensemble step keeps grid fields, grid tendencies and their
 * coefficients in one buffer per member; get_field of "gin" / "gout" / "sout" is then not meaningful after a step), "graphs",
 * "sppt_fold" (default 1: with device-drawn SPPT noise the spectral step prepares the next step's pattern itself instead of a
 * separate kernel per step).  None of the last four changes a bit of the results.  Returns <0 for an unknown name. */
int speedy_set_option(speedy_ctx* ctx, const char* name, int value);


/* bench/profiling: mean CUDA-event duration (ms) of each kernel of the main-loop body over
 * nsteps plain-launch steps; ms[10] in the order of speedy_kernel_names() */
int speedy_time_kernels(speedy_ctx* ctx, int nsteps, int flush_l2, double* ms);
const char* speedy_kernel_names(void);
/* in-graph timeline (debug aid): per-step duration of each kernel and the idle gap in front of it, microseconds;
 * out9 = 4 durations, 4 gaps, number of steps traced */
int speedy_trace(speedy_ctx* ctx, int on);
int speedy_trace_read(speedy_ctx* ctx, double* out9);
/* host-only: date.f90:109-157 newdate applied nsteps times to ymdhm[5] (no GPU needed) */
int speedy_host_calendar(int* ymdhm, int nsteps, double* tmonth, double* tyear, int* imont1);

/* ---- the caller of the path: program speedy (speedy.f90:1-54) ------------------------- */
/* The two namelist groups the reference reads from namelist.nml: &params (params.f90:46-70) and &date
 * (date.f90:54-71); datetimes are (year, month, day, hour, minute) as type(datetime), date.f90:14-20. */
typedef struct speedy_namelist {
    int nsteps_out;          /* params.f90:50  time steps between outputs, default 1 */
    int nstdia;              /* params.f90:49  period of the diagnostic print-out, default 36*5 */
    int start_datetime[5];   /* date.f90:62    default 1982-01-01 00:00 */
    int end_datetime[5];     /* date.f90:63    default 1982-02-01 00:00 */
} speedy_namelist;
/* the reference's defaults (host-only) */
int speedy_namelist_defaults(speedy_namelist* nml);
/* initialize_params + the namelist part of initialize_date (host-only): a missing file keeps the defaults, as the
 * reference's `inquire(exist=)` does; a file that exists must hold both groups.  Accepts `name = v`,
 * `start_datetime%year = v` and `start_datetime = y, m, d, h, mi`, `!` comments, any letter case. */
int speedy_read_namelist(const char* path, speedy_namelist* nml);
/* trips of `do while (.not. datetime_equal(model_datetime, end_datetime))` (speedy.f90:27) under newdate's calendar
 * (date.f90:109-157) at nsteps steps per day; -1 if the end date is never met (the reference would not terminate) */
long long speedy_steps_between(const int* start_ymdhm, const int* end_ymdhm, int nsteps);
/* info[4] = nmembers, time steps per day (params.f90:30), sppt_on, precision of the context */
int speedy_run_info(const speedy_ctx* ctx, int* info);
/* 0, or 1 once check_diagnostics has stopped the run (sticky): *step = the main-loop step that left the accepted range
 * (the reference's istep, diagnostics.f90:60-69) and diag = the diag(kx,3) of member 0 at that step; nothing is recomputed */
int speedy_range_failure(speedy_ctx* ctx, long long* step, double* diag);
/* The main program from after `call initialize` to `end` (speedy.f90:24-54) plus the step-0 print and file of
 * initialize_prognostics (prognostics.f90:120-126), on a context whose speedy_model_init ran with nml->start_datetime:
 * the state stays on the device; every nsteps_out steps member `member` (every member into out_dir/member<e>/ when
 * member < 0) is written as yyyymmddhhmm.nc into out_dir (NULL: no files); every nstdia steps the reference's
 * ' step = ... reke = / deke = / temp =' lines go to stdout when verbose != 0.  Returns 0 at the end date, 1 for
 * 'Model variables out of accepted range' (the lines of the failing step are printed, diagnostics.f90:60-69),
 * <0 on errors; *steps_done = main-loop steps completed. */
int speedy_main_loop(speedy_ctx* ctx, const speedy_namelist* nml, const char* out_dir, int member, int verbose, long long* steps_done);

#ifdef __cplusplus
}
#endif
#endif
