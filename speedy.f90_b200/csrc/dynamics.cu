// K3 / K6 — spectral-space kernels of the time step.
//   k_geopotential  : get_geopotential (geopotential.f90:33-57) outside the fused main loop (uvspec/grad live in K1's input stage)
//   k_spec_step     : vds/laplacian assembly of the transformed tendencies          (tendencies.f90:212-234)
//                     + get_spectral_tendencies (:242-293) + implicit_terms (implicit.f90:168-217)
//                     + 7 horizontal-diffusion passes, stratospheric drag (time_stepping.f90:63-96)
//                     + trunct, leapfrog and Robert-Asselin-Williams filter (:127-167)
//                     — all of it is local to one (m,n) coefficient across the 8 levels, so one
//                     thread owns a coefficient and the whole chain is ONE launch.
//   k_diagnostics   : check_diagnostics (diagnostics.f90:16-75) into the device clock
//   k_output        : output() conversions (input_output.f90:201-206)
// Compiled with --fmad=false (operation-by-operation agreement with the checker).
#include "model.h"
#include "spectral_ops.cuh"
#include "calendar.h"
#include "tma.cuh"
#include "close_step.cuh"

namespace spd {

#define KX 8

// SPPT pattern update (see sppt_point below)
struct SpptParams {
    unsigned long long seed; int* state; int member0; int draw;     // state[0] = updates done so far, state[1] = block ticket; state == nullptr: off
    double phi, f0;          // sppt.f90:32 and :76-80, evaluated once on the host (thirty exp() per thread otherwise)
};

struct SpecArgs {
    double* base; long long stride;
    Layout L;
    DevTables tv;
    const LevelConsts* lc;
    DevClock* clk;
    int j1, j2;
    double dt;
    int flag;
    double* partial;
    SpptParams sppt;         // state != nullptr: this launch also prepares the SPPT pattern of the next get_tendencies call (prologue)
    long long sout_off, sout_fs;   // the grid->spec output of the step: offset in the member block and doubles between fields (packed, or in place over the grid rows)
    unsigned* ready_reset;   // main-loop step: the per-member completion counts of this step's spec->grid kernel, zeroed here for the next step
};

__device__ __forceinline__ const double* sfield(const double* mb, long long off, int nsp, int f) { return mb + off + (size_t)f * nsp * 2; }
__device__ __forceinline__ double* sfield(double* mb, long long off, int nsp, int f) { return mb + off + (size_t)f * nsp * 2; }

// get_geopotential(t(:,:,:,1), phis)  geopotential.f90:33-57 for the paths outside the fused main loop.
// flag bit0: write the module variable phi; bit1: write phi_next (the field K1 transforms for the physics)
__global__ void k_geopotential(SpecArgs a) {
    const int mx = a.tv.mx, nx = a.tv.nx, nsp = mx * nx;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nsp) return;
    const int n = r / mx, m = r - n * mx;
    double* mb = a.base + (size_t)blockIdx.y * a.stride;
    const LevelConsts& lc = *a.lc;
    cd tt[KX], ph[KX];
#pragma unroll
    for (int kk = 0; kk < KX; kk++) tt[kk] = ld(sfield(mb, a.L.t, nsp, kk), mx, m, n);
    ph[KX - 1] = ld(mb + a.L.phis, mx, m, n) + lc.xgeop1[KX - 1] * tt[KX - 1];
#pragma unroll
    for (int kk = KX - 2; kk >= 0; kk--) ph[kk] = (ph[kk + 1] + lc.xgeop2[kk + 1] * tt[kk + 1]) + lc.xgeop1[kk] * tt[kk];
    if (m == 0) {
#pragma unroll
        for (int kk = 1; kk < KX - 1; kk++) ph[kk] = ph[kk] + lc.geop_corf[kk] * (tt[kk + 1] - tt[kk - 1]);
    }
#pragma unroll
    for (int kk = 0; kk < KX; kk++) {
        if (a.flag & 1) st(sfield(mb, a.L.phi, nsp, kk), mx, m, n, ph[kk]);
        if (a.flag & 2) st(sfield(mb, a.L.phi_next, nsp, kk), mx, m, n, ph[kk]);
    }
}

// k_spec_step — block = 32 coefficients x 8 levels (threadIdx.y = level, one warp per level).
// Everything that couples the levels of one coefficient (vertical sums, sigma-dot prefix,
// hydrostatic integration, the three 8x8 mat-vecs of the semi-implicit solve) goes through
// shared memory in the reference's summation order; everything else is one thread per (m,n,k).
// flag bit0: stop after implicit_terms and store the tendencies (get_tendencies drop-in)
// flag bit1: main-loop step: take qcorh from the day's transform when due, add the
//            check_diagnostics partial sums and flag the step as waiting to be closed (close_step.cuh: final
//            diagnostics reduction in fixed order, range guard, calendar advance — in the next spec->grid kernel).
#define SC 32
constexpr int SPEC_LC_DOUBLES = sizeof(LevelConsts) / sizeof(double);
static size_t spec_step_smem(int mx, int nx) { return sizeof(double) * (SPEC_LC_DOUBLES + 2 * KX * KX + (size_t)KX * KX * (mx + nx + 1)) + sizeof(uint64_t); }
// BATCH = false (single-member step): no register cap, every operand load of a thread is in flight at once;
// BATCH = true (ensemble batches): two CTAs per SM
// ---- SPPT AR(1) update in spectral space (sppt.f90:74-91) -------------------------------------------------
// eta is supplied (caller noise) or drawn with a counter-based generator keyed by (update count, global member, level, coefficient).
// One point = one (level, coefficient) of one member.  Two callers: the stand-alone kernel (first update of a run, supplied noise,
// the physics-only entry point) and the prologue of k_spec_step, which prepares the pattern of the NEXT get_tendencies call while it
// waits for its own inputs: a separate kernel between the spectral step and the next spec->grid launch holds every SM until the
// spectral step has drained (its blocks end in griddepcontrol.wait to keep the dependency chain transitive), so that the 216 KB
// transform CTAs cannot start their prologue underneath it.
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ double u01(unsigned long long h) { return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
// t = level * nspec + coefficient (the stand-alone kernel's thread index: the noise of a point does not depend on who computes it)
__device__ __forceinline__ void sppt_point(double* mb, const Layout& L, const DevTables& tv, const SpptParams& p, int counter, int e, int t, bool live) {
    const int mx = tv.mx, nx = tv.nx, nsp = mx * nx;
    const int k = live ? t / nsp : 0, r = live ? t - k * nsp : 0;
    const int n = r / mx, m = r - n * mx;
    const bool first = counter == 0;
    const double len_decorr = 500000.0;
    const double phi = p.phi, f0 = p.f0;
    const double sigma = f0 * exp(-0.25 * (len_decorr * len_decorr) * tv.el2[r]);
    cd eta;
    if (p.draw) {
        // counter-based Box-Muller (sppt.f90:103-117 shape: u = sqrt(-2 ln r1), v = 2*2pi*r2, sin only), clipped to +-10
        const unsigned long long id = ((unsigned long long)counter * 65536ull + (unsigned long long)(p.member0 + e)) * (unsigned long long)(KX * nsp) + (unsigned long long)t;
        const unsigned long long h = splitmix64(p.seed ^ splitmix64(id));
        const double r1 = u01(splitmix64(h + 1)), r2 = u01(splitmix64(h + 2)), r3 = u01(splitmix64(h + 3)), r4 = u01(splitmix64(h + 4));
        const double c = (double)(2.0f * 6.28318530718f);
        double gr = sqrt(-2.0 * log(r1)) * sin(c * r2), gi = sqrt(-2.0 * log(r3)) * sin(c * r4);
        gr = fmin(10.0, fabs(gr)) * copysign(1.0, gr);
        gi = fmin(10.0, fabs(gi)) * copysign(1.0, gi);
        eta = cd{gr, gi};
        if (live) st(sfield(mb, L.sppt_eta, nsp, k), mx, m, n, eta);
    } else {
        eta = ld(sfield(mb, L.sppt_eta, nsp, k), mx, m, n);
    }
    double* sp = sfield(mb, L.sppt_spec, nsp, k);
    cd v;
    if (first) {
        const double c = pow(1 - phi * phi, -0.5);
        v = (c * sigma) * eta;
    } else {
        v = phi * ld(sp, mx, m, n) + sigma * eta;
    }
    if (live) st(sp, mx, m, n, v);
}
// the update counter lives on the device so that a replayed CUDA graph draws fresh noise every step; the last block to finish
// advances it (every block has read it by then).  One thread per block, after the block's points are stored.
__device__ __forceinline__ void sppt_ticket(int* state, int counter, int nblk) {
    __threadfence();
    if (atomicAdd(&state[1], 1) == nblk - 1) { state[0] = counter + 1; state[1] = 0; __threadfence(); }
}

template <bool BATCH>
__global__ void __launch_bounds__(SC * KX, BATCH ? 2 : 1) k_spec_step(SpecArgs a) {
    const int mx = a.tv.mx, nx = a.tv.nx, nsp = mx * nx;
    const int c = threadIdx.x, k = threadIdx.y;
    const int r = blockIdx.x * SC + c;
    const bool valid = r < nsp;
    const int rr = valid ? r : nsp - 1;
    const int n = rr / mx, m = rr - n * mx;
    double* mb = a.base + (size_t)blockIdx.y * a.stride;
    const DevTables& tv = a.tv;
    const size_t q = rr;
    const cd zero{0.0, 0.0};
    __shared__ cd s_a[KX][SC], s_b[KX][SC], s_phi[KX][SC], s_sig[KX + 1][SC];
    __shared__ cd s_dmeanc[SC], s_psdt[SC];
    // The level constants and the three semi-implicit matrices (implicit.f90:36-165) are staged in shared
    // memory by bulk asynchronous copies; every other global operand of this thread's chain is loaded HERE,
    // before the first barrier, so that the kernel exposes one L2 round trip instead of one per stage.
    extern __shared__ __align__(16) double dsm[];
    double* sLc = dsm;
    double* sXd = sLc + SPEC_LC_DOUBLES;
    double* sXc = sXd + KX * KX;
    double* sXj = sXc + KX * KX;
    const int nl = mx + nx + 1;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sXj + (size_t)KX * KX * nl);
    const int tid = k * SC + c;
    if (tid == 0) trace_begin(tv.trace, 3);
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (tid == 0) {
        const uint32_t mb64 = KX * KX * sizeof(double);
        mbar_expect_tx(bar, (uint32_t)sizeof(LevelConsts) + 2 * mb64 + mb64 * nl);
        bulk_g2s(sLc, a.lc, (uint32_t)sizeof(LevelConsts), bar);
        bulk_g2s(sXd, tv.xd, mb64, bar);
        bulk_g2s(sXc, tv.xc, mb64, bar);
        bulk_g2s(sXj, tv.xjt, mb64 * nl, bar);
    }
    // SPPT pattern of the next get_tendencies call: independent of this step's data, so it runs while the grid->spec kernel drains
    // (and first of all: nothing of the step proper is live in registers yet)
    if (a.sppt.state) {
        const int sppt_counter = a.sppt.state[0];
        sppt_point(mb, a.L, tv, a.sppt, sppt_counter, blockIdx.y, k * nsp + rr, valid);
        __syncthreads();
        if (tid == 0) sppt_ticket(a.sppt.state, sppt_counter, gridDim.x * gridDim.y);    // here, not at the end: two of the paths below return early
    }
    const LevelConsts& lc = *reinterpret_cast<const LevelConsts*>(sLc);
    const double el2 = tv.el2[q], elz_q = tv.elz[q], trf_q = tv.trfilt[q], elm2_q = tv.elm2[q];
    // horizontal-diffusion factors: fetched ahead of the wait in the latency variant; the batch variant (128 registers for two blocks per SM)
    // would spill them across the whole kernel and reads them where they are used (L1 / L2 hits)
    double dmp = 0.0, dmpd = 0.0, dmps = 0.0, dmp1 = 0.0, dmp1d = 0.0, dmp1s = 0.0;
    if (!BATCH) { dmp = tv.dmp[q]; dmpd = tv.dmpd[q]; dmps = tv.dmps[q]; dmp1 = tv.dmp1[q]; dmp1d = tv.dmp1d[q]; dmp1s = tv.dmp1s[q]; }
    pdl_wait();                 // everything above reads constant tables only; the fields below come from the previous kernel
    pdl_trigger();
    if (a.ready_reset && blockIdx.x == 0 && tid == 0) a.ready_reset[blockIdx.y] = 0u;   // every column tile of the step has passed its wait
    const unsigned long long tk0 = tv.trace ? gtimer() : 0ull;
#define SSTAMP(i) do { if (tv.trace && tid == 0 && blockIdx.x == 5 && blockIdx.y == 0) tv.trace[56 + (i)] += gtimer() - tk0; } while (0)
    const cd tcorh = ld(mb + a.L.tcorh, mx, m, n);
    const cd qcorh_old = ld(mb + a.L.qcorh, mx, m, n);
    const cd qcorh_new = (a.flag & 2) ? ld((mb + a.sout_off + (size_t)(GO_QCORH) * a.sout_fs), mx, m, n) : zero;
    const int do_forcing = (a.flag & 2) ? a.clk->do_forcing : 0;
    const cd psdt_in = ld((mb + a.sout_off + (size_t)(GO_PSDT) * a.sout_fs), mx, m, n);
    const cd phis_q = ld(mb + a.L.phis, mx, m, n);
    // time level j1 of the prognostics for the filter (time_stepping.f90:163-166); j1 == 1 re-uses level 1
    cd vorj = zero, divj = zero, tj = zero, trj = zero, psj = zero;
    auto load_level_j = [&]() {
        vorj = ld(sfield(mb, a.L.vor, nsp, KX + k), mx, m, n);
        divj = ld(sfield(mb, a.L.div, nsp, KX + k), mx, m, n);
        tj = ld(sfield(mb, a.L.t, nsp, KX + k), mx, m, n);
        trj = ld(sfield(mb, a.L.tr, nsp, KX + k), mx, m, n);
        psj = ld(sfield(mb, a.L.ps, nsp, 1), mx, m, n);
    };
    if (!BATCH && a.j1 != 1) load_level_j();     // the batch variant fetches them in front of the filter (fewer values live across the solve)

    // ---- tendencies.f90:212-234: spectral assembly of the transformed grid-point tendencies
    cd vordt, divdt, tdt, trdt, psdt = zero;
    {
        const int f = GO_PER * k;
        // all stencil operands of the three vds calls are loaded unconditionally (neighbour index clamped at the edges, where the
        // value is not used) ahead of the n-dependent branches: one round trip instead of one per call
        const int nm = n > 0 ? n - 1 : 0, np = n < nx - 1 ? n + 1 : nx - 1;
        const double gx = tv.gradx[m], dym = tv.vddym[q], dyp = tv.vddyp[q];
        const double* F0 = (mb + a.sout_off + (size_t)(f + 0) * a.sout_fs); const double* F1 = (mb + a.sout_off + (size_t)(f + 1) * a.sout_fs);
        const double* F4 = (mb + a.sout_off + (size_t)(f + 4) * a.sout_fs); const double* F7 = (mb + a.sout_off + (size_t)(f + 7) * a.sout_fs);
        const cd um = ld(F0, mx, m, nm), u0 = ld(F0, mx, m, n), up = ld(F0, mx, m, np);
        const cd vm = ld(F1, mx, m, nm), v0 = ld(F1, mx, m, n), vp = ld(F1, mx, m, np);
        const cd ke = ld((mb + a.sout_off + (size_t)(f + 2) * a.sout_fs), mx, m, n);
        const cd ut0 = ld((mb + a.sout_off + (size_t)(f + 3) * a.sout_fs), mx, m, n), vtm = ld(F4, mx, m, nm), vtp = ld(F4, mx, m, np);
        const cd tt = ld((mb + a.sout_off + (size_t)(f + 5) * a.sout_fs), mx, m, n);
        const cd uq0 = ld((mb + a.sout_off + (size_t)(f + 6) * a.sout_fs), mx, m, n), vqm = ld(F7, mx, m, nm), vqp = ld(F7, mx, m, np);
        const cd qt = ld((mb + a.sout_off + (size_t)(f + 8) * a.sout_fs), mx, m, n);
        cd vo, di;
        dev_vds_r(nx, n, gx, dym, dyp, um, u0, up, vm, v0, vp, vo, di);
        vordt = vo;
        divdt = di - neg(el2 * ke);                                      // - laplacian(KE)
        tdt = dev_vds_div_r(nx, n, gx, dym, dyp, ut0, vtm, vtp) + tt;
        trdt = dev_vds_div_r(nx, n, gx, dym, dyp, uq0, vqm, vqp) + qt;
    }
    // time level 1 of the prognostics (all linear terms use it, alph = 0.5: tendencies.f90:32)
    const cd vor1 = ld(sfield(mb, a.L.vor, nsp, k), mx, m, n);
    const cd div1 = ld(sfield(mb, a.L.div, nsp, k), mx, m, n);
    const cd t1 = ld(sfield(mb, a.L.t, nsp, k), mx, m, n);
    const cd tr1 = ld(sfield(mb, a.L.tr, nsp, k), mx, m, n);
    const cd ps1 = ld(sfield(mb, a.L.ps, nsp, 0), mx, m, n);
    mbar_wait(bar, 0);
    SSTAMP(0);
    s_a[k][c] = div1;
    s_b[k][c] = t1;
    __syncthreads();
    // ---- get_spectral_tendencies  tendencies.f90:242-293 (level-coupled parts: one warp each)
    if (k == 0) {
        psdt = psdt_in;
        if (rr == 0) psdt = zero;                                        // tendencies.f90:126
        cd dmeanc = zero;
#pragma unroll
        for (int kk = 0; kk < KX; kk++) dmeanc = dmeanc + lc.dhs[kk] * s_a[kk][c];
        psdt = psdt - dmeanc;
        if (rr == 0) psdt = zero;
        s_dmeanc[c] = dmeanc;
        s_psdt[c] = psdt;
        cd sg = zero;
        s_sig[0][c] = zero;
#pragma unroll
        for (int kk = 0; kk < KX - 1; kk++) { sg = sg - lc.dhs[kk] * (s_a[kk][c] - dmeanc); s_sig[kk + 1][c] = sg; }
        s_sig[KX][c] = zero;
    } else if (k == 1) {
        // get_geopotential(t(:,:,:,1), phis)  geopotential.f90:33-57 (tendencies.f90:288)
        cd ph = phis_q + lc.xgeop1[KX - 1] * s_b[KX - 1][c];
        s_phi[KX - 1][c] = ph;
#pragma unroll
        for (int kk = KX - 2; kk >= 0; kk--) { ph = (ph + lc.xgeop2[kk + 1] * s_b[kk + 1][c]) + lc.xgeop1[kk] * s_b[kk][c]; s_phi[kk][c] = ph; }
        if (m == 0) {
#pragma unroll
            for (int kk = 1; kk < KX - 1; kk++) s_phi[kk][c] = s_phi[kk][c] + lc.geop_corf[kk] * (s_b[kk + 1][c] - s_b[kk - 1][c]);
        }
    }
    __syncthreads();
    {
        const cd dmeanc = s_dmeanc[c];
        const cd sg0 = s_sig[k][c], sg1 = s_sig[k + 1][c];
        const cd dk0 = (k >= 1) ? (lc.tref[k] - lc.tref[k - 1]) * sg0 : zero;
        const cd dk1 = (k + 1 <= KX - 1) ? (lc.tref[k + 1] - lc.tref[k]) * sg1 : zero;
        tdt = ((tdt - lc.dhsr[k] * (dk1 + dk0)) + lc.tref3[k] * (sg1 + sg0)) - lc.tref2[k] * dmeanc;
        const cd x = s_phi[k][c] + (lc.rgas * lc.tref[k]) * ps1;
        divdt = divdt - neg(el2 * x);
        if (valid) st(sfield(mb, a.L.phi, nsp, k), mx, m, n, s_phi[k][c]);   // module phi (tendencies.f90:288), read by output()
    }
    __syncthreads();          // s_a (div1) and s_b (t1) are free again
    SSTAMP(1);
    // ---- implicit_terms  implicit.f90:168-217
    s_a[k][c] = tdt;
    __syncthreads();
    {
        cd s = zero;
#pragma unroll
        for (int k1 = 0; k1 < KX; k1++) s = s + sXd[k + KX * k1] * s_a[k1][c];
        const cd ye = s + lc.tref1[k] * s_psdt[c];
        s_b[k][c] = divdt + elz_q * ye;       // yf
    }
    __syncthreads();
    divdt = zero;
    if (m + n != 0) {
        const double* xj = sXj + (m + n - 1);                      // xj(:,:,l) transposed, l = total wavenumber
#pragma unroll
        for (int k1 = 0; k1 < KX; k1++) divdt = divdt + xj[(size_t)(k + KX * k1) * nl] * s_b[k1][c];
    }
    __syncthreads();          // all reads of tdt (s_a) done
    s_a[k][c] = divdt;
    __syncthreads();
    if (k == 0) {
#pragma unroll
        for (int kk = 0; kk < KX; kk++) psdt = psdt - lc.dhsx[kk] * s_a[kk][c];
    }
#pragma unroll
    for (int k1 = 0; k1 < KX; k1++) tdt = tdt + sXc[k + KX * k1] * s_a[k1][c];

    if (a.flag & 1) {
        if (valid) {
            st(sfield(mb, a.L.vordt, nsp, k), mx, m, n, vordt);
            st(sfield(mb, a.L.divdt, nsp, k), mx, m, n, divdt);
            st(sfield(mb, a.L.tdt, nsp, k), mx, m, n, tdt);
            st(sfield(mb, a.L.trdt, nsp, k), mx, m, n, trdt);
            if (k == 0) st(mb + a.L.psdt, mx, m, n, psdt);
        }
        return;
    }
    SSTAMP(2);
    // ---- horizontal diffusion + drag  time_stepping.f90:63-96
    {
        cd qcorh;
        if (do_forcing) {       // qcorh = grid_to_spec(corh) of today's set_forcing (forcing.f90:99)
            qcorh = qcorh_new;
            if (k == 0 && valid) st(mb + a.L.qcorh, mx, m, n, qcorh);
        } else {
            qcorh = qcorh_old;
        }
        if (BATCH) { dmp = tv.dmp[q]; dmpd = tv.dmpd[q]; dmps = tv.dmps[q]; dmp1 = tv.dmp1[q]; dmp1d = tv.dmp1d[q]; dmp1s = tv.dmp1s[q]; }
        vordt = dmp1 * (vordt - dmp * vor1);
        divdt = dmp1d * (divdt - dmpd * div1);
        const cd ctmp = t1 + lc.tcorv[k] * tcorh;
        tdt = dmp1 * (tdt - dmp * ctmp);
        if (k == 0 && m == 0) {
            vordt = vordt - lc.sdrag * vor1;
            divdt = divdt - lc.sdrag * div1;
        }
        vordt = dmp1s * (vordt - dmps * vor1);
        divdt = dmp1s * (divdt - dmps * div1);
        tdt = dmp1s * (tdt - dmps * ctmp);
        const cd qtmp = tr1 + lc.qcorv[k] * qcorh;
        trdt = dmp1d * (trdt - dmpd * qtmp);
    }
    // ---- step_field_2d  time_stepping.f90:141-167
    cd vor2n = zero, div2n = zero, t2n = zero, t1n = zero;
    if (BATCH && a.j1 != 1) load_level_j();
    {
        const double eps = (a.j1 == 1) ? 0.0 : lc.rob;
        const double trf = trf_q;
        const double c1 = lc.wil * eps, c2 = (1.0 - lc.wil) * eps;
        auto stepf = [&](long long off, int nlev_fields, int kk, cd fdt, cd f1, cd fj_in, cd* new_level1 = nullptr) -> cd {
            double* p1 = sfield(mb, off, nsp, kk);
            double* p2 = sfield(mb, off, nsp, nlev_fields + kk);
            fdt = trf * fdt;
            const cd fj = (a.j1 == 1) ? f1 : fj_in;
            const cd fnew = f1 + a.dt * fdt;
            const cd f1n = fj + c1 * ((f1 - 2.0 * fj) + fnew);
            const cd fj2 = (a.j1 == 1) ? f1n : fj;     // :166 re-reads output(:,:,j1) after :163 overwrote level 1
            const cd f2n = fnew - c2 * ((f1n - 2.0 * fj2) + fnew);
            if (valid) { st(p1, mx, m, n, f1n); st(p2, mx, m, n, f2n); }
            if (new_level1) *new_level1 = f1n;
            return f2n;
        };
        if (k == 0) stepf(a.L.ps, 1, 0, psdt, ps1, psj);
        vor2n = stepf(a.L.vor, KX, k, vordt, vor1, vorj);
        div2n = stepf(a.L.div, KX, k, divdt, div1, divj);
        t2n = stepf(a.L.t, KX, k, tdt, t1, tj, &t1n);
        stepf(a.L.tr, KX, k, trdt, tr1, trj);
    }
    SSTAMP(3);
    if (!(a.flag & 2)) return;
    // ---- geopotential of the NEW time level 1: the field the next step's physics transforms (physics.f90:103,
    // tendencies.f90:203); the module variable phi above keeps the reference's one-step-old value for output()
    __syncthreads();
    s_b[k][c] = t1n;
    __syncthreads();
    if (k == 1) {
        cd ph = phis_q + lc.xgeop1[KX - 1] * s_b[KX - 1][c];
        s_phi[KX - 1][c] = ph;
#pragma unroll
        for (int kk = KX - 2; kk >= 0; kk--) { ph = (ph + lc.xgeop2[kk + 1] * s_b[kk + 1][c]) + lc.xgeop1[kk] * s_b[kk][c]; s_phi[kk][c] = ph; }
        if (m == 0) {
#pragma unroll
            for (int kk = 1; kk < KX - 1; kk++) s_phi[kk][c] = s_phi[kk][c] + lc.geop_corf[kk] * (s_b[kk + 1][c] - s_b[kk - 1][c]);
        }
    }
    __syncthreads();
    if (valid) st(sfield(mb, a.L.phi_next, nsp, k), mx, m, n, s_phi[k][c]);
    SSTAMP(4);
    // ---- check_diagnostics on the new time level 2 (diagnostics.f90:16-75): per-block partial sums
    {
        double s1 = 0.0, s2 = 0.0;
        if (valid && m >= 1) {
            const double e = elm2_q;
            const cd tv_ = neg(e * vor2n), td = neg(e * div2n);
            s1 = -(tv_.re * vor2n.re - tv_.im * (-vor2n.im));
            s2 = -(td.re * div2n.re - td.im * (-div2n.im));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
        double* part = a.partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (2 * KX);
        if (c == 0) { part[k] = s1; part[KX + k] = s2; }
        if (valid && rr == 0) a.partial[(size_t)gridDim.y * gridDim.x * 2 * KX + (size_t)blockIdx.y * KX + k] = t2n.re;
    }
    // ---- the step is closed (final diagnostics reduction, range guard, calendar) by an extra CTA of the next
    // step's spec->grid kernel, or by k_close_step when no step follows: nothing here waits for the other blocks
    if (blockIdx.x == 0 && blockIdx.y == 0 && c == 0 && k == 0) a.clk->close_pending = 1;
    SSTAMP(5);
#undef SSTAMP
    if (tv.trace) { __syncthreads(); if (tid == 0) trace_end(tv.trace, 3); }
}

__global__ void k_close_step(CloseArgs cl) { close_step_cta(cl, threadIdx.x); }

// check_diagnostics (diagnostics.f90:16-75): one block per level, tree reduction over (m,n)
__global__ void k_diagnostics(SpecArgs a) {
    const int mx = a.tv.mx, nx = a.tv.nx, nsp = mx * nx;
    const int k = blockIdx.x;
    const double* mb = a.base;    // member 0 drives the guard; other members: blockIdx.y
    mb += (size_t)blockIdx.y * a.stride;
    const int lev = a.j2 - 1;
    const double* vor = sfield(mb, a.L.vor, nsp, lev * KX + k);
    const double* div = sfield(mb, a.L.div, nsp, lev * KX + k);
    double s1 = 0.0, s2 = 0.0;
    for (int r = threadIdx.x; r < nsp; r += blockDim.x) {
        const int n = r / mx, m = r - n * mx;
        if (m >= 1) {
            const double e = a.tv.elm2[r];
            const cd v = ld(vor, mx, m, n), d = ld(div, mx, m, n);
            const cd tv_ = neg(e * v), td = neg(e * d);
            s1 -= tv_.re * v.re - tv_.im * (-v.im);
            s2 -= td.re * d.re - td.im * (-d.im);
        }
    }
    __shared__ double sh1[256], sh2[256];
    sh1[threadIdx.x] = s1; sh2[threadIdx.x] = s2;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) { sh1[threadIdx.x] += sh1[threadIdx.x + s]; sh2[threadIdx.x] += sh2[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double d1 = sh1[0], d2 = sh2[0];
        const double d3 = (double)sqrtf(0.5f) * sfield(mb, a.L.t, nsp, lev * KX + k)[0];
        const bool bad = !(d1 <= 500.0) || !(d2 <= 500.0) || !(d3 >= 180.0) || !(d3 <= 320.0);
        if (blockIdx.y == 0) {
            a.clk->diag[k] = d1; a.clk->diag[KX + k] = d2; a.clk->diag[2 * KX + k] = d3;
        }
        if (bad) atomicCAS(&a.clk->diag_fail, 0, a.clk->model_step > 0 ? a.clk->model_step : 1);
    }
}

// output() conversions input_output.f90:201-206 from the grid fields of the level-1 transform set
__global__ void k_output(const double* __restrict__ gin, int N, double grav, double p0, float* __restrict__ out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= N) return;
    for (int k = 0; k < KX; k++) {
        out[(size_t)(0 * KX + k) * N + q] = (float)gin[(size_t)(GI_U1 + k) * N + q];
        out[(size_t)(1 * KX + k) * N + q] = (float)gin[(size_t)(GI_V1 + k) * N + q];
        out[(size_t)(2 * KX + k) * N + q] = (float)gin[(size_t)(GI_T1 + k) * N + q];
        out[(size_t)(3 * KX + k) * N + q] = (float)(gin[(size_t)(GI_Q1 + k) * N + q] * (double)1.0e-3f);
        out[(size_t)(4 * KX + k) * N + q] = (float)(gin[(size_t)(GI_PHI + k) * N + q] / grav);
    }
    out[(size_t)(5 * KX) * N + q] = (float)(p0 * exp(gin[(size_t)GI_PSL * N + q]));
}

// sum and sum of squares over this context's members of the 41 output levels, in the units of
// output() (input_output.f90:201-206: u, v, t, q*1e-3, phi/grav, p0*exp(ps)) but kept in fp64;
// members are added in index order so that the partial sums are reproducible
__global__ void k_ensemble_sums(const double* __restrict__ base, long long stride, long long gin_off, int nmembers, int nvals, int N,
                                double grav, double p0, double* __restrict__ sum, double* __restrict__ sumsq) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nvals) return;
    const int f = q / N;   // 0..7 u, 8..15 v, 16..23 t, 24..31 q, 32..39 phi, 40 ps
    double s = 0.0, s2 = 0.0;
    for (int e = 0; e < nmembers; e++) {
        double v = base[(size_t)e * stride + gin_off + q];
        if (f >= 40) v = p0 * exp(v);
        else if (f >= 32) v = v / grav;
        else if (f >= 24) v = v * (double)1.0e-3f;
        s += v; s2 += v * v;
    }
    sum[q] = s; sumsq[q] = s2;
}

struct SpptArgs { double* base; long long stride; Layout L; DevTables tv; SpptParams p; };
__global__ void k_sppt_update(SpptArgs a) {
    // nothing here reads what the previous kernel writes: dependents may launch at once; the wait at the END keeps the chain
    // transitive (the next kernel's wait on this one implies everything before it)
    pdl_trigger();
    const int nsp = a.tv.mx * a.tv.nx;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int counter = a.p.state[0];
    sppt_point(a.base + (size_t)blockIdx.y * a.stride, a.L, a.tv, a.p, counter, blockIdx.y, t, t < KX * nsp);
    __syncthreads();
    if (threadIdx.x == 0) sppt_ticket(a.p.state, counter, gridDim.x * gridDim.y);
    pdl_wait();
}

// ---- launchers -------------------------------------------------------------------------------
static SpecArgs spec_args(speedy_ctx* ctx) {
    Model& M = *ctx->model;
    SpecArgs a;
    a.base = M.mem.p; a.stride = M.L.stride; a.L = M.L; a.tv = ctx->dv; a.lc = M.lc.p; a.clk = M.clock.p;
    a.j1 = 1; a.j2 = 1; a.dt = 0.0; a.flag = 0; a.partial = M.diag_partial.p; a.ready_reset = nullptr; a.sppt.state = nullptr;
    a.sout_off = M.L.sout; a.sout_fs = 2LL * ctx->d.nspec();
    return a;
}

void launch_geopotential(speedy_ctx* ctx, int which) {
    SpecArgs a = spec_args(ctx);
    a.flag = which;
    dim3 grid((ctx->d.nspec() + 63) / 64, ctx->nmembers);
    k_geopotential<<<grid, 64, 0, ctx->stream>>>(a);
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
}

static SpptParams sppt_params(speedy_ctx* ctx);
void launch_spec_step(speedy_ctx* ctx, int j1, int j2, double dt, int store_tend_only, int close_step) {
    Model& M = *ctx->model;
    SpecArgs a = spec_args(ctx);
    a.j1 = j1; a.j2 = j2; a.dt = dt; a.flag = (store_tend_only ? 1 : 0) | (close_step ? 2 : 0);
    if (close_step && M.ready_target) a.ready_reset = M.ready.p;
    // SPPT with device-drawn noise: the pattern of the next get_tendencies call is prepared here (model.cu sppt_next consumes it)
    if (ctx->sppt_on && M.sppt_draw && ctx->sppt_fold) { a.sppt = sppt_params(ctx); M.sppt_prepared = true; }
    if (close_step && M.alias_active) { a.sout_off = M.L.gin; a.sout_fs = ctx->d.ngrid(); }
    dim3 grid((ctx->d.nspec() + SC - 1) / SC, ctx->nmembers);
    const size_t need = (size_t)grid.x * grid.y * 2 * KX + (size_t)grid.y * KX;
    if (M.diag_partial.n < need) { M.diag_partial.alloc(need); a.partial = M.diag_partial.p; }
    const size_t smem = spec_step_smem(ctx->d.mx, ctx->d.nx);
    if (ctx->nmembers >= 4) CUDA_CHECK(launch_pdl(ctx->dv.trace == nullptr || ctx->trace_pdl, k_spec_step<true>, grid, dim3(SC, KX), smem, ctx->stream, a));
    else CUDA_CHECK(launch_pdl(ctx->dv.trace == nullptr || ctx->trace_pdl, k_spec_step<false>, grid, dim3(SC, KX), smem, ctx->stream, a));
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
}

// per device (speedy_create calls it after cudaSetDevice): the shared-memory opt-in is a per-device function attribute
void setup_spec_step_kernels() {
    CUDA_CHECK(cudaFuncSetAttribute(k_spec_step<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CUDA_CHECK(cudaFuncSetAttribute(k_spec_step<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
}

// Stand-alone forms of two operators that the main loop only runs fused inside k_spec_step, for the operator-level drop-ins of the
// C ABI (implicit.f90:168-217 implicit_terms, horizontal_diffusion.f90:86-105 do_horizontal_diffusion): one thread per (m,n),
// the sums in the reference's order.
__global__ void k_implicit_terms(double* __restrict__ divdt_, double* __restrict__ tdt_, double* __restrict__ psdt_, DevTables tv, const LevelConsts* __restrict__ lcp) {
    const int mx = tv.mx, nx = tv.nx, nsp = mx * nx;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nsp) return;
    const int n = r / mx, m = r - n * mx;
    const LevelConsts& lc = *lcp;
    const cd zero{0.0, 0.0};
    cd tdt[KX], divdt[KX], ye[KX], yf[KX];
#pragma unroll
    for (int k = 0; k < KX; k++) { tdt[k] = ld(tdt_ + (size_t)2 * nsp * k, mx, m, n); divdt[k] = ld(divdt_ + (size_t)2 * nsp * k, mx, m, n); ye[k] = zero; }
    cd psdt = ld(psdt_, mx, m, n);
#pragma unroll
    for (int k1 = 0; k1 < KX; k1++)
#pragma unroll
        for (int k = 0; k < KX; k++) ye[k] = ye[k] + tv.xd[k + KX * k1] * tdt[k1];
    const double elz = tv.elz[m + (size_t)mx * n];
#pragma unroll
    for (int k = 0; k < KX; k++) { ye[k] = ye[k] + lc.tref1[k] * psdt; yf[k] = divdt[k] + elz * ye[k]; divdt[k] = zero; }
    if (m + n != 0) {
        const double* xj = tv.xj + (size_t)KX * KX * (m + n - 1);      // xj(:,:,l), l = m + n - 2 in the reference's 1-based indices
#pragma unroll
        for (int k1 = 0; k1 < KX; k1++)
#pragma unroll
            for (int k = 0; k < KX; k++) divdt[k] = divdt[k] + xj[k + KX * k1] * yf[k1];
    }
#pragma unroll
    for (int k = 0; k < KX; k++) psdt = psdt - lc.dhsx[k] * divdt[k];
#pragma unroll
    for (int k = 0; k < KX; k++) {
        cd t = tdt[k];
#pragma unroll
        for (int k1 = 0; k1 < KX; k1++) t = t + tv.xc[k + KX * k1] * divdt[k1];
        st(tdt_ + (size_t)2 * nsp * k, mx, m, n, t);
        st(divdt_ + (size_t)2 * nsp * k, mx, m, n, divdt[k]);
    }
    st(psdt_, mx, m, n, psdt);
}

__global__ void k_horizontal_diffusion(const double* __restrict__ field, double* __restrict__ fdt, const double* __restrict__ dmp, const double* __restrict__ dmp1,
                                       int nsp, int nlev) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nsp * nlev) return;
    const int q = t % nsp;
    const cd f{field[2 * (size_t)t], field[2 * (size_t)t + 1]}, d{fdt[2 * (size_t)t], fdt[2 * (size_t)t + 1]};
    const cd o = dmp1[q] * (d - dmp[q] * f);      // (fdt_in - dmp*field)*dmp1
    fdt[2 * (size_t)t] = o.re; fdt[2 * (size_t)t + 1] = o.im;
}

void launch_implicit_terms(speedy_ctx* ctx, double* d_divdt, double* d_tdt, double* d_psdt) {
    k_implicit_terms<<<(ctx->d.nspec() + 63) / 64, 64, 0, ctx->stream>>>(d_divdt, d_tdt, d_psdt, ctx->dv, ctx->model->lc.p);
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
}

void launch_horizontal_diffusion(speedy_ctx* ctx, const double* d_field, double* d_fdt, const double* d_dmp, const double* d_dmp1, int nlev) {
    const int total = ctx->d.nspec() * nlev;
    k_horizontal_diffusion<<<(total + 127) / 128, 128, 0, ctx->stream>>>(d_field, d_fdt, d_dmp, d_dmp1, ctx->d.nspec(), nlev);
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
}

void launch_close_step(speedy_ctx* ctx) {
    Model& M = *ctx->model;
    CloseArgs cl{M.clock.p, M.diag_partial.p, (int)((ctx->d.nspec() + SC - 1) / SC), ctx->nmembers, ctx->dv.trace};
    k_close_step<<<1, 256, 0, ctx->stream>>>(cl);
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
}

void launch_diagnostics(speedy_ctx* ctx, int level) {
    SpecArgs a = spec_args(ctx);
    a.j2 = level;
    dim3 grid(KX, ctx->nmembers);
    k_diagnostics<<<grid, 256, 0, ctx->stream>>>(a);
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
}

void launch_output_convert(speedy_ctx* ctx, int member, float* d_out) {
    Model& M = *ctx->model;
    const int N = ctx->d.ngrid();
    k_output<<<(N + 127) / 128, 128, 0, ctx->stream>>>(M.mem.p + (size_t)member * M.L.stride + M.L.gin, N, ctx->tab.c.grav, ctx->tab.c.p0, d_out);
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
}

static SpptParams sppt_params(speedy_ctx* ctx) {
    Model& M = *ctx->model;
    SpptParams p;
    p.seed = ctx->seed; p.state = M.sppt_state.p; p.member0 = ctx->member_offset; p.draw = M.sppt_draw ? 1 : 0;
    {   // sppt.f90:32,76-80 in the reference's order of operations
        const double time_decorr = 6.0, len_decorr = 500000.0, stddev = (double)0.33f;
        p.phi = exp(-(24 / (double)ctx->tab.c.nsteps) / time_decorr);
        double f0 = 0.0;
        const double rr = len_decorr / ctx->tab.c.rearth;
        for (int nn = 1; nn <= ctx->d.trunc; nn++) f0 = f0 + (2 * nn + 1) * exp(-0.5 * (rr * rr) * nn * (nn + 1));
        p.f0 = sqrt(((stddev * stddev) * (1 - p.phi * p.phi)) / (2 * f0));
    }
    return p;
}

void launch_sppt_update(speedy_ctx* ctx) {
    Model& M = *ctx->model;
    SpptArgs a;
    a.base = M.mem.p; a.stride = M.L.stride; a.L = M.L; a.tv = ctx->dv; a.p = sppt_params(ctx);
    const int total = KX * ctx->d.nspec();
    dim3 grid((total + 127) / 128, ctx->nmembers);
    CUDA_CHECK(launch_pdl(ctx->dv.trace == nullptr || ctx->trace_pdl, k_sppt_update, grid, dim3(128), 0, ctx->stream, a));
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
}

void launch_ensemble_sums(speedy_ctx* ctx, double* d_sum, double* d_sumsq) {
    Model& M = *ctx->model;
    const int nvals = 41 * ctx->d.ngrid();
    k_ensemble_sums<<<(nvals + 255) / 256, 256, 0, ctx->stream>>>(M.mem.p, M.L.stride, M.L.gin + (long long)GI_U1 * ctx->d.ngrid(), ctx->nmembers, nvals, (int)ctx->d.ngrid(), ctx->tab.c.grav, ctx->tab.c.p0, d_sum, d_sumsq);
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace spd
