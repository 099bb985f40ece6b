// K4 + K5 — the grid-point column kernel: non-linear dynamics tendencies
// (tendencies.f90:109-197) and the whole physics sweep (physics.f90:110-205 dispatching
// convection.f90, large_scale_condensation.f90, shortwave_radiation.f90,
// longwave_radiation.f90, surface_fluxes.f90, vertical_diffusion.f90) fused into ONE pass
// over the (ix,il) columns: every parameterisation is column-local, so a CTA owns a tile of 32
// columns, stages it once in shared memory (tensor-map TMA) and works on it level-parallel
// (see "the column kernel" below); HBM is touched once per input field and once per output
// field.  Longitude is the fastest index = unit stride in every array.
// Also here: the per-step land/sea slab update (land_model.f90:184-239,
// sea_model.f90:253-444), the daily forcing (forcing.f90:55-99) and the device calendar.
//
// Compiled with --fmad=false so that the arithmetic is the reference's operation by
// operation (the checker is built with -ffp-contract=off).
#include "model.h"
#include "calendar.h"
#include "tma.cuh"
#include "member_ready.cuh"
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

namespace spd {

#define KX 8
#define F32(x) ((double)(x##f))

__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }   // Fortran min/max on non-NaN data
__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }

// The column kernel evaluates exp ~60 times per column; inlined, those copies are most of its
// 12k-instruction body and it stalls on instruction fetch (ncu: 40 % "no instruction").  One
// shared out-of-line copy keeps the body inside the instruction cache.  Same libdevice routine.
__device__ __noinline__ double exp_ol(double x) { return exp(x); }
#ifndef SPD_EXP_INLINE
#define exp(x) exp_ol(x)
#endif

// nint(T) -> row of fband(100:400,:) (longwave_radiation.f90:84,99).  The clamp only matters once the model has left its
// accepted range (diagnostics.f90:60-70): the reference would index out of bounds there, this must not fault the GPU.
__device__ __forceinline__ int band_row(double t) { const int n = (int)round(t) - 100; return min(max(n, 0), 300); }

// humidity.f90:44-78 for one point; p = sig*ps (or ps(1,1) for sig <= 0)
__device__ __forceinline__ double qsat_pt(double ta, double p) {
    const double e0 = 6.108e-3, c1 = F32(17.269), c2 = F32(21.875), t0 = F32(273.16), t1 = F32(35.86), t2 = F32(7.66);
    const bool warm = ta >= t0;
    const double q = e0 * exp((warm ? c1 : c2) * (ta - t0) / (ta - (warm ? t1 : t2)));
    return 622.0 * q / (p - F32(0.378) * q);
}

__device__ __forceinline__ void slab_point(double* mb, const Layout& L_, const SharedDev& sh_, const DevClock& c, const LevelConsts& lc, int N, int q, int day0);
__device__ __forceinline__ void forcing_point(double* mb, const Layout& L_, const SharedDev& sh_, const DevClock& clk, const LevelConsts& lc, int ix, int il, int q);

struct ColumnArgs {
    double* base; long long stride;
    int* ibase;
    Layout L;
    SharedDev sh;
    int merged;              // 1 (main loop): the pending slab update of the previous step and the daily forcing run here first
    const LevelConsts* lc;
    const DevClock* clk;
    const double* fband;     // (301,4) Fortran order
    const double* coriol;    // il
    const double* coa;       // il
    int ix, il;
    int mode;                // 0: dynamics + physics -> K2 inputs; 1: physics only, tendencies in/out in gout slots
    int csw_override;        // -1: take compute_shortwave from the device clock
    int sppt_on;
    unsigned long long* trace;
    int discard_row0;        // first grid-field row to drop (the rows below it are overwritten by this kernel's own output when the transient buffer is shared)
    int discard_gin;         // main-loop step: the grid fields are dead once the tile is staged (the next step's transform rewrites them): their L2 lines are dropped, not written back
    const unsigned* ready;   // main-loop step behind the quad transform: per-member completion counts (member_ready.cuh), else nullptr
    unsigned ready_target;
    // [member][row][column] tensor maps: box = rows x 32 columns, one TMA instruction per tile
    CUtensorMap m_dyn, m_phys, m_tau2, m_stratc, m_rsw;
};

// ---- the column kernel -------------------------------------------------------------------
// One CTA owns a tile of TC = 32 consecutive columns (lane = column).  The tile is staged once in
// shared memory by bulk asynchronous copies (TMA, one 256-byte row per field, two mbarriers) and
// worked on by TEN warps:
//   * eight LEVEL warps, warp k <-> sigma level k.  Everything that is local to a level — the
//     thermodynamic prep (qsat, rh, static energy), large-scale condensation, the long-wave source
//     terms, the whole of tendencies.f90:109-197, the short/long-wave transmissivities (all the exp()
//     of the radiation) and the closing sum of the tendencies — runs level-parallel, so the dependent
//     FP64 chain of a column is 1/8 as long as in a thread-per-column sweep (a B200 DFMA has 8.7 cycles of
//     dependent-issue latency, exp() 160: profiles/r1g).  The vertical sweeps that are inherently serial
//     (convection, the radiative flux recurrences, the surface-flux balance) run on level warp 1 (the sea
//     half of the surface fluxes on level warp 2) between the wide phases;
//   * SLAB: couple_sea_land of the previous step + set_forcing(1) when due, then stages the surface fields;
//   * VDIF: vertical_diffusion.f90 (column-serial, off the critical path).
// Hand-offs are named barriers; every sum keeps the reference's order of operations.
constexpr int TC = 32;
enum { W_SLAB = KX, W_VDIF = KX + 1, COL_WARPS = KX + 2 };
constexpr int COL_THREADS = COL_WARPS * 32;
constexpr int LEV_THREADS = KX * 32;
enum { BAR_LEV = 1, BAR_MID = 2, BAR_SEA = 3, BAR_END = 4 };
enum { SF_FMASK, SF_FSOL, SF_OZONE, SF_OZUPP, SF_ZENIT, SF_STRATZ, SF_ALBSFC, SF_PHIS0, SF_SST, SF_STL, SF_SOILW, SF_ALBL, SF_ALBS,
       SF_SNOWC, SF_FOROG, SF_SSRD, SF_N };
// scalar rows (one value per column)
enum { S_PSG, S_CLOUDC, S_QCLOUD, S_T1_1, S_T1_2, S_T2_1, S_DENVVS0, S_T0, S_U0, S_V0, S_USTR2, S_VSTR2, S_SHF2, S_EVAP2, S_SLRU2,
       S_UT8, S_VT8, S_SHFT, S_EVAPT, S_SLRD, S_FLX1, S_FLX2, S_FLX3, S_FLX4, S_N };
enum { I_ICNV, I_ICLTOP, I_LSC, I_N = I_LSC + KX };
// shared-memory rows of TC doubles
enum { R_GIN = 0, R_TAU2 = R_GIN + GI_N, R_STRATC = R_TAU2 + 4 * KX, R_RSW = R_STRATC + 2, R_SURF = R_RSW + KX, R_DYN = R_SURF + SF_N,
       R_SE = R_DYN + 4 * KX, R_QSAT = R_SE + KX, R_RH = R_QSAT + KX, R_QG = R_RH + KX, R_DFSE = R_QG + KX, R_DFQA = R_DFSE + KX,
       R_DTLSC = R_DFQA + KX, R_DQLSC = R_DTLSC + KX, R_VD = R_DQLSC + KX, R_LW = R_VD + 2 * KX, R_TAU1 = R_LW + 3 * KX,
       R_TAU2S = R_TAU1 + KX, R_SC = R_TAU2S + KX, R_END = R_SC + S_N };
constexpr int NFBAND = 301 * 4;
constexpr int LC_DOUBLES = sizeof(LevelConsts) / sizeof(double);
static_assert(sizeof(LevelConsts) % 16 == 0 && (NFBAND * 8) % 16 == 0, "bulk copies move multiples of 16 bytes");
constexpr size_t COL_SMEM = sizeof(double) * ((size_t)R_END * TC + NFBAND + LC_DOUBLES) + sizeof(int) * TC * I_N + 2 * sizeof(uint64_t);

template <bool BATCH>
__global__ void __launch_bounds__(COL_THREADS, BATCH ? 2 : 1) k_grid_columns(const __grid_constant__ ColumnArgs a) {
    extern __shared__ __align__(128) double smem[];
    double* sFband = smem + (size_t)R_END * TC;
    double* sLc = sFband + NFBAND;
    int* sInt = reinterpret_cast<int*>(sLc + LC_DOUBLES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sInt + TC * I_N);     // [0] physics inputs, [1] dynamics inputs
    const int ix = a.ix, il = a.il, N = ix * il;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) trace_begin(a.trace, 1);
    unsigned long long tk0 = 0ull;
#define STAMP(i) do { if (a.trace && lane == 0 && (warp == 0 || warp >= KX) && blockIdx.x == 40 && blockIdx.y == 0) a.trace[32 + (i)] += gtimer() - tk0; } while (0)
    const int col0 = blockIdx.x * TC, col = col0 + lane;
    const int e = blockIdx.y;
    const int j = col / ix;
    double* mb = a.base + (size_t)e * a.stride;
    int* ib = a.ibase + (size_t)e * a.L.istride;
    const LevelConsts& lc = *reinterpret_cast<const LevelConsts*>(sLc);
    double* gout = mb + a.L.gout;
#define SROW(r) smem[(size_t)(r) * TC + lane]
#define SG(f) SROW(R_GIN + (f))
#define SURF(i) SROW(R_SURF + (i))
#define STAU2(k, b) SROW(R_TAU2 + ((b)-1) * KX + ((k)-1))
#define STRATC(i) SROW(R_STRATC + (i))
#define RSW(k) SROW(R_RSW + (k)-1)
#define DYN(v, k) SROW(R_DYN + (v) * KX + (k)-1)
#define SE(k) SROW(R_SE + (k)-1)
#define QSAT(k) SROW(R_QSAT + (k)-1)
#define RH(k) SROW(R_RH + (k)-1)
#define QG(k) SROW(R_QG + (k)-1)
#define DFSE(k) SROW(R_DFSE + (k)-1)
#define DFQA(k) SROW(R_DFQA + (k)-1)
#define DTLSC(k) SROW(R_DTLSC + (k)-1)
#define DQLSC(k) SROW(R_DQLSC + (k)-1)
#define VD(v, k) SROW(R_VD + (v) * KX + (k)-1)
#define LWS(v, k) SROW(R_LW + (v) * KX + (k)-1)
#define TAU1(k) SROW(R_TAU1 + (k)-1)
#define TAU2S(k) SROW(R_TAU2S + (k)-1)
#define SC(i) SROW(R_SC + (i))
#define SI(i) sInt[(i) * TC + lane]
#define GOUT(f) gout[(size_t)(f) * N + col]
#define G2(off) mb[(off) + col]
#define G3(off, k) mb[(off) + (size_t)((k)-1) * N + col]
// diagnostics nobody on the device reads again (the reference's module variables for output / inspection): streaming stores, first out of L2
#define DIAG2(off, v) __stcs(&mb[(off) + col], (v))
#define TAU2W(k, b, v) do { const double v_ = (v); mb[a.L.tau2 + ((size_t)((b)-1) * KX + ((k)-1)) * N + col] = v_; STAU2(k, b) = v_; } while (0)

    // ---- stage the tile: one bulk copy per field row, spread over the threads -------------------
    const int ngin = a.sppt_on ? GI_N : GI_NBASE;
    const bool want_dyn = a.mode == 0;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    __syncthreads();
    if (tid == 0) {
        const uint32_t row = TC * sizeof(double);
        mbar_expect_tx(&bars[0], (uint32_t)((ngin - GI_U1) + 4 * KX + 2 + KX) * row + NFBAND * 8 + (uint32_t)sizeof(LevelConsts));
        if (want_dyn) mbar_expect_tx(&bars[1], (uint32_t)GI_U1 * row);
        // constant tables first: they may be fetched while the previous kernel is still draining (PDL)
        bulk_g2s(sFband, a.fband, NFBAND * 8, &bars[0]);
        bulk_g2s(sLc, a.lc, (uint32_t)sizeof(LevelConsts), &bars[0]);
    }
    if (a.ready) {                                     // this member's grid fields are complete (the transform may still be storing later members)
        if (tid == 0) ready_wait(a.ready + e, a.ready_target);
        __syncthreads();
    } else {
        pdl_wait();                                    // the grid fields of the previous kernel are complete
    }
    pdl_trigger();
    if (a.trace) tk0 = gtimer();
    if (tid == 0) {     // five tile copies (UTMALDG) instead of ~140 row copies: small bulk copies drain slowly through the TMA unit
        tensor_g2s_3d(&smem[(size_t)(R_GIN + GI_U1) * TC], &a.m_phys, col0, GI_U1, e, &bars[0]);
        tensor_g2s_3d(&smem[(size_t)R_TAU2 * TC], &a.m_tau2, col0, 0, e, &bars[0]);
        tensor_g2s_3d(&smem[(size_t)R_STRATC * TC], &a.m_stratc, col0, 0, e, &bars[0]);
        tensor_g2s_3d(&smem[(size_t)R_RSW * TC], &a.m_rsw, col0, 0, e, &bars[0]);
        if (want_dyn) tensor_g2s_3d(&smem[(size_t)R_GIN * TC], &a.m_dyn, col0, 0, e, &bars[1]);
    }

    const double cor = a.coriol[j];
    // ---------------- tendencies.f90:109-197 for one level ----------------
    // Needs the dynamics rows of the tile only, so any warp can run any level: level warp 1 goes straight from the prep into the
    // convection (the head of the tile's serial chain) and the slab warp, done early, takes level 1 in its place; level warp 3 runs
    // its level behind the downward long-wave sweep.  1.0 (one member) to 1.7 us (8 members) off the chain of every tile.
    auto dynamics_level = [&](const int kd) {
        if (a.mode == 0) {
            mbar_wait(&bars[0], 0);                    // every load of the tile has landed before its first store: the K2 inputs may share the rows
            mbar_wait(&bars[1], 0);
            const double px = SG(GI_PX), py = SG(GI_PY);
            double umean = 0.0, vmean = 0.0, dmean = 0.0;
    #pragma unroll
            for (int kk = 1; kk <= KX; kk++) {
                umean = umean + SG(GI_U + kk - 1) * lc.dhs[kk - 1];
                vmean = vmean + SG(GI_V + kk - 1) * lc.dhs[kk - 1];
                dmean = dmean + SG(GI_DIV + kk - 1) * lc.dhs[kk - 1];
            }
            if (kd == 1) GOUT(GO_PSDT) = -umean * px - vmean * py;   // :125
            // sigma-dot and its mean-flow part at the two interfaces of level kd (:140-143 prefix sums, re-done per level)
            double sd_lo = 0.0, sm_lo = 0.0, sd_hi = 0.0, sm_hi = 0.0;    // interfaces kd and kd+1
            double puvk = 0.0;
            {
                double sd = 0.0, sm = 0.0;
    #pragma unroll
                for (int kk = 1; kk <= KX; kk++) {
                    const double puv = (SG(GI_U + kk - 1) - umean) * px + (SG(GI_V + kk - 1) - vmean) * py;
                    if (kk == kd) { sd_lo = sd; sm_lo = sm; puvk = puv; }
                    sd = sd - lc.dhs[kk - 1] * (puv + SG(GI_DIV + kk - 1) - dmean);
                    sm = sm - lc.dhs[kk - 1] * puv;
                    if (kk == kd) { sd_hi = sd; sm_hi = sm; }
                }
            }
            const double ugk = SG(GI_U + kd - 1), vgk = SG(GI_V + kd - 1), tg2k = SG(GI_T + kd - 1), trgk = SG(GI_TR + kd - 1);
            const double divgk = SG(GI_DIV + kd - 1), vorgk = SG(GI_VOR + kd - 1) + cor;   // :103-107
            const double tggk = tg2k - lc.tref[kd - 1];
            // neighbours (level kd-1 for the lower-index interface, kd+1 for the upper one)
            const double ugm = (kd > 1) ? SG(GI_U + kd - 2) : 0.0, ugp = (kd < KX) ? SG(GI_U + kd) : 0.0;
            const double vgm = (kd > 1) ? SG(GI_V + kd - 2) : 0.0, vgp = (kd < KX) ? SG(GI_V + kd) : 0.0;
            const double tggm = (kd > 1) ? SG(GI_T + kd - 2) - lc.tref[kd - 2] : 0.0, tggp = (kd < KX) ? SG(GI_T + kd) - lc.tref[kd] : 0.0;
            const double trgm = (kd > 1) ? SG(GI_TR + kd - 2) : 0.0, trgp = (kd < KX) ? SG(GI_TR + kd) : 0.0;
            // temp(kd) lives on interface kd (kd = 2..kx), temp(1) = temp(kx+1) = 0
            double t_lo, t_hi;
            t_lo = (kd >= 2) ? sd_lo * (ugk - ugm) : 0.0;
            t_hi = (kd < KX) ? sd_hi * (ugp - ugk) : 0.0;
            const double utend = vgk * vorgk - tggk * lc.rgas * px - (t_hi + t_lo) * lc.dhsr[kd - 1];
            t_lo = (kd >= 2) ? sd_lo * (vgk - vgm) : 0.0;
            t_hi = (kd < KX) ? sd_hi * (vgp - vgk) : 0.0;
            const double vtend = -ugk * vorgk - tggk * lc.rgas * py - (t_hi + t_lo) * lc.dhsr[kd - 1];
            t_lo = (kd >= 2) ? sd_lo * (tggk - tggm) + sm_lo * (lc.tref[kd - 1] - lc.tref[kd - 2]) : 0.0;
            t_hi = (kd < KX) ? sd_hi * (tggp - tggk) + sm_hi * (lc.tref[kd] - lc.tref[kd - 1]) : 0.0;
            const double ttend = tggk * divgk - (t_hi + t_lo) * lc.dhsr[kd - 1] + lc.fsgr[kd - 1] * tggk * (sd_hi + sd_lo) +
                                 lc.tref3[kd - 1] * (sm_hi + sm_lo) + lc.akap * (tg2k * puvk - tggk * dmean);
            t_lo = (kd >= 4) ? sd_lo * (trgk - trgm) : 0.0;            // :192 the tracer flux is zeroed at interfaces 2 and 3
            t_hi = (kd < KX && kd + 1 >= 4) ? sd_hi * (trgp - trgk) : 0.0;
            const double qtend = trgk * divgk - (t_hi + t_lo) * lc.dhsr[kd - 1];
            DYN(0, kd) = utend; DYN(1, kd) = vtend; DYN(2, kd) = ttend; DYN(3, kd) = qtend;
            // products for the direct transforms (tendencies.f90:219-232)
            const int f = GO_PER * (kd - 1);
            GOUT(f + 2) = 0.5 * (ugk * ugk + vgk * vgk);
            GOUT(f + 3) = -ugk * tggk;
            GOUT(f + 4) = -vgk * tggk;
            GOUT(f + 6) = -ugk * trgk;
            GOUT(f + 7) = -vgk * trgk;
        } else {
            const int f = GO_PER * (kd - 1);
            DYN(0, kd) = GOUT(f + 0); DYN(1, kd) = GOUT(f + 1); DYN(2, kd) = GOUT(f + 5); DYN(3, kd) = GOUT(f + 8);
        }
    };

    if (warp == W_SLAB) {
        // ===== main loop only: couple_sea_land of the previous step (speedy.f90:53) and set_forcing(1) (speedy.f90:29-32),
        // both column-local, in front of the physics that consumes them
        if (a.merged) {
            const int pending = a.clk->slab_pending, forcing = a.clk->do_forcing;
            if (pending | forcing) mbar_wait(&bars[0], 0);     // level constants
            if (pending) slab_point(mb, a.L, a.sh, *a.clk, lc, N, col, 0);
            if (forcing) forcing_point(mb, a.L, a.sh, *a.clk, lc, ix, il, col);
        }
        SURF(SF_FMASK) = G2(a.L.fmask_l); SURF(SF_FSOL) = G2(a.L.fsol); SURF(SF_OZONE) = G2(a.L.ozone); SURF(SF_OZUPP) = G2(a.L.ozupp);
        SURF(SF_ZENIT) = G2(a.L.zenit); SURF(SF_STRATZ) = G2(a.L.stratz); SURF(SF_ALBSFC) = G2(a.L.albsfc); SURF(SF_PHIS0) = G2(a.L.phis0);
        SURF(SF_SST) = G2(a.L.sst_am); SURF(SF_STL) = G2(a.L.stl_am); SURF(SF_SOILW) = G2(a.L.soilw_am); SURF(SF_ALBL) = G2(a.L.alb_l);
        SURF(SF_ALBS) = G2(a.L.alb_s); SURF(SF_SNOWC) = G2(a.L.snowc); SURF(SF_FOROG) = G2(a.L.forog); SURF(SF_SSRD) = G2(a.L.ssrd);
        STAMP(11);
        named_arrive(BAR_MID, COL_THREADS);
        dynamics_level(1);
        named_arrive(BAR_END, COL_THREADS);
        return;
    }

    if (warp == W_VDIF) {
        mbar_wait(&bars[0], 0);
        if (a.discard_gin) {
            // an ensemble's transient fields (spec->grid output, read once here) are most of what streams through L2 in a step; left
            // alone they are written back to HBM when evicted and push the live state out (profiles/r2c_l2_*.csv)
            if (want_dyn) mbar_wait(&bars[1], 0);
            const char* g0 = reinterpret_cast<const char*>(mb + a.L.gin + col0);
            const int rfirst = want_dyn ? 0 : GI_U1, row0 = rfirst > a.discard_row0 ? rfirst : a.discard_row0, nrow = ngin - row0;
            for (int t = lane; t < 2 * nrow; t += 32)      // a tile row = TC doubles = two lines
                l2_discard_line(g0 + (size_t)(row0 + (t >> 1)) * N * sizeof(double) + (t & 1) * 128);
        }
        named_sync(BAR_MID, COL_THREADS);      // thermodynamic prep (level warps) and icnv (convection) are in shared memory
        // the column's rows are read in place (shared memory): private copies of five 8-level arrays do not fit the 96 registers of the batch kernel
        // ------------------- vertical_diffusion.f90:30-143 -------------------
        {
            const double trshc = 6.0, trvdi = 24.0, trvds = 6.0, redshc = 0.5, rhgrad = 0.5, segrad = F32(0.1);
            const int nl1 = KX - 1;
            double rsig[KX + 1], rsig1[KX + 1], ttenvd[KX + 1], qtenvd[KX + 1];
            const double cshc = lc.dhs[KX - 1] / 3600.0;
            const double cvdi = (lc.sigh[nl1] - lc.sigh[1]) / ((nl1 - 1) * 3600.0);
            const double fshcq = cshc / trshc, fshcse = cshc / (trshc * lc.cp);
            const double fvdiq = cvdi / trvdi, fvdise = cvdi / (trvds * lc.cp);
#pragma unroll
            for (int k = 1; k <= nl1; k++) { rsig[k] = 1.0 / lc.dhs[k - 1]; rsig1[k] = 1.0 / (1.0 - lc.sigh[k]); }
            rsig[KX] = 1.0 / lc.dhs[KX - 1];
#pragma unroll
            for (int k = 1; k <= KX; k++) { ttenvd[k] = 0.0; qtenvd[k] = 0.0; }
            double drh0 = rhgrad * (lc.fsg[KX - 1] - lc.fsg[nl1 - 1]);
            double fvdiq2 = fvdiq * lc.sigh[nl1];
            {
                const double dmse = SE(KX) - SE(nl1) + lc.alhc * (QG(KX) - QSAT(nl1));
                const double drh = RH(KX) - RH(nl1);
                double fcnv = 1.0;
                if (dmse >= 0.0) {
                    if (SI(I_ICNV) > 0) fcnv = redshc;
                    const double fluxse = fcnv * fshcse * dmse;
                    ttenvd[nl1] = fluxse * rsig[nl1];
                    ttenvd[KX] = -fluxse * rsig[KX];
                    if (drh >= 0.0) {
                        const double fluxq = fcnv * fshcq * QSAT(KX) * drh;
                        qtenvd[nl1] = fluxq * rsig[nl1];
                        qtenvd[KX] = -fluxq * rsig[KX];
                    }
                } else if (drh > drh0) {
                    const double fluxq = fvdiq2 * QSAT(nl1) * drh;
                    qtenvd[nl1] = fluxq * rsig[nl1];
                    qtenvd[KX] = -fluxq * rsig[KX];
                }
            }
#pragma unroll
            for (int k = 3; k <= KX - 2; k++) {
                if (lc.sigh[k] > 0.5) {
                    drh0 = rhgrad * (lc.fsg[k] - lc.fsg[k - 1]);
                    fvdiq2 = fvdiq * lc.sigh[k];
                    const double drh = RH(k + 1) - RH(k);
                    if (drh >= drh0) {
                        const double fluxq = fvdiq2 * QSAT(k) * drh;
                        qtenvd[k] = qtenvd[k] + fluxq * rsig[k];
                        qtenvd[k + 1] = qtenvd[k + 1] - fluxq * rsig[k + 1];
                    }
                }
            }
#pragma unroll
            for (int k = 1; k <= nl1; k++) {               // unrolled: static indices keep the column's arrays in registers
                const double se0 = SE(k + 1) + segrad * (SG(GI_PHI + (k) - 1) - SG(GI_PHI + (k + 1) - 1));
                if (SE(k) < se0) {
                    const double fluxse = fvdise * (se0 - SE(k));
                    ttenvd[k] = ttenvd[k] + fluxse * rsig[k];
#pragma unroll
                    for (int k1 = k + 1; k1 <= KX; k1++) ttenvd[k1] = ttenvd[k1] - fluxse * rsig1[k];
                }
            }

#pragma unroll
            for (int k = 1; k <= KX; k++) { VD(0, k) = ttenvd[k]; VD(1, k) = qtenvd[k]; }
        }
        STAMP(12);
        named_arrive(BAR_END, COL_THREADS);
        return;
    }


    // ======================= level warps: warp k-1 <-> sigma level k =======================
    const int k = warp + 1;
    const int nl1 = KX - 1, nlp = KX + 1;
    const int csw = (a.csw_override >= 0) ? a.csw_override : a.clk->csw;
    const double coa_j = a.coa[j];
    mbar_wait(&bars[0], 0);
    STAMP(0);
    // ---------------- phase A (wide): thermodynamic prep, physics.f90:110-122 ----------------
    const double tgk = SG(GI_T1 + k - 1), phigk = SG(GI_PHI + k - 1);
    const double psg = exp(SG(GI_PSL));
    const double rps = 1.0 / psg;
    const double qgk = dmax(SG(GI_Q1 + k - 1), 0.0);
    const double sek = lc.cp * tgk + phigk;
    const double qsatk = qsat_pt(tgk, lc.fsg[k - 1] * psg);
    const double rhk = qgk / qsatk;
    SE(k) = sek; QSAT(k) = qsatk; RH(k) = rhk; QG(k) = qgk;
    if (k == 1) SC(S_PSG) = psg;
    // ---------------- large_scale_condensation.f90:33-95, the level-local part ----------------
    {
        const double trlsc = 4.0, rhlsc = F32(0.9), drhlsc = F32(0.1), rhblsc = F32(0.95), qsmax = 10.0;
        const double rtlsc = 1.0 / (trlsc * 3600.0), tfact = lc.tfact;
        const double psa2 = psg * psg;
        double dtl = 0.0, dql = 0.0;
        int hit = 0;
        if (k >= 2) {
            const double sig2 = lc.fsg[k - 1] * lc.fsg[k - 1];
            double rhref = rhlsc + drhlsc * (sig2 - 1.0);
            if (k == KX) rhref = dmax(rhref, rhblsc);
            const double dqmax = qsmax * sig2 * rtlsc;
            const double dqa = rhref * qsatk - qgk;
            if (dqa < 0.0) {
                hit = 1;
                dql = dqa * rtlsc;
                dtl = tfact * dmin(-dql, dqmax * psa2);
            }
        }
        DTLSC(k) = dtl; DQLSC(k) = dql; SI(I_LSC + k - 1) = hit;
    }
    // ---------------- long-wave source terms, longwave_radiation.f90:40-76 ----------------
    {
        const double anis = 1.0;
        const double tgm = (k > 1) ? SG(GI_T1 + k - 2) : 0.0, tgp = (k < KX) ? SG(GI_T1 + k) : 0.0;
        const double si_k = (k <= nl1) ? tgk + lc.wvi[7 + k] * (tgp - tgk) : 0.0;            // st4a(k,1) before the power
        const double si_m = (k >= 2) ? tgm + lc.wvi[7 + k - 1] * (tgk - tgm) : 0.0;          // st4a(k-1,1)
        double s1, s2;
        if (k <= 2) {
            const double x = (k == 1) ? 0.75 * tgk + 0.25 * si_k : 0.50 * tgk + 0.25 * (si_m + si_k);
            s1 = lc.sbc * ((x * x) * (x * x));
            s2 = 0.0;
        } else {
            const double d = (k <= nl1) ? 0.5 * anis * dmax(si_k - si_m, 0.0) : anis * dmax(tgk - si_m, 0.0);
            const double x = tgk;
            const double st3a = lc.sbc * ((x * x) * x);
            s1 = st3a * tgk;
            s2 = 4.0 * st3a * d;
        }
        LWS(0, k) = s1; LWS(1, k) = s2;
    }
    STAMP(8);
    named_sync(BAR_LEV, LEV_THREADS);          // the prep of every level is in shared memory: the sweeps may start

    const double emisfc = F32(0.98), epslw = F32(0.05);
    // level warp 3 sweeps the long wave downward while level warp 1 is busy with convection (or the short-wave sweeps):
    // it needs only the transmissivities, the source terms of phase A and the band table
    auto lw_down = [&]() {
        // ------------------- downward longwave  longwave_radiation.f90:16-117 -------------------
        double st4a1[KX + 1], st4a2[KX + 1], tt_rlw[KX + 1], flux[5];
        double slrd;
#pragma unroll
        for (int kk = 1; kk <= KX; kk++) { st4a1[kk] = LWS(0, kk); st4a2[kk] = LWS(1, kk); }
        {
            double fsfcd = 0.0;
#pragma unroll
            for (int k = 1; k <= KX; k++) tt_rlw[k] = 0.0;
            // Band sweep with the level loop rolled (the body is fetched once and stays in the L0 instruction
            // cache; the unrolled 4 x 7 sweep was a third of this role's instruction stream).  For a fixed
            // level the bands are visited in order and for a fixed band the levels in order, i.e. every
            // tt_rlw(k) and flux(jb) sees the reference's sequence of operations (longwave_radiation.f90:93-105).
            {
                const int nt1 = band_row(SG(GI_T1));   // nint(T) -> row of fband(100:400,:)
                for (int jb = 1; jb <= 2; jb++) {
                    const double emis = 1.0 - STAU2(1, jb);
                    const double brad = sFband[nt1 + 301 * (jb - 1)] * (st4a1[1] + emis * st4a2[1]);
                    flux[jb] = emis * brad;
                    tt_rlw[1] = tt_rlw[1] - flux[jb];
                }
            }
            flux[3] = 0.0; flux[4] = 0.0;
            LWS(2, 1) = tt_rlw[1];
            {
                double f1 = flux[1], f2 = flux[2], f3 = flux[3], f4 = flux[4];
#pragma unroll 1
                for (int k = 2; k <= KX; k++) {
                    const double s1 = LWS(0, k), s2 = LWS(1, k);
                    const int ntk = band_row(SG(GI_T1 + k - 1));
                    double t = 0.0;
#define LW_BAND(fl, jb)                                                           \
    {                                                                             \
        const double tau = STAU2(k, jb);                                          \
        const double emis = 1.0 - tau;                                            \
        const double brad = sFband[ntk + 301 * ((jb)-1)] * (s1 + emis * s2);      \
        t = t + fl;                                                               \
        fl = tau * fl + emis * brad;                                              \
        t = t - fl;                                                               \
    }
                    LW_BAND(f1, 1) LW_BAND(f2, 2) LW_BAND(f3, 3) LW_BAND(f4, 4)
#undef LW_BAND
                    LWS(2, k) = t;
                }
                flux[1] = f1; flux[2] = f2; flux[3] = f3; flux[4] = f4;
            }
            for (int jb = 1; jb <= 4; jb++) fsfcd = fsfcd + emisfc * flux[jb];
            const double corlw = epslw * emisfc * st4a1[KX];
            LWS(2, KX) = LWS(2, KX) - corlw;
            fsfcd = fsfcd + corlw;
            slrd = fsfcd;
            DIAG2(a.L.slrd, slrd);
            SC(S_SLRD) = slrd;
            SC(S_FLX1) = flux[1]; SC(S_FLX2) = flux[2]; SC(S_FLX3) = flux[3]; SC(S_FLX4) = flux[4];
        }


        named_arrive(BAR_SEA, 96);
    };
    if (k == 3 && !csw) lw_down();
    if (k != 1) dynamics_level(k);
    STAMP(1);

    // ---------------- phase B (level warp 1): convection.f90:27-245 + the LSC reductions ----------------
    int iptop = 0, icltop = 0;
    double cloudc = 0.0, clstr = 0.0, qcloud = 0.0, precnv = 0.0, precls = 0.0;
    // the sweeps index their columns with run-time levels: they read the shared rows in place (private copies would be local-memory
    // arrays, and local memory is written through to L2: 41 MB per launch of an 8-member step before this, profiles/r2_prof_m8)
    if (k == 1) {
        double cbmf = 0.0;
        {
            const double psmin = F32(0.8), trcnv = 6.0, rhbl = F32(0.9), rhil = F32(0.7), entmax = 0.5, smf = F32(0.8);
            const int nl1 = KX - 1, nlp = KX + 1;
#pragma unroll
            for (int k = 1; k <= KX; k++) { DFSE(k) = 0.0; DFQA(k) = 0.0; }
            // diagnose_convection :170-245
            int itop = nlp;
            double qdif = 0.0;
            if (psg > psmin) {
                const double mse0 = SE(KX) + lc.alhc * QG(KX);
                double mse1 = SE(nl1) + lc.alhc * QG(nl1);
                mse1 = dmin(mse0, mse1);
                const double mss0 = dmax(mse0, SE(KX) + lc.alhc * QSAT(KX));
                int ktop1 = KX, ktop2 = KX;
                double msthr = 0.0;
                for (int k = KX - 3; k >= 3; k--) {
                    const double mssk = SE(k) + lc.alhc * QSAT(k), mssk1 = SE(k + 1) + lc.alhc * QSAT(k + 1);
                    const double mss2 = mssk + lc.wvi[(k) + 7] * (mssk1 - mssk);
                    if (mss0 > mss2) ktop1 = k;
                    if (mse1 > mss2) { ktop2 = k; msthr = mss2; }
                }
                if (ktop1 < KX) {
                    const double qthr0 = rhbl * QSAT(KX), qthr1 = rhbl * QSAT(nl1);
                    const bool lqthr = (QG(KX) > qthr0 && QG(nl1) > qthr1);
                    if (ktop2 < KX) {
                        itop = ktop1;
                        qdif = dmax(QG(KX) - qthr0, (mse0 - msthr) * lc.ralhc);
                    } else if (lqthr) {
                        itop = ktop1;
                        qdif = QG(KX) - qthr0;
                    }
                }
            }
            if (itop != nlp) {
                const double fqmax = 5.0;
                const double fm0 = lc.fm0, rdps = lc.rdps;          // p0*dhs(kx)/(grav*trcnv*3600), 2/(1-psmin): host-evaluated
                int k = KX, k1 = k - 1;
                const double qmax = dmax(F32(1.01) * QG(k), QSAT(k));
                double sb = SE(k1) + lc.wvi[(k1) + 7] * (SE(k) - SE(k1));
                double qb = QG(k1) + lc.wvi[(k1) + 7] * (QG(k) - QG(k1));
                qb = dmin(qb, QG(k));
                const double fpsa = psg * dmin(1.0, (psg - psmin) * rdps);
                double fmass = fm0 * fpsa * dmin(fqmax, qdif / (qmax - qb));
                cbmf = fmass;
                double fus = fmass * SE(k), fuq = fmass * qmax;
                double fds = fmass * sb, fdq = fmass * qb;
                DFSE(k) = fds - fus;
                DFQA(k) = fdq - fuq;
                for (k = KX - 1; k >= itop + 1; k--) {
                    k1 = k - 1;
                    DFSE(k) = fus - fds;
                    DFQA(k) = fuq - fdq;
                    const double enmass = lc.entr[k - 1] * psg * cbmf;     // convection.f90:118-131, host-evaluated
                    fmass = fmass + enmass;
                    fus = fus + enmass * SE(k);
                    fuq = fuq + enmass * QG(k);
                    sb = SE(k1) + lc.wvi[(k1) + 7] * (SE(k) - SE(k1));
                    qb = QG(k1) + lc.wvi[(k1) + 7] * (QG(k) - QG(k1));
                    fds = fmass * sb;
                    fdq = fmass * qb;
                    DFSE(k) = DFSE(k) + fds - fus;
                    DFQA(k) = DFQA(k) + fdq - fuq;
                    const double delq = rhil * QSAT(k) - QG(k);
                    if (delq > 0.0) {
                        const double fsq = smf * cbmf * delq;
                        DFQA(k) = DFQA(k) + fsq;
                        DFQA(KX) = DFQA(KX) - fsq;
                    }
                }
                k = itop;
                const double qsatb = QSAT(k) + lc.wvi[(k) + 7] * (QSAT(k + 1) - QSAT(k));
                precnv = dmax(fuq - fmass * qsatb, 0.0);
                DFSE(k) = fus - fds + lc.alhc * precnv;
                DFQA(k) = fuq - fdq - precnv;
            }
            iptop = itop;
            // physics.f90:127-138 (level 1 is not rescaled)
#pragma unroll
            for (int k = 2; k <= KX; k++) {
                DFSE(k) = DFSE(k) * rps * lc.grdscp[k - 1];
                DFQA(k) = DFQA(k) * rps * lc.grdsig[k - 1];
            }
        }
        const int icnv_ = KX - iptop;   // physics.f90:132, before LSC lowers iptop
        ib[a.L.icnv + col] = icnv_;
        SI(I_ICNV) = icnv_;
        {   // large_scale_condensation.f90:60-93: cloud top and precipitation from the level-local results
            const double prg = lc.prg;
#pragma unroll
            for (int kk = 2; kk <= KX; kk++) if (SI(I_LSC + kk - 1)) iptop = min(kk, iptop);
#pragma unroll
            for (int kk = 2; kk <= KX; kk++) {
                const double pfact = lc.dhs[kk - 1] * prg;
                precls = precls - pfact * DQLSC(kk);
            }
            precls = precls * psg;
        }
        DIAG2(a.L.precnv, precnv); DIAG2(a.L.precls, precls); DIAG2(a.L.cbmf, cbmf);
        ib[a.L.iptop + col] = iptop;
        STAMP(2);
    }
    named_sync(BAR_MID, COL_THREADS);          // + surface / forcing fields staged by the slab warp; icnv for the diffusion warp

    // ---------------- shortwave (every nstrad-th step), shortwave_radiation.f90 ----------------
    if (csw) {
        const double albcl = F32(0.43), albcls = 0.50;
        const double absdry = F32(0.033), absaer = F32(0.033), abswv1 = F32(0.022), abswv2 = 15.000, abscl1 = F32(0.015), abscl2 = F32(0.15);
        const double ablwin = F32(0.3), ablco2 = 6.0, ablwv1 = F32(0.7), ablwv2 = 50.0, ablcl1 = 12.0, ablcl2 = F32(0.6);
        const double epslw = F32(0.05);
        if (k == 1) {
            // clouds  shortwave_radiation.f90:332-410
            const double rhcl1 = F32(0.30), rhcl2 = 1.00, qacl = F32(0.20), wpcl = F32(0.2), pmaxcl = 10.0;
            const double clsmax = F32(0.60), clsminl = F32(0.15), gse_s0 = 0.25, gse_s1 = F32(0.40);
            const double gse = (SE(KX - 1) - SE(KX)) / (SG(GI_PHI + (KX - 1) - 1) - SG(GI_PHI + (KX) - 1));   // physics.f90:147
            const double rrcl = 1. / (rhcl2 - rhcl1);
            if (RH(nl1) > rhcl1) { cloudc = RH(nl1) - rhcl1; icltop = nl1; }
            else { cloudc = 0.0; icltop = nlp; }
            for (int kk = 3; kk <= KX - 2; kk++) {
                const double drh = RH(kk) - rhcl1;
                if (drh > cloudc && QG(kk) > qacl) { cloudc = drh; icltop = kk; }
            }
            {
                const double pr1 = dmin(pmaxcl, F32(86.4) * (precnv + precls));
                const double cc = dmin(1.0, cloudc * rrcl);
                cloudc = dmin(1.0, wpcl * sqrt(pr1) + cc * cc);
                icltop = min(iptop, icltop);
            }
            qcloud = QG(nl1);
            {
                const double clfact = F32(1.2), rgse = 1.0 / (gse_s1 - gse_s0);
                const double fstab = dmax(0.0, dmin(1.0, rgse * (gse - gse_s0)));
                clstr = fstab * dmax(clsmax - clfact * cloudc, 0.0);
                const double clstrl = dmax(clstr, clsminl) * RH(KX);
                const double fm = SURF(SF_FMASK);
                clstr = clstr + fm * (clstrl - clstr);
            }
            ib[a.L.icltop + col] = icltop;
            DIAG2(a.L.qcloud, qcloud); DIAG2(a.L.cloudc, cloudc); DIAG2(a.L.clstr, clstr);
            SC(S_CLOUDC) = cloudc; SC(S_QCLOUD) = qcloud; SI(I_ICLTOP) = icltop;
        }
        named_sync(BAR_LEV, LEV_THREADS);
        // ---- phase C (wide): every transmissivity of level k (shortwave_radiation.f90:130-150, 190-233)
        {
            const double cloudc_ = SC(S_CLOUDC), qcloud_ = SC(S_QCLOUD);
            const int icltop_ = SI(I_ICLTOP);
            const double zenit = SURF(SF_ZENIT);
            const double psaz = psg * zenit;
            double acloud = cloudc_ * dmin(abscl1 * qcloud_, abscl2);
            double t1;
            if (k == 1) {
                t1 = exp(-psaz * lc.dhs[0] * absdry);
            } else if (k <= nl1) {
                const double abs1 = absdry + absaer * (lc.fsg[k - 1] * lc.fsg[k - 1]);
                if (k >= icltop_) t1 = exp(-psaz * lc.dhs[k - 1] * (abs1 + abswv1 * qgk + acloud));
                else t1 = exp(-psaz * lc.dhs[k - 1] * (abs1 + abswv1 * qgk));
            } else {
                const double abs1 = absdry + absaer * (lc.fsg[KX - 1] * lc.fsg[KX - 1]);
                t1 = exp(-psaz * lc.dhs[KX - 1] * (abs1 + abswv1 * qgk));
            }
            TAU1(k) = t1;
            TAU2S(k) = (k == 1) ? 0.0 : exp(-psaz * lc.dhs[k - 1] * abswv2 * qgk);
            // longwave transmissivities :190-233 -> persistent tau2(ix,il,kx,4)
            if (k == 1) {
                TAU2W(1, 1, exp(-psg * lc.dhs[0] * ablwin));
                TAU2W(1, 2, exp(-psg * lc.dhs[0] * ablco2));
                TAU2W(1, 3, 1.0);
                TAU2W(1, 4, 1.0);
            } else if (k == 2 || k == KX) {
                TAU2W(k, 1, exp(-psg * lc.dhs[k - 1] * ablwin));
                TAU2W(k, 2, exp(-psg * lc.dhs[k - 1] * ablco2));
                TAU2W(k, 3, exp(-psg * lc.dhs[k - 1] * ablwv1 * qgk));
                TAU2W(k, 4, exp(-psg * lc.dhs[k - 1] * ablwv2 * qgk));
            } else {
                acloud = cloudc_ * ablcl2;
                const double deltap = psg * lc.dhs[k - 1];
                double acloud1;
                if (k < icltop_) acloud1 = acloud;
                else acloud1 = ablcl1 * cloudc_;
                TAU2W(k, 1, exp(-deltap * (ablwin + acloud1)));
                TAU2W(k, 2, exp(-deltap * ablco2));
                TAU2W(k, 3, exp(-deltap * dmax(ablwv1 * qgk, acloud)));
                TAU2W(k, 4, exp(-deltap * dmax(ablwv2 * qgk, acloud)));
            }
        }
        named_sync(BAR_LEV, LEV_THREADS);
        if (k == 3) lw_down();
        if (k == 1) {
            // get_shortwave_rad_fluxes  shortwave_radiation.f90:74-234: the flux sweeps
            const double fsol = SURF(SF_FSOL), ozone = SURF(SF_OZONE), ozupp = SURF(SF_OZUPP), stratz = SURF(SF_STRATZ);
            const double albsfc = SURF(SF_ALBSFC);
            const double fband2 = F32(0.05), fband1 = 1.0 - fband2;
            double tau1[KX + 1], tau2_[KX + 1], tau3[KX + 1], dfabs[KX + 1];
#pragma unroll
            for (int kk = 1; kk <= KX; kk++) { tau1[kk] = TAU1(kk); tau2_[kk] = TAU2S(kk); tau3[kk] = (kk == icltop) ? albcl * cloudc : 0.0; }   // static indices: registers
            tau3[KX] = albcls * clstr;
            double ftop = fsol;
            double flux1 = fsol * fband1, flux2 = fsol * fband2;
            dfabs[1] = flux1;
            flux1 = tau1[1] * (flux1 - ozupp * psg);
            dfabs[1] = dfabs[1] - flux1;
            dfabs[2] = flux1;
            flux1 = tau1[2] * (flux1 - ozone * psg);
            dfabs[2] = dfabs[2] - flux1;
#pragma unroll
            for (int kk = 3; kk <= KX; kk++) {
                tau3[kk] = flux1 * tau3[kk];
                flux1 = flux1 - tau3[kk];
                dfabs[kk] = flux1;
                flux1 = tau1[kk] * flux1;
                dfabs[kk] = dfabs[kk] - flux1;
            }
#pragma unroll
            for (int kk = 2; kk <= KX; kk++) {
                dfabs[kk] = dfabs[kk] + flux2;
                flux2 = tau2_[kk] * flux2;
                dfabs[kk] = dfabs[kk] - flux2;
            }
            const double fsfcd = flux1 + flux2;
            flux1 = flux1 * albsfc;
            const double fsfc = fsfcd - flux1;
#pragma unroll
            for (int kk = KX; kk >= 1; kk--) {
                dfabs[kk] = dfabs[kk] + flux1;
                flux1 = tau1[kk] * flux1;
                dfabs[kk] = dfabs[kk] - flux1;
                flux1 = flux1 + tau3[kk];
            }
            ftop = ftop - flux1;
            G2(a.L.ssrd) = fsfcd; SURF(SF_SSRD) = fsfcd; DIAG2(a.L.ssr, fsfc); DIAG2(a.L.tsr, ftop);
#pragma unroll
            for (int kk = 1; kk <= KX; kk++) { const double v = dfabs[kk] * rps * lc.grdscp[kk - 1]; G3(a.L.tt_rsw, kk) = v; RSW(kk) = v; }   // physics.f90:160-162
            const double eps1 = lc.eps1;
            mb[a.L.stratc + col] = STRATC(0) = stratz * psg;
            mb[a.L.stratc + N + col] = STRATC(1) = eps1 * psg;
        }
    }
    STAMP(3);

    // ---------------- surface_fluxes.f90:42-296: level warp 2 prepares the shared terms and the sea half ----------------
    const double fwind0 = F32(0.95), ftemp0 = 1.0, cdl = F32(2.4e-3), cds = F32(1.0e-3), chl = F32(1.2e-3), chs = F32(0.9e-3);
    const double vgust = 5.0, ctday = F32(1.0e-2), dtheta = 3.0, fstab = F32(0.67), clambda = 7.0, clambsn = 7.0;
    const double esbc = emisfc * lc.sbc;
    const double rdth = fstab / dtheta, astab = 0.5;
    if (k == 2) {
        const double ug8 = SG(GI_U1 + KX - 1), vg8 = SG(GI_V1 + KX - 1);   // only the lowest-level wind is used
        const double tg8 = SG(GI_T1 + KX - 1), tg7 = SG(GI_T1 + nl1 - 1), qg8 = QG(KX), phig8 = SG(GI_PHI + KX - 1);
        const double phi0 = SURF(SF_PHIS0), fmask = SURF(SF_FMASK), tsea = SURF(SF_SST);
        const double u0 = fwind0 * ug8, v0 = fwind0 * vg8;
        const double gtemp0 = 1.0 - ftemp0, rcp = lc.rcp;
        const double dt1 = lc.wvi[7 + KX] * (tg8 - tg7);
        double t1_1 = tg8 + dt1;
        double t1_2 = t1_1 - phi0 * dt1 / (lc.rgas * 288.0 * lc.sigl[KX - 1]);
        const double t2_2 = tg8 + rcp * phig8;
        const double t2_1 = t2_2 - rcp * phi0;
        if (tg8 > tg7) {
            t1_1 = ftemp0 * t1_1 + gtemp0 * t2_1;
            t1_2 = ftemp0 * t1_2 + gtemp0 * t2_2;
        } else {
            t1_1 = tg8;
            t1_2 = tg8;
        }
        const double t0 = t1_2 + fmask * (t1_1 - t1_2);
        const double denvvs0 = (lc.p0 * psg / (lc.rgas * t0)) * sqrt(u0 * u0 + v0 * v0 + vgust * vgust);
        double dths;
        if (tsea > t2_2) dths = dmin(dtheta, tsea - t2_2);
        else dths = dmax(-dtheta, astab * (tsea - t2_2));
        const double denvvs2 = denvvs0 * (1.0 + dths * rdth);
        const double q1_2 = qg8;
        const double cdsdv = cds * denvvs2;
        const double ustr2 = -cdsdv * ug8, vstr2 = -cdsdv * vg8;
        const double shf2 = chs * lc.cp * denvvs2 * (tsea - t1_2);
        const double qsat0_s = qsat_pt(tsea, psg);
        const double evap2 = chs * denvvs2 * (qsat0_s - q1_2);
        const double slru2 = esbc * ((tsea * tsea) * (tsea * tsea));
        SC(S_T1_1) = t1_1; SC(S_T1_2) = t1_2; SC(S_T2_1) = t2_1; SC(S_DENVVS0) = denvvs0; SC(S_T0) = t0; SC(S_U0) = u0; SC(S_V0) = v0;
        SC(S_USTR2) = ustr2; SC(S_VSTR2) = vstr2; SC(S_SHF2) = shf2; SC(S_EVAP2) = evap2; SC(S_SLRU2) = slru2;
        named_arrive(BAR_SEA, 96);
    }
    if (k == 1) {
        // ------------------- downward longwave: done by level warp 3 (lw_down), results in shared memory -------------------
        double st4a1[KX + 1], st4a2[KX + 1], tt_rlw[KX + 1], flux[5];
        STAMP(4);
        // ------------------------- surface_fluxes.f90:42-296 (lfluxland = .true.): land half and the blend -------------------------
        named_sync(BAR_SEA, 96);                // sea half of the fluxes (warp 2), long-wave down (warp 3)
        const double slrd = SC(S_SLRD);
#pragma unroll
        for (int kk = 1; kk <= KX; kk++) { st4a1[kk] = LWS(0, kk); st4a2[kk] = LWS(1, kk); }
        flux[1] = SC(S_FLX1); flux[2] = SC(S_FLX2); flux[3] = SC(S_FLX3); flux[4] = SC(S_FLX4);
        double ts, shf3, evap3, ustr3, vstr3, slru3;
        {
            const double ug8 = SG(GI_U1 + KX - 1), vg8 = SG(GI_V1 + KX - 1);
            const double fmask = SURF(SF_FMASK), tsea = SURF(SF_SST), stl_am = SURF(SF_STL);
            const double soilw_am = SURF(SF_SOILW), alb_l = SURF(SF_ALBL), alb_s = SURF(SF_ALBS), snowc = SURF(SF_SNOWC), forog = SURF(SF_FOROG);
            const double ssrd = SURF(SF_SSRD);
            const double t1_1 = SC(S_T1_1), t1_2 = SC(S_T1_2), t2_1 = SC(S_T2_1), denvvs0 = SC(S_DENVVS0);
            const double u0 = SC(S_U0), v0 = SC(S_V0);
            double tskin = stl_am + ctday * sqrt(coa_j) * ssrd * (1.0 - alb_l) * psg;
            double dthl;
            if (tskin > t2_1) dthl = dmin(dtheta, tskin - t2_1);
            else dthl = dmax(-dtheta, astab * (tskin - t2_1));
            const double denvvs1 = denvvs0 * (1.0 + dthl * rdth);
            const double cdldv = cdl * denvvs0 * forog;
            const double ustr1 = -cdldv * ug8, vstr1 = -cdldv * vg8;
            const double chlcp = chl * lc.cp;
            double shf1 = chlcp * denvvs1 * (tskin - t1_1);
            const double q1_1 = QG(KX);
            const double qsat0_1 = qsat_pt(tskin, psg);
            double evap1 = chl * denvvs1 * dmax(0.0, soilw_am * qsat0_1 - q1_1);
            const double tsk3 = (tskin * tskin) * tskin;
            const double dslr = 4.0 * esbc * tsk3;
            double slru1 = esbc * tsk3 * tskin;
            double hfluxn1 = ssrd * (1.0 - alb_l) + slrd - (slru1 + shf1 + lc.alhc * evap1);
            {   // lskineb
                const double clamb = clambda + snowc * (clambsn - clambda);
                hfluxn1 = hfluxn1 - clamb * (tskin - stl_am);
                double dtskin = tskin + 1.0;
                double qsat0_2 = qsat_pt(dtskin, psg);
                if (evap1 > 0.0) qsat0_2 = soilw_am * (qsat0_2 - qsat0_1);
                else qsat0_2 = 0.0;
                dtskin = hfluxn1 / (clamb + dslr + chl * denvvs1 * (lc.cp + lc.alhc * qsat0_2));
                tskin = tskin + dtskin;
                shf1 = shf1 + chlcp * denvvs1 * dtskin;
                evap1 = evap1 + chl * denvvs1 * qsat0_2 * dtskin;
                slru1 = slru1 + dslr * dtskin;
                hfluxn1 = clamb * (tskin - stl_am);
            }
            const double ustr2 = SC(S_USTR2), vstr2 = SC(S_VSTR2), shf2 = SC(S_SHF2), evap2 = SC(S_EVAP2), slru2 = SC(S_SLRU2);
            const double hfluxn2 = ssrd * (1.0 - alb_s) + slrd - slru2 + shf2 + lc.alhc * evap2;
            ustr3 = ustr2 + fmask * (ustr1 - ustr2);
            vstr3 = vstr2 + fmask * (vstr1 - vstr2);
            shf3 = shf2 + fmask * (shf1 - shf2);
            evap3 = evap2 + fmask * (evap1 - evap2);
            slru3 = slru2 + fmask * (slru1 - slru2);
            ts = tsea + fmask * (stl_am - tsea);
            tskin = tsea + fmask * (tskin - tsea);
            const double t0 = t1_2 + fmask * (t1_1 - t1_2);
            DIAG2(a.L.ustr, ustr1); DIAG2(a.L.ustr + N, ustr2); DIAG2(a.L.ustr + 2 * N, ustr3);
            DIAG2(a.L.vstr, vstr1); DIAG2(a.L.vstr + N, vstr2); DIAG2(a.L.vstr + 2 * N, vstr3);
            DIAG2(a.L.shf, shf1); mb[a.L.shf + N + col] = shf2; DIAG2(a.L.shf + 2 * N, shf3);        // (:,2) of shf / evap: read by the sea model of the next step
            DIAG2(a.L.evap, evap1); mb[a.L.evap + N + col] = evap2; DIAG2(a.L.evap + 2 * N, evap3);
            DIAG2(a.L.slru, slru1); DIAG2(a.L.slru + N, slru2); DIAG2(a.L.slru + 2 * N, slru3);
            mb[a.L.hfluxn + col] = hfluxn1; mb[a.L.hfluxn + N + col] = hfluxn2;
            DIAG2(a.L.ts, ts); DIAG2(a.L.tskin, tskin); DIAG2(a.L.u0, u0); DIAG2(a.L.v0, v0); DIAG2(a.L.t0, t0);
        }
        STAMP(5);
        // ------------------- upward longwave  longwave_radiation.f90:120-194 -------------------
        {
            const double refsfc = 1.0 - emisfc;
            const double fsfcu = slru3;
            DIAG2(a.L.slr, fsfcu - slrd);
            const int nts = band_row(ts);
            for (int jb = 1; jb <= 4; jb++) flux[jb] = sFband[nts + 301 * (jb - 1)] * fsfcu + refsfc * flux[jb];
            LWS(2, KX) = LWS(2, KX) + epslw * fsfcu;
            {
                double f1 = flux[1], f2 = flux[2], f3 = flux[3], f4 = flux[4];
#pragma unroll 1
                for (int k = KX; k >= 2; k--) {     // longwave_radiation.f90:155-167, level loop rolled as in the downward sweep
                    const double s1 = LWS(0, k), s2 = LWS(1, k);
                    const int ntk = band_row(SG(GI_T1 + k - 1));
                    double t = LWS(2, k);
#define LW_BAND(fl, jb)                                                           \
    {                                                                             \
        const double tau = STAU2(k, jb);                                          \
        const double emis = 1.0 - tau;                                            \
        const double brad = sFband[ntk + 301 * ((jb)-1)] * (s1 - emis * s2);      \
        t = t + fl;                                                               \
        fl = tau * fl + emis * brad;                                              \
        t = t - fl;                                                               \
    }
                    LW_BAND(f1, 1) LW_BAND(f2, 2) LW_BAND(f3, 3) LW_BAND(f4, 4)
#undef LW_BAND
                    LWS(2, k) = t;
                }
                flux[1] = f1; flux[2] = f2; flux[3] = f3; flux[4] = f4;
            }
            {
                const int nt1 = band_row(SG(GI_T1 + (1) - 1));
                double t = LWS(2, 1);
                for (int jb = 1; jb <= 2; jb++) {
                    const double tau = STAU2(1, jb);
                    const double emis = 1.0 - tau;
                    const double brad = sFband[nt1 + 301 * (jb - 1)] * (st4a1[1] - emis * st4a2[1]);
                    t = t + flux[jb];
                    flux[jb] = tau * flux[jb] + emis * brad;
                    t = t - flux[jb];
                }
                LWS(2, 1) = t;
            }
#pragma unroll
            for (int k = 1; k <= KX; k++) tt_rlw[k] = LWS(2, k);
            const double stratc1 = STRATC(0), stratc2 = STRATC(1);
            const double corlw1 = lc.dhs[0] * stratc2 * st4a1[1] + stratc1;
            const double corlw2 = lc.dhs[1] * stratc2 * st4a1[2];
            tt_rlw[1] = tt_rlw[1] - corlw1;
            tt_rlw[2] = tt_rlw[2] - corlw2;
            double ftop = corlw1 + corlw2;
            for (int jb = 1; jb <= 4; jb++) ftop = ftop + flux[jb];
            DIAG2(a.L.olr, ftop);
#pragma unroll
            for (int k = 1; k <= KX; k++) tt_rlw[k] = tt_rlw[k] * rps * lc.grdscp[k - 1];   // physics.f90:182-186, added in the closing stage
        }


        // hand the column-serial results to the level warps (physics.f90:197-205)
        SC(S_UT8) = 0.0 + ustr3 * rps * lc.grdsig[KX - 1];
        SC(S_VT8) = 0.0 + vstr3 * rps * lc.grdsig[KX - 1];
        SC(S_SHFT) = shf3 * rps * lc.grdscp[KX - 1];
        SC(S_EVAPT) = evap3 * rps * lc.grdsig[KX - 1];
#pragma unroll
        for (int kk = 1; kk <= KX; kk++) LWS(2, kk) = tt_rlw[kk];
        STAMP(6);
    }
    // ---------------- closing stage (wide): physics.f90:137-138, 182-186, 197-205, 208-222 for level k ----------------
    named_sync(BAR_END, COL_THREADS);     // serial results of level warp 1 and the vertical-diffusion fluxes are in shared memory
    {
        const double ut_dyn = DYN(0, k), vt_dyn = DYN(1, k), tt_dyn = DYN(2, k), qt_dyn = DYN(3, k);
        double ut = ut_dyn, vt = vt_dyn, tt = tt_dyn, qt = qt_dyn;
        tt = tt + DFSE(k) + DTLSC(k);
        qt = qt + DFQA(k) + DQLSC(k);
        tt = tt + RSW(k) + LWS(2, k);
        double ttenvd = VD(0, k), qtenvd = VD(1, k);
        if (k == KX) {
            ttenvd = ttenvd + SC(S_SHFT);
            qtenvd = qtenvd + SC(S_EVAPT);
            ut = ut + SC(S_UT8); vt = vt + SC(S_VT8);
        } else {
            ut = ut + 0.0; vt = vt + 0.0;
        }
        tt = tt + ttenvd; qt = qt + qtenvd;
        if (a.sppt_on) {
            double p = SG(GI_SPPT + k - 1);
            p = dmin(1.0, fabs(p)) * copysign(1.0, p);   // sppt.f90:98
            const double f = (1 + p * 1.0);
            ut = f * (ut - ut_dyn) + ut_dyn;
            vt = f * (vt - vt_dyn) + vt_dyn;
            tt = f * (tt - tt_dyn) + tt_dyn;
            qt = f * (qt - qt_dyn) + qt_dyn;
        }
        const int f0 = GO_PER * (k - 1);
        GOUT(f0 + 0) = ut; GOUT(f0 + 1) = vt; GOUT(f0 + 5) = tt; GOUT(f0 + 8) = qt;
    }
    STAMP(7);
    if (lane == 0) trace_end(a.trace, 1);
#undef SROW
#undef SG
#undef SURF
#undef STAU2
#undef STRATC
#undef RSW
#undef DYN
#undef SE
#undef QSAT
#undef RH
#undef QG
#undef DFSE
#undef DFQA
#undef DTLSC
#undef DQLSC
#undef VD
#undef LWS
#undef TAU1
#undef TAU2S
#undef SC
#undef SI
#undef TAU2W
#undef STAMP
}

// ------------------------------------------------------------------------------------------
// Per-step surface slabs: couple_land_atm + couple_sea_atm (land_model.f90:184-239,
// sea_model.f90:253-444, interpolation.f90:16-69) for one grid point per thread.
// ------------------------------------------------------------------------------------------
struct SlabArgs {
    double* base; long long stride;
    Layout L;
    SharedDev sh;
    DevClock* clk;
    const LevelConsts* lc;
    int N;
    int day0;    // 1: the `day == 0` initialisation call of initialize_coupler
};

__device__ __forceinline__ double forint_pt(const double* f12, int N, int q, int imon, double tmonth) {   // interpolation.f90:16-35
    int imon2;
    double wmon;
    if (tmonth <= 0.5) { imon2 = imon - 1; if (imon == 1) imon2 = 12; wmon = 0.5 - tmonth; }
    else { imon2 = imon + 1; if (imon == 12) imon2 = 1; wmon = tmonth - 0.5; }
    const double a0 = f12[(size_t)(imon - 1) * N + q], b0 = f12[(size_t)(imon2 - 1) * N + q];
    return a0 + wmon * (b0 - a0);
}
__device__ __forceinline__ double forin5_pt(const double* f12, int N, int q, int imon, double tmonth) {   // interpolation.f90:38-69
    int im2 = imon - 2, im1 = imon - 1, ip1 = imon + 1, ip2 = imon + 2;
    if (im2 < 1) im2 += 12;
    if (im1 < 1) im1 += 12;
    if (ip1 > 12) ip1 -= 12;
    if (ip2 > 12) ip2 -= 12;
    const double c0 = (double)(1.0f / 12.0f);
    const double t0 = c0 * tmonth, t1 = c0 * (1.0 - tmonth), t2 = 0.25 * tmonth * (1 - tmonth);
    const double wm2 = -t1 + t2, wm1 = -c0 + 8 * t1 - 6 * t2, w0 = 7 * c0 + 10 * t2, wp1 = -c0 + 8 * t0 - 6 * t2, wp2 = -t0 + t2;
    return wm2 * f12[(size_t)(im2 - 1) * N + q] + wm1 * f12[(size_t)(im1 - 1) * N + q] + w0 * f12[(size_t)(imon - 1) * N + q] +
           wp1 * f12[(size_t)(ip1 - 1) * N + q] + wp2 * f12[(size_t)(ip2 - 1) * N + q];
}

// one grid point of couple_land_atm + couple_sea_atm
__device__ __forceinline__ void slab_point(double* mb, const Layout& L_, const SharedDev& sh_, const DevClock& c, const LevelConsts& lc, int N, int q, int day0) {
    struct { const Layout& L; const SharedDev& sh; int day0; } a{L_, sh_, day0};
    const int imont1 = c.imont1;
    const double tmonth = c.tmonth;
#define S2(off) mb[(off) + q]
    // ---- land (land_model.f90:184-239)
    const double stlcl = forin5_pt(a.sh.stl12, N, q, imont1, tmonth);
    const double snowdcl = forint_pt(a.sh.snowd12, N, q, imont1, tmonth);
    const double soilwcl = forint_pt(a.sh.soilw12, N, q, imont1, tmonth);
    if (a.day0) {
        S2(a.L.stl_lm) = stlcl;
        S2(a.L.stl_am) = stlcl;
    } else {
        double tanom = S2(a.L.stl_lm) - stlcl;
        tanom = a.sh.cdland[q] * (tanom + a.sh.rhcapl[q] * mb[a.L.hfluxn + q]);
        const double v = tanom + stlcl;
        S2(a.L.stl_lm) = v;
        S2(a.L.stl_am) = v;
    }
    S2(a.L.stlcl_ob) = stlcl;
    S2(a.L.snowd_am) = snowdcl;
    S2(a.L.soilw_am) = soilwcl;
    // ---- sea (sea_model.f90:253-363)
    double sstcl = forin5_pt(a.sh.sst12, N, q, imont1, tmonth);
    double sicecl = forint_pt(a.sh.sice12, N, q, imont1, tmonth);
    double* an = mb + a.L.sstan3;
    if (!a.day0 && c.obs_ssta) {   // obs_ssta :366-385
        an[q] = an[N + q];
        an[N + q] = an[2 * N + q];
        const int rec = c.next_month;
        an[2 * N + q] = (rec >= 1 && rec <= c.nssta) ? (double)a.sh.ssta[(size_t)(rec - 1) * N + q] : 0.0;
    }
    double sstan_ob;
    {   // forint(2, sstan3, sstan_ob)
        int imon2;
        double wmon;
        if (tmonth <= 0.5) { imon2 = 1; wmon = 0.5 - tmonth; }
        else { imon2 = 3; wmon = tmonth - 0.5; }
        const double a0 = an[N + q], b0 = an[(size_t)(imon2 - 1) * N + q];
        sstan_ob = a0 + wmon * (b0 - a0);
    }
    const double sstfr = (double)(273.2f - 1.8f);   // :285 real32 subtraction
    double ticecl;
    if (sstcl > sstfr) {
        sicecl = dmin(0.5, sicecl);
        ticecl = sstfr;
        if (sicecl > 0.0) sstcl = sstfr + (sstcl - sstfr) / (1.0 - sicecl);
    } else {
        sicecl = dmax(0.5, sicecl);
        ticecl = sstfr + (sstcl - sstfr) / sicecl;
        sstcl = sstfr;
    }
    double sst_om, tice_om, sice_om;
    if (a.day0) {
        sst_om = 0.0;   // sea_coupling_flag <= 0
        tice_om = ticecl;
        sice_om = sicecl;
    } else {   // run_sea_model :387-444 (uses the pre-update sstcl_ob/sicecl_ob just computed, tice_am/sice_am of the last coupling)
        const double albsea = F32(0.07), albice = F32(0.60), emisfc = F32(0.98), beta = 1.0;
        const double tice_am = S2(a.L.tice_am), sice_am = S2(a.L.sice_am);
        const double ssrd = S2(a.L.ssrd), shf2 = mb[a.L.shf + N + q], evap2 = mb[a.L.evap + N + q], hfluxn2 = mb[a.L.hfluxn + N + q];
        sst_om = S2(a.L.sst_om); tice_om = S2(a.L.tice_om);
        const double sstfr4 = (sstfr * sstfr) * (sstfr * sstfr);
        const double difice = (albsea - albice) * ssrd + emisfc * lc.sbc * (sstfr4 - (tice_am * tice_am) * (tice_am * tice_am)) + shf2 + evap2 * lc.alhc;
        const double hflux_i = hfluxn2 + difice * (1.0 - sice_am);
        double hflux = hfluxn2 - 0.0 - sicecl * (hflux_i + beta * (sstfr - tice_om));
        double tanom = sst_om - sstcl;
        tanom = a.sh.cdsea[q] * (tanom + a.sh.rhcaps[q] * hflux);
        sst_om = tanom + sstcl;
        hflux = hflux_i + beta * (sstfr - tice_om);
        tanom = tice_om - ticecl;
        const double anom0 = 20.;
        const double cdis = a.sh.cdice[q] * (anom0 / (anom0 + fabs(tanom)));
        tanom = cdis * (tanom + a.sh.rhcapi[q] * hflux);
        tice_om = tanom + ticecl;
        sice_om = sicecl;
    }
    S2(a.L.sst_om) = sst_om; S2(a.L.tice_om) = tice_om; S2(a.L.sice_om) = sice_om;
    S2(a.L.sstcl_ob) = sstcl; S2(a.L.sicecl_ob) = sicecl; S2(a.L.ticecl_ob) = ticecl;
    double sst_am = sstcl + sstan_ob;
    const double sice_am = sice_om, tice_am = tice_om;
    sst_am = sst_am + sice_am * (tice_am - sst_am);
    S2(a.L.sst_am) = sst_am; S2(a.L.sice_am) = sice_am; S2(a.L.tice_am) = tice_am;
    S2(a.L.ssti_om) = sst_om + sice_am * (tice_am - sst_om);
#undef S2
}


__global__ void k_slab(SlabArgs a) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.N) return;
    if (q == 0 && blockIdx.y == 0 && !a.day0) a.clk->slab_pending = 0;   // the coupler call of the last step is being applied (no thread here reads the flag)
    slab_point(a.base + (size_t)blockIdx.y * a.stride, a.L, a.sh, *a.clk, *a.lc, a.N, q, a.day0);
}

// ------------------------------------------------------------------------------------------
// Daily forcing (forcing.f90:44-99): solar fields from the per-day table, albedos, snow
// cover and the grid-point humidity correction (its transform follows in K2).
// Runs every step inside the graph but returns at once unless the clock says it is due.
// ------------------------------------------------------------------------------------------
struct ForcingArgs {
    double* base; long long stride;
    Layout L;
    SharedDev sh;
    const DevClock* clk;
    const LevelConsts* lc;
    int ix, il;
    int force;
};

// one grid point of set_forcing's daily part
__device__ __forceinline__ void forcing_point(double* mb, const Layout& L_, const SharedDev& sh_, const DevClock& clk, const LevelConsts& lc, int ix, int il, int q) {
    struct { const Layout& L; const SharedDev& sh; int il; } a{L_, sh_, il};
    const int j = q / ix;
    const double* S = a.sh.solar + (size_t)clk.doy * 5 * a.il;
    mb[a.L.fsol + q] = S[0 * a.il + j];
    mb[a.L.ozone + q] = S[1 * a.il + j];
    mb[a.L.ozupp + q] = S[2 * a.il + j];
    mb[a.L.zenit + q] = S[3 * a.il + j];
    mb[a.L.stratz + q] = S[4 * a.il + j];
    const double albsea = F32(0.07), albice = F32(0.60), albsn = F32(0.60), sd2sc = 60.0;
    const double alb0 = mb[a.L.alb0 + q], fmask_l = mb[a.L.fmask_l + q], fmask_s = mb[a.L.fmask_s + q];
    const double snowc = dmin(1.0, mb[a.L.snowd_am + q] / sd2sc);
    const double alb_l = alb0 + snowc * (albsn - alb0);
    const double alb_s = albsea + mb[a.L.sice_am + q] * (albice - albsea);
    mb[a.L.snowc + q] = snowc;
    mb[a.L.alb_l + q] = alb_l;
    mb[a.L.alb_s + q] = alb_s;
    mb[a.L.albsfc + q] = alb_s + fmask_l * (alb_l - alb_s);
    // forcing.f90:77-99
    const double gamlat = lc.gamma / (1000. * lc.grav);
    const double corh = gamlat * mb[a.L.phis0 + q];
    const double pexp = 1. / (lc.rgas * gamlat);
    const double tsfc = fmask_l * mb[a.L.stl_am + q] + fmask_s * mb[a.L.sst_am + q];
    const double tref = tsfc + corh;
    const double psfc = pow(tsfc / tref, pexp);
    const double qref = qsat_pt(tref, psfc / psfc);   // get_qsat(tref, psfc/psfc, -1): ps(1,1) = 1 wherever psfc is finite
    const double qsfc = qsat_pt(tsfc, psfc);
    mb[a.L.qcorh_g + q] = lc.refrh1 * (qref - qsfc);
}


__global__ void k_daily_forcing(ForcingArgs a) {
    if (!a.force && !a.clk->do_forcing) return;
    const int N = a.ix * a.il;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= N) return;
    forcing_point(a.base + (size_t)blockIdx.y * a.stride, a.L, a.sh, *a.clk, *a.lc, a.ix, a.il, q);
}

// device calendar: the end-of-step bookkeeping of speedy.f90:44-47
__global__ void k_clock_advance(DevClock* c) {
    if (threadIdx.x == 0 && blockIdx.x == 0) cal_advance(*c);
}

// ---- tensor maps of the column kernel ---------------------------------------------------------
struct ColMaps { CUtensorMap dyn, phys, tau2, stratc, rsw; int sppt_on; };
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMap make_map3(EncodeTiledFn enc, double* base, long long N, int rows, int nmembers, long long member_stride, int box_rows) {
    CUtensorMap m;
    const cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)rows, (cuuint64_t)nmembers};
    const cuuint64_t strides[2] = {(cuuint64_t)N * sizeof(double), (cuuint64_t)member_stride * sizeof(double)};
    const cuuint32_t box[3] = {(cuuint32_t)TC, (cuuint32_t)box_rows, 1u};
    const cuuint32_t es[3] = {1u, 1u, 1u};
    const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return m;
}
static const ColMaps& column_maps(speedy_ctx* ctx) {
    Model& M = *ctx->model;
    if (M.colmaps && static_cast<ColMaps*>(M.colmaps)->sppt_on == ctx->sppt_on) return *static_cast<ColMaps*>(M.colmaps);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) throw std::runtime_error("cuTensorMapEncodeTiled is not available in this driver");
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
    const long long N = ctx->d.ngrid();
    const int ngin = ctx->sppt_on ? GI_N : GI_NBASE;
    ColMaps* cm = M.colmaps ? static_cast<ColMaps*>(M.colmaps) : new ColMaps();
    cm->sppt_on = ctx->sppt_on;
    cm->dyn = make_map3(enc, M.mem.p + M.L.gin, N, GI_N, ctx->nmembers, M.L.stride, GI_U1);
    cm->phys = make_map3(enc, M.mem.p + M.L.gin, N, GI_N, ctx->nmembers, M.L.stride, ngin - GI_U1);
    cm->tau2 = make_map3(enc, M.mem.p + M.L.tau2, N, 4 * KX, ctx->nmembers, M.L.stride, 4 * KX);
    cm->stratc = make_map3(enc, M.mem.p + M.L.stratc, N, 2, ctx->nmembers, M.L.stride, 2);
    cm->rsw = make_map3(enc, M.mem.p + M.L.tt_rsw, N, KX, ctx->nmembers, M.L.stride, KX);
    M.colmaps = cm;
    return *cm;
}

void free_column_maps(Model& M) { delete static_cast<ColMaps*>(M.colmaps); M.colmaps = nullptr; }

// per device (speedy_create calls it after cudaSetDevice): the shared-memory opt-in is a per-device function attribute
void setup_column_kernels() {
    CUDA_CHECK(cudaFuncSetAttribute(k_grid_columns<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COL_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(k_grid_columns<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COL_SMEM));
}

// ---- launchers ------------------------------------------------------------------------------
void launch_grid_columns(speedy_ctx* ctx, int mode, int csw_override, int merged) {
    Model& M = *ctx->model;
    ColumnArgs a;
    a.sh = M.sh; a.merged = merged;
    a.base = M.mem.p; a.stride = M.L.stride; a.ibase = M.imem.p; a.L = M.L; a.lc = M.lc.p; a.clk = M.clock.p;
    a.fband = ctx->dv.fband; a.coriol = ctx->dv.coriol; a.coa = ctx->dv.coa;
    a.ready = M.ready_target ? M.ready.p : nullptr; a.ready_target = M.ready_target;
    a.discard_row0 = 0;
    if (merged && mode == 0 && M.alias_active) {      // the K2 inputs go over the grid fields this tile has staged (rows 0..GO_N-1 of the same columns)
        a.L.gout = M.L.gin;
        a.L.qcorh_g = M.L.gin + (long long)GO_QCORH * ctx->d.ngrid();
        a.discard_row0 = GO_N;
    }
    a.discard_gin = merged && ctx->l2_discard && mode == 0 && (long long)(ctx->d.ngrid() / TC) * ctx->nmembers > ctx->num_sms;
    a.ix = ctx->d.ix; a.il = ctx->d.il; a.mode = mode; a.csw_override = csw_override; a.sppt_on = ctx->sppt_on; a.trace = ctx->dv.trace;
    {
        const ColMaps& cm = column_maps(ctx);
        a.m_dyn = cm.dyn; a.m_phys = cm.phys; a.m_tau2 = cm.tau2; a.m_stratc = cm.stratc; a.m_rsw = cm.rsw;
    }
    const int N = ctx->d.ngrid();
    if (N % TC) throw std::runtime_error("grid size must be a multiple of the column tile");
    dim3 grid(N / TC, ctx->nmembers);
    // more tiles than SMs (ensemble batches, T47): the two-CTAs-per-SM variant; otherwise the uncapped one
    if ((long long)grid.x * grid.y > ctx->num_sms) CUDA_CHECK(launch_pdl(ctx->dv.trace == nullptr || ctx->trace_pdl, k_grid_columns<true>, grid, dim3(COL_THREADS), COL_SMEM, ctx->stream, a));
    else CUDA_CHECK(launch_pdl(ctx->dv.trace == nullptr || ctx->trace_pdl, k_grid_columns<false>, grid, dim3(COL_THREADS), COL_SMEM, ctx->stream, a));
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
}

void launch_slab(speedy_ctx* ctx, int day0) {
    Model& M = *ctx->model;
    SlabArgs a;
    a.base = M.mem.p; a.stride = M.L.stride; a.L = M.L; a.sh = M.sh; a.clk = M.clock.p; a.lc = M.lc.p; a.N = ctx->d.ngrid(); a.day0 = day0;
    dim3 grid((a.N + 127) / 128, ctx->nmembers);
    k_slab<<<grid, 128, 0, ctx->stream>>>(a);
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
}

void launch_daily_forcing(speedy_ctx* ctx, int force) {
    Model& M = *ctx->model;
    ForcingArgs a;
    a.base = M.mem.p; a.stride = M.L.stride; a.L = M.L; a.sh = M.sh; a.clk = M.clock.p; a.lc = M.lc.p;
    a.ix = ctx->d.ix; a.il = ctx->d.il; a.force = force;
    const int N = ctx->d.ngrid();
    dim3 grid((N + 127) / 128, ctx->nmembers);
    k_daily_forcing<<<grid, 128, 0, ctx->stream>>>(a);
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
}

void launch_clock_advance(speedy_ctx* ctx) {
    k_clock_advance<<<1, 32, 0, ctx->stream>>>(ctx->model->clock.p);
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace spd
