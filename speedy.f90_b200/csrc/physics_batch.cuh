// The grid-point column work for ENSEMBLE BATCHES as two kernels (included by physics.cu; same arithmetic, expression by
// expression, as k_grid_columns — only the decomposition differs).
//
// The level-parallel kernel ties a 320-thread CTA and 97 KB of shared memory to 32 columns for the ~9 us of its column-serial
// chain (convection -> clouds -> short wave -> surface balance -> long wave up): two tiles in flight per SM, eight level warps
// waiting on named barriers most of the time (ncu at 8 members: 8.2 barrier stalls per issue, FP64 pipe 16 % active).  That is
// the right trade at one member (144 tiles for 148 SMs: latency is all that counts).  With >= 3 members there are more tiles
// than the SMs can hold that way, and the decomposition that wins is the opposite one:
//   k_col_levels  one thread per (column, level): everything that is local to a level — thermodynamic prep, the level-local
//                 part of large-scale condensation, the long-wave source terms, all of tendencies.f90:109-197.  No shared
//                 memory beyond the level constants, 256-thread blocks of 32 columns x 8 levels, thousands of them.
//   k_col_serial  one thread per column: the vertical sweeps (convection, clouds + short wave every third step, long wave down,
//                 surface fluxes, long wave up, vertical diffusion), the pending couple_sea_land / set_forcing(1), and the closing
//                 sum of the tendencies.  One warp = one tile of 32 columns, no barrier anywhere: every column of the batch is in
//                 flight at once (8 members: 1152 warps, 7.8 per SM), so the kernel takes one column's latency, not a queue of them.
// The two exchange the level-local results through 65 rows per member in global memory (L2-resident: 2.4 MB per member).
// Every sum keeps the reference's order of operations; integer fields are bit-exact.
#pragma once

namespace spd {

enum { CS_SE = 0, CS_QSAT = CS_SE + KX, CS_RH = CS_QSAT + KX, CS_QG = CS_RH + KX, CS_DTLSC = CS_QG + KX, CS_DQLSC = CS_DTLSC + KX,
       CS_LWS0 = CS_DQLSC + KX, CS_LWS1 = CS_LWS0 + KX, CS_PSG = CS_LWS1 + KX, CS_FB = CS_PSG + 1, CS_N = CS_FB + 4 * KX };
// CS_FB: fband(nint(T_k), jb) of the four long-wave bands (longwave_radiation.f90:83,98,145,160): a table row picked by the level's own
// temperature — looked up here so that the sweeps of the serial kernel have no data-dependent address in their chains
static_assert(CS_N == COLSCR_ROWS, "scratch rows of the batch column kernels (model.h)");

// ------------------------------------------------------------------------------------------------------------------------
// level-local work: thread = (column, level)
// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KX * 32, 3) k_col_levels(const __grid_constant__ ColumnArgs a) {
    __shared__ __align__(16) double sLc[LC_DOUBLES];
    const int ix = a.ix, il = a.il, N = ix * il;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int col = blockIdx.x * TC + lane, e = blockIdx.y, j = col / ix, k = warp + 1;
    double* mb = a.base + (size_t)e * a.stride;
    const double* gin = mb + a.L.gin;
    double* gout = mb + a.L.gout;
    double* scr = mb + a.L.colscr;
    if (a.trace && tid == 0) trace_begin(a.trace, 1);
    for (int t = tid; t < LC_DOUBLES; t += KX * 32) sLc[t] = reinterpret_cast<const double*>(a.lc)[t];     // constants: ahead of the dependency wait
    const double cor = a.coriol[j];
    __syncthreads();
    const LevelConsts& lc = *reinterpret_cast<const LevelConsts*>(sLc);
    pdl_wait();                                        // the grid fields of the previous kernel are complete
    pdl_trigger();
#define SG(f) gin[(size_t)(f) * N + col]
#define SCR(r) scr[(size_t)(r) * N + col]
#define GOUT(f) gout[(size_t)(f) * N + col]
    const int nl1 = KX - 1;
    // ---------------- thermodynamic prep, physics.f90:110-122 ----------------
    const double tgk = SG(GI_T1 + k - 1), phigk = SG(GI_PHI + k - 1);
    const double psg = exp(SG(GI_PSL));
    const double qgk = dmax(SG(GI_Q1 + k - 1), 0.0);
    const double sek = lc.cp * tgk + phigk;
    const double qsatk = qsat_pt(tgk, lc.fsg[k - 1] * psg);
    const double rhk = qgk / qsatk;
    SCR(CS_SE + k - 1) = sek; SCR(CS_QSAT + k - 1) = qsatk; SCR(CS_RH + k - 1) = rhk; SCR(CS_QG + k - 1) = qgk;
    if (k == 1) SCR(CS_PSG) = psg;
    // ---------------- large_scale_condensation.f90:33-95, the level-local part ----------------
    {
        const double trlsc = 4.0, rhlsc = F32(0.9), drhlsc = F32(0.1), rhblsc = F32(0.95), qsmax = 10.0;
        const double rtlsc = 1.0 / (trlsc * 3600.0), tfact = lc.tfact;
        const double psa2 = psg * psg;
        double dtl = 0.0, dql = 0.0;
        if (k >= 2) {
            const double sig2 = lc.fsg[k - 1] * lc.fsg[k - 1];
            double rhref = rhlsc + drhlsc * (sig2 - 1.0);
            if (k == KX) rhref = dmax(rhref, rhblsc);
            const double dqmax = qsmax * sig2 * rtlsc;
            const double dqa = rhref * qsatk - qgk;
            if (dqa < 0.0) {
                dql = dqa * rtlsc;
                dtl = tfact * dmin(-dql, dqmax * psa2);
            }
        }
        SCR(CS_DTLSC + k - 1) = dtl; SCR(CS_DQLSC + k - 1) = dql;      // dql < 0 <=> this level condenses (the serial kernel's iptop update)
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
        SCR(CS_LWS0 + k - 1) = s1; SCR(CS_LWS1 + k - 1) = s2;
        const int ntk = band_row(tgk);   // nint(T) -> row of fband(100:400,:)
#pragma unroll
        for (int jb = 1; jb <= 4; jb++) SCR(CS_FB + (jb - 1) * KX + k - 1) = a.fband[ntk + 301 * (jb - 1)];
    }
    // ---------------- tendencies.f90:109-197 for level k ----------------
    if (a.mode == 0) {
        const double px = SG(GI_PX), py = SG(GI_PY);
        double umean = 0.0, vmean = 0.0, dmean = 0.0;
#pragma unroll
        for (int kk = 1; kk <= KX; kk++) {
            umean = umean + SG(GI_U + kk - 1) * lc.dhs[kk - 1];
            vmean = vmean + SG(GI_V + kk - 1) * lc.dhs[kk - 1];
            dmean = dmean + SG(GI_DIV + kk - 1) * lc.dhs[kk - 1];
        }
        if (k == 1) GOUT(GO_PSDT) = -umean * px - vmean * py;   // :125
        double sd_lo = 0.0, sm_lo = 0.0, sd_hi = 0.0, sm_hi = 0.0;    // interfaces k and k+1
        double puvk = 0.0;
        {
            double sd = 0.0, sm = 0.0;
#pragma unroll
            for (int kk = 1; kk <= KX; kk++) {
                const double puv = (SG(GI_U + kk - 1) - umean) * px + (SG(GI_V + kk - 1) - vmean) * py;
                if (kk == k) { sd_lo = sd; sm_lo = sm; puvk = puv; }
                sd = sd - lc.dhs[kk - 1] * (puv + SG(GI_DIV + kk - 1) - dmean);
                sm = sm - lc.dhs[kk - 1] * puv;
                if (kk == k) { sd_hi = sd; sm_hi = sm; }
            }
        }
        const double ugk = SG(GI_U + k - 1), vgk = SG(GI_V + k - 1), tg2k = SG(GI_T + k - 1), trgk = SG(GI_TR + k - 1);
        const double divgk = SG(GI_DIV + k - 1), vorgk = SG(GI_VOR + k - 1) + cor;   // :103-107
        const double tggk = tg2k - lc.tref[k - 1];
        const double ugm = (k > 1) ? SG(GI_U + k - 2) : 0.0, ugp = (k < KX) ? SG(GI_U + k) : 0.0;
        const double vgm = (k > 1) ? SG(GI_V + k - 2) : 0.0, vgp = (k < KX) ? SG(GI_V + k) : 0.0;
        const double tggm = (k > 1) ? SG(GI_T + k - 2) - lc.tref[k - 2] : 0.0, tggp = (k < KX) ? SG(GI_T + k) - lc.tref[k] : 0.0;
        const double trgm = (k > 1) ? SG(GI_TR + k - 2) : 0.0, trgp = (k < KX) ? SG(GI_TR + k) : 0.0;
        double t_lo, t_hi;
        t_lo = (k >= 2) ? sd_lo * (ugk - ugm) : 0.0;
        t_hi = (k < KX) ? sd_hi * (ugp - ugk) : 0.0;
        const double utend = vgk * vorgk - tggk * lc.rgas * px - (t_hi + t_lo) * lc.dhsr[k - 1];
        t_lo = (k >= 2) ? sd_lo * (vgk - vgm) : 0.0;
        t_hi = (k < KX) ? sd_hi * (vgp - vgk) : 0.0;
        const double vtend = -ugk * vorgk - tggk * lc.rgas * py - (t_hi + t_lo) * lc.dhsr[k - 1];
        t_lo = (k >= 2) ? sd_lo * (tggk - tggm) + sm_lo * (lc.tref[k - 1] - lc.tref[k - 2]) : 0.0;
        t_hi = (k < KX) ? sd_hi * (tggp - tggk) + sm_hi * (lc.tref[k] - lc.tref[k - 1]) : 0.0;
        const double ttend = tggk * divgk - (t_hi + t_lo) * lc.dhsr[k - 1] + lc.fsgr[k - 1] * tggk * (sd_hi + sd_lo) +
                             lc.tref3[k - 1] * (sm_hi + sm_lo) + lc.akap * (tg2k * puvk - tggk * dmean);
        t_lo = (k >= 4) ? sd_lo * (trgk - trgm) : 0.0;            // :192 the tracer flux is zeroed at interfaces 2 and 3
        t_hi = (k < KX && k + 1 >= 4) ? sd_hi * (trgp - trgk) : 0.0;
        const double qtend = trgk * divgk - (t_hi + t_lo) * lc.dhsr[k - 1];
        // the dynamics tendencies wait in their K2 input slots for the closing sum of the serial kernel
        const int f = GO_PER * (k - 1);
        GOUT(f + 0) = utend; GOUT(f + 1) = vtend; GOUT(f + 5) = ttend; GOUT(f + 8) = qtend;
        GOUT(f + 2) = 0.5 * (ugk * ugk + vgk * vgk);             // products for the direct transforms (tendencies.f90:219-232)
        GOUT(f + 3) = -ugk * tggk;
        GOUT(f + 4) = -vgk * tggk;
        GOUT(f + 6) = -ugk * trgk;
        GOUT(f + 7) = -vgk * trgk;
    }
#undef SG
#undef SCR
#undef GOUT
}

// ------------------------------------------------------------------------------------------------------------------------
// column-serial work: thread = column
// ------------------------------------------------------------------------------------------------------------------------
// shared rows of one warp (lane = column: a lane only ever touches its own column, so no synchronisation is needed — the rows are an
// on-chip, conflict-free home for what the sweeps index with run-time levels or read inside their dependent chains)
enum { WR_TAU2 = 0, WR_FB = WR_TAU2 + 4 * KX, WR_SE = WR_FB + 4 * KX, WR_QG = WR_SE + KX, WR_QSAT = WR_QG + KX, WR_RH = WR_QSAT + KX, WR_LWS0 = WR_RH + KX, WR_N = WR_LWS0 + KX };
constexpr int SER_WARPS = 4;
constexpr size_t SER_SMEM = sizeof(double) * ((size_t)SER_WARPS * WR_N * 32 + LC_DOUBLES);
struct RowArr {                                       // KX values of this lane's column, 1-based
    const double* p;
    __device__ __forceinline__ double operator[](int k) const { return p[(k - 1) * 32]; }
};

__global__ void __launch_bounds__(SER_WARPS * 32, 2) k_col_serial(const __grid_constant__ ColumnArgs a) {
    extern __shared__ __align__(16) double ssm[];
    double* sLc = ssm + (size_t)SER_WARPS * WR_N * 32;
    const int ix = a.ix, il = a.il, N = ix * il;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    double* wsm = ssm + (size_t)wrp * WR_N * 32 + lane;
#define WROW(r) wsm[(r) * 32]
    const int col = (blockIdx.x * SER_WARPS + wrp) * TC + lane, e = blockIdx.y, j = col / ix;
    double* mb = a.base + (size_t)e * a.stride;
    int* ib = a.ibase + (size_t)e * a.L.istride;
    const double* gin = mb + a.L.gin;
    double* gout = mb + a.L.gout;
    const double* scr = mb + a.L.colscr;
    const double* sFband = a.fband;                    // (301,4): only the four surface-temperature lookups of the upward sweep read it here
    for (int t = threadIdx.x; t < LC_DOUBLES; t += SER_WARPS * 32) sLc[t] = reinterpret_cast<const double*>(a.lc)[t];
    const double coa_j = a.coa[j];
    __syncthreads();
    const LevelConsts& lc = *reinterpret_cast<const LevelConsts*>(sLc);
    pdl_wait();                                        // k_col_levels is complete
    pdl_trigger();
#define SG(f) gin[(size_t)(f) * N + col]
#define SCR(r) scr[(size_t)(r) * N + col]
#define GOUT(f) gout[(size_t)(f) * N + col]
#define G2(off) mb[(off) + col]
#define G3(off, k) mb[(off) + (size_t)((k)-1) * N + col]
#define STAU2(k, b) WROW(WR_TAU2 + ((b)-1) * KX + ((k)-1))
#define GTAU2(k, b) mb[a.L.tau2 + ((size_t)((b)-1) * KX + ((k)-1)) * N + col]
#define TAU2W(k, b, v) do { const double v_ = (v); GTAU2(k, b) = v_; STAU2(k, b) = v_; } while (0)
#define SFB(k, b) WROW(WR_FB + ((b)-1) * KX + ((k)-1))
    const int nl1 = KX - 1, nlp = KX + 1;
    const int csw = (a.csw_override >= 0) ? a.csw_override : a.clk->csw;
    // ===== main loop only: couple_sea_land of the previous step (speedy.f90:53) and set_forcing(1) (speedy.f90:29-32)
    if (a.merged) {
        if (a.clk->slab_pending) slab_point(mb, a.L, a.sh, *a.clk, lc, N, col, 0);
        if (a.clk->do_forcing) forcing_point(mb, a.L, a.sh, *a.clk, lc, ix, il, col);
    }
    // everything the sweeps read inside their chains, fetched at once (independent loads: one memory latency)
#pragma unroll
    for (int kk = 1; kk <= KX; kk++) {
        WROW(WR_SE + kk - 1) = SCR(CS_SE + kk - 1); WROW(WR_QG + kk - 1) = SCR(CS_QG + kk - 1);
        WROW(WR_QSAT + kk - 1) = SCR(CS_QSAT + kk - 1); WROW(WR_RH + kk - 1) = SCR(CS_RH + kk - 1);
        WROW(WR_LWS0 + kk - 1) = SCR(CS_LWS0 + kk - 1);
#pragma unroll
        for (int jb = 1; jb <= 4; jb++) {
            SFB(kk, jb) = SCR(CS_FB + (jb - 1) * KX + kk - 1);
            if (!csw) STAU2(kk, jb) = GTAU2(kk, jb);   // tau2 persists between the short-wave steps (mod_radcon.f90:47); recomputed below otherwise
        }
    }
    const RowArr se{&WROW(WR_SE)}, qg{&WROW(WR_QG)}, qsat{&WROW(WR_QSAT)}, rh{&WROW(WR_RH)}, st4a1{&WROW(WR_LWS0)};
    // read where they are used (static rows: the loads are independent of the chains and issue early)
#define st4a2_(k) SCR(CS_LWS1 + (k) - 1)
#define phig_(k) SG(GI_PHI + (k) - 1)
#define dtlsc_(k) SCR(CS_DTLSC + (k) - 1)
#define dqlsc_(k) SCR(CS_DQLSC + (k) - 1)
    const double tg_kx = SG(GI_T1 + KX - 1), tg_nl1 = SG(GI_T1 + KX - 2);
    const double psg = SCR(CS_PSG);
    const double rps = 1.0 / psg;
    const double wvi2[KX + 1] = {0, lc.wvi[8], lc.wvi[9], lc.wvi[10], lc.wvi[11], lc.wvi[12], lc.wvi[13], lc.wvi[14], lc.wvi[15]};
    const double emisfc = F32(0.98), epslw = F32(0.05);

    // ---------------- convection.f90:27-245 + the LSC reductions ----------------
    int iptop = 0, icltop = 0, icnv_ = 0;
    double cloudc = 0.0, clstr = 0.0, qcloud = 0.0, precnv = 0.0, precls = 0.0;
    double dfse[KX + 1], dfqa[KX + 1];
    {
        double cbmf = 0.0;
        {
            const double psmin = F32(0.8), rhbl = F32(0.9), rhil = F32(0.7), smf = F32(0.8);
#pragma unroll
            for (int k = 1; k <= KX; k++) { dfse[k] = 0.0; dfqa[k] = 0.0; }
            // diagnose_convection :170-245
            int itop = nlp;
            double qdif = 0.0;
            if (psg > psmin) {
                const double mse0 = se[KX] + lc.alhc * qg[KX];
                double mse1 = se[nl1] + lc.alhc * qg[nl1];
                mse1 = dmin(mse0, mse1);
                const double mss0 = dmax(mse0, se[KX] + lc.alhc * qsat[KX]);
                int ktop1 = KX, ktop2 = KX;
                double msthr = 0.0;
                for (int k = KX - 3; k >= 3; k--) {
                    const double mssk = se[k] + lc.alhc * qsat[k], mssk1 = se[k + 1] + lc.alhc * qsat[k + 1];
                    const double mss2 = mssk + wvi2[k] * (mssk1 - mssk);
                    if (mss0 > mss2) ktop1 = k;
                    if (mse1 > mss2) { ktop2 = k; msthr = mss2; }
                }
                if (ktop1 < KX) {
                    const double qthr0 = rhbl * qsat[KX], qthr1 = rhbl * qsat[nl1];
                    const bool lqthr = (qg[KX] > qthr0 && qg[nl1] > qthr1);
                    if (ktop2 < KX) {
                        itop = ktop1;
                        qdif = dmax(qg[KX] - qthr0, (mse0 - msthr) * lc.ralhc);
                    } else if (lqthr) {
                        itop = ktop1;
                        qdif = qg[KX] - qthr0;
                    }
                }
            }
            if (itop != nlp) {
                const double fqmax = 5.0;
                const double fm0 = lc.fm0, rdps = lc.rdps;          // p0*dhs(kx)/(grav*trcnv*3600), 2/(1-psmin): host-evaluated
                double entr[KX + 1];
#pragma unroll
                for (int k = 2; k <= nl1; k++) entr[k] = lc.entr[k - 1]; // convection.f90:118-131, host-evaluated
                int k = KX, k1 = k - 1;
                const double qmax = dmax(F32(1.01) * qg[k], qsat[k]);
                double sb = se[k1] + wvi2[k1] * (se[k] - se[k1]);
                double qb = qg[k1] + wvi2[k1] * (qg[k] - qg[k1]);
                qb = dmin(qb, qg[k]);
                const double fpsa = psg * dmin(1.0, (psg - psmin) * rdps);
                double fmass = fm0 * fpsa * dmin(fqmax, qdif / (qmax - qb));
                cbmf = fmass;
                double fus = fmass * se[k], fuq = fmass * qmax;
                double fds = fmass * sb, fdq = fmass * qb;
                dfse[k] = fds - fus;
                dfqa[k] = fdq - fuq;
                for (k = KX - 1; k >= itop + 1; k--) {
                    k1 = k - 1;
                    dfse[k] = fus - fds;
                    dfqa[k] = fuq - fdq;
                    const double enmass = entr[k] * psg * cbmf;
                    fmass = fmass + enmass;
                    fus = fus + enmass * se[k];
                    fuq = fuq + enmass * qg[k];
                    sb = se[k1] + wvi2[k1] * (se[k] - se[k1]);
                    qb = qg[k1] + wvi2[k1] * (qg[k] - qg[k1]);
                    fds = fmass * sb;
                    fdq = fmass * qb;
                    dfse[k] = dfse[k] + fds - fus;
                    dfqa[k] = dfqa[k] + fdq - fuq;
                    const double delq = rhil * qsat[k] - qg[k];
                    if (delq > 0.0) {
                        const double fsq = smf * cbmf * delq;
                        dfqa[k] = dfqa[k] + fsq;
                        dfqa[KX] = dfqa[KX] - fsq;
                    }
                }
                k = itop;
                const double qsatb = qsat[k] + wvi2[k] * (qsat[k + 1] - qsat[k]);
                precnv = dmax(fuq - fmass * qsatb, 0.0);
                dfse[k] = fus - fds + lc.alhc * precnv;
                dfqa[k] = fuq - fdq - precnv;
            }
            iptop = itop;
            // physics.f90:127-138 (level 1 is not rescaled)
#pragma unroll
            for (int k = 2; k <= KX; k++) {
                dfse[k] = dfse[k] * rps * lc.grdscp[k - 1];
                dfqa[k] = dfqa[k] * rps * lc.grdsig[k - 1];
            }
        }
        icnv_ = KX - iptop;   // physics.f90:132, before LSC lowers iptop
        ib[a.L.icnv + col] = icnv_;
        {   // large_scale_condensation.f90:60-93: cloud top and precipitation from the level-local results
            const double prg = lc.prg;
#pragma unroll
            for (int kk = 2; kk <= KX; kk++) if (dqlsc_(kk) < 0.0) iptop = min(kk, iptop);
#pragma unroll
            for (int kk = 2; kk <= KX; kk++) {
                const double pfact = lc.dhs[kk - 1] * prg;
                precls = precls - pfact * dqlsc_(kk);
            }
            precls = precls * psg;
        }
        G2(a.L.precnv) = precnv; G2(a.L.precls) = precls; G2(a.L.cbmf) = cbmf;
        ib[a.L.iptop + col] = iptop;
    }
    // surface / forcing fields of the column (after the pending slab update and the daily forcing above)
    const double s_fmask = G2(a.L.fmask_l), s_fsol = G2(a.L.fsol), s_ozone = G2(a.L.ozone), s_ozupp = G2(a.L.ozupp), s_zenit = G2(a.L.zenit),
                 s_stratz = G2(a.L.stratz), s_albsfc = G2(a.L.albsfc), s_phis0 = G2(a.L.phis0), s_sst = G2(a.L.sst_am), s_stl = G2(a.L.stl_am),
                 s_soilw = G2(a.L.soilw_am), s_albl = G2(a.L.alb_l), s_albs = G2(a.L.alb_s), s_snowc = G2(a.L.snowc), s_forog = G2(a.L.forog);
    double s_ssrd = G2(a.L.ssrd);
    double rsw[KX + 1], stratc1, stratc2;

    // ---------------- shortwave (every nstrad-th step), shortwave_radiation.f90 ----------------
    if (csw) {
        const double albcl = F32(0.43), albcls = 0.50;
        const double absdry = F32(0.033), absaer = F32(0.033), abswv1 = F32(0.022), abswv2 = 15.000, abscl1 = F32(0.015), abscl2 = F32(0.15);
        const double ablwin = F32(0.3), ablco2 = 6.0, ablwv1 = F32(0.7), ablwv2 = 50.0, ablcl1 = 12.0, ablcl2 = F32(0.6);
        {
            // clouds  shortwave_radiation.f90:332-410
            const double rhcl1 = F32(0.30), rhcl2 = 1.00, qacl = F32(0.20), wpcl = F32(0.2), pmaxcl = 10.0;
            const double clsmax = F32(0.60), clsminl = F32(0.15), gse_s0 = 0.25, gse_s1 = F32(0.40);
            const double gse = (se[KX - 1] - se[KX]) / (phig_(KX - 1) - phig_(KX));   // physics.f90:147
            const double rrcl = 1. / (rhcl2 - rhcl1);
            if (rh[nl1] > rhcl1) { cloudc = rh[nl1] - rhcl1; icltop = nl1; }
            else { cloudc = 0.0; icltop = nlp; }
            for (int kk = 3; kk <= KX - 2; kk++) {
                const double drh = rh[kk] - rhcl1;
                if (drh > cloudc && qg[kk] > qacl) { cloudc = drh; icltop = kk; }
            }
            {
                const double pr1 = dmin(pmaxcl, F32(86.4) * (precnv + precls));
                const double cc = dmin(1.0, cloudc * rrcl);
                cloudc = dmin(1.0, wpcl * sqrt(pr1) + cc * cc);
                icltop = min(iptop, icltop);
            }
            qcloud = qg[nl1];
            {
                const double clfact = F32(1.2), rgse = 1.0 / (gse_s1 - gse_s0);
                const double fstab = dmax(0.0, dmin(1.0, rgse * (gse - gse_s0)));
                clstr = fstab * dmax(clsmax - clfact * cloudc, 0.0);
                const double clstrl = dmax(clstr, clsminl) * rh[KX];
                const double fm = s_fmask;
                clstr = clstr + fm * (clstrl - clstr);
            }
            ib[a.L.icltop + col] = icltop;
            G2(a.L.qcloud) = qcloud; G2(a.L.cloudc) = cloudc; G2(a.L.clstr) = clstr;
        }
        // ---- every transmissivity of every level (shortwave_radiation.f90:130-150, 190-233): 48 independent exp()
        double tau1[KX + 1], tau2_[KX + 1], tau3[KX + 1], dfabs[KX + 1];
        {
            const double cloudc_ = cloudc, qcloud_ = qcloud;
            const int icltop_ = icltop;
            const double zenit = s_zenit;
            const double psaz = psg * zenit;
#pragma unroll
            for (int k = 1; k <= KX; k++) {
                const double qgk = qg[k];
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
                tau1[k] = t1;
                tau2_[k] = (k == 1) ? 0.0 : exp(-psaz * lc.dhs[k - 1] * abswv2 * qgk);
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
        }
        {
            // get_shortwave_rad_fluxes  shortwave_radiation.f90:74-234: the flux sweeps
            const double fsol = s_fsol, ozone = s_ozone, ozupp = s_ozupp, stratz = s_stratz;
            const double albsfc = s_albsfc;
            const double fband2 = F32(0.05), fband1 = 1.0 - fband2;
#pragma unroll
            for (int kk = 1; kk <= KX; kk++) tau3[kk] = 0.0;
            if (icltop <= KX) tau3[icltop] = albcl * cloudc;
            tau3[KX] = albcls * clstr;
            double ftop = fsol;
            double flux1 = fsol * fband1, flux2 = fsol * fband2;
            dfabs[1] = flux1;
            flux1 = tau1[1] * (flux1 - ozupp * psg);
            dfabs[1] = dfabs[1] - flux1;
            dfabs[2] = flux1;
            flux1 = tau1[2] * (flux1 - ozone * psg);
            dfabs[2] = dfabs[2] - flux1;
            for (int kk = 3; kk <= KX; kk++) {
                tau3[kk] = flux1 * tau3[kk];
                flux1 = flux1 - tau3[kk];
                dfabs[kk] = flux1;
                flux1 = tau1[kk] * flux1;
                dfabs[kk] = dfabs[kk] - flux1;
            }
            for (int kk = 2; kk <= KX; kk++) {
                dfabs[kk] = dfabs[kk] + flux2;
                flux2 = tau2_[kk] * flux2;
                dfabs[kk] = dfabs[kk] - flux2;
            }
            const double fsfcd = flux1 + flux2;
            flux1 = flux1 * albsfc;
            const double fsfc = fsfcd - flux1;
            for (int kk = KX; kk >= 1; kk--) {
                dfabs[kk] = dfabs[kk] + flux1;
                flux1 = tau1[kk] * flux1;
                dfabs[kk] = dfabs[kk] - flux1;
                flux1 = flux1 + tau3[kk];
            }
            ftop = ftop - flux1;
            G2(a.L.ssrd) = fsfcd; s_ssrd = fsfcd; G2(a.L.ssr) = fsfc; G2(a.L.tsr) = ftop;
#pragma unroll
            for (int kk = 1; kk <= KX; kk++) { const double v = dfabs[kk] * rps * lc.grdscp[kk - 1]; G3(a.L.tt_rsw, kk) = v; rsw[kk] = v; }   // physics.f90:160-162
            const double eps1 = lc.eps1;
            mb[a.L.stratc + col] = stratc1 = stratz * psg;
            mb[a.L.stratc + N + col] = stratc2 = eps1 * psg;
        }
    } else {
#pragma unroll
        for (int kk = 1; kk <= KX; kk++) rsw[kk] = G3(a.L.tt_rsw, kk);       // tt_rsw persists between the short-wave steps (physics.f90:79,186)
        stratc1 = mb[a.L.stratc + col]; stratc2 = mb[a.L.stratc + N + col];
    }

    // ------------------- downward longwave  longwave_radiation.f90:16-117 -------------------
    double tt_rlw[KX + 1], flux[5];
    double slrd;
    {
        double fsfcd = 0.0;
#pragma unroll
        for (int k = 1; k <= KX; k++) tt_rlw[k] = 0.0;
        {
            for (int jb = 1; jb <= 2; jb++) {
                const double emis = 1.0 - STAU2(1, jb);
                const double brad = SFB(1, jb) * (st4a1[1] + emis * st4a2_(1));
                flux[jb] = emis * brad;
                tt_rlw[1] = tt_rlw[1] - flux[jb];
            }
        }
        flux[3] = 0.0; flux[4] = 0.0;
        {
            double f1 = flux[1], f2 = flux[2], f3 = flux[3], f4 = flux[4];
#pragma unroll
            for (int k = 2; k <= KX; k++) {
                const double s1 = st4a1[k], s2 = st4a2_(k);
                double t = 0.0;
#define LW_BAND(fl, jb)                                                           \
    {                                                                             \
        const double tau = STAU2(k, jb);                                          \
        const double emis = 1.0 - tau;                                            \
        const double brad = SFB(k, jb) * (s1 + emis * s2);                        \
        t = t + fl;                                                               \
        fl = tau * fl + emis * brad;                                              \
        t = t - fl;                                                               \
    }
                LW_BAND(f1, 1) LW_BAND(f2, 2) LW_BAND(f3, 3) LW_BAND(f4, 4)
#undef LW_BAND
                tt_rlw[k] = t;
            }
            flux[1] = f1; flux[2] = f2; flux[3] = f3; flux[4] = f4;
        }
        for (int jb = 1; jb <= 4; jb++) fsfcd = fsfcd + emisfc * flux[jb];
        const double corlw = epslw * emisfc * st4a1[KX];
        tt_rlw[KX] = tt_rlw[KX] - corlw;
        fsfcd = fsfcd + corlw;
        slrd = fsfcd;
        G2(a.L.slrd) = slrd;
    }

    // ---------------- surface_fluxes.f90:42-296 (lfluxland = .true.): shared terms, sea half, land half, blend ----------------
    const double fwind0 = F32(0.95), ftemp0 = 1.0, cdl = F32(2.4e-3), cds = F32(1.0e-3), chl = F32(1.2e-3), chs = F32(0.9e-3);
    const double vgust = 5.0, ctday = F32(1.0e-2), dtheta = 3.0, fstab = F32(0.67), clambda = 7.0, clambsn = 7.0;
    const double esbc = emisfc * lc.sbc;
    const double rdth = fstab / dtheta, astab = 0.5;
    double ts, shf3, evap3, ustr3, vstr3, slru3;
    {
        const double ug8 = SG(GI_U1 + KX - 1), vg8 = SG(GI_V1 + KX - 1);   // only the lowest-level wind is used
        const double tg8 = tg_kx, tg7 = tg_nl1, qg8 = qg[KX], phig8 = phig_(KX);
        const double phi0 = s_phis0, fmask = s_fmask, tsea = s_sst;
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
        // land half and the blend
        const double stl_am = s_stl;
        const double soilw_am = s_soilw, alb_l = s_albl, alb_s = s_albs, snowc = s_snowc, forog = s_forog;
        const double ssrd = s_ssrd;
        double tskin = stl_am + ctday * sqrt(coa_j) * ssrd * (1.0 - alb_l) * psg;
        double dthl;
        if (tskin > t2_1) dthl = dmin(dtheta, tskin - t2_1);
        else dthl = dmax(-dtheta, astab * (tskin - t2_1));
        const double denvvs1 = denvvs0 * (1.0 + dthl * rdth);
        const double cdldv = cdl * denvvs0 * forog;
        const double ustr1 = -cdldv * ug8, vstr1 = -cdldv * vg8;
        const double chlcp = chl * lc.cp;
        double shf1 = chlcp * denvvs1 * (tskin - t1_1);
        const double q1_1 = qg[KX];
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
        const double hfluxn2 = ssrd * (1.0 - alb_s) + slrd - slru2 + shf2 + lc.alhc * evap2;
        ustr3 = ustr2 + fmask * (ustr1 - ustr2);
        vstr3 = vstr2 + fmask * (vstr1 - vstr2);
        shf3 = shf2 + fmask * (shf1 - shf2);
        evap3 = evap2 + fmask * (evap1 - evap2);
        slru3 = slru2 + fmask * (slru1 - slru2);
        ts = tsea + fmask * (stl_am - tsea);
        tskin = tsea + fmask * (tskin - tsea);
        const double t0b = t1_2 + fmask * (t1_1 - t1_2);
        mb[a.L.ustr + col] = ustr1; mb[a.L.ustr + N + col] = ustr2; mb[a.L.ustr + 2 * N + col] = ustr3;
        mb[a.L.vstr + col] = vstr1; mb[a.L.vstr + N + col] = vstr2; mb[a.L.vstr + 2 * N + col] = vstr3;
        mb[a.L.shf + col] = shf1; mb[a.L.shf + N + col] = shf2; mb[a.L.shf + 2 * N + col] = shf3;
        mb[a.L.evap + col] = evap1; mb[a.L.evap + N + col] = evap2; mb[a.L.evap + 2 * N + col] = evap3;
        mb[a.L.slru + col] = slru1; mb[a.L.slru + N + col] = slru2; mb[a.L.slru + 2 * N + col] = slru3;
        mb[a.L.hfluxn + col] = hfluxn1; mb[a.L.hfluxn + N + col] = hfluxn2;
        G2(a.L.ts) = ts; G2(a.L.tskin) = tskin; G2(a.L.u0) = u0; G2(a.L.v0) = v0; G2(a.L.t0) = t0b;
    }
    // ------------------- upward longwave  longwave_radiation.f90:120-194 -------------------
    {
        const double refsfc = 1.0 - emisfc;
        const double fsfcu = slru3;
        G2(a.L.slr) = fsfcu - slrd;
        const int nts = band_row(ts);
        for (int jb = 1; jb <= 4; jb++) flux[jb] = sFband[nts + 301 * (jb - 1)] * fsfcu + refsfc * flux[jb];
        tt_rlw[KX] = tt_rlw[KX] + epslw * fsfcu;
        {
            double f1 = flux[1], f2 = flux[2], f3 = flux[3], f4 = flux[4];
#pragma unroll
            for (int k = KX; k >= 2; k--) {     // longwave_radiation.f90:155-167
                const double s1 = st4a1[k], s2 = st4a2_(k);
                double t = tt_rlw[k];
#define LW_BAND(fl, jb)                                                           \
    {                                                                             \
        const double tau = STAU2(k, jb);                                          \
        const double emis = 1.0 - tau;                                            \
        const double brad = SFB(k, jb) * (s1 - emis * s2);                        \
        t = t + fl;                                                               \
        fl = tau * fl + emis * brad;                                              \
        t = t - fl;                                                               \
    }
                LW_BAND(f1, 1) LW_BAND(f2, 2) LW_BAND(f3, 3) LW_BAND(f4, 4)
#undef LW_BAND
                tt_rlw[k] = t;
            }
            flux[1] = f1; flux[2] = f2; flux[3] = f3; flux[4] = f4;
        }
        {
            double t = tt_rlw[1];
            for (int jb = 1; jb <= 2; jb++) {
                const double tau = STAU2(1, jb);
                const double emis = 1.0 - tau;
                const double brad = SFB(1, jb) * (st4a1[1] - emis * st4a2_(1));
                t = t + flux[jb];
                flux[jb] = tau * flux[jb] + emis * brad;
                t = t - flux[jb];
            }
            tt_rlw[1] = t;
        }
        const double corlw1 = lc.dhs[0] * stratc2 * st4a1[1] + stratc1;
        const double corlw2 = lc.dhs[1] * stratc2 * st4a1[2];
        tt_rlw[1] = tt_rlw[1] - corlw1;
        tt_rlw[2] = tt_rlw[2] - corlw2;
        double ftop = corlw1 + corlw2;
        for (int jb = 1; jb <= 4; jb++) ftop = ftop + flux[jb];
        G2(a.L.olr) = ftop;
#pragma unroll
        for (int k = 1; k <= KX; k++) tt_rlw[k] = tt_rlw[k] * rps * lc.grdscp[k - 1];   // physics.f90:182-186, added in the closing stage
    }
    // ------------------- vertical_diffusion.f90:30-143 -------------------
    double ttenvd[KX + 1], qtenvd[KX + 1];
    {
        const double trshc = 6.0, trvdi = 24.0, trvds = 6.0, redshc = 0.5, rhgrad = 0.5, segrad = F32(0.1);
        double rsig[KX + 1], rsig1[KX + 1];
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
            const double dmse = se[KX] - se[nl1] + lc.alhc * (qg[KX] - qsat[nl1]);
            const double drh = rh[KX] - rh[nl1];
            double fcnv = 1.0;
            if (dmse >= 0.0) {
                if (icnv_ > 0) fcnv = redshc;
                const double fluxse = fcnv * fshcse * dmse;
                ttenvd[nl1] = fluxse * rsig[nl1];
                ttenvd[KX] = -fluxse * rsig[KX];
                if (drh >= 0.0) {
                    const double fluxq = fcnv * fshcq * qsat[KX] * drh;
                    qtenvd[nl1] = fluxq * rsig[nl1];
                    qtenvd[KX] = -fluxq * rsig[KX];
                }
            } else if (drh > drh0) {
                const double fluxq = fvdiq2 * qsat[nl1] * drh;
                qtenvd[nl1] = fluxq * rsig[nl1];
                qtenvd[KX] = -fluxq * rsig[KX];
            }
        }
#pragma unroll
        for (int k = 3; k <= KX - 2; k++) {
            if (lc.sigh[k] > 0.5) {
                drh0 = rhgrad * (lc.fsg[k] - lc.fsg[k - 1]);
                fvdiq2 = fvdiq * lc.sigh[k];
                const double drh = rh[k + 1] - rh[k];
                if (drh >= drh0) {
                    const double fluxq = fvdiq2 * qsat[k] * drh;
                    qtenvd[k] = qtenvd[k] + fluxq * rsig[k];
                    qtenvd[k + 1] = qtenvd[k + 1] - fluxq * rsig[k + 1];
                }
            }
        }
#pragma unroll
        for (int k = 1; k <= nl1; k++) {
            const double se0 = se[k + 1] + segrad * (phig_(k) - phig_(k + 1));
            if (se[k] < se0) {
                const double fluxse = fvdise * (se0 - se[k]);
                ttenvd[k] = ttenvd[k] + fluxse * rsig[k];
#pragma unroll
                for (int k1 = k + 1; k1 <= KX; k1++) ttenvd[k1] = ttenvd[k1] - fluxse * rsig1[k];
            }
        }
    }
    // ---------------- closing: physics.f90:137-138, 182-186, 197-205, 208-222 ----------------
    const double ut8 = 0.0 + ustr3 * rps * lc.grdsig[KX - 1];
    const double vt8 = 0.0 + vstr3 * rps * lc.grdsig[KX - 1];
    const double shft = shf3 * rps * lc.grdscp[KX - 1];
    const double evapt = evap3 * rps * lc.grdsig[KX - 1];
#pragma unroll
    for (int k = 1; k <= KX; k++) {
        const int f0 = GO_PER * (k - 1);
        const double ut_dyn = GOUT(f0 + 0), vt_dyn = GOUT(f0 + 1), tt_dyn = GOUT(f0 + 5), qt_dyn = GOUT(f0 + 8);
        double ut = ut_dyn, vt = vt_dyn, tt = tt_dyn, qt = qt_dyn;
        tt = tt + dfse[k] + dtlsc_(k);
        qt = qt + dfqa[k] + dqlsc_(k);
        tt = tt + rsw[k] + tt_rlw[k];
        double tvd = ttenvd[k], qvd = qtenvd[k];
        if (k == KX) {
            tvd = tvd + shft;
            qvd = qvd + evapt;
            ut = ut + ut8; vt = vt + vt8;
        } else {
            ut = ut + 0.0; vt = vt + 0.0;
        }
        tt = tt + tvd; qt = qt + qvd;
        if (a.sppt_on) {
            double p = SG(GI_SPPT + k - 1);
            p = dmin(1.0, fabs(p)) * copysign(1.0, p);   // sppt.f90:98
            const double f = (1 + p * 1.0);
            ut = f * (ut - ut_dyn) + ut_dyn;
            vt = f * (vt - vt_dyn) + vt_dyn;
            tt = f * (tt - tt_dyn) + tt_dyn;
            qt = f * (qt - qt_dyn) + qt_dyn;
        }
        GOUT(f0 + 0) = ut; GOUT(f0 + 1) = vt; GOUT(f0 + 5) = tt; GOUT(f0 + 8) = qt;
    }
    if (a.trace && lane == 0) trace_end(a.trace, 1);
#undef SG
#undef SCR
#undef GOUT
#undef G2
#undef G3
#undef STAU2
#undef GTAU2
#undef TAU2W
#undef SFB
#undef st4a2_
#undef phig_
#undef dtlsc_
#undef dqlsc_
#undef WROW
}

}  // namespace spd
