// Batch ("quad") variants of the spherical-harmonic transforms for ensemble batches at T30: four fields at a time, so that
// the Legendre sums become FP64 tensor-core contractions (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4; tcgen05 has no FP64 kind)
// with the 8 columns of a tile = (re, im) x 4 fields, and the zonal transform is FFTPACK's own FFT (fft96f.cuh / fft96.cuh),
// not the dense operator.
//
//   k_g2s_quad  grid->spec = fourier_dir (fourier.f90:56-82) + legendre_dir (legendre.f90:114-155), incl. the cosgr / cosgr2
//               pre-scale of vdspec (spectral.f90:208-222).  One persistent CTA per SM owns WHOLE fields:
//     * the field arrives as tensor-map boxes of [16 longitudes x 8 latitudes] with the 128-byte swizzle, per band of 8
//       latitude pairs, through a ring of band buffers (two fields in flight per SM);
//     * stage A folds the two hemispheres while loading — the FFT is linear, so the Gaussian-weighted even / odd folds of
//       legendre.f90:127-133 and fourier_dir's 1/ix are applied to the grid rows and 48 FOLDED rows are transformed —
//       and runs radf3 + radf4 in registers; stage B (radf4 + radf2) writes the Fourier coefficients of the fold straight
//       into the quad buffer EO[field][coefficient row][folded row] — the B operand of the contraction, conflict-free;
//     * after four fields the direct Legendre sums of all wavenumbers run as DMMA tiles  out(n, (field, re/im)) =
//       sum_j P(m, n, j) * EO(j, (field, re/im))  with the P fragments of every tile resident in REGISTERS for the whole
//       kernel (loaded once per CTA from a fragment-ordered copy of the table): the only shared-memory traffic of the
//       Legendre stage is one 8-byte B element per lane and DMMA.
//   Rounding: the fold is taken before the FFT and the tile sums accumulate in DMMA order; per call the result agrees with
//   the reference order to ~1e-15 relative (tests: 1e-12 per call, 1e-10 after 48 h).
#include "ctx.h"
#include "tma.cuh"
#include "member_ready.cuh"
#include <cuda.h>
#include <map>
#include <mutex>
#include <tuple>
#include "fft96f.cuh"
#include "fft96.cuh"
#include "spectral_ops.cuh"
#include "close_step.cuh"
#include <algorithm>

namespace spd {

__device__ __forceinline__ void dmma884q(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct QCfg {
    static constexpr int TRUNC = 30, MX = 31, NX = 32, IX = 96, IL = 48, IY = 24, K2 = 2 * MX, NSPEC2 = NX * K2;
    static constexpr int WARPS = 24, THREADS = 32 * WARPS;
    static constexpr int NG = 3, GT = 128;           // latitude bands; threads per FFT group (a warpgroup)
    static constexpr int NPH = 2;                    // fields in the FFT stages at a time: group (band, phase) takes every NPH-th field
    static constexpr int RING = 3;                   // band buffers per group: one being transformed, two in flight
    static constexpr int NBOX = IX / 16;             // tensor-map boxes of [16 longitudes x 8 latitudes] = one 1 KB swizzle atom
    static constexpr int BAND = 2 * NBOX * 128;      // doubles per band buffer: [hemisphere][box][8 rows][16 longitudes]
    static constexpr int CS = 52, FS = K2 * CS;      // EO[field][c][slot]; CS = 4 mod 16 and FS = 8 mod 16: conflict-free B fragments
    static constexpr int SLOTS = 4, KS = IY / 4;     // tiles per warp, k-steps per tile
    static constexpr int TCOLS = 2 * SLOTS * KS;     // 32-bit TMEM columns of a warp's P fragments
    static constexpr int NBAR = RING * NPH;          // "band landed" barriers per band: field seq uses buffer seq % RING and barrier seq % NBAR, so
                                                     // that each barrier has ONE waiting group (phase parity is only safe for a waiter that cannot fall behind)
    static constexpr int NLIVE = 93, NTILE = 124;    // live / all (m, parity, n-tile) tiles
    static constexpr int LCAP = 128;                 // fields per CTA and launch (the host splits larger batches)
    static constexpr int RSTG = K2, FSTG = NX * RSTG + 2;   // output staging [field][n][K2]: a field is one contiguous run (one bulk store); field stride = 1 mod 8 sixteen-byte slots
    static constexpr size_t SMEM_G2S = sizeof(double) * (NG * RING * BAND + 4 * FS + IX) + sizeof(uint64_t) * (NG * NBAR + 2) + sizeof(int) * NTILE +
                                       LCAP * (3 * sizeof(int) + sizeof(long long));
    static_assert(NG * NPH * GT == THREADS && (WARPS / 4) * TCOLS <= 512 && TCOLS % 16 == 0, "FFT groups; tensor-memory columns");
    static_assert(FS % 16 == 8 && CS % 16 == 4 && (BAND * 8) % 1024 == 0 && BAND == IX * 16 && SMEM_G2S <= 232448, "layout");
    static_assert(WARPS * SLOTS >= NLIVE && NG * 8 == IY && SLOTS % 4 == 0, "tiling");
    static_assert((FSTG / 2) % 8 == 1 && RSTG == K2 && 4 * FSTG <= 4 * FS && (NSPEC2 * 8) % 16 == 0, "staging layout");
};


// host: tile list + P fragments in (warp, tile slot, k-step, lane) order
void build_quad_tables(const Tables& t, std::vector<int>& tiles, std::vector<double>& polyq) {
    using C = QCfg;
    tiles.assign(C::NTILE, 0);
    int nl = 0, nd = C::NLIVE;
    for (int m = 0; m < C::MX; m++)
        for (int p = 0; p < 2; p++)
            for (int tt = 0; tt < 2; tt++) {
                const int nmin = p + 16 * tt;
                const bool live = nmin <= C::TRUNC && m + nmin <= C::MX;
                const int packed = m | (p << 8) | (tt << 9) | (live ? 1 << 10 : 0);
                if (live) tiles[nl++] = packed; else tiles[nd++] = packed;
            }
    if (nl != C::NLIVE || nd != C::NTILE) throw std::runtime_error("quad tile enumeration");
    // the quad kernel scales the FOLDED rows by cosgr / cosgr2 (vdspec): the two hemispheres must carry the same factor
    for (int j = 0; j < C::IY; j++)
        if (t.cosgr[j] != t.cosgr[C::IL - 1 - j] || t.cosgr2[j] != t.cosgr2[C::IL - 1 - j]) throw std::runtime_error("cosgr is not symmetric about the equator");
    const double scale = (double)(1.0f / (float)C::IX);                 // fourier.f90:72, real32 quotient
    polyq.assign((size_t)C::WARPS * C::SLOTS * C::KS * 32, 0.0);
    for (int w = 0; w < C::WARPS; w++)
        for (int i = 0; i < C::SLOTS; i++) {
            const int L = i * C::WARPS + w;
            if (L >= C::NLIVE) continue;
            const int m = tiles[L] & 255, p = (tiles[L] >> 8) & 1, tt = (tiles[L] >> 9) & 1;
            for (int ks = 0; ks < C::KS; ks++)
                for (int lane = 0; lane < 32; lane++) {
                    const int g = lane >> 2, q = lane & 3;
                    const int n = p + 2 * (8 * tt + g), jh = 4 * ks + q;
                    const bool valid = n <= C::TRUNC && m + n <= C::MX;
                    // the Gaussian weight of the fold (legendre.f90:131-132) and fourier_dir's 1/ix ride in the P fragment
                    // fragments are stored in pairs (j, j + 1) per lane: one 16-byte load per lane, 512 contiguous bytes per warp
                    const int j = i * C::KS + ks;
                    polyq[((((size_t)w * C::SLOTS * C::KS) / 2 + j / 2) * 32 + lane) * 2 + (j & 1)] = valid ? t.poly[((size_t)jh * C::NX + n) * C::MX + m] * (t.wt[jh] * scale) : 0.0;
                }
        }
}

// One CTA = three warpgroups, one per band of 8 latitude pairs.  A group runs the whole FFT of its band's 16 folded rows by
// itself (its own ring of band buffers, its own named barriers), so the three groups hide each other's latencies and the CTA
// only synchronises around the Legendre phase of a quad:
//   wait band -> load + fold (registers) | group barrier | stage A, written IN PLACE over the band buffer as T[pos][16 slots]
//   (slot index XOR-swizzled with the 12-block of the position: conflict-free for both stages) | group barrier | stage B -> EO
//   | buffer handed back to the TMA for the band three fields ahead.
// A band of a field = two tensor-map boxes (southern and northern rows), each [8 latitudes][6 x 16 longitudes] through a 4-D view
// of the grid array (16 longitudes | 6 blocks of 16 | row | member): 128-byte segments, hardware 128-byte swizzle.
// The spectral coefficients leave through shared memory: the tile results are staged over the (dead) EO buffer in the output
// layout and written as whole 128-byte lines (a lane's own result is 16 bytes at a 496-byte stride: 32 sectors per store).
// tensor memory as a scratchpad for the P fragments: 32x32b shape, thread i of warp w owns TMEM lane 32 (w % 4) + i
__device__ __forceinline__ void tmem_st16(uint32_t addr, const double* v) {     // 8 doubles -> 16 columns
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { r[2 * i] = (uint32_t)__double2loint(v[i]); r[2 * i + 1] = (uint32_t)__double2hiint(v[i]); }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                    "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, double* v) {           // issue only: tmem_ld_wait() before the values are used
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                   "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(addr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}

// per-CTA timeline of the last traced launch (speedy_trace + SPEEDY_TRACE_STAMPS): %globaltimer at [start, first quad, after the quads, end]
#define CSTAMP(kslot, i) do { if (tv.trace && tid == 0 && blockIdx.x < 160) tv.trace[64 + 640 * (kslot) + 4 * blockIdx.x + (i)] = gtimer(); } while (0)

__global__ void __launch_bounds__(QCfg::THREADS, 1)
k_g2s_quad(const __grid_constant__ CUtensorMap gmap, const XDesc* __restrict__ desc, int nbatch, int idx_base, int idx_end, int unit,
           double* __restrict__ out_base, long long out_ms, DevTables tv, const int* __restrict__ gate, const double* __restrict__ dead_in, long long in_ms, long long out_fs) {
    using C = QCfg;
    extern __shared__ __align__(1024) double smem[];
    double* sRing = smem;                               // [NG][RING][BAND]
    double* sEO = sRing + C::NG * C::RING * C::BAND;    // [4][K2][CS]; after the tile sums: output staging [4][NX][RSTG]
    double* sWa = sEO + 4 * C::FS;                      // [IX] twiddles
    uint64_t* bars = reinterpret_cast<uint64_t*>(sWa + C::IX);     // [NG][NBAR] "band landed", then "EO free"; the last slot holds the TMEM base address
    uint64_t* eo_free = bars + C::NG * C::NBAR;
    long long* sOut = reinterpret_cast<long long*>(bars + C::NG * C::NBAR + 2);   // [LCAP] output offset of the field
    int* sRow0 = reinterpret_cast<int*>(sOut + C::LCAP);           // [LCAP] first row of the field in the tensor map
    int* sE = sRow0 + C::LCAP;                                      // [LCAP] member
    int* sFl = sE + C::LCAP;                                        // [LCAP] descriptor flags
    int* sTile = sFl + C::LCAP;
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(bars + C::NG * C::NBAR + 1);
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    if (tid == 0) trace_begin(tv.trace, 2);
    CSTAMP(1, 0);
    // this CTA's run of the flattened (member, field) list, in units of four fields (quads) — of two when there are fewer quads than SMs
    const int nunit = (idx_end - idx_base + unit - 1) / unit;
    const int q0 = (int)((long long)blockIdx.x * nunit / gridDim.x), q1 = (int)((long long)(blockIdx.x + 1) * nunit / gridDim.x);
    const int i0 = idx_base + unit * q0, cnt = min(idx_base + unit * q1, idx_end) - i0;
    if (tid == 0 && ((smem_u32(sRing) & 1023u) || cnt > C::LCAP)) __trap();
    if (tid == 0) {
        for (int k = 0; k < C::NG * C::NBAR + 1; k++) mbar_init(&bars[k], 1);
        mbar_fence_init();
    }
    if (w == 0) {                                       // all 512 columns of the SM's tensor memory (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(sTmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- prologue on constant tables only (may overlap the tail of the previous kernel: PDL)
    for (int t = tid; t < 2 * C::FS; t += C::THREADS) reinterpret_cast<double2*>(sEO)[t] = make_double2(0.0, 0.0);   // row c = 1 (Im m = 0, fourier.f90:76) stays zero
    for (int t = tid; t < C::IX; t += C::THREADS) sWa[t] = tv.fftwa[t];
    for (int t = tid; t < C::NTILE; t += C::THREADS) sTile[t] = tv.qtile[t];
    for (int t = tid; t < cnt; t += C::THREADS) {       // the descriptors are constants: the field list costs no round trip after the wait
        const int idx = i0 + t, e = idx / nbatch, f = idx - e * nbatch;
        const XDesc d = desc[f];
        const int row0 = (int)(d.off / C::IX);
        if (d.off != (long long)row0 * C::IX) __trap();
        sRow0[t] = row0; sE[t] = e; sFl[t] = d.flags; sOut[t] = (long long)e * out_ms + (long long)f * out_fs;     // out_fs > 2 nspec: each field's coefficients over its own (consumed) grid rows
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // P fragments of this warp's tiles: parked in tensor memory for the whole kernel (they would take 48 registers per thread of
    // a 768-thread CTA), fetched back at the start of every Legendre phase
    const uint32_t taddr = *sTmem + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)(C::TCOLS * (w >> 2));
    {
        double a[C::SLOTS * C::KS];
#pragma unroll
        for (int j = 0; j < C::SLOTS * C::KS; j += 2) {
            const double2 v = reinterpret_cast<const double2*>(tv.polyq)[((size_t)w * (C::SLOTS * C::KS / 2) + j / 2) * 32 + lane];
            a[j] = v.x; a[j + 1] = v.y;
        }
#pragma unroll
        for (int c = 0; c < C::TCOLS / 16; c++) tmem_st16(taddr + 16 * c, a + 8 * c);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    // roles: FFT group = (latitude band, field phase); stage A thread = (parity, latitude pair of the band, k of radf4); stage B thread = (set, slot)
    const int gi = w >> 2, b = gi % C::NG, ph = gi / C::NG, wl = w & 3, par = wl >> 1, ha = wl & 1;
    const int k3 = (lane & 3) + 4 * (lane >> 4), rsel = (lane >> 2) & 3;
    const int jl = 4 * ha + rsel, jh = 8 * b + jl;      // a warp reads 4 consecutive rows: distinct swizzle phases, conflict-free
    const int s16A = (8 * par + jl) ^ (4 * (k3 & 3));
    // element (row r, longitude 16 bx + l16) of a hemisphere block sits in 128-byte segment 6 r + bx, 16-byte chunk (l16 / 2) ^ (segment & 7)
    int oS[C::NBOX], oN[C::NBOX];
#pragma unroll
    for (int bx = 0; bx < C::NBOX; bx++) {
        const int sgS = 6 * jl + bx, sgN = 6 * (7 - jl) + bx;
        oS[bx] = sgS * 16 + (((k3 >> 1) ^ (sgS & 7)) << 1) + (k3 & 1);
        oN[bx] = C::BAND / 2 + sgN * 16 + (((k3 >> 1) ^ (sgN & 7)) << 1) + (k3 & 1);
    }
    const double cs1 = tv.cosgr[jh], cs2 = tv.cosgr2[jh];              // symmetric about the equator (checked by build_quad_tables)
    const double sgn = par ? -1.0 : 1.0;                                // even fold: north + south, odd fold: north - south
    const int t128 = tid & 127, setB = t128 >> 4, s16B = t128 & 15;
    const int slotB = 24 * (s16B >> 3) + 8 * b + (s16B & 7);
    double* ring = sRing + b * C::RING * C::BAND;
    uint64_t* gbar = bars + b * C::NBAR;
    pdl_wait();                                         // the grid fields of the previous kernel are complete
    pdl_trigger();
    const int gate_open = gate ? *gate : 1;
    auto skip = [&](int t) { return !gate_open && (sFl[t] & 4); };
    auto next_live = [&](int t) { while (t < cnt && skip(t)) t++; return t; };
    auto issue = [&](int t, int sq) {                   // one thread of the group: the two boxes of list entry t's band = live field number sq
        const int row0 = sRow0[t], e = sE[t];
        double* dst = ring + (sq % C::RING) * C::BAND;
        uint64_t* br = &gbar[sq % C::NBAR];
        fence_proxy_async();
        mbar_expect_tx(br, C::BAND * sizeof(double));
        tensor_g2s_4d(dst, &gmap, 0, 0, row0 + 8 * b, e, br);
        tensor_g2s_4d(dst + C::BAND / 2, &gmap, 0, 0, row0 + C::IL - 8 - 8 * b, e, br);
    };
    if (ph == 0 && wl == 0 && lane == 0) {              // the first RING fields of the band
        int tl = next_live(0);
        for (int k = 0; k < C::RING && tl < cnt; k++) { issue(tl, k); tl = next_live(tl + 1); }
    }
    int n = 0;                                          // live fields met so far (the same count in every thread)
    // phase stamps of CTA 0 (speedy_trace + SPEEDY_TRACE_STAMPS): cycles in [FFT of the quad, wait, tile sums, wait, staging, wait + copy-out, wait]
    long long tq = 0;
#define QSTAMP(i) do { if (tv.trace && tid == 0 && blockIdx.x == 0) { const long long t_ = clock64(); tv.trace[24 + (i)] += (unsigned long long)(t_ - tq); tq = t_; } } while (0)
    if (tv.trace) tq = clock64();
    CSTAMP(1, 1);
    for (int t0 = 0; t0 < cnt; t0 += 4) {
        unsigned present = 0;
        for (int fs = 0; fs < 4 && t0 + fs < cnt; fs++) {
            const int t = t0 + fs;
            if (skip(t)) continue;
            present |= 1u << fs;
            const int seq = n++;
            if (seq % C::NPH != ph) continue;           // the other group of this band takes it
            const int k = seq % C::RING, fl = sFl[t];
            double* G = ring + k * C::BAND;
            long long tf = 0;
#define FSTAMP(i) do { if (tv.trace && tid == 0 && blockIdx.x == 0) { const long long t_ = clock64(); tv.trace[1344 + (i)] += (unsigned long long)(t_ - tf); tf = t_; } } while (0)
            if (tv.trace) tf = clock64();
            mbar_wait(&gbar[seq % C::NBAR], (seq / C::NBAR) & 1);
            if (dead_in && t128 < 96) {
                // the band is in shared memory and nobody reads its source again (main-loop step): drop the 2 x 8 rows x 6 lines from L2
                // instead of letting them be written back to HBM on eviction
                const int hem = t128 / 48, r = (t128 - 48 * hem) / 6, l = t128 % 6;
                const int row = sRow0[t] + (hem ? C::IL - 8 - 8 * b : 8 * b) + r;
                l2_discard_line(reinterpret_cast<const char*>(dead_in + (size_t)sE[t] * in_ms + (size_t)row * C::IX) + l * 128);
            }
            FSTAMP(0);
            // ---- fold of the two hemispheres (legendre.f90:127-133, taken before the linear FFT; its Gaussian weight is in the P
            // fragments) with the cosgr / cosgr2 pre-scale of vdspec (spectral.f90:208-222)
            double x[12];
#pragma unroll
            for (int jj = 0; jj < 3; jj++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int bx = 2 * jj + (j >> 1), fl8 = 8 * (j & 1);       // second half of the box: chunk + 4
                    x[4 * jj + j] = fma(sgn, G[oS[bx] ^ fl8], G[oN[bx] ^ fl8]);
                }
            if (fl & 3) {
                const double sc = (fl & 1) ? cs1 : cs2;
#pragma unroll
                for (int u = 0; u < 12; u++) x[u] *= sc;
            }
            FSTAMP(1);
            named_sync(1 + gi, C::GT);                  // every row of the band is in registers: the buffer may be overwritten
            FSTAMP(2);
            // ---- stage A: radf3 + radf4 (fftpack.f90:774,844), in place
            Fft96F::stageA<16>(x, G + s16A, sWa, k3);
            FSTAMP(3);
            named_sync(1 + gi, C::GT);                  // T complete
            FSTAMP(4);
            // ---- stage B: radf4 + radf2 (fftpack.f90:844,722); half-complex position -> coefficient row, truncated at m = trunc
            if (t0 > 0) mbar_wait(eo_free, ((t0 >> 2) - 1) & 1);        // the previous quad's coefficients have left the buffer (bulk store)
            if (setB < 7) {
                double* E = sEO + fs * C::FS + slotB;
                auto st = [E](int pos, double v) { if (pos <= 2 * C::TRUNC) E[(pos + (pos > 0)) * C::CS] = v; };
                auto ld = [G, s16B](int blk, int off) { return G[(12 * blk + off) * 16 + (s16B ^ (4 * (blk & 3)))]; };
                // the five general sets share one instruction stream (the butterfly index is a run-time value): the two half-warps of
                // a warp work on different sets, and specialised copies would run one after the other
                if (setB < 5) Fft96F::stageB_general_g(ld, st, sWa, 3 + 2 * setB);
                else if (setB == 5) { Fft96F::stageB_first_g(ld, st, sWa); E[C::CS] = 0.0; }      // row c = 1: Im(m = 0) = 0 (fourier.f90:76)
                else Fft96F::stageB_last_g(ld, st, sWa);
            }
            FSTAMP(5);
            // the buffer goes back to the TMA (for the field RING places ahead) once the whole group has read T: only the issuing warp
            // waits (bar.sync / bar.arrive are warp-aligned instructions: the roles must be whole warps)
            if (wl == 0) {
                named_sync(7 + gi, C::GT);
                if (lane == 0) {
                    int t3 = t;
                    for (int u = 0; u < C::RING; u++) t3 = next_live(t3 + 1);
                    if (t3 < cnt) issue(t3, seq + C::RING);
                }
            } else {
                named_arrive(7 + gi, C::GT);
            }
            FSTAMP(6);
#undef FSTAMP
        }
        QSTAMP(0);
        __syncthreads();                                // EO of the quad complete
        QSTAMP(1);
        // ---- direct Legendre for the quad (legendre.f90:142-154): DMMA tiles, rows = 8 n of one parity, columns = (re, im) x 4 fields
        const double* Bq = sEO + (g >> 1) * C::FS + (g & 1) * C::CS + q;
        int tt[C::SLOTS];
        double c0[C::SLOTS], c1[C::SLOTS];
        {
            double a[C::SLOTS * C::KS];
#pragma unroll
            for (int c = 0; c < C::TCOLS / 16; c++) tmem_ld16(taddr + 16 * c, a + 8 * c);
            const double* Bp[C::SLOTS];
#pragma unroll
            for (int u = 0; u < C::SLOTS; u++) {
                const int L = u * C::WARPS + w;
                tt[u] = (L < C::NLIVE) ? sTile[L] : -1;
                const int tv_ = tt[u] < 0 ? 0 : tt[u];
                Bp[u] = Bq + 2 * (tv_ & 255) * C::CS + 24 * ((tv_ >> 8) & 1);
                c0[u] = 0.0; c1[u] = 0.0;
            }
#pragma unroll
            for (int ks = 0; ks < C::KS; ks++)
#pragma unroll
                for (int u = 0; u < C::SLOTS; u++) dmma884q(c0[u], c1[u], a[u * C::KS + ks], Bp[u][4 * ks]);
        }
        QSTAMP(2);
        __syncthreads();                                // every tile has read EO: the buffer becomes the output staging area
        QSTAMP(3);
        {
            // lane (g, q) holds (re, im) of coefficient (n, m) of field q: 16 bytes at [q][n][m]; row and field strides chosen so
            // that a quarter-warp's eight 16-byte stores fall in eight different bank groups
            double* stg = sEO + q * C::FSTG + 2 * g * C::RSTG;          // n = p + 2 (8 t + g): row n -> + (p + 16 t) rows
#pragma unroll
            for (int i = 0; i < C::SLOTS; i++)
                if (tt[i] >= 0)
                    *reinterpret_cast<double2*>(stg + (((tt[i] >> 8) & 1) + 16 * ((tt[i] >> 9) & 1)) * C::RSTG + 2 * (tt[i] & 255)) = make_double2(c0[i], c1[i]);
            for (int L = C::NLIVE + w; L < C::NTILE; L += C::WARPS) {   // n-tiles with no live coefficient: structural zeros
                const int td = sTile[L];
                *reinterpret_cast<double2*>(stg + (((td >> 8) & 1) + 16 * ((td >> 9) & 1)) * C::RSTG + 2 * (td & 255)) = make_double2(0.0, 0.0);
            }
        }
        QSTAMP(4);
        fence_proxy_async();                            // the staged coefficients (generic stores) become visible to the bulk-copy engine
        __syncthreads();
        // a field's 32 x 31 complex coefficients are one contiguous run: one bulk store each (TMA, SASS UBLKCP), asynchronous — the
        // next quad's FFT runs meanwhile; its stage B waits on `eo_free` before it writes the buffer again
        if (w == C::WARPS - 1 && lane == 0) {
            for (int fs = 0; fs < 4; fs++)
                if ((present >> fs) & 1) bulk_s2g(out_base + sOut[t0 + fs], sEO + fs * C::FSTG, C::NSPEC2 * sizeof(double));
            bulk_commit_wait_read();
            mbar_arrive(eo_free);
        }
        QSTAMP(5);
    }
#undef QSTAMP
    CSTAMP(1, 2);
    if (w == C::WARPS - 1 && lane == 0) bulk_wait_all();      // the last stores have left the SM's shared memory and are committed
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(*sTmem), "r"(512u) : "memory");
    if (tv.trace && tid == 0) trace_end(tv.trace, 2);
    CSTAMP(1, 3);
}

// ==========================================================================================================================
// k_s2g_quad  spec->grid = legendre_inv (legendre.f90:74-111) + fourier_inv (fourier.f90:23-53), incl. the uvspec / grad input
// stage (spectral.f90:124-196), the cosgr scale of fourier_inv (kcos /= 1) and the Coriolis add (tendencies.f90:103).  The mirror
// of k_g2s_quad: per quad of four fields
//   P0  the four spectral fields (or the source fields of derived PAIRS: ucos / vcos from (vor, div), d/dx / d/dy from ps)
//       arrive by bulk copies, one quad ahead;
//   P1  derived pairs are evaluated in registers and written back over their sources; coefficients outside the triangle
//       m + n <= trunc + 1 are zeroed (the reference never reads them, legendre.f90:38);
//   P2  inverse Legendre as DMMA tiles  X(j, (field, re/im)) = sum_n P(m, n, j) * S(n, (field, re/im))  per (m, latitude band),
//       even and odd n in two accumulators: row j (southern) = even - odd, row il+1-j (northern) = even + odd
//       (legendre.f90:105-109); the P fragments live in tensor memory; the half-complex rows (fourier.f90:33-45) go to
//       X[field][band][position][16 rows];
//   P3  six FFT groups (band x two fields at a time) run FFTPACK's backward FFT in place: stage 1 (radb2 + radb4) over X,
//       stage 2 (radb4 + radb3) into the band's grid rows in the swizzled layout of a tensor-map box, which leaves the SM
//       as two asynchronous tensor stores (southern and northern rows).
struct QICfg {
    static constexpr int TRUNC = 30, MX = 31, NX = 32, IX = 96, IL = 48, IY = 24, K2 = 2 * MX, NSPEC2 = NX * K2;
#ifndef SPD_K1Q_WARPS
#define SPD_K1Q_WARPS 12
#endif
    // 12 warps (three FFT groups, one per band, every field of the quad in turn; up to 168 registers per thread: the in-place stage 1
    // keeps a set's 16 inputs live across a barrier) or 24 warps (two fields at a time per band; 80 registers: spills)
    static constexpr int WARPS = SPD_K1Q_WARPS, THREADS = 32 * WARPS, NG = 3, GT = 128, NBOX = IX / 16, NPH = WARPS / 12;
    static constexpr int BAND = IX * 16;             // doubles per (field, band) buffer: X / T [96 positions][16 rows], then [2 hemispheres][8 rows][96]
    static constexpr int FSI = NSPEC2 + 2;           // spectral field stride in sIn: = 2 mod 16, conflict-free B fragments
    static constexpr int TSLOTS = 96 / WARPS, TCOLS = 16 * TSLOTS;   // tiles per warp; TMEM columns per warp (a tile = 8 fragments = 16 columns)
    static constexpr int NTILES = MX * NG;           // (m, band)
    static constexpr int LCAP = 128;                 // fields per CTA and launch
    static constexpr int NDEAD = (NX - 1) * (NX - 2) / 2;   // (m, n) of the mx x nx rectangle outside the triangle m + n <= mx (legendre.f90:38): n - 1 in row n
    static constexpr size_t SMEM = sizeof(double) * (4 * NG * BAND + 4 * FSI + IX) + sizeof(uint64_t) * 4 + sizeof(int) * WARPS * TSLOTS +
                                   LCAP * (2 * sizeof(long long) + 3 * sizeof(int)) + sizeof(unsigned short) * (NDEAD + 3);
    static_assert(FSI % 16 == 2 && (BAND * 8) % 1024 == 0 && SMEM <= 232448 && (WARPS / 4) * TCOLS <= 512, "layout");
};

// host: tiles (m, band) balanced over the warps by their k-step counts, P fragments per (warp, tile slot, fragment, lane):
// fragments 0..3 = even-n k-steps, 4..7 = odd-n k-steps of the tile; A(row = latitude pair of the band, k = n)
void build_quad_inverse_tables(const Tables& t, std::vector<int>& tiles, std::vector<double>& polyi) {
    using C = QICfg;
    struct Tl { int cost, m, b, ke, ko; };
    std::vector<Tl> all;
    for (int m = 0; m < C::MX; m++) {
        const int cnt = C::NX - m, ne = (cnt + 1) / 2, no = cnt / 2;     // n = 0..trunc+1-m (legendre.f90:38)
        for (int b = 0; b < C::NG; b++) all.push_back(Tl{(ne + 3) / 4 + (no + 3) / 4, m, b, (ne + 3) / 4, (no + 3) / 4});
    }
    std::stable_sort(all.begin(), all.end(), [](const Tl& x, const Tl& y) { return x.cost > y.cost; });
    std::vector<int> load(C::WARPS, 0), cnt(C::WARPS, 0);
    tiles.assign((size_t)C::WARPS * C::TSLOTS, -1);
    polyi.assign((size_t)C::WARPS * C::TSLOTS * 8 * 32, 0.0);
    for (const Tl& tl : all) {
        int w = 0;
        for (int i = 1; i < C::WARPS; i++) if (load[i] < load[w] || (load[i] == load[w] && cnt[i] < cnt[w])) w = i;
        if (cnt[w] >= C::TSLOTS) throw std::runtime_error("quad inverse tile balance");
        const int slot = cnt[w]++;
        load[w] += tl.cost;
        tiles[(size_t)w * C::TSLOTS + slot] = tl.m | (tl.b << 8) | (tl.ke << 12) | (tl.ko << 16);
        for (int fr = 0; fr < 8; fr++)
            for (int lane = 0; lane < 32; lane++) {
                const int g = lane >> 2, q = lane & 3, p = fr >> 2, ks = fr & 3;
                const int n = p + 2 * (4 * ks + q), jh = 8 * tl.b + g;
                const bool valid = n < C::NX && tl.m + n <= C::MX;
                polyi[((((size_t)w * C::TSLOTS + slot) * 4 + fr / 2) * 32 + lane) * 2 + (fr & 1)] = valid ? t.poly[((size_t)jh * C::NX + n) * C::MX + tl.m] : 0.0;   // pairs per lane: 16-byte loads
            }
    }
}

__global__ void __launch_bounds__(QICfg::THREADS, 1)
k_s2g_quad(const double* __restrict__ in_base, long long in_ms, const XDesc* __restrict__ desc, int nbatch, int q_base, int q_end, int nwork,
           const __grid_constant__ CUtensorMap omap, DevTables tv, CloseArgs cl) {
    using C = QICfg;
    if (blockIdx.x == (unsigned)nwork) {                // the extra CTA: closes the previous step, off the critical path
        pdl_wait();
        pdl_trigger();
        close_step_cta(cl, threadIdx.x);
        if (cl.ready) {                                 // the calendar is advanced: one count per member (member_ready.cuh)
            __syncthreads();
            for (int e = threadIdx.x; e < cl.ne; e += blockDim.x) ready_signal(cl.ready + e);
        }
        return;
    }
    extern __shared__ __align__(1024) double smem[];
    double* sX = smem;                                  // [4 fields][NG bands][BAND]
    double* sIn = sX + 4 * C::NG * C::BAND;             // [4][FSI] spectral coefficients of the quad, reference layout (m fastest)
    double* sWa = sIn + 4 * C::FSI;                     // [IX] twiddles
    uint64_t* bars = reinterpret_cast<uint64_t*>(sWa + C::IX);     // [0] quad sources landed; [3] holds the TMEM base address
    long long* sOff = reinterpret_cast<long long*>(bars + 4);      // [LCAP] source offset (first source of a derived pair: `off`)
    long long* sOff2 = sOff + C::LCAP;                              // [LCAP] second source (`off2` of a uvspec pair)
    int* sOp = reinterpret_cast<int*>(sOff2 + C::LCAP);             // [LCAP] op | flags << 8 | member << 16
    int* sOrow = sOp + C::LCAP;                                      // [LCAP] first row of the output field in the output tensor map
    int* sTile = sOrow + C::LCAP;                                    // [WARPS][TSLOTS]
    unsigned short* sDead = reinterpret_cast<unsigned short*>(sTile + C::WARPS * C::TSLOTS);   // [NDEAD] (n << 8 | m) outside the triangle
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(bars + 3);
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    if (tid == 0) trace_begin(tv.trace, 0);
    CSTAMP(0, 0);
    // this CTA's quads: quad Q = (member e, quad qm of the member's field list)
    const int qpm = (nbatch + 3) >> 2;
    const int nq = q_end - q_base;
    // quads are dealt out round robin: local quad cq = quad q_base + blockIdx.x + cq nwork, so the members complete in order
    const int ncq = (int)blockIdx.x < nq ? (nq - 1 - (int)blockIdx.x) / nwork + 1 : 0;
    auto quad_of = [&](int cq) { return q_base + (int)blockIdx.x + cq * nwork; };
    if (tid == 0 && ((smem_u32(sX) & 1023u) || 4 * ncq > C::LCAP)) __trap();
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_fence_init(); }
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(sTmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- prologue on constant tables only (may overlap the tail of the previous kernel: PDL)
    for (int t = tid; t < C::IX; t += C::THREADS) sWa[t] = tv.fftwa[t];
    for (int t = tid; t < C::WARPS * C::TSLOTS; t += C::THREADS) sTile[t] = tv.qtile_inv[t];
    for (int t = tid; t < C::MX * C::NX; t += C::THREADS) {
        const int n = t / C::MX, m = t - n * C::MX;
        if (m + n > C::MX) sDead[(n - 1) * (n - 2) / 2 + (m - (C::MX - n + 1))] = (unsigned short)((n << 8) | m);
    }
    for (int t = tid; t < 4 * ncq; t += C::THREADS) {
        const int Q = quad_of(t >> 2), e = Q / qpm, f = 4 * (Q - e * qpm) + (t & 3);
        if (f < nbatch) {
            const XDesc d = desc[f];
            sOff[t] = (long long)e * in_ms + d.off; sOff2[t] = (long long)e * in_ms + d.off2;
            sOp[t] = d.op | (d.flags << 8) | (e << 16);
            sOrow[t] = (d.oslot1 ? d.oslot1 - 1 : f) * C::IL;
        } else {
            sOp[t] = -1;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = *sTmem + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)(C::TCOLS * (w >> 2));
    // P fragments of this warp's tiles: parked in tensor memory for the whole kernel.  (Reading them from L2 inside the tile loop when a
    // CTA has only one or two quads was measured: the 196 KB per CTA take as long there as here — 8-member step 95.3 against 94.0 us.)
    // Only the k-steps a tile uses are fetched (high wavenumbers have few n: 43 % of the fragments are padding, and the fetch is
    // what the prologue costs — 148 CTAs pull the table out of L2 at once).
    const double2* pfrag = reinterpret_cast<const double2*>(tv.polyi) + (size_t)w * C::TSLOTS * 4 * 32 + lane;
    {
        double a[8];
#pragma unroll
        for (int i = 0; i < C::TSLOTS; i++) {
            const int tw = sTile[w * C::TSLOTS + i], ke = tw < 0 ? 0 : (tw >> 12) & 15, ko = tw < 0 ? 0 : (tw >> 16) & 15;
#pragma unroll
            for (int fr = 0; fr < 8; fr += 2) {
                const bool used = (fr < 4 ? fr : fr - 4) < (fr < 4 ? ke : ko);      // fragments fr, fr + 1 = k-steps of the even (0..3) / odd (4..7) n
                const double2 v = used ? pfrag[(i * 4 + fr / 2) * 32] : make_double2(0.0, 0.0);
                a[fr] = v.x; a[fr + 1] = v.y;
            }
            tmem_st16(taddr + 16 * i, a);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    // FFT roles: group = (band, field phase); stage 1 thread = (set, row); stage 2 thread = (row, k of radb4's third pass)
    const int gi = w >> 2, b = gi % C::NG, ph = gi / C::NG, wl = w & 3, t128 = tid & 127;
    static_assert(C::NPH == 1 || C::NPH == 2, "one or two fields at a time per band");
    const int set1 = t128 >> 4, row1 = t128 & 15;
    const int k3 = (lane & 3) + 4 * (lane >> 4), rsel = (lane >> 2) & 3;
    const int row2 = 4 * wl + rsel, hem = row2 >> 3, r8 = row2 & 7;          // a warp of stage 2 = 4 consecutive rows of one hemisphere block
    const int lat2 = hem ? (C::IL - 8 - 8 * b + r8) : (8 * b + r8);
    const double cg = tv.cosgr[lat2], cf = tv.coriol[lat2];
    int oG[C::NBOX];                                     // swizzled position of (row2, longitude 16 bx + k3) in the output box (as the input boxes of k_g2s_quad)
#pragma unroll
    for (int bx = 0; bx < C::NBOX; bx++) {
        const int sg = 6 * r8 + bx;
        oG[bx] = hem * (C::BAND / 2) + sg * 16 + (((k3 >> 1) ^ (sg & 7)) << 1) + (k3 & 1);
    }
    pdl_wait();                                         // the spectral fields of the previous kernel are complete
    pdl_trigger();
    auto issue_quad = [&](int cq) {                     // one thread: the sources of local quad cq
        uint32_t bytes = 0;
        for (int fs = 0; fs < 4; fs++) {
            const int t = 4 * cq + fs, op = sOp[t];
            if (op < 0) continue;
            const int o = op & 255;
            if (o == 0 || o == 1 || o == 3) bytes += C::NSPEC2 * sizeof(double);         // the field itself / first source of a pair
            else if (o == 2) bytes += C::NSPEC2 * sizeof(double);                        // second source of a uvspec pair
        }
        mbar_expect_tx(&bars[0], bytes);
        for (int fs = 0; fs < 4; fs++) {
            const int t = 4 * cq + fs, op = sOp[t];
            if (op < 0) continue;
            const int o = op & 255;
            if (o == 0 || o == 1 || o == 3) bulk_g2s(sIn + fs * C::FSI, in_base + sOff[t], C::NSPEC2 * sizeof(double), &bars[0]);
            else if (o == 2) bulk_g2s(sIn + fs * C::FSI, in_base + sOff2[t], C::NSPEC2 * sizeof(double), &bars[0]);
        }
    };
    if (tid == 0 && ncq > 0) issue_quad(0);
    // phase stamps of CTA 0 (speedy_trace + SPEEDY_TRACE_STAMPS): cycles in [sources + derived pairs, mask, wait, Legendre tiles, wait, FFT + stores, wait]
    long long tq = 0;
#define ISTAMP(i) do { if (tv.trace && tid == 0 && blockIdx.x == 0) { const long long t_ = clock64(); tv.trace[48 + (i)] += (unsigned long long)(t_ - tq); tq = t_; } } while (0)
    if (tv.trace) tq = clock64();
    CSTAMP(0, 1);
    for (int cq = 0; cq < ncq; cq++) {
        const int t0 = 4 * cq;
        // ---- P1: derived pairs in registers, then back over their sources; triangle mask
        bool waited = false;
#pragma unroll 1
        for (int pp = 0; pp < 2; pp++) {
            const int o = sOp[t0 + 2 * pp] < 0 ? 0 : (sOp[t0 + 2 * pp] & 255);
            if (o != 1 && o != 3) continue;             // (ucos, vcos) = uvspec(vor, div) / (d/dx, d/dy) = grad(ps) in slots 2 pp, 2 pp + 1
            constexpr int NU = (C::MX * C::NX + C::THREADS - 1) / C::THREADS;
            double pre[NU][3];                           // operator-table entries of this thread's two (m, n), fetched ahead of the wait
#pragma unroll
            for (int u = 0; u < NU; u++) {
                const int t = tid + u * C::THREADS;
                pre[u][0] = pre[u][1] = pre[u][2] = 0.0;
                if (t < C::MX * C::NX) {
                    const int m = t % C::MX;
                    pre[u][0] = (o == 1) ? tv.uvdx[t] : tv.gradx[m];
                    pre[u][1] = (o == 1) ? tv.uvdym[t] : tv.gradym[t];
                    pre[u][2] = (o == 1) ? tv.uvdyp[t] : tv.gradyp[t];
                }
            }
            if (!waited) { mbar_wait(&bars[0], cq & 1); waited = true; }
            double* A = sIn + (2 * pp) * C::FSI;
            double* B = sIn + (2 * pp + 1) * C::FSI;
            cd res[NU][2];
#pragma unroll
            for (int u = 0; u < NU; u++) {
                const int t = tid + u * C::THREADS;
                if (t < C::MX * C::NX) {
                    const int n = t / C::MX, m = t - n * C::MX;
                    if (o == 1) dev_uvspec_t(C::MX, C::NX, C::TRUNC, A, B, m, n, pre[u][0], pre[u][1], pre[u][2], res[u][0], res[u][1]);
                    else dev_grad_t(C::MX, C::NX, C::TRUNC, A, m, n, pre[u][0], pre[u][1], pre[u][2], res[u][0], res[u][1]);
                }
            }
            __syncthreads();                            // every stencil read is done: the sources may be overwritten
#pragma unroll
            for (int u = 0; u < NU; u++) {
                const int t = tid + u * C::THREADS;
                if (t < C::MX * C::NX) {
                    const int n = t / C::MX, m = t - n * C::MX;
                    const bool in = m + n <= C::MX;                  // outside the triangle: zero (legendre.f90:38)
                    st(A, C::MX, m, n, in ? res[u][0] : cd{0.0, 0.0});
                    st(B, C::MX, m, n, in ? res[u][1] : cd{0.0, 0.0});
                }
            }
        }
        if (!waited) mbar_wait(&bars[0], cq & 1);
        ISTAMP(0);
        for (int t = tid; t < 4 * C::NDEAD; t += C::THREADS) {           // plain fields: zero outside the triangle (a finite product with the zero P entries)
            const int fs = t / C::NDEAD, mn = sDead[t - fs * C::NDEAD];
            if ((sOp[t0 + fs] & 255) == 0) st(sIn + fs * C::FSI, C::MX, mn & 255, mn >> 8, cd{0.0, 0.0});
        }
        ISTAMP(1);
        __syncthreads();                                // sIn complete; X free (the previous quad's stores have read it: wait below)
        ISTAMP(2);
        // ---- P2: inverse Legendre, DMMA tiles (m, band), even and odd n
        {
            const double* Bq = sIn + (g >> 1) * C::FSI + (g & 1) + (2 * q) * C::K2;       // column = (field, re/im), k = n = parity + 2 (4 ks + q)
            double* Xq = sX + q * (C::NG * C::BAND);
            const int swx = 8 * (q & 1);
#pragma unroll
            for (int i = 0; i < C::TSLOTS; i += 2) {
                double a[2][8];
                tmem_ld16(taddr + 16 * i, a[0]);
                tmem_ld16(taddr + 16 * (i + 1), a[1]);
                int tl[2];
                double ce0[2], ce1[2], co0[2], co1[2];
#pragma unroll
                for (int u = 0; u < 2; u++) { tl[u] = sTile[w * C::TSLOTS + i + u]; ce0[u] = ce1[u] = co0[u] = co1[u] = 0.0; }
                // one warp-uniform trip count for the pair of tiles (their k-step counts differ by at most one: the tiles are dealt in
                // order of cost) and no branch around a DMMA: fragments beyond a tile's own count are zero
                const int tva = tl[0] < 0 ? 0 : tl[0], tvb = tl[1] < 0 ? 0 : tl[1];
                const int kmax = max(max((tva >> 12) & 15, (tva >> 16) & 15), max((tvb >> 12) & 15, (tvb >> 16) & 15));
                const double* Bpa = Bq + 2 * (tva & 255);
                const double* Bpb = Bq + 2 * (tvb & 255);
#pragma unroll
                for (int ks = 0; ks < 4; ks++) {
                    if (ks < kmax) {
                        dmma884q(ce0[0], ce1[0], a[0][ks], Bpa[ks * (8 * C::K2)]);
                        dmma884q(co0[0], co1[0], a[0][4 + ks], Bpa[ks * (8 * C::K2) + C::K2]);
                        dmma884q(ce0[1], ce1[1], a[1][ks], Bpb[ks * (8 * C::K2)]);
                        dmma884q(co0[1], co1[1], a[1][4 + ks], Bpb[ks * (8 * C::K2) + C::K2]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    if (tl[u] < 0) continue;
                    const int m = tl[u] & 255, bb = (tl[u] >> 8) & 15;
                    double* Xb = Xq + bb * C::BAND;
                    const int rs = g ^ swx, rn = (15 - g) ^ swx;           // southern row jl, northern row in ascending-latitude order
                    const int pr = m ? 2 * m - 1 : 0;                      // FFTPACK's half-complex order (fourier.f90:40-45); Im(m = 0) is dropped
                    Xb[pr * 16 + rs] = ce0[u] - co0[u];
                    Xb[pr * 16 + rn] = ce0[u] + co0[u];
                    if (m) {
                        Xb[(pr + 1) * 16 + rs] = ce1[u] - co1[u];
                        Xb[(pr + 1) * 16 + rn] = ce1[u] + co1[u];
                    }
                }
            }
        }
        ISTAMP(3);
        __syncthreads();                                // X of the quad complete; sIn free
        ISTAMP(4);
        if (tid == 0 && cq + 1 < ncq) issue_quad(cq + 1);
        // the previous quad's stores of this band group were issued a Legendre phase ago: complete by now, so its member can be told
        if (cl.ready && cq > 0 && wl == 0 && lane == 0) { bulk_wait_all(); ready_signal(cl.ready + quad_of(cq - 1) / qpm); }
        // ---- P3: backward FFT per (field, band), fields ph and ph + 2 of the quad
#pragma unroll 1
        for (int fs = ph; fs < 4; fs += C::NPH) {
            const int op = sOp[t0 + fs];
            if (op < 0) continue;
            const int fl = (op >> 8) & 255;
            double* Xb = sX + (fs * C::NG + b) * C::BAND;
            // stage 1: radb2 + radb4 (fftpack.f90:204,328), in place: every input of the set into registers, group barrier (at a point
            // where the warps converge: bar.sync is warp-aligned and the two half-warps of a warp work on different sets), then the stores
            {
                const int rx = row1 ^ (8 * (fs & 1));
                auto ld = [Xb, rx](int pos) { return pos <= 2 * C::TRUNC ? Xb[pos * 16 + rx] : 0.0; };       // zero padding above the truncation (fourier.f90:34-41)
                auto stt = [Xb, row1](int blk, int off, double v) { Xb[(12 * blk + off) * 16 + (row1 ^ (4 * (blk & 3)))] = v; };
                double v[16];
                if (set1 < 5) Fft96::stage1_general_ld(ld, 3 + 2 * set1, v);
                else if (set1 == 5) Fft96::stage1_first_ld(ld, v);
                else if (set1 == 6) Fft96::stage1_last_ld(ld, v);
                named_sync(1 + gi, C::GT);
                if (set1 < 5) Fft96::stage1_general_st(v, stt, sWa, 3 + 2 * set1);
                else if (set1 == 5) Fft96::stage1_first_st(v, stt, sWa);
                else if (set1 == 6) Fft96::stage1_last_st(v, stt, sWa);
            }
            named_sync(1 + gi, C::GT);                  // T complete
            // stage 2: radb4 + radb3 (fftpack.f90:328,256), then fourier_inv's cosgr scale and the Coriolis add
            double y[12];
            Fft96::stage2<16>(Xb + (row2 ^ (4 * (k3 & 3))), sWa, k3, y);
            named_sync(1 + gi, C::GT);                  // every T value is in registers: the buffer becomes the grid rows of the band
#pragma unroll
            for (int jj = 0; jj < 3; jj++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    double v = y[4 * jj + j];
                    if (fl & 1) v *= cg;
                    if (fl & 2) v += cf;
                    Xb[oG[2 * jj + (j >> 1)] ^ (8 * (j & 1))] = v;
                }
            fence_proxy_async();
            if (wl == 0) {
                named_sync(7 + gi, C::GT);
                if (lane == 0) {
                    const int e = op >> 16, row0 = sOrow[t0 + fs];
                    tensor_s2g_4d(&omap, 0, 0, row0 + 8 * b, e, Xb);
                    tensor_s2g_4d(&omap, 0, 0, row0 + C::IL - 8 - 8 * b, e, Xb + C::BAND / 2);
                    bulk_commit();
                }
            } else {
                named_arrive(7 + gi, C::GT);
            }
        }
        if (wl == 0 && lane == 0) bulk_wait_read_all(); // this group's stores have read their buffers
        ISTAMP(5);
        __syncthreads();
        ISTAMP(6);
    }
#undef ISTAMP
    CSTAMP(0, 2);
    if (wl == 0 && lane == 0) {
        bulk_wait_all();
        if (cl.ready && ncq > 0) ready_signal(cl.ready + quad_of(ncq - 1) / qpm);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(*sTmem), "r"(512u) : "memory");
    if (tv.trace && tid == 0) trace_end(tv.trace, 0);
    CSTAMP(0, 3);
}

// completion counts one member's fields add to its ready counter (member_ready.cuh): one per band group and quad
unsigned s2g_quad_ready_counts(int nbatch) { return (unsigned)(QICfg::NG * QICfg::NPH * ((nbatch + 3) / 4)); }

void setup_quad_kernels() {
    CUDA_CHECK(cudaFuncSetAttribute(k_g2s_quad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QCfg::SMEM_G2S));
    CUDA_CHECK(cudaFuncSetAttribute(k_s2g_quad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QICfg::SMEM));
}

// 4-D view of a batch of grid fields for the band loads: [16 longitudes][IX/16 blocks][row of IX doubles][member], box = one
// hemisphere block of a band ([16][6][8 rows][1]), SWIZZLE_128B.  A field at element offset `off` starts at row off / IX.
typedef CUresult (*EncodeTiledFnQ)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static const CUtensorMap& grid_band_map(const double* d_in, long long in_ms, int nmembers) {
    using C = QCfg;
    static std::map<std::tuple<const void*, long long, int>, CUtensorMap> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    const auto key = std::make_tuple((const void*)d_in, in_ms, nmembers);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    static EncodeTiledFnQ enc = nullptr;
    if (!enc) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
        if (!fn || qr != cudaDriverEntryPointSuccess) throw std::runtime_error("cuTensorMapEncodeTiled is not available in this driver");
        enc = reinterpret_cast<EncodeTiledFnQ>(fn);
    }
    if ((reinterpret_cast<uintptr_t>(d_in) & 15) || (nmembers > 1 && (in_ms * sizeof(double)) % 16)) throw std::runtime_error("grid_to_spec: input buffer must be 16-byte aligned");
    const cuuint64_t rows = (nmembers > 1 && in_ms > 0) ? (cuuint64_t)(in_ms / C::IX) : (cuuint64_t)1 << 22;
    const cuuint64_t dims[4] = {16, (cuuint64_t)C::NBOX, rows, (cuuint64_t)nmembers};
    const cuuint64_t strides[3] = {16 * sizeof(double), (cuuint64_t)C::IX * sizeof(double), (nmembers > 1 ? (cuuint64_t)in_ms : rows * (cuuint64_t)C::IX) * sizeof(double)};
    const cuuint32_t box[4] = {16u, (cuuint32_t)C::NBOX, 8u, 1u};
    const cuuint32_t es[4] = {1u, 1u, 1u, 1u};
    CUtensorMap m;
    const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(d_in), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed for the band map (" + std::to_string((int)r) + ")");
    return cache.emplace(key, m).first->second;
}

void launch_g2s_quad(speedy_ctx* ctx, const double* d_in, long long in_ms, const XDesc* d_desc, int nbatch, double* d_out, long long out_ms, int nmembers, const int* gate,
                     const G2sStepOpts& step) {
    using C = QCfg;
    const CUtensorMap& gmap = grid_band_map(d_in, in_ms, nmembers);
    const int nf = nbatch * nmembers, per = C::LCAP * ctx->num_sms;      // fields per launch: at most LCAP per CTA
    for (int base = 0; base < nf; base += per) {
        const int end = base + per < nf ? base + per : nf;
        // fewer quads than SMs (8 members of the model step: 70): hand out pairs instead, the FFT phase of a CTA halves
        const int nquad = (end - base + 3) / 4, unit = nquad < ctx->num_sms ? 2 : 4;
        const int nunit = (end - base + unit - 1) / unit;
        const int ncta = nunit < ctx->num_sms ? nunit : ctx->num_sms;
        CUDA_CHECK(launch_pdl(ctx->dv.trace == nullptr || ctx->trace_pdl, k_g2s_quad, dim3(ncta), dim3(C::THREADS), C::SMEM_G2S, ctx->stream, gmap, d_desc, nbatch, base, end, unit,
                              d_out, out_ms, ctx->dv, gate, (step.transient_input && ctx->l2_discard) ? d_in : (const double*)nullptr, in_ms,
                              step.out_field_stride ? step.out_field_stride : (long long)C::NSPEC2));
    }
}

void launch_s2g_quad(speedy_ctx* ctx, const double* d_in, long long in_ms, const XDesc* d_desc, int nbatch, double* d_out, long long out_ms, int nmembers,
                     const CloseArgs& cl) {
    using C = QICfg;
    const CUtensorMap& omap = grid_band_map(d_out, out_ms, nmembers);
    const int qpm = (nbatch + 3) / 4, nq = qpm * nmembers;
    const int reserve = cl.clk ? 1 : 0;                                   // one SM is left to the closing CTA when a step is to be closed
    const int per = (C::LCAP / 4) * (ctx->num_sms - reserve);             // quads per launch: at most LCAP / 4 per CTA
    bool first = true;
    for (int base = 0; base < nq; base += per) {
        const int end = base + per < nq ? base + per : nq;
        const int nwork = (end - base) < ctx->num_sms - reserve ? (end - base) : ctx->num_sms - reserve;
        CloseArgs c = first ? cl : CloseArgs{nullptr, nullptr, 0, 0, nullptr, nullptr};
        c.ready = cl.ready;
        CUDA_CHECK(launch_pdl(ctx->dv.trace == nullptr || ctx->trace_pdl, k_s2g_quad, dim3(nwork + (c.clk ? 1 : 0)), dim3(C::THREADS), C::SMEM, ctx->stream, d_in, in_ms, d_desc,
                              nbatch, base, end, nwork, omap, ctx->dv, c));
        first = false;
    }
}

}  // namespace spd
