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
#include <cuda.h>
#include <map>
#include <mutex>
#include <tuple>
#include "fft96f.cuh"

namespace spd {

__device__ __forceinline__ void dmma884q(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct QCfg {
    static constexpr int TRUNC = 30, MX = 31, NX = 32, IX = 96, IL = 48, IY = 24, K2 = 2 * MX, NSPEC2 = NX * K2;
    static constexpr int WARPS = 12, THREADS = 32 * WARPS;
    static constexpr int NG = 3, GT = 128;           // latitude bands = warpgroups; threads per group
    static constexpr int RING = 3;                   // band buffers per group: one being transformed, two in flight
    static constexpr int NBOX = IX / 16;             // tensor-map boxes of [16 longitudes x 8 latitudes] = one 1 KB swizzle atom
    static constexpr int BAND = 2 * NBOX * 128;      // doubles per band buffer: [hemisphere][box][8 rows][16 longitudes]
    static constexpr int CS = 52, FS = K2 * CS;      // EO[field][c][slot]; CS = 4 mod 16 and FS = 8 mod 16: conflict-free B fragments
    static constexpr int SLOTS = 8, KS = IY / 4;     // tiles per warp, k-steps per tile
    static constexpr int NLIVE = 93, NTILE = 124;    // live / all (m, parity, n-tile) tiles
    static constexpr int LCAP = 128;                 // fields per CTA and launch (the host splits larger batches)
    static constexpr int RSTG = 68, FSTG = NX * RSTG + 2;   // output staging [field][n][RSTG]: row = 2 mod 4 and field = 1 mod 8 sixteen-byte slots
    static constexpr size_t SMEM_G2S = sizeof(double) * (NG * RING * BAND + 4 * FS + IX) + sizeof(uint64_t) * (NG * RING + 1) + sizeof(int) * NTILE +
                                       LCAP * (3 * sizeof(int) + sizeof(long long));
    static_assert(FS % 16 == 8 && CS % 16 == 4 && (BAND * 8) % 1024 == 0 && BAND == IX * 16 && SMEM_G2S <= 232448, "layout");
    static_assert(WARPS * SLOTS >= NLIVE && NG * 8 == IY && NG * GT == THREADS && SLOTS % 4 == 0, "tiling");
    static_assert((RSTG / 2) % 4 == 2 && (FSTG / 2) % 8 == 1 && RSTG >= K2 && 4 * FSTG <= 4 * FS, "staging layout");
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
                    polyq[(((size_t)w * C::SLOTS + i) * C::KS + ks) * 32 + lane] = valid ? t.poly[((size_t)jh * C::NX + n) * C::MX + m] * (t.wt[jh] * scale) : 0.0;
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
__global__ void __launch_bounds__(QCfg::THREADS, 1)
k_g2s_quad(const __grid_constant__ CUtensorMap gmap, const XDesc* __restrict__ desc, int nbatch, int idx_base, int idx_end,
           double* __restrict__ out_base, long long out_ms, DevTables tv, const int* __restrict__ gate) {
    using C = QCfg;
    extern __shared__ __align__(1024) double smem[];
    double* sRing = smem;                               // [NG][RING][BAND]
    double* sEO = sRing + C::NG * C::RING * C::BAND;    // [4][K2][CS]; after the tile sums: output staging [4][NX][RSTG]
    double* sWa = sEO + 4 * C::FS;                      // [IX] twiddles
    uint64_t* bars = reinterpret_cast<uint64_t*>(sWa + C::IX);     // [NG][RING] band buffers
    long long* sOut = reinterpret_cast<long long*>(bars + C::NG * C::RING + 1);   // [LCAP] output offset of the field
    int* sRow0 = reinterpret_cast<int*>(sOut + C::LCAP);           // [LCAP] first row of the field in the tensor map
    int* sE = sRow0 + C::LCAP;                                      // [LCAP] member
    int* sFl = sE + C::LCAP;                                        // [LCAP] descriptor flags
    int* sTile = sFl + C::LCAP;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    if (tid == 0) trace_begin(tv.trace, 2);
    // this CTA's quads of the flattened (member, field) list
    const int nquad = (idx_end - idx_base + 3) >> 2;
    const int q0 = (int)((long long)blockIdx.x * nquad / gridDim.x), q1 = (int)((long long)(blockIdx.x + 1) * nquad / gridDim.x);
    const int i0 = idx_base + 4 * q0, cnt = min(idx_base + 4 * q1, idx_end) - i0;
    if (tid == 0 && ((smem_u32(sRing) & 1023u) || cnt > C::LCAP)) __trap();
    if (tid == 0) {
        for (int k = 0; k < C::NG * C::RING; k++) mbar_init(&bars[k], 1);
        mbar_fence_init();
    }
    // ---- prologue on constant tables only (may overlap the tail of the previous kernel: PDL)
    double a[C::SLOTS][C::KS];                          // P fragments of this warp's tiles, resident for the whole kernel
#pragma unroll
    for (int i = 0; i < C::SLOTS; i++)
#pragma unroll
        for (int ks = 0; ks < C::KS; ks++) a[i][ks] = tv.polyq[(((size_t)w * C::SLOTS + i) * C::KS + ks) * 32 + lane];
    for (int t = tid; t < 2 * C::FS; t += C::THREADS) reinterpret_cast<double2*>(sEO)[t] = make_double2(0.0, 0.0);   // row c = 1 (Im m = 0, fourier.f90:76) stays zero
    for (int t = tid; t < C::IX; t += C::THREADS) sWa[t] = tv.fftwa[t];
    for (int t = tid; t < C::NTILE; t += C::THREADS) sTile[t] = tv.qtile[t];
    for (int t = tid; t < cnt; t += C::THREADS) {       // the descriptors are constants: the field list costs no round trip after the wait
        const int idx = i0 + t, e = idx / nbatch, f = idx - e * nbatch;
        const XDesc d = desc[f];
        const int row0 = (int)(d.off / C::IX);
        if (d.off != (long long)row0 * C::IX) __trap();
        sRow0[t] = row0; sE[t] = e; sFl[t] = d.flags; sOut[t] = (long long)e * out_ms + (long long)f * C::NSPEC2;
    }
    // roles: group = latitude band; stage A thread = (parity, latitude pair of the band, k of radf4); stage B thread = (set, slot)
    const int b = w >> 2, wl = w & 3, par = wl >> 1, ha = wl & 1;
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
    uint64_t* gbar = bars + b * C::RING;
    __syncthreads();
    pdl_wait();                                         // the grid fields of the previous kernel are complete
    pdl_trigger();
    const int gate_open = gate ? *gate : 1;
    auto skip = [&](int t) { return !gate_open && (sFl[t] & 4); };
    auto next_live = [&](int t) { while (t < cnt && skip(t)) t++; return t; };
    auto issue = [&](int t, int k) {                    // one thread of the group: the two boxes of field t's band into ring buffer k
        const int row0 = sRow0[t], e = sE[t];
        double* dst = ring + k * C::BAND;
        fence_proxy_async();
        mbar_expect_tx(&gbar[k], C::BAND * sizeof(double));
        tensor_g2s_4d(dst, &gmap, 0, 0, row0 + 8 * b, e, &gbar[k]);
        tensor_g2s_4d(dst + C::BAND / 2, &gmap, 0, 0, row0 + C::IL - 8 - 8 * b, e, &gbar[k]);
    };
    int tl = next_live(0);                              // next field to load
    for (int k = 0; k < C::RING; k++)
        if (tl < cnt) { if (wl == 0 && lane == 0) issue(tl, k); tl = next_live(tl + 1); }
    int n = 0;                                          // fields consumed by this group
    // phase stamps of CTA 0 (speedy_trace + SPEEDY_TRACE_STAMPS): cycles in [FFT of the quad, wait, tile sums, wait, staging, wait + copy-out, wait]
    long long tq = 0;
#define QSTAMP(i) do { if (tv.trace && tid == 0 && blockIdx.x == 0) { const long long t_ = clock64(); tv.trace[24 + (i)] += (unsigned long long)(t_ - tq); tq = t_; } } while (0)
    if (tv.trace) tq = clock64();
    for (int t0 = 0; t0 < cnt; t0 += 4) {
        unsigned present = 0;
        for (int fs = 0; fs < 4 && t0 + fs < cnt; fs++) {
            const int t = t0 + fs;
            if (skip(t)) continue;
            present |= 1u << fs;
            const int k = n % C::RING, fl = sFl[t];
            double* G = ring + k * C::BAND;
            mbar_wait(&gbar[k], (n / C::RING) & 1);
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
            named_sync(1 + b, C::GT);                   // every row of the band is in registers: the buffer may be overwritten
            // ---- stage A: radf3 + radf4 (fftpack.f90:774,844), in place
            Fft96F::stageA<16>(x, G + s16A, sWa, k3);
            named_sync(1 + b, C::GT);                   // T complete
            // ---- stage B: radf4 + radf2 (fftpack.f90:844,722); half-complex position -> coefficient row, truncated at m = trunc
            if (setB < 7) {
                double* E = sEO + fs * C::FS + slotB;
                auto st = [E](int pos, double v) { if (pos <= 2 * C::TRUNC) E[(pos + (pos > 0)) * C::CS] = v; };
                auto ld = [G, s16B](int blk, int off) { return G[(12 * blk + off) * 16 + (s16B ^ (4 * (blk & 3)))]; };
                switch (setB) {                           // one instantiation per set: every position, twiddle index and truncation test is an immediate
                    case 0: Fft96F::stageB_general_g(ld, st, sWa, 3); break;
                    case 1: Fft96F::stageB_general_g(ld, st, sWa, 5); break;
                    case 2: Fft96F::stageB_general_g(ld, st, sWa, 7); break;
                    case 3: Fft96F::stageB_general_g(ld, st, sWa, 9); break;
                    case 4: Fft96F::stageB_general_g(ld, st, sWa, 11); break;
                    case 5: Fft96F::stageB_first_g(ld, st, sWa); break;
                    default: Fft96F::stageB_last_g(ld, st, sWa); break;
                }
            }
            // the buffer goes back to the TMA once the whole group has read T: only the issuing warp waits (bar.sync / bar.arrive
            // are warp-aligned instructions: the roles must be whole warps)
            if (wl == 0) {
                named_sync(4 + b, C::GT);
                if (lane == 0 && tl < cnt) issue(tl, k);
            } else {
                named_arrive(4 + b, C::GT);
            }
            if (tl < cnt) tl = next_live(tl + 1);
            n++;
        }
        QSTAMP(0);
        __syncthreads();                                // EO of the quad complete
        QSTAMP(1);
        // ---- direct Legendre for the quad (legendre.f90:142-154): DMMA tiles, rows = 8 n of one parity, columns = (re, im) x 4 fields
        const double* Bq = sEO + (g >> 1) * C::FS + (g & 1) * C::CS + q;
        int tt[C::SLOTS];
        double c0[C::SLOTS], c1[C::SLOTS];
#pragma unroll
        for (int i = 0; i < C::SLOTS; i += 4) {
            const double* Bp[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int L = (i + u) * C::WARPS + w;
                tt[i + u] = (L < C::NLIVE) ? sTile[L] : -1;
                const int tv_ = tt[i + u] < 0 ? 0 : tt[i + u];
                Bp[u] = Bq + 2 * (tv_ & 255) * C::CS + 24 * ((tv_ >> 8) & 1);
                c0[i + u] = 0.0; c1[i + u] = 0.0;
            }
#pragma unroll
            for (int ks = 0; ks < C::KS; ks++)
#pragma unroll
                for (int u = 0; u < 4; u++) dmma884q(c0[i + u], c1[i + u], a[i + u][ks], Bp[u][4 * ks]);
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
        __syncthreads();
        // whole 128-byte lines to global memory: a field's 32 x 31 complex coefficients are contiguous
        for (int fs = 0; fs < 4; fs++) {
            if (!((present >> fs) & 1)) continue;
            double2* outp = reinterpret_cast<double2*>(out_base + sOut[t0 + fs]);
            const double2* src = reinterpret_cast<const double2*>(sEO + fs * C::FSTG);
            for (int s2 = tid; s2 < C::NX * C::MX; s2 += C::THREADS) {
                const int nn = s2 / C::MX, mm = s2 - nn * C::MX;
                outp[s2] = src[nn * (C::RSTG / 2) + mm];
            }
        }
        QSTAMP(5);
        __syncthreads();                                // staging read: the buffer is EO again
        QSTAMP(6);
        for (int t = tid; t < 4 * C::IL; t += C::THREADS) sEO[(t / C::IL) * C::FS + C::CS + (t % C::IL)] = 0.0;   // restore row c = 1
    }
#undef QSTAMP
    if (tv.trace) { __syncthreads(); if (tid == 0) trace_end(tv.trace, 2); }
}

void setup_quad_kernels() {
    CUDA_CHECK(cudaFuncSetAttribute(k_g2s_quad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QCfg::SMEM_G2S));
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

void launch_g2s_quad(speedy_ctx* ctx, const double* d_in, long long in_ms, const XDesc* d_desc, int nbatch, double* d_out, long long out_ms, int nmembers, const int* gate) {
    using C = QCfg;
    const CUtensorMap& gmap = grid_band_map(d_in, in_ms, nmembers);
    const int nf = nbatch * nmembers, per = C::LCAP * ctx->num_sms;      // fields per launch: at most LCAP per CTA
    for (int base = 0; base < nf; base += per) {
        const int end = base + per < nf ? base + per : nf;
        const int nquad = (end - base + 3) / 4;
        const int ncta = nquad < ctx->num_sms ? nquad : ctx->num_sms;
        CUDA_CHECK(launch_pdl(ctx->dv.trace == nullptr || ctx->trace_pdl, k_g2s_quad, dim3(ncta), dim3(C::THREADS), C::SMEM_G2S, ctx->stream, gmap, d_desc, nbatch, base, end,
                              d_out, out_ms, ctx->dv, gate));
    }
}

}  // namespace spd
