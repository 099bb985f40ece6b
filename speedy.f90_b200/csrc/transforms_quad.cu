// Batch ("quad") variants of the spherical-harmonic transforms for ensemble batches at T30: four fields at a time, so that
// the Legendre sums become FP64 tensor-core contractions (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4; tcgen05 has no FP64 kind)
// with the 8 columns of a tile = (re, im) x 4 fields, and the zonal transform is FFTPACK's own FFT (fft96f.cuh / fft96.cuh),
// not the dense operator.
//
//   k_g2s_quad  grid->spec = fourier_dir (fourier.f90:56-82) + legendre_dir (legendre.f90:114-155), incl. the cosgr / cosgr2
//               pre-scale of vdspec (spectral.f90:208-222).  One persistent CTA per SM owns WHOLE fields:
//     * the field arrives as six tensor-map boxes with the 128-byte swizzle (as k_g2s_stream), double-buffered;
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
#include "fft96f.cuh"

namespace spd {

__device__ __forceinline__ void dmma884q(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct QCfg {
    static constexpr int TRUNC = 30, MX = 31, NX = 32, IX = 96, IL = 48, IY = 24, K2 = 2 * MX, NSPEC2 = NX * K2;
    static constexpr int NBOX = IX / 16, BOX = 16 * IL;
    static constexpr int WARPS = 12, THREADS = 32 * WARPS;
    static constexpr int XS = 49;                    // sT[pos][slot]: odd stride
    static constexpr int CS = 52, FS = K2 * CS;      // EO[field][c][slot]; CS = 4 mod 16 and FS = 8 mod 16: conflict-free B fragments
    static constexpr int GBUF = IX * IL;             // doubles per grid-field buffer (36 KB = whole 1 KB swizzle atoms)
    static constexpr int SLOTS = 8, KS = IY / 4;     // tiles per warp, k-steps per tile
    static constexpr int NLIVE = 93, NTILE = 124;    // live / all (m, parity, n-tile) tiles
    static constexpr size_t SMEM_G2S = sizeof(double) * (2 * GBUF + IX * XS + 4 * FS + IX) + 4 * sizeof(uint64_t) + sizeof(int) * NTILE;
    static_assert(FS % 16 == 8 && CS % 16 == 4 && (GBUF * 8) % 1024 == 0 && SMEM_G2S <= 232448, "layout");
    static_assert(WARPS * SLOTS >= NLIVE && 8 * IL == THREADS, "tiling");
};

// slot <-> latitude pair inside a hemisphere block of 8: a warp of stage A works on the latitude pairs {a, a+2, a+4, a+6}
// (distinct swizzle phases of the grid rows: conflict-free reads) and stores them as 4 consecutive slots (conflict-free writes)
__host__ __device__ constexpr int q_slot_of(int jh) { return (jh & ~7) | ((jh & 6) >> 1) | ((jh & 1) << 2); }
__host__ __device__ constexpr int q_jh_of(int s) { return (s & ~7) | ((s & 3) << 1) | ((s & 4) >> 2); }

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
    polyq.assign((size_t)C::WARPS * C::SLOTS * C::KS * 32, 0.0);
    for (int w = 0; w < C::WARPS; w++)
        for (int i = 0; i < C::SLOTS; i++) {
            const int L = i * C::WARPS + w;
            if (L >= C::NLIVE) continue;
            const int m = tiles[L] & 255, p = (tiles[L] >> 8) & 1, tt = (tiles[L] >> 9) & 1;
            for (int ks = 0; ks < C::KS; ks++)
                for (int lane = 0; lane < 32; lane++) {
                    const int g = lane >> 2, q = lane & 3;
                    const int n = p + 2 * (8 * tt + g), jh = q_jh_of(4 * ks + q);
                    const bool valid = n <= C::TRUNC && m + n <= C::MX;
                    polyq[(((size_t)w * C::SLOTS + i) * C::KS + ks) * 32 + lane] = valid ? t.poly[((size_t)jh * C::NX + n) * C::MX + m] : 0.0;
                }
        }
}

__global__ void __launch_bounds__(QCfg::THREADS, 1)
k_g2s_quad(const __grid_constant__ CUtensorMap gmap, const XDesc* __restrict__ desc, int nbatch, int nmembers,
           double* __restrict__ out_base, long long out_ms, DevTables tv, const int* __restrict__ gate) {
    using C = QCfg;
    extern __shared__ __align__(1024) double smem[];
    double* sG = smem;                                  // [2][NBOX][IL][16] swizzled grid fields
    double* sT = sG + 2 * C::GBUF;                      // [IX][XS] between the FFT stages
    double* sEO = sT + C::IX * C::XS;                   // [4][K2][CS]
    double* sWa = sEO + 4 * C::FS;                      // [IX] twiddles
    uint64_t* bars = reinterpret_cast<uint64_t*>(sWa + C::IX);     // [0], [1]: grid buffers
    int* sTile = reinterpret_cast<int*>(bars + 4);
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    if (tid == 0) trace_begin(tv.trace, 2);
    // this CTA's quads of the flattened (member, field) list
    const int nf = nbatch * nmembers, nquad = (nf + 3) >> 2;
    const int q0 = (int)((long long)blockIdx.x * nquad / gridDim.x), q1 = (int)((long long)(blockIdx.x + 1) * nquad / gridDim.x);
    const int i0 = 4 * q0, i1 = min(4 * q1, nf);
    if (tid == 0 && (smem_u32(sG) & 1023u)) __trap();
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    // ---- prologue on constant tables only (may overlap the tail of the previous kernel: PDL)
    double a[C::SLOTS][C::KS];                          // P fragments of this warp's tiles, resident for the whole kernel
#pragma unroll
    for (int i = 0; i < C::SLOTS; i++)
#pragma unroll
        for (int ks = 0; ks < C::KS; ks++) a[i][ks] = tv.polyq[(((size_t)w * C::SLOTS + i) * C::KS + ks) * 32 + lane];
    for (int t = tid; t < 4 * C::FS; t += C::THREADS) sEO[t] = 0.0;      // row c = 1 (Im m = 0, fourier.f90:76) stays zero
    for (int t = tid; t < C::IX; t += C::THREADS) sWa[t] = tv.fftwa[t];
    for (int t = tid; t < C::NTILE; t += C::THREADS) sTile[t] = tv.qtile[t];
    // stage A role: folded row (parity, latitude pair jh) x k of radf4
    const int par = w / 6, w6 = w - 6 * par;
    const int k3 = (lane & 3) + 4 * (lane >> 4), rsel = (lane >> 2) & 3;
    const int jh = 8 * (w6 >> 1) + (w6 & 1) + 2 * rsel;
    const int slotA = 24 * par + q_slot_of(jh);
    const int lat_s = jh, lat_n = C::IL - 1 - jh;
    const double wsc = tv.wt[jh] * (double)(1.0f / (float)C::IX);       // legendre.f90:131-132 weight x fourier.f90:72 scale
    const double cs1 = tv.cosgr[lat_s], cn1 = tv.cosgr[lat_n], cs2 = tv.cosgr2[lat_s], cn2 = tv.cosgr2[lat_n];
    // stage B role: (butterfly set, folded-row slot), consecutive lanes = consecutive slots
    const int setB = tid / C::IL, slotB = tid - setB * C::IL;
    __syncthreads();
    pdl_wait();                                         // the grid fields of the previous kernel are complete
    pdl_trigger();
    const int gate_open = gate ? *gate : 1;
    auto active = [&](int idx) { return gate_open || !(desc[idx % nbatch].flags & 4); };
    auto next_active = [&](int idx) { while (idx < i1 && !active(idx)) idx++; return idx; };
    auto issue = [&](int idx, int buf) {               // one thread: the six boxes of field idx into buffer buf
        const int e = idx / nbatch, f = idx - e * nbatch;
        const long long off = desc[f].off;
        const int row0 = (int)(off / C::IX);
        if (off != (long long)row0 * C::IX) __trap();
        fence_proxy_async();
        mbar_expect_tx(&bars[buf], C::GBUF * sizeof(double));
#pragma unroll
        for (int b = 0; b < C::NBOX; b++) tensor_g2s_3d(sG + buf * C::GBUF + b * C::BOX, &gmap, 16 * b, row0, e, &bars[buf]);
    };
    int cur = next_active(i0), ld = cur, nld = 0, ncons = 0;
    for (int k = 0; k < 2; k++)
        if (ld < i1) { if (tid == 0) issue(ld, nld & 1); nld++; ld = next_active(ld + 1); }

    while (cur < i1) {
        const int quad = cur >> 2;
        unsigned present = 0;                           // fields of this quad that are transformed
        while (cur < i1 && (cur >> 2) == quad) {
            const int fs = cur & 3, buf = ncons & 1;
            const int fl = desc[cur % nbatch].flags;
            present |= 1u << fs;
            mbar_wait(&bars[buf], (ncons >> 1) & 1);
            // ---- stage A: Gaussian-weighted fold of the two hemispheres + radf3 + radf4 (fftpack.f90:774,844)
            {
                const double* G = sG + buf * C::GBUF;
                const bool scl = (fl & 3) != 0;
                const double ss = (fl & 1) ? cs1 : cs2, sn = (fl & 1) ? cn1 : cn2;
                double x[12];
#pragma unroll
                for (int jj = 0; jj < 3; jj++)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int box = 2 * jj + (j >> 1), inb = k3 + 8 * (j & 1);
                        double vs = G[box * C::BOX + lat_s * 16 + ((((inb >> 1) ^ (lat_s & 7))) << 1) + (inb & 1)];
                        double vn = G[box * C::BOX + lat_n * 16 + ((((inb >> 1) ^ (lat_n & 7))) << 1) + (inb & 1)];
                        if (scl) { vs *= ss; vn *= sn; }
                        x[4 * jj + j] = (par ? (vn - vs) : (vn + vs)) * wsc;
                    }
                Fft96F::stageA<C::XS>(x, sT + slotA, sWa, k3);
            }
            __syncthreads();                            // sT complete, grid buffer free
            if (ld < i1) { if (tid == 0) issue(ld, buf); nld++; ld = next_active(ld + 1); }
            // ---- stage B: radf4 + radf2 (fftpack.f90:844,722); half-complex position -> coefficient row, truncated at m = trunc
            if (tid < 7 * C::IL) {
                double* E = sEO + fs * C::FS + slotB;
                auto st = [E](int pos, double v) { if (pos <= 2 * C::TRUNC) E[(pos + (pos > 0)) * C::CS] = v; };
                const double* T = sT + slotB;
                if (setB < 5) Fft96F::stageB_general_f<C::XS>(T, st, sWa, 3 + 2 * setB);
                else if (setB == 5) Fft96F::stageB_first_f<C::XS>(T, st, sWa);
                else Fft96F::stageB_last_f<C::XS>(T, st, sWa);
            }
            __syncthreads();                            // EO slot complete, sT free
            ncons++;
            cur = next_active(cur + 1);
        }
        // ---- direct Legendre for the quad (legendre.f90:142-154): DMMA tiles, rows = 8 n of one parity, columns = (re, im) x 4 fields
        {
            const int idx = 4 * quad + q;               // this lane's output field
            const bool st_ok = idx < nf && ((present >> q) & 1);
            const int e = idx / nbatch, f = idx - e * nbatch;
            double* outp = out_base + (size_t)(st_ok ? e : 0) * out_ms + (size_t)(st_ok ? f : 0) * C::NSPEC2;
            const double* Bq = sEO + (g >> 1) * C::FS + (g & 1) * C::CS + q;
#pragma unroll
            for (int i = 0; i < C::SLOTS; i += 2) {
                const int La = i * C::WARPS + w, Lb = La + C::WARPS;
                const bool va = La < C::NLIVE, vb = Lb < C::NLIVE;
                const int ta = sTile[va ? La : 0], tb = sTile[vb ? Lb : 0];
                const int ma = ta & 255, pa = (ta >> 8) & 1, mb = tb & 255, pb = (tb >> 8) & 1;
                const double* Ba = Bq + 2 * ma * C::CS + 24 * pa;
                const double* Bb = Bq + 2 * mb * C::CS + 24 * pb;
                double ca0 = 0.0, ca1 = 0.0, cb0 = 0.0, cb1 = 0.0;
#pragma unroll
                for (int ks = 0; ks < C::KS; ks++) {
                    dmma884q(ca0, ca1, a[i][ks], Ba[4 * ks]);
                    dmma884q(cb0, cb1, a[i + 1][ks], Bb[4 * ks]);
                }
                if (st_ok) {
                    if (va) *reinterpret_cast<double2*>(outp + (pa + 2 * (8 * ((ta >> 9) & 1) + g)) * C::K2 + 2 * ma) = make_double2(ca0, ca1);
                    if (vb) *reinterpret_cast<double2*>(outp + (pb + 2 * (8 * ((tb >> 9) & 1) + g)) * C::K2 + 2 * mb) = make_double2(cb0, cb1);
                }
            }
            // rows beyond the triangle (n-tiles with no live coefficient): structural zeros
            if (st_ok)
                for (int L = C::NLIVE + w; L < C::NTILE; L += C::WARPS) {
                    const int td = sTile[L];
                    *reinterpret_cast<double2*>(outp + (((td >> 8) & 1) + 2 * (8 * ((td >> 9) & 1) + g)) * C::K2 + 2 * (td & 255)) = make_double2(0.0, 0.0);
                }
        }
        __syncthreads();                                // EO free for the next quad
    }
    if (tv.trace) { __syncthreads(); if (tid == 0) trace_end(tv.trace, 2); }
}

void setup_quad_kernels() {
    CUDA_CHECK(cudaFuncSetAttribute(k_g2s_quad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QCfg::SMEM_G2S));
}

void launch_g2s_quad(speedy_ctx* ctx, const CUtensorMap& gmap, const XDesc* d_desc, int nbatch, double* d_out, long long out_ms, int nmembers, const int* gate) {
    using C = QCfg;
    const int nquad = (nbatch * nmembers + 3) / 4;
    const int ncta = nquad < ctx->num_sms ? nquad : ctx->num_sms;
    CUDA_CHECK(launch_pdl(ctx->dv.trace == nullptr || ctx->trace_pdl, k_g2s_quad, dim3(ncta), dim3(C::THREADS), C::SMEM_G2S, ctx->stream, gmap, d_desc, nbatch, nmembers,
                          d_out, out_ms, ctx->dv, gate));
}

}  // namespace spd
