// K1 / K2 — batched spherical-harmonic transforms for sm_100a.
//
//   K1 spec_to_grid  = legendre_inv (legendre.f90:74-111) + fourier_inv (fourier.f90:23-53)
//   K2 grid_to_spec  = fourier_dir (fourier.f90:56-82)    + legendre_dir (legendre.f90:114-155)
//
// A single member's step has only 91 + 74 transforms of a 96x48 field, far too few to fill
// 148 SMs with one CTA per transform, and each CTA would be a long dependent chain.  So every
// transform is cut into independent slices:
//   K1: 3 latitude groups (the Legendre sum and the zonal FFT are both per latitude), CTA =
//       (field, latitude group, member);
//   K2: 4 groups of Fourier rows (the forward DFT is per wavenumber row and the direct
//       Legendre sum is per (m,n)), CTA = (field, wavenumber group, member).
// The Legendre contraction runs on the FP64 pipe with every P_n^m load of a sum in flight at
// once (fixed trip count, zero-padded triangle); the zonal Fourier step is applied as a dense
// real operator — the reference's FFTPACK transform including its single-precision constants
// (SURVEY.md F13), extracted on the host by tables.cpp — with FP64 tensor-core MMAs
// (mma.sync.m8n8k4.f64; tcgen05 has no FP64 kind).  Intermediate Fourier coefficients never
// leave shared memory.  K1 can also build its input on the fly from the prognostic fields
// (uvspec / grad, spectral.f90:124-196), which removes a kernel from the time step.
#include "ctx.h"
#include "spectral_ops.cuh"
#include "tma.cuh"
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include <type_traits>
#include "fft96.cuh"
#include "fft144.cuh"
#include "fft96f.cuh"
#include "close_step.cuh"

namespace spd {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int padmod16(int n, int r) { return n + ((r - n % 16) + 16) % 16; }

template <int TRUNC>
struct TCfg {
    static constexpr int MX = TRUNC + 1, NX = TRUNC + 2;
    static constexpr int IX = (TRUNC == 30) ? 96 : 144, IY = IX / 4, IL = IX / 2;
    static constexpr int K2 = 2 * MX, KP = (K2 + 7) / 8 * 8;
    static constexpr int NSPEC2 = NX * K2;          // doubles per spectral field
    // K1: latitude groups
    static constexpr int LG = 3, JG = IY / LG, NR = 2 * JG;   // latitude pairs / rows per CTA
    static constexpr int XS = padmod16(NR, 4);      // sX row stride: conflict-free B fragments
    static constexpr int K1_THREADS = IX / 8 * 32;
    static constexpr size_t K1_SMEM = sizeof(double) * (3 * NSPEC2 + KP * XS);
    // K2: groups of Fourier rows
    static constexpr int CG = 4, RG = KP / CG;      // rows per CTA (16 / 24)
    static constexpr int GS = padmod16(IX, 4);      // sG row stride
    static constexpr int YS = padmod16(IL, 8);      // sY row stride: conflict-free double2 C stores
    static constexpr int ES = IY + 1;               // even/odd fold stride (odd)
    static constexpr int K2_THREADS = 384;
    static constexpr size_t K2_SMEM = sizeof(double) * (IL * GS + RG * YS);
    static_assert(IY % LG == 0 && NR % 8 == 0, "latitude groups must be whole 8-row tiles");
    static_assert(KP % (8 * CG) == 0 && IX % 8 == 0 && IL % 8 == 0, "tile sizes");
    static_assert(2 * RG * ES <= IL * GS, "fold buffers must fit in the grid staging buffer");
};

// mode: 0 full spec->grid, 1 legendre_inv only (out = (2mx,il)), 2 fourier_inv only (in = (2mx,il))
template <int TRUNC>
__global__ void __launch_bounds__(TCfg<TRUNC>::K1_THREADS)
k_spec_to_grid(const double* __restrict__ in_base, long long in_ms, const XDesc* __restrict__ desc,
               double* __restrict__ out_base, long long out_ms, DevTables tv, int mode) {
    using C = TCfg<TRUNC>;
    extern __shared__ double smem[];
    double* sIn = smem;
    double* sA = smem + C::NSPEC2;
    double* sB = smem + 2 * C::NSPEC2;
    double* sX = smem + 3 * C::NSPEC2;
    const int b = blockIdx.x / C::LG, grp = blockIdx.x - b * C::LG, e = blockIdx.y, tid = threadIdx.x, nthr = blockDim.x;
    const XDesc dsc = desc[b];
    const double* mbase = in_base + (size_t)e * in_ms;
    const double* in = mbase + dsc.off;
    const int j0 = grp * C::JG;                       // first latitude pair of this CTA
    auto row_lat = [&](int r) { return (r < C::JG) ? (j0 + r) : (C::IL - 1 - (j0 + (r - C::JG))); };

    for (int t = tid; t < (C::KP - C::K2) * C::XS; t += nthr) sX[C::K2 * C::XS + t] = 0.0;
    if (mode != 2) {
        // ---- input stage.  Coefficients outside the triangle m+n <= trunc+1 are never read by
        // the reference (legendre.f90:38 nsh2); they are zeroed so that the sums below have a
        // fixed trip count.
        if (dsc.op == 0) {
            for (int t = tid; t < C::NSPEC2; t += nthr) {
                const int n = t / C::K2, c = t - n * C::K2;
                sIn[t] = ((c >> 1) + n <= C::MX) ? in[t] : 0.0;
            }
        } else {
            // derived input: 1 ucos, 2 vcos = uvspec(vor, div) (spectral.f90:173-196); 3 d/dx, 4 d/dy = grad(ps) (:124-144)
            const double* in2 = mbase + dsc.off2;
            for (int t = tid; t < C::NSPEC2; t += nthr) { sA[t] = in[t]; if (dsc.op <= 2) sB[t] = in2[t]; }
            __syncthreads();
            for (int t = tid; t < C::MX * C::NX; t += nthr) {
                const int n = t / C::MX, m = t - n * C::MX;
                cd r0, r1;
                if (dsc.op <= 2) dev_uvspec(tv, sA, sB, m, n, r0, r1);
                else dev_grad(tv, sA, m, n, r0, r1);
                cd r = (dsc.op == 1 || dsc.op == 3) ? r0 : r1;
                if (m + n > C::MX) r = cd{0.0, 0.0};
                st(sIn, C::MX, m, n, r);
            }
        }
        __syncthreads();
        // ---- inverse Legendre for this CTA's latitude pairs: even/odd split in n, hemispheric symmetry
        for (int t = tid; t < C::JG * C::KP; t += nthr) {
            const int jl = t / C::KP, c = t - jl * C::KP;
            if (c < C::K2) {
                const int m = c >> 1;
                const double* P = tv.poly + (size_t)(j0 + jl) * C::NX * C::MX + m;
                double ev = 0.0, od = 0.0;
#pragma unroll
                for (int n = 0; n < C::NX; n += 2) ev += sIn[n * C::K2 + c] * P[n * C::MX];
#pragma unroll
                for (int n = 1; n < C::NX; n += 2) od += sIn[n * C::K2 + c] * P[n * C::MX];
                sX[c * C::XS + jl] = ev - od;               // row j (southern)
                sX[c * C::XS + C::JG + jl] = ev + od;       // row il+1-j (northern)
            }
        }
    } else {
        for (int t = tid; t < C::NR * C::K2; t += nthr) {
            const int r = t / C::K2, c = t - r * C::K2;
            sX[c * C::XS + r] = in[(size_t)row_lat(r) * C::K2 + c];
        }
    }
    __syncthreads();
    if (mode == 1) {
        double* out = out_base + (size_t)e * out_ms + (size_t)b * C::K2 * C::IL;
        for (int t = tid; t < C::NR * C::K2; t += nthr) {
            const int r = t / C::K2, c = t - r * C::K2;
            out[(size_t)row_lat(r) * C::K2 + c] = sX[c * C::XS + r];
        }
        return;
    }
    // ---- dense backward Fourier operator on the FP64 tensor pipe:
    //   grid[i][r] = sum_c finv[i][c] * X[c][r],  M = IX, N = NR, K = KP
    const int w = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    double a[C::KP / 4];
    {
        const double* A = tv.finv + (size_t)(8 * w + g) * C::KP + q;
#pragma unroll
        for (int ks = 0; ks < C::KP / 4; ks++) a[ks] = A[4 * ks];
    }
    double* out = out_base + (size_t)e * out_ms + (size_t)b * C::IX * C::IL;
    const int i = 8 * w + g;
    const bool sc = dsc.flags & 1, ad = dsc.flags & 2;
#pragma unroll
    for (int nt = 0; nt < C::NR / 8; nt++) {
        double c0 = 0.0, c1 = 0.0;
        const double* B = sX + q * C::XS + 8 * nt + g;
#pragma unroll
        for (int ks = 0; ks < C::KP / 4; ks++) dmma884(c0, c1, a[ks], B[4 * ks * C::XS]);
        const int ja = row_lat(8 * nt + 2 * q), jb = row_lat(8 * nt + 2 * q + 1);
        if (sc) { c0 *= tv.cosgr[ja]; c1 *= tv.cosgr[jb]; }
        if (ad) { c0 += tv.coriol[ja]; c1 += tv.coriol[jb]; }
        out[(size_t)ja * C::IX + i] = c0;
        out[(size_t)jb * C::IX + i] = c1;
    }
}

// mode: 0 full grid->spec, 1 fourier_dir only (out = (2mx,il)), 2 legendre_dir only (in = (2mx,il))
template <int TRUNC>
__global__ void __launch_bounds__(TCfg<TRUNC>::K2_THREADS)
k_grid_to_spec(const double* __restrict__ in_base, long long in_ms, const XDesc* __restrict__ desc,
               double* __restrict__ out_base, long long out_ms, DevTables tv, int mode, const int* __restrict__ gate) {
    using C = TCfg<TRUNC>;
    extern __shared__ double smem[];
    double* sG = smem;
    double* sY = smem + C::IL * C::GS;
    const int b = blockIdx.x / C::CG, grp = blockIdx.x - b * C::CG, e = blockIdx.y, tid = threadIdx.x, nthr = blockDim.x;
    const XDesc dsc = desc[b];
    if (gate && (dsc.flags & 4) && !*gate) return;     // in-graph conditional work (the daily forcing transform)
    const double* in = in_base + (size_t)e * in_ms + dsc.off;
    const int w = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3, nw = nthr >> 5;
    const int c0row = grp * C::RG;                     // first Fourier row of this CTA

    if (mode != 2) {
        const double* scl = (dsc.flags & 1) ? tv.cosgr : ((dsc.flags & 2) ? tv.cosgr2 : nullptr);
        for (int t = tid; t < C::IL * C::IX; t += nthr) {
            const int j = t / C::IX, i = t - j * C::IX;
            double v = in[t];
            if (scl) v *= scl[j];
            sG[j * C::GS + i] = v;
        }
        __syncthreads();
        // dense forward Fourier operator for this CTA's rows: Y[c][j] = sum_i ffwd[c][i] * g[i][j],  M = RG, N = IL, K = IX
        constexpr int MT = C::RG / 8, NT = C::IL / 8;
        for (int tile = w; tile < MT * NT; tile += nw) {
            const int mt = tile / NT, nt = tile - mt * NT;
            const int crow = c0row + 8 * mt + g;
            const double* A = tv.ffwd + (size_t)crow * C::IX + q;
            double c0 = 0.0, c1 = 0.0;
            const double* B = sG + (8 * nt + g) * C::GS + q;
#pragma unroll
            for (int ks = 0; ks < C::IX / 4; ks++) dmma884(c0, c1, A[4 * ks], B[4 * ks]);
            *reinterpret_cast<double2*>(sY + (8 * mt + g) * C::YS + 8 * nt + 2 * q) = make_double2(c0, c1);
        }
    } else {
        for (int t = tid; t < C::IL * C::RG; t += nthr) {
            const int j = t / C::RG, cl = t - j * C::RG;
            const int c = c0row + cl;
            sY[cl * C::YS + j] = (c < C::K2) ? in[(size_t)j * C::K2 + c] : 0.0;
        }
    }
    __syncthreads();
    if (mode == 1) {
        double* out = out_base + (size_t)e * out_ms + (size_t)b * C::K2 * C::IL;
        for (int t = tid; t < C::IL * C::RG; t += nthr) {
            const int j = t / C::RG, cl = t - j * C::RG;
            const int c = c0row + cl;
            if (c < C::K2) out[(size_t)j * C::K2 + c] = sY[cl * C::YS + j];
        }
        return;
    }
    // Gaussian-weighted even/odd fold (legendre.f90:127-133); sG is free now
    double* sE = sG;
    double* sO = sG + C::RG * C::ES;
    for (int t = tid; t < C::RG * C::IY; t += nthr) {
        const int cl = t / C::IY, jh = t - cl * C::IY;
        const double south = sY[cl * C::YS + jh], north = sY[cl * C::YS + (C::IL - 1 - jh)];
        const double wgt = tv.wt[jh];
        sE[cl * C::ES + jh] = (north + south) * wgt;
        sO[cl * C::ES + jh] = (north - south) * wgt;
    }
    __syncthreads();
    // direct Legendre: out(c,n) = sum_j P(m,n,j) * {even|odd}(c,j), n <= trunc (legendre.f90:142-154)
    double* out = out_base + (size_t)e * out_ms + (size_t)b * C::K2 * C::NX;
    for (int t = tid; t < C::NX * C::RG; t += nthr) {
        const int n = t / C::RG, cl = t - n * C::RG;
        const int c = c0row + cl;
        if (c >= C::K2) continue;
        const int m = c >> 1;
        double s = 0.0;
        if (n <= TRUNC && m + n <= C::MX) {
            const double* P = tv.poly + (size_t)n * C::MX + m;
            const double* F = ((n & 1) ? sO : sE) + cl * C::ES;
#pragma unroll
            for (int jh = 0; jh < C::IY; jh++) s += P[(size_t)jh * C::NX * C::MX] * F[jh];
        }
        out[n * C::K2 + c] = s;
    }
}


// ==========================================================================================
// Streaming variants of K1 / K2 (mode 0, the time-step path).
//
// One persistent CTA per SM owns a slice of the transform — a latitude group (K1) or a group of
// Fourier rows (K2) — and streams a chunk of FIELDS through it.  Everything that does not depend
// on the field is staged once per CTA by bulk asynchronous copies (TMA): the P_n^m tile of the
// slice (contiguous in [j][n][m]; re-laid out per wavenumber group for the direct transform) and
// the rows of the dense Fourier operator.  The fields themselves arrive through one mbarrier-
// tracked staging buffer: as soon as a field has been consumed into the working buffer the copy
// of the next field is issued, so its L2/HBM latency hides behind the Legendre + Fourier work of
// the current one.  The only global loads left on the dependent chain are per-field scalars.
// ==========================================================================================
template <int TRUNC>
struct SCfg : TCfg<TRUNC> {
    using B = TCfg<TRUNC>;
    // K1: latitude groups sized so that the P tile + staging fit in 227 KB
    static constexpr int LG = (TRUNC == 30) ? 3 : 9, JG = B::IY / LG, NR = 2 * JG;
    static constexpr int XS = padmod16(NR, 4);
    // P tile: per latitude only the triangle m + n <= trunc + 1 (legendre.f90:38), rows of n packed back to back.
    // A read past the end of a row meets the next row (finite) times a masked-out coefficient (zero).
    __host__ __device__ static constexpr int tri_cnt(int n) { return (B::MX - n + 1 < B::MX) ? (B::MX - n + 1) : B::MX; }
    __host__ __device__ static constexpr int tri_off(int n) { int o = 0; for (int i = 0; i < n; i++) o += tri_cnt(i); return o; }
    // terms of the inverse Legendre sum that can be non-zero for wavenumbers m >= mw * w (m + n <= MX)
    __host__ __device__ static constexpr int leg_terms(int w, int mw) { int t = B::MX + 1 - mw * w; return t > B::NX ? B::NX : (t < 1 ? 1 : t); }
    static constexpr int TR = (tri_off(B::NX) + 1) / 2 * 2;       // doubles per latitude in the packed table (even: 16-byte rows)
    static constexpr int PS = TR + B::MX + (24 - (TR + B::MX) % 16) % 16;   // shared row: >= MX zeroed pad, stride = 8 mod 16 (alternating bank halves)
    static constexpr int PT = JG * PS;                            // P tile, doubles
    static constexpr int MP = (B::MX + 31) / 32 * 32;             // m padded to whole warps
    static constexpr int K1_THREADS = B::IX / 8 * 32;
    static constexpr size_t K1_SMEM = sizeof(double) * (PT + 3 * B::NSPEC2 + B::KP * XS) + 2 * sizeof(uint64_t);
    // FFT variant of the Fourier stage (fft96.cuh / fft144.cuh): half-complex rows + the buffer between its two stages,
    // position-major with an odd row stride (the Legendre stage's scattered writes and the FFT's row-wise reads spread over the banks), + twiddles
    static constexpr int XF = NR + 1;
    static constexpr bool HAS_FFT = (B::IX == 96 && NR == 16) || (B::IX == 144 && NR == 8);
    static constexpr size_t K1_SMEM_FFT = sizeof(double) * (PT + 3 * B::NSPEC2 + 2 * B::IX * XF + B::IX) + 2 * sizeof(uint64_t);
    // K2: groups of 16 Fourier rows = 8 zonal wavenumbers
    static constexpr int RG = 16, CG = B::KP / RG, MG = RG / 2;
    static constexpr int GS = padmod16(B::IX, 4), FS = GS, YS = padmod16(B::IL, 8), ES = B::IY + 1 - (B::IY & 1);   // ES odd
    static constexpr bool P_SMEM = (TRUNC == 30);                 // T47: 113 KB tile does not fit beside the grid buffer
    static constexpr int PD = B::IY * B::NX * MG;                 // re-laid-out P tile [jh][n][mloc]
    static constexpr int MT = RG / 8, NT = B::IL / 8;
    static constexpr int K2_THREADS = 32 * (MT * NT < 18 ? MT * NT : 18);
    static constexpr int OOFF = RG * ES + 1;                      // sO behind sE, shifted one bank
    // latency variant: the even/odd folds alias the grid buffer (dead after the Fourier stage), 107 KB at T30: two CTAs per SM;
    // batch variant: separate fold buffers (the next field is already streaming into the grid buffer) + the raw field buffer
    // the grid field arrives as IX/16 tensor-map boxes of [16 longitudes x IL latitudes] (128-byte rows, SWIZZLE_128B): the B
    // fragments of the DMMA read it conflict-free without row padding, and no re-layout pass is needed
    static constexpr int NBOX = B::IX / 16, BOX = 16 * B::IL;
    static constexpr size_t K2_SMEM = sizeof(double) * (B::IL * B::IX + RG * FS + RG * YS + (P_SMEM ? PD : 0)) + 2 * sizeof(uint64_t);
    static constexpr size_t K2_SMEM_BATCH = K2_SMEM + sizeof(double) * (2 * RG * ES + 2);
    static_assert(2 * RG * ES + 2 <= B::IL * B::IX && B::IX % 16 == 0 && (BOX * 8) % 1024 == 0, "fold buffers fit in the grid buffer; whole swizzle atoms");
    static_assert(B::IY % LG == 0 && NR % 8 == 0 && JG % 4 == 0 && 32 % (JG / 2) == 0 && JG * MP <= K1_THREADS, "K1 tiling");
    static_assert(B::KP % RG == 0 && (PS * 8) % 16 == 0 && PS % 16 == 8 && PS >= TR + B::MX && (TR * 8) % 16 == 0 && (B::NSPEC2 * 8) % 16 == 0 && (PD * 8) % 16 == 0, "K2 tiling / bulk-copy sizes");
    static_assert(K1_SMEM <= 232448 && K1_SMEM_FFT <= 232448 && K2_SMEM_BATCH <= 232448, "shared memory budget");
};

template <int TRUNC, bool BATCH, bool FFT>
__global__ void __maxnreg__(TRUNC == 30 ? 80 : 96)    // T30: <= 80 registers, two CTAs per SM (their Legendre and Fourier phases overlap); T47: one CTA of 576 threads
k_s2g_stream(const double* __restrict__ in_base, long long in_ms, const XDesc* __restrict__ desc, int nbatch, int nchunk,
             double* __restrict__ out_base, long long out_ms, DevTables tv, CloseArgs cl) {
    using C = SCfg<TRUNC>;
    if (blockIdx.x == (unsigned)(nchunk * C::LG)) {     // the extra CTA: closes the previous step, off the critical path
        pdl_wait();
        pdl_trigger();
        if (blockIdx.y == 0) close_step_cta(cl, threadIdx.x);     // one closer for the whole member batch
        return;
    }
    extern __shared__ __align__(16) double smem[];
    double* sP = smem;
    double* sA = sP + C::PT;                 // staging: source field(s) of the current transform
    double* sB = sA + C::NSPEC2;
    double* sIn = sB + C::NSPEC2;
    double* sX = sIn + C::NSPEC2;
    double* sT = sX + C::IX * C::XF;         // FFT variant: between the two FFT stages
    double* sWa = sT + C::IX * C::XF;        // FFT variant: twiddle table of rffti1
    uint64_t* bars = reinterpret_cast<uint64_t*>(FFT ? sWa + C::IX : sX + C::KP * C::XS);   // [0] P tile, [1] staging
    const int tid = threadIdx.x, nthr = blockDim.x;
    if (tid == 0) trace_begin(tv.trace, 0);
    const int grp = blockIdx.x % C::LG, chunk = blockIdx.x / C::LG, e = blockIdx.y;
    const int f0 = (int)((long long)chunk * nbatch / nchunk), f1 = (int)((long long)(chunk + 1) * nbatch / nchunk);
    const double* mbase = in_base + (size_t)e * in_ms;
    const int j0 = grp * C::JG;
    auto row_lat = [&](int r) { return (r < C::JG) ? (j0 + r) : (C::IL - 1 - (j0 + (r - C::JG))); };
    __shared__ XDesc sDesc[104];             // descriptors of this CTA's fields (constant data: fetched in the prologue)
    auto issue = [&](int f) {                // one thread: bulk copies of field f's source(s)
        const XDesc d = (f - f0 < 104) ? sDesc[f - f0] : desc[f];
        const uint32_t bytes = C::NSPEC2 * sizeof(double);
        const bool two = d.op == 1 || d.op == 2;
        mbar_expect_tx(&bars[1], two ? 2 * bytes : bytes);
        bulk_g2s(sA, mbase + d.off, bytes, &bars[1]);
        if (two) bulk_g2s(sB, mbase + d.off2, bytes, &bars[1]);
    };
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    __syncthreads();
    // prologue on constant tables only (may overlap the tail of the previous kernel: PDL)
    if (tid == 0) mbar_expect_tx(&bars[0], C::JG * C::TR * sizeof(double));
    if (tid < C::JG) bulk_g2s(sP + tid * C::PS, tv.polyt + (size_t)(j0 + tid) * C::TR, C::TR * sizeof(double), &bars[0]);
    for (int t = tid; t < C::JG * (C::PS - C::TR); t += nthr) sP[(t / (C::PS - C::TR)) * C::PS + C::TR + t % (C::PS - C::TR)] = 0.0;   // row pads
    for (int t = tid; t < f1 - f0 && t < 104; t += nthr) sDesc[t] = desc[f0 + t];
    // this warp's A fragments of the dense backward Fourier operator stay in registers for every field
    const int w = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    double a[FFT ? 1 : C::KP / 4];
    // latitude factors of the epilogue (fourier.f90:47-51, tendencies.f90:103) for this thread's output rows
    double cg[FFT ? 1 : C::NR / 4], cf[FFT ? 1 : C::NR / 4];
    if constexpr (FFT) {
        // half-complex rows hold positions 0..2*trunc (fourier.f90:40-45); the zero padding up to ix is written once
        for (int t = tid; t < (C::IX - (C::K2 - 1)) * C::XF; t += nthr) sX[(C::K2 - 1) * C::XF + t] = 0.0;
        for (int t = tid; t < C::IX; t += nthr) sWa[t] = tv.fftwa[t];
        const int jr = row_lat((tid >> (C::IX == 96 ? 3 : 4)) & (C::NR - 1));      // stage 2: thread = (row, k of the third pass)
        cg[0] = tv.cosgr[jr]; cf[0] = tv.coriol[jr];
    } else {
        const double* A = tv.finv + (size_t)(8 * w + g) * C::KP + q;
#pragma unroll
        for (int ks = 0; ks < C::KP / 4; ks++) a[ks] = A[4 * ks];
        for (int t = tid; t < (C::KP - C::K2) * C::XS; t += nthr) sX[C::K2 * C::XS + t] = 0.0;
#pragma unroll
        for (int nt = 0; nt < C::NR / 8; nt++) {
            const int ja = row_lat(8 * nt + 2 * q), jb = row_lat(8 * nt + 2 * q + 1);
            cg[2 * nt] = tv.cosgr[ja]; cg[2 * nt + 1] = tv.cosgr[jb];
            cf[2 * nt] = tv.coriol[ja]; cf[2 * nt + 1] = tv.coriol[jb];
        }
    }
    // FFT variant (registers to spare): the operator-table entries of the FIRST field's derived input (uvspec / grad) are
    // fetched here, ahead of the wait — they are constants, and on the single-member step a CTA has one field
    constexpr int NE = (C::MX * C::NX + C::K1_THREADS - 1) / C::K1_THREADS;
    constexpr bool PRE = FFT && !BATCH;
    double pre[PRE ? NE : 1][3];
    if constexpr (PRE) {
        const int op0 = (f0 < f1) ? desc[f0].op : 0;
#pragma unroll
        for (int u = 0; u < NE; u++) {
            const int t = tid + u * C::K1_THREADS;
            pre[u][0] = pre[u][1] = pre[u][2] = 0.0;
            if (op0 != 0 && t < C::MX * C::NX) {
                const int m = t % C::MX;
                pre[u][0] = (op0 <= 2) ? tv.uvdx[t] : tv.gradx[m];
                pre[u][1] = (op0 <= 2) ? tv.uvdym[t] : tv.gradym[t];
                pre[u][2] = (op0 <= 2) ? tv.uvdyp[t] : tv.gradyp[t];
            }
        }
    }
    pdl_wait();                                        // the spectral fields of the previous kernel are complete
    pdl_trigger();
    const unsigned long long tk0 = tv.trace ? gtimer() : 0ull;
    // block 96: a derived (uvspec) field at T30; batch variant: block 4, third and fourth field of its chunk (steady state)
#define KSTAMP(i) do { if (tv.trace && tid == 0 && blockIdx.x == (BATCH ? 4 : 96) && blockIdx.y == 0) { const int si_ = (i) - (BATCH ? 8 : 0); if (si_ >= 0 && si_ < 8) tv.trace[48 + si_] += gtimer() - tk0; } } while (0)
    if (tid == 0 && f0 < f1) issue(f0);
    // Legendre work items: one thread owns a zonal wavenumber m and TWO latitude pairs (jlA, jlA + JG/2), so that the
    // spectral coefficients it reads serve 8 sums (the stage is bound by shared-memory bandwidth); a warp covers
    // JG/2 latitude pairs x 32/(JG/2) wavenumbers: P reads fall in alternating bank halves (row stride = 8 mod 16)
    constexpr int JH = C::JG / 2, MW = 32 / JH;
    const int jlA = lane / MW, jlB = jlA + JH, m2 = MW * w + lane % MW;
    const bool leg2 = w < C::MP / MW && m2 < C::MX;
    __syncthreads();                                   // sDesc

    for (int f = f0; f < f1; f++) {
        const XDesc dsc = (f - f0 < 104) ? sDesc[f - f0] : desc[f];      // beyond the prefetched window: a plain (cached) load
        const bool sc = dsc.flags & 1, ad = dsc.flags & 2;
        mbar_wait(&bars[1], (f - f0) & 1);
        KSTAMP(4 * (f - f0) + 0);
        // ---- input stage.  Coefficients outside the triangle m+n <= trunc+1 are never read by the
        // reference (legendre.f90:38 nsh2); they are zeroed so that the sums have a fixed trip count.
        if (dsc.op == 0) {
            for (int t = tid; t < C::NSPEC2; t += nthr) {
                const int n = t / C::K2, c = t - n * C::K2;
                sIn[t] = ((c >> 1) + n <= C::MX) ? sA[t] : 0.0;
            }
        } else {
            // derived input: 1 ucos, 2 vcos = uvspec(vor, div) (spectral.f90:173-196); 3 d/dx, 4 d/dy = grad(ps) (:124-144)
            // fixed trip count, fully unrolled: the operator-table loads of all of a thread's coefficients are in flight together;
            // one instantiation per component, each reading only its own stencil (half the shared-memory loads of a full uvspec)
            // (the single-member T30 variant keeps ONE instruction stream that evaluates both components: its code runs once per
            // CTA, and four specialised copies measured 0.1 us slower per step there, against -1.5 % at 8 members and -5 % at T47)
            constexpr bool SPLIT = BATCH || TRUNC != 30;
            if constexpr (!SPLIT) {
#pragma unroll
                for (int u = 0; u < NE; u++) {
                    const int t = tid + u * C::K1_THREADS;
                    if (t < C::MX * C::NX) {
                        const int n = t / C::MX, m = t - n * C::MX;
                        cd r0, r1;
                        double t0, t1, t2;
                        if (PRE && f == f0) { t0 = pre[u][0]; t1 = pre[u][1]; t2 = pre[u][2]; }
                        else if (dsc.op <= 2) { t0 = tv.uvdx[t]; t1 = tv.uvdym[t]; t2 = tv.uvdyp[t]; }
                        else { t0 = tv.gradx[m]; t1 = tv.gradym[t]; t2 = tv.gradyp[t]; }
                        if (dsc.op <= 2) dev_uvspec_t(C::MX, C::NX, TRUNC, sA, sB, m, n, t0, t1, t2, r0, r1);
                        else dev_grad_t(C::MX, C::NX, TRUNC, sA, m, n, t0, t1, t2, r0, r1);
                        cd r = (dsc.op == 1 || dsc.op == 3) ? r0 : r1;
                        if (m + n > C::MX) r = cd{0.0, 0.0};
                        st(sIn, C::MX, m, n, r);
                    }
                }
            } else {
                auto derived = [&](auto opc) {
                    constexpr int OP = decltype(opc)::value;
#pragma unroll
                    for (int u = 0; u < NE; u++) {
                        const int t = tid + u * C::K1_THREADS;
                        if (t < C::MX * C::NX) {
                            const int n = t / C::MX, m = t - n * C::MX;
                            double t0, t1, t2;
                            if (PRE && f == f0) { t0 = pre[u][0]; t1 = pre[u][1]; t2 = pre[u][2]; }
                            else if (OP <= 2) { t0 = tv.uvdx[t]; t1 = tv.uvdym[t]; t2 = tv.uvdyp[t]; }
                            else { t0 = tv.gradx[m]; t1 = tv.gradym[t]; t2 = tv.gradyp[t]; }
                            cd r;
                            if constexpr (OP == 1) r = dev_ucos_t(C::MX, C::NX, TRUNC, sA, sB, m, n, t0, t1, t2);
                            else if constexpr (OP == 2) r = dev_vcos_t(C::MX, C::NX, TRUNC, sA, sB, m, n, t0, t1, t2);
                            else if constexpr (OP == 3) r = dev_gradx_t(C::MX, sA, m, n, t0);
                            else r = dev_grady_t(C::MX, C::NX, TRUNC, sA, m, n, t1, t2);
                            if (m + n > C::MX) r = cd{0.0, 0.0};
                            st(sIn, C::MX, m, n, r);
                        }
                    }
                };
                switch (dsc.op) {
                    case 1: derived(std::integral_constant<int, 1>{}); break;
                    case 2: derived(std::integral_constant<int, 2>{}); break;
                    case 3: derived(std::integral_constant<int, 3>{}); break;
                    default: derived(std::integral_constant<int, 4>{}); break;
                }
            }
        }
        __syncthreads();                      // sIn complete, staging free
        if (tid == 0 && f + 1 < f1) issue(f + 1);
        if (f == f0) mbar_wait(&bars[0], 0);
        KSTAMP(4 * (f - f0) + 1);
        // ---- inverse Legendre for this CTA's latitude pairs (legendre.f90:74-111): one thread per
        // (latitude pair, m); real and imaginary sums share the P values
        if (leg2) {
            const double* PA = sP + (size_t)jlA * C::PS + m2;
            const double* PB = sP + (size_t)jlB * C::PS + m2;
            const double2* X = reinterpret_cast<const double2*>(sIn) + m2;
            // the wavenumbers of warp w start at MW*w: their coefficients vanish for n > MX - MW*w (triangular truncation), so
            // the higher warps leave the sum early.  One instruction stream for all warps: blocks of 8 terms, fully unrolled,
            // with a warp-uniform exit between blocks; each accumulator still adds its terms in ascending n.
            double ear = 0.0, eai = 0.0, oar = 0.0, oai = 0.0, ebr = 0.0, ebi = 0.0, obr = 0.0, obi = 0.0;
            const int nmax = C::leg_terms(w, MW);
#pragma unroll
            for (int nb = 0; nb < C::NX; nb += 8) {
                if (nb >= nmax) break;
#pragma unroll
                for (int n = nb; n < nb + 8 && n < C::NX; n += 2) {
                    const double2 x = X[n * C::MX]; const double pa = PA[C::tri_off(n)], pb = PB[C::tri_off(n)];
                    ear += x.x * pa; eai += x.y * pa; ebr += x.x * pb; ebi += x.y * pb;
                }
#pragma unroll
                for (int n = nb + 1; n < nb + 8 && n < C::NX; n += 2) {
                    const double2 x = X[n * C::MX]; const double pa = PA[C::tri_off(n)], pb = PB[C::tri_off(n)];
                    oar += x.x * pa; oai += x.y * pa; obr += x.x * pb; obi += x.y * pb;
                }
            }
            if constexpr (FFT) {
                // FFTPACK's half-complex order (fourier.f90:40-45): a0, then (re, im) of m = 1..; the imaginary part of m = 0 is dropped
                double* xr = sX + (m2 == 0 ? 0 : 2 * m2 - 1) * C::XF;
                xr[jlA] = ear - oar;  xr[C::JG + jlA] = ear + oar;
                xr[jlB] = ebr - obr;  xr[C::JG + jlB] = ebr + obr;
                if (m2) {
                    xr[C::XF + jlA] = eai - oai;  xr[C::XF + C::JG + jlA] = eai + oai;
                    xr[C::XF + jlB] = ebi - obi;  xr[C::XF + C::JG + jlB] = ebi + obi;
                }
            } else {
                double* xr = sX + (2 * m2) * C::XS;
                xr[jlA] = ear - oar;  xr[C::XS + jlA] = eai - oai;                       // row j (southern)
                xr[C::JG + jlA] = ear + oar;  xr[C::XS + C::JG + jlA] = eai + oai;       // row il+1-j (northern)
                xr[jlB] = ebr - obr;  xr[C::XS + jlB] = ebi - obi;
                xr[C::JG + jlB] = ebr + obr;  xr[C::XS + C::JG + jlB] = ebi + obi;
            }
        }
        __syncthreads();
        KSTAMP(4 * (f - f0) + 2);
        double* out = out_base + (size_t)e * out_ms + (size_t)(dsc.oslot1 ? dsc.oslot1 - 1 : f) * C::IX * C::IL;
        if constexpr (FFT) {
            // ---- backward FFT per latitude row, FFTPACK's passes regrouped into two register-resident stages
            // (fft96.cuh at T30: 16 rows per CTA; fft144.cuh at T47: 8 rows per CTA)
            if constexpr (C::IX == 96) {
                // stage 1: half-warp = one closed input set x the 16 rows
                if (tid < 128) {
                    const int h = tid >> 4, r = tid & 15;
                    if (h < 5) Fft96::stage1_general<C::XF>(sX + r, sT + r, sWa, 3 + 2 * h);
                    else if (h == 6) Fft96::stage1_first<C::XF>(sX + r, sT + r, sWa);
                    else if (h == 7) Fft96::stage1_last<C::XF>(sX + r, sT + r, sWa);
                }
                __syncthreads();
                // stage 2: thread = (row, k): 8 consecutive lanes write 8 consecutive longitudes
                if (tid < 128) {
                    const int r = tid >> 3, k3 = tid & 7;
                    double y[12];
                    Fft96::stage2<C::XF>(sT + r, sWa, k3, y);
                    double* orow = out + (size_t)row_lat(r) * C::IX + k3;
#pragma unroll
                    for (int t = 0; t < 12; t++) {
                        double v = y[t];
                        if (sc) v *= cg[0];
                        if (ad) v += cf[0];
                        orow[8 * (t & 3) + 32 * (t >> 2)] = v;
                    }
                }
            } else {
                // stage 1: quarter-warp = one (input set, k pair) x the 8 rows; warp 0: k = 1,2, warp 1: k = 3,4, warp 2: the i = 1 sets
                if (tid < 80) {
                    const int role = tid >> 3, r = tid & 7;
                    if (role < 8) Fft144::stage1_general<C::XF>(sX + r, sT + r, sWa, 3 + 2 * (role & 3), role >> 2);
                    else Fft144::stage1_first<C::XF>(sX + r, sT + r, sWa, role - 8);
                }
                __syncthreads();
                // stage 2: thread = (row, k): 16 consecutive lanes write 16 consecutive longitudes
                if (tid < 128) {
                    const int r = tid >> 4, k3 = tid & 15;
                    double y[9];
                    Fft144::stage2<C::XF>(sT + r, sWa, k3, y);
                    double* orow = out + (size_t)row_lat(r) * C::IX + k3;
#pragma unroll
                    for (int t = 0; t < 9; t++) {
                        double v = y[t];
                        if (sc) v *= cg[0];
                        if (ad) v += cf[0];
                        orow[16 * (t % 3) + 48 * (t / 3)] = v;
                    }
                }
            }
        } else {
        // ---- dense backward Fourier operator on the FP64 tensor pipe:
        //   grid[i][r] = sum_c finv[i][c] * X[c][r],  M = IX, N = NR, K = KP
        const int i = 8 * w + g;
#pragma unroll
        for (int nt = 0; nt < C::NR / 8; nt++) {
            double c0 = 0.0, c1 = 0.0;
            const double* Bf = sX + q * C::XS + 8 * nt + g;
#pragma unroll
            for (int ks = 0; ks < C::KP / 4; ks++) dmma884(c0, c1, a[ks], Bf[4 * ks * C::XS]);
            const int ja = row_lat(8 * nt + 2 * q), jb = row_lat(8 * nt + 2 * q + 1);
            if (sc) { c0 *= cg[2 * nt]; c1 *= cg[2 * nt + 1]; }
            if (ad) { c0 += cf[2 * nt]; c1 += cf[2 * nt + 1]; }
            out[(size_t)ja * C::IX + i] = c0;
            out[(size_t)jb * C::IX + i] = c1;
        }
        }
        KSTAMP(4 * (f - f0) + 3);
        // the next field's Legendre stage rewrites sX only after the next __syncthreads pair
    }
#undef KSTAMP
    if (tv.trace) { __syncthreads(); if (tid == 0) trace_end(tv.trace, 0); }
}

template <int TRUNC, bool BATCH>
__global__ void __launch_bounds__(SCfg<TRUNC>::K2_THREADS, TRUNC == 30 ? 2 : 1)
k_g2s_stream(const __grid_constant__ CUtensorMap gmap, const XDesc* __restrict__ desc, int nbatch, int nchunk,
             double* __restrict__ out_base, long long out_ms, DevTables tv, const int* __restrict__ gate) {
    using C = SCfg<TRUNC>;
    extern __shared__ __align__(1024) double smem[];
    double* sG = smem;                                  // [NBOX][IL][16] grid field, 16-byte chunks of a row XOR-swizzled with the row index (TMA SWIZZLE_128B)
    double* sF = sG + C::IL * C::IX;                    // [RG][FS] rows of the dense forward Fourier operator
    double* sY = sF + C::RG * C::FS;                    // [RG][YS] Fourier coefficients of this group
    double* sPd = sY + C::RG * C::YS;                   // [IY][NX][MG]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sPd + (C::P_SMEM ? C::PD : 0));   // [0] operator + P tiles, [1] grid field
    double* sE = BATCH ? reinterpret_cast<double*>(bars + 2) : sG;   // even / odd folds (latency variant: in the dead grid buffer)
    double* sO = sE + C::OOFF;
    const int tid = threadIdx.x, nthr = blockDim.x;
    if (tid == 0) trace_begin(tv.trace, 2);
    const int grp = blockIdx.x % C::CG, chunk = blockIdx.x / C::CG, e = blockIdx.y;
    const int f0 = (int)((long long)chunk * nbatch / nchunk), f1 = (int)((long long)(chunk + 1) * nbatch / nchunk);
    const int c0row = grp * C::RG;
    int gate_open = 1;
    auto live = [&](int f) { return gate_open || !(desc[f].flags & 4); };
    auto next_live = [&](int f) { while (f < f1 && !live(f)) f++; return f; };
    const uint32_t rowb = C::IX * sizeof(double);
    // a field = NBOX tensor-map boxes issued by one thread (48 row copies took ~1 us to drain through the TMA unit); the map is
    // [longitude][row of IX doubles][member], a field at element offset `off` starts at row off / IX
    auto issue_off = [&](long long off) {
        if (tid == 0) {
            const int row0 = (int)(off / C::IX);
            if (off != (long long)row0 * C::IX) __trap();   // fields must start on whole rows of the input map
            fence_proxy_async();                        // the folds of the previous field (generic writes) alias this buffer
            mbar_expect_tx(&bars[1], C::IL * rowb);
#pragma unroll
            for (int b = 0; b < C::NBOX; b++) tensor_g2s_3d(sG + b * C::BOX, &gmap, 16 * b, row0, e, &bars[1]);
        }
    };
    auto issue = [&](int f) { if (tid == 0) issue_off(desc[f].off); };
    if (tid == 0 && (smem_u32(sG) & 1023u)) __trap();   // the swizzle pattern is a function of the address: the boxes must start on 1 KB
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    __syncthreads();
    // prologue on constant tables only (may overlap the tail of the previous kernel: PDL)
    if (tid == 0) mbar_expect_tx(&bars[0], C::RG * rowb + (C::P_SMEM ? C::PD * sizeof(double) : 0));
    if (tid < C::RG) bulk_g2s(sF + tid * C::FS, tv.ffwd + (size_t)(c0row + tid) * C::IX, rowb, &bars[0]);
    if (C::P_SMEM && tid == C::RG) bulk_g2s(sPd, tv.polyd + (size_t)grp * C::PD, C::PD * sizeof(double), &bars[0]);
    // constants this thread needs later are fetched here too, ahead of the dependency wait: the first field's descriptor, the
    // latitude factors of this warp's DMMA tile (one tile per warp) and this thread's Gaussian weight of the fold — each would
    // otherwise cost an L2 round trip on the critical path of a CTA that handles a single field
    const int w = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3, nw = nthr >> 5;
    static_assert(C::MT * C::NT * 32 == C::K2_THREADS && C::RG * C::IY == C::K2_THREADS, "one DMMA tile per warp, one fold element per thread");
    const XDesc d0 = (f0 < f1) ? desc[f0] : XDesc{0, 0, 0, 0, 0};
    const int nt_w = w % C::NT;
    const double sj_cosgr = tv.cosgr[8 * nt_w + g], sj_cosgr2 = tv.cosgr2[8 * nt_w + g];
    const double wgt_t = tv.wt[tid % C::IY];
    pdl_wait();                                        // the grid fields of the previous kernel are complete
    pdl_trigger();
    int f = f0;
    if (f0 < f1 && !(d0.flags & 4)) {                   // not a gated field: its copy does not wait for the gate
        issue_off(d0.off);
        gate_open = gate ? *gate : 1;
    } else {
        gate_open = gate ? *gate : 1;                   // in-graph conditional work (the daily forcing transform)
        f = next_live(f0);
        if (f < f1) issue(f);
    }
    int it = 0;
    for (; f < f1; it++) {
        const XDesc dsc = (f == f0) ? d0 : desc[f];
        const bool scl = (dsc.flags & 3) != 0;
        const double sj = (dsc.flags & 1) ? sj_cosgr : sj_cosgr2;
        if (it == 0) mbar_wait(&bars[0], 0);
        mbar_wait(&bars[1], it & 1);
        const int fn = next_live(f + 1);
        // ---- dense forward Fourier operator for this CTA's rows (fourier.f90:56-82), with the cosgr / cosgr2 pre-scale of
        // vdspec (spectral.f90:208-222):  Y[c][j] = sum_i ffwd[c][i] * (g[i][j] * scl[j]),  M = RG, N = IL, K = IX.
        // B fragment of lane (g, q) at k-step ks: latitude 8 nt + g, longitude 4 ks + q = box ks / 4, 16-byte chunk
        // 2 (ks % 4) + q / 2 of the row, XORed with the row index mod 8 (= g)
        int swz[4];
#pragma unroll
        for (int kk = 0; kk < 4; kk++) swz[kk] = (((2 * kk + (q >> 1)) ^ g) << 1) + (q & 1);
        for (int tile = w; tile < C::MT * C::NT; tile += nw) {
            const int mt = tile / C::NT, nt = tile - mt * C::NT;
            const double* A = sF + (8 * mt + g) * C::FS + q;
            const double* Bf = sG + (8 * nt + g) * 16;
            double c0 = 0.0, c1 = 0.0;
            if (scl) {
#pragma unroll
                for (int ks = 0; ks < C::IX / 4; ks++) dmma884(c0, c1, A[4 * ks], Bf[(ks >> 2) * C::BOX + swz[ks & 3]] * sj);
            } else {
#pragma unroll
                for (int ks = 0; ks < C::IX / 4; ks++) dmma884(c0, c1, A[4 * ks], Bf[(ks >> 2) * C::BOX + swz[ks & 3]]);
            }
            *reinterpret_cast<double2*>(sY + (8 * mt + g) * C::YS + 8 * nt + 2 * q) = make_double2(c0, c1);
        }
        __syncthreads();                                // sY complete, grid buffer free
        if (BATCH && fn < f1) issue(fn);                // batch variant: separate fold buffers, the next field streams in now
        // Gaussian-weighted even/odd fold (legendre.f90:127-133)
        for (int t = tid; t < C::RG * C::IY; t += nthr) {
            const int cl = t / C::IY, jh = t - cl * C::IY;
            const double south = sY[cl * C::YS + jh], north = sY[cl * C::YS + (C::IL - 1 - jh)];
            const double wgt = wgt_t;
            sE[cl * C::ES + jh] = (north + south) * wgt;
            sO[cl * C::ES + jh] = (north - south) * wgt;
        }
        __syncthreads();
        // direct Legendre (legendre.f90:142-154): one thread per (m, two n of equal parity): the real and imaginary
        // sums share P, the two n share the folded Fourier coefficients (the stage is shared-memory-bandwidth-bound)
        double* out = out_base + (size_t)e * out_ms + (size_t)f * C::K2 * C::NX;
        constexpr int NP = 2 * ((C::NX + 3) / 4);
        for (int t = tid; t < NP * C::MG; t += nthr) {
            const int np = t / C::MG, ml = t - np * C::MG;
            const int m = grp * C::MG + ml;
            if (m >= C::MX) continue;
            const int nA = 4 * (np >> 1) + (np & 1), nB = nA + 2;
            const bool vA = nA < C::NX, vB = nB < C::NX;
            const bool cA = vA && nA <= TRUNC && m + nA <= C::MX, cB = vB && nB <= TRUNC && m + nB <= C::MX;
            double ar = 0.0, ai = 0.0, br = 0.0, bi = 0.0;
            if (cA) {                                   // m + nB <= MX implies m + nA <= MX: cB only if cA
                const double* Fr = ((np & 1) ? sO : sE) + (2 * ml) * C::ES;
                if (C::P_SMEM) {
                    const double* PA = sPd + (size_t)nA * C::MG + ml;
                    const double* PB = sPd + (size_t)(cB ? nB : nA) * C::MG + ml;
#pragma unroll
                    for (int jh = 0; jh < C::IY; jh++) {
                        const double fr = Fr[jh], fi = Fr[C::ES + jh];
                        const double pa = PA[(size_t)jh * C::NX * C::MG], pb = PB[(size_t)jh * C::NX * C::MG];
                        ar += pa * fr; ai += pa * fi; br += pb * fr; bi += pb * fi;
                    }
                } else {
                    const double* PA = tv.poly + (size_t)nA * C::MX + m;
                    const double* PB = tv.poly + (size_t)(cB ? nB : nA) * C::MX + m;
#pragma unroll
                    for (int jh = 0; jh < C::IY; jh++) {
                        const double fr = Fr[jh], fi = Fr[C::ES + jh];
                        const double pa = PA[(size_t)jh * C::NX * C::MX], pb = PB[(size_t)jh * C::NX * C::MX];
                        ar += pa * fr; ai += pa * fi; br += pb * fr; bi += pb * fi;
                    }
                }
                if (!cB) { br = 0.0; bi = 0.0; }
            }
            if (vA) *reinterpret_cast<double2*>(out + nA * C::K2 + 2 * m) = make_double2(ar, ai);
            if (vB) *reinterpret_cast<double2*>(out + nB * C::K2 + 2 * m) = make_double2(br, bi);
        }
        if (!BATCH && fn < f1) {                        // the folds live in the grid buffer: the next field may only come now
            __syncthreads();
            issue(fn);
        }
        f = fn;
    }
    if (tv.trace) { __syncthreads(); if (tid == 0) trace_end(tv.trace, 2); }
}

// ---------------------------------------------------------------------------------------------------------------------
// EXPERIMENTAL (SPEEDY_K2_FIELD=1, T30 ensemble batches only; written at the end of round 1 without GPU time left — it has NOT
// run on a device yet and is off by default): grid->spec with one CTA per WHOLE field instead of four wavenumber-group CTAs.
// The field is read once, the zonal transform is FFTPACK's forward FFT regrouped into two register-resident stages
// (fft96f.cuh, bit-identical to rfftf1 on the host) instead of the dense operator, and the direct Legendre sums of all
// wavenumbers read the triangle-packed P table staged once per CTA.  Estimated shared-memory wavefronts per field: ~3 000
// against ~10 000 for the four-CTA form.
//   sG  [NBOX][IL][16]  grid field, tensor-map boxes with the 128-byte swizzle (as k_g2s_stream)
//   sT  [IX][XP]        between the FFT stages, position-major; afterwards the even/odd folds EO[parity][jh][64]
//   sY  [IX][XP]        half-complex rows, position-major
//   sP  [IY][TR]        P table, per latitude the triangle m + n <= trunc + 1 packed row by row (polyt)
template <int TRUNC>
struct FieldCfg : SCfg<TRUNC> {
    using C = SCfg<TRUNC>;
    static constexpr int XP = C::IL + 1;                                   // odd row stride: conflict-free for consecutive rows
    static constexpr int THREADS = 8 * C::IL;                              // stage A: one thread per (row, k of radf4)
    static constexpr int EOW = 2 * ((C::MX + 1) / 2 * 2);                  // doubles per latitude of a fold row (re, im per m)
    static constexpr size_t SMEM = sizeof(double) * ((size_t)C::IL * C::IX + 2 * (size_t)C::IX * XP + (size_t)C::IY * C::TR + C::IX) + 2 * sizeof(uint64_t) + sizeof(int) * (C::NX + 2);
    static_assert(TRUNC == 30 && C::IX == 96, "whole-field kernel: T30 only (fft96f.cuh)");
    static_assert(2 * C::IY * EOW <= C::IX * XP && SMEM <= 232448 && ((size_t)C::IY * C::TR * 8) % 16 == 0, "fold buffer fits in sT; shared memory budget");
};

template <int TRUNC>
__global__ void __launch_bounds__(FieldCfg<TRUNC>::THREADS, 1)
k_g2s_field(const __grid_constant__ CUtensorMap gmap, const XDesc* __restrict__ desc, int nbatch, int nchunk,
            double* __restrict__ out_base, long long out_ms, DevTables tv, const int* __restrict__ gate) {
    using C = FieldCfg<TRUNC>;
    extern __shared__ __align__(1024) double smem[];
    double* sG = smem;
    double* sT = sG + C::IL * C::IX;
    double* sY = sT + C::IX * C::XP;
    double* sP = sY + C::IX * C::XP;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + (size_t)C::IY * C::TR);   // [0] P table, [1] grid field
    double* sEO = sT;                                                      // [2][IY][EOW], after the FFT
    // no static shared arrays here: the dynamic buffer must start on the 1 KB boundary the swizzle pattern is tied to
    double* sWa = reinterpret_cast<double*>(bars + 2);                     // [IX]
    int* sTri = reinterpret_cast<int*>(sWa + C::IX);                       // [NX + 1]
    const int tid = threadIdx.x, nthr = blockDim.x;
    if (tid == 0) trace_begin(tv.trace, 2);
    const int chunk = blockIdx.x, e = blockIdx.y;
    const int f0 = (int)((long long)chunk * nbatch / nchunk), f1 = (int)((long long)(chunk + 1) * nbatch / nchunk);
    int gate_open = 1;
    auto live = [&](int f) { return gate_open || !(desc[f].flags & 4); };
    auto next_live = [&](int f) { while (f < f1 && !live(f)) f++; return f; };
    auto issue = [&](int f) {
        if (tid == 0) {
            const long long off = desc[f].off;
            const int row0 = (int)(off / C::IX);
            if (off != (long long)row0 * C::IX) __trap();
            fence_proxy_async();
            mbar_expect_tx(&bars[1], C::IL * C::IX * sizeof(double));
#pragma unroll
            for (int b = 0; b < C::NBOX; b++) tensor_g2s_3d(sG + b * C::BOX, &gmap, 16 * b, row0, e, &bars[1]);
        }
    };
    if (tid == 0 && (smem_u32(sG) & 1023u)) __trap();
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    __syncthreads();
    // prologue on constant tables only
    if (tid == 0) {
        mbar_expect_tx(&bars[0], (uint32_t)((size_t)C::IY * C::TR * sizeof(double)));
        bulk_g2s(sP, tv.polyt, (uint32_t)((size_t)C::IY * C::TR * sizeof(double)), &bars[0]);
    }
    for (int t = tid; t < C::IX; t += nthr) sWa[t] = tv.fftwa[t];
    for (int t = tid; t <= C::NX; t += nthr) sTri[t] = C::tri_off(t);
    const double scale = (double)(1.0f / (float)C::IX);                    // fourier.f90:72
    // stage A mapping: k fastest (8 consecutive lanes read 8 consecutive longitudes of a row)
    const int rA = tid >> 3, k3 = tid & 7;
    const double sjA1 = tv.cosgr[rA], sjA2 = tv.cosgr2[rA];
    pdl_wait();
    pdl_trigger();
    gate_open = gate ? *gate : 1;
    int f = next_live(f0);
    if (f < f1) issue(f);
    __syncthreads();                                                       // sWa, sTri
    int it = 0;
    for (; f < f1; it++) {
        const XDesc dsc = desc[f];
        const int fn = next_live(f + 1);
        if (it == 0) mbar_wait(&bars[0], 0);
        mbar_wait(&bars[1], it & 1);
        // ---- forward FFT, stage A (radf3 + radf4): grid points k3 + 8 j + 32 jj of row rA, with the cosgr / cosgr2 pre-scale
        // of vdspec (spectral.f90:208-222).  Element (lat, lon): box lon / 16, 16-byte chunk ((lon % 16) / 2) ^ (lat % 8).
        {
            double x[12];
            const double sj = (dsc.flags & 1) ? sjA1 : ((dsc.flags & 2) ? sjA2 : 1.0);
            const bool scl = (dsc.flags & 3) != 0;
#pragma unroll
            for (int jj = 0; jj < 3; jj++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int lon = k3 + 8 * j + 32 * jj;                   // box 2 jj + j / 2, in-box longitude k3 + 8 (j % 2)
                    const int chunkc = ((k3 >> 1) + 4 * (j & 1)) ^ (rA & 7);
                    const double v = sG[(2 * jj + (j >> 1)) * C::BOX + rA * 16 + (chunkc << 1) + (lon & 1)];
                    x[4 * jj + j] = scl ? v * sj : v;
                }
            Fft96F::stageA<C::XP>(x, sT + rA, sWa, k3);
        }
        __syncthreads();                                                   // sT complete, grid buffer free
        if (fn < f1) issue(fn);
        // ---- stage B (radf4 + radf2): consecutive lanes = consecutive rows; sets 0..4 general (i = 3 + 2 set), 5 first, 6 last
        if (tid < 7 * C::IL) {
            const int set = tid / C::IL, r = tid - set * C::IL;
            if (set < 5) Fft96F::stageB_general<C::XP>(sT + r, sY + r, sWa, 3 + 2 * set);
            else if (set == 5) Fft96F::stageB_first<C::XP>(sT + r, sY + r, sWa);
            else Fft96F::stageB_last<C::XP>(sT + r, sY + r, sWa);
        }
        __syncthreads();                                                   // sY complete, sT free for the folds
        // ---- fourier_dir's 1/ix (fourier.f90:72-80) and the Gaussian-weighted even/odd fold (legendre.f90:127-133):
        // coefficient row c = 2 m (re), 2 m + 1 (im) <-> half-complex position 0 / 2 m - 1, 2 m; Im(m = 0) = 0
        for (int t = tid; t < C::K2 * C::IY; t += nthr) {
            const int jh = t / C::K2, c = t - jh * C::K2;                 // c fastest: the fold rows are written contiguously
            double ev = 0.0, od = 0.0;
            if (c != 1) {
                const int pos = (c == 0) ? 0 : c - 1;
                const double south = sY[pos * C::XP + jh] * scale, north = sY[pos * C::XP + (C::IL - 1 - jh)] * scale;
                const double wgt = tv.wt[jh];
                ev = (north + south) * wgt;
                od = (north - south) * wgt;
            }
            sEO[(size_t)jh * C::EOW + c] = ev;
            sEO[(size_t)(C::IY + jh) * C::EOW + c] = od;
        }
        __syncthreads();
        // ---- direct Legendre (legendre.f90:142-154) for all wavenumbers: one thread per (m, two n of equal parity)
        double* out = out_base + (size_t)e * out_ms + (size_t)f * C::K2 * C::NX;
        constexpr int NP = 2 * ((C::NX + 3) / 4);
        for (int t = tid; t < NP * C::MX; t += nthr) {
            const int np = t / C::MX, m = t - np * C::MX;
            const int nA = 4 * (np >> 1) + (np & 1), nB = nA + 2;
            const bool vA = nA < C::NX, vB = nB < C::NX;
            const bool cA = vA && nA <= TRUNC && m + nA <= C::MX, cB = vB && nB <= TRUNC && m + nB <= C::MX;
            double ar = 0.0, ai = 0.0, br = 0.0, bi = 0.0;
            if (cA) {
                const double2* F = reinterpret_cast<const double2*>(sEO + (size_t)((np & 1) ? C::IY : 0) * C::EOW) + m;
                const double* PA = sP + sTri[nA] + m;
                const double* PB = sP + sTri[cB ? nB : nA] + m;
#pragma unroll
                for (int jh = 0; jh < C::IY; jh++) {
                    const double2 fv = F[(size_t)jh * (C::EOW / 2)];
                    const double pa = PA[(size_t)jh * C::TR], pb = PB[(size_t)jh * C::TR];
                    ar += pa * fv.x; ai += pa * fv.y; br += pb * fv.x; bi += pb * fv.y;
                }
                if (!cB) { br = 0.0; bi = 0.0; }
            }
            if (vA) *reinterpret_cast<double2*>(out + nA * C::K2 + 2 * m) = make_double2(ar, ai);
            if (vB) *reinterpret_cast<double2*>(out + nB * C::K2 + 2 * m) = make_double2(br, bi);
        }
        __syncthreads();                                                   // the folds (in sT) are free for the next field
        f = fn;
    }
    if (tv.trace && tid == 0) trace_end(tv.trace, 2);
}

void launch_g2s_quad(speedy_ctx* ctx, const double* d_in, long long in_ms, const XDesc* d_desc, int nbatch, double* d_out, long long out_ms, int nmembers, const int* gate,
                     const G2sStepOpts& step);

void setup_transform_kernels() {
    CUDA_CHECK(cudaFuncSetAttribute(k_spec_to_grid<30>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TCfg<30>::K1_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(k_spec_to_grid<47>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TCfg<47>::K1_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(k_grid_to_spec<30>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TCfg<30>::K2_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(k_grid_to_spec<47>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TCfg<47>::K2_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(k_s2g_stream<30, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCfg<30>::K1_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(k_s2g_stream<47, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCfg<47>::K1_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(k_g2s_stream<30, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCfg<30>::K2_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(k_g2s_stream<47, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCfg<47>::K2_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(k_s2g_stream<30, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCfg<30>::K1_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(k_s2g_stream<47, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCfg<47>::K1_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(k_s2g_stream<30, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCfg<30>::K1_SMEM_FFT));
    CUDA_CHECK(cudaFuncSetAttribute(k_s2g_stream<30, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCfg<30>::K1_SMEM_FFT));
    CUDA_CHECK(cudaFuncSetAttribute(k_s2g_stream<47, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCfg<47>::K1_SMEM_FFT));
    CUDA_CHECK(cudaFuncSetAttribute(k_s2g_stream<47, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCfg<47>::K1_SMEM_FFT));
    CUDA_CHECK(cudaFuncSetAttribute(k_g2s_stream<30, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCfg<30>::K2_SMEM_BATCH));
    CUDA_CHECK(cudaFuncSetAttribute(k_g2s_stream<47, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCfg<47>::K2_SMEM_BATCH));
    CUDA_CHECK(cudaFuncSetAttribute(k_g2s_field<30>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FieldCfg<30>::SMEM));
}

// fields per persistent CTA: spread (slices x members x chunks) over the SMs, one CTA each
static int stream_chunks(speedy_ctx* ctx, int slices, int nmembers, int nbatch, int reserved_sms = 0, int ctas_per_sm = 1) {
    int nchunk = (ctx->num_sms * ctas_per_sm - reserved_sms) / (slices * nmembers);
    if (nchunk < 1) nchunk = 1;
    if (nchunk > nbatch) nchunk = nbatch;
    return nchunk;
}

void launch_s2g_quad(speedy_ctx* ctx, const double* d_in, long long in_ms, const XDesc* d_desc, int nbatch, double* d_out, long long out_ms, int nmembers,
                     const CloseArgs& cl);

// ensemble batches at T30: four fields at a time, FFT + DMMA Legendre (transforms_quad.cu); the list must hold derived fields as
// aligned pairs (quad_ok), and three or more fields per SM must be there to fill the quads
bool s2g_quad_selected(const speedy_ctx* ctx, int nbatch, int nmembers, bool quad_ok) {
    return ctx->d.trunc == 30 && ctx->precision == 0 && quad_ok && ctx->k1_quad && ctx->fft_inverse && (long long)nbatch * nmembers >= 3ll * ctx->num_sms;
}

template <int TRUNC>
static void launch_s2g_stream(speedy_ctx* ctx, const double* d_in, long long in_ms, const XDesc* d_desc, int nbatch,
                              double* d_out, long long out_ms, int nmembers, const CloseArgs& cl, bool quad_ok) {
    using C = SCfg<TRUNC>;
    if constexpr (TRUNC == 30) {
        if (s2g_quad_selected(ctx, nbatch, nmembers, quad_ok)) {
            launch_s2g_quad(ctx, d_in, in_ms, d_desc, nbatch, d_out, out_ms, nmembers, cl);
            return;
        }
    }
    // one SM is left to the closing CTA when a step is to be closed
    // the Fourier stage is the regrouped FFTPACK FFT (fft96.cuh / fft144.cuh) unless SPEEDY_DENSE_INVERSE asks for the dense operator on the FP64 tensor pipe
    constexpr bool HF = C::HAS_FFT;
    const bool fft = HF && ctx->fft_inverse;
    const size_t smem = fft ? C::K1_SMEM_FFT : C::K1_SMEM;
    int& occ = ctx->occ_k1[fft ? 1 : 0];   // resident CTAs per SM (2 at T30, 1 at T47), per variant; cached per context (= per device)
    if (!occ) {
        if (fft) CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_s2g_stream<TRUNC, false, HF>, C::K1_THREADS, smem));
        else CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_s2g_stream<TRUNC, false, false>, C::K1_THREADS, smem));
        if (occ < 1) occ = 1;
        if (occ > 2) occ = 2;
    }
    const int nchunk = stream_chunks(ctx, C::LG, nmembers, nbatch, cl.clk ? 1 : 0, occ);
    dim3 grid(nchunk * C::LG + (cl.clk ? 1 : 0), nmembers);
    const bool pdl = ctx->dv.trace == nullptr || ctx->trace_pdl;
    // chunks of >= 3 fields (ensemble batches) take the throughput-oriented variant, the single-member step the latency-oriented one
    const bool batch = (nbatch + nchunk - 1) / nchunk >= 3;
#define S2G_LAUNCH(B, F) CUDA_CHECK(launch_pdl(pdl, k_s2g_stream<TRUNC, B, F>, grid, dim3(C::K1_THREADS), smem, ctx->stream, d_in, in_ms, d_desc, nbatch, nchunk, d_out, out_ms, ctx->dv, cl))
    if (fft) { if (batch) S2G_LAUNCH(true, HF); else S2G_LAUNCH(false, HF); }
    else { if (batch) S2G_LAUNCH(true, false); else S2G_LAUNCH(false, false); }
#undef S2G_LAUNCH
}
// tensor map of K2's input: [longitude][row of IX doubles][member] over the caller's buffer, boxes of 16 longitudes x IL rows,
// SWIZZLE_128B.  A field at element offset `off` (a multiple of IX) starts at row off / IX.  Maps are cached per buffer.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static const CUtensorMap& grid_field_map(const double* d_in, long long in_ms, int nmembers, int ix, int il) {   // il: latitudes per box
    static std::map<std::tuple<const void*, long long, int, int, int>, CUtensorMap> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    const auto key = std::make_tuple((const void*)d_in, in_ms, nmembers, ix, il);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    static EncodeTiledFn enc = nullptr;
    if (!enc) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) throw std::runtime_error("cuTensorMapEncodeTiled is not available in this driver");
        enc = reinterpret_cast<EncodeTiledFn>(fn);
    }
    if ((reinterpret_cast<uintptr_t>(d_in) & 15) || (nmembers > 1 && (in_ms * sizeof(double)) % 16)) throw std::runtime_error("grid_to_spec: input buffer must be 16-byte aligned");
    // rows: everything a member's descriptors may address; the extent is a clipping bound, not an allocation size
    const cuuint64_t rows = (nmembers > 1 && in_ms > 0) ? (cuuint64_t)(in_ms / ix) : (cuuint64_t)1 << 22;
    const cuuint64_t dims[3] = {(cuuint64_t)ix, rows, (cuuint64_t)nmembers};
    const cuuint64_t strides[2] = {(cuuint64_t)ix * sizeof(double), (nmembers > 1 ? (cuuint64_t)in_ms : rows * (cuuint64_t)ix) * sizeof(double)};
    const cuuint32_t box[3] = {16u, (cuuint32_t)il, 1u};
    const cuuint32_t es[3] = {1u, 1u, 1u};
    CUtensorMap m;
    const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(d_in), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed for the grid_to_spec input (" + std::to_string((int)r) + ")");
    return cache.emplace(key, m).first->second;
}

// field chunks per member of the streaming grid->spec kernel (three or more fields per chunk: its batch variant, and the quad kernel at T30)
template <int TRUNC>
static int g2s_chunks(speedy_ctx* ctx, int nbatch, int nmembers) {
    using C = SCfg<TRUNC>;
    int &occ = ctx->occ_k2, &occ_b = ctx->occ_k2b;       // resident CTAs per SM of the latency / batch variant; cached per context (= per device)
    if (!occ) { CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_g2s_stream<TRUNC, false>, C::K2_THREADS, C::K2_SMEM)); if (occ < 1) occ = 1; if (occ > 2) occ = 2; }
    if (!occ_b) { CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_b, k_g2s_stream<TRUNC, true>, C::K2_THREADS, C::K2_SMEM_BATCH)); if (occ_b < 1) occ_b = 1; if (occ_b > 2) occ_b = 2; }
    int nchunk = stream_chunks(ctx, C::CG, nmembers, nbatch, 0, occ);
    if ((nbatch + nchunk - 1) / nchunk >= 3) nchunk = stream_chunks(ctx, C::CG, nmembers, nbatch, 0, occ_b);   // batch variant
    return nchunk;
}
bool g2s_quad_selected(speedy_ctx* ctx, int nbatch, int nmembers) {
    if (ctx->d.trunc != 30 || ctx->precision != 0 || !ctx->k2_quad || nbatch <= 0) return false;
    const int nchunk = g2s_chunks<30>(ctx, nbatch, nmembers);
    return (nbatch + nchunk - 1) / nchunk >= 3;
}

template <int TRUNC>
static void launch_g2s_stream(speedy_ctx* ctx, const double* d_in, long long in_ms, const XDesc* d_desc, int nbatch,
                              double* d_out, long long out_ms, int nmembers, const int* gate, const G2sStepOpts& step) {
    using C = SCfg<TRUNC>;
    const int nchunk = g2s_chunks<TRUNC>(ctx, nbatch, nmembers);
    const CUtensorMap& gmap = grid_field_map(d_in, in_ms, nmembers, C::IX, C::IL);
    if constexpr (TRUNC == 30) {
        if (ctx->k2_quad && (nbatch + nchunk - 1) / nchunk >= 3) {     // four fields at a time: FFT + DMMA Legendre (transforms_quad.cu)
            launch_g2s_quad(ctx, d_in, in_ms, d_desc, nbatch, d_out, out_ms, nmembers, gate, step);
            return;
        }
    }
    // several CTAs share a field in the other variants: an output over the field's own grid rows would race with their reads
    if (step.out_field_stride) throw std::runtime_error("grid->spec: an in-place output needs the quad kernel");
    if constexpr (TRUNC == 30) {
        // experimental whole-field kernel (see k_g2s_field): opt-in, ensemble batches only
        if (ctx->k2_field == 2 || (ctx->k2_field && (nbatch + nchunk - 1) / nchunk >= 3)) {
            using F = FieldCfg<TRUNC>;
            const int nch = stream_chunks(ctx, 1, nmembers, nbatch);
            CUDA_CHECK(launch_pdl(ctx->dv.trace == nullptr || ctx->trace_pdl, k_g2s_field<TRUNC>, dim3(nch, nmembers), dim3(F::THREADS), F::SMEM, ctx->stream, gmap, d_desc, nbatch, nch, d_out, out_ms, ctx->dv, gate));
            return;
        }
    }
    dim3 grid(nchunk * C::CG, nmembers);
    if ((nbatch + nchunk - 1) / nchunk >= 3)
        CUDA_CHECK(launch_pdl(ctx->dv.trace == nullptr || ctx->trace_pdl, k_g2s_stream<TRUNC, true>, grid, dim3(C::K2_THREADS), C::K2_SMEM_BATCH, ctx->stream, gmap, d_desc, nbatch, nchunk, d_out, out_ms, ctx->dv, gate));
    else
        CUDA_CHECK(launch_pdl(ctx->dv.trace == nullptr || ctx->trace_pdl, k_g2s_stream<TRUNC, false>, grid, dim3(C::K2_THREADS), C::K2_SMEM, ctx->stream, gmap, d_desc, nbatch, nchunk, d_out, out_ms, ctx->dv, gate));
}

// layout of the per-wavenumber-group P tiles of the streaming direct transform: [grp][jh][n][mloc]
int polyd_groups(int trunc) { return trunc == 30 ? SCfg<30>::CG : SCfg<47>::CG; }
int polyd_mg(int trunc) { return trunc == 30 ? SCfg<30>::MG : SCfg<47>::MG; }
int polyt_row(int trunc) { return trunc == 30 ? SCfg<30>::TR : SCfg<47>::TR; }

void launch_spec_to_grid(speedy_ctx* ctx, const double* d_in, long long in_ms, const XDesc* d_desc, int nbatch,
                         double* d_out, long long out_ms, int nmembers, int mode, const CloseArgs* close, bool quad_ok) {
    if (nbatch <= 0) return;
    if (mode == 0 && ctx->precision == 1) {
        const CloseArgs cl = close ? *close : CloseArgs{nullptr, nullptr, 0, 0, nullptr};
        launch_spec_to_grid_f32(ctx, d_in, in_ms, d_desc, nbatch, d_out, out_ms, nmembers, cl);
    } else if (mode == 0) {
        const CloseArgs cl = close ? *close : CloseArgs{nullptr, nullptr, 0, 0, nullptr};
        if (ctx->d.trunc == 30) launch_s2g_stream<30>(ctx, d_in, in_ms, d_desc, nbatch, d_out, out_ms, nmembers, cl, quad_ok);
        else launch_s2g_stream<47>(ctx, d_in, in_ms, d_desc, nbatch, d_out, out_ms, nmembers, cl, false);
    } else if (ctx->d.trunc == 30) {
        dim3 grid(nbatch * TCfg<30>::LG, nmembers);
        k_spec_to_grid<30><<<grid, TCfg<30>::K1_THREADS, TCfg<30>::K1_SMEM, ctx->stream>>>(d_in, in_ms, d_desc, d_out, out_ms, ctx->dv, mode);
    } else {
        dim3 grid(nbatch * TCfg<47>::LG, nmembers);
        k_spec_to_grid<47><<<grid, TCfg<47>::K1_THREADS, TCfg<47>::K1_SMEM, ctx->stream>>>(d_in, in_ms, d_desc, d_out, out_ms, ctx->dv, mode);
    }
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
}

void launch_grid_to_spec(speedy_ctx* ctx, const double* d_in, long long in_ms, const XDesc* d_desc, int nbatch,
                         double* d_out, long long out_ms, int nmembers, int mode, const int* gate, G2sStepOpts step) {
    if (nbatch <= 0) return;
    if (mode == 0 && ctx->precision == 1) {
        launch_grid_to_spec_f32(ctx, d_in, in_ms, d_desc, nbatch, d_out, out_ms, nmembers, gate);
    } else if (mode == 0) {
        if (ctx->d.trunc == 30) launch_g2s_stream<30>(ctx, d_in, in_ms, d_desc, nbatch, d_out, out_ms, nmembers, gate, step);
        else launch_g2s_stream<47>(ctx, d_in, in_ms, d_desc, nbatch, d_out, out_ms, nmembers, gate, step);
    } else if (ctx->d.trunc == 30) {
        dim3 grid(nbatch * TCfg<30>::CG, nmembers);
        k_grid_to_spec<30><<<grid, TCfg<30>::K2_THREADS, TCfg<30>::K2_SMEM, ctx->stream>>>(d_in, in_ms, d_desc, d_out, out_ms, ctx->dv, mode, gate);
    } else {
        dim3 grid(nbatch * TCfg<47>::CG, nmembers);
        k_grid_to_spec<47><<<grid, TCfg<47>::K2_THREADS, TCfg<47>::K2_SMEM, ctx->stream>>>(d_in, in_ms, d_desc, d_out, out_ms, ctx->dv, mode, gate);
    }
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace spd
