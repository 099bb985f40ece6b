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

void setup_transform_kernels() {
    CUDA_CHECK(cudaFuncSetAttribute(k_spec_to_grid<30>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TCfg<30>::K1_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(k_spec_to_grid<47>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TCfg<47>::K1_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(k_grid_to_spec<30>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TCfg<30>::K2_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(k_grid_to_spec<47>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TCfg<47>::K2_SMEM));
}

void launch_spec_to_grid(speedy_ctx* ctx, const double* d_in, long long in_ms, const XDesc* d_desc, int nbatch,
                         double* d_out, long long out_ms, int nmembers, int mode) {
    if (nbatch <= 0) return;
    if (ctx->d.trunc == 30) {
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
                         double* d_out, long long out_ms, int nmembers, int mode, const int* gate) {
    if (nbatch <= 0) return;
    if (ctx->d.trunc == 30) {
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
