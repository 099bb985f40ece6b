// K1 / K2 — batched spherical-harmonic transforms for sm_100a.
//
//   K1 spec_to_grid  = legendre_inv (legendre.f90:74-111) + fourier_inv (fourier.f90:23-53)
//   K2 grid_to_spec  = fourier_dir (fourier.f90:56-82)    + legendre_dir (legendre.f90:114-155)
//
// One CTA per (transform, member).  The Legendre contraction runs on the FP64 CUDA cores
// with the P_n^m table streamed from L2; the zonal Fourier step is applied as a dense
// real operator (the reference's FFTPACK transform including its single-precision
// constants, SURVEY.md F13, extracted on the host by tables.cpp) with FP64 tensor-core
// MMAs (mma.sync.m8n8k4.f64 — tcgen05 has no FP64 kind).  Intermediate Fourier
// coefficients never leave shared memory.
#include "ctx.h"

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
    static constexpr int K2 = 2 * MX, KP = (K2 + 3) / 4 * 4;
    static constexpr int XS = padmod16(IL, 4);   // sX row stride: conflict-free B fragments
    static constexpr int GS = padmod16(IX, 4);   // sG row stride
    static constexpr int YS = padmod16(IL, 8);   // sY row stride: conflict-free double2 C stores
    static constexpr int ES = IY + 1;            // even/odd fold stride (odd)
    static constexpr int K1_THREADS = IX / 8 * 32;
    static constexpr int K2_THREADS = KP / 8 * 32;
    static constexpr size_t K1_SMEM = sizeof(double) * (NX * K2 + KP * XS);
    static constexpr size_t K2_SMEM = sizeof(double) * (IL * GS + KP * YS);
    static_assert(KP % 8 == 0 && IX % 8 == 0 && IL % 8 == 0, "tile sizes");
    static_assert(2 * K2 * ES <= IL * GS, "fold buffers must fit in the grid staging buffer");
};

template <int TRUNC>
__global__ void __launch_bounds__(TCfg<TRUNC>::K1_THREADS)
k_spec_to_grid(const double* __restrict__ in_base, long long in_ms, const XDesc* __restrict__ desc,
               double* __restrict__ out_base, long long out_ms, DevTables tv, int mode) {
    using C = TCfg<TRUNC>;
    extern __shared__ double smem[];
    double* sIn = smem;
    double* sX = smem + C::NX * C::K2;
    const int b = blockIdx.x, e = blockIdx.y, tid = threadIdx.x, nthr = blockDim.x;
    const XDesc dsc = desc[b];
    const double* in = in_base + (size_t)e * in_ms + dsc.off;

    for (int t = tid; t < (C::KP - C::K2) * C::XS; t += nthr) sX[C::K2 * C::XS + t] = 0.0;
    if (mode != 2) {
        // coefficients outside the triangle m+n <= trunc+1 are never read by the reference
        // (legendre.f90:38 nsh2); they are zeroed here so that the sums below have a fixed trip
        // count and every P load of a sum is in flight at once
        for (int t = tid; t < C::NX * C::K2; t += nthr) {
            const int n = t / C::K2, c = t - n * C::K2;
            sIn[t] = ((c >> 1) + n <= C::MX) ? in[t] : 0.0;
        }
        __syncthreads();
        // inverse Legendre: even/odd split in n, hemispheric symmetry
        for (int t = tid; t < C::IY * C::KP; t += nthr) {
            const int jh = t / C::KP, c = t - jh * C::KP;
            if (c < C::K2) {
                const int m = c >> 1;
                const double* P = tv.poly + (size_t)jh * C::NX * C::MX + m;
                double ev = 0.0, od = 0.0;
#pragma unroll
                for (int n = 0; n < C::NX; n += 2) ev += sIn[n * C::K2 + c] * P[n * C::MX];
#pragma unroll
                for (int n = 1; n < C::NX; n += 2) od += sIn[n * C::K2 + c] * P[n * C::MX];
                sX[c * C::XS + jh] = ev - od;                  // j = jh (southern row)
                sX[c * C::XS + (C::IL - 1 - jh)] = ev + od;    // j = il+1-j (northern row)
            }
        }
    } else {
        for (int t = tid; t < C::IL * C::K2; t += nthr) {
            const int j = t / C::K2, c = t - j * C::K2;
            sX[c * C::XS + j] = in[t];
        }
    }
    __syncthreads();
    if (mode == 1) {
        double* out = out_base + (size_t)e * out_ms + (size_t)b * C::K2 * C::IL;
        for (int t = tid; t < C::IL * C::K2; t += nthr) {
            const int j = t / C::K2, c = t - j * C::K2;
            out[t] = sX[c * C::XS + j];
        }
        return;
    }
    // dense backward Fourier operator on the FP64 tensor pipe:
    //   grid[i][j] = sum_c finv[i][c] * X[c][j],  M = IX, N = IL, K = KP
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
#pragma unroll 2
    for (int nt = 0; nt < C::IL / 8; nt++) {
        double c0 = 0.0, c1 = 0.0;
        const double* B = sX + q * C::XS + 8 * nt + g;
#pragma unroll
        for (int ks = 0; ks < C::KP / 4; ks++) dmma884(c0, c1, a[ks], B[4 * ks * C::XS]);
        const int j0 = 8 * nt + 2 * q;
        if (sc) { c0 *= tv.cosgr[j0]; c1 *= tv.cosgr[j0 + 1]; }
        if (ad) { c0 += tv.coriol[j0]; c1 += tv.coriol[j0 + 1]; }
        out[(size_t)j0 * C::IX + i] = c0;
        out[(size_t)(j0 + 1) * C::IX + i] = c1;
    }
}

template <int TRUNC>
__global__ void __launch_bounds__(TCfg<TRUNC>::K2_THREADS)
k_grid_to_spec(const double* __restrict__ in_base, long long in_ms, const XDesc* __restrict__ desc,
               double* __restrict__ out_base, long long out_ms, DevTables tv, int mode, const int* __restrict__ gate) {
    using C = TCfg<TRUNC>;
    extern __shared__ double smem[];
    double* sG = smem;
    double* sY = smem + C::IL * C::GS;
    const int b = blockIdx.x, e = blockIdx.y, tid = threadIdx.x, nthr = blockDim.x;
    const XDesc dsc = desc[b];
    if (gate && (dsc.flags & 4) && !*gate) return;     // in-graph conditional work (the daily forcing transform)
    const double* in = in_base + (size_t)e * in_ms + dsc.off;
    const int w = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;

    if (mode != 2) {
        const double* scl = (dsc.flags & 1) ? tv.cosgr : ((dsc.flags & 2) ? tv.cosgr2 : nullptr);
        for (int t = tid; t < C::IL * C::IX; t += nthr) {
            const int j = t / C::IX, i = t - j * C::IX;
            double v = in[t];
            if (scl) v *= scl[j];
            sG[j * C::GS + i] = v;
        }
        __syncthreads();
        // dense forward Fourier operator: Y[c][j] = sum_i ffwd[c][i] * g[i][j],  M = KP, N = IL, K = IX
        double a[C::IX / 4];
        {
            const double* A = tv.ffwd + (size_t)(8 * w + g) * C::IX + q;
#pragma unroll
            for (int ks = 0; ks < C::IX / 4; ks++) a[ks] = A[4 * ks];
        }
#pragma unroll 2
        for (int nt = 0; nt < C::IL / 8; nt++) {
            double c0 = 0.0, c1 = 0.0;
            const double* B = sG + (8 * nt + g) * C::GS + q;
#pragma unroll
            for (int ks = 0; ks < C::IX / 4; ks++) dmma884(c0, c1, a[ks], B[4 * ks]);
            *reinterpret_cast<double2*>(sY + (8 * w + g) * C::YS + 8 * nt + 2 * q) = make_double2(c0, c1);
        }
    } else {
        for (int t = tid; t < C::IL * C::K2; t += nthr) {
            const int j = t / C::K2, c = t - j * C::K2;
            sY[c * C::YS + j] = in[t];
        }
    }
    __syncthreads();
    if (mode == 1) {
        double* out = out_base + (size_t)e * out_ms + (size_t)b * C::K2 * C::IL;
        for (int t = tid; t < C::IL * C::K2; t += nthr) {
            const int j = t / C::K2, c = t - j * C::K2;
            out[t] = sY[c * C::YS + j];
        }
        return;
    }
    // Gaussian-weighted even/odd fold (legendre.f90:127-133); sG is free now
    double* sE = sG;
    double* sO = sG + C::K2 * C::ES;
    for (int t = tid; t < C::K2 * C::IY; t += nthr) {
        const int c = t / C::IY, jh = t - c * C::IY;
        const double south = sY[c * C::YS + jh], north = sY[c * C::YS + (C::IL - 1 - jh)];
        const double wgt = tv.wt[jh];
        sE[c * C::ES + jh] = (north + south) * wgt;
        sO[c * C::ES + jh] = (north - south) * wgt;
    }
    __syncthreads();
    // direct Legendre: out(c,n) = sum_j P(m,n,j) * {even|odd}(c,j), n <= trunc (legendre.f90:142-154)
    double* out = out_base + (size_t)e * out_ms + (size_t)b * C::K2 * C::NX;
    for (int t = tid; t < C::NX * C::KP; t += nthr) {
        const int n = t / C::KP, c = t - n * C::KP;
        if (c >= C::K2) continue;
        const int m = c >> 1;
        double s = 0.0;
        if (n <= TRUNC && m + n <= C::MX) {
            const double* P = tv.poly + (size_t)n * C::MX + m;
            const double* F = ((n & 1) ? sO : sE) + c * C::ES;
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
    dim3 grid(nbatch, nmembers);
    if (ctx->d.trunc == 30)
        k_spec_to_grid<30><<<grid, TCfg<30>::K1_THREADS, TCfg<30>::K1_SMEM, ctx->stream>>>(d_in, in_ms, d_desc, d_out, out_ms, ctx->dv, mode);
    else
        k_spec_to_grid<47><<<grid, TCfg<47>::K1_THREADS, TCfg<47>::K1_SMEM, ctx->stream>>>(d_in, in_ms, d_desc, d_out, out_ms, ctx->dv, mode);
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
}

void launch_grid_to_spec(speedy_ctx* ctx, const double* d_in, long long in_ms, const XDesc* d_desc, int nbatch,
                         double* d_out, long long out_ms, int nmembers, int mode, const int* gate) {
    if (nbatch <= 0) return;
    dim3 grid(nbatch, nmembers);
    if (ctx->d.trunc == 30)
        k_grid_to_spec<30><<<grid, TCfg<30>::K2_THREADS, TCfg<30>::K2_SMEM, ctx->stream>>>(d_in, in_ms, d_desc, d_out, out_ms, ctx->dv, mode, gate);
    else
        k_grid_to_spec<47><<<grid, TCfg<47>::K2_THREADS, TCfg<47>::K2_SMEM, ctx->stream>>>(d_in, in_ms, d_desc, d_out, out_ms, ctx->dv, mode, gate);
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace spd
