// Single-precision spherical-harmonic transforms: the `precision = 1` mode of BASELINE configs[4]
// ("mixed fp32 dynamical core + fp64 implicit solve").  The Legendre sums and the dense Fourier
// operator are evaluated entirely in real32 (operands converted on load, FFMA accumulation); the
// prognostic state, the grid-point column work, the semi-implicit solve and the time stepping stay
// fp64.  These kernels exist for the tolerance study (tools/precision_study.py, tools/configs4_ensemble.py):
// one CTA per (field, slice), no tensor cores; the P_n^m slice and the Fourier operator rows of a CTA are
// staged in shared memory from real32 copies of the tables (built once per context), so the inner loops
// touch shared memory only.  The sums run in the order of the first version of these kernels (operands
// converted on load): results are bit-identical to it.
//
//   k_s2g_f32 = uvspec/grad input stage + legendre_inv (legendre.f90:74-111) + fourier_inv (fourier.f90:23-53)
//   k_g2s_f32 = fourier_dir (fourier.f90:56-82) + legendre_dir (legendre.f90:114-155)
#include "ctx.h"
#include "spectral_ops.cuh"
#include "tma.cuh"
#include "close_step.cuh"

namespace spd {

template <int TRUNC>
struct FCfg {
    static constexpr int MX = TRUNC + 1, NX = TRUNC + 2;
    static constexpr int IX = (TRUNC == 30) ? 96 : 144, IY = IX / 4, IL = IX / 2;
    static constexpr int K2 = 2 * MX, KP = (K2 + 7) / 8 * 8;
    static constexpr int NSPEC2 = NX * K2, NSPEC2P = (NSPEC2 + 3) / 4 * 4;          // padded: the regions behind it are filled with 16-byte stores
    static constexpr int LG = (TRUNC == 30) ? 3 : 9, JG = IY / LG, NR = 2 * JG;     // K1: latitude pairs per CTA
    static constexpr int RG = 16, CG = KP / RG;                                     // K2: Fourier rows per CTA
    static constexpr int THREADS = 384;
    static constexpr int MG = RG / 2;                                               // K2: zonal wavenumbers per CTA
    static constexpr int GS = IX + 1;                                               // K2: padded row of the grid field (bank-conflict-free column reads)
    // real32 tables of a context, one buffer: P[iy][nx][mx] | inverse operator transposed [K2][ix] | forward operator [KP][ix] | P per wavenumber group [CG][iy][nx][MG]
    static constexpr size_t T_POLY = 0, T_FINV = T_POLY + (size_t)IY * NX * MX, T_FFWD = T_FINV + (size_t)K2 * IX, T_POLYD = T_FFWD + (size_t)KP * IX,
                            T_END = T_POLYD + (size_t)CG * IY * NX * MG;
    static constexpr size_t S2G_SMEM = sizeof(double) * 2 * NSPEC2 + sizeof(float) * (NSPEC2P + K2 * NR + JG * NX * MX + K2 * IX);
    static_assert((JG * NX * MX) % 4 == 0 && (K2 * IX) % 4 == 0 && (K2 * NR) % 4 == 0 && (IL * GS) % 4 == 0 && (RG * IY) % 4 == 0 && (IY * NX * MG) % 4 == 0 &&
                  T_FINV % 4 == 0 && T_FFWD % 4 == 0 && T_POLYD % 4 == 0 && (2 * NSPEC2 * sizeof(double)) % 16 == 0, "16-byte staging");
    static constexpr size_t G2S_SMEM = sizeof(float) * (IL * GS + RG * IL + 2 * RG * IY + RG * IX + IY * NX * MG);
};

template <int TRUNC>
__global__ void __launch_bounds__(FCfg<TRUNC>::THREADS)
k_s2g_f32(const double* __restrict__ in_base, long long in_ms, const XDesc* __restrict__ desc,
          double* __restrict__ out_base, long long out_ms, DevTables tv, const float* __restrict__ ftab, CloseArgs cl, int nwork) {
    using C = FCfg<TRUNC>;
    if (blockIdx.x == (unsigned)nwork) {                         // the extra CTAs: member 0's closes the previous main-loop step, off the critical path (close_step.cuh)
        if (blockIdx.y == 0 && cl.clk) { pdl_wait(); pdl_trigger(); close_step_cta(cl, threadIdx.x); }
        return;
    }
    extern __shared__ __align__(16) unsigned char raw[];
    double* sA = reinterpret_cast<double*>(raw);                 // source fields of a derived input (fp64 state)
    double* sB = sA + C::NSPEC2;
    float* sIn = reinterpret_cast<float*>(sB + C::NSPEC2);       // the field to transform, real32
    float* sX = sIn + C::NSPEC2P;                                // [K2][NR] Fourier coefficients of this CTA's rows
    float* sP = sX + C::K2 * C::NR;                              // [JG][NX][MX] P of this CTA's latitude pairs
    float* sF = sP + C::JG * C::NX * C::MX;                      // [K2][IX] backward Fourier operator, coefficient-major
    const int b = blockIdx.x / C::LG, grp = blockIdx.x - b * C::LG, e = blockIdx.y, tid = threadIdx.x, nthr = blockDim.x;
    {   // constant tables first: contiguous slices of the context's real32 tables
        const float4* gp = reinterpret_cast<const float4*>(ftab + C::T_POLY + (size_t)grp * C::JG * C::NX * C::MX);
        for (int t = tid; t < C::JG * C::NX * C::MX / 4; t += nthr) reinterpret_cast<float4*>(sP)[t] = gp[t];
        const float4* gf = reinterpret_cast<const float4*>(ftab + C::T_FINV);
        for (int t = tid; t < C::K2 * C::IX / 4; t += nthr) reinterpret_cast<float4*>(sF)[t] = gf[t];
    }
    pdl_wait();                                                  // the spectral fields of the previous kernel are complete
    pdl_trigger();
    const XDesc dsc = desc[b];
    const double* mbase = in_base + (size_t)e * in_ms;
    const double* in = mbase + dsc.off;
    const int j0 = grp * C::JG;
    auto row_lat = [&](int r) { return (r < C::JG) ? (j0 + r) : (C::IL - 1 - (j0 + (r - C::JG))); };
    if (dsc.op == 0) {
        for (int t = tid; t < C::NSPEC2; t += nthr) {
            const int n = t / C::K2, c = t - n * C::K2;
            sIn[t] = ((c >> 1) + n <= C::MX) ? (float)in[t] : 0.0f;
        }
    } else {
        const double* in2 = mbase + dsc.off2;
        for (int t = tid; t < C::NSPEC2; t += nthr) { sA[t] = in[t]; if (dsc.op <= 2) sB[t] = in2[t]; }
        __syncthreads();
        for (int t = tid; t < C::MX * C::NX; t += nthr) {        // uvspec / grad in real32 (spectral.f90:124-196)
            const int n = t / C::MX, m = t - n * C::MX;
            auto L = [&](const double* f, int mm, int nn, float& re, float& im) { re = (float)f[2 * (mm + C::MX * nn)]; im = (float)f[2 * (mm + C::MX * nn) + 1]; };
            float r0r, r0i, r1r, r1i;
            if (dsc.op <= 2) {
                const float dx = (float)tv.uvdx[t], dym = (float)tv.uvdym[t], dyp = (float)tv.uvdyp[t];
                float vr, vi, dr, di; L(sA, m, n, vr, vi); L(sB, m, n, dr, di);
                const float zpr = -(dx * vi), zpi = dx * vr, zcr = -(dx * di), zci = dx * dr;     // times_i
                if (n == 0) {
                    float ar, ai, br, bi; L(sA, m, 1, ar, ai); L(sB, m, 1, br, bi);
                    r0r = zcr - dyp * ar; r0i = zci - dyp * ai; r1r = zpr + dyp * br; r1i = zpi + dyp * bi;
                } else if (n == C::NX - 1) {
                    float ar, ai, br, bi; L(sA, m, TRUNC, ar, ai); L(sB, m, TRUNC, br, bi);
                    r0r = dym * ar; r0i = dym * ai; r1r = -(dym * br); r1i = -(dym * bi);
                } else {
                    float am_r, am_i, ap_r, ap_i, bm_r, bm_i, bp_r, bp_i;
                    L(sA, m, n - 1, am_r, am_i); L(sA, m, n + 1, ap_r, ap_i); L(sB, m, n - 1, bm_r, bm_i); L(sB, m, n + 1, bp_r, bp_i);
                    r1r = (-(dym * bm_r) + dyp * bp_r) + zpr; r1i = (-(dym * bm_i) + dyp * bp_i) + zpi;
                    r0r = (dym * am_r - dyp * ap_r) + zcr; r0i = (dym * am_i - dyp * ap_i) + zci;
                }
            } else {
                const float gx = (float)tv.gradx[m], gym = (float)tv.gradym[t], gyp = (float)tv.gradyp[t];
                float pr, pi; L(sA, m, n, pr, pi);
                r0r = -(gx * pi); r0i = gx * pr;
                if (n == 0) { float ar, ai; L(sA, m, 1, ar, ai); r1r = gyp * ar; r1i = gyp * ai; }
                else if (n == C::NX - 1) { float ar, ai; L(sA, m, TRUNC, ar, ai); r1r = -(gym * ar); r1i = -(gym * ai); }
                else { float ar, ai, br, bi; L(sA, m, n - 1, ar, ai); L(sA, m, n + 1, br, bi); r1r = -(gym * ar) + gyp * br; r1i = -(gym * ai) + gyp * bi; }
            }
            const bool first = dsc.op == 1 || dsc.op == 3;
            float rr = first ? r0r : r1r, ri = first ? r0i : r1i;
            if (m + n > C::MX) { rr = 0.0f; ri = 0.0f; }
            sIn[2 * (m + C::MX * n)] = rr; sIn[2 * (m + C::MX * n) + 1] = ri;
        }
    }
    __syncthreads();
    // inverse Legendre, real32 FFMA
    for (int t = tid; t < C::JG * C::K2; t += nthr) {
        const int jl = t / C::K2, c = t - jl * C::K2, m = c >> 1;
        const float* P = sP + (size_t)jl * C::NX * C::MX + m;
        float ev = 0.0f, od = 0.0f;
#pragma unroll 4
        for (int n = 0; n < C::NX; n += 2) ev = fmaf(sIn[n * C::K2 + c], P[n * C::MX], ev);
#pragma unroll 4
        for (int n = 1; n < C::NX; n += 2) od = fmaf(sIn[n * C::K2 + c], P[n * C::MX], od);
        sX[c * C::NR + jl] = ev - od;
        sX[c * C::NR + C::JG + jl] = ev + od;
    }
    __syncthreads();
    // dense backward Fourier operator, real32 FFMA: grid[i][r] = sum_c finv[i][c] * X[c][r]
    double* out = out_base + (size_t)e * out_ms + (size_t)(dsc.oslot1 ? dsc.oslot1 - 1 : b) * C::IX * C::IL;
    const bool sc = dsc.flags & 1, ad = dsc.flags & 2;
    for (int t = tid; t < C::IX * C::NR; t += nthr) {
        const int r = t / C::IX, i = t - r * C::IX;
        float s = 0.0f;
#pragma unroll 4
        for (int c = 0; c < C::K2; c++) s = fmaf(sF[c * C::IX + i], sX[c * C::NR + r], s);
        const int j = row_lat(r);
        if (sc) s *= (float)tv.cosgr[j];
        if (ad) s += (float)tv.coriol[j];
        out[(size_t)j * C::IX + i] = (double)s;
    }
}

template <int TRUNC>
__global__ void __launch_bounds__(FCfg<TRUNC>::THREADS)
k_g2s_f32(const double* __restrict__ in_base, long long in_ms, const XDesc* __restrict__ desc,
          double* __restrict__ out_base, long long out_ms, DevTables tv, const int* __restrict__ gate, const float* __restrict__ ftab) {
    using C = FCfg<TRUNC>;
    extern __shared__ __align__(16) unsigned char raw[];
    float* sG = reinterpret_cast<float*>(raw);       // [IL][GS] the grid field, rows padded by one
    float* sY = sG + C::IL * C::GS;                  // [RG][IL]
    float* sE = sY + C::RG * C::IL;                  // [IY][RG] even fold, latitude-major
    float* sO = sE + C::RG * C::IY;                  // [IY][RG] odd fold
    float* sA = sO + C::RG * C::IY;                  // [RG][IX] forward Fourier operator rows of this group
    float* sP = sA + C::RG * C::IX;                  // [IY][NX][MG] P of this group's wavenumbers
    const int b = blockIdx.x / C::CG, grp = blockIdx.x - b * C::CG, e = blockIdx.y, tid = threadIdx.x, nthr = blockDim.x;
    const XDesc dsc = desc[b];
    const double* in = in_base + (size_t)e * in_ms + dsc.off;
    const int c0row = grp * C::RG;
    {
        const float4* ga = reinterpret_cast<const float4*>(ftab + C::T_FFWD + (size_t)c0row * C::IX);
        for (int t = tid; t < C::RG * C::IX / 4; t += nthr) reinterpret_cast<float4*>(sA)[t] = ga[t];
        const float4* gp = reinterpret_cast<const float4*>(ftab + C::T_POLYD + (size_t)grp * C::IY * C::NX * C::MG);
        for (int t = tid; t < C::IY * C::NX * C::MG / 4; t += nthr) reinterpret_cast<float4*>(sP)[t] = gp[t];
    }
    pdl_wait();                                                  // the grid fields (and the clock flag behind `gate`) of the previous kernels are complete
    pdl_trigger();
    if (gate && (dsc.flags & 4) && !*gate) return;
    const double* scl = (dsc.flags & 1) ? tv.cosgr : ((dsc.flags & 2) ? tv.cosgr2 : nullptr);
    for (int t = tid; t < C::IL * C::IX; t += nthr) {
        const int j = t / C::IX, i = t - j * C::IX;
        float v = (float)in[t];
        if (scl) v *= (float)scl[j];
        sG[j * C::GS + i] = v;
    }
    __syncthreads();
    for (int t = tid; t < C::RG * C::IL; t += nthr) {        // forward Fourier operator rows of this group
        const int cl = t / C::IL, j = t - cl * C::IL, c = c0row + cl;
        float s = 0.0f;
        if (c < C::K2) {
            const float* A = sA + cl * C::IX;
            const float* G = sG + j * C::GS;
#pragma unroll 4
            for (int i = 0; i < C::IX; i++) s = fmaf(A[i], G[i], s);
        }
        sY[cl * C::IL + j] = s;
    }
    __syncthreads();
    for (int t = tid; t < C::RG * C::IY; t += nthr) {        // Gaussian-weighted even/odd fold (legendre.f90:127-133)
        const int cl = t / C::IY, jh = t - cl * C::IY;
        const float south = sY[cl * C::IL + jh], north = sY[cl * C::IL + (C::IL - 1 - jh)], wgt = (float)tv.wt[jh];
        sE[jh * C::RG + cl] = (north + south) * wgt;
        sO[jh * C::RG + cl] = (north - south) * wgt;
    }
    __syncthreads();
    double* out = out_base + (size_t)e * out_ms + (size_t)b * C::K2 * C::NX;
    for (int t = tid; t < C::NX * C::RG; t += nthr) {        // direct Legendre (legendre.f90:142-154)
        const int n = t / C::RG, cl = t - n * C::RG, c = c0row + cl;
        if (c >= C::K2) continue;
        const int m = c >> 1;
        float s = 0.0f;
        if (n <= TRUNC && m + n <= C::MX) {
            const float* P = sP + n * C::MG + (cl >> 1);
            const float* F = ((n & 1) ? sO : sE) + cl;
#pragma unroll 4
            for (int jh = 0; jh < C::IY; jh++) s = fmaf(P[jh * C::NX * C::MG], F[jh * C::RG], s);
        }
        out[n * C::K2 + c] = (double)s;
    }
}

// per device (speedy_create calls it after cudaSetDevice); every real32 kernel, whichever launcher runs first (at T47 the direct
// transform needs more than the default 48 KB and is the FIRST transform of speedy_model_init)
void setup_f32_kernels() {
    CUDA_CHECK(cudaFuncSetAttribute(k_s2g_f32<30>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FCfg<30>::S2G_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(k_s2g_f32<47>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FCfg<47>::S2G_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(k_g2s_f32<30>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FCfg<30>::G2S_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(k_g2s_f32<47>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FCfg<47>::G2S_SMEM));
}

// real32 copies of the tables, built on first use (one buffer per context, laid out as FCfg::T_*)
template <int TRUNC>
static const float* f32_tables(speedy_ctx* ctx) {
    using C = FCfg<TRUNC>;
    if (ctx->f32tab.n == C::T_END) return ctx->f32tab.p;
    const Tables& t = ctx->tab;
    std::vector<float> h(C::T_END, 0.0f);
    for (size_t k = 0; k < (size_t)C::IY * C::NX * C::MX; k++) h[C::T_POLY + k] = (float)t.poly[k];
    for (int c = 0; c < C::K2; c++)
        for (int i = 0; i < C::IX; i++) h[C::T_FINV + (size_t)c * C::IX + i] = (float)t.finv[(size_t)i * C::KP + c];
    for (size_t k = 0; k < (size_t)C::KP * C::IX; k++) h[C::T_FFWD + k] = (float)t.ffwd[k];
    for (int g = 0; g < C::CG; g++)
        for (int j = 0; j < C::IY; j++)
            for (int n = 0; n < C::NX; n++)
                for (int ml = 0; ml < C::MG; ml++) {
                    const int m = g * C::MG + ml;
                    if (m < C::MX) h[C::T_POLYD + (((size_t)g * C::IY + j) * C::NX + n) * C::MG + ml] = (float)t.poly[((size_t)j * C::NX + n) * C::MX + m];
                }
    ctx->f32tab.upload(h);
    return ctx->f32tab.p;
}

void launch_spec_to_grid_f32(speedy_ctx* ctx, const double* d_in, long long in_ms, const XDesc* d_desc, int nbatch,
                             double* d_out, long long out_ms, int nmembers, const CloseArgs& cl) {
    const bool pdl = ctx->dv.trace == nullptr || ctx->trace_pdl;
    if (ctx->d.trunc == 30) {
        const int nwork = nbatch * FCfg<30>::LG;
        CUDA_CHECK(launch_pdl(pdl, k_s2g_f32<30>, dim3(nwork + (cl.clk ? 1 : 0), nmembers), dim3(FCfg<30>::THREADS), FCfg<30>::S2G_SMEM, ctx->stream,
                              d_in, in_ms, d_desc, d_out, out_ms, ctx->dv, f32_tables<30>(ctx), cl, nwork));
    } else {
        const int nwork = nbatch * FCfg<47>::LG;
        CUDA_CHECK(launch_pdl(pdl, k_s2g_f32<47>, dim3(nwork + (cl.clk ? 1 : 0), nmembers), dim3(FCfg<47>::THREADS), FCfg<47>::S2G_SMEM, ctx->stream,
                              d_in, in_ms, d_desc, d_out, out_ms, ctx->dv, f32_tables<47>(ctx), cl, nwork));
    }
}

void launch_grid_to_spec_f32(speedy_ctx* ctx, const double* d_in, long long in_ms, const XDesc* d_desc, int nbatch,
                             double* d_out, long long out_ms, int nmembers, const int* gate) {
    const bool pdl = ctx->dv.trace == nullptr || ctx->trace_pdl;
    if (ctx->d.trunc == 30)
        CUDA_CHECK(launch_pdl(pdl, k_g2s_f32<30>, dim3(nbatch * FCfg<30>::CG, nmembers), dim3(FCfg<30>::THREADS), FCfg<30>::G2S_SMEM, ctx->stream,
                              d_in, in_ms, d_desc, d_out, out_ms, ctx->dv, gate, f32_tables<30>(ctx)));
    else
        CUDA_CHECK(launch_pdl(pdl, k_g2s_f32<47>, dim3(nbatch * FCfg<47>::CG, nmembers), dim3(FCfg<47>::THREADS), FCfg<47>::G2S_SMEM, ctx->stream,
                              d_in, in_ms, d_desc, d_out, out_ms, ctx->dv, gate, f32_tables<47>(ctx)));
}

}  // namespace spd
