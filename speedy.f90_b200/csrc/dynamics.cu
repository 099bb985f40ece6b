// K3 / K6 — spectral-space kernels of the time step.
//   k_spec_prologue : uvspec (both time levels), grad(ps), get_geopotential        (spectral.f90:124-196, geopotential.f90:33-57)
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

namespace spd {

#define KX 8

struct SpecArgs {
    double* base; long long stride;
    Layout L;
    DevTables tv;
    const LevelConsts* lc;
    DevClock* clk;
    int j1, j2;
    double dt;
    int flag;
};

__device__ __forceinline__ const double* sfield(const double* mb, long long off, int nsp, int f) { return mb + off + (size_t)f * nsp * 2; }
__device__ __forceinline__ double* sfield(double* mb, long long off, int nsp, int f) { return mb + off + (size_t)f * nsp * 2; }

// flag bit0: also refresh phi <- get_geopotential(t(:,:,:,1))
__global__ void k_spec_prologue(SpecArgs a) {
    const int mx = a.tv.mx, nx = a.tv.nx, nsp = mx * nx;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= KX * nsp) return;
    const int k = t / nsp, r = t - k * nsp;
    const int n = r / mx, m = r - n * mx;
    double* mb = a.base + (size_t)blockIdx.y * a.stride;
    const LevelConsts& lc = *a.lc;
    cd uc, vc;
    // dynamics set: time level j2 (tendencies.f90:99-101)
    dev_uvspec(a.tv, sfield(mb, a.L.vor, nsp, (a.j2 - 1) * KX + k), sfield(mb, a.L.div, nsp, (a.j2 - 1) * KX + k), m, n, uc, vc);
    st(sfield(mb, a.L.sprep, nsp, SP_U2 + k), mx, m, n, uc);
    st(sfield(mb, a.L.sprep, nsp, SP_V2 + k), mx, m, n, vc);
    // physics set: time level 1 (physics.f90:96-98)
    dev_uvspec(a.tv, sfield(mb, a.L.vor, nsp, k), sfield(mb, a.L.div, nsp, k), m, n, uc, vc);
    st(sfield(mb, a.L.sprep, nsp, SP_U1 + k), mx, m, n, uc);
    st(sfield(mb, a.L.sprep, nsp, SP_V1 + k), mx, m, n, vc);
    if (k == 0) {
        cd dx, dy;
        dev_grad(a.tv, sfield(mb, a.L.ps, nsp, a.j2 - 1), m, n, dx, dy);   // tendencies.f90:121
        st(sfield(mb, a.L.sprep, nsp, SP_PX), mx, m, n, dx);
        st(sfield(mb, a.L.sprep, nsp, SP_PY), mx, m, n, dy);
        if (a.flag & 1) {
            // get_geopotential(t(:,:,:,1), phis)  geopotential.f90:33-57
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
            for (int kk = 0; kk < KX; kk++) st(sfield(mb, a.L.phi, nsp, kk), mx, m, n, ph[kk]);
        }
    }
}

// flag bit0: stop after implicit_terms and store the tendencies (get_tendencies drop-in)
__global__ void __launch_bounds__(64) k_spec_step(SpecArgs a) {
    const int mx = a.tv.mx, nx = a.tv.nx, nsp = mx * nx;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nsp) return;
    const int n = r / mx, m = r - n * mx;
    double* mb = a.base + (size_t)blockIdx.y * a.stride;
    const LevelConsts& lc = *a.lc;
    const DevTables& tv = a.tv;
    const size_t q = r;
    const double el2 = tv.el2[q];
    const cd zero{0.0, 0.0};

    cd vordt[KX], divdt[KX], tdt[KX], trdt[KX], psdt;
    // ---- tendencies.f90:212-234: spectral assembly of the transformed grid-point tendencies
#pragma unroll
    for (int k = 0; k < KX; k++) {
        const int f = GO_PER * k;
        cd vo, di, dum;
        dev_vds(tv, sfield(mb, a.L.sout, nsp, f + 0), sfield(mb, a.L.sout, nsp, f + 1), m, n, vo, di);
        vordt[k] = vo;
        const cd ke = ld(sfield(mb, a.L.sout, nsp, f + 2), mx, m, n);
        divdt[k] = di - neg(el2 * ke);                                   // - laplacian(KE)
        dev_vds(tv, sfield(mb, a.L.sout, nsp, f + 3), sfield(mb, a.L.sout, nsp, f + 4), m, n, dum, di);
        tdt[k] = di + ld(sfield(mb, a.L.sout, nsp, f + 5), mx, m, n);
        dev_vds(tv, sfield(mb, a.L.sout, nsp, f + 6), sfield(mb, a.L.sout, nsp, f + 7), m, n, dum, di);
        trdt[k] = di + ld(sfield(mb, a.L.sout, nsp, f + 8), mx, m, n);
    }
    psdt = ld(sfield(mb, a.L.sout, nsp, GO_PSDT), mx, m, n);
    if (r == 0) psdt = zero;                                             // tendencies.f90:126

    // time level 1 of the prognostics (all linear terms use it, alph = 0.5: tendencies.f90:32)
    cd vor1[KX], div1[KX], t1[KX], tr1[KX];
#pragma unroll
    for (int k = 0; k < KX; k++) {
        vor1[k] = ld(sfield(mb, a.L.vor, nsp, k), mx, m, n);
        div1[k] = ld(sfield(mb, a.L.div, nsp, k), mx, m, n);
        t1[k] = ld(sfield(mb, a.L.t, nsp, k), mx, m, n);
        tr1[k] = ld(sfield(mb, a.L.tr, nsp, k), mx, m, n);
    }
    const cd ps1 = ld(sfield(mb, a.L.ps, nsp, 0), mx, m, n);

    // ---- get_spectral_tendencies  tendencies.f90:242-293
    {
        cd dmeanc = zero;
#pragma unroll
        for (int k = 0; k < KX; k++) dmeanc = dmeanc + lc.dhs[k] * div1[k];
        psdt = psdt - dmeanc;
        if (r == 0) psdt = zero;
        cd sigdtc[KX + 1], dumk[KX + 1];
        sigdtc[0] = zero; sigdtc[KX] = zero;
#pragma unroll
        for (int k = 0; k < KX - 1; k++) sigdtc[k + 1] = sigdtc[k] - lc.dhs[k] * (div1[k] - dmeanc);
        dumk[0] = zero; dumk[KX] = zero;
#pragma unroll
        for (int k = 1; k < KX; k++) dumk[k] = (lc.tref[k] - lc.tref[k - 1]) * sigdtc[k];
#pragma unroll
        for (int k = 0; k < KX; k++)
            tdt[k] = ((tdt[k] - lc.dhsr[k] * (dumk[k + 1] + dumk[k])) + lc.tref3[k] * (sigdtc[k + 1] + sigdtc[k])) - lc.tref2[k] * dmeanc;
        // phi was refreshed from t(:,:,:,1) by the prologue of this step (same values as :288)
#pragma unroll
        for (int k = 0; k < KX; k++) {
            const cd x = ld(sfield(mb, a.L.phi, nsp, k), mx, m, n) + (lc.rgas * lc.tref[k]) * ps1;
            divdt[k] = divdt[k] - neg(el2 * x);
        }
    }
    // ---- implicit_terms  implicit.f90:168-217
    {
        cd ye[KX], yf[KX];
#pragma unroll
        for (int k = 0; k < KX; k++) {
            cd s = zero;
#pragma unroll
            for (int k1 = 0; k1 < KX; k1++) s = s + tv.xd[k + KX * k1] * tdt[k1];
            ye[k] = s + lc.tref1[k] * psdt;
        }
        const double elz = tv.elz[q];
#pragma unroll
        for (int k = 0; k < KX; k++) yf[k] = divdt[k] + elz * ye[k];
#pragma unroll
        for (int k = 0; k < KX; k++) divdt[k] = zero;
        if (m + n != 0) {
            const double* xj = tv.xj + (size_t)KX * KX * (m + n - 1);   // xj(:,:,l), l = total wavenumber
            for (int k1 = 0; k1 < KX; k1++) {
#pragma unroll
                for (int k = 0; k < KX; k++) divdt[k] = divdt[k] + xj[k + KX * k1] * yf[k1];
            }
        }
#pragma unroll
        for (int k = 0; k < KX; k++) psdt = psdt - lc.dhsx[k] * divdt[k];
#pragma unroll
        for (int k = 0; k < KX; k++) {
#pragma unroll
            for (int k1 = 0; k1 < KX; k1++) tdt[k] = tdt[k] + tv.xc[k + KX * k1] * divdt[k1];
        }
    }
    if (a.flag & 1) {
#pragma unroll
        for (int k = 0; k < KX; k++) {
            st(sfield(mb, a.L.vordt, nsp, k), mx, m, n, vordt[k]);
            st(sfield(mb, a.L.divdt, nsp, k), mx, m, n, divdt[k]);
            st(sfield(mb, a.L.tdt, nsp, k), mx, m, n, tdt[k]);
            st(sfield(mb, a.L.trdt, nsp, k), mx, m, n, trdt[k]);
        }
        st(mb + a.L.psdt, mx, m, n, psdt);
        return;
    }
    // ---- horizontal diffusion + drag  time_stepping.f90:63-96
    {
        const double dmp = tv.dmp[q], dmpd = tv.dmpd[q], dmps = tv.dmps[q], dmp1 = tv.dmp1[q], dmp1d = tv.dmp1d[q], dmp1s = tv.dmp1s[q];
        const cd tcorh = ld(mb + a.L.tcorh, mx, m, n), qcorh = ld(mb + a.L.qcorh, mx, m, n);
#pragma unroll
        for (int k = 0; k < KX; k++) {
            vordt[k] = dmp1 * (vordt[k] - dmp * vor1[k]);
            divdt[k] = dmp1d * (divdt[k] - dmpd * div1[k]);
            const cd ctmp = t1[k] + lc.tcorv[k] * tcorh;
            tdt[k] = dmp1 * (tdt[k] - dmp * ctmp);
            if (k == 0 && m == 0) {
                vordt[0] = vordt[0] - lc.sdrag * vor1[0];
                divdt[0] = divdt[0] - lc.sdrag * div1[0];
            }
            vordt[k] = dmp1s * (vordt[k] - dmps * vor1[k]);
            divdt[k] = dmp1s * (divdt[k] - dmps * div1[k]);
            tdt[k] = dmp1s * (tdt[k] - dmps * ctmp);
            const cd qtmp = tr1[k] + lc.qcorv[k] * qcorh;
            trdt[k] = dmp1d * (trdt[k] - dmpd * qtmp);
        }
    }
    // ---- step_field_2d  time_stepping.f90:141-167
    {
        const double eps = (a.j1 == 1) ? 0.0 : lc.rob;
        const double trf = tv.trfilt[q];
        const double c1 = lc.wil * eps, c2 = (1.0 - lc.wil) * eps;
        auto stepf = [&](long long off, int nlev_fields, int k, cd fdt) {
            double* p1 = sfield(mb, off, nsp, k);
            double* p2 = sfield(mb, off, nsp, nlev_fields + k);
            fdt = trf * fdt;
            const cd f1 = ld(p1, mx, m, n);
            const cd fj = (a.j1 == 1) ? f1 : ld(p2, mx, m, n);
            const cd fnew = f1 + a.dt * fdt;
            const cd f1n = fj + c1 * ((f1 - 2.0 * fj) + fnew);
            const cd fj2 = (a.j1 == 1) ? f1n : fj;     // :166 re-reads output(:,:,j1) after :163 overwrote level 1
            const cd f2n = fnew - c2 * ((f1n - 2.0 * fj2) + fnew);
            st(p1, mx, m, n, f1n);
            st(p2, mx, m, n, f2n);
        };
        stepf(a.L.ps, 1, 0, psdt);
#pragma unroll
        for (int k = 0; k < KX; k++) {
            stepf(a.L.vor, KX, k, vordt[k]);
            stepf(a.L.div, KX, k, divdt[k]);
            stepf(a.L.t, KX, k, tdt[k]);
            stepf(a.L.tr, KX, k, trdt[k]);
        }
    }
}

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

// sum and sum of squares over this context's members of the 41 level-1 grid fields
__global__ void k_ensemble_sums(const double* __restrict__ base, long long stride, long long gin_off, int nmembers, int nvals,
                                double* __restrict__ sum, double* __restrict__ sumsq) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nvals) return;
    double s = 0.0, s2 = 0.0;
    for (int e = 0; e < nmembers; e++) {
        const double v = base[(size_t)e * stride + gin_off + q];
        s += v; s2 += v * v;
    }
    sum[q] = s; sumsq[q] = s2;
}

// SPPT AR(1) update in spectral space (sppt.f90:74-91); eta is supplied (caller noise) or
// drawn with a counter-based generator: flag bit0 = first step, bit1 = draw eta on device
struct SpptArgs {
    double* base; long long stride; Layout L; DevTables tv;
    unsigned long long seed; long long counter; int first; int draw; double rearth;
};
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ double u01(unsigned long long h) { return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
__global__ void k_sppt_update(SpptArgs a) {
    const int mx = a.tv.mx, nx = a.tv.nx, nsp = mx * nx;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= KX * nsp) return;
    const int k = t / nsp, r = t - k * nsp;
    const int n = r / mx, m = r - n * mx;
    const int e = blockIdx.y;
    double* mb = a.base + (size_t)e * a.stride;
    const double time_decorr = 6.0, len_decorr = 500000.0, stddev = (double)0.33f;
    const double phi = exp(-(24 / 36.0) / time_decorr);
    double f0 = 0.0;
    const double rr = len_decorr / a.rearth;
    for (int nn = 1; nn <= a.tv.trunc; nn++) f0 = f0 + (2 * nn + 1) * exp(-0.5 * (rr * rr) * nn * (nn + 1));
    f0 = sqrt(((stddev * stddev) * (1 - phi * phi)) / (2 * f0));
    const double sigma = f0 * exp(-0.25 * (len_decorr * len_decorr) * a.tv.el2[r]);
    cd eta;
    if (a.draw) {
        // counter-based Box-Muller (sppt.f90:103-117 shape: u = sqrt(-2 ln r1), v = 2*2pi*r2, sin only), clipped to +-10
        const unsigned long long id = ((unsigned long long)a.counter * 64ull + (unsigned long long)e) * (unsigned long long)(KX * nsp) + (unsigned long long)t;
        const unsigned long long h = splitmix64(a.seed ^ splitmix64(id));
        const double r1 = u01(splitmix64(h + 1)), r2 = u01(splitmix64(h + 2)), r3 = u01(splitmix64(h + 3)), r4 = u01(splitmix64(h + 4));
        const double c = (double)(2.0f * 6.28318530718f);
        double gr = sqrt(-2.0 * log(r1)) * sin(c * r2), gi = sqrt(-2.0 * log(r3)) * sin(c * r4);
        gr = fmin(10.0, fabs(gr)) * copysign(1.0, gr);
        gi = fmin(10.0, fabs(gi)) * copysign(1.0, gi);
        eta = cd{gr, gi};
        st(sfield(mb, a.L.sppt_eta, nsp, k), mx, m, n, eta);
    } else {
        eta = ld(sfield(mb, a.L.sppt_eta, nsp, k), mx, m, n);
    }
    double* sp = sfield(mb, a.L.sppt_spec, nsp, k);
    cd v;
    if (a.first) {
        const double c = pow(1 - phi * phi, -0.5);
        v = (c * sigma) * eta;
    } else {
        v = phi * ld(sp, mx, m, n) + sigma * eta;
    }
    st(sp, mx, m, n, v);
}

// ---- launchers -------------------------------------------------------------------------------
static SpecArgs spec_args(speedy_ctx* ctx) {
    Model& M = *ctx->model;
    SpecArgs a;
    a.base = M.mem.p; a.stride = M.L.stride; a.L = M.L; a.tv = ctx->dv; a.lc = M.lc.p; a.clk = M.clock.p;
    a.j1 = 1; a.j2 = 1; a.dt = 0.0; a.flag = 0;
    return a;
}

void launch_spec_prologue(speedy_ctx* ctx, int j2, int refresh_phi) {
    SpecArgs a = spec_args(ctx);
    a.j2 = j2; a.flag = refresh_phi ? 1 : 0;
    const int total = KX * ctx->d.nspec();
    dim3 grid((total + 127) / 128, ctx->nmembers);
    k_spec_prologue<<<grid, 128, 0, ctx->stream>>>(a);
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
}

void launch_spec_step(speedy_ctx* ctx, int j1, int j2, double dt, int store_tend_only) {
    SpecArgs a = spec_args(ctx);
    a.j1 = j1; a.j2 = j2; a.dt = dt; a.flag = store_tend_only ? 1 : 0;
    dim3 grid((ctx->d.nspec() + 63) / 64, ctx->nmembers);
    k_spec_step<<<grid, 64, 0, ctx->stream>>>(a);
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

void launch_sppt_update(speedy_ctx* ctx) {
    Model& M = *ctx->model;
    SpptArgs a;
    a.base = M.mem.p; a.stride = M.L.stride; a.L = M.L; a.tv = ctx->dv;
    a.seed = ctx->seed; a.counter = M.sppt_counter; a.first = M.sppt_first ? 1 : 0; a.draw = M.sppt_draw ? 1 : 0; a.rearth = ctx->tab.c.rearth;
    const int total = KX * ctx->d.nspec();
    dim3 grid((total + 127) / 128, ctx->nmembers);
    k_sppt_update<<<grid, 128, 0, ctx->stream>>>(a);
    M.sppt_first = false;
    M.sppt_counter++;
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
}

void launch_ensemble_sums(speedy_ctx* ctx, double* d_sum, double* d_sumsq) {
    Model& M = *ctx->model;
    const int nvals = 41 * ctx->d.ngrid();
    k_ensemble_sums<<<(nvals + 255) / 256, 256, 0, ctx->stream>>>(M.mem.p, M.L.stride, M.L.gin + (long long)GI_U1 * ctx->d.ngrid(), ctx->nmembers, nvals, d_sum, d_sumsq);
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace spd
