// Internal context of the speedy_b200 library (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include <map>
#include <stdexcept>
#include "host/tables.h"

namespace spd {

// Device-side view of the constant tables (passed to kernels by value; < 1 KB)
struct DevTables {
    int trunc, ix, iy, il, kx, nx, mx;
    const double* poly;     // [iy][nx][mx]
    const double* polyt;    // [iy][TR]: per latitude the triangle m+n <= trunc+1 of P, rows of n packed (streaming inverse transform)
    const double* polyd;    // [grp][iy][nx][mg]: P re-laid out per group of mg zonal wavenumbers (streaming direct transform)
    const double* finv;     // [ix][k2pad]
    const double* ffwd;     // [k2pad][ix]
    const double* fftwa;    // [ix] twiddle table of rffti1 (fftpack.f90:1-67)
    const double* wt;       // iy
    const double* cosgr;    // il
    const double* cosgr2;   // il
    const double* coriol;   // il
    const double* cosg;     // il
    const double* sia;      // il
    const double* coa;      // il
    // spectral operator tables (mx,nx) m fastest
    const double *el2, *elm2, *trfilt, *gradx, *gradym, *gradyp, *uvdx, *uvdym, *uvdyp, *vddym, *vddyp;
    // dynamics
    const double *dmp, *dmpd, *dmps, *dmp1, *dmp1d, *dmp1s, *elz;
    const double *xj, *xc, *xd;          // Fortran order (kx,kx[,l])
    const double* xjt;                   // xj transposed: [k + kx*k1][l]
    const double* fband;                 // (301,4)
    const double* polyq;                 // T30: P fragments of the quad kernels' DMMA tiles, (warp, tile slot, k-step, lane) order (transforms_quad.cu)
    const int* qtile;                    // T30: (m, parity, n-tile) of every tile of the direct quad transform
    const double* polyi;                 // T30: P fragments of the inverse quad transform, (warp, tile slot, fragment, lane)
    const int* qtile_inv;                // T30: (m, band, even / odd k-steps) per (warp, tile slot) of the inverse quad transform
    unsigned long long* trace;           // nullptr, or the in-graph timeline buffer (speedy_trace)
};

// small per-level constants go to __constant__ memory (see consts.cuh)
struct LevelConsts {
    double hsg[9], dhs[8], fsg[8], dhsr[8], fsgr[8];
    double tref[8], tref1[8], tref2[8], tref3[8], dhsx[8];
    double xgeop1[8], xgeop2[8], geop_corf[8];
    double tcorv[8], qcorv[8];
    double sigl[8], sigh[9], grdsig[8], grdscp[8], wvi[16];
    double rgas, akap, cp, p0, grav, alhc, alhs, sbc, rearth, refrh1, gamma;
    double rob, wil, sdrag;
    // loop-invariant quotients of the column-serial sweeps, evaluated once on the host with the reference's expressions
    // (an fp64 division is ~130 cycles of dependent latency on the critical warp): convection.f90:118-131, LSC :52, SW :227
    double entr[8], ralhc, fm0, rdps, prg, tfact, eps1, rcp, pad_;
};

// descriptor of one transform in a batch
struct XDesc {
    long long off;   // element (double) offset of the field relative to the member base pointer
    int flags;       // K1: bit0 scale by cosgr(j), bit1 add coriol(j);  K2: bit0 scale by cosgr, bit1 scale by cosgr2, bit2 gated
    int op;          // K1 input: 0 the field at `off`; 1 ucos, 2 vcos of uvspec(vor@off, div@off2); 3 d/dx, 4 d/dy of grad(ps@off)
    long long off2;
    int oslot1;      // K1: 0 -> the transform's own index is its output slot; else output slot + 1 (compact field lists)
};

#define CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw std::runtime_error(std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

template <class T>
struct DevBuf {
    T* p = nullptr; size_t n = 0;
    // the memset runs on the legacy default stream, which does not order against the context's non-blocking stream: wait for it
    void alloc(size_t cnt) { free(); n = cnt; if (cnt) { CUDA_CHECK(cudaMalloc(&p, cnt * sizeof(T))); CUDA_CHECK(cudaMemset(p, 0, cnt * sizeof(T))); CUDA_CHECK(cudaDeviceSynchronize()); } }
    void upload(const std::vector<T>& v) { if (v.size() != n) alloc(v.size()); if (n) CUDA_CHECK(cudaMemcpy(p, v.data(), n * sizeof(T), cudaMemcpyHostToDevice)); }
    void free() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DevBuf() { free(); }
};

struct Model;   // model.h

}  // namespace spd

struct speedy_ctx {
    spd::Tables tab;
    spd::Dims d;
    int nmembers = 1;
    int device = 0;
    int sppt_on = 0;
    unsigned long long seed = 0;
    bool trace_pdl = false;  // SPEEDY_TRACE_PDL=1: keep programmatic dependent launch on while tracing (stamps under production overlap; the kernel timeline is then not meaningful)
    bool fft_inverse = true; // spec->grid Fourier stage: regrouped FFTPACK FFT (fft96.cuh / fft144.cuh); SPEEDY_DENSE_INVERSE=1 selects the dense DMMA operator
    bool member_ready = true; // main-loop step on the quad transforms: the column tiles of a member start when that member's grid fields are stored (member_ready.cuh); 0: when the whole transform is
    bool sppt_fold = true;             // SPPT with device-drawn noise: the spectral step prepares the next step's pattern in its prologue (no separate kernel in the step)
    bool transient_alias = true;       // ensemble main-loop step on the quad transforms: one buffer per member carries the step's transient fields — the column
                                       // kernel writes its grid tendencies over the grid fields it has staged, grid->spec writes each field's coefficients over
                                       // that field's own grid rows — so a step allocates a third of the L2 lines ("one transient buffer")
    bool l2_discard = true;  // ensemble steps: transient grid fields are dropped from L2 after their only read (discard.global.L2) instead of being written back
    bool k1_quad = true;     // spec->grid ensemble batches at T30: four fields at a time (k_s2g_quad); 0: the streaming kernel
    bool k2_quad = true;     // grid->spec ensemble batches at T30: four fields at a time, FFT + DMMA Legendre (k_g2s_quad); 0: the streaming kernel with the dense operator
    int k2_field = 0;        // 1: ensemble batches, 2: every launch incl. the single-member step (one CTA per field)      // grid->spec ensemble batches: whole-field FFT kernel (k_g2s_field) instead of the four wavenumber-group CTAs with the dense operator
    int precision = 0;       // 0 fp64 everywhere; 1 real32 spherical-harmonic transforms (transforms_f32.cu), fp64 elsewhere
    int num_sms = 148;
    int occ_k1[2] = {0, 0}, occ_k2 = 0, occ_k2b = 0;   // occupancy of the streaming transform kernels on this context's device (filled on first use)
    spd::DevBuf<unsigned long long> trace;
    int member_offset = 0;   // global index of member 0 of this context (SPPT stream id of a sharded ensemble)
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // device->host copy of the state, overlapped with the output conversions (speedy_run_steps_host)
    cudaEvent_t copy_event = nullptr;
    long long launches = 0;
    bool use_graphs = true;
    // device tables
    std::map<std::string, spd::DevBuf<double>> dtab;
    spd::DevBuf<int> d_qtile, d_qtile_inv;
    spd::DevBuf<float> f32tab;          // real32 copies of the transform tables (precision = 1, transforms_f32.cu)
    spd::DevTables dv{};
    // scratch for host-pointer API calls
    spd::DevBuf<double> scratch_a, scratch_b, scratch_c, scratch_d;
    std::map<unsigned long long, spd::DevBuf<spd::XDesc>> desc_cache;
    spd::Model* model = nullptr;
    void ensure_scratch(spd::DevBuf<double>& b, size_t n) { if (b.n < n) b.alloc(n); }
};

namespace spd {
void upload_tables(speedy_ctx* ctx);            // abi.cu
void upload_implicit(speedy_ctx* ctx);          // abi.cu
void upload_level_consts(speedy_ctx* ctx);      // consts in dynamics.cu

// transforms.cu ----------------------------------------------------------------------
// mode: 0 full spec->grid, 1 legendre_inv only (out = (2mx,il)), 2 fourier_inv only (in = (2mx,il))
struct CloseArgs;   // close_step.cuh
void launch_spec_to_grid(speedy_ctx* ctx, const double* d_in, long long in_member_stride,
                         const XDesc* d_desc, int nbatch, double* d_out, long long out_member_stride,
                         int nmembers, int mode, const CloseArgs* close = nullptr, bool quad_ok = false);   // quad_ok: derived fields (op != 0) come as aligned pairs
// mode: 0 full grid->spec, 1 fourier_dir only (out = (2mx,il)), 2 legendre_dir only (in = (2mx,il))
// what the main-loop step of an ensemble tells its grid->spec launch (the quad kernel honours it; the other variants ignore it,
// and the step only sets it when the quad kernel is the one selected)
struct G2sStepOpts {
    bool transient_input = false;      // the input fields are dead once read (the next step rewrites them): their L2 lines may be dropped
    long long out_field_stride = 0;    // doubles between consecutive output fields (0: packed, 2 * nspec)
};
void launch_grid_to_spec(speedy_ctx* ctx, const double* d_in, long long in_member_stride,
                         const XDesc* d_desc, int nbatch, double* d_out, long long out_member_stride,
                         int nmembers, int mode, const int* gate = nullptr, G2sStepOpts step = G2sStepOpts());
void setup_transform_kernels();
void setup_f32_kernels();          // transforms_f32.cu
void setup_column_kernels();       // physics.cu
void setup_spec_step_kernels();    // dynamics.cu
// transforms_quad.cu (T30 ensemble batches)
void setup_quad_kernels();
bool s2g_quad_selected(const speedy_ctx* ctx, int nbatch, int nmembers, bool quad_ok);
bool g2s_quad_selected(speedy_ctx* ctx, int nbatch, int nmembers);                       // transforms.cu: would this grid->spec launch take the quad kernel   // transforms.cu: would this spec->grid launch take the quad kernel
unsigned s2g_quad_ready_counts(int nbatch);                                             // member_ready.cuh: counts per member of one launch
void build_quad_tables(const Tables& t, std::vector<int>& tiles, std::vector<double>& polyq);
void build_quad_inverse_tables(const Tables& t, std::vector<int>& tiles, std::vector<double>& polyi);
// transforms_f32.cu (precision = 1)
void launch_spec_to_grid_f32(speedy_ctx* ctx, const double* d_in, long long in_ms, const XDesc* d_desc, int nbatch, double* d_out, long long out_ms, int nmembers, const CloseArgs& cl);
void launch_grid_to_spec_f32(speedy_ctx* ctx, const double* d_in, long long in_ms, const XDesc* d_desc, int nbatch, double* d_out, long long out_ms, int nmembers, const int* gate);
int polyd_groups(int trunc);
int polyd_mg(int trunc);
int polyt_row(int trunc);
// spectral_ops.cu ---------------------------------------------------------------------
void launch_spectral_op(speedy_ctx* ctx, int op, const double* a, const double* b, double* o1, double* o2, int nbatch);
enum { OP_LAPLACIAN = 0, OP_INVLAPLACIAN, OP_GRAD, OP_VDS, OP_UVSPEC, OP_TRUNCT };
}  // namespace spd
