// K3 (stand-alone form): batched spectral operators behind the C ABI
// (spectral.f90:84-96 laplacian/inverse_laplacian, :124 grad, :146 vds, :173 uvspec, :229 trunct).
#include "spectral_ops.cuh"

namespace spd {

__global__ void k_spectral_op(int op, const double* __restrict__ a, const double* __restrict__ b,
                              double* __restrict__ o1, double* __restrict__ o2, int nbatch, DevTables tv) {
    const int nsp = tv.mx * tv.nx;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nbatch * nsp) return;
    const int f = t / nsp, r = t - f * nsp;
    const int n = r / tv.mx, m = r - n * tv.mx;
    const size_t fo = (size_t)f * nsp * 2;
    const size_t q = m + (size_t)tv.mx * n;
    switch (op) {
        case OP_LAPLACIAN: { cd v = ld(a + fo, tv.mx, m, n); st(o1 + fo, tv.mx, m, n, cd{-v.re * tv.el2[q], -v.im * tv.el2[q]}); break; }
        case OP_INVLAPLACIAN: { cd v = ld(a + fo, tv.mx, m, n); st(o1 + fo, tv.mx, m, n, cd{-v.re * tv.elm2[q], -v.im * tv.elm2[q]}); break; }
        case OP_GRAD: { cd dx, dy; dev_grad(tv, a + fo, m, n, dx, dy); st(o1 + fo, tv.mx, m, n, dx); st(o2 + fo, tv.mx, m, n, dy); break; }
        case OP_VDS: { cd vo, di; dev_vds(tv, a + fo, b + fo, m, n, vo, di); st(o1 + fo, tv.mx, m, n, vo); st(o2 + fo, tv.mx, m, n, di); break; }
        case OP_UVSPEC: { cd uc, vc; dev_uvspec(tv, a + fo, b + fo, m, n, uc, vc); st(o1 + fo, tv.mx, m, n, uc); st(o2 + fo, tv.mx, m, n, vc); break; }
        case OP_TRUNCT: { cd v = ld(a + fo, tv.mx, m, n); st(o1 + fo, tv.mx, m, n, tv.trfilt[q] * v); break; }
    }
}

void launch_spectral_op(speedy_ctx* ctx, int op, const double* a, const double* b, double* o1, double* o2, int nbatch) {
    const int total = nbatch * ctx->d.mx * ctx->d.nx;
    k_spectral_op<<<(total + 127) / 128, 128, 0, ctx->stream>>>(op, a, b, o1, o2, nbatch, ctx->dv);
    ctx->launches++;
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace spd
