// Probe: tensor memory (TMEM) as a per-lane scratchpad — every warp of a 24-warp CTA parks 24 doubles per lane in its lane quarter with
// tcgen05.st (SASS STTM) and reads them back with tcgen05.ld (LDTM) after the registers were clobbered.  Checks the addressing used by
// the quad transform kernels (lane field = 32 x (warp % 4), one column range per warp of a quarter).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tmem_probe tools/tmem_probe.cu && tools/tmem_probe
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_st16(uint32_t addr, const double* v) {     // 8 doubles = 16 columns
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { r[2 * i] = (uint32_t)__double2loint(v[i]); r[2 * i + 1] = (uint32_t)__double2hiint(v[i]); }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                    "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, double* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                   "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(addr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}

__global__ void __launch_bounds__(768, 1) k(const double* in, double* out) {
    __shared__ uint32_t taddr_s;
    const int w = threadIdx.x >> 5;
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&taddr_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = taddr_s;
    const uint32_t addr = base + ((uint32_t)(32 * (w & 3)) << 16) + 48 * (w >> 2);
    double v[24];
    for (int i = 0; i < 24; i++) v[i] = in[(size_t)(blockIdx.x * 768 + threadIdx.x) * 24 + i];
    for (int c = 0; c < 3; c++) tmem_st16(addr + 16 * c, v + 8 * c);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    for (int i = 0; i < 24; i++) v[i] = -1.0;
    __syncthreads();
    for (int rep = 0; rep < 2; rep++)
        for (int c = 0; c < 3; c++) tmem_ld16(addr + 16 * c, v + 8 * c);
    for (int i = 0; i < 24; i++) out[(size_t)(blockIdx.x * 768 + threadIdx.x) * 24 + i] = v[i];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "r"(512u) : "memory");
}

int main() {
    const int nb = 296, n = nb * 768 * 24;
    std::vector<double> h(n), o(n);
    for (int i = 0; i < n; i++) h[i] = 1.0 + i * 1e-3;
    double *din, *dout;
    cudaMalloc(&din, n * 8); cudaMalloc(&dout, n * 8);
    cudaMemcpy(din, h.data(), n * 8, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 3; rep++) k<<<nb, 768>>>(din, dout);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(o.data(), dout, n * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < n; i++) bad += o[i] != h[i];
    printf("tmem probe: %s, %d mismatches of %d (%s)\n", bad == 0 && e == cudaSuccess ? "PASS" : "FAIL", bad, n, cudaGetErrorString(e));
    return bad != 0;
}
