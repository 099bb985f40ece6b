// Probe: FP64 DFMA vs DMMA (mma.sync.m8n8k4.f64) throughput and dependent-issue latency on this GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_probe tools/dmma_probe.cu && tools/dmma_probe
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double* out, int iters) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_dmma(double* out, int iters) {
    double c[8][2]; for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = threadIdx.x * 1e-3;
    double a = 1.0 + threadIdx.x * 1e-9, b = 0.5;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    }
    double s = 0; for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// dependent chains, one warp: cycles per op
__global__ void k_lat(double* out, long long* cyc, int iters) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (double)((i * 37 + 11) % 1024);
    __syncthreads();
    double x = threadIdx.x * 1e-3, c0 = x, c1 = x;
    const double b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) x = fma(x, b, c);
    long long t1 = clock64();
    for (int i = 0; i < iters; i++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(b), "d"(c));
    long long t2 = clock64();
    int idx = threadIdx.x;
    for (int i = 0; i < iters; i++) idx = (int)sm[idx & 1023];      // LDS.64 -> cvt -> address
    long long t3 = clock64();
    double y = x;
    for (int i = 0; i < iters; i++) y = y * b;                        // DMUL chain
    long long t4 = clock64();
    double z = x + 2.0;
    for (int i = 0; i < iters; i++) z = exp(z * 1e-9);                 // libdevice exp chain
    long long t5 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; }
    out[threadIdx.x] = x + c0 + c1 + idx + y + z;
}
int main() {
    double* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
    long long* cyc; cudaMallocManaged(&cyc, 8 * sizeof(long long));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, blocks = 148 * 4, threads = 512;
    for (int rep = 0; rep < 2; rep++) {
        float ms;
        cudaEventRecord(e0); k_dfma<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 8 * iters * (double)blocks * threads;
        printf("DFMA  %.3f ms  %.2f TFLOP/s\n", ms, fl / ms / 1e9);
        cudaEventRecord(e0); k_dmma<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        fl = 2.0 * 8 * 8 * 4 * 8 * iters * (double)blocks * (threads / 32);
        printf("DMMA  %.3f ms  %.2f TFLOP/s\n", ms, fl / ms / 1e9);
    }
    const int n = 4096;
    k_lat<<<1, 32>>>(out, cyc, n); cudaDeviceSynchronize();
    k_lat<<<1, 32>>>(out, cyc, n); cudaDeviceSynchronize();
    printf("dependent-chain cycles/op (1 warp): DFMA %.1f  DMMA.8x8x4 %.1f  LDS.64+cvt %.1f  DMUL %.1f  exp() %.1f\n",
           (double)cyc[0] / n, (double)cyc[1] / n, (double)cyc[2] / n, (double)cyc[3] / n, (double)cyc[4] / n);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
