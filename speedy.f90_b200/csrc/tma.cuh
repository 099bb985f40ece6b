// Bulk asynchronous copies (the TMA engine's 1-D path, SASS UBLKCP) + mbarrier completion, and
// named barriers for warp-role hand-offs.  sm_90+/sm_100a PTX, no tensor maps needed: every
// tile staged by the kernels of this library is a contiguous run of >= 16 bytes.
#pragma once
#include <cstdint>

namespace spd {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the initialised barriers visible to the async proxy (call once, then __syncthreads())
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned; completion is
// signalled on `bar` as `bytes` transaction bytes
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// global -> shared TILE copy through a tensor map (cp.async.bulk.tensor, SASS UTMALDG): one instruction moves a
// [rows x 32-column] box of a [member][field][column] array; coordinates are element indices, innermost first
__device__ __forceinline__ void tensor_g2s_3d(void* dst_smem, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tensor_g2s_4d(void* dst_smem, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}

// shared -> global bulk copy (TMA store, bulk-group completion); bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
// shared -> global TILE store through a tensor map (cp.async.bulk.tensor, SASS UTMASTG), bulk-group completion
__device__ __forceinline__ void tensor_s2g_4d(const void* tmap, int c0, int c1, int c2, int c3, const void* src_smem) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                 ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(src_smem)) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// close the bulk group and wait until its copies have finished READING shared memory (the buffer may be reused)
__device__ __forceinline__ void bulk_commit_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// wait until every bulk store of this thread is complete (writes performed)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// order earlier generic-proxy accesses to shared memory before later async-proxy (TMA) writes to the same locations
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// named barriers (ids 1..15; 0 is __syncthreads): producer/consumer hand-off between warp roles
__device__ __forceinline__ void named_arrive(int id, int nthreads) {
    __threadfence_block();   // the arriving side's shared/global writes are ordered before the arrival
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- programmatic dependent launch (PDL): a kernel launched with launch_pdl() may start while its
// predecessor in the stream / graph is still draining; everything before pdl_wait() must not touch data
// the predecessor writes (barrier set-up, staging of constant tables), everything after sees it complete.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- in-graph timeline (debug aid, speedy_trace): per kernel slot the earliest CTA start and the latest
// CTA end on the GPU's global nanosecond timer.  trace == nullptr in production launches.
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// The 128-byte line at p (aligned) holds data nobody will read again before it is rewritten: L2 may drop it without writing it back.
__device__ __forceinline__ void l2_discard_line(const void* p) { asm volatile("discard.global.L2 [%0], 128;" :: "l"(p) : "memory"); }

__device__ __forceinline__ void trace_begin(unsigned long long* trace, int slot) { if (trace) atomicMin(&trace[slot], gtimer()); }
__device__ __forceinline__ void trace_end(unsigned long long* trace, int slot) { if (trace) atomicMax(&trace[4 + slot], gtimer()); }

#ifdef __CUDACC__
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

}  // namespace spd
