// Per-member hand-off from the quad spec->grid kernel to the column kernel of the same model step.
// The grid fields of a member are complete when every quad of that member has been stored; with 77-85 fields per member an 8-member
// step holds 160-176 quads for 147 CTAs, so a few CTAs run a second quad while the others are done: their late quads are the last
// member's (quads are dealt out in member order, round robin), and the column tiles of the earlier members need not wait for them.
//   producer (k_s2g_quad): the thread that issued a band group's TMA stores waits for their completion (cp.async.bulk.wait_group 0:
//     writes done), then release-increments ready[member]; the closing CTA adds one to every member after the calendar is advanced.
//   consumer (k_grid_columns): one thread acquire-polls ready[member] until it reaches the target (3 band groups x quads per member
//     + 1), in place of griddepcontrol.wait; the kernel is a programmatic dependent of the transform, so its CTAs start on the SMs the
//     finished transform CTAs leave (every transform CTA is resident before the dependent launches: no SM is taken from the producer).
//   k_spec_step (a full dependency later) zeroes the counts for the next step.
// A poll that lasts longer than READY_TIMEOUT_NS (5 s) traps instead of hanging the device.
#pragma once
#include "tma.cuh"

namespace spd {

constexpr unsigned long long READY_TIMEOUT_NS = 5000ull * 1000 * 1000;   // 5 s: far beyond any real wait, also under compute-sanitizer

__device__ __forceinline__ void ready_signal(unsigned* cnt) {
    asm volatile("fence.proxy.async;" ::: "memory");   // the stores went through the async proxy
    __threadfence();
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(cnt) : "memory");
}
__device__ __forceinline__ void ready_wait(const unsigned* cnt, unsigned target) {
    const unsigned long long t0 = gtimer();
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
        if (v >= target) break;
        __nanosleep(200);
        if (gtimer() - t0 > READY_TIMEOUT_NS) __trap();
    }
    asm volatile("fence.proxy.async;" ::: "memory");   // the tile copies that follow read through the async proxy
}

}  // namespace spd
