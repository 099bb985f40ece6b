// Closing of a main-loop step (speedy.f90:41-47): the final check_diagnostics reduction in a fixed order
// (diagnostics.f90:16-75), the range guard, and the calendar advance with the flags of the next step.
// One CTA of >= 256 threads runs it.  In the production graph this is an extra CTA of the NEXT step's
// spec->grid kernel (which does not read the clock), so it is off the critical path; a stand-alone kernel
// flushes it when no step follows.
#pragma once
#include "model.h"
#include "calendar.h"

namespace spd {

struct CloseArgs {
    DevClock* clk;            // nullptr: nothing to close
    const double* partial;    // [member][block][kx][2] + [member][kx]: partial sums written by k_spec_step
    int nb, ne;               // spectral-step blocks per member, members
    unsigned long long* trace; // in-graph timeline buffer (stand-alone closer in trace mode), else nullptr
    unsigned* ready;          // [member] completion counts of the spec->grid kernel that carries this closer (member_ready.cuh), or nullptr
};

// threads 0..255 of the calling CTA (warp w <-> (member, level) pairs w, w+8, ...; lanes take blocks in a fixed order)
__device__ __forceinline__ void close_step_cta(const CloseArgs& a, int tid) {
    constexpr int KXL = 8;
    if (tid >= 32 * KXL) return;
    const int c = tid & 31, k = tid >> 5;
    if (!a.clk->close_pending) return;
    const int nb = a.nb, ne = a.ne;
    // warp k = level k; four members at a time, their loads issued together (one L2 round trip per four members)
    const int kk = k;
    for (int e0 = 0; e0 < ne; e0 += 4) {
        double d1[4], d2[4], d3[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int e = e0 + u;
            d1[u] = 0.0; d2[u] = 0.0; d3[u] = 0.0;
            if (e < ne) {
                for (int b = c; b < nb; b += 32) {
                    const double* part = a.partial + ((size_t)e * nb + b) * (2 * KXL);
                    d1[u] += __ldcg(part + kk); d2[u] += __ldcg(part + KXL + kk);
                }
                if (c == 0) d3[u] = __ldcg(a.partial + (size_t)ne * nb * 2 * KXL + (size_t)e * KXL + kk);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { d1[u] += __shfl_xor_sync(0xffffffffu, d1[u], o); d2[u] += __shfl_xor_sync(0xffffffffu, d2[u], o); }
        }
        if (c == 0) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int e = e0 + u;
                if (e >= ne) break;
                const double t3 = (double)sqrtf(0.5f) * d3[u];
                // after a range failure the numbers of the failing step stay (the reference prints them and stops, diagnostics.f90:60-69)
                const int failed = a.clk->diag_fail;
                if (e == 0 && (failed == 0 || failed == a.clk->model_step)) { a.clk->diag[kk] = d1[u]; a.clk->diag[KXL + kk] = d2[u]; a.clk->diag[2 * KXL + kk] = t3; }
                const bool bad = !(d1[u] <= 500.0) || !(d2[u] <= 500.0) || !(t3 >= 180.0) || !(t3 <= 320.0);
                if (bad) atomicCAS(&a.clk->diag_fail, 0, a.clk->model_step);
            }
        }
    }
    // all 256 threads of the closing group (named barrier: the rest of the CTA may be elsewhere)
    asm volatile("bar.sync 15, 256;" ::: "memory");
    if (tid == 0) {
        cal_advance(*a.clk);          // speedy.f90:44-47
        a.clk->slab_pending = 1;      // couple_sea_land of this step rides in the next column kernel
        a.clk->close_pending = 0;
        if (a.trace) {                // close the step's timeline: durations [8..11], gaps before each kernel [12..15], steps [16]
            unsigned long long* tr = a.trace;
            unsigned long long prev = tr[17];
            for (int sl = 0; sl < 4; sl++) {
                const unsigned long long t0 = tr[sl], t1 = tr[4 + sl];
                tr[8 + sl] += t1 - t0;
                if (prev && t0 > prev) tr[12 + sl] += t0 - prev;
                prev = t1;
                tr[sl] = ~0ull; tr[4 + sl] = 0ull;
            }
            tr[17] = prev;
            tr[16] += 1;
        }
    }
}

}  // namespace spd
