// Warp-cooperative backward real FFT — the reference's FFTPACK passes (fftpack.f90:69-134 rfftb1, :204-424
// radb2 / radb3 / radb4) with the butterflies of one pass spread over the 32 lanes of a warp.  One warp
// transforms one latitude row that lives in shared memory; the arithmetic of every butterfly is the
// reference's, expression by expression, with the same twiddle table (single-precision 2*pi seed,
// fftpack.f90:39), so the result agrees with the reference's transform to the last bits — unlike a generic
// FFT, which differs at 1e-7 — at 1/8 of the flops of the dense-operator form.
//
// cc / ch are the pass input / output as in FFTPACK: cc(ido,ip,l1), ch(ido,l1,ip), 0-based storage of the
// reference's 1-based arrays; wa1.. point so that wa[i-2], wa[i-1] (i = 3,5,..) are the reference's entries.
#pragma once

namespace spd {

#define FCC(i, j, k) cc[((i)-1) + ido * (((j)-1) + IP * ((k)-1))]
#define FCH(i, k, j) ch[((i)-1) + ido * (((k)-1) + l1 * ((j)-1))]

__device__ __forceinline__ void w_radb2(int ido, int l1, const double* cc, double* ch, const double* wa1, int lane) {
    constexpr int IP = 2;
    for (int k = 1 + lane; k <= l1; k += 32) {
        FCH(1, k, 1) = FCC(1, 1, k) + FCC(ido, 2, k);
        FCH(1, k, 2) = FCC(1, 1, k) - FCC(ido, 2, k);
    }
    if (ido < 2) return;
    if (ido > 2) {
        const int idp2 = ido + 2, ni = (ido - 1) / 2;
        for (int t = lane; t < l1 * ni; t += 32) {
            const int k = t / ni + 1, i = 3 + 2 * (t - (k - 1) * ni);
            const int ic = idp2 - i;
            FCH(i - 1, k, 1) = FCC(i - 1, 1, k) + FCC(ic - 1, 2, k);
            const double tr2 = FCC(i - 1, 1, k) - FCC(ic - 1, 2, k);
            FCH(i, k, 1) = FCC(i, 1, k) - FCC(ic, 2, k);
            const double ti2 = FCC(i, 1, k) + FCC(ic, 2, k);
            FCH(i - 1, k, 2) = wa1[i - 2] * tr2 - wa1[i - 1] * ti2;
            FCH(i, k, 2) = wa1[i - 2] * ti2 + wa1[i - 1] * tr2;
        }
        if (ido % 2 == 1) return;
    }
    for (int k = 1 + lane; k <= l1; k += 32) {
        FCH(ido, k, 1) = FCC(ido, 1, k) + FCC(ido, 1, k);
        FCH(ido, k, 2) = -(FCC(1, 2, k) + FCC(1, 2, k));
    }
}

__device__ __forceinline__ void w_radb3(int ido, int l1, const double* cc, double* ch, const double* wa1, const double* wa2, int lane) {
    constexpr int IP = 3;
    const double taur = -.5;
    const double taui = (double)(.5f * sqrtf(3.f));       // .5*sqrt(3.) in real32 (fftpack.f90:269)
    for (int k = 1 + lane; k <= l1; k += 32) {
        const double tr2 = FCC(ido, 2, k) + FCC(ido, 2, k);
        const double cr2 = FCC(1, 1, k) + taur * tr2;
        FCH(1, k, 1) = FCC(1, 1, k) + tr2;
        const double ci3 = taui * (FCC(1, 3, k) + FCC(1, 3, k));
        FCH(1, k, 2) = cr2 - ci3;
        FCH(1, k, 3) = cr2 + ci3;
    }
    if (ido == 1) return;
    const int idp2 = ido + 2, ni = (ido - 1) / 2;
    for (int t = lane; t < l1 * ni; t += 32) {
        const int k = t / ni + 1, i = 3 + 2 * (t - (k - 1) * ni);
        const int ic = idp2 - i;
        const double tr2 = FCC(i - 1, 3, k) + FCC(ic - 1, 2, k);
        const double cr2 = FCC(i - 1, 1, k) + taur * tr2;
        FCH(i - 1, k, 1) = FCC(i - 1, 1, k) + tr2;
        const double ti2 = FCC(i, 3, k) - FCC(ic, 2, k);
        const double ci2 = FCC(i, 1, k) + taur * ti2;
        FCH(i, k, 1) = FCC(i, 1, k) + ti2;
        const double cr3 = taui * (FCC(i - 1, 3, k) - FCC(ic - 1, 2, k));
        const double ci3 = taui * (FCC(i, 3, k) + FCC(ic, 2, k));
        const double dr2 = cr2 - ci3;
        const double dr3 = cr2 + ci3;
        const double di2 = ci2 + cr3;
        const double di3 = ci2 - cr3;
        FCH(i - 1, k, 2) = wa1[i - 2] * dr2 - wa1[i - 1] * di2;
        FCH(i, k, 2) = wa1[i - 2] * di2 + wa1[i - 1] * dr2;
        FCH(i - 1, k, 3) = wa2[i - 2] * dr3 - wa2[i - 1] * di3;
        FCH(i, k, 3) = wa2[i - 2] * di3 + wa2[i - 1] * dr3;
    }
}

__device__ __forceinline__ void w_radb4(int ido, int l1, const double* cc, double* ch, const double* wa1, const double* wa2, const double* wa3, int lane) {
    constexpr int IP = 4;
    const double sqrt2 = (double)sqrtf(2.f);               // sqrt(2.) in real32 (fftpack.f90:341)
    for (int k = 1 + lane; k <= l1; k += 32) {
        const double tr1 = FCC(1, 1, k) - FCC(ido, 4, k);
        const double tr2 = FCC(1, 1, k) + FCC(ido, 4, k);
        const double tr3 = FCC(ido, 2, k) + FCC(ido, 2, k);
        const double tr4 = FCC(1, 3, k) + FCC(1, 3, k);
        FCH(1, k, 1) = tr2 + tr3;
        FCH(1, k, 2) = tr1 - tr4;
        FCH(1, k, 3) = tr2 - tr3;
        FCH(1, k, 4) = tr1 + tr4;
    }
    if (ido < 2) return;
    if (ido > 2) {
        const int idp2 = ido + 2, ni = (ido - 1) / 2;
        for (int t = lane; t < l1 * ni; t += 32) {
            const int k = t / ni + 1, i = 3 + 2 * (t - (k - 1) * ni);
            const int ic = idp2 - i;
            const double ti1 = FCC(i, 1, k) + FCC(ic, 4, k);
            const double ti2 = FCC(i, 1, k) - FCC(ic, 4, k);
            const double ti3 = FCC(i, 3, k) - FCC(ic, 2, k);
            const double tr4 = FCC(i, 3, k) + FCC(ic, 2, k);
            const double tr1 = FCC(i - 1, 1, k) - FCC(ic - 1, 4, k);
            const double tr2 = FCC(i - 1, 1, k) + FCC(ic - 1, 4, k);
            const double ti4 = FCC(i - 1, 3, k) - FCC(ic - 1, 2, k);
            const double tr3 = FCC(i - 1, 3, k) + FCC(ic - 1, 2, k);
            FCH(i - 1, k, 1) = tr2 + tr3;
            const double cr3 = tr2 - tr3;
            FCH(i, k, 1) = ti2 + ti3;
            const double ci3 = ti2 - ti3;
            const double cr2 = tr1 - tr4;
            const double cr4 = tr1 + tr4;
            const double ci2 = ti1 + ti4;
            const double ci4 = ti1 - ti4;
            FCH(i - 1, k, 2) = wa1[i - 2] * cr2 - wa1[i - 1] * ci2;
            FCH(i, k, 2) = wa1[i - 2] * ci2 + wa1[i - 1] * cr2;
            FCH(i - 1, k, 3) = wa2[i - 2] * cr3 - wa2[i - 1] * ci3;
            FCH(i, k, 3) = wa2[i - 2] * ci3 + wa2[i - 1] * cr3;
            FCH(i - 1, k, 4) = wa3[i - 2] * cr4 - wa3[i - 1] * ci4;
            FCH(i, k, 4) = wa3[i - 2] * ci4 + wa3[i - 1] * cr4;
        }
        if (ido % 2 == 1) return;
    }
    for (int k = 1 + lane; k <= l1; k += 32) {
        const double ti1 = FCC(1, 2, k) + FCC(1, 4, k);
        const double ti2 = FCC(1, 4, k) - FCC(1, 2, k);
        const double tr1 = FCC(ido, 1, k) - FCC(ido, 3, k);
        const double tr2 = FCC(ido, 1, k) + FCC(ido, 3, k);
        FCH(ido, k, 1) = tr2 + tr2;
        FCH(ido, k, 2) = sqrt2 * (tr1 - ti1);
        FCH(ido, k, 3) = ti2 + ti2;
        FCH(ido, k, 4) = -sqrt2 * (tr1 + ti1);
    }
}
#undef FCC
#undef FCH

// rfftb1 (fftpack.f90:69-134) for N = 96 (factors 2,4,4,3) or N = 144 (4,4,3,3): the row is in c on entry and on
// exit (both factor lists have four passes, so the ping-pong ends where it started); ch is scratch; wa is the 0-based
// twiddle table of rffti1.  All 32 lanes of the warp call this together.
template <int N>
__device__ __forceinline__ void warp_rfftb(double* c, double* ch, const double* wa, int lane) {
    constexpr int F0 = (N == 96) ? 2 : 4, F1 = 4, F2 = (N == 96) ? 4 : 3, F3 = 3;
    constexpr int fac[4] = {F0, F1, F2, F3};
    int l1 = 1, iw = 1;
    double* in = c;
    double* out = ch;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        const int ip = fac[p], l2 = ip * l1, ido = N / l2;
        const double* w1 = wa + iw - 2;          // so that w1[i-2] is the reference's wa(iw + i - 3), i = 3,5,..
        if (ip == 4) w_radb4(ido, l1, in, out, w1, w1 + ido, w1 + 2 * ido, lane);
        else if (ip == 2) w_radb2(ido, l1, in, out, w1, lane);
        else w_radb3(ido, l1, in, out, w1, w1 + ido, lane);
        __syncwarp();
        double* tsw = in; in = out; out = tsw;
        l1 = l2;
        iw += (ip - 1) * ido;
    }
}

}  // namespace spd
