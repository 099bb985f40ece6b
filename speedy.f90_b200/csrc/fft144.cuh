// Backward real FFT of length 144 (T47) — FFTPACK's passes for the factor list 4,4,3,3 of rffti1 (fftpack.f90:69-134
// rfftb1; :328 radb4, :256 radb3), regrouped like fft96.cuh so that a thread keeps its data in registers across two passes:
//   stage 1 = radb4 (ido 36, l1 1) + radb4 (ido 9, l1 4).  The second pass's general butterfly i (3,5,7,9), for all four
//             k, closes over the four first-pass butterflies i, i+18, 20-i, 38-i: 32 reals.  A thread takes one such set
//             for TWO of the four k (it evaluates the four first-pass butterflies and keeps the two outputs it needs),
//             so that its live data stay at 16 reals.  The i = 1 case closes over the first pass's i = 1, i = ido and
//             i = 19 butterflies (16 reals), likewise split by k pair;
//   stage 2 = radb3 (ido 3, l1 16) + radb3 (ido 1, l1 48): sixteen closed sets of 9 contiguous reals.
// Butterflies, constants and twiddles are the reference's, expression by expression (see fft96.cuh).
// Data layout: element p of row r at X[p * XS + r]; X and T are already offset by the row; wa is rffti1's table, 0-based.
#pragma once

namespace spd {

struct Fft144 {
    // one first-pass radb4 butterfly (fftpack.f90:356-393, ido 36) at i1 (ic = 38 - i1); returns ch(i1-1,1,j), ch(i1,1,j)
    // for j = 2*kh + 1 (a) and j = 2*kh + 2 (b)
    template <int XS>
    static __device__ __forceinline__ void pass1_at(const double* X, const double* wa, int i1, int kh, double& ar, double& ai, double& br, double& bi) {
        const int ic = 38 - i1;
        const double c1r = X[(i1 - 2) * XS], c1i = X[(i1 - 1) * XS], c3r = X[(i1 + 70) * XS], c3i = X[(i1 + 71) * XS];
        const double c2r = X[(ic + 34) * XS], c2i = X[(ic + 35) * XS], c4r = X[(ic + 106) * XS], c4i = X[(ic + 107) * XS];
        const double ti1 = c1i + c4i;
        const double ti2 = c1i - c4i;
        const double ti3 = c3i - c2i;
        const double tr4 = c3i + c2i;
        const double tr1 = c1r - c4r;
        const double tr2 = c1r + c4r;
        const double ti4 = c3r - c2r;
        const double tr3 = c3r + c2r;
        if (kh == 0) {
            const double w1r = wa[i1 - 3], w1i = wa[i1 - 2];
            const double cr2 = tr1 - tr4, ci2 = ti1 + ti4;
            ar = tr2 + tr3;
            ai = ti2 + ti3;
            br = w1r * cr2 - w1i * ci2;
            bi = w1r * ci2 + w1i * cr2;
        } else {
            const double w2r = wa[i1 + 33], w2i = wa[i1 + 34], w3r = wa[i1 + 69], w3i = wa[i1 + 70];
            const double cr3 = tr2 - tr3, ci3 = ti2 - ti3, cr4 = tr1 + tr4, ci4 = ti1 - ti4;
            ar = w2r * cr3 - w2i * ci3;
            ai = w2r * ci3 + w2i * cr3;
            br = w3r * cr4 - w3i * ci4;
            bi = w3r * ci4 + w3i * cr4;
        }
    }

    // stage 1, general set of the second pass's butterfly i (3,5,7,9), k = 2*kh + 1 and 2*kh + 2
    template <int XS>
    static __device__ __forceinline__ void stage1_general(const double* X, double* T, const double* wa, int i, int kh) {
        // [0] i, [1] i+18, [2] 20-i, [3] 38-i  <->  the second pass's cc(.,1,k), cc(.,3,k), cc(ic..,2,k), cc(ic..,4,k)
        double ar[4], ai[4], br[4], bi[4];
        pass1_at<XS>(X, wa, i, kh, ar[0], ai[0], br[0], bi[0]);
        pass1_at<XS>(X, wa, i + 18, kh, ar[1], ai[1], br[1], bi[1]);
        pass1_at<XS>(X, wa, 20 - i, kh, ar[2], ai[2], br[2], bi[2]);
        pass1_at<XS>(X, wa, 38 - i, kh, ar[3], ai[3], br[3], bi[3]);
        const double w1r = wa[105 + i], w1i = wa[106 + i], w2r = wa[114 + i], w2i = wa[115 + i], w3r = wa[123 + i], w3i = wa[124 + i];
        auto radb4 = [&](const double* yr, const double* yi, int o) {   // o: position of ch(i-1,k,1)
            const double ti1 = yi[0] + yi[3];
            const double ti2 = yi[0] - yi[3];
            const double ti3 = yi[1] - yi[2];
            const double tr4 = yi[1] + yi[2];
            const double tr1 = yr[0] - yr[3];
            const double tr2 = yr[0] + yr[3];
            const double ti4 = yr[1] - yr[2];
            const double tr3 = yr[1] + yr[2];
            T[o * XS] = tr2 + tr3;
            const double cr3 = tr2 - tr3;
            T[(o + 1) * XS] = ti2 + ti3;
            const double ci3 = ti2 - ti3;
            const double cr2 = tr1 - tr4;
            const double cr4 = tr1 + tr4;
            const double ci2 = ti1 + ti4;
            const double ci4 = ti1 - ti4;
            T[(o + 36) * XS] = w1r * cr2 - w1i * ci2;
            T[(o + 37) * XS] = w1r * ci2 + w1i * cr2;
            T[(o + 72) * XS] = w2r * cr3 - w2i * ci3;
            T[(o + 73) * XS] = w2r * ci3 + w2i * cr3;
            T[(o + 108) * XS] = w3r * cr4 - w3i * ci4;
            T[(o + 109) * XS] = w3r * ci4 + w3i * cr4;
        };
        radb4(ar, ai, (i - 2) + 18 * kh);          // k = 2 kh + 1: position (i-2) + 9 (k-1)
        radb4(br, bi, (i - 2) + 18 * kh + 9);      // k = 2 kh + 2
    }

    // stage 1, the second pass's i = 1 case (fftpack.f90:343-353) for k = 2*kh + 1, 2*kh + 2: fed by the first pass's
    // i = 1 (:343-353) and i = ido (:397-408) cases and its butterfly 19
    template <int XS>
    static __device__ __forceinline__ void stage1_first(const double* X, double* T, const double* wa, int kh) {
        const double sqrt2 = (double)sqrtf(2.f);           // sqrt(2.) in real32 (fftpack.f90:341)
        double p[2], q[2];                                  // ch(1,1,j), ch(36,1,j) for the two j of this k pair
        {
            const double tr1 = X[0] - X[143 * XS];
            const double tr2 = X[0] + X[143 * XS];
            const double tr3 = X[71 * XS] + X[71 * XS];
            const double tr4 = X[72 * XS] + X[72 * XS];
            p[0] = kh ? tr2 - tr3 : tr2 + tr3;
            p[1] = kh ? tr1 + tr4 : tr1 - tr4;
        }
        {
            const double ti1 = X[36 * XS] + X[108 * XS];
            const double ti2 = X[108 * XS] - X[36 * XS];
            const double tr1 = X[35 * XS] - X[107 * XS];
            const double tr2 = X[35 * XS] + X[107 * XS];
            q[0] = kh ? ti2 + ti2 : tr2 + tr2;
            q[1] = kh ? -sqrt2 * (tr1 + ti1) : sqrt2 * (tr1 - ti1);
        }
        double yr[2], yi[2];                                // ch(18,1,j), ch(19,1,j)
        pass1_at<XS>(X, wa, 19, kh, yr[0], yi[0], yr[1], yi[1]);
#pragma unroll
        for (int t = 0; t < 2; t++) {
            const double tr1 = p[t] - q[t];
            const double tr2 = p[t] + q[t];
            const double tr3 = yr[t] + yr[t];
            const double tr4 = yi[t] + yi[t];
            const int o = 9 * (2 * kh + t);
            T[o * XS] = tr2 + tr3;
            T[(o + 36) * XS] = tr1 - tr4;
            T[(o + 72) * XS] = tr2 - tr3;
            T[(o + 108) * XS] = tr1 + tr4;
        }
    }

    // stage 2 for the third pass's k = k3 + 1 (k3 = 0..15): y[3*jj + j] = grid point k3 + 16 j + 48 jj (j, jj = 0..2)
    template <int XS>
    static __device__ __forceinline__ void stage2(const double* T, const double* wa, int k3, double (&y)[9]) {
        const double taur = -.5;
        const double taui = (double)(.5f * sqrtf(3.f));    // .5*sqrt(3.) in real32 (fftpack.f90:269)
        double e[9];
#pragma unroll
        for (int t = 0; t < 9; t++) e[t] = T[(9 * k3 + t) * XS];       // cc(i,j,k): e[(i-1) + 3 (j-1)]
        double z[3][3];                                                // ch(i,k,j)
        {   // i = 1 (fftpack.f90:271-278)
            const double tr2 = e[5] + e[5];
            const double cr2 = e[0] + taur * tr2;
            z[0][0] = e[0] + tr2;
            const double ci3 = taui * (e[6] + e[6]);
            z[0][1] = cr2 - ci3;
            z[0][2] = cr2 + ci3;
        }
        {   // i = 3, ic = 2 (fftpack.f90:283-301); twiddles wa(136 + ..) of the third pass
            const double w1r = wa[135], w1i = wa[136], w2r = wa[138], w2i = wa[139];
            const double tr2 = e[7] + e[3];
            const double cr2 = e[1] + taur * tr2;
            z[1][0] = e[1] + tr2;
            const double ti2 = e[8] - e[4];
            const double ci2 = e[2] + taur * ti2;
            z[2][0] = e[2] + ti2;
            const double cr3 = taui * (e[7] - e[3]);
            const double ci3 = taui * (e[8] + e[4]);
            const double dr2 = cr2 - ci3;
            const double dr3 = cr2 + ci3;
            const double di2 = ci2 + cr3;
            const double di3 = ci2 - cr3;
            z[1][1] = w1r * dr2 - w1i * di2;
            z[2][1] = w1r * di2 + w1i * dr2;
            z[1][2] = w2r * dr3 - w2i * di3;
            z[2][2] = w2r * di3 + w2i * dr3;
        }
#pragma unroll
        for (int j = 0; j < 3; j++) {   // radb3, ido = 1, for k = k3 + 16 j + 1
            const double tr2 = z[1][j] + z[1][j];
            const double cr2 = z[0][j] + taur * tr2;
            y[j] = z[0][j] + tr2;
            const double ci3 = taui * (z[2][j] + z[2][j]);
            y[3 + j] = cr2 - ci3;
            y[6 + j] = cr2 + ci3;
        }
    }
};

}  // namespace spd
