// Forward real FFT of length 96 — FFTPACK's passes (fftpack.f90:136-202 rfftf1 with rffti1's factor list 2,4,4,3 taken
// backwards: :774 radf3, :844 radf4 twice, :722 radf2), regrouped like fft96.cuh (its transpose) so that a thread keeps its
// data in registers across two passes:
//   stage A = radf3 (ido 1, l1 32) + radf4 (ido 3, l1 8): eight closed sets of 12 reals — grid points k + 8 j + 32 jj in,
//             12 contiguous reals out;
//   stage B = radf4 (ido 12, l1 2) + radf2 (ido 48, l1 1): five closed sets of 16 reals (radf4's general butterflies
//             i = 3,5,..,11 for both k, and the four radf2 butterflies they feed) and two of 8 (radf4's i = 1 and i = ido).
// Butterflies, constants and twiddles are the reference's, expression by expression.  Used by the whole-field grid->spec kernels
// (k_g2s_field in transforms.cu, k_g2s_quad in transforms_quad.cu); tests/test_fft96_cpu.py checks a host build bit for bit
// against the oracle's pass-by-pass rfftf1.
// Data layout as in fft96.cuh: element p of row r at X[p * XS + r]; wa is rffti1's table, 0-based.
#pragma once

namespace spd {

struct Fft96F {
    // stage A for radf4's k = k3 + 1 (k3 = 0..7): x[4*jj + j] = grid point k3 + 8 j + 32 jj in, T[12 k3 .. 12 k3 + 11] out
    template <int XS>
    static __device__ __forceinline__ void stageA(const double (&x)[12], double* T, const double* wa, int k3) {
        const double taur = -.5;
        const double taui = (double)(.5f * sqrtf(3.f));    // .5*sqrt(3.) in real32 (fftpack.f90:787)
        double c[3][4];                                    // radf4's cc(i,k,j) = radf3's ch(1,i,k + 8 (j-1))
#pragma unroll
        for (int j = 0; j < 4; j++) {                      // radf3, ido = 1 (fftpack.f90:789-794)
            const double a1 = x[j], a2 = x[4 + j], a3 = x[8 + j];
            const double cr2 = a2 + a3;
            c[0][j] = a1 + cr2;
            c[2][j] = taui * (a3 - a2);
            c[1][j] = a1 + taur * cr2;
        }
        double* o = T + (size_t)(12 * k3) * XS;            // ch(i,j,k) at (i-1) + 3 (j-1)
        {   // i = 1 (fftpack.f90:859-866)
            const double tr1 = c[0][1] + c[0][3];
            const double tr2 = c[0][0] + c[0][2];
            o[0 * XS] = tr1 + tr2;                         // ch(1,1,k)
            o[11 * XS] = tr2 - tr1;                        // ch(3,4,k)
            o[5 * XS] = c[0][0] - c[0][2];                 // ch(3,2,k)
            o[6 * XS] = c[0][3] - c[0][1];                 // ch(1,3,k)
        }
        {   // i = 3, ic = 2 (fftpack.f90:872-905); twiddles wa(85 + ..)
            const double w1r = wa[84], w1i = wa[85], w2r = wa[87], w2i = wa[88], w3r = wa[90], w3i = wa[91];
            const double cr2 = w1r * c[1][1] + w1i * c[2][1];
            const double ci2 = w1r * c[2][1] - w1i * c[1][1];
            const double cr3 = w2r * c[1][2] + w2i * c[2][2];
            const double ci3 = w2r * c[2][2] - w2i * c[1][2];
            const double cr4 = w3r * c[1][3] + w3i * c[2][3];
            const double ci4 = w3r * c[2][3] - w3i * c[1][3];
            const double tr1 = cr2 + cr4;
            const double tr4 = cr4 - cr2;
            const double ti1 = ci2 + ci4;
            const double ti4 = ci2 - ci4;
            const double ti2 = c[2][0] + ci3;
            const double ti3 = c[2][0] - ci3;
            const double tr2 = c[1][0] + cr3;
            const double tr3 = c[1][0] - cr3;
            o[1 * XS] = tr1 + tr2;                         // ch(2,1,k)
            o[9 * XS] = tr2 - tr1;                         // ch(1,4,k)
            o[2 * XS] = ti1 + ti2;                         // ch(3,1,k)
            o[10 * XS] = ti1 - ti2;                        // ch(2,4,k)
            o[7 * XS] = ti4 + tr3;                         // ch(2,3,k)
            o[3 * XS] = tr3 - ti4;                         // ch(1,2,k)
            o[8 * XS] = tr4 + ti3;                         // ch(3,3,k)
            o[4 * XS] = tr4 - ti3;                         // ch(2,2,k)
        }
    }

    // one radf2 butterfly (fftpack.f90:741-752) at i1 (ic = 50 - i1) on radf4's outputs of both k: (ar, ai) = ch(i1-1, ., 1),
    // ch(i1, ., 1) of radf4 (k = 1), (br, bi) the same for k = 2; writes the half-complex result
    // `St` is a store functor st(pos, value) for half-complex position pos (the pointer forms below store to Y[pos * XS]; the
    // quad kernel maps the position to a coefficient row and drops the positions above the truncation)
    template <int XS>
    struct StoreY {
        double* Y;
        __device__ __forceinline__ void operator()(int pos, double v) const { Y[pos * XS] = v; }
    };
    template <class St>
    static __device__ __forceinline__ void radf2_at_f(const St& st, const double* wa, int i1, double ar, double ai, double br, double bi) {
        const int ic = 50 - i1;
        const double wr = wa[i1 - 3], wi = wa[i1 - 2];
        const double tr2 = wr * br + wi * bi;
        const double ti2 = wr * bi - wi * br;
        st(i1 - 1, ai + ti2);                              // ch(i,1,k)
        st(47 + ic, ti2 - ai);                             // ch(ic,2,k)
        st(i1 - 2, ar + tr2);                              // ch(i-1,1,k)
        st(46 + ic, ar - tr2);                             // ch(ic-1,2,k)
    }

    // stage B, general set of radf4's butterfly i (3,5,..,11), both k
    template <int XS>
    static __device__ __forceinline__ void stageB_general(const double* T, double* Y, const double* wa, int i) { stageB_general_f<XS>(T, StoreY<XS>{Y}, wa, i); }
    template <int XS, class St>
    static __device__ __forceinline__ void stageB_general_f(const double* T, const St& st, const double* wa, int i) { stageB_general_g(LoadT<XS>{T}, st, wa, i); }
    // `Ld` is a load functor ld(blk, off) for stage A's output at position 12 blk + off (blk = k + 2 j of radf4's cc(., k, j))
    template <int XS>
    struct LoadT {
        const double* T;
        __device__ __forceinline__ double operator()(int blk, int off) const { return T[(12 * blk + off) * XS]; }
    };
    template <class Ld, class St>
    static __device__ __forceinline__ void stageB_general_g(const Ld& ld, const St& st, const double* wa, int i) {
        const double w1r = wa[45 + i], w1i = wa[46 + i], w2r = wa[57 + i], w2i = wa[58 + i], w3r = wa[69 + i], w3i = wa[70 + i];
        // radf4 outputs of k = 1, 2 at the four radf2 positions: [0] i (j 1), [1] i+24 (j 3), [2] 26-i (ic, j 2), [3] 50-i (ic, j 4)
        double pr[2][4], pi[2][4];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            // cc(i-1,k,j), cc(i,k,j): position (i - 2) + 12 k + 24 (j - 1) and the next
            const double c1r = ld(k, i - 2), c1i = ld(k, i - 1), c2r = ld(k + 2, i - 2), c2i = ld(k + 2, i - 1), c3r = ld(k + 4, i - 2), c3i = ld(k + 4, i - 1),
                         c4r = ld(k + 6, i - 2), c4i = ld(k + 6, i - 1);
            const double cr2 = w1r * c2r + w1i * c2i;
            const double ci2 = w1r * c2i - w1i * c2r;
            const double cr3 = w2r * c3r + w2i * c3i;
            const double ci3 = w2r * c3i - w2i * c3r;
            const double cr4 = w3r * c4r + w3i * c4i;
            const double ci4 = w3r * c4i - w3i * c4r;
            const double tr1 = cr2 + cr4;
            const double tr4 = cr4 - cr2;
            const double ti1 = ci2 + ci4;
            const double ti4 = ci2 - ci4;
            const double ti2 = c1i + ci3;
            const double ti3 = c1i - ci3;
            const double tr2 = c1r + cr3;
            const double tr3 = c1r - cr3;
            pr[k][0] = tr1 + tr2;                          // ch(i-1,1,k)
            pr[k][3] = tr2 - tr1;                          // ch(ic-1,4,k)
            pi[k][0] = ti1 + ti2;                          // ch(i,1,k)
            pi[k][3] = ti1 - ti2;                          // ch(ic,4,k)
            pr[k][1] = ti4 + tr3;                          // ch(i-1,3,k)
            pr[k][2] = tr3 - ti4;                          // ch(ic-1,2,k)
            pi[k][1] = tr4 + ti3;                          // ch(i,3,k)
            pi[k][2] = tr4 - ti3;                          // ch(ic,2,k)
        }
        radf2_at_f(st, wa, i, pr[0][0], pi[0][0], pr[1][0], pi[1][0]);
        radf2_at_f(st, wa, i + 24, pr[0][1], pi[0][1], pr[1][1], pi[1][1]);
        radf2_at_f(st, wa, 26 - i, pr[0][2], pi[0][2], pr[1][2], pi[1][2]);
        radf2_at_f(st, wa, 50 - i, pr[0][3], pi[0][3], pr[1][3], pi[1][3]);
    }

    // stage B, radf4's i = 1 case (fftpack.f90:859-866) for both k, feeding radf2's i = 1 (:729-732) and i = ido (:757-760)
    // cases and its butterfly 25
    template <int XS>
    static __device__ __forceinline__ void stageB_first(const double* T, double* Y, const double* wa) { stageB_first_f<XS>(T, StoreY<XS>{Y}, wa); }
    template <int XS, class St>
    static __device__ __forceinline__ void stageB_first_f(const double* T, const St& st, const double* wa) { stageB_first_g(LoadT<XS>{T}, st, wa); }
    template <class Ld, class St>
    static __device__ __forceinline__ void stageB_first_g(const Ld& ld, const St& st, const double* wa) {
        double a[2], d[2], yr[2], yi[2];                   // ch(1,1,k), ch(12,4,k), ch(12,2,k), ch(1,3,k)
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const double c1 = ld(k, 0), c2 = ld(k + 2, 0), c3 = ld(k + 4, 0), c4 = ld(k + 6, 0);
            const double tr1 = c2 + c4;
            const double tr2 = c1 + c3;
            a[k] = tr1 + tr2;
            d[k] = tr2 - tr1;
            yr[k] = c1 - c3;
            yi[k] = c4 - c2;
        }
        st(0, a[0] + a[1]);                                // radf2: ch(1,1,1)
        st(95, a[0] - a[1]);                               // ch(48,2,1)
        st(48, -d[1]);                                     // ch(1,2,1) = -cc(48,1,2)
        st(47, d[0]);                                      // ch(48,1,1) =  cc(48,1,1)
        radf2_at_f(st, wa, 25, yr[0], yi[0], yr[1], yi[1]);
    }

    // stage B, radf4's i = ido case (fftpack.f90:911-920) for both k, feeding radf2's butterflies 13 and 37
    template <int XS>
    static __device__ __forceinline__ void stageB_last(const double* T, double* Y, const double* wa) { stageB_last_f<XS>(T, StoreY<XS>{Y}, wa); }
    template <int XS, class St>
    static __device__ __forceinline__ void stageB_last_f(const double* T, const St& st, const double* wa) { stageB_last_g(LoadT<XS>{T}, st, wa); }
    template <class Ld, class St>
    static __device__ __forceinline__ void stageB_last_g(const Ld& ld, const St& st, const double* wa) {
        const double hsqt2 = (double)(.5f * sqrtf(2.f));   // .5*sqrt(2.) in real32 (fftpack.f90:857)
        double pr[2], pi[2], qr[2], qi[2];                 // ch(12,1,k), ch(1,2,k) -> positions 11, 12; ch(12,3,k), ch(1,4,k) -> 35, 36
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const double c1 = ld(k, 11), c2 = ld(k + 2, 11), c3 = ld(k + 4, 11), c4 = ld(k + 6, 11);
            const double ti1 = -hsqt2 * (c2 + c4);
            const double tr1 = hsqt2 * (c2 - c4);
            pr[k] = tr1 + c1;
            qr[k] = c1 - tr1;
            pi[k] = ti1 - c3;
            qi[k] = ti1 + c3;
        }
        radf2_at_f(st, wa, 13, pr[0], pi[0], pr[1], pi[1]);
        radf2_at_f(st, wa, 37, qr[0], qi[0], qr[1], qi[1]);
    }
};

}  // namespace spd
