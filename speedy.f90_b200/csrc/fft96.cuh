// Backward real FFT of length 96 — the reference's FFTPACK passes (fftpack.f90:69-134 rfftb1 with the factor list
// 2,4,4,3 of rffti1; :204 radb2, :328 radb4, :256 radb3) regrouped so that a thread keeps its data in registers across
// two passes:
//   stage 1 = radb2 (ido 48, l1 1) + radb4 (ido 12, l1 2): the 96 half-complex inputs fall into five closed sets of 16
//             reals (the general butterflies i = 3,5,..,11 of radb4 with the four radb2 butterflies that feed them) and
//             two sets of 8 (radb4's i = 1 and i = ido cases);
//   stage 2 = radb4 (ido 3, l1 8) + radb3 (ido 1, l1 32): eight closed sets of 12 contiguous reals, one per k of radb4.
// Every butterfly is the reference's, expression by expression, with the twiddle table of rffti1 (single-precision
// 2*pi seed, fftpack.f90:39): the result matches the reference's transform to rounding (fused multiply-adds aside),
// which a generic FFT does not, at 1/10 of the flops of the dense-operator form.
//
// Data sit in shared memory "position-major": element p of row r at X[p * XS + r], so that the 16 rows of a half-warp
// read and write 16 consecutive doubles.  X and T are already offset by the row.  wa is the 0-based table of rffti1.
#pragma once

namespace spd {

struct Fft96 {
    // one radb2 butterfly (fftpack.f90:223-236) at i1 (ic = 50 - i1): returns ch(i1-1,1,1), ch(i1,1,1), ch(i1-1,1,2), ch(i1,1,2)
    // functor forms (quad kernel, transforms_quad.cu): ld(pos) reads half-complex position pos of the row, st(blk, off, v) writes stage 1's
    // output at position 12 blk + off, sync() separates ALL loads of a stage from its stores (in-place transforms).  The pointer forms
    // below are the same code with X[pos * XS] / T[pos * XS] and no barrier.
    template <int XS>
    struct LoadX {
        const double* X;
        __device__ __forceinline__ double operator()(int pos) const { return X[pos * XS]; }
    };
    template <int XS>
    struct StoreT {
        double* T;
        __device__ __forceinline__ void operator()(int blk, int off, double v) const { T[(12 * blk + off) * XS] = v; }
    };
    struct NoSync { __device__ __forceinline__ void operator()() const {} };

    template <int XS>
    static __device__ __forceinline__ void radb2_at(const double* X, const double* wa, int i1, double& y1r, double& y1i, double& y2r, double& y2i) {
        radb2_at_g(LoadX<XS>{X}, wa, i1, y1r, y1i, y2r, y2i);
    }
    template <class Ld>
    static __device__ __forceinline__ void radb2_at_g(const Ld& ld, const double* wa, int i1, double& y1r, double& y1i, double& y2r, double& y2i) {
        const int ic = 50 - i1;
        const double ar = ld(i1 - 2), ai = ld(i1 - 1), br = ld(46 + ic), bi = ld(47 + ic);
        const double wr = wa[i1 - 3], wi = wa[i1 - 2];
        y1r = ar + br;
        const double tr2 = ar - br;
        y1i = ai - bi;
        const double ti2 = ai + bi;
        y2r = wr * tr2 - wi * ti2;
        y2i = wr * ti2 + wi * tr2;
    }

    // stage 1, general set of radb4's butterfly i (3,5,..,11), both k
    template <int XS>
    static __device__ __forceinline__ void stage1_general(const double* X, double* T, const double* wa, int i) { stage1_general_g(LoadX<XS>{X}, StoreT<XS>{T}, NoSync{}, wa, i); }
    template <class Ld, class St, class Sy>
    static __device__ __forceinline__ void stage1_general_g(const Ld& ld, const St& st, const Sy& sync, const double* wa, int i) {
        double v[16];
        stage1_general_ld(ld, i, v);
        sync();
        stage1_general_st(v, st, wa, i);
    }
    // the two halves of a set, for in-place transforms whose barrier must sit at a point where the whole warp converges:
    // _ld reads the set's inputs into v (general: 16, first: 8, last: 8 values), _st computes and stores
    template <class Ld>
    static __device__ __forceinline__ void stage1_general_ld(const Ld& ld, int i, double* v) {
        const int i1s[4] = {i, i + 24, 26 - i, 50 - i};
#pragma unroll
        for (int j = 0; j < 4; j++) { const int i1 = i1s[j]; v[4 * j] = ld(i1 - 2); v[4 * j + 1] = ld(i1 - 1); v[4 * j + 2] = ld(96 - i1); v[4 * j + 3] = ld(97 - i1); }
    }
    // radb2 butterfly (fftpack.f90:223-236) at i1 on loaded values (ar, ai, br, bi)
    static __device__ __forceinline__ void radb2_v(const double* v, const double* wa, int i1, double& y1r, double& y1i, double& y2r, double& y2i) {
        const double ar = v[0], ai = v[1], br = v[2], bi = v[3];
        const double wr = wa[i1 - 3], wi = wa[i1 - 2];
        y1r = ar + br;
        const double tr2 = ar - br;
        y1i = ai - bi;
        const double ti2 = ai + bi;
        y2r = wr * tr2 - wi * ti2;
        y2i = wr * ti2 + wi * tr2;
    }
    template <class St>
    static __device__ __forceinline__ void stage1_general_st(const double* v, const St& st, const double* wa, int i) {
        double y1r[4], y1i[4], y2r[4], y2i[4];     // [0] i, [1] i+24, [2] 26-i, [3] 50-i  <->  radb4's cc(.,1,k), cc(.,3,k), cc(ic..,2,k), cc(ic..,4,k)
        radb2_v(v, wa, i, y1r[0], y1i[0], y2r[0], y2i[0]);
        radb2_v(v + 4, wa, i + 24, y1r[1], y1i[1], y2r[1], y2i[1]);
        radb2_v(v + 8, wa, 26 - i, y1r[2], y1i[2], y2r[2], y2i[2]);
        radb2_v(v + 12, wa, 50 - i, y1r[3], y1i[3], y2r[3], y2i[3]);
        const double w1r = wa[45 + i], w1i = wa[46 + i], w2r = wa[57 + i], w2i = wa[58 + i], w3r = wa[69 + i], w3i = wa[70 + i];
        const int off = i - 2;
        auto radb4 = [&](const double* yr, const double* yi, int kb) {   // fftpack.f90:356-393; kb = k - 1: ch(i-1,k,j) at block kb + 2 (j-1), offset i - 2
            const double ti1 = yi[0] + yi[3];
            const double ti2 = yi[0] - yi[3];
            const double ti3 = yi[1] - yi[2];
            const double tr4 = yi[1] + yi[2];
            const double tr1 = yr[0] - yr[3];
            const double tr2 = yr[0] + yr[3];
            const double ti4 = yr[1] - yr[2];
            const double tr3 = yr[1] + yr[2];
            st(kb, off, tr2 + tr3);
            const double cr3 = tr2 - tr3;
            st(kb, off + 1, ti2 + ti3);
            const double ci3 = ti2 - ti3;
            const double cr2 = tr1 - tr4;
            const double cr4 = tr1 + tr4;
            const double ci2 = ti1 + ti4;
            const double ci4 = ti1 - ti4;
            st(kb + 2, off, w1r * cr2 - w1i * ci2);
            st(kb + 2, off + 1, w1r * ci2 + w1i * cr2);
            st(kb + 4, off, w2r * cr3 - w2i * ci3);
            st(kb + 4, off + 1, w2r * ci3 + w2i * cr3);
            st(kb + 6, off, w3r * cr4 - w3i * ci4);
            st(kb + 6, off + 1, w3r * ci4 + w3i * cr4);
        };
        radb4(y1r, y1i, 0);
        radb4(y2r, y2i, 1);
    }

    // stage 1, radb4's i = 1 case (fftpack.f90:343-353): fed by radb2's i = 1 and i = ido cases and its butterfly 25
    template <int XS>
    static __device__ __forceinline__ void stage1_first(const double* X, double* T, const double* wa) { stage1_first_g(LoadX<XS>{X}, StoreT<XS>{T}, NoSync{}, wa); }
    template <class Ld, class St, class Sy>
    static __device__ __forceinline__ void stage1_first_g(const Ld& ld, const St& st, const Sy& sync, const double* wa) {
        double v[8];
        stage1_first_ld(ld, v);
        sync();
        stage1_first_st(v, st, wa);
    }
    template <class Ld>
    static __device__ __forceinline__ void stage1_first_ld(const Ld& ld, double* v) {
        v[0] = ld(0); v[1] = ld(95); v[2] = ld(47); v[3] = ld(48);
        v[4] = ld(23); v[5] = ld(24); v[6] = ld(71); v[7] = ld(72);      // radb2's butterfly 25
    }
    template <class St>
    static __device__ __forceinline__ void stage1_first_st(const double* v, const St& st, const double* wa) {
        const double c0 = v[0], c95 = v[1], c47 = v[2], c48 = v[3];
        double a[2], d[2], yr[2], yi[2];
        a[0] = c0 + c95;                 // ch(1,1,1)
        a[1] = c0 - c95;                 // ch(1,1,2)
        d[0] = c47 + c47;                // ch(48,1,1)
        d[1] = -(c48 + c48);             // ch(48,1,2)
        radb2_v(v + 4, wa, 25, yr[0], yi[0], yr[1], yi[1]);
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const double tr1 = a[k] - d[k];
            const double tr2 = a[k] + d[k];
            const double tr3 = yr[k] + yr[k];
            const double tr4 = yi[k] + yi[k];
            st(k, 0, tr2 + tr3);
            st(k + 2, 0, tr1 - tr4);
            st(k + 4, 0, tr2 - tr3);
            st(k + 6, 0, tr1 + tr4);
        }
    }

    // stage 1, radb4's i = ido case (fftpack.f90:397-408): fed by radb2's butterflies 13 and 37
    template <int XS>
    static __device__ __forceinline__ void stage1_last(const double* X, double* T, const double* wa) { stage1_last_g(LoadX<XS>{X}, StoreT<XS>{T}, NoSync{}, wa); }
    template <class Ld, class St, class Sy>
    static __device__ __forceinline__ void stage1_last_g(const Ld& ld, const St& st, const Sy& sync, const double* wa) {
        double v[8];
        stage1_last_ld(ld, v);
        sync();
        stage1_last_st(v, st, wa);
    }
    template <class Ld>
    static __device__ __forceinline__ void stage1_last_ld(const Ld& ld, double* v) {
        v[0] = ld(11); v[1] = ld(12); v[2] = ld(83); v[3] = ld(84);      // radb2's butterfly 13
        v[4] = ld(35); v[5] = ld(36); v[6] = ld(59); v[7] = ld(60);      // radb2's butterfly 37
    }
    template <class St>
    static __device__ __forceinline__ void stage1_last_st(const double* v, const St& st, const double* wa) {
        const double sqrt2 = (double)sqrtf(2.f);           // sqrt(2.) in real32 (fftpack.f90:341)
        double pr[2], pi[2], qr[2], qi[2];
        radb2_v(v, wa, 13, pr[0], pi[0], pr[1], pi[1]);     // ch(12,1,.), ch(13,1,.)
        radb2_v(v + 4, wa, 37, qr[0], qi[0], qr[1], qi[1]); // ch(36,1,.), ch(37,1,.)
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const double ti1 = pi[k] + qi[k];
            const double ti2 = qi[k] - pi[k];
            const double tr1 = pr[k] - qr[k];
            const double tr2 = pr[k] + qr[k];
            st(k, 11, tr2 + tr2);
            st(k + 2, 11, sqrt2 * (tr1 - ti1));
            st(k + 4, 11, ti2 + ti2);
            st(k + 6, 11, -sqrt2 * (tr1 + ti1));
        }
    }

    // stage 2 for radb4's k = k3 + 1 (k3 = 0..7): y[4*(jj) + j] = grid point k3 + 8 j + 32 jj (j = 0..3, jj = 0..2)
    template <int XS>
    static __device__ __forceinline__ void stage2(const double* T, const double* wa, int k3, double (&y)[12]) {
        const double taur = -.5;
        const double taui = (double)(.5f * sqrtf(3.f));    // .5*sqrt(3.) in real32 (fftpack.f90:269)
        double e[12];
#pragma unroll
        for (int t = 0; t < 12; t++) e[t] = T[(12 * k3 + t) * XS];     // cc(i,j,k): e[(i-1) + 3 (j-1)]
        double z[3][4];                                                // ch(i,k,j)
        {   // i = 1 (fftpack.f90:343-353)
            const double tr1 = e[0] - e[11];
            const double tr2 = e[0] + e[11];
            const double tr3 = e[5] + e[5];
            const double tr4 = e[6] + e[6];
            z[0][0] = tr2 + tr3;
            z[0][1] = tr1 - tr4;
            z[0][2] = tr2 - tr3;
            z[0][3] = tr1 + tr4;
        }
        {   // i = 3, ic = 2 (fftpack.f90:356-393); twiddles wa(85 + ..) of the third pass
            const double w1r = wa[84], w1i = wa[85], w2r = wa[87], w2i = wa[88], w3r = wa[90], w3i = wa[91];
            const double ti1 = e[2] + e[10];
            const double ti2 = e[2] - e[10];
            const double ti3 = e[8] - e[4];
            const double tr4 = e[8] + e[4];
            const double tr1 = e[1] - e[9];
            const double tr2 = e[1] + e[9];
            const double ti4 = e[7] - e[3];
            const double tr3 = e[7] + e[3];
            z[1][0] = tr2 + tr3;
            const double cr3 = tr2 - tr3;
            z[2][0] = ti2 + ti3;
            const double ci3 = ti2 - ti3;
            const double cr2 = tr1 - tr4;
            const double cr4 = tr1 + tr4;
            const double ci2 = ti1 + ti4;
            const double ci4 = ti1 - ti4;
            z[1][1] = w1r * cr2 - w1i * ci2;
            z[2][1] = w1r * ci2 + w1i * cr2;
            z[1][2] = w2r * cr3 - w2i * ci3;
            z[2][2] = w2r * ci3 + w2i * cr3;
            z[1][3] = w3r * cr4 - w3i * ci4;
            z[2][3] = w3r * ci4 + w3i * cr4;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {   // radb3, ido = 1 (fftpack.f90:271-278) for k = k3 + 8 j + 1
            const double tr2 = z[1][j] + z[1][j];
            const double cr2 = z[0][j] + taur * tr2;
            y[j] = z[0][j] + tr2;
            const double ci3 = taui * (z[2][j] + z[2][j]);
            y[4 + j] = cr2 - ci3;
            y[8 + j] = cr2 + ci3;
        }
    }
};

}  // namespace spd
