// Host build of csrc/fft96.cuh (the device qualifiers compiled away) so that the regrouped FFTPACK passes can be
// checked bit for bit against the oracle's pass-by-pass transform without a GPU.  Test helper only.
#include <cmath>
#define __device__
#define __forceinline__ inline
#include "../../speedy.f90_b200/csrc/fft96.cuh"

extern "C" void fft96_backward(const double* c, const double* wa, double* out) {
    using namespace spd;
    double T[96];
    for (int i = 3; i <= 11; i += 2) Fft96::stage1_general<1>(c, T, wa, i);
    Fft96::stage1_first<1>(c, T, wa);
    Fft96::stage1_last<1>(c, T, wa);
    for (int k3 = 0; k3 < 8; k3++) {
        double y[12];
        Fft96::stage2<1>(T, wa, k3, y);
        for (int jj = 0; jj < 3; jj++)
            for (int j = 0; j < 4; j++) out[k3 + 8 * j + 32 * jj] = y[4 * jj + j];
    }
}
