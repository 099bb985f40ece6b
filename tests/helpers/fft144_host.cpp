// Host build of csrc/fft144.cuh (device qualifiers compiled away): checked bit for bit against the oracle's
// pass-by-pass rfftb1 at T47 without a GPU.  Test helper only.
#include <cmath>
#define __device__
#define __forceinline__ inline
#include "../../speedy.f90_b200/csrc/fft144.cuh"

extern "C" void fft144_backward(const double* c, const double* wa, double* out) {
    using namespace spd;
    double T[144];
    for (int kh = 0; kh < 2; kh++) {
        for (int i = 3; i <= 9; i += 2) Fft144::stage1_general<1>(c, T, wa, i, kh);
        Fft144::stage1_first<1>(c, T, wa, kh);
    }
    for (int k3 = 0; k3 < 16; k3++) {
        double y[9];
        Fft144::stage2<1>(T, wa, k3, y);
        for (int jj = 0; jj < 3; jj++)
            for (int j = 0; j < 3; j++) out[k3 + 16 * j + 48 * jj] = y[3 * jj + j];
    }
}
