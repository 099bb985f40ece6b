// Host build of csrc/fft96f.cuh (device qualifiers compiled away): the regrouped forward FFTPACK passes, checked bit for
// bit against the oracle's pass-by-pass rfftf1 without a GPU.  Test helper only.
#include <cmath>
#include <cstddef>
#define __device__
#define __forceinline__ inline
#include "../../speedy.f90_b200/csrc/fft96f.cuh"

extern "C" void fft96_forward(const double* g, const double* wa, double* out) {
    using namespace spd;
    double T[96];
    for (int k3 = 0; k3 < 8; k3++) {
        double x[12];
        for (int jj = 0; jj < 3; jj++)
            for (int j = 0; j < 4; j++) x[4 * jj + j] = g[k3 + 8 * j + 32 * jj];
        Fft96F::stageA<1>(x, T, wa, k3);
    }
    for (int i = 3; i <= 11; i += 2) Fft96F::stageB_general<1>(T, out, wa, i);
    Fft96F::stageB_first<1>(T, out, wa);
    Fft96F::stageB_last<1>(T, out, wa);
}
