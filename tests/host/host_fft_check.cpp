// CPU emulation of the 64-lane cooperative FFT in sarssl_b200/csrc/fft512.cuh against an O(N^2) double DFT.
// Build+run: g++ -O1 -I sarssl_b200/csrc tests/host/host_fft_check.cpp -o /tmp/host_fft_check && /tmp/host_fft_check
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <complex>
#include "fft512.cuh"
using namespace sarssl;
int main() {
    std::vector<float> x0(512), x1(512);
    srand(3);
    for (int n = 0; n < 512; ++n) { x0[n] = (rand() / (float)RAND_MAX) - 0.5f; x1[n] = (rand() / (float)RAND_MAX) - 0.5f; }
    static float sre[kFftPlane], sim[kFftPlane];
    FftLane lanes[64];
    float2 v[64][8];
    for (int l = 0; l < 64; ++l) lanes[l].init(l);
    for (int l = 0; l < 64; ++l) for (int r = 0; r < 8; ++r) { int n = l + 64 * r; v[l][r] = make_float2(x0[n] * lanes[l].win[r], x1[n] * lanes[l].win[r]); }
    for (int l = 0; l < 64; ++l) fft_pass1(v[l], lanes[l], sre, sim, l);
    for (int l = 0; l < 64; ++l) fft_pass2_load(v[l], sre, sim, l);
    for (int l = 0; l < 64; ++l) fft_pass2_store(v[l], lanes[l], sre, sim, l);
    for (int l = 0; l < 64; ++l) fft_pass3(v[l], sre, sim, l);
    for (int l = 0; l < 64; ++l) split_store(v[l], sre, sim, l);
    double maxerr = 0, maxref = 0;
    const double PI = 3.14159265358979323846;
    for (int t = 0; t < 64; ++t) for (int k3 = 0; k3 < 5; ++k3) {
        int k = t + 64 * k3; if (k > 256) continue;
        float4 o = split_bin(v[t], sre, sim, t, k3);
        std::complex<double> X0 = 0, X1 = 0;
        for (int n = 0; n < 512; ++n) {
            double w = 0.5 - 0.5 * cos(2 * PI * n / 512);
            std::complex<double> e = std::polar(1.0, -2 * PI * n * k / 512.0);
            X0 += (double)x0[n] * w * e; X1 += (double)x1[n] * w * e;
        }
        double e = fabs(o.x - X0.real()); e = fmax(e, fabs(o.z - X0.imag())); e = fmax(e, fabs(o.y - X1.real())); e = fmax(e, fabs(o.w - X1.imag()));
        maxerr = fmax(maxerr, e); maxref = fmax(maxref, std::abs(X0));
    }
    printf("max abs err %.3e (max |X| %.3f)\n", maxerr, maxref);
    return maxerr < 1e-4 * maxref ? 0 : 1;
}
