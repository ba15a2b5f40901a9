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

// ---- warp-per-transform variant (fft512w.cuh): emulate 32 lanes; the shuffle steps are re-stated from the documented algebra ----
#include "fft512w.cuh"
static int check_warp_fft() {
    std::vector<float> x0(512), x1(512);
    srand(11);
    for (int n = 0; n < 512; ++n) { x0[n] = (rand() / (float)RAND_MAX) - 0.5f; x1[n] = (rand() / (float)RAND_MAX) - 0.5f; }
    static float2 tb[kWTransFloat2];
    WarpFftLane lanes[32];
    float2 v[32][16];
    const double PI = 3.14159265358979323846;
    for (int l = 0; l < 32; ++l) lanes[l].init(l);
    for (int l = 0; l < 32; ++l) for (int n1 = 0; n1 < 16; ++n1) {
        int n = 32 * n1 + l; float w = (float)(0.5 - 0.5 * cos(2 * PI * n / 512));
        v[l][n1] = make_float2(x0[n] * w, x1[n] * w);
    }
    for (int l = 0; l < 32; ++l) wfft_stage1(v[l], lanes[l], tb, l);
    for (int l = 0; l < 32; ++l) wfft_stage2(v[l], tb, l);
    // combine (lane ^ 16): Z[k1 + 256 h + 16 s]
    float2 Z[32][16];
    for (int l = 0; l < 32; ++l) for (int s = 0; s < 16; ++s) {
        float2 mine = v[l][s], oth = v[l ^ 16][s];
        Z[l][s] = (l >= 16) ? make_float2(oth.x - mine.x, oth.y - mine.y) : make_float2(mine.x + oth.x, mine.y + oth.y);
    }
    double maxerr = 0, maxref = 0;
    for (int l = 0; l < 32; ++l) for (int s = 0; s < 16; ++s) {
        int k = (l & 15) + 256 * (l >> 4) + 16 * s;
        std::complex<double> X = 0;
        for (int n = 0; n < 512; ++n) {
            double w = 0.5 - 0.5 * cos(2 * PI * n / 512);
            X += std::complex<double>(x0[n] * w, x1[n] * w) * std::polar(1.0, -2 * PI * n * k / 512.0);
        }
        maxerr = fmax(maxerr, fmax(fabs(Z[l][s].x - X.real()), fabs(Z[l][s].y - X.imag()))); maxref = fmax(maxref, std::abs(X));
    }
    // natural-order Z buffer (wfft_store_z) and the split the kernel does from it: bins k = 1 + lane + 32 r against X_ch0 / X_ch1 computed directly
    static float2 zb[kWTransFloat2];
    for (int l = 0; l < 32; ++l) {
        float2 sgn = l < 16 ? make_float2(1.f, 1.f) : make_float2(-1.f, -1.f);
        wfft_store_z(v[l], v[l ^ 16], sgn, zb, l);
    }
    int bad = 0;
    for (int l = 0; l < 32; ++l) for (int s = 0; s < 16; ++s) {
        int k = (l & 15) + 256 * (l >> 4) + 16 * s;
        if (zb[wz_pos(k)].x != Z[l][s].x || zb[wz_pos(k)].y != Z[l][s].y) ++bad;
    }
    double maxsplit = 0;
    for (int k = 1; k <= 256; ++k) {
        float4 o = wfft_split2(zb[wz_pos(k)], zb[wz_pos(512 - k)]);
        std::complex<double> X0 = 0, X1 = 0;
        for (int n = 0; n < 512; ++n) {
            double w = 0.5 - 0.5 * cos(2 * PI * n / 512);
            std::complex<double> e = std::polar(1.0, -2 * PI * n * k / 512.0);
            X0 += x0[n] * w * e; X1 += x1[n] * w * e;
        }
        maxsplit = fmax(maxsplit, fmax(fmax(fabs(0.5 * o.x - X0.real()), fabs(0.5 * o.z - X0.imag())), fmax(fabs(0.5 * o.y - X1.real()), fabs(0.5 * o.w - X1.imag()))));
    }
    if (maxsplit > 1e-4 * maxref) ++bad;
    printf("warp fft: max abs err %.3e (max |Z| %.3f), split err %.3e, mismatches %d\n", maxerr, maxref, maxsplit, bad);
    return (maxerr < 1e-4 * maxref && bad == 0) ? 0 : 1;
}
struct RunWarpCheck { RunWarpCheck() { if (check_warp_fft()) { fprintf(stderr, "warp fft check FAILED\n"); exit(2); } } } run_warp_check_instance;
