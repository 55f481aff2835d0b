// host_tables.cpp -- exact host-side coefficient tables for the b200cv kernels.
// Compiled with -ffp-contract=off: every expression below must round exactly like the reference's
// softfloat (IEEE-754 binary64, round-to-nearest-even, no fused operations unless written as fma()).
//
// Reference behaviour restated here (not copied):
//   getGaussianKernelBitExact      modules/imgproc/src/smooth.dispatch.cpp:81-198
//   getGaussianKernelFixedPoint_ED modules/imgproc/src/smooth.dispatch.cpp:224-258
//   softdouble exp()               modules/core/src/softfloat.cpp:3429-3563 (table-driven 2^(k/64) * degree-5 polynomial)
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "host_tables.h"

namespace b200cv {

static inline double from_bits(uint64_t u) { double d; std::memcpy(&d, &u, 8); return d; }
static inline uint64_t to_bits(double d) { uint64_t u; std::memcpy(&u, &d, 8); return u; }

// 2^(i/64), i = 0..63, correctly rounded to binary64 (computed once in extended precision).
static const double* exp2_table64()
{
    static double tab[64];
    static bool init = false;
    if (!init) {
        for (int i = 0; i < 64; i++) tab[i] = (double)exp2l((long double)i / 64.0L);
        init = true;
    }
    return tab;
}

// exp() with the reference's exact operation sequence: x*(64/ln2) -> nearest integer n; 2^(n>>6) * 2^((n&63)/64) * P(frac)
double softdouble_exp(double x)
{
    if (std::isnan(x)) return std::nan("");
    if (std::isinf(x)) return x > 0 ? x : 0.0;
    const double C0 = from_bits(0x3f83ce0f3e46f431ULL);                 // common polynomial scale
    static const double A5 = 1.0 / C0,
                        A4 = from_bits(0x3fe62e42fefa39f1ULL) / C0,     // ln2
                        A3 = from_bits(0x3fcebfbdff82a45aULL) / C0,     // ~ln2^2/2
                        A2 = from_bits(0x3fac6b08d81fec75ULL) / C0,     // ~ln2^3/6
                        A1 = from_bits(0x3f83b2a72b4f3cd3ULL) / C0,     // minimax degree-4 coefficient
                        A0 = from_bits(0x3f55e7aa1566c2a4ULL) / C0;     // minimax degree-5 coefficient
    static const double prescale = from_bits(0x3ff71547652b82feULL) * 64.0;  // 64/ln2
    static const double postscale = 1.0 / 64.0;
    const double max_val = 3000.0 * 64.0;

    double x0;
    int e = (int)((to_bits(x) >> 52) & 0x7ff);
    if (e > 1023 + 10) x0 = (to_bits(x) >> 63) ? -max_val : max_val;
    else x0 = x * prescale;

    int val0 = (int)std::lrint(x0);   // round-half-even (default rounding mode)
    int t = (val0 >> 6) + 1023;
    t = t < 0 ? 0 : (t > 2047 ? 2047 : t);
    double buf = from_bits((uint64_t)t << 52);
    x0 = (x0 - std::nearbyint(x0)) * postscale;

    double p = A0 * x0;
    p = p + A1; p = p * x0;
    p = p + A2; p = p * x0;
    p = p + A3; p = p * x0;
    p = p + A4; p = p * x0;
    p = p + A5;
    double r = buf * C0;
    r = r * exp2_table64()[val0 & 63];
    r = r * p;
    return r;
}

void gaussian_kernel_bitexact(int n, double sigma, std::vector<double>& out)
{
    out.assign((size_t)n, 0.0);
    if (sigma <= 0) {
        static const double k3[] = {0.25, 0.5, 0.25};
        static const double k5[] = {0.0625, 0.25, 0.375, 0.25, 0.0625};
        static const double k7[] = {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125};
        static const double k9[] = {4 / 256., 13 / 256., 30 / 256., 51 / 256., 60 / 256., 51 / 256., 30 / 256., 13 / 256., 4 / 256.};
        const double* fixed = n == 1 ? nullptr : n == 3 ? k3 : n == 5 ? k5 : n == 7 ? k7 : n == 9 ? k9 : nullptr;
        if (n == 1) { out[0] = 1.0; return; }
        if (fixed) { for (int i = 0; i < n; i++) out[i] = fixed[i]; return; }
    }
    const double c015 = from_bits(0x3fc3333333333333ULL), c035 = from_bits(0x3fd6666666666666ULL);
    double sigmaX = sigma > 0 ? sigma : std::fma((double)n, c015, c035);   // the reference uses a fused mulAdd here
    double scale2X = -0.125 / (sigmaX * sigmaX);
    int n2 = (n - 1) / 2;
    std::vector<double> values((size_t)n2 + 1);
    double sum = 0.0;
    for (int i = 0, x = 1 - n; i < n2; i++, x += 2) {
        double t = softdouble_exp((double)(x * x) * scale2X);
        values[i] = t;
        sum = sum + t;
    }
    sum = sum * 2.0;
    sum = sum + 1.0;
    if ((n & 1) == 0) sum = sum + 1.0;
    double mul1 = 1.0 / sum;
    for (int i = 0; i < n2; i++) {
        double t = values[i] * mul1;
        out[i] = t;
        out[n - 1 - i] = t;
    }
    out[n2] = 1.0 * mul1;
    if ((n & 1) == 0) out[n2 + 1] = out[n2];
}

// error-diffusion rounding of the bit-exact taps to `bits` fractional bits; centre tap takes the remainder
void gaussian_kernel_fixed(int n, double sigma, int bits, std::vector<int64_t>& out)
{
    std::vector<double> k;
    gaussian_kernel_bitexact(n, sigma, k);
    out.assign((size_t)n, 0);
    const double mult = (double)((int64_t)1 << bits);
    int n2 = n / 2;
    double err = 0.0;
    int64_t sum = 0;
    for (int i = 0; i < n2; i++) {
        double adj = k[i] * mult;
        adj = adj + err;
        int64_t v0 = (int64_t)std::lrint(adj);
        err = adj - (double)v0;
        out[i] = v0;
        out[n - 1 - i] = v0;
        sum += v0;
    }
    out[n2] = ((int64_t)1 << bits) - 2 * sum;
}

// automatic kernel size from sigma: smooth.dispatch.cpp:288-291
int gaussian_auto_ksize(double sigma, bool is_u8)
{
    return ((int)std::lrint(sigma * (is_u8 ? 3 : 4) * 2 + 1)) | 1;
}

}  // namespace b200cv

// ---- cv::remap interpolation tables (reference behaviour: initInterTab2D, modules/imgproc/src/imgwarp.cpp:213-287) ----
// 32 x 32 sub-pixel positions; entry (iy*32+ix) holds the outer product of the 1-D tap vectors for fy=iy/32, fx=ix/32.
// Fixed-point version: taps scaled by 2^15 and rounded, then the rounding residue of the whole ksize x ksize block is
// pushed into the largest (residue < 0) or smallest (residue > 0) of the four central taps so that the block sums to 2^15.
namespace b200cv {

static void cubic_taps(float x, float* c)
{
    const float A = -0.75f;
    c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
    c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
    c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
    c[3] = 1.f - c[0] - c[1] - c[2];
}

// interpolateLanczos4 as cv::remap's tables use it (imgwarp.cpp:162-188): sin / cos of the first tap's angle in double, the other seven by
// the 45-degree rotation table, normalised in float; x < FLT_EPSILON is the unit tap
static void lanczos4_taps_remap(float x, float* c)
{
    static const double s45 = 0.70710678118654752440084436210485;
    static const double cs[][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
    if (x < 1.1920928955078125e-07f) {
        for (int i = 0; i < 8; i++) c[i] = 0;
        c[3] = 1;
        return;
    }
    float sum = 0;
    const double y0 = -(x + 3) * 3.1415926535897932384626433832795 * 0.25, s0 = std::sin(y0), c0 = std::cos(y0);
    for (int i = 0; i < 8; i++) {
        const double y = -(x + 3 - i) * 3.1415926535897932384626433832795 * 0.25;
        c[i] = (float)((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
        sum += c[i];
    }
    sum = 1.f / sum;
    for (int i = 0; i < 8; i++) c[i] *= sum;
}

static void inter_tab_2d(int ksize, std::vector<float>& ftab, std::vector<short>& itab)
{
    const int N = 32;
    std::vector<float> t1(N * ksize);
    const float step = 1.f / N;
    for (int i = 0; i < N; i++) {
        float x = i * step;
        if (ksize == 2) { t1[i * 2] = 1.f - x; t1[i * 2 + 1] = x; }
        else if (ksize == 4) cubic_taps(x, &t1[i * 4]);
        else lanczos4_taps_remap(x, &t1[i * 8]);
    }
    const int kk = ksize * ksize, c0 = ksize / 2;
    ftab.assign((size_t)N * N * kk, 0.f);
    itab.assign((size_t)N * N * kk, 0);
    for (int iy = 0; iy < N; iy++)
        for (int ix = 0; ix < N; ix++) {
            float* f = &ftab[(size_t)(iy * N + ix) * kk];
            short* q = &itab[(size_t)(iy * N + ix) * kk];
            int total = 0;
            for (int a = 0; a < ksize; a++)
                for (int b = 0; b < ksize; b++) {
                    float v = t1[iy * ksize + a] * t1[ix * ksize + b];
                    f[a * ksize + b] = v;
                    long r = lrintf(v * 32768.f);
                    r = r < -32768 ? -32768 : r > 32767 ? 32767 : r;
                    q[a * ksize + b] = (short)r;
                    total += (int)r;
                }
            int residue = total - 32768;
            if (residue != 0 && ksize == 2) {
                // 2x2 block: the reference's scan starts at the last tap and only ever meets zeros beyond the block,
                // so the residue (only -1 at the all-integer position, where 32768 saturates to 32767) lands on tap 3
                q[3] = (short)(q[3] - residue);
            } else if (residue != 0) {
                int lo = c0 * ksize + c0, hi = lo;          // scan taps (2..3, 2..3) in row-major order
                for (int a = c0; a < c0 + 2; a++)
                    for (int b = c0; b < c0 + 2; b++) {
                        int idx = a * ksize + b;
                        if (q[idx] < q[lo]) lo = idx;
                        else if (q[idx] > q[hi]) hi = idx;
                    }
                int tgt = residue < 0 ? hi : lo;
                q[tgt] = (short)(q[tgt] - residue);
            }
        }
}

void bilinear_tab(std::vector<float>& f, std::vector<short>& i) { inter_tab_2d(2, f, i); }
void bicubic_tab(std::vector<float>& f, std::vector<short>& i) { inter_tab_2d(4, f, i); }
void lanczos4_tab(std::vector<float>& f, std::vector<short>& i) { inter_tab_2d(8, f, i); }

}  // namespace b200cv
