/* oracle/port/port_gauss.c -- Gaussian taps exactly as the reference derives them (TEST INFRASTRUCTURE ONLY).
 *   getGaussianKernelBitExact       modules/imgproc/src/smooth.dispatch.cpp:81-198
 *   getGaussianKernelFixedPoint_ED  modules/imgproc/src/smooth.dispatch.cpp:224-258
 *   softdouble exp                  modules/core/src/softfloat.cpp:3429-3563
 * Native IEEE doubles reproduce softdouble bit for bit as long as nothing is contracted: build with -ffp-contract=off.
 */
#include "port_common.h"

static double bits2d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }

static double sd_exp(double x)
{
    static double tab[64];
    static int init = 0;
    if (!init) { for (int i = 0; i < 64; i++) tab[i] = (double)powl(2.0L, (long double)i / 64.0L); init = 1; }
    const double c = bits2d(0x3f83ce0f3e46f431ULL);
    const double a5 = 1.0 / c, a4 = bits2d(0x3fe62e42fefa39f1ULL) / c, a3 = bits2d(0x3fcebfbdff82a45aULL) / c,
                 a2 = bits2d(0x3fac6b08d81fec75ULL) / c, a1 = bits2d(0x3f83b2a72b4f3cd3ULL) / c, a0 = bits2d(0x3f55e7aa1566c2a4ULL) / c;
    const double pre = bits2d(0x3ff71547652b82feULL) * 64.0, post = 1.0 / 64.0;
    if (isnan(x)) return x;
    if (isinf(x)) return x > 0 ? x : 0.0;
    uint64_t u; memcpy(&u, &x, 8);
    double x0 = ((int)((u >> 52) & 0x7ff) > 1033) ? ((u >> 63) ? -192000.0 : 192000.0) : x * pre;
    int v = (int)lrint(x0);
    int t = (v >> 6) + 1023;
    if (t < 0) t = 0; if (t > 2047) t = 2047;
    double scale2 = bits2d((uint64_t)t << 52);
    x0 = (x0 - nearbyint(x0)) * post;
    double p = a0 * x0; p += a1; p *= x0; p += a2; p *= x0; p += a3; p *= x0; p += a4; p *= x0; p += a5;
    double r = scale2 * c; r *= tab[v & 63]; r *= p;
    return r;
}

void port_gaussian_taps(int n, double sigma, double* out)
{
    if (sigma <= 0 && (n == 1 || n == 3 || n == 5 || n == 7 || n == 9)) {
        static const double t3[] = {.25, .5, .25}, t5[] = {.0625, .25, .375, .25, .0625},
                            t7[] = {.03125, .109375, .21875, .28125, .21875, .109375, .03125},
                            t9[] = {4 / 256., 13 / 256., 30 / 256., 51 / 256., 60 / 256., 51 / 256., 30 / 256., 13 / 256., 4 / 256.};
        static const double t1[] = {1.0};
        const double* t = n == 1 ? t1 : n == 3 ? t3 : n == 5 ? t5 : n == 7 ? t7 : t9;
        memcpy(out, t, n * sizeof(double));
        return;
    }
    double sig = sigma > 0 ? sigma : fma((double)n, bits2d(0x3fc3333333333333ULL), bits2d(0x3fd6666666666666ULL));
    double s2 = -0.125 / (sig * sig);
    int half = (n - 1) / 2;
    double sum = 0;
    for (int i = 0, x = 1 - n; i < half; i++, x += 2) { out[i] = sd_exp((double)(x * x) * s2); sum += out[i]; }
    sum *= 2.0; sum += 1.0;
    if (!(n & 1)) sum += 1.0;
    double inv = 1.0 / sum;
    for (int i = 0; i < half; i++) { out[i] *= inv; out[n - 1 - i] = out[i]; }
    out[half] = inv;
    if (!(n & 1)) out[half + 1] = inv;
}

PORT_API void port_gaussian_taps_fixed(int n, double sigma, int bits, long long* out)
{
    double* k = (double*)malloc(sizeof(double) * n);
    port_gaussian_taps(n, sigma, k);
    double one = (double)(1LL << bits), err = 0;
    long long sum = 0;
    for (int i = 0; i < n / 2; i++) {
        double a = k[i] * one; a += err;
        long long v = llrint(a);
        err = a - (double)v;
        out[i] = out[n - 1 - i] = v;
        sum += v;
    }
    out[n / 2] = (1LL << bits) - 2 * sum;
    free(k);
}

PORT_API int port_gaussian_kernel(int n, double sigma, int ktype, void* out)
{
    double* k = (double*)malloc(sizeof(double) * n);
    port_gaussian_taps(n, sigma, k);
    if (ktype == 6) memcpy(out, k, sizeof(double) * n);
    else for (int i = 0; i < n; i++) ((float*)out)[i] = (float)k[i];
    free(k);
    return 0;
}
