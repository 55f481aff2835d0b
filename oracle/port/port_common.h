/* oracle/port/port_common.h -- shared helpers of the plain-C restatement ("port") of the reference's CPU hot path.
 *
 * TEST INFRASTRUCTURE ONLY: this code is the checker the CUDA path is compared against (tests/, smoke(),
 * bench.py's cpu_baseline leg).  Nothing under opencv_b200/ links or calls it.
 * Parity status: PINNED -- every function here is itself checked against (a) the reference's own golden
 * vectors / known-answer hashes (tests/golden/, tests/test_oracle_*.py) and (b) the unmodified reference built
 * from /root/reference by oracle/build_ref.py when that library is present.
 */
#ifndef PORT_COMMON_H
#define PORT_COMMON_H
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PORT_API __attribute__((visibility("default")))

typedef unsigned char uchar;

enum { P_8U = 0, P_16S = 3, P_32F = 5 };
#define P_DEPTH(t) ((t) & 7)
#define P_CN(t) ((((t) >> 3) & 511) + 1)

enum { PB_CONSTANT = 0, PB_REPLICATE = 1, PB_REFLECT = 2, PB_WRAP = 3, PB_REFLECT_101 = 4 };

/* cv::borderInterpolate -- modules/core/src/copy.cpp:748-793 */
static inline int port_border(int p, int len, int type)
{
    if ((unsigned)p < (unsigned)len) return p;
    type &= ~16;
    if (type == PB_REPLICATE) return p < 0 ? 0 : len - 1;
    if (type == PB_REFLECT || type == PB_REFLECT_101) {
        int delta = type == PB_REFLECT_101;
        if (len == 1) return 0;
        do { p = p < 0 ? -p - 1 + delta : len - 1 - (p - len) - delta; } while ((unsigned)p >= (unsigned)len);
        return p;
    }
    if (type == PB_WRAP) {
        if (p < 0) p -= ((p - len + 1) / len) * len;
        if (p >= len) p %= len;
        return p;
    }
    return -1;
}

/* saturate_cast family -- modules/core/include/opencv2/core/saturate.hpp:103-133; cvRound = round-half-even */
static inline int port_round(double v) { return (int)lrint(v); }
static inline uchar port_sat_u8i(int v) { return (uchar)(v < 0 ? 0 : v > 255 ? 255 : v); }
static inline uchar port_sat_u8f(float v) { return port_sat_u8i((int)lrintf(v)); }
static inline short port_sat_s16i(int v) { return (short)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }
static inline short port_sat_s16f(float v) { return port_sat_s16i((int)lrintf(v)); }

static inline size_t port_esz(int depth) { return depth == P_8U ? 1 : depth == P_16S ? 2 : 4; }

/* gaussian taps (port_gauss.c) */
void port_gaussian_taps(int n, double sigma, double* out);
void port_gaussian_taps_fixed(int n, double sigma, int bits, long long* out);
int port_sep_filter_core(const void* src, size_t sstep, void* dst, size_t dstep, int w, int h, int stype, int ddepth,
                         const float* kx, int nx, const float* ky, int ny, int ax, int ay, double delta, int border);
#endif
