/* oracle/port/port_filter.c -- GaussianBlur / sepFilter2D / filter2D / Sobel restated in scalar C.
 * TEST INFRASTRUCTURE ONLY (see port_common.h).
 *   cv::GaussianBlur          modules/imgproc/src/smooth.dispatch.cpp:609-826
 *   fixedSmoothInvoker (u8)   modules/imgproc/src/smooth.simd.hpp:1925-2197  == the evaluator of
 *                             modules/imgproc/test/test_smooth_bitexact.cpp:40-53
 *   createSeparableLinearFilter modules/imgproc/src/filter.dispatch.cpp:305-383 (bit-exact int mode :334-362)
 *   RowFilter / ColumnFilter  modules/imgproc/src/filter.simd.hpp:2386-2447, 2580-2757
 *   Filter2D                  modules/imgproc/src/filter.simd.hpp:3103-3175
 *   cv::Sobel                 modules/imgproc/src/deriv.cpp:414-465, getSobelKernels :96-160
 */
#include "port_common.h"

static float load_f(const void* row, int depth, int idx)
{
    return depth == P_8U ? (float)((const uchar*)row)[idx] : depth == P_16S ? (float)((const short*)row)[idx] : ((const float*)row)[idx];
}

static void store_f(void* row, int depth, int idx, float v)
{
    if (depth == P_8U) ((uchar*)row)[idx] = port_sat_u8f(v);
    else if (depth == P_16S) ((short*)row)[idx] = port_sat_s16f(v);
    else ((float*)row)[idx] = v;
}

/* kernel classification: getKernelType, filter.dispatch.cpp:225-259 */
enum { KT_SYMM = 1, KT_ASYMM = 2, KT_SMOOTH = 4, KT_INT = 8 };
static int ktype_of(const float* k, int n, int anchor)
{
    int t = KT_SMOOTH | KT_INT;
    double sum = 0;
    if (anchor * 2 + 1 == n) t |= KT_SYMM | KT_ASYMM;
    for (int i = 0; i < n; i++) {
        double a = k[i], b = k[n - 1 - i];
        if (a != b) t &= ~KT_SYMM;
        if (a != -b) t &= ~KT_ASYMM;
        if (a < 0) t &= ~KT_SMOOTH;
        if (a != (double)port_round(a)) t &= ~KT_INT;
        sum += a;
    }
    if (fabs(sum - 1) > 1.1920928955078125e-07 * (fabs(sum) + 1)) t &= ~KT_SMOOTH;
    return t;
}

static int exact_int_kernel(const float* k, int n, int bits, int* out)
{
    double eps = 10 * 1.1920928955078125e-07 * (1 << bits);
    for (int i = 0; i < n; i++) {
        double a = (double)k[i] * (1 << bits);
        out[i] = port_round(a);
        if (fabs(a - out[i]) > eps) return 0;
    }
    return 1;
}

/* integer separable pass: dst = cast((sum_j ky[j] * sum_i kx[i]*src + delta + round) >> shift) */
static void sep_int(const uchar* src, size_t sstep, void* dst, size_t dstep, int w, int h, int cn, int ddepth,
                    const long long* kx, int nx, const long long* ky, int ny, int ax, int ay, long long delta, int shift, int border,
                    int even_limit)
{
    long long rnd = shift ? (1LL << (shift - 1)) : 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            for (int c = 0; c < cn; c++) {
                long long acc = 0;
                for (int j = 0; j < ny; j++) {
                    int sy = port_border(y + j - ay, h, border);
                    if (sy < 0) continue;
                    const uchar* row = src + (size_t)sy * sstep;
                    long long line = 0;
                    for (int i = 0; i < nx; i++) {
                        int sx = port_border(x + i - ax, w, border);
                        if (sx >= 0) line += kx[i] * row[sx * cn + c];
                    }
                    acc += ky[j] * line;
                }
                long long v = (acc + delta + rnd) >> shift;
                if (x * cn + c < even_limit) {      /* vector body of SymmColumnVec_32s8u rounds the exact value half-to-even */
                    long long t = acc + delta, q = t >> shift, r = t & ((1LL << shift) - 1), half = 1LL << (shift - 1);
                    v = q + (r > half || (r == half && (q & 1)));
                }
                if (ddepth == P_8U) ((uchar*)((char*)dst + (size_t)y * dstep))[x * cn + c] = (uchar)(v < 0 ? 0 : v > 255 ? 255 : v);
                else ((short*)((char*)dst + (size_t)y * dstep))[x * cn + c] = (short)(v < -32768 ? -32768 : v > 32767 ? 32767 : v);
            }
}

static float mul_rn(float a, float b) { volatile float p = a * b; return p; }   /* a product rounded to float on its own */

/* float separable pass, in the operation order of the reference's AVX2 objects (the ones its dispatcher picks on any AVX2 host):
   rows     tap order, s = fma(x[i], kx[i], s) from 0            RowVec_32f, filter.simd.hpp:1632-1650
            row_mode 1|2: float source, 3 or 5 (anti)symmetric taps, centre-out
              k=3: fma(x0, k0, (x-1 + x1)*k1)   k=5: fma(x-2 + x2, k2, <k=3 expression>)   (differences and no centre for mode 2)
                                                                 SymmRowSmallVec_32f, :1768-1844
   columns  col_mode 1|2: (anti)symmetric odd kernel, mirrored rows are added first:
              s = fma(ky[c], S[c], delta); s = fma(ky[c+k], S[c+k] +/- S[c-k], s)
                                                                 SymmColumnVec_32f / _32f8u, :1878-1949, :1158-1202
            col_mode 0: any other kernel runs the scalar ColumnFilter, which is not contracted:
              s = fma(ky[0], S[0], delta); s += round(ky[j]*S[j])   :2590-2640 (as compiled: only the delta term is fused)
   The last (w*cn mod 8) elements of a float-source row (mod 32 for an 8-bit source feeding a float intermediate) come from the reference's scalar remainder loops, whose rounding depends on
   how its compiler contracted each of them; this port uses the formulas above there too (differences <= 1 ulp, tests/ mask them). */
static void sep_float(const void* src, size_t sstep, void* dst, size_t dstep, int w, int h, int cn, int sdepth, int ddepth,
                      const float* kx, int nx, const float* ky, int ny, int ax, int ay, float delta, int border, int col_mode, int row_mode)
{
    int we = w * cn;
    float* mid = (float*)malloc(sizeof(float) * (size_t)we * h);
    for (int y = 0; y < h; y++) {
        const void* row = (const char*)src + (size_t)y * sstep;
        for (int x = 0; x < w; x++)
            for (int c = 0; c < cn; c++) {
                float s = 0.f;
                if (row_mode) {
                    float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
                    for (int i = 0; i < nx; i++) {
                        int sx = port_border(x + i - ax, w, border);
                        v[i] = sx < 0 ? 0.f : load_f(row, sdepth, sx * cn + c);
                    }
                    int m = nx / 2;
                    if (row_mode == 1) {
                        s = fmaf(v[m], kx[m], mul_rn(v[m - 1] + v[m + 1], kx[m + 1]));
                        if (nx == 5) s = fmaf(v[m - 2] + v[m + 2], kx[m + 2], s);
                    } else {
                        s = mul_rn(v[m + 1] - v[m - 1], kx[m + 1]);
                        if (nx == 5) s = fmaf(v[m + 2] - v[m - 2], kx[m + 2], s);
                    }
                } else
                for (int i = 0; i < nx; i++) {
                    int sx = port_border(x + i - ax, w, border);
                    float v = sx < 0 ? 0.f : load_f(row, sdepth, sx * cn + c);
                    s = fmaf(v, kx[i], s);
                }
                mid[(size_t)y * we + x * cn + c] = s;
            }
    }
    for (int y = 0; y < h; y++) {
        void* drow = (char*)dst + (size_t)y * dstep;
        for (int e = 0; e < we; e++) {
            float s = delta;
            if (col_mode) {
                int c = ny / 2;
                if (col_mode == 1) {
                    int sy = port_border(y, h, border);
                    s = fmaf(ky[c], sy < 0 ? 0.f : mid[(size_t)sy * we + e], delta);
                }
                for (int k = 1; k <= c; k++) {
                    int s0 = port_border(y + k, h, border), s1 = port_border(y - k, h, border);
                    float a = s0 < 0 ? 0.f : mid[(size_t)s0 * we + e], b = s1 < 0 ? 0.f : mid[(size_t)s1 * we + e];
                    s = fmaf(ky[c + k], col_mode == 1 ? a + b : a - b, s);
                }
            } else {
                for (int j = 0; j < ny; j++) {
                    int sy = port_border(y + j - ay, h, border);
                    float v = sy < 0 ? 0.f : mid[(size_t)sy * we + e];
                    s = j == 0 ? fmaf(v, ky[0], delta) : s + mul_rn(v, ky[j]);
                }
            }
            store_f(drow, ddepth, e, s);
        }
    }
    free(mid);
}

int port_sep_filter_core(const void* src, size_t sstep, void* dst, size_t dstep, int w, int h, int stype, int ddepth,
                         const float* kx, int nx, const float* ky, int ny, int ax, int ay, double delta, int border)
{
    int sdepth = P_DEPTH(stype), cn = P_CN(stype);
    if (ddepth < 0) ddepth = sdepth;
    if (ax < 0) ax = nx / 2;
    if (ay < 0) ay = ny / 2;
    border &= ~16;
    if (sdepth == P_8U && (ddepth == P_8U || ddepth == P_16S)) {
        int rt = ktype_of(kx, nx, ax), ct = ktype_of(ky, ny, ay);
        int smooth8 = ddepth == P_8U && rt == (KT_SMOOTH | KT_SYMM) && ct == (KT_SMOOTH | KT_SYMM);
        int int16 = ddepth == P_16S && (rt & (KT_SYMM | KT_ASYMM)) && (ct & (KT_SYMM | KT_ASYMM)) && (rt & ct & KT_INT);
        if (smooth8 || int16) {
            int bits = ddepth == P_8U ? 8 : 0;
            int ikx[64], iky[64];
            if (nx <= 64 && ny <= 64 && exact_int_kernel(kx, nx, bits, ikx) && exact_int_kernel(ky, ny, bits, iky)) {
                long long lkx[64], lky[64];
                for (int i = 0; i < nx; i++) lkx[i] = ikx[i];
                for (int i = 0; i < ny; i++) lky[i] = iky[i];
                double dd = delta * (double)(1 << (2 * bits));
                long long di = llrint(dd);
                if (di > 2147483647LL) di = 2147483647LL; if (di < -2147483648LL) di = -2147483648LL;
                /* SymmColumnVec_32s8u (filter.simd.hpp:1011-1100): the first floor16(w*cn) elements of a row go through a float
                   vector body that rounds half-to-even; the scalar tail uses FixedPtCastEx (half-up), :2937-2946 */
                int even_limit = (bits && ny > 1) ? ((w * cn) / 16) * 16 : 0;
                sep_int((const uchar*)src, sstep, dst, dstep, w, h, cn, ddepth, lkx, nx, lky, ny, ax, ay, di, 2 * bits, border, even_limit);
                return 0;
            }
        }
    }
    int col_mode = 0, row_mode = 0;
    if ((ny & 1) && ay == ny / 2) {
        int ct = ktype_of(ky, ny, ay);
        col_mode = (ct & KT_SYMM) ? 1 : (ct & KT_ASYMM) ? 2 : 0;
    }
    if (sdepth == P_32F && (nx == 3 || nx == 5) && ax == nx / 2) {
        int rt = ktype_of(kx, nx, ax);
        row_mode = (rt & KT_SYMM) ? 1 : (rt & KT_ASYMM) ? 2 : 0;
    }
    sep_float(src, sstep, dst, dstep, w, h, cn, sdepth, ddepth, kx, nx, ky, ny, ax, ay, (float)delta, border, col_mode, row_mode);
    return 0;
}

PORT_API int port_sep_filter2d(const void* src, size_t sstep, void* dst, size_t dstep, int w, int h, int stype, int ddepth,
                               const float* kx, int nx, const float* ky, int ny, int ax, int ay, double delta, int border)
{
    return port_sep_filter_core(src, sstep, dst, dstep, w, h, stype, ddepth, kx, nx, ky, ny, ax, ay, delta, border);
}

PORT_API int port_gaussian_blur(const void* src, size_t sstep, void* dst, size_t dstep, int w, int h, int type,
                                int kw, int kh, double s1, double s2, int border)
{
    int depth = P_DEPTH(type), cn = P_CN(type);
    int b = border & ~16;
    if (b != PB_CONSTANT) { if (h == 1) kh = 1; if (w == 1) kw = 1; }
    if (kw == 1 && kh == 1) {
        for (int y = 0; y < h; y++) memcpy((char*)dst + (size_t)y * dstep, (const char*)src + (size_t)y * sstep, (size_t)w * cn * (depth == 2 ? 2 : port_esz(depth)));
        return 0;
    }
    if (s2 <= 0) s2 = s1;
    if (kw <= 0 && s1 > 0) kw = port_round(s1 * (depth == P_8U ? 3 : 4) * 2 + 1) | 1;
    if (kh <= 0 && s2 > 0) kh = port_round(s2 * (depth == P_8U ? 3 : 4) * 2 + 1) | 1;
    if (kw <= 0 || kh <= 0 || !(kw & 1) || !(kh & 1) || kw > 255 || kh > 255) return -1;
    if (s1 < 0) s1 = 0; if (s2 < 0) s2 = 0;
    if (depth == 2) {
        /* CV_16U: the 16.16 fixed-point path (fixedSmoothInvoker<uint16_t, ufixedpoint32>, smooth.simd.hpp:1925-2197; fixedpoint.inl.hpp:...):
         * rows   H = sum tap_x * p   in 32 bits (ufixedpoint32 * uint16 and + saturate; taps sum to 2^16, so nothing saturates for real taps)
         * columns   sum tap_y * H   in 64 bits (32.32), result (v + 2^31) >> 32 saturated to 16 bits */
        long long fx[256], fy[256];
        port_gaussian_taps_fixed(kw, s1, 16, fx);
        port_gaussian_taps_fixed(kh, s2, 16, fy);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
                for (int c = 0; c < cn; c++) {
                    unsigned long long acc = 0;
                    for (int j = 0; j < kh; j++) {
                        int sy = port_border(y + j - kh / 2, h, b);
                        if (sy < 0) continue;
                        const unsigned short* row = (const unsigned short*)((const char*)src + (size_t)sy * sstep);
                        unsigned long long line = 0;
                        for (int i = 0; i < kw; i++) {
                            int sx = port_border(x + i - kw / 2, w, b);
                            if (sx < 0) continue;
                            unsigned long long pr = (unsigned long long)fx[i] * row[sx * cn + c];
                            if (pr > 0xFFFFFFFFull) pr = 0xFFFFFFFFull;
                            line += pr; if (line > 0xFFFFFFFFull) line = 0xFFFFFFFFull;
                        }
                        unsigned long long pr = (unsigned long long)fy[j] * line, s = acc + pr;
                        acc = s < acc ? ~0ull : s;
                    }
                    unsigned long long r = (acc + (1ull << 31)) >> 32;
                    ((unsigned short*)((char*)dst + (size_t)y * dstep))[x * cn + c] = (unsigned short)(r > 65535 ? 65535 : r);
                }
        return 0;
    }
    if (depth == P_8U) {
        long long fx[256], fy[256];
        port_gaussian_taps_fixed(kw, s1, 8, fx);
        port_gaussian_taps_fixed(kh, s2, 8, fy);
        sep_int((const uchar*)src, sstep, dst, dstep, w, h, cn, P_8U, fx, kw, fy, kh, kw / 2, kh / 2, 0, 16, b, 0);
        return 0;
    }
    double dx[256], dy[256];
    float kx[256], ky[256];
    port_gaussian_taps(kw, s1, dx);
    port_gaussian_taps(kh, s2, dy);
    for (int i = 0; i < kw; i++) kx[i] = (float)dx[i];
    for (int i = 0; i < kh; i++) ky[i] = (float)dy[i];
    return port_sep_filter_core(src, sstep, dst, dstep, w, h, type, depth, kx, kw, ky, kh, kw / 2, kh / 2, 0.0, b);
}

PORT_API int port_filter2d(const void* src, size_t sstep, void* dst, size_t dstep, int w, int h, int stype, int ddepth,
                           const float* k, int kw, int kh, int ax, int ay, double delta, int border)
{
    int sdepth = P_DEPTH(stype), cn = P_CN(stype);
    if (ddepth < 0) ddepth = sdepth;
    if (ax < 0) ax = kw / 2;
    if (ay < 0) ay = kh / 2;
    border &= ~16;
    for (int y = 0; y < h; y++) {
        void* drow = (char*)dst + (size_t)y * dstep;
        for (int x = 0; x < w; x++)
            for (int c = 0; c < cn; c++) {
                float s = (float)delta;
                for (int j = 0; j < kh; j++) {
                    int sy = port_border(y + j - ay, h, border);
                    const void* row = sy < 0 ? NULL : (const char*)src + (size_t)sy * sstep;
                    for (int i = 0; i < kw; i++) {
                        float kv = k[j * kw + i];
                        if (kv == 0) continue;
                        int sx = port_border(x + i - ax, w, border);
                        float v = (row && sx >= 0) ? load_f(row, sdepth, sx * cn + c) : 0.f;
                        /* 8-bit source, float destination: the reference runs the scalar Filter2D<uchar, Cast<float,float>, FilterNoVec>, whose
                           products are rounded before they are added (filter.simd.hpp:3160-3172); every other pair has a vector body with FMA */
                        s = (sdepth == P_8U && ddepth == P_32F) ? s + mul_rn(kv, v) : fmaf(kv, v, s);
                    }
                }
                store_f(drow, ddepth, x * cn + c, s);
            }
    }
    return 0;
}

/* (1+z)^(n-d-1) (z-1)^d */
static void sobel_taps_1d(int n, int d, float* out)
{
    long long p[40] = {1};
    int len = 1;
    if (n > 1) {
        for (int i = 0; i < n - d - 1; i++) { p[len] = 0; for (int j = len; j > 0; j--) p[j] += p[j - 1]; len++; }
        for (int i = 0; i < d; i++) { p[len] = 0; for (int j = len; j > 0; j--) p[j] = p[j - 1] - p[j]; p[0] = -p[0]; len++; }
    }
    for (int i = 0; i < len; i++) out[i] = (float)p[i];
}

PORT_API int port_sobel(const void* src, size_t sstep, void* dst, size_t dstep, int w, int h, int stype, int ddepth,
                        int dx, int dy, int ksize, double scale, double delta, int border)
{
    float kx[40], ky[40];
    int nx, ny;
    if (ksize == -1) {
        static const float d[] = {-1, 0, 1}, s[] = {3, 10, 3};
        memcpy(kx, dx ? d : s, sizeof(d)); memcpy(ky, dy ? d : s, sizeof(d));
        nx = ny = 3;
    } else {
        nx = ny = ksize;
        if (nx == 1 && dx > 0) nx = 3;
        if (ny == 1 && dy > 0) ny = 3;
        sobel_taps_1d(nx, dx, kx);
        sobel_taps_1d(ny, dy, ky);
    }
    if (scale != 1) {
        float fs = (float)scale;
        float* k = dx == 0 ? kx : ky;
        int n = dx == 0 ? nx : ny;
        for (int i = 0; i < n; i++) k[i] = k[i] * fs;
    }
    return port_sep_filter_core(src, sstep, dst, dstep, w, h, stype, ddepth, kx, nx, ky, ny, -1, -1, delta, border);
}

/* cv::integral, 8UC1 -> 32S sum (+ 64F sum of squares): sumpixels.dispatch.cpp:192-263.  sum[y+1][x+1] = sum[y][x+1] + row prefix. */
PORT_API int port_integral(const void* src, size_t sstep, int w, int h, int* sum, size_t sumstep, double* sqsum, size_t sqstep)
{
    memset(sum, 0, sizeof(int) * (size_t)(w + 1));
    if (sqsum) memset(sqsum, 0, sizeof(double) * (size_t)(w + 1));
    for (int y = 0; y < h; y++) {
        const uchar* s = (const uchar*)src + (size_t)y * sstep;
        const int* prev = (const int*)((const char*)sum + (size_t)y * sumstep); int* cur = (int*)((char*)sum + (size_t)(y + 1) * sumstep);
        unsigned run = 0; double qrun = 0;
        cur[0] = 0;
        for (int x = 0; x < w; x++) { run += s[x]; cur[x + 1] = (int)((unsigned)prev[x + 1] + run); }
        if (sqsum) {
            const double* qp = (const double*)((const char*)sqsum + (size_t)y * sqstep); double* qc = (double*)((char*)sqsum + (size_t)(y + 1) * sqstep);
            qc[0] = 0;
            for (int x = 0; x < w; x++) { qrun += (double)s[x] * s[x]; qc[x + 1] = qp[x + 1] + qrun; }
        }
    }
    return 0;
}

/* ---- cv::boxFilter / cv::blur (box_filter.dispatch.cpp:440-498; box_filter.simd.hpp) ------------------------------------------------
 * Sum type as createBoxFilter picks it (simd.hpp:1250-1272): 8U->8U with kw*kh <= 256 -> 16-bit sums and the integer divide of
 * ColumnSum<ushort,uchar> (simd.hpp:430-600: d = cvRound(1/scale), (s + divDelta) * divScale >> 23); other 8U sources -> int sums,
 * scaled as float(s) * float(scale) in the SIMD body and as double s * scale in the scalar remainder (ColumnSum<int,uchar> :275-428,
 * <int,float> :1039-1158; remainder = last (w*cn) mod 8 resp. mod 4 elements); 32F -> double sums, RowSum direct for 3 / 5 taps and
 * sliding otherwise (:64-172), ColumnSum sliding down the whole image (:175-272), result (float)(s * scale).
 * Borders through cv::borderInterpolate; BORDER_CONSTANT pads with zeros (no borderValue on this path). */
PORT_API int port_box_filter(const void* src, size_t sstep, int w, int h, int stype, void* dst, size_t dstep, int ddepth,
                             int kw, int kh, int ax, int ay, int normalize, int border)
{
    int sdepth = P_DEPTH(stype), cn = P_CN(stype), b = border & ~16;
    if (ddepth < 0) ddepth = sdepth;
    if (ax < 0) ax = kw / 2;
    if (ay < 0) ay = kh / 2;
    if (kw <= 0 || kh <= 0 || ax >= kw || ay >= kh || b == PB_WRAP) return -1;
    if (!((sdepth == P_8U && (ddepth == P_8U || ddepth == P_32F)) || (sdepth == P_32F && ddepth == P_32F))) return -1;
    const int wn = w * cn;
    const double scale = normalize ? 1. / ((double)kw * kh) : 1.;
    const int have_scale = scale != 1;
    const int u16path = sdepth == P_8U && ddepth == P_8U && kw * kh <= 256;
    int div_scale = 1, div_delta = 0;
    if (u16path && have_scale) {
        int d = port_round(1. / scale);
        double sf = (double)(1 << 23) / d;
        div_scale = (int)floor(sf);
        sf -= div_scale;
        div_delta = d / 2;
        if (sf < 0.5) div_delta++; else div_scale++;
    }
    /* padded source row indices */
    int* xtab = (int*)malloc(sizeof(int) * (size_t)(w + kw));
    for (int j = 0; j < w + kw - 1; j++) xtab[j] = port_border(j - ax, w, b);
    const int isf = sdepth == P_32F;
    /* row sums of every source row (int or double), then the sliding column sum in image order */
    double* rsd = isf ? (double*)malloc(sizeof(double) * (size_t)wn * h) : NULL;
    int* rsi = isf ? NULL : (int*)malloc(sizeof(int) * (size_t)wn * h);
    double* padd = isf ? (double*)malloc(sizeof(double) * (size_t)(w + kw) * cn) : NULL;
    for (int y = 0; y < h; y++) {
        const uchar* su = (const uchar*)src + (size_t)y * sstep;
        const float* sf = (const float*)su;
        if (!isf) {
            for (int e = 0; e < wn; e++) {
                int x = e / cn, c = e - x * cn, s = 0;
                for (int k = 0; k < kw; k++) { int sx = xtab[x + k]; s += sx < 0 ? 0 : su[sx * cn + c]; }
                rsi[(size_t)y * wn + e] = s;
            }
        } else {
            for (int j = 0; j < w + kw - 1; j++)
                for (int c = 0; c < cn; c++) padd[j * cn + c] = xtab[j] < 0 ? 0. : (double)sf[xtab[j] * cn + c];
            double* D = rsd + (size_t)y * wn;
            if (kw == 3) for (int e = 0; e < wn; e++) D[e] = padd[e] + padd[e + cn] + padd[e + 2 * cn];
            else if (kw == 5) for (int e = 0; e < wn; e++) D[e] = padd[e] + padd[e + cn] + padd[e + 2 * cn] + padd[e + 3 * cn] + padd[e + 4 * cn];
            else
                for (int c = 0; c < cn; c++) {
                    double s = 0;
                    for (int k = 0; k < kw; k++) s += padd[k * cn + c];
                    D[c] = s;
                    for (int x = 0; x < w - 1; x++) { s += padd[(x + kw) * cn + c] - padd[x * cn + c]; D[(x + 1) * cn + c] = s; }
                }
        }
    }
    const float scale_f = (float)scale;
    const int tail8 = wn - wn % 8, tail4 = wn - wn % 4;
    if (!isf) {
        int* SUM = (int*)calloc((size_t)wn, sizeof(int));
        for (int r = 0; r < kh - 1; r++) {
            int sy = port_border(r - ay, h, b);
            if (sy >= 0) for (int e = 0; e < wn; e++) SUM[e] += rsi[(size_t)sy * wn + e];
        }
        for (int y = 0; y < h; y++) {
            int sp = port_border(y - ay + kh - 1, h, b), sm = port_border(y - ay, h, b);
            uchar* du = (uchar*)dst + (size_t)y * dstep;
            float* df = (float*)du;
            for (int e = 0; e < wn; e++) {
                int s0 = SUM[e] + (sp < 0 ? 0 : rsi[(size_t)sp * wn + e]);
                if (ddepth == P_8U) {
                    if (u16path) du[e] = have_scale ? (uchar)(((unsigned)(s0 + div_delta) * (unsigned)div_scale) >> 23) : port_sat_u8i(s0);
                    else if (!have_scale) du[e] = port_sat_u8i(s0);
                    else if (e < tail8) du[e] = port_sat_u8i((int)lrintf(mul_rn((float)s0, scale_f)));
                    else du[e] = port_sat_u8i(port_round(s0 * scale));
                } else {
                    if (!have_scale) df[e] = (float)s0;
                    else if (e < tail4) df[e] = mul_rn((float)s0, scale_f);
                    else df[e] = (float)(s0 * scale);
                }
                SUM[e] = s0 - (sm < 0 ? 0 : rsi[(size_t)sm * wn + e]);
            }
        }
        free(SUM);
    } else {
        double* SUM = (double*)calloc((size_t)wn, sizeof(double));
        for (int r = 0; r < kh - 1; r++) {
            int sy = port_border(r - ay, h, b);
            for (int e = 0; e < wn; e++) SUM[e] += sy < 0 ? 0. : rsd[(size_t)sy * wn + e];
        }
        for (int y = 0; y < h; y++) {
            int sp = port_border(y - ay + kh - 1, h, b), sm = port_border(y - ay, h, b);
            float* df = (float*)((uchar*)dst + (size_t)y * dstep);
            for (int e = 0; e < wn; e++) {
                double s0 = SUM[e] + (sp < 0 ? 0. : rsd[(size_t)sp * wn + e]);
                df[e] = have_scale ? (float)(s0 * scale) : (float)s0;
                SUM[e] = s0 - (sm < 0 ? 0. : rsd[(size_t)sm * wn + e]);
            }
        }
        free(SUM);
    }
    free(xtab); free(rsd); free(rsi); free(padd);
    return 0;
}
