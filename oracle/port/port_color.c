/* oracle/port/port_color.c -- cvtColor (8-bit BGR/RGB(A) <-> GRAY / YUV / YCrCb / HSV / BGR(A)) restated in scalar C.
 * TEST INFRASTRUCTURE ONLY (see port_common.h).
 *   dispatch            modules/imgproc/src/color.cpp:208-390
 *   RGB2Gray<uchar>     modules/imgproc/src/color_rgb.simd.hpp:660-750
 *   RGB2YCrCb_i<uchar>  modules/imgproc/src/color_yuv.simd.hpp:397-572 ; YCrCb2RGB_i<uchar> :738-888
 *   RGB2HSV_b           modules/imgproc/src/color_hsv.simd.hpp:47-268 ; HSV2RGB_b :518-672 (vector body truncates, scalar tail rounds)
 */
#include "port_common.h"

#define DESCALE(x, n) (((x) + (1 << ((n)-1))) >> (n))

static void hsv2bgr_px(const uchar* s, uchar* d, int bidx, float hscale, int truncate)
{
    float h = (float)s[0], sv = (float)s[1] * (1.0f / 255.0f), v = (float)s[2] * (1.0f / 255.0f);
    h = h * hscale;
    float pre = truncf(h);
    h = h - pre;
    /* the reference's AVX2 unit is built with -mfma and GCC contracts 1 - s*h into one fnmadd (observed: 0 mismatches
       against the built reference with the fused form, 6/67065 pixels off by one without it) */
    float t0 = v, t1 = v * (1.0f - sv), t2 = v * fmaf(-sv, h, 1.0f), t3 = v * fmaf(-sv, 1.0f - h, 1.0f);
    float sec = truncf(pre * (1.0f / 6.0f));
    int sector = (int)(pre - sec * 6.0f);
    float b, g, r;
    switch (sector) {
    case 0: b = t1; g = t3; r = t0; break;
    case 1: b = t1; g = t0; r = t2; break;
    case 2: b = t3; g = t0; r = t1; break;
    case 3: b = t0; g = t2; r = t1; break;
    case 4: b = t0; g = t1; r = t3; break;
    default: b = t2; g = t1; r = t0; break;
    }
    b *= 255.0f; g *= 255.0f; r *= 255.0f;
    if (truncate) { d[bidx] = port_sat_u8i((int)b); d[1] = port_sat_u8i((int)g); d[bidx ^ 2] = port_sat_u8i((int)r); }
    else { d[bidx] = port_sat_u8f(b); d[1] = port_sat_u8f(g); d[bidx ^ 2] = port_sat_u8f(r); }
}

PORT_API int port_cvt_color(const void* src_, size_t sstep, void* dst_, size_t dstep, int w, int h, int stype, int dtype, int code)
{
    if (P_DEPTH(stype) != P_8U || P_DEPTH(dtype) != P_8U) return -1;
    int scn = P_CN(stype), dcn = P_CN(dtype);
    static int sdiv[256], hdiv180[256], hdiv256[256], tables = 0;
    if (!tables) {
        for (int i = 1; i < 256; i++) {
            sdiv[i] = port_round((255 << 12) / (1. * i));
            hdiv180[i] = port_round((180 << 12) / (6. * i));
            hdiv256[i] = port_round((256 << 12) / (6. * i));
        }
        tables = 1;
    }
    for (int y = 0; y < h; y++) {
        const uchar* s = (const uchar*)src_ + (size_t)y * sstep;
        uchar* d = (uchar*)dst_ + (size_t)y * dstep;
        for (int x = 0; x < w; x++, s += scn, d += dcn) {
            switch (code) {
            case 0: case 1: case 2: case 3: case 4: case 5: {
                int swap = code >= 2;
                uchar b = s[swap ? 2 : 0], g = s[1], r = s[swap ? 0 : 2];
                d[0] = b; d[1] = g; d[2] = r;
                if (dcn == 4) d[3] = scn == 4 ? s[3] : 255;
                break;
            }
            case 6: case 10: d[0] = (uchar)DESCALE(s[0] * 3735 + s[1] * 19235 + s[2] * 9798, 15); break;
            case 7: case 11: d[0] = (uchar)DESCALE(s[0] * 9798 + s[1] * 19235 + s[2] * 3735, 15); break;
            case 8: case 9: d[0] = d[1] = d[2] = s[0]; if (dcn == 4) d[3] = 255; break;
            case 36: case 37: case 82: case 83: {
                int crcb = code < 40, bidx = (code == 36 || code == 82) ? 0 : 2;
                int c[5] = {4899, 9617, 1868, crcb ? 11682 : 14369, crcb ? 9241 : 8061};
                if (bidx == 0) { int t = c[0]; c[0] = c[2]; c[2] = t; }
                int yuv = !crcb, delta = 128 << 14;
                int Y = DESCALE(s[0] * c[0] + s[1] * c[1] + s[2] * c[2], 14);
                int Cr = DESCALE((s[bidx ^ 2] - Y) * c[3] + delta, 14);
                int Cb = DESCALE((s[bidx] - Y) * c[4] + delta, 14);
                d[0] = port_sat_u8i(Y); d[1 + yuv] = port_sat_u8i(Cr); d[2 - yuv] = port_sat_u8i(Cb);
                break;
            }
            case 38: case 39: case 84: case 85: {
                int crcb = code < 40, bidx = (code == 38 || code == 84) ? 0 : 2, yuv = !crcb;
                int c0 = crcb ? 22987 : 18678, c1 = crcb ? -11698 : -9519, c2 = crcb ? -5636 : -6472, c3 = crcb ? 29049 : 33292;
                int Y = s[0], Cr = s[1 + yuv], Cb = s[2 - yuv];
                d[bidx] = port_sat_u8i(Y + DESCALE((Cb - 128) * c3, 14));
                d[1] = port_sat_u8i(Y + DESCALE((Cb - 128) * c2 + (Cr - 128) * c1, 14));
                d[bidx ^ 2] = port_sat_u8i(Y + DESCALE((Cr - 128) * c0, 14));
                if (dcn == 4) d[3] = 255;
                break;
            }
            case 40: case 41: case 66: case 67: {
                int bidx = (code == 40 || code == 66) ? 0 : 2, hr = code < 60 ? 180 : 256;
                const int* hdiv = hr == 180 ? hdiv180 : hdiv256;
                int b = s[bidx], g = s[1], r = s[bidx ^ 2];
                int v = b > g ? b : g; if (r > v) v = r;
                int mn = b < g ? b : g; if (r < mn) mn = r;
                int diff = v - mn, vr = v == r ? -1 : 0, vg = v == g ? -1 : 0;
                int sat = (diff * sdiv[v] + (1 << 11)) >> 12;
                int hh = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
                hh = (hh * hdiv[diff] + (1 << 11)) >> 12;
                hh += hh < 0 ? hr : 0;
                d[0] = port_sat_u8i(hh); d[1] = (uchar)sat; d[2] = (uchar)v;
                break;
            }
            case 54: case 55: case 70: case 71: {
                int bidx = (code == 54 || code == 70) ? 0 : 2;
                float hscale = 6.0f / (code < 60 ? 180 : 255);   /* inverse _FULL uses 255: color_hsv.simd.hpp:1302 */
                int trunc_cols = w >= 32 ? (w / 32) * 32 : 0;   /* AVX2 vector body covers 32 pixels per iteration */
                hsv2bgr_px(s, d, bidx, hscale, x < trunc_cols);
                if (dcn == 4) d[3] = 255;
                break;
            }
            default: return 1;
            }
        }
    }
    return 0;
}

/* ---- subsampled YUV wire formats (color_yuv.simd.hpp:1018-2202; BT.601 limited range, 20-bit fixed point) -----------------------------
 * 4:2:0 two-plane NV12 / NV21 (codes 90-97), three-plane YV12 / IYUV (98-105), Y extraction (106), 4:2:2 UYVY / YUY2 / YVYU (107-124),
 * BGR(A) / RGB(A) -> IYUV / YV12 (127-134).  A 4:2:0 image of W x H pixels is one 8-bit plane of H*3/2 rows: Y rows, then the chroma:
 * interleaved (NV) rows of W bytes, or planar half rows of W/2 bytes, two to a row, all U (V for YV12) rows before the others.
 *   ruv = 2^19 + 1673527 (v-128);  guv = 2^19 - 852492 (v-128) - 409993 (u-128);  buv = 2^19 + 2116026 (u-128)      (:1043-1052)
 *   y' = max(0, y-16) * 1220542;   c = saturate((y' + cuv) >> 20)                                                     (:1090-1099)
 *   Y = (269484 r + 528482 g + 102760 b + 2^19 + (16 << 20)) >> 20;  U, V likewise from the EVEN row, EVEN column pixel (:1473-1523) */
static void yuv_px(int y, int u, int v, int bidx, int dcn, uchar* d)
{
    int uu = u - 128, vv = v - 128;
    int ruv = (1 << 19) + 1673527 * vv, guv = (1 << 19) - 852492 * vv - 409993 * uu, buv = (1 << 19) + 2116026 * uu;
    int yy = (y - 16 > 0 ? y - 16 : 0) * 1220542;
    d[2 - bidx] = port_sat_u8i((yy + ruv) >> 20);
    d[1] = port_sat_u8i((yy + guv) >> 20);
    d[bidx] = port_sat_u8i((yy + buv) >> 20);
    if (dcn == 4) d[3] = 255;
}

/* cv::cvtColorTwoPlane (color.cpp:171-185): NV12 / NV21 with separate luma / chroma buffers */
PORT_API int port_cvt_color_two_plane(const void* y_, size_t ystep, const void* uv_, size_t uvstep, int w, int h, void* dst_, size_t dstep, int dcn, int code)
{
    if (code < 90 || code > 97 || (w & 1) || (h & 1) || (dcn != 3 && dcn != 4)) return -1;
    const int c = code - 90, rgb = !(c & 1), uidx = (c >> 1) & 1, bidx = rgb ? 2 : 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uchar* uv = (const uchar*)uv_ + (size_t)(y / 2) * uvstep + (x & ~1);
            yuv_px(((const uchar*)y_ + (size_t)y * ystep)[x], uv[uidx], uv[1 - uidx], bidx, dcn, (uchar*)dst_ + (size_t)y * dstep + x * dcn);
        }
    return 0;
}

PORT_API int port_cvt_color_yuv(const void* src_, size_t sstep, int sw, int sh, int scn, void* dst_, size_t dstep, int dw, int dh, int dcn, int code)
{
    const uchar* src = (const uchar*)src_;
    uchar* dst = (uchar*)dst_;
    if (code >= 90 && code <= 105) {                         /* 4:2:0 -> BGR family */
        if (scn != 1 || sw != dw || sh != dh * 3 / 2 || (dw & 1) || (dh & 1) || (dcn != 3 && dcn != 4)) return -1;
        const int planar = code >= 98;
        int rgb, uidx;                                       /* uidx 1: V comes first */
        if (!planar) { int c = code - 90; rgb = !(c & 1); uidx = (c >> 1) & 1; }
        else { int c = (code - 98) & 3; rgb = !(c & 1); uidx = c < 2; }
        const int bidx = rgb ? 2 : 0;
        for (int y = 0; y < dh; y++)
            for (int x = 0; x < dw; x++) {
                int u, v;
                if (!planar) {
                    const uchar* uv = src + (size_t)(dh + y / 2) * sstep + (x & ~1);
                    u = uv[uidx]; v = uv[1 - uidx];
                } else {
                    int k0 = y / 2, k1 = dh / 2 + y / 2;      /* half-row index of the first / second chroma plane */
                    int a = src[(size_t)(dh + k0 / 2) * sstep + (k0 & 1) * (dw / 2) + x / 2];
                    int b = src[(size_t)(dh + k1 / 2) * sstep + (k1 & 1) * (dw / 2) + x / 2];
                    u = uidx ? b : a; v = uidx ? a : b;
                }
                yuv_px(src[(size_t)y * sstep + x], u, v, bidx, dcn, dst + (size_t)y * dstep + x * dcn);
            }
        return 0;
    }
    if (code == 106) {                                       /* Y plane of a 4:2:0 image */
        if (scn != 1 || dcn != 1 || sw != dw || sh != dh * 3 / 2) return -1;
        for (int y = 0; y < dh; y++) memcpy(dst + (size_t)y * dstep, src + (size_t)y * sstep, (size_t)dw);
        return 0;
    }
    if ((code >= 107 && code <= 124) && code != 109 && code != 110 && code != 113 && code != 114) {   /* 4:2:2 interleaved, 2 bytes / pixel */
        if (scn != 2 || sw != dw || sh != dh || (dw & 1)) return -1;
        if (code >= 123) {
            if (dcn != 1) return -1;
            for (int y = 0; y < dh; y++)
                for (int x = 0; x < dw; x++) dst[(size_t)y * dstep + x] = src[(size_t)y * sstep + 2 * x + (code == 123 ? 1 : 0)];
            return 0;
        }
        if (dcn != 3 && dcn != 4) return -1;
        const int ycn = (code == 107 || code == 108 || code == 111 || code == 112) ? 1 : 0;      /* UYVY: luma in the odd bytes */
        const int yvyu = code == 117 || code == 118 || code == 121 || code == 122;
        const int rgb = code == 107 || code == 111 || code == 115 || code == 117 || code == 119 || code == 121;
        const int uoff = 1 - ycn + yvyu * 2, voff = (2 + uoff) % 4, bidx = rgb ? 2 : 0;
        for (int y = 0; y < dh; y++)
            for (int x = 0; x < dw; x++) {
                const uchar* q = src + (size_t)y * sstep + (x >> 1) * 4;
                yuv_px(q[ycn + (x & 1) * 2], q[uoff], q[voff], bidx, dcn, dst + (size_t)y * dstep + x * dcn);
            }
        return 0;
    }
    if ((code >= 46 && code <= 49) || (code >= 139 && code <= 142)) {
        /* Bayer mosaic -> BGR / BGRA, bilinear (demosaicing.cpp:806-1056): interior rows / columns as the reference walks them (the
         * non-green sites of a row carry `blue`-side colour, rows alternate), then first / last columns and rows copied from inside */
        const int four = code >= 139, c = four ? code - 139 : code - 46;
        if (scn != 1 || dcn != (four ? 4 : 3) || sw != dw || sh != dh || sw < 3 || sh < 3) return -1;
        int blue = c < 2 ? -1 : 1, swg = c & 1;
        for (int i = 0; i < sh - 2; i++, blue = -blue, swg = !swg) {
            const uchar* b0 = src + (size_t)i * sstep; const uchar* b1 = b0 + sstep; const uchar* b2 = b1 + sstep;
            uchar* drow = dst + (size_t)(i + 1) * dstep;
            for (int k = 0; k < sw - 2; k++) {
                uchar* d = drow + (k + 1) * dcn + 1;                    /* the green channel of interior pixel k */
                int green = ((k & 1) == 0) == (swg != 0);
                if (green) {
                    d[-blue] = (uchar)((b0[k + 1] + b2[k + 1] + 1) >> 1);
                    d[0] = b1[k + 1];
                    d[blue] = (uchar)((b1[k] + b1[k + 2] + 1) >> 1);
                } else {
                    d[-blue] = (uchar)((b0[k] + b0[k + 2] + b2[k] + b2[k + 2] + 2) >> 2);
                    d[0] = (uchar)((b0[k + 1] + b1[k] + b1[k + 2] + b2[k + 1] + 2) >> 2);
                    d[blue] = b1[k + 1];
                }
                if (dcn == 4) d[2] = 255;
            }
            memcpy(drow, drow + dcn, dcn);
            memcpy(drow + (sw - 1) * dcn, drow + (sw - 2) * dcn, dcn);
        }
        memcpy(dst, dst + dstep, (size_t)sw * dcn);
        memcpy(dst + (size_t)(sh - 1) * dstep, dst + (size_t)(sh - 2) * dstep, (size_t)sw * dcn);
        return 0;
    }
    if (code >= 143 && code <= 154) {                        /* BGR family -> 4:2:2 UYVY / YUY2 / YVYU (color_yuv.simd.hpp:1862-1958, 14-bit fixed point) */
        if (dcn != 2 || sw != dw || sh != dh || (sw & 1) || (scn != 3 && scn != 4)) return -1;
        const int uyvy = code <= 146, yvyu = code == 149 || code == 150 || code == 153 || code == 154;
        const int rgb = (code & 1), ycn = uyvy ? 1 : 0;      /* odd codes are the RGB(A) ones */
        const int uoff = 1 - ycn + yvyu * 2, voff = (2 + uoff) % 4, bidx = rgb ? 2 : 0;
        for (int y = 0; y < sh; y++)
            for (int x = 0; x < sw; x += 2) {
                const uchar* p = src + (size_t)y * sstep + x * scn; uchar* q = dst + (size_t)y * dstep + x * 2;
                int r1 = p[2 - bidx], g1 = p[1], b1 = p[bidx], r2 = p[scn + 2 - bidx], g2 = p[scn + 1], b2 = p[scn + bidx];
                q[ycn] = port_sat_u8i(((1 << 13) + r1 * 4211 + g1 * 8258 + b1 * 1606 + (1 << 14) * 16) >> 14);
                q[ycn + 2] = port_sat_u8i(((1 << 13) + r2 * 4211 + g2 * 8258 + b2 * 1606 + (1 << 14) * 16) >> 14);
                int sr = r1 + r2, sg = g1 + g2, sb = b1 + b2;
                q[uoff] = port_sat_u8i(((1 << 13) + sr * -1212 + sg * -2384 + sb * 3596 + (1 << 13) * 256) >> 14);
                q[voff] = port_sat_u8i(((1 << 13) + sr * 3596 + sg * -3015 + sb * -582 + (1 << 13) * 256) >> 14);
            }
        return 0;
    }
    if (code >= 127 && code <= 134) {                        /* BGR family -> IYUV / YV12 */
        if (dcn != 1 || sw != dw || dh != sh * 3 / 2 || (sw & 1) || (sh & 1) || (scn != 3 && scn != 4)) return -1;
        const int c = (code - 127) & 3, rgb = !(c & 1), yv12 = code >= 131;   /* the channel count is the source's, whatever the code says (the reference's own KAT feeds 3 channels to the RGBA codes) */
        const int bidx = rgb ? 2 : 0, w = sw, h = sh;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                const uchar* p = src + (size_t)y * sstep + x * scn;
                int b = p[bidx], g = p[1], r = p[2 - bidx];
                dst[(size_t)y * dstep + x] = port_sat_u8i((269484 * r + 528482 * g + 102760 * b + (1 << 19) + (16 << 20)) >> 20);
                if (!(y & 1) && !(x & 1)) {
                    int uu = (-155188 * r - 305135 * g + 460324 * b + (1 << 19) + (128 << 20)) >> 20;
                    int vv = (460324 * r - 385875 * g - 74448 * b + (1 << 19) + (128 << 20)) >> 20;
                    int ku = y / 2 + (yv12 ? h / 2 : 0), kv = y / 2 + (yv12 ? 0 : h / 2);
                    dst[(size_t)(h + ku / 2) * dstep + (ku & 1) * (w / 2) + x / 2] = port_sat_u8i(uu);
                    dst[(size_t)(h + kv / 2) * dstep + (kv & 1) * (w / 2) + x / 2] = port_sat_u8i(vv);
                }
            }
        return 0;
    }
    return -1;
}

/* ---- 8-bit BGR / RGB -> Lab (color_lab.cpp:1573-1890, RGB2Lab_b; tables :1225-1280): all integer after the two tables ------------------
 *   g = gamma table (sRGB curve at 255 * 8 resolution, or linear), X/Y/Z = (R*C0 + G*C1 + B*C2 + 2^11) >> 12 with the XYZ matrix divided by
 *   the white point, f = cube-root table at 2^15 scale, L = (296 f(Y) - Lshift' ) >> 15, a = (500 (fX - fY) + ...) >> 15, b likewise.
 * The tables come from softfloat pow / cbrt in the reference; here from libm (double pow -> float, cbrtf): pinned by running the whole 2^24
 * colour cube against the reference (tests/test_oracle.py). */
/* cv::cbrt(softfloat) (core/src/softfloat.cpp:3897-3930): exponent split by three, a quartic rational polynomial of the mantissa in double
 * (IEEE operations, no fusing), the result's mantissa TRUNCATED to 23 bits -- not the correctly rounded cbrtf, and the table needs its bits */
static float soft_cbrtf(float x)
{
    uint32_t v; memcpy(&v, &x, 4);
    if ((v & 0x7fffffffu) == 0) return 0.f;
    const uint32_t s = v >> 31;
    int ex = (int)((v >> 23) & 255) - 127, shx = ex % 3;
    shx -= shx >= 0 ? 3 : 0;
    ex = (ex - shx) / 3 - 1;
    uint64_t fv = ((uint64_t)(shx + 1023) << 52) | ((uint64_t)(v & 0x7fffffu) << 29);
    double fr; memcpy(&fr, &fv, 8);
    static const uint64_t K[9] = {0x4046a09e6653ba70ull, 0x406808f46c6116e0ull, 0x405dca97439cae14ull, 0x402add70d2827500ull, 0x3fc4f15f83f55d2dull,
                                  0x402d9e20660edb21ull, 0x4062ff15c0285815ull, 0x406510d06a8112ceull, 0x4040fecbc9e2c375ull};
    double A[9]; memcpy(A, K, sizeof(A));
    volatile double num = A[0] * fr; num = num + A[1]; num = num * fr; num = num + A[2]; num = num * fr; num = num + A[3]; num = num * fr; num = num + A[4];
    volatile double den = A[5] * fr; den = den + A[6]; den = den * fr; den = den + A[7]; den = den * fr; den = den + A[8]; den = den * fr; den = den + 1.0;
    double q = num / den;
    uint64_t r; memcpy(&r, &q, 8);
    uint32_t y = (s << 31) | ((uint32_t)(ex + 127) << 23) | (uint32_t)((r & 0xFFFFFFFFFFFFFull) >> 29);
    float out; memcpy(&out, &y, 4);
    return out;
}

static unsigned short g_lab_gamma[256], g_lab_lin[256], g_lab_cbrt[256 * 3 / 2 * 8];
static int g_lab_ready = 0;
static void lab_tabs(void)
{
    if (g_lab_ready) return;
    const float intScale = 255 * 8;
    for (int i = 0; i < 256; i++) {
        float x = (float)i / 255.f;
        double xd = x;
        float g = (float)(xd <= 809. / 20000. ? xd / (323. / 25.) : pow((xd + 11. / 200.) / (1. + 11. / 200.), 12. / 5.));
        g_lab_gamma[i] = (unsigned short)lrintf(intScale * g);
        g_lab_lin[i] = (unsigned short)(i * 8);
    }
    const float cbScale = 1.f / (255.f * 8), lthresh = 216.f / 24389.f, lscale = 841.f / 108.f, lbias = 16.f / 116.f, lshift2 = 32768.f;
    for (int i = 0; i < 256 * 3 / 2 * 8; i++) {
        float x = cbScale * (float)i;
        float f = x < lthresh ? fmaf(x, lscale, lbias) : soft_cbrtf(x);
        g_lab_cbrt[i] = (unsigned short)lrintf(lshift2 * f);
    }
    g_lab_ready = 1;
}

PORT_API int port_cvt_color_lab(const void* src_, size_t sstep, void* dst_, size_t dstep, int w, int h, int scn, int code)
{
    /* 44 BGR2Lab, 45 RGB2Lab (sRGB gamma); 74 LBGR2Lab, 75 LRGB2Lab (linear) */
    if ((code != 44 && code != 45 && code != 74 && code != 75) || (scn != 3 && scn != 4)) return -1;
    lab_tabs();
    const int bidx = (code == 44 || code == 74) ? 0 : 2, srgb = code < 70;
    const unsigned short* tab = srgb ? g_lab_gamma : g_lab_lin;
    static const double M[9] = {0.412453, 0.357580, 0.180423, 0.212671, 0.715160, 0.072169, 0.019334, 0.119193, 0.950227};
    static const double wp[3] = {0.950456, 1., 1.088754};
    int C[9];
    for (int i = 0; i < 3; i++) {
        C[i * 3 + (bidx ^ 2)] = port_round(4096. * M[i * 3] / wp[i]);
        C[i * 3 + 1] = port_round(4096. * M[i * 3 + 1] / wp[i]);
        C[i * 3 + bidx] = port_round(4096. * M[i * 3 + 2] / wp[i]);
    }
    const int Lscale = (116 * 255 + 50) / 100, Lshift = -((16 * 255 * (1 << 15) + 50) / 100);
    for (int y = 0; y < h; y++) {
        const uchar* s = (const uchar*)src_ + (size_t)y * sstep; uchar* d = (uchar*)dst_ + (size_t)y * dstep;
        for (int x = 0; x < w; x++, s += scn, d += 3) {
            int R = tab[s[0]], G = tab[s[1]], B = tab[s[2]];
            int fX = g_lab_cbrt[(R * C[0] + G * C[1] + B * C[2] + (1 << 11)) >> 12];
            int fY = g_lab_cbrt[(R * C[3] + G * C[4] + B * C[5] + (1 << 11)) >> 12];
            int fZ = g_lab_cbrt[(R * C[6] + G * C[7] + B * C[8] + (1 << 11)) >> 12];
            int L = (Lscale * fY + Lshift + (1 << 14)) >> 15;
            int a = (500 * (fX - fY) + 128 * (1 << 15) + (1 << 14)) >> 15;
            int b = (200 * (fY - fZ) + 128 * (1 << 15) + (1 << 14)) >> 15;
            d[0] = port_sat_u8i(L); d[1] = port_sat_u8i(a); d[2] = port_sat_u8i(b);
        }
    }
    return 0;
}

/* ---- 8-bit BGR / RGB <-> CIE XYZ (RGB2XYZ_i<uchar> color_lab.cpp:250-330, XYZ2RGB_i<uchar> :650-730; integer matrices :132-144, 12-bit) ------
 * codes 32 BGR2XYZ, 33 RGB2XYZ, 34 XYZ2BGR, 35 XYZ2RGB: out = saturate((M * in + 2^11) >> 12); BGR order swaps matrix columns (to XYZ) or rows. */
static void xyz_matrix(int code, int* k)
{
    static const int fwd[9] = {1689, 1465, 739, 871, 2929, 296, 79, 488, 3892}, inv[9] = {13273, -6296, -2042, -3970, 7684, 170, 228, -836, 4331};
    const int to_xyz = code == 32 || code == 33, bgr = code == 32 || code == 34;
    for (int i = 0; i < 9; i++) k[i] = to_xyz ? fwd[i] : inv[i];
    if (bgr) {
        if (to_xyz) for (int r = 0; r < 3; r++) { int t = k[3 * r]; k[3 * r] = k[3 * r + 2]; k[3 * r + 2] = t; }
        else for (int c = 0; c < 3; c++) { int t = k[c]; k[c] = k[6 + c]; k[6 + c] = t; }
    }
}

PORT_API int port_cvt_color_xyz(const void* src_, size_t sstep, void* dst_, size_t dstep, int w, int h, int scn, int dcn, int code)
{
    if (code < 32 || code > 35 || (scn != 3 && scn != 4) || (dcn != 3 && dcn != 4)) return -1;
    if ((code <= 33 && dcn != 3) || (code >= 34 && scn != 3)) return -1;
    int k[9]; xyz_matrix(code, k);
    for (int y = 0; y < h; y++) {
        const uchar* s = (const uchar*)src_ + (size_t)y * sstep; uchar* d = (uchar*)dst_ + (size_t)y * dstep;
        for (int x = 0; x < w; x++, s += scn, d += dcn) {
            for (int r = 0; r < 3; r++) d[r] = port_sat_u8i((s[0] * k[3 * r] + s[1] * k[3 * r + 1] + s[2] * k[3 * r + 2] + (1 << 11)) >> 12);
            if (dcn == 4) d[3] = 255;
        }
    }
    return 0;
}

/* ---- 8-bit Lab -> BGR / RGB (Lab2RGBinteger, color_lab.cpp:2399-2700; tables :1263-1308, :1086-1107): integer after the tables ---------
 *   (y, ify) = LabToYF_b[L];  x = abToXZ_b[ify + adiv(a)], z = abToXZ_b[ify - bdiv(b)]  (piecewise linear / cubic in 14-bit fixed point);
 *   rgb = (C * xyz + 2^13) >> 14 clipped to [0, 4095], then the inverse-gamma table (sRGB) or (v * 255) >> 12 (linear). */
static unsigned short g_lab_yf[512], g_lab_invgamma[4096];
static int* g_lab_abxz = 0;
static void lab_inv_tabs(void)
{
    if (g_lab_abxz) return;
    const int BASE = 1 << 14;
    for (int i = 0; i < 256; i++) {
        int yv, ify;
        if (i <= 20) {
            yv = (int)lrintf((float)(i * BASE * 20 * 9) / (float)(17 * 29 * 29 * 29));
            ify = (int)lrintf((float)BASE * ((float)16 / (float)116 + (float)(i * 5) / (float)(3 * 17 * 29)));
        } else {
            float fy = (float)(i * 100 * BASE) / (float)(255 * 116) + (float)(16 * BASE) / (float)116;
            ify = (int)lrintf(fy);
            volatile float f2 = fy * fy; volatile float f3 = f2 * fy;
            yv = (int)lrintf(f3 / (float)(BASE * BASE));
        }
        g_lab_yf[2 * i] = (unsigned short)yv; g_lab_yf[2 * i + 1] = (unsigned short)ify;
    }
    for (int i = 0; i < 4096; i++) {
        float x = (1.f / 4096.f) * (float)i;
        double xd = x;
        float ig = (float)(xd <= 7827. / 2500000. ? xd * (323. / 25.) : pow(xd, 1. / (12. / 5.)) * (1. + 11. / 200.) - 11. / 200.);
        g_lab_invgamma[i] = (unsigned short)lrintf(255.f * ig);
    }
    const int minAB = -8145, n = BASE * 9 / 4;
    int* t = (int*)malloc(sizeof(int) * (size_t)n);
    for (int i = minAB; i < n + minAB; i++)
        t[i - minAB] = i <= 3390 ? i * 108 / 841 - BASE * 16 / 116 * 108 / 841 : i * i / BASE * i / BASE;
    g_lab_abxz = t;
}

PORT_API int port_cvt_color_lab_inv(const void* src_, size_t sstep, void* dst_, size_t dstep, int w, int h, int dcn, int code)
{
    /* 56 Lab2BGR, 57 Lab2RGB (sRGB); 78 Lab2LBGR, 79 Lab2LRGB (linear) */
    if ((code != 56 && code != 57 && code != 78 && code != 79) || (dcn != 3 && dcn != 4)) return -1;
    lab_inv_tabs();
    const int bidx = (code == 56 || code == 78) ? 0 : 2, srgb = code < 70, BASE = 1 << 14, minAB = -8145;
    static const double M[9] = {3.240479, -1.53715, -0.498535, -0.969256, 1.875991, 0.041556, 0.055648, -0.204043, 1.057311};
    static const double wp[3] = {0.950456, 1., 1.088754};
    int C[9];
    for (int i = 0; i < 3; i++) {
        C[i + bidx * 3] = port_round(4096. * M[i] * wp[i]);
        C[i + 3] = port_round(4096. * M[i + 3] * wp[i]);
        C[i + (bidx ^ 2) * 3] = port_round(4096. * M[i + 6] * wp[i]);
    }
    for (int y = 0; y < h; y++) {
        const uchar* s = (const uchar*)src_ + (size_t)y * sstep; uchar* d = (uchar*)dst_ + (size_t)y * dstep;
        for (int x = 0; x < w; x++, s += 3, d += dcn) {
            const int LL = s[0], aa = s[1], bb = s[2];
            const int yv = g_lab_yf[LL * 2], ify = g_lab_yf[LL * 2 + 1];
            const int adiv = ((5 * aa * 53687 + (1 << 7)) >> 13) - 128 * BASE / 500, bdiv = ((bb * 41943 + (1 << 4)) >> 9) - 128 * BASE / 200 + 1;
            const int xv = g_lab_abxz[ify + adiv - minAB], zv = g_lab_abxz[ify - bdiv - minAB];
            int ro = (C[0] * xv + C[1] * yv + C[2] * zv + (1 << 13)) >> 14, go = (C[3] * xv + C[4] * yv + C[5] * zv + (1 << 13)) >> 14,
                bo = (C[6] * xv + C[7] * yv + C[8] * zv + (1 << 13)) >> 14;
            ro = ro < 0 ? 0 : ro > 4095 ? 4095 : ro; go = go < 0 ? 0 : go > 4095 ? 4095 : go; bo = bo < 0 ? 0 : bo > 4095 ? 4095 : bo;
            if (srgb) { ro = g_lab_invgamma[ro]; go = g_lab_invgamma[go]; bo = g_lab_invgamma[bo]; }
            else { ro = ((ro << 8) - ro) >> 12; go = ((go << 8) - go) >> 12; bo = ((bo << 8) - bo) >> 12; }
            d[0] = (uchar)(bo > 255 ? 255 : bo); d[1] = (uchar)(go > 255 ? 255 : go); d[2] = (uchar)(ro > 255 ? 255 : ro);   /* rows are placed by blueIdx */
            if (dcn == 4) d[3] = 255;
        }
    }
    return 0;
}
