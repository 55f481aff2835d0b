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
