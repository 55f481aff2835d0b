/* oracle/port/port_features.c -- matchTemplate, cornerHarris / cornerMinEigenVal, goodFeaturesToTrack, SIFT pyramid in scalar C.
 * TEST INFRASTRUCTURE ONLY (see port_common.h).
 *   cv::matchTemplate / common_matchTemplate  modules/imgproc/src/templmatch.cpp:1158-1194, :906-1029 (numerator: direct sum in
 *                                             place of the reference's block DFT :566-760 -- same quantity, exact instead of ~1e-7)
 *   cornerEigenValsVecs / calcHarris / calcMinEigenVal  modules/imgproc/src/corner.cpp:237-322, :104-155, :55-101
 *   cv::goodFeaturesToTrack                   modules/imgproc/src/featureselect.cpp:382-548
 *   SIFT createInitialImage / buildGaussianPyramid / buildDoGPyramid  modules/features2d/src/sift.dispatch.cpp:176-310
 *     (parity of the pyramid is NOT pinned by any reference test; this is the same composition of public calls)
 */
#include "port_common.h"

int port_resize(const void*, size_t, int, int, void*, size_t, int, int, int, int);
int port_warp_affine(const void*, size_t, int, int, void*, size_t, int, int, int, const double*, int, int, const double*);
int port_gaussian_blur(const void*, size_t, void*, size_t, int, int, int, int, int, double, double, int);
int port_sobel(const void*, size_t, void*, size_t, int, int, int, int, int, int, int, double, double, int);

PORT_API int port_match_template(const void* img, size_t istep, int iw, int ih, const void* tpl, size_t tstep, int tw, int th, int type,
                                 float* result, size_t rstep, int method)
{
    int depth = P_DEPTH(type);
    if (P_CN(type) != 1 || (depth != P_8U && depth != P_32F)) return 1;
    int ow = iw - tw + 1, oh = ih - th + 1;
#define IM(y, x) (depth == P_8U ? (double)((const uchar*)img + (size_t)(y) * istep)[x] : (double)((const float*)((const char*)img + (size_t)(y) * istep))[x])
#define TP(y, x) (depth == P_8U ? (double)((const uchar*)tpl + (size_t)(y) * tstep)[x] : (double)((const float*)((const char*)tpl + (size_t)(y) * tstep))[x])
    /* integral images (f64) */
    size_t isz = (size_t)(iw + 1) * (ih + 1);
    double* sum = (double*)calloc(isz * 2, sizeof(double)); double* sq = sum + isz;
    for (int y = 0; y < ih; y++) {
        double rs = 0, rq = 0;
        for (int x = 0; x < iw; x++) {
            double v = IM(y, x); rs += v; rq += v * v;
            sum[(size_t)(y + 1) * (iw + 1) + x + 1] = sum[(size_t)y * (iw + 1) + x + 1] + rs;
            sq[(size_t)(y + 1) * (iw + 1) + x + 1] = sq[(size_t)y * (iw + 1) + x + 1] + rq;
        }
    }
    double ts = 0, tq = 0, n = (double)tw * th;
    for (int y = 0; y < th; y++) for (int x = 0; x < tw; x++) { double v = TP(y, x); ts += v; tq += v * v; }
    double scale = 1. / n, mean = ts * scale, var = tq * scale - mean * mean; if (var < 0) var = 0;
    double sdv = sqrt(var), templNorm = sdv * sdv, invArea = 1. / ((double)th * tw);
    int numType = (method == 2 || method == 3) ? 0 : (method == 4 || method == 5) ? 1 : 2;
    int normed = method == 1 || method == 3 || method == 5;
    int flat = templNorm < 2.220446049250313e-16 && method == 5;
    double templSum2 = templNorm + mean * mean;
    if (numType != 1) { mean = 0; templNorm = templSum2; }
    templSum2 /= invArea; templNorm = sqrt(templNorm); templNorm /= sqrt(invArea);
    for (int y = 0; y < oh; y++) {
        float* rr = (float*)((char*)result + (size_t)y * rstep);
        for (int x = 0; x < ow; x++) {
            double acc = 0;
            for (int v = 0; v < th; v++) for (int u = 0; u < tw; u++) acc += TP(v, u) * IM(y + v, x + u);
            double num = (double)(float)acc, t, wm2 = 0, ws2 = 0;
            if (method == 2) { rr[x] = (float)num; continue; }
            if (flat) { rr[x] = 1.f; continue; }
#define BOX(p) (p[(size_t)y * (iw + 1) + x] - p[(size_t)y * (iw + 1) + x + tw] - p[(size_t)(y + th) * (iw + 1) + x] + p[(size_t)(y + th) * (iw + 1) + x + tw])
            if (numType == 1) { t = BOX(sum); wm2 += t * t; num -= t * mean; wm2 *= invArea; }
            if (normed || numType == 2) { ws2 += BOX(sq); if (numType == 2) { num = ws2 - 2 * num + templSum2; if (num < 0) num = 0; } }
            if (normed) {
                double d2 = ws2 - wm2; if (d2 < 0) d2 = 0;
                double lim = 10 * 1.1920928955078125e-07 * ws2; if (lim > 0.5) lim = 0.5;
                t = d2 <= lim ? 0 : sqrt(d2) * templNorm;
                if (fabs(num) < t) num /= t; else if (fabs(num) < t * 1.125) num = num > 0 ? 1 : -1; else num = method != 1 ? 0 : 1;
            }
            rr[x] = (float)num;
        }
    }
    free(sum);
    return 0;
}

static int corner_impl(const void* src, size_t sstep, int w, int h, int type, float* dst, size_t dstep, int bs, int ks, double k, int border, int op)
{
    int depth = P_DEPTH(type);
    if (P_CN(type) != 1 || (depth != P_8U && depth != P_32F)) return 1;
    double scale = (double)(1 << ((ks > 0 ? ks : 3) - 1)) * bs;
    if (depth == P_8U) scale *= 255.0;
    scale = 1.0 / scale;
    size_t n = (size_t)w * h;
    float* dx = (float*)malloc(sizeof(float) * n * 5); float* dy = dx + n; float* a = dy + n; float* b = a + n; float* c = b + n;
    port_sobel(src, sstep, dx, (size_t)w * 4, w, h, type, P_32F, 1, 0, ks, scale, 0, border);
    port_sobel(src, sstep, dy, (size_t)w * 4, w, h, type, P_32F, 0, 1, ks, scale, 0, border);
    for (size_t i = 0; i < n; i++) { a[i] = dx[i] * dx[i]; b[i] = dx[i] * dy[i]; c[i] = dy[i] * dy[i]; }
    int an = bs / 2;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            double sa = 0, sb = 0, sc = 0;
            for (int j = 0; j < bs; j++) {
                int yy = port_border(y + j - an, h, border);
                double ra = 0, rb = 0, rc = 0;
                for (int i = 0; i < bs; i++) {
                    int xx = port_border(x + i - an, w, border);
                    if (yy >= 0 && xx >= 0) { size_t o = (size_t)yy * w + xx; ra += a[o]; rb += b[o]; rc += c[o]; }
                }
                sa += ra; sb += rb; sc += rc;
            }
            float fa = (float)sa, fb = (float)sb, fc = (float)sc, out;
            /* calcHarrisLine_AVX (corner.avx.cpp:145-160, plain AVX object: no FMA): (a*c - b*b) - k*((a+c)*(a+c)) */
            if (op == 0) { float acbb = fa * fc - fb * fb, ac = fa + fc; out = acbb - (float)k * (ac * ac); }
            else { float ha = fa * 0.5f, hc = fc * 0.5f, t = ha - hc; t = fb * fb + t * t; out = (ha + hc) - sqrtf(t); }
            ((float*)((char*)dst + (size_t)y * dstep))[x] = out;
        }
    free(dx);
    return 0;
}

PORT_API int port_corner_harris(const void* src, size_t sstep, int w, int h, int type, float* dst, size_t dstep, int bs, int ks, double k, int border)
{ return corner_impl(src, sstep, w, h, type, dst, dstep, bs, ks, k, border, 0); }
PORT_API int port_corner_min_eigen_val(const void* src, size_t sstep, int w, int h, int type, float* dst, size_t dstep, int bs, int ks, int border)
{ return corner_impl(src, sstep, w, h, type, dst, dstep, bs, ks, 0, border, 1); }

typedef struct { float v; int pos; } cand_t;
static int cand_cmp(const void* pa, const void* pb)
{
    const cand_t* a = (const cand_t*)pa; const cand_t* b = (const cand_t*)pb;
    if (a->v > b->v) return -1; if (a->v < b->v) return 1;
    return a->pos > b->pos ? -1 : a->pos < b->pos ? 1 : 0;
}

PORT_API int port_good_features_to_track(const void* src, size_t sstep, int w, int h, int type, float* corners, float* quality, int max_out, int* nout,
                                         int max_corners, double ql, double min_dist, int bs, int gs, int harris, double k)
{
    float* eig = (float*)malloc(sizeof(float) * (size_t)w * h);
    int rc = corner_impl(src, sstep, w, h, type, eig, (size_t)w * 4, bs, gs, k, PB_REFLECT_101, harris ? 0 : 1);
    if (rc) { free(eig); return rc; }
    double maxv = -1e300;
    for (size_t i = 0; i < (size_t)w * h; i++) if (eig[i] > maxv) maxv = eig[i];
    float thr = (float)(maxv * ql);
    cand_t* cand = (cand_t*)malloc(sizeof(cand_t) * (size_t)w * h); size_t nc = 0;
    for (int y = 1; y < h - 1; y++)
        for (int x = 1; x < w - 1; x++) {
            float v = eig[(size_t)y * w + x];
            if (!(v > thr) || v == 0) continue;
            int ismax = 1;
            for (int dy = -1; dy <= 1 && ismax; dy++) for (int dx = -1; dx <= 1; dx++) { float nb = eig[(size_t)(y + dy) * w + x + dx]; nb = nb > thr ? nb : 0.f; if (nb > v) { ismax = 0; break; } }
            if (ismax) { cand[nc].v = v; cand[nc].pos = y * w + x; nc++; }
        }
    qsort(cand, nc, sizeof(cand_t), cand_cmp);
    int acc = 0;
    double md2 = min_dist * min_dist;
    float* ax = (float*)malloc(sizeof(float) * 2 * (nc + 1));
    for (size_t i = 0; i < nc; i++) {
        int y = cand[i].pos / w, x = cand[i].pos - y * w, good = 1;
        if (min_dist >= 1) for (int j = 0; j < acc; j++) { float dx = x - ax[2 * j], dy = y - ax[2 * j + 1]; if (dx * dx + dy * dy < md2) { good = 0; break; } }
        if (!good) continue;
        ax[2 * acc] = (float)x; ax[2 * acc + 1] = (float)y;
        if (acc < max_out) { corners[2 * acc] = (float)x; corners[2 * acc + 1] = (float)y; if (quality) quality[acc] = cand[i].v; }
        acc++;
        if (max_corners > 0 && acc == max_corners) break;
    }
    *nout = acc;
    free(ax); free(cand); free(eig);
    return 0;
}

PORT_API int port_sift_pyramid(const void* gray, size_t step, int w, int h, int nl, double sigma, int upscale, float* gauss, size_t* ge_out, float* dog,
                               size_t* de_out, int* n_oct, int* dims)
{
    int bw = upscale ? 2 * w : w, bh = upscale ? 2 * h : h;
    int no = port_round(log((double)(bw < bh ? bw : bh)) / log(2.) - 2) - (upscale ? -1 : 0);
    size_t ge = 0, de = 0; int cw = bw, ch = bh;
    for (int o = 0; o < no; o++) { if (dims) { dims[2 * o] = cw; dims[2 * o + 1] = ch; } ge += (size_t)cw * ch * (nl + 3); de += (size_t)cw * ch * (nl + 2); cw /= 2; ch /= 2; }
    if (ge_out) *ge_out = ge; if (de_out) *de_out = de; if (n_oct) *n_oct = no;
    if (!gauss && !dog) return 0;
    float* G = gauss ? gauss : (float*)malloc(sizeof(float) * ge);
    float* gf = (float*)malloc(sizeof(float) * (size_t)w * h);
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) gf[(size_t)y * w + x] = (float)((const uchar*)gray + (size_t)y * step)[x];
    float fs = (float)sigma;
    const int F32 = P_32F;
    if (upscale) {
        float sd = sqrtf(fmaxf(fs * fs - 0.5f * 0.5f * 4, 0.01f));
        float* dbl = (float*)malloc(sizeof(float) * (size_t)bw * bh);
        const double Mh[6] = {0.5, 0, 0, 0, 0.5, 0}, zero[4] = {0, 0, 0, 0};
        if (upscale == 2) port_resize(gf, (size_t)w * 4, w, h, dbl, (size_t)bw * 4, bw, bh, F32, 1);      /* enable_precise_upscale = false: cv::resize LINEAR */
        else port_warp_affine(gf, (size_t)w * 4, w, h, dbl, (size_t)bw * 4, bw, bh, F32, Mh, 1 | 16, PB_REFLECT, zero);
        port_gaussian_blur(dbl, (size_t)bw * 4, G, (size_t)bw * 4, bw, bh, F32, 0, 0, sd, sd, PB_REFLECT_101);
        free(dbl);
    } else {
        float sd = sqrtf(fmaxf(fs * fs - 0.5f * 0.5f, 0.01f));
        port_gaussian_blur(gf, (size_t)w * 4, G, (size_t)w * 4, w, h, F32, 0, 0, sd, sd, PB_REFLECT_101);
    }
    free(gf);
    double sig[16]; sig[0] = sigma; double kk = pow(2., 1. / nl);
    for (int i = 1; i < nl + 3; i++) { double sp = pow(kk, (double)(i - 1)) * sigma, st = sp * kk; sig[i] = sqrt(st * st - sp * sp); }
    size_t goff = 0, doff = 0; cw = bw; ch = bh;
    for (int o = 0; o < no; o++) {
        size_t n = (size_t)cw * ch;
        if (o > 0) {
            int pw = cw * 2 + (dims ? 0 : 0); (void)pw;
            int ppw = dims ? dims[2 * (o - 1)] : cw * 2, pph = dims ? dims[2 * (o - 1) + 1] : ch * 2;
            size_t pn = (size_t)ppw * pph;
            port_resize(G + goff - pn * (nl + 3) + pn * nl, (size_t)ppw * 4, ppw, pph, G + goff, (size_t)cw * 4, cw, ch, F32, 0);
        }
        for (int i = 1; i < nl + 3; i++) port_gaussian_blur(G + goff + (size_t)(i - 1) * n, (size_t)cw * 4, G + goff + (size_t)i * n, (size_t)cw * 4, cw, ch, F32, 0, 0, sig[i], sig[i], PB_REFLECT_101);
        if (dog) for (int i = 0; i < nl + 2; i++) for (size_t e = 0; e < n; e++) dog[doff + (size_t)i * n + e] = G[goff + (size_t)(i + 1) * n + e] - G[goff + (size_t)i * n + e];
        goff += n * (nl + 3); doff += n * (nl + 2); cw /= 2; ch /= 2;
    }
    if (!gauss) free(G);
    return 0;
}

/* ---- SIFT scale-space extrema, sub-pixel refinement, orientation assignment (sift.simd.hpp:160-681, sift.dispatch.cpp:368-402, :529-560) -----
 * Input: the packed Gaussian / DoG pyramids (layout of port_sift_pyramid: octave-major, nl+3 / nl+2 images of dims[2o] x dims[2o+1]).
 * Output: keypoints as 6 floats (x, y, size, angle, response, octave bits) after KeyPointsFilter::removeDuplicatedSorted and the
 * first-octave rescaling, i.e. what cv::SIFT::detect returns (nfeatures = 0, no mask).
 * The reference's AVX2 / AVX-512 objects of this file are built with FMA contraction, its exp / atan2 are OpenCV's own approximations
 * (mathfuncs_core.simd.hpp): this restatement uses plain float expressions, fastAtan2's polynomial and expf.  It is therefore pinned by
 * tolerance, not bit for bit: tests/test_oracle.py states the agreement that is measured against the reference. */
typedef struct { float x, y, size, angle, response; int octave; } PortKp;

static float port_fast_atan2_deg(float y, float x)          /* atan_f32, mathfuncs_core.simd.hpp:52-72 */
{
    const float p1 = 0.9997878412794807f * (float)(180 / 3.1415926535897932384626433832795), p3 = -0.3258083974640975f * (float)(180 / 3.1415926535897932384626433832795);
    const float p5 = 0.1555786518463281f * (float)(180 / 3.1415926535897932384626433832795), p7 = -0.04432655554792128f * (float)(180 / 3.1415926535897932384626433832795);
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) { c = ay / (ax + (float)2.2204460492503131e-16); c2 = c * c; a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
    else { c = ax / (ay + (float)2.2204460492503131e-16); c2 = c * c; a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

static int sift_adjust(const float* dogo, int cw, int ch, int nl, int octv, int* layer, int* r_, int* c_, float contrastThreshold, float edgeThreshold,
                       float sigma, PortKp* kpt)
{
    const float img_scale = 1.f / 255, deriv_scale = img_scale * 0.5f, second_deriv_scale = img_scale, cross_deriv_scale = img_scale * 0.25f;
    const size_t n = (size_t)cw * ch;
    float xi = 0, xr = 0, xc = 0, contr = 0;
    int i = 0, r = *r_, c = *c_, l = *layer;
#define AT(P, R, C) (P)[(size_t)(R) * cw + (C)]
    for (; i < 5; i++) {
        const float* img = dogo + (size_t)l * n; const float* prev = img - n; const float* next = img + n;
        float dD[3] = {(AT(img, r, c + 1) - AT(img, r, c - 1)) * deriv_scale, (AT(img, r + 1, c) - AT(img, r - 1, c)) * deriv_scale,
                       (AT(next, r, c) - AT(prev, r, c)) * deriv_scale};
        float v2 = AT(img, r, c) * 2;
        float dxx = (AT(img, r, c + 1) + AT(img, r, c - 1) - v2) * second_deriv_scale;
        float dyy = (AT(img, r + 1, c) + AT(img, r - 1, c) - v2) * second_deriv_scale;
        float dss = (AT(next, r, c) + AT(prev, r, c) - v2) * second_deriv_scale;
        float dxy = (AT(img, r + 1, c + 1) - AT(img, r + 1, c - 1) - AT(img, r - 1, c + 1) + AT(img, r - 1, c - 1)) * cross_deriv_scale;
        float dxs = (AT(next, r, c + 1) - AT(next, r, c - 1) - AT(prev, r, c + 1) + AT(prev, r, c - 1)) * cross_deriv_scale;
        float dys = (AT(next, r + 1, c) - AT(next, r - 1, c) - AT(prev, r + 1, c) + AT(prev, r - 1, c)) * cross_deriv_scale;
        /* Matx33f::solve(DECOMP_LU) = Cramer's rule in float (operations.hpp:191-213, matx.inl.hpp:53-61) */
        float a00 = dxx, a01 = dxy, a02 = dxs, a10 = dxy, a11 = dyy, a12 = dys, a20 = dxs, a21 = dys, a22 = dss;
        float d = (float)(double)(a00 * (a11 * a22 - a21 * a12) - a01 * (a10 * a22 - a20 * a12) + a02 * (a10 * a21 - a20 * a11));
        float X[3] = {0, 0, 0};
        if (d != 0) {
            d = 1 / d;
            X[0] = d * (dD[0] * (a11 * a22 - a12 * a21) - a01 * (dD[1] * a22 - a12 * dD[2]) + a02 * (dD[1] * a21 - a11 * dD[2]));
            X[1] = d * (a00 * (dD[1] * a22 - a12 * dD[2]) - dD[0] * (a10 * a22 - a12 * a20) + a02 * (a10 * dD[2] - dD[1] * a20));
            X[2] = d * (a00 * (a11 * dD[2] - dD[1] * a21) - a01 * (a10 * dD[2] - dD[1] * a20) + dD[0] * (a10 * a21 - a11 * a20));
        }
        xi = -X[2]; xr = -X[1]; xc = -X[0];
        if (fabsf(xi) < 0.5f && fabsf(xr) < 0.5f && fabsf(xc) < 0.5f) break;
        if (fabsf(xi) > (float)(2147483647 / 3) || fabsf(xr) > (float)(2147483647 / 3) || fabsf(xc) > (float)(2147483647 / 3)) return 0;
        c += (int)lrintf(xc); r += (int)lrintf(xr); l += (int)lrintf(xi);
        if (l < 1 || l > nl || c < 5 || c >= cw - 5 || r < 5 || r >= ch - 5) return 0;
    }
    if (i >= 5) return 0;
    {
        const float* img = dogo + (size_t)l * n; const float* prev = img - n; const float* next = img + n;
        float dD[3] = {(AT(img, r, c + 1) - AT(img, r, c - 1)) * deriv_scale, (AT(img, r + 1, c) - AT(img, r - 1, c)) * deriv_scale,
                       (AT(next, r, c) - AT(prev, r, c)) * deriv_scale};
        float t = dD[0] * xc + dD[1] * xr + dD[2] * xi;
        contr = AT(img, r, c) * img_scale + t * 0.5f;
        if (fabsf(contr) * nl < contrastThreshold) return 0;
        float v2 = AT(img, r, c) * 2.f;
        float dxx = (AT(img, r, c + 1) + AT(img, r, c - 1) - v2) * second_deriv_scale;
        float dyy = (AT(img, r + 1, c) + AT(img, r - 1, c) - v2) * second_deriv_scale;
        float dxy = (AT(img, r + 1, c + 1) - AT(img, r + 1, c - 1) - AT(img, r - 1, c + 1) + AT(img, r - 1, c - 1)) * cross_deriv_scale;
        float tr = dxx + dyy, det = dxx * dyy - dxy * dxy;
        if (det <= 0 || tr * tr * edgeThreshold >= (edgeThreshold + 1) * (edgeThreshold + 1) * det) return 0;
    }
#undef AT
    kpt->x = (c + xc) * (1 << octv);
    kpt->y = (r + xr) * (1 << octv);
    kpt->octave = octv + (l << 8) + ((int)lrint((xi + 0.5) * 255) << 16);
    kpt->size = sigma * powf(2.f, (l + xi) / nl) * (1 << octv) * 2;
    kpt->response = fabsf(contr);
    *layer = l; *r_ = r; *c_ = c;
    return 1;
}

static float sift_ori_hist(const float* img, int cw, int ch, int px, int py, int radius, float sigma, float* hist, int n)
{
    float temp[36 + 4]; float* th = temp + 2;
    const float expf_scale = -1.f / (2.f * sigma * sigma);
    for (int i = 0; i < n; i++) th[i] = 0.f;
    for (int i = -radius; i <= radius; i++) {
        int y = py + i;
        if (y <= 0 || y >= ch - 1) continue;
        for (int j = -radius; j <= radius; j++) {
            int x = px + j;
            if (x <= 0 || x >= cw - 1) continue;
            float dx = img[(size_t)y * cw + x + 1] - img[(size_t)y * cw + x - 1];
            float dy = img[(size_t)(y - 1) * cw + x] - img[(size_t)(y + 1) * cw + x];
            float w = expf((i * i + j * j) * expf_scale), ori = port_fast_atan2_deg(dy, dx), mag = sqrtf(dx * dx + dy * dy);
            int bin = (int)lrintf((n / 360.f) * ori);
            if (bin >= n) bin -= n;
            if (bin < 0) bin += n;
            th[bin] += w * mag;
        }
    }
    th[-1] = th[n - 1]; th[-2] = th[n - 2]; th[n] = th[0]; th[n + 1] = th[1];
    for (int i = 0; i < n; i++) hist[i] = (th[i - 2] + th[i + 2]) * (1.f / 16.f) + (th[i - 1] + th[i + 1]) * (4.f / 16.f) + th[i] * (6.f / 16.f);
    float m = hist[0];
    for (int i = 1; i < n; i++) if (hist[i] > m) m = hist[i];
    return m;
}

static int kp_cmp(const void* pa, const void* pb)        /* KeyPoint12_LessThan, keypoint.cpp:253-271 (class_id is always -1 here) */
{
    const PortKp* a = (const PortKp*)pa; const PortKp* b = (const PortKp*)pb;
    if (a->x != b->x) return a->x < b->x ? -1 : 1;
    if (a->y != b->y) return a->y < b->y ? -1 : 1;
    if (a->size != b->size) return a->size > b->size ? -1 : 1;
    if (a->angle != b->angle) return a->angle < b->angle ? -1 : 1;
    if (a->response != b->response) return a->response > b->response ? -1 : 1;
    if (a->octave != b->octave) return a->octave > b->octave ? -1 : 1;
    return 0;
}

PORT_API int port_sift_detect(const float* gauss, const float* dog, const int* dims, int n_oct, int nl, double contrastThreshold, double edgeThreshold,
                              double sigma, int first_octave, int nfeatures, int max_kp, float* kp_out, int* n_out)
{
    const int threshold = (int)floor(0.5 * contrastThreshold / nl * 255);
    size_t cap = 1024, cnt = 0;
    PortKp* kps = (PortKp*)malloc(sizeof(PortKp) * cap);
    size_t goff = 0, doff = 0;
    for (int o = 0; o < n_oct; o++) {
        const int cw = dims[2 * o], ch = dims[2 * o + 1];
        const size_t n = (size_t)cw * ch;
        const float* dogo = dog + doff; const float* go = gauss + goff;
        for (int i = 1; i <= nl; i++) {
            const float* img = dogo + (size_t)i * n; const float* prev = img - n; const float* next = img + n;
            for (int r = 5; r < ch - 5; r++)
                for (int c = 5; c < cw - 5; c++) {
                    const float val = img[(size_t)r * cw + c];
                    if (fabsf(val) <= threshold) continue;
                    int ok = 1;
                    for (int dz = 0; dz < 3 && ok; dz++) {
                        const float* p = dz == 0 ? img : dz == 1 ? prev : next;
                        for (int dy = -1; dy <= 1 && ok; dy++)
                            for (int dx = -1; dx <= 1; dx++) {
                                const float v = p[(size_t)(r + dy) * cw + c + dx];
                                if (val > 0 ? val < v : val > v) { ok = 0; break; }
                            }
                    }
                    if (!ok) continue;
                    PortKp kpt; int r1 = r, c1 = c, layer = i;
                    if (!sift_adjust(dogo, cw, ch, nl, o, &layer, &r1, &c1, (float)contrastThreshold, (float)edgeThreshold, (float)sigma, &kpt)) continue;
                    float scl_octv = kpt.size * 0.5f / (1 << o), hist[36];
                    float omax = sift_ori_hist(go + (size_t)layer * n, cw, ch, c1, r1, (int)lrintf(4.5f * scl_octv), 1.5f * scl_octv, hist, 36);
                    float mag_thr = omax * 0.8f;
                    for (int j = 0; j < 36; j++) {
                        int l = j > 0 ? j - 1 : 35, r2 = j < 35 ? j + 1 : 0;
                        if (hist[j] > hist[l] && hist[j] > hist[r2] && hist[j] >= mag_thr) {
                            float bin = j + 0.5f * (hist[l] - hist[r2]) / (hist[l] - 2 * hist[j] + hist[r2]);
                            bin = bin < 0 ? 36 + bin : bin >= 36 ? bin - 36 : bin;
                            kpt.angle = 360.f - (float)((360.f / 36) * bin);
                            if (fabsf(kpt.angle - 360.f) < 1.1920929e-07f) kpt.angle = 0.f;
                            if (cnt == cap) { cap *= 2; kps = (PortKp*)realloc(kps, sizeof(PortKp) * cap); }
                            kps[cnt++] = kpt;
                        }
                    }
                }
        }
        goff += n * (nl + 3); doff += n * (nl + 2);
    }
    /* KeyPointsFilter::removeDuplicatedSorted (keypoint.cpp:273-291), then the first-octave rescaling (sift.dispatch.cpp:548-555) */
    size_t m = 0;
    if (cnt) {
        qsort(kps, cnt, sizeof(PortKp), kp_cmp);
        for (size_t j = 1; j < cnt; j++)
            if (kps[m].x != kps[j].x || kps[m].y != kps[j].y || kps[m].size != kps[j].size || kps[m].angle != kps[j].angle) kps[++m] = kps[j];
        m++;
    }
    /* KeyPointsFilter::retainBest (keypoint.cpp:69-90): everything at least as strong as the nfeatures-th response.  The reference's order after
     * std::nth_element / std::partition is the library's; here the survivors keep the sorted order: compare as sets. */
    if (nfeatures > 0 && m > (size_t)nfeatures) {
        float* resp = (float*)malloc(sizeof(float) * m);
        for (size_t j = 0; j < m; j++) resp[j] = kps[j].response;
        for (size_t a = 0; a < (size_t)nfeatures; a++) {        /* partial selection sort of the responses, descending */
            size_t best = a;
            for (size_t b2 = a + 1; b2 < m; b2++) if (resp[b2] > resp[best]) best = b2;
            float t = resp[a]; resp[a] = resp[best]; resp[best] = t;
        }
        const float ambiguous = resp[nfeatures - 1];
        free(resp);
        size_t w = 0;
        for (size_t j = 0; j < m; j++) if (kps[j].response >= ambiguous) kps[w++] = kps[j];
        m = w;
    }
    if (first_octave < 0) {
        const float scale = 1.f / (float)(1 << -first_octave);
        for (size_t j = 0; j < m; j++) {
            kps[j].octave = (kps[j].octave & ~255) | ((kps[j].octave + first_octave) & 255);
            kps[j].x *= scale; kps[j].y *= scale; kps[j].size *= scale;
        }
    }
    *n_out = (int)m;
    for (size_t j = 0; j < m && j < (size_t)max_kp; j++) {
        float* o6 = kp_out + 6 * j;
        o6[0] = kps[j].x; o6[1] = kps[j].y; o6[2] = kps[j].size; o6[3] = kps[j].angle; o6[4] = kps[j].response; memcpy(o6 + 5, &kps[j].octave, 4);
    }
    free(kps);
    return 0;
}

/* ---- SIFT descriptors (calcSIFTDescriptor, sift.simd.hpp:709-1035; calcDescriptorsComputer, sift.dispatch.cpp:417-462): 4 x 4 cells x 8
 * orientation bins, Gaussian-weighted gradient magnitudes spread by tri-linear interpolation, clipped at 0.2 of the norm, renormalised to
 * 512 and saturated to bytes (stored as floats).  Same remark on exactness as port_sift_detect: plain float, fastAtan2's polynomial, expf. */
PORT_API int port_sift_descriptors(const float* gauss, const int* dims, int n_oct, int nl, int first_octave, const float* kp, int nkp, float* desc)
{
    enum { d = 4, n = 8 };
    size_t* offs = (size_t*)malloc(sizeof(size_t) * (size_t)n_oct);
    size_t go = 0;
    for (int o = 0; o < n_oct; o++) { offs[o] = go; go += (size_t)dims[2 * o] * dims[2 * o + 1] * (nl + 3); }
    for (int q = 0; q < nkp; q++) {
        const float* k6 = kp + 6 * q; int oct; memcpy(&oct, k6 + 5, 4);
        int octave = oct & 255, layer = (oct >> 8) & 255;
        octave = octave < 128 ? octave : (-128 | octave);
        const float scale = octave >= 0 ? 1.f / (1 << octave) : (float)(1 << -octave);
        const int oi = octave - first_octave;
        if (oi < 0 || oi >= n_oct || layer > nl + 2) { free(offs); return -1; }
        const int cols = dims[2 * oi], rows = dims[2 * oi + 1];
        const float* img = gauss + offs[oi] + (size_t)layer * cols * rows;
        const float size = k6[2] * scale, ptx = k6[0] * scale, pty = k6[1] * scale;
        float ori = 360.f - k6[3];
        if (fabsf(ori - 360.f) < 1.1920929e-07f) ori = 0.f;
        const float scl = size * 0.5f;
        const int px = (int)lrintf(ptx), py = (int)lrintf(pty);
        float cos_t = cosf(ori * (float)(3.1415926535897932384626433832795 / 180)), sin_t = sinf(ori * (float)(3.1415926535897932384626433832795 / 180));
        const float bins_per_rad = n / 360.f, exp_scale = -1.f / (d * d * 0.5f), hist_width = 3.f * scl;
        int radius = (int)lrintf(hist_width * 1.4142135623730951f * (d + 1) * 0.5f);
        const int diag = (int)sqrt((double)cols * cols + (double)rows * rows);
        if (radius > diag) radius = diag;
        cos_t /= hist_width; sin_t /= hist_width;
        float hist[(d + 2) * (d + 2) * (n + 2)], raw[d * d * n];
        for (int i = 0; i < (d + 2) * (d + 2) * (n + 2); i++) hist[i] = 0.f;
        for (int i = -radius; i <= radius; i++)
            for (int j = -radius; j <= radius; j++) {
                float c_rot = j * cos_t - i * sin_t, r_rot = j * sin_t + i * cos_t;
                float rbin = r_rot + d / 2 - 0.5f, cbin = c_rot + d / 2 - 0.5f;
                int r = py + i, c = px + j;
                if (!(rbin > -1 && rbin < d && cbin > -1 && cbin < d && r > 0 && r < rows - 1 && c > 0 && c < cols - 1)) continue;
                float dx = img[(size_t)r * cols + c + 1] - img[(size_t)r * cols + c - 1];
                float dy = img[(size_t)(r - 1) * cols + c] - img[(size_t)(r + 1) * cols + c];
                float w = expf((c_rot * c_rot + r_rot * r_rot) * exp_scale);
                float obin = (port_fast_atan2_deg(dy, dx) - ori) * bins_per_rad, mag = sqrtf(dx * dx + dy * dy) * w;
                int r0 = (int)floorf(rbin), c0 = (int)floorf(cbin), o0 = (int)floorf(obin);
                rbin -= r0; cbin -= c0; obin -= o0;
                if (o0 < 0) o0 += n;
                if (o0 >= n) o0 -= n;
                /* tri-linear split: the upper share of each axis is weight * fraction, the lower share the remainder (sift.simd.hpp:864-882) */
                const int idx = ((r0 + 1) * (d + 2) + c0 + 1) * (n + 2) + o0;
                const float up_r = mag * rbin, wr[2] = {mag - up_r, up_r};
                for (int ri = 0; ri < 2; ri++) {
                    const float up_c = wr[ri] * cbin, wc[2] = {wr[ri] - up_c, up_c};
                    for (int ci = 0; ci < 2; ci++) {
                        const float up_o = wc[ci] * obin;
                        float* cell = hist + idx + ri * (d + 2) * (n + 2) + ci * (n + 2);
                        cell[0] += wc[ci] - up_o; cell[1] += up_o;
                    }
                }
            }
        for (int i = 0; i < d; i++)
            for (int j = 0; j < d; j++) {
                int idx = ((i + 1) * (d + 2) + (j + 1)) * (n + 2);
                hist[idx] += hist[idx + n]; hist[idx + 1] += hist[idx + n + 1];
                for (int k = 0; k < n; k++) raw[(i * d + j) * n + k] = hist[idx + k];
            }
        float nrm2 = 0;
        for (int k = 0; k < d * d * n; k++) nrm2 += raw[k] * raw[k];
        const float thr = sqrtf(nrm2) * 0.2f;
        nrm2 = 0;
        for (int k = 0; k < d * d * n; k++) { float v = raw[k] < thr ? raw[k] : thr; raw[k] = v; nrm2 += v * v; }
        float s = sqrtf(nrm2); if (s < 1.1920929e-07f) s = 1.1920929e-07f;
        nrm2 = 512.f / s;
        for (int k = 0; k < d * d * n; k++) desc[(size_t)q * 128 + k] = (float)port_sat_u8f(raw[k] * nrm2);
    }
    free(offs);
    return 0;
}

/* ---- cv::matchTemplate with a mask (matchTemplateMask, templmatch.cpp:762-905), one channel ---------------------------------------------
 * The reference turns everything into float and evaluates up to four cross-correlations with its block DFT:
 *   S1 = CC(I, W1), S2 = CC(I^2, M^2), S3 = CC(I, M), S4 = CC(I, M^2);  W1 = T M^2 (SQDIFF / CCORR) or M^2 (T - mu), mu = sum(M T) / sum(M) (CCOEFF)
 *   SQDIFF  -2 S1 + S2 + c              c = sum((T M)^2)            NORMED: / sqrt(c S2)
 *   CCORR   S1                                                        NORMED: / sqrt(c S2)
 *   CCOEFF  S1 - S3 sum(W1) / sum(M)                                  NORMED: / (sqrt(S2 + S3 / sum(M) (S3 sum(M^2) / sum(M) - 2 S4)) |M (T - mu)|)
 * An 8-bit mask is binarised (non-zero -> 1), a float mask is a weight.  Here the sums are direct, in double: parity with the reference is by
 * tolerance (its own test allows the DFT's error), as for the unmasked methods. */
PORT_API int port_match_template_masked(const void* img, size_t istep, int iw, int ih, const void* tpl, size_t tstep, int tw, int th, int type,
                                        const void* mask, size_t mstep, int mask_type, float* result, size_t rstep, int method)
{
    const int depth = P_DEPTH(type), mdepth = P_DEPTH(mask_type);
    if (P_CN(type) != 1 || P_CN(mask_type) != 1 || (depth != P_8U && depth != P_32F) || (mdepth != P_8U && mdepth != P_32F) || method < 0 || method > 5) return 1;
    const int ow = iw - tw + 1, oh = ih - th + 1, n = tw * th;
    float* T = (float*)malloc(sizeof(float) * (size_t)n * 4); float* M = T + n; float* M2 = M + n; float* W1 = M2 + n;
    for (int y = 0; y < th; y++)
        for (int x = 0; x < tw; x++) {
            T[y * tw + x] = depth == P_8U ? (float)((const uchar*)tpl + (size_t)y * tstep)[x] : ((const float*)((const char*)tpl + (size_t)y * tstep))[x];
            M[y * tw + x] = mdepth == P_8U ? (((const uchar*)mask + (size_t)y * mstep)[x] ? 1.f : 0.f) : ((const float*)((const char*)mask + (size_t)y * mstep))[x];
        }
    double sumM = 0, sumMT = 0, sumM2 = 0, c = 0;
    for (int i = 0; i < n; i++) { M2[i] = M[i] * M[i]; sumM += M[i]; sumMT += (double)(M[i] * T[i]); sumM2 += M2[i]; float tm = T[i] * M[i]; c += (double)tm * tm; }
    const int coeff = method >= 4;
    const float mu = (float)(sumMT / sumM);
    double sumW1 = 0, nt2 = 0;
    for (int i = 0; i < n; i++) {
        W1[i] = coeff ? M[i] * (M[i] * (T[i] - mu)) : T[i] * M2[i];
        sumW1 += W1[i];
        float q = M[i] * (T[i] - mu); nt2 += (double)q * q;
    }
    const double norm_templx = sqrt(nt2);
    for (int y = 0; y < oh; y++)
        for (int x = 0; x < ow; x++) {
            double S1 = 0, S2 = 0, S3 = 0, S4 = 0;
            for (int v = 0; v < th; v++)
                for (int u = 0; u < tw; u++) {
                    const double I = depth == P_8U ? (double)((const uchar*)img + (size_t)(y + v) * istep)[x + u]
                                                   : (double)((const float*)((const char*)img + (size_t)(y + v) * istep))[x + u];
                    const int i = v * tw + u;
                    S1 += I * W1[i]; S2 += I * I * M2[i]; S3 += I * M[i]; S4 += I * M2[i];
                }
            double r;
            if (method <= 1) { r = -2 * S1 + S2 + c; if (method == 1) r /= sqrt(c * S2); }
            else if (method <= 3) { r = S1; if (method == 3) r /= sqrt(c * S2); }
            else {
                r = S1 - S3 * (sumW1 / sumM);
                if (method == 5) { const double nimg = S2 + (S3 / sumM) * (S3 * (sumM2 / sumM) - 2 * S4); r /= sqrt(nimg) * norm_templx; }
            }
            ((float*)((char*)result + (size_t)y * rstep))[x] = (float)r;
        }
    free(T);
    return 0;
}
