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
        port_warp_affine(gf, (size_t)w * 4, w, h, dbl, (size_t)bw * 4, bw, bh, F32, Mh, 1 | 16, PB_REFLECT, zero);
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
