// warp.cu -- cv::warpAffine / cv::warpPerspective (INTER_NEAREST / LINEAR / CUBIC; 8-bit and float; 1/3/4 channels).
//
// One fused kernel per call (coordinate generation + gather + blend); the reference does the same per 64x64-ish block
// through cv::remap (WarpAffineInvoker imgwarp.cpp:2233-2298, WarpPerspectiveInvoker :3160-3226).
//
// Coordinates reproduce the reference's fixed-point pipeline exactly (explicit _rn fp64 intrinsics, no contraction):
//   affine       adelta=rint(M0*x*1024), bdelta=rint(M3*x*1024), X0=rint((M1*y+M2)*1024)+rd, Y0=rint((M4*y+M5)*1024)+rd,
//                rd = 512 (NEAREST) | 16 (others); NEAREST: (X0+adelta)>>10; others: X=(X0+adelta)>>5, sx=X>>5, fx=X&31
//                (hal::warpAffine :2673-2700, warpAffineBlockline[NN] :2702-2782)
//   perspective  block-relative fp64: x_b = x - x%bw0, X0=M0*x_b+M1*y+M2 (same for Y0,W0); W=W0+M6*x1; W = W ? 32/W : 0
//                (1/W for NEAREST); X=rint(clamp((X0+M0*x1)*W, INT_MIN, INT_MAX))  (:3199-3201, :3299-3365)
// Sampling follows remapNearest / remapBilinear / remapBicubic (:329-430, :675-904, :907-1010): 5-bit sub-pixel index into
// the 32x32 tap tables (initInterTab2D :213-287; built on the host with the same float code and uploaded once),
// u8: sat_u8((sum + 2^14) >> 15); f32: float sums in the reference's order; borders CONSTANT/REPLICATE/REFLECT/REFLECT_101/WRAP.
#include <vector>
#include "common.cuh"
#include "host_tables.h"

namespace b200cv {

__device__ short g_bilin_i[1024 * 4];
__device__ float g_bilin_f[1024 * 4];
__device__ short g_bicub_i[1024 * 16];
__device__ float g_bicub_f[1024 * 16];

struct WarpParams {
    double M[9];
    float cval_f[4];
    int cval_i[4];
    int sw, sh, dw, dh;
    int border, persp, bw0;
};

enum { W_NN = 0, W_LIN = 1, W_CUB = 2 };

template <typename T> struct TabOf;
template <> struct TabOf<uchar> { typedef short type; __device__ static const short* lin() { return g_bilin_i; } __device__ static const short* cub() { return g_bicub_i; } };
template <> struct TabOf<float> { typedef float type; __device__ static const float* lin() { return g_bilin_f; } __device__ static const float* cub() { return g_bicub_f; } };

__device__ __forceinline__ int clipi(int x, int a, int b) { return x >= a ? (x < b ? x : b - 1) : a; }

template <int INTERP>
__device__ __forceinline__ void warp_coords(const WarpParams& p, int x, int y, int& sx, int& sy, int& a)
{
    if (!p.persp) {
        const int rd = INTERP == W_NN ? 512 : 16;
        int adelta = __double2int_rn(__dmul_rn(__dmul_rn(p.M[0], (double)x), 1024.0));
        int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(p.M[3], (double)x), 1024.0));
        int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(p.M[1], (double)y), p.M[2]), 1024.0)) + rd;
        int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(p.M[4], (double)y), p.M[5]), 1024.0)) + rd;
        if (INTERP == W_NN) {
            sx = sat_s16((X0 + adelta) >> 10);
            sy = sat_s16((Y0 + bdelta) >> 10);
            a = 0;
        } else {
            int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
            sx = sat_s16(X >> 5);
            sy = sat_s16(Y >> 5);
            a = (Y & 31) * 32 + (X & 31);
        }
    } else {
        int xb = (x / p.bw0) * p.bw0, x1 = x - xb;
        double X0 = __dadd_rn(__dadd_rn(__dmul_rn(p.M[0], (double)xb), __dmul_rn(p.M[1], (double)y)), p.M[2]);
        double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(p.M[3], (double)xb), __dmul_rn(p.M[4], (double)y)), p.M[5]);
        double W0 = __dadd_rn(__dadd_rn(__dmul_rn(p.M[6], (double)xb), __dmul_rn(p.M[7], (double)y)), p.M[8]);
        double W = __dadd_rn(W0, __dmul_rn(p.M[6], (double)x1));
        W = W != 0.0 ? __ddiv_rn(INTERP == W_NN ? 1.0 : 32.0, W) : 0.0;
        double fX = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(X0, __dmul_rn(p.M[0], (double)x1)), W)));
        double fY = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(Y0, __dmul_rn(p.M[3], (double)x1)), W)));
        int X = __double2int_rn(fX), Y = __double2int_rn(fY);
        if (INTERP == W_NN) { sx = sat_s16(X); sy = sat_s16(Y); a = 0; }
        else { sx = sat_s16(X >> 5); sy = sat_s16(Y >> 5); a = (Y & 31) * 32 + (X & 31); }
    }
}

template <typename T, int CN, int INTERP>
__global__ void __launch_bounds__(256) warp_kernel(Img src, Img dst, const __grid_constant__ WarpParams p)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, f = blockIdx.z;
    if (x >= p.dw) return;
    int sx, sy, a;
    warp_coords<INTERP>(p, x, y, sx, sy, a);
    T* d = dst.row<T>(f, y) + (size_t)x * CN;
    const int sw = p.sw, sh = p.sh, border = p.border;
    T cval[4];
#pragma unroll
    for (int c = 0; c < 4; c++) { if constexpr (sizeof(T) == 1) cval[c] = (T)p.cval_i[c]; else cval[c] = p.cval_f[c]; }

    if constexpr (INTERP == W_NN) {
        const T* s;
        if ((unsigned)sx < (unsigned)sw && (unsigned)sy < (unsigned)sh) s = src.row<T>(f, sy) + (size_t)sx * CN;
        else if (border == B200CV_BORDER_REPLICATE) s = src.row<T>(f, clipi(sy, 0, sh)) + (size_t)clipi(sx, 0, sw) * CN;
        else if (border == B200CV_BORDER_CONSTANT) s = nullptr;
        else s = src.row<T>(f, border_interpolate(sy, sh, border)) + (size_t)border_interpolate(sx, sw, border) * CN;
#pragma unroll
        for (int c = 0; c < CN; c++) d[c] = s ? s[c] : cval[c];
    } else if constexpr (INTERP == W_LIN) {
        typedef typename TabOf<T>::type AT;
        const AT* w = TabOf<T>::lin() + a * 4;
        if (border == B200CV_BORDER_CONSTANT && (sx >= sw || sx + 1 < 0 || sy >= sh || sy + 1 < 0)) {
#pragma unroll
            for (int c = 0; c < CN; c++) d[c] = cval[c];
            return;
        }
        int sx0, sx1, sy0, sy1;
        if ((unsigned)sx < (unsigned)(sw - 1) && (unsigned)sy < (unsigned)(sh - 1)) { sx0 = sx; sx1 = sx + 1; sy0 = sy; sy1 = sy + 1; }
        else if (border == B200CV_BORDER_REPLICATE) { sx0 = clipi(sx, 0, sw); sx1 = clipi(sx + 1, 0, sw); sy0 = clipi(sy, 0, sh); sy1 = clipi(sy + 1, 0, sh); }
        else {
            sx0 = border_interpolate(sx, sw, border); sx1 = border_interpolate(sx + 1, sw, border);
            sy0 = border_interpolate(sy, sh, border); sy1 = border_interpolate(sy + 1, sh, border);
        }
        const T* r0 = sy0 >= 0 ? src.row<T>(f, sy0) : nullptr;
        const T* r1 = sy1 >= 0 ? src.row<T>(f, sy1) : nullptr;
#pragma unroll
        for (int c = 0; c < CN; c++) {
            T v0 = (r0 && sx0 >= 0) ? r0[sx0 * CN + c] : cval[c];
            T v1 = (r0 && sx1 >= 0) ? r0[sx1 * CN + c] : cval[c];
            T v2 = (r1 && sx0 >= 0) ? r1[sx0 * CN + c] : cval[c];
            T v3 = (r1 && sx1 >= 0) ? r1[sx1 * CN + c] : cval[c];
            if constexpr (sizeof(T) == 1) d[c] = sat_u8((v0 * w[0] + v1 * w[1] + v2 * w[2] + v3 * w[3] + (1 << 14)) >> 15);
            else d[c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v0, w[0]), __fmul_rn(v1, w[1])), __fmul_rn(v2, w[2])), __fmul_rn(v3, w[3]));
        }
    } else {
        typedef typename TabOf<T>::type AT;
        const AT* w = TabOf<T>::cub() + a * 16;
        sx -= 1; sy -= 1;
        const bool inlier = (unsigned)sx < (unsigned)max(sw - 3, 0) && (unsigned)sy < (unsigned)max(sh - 3, 0);
        if (!inlier && border == B200CV_BORDER_CONSTANT && (sx >= sw || sx + 4 <= 0 || sy >= sh || sy + 4 <= 0)) {
#pragma unroll
            for (int c = 0; c < CN; c++) d[c] = cval[c];
            return;
        }
        int xs[4], ys[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            xs[i] = inlier ? sx + i : border_interpolate(sx + i, sw, border);
            ys[i] = inlier ? sy + i : border_interpolate(sy + i, sh, border);
        }
#pragma unroll
        for (int c = 0; c < CN; c++) {
            if constexpr (sizeof(T) == 1) {
                int sum = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uchar* r = ys[i] >= 0 ? src.row<uchar>(f, ys[i]) : nullptr;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        int v = (r && xs[j] >= 0) ? r[xs[j] * CN + c] : cval[c];
                        sum += v * w[i * 4 + j];
                    }
                }
                d[c] = sat_u8((sum + (1 << 14)) >> 15);
            } else {
                if (inlier) {
                    float sum = 0.f;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const float* r = src.row<float>(f, ys[i]) + (size_t)xs[0] * CN + c;
                        float rs = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r[0], w[i * 4]), __fmul_rn(r[CN], w[i * 4 + 1])),
                                                       __fmul_rn(r[2 * CN], w[i * 4 + 2])), __fmul_rn(r[3 * CN], w[i * 4 + 3]));
                        sum = i == 0 ? rs : __fadd_rn(sum, rs);
                    }
                    d[c] = sum;
                } else {
                    float cv = cval[c], sum = cv;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        if (ys[i] < 0) continue;
                        const float* r = src.row<float>(f, ys[i]);
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            if (xs[j] >= 0) sum = __fadd_rn(sum, __fmul_rn(__fsub_rn(r[xs[j] * CN + c], cv), w[i * 4 + j]));
                    }
                    d[c] = sum;
                }
            }
        }
    }
}

static int ensure_warp_tables()
{
    static bool done = false;
    if (done) return B200CV_OK;
    std::vector<float> f; std::vector<short> q;
    bilinear_tab(f, q);
    B200_CUDA(cudaMemcpyToSymbol(g_bilin_f, f.data(), f.size() * sizeof(float)));
    B200_CUDA(cudaMemcpyToSymbol(g_bilin_i, q.data(), q.size() * sizeof(short)));
    bicubic_tab(f, q);
    B200_CUDA(cudaMemcpyToSymbol(g_bicub_f, f.data(), f.size() * sizeof(float)));
    B200_CUDA(cudaMemcpyToSymbol(g_bicub_i, q.data(), q.size() * sizeof(short)));
    done = true;
    return B200CV_OK;
}

template <typename T, int CN>
static int launch_warp(int interp, const Img& s, const Img& d, const WarpParams& p, cudaStream_t st)
{
    dim3 grid(div_up((unsigned)p.dw, 256), (unsigned)p.dh, (unsigned)s.frames);
    if (interp == W_NN) warp_kernel<T, CN, W_NN><<<grid, 256, 0, st>>>(s, d, p);
    else if (interp == W_LIN) warp_kernel<T, CN, W_LIN><<<grid, 256, 0, st>>>(s, d, p);
    else warp_kernel<T, CN, W_CUB><<<grid, 256, 0, st>>>(s, d, p);
    B200_LAUNCH_CHECK();
    return B200CV_OK;
}

static int warp_common(const b200cvMat* src, const b200cvMat* dst, const double* Minv, int persp, int flags, int border,
                       const double* bv, void* stream)
{
    int rc;
    if ((rc = check_mat(src, "src")) || (rc = check_mat(dst, "dst"))) return rc;
    B200_REQUIRE(src->type == dst->type, "warp: dst type must equal src type");
    B200_REQUIRE(src->data != dst->data, "warp: in-place is not supported");
    const int depth = B200CV_DEPTH(src->type), cn = B200CV_CN(src->type);
    if ((depth != B200CV_8U && depth != B200CV_32F) || (cn != 1 && cn != 3 && cn != 4)) return B200CV_NOT_IMPLEMENTED;
    int interp = flags & 7;
    if (interp == B200CV_INTER_AREA) interp = B200CV_INTER_LINEAR;            // imgwarp.cpp:2816, :3393
    if (interp > B200CV_INTER_CUBIC) return B200CV_NOT_IMPLEMENTED;
    border &= ~B200CV_BORDER_ISOLATED;
    if (border < 0 || border > B200CV_BORDER_REFLECT_101) return B200CV_NOT_IMPLEMENTED;   // BORDER_TRANSPARENT: not on the device path
    if (src->cols >= 32767 || src->rows >= 32767 || dst->rows >= 65536) return B200CV_NOT_IMPLEMENTED;   // CV_Assert(cols,rows < SHRT_MAX) imgwarp.cpp:1813
    if ((rc = ensure_warp_tables())) return rc;
    Img s = make_img(src), d = make_img(dst);
    B200_REQUIRE(s.frames == d.frames, "src/dst batch mismatch");
    WarpParams p;
    for (int i = 0; i < 9; i++) p.M[i] = Minv[i];
    for (int c = 0; c < 4; c++) {
        double v = bv ? bv[c] : 0.0;
        long r = lrint(v);
        p.cval_i[c] = (int)(r < 0 ? 0 : r > 255 ? 255 : r);
        p.cval_f[c] = (float)v;
    }
    p.sw = src->cols; p.sh = src->rows; p.dw = dst->cols; p.dh = dst->rows; p.border = border; p.persp = persp;
    {   // WarpPerspectiveInvoker block width (imgwarp.cpp:3182-3184)
        int bh0 = p.dh < 16 ? p.dh : 16;
        int bw0 = 1024 / bh0; if (bw0 > p.dw) bw0 = p.dw;
        p.bw0 = bw0;
    }
    cudaStream_t st = as_stream(stream);
    if (depth == B200CV_8U) {
        if (cn == 1) return launch_warp<uchar, 1>(interp, s, d, p, st);
        if (cn == 3) return launch_warp<uchar, 3>(interp, s, d, p, st);
        return launch_warp<uchar, 4>(interp, s, d, p, st);
    }
    if (cn == 1) return launch_warp<float, 1>(interp, s, d, p, st);
    if (cn == 3) return launch_warp<float, 3>(interp, s, d, p, st);
    return launch_warp<float, 4>(interp, s, d, p, st);
}

}  // namespace b200cv

using namespace b200cv;

extern "C" int b200cv_warp_affine(const b200cvMat* src, const b200cvMat* dst, const double* M0, int flags, int border,
                                  const double* border_value, void* stream)
{
    B200_REQUIRE(M0, "null matrix");
    double M[9] = {M0[0], M0[1], M0[2], M0[3], M0[4], M0[5], 0, 0, 1};
    if (!(flags & B200CV_WARP_INVERSE_MAP)) {      // cv::warpAffine, imgwarp.cpp:2824-2834
        double D = M[0] * M[4] - M[1] * M[3];
        D = D != 0 ? 1. / D : 0;
        double A11 = M[4] * D, A22 = M[0] * D;
        M[0] = A11; M[1] *= -D;
        M[3] *= -D; M[4] = A22;
        double b1 = -M[0] * M[2] - M[1] * M[5];
        double b2 = -M[3] * M[2] - M[4] * M[5];
        M[2] = b1; M[5] = b2;
    }
    return warp_common(src, dst, M, 0, flags, border, border_value, stream);
}

extern "C" int b200cv_warp_perspective(const b200cvMat* src, const b200cvMat* dst, const double* M0, int flags, int border,
                                       const double* border_value, void* stream)
{
    B200_REQUIRE(M0, "null matrix");
    double M[9];
    for (int i = 0; i < 9; i++) M[i] = M0[i];
    if (!(flags & B200CV_WARP_INVERSE_MAP)) {      // cv::invert of a 3x3 CV_64F matrix: adjugate / determinant (core/src/lapack.cpp:944-970)
        const double* m = M0;
        double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
        if (det != 0.) {
            double d = 1. / det;
            M[0] = (m[4] * m[8] - m[5] * m[7]) * d; M[1] = (m[2] * m[7] - m[1] * m[8]) * d; M[2] = (m[1] * m[5] - m[2] * m[4]) * d;
            M[3] = (m[5] * m[6] - m[3] * m[8]) * d; M[4] = (m[0] * m[8] - m[2] * m[6]) * d; M[5] = (m[2] * m[3] - m[0] * m[5]) * d;
            M[6] = (m[3] * m[7] - m[4] * m[6]) * d; M[7] = (m[1] * m[6] - m[0] * m[7]) * d; M[8] = (m[0] * m[4] - m[1] * m[3]) * d;
        } else {
            for (int i = 0; i < 9; i++) M[i] = 0;   // cv::invert leaves dst = 0 for a singular matrix
        }
    }
    return warp_common(src, dst, M, 1, flags, border, border_value, stream);
}
