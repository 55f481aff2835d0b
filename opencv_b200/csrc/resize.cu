// resize.cu -- cv::resize: INTER_NEAREST, INTER_LINEAR (incl. the exact-1/2 case the reference rewrites to
// INTER_AREA), INTER_AREA with an exact 2x2 box, INTER_CUBIC; 8-bit and float, 1/3/4 channels.
//
// Everything the reference tabulates on the host (source offsets and fixed-point / float taps per destination
// column and row, resize.cpp:4097-4190) is recomputed per thread here with the same IEEE operations in the same
// order (explicit _rn intrinsics: no FMA contraction), so no tables travel through HBM:
//   NEAREST   sx = min(floor(x * (1/fx)), sw-1) in fp64                               (resizeNN, resize.cpp:1121-1172)
//   LINEAR    fx = (float)((dx+0.5)*scale - 0.5); sx = floor(fx); fx -= sx; edge clamps (:4099-4124);
//             u8: taps = sat_s16(round(c*2048)), H pass int, V pass ((b0*(T0>>4))>>16)+((b1*(T1>>4))>>16)+2)>>2
//             (HResizeLinear :1877-1928, VResizeLinear<uchar> :1963-1989);  f32: T = S0*a0 + S1*a1 (mul, mul, add)
//   AREA 2x2  u8 (a+b+c+d+2)>>2 ; f32 ((a+b)+(c+d))*0.25 (1/4 ch) or (((a+b)+c)+d)*0.25   (:2919-3068, :2857-2900)
//   CUBIC     A=-0.75 taps (interpolateCubic :964-972), per-tap index clamping at the edges (HResizeCubic :1993-2041);
//             u8 V pass: the reference's SSE body evaluates S0*b0+(S1*b1+(S2*b2+S3*b3)) in float with b=beta/2^22 and
//             rounds half-even for the first floor8(width*cn) elements of a row, and uses (sum+2^21)>>22 for the tail
//             (VResizeCubicVec_32s8u :1408-1444, FixedPtCast :2059-2060) -- both are reproduced, so u8 CUBIC is bit-exact.
#include "common.cuh"

namespace b200cv {

struct ResizeParams {
    double ifx, ify;        // NEAREST: 1/fx, 1/fy
    double scale_x, scale_y;  // LINEAR/CUBIC: 1/inv_scale
    int sw, sh, dw, dh;
};

__device__ __forceinline__ int clip_i(int x, int a, int b) { return x >= a ? (x < b ? x : b - 1) : a; }

// ---- NEAREST ------------------------------------------------------------------------------------------------------
template <int PIX>   // pixel size in bytes
__global__ void __launch_bounds__(256) resize_nn_kernel(Img src, Img dst, ResizeParams p)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int f = blockIdx.z;
    if (x >= p.dw) return;
    int sx = min((int)floor(__dmul_rn((double)x, p.ifx)), p.sw - 1);
    int sy = min((int)floor(__dmul_rn((double)y, p.ify)), p.sh - 1);
    const uchar* s = src.row<uchar>(f, sy) + (size_t)sx * PIX;
    uchar* d = dst.row<uchar>(f, y) + (size_t)x * PIX;
    if constexpr (PIX == 4) *(uint32_t*)d = *(const uint32_t*)s;
    else if constexpr (PIX == 8) *(uint2*)d = *(const uint2*)s;
    else if constexpr (PIX == 16) *(uint4*)d = *(const uint4*)s;
    else if constexpr (PIX == 12) { const uint32_t* s4 = (const uint32_t*)s; uint32_t* d4 = (uint32_t*)d; d4[0] = s4[0]; d4[1] = s4[1]; d4[2] = s4[2]; }
    else {
#pragma unroll
        for (int i = 0; i < PIX; i++) d[i] = s[i];
    }
}

// ---- coefficient helpers -------------------------------------------------------------------------------------------
// linear: returns source index and fractional weight with the reference's edge clamps (ksize2 == 1)
__device__ __forceinline__ void linear_coef(int d, double scale, int ssize, int& s, float& fr, bool clamp_edges)
{
    float fx = (float)__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5);
    int sx = (int)floorf(fx);
    fx = __fsub_rn(fx, (float)sx);
    if (clamp_edges) {
        if (sx < 0) { fx = 0.f; sx = 0; }
        if (sx >= ssize - 1) { fx = 0.f; sx = ssize - 1; }
    }
    s = sx; fr = fx;
}

__device__ __forceinline__ void cubic_coeffs(float x, float* c)
{
    const float A = -0.75f;
    float x1 = __fadd_rn(x, 1.f);
    c[0] = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, x1), 5 * A), x1), 8 * A), x1), 4 * A);
    c[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2, x), A + 3), x), x), 1.f);
    float ix = __fsub_rn(1.f, x);
    c[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2, ix), A + 3), ix), ix), 1.f);
    c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c[0]), c[1]), c[2]);
}

__device__ __forceinline__ short coef_s16(float c) { return sat_s16(__float2int_rn(__fmul_rn(c, 2048.f))); }

// ---- LINEAR ----------------------------------------------------------------------------------------------------------
template <typename T, int CN>
__global__ void __launch_bounds__(256) resize_linear_kernel(Img src, Img dst, ResizeParams p)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int f = blockIdx.z;
    if (x >= p.dw) return;
    int sx, sy; float fx, fy;
    linear_coef(x, p.scale_x, p.sw, sx, fx, true);
    linear_coef(y, p.scale_y, p.sh, sy, fy, false);      // rows are clipped when fetched, the taps keep fy (:4160-4170, :2211)
    const int sy0 = clip_i(sy, 0, p.sh), sy1 = clip_i(sy + 1, 0, p.sh);
    const bool last_col = sx >= p.sw - 1;                  // dx >= xmax: D = S[sx] * ONE
    const T* r0 = src.row<T>(f, sy0) + (size_t)sx * CN;
    const T* r1 = src.row<T>(f, sy1) + (size_t)sx * CN;
    T* d = dst.row<T>(f, y) + (size_t)x * CN;
    if constexpr (sizeof(T) == 1) {
        const int a0 = coef_s16(__fsub_rn(1.f, fx)), a1 = coef_s16(fx);
        const int b0 = coef_s16(__fsub_rn(1.f, fy)), b1 = coef_s16(fy);
#pragma unroll
        for (int c = 0; c < CN; c++) {
            int t0 = last_col ? r0[c] * 2048 : r0[c] * a0 + r0[c + CN] * a1;
            int t1 = last_col ? r1[c] * 2048 : r1[c] * a0 + r1[c + CN] * a1;
            d[c] = (uchar)((((b0 * (t0 >> 4)) >> 16) + ((b1 * (t1 >> 4)) >> 16) + 2) >> 2);
        }
    } else {
        const float a0 = __fsub_rn(1.f, fx), a1 = fx, b0 = __fsub_rn(1.f, fy), b1 = fy;
#pragma unroll
        for (int c = 0; c < CN; c++) {
            float t0 = last_col ? r0[c] : __fadd_rn(__fmul_rn(r0[c], a0), __fmul_rn(r0[c + CN], a1));
            float t1 = last_col ? r1[c] : __fadd_rn(__fmul_rn(r1[c], a0), __fmul_rn(r1[c + CN], a1));
            d[c] = __fadd_rn(__fmul_rn(t0, b0), __fmul_rn(t1, b1));
        }
    }
}

// ---- AREA, exact 2x2 ---------------------------------------------------------------------------------------------
// u8: each thread produces 4 destination pixels from 2 rows x 8 source pixels (contiguous 8*CN bytes per row)
template <int CN>
__global__ void __launch_bounds__(256) resize_area2_u8_kernel(Img src, Img dst, int dw, int vec_ok)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y, f = blockIdx.z;
    if (x0 >= dw) return;
    const uchar* s0 = src.row<uchar>(f, 2 * y) + (size_t)x0 * 2 * CN;
    const uchar* s1 = src.row<uchar>(f, 2 * y + 1) + (size_t)x0 * 2 * CN;
    uchar* d = dst.row<uchar>(f, y) + (size_t)x0 * CN;
    const int n = min(4, dw - x0);
    if (vec_ok && n == 4) {
        constexpr int NB = 8 * CN;                 // source bytes per row: 8 / 24 / 32
        uint32_t a[NB / 4], b[NB / 4], o[CN];
#pragma unroll
        for (int i = 0; i < NB / 4; i++) { a[i] = __ldg((const uint32_t*)s0 + i); b[i] = __ldg((const uint32_t*)s1 + i); }
        const uchar* pa = (const uchar*)a; const uchar* pb = (const uchar*)b; uchar* po = (uchar*)o;
#pragma unroll
        for (int px = 0; px < 4; px++)
#pragma unroll
            for (int c = 0; c < CN; c++)
                po[px * CN + c] = (uchar)((pa[px * 2 * CN + c] + pa[px * 2 * CN + CN + c] + pb[px * 2 * CN + c] + pb[px * 2 * CN + CN + c] + 2) >> 2);
#pragma unroll
        for (int i = 0; i < CN; i++) ((uint32_t*)d)[i] = o[i];
    } else {
        for (int px = 0; px < n; px++)
#pragma unroll
            for (int c = 0; c < CN; c++)
                d[px * CN + c] = (uchar)((s0[px * 2 * CN + c] + s0[px * 2 * CN + CN + c] + s1[px * 2 * CN + c] + s1[px * 2 * CN + CN + c] + 2) >> 2);
    }
}

template <int CN>
__global__ void __launch_bounds__(256) resize_area2_f32_kernel(Img src, Img dst, int dw)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, f = blockIdx.z;
    if (x >= dw) return;
    const float* s0 = src.row<float>(f, 2 * y) + (size_t)x * 2 * CN;
    const float* s1 = src.row<float>(f, 2 * y + 1) + (size_t)x * 2 * CN;
    float* d = dst.row<float>(f, y) + (size_t)x * CN;
#pragma unroll
    for (int c = 0; c < CN; c++) {
        float a = s0[c], b = s0[c + CN], e = s1[c], g = s1[c + CN];
        float sum = (CN == 1 || CN == 4) ? __fadd_rn(__fadd_rn(a, b), __fadd_rn(e, g)) : __fadd_rn(__fadd_rn(__fadd_rn(a, b), e), g);
        d[c] = __fmul_rn(sum, 0.25f);
    }
}

// ---- CUBIC ---------------------------------------------------------------------------------------------------------------
template <typename T, int CN>
__global__ void __launch_bounds__(256) resize_cubic_kernel(Img src, Img dst, ResizeParams p)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int f = blockIdx.z;
    if (x >= p.dw) return;
    int sx, sy; float fx, fy;
    linear_coef(x, p.scale_x, p.sw, sx, fx, false);
    linear_coef(y, p.scale_y, p.sh, sy, fy, false);
    float cx[4], cy[4];
    cubic_coeffs(fx, cx);
    cubic_coeffs(fy, cy);
    int xi[4];
#pragma unroll
    for (int j = 0; j < 4; j++) xi[j] = min(max(sx - 1 + j, 0), p.sw - 1);    // per-tap clamping == the while-loops of HResizeCubic
    const T* rows[4];
#pragma unroll
    for (int k = 0; k < 4; k++) rows[k] = src.row<T>(f, clip_i(sy - 1 + k, 0, p.sh));
    T* d = dst.row<T>(f, y) + (size_t)x * CN;
    if constexpr (sizeof(T) == 1) {
        int ia[4], ib[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { ia[j] = coef_s16(cx[j]); ib[j] = coef_s16(cy[j]); }
        const int vec_limit = ((p.dw * CN) / 8) * 8;
        const float sc = 1.f / (2048.f * 2048.f);
#pragma unroll
        for (int c = 0; c < CN; c++) {
            int t[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uchar* r = rows[k];
                t[k] = r[xi[0] * CN + c] * ia[0] + r[xi[1] * CN + c] * ia[1] + r[xi[2] * CN + c] * ia[2] + r[xi[3] * CN + c] * ia[3];
            }
            if (x * CN + c < vec_limit) {
                float v = __fmul_rn((float)t[3], __fmul_rn((float)ib[3], sc));
                v = __fadd_rn(__fmul_rn((float)t[2], __fmul_rn((float)ib[2], sc)), v);
                v = __fadd_rn(__fmul_rn((float)t[1], __fmul_rn((float)ib[1], sc)), v);
                v = __fadd_rn(__fmul_rn((float)t[0], __fmul_rn((float)ib[0], sc)), v);
                d[c] = sat_u8(__float2int_rn(v));
            } else {
                d[c] = sat_u8((t[0] * ib[0] + t[1] * ib[1] + t[2] * ib[2] + t[3] * ib[3] + (1 << 21)) >> 22);
            }
        }
    } else {
        const int vec_limit = ((p.dw * CN) / 4) * 4;
#pragma unroll
        for (int c = 0; c < CN; c++) {
            float t[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float* r = rows[k];
                float v = __fmul_rn(r[xi[0] * CN + c], cx[0]);
                v = __fadd_rn(v, __fmul_rn(r[xi[1] * CN + c], cx[1]));
                v = __fadd_rn(v, __fmul_rn(r[xi[2] * CN + c], cx[2]));
                v = __fadd_rn(v, __fmul_rn(r[xi[3] * CN + c], cx[3]));
                t[k] = v;
            }
            float o;
            if (x * CN + c < vec_limit) {
                o = __fmul_rn(t[3], cy[3]);
                o = __fadd_rn(__fmul_rn(t[2], cy[2]), o);
                o = __fadd_rn(__fmul_rn(t[1], cy[1]), o);
                o = __fadd_rn(__fmul_rn(t[0], cy[0]), o);
            } else {
                o = __fmul_rn(t[0], cy[0]);
                o = __fadd_rn(o, __fmul_rn(t[1], cy[1]));
                o = __fadd_rn(o, __fmul_rn(t[2], cy[2]));
                o = __fadd_rn(o, __fmul_rn(t[3], cy[3]));
            }
            d[c] = o;
        }
    }
}

template <typename T>
static int launch_by_cn(int cn, int interp, const Img& s, const Img& d, const ResizeParams& p, cudaStream_t st)
{
    dim3 grid(div_up((unsigned)p.dw, 256), (unsigned)p.dh, (unsigned)s.frames);
#define L(K, CN) K<T, CN><<<grid, 256, 0, st>>>(s, d, p)
    if (interp == B200CV_INTER_LINEAR) {
        if (cn == 1) L(resize_linear_kernel, 1); else if (cn == 3) L(resize_linear_kernel, 3); else L(resize_linear_kernel, 4);
    } else {
        if (cn == 1) L(resize_cubic_kernel, 1); else if (cn == 3) L(resize_cubic_kernel, 3); else L(resize_cubic_kernel, 4);
    }
#undef L
    B200_LAUNCH_CHECK();
    return B200CV_OK;
}

int copy_impl(const b200cvMat* src, const b200cvMat* dst, void* stream);

}  // namespace b200cv

using namespace b200cv;

extern "C" int b200cv_resize(const b200cvMat* src, const b200cvMat* dst, int interpolation, void* stream)
{
    int rc;
    if ((rc = check_mat(src, "src")) || (rc = check_mat(dst, "dst"))) return rc;
    B200_REQUIRE(src->type == dst->type, "resize: dst type must equal src type");
    B200_REQUIRE(src->data != dst->data, "resize: in-place is not supported");
    const int depth = B200CV_DEPTH(src->type), cn = B200CV_CN(src->type);
    if ((depth != B200CV_8U && depth != B200CV_32F) || (cn != 1 && cn != 3 && cn != 4)) return B200CV_NOT_IMPLEMENTED;
    Img s = make_img(src), d = make_img(dst);
    B200_REQUIRE(s.frames == d.frames, "src/dst batch mismatch");
    if (s.rows >= 65536 || d.rows >= 65536 || s.frames >= 65536) return B200CV_NOT_IMPLEMENTED;
    cudaStream_t st = as_stream(stream);
    if (src->cols == dst->cols && src->rows == dst->rows) return copy_impl(src, dst, stream);   // resize.cpp:4238

    ResizeParams p;
    p.sw = src->cols; p.sh = src->rows; p.dw = dst->cols; p.dh = dst->rows;
    const double inv_x = (double)p.dw / p.sw, inv_y = (double)p.dh / p.sh;   // hal::resize, resize.cpp:3835-3839
    p.ifx = 1. / inv_x; p.ify = 1. / inv_y;
    p.scale_x = 1. / inv_x; p.scale_y = 1. / inv_y;
    const int pix = (int)elem_size(src->type);
    dim3 grid(div_up((unsigned)p.dw, 256), (unsigned)p.dh, (unsigned)s.frames);

    if (interpolation == B200CV_INTER_NEAREST) {
        switch (pix) {
        case 1: resize_nn_kernel<1><<<grid, 256, 0, st>>>(s, d, p); break;
        case 3: resize_nn_kernel<3><<<grid, 256, 0, st>>>(s, d, p); break;
        case 4: resize_nn_kernel<4><<<grid, 256, 0, st>>>(s, d, p); break;
        case 12: resize_nn_kernel<12><<<grid, 256, 0, st>>>(s, d, p); break;
        case 16: resize_nn_kernel<16><<<grid, 256, 0, st>>>(s, d, p); break;
        default: return B200CV_NOT_IMPLEMENTED;
        }
        B200_LAUNCH_CHECK();
        return B200CV_OK;
    }
    // exact 2x2 decimation: INTER_AREA, and INTER_LINEAR which the reference rewrites to it (resize.cpp:4009-4012)
    const int isx = (int)lrint(p.scale_x), isy = (int)lrint(p.scale_y);
    const bool area_fast = fabs(p.scale_x - isx) < 2.220446049250313e-16 && fabs(p.scale_y - isy) < 2.220446049250313e-16;
    if ((interpolation == B200CV_INTER_LINEAR || interpolation == B200CV_INTER_AREA) && area_fast && isx == 2 && isy == 2) {
        if (depth == B200CV_8U) {
            int vec_ok = (((uintptr_t)s.data | s.step | s.fstep | (uintptr_t)d.data | d.step | d.fstep) & 3) == 0;
            dim3 g4(div_up((unsigned)div_up((unsigned)p.dw, 4), 256), (unsigned)p.dh, (unsigned)s.frames);
            if (cn == 1) resize_area2_u8_kernel<1><<<g4, 256, 0, st>>>(s, d, p.dw, vec_ok);
            else if (cn == 3) resize_area2_u8_kernel<3><<<g4, 256, 0, st>>>(s, d, p.dw, vec_ok);
            else resize_area2_u8_kernel<4><<<g4, 256, 0, st>>>(s, d, p.dw, vec_ok);
        } else {
            if (cn == 1) resize_area2_f32_kernel<1><<<grid, 256, 0, st>>>(s, d, p.dw);
            else if (cn == 3) resize_area2_f32_kernel<3><<<grid, 256, 0, st>>>(s, d, p.dw);
            else resize_area2_f32_kernel<4><<<grid, 256, 0, st>>>(s, d, p.dw);
        }
        B200_LAUNCH_CHECK();
        return B200CV_OK;
    }
    if (interpolation != B200CV_INTER_LINEAR && interpolation != B200CV_INTER_CUBIC) return B200CV_NOT_IMPLEMENTED;
    return depth == B200CV_8U ? launch_by_cn<uchar>(cn, interpolation, s, d, p, st) : launch_by_cn<float>(cn, interpolation, s, d, p, st);
}
