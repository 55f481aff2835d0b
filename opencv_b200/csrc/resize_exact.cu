// resize_exact.cu -- cv::resize INTER_NEAREST_EXACT and INTER_LINEAR_EXACT (8-bit): the reference's integer-only resizers.
//
// NEAREST_EXACT (resizeNN_bitexact, resize.cpp:1267-1288): source index of the destination pixel CENTRE in 16.16 fixed point,
//     ifx = ((sw << 16) + dw/2) / dw,  ifx0 = ifx/2 - sw % 2,  sx = min((ifx * x + ifx0) >> 16, sw - 1)       (any pixel size)
// LINEAR_EXACT, 8-bit (resize_bitExact<uint8_t, interpolationLinear>, resize.cpp:776-960; fixedpoint.inl.hpp:326-374):
//     f = (1 / inv_scale) * (d + 0.5) - 0.5 in double;  i = floor f;  weights in 8.8 fixed point c1 = cvRound((f - i) * 256), c0 = 256 - c1;
//     i < 0 (or a 1-pixel source): copy the first sample; i >= size - 1: copy the last;
//     row pass   H = c0 * p[i] + c1 * p[i+1]            (16-bit, at most 255 * 256)
//     column pass (H0 * b0 + H1 * b1 + 2^15) >> 16, or (H + 128) >> 8 for the copied rows;  all unsigned integer, bit-exact.
//     (float data has no exact mode -- cv::resize turns it into INTER_LINEAR, resize.cpp:4223; exact 2 x 2 decimation is the
//      INTER_AREA fast path, :3976-3981: both are routed by b200cv_resize before this file is reached.)
// One thread per destination element; the weights are derived in the thread from d and the scale with the reference's expressions
// (explicit _rn intrinsics: no contraction), so there are no tables.  Gather-bound like INTER_LINEAR.
#include "common.cuh"

namespace b200cv {

namespace {

struct ExactParams {
    int sw, sh, dw, dh;
    int ifx, ifx0, ify, ify0;   // NEAREST_EXACT
    double scale_x, scale_y;    // LINEAR_EXACT: 1 / inv_scale
};

template <int PIX>
__global__ void __launch_bounds__(256) resize_nn_exact_kernel(Img src, Img dst, ExactParams p)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, f = blockIdx.z;
    if (x >= p.dw) return;
    const int sy = min((p.ify * y + p.ify0) >> 16, p.sh - 1);
    const int sx = min((p.ifx * x + p.ifx0) >> 16, p.sw - 1);
    const uchar* s = src.row<uchar>(f, sy) + (size_t)sx * PIX;
    uchar* d = dst.row<uchar>(f, y) + (size_t)x * PIX;
#pragma unroll
    for (int i = 0; i < PIX; i++) d[i] = s[i];
}

// position d along one axis: kind 0 = interpolate between i and i + 1 with weights (256 - c1, c1), 1 = copy the first sample, 2 = copy the last
struct ExactTap { int kind, i; unsigned c1; };

__device__ __forceinline__ ExactTap exact_tap(int d, double scale, int ssize)
{
    ExactTap t;
    const double fval = __dsub_rn(__dmul_rn(scale, (double)d + 0.5), 0.5);     // d + 0.5 is exact
    const int ival = (int)floor(fval);
    t.i = 0; t.c1 = 0;
    if (ival >= 0 && ssize > 1) {
        if (ival < ssize - 1) {
            t.kind = 0; t.i = ival;
            t.c1 = (unsigned)__double2int_rn(__dmul_rn(__dsub_rn(fval, (double)ival), 256.0));
        } else t.kind = 2;
    } else t.kind = 1;
    return t;
}

template <int CN>
__device__ __forceinline__ unsigned exact_hline(const uchar* s, int c, const ExactTap& tx, int sw)
{
    if (tx.kind == 1) return (unsigned)s[c] << 8;
    if (tx.kind == 2) return (unsigned)s[(sw - 1) * CN + c] << 8;
    const unsigned a = min((256u - tx.c1) * s[tx.i * CN + c], 65535u);
    const unsigned b = tx.c1 ? min(tx.c1 * s[(tx.i + 1) * CN + c], 65535u) : 0u;
    return min(a + b, 65535u);
}

template <int CN>
__global__ void __launch_bounds__(256) resize_linear_exact_kernel(Img src, Img dst, ExactParams p)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;           // destination element x * CN + c
    const int y = blockIdx.y, f = blockIdx.z;
    if (e >= p.dw * CN) return;
    const int x = e / CN, c = e - x * CN;
    const ExactTap tx = exact_tap(x, p.scale_x, p.sw), ty = exact_tap(y, p.scale_y, p.sh);
    unsigned v;
    if (ty.kind == 0) {
        const unsigned h0 = exact_hline<CN>(src.row<uchar>(f, ty.i), c, tx, p.sw);
        const unsigned h1 = exact_hline<CN>(src.row<uchar>(f, ty.i + 1), c, tx, p.sw);
        v = (h0 * (256u - ty.c1) + h1 * ty.c1 + 32768u) >> 16;
    } else {
        const unsigned h = exact_hline<CN>(src.row<uchar>(f, ty.kind == 1 ? 0 : p.sh - 1), c, tx, p.sw);
        v = ((h + 128u) & 0xFFFFu) >> 8;                           // fixedround() is a 16-bit add
    }
    dst.row<uchar>(f, y)[e] = (uchar)min(v, 255u);
}

}  // namespace

// called by b200cv_resize (types, channel counts and batch sizes already checked); exact: 5 = INTER_LINEAR_EXACT (8-bit), 6 = INTER_NEAREST_EXACT
int resize_exact_impl(const Img& s, const Img& d, int depth, int cn, int interpolation, cudaStream_t st)
{
    ExactParams p;
    p.sw = s.cols; p.sh = s.rows; p.dw = d.cols; p.dh = d.rows;
    if (d.rows >= 65536 || s.frames >= 65536) return B200CV_NOT_IMPLEMENTED;
    const dim3 block(256);
    if (interpolation == 6) {
        if (p.sw >= 32768 || p.sh >= 32768) return B200CV_NOT_IMPLEMENTED;       // the reference's 16.16 products overflow int beyond that
        p.ifx = ((p.sw << 16) + p.dw / 2) / p.dw; p.ifx0 = p.ifx / 2 - p.sw % 2;
        p.ify = ((p.sh << 16) + p.dh / 2) / p.dh; p.ify0 = p.ify / 2 - p.sh % 2;
        const int pix = cn * (depth == B200CV_8U ? 1 : 4);
        const dim3 grid(div_up((unsigned)p.dw, 256), (unsigned)p.dh, (unsigned)s.frames);
        switch (pix) {
        case 1: resize_nn_exact_kernel<1><<<grid, block, 0, st>>>(s, d, p); break;
        case 3: resize_nn_exact_kernel<3><<<grid, block, 0, st>>>(s, d, p); break;
        case 4: resize_nn_exact_kernel<4><<<grid, block, 0, st>>>(s, d, p); break;
        case 12: resize_nn_exact_kernel<12><<<grid, block, 0, st>>>(s, d, p); break;
        case 16: resize_nn_exact_kernel<16><<<grid, block, 0, st>>>(s, d, p); break;
        default: return B200CV_NOT_IMPLEMENTED;
        }
        B200_LAUNCH_CHECK();
        return B200CV_OK;
    }
    if (interpolation != 5 || depth != B200CV_8U) return B200CV_NOT_IMPLEMENTED;
    const double inv_x = (double)p.dw / p.sw, inv_y = (double)p.dh / p.sh;      // hal::resize, resize.cpp:3835-3839
    p.scale_x = 1. / inv_x; p.scale_y = 1. / inv_y;
    const dim3 grid(div_up((unsigned)(p.dw * cn), 256), (unsigned)p.dh, (unsigned)s.frames);
    if (cn == 1) resize_linear_exact_kernel<1><<<grid, block, 0, st>>>(s, d, p);
    else if (cn == 3) resize_linear_exact_kernel<3><<<grid, block, 0, st>>>(s, d, p);
    else resize_linear_exact_kernel<4><<<grid, block, 0, st>>>(s, d, p);
    B200_LAUNCH_CHECK();
    return B200CV_OK;
}

}  // namespace b200cv
