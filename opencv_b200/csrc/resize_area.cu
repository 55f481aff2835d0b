// resize_area.cu -- cv::resize INTER_AREA in its true area mode (both scale factors >= 1; resize.cpp:4016-4064).
//
//   integer scale factors (resizeAreaFast_Invoker, resize.cpp:2969-3060): plain sx x sy window sum times float(1 / area);
//       8-bit: int sum, cvRound(float(sum) * scale);  float: sums in groups of four (((a+b)+c)+d, CV_ENABLE_UNROLLED), then sum * scale.
//       (2 x 2 with 1 / 3 / 4 channels is the (a+b+c+d+2)>>2 path of resize.cu.)
//   any other factors (computeResizeAreaTab :3334-3373, ResizeArea_Invoker :3183-3297): every destination cell [dx*scale, (dx+1)*scale)
//       is covered by an optional partial first source pixel, whole pixels, an optional partial last pixel; weights are double
//       quotients rounded to float.  Per source row  buf = ((0 + S0*a0) + S1*a1) + ...;  per destination row  sum = b0*buf0, then
//       sum += bj*bufj  -- float, multiply and add rounded separately (the translation unit is built without FMA).  cvRound for 8-bit.
// Both are bit-exact: the kernels issue the same float / double operations in the same order (explicit _rn intrinsics: nvcc would
// otherwise contract a*b+c).  One thread per destination element; the taps of a cell are derived in the thread from dx and the scale
// in double arithmetic (the same expressions as the reference's table builder), so there is no table, no allocation and no upload.
// Scales above 1 are decimations: every source byte is read once, the bound is HBM.
#include "common.cuh"

namespace b200cv {

namespace {

struct AreaParams {
    int sw, sh, dw, dh;
    int isx, isy;               // integer-scale kernel
    float inv_area;
    double scale_x, scale_y;    // general kernel
};

template <typename T> __device__ __forceinline__ T area_store(float v);
template <> __device__ __forceinline__ uchar area_store<uchar>(float v) { return sat_u8(__float2int_rn(v)); }
template <> __device__ __forceinline__ float area_store<float>(float v) { return v; }

template <typename T, int CN>
__global__ void __launch_bounds__(256) resize_area_int_kernel(Img src, Img dst, AreaParams p)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;           // destination element x * CN + c
    const int y = blockIdx.y, f = blockIdx.z;
    if (e >= p.dw * CN) return;
    const int x = e / CN, c = e - x * CN;
    const int area = p.isx * p.isy;
    if constexpr (sizeof(T) == 1) {
        int sum = 0;
        for (int j = 0; j < p.isy; j++) {
            const T* s = src.row<T>(f, y * p.isy + j) + (x * p.isx) * CN + c;
            for (int i = 0; i < p.isx; i++) sum += s[i * CN];
        }
        dst.row<T>(f, y)[e] = area_store<T>(__fmul_rn(__int2float_rn(sum), p.inv_area));
    } else {
        float sum = 0.f, v0 = 0.f, v1 = 0.f, v2 = 0.f;
        int k = 0, n = 0;                                         // k: elements consumed by complete groups / singles, n: fill of the open group
        for (int j = 0; j < p.isy; j++) {
            const T* s = src.row<T>(f, y * p.isy + j) + (x * p.isx) * CN + c;
            for (int i = 0; i < p.isx; i++) {
                const float sv = s[i * CN];
                if (k <= area - 4 || n) {
                    if (n == 0) { v0 = sv; n = 1; }
                    else if (n == 1) { v1 = sv; n = 2; }
                    else if (n == 2) { v2 = sv; n = 3; }
                    else { sum = __fadd_rn(sum, __fadd_rn(__fadd_rn(__fadd_rn(v0, v1), v2), sv)); n = 0; k += 4; }
                } else {
                    sum = __fadd_rn(sum, sv);
                    k++;
                }
            }
        }
        dst.row<T>(f, y)[e] = area_store<T>(__fmul_rn(sum, p.inv_area));
    }
}

// the taps of destination cell d along one axis: [first partial] [whole pixels s1 .. s2-1] [last partial] (computeResizeAreaTab)
struct AreaTaps {
    int s1, s2;                 // whole pixels
    bool has_first, has_last;   // partial pixel s1 - 1 / s2
    float a_first, a_mid, a_last;
};

__device__ __forceinline__ AreaTaps area_taps(int d, double scale, int ssize)
{
    AreaTaps t;
    const double fs1 = __dmul_rn((double)d, scale);
    const double fs2 = __dadd_rn(fs1, scale);
    const double rest = __dsub_rn((double)ssize, fs1);
    const double cell = scale < rest ? scale : rest;               // std::min(scale, ssize - fsx1)
    int s1 = (int)ceil(fs1), s2 = (int)floor(fs2);
    s2 = min(s2, ssize - 1);
    s1 = min(s1, s2);
    t.s1 = s1; t.s2 = s2;
    const double d1 = __dsub_rn((double)s1, fs1), d2 = __dsub_rn(fs2, (double)s2);
    t.has_first = d1 > 1e-3;
    t.has_last = d2 > 1e-3;
    t.a_first = __double2float_rn(__ddiv_rn(d1, cell));
    t.a_mid = __double2float_rn(__ddiv_rn(1.0, cell));
    double m = d2 < 1. ? d2 : 1.;                                  // std::min(std::min(fsx2 - sx2, 1.), cellWidth)
    m = m < cell ? m : cell;
    t.a_last = __double2float_rn(__ddiv_rn(m, cell));
    return t;
}

template <typename T, int CN>
__device__ __forceinline__ float area_row(const T* s, int c, const AreaTaps& tx)
{
    float buf = 0.f;
    if (tx.has_first) buf = __fadd_rn(buf, __fmul_rn((float)s[(tx.s1 - 1) * CN + c], tx.a_first));
    for (int sx = tx.s1; sx < tx.s2; sx++) buf = __fadd_rn(buf, __fmul_rn((float)s[sx * CN + c], tx.a_mid));
    if (tx.has_last) buf = __fadd_rn(buf, __fmul_rn((float)s[tx.s2 * CN + c], tx.a_last));
    return buf;
}

template <typename T, int CN>
__global__ void __launch_bounds__(256) resize_area_kernel(Img src, Img dst, AreaParams p)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, f = blockIdx.z;
    if (e >= p.dw * CN) return;
    const int x = e / CN, c = e - x * CN;
    const AreaTaps tx = area_taps(x, p.scale_x, p.sw), ty = area_taps(y, p.scale_y, p.sh);
    float sum = 0.f;
    bool first = true;                                            // first source row of this destination row: sum = beta * buf
    if (ty.has_first) {
        sum = __fmul_rn(ty.a_first, area_row<T, CN>(src.row<T>(f, ty.s1 - 1), c, tx));
        first = false;
    }
    for (int sy = ty.s1; sy < ty.s2; sy++) {
        const float t = __fmul_rn(ty.a_mid, area_row<T, CN>(src.row<T>(f, sy), c, tx));
        sum = first ? t : __fadd_rn(sum, t);
        first = false;
    }
    if (ty.has_last) {
        const float t = __fmul_rn(ty.a_last, area_row<T, CN>(src.row<T>(f, ty.s2), c, tx));
        sum = first ? t : __fadd_rn(sum, t);
    }
    dst.row<T>(f, y)[e] = area_store<T>(sum);
}

template <typename T>
int launch_area(bool integer, int cn, const Img& s, const Img& d, const AreaParams& p, cudaStream_t st)
{
    const dim3 block(256);
    const dim3 grid(div_up((unsigned)(p.dw * cn), 256), (unsigned)p.dh, (unsigned)s.frames);
    if (integer) {
        if (cn == 1) resize_area_int_kernel<T, 1><<<grid, block, 0, st>>>(s, d, p);
        else if (cn == 3) resize_area_int_kernel<T, 3><<<grid, block, 0, st>>>(s, d, p);
        else resize_area_int_kernel<T, 4><<<grid, block, 0, st>>>(s, d, p);
    } else {
        if (cn == 1) resize_area_kernel<T, 1><<<grid, block, 0, st>>>(s, d, p);
        else if (cn == 3) resize_area_kernel<T, 3><<<grid, block, 0, st>>>(s, d, p);
        else resize_area_kernel<T, 4><<<grid, block, 0, st>>>(s, d, p);
    }
    B200_LAUNCH_CHECK();
    return B200CV_OK;
}

}  // namespace

// called by b200cv_resize for INTER_AREA with both scale factors >= 1 (types, channel counts and batch sizes already checked)
int resize_area_impl(const Img& s, const Img& d, int depth, int cn, cudaStream_t st)
{
    AreaParams p;
    p.sw = s.cols; p.sh = s.rows; p.dw = d.cols; p.dh = d.rows;
    const double inv_x = (double)p.dw / p.sw, inv_y = (double)p.dh / p.sh;      // hal::resize, resize.cpp:3835-3839
    p.scale_x = 1. / inv_x; p.scale_y = 1. / inv_y;
    if (p.scale_x < 1 || p.scale_y < 1 || d.rows >= 65536 || s.frames >= 65536) return B200CV_NOT_IMPLEMENTED;
    p.isx = (int)nearbyint(p.scale_x); p.isy = (int)nearbyint(p.scale_y);
    const bool integer = fabs(p.scale_x - p.isx) < 2.220446049250313e-16 && fabs(p.scale_y - p.isy) < 2.220446049250313e-16;
    p.inv_area = 1.f / (p.isx * p.isy);
    return depth == B200CV_8U ? launch_area<uchar>(integer, cn, s, d, p, st) : launch_area<float>(integer, cn, s, d, p, st);
}

}  // namespace b200cv
