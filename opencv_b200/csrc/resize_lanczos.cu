// resize_lanczos.cu -- cv::resize INTER_LANCZOS4 (8 x 8 taps), 8-bit and float, 1 / 3 / 4 channels.
//
// Reference (resize.cpp): s = floor(f), f = (d + 0.5) * scale - 0.5 as float; taps s-3 .. s+4, indices clamped to the image (:2083-2100 rows,
// clip() :2160 columns); weights interpolateLanczos4 (:974-1003): sin / cos of the first tap's angle in DOUBLE, the other seven through a
// 45-degree rotation table, divided by y^2, normalised in float.
//   8-bit:  weights cvRound(w * 2048) as shorts; int row sums, int column sum (wrapping like the reference's), (v + 2^21) >> 22, saturate
//   float:  rows left to right;  columns S0*b0 + (S1*b1 + ... (S6*b6 + S7*b7)) in the 4-lane SIMD body (VResizeLanczos4Vec_32f :1596-1621)
//           and left to right in the last (dw * cn) % 4 elements (:2131-2156); multiply and add rounded separately (no FMA in that unit)
// The weights need the host's libm sin / cos to be the reference's, bit for bit: the two tables (dw + dh entries of an offset and 8
// weights) are built on the host with the reference's expressions and uploaded with the call (<= 0.5 MB for 8K); everything per pixel
// runs on the device.  Float: one thread per destination element, 64 taps (gather / issue bound, like CUBIC); 8-bit: the tiled separable
// kernel below (B200CV_RESIZE_LANCZOS_PATH=v1 keeps the per-element kernel).
#include <math.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "common.cuh"
#include "resize.cuh"

namespace b200cv {

namespace {


// interpolateLanczos4, resize.cpp:974-1003 (host; this file is compiled with -ffp-contract=off)
void lanczos4_weights(float x, float* c)
{
    static const double s45 = 0.70710678118654752440084436210485;
    static const double cs[][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
    float sum = 0;
    const double y0 = -(x + 3) * 3.1415926535897932384626433832795 * 0.25, s0 = sin(y0), c0 = cos(y0);
    for (int i = 0; i < 8; i++) {
        const float y0_ = (x + 3 - i);
        if (fabsf(y0_) >= 1e-6f) {
            const double y = -y0_ * 3.1415926535897932384626433832795 * 0.25;
            c[i] = (float)((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
        } else {
            c[i] = 1e30f;
        }
        sum += c[i];
    }
    sum = 1.f / sum;
    for (int i = 0; i < 8; i++) c[i] *= sum;
}

void lanczos4_table(int dn, double scale, LzTap* tab)
{
    for (int d = 0; d < dn; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        const int s = (int)floorf(f);
        f -= s;
        LzTap& t = tab[d];
        t.s = s; t.pad[0] = t.pad[1] = 0;
        lanczos4_weights(f, t.fc);
        for (int k = 0; k < 8; k++) {                       // saturate_cast<short>(w * INTER_RESIZE_COEF_SCALE), resize.cpp:4141
            const long r = lrintf(t.fc[k] * 2048.f);
            t.ic[k] = (short)(r < -32768 ? -32768 : r > 32767 ? 32767 : r);
        }
    }
}


template <typename T, int CN>
__global__ void __launch_bounds__(256) resize_lanczos4_kernel(Img src, Img dst, const LzTap* __restrict__ xt, const LzTap* __restrict__ yt, int sw, int sh, int dw)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;           // destination element x * CN + c
    const int y = blockIdx.y, f = blockIdx.z;
    if (e >= dw * CN) return;
    const int x = e / CN, c = e - x * CN;
    const LzTap tx = xt[x], ty = yt[y];
    int xi[8];
#pragma unroll
    for (int j = 0; j < 8; j++) xi[j] = lz_clip(tx.s - 3 + j, sw) * CN + c;
    if constexpr (sizeof(T) == 1) {
        unsigned v = 0;                                           // unsigned: the reference's int sums wrap
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uchar* r = src.row<uchar>(f, lz_clip(ty.s - 3 + k, sh));
            int t = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) t += r[xi[j]] * tx.ic[j];
            v += (unsigned)t * (unsigned)(int)ty.ic[k];
        }
        dst.row<uchar>(f, y)[e] = sat_u8((int)(v + (1u << 21)) >> 22);
    } else {
        float t[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float* r = src.row<float>(f, lz_clip(ty.s - 3 + k, sh));
            float v = __fmul_rn(r[xi[0]], tx.fc[0]);
#pragma unroll
            for (int j = 1; j < 8; j++) v = __fadd_rn(v, __fmul_rn(r[xi[j]], tx.fc[j]));
            t[k] = v;
        }
        float o;
        if (e < ((dw * CN) & ~3)) {
            o = __fmul_rn(t[7], ty.fc[7]);
#pragma unroll
            for (int k = 6; k >= 0; k--) o = __fadd_rn(__fmul_rn(t[k], ty.fc[k]), o);
        } else {
            o = __fmul_rn(t[0], ty.fc[0]);
#pragma unroll
            for (int k = 1; k < 8; k++) o = __fadd_rn(o, __fmul_rn(t[k], ty.fc[k]));
        }
        dst.row<float>(f, y)[e] = o;
    }
}

}  // namespace

// called by b200cv_resize (types, channel counts and batch sizes already checked)
int resize_lanczos_impl(const Img& s, const Img& d, int depth, int cn, cudaStream_t st)
{
    const int sw = s.cols, sh = s.rows, dw = d.cols, dh = d.rows;
    if (dh >= 65536 || s.frames >= 65536) return B200CV_NOT_IMPLEMENTED;
    const double inv_x = (double)dw / sw, inv_y = (double)dh / sh;               // hal::resize, resize.cpp:3835-3839
    std::vector<LzTap> tab((size_t)dw + dh);
    lanczos4_table(dw, 1. / inv_x, tab.data());
    lanczos4_table(dh, 1. / inv_y, tab.data() + dw);
    LzTap* dtab = nullptr;
    B200_CUDA(cudaMallocAsync((void**)&dtab, sizeof(LzTap) * tab.size(), st));
    cudaError_t ce = cudaMemcpyAsync(dtab, tab.data(), sizeof(LzTap) * tab.size(), cudaMemcpyHostToDevice, st);   // pageable source: staged before the call returns
    if (ce != cudaSuccess) { cudaFreeAsync(dtab, st); return cuda_fail(ce, "cudaMemcpyAsync(lanczos tables)", __FILE__, __LINE__); }
    const LzTap *xt = dtab, *yt = dtab + dw;
    const dim3 block(256);
    const dim3 grid(div_up((unsigned)(dw * cn), 256), (unsigned)dh, (unsigned)s.frames);
    bool done = false;
#ifndef B200CV_HOST_EMULATION
    const char* lz_env = getenv("B200CV_RESIZE_LANCZOS_PATH");                 // "v1": the per-element kernel for 8-bit images too
    if (depth == B200CV_8U && !(lz_env && !strcmp(lz_env, "v1")))
        done = resize_lanczos_sep_u8(s, d, cn, tab.data() + dw, xt, yt, st);    // resize_lanczos_sep.cu: tiled, separable
#endif
    if (done) {
    } else if (depth == B200CV_8U) {
        if (cn == 1) resize_lanczos4_kernel<uchar, 1><<<grid, block, 0, st>>>(s, d, xt, yt, sw, sh, dw);
        else if (cn == 3) resize_lanczos4_kernel<uchar, 3><<<grid, block, 0, st>>>(s, d, xt, yt, sw, sh, dw);
        else resize_lanczos4_kernel<uchar, 4><<<grid, block, 0, st>>>(s, d, xt, yt, sw, sh, dw);
    } else {
        if (cn == 1) resize_lanczos4_kernel<float, 1><<<grid, block, 0, st>>>(s, d, xt, yt, sw, sh, dw);
        else if (cn == 3) resize_lanczos4_kernel<float, 3><<<grid, block, 0, st>>>(s, d, xt, yt, sw, sh, dw);
        else resize_lanczos4_kernel<float, 4><<<grid, block, 0, st>>>(s, d, xt, yt, sw, sh, dw);
    }
    ce = cudaGetLastError();
    count_launch();
    cudaFreeAsync(dtab, st);
    if (ce != cudaSuccess) return cuda_fail(ce, "kernel launch", __FILE__, __LINE__);
    return B200CV_OK;
}

}  // namespace b200cv
