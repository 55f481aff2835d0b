// resize.cuh -- types shared by resize.cu (tables, per-pixel kernels) and resize_sep.cu (tiled separable 8-bit kernels)
#pragma once
#include "common.cuh"

namespace b200cv {

struct ResizeParams {
    double ifx, ify;          // NEAREST: 1/fx, 1/fy
    double scale_x, scale_y;  // LINEAR/CUBIC: 1/inv_scale
    double inv_x, inv_y;      // INTER_AREA on an enlarging axis: weights from (d + 1) - (s + 1) * inv_scale
    int sw, sh, dw, dh;
    int area_mode;
};

// The reference tabulates per destination column / row the source index and the taps once per call on the host
// (resize.cpp:4097-4190).  Same here, on the device (resize_tab_kernel): the main kernels only read the tables.
struct ResTab {            // 32 bytes
    int s;                 // source index (unclamped for rows / cubic columns, clamped for linear columns)
    int last;              // linear columns: taps collapse to S[s]*ONE (dx >= xmax)
    int pad[2];
    union { int ic[4]; float fc[4]; };
};

// INTER_LANCZOS4 tables (resize_lanczos.cu builds them on the host; resize_lanczos_sep.cu reads them too)
struct LzTap {
    int s;                 // source index of tap 3 (floor of the source coordinate)
    float fc[8];
    short ic[8];
    short pad[2];
};
__host__ __device__ __forceinline__ int lz_clip(int x, int n) { return x < 0 ? 0 : (x < n ? x : n - 1); }
bool resize_lanczos_sep_u8(const Img& s, const Img& d, int cn, const LzTap* host_yt, const LzTap* xt, const LzTap* yt, cudaStream_t st);   // resize_lanczos_sep.cu

__host__ __device__ __forceinline__ int clip_i(int x, int a, int b) { return x >= a ? (x < b ? x : b - 1) : a; }

// ---- coefficient helpers -------------------------------------------------------------------------------------------
// linear: returns source index and fractional weight with the reference's edge clamps (ksize2 == 1)
// area_mode (INTER_AREA when an axis is enlarged, resize.cpp:4104-4109): s = floor(d * scale), f = (d+1) - (s+1) * inv_scale, <= 0 -> 0, else its fraction
__device__ __forceinline__ void linear_coef(int d, double scale, int ssize, int& s, float& fr, bool clamp_edges, bool area_mode = false, double inv_scale = 0.)
{
    float fx;
    int sx;
    if (!area_mode) {
        fx = (float)__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5);
        sx = (int)floorf(fx);
        fx = __fsub_rn(fx, (float)sx);
    } else {
        sx = (int)floor(__dmul_rn((double)d, scale));
        fx = (float)__dsub_rn((double)(d + 1), __dmul_rn((double)(sx + 1), inv_scale));
        fx = fx <= 0.f ? 0.f : __fsub_rn(fx, floorf(fx));
    }
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


// one table entry: destination column (is_y = false) or row d
template <bool CUBIC, bool FIXPT>
__device__ __forceinline__ ResTab res_tab_entry(int d, bool is_y, const ResizeParams& p)
{
    int s; float fr;
    linear_coef(d, is_y ? p.scale_y : p.scale_x, is_y ? p.sh : p.sw, s, fr, !CUBIC && !is_y, !CUBIC && p.area_mode, is_y ? p.inv_y : p.inv_x);
    ResTab t;
    t.s = s; t.last = (!CUBIC && !is_y && s >= p.sw - 1); t.pad[0] = t.pad[1] = 0;
    float c[4];
    if (CUBIC) cubic_coeffs(fr, c);
    else { c[0] = __fsub_rn(1.f, fr); c[1] = fr; c[2] = c[3] = 0.f; }
#pragma unroll
    for (int k = 0; k < 4; k++) { if (FIXPT) t.ic[k] = coef_s16(c[k]); else t.fc[k] = c[k]; }
    return t;
}

// 8-bit LINEAR / CUBIC through a shared-memory tile of horizontally filtered rows (resize_sep.cu); NOT_IMPLEMENTED when the
// configuration does not fit (the caller then runs the per-pixel kernels)
int resize_sep_u8(const Img& s, const Img& d, int cn, bool cubic, const ResizeParams& p, const ResTab* xt, const ResTab* yt, cudaStream_t st);

}  // namespace b200cv
