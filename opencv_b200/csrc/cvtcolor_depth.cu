// cvtcolor_depth.cu -- cv::cvtColor for CV_16U and CV_32F images: channel reorders, BGR / RGB (A) <-> GRAY, BGR / RGB <-> YCrCb / YUV.
//
// Reference arithmetic (modules/imgproc/src/color_rgb.simd.hpp, color_yuv.simd.hpp), pinned against the compiled reference on the CPU
// (every formula below reproduced it with 0 differing elements before this file was written):
//   RGB2RGB<T>            channel copy, alpha = ColorChannel<T>::max() (65535 / 1.0f)                                        (:24-120)
//   RGB2Gray<ushort>      (c0*cb + c1*cg + c2*cr + 2^14) >> 15, unsigned, coefficients 3735 / 19235 / 9798                     (:752-841)
//   RGB2Gray<float>       vector body  fma(c2, k2, fma(c1, k1, c0*k0));  the last width % 8 pixels of a row run the scalar
//                         statement c0*k0 + c1*k1 + c2*k2, which GCC contracts to fma(c2, k2, fma(c0, k0, c1*k1))             (:608-657)
//   Gray2RGB<T>           three copies + alpha                                                                               (:851-937)
//   RGB2YCrCb_f<float>    Y: vector fma(c0, C0, fma(c1, C1, c2*C2)), scalar tail fma(c2, C2, fma(c0, C0, c1*C1));
//                         Cr = fma(R - Y, C3, 0.5), Cb = fma(B - Y, C4, 0.5)                                                   (color_yuv.simd.hpp:134-211)
//   RGB2YCrCb_i<ushort>   Y = DESCALE(c0*C0 + c1*C1 + c2*C2, 14), Cr = DESCALE((R - Y)*C3 + 2^29, 14), saturate_cast<ushort>   (:214-396)
//   YCrCb2RGB_f<float>    b = fma(Cb - .5, C3, Y), g = fma(Cr - .5, C1, fma(Cb - .5, C2, Y)), r = fma(Cr - .5, C0, Y)          (:616-689)
//   YCrCb2RGB_i<ushort>   b = Y + DESCALE((Cb - 32768)*C3, 14) ..., saturate_cast<ushort>                                      (:692-735, :880-1013)
//   RGB2HSV_f / HSV2RGB_f  float only; see the operators below                                                                  (color_hsv.simd.hpp:269-515)
//   RGB2XYZ_f / XYZ2RGB_f, RGB2XYZ_i / XYZ2RGB_i<ushort>   3x3 matrix; float: 4-lane vectors c0*C0 + (c1*C1 + c2*C2), tail (c0*C0 + c1*C1) + c2*C2, no FMA;
//                         16-bit: DESCALE(.., 12), saturate_cast<ushort>                                                        (color_lab.cpp:172-700)
// Pure streaming: a thread converts 4 adjacent pixels (8-byte vector accesses when the row allows), nothing is staged.
#include "common.cuh"

namespace b200cv {

namespace {

template <typename T> struct ChanMax;
template <> struct ChanMax<unsigned short> { static __device__ __forceinline__ unsigned short v() { return 65535; } };
template <> struct ChanMax<float> { static __device__ __forceinline__ float v() { return 1.0f; } };

__device__ __forceinline__ unsigned short sat_u16(int v) { return (unsigned short)min(max(v, 0), 65535); }
__device__ __forceinline__ int descale14(int v) { return (v + (1 << 13)) >> 14; }

template <typename T, int SCN, int DCN, bool SWAP> struct DOpReorder {
    __device__ __forceinline__ void operator()(const T* s, T* d, bool) const
    {
        d[0] = s[SWAP ? 2 : 0]; d[1] = s[1]; d[2] = s[SWAP ? 0 : 2];
        if (DCN == 4) d[3] = SCN == 4 ? s[3] : ChanMax<T>::v();
    }
};
template <typename T, int DCN> struct DOpGray2BGR {
    __device__ __forceinline__ void operator()(const T* s, T* d, bool) const
    {
        d[0] = d[1] = d[2] = s[0];
        if (DCN == 4) d[3] = ChanMax<T>::v();
    }
};
struct DOpGray16 {      // k0 applies to channel 0
    unsigned k0, k1, k2;
    __device__ __forceinline__ void operator()(const unsigned short* s, unsigned short* d, bool) const { d[0] = (unsigned short)((s[0] * k0 + s[1] * k1 + s[2] * k2 + (1u << 14)) >> 15); }
};
struct DOpGray32 {
    float k0, k1, k2;
    __device__ __forceinline__ void operator()(const float* s, float* d, bool vec) const
    {
        d[0] = vec ? fmaf(s[2], k2, fmaf(s[1], k1, __fmul_rn(s[0], k0))) : fmaf(s[2], k2, fmaf(s[0], k0, __fmul_rn(s[1], k1)));
    }
};
template <int BIDX, int YUV> struct DOpToYCC16 {     // destination order Y Cr Cb (YUV = 0) or Y U V = Y Cb Cr (YUV = 1)
    int c0, c1, c2, c3, c4;
    __device__ __forceinline__ void operator()(const unsigned short* s, unsigned short* d, bool) const
    {
        const int Y = descale14(s[0] * c0 + s[1] * c1 + s[2] * c2);
        const int Cr = descale14((s[BIDX ^ 2] - Y) * c3 + (32768 << 14)), Cb = descale14((s[BIDX] - Y) * c4 + (32768 << 14));
        d[0] = sat_u16(Y); d[1 + YUV] = sat_u16(Cr); d[2 - YUV] = sat_u16(Cb);
    }
};
template <int BIDX, int YUV> struct DOpToYCC32 {
    float c0, c1, c2, c3, c4;
    __device__ __forceinline__ void operator()(const float* s, float* d, bool vec) const
    {
        const float Y = vec ? fmaf(s[0], c0, fmaf(s[1], c1, __fmul_rn(s[2], c2))) : fmaf(s[2], c2, fmaf(s[0], c0, __fmul_rn(s[1], c1)));
        d[0] = Y;
        d[1 + YUV] = fmaf(__fsub_rn(s[BIDX ^ 2], Y), c3, 0.5f);
        d[2 - YUV] = fmaf(__fsub_rn(s[BIDX], Y), c4, 0.5f);
    }
};
template <int BIDX, int YUV, int DCN> struct DOpFromYCC16 {
    int c0, c1, c2, c3;
    __device__ __forceinline__ void operator()(const unsigned short* s, unsigned short* d, bool) const
    {
        const int Y = s[0], Cr = s[1 + YUV] - 32768, Cb = s[2 - YUV] - 32768;
        d[BIDX] = sat_u16(Y + descale14(Cb * c3));
        d[1] = sat_u16(Y + descale14(Cb * c2 + Cr * c1));
        d[BIDX ^ 2] = sat_u16(Y + descale14(Cr * c0));
        if (DCN == 4) d[3] = 65535;
    }
};
template <int BIDX, int YUV, int DCN> struct DOpFromYCC32 {
    float c0, c1, c2, c3;
    __device__ __forceinline__ void operator()(const float* s, float* d, bool) const
    {
        const float Y = s[0], cr = __fsub_rn(s[1 + YUV], 0.5f), cb = __fsub_rn(s[2 - YUV], 0.5f);
        d[BIDX] = fmaf(cb, c3, Y);
        d[1] = fmaf(cr, c1, fmaf(cb, c2, Y));
        d[BIDX ^ 2] = fmaf(cr, c0, Y);
        if (DCN == 4) d[3] = 1.0f;
    }
};

// float HSV (color_hsv.simd.hpp:269-368 RGB2HSV_f, :371-515 HSV2RGB_f; hue range 360 for float images, the _FULL codes are the same).
// Vector body (8 lanes): s = diff / (|max| + eps), h = fma(hsel, 60 / (diff + eps), res) * hscale with res in {0, 360, 120, 240};
// scalar tail: 60. / (diff + eps) in DOUBLE, rounded to float; h = (g - b) * d, or fma(b - r, d, 120), fma(r - g, d, 240) (GCC's contraction), + 360 if negative.
template <int BIDX> struct DOpToHSV32 {
    float hscale;
    __device__ __forceinline__ void operator()(const float* sp, float* d, bool vec) const
    {
        const float b = sp[BIDX], g = sp[1], r = sp[BIDX ^ 2];
        const float vmax = fmaxf(fmaxf(r, g), b), vmin = fminf(fminf(r, g), b);
        const float diff = __fsub_rn(vmax, vmin);
        const float s = __fdiv_rn(diff, __fadd_rn(fabsf(vmax), 1.1920929e-07f));
        const bool req = r == vmax, geq = g == vmax;
        float h;
        if (vec) {
            const float hsel = req ? __fsub_rn(g, b) : geq ? __fsub_rn(b, r) : __fsub_rn(r, g);
            const float res = req ? (g < b ? 360.f : 0.f) : geq ? 120.f : 240.f;
            h = fmaf(hsel, __fdiv_rn(60.f, __fadd_rn(diff, 1.1920929e-07f)), res);
        } else {
            const float dd = __double2float_rn(__ddiv_rn(60.0, (double)__fadd_rn(diff, 1.1920929e-07f)));
            h = req ? __fmul_rn(__fsub_rn(g, b), dd) : geq ? fmaf(__fsub_rn(b, r), dd, 120.f) : fmaf(__fsub_rn(r, g), dd, 240.f);
            if (h < 0.f) h = __fadd_rn(h, 360.f);
        }
        d[0] = __fmul_rn(h, hscale); d[1] = s; d[2] = vmax;
    }
};
// inverse: tab1 = v(1 - s), tab2 = v * fma(-s, h, 1), tab3 = v * fma(-s, 1 - h, 1) (the compiler's contraction in both the vector unit and the tail);
// vector body: sector from TRUNCATED h through float arithmetic and a chain of selects (negative hues fall through it as written);
// tail (HSV2RGB_native): s == 0 -> grey, FLOORED h, sector mod 6 made non-negative, table look-up
template <int BIDX, int DCN> struct DOpFromHSV32 {
    float hscale;
    __device__ __forceinline__ void operator()(const float* sp, float* d, bool vec) const
    {
        float h = __fmul_rn(sp[0], hscale);
        const float s = sp[1], v = sp[2];
        float b, g, r;
        if (vec) {
            const float pre = truncf(h);
            h = __fsub_rn(h, pre);
            const float tab0 = v, tab1 = __fmul_rn(v, __fsub_rn(1.f, s)), tab2 = __fmul_rn(v, fmaf(-s, h, 1.f)), tab3 = __fmul_rn(v, fmaf(-s, __fsub_rn(1.f, h), 1.f));
            const float sec = fmaf(-truncf(__fmul_rn(pre, 1.0f / 6.0f)), 6.f, pre);       // exact: small integers
            b = sec < 2.f ? tab1 : 0.f; b = sec == 2.f ? tab3 : b; b = sec == 3.f ? tab0 : b; b = sec == 4.f ? tab0 : b; b = sec > 4.f ? tab2 : b;
            g = sec < 1.f ? tab3 : s; g = sec == 1.f ? tab0 : g; g = sec == 2.f ? tab0 : g; g = sec == 3.f ? tab2 : g; g = sec > 3.f ? tab1 : g;
            r = sec < 1.f ? tab0 : v; r = sec == 1.f ? tab2 : r; r = sec == 2.f ? tab1 : r; r = sec == 3.f ? tab1 : r; r = sec == 4.f ? tab3 : r; r = sec > 4.f ? tab0 : r;
        } else if (s == 0.f) {
            b = g = r = v;
        } else {
            const float fl = floorf(h);
            h = __fsub_rn(h, fl);
            int sector = (int)fl % 6;
            sector += sector < 0 ? 6 : 0;
            const float tab0 = v, tab1 = __fmul_rn(v, __fsub_rn(1.f, s)), tab2 = __fmul_rn(v, fmaf(-s, h, 1.f)), tab3 = __fmul_rn(v, fmaf(-s, __fsub_rn(1.f, h), 1.f));
            // sector_data rows {b, g, r} = {1,3,0},{1,0,2},{3,0,1},{0,2,1},{0,1,3},{2,1,0}
            b = sector < 2 ? tab1 : sector == 2 ? tab3 : sector < 5 ? tab0 : tab2;
            g = sector == 0 ? tab3 : sector < 3 ? tab0 : sector == 3 ? tab2 : tab1;
            r = sector == 0 ? tab0 : sector == 1 ? tab2 : sector < 4 ? tab1 : sector == 4 ? tab3 : tab0;
        }
        d[BIDX] = b; d[1] = g; d[BIDX ^ 2] = r;
        if (DCN == 4) d[3] = 1.0f;
    }
};

// CIE XYZ (color_lab.cpp is not a dispatched unit: it is compiled for the baseline ISA, so the float vectors have 4 lanes and v_fma is mul + add)
template <int SCN, int DCN> struct DOpXYZ16 {      // rows = output channels, columns = input channels, 12-bit fixed point
    int c[9];
    __device__ __forceinline__ void operator()(const unsigned short* s, unsigned short* d, bool) const
    {
#pragma unroll
        for (int r = 0; r < 3; r++) d[r] = sat_u16((s[0] * c[3 * r] + s[1] * c[3 * r + 1] + s[2] * c[3 * r + 2] + (1 << 11)) >> 12);
        if (DCN == 4) d[3] = 65535;
    }
};
template <int SCN, int DCN> struct DOpXYZ32 {
    float c[9];
    __device__ __forceinline__ void operator()(const float* s, float* d, bool vec) const
    {
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const float p0 = __fmul_rn(s[0], c[3 * r]), p1 = __fmul_rn(s[1], c[3 * r + 1]), p2 = __fmul_rn(s[2], c[3 * r + 2]);
            d[r] = vec ? __fadd_rn(p0, __fadd_rn(p1, p2)) : __fadd_rn(__fadd_rn(p0, p1), p2);
        }
        if (DCN == 4) d[3] = 1.0f;
    }
};

// thread = 4 adjacent pixels of one row; 4 * CN * sizeof(T) is a multiple of 8 for every case (of 16 for float): aligned rows move as uint2 / uint4
template <typename T, int SCN, int DCN, class Op>
__global__ void __launch_bounds__(256) cvt_depth_kernel(Img src, Img dst, Op op, int vec_cols)
{
    const int x0 = (blockIdx.x * 256 + threadIdx.x) * 4, y = blockIdx.y, f = blockIdx.z;
    if (x0 >= src.cols) return;
    const T* sp = src.row<T>(f, y) + (size_t)x0 * SCN;
    T* dp = dst.row<T>(f, y) + (size_t)x0 * DCN;
    const int n = min(4, src.cols - x0);
    // widest vector that divides the thread's bytes: 16 (float, and 4-channel 16-bit) or 8
    constexpr int SB = 4 * SCN * (int)sizeof(T), DB = 4 * DCN * (int)sizeof(T);
    constexpr int SV = SB % 16 == 0 ? 16 : 8, DV = DB % 16 == 0 ? 16 : 8;
    union { uint4 q[SB / 16 ? SB / 16 : 1]; uint2 w[SB / 8]; T e[4 * SCN]; } in;
    union { uint4 q[DB / 16 ? DB / 16 : 1]; uint2 w[DB / 8]; T e[4 * DCN]; } out;
    const bool full = n == 4;
    if (full && (((uintptr_t)sp) & (SV - 1)) == 0) {
        if constexpr (SV == 16) {
#pragma unroll
            for (int i = 0; i < SB / 16; i++) in.q[i] = __ldg((const uint4*)sp + i);
        } else {
#pragma unroll
            for (int i = 0; i < SB / 8; i++) in.w[i] = __ldg((const uint2*)sp + i);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4 * SCN; i++) in.e[i] = i < n * SCN ? sp[i] : T(0);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) op(in.e + k * SCN, out.e + k * DCN, x0 + k < vec_cols);
    if (full && (((uintptr_t)dp) & (DV - 1)) == 0) {
        if constexpr (DV == 16) {
#pragma unroll
            for (int i = 0; i < DB / 16; i++) ((uint4*)dp)[i] = out.q[i];
        } else {
#pragma unroll
            for (int i = 0; i < DB / 8; i++) ((uint2*)dp)[i] = out.w[i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4 * DCN; i++) if (i < n * DCN) dp[i] = out.e[i];
    }
}

template <typename T, int SCN, int DCN, class Op>
int launch_depth(const Img& s, const Img& d, const Op& op, cudaStream_t st, int lanes = 8)      // lanes of the reference's float vector: 8 (AVX2 units), 4 (baseline units)
{
    if (s.rows > 65535 || s.frames > 65535) return B200CV_NOT_IMPLEMENTED;
    const dim3 grid(div_up((unsigned)s.cols, 1024), (unsigned)s.rows, (unsigned)s.frames), block(256);
    cvt_depth_kernel<T, SCN, DCN, Op><<<grid, block, 0, st>>>(s, d, op, (s.cols / lanes) * lanes);
    B200_LAUNCH_CHECK();
    return B200CV_OK;
}

template <typename T>
int cvt_depth_typed(const Img& s, const Img& d, int scn, int dcn, int code, cudaStream_t st)
{
    constexpr bool F = sizeof(T) == 4;
#define NEED(sc_ok, dc_ok) B200_REQUIRE((sc_ok) && (dc_ok), "channel count does not match the colour code")
    switch (code) {
    case 0: NEED(scn == 3, dcn == 4); return launch_depth<T, 3, 4>(s, d, DOpReorder<T, 3, 4, false>(), st);     // BGR2BGRA
    case 1: NEED(scn == 4, dcn == 3); return launch_depth<T, 4, 3>(s, d, DOpReorder<T, 4, 3, false>(), st);     // BGRA2BGR
    case 2: NEED(scn == 3, dcn == 4); return launch_depth<T, 3, 4>(s, d, DOpReorder<T, 3, 4, true>(), st);      // BGR2RGBA
    case 3: NEED(scn == 4, dcn == 3); return launch_depth<T, 4, 3>(s, d, DOpReorder<T, 4, 3, true>(), st);      // RGBA2BGR
    case 4: NEED(scn == 3, dcn == 3); return launch_depth<T, 3, 3>(s, d, DOpReorder<T, 3, 3, true>(), st);      // BGR2RGB
    case 5: NEED(scn == 4, dcn == 4); return launch_depth<T, 4, 4>(s, d, DOpReorder<T, 4, 4, true>(), st);      // BGRA2RGBA
    case 6: case 7: case 10: case 11: {                                                                         // BGR2GRAY RGB2GRAY BGRA2GRAY RGBA2GRAY
        NEED(scn == ((code == 6 || code == 7) ? 3 : 4), dcn == 1);
        const bool rgb = code == 7 || code == 11;
        if constexpr (F) {
            DOpGray32 op = {rgb ? 0.299f : 0.114f, 0.587f, rgb ? 0.114f : 0.299f};
            return scn == 3 ? launch_depth<T, 3, 1>(s, d, op, st) : launch_depth<T, 4, 1>(s, d, op, st);
        } else {
            DOpGray16 op = {rgb ? 9798u : 3735u, 19235u, rgb ? 3735u : 9798u};
            return scn == 3 ? launch_depth<T, 3, 1>(s, d, op, st) : launch_depth<T, 4, 1>(s, d, op, st);
        }
    }
    case 8: NEED(scn == 1, dcn == 3); return launch_depth<T, 1, 3>(s, d, DOpGray2BGR<T, 3>(), st);
    case 9: NEED(scn == 1, dcn == 4); return launch_depth<T, 1, 4>(s, d, DOpGray2BGR<T, 4>(), st);
    case 36: case 37: case 82: case 83: {                                                                       // BGR2YCrCb RGB2YCrCb BGR2YUV RGB2YUV
        NEED(scn == 3 || scn == 4, dcn == 3);
        const bool crcb = code == 36 || code == 37;
        const int bidx = (code == 36 || code == 82) ? 0 : 2;
#define GO(OPT, B, Y, ...) do { OPT<B, Y> op = {__VA_ARGS__}; return scn == 3 ? launch_depth<T, 3, 3>(s, d, op, st) : launch_depth<T, 4, 3>(s, d, op, st); } while (0)
        if constexpr (F) {
            float c[5] = {0.299f, 0.587f, 0.114f, crcb ? 0.713f : 0.877f, crcb ? 0.564f : 0.492f};
            if (bidx == 0) { float t = c[0]; c[0] = c[2]; c[2] = t; }
            if (bidx == 0 && crcb) GO(DOpToYCC32, 0, 0, c[0], c[1], c[2], c[3], c[4]);
            if (bidx == 0) GO(DOpToYCC32, 0, 1, c[0], c[1], c[2], c[3], c[4]);
            if (crcb) GO(DOpToYCC32, 2, 0, c[0], c[1], c[2], c[3], c[4]);
            GO(DOpToYCC32, 2, 1, c[0], c[1], c[2], c[3], c[4]);
        } else {
            int c[5] = {4899, 9617, 1868, crcb ? 11682 : 14369, crcb ? 9241 : 8061};
            if (bidx == 0) { int t = c[0]; c[0] = c[2]; c[2] = t; }
            if (bidx == 0 && crcb) GO(DOpToYCC16, 0, 0, c[0], c[1], c[2], c[3], c[4]);
            if (bidx == 0) GO(DOpToYCC16, 0, 1, c[0], c[1], c[2], c[3], c[4]);
            if (crcb) GO(DOpToYCC16, 2, 0, c[0], c[1], c[2], c[3], c[4]);
            GO(DOpToYCC16, 2, 1, c[0], c[1], c[2], c[3], c[4]);
        }
#undef GO
    }
    case 38: case 39: case 84: case 85: {                                                                       // YCrCb2BGR YCrCb2RGB YUV2BGR YUV2RGB
        NEED(scn == 3, dcn == 3 || dcn == 4);
        const bool crcb = code == 38 || code == 39;
        const int bidx = (code == 38 || code == 84) ? 0 : 2;
#define GO(OPT, B, Y, ...) do { if (dcn == 3) { OPT<B, Y, 3> op = {__VA_ARGS__}; return launch_depth<T, 3, 3>(s, d, op, st); } \
                                else { OPT<B, Y, 4> op = {__VA_ARGS__}; return launch_depth<T, 3, 4>(s, d, op, st); } } while (0)
        if constexpr (F) {
            const float c0 = crcb ? 1.403f : 1.140f, c1 = crcb ? -0.714f : -0.581f, c2 = crcb ? -0.344f : -0.395f, c3 = crcb ? 1.773f : 2.032f;
            if (bidx == 0 && crcb) GO(DOpFromYCC32, 0, 0, c0, c1, c2, c3);
            if (bidx == 0) GO(DOpFromYCC32, 0, 1, c0, c1, c2, c3);
            if (crcb) GO(DOpFromYCC32, 2, 0, c0, c1, c2, c3);
            GO(DOpFromYCC32, 2, 1, c0, c1, c2, c3);
        } else {
            const int c0 = crcb ? 22987 : 18678, c1 = crcb ? -11698 : -9519, c2 = crcb ? -5636 : -6472, c3 = crcb ? 29049 : 33292;
            if (bidx == 0 && crcb) GO(DOpFromYCC16, 0, 0, c0, c1, c2, c3);
            if (bidx == 0) GO(DOpFromYCC16, 0, 1, c0, c1, c2, c3);
            if (crcb) GO(DOpFromYCC16, 2, 0, c0, c1, c2, c3);
            GO(DOpFromYCC16, 2, 1, c0, c1, c2, c3);
        }
#undef GO
    }
    case 40: case 41: case 66: case 67: case 54: case 55: case 70: case 71: {                                   // BGR/RGB <-> HSV (_FULL): float only
        if constexpr (!F) return B200CV_NOT_IMPLEMENTED;
        else {
            const bool fwd = code == 40 || code == 41 || code == 66 || code == 67;
            const bool blue_first = code == 40 || code == 66 || code == 54 || code == 70;
            NEED(fwd ? (scn == 3 || scn == 4) : scn == 3, fwd ? dcn == 3 : (dcn == 3 || dcn == 4));
            if (fwd) {
                const float hscale = 360.f * (1.f / 360.f);        // hrange * (1.f / 360.f), hrange = 360 for float images (color_hsv.simd.hpp:308, color_hsv.dispatch.cpp)
                if (blue_first) { DOpToHSV32<0> op = {hscale}; return scn == 3 ? launch_depth<T, 3, 3>(s, d, op, st) : launch_depth<T, 4, 3>(s, d, op, st); }
                DOpToHSV32<2> op = {hscale};
                return scn == 3 ? launch_depth<T, 3, 3>(s, d, op, st) : launch_depth<T, 4, 3>(s, d, op, st);
            }
            const float hscale = 6.f / 360.f;
            if (blue_first && dcn == 3) { DOpFromHSV32<0, 3> op = {hscale}; return launch_depth<T, 3, 3>(s, d, op, st); }
            if (blue_first) { DOpFromHSV32<0, 4> op = {hscale}; return launch_depth<T, 3, 4>(s, d, op, st); }
            if (dcn == 3) { DOpFromHSV32<2, 3> op = {hscale}; return launch_depth<T, 3, 3>(s, d, op, st); }
            DOpFromHSV32<2, 4> op = {hscale};
            return launch_depth<T, 3, 4>(s, d, op, st);
        }
    }
    case 32: case 33: case 34: case 35: {                                                                       // BGR2XYZ RGB2XYZ XYZ2BGR XYZ2RGB
        const bool fwd = code <= 33;
        NEED(fwd ? (scn == 3 || scn == 4) : scn == 3, fwd ? dcn == 3 : (dcn == 3 || dcn == 4));
        // sRGB2XYZ_D65 / XYZ2sRGB_D65 (color_lab.cpp:103-144) as float(softdouble) and as 12-bit integers; blue-first images swap the
        // matrix's columns (to XYZ) or rows (from XYZ)
        static const float FWD_F[9] = {0.412453f, 0.357580f, 0.180423f, 0.212671f, 0.715160f, 0.072169f, 0.019334f, 0.119193f, 0.950227f};
        static const float INV_F[9] = {3.240479f, -1.53715f, -0.498535f, -0.969256f, 1.875991f, 0.041556f, 0.055648f, -0.204043f, 1.057311f};
        static const int FWD_I[9] = {1689, 1465, 739, 871, 2929, 296, 79, 488, 3892}, INV_I[9] = {13273, -6296, -2042, -3970, 7684, 170, 228, -836, 4331};
        const bool blue_first = code == 32 || code == 34;
        int idx[9];
        for (int r = 0; r < 3; r++)
            for (int k = 0; k < 3; k++) idx[3 * r + k] = fwd ? 3 * r + (blue_first ? 2 - k : k) : 3 * (blue_first ? 2 - r : r) + k;
#define GO(OPT, TAB, ...) do { if (scn == 3 && dcn == 3) { OPT<3, 3> op; for (int i = 0; i < 9; i++) op.c[i] = TAB[idx[i]]; return launch_depth<T, 3, 3>(s, d, op, st, __VA_ARGS__); } \
                               if (scn == 4) { OPT<4, 3> op; for (int i = 0; i < 9; i++) op.c[i] = TAB[idx[i]]; return launch_depth<T, 4, 3>(s, d, op, st, __VA_ARGS__); } \
                               OPT<3, 4> op; for (int i = 0; i < 9; i++) op.c[i] = TAB[idx[i]]; return launch_depth<T, 3, 4>(s, d, op, st, __VA_ARGS__); } while (0)
        if constexpr (F) { if (fwd) GO(DOpXYZ32, FWD_F, 4); else GO(DOpXYZ32, INV_F, 4); }
        else { if (fwd) GO(DOpXYZ16, FWD_I, 8); else GO(DOpXYZ16, INV_I, 8); }
#undef GO
    }
    default: return B200CV_NOT_IMPLEMENTED;
    }
#undef NEED
}

}  // namespace

// 16-bit unsigned and float images (same depth on both sides); every other code / depth: NOT_IMPLEMENTED (a stock OpenCV then runs its own code)
int cvt_color_depth(const b200cvMat* src, const b200cvMat* dst, int code, cudaStream_t st)
{
    const int sd = B200CV_DEPTH(src->type), dd = B200CV_DEPTH(dst->type);
    if (sd != dd || (sd != B200CV_16U && sd != B200CV_32F)) return B200CV_NOT_IMPLEMENTED;
    B200_REQUIRE(src->data != dst->data, "cvtColor: in-place is not supported");
    const Img s = make_img(src), d = make_img(dst);
    const int scn = B200CV_CN(src->type), dcn = B200CV_CN(dst->type);
    return sd == B200CV_32F ? cvt_depth_typed<float>(s, d, scn, dcn, code, st) : cvt_depth_typed<unsigned short>(s, d, scn, dcn, code, st);
}

}  // namespace b200cv
