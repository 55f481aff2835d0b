// resize_sep.cu -- cv::resize INTER_LINEAR / INTER_CUBIC for 8-bit images as a TILED SEPARABLE filter (second version of the
// per-pixel kernels in resize.cu, which stay for float images and as the fallback).
//
// Arithmetic = the reference's, bit for bit (resize.cpp): horizontal pass in integers with 11-bit taps (HResizeLinear :1877-1928,
// HResizeCubic :1993-2041, per-tap index clamping at the image edges), vertical pass
//   LINEAR  ((b0 * (T0 >> 4)) >> 16) + ((b1 * (T1 >> 4)) >> 16) + 2) >> 2                                (VResizeLinear<uchar> :1963-1989)
//   CUBIC   float S0*b0 + (S1*b1 + (S2*b2 + S3*b3)), b = beta / 2^22, round half even, for the first floor8(width*cn) elements of a row
//           and (sum + 2^21) >> 22 in integers for the rest                         (VResizeCubicVec_32s8u :1408-1444, FixedPtCast :2059-2060)
//
// Why: the per-pixel kernels filter every source row once per destination row that uses it (2x / 4x the horizontal work), convert with
// I2F / F2I (quarter-rate unit: 15 conversions per 3-channel CUBIC pixel) and leave through byte stores; they ran at 0.06-0.36 of the
// HBM roofline, LSU / issue bound.  Here one CTA owns DW x DH destination pixels:
//   H pass   thread = one destination pixel column (its source offset and taps stay in registers) walking down the R source rows the tile
//            needs: aligned word loads + funnel-shift realignment + PRMT tap pairs + IDP2A, exactly once per (column, source row); the
//            filtered row goes to shared memory as u16 (LINEAR, T >> 4 < 2^15) or float (CUBIC: the IDP2A chain starts at the bit pattern
//            of 1.5 * 2^23, so the integer sum IS a float after one FADD -- no I2F);
//   V pass   item = 4 adjacent byte elements of one destination row: 64-/128-bit shared loads, LINEAR two IMAD.HI per element against
//            taps pre-shifted by 16, CUBIC the reference's float chain and rounding by adding 1.5 * 2^23 (no F2I); 32-bit stores.
// Tables (source index + taps per destination column / row) come from resize_tab_kernel (resize.cu), built with the reference's fp64 / float sequence.
#include <string.h>
#include <math.h>
#include <type_traits>
#include "resize.cuh"

namespace b200cv {

template <int CN, bool CUBIC> struct RSCfg {
    static constexpr int DW = CUBIC ? 128 : 256;         // destination pixels per tile row
    static constexpr int E = DW * CN;                    // byte elements per tile row
    static constexpr int NT = CUBIC ? 4 : 2;             // taps
    static constexpr int MIDB = CUBIC ? 4 : 2;           // bytes per element of the filtered rows
    static constexpr int NQ = E / 4;                     // V-pass items per row
};

struct RSRow {              // per destination row of the tile (shared memory)
    int r[4];               // filtered-row index (relative to the tile's first source row) per tap
    union { unsigned bs[4]; float bf[4]; int bi[4]; };    // LINEAR: taps << 16; CUBIC: float(beta) * 2^-22
    int ib[4];              // CUBIC: the integer taps (tail elements)
};

#ifdef B200CV_HOST_EMULATION
static inline int rs_dp2a_s(int a, unsigned b, int c) { return c + (int)(short)(a & 0xffff) * (int)(b & 0xffu) + (a >> 16) * (int)((b >> 8) & 0xffu); }
static inline unsigned rs_umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned rs_funnel_r(unsigned lo, unsigned hi, unsigned sh) { return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
static inline unsigned rs_madhi(unsigned a, unsigned b, unsigned c) { return rs_umulhi(a, b) + c; }
#else
// dp2a.lo.s32.u32: a = two s16 taps, b = bytes 0 and 1: c + a.lo * b.byte0 + a.hi * b.byte1
__device__ __forceinline__ int rs_dp2a_s(int a, unsigned b, int c) { int d; asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ unsigned rs_umulhi(unsigned a, unsigned b) { return __umulhi(a, b); }
__device__ __forceinline__ unsigned rs_funnel_r(unsigned lo, unsigned hi, unsigned sh) { return __funnelshift_r(lo, hi, sh); }
// hi(a * b) + c in one IMAD.HI (the compiler fuses one addend on its own, not the rounding constant)
__device__ __forceinline__ unsigned rs_madhi(unsigned a, unsigned b, unsigned c) { unsigned d; asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
#endif

// NW words starting AT byte A of the row (A need not be aligned): NW + 1 aligned loads + funnel shifts
template <int NW>
__device__ __forceinline__ void rs_load_realigned(const uchar* row, unsigned A, unsigned* w)
{
    const unsigned* q = (const unsigned*)(row + (A & ~3u));
    const unsigned sh8 = 8 * (A & 3u);
    unsigned t[NW + 1];
#pragma unroll
    for (int i = 0; i <= NW; i++) t[i] = q[i];
#pragma unroll
    for (int i = 0; i < NW; i++) w[i] = rs_funnel_r(t[i], t[i + 1], sh8);
}
// bytes (c + j0*CN, c + (j0+1)*CN) of the realigned words in bytes 0, 1 (what dp2a.lo reads)
template <int CN> __device__ __forceinline__ unsigned rs_pair(const unsigned* w, int c, int j0)
{
    const int p0 = c + j0 * CN, p1 = p0 + CN;
    return __byte_perm(w[p0 >> 2], w[p1 >> 2], (unsigned)((p0 & 3) | ((4 + (p1 & 3)) << 4)));
}

constexpr float RS_MAGIC = 12582912.f;              // 1.5 * 2^23: float(0x4B400000 + t) == RS_MAGIC + t for |t| < 2^22
constexpr int RS_MAGIC_I = 0x4B400000;

// ---- H pass: one destination pixel column, rows r = r_first, r_first + r_step, ... < R of the tile --------------------------------------
// Rows are taken four at a time: all the loads of a group are issued before the first is consumed (the pass is latency bound otherwise), and
// the global / shared pointers advance by additions (the first version recomputed 64-bit row addresses per row: 30 IMAD per pixel).
template <int CN, bool CUBIC>
__device__ __forceinline__ void rs_hpass_thread(const Img& src, int f, const ResizeParams& p, const ResTab& tx, int row_lo, int R, int r_first, int r_step,
                                                unsigned char* mid_col /* &mid[0][col * CN] */)
{
    constexpr int E = RSCfg<CN, CUBIC>::E;
    constexpr int NT = CUBIC ? 4 : 2;
    constexpr int NW = (NT * CN + 3) / 4;
    typedef typename std::conditional<CUBIC, float, unsigned short>::type MidT;
    const bool words_ok = (((uintptr_t)src.data | src.step | src.fstep) & 3) == 0;
    const int row_bytes = p.sw * CN;
    const int A = CUBIC ? (tx.s - 1) * CN : tx.s * CN;                     // first source byte of the taps (before clamping)
    const bool last_col = !CUBIC && tx.last != 0;
    const bool fast = words_ok && (CUBIC ? (tx.s >= 1 && tx.s + 2 <= p.sw - 1) : !last_col) && A + NT * CN + 7 <= row_bytes;   // the NW + 1 words stay inside the row
    const int c01 = (tx.ic[0] & 0xffff) | (tx.ic[1] << 16), c23 = CUBIC ? (tx.ic[2] & 0xffff) | (tx.ic[3] << 16) : 0;
    const size_t gstep = (size_t)r_step * src.step;
    const uchar* gp = src.data + (size_t)f * src.fstep + (size_t)(row_lo + r_first) * src.step;
    MidT* m = (MidT*)mid_col + (size_t)r_first * E;
    const int mstep = r_step * E;
    int r = r_first;
    if (fast) {
        const uchar* wp = gp + (A & ~3);
        const unsigned sh8 = 8u * ((unsigned)A & 3u);
        auto finish = [&](const unsigned* t, MidT* mo) {
            unsigned w[NW];
#pragma unroll
            for (int i = 0; i < NW; i++) w[i] = rs_funnel_r(t[i], t[i + 1], sh8);
#pragma unroll
            for (int c = 0; c < CN; c++) {
                if constexpr (CUBIC) {
                    const int v = rs_dp2a_s(c23, rs_pair<CN>(w, c, 2), rs_dp2a_s(c01, rs_pair<CN>(w, c, 0), RS_MAGIC_I));
                    mo[c] = __fsub_rn(__int_as_float(v), RS_MAGIC);        // exact: |sum| < 2^22
                } else {
                    mo[c] = (unsigned short)(rs_dp2a_s(c01, rs_pair<CN>(w, c, 0), 0) >> 4);
                }
            }
        };
        for (; r + 3 * r_step < R; r += 4 * r_step) {
            unsigned t[4][NW + 1];
#pragma unroll
            for (int g = 0; g < 4; g++)
#pragma unroll
                for (int i = 0; i <= NW; i++) t[g][i] = ((const unsigned*)(wp + g * gstep))[i];
#pragma unroll
            for (int g = 0; g < 4; g++) finish(t[g], m + g * mstep);
            wp += 4 * gstep; m += 4 * mstep;
        }
        for (; r < R; r += r_step) {
            unsigned t[NW + 1];
#pragma unroll
            for (int i = 0; i <= NW; i++) t[i] = ((const unsigned*)wp)[i];
            finish(t, m);
            wp += gstep; m += mstep;
        }
    } else {
        int xi[4];
#pragma unroll
        for (int j = 0; j < 4; j++) xi[j] = CUBIC ? min(max(tx.s - 1 + j, 0), p.sw - 1) * CN : (tx.s + (j & 1)) * CN;   // per-tap clamping == the while-loops of HResizeCubic
        for (; r < R; r += r_step, gp += gstep, m += mstep) {
#pragma unroll
            for (int c = 0; c < CN; c++) {
                if constexpr (CUBIC) {
                    const int v = RS_MAGIC_I + gp[xi[0] + c] * tx.ic[0] + gp[xi[1] + c] * tx.ic[1] + gp[xi[2] + c] * tx.ic[2] + gp[xi[3] + c] * tx.ic[3];
                    m[c] = __fsub_rn(__int_as_float(v), RS_MAGIC);
                } else {
                    const int v = last_col ? gp[xi[0] + c] * 2048 : gp[xi[0] + c] * tx.ic[0] + gp[xi[1] + c] * tx.ic[1];
                    m[c] = (unsigned short)(v >> 4);
                }
            }
        }
    }
}

// ---- V pass: item = 4 adjacent byte elements (tile element e0 .. e0+3) of destination row `row` of the tile -------------------------------
template <int CN, bool CUBIC>
__device__ __forceinline__ unsigned rs_vpass_item(const unsigned char* const* mrow /* filtered row per tap */, const RSRow& yr, int e0, int ge0, int vec_limit)
{
    unsigned out = 0;
    if constexpr (!CUBIC) {
        const uint2 m0 = *(const uint2*)(mrow[0] + e0 * 2), m1 = *(const uint2*)(mrow[1] + e0 * 2);
        const unsigned T0[4] = {m0.x & 0xffffu, m0.x >> 16, m0.y & 0xffffu, m0.y >> 16}, T1[4] = {m1.x & 0xffffu, m1.x >> 16, m1.y & 0xffffu, m1.y >> 16};
        // (b * T) >> 16 per tap and element (VResizeLinear's two separately floored products): the products fit 32 bits (b <= 2048, T <= 32640), so
        // a plain IMAD and the upper half-word do it -- IMAD.HI issues at a fraction of IMAD's rate.  One PRMT takes the upper halves of TWO
        // elements' products into one register; the sums (<= 1023 + 2) of two elements are then added, rounded and shifted together.
        const unsigned b0 = yr.bs[0] >> 16, b1 = yr.bs[1] >> 16;
        const unsigned s01 = __byte_perm(b0 * T0[0], b0 * T0[1], 0x7632) + __byte_perm(b1 * T1[0], b1 * T1[1], 0x7632) + 0x00020002u;
        const unsigned s23 = __byte_perm(b0 * T0[2], b0 * T0[3], 0x7632) + __byte_perm(b1 * T1[2], b1 * T1[3], 0x7632) + 0x00020002u;
        out = __byte_perm(s01 >> 2, s23 >> 2, 0x6420);                     // ((sum + 2) >> 2) <= 255: bytes 0 and 2 of each pair, no mask
    } else {
        float S[4][4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint4 q = *(const uint4*)(mrow[k] + e0 * 4);
            S[k][0] = __int_as_float((int)q.x); S[k][1] = __int_as_float((int)q.y); S[k][2] = __int_as_float((int)q.z); S[k][3] = __int_as_float((int)q.w);
        }
        int v[4];
        if (ge0 < vec_limit) {          // ge0 is a multiple of 4, vec_limit of 8: the 4 elements are on the same side
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float a = __fmul_rn(S[3][i], yr.bf[3]);
                a = __fadd_rn(__fmul_rn(S[2][i], yr.bf[2]), a);
                a = __fadd_rn(__fmul_rn(S[1][i], yr.bf[1]), a);
                a = __fadd_rn(__fmul_rn(S[0][i], yr.bf[0]), a);
                // cvRound: |a| < 2^22, so adding 1.5 * 2^23 rounds to the nearest integer, ties to even, like cvtps2dq
                v[i] = __float_as_int(__fadd_rn(a, RS_MAGIC)) - RS_MAGIC_I;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int t0 = __float2int_rn(S[0][i]), t1 = __float2int_rn(S[1][i]), t2 = __float2int_rn(S[2][i]), t3 = __float2int_rn(S[3][i]);
                v[i] = (t0 * yr.ib[0] + t1 * yr.ib[1] + t2 * yr.ib[2] + t3 * yr.ib[3] + (1 << 21)) >> 22;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = min(max(v[i], 0), 255);
        out = __byte_perm(__byte_perm((unsigned)v[0], (unsigned)v[1], 0x0040), __byte_perm((unsigned)v[2], (unsigned)v[3], 0x0040), 0x5410);
    }
    (void)ge0; (void)vec_limit;
    return out;
}

// per-tile set-up shared by the kernel and the host emulation: the tile's source row range and the per-row records
template <bool CUBIC>
__device__ __forceinline__ void rs_tile_rows(const ResTab* yt, int y0, int nrows_out, int sh, int& row_lo, int& R)
{
    const int first = yt[y0].s - (CUBIC ? 1 : 0), last = yt[y0 + nrows_out - 1].s + (CUBIC ? 2 : 1);
    row_lo = clip_i(first, 0, sh);
    R = clip_i(last, 0, sh) - row_lo + 1;
}
template <bool CUBIC>
__device__ __forceinline__ void rs_fill_row(RSRow& o, const ResTab& ty, int row_lo, int sh)
{
    constexpr int NT = CUBIC ? 4 : 2;
#pragma unroll
    for (int k = 0; k < 4; k++) { o.r[k] = 0; o.bi[k] = 0; o.ib[k] = 0; }
#pragma unroll
    for (int k = 0; k < NT; k++) {
        o.r[k] = clip_i(ty.s - (CUBIC ? 1 : 0) + k, 0, sh) - row_lo;            // rows are clipped when fetched, the taps keep their values (:2211)
        if (CUBIC) { o.bf[k] = __fmul_rn(__int2float_rn(ty.ic[k]), 1.f / (2048.f * 2048.f)); o.ib[k] = ty.ic[k]; }
        else o.bs[k] = (unsigned)ty.ic[k] << 16;
    }
}

// a warp takes whole destination rows of the tile (its lanes the 4-element items of the row): the row record is read once per row,
// pointers advance by additions, no divisions
template <int CN, bool CUBIC>
__device__ __forceinline__ void rs_vpass_thread(int tid, int nthreads, const unsigned char* mid, const RSRow* yrow, const Img& dst, int f, const ResizeParams& p,
                                                int x0, int y0, int nrows_out, int ncols_out)
{
    constexpr int NQ = RSCfg<CN, CUBIC>::NQ;
    const int vec_limit = ((p.dw * CN) / 8) * 8;
    const bool vec_store = (((uintptr_t)dst.data | dst.step | dst.fstep) & 3) == 0;
    const int ne = ncols_out * CN;                                   // valid elements per row of this tile
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
    const int nq = min(NQ, (ne + 3) >> 2);
    uchar* drow = dst.data + (size_t)f * dst.fstep + (size_t)(y0 + warp) * dst.step + (size_t)x0 * CN;
    const size_t dstep = (size_t)nwarps * dst.step;
    for (int row = warp; row < nrows_out; row += nwarps, drow += dstep) {
        const RSRow yr = yrow[row];
        const unsigned char* mrow[4];
#pragma unroll
        for (int k = 0; k < 4; k++) mrow[k] = mid + (size_t)yr.r[k] * (RSCfg<CN, CUBIC>::E * RSCfg<CN, CUBIC>::MIDB);
        if (vec_store && ne == RSCfg<CN, CUBIC>::E) {          // full-width tile: every offset of the row is a compile-time constant past the lane's base
#pragma unroll
            for (int i = 0; i < NQ / 32; i++) {
                const int e0 = 4 * (lane + 32 * i);
                *(unsigned*)(drow + e0) = rs_vpass_item<CN, CUBIC>(mrow, yr, e0, x0 * CN + e0, vec_limit);
            }
            continue;
        }
        for (int q = lane; q < nq; q += 32) {
            const int e0 = 4 * q;
            const unsigned out = rs_vpass_item<CN, CUBIC>(mrow, yr, e0, x0 * CN + e0, vec_limit);
            uchar* d = drow + e0;
            if (vec_store && e0 + 4 <= ne) *(unsigned*)d = out;
            else for (int i = 0; i < 4 && e0 + i < ne; i++) d[i] = (uchar)(out >> (8 * i));
        }
    }
}

#ifndef B200CV_HOST_EMULATION
template <int CN, bool CUBIC>
__global__ void __launch_bounds__(256) resize_sep_u8_kernel(const Img src, const Img dst, const ResizeParams p, const ResTab* __restrict__ xt, const ResTab* __restrict__ yt,
                                                            int DH, int RMAX)
{
    typedef RSCfg<CN, CUBIC> C;
    extern __shared__ __align__(16) unsigned char rs_smem[];
    unsigned char* mid = rs_smem;                                              // [RMAX][E] u16 / float
    RSRow* yrow = (RSRow*)(rs_smem + (size_t)RMAX * C::E * C::MIDB);           // [DH]
    const int tid = threadIdx.x, f = blockIdx.z;
    const int x0 = blockIdx.x * C::DW, y0 = blockIdx.y * DH;
    const int nrows_out = min(DH, p.dh - y0), ncols_out = min(C::DW, p.dw - x0);
    int row_lo, R;
    rs_tile_rows<CUBIC>(yt, y0, nrows_out, p.sh, row_lo, R);
    if (R > RMAX) R = RMAX;                                                    // never true: RMAX is the exact maximum over all tiles (host)
    if (tid < nrows_out) rs_fill_row<CUBIC>(yrow[tid], yt[y0 + tid], row_lo, p.sh);
    {
        constexpr int TPC = 256 / C::DW;                                       // threads per pixel column (1 or 2): they interleave the rows
        const int col = tid % C::DW, rpar = tid / C::DW;
        if (col < ncols_out) rs_hpass_thread<CN, CUBIC>(src, f, p, xt[x0 + col], row_lo, R, rpar, TPC, mid + (size_t)col * CN * C::MIDB);
    }
    __syncthreads();
    rs_vpass_thread<CN, CUBIC>(tid, 256, mid, yrow, dst, f, p, x0, y0, nrows_out, ncols_out);
}
#endif

// the reference's source row of destination row d (resize.cpp:4097-4124 for rows: no edge clamping of the index itself)
static int rs_host_src_row(int d, double scale, bool area_mode)
{
    if (!area_mode) {
        volatile double t = ((double)d + 0.5) * scale;
        const float fx = (float)(t - 0.5);
        return (int)floorf(fx);
    }
    volatile double t = (double)d * scale;
    return (int)floor(t);
}

// exact maximum, over the tiles of DH destination rows, of the number of distinct source rows a tile reads
static int rs_host_rmax(const ResizeParams& p, bool cubic, int DH)
{
    int rmax = 1;
    for (int y0 = 0; y0 < p.dh; y0 += DH) {
        const int y1 = (y0 + DH < p.dh ? y0 + DH : p.dh) - 1;
        const int first = rs_host_src_row(y0, p.scale_y, p.area_mode != 0) - (cubic ? 1 : 0), last = rs_host_src_row(y1, p.scale_y, p.area_mode != 0) + (cubic ? 2 : 1);
        const int R = clip_i(last, 0, p.sh) - clip_i(first, 0, p.sh) + 1;
        if (R > rmax) rmax = R;
    }
    return rmax;
}

#ifndef B200CV_HOST_EMULATION
template <int CN, bool CUBIC>
static int launch_rs(const Img& s, const Img& d, const ResizeParams& p, const ResTab* xt, const ResTab* yt, cudaStream_t st)
{
    typedef RSCfg<CN, CUBIC> C;
    // destination rows per tile: as many as keep the filtered rows + row records under ~40 KB (4-5 CTAs per SM)
    int DH = 0, RMAX = 0;
    for (int cand : {64, 48, 32, 24, 16, 8, 4}) {
        const int r = rs_host_rmax(p, CUBIC, cand);
        if ((size_t)r * C::E * C::MIDB + (size_t)cand * sizeof(RSRow) <= 40 * 1024) { DH = cand; RMAX = r; break; }
    }
    if (!DH) return B200CV_NOT_IMPLEMENTED;                 // extreme decimation: the per-pixel kernels take it
    const size_t smem = (size_t)RMAX * C::E * C::MIDB + (size_t)DH * sizeof(RSRow);
    dim3 grid(div_up((unsigned)p.dw, C::DW), div_up((unsigned)p.dh, (unsigned)DH), (unsigned)s.frames);
    resize_sep_u8_kernel<CN, CUBIC><<<grid, 256, smem, st>>>(s, d, p, xt, yt, DH, RMAX);
    cudaError_t e = cudaGetLastError();
    count_launch();
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch", __FILE__, __LINE__);
    return B200CV_OK;
}

int resize_sep_u8(const Img& s, const Img& d, int cn, bool cubic, const ResizeParams& p, const ResTab* xt, const ResTab* yt, cudaStream_t st)
{
    if (d.rows >= 65536 * 4) return B200CV_NOT_IMPLEMENTED;
    if (cubic) {
        if (cn == 1) return launch_rs<1, true>(s, d, p, xt, yt, st);
        if (cn == 3) return launch_rs<3, true>(s, d, p, xt, yt, st);
        if (cn == 4) return launch_rs<4, true>(s, d, p, xt, yt, st);
    } else {
        if (cn == 1) return launch_rs<1, false>(s, d, p, xt, yt, st);
        if (cn == 3) return launch_rs<3, false>(s, d, p, xt, yt, st);
        if (cn == 4) return launch_rs<4, false>(s, d, p, xt, yt, st);
    }
    return B200CV_NOT_IMPLEMENTED;
}
#endif

}  // namespace b200cv
