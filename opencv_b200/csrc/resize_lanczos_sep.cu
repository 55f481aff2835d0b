// resize_lanczos_sep.cu -- cv::resize INTER_LANCZOS4 for 8-bit images, second version: tiled and separable (the tables, the first version and the
// float path are in resize_lanczos.cu).
#include <algorithm>
#include "common.cuh"
#include "resize.cuh"

namespace b200cv {

namespace {

// The first version filters every source row once per destination row that uses it (8 x): 64 byte gathers + 64 MACs per element.  Here a CTA
// owns LS_DW x DH destination pixels: H pass = every (source row of the tile, destination column, channel) once -- 8 byte gathers through L1 +
// 8 IMAD into an int32 row in shared memory (the reference's HResizeLanczos4 buf row, resize.cpp:2066-2118); V pass = 4 adjacent elements per
// item: 8 x LDS.128 + 32 IMAD in wrapping 32-bit arithmetic like the reference's int sums (:2120-2158), (v + 2^21) >> 22, saturate, one
// 32-bit store.  Same tables, same integers: bit-identical to the first version.
constexpr int LS_DW = 64;

template <int CN> struct LSCfg { static constexpr int NT = CN == 3 ? 384 : 256; };     // threads: a multiple of the tile's element columns (64 * CN)

template <int CN>
__global__ void __launch_bounds__(LSCfg<CN>::NT, CN == 3 ? 3 : 4) resize_lanczos4_sep_kernel(Img src, Img dst, const LzTap* __restrict__ xt, const LzTap* __restrict__ yt, int sw, int sh, int dw, int dh,
                                                                  int DH, int RMAX)
{
    constexpr int E = LS_DW * CN;
    extern __shared__ __align__(16) int ls_mid[];                 // [RMAX][E]
    const int tid = threadIdx.x, f = blockIdx.z;
    const int x0 = blockIdx.x * LS_DW, y0 = blockIdx.y * DH;
    const int nrows = min(DH, dh - y0), ncols = min(LS_DW, dw - x0), ne = ncols * CN;
    const int row_first = yt[y0].s - 3;
    const int R = min(yt[y0 + nrows - 1].s + 4 - row_first + 1, RMAX);
    // H pass: a thread owns one element column (offsets and taps stay in registers) and every G-th source row of the tile; four rows are in flight
    // at a time (32 independent byte gathers: the pass is latency bound otherwise)
    constexpr int NT = LSCfg<CN>::NT, G = NT / E;
    {
        const int e = tid % E, g = tid / E;
        if (e < ne) {
            const int x = e / CN, c = e - x * CN;
            const LzTap tx = xt[x0 + x];
            int xi[8];
#pragma unroll
            for (int j = 0; j < 8; j++) xi[j] = lz_clip(tx.s - 3 + j, sw) * CN + c;
            int r = g;
            for (; r + 3 * G < R; r += 4 * G) {
                int t[4] = {0, 0, 0, 0};
                unsigned char b[4][8];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uchar* rp = src.row<uchar>(f, lz_clip(row_first + r + u * G, sh));
#pragma unroll
                    for (int j = 0; j < 8; j++) b[u][j] = __ldg(rp + xi[j]);
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
#pragma unroll
                    for (int j = 0; j < 8; j++) t[u] += b[u][j] * tx.ic[j];
                    ls_mid[(r + u * G) * E + e] = t[u];
                }
            }
            for (; r < R; r += G) {
                const uchar* rp = src.row<uchar>(f, lz_clip(row_first + r, sh));
                int t = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) t += __ldg(rp + xi[j]) * tx.ic[j];
                ls_mid[r * E + e] = t;
            }
        }
    }
    __syncthreads();
    // V pass: item = 4 adjacent elements of one destination row
    const bool vec_store = (((uintptr_t)dst.data | dst.step | dst.fstep) & 3) == 0 && ((x0 * CN) & 3) == 0;
    const int nq = (ne + 3) >> 2;
    for (int it = tid; it < nrows * (E / 4); it += NT) {
        const int row = it / (E / 4), q = it - row * (E / 4);
        if (q >= nq) continue;
        const LzTap ty = yt[y0 + row];
        const int m0 = ty.s - 3 - row_first;
        unsigned v[4] = {0, 0, 0, 0};                                // unsigned: the reference's int sums wrap
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int4 t = *(const int4*)(ls_mid + (m0 + k) * E + 4 * q);
            const unsigned b = (unsigned)(int)ty.ic[k];
            v[0] += (unsigned)t.x * b; v[1] += (unsigned)t.y * b; v[2] += (unsigned)t.z * b; v[3] += (unsigned)t.w * b;
        }
        unsigned out = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) out |= (unsigned)sat_u8((int)(v[i] + (1u << 21)) >> 22) << (8 * i);
        uchar* d = dst.row<uchar>(f, y0 + row) + (size_t)x0 * CN + 4 * q;
        if (vec_store && 4 * q + 4 <= ne) *(unsigned*)d = out;
        else for (int i = 0; i < 4 && 4 * q + i < ne; i++) d[i] = (uchar)(out >> (8 * i));
    }
}

template <int CN>
static bool launch_lanczos_sep(const Img& s, const Img& d, const LzTap* host_yt, const LzTap* xt, const LzTap* yt, cudaStream_t st)
{
    constexpr int E = LS_DW * CN;
    const int dh = d.rows;
    for (int DH : {32, 16, 8, 4}) {
        int rmax = 0;
        for (int y0 = 0; y0 < dh; y0 += DH) rmax = std::max(rmax, host_yt[std::min(y0 + DH, dh) - 1].s + 4 - (host_yt[y0].s - 3) + 1);
        const size_t smem = (size_t)rmax * E * 4;
        if (smem > 44 * 1024) continue;
        const dim3 grid(div_up((unsigned)d.cols, LS_DW), div_up((unsigned)dh, (unsigned)DH), (unsigned)s.frames);
        resize_lanczos4_sep_kernel<CN><<<grid, LSCfg<CN>::NT, smem, st>>>(s, d, xt, yt, s.cols, s.rows, d.cols, dh, DH, rmax);
        return true;
    }
    return false;       // extreme decimation: the per-element kernel
}

}  // namespace

bool resize_lanczos_sep_u8(const Img& s, const Img& d, int cn, const LzTap* host_yt, const LzTap* xt, const LzTap* yt, cudaStream_t st)
{
    if (d.rows >= 65536 * 4 || s.frames >= 65536) return false;
    const bool ok = cn == 1 ? launch_lanczos_sep<1>(s, d, host_yt, xt, yt, st) : cn == 3 ? launch_lanczos_sep<3>(s, d, host_yt, xt, yt, st)
                                                                                       : cn == 4 ? launch_lanczos_sep<4>(s, d, host_yt, xt, yt, st) : false;
    return ok;
}

}  // namespace b200cv
