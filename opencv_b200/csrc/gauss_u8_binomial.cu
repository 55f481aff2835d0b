// gauss_u8_binomial.cu -- cv::GaussianBlur on 8-bit single-channel images for the two kernels almost every caller uses: 3 x 3 and 5 x 5 with
// sigma = 0, whose 8.8 fixed-point taps are the binomial rows {64, 128, 64} and {16, 64, 96, 64, 16} (smooth.dispatch.cpp:87-146, 224-258).
//
// Reference arithmetic (fixedSmoothInvoker<uint8_t, ufixedpoint16>, smooth.simd.hpp:1925-2197): rows in 8.8, columns in 16.16,
//   dst = (sum_j ky[j] * sum_i kx[i] * src + 32768) >> 16,  no intermediate saturation (the taps add up to 256).
// With kx = ky = 64 * {1, 2, 1} this is exactly (B + 8) >> 4, B = the 3 x 3 binomial sum (<= 16 * 255); with 16 * {1, 4, 6, 4, 1} it is
// (B + 128) >> 8, B <= 256 * 255 -- both fit 16 bits, rounding constant included.  So two pixels ride in one 32-bit register as 16-bit halves
// and the whole filter is adds and shifts on the integer pipes, with no multiplier in the loop:
//   IDP4A / IDP2A issue at half rate on the FMA-heavy pipe; the general tile kernel (gauss_u8.cu) spends ~3.5 of the ~11 issue slots per pixel that
//   the HBM roofline leaves on them alone (12.1 thread-instructions per pixel at 3 x 3: 0.58 of the roofline, profiles/r02_notes.md section 2).
// A thread owns 4 adjacent columns (one 32-bit word of pixels) and walks down GB_SEG output rows: per input row three aligned word loads
// (its own and its neighbours' -- L1 hits), E / O = even / odd pixels as halves, the shifted pairs by PRMT, row sums for the even and the odd
// pixels; the last K - 1 row sums stay in registers (ring resolved at compile time by unrolling K rows); column sums, rounding, one PRMT to
// bytes, one 32-bit store.  The next row's words are fetched before the current row is processed.  Image edges: the first / last thread of a row
// builds its words byte by byte through cv::borderInterpolate; rows outside the image are fetched from the reflected row (or are zero for
// BORDER_CONSTANT, which contributes zero here: smooth.simd.hpp:963-995).
#include <cstdlib>
#include <cstring>
#include "common.cuh"

namespace b200cv {

namespace {

constexpr int GB_SEG = 64;      // output rows per thread

// the first / last thread of a row: its 8 byte columns x0 - 2 .. x0 + 5 through the border rule -- resolved ONCE per thread (the same for every row;
// resolving them per row cost more than the whole rest of the kernel: 2 of the 30 warps of a row took 13x the time of the others)
struct GbEdge { int sx[8]; };
__device__ __noinline__ GbEdge gb_edge_columns(int x0, int W, int border)
{
    GbEdge e;
    for (int k = 0; k < 8; k++) {
        const int x = x0 - 2 + k;
        e.sx[k] = (unsigned)x < (unsigned)W ? x : border_interpolate(x, W, border);      // -1: BORDER_CONSTANT outside (contributes zero)
    }
    return e;
}

template <int K>
__device__ __forceinline__ void gb_fetch(const uchar* rp, int x0, bool interior, const GbEdge& e, unsigned& wl, unsigned& w0, unsigned& wr)
{
    if (!rp) { wl = w0 = wr = 0; return; }
    if (interior) {
        const unsigned* q = (const unsigned*)(rp + x0);
        wl = __ldg(q - 1); w0 = __ldg(q); wr = __ldg(q + 1);
        return;
    }
    unsigned b[8];
#pragma unroll
    for (int k = 0; k < 8; k++) b[k] = e.sx[k] < 0 ? 0u : (unsigned)rp[e.sx[k]];
    wl = (b[0] << 16) | (b[1] << 24);
    w0 = b[2] | (b[3] << 8) | (b[4] << 16) | (b[5] << 24);
    wr = b[6] | (b[7] << 8);
}

// row sums of the even (pixels 0, 2) and odd (pixels 1, 3) columns as 16-bit halves
template <int K>
__device__ __forceinline__ void gb_row(unsigned wl, unsigned w0, unsigned wr, unsigned& he, unsigned& ho)
{
    const unsigned E = __byte_perm(w0, 0u, 0x4240), O = __byte_perm(w0, 0u, 0x4341);       // (b0, b2), (b1, b3)
    const unsigned Om1 = __byte_perm(wl, O, 0x5453);                                          // (b-1, b1)
    const unsigned Ep1 = __byte_perm(E, wr, 0x1412);                                          // (b2, b4)
    if constexpr (K == 3) {
        he = Om1 + O + 2u * E;
        ho = E + Ep1 + 2u * O;
    } else {
        const unsigned Em1 = __byte_perm(wl, E, 0x5452);                                      // (b-2, b0)
        const unsigned Op1 = __byte_perm(O, wr, 0x1512);                                      // (b3, b5)
        he = (Em1 + Ep1) + 4u * (Om1 + O) + 6u * E;
        ho = (Om1 + Op1) + 4u * (E + Ep1) + 6u * O;
    }
}

// FAST: the segment's rows y0 - R .. y1 - 1 + R are all inside the image and the thread's three words are inside the row -- pointers advance by
// additions, no border logic in the loop; otherwise rows and bytes go through cv::borderInterpolate (edge threads, first / last segments)
template <int K, bool FAST>
__device__ __forceinline__ void gb_walk(const Img& src, const Img& dst, int border, int x0, int y0, int y1, int f, bool interior)
{
    constexpr int R = K / 2;
    const int W = src.cols, H = src.rows;
    auto row_ptr = [&](int r) -> const uchar* {
        const int sr = (unsigned)r < (unsigned)H ? r : border_interpolate(r, H, border);
        return sr < 0 ? nullptr : src.row<uchar>(f, sr);
    };
    GbEdge edge;
    if (!FAST && !interior) edge = gb_edge_columns(x0, W, border);
    const unsigned* qf = FAST ? (const unsigned*)(src.row<uchar>(f, y0 - R) + x0) : nullptr;
    const size_t stepw = src.step >> 2;
    int fi = y0 - R;                                   // next image row to fetch
    auto fetch = [&](unsigned& a, unsigned& b, unsigned& c) {
        if constexpr (FAST) { a = __ldg(qf - 1); b = __ldg(qf); c = __ldg(qf + 1); qf += stepw; }
        else gb_fetch<K>(row_ptr(fi), x0, interior, edge, a, b, c);
        fi++;
    };
    // input row i of this segment = image row y0 - R + i (N of them); its words sit in slot i % PF, its row sums in slot i % K.  The words of row
    // i + PF are requested as soon as row i's are consumed: PF rows of loads in flight per thread
    constexpr int PF = K == 3 ? 6 : 5;              // a multiple of K: both ring indices are compile-time constants in the unrolled body
    const int N = (y1 - y0) + K - 1;
    unsigned he[K], ho[K];
    unsigned wl[PF], w0[PF], wr[PF];
#pragma unroll
    for (int u = 0; u < PF; u++) {
        wl[u] = w0[u] = wr[u] = 0;
        if (u < N) fetch(wl[u], w0[u], wr[u]);
    }
    uchar* dp = dst.row<uchar>(f, y0) + x0;
    const bool store_word = FAST || x0 + 4 <= W;
    for (int base = 0; base < N; base += PF) {
#pragma unroll
        for (int u = 0; u < PF; u++) {
            const int i = base + u;
            if (i < N) {
                gb_row<K>(wl[u], w0[u], wr[u], he[u % K], ho[u % K]);
                if (i + PF < N) fetch(wl[u], w0[u], wr[u]);
                if (i >= K - 1) {                                   // output row y0 + i - (K - 1): input rows i - K + 1 .. i, oldest in slot (u + 1) % K
                    unsigned se, so, out;
                    if constexpr (K == 3) {
                        se = he[(u + 1) % K] + he[u % K] + 2u * he[(u + 2) % K] + 0x00080008u;
                        so = ho[(u + 1) % K] + ho[u % K] + 2u * ho[(u + 2) % K] + 0x00080008u;
                        out = __byte_perm(se >> 4, so >> 4, 0x6240);
                    } else {
                        se = (he[(u + 1) % K] + he[u % K]) + 4u * (he[(u + 2) % K] + he[(u + 4) % K]) + 6u * he[(u + 3) % K] + 0x00800080u;
                        so = (ho[(u + 1) % K] + ho[u % K]) + 4u * (ho[(u + 2) % K] + ho[(u + 4) % K]) + 6u * ho[(u + 3) % K] + 0x00800080u;
                        out = __byte_perm(se, so, 0x7351);
                    }
                    if (store_word) *(unsigned*)dp = out;
                    else for (int k = 0; k < 4 && x0 + k < W; k++) dp[k] = (uchar)(out >> (8 * k));
                    dp += dst.step;
                }
            }
        }
    }
}

template <int K>
__global__ void __launch_bounds__(128) gauss_u8_binomial_kernel(Img src, Img dst, int border)
{
    constexpr int R = K / 2;
    const int W = src.cols, H = src.rows;
    const int x0 = (blockIdx.x * 128 + threadIdx.x) * 4, f = blockIdx.z;
    const int y0 = blockIdx.y * GB_SEG, y1 = min(y0 + GB_SEG, H);
    if (x0 >= W) return;
    const bool interior = x0 >= 4 && x0 + 8 <= W;
    if (interior && y0 - R >= 0 && y1 + R <= H) gb_walk<K, true>(src, dst, border, x0, y0, y1, f, true);
    else gb_walk<K, false>(src, dst, border, x0, y0, y1, f, interior);
}

}  // namespace

// taps as 8.8 integers (what gauss_u8_fast receives); returns NOT_IMPLEMENTED when the case is not the binomial one
int gauss_u8_binomial(const Img& s, const Img& d, int cn, const int64_t* fx, int kw, const int64_t* fy, int kh, int border, cudaStream_t st)
{
    if (cn != 1 || kw != kh || (kw != 3 && kw != 5) || s.cols < 8 || s.rows < kw) return B200CV_NOT_IMPLEMENTED;
    static const int64_t t3[3] = {64, 128, 64}, t5[5] = {16, 64, 96, 64, 16};
    const int64_t* t = kw == 3 ? t3 : t5;
    for (int i = 0; i < kw; i++) if (fx[i] != t[i] || fy[i] != t[i]) return B200CV_NOT_IMPLEMENTED;
    if (s.rows >= 65536 * GB_SEG || s.frames >= 65536) return B200CV_NOT_IMPLEMENTED;
    const int words_ok = ((((uintptr_t)s.data | s.step | s.fstep) & 3) == 0 && (((uintptr_t)d.data | d.step | d.fstep) & 3) == 0) ? 1 : 0;
    if (!words_ok) return B200CV_NOT_IMPLEMENTED;         // unaligned rows: the tile kernel
    const dim3 grid(div_up(div_up((unsigned)s.cols, 4), 128), div_up((unsigned)s.rows, GB_SEG), (unsigned)s.frames);
    if (kw == 3) gauss_u8_binomial_kernel<3><<<grid, 128, 0, st>>>(s, d, border);
    else gauss_u8_binomial_kernel<5><<<grid, 128, 0, st>>>(s, d, border);
    B200_LAUNCH_CHECK();
    return B200CV_OK;
}

}  // namespace b200cv
