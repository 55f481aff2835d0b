// demosaic.cu -- cv::cvtColor / cv::demosaicing for the Bayer patterns, bilinear interpolation, 8-bit (the wire
// format of a raw sensor).  Codes COLOR_BayerBG/GB/RG/GR2BGR = 46..49 (the 2RGB names are the same numbers permuted) and 2BGRA = 139..142.
//
// Reference (demosaicing.cpp:806-1056, Bayer2RGB_Invoker / Bayer2RGB_): for every interior pixel of the mosaic
//   at a red / blue site:  that colour = the sample, green = (4 edge neighbours + 2) >> 2, the other colour = (4 diagonal neighbours + 2) >> 2
//   at a green site:       one colour = (left + right + 1) >> 1, the other = (up + down + 1) >> 1, green = the sample
// the first / last columns repeat their interior neighbour, then the first / last rows repeat theirs (:989-1008, :1041-1055), so every
// border pixel equals the value computed at the nearest interior site -- which is how the kernel evaluates it.  All integer: bit-exact.
// `blue` (+1 / -1: which of the two colours the non-green sites of the first interior row carry) and `start_with_green` alternate row by row.
// Eight destination pixels per thread (one 24- / 32-byte store when aligned), 3 x 3 byte neighbourhoods through L1: the bound is HBM
// (1 + 3 bytes per pixel).
#include "common.cuh"
#include "bytes.cuh"

namespace b200cv {

namespace {

template <int DCN>
__global__ void __launch_bounds__(256) bayer_bilinear_kernel(Img src, Img dst, int W, int H, int blue0, int swg0)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;          // 8 destination pixels per thread, one 24- / 32-byte row piece out
    const int y = blockIdx.y, f = blockIdx.z;
    if (x0 >= W) return;
    const int n = min(8, W - x0);
    const int yi = min(max(y, 1), H - 2);                                 // border pixels repeat the nearest interior site
    const int i = yi - 1;                                                 // interior row, 0-based
    const int blue = (i & 1) ? -blue0 : blue0;
    const bool swg = ((i & 1) != 0) != (swg0 != 0);                       // this interior row starts with a green site
    const uchar* q0 = src.row<uchar>(f, yi - 1);
    const uchar* q1 = src.row<uchar>(f, yi);
    const uchar* q2 = src.row<uchar>(f, yi + 1);
    uchar o[8 * DCN];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int xi = min(max(min(x0 + j, W - 1), 1), W - 2), k = xi - 1;  // interior column, 0-based (columns past the row end are computed but not stored)
        const bool green = ((k & 1) == 0) == swg;
        const uchar *r0 = q0 + xi, *r1 = q1 + xi, *r2 = q2 + xi;
        int c_lo, c_hi, g;                                                // channel 1 - blue, channel 1 + blue, green
        if (green) {
            c_lo = (r0[0] + r2[0] + 1) >> 1;                              // vertical neighbours
            c_hi = (r1[-1] + r1[1] + 1) >> 1;                             // horizontal neighbours
            g = r1[0];
        } else {
            c_lo = (r0[-1] + r0[1] + r2[-1] + r2[1] + 2) >> 2;            // diagonals
            g = (r0[0] + r1[-1] + r1[1] + r2[0] + 2) >> 2;                // edge neighbours
            c_hi = r1[0];
        }
        o[j * DCN] = (uchar)(blue > 0 ? c_lo : c_hi);
        o[j * DCN + 1] = (uchar)g;
        o[j * DCN + 2] = (uchar)(blue > 0 ? c_hi : c_lo);
        if constexpr (DCN == 4) o[j * DCN + 3] = 255;
    }
    store_bytes<8 * DCN>(dst.row<uchar>(f, y) + (size_t)x0 * DCN, n * DCN, o);
}

}  // namespace

// called by b200cv_cvt_color for codes 46-49 and 139-142 (8-bit matrices of equal size and batch already checked)
int demosaic_bilinear(const b200cvMat* src, const b200cvMat* dst, int code, cudaStream_t st)
{
    const int scn = B200CV_CN(src->type), dcn = B200CV_CN(dst->type);
    const bool four = code >= 139;
    B200_REQUIRE(scn == 1 && dcn == (four ? 4 : 3), "Bayer -> BGR needs a 1-channel mosaic and a 3-channel (BGRA codes: 4-channel) destination");
    const int W = src->cols, H = src->rows;
    if (W < 3 || H < 3) return B200CV_NOT_IMPLEMENTED;                  // the reference zero-fills such images (demosaicing.cpp:836-851, :1051-1055)
    Img s = make_img(src), d = make_img(dst);
    if (H >= 65536 || s.frames >= 65536) return B200CV_NOT_IMPLEMENTED;
    const int c = four ? code - 139 : code - 46;                        // 0 BG, 1 GB, 2 RG, 3 GR
    const int blue0 = c < 2 ? -1 : 1, swg0 = (c & 1);
    const dim3 block(256);
    const dim3 grid(div_up(div_up((unsigned)W, 8), 256), (unsigned)H, (unsigned)s.frames);
    if (dcn == 3) bayer_bilinear_kernel<3><<<grid, block, 0, st>>>(s, d, W, H, blue0, swg0);
    else bayer_bilinear_kernel<4><<<grid, block, 0, st>>>(s, d, W, H, blue0, swg0);
    B200_LAUNCH_CHECK();
    return B200CV_OK;
}

}  // namespace b200cv
