// matchtemplate_tc.cu -- the matchTemplate correlation numerator on 5th-generation tensor cores (tcgen05, kind::i8).
//
//   R(y, x) = sum_{v<h} sum_{u<w} T(v,u) * I(y+v, x+u)          8-bit image and template, exact 32-bit integer result
//
// is a dense contraction with N = 1 per window.  To give the MMA a useful N the template row is expanded into a banded
// (Toeplitz) operand: for an x-tile of 64 outputs and one template row v
//   D[m][j] += sum_{k<128} A_v[m][k] * B_v[k][j],   A_v[m][k] = I(y0+m+v, x0+k),   B_v[k][j] = T(v, k-j) (0 outside [0,w))
// A_v is a plain 128-byte-wide box of image rows -- NO im2col is materialised: the CTA stages rows y0 .. y0+M+h-2 once
// (TMA, 16-byte-column boxes = the canonical K-major no-swizzle core-matrix layout) and the A descriptor of step v simply
// starts v rows (v*16 bytes) further down.  B_v (8 KB) is pre-expanded once per call into global memory in the same layout
// and streamed through a 4-stage cp.async.bulk ring.  One elected thread issues tcgen05.mma (M=128, N=64, K=32, u8 x u8 ->
// s32, accumulators in TMEM, two M-tiles per CTA share every B_v), tcgen05.commit frees ring slots and publishes the
// accumulator; all four warps read it back with tcgen05.ld and store float(R).  Half of B is zeros (2.1x redundant MACs),
// the price of a well-shaped MMA; the result is exact because products <= 65025 and sums < 2^31.
//
//
// (8-bit filter2D with >= 11x11 taps uses the same Toeplitz formulation in a persistent kernel: filter2d_tc.cu.)
//
// Reference: crossCorr, modules/imgproc/src/templmatch.cpp:566-760 (block DFT in float on one thread).
#include "common.cuh"
#include "tma.cuh"
#include <cmath>
#include <vector>

namespace b200cv {

constexpr int TC_N = 64;          // outputs per x-tile (MMA N)
constexpr int TC_K = 128;         // K per template row: w + N - 1 <= 128
constexpr int TC_MT = 2;          // M-tiles (of 128 rows) per CTA
constexpr int TC_NS = 4;          // B ring stages
constexpr int TC_BBYTES = TC_N * TC_K;   // 8192


// B_v in smem/global: [k-chunk c (8)][column j (64)][16 bytes]: byte b = T(v, 16c + b - j)
__global__ void toeplitz_kernel(Img templ, int w, int h, unsigned char* out)
{
    const int v = blockIdx.x;
    for (int idx = threadIdx.x; idx < TC_BBYTES; idx += blockDim.x) {
        int c = idx >> 10, j = (idx >> 4) & 63, b = idx & 15;
        int u = 16 * c + b - j;
        out[(size_t)v * TC_BBYTES + idx] = (u >= 0 && u < w) ? templ.row<uchar>(0, v)[u] : (uchar)0;
    }
}

// border-extended copy of an 8-bit single-channel image: P(y, x) = src(y - ay, x - ax) under `border`.
// 16 output bytes per thread; away from the borders the (generally unaligned) source run is read as 5 aligned words and realigned
// with funnel shifts.
__global__ void __launch_bounds__(128) pad_u8_kernel(Img src, Img dst, int ax, int ay, int border, int words_ok)
{
    const int f = blockIdx.z, y = blockIdx.y;
    const int x16 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (x16 >= dst.cols) return;
    const int sy = border_interpolate(y - ay, src.rows, border);
    uint32_t o[4] = {0, 0, 0, 0};
    if (sy >= 0) {
        const uchar* sp = src.row<uchar>(f, sy);
        const int sx0 = x16 - ax;
        if (words_ok && sx0 >= 0 && sx0 + 20 <= src.cols) {
            const int a = sx0 & 3;
            const uint32_t* wp = (const uint32_t*)(sp + (sx0 - a));
            uint32_t w[5];
#pragma unroll
            for (int i = 0; i < 5; i++) w[i] = wp[i];
#pragma unroll
            for (int i = 0; i < 4; i++) o[i] = __funnelshift_r(w[i], w[i + 1], 8 * a);
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                int sx = border_interpolate(sx0 + i, src.cols, border);
                o[i >> 2] |= (uint32_t)(sx >= 0 ? sp[sx] : (uchar)0) << (8 * (i & 3));
            }
        }
    }
    *(uint4*)(dst.row<uchar>(f, y) + x16) = make_uint4(o[0], o[1], o[2], o[3]);
}

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    // K-major, no swizzle: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout_type=0 [61,64)
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// one lane of a converged warp: the issuer warp runs its loop warp-uniformly and elects the issuing thread, so the descriptors stay in
// uniform registers (issued from inside `if (lane == 0)` every operand went through an R2UR waterfall: ~15 instructions per MMA)
__device__ __forceinline__ bool umma_elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
          "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
          "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}

struct TCParams {
    int h, ra_alloc, box_h, nbox;       // template rows; staged image rows (allocated), rows per TMA box, boxes per column chunk
    int ow, oh;
    int kch;                            // K per template row = 32 * kch (w + N - 1 <= 32 * kch)
};

__global__ void __launch_bounds__(128) ccorr_u8_tc_kernel(const __grid_constant__ CUtensorMap tmap, const unsigned char* __restrict__ bglob, Img res, TCParams p)
{
    constexpr int NN = TC_N;                                           // MMA N
    constexpr int TCOLS = TC_MT * NN <= 128 ? 128 : TC_MT * NN <= 256 ? 256 : 512;   // TMEM columns (power of two)
    extern __shared__ __align__(128) unsigned char smem[];
    const int nchunk = 2 * p.kch;                                      // 16-byte K chunks per row
    const uint32_t bbytes = (uint32_t)nchunk * NN * 16;                // one B slab
    unsigned char* sA = smem;                                          // nchunk x ra_alloc rows x 16 B
    unsigned char* sB = smem + (size_t)nchunk * p.ra_alloc * 16;       // TC_NS slabs
    __shared__ __align__(8) uint64_t full[TC_NS], empty[TC_NS], a_full, acc_full;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.z, x0 = blockIdx.x * TC_N, y0 = blockIdx.y * (128 * TC_MT);
    const uint32_t lbo_a = (uint32_t)p.ra_alloc * 16u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_NS; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(&a_full, 1); mbar_init(&acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) {   // TMEM: TC_MT x NN columns of 32-bit accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    if (warp == 0 && lane == 0) {
        // ---- producer: image rows once, then the Toeplitz operand of every template row through the ring ----
        mbar_arrive_expect_tx(&a_full, (uint32_t)p.ra_alloc * 16u * nchunk);
        for (int c = 0; c < nchunk; c++)
            for (int b = 0; b < p.nbox; b++)
                tma_load_3d(sA + (size_t)c * lbo_a + (size_t)b * p.box_h * 16, &tmap, x0 + 16 * c, y0 + b * p.box_h, f, &a_full);
        for (int v = 0; v < p.h; v++) {
            const int s = v % TC_NS;
            mbar_wait(&empty[s], ((v / TC_NS) & 1) ^ 1);
            mbar_arrive_expect_tx(&full[s], bbytes);
            bulk_load(sB + (size_t)s * bbytes, bglob + (size_t)v * bbytes, bbytes, &full[s]);
        }
    } else if (warp == 1) {
        // ---- MMA issuer: the whole warp runs the loop (uniform values), one elected lane issues ----
        // instruction descriptor: D = S32 (2<<4), A = B = unsigned 8 bit (0 at [7,10) and [10,13)),
        // K-major both, N>>3 at [17,23), M>>4 at [24,29)
        const uint32_t idesc = (2u << 4) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        mbar_wait(&a_full, 0);
        tc_fence_after();
        // descriptors (K-major, no swizzle): low word = start >> 4 [0,14) | LBO >> 4 [16,30); high word = SBO >> 4 | version 1 at bit 14
        const uint32_t hi = (128u >> 4) | (1u << 14);
        const uint32_t a_lo0 = ((smem_u32(sA) & 0x3FFFFu) >> 4) | ((lbo_a >> 4) << 16);
        const uint32_t b_lo0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (((uint32_t)(NN * 16) >> 4) << 16);
        const uint32_t a_ks = (2u * lbo_a) >> 4;
        for (int v = 0; v < p.h; v++) {
            const int s = v % TC_NS;
            mbar_wait(&full[s], (v / TC_NS) & 1);
            tc_fence_after();
            if (umma_elect_one()) {
#pragma unroll
                for (int mt = 0; mt < TC_MT; mt++)
#pragma unroll
                    for (int ks = 0; ks < TC_K / 32; ks++) {
                        if (ks < p.kch) {
                            const uint64_t ad = ((uint64_t)hi << 32) | (a_lo0 + (uint32_t)ks * a_ks + (uint32_t)(mt * 128 + v));
                            const uint64_t bd = ((uint64_t)hi << 32) | (b_lo0 + (((uint32_t)s * bbytes) >> 4) + (uint32_t)ks * ((2u * NN * 16) >> 4));
                            umma_i8(tmem + mt * NN, ad, bd, idesc, (v | ks) != 0);
                        }
                    }
                umma_commit(&empty[s]);
            }
            __syncwarp();
        }
        if (umma_elect_one()) umma_commit(&acc_full);
    }
    // ---- epilogue: all four warps; warp w owns TMEM lanes (= accumulator rows) 32w .. 32w+31 ----
    __syncwarp();
    mbar_wait(&acc_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int mt = 0; mt < TC_MT; mt++) {
        const int gy = y0 + mt * 128 + warp * 32 + lane;
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16) + mt * NN;
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            uint32_t r[32];
            tmem_ld32(trow + half * 32, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float* rp = gy < p.oh ? res.row<float>(f, gy) + x0 : nullptr;
            if (rp) {
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    int gx = x0 + half * 32 + j;
                    if (gx < p.ow) rp[half * 32 + j] = (float)(int)r[j];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCOLS) : "memory");
}

// returns B200CV_NOT_IMPLEMENTED when the tensor-core path does not apply (caller uses the IDP4A kernel)
int ccorr_u8_tensor(const Img& im, const Img& tp, const Img& rs, int w, int h, cudaStream_t st)
{
    if (w > TC_K - TC_N + 1 || h > 512 || !tma_compatible(im) || im.frames >= 65536) return B200CV_NOT_IMPLEMENTED;
    const int ow = im.cols - w + 1, oh = im.rows - h + 1;
    TCParams p;
    p.h = h; p.ow = ow; p.oh = oh; p.kch = TC_K / 32;
    int ra = 128 * TC_MT + h - 1;
    p.nbox = (ra + 255) / 256;
    p.box_h = (((ra + p.nbox - 1) / p.nbox) + 7) & ~7;
    p.ra_alloc = p.nbox * p.box_h;
    size_t smem = (size_t)8 * p.ra_alloc * 16 + (size_t)TC_NS * TC_BBYTES;
    if (smem > 200 * 1024) return B200CV_NOT_IMPLEMENTED;
    static PerDeviceFlag attr_pd; bool& attr = attr_pd.cur();
    if (!attr) { B200_CUDA(cudaFuncSetAttribute(ccorr_u8_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
    unsigned char* bglob = nullptr;
    B200_CUDA(cudaMallocAsync(&bglob, (size_t)h * TC_BBYTES, st));
    toeplitz_kernel<<<h, 256, 0, st>>>(tp, w, h, bglob);
    count_launch();
    CUtensorMap tm;
    int rc = make_tensor_map_3d(&tm, im.data, 1, im.cols, im.rows, im.frames, im.step, im.fstep, 16, p.box_h);
    if (rc) { cudaFreeAsync(bglob, st); return rc; }
    dim3 grid(div_up((unsigned)ow, TC_N), div_up((unsigned)oh, 128 * TC_MT), (unsigned)im.frames);
    ccorr_u8_tc_kernel<<<grid, 128, smem, st>>>(tm, bglob, rs, p);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(bglob, st);
    count_launch();
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch", __FILE__, __LINE__);
    return B200CV_OK;
}

void launch_pad_u8(const Img& src, const Img& dst, int ax, int ay, int border, cudaStream_t st)
{
    pad_u8_kernel<<<dim3(div_up((unsigned)dst.cols / 16, 128), (unsigned)dst.rows, (unsigned)src.frames), 128, 0, st>>>(
        src, dst, ax, ay, border, (((uintptr_t)src.data | src.step | src.fstep) & 3) == 0);
    count_launch();
}

}  // namespace b200cv
