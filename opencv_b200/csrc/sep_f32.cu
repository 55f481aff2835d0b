// sep_f32.cu -- single-channel float separable filter fast path (GaussianBlur f32, sepFilter2D f32, the SIFT pyramid blurs):
// TMA tile load + FFMA row/column passes, optional fused difference-of-Gaussians output.
//
// Same arithmetic as sep_fast_kernel<float,float,M_FLOAT,KB> in sepfilter.cu (row pass s = fma(src, kx[i], s) in tap order,
// column pass starting from delta; reference: RowVec_32f / SymmColumnFilter, modules/imgproc/src/filter.simd.hpp:1634-1648,
// 2652-2757) -- only the data movement differs: ONE thread issues a 3-D cp.async.bulk.tensor box load of the
// (192+2r) x (32+2r) float tile into shared memory and the CTA waits on an mbarrier, so the loads of one CTA overlap the
// arithmetic of the other CTAs resident on the SM instead of stalling every thread on its own LDG (the generic kernel was
// long-scoreboard bound: ncu profiles/r01_prof1_summary.txt).  TMA zero-fills outside the image (= BORDER_CONSTANT); for
// REPLICATE / REFLECT / REFLECT_101 only boundary CTAs patch their apron from the mirrored in-tile cells.
#include <cstring>
#include "common.cuh"
#include "tma.cuh"

namespace b200cv {

constexpr int SF_TW = 192, SF_TH = 32, SF_IW = 224;

struct SF32Params {
    float kx[32], ky[32];      // zero padded + centred to KB taps
    float delta;
    int W, H, border;
    Img dog;
    int has_dog;
};

template <int KB>
__global__ void __launch_bounds__(256, 2) sep_f32_tma_kernel(const CUtensorMap* __restrict__ tmap, Img dst, const __grid_constant__ SF32Params p)
{
    constexpr int RB = KB / 2;
    constexpr int RA = ((RB + 3) / 4) * 4;               // left apron staged: TMA needs the box to start on a 16-byte boundary
    constexpr int OFF = RA - RB;
    constexpr int IH = SF_TH + KB - 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_in = (float*)smem_raw;                       // IH x SF_IW
    float* s_mid = s_in + IH * SF_IW;                     // IH x SF_TW
    __shared__ __align__(8) uint64_t s_bar;
    const int f = blockIdx.z, x0 = blockIdx.x * SF_TW, y0 = blockIdx.y * SF_TH;
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_barrier_init();
        mbar_arrive_expect_tx(&s_bar, (uint32_t)(SF_IW * IH * sizeof(float)));
        tma_load_3d(s_in, tmap, x0 - RA, y0 - RB, f, &s_bar);
    }
    __syncthreads();
    mbar_wait(&s_bar, 0);

    const int tx0 = x0 - RA;
    const bool edge = (tx0 < 0) || (y0 - RB < 0) || (tx0 + SF_IW > p.W) || (y0 - RB + IH > p.H);
    if (edge && p.border != B200CV_BORDER_CONSTANT) {
        for (int idx = tid; idx < IH * SF_IW; idx += 256) {
            int r = idx / SF_IW, c = idx - r * SF_IW;
            int gy = y0 - RB + r;
            if ((unsigned)gy < (unsigned)p.H) continue;
            int sr = border_interpolate(gy, p.H, p.border) - (y0 - RB);
            if ((unsigned)sr < (unsigned)IH) s_in[idx] = s_in[sr * SF_IW + c];   // rows beyond the apron feed no valid output
        }
        __syncthreads();
        const int c_first = p.W - tx0;                        // first tile column right of the image (may be >= SF_IW)
        const int nright = c_first < SF_IW ? min(SF_IW - c_first, RB + 4) : 0;
        const int nleft = tx0 < 0 ? RA : 0;
        const int ncol = nleft + nright;
        for (int idx = tid; idx < IH * ncol; idx += 256) {
            int r = idx / ncol, k = idx - r * ncol;
            int c = k < nleft ? k : c_first + (k - nleft);
            int sc = border_interpolate(tx0 + c, p.W, p.border) - tx0;
            if ((unsigned)sc < (unsigned)SF_IW) s_in[r * SF_IW + c] = s_in[r * SF_IW + sc];
        }
        __syncthreads();
    }

    // ---- row pass: item = 8 consecutive outputs of one staged row ----
    {
        constexpr int GPR = SF_TW / 8;                    // 24 items per row
        constexpr int NEED = 8 + KB - 1;
        constexpr int NV = (OFF + NEED + 3) / 4;
#pragma unroll 1
        for (int it = tid; it < IH * GPR; it += 256) {
            const int r = it / GPR, g = it - r * GPR;
            const float4* vp = (const float4*)(s_in + r * SF_IW + g * 8);
            float acc[8];
#pragma unroll
            for (int o = 0; o < 8; o++) acc[o] = 0.f;
#pragma unroll
            for (int w = 0; w < NV; w++) {
                float4 q = vp[w];
                const float vals[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int e = w * 4 + b - OFF;
                    if (e >= 0 && e < NEED) {
#pragma unroll
                        for (int o = 0; o < 8; o++) {
                            const int i = e - o;
                            if (i >= 0 && i < KB) acc[o] = fmaf(vals[b], p.kx[i], acc[o]);
                        }
                    }
                }
            }
            float4* mp = (float4*)(s_mid + r * SF_TW + g * 8);
            mp[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
            mp[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
    }
    __syncthreads();

    // ---- column pass: item = 4 columns x 8 rows ----
    {
        const bool dvec = (((uintptr_t)dst.data | dst.step | dst.fstep) & 15) == 0;
        const bool gvec = p.has_dog && (((uintptr_t)p.dog.data | p.dog.step | p.dog.fstep) & 15) == 0;
#pragma unroll 1
        for (int it = tid; it < (SF_TW / 4) * (SF_TH / 8); it += 256) {
            const int q = it / (SF_TW / 4), c4 = it - q * (SF_TW / 4);
            const float* mbase = s_mid + (q * 8) * SF_TW + c4 * 4;
            float acc[8][4];
#pragma unroll
            for (int o = 0; o < 8; o++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[o][c] = p.delta;
#pragma unroll
            for (int m = 0; m < 8 + KB - 1; m++) {
                const float4 v = *(const float4*)(mbase + m * SF_TW);
#pragma unroll
                for (int o = 0; o < 8; o++) {
                    const int j = m - o;
                    if (j >= 0 && j < KB) {
                        const float t = p.ky[j];
                        acc[o][0] = fmaf(v.x, t, acc[o][0]); acc[o][1] = fmaf(v.y, t, acc[o][1]);
                        acc[o][2] = fmaf(v.z, t, acc[o][2]); acc[o][3] = fmaf(v.w, t, acc[o][3]);
                    }
                }
            }
            const int gx = x0 + c4 * 4;
            if (gx >= p.W) continue;
#pragma unroll
            for (int o = 0; o < 8; o++) {
                const int gy = y0 + q * 8 + o;
                if (gy >= p.H) break;
                float* dp = dst.row<float>(f, gy) + gx;
                if (dvec && gx + 4 <= p.W) *(float4*)dp = make_float4(acc[o][0], acc[o][1], acc[o][2], acc[o][3]);
                else {
#pragma unroll
                    for (int c = 0; c < 4; c++) if (gx + c < p.W) dp[c] = acc[o][c];
                }
                if (p.has_dog) {
                    const float* ctr = s_in + (q * 8 + o + RB) * SF_IW + RA + c4 * 4;
                    float* gp = p.dog.row<float>(f, gy) + gx;
                    if (gvec && gx + 4 <= p.W)
                        *(float4*)gp = make_float4(__fsub_rn(acc[o][0], ctr[0]), __fsub_rn(acc[o][1], ctr[1]), __fsub_rn(acc[o][2], ctr[2]), __fsub_rn(acc[o][3], ctr[3]));
                    else {
#pragma unroll
                        for (int c = 0; c < 4; c++) if (gx + c < p.W) gp[c] = __fsub_rn(acc[o][c], ctr[c]);
                    }
                }
            }
        }
    }
}

template <int KB>
static int launch_sf32(const CUtensorMap& tm, const Img& d, const SF32Params& p, int frames, cudaStream_t st)
{
    constexpr int IH = SF_TH + KB - 1;
    const size_t smem = (size_t)IH * (SF_IW + SF_TW) * sizeof(float);
    auto kern = sep_f32_tma_kernel<KB>;
    static bool attr = false;
    if (!attr) { B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    CUtensorMap* dtm = nullptr;
    int rc = upload_tensor_map(tm, &dtm, st);
    if (rc) return rc;
    dim3 grid(div_up((unsigned)p.W, SF_TW), div_up((unsigned)p.H, SF_TH), (unsigned)frames);
    kern<<<grid, 256, smem, st>>>(dtm, d, p);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(dtm, st);
    count_launch();
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch", __FILE__, __LINE__);
    return B200CV_OK;
}

// returns B200CV_NOT_IMPLEMENTED when the fast path does not apply
int sep_f32_fast(const Img& s, const Img& d, const float* kx, int nx, const float* ky, int ny, float delta, int border, const Img* dog, cudaStream_t st)
{
    if (!(nx & 1) || !(ny & 1) || nx > 31 || ny > 31) return B200CV_NOT_IMPLEMENTED;
    if (border == B200CV_BORDER_WRAP || !tma_compatible(s) || s.frames >= 65536) return B200CV_NOT_IMPLEMENTED;
    static const int buckets[] = {3, 5, 7, 9, 11, 13, 15, 17, 21, 25, 27, 31};
    int kmax = nx > ny ? nx : ny, KB = 0;
    for (int b : buckets) if (kmax <= b) { KB = b; break; }
    if (!KB || s.cols < KB || s.rows < KB) return B200CV_NOT_IMPLEMENTED;
    SF32Params p;
    memset(&p, 0, sizeof(p));
    for (int i = 0; i < nx; i++) p.kx[(KB - nx) / 2 + i] = kx[i];
    for (int i = 0; i < ny; i++) p.ky[(KB - ny) / 2 + i] = ky[i];
    p.delta = delta; p.W = s.cols; p.H = s.rows; p.border = border;
    if (dog) { p.dog = *dog; p.has_dog = 1; }
    CUtensorMap tm;
    int rc = make_tensor_map_3d(&tm, s.data, 4, s.cols, s.rows, s.frames, s.step, s.fstep, SF_IW, SF_TH + KB - 1);
    if (rc) return rc;
    switch (KB) {
    case 3: return launch_sf32<3>(tm, d, p, s.frames, st);
    case 5: return launch_sf32<5>(tm, d, p, s.frames, st);
    case 7: return launch_sf32<7>(tm, d, p, s.frames, st);
    case 9: return launch_sf32<9>(tm, d, p, s.frames, st);
    case 11: return launch_sf32<11>(tm, d, p, s.frames, st);
    case 13: return launch_sf32<13>(tm, d, p, s.frames, st);
    case 15: return launch_sf32<15>(tm, d, p, s.frames, st);
    case 17: return launch_sf32<17>(tm, d, p, s.frames, st);
    case 21: return launch_sf32<21>(tm, d, p, s.frames, st);
    case 25: return launch_sf32<25>(tm, d, p, s.frames, st);
    case 27: return launch_sf32<27>(tm, d, p, s.frames, st);
    case 31: return launch_sf32<31>(tm, d, p, s.frames, st);
    }
    return B200CV_NOT_IMPLEMENTED;
}

}  // namespace b200cv
