// minimal TMA box-load check: build with nvcc, run on the GPU box (development aid)
#include <cstdio>
#include <cstring>
#include <vector>
#include "../opencv_b200/csrc/tma.cuh"
using namespace b200cv;
namespace b200cv { void set_error(const char* fmt, ...) {} int cuda_fail(cudaError_t e, const char* w, const char*, int) { printf("cuda fail %s: %s\n", w, cudaGetErrorString(e)); return -3; } void count_launch(int) {} }

__global__ void kg(const CUtensorMap* tmp, unsigned char* out, int x, int y, int IH, int BW)
{
    __shared__ __align__(128) unsigned char s[256 * 64];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        mbar_arrive_expect_tx(&bar, BW * IH);
        tma_load_3d(s, tmp, x, y, 0, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < BW * IH; i += blockDim.x) out[i] = s[i];
}

__global__ void k(const __grid_constant__ CUtensorMap tm, unsigned char* out, int x, int y, int IH)
{
    __shared__ __align__(128) unsigned char s[256 * 64];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        mbar_arrive_expect_tx(&bar, 256 * IH);
        tma_load_3d(s, &tm, x, y, 0, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < 256 * IH; i += blockDim.x) out[i] = s[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv)
{
    const bool runA = argc > 1;
    const int W = 3840, H = 2160, IH = 62;
    std::vector<unsigned char> h((size_t)W * H);
    for (size_t i = 0; i < h.size(); i++) h[i] = (unsigned char)((i % W) ^ (i / W));
    unsigned char *d, *o;
    cudaMalloc(&d, h.size()); cudaMalloc(&o, 256 * 64);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: %s q=%d p=%p\n", cudaGetErrorString(e), (int)q, p);
    CUtensorMap tm; memset(&tm, 0, sizeof(tm));
    for (int rank = 3; rank >= 3; rank--) {
        cuuint64_t dims[3] = {W, H, 1}; cuuint64_t strides[2] = {W, (cuuint64_t)W * H}; cuuint32_t box[3] = {256, IH, 1}; cuuint32_t es[3] = {1, 1, 1};
        CUresult r = ((EncodeTiledFn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("rank %d encode result %d\n", rank, (int)r);
        if (rank == 3 && runA) {
            k<<<1, 256>>>(tm, o, 16, 16, IH);
            e = cudaDeviceSynchronize();
            printf("rank3 kernel: %s\n", cudaGetErrorString(e));
            if (e == cudaSuccess) {
                std::vector<unsigned char> ho(256 * 64); cudaMemcpy(ho.data(), o, ho.size(), cudaMemcpyDeviceToHost);
                int bad = 0;
                for (int r2 = 0; r2 < IH; r2++) for (int c = 0; c < 256; c++) {
                    int gx = c + 16, gy = r2 + 16; unsigned char want = (gx < 0 || gy < 0) ? 0 : (unsigned char)(gx ^ gy);
                    if (ho[r2 * 256 + c] != want) bad++;
                }
                printf("mismatches %d\n", bad);
            }
        }
    }
    // variant B: tensor map in global memory, several box widths
    for (int bw : {64, 128, 256}) {
        cuuint64_t dims[3] = {W, H, 1}; cuuint64_t strides[2] = {W, (cuuint64_t)W * H}; cuuint32_t box[3] = {(cuuint32_t)bw, IH, 1}; cuuint32_t es[3] = {1, 1, 1};
        CUresult r = ((EncodeTiledFn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUtensorMap* dtm; cudaMalloc(&dtm, sizeof(tm)); cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
        for (int xy : {16, 0, 3800, -1}) {
            kg<<<1, 256>>>(dtm, o, xy, xy < 2100 ? xy : 2150, IH, bw);
            e = cudaDeviceSynchronize();
            printf("global-map bw=%d xy=%d encode %d kernel: %s\n", bw, xy, (int)r, cudaGetErrorString(e));
            if (e != cudaSuccess) return 1;
        }
    }
    return 0;
}
