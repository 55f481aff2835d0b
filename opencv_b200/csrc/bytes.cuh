// bytes.cuh -- N consecutive bytes of a row into / out of registers: one 16- / 8- / 4-byte access when the address allows it and all N bytes
// are inside the row, single bytes otherwise (row tails, odd pitches).  Shared by the streaming per-pixel kernels (cvtcolor_yuv.cu,
// cvtcolor_lab.cu); plain global-memory accesses only, so the kernels also run under tests/emu.
#pragma once
#include "common.cuh"

namespace b200cv {

__device__ __forceinline__ bool aligned_to(const void* p, unsigned a) { return ((uintptr_t)p & (a - 1)) == 0; }

// N bytes from p (only the first n are inside the row; the rest read as 0)
template <int N>
__device__ __forceinline__ void load_bytes(const uchar* p, int n, uchar (&o)[N])
{
    static_assert(N == 4 || N == 8 || N == 16 || N == 24 || N == 32, "load_bytes");
    if (n == N && aligned_to(p, N % 16 == 0 ? 16 : N % 8 == 0 ? 8 : 4)) {
        if constexpr (N % 16 == 0) {
#pragma unroll
            for (int k = 0; k < N / 16; k++) {
                const uint4 v = ((const uint4*)p)[k];
                const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 16; i++) o[16 * k + i] = (uchar)(w[i >> 2] >> (8 * (i & 3)));
            }
        } else if constexpr (N % 8 == 0) {
#pragma unroll
            for (int k = 0; k < N / 8; k++) {
                const uint2 v = ((const uint2*)p)[k];
#pragma unroll
                for (int i = 0; i < 8; i++) o[8 * k + i] = (uchar)((i < 4 ? v.x : v.y) >> (8 * (i & 3)));
            }
        } else {
            const unsigned v = *(const unsigned*)p;
#pragma unroll
            for (int i = 0; i < 4; i++) o[i] = (uchar)(v >> (8 * i));
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; i++) o[i] = i < n ? p[i] : (uchar)0;
    }
}

template <int N>
__device__ __forceinline__ void store_bytes(uchar* p, int n, const uchar (&o)[N])
{
    static_assert(N == 4 || N == 8 || N == 16 || N == 24 || N == 32, "store_bytes");
    if (n == N && aligned_to(p, N % 8 == 0 ? 8 : 4)) {
        if constexpr (N % 8 == 0) {
#pragma unroll
            for (int k = 0; k < N / 8; k++) {
                uint2 v;
                v.x = o[8 * k] | (o[8 * k + 1] << 8) | (o[8 * k + 2] << 16) | ((unsigned)o[8 * k + 3] << 24);
                v.y = o[8 * k + 4] | (o[8 * k + 5] << 8) | (o[8 * k + 6] << 16) | ((unsigned)o[8 * k + 7] << 24);
                ((uint2*)p)[k] = v;
            }
        } else {
            *(unsigned*)p = o[0] | (o[1] << 8) | (o[2] << 16) | ((unsigned)o[3] << 24);
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; i++) if (i < n) p[i] = o[i];
    }
}

}  // namespace b200cv
