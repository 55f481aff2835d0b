// cuda_emu.h -- just enough of the CUDA programming model to run SIMPLE kernels (no shared memory, no warp intrinsics, no barriers) on
// the host, one thread at a time.  TEST INFRASTRUCTURE: tests/test_kernel_emulation.py compiles a kernel source with g++ against this
// header (B200CV_HOST_EMULATION), replaces its <<<grid, block, 0, st>>> launches by EMU_LAUNCH and compares the result with the oracle.
// It checks index arithmetic, plane layouts, tails and the aligned / unaligned access paths on machines without a GPU; it says nothing
// about performance and is no substitute for the -m gpu parity tests.
#pragma once
#include <math.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ static
#define __launch_bounds__(...)
#define __grid_constant__
#define __restrict__

struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
static dim3 blockIdx, threadIdx, blockDim, gridDim;

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { *p = malloc(n); return *p ? cudaSuccess : 2; }
static inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMemcpyToSymbol(T& sym, const void* src, size_t n) { memcpy(&sym, src, n); return cudaSuccess; }
static inline int atomicAdd(int* p, int v) { const int old = *p; *p = old + v; return old; }      // threads run one after another

template <typename T> static inline T min(T a, T b) { return b < a ? b : a; }
template <typename T> static inline T max(T a, T b) { return a < b ? b : a; }
static inline int __float2int_rn(float v) { return (int)lrintf(v); }
// single operations rounded on their own (volatile: the host compiler must not contract or re-associate them either)
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __int2float_rn(int a) { return (float)a; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __int_as_float(int a) { float f; memcpy(&f, &a, 4); return f; }
static inline int __float_as_int(float f) { int a; memcpy(&a, &f, 4); return a; }
static inline short sat_s16_emu(int v) { return (short)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline float __double2float_rn(double a) { return (float)a; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline int __double2int_rn(double a) { return (int)lrint(a); }

// packed-integer intrinsics (PTX dp4a.u32.u32, dp2a.lo/hi.u32.u32, prmt) for kernels whose arithmetic core is host-testable
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c)
{
    for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xffu) * ((b >> (8 * i)) & 0xffu);
    return c;
}
static inline unsigned __dp2a_lo(unsigned a, unsigned b, unsigned c) { return c + (a & 0xffffu) * (b & 0xffu) + (a >> 16) * ((b >> 8) & 0xffu); }
static inline unsigned __dp2a_hi(unsigned a, unsigned b, unsigned c) { return c + (a & 0xffffu) * ((b >> 16) & 0xffu) + (a >> 16) * ((b >> 24) & 0xffu); }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s)
{
    const unsigned long long v = ((unsigned long long)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xffu) << (8 * i);
    return r;
}

template <typename F>
static inline void emu_launch(dim3 grid, dim3 block, F thread_body)
{
    gridDim = grid; blockDim = block;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++)
                for (unsigned tz = 0; tz < block.z; tz++)
                    for (unsigned ty = 0; ty < block.y; ty++)
                        for (unsigned tx = 0; tx < block.x; tx++) {
                            blockIdx = dim3(bx, by, bz); threadIdx = dim3(tx, ty, tz);
                            thread_body();
                        }
}
#define EMU_LAUNCH(grid, block, ...) emu_launch(grid, block, [&] { __VA_ARGS__; })
