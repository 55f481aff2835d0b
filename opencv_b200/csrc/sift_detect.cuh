// sift_detect.cuh -- types shared by the SIFT front-end kernels (sift_detect.cu, sift_desc_warp.cu)
#pragma once
#include "common.cuh"

namespace b200cv {

enum { SIFT_MAX_OCT = 16, SIFT_BORDER = 5, SIFT_ORI_BINS = 36 };

struct SiftPyr {
    const float* gauss;
    const float* dog;
    int n_oct, nl;
    int w[SIFT_MAX_OCT], h[SIFT_MAX_OCT];
    unsigned long long goff[SIFT_MAX_OCT], doff[SIFT_MAX_OCT];      // element offsets of octave o in the packed buffers
};

struct SiftCand { int o, layer, r, c; };
struct SiftKp { float x, y, size, angle, response; int octave; };

__device__ __forceinline__ float fast_atan2_deg(float y, float x)            // atan_f32, mathfuncs_core.simd.hpp:52-72
{
    const float p1 = 0.9997878412794807f * (float)(180 / 3.1415926535897932384626433832795), p3 = -0.3258083974640975f * (float)(180 / 3.1415926535897932384626433832795);
    const float p5 = 0.1555786518463281f * (float)(180 / 3.1415926535897932384626433832795), p7 = -0.04432655554792128f * (float)(180 / 3.1415926535897932384626433832795);
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) { c = ay / (ax + (float)2.2204460492503131e-16); c2 = c * c; a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
    else { c = ax / (ay + (float)2.2204460492503131e-16); c2 = c * c; a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}


}  // namespace b200cv
