// host_tables.h -- exact host-side coefficient tables (see host_tables.cpp)
#pragma once
#include <cstdint>
#include <vector>

namespace b200cv {

double softdouble_exp(double x);
void gaussian_kernel_bitexact(int n, double sigma, std::vector<double>& out);
void gaussian_kernel_fixed(int n, double sigma, int bits, std::vector<int64_t>& out);
int gaussian_auto_ksize(double sigma, bool is_u8);

// cv::remap interpolation tables (imgwarp.cpp:213-287): 32x32 sub-pixel positions x (ksize*ksize) taps, float and 2^15 fixed point
void bilinear_tab(std::vector<float>& f, std::vector<short>& i);
void bicubic_tab(std::vector<float>& f, std::vector<short>& i);
void lanczos4_tab(std::vector<float>& f, std::vector<short>& i);

}  // namespace b200cv
