// host_tables.h -- exact host-side coefficient tables (see host_tables.cpp)
#pragma once
#include <cstdint>
#include <vector>

namespace b200cv {

double softdouble_exp(double x);
void gaussian_kernel_bitexact(int n, double sigma, std::vector<double>& out);
void gaussian_kernel_fixed(int n, double sigma, int bits, std::vector<int64_t>& out);
int gaussian_auto_ksize(double sigma, bool is_u8);

// cv::resize coefficient tables (resize.cpp:4097-4190): per destination column/row source index + taps
struct ResizeTab {
    std::vector<int> ofs;        // first source index (already multiplied by cn for x tables when cn>0)
    std::vector<short> ialpha;   // ksize fixed-point taps per entry (u8 path)
    std::vector<float> alpha;    // ksize float taps per entry (f32 path)
    int dmax;                    // first destination index whose taps reach past the last source sample (xmax)
    int dmin;                    // last+1 destination index whose taps start before sample 0 (xmin)
};
void resize_linear_tab(int ssize, int dsize, ResizeTab& t);
void resize_cubic_tab(int ssize, int dsize, ResizeTab& t);

// cv::remap interpolation tables (imgwarp.cpp:213-287): 32x32 sub-pixel positions
void bicubic_tab_i16(std::vector<short>& tab /*1024*16*/);
void bicubic_tab_f32(std::vector<float>& tab /*1024*16*/);
void bilinear_tab_f32(std::vector<float>& tab /*1024*4*/);

}  // namespace b200cv
