// C++ smoke test of the host-side mirror (opencv_b200/host/b200cv.hpp): compiled on the CPU box by tests/test_abi.py, run on the GPU
// by tests/test_gpu_hal.py.  Exit code 0 = the filter object, stream, events and free functions all work and agree with each other.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../opencv_b200/host/b200cv.hpp"
using namespace b200cv;
int main()
{
    try {
        check(b200cv_init(0), "init");
        const int W = 640, H = 480;
        std::vector<unsigned char> img((size_t)W * H * 3), out1(img.size()), out2(img.size()), gray((size_t)W * H);
        for (size_t i = 0; i < img.size(); i++) img[i] = (unsigned char)((i * 2654435761u) >> 24);
        Stream stream;
        Event e0, e1;
        GpuMat d_src(H, W, makeType(CV_8U, 3)), d_dst, d_dst2, d_gray;
        d_src.upload(img.data(), (size_t)W * 3, stream);
        e0.record(stream);
        FilterPtr gauss = createGaussianFilter(d_src.type(), d_src.type(), Size(5, 5), 0);      // the cv::cuda::Filter usage pattern
        gauss->apply(d_src, d_dst, stream);
        GaussianBlur(d_src, d_dst2, Size(5, 5), 0, 0, BORDER_DEFAULT, stream);
        cvtColor(d_src, d_gray, 6 /*COLOR_BGR2GRAY*/, 1, stream);
        e1.record(stream);
        d_dst.download(out1.data(), (size_t)W * 3, stream);
        d_dst2.download(out2.data(), (size_t)W * 3, stream);
        d_gray.download(gray.data(), (size_t)W, stream);
        stream.waitForCompletion();
        if (!stream.queryIfComplete()) return 2;
        if (memcmp(out1.data(), out2.data(), out1.size()) != 0) return 3;
        // interior check against the closed form (16,64,96,64,16)/256 squared, and the gray formula
        int x = 100, y = 50;
        for (int c = 0; c < 3; c++) {
            const int k[5] = {16, 64, 96, 64, 16}; long acc = 0;
            for (int j = 0; j < 5; j++) for (int i = 0; i < 5; i++) acc += (long)k[j] * k[i] * img[((size_t)(y + j - 2) * W + (x + i - 2)) * 3 + c];
            if (out1[((size_t)y * W + x) * 3 + c] != (unsigned char)((acc + 32768) >> 16)) return 4;
        }
        const unsigned char* p = &img[((size_t)y * W + x) * 3];
        if (gray[(size_t)y * W + x] != (unsigned char)((p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + 16384) >> 15)) return 5;
        printf("host api ok: %.3f ms for 2 blurs + 1 cvtColor on 640x480x3\n", Event::elapsedTime(e0, e1));
        return 0;
    } catch (const std::exception& e) {
        fprintf(stderr, "exception: %s\n", e.what());
        return 1;
    }
}
