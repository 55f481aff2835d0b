// sift.cu -- SIFT scale space: initial image, Gaussian pyramid and difference-of-Gaussians pyramid.
//
// Reference (modules/features2d/src/sift.dispatch.cpp): createInitialImage :176-221 (u8 -> f32, optional 2x upsample through
// warpAffine(INTER_LINEAR | WARP_INVERSE_MAP, BORDER_REFLECT), GaussianBlur(sig_diff)), buildGaussianPyramid :224-263
// (octave base = INTER_NEAREST half-size of layer nOctaveLayers of the previous octave; layer i = GaussianBlur(layer i-1,
// Size(), sig[i])), buildDoGPyramid :266-310 (DoG[i] = G[i+1] - G[i]).  On the CPU the 66 blurs of a 4K frame run on the
// single-threaded FilterEngine.
//
// Here the pyramid of a whole batch of frames is a fixed launch sequence on one stream: every step is one of this
// library's kernels over all frames at once (grid z = frame), and each DoG level is produced by the blur kernel that
// writes its minuend (the subtrahend is the blur's own input, already in shared memory) -- no separate subtract pass and
// no re-read of either Gaussian level.
#include <cmath>
#include <vector>
#include "common.cuh"

namespace b200cv {

int gaussian_blur_impl(const b200cvMat* src, const b200cvMat* dst, int kw, int kh, double sigma1, double sigma2, int border, void* stream,
                       const b200cvMat* dog);

__global__ void __launch_bounds__(256) u8_to_f32_kernel(Img src, Img dst)
{
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
    if (x < src.cols) dst.row<float>(f, y)[x] = (float)src.row<uchar>(f, y)[x];
}

static int octave_count(int base_w, int base_h, int first_octave)
{
    // sift.dispatch.cpp:538
    return (int)lrint(std::log((double)(base_w < base_h ? base_w : base_h)) / std::log(2.) - 2) - first_octave;
}

}  // namespace b200cv

using namespace b200cv;

extern "C" int b200cv_sift_pyramid_layout(int width, int height, int n_layers, int upscale, int* n_octaves, size_t* gauss_elems,
                                          size_t* dog_elems, int* dims)
{
    B200_REQUIRE(width > 0 && height > 0 && n_layers > 0, "bad arguments");
    int bw = upscale ? width * 2 : width, bh = upscale ? height * 2 : height;
    int no = octave_count(bw, bh, upscale ? -1 : 0);
    B200_REQUIRE(no > 0, "image too small for a SIFT pyramid");
    size_t ge = 0, de = 0;
    int cw = bw, ch = bh;
    for (int o = 0; o < no; o++) {
        if (dims) { dims[2 * o] = cw; dims[2 * o + 1] = ch; }
        ge += (size_t)cw * ch * (n_layers + 3);
        de += (size_t)cw * ch * (n_layers + 2);
        cw /= 2; ch /= 2;
    }
    if (n_octaves) *n_octaves = no;
    if (gauss_elems) *gauss_elems = ge;
    if (dog_elems) *dog_elems = de;
    return B200CV_OK;
}

extern "C" int b200cv_sift_pyramid(const b200cvMat* src, int n_layers, double sigma, int upscale, float* gauss, size_t gauss_frame_elems,
                                   float* dog, size_t dog_frame_elems, void* stream)
{
    int rc;
    if ((rc = check_mat(src, "src"))) return rc;
    B200_REQUIRE(src->type == B200CV_MAKETYPE(B200CV_8U, 1), "SIFT pyramid input must be CV_8UC1");
    B200_REQUIRE(gauss && n_layers > 0 && n_layers <= 8, "bad arguments");
    const int W = src->cols, H = src->rows, frames = src->frames > 1 ? src->frames : 1;
    int no; size_t ge, de;
    std::vector<int> dims(64);
    if ((rc = b200cv_sift_pyramid_layout(W, H, n_layers, upscale, &no, &ge, &de, dims.data()))) return rc;
    B200_REQUIRE(gauss_frame_elems >= ge && (!dog || dog_frame_elems >= de), "output buffers too small");
    cudaStream_t st = as_stream(stream);
    const int F32 = B200CV_MAKETYPE(B200CV_32F, 1);
    auto level = [&](float* base, size_t frame_elems, size_t off, int w, int h) {
        b200cvMat m = {base + off, (size_t)w * 4, w, h, F32, frames, frame_elems * 4};
        return m;
    };

    // ---- createInitialImage ----
    float* tmp = nullptr;   // gray_fpt (W x H) followed by dbl (2W x 2H), per frame
    const size_t gray_elems = (size_t)W * H, dbl_elems = upscale ? gray_elems * 4 : 0;
    const size_t tmp_frame = (gray_elems + dbl_elems + 3) & ~(size_t)3;
    B200_CUDA(cudaMallocAsync(&tmp, tmp_frame * frames * sizeof(float), st));
    b200cvMat gray = level(tmp, tmp_frame, 0, W, H);
    {
        Img s = make_img(src), d = make_img(&gray);
        u8_to_f32_kernel<<<dim3(div_up((unsigned)W, 256), H, frames), 256, 0, st>>>(s, d);
        count_launch();
    }
    const float fsigma = (float)sigma;
    b200cvMat g00 = level(gauss, gauss_frame_elems, 0, dims[0], dims[1]);
    if (upscale) {
        float sig_diff = sqrtf(fmaxf(fsigma * fsigma - 0.5f * 0.5f * 4, 0.01f));
        b200cvMat dbl = level(tmp, tmp_frame, gray_elems, 2 * W, 2 * H);
        const double Mh[6] = {0.5, 0, 0, 0, 0.5, 0};
        // upscale == 2: SIFT::create's default, enable_precise_upscale = false -> cv::resize(INTER_LINEAR) (sift.dispatch.cpp:203-208)
        if (upscale == 2) rc = b200cv_resize(&gray, &dbl, B200CV_INTER_LINEAR, stream);
        else rc = b200cv_warp_affine(&gray, &dbl, Mh, B200CV_INTER_LINEAR | B200CV_WARP_INVERSE_MAP, B200CV_BORDER_REFLECT, nullptr, stream);
        if (!rc) rc = gaussian_blur_impl(&dbl, &g00, 0, 0, sig_diff, sig_diff, B200CV_BORDER_REFLECT_101, stream, nullptr);
    } else {
        float sig_diff = sqrtf(fmaxf(fsigma * fsigma - 0.5f * 0.5f, 0.01f));
        rc = gaussian_blur_impl(&gray, &g00, 0, 0, sig_diff, sig_diff, B200CV_BORDER_REFLECT_101, stream, nullptr);
    }
    if (rc) { cudaFreeAsync(tmp, st); return rc; }

    // ---- buildGaussianPyramid + buildDoGPyramid ----
    std::vector<double> sig(n_layers + 3);
    sig[0] = sigma;
    const double k = std::pow(2., 1. / n_layers);
    for (int i = 1; i < n_layers + 3; i++) {
        double sig_prev = std::pow(k, (double)(i - 1)) * sigma;
        double sig_total = sig_prev * k;
        sig[i] = std::sqrt(sig_total * sig_total - sig_prev * sig_prev);
    }
    size_t goff = 0, doff = 0;
    for (int o = 0; o < no && !rc; o++) {
        const int w = dims[2 * o], h = dims[2 * o + 1];
        const size_t n = (size_t)w * h;
        if (o > 0) {
            const int pw = dims[2 * (o - 1)], ph = dims[2 * (o - 1) + 1];
            b200cvMat prev = level(gauss, gauss_frame_elems, goff - (size_t)pw * ph * (n_layers + 3) + (size_t)pw * ph * n_layers, pw, ph);
            b200cvMat base = level(gauss, gauss_frame_elems, goff, w, h);
            rc = b200cv_resize(&prev, &base, B200CV_INTER_NEAREST, stream);
        }
        for (int i = 1; i < n_layers + 3 && !rc; i++) {
            b200cvMat a = level(gauss, gauss_frame_elems, goff + (size_t)(i - 1) * n, w, h);
            b200cvMat b = level(gauss, gauss_frame_elems, goff + (size_t)i * n, w, h);
            b200cvMat dg;
            if (dog) dg = level(dog, dog_frame_elems, doff + (size_t)(i - 1) * n, w, h);
            rc = gaussian_blur_impl(&a, &b, 0, 0, sig[i], sig[i], B200CV_BORDER_REFLECT_101, stream, dog ? &dg : nullptr);
        }
        goff += n * (n_layers + 3);
        doff += n * (n_layers + 2);
    }
    cudaFreeAsync(tmp, st);
    return rc;
}
