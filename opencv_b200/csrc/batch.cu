// batch.cu -- multi-GPU batch driver (include/b200cv_batch.h): frames of a host batch are sharded over the devices of one box, one persistent
// host thread per device, NUMA-local page-locked staging, ncclBroadcast of the one shared operand.  No arithmetic here.
// Pattern in the reference: samples/gpu/multi.cpp:27-68 (one worker per device, setDevice, independent work).
#include <dlfcn.h>
#include <pthread.h>
#include <sched.h>
#include <sys/mman.h>
#include <map>
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "common.cuh"
#include "../../include/b200cv_batch.h"
#include "../../include/b200cv_hal.h"

namespace b200cv {

int host_match_template_dev(const b200cvMat* image, const b200cvMat* dtempl, const b200cvMat* result, int method);   // hal_api.cu

// ---- NCCL, loaded at run time (the library must load on boxes without it; a one-device driver never needs it) ---------------------------
struct NcclApi {
    void* h = nullptr;
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok() const { return h && CommInitAll && CommDestroy && GroupStart && GroupEnd && Broadcast; }
};
static NcclApi& nccl()
{
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) if ((a.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!a.h) return;
        a.CommInitAll = (int (*)(void**, int, const int*))dlsym(a.h, "ncclCommInitAll");
        a.CommDestroy = (int (*)(void*))dlsym(a.h, "ncclCommDestroy");
        a.GroupStart = (int (*)())dlsym(a.h, "ncclGroupStart");
        a.GroupEnd = (int (*)())dlsym(a.h, "ncclGroupEnd");
        a.Broadcast = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(a.h, "ncclBroadcast");
        a.GetErrorString = (const char* (*)(int))dlsym(a.h, "ncclGetErrorString");
    });
    return a;
}
constexpr int NCCL_UINT8 = 1;       // ncclUint8 (nccl.h: ncclInt8 = 0, ncclUint8 = 1)

// ---- worker ---------------------------------------------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (cap >= bytes) return B200CV_OK;
        if (p) { B200_CUDA(cudaFree(p)); p = nullptr; cap = 0; }
        B200_CUDA(cudaMalloc(&p, bytes));
        cap = bytes;
        return B200CV_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct Worker {
    int device = 0, index = 0;
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, quit = false, done = false;
    int rc = 0;
    char err[512] = "";
    int last_count = 0;
    int init_rc = 0;
    // device-side state owned by the worker thread
    cudaStream_t s_up = nullptr, s_k = nullptr, s_down = nullptr;
    cudaEvent_t ev_up[2] = {}, ev_k[2] = {}, ev_down[2] = {};
    DevBuf templ, gauss, dog, src[2], har[2];
};

static void pin_thread_near_gpu(int device)
{
    char bus[32] = "";
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return; }
    for (char* c = bus; *c; c++) if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
    std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return;
    char line[1024] = "";
    const bool got = fgets(line, sizeof(line), f) != nullptr;
    fclose(f);
    if (!got) return;
    cpu_set_t set;
    CPU_ZERO(&set);
    int n = 0;
    for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        const int k = sscanf(tok, "%d-%d", &a, &b);
        if (k == 1) b = a;
        if (k >= 1) for (int c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET(c, &set); n++; }
    }
    if (n > 0) pthread_setaffinity_np(pthread_self(), sizeof(set), &set);     // refused inside a narrower cpuset: stay where we are
}

static void worker_main(Worker* w)
{
    int rc = b200cv_init(w->device);
    if (rc == B200CV_OK) {
        pin_thread_near_gpu(w->device);
        cudaError_t e = cudaStreamCreateWithFlags(&w->s_up, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&w->s_k, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&w->s_down, cudaStreamNonBlocking);
        for (int i = 0; i < 2 && e == cudaSuccess; i++) {
            e = cudaEventCreateWithFlags(&w->ev_up[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w->ev_k[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w->ev_down[i], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) rc = cuda_fail(e, "worker stream setup", __FILE__, __LINE__);
    }
    {
        std::lock_guard<std::mutex> g(w->m);
        w->init_rc = rc;
        if (rc) strncpy(w->err, b200cv_last_error(), sizeof(w->err) - 1);
        w->done = true;
    }
    w->cv.notify_all();
    while (true) {
        std::function<int()> job;
        {
            std::unique_lock<std::mutex> lk(w->m);
            w->cv.wait(lk, [&] { return w->has_job || w->quit; });
            if (w->quit) break;
            job = std::move(w->job);
            w->has_job = false;
        }
        set_error("%s", "");
        const int r = job();
        {
            std::lock_guard<std::mutex> g(w->m);
            w->rc = r;
            strncpy(w->err, b200cv_last_error(), sizeof(w->err) - 1);
            w->done = true;
        }
        w->cv.notify_all();
    }
    if (w->init_rc == B200CV_OK) {
        cudaDeviceSynchronize();
        w->templ.release(); w->gauss.release(); w->dog.release();
        for (int i = 0; i < 2; i++) { w->src[i].release(); w->har[i].release(); }
        for (int i = 0; i < 2; i++) { if (w->ev_up[i]) cudaEventDestroy(w->ev_up[i]); if (w->ev_k[i]) cudaEventDestroy(w->ev_k[i]); if (w->ev_down[i]) cudaEventDestroy(w->ev_down[i]); }
        if (w->s_up) cudaStreamDestroy(w->s_up);
        if (w->s_k) cudaStreamDestroy(w->s_k);
        if (w->s_down) cudaStreamDestroy(w->s_down);
    }
}

}  // namespace b200cv

using namespace b200cv;

struct b200cvBatch {
    std::vector<Worker*> w;
    std::vector<void*> comm;         // ncclComm_t per worker (empty: one device, or NCCL absent)
    std::mutex call;                 // one batch call at a time per handle
    std::map<void*, size_t> regions; // b200cv_batch_host_alloc_frames: base -> bytes
};

namespace b200cv {

// post fn(worker) to the selected workers and wait; first failure wins (its message moves to the caller's thread)
static int run_on(b200cvBatch* b, const std::function<int(Worker&)>& fn, int only = -1)
{
    const int n = (int)b->w.size();
    for (int i = 0; i < n; i++) {
        if (only >= 0 && i != only) continue;
        Worker* w = b->w[i];
        {
            std::lock_guard<std::mutex> g(w->m);
            w->job = [w, &fn] { return fn(*w); };
            w->has_job = true; w->done = false;
        }
        w->cv.notify_all();
    }
    int rc = B200CV_OK;
    for (int i = 0; i < n; i++) {
        if (only >= 0 && i != only) continue;
        Worker* w = b->w[i];
        std::unique_lock<std::mutex> lk(w->m);
        w->cv.wait(lk, [&] { return w->done; });
        if (w->rc != B200CV_OK && rc == B200CV_OK) { rc = w->rc; set_error("device %d: %s", w->device, w->err); }
    }
    return rc;
}

static inline b200cvMat sub_batch(const b200cvMat* m, int first, int count)
{
    b200cvMat s = *m;
    s.data = (char*)m->data + (size_t)first * m->frame_step;
    s.frames = count;
    return s;
}

static inline int frames_of(const b200cvMat* m) { return m->frames > 1 ? m->frames : 1; }

// shard a (src, dst) pair of host batches and run `op` on every worker's block
static int sharded(b200cvBatch* b, const b200cvMat* src, const b200cvMat* dst, const std::function<int(const b200cvMat*, const b200cvMat*)>& op)
{
    int rc;
    if (!b) { set_error("null batch handle"); return B200CV_ERR_BAD_ARG; }
    if ((rc = check_mat(src, "src")) || (rc = check_mat(dst, "dst"))) return rc;
    const int frames = frames_of(src);
    B200_REQUIRE(frames_of(dst) == frames, "src/dst batch mismatch");
    std::lock_guard<std::mutex> g(b->call);
    const int n = (int)b->w.size();
    return run_on(b, [&](Worker& w) {
        int f0 = 0, cnt = 0;
        b200cv_batch_shard(frames, w.index, n, &f0, &cnt);
        w.last_count = cnt;
        if (cnt == 0) return (int)B200CV_OK;
        const b200cvMat s = sub_batch(src, f0, cnt), d = sub_batch(dst, f0, cnt);
        return op(&s, &d);
    });
}

static inline size_t pitch256(size_t row_bytes) { return (row_bytes + 255) & ~(size_t)255; }

static int copy_frames(void* dst, size_t dstep, size_t dfstep, const void* src, size_t sstep, size_t sfstep, size_t row_bytes, int rows, int n, cudaMemcpyKind kind, cudaStream_t st)
{
    if (dstep == sstep && dfstep == sfstep && dfstep == dstep * (size_t)rows) {
        B200_CUDA(cudaMemcpyAsync(dst, src, dfstep * n, kind, st));
    } else if (dfstep == dstep * (size_t)rows && sfstep == sstep * (size_t)rows) {
        B200_CUDA(cudaMemcpy2DAsync(dst, dstep, src, sstep, row_bytes, (size_t)rows * n, kind, st));
    } else {
        for (int f = 0; f < n; f++)
            B200_CUDA(cudaMemcpy2DAsync((char*)dst + (size_t)f * dfstep, dstep, (const char*)src + (size_t)f * sfstep, sstep, row_bytes, rows, kind, st));
    }
    return B200CV_OK;
}

struct SiftHarrisArgs {
    const b200cvMat* src; const b200cvMat* harris;
    int nol; double sigma; int upscale, bs, ks; double k; int wave;
    b200cvWaveConsumer consumer; void* user;
    int n_workers, frames;
};

static int sift_harris_worker(Worker& w, const SiftHarrisArgs& a)
{
    int rc, f0 = 0, cnt = 0;
    b200cv_batch_shard(a.frames, w.index, a.n_workers, &f0, &cnt);
    w.last_count = cnt;
    if (cnt == 0) return B200CV_OK;
    const int W = a.src->cols, H = a.src->rows;
    int noct = 0; size_t gel = 0, del = 0;
    if ((rc = b200cv_sift_pyramid_layout(W, H, a.nol, a.upscale, &noct, &gel, &del, nullptr))) return rc;
    const int wave = std::max(1, std::min(a.wave, cnt));
    const size_t sp = pitch256((size_t)W), sfb = sp * H, hp = pitch256((size_t)W * 4), hfb = hp * H;
    if ((rc = w.gauss.ensure(gel * sizeof(float) * wave)) || (rc = w.dog.ensure(del * sizeof(float) * wave))) return rc;
    for (int i = 0; i < 2; i++)
        if ((rc = w.src[i].ensure(sfb * wave)) || (rc = w.har[i].ensure(hfb * wave))) return rc;
    const int nw = (cnt + wave - 1) / wave;
    const size_t h_fstep = a.src->frames > 1 ? a.src->frame_step : a.src->step * (size_t)H;
    for (int wi = 0; wi < nw; wi++) {
        const int b = wi & 1, fw0 = f0 + wi * wave, nf = std::min(wave, f0 + cnt - fw0);
        if (wi >= 2) B200_CUDA(cudaStreamWaitEvent(w.s_up, w.ev_k[b], 0));          // the kernels of wave wi-2 are done with this source buffer
        if ((rc = copy_frames(w.src[b].p, sp, sfb, (const char*)a.src->data + (size_t)fw0 * h_fstep, a.src->step, h_fstep, (size_t)W, H, nf, cudaMemcpyHostToDevice, w.s_up))) return rc;
        B200_CUDA(cudaEventRecord(w.ev_up[b], w.s_up));
        B200_CUDA(cudaStreamWaitEvent(w.s_k, w.ev_up[b], 0));
        if (wi >= 2 && a.harris) B200_CUDA(cudaStreamWaitEvent(w.s_k, w.ev_down[b], 0));   // the download of wave wi-2 is done with this response buffer
        const b200cvMat ds = {w.src[b].p, sp, W, H, B200CV_MAKETYPE(B200CV_8U, 1), nf, sfb};
        const b200cvMat dh = {w.har[b].p, hp, W, H, B200CV_MAKETYPE(B200CV_32F, 1), nf, hfb};
        if ((rc = b200cv_sift_pyramid(&ds, a.nol, a.sigma, a.upscale, (float*)w.gauss.p, gel, (float*)w.dog.p, del, (void*)w.s_k))) return rc;
        if ((rc = b200cv_corner_harris(&ds, &dh, a.bs, a.ks, a.k, B200CV_BORDER_REFLECT_101, (void*)w.s_k))) return rc;
        B200_CUDA(cudaEventRecord(w.ev_k[b], w.s_k));
        if (a.harris) {
            const size_t o_fstep = a.harris->frames > 1 ? a.harris->frame_step : a.harris->step * (size_t)H;
            B200_CUDA(cudaStreamWaitEvent(w.s_down, w.ev_k[b], 0));
            if ((rc = copy_frames((char*)a.harris->data + (size_t)fw0 * o_fstep, a.harris->step, o_fstep, w.har[b].p, hp, hfb, (size_t)W * 4, H, nf, cudaMemcpyDeviceToHost, w.s_down))) return rc;
            B200_CUDA(cudaEventRecord(w.ev_down[b], w.s_down));
        }
        if (a.consumer) {
            B200_CUDA(cudaStreamSynchronize(w.s_k));
            if ((rc = a.consumer(a.user, w.index, fw0, nf, (const float*)w.gauss.p, gel, (const float*)w.dog.p, del, (const float*)w.har[b].p, hp, hfb))) {
                set_error("wave consumer returned %d", rc);
                cudaStreamSynchronize(w.s_up); cudaStreamSynchronize(w.s_down);
                return rc;
            }
        }
    }
    B200_CUDA(cudaStreamSynchronize(w.s_up));
    B200_CUDA(cudaStreamSynchronize(w.s_k));
    B200_CUDA(cudaStreamSynchronize(w.s_down));
    return B200CV_OK;
}

}  // namespace b200cv

extern "C" {

int b200cv_batch_shard(int frames, int index, int n_workers, int* first, int* count)
{
    B200_REQUIRE(frames >= 0 && n_workers > 0 && index >= 0 && index < n_workers && first && count, "bad shard arguments");
    const int base = frames / n_workers, rem = frames % n_workers;
    *first = index * base + std::min(index, rem);
    *count = base + (index < rem ? 1 : 0);
    return B200CV_OK;
}

int b200cv_batch_create(b200cvBatch** out, const int* devices, int n_devices)
{
    B200_REQUIRE(out, "null out pointer");
    *out = nullptr;
    const int visible = b200cv_device_count();
    if (visible == 0) { set_error("no CUDA device visible: the b200cv hot path has no CPU fallback"); return B200CV_ERR_NO_DEVICE; }
    std::vector<int> devs;
    if (!devices || n_devices <= 0) for (int i = 0; i < visible; i++) devs.push_back(i);
    else for (int i = 0; i < n_devices; i++) {
        B200_REQUIRE(devices[i] >= 0 && devices[i] < visible, "bad device index");
        B200_REQUIRE(std::find(devs.begin(), devs.end(), devices[i]) == devs.end(), "device listed twice");
        devs.push_back(devices[i]);
    }
    b200cvBatch* b = new b200cvBatch();
    for (size_t i = 0; i < devs.size(); i++) {
        Worker* w = new Worker();
        w->device = devs[i]; w->index = (int)i;
        b->w.push_back(w);
        w->th = std::thread(worker_main, w);
    }
    int rc = B200CV_OK;
    for (Worker* w : b->w) {
        std::unique_lock<std::mutex> lk(w->m);
        w->cv.wait(lk, [&] { return w->done; });
        if (w->init_rc && !rc) { rc = w->init_rc; set_error("device %d: %s", w->device, w->err); }
    }
    if (rc == B200CV_OK && devs.size() > 1) {
        // the collective of the path: a communicator over the driver's devices (ncclCommInitAll: one process, one rank per device)
        NcclApi& n = nccl();
        if (!n.ok()) { set_error("libnccl.so.2 not found: a multi-device batch driver broadcasts its shared operand through NCCL"); rc = B200CV_ERR_NO_DEVICE; }
        else {
            b->comm.assign(devs.size(), nullptr);
            const int r = n.CommInitAll(b->comm.data(), (int)devs.size(), devs.data());
            if (r != 0) { set_error("ncclCommInitAll failed: %s", n.GetErrorString ? n.GetErrorString(r) : "?"); b->comm.clear(); rc = B200CV_ERR_CUDA; }
        }
    }
    if (rc) { b200cv_batch_destroy(b); return rc; }
    *out = b;
    return B200CV_OK;
}

int b200cv_batch_destroy(b200cvBatch* b)
{
    if (!b) return B200CV_OK;
    for (auto& r : b->regions) { cudaHostUnregister(r.first); munmap(r.first, r.second); }
    for (void* c : b->comm) if (c) nccl().CommDestroy(c);
    for (Worker* w : b->w) {
        { std::lock_guard<std::mutex> g(w->m); w->quit = true; }
        w->cv.notify_all();
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    delete b;
    return B200CV_OK;
}

int b200cv_batch_device_count(const b200cvBatch* b) { return b ? (int)b->w.size() : 0; }
int b200cv_batch_device(const b200cvBatch* b, int index) { return b && index >= 0 && index < (int)b->w.size() ? b->w[index]->device : -1; }
int b200cv_batch_uses_nccl(const b200cvBatch* b) { return b && !b->comm.empty() ? 1 : 0; }
int b200cv_batch_last_count(const b200cvBatch* b, int index) { return b && index >= 0 && index < (int)b->w.size() ? b->w[index]->last_count : 0; }

int b200cv_batch_host_alloc(b200cvBatch* b, int index, void** hptr, size_t bytes)
{
    B200_REQUIRE(b && hptr && bytes > 0 && index >= 0 && index < (int)b->w.size(), "bad host_alloc arguments");
    std::lock_guard<std::mutex> g(b->call);
    return run_on(b, [&](Worker&) {
        B200_CUDA(cudaHostAlloc(hptr, bytes, cudaHostAllocPortable));     // pinned by the worker thread: pages come from the NUMA node of its CPUs
        memset(*hptr, 0, bytes);
        return (int)B200CV_OK;
    }, index);
}

int b200cv_batch_host_free(b200cvBatch* b, int index, void* hptr)
{
    B200_REQUIRE(b && index >= 0 && index < (int)b->w.size(), "bad host_free arguments");
    std::lock_guard<std::mutex> g(b->call);
    return run_on(b, [&](Worker&) { B200_CUDA(cudaFreeHost(hptr)); return (int)B200CV_OK; }, index);
}

// One contiguous host batch whose frame blocks live on the NUMA node of the device that will process them: anonymous pages, first touched by
// the owning worker (its thread runs on the CPUs next to its GPU), then page-locked in place (cudaHostRegister keeps the physical placement).
int b200cv_batch_host_alloc_frames(b200cvBatch* b, void** hptr, size_t frame_bytes, int frames)
{
    B200_REQUIRE(b && hptr && frame_bytes > 0 && frames > 0, "bad host_alloc_frames arguments");
    const size_t total = frame_bytes * (size_t)frames;
    void* base = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (base == MAP_FAILED) { set_error("mmap of %zu bytes failed", total); return B200CV_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(b->call);
    const int n = (int)b->w.size();
    int rc = run_on(b, [&](Worker& w) {
        int f0 = 0, cnt = 0;
        b200cv_batch_shard(frames, w.index, n, &f0, &cnt);
        if (cnt > 0) memset((char*)base + (size_t)f0 * frame_bytes, 0, (size_t)cnt * frame_bytes);
        return (int)B200CV_OK;
    });
    if (rc == B200CV_OK) {
        cudaError_t e = cudaHostRegister(base, total, cudaHostRegisterPortable);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaHostRegister", __FILE__, __LINE__);
    }
    if (rc) { munmap(base, total); return rc; }
    b->regions[base] = total;
    *hptr = base;
    return B200CV_OK;
}

int b200cv_batch_host_free_frames(b200cvBatch* b, void* hptr)
{
    B200_REQUIRE(b && hptr, "bad host_free_frames arguments");
    std::lock_guard<std::mutex> g(b->call);
    auto it = b->regions.find(hptr);
    B200_REQUIRE(it != b->regions.end(), "not a b200cv_batch_host_alloc_frames region");
    cudaHostUnregister(hptr);
    munmap(hptr, it->second);
    b->regions.erase(it);
    return B200CV_OK;
}

int b200cv_batch_gaussian_blur(b200cvBatch* b, const b200cvMat* s, const b200cvMat* d, int kw, int kh, double sx, double sy, int border)
{ return sharded(b, s, d, [=](const b200cvMat* a, const b200cvMat* c) { return b200cv_host_gaussian_blur(a, c, kw, kh, sx, sy, border); }); }
int b200cv_batch_sep_filter2d(b200cvBatch* b, const b200cvMat* s, const b200cvMat* d, const float* kx, int nx, const float* ky, int ny, int ax, int ay, double delta, int border)
{ return sharded(b, s, d, [=](const b200cvMat* a, const b200cvMat* c) { return b200cv_host_sep_filter2d(a, c, kx, nx, ky, ny, ax, ay, delta, border); }); }
int b200cv_batch_filter2d(b200cvBatch* b, const b200cvMat* s, const b200cvMat* d, const float* k, int kw, int kh, int ax, int ay, double delta, int border)
{ return sharded(b, s, d, [=](const b200cvMat* a, const b200cvMat* c) { return b200cv_host_filter2d(a, c, k, kw, kh, ax, ay, delta, border); }); }
int b200cv_batch_resize(b200cvBatch* b, const b200cvMat* s, const b200cvMat* d, int interp, double fx, double fy)
{ return sharded(b, s, d, [=](const b200cvMat* a, const b200cvMat* c) { return b200cv_host_resize_scaled(a, c, interp, fx, fy); }); }
int b200cv_batch_warp_affine(b200cvBatch* b, const b200cvMat* s, const b200cvMat* d, const double* M, int flags, int border, const double* bv)
{ return sharded(b, s, d, [=](const b200cvMat* a, const b200cvMat* c) { return b200cv_host_warp_affine(a, c, M, flags, border, bv); }); }
int b200cv_batch_warp_perspective(b200cvBatch* b, const b200cvMat* s, const b200cvMat* d, const double* M, int flags, int border, const double* bv)
{ return sharded(b, s, d, [=](const b200cvMat* a, const b200cvMat* c) { return b200cv_host_warp_perspective(a, c, M, flags, border, bv); }); }
int b200cv_batch_cvt_color(b200cvBatch* b, const b200cvMat* s, const b200cvMat* d, int code)
{ return sharded(b, s, d, [=](const b200cvMat* a, const b200cvMat* c) { return b200cv_host_cvt_color(a, c, code); }); }
int b200cv_batch_corner_harris(b200cvBatch* b, const b200cvMat* s, const b200cvMat* d, int bs, int ks, double k, int border)
{ return sharded(b, s, d, [=](const b200cvMat* a, const b200cvMat* c) { return b200cv_host_corner_harris(a, c, bs, ks, k, border); }); }

int b200cv_batch_match_template(b200cvBatch* b, const b200cvMat* image, const b200cvMat* templ, const b200cvMat* result, int method)
{
    int rc;
    if (!b) { set_error("null batch handle"); return B200CV_ERR_BAD_ARG; }
    if ((rc = check_mat(templ, "templ"))) return rc;
    const size_t tp = pitch256((size_t)templ->cols * elem_size(templ->type)), tbytes = tp * templ->rows;
    {
        // shared operand: on the first device from the host, to the others by ncclBroadcast (NVLink), each on its worker's kernel stream
        std::lock_guard<std::mutex> g(b->call);
        rc = run_on(b, [&](Worker& w) {
            int r = w.templ.ensure(tbytes);
            if (r) return r;
            if (w.index == 0) {
                B200_CUDA(cudaMemsetAsync(w.templ.p, 0, tbytes, w.s_k));
                B200_CUDA(cudaMemcpy2DAsync(w.templ.p, tp, templ->data, templ->step, (size_t)templ->cols * elem_size(templ->type), templ->rows, cudaMemcpyHostToDevice, w.s_k));
                B200_CUDA(cudaStreamSynchronize(w.s_k));
            }
            return (int)B200CV_OK;
        });
        if (rc) return rc;
        if (!b->comm.empty()) {
            NcclApi& n = nccl();
            int r = n.GroupStart();
            for (size_t i = 0; i < b->w.size() && r == 0; i++)
                r = n.Broadcast(b->w[i]->templ.p, b->w[i]->templ.p, tbytes, NCCL_UINT8, 0, b->comm[i], b->w[i]->s_k);
            const int r2 = n.GroupEnd();
            if (r == 0) r = r2;
            if (r != 0) { set_error("ncclBroadcast failed: %s", n.GetErrorString ? n.GetErrorString(r) : "?"); return B200CV_ERR_CUDA; }
            for (Worker* w : b->w) B200_CUDA(cudaStreamSynchronize(w->s_k));
        }
    }
    return sharded(b, image, result, [=](const b200cvMat* a, const b200cvMat* c) {
        int dev = 0;
        cudaGetDevice(&dev);
        Worker* me = nullptr;
        for (Worker* w : b->w) if (w->device == dev) me = w;
        if (!me) { set_error("worker/device mismatch"); return (int)B200CV_ERR_BAD_ARG; }
        const b200cvMat dt = {me->templ.p, tp, templ->cols, templ->rows, templ->type, 1, 0};
        return host_match_template_dev(a, &dt, c, method);
    });
}

int b200cv_batch_sift_harris(b200cvBatch* b, const b200cvMat* src, const b200cvMat* harris, int nol, double sigma, int upscale, int bs, int ks, double k, int wave,
                             b200cvWaveConsumer consumer, void* user)
{
    int rc;
    if (!b) { set_error("null batch handle"); return B200CV_ERR_BAD_ARG; }
    if ((rc = check_mat(src, "src"))) return rc;
    B200_REQUIRE(src->type == B200CV_MAKETYPE(B200CV_8U, 1), "SIFT pyramid: 8UC1 frames");
    if (harris && harris->data) {
        if ((rc = check_mat(harris, "harris"))) return rc;
        B200_REQUIRE(harris->type == B200CV_MAKETYPE(B200CV_32F, 1) && harris->cols == src->cols && harris->rows == src->rows && frames_of(harris) == frames_of(src),
                     "harris: CV_32FC1 batch of the source's size");
    } else harris = nullptr;
    B200_REQUIRE(wave >= 1, "wave >= 1");
    SiftHarrisArgs a = {src, harris, nol, sigma, upscale, bs, ks, k, wave, consumer, user, (int)b->w.size(), frames_of(src)};
    std::lock_guard<std::mutex> g(b->call);
    return run_on(b, [&](Worker& w) { return sift_harris_worker(w, a); });
}

}  // extern "C"
