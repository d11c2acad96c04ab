// C-ABI plumbing: version / error reporting / launch counter, and the host-buffer entry points.
#include <atomic>
#include <stdlib.h>
#include <string.h>

#include "mac_common.h"

namespace mac {

namespace {
thread_local char g_error[512] = "";
std::atomic<unsigned long long> g_launches{0};

struct HostCache {  // per-thread device staging for the *_host entry points
    int device = -1;
    void *buf = nullptr;
    size_t bytes = 0;
    cudaStream_t stream = nullptr;       // compute
    cudaStream_t copy_stream = nullptr;  // host -> device slices
    cudaStream_t copy_stream2 = nullptr; // optional second copy queue (MAC_HOST_COPY_STREAMS=2)
    cudaEvent_t landed[64] = {};
};
thread_local HostCache g_cache;
}  // namespace

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
void uncount_launch(unsigned n) { g_launches.fetch_sub(n, std::memory_order_relaxed); }

int sm_count(int device)
{
    static int cached[64] = {0};
    if (device < 0 || device >= 64) return 148;
    if (cached[device] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
        cached[device] = n;
    }
    return cached[device];
}

}  // namespace mac

extern "C" int mac_version(void) { return 100; /* 0.1.0 */ }

extern "C" const char *mac_last_error(void) { return mac::g_error; }

extern "C" int mac_built_for_sm(void) { return 100; }

extern "C" unsigned long long mac_launch_count(void) { return mac::g_launches.load(std::memory_order_relaxed); }

extern "C" void mac_host_release(void)
{
    mac::HostCache &c = mac::g_cache;
    if (c.buf) {
        cudaSetDevice(c.device);
        cudaFree(c.buf);
    }
    if (c.stream) cudaStreamDestroy(c.stream);
    if (c.copy_stream) cudaStreamDestroy(c.copy_stream);
    if (c.copy_stream2) cudaStreamDestroy(c.copy_stream2);
    for (cudaEvent_t e : c.landed)
        if (e) cudaEventDestroy(e);
    c = mac::HostCache();
}

extern "C" int mac_covgain_host(const float *pts, int pts_dim, const float *harmonics, const float *cams, float *out,
                                int B, int P, int C, int cam_begin, int cam_end, int act, int device)
{
    using namespace mac;
    MAC_REQUIRE(pts && harmonics && cams && out, "null host pointer");
    MAC_REQUIRE(B > 0 && P > 0 && C > 0 && pts_dim >= 3, "bad shape B=%d P=%d C=%d pts_dim=%d", B, P, C, pts_dim);
    MAC_REQUIRE(0 <= cam_begin && cam_begin <= cam_end && cam_end <= C, "bad camera range [%d, %d) for C=%d",
                cam_begin, cam_end, C);
    // the call works on `device` and leaves the caller's current device as it found it
    struct DeviceGuard {
        int prev = -1;
        ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    } guard;
    MAC_CUDA(cudaGetDevice(&guard.prev));
    MAC_CUDA(cudaSetDevice(device));
    HostCache &c = g_cache;
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t n_pts = up(sizeof(float) * B * static_cast<size_t>(P) * pts_dim);
    const size_t n_harm = up(sizeof(float) * B * static_cast<size_t>(P) * MAC_N_HARMONICS);
    const size_t n_cams = up(sizeof(float) * B * static_cast<size_t>(C) * 3);
    const size_t n_out = up(sizeof(float) * B * static_cast<size_t>(C));
    const size_t n_ws = up(mac_covgain_workspace_bytes(B, C));
    const size_t need = n_pts + n_harm + n_cams + n_out + n_ws;
    if (c.device != device || c.bytes < need) {
        mac_host_release();
        MAC_CUDA(cudaSetDevice(device));
        MAC_CUDA(cudaMalloc(&c.buf, need));
        MAC_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        MAC_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
        MAC_CUDA(cudaStreamCreateWithFlags(&c.copy_stream2, cudaStreamNonBlocking));
        for (cudaEvent_t &e : c.landed) MAC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c.device = device;
        c.bytes = need;
    }
    unsigned char *base = static_cast<unsigned char *>(c.buf);
    float *d_pts = reinterpret_cast<float *>(base);
    float *d_harm = reinterpret_cast<float *>(base + n_pts);
    float *d_cams = reinterpret_cast<float *>(base + n_pts + n_harm);
    float *d_out = reinterpret_cast<float *>(base + n_pts + n_harm + n_cams);
    void *d_ws = base + n_pts + n_harm + n_cams + n_out;
    MAC_CUDA(cudaMemsetAsync(d_ws, 0, n_ws, c.stream));
    // The (small) camera upload goes FIRST into the queue that carries the point slices.  On the compute stream it would be
    // ordered behind the memset; by the time it became ready the host had usually queued all point slices on the copy
    // engine, which serves one direction in submission order: the cameras -- and with them every kernel -- then waited
    // for the whole 54.6 MB (measured with MAC_HOST_TRACE: first kernel done at 1.14 ms instead of 0.21 ms, a race that
    // made a call take 1.15 or 1.54 ms).
    const bool sliced_upload = (B == 1 && P >= (1 << 16));
    MAC_CUDA(cudaMemcpyAsync(d_cams, cams, sizeof(float) * B * static_cast<size_t>(C) * 3, cudaMemcpyHostToDevice,
                             sliced_upload ? c.copy_stream : c.stream));
    // One cloud with many points: the points arrive in slices on a copy stream while the kernel integrates the
    // previous slice (the partial sums stay in the fixed-point workspace between the slice launches).
    // MAC_HOST_SLICES / MAC_HOST_SKIP_KERNEL: tuning / diagnosis knobs of tools/bench_e2e.py (read once)
    static const int env_slices = [] { const char *e = getenv("MAC_HOST_SLICES"); return e ? atoi(e) : 0; }();
    static const bool skip_kernel = [] { const char *e = getenv("MAC_HOST_SKIP_KERNEL"); return e && atoi(e) != 0; }();
    int n_slices = sliced_upload ? 12 : 1;   // 8, 12 and 16 measure within 3 % of each other (1.19-1.26 ms at a 1.00 ms link floor)
    if (env_slices >= 1 && env_slices <= 64 && sliced_upload) n_slices = env_slices;
    if (sliced_upload && n_slices == 1) {   // single piece: everything on the compute stream, order the cameras there too
        MAC_CUDA(cudaEventRecord(c.landed[0], c.copy_stream));
        MAC_CUDA(cudaStreamWaitEvent(c.stream, c.landed[0], 0));
    }
    if (n_slices > 1) {
        const int per = ((P + n_slices - 1) / n_slices + 31) / 32 * 32;
        static const int copy_streams = [] { const char *e = getenv("MAC_HOST_COPY_STREAMS"); return e ? atoi(e) : 1; }();
        // MAC_HOST_TRACE=1: device timeline of every call on stderr (copy k landed / kernel k done, ms since the first copy)
        static const bool trace = [] { const char *e = getenv("MAC_HOST_TRACE"); return e && atoi(e) != 0; }();
        static thread_local cudaEvent_t tr_start = nullptr, tr_landed[64], tr_kdone[64];
        if (trace && !tr_start) {
            MAC_CUDA(cudaEventCreate(&tr_start));
            for (int i = 0; i < 64; ++i) {
                MAC_CUDA(cudaEventCreate(&tr_landed[i]));
                MAC_CUDA(cudaEventCreate(&tr_kdone[i]));
            }
        }
        if (trace) MAC_CUDA(cudaEventRecord(tr_start, c.copy_stream));
        int k = 0;
        for (int p0 = 0; p0 < P; p0 += per, ++k) {
            const int np = P - p0 < per ? P - p0 : per;
            cudaStream_t cs = (copy_streams == 2 && (k & 1)) ? c.copy_stream2 : c.copy_stream;
            MAC_CUDA(cudaMemcpyAsync(d_pts + static_cast<size_t>(p0) * pts_dim, pts + static_cast<size_t>(p0) * pts_dim,
                                     sizeof(float) * np * static_cast<size_t>(pts_dim), cudaMemcpyHostToDevice, cs));
            MAC_CUDA(cudaMemcpyAsync(d_harm + static_cast<size_t>(p0) * MAC_N_HARMONICS,
                                     harmonics + static_cast<size_t>(p0) * MAC_N_HARMONICS,
                                     sizeof(float) * np * static_cast<size_t>(MAC_N_HARMONICS), cudaMemcpyHostToDevice,
                                     cs));
            MAC_CUDA(cudaEventRecord(c.landed[k], cs));
            if (trace) MAC_CUDA(cudaEventRecord(tr_landed[k], cs));
            MAC_CUDA(cudaStreamWaitEvent(c.stream, c.landed[k], 0));
            if (skip_kernel) continue;
            const int rc = covgain_accumulate(d_pts + static_cast<size_t>(p0) * pts_dim, pts_dim,
                                              d_harm + static_cast<size_t>(p0) * MAC_N_HARMONICS, d_cams, d_out, np, C, cam_begin,
                                              cam_end, act, d_ws, n_ws, P, p0 + np >= P ? 1 : 0, c.stream);
            if (rc != MAC_OK) return rc;
            if (trace) MAC_CUDA(cudaEventRecord(tr_kdone[k], c.stream));
        }
        if (trace) {
            MAC_CUDA(cudaStreamSynchronize(c.stream));
            char line[2048];
            int n = 0;
            for (int i = 0; i < k; ++i) {
                float a = 0.f, b = 0.f;
                cudaEventElapsedTime(&a, tr_start, tr_landed[i]);
                cudaEventElapsedTime(&b, tr_start, tr_kdone[i]);
                n += snprintf(line + n, sizeof(line) - n, " %.3f/%.3f", a, b);
            }
            fprintf(stderr, "[mac_covgain_host] landed/kernel-done ms:%s\n", line);
        }
    } else {
        MAC_CUDA(cudaMemcpyAsync(d_pts, pts, sizeof(float) * B * static_cast<size_t>(P) * pts_dim, cudaMemcpyHostToDevice,
                                 c.stream));
        MAC_CUDA(cudaMemcpyAsync(d_harm, harmonics, sizeof(float) * B * static_cast<size_t>(P) * MAC_N_HARMONICS,
                                 cudaMemcpyHostToDevice, c.stream));
        const int rc = mac_covgain_f32(d_pts, pts_dim, d_harm, d_cams, d_out, B, P, C, cam_begin, cam_end, act, d_ws, n_ws,
                                       c.stream);
        if (rc != MAC_OK) return rc;
    }
    if (cam_end > cam_begin) {
        MAC_CUDA(cudaMemcpy2DAsync(out + cam_begin, sizeof(float) * C, d_out + cam_begin, sizeof(float) * C,
                                   sizeof(float) * (cam_end - cam_begin), B, cudaMemcpyDeviceToHost, c.stream));
    }
    MAC_CUDA(cudaStreamSynchronize(c.stream));
    return MAC_OK;
}
