// Shared host-side plumbing of the C ABI: thread-local error text, CUDA error mapping, launch counter.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/macarons_b200.h"

namespace mac {

void set_error(const char *fmt, ...);
void count_launch(unsigned n = 1);
void uncount_launch(unsigned n);   // launches recorded into a CUDA graph are counted when the graph runs
int sm_count(int device);
// covgain.cu: accumulate one slice of the points of a single cloud (see mac_covgain_host)
int covgain_accumulate(const float *pts, int pts_dim, const float *harmonics, const float *cams, float *out, int P, int C,
                       int cam_begin, int cam_end, int act, void *workspace, size_t workspace_bytes, int p_total,
                       int finalize, cudaStream_t stream);

#define MAC_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            ::mac::set_error(__VA_ARGS__);     \
            return MAC_ERR_INVALID_ARGUMENT;   \
        }                                      \
    } while (0)

#define MAC_CUDA(expr)                                                                       \
    do {                                                                                     \
        cudaError_t err__ = (expr);                                                          \
        if (err__ != cudaSuccess) {                                                          \
            ::mac::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__),      \
                             __FILE__, __LINE__);                                            \
            return MAC_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: every launcher keeps one
// DeviceOnce per kernel instantiation and configures the kernel the first time it is launched on each device.
// Two threads racing on the first call both set the same value, which is harmless.
struct DeviceOnce {
    std::atomic<unsigned char> done[64];
};
template <class Kernel>
inline int ensure_dynamic_smem(DeviceOnce &once, Kernel kernel, int bytes)
{
    int device = 0;
    MAC_CUDA(cudaGetDevice(&device));
    const bool tracked = device >= 0 && device < 64;
    if (!tracked || !once.done[device].load(std::memory_order_acquire)) {
        MAC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        if (tracked) once.done[device].store(1, std::memory_order_release);
    }
    return MAC_OK;
}

// ---- small PTX wrappers shared by the kernels (sm_100a) ---------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// order earlier generic-proxy accesses to shared memory before later async-proxy (bulk copy) writes
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP); 16-B aligned, size % 16 == 0
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// TMA tiled load of one box of a 3-D tensor map (SASS UTMALDG); coordinates fastest dimension first
__device__ __forceinline__ void tma_load_3d(void *dst_smem, const void *tensor_map, int c0, int c1, int c2, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst_smem)),
        "l"(tensor_map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
// system-scope release store / acquire load: flags that live in (possibly peer) device memory
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float rsqrt_approx(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#endif  // __CUDACC__

}  // namespace mac
