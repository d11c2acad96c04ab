// tcgen05 / TMEM / TMA wrappers shared by the tensor-core kernels (sm_100a only).
// Bit layouts follow the PTX ISA "tcgen05 matrix descriptor / instruction descriptor" tables.
#pragma once
#include <cuda.h>

#include "mac_common.h"

namespace mac {

// 2-D fp32 tensor map, row-major (rows, cols) with row stride `ld` floats, box = box_rows x 32 floats (128 B),
// SWIZZLE_128B.  Out-of-range rows / columns are zero-filled by the TMA unit.
int make_tensor_map_2d(CUtensorMap *map, const float *base, int rows, int cols, int ld, int box_rows);

#ifdef __CUDACC__
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst_smem, const void *tensor_map, int c0, int c1, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst_smem)),
        "l"(tensor_map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
// TMA tiled store of one box shared -> global (bulk async-group completion); rows / columns outside the tensor are dropped
__device__ __forceinline__ void tma_store_2d(const void *tensor_map, const void *src_smem, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tensor_map), "r"(c0), "r"(c1),
                 "r"(smem_u32(src_smem))
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk stores committed by this thread have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const void *tensor_map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(tensor_map) : "memory");
}

// ---- TMEM ------------------------------------------------------------------------------------
// one full warp; writes the TMEM base address (lane 0, column c) to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 consecutive fp32 columns of this thread's TMEM lane (warp w of the CTA owns lanes 32*(w%4) .. +31)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// write 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
        "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
        "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
        "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// 16-byte async copy global -> shared (LDGSTS), L2 only
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
// 16-byte async copy with zero fill: src_bytes = 16 copies, 0 writes zeros (nothing is read from src)
__device__ __forceinline__ void cp_async16_zfill(void *dst_smem, const void *src, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one arrival when all cp.async issued so far by this thread have landed (the arrival counts against
// the barrier's expected count: .noinc)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- UMMA (tcgen05.mma) ------------------------------------------------------------------------
// Shared-memory matrix descriptor of a K-major operand tile stored as 128-byte rows (32 fp32) with the
// 128-byte swizzle TMA writes: 8-row groups are 1024 B apart (SBO), the leading offset is unused.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);  // start address, 16-B units       [0,14)
    d |= static_cast<uint64_t>(1) << 16;                      // leading byte offset (ignored)   [16,30)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;              // stride byte offset: 8 rows      [32,46)
    d |= static_cast<uint64_t>(1) << 46;                      // descriptor version (Blackwell)  [46,48)
    d |= static_cast<uint64_t>(2) << 61;                      // SWIZZLE_128B                    [61,64)
    return d;
}
// Instruction descriptor: D fp32, A and B tf32, both K-major, M x N tile.
__device__ __host__ constexpr uint32_t umma_idesc_tf32(int m, int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(m >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread for the CTA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T: the A operand (M rows on the TMEM lanes, one tf32 per 32-bit column, 8 columns per
// k-step) is read from tensor memory, e.g. softmax weights written there with tcgen05.st
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive when all previously issued MMAs of this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// fp32 -> tf32 (round to nearest, ties away), result has the 13 low mantissa bits cleared
__device__ __forceinline__ float to_tf32(float x)
{
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
// Split of an fp32 value into two TF32 halves with full-rate integer / FADD instructions (cvt.rna.tf32.f32 is emulated
// by ~5 instructions on sm_100a): hi = x with the 13 low mantissa bits cleared (NaN / inf preserved), lo = x - hi exactly
// (|lo| < 2^-10 |x|), rounded to nearest TF32 by adding half an ulp to the bit pattern.  x = hi + lo to 2^-22 relative.
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ float tf32_lo(float x, float hi)
{
    return __uint_as_float((__float_as_uint(x - hi) + 0x1000u) & 0xffffe000u);
}
#endif  // __CUDACC__

}  // namespace mac

namespace mac {
// linear.cu: see mac_linear_f32 in include/macarons_b200.h
int linear_forward(const float *X, int ldx, const float *W_hi, const float *W_lo, int ldw, const float *bias, float *out,
                   int ldo, int M, int N, int K, int act, const float *res, int ldr, float *ln_out, int ldl,
                   const float *ln_g, const float *ln_b, float ln_eps, int pool, cudaStream_t stream, int res_first = 0,
                   const float *lnin_stats = nullptr, const float *lnin_g = nullptr, const float *lnin_b = nullptr,
                   float *stats_out = nullptr, int act_in = 0 /* MAC_LIN_NONE; MAC_LIN_GELU: exact GELU applied to X on load */);
}  // namespace mac
