// Dense multi-head self-attention on the 5th-generation tensor cores (reference networks/Attention.py:8-36, 182-204,
// mask = None): softmax(Q K^T / sqrt(d)) V for clouds of S tokens, 4 heads, per-head dims (8, 32) [SconeOcc global
// transformer] or (16, 64) [SconeVis].
//
// Two kernels:
//  1. attn_prep: reads the fused qkv rows once and writes the tensor-core operands, already split into TF32 halves
//     (x_hi = tf32(x), x_lo = tf32(x - x_hi), see linear.cu):
//        Qp, Kp  [B*H*S][32]   per-head query / key rows zero-padded to one 128-byte swizzle span
//                              (Q pre-multiplied by log2(e)/sqrt(d): the softmax runs in base 2)
//        Vt      [B*H*DV][Sp]  per-head value matrix TRANSPOSED (keys contiguous): the K-major B operand of P.V
//  2. attn_tc: one CTA per (cloud, head, 128-query tile), flash-style over blocks of 64 keys:
//        warp 0   TMA producer: Q once; (K, Vt) blocks into a 2-stage ring
//        warp 1   MMA issuer:  S = Q.K^T (128x64, TMEM)  and  O += P.V (128xDV, TMEM), software-pipelined so that
//                 S of block j+1 is computed while the softmax of block j runs
//        warps 2-5 softmax, one query row per thread: tcgen05.ld S, online max / sum in base 2, rescale O in TMEM
//                 (tcgen05.ld/st), write P as TF32 halves straight into the swizzled K-major A-operand layout in shared
//                 memory, fence to the async proxy, hand over by mbarrier; finally O / l -> global.
// Every product (Q.K and P.V) is the 3-term split x_hi y_hi + x_lo y_hi + x_hi y_lo, i.e. fp32-accurate.
#include <float.h>
#include <math.h>

#include "nets.h"
#include "tc_common.h"

namespace mac {

namespace {

constexpr int kH = 4;        // heads
constexpr int kQT = 128;     // queries per CTA
constexpr int kKB = 64;      // keys per block
constexpr int kPad = 32;     // padded per-head q/k width (one SW128 span)

// ---- 1. operand preparation ---------------------------------------------------------------------
template <int DQK, int DV>
__global__ void __launch_bounds__(256) attn_prep_kernel(const float *__restrict__ qkv, int ldq, float *__restrict__ qp_hi,
                                                        float *__restrict__ qp_lo, float *__restrict__ kp_hi,
                                                        float *__restrict__ kp_lo, float *__restrict__ vt_hi,
                                                        float *__restrict__ vt_lo, int B, int S, int Sp)
{
    __shared__ float tile[32][33];
    const float scale = rsqrtf(static_cast<float>(DQK)) * 1.4426950408889634f;
    const int b = blockIdx.z, h = blockIdx.y;
    const int s0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const float *base = qkv + static_cast<size_t>(b) * S * ldq;
    const size_t head_row0 = (static_cast<size_t>(b) * kH + h) * S;
    // q / k rows: 32 tokens x 32 padded columns
    for (int r = ty; r < 32; r += 8) {
        const int s = s0 + r;
        if (s < S) {
            const float q = tx < DQK ? base[static_cast<size_t>(s) * ldq + h * DQK + tx] * scale : 0.f;
            const float k = tx < DQK ? base[static_cast<size_t>(s) * ldq + kH * DQK + h * DQK + tx] : 0.f;
            const float qh = to_tf32(q), kh = to_tf32(k);
            const size_t o = (head_row0 + s) * kPad + tx;
            qp_hi[o] = qh, qp_lo[o] = to_tf32(q - qh);
            kp_hi[o] = kh, kp_lo[o] = to_tf32(k - kh);
        }
    }
    // v: transpose 32 tokens x DV dims through shared memory, 32 dims at a time
    for (int d0 = 0; d0 < DV; d0 += 32) {
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int s = s0 + r;
            tile[r][tx] = s < S ? base[static_cast<size_t>(s) * ldq + 2 * kH * DQK + h * DV + d0 + tx] : 0.f;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {   // r = dim, tx = token
            const int s = s0 + tx;
            if (s < Sp) {
                const float v = tile[tx][r];
                const float vh = to_tf32(v);
                const size_t o = ((static_cast<size_t>(b) * kH + h) * DV + d0 + r) * Sp + s;
                vt_hi[o] = vh, vt_lo[o] = to_tf32(v - vh);
            }
        }
    }
}

// ---- 2. attention ---------------------------------------------------------------------------------
template <int DV>
struct AttnCfg {
    static constexpr int kQBytes = kQT * kPad * 4;                 // 16 KB per half
    static constexpr int kKBytes = kKB * kPad * 4;                 // 8 KB per half
    static constexpr int kVChunk = DV * 32 * 4;                    // one 32-key chunk of Vt (DV rows x 128 B)
    static constexpr int kVBytes = 2 * kVChunk;                    // 64 keys, per half
    static constexpr int kStage = 2 * kKBytes + 2 * kVBytes;       // K hi|lo, Vt hi|lo
    static constexpr int kPChunk = kQT * 32 * 4;                   // 16 KB: 128 rows x 32 keys
    static constexpr int kPBytes = 2 * kPChunk;                    // per half
    static constexpr int kSmem = 2 * kQBytes + 2 * kStage + 2 * kPBytes + 1024 + 256;
    static constexpr uint32_t kStageTx = kStage;
};

struct AttnParams {
    float *out;
    int ldo, S, n_qt;
    const int *lens;   // optional (B): ragged batches, cloud b holds lens[b] <= S tokens (rows beyond are padding)
};

template <int DV>
__global__ void __launch_bounds__(192, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap mapQhi, const __grid_constant__ CUtensorMap mapQlo,
               const __grid_constant__ CUtensorMap mapKhi, const __grid_constant__ CUtensorMap mapKlo,
               const __grid_constant__ CUtensorMap mapVhi, const __grid_constant__ CUtensorMap mapVlo, const AttnParams p)
{
    using C = AttnCfg<DV>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *q_hi = smem, *q_lo = smem + C::kQBytes;
    uint8_t *stage0 = smem + 2 * C::kQBytes;
    uint8_t *p_hi = stage0 + 2 * C::kStage, *p_lo = p_hi + C::kPBytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(p_lo + C::kPBytes);
    uint64_t *q_full = bars;            // Q landed
    uint64_t *kv_full = bars + 1;       // [2]
    uint64_t *kv_empty = bars + 3;      // [2]
    uint64_t *s_full = bars + 5;        // S = Q.K^T of the current block complete
    uint64_t *s_free = bars + 6;        // all softmax threads have loaded S
    uint64_t *p_ready = bars + 7;       // P written (and O rescaled)
    uint64_t *o_free = bars + 8;        // O += P.V of the previous block complete (P buffer and O reusable)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 9);
    auto k_hi = [&](int st) { return stage0 + st * C::kStage; };
    auto k_lo = [&](int st) { return stage0 + st * C::kStage + C::kKBytes; };
    auto v_hi = [&](int st) { return stage0 + st * C::kStage + 2 * C::kKBytes; };
    auto v_lo = [&](int st) { return stage0 + st * C::kStage + 2 * C::kKBytes + C::kVBytes; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x % p.n_qt, bh = blockIdx.x / p.n_qt;   // bh = b * H + h
    const int q0 = qt * kQT;
    const int len = p.lens ? p.lens[bh / kH] : p.S;   // tokens of this cloud (keys >= len are masked, queries skipped)
    if (q0 >= len) {   // padding-only tile (whole CTA, before any barrier / TMEM allocation): keep the rows finite
        for (int e = threadIdx.x; e < kQT * (DV / 4); e += blockDim.x) {
            const int qi = q0 + e / (DV / 4);
            if (qi < p.S)
                *reinterpret_cast<float4 *>(p.out + (static_cast<size_t>(bh / kH) * p.S + qi) * p.ldo + (bh % kH) * DV +
                                            (e % (DV / 4)) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return;
    }
    const int nb = (len + kKB - 1) / kKB;

    if (threadIdx.x == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_free, 128);
        mbar_init(p_ready, 128);
        mbar_init(o_free, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 128);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t t_s = tmem_base, t_o = tmem_base + 64;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, 2 * C::kQBytes);
            tma_load_2d(q_hi, &mapQhi, 0, bh * p.S + q0, q_full);
            tma_load_2d(q_lo, &mapQlo, 0, bh * p.S + q0, q_full);
            for (int j = 0; j < nb; ++j) {
                const int st = j & 1;
                mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&kv_full[st], C::kStageTx);
                tma_load_2d(k_hi(st), &mapKhi, 0, bh * p.S + j * kKB, &kv_full[st]);
                tma_load_2d(k_lo(st), &mapKlo, 0, bh * p.S + j * kKB, &kv_full[st]);
                for (int c = 0; c < 2; ++c) {
                    tma_load_2d(v_hi(st) + c * C::kVChunk, &mapVhi, j * kKB + c * 32, bh * DV, &kv_full[st]);
                    tma_load_2d(v_lo(st) + c * C::kVChunk, &mapVlo, j * kKB + c * 32, bh * DV, &kv_full[st]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = umma_idesc_tf32(kQT, kKB);
            constexpr uint32_t idesc_o = umma_idesc_tf32(kQT, DV);
            auto mma_pv = [&](int j) {   // O (+)= P(j) . V(j)
                const int st = j & 1;
                mbar_wait(p_ready, j & 1);
                tc_fence_after_sync();
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const uint32_t ah = smem_u32(p_hi) + c * C::kPChunk, al = smem_u32(p_lo) + c * C::kPChunk;
                    const uint32_t bh_ = smem_u32(v_hi(st)) + c * C::kVChunk, bl = smem_u32(v_lo(st)) + c * C::kVChunk;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t off = k * 32;
                        umma_tf32(t_o, umma_desc_k_sw128(al + off), umma_desc_k_sw128(bh_ + off), idesc_o, (j | c | k) != 0);
                        umma_tf32(t_o, umma_desc_k_sw128(ah + off), umma_desc_k_sw128(bl + off), idesc_o, 1);
                        umma_tf32(t_o, umma_desc_k_sw128(ah + off), umma_desc_k_sw128(bh_ + off), idesc_o, 1);
                    }
                }
                umma_commit(o_free);
                umma_commit(&kv_empty[st]);
            };
            mbar_wait(q_full, 0);
            for (int j = 0; j < nb; ++j) {
                const int st = j & 1;
                mbar_wait(&kv_full[st], (j >> 1) & 1);
                if (j > 0) mbar_wait(s_free, (j - 1) & 1);   // the softmax has read S of block j-1
                tc_fence_after_sync();
                const uint32_t qh = smem_u32(q_hi), ql = smem_u32(q_lo), kh = smem_u32(k_hi(st)), kl = smem_u32(k_lo(st));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t off = k * 32;
                    umma_tf32(t_s, umma_desc_k_sw128(ql + off), umma_desc_k_sw128(kh + off), idesc_s, k != 0);
                    umma_tf32(t_s, umma_desc_k_sw128(qh + off), umma_desc_k_sw128(kl + off), idesc_s, 1);
                    umma_tf32(t_s, umma_desc_k_sw128(qh + off), umma_desc_k_sw128(kh + off), idesc_s, 1);
                }
                umma_commit(s_full);
                if (j > 0) mma_pv(j - 1);
            }
            mma_pv(nb - 1);
        }
    } else {
        // ===== softmax: one query row per thread =====
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        float m = -1.0e30f, l = 0.f;
        float v[32];
        const uint32_t row_off = (r >> 3) * 1024 + (r & 7) * 128;   // swizzled K-major tile: 8-row groups of 1024 B
        for (int j = 0; j < nb; ++j) {
            float s[kKB];
            mbar_wait(s_full, j & 1);
            tc_fence_after_sync();
            tmem_ld32(t_s + lane_addr, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) s[i] = v[i];
            tmem_ld32(t_s + lane_addr + 32, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) s[32 + i] = v[i];
            tc_fence_before_sync();
            mbar_arrive(s_free);
            const int nvalid = len - j * kKB;   // keys of this block inside the cloud
            float bm = m;
#pragma unroll
            for (int i = 0; i < kKB; ++i) {
                if (i >= nvalid) s[i] = -1.0e30f;
                bm = fmaxf(bm, s[i]);
            }
            const float corr = exp2f(m - bm);
            m = bm;
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < kKB; ++i) {
                s[i] = i < nvalid ? exp2f(s[i] - m) : 0.f;
                sum += s[i];
            }
            l = l * corr + sum;
            if (j > 0) {
                // the previous P.V has completed: O may be rescaled and the P buffer overwritten
                mbar_wait(o_free, (j - 1) & 1);
                tc_fence_after_sync();
#pragma unroll
                for (int c0 = 0; c0 < DV; c0 += 32) {
                    tmem_ld32(t_o + lane_addr + c0, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] *= corr;
                    tmem_st32(t_o + lane_addr + c0, v);
                }
            }
            // P -> shared memory, TF32 halves, SWIZZLE_128B K-major layout (16-byte piece index XOR row & 7)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int pc = 0; pc < 8; ++pc) {
                    float4 h4, l4;
                    const float *src = s + c * 32 + pc * 4;
                    h4.x = to_tf32(src[0]), h4.y = to_tf32(src[1]), h4.z = to_tf32(src[2]), h4.w = to_tf32(src[3]);
                    l4.x = to_tf32(src[0] - h4.x), l4.y = to_tf32(src[1] - h4.y), l4.z = to_tf32(src[2] - h4.z), l4.w = to_tf32(src[3] - h4.w);
                    const uint32_t off = c * C::kPChunk + row_off + ((pc ^ (r & 7)) << 4);
                    *reinterpret_cast<float4 *>(p_hi + off) = h4;
                    *reinterpret_cast<float4 *>(p_lo + off) = l4;
                }
            }
            fence_proxy_async_smem();
            tc_fence_before_sync();
            mbar_arrive(p_ready);
        }
        // ---- O / l -> global ----
        mbar_wait(o_free, (nb - 1) & 1);
        tc_fence_after_sync();
        const int qi = q0 + r;
        const float inv = 1.0f / l;
        const int b = bh / kH, h = bh % kH;
#pragma unroll
        for (int c0 = 0; c0 < DV; c0 += 32) {
            tmem_ld32(t_o + lane_addr + c0, v);
            if (qi < p.S) {   // padding rows (qi >= len) are written as zeros so that later layers stay finite
                float *dst = p.out + (static_cast<size_t>(b) * p.S + qi) * p.ldo + h * DV + c0;
                const float sc = qi < len ? inv : 0.f;
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4 *>(dst + i) = make_float4(v[i] * sc, v[i + 1] * sc, v[i + 2] * sc, v[i + 3] * sc);
            }
        }
        tc_fence_before_sync();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 128);
    }
}

template <int DQK, int DV>
int run(const float *qkv, int ldq, float *out, int ldo, int B, int S, float *scratch, cudaStream_t stream, const int *lens)
{
    using C = AttnCfg<DV>;
    const int Sp = (S + 3) / 4 * 4;
    const size_t n_qk = static_cast<size_t>(B) * kH * S * kPad, n_v = static_cast<size_t>(B) * kH * DV * Sp;
    float *qp_hi = scratch, *qp_lo = qp_hi + n_qk, *kp_hi = qp_lo + n_qk, *kp_lo = kp_hi + n_qk;
    float *vt_hi = kp_lo + n_qk, *vt_lo = vt_hi + n_v;
    dim3 pgrid((Sp + 31) / 32, kH, B);
    attn_prep_kernel<DQK, DV><<<pgrid, 256, 0, stream>>>(qkv, ldq, qp_hi, qp_lo, kp_hi, kp_lo, vt_hi, vt_lo, B, S, Sp);
    MAC_CUDA(cudaGetLastError());

    CUtensorMap mQh, mQl, mKh, mKl, mVh, mVl;
    const int rows_qk = B * kH * S;
    if (int rc = make_tensor_map_2d(&mQh, qp_hi, rows_qk, kPad, kPad, kQT)) return rc;
    if (int rc = make_tensor_map_2d(&mQl, qp_lo, rows_qk, kPad, kPad, kQT)) return rc;
    if (int rc = make_tensor_map_2d(&mKh, kp_hi, rows_qk, kPad, kPad, kKB)) return rc;
    if (int rc = make_tensor_map_2d(&mKl, kp_lo, rows_qk, kPad, kPad, kKB)) return rc;
    if (int rc = make_tensor_map_2d(&mVh, vt_hi, B * kH * DV, S, Sp, DV)) return rc;
    if (int rc = make_tensor_map_2d(&mVl, vt_lo, B * kH * DV, S, Sp, DV)) return rc;
    static bool configured = false;
    if (!configured) {
        MAC_CUDA(cudaFuncSetAttribute(attn_tc_kernel<DV>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
        configured = true;
    }
    AttnParams p{out, ldo, S, (S + kQT - 1) / kQT, lens};
    attn_tc_kernel<DV><<<B * kH * p.n_qt, 192, C::kSmem, stream>>>(mQh, mQl, mKh, mKl, mVh, mVl, p);
    MAC_CUDA(cudaGetLastError());
    count_launch(2);
    return MAC_OK;
}

}  // namespace

size_t attn_dense_tc_scratch_floats(int B, int S, int dv)
{
    const size_t Sp = (static_cast<size_t>(S) + 3) / 4 * 4;
    return 4 * static_cast<size_t>(B) * kH * S * kPad + 2 * static_cast<size_t>(B) * kH * dv * Sp + 64;
}

int attn_dense_tc(const float *qkv, int ldq, float *out, int ldo, int B, int S, int dqk, int dv, float *scratch,
                  cudaStream_t stream, const int *lens)
{
    MAC_REQUIRE(qkv && out && scratch && B > 0 && S > 0, "null tensor pointer");
    MAC_REQUIRE(ldq % 4 == 0 && ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(scratch) & 15u) == 0, "attention buffers must be 16-byte aligned");
    if (dqk == 8 && dv == 32) return run<8, 32>(qkv, ldq, out, ldo, B, S, scratch, stream, lens);
    if (dqk == 16 && dv == 64) return run<16, 64>(qkv, ldq, out, ldo, B, S, scratch, stream, lens);
    set_error("attn_dense_tc is built for 4 heads of (8, 32) or (16, 64) dims, got (%d, %d)", dqk, dv);
    return MAC_ERR_UNSUPPORTED;
}

}  // namespace mac
