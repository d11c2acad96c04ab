// Dense multi-head self-attention on the 5th-generation tensor cores (reference networks/Attention.py:8-36, 182-204,
// mask = None): softmax(Q K^T / sqrt(d)) V for clouds of S tokens, 4 heads, per-head dims (8, 32) [SconeOcc global
// transformer] or (16, 64) [SconeVis].
//
// Two kernels:
//  1. attn_prep: reads the fused qkv rows once and writes the tensor-core operands, already split into TF32 halves
//     (x_hi = tf32(x), x_lo = tf32(x - x_hi), see linear.cu):
//        Qp, Kp  [B*H*S][32]   one 128-byte row per (head, token): [hi (DQK) | lo (DQK) | 0], i.e. both halves of a
//                              query / key share one swizzle span (Q pre-multiplied by log2(e)/sqrt(d): base-2 softmax)
//        Vt      [B*H*DV][Sp]  per-head value matrix TRANSPOSED (keys contiguous): the K-major B operand of P.V
//  2. attn_tc: one CTA per (cloud, head, 128-query tile), two CTAs per SM, flash-style over blocks of 64 keys:
//        warp 0   TMA producer: Q once; (K, Vt) blocks into a 2-stage ring
//        warp 1   MMA issuer:  S = Q.K^T (128x64, TMEM)  and  O += P.V (128xDV, TMEM), software-pipelined so that
//                 S of block j+1 is computed while the softmax of block j runs
//        warps 2-5 softmax, one query row per thread: tcgen05.ld S, online max / sum in base 2 with a lazily updated
//                 reference maximum (O in TMEM is only rescaled when the maximum moved by more than 2^8), P written
//                 back to TENSOR MEMORY as TF32 halves (tcgen05.st) and consumed from there as the A operand of P.V
//                 (tcgen05.mma with A in TMEM), so P never touches shared memory; finally O / l -> global.
//     The kernel was shared-memory-bandwidth bound when P went through shared memory (ncu: tensor pipe 30 %, the
//     softmax warps 58 % of their time waiting for S): per 64-key block the UMMAs read 216 KB of operands and the
//     softmax wrote 64 KB; now 84 KB are read and nothing is written (profiles/r01t_attention.md).
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
constexpr int kRow = 32;     // floats per packed q / k row (one SW128 span): [hi | lo | 0]
constexpr float kSlack = 8.f;  // the running softmax maximum may lag the true one by this much (base-2 exponent)

// ---- 1. operand preparation ---------------------------------------------------------------------
template <int DQK, int DV>
__global__ void __launch_bounds__(256) attn_prep_kernel(const float *__restrict__ qkv, int ldq, float *__restrict__ qp,
                                                        float *__restrict__ kp, float *__restrict__ vt_hi,
                                                        float *__restrict__ vt_lo, int B, int S, int Sp)
{
    static_assert(2 * DQK <= kRow, "hi and lo halves of a query / key row must fit one 128-byte span");
    __shared__ float tile[32][33];
    const float scale = rsqrtf(static_cast<float>(DQK)) * 1.4426950408889634f;
    const int b = blockIdx.z, h = blockIdx.y;
    const int s0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const float *base = qkv + static_cast<size_t>(b) * S * ldq;
    const size_t head_row0 = (static_cast<size_t>(b) * kH + h) * S;
    // q / k rows: 32 tokens x [hi (DQK) | lo (DQK) | 0]
    for (int r = ty; r < 32; r += 8) {
        const int s = s0 + r;
        if (s < S) {
            float qv = 0.f, kv = 0.f;
            if (tx < 2 * DQK) {
                const int e = tx < DQK ? tx : tx - DQK;
                const float q = base[static_cast<size_t>(s) * ldq + h * DQK + e] * scale;
                const float k = base[static_cast<size_t>(s) * ldq + kH * DQK + h * DQK + e];
                const float qh = to_tf32(q), kh = to_tf32(k);
                qv = tx < DQK ? qh : to_tf32(q - qh);
                kv = tx < DQK ? kh : to_tf32(k - kh);
            }
            const size_t o = (head_row0 + s) * kRow + tx;
            qp[o] = qv;
            kp[o] = kv;
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
    static constexpr int kQBytes = kQT * kRow * 4;                 // 16 KB: 128 packed query rows
    static constexpr int kKBytes = kKB * kRow * 4;                 // 8 KB: 64 packed key rows
    static constexpr int kVChunk = DV * 32 * 4;                    // one 32-key chunk of Vt (DV rows x 128 B)
    static constexpr int kVBytes = 2 * kVChunk;                    // 64 keys, per half
    static constexpr int kStage = kKBytes + 2 * kVBytes;           // K, Vt hi | lo
    static constexpr int kSmem = kQBytes + 2 * kStage + 1024 + 256;   // 97.25 KB (DV = 64): two CTAs per SM
    static constexpr uint32_t kStageTx = kStage;
    // tensor memory columns: S | O | P_hi | P_lo
    static constexpr int kColS = 0, kColO = kKB, kColPhi = 2 * kKB, kColPlo = 3 * kKB, kTmemCols = 4 * kKB;
    static_assert(DV <= kKB, "O must fit between S and P");
};

struct AttnParams {
    float *out;
    int ldo, S, n_qt;
    const int *lens;   // optional (B): ragged batches, cloud b holds lens[b] <= S tokens (rows beyond are padding)
};

template <int DQK, int DV>
__global__ void __launch_bounds__(192, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
               const __grid_constant__ CUtensorMap mapVhi, const __grid_constant__ CUtensorMap mapVlo, const AttnParams p)
{
    using C = AttnCfg<DV>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *q_sm = smem;
    uint8_t *stage0 = smem + C::kQBytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(stage0 + 2 * C::kStage);
    uint64_t *q_full = bars;            // Q landed
    uint64_t *kv_full = bars + 1;       // [2]
    uint64_t *kv_empty = bars + 3;      // [2]
    uint64_t *s_full = bars + 5;        // S = Q.K^T of the current block complete
    uint64_t *s_free = bars + 6;        // all softmax threads have loaded S
    uint64_t *p_ready = bars + 7;       // P written (and O rescaled)
    uint64_t *o_free = bars + 8;        // O += P.V of the previous block complete (P buffer and O reusable)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 9);
    auto k_sm = [&](int st) { return stage0 + st * C::kStage; };
    auto v_hi = [&](int st) { return stage0 + st * C::kStage + C::kKBytes; };
    auto v_lo = [&](int st) { return stage0 + st * C::kStage + C::kKBytes + C::kVBytes; };

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
    if (warp == 1) tmem_alloc(tmem_slot, C::kTmemCols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t t_s = tmem_base + C::kColS, t_o = tmem_base + C::kColO;
    const uint32_t t_phi = tmem_base + C::kColPhi, t_plo = tmem_base + C::kColPlo;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, C::kQBytes);
            tma_load_2d(q_sm, &mapQ, 0, bh * p.S + q0, q_full);
            for (int j = 0; j < nb; ++j) {
                const int st = j & 1;
                mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&kv_full[st], C::kStageTx);
                tma_load_2d(k_sm(st), &mapK, 0, bh * p.S + j * kKB, &kv_full[st]);
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
            auto mma_pv = [&](int j) {   // O (+)= P(j) . V(j), P read from tensor memory (one 8-key k-step = 8 columns)
                const int st = j & 1;
                mbar_wait(p_ready, j & 1);
                tc_fence_after_sync();
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const uint32_t bh_ = smem_u32(v_hi(st)) + c * C::kVChunk, bl = smem_u32(v_lo(st)) + c * C::kVChunk;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t off = k * 32, col = (c * 4 + k) * 8;
                        umma_tf32_ts(t_o, t_plo + col, umma_desc_k_sw128(bh_ + off), idesc_o, (j | c | k) != 0);
                        umma_tf32_ts(t_o, t_phi + col, umma_desc_k_sw128(bl + off), idesc_o, 1);
                        umma_tf32_ts(t_o, t_phi + col, umma_desc_k_sw128(bh_ + off), idesc_o, 1);
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
                // packed rows: hi halves at byte 0, lo halves at byte 4 * DQK of every 128-byte row
                const uint32_t qh = smem_u32(q_sm), ql = qh + 4 * DQK, kh = smem_u32(k_sm(st)), kl = kh + 4 * DQK;
#pragma unroll
                for (int k = 0; k < DQK / 8; ++k) {
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
            if (nvalid < kKB) {                 // tail block only: keys beyond the cloud get zero weight
#pragma unroll
                for (int i = 0; i < kKB; ++i)
                    if (i >= nvalid) s[i] = -1.0e30f;
            }
            float bm = s[0];
#pragma unroll
            for (int i = 1; i < kKB; ++i) bm = fmaxf(bm, s[i]);
            // Lazy running maximum: the reference point m of the base-2 exponentials only moves when some row of
            // this warp exceeds it by more than kSlack (warp-uniform decision: the TMEM accesses are warp-collective),
            // so the weights stay <= 2^kSlack and the rescaling of O in TMEM (ld / mul / st of DV columns) is skipped
            // for almost every block after the first few.  softmax is invariant to the choice of m.
            const bool moved = __any_sync(0xffffffffu, bm > m + kSlack);
            float corr = 1.f;
            if (moved) {
                const float mn = fmaxf(m, bm);
                corr = ex2_approx(m - mn);
                m = mn;
            }
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < kKB; ++i) {
                s[i] = ex2_approx(s[i] - m);
                sum += s[i];
            }
            l = l * corr + sum;
            if (j > 0) {
                // the previous P.V has completed: O may be rescaled and the P buffer overwritten
                mbar_wait(o_free, (j - 1) & 1);
                tc_fence_after_sync();
                if (moved) {
#pragma unroll
                    for (int c0 = 0; c0 < DV; c0 += 32) {
                        tmem_ld32(t_o + lane_addr + c0, v);
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] *= corr;
                        tmem_st32(t_o + lane_addr + c0, v);
                    }
                }
            }
            // P -> tensor memory as TF32 halves, this thread's lane, one column per key: the A operand of P.V.
            // hi = p with the 13 low mantissa bits cleared (a TF32 value), lo = p - hi exactly (<= 2^-10 p; the tensor
            // core reads its top 19 bits): p = hi + lo to 2^-21 relative, with 2 full-rate instructions per weight
            // instead of two emulated round-to-nearest conversions.
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float lo[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    v[i] = tf32_hi(s[c * 32 + i]);
                    lo[i] = s[c * 32 + i] - v[i];
                }
                tmem_st32(t_phi + lane_addr + c * 32, v);
                tmem_st32(t_plo + lane_addr + c * 32, lo);
            }
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
        tmem_dealloc(tmem_base, C::kTmemCols);
    }
}

template <int DQK, int DV>
int run(const float *qkv, int ldq, float *out, int ldo, int B, int S, float *scratch, cudaStream_t stream, const int *lens)
{
    using C = AttnCfg<DV>;
    const int Sp = (S + 3) / 4 * 4;
    const size_t n_qk = static_cast<size_t>(B) * kH * S * kRow, n_v = static_cast<size_t>(B) * kH * DV * Sp;
    float *qp = scratch, *kp = qp + n_qk;
    float *vt_hi = kp + n_qk, *vt_lo = vt_hi + n_v;
    dim3 pgrid((Sp + 31) / 32, kH, B);
    attn_prep_kernel<DQK, DV><<<pgrid, 256, 0, stream>>>(qkv, ldq, qp, kp, vt_hi, vt_lo, B, S, Sp);
    MAC_CUDA(cudaGetLastError());

    CUtensorMap mQ, mK, mVh, mVl;
    const int rows_qk = B * kH * S;
    if (int rc = make_tensor_map_2d(&mQ, qp, rows_qk, kRow, kRow, kQT)) return rc;
    if (int rc = make_tensor_map_2d(&mK, kp, rows_qk, kRow, kRow, kKB)) return rc;
    if (int rc = make_tensor_map_2d(&mVh, vt_hi, B * kH * DV, S, Sp, DV)) return rc;
    if (int rc = make_tensor_map_2d(&mVl, vt_lo, B * kH * DV, S, Sp, DV)) return rc;
    static DeviceOnce once;
    if (int rc = ensure_dynamic_smem(once, attn_tc_kernel<DQK, DV>, C::kSmem)) return rc;
    AttnParams p{out, ldo, S, (S + kQT - 1) / kQT, lens};
    attn_tc_kernel<DQK, DV><<<B * kH * p.n_qt, 192, C::kSmem, stream>>>(mQ, mK, mVh, mVl, p);
    MAC_CUDA(cudaGetLastError());
    count_launch(2);
    return MAC_OK;
}

}  // namespace

size_t attn_dense_tc_scratch_floats(int B, int S, int dv)
{
    const size_t Sp = (static_cast<size_t>(S) + 3) / 4 * 4;
    return 2 * static_cast<size_t>(B) * kH * S * kRow + 2 * static_cast<size_t>(B) * kH * dv * Sp + 64;
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
