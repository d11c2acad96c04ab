// Fused linear layer on the 5th-generation tensor cores:
//     Y = epilogue( X (M,K) . W (N,K)^T + bias )
// the building block of every dense contraction of SconeOcc / SconeVis (reference nn.Linear calls in
// networks/Attention.py:96-98,160-162,173,225-226, networks/SconeOcc.py:18-22,103,227-229,
// networks/SconeVis.py:114-119).
//
// Design (sm_100a):
//   * CTA tile 128 (M) x BN (N), K swept in chunks of 32 fp32 = 128-byte rows.  Warp 0 streams the X and W
//     tiles global -> shared with TMA (SWIZZLE_128B tensor maps, mbarrier completion, out-of-range rows
//     and columns zero-filled by the hardware, so K, N, M need no padding).
//   * warp 1 issues tcgen05.mma kind::tf32 (UMMA 128 x BN x 8) with the fp32 accumulator in TMEM.
//   * fp32 accuracy: TF32 keeps 10 mantissa bits, far too few for parity with the fp32 reference, so
//     every product is evaluated as  x_hi w_hi + x_lo w_hi + x_hi w_lo  (x = x_hi + x_lo with both parts
//     exactly representable in TF32; the dropped x_lo w_lo term is 2^-22 relative).  W is split once when
//     the weights are packed; X is split in shared memory by 4 dedicated warps (element-wise, hence
//     oblivious to the swizzle; x_hi by clearing the low mantissa bits, x_lo = x - x_hi rounded with integer
//     arithmetic: 4 full-rate instructions per element, tc_common.h), fenced to the async proxy and handed to
//     the MMA warp through a second mbarrier.
//   * persistent CTAs (one per SM) walk the tiles; two TMEM accumulators let the TMA / split / MMA warps work
//     on tile t+1 while the 4 epilogue warps drain tile t.
//   * epilogue: one output row per thread, 32 columns at a time: tcgen05.ld -> bias -> activation (ReLU /
//     exact GELU) -> residual add.  Residual chunks arrive by cp.async into a padded staging buffer and the
//     results leave through the same buffer, so that every global access is a full 128-byte row segment
//     (row-per-thread accesses straight to global memory made the first version 10x slower).  Optionally
//     the rows are also LayerNorm-ed for the next layer (y is parked in TMEM with tcgen05.st; mean, centred
//     variance and the normalised output are three cheap TMEM passes; the tile spans all N) or max / mean
//     pooled over groups of 16 consecutive rows (the 16-token neighbourhoods of SconeOcc).
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "nets.h"
#include "tc_common.h"

namespace mac {

int make_tensor_map_2d(CUtensorMap *map, const float *base, int rows, int cols, int ld, int box_rows)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        MAC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled is not available from this driver");
            return MAC_ERR_CUDA;
        }
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    MAC_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15u) == 0 && ld % 4 == 0,
                "TMA operand must be 16-byte aligned with a row stride that is a multiple of 4 floats (ld=%d)", ld);
    MAC_REQUIRE(rows > 0 && cols > 0 && cols <= ld && box_rows >= 1 && box_rows <= 256, "bad tensor map shape");
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * sizeof(float)};
    const cuuint32_t box[2] = {32, static_cast<cuuint32_t>(box_rows)};
    const cuuint32_t elem[2] = {1, 1};
    const CUresult rc = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box,
                               elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%d cols=%d ld=%d box_rows=%d)",
                  static_cast<int>(rc), rows, cols, ld, box_rows);
        return MAC_ERR_CUDA;
    }
    return MAC_OK;
}

namespace {

constexpr int kBM = 128;
constexpr int kBK = 32;             // fp32 per k-chunk: one 128-byte swizzle span
constexpr int kUmmaK = 8;           // tf32 MMA depth (32 bytes)
constexpr int kATileBytes = kBM * kBK * 4;
constexpr int kBaseThreads = 256;   // warp 0 X TMA, warp 1 MMA, warp 2 W TMA, warps 4-7 X split; then 4 warps per epilogue group
constexpr int kMaxStages = 8;
constexpr int kEpiLd = 36;          // epilogue staging row stride (floats): 32 columns + 4 pad, conflict-free float4 rows
constexpr int kEpiPadBytes = kBM * kEpiLd * 4;   // plain-store path: one padded 128 x 36 staging tile per epilogue group
constexpr int kEpiBar = 1;          // named barrier of the 128 epilogue threads

struct LinearParams {
    int M, N, K;
    int n_tiles_m, n_tiles_n;
    const float *bias;
    float *out;
    int ldo;
    const float *res;
    int ldr;
    float *ln_out;
    int ldl;
    const float *ln_g, *ln_b;
    float ln_eps;
    int act;
    int pool;       // 0, or 16: max | mean over groups of 16 rows -> out (M/16, 2N)
    int res_first;  // 1: out = act(acc + bias + res) (ResNet blocks); 0: out = act(acc + bias) + res (transformer residuals)
    // LayerNorm applied to X on load (by the split warps): x' = (x - mean[row]) * rstd[row] * g[col] + b[col]
    const float *lnin_stats;   // (M, 2) = (mean, rstd) per row, or null
    const float *lnin_g, *lnin_b;   // (K)
    float *stats_out;          // (M, 2): (mean, rstd) of every output row (for the LayerNorm-on-load of the next layer)
    ConvGather conv;           // GATHER kernels only: X is the implicit im2col matrix of this convolution
    // split-K: tile = (m tile, n tile, split); split s contracts the k-chunks [s * chunks_per_split, ...) and writes its raw
    // fp32 partial tile to out + s * split_stride (bias / activation / residual are then applied by splitk_finish_kernel)
    int splits, chunks_per_split;
    long long split_stride;
    int act_in;                // MAC_LIN_GELU: the exact GELU is applied to X on load (by the split warps), else MAC_LIN_NONE
    int tma_out;               // 1: `out` leaves through TMA stores (mapOut) from swizzled, double-buffered staging tiles
};

#ifdef MAC_LINEAR_PROFILE
// debug build only (MAC_EXTRA_NVCC_FLAGS=-DMAC_LINEAR_PROFILE): cycles the UMMA-issuing thread spends waiting on each
// barrier class, summed over CTAs: [0] accumulator free, [1] X (hi / lo) ready, [2] W ready, [3] whole role, [4] tiles
__device__ unsigned long long g_linear_prof[12];
#endif

struct TileCoord {
    int m0, n0, k0, k1, sp;
};
template <int BN>
__device__ __forceinline__ TileCoord tile_coord(const LinearParams &p, const int tile, const int nk)
{
    const int sp = tile % p.splits, mn = tile / p.splits;
    const int k0 = sp * p.chunks_per_split;
    return TileCoord{(mn / p.n_tiles_n) * kBM, (mn % p.n_tiles_n) * BN, k0, min(nk, k0 + p.chunks_per_split), sp};
}

// GELU(x) = x/2 (1 + erf(x / sqrt 2)) with erf from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, the accuracy class of
// erff itself): erf(z) = 1 - (a1 t + ... + a5 t^5) exp(-z^2), t = 1 / (1 + p z), z >= 0.  14 instructions with two MUFU
// (rcp, ex2) instead of ~30 for erff: the GELU epilogues (ff1 layers, N = 256 / 512 per row) were issue-bound on it.
__device__ __forceinline__ float gelu_exact(float x)
{
    const float h = 0.5f * x;
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float e = ex2_approx(z * z * -1.4426950408889634f);
    const float erf_abs = fmaf(-p * t, e, 1.0f);          // erf(|x| / sqrt 2)
    return fmaf(fabsf(h), erf_abs, h);                     // x/2 + |x|/2 erf(|x|/sqrt 2) = x/2 (1 + erf(x/sqrt 2))
}

template <int ACT>
__device__ __forceinline__ float apply_act(float y)
{
    if (ACT == MAC_LIN_RELU) return fmaxf(y, 0.f);
    if (ACT == MAC_LIN_GELU) return gelu_exact(y);
    if (ACT == MAC_LIN_ELU) return y > 0.f ? y : expf(y) - 1.0f;
    if (ACT == MAC_LIN_SIGMOID) return 1.0f / (1.0f + expf(-y));
    return y;
}

// Shared-memory plan.  Independent rings so that the number of X bytes in flight from HBM is not tied to the (large,
// L2-resident) weight tiles:
//   raw X ring   kRawSlots x 16 KB   TMA / gather destination, read once by the split warps (SPLIT) or by the UMMAs
//   W ring       kWSlots x (1 or 2) x BN x 128 B   W_hi (and W_lo) k-chunks
//   staging      per epilogue group: two (one at BN = 256) swizzled 128 x 32 tiles for the TMA stores
//
// A operand in tensor memory (kATmem = every SPLIT layer).  Per k-chunk the operands of the 12 UMMAs are read 3 times: with
// A in shared memory (first version) that was 96 KB of operand reads on top of the 48 KB TMA writes and the 16 KB read +
// 32 KB written by the split -- 192 KB per chunk against a shared-memory port of 128 B / clock = 1536 clocks, MORE than the
// tensor pipe needs: the layers were shared-memory-bandwidth bound (measured 7.5 k clocks per K = 128 tile, floor 6.1 k).
// The split warps hold a row per thread and write tf32(x) / tf32(x - tf32(x)) to TENSOR memory with tcgen05.st (2 x 64
// columns behind the accumulators); the UMMAs read A from there ("ts" form) and only W from shared memory: 112 KB per
// chunk, no lo ring, and the raw X slot is recycled as soon as the split has read it.
template <int BN, bool SPLIT>
struct Cfg {
    static constexpr bool kATmem = SPLIT;
    // 512 TMEM columns: two accumulators + the A slots up to BN = 192; at BN = 256 ONE accumulator (the UMMAs of tile t+1 then
    // wait for the epilogue of tile t; with A in shared memory and one epilogue group that tile shape ran at 1.5 TB/s)
    static constexpr int kAccBufs = (!kATmem || 2 * BN + 128 <= 512) ? 2 : 1;
    static constexpr int kStageTiles = BN <= 192 ? 2 : 1;   // staging tiles per epilogue group (TMA-store path)
    static constexpr int kBTileBytes = BN * kBK * 4;
    static constexpr int kWSlotBytes = (SPLIT ? 2 : 1) * kBTileBytes;
    static constexpr int kEpiGroups2 = BN <= 128 || kATmem;   // room for a second epilogue group (no lo ring)
    static constexpr int kEpiBufs = kEpiGroups2 ? 2 : 1;
    static constexpr bool kTmaOut = kATmem;   // (the kernels with the lo ring have no room for the second staging tile)
    // (the plain-store path of the same kernel -- LayerNorm-ed second output, pooling, split-K partials -- needs the padded tile)
    static constexpr int kEpiBufBytes = (kTmaOut && kStageTiles * kBM * 128 > kEpiPadBytes) ? kStageTiles * kBM * 128 : kEpiPadBytes;
    static constexpr int kWSlots = 2;
    static constexpr int kLnInMaxK = 512;
    static constexpr int kVecBytes = 3 * BN * 4 + 4 * 128 * 4 +   // bias | gamma | beta of the tile's columns, LayerNorm partials (sum | variance)
                                     (SPLIT ? 2 * kLnInMaxK * 4 : 0);   // gamma | beta (K) of the LayerNorm applied on load
    static constexpr int kThreads = kBaseThreads + 128 * kEpiBufs;
    static constexpr int kBudget = 226 * 1024 - 1024 - 512 - kVecBytes - kEpiBufs * kEpiBufBytes - kWSlots * kWSlotBytes;
    static constexpr int kRawSlots = (kBudget / kATileBytes) < kMaxStages ? (kBudget / kATileBytes) : kMaxStages;
    static constexpr int kOperandBytes = kRawSlots * kATileBytes + kWSlots * kWSlotBytes;
    static constexpr int kSmemBytes = kOperandBytes + kEpiBufs * kEpiBufBytes + kVecBytes + 1024 /*alignment*/ + 512 /*barriers*/;
    static constexpr uint32_t kWTxBytes = kWSlotBytes;
    static_assert(kRawSlots >= 2, "need at least a double-buffered X pipeline");
};

// Persistent: every CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the TMA and MMA warps run ahead of the
// epilogue through a ring of operand stages and two TMEM accumulators.
// GATHER: X is never materialised; warps 0 and 3 build the X k-chunks of a convolution (ConvGather) straight from the NHWC
// activations with 16-byte cp.async copies laid out as the SWIZZLE_128B pattern the split warps and the UMMA descriptors
// expect (zero fill for padding, rows beyond M and columns beyond K), completing on the same full_x barriers.
template <int BN, bool SPLIT, int ACT, bool GATHER>
__global__ void __launch_bounds__((Cfg<BN, SPLIT>::kThreads), 1)
linear_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBhi,
              const __grid_constant__ CUtensorMap mapBlo, const __grid_constant__ CUtensorMap mapOut, const LinearParams p)
{
    using C = Cfg<BN, SPLIT>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t *smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t *raw_base = smem;
    uint8_t *w_base = raw_base + C::kRawSlots * kATileBytes;
    float *ebuf = reinterpret_cast<float *>(smem + C::kOperandBytes);
    float *svec = ebuf + C::kEpiBufs * (C::kEpiBufBytes / 4);   // [3][BN]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::kOperandBytes + C::kEpiBufs * C::kEpiBufBytes + C::kVecBytes);
    uint64_t *full_x = bars;                      // [kRawSlots] X chunk landed
    uint64_t *empty_x = bars + kMaxStages;        // [kRawSlots] MMAs that read the raw slot have completed
    uint64_t *ready_lo = bars + 2 * kMaxStages;   // [2] split done: tf32 hi / lo of the chunk written to A slot s of tensor memory
    uint64_t *empty_lo = ready_lo + 2;            // [2] UMMAs that read A slot s have completed
    uint64_t *full_w = empty_lo + 2;              // [2] W chunk landed
    uint64_t *empty_w = full_w + 2;               // [2]
    uint64_t *acc_full = empty_w + 2;             // [2] accumulator complete
    uint64_t *acc_empty = acc_full + 2;           // [2] accumulator drained by the epilogue
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    auto a_hi = [&](int s) { return raw_base + s * kATileBytes; };
    auto b_hi = [&](int s) { return w_base + s * C::kWSlotBytes; };
    auto b_lo = [&](int s) { return w_base + s * C::kWSlotBytes + C::kBTileBytes; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = (p.K + kBK - 1) / kBK;
    const int n_tiles = p.n_tiles_m * p.n_tiles_n * p.splits;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapBhi);
        if (SPLIT) tma_prefetch_desc(&mapBlo);
        for (int s = 0; s < C::kRawSlots; ++s) {
            mbar_init(&full_x[s], GATHER ? 64 : 1);   // GATHER: one arrival per lane of the two gather warps
            mbar_init(&empty_x[s], C::kATmem ? 128 : 1);   // A in TMEM: released by the 128 split threads, else by the UMMAs
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&ready_lo[s], 128);
            mbar_init(&empty_lo[s], 1);
            mbar_init(&full_w[s], 1);
            mbar_init(&empty_w[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 128 * C::kEpiBufs);
        }
        fence_mbar_init();
    }
    constexpr int kTmemNeed = C::kAccBufs * BN + (C::kATmem ? 128 : 0);   // accumulators (+ two A slots of hi | lo, 32 columns each)
    constexpr uint32_t kTmemCols = kTmemNeed <= 128 ? 128u : (kTmemNeed <= 256 ? 256u : 512u);   // power of two
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (GATHER && (warp == 0 || warp == 3)) {
        // ===== X producer (implicit GEMM): warp 0 gathers rows 0-63 of the tile, warp 3 rows 64-127; a lane owns two
        // consecutive output pixels and copies their 128-byte k-chunk rows as eight 16-byte pieces =====
        static_assert(!GATHER || SPLIT, "the gathered operand is consumed by the split warps");
        const ConvGather &g = p.conv;
        const int Cin = g.Ca + g.Cb;
        const int seg_len = Cin < kBK ? Cin : kBK;     // 32, or 16: two taps per k-chunk
        const int n_seg = kBK / seg_len;
        const int taps = g.k * g.k;
        const int r0 = (warp == 0 ? 0 : 64) + 2 * lane;
        const bool resample = g.Ha != g.H || g.Wa != g.W;
        int kt = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const TileCoord tc = tile_coord<BN>(p, tile, nk);
            const int m0 = tc.m0;
            int img[2], iy0[2], ix0[2];
            bool rok[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = m0 + r0 + i;
                rok[i] = row < p.M;
                const int rr = rok[i] ? row : 0;
                img[i] = rr / (g.Ho * g.Wo);
                const int rem = rr - img[i] * (g.Ho * g.Wo);
                const int oy = rem / g.Wo;
                iy0[i] = oy * g.stride - g.pad;
                ix0[i] = (rem - oy * g.Wo) * g.stride - g.pad;
            }
            for (int kc = tc.k0; kc < tc.k1; ++kc, ++kt) {
                const int s = kt % C::kRawSlots;
                const uint32_t ph = (kt / C::kRawSlots) & 1;
                mbar_wait(&empty_x[s], ph ^ 1);
                uint8_t *slot = a_hi(s);
                for (int sg = 0; sg < n_seg; ++sg) {
                    const int kk = kc * kBK + sg * seg_len;
                    const int tap = kk / Cin, c = kk - tap * Cin;
                    const int ky = tap / g.k, kx = tap - ky * g.k;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        int y = iy0[i] + ky, x = ix0[i] + kx;
                        bool ok = rok[i] && tap < taps;
                        if (g.reflect) {
                            y = y < 0 ? -y : (y >= g.H ? 2 * g.H - 2 - y : y);
                            x = x < 0 ? -x : (x >= g.W ? 2 * g.W - 2 - x : x);
                        } else {
                            ok = ok && y >= 0 && y < g.H && x >= 0 && x < g.W;
                        }
                        const float *src = g.a;
                        if (ok) {
                            if (c < g.Ca) {
                                int ya = y, xa = x;
                                if (resample) {
                                    ya = min(static_cast<int>(floorf(static_cast<float>(y) * g.scale_h)), g.Ha - 1);
                                    xa = min(static_cast<int>(floorf(static_cast<float>(x) * g.scale_w)), g.Wa - 1);
                                }
                                src = g.a + (static_cast<size_t>(img[i] * g.Ha + ya) * g.Wa + xa) * g.lda + c;
                            } else {
                                src = g.b + (static_cast<size_t>(img[i] * g.H + y) * g.W + x) * g.ldb + (c - g.Ca);
                            }
                        }
                        const int r = r0 + i;
                        uint8_t *drow = slot + r * 128;
                        const uint32_t nbytes = ok ? 16u : 0u;
                        for (int pc = 0; pc < seg_len / 4; ++pc) {
                            const int piece = sg * (seg_len / 4) + pc;
                            cp_async16_zfill(drow + ((piece ^ (r & 7)) << 4), src + pc * 4, nbytes);
                        }
                    }
                }
                cp_async_mbar_arrive_noinc(&full_x[s]);
            }
        }
    } else if (warp == 0) {
        // ===== X producer (TMA): runs up to kRawSlots k-chunks ahead of the tensor core =====
        if (lane == 0) {
            int kt = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const TileCoord tc = tile_coord<BN>(p, tile, nk);
                const int m0 = tc.m0;
                for (int kc = tc.k0; kc < tc.k1; ++kc, ++kt) {
                    const int s = kt % C::kRawSlots;
                    const uint32_t ph = (kt / C::kRawSlots) & 1;
                    mbar_wait(&empty_x[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full_x[s], kATileBytes);
                    tma_load_2d(a_hi(s), &mapA, kc * kBK, m0, &full_x[s]);
                }
            }
        }
    } else if (warp == 2) {
        // ===== W producer (TMA): weight k-chunks come from L2 =====
        if (lane == 0) {
            int kt = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const TileCoord tc = tile_coord<BN>(p, tile, nk);
                const int n0 = tc.n0;
                for (int kc = tc.k0; kc < tc.k1; ++kc, ++kt) {
                    const int s = kt & 1;
                    const uint32_t ph = (kt >> 1) & 1;
                    mbar_wait(&empty_w[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full_w[s], C::kWTxBytes);
                    tma_load_2d(b_hi(s), &mapBhi, kc * kBK, n0, &full_w[s]);
                    if (SPLIT) tma_load_2d(b_lo(s), &mapBlo, kc * kBK, n0, &full_w[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(kBM, BN);
            int kt = 0, it = 0;
#ifdef MAC_LINEAR_PROFILE
            long long w_acc = 0, w_x = 0, w_w = 0, t_role = clock64(), t_;
#define MAC_PROF_WAIT(ACC, STMT) t_ = clock64(); STMT; ACC += clock64() - t_
#else
#define MAC_PROF_WAIT(ACC, STMT) STMT
#endif
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const int ab = C::kAccBufs == 2 ? (it & 1) : 0;
                const uint32_t aph = C::kAccBufs == 2 ? ((it >> 1) & 1) : (it & 1);
                MAC_PROF_WAIT(w_acc, mbar_wait(&acc_empty[ab], aph ^ 1));
                tc_fence_after_sync();
                const uint32_t tacc = tmem_base + ab * BN;
                const TileCoord tc = tile_coord<BN>(p, tile, nk);
                for (int kc = tc.k0; kc < tc.k1; ++kc, ++kt) {
                    const int sx = kt % C::kRawSlots, s2 = kt & 1;
                    const uint32_t phx = (kt / C::kRawSlots) & 1, ph2 = (kt >> 1) & 1;
                    if (SPLIT) { MAC_PROF_WAIT(w_x, mbar_wait(&ready_lo[s2], ph2)); }
                    else { MAC_PROF_WAIT(w_x, mbar_wait(&full_x[sx], phx)); }
                    MAC_PROF_WAIT(w_w, mbar_wait(&full_w[s2], ph2));
                    tc_fence_after_sync();
                    const uint32_t ah = smem_u32(a_hi(sx)), bh = smem_u32(b_hi(s2));
                    const uint32_t bl = smem_u32(b_lo(s2));
                    const uint32_t ta_hi = tmem_base + C::kAccBufs * BN + s2 * 64, ta_lo = ta_hi + 32;   // A slot s2 in tensor memory
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k) {
                        const uint32_t off = k * kUmmaK * 4;
                        if (C::kATmem) {
                            umma_tf32_ts(tacc, ta_lo + k * kUmmaK, umma_desc_k_sw128(bh + off), idesc, kc != tc.k0 || k != 0);
                            umma_tf32_ts(tacc, ta_hi + k * kUmmaK, umma_desc_k_sw128(bl + off), idesc, 1);
                            umma_tf32_ts(tacc, ta_hi + k * kUmmaK, umma_desc_k_sw128(bh + off), idesc, 1);
                        } else {
                            umma_tf32(tacc, umma_desc_k_sw128(ah + off), umma_desc_k_sw128(bh + off), idesc, kc != tc.k0 || k != 0);
                        }
                    }
                    if (!C::kATmem) umma_commit(&empty_x[sx]);
                    umma_commit(&empty_w[s2]);
                    if (SPLIT) umma_commit(&empty_lo[s2]);
                }
                umma_commit(&acc_full[ab]);
            }
#ifdef MAC_LINEAR_PROFILE
            atomicAdd(&g_linear_prof[0], static_cast<unsigned long long>(w_acc));
            atomicAdd(&g_linear_prof[1], static_cast<unsigned long long>(w_x));
            atomicAdd(&g_linear_prof[2], static_cast<unsigned long long>(w_w));
            atomicAdd(&g_linear_prof[3], static_cast<unsigned long long>(clock64() - t_role));
            atomicAdd(&g_linear_prof[4], static_cast<unsigned long long>(it));
#endif
#undef MAC_PROF_WAIT
        }
    } else if (warp >= 4 && warp < 8) {
        // ===== X split: x -> (tf32(x), tf32(x - tf32(x))) =====
        if (C::kATmem) {
            // one row of the tile per thread (warp q of the four owns TMEM lanes 32 q .. 32 q + 31): 8 conflict-free LDS.128
            // un-swizzle the row, the halves go to tensor memory, then the raw slot returns to the producer
            const int r = threadIdx.x - 128;   // 0..127 = tile row = TMEM lane
            const uint32_t lane_addr = static_cast<uint32_t>(r & ~31) << 16;
            const bool lnin = p.lnin_stats != nullptr;
            float *kvec = svec + 3 * BN + 4 * 128;   // [2][kLnInMaxK]: gamma | beta over K
            if (lnin) {
                for (int c = r; c < C::kLnInMaxK; c += 128) {
                    kvec[c] = c < p.K ? p.lnin_g[c] : 0.f;
                    kvec[C::kLnInMaxK + c] = c < p.K ? p.lnin_b[c] : 0.f;
                }
                named_bar_sync(8, 128);   // the 4 split warps only (ids 1-3 belong to the epilogue groups)
            }
            int kt = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const TileCoord tc = tile_coord<BN>(p, tile, nk);
                float mean = 0.f, rstd = 0.f;
                if (lnin && tc.m0 + r < p.M) {
                    const float2 st = __ldg(reinterpret_cast<const float2 *>(p.lnin_stats) + tc.m0 + r);
                    mean = st.x, rstd = st.y;
                }
                for (int kc = tc.k0; kc < tc.k1; ++kc, ++kt) {
                    const int sx = kt % C::kRawSlots, s2 = kt & 1;
                    const uint32_t phx = (kt / C::kRawSlots) & 1, ph2 = (kt >> 1) & 1;
                    mbar_wait(&full_x[sx], phx);
                    float v[32], h[32];
                    const uint8_t *row = a_hi(sx) + r * 128;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 x4 = *reinterpret_cast<const float4 *>(row + ((q ^ (r & 7)) << 4));
                        v[4 * q] = x4.x, v[4 * q + 1] = x4.y, v[4 * q + 2] = x4.z, v[4 * q + 3] = x4.w;
                    }
                    if (lnin) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 g4 = *reinterpret_cast<const float4 *>(kvec + kc * kBK + 4 * q);
                            const float4 b4 = *reinterpret_cast<const float4 *>(kvec + C::kLnInMaxK + kc * kBK + 4 * q);
                            v[4 * q] = fmaf((v[4 * q] - mean) * rstd, g4.x, b4.x);
                            v[4 * q + 1] = fmaf((v[4 * q + 1] - mean) * rstd, g4.y, b4.y);
                            v[4 * q + 2] = fmaf((v[4 * q + 2] - mean) * rstd, g4.z, b4.z);
                            v[4 * q + 3] = fmaf((v[4 * q + 3] - mean) * rstd, g4.w, b4.w);
                        }
                    }
                    if (p.act_in == MAC_LIN_GELU) {   // the producer layer stored its pre-activation (its epilogue was GELU-bound)
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = gelu_exact(v[j]);
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        h[j] = tf32_hi(v[j]);
                        v[j] = tf32_lo(v[j], h[j]);
                    }
                    mbar_wait(&empty_lo[s2], ph2 ^ 1);   // the UMMAs that read A slot s2 two chunks ago have completed
                    tc_fence_after_sync();
                    const uint32_t ta = tmem_base + C::kAccBufs * BN + s2 * 64 + lane_addr;
                    tmem_st32(ta, h);
                    tmem_st32(ta + 32, v);
                    // every value read from the raw slot has been consumed by the two stores above (volatile, in program
                    // order), so the loads have completed: order them before the async-proxy refill and release the slot
                    fence_proxy_async_smem();
                    mbar_arrive(&empty_x[sx]);
                    tc_fence_before_sync();
                    mbar_arrive(&ready_lo[s2]);
                }
            }
        }
    } else if (warp < 8) {
        // warp 3: idle (keeps the epilogue warps aligned to TMEM lane quarters)
    } else {
        // ===== epilogue: TMEM -> registers -> (bias, activation, residual, LayerNorm, pool) -> staged, coalesced stores =====
        // kEpiGroups groups of 4 warps; group g owns staging buffer g and the 32-column chunks g, g + G, ... of a tile.
        constexpr int G = C::kEpiBufs;
        const int g = (warp - 8) >> 2;          // group of this warp
        const int t = threadIdx.x - 256 - g * 128;  // 0..127 within the group
        const int q = warp & 3;                 // TMEM lane quarter this warp may read
        const int rl = q * 32 + lane;           // row of the tile owned by this thread
        const int gbar = kEpiBar + g;           // named barrier of this group
        float *mybuf = ebuf + g * (C::kEpiBufBytes / 4);
        float *red = svec + 3 * BN;             // [G][128] cross-group partial row sums, then [G][128] partial centred squares: separate
        float *red_var = red + 2 * 128;         // arrays, so that neither needs a barrier before it is rewritten for the next tile
        float v[32];
        int it = 0;
        // TMA-store path: staging tile (cb & 1) of this group holds chunk number cb of the group's chunk sequence; the
        // residual of the NEXT chunk is copied into the other tile while the current one is computed (res_inflight)
        const bool tma_out = C::kTmaOut && p.tma_out != 0;
        const bool fixed_cols = p.n_tiles_n == 1;   // the column vectors (bias, gamma, beta) are the same for every tile
        uint8_t *const stage = reinterpret_cast<uint8_t *>(mybuf);
        int cb = 0;
        bool res_inflight = false;
        // residual chunk c of the tile at (rm0, rn0) -> swizzled staging tile, 16-byte async copies, coalesced
        auto fetch_res_sw = [&](int rm0, int rn0, int c, uint8_t *dst) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int piece = t + i * 128, r = piece >> 3, sg = piece & 7;
                const int col = rn0 + c * 32 + sg * 4;
                if (rm0 + r < p.M && col < p.N)
                    cp_async16(dst + r * 128 + ((sg ^ (r & 7)) << 4), p.res + static_cast<size_t>(rm0 + r) * p.ldr + col);
            }
            cp_async_commit();
        };
        if (tma_out && p.res && blockIdx.x < n_tiles) {
            const TileCoord t0 = tile_coord<BN>(p, blockIdx.x, nk);
            if (g < (min(BN, p.N - t0.n0) + 31) / 32) {
                fetch_res_sw(t0.m0, t0.n0, g, stage);
                res_inflight = true;
            }
        }
#ifdef MAC_LINEAR_PROFILE
        // epilogue phases of group 0 (thread 0): [5] tile prologue (syncs + column vectors), [6] wait for the accumulator,
        // [7] chunk loop (residual wait, TMEM load, math, staging, stores), [8] row statistics / LayerNorm passes
        long long e_pro = 0, e_acc = 0, e_chunks = 0, e_stats = 0, e_t = 0;
#define MAC_EPI_MARK(ACC) if (g == 0 && t == 0) { const long long now_ = clock64(); ACC += now_ - e_t; e_t = now_; }
#else
#define MAC_EPI_MARK(ACC)
#endif
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
#ifdef MAC_LINEAR_PROFILE
            if (g == 0 && t == 0) e_t = clock64();
#endif
            const TileCoord tc = tile_coord<BN>(p, tile, nk);
            const int m0 = tc.m0, n0 = tc.n0;
            float *const out_base = p.out ? p.out + tc.sp * p.split_stride : nullptr;   // split-K: this split's partial tile
            const int ab = C::kAccBufs == 2 ? (it & 1) : 0;
            const uint32_t aph = C::kAccBufs == 2 ? ((it >> 1) & 1) : (it & 1);
            const int row = m0 + rl;
            const bool row_ok = row < p.M;
            const int ncols = min(BN, p.N - n0);          // valid columns of this tile
            const int nch = (ncols + 31) / 32;
            const uint32_t taddr = tmem_base + ab * BN + (static_cast<uint32_t>(q * 32) << 16);

            // residual chunk c (128 rows x 32 columns) -> staging buffer, 16-byte async copies, coalesced
            auto fetch_res = [&](int c) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int piece = t + i * 128, r = piece >> 3, sg = piece & 7;
                    const int col = n0 + c * 32 + sg * 4;
                    if (m0 + r < p.M && col < p.N)
                        cp_async16(mybuf + r * kEpiLd + sg * 4, p.res + static_cast<size_t>(m0 + r) * p.ldr + col);
                }
                cp_async_commit();
            };
            // staging buffer -> global, coalesced (8 threads cover one 128-byte row segment)
            auto store_chunk = [&](int c, float *base, int ld) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int piece = t + i * 128, r = piece >> 3, sg = piece & 7;
                    const int col = n0 + c * 32 + sg * 4;
                    if (m0 + r < p.M && col < p.N) {
                        const float4 y = *reinterpret_cast<const float4 *>(mybuf + r * kEpiLd + sg * 4);
                        float *dst = base + static_cast<size_t>(m0 + r) * ld + col;
                        if (col + 3 < p.N) *reinterpret_cast<float4 *>(dst) = y;
                        else {
                            dst[0] = y.x;
                            if (col + 1 < p.N) dst[1] = y.y;
                            if (col + 2 < p.N) dst[2] = y.z;
                        }
                    }
                }
            };
            auto all_groups_sync = [&]() { named_bar_sync(kEpiBar + G, 128 * G); };

            // per-column vectors of this tile -> shared memory (broadcast float4 reads instead of one LDG per element)
            // (with fixed column vectors the groups run decoupled: their staging tiles are their own, the partial-sum arrays
            // are ordered by the barriers of the statistics passes, the accumulator hand-over counts both groups' arrivals)
            if (!fixed_cols || it == 0) all_groups_sync();  // previous tile: vectors no longer read
            if (!fixed_cols || it == 0) {
                for (int c = t + g * 128; c < BN; c += 128 * G) {
                    const bool ok = c < ncols;
                    svec[c] = (ok && p.bias) ? p.bias[n0 + c] : 0.f;
                    svec[BN + c] = (ok && p.ln_out) ? p.ln_g[n0 + c] : 0.f;
                    svec[2 * BN + c] = (ok && p.ln_out) ? p.ln_b[n0 + c] : 0.f;
                }
            }
            if (!tma_out && p.res && g < nch) fetch_res(g);   // overlaps with the main loop of this tile
            if (!fixed_cols || it == 0) all_groups_sync();
            MAC_EPI_MARK(e_pro);
            mbar_wait(&acc_full[ab], aph);
            tc_fence_after_sync();
            MAC_EPI_MARK(e_acc);
            if (p.pool) {
                // groups of 16 consecutive rows -> out[row/16, col] = max, out[row/16, N + col] = mean
                const int grp = row >> 4, gl = lane & 15;
                for (int c = g; c < nch; c += G) {
                    tmem_ld32(taddr + c * 32, v);
                    float mx[2] = {0.f, 0.f}, sm[2] = {0.f, 0.f};
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float y = apply_act<ACT>(v[j] + svec[c * 32 + j]);
                        float a = y, b = y;
#pragma unroll
                        for (int d = 8; d >= 1; d >>= 1) {
                            a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, d));
                            b += __shfl_xor_sync(0xffffffffu, b, d);
                        }
                        if ((j & 15) == gl) mx[j >> 4] = a, sm[j >> 4] = b;
                    }
                    if (row_ok) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int col = n0 + c * 32 + h * 16 + gl;
                            if (col < p.N) {
                                p.out[static_cast<size_t>(grp) * p.ldo + col] = mx[h];
                                p.out[static_cast<size_t>(grp) * p.ldo + p.N + col] = sm[h] * (1.0f / 16.0f);
                            }
                        }
                    }
                }
            } else {
                float *mine = mybuf + rl * kEpiLd;
                const bool has_res = p.res != nullptr;
                const float pre = p.res_first ? 1.f : 0.f, post = 1.f - pre;   // where the residual enters
                float sum = 0.f;
                for (int c = g; tma_out && c < nch; c += G, ++cb) {
                    uint8_t *const sb = stage + (cb % C::kStageTiles) * (kBM * 128);
                    tmem_ld32(taddr + c * 32, v);
                    if (t == 0) bulk_wait_read_all();   // the store of the previous chunk has read its staging tile
                    if (has_res && !res_inflight) {     // one staging tile, or after a tile this group had no chunk of
                        if (C::kStageTiles == 1) named_bar_sync(gbar, 128);   // the tile is this chunk's own: wait for its last store
                        fetch_res_sw(m0, n0, c, sb);
                    }
                    if (has_res) cp_async_wait_all();
                    named_bar_sync(gbar, 128);          // residual chunk visible to the group; the other tile is free
                    res_inflight = false;
                    if (has_res && C::kStageTiles == 2) {   // residual of the group's next chunk -> the other tile, while this one is computed
                        int c2 = c + G, tile2 = tile;
                        if (c2 >= nch) c2 = g, tile2 = tile + gridDim.x;
                        if (tile2 < n_tiles) {
                            const TileCoord t2 = tile_coord<BN>(p, tile2, nk);
                            if (c2 < (min(BN, p.N - t2.n0) + 31) / 32) {
                                fetch_res_sw(t2.m0, t2.n0, c2, stage + ((cb + 1) & 1) * (kBM * 128));
                                res_inflight = true;
                            }
                        }
                    }
                    uint8_t *const rowp = sb + rl * 128;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 *const ptr = reinterpret_cast<float4 *>(rowp + (((j >> 2) ^ (rl & 7)) << 4));
                        const float4 b4 = *reinterpret_cast<const float4 *>(svec + c * 32 + j);
                        float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (has_res) r4 = *ptr;
                        float4 y4;
                        y4.x = apply_act<ACT>(v[j] + b4.x + pre * r4.x) + post * r4.x;
                        y4.y = apply_act<ACT>(v[j + 1] + b4.y + pre * r4.y) + post * r4.y;
                        y4.z = apply_act<ACT>(v[j + 2] + b4.z + pre * r4.z) + post * r4.z;
                        y4.w = apply_act<ACT>(v[j + 3] + b4.w + pre * r4.w) + post * r4.w;
                        if (c * 32 + j + 3 >= ncols) {  // ragged right edge: columns >= N contribute nothing
                            if (c * 32 + j >= ncols) y4.x = 0.f;
                            if (c * 32 + j + 1 >= ncols) y4.y = 0.f;
                            if (c * 32 + j + 2 >= ncols) y4.z = 0.f;
                            y4.w = 0.f;
                        }
                        v[j] = y4.x, v[j + 1] = y4.y, v[j + 2] = y4.z, v[j + 3] = y4.w;
                        sum += (y4.x + y4.y) + (y4.z + y4.w);
                        *ptr = y4;
                    }
                    if (p.stats_out) tmem_st32(taddr + c * 32, v);   // keep y for the row statistics
                    fence_proxy_async_smem();           // generic writes of the tile -> visible to the TMA engine
                    named_bar_sync(gbar, 128);          // the 128 x 32 chunk of y is staged
                    if (t == 0) {
                        tma_store_2d(&mapOut, sb, n0 + c * 32, m0);   // rows >= M / columns >= N are dropped by the TMA unit
                        bulk_commit_group();
                    }
                }
                for (int c = g; !tma_out && c < nch; c += G) {
                    if (has_res) {
                        cp_async_wait_all();
                        named_bar_sync(gbar, 128);  // residual chunk c visible to the whole group
                    }
                    tmem_ld32(taddr + c * 32, v);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b4 = *reinterpret_cast<const float4 *>(svec + c * 32 + j);
                        float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (has_res) r4 = *reinterpret_cast<const float4 *>(mine + j);
                        float4 y4;
                        y4.x = apply_act<ACT>(v[j] + b4.x + pre * r4.x) + post * r4.x;
                        y4.y = apply_act<ACT>(v[j + 1] + b4.y + pre * r4.y) + post * r4.y;
                        y4.z = apply_act<ACT>(v[j + 2] + b4.z + pre * r4.z) + post * r4.z;
                        y4.w = apply_act<ACT>(v[j + 3] + b4.w + pre * r4.w) + post * r4.w;
                        if (c * 32 + j + 3 >= ncols) {  // ragged right edge: columns >= N contribute nothing
                            if (c * 32 + j >= ncols) y4.x = 0.f;
                            if (c * 32 + j + 1 >= ncols) y4.y = 0.f;
                            if (c * 32 + j + 2 >= ncols) y4.z = 0.f;
                            y4.w = 0.f;
                        }
                        v[j] = y4.x, v[j + 1] = y4.y, v[j + 2] = y4.z, v[j + 3] = y4.w;
                        sum += (y4.x + y4.y) + (y4.z + y4.w);
                        *reinterpret_cast<float4 *>(mine + j) = y4;
                    }
                    if (p.ln_out || p.stats_out) tmem_st32(taddr + c * 32, v);  // keep y for the LayerNorm passes
                    named_bar_sync(gbar, 128);                   // the 128 x 32 chunk of y is staged
                    if (p.out) store_chunk(c, out_base, p.ldo);
                    named_bar_sync(gbar, 128);                   // staging buffer drained
                    if (has_res && c + G < nch) fetch_res(c + G);
                }
                MAC_EPI_MARK(e_chunks);
                if (p.ln_out || p.stats_out) {
                    // LayerNorm over the N columns of this row (the tile spans all of N): mean, then centred variance;
                    // the groups hold disjoint column chunks and combine their partial sums through shared memory
                    if (G > 1) {
                        red[g * 128 + rl] = sum;
                        all_groups_sync();
                        sum = red[rl] + red[128 + rl];
                    }
                    const float mean = sum / static_cast<float>(p.N);
                    float var = 0.f;
                    for (int c = g; c < nch; c += G) {
                        tmem_ld32(taddr + c * 32, v);
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float d = (c * 32 + j < ncols) ? v[j] - mean : 0.f;
                            var = fmaf(d, d, var);
                        }
                    }
                    if (G > 1) {
                        red_var[g * 128 + rl] = var;
                        all_groups_sync();
                        var = red_var[rl] + red_var[128 + rl];
                    }
                    const float rstd = 1.0f / sqrtf(var / static_cast<float>(p.N) + p.ln_eps);
                    if (p.stats_out && g == 0 && row_ok)   // the next layer normalises on load: 8 bytes instead of a row
                        *reinterpret_cast<float2 *>(p.stats_out + 2 * static_cast<size_t>(row)) = make_float2(mean, rstd);
                    for (int c = g; p.ln_out && c < nch; c += G) {
                        tmem_ld32(taddr + c * 32, v);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 g4 = *reinterpret_cast<const float4 *>(svec + BN + c * 32 + j);
                            const float4 b4 = *reinterpret_cast<const float4 *>(svec + 2 * BN + c * 32 + j);
                            float4 y4;
                            y4.x = (v[j] - mean) * rstd * g4.x + b4.x;
                            y4.y = (v[j + 1] - mean) * rstd * g4.y + b4.y;
                            y4.z = (v[j + 2] - mean) * rstd * g4.z + b4.z;
                            y4.w = (v[j + 3] - mean) * rstd * g4.w + b4.w;
                            *reinterpret_cast<float4 *>(mine + j) = y4;
                        }
                        named_bar_sync(gbar, 128);
                        store_chunk(c, p.ln_out, p.ldl);
                        named_bar_sync(gbar, 128);
                    }
                }
            }
            MAC_EPI_MARK(e_stats);
            tc_fence_before_sync();
            mbar_arrive(&acc_empty[ab]);
        }
        if (tma_out && t == 0) bulk_wait_read_all();   // the staging tiles must outlive the last stores' reads
#ifdef MAC_LINEAR_PROFILE
        if (g == 0 && t == 0) {
            atomicAdd(&g_linear_prof[5], static_cast<unsigned long long>(e_pro));
            atomicAdd(&g_linear_prof[6], static_cast<unsigned long long>(e_acc));
            atomicAdd(&g_linear_prof[7], static_cast<unsigned long long>(e_chunks));
            atomicAdd(&g_linear_prof[8], static_cast<unsigned long long>(e_stats));
        }
#endif
#undef MAC_EPI_MARK
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <int BN, bool SPLIT, int ACT, bool GATHER = false>
int launch_act(const CUtensorMap &mapA, const CUtensorMap &mapBhi, const CUtensorMap &mapBlo, LinearParams &p,
               cudaStream_t stream)
{
    // `out` through TMA stores whenever it is a plain (M, N) matrix written once: not pooled, not a split-K partial, no
    // second (LayerNorm-ed) output
    static const int no_tma_out = [] { const char *e = getenv("MAC_LINEAR_NO_TMA_STORE"); return e ? atoi(e) : 0; }();   // A/B knob
    CUtensorMap mapOut = mapBhi;   // placeholder when unused
    p.tma_out = 0;
    if (Cfg<BN, SPLIT>::kTmaOut && !no_tma_out && p.out && !p.pool && !p.ln_out && p.splits <= 1 && p.ldo % 4 == 0 &&
        (reinterpret_cast<uintptr_t>(p.out) & 15u) == 0) {
        if (int rc = make_tensor_map_2d(&mapOut, p.out, p.M, p.N, p.ldo, kBM)) return rc;
        p.tma_out = 1;
    }
    using C = Cfg<BN, SPLIT>;
    static DeviceOnce once;
    if (int rc = ensure_dynamic_smem(once, linear_kernel<BN, SPLIT, ACT, GATHER>, C::kSmemBytes)) return rc;
    p.n_tiles_m = (p.M + kBM - 1) / kBM;
    p.n_tiles_n = (p.N + BN - 1) / BN;
    int device = 0;
    MAC_CUDA(cudaGetDevice(&device));
    if (p.splits < 1) p.splits = 1, p.chunks_per_split = (p.K + kBK - 1) / kBK, p.split_stride = 0;
    const int n_tiles = p.n_tiles_m * p.n_tiles_n * p.splits;
    const int grid = n_tiles < sm_count(device) ? n_tiles : sm_count(device);
    linear_kernel<BN, SPLIT, ACT, GATHER><<<grid, C::kThreads, C::kSmemBytes, stream>>>(mapA, mapBhi, mapBlo, mapOut, p);
    MAC_CUDA(cudaGetLastError());
    count_launch();
    return MAC_OK;
}

template <int BN, bool SPLIT>
int launch(const CUtensorMap &mapA, const CUtensorMap &mapBhi, const CUtensorMap &mapBlo, LinearParams &p,
           cudaStream_t stream)
{
    if (p.act == MAC_LIN_GELU) return launch_act<BN, SPLIT, MAC_LIN_GELU>(mapA, mapBhi, mapBlo, p, stream);
    if (p.act == MAC_LIN_RELU) return launch_act<BN, SPLIT, MAC_LIN_RELU>(mapA, mapBhi, mapBlo, p, stream);
    if (p.act == MAC_LIN_ELU) return launch_act<BN, SPLIT, MAC_LIN_ELU>(mapA, mapBhi, mapBlo, p, stream);
    if (p.act == MAC_LIN_SIGMOID) return launch_act<BN, SPLIT, MAC_LIN_SIGMOID>(mapA, mapBhi, mapBlo, p, stream);
    return launch_act<BN, SPLIT, MAC_LIN_NONE>(mapA, mapBhi, mapBlo, p, stream);
}

// convolution epilogues of the depth network: ReLU (encoder), ELU (decoder), sigmoid (disparity heads), none
template <int BN>
int launch_gather(const CUtensorMap &mapBhi, const CUtensorMap &mapBlo, LinearParams &p, cudaStream_t stream)
{
    if (p.act == MAC_LIN_RELU) return launch_act<BN, true, MAC_LIN_RELU, true>(mapBhi, mapBhi, mapBlo, p, stream);
    if (p.act == MAC_LIN_ELU) return launch_act<BN, true, MAC_LIN_ELU, true>(mapBhi, mapBhi, mapBlo, p, stream);
    if (p.act == MAC_LIN_SIGMOID) return launch_act<BN, true, MAC_LIN_SIGMOID, true>(mapBhi, mapBhi, mapBlo, p, stream);
    return launch_act<BN, true, MAC_LIN_NONE, true>(mapBhi, mapBhi, mapBlo, p, stream);
}

// Second pass of a split-K layer: out = epilogue(sum over splits of the raw partial tiles), summed in split order
// (deterministic).  One thread per 4 columns.
template <int ACT>
__global__ void __launch_bounds__(256) splitk_finish_kernel(const float *__restrict__ part, int splits, long long stride, int ldp,
                                                            int M, int N, const float *__restrict__ bias,
                                                            const float *__restrict__ res, int ldr, int res_first,
                                                            float *__restrict__ out, int ldo)
{
    const int n4 = N / 4;
    const long long total = static_cast<long long>(M) * n4;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
        const int row = static_cast<int>(i / n4), col = static_cast<int>(i - static_cast<long long>(row) * n4) * 4;
        const float *src = part + static_cast<size_t>(row) * ldp + col;
        float4 a = *reinterpret_cast<const float4 *>(src);
        for (int sp = 1; sp < splits; ++sp) {
            const float4 b = *reinterpret_cast<const float4 *>(src + sp * stride);
            a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
        }
        if (bias) {
            const float4 b = *reinterpret_cast<const float4 *>(bias + col);
            a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
        }
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (res) r = *reinterpret_cast<const float4 *>(res + static_cast<size_t>(row) * ldr + col);
        const float pre = res_first ? 1.f : 0.f, post = 1.f - pre;
        float4 y;
        y.x = apply_act<ACT>(a.x + pre * r.x) + post * r.x;
        y.y = apply_act<ACT>(a.y + pre * r.y) + post * r.y;
        y.z = apply_act<ACT>(a.z + pre * r.z) + post * r.z;
        y.w = apply_act<ACT>(a.w + pre * r.w) + post * r.w;
        *reinterpret_cast<float4 *>(out + static_cast<size_t>(row) * ldo + col) = y;
    }
}

}  // namespace

bool conv_gather_supported(const ConvGather &g)
{
    const bool aligned = (reinterpret_cast<uintptr_t>(g.a) & 15u) == 0 && g.lda % 4 == 0 &&
                         (!g.b || ((reinterpret_cast<uintptr_t>(g.b) & 15u) == 0 && g.ldb % 4 == 0));
    if (!aligned || g.k < 1 || g.Ca <= 0) return false;
    if (g.Ca % kBK == 0 && g.Cb % kBK == 0) return true;   // every k-chunk inside one tap of one source
    return g.Cb == 0 && g.Ca == 16;                        // two taps per k-chunk
}

int linear_forward_conv(const ConvGather &g, const float *W_hi, const float *W_lo, int ldw, const float *bias, float *out,
                        int ldo, int M, int N, int K, int act, const float *res, int ldr, cudaStream_t stream, int res_first,
                        float *ws, size_t ws_floats)
{
    MAC_REQUIRE(g.a && W_hi && W_lo && out, "null tensor pointer");
    MAC_REQUIRE(conv_gather_supported(g), "convolution not supported by the gather producer (Ca=%d Cb=%d)", g.Ca, g.Cb);
    MAC_REQUIRE(K == g.k * g.k * (g.Ca + g.Cb), "K=%d does not match the convolution", K);
    MAC_REQUIRE(M > 0 && N > 0 && act >= MAC_LIN_NONE && act <= MAC_LIN_SIGMOID && act != MAC_LIN_GELU, "bad shape or activation");
    MAC_REQUIRE(ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0, "out must be 16-byte aligned with ldo %% 4 == 0");
    MAC_REQUIRE(!res || (N % 4 == 0 && ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(res) & 15u) == 0),
                "the residual needs N %% 4 == 0 and a 16-byte aligned base and row stride");
    const int bn = N <= 64 ? 64 : 128;
    CUtensorMap mapBhi, mapBlo;
    if (int rc = make_tensor_map_2d(&mapBhi, W_hi, N, K, ldw, bn)) return rc;
    if (int rc = make_tensor_map_2d(&mapBlo, W_lo, N, K, ldw, bn)) return rc;
    LinearParams p{};
    p.M = M, p.N = N, p.K = K;
    p.bias = bias;
    p.out = out, p.ldo = ldo;
    p.res = res, p.ldr = ldr;
    p.act = act, p.res_first = res_first;
    p.conv = g;

    // Split-K for the low-resolution layers: a handful of 128-row tiles with thousands of k-chunks would stream the
    // (large) weight matrix through 2-8 SMs; with the chunks divided over ~one CTA per SM every SM pulls its own weight columns.
    int device = 0;
    MAC_CUDA(cudaGetDevice(&device));
    const int sms = sm_count(device);
    const int tiles_mn = ((M + kBM - 1) / kBM) * ((N + bn - 1) / bn);
    const int nk = (K + kBK - 1) / kBK;
    static const int no_split = [] { const char *e = getenv("MAC_LINEAR_NO_SPLITK"); return e ? atoi(e) : 0; }();   // A/B timing knob
    int want = sms / tiles_mn < nk / 4 ? sms / tiles_mn : nk / 4;
    if (ws && !no_split && N % 4 == 0 && (reinterpret_cast<uintptr_t>(ws) & 15u) == 0 && want >= 2) {
        const int cps = (nk + want - 1) / want;
        const int splits = (nk + cps - 1) / cps;        // every split has at least one chunk
        const size_t need = static_cast<size_t>(splits) * M * N;
        if (splits >= 2 && need <= ws_floats) {
            p.splits = splits, p.chunks_per_split = cps, p.split_stride = static_cast<long long>(M) * N;
            p.out = ws, p.ldo = N;
            p.bias = nullptr, p.res = nullptr, p.act = MAC_LIN_NONE;
            if (int rc = bn == 64 ? launch_gather<64>(mapBhi, mapBlo, p, stream) : launch_gather<128>(mapBhi, mapBlo, p, stream))
                return rc;
            const long long total = static_cast<long long>(M) * (N / 4);
            const int grid = static_cast<int>((total + 255) / 256 < 4 * sms ? (total + 255) / 256 : 4 * sms);
#define MAC_FINISH(ACT)                                                                                                       \
    splitk_finish_kernel<ACT><<<grid, 256, 0, stream>>>(ws, splits, p.split_stride, N, M, N, bias, res, ldr, res_first, out, ldo)
            if (act == MAC_LIN_RELU) MAC_FINISH(MAC_LIN_RELU);
            else if (act == MAC_LIN_ELU) MAC_FINISH(MAC_LIN_ELU);
            else if (act == MAC_LIN_SIGMOID) MAC_FINISH(MAC_LIN_SIGMOID);
            else MAC_FINISH(MAC_LIN_NONE);
#undef MAC_FINISH
            MAC_CUDA(cudaGetLastError());
            count_launch();
            return MAC_OK;
        }
    }
    if (bn == 64) return launch_gather<64>(mapBhi, mapBlo, p, stream);
    return launch_gather<128>(mapBhi, mapBlo, p, stream);
}

int linear_forward(const float *X, int ldx, const float *W_hi, const float *W_lo, int ldw, const float *bias, float *out,
                   int ldo, int M, int N, int K, int act, const float *res, int ldr, float *ln_out, int ldl,
                   const float *ln_g, const float *ln_b, float ln_eps, int pool, cudaStream_t stream, int res_first,
                   const float *lnin_stats, const float *lnin_g, const float *lnin_b, float *stats_out, int act_in)
{
    MAC_REQUIRE(X && W_hi && (out || ln_out), "null tensor pointer");
    MAC_REQUIRE(act_in == MAC_LIN_NONE || (act_in == MAC_LIN_GELU && W_lo && !lnin_stats),
                "an activation on load needs split weights and no LayerNorm on load");
    MAC_REQUIRE(!lnin_stats || (W_lo && lnin_g && lnin_b && K <= 512 && K % 4 == 0 &&
                                (reinterpret_cast<uintptr_t>(lnin_stats) & 7u) == 0),
                "LayerNorm on load needs split weights, gamma / beta and K <= 512");
    MAC_REQUIRE(!stats_out || (N <= 256 && !pool && (reinterpret_cast<uintptr_t>(stats_out) & 7u) == 0),
                "row statistics need N <= 256 and no pooling");
    MAC_REQUIRE(M > 0 && N > 0 && K > 0, "M, N, K must be positive (got %d, %d, %d)", M, N, K);
    MAC_REQUIRE(act >= MAC_LIN_NONE && act <= MAC_LIN_SIGMOID, "bad activation %d", act);
    MAC_REQUIRE(pool == 0 || pool == 16, "pool must be 0 or 16");
    MAC_REQUIRE(!pool || (M % 16 == 0 && out && !ln_out), "pooling needs M %% 16 == 0 and no LayerNorm output");
    MAC_REQUIRE(!ln_out || (N <= 256 && ln_g && ln_b && ldl % 4 == 0), "fused LayerNorm needs N <= 256, gamma and beta");
    MAC_REQUIRE(!out || pool || (ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0),
                "out must be 16-byte aligned with ldo %% 4 == 0");
    const bool split = W_lo != nullptr;

    MAC_REQUIRE(!res || (N % 4 == 0 && ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(res) & 15u) == 0 && !pool),
                "the residual needs N %% 4 == 0, a 16-byte aligned base and row stride, and no pooling");
    MAC_REQUIRE(!ln_out || (reinterpret_cast<uintptr_t>(ln_out) & 15u) == 0, "ln_out must be 16-byte aligned");
    // 256-wide tiles only where a row-wise epilogue (LayerNorm) must see all of N; otherwise 128-wide tiles
    // (two per 256 columns re-read X from L2 but run two-deep epilogue staging and twice the CTAs).
    int bn;
    if (ln_out || pool || stats_out) bn = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
    else if (N > 128 && N <= 192 && !res && !stats_out) bn = 192;   // e.g. the fused q|k|v projection of the 128-wide transformers: one
                                                      // tile per row block instead of a full and a half-empty 128-wide one
    else bn = N <= 64 ? 64 : 128;
    static const int wide = [] { const char *e = getenv("MAC_LINEAR_WIDE"); return e ? atoi(e) : 0; }();   // tuning knob
    if (wide && !(ln_out || pool || stats_out) && N % 256 == 0) bn = 256;
    MAC_REQUIRE(!(ln_out || pool || stats_out) || N <= bn, "row-wise epilogues need the tile to span N");

    CUtensorMap mapA, mapBhi, mapBlo;
    if (int rc = make_tensor_map_2d(&mapA, X, M, K, ldx, kBM)) return rc;
    if (int rc = make_tensor_map_2d(&mapBhi, W_hi, N, K, ldw, bn)) return rc;
    if (int rc = make_tensor_map_2d(&mapBlo, split ? W_lo : W_hi, N, K, ldw, bn)) return rc;

    LinearParams p{};
    p.M = M, p.N = N, p.K = K;
    p.bias = bias;
    p.out = out, p.ldo = ldo;
    p.res = res, p.ldr = ldr;
    p.ln_out = ln_out, p.ldl = ldl, p.ln_g = ln_g, p.ln_b = ln_b, p.ln_eps = ln_eps;
    p.act = act, p.pool = pool, p.res_first = res_first;
    p.lnin_stats = lnin_stats, p.lnin_g = lnin_g, p.lnin_b = lnin_b, p.stats_out = stats_out;
    p.act_in = act_in;
    if (stats_out && !ln_out) p.ln_eps = ln_eps;

    if (split) {
        if (bn == 64) return launch<64, true>(mapA, mapBhi, mapBlo, p, stream);
        if (bn == 128) return launch<128, true>(mapA, mapBhi, mapBlo, p, stream);
        if (bn == 192) return launch<192, true>(mapA, mapBhi, mapBlo, p, stream);
        return launch<256, true>(mapA, mapBhi, mapBlo, p, stream);
    }
    if (bn == 64) return launch<64, false>(mapA, mapBhi, mapBlo, p, stream);
    if (bn == 128) return launch<128, false>(mapA, mapBhi, mapBlo, p, stream);
    if (bn == 192) return launch<192, false>(mapA, mapBhi, mapBlo, p, stream);
    return launch<256, false>(mapA, mapBhi, mapBlo, p, stream);
}

}  // namespace mac

#ifdef MAC_LINEAR_PROFILE
// debug build only: read and clear the wait counters of the UMMA thread (see g_linear_prof)
extern "C" int mac_linear_profile_read(unsigned long long *out12)
{
    MAC_CUDA(cudaDeviceSynchronize());
    MAC_CUDA(cudaMemcpyFromSymbol(out12, mac::g_linear_prof, sizeof(unsigned long long) * 12));
    unsigned long long zero[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    MAC_CUDA(cudaMemcpyToSymbol(mac::g_linear_prof, zero, sizeof(zero)));
    return MAC_OK;
}
#endif

extern "C" int mac_linear_f32(const float *X, int ldx, const float *W_hi, const float *W_lo, int ldw, const float *bias,
                              float *out, int ldo, int M, int N, int K, int act, const float *res, int ldr,
                              float *ln_out, int ldl, const float *ln_g, const float *ln_b, float ln_eps, int pool,
                              void *stream)
{
    return mac::linear_forward(X, ldx, W_hi, W_lo, ldw, bias, out, ldo, M, N, K, act, res, ldr, ln_out, ldl, ln_g, ln_b,
                               ln_eps, pool, static_cast<cudaStream_t>(stream));
}

// mac_linear_f32 with the LayerNorm of the INPUT rows applied on load (lnin_stats (M, 2) = (mean, rstd) per row,
// lnin_g / lnin_b (K)) and / or the (mean, rstd) of every OUTPUT row written to stats_out (M, 2) with ln_eps: a chain of
// layers then passes 8 bytes per row instead of a normalised copy of the activations.
extern "C" int mac_linear_lnio_f32(const float *X, int ldx, const float *W_hi, const float *W_lo, int ldw, const float *bias,
                                   float *out, int ldo, int M, int N, int K, int act, const float *res, int ldr,
                                   const float *lnin_stats, const float *lnin_g, const float *lnin_b, float *stats_out,
                                   float ln_eps, void *stream)
{
    return mac::linear_forward(X, ldx, W_hi, W_lo, ldw, bias, out, ldo, M, N, K, act, res, ldr, nullptr, 0, nullptr, nullptr,
                               ln_eps, 0, static_cast<cudaStream_t>(stream), 0, lnin_stats, lnin_g, lnin_b, stats_out);
}
