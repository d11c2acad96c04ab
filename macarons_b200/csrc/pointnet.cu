// CUDA-core kernels around the tensor-core linear layers of SconeOcc / SconeVis:
//   knn16           16 nearest neighbours of every query in a cloud     (reference utility/utils.py:1497-1509)
//   embed_first     first embedding layer (K = 3 or 4 inputs) + GELU, optional neighbour gather / offset and
//                   input concatenation                                  (networks/Attention.py:96-98,124-126;
//                                                                        networks/SconeOcc.py:293-296, 18-36)
//   attn16          self-attention over 16-token neighbourhoods         (networks/Attention.py:8-36,182-204)
//   attn_dense      self-attention over one cloud (<= a few thousand tokens), flash-style online softmax
//   colpool         max / mean over the tokens of a cloud               (networks/SconeOcc.py:120-127,
//                                                                        networks/Attention.py:110-114)
//   vis_embed_finish  [features | global max | input] concatenation + LayerNorm (networks/Attention.py:110-126)
//   bias_gemv       per-cloud bias = b + W_g . global_feature (folds the broadcast global feature of
//                   networks/SconeOcc.py:330-334 into the first head layer)
#include <float.h>
#include <math.h>

#include "tc_common.h"
#include "nets.h"

namespace mac {

namespace {

// GELU with the A&S 7.1.26 erf (|error| <= 1.5e-7), see linear.cu
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
    return fmaf(fabsf(h), fmaf(-p * t, e, 1.0f), h);
}

// ------------------------------------------------------------------------------------------------
// kNN: one thread per query, cloud tiled through shared memory, sorted top-16 kept in registers.
// ------------------------------------------------------------------------------------------------
constexpr int kKnn = 16;
constexpr int kKnnTile = 1024;

__global__ void __launch_bounds__(64) knn16_kernel(const float *__restrict__ x, const float *__restrict__ pc,
                                                    int *__restrict__ idx_out, float *__restrict__ dist_out, int Q, int N)
{
    __shared__ float sp[kKnnTile * 3];
    const int b = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = q < Q;
    const float *xb = x + (static_cast<size_t>(b) * Q + (live ? q : 0)) * 3;
    const float qx = xb[0], qy = xb[1], qz = xb[2];
    const float *pcb = pc + static_cast<size_t>(b) * N * 3;

    float bd[kKnn];
    int bi[kKnn];
#pragma unroll
    for (int i = 0; i < kKnn; ++i) bd[i] = FLT_MAX, bi[i] = 0;

    for (int t0 = 0; t0 < N; t0 += kKnnTile) {
        const int n = min(kKnnTile, N - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < n * 3; i += blockDim.x) sp[i] = pcb[static_cast<size_t>(t0) * 3 + i];
        __syncthreads();
        for (int j = 0; j < n; ++j) {
            const float dx = qx - sp[3 * j], dy = qy - sp[3 * j + 1], dz = qz - sp[3 * j + 2];
            const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            if (d < bd[kKnn - 1]) {
                bd[kKnn - 1] = d;
                bi[kKnn - 1] = t0 + j;
#pragma unroll
                for (int i = kKnn - 1; i > 0; --i) {
                    if (bd[i] < bd[i - 1]) {
                        const float td = bd[i];
                        bd[i] = bd[i - 1];
                        bd[i - 1] = td;
                        const int ti = bi[i];
                        bi[i] = bi[i - 1];
                        bi[i - 1] = ti;
                    }
                }
            }
        }
    }
    if (live) {
        int *io = idx_out + (static_cast<size_t>(b) * Q + q) * kKnn;
#pragma unroll
        for (int i = 0; i < kKnn; i += 4) *reinterpret_cast<int4 *>(io + i) = make_int4(bi[i], bi[i + 1], bi[i + 2], bi[i + 3]);
        if (dist_out) {
            float *dout = dist_out + (static_cast<size_t>(b) * Q + q) * kKnn;
#pragma unroll
            for (int i = 0; i < kKnn; ++i) dout[i] = sqrtf(bd[i]);
        }
    }
}

// Ragged form for a batch of cells: cell c owns the queries [q_off[c], q_off[c+1]) of one flat
// query array and the points [n_off[c], n_off[c+1]) of one concatenated cloud; the indices written are GLOBAL rows of the
// concatenated cloud, so the neighbour gather of embed_first works on the flat arrays.  Same scan as knn16_kernel.
__global__ void __launch_bounds__(64) knn16_cells_kernel(const float *__restrict__ x, const float *__restrict__ pc,
                                                          const int *__restrict__ q_off, const int *__restrict__ n_off,
                                                          int *__restrict__ idx_out)
{
    __shared__ float sp[kKnnTile * 3];
    const int c = blockIdx.y;
    const int q0 = q_off[c], Q = q_off[c + 1] - q0;
    if (static_cast<int>(blockIdx.x) * 64 >= Q) return;   // uniform per block
    const int n0 = n_off[c], N = n_off[c + 1] - n0;
    const int q = blockIdx.x * 64 + threadIdx.x;
    const bool live = q < Q;
    const float *xb = x + static_cast<size_t>(q0 + (live ? q : 0)) * 3;
    const float qx = xb[0], qy = xb[1], qz = xb[2];
    const float *pcb = pc + static_cast<size_t>(n0) * 3;

    float bd[kKnn];
    int bi[kKnn];
#pragma unroll
    for (int i = 0; i < kKnn; ++i) bd[i] = FLT_MAX, bi[i] = 0;
    for (int t0 = 0; t0 < N; t0 += kKnnTile) {
        const int n = min(kKnnTile, N - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < n * 3; i += blockDim.x) sp[i] = pcb[static_cast<size_t>(t0) * 3 + i];
        __syncthreads();
        for (int j = 0; j < n; ++j) {
            const float dx = qx - sp[3 * j], dy = qy - sp[3 * j + 1], dz = qz - sp[3 * j + 2];
            const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            if (d < bd[kKnn - 1]) {
                bd[kKnn - 1] = d;
                bi[kKnn - 1] = t0 + j;
#pragma unroll
                for (int i = kKnn - 1; i > 0; --i) {
                    if (bd[i] < bd[i - 1]) {
                        const float td = bd[i];
                        bd[i] = bd[i - 1];
                        bd[i - 1] = td;
                        const int ti = bi[i];
                        bi[i] = bi[i - 1];
                        bi[i - 1] = ti;
                    }
                }
            }
        }
    }
    if (live) {
        int *io = idx_out + static_cast<size_t>(q0 + q) * kKnn;
#pragma unroll
        for (int i = 0; i < kKnn; i += 4)
            *reinterpret_cast<int4 *>(io + i) = make_int4(n0 + bi[i], n0 + bi[i + 1], n0 + bi[i + 2], n0 + bi[i + 3]);
    }
}

// out[r, :] = table[row_of[r], :]   (N % 4 == 0): the per-cell bias of the first head layer for a flat batch of queries
__global__ void __launch_bounds__(256) gather_rows_kernel(const float *__restrict__ table, const int *__restrict__ row_of,
                                                          float *__restrict__ out, long long rows, int n4)
{
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < rows * n4; i += gridDim.x * 256ll) {
        const long long r = i / n4;
        const int c = static_cast<int>(i - r * n4);
        reinterpret_cast<float4 *>(out)[i] = __ldg(reinterpret_cast<const float4 *>(table) + static_cast<size_t>(row_of[r]) * n4 + c);
    }
}

// ------------------------------------------------------------------------------------------------
// First embedding layer: h = GELU(W1 p + b1), p in R^IN (IN = 3 or 4); one thread per (token, 4 features).
// GATHER: p = pc[b, idx[b, q, j]] - x[b, q]  (token = (b*Q + q)*16 + j); otherwise p = in[token].
// append: copy p behind the features (columns inner .. inner+IN-1).
// ------------------------------------------------------------------------------------------------
struct EmbedParams {
    const float *in;      // (T, ld_in) rows, first IN entries used            [!GATHER]
    int ld_in;
    const float *pc;      // (B, N, 3)                                          [GATHER]
    const float *x;       // (B, Q, 3)
    const int *idx;       // (B, Q, 16)
    int Q, N;
    const float *w, *b;   // (inner, IN), (inner)
    int inner, append;
    float *out;           // (T, ldo)
    int ldo;
    long long T;
};

template <int IN, bool GATHER>
__global__ void __launch_bounds__(256) embed_first_kernel(const EmbedParams p)
{
    extern __shared__ float sw[];  // w (inner*IN) | b (inner)
    for (int i = threadIdx.x; i < p.inner * IN; i += blockDim.x) sw[i] = p.w[i];
    for (int i = threadIdx.x; i < p.inner; i += blockDim.x) sw[p.inner * IN + i] = p.b[i];
    __syncthreads();
    const float *sb = sw + p.inner * IN;
    const int groups = (p.inner + (p.append ? IN : 0) + 3) / 4;  // float4 groups per token
    // the block size and the grid stride are multiples of 32, so a thread serves the same 4 output columns for every token:
    // its 4 x IN weights and 4 biases live in registers (read from shared memory inside the loop, the stride-4 / stride-12
    // accesses were 4-way bank conflicts: ~60 wavefronts per token made the kernel LSU-bound at 2 TB/s)
    const int c4 = static_cast<int>(threadIdx.x & 31);
    float wr[4][IN], br[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int c = c4 * 4 + e;
        br[e] = c < p.inner ? sb[c] : 0.f;
#pragma unroll
        for (int k = 0; k < IN; ++k) wr[e][k] = c < p.inner ? sw[c * IN + k] : 0.f;
    }
    // A warp takes 32 consecutive tokens: lane l fetches the input of token l once (index, point, query: with one token per
    // warp-trip every lane repeated that address arithmetic and those loads -- 228 warp instructions per token, issue-bound at
    // 76 % issue-active), then the 32 inputs are broadcast one after the other with shuffles and lane l produces its 4 columns.
    const int lane = threadIdx.x & 31;
    const bool has_cols = c4 < groups;
    const long long n_tiles = (p.T + 31) >> 5;
    const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long tile = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5; tile < n_tiles; tile += n_warps) {
        const long long t = tile * 32 + lane;
        float v[IN];
#pragma unroll
        for (int e = 0; e < IN; ++e) v[e] = 0.f;
        if (t < p.T) {
            if (GATHER) {
                const long long bq = t >> 4;            // b*Q + q
                const int b = static_cast<int>(static_cast<unsigned>(bq) / static_cast<unsigned>(p.Q));   // bq < 2^32 (launcher)
                const float *pp = p.pc + (static_cast<size_t>(b) * p.N + __ldg(p.idx + t)) * 3;
                const float *xx = p.x + bq * 3;
#pragma unroll
                for (int e = 0; e < IN; ++e) v[e] = __ldg(pp + e) - __ldg(xx + e);
            } else {
                const float *pp = p.in + t * p.ld_in;
#pragma unroll
                for (int e = 0; e < IN; ++e) v[e] = __ldg(pp + e);
            }
        }
        const int n = static_cast<int>(p.T - tile * 32 < 32 ? p.T - tile * 32 : 32);
        float *orow = p.out + tile * 32 * p.ldo + c4 * 4;
        for (int j = 0; j < n; ++j, orow += p.ldo) {
            float a[IN];
#pragma unroll
            for (int e = 0; e < IN; ++e) a[e] = __shfl_sync(0xffffffffu, v[e], j);
            if (!has_cols) continue;
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = c4 * 4 + e;
                float y = 0.f;
                if (c < p.inner) {
                    y = br[e];
#pragma unroll
                    for (int k = 0; k < IN; ++k) y = fmaf(wr[e][k], a[k], y);
                    y = gelu_exact(y);
                } else if (p.append && c < p.inner + IN) {
                    y = a[c - p.inner];
                }
                o[e] = y;
            }
            *reinterpret_cast<float4 *>(orow) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// attn16: one warp per 16-token sequence, 4 heads; lane = (query i, head pair).
// qkv rows: [q (H*DQK) | k (H*DQK) | v (H*DV)], out rows: (H*DV).
// ------------------------------------------------------------------------------------------------
template <int DQK, int DV>
__global__ void __launch_bounds__(256) attn16_kernel(const float *__restrict__ qkv, int ldq, float *__restrict__ out, int ldo,
                                                     long long n_seq)
{
    constexpr int H = 4;
    constexpr int W = 2 * H * DQK + H * DV;  // 192
    constexpr int LD = W + 4;                // padded row: conflict-free 128-bit row reads
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *t = smem + warp * 16 * LD;
    const int i = lane & 15, hh = lane >> 4;
    const float scale = rsqrtf(static_cast<float>(DQK)) * 1.4426950408889634f;  // 1/sqrt(d) * log2(e)

    for (long long s = blockIdx.x * 8ll + warp; s < n_seq; s += gridDim.x * 8ll) {
        const float *src = qkv + s * 16 * ldq;
        __syncwarp();
        // the whole 16 x 192 tile as 16-byte async copies: all 24 per lane in flight at once (through registers the compiler
        // kept 4 loads in flight, i.e. 32 KB per SM: the kernel was latency-bound at 63 % of the HBM roofline)
#pragma unroll
        for (int e = lane; e < 16 * (W / 4); e += 32) {
            const int r = e / (W / 4), c = e % (W / 4);
            cp_async16(t + r * LD + c * 4, src + r * ldq + c * 4);
        }
        cp_async_commit();
        cp_async_wait_all();
        __syncwarp();
        float o[2][DV];
#pragma unroll
        for (int hq = 0; hq < 2; ++hq) {
            const int h = 2 * hh + hq;
            float q[DQK];
#pragma unroll
            for (int c = 0; c < DQK; c += 4) {
                const float4 v = *reinterpret_cast<const float4 *>(t + i * LD + h * DQK + c);
                q[c] = v.x, q[c + 1] = v.y, q[c + 2] = v.z, q[c + 3] = v.w;
            }
            float sc[16];
            float m = -FLT_MAX;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float a = 0.f;
#pragma unroll
                for (int c = 0; c < DQK; c += 4) {
                    const float4 k = *reinterpret_cast<const float4 *>(t + j * LD + H * DQK + h * DQK + c);
                    a = fmaf(q[c], k.x, a), a = fmaf(q[c + 1], k.y, a), a = fmaf(q[c + 2], k.z, a), a = fmaf(q[c + 3], k.w, a);
                }
                sc[j] = a * scale;
                m = fmaxf(m, sc[j]);
            }
            float l = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                sc[j] = ex2_approx(sc[j] - m);   // arguments <= 0: the flush-to-zero of tiny weights is harmless
                l += sc[j];
            }
            const float inv = 1.0f / l;
#pragma unroll
            for (int d = 0; d < DV; ++d) o[hq][d] = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float pj = sc[j] * inv;
#pragma unroll
                for (int d = 0; d < DV; d += 4) {
                    const float4 v = *reinterpret_cast<const float4 *>(t + j * LD + 2 * H * DQK + h * DV + d);
                    o[hq][d] = fmaf(pj, v.x, o[hq][d]), o[hq][d + 1] = fmaf(pj, v.y, o[hq][d + 1]);
                    o[hq][d + 2] = fmaf(pj, v.z, o[hq][d + 2]), o[hq][d + 3] = fmaf(pj, v.w, o[hq][d + 3]);
                }
            }
        }
        __syncwarp();  // everyone is done reading q/k/v: reuse the tile (columns 0 .. H*DV-1) as the output stage
#pragma unroll
        for (int hq = 0; hq < 2; ++hq)
#pragma unroll
            for (int d = 0; d < DV; d += 4)
                *reinterpret_cast<float4 *>(t + i * LD + (2 * hh + hq) * DV + d) =
                    make_float4(o[hq][d], o[hq][d + 1], o[hq][d + 2], o[hq][d + 3]);
        __syncwarp();
        float *dst = out + s * 16 * ldo;
        for (int e = lane; e < 16 * (H * DV / 4); e += 32) {
            const int r = e / (H * DV / 4), c = e % (H * DV / 4);
            *reinterpret_cast<float4 *>(dst + r * ldo + c * 4) = *reinterpret_cast<const float4 *>(t + r * LD + c * 4);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// attn_dense: CTA = 64 queries of one (cloud, head), 8 warps = (value-dim half) x (4 key splits); lane = a PAIR of
// queries, so every K / V row read from shared memory (a warp-wide broadcast) feeds 2 x (DQK + DV/2) FMAs.  Each
// key split runs its own online softmax (flash style); the 4 partial (max, sum, accumulator) triples of a query are
// merged through shared memory at the end.
// ------------------------------------------------------------------------------------------------
template <int DQK, int DV>
__global__ void __launch_bounds__(256) attn_dense_kernel(const float *__restrict__ qkv, int ldq, float *__restrict__ out,
                                                         int ldo, int S_stride, const int *__restrict__ lens)
{
    constexpr int H = 4, KT = 32, NS = 4, DH = DV / 2;
    extern __shared__ __align__(16) float sm[];
    float *sk = sm;                       // [NS][KT][DQK]
    float *sv = sm + NS * KT * DQK;       // [NS][KT][DV]
    const int b = blockIdx.z, h = blockIdx.y;
    // ragged batches: cloud b occupies rows [b * S_stride, b * S_stride + S); rows beyond S are padding
    const int S = lens ? lens[b] : S_stride;
    if (blockIdx.x * 64 >= S) {   // padding-only tile: keep the rows finite for the layers that follow
        for (int e = threadIdx.x; e < 64 * (DV / 4); e += 256) {
            const int qi = blockIdx.x * 64 + e / (DV / 4);
            if (qi < S_stride)
                *reinterpret_cast<float4 *>(out + (static_cast<size_t>(b) * S_stride + qi) * ldo + h * DV + (e % (DV / 4)) * 4) =
                    make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = warp & 1, ks = warp >> 1;
    const int qa = blockIdx.x * 64 + 2 * lane;
    const float *base = qkv + static_cast<size_t>(b) * S_stride * ldq;
    const float scale = rsqrtf(static_cast<float>(DQK)) * 1.4426950408889634f;   // 1/sqrt(d) * log2(e)

    float q[2][DQK];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const float *qp = base + static_cast<size_t>(min(qa + u, S - 1)) * ldq + h * DQK;
#pragma unroll
        for (int c = 0; c < DQK; c += 4) {
            const float4 v = *reinterpret_cast<const float4 *>(qp + c);
            q[u][c] = v.x * scale, q[u][c + 1] = v.y * scale, q[u][c + 2] = v.z * scale, q[u][c + 3] = v.w * scale;
        }
    }
    float acc[2][DH];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int d = 0; d < DH; ++d) acc[u][d] = 0.f;
    float m[2] = {-FLT_MAX, -FLT_MAX}, l[2] = {0.f, 0.f};

    const int per = ((S + NS - 1) / NS + KT - 1) / KT * KT;   // keys per split, a multiple of the staging tile
    const int k_begin = ks * per, k_end = min(k_begin + per, S);
    for (int it = 0; it < per / KT; ++it) {
        __syncthreads();
        // stage KT keys of every split: K rows then V rows, zero beyond the end of the cloud
        for (int e = threadIdx.x; e < NS * KT * (DQK / 4); e += 256) {
            const int r = e / (DQK / 4), c = e % (DQK / 4);
            const int key = (r / KT) * per + it * KT + (r % KT);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (key < S) v = *reinterpret_cast<const float4 *>(base + static_cast<size_t>(key) * ldq + H * DQK + h * DQK + c * 4);
            *reinterpret_cast<float4 *>(sk + r * DQK + c * 4) = v;
        }
        for (int e = threadIdx.x; e < NS * KT * (DV / 4); e += 256) {
            const int r = e / (DV / 4), c = e % (DV / 4);
            const int key = (r / KT) * per + it * KT + (r % KT);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (key < S) v = *reinterpret_cast<const float4 *>(base + static_cast<size_t>(key) * ldq + 2 * H * DQK + h * DV + c * 4);
            *reinterpret_cast<float4 *>(sv + r * DV + c * 4) = v;
        }
        __syncthreads();
        const int nvalid = min(KT, k_end - (k_begin + it * KT));   // valid keys of this warp's tile (may be <= 0)
        const float *wk = sk + ks * KT * DQK, *wv = sv + ks * KT * DV + half * DH;
#pragma unroll 1
        for (int j0 = 0; j0 < KT; j0 += 8) {
            if (j0 >= nvalid) break;
            float sc[2][8];
            float tm0 = m[0], tm1 = m[1];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int c = 0; c < DQK; c += 4) {
                    const float4 k = *reinterpret_cast<const float4 *>(wk + (j0 + j) * DQK + c);
                    a0 = fmaf(q[0][c], k.x, a0), a0 = fmaf(q[0][c + 1], k.y, a0), a0 = fmaf(q[0][c + 2], k.z, a0), a0 = fmaf(q[0][c + 3], k.w, a0);
                    a1 = fmaf(q[1][c], k.x, a1), a1 = fmaf(q[1][c + 1], k.y, a1), a1 = fmaf(q[1][c + 2], k.z, a1), a1 = fmaf(q[1][c + 3], k.w, a1);
                }
                const bool ok = j0 + j < nvalid;
                sc[0][j] = ok ? a0 : -FLT_MAX;
                sc[1][j] = ok ? a1 : -FLT_MAX;
                tm0 = fmaxf(tm0, sc[0][j]);
                tm1 = fmaxf(tm1, sc[1][j]);
            }
            const float c0 = exp2f(m[0] - tm0), c1 = exp2f(m[1] - tm1);
            m[0] = tm0, m[1] = tm1;
            l[0] *= c0, l[1] *= c1;
#pragma unroll
            for (int d = 0; d < DH; ++d) acc[0][d] *= c0, acc[1][d] *= c1;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool ok = j0 + j < nvalid;
                const float p0 = ok ? exp2f(sc[0][j] - m[0]) : 0.f, p1 = ok ? exp2f(sc[1][j] - m[1]) : 0.f;
                l[0] += p0, l[1] += p1;
#pragma unroll
                for (int d = 0; d < DH; d += 4) {
                    const float4 v = *reinterpret_cast<const float4 *>(wv + (j0 + j) * DV + d);
                    acc[0][d] = fmaf(p0, v.x, acc[0][d]), acc[0][d + 1] = fmaf(p0, v.y, acc[0][d + 1]);
                    acc[0][d + 2] = fmaf(p0, v.z, acc[0][d + 2]), acc[0][d + 3] = fmaf(p0, v.w, acc[0][d + 3]);
                    acc[1][d] = fmaf(p1, v.x, acc[1][d]), acc[1][d + 1] = fmaf(p1, v.y, acc[1][d + 1]);
                    acc[1][d + 2] = fmaf(p1, v.z, acc[1][d + 2]), acc[1][d + 3] = fmaf(p1, v.w, acc[1][d + 3]);
                }
            }
        }
    }

    // merge the 4 key splits: (2,3) -> (0,1), then 1 -> 0; record = [m, l, acc[DH]] per (query of the pair), lane-major
    constexpr int REC = DH + 2;
    float *mg = sm;   // [2 splits][2 halves][2 queries][REC][32 lanes]  (fits in the staging area)
    auto slot = [&](int sp, int u, int f) { return mg + (((sp * 2 + half) * 2 + u) * REC + f) * 32 + lane; };
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
        const int senders_from = round == 0 ? 2 : 1, n_send = round == 0 ? 2 : 1;
        __syncthreads();
        if (ks >= senders_from && ks < senders_from + n_send) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                *slot(ks - senders_from, u, 0) = m[u];
                *slot(ks - senders_from, u, 1) = l[u];
#pragma unroll
                for (int d = 0; d < DH; ++d) *slot(ks - senders_from, u, 2 + d) = acc[u][d];
            }
        }
        __syncthreads();
        if (ks < n_send) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float mo = *slot(ks, u, 0), lo = *slot(ks, u, 1);
                const float mn = fmaxf(m[u], mo);
                const float ca = exp2f(m[u] - mn), cb = exp2f(mo - mn);
                l[u] = l[u] * ca + lo * cb;
#pragma unroll
                for (int d = 0; d < DH; ++d) acc[u][d] = acc[u][d] * ca + *slot(ks, u, 2 + d) * cb;
                m[u] = mn;
            }
        }
    }
    if (ks == 0) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (qa + u < S_stride) {
                const float inv = qa + u < S ? 1.0f / l[u] : 0.f;   // padding rows of a ragged batch: zeros
                float *dst = out + (static_cast<size_t>(b) * S_stride + qa + u) * ldo + h * DV + half * DH;
#pragma unroll
                for (int d = 0; d < DH; d += 4)
                    *reinterpret_cast<float4 *>(dst + d) =
                        make_float4(acc[u][d] * inv, acc[u][d + 1] * inv, acc[u][d + 2] * inv, acc[u][d + 3] * inv);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// colpool: out_max[b, c] = max_s in[b, s, c]; out_mean[b, c] = mean_s in[b, s, c]   (either may be null)
// block = 32 columns x 32 row lanes
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) colpool_kernel(const float *__restrict__ in, int ld, int S_stride, int N,
                                                       float *__restrict__ out_max, float *__restrict__ out_mean, int ldo,
                                                       const int *__restrict__ lens)
{
    constexpr int RL = 32;   // row lanes per column (the first version had 8 and one dependent load chain per lane: 73 us
                             // for one 2048-token cloud)
    __shared__ float smax[RL][33], ssum[RL][33];
    const int b = blockIdx.y;
    const int S = lens ? lens[b] : S_stride;   // ragged batches: only the first lens[b] rows of cloud b are tokens
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ry = threadIdx.x >> 5;
    float mx[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX}, sm[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < N) {
        const float *p = in + static_cast<size_t>(b) * S_stride * ld + c;
        int s = ry;
        for (; s + 3 * RL < S; s += 4 * RL) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float v = p[static_cast<size_t>(s + u * RL) * ld];
                mx[u] = fmaxf(mx[u], v);
                sm[u] += v;
            }
        }
        for (; s < S; s += RL) {
            const float v = p[static_cast<size_t>(s) * ld];
            mx[0] = fmaxf(mx[0], v);
            sm[0] += v;
        }
    }
    smax[ry][threadIdx.x & 31] = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
    ssum[ry][threadIdx.x & 31] = (sm[0] + sm[1]) + (sm[2] + sm[3]);
    __syncthreads();
    if (ry == 0 && c < N) {
        float m = smax[0][threadIdx.x], t = ssum[0][threadIdx.x];
#pragma unroll
        for (int r = 1; r < RL; ++r) {
            m = fmaxf(m, smax[r][threadIdx.x]);
            t += ssum[r][threadIdx.x];
        }
        if (out_max) out_max[static_cast<size_t>(b) * ldo + c] = S > 0 ? m : 0.f;
        if (out_mean) out_mean[static_cast<size_t>(b) * ldo + c] = S > 0 ? t / static_cast<float>(S) : 0.f;
    }
}

// ------------------------------------------------------------------------------------------------
// vis_embed_finish: x0[t] = [e (F) | gmax[b] (F) | pts[t] (IN)]  (width D = 2F + IN <= 256), ln = LayerNorm(x0)
// e is already in x0[:, :F]; one warp per token, 8 columns per lane.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vis_embed_finish_kernel(float *__restrict__ x0, int ld, const float *__restrict__ gmax,
                                                               int ldg, const float *__restrict__ pts, int ldp, int F, int IN,
                                                               int S, long long T, const float *__restrict__ g,
                                                               const float *__restrict__ bta, float eps, float *__restrict__ ln,
                                                               int ldl)
{
    const int D = 2 * F + IN;
    const int lane = threadIdx.x & 31;
    const long long t = blockIdx.x * 8ll + (threadIdx.x >> 5);
    if (t >= T) return;
    const int b = static_cast<int>(t / S);
    float v[8];
    float sum = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = lane + 32 * e;
        float y = 0.f;
        if (c < F) y = x0[t * ld + c];
        else if (c < 2 * F) y = gmax[static_cast<size_t>(b) * ldg + c - F];
        else if (c < D) y = pts[t * ldp + c - 2 * F];
        v[e] = y;
        if (c >= F && c < D) x0[t * ld + c] = y;
        sum += y;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    const float mean = sum / static_cast<float>(D);
    float var = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = lane + 32 * e;
        if (c < D) {
            const float d = v[e] - mean;
            var = fmaf(d, d, var);
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) var += __shfl_xor_sync(0xffffffffu, var, d);
    const float rstd = 1.0f / sqrtf(var / static_cast<float>(D) + eps);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = lane + 32 * e;
        if (c < D) ln[t * ldl + c] = (v[e] - mean) * rstd * g[c] + bta[c];
    }
}

// bias_out[b, n] = bias[n] + sum_k W[n, k] * g[b, k];   one warp per (b, n)
__global__ void __launch_bounds__(256) bias_gemv_kernel(const float *__restrict__ W, int ldw, const float *__restrict__ bias,
                                                        const float *__restrict__ g, int ldg, int N, int K,
                                                        float *__restrict__ out, int B)
{
    const int lane = threadIdx.x & 31;
    const long long w = blockIdx.x * 8ll + (threadIdx.x >> 5);
    if (w >= static_cast<long long>(B) * N) return;
    const int b = static_cast<int>(w / N), n = static_cast<int>(w % N);
    float a = 0.f;
    for (int k = lane; k < K; k += 32) a = fmaf(W[static_cast<size_t>(n) * ldw + k], g[static_cast<size_t>(b) * ldg + k], a);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
    if (lane == 0) out[static_cast<size_t>(b) * N + n] = a + bias[n];
}

}  // namespace

// ---- host launchers ----------------------------------------------------------------------------
int knn16(const float *x, const float *pc, int *idx, float *dist, int B, int Q, int N, cudaStream_t stream)
{
    MAC_REQUIRE(x && pc && idx, "null tensor pointer");
    MAC_REQUIRE(B > 0 && Q > 0 && N >= kKnn, "kNN needs B > 0, Q > 0 and at least 16 cloud points (got B=%d Q=%d N=%d)", B, Q, N);
    MAC_REQUIRE((reinterpret_cast<uintptr_t>(idx) & 15u) == 0, "idx must be 16-byte aligned");
    dim3 grid((Q + 63) / 64, B);   // 64 queries per CTA: 16384-query passes still fill the 148 SMs
    knn16_kernel<<<grid, 64, 0, stream>>>(x, pc, idx, dist, Q, N);
    MAC_CUDA(cudaGetLastError());
    count_launch();
    return MAC_OK;
}

int knn16_cells(const float *x, const float *pc, const int *q_off, const int *n_off, int *idx, int n_cells, int max_q,
                cudaStream_t stream)
{
    MAC_REQUIRE(x && pc && q_off && n_off && idx, "null tensor pointer");
    MAC_REQUIRE(n_cells > 0 && n_cells <= 65535 && max_q > 0, "kNN over cells needs 1..65535 cells (got %d) and queries", n_cells);
    MAC_REQUIRE((reinterpret_cast<uintptr_t>(idx) & 15u) == 0, "idx must be 16-byte aligned");
    dim3 grid((max_q + 63) / 64, n_cells);
    knn16_cells_kernel<<<grid, 64, 0, stream>>>(x, pc, q_off, n_off, idx);
    MAC_CUDA(cudaGetLastError());
    count_launch();
    return MAC_OK;
}

int gather_rows(const float *table, const int *row_of, float *out, long long rows, int N, cudaStream_t stream)
{
    MAC_REQUIRE(table && row_of && out && rows > 0 && N > 0 && N % 4 == 0, "gather_rows needs N %% 4 == 0");
    const long long want = (rows * (N / 4) + 255) / 256;
    gather_rows_kernel<<<static_cast<unsigned>(want < 148 * 16 ? want : 148 * 16), 256, 0, stream>>>(table, row_of, out, rows, N / 4);
    MAC_CUDA(cudaGetLastError());
    count_launch();
    return MAC_OK;
}

int embed_first(const float *in, int ld_in, int in_dim, const float *pc, const float *x, const int *idx, int Q, int N,
                const float *w, const float *b, int inner, int append, float *out, int ldo, long long T,
                cudaStream_t stream)
{
    MAC_REQUIRE(w && b && out && T > 0, "null tensor pointer");
    MAC_REQUIRE(in_dim == 3 || in_dim == 4, "embedding input dimension must be 3 or 4");
    MAC_REQUIRE(inner + (append ? in_dim : 0) <= 128 && ldo % 4 == 0 && ldo >= (inner + (append ? in_dim : 0) + 3) / 4 * 4,
                "embedding width %d does not fit the 128-column staging row (ldo=%d)", inner, ldo);
    const bool gather = idx != nullptr;
    MAC_REQUIRE(gather ? (pc && x && in_dim == 3) : (in != nullptr), "embedding input missing");
    MAC_REQUIRE((T >> 4) < (1ll << 32), "too many tokens for one launch");
    EmbedParams p{};
    p.in = in, p.ld_in = ld_in, p.pc = pc, p.x = x, p.idx = idx, p.Q = Q, p.N = N;
    p.w = w, p.b = b, p.inner = inner, p.append = append, p.out = out, p.ldo = ldo, p.T = T;
    const size_t smem = static_cast<size_t>(inner) * (in_dim + 1) * sizeof(float);
    // one 32-token tile per warp; 8 warps per CTA, or one when there are too few tiles to fill the machine that way
    const long long n_tiles = (T + 31) / 32;
    const int block = n_tiles < 148 * 8 ? 32 : 256;
    const long long want = (n_tiles * 32 + block - 1) / block;
    const int grid = static_cast<int>(want < 148 * 16 ? want : 148 * 16);
    if (gather) embed_first_kernel<3, true><<<grid, block, smem, stream>>>(p);
    else if (in_dim == 3) embed_first_kernel<3, false><<<grid, block, smem, stream>>>(p);
    else embed_first_kernel<4, false><<<grid, block, smem, stream>>>(p);
    MAC_CUDA(cudaGetLastError());
    count_launch();
    return MAC_OK;
}

int attn16(const float *qkv, int ldq, float *out, int ldo, long long n_seq, int dqk, int dv, cudaStream_t stream)
{
    MAC_REQUIRE(qkv && out && n_seq > 0, "null tensor pointer");
    MAC_REQUIRE(dqk == 8 && dv == 32, "attn16 is built for 4 heads of (8, 32) dims");
    MAC_REQUIRE(ldq % 4 == 0 && ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0,
                "attention rows must be 16-byte aligned");
    constexpr int LD = 2 * 4 * 8 + 4 * 32 + 4;
    const size_t smem = 8 * 16 * LD * sizeof(float);
    static DeviceOnce once;
    if (int rc = ensure_dynamic_smem(once, attn16_kernel<8, 32>, static_cast<int>(smem))) return rc;
    const long long want = (n_seq + 7) / 8;
    const int grid = static_cast<int>(want < 148 * 2 ? want : 148 * 2);
    attn16_kernel<8, 32><<<grid, 256, smem, stream>>>(qkv, ldq, out, ldo, n_seq);
    MAC_CUDA(cudaGetLastError());
    count_launch();
    return MAC_OK;
}

int attn_dense(const float *qkv, int ldq, float *out, int ldo, int B, int S, int dqk, int dv, cudaStream_t stream,
               const int *lens)
{
    MAC_REQUIRE(qkv && out && B > 0 && S > 0, "null tensor pointer");
    MAC_REQUIRE(ldq % 4 == 0 && ldo % 4 == 0, "attention rows must be 16-byte aligned");
    dim3 grid((S + 63) / 64, 4, B);
    static DeviceOnce once;
    if (int rc = ensure_dynamic_smem(once, attn_dense_kernel<16, 64>, 4 * 32 * (16 + 64) * 4)) return rc;
    if (dqk == 8 && dv == 32) attn_dense_kernel<8, 32><<<grid, 256, 4 * 32 * (8 + 32) * 4, stream>>>(qkv, ldq, out, ldo, S, lens);
    else if (dqk == 16 && dv == 64) attn_dense_kernel<16, 64><<<grid, 256, 4 * 32 * (16 + 64) * 4, stream>>>(qkv, ldq, out, ldo, S, lens);
    else {
        set_error("attn_dense is built for 4 heads of (8, 32) or (16, 64) dims, got (%d, %d)", dqk, dv);
        return MAC_ERR_UNSUPPORTED;
    }
    MAC_CUDA(cudaGetLastError());
    count_launch();
    return MAC_OK;
}

int colpool(const float *in, int ld, int B, int S, int N, float *out_max, float *out_mean, int ldo, cudaStream_t stream,
            const int *lens)
{
    MAC_REQUIRE(in && (out_max || out_mean) && B > 0 && S > 0 && N > 0, "null tensor pointer");
    dim3 grid((N + 31) / 32, B);
    colpool_kernel<<<grid, 1024, 0, stream>>>(in, ld, S, N, out_max, out_mean, ldo, lens);
    MAC_CUDA(cudaGetLastError());
    count_launch();
    return MAC_OK;
}

int vis_embed_finish(float *x0, int ld, const float *gmax, int ldg, const float *pts, int ldp, int F, int in_dim, int S,
                     long long T, const float *g, const float *b, float eps, float *ln, int ldl, cudaStream_t stream)
{
    MAC_REQUIRE(x0 && gmax && pts && g && b && ln && T > 0, "null tensor pointer");
    MAC_REQUIRE(2 * F + in_dim <= 256, "embedding width %d exceeds 256", 2 * F + in_dim);
    vis_embed_finish_kernel<<<static_cast<unsigned>((T + 7) / 8), 256, 0, stream>>>(x0, ld, gmax, ldg, pts, ldp, F, in_dim, S, T, g,
                                                                                   b, eps, ln, ldl);
    MAC_CUDA(cudaGetLastError());
    count_launch();
    return MAC_OK;
}

int bias_gemv(const float *W, int ldw, const float *bias, const float *g, int ldg, int N, int K, float *out, int B,
              cudaStream_t stream)
{
    MAC_REQUIRE(W && bias && g && out && B > 0, "null tensor pointer");
    const long long warps = static_cast<long long>(B) * N;
    bias_gemv_kernel<<<static_cast<unsigned>((warps + 7) / 8), 256, 0, stream>>>(W, ldw, bias, g, ldg, N, K, out, B);
    MAC_CUDA(cudaGetLastError());
    count_launch();
    return MAC_OK;
}

}  // namespace mac

extern "C" int mac_knn16_f32(const float *x, const float *pc, int *idx, float *dist, int B, int Q, int N, void *stream)
{
    return mac::knn16(x, pc, idx, dist, B, Q, N, static_cast<cudaStream_t>(stream));
}
