// Coverage-gain integration: for every (cloud b, camera c) the mean over surface points p of
//     act( sum_k Y_k(direction of cam_c - pt_p) * H[b,p,k] ),   k = l*l+l+m, l < 8  (64 terms)
// Replaces SconeVis.compute_coverage_gain / compute_visibilities and
// Macarons.compute_visibility_gains (reference networks/SconeVis.py:164-252, Macarons.py:138-178).
//
// Design (sm_100a, CUDA cores; see DESIGN.md section "coverage-gain kernel"):
//   * one surface point per lane.  The 32 x 64 coefficient tile of a warp is staged global -> shared by
//     TMA (two 32-row x 128-B boxes of a 3-D tensor map over (64, P, B), SWIZZLE_128B, mbarrier
//     completion; rows past P are zero-filled by the hardware), pulled into registers with 16
//     conflict-free LDS.128 (the swizzle un-does the 128-B row stride) and converted ONCE to 64 "Horner-ready"
//     coefficients (sh_horner_gen.h): on the unit sphere the SH sum is P(ct, uz) + ux Q(ct, uz) with u the unit
//     ray (deg P <= 7, deg Q <= 6: 64 monomial coefficients), i.e. 63 FMAs per (point, camera) pair and no
//     trigonometry.
//   * the warp then sweeps a chunk of cameras (broadcast LDS.128 from a per-warp table).  Per-pair
//     values go through a padded per-warp transpose buffer so that the sum over the 32 points costs
//     ~2 instructions per pair instead of a 10-instruction shuffle tree.
//   * the 32-point sum of every (tile, camera) is converted to exact fixed point and all further additions
//     (per-warp accumulators in shared memory, one 64-bit atomic per camera when a warp changes camera chunk)
//     are integer additions: associative, so the result is bitwise independent of scheduling, launch geometry
//     and of how cameras are partitioned across GPUs.  Warps are persistent and pull (camera chunk, tile)
//     tasks from a global counter.  The last CTA to finish converts the accumulators to the fp32 mean and
//     re-zeroes the workspace.
#include <cuda.h>
#include <stdlib.h>

#include "gather_dev.h"
#include "mac_common.h"
#include "sh_eval_mixed.h"
#include "sh_horner_gen.h"

namespace mac {

namespace {

constexpr int kBoxBytes = 32 * 128;  // one TMA box: 32 points x 32 coefficients (128 B, the SWIZZLE_128B span)
constexpr int kRedStride = 36;     // transpose buffer row stride (floats)
constexpr int kCamChunkMax = 128;  // cameras per warp task

constexpr float kNegLog2e = -1.4426950408889634f;
constexpr float kFixSigmoid = 4294967296.0f;  // 2^32: sums of sigmoid values, |s| <= 32 * tiles_per_task
constexpr float kFixRelu = 16777216.0f;       // 2^24: sums of relu values (unbounded inputs)

struct __align__(1024) WarpSmem {
    float stage[2 * kBoxBytes / 4];  // TMA landing zone (coefficients 0-31 | 32-63); re-used as the 32x36 transpose buffer
    float4 cams[kCamChunkMax + 2];  // +2: the pipelined sweep normalises one pair ahead
    long long wacc[kCamChunkMax];   // per-camera sums of this warp, exact fixed point (2^-32 or 2^-24 units)
    uint64_t bar;
};
static_assert(sizeof(WarpSmem) % 1024 == 0, "SWIZZLE_128B boxes need 1024-B aligned shared memory");
static_assert(32 * kRedStride * 4 <= 2 * kBoxBytes, "transpose buffer must fit in the staging area");

struct CovgainParams {
    const float *pts;
    const float *harm;
    const float *cams;
    float *out;
    unsigned long long *acc;   // (B*C) fixed-point accumulators   [REDUCE only]
    unsigned int *flags;       // (B*C) non-finite markers         [REDUCE only]
    unsigned int *done;        // CTA completion ticket            [REDUCE only]
    unsigned int *next_task;   // dynamic task counter (tasks beyond the first one of every warp)
    int pts_dim;
    int B, P, C;
    int cam_begin, cam_end;
    int mean_count;            // divisor of the mean (P, or the total point count when the points arrive in slices)
    int finalize;              // 0: leave the partial sums in the workspace (more point slices follow)
    int cams_per_task;         // multiple of 32, <= kCamChunkMax
    int n_cam_chunks;
    int tiles_per_cloud;       // ceil(P / 32)
    int tiles_per_task;
    int runs_per_cloud;        // ceil(tiles_per_cloud / tiles_per_task)
    int total_tasks;           // B * runs_per_cloud * n_cam_chunks
    // fused all-gather (REDUCE only, push_world > 0): the finishing CTA stores this rank's score columns into the
    // (B, C) score board of every peer over NVLink and then raises its arrival flag there.
    int push_world, push_rank;
    unsigned int push_epoch;
    float *push_dst[MAC_MAX_PEERS];
    unsigned int *push_flag[MAC_MAX_PEERS];
    // fused wait + argmax (push_world > 0 and best != null): after raising its flags the finishing CTA waits for the
    // flags of all ranks on the LOCAL board and writes the replicated NBV index: the whole sharded step is one launch
    long long *best;
    int *status;
    // point-sharded form (push_partial != 0): the finishing CTA pushes the exact partial sums of ALL cameras over this rank's
    // points into slot push_rank of every peer's partial region, waits, adds the push_world partials and takes the argmax
    int push_partial;
    unsigned long long *part_dst[MAC_MAX_PEERS];
};

struct Ray {
    float ux, ct, uz;  // unit vector from the surface point to the camera: (x, y, z) / r
};

__device__ __forceinline__ Ray make_ray(const float4 cam, const float px, const float py, const float pz)
{
    const float dx = cam.x - px, dy = cam.y - py, dz = cam.z - pz;
    const float rinv = rsqrt_approx(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
    return Ray{dx * rinv, dy * rinv, dz * rinv};
}

template <bool PACKED, int NG, int NG0, int NGP>
__device__ __forceinline__ void load_coefficients(const float (&h)[64], float (&g)[NG], float (&g0)[NG0],
                                                  unsigned long long (&gp)[NGP], const float scale)
{
    if constexpr (PACKED) mac_sh_pretransform_pq(h, g0, gp, scale);
    else mac_sh_pretransform(h, g, scale);
}
// EVAL: 0 = P + u_x Q form with scalar FFMA (63 per ray; default), 1 = scalar A/B form (75 FFMA per ray),
// 2 = P + u_x Q form with packed fma.rn.f32x2 (27 FFMA2 + 9 FFMA per ray).  0 and 2 are bitwise identical.
template <int EVAL, int NG, int NG0, int NGP>
__device__ __forceinline__ void eval_two_rays(const float (&g)[NG], const float (&g0)[NG0],
                                              const unsigned long long (&gp)[NGP], const Ray a, const Ray b, float &za,
                                              float &zb)
{
    if constexpr (EVAL == 0) mac_sh_eval2_pq_mixed<0, false>(g0, gp, a.ux, a.ct, a.uz, b.ux, b.ct, b.uz, za, zb);
    else if constexpr (EVAL == 1) mac_sh_eval2(g, a.ux, a.ct, a.uz, b.ux, b.ct, b.uz, za, zb);
    else mac_sh_eval2_pq(g0, gp, a.ux, a.ct, a.uz, b.ux, b.ct, b.uz, za, zb);
}

// The activation is split so that its two MUFU results are consumed one loop iteration apart.
// Sigmoid: g was pre-scaled by -log2(e), so z = -x*log2(e) and sigmoid(x) = 1 / (1 + 2^z).
template <bool SIGMOID>
__device__ __forceinline__ float activation_start(const float z)
{
    return SIGMOID ? ex2_approx(z) : fmaxf(z, 0.0f);
}
template <bool SIGMOID>
__device__ __forceinline__ float activation_finish(const float e)
{
    return SIGMOID ? rcp_approx(1.0f + e) : e;
}

// Sum over the 32 points of a tile for the 32 cameras of one batch: lane j adds up row j of the transpose
// buffer (8 conflict-free LDS.128), converts the 32-point sum to fixed point and adds it to the per-warp,
// per-camera integer accumulator.  The float sum covers exactly one tile in a fixed order and every later addition
// is an integer addition, so the result does not depend on which warp processes which tile, on the launch geometry
// or on the camera partition.  Returns a non-zero mask if the sum is not finite (the camera's score becomes NaN).
template <bool SIGMOID>
__device__ __forceinline__ unsigned reduce_batch(const float *red, long long *wacc, const int cb, const int lane)
{
    __syncwarp();
    const float4 *r4 = reinterpret_cast<const float4 *>(&red[lane * kRedStride]);
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float4 v = r4[q];
        s += (v.x + v.y) + (v.z + v.w);
    }
    const bool ok = fabsf(s) < 1.0e12f;
    if (ok) wacc[cb + lane] += __float2ll_rn(s * (SIGMOID ? kFixSigmoid : kFixRelu));  // lane j owns camera cb + j
    __syncwarp();
    return ok ? 0u : (1u << (cb >> 5));
}

// WARPS x MINB = resident warps per SM (register budget 65536 / (32 * WARPS * MINB)); EVAL selects the evaluator.
template <bool SIGMOID, bool REDUCE, int EVAL, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) covgain_kernel(const CovgainParams prm, const __grid_constant__ CUtensorMap harm_map)
{
    constexpr bool PACKED = EVAL != 1;   // coefficients held as 8 floats + 28 (P, Q) register pairs
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    WarpSmem &ws = reinterpret_cast<WarpSmem *>(smem_raw)[warp];
    float *red = ws.stage;

    if (lane == 0) {
        mbar_init(&ws.bar, 1);
        fence_mbar_init();
    }
    __syncwarp();

    // Persistent warps: the first task of a warp is static, further ones come from a global counter.  Tasks are ordered
    // (camera chunk, cloud, run of tiles), so consecutive tasks of a warp almost always share the camera chunk: the camera
    // table is staged and the per-camera accumulators are flushed (one 64-bit atomic per camera) only when it changes.
    int task = blockIdx.x * WARPS + warp;
    int key = -1, b = 0, cam0 = 0, ncam = 0, ncam32 = 0;
    unsigned bad = 0;
    uint32_t parity = 0;
    auto flush = [&]() {
        if (!REDUCE || key < 0) return;
        __syncwarp();
        for (int j = lane; j < ncam; j += 32) {
            const size_t idx = static_cast<size_t>(b) * prm.C + cam0 + j;
            const long long q = ws.wacc[j];
            if (q != 0) atomicAdd(prm.acc + idx, static_cast<unsigned long long>(q));
            if (bad & (1u << (j >> 5))) atomicOr(prm.flags + idx, 1u);  // NaN / inf / overflow -> result is NaN
        }
        bad = 0;
    };
    const int tasks_per_chunk = prm.B * prm.runs_per_cloud;
    while (task < prm.total_tasks) {
        const int cc = task / tasks_per_chunk;
        const int rest = task - cc * tasks_per_chunk;
        const int run = rest % prm.runs_per_cloud;
        const int tb = rest / prm.runs_per_cloud;
        if (tb * prm.n_cam_chunks + cc != key) {
            flush();
            key = tb * prm.n_cam_chunks + cc;
            b = tb;
            cam0 = prm.cam_begin + cc * prm.cams_per_task;
            ncam = min(prm.cams_per_task, prm.cam_end - cam0);
            ncam32 = (ncam + 31) & ~31;
            __syncwarp();
            for (int j = lane; j < ncam32 + 2; j += 32) {
                float4 c = make_float4(0.f, 0.f, 1048576.f, 0.f);  // padding camera: finite, result discarded
                if (j < ncam) {
                    const float *src = prm.cams + (static_cast<size_t>(b) * prm.C + cam0 + j) * 3;
                    c = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), 0.f);
                }
                ws.cams[j] = c;
                if (j < ncam32) ws.wacc[j] = 0;
            }
            __syncwarp();
        }

        const int tile_end = min(prm.tiles_per_cloud, (run + 1) * prm.tiles_per_task);
        for (int tile = run * prm.tiles_per_task; tile < tile_end; ++tile) {
            const int p = tile * 32 + lane;
            const bool valid = p < prm.P;
            const int nvalid = min(32, prm.P - tile * 32);

            // ---- stage this tile's coefficient rows (async proxy), fetch the point meanwhile ----
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_expect_tx(&ws.bar, 2u * kBoxBytes);
                tma_load_3d(&ws.stage[0], &harm_map, 0, tile * 32, b, &ws.bar);
                tma_load_3d(&ws.stage[kBoxBytes / 4], &harm_map, 32, tile * 32, b, &ws.bar);
            }
            float px = 0.f, py = 0.f, pz = 0.f;
            if (valid) {
                const float *pp = prm.pts + (static_cast<size_t>(b) * prm.P + p) * prm.pts_dim;
                px = __ldg(pp);
                py = __ldg(pp + 1);
                pz = __ldg(pp + 2);
            }
            mbar_wait(&ws.bar, parity);
            parity ^= 1u;

            float g[PACKED ? 1 : 64];
            float g0[PACKED ? 8 : 1];
            unsigned long long gp[PACKED ? 28 : 1];
            {
                // element (row r, 16-B chunk c) of a SWIZZLE_128B box sits at r*128 + ((c ^ (r & 7)) << 4)
                float h[64];
                const unsigned char *row = reinterpret_cast<const unsigned char *>(ws.stage) + lane * 128;
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const float4 v = *reinterpret_cast<const float4 *>(row + (q >> 3) * kBoxBytes +
                                                                       (((q & 7) ^ (lane & 7)) << 4));
                    h[4 * q + 0] = v.x;
                    h[4 * q + 1] = v.y;
                    h[4 * q + 2] = v.z;
                    h[4 * q + 3] = v.w;
                }
                load_coefficients<PACKED>(h, g, g0, gp, SIGMOID ? kNegLog2e : 1.0f);
            }
            __syncwarp();  // every lane has its row in registers: the staging rows become `red`

            if (REDUCE && nvalid < 32) {  // tail tile: columns of absent points must read as zero
                for (int j = 0; j < 32; ++j) red[j * kRedStride + lane] = 0.f;
                __syncwarp();
            }

            // ---- camera sweep, two rays per iteration, software-pipelined by hand: iteration jj evaluates
            // pair jj (rays normalised in iteration jj-1), finishes the activation of pair jj-1 (its ex2 was
            // issued in iteration jj-1) and normalises the rays of pair jj+1, so that no LDS / MUFU result
            // is consumed in the iteration that produced it.
            Ray rA = make_ray(ws.cams[0], px, py, pz), rB = make_ray(ws.cams[1], px, py, pz);
            float eA = 0.f, eB = 0.f;
            const int npairs = ncam32 >> 1;
#pragma unroll 1
            for (int jj = 0; jj < npairs; ++jj) {
                const float4 nA = ws.cams[2 * jj + 2], nB = ws.cams[2 * jj + 3];  // table has 2 pad entries
                const Ray qA = make_ray(nA, px, py, pz), qB = make_ray(nB, px, py, pz);
                float zA, zB;
                eval_two_rays<EVAL>(g, g0, gp, rA, rB, zA, zB);
                const float vA = activation_finish<SIGMOID>(eA), vB = activation_finish<SIGMOID>(eB);
                // camera (within the chunk) of the pair being finished; at jj == 0 there is none: c = -2 and the
                // value lands in rows 30/31 of the transpose buffer, which are rewritten before they are summed.
                const int c = 2 * jj - 2;
                if (REDUCE) {
                    if (valid) {
                        red[(c & 31) * kRedStride + lane] = vA;
                        red[((c & 31) + 1) * kRedStride + lane] = vB;
                    }
                } else if (valid) {
                    float *o = prm.out + (static_cast<size_t>(b) * prm.C + cam0 + c) * prm.P + p;
                    if (c >= 0 && c < ncam) o[0] = vA;
                    if (c >= 0 && c + 1 < ncam) o[prm.P] = vB;
                }
                eA = activation_start<SIGMOID>(zA);
                eB = activation_start<SIGMOID>(zB);
                rA = qA;
                rB = qB;
                if (REDUCE && (jj & 15) == 0 && jj > 0) bad |= reduce_batch<SIGMOID>(red, ws.wacc, c & ~31, lane);
            }
            {
                const float vA = activation_finish<SIGMOID>(eA), vB = activation_finish<SIGMOID>(eB);
                const int c = ncam32 - 2;
                if (REDUCE) {
                    if (valid) {
                        red[(c & 31) * kRedStride + lane] = vA;
                        red[((c & 31) + 1) * kRedStride + lane] = vB;
                    }
                    bad |= reduce_batch<SIGMOID>(red, ws.wacc, c & ~31, lane);
                } else if (valid) {
                    float *o = prm.out + (static_cast<size_t>(b) * prm.C + cam0 + c) * prm.P + p;
                    if (c < ncam) o[0] = vA;
                    if (c + 1 < ncam) o[prm.P] = vB;
                }
            }
        }

        // next task of this warp: dynamic for the reducing kernel (the workspace holds the counter), strided otherwise
        if (REDUCE) {
            int nt = 0;
            if (lane == 0) nt = static_cast<int>(atomicAdd(prm.next_task, 1u)) + static_cast<int>(gridDim.x) * WARPS;
            task = __shfl_sync(0xffffffffu, nt, 0);
        } else {
            task += static_cast<int>(gridDim.x) * WARPS;
        }
    }
    flush();

    if (REDUCE) {
        // ---- last CTA: fixed-point accumulators -> fp32 mean; leave the workspace zeroed ----
        __shared__ bool is_last;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            is_last = (atomicAdd(prm.done, 1u) == gridDim.x - 1);
        }
        __syncthreads();
        if (is_last && !prm.finalize) {
            if (threadIdx.x == 0) *prm.done = 0u, *prm.next_task = 0u;   // partial sums stay in the workspace for the next point slice
        } else if (is_last && prm.push_partial) {
            __threadfence();
            const size_t BC = static_cast<size_t>(prm.B) * prm.C;
            for (size_t i = threadIdx.x; i < BC; i += WARPS * 32) {
                const unsigned long long q = atomicExch(prm.acc + i, 0ull);
                const unsigned int bad = atomicExch(prm.flags + i, 0u);
                for (int r = 0; r < prm.push_world; ++r) {   // peer-mapped stores into slot push_rank
                    prm.part_dst[r][prm.push_rank * BC + i] = q;
                    reinterpret_cast<unsigned int *>(prm.part_dst[r] + prm.push_world * BC)[prm.push_rank * BC + i] = bad;
                }
            }
            __threadfence_system();
            __syncthreads();
            if (threadIdx.x < prm.push_world) st_release_sys(prm.push_flag[threadIdx.x] + prm.push_rank, prm.push_epoch);
            if (threadIdx.x == 0) *prm.done = 0u, *prm.next_task = 0u;
            __syncthreads();
            if (wait_for_ranks<WARPS * 32>(prm.push_flag[prm.push_rank], prm.push_world, prm.push_epoch, prm.status)) {
                const unsigned long long *mine = prm.part_dst[prm.push_rank];
                const unsigned int *mine_bad = reinterpret_cast<const unsigned int *>(mine + prm.push_world * BC);
                float *scores = prm.push_dst[prm.push_rank];
                const double unfix = 1.0 / static_cast<double>(SIGMOID ? kFixSigmoid : kFixRelu);
                for (size_t i = threadIdx.x; i < BC; i += WARPS * 32) {
                    long long total = 0;
                    unsigned int bad = 0;
                    for (int r = 0; r < prm.push_world; ++r) {
                        total += static_cast<long long>(__ldcg(mine + r * BC + i));   // integer sum: order-independent, exact
                        bad |= __ldcg(mine_bad + r * BC + i);
                    }
                    const float t = static_cast<float>(static_cast<double>(total) * unfix);
                    const float score = bad ? __int_as_float(0x7fc00000) : __fdiv_rn(t, static_cast<float>(prm.mean_count));
                    scores[i] = score;
                    if (prm.out) prm.out[i] = score;
                }
                __threadfence();
                __syncthreads();
                argmax_rows<WARPS * 32>(scores, prm.B, prm.C, prm.best);
            }
        } else if (is_last) {
            __threadfence();
            const int nloc = prm.cam_end - prm.cam_begin;
            const double unfix = 1.0 / static_cast<double>(SIGMOID ? kFixSigmoid : kFixRelu);
            for (int i = threadIdx.x; i < prm.B * nloc; i += WARPS * 32) {
                const size_t idx = static_cast<size_t>(i / nloc) * prm.C + prm.cam_begin + (i % nloc);
                const long long q = static_cast<long long>(atomicExch(prm.acc + idx, 0ull));
                const unsigned int bad = atomicExch(prm.flags + idx, 0u);
                const float total = static_cast<float>(static_cast<double>(q) * unfix);
                const float score = bad ? __int_as_float(0x7fc00000) : __fdiv_rn(total, static_cast<float>(prm.mean_count));
                if (prm.out) prm.out[idx] = score;
                for (int r = 0; r < prm.push_world; ++r) prm.push_dst[r][idx] = score;  // peer-mapped stores
            }
            if (prm.push_world > 0) {
                __threadfence_system();  // every thread: its score stores are visible system-wide ...
                __syncthreads();
                if (threadIdx.x < prm.push_world)  // ... before the arrival flag of this rank is
                    st_release_sys(prm.push_flag[threadIdx.x] + prm.push_rank, prm.push_epoch);
            }
            if (threadIdx.x == 0) *prm.done = 0u, *prm.next_task = 0u;
            if (prm.push_world > 0 && prm.best) {
                __syncthreads();
                wait_scores_and_argmax<WARPS * 32>(prm.push_dst[prm.push_rank], prm.push_flag[prm.push_rank], prm.push_world,
                                                   prm.push_epoch, prm.B, prm.C, prm.best, prm.status);
            }
        }
    }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// 3-D tensor map over the harmonics (fastest first: 64 coefficients, P points, B clouds); box = 32 coefficients
// (128 B) x 32 points, SWIZZLE_128B.  Out-of-range points of the last tile of a cloud are zero-filled.
int make_harmonics_map(CUtensorMap *map, const float *harm, int B, int P)
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
    const cuuint64_t dims[3] = {MAC_N_HARMONICS, static_cast<cuuint64_t>(P), static_cast<cuuint64_t>(B)};
    const cuuint64_t strides[2] = {MAC_N_HARMONICS * sizeof(float),
                                   static_cast<cuuint64_t>(P) * MAC_N_HARMONICS * sizeof(float)};
    const cuuint32_t box[3] = {32, 32, 1};
    const cuuint32_t elem[3] = {1, 1, 1};
    const CUresult rc = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(harm), dims, strides, box,
                               elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(rc));
        return MAC_ERR_CUDA;
    }
    return MAC_OK;
}

int plan_and_launch(bool reduce, const float *pts, int pts_dim, const float *harm, const float *cams, float *out,
                    int B, int P, int C, int cam_begin, int cam_end, int act, void *workspace,
                    size_t workspace_bytes, void *stream, const mac_peer_board_t *board = nullptr, int mean_count = 0,
                    int finalize = 1, long long *best = nullptr, int *status = nullptr, bool push_partial = false)
{
    MAC_REQUIRE(pts && harm && cams && (out || board), "null tensor pointer");
    if (board) {
        MAC_REQUIRE(reduce, "score push needs the reducing kernel");
        MAC_REQUIRE(board->world >= 1 && board->world <= MAC_MAX_PEERS && board->rank >= 0 && board->rank < board->world,
                    "bad peer board: rank %d of %d (max %d peers)", board->rank, board->world, MAC_MAX_PEERS);
        for (int r = 0; r < board->world; ++r)
            MAC_REQUIRE(board->scores[r] && board->flags[r], "peer board pointer %d is null", r);
    }
    MAC_REQUIRE(B > 0 && P > 0 && C > 0, "B, P, C must be positive (got %d, %d, %d)", B, P, C);
    MAC_REQUIRE(pts_dim >= 3, "pts_dim must be >= 3 (got %d)", pts_dim);
    MAC_REQUIRE(0 <= cam_begin && cam_begin <= cam_end && cam_end <= C, "bad camera range [%d, %d) for C=%d",
                cam_begin, cam_end, C);
    MAC_REQUIRE(act == MAC_ACT_RELU || act == MAC_ACT_SIGMOID, "act must be MAC_ACT_RELU or MAC_ACT_SIGMOID");
    MAC_REQUIRE((reinterpret_cast<uintptr_t>(harm) & 15u) == 0, "harmonics must be 16-byte aligned");
    if (cam_begin == cam_end && !board) return MAC_OK;

    CovgainParams prm{};
    prm.pts = pts;
    prm.harm = harm;
    prm.cams = cams;
    prm.out = out;
    prm.pts_dim = pts_dim;
    prm.B = B;
    prm.P = P;
    prm.C = C;
    prm.cam_begin = cam_begin;
    prm.cam_end = cam_end;
    prm.mean_count = mean_count > 0 ? mean_count : P;
    prm.finalize = finalize;
    if (board) {
        prm.push_world = board->world;
        prm.push_rank = board->rank;
        prm.push_epoch = board->epoch;
        for (int r = 0; r < board->world; ++r) {
            prm.push_dst[r] = board->scores[r];
            prm.push_flag[r] = board->flags[r];
        }
        prm.best = best;
        prm.status = status;
        if (push_partial) {
            prm.push_partial = 1;
            for (int r = 0; r < board->world; ++r) {
                MAC_REQUIRE(board->partials[r], "peer board: partial region %d is null", r);
                prm.part_dst[r] = static_cast<unsigned long long *>(board->partials[r]);
            }
        }
    }
    if (reduce) {
        const size_t need = mac_covgain_workspace_bytes(B, C);
        if (!workspace || workspace_bytes < need) {
            set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
            return MAC_ERR_WORKSPACE;
        }
        MAC_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0, "workspace must be 16-byte aligned");
        unsigned char *w = static_cast<unsigned char *>(workspace);
        prm.acc = reinterpret_cast<unsigned long long *>(w);
        w += align_up(static_cast<size_t>(B) * C * sizeof(unsigned long long), 16);
        prm.flags = reinterpret_cast<unsigned int *>(w);
        w += align_up(static_cast<size_t>(B) * C * sizeof(unsigned int), 16);
        prm.done = reinterpret_cast<unsigned int *>(w);
        prm.next_task = prm.done + 1;
    }

    CUtensorMap harm_map;
    if (int rc = make_harmonics_map(&harm_map, harm, B, P)) return rc;

    int device = 0;
    MAC_CUDA(cudaGetDevice(&device));
    // MAC_COVGAIN_VARIANT (ablation knob for tools/bench_covgain.py): 0 = default, scalar FFMA Horner steps of the
    // P + u_x Q form; 1 = scalar A/B form; 2 = packed fma.rn.f32x2 Horner steps (bitwise identical to 0, 2.7 % slower:
    // an FFMA2 holds the issue port for two cycles).
    static const int variant = [] {
        const char *e = getenv("MAC_COVGAIN_VARIANT");
        return e ? atoi(e) : 0;
    }();
    // 8 warps x 2 CTAs/SM (128 registers) measured fastest; 12 x 1 (168 regs) and 16 x 1 were 9-11 % slower
    // (DESIGN.md section 4.3).
    const int warps_per_cta = 8;
    const int ctas_per_sm = 2;
    const long long slots = static_cast<long long>(sm_count(device)) * ctas_per_sm * warps_per_cta;  // resident warps
    const int n_local = cam_end - cam_begin;
    prm.tiles_per_cloud = (P + 31) / 32;
    // Camera chunk per warp task: as large as possible (amortises the per-tile load + change of basis)
    // while leaving enough tasks to balance the machine.
    int cpt = kCamChunkMax;
    while (cpt > 32) {
        const long long tasks = static_cast<long long>(B) * prm.tiles_per_cloud * ((n_local + cpt - 1) / cpt);
        if (cpt / 2 >= n_local || tasks < 2 * slots) cpt /= 2;   // dynamic scheduling balances ~2.5 tasks per warp well
        else break;
    }
    while (cpt > 32 && cpt / 2 >= n_local) cpt /= 2;
    // MAC_COVGAIN_CPT / MAC_COVGAIN_TPT override the heuristics (tuning knobs for tools/bench_covgain.py)
    static const int force_cpt = [] { const char *e = getenv("MAC_COVGAIN_CPT"); return e ? atoi(e) : 0; }();
    static const int force_tpt = [] { const char *e = getenv("MAC_COVGAIN_TPT"); return e ? atoi(e) : 0; }();
    if (force_cpt == 32 || force_cpt == 64 || force_cpt == 96 || force_cpt == 128) cpt = force_cpt;
    prm.cams_per_task = cpt;
    prm.n_cam_chunks = (n_local + cpt - 1) / cpt;
    // One point tile per task: the warps are persistent and pull tasks dynamically, so small tasks only cost one
    // counter atomic each and keep the tail short (several tiles per task only for enormous launches).
    const long long base_tasks = static_cast<long long>(B) * prm.tiles_per_cloud * prm.n_cam_chunks;
    long long tpt = base_tasks / (256 * slots);
    if (tpt < 1) tpt = 1;
    if (tpt > 64) tpt = 64;
    if (force_tpt > 0) tpt = force_tpt;
    prm.tiles_per_task = static_cast<int>(tpt);
    prm.runs_per_cloud = (prm.tiles_per_cloud + prm.tiles_per_task - 1) / prm.tiles_per_task;
    const long long total = static_cast<long long>(B) * prm.runs_per_cloud * prm.n_cam_chunks;
    MAC_REQUIRE(total < (1ll << 31) - 64, "problem too large for one launch (%lld tasks)", total);
    prm.total_tasks = static_cast<int>(total);

    long long n_ctas = total > 0 ? (total + warps_per_cta - 1) / warps_per_cta : 1;
    if (n_ctas > slots / warps_per_cta) n_ctas = slots / warps_per_cta;   // persistent: one resident wave
    const dim3 grid(static_cast<unsigned>(n_ctas));
    const size_t smem = sizeof(WarpSmem) * warps_per_cta;
    cudaStream_t st = static_cast<cudaStream_t>(stream);

#define MAC_LAUNCH_V(SIG, RED, EVAL, WARPS, MINB)                                                                \
    do {                                                                                                          \
        static DeviceOnce once;                                                                                   \
        if (int rc = ensure_dynamic_smem(once, covgain_kernel<SIG, RED, EVAL, WARPS, MINB>, static_cast<int>(smem))) \
            return rc;                                                                                            \
        covgain_kernel<SIG, RED, EVAL, WARPS, MINB><<<grid, WARPS * 32, smem, st>>>(prm, harm_map);             \
    } while (0)
#define MAC_LAUNCH(SIG, RED)                                       \
    do {                                                           \
        switch (variant) {                                         \
        case 1: MAC_LAUNCH_V(SIG, RED, 1, 8, 2); break;            \
        case 2: MAC_LAUNCH_V(SIG, RED, 2, 8, 2); break;            \
        default: MAC_LAUNCH_V(SIG, RED, 0, 8, 2); break;           \
        }                                                          \
    } while (0)

    if (reduce) {
        if (act == MAC_ACT_SIGMOID) MAC_LAUNCH(true, true);
        else MAC_LAUNCH(false, true);
    } else {
        if (act == MAC_ACT_SIGMOID) MAC_LAUNCH(true, false);
        else MAC_LAUNCH(false, false);
    }
#undef MAC_LAUNCH_V
#undef MAC_LAUNCH
    MAC_CUDA(cudaGetLastError());
    count_launch();
    return MAC_OK;
}

}  // namespace
}  // namespace mac

extern "C" size_t mac_covgain_workspace_bytes(int B, int C)
{
    if (B <= 0 || C <= 0) return 0;
    const size_t n = static_cast<size_t>(B) * C;
    return mac::align_up(n * sizeof(unsigned long long), 16) + mac::align_up(n * sizeof(unsigned int), 16) + 16;
}

extern "C" int mac_covgain_f32(const float *pts, int pts_dim, const float *harmonics, const float *cams, float *out,
                               int B, int P, int C, int cam_begin, int cam_end, int act, void *workspace,
                               size_t workspace_bytes, void *stream)
{
    return mac::plan_and_launch(true, pts, pts_dim, harmonics, cams, out, B, P, C, cam_begin, cam_end, act,
                                workspace, workspace_bytes, stream);
}

namespace mac {
// Point-sliced form used by the host-buffer entry point: accumulate the slice [pts, pts + P) of a cloud of `p_total`
// points into the workspace; the call with finalize = 1 converts the sums to the mean over p_total.
int covgain_accumulate(const float *pts, int pts_dim, const float *harmonics, const float *cams, float *out, int P, int C,
                       int cam_begin, int cam_end, int act, void *workspace, size_t workspace_bytes, int p_total,
                       int finalize, cudaStream_t stream)
{
    return plan_and_launch(true, pts, pts_dim, harmonics, cams, out, 1, P, C, cam_begin, cam_end, act, workspace,
                                workspace_bytes, stream, nullptr, p_total, finalize);
}
}  // namespace mac

extern "C" int mac_visibility_f32(const float *pts, int pts_dim, const float *harmonics, const float *cams,
                                  float *out, int B, int P, int C, int cam_begin, int cam_end, int act, void *stream)
{
    return mac::plan_and_launch(false, pts, pts_dim, harmonics, cams, out, B, P, C, cam_begin, cam_end, act, nullptr,
                                0, stream);
}

extern "C" int mac_covgain_push_f32(const float *pts, int pts_dim, const float *harmonics, const float *cams, int B,
                                    int P, int C, int cam_begin, int cam_end, int act, void *workspace,
                                    size_t workspace_bytes, const mac_peer_board_t *board, void *stream)
{
    if (!board) {
        mac::set_error("mac_covgain_push_f32 needs a peer board");
        return MAC_ERR_INVALID_ARGUMENT;
    }
    return mac::plan_and_launch(true, pts, pts_dim, harmonics, cams, nullptr, B, P, C, cam_begin, cam_end, act,
                                workspace, workspace_bytes, stream, board);
}

extern "C" int mac_covgain_push_argmax_f32(const float *pts, int pts_dim, const float *harmonics, const float *cams, int B,
                                           int P, int C, int cam_begin, int cam_end, int act, void *workspace,
                                           size_t workspace_bytes, const mac_peer_board_t *board, long long *best,
                                           int *status, void *ev_begin, void *ev_end, void *stream)
{
    if (!board || !best || !status) {
        mac::set_error("mac_covgain_push_argmax_f32 needs a peer board, best and status");
        return MAC_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (ev_begin) MAC_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(ev_begin), st));
    // the finishing CTA of the scoring kernel also waits for the peers and takes the argmax: one launch per step
    if (int rc = mac::plan_and_launch(true, pts, pts_dim, harmonics, cams, nullptr, B, P, C, cam_begin, cam_end, act, workspace,
                                      workspace_bytes, stream, board, 0, 1, best, status))
        return rc;
    if (ev_end) MAC_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(ev_end), st));
    return MAC_OK;
}

extern "C" size_t mac_covgain_partial_region_bytes(int world, int B, int C)
{
    if (world <= 0 || B <= 0 || C <= 0) return 0;
    return static_cast<size_t>(world) * B * C * (sizeof(unsigned long long) + sizeof(unsigned int));
}

extern "C" int mac_covgain_push_partial_argmax_f32(const float *pts, int pts_dim, const float *harmonics, const float *cams,
                                                   int B, int P_local, int p_total, int C, int act, void *workspace,
                                                   size_t workspace_bytes, const mac_peer_board_t *board, long long *best,
                                                   int *status, void *stream)
{
    if (!board || !best || !status) {
        mac::set_error("mac_covgain_push_partial_argmax_f32 needs a peer board, best and status");
        return MAC_ERR_INVALID_ARGUMENT;
    }
    if (p_total < P_local || P_local <= 0) {
        mac::set_error("bad point partition: P_local=%d of p_total=%d", P_local, p_total);
        return MAC_ERR_INVALID_ARGUMENT;
    }
    return mac::plan_and_launch(true, pts, pts_dim, harmonics, cams, nullptr, B, P_local, C, 0, C, act, workspace,
                                workspace_bytes, stream, board, p_total, 1, best, status, true);
}

extern "C" int mac_covgain_accumulate_f32(const float *pts, int pts_dim, const float *harmonics, const float *cams, int P_slice,
                                          int C, int act, void *workspace, size_t workspace_bytes, void *stream)
{
    // finalize = 0: the exact partial sums of this slice of ONE cloud stay in the workspace; a later
    // mac_covgain_f32-family call on the same workspace (e.g. mac_covgain_push_partial_argmax_f32 with the last slice) adds
    // its own points and finishes.  mean_count is irrelevant here.
    return mac::plan_and_launch(true, pts, pts_dim, harmonics, cams, reinterpret_cast<float *>(workspace) /* unused, non-null */,
                                1, P_slice, C, 0, C, act, workspace, workspace_bytes, stream, nullptr, P_slice, 0);
}
