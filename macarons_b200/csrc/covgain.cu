// Coverage-gain integration: for every (cloud b, camera c) the mean over surface points p of
//     act( sum_k Y_k(direction of cam_c - pt_p) * H[b,p,k] ),   k = l*l+l+m, l < 8  (64 terms)
// Replaces SconeVis.compute_coverage_gain / compute_visibilities and
// Macarons.compute_visibility_gains (reference networks/SconeVis.py:164-252, Macarons.py:138-178).
//
// Design (sm_100a, CUDA cores; see DESIGN.md section "coverage-gain kernel"):
//   * one surface point per lane.  Its 64 SH coefficients are staged global -> shared with the bulk
//     async-copy engine (one 256-B cp.async.bulk per lane into a padded row, mbarrier completion),
//     pulled into registers with 16 conflict-free LDS.128 and converted ONCE to 64 "Horner-ready"
//     coefficients (sh_horner_gen.h): the SH sum becomes Re sum_m (A_m(ct) - i B_m(ct)) (uz + i ux)^m
//     with u the unit ray, i.e. 49 + 26 FMAs per (point, camera) pair and no trigonometry.
//   * the warp then sweeps a chunk of cameras (broadcast LDS.128 from a per-warp table).  Per-pair
//     values go through a padded per-warp transpose buffer so that the sum over the 32 points costs
//     ~2 instructions per pair instead of a 10-instruction shuffle tree.
//   * per-task partial sums are accumulated as exact fixed-point int64 with one atomic per
//     (task, camera); integer addition is associative, so the result is bitwise independent of
//     scheduling, launch geometry and of how cameras are partitioned across GPUs.  The last CTA to
//     finish converts the accumulators to the fp32 mean and re-zeroes the workspace.
#include "mac_common.h"
#include "sh_horner_gen.h"

namespace mac {

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kRowFloats = 68;     // 64 coefficients + 4 pad: rows 272 B apart -> LDS.128 conflict-free
constexpr int kRedStride = 36;     // transpose buffer row stride (floats)
constexpr int kCamChunkMax = 128;  // cameras per warp task

constexpr float kNegLog2e = -1.4426950408889634f;
constexpr float kFixSigmoid = 4294967296.0f;  // 2^32: sums of sigmoid values, |s| <= 32 * tiles_per_task
constexpr float kFixRelu = 16777216.0f;       // 2^24: sums of relu values (unbounded inputs)

struct __align__(128) WarpSmem {
    float stage[32 * kRowFloats];  // bulk-copy landing zone; re-used as the 32x36 transpose buffer
    float4 cams[kCamChunkMax];
    float wacc[kCamChunkMax];
    uint64_t bar;
    uint64_t pad_[15];
};
static_assert(sizeof(WarpSmem) % 128 == 0, "per-warp shared block must keep 128-B alignment");
static_assert(32 * kRedStride <= 32 * kRowFloats, "transpose buffer must fit in the staging rows");

struct CovgainParams {
    const float *pts;
    const float *harm;
    const float *cams;
    float *out;
    unsigned long long *acc;   // (B*C) fixed-point accumulators   [REDUCE only]
    unsigned int *flags;       // (B*C) non-finite markers         [REDUCE only]
    unsigned int *done;        // CTA completion ticket            [REDUCE only]
    int pts_dim;
    int B, P, C;
    int cam_begin, cam_end;
    int cams_per_task;         // multiple of 32, <= kCamChunkMax
    int n_cam_chunks;
    int tiles_per_cloud;       // ceil(P / 32)
    int tiles_per_task;
    int runs_per_cloud;        // ceil(tiles_per_cloud / tiles_per_task)
    int total_tasks;           // B * runs_per_cloud * n_cam_chunks
};

template <bool SIGMOID>
__device__ __forceinline__ float pair_value(const float (&g)[64], float px, float py, float pz, float4 cam)
{
    const float dx = cam.x - px, dy = cam.y - py, dz = cam.z - pz;
    const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    const float rinv = rsqrt_approx(r2);
    const float z = mac_sh_eval(g, dx * rinv, dy * rinv, dz * rinv);
    if (SIGMOID) {
        // g was pre-scaled by -log2(e): z = -x*log2(e), sigmoid(x) = 1 / (1 + 2^z)
        return rcp_approx(1.0f + ex2_approx(z));
    }
    return fmaxf(z, 0.0f);
}

template <bool SIGMOID, bool REDUCE>
__global__ void __launch_bounds__(kThreads, 2) covgain_kernel(const CovgainParams prm)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    WarpSmem &ws = reinterpret_cast<WarpSmem *>(smem_raw)[warp];
    float *red = ws.stage;

    if (lane == 0) {
        mbar_init(&ws.bar, 1);
        fence_mbar_init();
    }
    __syncwarp();

    const int task = blockIdx.x * kWarpsPerCta + warp;
    if (task < prm.total_tasks) {
        // task -> (cloud, run of point tiles, camera chunk); camera chunk fastest so that the warps of
        // one CTA share their coefficient rows through L1/L2.
        const int cc = task % prm.n_cam_chunks;
        const int rest = task / prm.n_cam_chunks;
        const int run = rest % prm.runs_per_cloud;
        const int b = rest / prm.runs_per_cloud;
        const int cam0 = prm.cam_begin + cc * prm.cams_per_task;
        const int ncam = min(prm.cams_per_task, prm.cam_end - cam0);
        const int ncam32 = (ncam + 31) & ~31;

        for (int j = lane; j < ncam32; j += 32) {
            float4 c = make_float4(0.f, 0.f, 1048576.f, 0.f);  // padding camera: finite, result discarded
            if (j < ncam) {
                const float *src = prm.cams + (static_cast<size_t>(b) * prm.C + cam0 + j) * 3;
                c = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), 0.f);
            }
            ws.cams[j] = c;
            ws.wacc[j] = 0.f;
        }

        uint32_t parity = 0;
        const int tile_end = min(prm.tiles_per_cloud, (run + 1) * prm.tiles_per_task);
        for (int tile = run * prm.tiles_per_task; tile < tile_end; ++tile) {
            const int p = tile * 32 + lane;
            const bool valid = p < prm.P;
            const int nvalid = min(32, prm.P - tile * 32);

            // ---- stage this tile's coefficient rows (async proxy), fetch the point meanwhile ----
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_expect_tx(&ws.bar, static_cast<uint32_t>(nvalid) * 256u);
            __syncwarp();
            float px = 0.f, py = 0.f, pz = 0.f;
            if (valid) {
                const size_t row = static_cast<size_t>(b) * prm.P + p;
                bulk_g2s(&ws.stage[lane * kRowFloats], prm.harm + row * MAC_N_HARMONICS, 256u, &ws.bar);
                const float *pp = prm.pts + row * prm.pts_dim;
                px = __ldg(pp);
                py = __ldg(pp + 1);
                pz = __ldg(pp + 2);
            }
            mbar_wait(&ws.bar, parity);
            parity ^= 1u;

            float g[64];
            {
                float h[64];
                const float4 *row4 = reinterpret_cast<const float4 *>(&ws.stage[lane * kRowFloats]);
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const float4 v = valid ? row4[q] : make_float4(0.f, 0.f, 0.f, 0.f);
                    h[4 * q + 0] = v.x;
                    h[4 * q + 1] = v.y;
                    h[4 * q + 2] = v.z;
                    h[4 * q + 3] = v.w;
                }
                mac_sh_pretransform(h, g, SIGMOID ? kNegLog2e : 1.0f);
            }
            __syncwarp();  // every lane has its row in registers: the staging rows become `red`

            if (REDUCE && nvalid < 32) {  // tail tile: columns of absent points must read as zero
                for (int j = 0; j < 32; ++j) red[j * kRedStride + lane] = 0.f;
                __syncwarp();
            }

            for (int cb = 0; cb < ncam32; cb += 32) {
#pragma unroll 2
                for (int j = 0; j < 32; ++j) {
                    const float v = pair_value<SIGMOID>(g, px, py, pz, ws.cams[cb + j]);
                    if (REDUCE) {
                        if (valid) red[j * kRedStride + lane] = v;
                    } else if (valid && cb + j < ncam) {
                        prm.out[(static_cast<size_t>(b) * prm.C + cam0 + cb + j) * prm.P + p] = v;
                    }
                }
                if (REDUCE) {
                    __syncwarp();
                    const float4 *r4 = reinterpret_cast<const float4 *>(&red[lane * kRedStride]);
                    float s = 0.f;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 v = r4[q];
                        s += (v.x + v.y) + (v.z + v.w);
                    }
                    ws.wacc[cb + lane] += s;  // lane j owns camera cb + j
                    __syncwarp();
                }
            }
        }

        if (REDUCE) {
            for (int j = lane; j < ncam; j += 32) {
                const float s = ws.wacc[j];
                const size_t idx = static_cast<size_t>(b) * prm.C + cam0 + j;
                if (fabsf(s) < 1.0e12f) {
                    const long long q = __float2ll_rn(s * (SIGMOID ? kFixSigmoid : kFixRelu));
                    atomicAdd(prm.acc + idx, static_cast<unsigned long long>(q));
                } else {
                    atomicOr(prm.flags + idx, 1u);  // NaN / inf / overflow -> result is NaN
                }
            }
        }
    }

    if (REDUCE) {
        // ---- last CTA: fixed-point accumulators -> fp32 mean; leave the workspace zeroed ----
        __shared__ bool is_last;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            is_last = (atomicAdd(prm.done, 1u) == gridDim.x - 1);
        }
        __syncthreads();
        if (is_last) {
            __threadfence();
            const int nloc = prm.cam_end - prm.cam_begin;
            const double unfix = 1.0 / static_cast<double>(SIGMOID ? kFixSigmoid : kFixRelu);
            for (int i = threadIdx.x; i < prm.B * nloc; i += kThreads) {
                const size_t idx = static_cast<size_t>(i / nloc) * prm.C + prm.cam_begin + (i % nloc);
                const long long q = static_cast<long long>(atomicExch(prm.acc + idx, 0ull));
                const unsigned int bad = atomicExch(prm.flags + idx, 0u);
                const float total = static_cast<float>(static_cast<double>(q) * unfix);
                prm.out[idx] = bad ? __int_as_float(0x7fc00000) : __fdiv_rn(total, static_cast<float>(prm.P));
            }
            if (threadIdx.x == 0) *prm.done = 0u;
        }
    }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int plan_and_launch(bool reduce, const float *pts, int pts_dim, const float *harm, const float *cams, float *out,
                    int B, int P, int C, int cam_begin, int cam_end, int act, void *workspace,
                    size_t workspace_bytes, void *stream)
{
    MAC_REQUIRE(pts && harm && cams && out, "null tensor pointer");
    MAC_REQUIRE(B > 0 && P > 0 && C > 0, "B, P, C must be positive (got %d, %d, %d)", B, P, C);
    MAC_REQUIRE(pts_dim >= 3, "pts_dim must be >= 3 (got %d)", pts_dim);
    MAC_REQUIRE(0 <= cam_begin && cam_begin <= cam_end && cam_end <= C, "bad camera range [%d, %d) for C=%d",
                cam_begin, cam_end, C);
    MAC_REQUIRE(act == MAC_ACT_RELU || act == MAC_ACT_SIGMOID, "act must be MAC_ACT_RELU or MAC_ACT_SIGMOID");
    MAC_REQUIRE((reinterpret_cast<uintptr_t>(harm) & 15u) == 0, "harmonics must be 16-byte aligned");
    if (cam_begin == cam_end) return MAC_OK;

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
    }

    int device = 0;
    MAC_CUDA(cudaGetDevice(&device));
    const long long slots = static_cast<long long>(sm_count(device)) * 2 * kWarpsPerCta;  // resident warps
    const int n_local = cam_end - cam_begin;
    prm.tiles_per_cloud = (P + 31) / 32;
    // Camera chunk per warp task: as large as possible (amortises the per-tile load + change of basis)
    // while leaving enough tasks to balance the machine.
    int cpt = kCamChunkMax;
    while (cpt > 32) {
        const long long tasks = static_cast<long long>(B) * prm.tiles_per_cloud * ((n_local + cpt - 1) / cpt);
        if (cpt / 2 >= n_local || tasks < 6 * slots) cpt /= 2;
        else break;
    }
    while (cpt > 32 && cpt / 2 >= n_local) cpt /= 2;
    prm.cams_per_task = cpt;
    prm.n_cam_chunks = (n_local + cpt - 1) / cpt;
    // Several point tiles per task only when there are far more tasks than warp slots (fewer atomics).
    const long long base_tasks = static_cast<long long>(B) * prm.tiles_per_cloud * prm.n_cam_chunks;
    long long tpt = base_tasks / (24 * slots);
    if (tpt < 1) tpt = 1;
    if (tpt > 64) tpt = 64;
    prm.tiles_per_task = static_cast<int>(tpt);
    prm.runs_per_cloud = (prm.tiles_per_cloud + prm.tiles_per_task - 1) / prm.tiles_per_task;
    const long long total = static_cast<long long>(B) * prm.runs_per_cloud * prm.n_cam_chunks;
    MAC_REQUIRE(total < (1ll << 31) - kWarpsPerCta, "problem too large for one launch (%lld tasks)", total);
    prm.total_tasks = static_cast<int>(total);

    const dim3 grid(static_cast<unsigned>((total + kWarpsPerCta - 1) / kWarpsPerCta));
    const size_t smem = sizeof(WarpSmem) * kWarpsPerCta;
    cudaStream_t st = static_cast<cudaStream_t>(stream);

#define MAC_LAUNCH(SIG, RED)                                                                                      \
    do {                                                                                                          \
        static bool attr_done = false;                                                                            \
        if (!attr_done) {                                                                                         \
            MAC_CUDA(cudaFuncSetAttribute(covgain_kernel<SIG, RED>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                          static_cast<int>(smem)));                                               \
            attr_done = true;                                                                                     \
        }                                                                                                         \
        covgain_kernel<SIG, RED><<<grid, kThreads, smem, st>>>(prm);                                              \
    } while (0)

    if (reduce) {
        if (act == MAC_ACT_SIGMOID) MAC_LAUNCH(true, true);
        else MAC_LAUNCH(false, true);
    } else {
        if (act == MAC_ACT_SIGMOID) MAC_LAUNCH(true, false);
        else MAC_LAUNCH(false, false);
    }
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

extern "C" int mac_visibility_f32(const float *pts, int pts_dim, const float *harmonics, const float *cams,
                                  float *out, int B, int P, int C, int cam_begin, int cam_end, int act, void *stream)
{
    return mac::plan_and_launch(false, pts, pts_dim, harmonics, cams, out, B, P, C, cam_begin, cam_end, act, nullptr,
                                0, stream);
}
