// View-state kernels (rows a10-a12 of the scope table):
//   view_state     spherical-histogram binning of the visited cameras around every point
//                  (reference utility/scone_utils.py:799-860, utils.floor_divide :113-117,
//                   CustomGeometry.get_spherical_coords :27-45)
//   view_harmonics projection of the (n_elev x n_azim) histogram onto the 64 SH basis functions
//                  (reference utility/scone_utils.py:934-960); never materialises the (B,P,64,98) product
//   gather_bins    permutation of the bins used by move_view_state_to_view_space (:863-930)
// HBM-bound element-wise work: one point per thread (binning) / per thread group (projection), coalesced rows.
#include <math.h>

#include "mac_common.h"

namespace mac {

namespace {

constexpr int kMaxBins = 128;

// torch.remainder on fp32 for d > 0: fmod, then shifted into [0, d).  fmodf itself (a bit-serial long division in libdevice,
// ~60 divergent instructions) is replaced by an exact equivalent for the small quotients that occur here (|x / d| < 2^22):
// with n = trunc(x / d) the remainder x - n d is exactly representable, so ONE fma returns it without rounding; n comes
// from a reciprocal multiply and is corrected by at most one step.  Non-finite or huge inputs take the library path.
__device__ __forceinline__ float exact_fmod_pos(float ax, float d, float rd)   // ax >= 0
{
    float n = truncf(ax * rd);
    float r = fmaf(-n, d, ax);
    if (r < 0.f) {
        n -= 1.f;
        r = fmaf(-n, d, ax);
    } else if (r >= d) {
        n += 1.f;
        r = fmaf(-n, d, ax);
    }
    return r;
}
__device__ __forceinline__ float torch_remainder(float x, float d, float rd)
{
    const float ax = fabsf(x);
    float m;
    if (ax < 4194304.f * d) {
        m = exact_fmod_pos(ax, d, rd);
        m = x < 0.f ? -m : m;   // fmod carries the sign of x (and -0.0 for exact multiples, like fmodf)
    } else {
        m = fmodf(x, d);
    }
    if (m != 0.f && m < 0.f) m += d;
    return m;
}

struct ViewStateParams {
    const float *pts;    // (B*P, pts_dim)
    const float *views;  // (V, 3)
    float *state;        // (B*P, n_bins)
    long long n_pts;
    int pts_dim, V, n_elev, n_azim;
    float elev_step, azim_step, half_elev_step, half_azim_step;
    float inv_elev_step, inv_azim_step;   // approximate reciprocals (exact_fmod_pos corrects the quotient)
    int elev_lo;         // Python -n_elev // 2
    int azim_hi;         // n_azim // 2
    int azim_wrap;       // Python -n_azim // 2
    int elev_shift;      // n_elev // 2
};

// atan2(y, x) without branches or library calls: atan(min/max) by a degree-8 minimax polynomial in (min/max)^2 (fitted at
// Chebyshev nodes; 1.2e-7 max error on [0, 1], i.e. fp32 rounding level), then the octant is restored with selects.
__device__ __forceinline__ float atan2_poly(float y, float x)
{
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float a = mx > 0.f ? __fdividef(mn, mx) : 0.f;
    const float s = a * a;
    float r = 0.0028340641874819994f;
    r = fmaf(r, s, -0.016005029901862144f);
    r = fmaf(r, s, 0.042587608098983765f);
    r = fmaf(r, s, -0.07495445758104324f);
    r = fmaf(r, s, 0.10636754333972931f);
    r = fmaf(r, s, -0.14202570915222168f);
    r = fmaf(r, s, 0.19992484152317047f);
    r = fmaf(r, s, -0.3333306610584259f);
    r = fmaf(r * s, a, a);
    r = ay > ax ? 1.57079632679489661923f - r : r;
    r = x < 0.f ? 3.14159265358979323846f - r : r;
    return y < 0.f ? -r : r;
}

// Bin of the ray pt -> view (scone_utils.py:815-849).  The reference obtains the angles as elev = asin(dy / r) and
// azim = +-acos(dz / (r cos(elev))) (CustomGeometry.py:27-45); the first version of this kernel followed that statement by
// statement with libdevice asinf / cosf / acosf (~330 issued instructions per ray with divergent slow paths, a dependent
// chain that made the kernel latency-bound, and -- like the reference itself -- ill-conditioned near the poles and near
// |cos(azim)| = 1, where the two fp32 evaluations disagree by up to 3e-4 rad).  The angles are now the well-conditioned
// elev = atan2(dy, rho), azim = atan2(dx, dz) (rho = sqrt(dx^2 + dz^2)), equal to the reference's definition in exact
// arithmetic and accurate to ~2e-7 rad everywhere; everything after the angles (torch.remainder based floor division, the
// half-step rounding, the Python `-n // 2` clamps, the truncating cast and the wrap) still follows the reference operation
// by operation.  Measured on 5 M rays against the reference's fp32 arithmetic: 8 differing bins (1.6e-6), all of them rays
// within 1e-5 rad of a bin boundary or within 2e-3 rad of a pole, where the reference's own azimuth is off by 0.1-2 rad.
__device__ __forceinline__ int view_bin(const ViewStateParams &p, float dx, float dy, float dz)
{
    const float rho2 = fmaf(dx, dx, dz * dz);
    const float rho = rho2 > 0.f ? rho2 * rsqrt_approx(rho2) : 0.f;
    const float elev = atan2_poly(dy, rho);
    // exactly above / below the point the azimuth is undefined; the reference's formula returns acos(+-0) = pi/2 there
    const float azim = rho2 > 0.f ? atan2_poly(dx, dz) : 1.57079632679489661923f;

    const float re = torch_remainder(elev, p.elev_step, p.inv_elev_step), ra = torch_remainder(azim, p.azim_step, p.inv_azim_step);
    float ie = (elev - re) / p.elev_step, ia = (azim - ra) / p.azim_step;
    if (re > p.half_elev_step) ie += 1.f;
    if (ra > p.half_azim_step) ia += 1.f;
    if (ie >= static_cast<float>(p.n_elev)) ie = static_cast<float>(p.n_elev - 1);
    if (ie < static_cast<float>(p.elev_lo)) ie = static_cast<float>(p.elev_lo);
    if (ia > static_cast<float>(p.azim_hi)) ia = static_cast<float>(p.azim_wrap);
    ie += static_cast<float>(p.elev_shift);
    if (ia < 0.f) ia += static_cast<float>(p.n_azim);
    const int n_bins = p.n_elev * p.n_azim;
    int idx = static_cast<int>(ie) * p.n_azim + static_cast<int>(ia);   // NaN rays -> 0 like a garbage long cast
    // Python modulo n_bins: one conditional step covers every finite ray (idx in [-n_azim, n_bins + n_azim))
    if (idx >= -n_bins && idx < 2 * n_bins) {
        if (idx >= n_bins) idx -= n_bins;
        if (idx < 0) idx += n_bins;
    } else {
        idx %= n_bins;
        if (idx < 0) idx += n_bins;
    }
    return idx;
}

// one warp per point: lanes split the views, bins are OR-ed into a 128-bit mask, the row is written coalesced
__global__ void __launch_bounds__(256) view_state_kernel(const ViewStateParams p)
{
    extern __shared__ float sviews[];
    for (int i = threadIdx.x; i < p.V * 3; i += blockDim.x) sviews[i] = p.views[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int n_bins = p.n_elev * p.n_azim;
    for (long long pt = blockIdx.x * 8ll + (threadIdx.x >> 5); pt < p.n_pts; pt += gridDim.x * 8ll) {
        const float *x = p.pts + pt * p.pts_dim;
        const float px = x[0], py = x[1], pz = x[2];
        unsigned m[4] = {0u, 0u, 0u, 0u};
        for (int v = lane; v < p.V; v += 32) {
            const int b = view_bin(p, sviews[3 * v] - px, sviews[3 * v + 1] - py, sviews[3 * v + 2] - pz);
            m[b >> 5] |= 1u << (b & 31);
        }
#pragma unroll
        for (int w = 0; w < 4; ++w)
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) m[w] |= __shfl_xor_sync(0xffffffffu, m[w], d);
        float *dst = p.state + pt * n_bins;
        for (int j = lane; j < n_bins; j += 32) dst[j] = ((m[j >> 5] >> (j & 31)) & 1u) ? 1.0f : 0.0f;
    }
}

// out[pt, k] = sum_j state[pt, j] * base[k, j] * sin(polar_j) * polar_step * azim_step
// block: 32 points; W (n_bins x 64) rebuilt in shared memory per block (cheap: 6272 entries).  The state is a {0,1}
// histogram with at most V non-zero bins: bins whose state is exactly 0 contribute an exact 0 and are skipped (the
// branch is uniform: the 64 threads of a group share the point), the others keep the reference's multiply order.
__global__ void __launch_bounds__(256) view_harmonics_kernel(const float *__restrict__ state, const float *__restrict__ base,
                                                             const float *__restrict__ h_polar, float *__restrict__ out,
                                                             long long n_pts, int n_bins, float polar_step, float azim_step)
{
    extern __shared__ float sm[];
    float *W = sm;                       // [n_bins][64]
    float *st = sm + n_bins * 64;        // [32][n_bins + 1]
    for (int i = threadIdx.x; i < n_bins * 64; i += blockDim.x) {
        const int j = i >> 6, k = i & 63;
        W[i] = base[k * n_bins + j];
    }
    float *sinp = st + 32 * (n_bins + 1);  // [n_bins]
    for (int j = threadIdx.x; j < n_bins; j += blockDim.x) sinp[j] = sinf(h_polar[j]);
    const int k = threadIdx.x & 63, g = threadIdx.x >> 6;  // 4 point groups x 64 coefficients
    const int lane = threadIdx.x & 31;
    for (long long p0 = blockIdx.x * 32ll; p0 < n_pts; p0 += gridDim.x * 32ll) {
        __syncthreads();
        const int n = static_cast<int>(n_pts - p0 < 32 ? n_pts - p0 : 32);
        for (int i = threadIdx.x; i < n * n_bins; i += blockDim.x) st[(i / n_bins) * (n_bins + 1) + i % n_bins] = state[p0 * n_bins + i];
        __syncthreads();
        for (int q = g; q < n; q += 4) {   // q is uniform across each warp (a group = 2 warps)
            const float *s = st + q * (n_bins + 1);
            float acc = 0.f;
            for (int j0 = 0; j0 < n_bins; j0 += 32) {
                unsigned m = __ballot_sync(0xffffffffu, j0 + lane < n_bins && s[j0 + lane] != 0.f);
                while (m) {   // ascending bins; products rounded one by one like the reference's element-wise torch ops
                    const int j = j0 + __ffs(m) - 1;
                    m &= m - 1;
                    const float term = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(s[j], W[j * 64 + k]), sinp[j]), polar_step), azim_step);
                    acc = __fadd_rn(acc, term);
                }
            }
            out[(p0 + q) * 64 + k] = acc;
        }
    }
}

// Fused view state -> view harmonics (compute_view_state followed by compute_view_harmonics, scone_utils.py:799-860 and
// :934-960) without the (B, P, n_bins) histogram in HBM: per point, the bins of the V visited cameras are OR-ed into a
// 128-bit mask in shared memory (one thread per (point, view) pair), then every set bin j adds the table row
// T[j][k] = ((1 * base[k][j]) * sin(polar_j)) * polar_step * azim_step  -- bit for bit the term the two-kernel path adds for a
// state of 1.0, in the same ascending-bin order, so both paths give identical results.
// Traffic: 4*pts_dim B read + 256 B written per point (vs + 2 * 4 * n_bins B for the materialised histogram).
// Warp-autonomous mapping (no block barriers, no shared atomics, balance at the granularity of 32 points):
//   phase 1  lane = point: the lane evaluates the reference's binning rule for its own point against the V views in turn
//            (same view for all lanes: the long dependent chain never diverges) and ORs the bins into a 128-bit mask held in
//            4 registers;
//   phase 2  two points at a time, one per half-warp: the mask of the point is fetched with 4 shuffles, lane l adds the table
//            entries T[j][4l..4l+3] of every set bin j (one LDS.128 per bin) and the half-warp writes the 256-byte output row.
//            (ncu: with one point per warp and 2 coefficients per lane the bit decode -- 10 instructions per set bin,
//            replicated by all 32 lanes -- was 62 % of the kernel's 52 M warp instructions.)
// The binning is ~330 issued instructions per ray; 8 CTAs x 8 warps per SM (32 registers) keep the issue slots busy.
__global__ void __launch_bounds__(256, 8) viewstate_harm_kernel(const ViewStateParams p, const float *__restrict__ base,
                                                                const float *__restrict__ h_polar, float *__restrict__ out,
                                                                float polar_step, float azim_step)
{
    extern __shared__ float sm[];
    const int n_bins = p.n_elev * p.n_azim;
    // table rows padded to kTS = 68 floats: in the transposed build below lane j writes T[j][k], and with a stride of 64 all
    // 32 lanes hit one bank (a 32-way conflict on every one of the 196 stores of each CTA, 8 CTAs per SM); 68 makes it 4-way
    // and keeps the rows 16-byte aligned for the LDS.128 of phase 2
    constexpr int kTS = 68;
    float *T = sm;                         // [n_bins][kTS]
    float *sinp = T + n_bins * kTS;        // [n_bins]
    float *sviews = sinp + kMaxBins;       // [V][3]
    for (int j = threadIdx.x; j < n_bins; j += blockDim.x) sinp[j] = sinf(h_polar[j]);
    for (int i = threadIdx.x; i < p.V * 3; i += blockDim.x) sviews[i] = p.views[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int k = threadIdx.x >> 5; k < 64; k += 8)              // row k of base (64, n_bins): coalesced read, transposed store
        for (int j = lane; j < n_bins; j += 32)
            T[j * kTS + k] = __fmul_rn(__fmul_rn(__fmul_rn(base[k * n_bins + j], sinp[j]), polar_step), azim_step);
    __syncthreads();
    const long long n_tiles = (p.n_pts + 31) / 32;
    const long long warp0 = blockIdx.x * 8ll + (threadIdx.x >> 5), n_warps = gridDim.x * 8ll;
    const int half = lane >> 4;                                  // phase 2: each half-warp sums one point, 4 coefficients per lane
    const float4 *T4 = reinterpret_cast<const float4 *>(T) + (lane & 15);
    for (long long tile = warp0; tile < n_tiles; tile += n_warps) {
        const long long pt = tile * 32 + lane;
        unsigned m0 = 0u, m1 = 0u, m2 = 0u, m3 = 0u;
        if (pt < p.n_pts) {
            const float *x = p.pts + pt * p.pts_dim;
            const float px = x[0], py = x[1], pz = x[2];
            for (int v = 0; v < p.V; ++v) {
                const int b = view_bin(p, sviews[3 * v] - px, sviews[3 * v + 1] - py, sviews[3 * v + 2] - pz);
                const unsigned bit = 1u << (b & 31);
                const int w = b >> 5;
                m0 |= w == 0 ? bit : 0u;
                m1 |= w == 1 ? bit : 0u;
                m2 |= w == 2 ? bit : 0u;
                m3 |= w == 3 ? bit : 0u;
            }
        }
        const int n = static_cast<int>(p.n_pts - tile * 32 < 32 ? p.n_pts - tile * 32 : 32);
        float4 *o4 = reinterpret_cast<float4 *>(out + tile * 32 * 64) + (lane & 15);
        for (int q0 = 0; q0 < n; q0 += 2) {
            const int q = q0 + half;                             // absent points carry an empty mask
            unsigned mq[4] = {__shfl_sync(0xffffffffu, m0, q & 31), __shfl_sync(0xffffffffu, m1, q & 31),
                              __shfl_sync(0xffffffffu, m2, q & 31), __shfl_sync(0xffffffffu, m3, q & 31)};
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                unsigned m = mq[w];
                while (m) {   // ascending bins; the two half-warps run their own bit lists under predication
                    const int j = w * 32 + __ffs(m) - 1;
                    m &= m - 1;
                    const float4 t = T4[j * (kTS / 4)];
                    acc.x = __fadd_rn(acc.x, t.x), acc.y = __fadd_rn(acc.y, t.y);
                    acc.z = __fadd_rn(acc.z, t.z), acc.w = __fadd_rn(acc.w, t.w);
                }
            }
            if (q < n) o4[q * 16] = acc;
        }
    }
}

__global__ void __launch_bounds__(256) gather_bins_kernel(const float *__restrict__ in, const int *__restrict__ index,
                                                          float *__restrict__ out, long long total, int n_bins)
{
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
        const long long pt = i / n_bins;
        const int j = static_cast<int>(i - pt * n_bins);
        out[i] = in[pt * n_bins + index[j]];
    }
}

}  // namespace

}  // namespace mac

using namespace mac;

extern "C" int mac_view_state_f32(const float *pts, int pts_dim, const float *views, float *state, int B, int P, int V,
                                  int n_elev, int n_azim, void *stream)
{
    MAC_REQUIRE(pts && views && state, "null tensor pointer");
    MAC_REQUIRE(B > 0 && P > 0 && V >= 0 && pts_dim >= 3, "bad shape B=%d P=%d V=%d pts_dim=%d", B, P, V, pts_dim);
    MAC_REQUIRE(n_elev > 0 && n_azim > 0 && n_elev * n_azim <= kMaxBins, "view state supports at most %d bins", kMaxBins);
    ViewStateParams p{};
    p.pts = pts, p.views = views, p.state = state;
    p.n_pts = static_cast<long long>(B) * P;
    p.pts_dim = pts_dim, p.V = V, p.n_elev = n_elev, p.n_azim = n_azim;
    const double es = M_PI / (n_elev + 1), as = 2.0 * M_PI / n_azim;
    p.elev_step = static_cast<float>(es), p.azim_step = static_cast<float>(as);
    p.half_elev_step = static_cast<float>(es / 2.0), p.half_azim_step = static_cast<float>(as / 2.0);
    p.inv_elev_step = 1.0f / p.elev_step, p.inv_azim_step = 1.0f / p.azim_step;
    p.elev_lo = -((n_elev + 1) / 2);    // Python floor division -n_elev // 2
    p.azim_hi = n_azim / 2;
    p.azim_wrap = -((n_azim + 1) / 2);  // Python -n_azim // 2
    p.elev_shift = n_elev / 2;
    const long long want = (p.n_pts + 7) / 8;
    const int grid = static_cast<int>(want < 148 * 8 ? want : 148 * 8);
    view_state_kernel<<<grid, 256, static_cast<size_t>(V) * 3 * sizeof(float), static_cast<cudaStream_t>(stream)>>>(p);
    MAC_CUDA(cudaGetLastError());
    count_launch();
    return MAC_OK;
}

extern "C" int mac_view_harmonics_f32(const float *state, const float *base, const float *h_polar, float *out, int B, int P,
                                      int n_elev, int n_azim, void *stream)
{
    MAC_REQUIRE(state && base && h_polar && out, "null tensor pointer");
    MAC_REQUIRE(B > 0 && P > 0, "B and P must be positive");
    const int n_bins = n_elev * n_azim;
    MAC_REQUIRE(n_bins > 0 && n_bins <= kMaxBins, "view state supports at most %d bins", kMaxBins);
    const long long n_pts = static_cast<long long>(B) * P;
    const size_t smem = (static_cast<size_t>(n_bins) * 64 + 32 * (n_bins + 1) + n_bins) * sizeof(float);
    static DeviceOnce once;
    if (int rc = ensure_dynamic_smem(once, view_harmonics_kernel, 64 * 1024)) return rc;
    const long long want = (n_pts + 31) / 32;
    const int grid = static_cast<int>(want < 148 * 4 ? want : 148 * 4);
    view_harmonics_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
        state, base, h_polar, out, n_pts, n_bins, static_cast<float>(M_PI / (n_elev + 1)), static_cast<float>(2.0 * M_PI / n_azim));
    MAC_CUDA(cudaGetLastError());
    count_launch();
    return MAC_OK;
}

extern "C" int mac_gather_bins_f32(const float *in, const int *index, float *out, int B, int P, int n_bins, void *stream)
{
    MAC_REQUIRE(in && index && out && B > 0 && P > 0 && n_bins > 0, "bad arguments");
    const long long total = static_cast<long long>(B) * P * n_bins;
    const long long want = (total + 255) / 256;
    const int grid = static_cast<int>(want < 148 * 16 ? want : 148 * 16);
    gather_bins_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, index, out, total, n_bins);
    MAC_CUDA(cudaGetLastError());
    count_launch();
    return MAC_OK;
}

extern "C" int mac_viewstate_harm_f32(const float *pts, int pts_dim, const float *views, const float *base,
                                      const float *h_polar, float *out, int B, int P, int V, int n_elev, int n_azim,
                                      void *stream)
{
    MAC_REQUIRE(pts && views && base && h_polar && out, "null tensor pointer");
    MAC_REQUIRE(B > 0 && P > 0 && V >= 0 && pts_dim >= 3, "bad shape B=%d P=%d V=%d pts_dim=%d", B, P, V, pts_dim);
    MAC_REQUIRE(n_elev > 0 && n_azim > 0 && n_elev * n_azim <= kMaxBins, "view state supports at most %d bins", kMaxBins);
    MAC_REQUIRE(V <= 4096, "at most 4096 visited cameras per call (got %d)", V);
    ViewStateParams p{};
    p.pts = pts, p.views = views, p.state = nullptr;
    p.n_pts = static_cast<long long>(B) * P;
    p.pts_dim = pts_dim, p.V = V, p.n_elev = n_elev, p.n_azim = n_azim;
    const double es = M_PI / (n_elev + 1), as = 2.0 * M_PI / n_azim;
    p.elev_step = static_cast<float>(es), p.azim_step = static_cast<float>(as);
    p.half_elev_step = static_cast<float>(es / 2.0), p.half_azim_step = static_cast<float>(as / 2.0);
    p.inv_elev_step = 1.0f / p.elev_step, p.inv_azim_step = 1.0f / p.azim_step;
    p.elev_lo = -((n_elev + 1) / 2);
    p.azim_hi = n_azim / 2;
    p.azim_wrap = -((n_azim + 1) / 2);
    p.elev_shift = n_elev / 2;
    const int n_bins = n_elev * n_azim;
    const size_t smem = (68 * static_cast<size_t>(n_bins) + kMaxBins + 3 * static_cast<size_t>(V)) * sizeof(float);   // kTS = 68
    static DeviceOnce once;
    if (int rc = ensure_dynamic_smem(once, viewstate_harm_kernel, 96 * 1024)) return rc;
    const long long want = ((p.n_pts + 31) / 32 + 7) / 8;   // one 32-point tile per warp, 8 warps per CTA
    int device = 0;
    MAC_CUDA(cudaGetDevice(&device));
    const long long resident = static_cast<long long>(sm_count(device)) * 8;   // 8 CTAs / SM (~27 KB of shared memory each)
    const int grid = static_cast<int>(want < resident ? want : resident);
    viewstate_harm_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
        p, base, h_polar, out, static_cast<float>(M_PI / (n_elev + 1)), static_cast<float>(2.0 * M_PI / n_azim));
    MAC_CUDA(cudaGetLastError());
    count_launch();
    return MAC_OK;
}
