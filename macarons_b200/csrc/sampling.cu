// Occupancy-weighted inverse-CDF sampling of proxy points (row a13; reference
// utility/scone_utils.py:1030-1076).  The reference builds an (n_sample x N) difference matrix (0.8 GB at
// N = 100k) and takes a row-wise argmin; that is a searchsorted of the uniforms in the CDF.  Here:
//   kernel 1 (one CTA): occupancy > min_occ mask -> compacted index list, fp64 sum, probabilities p / sum (fp32),
//                       CDF as a running fp64 sum rounded to fp32 per element (torch's CPU cumsum accumulates
//                       in double as well), all in two passes over the points;
//   kernel 2 (one CTA): binary search of every uniform (first CDF entry >= u, "none" -> 0), bitonic sort of the
//                       picks in shared memory, unique + inverse map (torch.unique semantics: sorted values),
//                       and the gather of the selected rows [xyz, occupancy] and their 64 view harmonics.
// Batched form (MACARONS candidate scoring, reference utility/macarons_utils.py:1603-1628 + Camera.get_points_in_fov
// :2400-2435): one CTA per candidate camera; a point is kept when it projects inside the camera's image, lies in front
// of it and within the sensor range, AND its occupancy exceeds min_occ.  Everything else is the same code.
#include "mac_common.h"

namespace mac {

namespace {

constexpr int kScanThreads = 1024;
constexpr int kMaxSamples = 4096;

struct SampleParams {
    const float *X;      // (N, 3)
    const float *preds;  // (N, 1)
    const float *vh;     // (N, 64)
    const float *u;      // (n_sample)
    int N, n_sample;
    float min_occ;
    int *kept;           // (N) compacted indices of points with occupancy > min_occ
    float *cdf;          // (N)
    int *n_kept;         // [0] = number kept, [1] = number of unique picks
    float *res;          // (n_sample, 4)
    float *res_h;        // (n_sample, 64)
    long long *inverse;  // (n_sample)
    long long *picks;    // (n_sample) picked index into the compacted list, unsorted (optional)
    // batched form: blockIdx.x = candidate camera; u / kept / cdf / n_kept / res / res_h / inverse advance per candidate
    const float *cams;   // (C, kCamFloats) or null: [full projection 4x4 | world-to-view 4x4 | centre 3 | pad], row-vector
    float ndc[4];        // min_x, max_x, min_y, max_y
    float fov_range;     // < 0: no range test
    float *volume;       // (C) sum of the kept occupancies (optional)
};

constexpr int kCamFloats = 36;

struct Selector {   // keep(i): occupancy above the threshold and (batched form) inside the candidate's field of view
    const float *X, *preds;
    float min_occ;
    bool use_cam;
    float P[16], Vz[4], centre[3], ndc[4], range;
    __device__ bool operator()(int i, float &v) const
    {
        v = preds[i];
        if (!(v > min_occ)) return false;
        if (!use_cam) return true;
        const float x = X[3 * i], y = X[3 * i + 1], z = X[3 * i + 2];
        // [x y z 1] @ M, M row-major (pytorch3d Transform3d.transform_points), then the perspective divide
        const float px = fmaf(x, P[0], fmaf(y, P[4], fmaf(z, P[8], P[12])));
        const float py = fmaf(x, P[1], fmaf(y, P[5], fmaf(z, P[9], P[13])));
        const float pw = fmaf(x, P[3], fmaf(y, P[7], fmaf(z, P[11], P[15])));
        const float vz = fmaf(x, Vz[0], fmaf(y, Vz[1], fmaf(z, Vz[2], Vz[3])));
        const float nx = __fdiv_rn(px, pw), ny = __fdiv_rn(py, pw);
        bool in = nx >= ndc[0] && nx <= ndc[1] && ny >= ndc[2] && ny <= ndc[3] && vz > 0.f;
        if (range >= 0.f) {
            const float dx = x - centre[0], dy = y - centre[1], dz = z - centre[2];
            in = in && sqrtf(dx * dx + dy * dy + dz * dz) < range;
        }
        return in;
    }
};

__device__ Selector make_selector(const SampleParams &p, int cand)
{
    Selector s;
    s.X = p.X, s.preds = p.preds, s.min_occ = p.min_occ, s.use_cam = p.cams != nullptr, s.range = p.fov_range;
    if (s.use_cam) {
        const float *c = p.cams + static_cast<size_t>(cand) * kCamFloats;
        for (int i = 0; i < 16; ++i) s.P[i] = c[i];
        for (int i = 0; i < 4; ++i) s.Vz[i] = c[16 + 4 * i + 2], s.ndc[i] = p.ndc[i];
        for (int i = 0; i < 3; ++i) s.centre[i] = c[32 + i];
    }
    return s;
}

__global__ void __launch_bounds__(kScanThreads) sample_scan_kernel(const SampleParams p)
{
    __shared__ int s_cnt[kScanThreads];
    __shared__ double s_sum[kScanThreads];
    const int t = threadIdx.x;
    const int cand = blockIdx.x;
    const Selector keep = make_selector(p, cand);
    int *kept = p.kept + static_cast<size_t>(cand) * p.N;
    float *cdf = p.cdf + static_cast<size_t>(cand) * p.N;
    const int per = (p.N + kScanThreads - 1) / kScanThreads;
    const int lo = min(t * per, p.N), hi = min(lo + per, p.N);
    int cnt = 0;
    double sum = 0.0;
    for (int i = lo; i < hi; ++i) {
        float v;
        if (keep(i, v)) {
            ++cnt;
            sum += static_cast<double>(v);
        }
    }
    s_cnt[t] = cnt;
    s_sum[t] = sum;
    __syncthreads();
    // inclusive Hillis-Steele scan over the 1024 partials
    for (int d = 1; d < kScanThreads; d <<= 1) {
        const int c = t >= d ? s_cnt[t - d] : 0;
        const double s = t >= d ? s_sum[t - d] : 0.0;
        __syncthreads();
        s_cnt[t] += c;
        s_sum[t] += s;
        __syncthreads();
    }
    const float total = static_cast<float>(s_sum[kScanThreads - 1]);  // torch.sum(res_preds) in fp32
    int pos = s_cnt[t] - cnt;
    // exclusive prefix of the probabilities in fp64: sum of (fp32) p_i / total over the kept points before `lo`.
    // Recomputed per segment from the fp32 quotients so that it is the sum the sequential cumsum would see.
    __shared__ double s_psum[kScanThreads];
    double psum = 0.0;
    for (int i = lo; i < hi; ++i) {
        float v;
        if (keep(i, v)) psum += static_cast<double>(v / total);
    }
    s_psum[t] = psum;
    __syncthreads();
    for (int d = 1; d < kScanThreads; d <<= 1) {
        const double s = t >= d ? s_psum[t - d] : 0.0;
        __syncthreads();
        s_psum[t] += s;
        __syncthreads();
    }
    double run = s_psum[t] - psum;
    for (int i = lo; i < hi; ++i) {
        float v;
        if (keep(i, v)) {
            run += static_cast<double>(v / total);
            kept[pos] = i;
            cdf[pos] = static_cast<float>(run);
            ++pos;
        }
    }
    if (t == kScanThreads - 1) {
        p.n_kept[2 * cand] = s_cnt[t];
        if (p.volume) p.volume[cand] = total;
    }
}

__global__ void __launch_bounds__(1024) sample_pick_kernel(SampleParams p)
{
    __shared__ int s_pick[kMaxSamples];   // sorted picks
    __shared__ int s_flag[kMaxSamples];   // unique flags -> positions
    __shared__ int s_total;
    const int t = threadIdx.x;
    const int n = p.n_sample;
    {   // batched form: this CTA's candidate
        const size_t cand = blockIdx.x;
        p.kept += cand * p.N, p.cdf += cand * p.N, p.n_kept += 2 * cand, p.u += cand * n;
        p.res += cand * n * 4, p.res_h += cand * n * 64, p.inverse += cand * n;
        if (p.picks) p.picks += cand * n;
    }
    const int n_kept = p.n_kept[0];
    if (n_kept == 0) {  // nothing above the occupancy threshold: empty result (the reference divides by zero here)
        if (t == 0) p.n_kept[1] = 0;
        for (int i = t; i < n; i += blockDim.x) p.inverse[i] = 0;
        return;
    }
    int n_pow2 = 1;
    while (n_pow2 < n) n_pow2 <<= 1;
    // searchsorted(cdf, u, left): first index with cdf >= u; none -> 0 (scone_utils.py:1054-1056)
    for (int i = t; i < n_pow2; i += blockDim.x) {
        int pick = 0x7fffffff;
        if (i < n) {
            const float u = p.u[i];
            int lo = 0, hi = n_kept;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (p.cdf[mid] < u) lo = mid + 1;
                else hi = mid;
            }
            pick = lo >= n_kept ? 0 : lo;
            if (p.picks) p.picks[i] = pick;
        }
        s_pick[i] = pick;
    }
    __syncthreads();
    // bitonic sort (ascending); padding entries are INT_MAX
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = t; i < n_pow2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const int a = s_pick[i], b = s_pick[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        s_pick[i] = b;
                        s_pick[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    // unique: flag the first occurrence, inclusive scan of the flags
    for (int i = t; i < n_pow2; i += blockDim.x) s_flag[i] = (i < n && (i == 0 || s_pick[i] != s_pick[i - 1])) ? 1 : 0;
    __syncthreads();
    for (int d = 1; d < n_pow2; d <<= 1) {
        int add[kMaxSamples / 1024];
        for (int i = t, c = 0; i < n_pow2; i += blockDim.x, ++c) add[c] = i >= d ? s_flag[i - d] : 0;
        __syncthreads();
        for (int i = t, c = 0; i < n_pow2; i += blockDim.x, ++c) s_flag[i] += add[c];
        __syncthreads();
    }
    if (t == 0) {
        s_total = s_flag[n - 1];
        p.n_kept[1] = s_total;
    }
    __syncthreads();
    // compact the unique values to the front of a second list (re-using the tail of s_flag is not possible: it is
    // still read), so write them into global `kept`-independent scratch: the res rows themselves are the output.
    // Row r of the output = unique value number r (ascending).
    for (int i = t; i < n; i += blockDim.x) {
        const bool first = (i == 0) || (s_pick[i] != s_pick[i - 1]);
        if (first) {
            const int r = s_flag[i] - 1;
            const int src = p.kept[s_pick[i]];
            p.res[4 * r + 0] = p.X[3 * src + 0];
            p.res[4 * r + 1] = p.X[3 * src + 1];
            p.res[4 * r + 2] = p.X[3 * src + 2];
            p.res[4 * r + 3] = p.preds[src];
        }
    }
    // view harmonics of the unique rows: 16 threads x float4 per row
    for (int e = t; e < n * 16; e += blockDim.x) {
        const int i = e >> 4, c = e & 15;
        const bool first = (i == 0) || (s_pick[i] != s_pick[i - 1]);
        if (first) {
            const int r = s_flag[i] - 1;
            const int src = p.kept[s_pick[i]];
            reinterpret_cast<float4 *>(p.res_h + 64 * static_cast<size_t>(r))[c] =
                reinterpret_cast<const float4 *>(p.vh + 64 * static_cast<size_t>(src))[c];
        }
    }
    // inverse map: rank of each (unsorted) pick among the unique values = (scan value at its last sorted position) - 1
    for (int i = t; i < n; i += blockDim.x) {
        const float u = p.u[i];
        int lo = 0, hi = n_kept;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (p.cdf[mid] < u) lo = mid + 1;
            else hi = mid;
        }
        const int pick = lo >= n_kept ? 0 : lo;
        int a = 0, b = n;   // first sorted position holding `pick`
        while (a < b) {
            const int mid = (a + b) >> 1;
            if (s_pick[mid] < pick) a = mid + 1;
            else b = mid;
        }
        p.inverse[i] = s_flag[a] - 1;
    }
}


// Field-of-view mask of N points for ONE camera (Camera.get_points_in_fov, reference utility/macarons_utils.py:2400-2435):
// the Selector's geometric predicate without the occupancy test.
__global__ void __launch_bounds__(256) points_in_fov_kernel(const SampleParams p, unsigned char *__restrict__ mask)
{
    __shared__ Selector sel;
    if (threadIdx.x == 0) {
        sel = make_selector(p, 0);
        sel.min_occ = -1.f;
    }
    __syncthreads();
    for (int i = blockIdx.x * 256 + threadIdx.x; i < p.N; i += gridDim.x * 256) {
        const float x = sel.X[3 * i], y = sel.X[3 * i + 1], z = sel.X[3 * i + 2];
        const float px = fmaf(x, sel.P[0], fmaf(y, sel.P[4], fmaf(z, sel.P[8], sel.P[12])));
        const float py = fmaf(x, sel.P[1], fmaf(y, sel.P[5], fmaf(z, sel.P[9], sel.P[13])));
        const float pw = fmaf(x, sel.P[3], fmaf(y, sel.P[7], fmaf(z, sel.P[11], sel.P[15])));
        const float vz = fmaf(x, sel.Vz[0], fmaf(y, sel.Vz[1], fmaf(z, sel.Vz[2], sel.Vz[3])));
        const float nx = __fdiv_rn(px, pw), ny = __fdiv_rn(py, pw);
        bool in = nx >= sel.ndc[0] && nx <= sel.ndc[1] && ny >= sel.ndc[2] && ny <= sel.ndc[3] && vz > 0.f;
        if (sel.range >= 0.f) {
            const float dx = x - sel.centre[0], dy = y - sel.centre[1], dz = z - sel.centre[2];
            in = in && sqrtf(dx * dx + dy * dy + dz * dz) < sel.range;
        }
        mask[i] = in ? 1 : 0;
    }
}

}  // namespace

}  // namespace mac

using namespace mac;

extern "C" size_t mac_sample_proxy_workspace_bytes(int N) { return static_cast<size_t>(N) * 8 + 256; }

extern "C" int mac_sample_proxy_points_f32(const float *X, const float *preds, const float *view_harmonics, const float *u, int N,
                                           int n_sample, float min_occ, float *res, float *res_harmonics, long long *inverse,
                                           int *counts, void *workspace, size_t workspace_bytes, void *stream)
{
    MAC_REQUIRE(X && preds && view_harmonics && u && res && res_harmonics && inverse && counts && workspace, "null pointer");
    MAC_REQUIRE(N > 0 && n_sample > 0 && n_sample <= kMaxSamples, "need N > 0 and 0 < n_sample <= %d", kMaxSamples);
    MAC_REQUIRE((reinterpret_cast<uintptr_t>(view_harmonics) & 15u) == 0 && (reinterpret_cast<uintptr_t>(res_harmonics) & 15u) == 0,
                "harmonics must be 16-byte aligned");
    if (workspace_bytes < mac_sample_proxy_workspace_bytes(N)) {
        set_error("workspace too small: need %zu bytes, got %zu", mac_sample_proxy_workspace_bytes(N), workspace_bytes);
        return MAC_ERR_WORKSPACE;
    }
    SampleParams p{};
    p.X = X, p.preds = preds, p.vh = view_harmonics, p.u = u, p.N = N, p.n_sample = n_sample, p.min_occ = min_occ;
    p.kept = static_cast<int *>(workspace);
    p.cdf = reinterpret_cast<float *>(static_cast<unsigned char *>(workspace) + static_cast<size_t>(N) * 4);
    p.n_kept = counts;
    p.res = res, p.res_h = res_harmonics, p.inverse = inverse, p.picks = nullptr;
    p.cams = nullptr, p.fov_range = -1.f, p.volume = nullptr;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    sample_scan_kernel<<<1, kScanThreads, 0, st>>>(p);
    MAC_CUDA(cudaGetLastError());
    sample_pick_kernel<<<1, 1024, 0, st>>>(p);
    MAC_CUDA(cudaGetLastError());
    count_launch(2);
    return MAC_OK;
}

extern "C" size_t mac_fov_sample_proxy_workspace_bytes(int N, int C) { return static_cast<size_t>(N) * C * 8 + 256; }

extern "C" int mac_fov_sample_proxy_f32(const float *X, const float *preds, const float *view_harmonics, const float *cams,
                                        const float *ndc_bounds, float fov_range, float min_occ, const float *u, int N, int C,
                                        int n_sample, float *res, float *res_harmonics, long long *inverse, int *counts,
                                        float *volume, void *workspace, size_t workspace_bytes, void *stream)
{
    MAC_REQUIRE(X && preds && view_harmonics && cams && ndc_bounds && u && res && res_harmonics && inverse && counts && workspace,
                "null pointer");
    MAC_REQUIRE(N > 0 && C > 0 && n_sample > 0 && n_sample <= kMaxSamples, "need N, C > 0 and 0 < n_sample <= %d", kMaxSamples);
    MAC_REQUIRE((reinterpret_cast<uintptr_t>(view_harmonics) & 15u) == 0 && (reinterpret_cast<uintptr_t>(res_harmonics) & 15u) == 0,
                "harmonics must be 16-byte aligned");
    if (workspace_bytes < mac_fov_sample_proxy_workspace_bytes(N, C)) {
        set_error("workspace too small: need %zu bytes, got %zu", mac_fov_sample_proxy_workspace_bytes(N, C), workspace_bytes);
        return MAC_ERR_WORKSPACE;
    }
    SampleParams p{};
    p.X = X, p.preds = preds, p.vh = view_harmonics, p.u = u, p.N = N, p.n_sample = n_sample, p.min_occ = min_occ;
    p.kept = static_cast<int *>(workspace);
    p.cdf = reinterpret_cast<float *>(static_cast<unsigned char *>(workspace) + static_cast<size_t>(N) * C * 4);
    p.n_kept = counts;
    p.res = res, p.res_h = res_harmonics, p.inverse = inverse, p.picks = nullptr;
    p.cams = cams, p.fov_range = fov_range, p.volume = volume;
    for (int i = 0; i < 4; ++i) p.ndc[i] = ndc_bounds[i];   // host array
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    sample_scan_kernel<<<C, kScanThreads, 0, st>>>(p);
    MAC_CUDA(cudaGetLastError());
    sample_pick_kernel<<<C, 1024, 0, st>>>(p);
    MAC_CUDA(cudaGetLastError());
    count_launch(2);
    return MAC_OK;
}

extern "C" int mac_points_in_fov_f32(const float *X, const float *cam, const float *ndc_bounds, float fov_range, int N,
                                     unsigned char *mask, void *stream)
{
    MAC_REQUIRE(X && cam && ndc_bounds && mask, "null pointer");
    MAC_REQUIRE(N >= 0, "N must be non-negative");
    if (N == 0) return MAC_OK;
    SampleParams p{};
    p.X = X, p.N = N, p.cams = cam, p.fov_range = fov_range;
    for (int i = 0; i < 4; ++i) p.ndc[i] = ndc_bounds[i];
    const int want = (N + 255) / 256;
    points_in_fov_kernel<<<want < 148 * 8 ? want : 148 * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, mask);
    MAC_CUDA(cudaGetLastError());
    count_launch();
    return MAC_OK;
}
