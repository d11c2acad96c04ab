// Depth-side helpers on either side of the depth network (reference
// utility/macarons_utils.py, methods of `Camera`):
//   unproject_depth   project_depth_in_3D :2339-2360 -- every pixel's NDC coordinates (tables of Camera.__init__
//                     :1929-1938) + metric depth -> world point, through the inverse full projection of the camera
//                     (pytorch3d FoVPerspectiveCameras.unproject_points, scaled_depth_input = False)
//   signed_distance   get_signed_distance_to_depth_maps :2451-2500 -- view-space z of a 3-D point minus the depth map
//                     sampled bilinearly (border padding, align_corners = False) at its projection; masked pixels count
//                     as `fill` (= 1.1 zfar)
// Both are one coalesced pass over their output (HBM-bound, a few hundred KB to a few MB): the reference needs a
// (B, HW, 3) concatenation + a batched 4x4 transform, resp. two transforms, two transposes and grid_sample.
#include "mac_common.h"

namespace mac {
namespace {

constexpr int kCamUnproject = 18;   // inverse full projection 4x4 (row-vector convention) | f1 | f2
constexpr int kCamDistance = 32;    // full projection 4x4 | world-to-view 4x4

__global__ void __launch_bounds__(256) unproject_depth_kernel(const float *__restrict__ depth, const float *__restrict__ cam,
                                                              float *__restrict__ out, int H, int W, float ndc_x0, float ndc_y0,
                                                              float inv_m1)
{
    const int b = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= H * W) return;
    const float *M = cam + static_cast<size_t>(b) * kCamUnproject;
    const int i = idx / W, j = idx - i * W;
    // ndc_x_tab = W / m - (j / (m - 1)) * 2, ndc_y_tab = H / m - (i / (m - 1)) * 2   (fp32, the reference's operation order)
    const float x = __fsub_rn(ndc_x0, __fmul_rn(__fdiv_rn(static_cast<float>(j), inv_m1), 2.0f));
    const float y = __fsub_rn(ndc_y0, __fmul_rn(__fdiv_rn(static_cast<float>(i), inv_m1), 2.0f));
    const float d = depth[static_cast<size_t>(b) * H * W + idx];
    const float sd = __fdiv_rn(__fadd_rn(__fmul_rn(M[16], d), M[17]), d);     // (f1 d + f2) / d
    float o[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) o[c] = fmaf(x, M[c], fmaf(y, M[4 + c], fmaf(sd, M[8 + c], M[12 + c])));
    float *dst = out + (static_cast<size_t>(b) * H * W + idx) * 3;
    dst[0] = __fdiv_rn(o[0], o[3]);
    dst[1] = __fdiv_rn(o[1], o[3]);
    dst[2] = __fdiv_rn(o[2], o[3]);
}

__global__ void __launch_bounds__(256) signed_distance_kernel(const float *__restrict__ pts, const float *__restrict__ depth,
                                                              const unsigned char *__restrict__ mask,
                                                              const float *__restrict__ cam, float *__restrict__ out, int P,
                                                              int H, int W, float fill, float sx, float sy)
{
    const int n = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const float *Pm = cam + static_cast<size_t>(n) * kCamDistance, *V = Pm + 16;
    const float x = pts[3 * p], y = pts[3 * p + 1], z = pts[3 * p + 2];
    const float vz = fmaf(x, V[2], fmaf(y, V[6], fmaf(z, V[10], V[14])));
    const float vw = fmaf(x, V[3], fmaf(y, V[7], fmaf(z, V[11], V[15])));
    const float px = fmaf(x, Pm[0], fmaf(y, Pm[4], fmaf(z, Pm[8], Pm[12])));
    const float py = fmaf(x, Pm[1], fmaf(y, Pm[5], fmaf(z, Pm[9], Pm[13])));
    const float pw = fmaf(x, Pm[3], fmaf(y, Pm[7], fmaf(z, Pm[11], Pm[15])));
    // grid coordinates of grid_sample: (-m / W) ndc_x, (-m / H) ndc_y; align_corners = False pixel coordinates, clamped
    const float gx = sx * __fdiv_rn(px, pw), gy = sy * __fdiv_rn(py, pw);
    float ix = ((gx + 1.0f) * static_cast<float>(W) - 1.0f) * 0.5f;
    float iy = ((gy + 1.0f) * static_cast<float>(H) - 1.0f) * 0.5f;
    ix = fminf(fmaxf(ix, 0.0f), static_cast<float>(W - 1));
    iy = fminf(fmaxf(iy, 0.0f), static_cast<float>(H - 1));
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
    const float tx = ix - fx, ty = iy - fy;
    const float *dm = depth + static_cast<size_t>(n) * H * W;
    const unsigned char *mm = mask + static_cast<size_t>(n) * H * W;
    auto tap = [&](int yy, int xx) -> float {
        if (xx < 0 || xx >= W || yy < 0 || yy >= H) return 0.0f;      // weight is zero there (clamped coordinates)
        const int o = yy * W + xx;
        return mm[o] ? dm[o] : fill;
    };
    const float v = tap(y0, x0) * (1.0f - tx) * (1.0f - ty) + tap(y0, x0 + 1) * tx * (1.0f - ty) +
                    tap(y0 + 1, x0) * (1.0f - tx) * ty + tap(y0 + 1, x0 + 1) * tx * ty;
    out[static_cast<size_t>(n) * P + p] = __fdiv_rn(vz, vw) - v;
}

}  // namespace
}  // namespace mac

using namespace mac;

extern "C" int mac_unproject_depth_f32(const float *depth, const float *cams, float *out, int B, int H, int W, void *stream)
{
    MAC_REQUIRE(depth && cams && out, "null pointer");
    MAC_REQUIRE(B > 0 && H > 1 && W > 1, "need B > 0 and an image of at least 2 x 2 pixels (got %d, %d x %d)", B, H, W);
    const int m = H < W ? H : W;
    dim3 grid((H * W + 255) / 256, B);
    unproject_depth_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        depth, cams, out, H, W, static_cast<float>(static_cast<double>(W) / m), static_cast<float>(static_cast<double>(H) / m),
        static_cast<float>(m - 1));
    MAC_CUDA(cudaGetLastError());
    count_launch();
    return MAC_OK;
}

extern "C" int mac_signed_distance_f32(const float *pts, const float *depth_maps, const unsigned char *mask, const float *cams,
                                       float *out, int n_depth, int P, int H, int W, float fill, void *stream)
{
    MAC_REQUIRE(pts && depth_maps && mask && cams && out, "null pointer");
    MAC_REQUIRE(n_depth > 0 && P > 0 && H > 0 && W > 0, "bad shape n=%d P=%d image %d x %d", n_depth, P, H, W);
    const int m = H < W ? H : W;
    dim3 grid((P + 255) / 256, n_depth);
    signed_distance_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        pts, depth_maps, mask, cams, out, P, H, W, fill, static_cast<float>(-static_cast<double>(m) / W),
        static_cast<float>(-static_cast<double>(m) / H));
    MAC_CUDA(cudaGetLastError());
    count_launch();
    return MAC_OK;
}

// ------------------------------------------------------------------------------------------------
// Cell.fill resolution filter for ALL cells of a Scene.fill_cells call at once (reference
// utility/macarons_utils.py:2556-2561: `torch.min(torch.cdist(new.double(), stored.double()), -1)[0] > resolution`, once per
// cell): new point i belongs to cell slot[i]; the stored points of slot s are rows [off[s], off[s+1]) of `stored`.
// out[i] = distance (float64) to the nearest stored point of its own cell, +inf if the cell is empty.
// ------------------------------------------------------------------------------------------------
namespace mac {
namespace {
__global__ void __launch_bounds__(128) cell_min_dist_kernel(const float *__restrict__ pts, const int *__restrict__ slot,
                                                            const float *__restrict__ stored, const int *__restrict__ off,
                                                            double *__restrict__ out, int M)
{
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= M) return;
    const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    const int s = slot[i];
    double best = INFINITY;
    for (int j = off[s]; j < off[s + 1]; ++j) {
        const double dx = x - static_cast<double>(__ldg(stored + 3 * j)), dy = y - static_cast<double>(__ldg(stored + 3 * j + 1)),
                     dz = z - static_cast<double>(__ldg(stored + 3 * j + 2));
        best = fmin(best, dx * dx + dy * dy + dz * dz);
    }
    out[i] = sqrt(best);
}
}  // namespace
}  // namespace mac

extern "C" int mac_cell_min_dist_f64(const float *pts, const int *slot, const float *stored, const int *off, double *out, int M,
                                     void *stream)
{
    using namespace mac;
    MAC_REQUIRE(M >= 0, "M must be non-negative");
    if (M == 0) return MAC_OK;
    MAC_REQUIRE(pts && slot && off && out, "null pointer");
    cell_min_dist_kernel<<<(M + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(pts, slot, stored, off, out, M);
    MAC_CUDA(cudaGetLastError());
    count_launch();
    return MAC_OK;
}
