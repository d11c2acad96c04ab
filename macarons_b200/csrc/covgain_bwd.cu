// Backward pass of the coverage-gain / visibility-gain integration with respect to the SH coefficients:
//     coverage:    dL/dH[b,p,k] = sum_c  g[b,c] / P   * act'(z_pc) * Y_k(u_pc)
//     visibility:  dL/dH[b,p,k] = sum_c  g[b,c,p]     * act'(z_pc) * Y_k(u_pc),      z_pc = sum_k Y_k(u_pc) H[b,p,k]
// (reference: autograd through SconeVis.compute_coverage_gain / compute_visibilities, networks/SconeVis.py:164-252, as
// used by trainers/pretrain_scone_vis.py:162-225 and trainers/train_macarons.py:423-444; the points and the camera
// positions are data there and receive no gradient).
//
// One surface point per thread: its 64 coefficients (pre-multiplied by the SH normalisation N_lm) and its 64 gradient
// accumulators live in registers; the cameras of the cloud are staged in shared memory and swept once.  Per ray the
// real SH basis is evaluated without trigonometry from the unit ray u = (ux, ct, uz):
//     sin^m(theta) (cos m phi + i sin m phi) = (uz + i ux)^m          (A_m + i B_m, complex recurrence)
//     P_l^m(ct) / sin^m(theta) =: q_l^m,  q_m^m = (-1)^m (2m-1)!!,  q_l^m = ((2l-1) ct q_{l-1}^m - (l+m-1) q_{l-2}^m) / (l-m)
//     Y_{l,+m} = N_lm q_l^m A_m,   Y_{l,-m} = N_lm q_l^m B_m,   Y_{l,0} = N_l0 q_l^0        (k = l*l + l + m)
// which is the closed form of SURVEY.md appendix A.1 (oracle/sh_cov.py::sh_basis_closed_form_f64).  The sum over cameras
// runs in camera order in one thread: the result is deterministic.  HBM traffic: one read of (pts, H), one write of dH.
#include "mac_common.h"

namespace mac {

namespace {

constexpr int kBwdThreads = 128;
constexpr int kBwdCamChunk = 512;  // cameras staged per pass (8 KB of shared memory)

// N_lm for 0 <= m <= l < 8 (reference utility/spherical_harmonics.py:126,138)
__device__ constexpr float kShNorm[8][8] = {
    {2.820947918e-01f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {4.886025119e-01f, 4.886025119e-01f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {6.307831305e-01f, 3.641828102e-01f, 1.820914051e-01f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {7.463526652e-01f, 3.046971996e-01f, 9.635371475e-02f, 3.933623933e-02f, 0.f, 0.f, 0.f, 0.f},
    {8.462843753e-01f, 2.676186174e-01f, 6.307831305e-02f, 1.685838828e-02f, 5.960340338e-03f, 0.f, 0.f, 0.f},
    {9.356025796e-01f, 2.415715473e-01f, 4.565273129e-02f, 9.318824751e-03f, 2.196468058e-03f, 6.945841871e-04f, 0.f, 0.f},
    {1.017107236e+00f, 2.219509952e-01f, 3.509353370e-02f, 5.848922283e-03f, 1.067862224e-03f, 2.276689911e-04f,
     6.572237664e-05f, 0.f},
    {1.092548431e+00f, 2.064722459e-01f, 2.809731381e-02f, 3.973560225e-03f, 5.990367431e-04f, 9.983945719e-05f,
     1.958012848e-05f, 5.233009454e-06f}};
// (-1)^m (2m-1)!!
__device__ constexpr float kQmm[8] = {1.f, -1.f, 3.f, -15.f, 105.f, -945.f, 10395.f, -135135.f};

struct BwdParams {
    const float *pts;
    const float *harm;
    const float *cams;
    const float *gout;  // (B, C) or (B, C, P)
    float *dharm;       // (B, P, 64)
    int pts_dim, B, P, C;
    float inv_count;    // 1 / P for the coverage mean, unused per point
};

// Un-normalised basis of one ray, grouped by order m: q[idx(l, m)] and the azimuthal factors A_m, B_m.
// Y_{l,+-m} = N_lm * q_l^m * {A_m, B_m}.  Everything is indexed at compile time after unrolling.
__device__ __forceinline__ constexpr int qidx(int l, int m) { return m * 8 - (m * (m - 1)) / 2 + (l - m); }  // 36 entries

template <bool SIGMOID, bool PER_POINT>
__global__ void __launch_bounds__(kBwdThreads) covgain_bwd_kernel(const BwdParams prm)
{
    __shared__ float4 scam[kBwdCamChunk];
    const int b = blockIdx.y;
    const int p = blockIdx.x * kBwdThreads + threadIdx.x;
    const bool valid = p < prm.P;
    const size_t row = static_cast<size_t>(b) * prm.P + (valid ? p : 0);

    float hn[64];   // N_lm * H
    float acc[64];  // sum_c w * q * {A, B}   (multiplied by N_lm at the end)
    {
        const float4 *h4 = reinterpret_cast<const float4 *>(prm.harm + row * 64);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float4 v = valid ? __ldg(h4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            hn[4 * i] = v.x, hn[4 * i + 1] = v.y, hn[4 * i + 2] = v.z, hn[4 * i + 3] = v.w;
        }
#pragma unroll
        for (int l = 0; l < 8; ++l)
#pragma unroll
            for (int m = -l; m <= l; ++m) hn[l * l + l + m] *= kShNorm[l][m < 0 ? -m : m];
#pragma unroll
        for (int k = 0; k < 64; ++k) acc[k] = 0.f;
    }
    const float *pp = prm.pts + row * prm.pts_dim;
    const float px = __ldg(pp), py = __ldg(pp + 1), pz = __ldg(pp + 2);

    for (int c0 = 0; c0 < prm.C; c0 += kBwdCamChunk) {
        const int nc = min(kBwdCamChunk, prm.C - c0);
        __syncthreads();
        for (int j = threadIdx.x; j < nc; j += kBwdThreads) {
            const size_t ci = static_cast<size_t>(b) * prm.C + c0 + j;
            const float *src = prm.cams + ci * 3;
            scam[j] = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2),
                                  PER_POINT ? 0.f : __ldg(prm.gout + ci) * prm.inv_count);
        }
        __syncthreads();
        if (!valid) continue;
#pragma unroll 1
        for (int j = 0; j < nc; ++j) {
            const float4 cam = scam[j];
            const float dx = cam.x - px, dy = cam.y - py, dz = cam.z - pz;
            const float rinv = rsqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
            const float ux = dx * rinv, ct = dy * rinv, uz = dz * rinv;
            // azimuthal factors (A_m + i B_m) = (uz + i ux)^m
            float A[8], Bm[8];
            A[0] = 1.f, Bm[0] = 0.f;
#pragma unroll
            for (int m = 1; m < 8; ++m) {
                A[m] = A[m - 1] * uz - Bm[m - 1] * ux;
                Bm[m] = Bm[m - 1] * uz + A[m - 1] * ux;
            }
            // q_l^m(ct)
            float q[36];
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                q[qidx(m, m)] = kQmm[m];
                if (m + 1 < 8) q[qidx(m + 1, m)] = (2 * m + 1) * ct * kQmm[m];
#pragma unroll
                for (int l = m + 2; l < 8; ++l)
                    q[qidx(l, m)] = ((2 * l - 1) * ct * q[qidx(l - 1, m)] - (l + m - 1) * q[qidx(l - 2, m)]) * (1.0f / (l - m));
            }
            // z = sum_k Y_k H_k
            float z = 0.f;
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                z = fmaf(q[qidx(l, 0)], hn[l * l + l], z);
#pragma unroll
                for (int m = 1; m <= l; ++m) {
                    const float t = fmaf(A[m], hn[l * l + l + m], Bm[m] * hn[l * l + l - m]);
                    z = fmaf(q[qidx(l, m)], t, z);
                }
            }
            float g = cam.w;
            if (PER_POINT) g = __ldg(prm.gout + (static_cast<size_t>(b) * prm.C + c0 + j) * prm.P + p);
            float w;
            if (SIGMOID) {
                const float v = 1.0f / (1.0f + __expf(-z));
                w = g * v * (1.0f - v);
            } else {
                w = z > 0.f ? g : 0.f;   // torch: relu'(0) = 0
            }
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                acc[l * l + l] = fmaf(w, q[qidx(l, 0)], acc[l * l + l]);
#pragma unroll
                for (int m = 1; m <= l; ++m) {
                    const float wq = w * q[qidx(l, m)];
                    acc[l * l + l + m] = fmaf(wq, A[m], acc[l * l + l + m]);
                    acc[l * l + l - m] = fmaf(wq, Bm[m], acc[l * l + l - m]);
                }
            }
        }
    }
    if (valid) {
#pragma unroll
        for (int l = 0; l < 8; ++l)
#pragma unroll
            for (int m = -l; m <= l; ++m) acc[l * l + l + m] *= kShNorm[l][m < 0 ? -m : m];
        float4 *d4 = reinterpret_cast<float4 *>(prm.dharm + row * 64);
#pragma unroll
        for (int i = 0; i < 16; ++i) d4[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
    }
}

}  // namespace

}  // namespace mac

using namespace mac;

extern "C" int mac_covgain_backward_f32(const float *pts, int pts_dim, const float *harmonics, const float *cams,
                                        const float *grad_out, float *grad_harmonics, int B, int P, int C, int act,
                                        int per_point, void *stream)
{
    MAC_REQUIRE(pts && harmonics && cams && grad_out && grad_harmonics, "null tensor pointer");
    MAC_REQUIRE(B > 0 && P > 0 && C > 0 && pts_dim >= 3, "bad shape B=%d P=%d C=%d pts_dim=%d", B, P, C, pts_dim);
    MAC_REQUIRE(B <= 65535, "at most 65535 clouds per call (got %d)", B);
    MAC_REQUIRE(act == MAC_ACT_RELU || act == MAC_ACT_SIGMOID, "act must be MAC_ACT_RELU or MAC_ACT_SIGMOID");
    MAC_REQUIRE(((reinterpret_cast<uintptr_t>(harmonics) | reinterpret_cast<uintptr_t>(grad_harmonics)) & 15u) == 0,
                "harmonics and grad_harmonics must be 16-byte aligned");
    BwdParams prm{pts, harmonics, cams, grad_out, grad_harmonics, pts_dim, B, P, C, 1.0f / static_cast<float>(P)};
    const dim3 grid((P + kBwdThreads - 1) / kBwdThreads, B);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (act == MAC_ACT_SIGMOID) {
        if (per_point) covgain_bwd_kernel<true, true><<<grid, kBwdThreads, 0, st>>>(prm);
        else covgain_bwd_kernel<true, false><<<grid, kBwdThreads, 0, st>>>(prm);
    } else {
        if (per_point) covgain_bwd_kernel<false, true><<<grid, kBwdThreads, 0, st>>>(prm);
        else covgain_bwd_kernel<false, false><<<grid, kBwdThreads, 0, st>>>(prm);
    }
    MAC_CUDA(cudaGetLastError());
    count_launch();
    return MAC_OK;
}
