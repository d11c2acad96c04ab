// Internal launchers of the SconeOcc / SconeVis building blocks (pointnet.cu, linear.cu).
#pragma once
#include <cuda_runtime.h>

#include "mac_common.h"

namespace mac {

// Implicit-GEMM A operand of a convolution (linear.cu, depth.cu): row m of the (M, K) matrix is output pixel m of an
// NHWC convolution, K = k*k*(Ca + Cb) ordered (ky, kx, channel).  Source A (n, Ha, Wa, Ca) is nearest-neighbour
// up-sampled to the (H, W) input grid when its size differs; source B (n, H, W, Cb), if any, is concatenated behind A's
// channels.  Zero or reflect padding.
struct ConvGather {
    const float *a, *b;
    int lda, ldb, Ca, Cb, Ha, Wa;
    int H, W, Ho, Wo, k, stride, pad, reflect;
    float scale_h, scale_w;   // Ha / H, Wa / W (torch nearest: src = min(floor(dst * scale), size - 1))
};
// true if linear_forward_conv can gather this convolution (every 16-byte piece of a k-chunk lies in one tap of one source)
bool conv_gather_supported(const ConvGather &g);
// `ws` (optional, 16-byte aligned, ws_floats floats): scratch for the split-K partial tiles of layers with few output tiles
int linear_forward_conv(const ConvGather &g, const float *W_hi, const float *W_lo, int ldw, const float *bias, float *out,
                        int ldo, int M, int N, int K, int act, const float *res, int ldr, cudaStream_t stream, int res_first,
                        float *ws = nullptr, size_t ws_floats = 0);

int knn16(const float *x, const float *pc, int *idx, float *dist, int B, int Q, int N, cudaStream_t stream);
// ragged batch of cells: queries [q_off[c], q_off[c+1]) against points [n_off[c], n_off[c+1]) of a concatenated cloud;
// writes GLOBAL point rows (offsets on the device, n_cells + 1 ints each; max_q = largest query count of a cell)
int knn16_cells(const float *x, const float *pc, const int *q_off, const int *n_off, int *idx, int n_cells, int max_q,
                cudaStream_t stream);
int gather_rows(const float *table, const int *row_of, float *out, long long rows, int N, cudaStream_t stream);
int embed_first(const float *in, int ld_in, int in_dim, const float *pc, const float *x, const int *idx, int Q, int N,
                const float *w, const float *b, int inner, int append, float *out, int ldo, long long T,
                cudaStream_t stream);
int attn16(const float *qkv, int ldq, float *out, int ldo, long long n_seq, int dqk, int dv, cudaStream_t stream);
// `lens` (optional, device, B ints): ragged batches -- cloud b holds lens[b] <= S tokens in rows [b*S, b*S + lens[b])
int attn_dense(const float *qkv, int ldq, float *out, int ldo, int B, int S, int dqk, int dv, cudaStream_t stream,
               const int *lens = nullptr);
// attention_tc.cu: the same contract on tcgen05 tensor cores; `scratch` holds attn_dense_tc_scratch_floats() floats
size_t attn_dense_tc_scratch_floats(int B, int S, int dv);
int attn_dense_tc(const float *qkv, int ldq, float *out, int ldo, int B, int S, int dqk, int dv, float *scratch,
                  cudaStream_t stream, const int *lens = nullptr);
int colpool(const float *in, int ld, int B, int S, int N, float *out_max, float *out_mean, int ldo, cudaStream_t stream,
            const int *lens = nullptr);
int vis_embed_finish(float *x0, int ld, const float *gmax, int ldg, const float *pts, int ldp, int F, int in_dim, int S,
                     long long T, const float *g, const float *b, float eps, float *ln, int ldl, cudaStream_t stream);
int bias_gemv(const float *W, int ldw, const float *bias, const float *g, int ldg, int N, int K, float *out, int B,
              cudaStream_t stream);

}  // namespace mac
