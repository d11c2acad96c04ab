// Forward passes of SconeOcc and SconeVis assembled from the tensor-core linear layer (linear.cu) and the
// CUDA-core kernels of pointnet.cu.  Reference: networks/SconeOcc.py:250-347, networks/SconeVis.py:121-162,
// networks/Attention.py.  Everything is enqueued on one stream; the caller owns weights and workspace.
#include <stdlib.h>

#include "nets.h"
#include "tc_common.h"

namespace mac {

namespace {

constexpr float kLnEps = 1e-5f;  // torch.nn.LayerNorm default
constexpr int kHeads = 4;

size_t align256(size_t v) { return (v + 255) / 256 * 256; }

struct Bump {  // carve the caller's workspace
    unsigned char *base;
    size_t used, cap;
    float *f(size_t n)
    {
        float *p = reinterpret_cast<float *>(base + used);
        used += align256(n * sizeof(float));
        return p;
    }
};

// `lnin`: LayerNorm of the rows of X applied on load (statistics from the producing layer, gamma / beta of the consumer's
// norm); `stats_out`: (mean, rstd) of the output rows for the next layer's on-load LayerNorm.
struct LnIn {
    const float *stats = nullptr, *g = nullptr, *b = nullptr;
};
int lin(const float *X, int ldx, const mac_linear_w_t &w, const float *bias, float *out, int ldo, long long M, int act,
        const float *res, int ldr, float *ln_out, int ldl, const float *g, const float *b, int pool, cudaStream_t st,
        LnIn lnin = LnIn(), float *stats_out = nullptr, int act_in = MAC_LIN_NONE)
{
    return linear_forward(X, ldx, w.hi, w.lo, w.ldw, bias, out, ldo, static_cast<int>(M), w.N, w.K, act, res, ldr, ln_out, ldl,
                          g, b, kLnEps, pool, st, 0, lnin.stats, lnin.g, lnin.b, stats_out, act_in);
}

// Buffers of one encoder stack over T tokens of width D.
struct EncBufs {
    float *x, *x2, *ln, *qkv, *att, *ff;
    float *attn_scratch;  // tensor-core attention operands (dense attention only)
    float *stats;         // (T, 2) row statistics (mean, rstd) of the residual stream: LayerNorm is applied on load
    int ldqkv;
};

// MAC_ATTN_CUDA_CORES=1 selects the CUDA-core dense attention kernel (A/B knob for tools/bench_nets.py and the tests)
bool use_tc_attention(int S)
{
    static const bool force_cuda_cores = [] {
        const char *e = getenv("MAC_ATTN_CUDA_CORES");
        return e && atoi(e) != 0;
    }();
    return !force_cuda_cores && S >= 64;
}
size_t enc_floats_per_token(int D, int qkv_n) { return 3 * static_cast<size_t>(D) + qkv_n + D + 2 * D + 2; }

EncBufs carve_enc(Bump &ws, long long T, int D, int qkv_n, size_t attn_scratch_floats = 0)
{
    EncBufs e;
    e.x = ws.f(T * D), e.x2 = ws.f(T * D), e.ln = ws.f(T * D);
    e.qkv = ws.f(T * qkv_n), e.att = ws.f(T * D), e.ff = ws.f(T * 2 * D);
    e.attn_scratch = attn_scratch_floats ? ws.f(attn_scratch_floats) : nullptr;
    e.stats = ws.f(T * 2);
    e.ldqkv = qkv_n;
    return e;
}

// On entry e.x = residual stream and either e.ln = norm1 of encoder 0 applied to it (first_is_materialised) or e.stats =
// its row statistics.  Inside the stack no normalised copy is written: every layer that ends a residual branch emits the
// (mean, rstd) of its rows and the consumer applies its own LayerNorm (gamma, beta) while splitting X for the tensor
// cores.  On exit e.x = residual stream, e.stats = its statistics: the caller's next layer normalises with (fin_g, fin_b).
// seq16: attention over groups of 16 tokens; otherwise over B clouds of S tokens (lens: ragged batches).
int encoder_stack(const mac_encoder_w_t *enc, int n_enc, EncBufs &e, long long T, int D, int dqk, int dv, bool seq16, int B, int S,
                  cudaStream_t st, const int *lens = nullptr, bool first_is_materialised = false)
{
    for (int i = 0; i < n_enc; ++i) {
        const mac_encoder_w_t &w = enc[i];
        const bool mat = first_is_materialised && i == 0;
        LnIn n1;
        if (!mat) n1.stats = e.stats, n1.g = w.ln1_g, n1.b = w.ln1_b;
        if (int rc = lin(mat ? e.ln : e.x, D, w.qkv, w.qkv.bias, e.qkv, e.ldqkv, T, MAC_LIN_NONE, nullptr, 0, nullptr, 0, nullptr,
                         nullptr, 0, st, n1))
            return rc;
        if (seq16) {
            if (int rc = attn16(e.qkv, e.ldqkv, e.att, D, T / 16, dqk, dv, st)) return rc;
        } else {
            if (use_tc_attention(S) && e.attn_scratch) {
                if (int rc = attn_dense_tc(e.qkv, e.ldqkv, e.att, D, B, S, dqk, dv, e.attn_scratch, st, lens)) return rc;
            } else {
                if (int rc = attn_dense(e.qkv, e.ldqkv, e.att, D, B, S, dqk, dv, st, lens)) return rc;
            }
        }
        // x2 = x + out(att), statistics of x2 for norm2
        if (int rc = lin(e.att, D, w.out, w.out.bias, e.x2, D, T, MAC_LIN_NONE, e.x, D, nullptr, 0, nullptr, nullptr, 0, st, LnIn(),
                         e.stats))
            return rc;
        LnIn n2;
        n2.stats = e.stats, n2.g = w.ln2_g, n2.b = w.ln2_b;
        // The GELU between ff1 and ff2 is applied by ff2 ON LOAD (its split warps have slack; the 2 MUFU + 13 FP instructions per
        // element made ff1's epilogue the slowest stage of the block): e.ff holds the pre-activation.  Same function on the same
        // fp32 values, so the result is unchanged bit for bit.  MAC_FF_GELU_IN_EPILOGUE=1 restores the first form (A/B timing).
        static const int gelu_in_epilogue = [] { const char *v = getenv("MAC_FF_GELU_IN_EPILOGUE"); return v ? atoi(v) : 0; }();
        if (int rc = lin(e.x2, D, w.ff1, w.ff1.bias, e.ff, 2 * D, T, gelu_in_epilogue ? MAC_LIN_GELU : MAC_LIN_NONE, nullptr, 0, nullptr,
                         0, nullptr, nullptr, 0, st, n2))
            return rc;
        // x = x2 + ff2(gelu(ff)), statistics of x for the next norm1 (or the final norm)
        if (int rc = lin(e.ff, 2 * D, w.ff2, w.ff2.bias, e.x, D, T, MAC_LIN_NONE, e.x2, D, nullptr, 0, nullptr, nullptr, 0, st, LnIn(),
                         e.stats, gelu_in_epilogue ? MAC_LIN_NONE : MAC_LIN_GELU))
            return rc;
    }
    return MAC_OK;
}

int check_pct(const mac_pct_w_t &w)
{
    MAC_REQUIRE(w.n_enc >= 1 && w.n_enc <= MAC_MAX_ENCODERS, "bad encoder count %d", w.n_enc);
    MAC_REQUIRE(w.d_model == 128 && w.dqk == 8 && w.dv == 32, "PCTransformer kernels are built for d_model 128, 4 heads of (8, 32)");
    MAC_REQUIRE(w.in_dim == 3 && w.inner + w.in_dim == w.d_model && w.emb2.K == w.d_model && w.emb2.N == w.d_model,
                "PCTransformer embedding must be 3 -> %d -> %d (+3 input)", w.inner, w.inner);
    return MAC_OK;
}

}  // namespace

}  // namespace mac

using namespace mac;

// ------------------------------------------------------------------------------------------------
// SconeVis
// ------------------------------------------------------------------------------------------------
extern "C" size_t mac_sconevis_workspace_bytes(int B, int S)
{
    const size_t T = static_cast<size_t>(B) * S;
    size_t n = 0;
    n += align256(T * 128 * 4);                       // h
    n += align256(static_cast<size_t>(B) * 128 * 4);  // gmax
    n += 6 * align256(T * 512 * 4) + align256(T * 2 * 4);   // encoder buffers (upper bound: D = 256, qkv 384, ff 512), row stats
    n += align256(T * 256 * 4) + align256(T * 128 * 4);
    n += align256(attn_dense_tc_scratch_floats(B, S, 64) * 4);
    return n + 4096;
}

namespace {
int sconevis_forward(const mac_sconevis_w_t *w, const float *pts, const float *vh, float *out, int B, int S, const int *lens,
                     void *workspace, size_t workspace_bytes, void *stream)
{
    MAC_REQUIRE(w && pts && vh && out && workspace, "null pointer");
    MAC_REQUIRE(B > 0 && S > 0, "B and S must be positive (got %d, %d)", B, S);
    MAC_REQUIRE(w->d_model == 256 && w->dqk == 16 && w->dv == 64 && w->in_dim == 4 && 2 * w->inner + w->in_dim == 256,
                "SconeVis kernels are built for pts_dim 4, d_model 256, 4 heads of (16, 64)");
    MAC_REQUIRE(w->n_enc >= 1 && w->n_enc <= MAC_MAX_ENCODERS, "bad encoder count %d", w->n_enc);
    MAC_REQUIRE(w->fc1.N == 192 && w->fc2.K == 256 && w->fc3.N == MAC_N_HARMONICS, "unexpected SconeVis head shape");
    if (workspace_bytes < mac_sconevis_workspace_bytes(B, S)) {
        set_error("workspace too small: need %zu bytes, got %zu", mac_sconevis_workspace_bytes(B, S), workspace_bytes);
        return MAC_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long T = static_cast<long long>(B) * S;
    const int D = 256, F = w->inner;
    Bump ws{static_cast<unsigned char *>(workspace), 0, workspace_bytes};
    float *h = ws.f(T * 128);
    float *gmax = ws.f(static_cast<size_t>(B) * 128);
    EncBufs e = carve_enc(ws, T, D, w->enc[0].qkv.N, attn_dense_tc_scratch_floats(B, S, 64));
    float *hb = ws.f(T * 256);
    float *h2 = ws.f(T * 128);

    // embedding: h = GELU(W1 p + b1); e = W2 h + b2 -> x[:, :F]; global max; concat + norm1
    if (int rc = embed_first(pts, 4, 4, nullptr, nullptr, nullptr, 0, 0, w->emb1_w, w->emb1_b, F, 0, h, 128, T, st)) return rc;
    if (int rc = lin(h, 128, w->emb2, w->emb2.bias, e.x, D, T, MAC_LIN_NONE, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, st)) return rc;
    if (int rc = colpool(e.x, D, B, S, F, gmax, nullptr, 128, st, lens)) return rc;
    if (int rc = vis_embed_finish(e.x, D, gmax, 128, pts, 4, F, 4, S, T, w->enc[0].ln1_g, w->enc[0].ln1_b, kLnEps, e.ln, D, st))
        return rc;
    if (int rc = encoder_stack(w->enc, w->n_enc, e, T, D, w->dqk, w->dv, false, B, S, st, lens, true)) return rc;
    // head: final norm (on load) -> fc1 + GELU | view harmonics -> fc2 + GELU -> fc3
    LnIn fin;
    fin.stats = e.stats, fin.g = w->ln_g, fin.b = w->ln_b;
    if (int rc = lin(e.x, D, w->fc1, w->fc1.bias, hb, 256, T, MAC_LIN_GELU, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, st, fin)) return rc;
    MAC_CUDA(cudaMemcpy2DAsync(hb + 192, 256 * sizeof(float), vh, 64 * sizeof(float), 64 * sizeof(float), T,
                               cudaMemcpyDeviceToDevice, st));
    if (int rc = lin(hb, 256, w->fc2, w->fc2.bias, h2, 128, T, MAC_LIN_GELU, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, st)) return rc;
    if (int rc = lin(h2, 128, w->fc3, w->fc3.bias, out, 64, T, MAC_LIN_NONE, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, st)) return rc;
    return MAC_OK;
}
}  // namespace

extern "C" int mac_sconevis_forward_f32(const mac_sconevis_w_t *w, const float *pts, const float *vh, float *out, int B, int S,
                                        void *workspace, size_t workspace_bytes, void *stream)
{
    return sconevis_forward(w, pts, vh, out, B, S, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int mac_sconevis_forward_ragged_f32(const mac_sconevis_w_t *w, const float *pts, const float *vh, float *out, int B,
                                               int S, const int *lens, void *workspace, size_t workspace_bytes, void *stream)
{
    MAC_REQUIRE(lens, "lens is null");
    return sconevis_forward(w, pts, vh, out, B, S, lens, workspace, workspace_bytes, stream);
}

// ------------------------------------------------------------------------------------------------
// SconeOcc
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int kFeat = 3 * 256 + 512 + 64;  // [3 x local | x embedding | view harmonics] = 1344

struct OccLayout {
    size_t global_T, local_T;
};
}  // namespace

namespace {
size_t sconeocc_workspace(int B, int Sg, int chunk, int Q)
{
    const size_t Tg = static_cast<size_t>(B) * Sg, Tl = static_cast<size_t>(chunk) * 16;
    const size_t Tm = Tg > Tl ? Tg : Tl;
    size_t n = 0;
    n += align256(Tm * 128 * 4);                                  // h
    n += 3 * align256(Tm * 128 * 4) + align256(Tm * 192 * 4) + align256(Tm * 128 * 4) + align256(Tm * 256 * 4) +
         align256(Tm * 2 * 4);  // encoder, row statistics
    n += align256(Tg * 256 * 4);                                  // global linear0 output
    n += align256(attn_dense_tc_scratch_floats(B, Sg, 32) * 4);   // tensor-core attention operands (global transformer)
    n += 2 * align256(static_cast<size_t>(B) * 512 * 4);          // global feature, per-cloud bias
    n += 3 * align256(static_cast<size_t>(Q) * 16 * 4);           // kNN indices of all queries of one cloud, 3 scales
    n += align256(static_cast<size_t>(chunk) * kFeat * 4);        // feature rows
    n += align256(static_cast<size_t>(chunk) * 128 * 4) + align256(static_cast<size_t>(chunk) * 256 * 4);  // x embedding
    n += align256(static_cast<size_t>(chunk) * 512 * 4) + align256(static_cast<size_t>(chunk) * 256 * 4) +
         align256(static_cast<size_t>(chunk) * 4 * 4);            // head
    return n + 4096;
}
}  // namespace

extern "C" size_t mac_sconeocc_workspace_bytes(int B, int Sg, int chunk, int Q) { return sconeocc_workspace(B, Sg, chunk, Q); }

extern "C" int mac_sconeocc_forward_f32(const mac_sconeocc_w_t *w, const float *pc_global, int Sg, const float *const *pc_scale,
                                        const int *n_scale_pts, const float *x, const float *vh, float *out, int B, int Q,
                                        int chunk, void *workspace, size_t workspace_bytes, void *stream)
{
    MAC_REQUIRE(w && pc_global && pc_scale && n_scale_pts && x && vh && out && workspace, "null pointer");
    MAC_REQUIRE(B > 0 && Q > 0 && Sg > 0 && chunk > 0, "B, Q, Sg, chunk must be positive");
    MAC_REQUIRE(w->n_scale == 3, "SconeOcc kernels are built for 3 neighbourhood scales");
    if (int rc = check_pct(w->global_pct)) return rc;
    for (int s = 0; s < w->n_scale; ++s) {
        if (int rc = check_pct(w->local_pct[s])) return rc;
        MAC_REQUIRE(w->local_pct[s].linear0.N == 128, "local feature must be 2 x 128");
        MAC_REQUIRE(pc_scale[s] && n_scale_pts[s] >= 16, "scale %d: kNN needs at least 16 cloud points (got %d)", s, n_scale_pts[s]);
    }
    MAC_REQUIRE(w->global_pct.linear0.N == 256 && w->global_dim == 512 && w->lin1.K == kFeat && w->lin1.N == 512 &&
                    w->xemb1_n == 128 && w->xemb3.N == 512 && w->lin3.N == 1,
                "unexpected SconeOcc head shape");
    if (workspace_bytes < mac_sconeocc_workspace_bytes(B, Sg, chunk, Q)) {
        set_error("workspace too small: need %zu bytes, got %zu", mac_sconeocc_workspace_bytes(B, Sg, chunk, Q), workspace_bytes);
        return MAC_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int D = 128;
    const long long Tg = static_cast<long long>(B) * Sg, Tl = static_cast<long long>(chunk) * 16;
    const long long Tm = Tg > Tl ? Tg : Tl;
    Bump ws{static_cast<unsigned char *>(workspace), 0, workspace_bytes};
    float *h = ws.f(Tm * 128);
    EncBufs e = carve_enc(ws, Tm, D, 192, attn_dense_tc_scratch_floats(B, Sg, 32));
    float *g0 = ws.f(Tg * 256);
    float *gfeat = ws.f(static_cast<size_t>(B) * 512);
    float *bias1 = ws.f(static_cast<size_t>(B) * 512);
    int *idx_all[3];
    for (int s = 0; s < 3; ++s) idx_all[s] = reinterpret_cast<int *>(ws.f(static_cast<size_t>(Q) * 16));
    float *feat = ws.f(static_cast<size_t>(chunk) * kFeat);
    float *xe1 = ws.f(static_cast<size_t>(chunk) * 128);
    float *xe2 = ws.f(static_cast<size_t>(chunk) * 256);
    float *l1 = ws.f(static_cast<size_t>(chunk) * 512);
    float *l2 = ws.f(static_cast<size_t>(chunk) * 256);
    float *l3 = ws.f(static_cast<size_t>(chunk) * 4);

    // ---- global point-cloud feature (SconeOcc.py:269-275, PCTransformer.forward :105-130) ----
    {
        const mac_pct_w_t &g = w->global_pct;
        if (int rc = embed_first(pc_global, 3, 3, nullptr, nullptr, nullptr, 0, 0, g.emb1_w, g.emb1_b, g.inner, 1, h, 128, Tg, st)) return rc;
        if (int rc = lin(h, 128, g.emb2, g.emb2.bias, e.x, D, Tg, MAC_LIN_NONE, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, st, LnIn(),
                         e.stats))
            return rc;
        if (int rc = encoder_stack(g.enc, g.n_enc, e, Tg, D, g.dqk, g.dv, false, B, Sg, st)) return rc;
        LnIn fin;
        fin.stats = e.stats, fin.g = g.ln_g, fin.b = g.ln_b;
        if (int rc = lin(e.x, D, g.linear0, g.linear0.bias, g0, 256, Tg, MAC_LIN_NONE, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, st, fin))
            return rc;
        if (int rc = colpool(g0, 256, B, Sg, 256, gfeat, gfeat + 256, 512, st)) return rc;
        // the global feature is the same for every query of a cloud: fold it into the bias of linear1
        if (int rc = bias_gemv(w->lin1_wg, w->lin1_wg_ld, w->lin1_b, gfeat, 512, 512, 512, bias1, B, st)) return rc;
    }

    // ---- per query: 3 neighbourhood transformers + x embedding + head (SconeOcc.py:293-342) ----
    for (int b = 0; b < B; ++b) {
        // 16 nearest cloud points of EVERY query of this cloud at the three scales, before the chunk loop: one thread per
        // query needs all Q queries in flight to fill the machine (per 16384-query chunk the kernel ran < 1 warp per SM
        // sub-partition: 343 us per call at N = 4096, 13x the throughput bound)
        for (int s = 0; s < w->n_scale; ++s) {
            const int N = n_scale_pts[s];
            if (int rc = knn16(x + static_cast<size_t>(b) * Q * 3, pc_scale[s] + static_cast<size_t>(b) * N * 3, idx_all[s], nullptr,
                               1, Q, N, st))
                return rc;
        }
        for (int q0 = 0; q0 < Q; q0 += chunk) {
            const int nq = Q - q0 < chunk ? Q - q0 : chunk;
            const long long T = static_cast<long long>(nq) * 16;
            const float *xq = x + (static_cast<size_t>(b) * Q + q0) * 3;
            for (int s = 0; s < w->n_scale; ++s) {
                const mac_pct_w_t &l = w->local_pct[s];
                const int N = n_scale_pts[s];
                const float *pcs = pc_scale[s] + static_cast<size_t>(b) * N * 3;
                if (int rc = embed_first(nullptr, 0, 3, pcs, xq, idx_all[s] + static_cast<size_t>(q0) * 16, nq, N, l.emb1_w, l.emb1_b,
                                         l.inner, 1, h, 128, T, st))
                    return rc;
                if (int rc = lin(h, 128, l.emb2, l.emb2.bias, e.x, D, T, MAC_LIN_NONE, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, st,
                                 LnIn(), e.stats))
                    return rc;
                if (int rc = encoder_stack(l.enc, l.n_enc, e, T, D, l.dqk, l.dv, true, 0, 0, st)) return rc;
                LnIn fin;
                fin.stats = e.stats, fin.g = l.ln_g, fin.b = l.ln_b;
                if (int rc = lin(e.x, D, l.linear0, l.linear0.bias, feat + s * 256, kFeat, T, MAC_LIN_NONE, nullptr, 0, nullptr, 0,
                                 nullptr, nullptr, 16, st, fin))
                    return rc;
            }
            if (int rc = embed_first(xq, 3, 3, nullptr, nullptr, nullptr, 0, 0, w->xemb1_w, w->xemb1_b, w->xemb1_n, 0, xe1, 128, nq, st))
                return rc;
            if (int rc = lin(xe1, 128, w->xemb2, w->xemb2.bias, xe2, 256, nq, MAC_LIN_GELU, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, st))
                return rc;
            if (int rc = lin(xe2, 256, w->xemb3, w->xemb3.bias, feat + 768, kFeat, nq, MAC_LIN_GELU, nullptr, 0, nullptr, 0, nullptr,
                             nullptr, 0, st))
                return rc;
            MAC_CUDA(cudaMemcpy2DAsync(feat + 1280, kFeat * sizeof(float), vh + (static_cast<size_t>(b) * Q + q0) * 64,
                                       64 * sizeof(float), 64 * sizeof(float), nq, cudaMemcpyDeviceToDevice, st));
            if (int rc = lin(feat, kFeat, w->lin1, bias1 + static_cast<size_t>(b) * 512, l1, 512, nq, MAC_LIN_GELU, nullptr, 0, nullptr,
                             0, nullptr, nullptr, 0, st))
                return rc;
            if (int rc = lin(l1, 512, w->lin2, w->lin2.bias, l2, 256, nq, MAC_LIN_GELU, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, st))
                return rc;
            if (int rc = lin(l2, 256, w->lin3, w->lin3.bias, l3, 4, nq, MAC_LIN_GELU, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, st))
                return rc;
            MAC_CUDA(cudaMemcpy2DAsync(out + static_cast<size_t>(b) * Q + q0, sizeof(float), l3, 4 * sizeof(float), sizeof(float), nq,
                                       cudaMemcpyDeviceToDevice, st));
        }
    }
    return MAC_OK;
}

// ------------------------------------------------------------------------------------------------
// SconeOcc over a ragged batch of cells: the per-cell `macarons(mode='occupancy', ...)` calls
// of compute_scene_occupancy_probability_field (/root/reference/macarons/utility/macarons_utils.py:1443-1518) as ONE forward.
// Every cell has its own partial cloud (own sub-samples, own global feature) and its own queries; the queries of all cells
// form one flat array, so the neighbourhood transformers, the x embedding and the head run over full 128-row tiles whatever
// the cell sizes are.  Per query the arithmetic is that of mac_sconeocc_forward_f32 for its cell alone.
// ------------------------------------------------------------------------------------------------
extern "C" size_t mac_sconeocc_cells_workspace_bytes(int n_cells, int Sg, int chunk, long long Qtot)
{
    if (n_cells <= 0 || Sg <= 0 || chunk <= 0 || Qtot <= 0 || Qtot > (1ll << 30)) return 0;
    return sconeocc_workspace(n_cells, Sg, chunk, static_cast<int>(Qtot)) + align256(static_cast<size_t>(chunk) * 512 * 4);
}

extern "C" int mac_sconeocc_forward_cells_f32(const mac_sconeocc_w_t *w, int n_cells, const float *pc_global, int Sg,
                                              const int *lens_g, const float *const *pc_scale, const int *const *scale_off,
                                              const float *x, const float *vh, const int *q_off, const int *cell_of_q,
                                              int max_q, float *out, long long Qtot, int chunk, void *workspace,
                                              size_t workspace_bytes, void *stream)
{
    MAC_REQUIRE(w && pc_global && lens_g && pc_scale && scale_off && x && vh && q_off && cell_of_q && out && workspace,
                "null pointer");
    MAC_REQUIRE(n_cells > 0 && Qtot > 0 && Qtot <= (1ll << 30) && Sg > 0 && chunk > 0 && max_q > 0,
                "n_cells, Qtot, Sg, chunk, max_q must be positive");
    MAC_REQUIRE(w->n_scale == 3, "SconeOcc kernels are built for 3 neighbourhood scales");
    if (int rc = check_pct(w->global_pct)) return rc;
    for (int s = 0; s < w->n_scale; ++s) {
        if (int rc = check_pct(w->local_pct[s])) return rc;
        MAC_REQUIRE(w->local_pct[s].linear0.N == 128, "local feature must be 2 x 128");
        MAC_REQUIRE(pc_scale[s] && scale_off[s], "scale %d: null cloud / offsets", s);
    }
    MAC_REQUIRE(w->global_pct.linear0.N == 256 && w->global_dim == 512 && w->lin1.K == kFeat && w->lin1.N == 512 &&
                    w->xemb1_n == 128 && w->xemb3.N == 512 && w->lin3.N == 1,
                "unexpected SconeOcc head shape");
    const size_t need = mac_sconeocc_cells_workspace_bytes(n_cells, Sg, chunk, Qtot);
    if (workspace_bytes < need) {
        set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
        return MAC_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int D = 128, B = n_cells;
    const long long Q = Qtot;
    const long long Tg = static_cast<long long>(B) * Sg, Tl = static_cast<long long>(chunk) * 16;
    const long long Tm = Tg > Tl ? Tg : Tl;
    Bump ws{static_cast<unsigned char *>(workspace), 0, workspace_bytes};
    float *h = ws.f(Tm * 128);
    EncBufs e = carve_enc(ws, Tm, D, 192, attn_dense_tc_scratch_floats(B, Sg, 32));
    float *g0 = ws.f(Tg * 256);
    float *gfeat = ws.f(static_cast<size_t>(B) * 512);
    float *bias1 = ws.f(static_cast<size_t>(B) * 512);
    int *idx_all[3];
    for (int s = 0; s < 3; ++s) idx_all[s] = reinterpret_cast<int *>(ws.f(static_cast<size_t>(Q) * 16));
    float *feat = ws.f(static_cast<size_t>(chunk) * kFeat);
    float *xe1 = ws.f(static_cast<size_t>(chunk) * 128);
    float *xe2 = ws.f(static_cast<size_t>(chunk) * 256);
    float *l1 = ws.f(static_cast<size_t>(chunk) * 512);
    float *l2 = ws.f(static_cast<size_t>(chunk) * 256);
    float *l3 = ws.f(static_cast<size_t>(chunk) * 4);
    float *qbias = ws.f(static_cast<size_t>(chunk) * 512);

    // ---- global feature of every cell: ragged batch of B clouds of lens_g[c] <= Sg points ----
    {
        const mac_pct_w_t &g = w->global_pct;
        if (int rc = embed_first(pc_global, 3, 3, nullptr, nullptr, nullptr, 0, 0, g.emb1_w, g.emb1_b, g.inner, 1, h, 128, Tg, st)) return rc;
        if (int rc = lin(h, 128, g.emb2, g.emb2.bias, e.x, D, Tg, MAC_LIN_NONE, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, st, LnIn(),
                         e.stats))
            return rc;
        if (int rc = encoder_stack(g.enc, g.n_enc, e, Tg, D, g.dqk, g.dv, false, B, Sg, st, lens_g)) return rc;
        LnIn fin;
        fin.stats = e.stats, fin.g = g.ln_g, fin.b = g.ln_b;
        if (int rc = lin(e.x, D, g.linear0, g.linear0.bias, g0, 256, Tg, MAC_LIN_NONE, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, st, fin))
            return rc;
        if (int rc = colpool(g0, 256, B, Sg, 256, gfeat, gfeat + 256, 512, st, lens_g)) return rc;
        if (int rc = bias_gemv(w->lin1_wg, w->lin1_wg_ld, w->lin1_b, gfeat, 512, 512, 512, bias1, B, st)) return rc;
    }

    // ---- 16 nearest points of every query in its own cell's cloud, three scales (global rows of the concatenated clouds) ----
    for (int s = 0; s < w->n_scale; ++s)
        if (int rc = knn16_cells(x, pc_scale[s], q_off, scale_off[s], idx_all[s], B, max_q, st)) return rc;

    // ---- flat query loop: neighbourhood transformers + x embedding + head ----
    for (long long q0 = 0; q0 < Q; q0 += chunk) {
        const int nq = static_cast<int>(Q - q0 < chunk ? Q - q0 : chunk);
        const long long T = static_cast<long long>(nq) * 16;
        const float *xq = x + static_cast<size_t>(q0) * 3;
        for (int s = 0; s < w->n_scale; ++s) {
            const mac_pct_w_t &l = w->local_pct[s];
            // indices are global rows: "one cloud" of unspecified size holding every cell's points
            if (int rc = embed_first(nullptr, 0, 3, pc_scale[s], xq, idx_all[s] + static_cast<size_t>(q0) * 16, nq, 0, l.emb1_w,
                                     l.emb1_b, l.inner, 1, h, 128, T, st))
                return rc;
            if (int rc = lin(h, 128, l.emb2, l.emb2.bias, e.x, D, T, MAC_LIN_NONE, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, st,
                             LnIn(), e.stats))
                return rc;
            if (int rc = encoder_stack(l.enc, l.n_enc, e, T, D, l.dqk, l.dv, true, 0, 0, st)) return rc;
            LnIn fin;
            fin.stats = e.stats, fin.g = l.ln_g, fin.b = l.ln_b;
            if (int rc = lin(e.x, D, l.linear0, l.linear0.bias, feat + s * 256, kFeat, T, MAC_LIN_NONE, nullptr, 0, nullptr, 0,
                             nullptr, nullptr, 16, st, fin))
                return rc;
        }
        if (int rc = embed_first(xq, 3, 3, nullptr, nullptr, nullptr, 0, 0, w->xemb1_w, w->xemb1_b, w->xemb1_n, 0, xe1, 128, nq, st))
            return rc;
        if (int rc = lin(xe1, 128, w->xemb2, w->xemb2.bias, xe2, 256, nq, MAC_LIN_GELU, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, st))
            return rc;
        if (int rc = lin(xe2, 256, w->xemb3, w->xemb3.bias, feat + 768, kFeat, nq, MAC_LIN_GELU, nullptr, 0, nullptr, 0, nullptr,
                         nullptr, 0, st))
            return rc;
        MAC_CUDA(cudaMemcpy2DAsync(feat + 1280, kFeat * sizeof(float), vh + static_cast<size_t>(q0) * 64, 64 * sizeof(float),
                                   64 * sizeof(float), nq, cudaMemcpyDeviceToDevice, st));
        // per-query bias of linear1 = its cell's [b1 + W_g . global feature]: added before the GELU as a "residual"
        if (int rc = gather_rows(bias1, cell_of_q + q0, qbias, nq, 512, st)) return rc;
        if (int rc = linear_forward(feat, kFeat, w->lin1.hi, w->lin1.lo, w->lin1.ldw, nullptr, l1, 512, nq, w->lin1.N, w->lin1.K,
                                    MAC_LIN_GELU, qbias, 512, nullptr, 0, nullptr, nullptr, kLnEps, 0, st, /*res_first=*/1, nullptr,
                                    nullptr, nullptr, nullptr))
            return rc;
        if (int rc = lin(l1, 512, w->lin2, w->lin2.bias, l2, 256, nq, MAC_LIN_GELU, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, st))
            return rc;
        if (int rc = lin(l2, 256, w->lin3, w->lin3.bias, l3, 4, nq, MAC_LIN_GELU, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, st))
            return rc;
        MAC_CUDA(cudaMemcpy2DAsync(out + q0, sizeof(float), l3, 4 * sizeof(float), sizeof(float), nq, cudaMemcpyDeviceToDevice, st));
    }
    return MAC_OK;
}
