/*
 * macarons_b200 -- C ABI of the Blackwell (sm_100a) next-best-view scoring path.
 *
 * The reference (Anttwo/MACARONS, pure Python) has no FFI layer of its own; its boundary for this
 * path is the Python module API.  Each entry point below is the native
 * replacement for the body of one reference function and is bound with ctypes from the
 * reference-shaped Python classes in macarons_b200/networks and macarons_b200/utility (see
 * INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain C types only; all tensors are caller-owned, dense, row-major fp32;
 *   - `*_f32` entry points take DEVICE pointers and enqueue asynchronously on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - `*_host` entry points take HOST pointers, copy in/out themselves and return after the
 *     result is in host memory;
 *   - return value 0 = success, negative = error (mac_last_error() gives the message of the last
 *     failure on the calling thread);
 *   - no entry point allocates or frees caller memory; workspaces are sized by the matching
 *     `*_workspace_bytes` query and must be zero-filled ONCE by the caller before first use
 *     (every successful call leaves them zero-filled again).
 */
#ifndef MACARONS_B200_H
#define MACARONS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAC_OK 0
#define MAC_ERR_INVALID_ARGUMENT (-1)
#define MAC_ERR_CUDA (-2)
#define MAC_ERR_WORKSPACE (-3)
#define MAC_ERR_UNSUPPORTED (-4)

#define MAC_ACT_RELU 0    /* use_sigmoid=False branch, networks/SconeVis.py:245-246 */
#define MAC_ACT_SIGMOID 1 /* use_sigmoid=True  branch, networks/SconeVis.py:243-244 */

#define MAC_N_HARMONICS 64 /* l < 8; hard-coded 64 at networks/SconeVis.py:241 */

/* Library version: major*10000 + minor*100 + patch. */
int mac_version(void);
/* Message of the last error raised on this thread ("" if none). */
const char *mac_last_error(void);
/* Compute capability the device code was built for (100 for sm_100a). */
int mac_built_for_sm(void);

/* ---------------------------------------------------------------------------------------------
 * Coverage gain  --  replaces the bodies of
 *     SconeVis.compute_coverage_gain      /root/reference/macarons/networks/SconeVis.py:210-252
 *     SconeVis.compute_visibilities       /root/reference/macarons/networks/SconeVis.py:164-208
 *     Macarons.compute_visibility_gains   /root/reference/macarons/networks/Macarons.py:138-178
 * (and with them get_spherical_coords, utility/CustomGeometry.py:27-45, and
 *  get_spherical_harmonics, utility/spherical_harmonics.py:143-156, on the ray tensor).
 *
 *   pts        (B, P, pts_dim) fp32, only [..., :3] is read; pts_dim >= 3
 *   harmonics  (B, P, 64) fp32, coefficient k = l*l + l + m of each point's visibility-gain function
 *   cams       (B, C, 3)  fp32 camera centres
 *   cam_begin, cam_end    this call scores cameras [cam_begin, cam_end) of every cloud
 *                         (the multi-GPU partition: rank r passes its own slice)
 *   act        MAC_ACT_SIGMOID / MAC_ACT_RELU
 *
 * mac_covgain_f32:      out (B, C) fp32; only columns [cam_begin, cam_end) are written:
 *                       out[b,c] = (1/P) sum_p act( sum_k Y_k(cam_c - pt_p) * harmonics[b,p,k] )
 * mac_visibility_f32:   out (B, C, P) fp32; only rows [cam_begin, cam_end) are written (no mean).
 *
 * Workspace (mac_covgain_f32 only): mac_covgain_workspace_bytes(B, C) bytes of device memory,
 * zero-filled once by the caller.  Results are bitwise independent of the launch configuration,
 * of scheduling and of how the camera axis is partitioned across calls / GPUs.
 * ------------------------------------------------------------------------------------------- */
size_t mac_covgain_workspace_bytes(int B, int C);

int mac_covgain_f32(const float *pts, int pts_dim, const float *harmonics, const float *cams,
                    float *out, int B, int P, int C, int cam_begin, int cam_end, int act,
                    void *workspace, size_t workspace_bytes, void *stream);

int mac_visibility_f32(const float *pts, int pts_dim, const float *harmonics, const float *cams,
                       float *out, int B, int P, int C, int cam_begin, int cam_end, int act,
                       void *stream);

/* Backward of mac_covgain_f32 (per_point = 0: grad_out (B, C)) and of mac_visibility_f32 (per_point = 1: grad_out
 * (B, C, P)) with respect to the harmonics: grad_harmonics (B, P, 64) = sum over ALL C cameras of
 * grad * act'(z) * Y_k(ray).  This is what autograd computes through SconeVis.compute_coverage_gain /
 * compute_visibilities (/root/reference/macarons/networks/SconeVis.py:164-252) in the reference's training loops
 * (/root/reference/macarons/trainers/pretrain_scone_vis.py:162-225, train_macarons.py:423-444), where the points and
 * the camera positions are data (no gradient).  Deterministic (fixed summation order). */
int mac_covgain_backward_f32(const float *pts, int pts_dim, const float *harmonics, const float *cams,
                             const float *grad_out, float *grad_harmonics, int B, int P, int C, int act, int per_point,
                             void *stream);

/* Host-buffer form of mac_covgain_f32 (the call a non-torch caller makes): copies pts/harmonics/
 * cams to the device `device` (pinned staging is the caller's business), runs the kernel, copies
 * the [cam_begin, cam_end) columns of `out` back and synchronises.  Device buffers are cached per
 * thread between calls and released by mac_host_release(). */
int mac_covgain_host(const float *pts, int pts_dim, const float *harmonics, const float *cams,
                     float *out, int B, int P, int C, int cam_begin, int cam_end, int act,
                     int device);
void mac_host_release(void);

/* ---------------------------------------------------------------------------------------------
 * Fused score all-gather over NVLink peer memory (multi-GPU, one process per GPU).
 *
 * Every rank owns a "score board" in peer-mapped device memory (e.g. torch symmetric memory):
 *     scores  (B, C) fp32   the assembled score matrix
 *     flags   (world) u32   flags[r] = last epoch for which rank r's columns have landed
 * mac_covgain_push_f32 is mac_covgain_f32 whose finishing CTA, instead of (only) writing a local `out`,
 * stores this rank's columns [cam_begin, cam_end) straight into the score board of EVERY rank (its own
 * included) and then release-stores `epoch` into flags[rank] on every board: compute and exchange are one
 * kernel, there is no separate collective.  `board->scores[r]` / `board->flags[r]` are the addresses of
 * rank r's board as mapped into THIS process.  Epochs must increase by one per step; with two boards
 * used alternately (even / odd epochs) a rank can never overwrite scores a peer is still reading.
 *
 * mac_gather_wait_argmax enqueues a one-CTA kernel that waits (acquire loads, bounded by ~2 s) until all
 * `world` flags of the local board have reached `epoch`, then writes best[b] = index of the first maximum
 * of scores[b, :] (NaN counts as maximal, like torch.argmax); on timeout it sets status[0] = 1 and leaves best
 * untouched.  `status` points to FOUR ints: status[0] is sticky (never reset by the device): the caller zeroes it before the
 * first step and after having handled a timeout; status[1] = nanoseconds this rank waited for its slowest peer in the last
 * step, status[2] / status[3] = accumulated nanoseconds / number of steps since the caller last zeroed them.
 * ------------------------------------------------------------------------------------------- */
#define MAC_MAX_PEERS 16
typedef struct mac_peer_board {
    int world;
    int rank;
    unsigned int epoch;
    float *scores[MAC_MAX_PEERS];
    unsigned int *flags[MAC_MAX_PEERS];
    /* point-sharded steps only (mac_covgain_push_partial_argmax_f32), else unused: partials[r] = rank r's region of
     * world * B * C exact fixed-point sums (int64, slot [s * B * C + i] written by rank s) followed by world * B * C
     * u32 non-finite markers, as mapped into this process */
    void *partials[MAC_MAX_PEERS];
} mac_peer_board_t;

int mac_covgain_push_f32(const float *pts, int pts_dim, const float *harmonics, const float *cams, int B,
                         int P, int C, int cam_begin, int cam_end, int act, void *workspace,
                         size_t workspace_bytes, const mac_peer_board_t *board, void *stream);

int mac_gather_wait_argmax(const float *scores, const unsigned int *flags, int world, unsigned int epoch,
                           int B, int C, long long *best, int *status, void *stream);

/* One sharded scoring step in ONE launch: mac_covgain_push_f32 whose finishing CTA, after raising this rank's flags,
 * also does the work of mac_gather_wait_argmax on this rank's own board (board->scores[board->rank],
 * board->flags[board->rank]): it waits for the flags of all ranks and writes best / status.  At 64 cameras per GPU the
 * step is ~70 us of device time, so neither a second launch nor two host calls with argument marshalling are affordable.
 * ev_begin / ev_end: optional cudaEvent_t recorded on `stream` around the scoring kernel (per-kernel timing). */
int mac_covgain_push_argmax_f32(const float *pts, int pts_dim, const float *harmonics, const float *cams, int B, int P,
                                int C, int cam_begin, int cam_end, int act, void *workspace, size_t workspace_bytes,
                                const mac_peer_board_t *board, long long *best, int *status, void *ev_begin,
                                void *ev_end, void *stream);

/* The same step partitioned over the POINTS instead of the cameras (an extension for callers whose points arrive from
 * the host: every rank then uploads and reads only its own rows): this rank integrates ALL C cameras over its P_local
 * points (pts / harmonics hold only those rows), the finishing CTA stores the exact fixed-point partial sums into slot
 * `rank` of every peer's partial region (board->partials), raises the flags, waits for all ranks, adds the `world` partial
 * sums of every camera -- integer additions, so the scores are bitwise those of one GPU integrating all p_total points --
 * writes the (B, C) means into its own score board (board->scores[rank]) and takes the argmax. */
/* Accumulate a slice of the points of ONE cloud (B = 1) into the workspace without finishing: lets a caller whose points
 * arrive in slices (host -> device copies on another stream) integrate slice k while slice k+1 is in flight; the call that
 * carries the last slice (mac_covgain_push_partial_argmax_f32, or mac_covgain_f32 over the remaining rows with the same
 * workspace) adds these sums to its own. */
int mac_covgain_accumulate_f32(const float *pts, int pts_dim, const float *harmonics, const float *cams, int P_slice, int C,
                               int act, void *workspace, size_t workspace_bytes, void *stream);
size_t mac_covgain_partial_region_bytes(int world, int B, int C);
int mac_covgain_push_partial_argmax_f32(const float *pts, int pts_dim, const float *harmonics, const float *cams, int B,
                                        int P_local, int p_total, int C, int act, void *workspace, size_t workspace_bytes,
                                        const mac_peer_board_t *board, long long *best, int *status, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Fused linear layer on tcgen05 tensor cores (building block of SconeOcc / SconeVis; replaces the
 * nn.Linear + LayerNorm + GELU + residual sequences of /root/reference/macarons/networks/Attention.py:96-98,
 * 160-162, 225-226, 281-298 and the pooling of networks/SconeOcc.py:120-127).
 *
 *   out[m, n] = res[m, n] + act( sum_k X[m, k] * W[n, k] + bias[n] )          (res, bias optional)
 *   ln_out    = LayerNorm_N(out) * ln_g + ln_b                                  (optional, N <= 256)
 *   pool = 16: out (M/16, 2N) = [max | mean] of `out` over groups of 16 consecutive rows
 *
 * X (M, K) row stride ldx, W (N, K) row stride ldw, both 16-byte aligned with strides that are multiples
 * of 4 floats.  W_hi / W_lo are the two TF32 halves of the fp32 weight (W_hi = tf32(W), W_lo = tf32(W - W_hi),
 * see macarons_b200/packing.py); the kernel splits X the same way on chip and accumulates the three cross
 * products in fp32, which reproduces an fp32 matmul to ~1e-6 relative.  W_lo = NULL selects plain
 * single-pass TF32 (10 mantissa bits).
 * ------------------------------------------------------------------------------------------- */
#define MAC_LIN_NONE 0
#define MAC_LIN_RELU 1
#define MAC_LIN_GELU 2 /* exact (erf) GELU, torch.nn.GELU() default */
#define MAC_LIN_ELU 3
#define MAC_LIN_SIGMOID 4

int mac_linear_f32(const float *X, int ldx, const float *W_hi, const float *W_lo, int ldw, const float *bias,
                   float *out, int ldo, int M, int N, int K, int act, const float *res, int ldr, float *ln_out,
                   int ldl, const float *ln_g, const float *ln_b, float ln_eps, int pool, void *stream);
/* mac_linear_f32 with the LayerNorm of the INPUT rows applied on load -- lnin_stats (M, 2) = (mean, rstd) per row,
 * lnin_g / lnin_b (K <= 512), split weights required -- and / or the (mean, rstd) of every OUTPUT row written to
 * stats_out (M, 2) with ln_eps (N <= 256): consecutive layers of /root/reference/macarons/networks/Attention.py:281-298
 * (norm1 -> attention -> residual -> norm2 -> feed-forward) pass 8 bytes per row instead of a normalised copy. */
int mac_linear_lnio_f32(const float *X, int ldx, const float *W_hi, const float *W_lo, int ldw, const float *bias, float *out,
                        int ldo, int M, int N, int K, int act, const float *res, int ldr, const float *lnin_stats,
                        const float *lnin_g, const float *lnin_b, float *stats_out, float ln_eps, void *stream);


/* ---------------------------------------------------------------------------------------------
 * k nearest neighbours (k = 16)  --  replaces get_knn_points, /root/reference/macarons/utility/utils.py:1497-1509
 * (torch.cdist + topk + knn_gather): idx (B, Q, 16) int32 indices into pc (B, N, 3) of the 16 nearest
 * points of every query x (B, Q, 3), nearest first; dist (B, Q, 16) Euclidean distances (may be NULL).
 * No (B, Q, N) distance matrix is materialised.
 * ------------------------------------------------------------------------------------------- */
int mac_knn16_f32(const float *x, const float *pc, int *idx, float *dist, int B, int Q, int N, void *stream);

/* ---------------------------------------------------------------------------------------------
 * SconeOcc / SconeVis forward passes.  Weights are packed once by the caller (macarons_b200/packing.py)
 * into the structs below; all pointers are device pointers owned by the caller.
 *   mac_linear_w_t   one nn.Linear as the two TF32 halves of its (N, K) weight (row stride ldw) + bias
 *   mac_encoder_w_t  one pre-LayerNorm encoder, networks/Attention.py:239-300; qkv = rows [w_q; w_k; w_v]
 *   mac_pct_w_t      PCTransformer, networks/SconeOcc.py:45-130.  emb2 is Embedding.linear2 extended by
 *                    identity rows so that the GEMM also performs the `cat((res, x))` of Attention.py:125
 * ------------------------------------------------------------------------------------------- */
#define MAC_MAX_ENCODERS 4
#define MAC_MAX_SCALES 4
typedef struct mac_linear_w {
    const float *hi, *lo, *bias;
    int N, K, ldw;
} mac_linear_w_t;
typedef struct mac_encoder_w {
    const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
    mac_linear_w_t qkv, out, ff1, ff2;
} mac_encoder_w_t;
typedef struct mac_pct_w {
    const float *emb1_w, *emb1_b; /* Embedding.linear1, plain fp32 (inner, in_dim) */
    int in_dim, inner;
    mac_linear_w_t emb2;
    int n_enc, d_model, dqk, dv;  /* dqk, dv: per-head dims (4 heads) */
    mac_encoder_w_t enc[MAC_MAX_ENCODERS];
    const float *ln_g, *ln_b;
    mac_linear_w_t linear0;
} mac_pct_w_t;
typedef struct mac_sconeocc_w {
    mac_pct_w_t global_pct;
    int n_scale;
    mac_pct_w_t local_pct[MAC_MAX_SCALES];
    const float *xemb1_w, *xemb1_b; /* XEmbedding.linear1, plain fp32 (128, 3) */
    int xemb1_n;
    mac_linear_w_t xemb2, xemb3;
    const float *lin1_wg;           /* linear1 columns that multiply the global feature, plain fp32 (512, 512) */
    int lin1_wg_ld, global_dim;
    const float *lin1_b;
    mac_linear_w_t lin1;            /* remaining columns [local | x | view harmonics]; bias unused */
    mac_linear_w_t lin2, lin3;
} mac_sconeocc_w_t;
typedef struct mac_sconevis_w {
    const float *emb1_w, *emb1_b;   /* Embedding.linear1, plain fp32 (126, 4) */
    int in_dim, inner;
    mac_linear_w_t emb2;
    int n_enc, d_model, dqk, dv;
    mac_encoder_w_t enc[MAC_MAX_ENCODERS];
    const float *ln_g, *ln_b;
    mac_linear_w_t fc1, fc2, fc3;
} mac_sconevis_w_t;

/* SconeVis.forward, /root/reference/macarons/networks/SconeVis.py:121-162 (default architecture:
 * global max-pooled feature, view harmonics concatenated before fc2, no mask):
 *   pts (B, S, 4), view_harmonics (B, S, 64) -> out (B, S, 64) SH coefficients of the visibility gains. */
size_t mac_sconevis_workspace_bytes(int B, int S);
int mac_sconevis_forward_f32(const mac_sconevis_w_t *w, const float *pts, const float *view_harmonics, float *out, int B,
                             int S, void *workspace, size_t workspace_bytes, void *stream);
/* Ragged batch: cloud b holds lens[b] <= S tokens in rows [b*S, b*S + lens[b]) (lens: device, B ints); the padding rows
 * must hold finite values (zeros), are excluded from the global max feature and masked as attention keys; their outputs
 * are unspecified.  One call replaces B separate forwards of different lengths: the per-candidate
 * `macarons(mode='visibility', ...)` calls of /root/reference/macarons/utility/macarons_utils.py:1663. */
int mac_sconevis_forward_ragged_f32(const mac_sconevis_w_t *w, const float *pts, const float *view_harmonics, float *out,
                                    int B, int S, const int *lens, void *workspace, size_t workspace_bytes, void *stream);

/* SconeOcc.forward, /root/reference/macarons/networks/SconeOcc.py:250-347, after the caller has drawn the
 * random sub-samples (torch.randperm, :269 and :311, stays on the host so that the RNG stream matches):
 *   pc_global (B, Sg, 3)           the <= seq_len points fed to the global transformer
 *   pc_scale[s] (B, n_scale_pts[s], 3)   the cloud kNN is taken in at scale s (full, /ds, /ds^2)
 *   x (B, Q, 3) queries, view_harmonics (B, Q, 64)  ->  out (B, Q, 1) = GELU(MLP(...)) occupancy values.
 * Queries are processed `chunk` at a time (results do not depend on it). */
size_t mac_sconeocc_workspace_bytes(int B, int Sg, int chunk, int Q);
int mac_sconeocc_forward_f32(const mac_sconeocc_w_t *w, const float *pc_global, int Sg, const float *const *pc_scale,
                             const int *n_scale_pts, const float *x, const float *view_harmonics, float *out, int B,
                             int Q, int chunk, void *workspace, size_t workspace_bytes, void *stream);

/* SconeOcc over a ragged batch of cells: the per-cell `macarons(mode='occupancy', ...)` calls of
 * compute_scene_occupancy_probability_field, /root/reference/macarons/utility/macarons_utils.py:1443-1518 (one call per
 * occupied cell, each with its own neighbourhood cloud and sub-samples) as ONE forward.
 *   pc_global (n_cells, Sg, 3): the <= seq_len global-transformer points of every cell, zero rows after lens_g[c]
 *   lens_g (n_cells) device ints
 *   pc_scale[s] (total_s, 3), scale_off[s] (n_cells + 1 device ints): the clouds of all cells at scale s, concatenated,
 *       cell c = rows [scale_off[s][c], scale_off[s][c+1]) (>= 16 each); pc_scale / scale_off are HOST arrays of 3 pointers
 *   x (Qtot, 3), view_harmonics (Qtot, 64): the queries of all cells, cell-major; q_off (n_cells + 1 device ints);
 *   cell_of_q (Qtot device ints); max_q = largest number of queries of one cell
 *   -> out (Qtot) occupancy values; per query the arithmetic of mac_sconeocc_forward_f32 on its own cell. */
size_t mac_sconeocc_cells_workspace_bytes(int n_cells, int Sg, int chunk, long long Qtot);
int mac_sconeocc_forward_cells_f32(const mac_sconeocc_w_t *w, int n_cells, const float *pc_global, int Sg, const int *lens_g,
                                   const float *const *pc_scale, const int *const *scale_off, const float *x,
                                   const float *view_harmonics, const int *q_off, const int *cell_of_q, int max_q, float *out,
                                   long long Qtot, int chunk, void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * View state (rows a10-a12): spherical histogram of the visited cameras around every point and its
 * projection onto the SH basis.
 *   mac_view_state_f32      compute_view_state, /root/reference/macarons/utility/scone_utils.py:799-860
 *       pts (B, P, pts_dim) [xyz read], views (V, 3) -> state (B, P, n_elev*n_azim) fp32 in {0, 1};
 *       bin arithmetic (fp32 asin / acos, torch.remainder based floor division, Python `-n // 2` clamps and
 *       the final wrap modulo n_elev*n_azim) follows the reference statement by statement.
 *   mac_view_harmonics_f32  compute_view_harmonics, scone_utils.py:934-960
 *       state (B, P, n_bins), base (64, n_bins) and h_polar (n_bins) from get_all_harmonics_under_degree
 *       -> out (B, P, 64) = sum_j state_j * base_kj * sin(polar_j) * polar_step * azim_step.
 *   mac_gather_bins_f32     the bin permutation of move_view_state_to_view_space, scone_utils.py:928:
 *       out[b, p, j] = in[b, p, index[j]].
 *   mac_viewstate_harm_f32  compute_view_state followed by compute_view_harmonics (the pair of calls at
 *       /root/reference/macarons/testers/shapenet.py:126-131, utility/macarons_utils.py:2818-2877) in one pass:
 *       pts (B, P, pts_dim), views (V, 3), base (64, n_bins), h_polar (n_bins) -> out (B, P, 64); the histogram
 *       never reaches HBM; results are bitwise those of the two separate calls.
 * ------------------------------------------------------------------------------------------- */
int mac_view_state_f32(const float *pts, int pts_dim, const float *views, float *state, int B, int P, int V, int n_elev,
                       int n_azim, void *stream);
int mac_view_harmonics_f32(const float *state, const float *base, const float *h_polar, float *out, int B, int P,
                           int n_elev, int n_azim, void *stream);
int mac_gather_bins_f32(const float *in, const int *index, float *out, int B, int P, int n_bins, void *stream);
int mac_viewstate_harm_f32(const float *pts, int pts_dim, const float *views, const float *base, const float *h_polar,
                           float *out, int B, int P, int V, int n_elev, int n_azim, void *stream);

/* ---------------------------------------------------------------------------------------------
 * sample_proxy_points (row a13), /root/reference/macarons/utility/scone_utils.py:1030-1076, with the uniforms
 * supplied by the caller (the reference draws torch.rand(n_sample, 1) on the tensors' device):
 *   X (N, 3), preds (N, 1), view_harmonics (N, 64), u (n_sample) in [0, 1)
 *   -> res (<= n_sample, 4) [xyz, occupancy] and res_harmonics (<= n_sample, 64) of the UNIQUE picked points in
 *      ascending index order, inverse (n_sample) int64 position of every draw in that list,
 *      counts[0] = points with occupancy > min_occ, counts[1] = number of unique picks (rows of res that are valid).
 * Outputs must be sized for n_sample rows; n_sample <= 4096.
 * ------------------------------------------------------------------------------------------- */
size_t mac_sample_proxy_workspace_bytes(int N);
int mac_sample_proxy_points_f32(const float *X, const float *preds, const float *view_harmonics, const float *u, int N,
                                int n_sample, float min_occ, float *res, float *res_harmonics, long long *inverse,
                                int *counts, void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Field-of-view / occupancy selection + proxy sampling for C candidate cameras at once: the first half of
 * predict_coverage_gain_for_single_camera, /root/reference/macarons/utility/macarons_utils.py:1603-1628, i.e.
 * Camera.get_points_in_fov (:2400-2435: NDC bounds, view z > 0, distance to the camera centre < fov_range),
 * the occupancy threshold (:1610-1613) and sample_proxy_points (scone_utils.py:1030-1076) per candidate.
 *   X (N, 3), preds (N, 1), view_harmonics (N, 64): the scene's proxy points (shared by all candidates)
 *   cams (C, 36): per candidate [full projection 4x4 | world-to-view 4x4 | camera centre 3 | pad], matrices row-major in
 *                 the row-vector convention of pytorch3d (`Transform3d.get_matrix()`: p' = [p, 1] @ M)
 *   ndc_bounds: HOST array (min_x, max_x, min_y, max_y); fov_range < 0 disables the range test
 *   u (C, n_sample) uniforms in [0, 1)
 *   -> per candidate c: res (C, n_sample, 4), res_harmonics (C, n_sample, 64), inverse (C, n_sample) as in
 *      mac_sample_proxy_points_f32 (rows >= counts[2c+1] are left untouched), counts (C, 2) = (points kept, unique
 *      picks), volume (C) = sum of the kept occupancies (`fov_proxy_volume`, :1621; may be null).
 * ------------------------------------------------------------------------------------------- */
/* The resolution filter of Cell.fill for all cells of one Scene.fill_cells call, /root/reference/macarons/utility/
 * macarons_utils.py:2556-2561 (`min(cdist(new.double(), stored.double())) > resolution`, one cdist per cell there):
 * pts (M, 3) new points, slot (M) the cell slot of each, stored (E, 3) the stored points of all slots concatenated,
 * off (n_slots + 1) their row offsets -> out (M) float64 distance to the nearest stored point of the own cell (+inf if none). */
int mac_cell_min_dist_f64(const float *pts, const int *slot, const float *stored, const int *off, double *out, int M,
                          void *stream);

/* Camera.get_points_in_fov for ONE camera, /root/reference/macarons/utility/macarons_utils.py:2400-2435:
 * X (N, 3), cam (36 floats, same layout as a row of `cams` above), ndc_bounds (HOST, 4 floats), fov_range < 0 = no range
 * test -> mask (N) bytes, 1 where the point projects inside the image, lies in front of the camera and within range. */
int mac_points_in_fov_f32(const float *X, const float *cam, const float *ndc_bounds, float fov_range, int N,
                          unsigned char *mask, void *stream);
size_t mac_fov_sample_proxy_workspace_bytes(int N, int C);
int mac_fov_sample_proxy_f32(const float *X, const float *preds, const float *view_harmonics, const float *cams,
                             const float *ndc_bounds, float fov_range, float min_occ, const float *u, int N, int C,
                             int n_sample, float *res, float *res_harmonics, long long *inverse, int *counts, float *volume,
                             void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Depth-side helpers, methods of `Camera` in /root/reference/macarons/utility/macarons_utils.py:
 *   mac_unproject_depth_f32  project_depth_in_3D :2339-2360: depth (B, H, W) metric depth -> out (B, H*W, 3) world points;
 *       cams (B, 18) = [inverse full projection 4x4, row-vector convention | f1 | f2] with f1 = K[2][2], f2 = K[3][2] of
 *       the pytorch3d projection matrix (unproject_points, scaled_depth_input = False); the NDC pixel tables are those
 *       of Camera.__init__ :1929-1938.
 *   mac_signed_distance_f32  get_signed_distance_to_depth_maps :2451-2500: pts (P, 3), depth_maps (n, H, W), mask (n, H, W)
 *       bytes, cams (n, 32) = [full projection 4x4 | world-to-view 4x4] -> out (n, P) = view z - bilinear sample of the
 *       depth map (border padding, align_corners = False; masked pixels read as `fill` = 1.1 zfar).
 * ------------------------------------------------------------------------------------------- */
int mac_unproject_depth_f32(const float *depth, const float *cams, float *out, int B, int H, int W, void *stream);
int mac_signed_distance_f32(const float *pts, const float *depth_maps, const unsigned char *mask, const float *cams, float *out,
                            int n_depth, int P, int H, int W, float fill, void *stream);

/* ---------------------------------------------------------------------------------------------
 * ManyDepth.forward (row a14), /root/reference/macarons/networks/ManyDepth.py:719-758 -> DepthDecoder.forward
 * :474-531 -> CostVolumeBuilder.forward :207-305, inference mode (BatchNorm folded into the convolutions at packing
 * time, ground-truth relative poses already composed into per-frame cameras by the caller, ManyDepth.py:740-750).
 *   mac_conv_w_t   one convolution as the (Cout, kh*kw*Cin) matrix of its weights in (ky, kx, c) order (TF32 halves),
 *                  its bias, geometry, padding mode (reflect = 1 / zeros = 0) and fused activation (MAC_LIN_*)
 *   x (B, 3, H, W) target frames, x_alpha (B, n_alpha, 3, H, W) source frames, NCHW fp32 as the reference passes them
 *   cam (B, 1 + n_alpha, 13): per frame R (9, row-major, X_view = X_world R + T), T (3), zfar (1); record 0 = target
 *   -> disp1 (B, 1, H, W), disp2 (B, 1, H/2, ceil(W/2)), disp3 (.. /4), disp4 (.. /8) sigmoid disparities.
 * ------------------------------------------------------------------------------------------- */
typedef struct mac_conv_w {
    mac_linear_w_t lin;
    int k, stride, pad, reflect, act;
} mac_conv_w_t;
typedef struct mac_block_w { /* torchvision BasicBlock */
    mac_conv_w_t conv1, conv2, down;
    int has_down;
} mac_block_w_t;
typedef struct mac_expansion_w { /* ExpansionLayer, ManyDepth.py:308-365 */
    mac_conv_w_t upconv, iconv;
} mac_expansion_w_t;
typedef struct mac_manydepth_w {
    mac_conv_w_t conv1;
    mac_block_w_t layer1[2], layer2[2], layer3[2], layer4[2];
    mac_conv_w_t conv_reduce;
    mac_expansion_w_t expansion[5]; /* expansion5 .. expansion1 */
    mac_conv_w_t disp[4];           /* disp1 .. disp4 */
    int n_depth;
    float d_min, d_max;
} mac_manydepth_w_t;

size_t mac_manydepth_workspace_bytes(const mac_manydepth_w_t *w, int B, int n_alpha, int H, int W);
int mac_manydepth_forward_f32(const mac_manydepth_w_t *w, const float *x, const float *x_alpha, const float *cam,
                              float *disp1, float *disp2, float *disp3, float *disp4, int B, int n_alpha, int H, int W,
                              void *workspace, size_t workspace_bytes, void *stream);

/* Number of kernel launches the library has enqueued since load (for bench.py's gpu_launches). */
unsigned long long mac_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MACARONS_B200_H */
