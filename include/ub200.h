/*
 * ub200.h -- C ABI of the B200-native uncertainty rendering-and-scoring hot path.
 *
 * Drop-in boundary for the hot path of AaltoML/uncertainty-nerf-gs (the "reference";
 * citations are file:line under /root/reference/nerfuncertainty).  The reference is 100 %
 * Python and has no FFI of its own: this path sits behind its Python plugin surface
 * (model get_outputs dicts, nerfuncertainty.metrics.ause / auce).  Each entry point below
 * replaces the torch / numpy arithmetic of the cited reference lines; the Python host layer
 * (uncertainty_nerf_gs_b200/) binds them with ctypes and mirrors the reference interface.
 * INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - every pointer is a DEVICE pointer unless the name ends in _host.
 *   - the caller allocates all outputs and the workspace; *_workspace_bytes() is the query.
 *   - every call enqueues on `stream` (a cudaStream_t passed as void*) and returns without
 *     synchronising; no global mutable state, no internal threads; re-entrant across
 *     streams and devices (uses the caller's current device).
 *   - return value: UB_OK (0) or a negative ub_status; ub_last_error() returns a
 *     thread-local message for the last failing call on this thread.
 *   - built for sm_100a only; there is no CPU path.
 */
#ifndef UB200_H_
#define UB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UB_ABI_VERSION 2

typedef enum ub_status {
  UB_OK = 0,
  UB_ERR_BAD_ARG = -1,      /* null pointer, negative size, inconsistent arguments        */
  UB_ERR_UNSUPPORTED = -2,  /* shape / alignment / option not supported by this build       */
  UB_ERR_WORKSPACE = -3,    /* workspace null or smaller than *_workspace_bytes()           */
  UB_ERR_LAUNCH = -4        /* CUDA reported an error while enqueuing                       */
} ub_status;

int ub_abi_version(void);
const char* ub_last_error(void);
/* SM count of the current device (grid sizing is a multiple of it). */
int ub_sm_count(void);

/* ------------------------------------------------------------------------------------------
 * (A1) Per-ray front-to-back compositing with the variance term.
 * Replaces, for one batch of rays downstream of the field:
 *   RaySamples.get_weights          models/laplace/laplace_model.py:47-62 (in-repo restatement),
 *                                   called models/activenerfacto/activenerfacto_model.py:94
 *   renderer_rgb/depth/expected/acc models/activenerfacto/activenerfacto_model.py:98-102
 *   beta NaN guard + Sum w^2 beta    models/activenerfacto/activenerfacto_model.py:105-107,
 *                                   models/laplace/laplace_model.py:478-480
 *   depth variance                  models/activenerfacto/activenerfacto_model.py:111-112
 * Inputs are [num_rays, num_samples] row-major float32 (rgb: [num_rays, num_samples, 3]).
 * Outputs are [num_rays] float32 (rgb: [num_rays, 3]); any output pointer may be NULL.
 * ---------------------------------------------------------------------------------------- */
enum {
  UB_BG_LAST_SAMPLE = 0,  /* nerfacto default: background = colour of the last sample        */
  UB_BG_NONE = 1,         /* "random" at eval: return the unblended sum                       */
  UB_BG_FIXED = 2         /* fixed colour in background_rgb                                   */
};
enum {
  UB_BETA_RAW = 0,        /* nerfacto-laplace: beta used as is                                */
  UB_BETA_NAN_GUARD = 1   /* active-nerfacto: if any NaN in the chunk, beta = nan_to_num(beta) */
};

typedef struct ub_composite_rays_args {
  const float* density;      /* [R,S]                                                        */
  const float* deltas;       /* [R,S]; NULL: ends - starts in float32 (what RaySamples.deltas  */
                             /* holds after RayBundle.get_ray_samples) -- one stream less     */
  const float* starts;       /* [R,S]                                                        */
  const float* ends;         /* [R,S]                                                        */
  const float* rgb;          /* [R,S,3]                                                      */
  const float* beta;         /* [R,S] per-sample variance (field "rgb_var"); may be NULL      */
  int64_t num_rays;          /* R >= 0                                                       */
  int32_t num_samples;       /* S >= 1                                                       */
  int32_t background_mode;   /* UB_BG_*                                                      */
  float background_rgb[3];   /* used when background_mode == UB_BG_FIXED                     */
  int32_t beta_mode;         /* UB_BETA_*                                                    */
  int64_t rays_per_chunk;    /* eval chunk (reference: 1 << 15); <= 0 means one chunk.       *
                              * Chunk-wide reductions (clip bounds of the expected depth,    *
                              * the beta NaN guard) are evaluated per chunk, as the          *
                              * reference's chunk loop does.                                 */
  int32_t eval_mode;         /* 1: nan_to_num(rgb) in, clamp [0,1] out (not self.training)   */
  float* out_rgb;            /* [R,3]                                                        */
  float* out_accumulation;   /* [R]                                                          */
  float* out_depth;          /* [R] median depth                                             */
  float* out_expected_depth; /* [R]                                                          */
  float* out_rgb_var;        /* [R]                                                          */
  float* out_rgb_std;        /* [R]                                                          */
  float* out_depth_var;      /* [R]                                                          */
  float* out_depth_std;      /* [R]                                                          */
  float* out_weights;        /* [R,S] optional: the volume-rendering weights                 */
} ub_composite_rays_args;

/* Workspace = [per-chunk table, 16 B per chunk: clip bounds of the expected depth, beta-NaN flag]
 *             [candidate flags, 8 B per 8-ray tile].
 * The compositing kernel flags the rays that the chunk-wide values can change (expected depth outside the ray's own
 * step range; non-finite sum w^2 beta) and the finalize launch visits only those.  A workspace that holds just the
 * chunk table is accepted too: finalize then visits every ray.  The chunk table is what
 * ub_composite_rays_backward reads as `chunk_workspace`. */
size_t ub_composite_rays_workspace_bytes(int64_t num_rays, int64_t rays_per_chunk);
int ub_composite_rays(const ub_composite_rays_args* args, void* workspace, size_t workspace_bytes,
                      void* stream);

/* The same for up to UB_MAX_COMPOSITE_BATCH independent ray batches (the M members of one view:
 * ensemble_pipeline.py:150-157 renders them one after the other): one workspace memset, the compositing kernels
 * back to back, one finalize launch for all batches.  The workspace is one 16-byte aligned region of
 * ub_composite_rays_batch_workspace_bytes(args, n) bytes. */
#define UB_MAX_COMPOSITE_BATCH 8
size_t ub_composite_rays_batch_workspace_bytes(const ub_composite_rays_args* args, int32_t num_batches);
int ub_composite_rays_batch(const ub_composite_rays_args* args, int32_t num_batches, void* workspace,
                            size_t workspace_bytes, void* stream);

/* Backward of ub_composite_rays in training mode (eval_mode 0, beta used as is) w.r.t. density, sample
 * colours and beta: what `ns-train` needs to run active-nerfacto through the fused compositor (losses:
 * models/activenerfacto/activenerfacto_model.py:155-191; the interlevel / distortion losses consume the
 * `weights` output, whose gradient is an input here).  The forward is recomputed from its inputs; only
 * the forward's median depth (a constant: computed under no_grad in the reference, :99-100) and its
 * workspace (clip bounds of the expected depth) are needed.  Every g_* pointer may be NULL (zero
 * gradient); every out_g_* pointer may be NULL.  num_samples in {16, 32, 48, 64, 96}; per-sample tensors
 * 16-byte aligned. */
typedef struct ub_composite_rays_bwd_args {
  const float* density;          /* [R,S] forward inputs                                              */
  const float* deltas;
  const float* starts;
  const float* ends;
  const float* rgb;              /* [R,S,3]                                                           */
  const float* beta;             /* [R,S] or NULL                                                     */
  int64_t num_rays;
  int32_t num_samples;
  int32_t background_mode;       /* UB_BG_*                                                           */
  float background_rgb[3];
  int64_t rays_per_chunk;
  const float* depth;            /* [R] median depth returned by the forward                          */
  const void* chunk_workspace;   /* the forward's workspace (needed with g_expected_depth)            */
  const float* g_rgb;            /* [R,3] incoming gradients                                          */
  const float* g_accumulation;   /* [R]                                                               */
  const float* g_expected_depth; /* [R]                                                               */
  const float* g_rgb_var;        /* [R]                                                               */
  const float* g_rgb_std;        /* [R]                                                               */
  const float* g_depth_var;      /* [R]                                                               */
  const float* g_depth_std;      /* [R]                                                               */
  const float* g_weights;        /* [R,S]                                                             */
  float* out_g_density;          /* [R,S]                                                             */
  float* out_g_rgb;              /* [R,S,3]                                                           */
  float* out_g_beta;             /* [R,S]                                                             */
} ub_composite_rays_bwd_args;

int ub_composite_rays_backward(const ub_composite_rays_bwd_args* args, void* stream);

/* Same renderers driven by given weights instead of densities:
 *   prop_depth_i   models/activenerfacto/activenerfacto_model.py:150-151
 *   depth / depth_var / expected_depth / accumulation from the averaged sampled weights,
 *                  models/laplace/laplace_model.py:509-521
 * Any output may be NULL. */
typedef struct ub_render_weights_args {
  const float* weights;      /* [R,S] */
  const float* starts;       /* [R,S] */
  const float* ends;         /* [R,S] */
  int64_t num_rays;
  int32_t num_samples;
  int64_t rays_per_chunk;
  float* out_accumulation;
  float* out_depth;
  float* out_expected_depth;
  float* out_depth_var;
  float* out_depth_std;
} ub_render_weights_args;

size_t ub_render_weights_workspace_bytes(int64_t num_rays, int64_t rays_per_chunk);
int ub_render_weights(const ub_render_weights_args* args, void* workspace, size_t workspace_bytes,
                      void* stream);

/* Mean volume-rendering weights over num_draws Gaussian density draws (nerfacto-laplace with sampled
 * density): replaces models/laplace/laplace_model.py:486-507 without materialising [draws, R, S].
 * density, density_var, deltas [R,S]; noise: optional standard-normal draws [num_draws, R, S] (NULL:
 * in-kernel Philox4x32-10 seeded with `seed`, statistical parity only); out_weights [R,S] feeds
 * ub_render_weights.  num_samples <= 256. */
int ub_average_sampled_weights(const float* density, const float* density_var, const float* deltas,
                               const float* noise, int64_t num_rays, int32_t num_samples,
                               int32_t num_draws, uint64_t seed, float* out_weights, void* stream);

/* ------------------------------------------------------------------------------------------
 * (B) Fused per-pixel mean / variance across K ensemble members or MC-dropout passes.
 * Replaces torch.stack(...).mean(0) / .std(0).mean(-1) / .var(0).mean(-1):
 *   models/mcdropout/mcdropout_models.py:121-126
 *   models/ensemble/ensemble_pipeline.py:159-190
 * members_host: HOST array of num_members DEVICE pointers, each [num_pixels, channels] float32
 * (no stacked copy is made).  out_mean [num_pixels, channels]; out_spread [num_pixels]:
 * unbiased std (spread_mode 1) or unbiased variance (spread_mode 2) over members, averaged
 * over channels; spread_mode 0 / out_spread NULL skips it.
 * ---------------------------------------------------------------------------------------- */
enum { UB_SPREAD_NONE = 0, UB_SPREAD_STD = 1, UB_SPREAD_VAR = 2 };
#define UB_MAX_MEMBERS 32
#define UB_MAX_REDUCE_JOBS 16
#define UB_MAX_REDUCE_BATCH_MEMBERS UB_MAX_MEMBERS

int ub_reduce_members(const float* const* members_host, int32_t num_members, int64_t num_pixels,
                      int32_t channels, int32_t spread_mode, float* out_mean, float* out_spread,
                      void* stream);

/* All output keys of one view in ONE launch: job j reduces its own set of num_members member
 * tensors (the loop over keys of mcdropout_models.py:121-126 / ensemble_pipeline.py:159-190). */
typedef struct ub_reduce_job {
  const float* const* members_host; /* HOST array of num_members DEVICE pointers [num_pixels, channels] */
  int64_t num_pixels;
  int32_t channels;
  int32_t spread_mode;              /* UB_SPREAD_* */
  float* out_mean;                  /* [num_pixels, channels] or NULL */
  float* out_spread;                /* [num_pixels] or NULL */
} ub_reduce_job;

int ub_reduce_members_batched(const ub_reduce_job* jobs_host, int32_t num_jobs, int32_t num_members,
                              void* stream);

/* ------------------------------------------------------------------------------------------
 * (C1) Metric prologue, NLL and AUCE interval histogram for a batch of images (segments).
 * Replaces scripts/eval_uncertainty.py:323-333 (se / ae / var), :404-412 (NLL), and the 99
 * interval-coverage passes of metrics/auce.py:18-28.
 * pred, target: [total_pixels, channels]; std: [total_pixels] (one sigma per pixel, shared by
 * the channels, as eval_uncertainty.py:371-374 repeats it).  seg_offsets: DEVICE array of
 * num_segments+1 int64 pixel offsets (segment s = [off[s], off[s+1])); segment tables live on the
 * device so that no call on this path performs a host->device copy (a pageable copy would
 * synchronise the stream).
 * z_values: DEVICE [num_z] float64, strictly decreasing (norm.ppf(1 - alpha/2)).
 * Outputs: out_sq_err / out_abs_err / out_var [total_pixels] (may be NULL);
 * out_sums [num_segments, UB_PROLOGUE_NSUMS] float64:
 *     0 sum(se) 1 sum(ae) 2 sum(var) 3 sum(nll over pixels and channels) 4 sum(interval sigma)
 * out_hist [num_segments, num_z + 1] int64: hist[c] = number of (pixel, channel) elements
 *     whose interval predicate  t >= m - z_k s  &&  t <= m + z_k s  (float64, NumPy >= 2
 *     promotion) holds for exactly the first c thresholds; coverage count at k = sum_{c>k} hist[c].
 * ---------------------------------------------------------------------------------------- */
#define UB_PROLOGUE_NSUMS 5

typedef struct ub_score_prologue_args {
  const float* pred;
  const float* target;
  const float* std;
  int32_t channels;               /* 1 or 3 */
  int32_t num_segments;
  const int64_t* seg_offsets;     /* DEVICE [num_segments + 1] */
  int64_t max_segment_len;        /* length of the longest segment */
  float nll_min_std;              /* eps of negative_gaussian_loglikelihood */
  int32_t sigma_from_var;         /* 1: interval sigma = sqrt(std*std) as eval_uncertainty.py:371 (rgb); *
                                   * 0: sigma = std (depth, eval_uncertainty.py:612-614)                */
  const double* z_values;
  int32_t num_z;                  /* <= 127 */
  float* out_sq_err;
  float* out_abs_err;
  float* out_var;
  double* out_sums;
  int64_t* out_hist;
  uint32_t* out_coarse_hist;      /* optional DEVICE [3, num_segments, 4096]: histograms of the top 12 bits of the     *
                                   * order-preserving keys of out_var / out_abs_err / out_sq_err (in this order), the   *
                                   * `coarse_hist` input of ub_cut_select_sums_ex; needs the three vectors; NULL: skip   */
} ub_score_prologue_args;

/* max_segment_len: length of the longest segment (workspace scales with it, not with the total). */
size_t ub_score_prologue_workspace_bytes(int32_t num_segments, int64_t max_segment_len, int32_t num_z);
int ub_score_prologue(const ub_score_prologue_args* args, void* workspace, size_t workspace_bytes,
                      void* stream);

/* ------------------------------------------------------------------------------------------
 * (C2) Per-image segmented stable radix sort and cumulative-error scans for AUSE.
 * Replaces torch.sort + the 2 x 100 slice means of metrics/ause.py:10-34.
 * Ordering contract == torch.sort(stable=True) on float32: ascending, ties keep ascending
 * original index, -0.0 == +0.0, every NaN after +inf.
 * keys [total] float32 segmented by seg_offsets (DEVICE int64 [num_segments+1]); total and
 * max_segment_len are the host-side copies of off[num_segments] and the longest segment.
 * out_sorted_keys [total] (may be NULL); out_perm [total] int32: index *within its segment* of
 * the element at each sorted position (NULL for a keys-only sort).
 * ---------------------------------------------------------------------------------------- */
size_t ub_segmented_sort_workspace_bytes(int32_t num_segments, int64_t total, int64_t max_segment_len,
                                         int32_t with_perm);
int ub_segmented_sort(const float* keys, int32_t num_segments, const int64_t* seg_offsets,
                      int64_t total, int64_t max_segment_len, float* out_sorted_keys, int32_t* out_perm, void* workspace,
                      size_t workspace_bytes, void* stream);

/* Prefix sums at cut points, float64 accumulation:
 *   out_sums[s, v, c] = sum_{i < cuts[s, c]} values_v[off[s] + (perm_v ? perm_v[off[s] + i] : i)]
 * values_host: HOST array of num_values (<= 8) DEVICE pointers [total] float32; perms_host: HOST
 * array of num_values DEVICE int32 pointers as produced by ub_segmented_sort (entries, or the whole
 * array, may be NULL = identity); cuts: DEVICE [num_segments, num_cuts] int64, each in
 * [0, segment length] (caller's contract); out_sums: DEVICE [num_segments, num_values, num_cuts] float64. */
size_t ub_cut_prefix_sums_workspace_bytes(int32_t num_segments, int64_t max_segment_len,
                                          int32_t num_values, int32_t num_cuts);
int ub_cut_prefix_sums(const float* const* values_host, const int32_t* const* perms_host,
                       int32_t num_values, int32_t num_segments, const int64_t* seg_offsets,
                       int64_t max_segment_len, const int64_t* cuts, int32_t num_cuts, double* out_sums,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Host tail of the batched scorer (HOST function, HOST pointers, no stream): from the packed results of a batch --
 * packed = [num_views][4][num_cuts] AUSE slice sums (abs err by var, sq err by var, abs err ascending, sq err
 * ascending) | [num_views][5] prologue sums (se, ae, var, nll, sigma) | [num_views][num_z + 1] interval histogram
 * (int64 bit patterns), as ub_cut_select_sums / ub_score_prologue wrote them and one device->host copy brought them
 * over -- to what get_unc_metrics_rgb returns per image (eval_uncertainty.py:323-402): for the error types (mae, mse,
 * rmse) the normalised curves of metrics/ause.py:15-44 -- out_by_unc [V][3][num_cuts] float64, the oracle curve as
 * float64 (out_oracle64) and as float32 (out_oracle32) with out_oracle_is64 [V][3] saying which of the two the
 * reference's dtype rules produce -- and out_ause [V][3]; out_scalars [V][3] = nll, avg_var, mse mean (float32);
 * out_auce_curves [V][5][num_z] = coverage, interval length, coverage error, abs, neg (metrics/auce.py:24-45);
 * out_auc [V][3] = areas of abs error, length, neg error (:47-54).  num_pixels [V]: pixels per view (ragged for the
 * masked depth modality, get_unc_metrics_depth :415-644 with channels = 1); cuts [V][num_cuts]: its slice lengths
 * int((1-r) n);
 * ratio_steps [num_cuts - 1] = diff of the removal ratios; one_minus_alpha [num_z], alpha_steps [num_z - 1].
 * Every operation runs in the dtype and order numpy uses for the reference's expressions (pairwise row sums of
 * np.trapz included): the results equal the numpy evaluation bit for bit. */
int ub_score_tail_host(const double* packed, int32_t num_views, const int64_t* num_pixels, int32_t channels,
                       const int64_t* cuts, int32_t num_cuts, const double* ratio_steps, const double* z_values, int32_t num_z,
                       const double* one_minus_alpha, const double* alpha_steps, double* out_by_unc,
                       double* out_oracle64, float* out_oracle32, int32_t* out_oracle_is64, double* out_ause,
                       float* out_scalars, double* out_auce_curves, double* out_auc);

/* AUSE cut-point sums without a full sort (multi-cut radix select).  Same result as ub_segmented_sort
 * followed by ub_cut_prefix_sums -- the sums over the first cuts[s, c] elements of the stable ascending
 * order of the keys (metrics/ause.py:10-20 with the error as its own key, :25-34 with the uncertainty as
 * key and the errors as payload) -- when the permutation itself is not needed: keys are classified against
 * the cut positions (count-proportional fine bins over a 4096-bin coarse histogram, tie groups resolved by element
 * index, the few undecided keys around every cut ranked per cell), payloads summed per class exactly (fixed point)
 * and returned as float64.
 * Family f (<= 4) has keys keys_host[f] [total] and one or two payload arrays pay0_host[f], pay1_host[f]
 * (HOST arrays of DEVICE pointers; pay1_host or its entries may be NULL; pay0 == keys is recognised).
 * All families share seg_offsets (DEVICE int64 [num_views + 1], off[0] == 0) and cuts (DEVICE
 * [num_views, num_cuts] int64, num_cuts <= 128, any order, clamped to the segment length).
 * out_sums: DEVICE [num_views, V, num_cuts] float64, V = total number of payload arrays, rows in family
 * order.  max_segment_len <= 2^24 (UB_ERR_UNSUPPORTED above: use the sort). */
size_t ub_cut_select_sums_workspace_bytes(int32_t num_families, int32_t num_views, int64_t total,
                                          int64_t max_segment_len, int32_t num_cuts);
int ub_cut_select_sums(const float* const* keys_host, const float* const* pay0_host,
                       const float* const* pay1_host, int32_t num_families, int32_t num_views,
                       const int64_t* seg_offsets, int64_t total, int64_t max_segment_len, const int64_t* cuts,
                       int32_t num_cuts, double* out_sums, void* workspace, size_t workspace_bytes,
                       void* stream);
/* Same, with the coarse key histograms supplied by the caller: coarse_hist DEVICE [num_families, num_views, 4096]
 * uint32, entry [f][v][b] = number of keys of family f in segment v whose order-preserving key (float32 bits made
 * monotone: -0.0 == +0.0, NaN last) has top 12 bits b -- what ub_score_prologue writes to out_coarse_hist for the
 * families (var, abs err, sq err).  Saves the first pass over the keys.  NULL: computed here. */
int ub_cut_select_sums_ex(const float* const* keys_host, const float* const* pay0_host,
                          const float* const* pay1_host, int32_t num_families, int32_t num_views,
                          const int64_t* seg_offsets, int64_t total, int64_t max_segment_len, const int64_t* cuts,
                          int32_t num_cuts, const uint32_t* coarse_hist, double* out_sums, void* workspace,
                          size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * (A3) Last-layer diagonal-Laplace MC moments.
 * Replaces NerfactoLaplaceField.sample_laplace, models/laplace/laplace_field.py:528-568, for a
 * linear head out = act(x W^T + b): for each of n_samples parameter draws theta_s (rows of
 * sampled_params [n_samples, out_dim*hidden + out_dim], weight row-major then bias -- the
 * order of torch parameters_to_vector) accumulate E[y] and E[y^2]; sigma2 = E[y^2] - E[y]^2.
 * activation: 0 identity, 1 sigmoid (rgb head, :468-476), 2 trunc_exp == exp (density head,
 * :331-339).  out_mean / out_sigma2 [num_points, out_dim]; out_mean2 optional.
 * The rgb head (out_dim 3, 3 * n_samples <= 304) runs as a [P,64]x[64,300] GEMM on the tensor cores
 * (tcgen05.mma kind::tf32 with a 3-term hi/lo split, float32 accumulators in tensor memory); the
 * density head and other shapes use the fp32-FMA kernel.
 * ---------------------------------------------------------------------------------------- */
enum { UB_ACT_IDENTITY = 0, UB_ACT_SIGMOID = 1, UB_ACT_EXP = 2 };
/* OR-ed into `activation`: keep the rgb head on the fp32-FMA kernel instead of the tcgen05 3xTF32 path */
#define UB_ACT_FLAG_NO_TENSOR_CORES 0x100

int ub_laplace_ll_moments(const float* x, int64_t num_points, int32_t hidden, int32_t out_dim,
                          const float* sampled_params, int32_t n_samples, int32_t activation,
                          float* out_mean, float* out_mean2, float* out_sigma2, void* stream);

/* ------------------------------------------------------------------------------------------
 * (A2) Alpha compositing over pre-binned per-tile splat lists (16x16 tiles).
 * Replaces the four gsplat.rasterize_gaussians passes of
 * models/activesplatfacto/activesplatfacto_model.py:260-356 by one fused pass over
 * rgb[3] + beta[1] + depth[1] and a second pass for the depth variance.
 * xys [G,2], conics [G,3], opacities [G], colors [G, channels] float32;
 * gaussian_ids [num_intersects] int32 sorted by (tile, depth); tile_bins [tiles, 2] int32.
 * out [H, W, channels] = sum_i c_i alpha_i T_i + T_final * background; out_alpha [H, W] = 1 - T_final.
 * ---------------------------------------------------------------------------------------- */
#define UB_TILE 16
#define UB_MAX_SPLAT_CHANNELS 8
#define UB_MAX_SPLAT_PLANES 4

int ub_composite_tiles(const float* xys, const float* conics, const float* opacities,
                       const float* colors, int32_t channels, const int32_t* gaussian_ids,
                       const int32_t* tile_bins, int32_t img_height, int32_t img_width,
                       const float* background_host, float* out, float* out_alpha, void* stream);

/* Same pass with the colours given as up to UB_MAX_SPLAT_PLANES separate tensors ("planes", e.g. rgb [G,3],
 * beta [G,1], depth [G,1]; at most UB_MAX_SPLAT_CHANNELS channels in total) and one output image per plane
 * (outs_host[p]: [H, W, plane_channels[p]], may be NULL).  planes_host / plane_channels_host / outs_host are
 * HOST arrays.  background_host: one value per channel in plane order (NULL = 0).  channel_max_keys:
 * optional DEVICE uint32 [total channels]; receives, per channel, the order-preserving key of the maximum
 * of the output image (the `depth_im.max()` of activesplatfacto_model.py:319; consumed by
 * ub_splat_normalize). */
int ub_composite_tiles_planes(const float* xys, const float* conics, const float* opacities,
                              const float* const* planes_host, const int32_t* plane_channels_host,
                              int32_t num_planes, const int32_t* gaussian_ids, const int32_t* tile_bins,
                              int32_t img_height, int32_t img_width, const float* background_host,
                              float* const* outs_host, float* out_alpha, uint32_t* channel_max_keys,
                              void* stream);

/* Diagnostic (parity tests): sigma and alpha = min(0.999, opacity * __expf(-sigma)) of the list entries
 * gaussian_ids[first .. first + count) for the 256 pixel centres of tile (tile_x, tile_y), exactly as the compositing
 * kernels evaluate them (same device function, rounding pinned).  out_sigma / out_alpha: DEVICE [count, 16, 16],
 * row-major inside the tile.  Lets a host oracle replay gsplat's threshold decisions (sigma < 0, alpha < 1/255,
 * T (1 - alpha) <= 1e-4; gsplat 0.1.11 rasterize_forward, reached from activesplatfacto_model.py:260-355) in the
 * same float32 arithmetic. */
int ub_tile_alpha_probe(const float* xys, const float* conics, const float* opacities,
                        const int32_t* gaussian_ids, int32_t first, int32_t count, int32_t tile_x,
                        int32_t tile_y, float* out_sigma, float* out_alpha, void* stream);

/* In-place post-processing of one output plane [num_pixels, channels]:
 *   clamp_max_one   : image = min(image, 1)                       activesplatfacto_model.py:275
 *   divide_by_alpha : image = alpha > 0 ? image / alpha : max     activesplatfacto_model.py:319, 356
 * max_key: DEVICE pointer to the plane's entry of channel_max_keys.  Optional derived planes of the same shape,
 * written from the post-processed image: out_square = image^2 (rgb_var = uncertainty ** 2, :364), out_sqrt =
 * sqrt(image) (depth_std = depth_var.sqrt(), :367); NULL to skip. */
int ub_splat_normalize(float* image, int32_t channels, const float* alpha, int64_t num_pixels,
                       int32_t clamp_max_one, int32_t divide_by_alpha, const uint32_t* max_key, float* out_square,
                       float* out_sqrt, void* stream);

/* out_sq_residual[g] = (depth_g - depth_image[floor(y_g), floor(x_g)])^2 when the centre pixel satisfies
 * 0 < x < W and 0 < y < H (strict, as the reference's mask), else depth_g^2: the colours of the
 * depth-variance pass, activesplatfacto_model.py:325-349.  depth_image [H, W] is the normalised depth. */
int ub_splat_depth_residual(const float* xys, const float* depths, const float* depth_image,
                            int32_t img_height, int32_t img_width, int64_t num_gaussians,
                            float* out_sq_residual, void* stream);

/* ------------------------------------------------------------------------------------------
 * (a15) Inputs of the depth scorer.  Replaces the per-view torch lines of get_unc_metrics_depth
 * (scripts/eval_uncertainty.py:461-462 scale, :511-513 clamp to [1e-3, max gt], :552-560 keep gt > 0).
 * depth, depth_std, depth_gt [B, N] float32, scales [B] float32 (DEVICE).  Outputs: the kept pixels of all views
 * in row-major order, view after view (capacity B*N each), and out_offsets [B+1] int64 (DEVICE): view b owns
 * [out_offsets[b], out_offsets[b+1]) -- ragged segments for ub_score_prologue / ub_segmented_sort.
 * ---------------------------------------------------------------------------------------- */
size_t ub_depth_prepare_workspace_bytes(int32_t num_views, int64_t pixels_per_view);
int ub_depth_prepare(const float* depth, const float* depth_std, const float* depth_gt, const float* scales,
                     int32_t num_views, int64_t pixels_per_view, float* out_pred, float* out_std, float* out_gt,
                     int64_t* out_offsets, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * (f1) Backward of ub_composite_tiles_planes, for training through the fused pass (the reference's losses,
 * activesplatfacto_model.py:369-441, read outputs["rgb"] and outputs["uncertainty"]; gsplat supplies this
 * gradient inside rasterize_gaussians' autograd function).
 * Inputs as the forward call, plus v_outs_host[p] = dL/d(out plane p) [H,W,ch_p] (NULL: zero) and
 * v_alpha = dL/d(alpha) [H,W] (NULL: zero).  Outputs (zeroed here, then accumulated with float32 atomics):
 * v_xys [G,2], v_conics [G,3], v_opacities [G], v_planes_host[p] [G,ch_p] (NULL: not wanted).
 * ---------------------------------------------------------------------------------------- */
int ub_composite_tiles_planes_backward(const float* xys, const float* conics, const float* opacities,
                                       const float* const* planes_host, const int32_t* plane_channels_host,
                                       int32_t num_planes, const int32_t* gaussian_ids, const int32_t* tile_bins,
                                       int32_t img_height, int32_t img_width, const float* background_host,
                                       const float* const* v_outs_host, const float* v_alpha,
                                       int64_t num_gaussians, float* v_xys, float* v_conics, float* v_opacities,
                                       float* const* v_planes_host, void* stream);

/* ------------------------------------------------------------------------------------------
 * (f2) Producers of the splat path: EWA projection and view-dependent colours.  Replace gsplat's
 * project_gaussians (activesplatfacto_model.py:221-234) and spherical_harmonics (:242-249), forward only (eval).
 *   means3d [G,3], scales [G,3] (already exp'ed), quats [G,4] (w,x,y,z; normalised inside),
 *   viewmat: DEVICE pointer to the [3,4] row-major world->camera matrix, clip_thresh: near plane (gsplat: 0.01).
 *   Outputs: xys [G,2], depths [G], radii [G] int32, conics [G,3]; optional (may be NULL): compensation [G],
 *   num_tiles_hit [G] int32, cov3d [G,6].  Culled Gaussians (behind the near plane, zero determinant, no tile)
 *   get radius 0 and zeros.
 *   ub_spherical_harmonics: degree in [0,3] gives the coefficient layout coeffs [G,(degree+1)^2,3]; the first
 *   (degrees_to_use+1)^2 bases are evaluated at viewdirs [G,3] (normalised inside) -> out_colors [G,3].
 * ---------------------------------------------------------------------------------------- */
int ub_project_gaussians(const float* means3d, const float* scales, float glob_scale, const float* quats,
                         const float* viewmat, float fx, float fy, float cx, float cy, int32_t img_height,
                         int32_t img_width, float clip_thresh, int64_t num_gaussians, float* out_xys,
                         float* out_depths, int32_t* out_radii, float* out_conics, float* out_compensation,
                         int32_t* out_num_tiles_hit, float* out_cov3d, void* stream);
int ub_spherical_harmonics(int32_t degree, int32_t degrees_to_use, const float* viewdirs, const float* coeffs,
                           int64_t num_gaussians, float* out_colors, void* stream);

/* ------------------------------------------------------------------------------------------
 * (f2) Tile binning of projected Gaussians (setup step; the compositing above takes its outputs).
 * Replaces what gsplat does inside every rasterize_gaussians call of activesplatfacto_model.py:260-355
 * (tile rectangle = centre +- radius in tile units, (tile << 32 | depth) keys, stable sort, tile ranges).
 * Two phases because the output size is data dependent:
 *   ub_bin_count      out_offsets [G] = exclusive prefix of the per-Gaussian tile counts, *out_total (DEVICE
 *                     int64) = number of intersections I; the caller reads it back to size the outputs;
 *   ub_bin_gaussians  out_gaussian_ids [I] int32 sorted by (tile, depth, intersection order),
 *                     out_tile_bins [tiles, 2] int32 ([p, p) for an empty tile, p = start of the next one).
 * xys [G,2], depths [G] float32, radii [G] int32 (radius <= 0: culled).
 * ---------------------------------------------------------------------------------------- */
size_t ub_bin_count_workspace_bytes(int64_t num_gaussians);
int ub_bin_count(const float* xys, const int32_t* radii, int64_t num_gaussians, int32_t img_height,
                 int32_t img_width, int64_t* out_offsets, int64_t* out_total, void* workspace,
                 size_t workspace_bytes, void* stream);
size_t ub_bin_gaussians_workspace_bytes(int64_t num_intersections);
int ub_bin_gaussians(const float* xys, const float* depths, const int32_t* radii, int64_t num_gaussians,
                     int32_t img_height, int32_t img_width, const int64_t* offsets, int64_t num_intersections,
                     int32_t* out_gaussian_ids, int32_t* out_tile_bins, void* workspace, size_t workspace_bytes,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UB200_H_ */
