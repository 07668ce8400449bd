// (A1) Per-ray front-to-back compositing with the variance term, one pass.
//
// Reference arithmetic (file:line under /root/reference/nerfuncertainty):
//   weights        models/laplace/laplace_model.py:47-62   (== nerfstudio RaySamples.get_weights)
//   rgb/depth/...  models/activenerfacto/activenerfacto_model.py:98-112
//   Sum w^2 beta   models/activenerfacto/activenerfacto_model.py:105-107, laplace_model.py:478-480
//
// Data layout in HBM: six streams of [R,S] float32 rows (rgb: [R,S,3]); a tile of 8 consecutive
// rays is contiguous in every stream.  Fast path (composite_rays_tma<S, NCW, NST>):
//   * persistent CTAs, one per SM: warps 0..NCW-1 = consumers, warp NCW = producer;
//   * the producer's elected lane stages 8-ray tiles (12 KB at S = 48) into a ring of NST
//     shared-memory stages with 1-D bulk async copies (cp.async.bulk -> UBLKCP) signalled on
//     full/empty mbarriers; tiles are dealt to consumer warps round-robin;
//   * a consumer warp owns a tile: 4 lanes per ray, S/4 consecutive samples per lane,
//     conflict-free LDS.128 reads, arithmetic from registers, 2-step shuffles inside the ray;
//   * the two prefix scans (optical depth, cumulative weight) run in float64 and every prefix is
//     rounded to float32 -- the semantics of torch.cumsum on CPU float32, which decides the
//     median-depth index.  (float64 sums of float32 terms are exact unless the terms span more
//     than 2^29, so the lane-parallel association equals the sequential one.)
//   * nan_to_num of weights / colours and the beta NaN guard are applied on a rare slow path that
//     is entered only when a per-ray sum comes out non-finite (a non-finite term always makes the
//     sum non-finite), so the common path carries no per-sample NaN tests.
// Generic path (composite_rays_generic): warp per ray, any S / alignment, same outputs.
//
// Chunk-wide reductions of the reference (clip bounds of expected depth = min/max of steps over
// the eval chunk; the `isnan(beta).any()` guard) are accumulated per chunk in the workspace and
// applied by a small finalize kernel.  That kernel does not re-read every ray: the compositing kernel
// flags, per 8-ray tile, the few rays the chunk-wide values can possibly change (expected depth outside
// the ray's own step range: empty / near-empty rays; non-finite sum w^2 beta), and finalize reads one
// 8-byte word per tile and only visits flagged rays (1 B/ray instead of a 12 B/ray pass over the outputs).
#include <stdlib.h>

#include "ub_common.cuh"

namespace ub {

struct CompositeParams {
  const float* density;
  const float* deltas;
  const float* starts;
  const float* ends;
  const float* rgb;
  const float* beta;
  const float* weights_in;  // render-from-weights mode (generic kernel only)
  long long num_rays;
  int num_samples;
  int bg_mode;
  float bg[3];
  int beta_mode;
  long long rays_per_chunk;
  int tiles_per_chunk;  // rays_per_chunk / 8 on the fast path (0: one chunk)
  int chunk_shift;      // log2(tiles_per_chunk) when it is a power of two (the usual 32768-ray chunk), else -1
  int eval_mode;
  float* o_rgb;
  float* o_acc;
  float* o_depth;
  float* o_exp;
  float* o_rgb_var;
  float* o_rgb_std;
  float* o_dvar;
  float* o_dstd;
  float* o_w;
  unsigned* chunk_ws;  // [num_chunks][4]: max key(steps), max ~key(steps), beta-has-NaN, pad
  // Candidate flags of the finalize pass (fast path), one 64-bit word per 8-ray tile, written unconditionally by the
  // tile's warp (no atomics, no memset): low half = ballot of the rays whose expected depth lies outside the [min, max]
  // of their OWN steps -- a superset of the rays the chunk-wide clip can change, since the chunk's bounds enclose the
  // ray's --, high half = ballot of the rays whose sum w^2 beta came out non-finite (the only ones the NaN-guard redo
  // can change).  NULL: finalize visits every ray.
  unsigned long long* cand_flags;
};

constexpr int kLanesPerRay = 4;
constexpr int kRaysPerTile = 32 / kLanesPerRay;  // 8

__device__ __forceinline__ float clamp01_keep_nan(float v) {  // torch.clamp_ propagates NaN
  return v != v ? v : fminf(fmaxf(v, 0.0f), 1.0f);
}
__device__ __forceinline__ float reduce4(float v) {
  v += __shfl_xor_sync(FULL_MASK, v, 1);
  v += __shfl_xor_sync(FULL_MASK, v, 2);
  return v;
}
__device__ __forceinline__ bool non_finite(float v) { return !(fabsf(v) <= FLT_MAX); }

// Per-warp running min/max of `steps` for the current chunk, flushed with two atomics when the
// chunk changes (identity of both atomicMax targets is 0, so the workspace is memset to 0).
struct ChunkBounds {
  long long chunk = -1;
  float lo = INFINITY, hi = -INFINITY;
  bool has_nan = false;
  __device__ __forceinline__ void flush(unsigned* ws) {
    if (chunk >= 0) {
      float a = lo, b = hi;
      int n = has_nan ? 1 : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a = fminf(a, __shfl_xor_sync(FULL_MASK, a, o));
        b = fmaxf(b, __shfl_xor_sync(FULL_MASK, b, o));
        n |= __shfl_xor_sync(FULL_MASK, n, o);
      }
      if ((threadIdx.x & 31) == 0) {
        if (a <= b) {
          atomicMax(ws + chunk * 4 + 0, order_key(b));
          atomicMax(ws + chunk * 4 + 1, ~order_key(a));
        }
        if (n) ws[chunk * 4 + 2] = 1u;
      }
    }
    lo = INFINITY;
    hi = -INFINITY;
    has_nan = false;
  }
  __device__ __forceinline__ void enter(long long c, unsigned* ws) {
    if (c != chunk) {  // warp-uniform
      flush(ws);
      chunk = c;
    }
  }
  __device__ __forceinline__ void add_step(float s) {
    lo = fminf(lo, s);
    hi = fmaxf(hi, s);
  }
};

// deltas == NULL: the bin widths are ends - starts in float32, which is what nerfstudio's RayBundle.get_ray_samples stores
__device__ __forceinline__ float delta_at(const CompositeParams& p, size_t i) {
  return p.deltas ? p.deltas[i] : __fsub_rn(p.ends[i], p.starts[i]);
}

// exclusive prefix over the 4 lanes of a ray of the lanes' float64 totals: an inclusive scan in two shuffle-up steps
// minus the lane's own total (float64 sums of float32 terms: exact, so the association does not matter)
__device__ __forceinline__ double ray_exclusive_offset(double total, int q) {
  double v = total;
  double u = shfl_up_double(FULL_MASK, v, 1);
  if (q >= 1) v += u;
  u = shfl_up_double(FULL_MASK, v, 2);
  if (q >= 2) v += u;
  return v - total;
}

template <int S, int NCW, int NST>
__global__ void __launch_bounds__(32 * (1 + NCW), 1) composite_rays_tma(const CompositeParams p) {
  constexpr int P = S / kLanesPerRay;  // samples per lane
  constexpr int V = P / 4;             // float4 per lane per scalar stream
  static_assert(S % 16 == 0, "fast path needs S % 16 == 0");
  // Tiles are dealt round-robin to the NCW consumer warps and to the NST ring stages.  NST must be a
  // multiple of NCW so that every stage is always consumed by the same warp: bulk copies of different
  // stages complete out of order, and a *different* warp waiting for tile t+NST on a stage whose
  // tile t is still in flight would see the opposite phase parity as "complete" and read early.
  static_assert(NST % NCW == 0, "ring stages must be a multiple of the consumer warps");
  constexpr int kTileFloats = kRaysPerTile * S;  // one scalar stream of one tile
  constexpr int kStageFloats = kTileFloats * 8;  // 5 scalar streams + rgb (3x)
  constexpr uint32_t kStageBytes = kStageFloats * 4;
  // (float)x >= 0.5f  <=>  x >= 0.5 - 2^-26 for a double x under round-to-nearest-even
  const double kHalf = 0.5 - 1.4901161193847656e-08;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stages = reinterpret_cast<float*>(smem_raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NST * kStageBytes);
  uint64_t* empty_bar = full_bar + NST;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = (int)((p.num_rays + kRaysPerTile - 1) / kRaysPerTile);
  const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const bool has_beta = p.beta != nullptr;
  const bool derive = p.deltas == nullptr;  // deltas = ends - starts (what RayBundle.get_ray_samples stores): one stream less

  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == NCW) {
    // ===== producer (the highest warp id: the issue arbiter favours it over the consumers, so the
    // copy stream never starves): one elected lane issues all bulk copies =====
    if (lane == 0) {
      for (int j = 0; j < my_tiles; ++j) {
        const int st = j % NST;
        const uint32_t ph = (uint32_t)((j / NST) & 1);
        mbar_wait(&empty_bar[st], ph ^ 1u, 1000000 + j);
        const long long ray0 = ((long long)j * gridDim.x + blockIdx.x) * kRaysPerTile;
        const int n = (int)min((long long)kRaysPerTile, p.num_rays - ray0);
        const uint32_t sb = (uint32_t)n * S * 4u;  // bytes of one scalar stream
        float* dst = stages + (size_t)st * kStageFloats;
        const size_t off = (size_t)ray0 * S;
        mbar_arrive_expect_tx(&full_bar[st], sb * ((has_beta ? 8u : 7u) - (derive ? 1u : 0u)));
        bulk_g2s(dst + 0 * kTileFloats, p.density + off, sb, &full_bar[st]);
        if (!derive) bulk_g2s(dst + 1 * kTileFloats, p.deltas + off, sb, &full_bar[st]);
        bulk_g2s(dst + 2 * kTileFloats, p.starts + off, sb, &full_bar[st]);
        bulk_g2s(dst + 3 * kTileFloats, p.ends + off, sb, &full_bar[st]);
        if (has_beta) bulk_g2s(dst + 4 * kTileFloats, p.beta + off, sb, &full_bar[st]);
        bulk_g2s(dst + 5 * kTileFloats, p.rgb + off * 3, sb * 3u, &full_bar[st]);
      }
    }
    return;
  }

  // ===== consumers =====
  const int w = warp;
  const int r = lane >> 2;  // ray within the tile
  const int q = lane & 3;   // quarter of the ray this lane owns
  const int lane_off = r * S + q * P;
  ChunkBounds bounds;
  // the (up to) three outputs this lane stores per ray, by its quarter q -- loop-invariant, selected once:
  //   q == 0: rgb[3 ray + 0, 1, 2]    q == 1: accumulation, depth, expected depth    q == 2: rgb_var, rgb_std    q == 3: depth_var, depth_std
  float* const out0 = q == 0 ? p.o_rgb : (q == 1 ? p.o_acc : (q == 2 ? p.o_rgb_var : p.o_dvar));
  float* const out1 = q == 0 ? (p.o_rgb ? p.o_rgb + 1 : nullptr)
                             : (q == 1 ? p.o_depth : (q == 2 ? p.o_rgb_std : p.o_dstd));
  float* const out2 = q == 0 ? (p.o_rgb ? p.o_rgb + 2 : nullptr) : (q == 1 ? p.o_exp : nullptr);
  const int out_mult = q == 0 ? 3 : 1;

  for (int j = w; j < my_tiles; j += NCW) {
    const int st = j % NST;
    const uint32_t ph = (uint32_t)((j / NST) & 1);
    const int tile = j * (int)gridDim.x + (int)blockIdx.x;
    const int ray0 = tile * kRaysPerTile;
    const int ray = ray0 + r;
    const bool active = ray < (int)p.num_rays;  // the fast path takes num_rays < 2^30

    mbar_wait(&full_bar[st], ph, j);
    const float* sb = stages + (size_t)st * kStageFloats;
#ifdef UB_COMPOSITE_DRY  // measurement aid (UB_NVCC_EXTRA=-DUB_COMPOSITE_DRY): the copy stream alone, consumers only
                         // release their stages -- 259 us per 1 089 480 rays = 6.63 TB/s, the ceiling of this access pattern
    if (sb[lane_off] == 1234.5f && p.o_rgb_var) p.o_rgb_var[ray] = 0.f;
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[st]);
    continue;
#endif

    // ---- optical depth: dd = delta * sigma, float64 exclusive prefix along the ray ----
    float dd[P];
    {
      const float4* a = reinterpret_cast<const float4*>(sb + 0 * kTileFloats + lane_off);
      const float4* b = reinterpret_cast<const float4*>(sb + 1 * kTileFloats + lane_off);
      const float4* c = reinterpret_cast<const float4*>(sb + 2 * kTileFloats + lane_off);
      const float4* d = reinterpret_cast<const float4*>(sb + 3 * kTileFloats + lane_off);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float4 x = a[v];
        float4 y;
        if (derive) {
          const float4 s0 = c[v], s1 = d[v];
          y = make_float4(__fsub_rn(s1.x, s0.x), __fsub_rn(s1.y, s0.y), __fsub_rn(s1.z, s0.z), __fsub_rn(s1.w, s0.w));
        } else {
          y = b[v];
        }
        dd[4 * v + 0] = __fmul_rn(y.x, x.x);
        dd[4 * v + 1] = __fmul_rn(y.y, x.y);
        dd[4 * v + 2] = __fmul_rn(y.z, x.z);
        dd[4 * v + 3] = __fmul_rn(y.w, x.w);
      }
    }
    double pre[P];
    double run = 0.0;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      pre[i] = run;
      run += (double)dd[i];
    }
    double off = ray_exclusive_offset(run, q);

    // ---- weights ----
    float wgt[P];
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      const float trans = expf(-(float)(off + pre[i]));
      const float alpha = 1.0f - expf(-dd[i]);
      wgt[i] = __fmul_rn(alpha, trans);
      acc += wgt[i];
    }
    acc = reduce4(acc);
    if (__any_sync(FULL_MASK, non_finite(acc))) {  // rare: some weight is NaN / inf -> torch.nan_to_num
      acc = 0.f;
#pragma unroll
      for (int i = 0; i < P; ++i) {
        wgt[i] = nan_to_num(wgt[i]);
        acc += wgt[i];
      }
      acc = reduce4(acc);
    }

    // ---- sample midpoints, chunk bounds ----
    float step[P];
    {
      const float4* c = reinterpret_cast<const float4*>(sb + 2 * kTileFloats + lane_off);
      const float4* d = reinterpret_cast<const float4*>(sb + 3 * kTileFloats + lane_off);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float4 x = c[v], y = d[v];
        step[4 * v + 0] = __fadd_rn(x.x, y.x) * 0.5f;
        step[4 * v + 1] = __fadd_rn(x.y, y.y) * 0.5f;
        step[4 * v + 2] = __fadd_rn(x.z, y.z) * 0.5f;
        step[4 * v + 3] = __fadd_rn(x.w, y.w) * 0.5f;
      }
    }
    bounds.enter(p.chunk_shift >= 0 ? (tile >> p.chunk_shift) : (p.tiles_per_chunk > 0 ? tile / p.tiles_per_chunk : 0),
                 p.chunk_ws);
    float rmin = step[0], rmax = step[0];  // step range of the ray (fminf / fmaxf skip NaN, like the chunk bounds)
#pragma unroll
    for (int i = 1; i < P; ++i) {
      rmin = fminf(rmin, step[i]);
      rmax = fmaxf(rmax, step[i]);
    }
    if (active) {
      bounds.add_step(rmin);
      bounds.add_step(rmax);
    }
    rmin = fminf(rmin, __shfl_xor_sync(FULL_MASK, rmin, 1));
    rmax = fmaxf(rmax, __shfl_xor_sync(FULL_MASK, rmax, 1));
    rmin = fminf(rmin, __shfl_xor_sync(FULL_MASK, rmin, 2));
    rmax = fmaxf(rmax, __shfl_xor_sync(FULL_MASK, rmax, 2));

    // ---- cumulative weight in float64, first sample with (float)cw >= 0.5 ----
    run = 0.0;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      run += (double)wgt[i];
      pre[i] = run;
    }
    off = ray_exclusive_offset(run, q);
    int first = S;
    const double half_here = kHalf - off;  // off + pre[i] >= kHalf, with the lane's offset moved to the other side
#pragma unroll
    for (int i = P - 1; i >= 0; --i)
      if (pre[i] >= half_here) first = q * P + i;
    first = min(first, __shfl_xor_sync(FULL_MASK, first, 1));
    first = min(first, __shfl_xor_sync(FULL_MASK, first, 2));
    first = min(first, S - 1);  // clamp(idx, 0, S-1)
    // the midpoint of sample `first`, recomputed from the stage (the 4 lanes of a ray read the same two words) instead of
    // a select chain over the lane's 12 midpoints plus a shuffle
    const float depth = __fadd_rn(sb[2 * kTileFloats + r * S + first], sb[3 * kTileFloats + r * S + first]) * 0.5f;

    // ---- expected depth numerator, depth variance ----
    float e_num = 0.f, dvar = 0.f;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      e_num = fmaf(wgt[i], step[i], e_num);
      const float t = step[i] - depth;
      dvar = fmaf(wgt[i], t * t, dvar);
    }

    // ---- sum w^2 beta ----
    float var = 0.f;
    if (has_beta) {
      const float4* e = reinterpret_cast<const float4*>(sb + 4 * kTileFloats + lane_off);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float4 x = e[v];
        var = fmaf(wgt[4 * v + 0] * wgt[4 * v + 0], x.x, var);
        var = fmaf(wgt[4 * v + 1] * wgt[4 * v + 1], x.y, var);
        var = fmaf(wgt[4 * v + 2] * wgt[4 * v + 2], x.z, var);
        var = fmaf(wgt[4 * v + 3] * wgt[4 * v + 3], x.w, var);
      }
      var = reduce4(var);
      if (p.beta_mode == UB_BETA_NAN_GUARD && __any_sync(FULL_MASK, var != var)) {
        // rare: a NaN sum -- redo with NaN betas zeroed and remember that this chunk had a NaN
        const float* bs = sb + 4 * kTileFloats + lane_off;
        bool saw_nan = false;
        var = 0.f;
#pragma unroll
        for (int i = 0; i < P; ++i) {
          float bi = bs[i];
          if (bi != bi) {
            bi = 0.f;
            saw_nan = true;
          }
          var = fmaf(wgt[i] * wgt[i], bi, var);
        }
        var = reduce4(var);
        if (active && saw_nan) bounds.has_nan = true;
      }
    }

    // ---- colours (read straight from the stage) ----
    float col[3] = {0.f, 0.f, 0.f};
    const float* cs = sb + 5 * kTileFloats + 3 * lane_off;
    {
      const float4* f = reinterpret_cast<const float4*>(cs);
#pragma unroll
      for (int v = 0; v < 3 * V; ++v) {
        const float4 x = f[v];
        col[(4 * v + 0) % 3] = fmaf(wgt[(4 * v + 0) / 3], x.x, col[(4 * v + 0) % 3]);
        col[(4 * v + 1) % 3] = fmaf(wgt[(4 * v + 1) / 3], x.y, col[(4 * v + 1) % 3]);
        col[(4 * v + 2) % 3] = fmaf(wgt[(4 * v + 2) / 3], x.z, col[(4 * v + 2) % 3]);
        col[(4 * v + 3) % 3] = fmaf(wgt[(4 * v + 3) / 3], x.w, col[(4 * v + 3) % 3]);
      }
    }
    float cr = reduce4(col[0]), cg = reduce4(col[1]), cb = reduce4(col[2]);
    const float* last = sb + 5 * kTileFloats + 3 * (r * S + S - 1);  // colour of the ray's last sample
    float b0 = last[0], b1 = last[1], b2 = last[2];
    if (p.eval_mode &&
        __any_sync(FULL_MASK, non_finite(cr) || non_finite(cg) || non_finite(cb))) {
      // rare: a non-finite colour -> eval-mode nan_to_num(rgb) before the weighted sum
      col[0] = col[1] = col[2] = 0.f;
#pragma unroll
      for (int k = 0; k < 3 * P; ++k) col[k % 3] = fmaf(wgt[k / 3], nan_to_num(cs[k]), col[k % 3]);
      cr = reduce4(col[0]);
      cg = reduce4(col[1]);
      cb = reduce4(col[2]);
      b0 = nan_to_num(b0);
      b1 = nan_to_num(b1);
      b2 = nan_to_num(b2);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[st]);  // the stage can be refilled

    e_num = reduce4(e_num);
    dvar = reduce4(dvar) + 1e-5f;
    if (p.bg_mode == UB_BG_FIXED) {
      b0 = p.bg[0];
      b1 = p.bg[1];
      b2 = p.bg[2];
    }
    if (p.bg_mode != UB_BG_NONE) {
      const float rem = 1.0f - acc;
      cr = cr + b0 * rem;
      cg = cg + b1 * rem;
      cb = cb + b2 * rem;
    }
    if (p.eval_mode) {
      cr = clamp01_keep_nan(cr);
      cg = clamp01_keep_nan(cg);
      cb = clamp01_keep_nan(cb);
    }

    // ---- candidates of the finalize pass: the chunk-wide clip can only change an expected depth that lies outside
    // the ray's own [min, max] of steps; the NaN-guard redo only touches non-finite sums ----
    const float e_exp = e_num / (acc + 1e-10f);
    if (p.cand_flags) {
      const unsigned clip_m = __ballot_sync(FULL_MASK, active && q == 0 && e_exp == e_exp && !(e_exp >= rmin && e_exp <= rmax));
      const unsigned var_m = __ballot_sync(FULL_MASK, active && q == 0 && has_beta && p.beta_mode == UB_BETA_NAN_GUARD &&
                                                          non_finite(var));
      if (lane == 0) p.cand_flags[tile] = ((unsigned long long)var_m << 32) | clip_m;   // bit 4 r <-> ray r of the tile
    }

    // ---- outputs: the 4 lanes of a ray split the stores.  Branch-free: a four-way branch on q runs its four bodies one
    // after the other with a quarter of the lanes each (two of them a square root); here every lane fills up to three
    // store slots selected by q, and one square root serves rgb_std (q == 2) and depth_std (q == 3) ----
    if (active) {
      const float sq = sqrtf(q == 2 ? var : dvar);
      const float v0 = q == 0 ? cr : (q == 1 ? acc : (q == 2 ? var : dvar));
      const float v1 = q == 0 ? cg : (q == 1 ? depth : sq);
      const float v2 = q == 0 ? cb : e_exp;
      const size_t at = (size_t)ray * out_mult;
      if (out0) out0[at] = v0;
      if (out1) out1[at] = v1;
      if (out2) out2[at] = v2;
      if (p.o_w) {
        float4* ow = reinterpret_cast<float4*>(p.o_w + (size_t)ray * S + q * P);
#pragma unroll
        for (int v = 0; v < V; ++v)
          ow[v] = make_float4(wgt[4 * v], wgt[4 * v + 1], wgt[4 * v + 2], wgt[4 * v + 3]);
      }
    }
  }
  bounds.flush(p.chunk_ws);
}

// Inclusive float64 warp scan (Kogge-Stone).  The association differs from a sequential sum only
// below float64 rounding; every prefix is rounded to float32 by the caller.
__device__ __forceinline__ double warp_inclusive_scan(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double up = shfl_up_double(FULL_MASK, v, o);
    if (lane >= o) v += up;
  }
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}

// Generic path: one warp per ray, any S, any alignment, optional render-from-weights mode.
template <bool FROM_WEIGHTS>
__global__ void __launch_bounds__(256) composite_rays_generic(const CompositeParams p) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long num_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int S = p.num_samples;
  ChunkBounds bounds;
  // each warp owns a contiguous range of rays, so it crosses at most a few chunk boundaries
  const long long per_warp = (p.num_rays + num_warps - 1) / num_warps;
  const long long ray_begin = warp_global * per_warp;
  const long long ray_end = min(p.num_rays, ray_begin + per_warp);

  for (long long ray = ray_begin; ray < ray_end; ++ray) {
    const size_t row = (size_t)ray * S;
    const long long chunk = p.rays_per_chunk > 0 ? ray / p.rays_per_chunk : 0;
    bounds.enter(chunk, p.chunk_ws);

    float acc = 0.f, e_num = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, var = 0.f;
    double carry_dd = 0.0, carry_w = 0.0;
    int first = S;
    bool nan_beta = false;
    float last0 = 0.f, last1 = 0.f, last2 = 0.f;

    // pass 1: weights, sums, median index
    for (int base = 0; base < S; base += 32) {
      const int i = base + lane;
      const bool ok = i < S;
      float wi = 0.f, st = 0.f;
      if (ok) st = (p.starts[row + i] + p.ends[row + i]) / 2;
      if (FROM_WEIGHTS) {
        if (ok) wi = p.weights_in[row + i];
      } else {
        float ddi = ok ? delta_at(p, row + i) * p.density[row + i] : 0.f;
        double incl = warp_inclusive_scan((double)ddi, lane);
        // exclusive prefix = carry + inclusive prefix of the previous lane
        double prev = shfl_up_double(FULL_MASK, incl, 1);
        double excl = lane == 0 ? carry_dd : carry_dd + prev;
        carry_dd += shfl_double(FULL_MASK, incl, 31);
        if (ok) {
          const float trans = expf(-(float)excl);
          const float alpha = 1.0f - expf(-ddi);
          wi = nan_to_num(alpha * trans);
        }
      }
      double cwi = carry_w + warp_inclusive_scan((double)wi, lane);
      carry_w = shfl_double(FULL_MASK, cwi, 31);
      const unsigned hit = __ballot_sync(FULL_MASK, ok && (float)cwi >= 0.5f);
      if (first == S && hit) first = base + __ffs(hit) - 1;
      if (ok) {
        bounds.add_step(st);
        acc += wi;
        e_num += wi * st;
        if (!FROM_WEIGHTS) {
          float c0 = p.rgb[(row + i) * 3 + 0], c1 = p.rgb[(row + i) * 3 + 1],
                c2 = p.rgb[(row + i) * 3 + 2];
          if (p.eval_mode) {
            c0 = nan_to_num(c0);
            c1 = nan_to_num(c1);
            c2 = nan_to_num(c2);
          }
          cr += wi * c0;
          cg += wi * c1;
          cb += wi * c2;
          if (i == S - 1) {
            last0 = c0;
            last1 = c1;
            last2 = c2;
          }
          if (p.beta) {
            float bi = p.beta[row + i];
            if (p.beta_mode == UB_BETA_NAN_GUARD && bi != bi) {
              bi = 0.0f;
              nan_beta = true;
            }
            var += (wi * wi) * bi;
          }
          if (p.o_w) p.o_w[row + i] = wi;
        }
      }
    }
    first = min(first, S - 1);
    const float depth = (p.starts[row + first] + p.ends[row + first]) / 2;

    // pass 2: depth variance around the median (weights recomputed identically)
    float dvar = 0.f;
    if (p.o_dvar || p.o_dstd) {
      carry_dd = 0.0;
      for (int base = 0; base < S; base += 32) {
        const int i = base + lane;
        const bool ok = i < S;
        float wi = 0.f;
        if (FROM_WEIGHTS) {
          if (ok) wi = p.weights_in[row + i];
        } else {
          float ddi = ok ? delta_at(p, row + i) * p.density[row + i] : 0.f;
          double incl = warp_inclusive_scan((double)ddi, lane);
          double prev = shfl_up_double(FULL_MASK, incl, 1);
          double excl = lane == 0 ? carry_dd : carry_dd + prev;
          carry_dd += shfl_double(FULL_MASK, incl, 31);
          if (ok) wi = nan_to_num((1.0f - expf(-ddi)) * expf(-(float)excl));
        }
        if (ok) {
          const float t = (p.starts[row + i] + p.ends[row + i]) / 2 - depth;
          dvar += wi * (t * t);
        }
      }
    }

    acc = warp_sum(acc);
    e_num = warp_sum(e_num);
    dvar = warp_sum(dvar) + 1e-5f;
    bounds.has_nan |= nan_beta;
    if (!FROM_WEIGHTS) {
      cr = warp_sum(cr);
      cg = warp_sum(cg);
      cb = warp_sum(cb);
      var = warp_sum(var);
      const int src = (S - 1) & 31;
      float b0 = __shfl_sync(FULL_MASK, last0, src);
      float b1 = __shfl_sync(FULL_MASK, last1, src);
      float b2 = __shfl_sync(FULL_MASK, last2, src);
      if (p.bg_mode == UB_BG_FIXED) {
        b0 = p.bg[0];
        b1 = p.bg[1];
        b2 = p.bg[2];
      }
      if (p.bg_mode != UB_BG_NONE) {
        const float rem = 1.0f - acc;
        cr = cr + b0 * rem;
        cg = cg + b1 * rem;
        cb = cb + b2 * rem;
      }
      if (p.eval_mode) {
        cr = clamp01_keep_nan(cr);
        cg = clamp01_keep_nan(cg);
        cb = clamp01_keep_nan(cb);
      }
    }
    if (lane == 0) {
      if (!FROM_WEIGHTS) {
        if (p.o_rgb) {
          p.o_rgb[ray * 3 + 0] = cr;
          p.o_rgb[ray * 3 + 1] = cg;
          p.o_rgb[ray * 3 + 2] = cb;
        }
        if (p.o_rgb_var) p.o_rgb_var[ray] = var;
        if (p.o_rgb_std) p.o_rgb_std[ray] = sqrtf(var);
      }
      if (p.o_acc) p.o_acc[ray] = acc;
      if (p.o_depth) p.o_depth[ray] = depth;
      if (p.o_exp) p.o_exp[ray] = e_num / (acc + 1e-10f);
      if (p.o_dvar) p.o_dvar[ray] = dvar;
      if (p.o_dstd) p.o_dstd[ray] = sqrtf(dvar);
    }
  }
  bounds.flush(p.chunk_ws);
}

// Finalize: clip the expected depth to the chunk's [min(steps), max(steps)] (torch.clip keeps a
// NaN input); and, in chunks whose beta contained a NaN, redo rgb_var for the (rare) rays whose
// beta also holds +-inf, because the reference then applied nan_to_num to the whole chunk.
__device__ __forceinline__ void finalize_ray(const CompositeParams& p, long long ray) {
  const long long chunk = p.rays_per_chunk > 0 ? ray / p.rays_per_chunk : 0;
  const unsigned* ws = p.chunk_ws + chunk * 4;
  if (p.o_exp) {
    const float lo = order_key_inv(~ws[1]);
    const float hi = order_key_inv(ws[0]);
    float e = p.o_exp[ray];
    if (e == e) e = fminf(fmaxf(e, lo), hi);
    p.o_exp[ray] = e;
  }
  if (p.beta && p.beta_mode == UB_BETA_NAN_GUARD && ws[2] && p.o_rgb_var) {
    const float v = p.o_rgb_var[ray];
    if (!isfinite(v)) {
      const int S = p.num_samples;
      const size_t row = (size_t)ray * S;
      double carry = 0.0;
      float var = 0.f;
      for (int i = 0; i < S; ++i) {
        const float ddi = delta_at(p, row + i) * p.density[row + i];
        const float wi = nan_to_num((1.0f - expf(-ddi)) * expf(-(float)carry));
        carry += (double)ddi;
        var += (wi * wi) * nan_to_num(p.beta[row + i]);
      }
      p.o_rgb_var[ray] = var;
      if (p.o_rgb_std) p.o_rgb_std[ray] = sqrtf(var);
    }
  }
}

// visits the flagged rays when the batch has candidate flags (one thread per tile, grid-stride), else every ray
__device__ __forceinline__ void finalize_batch_rays(const CompositeParams& p) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p.cand_flags) {
    const long long tiles = (p.num_rays + kRaysPerTile - 1) / kRaysPerTile;
    for (long long tile = t; tile < tiles; tile += stride) {
      const unsigned long long f = p.cand_flags[tile];
      unsigned m = (unsigned)f | (unsigned)(f >> 32);
      while (m) {
        const int bit = __ffs(m) - 1;
        m &= m - 1;
        finalize_ray(p, tile * kRaysPerTile + (bit >> 2));
      }
    }
  } else {
    for (long long ray = t; ray < p.num_rays; ray += stride) finalize_ray(p, ray);
  }
}

__global__ void __launch_bounds__(256) composite_finalize(const CompositeParams p) { finalize_batch_rays(p); }

// the finalize passes of up to UB_MAX_COMPOSITE_BATCH independent ray batches in one launch (blockIdx.y = batch)
struct FinalizeBatch {
  CompositeParams p[UB_MAX_COMPOSITE_BATCH];
};
__global__ void __launch_bounds__(256) composite_finalize_batch(const __grid_constant__ FinalizeBatch b) {
  finalize_batch_rays(b.p[blockIdx.y]);
}


static size_t chunk_ws_bytes(long long num_rays, long long rays_per_chunk) {
  long long chunks = 1;
  if (rays_per_chunk > 0 && num_rays > 0) chunks = (num_rays + rays_per_chunk - 1) / rays_per_chunk;
  return (size_t)chunks * 4 * sizeof(unsigned);
}

template <int S, int NCW, int NST>
static int launch_tma(const CompositeParams& p, cudaStream_t stream) {
  constexpr size_t stage_bytes = (size_t)kRaysPerTile * S * 8 * 4;
  constexpr size_t smem = NST * stage_bytes + 2 * NST * sizeof(uint64_t);
  static_assert(smem <= 227 * 1024, "stage ring exceeds shared memory");
  auto kern = composite_rays_tma<S, NCW, NST>;
  static bool configured[64] = {};  // per device: the attribute call costs a few microseconds per launch
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("composite_rays: cannot reserve %zu B shared memory (%s)", smem, cudaGetErrorString(e));
      return UB_ERR_LAUNCH;
    }
    // Ask for the largest shared-memory carveout, not the smallest that holds the ring: the persistent CTA stays on
    // its SM for the whole view, and the ~55 KB beyond its ring is what lets a block of the previous view's scoring
    // kernels (up to 47 KB each) run beside it instead of waiting for the gap between two compositing kernels.
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const long long tiles = (p.num_rays + kRaysPerTile - 1) / kRaysPerTile;
  int grid = sm_count();
  if (grid <= 0) grid = 148;
  if (tiles < grid) grid = (int)tiles;
  kern<<<grid, 32 * (1 + NCW), smem, stream>>>(p);
  return check_launch("composite_rays_tma");
}

static bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; }

}  // namespace ub

extern "C" {

// workspace of one ray batch: [chunk table (zeroed by the call) | candidate flags, 8 B per 8-ray tile]
static size_t composite_head_bytes(int64_t num_rays, int64_t rays_per_chunk) {
  return ub::align_up(ub::chunk_ws_bytes(num_rays, rays_per_chunk), 16);
}
static size_t composite_list_bytes(int64_t num_rays) {
  return (size_t)((num_rays > 0 ? num_rays : 0) + ub::kRaysPerTile - 1) / ub::kRaysPerTile * 8;
}

size_t ub_composite_rays_workspace_bytes(int64_t num_rays, int64_t rays_per_chunk) {
  return composite_head_bytes(num_rays, rays_per_chunk) + composite_list_bytes(num_rays);
}
size_t ub_render_weights_workspace_bytes(int64_t num_rays, int64_t rays_per_chunk) {
  return ub::chunk_ws_bytes(num_rays, rays_per_chunk);
}

static bool composite_fast_path(const ub_composite_rays_args* a) {
  using namespace ub;
  const bool chunk_ok = (a->rays_per_chunk <= 0 || a->rays_per_chunk % kRaysPerTile == 0) &&
                        a->num_rays < (1LL << 30);
  const bool align_ok = aligned16(a->density) && aligned16(a->deltas) && aligned16(a->starts) &&
                        aligned16(a->ends) && aligned16(a->rgb) &&
                        (a->beta == nullptr || aligned16(a->beta)) &&
                        (a->out_weights == nullptr || aligned16(a->out_weights));
  const int s = a->num_samples;
  return chunk_ok && align_ok && (s == 32 || s == 48 || s == 64 || s == 96);
}

// validation + parameter block of one ray batch (no launches)
static int composite_prepare(const ub_composite_rays_args* a, void* workspace, size_t workspace_bytes,
                             void* cand_list, ub::CompositeParams& p) {
  using namespace ub;
  UB_REQUIRE(a != nullptr, UB_ERR_BAD_ARG, "composite_rays: args is NULL");
  UB_REQUIRE(a->num_rays >= 0 && a->num_samples >= 1, UB_ERR_BAD_ARG,
             "composite_rays: bad shape R=%lld S=%d", (long long)a->num_rays, a->num_samples);
  if (a->num_rays == 0) return UB_OK;
  UB_REQUIRE(a->density && a->starts && a->ends && a->rgb, UB_ERR_BAD_ARG,
             "composite_rays: density/starts/ends/rgb must be non-NULL");  // deltas == NULL: ends - starts
  UB_REQUIRE(a->background_mode >= UB_BG_LAST_SAMPLE && a->background_mode <= UB_BG_FIXED,
             UB_ERR_BAD_ARG, "composite_rays: bad background_mode %d", a->background_mode);
  UB_REQUIRE(a->beta_mode == UB_BETA_RAW || a->beta_mode == UB_BETA_NAN_GUARD, UB_ERR_BAD_ARG,
             "composite_rays: bad beta_mode %d", a->beta_mode);
  UB_REQUIRE((a->out_rgb_var == nullptr && a->out_rgb_std == nullptr) || a->beta != nullptr,
             UB_ERR_BAD_ARG, "composite_rays: rgb_var/rgb_std requested without beta");
  const size_t need = composite_head_bytes(a->num_rays, a->rays_per_chunk);
  UB_REQUIRE(workspace != nullptr && workspace_bytes >= need, UB_ERR_WORKSPACE,
             "composite_rays: workspace %zu B < required %zu B", workspace_bytes, need);
  p = CompositeParams{};
  p.density = a->density;
  p.deltas = a->deltas;
  p.starts = a->starts;
  p.ends = a->ends;
  p.rgb = a->rgb;
  p.beta = a->beta;
  p.weights_in = nullptr;
  p.num_rays = a->num_rays;
  p.num_samples = a->num_samples;
  p.bg_mode = a->background_mode;
  p.bg[0] = a->background_rgb[0];
  p.bg[1] = a->background_rgb[1];
  p.bg[2] = a->background_rgb[2];
  p.beta_mode = a->beta_mode;
  p.rays_per_chunk = a->rays_per_chunk;
  p.eval_mode = a->eval_mode;
  p.o_rgb = a->out_rgb;
  p.o_acc = a->out_accumulation;
  p.o_depth = a->out_depth;
  p.o_exp = a->out_expected_depth;
  p.o_rgb_var = a->out_rgb_var;
  p.o_rgb_std = a->out_rgb_std;
  p.o_dvar = a->out_depth_var;
  p.o_dstd = a->out_depth_std;
  p.o_w = a->out_weights;
  p.chunk_ws = static_cast<unsigned*>(workspace);
  p.tiles_per_chunk = a->rays_per_chunk > 0 ? (int)(a->rays_per_chunk / kRaysPerTile) : 0;
  p.chunk_shift = -1;
  if (p.tiles_per_chunk > 0 && (p.tiles_per_chunk & (p.tiles_per_chunk - 1)) == 0) {
    p.chunk_shift = 0;
    while ((1 << p.chunk_shift) < p.tiles_per_chunk) ++p.chunk_shift;
  }
  if (cand_list != nullptr && composite_fast_path(a))  // the generic kernel keeps the visit-every-ray finalize
    p.cand_flags = static_cast<unsigned long long*>(cand_list);
  return UB_OK;
}

// the compositing kernel of one prepared batch (the chunk workspace must already be zero)
static int composite_launch_main(const ub_composite_rays_args* a, const ub::CompositeParams& p, cudaStream_t stream) {
  using namespace ub;
  int rc = UB_OK;
  bool fast = composite_fast_path(a);
  if (fast) {
    switch (a->num_samples) {
      case 32: rc = launch_tma<32, 8, 16>(p, stream); break;
      case 48: {
        // tuning hook: consumer warps per CTA / ring stages (default 7 / 14, the fastest measured on B200)
        static const int ncw = [] { const char* e = getenv("UB_COMPOSITE_NCW"); return e ? atoi(e) : 7; }();
        static const int nst = [] { const char* e = getenv("UB_COMPOSITE_NST"); return e ? atoi(e) : 14; }();
#define UB_TRY(N, T) if (ncw == N && nst == T) { rc = launch_tma<48, N, T>(p, stream); break; }
        UB_TRY(4, 16) UB_TRY(8, 16) UB_TRY(16, 16) UB_TRY(6, 18) UB_TRY(9, 18) UB_TRY(6, 12) UB_TRY(12, 12)
        UB_TRY(14, 14) UB_TRY(10, 10) UB_TRY(7, 14) UB_TRY(5, 15) UB_TRY(8, 8)
#undef UB_TRY
        rc = launch_tma<48, 7, 14>(p, stream);
        break;
      }
      case 64: rc = launch_tma<64, 8, 8>(p, stream); break;
      case 96: rc = launch_tma<96, 4, 8>(p, stream); break;
      default: fast = false;
    }
  }
  if (!fast) {
    const long long warps_needed = a->num_rays;
    long long blocks = (warps_needed + 7) / 8;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap && cap > 0) blocks = cap;
    composite_rays_generic<false><<<(unsigned)blocks, 256, 0, stream>>>(p);
    rc = check_launch("composite_rays_generic");
  }
  return rc;
}

int ub_composite_rays(const ub_composite_rays_args* a, void* workspace, size_t workspace_bytes,
                      void* stream_v) {
  using namespace ub;
  CompositeParams p{};
  // a workspace that only holds the chunk table (the size this call needed before the candidate list existed) still
  // works: finalize then visits every ray
  const bool has_shape = a != nullptr && a->num_rays > 0;
  const size_t head = has_shape ? composite_head_bytes(a->num_rays, a->rays_per_chunk) : 0;
  const bool room = has_shape && workspace != nullptr &&
                    workspace_bytes >= ub_composite_rays_workspace_bytes(a->num_rays, a->rays_per_chunk);
  int rc = composite_prepare(a, workspace, workspace_bytes, room ? static_cast<char*>(workspace) + head : nullptr, p);
  if (rc != UB_OK || a->num_rays == 0) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (cudaMemsetAsync(workspace, 0, head, stream) != cudaSuccess) return check_launch("composite_rays memset");
  rc = composite_launch_main(a, p, stream);
  if (rc != UB_OK) return rc;
  const long long fwork = p.cand_flags ? (a->num_rays + kRaysPerTile - 1) / kRaysPerTile : a->num_rays;
  const unsigned fblocks = (unsigned)min((long long)65535, (long long)((fwork + 255) / 256));
  composite_finalize<<<fblocks, 256, 0, stream>>>(p);
  return check_launch("composite_finalize");
}

// batch workspace: the heads (chunk table + candidate counter) of all batches first -- one memset --, then the lists
static size_t batch_heads_bytes(const ub_composite_rays_args* args, int32_t num_batches) {
  size_t total = 0;
  for (int i = 0; i < num_batches; ++i) total += composite_head_bytes(args[i].num_rays, args[i].rays_per_chunk);
  return total;
}

size_t ub_composite_rays_batch_workspace_bytes(const ub_composite_rays_args* args, int32_t num_batches) {
  if (args == nullptr) return 16;
  size_t total = batch_heads_bytes(args, num_batches);
  for (int i = 0; i < num_batches; ++i) total += composite_list_bytes(args[i].num_rays);
  return total > 0 ? total : 16;
}

int ub_composite_rays_batch(const ub_composite_rays_args* args, int32_t num_batches, void* workspace,
                            size_t workspace_bytes, void* stream_v) {
  using namespace ub;
  UB_REQUIRE(args != nullptr && num_batches >= 1 && num_batches <= UB_MAX_COMPOSITE_BATCH, UB_ERR_BAD_ARG,
             "composite_rays_batch: between 1 and %d batches", UB_MAX_COMPOSITE_BATCH);
  const size_t need = ub_composite_rays_batch_workspace_bytes(args, num_batches);
  UB_REQUIRE(workspace != nullptr && workspace_bytes >= need, UB_ERR_WORKSPACE,
             "composite_rays_batch: workspace %zu B < required %zu B", workspace_bytes, need);
  UB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0, UB_ERR_BAD_ARG,
             "composite_rays_batch: workspace must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  FinalizeBatch fb{};
  char* ws = static_cast<char*>(workspace);
  const size_t heads = batch_heads_bytes(args, num_batches);
  char* lists = ws + heads;
  long long full_rays = 0;  // finalize work items of the largest batch: tiles (candidate flags) or rays (generic path)
  int live = 0;
  const ub_composite_rays_args* live_args[UB_MAX_COMPOSITE_BATCH];
  for (int i = 0; i < num_batches; ++i) {
    const size_t head = composite_head_bytes(args[i].num_rays, args[i].rays_per_chunk);
    CompositeParams p{};
    int rc = composite_prepare(&args[i], ws, head, lists, p);
    ws += head;
    lists += composite_list_bytes(args[i].num_rays);
    if (rc != UB_OK) return rc;
    if (args[i].num_rays == 0) continue;
    fb.p[live] = p;
    live_args[live] = &args[i];
    ++live;
    const long long work = p.cand_flags ? (args[i].num_rays + kRaysPerTile - 1) / kRaysPerTile : args[i].num_rays;
    if (work > full_rays) full_rays = work;
  }
  if (live == 0) return UB_OK;
  // one memset (heads only), the compositing kernels back to back, one finalize launch for all batches
  if (cudaMemsetAsync(workspace, 0, heads, stream) != cudaSuccess) return check_launch("composite_rays_batch memset");
  for (int i = 0; i < live; ++i) {
    int rc = composite_launch_main(live_args[i], fb.p[i], stream);
    if (rc != UB_OK) return rc;
  }
  const unsigned gx = (unsigned)min((long long)65535, (long long)((full_rays + 255) / 256));
  composite_finalize_batch<<<dim3(gx, (unsigned)live), 256, 0, stream>>>(fb);
  return check_launch("composite_finalize_batch");
}

int ub_render_weights(const ub_render_weights_args* a, void* workspace, size_t workspace_bytes,
                      void* stream_v) {
  using namespace ub;
  UB_REQUIRE(a != nullptr, UB_ERR_BAD_ARG, "render_weights: args is NULL");
  UB_REQUIRE(a->num_rays >= 0 && a->num_samples >= 1, UB_ERR_BAD_ARG,
             "render_weights: bad shape R=%lld S=%d", (long long)a->num_rays, a->num_samples);
  if (a->num_rays == 0) return UB_OK;
  UB_REQUIRE(a->weights && a->starts && a->ends, UB_ERR_BAD_ARG,
             "render_weights: weights/starts/ends must be non-NULL");
  const size_t need = chunk_ws_bytes(a->num_rays, a->rays_per_chunk);
  UB_REQUIRE(workspace != nullptr && workspace_bytes >= need, UB_ERR_WORKSPACE,
             "render_weights: workspace %zu B < required %zu B", workspace_bytes, need);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  CompositeParams p{};
  p.weights_in = a->weights;
  p.starts = a->starts;
  p.ends = a->ends;
  p.num_rays = a->num_rays;
  p.num_samples = a->num_samples;
  p.bg_mode = UB_BG_NONE;
  p.rays_per_chunk = a->rays_per_chunk;
  p.o_acc = a->out_accumulation;
  p.o_depth = a->out_depth;
  p.o_exp = a->out_expected_depth;
  p.o_dvar = a->out_depth_var;
  p.o_dstd = a->out_depth_std;
  p.chunk_ws = static_cast<unsigned*>(workspace);
  if (cudaMemsetAsync(workspace, 0, need, stream) != cudaSuccess) return check_launch("render_weights memset");
  long long blocks = (a->num_rays + 7) / 8;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap && cap > 0) blocks = cap;
  composite_rays_generic<true><<<(unsigned)blocks, 256, 0, stream>>>(p);
  int rc = check_launch("render_weights");
  if (rc != UB_OK) return rc;
  if (a->out_expected_depth) {
    const unsigned fblocks = (unsigned)min((long long)65535, (long long)((a->num_rays + 255) / 256));
    composite_finalize<<<fblocks, 256, 0, stream>>>(p);
    rc = check_launch("render_weights finalize");
  }
  return rc;
}

}  // extern "C"
