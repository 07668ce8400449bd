// (A1) Per-ray front-to-back compositing with the variance term, one pass.
//
// Reference arithmetic (file:line under /root/reference/nerfuncertainty):
//   weights        models/laplace/laplace_model.py:47-62   (== nerfstudio RaySamples.get_weights)
//   rgb/depth/...  models/activenerfacto/activenerfacto_model.py:98-112
//   Sum w^2 beta   models/activenerfacto/activenerfacto_model.py:105-107, laplace_model.py:478-480
//
// Data layout in HBM: six streams of [R,S] float32 rows (rgb: [R,S,3]); a tile of 8 consecutive
// rays is contiguous in every stream.  Fast path (composite_rays_tma<S>):
//   * persistent CTAs, one per SM: warp 0 = producer, warps 1..8 = consumers;
//   * the producer's elected lane stages 8-ray tiles into shared memory with 1-D bulk async
//     copies (cp.async.bulk -> UBLKCP), two 12 KB stages per consumer warp, full/empty mbarriers;
//   * a consumer warp owns a tile: 4 lanes per ray, 12 consecutive samples per lane (S = 48),
//     conflict-free LDS.128 reads, then everything from registers;
//   * the two prefix scans (optical depth, cumulative weight) run in float64 *sequentially* along
//     the ray (lane after lane inside the 4-lane group), rounding every prefix to float32 -- the
//     exact semantics of torch.cumsum on CPU float32, which decides the median-depth index.
// Generic path (composite_rays_generic): warp per ray, any S / alignment, same outputs.
//
// Chunk-wide reductions of the reference (clip bounds of expected depth = min/max of steps over
// the eval chunk; the `isnan(beta).any()` guard) are accumulated per chunk in the workspace and
// applied by a small finalize kernel.
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
};

constexpr int kStagesPerWarp = 2;
// consumer warps per CTA: as many as fit two 8-ray stages each in 227 KB of shared memory
constexpr int consumer_warps_for(int S) { return S <= 48 ? 8 : (S <= 64 ? 6 : 4); }
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

// Per-warp running min/max of `steps` for the current chunk, flushed with two atomics when the
// chunk changes (identity of both atomicMax targets is 0, so the workspace is memset to 0).
struct ChunkBounds {
  long long chunk = -1;
  unsigned kmax = 0u, kmin_inv = 0u;
  bool has_nan = false;
  __device__ __forceinline__ void flush(unsigned* ws) {
    if (chunk >= 0) {
      unsigned a = kmax, b = kmin_inv;
      int n = has_nan ? 1 : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a = max(a, __shfl_xor_sync(FULL_MASK, a, o));
        b = max(b, __shfl_xor_sync(FULL_MASK, b, o));
        n |= __shfl_xor_sync(FULL_MASK, n, o);
      }
      if ((threadIdx.x & 31) == 0) {
        atomicMax(ws + chunk * 4 + 0, a);
        atomicMax(ws + chunk * 4 + 1, b);
        if (n) ws[chunk * 4 + 2] = 1u;
      }
    }
    kmax = 0u;
    kmin_inv = 0u;
    has_nan = false;
  }
  __device__ __forceinline__ void enter(long long c, unsigned* ws) {
    if (c != chunk) {  // warp-uniform
      flush(ws);
      chunk = c;
    }
  }
  __device__ __forceinline__ void add_step(float s) {
    unsigned k = order_key(s);
    kmax = max(kmax, k);
    kmin_inv = max(kmin_inv, ~k);
  }
};

template <int S>
__global__ void __launch_bounds__(32 * (1 + consumer_warps_for(S)), 1)
composite_rays_tma(const CompositeParams p) {
  constexpr int kConsumerWarps = consumer_warps_for(S);
  constexpr int P = S / kLanesPerRay;  // samples per lane
  constexpr int V = P / 4;             // float4 per lane per scalar stream
  static_assert(S % 16 == 0, "fast path needs S % 16 == 0");
  constexpr int kTileFloats = kRaysPerTile * S;      // one scalar stream of one tile
  constexpr int kStageFloats = kTileFloats * 8;      // 5 scalar streams + rgb (3x)
  constexpr uint32_t kStageBytes = kStageFloats * 4;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stages = reinterpret_cast<float*>(smem_raw);
  uint64_t* full_bar =
      reinterpret_cast<uint64_t*>(smem_raw + (size_t)kConsumerWarps * kStagesPerWarp * kStageBytes);
  uint64_t* empty_bar = full_bar + kConsumerWarps * kStagesPerWarp;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long num_tiles = (p.num_rays + kRaysPerTile - 1) / kRaysPerTile;
  const bool has_beta = p.beta != nullptr;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kConsumerWarps * kStagesPerWarp; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == 0) {
    // ===== producer: one elected lane issues all bulk copies =====
    if (lane == 0) {
      for (long long k = 0;; ++k) {
        const long long base_tile = (k * gridDim.x + blockIdx.x) * kConsumerWarps;
        if (base_tile >= num_tiles) break;
        const int s = (int)(k % kStagesPerWarp);
        const uint32_t ph = (uint32_t)((k / kStagesPerWarp) & 1);
        for (int w = 0; w < kConsumerWarps; ++w) {
          const long long tile = base_tile + w;
          if (tile >= num_tiles) break;
          const int slot = w * kStagesPerWarp + s;
          mbar_wait(&empty_bar[slot], ph ^ 1u);
          const long long ray0 = tile * kRaysPerTile;
          const int n = (int)min((long long)kRaysPerTile, p.num_rays - ray0);
          const uint32_t sb = (uint32_t)n * S * 4u;  // bytes of one scalar stream
          float* dst = stages + (size_t)slot * kStageFloats;
          const size_t off = (size_t)ray0 * S;
          mbar_arrive_expect_tx(&full_bar[slot], sb * (has_beta ? 8u : 7u));
          bulk_g2s(dst + 0 * kTileFloats, p.density + off, sb, &full_bar[slot]);
          bulk_g2s(dst + 1 * kTileFloats, p.deltas + off, sb, &full_bar[slot]);
          bulk_g2s(dst + 2 * kTileFloats, p.starts + off, sb, &full_bar[slot]);
          bulk_g2s(dst + 3 * kTileFloats, p.ends + off, sb, &full_bar[slot]);
          if (has_beta) bulk_g2s(dst + 4 * kTileFloats, p.beta + off, sb, &full_bar[slot]);
          bulk_g2s(dst + 5 * kTileFloats, p.rgb + off * 3, sb * 3u, &full_bar[slot]);
        }
      }
    }
    return;
  }

  // ===== consumers =====
  const int w = warp - 1;
  const int r = lane >> 2;  // ray within the tile
  const int q = lane & 3;   // quarter of the ray this lane owns
  const int group_base = lane & ~3;
  ChunkBounds bounds;

  for (long long k = 0;; ++k) {
    const long long tile = (k * gridDim.x + blockIdx.x) * kConsumerWarps + w;
    if (tile >= num_tiles) break;
    const int s = (int)(k % kStagesPerWarp);
    const uint32_t ph = (uint32_t)((k / kStagesPerWarp) & 1);
    const int slot = w * kStagesPerWarp + s;
    const long long ray0 = tile * kRaysPerTile;
    const int n = (int)min((long long)kRaysPerTile, p.num_rays - ray0);
    const bool active = r < n;
    const long long ray = ray0 + r;

    mbar_wait(&full_bar[slot], ph);
    const float* st = stages + (size_t)slot * kStageFloats;
    const int lane_off = r * S + q * P;

    float dd[P], step[P], beta[P], col[3 * P];
    {
      const float4* a = reinterpret_cast<const float4*>(st + 0 * kTileFloats + lane_off);
      const float4* b = reinterpret_cast<const float4*>(st + 1 * kTileFloats + lane_off);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float4 x = a[j], y = b[j];
        dd[4 * j + 0] = y.x * x.x;
        dd[4 * j + 1] = y.y * x.y;
        dd[4 * j + 2] = y.z * x.z;
        dd[4 * j + 3] = y.w * x.w;
      }
      const float4* c = reinterpret_cast<const float4*>(st + 2 * kTileFloats + lane_off);
      const float4* d = reinterpret_cast<const float4*>(st + 3 * kTileFloats + lane_off);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float4 x = c[j], y = d[j];
        step[4 * j + 0] = (x.x + y.x) / 2;
        step[4 * j + 1] = (x.y + y.y) / 2;
        step[4 * j + 2] = (x.z + y.z) / 2;
        step[4 * j + 3] = (x.w + y.w) / 2;
      }
      if (has_beta) {
        const float4* e = reinterpret_cast<const float4*>(st + 4 * kTileFloats + lane_off);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          float4 x = e[j];
          beta[4 * j + 0] = x.x;
          beta[4 * j + 1] = x.y;
          beta[4 * j + 2] = x.z;
          beta[4 * j + 3] = x.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < P; ++i) beta[i] = 0.0f;
      }
      const float4* f = reinterpret_cast<const float4*>(st + 5 * kTileFloats + 3 * lane_off);
#pragma unroll
      for (int j = 0; j < 3 * V; ++j) {
        float4 x = f[j];
        col[4 * j + 0] = x.x;
        col[4 * j + 1] = x.y;
        col[4 * j + 2] = x.z;
        col[4 * j + 3] = x.w;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[slot]);  // stage can be refilled while we compute

    // ---- scan 1: exclusive prefix of dd in float64, sequential along the ray ----
    double dd64[P];
#pragma unroll
    for (int i = 0; i < P; ++i) dd64[i] = (double)dd[i];
    double pre[P];
    double carry = 0.0;
#pragma unroll 1
    for (int qq = 0; qq < kLanesPerRay; ++qq) {
      if (q == qq) {
#pragma unroll
        for (int i = 0; i < P; ++i) {
          pre[i] = carry;
          carry += dd64[i];
        }
      }
      carry = shfl_double(FULL_MASK, carry, group_base | qq);
    }

    float wgt[P];
    float acc = 0.f, e_num = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, var = 0.f;
    bool nan_beta = false;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      const float trans = expf(-(float)pre[i]);
      const float alpha = 1.0f - expf(-dd[i]);
      const float wi = nan_to_num(alpha * trans);
      wgt[i] = wi;
      acc += wi;
      e_num += wi * step[i];
      float c0 = col[3 * i + 0], c1 = col[3 * i + 1], c2 = col[3 * i + 2];
      if (p.eval_mode) {
        c0 = nan_to_num(c0);
        c1 = nan_to_num(c1);
        c2 = nan_to_num(c2);
        col[3 * i + 0] = c0;
        col[3 * i + 1] = c1;
        col[3 * i + 2] = c2;
      }
      cr += wi * c0;
      cg += wi * c1;
      cb += wi * c2;
      float bi = beta[i];
      if (p.beta_mode == UB_BETA_NAN_GUARD && bi != bi) {
        bi = 0.0f;
        nan_beta = true;
      }
      var += (wi * wi) * bi;
    }

    // ---- scan 2: inclusive prefix of the weights in float64, sequential; median index ----
    double w64[P];
#pragma unroll
    for (int i = 0; i < P; ++i) w64[i] = (double)wgt[i];
    double cw[P];
    carry = 0.0;
#pragma unroll 1
    for (int qq = 0; qq < kLanesPerRay; ++qq) {
      if (q == qq) {
#pragma unroll
        for (int i = 0; i < P; ++i) {
          carry += w64[i];
          cw[i] = carry;
        }
      }
      carry = shfl_double(FULL_MASK, carry, group_base | qq);
    }
    int first = S;  // first sample index with cumulative weight >= 0.5 (searchsorted side=left)
#pragma unroll
    for (int i = P - 1; i >= 0; --i)
      if ((float)cw[i] >= 0.5f) first = q * P + i;
    first = min(first, __shfl_xor_sync(FULL_MASK, first, 1));
    first = min(first, __shfl_xor_sync(FULL_MASK, first, 2));
    first = min(first, S - 1);  // clamp(idx, 0, S-1)
    float depth = step[0];
#pragma unroll
    for (int i = 1; i < P; ++i)
      if (i == first % P) depth = step[i];
    depth = __shfl_sync(FULL_MASK, depth, group_base | (first / P));

    float dvar = 0.f;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      const float t = step[i] - depth;
      dvar += wgt[i] * (t * t);
    }

    acc = reduce4(acc);
    e_num = reduce4(e_num);
    cr = reduce4(cr);
    cg = reduce4(cg);
    cb = reduce4(cb);
    var = reduce4(var);
    dvar = reduce4(dvar) + 1e-5f;

    // background: colour of the last sample lives in lane q == 3, local index P-1
    float b0 = __shfl_sync(FULL_MASK, col[3 * P - 3], group_base | 3);
    float b1 = __shfl_sync(FULL_MASK, col[3 * P - 2], group_base | 3);
    float b2 = __shfl_sync(FULL_MASK, col[3 * P - 1], group_base | 3);
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
    const float expected = e_num / (acc + 1e-10f);

    // ---- chunk-wide bookkeeping (clip bounds of expected depth, beta NaN flag) ----
    const long long chunk = p.rays_per_chunk > 0 ? ray0 / p.rays_per_chunk : 0;
    bounds.enter(chunk, p.chunk_ws);
    if (active) {
#pragma unroll
      for (int i = 0; i < P; ++i) bounds.add_step(step[i]);
      bounds.has_nan |= nan_beta;
    }

    // ---- outputs: the 4 lanes of a ray split the stores ----
    if (active) {
      if (q == 0) {
        if (p.o_rgb) {
          p.o_rgb[ray * 3 + 0] = cr;
          p.o_rgb[ray * 3 + 1] = cg;
          p.o_rgb[ray * 3 + 2] = cb;
        }
      } else if (q == 1) {
        if (p.o_acc) p.o_acc[ray] = acc;
        if (p.o_depth) p.o_depth[ray] = depth;
        if (p.o_exp) p.o_exp[ray] = expected;
      } else if (q == 2) {
        if (p.o_rgb_var) p.o_rgb_var[ray] = var;
        if (p.o_rgb_std) p.o_rgb_std[ray] = sqrtf(var);
      } else {
        if (p.o_dvar) p.o_dvar[ray] = dvar;
        if (p.o_dstd) p.o_dstd[ray] = sqrtf(dvar);
      }
      if (p.o_w) {
        float4* ow = reinterpret_cast<float4*>(p.o_w + (size_t)ray * S + q * P);
#pragma unroll
        for (int j = 0; j < V; ++j)
          ow[j] = make_float4(wgt[4 * j], wgt[4 * j + 1], wgt[4 * j + 2], wgt[4 * j + 3]);
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
        float ddi = ok ? p.deltas[row + i] * p.density[row + i] : 0.f;
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
          float ddi = ok ? p.deltas[row + i] * p.density[row + i] : 0.f;
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
__global__ void __launch_bounds__(256) composite_finalize(const CompositeParams p) {
  const long long ray = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= p.num_rays) return;
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
        const float ddi = p.deltas[row + i] * p.density[row + i];
        const float wi = nan_to_num((1.0f - expf(-ddi)) * expf(-(float)carry));
        carry += (double)ddi;
        var += (wi * wi) * nan_to_num(p.beta[row + i]);
      }
      p.o_rgb_var[ray] = var;
      if (p.o_rgb_std) p.o_rgb_std[ray] = sqrtf(var);
    }
  }
}

static size_t chunk_ws_bytes(long long num_rays, long long rays_per_chunk) {
  long long chunks = 1;
  if (rays_per_chunk > 0 && num_rays > 0) chunks = (num_rays + rays_per_chunk - 1) / rays_per_chunk;
  return (size_t)chunks * 4 * sizeof(unsigned);
}

template <int S>
static int launch_tma(const CompositeParams& p, cudaStream_t stream) {
  constexpr int kConsumerWarps = consumer_warps_for(S);
  constexpr size_t stage_bytes = (size_t)kRaysPerTile * S * 8 * 4;
  constexpr size_t smem = kConsumerWarps * kStagesPerWarp * stage_bytes +
                          2 * kConsumerWarps * kStagesPerWarp * sizeof(uint64_t);
  static_assert(smem <= 227 * 1024, "stage ring exceeds shared memory");
  cudaError_t e = cudaFuncSetAttribute(composite_rays_tma<S>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("composite_rays: cannot reserve %zu B shared memory (%s)", smem, cudaGetErrorString(e));
    return UB_ERR_LAUNCH;
  }
  const long long tiles = (p.num_rays + kRaysPerTile - 1) / kRaysPerTile;
  const long long want = (tiles + kConsumerWarps - 1) / kConsumerWarps;
  int grid = sm_count();
  if (grid <= 0) grid = 148;
  if (want < grid) grid = (int)want;
  composite_rays_tma<S><<<grid, 32 * (1 + kConsumerWarps), smem, stream>>>(p);
  return check_launch("composite_rays_tma");
}

static bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; }

}  // namespace ub

extern "C" {

size_t ub_composite_rays_workspace_bytes(int64_t num_rays, int64_t rays_per_chunk) {
  return ub::chunk_ws_bytes(num_rays, rays_per_chunk);
}
size_t ub_render_weights_workspace_bytes(int64_t num_rays, int64_t rays_per_chunk) {
  return ub::chunk_ws_bytes(num_rays, rays_per_chunk);
}

int ub_composite_rays(const ub_composite_rays_args* a, void* workspace, size_t workspace_bytes,
                      void* stream_v) {
  using namespace ub;
  UB_REQUIRE(a != nullptr, UB_ERR_BAD_ARG, "composite_rays: args is NULL");
  UB_REQUIRE(a->num_rays >= 0 && a->num_samples >= 1, UB_ERR_BAD_ARG,
             "composite_rays: bad shape R=%lld S=%d", (long long)a->num_rays, a->num_samples);
  if (a->num_rays == 0) return UB_OK;
  UB_REQUIRE(a->density && a->deltas && a->starts && a->ends && a->rgb, UB_ERR_BAD_ARG,
             "composite_rays: density/deltas/starts/ends/rgb must be non-NULL");
  UB_REQUIRE(a->background_mode >= UB_BG_LAST_SAMPLE && a->background_mode <= UB_BG_FIXED,
             UB_ERR_BAD_ARG, "composite_rays: bad background_mode %d", a->background_mode);
  UB_REQUIRE(a->beta_mode == UB_BETA_RAW || a->beta_mode == UB_BETA_NAN_GUARD, UB_ERR_BAD_ARG,
             "composite_rays: bad beta_mode %d", a->beta_mode);
  UB_REQUIRE((a->out_rgb_var == nullptr && a->out_rgb_std == nullptr) || a->beta != nullptr,
             UB_ERR_BAD_ARG, "composite_rays: rgb_var/rgb_std requested without beta");
  const size_t need = chunk_ws_bytes(a->num_rays, a->rays_per_chunk);
  UB_REQUIRE(workspace != nullptr && workspace_bytes >= need, UB_ERR_WORKSPACE,
             "composite_rays: workspace %zu B < required %zu B", workspace_bytes, need);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);

  CompositeParams p{};
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

  if (cudaMemsetAsync(workspace, 0, need, stream) != cudaSuccess) return check_launch("composite_rays memset");

  const bool chunk_ok = a->rays_per_chunk <= 0 || a->rays_per_chunk % kRaysPerTile == 0;
  const bool align_ok = aligned16(a->density) && aligned16(a->deltas) && aligned16(a->starts) &&
                        aligned16(a->ends) && aligned16(a->rgb) &&
                        (a->beta == nullptr || aligned16(a->beta)) &&
                        (a->out_weights == nullptr || aligned16(a->out_weights));
  int rc = UB_OK;
  bool fast = chunk_ok && align_ok;
  if (fast) {
    switch (a->num_samples) {
      case 32: rc = launch_tma<32>(p, stream); break;
      case 48: rc = launch_tma<48>(p, stream); break;
      case 64: rc = launch_tma<64>(p, stream); break;
      case 96: rc = launch_tma<96>(p, stream); break;
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
  if (rc != UB_OK) return rc;
  const unsigned fblocks = (unsigned)((a->num_rays + 255) / 256);
  composite_finalize<<<fblocks, 256, 0, stream>>>(p);
  return check_launch("composite_finalize");
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
    const unsigned fblocks = (unsigned)((a->num_rays + 255) / 256);
    composite_finalize<<<fblocks, 256, 0, stream>>>(p);
    rc = check_launch("render_weights finalize");
  }
  return rc;
}

}  // extern "C"
