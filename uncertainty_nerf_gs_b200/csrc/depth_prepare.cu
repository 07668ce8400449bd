// (a15) Preparation of the depth-scoring inputs: per view scale, clamp and keep the pixels with ground truth.
//
// Reference arithmetic (file:line under /root/reference/nerfuncertainty/scripts/eval_uncertainty.py):
//   :461-462   depth = a * depth;  depth_std = a * depth_std          (a = per-dataset scale)
//   :455-456   MIN_DEPTH = 1e-3;  MAX_DEPTH = depth_gt.max()
//   :558-560   depth[depth < MIN_DEPTH] = MIN_DEPTH;  depth[depth > MAX_DEPTH] = MAX_DEPTH   (masked assignments)
//   :552-560   mask = depth_gt > 0;  depth[mask], depth_std[mask], depth_gt[mask]
// The masked selections keep row-major pixel order (it decides the ties of the stable ranking downstream), so
// this is a *stable* stream compaction per view: block counts -> per-view exclusive scan -> scatter.  Views
// become ragged segments of one flat array; `out_offsets` (device int64 [B+1]) is what the caller reads back
// to size the scoring launches -- the only host synchronisation of the depth path.
// NaN semantics follow those masked assignments: a NaN depth fails both comparisons and stays NaN; torch's max()
// propagates NaN, so a NaN anywhere in a view's ground truth makes MAX_DEPTH NaN, `depth > NaN` is false everywhere
// and the view's depths are left unclamped above (not turned into NaN); NaN > 0 is false in the mask.
#include "ub_common.cuh"

namespace ub {

constexpr int kDepthThreads = 256;
constexpr int kDepthPerThread = 4;
constexpr int kDepthChunk = kDepthThreads * kDepthPerThread;  // pixels per block

struct DepthPrepParams {
  const float* depth;
  const float* depth_std;
  const float* depth_gt;
  const float* scales;  // device [B]
  long long pixels;     // per view
  int blocks_per_view;
  unsigned* block_counts;  // [B][blocks_per_view], exclusive-scanned in place by the second kernel
  unsigned* gt_max_key;    // [B] order keys
  unsigned* gt_has_nan;    // [B]
  long long* offsets;      // [B + 1]
  float* out_pred;
  float* out_std;
  float* out_gt;
};

__global__ void __launch_bounds__(kDepthThreads) depth_count_kernel(const DepthPrepParams p) {
  __shared__ unsigned warp_cnt[kDepthThreads / 32];
  __shared__ float warp_max[kDepthThreads / 32];
  __shared__ unsigned warp_nan[kDepthThreads / 32];
  const int view = blockIdx.y;
  const long long base = (long long)blockIdx.x * kDepthChunk;
  const float* gt = p.depth_gt + (long long)view * p.pixels;
  unsigned cnt = 0, has_nan = 0;
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < kDepthPerThread; ++k) {
    const long long i = base + (long long)k * kDepthThreads + threadIdx.x;
    if (i < p.pixels) {
      const float g = gt[i];
      cnt += g > 0.0f ? 1u : 0u;
      has_nan |= g != g ? 1u : 0u;
      mx = fmaxf(mx, g);  // fmaxf drops NaN; tracked separately
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(FULL_MASK, cnt, o);
    has_nan |= __shfl_xor_sync(FULL_MASK, has_nan, o);
    mx = fmaxf(mx, __shfl_xor_sync(FULL_MASK, mx, o));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    warp_cnt[warp] = cnt;
    warp_max[warp] = mx;
    warp_nan[warp] = has_nan;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned c = 0, n = 0;
    float m = -INFINITY;
    for (int w = 0; w < kDepthThreads / 32; ++w) {
      c += warp_cnt[w];
      n |= warp_nan[w];
      m = fmaxf(m, warp_max[w]);
    }
    p.block_counts[(size_t)view * p.blocks_per_view + blockIdx.x] = c;
    if (m > -INFINITY) atomicMax(&p.gt_max_key[view], order_key(m));
    if (n) p.gt_has_nan[view] = 1u;
  }
}

// one block per view: exclusive scan of the view's block counts in place, total into offsets[view + 1] (as a count)
__global__ void __launch_bounds__(256) depth_scan_kernel(const DepthPrepParams p) {
  __shared__ unsigned warp_tot[8];
  __shared__ unsigned carry_s;
  const int view = blockIdx.x;
  unsigned* row = p.block_counts + (size_t)view * p.blocks_per_view;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int b0 = 0; b0 < p.blocks_per_view; b0 += 256) {
    const int i = b0 + threadIdx.x;
    const unsigned v = i < p.blocks_per_view ? row[i] : 0u;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned u = __shfl_up_sync(FULL_MASK, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    unsigned wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += warp_tot[w];
    const unsigned carry = carry_s;
    if (i < p.blocks_per_view) row[i] = carry + wbase + incl - v;
    __syncthreads();
    if (threadIdx.x == 255) carry_s = carry + wbase + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) p.offsets[view + 1] = (long long)carry_s;
}

// offsets[0] = 0, offsets[v + 1] = running sum of the per-view counts (B is small: one thread)
__global__ void depth_offsets_kernel(long long* offsets, int num_views) {
  long long run = 0;
  offsets[0] = 0;
  for (int v = 0; v < num_views; ++v) {
    run += offsets[v + 1];
    offsets[v + 1] = run;
  }
}

__global__ void __launch_bounds__(kDepthThreads) depth_scatter_kernel(const DepthPrepParams p) {
  __shared__ unsigned warp_base[kDepthThreads / 32];
  const int view = blockIdx.y;
  const long long base = (long long)blockIdx.x * kDepthChunk;
  const long long voff = (long long)view * p.pixels;
  const float scale = p.scales[view];
  const float max_d = p.gt_has_nan[view] ? NAN : order_key_inv(p.gt_max_key[view]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long out = p.offsets[view] + p.block_counts[(size_t)view * p.blocks_per_view + blockIdx.x];
  // pixel order inside the block: k-major (k * 256 + thread), i.e. ascending pixel index
#pragma unroll
  for (int k = 0; k < kDepthPerThread; ++k) {
    const long long i = base + (long long)k * kDepthThreads + threadIdx.x;
    float g = 0.f;
    bool keep = false;
    if (i < p.pixels) {
      g = p.depth_gt[voff + i];
      keep = g > 0.0f;
    }
    const unsigned m = __ballot_sync(FULL_MASK, keep);
    if (lane == 0) warp_base[warp] = __popc(m);
    __syncthreads();
    unsigned before = 0, total = 0;
    for (int w = 0; w < kDepthThreads / 32; ++w) {
      const unsigned c = warp_base[w];
      before += w < warp ? c : 0u;
      total += c;
    }
    if (keep) {
      const long long dst = out + before + __popc(m & ((1u << lane) - 1u));
      float d = __fmul_rn(scale, p.depth[voff + i]);
      if (d < 1e-3f) d = 1e-3f;    // depth[depth < MIN_DEPTH] = MIN_DEPTH: false for a NaN depth
      if (d > max_d) d = max_d;    // depth[depth > MAX_DEPTH] = MAX_DEPTH: false for a NaN depth or a NaN MAX_DEPTH
      p.out_pred[dst] = d;
      p.out_std[dst] = __fmul_rn(scale, p.depth_std[voff + i]);
      p.out_gt[dst] = g;
    }
    out += total;
    __syncthreads();
  }
}

struct DepthPrepLayout {
  size_t off_counts, off_max, off_nan, total;
  int blocks_per_view;
};
static DepthPrepLayout depth_layout(int num_views, long long pixels) {
  DepthPrepLayout l{};
  l.blocks_per_view = (int)((pixels + kDepthChunk - 1) / kDepthChunk);
  if (l.blocks_per_view < 1) l.blocks_per_view = 1;
  size_t o = 0;
  l.off_counts = o;
  o = align_up(o + (size_t)num_views * l.blocks_per_view * sizeof(unsigned), 256);
  l.off_max = o;
  o = align_up(o + (size_t)num_views * sizeof(unsigned), 256);
  l.off_nan = o;
  o = align_up(o + (size_t)num_views * sizeof(unsigned), 256);
  l.total = o;
  return l;
}

}  // namespace ub

extern "C" {

size_t ub_depth_prepare_workspace_bytes(int32_t num_views, int64_t pixels_per_view) {
  if (num_views < 1 || pixels_per_view < 0) return 256;
  return ub::depth_layout(num_views, pixels_per_view).total;
}

int ub_depth_prepare(const float* depth, const float* depth_std, const float* depth_gt, const float* scales,
                     int32_t num_views, int64_t pixels_per_view, float* out_pred, float* out_std, float* out_gt,
                     int64_t* out_offsets, void* workspace, size_t workspace_bytes, void* stream_v) {
  using namespace ub;
  UB_REQUIRE(num_views >= 1 && num_views <= 65535 && pixels_per_view >= 0, UB_ERR_BAD_ARG,
             "depth_prepare: bad sizes (views %d, pixels %lld)", num_views, (long long)pixels_per_view);
  UB_REQUIRE(out_offsets != nullptr && scales != nullptr, UB_ERR_BAD_ARG, "depth_prepare: offsets / scales are NULL");
  UB_REQUIRE(pixels_per_view == 0 || (depth && depth_std && depth_gt && out_pred && out_std && out_gt),
             UB_ERR_BAD_ARG, "depth_prepare: input / output pointer is NULL");
  const DepthPrepLayout lay = depth_layout(num_views, pixels_per_view);
  UB_REQUIRE(workspace != nullptr && workspace_bytes >= lay.total, UB_ERR_WORKSPACE,
             "depth_prepare: workspace %zu B < required %zu B", workspace_bytes, lay.total);
  UB_REQUIRE(pixels_per_view < (1LL << 31), UB_ERR_UNSUPPORTED, "depth_prepare: view larger than 2^31 pixels");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  char* ws = static_cast<char*>(workspace);
  if (cudaMemsetAsync(ws, 0, lay.total, stream) != cudaSuccess) return check_launch("depth_prepare memset");
  DepthPrepParams p{};
  p.depth = depth;
  p.depth_std = depth_std;
  p.depth_gt = depth_gt;
  p.scales = scales;
  p.pixels = pixels_per_view;
  p.blocks_per_view = lay.blocks_per_view;
  p.block_counts = reinterpret_cast<unsigned*>(ws + lay.off_counts);
  p.gt_max_key = reinterpret_cast<unsigned*>(ws + lay.off_max);
  p.gt_has_nan = reinterpret_cast<unsigned*>(ws + lay.off_nan);
  p.offsets = reinterpret_cast<long long*>(out_offsets);
  p.out_pred = out_pred;
  p.out_std = out_std;
  p.out_gt = out_gt;
  dim3 grid((unsigned)lay.blocks_per_view, (unsigned)num_views);
  depth_count_kernel<<<grid, kDepthThreads, 0, stream>>>(p);
  depth_scan_kernel<<<num_views, 256, 0, stream>>>(p);
  depth_offsets_kernel<<<1, 1, 0, stream>>>(p.offsets, num_views);
  depth_scatter_kernel<<<grid, kDepthThreads, 0, stream>>>(p);
  return check_launch("depth_prepare");
}

}  // extern "C"
