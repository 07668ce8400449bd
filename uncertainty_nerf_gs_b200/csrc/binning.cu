// (f2) Tile binning of projected 2-D Gaussians: (tile, depth)-sorted intersection lists for composite_tiles.
//
// The reference gets these lists from gsplat 0.1.x inside every `rasterize_gaussians` call
// (models/activesplatfacto/activesplatfacto_model.py:260-355: four times per view); gsplat is not vendored,
// the scheme below is its published one: tile rectangle of a Gaussian = centre +- radius in tile units
// ((int) truncation, clamped to the tile grid), one (tile << 32 | depth bits) key per covered tile, stable
// radix sort, per-tile [start, end) ranges.  BASELINE.json's splat configuration is "pre-binned", so this
// runs outside the timed compositing path; it exists so that the path does not depend on torch for it.
//
// Pipeline (the 64-bit key sort is done as two stable 32-bit sorts with the segmented radix sort of
// radix_sort.cu: first by depth, then by tile):
//   bin_count      per-Gaussian number of covered tiles
//   scan           exclusive prefix over Gaussians (block sums + one-block scan + fix-up)
//   bin_expand     one (tile id, depth, Gaussian id) record per intersection, in Gaussian order
//   [sort by depth] -> perm1;  gather tile ids by perm1;  [stable sort by tile] -> perm2
//   bin_finish     gaussian_ids = gid[perm1[perm2]], tile_bins from the boundaries of the sorted tile ids
#include "ub_common.cuh"

namespace ub {

__device__ __forceinline__ void tile_rect(float x, float y, int radius, int tiles_x, int tiles_y, int& x0, int& x1,
                                          int& y0, int& y1) {
  const float cx = x / UB_TILE, cy = y / UB_TILE, tr = (float)radius / UB_TILE;
  x0 = min(max(0, (int)(cx - tr)), tiles_x);
  x1 = min(max(0, (int)(cx + tr + 1.0f)), tiles_x);
  y0 = min(max(0, (int)(cy - tr)), tiles_y);
  y1 = min(max(0, (int)(cy + tr + 1.0f)), tiles_y);
}

__global__ void __launch_bounds__(256)
bin_count_kernel(const float* xys, const int32_t* radii, long long g_count, int tiles_x, int tiles_y,
                 long long* counts) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= g_count) return;
  long long n = 0;
  if (radii[g] > 0) {
    int x0, x1, y0, y1;
    tile_rect(xys[2 * g], xys[2 * g + 1], radii[g], tiles_x, tiles_y, x0, x1, y0, y1);
    n = (long long)max(0, x1 - x0) * max(0, y1 - y0);
  }
  counts[g] = n;
}

constexpr int kScanBlock = 1024;

// level 1: in-place exclusive scan inside each block of 1024 elements, block total -> sums[block]
__global__ void __launch_bounds__(kScanBlock) scan_blocks_kernel(long long* data, long long n, long long* sums) {
  __shared__ long long warp_tot[32];
  const long long i = (long long)blockIdx.x * kScanBlock + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long v = i < n ? data[i] : 0;
  long long incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long u = __shfl_up_sync(FULL_MASK, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    long long t = warp_tot[lane], ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long u = __shfl_up_sync(FULL_MASK, ti, o);
      if (lane >= o) ti += u;
    }
    warp_tot[lane] = ti - t;
    if (lane == 31 && sums) sums[blockIdx.x] = ti;
  }
  __syncthreads();
  if (i < n) data[i] = warp_tot[warp] + incl - v;
}

__global__ void __launch_bounds__(kScanBlock) scan_add_kernel(long long* data, long long n, const long long* offs) {
  const long long i = (long long)blockIdx.x * kScanBlock + threadIdx.x;
  if (i < n) data[i] += offs[blockIdx.x];
}

// one block: exclusive scan of up to kScanBlock * kScanBlock block sums; total -> *total_out
__global__ void __launch_bounds__(kScanBlock) scan_sums_kernel(long long* sums, long long m, long long* total_out) {
  __shared__ long long warp_tot[32];
  __shared__ long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long base = 0; base < m; base += kScanBlock) {
    const long long i = base + threadIdx.x;
    const long long v = i < m ? sums[i] : 0;
    long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long u = __shfl_up_sync(FULL_MASK, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      long long t = warp_tot[lane], ti = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long u = __shfl_up_sync(FULL_MASK, ti, o);
        if (lane >= o) ti += u;
      }
      warp_tot[lane] = ti - t;
    }
    __syncthreads();
    const long long carry = carry_s;
    if (i < m) sums[i] = carry + warp_tot[warp] + incl - v;
    __syncthreads();
    if (threadIdx.x == kScanBlock - 1) carry_s = carry + warp_tot[warp] + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry_s;
}

__global__ void __launch_bounds__(256)
bin_expand_kernel(const float* xys, const float* depths, const int32_t* radii, long long g_count, int tiles_x,
                  int tiles_y, const long long* offsets, float* tile_keys, float* depth_keys, int32_t* gids) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= g_count || radii[g] <= 0) return;
  int x0, x1, y0, y1;
  tile_rect(xys[2 * g], xys[2 * g + 1], radii[g], tiles_x, tiles_y, x0, x1, y0, y1);
  long long o = offsets[g];
  const float d = depths[g];
  for (int ty = y0; ty < y1; ++ty)
    for (int tx = x0; tx < x1; ++tx) {
      tile_keys[o] = (float)(ty * tiles_x + tx);  // exact: tile ids < 2^24
      depth_keys[o] = d;
      gids[o] = (int32_t)g;
      ++o;
    }
}

__global__ void __launch_bounds__(256)
gather_f32_kernel(const float* src, const int32_t* perm, long long n, float* dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[perm[i]];
}

__global__ void __launch_bounds__(256)
bin_finish_kernel(const float* sorted_tiles, const int32_t* perm1, const int32_t* perm2, const int32_t* gids,
                  long long n, int32_t* gaussian_ids, int32_t* tile_bins) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  gaussian_ids[i] = gids[perm1[perm2[i]]];
  const int t = (int)sorted_tiles[i];
  if (i == 0 || (int)sorted_tiles[i - 1] != t) tile_bins[2 * t + 0] = (int32_t)i;
  if (i == n - 1 || (int)sorted_tiles[i + 1] != t) tile_bins[2 * t + 1] = (int32_t)(i + 1);
}

// empty tiles: [start, end) = [p, p] with p = start of the next non-empty tile (what searchsorted yields)
__global__ void bin_fill_empty_kernel(int32_t* tile_bins, int tiles, int32_t total) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int32_t next = total;
  for (int t = tiles - 1; t >= 0; --t) {
    if (tile_bins[2 * t + 0] < 0) {
      tile_bins[2 * t + 0] = next;
      tile_bins[2 * t + 1] = next;
    } else {
      next = tile_bins[2 * t + 0];
    }
  }
}

}  // namespace ub

extern "C" {

size_t ub_bin_count_workspace_bytes(int64_t num_gaussians) {
  const size_t blocks = (size_t)((num_gaussians + ub::kScanBlock - 1) / ub::kScanBlock);
  return ub::align_up((size_t)(num_gaussians + 1) * 8, 256) + ub::align_up((blocks + 1) * 8, 256) + 256;
}

int ub_bin_count(const float* xys, const int32_t* radii, int64_t num_gaussians, int32_t img_height,
                 int32_t img_width, int64_t* out_offsets, int64_t* out_total, void* workspace,
                 size_t workspace_bytes, void* stream_v) {
  using namespace ub;
  UB_REQUIRE(num_gaussians >= 0 && img_height >= 1 && img_width >= 1, UB_ERR_BAD_ARG, "bin_count: bad sizes");
  UB_REQUIRE(out_offsets && out_total, UB_ERR_BAD_ARG, "bin_count: NULL output");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (num_gaussians == 0) {
    if (cudaMemsetAsync(out_total, 0, 8, stream) != cudaSuccess) return check_launch("bin_count memset");
    return UB_OK;
  }
  UB_REQUIRE(xys && radii, UB_ERR_BAD_ARG, "bin_count: NULL input");
  const long long blocks = (num_gaussians + kScanBlock - 1) / kScanBlock;
  UB_REQUIRE(blocks <= (long long)kScanBlock * kScanBlock, UB_ERR_UNSUPPORTED, "bin_count: too many Gaussians");
  const size_t need = align_up((size_t)(blocks + 1) * 8, 256);
  UB_REQUIRE(workspace != nullptr && workspace_bytes >= need, UB_ERR_WORKSPACE,
             "bin_count: workspace %zu B < required %zu B", workspace_bytes, need);
  long long* sums = static_cast<long long*>(workspace);
  const int tiles_x = (img_width + UB_TILE - 1) / UB_TILE, tiles_y = (img_height + UB_TILE - 1) / UB_TILE;
  long long* offs = reinterpret_cast<long long*>(out_offsets);
  bin_count_kernel<<<(unsigned)((num_gaussians + 255) / 256), 256, 0, stream>>>(xys, radii, num_gaussians, tiles_x,
                                                                               tiles_y, offs);
  scan_blocks_kernel<<<(unsigned)blocks, kScanBlock, 0, stream>>>(offs, num_gaussians, sums);
  scan_sums_kernel<<<1, kScanBlock, 0, stream>>>(sums, blocks, reinterpret_cast<long long*>(out_total));
  scan_add_kernel<<<(unsigned)blocks, kScanBlock, 0, stream>>>(offs, num_gaussians, sums);
  return check_launch("bin_count");
}

size_t ub_bin_gaussians_workspace_bytes(int64_t num_intersections) {
  const size_t n = (size_t)(num_intersections > 0 ? num_intersections : 1);
  // tile keys, depth keys, gathered tile keys, sorted tile keys (float) + gids, perm1, perm2 (int32) + offsets table
  return 7 * ub::align_up(n * 4, 256) + 512 + ub_segmented_sort_workspace_bytes(1, (int64_t)n, (int64_t)n, 1);
}

int ub_bin_gaussians(const float* xys, const float* depths, const int32_t* radii, int64_t num_gaussians,
                     int32_t img_height, int32_t img_width, const int64_t* offsets, int64_t num_intersections,
                     int32_t* out_gaussian_ids, int32_t* out_tile_bins, void* workspace, size_t workspace_bytes,
                     void* stream_v) {
  using namespace ub;
  UB_REQUIRE(num_gaussians >= 0 && num_intersections >= 0 && img_height >= 1 && img_width >= 1, UB_ERR_BAD_ARG,
             "bin_gaussians: bad sizes");
  UB_REQUIRE(num_intersections <= 0x7FFFFFFFLL, UB_ERR_UNSUPPORTED, "bin_gaussians: more than 2^31-1 intersections");
  UB_REQUIRE(out_tile_bins != nullptr, UB_ERR_BAD_ARG, "bin_gaussians: NULL output");
  const int tiles_x = (img_width + UB_TILE - 1) / UB_TILE, tiles_y = (img_height + UB_TILE - 1) / UB_TILE;
  const int tiles = tiles_x * tiles_y;
  UB_REQUIRE(tiles < (1 << 24), UB_ERR_UNSUPPORTED, "bin_gaussians: more than 2^24 tiles");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (cudaMemsetAsync(out_tile_bins, 0xFF, (size_t)tiles * 2 * sizeof(int32_t), stream) != cudaSuccess)
    return check_launch("bin_gaussians memset");
  if (num_intersections > 0) {
    UB_REQUIRE(xys && depths && radii && offsets && out_gaussian_ids, UB_ERR_BAD_ARG, "bin_gaussians: NULL pointer");
    const size_t need = ub_bin_gaussians_workspace_bytes(num_intersections);
    UB_REQUIRE(workspace != nullptr && workspace_bytes >= need, UB_ERR_WORKSPACE,
               "bin_gaussians: workspace %zu B < required %zu B", workspace_bytes, need);
    const size_t n = (size_t)num_intersections, stride = align_up(n * 4, 256);
    char* ws = static_cast<char*>(workspace);
    float* tile_keys = reinterpret_cast<float*>(ws + 0 * stride);
    float* depth_keys = reinterpret_cast<float*>(ws + 1 * stride);
    float* tile_by_depth = reinterpret_cast<float*>(ws + 2 * stride);
    float* tile_sorted = reinterpret_cast<float*>(ws + 3 * stride);
    int32_t* gids = reinterpret_cast<int32_t*>(ws + 4 * stride);
    int32_t* perm1 = reinterpret_cast<int32_t*>(ws + 5 * stride);
    int32_t* perm2 = reinterpret_cast<int32_t*>(ws + 6 * stride);
    long long* seg = reinterpret_cast<long long*>(ws + 7 * stride);
    char* sort_ws = ws + 7 * stride + 512;
    const size_t sort_ws_bytes = workspace_bytes - (7 * stride + 512);
    const long long seg_host[2] = {0, (long long)n};
    // 16-byte table, enqueued before anything reads it; tiny pageable copy (binning is a setup step)
    if (cudaMemcpyAsync(seg, seg_host, sizeof(seg_host), cudaMemcpyHostToDevice, stream) != cudaSuccess)
      return check_launch("bin_gaussians segment table");
    const unsigned gb = (unsigned)((num_gaussians + 255) / 256), ib = (unsigned)((n + 255) / 256);
    bin_expand_kernel<<<gb, 256, 0, stream>>>(xys, depths, radii, num_gaussians, tiles_x, tiles_y,
                                              reinterpret_cast<const long long*>(offsets), tile_keys, depth_keys, gids);
    int rc = ub_segmented_sort(depth_keys, 1, reinterpret_cast<const int64_t*>(seg), (int64_t)n, (int64_t)n, nullptr,
                               perm1, sort_ws, sort_ws_bytes, stream_v);
    if (rc != UB_OK) return rc;
    gather_f32_kernel<<<ib, 256, 0, stream>>>(tile_keys, perm1, (long long)n, tile_by_depth);
    rc = ub_segmented_sort(tile_by_depth, 1, reinterpret_cast<const int64_t*>(seg), (int64_t)n, (int64_t)n,
                           tile_sorted, perm2, sort_ws, sort_ws_bytes, stream_v);
    if (rc != UB_OK) return rc;
    bin_finish_kernel<<<ib, 256, 0, stream>>>(tile_sorted, perm1, perm2, gids, (long long)n, out_gaussian_ids,
                                              out_tile_bins);
  }
  bin_fill_empty_kernel<<<1, 32, 0, stream>>>(out_tile_bins, tiles, (int32_t)num_intersections);
  return check_launch("bin_gaussians");
}

}  // extern "C"
