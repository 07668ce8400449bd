// (C2) Per-image segmented stable LSD radix sort of uncertainty keys + prefix sums at cut points.
//
// Reference arithmetic (file:line under /root/reference/nerfuncertainty):
//   metrics/ause.py:10      err_vec_sorted, _ = torch.sort(err_vec)           (keys-only sort)
//   metrics/ause.py:25-26   _, idx = torch.sort(unc_vec); err_vec[idx]        (pair sort + gather)
//   metrics/ause.py:15-20, 29-34   mean of the first int((1-r) n) sorted errors (100 cut points)
//
// Ordering contract == torch.sort(stable=True) on float32: ascending,
// ties keep ascending original index, -0.0 == +0.0, every NaN after +inf.
//
// Layout: a batch of images is one flat key array split into segments.  Four 8-bit passes, each
//   upsweep   : per-tile digit histogram                       -> counts[seg][digit][tile]
//   scan      : one block per (digit, segment) scans its row of tile counts and records the digit
//               total; the downsweep turns the 256 totals into digit bases itself
//   downsweep : stable rank inside the tile (warp multi-split by 8 ballots, warp-private counters),
//               local reorder through shared memory, run-coalesced scatter.
// Pass 0 reads the float keys and synthesises the payload (index within the segment); pass 3
// writes straight into the caller's outputs.  Ping-pong buffers live in the workspace and stay
// L2-resident for image-sized segments.
#include "ub_common.cuh"

namespace ub {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kItems = 16;
constexpr int kTile = kSortThreads * kItems;  // 4096 keys per tile
constexpr int kRadix = 256;

struct SortPass {
  const float* in_float;     // pass 0 only
  const uint32_t* in_keys;   // passes 1..3
  const int32_t* in_vals;    // passes 1..3 when with_vals
  uint32_t* out_keys;        // may be NULL on the last pass
  float* out_float;          // last pass: sorted keys as float (may be NULL)
  int32_t* out_vals;         // NULL for keys-only
  const long long* seg_offsets;
  uint32_t* counts;          // [num_segments][kRadix][max_tiles]
  uint32_t* totals;          // [num_segments][kRadix]
  int max_tiles;
  int shift;
  int first_pass;
  int last_pass;
};

__device__ __forceinline__ uint32_t load_key(const SortPass& p, long long gidx) {
  return p.first_pass ? sort_key_from_float(p.in_float[gidx]) : p.in_keys[gidx];
}

// block-wide exclusive scan of one value per thread (256 threads); returns the exclusive prefix and
// leaves the block total in *total_out (shared)
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* warp_tmp,
                                                             uint32_t* total_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(FULL_MASK, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) warp_tmp[warp] = incl;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w) {
    const uint32_t t = warp_tmp[w];
    if (w < warp) base += t;
    tot += t;
  }
  if (threadIdx.x == 0 && total_out) *total_out = tot;
  __syncthreads();
  return base + incl - v;
}

__global__ void __launch_bounds__(kSortThreads) sort_upsweep(const SortPass p) {
  __shared__ uint32_t hist[kSortWarps][kRadix];
  const int seg = blockIdx.y, tile = blockIdx.x;
  const long long seg_lo = p.seg_offsets[seg];
  const long long len = p.seg_offsets[seg + 1] - seg_lo;
  const long long tile_lo = (long long)tile * kTile;
  if (tile_lo >= len) return;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kSortWarps * kRadix; i += kSortThreads) (&hist[0][0])[i] = 0u;
  __syncthreads();
  const int count = (int)min((long long)kTile, len - tile_lo);
  for (int i = threadIdx.x; i < count; i += kSortThreads) {
    const uint32_t k = load_key(p, seg_lo + tile_lo + i);
    atomicAdd(&hist[warp][(k >> p.shift) & 0xFFu], 1u);
  }
  __syncthreads();
  uint32_t v = 0;
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w) v += hist[w][threadIdx.x];
  p.counts[((size_t)seg * kRadix + threadIdx.x) * p.max_tiles + tile] = v;
}

// One block per (digit, segment): exclusive scan of that digit's tile counts in place + digit total.
__global__ void __launch_bounds__(kSortThreads) sort_scan_rows(const SortPass p) {
  __shared__ uint32_t warp_tmp[kSortWarps];
  __shared__ uint32_t chunk_total;
  const int d = blockIdx.x, seg = blockIdx.y;
  const long long len = p.seg_offsets[seg + 1] - p.seg_offsets[seg];
  const int tiles = (int)((len + kTile - 1) / kTile);
  uint32_t* row = p.counts + ((size_t)seg * kRadix + d) * p.max_tiles;
  uint32_t carry = 0;
  for (int base = 0; base < tiles; base += kSortThreads) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < tiles ? row[i] : 0u;
    const uint32_t ex = block_exclusive_scan_256(v, warp_tmp, &chunk_total);
    if (i < tiles) row[i] = carry + ex;
    carry += chunk_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) p.totals[(size_t)seg * kRadix + d] = carry;
}

template <bool WITH_VALS>
__global__ void __launch_bounds__(kSortThreads, 3) sort_downsweep(const SortPass p) {
  __shared__ uint32_t s_keys[kTile];
  __shared__ int32_t s_vals[WITH_VALS ? kTile : 1];
  __shared__ uint32_t cnt[kSortWarps][kRadix];
  __shared__ uint32_t digit_start[kRadix];
  __shared__ uint32_t gofs[kRadix];
  __shared__ uint32_t scan_tmp[kSortWarps];

  const int seg = blockIdx.y, tile = blockIdx.x;
  const long long seg_lo = p.seg_offsets[seg];
  const long long len = p.seg_offsets[seg + 1] - seg_lo;
  const long long tile_lo = (long long)tile * kTile;
  if (tile_lo >= len) return;
  const int count = (int)min((long long)kTile, len - tile_lo);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;

  for (int i = threadIdx.x; i < kSortWarps * kRadix; i += kSortThreads) (&cnt[0][0])[i] = 0u;

  // load: warp w owns positions [w*512, (w+1)*512) of the tile, item i of lane l = w*512 + i*32 + l
  // All 16 loads are issued before the first use (clamped addresses, no branches in between): a load
  // guarded by its own branch serialises into 16 dependent DRAM round trips.
  uint32_t key[kItems];
  int32_t val[kItems];
  uint32_t rank2[kItems / 2];  // two 16-bit ranks per register (rank < 4096)
  const uint32_t* src_keys = p.first_pass ? reinterpret_cast<const uint32_t*>(p.in_float) : p.in_keys;
  const long long tile_base = seg_lo + tile_lo;
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const int pos = warp * (32 * kItems) + i * 32 + lane;
    key[i] = __ldcs(src_keys + tile_base + min(pos, count - 1));
  }
  if (WITH_VALS && !p.first_pass) {
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      const int pos = warp * (32 * kItems) + i * 32 + lane;
      val[i] = __ldcs(p.in_vals + tile_base + min(pos, count - 1));
    }
  }
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const int pos = warp * (32 * kItems) + i * 32 + lane;
    if (p.first_pass) {
      key[i] = sort_key_from_float(__uint_as_float(key[i]));
      if (WITH_VALS) val[i] = (int32_t)(tile_lo + pos);
    }
    if (pos >= count) key[i] = 0xFFFFFFFFu;
  }
  // global base of every digit for this tile: exclusive scan of the 256 digit totals + row prefix
  {
    const int d = threadIdx.x;
    const uint32_t tot = p.totals[(size_t)seg * kRadix + d];
    const uint32_t base = block_exclusive_scan_256(tot, scan_tmp, nullptr);  // also syncs: counters zeroed
    gofs[d] = base + p.counts[((size_t)seg * kRadix + d) * p.max_tiles + tile];
  }

  // stable rank within the warp, digit by digit group (match-any multi-split)
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const int pos = warp * (32 * kItems) + i * 32 + lane;
    // positions past the end carry key 0xFFFFFFFF: digit 255, ranked after every real key (they are the
    // last positions of the tile), never stored
    const uint32_t d = (key[i] >> p.shift) & 0xFFu;
    (void)pos;
#ifdef UB_SORT_MATCH_ANY
    const unsigned peers = __match_any_sync(FULL_MASK, d);
#else
    unsigned peers = FULL_MASK;  // 8 ballots: MATCH.ANY serialises over the distinct values of the warp
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const bool bit = (d >> b) & 1u;
      const unsigned bal = __ballot_sync(FULL_MASK, bit);
      peers &= bit ? bal : ~bal;
    }
#endif
    const int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (lane == leader) {
      old = cnt[warp][d];
      cnt[warp][d] = old + __popc(peers);
    }
    old = __shfl_sync(FULL_MASK, old, leader);
    const uint32_t rk = old + __popc(peers & lt_mask);
    if (i & 1) rank2[i >> 1] |= rk << 16; else rank2[i >> 1] = rk;
    __syncwarp();
  }
  __syncthreads();

  // per digit: exclusive scan over warps (thread d owns digit d), then over digits
  {
    const int d = threadIdx.x;
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t t = cnt[w][d];
      cnt[w][d] = total;
      total += t;
    }
    digit_start[d] = block_exclusive_scan_256(total, scan_tmp, nullptr);
  }
  __syncthreads();

  // local reorder through shared memory
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const int pos = warp * (32 * kItems) + i * 32 + lane;
    if (pos < count) {
      const uint32_t d = (key[i] >> p.shift) & 0xFFu;
      const uint32_t rk = (i & 1) ? (rank2[i >> 1] >> 16) : (rank2[i >> 1] & 0xFFFFu);
      const uint32_t dst = digit_start[d] + cnt[warp][d] + rk;
      s_keys[dst] = key[i];
      if (WITH_VALS) s_vals[dst] = val[i];
    }
  }
  __syncthreads();

  // run-coalesced scatter
  for (int j = threadIdx.x; j < count; j += kSortThreads) {
    const uint32_t k = s_keys[j];
    const uint32_t d = (k >> p.shift) & 0xFFu;
    const long long dst = seg_lo + gofs[d] + (j - digit_start[d]);
    if (p.last_pass) {
      if (p.out_float) p.out_float[dst] = order_key_inv(k);
    } else {
      p.out_keys[dst] = k;
    }
    if (WITH_VALS) p.out_vals[dst] = s_vals[j];
  }
}

struct SortLayout {
  size_t off_offsets, off_counts, off_totals, off_keys_a, off_keys_b, off_vals_a, off_vals_b, total;
  int max_tiles;
};
static SortLayout sort_layout(int num_segments, long long total, long long max_len, bool with_vals) {
  SortLayout l{};
  l.max_tiles = (int)((max_len + kTile - 1) / kTile);
  if (l.max_tiles < 1) l.max_tiles = 1;
  size_t o = 0;
  l.off_offsets = o;
  o = align_up(o + (size_t)(num_segments + 1) * sizeof(long long), 256);
  l.off_counts = o;
  o = align_up(o + (size_t)num_segments * kRadix * l.max_tiles * sizeof(uint32_t), 256);
  l.off_totals = o;
  o = align_up(o + (size_t)num_segments * kRadix * sizeof(uint32_t), 256);
  l.off_keys_a = o;
  o = align_up(o + (size_t)total * 4, 256);
  l.off_keys_b = o;
  o = align_up(o + (size_t)total * 4, 256);
  l.off_vals_a = o;
  if (with_vals) o = align_up(o + (size_t)total * 4, 256);
  l.off_vals_b = o;
  if (with_vals) o = align_up(o + (size_t)total * 4, 256);
  l.total = o;
  return l;
}

// ---------------------------------------------------------------------------------------------
// Prefix sums at cut points (float64, deterministic two-level reduction).
constexpr int kCutChunk = 2048;
constexpr int kCutThreads = 256;
constexpr int kMaxCutValues = 8;

struct CutParams {
  const float* values[kMaxCutValues];
  const int32_t* perms[kMaxCutValues];  // per value array: gather permutation or NULL
  int num_values;
  const long long* seg_offsets;
  const long long* cuts;  // [num_segments][num_cuts]
  int num_cuts;
  double* block_tot;  // [num_segments][num_values][max_blocks]
  int max_blocks;
  double* out;        // [num_segments][num_values][num_cuts]
};

__device__ __forceinline__ double block_sum_256(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += shfl_xor_double(FULL_MASK, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < kCutThreads / 32; ++w) t += red[w];
  __syncthreads();
  return t;  // valid in thread 0
}

// grid (block, value, segment)
__global__ void __launch_bounds__(kCutThreads) cut_block_totals(const CutParams p) {
  __shared__ double red[kCutThreads / 32];
  const int blk = blockIdx.x, v = blockIdx.y, seg = blockIdx.z;
  const long long seg_lo = p.seg_offsets[seg];
  const long long len = p.seg_offsets[seg + 1] - seg_lo;
  const long long lo = (long long)blk * kCutChunk;
  if (lo >= len) return;
  const long long hi = min(len, lo + kCutChunk);
  const float* val = p.values[v];
  const int32_t* perm = p.perms[v];
  double acc = 0.0;
  for (long long i = lo + threadIdx.x; i < hi; i += kCutThreads) {
    const long long src = seg_lo + (perm ? (long long)perm[seg_lo + i] : i);
    acc += (double)val[src];
  }
  const double t = block_sum_256(acc, red);
  if (threadIdx.x == 0) p.block_tot[((size_t)seg * p.num_values + v) * p.max_blocks + blk] = t;
}

// one block per (cut, value, segment)
__global__ void __launch_bounds__(kCutThreads) cut_prefix_finish(const CutParams p) {
  __shared__ double red[kCutThreads / 32];
  const int c = blockIdx.x, v = blockIdx.y, seg = blockIdx.z;
  const long long seg_lo = p.seg_offsets[seg];
  const long long cut = p.cuts[(size_t)seg * p.num_cuts + c];
  const long long full_blocks = cut / kCutChunk;
  const double* bt = p.block_tot + ((size_t)seg * p.num_values + v) * p.max_blocks;
  const float* val = p.values[v];
  const int32_t* perm = p.perms[v];
  double acc = 0.0;
  for (long long b = threadIdx.x; b < full_blocks; b += kCutThreads) acc += bt[b];
  for (long long i = full_blocks * kCutChunk + threadIdx.x; i < cut; i += kCutThreads) {
    const long long src = seg_lo + (perm ? (long long)perm[seg_lo + i] : i);
    acc += (double)val[src];
  }
  const double t = block_sum_256(acc, red);
  if (threadIdx.x == 0) p.out[((size_t)seg * p.num_values + v) * p.num_cuts + c] = t;
}

struct CutLayout {
  size_t off_offsets, off_cuts, off_tot, total;
  int max_blocks;
};
static CutLayout cut_layout(int num_segments, long long max_len, int num_values, int num_cuts) {
  CutLayout l{};
  l.max_blocks = (int)((max_len + kCutChunk - 1) / kCutChunk);
  if (l.max_blocks < 1) l.max_blocks = 1;
  size_t o = 0;
  l.off_offsets = o;
  o = align_up(o + (size_t)(num_segments + 1) * sizeof(long long), 256);
  l.off_cuts = o;
  o = align_up(o + (size_t)num_segments * num_cuts * sizeof(long long), 256);
  l.off_tot = o;
  o = align_up(o + (size_t)num_segments * num_values * l.max_blocks * sizeof(double), 256);
  l.total = o;
  return l;
}

}  // namespace ub

extern "C" {

size_t ub_segmented_sort_workspace_bytes(int32_t num_segments, int64_t total, int64_t max_segment_len,
                                         int32_t with_perm) {
  if (num_segments < 1 || total < 0 || max_segment_len < 0) return 256;
  return ub::sort_layout(num_segments, total, max_segment_len, with_perm != 0).total;
}

int ub_segmented_sort(const float* keys, int32_t num_segments, const int64_t* seg_offsets, int64_t total,
                      int64_t max_segment_len, float* out_sorted_keys, int32_t* out_perm, void* workspace,
                      size_t workspace_bytes, void* stream_v) {
  using namespace ub;
  UB_REQUIRE(num_segments >= 1 && num_segments <= 65535 && seg_offsets != nullptr, UB_ERR_BAD_ARG,
             "segmented_sort: bad segments");
  UB_REQUIRE(total >= 0 && max_segment_len >= 0 && max_segment_len <= total, UB_ERR_BAD_ARG,
             "segmented_sort: bad total / max_segment_len");
  const long long max_len = max_segment_len;
  UB_REQUIRE(max_len <= 0x7FFFFFFFLL, UB_ERR_UNSUPPORTED, "segmented_sort: segment longer than 2^31-1");
  UB_REQUIRE(out_sorted_keys != nullptr || out_perm != nullptr, UB_ERR_BAD_ARG,
             "segmented_sort: nothing to output");
  if (total == 0) return UB_OK;
  UB_REQUIRE(keys != nullptr, UB_ERR_BAD_ARG, "segmented_sort: keys is NULL");
  const bool with_vals = out_perm != nullptr;
  const SortLayout lay = sort_layout(num_segments, total, max_len, with_vals);
  UB_REQUIRE(workspace != nullptr && workspace_bytes >= lay.total, UB_ERR_WORKSPACE,
             "segmented_sort: workspace %zu B < required %zu B", workspace_bytes, lay.total);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  char* ws = static_cast<char*>(workspace);

  uint32_t* ka = reinterpret_cast<uint32_t*>(ws + lay.off_keys_a);
  uint32_t* kb = reinterpret_cast<uint32_t*>(ws + lay.off_keys_b);
  int32_t* va = reinterpret_cast<int32_t*>(ws + lay.off_vals_a);
  int32_t* vb = reinterpret_cast<int32_t*>(ws + lay.off_vals_b);

  dim3 grid((unsigned)lay.max_tiles, (unsigned)num_segments);
  dim3 grid_scan(kRadix, (unsigned)num_segments);
  for (int pass = 0; pass < 4; ++pass) {
    SortPass p{};
    p.seg_offsets = reinterpret_cast<const long long*>(seg_offsets);
    p.counts = reinterpret_cast<uint32_t*>(ws + lay.off_counts);
    p.totals = reinterpret_cast<uint32_t*>(ws + lay.off_totals);
    p.max_tiles = lay.max_tiles;
    p.shift = pass * 8;
    p.first_pass = pass == 0;
    p.last_pass = pass == 3;
    p.in_float = keys;
    p.in_keys = (pass & 1) ? ka : kb;   // pass 0 writes A, 1 reads A writes B, 2 reads B writes A, 3 reads A
    p.in_vals = (pass & 1) ? va : vb;
    p.out_keys = (pass & 1) ? kb : ka;
    p.out_vals = (pass & 1) ? vb : va;
    if (p.last_pass) {
      p.out_keys = nullptr;
      p.out_float = out_sorted_keys;
      p.out_vals = out_perm;
    }
    sort_upsweep<<<grid, kSortThreads, 0, stream>>>(p);
    sort_scan_rows<<<grid_scan, kSortThreads, 0, stream>>>(p);
    if (with_vals)
      sort_downsweep<true><<<grid, kSortThreads, 0, stream>>>(p);
    else
      sort_downsweep<false><<<grid, kSortThreads, 0, stream>>>(p);
    int rc = check_launch("segmented_sort pass");
    if (rc != UB_OK) return rc;
  }
  return UB_OK;
}

size_t ub_cut_prefix_sums_workspace_bytes(int32_t num_segments, int64_t max_segment_len,
                                          int32_t num_values, int32_t num_cuts) {
  if (num_segments < 1 || num_values < 1 || num_cuts < 1 || max_segment_len < 0) return 256;
  return ub::cut_layout(num_segments, max_segment_len, num_values, num_cuts).total;
}

int ub_cut_prefix_sums(const float* const* values_host, const int32_t* const* perms_host,
                       int32_t num_values, int32_t num_segments, const int64_t* seg_offsets,
                       int64_t max_segment_len, const int64_t* cuts, int32_t num_cuts, double* out_sums,
                       void* workspace, size_t workspace_bytes, void* stream_v) {
  using namespace ub;
  UB_REQUIRE(values_host != nullptr && num_values >= 1 && num_values <= kMaxCutValues, UB_ERR_BAD_ARG,
             "cut_prefix_sums: num_values must be in [1, %d]", kMaxCutValues);
  UB_REQUIRE(num_segments >= 1 && num_segments <= 65535 && seg_offsets != nullptr, UB_ERR_BAD_ARG,
             "cut_prefix_sums: bad segments");
  UB_REQUIRE(num_cuts >= 1 && cuts != nullptr && out_sums != nullptr && max_segment_len >= 0, UB_ERR_BAD_ARG,
             "cut_prefix_sums: bad cuts / output");
  for (int v = 0; v < num_values; ++v)
    UB_REQUIRE(values_host[v] != nullptr || max_segment_len == 0, UB_ERR_BAD_ARG,
               "cut_prefix_sums: values[%d] is NULL", v);
  const long long max_len = max_segment_len;
  const CutLayout lay = cut_layout(num_segments, max_len, num_values, num_cuts);
  UB_REQUIRE(workspace != nullptr && workspace_bytes >= lay.total, UB_ERR_WORKSPACE,
             "cut_prefix_sums: workspace %zu B < required %zu B", workspace_bytes, lay.total);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  char* ws = static_cast<char*>(workspace);
  CutParams p{};
  for (int v = 0; v < num_values; ++v) {
    p.values[v] = values_host[v];
    p.perms[v] = perms_host ? perms_host[v] : nullptr;
  }
  p.num_values = num_values;
  p.seg_offsets = reinterpret_cast<const long long*>(seg_offsets);
  p.cuts = reinterpret_cast<const long long*>(cuts);
  p.num_cuts = num_cuts;
  p.block_tot = reinterpret_cast<double*>(ws + lay.off_tot);
  p.max_blocks = lay.max_blocks;
  p.out = out_sums;
  if (max_len > 0) {
    dim3 g1((unsigned)lay.max_blocks, (unsigned)num_values, (unsigned)num_segments);
    cut_block_totals<<<g1, kCutThreads, 0, stream>>>(p);
  }
  dim3 g2((unsigned)num_cuts, (unsigned)num_values, (unsigned)num_segments);
  cut_prefix_finish<<<g2, kCutThreads, 0, stream>>>(p);
  return check_launch("cut_prefix_sums");
}

}  // extern "C"
