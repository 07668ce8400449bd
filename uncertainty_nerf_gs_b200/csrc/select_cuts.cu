// (C2') AUSE cut-point sums by a multi-cut radix select instead of a full sort.
//
// Reference arithmetic (file:line under /root/reference/nerfuncertainty):
//   metrics/ause.py:10, 15-20     err sorted ascending, mean of the first int((1-r) n) values, 100 ratios
//   metrics/ause.py:25-26, 29-34  idx = sort(unc); err[idx]; the same 100 prefix means
// The curves only need  P_k = sum of the payload over the FIRST c_k ELEMENTS of the stable ascending order
// of the keys  at the 100 cut counts c_k; the permutation itself is never returned.  So instead of four
// scatter passes over every key (ub_segmented_sort) the keys are only *classified* against the cuts:
//
//   coarse     per segment: histogram of the top 12 bits of the order-preserving uint32 key
//   alloc      every coarse bin is split into 2^k fine bins in proportion to its count (<= 8192 fine bins
//              per segment, each holding <= n/2048 keys unless the keys tie): an equal-frequency binning
//              that needs neither the key range nor splitters; a sorted 1024-key sample finds the heavy
//              tie values (a clamped uncertainty floor), each of which gets a bin of its own
//   fine       histogram over the fine bins; per 2048-key tile counts of the heavy tie values
//   locate     prefix over the fine bins; every cut falls into one bin (its "cell") at a residual rank r;
//              a bin that holds no cut gets a class = number of cuts that exclude it
//   (locate)   a big cell whose keys are all equal (a tie group) is resolved by index: stable order inside
//              it is the element order, so the tile where the running count crosses r follows from the
//              per-tile counts and only that tile's members stay undecided; other cells stay undecided whole
//   classify   one pass over keys + payloads: decided elements add their payload to the sum of their class
//              (exact fixed-point limbs with native 32-bit shared-memory atomics, folded into float64 per
//              block); undecided ones (typically 1-3 % of the keys) go to their cell's slot of a side list
//              as (key, index, payloads) records
//   resolve    per cell (or per undecided tile of a tie group): records bucketed by the top 8 bits in which
//              their (key, index) differ, the members of a bucket that a cut splits ranked pairwise, hence
//              the class of every record; a radix select over the records takes over for cells beyond 2048
//              records (an unsampled tie group next to other keys of its fine bin -- correct, only slower)
//   finish     class sums -> prefix over the classes = the sums under every cut.
//
// The sets of elements under every cut are exactly those of torch.sort(stable=True) (same key transform:
// -0.0 == +0.0, NaN last, ties by index), so the sums equal the sort path's up to float64 summation order.
#include "ub_common.cuh"

namespace ub {

constexpr int kSelThreads = 256;
constexpr int kSelWarps = kSelThreads / 32;
constexpr int kSelLocThreads = 1024;  // sel_locate: one block per segment, latency-bound
constexpr int kSelResThreads = 128;   // sel_resolve: many small runs
constexpr int kSelCoarse = 4096;   // top 12 key bits
constexpr int kSelLowBits = 20;
constexpr uint32_t kSelLowMask = (1u << kSelLowBits) - 1u;
constexpr int kSelFine = 8192;     // fine bins per segment (upper bound by construction)
constexpr int kSelMaxHeavy = 16;   // tie values that get a bin of their own
constexpr int kSelBins = kSelFine + 2 * kSelMaxHeavy;
constexpr int kSelSample = 1024;   // keys sampled per segment to find the heavy tie values
constexpr int kSelHeavyHits = 8;   // sample hits that make a key heavy (~0.8 % of the segment)
constexpr int kSelTile = 2048;     // classification tile
constexpr int kSelItems = kSelTile / kSelThreads;
constexpr int kSelMaxTilesPerBlock = 8;   // 16384 keys: the 16-bit fixed-point limbs of a block stay below 2^31
constexpr int kSelChunk = 8192;    // keys per block in the histogram kernels
constexpr int kSelMaxCuts = 128;
constexpr int kSelMaxFamilies = 4;
constexpr uint32_t kSelTiledMin = 2 * kSelTile;  // tie groups above this size are resolved by tile
constexpr int kSelBrute = 2048;    // records of a run handled in shared memory
constexpr uint16_t kSelCellFlag = 0x8000u;
constexpr uint8_t kSelCompact = 0xFFu;

struct RunDesc {        // what the resolve block of cut j works on
  int leader;           // 0: nothing to do for this cut's block
  int cell, j0, nj;     // the cell and its cuts [j0, j0 + nj)
  uint32_t start, len;  // records [start, start + len) of the cell's slot
  uint32_t base;        // first record of the cell's slot in the segment's side list
  uint32_t pad;
};

struct SegPlan {
  int nfine, nbins, ncells, nheavy, ntiled;
  int emax;  // biased exponent of the largest finite |key| of the segment
  int flags;  // bit 0: the segment holds negative keys, bit 1: it holds NaN keys (from the coarse histogram)
  int pad2;
  uint32_t heavy[kSelMaxHeavy];      // heavy tie values (any order)
  int tcell[kSelMaxHeavy];           // tiled cells: tie groups that hold a cut and are resolved by tile
  int theavy[kSelMaxHeavy];          // tiled cell -> index of its key among the heavy tie values (row of tilecounts)
  int cut_k[kSelMaxCuts];            // original index of the j-th smallest cut
  int cut_T[kSelMaxCuts];            // bins < T are under the cut
  int cut_cell[kSelMaxCuts];         // cell that holds the cut, -1 if the cut is a bin boundary
  uint32_t cut_r[kSelMaxCuts];       // elements of the cell under the cut (stable order)
  uint32_t cut_posoff[kSelMaxCuts];  // records of the cell's slot under the cut (in (key, index) order)
  RunDesc run[kSelMaxCuts];
  int cell_bin[kSelMaxCuts];
  uint32_t cell_count[kSelMaxCuts];
  uint32_t cell_comp[kSelMaxCuts];   // records in the cell's slot
  int cell_j0[kSelMaxCuts], cell_j1[kSelMaxCuts];  // cuts [j0, j1) lie inside the cell
  int cell_tiled[kSelMaxCuts];       // index into tcell / theavy, -1: the whole cell goes to the side list
};

struct SelParams {
  const float* keys[kSelMaxFamilies];
  const float* pay0[kSelMaxFamilies];
  const float* pay1[kSelMaxFamilies];  // NULL: one payload
  int self_payload[kSelMaxFamilies];   // payload 0 is the key itself
  int row0[kSelMaxFamilies];           // first output row of the family
  int npay[kSelMaxFamilies];
  int pay_src[kSelMaxFamilies][2];     // family whose keys are this payload array (its magnitude bound is known), or -1
  float* cpay0[kSelMaxFamilies];       // record payloads [total] (NULL when self_payload)
  float* cpay1[kSelMaxFamilies];
  int num_families, num_views, num_values, num_cuts;
  long long total;
  int max_tiles, max_blocks, tiles_per_block;
  const long long* view_offsets;  // [B + 1]
  const long long* cuts;          // [B][num_cuts]
  double* out;                    // [B][V][num_cuts]
  uint32_t* hist_c;       // [G][kSelCoarse]                               zeroed
  uint32_t* hist_f;       // [G][kSelBins]                                 zeroed
  uint32_t* cursor;       // [G][kSelMaxCuts] records written to each cell's slot   zeroed
  double* ssum;           // [G][kSelMaxCuts + 1][2] class sums of the records      zeroed
  uint32_t* table;        // [G][kSelCoarse]: (first fine bin << 5) | shift of the low 20 key bits
  SegPlan* plan;          // [G]
  uint16_t* binmap;       // [G][kSelBins]: class, or kSelCellFlag | cell
  uint32_t* tilecounts;   // [G][kSelMaxHeavy][max_tiles] per heavy tie value; after sel_locate, rows of tiled cells:
                          // first record of (tiled cell, tile)
  uint8_t* tilemode;      // [G][max_tiles][kSelMaxHeavy]: class of the tie group's members in the tile, or kSelCompact
  double* spart;          // [G][max_blocks][num_cuts + 1][2]
  double* fpart;          // [G][kSelFinGroups][kSelFinStride] group totals of sel_finish
  uint32_t* fcount;       // [G] sel_finish blocks that have arrived                       zeroed
  uint32_t* ckeys;        // [F * total] records: order-preserving key
  uint32_t* cidx;         // [F * total] records: index within the segment
};

__device__ __forceinline__ void sel_segment(const SelParams& p, int g, int& f, int& b, long long& lo,
                                            long long& len) {
  f = g / p.num_views;
  b = g - f * p.num_views;
  lo = p.view_offsets[b];
  len = p.view_offsets[b + 1] - lo;
}

// bin of a key: its fine bin, shifted so that every heavy tie value K owns a bin (keys of K's fine bin below
// K, K itself and the keys above K get three consecutive indices): monotone in the key.
__device__ __forceinline__ int sel_bin(uint32_t u, uint32_t t, const uint32_t* heavy, int nheavy) {
  int bin = (int)(t >> 5) + (int)((u & kSelLowMask) >> (t & 31u));
  for (int h = 0; h < nheavy; ++h) bin += (int)(heavy[h] < u) + (int)(heavy[h] <= u);
  return bin;
}

// exclusive scan of one value per thread over a block of 32 * WARPS threads
template <int WARPS = kSelWarps>
__device__ __forceinline__ uint32_t sel_block_excl_scan(uint32_t v, uint32_t* warp_tmp, uint32_t* total_out) {
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
  for (int w = 0; w < WARPS; ++w) {
    const uint32_t t = warp_tmp[w];
    if (w < warp) base += t;
    tot += t;
  }
  if (total_out) *total_out = tot;
  __syncthreads();
  return base + incl - v;
}

// ---- coarse histogram: top 12 key bits ----------------------------------------------------------------
__global__ void __launch_bounds__(kSelThreads) sel_coarse_hist(const SelParams p) {
  __shared__ uint32_t h[kSelCoarse];
  const int g = blockIdx.y;
  int f, b;
  long long lo, len;
  sel_segment(p, g, f, b, lo, len);
  const long long start = (long long)blockIdx.x * kSelChunk;
  if (start >= len) return;
  for (int i = threadIdx.x; i < kSelCoarse; i += kSelThreads) h[i] = 0u;
  __syncthreads();
  const int count = (int)min((long long)kSelChunk, len - start);
  const float* k = p.keys[f] + lo + start;
  constexpr int per = kSelChunk / kSelThreads;
  for (int base = 0; base < per; base += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldg(k + min((base + i) * kSelThreads + (int)threadIdx.x, count - 1));
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if ((base + i) * kSelThreads + (int)threadIdx.x < count)
        atomicAdd(&h[sort_key_from_float(v[i]) >> kSelLowBits], 1u);
  }
  __syncthreads();
  uint32_t* gh = p.hist_c + (size_t)g * kSelCoarse;
  for (int i = threadIdx.x; i < kSelCoarse; i += kSelThreads) {
    const uint32_t v = h[i];
    if (v) atomicAdd(gh + i, v);
  }
}

// ---- fine-bin allocation + heavy tie values: one block of 1024 threads per segment (a latency chain) ----------
constexpr int kSelAllocThreads = 1024;
__global__ void __launch_bounds__(kSelAllocThreads, 2) sel_alloc(const SelParams p) {
  constexpr int kWarps = kSelAllocThreads / 32;
  static_assert(kSelSample == kSelAllocThreads, "one sample per thread");
  __shared__ uint32_t warp_tmp[kWarps];
  __shared__ uint32_t s_sample[kSelSample];
  const int g = blockIdx.x, tid = threadIdx.x;
  int f, b;
  long long lo, len;
  sel_segment(p, g, f, b, lo, len);
  SegPlan& pl = p.plan[g];
  // the sample first: its loads are the longest latency of the kernel
  const int ns = (int)min((long long)kSelSample, len);
  const float* k = p.keys[f] + lo;
  uint32_t x = tid < ns ? sort_key_from_float(__ldg(k + (long long)tid * len / ns)) : 0xFFFFFFFFu;
  const uint32_t target = (uint32_t)max(1LL, (len + 2047) / 2048);
  constexpr int per = kSelCoarse / kSelAllocThreads;
  const uint32_t* gh = p.hist_c + (size_t)g * kSelCoarse;
  uint32_t nsub[per], lg[per], sum = 0;
  int emax = 0, flags = 0;
#pragma unroll
  for (int i = 0; i < per; ++i) {
    const uint32_t cnt = gh[tid * per + i];
    if (cnt) {  // exponent field of the keys of this coarse bin (sign, 8 exponent bits, 3 mantissa bits)
      const int c = tid * per + i;
      const int ex = ((c >= 2048 ? c : ~c) >> 3) & 0xFF;
      if (ex != 255) emax = max(emax, ex);
      if (c < 2048) flags |= 1;              // negative keys (-0.0 counts as +0.0)
      if (c == kSelCoarse - 1) flags |= 2;   // the NaN key 0xFFFFFFFE
    }
    uint32_t l = 0;
    if (cnt > target) {
      const uint32_t q = (cnt + target - 1) / target;  // >= 2
      l = min(32u - (uint32_t)__clz(q - 1u), (uint32_t)kSelLowBits);
    }
    lg[i] = l;
    nsub[i] = cnt ? (1u << l) : 0u;
    sum += nsub[i];
  }
  uint32_t tot;
  uint32_t run = sel_block_excl_scan<kWarps>(sum, warp_tmp, &tot);
  uint32_t* tab = p.table + (size_t)g * kSelCoarse;
#pragma unroll
  for (int i = 0; i < per; ++i) {
    tab[tid * per + i] = (run << 5) | ((uint32_t)kSelLowBits - lg[i]);
    run += nsub[i];
  }
  // heavy tie values from a sorted sample: a key that fills >= 8 of 1024 evenly spaced probes.  A tie group
  // next to other keys of the same fine bin would make its cell "large and not one tie group" (the slow
  // radix-select corner of sel_resolve); owning a bin makes it a tie-group cell, resolved by index.
  // Bitonic sort, ascending, one key per thread: partners closer than a warp exchange by shuffle.
  for (int size = 2; size <= kSelSample; size <<= 1) {
    const bool up = (tid & size) == 0;
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      uint32_t y;
      if (stride >= 32) {
        s_sample[tid] = x;
        __syncthreads();
        y = s_sample[tid ^ stride];
        __syncthreads();
      } else {
        y = __shfl_xor_sync(FULL_MASK, x, stride);
      }
      const bool low = (tid & stride) == 0;  // the lower index of the pair keeps the smaller key in an ascending run
      x = (low == up) ? min(x, y) : max(x, y);
    }
  }
  s_sample[tid] = x;
  __shared__ int s_nh, s_emax, s_flags;
  __shared__ uint32_t s_hv[kSelMaxHeavy];
  if (tid == 0) {
    s_emax = 0;
    s_flags = 0;
  }
  __syncthreads();
  emax = __reduce_max_sync(FULL_MASK, emax);
  flags = __reduce_or_sync(FULL_MASK, flags);
  if ((tid & 31) == 0 && emax) atomicMax(&s_emax, emax);
  if ((tid & 31) == 0 && flags) atomicOr(&s_flags, flags);
  for (int thresh = kSelHeavyHits;; thresh *= 2) {  // at most kSelMaxHeavy values: raise the bar until they fit
    if (tid == 0) s_nh = 0;
    __syncthreads();
    if (tid < ns) {
      const bool first = tid == 0 || s_sample[tid - 1] != x;
      if (first && tid + thresh - 1 < ns && s_sample[tid + thresh - 1] == x) {
        const int slot = atomicAdd(&s_nh, 1);
        if (slot < kSelMaxHeavy) s_hv[slot] = x;
      }
    }
    __syncthreads();
    if (s_nh <= kSelMaxHeavy) break;
    __syncthreads();
  }
  const int nh = s_nh;
  if (tid < kSelMaxHeavy) pl.heavy[tid] = tid < nh ? s_hv[tid] : 0xFFFFFFFFu;
  if (tid == 0) {
    pl.emax = s_emax;
    pl.flags = s_flags;
    pl.nheavy = nh;
    pl.nfine = (int)tot;  // <= kSelFine: sum pow2ceil(cnt / target) <= 4096 + 2 n / target
    pl.nbins = (int)tot + 2 * nh;
  }
}

// ---- fine histogram --------------------------------------------------------------------------------------
// Order-preserving key of a segment known (from its coarse histogram) to hold neither negative values nor NaN:
// one OR instead of the general transform (-0.0 and +0.0 both map to 0x80000000, as in sort_key_from_float).
template <bool NONNEG>
__device__ __forceinline__ uint32_t sel_key(float f) {
  return NONNEG ? (__float_as_uint(f) | 0x80000000u) : sort_key_from_float(f);
}

// bins of eight keys at once: the eight table loads are in flight together, and the heavy-value adjustment runs
// value by value over all eight keys (one shared-memory load per heavy value instead of eight)
// `tcnt` (fine histogram only): per heavy value, the number of its occurrences among these eight keys per thread --
// one 2048-key tile per block-wide call -- is added to tcnt[h] (stable order inside a tie group is the element order,
// so a tie group that turns out to hold a cut is resolved from these per-tile counts, see sel_plan_tiled_cell)
template <bool NONNEG, bool COUNT = false>
__device__ __forceinline__ void sel_bins8(const float (&v)[8], const uint32_t* __restrict__ tab,
                                          const uint32_t* heavy, int nh, uint32_t (&u)[8], int (&bin)[8],
                                          uint32_t* tcnt = nullptr, uint32_t valid = 0xFFu) {
  uint32_t t[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    u[i] = sel_key<NONNEG>(v[i]);
    t[i] = __ldg(tab + (u[i] >> kSelLowBits));
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) bin[i] = (int)(t[i] >> 5) + (int)((u[i] & kSelLowMask) >> (t[i] & 31u));
  for (int h = 0; h < nh; ++h) {
    const uint32_t hv = heavy[h];
#pragma unroll
    for (int i = 0; i < 8; ++i) bin[i] += (int)(hv < u[i]) + (int)(hv <= u[i]);
    if (COUNT) {
      uint32_t c = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) c += (uint32_t)(hv == u[i]) & (valid >> i);
      c = __reduce_add_sync(FULL_MASK, c);
      if ((threadIdx.x & 31) == 0 && c) atomicAdd(tcnt + h, c);
    }
  }
}

template <bool NONNEG>
__device__ __forceinline__ void sel_fine_hist_body(const float* __restrict__ k, int count,
                                                   const uint32_t* __restrict__ tab, const uint32_t* s_heavy, int nh,
                                                   uint32_t* h, uint32_t (*tcnt)[kSelMaxHeavy]) {
  constexpr int per = kSelChunk / kSelThreads;
  const int tid = threadIdx.x;
  if (count == kSelChunk) {  // whole chunk: no clamps, no predicates
    for (int base = 0; base < per; base += 8) {
      float v[8];
      uint32_t u[8];
      int bin[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __ldg(k + (base + i) * kSelThreads + tid);
      sel_bins8<NONNEG, true>(v, tab, s_heavy, nh, u, bin, tcnt[base / 8]);
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&h[bin[i]], 1u);
    }
  } else {
    for (int base = 0; base < per; base += 8) {
      float v[8];
      uint32_t u[8];
      int bin[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __ldg(k + min((base + i) * kSelThreads + tid, count - 1));
      uint32_t valid = 0u;
#pragma unroll
      for (int i = 0; i < 8; ++i) valid |= (uint32_t)((base + i) * kSelThreads + tid < count) << i;
      sel_bins8<NONNEG, true>(v, tab, s_heavy, nh, u, bin, tcnt[base / 8], valid);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if ((valid >> i) & 1u) atomicAdd(&h[bin[i]], 1u);
    }
  }
}

__global__ void __launch_bounds__(kSelThreads) sel_fine_hist(const SelParams p) {
  __shared__ uint32_t h[kSelBins];
  __shared__ uint32_t s_heavy[kSelMaxHeavy];
  const int g = blockIdx.y;
  int f, b;
  long long lo, len;
  sel_segment(p, g, f, b, lo, len);
  const long long start = (long long)blockIdx.x * kSelChunk;
  if (start >= len) return;
  const SegPlan& pl = p.plan[g];
  const int nb = pl.nbins, nh = pl.nheavy;
  constexpr int kTilesPerChunk = kSelChunk / kSelTile;
  static_assert(kSelTile == 8 * kSelThreads, "one sel_bins8 call of the block covers one tile");
  __shared__ uint32_t s_tcnt[kTilesPerChunk][kSelMaxHeavy];
  for (int i = threadIdx.x; i < nb; i += kSelThreads) h[i] = 0u;
  if (threadIdx.x < kSelMaxHeavy) s_heavy[threadIdx.x] = pl.heavy[threadIdx.x];
  if (threadIdx.x < kTilesPerChunk * kSelMaxHeavy) (&s_tcnt[0][0])[threadIdx.x] = 0u;
  __syncthreads();
  const int count = (int)min((long long)kSelChunk, len - start);
  const float* k = p.keys[f] + lo + start;
  const uint32_t* tab = p.table + (size_t)g * kSelCoarse;
  if ((pl.flags & 3) == 0) sel_fine_hist_body<true>(k, count, tab, s_heavy, nh, h, s_tcnt);
  else sel_fine_hist_body<false>(k, count, tab, s_heavy, nh, h, s_tcnt);
  __syncthreads();
  if (threadIdx.x < kTilesPerChunk * kSelMaxHeavy) {  // per-tile counts of every heavy tie value
    const int tl = threadIdx.x / kSelMaxHeavy, hh = threadIdx.x % kSelMaxHeavy;
    const int tile = blockIdx.x * kTilesPerChunk + tl;
    if (hh < nh && (long long)tile * kSelTile < len)
      p.tilecounts[((size_t)g * kSelMaxHeavy + hh) * p.max_tiles + tile] = s_tcnt[tl][hh];
  }
  uint32_t* gh = p.hist_f + (size_t)g * kSelBins;
  for (int i = threadIdx.x; i < nb; i += kSelThreads) {
    const uint32_t v = h[i];
    if (v) atomicAdd(gh + i, v);
  }
}

// ---- plan of one tie group that holds cuts (tiled cell h of segment g): called by every thread of sel_locate's block
template <int WARPS>
__device__ __forceinline__ void sel_plan_tiled_cell(const SelParams& p, int g, int h) {
  constexpr int kThreads = WARPS * 32;
  __shared__ uint32_t warp_tmp[WARPS];
  __shared__ int s_tstar[kSelMaxCuts];
  __shared__ uint32_t s_rho[kSelMaxCuts];
  const int tid = threadIdx.x;
  SegPlan& pl = p.plan[g];
  const int cell = pl.tcell[h];
  int f, b;
  long long lo, len;
  sel_segment(p, g, f, b, lo, len);
  const int ntiles = (int)((len + kSelTile - 1) / kSelTile);
  uint32_t* row = p.tilecounts + ((size_t)g * kSelMaxHeavy + pl.theavy[h]) * p.max_tiles;
  uint32_t carry = 0;
  for (int base = 0; base < ntiles; base += kThreads) {
    const int i = base + tid;
    const uint32_t v = i < ntiles ? row[i] : 0u;
    uint32_t tot;
    const uint32_t ex = sel_block_excl_scan<WARPS>(v, warp_tmp, &tot);
    if (i < ntiles) row[i] = carry + ex;
    carry += tot;
  }
  __syncthreads();
  const uint32_t cell_count = carry;  // == pl.cell_count[cell]
  const int j0 = pl.cell_j0[cell], nj = pl.cell_j1[cell] - j0;  // nj <= kSelMaxCuts <= kThreads
  uint8_t* mode = p.tilemode + (size_t)g * p.max_tiles * kSelMaxHeavy + h;
  if (tid < nj) {
    const uint32_t r = pl.cut_r[j0 + tid];  // 1 <= r < cell_count
    int l = 0, hi = ntiles - 1;             // largest tile with row[tile] < r
    while (l < hi) {
      const int mid = (l + hi + 1) >> 1;
      if (row[mid] < r) l = mid; else hi = mid - 1;
    }
    s_tstar[tid] = l;
    s_rho[tid] = r - row[l];
  }
  __syncthreads();
  for (int t = tid; t < ntiles; t += kThreads) {
    int l = 0, hi = nj;  // number of cuts whose tile lies before t
    while (l < hi) {
      const int mid = (l + hi) >> 1;
      if (s_tstar[mid] < t) l = mid + 1; else hi = mid;
    }
    const bool star = l < nj && s_tstar[l] == t;
    mode[(size_t)t * kSelMaxHeavy] = star ? kSelCompact : (uint8_t)(j0 + l);
  }
  // the undecided tiles, in ascending order, share the cell's slot of the side list: the first record of a tile is
  // the number of members in the undecided tiles before it (a scan over the cuts, one thread per cut)
  bool fresh = false;
  uint32_t raw = 0u;
  int t = 0;
  if (tid < nj) {
    t = s_tstar[tid];
    fresh = tid == 0 || s_tstar[tid - 1] != t;
    raw = (t + 1 < ntiles ? row[t + 1] : cell_count) - row[t];  // members of the cut's tile
  }
  uint32_t acc;
  const uint32_t ex = sel_block_excl_scan<WARPS>(fresh ? raw : 0u, warp_tmp, &acc);  // syncs: every row[] read above is done
  if (tid < nj) {
    const uint32_t cur = fresh ? ex : ex - raw;  // a cut that shares its tile with its predecessor: the tile's start
    if (fresh) row[t] = cur;  // first record of this tile inside the cell's slot (what sel_classify reads)
    pl.cut_posoff[j0 + tid] = cur + s_rho[tid];
    RunDesc rd{};
    rd.leader = fresh;
    rd.cell = cell;
    rd.j0 = j0;
    rd.nj = nj;
    rd.start = cur;
    rd.len = raw;
    pl.run[j0 + tid] = rd;
  }
  if (tid == 0) pl.cell_comp[cell] = acc;
  __syncthreads();  // the shared arrays are reused by the next tie group
}

// ---- locate: one block of 1024 threads per segment -----------------------------------------------------------
__global__ void __launch_bounds__(kSelLocThreads, 2) sel_locate(const SelParams p) {
  constexpr int kLocWarps = kSelLocThreads / 32;
  constexpr int kPad = (kSelBins + kLocWarps * 32 - 1) / (kLocWarps * 32) * (kLocWarps * 32);
  constexpr int span = kPad / kLocWarps, rounds = span / 32;
  __shared__ uint32_t excl[kPad + 1];
  __shared__ uint32_t s_wtot[kLocWarps];
  __shared__ uint32_t s_raw[kSelMaxCuts], s_c[kSelMaxCuts], s_r[kSelMaxCuts];
  __shared__ int s_k[kSelMaxCuts], s_T[kSelMaxCuts], s_part[kSelMaxCuts], s_cutcell[kSelMaxCuts];
  __shared__ int s_cellbin[kSelMaxCuts], s_celltiled[kSelMaxCuts], s_cellj0[kSelMaxCuts], s_cellj1[kSelMaxCuts];
  __shared__ uint32_t s_cellcount[kSelMaxCuts];
  __shared__ int s_hbin[kSelMaxHeavy];
  __shared__ int s_ncells, s_wc[4], s_wc2[4], s_tmp[kSelMaxCuts];
  static_assert(kPad * 4 < 44 * 1024, "locate scan layout");
  const int g = blockIdx.x, tid = threadIdx.x, nc = p.num_cuts, lane = tid & 31, warp = tid >> 5;
  int f, b;
  long long lo, len;
  sel_segment(p, g, f, b, lo, len);
  SegPlan& pl = p.plan[g];
  const int nb = max(1, pl.nbins), nh = pl.nheavy;

  // cuts ascending (stable rank by counting)
  if (tid < nc) {
    long long c = p.cuts[(size_t)b * nc + tid];
    c = max(0LL, min(c, len));
    s_raw[tid] = (uint32_t)c;
  }
  if (tid < nh) {  // the bin every heavy tie value owns
    const uint32_t key = pl.heavy[tid];
    s_hbin[tid] = sel_bin(key, p.table[(size_t)g * kSelCoarse + (key >> kSelLowBits)], pl.heavy, nh);
  }
  __syncthreads();
  if (tid < nc) {
    const uint32_t c = s_raw[tid];
    int rank = 0;
    for (int i = 0; i < nc; ++i) rank += (s_raw[i] < c) || (s_raw[i] == c && i < tid);
    s_c[rank] = c;
    s_k[rank] = tid;
  }
  // exclusive prefix of the bin counts: warp w owns bins [span w, span (w + 1)), 32 consecutive bins per round
  {
    const uint32_t* gh = p.hist_f + (size_t)g * kSelBins;
    uint32_t v[rounds];
#pragma unroll
    for (int q = 0; q < rounds; ++q) {
      const int bin = warp * span + q * 32 + lane;
      v[q] = bin < nb ? gh[bin] : 0u;
    }
    uint32_t carry = 0;
#pragma unroll
    for (int q = 0; q < rounds; ++q) {
      uint32_t incl = v[q];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane >= o) incl += u;
      }
      excl[warp * span + q * 32 + lane] = carry + incl - v[q];
      carry += __shfl_sync(FULL_MASK, incl, 31);
    }
    if (lane == 0) s_wtot[warp] = carry;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; ++w) base += s_wtot[w];
    for (int i = lane; i < span; i += 32) excl[warp * span + i] += base;  // excl[b] for b >= nbins = the total
  }
  __syncthreads();

  if (tid < nc) {
    const uint32_t c = s_c[tid];
    int T = 0, part = 0;
    uint32_t r = 0;
    if (c > 0) {
      int l = 0, h = nb - 1;  // smallest bin with excl[bin + 1] >= c
      while (l < h) {
        const int mid = (l + h) >> 1;
        if (excl[mid + 1] >= c) h = mid; else l = mid + 1;
      }
      const uint32_t cnt = excl[l + 1] - excl[l];
      r = c - excl[l];
      if (r >= cnt) { T = l + 1; r = 0; } else { T = l; part = 1; }
    }
    s_T[tid] = T;
    s_part[tid] = part;
    s_r[tid] = r;
  }
  __syncthreads();

  // cells = runs of consecutive cuts inside one bin: numbered by a ballot prefix over the (<= 128) cuts
  static_assert(kSelMaxCuts == 128, "four warps cover the cuts");
  const unsigned le_mask = 0xFFFFFFFFu >> (31 - lane);
  bool fresh = false;
  if (tid < kSelMaxCuts) {
    fresh = tid < nc && s_part[tid] && (tid == 0 || !s_part[tid - 1] || s_T[tid - 1] != s_T[tid]);
    const unsigned bal = __ballot_sync(FULL_MASK, fresh);
    if (lane == 0) s_wc[warp] = __popc(bal);
    s_cutcell[tid] = __popc(bal & le_mask);  // starts at or before this cut inside its warp
    s_cellj1[tid] = 0;
  }
  __syncthreads();
  if (tid < kSelMaxCuts) {
    int before = 0;
    for (int w = 0; w < warp; ++w) before += s_wc[w];
    const int cell = before + s_cutcell[tid] - 1;
    s_cutcell[tid] = (tid < nc && s_part[tid]) ? cell : -1;
    if (fresh) {
      const int bin = s_T[tid];
      const uint32_t cnt = excl[bin + 1] - excl[bin];
      int hk = -1;
      if (cnt > kSelTiledMin)
        for (int h = 0; h < nh; ++h)
          if (s_hbin[h] == bin) hk = h;  // the cell is exactly one tie group
      s_cellbin[cell] = bin;
      s_cellcount[cell] = cnt;
      s_celltiled[cell] = hk;  // heavy-key index for now
      s_cellj0[cell] = tid;
    }
    if (tid == 0) s_ncells = s_wc[0] + s_wc[1] + s_wc[2] + s_wc[3];
  }
  __syncthreads();
  if (tid < nc && s_cutcell[tid] >= 0) atomicMax(&s_cellj1[s_cutcell[tid]], tid + 1);
  if (tid < kSelMaxCuts) {  // number the tie-group cells
    const bool is_t = tid < s_ncells && s_celltiled[tid] >= 0;
    const unsigned bal = __ballot_sync(FULL_MASK, is_t);
    if (lane == 0) s_wc2[warp] = __popc(bal);
    s_tmp[tid] = __popc(bal & le_mask) - 1;
  }
  __syncthreads();
  if (tid < s_ncells) {
    const int cell = tid;
    int ti = -1;
    if (s_celltiled[cell] >= 0) {
      ti = s_tmp[cell];
      for (int w = 0; w < warp; ++w) ti += s_wc2[w];
      pl.tcell[ti] = cell;
      pl.theavy[ti] = s_celltiled[cell];
    }
    s_celltiled[cell] = ti;
    pl.cell_bin[cell] = s_cellbin[cell];
    pl.cell_count[cell] = s_cellcount[cell];
    pl.cell_tiled[cell] = ti;
    pl.cell_j0[cell] = s_cellj0[cell];
    pl.cell_j1[cell] = s_cellj1[cell];
    if (ti < 0) pl.cell_comp[cell] = s_cellcount[cell];
  }
  if (tid == 0) {
    pl.ncells = s_ncells;
    pl.ntiled = s_wc2[0] + s_wc2[1] + s_wc2[2] + s_wc2[3];
  }
  __syncthreads();
  const int ncells = s_ncells;
  if (tid < nc) {
    const int cell = s_cutcell[tid];
    pl.cut_k[tid] = s_k[tid];
    pl.cut_T[tid] = s_T[tid];
    pl.cut_cell[tid] = cell;
    pl.cut_r[tid] = s_r[tid];
    RunDesc rd{};
    if (cell >= 0 && s_celltiled[cell] < 0) {  // the whole cell is one run, resolved by its first cut's block
      pl.cut_posoff[tid] = s_r[tid];
      rd.leader = tid == s_cellj0[cell];
      rd.cell = cell;
      rd.j0 = s_cellj0[cell];
      rd.nj = s_cellj1[cell] - s_cellj0[cell];
      rd.start = 0u;
      rd.len = s_cellcount[cell];
    }
    if (cell < 0 || s_celltiled[cell] < 0) pl.run[tid] = rd;  // tiled cells: sel_plan_tiled_cell writes their runs
  }
  uint16_t* map = p.binmap + (size_t)g * kSelBins;
  for (int bin = tid; bin < nb; bin += kSelLocThreads) {
    int l = 0, h = nc;  // number of cuts with T <= bin
    while (l < h) {
      const int mid = (l + h) >> 1;
      if (s_T[mid] <= bin) l = mid + 1; else h = mid;
    }
    uint16_t v = (uint16_t)l;
    int a = 0, z = ncells;  // is the bin a cell?
    while (a < z) {
      const int mid = (a + z) >> 1;
      if (s_cellbin[mid] < bin) a = mid + 1; else z = mid;
    }
    if (a < ncells && s_cellbin[a] == bin) v = (uint16_t)(kSelCellFlag | a);
    map[bin] = v;
  }
  // the tie groups that hold cuts (their per-tile counts come from the fine histogram pass)
  __syncthreads();  // the plan fields written above are read back by other threads
  const int ntiled = pl.ntiled;
  for (int h = 0; h < ntiled; ++h) sel_plan_tiled_cell<kLocWarps>(p, g, h);
}

// first record of every cell's slot within the segment's side list (s_base[ncells] = all records); every thread of
// the block must call it
__device__ __forceinline__ void sel_cell_bases(const SegPlan& pl, uint32_t* s_base, uint32_t* warp_tmp) {
  const int tid = threadIdx.x, ncells = pl.ncells;
  const uint32_t v = tid < ncells ? pl.cell_comp[tid] : 0u;
  const uint32_t ex = sel_block_excl_scan(v, warp_tmp, nullptr);  // syncs
  if (tid <= ncells && tid <= kSelMaxCuts) s_base[tid] = ex;
  __syncthreads();
}

// ---- classify: class sums of the decided elements, records of the undecided ones ---------------------------
// Class sums without 64-bit shared-memory atomics (those are compare-and-swap loops: ATOMS.CAST.SPIN.64).  Per segment
// (or block) and payload array the largest finite magnitude fixes a scale 2^(E - 150 - kSelWindow) (E = its biased
// exponent); a value whose lowest mantissa bit is a multiple of that scale -- everything within 2^33 of the maximum
// -- is an exact integer q < 2^57, added as three 19-bit limbs with native 32-bit integer reductions (exact,
// order-independent).  Every (payload, limb, class) owns kSelCols words and a lane adds to column lane % kSelCols,
// which spreads the members of one class -- a tie group's thousands of equal-class keys included -- over kSelCols
// words and consecutive classes over the banks.  A word sees at most 2 x 16384 / kSelCols = 4096 adds of < 2^19 (a
// lane adds its own 8 keys of a tile or, in the general path, up to 8 queued ones of its warp: at most 16).  The few
// values below the window, denormals, NaN and inf take the float64 compare-and-swap add.  (16 columns of 20 bits
// halve the bank conflicts but cost a resident block per SM: 157 -> 142 us for 16 views of 800x800.)
constexpr int kSelCols = 8;
constexpr int kSelLimbBits = 19;
constexpr uint32_t kSelLimbMask = (1u << kSelLimbBits) - 1u;
constexpr int kSelWindow = 33;  // q = mantissa << (be - (emax - kSelWindow)) < 2^(24 + 33)
static_assert(2 * (kSelMaxTilesPerBlock * kSelTile / kSelCols) <= (1 << (31 - kSelLimbBits)), "limb words cannot overflow");

constexpr uint16_t kSelTiledFlag = 0x4000u;  // map code of a tie group resolved by tile: look up the tile's code
struct ClassifyShared {
  uint16_t map[kSelBins];           // bin -> class, kSelCellFlag | cell (undecided: a record) or kSelTiledFlag | tiled
                                    // index (-> tmode of the tile); the coarse table is read through L1
  uint32_t heavy[kSelMaxHeavy];
  double sum[kSelMaxCuts + 1][2];
  uint32_t base[kSelMaxCuts + 1];
  int8_t tiled[kSelMaxCuts];        // cell -> tiled index or -1
  uint32_t slot[kSelMaxTilesPerBlock][kSelMaxHeavy];  // first record of (tile, tiled cell) within the segment's side list
  uint32_t cur[kSelMaxTilesPerBlock][kSelMaxHeavy];   // records of (tile, tiled cell) written so far
  uint16_t tmode[kSelMaxTilesPerBlock][kSelMaxHeavy]; // class of the tie group's members in the tile, or kSelCellFlag | cell
  uint32_t warp_tmp[kSelWarps];
  uint32_t vmax[2];                 // bits of the largest finite |payload| of the block
  uint32_t vneg;                    // a payload of the block is negative
  uint32_t rqn[kSelWarps];          // per warp: positions queued for the general path
  uint16_t rq[kSelWarps][kSelTile / kSelWarps];
};
// dynamic shared memory: int limb[2 payloads][3 limbs][num_cuts + 1 classes][kSelCols]; a family with one payload
// array uses the same words as [3 limbs][num_cuts + 1 classes][2 kSelCols]
__host__ __device__ inline size_t sel_limb_words(int num_cuts) { return (size_t)2 * 3 * (num_cuts + 1) * kSelCols; }

struct SelScale {
  int lo_e, width, sh_e;  // q exists iff (unsigned)(be - lo_e) <= width; q = mantissa << (be - sh_e)
};
__device__ __forceinline__ SelScale sel_scale(int emax) {
  SelScale s;
  s.lo_e = emax >= 1 ? max(emax - kSelWindow, 1) : 1000;
  s.width = emax >= 1 ? emax - s.lo_e : 0;
  s.sh_e = emax - kSelWindow;
  return s;
}
// window test of one payload value: `fast` = it has an exact fixed-point image, `slow` = it is outside the window and
// not zero.  SIGNED = false: the caller knows that no payload is negative (-0.0 lands outside the window and adds
// nothing, like +0.0).
template <bool SIGNED>
__device__ __forceinline__ void sel_window(uint32_t bits, const SelScale& sc, uint32_t& sft, bool& fast, bool& slow) {
  const uint32_t be = SIGNED ? ((bits >> 23) & 0xFFu) : (bits >> 23);
  fast = (uint32_t)(be - (uint32_t)sc.lo_e) <= (uint32_t)sc.width;
  sft = (be - (uint32_t)sc.sh_e) & 63u;  // in [0, kSelWindow] when fast (unused otherwise)
  slow = !fast && (bits << 1) != 0u;
}
// the three limbs of q = mant << sft (negated when `neg`); mant = 0 gives three zero limbs
__device__ __forceinline__ void sel_limbs(uint32_t mant, uint32_t sft, bool neg, int (&l)[3]) {
  const unsigned long long q = (unsigned long long)mant << sft;
  const uint32_t qlo = (uint32_t)q, qhi = (uint32_t)(q >> 32);
  l[0] = (int)(qlo & kSelLimbMask);
  l[1] = (int)(__funnelshift_r(qlo, qhi, kSelLimbBits) & kSelLimbMask);
  l[2] = (int)(qhi >> (2 * kSelLimbBits - 32));
  if (neg) {
    l[0] = -l[0];
    l[1] = -l[1];
    l[2] = -l[2];
  }
}
// Three unconditional reductions (a zero limb adds zero): ptxas turns a predicated shared-memory atomic into a
// divergence region of four instructions, which costs more issue slots than the reduction it saves.
// `w`: shared-memory address of the word of (payload, limb 0, class, this lane's column); `stride`: bytes between limbs.
__device__ __forceinline__ void sel_red(uint32_t w, int v) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(w), "r"(v) : "memory");
}
__device__ __forceinline__ void sel_add_limbs(uint32_t w, uint32_t stride, const int (&l)[3]) {
  sel_red(w, l[0]);
  sel_red(w + stride, l[1]);
  sel_red(w + 2u * stride, l[2]);
}

// NONNEG: neither the keys nor the payloads of the segment hold a negative value or (the keys) a NaN -- known from
// the coarse histograms -- which shortens the key transform and the fixed-point split.
template <int NPAY, bool SELF, bool NONNEG>
__device__ __forceinline__ void sel_classify_tiles(const SelParams& p, ClassifyShared& sh, int* limb, int g, int f,
                                                   long long lo, long long len, int t0, const SelScale sc0,
                                                   const SelScale sc1) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const SegPlan& pl = p.plan[g];
  const int nh = pl.nheavy, nt = pl.ntiled;
  const uint32_t* tab = p.table + (size_t)g * kSelCoarse;
  const float* k = p.keys[f] + lo;
  const float* q0 = SELF ? nullptr : p.pay0[f] + lo;
  const float* q1 = NPAY == 2 ? p.pay1[f] + lo : nullptr;
  const long long side = (long long)f * p.total + lo;
  uint32_t* ck = p.ckeys + side;
  uint32_t* ci = p.cidx + side;
  float* c0 = SELF ? nullptr : p.cpay0[f] + lo;
  float* c1 = NPAY == 2 ? p.cpay1[f] + lo : nullptr;
  uint32_t* cursor = p.cursor + (size_t)g * kSelMaxCuts;
  const uint8_t* modes = p.tilemode + (size_t)g * p.max_tiles * kSelMaxHeavy;
  const uint32_t* rows = p.tilecounts + (size_t)g * kSelMaxHeavy * p.max_tiles;
  uint16_t* rq = sh.rq[warp];
  // one payload array: its limbs take the whole table, 16 columns instead of 8 (half the bank conflicts)
  constexpr int kCols = NPAY == 2 ? kSelCols : 2 * kSelCols;
  const uint32_t lstride = (uint32_t)(p.num_cuts + 1) * kCols * 4u;     // bytes between the limbs of a payload
  const uint32_t col0 = smem_u32(limb + (lane & (kCols - 1)));          // this lane's column of payload 0
  const uint32_t col1 = col0 + 3u * lstride;
  constexpr uint32_t kClassBytes = kCols * 4u;

  if (nt > 0) {  // tables of the tie groups for every tile of the block (no barrier inside the tile loop)
    for (int idx = tid; idx < kSelMaxTilesPerBlock * kSelMaxHeavy; idx += kSelThreads) {
      const int tt = idx / kSelMaxHeavy, ti = idx % kSelMaxHeavy;
      const int t = t0 + tt;
      if (ti < nt && tt < p.tiles_per_block && (long long)t * kSelTile < len) {
        const uint8_t md = modes[(size_t)t * kSelMaxHeavy + ti];
        const int cell = pl.tcell[ti];
        sh.slot[tt][ti] = sh.base[cell] + rows[(size_t)pl.theavy[ti] * p.max_tiles + t];
        sh.cur[tt][ti] = 0u;
        sh.tmode[tt][ti] = md == kSelCompact ? (uint16_t)(kSelCellFlag | cell) : (uint16_t)md;
      }
    }
    if (tid < nt) sh.map[pl.cell_bin[pl.tcell[tid]]] = (uint16_t)(kSelTiledFlag | tid);
    __syncthreads();
  }

  for (int tt = 0; tt < p.tiles_per_block; ++tt) {
    const int t = t0 + tt;
    const long long tile_lo = (long long)t * kSelTile;
    if (tile_lo >= len) break;
    const int count = (int)min((long long)kSelTile, len - tile_lo);

    // everything the fast pass does not cover -- an undecided element (a record of its cell), a payload outside the
    // fixed-point window, every element of the segment's last, partial tile -- goes through this
    auto general = [&](int pos) {
      const float kv = k[tile_lo + pos];
      const float v0 = SELF ? kv : q0[tile_lo + pos];
      const float v1 = NPAY == 2 ? q1[tile_lo + pos] : 0.f;
      const uint32_t key = sort_key_from_float(kv);
      uint16_t m = sh.map[sel_bin(key, __ldg(tab + (key >> kSelLowBits)), sh.heavy, nh)];
      if (m & kSelTiledFlag) m = sh.tmode[tt][m & (kSelMaxHeavy - 1)];
      if (m & kSelCellFlag) {  // (any order: records are ranked by (key, index) later)
        const int cell = m & 0x3FFF;
        const int ti = sh.tiled[cell];
        const uint32_t d = ti >= 0 ? sh.slot[tt][ti] + atomicAdd(&sh.cur[tt][ti], 1u)
                                   : sh.base[cell] + atomicAdd(cursor + cell, 1u);
        ck[d] = key;
        ci[d] = (uint32_t)(tile_lo + pos);
        if (!SELF) c0[d] = v0;
        if (NPAY == 2) c1[d] = v1;
      } else {
        int l[3];
        uint32_t sft;
        bool fast, slow;
        const uint32_t b0 = __float_as_uint(v0);
        sel_window<true>(b0, sc0, sft, fast, slow);
        sel_limbs(fast ? ((b0 & 0x7FFFFFu) | 0x800000u) : 0u, sft, (int)b0 < 0, l);
        sel_add_limbs(col0 + (uint32_t)m * kClassBytes, lstride, l);
        if (slow) atomicAdd(&sh.sum[m][0], (double)v0);
        if (NPAY == 2) {
          const uint32_t b1 = __float_as_uint(v1);
          sel_window<true>(b1, sc1, sft, fast, slow);
          sel_limbs(fast ? ((b1 & 0x7FFFFFu) | 0x800000u) : 0u, sft, (int)b1 < 0, l);
          sel_add_limbs(col1 + (uint32_t)m * kClassBytes, lstride, l);
          if (slow) atomicAdd(&sh.sum[m][1], (double)v1);
        }
      }
    };

    if (count == kSelTile) {
      float kf[kSelItems], a0[kSelItems], a1[kSelItems];
#pragma unroll
      for (int i = 0; i < kSelItems; ++i) kf[i] = __ldcs(k + tile_lo + i * kSelThreads + tid);
      if (!SELF) {
#pragma unroll
        for (int i = 0; i < kSelItems; ++i) a0[i] = __ldcs(q0 + tile_lo + i * kSelThreads + tid);
      }
      if (NPAY == 2) {
#pragma unroll
        for (int i = 0; i < kSelItems; ++i) a1[i] = __ldcs(q1 + tile_lo + i * kSelThreads + tid);
      }
      static_assert(kSelItems == 8, "sel_bins8");
      uint32_t key[kSelItems];
      int bin[kSelItems];
      sel_bins8<NONNEG>(kf, tab, sh.heavy, nh, key, bin);
      uint32_t m[kSelItems];
#pragma unroll
      for (int i = 0; i < kSelItems; ++i) m[i] = sh.map[bin[i]];
      if (nt > 0) {  // uniform over the block
#pragma unroll
        for (int i = 0; i < kSelItems; ++i)
          if (m[i] & kSelTiledFlag) m[i] = sh.tmode[tt][m[i] & (kSelMaxHeavy - 1)];
      }
      uint32_t rare = 0u;
#pragma unroll
      for (int i = 0; i < kSelItems; ++i) {
        // an element the fast pass does not own (a record, a payload outside the window) adds nothing here: its
        // mantissas are zeroed, hence its limbs, hence no reduction is issued
        int l[3];
        uint32_t sft0, sft1 = 0u;
        bool fast0, slow0, fast1 = false, slow1 = false;
        const uint32_t b0 = __float_as_uint(SELF ? kf[i] : a0[i]);
        const uint32_t b1 = NPAY == 2 ? __float_as_uint(a1[i]) : 0u;
        sel_window<!NONNEG>(b0, sc0, sft0, fast0, slow0);
        if (NPAY == 2) sel_window<!NONNEG>(b1, sc1, sft1, fast1, slow1);
        const bool ok = !(m[i] & kSelCellFlag) && !slow0 && !slow1;
        rare |= (uint32_t)(!ok) << i;
        const uint32_t w = (m[i] & 0xFFu) * kClassBytes;
        sel_limbs((ok && fast0) ? ((b0 & 0x7FFFFFu) | 0x800000u) : 0u, sft0, !NONNEG && (int)b0 < 0, l);
        sel_add_limbs(col0 + w, lstride, l);
        if (NPAY == 2) {
          sel_limbs((ok && fast1) ? ((b1 & 0x7FFFFFu) | 0x800000u) : 0u, sft1, !NONNEG && (int)b1 < 0, l);
          sel_add_limbs(col1 + w, lstride, l);
        }
      }
      // queue the rare positions per warp, then run them through the general path with every lane busy
      while (rare) {
        const int i = __ffs((int)rare) - 1;
        rare &= rare - 1u;
        rq[atomicAdd(&sh.rqn[warp], 1u)] = (uint16_t)(i * kSelThreads + tid);
      }
      __syncwarp();
      const int nq = (int)sh.rqn[warp];
      for (int e = lane; e < nq; e += 32) general((int)rq[e]);
      __syncwarp();
      if (lane == 0) sh.rqn[warp] = 0u;
    } else {
      for (int pos = tid; pos < count; pos += kSelThreads) general(pos);
    }
  }
}

template <int NPAY, bool SELF>
__device__ __forceinline__ void sel_classify_body(const SelParams& p, ClassifyShared& sh, int* limb, int g, int f,
                                                  int b, long long lo, long long len) {
  const int tid = threadIdx.x, lane = tid & 31;
  SegPlan& pl = p.plan[g];
  const int ncells = pl.ncells, nc = p.num_cuts;
  const int t0 = blockIdx.x * p.tiles_per_block;
  {
    const uint16_t* map = p.binmap + (size_t)g * kSelBins;
    const int nb = pl.nbins;
    for (int i = tid; i < nb; i += kSelThreads) sh.map[i] = map[i];
    if (tid < kSelMaxHeavy) sh.heavy[tid] = pl.heavy[tid];
  }
  const int lwords = (int)sel_limb_words(nc);  // two payloads x 8 columns or one payload x 16 columns
  for (int i = tid; i < (kSelMaxCuts + 1) * 2; i += kSelThreads) (&sh.sum[0][0])[i] = 0.0;
  for (int i = tid; i < lwords / 4; i += kSelThreads) reinterpret_cast<int4*>(limb)[i] = make_int4(0, 0, 0, 0);
  if (tid < 2) sh.vmax[tid] = 0u;
  if (tid == 2) sh.vneg = 0u;
  if (tid < kSelWarps) sh.rqn[tid] = 0u;
  if (tid < kSelMaxCuts) sh.tiled[tid] = tid < ncells ? (int8_t)pl.cell_tiled[tid] : (int8_t)-1;
  sel_cell_bases(pl, sh.base, sh.warp_tmp);
  if (blockIdx.x == 0 && tid < nc) {  // the resolve blocks find their slot without redoing the scan
    const int cell = pl.cut_cell[tid];
    if (cell >= 0) pl.run[tid].base = sh.base[cell];
  }

  // scale of the fixed-point class sums: exponent of the largest finite magnitude of each payload array over
  // the segment -- known from the coarse histogram when the payload is the key array of a family (the AUSE
  // call: the errors are the keys of the two error-sorted families), else the largest of this block's values;
  // likewise whether any payload is negative
  const int src0 = p.pay_src[f][0], src1 = NPAY == 2 ? p.pay_src[f][1] : 0;
  int emax0 = src0 >= 0 ? p.plan[src0 * p.num_views + b].emax : -1;
  int emax1 = NPAY == 2 ? (src1 >= 0 ? p.plan[src1 * p.num_views + b].emax : -1) : 0;
  int neg = (pl.flags & 3) | (src0 >= 0 ? (p.plan[src0 * p.num_views + b].flags & 1) : 0);
  if (NPAY == 2 && src1 >= 0) neg |= p.plan[src1 * p.num_views + b].flags & 1;
  if (emax0 < 0 || emax1 < 0) {  // uniform over the block
    const float* k = p.keys[f] + lo;
    const float* q0 = SELF ? nullptr : p.pay0[f] + lo;
    const float* q1 = NPAY == 2 ? p.pay1[f] + lo : nullptr;
    uint32_t m0 = 0u, m1 = 0u, sg = 0u;
    const long long blk_lo = (long long)t0 * kSelTile;
    const long long blk_hi = min(len, blk_lo + (long long)p.tiles_per_block * kSelTile);
    for (long long i = blk_lo + tid; i < blk_hi; i += kSelThreads) {
      if (emax0 < 0) {
        const uint32_t w = __float_as_uint(SELF ? k[i] : q0[i]);
        const uint32_t b0 = w & 0x7FFFFFFFu;
        if (b0 < 0x7F800000u) m0 = max(m0, b0);
        sg |= w;
      }
      if (NPAY == 2 && emax1 < 0) {
        const uint32_t w = __float_as_uint(q1[i]);
        const uint32_t b1 = w & 0x7FFFFFFFu;
        if (b1 < 0x7F800000u) m1 = max(m1, b1);
        sg |= w;
      }
    }
    m0 = __reduce_max_sync(FULL_MASK, m0);
    m1 = __reduce_max_sync(FULL_MASK, m1);
    sg = __reduce_or_sync(FULL_MASK, sg);
    if (lane == 0) {
      if (m0) atomicMax(&sh.vmax[0], m0);
      if (m1) atomicMax(&sh.vmax[1], m1);
      if (sg >> 31) sh.vneg = 1u;
    }
    __syncthreads();
    if (emax0 < 0) emax0 = (int)(sh.vmax[0] >> 23);
    if (emax1 < 0) emax1 = (int)(sh.vmax[1] >> 23);
    neg |= (int)sh.vneg;
  }
  const SelScale sc0 = sel_scale(emax0), sc1 = sel_scale(emax1);
  if (neg == 0) sel_classify_tiles<NPAY, SELF, true>(p, sh, limb, g, f, lo, len, t0, sc0, sc1);
  else sel_classify_tiles<NPAY, SELF, false>(p, sh, limb, g, f, lo, len, t0, sc0, sc1);

  __syncthreads();  // all adds of the block done: fold the limb columns into the float64 sums
  double* sp = p.spart + ((size_t)g * p.max_blocks + blockIdx.x) * (size_t)(nc + 1) * 2;
  constexpr int kCols = NPAY == 2 ? kSelCols : 2 * kSelCols;
  const int lstride = (nc + 1) * kCols;
  for (int i = tid; i < (nc + 1) * 2; i += kSelThreads) {
    const int cls = i >> 1, pidx = i & 1;
    double v = (&sh.sum[0][0])[i];
    if (pidx < NPAY) {
      const int* w = limb + pidx * 3 * lstride + cls * kCols;
      long long t0s = 0, t1s = 0, t2s = 0;
#pragma unroll
      for (int c = 0; c < kCols; ++c) {
        const int cc = (c + tid) & (kCols - 1);  // threads start at different columns: no bank conflict
        t0s += w[cc];
        t1s += w[lstride + cc];
        t2s += w[2 * lstride + cc];
      }
      // q total = t0 + t1 2^19 + t2 2^38 (|.| < 2^14 2^57 may exceed 64 bits: add as three exactly scaled doubles)
      const int e = (pidx ? emax1 : emax0) - 150 - kSelWindow;
      if (t0s | t1s | t2s)
        v += ldexp((double)t0s, e) + ldexp((double)t1s, e + kSelLimbBits) + ldexp((double)t2s, e + 2 * kSelLimbBits);
    }
    sp[i] = v;
  }
}

__global__ void __launch_bounds__(kSelThreads, 4) sel_classify(const SelParams p) {
  __shared__ ClassifyShared sh;
  extern __shared__ __align__(16) int sel_limb_smem[];
  const int g = blockIdx.y;
  int f, b;
  long long lo, len;
  sel_segment(p, g, f, b, lo, len);
  if ((long long)blockIdx.x * p.tiles_per_block * kSelTile >= len) return;
  if (p.self_payload[f]) sel_classify_body<1, true>(p, sh, sel_limb_smem, g, f, b, lo, len);
  else if (p.pay1[f]) sel_classify_body<2, false>(p, sh, sel_limb_smem, g, f, b, lo, len);
  else sel_classify_body<1, false>(p, sh, sel_limb_smem, g, f, b, lo, len);
}

// ---- resolve: exact class of every record, one block per run of records ------------------------------------
__global__ void __launch_bounds__(kSelResThreads) sel_resolve(const SelParams p) {
  __shared__ unsigned long long s_comp[kSelBrute];
  __shared__ double s_sum[kSelMaxCuts + 1][2];
  __shared__ uint32_t s_pos[kSelMaxCuts];
  __shared__ unsigned long long s_thr[kSelMaxCuts];
  __shared__ uint32_t s_hist[256], s_dstart[256], s_hotd[256];
  __shared__ uint16_t s_list[kSelBrute];
  __shared__ uint32_t s_or[2];
  __shared__ unsigned long long s_prefix;
  __shared__ uint32_t s_rank, s_nl;
  const int j = blockIdx.x, g = blockIdx.y, tid = threadIdx.x;
  const SegPlan& pl = p.plan[g];
  const RunDesc rd = pl.run[j];
  if (!rd.leader) return;
  int f, b;
  long long lo, len;
  sel_segment(p, g, f, b, lo, len);
  const int j0 = rd.j0, nj = rd.nj;
  const uint32_t start = rd.start, rlen = rd.len;
  const long long first = lo + rd.base + start;
  const uint32_t* ck = p.ckeys + (long long)f * p.total + first;
  const uint32_t* ci = p.cidx + (long long)f * p.total + first;
  const bool self = p.self_payload[f] != 0;
  const float* c0 = self ? nullptr : p.cpay0[f] + first;
  const float* c1 = p.pay1[f] ? p.cpay1[f] + first : nullptr;
  for (int i = tid; i < nj; i += kSelResThreads) s_pos[i] = pl.cut_posoff[j0 + i];
  for (int i = tid; i < (nj + 1) * 2; i += kSelResThreads) (&s_sum[0][0])[i] = 0.0;

  if (rlen <= (uint32_t)kSelBrute) {
    // bucket the records by the top 8 bits in which their (key, index) composites differ: a bucket that no cut
    // splits is classified as a whole, the members of the others (a few records) are ranked pairwise
    if (tid == 0) {
      s_or[0] = 0u;
      s_or[1] = 0u;
      s_nl = 0u;
    }
    for (int i = tid; i < 256; i += kSelResThreads) {
      s_hist[i] = 0u;
      s_hotd[i] = 0u;
    }
    for (uint32_t i = tid; i < rlen; i += kSelResThreads)
      s_comp[i] = ((unsigned long long)ck[i] << 32) | ci[i];
    __syncthreads();
    {
      const unsigned long long first_c = s_comp[0];
      uint32_t xl = 0u, xh = 0u;
      for (uint32_t i = tid; i < rlen; i += kSelResThreads) {
        const unsigned long long x = s_comp[i] ^ first_c;
        xl |= (uint32_t)x;
        xh |= (uint32_t)(x >> 32);
      }
      xl = __reduce_or_sync(FULL_MASK, xl);
      xh = __reduce_or_sync(FULL_MASK, xh);
      if ((tid & 31) == 0) {
        if (xl) atomicOr(&s_or[0], xl);
        if (xh) atomicOr(&s_or[1], xh);
      }
    }
    __syncthreads();
    const unsigned long long diff = ((unsigned long long)s_or[1] << 32) | s_or[0];
    const int sh = diff ? max(0, 63 - __clzll((long long)diff) - 7) : 0;
    for (uint32_t i = tid; i < rlen; i += kSelResThreads)
      atomicAdd(&s_hist[(uint32_t)(s_comp[i] >> sh) & 0xFFu], 1u);
    __syncthreads();
    if (tid < 32) {  // exclusive scan of the 256 bucket sizes by one warp
      uint32_t carry = 0;
      for (int r = 0; r < 8; ++r) {
        const uint32_t v = s_hist[r * 32 + tid];
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t u = __shfl_up_sync(FULL_MASK, incl, o);
          if (tid >= o) incl += u;
        }
        s_dstart[r * 32 + tid] = carry + incl - v;
        carry += __shfl_sync(FULL_MASK, incl, 31);
      }
    }
    __syncthreads();
    if (tid < nj) {  // the bucket that holds the last record under the cut is split iff it also holds the next one
      const long long rel = (long long)s_pos[tid] - (long long)start;
      if (rel > 0 && rel < (long long)rlen) {
        int l = 0, h = 255;  // largest digit with dstart <= rel - 1
        while (l < h) {
          const int mid = (l + h + 1) >> 1;
          if ((long long)s_dstart[mid] <= rel - 1) l = mid; else h = mid - 1;
        }
        if (rel < (long long)(s_dstart[l] + s_hist[l])) s_hotd[l] = 1u;
      }
    }
    __syncthreads();
    for (uint32_t i = tid; i < rlen; i += kSelResThreads)
      if (s_hotd[(uint32_t)(s_comp[i] >> sh) & 0xFFu]) s_list[atomicAdd(&s_nl, 1u)] = (uint16_t)i;
    __syncthreads();
    const uint32_t nl = s_nl;
    int cls_lo = 0, cls_hi = 0;  // classes are monotone in the position: the run spans [cls_lo, cls_hi]
    for (int q = 0; q < nj; ++q) {
      cls_lo += s_pos[q] <= start;
      cls_hi += s_pos[q] <= start + rlen - 1u;
    }
    const bool few = cls_hi - cls_lo <= 3;
    double acc0[4] = {0.0, 0.0, 0.0, 0.0}, acc1[4] = {0.0, 0.0, 0.0, 0.0};
    for (uint32_t i = tid; i < rlen; i += kSelResThreads) {
      const unsigned long long mine = s_comp[i];
      const uint32_t d = (uint32_t)(mine >> sh) & 0xFFu;
      uint32_t rank = 0;
      if (s_hotd[d])
        for (uint32_t q = 0; q < nl; ++q) {
          const unsigned long long other = s_comp[s_list[q]];
          rank += (((uint32_t)(other >> sh) & 0xFFu) == d) && other < mine;
        }
      const uint32_t pos = start + s_dstart[d] + rank;  // position inside the cell's slot in (key, index) order
      int cls = 0;
      for (int q = 0; q < nj; ++q) cls += s_pos[q] <= pos;
      const double v0 = self ? (double)order_key_inv((uint32_t)(mine >> 32)) : (double)c0[i];
      const double v1 = c1 ? (double)c1[i] : 0.0;
      if (few) {  // all the threads would hammer the same few shared-memory words: sum in registers first
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          acc0[q] += cls - cls_lo == q ? v0 : 0.0;
          acc1[q] += cls - cls_lo == q ? v1 : 0.0;
        }
      } else {
        atomicAdd(&s_sum[cls][0], v0);
        if (c1) atomicAdd(&s_sum[cls][1], v1);
      }
    }
    if (few) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          acc0[q] += shfl_xor_double(FULL_MASK, acc0[q], o);
          acc1[q] += shfl_xor_double(FULL_MASK, acc1[q], o);
        }
        if ((tid & 31) == 0 && cls_lo + q <= nj) {
          if (acc0[q] != 0.0) atomicAdd(&s_sum[cls_lo + q][0], acc0[q]);
          if (acc1[q] != 0.0) atomicAdd(&s_sum[cls_lo + q][1], acc1[q]);
        }
      }
    }
  } else {
    // a cell too big for shared memory (never a tie-group tile: those hold <= kSelTile records): for every
    // cut of the cell, radix-select the record at position posoff - 1, then class = number of thresholds below
    __syncthreads();
    for (int q = 0; q < nj; ++q) {
      unsigned long long prefix = 0ull, mask = 0ull;
      uint32_t want = s_pos[q] - 1u;  // posoff >= 1
      for (int pass = 7; pass >= 0; --pass) {
        for (int i = tid; i < 256; i += kSelResThreads) s_hist[i] = 0u;
        __syncthreads();
        for (uint32_t i = tid; i < rlen; i += kSelResThreads) {
          const unsigned long long c = ((unsigned long long)ck[i] << 32) | ci[i];
          if ((c & mask) == prefix) atomicAdd(&s_hist[(uint32_t)(c >> (8 * pass)) & 0xFFu], 1u);
        }
        __syncthreads();
        if (tid == 0) {
          uint32_t acc = 0;
          int d = 0;
          for (; d < 255; ++d) {
            if (acc + s_hist[d] > want) break;
            acc += s_hist[d];
          }
          s_rank = want - acc;
          s_prefix = prefix | ((unsigned long long)d << (8 * pass));
        }
        __syncthreads();
        prefix = s_prefix;
        want = s_rank;
        mask |= 0xFFull << (8 * pass);
        __syncthreads();
      }
      if (tid == 0) s_thr[q] = prefix;
    }
    __syncthreads();
    for (uint32_t i = tid; i < rlen; i += kSelResThreads) {
      const unsigned long long c = ((unsigned long long)ck[i] << 32) | ci[i];
      int cls = 0;
      for (int q = 0; q < nj; ++q) cls += s_thr[q] < c;  // excluded from the cuts whose last record precedes it
      const double v0 = self ? (double)order_key_inv(ck[i]) : (double)c0[i];
      atomicAdd(&s_sum[cls][0], v0);
      if (c1) atomicAdd(&s_sum[cls][1], (double)c1[i]);
    }
  }
  __syncthreads();
  double* gs = p.ssum + ((size_t)g * (kSelMaxCuts + 1) + j0) * 2;
  for (int i = tid; i < (nj + 1) * 2; i += kSelResThreads) {
    const double v = (&s_sum[0][0])[i];
    if (v != 0.0) atomicAdd(gs + i, v);
  }
}

// ---- finish: class sums -> sums under every cut ------------------------------------------------------------------
// kSelFinGroups blocks per segment each add a fixed share of the classify blocks' partial sums (block r: partials
// r, r + kSelFinGroups, ..., eight loads in flight); the last of them to arrive adds the group totals and the
// record sums in a fixed order and runs the prefix over the classes.  Deterministic: no sum depends on arrival order.
constexpr int kSelFinGroups = 8;
constexpr int kSelFinStride = (kSelMaxCuts + 1) * 2;
__global__ void __launch_bounds__(kSelThreads) sel_finish(const SelParams p) {
  __shared__ double s_tot[kSelFinStride];
  __shared__ uint32_t s_last;
  const int r = blockIdx.x, g = blockIdx.y, tid = threadIdx.x, nc = p.num_cuts;
  int f, b;
  long long lo, len;
  sel_segment(p, g, f, b, lo, len);
  const int ntiles = (int)((len + kSelTile - 1) / kSelTile);
  const int nblk = (ntiles + p.tiles_per_block - 1) / p.tiles_per_block;
  const int stride = (nc + 1) * 2;
  const double* sp = p.spart + (size_t)g * p.max_blocks * stride;
  double* fp = p.fpart + (size_t)g * kSelFinGroups * kSelFinStride;
  for (int i = tid; i < stride; i += kSelThreads) {
    double s = 0.0;
    for (int blk0 = r; blk0 < nblk; blk0 += 8 * kSelFinGroups) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int blk = blk0 + u * kSelFinGroups;
        v[u] = blk < nblk ? sp[(size_t)blk * stride + i] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    fp[(size_t)r * kSelFinStride + i] = s;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(p.fcount + g, 1u) == (uint32_t)(kSelFinGroups - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const double* gs = p.ssum + (size_t)g * (kSelMaxCuts + 1) * 2;
  for (int i = tid; i < stride; i += kSelThreads) {
    double s = gs[i];
#pragma unroll
    for (int q = 0; q < kSelFinGroups; ++q) s += __ldcg(fp + (size_t)q * kSelFinStride + i);
    s_tot[i] = s;
  }
  __syncthreads();
  // prefix over the classes: warp w scans payload w (32 cuts per round, fixed shuffle order)
  const SegPlan& pl = p.plan[g];
  const int w = tid >> 5, lane = tid & 31;
  if (w < p.npay[f]) {
    double* out = p.out + ((size_t)b * p.num_values + p.row0[f] + w) * nc;
    double carry = 0.0;
    for (int j0 = 0; j0 < nc; j0 += 32) {
      const int j = j0 + lane;
      double v = j < nc ? s_tot[j * 2 + w] : 0.0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double u = shfl_up_double(FULL_MASK, v, o);
        if (lane >= o) v += u;
      }
      v += carry;
      if (j < nc) out[pl.cut_k[j]] = v;
      carry = shfl_double(FULL_MASK, v, 31);
    }
  }
}

struct SelLayout {
  size_t off_zero, zero_bytes, off_histc, off_histf, off_cursor, off_ssum, off_fcount, off_fpart;
  size_t off_table, off_plan, off_binmap, off_tilecounts, off_tilemode, off_spart, off_ckeys, off_cidx, off_cpay,
      total;
  int max_tiles, max_blocks, tiles_per_block, G;
};

static SelLayout sel_layout(int F, int B, int num_cuts, int num_side_arrays, long long total, long long max_len) {
  SelLayout l{};
  l.G = F * B;
  l.max_tiles = (int)((max_len + kSelTile - 1) / kSelTile);
  if (l.max_tiles < 1) l.max_tiles = 1;
  // blocks of the classifying pass: as long as possible (a block stages, zeroes and folds ~60 KB of tables) while one
  // wave of them -- three blocks per SM -- still covers the call; eight tiles each once the call is larger than that
  const int sms = sm_count();
  const long long slots = 4LL * (sms > 0 ? sms : 148);
  long long tpb = ((long long)l.max_tiles * l.G + slots - 1) / slots;
  l.tiles_per_block = (int)(tpb < 1 ? 1 : tpb > kSelMaxTilesPerBlock ? kSelMaxTilesPerBlock : tpb);
  l.max_blocks = (l.max_tiles + l.tiles_per_block - 1) / l.tiles_per_block;
  const size_t G = (size_t)l.G;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    const size_t at = o;
    o = align_up(o + bytes, 256);
    return at;
  };
  l.off_zero = o;
  l.off_histc = take(G * kSelCoarse * sizeof(uint32_t));
  l.off_histf = take(G * kSelBins * sizeof(uint32_t));
  l.off_cursor = take(G * kSelMaxCuts * sizeof(uint32_t));
  l.off_ssum = take(G * (kSelMaxCuts + 1) * 2 * sizeof(double));
  l.off_fcount = take(G * sizeof(uint32_t));
  l.zero_bytes = o - l.off_zero;
  l.off_fpart = take(G * (size_t)kSelFinGroups * kSelFinStride * sizeof(double));
  l.off_table = take(G * kSelCoarse * sizeof(uint32_t));
  l.off_plan = take(G * sizeof(SegPlan));
  l.off_binmap = take(G * kSelBins * sizeof(uint16_t));
  l.off_tilecounts = take(G * (size_t)kSelMaxHeavy * l.max_tiles * sizeof(uint32_t));
  l.off_tilemode = take(G * (size_t)l.max_tiles * kSelMaxHeavy);
  l.off_spart = take(G * (size_t)l.max_blocks * (num_cuts + 1) * 2 * sizeof(double));
  l.off_ckeys = take((size_t)F * total * sizeof(uint32_t));
  l.off_cidx = take((size_t)F * total * sizeof(uint32_t));
  l.off_cpay = take((size_t)num_side_arrays * total * sizeof(float));
  l.total = o;
  return l;
}

}  // namespace ub

extern "C" {

size_t ub_cut_select_sums_workspace_bytes(int32_t num_families, int32_t num_views, int64_t total,
                                          int64_t max_segment_len, int32_t num_cuts) {
  if (num_families < 1 || num_families > ub::kSelMaxFamilies || num_views < 1 || total < 0 ||
      max_segment_len < 0 || num_cuts < 1)
    return 256;
  // worst case: two record payload arrays per family
  return ub::sel_layout(num_families, num_views, num_cuts, 2 * num_families, total, max_segment_len).total;
}

int ub_cut_select_sums(const float* const* keys_host, const float* const* pay0_host,
                       const float* const* pay1_host, int32_t num_families, int32_t num_views,
                       const int64_t* seg_offsets, int64_t total, int64_t max_segment_len, const int64_t* cuts,
                       int32_t num_cuts, double* out_sums, void* workspace, size_t workspace_bytes,
                       void* stream_v) {
  return ub_cut_select_sums_ex(keys_host, pay0_host, pay1_host, num_families, num_views, seg_offsets, total,
                               max_segment_len, cuts, num_cuts, nullptr, out_sums, workspace, workspace_bytes, stream_v);
}

int ub_cut_select_sums_ex(const float* const* keys_host, const float* const* pay0_host,
                          const float* const* pay1_host, int32_t num_families, int32_t num_views,
                          const int64_t* seg_offsets, int64_t total, int64_t max_segment_len, const int64_t* cuts,
                          int32_t num_cuts, const uint32_t* coarse_hist, double* out_sums, void* workspace,
                          size_t workspace_bytes, void* stream_v) {
  using namespace ub;
  UB_REQUIRE(num_families >= 1 && num_families <= kSelMaxFamilies && keys_host && pay0_host, UB_ERR_BAD_ARG,
             "cut_select_sums: num_families must be in [1, %d]", kSelMaxFamilies);
  UB_REQUIRE(num_views >= 1 && (long long)num_views * num_families <= 65535 && seg_offsets != nullptr,
             UB_ERR_BAD_ARG, "cut_select_sums: bad segments");
  UB_REQUIRE(num_cuts >= 1 && num_cuts <= kSelMaxCuts && cuts != nullptr && out_sums != nullptr, UB_ERR_BAD_ARG,
             "cut_select_sums: num_cuts must be in [1, %d]", kSelMaxCuts);
  UB_REQUIRE(total >= 0 && max_segment_len >= 0 && max_segment_len <= total, UB_ERR_BAD_ARG,
             "cut_select_sums: bad total / max_segment_len");
  UB_REQUIRE(max_segment_len <= (1LL << 24), UB_ERR_UNSUPPORTED,
             "cut_select_sums: segments longer than 2^24 keys are not supported (use ub_segmented_sort)");
  SelParams p{};
  int V = 0, side = 0;
  for (int f = 0; f < num_families; ++f) {
    UB_REQUIRE((keys_host[f] != nullptr && pay0_host[f] != nullptr) || total == 0, UB_ERR_BAD_ARG,
               "cut_select_sums: keys / payload of family %d is NULL", f);
    p.keys[f] = keys_host[f];
    p.pay0[f] = pay0_host[f];
    p.pay1[f] = pay1_host ? pay1_host[f] : nullptr;
    p.self_payload[f] = p.pay0[f] == p.keys[f] && p.pay1[f] == nullptr;
    p.npay[f] = p.pay1[f] ? 2 : 1;
    p.pay_src[f][0] = p.pay_src[f][1] = -1;
    for (int f2 = 0; f2 < num_families; ++f2) {
      if (keys_host[f2] == p.pay0[f]) p.pay_src[f][0] = f2;
      if (p.pay1[f] && keys_host[f2] == p.pay1[f]) p.pay_src[f][1] = f2;
    }
    p.row0[f] = V;
    V += p.npay[f];
    if (!p.self_payload[f]) side += p.npay[f];
  }
  const SelLayout lay = sel_layout(num_families, num_views, num_cuts, side, total, max_segment_len);
  UB_REQUIRE(workspace != nullptr && workspace_bytes >= lay.total, UB_ERR_WORKSPACE,
             "cut_select_sums: workspace %zu B < required %zu B", workspace_bytes, lay.total);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  char* ws = static_cast<char*>(workspace);
  p.num_families = num_families;
  p.num_views = num_views;
  p.num_values = V;
  p.num_cuts = num_cuts;
  p.total = total;
  p.max_tiles = lay.max_tiles;
  p.max_blocks = lay.max_blocks;
  p.tiles_per_block = lay.tiles_per_block;
  p.view_offsets = reinterpret_cast<const long long*>(seg_offsets);
  p.cuts = reinterpret_cast<const long long*>(cuts);
  p.out = out_sums;
  p.hist_c = reinterpret_cast<uint32_t*>(ws + lay.off_histc);
  p.hist_f = reinterpret_cast<uint32_t*>(ws + lay.off_histf);
  p.cursor = reinterpret_cast<uint32_t*>(ws + lay.off_cursor);
  p.ssum = reinterpret_cast<double*>(ws + lay.off_ssum);
  p.table = reinterpret_cast<uint32_t*>(ws + lay.off_table);
  p.plan = reinterpret_cast<SegPlan*>(ws + lay.off_plan);
  p.binmap = reinterpret_cast<uint16_t*>(ws + lay.off_binmap);
  p.tilecounts = reinterpret_cast<uint32_t*>(ws + lay.off_tilecounts);
  p.tilemode = reinterpret_cast<uint8_t*>(ws + lay.off_tilemode);
  p.spart = reinterpret_cast<double*>(ws + lay.off_spart);
  p.fpart = reinterpret_cast<double*>(ws + lay.off_fpart);
  p.fcount = reinterpret_cast<uint32_t*>(ws + lay.off_fcount);
  p.ckeys = reinterpret_cast<uint32_t*>(ws + lay.off_ckeys);
  p.cidx = reinterpret_cast<uint32_t*>(ws + lay.off_cidx);
  float* cpay = reinterpret_cast<float*>(ws + lay.off_cpay);
  {
    int s = 0;
    for (int f = 0; f < num_families; ++f) {
      if (p.self_payload[f]) continue;
      p.cpay0[f] = cpay + (size_t)(s++) * total;
      if (p.pay1[f]) p.cpay1[f] = cpay + (size_t)(s++) * total;
    }
  }
  const int G = lay.G;
  if (cudaMemsetAsync(ws + lay.off_zero, 0, lay.zero_bytes, stream) != cudaSuccess)
    return check_launch("cut_select_sums memset");
  const unsigned chunks = (unsigned)((max_segment_len + kSelChunk - 1) / kSelChunk);
  dim3 grid_chunks(chunks < 1 ? 1 : chunks, (unsigned)G);
  dim3 grid_blocks((unsigned)lay.max_blocks, (unsigned)G);
  dim3 grid_cells((unsigned)num_cuts, (unsigned)G);
  if (coarse_hist != nullptr)
    p.hist_c = const_cast<uint32_t*>(coarse_hist);  // read-only from here on (sel_alloc)
  else
    sel_coarse_hist<<<grid_chunks, kSelThreads, 0, stream>>>(p);
  sel_alloc<<<G, kSelAllocThreads, 0, stream>>>(p);
  sel_fine_hist<<<grid_chunks, kSelThreads, 0, stream>>>(p);
  sel_locate<<<G, kSelLocThreads, 0, stream>>>(p);
  {
    const size_t limb_bytes = sel_limb_words(num_cuts) * sizeof(int);
    // opt in to > 48 KB of shared memory, once per device (an idempotent kernel attribute, not library state)
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      if (cudaFuncSetAttribute(sel_classify, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)(sel_limb_words(kSelMaxCuts) * sizeof(int))) != cudaSuccess)
        return check_launch("cut_select_sums smem attribute");
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    sel_classify<<<grid_blocks, kSelThreads, limb_bytes, stream>>>(p);
  }
  sel_resolve<<<grid_cells, kSelResThreads, 0, stream>>>(p);
  sel_finish<<<dim3(kSelFinGroups, (unsigned)G), kSelThreads, 0, stream>>>(p);
  return check_launch("cut_select_sums");
}

}  // extern "C"
