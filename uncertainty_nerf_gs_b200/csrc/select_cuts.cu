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
//              that needs neither the key range nor a sample
//   fine       histogram over the fine bins
//   locate     prefix over the fine bins; every cut falls into one bin (its "cell") at a residual rank r;
//              a bin that holds no cut gets a class = number of cuts that exclude it
//   cells      per (cell, 2048-key tile) counts + key range of every cell
//   plan       a big cell whose keys are all equal (a tie group) is resolved by index: stable order inside
//              it is the element order, so the tile where the running count crosses r follows from the
//              per-tile counts and only that tile's members stay undecided; other cells stay undecided whole
//   classify   one pass over keys + payloads: decided elements add their payload to the float64 sum of their
//              class; undecided ones (typically 1-3 % of the keys) go to their cell's slot of a side list
//              as (key, index, payloads) records
//   resolve    per cell (or per undecided tile of a tie group): exact rank of every record by (key, index),
//              hence its class; a radix select over the records takes over for cells too big to rank
//              pairwise (several tie groups inside one fine bin -- correct, only slower)
//   finish     class sums -> prefix over the classes = the sums under every cut.
//
// The sets of elements under every cut are exactly those of torch.sort(stable=True) (same key transform:
// -0.0 == +0.0, NaN last, ties by index), so the sums equal the sort path's up to float64 summation order.
#include "ub_common.cuh"

namespace ub {

constexpr int kSelThreads = 256;
constexpr int kSelWarps = kSelThreads / 32;
constexpr int kSelCoarse = 4096;   // top 12 key bits
constexpr int kSelLowBits = 20;
constexpr uint32_t kSelLowMask = (1u << kSelLowBits) - 1u;
constexpr int kSelFine = 8192;     // fine bins per segment (upper bound by construction)
constexpr int kSelMaxHeavy = 16;   // tie values that get a bin of their own
constexpr int kSelBins = kSelFine + 2 * kSelMaxHeavy;
constexpr int kSelSample = 1024;   // keys sampled per segment to find the heavy tie values
constexpr int kSelHeavyHits = 8;   // sample hits that make a key heavy (~0.8 % of the segment)
constexpr int kSelTile = 2048;     // counting / compaction tile
constexpr int kSelItems = kSelTile / kSelThreads;
constexpr int kSelMaxTilesPerBlock = 16;
constexpr int kSelChunk = 8192;    // keys per block in the histogram kernels
constexpr int kSelMaxCuts = 128;
constexpr int kSelMaxFamilies = 4;
constexpr uint32_t kSelTiledMin = 2 * kSelTile;  // tie groups above this size are resolved by tile
constexpr int kSelBrute = 2048;    // records ranked pairwise in shared memory
constexpr uint16_t kSelCellFlag = 0x8000u;
constexpr uint8_t kSelCompact = 0xFFu;

struct SegPlan {
  int nfine, nbins, ncells, nheavy;
  uint32_t heavy[kSelMaxHeavy];      // ascending
  int cut_k[kSelMaxCuts];            // original index of the j-th smallest cut
  int cut_T[kSelMaxCuts];            // fine bins < T are under the cut
  int cut_cell[kSelMaxCuts];         // cell that holds the cut, -1 if the cut is a bin boundary
  uint32_t cut_r[kSelMaxCuts];       // elements of the cell under the cut (stable order)
  uint32_t cut_posoff[kSelMaxCuts];  // records of the cell's slot under the cut (in (key, index) order)
  int cut_leader[kSelMaxCuts];       // the cut's block resolves a run of records
  uint32_t cut_run_start[kSelMaxCuts], cut_run_len[kSelMaxCuts];
  int cell_bin[kSelMaxCuts];
  uint32_t cell_count[kSelMaxCuts];
  uint32_t cell_comp[kSelMaxCuts];   // records in the cell's slot
  int cell_j0[kSelMaxCuts], cell_j1[kSelMaxCuts];  // cuts [j0, j1) lie inside the cell
  int cell_tiled[kSelMaxCuts];
};

struct SelParams {
  const float* keys[kSelMaxFamilies];
  const float* pay0[kSelMaxFamilies];
  const float* pay1[kSelMaxFamilies];  // NULL: one payload
  int self_payload[kSelMaxFamilies];   // payload 0 is the key itself
  int row0[kSelMaxFamilies];           // first output row of the family
  int npay[kSelMaxFamilies];
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
  uint32_t* cell_mm;      // [G][kSelMaxCuts][2] {max(~u), max(u)}         zeroed
  double* ssum;           // [G][kSelMaxCuts + 1][2] class sums of the records  zeroed
  uint32_t* table;        // [G][kSelCoarse]: (first fine bin << 5) | shift of the low 20 key bits
  SegPlan* plan;          // [G]
  uint16_t* binmap;       // [G][kSelBins]: class, or kSelCellFlag | cell
  uint32_t* tilecounts;   // [G][num_cuts][max_tiles]; after sel_plan_cells: first record of (cell, tile)
  uint8_t* tilemode;      // [G][max_tiles][kSelMaxCuts]: class of the cell's members in the tile, or kSelCompact
  double* spart;          // [G][max_blocks][num_cuts + 1][2]
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
  for (int w = 0; w < kSelWarps; ++w) {
    const uint32_t t = warp_tmp[w];
    if (w < warp) base += t;
    tot += t;
  }
  if (total_out) *total_out = tot;
  __syncthreads();
  return base + incl - v;
}

// per-block copy of the segment's binning tables
struct BinTables {
  uint16_t map[kSelBins];
  uint32_t tab[kSelCoarse];
  uint32_t heavy[kSelMaxHeavy];
};

__device__ __forceinline__ void sel_load_tables(const SelParams& p, int g, const SegPlan& pl, BinTables& bt,
                                                bool with_map) {
  const uint32_t* tab = p.table + (size_t)g * kSelCoarse;
  for (int i = threadIdx.x; i < kSelCoarse; i += kSelThreads) bt.tab[i] = tab[i];
  if (with_map) {
    const uint16_t* map = p.binmap + (size_t)g * kSelBins;
    const int nb = pl.nbins;
    for (int i = threadIdx.x; i < nb; i += kSelThreads) bt.map[i] = map[i];
  }
  if (threadIdx.x < kSelMaxHeavy) bt.heavy[threadIdx.x] = pl.heavy[threadIdx.x];
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

// ---- fine-bin allocation + heavy tie values: one block per segment ----------------------------------------
__global__ void __launch_bounds__(kSelThreads) sel_alloc(const SelParams p) {
  __shared__ uint32_t warp_tmp[kSelWarps];
  __shared__ uint32_t s_sample[kSelSample];
  const int g = blockIdx.x, tid = threadIdx.x;
  int f, b;
  long long lo, len;
  sel_segment(p, g, f, b, lo, len);
  SegPlan& pl = p.plan[g];
  const uint32_t target = (uint32_t)max(1LL, (len + 2047) / 2048);
  constexpr int per = kSelCoarse / kSelThreads;
  const uint32_t* gh = p.hist_c + (size_t)g * kSelCoarse;
  uint32_t nsub[per], lg[per], sum = 0;
#pragma unroll
  for (int i = 0; i < per; ++i) {
    const uint32_t cnt = gh[tid * per + i];
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
  uint32_t run = sel_block_excl_scan(sum, warp_tmp, &tot);
  uint32_t* tab = p.table + (size_t)g * kSelCoarse;
#pragma unroll
  for (int i = 0; i < per; ++i) {
    tab[tid * per + i] = (run << 5) | ((uint32_t)kSelLowBits - lg[i]);
    run += nsub[i];
  }
  // heavy tie values from a sorted sample: a key that fills >= 8 of 1024 evenly spaced probes.  A tie group
  // next to other keys of the same fine bin would make its cell "large and not one tie group" (the slow
  // radix-select corner of sel_resolve); owning a bin makes it a tie-group cell, resolved by index.
  const int ns = (int)min((long long)kSelSample, len);
  const float* k = p.keys[f] + lo;
  for (int i = tid; i < kSelSample; i += kSelThreads)
    s_sample[i] = i < ns ? sort_key_from_float(__ldg(k + (long long)i * len / ns)) : 0xFFFFFFFFu;
  __syncthreads();
  for (int size = 2; size <= kSelSample; size <<= 1) {      // bitonic sort, ascending
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < kSelSample / 2; i += kSelThreads) {
        const int a = 2 * i - (i & (stride - 1)), c = a + stride;
        const bool up = (a & size) == 0;
        const uint32_t x = s_sample[a], y = s_sample[c];
        if ((x > y) == up) {
          s_sample[a] = y;
          s_sample[c] = x;
        }
      }
      __syncthreads();
    }
  }
  __shared__ int s_nh;
  __shared__ uint32_t s_hv[kSelMaxHeavy];
  for (int thresh = kSelHeavyHits;; thresh *= 2) {  // at most kSelMaxHeavy values: raise the bar until they fit
    if (tid == 0) s_nh = 0;
    __syncthreads();
    for (int i = tid; i < ns; i += kSelThreads) {
      const uint32_t v = s_sample[i];
      const bool first = i == 0 || s_sample[i - 1] != v;
      if (first && i + thresh - 1 < ns && s_sample[i + thresh - 1] == v) {
        const int slot = atomicAdd(&s_nh, 1);
        if (slot < kSelMaxHeavy) s_hv[slot] = v;
      }
    }
    __syncthreads();
    if (s_nh <= kSelMaxHeavy) break;
    __syncthreads();
  }
  const int nh = s_nh;
  if (tid < kSelMaxHeavy) pl.heavy[tid] = tid < nh ? s_hv[tid] : 0xFFFFFFFFu;
  if (tid == 0) {
    pl.nheavy = nh;
    pl.nfine = (int)tot;  // <= kSelFine: sum pow2ceil(cnt / target) <= 4096 + 2 n / target
    pl.nbins = (int)tot + 2 * nh;
  }
}

// ---- fine histogram --------------------------------------------------------------------------------------
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
  for (int i = threadIdx.x; i < nb; i += kSelThreads) h[i] = 0u;
  if (threadIdx.x < kSelMaxHeavy) s_heavy[threadIdx.x] = pl.heavy[threadIdx.x];
  __syncthreads();
  const int count = (int)min((long long)kSelChunk, len - start);
  const float* k = p.keys[f] + lo + start;
  const uint32_t* tab = p.table + (size_t)g * kSelCoarse;
  constexpr int per = kSelChunk / kSelThreads;
  for (int base = 0; base < per; base += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldg(k + min((base + i) * kSelThreads + (int)threadIdx.x, count - 1));
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if ((base + i) * kSelThreads + (int)threadIdx.x < count) {
        const uint32_t u = sort_key_from_float(v[i]);
        atomicAdd(&h[sel_bin(u, __ldg(tab + (u >> kSelLowBits)), s_heavy, nh)], 1u);
      }
  }
  __syncthreads();
  uint32_t* gh = p.hist_f + (size_t)g * kSelBins;
  for (int i = threadIdx.x; i < nb; i += kSelThreads) {
    const uint32_t v = h[i];
    if (v) atomicAdd(gh + i, v);
  }
}

// ---- locate: one block per segment -----------------------------------------------------------------------
__global__ void __launch_bounds__(kSelThreads) sel_locate(const SelParams p) {
  constexpr int kPad = (kSelBins + 2047) / 2048 * 2048;  // 8 warps x (a multiple of 8 rounds of 32 bins)
  __shared__ uint32_t excl[kPad + 1];
  __shared__ uint32_t s_wtot[kSelWarps];
  __shared__ uint32_t s_raw[kSelMaxCuts], s_c[kSelMaxCuts], s_r[kSelMaxCuts];
  __shared__ int s_k[kSelMaxCuts], s_T[kSelMaxCuts], s_part[kSelMaxCuts], s_cutcell[kSelMaxCuts];
  __shared__ int s_cellbin[kSelMaxCuts];
  __shared__ int s_ncells;
  static_assert(kPad % (kSelWarps * 32 * 8) == 0 && kPad * 4 < 44 * 1024, "locate scan layout");
  const int g = blockIdx.x, tid = threadIdx.x, nc = p.num_cuts, lane = tid & 31, warp = tid >> 5;
  int f, b;
  long long lo, len;
  sel_segment(p, g, f, b, lo, len);
  SegPlan& pl = p.plan[g];
  const int nb = max(1, pl.nbins);

  // cuts ascending (stable rank by counting)
  if (tid < nc) {
    long long c = p.cuts[(size_t)b * nc + tid];
    c = max(0LL, min(c, len));
    s_raw[tid] = (uint32_t)c;
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
    constexpr int span = kPad / kSelWarps, rounds = span / 32;
    const uint32_t* gh = p.hist_f + (size_t)g * kSelBins;
    uint32_t carry = 0;
    for (int r0 = 0; r0 < rounds; r0 += 8) {
      uint32_t v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int bin = warp * span + (r0 + q) * 32 + lane;
        v[q] = bin < nb ? gh[bin] : 0u;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint32_t incl = v[q];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t u = __shfl_up_sync(FULL_MASK, incl, o);
          if (lane >= o) incl += u;
        }
        excl[warp * span + (r0 + q) * 32 + lane] = carry + incl - v[q];
        carry += __shfl_sync(FULL_MASK, incl, 31);
      }
    }
    if (lane == 0) s_wtot[warp] = carry;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kSelWarps; ++w) {
      if (w < warp) base += s_wtot[w];
      tot += s_wtot[w];
    }
    for (int i = lane; i < span; i += 32) excl[warp * span + i] += base;
    if (tid == 0) excl[kPad] = tot;
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

  if (tid == 0) {
    int ncells = 0;
    for (int j = 0; j < nc; ++j) {
      if (!s_part[j]) { s_cutcell[j] = -1; continue; }
      const bool fresh = j == 0 || !s_part[j - 1] || s_T[j - 1] != s_T[j];
      if (fresh) {
        s_cellbin[ncells] = s_T[j];
        pl.cell_bin[ncells] = s_T[j];
        pl.cell_count[ncells] = excl[s_T[j] + 1] - excl[s_T[j]];
        pl.cell_j0[ncells] = j;
        ++ncells;
      }
      s_cutcell[j] = ncells - 1;
      pl.cell_j1[ncells - 1] = j + 1;
    }
    s_ncells = ncells;
    pl.ncells = ncells;
  }
  __syncthreads();
  if (tid < nc) {
    pl.cut_k[tid] = s_k[tid];
    pl.cut_T[tid] = s_T[tid];
    pl.cut_cell[tid] = s_cutcell[tid];
    pl.cut_r[tid] = s_r[tid];
    pl.cut_posoff[tid] = 0;
    pl.cut_leader[tid] = 0;
  }
  const int ncells = s_ncells;
  uint16_t* map = p.binmap + (size_t)g * kSelBins;
  for (int bin = tid; bin < nb; bin += kSelThreads) {
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
}

// ---- per (cell, tile) counts and the key range of every cell ---------------------------------------------
__global__ void __launch_bounds__(kSelThreads) sel_cell_counts(const SelParams p) {
  __shared__ BinTables bt;
  __shared__ uint32_t s_cnt[kSelMaxCuts];
  __shared__ uint32_t s_mm[kSelMaxCuts][2];
  const int g = blockIdx.y, tid = threadIdx.x;
  const SegPlan& pl = p.plan[g];
  const int ncells = pl.ncells;
  if (ncells == 0) return;
  int f, b;
  long long lo, len;
  sel_segment(p, g, f, b, lo, len);
  const int t0 = blockIdx.x * p.tiles_per_block;
  if ((long long)t0 * kSelTile >= len) return;
  const int nh = pl.nheavy;
  sel_load_tables(p, g, pl, bt, true);
  if (tid < kSelMaxCuts) {
    s_cnt[tid] = 0u;
    s_mm[tid][0] = 0u;
    s_mm[tid][1] = 0u;
  }
  __syncthreads();
  const float* k = p.keys[f] + lo;
  for (int tt = 0; tt < p.tiles_per_block; ++tt) {
    const int t = t0 + tt;
    const long long tile_lo = (long long)t * kSelTile;
    if (tile_lo >= len) break;
    const int count = (int)min((long long)kSelTile, len - tile_lo);
    float v[kSelItems];
#pragma unroll
    for (int i = 0; i < kSelItems; ++i) v[i] = __ldg(k + tile_lo + min(i * kSelThreads + tid, count - 1));
#pragma unroll
    for (int i = 0; i < kSelItems; ++i) {
      if (i * kSelThreads + tid < count) {
        const uint32_t key = sort_key_from_float(v[i]);
        const uint16_t m = bt.map[sel_bin(key, bt.tab[key >> kSelLowBits], bt.heavy, nh)];
        if (m & kSelCellFlag) {
          const int cell = m & 0x7FFF;
          atomicAdd(&s_cnt[cell], 1u);
          atomicMax(&s_mm[cell][0], ~key);
          atomicMax(&s_mm[cell][1], key);
        }
      }
    }
    __syncthreads();
    if (tid < ncells) {
      p.tilecounts[((size_t)g * p.num_cuts + tid) * p.max_tiles + t] = s_cnt[tid];
      s_cnt[tid] = 0u;
    }
    __syncthreads();
  }
  if (tid < ncells && (s_mm[tid][0] | s_mm[tid][1])) {
    atomicMax(p.cell_mm + ((size_t)g * kSelMaxCuts + tid) * 2, s_mm[tid][0]);
    atomicMax(p.cell_mm + ((size_t)g * kSelMaxCuts + tid) * 2 + 1, s_mm[tid][1]);
  }
}

// ---- plan: one block per (cell, segment) ---------------------------------------------------------------
__global__ void __launch_bounds__(kSelThreads) sel_plan_cells(const SelParams p) {
  __shared__ uint32_t warp_tmp[kSelWarps];
  __shared__ int s_tstar[kSelMaxCuts];
  __shared__ uint32_t s_rho[kSelMaxCuts];
  const int cell = blockIdx.x, g = blockIdx.y, tid = threadIdx.x;
  SegPlan& pl = p.plan[g];
  if (cell >= pl.ncells) return;
  int f, b;
  long long lo, len;
  sel_segment(p, g, f, b, lo, len);
  const int ntiles = (int)((len + kSelTile - 1) / kSelTile);
  uint32_t* row = p.tilecounts + ((size_t)g * p.num_cuts + cell) * p.max_tiles;
  uint32_t carry = 0;
  for (int base = 0; base < ntiles; base += kSelThreads) {
    const int i = base + tid;
    const uint32_t v = i < ntiles ? row[i] : 0u;
    uint32_t tot;
    const uint32_t ex = sel_block_excl_scan(v, warp_tmp, &tot);
    if (i < ntiles) row[i] = carry + ex;
    carry += tot;
  }
  __syncthreads();
  const uint32_t cell_count = carry;  // == pl.cell_count[cell]
  const uint32_t* mm = p.cell_mm + ((size_t)g * kSelMaxCuts + cell) * 2;
  const bool pure = (~mm[0]) == mm[1];
  const bool tiled = pure && cell_count > kSelTiledMin;
  const int j0 = pl.cell_j0[cell], nj = pl.cell_j1[cell] - j0;
  uint8_t* mode = p.tilemode + (size_t)g * p.max_tiles * kSelMaxCuts + cell;
  if (!tiled) {
    for (int t = tid; t < ntiles; t += kSelThreads) mode[(size_t)t * kSelMaxCuts] = kSelCompact;
    for (int jj = tid; jj < nj; jj += kSelThreads) {
      pl.cut_posoff[j0 + jj] = pl.cut_r[j0 + jj];
      pl.cut_leader[j0 + jj] = jj == 0;
      pl.cut_run_start[j0 + jj] = 0u;
      pl.cut_run_len[j0 + jj] = cell_count;
    }
    if (tid == 0) {
      pl.cell_comp[cell] = cell_count;
      pl.cell_tiled[cell] = 0;
    }
    return;
  }
  for (int jj = tid; jj < nj; jj += kSelThreads) {
    const uint32_t r = pl.cut_r[j0 + jj];  // 1 <= r < cell_count
    int l = 0, h = ntiles - 1;             // largest tile with row[tile] < r
    while (l < h) {
      const int mid = (l + h + 1) >> 1;
      if (row[mid] < r) l = mid; else h = mid - 1;
    }
    s_tstar[jj] = l;
    s_rho[jj] = r - row[l];
  }
  __syncthreads();
  for (int t = tid; t < ntiles; t += kSelThreads) {
    int l = 0, h = nj;  // number of cuts whose tile lies before t
    while (l < h) {
      const int mid = (l + h) >> 1;
      if (s_tstar[mid] < t) l = mid + 1; else h = mid;
    }
    const bool star = l < nj && s_tstar[l] == t;
    mode[(size_t)t * kSelMaxCuts] = star ? kSelCompact : (uint8_t)(j0 + l);
  }
  if (tid == 0) {
    uint32_t acc = 0, cur = 0, raw = 0;
    for (int jj = 0; jj < nj; ++jj) {
      const bool fresh = jj == 0 || s_tstar[jj] != s_tstar[jj - 1];
      if (fresh) {
        const int t = s_tstar[jj];
        cur = acc;
        raw = (t + 1 < ntiles ? row[t + 1] : cell_count) - row[t];
        acc += raw;
        row[t] = cur;  // first record of this tile inside the cell's slot (tiles ascend: row[t + 1] is still a prefix)
      }
      pl.cut_posoff[j0 + jj] = cur + s_rho[jj];
      pl.cut_leader[j0 + jj] = fresh;
      pl.cut_run_start[j0 + jj] = cur;
      pl.cut_run_len[j0 + jj] = raw;
    }
    pl.cell_comp[cell] = acc;
    pl.cell_tiled[cell] = 1;
  }
}

// first record of every cell's slot within the segment's side list (s_base[ncells] = all records) and the
// biggest tie-group cell; every thread of the block must call it
__device__ __forceinline__ void sel_cell_bases(const SegPlan& pl, uint32_t* s_base, uint32_t* warp_tmp,
                                               uint32_t* s_hot) {
  const int tid = threadIdx.x, ncells = pl.ncells;
  if (tid == 0) *s_hot = 0u;
  const uint32_t v = tid < ncells ? pl.cell_comp[tid] : 0u;
  uint32_t tot;
  const uint32_t ex = sel_block_excl_scan(v, warp_tmp, &tot);  // syncs
  if (tid <= ncells && tid <= kSelMaxCuts) s_base[tid] = ex;
  if (tid < ncells && pl.cell_tiled[tid]) atomicMax(s_hot, (pl.cell_count[tid] << 7) | (uint32_t)tid);
  __syncthreads();
}

// ---- classify: class sums of the decided elements, records of the undecided ones ---------------------------
struct ClassifyShared {
  BinTables bt;
  double sum[kSelMaxCuts + 1][2];
  uint32_t base[kSelMaxCuts + 1];
  uint32_t slot[kSelMaxCuts];  // first record of (cell, tile) within the segment's side list
  uint32_t cur[kSelMaxCuts];
  uint32_t warp_tmp[kSelWarps];
  uint32_t hot;
  int limb[kSelMaxCuts + 1][2][3];  // per-tile fixed-point class sums (3 x 16 bits), native 32-bit atomics
  uint32_t vmax[2];                 // bits of the largest finite |payload| of the tile
  uint8_t mode[kSelMaxCuts];
};

// Class sums without 64-bit shared-memory atomics (those are compare-and-swap loops: ATOMS.CAST.SPIN.64).  Per
// tile and payload array the largest finite magnitude fixes a scale 2^(E - 174) (E = its biased exponent); a value
// whose lowest mantissa bit is a multiple of that scale -- everything within 2^24 of the maximum -- is an exact
// integer q < 2^49 and is added as three 16-bit limbs with native integer atomics (exact, order-independent);
// the few values below that, denormals, NaN and inf take the float64 compare-and-swap add.
__device__ __forceinline__ void sel_add_payload(ClassifyShared& sh, int cls, int pidx, float v, int emax) {
  const uint32_t bits = __float_as_uint(v);
  const int be = (int)((bits >> 23) & 0xFFu);
  if ((bits << 1) == 0u) return;  // +-0 adds nothing
  if (be != 0 && be != 255 && be + 24 >= emax) {
    long long q = (long long)((bits & 0x7FFFFFu) | 0x800000u) << (be + 24 - emax);
    if (bits >> 31) q = -q;
    atomicAdd(&sh.limb[cls][pidx][0], (int)(q & 0xFFFF));
    atomicAdd(&sh.limb[cls][pidx][1], (int)((q >> 16) & 0xFFFF));
    atomicAdd(&sh.limb[cls][pidx][2], (int)(q >> 32));
  } else {
    atomicAdd(&sh.sum[cls][pidx], (double)v);
  }
}

template <int NPAY, bool SELF>
__device__ __forceinline__ void sel_classify_body(const SelParams& p, ClassifyShared& sh, int g, int f,
                                                  long long lo, long long len) {
  const int tid = threadIdx.x, lane = tid & 31;
  const SegPlan& pl = p.plan[g];
  const int ncells = pl.ncells, nc = p.num_cuts, nh = pl.nheavy;
  const int t0 = blockIdx.x * p.tiles_per_block;
  sel_load_tables(p, g, pl, sh.bt, true);
  for (int i = tid; i < (kSelMaxCuts + 1) * 2; i += kSelThreads) (&sh.sum[0][0])[i] = 0.0;
  for (int i = tid; i < (kSelMaxCuts + 1) * 6; i += kSelThreads) (&sh.limb[0][0][0])[i] = 0;
  if (tid < 2) sh.vmax[tid] = 0u;
  sel_cell_bases(pl, sh.base, sh.warp_tmp, &sh.hot);
  const int hot = sh.hot ? (int)(sh.hot & 127u) : -1;
  const float* k = p.keys[f] + lo;
  const float* q0 = SELF ? nullptr : p.pay0[f] + lo;
  const float* q1 = NPAY == 2 ? p.pay1[f] + lo : nullptr;
  const long long side = (long long)f * p.total + lo;
  uint32_t* ck = p.ckeys + side;
  uint32_t* ci = p.cidx + side;
  float* c0 = SELF ? nullptr : p.cpay0[f] + lo;
  float* c1 = NPAY == 2 ? p.cpay1[f] + lo : nullptr;
  const uint8_t* modes = p.tilemode + (size_t)g * p.max_tiles * kSelMaxCuts;
  const uint32_t* rows = p.tilecounts + (size_t)g * p.num_cuts * p.max_tiles;

  for (int tt = 0; tt < p.tiles_per_block; ++tt) {
    const int t = t0 + tt;
    const long long tile_lo = (long long)t * kSelTile;
    if (tile_lo >= len) break;
    const int count = (int)min((long long)kSelTile, len - tile_lo);
    __syncthreads();  // previous tile done with the per-tile tables
    if (tid < ncells) {
      sh.mode[tid] = modes[(size_t)t * kSelMaxCuts + tid];
      sh.slot[tid] = sh.base[tid] + rows[(size_t)tid * p.max_tiles + t];
      sh.cur[tid] = 0u;
    }
    float kf[kSelItems], a0[kSelItems], a1[kSelItems];
#pragma unroll
    for (int i = 0; i < kSelItems; ++i) kf[i] = __ldcs(k + tile_lo + min(i * kSelThreads + tid, count - 1));
    if (!SELF) {
#pragma unroll
      for (int i = 0; i < kSelItems; ++i) a0[i] = __ldcs(q0 + tile_lo + min(i * kSelThreads + tid, count - 1));
    }
    if (NPAY == 2) {
#pragma unroll
      for (int i = 0; i < kSelItems; ++i) a1[i] = __ldcs(q1 + tile_lo + min(i * kSelThreads + tid, count - 1));
    }
    {  // largest finite payload magnitudes of the tile
      uint32_t m0 = 0u, m1 = 0u;
#pragma unroll
      for (int i = 0; i < kSelItems; ++i) {
        if (i * kSelThreads + tid < count) {
          const uint32_t b0 = __float_as_uint(SELF ? kf[i] : a0[i]) & 0x7FFFFFFFu;
          if (b0 < 0x7F800000u) m0 = max(m0, b0);
          if (NPAY == 2) {
            const uint32_t b1 = __float_as_uint(a1[i]) & 0x7FFFFFFFu;
            if (b1 < 0x7F800000u) m1 = max(m1, b1);
          }
        }
      }
      m0 = __reduce_max_sync(FULL_MASK, m0);
      if (NPAY == 2) m1 = __reduce_max_sync(FULL_MASK, m1);
      if (lane == 0) {
        if (m0) atomicMax(&sh.vmax[0], m0);
        if (NPAY == 2 && m1) atomicMax(&sh.vmax[1], m1);
      }
    }
    __syncthreads();  // per-tile tables and maxima visible
    const int emax0 = (int)(sh.vmax[0] >> 23), emax1 = (int)(sh.vmax[1] >> 23);
    const int hot_cls = hot >= 0 ? (int)sh.mode[hot] : (int)kSelCompact;
    double hot0 = 0.0, hot1 = 0.0;
#pragma unroll
    for (int i = 0; i < kSelItems; ++i) {
      const int pos = i * kSelThreads + tid;
      if (pos < count) {
        const uint32_t key = sort_key_from_float(kf[i]);
        const uint16_t m = sh.bt.map[sel_bin(key, sh.bt.tab[key >> kSelLowBits], sh.bt.heavy, nh)];
        int cls = m;
        int cell = -1;
        if (m & kSelCellFlag) {
          cell = m & 0x7FFF;
          cls = sh.mode[cell];
        }
        if (cls == (int)kSelCompact) {  // undecided: a record in the cell's slot (order inside the tile is free:
          const uint32_t d = sh.slot[cell] + atomicAdd(&sh.cur[cell], 1u);  // records are ranked by (key, index))
          ck[d] = key;
          ci[d] = (uint32_t)(tile_lo + pos);
          if (!SELF) c0[d] = a0[i];
          if (NPAY == 2) c1[d] = a1[i];
        } else {
          const double v0 = (double)(SELF ? kf[i] : a0[i]);
          if (cell >= 0 && cell == hot) {  // the big tie group: one class per tile, summed in registers
            hot0 += v0;
            if (NPAY == 2) hot1 += (double)a1[i];
          } else {
            sel_add_payload(sh, cls, 0, SELF ? kf[i] : a0[i], emax0);
            if (NPAY == 2) sel_add_payload(sh, cls, 1, a1[i], emax1);
          }
        }
      }
    }
    if (hot >= 0 && hot_cls != (int)kSelCompact) {  // uniform over the block
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        hot0 += shfl_xor_double(FULL_MASK, hot0, o);
        if (NPAY == 2) hot1 += shfl_xor_double(FULL_MASK, hot1, o);
      }
      if (lane == 0) {
        atomicAdd(&sh.sum[hot_cls][0], hot0);
        if (NPAY == 2) atomicAdd(&sh.sum[hot_cls][1], hot1);
      }
    }
    __syncthreads();  // all adds of the tile done: fold the limbs into the float64 sums
    for (int i = tid; i < (nc + 1) * 2; i += kSelThreads) {
      int* l = &sh.limb[0][0][0] + i * 3;
      const long long tot = ((long long)l[2] << 32) + ((long long)l[1] << 16) + (long long)l[0];
      if (tot != 0) (&sh.sum[0][0])[i] += ldexp((double)tot, (int)(sh.vmax[i & 1] >> 23) - 174);
      l[0] = 0;
      l[1] = 0;
      l[2] = 0;
    }
    __syncthreads();
    if (tid < 2) sh.vmax[tid] = 0u;
  }
  __syncthreads();
  double* sp = p.spart + ((size_t)g * p.max_blocks + blockIdx.x) * (size_t)(nc + 1) * 2;
  for (int i = tid; i < (nc + 1) * 2; i += kSelThreads) sp[i] = (&sh.sum[0][0])[i];
}

__global__ void __launch_bounds__(kSelThreads) sel_classify(const SelParams p) {
  __shared__ ClassifyShared sh;
  const int g = blockIdx.y;
  int f, b;
  long long lo, len;
  sel_segment(p, g, f, b, lo, len);
  if ((long long)blockIdx.x * p.tiles_per_block * kSelTile >= len) return;
  if (p.self_payload[f]) sel_classify_body<1, true>(p, sh, g, f, lo, len);
  else if (p.pay1[f]) sel_classify_body<2, false>(p, sh, g, f, lo, len);
  else sel_classify_body<1, false>(p, sh, g, f, lo, len);
}

// ---- resolve: exact class of every record, one block per run of records ------------------------------------
__global__ void __launch_bounds__(kSelThreads) sel_resolve(const SelParams p) {
  __shared__ unsigned long long s_comp[kSelBrute];
  __shared__ double s_sum[kSelMaxCuts + 1][2];
  __shared__ uint32_t s_pos[kSelMaxCuts];
  __shared__ unsigned long long s_thr[kSelMaxCuts];
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_base[kSelMaxCuts + 1];
  __shared__ uint32_t warp_tmp[kSelWarps];
  __shared__ unsigned long long s_prefix;
  __shared__ uint32_t s_rank, s_hot, s_nl;
  __shared__ uint32_t s_or[2], s_dstart[256], s_hotd[256];
  __shared__ uint16_t s_list[kSelBrute];
  const int j = blockIdx.x, g = blockIdx.y, tid = threadIdx.x;
  const SegPlan& pl = p.plan[g];
  if (j >= p.num_cuts) return;
  const int cell = pl.cut_cell[j];
  if (cell < 0 || !pl.cut_leader[j]) return;
  int f, b;
  long long lo, len;
  sel_segment(p, g, f, b, lo, len);
  sel_cell_bases(pl, s_base, warp_tmp, &s_hot);
  const int j0 = pl.cell_j0[cell], nj = pl.cell_j1[cell] - j0;
  const uint32_t start = pl.cut_run_start[j], rlen = pl.cut_run_len[j];
  const long long first = lo + s_base[cell] + start;
  const uint32_t* ck = p.ckeys + (long long)f * p.total + first;
  const uint32_t* ci = p.cidx + (long long)f * p.total + first;
  const bool self = p.self_payload[f] != 0;
  const float* c0 = self ? nullptr : p.cpay0[f] + first;
  const float* c1 = p.pay1[f] ? p.cpay1[f] + first : nullptr;
  for (int i = tid; i < nj; i += kSelThreads) s_pos[i] = pl.cut_posoff[j0 + i];
  for (int i = tid; i < (nj + 1) * 2; i += kSelThreads) (&s_sum[0][0])[i] = 0.0;
  __syncthreads();

  if (rlen <= (uint32_t)kSelBrute) {
    // bucket the records by the top 8 bits in which their (key, index) composites differ: a bucket that no cut
    // splits is classified as a whole, the members of the others (a few records) are ranked pairwise
    if (tid == 0) {
      s_or[0] = 0u;
      s_or[1] = 0u;
      s_nl = 0u;
    }
    s_hist[tid] = 0u;
    s_hotd[tid] = 0u;
    for (uint32_t i = tid; i < rlen; i += kSelThreads)
      s_comp[i] = ((unsigned long long)ck[i] << 32) | ci[i];
    __syncthreads();
    {
      const unsigned long long c0 = s_comp[0];
      uint32_t xl = 0u, xh = 0u;
      for (uint32_t i = tid; i < rlen; i += kSelThreads) {
        const unsigned long long x = s_comp[i] ^ c0;
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
    for (uint32_t i = tid; i < rlen; i += kSelThreads) atomicAdd(&s_hist[(uint32_t)(s_comp[i] >> sh) & 0xFFu], 1u);
    __syncthreads();
    {
      const uint32_t cnt = s_hist[tid];
      const uint32_t ex = sel_block_excl_scan(cnt, warp_tmp, nullptr);
      s_dstart[tid] = ex;
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
    for (uint32_t i = tid; i < rlen; i += kSelThreads)
      if (s_hotd[(uint32_t)(s_comp[i] >> sh) & 0xFFu]) s_list[atomicAdd(&s_nl, 1u)] = (uint16_t)i;
    __syncthreads();
    const uint32_t nl = s_nl;
    for (uint32_t i = tid; i < rlen; i += kSelThreads) {
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
      atomicAdd(&s_sum[cls][0], v0);
      if (c1) atomicAdd(&s_sum[cls][1], (double)c1[i]);
    }
  } else {
    // a cell too big to rank pairwise (never a tie-group tile: those hold <= kSelTile records): for every
    // cut of the cell, radix-select the record at position posoff - 1, then class = number of thresholds below
    for (int q = 0; q < nj; ++q) {
      unsigned long long prefix = 0ull, mask = 0ull;
      uint32_t want = s_pos[q] - 1u;  // posoff >= 1
      for (int pass = 7; pass >= 0; --pass) {
        s_hist[tid] = 0u;
        __syncthreads();
        for (uint32_t i = tid; i < rlen; i += kSelThreads) {
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
    for (uint32_t i = tid; i < rlen; i += kSelThreads) {
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
  for (int i = tid; i < (nj + 1) * 2; i += kSelThreads) {
    const double v = (&s_sum[0][0])[i];
    if (v != 0.0) atomicAdd(gs + i, v);
  }
}

// ---- finish: class sums -> sums under every cut, one block per segment ---------------------------------------
__global__ void __launch_bounds__(kSelThreads) sel_finish(const SelParams p) {
  __shared__ double s_tot[(kSelMaxCuts + 1) * 2];
  const int g = blockIdx.x, tid = threadIdx.x, nc = p.num_cuts;
  int f, b;
  long long lo, len;
  sel_segment(p, g, f, b, lo, len);
  const int ntiles = (int)((len + kSelTile - 1) / kSelTile);
  const int nblk = (ntiles + p.tiles_per_block - 1) / p.tiles_per_block;
  const size_t stride = (size_t)(nc + 1) * 2;
  const double* sp = p.spart + (size_t)g * p.max_blocks * stride;
  const double* gs = p.ssum + (size_t)g * (kSelMaxCuts + 1) * 2;
  for (int i = tid; i < (int)stride; i += kSelThreads) {
    double s = gs[i];
    for (int blk = 0; blk < nblk; ++blk) s += sp[(size_t)blk * stride + i];
    s_tot[i] = s;
  }
  __syncthreads();
  const SegPlan& pl = p.plan[g];
  if (tid < p.npay[f]) {
    double run = 0.0;
    double* out = p.out + ((size_t)b * p.num_values + p.row0[f] + tid) * nc;
    for (int j = 0; j < nc; ++j) {
      run += s_tot[j * 2 + tid];
      out[pl.cut_k[j]] = run;
    }
  }
}

struct SelLayout {
  size_t off_zero, zero_bytes, off_histc, off_histf, off_cellmm, off_ssum;
  size_t off_table, off_plan, off_binmap, off_tilecounts, off_tilemode, off_spart, off_ckeys, off_cidx, off_cpay,
      total;
  int max_tiles, max_blocks, tiles_per_block, G;
};

static SelLayout sel_layout(int F, int B, int num_cuts, int num_side_arrays, long long total, long long max_len) {
  SelLayout l{};
  l.G = F * B;
  l.max_tiles = (int)((max_len + kSelTile - 1) / kSelTile);
  if (l.max_tiles < 1) l.max_tiles = 1;
  // blocks of the counting / classifying passes: enough of them to fill the device, each as long as that
  // allows (a block stages 50 KB of binning tables)
  long long tpb = (long long)l.max_tiles * l.G / 1184;
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
  l.off_cellmm = take(G * kSelMaxCuts * 2 * sizeof(uint32_t));
  l.off_ssum = take(G * (kSelMaxCuts + 1) * 2 * sizeof(double));
  l.zero_bytes = o - l.off_zero;
  l.off_table = take(G * kSelCoarse * sizeof(uint32_t));
  l.off_plan = take(G * sizeof(SegPlan));
  l.off_binmap = take(G * kSelBins * sizeof(uint16_t));
  l.off_tilecounts = take(G * (size_t)num_cuts * l.max_tiles * sizeof(uint32_t));
  l.off_tilemode = take(G * (size_t)l.max_tiles * kSelMaxCuts);
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
  p.cell_mm = reinterpret_cast<uint32_t*>(ws + lay.off_cellmm);
  p.ssum = reinterpret_cast<double*>(ws + lay.off_ssum);
  p.table = reinterpret_cast<uint32_t*>(ws + lay.off_table);
  p.plan = reinterpret_cast<SegPlan*>(ws + lay.off_plan);
  p.binmap = reinterpret_cast<uint16_t*>(ws + lay.off_binmap);
  p.tilecounts = reinterpret_cast<uint32_t*>(ws + lay.off_tilecounts);
  p.tilemode = reinterpret_cast<uint8_t*>(ws + lay.off_tilemode);
  p.spart = reinterpret_cast<double*>(ws + lay.off_spart);
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
  sel_coarse_hist<<<grid_chunks, kSelThreads, 0, stream>>>(p);
  sel_alloc<<<G, kSelThreads, 0, stream>>>(p);
  sel_fine_hist<<<grid_chunks, kSelThreads, 0, stream>>>(p);
  sel_locate<<<G, kSelThreads, 0, stream>>>(p);
  sel_cell_counts<<<grid_blocks, kSelThreads, 0, stream>>>(p);
  sel_plan_cells<<<grid_cells, kSelThreads, 0, stream>>>(p);
  sel_classify<<<grid_blocks, kSelThreads, 0, stream>>>(p);
  sel_resolve<<<grid_cells, kSelThreads, 0, stream>>>(p);
  sel_finish<<<G, kSelThreads, 0, stream>>>(p);
  return check_launch("cut_select_sums");
}

}  // extern "C"
