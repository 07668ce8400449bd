// (C2') AUSE cut-point sums by a multi-cut radix select instead of a full sort.
//
// Reference arithmetic (file:line under /root/reference/nerfuncertainty):
//   metrics/ause.py:10, 15-20     err sorted ascending, mean of the first int((1-r) n) values, 100 ratios
//   metrics/ause.py:25-26, 29-34  idx = sort(unc); err[idx]; the same 100 prefix means
// The curves only need  P_k = sum of the payload over the FIRST c_k ELEMENTS of the stable ascending order
// of the keys  at the 100 cut counts c_k; the permutation itself is never returned.  So instead of four
// scatter passes over every key (ub_segmented_sort) the keys are only *classified*:
//
//   range      per segment: min / max of the order-preserving uint32 key
//   hist       4096 bins over [min, max] (adaptive shift), counts only
//   locate     prefix over the bins; every cut falls into one bin ("cell") at a residual rank r; bins that
//              hold no cut get a class = number of cuts that exclude them
//   cells      per (cell, 2048-key tile) counts + key range of each cell
//   plan       a cell whose keys are all equal (a tie group) is resolved by index: stable order inside it is
//              the element order, so the tile where the running count crosses r is found from the per-tile
//              counts and only that tile's members stay undecided; other cells stay undecided as a whole
//   classify   one pass over keys + payloads: decided elements add their payload to their class sum
//              (float64), undecided ones are compacted, in element order, into a small side list
//   finish     the side lists go through the existing stable segmented sort + cut-point prefix sums; the
//              class sums are prefix-summed over the classes and added.
//
// The sets of elements under every cut are exactly those of torch.sort(stable=True) (same key transform:
// -0.0 == +0.0, NaN last, ties by index), so the sums equal the sort path's up to float64 summation order.
// Typical images leave 2-5 % of the keys undecided; a segment whose cells are large and not tie groups
// degrades to sorting those cells, never to a wrong answer.
#include "ub_common.cuh"
#include "sort_internal.cuh"

namespace ub {

constexpr int kSelThreads = 256;
constexpr int kSelWarps = kSelThreads / 32;
constexpr int kSelBins = 4096;
constexpr int kSelBinBits = 12;
constexpr int kSelTile = 2048;  // counting / compaction tile
constexpr int kSelItems = kSelTile / kSelThreads;
constexpr int kSelTilesPerBlock = 4;
constexpr int kSelChunk = 8192;  // keys per block in the range / histogram kernels
constexpr int kSelMaxCuts = 128;
constexpr int kSelMaxFamilies = 4;
constexpr uint32_t kSelTiledMin = 2 * kSelTile;  // tie groups above this size are resolved by tile
constexpr uint16_t kSelCellFlag = 0x8000u;
constexpr uint8_t kSelCompact = 0xFFu;

struct SegPlan {
  uint32_t umin;
  int shift, nb, ncells, hot_cell;
  int cut_k[kSelMaxCuts];          // original index of the j-th smallest cut
  int cut_T[kSelMaxCuts];          // bins < T are under the cut
  int cut_cell[kSelMaxCuts];       // cell that holds the cut, -1 if the cut is a bin boundary
  uint32_t cut_r[kSelMaxCuts];     // elements of the cell under the cut (stable order)
  uint32_t cut_rho[kSelMaxCuts];   // tiled cells: elements of tile t* under the cut
  uint32_t cut_posoff[kSelMaxCuts];  // position of the cut inside the cell's part of the side list
  int cell_bin[kSelMaxCuts];
  uint32_t cell_count[kSelMaxCuts];
  uint32_t cell_comp[kSelMaxCuts];   // elements of the cell in the side list
  int cell_j0[kSelMaxCuts], cell_j1[kSelMaxCuts];  // cuts [j0, j1) lie inside the cell
  int cell_tiled[kSelMaxCuts];
};

struct SelParams {
  const float* keys[kSelMaxFamilies];
  const float* pay0[kSelMaxFamilies];
  const float* pay1[kSelMaxFamilies];  // NULL: one payload
  int self_payload[kSelMaxFamilies];   // payload 0 is the key itself: nothing but keys is compacted
  int row0[kSelMaxFamilies];           // first output row of the family
  int npay[kSelMaxFamilies];
  float* cpay0[kSelMaxFamilies];       // side lists [total] (NULL when self_payload)
  float* cpay1[kSelMaxFamilies];
  int num_families, num_views, num_values, num_cuts;
  long long total;
  int max_tiles, max_blocks;
  const long long* view_offsets;  // [B + 1]
  const long long* cuts;          // [B][num_cuts]
  uint32_t* range;        // [G][2] {max(~u), max(u)}                     zeroed
  uint32_t* hist;         // [G][kSelBins]                                  zeroed
  uint32_t* cell_mm;      // [G][kSelMaxCuts][2] {max(~u), max(u)}          zeroed
  uint32_t* tile_ncomp;   // [G][max_tiles]                                 zeroed
  long long* seg_offsets; // [G + 1] slot of every segment in the [F * total] side arrays
  SegPlan* plan;          // [G]
  uint16_t* binmap;       // [G][kSelBins]: class, or kSelCellFlag | cell
  uint32_t* tilecounts;   // [G][num_cuts][max_tiles]; exclusive prefix over the tiles after sel_plan_cells
  uint8_t* tilemode;      // [G][max_tiles][kSelMaxCuts]: class of the cell's members in the tile, or kSelCompact
  uint32_t* tilebase;     // [G][max_tiles] first side-list position of the tile
  long long* complen;     // [G] side-list length
  long long* lens_v;      // [V][B]
  long long* pos_v;       // [V][B][num_cuts] cut positions in the sorted side list
  double* spart;          // [G][max_blocks][num_cuts + 1][2]
  double* coarse;         // [B][V][num_cuts]
  float* ckeys;           // [F * total]
};

__device__ __forceinline__ void sel_segment(const SelParams& p, int g, int& f, int& b, long long& lo,
                                            long long& len) {
  f = g / p.num_views;
  b = g - f * p.num_views;
  lo = p.view_offsets[b];
  len = p.view_offsets[b + 1] - lo;
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

// ---- range -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSelThreads) sel_range(const SelParams p) {
  const int g = blockIdx.y;
  int f, b;
  long long lo, len;
  sel_segment(p, g, f, b, lo, len);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    p.seg_offsets[g] = (long long)f * p.total + lo;
    if (g == (int)gridDim.y - 1) p.seg_offsets[g + 1] = (long long)p.num_families * p.total;
  }
  const long long start = (long long)blockIdx.x * kSelChunk;
  if (start >= len) return;
  const int count = (int)min((long long)kSelChunk, len - start);
  const float* k = p.keys[f] + lo + start;
  uint32_t mn = 0xFFFFFFFFu, mx = 0u;
#pragma unroll 8
  for (int i = threadIdx.x; i < count; i += kSelThreads) {
    const uint32_t u = sort_key_from_float(__ldg(k + i));
    mn = min(mn, u);
    mx = max(mx, u);
  }
  mn = __reduce_min_sync(FULL_MASK, mn);
  mx = __reduce_max_sync(FULL_MASK, mx);
  if ((threadIdx.x & 31) == 0 && mn <= mx) {
    atomicMax(p.range + 2 * g, ~mn);
    atomicMax(p.range + 2 * g + 1, mx);
  }
}

__device__ __forceinline__ void sel_binning(const SelParams& p, int g, uint32_t& umin, int& shift, int& nb) {
  umin = ~p.range[2 * g];
  const uint32_t umax = p.range[2 * g + 1];
  const uint32_t span = umax >= umin ? umax - umin : 0u;
  const int bits = 32 - __clz(span);
  shift = max(0, bits - kSelBinBits);
  nb = (int)(span >> shift) + 1;
}

// ---- histogram -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSelThreads) sel_hist(const SelParams p) {
  __shared__ uint32_t h[kSelBins];
  const int g = blockIdx.y;
  int f, b;
  long long lo, len;
  sel_segment(p, g, f, b, lo, len);
  const long long start = (long long)blockIdx.x * kSelChunk;
  if (start >= len) return;
  uint32_t umin;
  int shift, nb;
  sel_binning(p, g, umin, shift, nb);
  for (int i = threadIdx.x; i < nb; i += kSelThreads) h[i] = 0u;
  __syncthreads();
  const int count = (int)min((long long)kSelChunk, len - start);
  const float* k = p.keys[f] + lo + start;
#pragma unroll 8
  for (int i = threadIdx.x; i < count; i += kSelThreads) {
    const uint32_t u = sort_key_from_float(__ldg(k + i));
    atomicAdd(&h[(u - umin) >> shift], 1u);
  }
  __syncthreads();
  uint32_t* gh = p.hist + (size_t)g * kSelBins;
  for (int i = threadIdx.x; i < nb; i += kSelThreads) {
    const uint32_t v = h[i];
    if (v) atomicAdd(gh + i, v);
  }
}

// ---- locate: one block per segment -----------------------------------------------------------------
__global__ void __launch_bounds__(kSelThreads) sel_locate(const SelParams p) {
  __shared__ uint32_t excl[kSelBins + 1];
  __shared__ uint32_t warp_tmp[kSelWarps];
  __shared__ uint32_t s_raw[kSelMaxCuts], s_c[kSelMaxCuts], s_r[kSelMaxCuts];
  __shared__ int s_k[kSelMaxCuts], s_T[kSelMaxCuts], s_part[kSelMaxCuts], s_cutcell[kSelMaxCuts];
  __shared__ int s_cellbin[kSelMaxCuts];
  __shared__ int s_ncells;
  const int g = blockIdx.x, tid = threadIdx.x, nc = p.num_cuts;
  int f, b;
  long long lo, len;
  sel_segment(p, g, f, b, lo, len);
  uint32_t umin;
  int shift, nb;
  sel_binning(p, g, umin, shift, nb);
  if (len == 0) nb = 1;
  SegPlan& pl = p.plan[g];

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
  // exclusive prefix of the bin counts: thread t owns bins [16 t, 16 t + 16)
  {
    constexpr int per = kSelBins / kSelThreads;
    const uint32_t* gh = p.hist + (size_t)g * kSelBins;
    uint32_t loc[per];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < per; ++i) {
      const int bin = tid * per + i;
      loc[i] = bin < nb ? gh[bin] : 0u;
      sum += loc[i];
    }
    uint32_t run = sel_block_excl_scan(sum, warp_tmp, nullptr);
#pragma unroll
    for (int i = 0; i < per; ++i) {
      excl[tid * per + i] = run;
      run += loc[i];
    }
    if (tid == kSelThreads - 1) excl[kSelBins] = run;
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
    pl.umin = umin;
    pl.shift = shift;
    pl.nb = nb;
    pl.ncells = ncells;
    pl.hot_cell = -1;
  }
  __syncthreads();
  if (tid < nc) {
    pl.cut_k[tid] = s_k[tid];
    pl.cut_T[tid] = s_T[tid];
    pl.cut_cell[tid] = s_cutcell[tid];
    pl.cut_r[tid] = s_r[tid];
    pl.cut_rho[tid] = 0;
    pl.cut_posoff[tid] = 0;
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

// ---- per (cell, tile) counts and the key range of every cell -----------------------------------------
__global__ void __launch_bounds__(kSelThreads) sel_cell_counts(const SelParams p) {
  __shared__ uint16_t s_map[kSelBins];
  __shared__ uint32_t s_cnt[kSelMaxCuts];
  __shared__ uint32_t s_mm[kSelMaxCuts][2];
  const int g = blockIdx.y, tid = threadIdx.x;
  const SegPlan& pl = p.plan[g];
  const int ncells = pl.ncells;
  if (ncells == 0) return;
  int f, b;
  long long lo, len;
  sel_segment(p, g, f, b, lo, len);
  const int t0 = blockIdx.x * kSelTilesPerBlock;
  if ((long long)t0 * kSelTile >= len) return;
  const uint32_t umin = pl.umin;
  const int shift = pl.shift, nb = pl.nb;
  const uint16_t* map = p.binmap + (size_t)g * kSelBins;
  for (int i = tid; i < nb; i += kSelThreads) s_map[i] = map[i];
  if (tid < kSelMaxCuts) {
    s_cnt[tid] = 0u;
    s_mm[tid][0] = 0u;
    s_mm[tid][1] = 0u;
  }
  __syncthreads();
  const float* k = p.keys[f] + lo;
  for (int tt = 0; tt < kSelTilesPerBlock; ++tt) {
    const int t = t0 + tt;
    const long long tile_lo = (long long)t * kSelTile;
    if (tile_lo >= len) break;
    const int count = (int)min((long long)kSelTile, len - tile_lo);
    uint32_t u[kSelItems];
#pragma unroll
    for (int i = 0; i < kSelItems; ++i)
      u[i] = __float_as_uint(__ldg(k + tile_lo + min(i * kSelThreads + tid, count - 1)));
#pragma unroll
    for (int i = 0; i < kSelItems; ++i) {
      if (i * kSelThreads + tid < count) {
        const uint32_t key = sort_key_from_float(__uint_as_float(u[i]));
        const uint16_t m = s_map[(key - umin) >> shift];
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

// ---- plan, part 1: one block per (cell, segment) ------------------------------------------------------
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
  uint32_t* ncomp = p.tile_ncomp + (size_t)g * p.max_tiles;
  if (!tiled) {
    for (int t = tid; t < ntiles; t += kSelThreads) {
      mode[(size_t)t * kSelMaxCuts] = kSelCompact;
      const uint32_t raw = (t + 1 < ntiles ? row[t + 1] : cell_count) - row[t];
      if (raw) atomicAdd(ncomp + t, raw);
    }
    for (int jj = tid; jj < nj; jj += kSelThreads) pl.cut_posoff[j0 + jj] = pl.cut_r[j0 + jj];
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
    pl.cut_rho[j0 + jj] = r - row[l];
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
    if (star) {
      const uint32_t raw = (t + 1 < ntiles ? row[t + 1] : cell_count) - row[t];
      if (raw) atomicAdd(ncomp + t, raw);
    }
  }
  if (tid == 0) {
    uint32_t acc = 0, cur = 0;
    for (int jj = 0; jj < nj; ++jj) {
      if (jj == 0 || s_tstar[jj] != s_tstar[jj - 1]) {
        const int t = s_tstar[jj];
        cur = acc;
        acc += (t + 1 < ntiles ? row[t + 1] : cell_count) - row[t];
      }
      pl.cut_posoff[j0 + jj] = cur + s_rho[jj];
    }
    pl.cell_comp[cell] = acc;
    pl.cell_tiled[cell] = 1;
  }
}

// ---- plan, part 2: one block per segment --------------------------------------------------------------
__global__ void __launch_bounds__(kSelThreads) sel_plan_tiles(const SelParams p) {
  __shared__ uint32_t warp_tmp[kSelWarps];
  __shared__ uint32_t s_cbase[kSelMaxCuts + 1];
  const int g = blockIdx.x, tid = threadIdx.x, nc = p.num_cuts;
  SegPlan& pl = p.plan[g];
  int f, b;
  long long lo, len;
  sel_segment(p, g, f, b, lo, len);
  const int ncells = pl.ncells;
  if (tid == 0) {
    uint32_t acc = 0, best = 0;
    int hot = -1;
    for (int c = 0; c < ncells; ++c) {
      s_cbase[c] = acc;
      acc += pl.cell_comp[c];
      if (pl.cell_tiled[c] && pl.cell_count[c] > best) {
        best = pl.cell_count[c];
        hot = c;
      }
    }
    s_cbase[ncells] = acc;
    p.complen[g] = acc;
    pl.hot_cell = hot;
    for (int r = 0; r < p.npay[f]; ++r) p.lens_v[(size_t)(p.row0[f] + r) * p.num_views + b] = acc;
  }
  __syncthreads();
  for (int j = tid; j < nc; j += kSelThreads) {
    long long pos;
    const int cell = pl.cut_cell[j];
    if (cell >= 0) {
      pos = (long long)s_cbase[cell] + pl.cut_posoff[j];
    } else {
      const int T = pl.cut_T[j];
      int l = 0, h = ncells;  // cells with bin < T
      while (l < h) {
        const int mid = (l + h) >> 1;
        if (pl.cell_bin[mid] < T) l = mid + 1; else h = mid;
      }
      pos = s_cbase[l];
    }
    for (int r = 0; r < p.npay[f]; ++r)
      p.pos_v[((size_t)(p.row0[f] + r) * p.num_views + b) * nc + pl.cut_k[j]] = pos;
  }
  const int ntiles = (int)((len + kSelTile - 1) / kSelTile);
  const uint32_t* ncomp = p.tile_ncomp + (size_t)g * p.max_tiles;
  uint32_t* tb = p.tilebase + (size_t)g * p.max_tiles;
  uint32_t carry = 0;
  for (int base = 0; base < ntiles; base += kSelThreads) {
    const int i = base + tid;
    const uint32_t v = i < ntiles ? ncomp[i] : 0u;
    uint32_t tot;
    const uint32_t ex = sel_block_excl_scan(v, warp_tmp, &tot);
    if (i < ntiles) tb[i] = carry + ex;
    carry += tot;
  }
}

// ---- classify: class sums + stable compaction of the undecided elements ---------------------------------
template <int NPAY, bool SELF>
__device__ __forceinline__ void sel_classify_body(const SelParams& p, int g, int f, long long lo, long long len) {
  __shared__ uint16_t s_map[kSelBins];
  __shared__ double s_sum[kSelMaxCuts + 1][2];
  __shared__ uint8_t s_mode[kSelMaxCuts];
  __shared__ uint32_t s_wtot[kSelWarps];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const SegPlan& pl = p.plan[g];
  const int ncells = pl.ncells, nc = p.num_cuts, hot = pl.hot_cell;
  const uint32_t umin = pl.umin;
  const int shift = pl.shift, nb = pl.nb;
  const int t0 = blockIdx.x * kSelTilesPerBlock;
  const uint16_t* map = p.binmap + (size_t)g * kSelBins;
  for (int i = tid; i < nb; i += kSelThreads) s_map[i] = map[i];
  for (int i = tid; i < (kSelMaxCuts + 1) * 2; i += kSelThreads) (&s_sum[0][0])[i] = 0.0;
  const float* k = p.keys[f] + lo;
  const float* q0 = SELF ? nullptr : p.pay0[f] + lo;
  const float* q1 = NPAY == 2 ? p.pay1[f] + lo : nullptr;
  const long long slot = p.seg_offsets[g];
  float* ck = p.ckeys + slot;
  float* c0 = SELF ? nullptr : p.cpay0[f] + lo;
  float* c1 = NPAY == 2 ? p.cpay1[f] + lo : nullptr;
  const uint8_t* modes = p.tilemode + (size_t)g * p.max_tiles * kSelMaxCuts;
  const uint32_t* tb = p.tilebase + (size_t)g * p.max_tiles;

  for (int tt = 0; tt < kSelTilesPerBlock; ++tt) {
    const int t = t0 + tt;
    const long long tile_lo = (long long)t * kSelTile;
    if (tile_lo >= len) break;
    const int count = (int)min((long long)kSelTile, len - tile_lo);
    __syncthreads();  // previous tile done with s_mode / s_wtot; first tile: tables loaded
    if (tid < ncells) s_mode[tid] = modes[(size_t)t * kSelMaxCuts + tid];
    // element order inside the tile: (warp, item, lane)
    float kf[kSelItems], a0[kSelItems], a1[kSelItems];
#pragma unroll
    for (int i = 0; i < kSelItems; ++i) {
      const int pos = min(warp * (32 * kSelItems) + i * 32 + lane, count - 1);
      kf[i] = __ldcs(k + tile_lo + pos);
    }
    if (!SELF) {
#pragma unroll
      for (int i = 0; i < kSelItems; ++i) {
        const int pos = min(warp * (32 * kSelItems) + i * 32 + lane, count - 1);
        a0[i] = __ldcs(q0 + tile_lo + pos);
      }
    }
    if (NPAY == 2) {
#pragma unroll
      for (int i = 0; i < kSelItems; ++i) {
        const int pos = min(warp * (32 * kSelItems) + i * 32 + lane, count - 1);
        a1[i] = __ldcs(q1 + tile_lo + pos);
      }
    }
    __syncthreads();  // s_mode visible
    const int hot_cls = hot >= 0 ? (int)s_mode[hot] : (int)kSelCompact;
    double hot0 = 0.0, hot1 = 0.0;
    uint32_t rk[kSelItems];
    uint32_t cmask = 0u, running = 0u;
#pragma unroll
    for (int i = 0; i < kSelItems; ++i) {
      const int pos = warp * (32 * kSelItems) + i * 32 + lane;
      const bool valid = pos < count;
      bool compact = false;
      if (valid) {
        const uint32_t key = sort_key_from_float(kf[i]);
        const uint16_t m = s_map[(key - umin) >> shift];
        int cls;
        bool is_hot = false;
        if (m & kSelCellFlag) {
          const int cell = m & 0x7FFF;
          cls = s_mode[cell];
          compact = cls == (int)kSelCompact;
          is_hot = cell == hot;
        } else {
          cls = m;
        }
        if (!compact) {
          const double v0 = (double)(SELF ? kf[i] : a0[i]);
          if (is_hot) {
            hot0 += v0;
            if (NPAY == 2) hot1 += (double)a1[i];
          } else {
            atomicAdd(&s_sum[cls][0], v0);
            if (NPAY == 2) atomicAdd(&s_sum[cls][1], (double)a1[i]);
          }
        }
      }
      const unsigned bal = __ballot_sync(FULL_MASK, compact);
      rk[i] = running + __popc(bal & lt_mask);
      running += __popc(bal);
      if (compact) cmask |= 1u << i;
    }
    if (hot >= 0 && hot_cls != (int)kSelCompact) {  // uniform over the block
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        hot0 += shfl_xor_double(FULL_MASK, hot0, o);
        if (NPAY == 2) hot1 += shfl_xor_double(FULL_MASK, hot1, o);
      }
      if (lane == 0) {
        atomicAdd(&s_sum[hot_cls][0], hot0);
        if (NPAY == 2) atomicAdd(&s_sum[hot_cls][1], hot1);
      }
    }
    if (lane == 0) s_wtot[warp] = running;
    __syncthreads();
    uint32_t wbase = tb[t];
#pragma unroll
    for (int w = 0; w < kSelWarps; ++w)
      if (w < warp) wbase += s_wtot[w];
    if (cmask) {
#pragma unroll
      for (int i = 0; i < kSelItems; ++i) {
        if (cmask & (1u << i)) {
          const uint32_t d = wbase + rk[i];
          ck[d] = kf[i];
          if (!SELF) c0[d] = a0[i];
          if (NPAY == 2) c1[d] = a1[i];
        }
      }
    }
  }
  __syncthreads();
  double* sp = p.spart + ((size_t)g * p.max_blocks + blockIdx.x) * (size_t)(nc + 1) * 2;
  for (int i = tid; i < (nc + 1) * 2; i += kSelThreads) sp[i] = (&s_sum[0][0])[i];
}

__global__ void __launch_bounds__(kSelThreads) sel_classify(const SelParams p) {
  const int g = blockIdx.y;
  int f, b;
  long long lo, len;
  sel_segment(p, g, f, b, lo, len);
  if ((long long)blockIdx.x * kSelTilesPerBlock * kSelTile >= len) return;
  if (p.self_payload[f]) sel_classify_body<1, true>(p, g, f, lo, len);
  else if (p.pay1[f]) sel_classify_body<2, false>(p, g, f, lo, len);
  else sel_classify_body<1, false>(p, g, f, lo, len);
}

// ---- class sums -> per-cut sums of the decided elements, one block per segment ---------------------------
__global__ void __launch_bounds__(kSelThreads) sel_coarse_finish(const SelParams p) {
  __shared__ double s_tot[(kSelMaxCuts + 1) * 2];
  const int g = blockIdx.x, tid = threadIdx.x, nc = p.num_cuts;
  int f, b;
  long long lo, len;
  sel_segment(p, g, f, b, lo, len);
  const int ntiles = (int)((len + kSelTile - 1) / kSelTile);
  const int nblk = (ntiles + kSelTilesPerBlock - 1) / kSelTilesPerBlock;
  const size_t stride = (size_t)(nc + 1) * 2;
  const double* sp = p.spart + (size_t)g * p.max_blocks * stride;
  for (int i = tid; i < (int)stride; i += kSelThreads) {
    double s = 0.0;
    for (int blk = 0; blk < nblk; ++blk) s += sp[(size_t)blk * stride + i];
    s_tot[i] = s;
  }
  __syncthreads();
  const SegPlan& pl = p.plan[g];
  if (tid < p.npay[f]) {
    double run = 0.0;
    double* out = p.coarse + ((size_t)b * p.num_values + p.row0[f] + tid) * nc;
    for (int j = 0; j < nc; ++j) {
      run += s_tot[j * 2 + tid];
      out[pl.cut_k[j]] = run;
    }
  }
}

struct SelLayout {
  size_t off_zero, zero_bytes, off_range, off_hist, off_cellmm, off_ncomp;
  size_t off_segoff, off_plan, off_binmap, off_tilecounts, off_tilemode, off_tilebase, off_complen, off_lens,
      off_pos, off_spart, off_coarse, off_ckeys, off_csorted, off_cperm, off_cpay, off_sortws, off_cutws, total;
  size_t sortws_bytes, cutws_bytes;
  int max_tiles, max_blocks, G;
};

static SelLayout sel_layout(int F, int B, int V, int num_cuts, int num_side_arrays, long long total,
                            long long max_len) {
  SelLayout l{};
  l.G = F * B;
  l.max_tiles = (int)((max_len + kSelTile - 1) / kSelTile);
  if (l.max_tiles < 1) l.max_tiles = 1;
  l.max_blocks = (l.max_tiles + kSelTilesPerBlock - 1) / kSelTilesPerBlock;
  const size_t G = (size_t)l.G;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    const size_t at = o;
    o = align_up(o + bytes, 256);
    return at;
  };
  l.off_zero = o;
  l.off_range = take(G * 2 * sizeof(uint32_t));
  l.off_hist = take(G * kSelBins * sizeof(uint32_t));
  l.off_cellmm = take(G * kSelMaxCuts * 2 * sizeof(uint32_t));
  l.off_ncomp = take(G * l.max_tiles * sizeof(uint32_t));
  l.zero_bytes = o - l.off_zero;
  l.off_segoff = take((G + 1) * sizeof(long long));
  l.off_plan = take(G * sizeof(SegPlan));
  l.off_binmap = take(G * kSelBins * sizeof(uint16_t));
  l.off_tilecounts = take(G * (size_t)num_cuts * l.max_tiles * sizeof(uint32_t));
  l.off_tilemode = take(G * (size_t)l.max_tiles * kSelMaxCuts);
  l.off_tilebase = take(G * (size_t)l.max_tiles * sizeof(uint32_t));
  l.off_complen = take(G * sizeof(long long));
  l.off_lens = take((size_t)V * B * sizeof(long long));
  l.off_pos = take((size_t)V * B * num_cuts * sizeof(long long));
  l.off_spart = take(G * (size_t)l.max_blocks * (num_cuts + 1) * 2 * sizeof(double));
  l.off_coarse = take((size_t)B * V * num_cuts * sizeof(double));
  l.off_ckeys = take((size_t)F * total * sizeof(float));
  l.off_csorted = take((size_t)F * total * sizeof(float));
  l.off_cperm = take((size_t)F * total * sizeof(int32_t));
  l.off_cpay = take((size_t)num_side_arrays * total * sizeof(float));
  l.sortws_bytes = segmented_sort_workspace(l.G, (long long)F * total, max_len, true);
  l.off_sortws = take(l.sortws_bytes);
  l.cutws_bytes = cut_prefix_workspace(B, max_len, V, num_cuts);
  l.off_cutws = take(l.cutws_bytes);
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
  // worst case: two payload rows and two side arrays per family
  return ub::sel_layout(num_families, num_views, 2 * num_families, num_cuts, 2 * num_families, total,
                        max_segment_len).total;
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
  const SelLayout lay = sel_layout(num_families, num_views, V, num_cuts, side, total, max_segment_len);
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
  p.view_offsets = reinterpret_cast<const long long*>(seg_offsets);
  p.cuts = reinterpret_cast<const long long*>(cuts);
  p.range = reinterpret_cast<uint32_t*>(ws + lay.off_range);
  p.hist = reinterpret_cast<uint32_t*>(ws + lay.off_hist);
  p.cell_mm = reinterpret_cast<uint32_t*>(ws + lay.off_cellmm);
  p.tile_ncomp = reinterpret_cast<uint32_t*>(ws + lay.off_ncomp);
  p.seg_offsets = reinterpret_cast<long long*>(ws + lay.off_segoff);
  p.plan = reinterpret_cast<SegPlan*>(ws + lay.off_plan);
  p.binmap = reinterpret_cast<uint16_t*>(ws + lay.off_binmap);
  p.tilecounts = reinterpret_cast<uint32_t*>(ws + lay.off_tilecounts);
  p.tilemode = reinterpret_cast<uint8_t*>(ws + lay.off_tilemode);
  p.tilebase = reinterpret_cast<uint32_t*>(ws + lay.off_tilebase);
  p.complen = reinterpret_cast<long long*>(ws + lay.off_complen);
  p.lens_v = reinterpret_cast<long long*>(ws + lay.off_lens);
  p.pos_v = reinterpret_cast<long long*>(ws + lay.off_pos);
  p.spart = reinterpret_cast<double*>(ws + lay.off_spart);
  p.coarse = reinterpret_cast<double*>(ws + lay.off_coarse);
  p.ckeys = reinterpret_cast<float*>(ws + lay.off_ckeys);
  float* csorted = reinterpret_cast<float*>(ws + lay.off_csorted);
  int32_t* cperm = reinterpret_cast<int32_t*>(ws + lay.off_cperm);
  float* cpay = reinterpret_cast<float*>(ws + lay.off_cpay);
  const float* values[2 * kSelMaxFamilies];
  const int32_t* perms[2 * kSelMaxFamilies];
  {
    int s = 0;
    for (int f = 0; f < num_families; ++f) {
      if (p.self_payload[f]) {
        values[p.row0[f]] = csorted + (size_t)f * total;
        perms[p.row0[f]] = nullptr;
        continue;
      }
      p.cpay0[f] = cpay + (size_t)(s++) * total;
      values[p.row0[f]] = p.cpay0[f];
      perms[p.row0[f]] = cperm + (size_t)f * total;
      if (p.pay1[f]) {
        p.cpay1[f] = cpay + (size_t)(s++) * total;
        values[p.row0[f] + 1] = p.cpay1[f];
        perms[p.row0[f] + 1] = cperm + (size_t)f * total;
      }
    }
  }
  const int G = lay.G;
  if (cudaMemsetAsync(ws + lay.off_zero, 0, lay.zero_bytes, stream) != cudaSuccess)
    return check_launch("cut_select_sums memset");
  const unsigned chunks = (unsigned)((max_segment_len + kSelChunk - 1) / kSelChunk);
  dim3 grid_chunks(chunks < 1 ? 1 : chunks, (unsigned)G);
  dim3 grid_blocks((unsigned)lay.max_blocks, (unsigned)G);
  sel_range<<<grid_chunks, kSelThreads, 0, stream>>>(p);
  sel_hist<<<grid_chunks, kSelThreads, 0, stream>>>(p);
  sel_locate<<<G, kSelThreads, 0, stream>>>(p);
  sel_cell_counts<<<grid_blocks, kSelThreads, 0, stream>>>(p);
  sel_plan_cells<<<dim3((unsigned)num_cuts, (unsigned)G), kSelThreads, 0, stream>>>(p);
  sel_plan_tiles<<<G, kSelThreads, 0, stream>>>(p);
  sel_classify<<<grid_blocks, kSelThreads, 0, stream>>>(p);
  sel_coarse_finish<<<G, kSelThreads, 0, stream>>>(p);
  int rc = check_launch("cut_select_sums");
  if (rc != UB_OK) return rc;
  if (total > 0 && max_segment_len > 0) {
    rc = segmented_sort_impl(p.ckeys, G, reinterpret_cast<const int64_t*>(p.seg_offsets),
                             reinterpret_cast<const int64_t*>(p.complen), (int64_t)num_families * total,
                             max_segment_len, csorted, cperm, ws + lay.off_sortws, lay.sortws_bytes, stream_v);
    if (rc != UB_OK) return rc;
  }
  return cut_prefix_impl(values, perms, V, num_views, seg_offsets, reinterpret_cast<const int64_t*>(p.lens_v),
                         num_views, max_segment_len, reinterpret_cast<const int64_t*>(p.pos_v),
                         (int64_t)num_views * num_cuts, num_cuts, p.coarse, out_sums, ws + lay.off_cutws,
                         lay.cutws_bytes, stream_v);
}

}  // extern "C"
