// (C1) Metric prologue + NLL + AUCE interval histogram, one pass over (pred, target, std).
//
// Reference arithmetic (file:line under /root/reference/nerfuncertainty):
//   scripts/eval_uncertainty.py:323-333   se = sum_c (p-t)^2, ae = sum_c |p-t|, var = std^2
//   scripts/eval_uncertainty.py:404-412   -Normal(p, max(std, eps)).log_prob(t)
//   scripts/eval_uncertainty.py:371-378   sigma = sqrt(var) repeated over channels -> auce()
//   metrics/auce.py:18-28                 99 x count(t >= m - z s  and  t <= m + z s)
//
// The 99 coverage passes collapse into one histogram: the interval predicate is monotone in z
// (z decreasing in k), so every element is characterised by the number c of leading thresholds
// it satisfies; coverage_k = #{c > k}.  The exact predicate is float64 (NumPy >= 2 promotion:
// np.float64 z * float32 sigma -> float64, no fused multiply-add).  Evaluating it 99 (or even 7) times
// per element makes the pass instruction-bound, so c is located in three tiers:
//   1. r = |t - m| / sigma in float32 (relative error < 1e-6) indexes a shared-memory table over
//      uniform bins of r, built per block from the z table.  A bin that holds no threshold within a
//      relative guard band of 1e-5 determines c outright: the float64 outcome cannot differ, because
//      the estimate error (4e-7 relative) plus the float64 rounding of m -+ z sigma (2^-52 (|m|/sigma
//      + z) absolute, < 3e-10 once sigma >= 1e-6 |m| is required) is far inside the band (which
//      also has an absolute part of 1e-9).                                  (~99 % of elements)
//   2. A bin with exactly one threshold k in the band: c = k + [exact float64 predicate for z_k].
//   3. Anything else (NaN / negative ratio, tiny sigma, several thresholds per bin): the float32
//      binary search + float64 verification + exact float64 binary search of coverage_count().
// The counts are therefore bit-exact whatever the estimate does.  Tier 2 is evaluated inline for the whole warp
// whenever any lane needs it (a call per lane ran with one thread active and cost a quarter of the kernel's
// instructions: profiles/r1_select_ncu_summary.txt -> r2).
// Sums are accumulated in float64 per block and combined in a fixed order (deterministic).
// NLL: d^2 / (2 s^2) and log s use the fast division / logarithm (relative error ~1e-7 per element, far inside the
// 1e-5 contract of the mean); everything that feeds an integer result (the coverage counts) stays exact.
// Optionally the kernel also counts, per segment, the top 12 bits of the order-preserving keys of the three vectors
// it writes (var, abs err, sq err): the coarse histograms ub_cut_select_sums would otherwise rebuild with a pass of
// its own over the vectors it has just been handed (-12 B/pixel, -1 launch).
#include "ub_common.cuh"

namespace ub {

constexpr int kPrologueThreads = 256;
constexpr int kPixPerThread = 4;
constexpr int kPixPerChunk = kPrologueThreads * kPixPerThread;
// chunks per block: amortise the table build and the flush of the block's histograms (4096 pixels per block; 8192 once
// the batch is large enough to fill the device several times over anyway)
__host__ __device__ inline int prologue_chunks_per_block(long long num_segments, long long max_len) {
  return num_segments * max_len >= (4LL << 20) ? 8 : 4;
}
constexpr int kMaxZ = 127;
constexpr int kLutBins = 16384;           // uniform bins of r over [0, 1.001 z_0)
constexpr float kLutGuard = 1e-5f;        // relative guard band around every bin (float32 estimate error)
constexpr double kLutGuardAbs = 1e-9;     // absolute guard band (float64 rounding of m -+ z sigma)
constexpr float kSigmaGuard = 1e6f;       // tiers 1/2 need sigma >= 1e-6 |m|: 2^-52 (|m|/sigma + z) << 1e-9
constexpr unsigned kLutOneThreshold = 128u;  // entry = 128 + k: only threshold k is undecided
constexpr unsigned kLutGeneric = 255u;

struct PrologueParams {
  const float* pred;
  const float* target;
  const float* std;
  int channels;
  int num_segments;
  const long long* seg_offsets;  // device copy
  float nll_min_std;
  int sigma_from_var;
  const double* z;
  int num_z;
  float* o_se;
  float* o_ae;
  float* o_var;
  double* partial;  // [num_segments][blocks_per_seg][NSUMS]
  unsigned* done;   // [num_segments] blocks of the segment that have written their partial sums (zeroed)
  double* out_sums; // [num_segments][NSUMS]
  int blocks_per_seg;
  int chunks_per_block;
  unsigned long long* hist;  // [num_segments][num_z + 1]
  const unsigned char* lut;  // [kLutBins + 16] ratio table (workspace)
  int vec_ok;
  unsigned* coarse;          // [3][num_segments][kCoarseBins] or NULL: top-12-bit key histograms of var / ae / se
};

constexpr int kCoarseBins = 4096;
constexpr int kCoarseShift = 20;

// sort_key_from_float(x) >> 20 for a value that is never negative (a square, a sum of magnitudes): the sign bit
// of the key is set, NaN is the one largest key
__device__ __forceinline__ unsigned coarse_bin_nonneg(float x) {
  return x != x ? (unsigned)(kCoarseBins - 1) : (0x800u | (__float_as_uint(x) >> kCoarseShift));
}

__device__ __forceinline__ bool interval_holds(double z, float m, float s, float t) {
  const double zs = __dmul_rn(z, (double)s);
  const double lo = __dsub_rn((double)m, zs);
  const double hi = __dadd_rn((double)m, zs);
  const double td = (double)t;
  return td >= lo && td <= hi;
}

// number of leading thresholds (z strictly decreasing) whose interval contains t
__device__ __noinline__ int coverage_count(const double* __restrict__ zs_d,
                                              const float* __restrict__ zs_f, int nz, float m,
                                              float s, float t) {
  // float32 estimate: count of z_k >= |t - m| / s
  const float r = fabsf(t - m) / s;
  int lo = 0, hi = nz;  // invariant: z[lo-1] >= r (or lo == 0), z[hi] < r (or hi == nz)
  if (r == r) {
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (zs_f[mid] >= r) lo = mid + 1; else hi = mid;
    }
  } else {
    lo = 0;  // NaN ratio (0/0, NaN inputs): let the exact check decide
  }
  int c = lo;
  const bool left_ok = c == 0 || interval_holds(zs_d[c - 1], m, s, t);
  const bool right_ok = c == nz || !interval_holds(zs_d[c], m, s, t);
  if (left_ok && right_ok) return c;
  // exact float64 binary search (predicate true on a prefix of k)
  lo = 0;
  hi = nz;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (interval_holds(zs_d[mid], m, s, t)) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Ratio table, built once per call into the workspace: one bin per thread.  z strictly decreasing.
// Entry for the bin [i w, (i+1) w) widened by the guard bands: see the tiers above.
__global__ void __launch_bounds__(256) prologue_lut_kernel(const double* __restrict__ z, int nz,
                                                           unsigned char* __restrict__ lut) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > kLutBins) return;
  if (i == kLutBins) {  // r >= r_max > z_0 (1 + guard): no threshold holds
    lut[i] = 0;
    return;
  }
  const double r_max = fmax(z[0] * 1.001, 1e-30);
  const double bin_w = r_max / kLutBins;
  const double a = i * bin_w * (1.0 - (double)kLutGuard) - kLutGuardAbs;
  const double b = (i + 1) * bin_w * (1.0 + (double)kLutGuard) + kLutGuardAbs;
  int lo = 0, hi = nz;  // above = #{k : z_k > b}
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (z[mid] > b) lo = mid + 1; else hi = mid;
  }
  const int above = lo;
  int inside = 0;
  while (above + inside < nz && z[above + inside] >= a && inside < 2) ++inside;
  lut[i] = inside == 0 ? (unsigned char)above
                       : (inside == 1 ? (unsigned char)(kLutOneThreshold + above) : (unsigned char)kLutGeneric);
}

template <int C, bool COARSE>
__global__ void __launch_bounds__(kPrologueThreads, 3) score_prologue_kernel(const PrologueParams p) {
  __shared__ double z_d[kMaxZ + 1];
  __shared__ float z_f[kMaxZ + 1];
  __shared__ unsigned int hist_s[kPrologueThreads / 32][kMaxZ + 1];
  __shared__ double red[kPrologueThreads / 32][UB_PROLOGUE_NSUMS];
  __shared__ __align__(16) unsigned char lut[kLutBins + 16];
  // coarse key histograms of the block, two 16-bit counters per word (a block holds <= 8192 pixels < 2^16)
  __shared__ unsigned int coarse_s[COARSE ? 3 : 1][COARSE ? kCoarseBins / 2 : 1];

  const int seg = blockIdx.y;
  const long long seg_lo = p.seg_offsets[seg], seg_hi = p.seg_offsets[seg + 1];
  const long long blk_lo = seg_lo + (long long)blockIdx.x * (kPixPerChunk * p.chunks_per_block);
  if (blk_lo >= seg_hi && blockIdx.x > 0) return;  // ragged batches: nothing in this block (block 0 writes the zero partials)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nz = p.num_z;

  for (int i = threadIdx.x; i <= kMaxZ; i += blockDim.x) {
    z_d[i] = i < nz ? p.z[i] : 0.0;
    z_f[i] = i < nz ? (float)p.z[i] : 0.f;
  }
  for (int i = threadIdx.x; i < (kPrologueThreads / 32) * (kMaxZ + 1); i += blockDim.x)
    (&hist_s[0][0])[i] = 0u;
  if (COARSE)
    for (int i = threadIdx.x; i < 3 * (kCoarseBins / 2); i += blockDim.x) (&coarse_s[0][0])[i] = 0u;
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.lut);
    uint4* dst = reinterpret_cast<uint4*>(lut);
    for (int i = threadIdx.x; i < (kLutBins + 16) / 16; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const float inv_w = (float)((double)kLutBins / fmax(z_d[0] * 1.001, 1e-30));

  double sums[UB_PROLOGUE_NSUMS] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
  for (int chunk = 0; chunk < p.chunks_per_block; ++chunk) {
    const long long pix0 = blk_lo + (long long)chunk * kPixPerChunk + (long long)threadIdx.x * kPixPerThread;
    const int npix = (int)max(0LL, min((long long)kPixPerThread, seg_hi - pix0));
    if (__syncthreads_count(npix > 0) == 0) break;  // the whole block is past the segment's end
    float pr[kPixPerThread * C], tg[kPixPerThread * C], sd[kPixPerThread];
    const bool vec = p.vec_ok && npix == kPixPerThread && (pix0 & 3) == 0;  // 16-byte aligned rows
    if (vec) {
      const float4* a = reinterpret_cast<const float4*>(p.pred + pix0 * C);
      const float4* b = reinterpret_cast<const float4*>(p.target + pix0 * C);
#pragma unroll
      for (int j = 0; j < C; ++j) {
        const float4 x = __ldcs(a + j), y = __ldcs(b + j);
        pr[4 * j] = x.x; pr[4 * j + 1] = x.y; pr[4 * j + 2] = x.z; pr[4 * j + 3] = x.w;
        tg[4 * j] = y.x; tg[4 * j + 1] = y.y; tg[4 * j + 2] = y.z; tg[4 * j + 3] = y.w;
      }
      const float4 s4 = __ldcs(reinterpret_cast<const float4*>(p.std + pix0));
      sd[0] = s4.x; sd[1] = s4.y; sd[2] = s4.z; sd[3] = s4.w;
    } else {
#pragma unroll
      for (int i = 0; i < kPixPerThread * C; ++i) {
        const bool ok = i < npix * C;
        pr[i] = ok ? p.pred[pix0 * C + i] : 0.f;
        tg[i] = ok ? p.target[pix0 * C + i] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < kPixPerThread; ++i) sd[i] = i < npix ? p.std[pix0 + i] : 1.f;
    }
    // every lane runs every pixel slot (dummy data past the end), so the warp votes below are convergent
    float se[kPixPerThread], ae[kPixPerThread], vr[kPixPerThread];
#pragma unroll
    for (int px = 0; px < kPixPerThread; ++px) {
      const bool valid = px < npix;
      const float sdv = sd[px];
      const float var = __fmul_rn(sdv, sdv);
      const float sigma = p.sigma_from_var ? sqrtf(var) : sdv;
      const float nll_s = fmaxf(sdv, p.nll_min_std);  // torch.maximum propagates NaN:
      const float s_nll = sdv != sdv ? sdv : nll_s;
      const float inv_two_var = __fdividef(0.5f, __fmul_rn(s_nll, s_nll));
      const float log_s = __logf(s_nll) + 0.91893853320467274178f;
      const float inv_sigma = __fdividef(1.0f, sigma);
      // tiers 1 / 2 need a positive, not-tiny sigma with sigma >= 1e-6 |m|; tested once per pixel against the largest
      // |m| of its channels (conservative: a channel sent to tier 3 needlessly is still counted exactly) and folded into
      // the ratio's multiplier, so that a failed guard shows up as a NaN ratio
      float m_max = fabsf(pr[px * C]);
#pragma unroll
      for (int c = 1; c < C; ++c) m_max = fmaxf(m_max, fabsf(pr[px * C + c]));
      const bool guard_ok = sigma > 1e-30f && sigma < 1e30f && sigma * kSigmaGuard >= m_max;
      // sigma == 0 exactly (an empty ray: all weights zero) is common enough to keep out of tier 3: every interval is
      // the point m, so the count is 0 unless t == m -- an infinite multiplier makes the ratio inf (table entry "no
      // threshold holds") for d != 0 and NaN (-> the exact path) for d == 0
      const float r_mul = guard_ok ? inv_sigma : (sigma == 0.f ? __int_as_float(0x7F800000) : __int_as_float(0x7FC00000));
      float se_px = 0.f, ae_px = 0.f, nll_px = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float m = pr[px * C + c], t = tg[px * C + c];
        const float d = m - t;
        const float d2 = __fmul_rn(d, d);
        se_px = c == 0 ? d2 : __fadd_rn(se_px, d2);
        ae_px = c == 0 ? fabsf(d) : __fadd_rn(ae_px, fabsf(d));
        nll_px += fmaf(d2, inv_two_var, log_s);
        // coverage count, tiers 1-3
        const float r = fabsf(d) * r_mul;
        const unsigned bin = min((unsigned)__float2int_rz(r * inv_w), (unsigned)kLutBins);
        const unsigned looked_up = lut[bin];
        const unsigned e = r >= 0.f ? looked_up : kLutGeneric;  // false for the NaN of a failed guard / NaN inputs
        int cnt = (int)e;
        if (__any_sync(FULL_MASK, valid && e >= kLutOneThreshold)) {  // warp-uniform
          // tier 2 for the whole warp, branch-free: one exact float64 predicate at threshold k
          const int k = (int)(e & 127u);
          const bool holds = interval_holds(z_d[k], m, sigma, t);
          if (e != kLutGeneric && e >= kLutOneThreshold) cnt = k + (holds ? 1 : 0);
          if (__any_sync(FULL_MASK, valid && e == kLutGeneric)) {  // tier 3: NaN / negative ratios, tiny sigma
            if (valid && e == kLutGeneric) cnt = coverage_count(z_d, z_f, nz, m, sigma, t);
          }
        }
        if (valid) atomicAdd(&hist_s[warp][cnt], 1u);
      }
      se[px] = se_px;
      ae[px] = ae_px;
      vr[px] = var;
      if (valid) {
        sums[0] += (double)se_px;
        sums[1] += (double)ae_px;
        sums[2] += (double)var;
        sums[3] += (double)nll_px;
        sums[4] += (double)sigma;
        if (COARSE) {
          const unsigned kv = coarse_bin_nonneg(var);
          const unsigned ka = coarse_bin_nonneg(ae_px);
          const unsigned ks = coarse_bin_nonneg(se_px);
          atomicAdd(&coarse_s[0][kv >> 1], 1u << (16 * (kv & 1u)));
          atomicAdd(&coarse_s[1][ka >> 1], 1u << (16 * (ka & 1u)));
          atomicAdd(&coarse_s[2][ks >> 1], 1u << (16 * (ks & 1u)));
        }
      }
    }
    if (vec) {
      if (p.o_se) *reinterpret_cast<float4*>(p.o_se + pix0) = make_float4(se[0], se[1], se[2], se[3]);
      if (p.o_ae) *reinterpret_cast<float4*>(p.o_ae + pix0) = make_float4(ae[0], ae[1], ae[2], ae[3]);
      if (p.o_var) *reinterpret_cast<float4*>(p.o_var + pix0) = make_float4(vr[0], vr[1], vr[2], vr[3]);
    } else {
#pragma unroll
      for (int px = 0; px < kPixPerThread; ++px)
        if (px < npix) {
          if (p.o_se) p.o_se[pix0 + px] = se[px];
          if (p.o_ae) p.o_ae[pix0 + px] = ae[px];
          if (p.o_var) p.o_var[pix0 + px] = vr[px];
        }
    }
  }

  // block reduction of the float64 sums in a fixed order
#pragma unroll
  for (int j = 0; j < UB_PROLOGUE_NSUMS; ++j) {
    double v = sums[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_xor_double(FULL_MASK, v, o);
    if (lane == 0) red[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < UB_PROLOGUE_NSUMS) {
    double v = 0.0;
    for (int w = 0; w < kPrologueThreads / 32; ++w) v += red[w][threadIdx.x];
    p.partial[((size_t)seg * p.blocks_per_seg + blockIdx.x) * UB_PROLOGUE_NSUMS + threadIdx.x] = v;
    __threadfence();
  }
  for (int c = threadIdx.x; c <= nz; c += blockDim.x) {
    unsigned int v = 0;
    for (int w = 0; w < kPrologueThreads / 32; ++w) v += hist_s[w][c];
    if (v) atomicAdd(&p.hist[(size_t)seg * (nz + 1) + c], (unsigned long long)v);
  }
  if (COARSE) {
    for (int i = threadIdx.x; i < 3 * (kCoarseBins / 2); i += blockDim.x) {
      const unsigned v = (&coarse_s[0][0])[i];
      if (v) {
        const int f = i / (kCoarseBins / 2), w2 = i - f * (kCoarseBins / 2);
        unsigned* g = p.coarse + ((size_t)f * p.num_segments + seg) * kCoarseBins + 2 * w2;
        if (v & 0xFFFFu) atomicAdd(g, v & 0xFFFFu);
        if (v >> 16) atomicAdd(g + 1, v >> 16);
      }
    }
  }
  // the last block of the segment to get here adds the partial sums of all of them, in block order (deterministic):
  // one warp per sum, lanes stride over the blocks, then a fixed-order shuffle tree
  __shared__ unsigned s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long pix_per_block = (long long)kPixPerChunk * p.chunks_per_block;
    const unsigned nblk = (unsigned)max(1LL, (seg_hi - seg_lo + pix_per_block - 1) / pix_per_block);
    s_last = atomicAdd(p.done + seg, 1u) == nblk - 1u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (warp < UB_PROLOGUE_NSUMS) {
    double v = 0.0;
    for (int b = lane; b < p.blocks_per_seg; b += 32)
      v += __ldcg(p.partial + ((size_t)seg * p.blocks_per_seg + b) * UB_PROLOGUE_NSUMS + warp);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_xor_double(FULL_MASK, v, o);
    if (lane == 0) p.out_sums[(size_t)seg * UB_PROLOGUE_NSUMS + warp] = v;
  }
}

struct PrologueLayout {
  size_t off_offsets, off_partial, partial_bytes, off_done, off_lut, total;
  int blocks_per_seg, chunks_per_block;
};
static PrologueLayout prologue_layout(int num_segments, long long max_len) {
  PrologueLayout l{};
  l.chunks_per_block = prologue_chunks_per_block(num_segments, max_len);
  const long long pix_per_block = (long long)kPixPerChunk * l.chunks_per_block;
  l.blocks_per_seg = (int)((max_len + pix_per_block - 1) / pix_per_block);
  if (l.blocks_per_seg < 1) l.blocks_per_seg = 1;
  l.off_offsets = 0;
  l.off_partial = align_up((size_t)(num_segments + 1) * sizeof(long long), 256);
  l.partial_bytes = (size_t)num_segments * l.blocks_per_seg * UB_PROLOGUE_NSUMS * sizeof(double);
  l.off_done = l.off_partial + l.partial_bytes;  // zeroed together with the partial sums
  l.off_lut = align_up(l.off_done + (size_t)num_segments * sizeof(unsigned), 256);
  l.total = l.off_lut + kLutBins + 16;
  return l;
}

}  // namespace ub

extern "C" {

size_t ub_score_prologue_workspace_bytes(int32_t num_segments, int64_t max_segment_len, int32_t num_z) {
  (void)num_z;
  if (num_segments < 1 || max_segment_len < 0) return 256;
  return ub::prologue_layout(num_segments, max_segment_len).total;
}

int ub_score_prologue(const ub_score_prologue_args* a, void* workspace, size_t workspace_bytes,
                      void* stream_v) {
  using namespace ub;
  UB_REQUIRE(a != nullptr, UB_ERR_BAD_ARG, "score_prologue: args is NULL");
  UB_REQUIRE(a->channels == 1 || a->channels == 3, UB_ERR_UNSUPPORTED,
             "score_prologue: channels must be 1 or 3 (got %d)", a->channels);
  UB_REQUIRE(a->num_segments >= 1 && a->num_segments <= 65535 && a->seg_offsets != nullptr,
             UB_ERR_BAD_ARG, "score_prologue: bad segments");
  UB_REQUIRE(a->max_segment_len >= 0, UB_ERR_BAD_ARG, "score_prologue: bad max_segment_len");
  UB_REQUIRE(a->num_z >= 1 && a->num_z <= kMaxZ && a->z_values != nullptr, UB_ERR_BAD_ARG,
             "score_prologue: num_z must be in [1, %d]", kMaxZ);
  UB_REQUIRE(a->out_sums != nullptr && a->out_hist != nullptr, UB_ERR_BAD_ARG,
             "score_prologue: out_sums / out_hist must be non-NULL");
  UB_REQUIRE(a->max_segment_len == 0 || (a->pred && a->target && a->std), UB_ERR_BAD_ARG,
             "score_prologue: pred/target/std must be non-NULL");
  const PrologueLayout lay = prologue_layout(a->num_segments, a->max_segment_len);
  UB_REQUIRE(workspace != nullptr && workspace_bytes >= lay.total, UB_ERR_WORKSPACE,
             "score_prologue: workspace %zu B < required %zu B", workspace_bytes, lay.total);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  char* ws = static_cast<char*>(workspace);
  if (cudaMemsetAsync(a->out_hist, 0, (size_t)a->num_segments * (a->num_z + 1) * sizeof(int64_t),
                      stream) != cudaSuccess)
    return check_launch("score_prologue hist memset");
  if (a->out_coarse_hist) {
    UB_REQUIRE(a->out_sq_err && a->out_abs_err && a->out_var, UB_ERR_BAD_ARG,
               "score_prologue: out_coarse_hist needs the three output vectors");
    if (cudaMemsetAsync(a->out_coarse_hist, 0, (size_t)3 * a->num_segments * kCoarseBins * sizeof(uint32_t), stream) !=
        cudaSuccess)
      return check_launch("score_prologue coarse memset");
  }
  // blocks past the end of a short segment exit early: their partial sums must read as zero
  if (cudaMemsetAsync(ws + lay.off_partial, 0, lay.partial_bytes + (size_t)a->num_segments * sizeof(unsigned), stream) !=
      cudaSuccess)
    return check_launch("score_prologue partial memset");

  PrologueParams p{};
  p.pred = a->pred;
  p.target = a->target;
  p.std = a->std;
  p.channels = a->channels;
  p.num_segments = a->num_segments;
  p.seg_offsets = reinterpret_cast<const long long*>(a->seg_offsets);
  p.nll_min_std = a->nll_min_std;
  p.sigma_from_var = a->sigma_from_var;
  p.z = a->z_values;
  p.num_z = a->num_z;
  p.o_se = a->out_sq_err;
  p.o_ae = a->out_abs_err;
  p.o_var = a->out_var;
  p.partial = reinterpret_cast<double*>(ws + lay.off_partial);
  p.done = reinterpret_cast<unsigned*>(ws + lay.off_done);
  p.out_sums = a->out_sums;
  p.blocks_per_seg = lay.blocks_per_seg;
  p.chunks_per_block = lay.chunks_per_block;
  p.hist = reinterpret_cast<unsigned long long*>(a->out_hist);
  unsigned char* lut = reinterpret_cast<unsigned char*>(ws + lay.off_lut);
  p.lut = lut;
  prologue_lut_kernel<<<(kLutBins + 1 + 255) / 256, 256, 0, stream>>>(a->z_values, a->num_z, lut);
  auto al16 = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  bool vec_ok = al16(a->pred) && al16(a->target) && al16(a->std) && al16(a->out_sq_err) &&
                al16(a->out_abs_err) && al16(a->out_var);
  p.vec_ok = vec_ok ? 1 : 0;  // the kernel additionally requires 4-pixel-aligned positions
  p.coarse = a->out_coarse_hist;

  dim3 grid((unsigned)lay.blocks_per_seg, (unsigned)a->num_segments);
  if (a->channels == 3) {
    if (p.coarse) score_prologue_kernel<3, true><<<grid, kPrologueThreads, 0, stream>>>(p);
    else score_prologue_kernel<3, false><<<grid, kPrologueThreads, 0, stream>>>(p);
  } else {
    if (p.coarse) score_prologue_kernel<1, true><<<grid, kPrologueThreads, 0, stream>>>(p);
    else score_prologue_kernel<1, false><<<grid, kPrologueThreads, 0, stream>>>(p);
  }
  return check_launch("score_prologue");
}

}  // extern "C"
