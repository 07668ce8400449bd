// (B) Fused per-pixel mean / spread across K ensemble members or MC-dropout passes.
//
// Reference arithmetic (file:line under /root/reference/nerfuncertainty):
//   models/mcdropout/mcdropout_models.py:121-126   stack -> mean(0); std(0).mean(-1)[..., None]
//   models/ensemble/ensemble_pipeline.py:159-190   same, plus var(0).mean(-1) for the epistemic term
//
// One streaming pass: every member image is read exactly once through its own pointer (no
// torch.stack copy), each thread owns 4 consecutive pixels of a one-channel key (128-bit loads) or 2 of a three-
// channel key (64-bit loads: half the registers of 4 pixels, twice the resident warps), and the K values of an
// element never leave registers.  All output keys of a view (rgb, depth, accumulation, ...) go
// through ONE launch: blockIdx.y selects the key ("job").  The spread uses shifted sums in float64
// (shift = first member), which is exact for identical members and has no mean^2/var
// cancellation; torch's CPU kernel (Welford with float64 accumulators) agrees to float32 rounding.
#include <algorithm>

#include "ub_common.cuh"

namespace ub {

constexpr int kMaxJobs = UB_MAX_REDUCE_JOBS;
constexpr int kMaxBatchMembers = UB_MAX_REDUCE_BATCH_MEMBERS;

struct ReduceJob {
  long long num_pixels;
  int channels;
  int spread_mode;
  float* out_mean;
  float* out_spread;
  int vec_ok;
};
struct ReduceBatch {
  const float* member[kMaxJobs][kMaxBatchMembers];
  ReduceJob job[kMaxJobs];
  int num_members;
};

template <int C, int P, int KMAX, bool GROUPED>
__device__ __forceinline__ void reduce_pixels(const float* const* member, int K, const ReduceJob& jb,
                                              long long pix0, int npix, bool vec) {
  constexpr int E = P * C;  // elements per thread
  static_assert((C == 1 && P == 4) || (C == 3 && P == 2), "vector widths below");
  const long long e0 = pix0 * C;
  const int ne = npix * C;
  auto load = [&](const float* m, float (&dst)[E]) {
    if (vec) {
      if constexpr (C == 1) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(m + e0));
        dst[0] = v.x;
        dst[1] = v.y;
        dst[2] = v.z;
        dst[3] = v.w;
      } else {  // two pixels = 24 bytes at an 8-byte aligned offset (pix0 even, base 16-byte aligned)
        const float2* src = reinterpret_cast<const float2*>(m + e0);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float2 v = __ldcs(src + j);
          dst[2 * j] = v.x;
          dst[2 * j + 1] = v.y;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < E; ++i) dst[i] = i < ne ? m[e0 + i] : 0.f;
    }
  };
  // GROUPED == false (K <= KMAX): all members are loaded before the first use -- up to KMAX * 24 bytes in flight per
  // thread, which together with 3-4 CTAs / SM is what hides the DRAM latency -- and the two float64 sums
  // of an element live only while that element is reduced.  GROUPED == true (any K): member 0 is the shift, the others
  // arrive in groups of KMAX - 1 whose loads are issued together; the sums persist across the groups.  The float64
  // additions run over the members in ascending order either way, so the results are identical.
  float x0[E];
  double s1[E], s2[E];
  if constexpr (!GROUPED) {
    float x[KMAX][E];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) load(member[k < K ? k : 0], x[k]);
#pragma unroll
    for (int i = 0; i < E; ++i) {
      x0[i] = x[0][i];
      const double b0 = (double)x0[i];
      double a = 0.0, b2 = 0.0;
#pragma unroll
      for (int k = 1; k < KMAX; ++k) {
        if (k < K) {
          const double d = (double)x[k][i] - b0;
          a += d;
          b2 = fma(d, d, b2);
        }
      }
      s1[i] = a;
      s2[i] = b2;
    }
  } else {
    constexpr int G = KMAX - 1;
    load(member[0], x0);
#pragma unroll
    for (int i = 0; i < E; ++i) {
      s1[i] = 0.0;
      s2[i] = 0.0;
    }
    for (int k0 = 1; k0 < K; k0 += G) {
      float x[G][E];
#pragma unroll
      for (int u = 0; u < G; ++u) {
        if (k0 + u < K) {  // warp-uniform
          load(member[k0 + u], x[u]);
        } else {
#pragma unroll
          for (int i = 0; i < E; ++i) x[u][i] = x0[i];  // d = 0: adds nothing
        }
      }
#pragma unroll
      for (int i = 0; i < E; ++i) {
        const double b0 = (double)x0[i];
        double a = s1[i], b2 = s2[i];
#pragma unroll
        for (int u = 0; u < G; ++u) {
          const double d = (double)x[u][i] - b0;
          a += d;
          b2 = fma(d, d, b2);
        }
        s1[i] = a;
        s2[i] = b2;
      }
    }
  }
  // Only the cancellation-prone part stays in float64: the shifted sums and  s2 - s1^2 / K  (two float64 operations
  // per element).  Everything downstream of that difference -- the 1 / (K - 1) scale, the square root, the mean
  // x0 + s1 / K -- is float32 arithmetic on values that are already well conditioned (<= 3 ulp of a float in total,
  // 1e-7 relative), which takes ~2/3 of the float64-pipe work (conversions included) out of the kernel.
  const double inv_k = 1.0 / (double)K;
  const float inv_kf = 1.0f / (float)K;
  const float inv_km1 = 1.0f / (float)(K - 1);     // K == 1 -> inf: 0 * inf = NaN like torch's unbiased std of one sample
  float mean[E];
#pragma unroll
  for (int i = 0; i < E; ++i) mean[i] = fmaf((float)s1[i], inv_kf, x0[i]);
  if (jb.out_mean) {
    if (vec) {
      if constexpr (C == 1) {
        *reinterpret_cast<float4*>(jb.out_mean + e0) = make_float4(mean[0], mean[1], mean[2], mean[3]);
      } else {
        float2* dst = reinterpret_cast<float2*>(jb.out_mean + e0);
#pragma unroll
        for (int j = 0; j < 3; ++j) dst[j] = make_float2(mean[2 * j], mean[2 * j + 1]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < E; ++i)
        if (i < ne) jb.out_mean[e0 + i] = mean[i];
    }
  }
  if (jb.spread_mode != UB_SPREAD_NONE && jb.out_spread) {
    float spread[P];
#pragma unroll
    for (int px = 0; px < P; ++px) {
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int i = px * C + c;
        float var = (float)fma(-s1[i] * inv_k, s1[i], s2[i]) * inv_km1;   // sum (d - mean d)^2 / (K - 1)
        if (var < 0.0f) var = 0.0f;
        const float v = jb.spread_mode == UB_SPREAD_STD ? sqrtf(var) : var;
        acc = c == 0 ? v : acc + v;
      }
      spread[px] = C == 1 ? acc : acc / (float)C;
    }
    if (vec) {
      if constexpr (P == 4)
        *reinterpret_cast<float4*>(jb.out_spread + pix0) = make_float4(spread[0], spread[1], spread[2], spread[3]);
      else
        *reinterpret_cast<float2*>(jb.out_spread + pix0) = make_float2(spread[0], spread[1]);
    } else {
#pragma unroll
      for (int px = 0; px < P; ++px)
        if (px < npix) jb.out_spread[pix0 + px] = spread[px];
    }
  }
}

// mean-only fast path (the bulk of a view's keys, e.g. the [H,W,48] density tensor): a flat stream,
// 8 consecutive floats per thread, the loads of up to 4 members issued before any arithmetic, float32
// shifted sums (x0 + sum(x_k - x0) / K: exact for identical members, <= 2 ulp otherwise).
__device__ __forceinline__ void reduce_flat_mean8(const float* const* member, int K, float* out, long long e0) {
  float4 a0 = *reinterpret_cast<const float4*>(member[0] + e0);
  float4 a1 = *reinterpret_cast<const float4*>(member[0] + e0 + 4);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int k = 1; k < K; k += 4) {
    float4 v[4][2];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (k + u < K) {
        v[u][0] = *reinterpret_cast<const float4*>(member[k + u] + e0);
        v[u][1] = *reinterpret_cast<const float4*>(member[k + u] + e0 + 4);
      } else {
        v[u][0] = a0;
        v[u][1] = a1;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s[0] += v[u][0].x - a0.x;
      s[1] += v[u][0].y - a0.y;
      s[2] += v[u][0].z - a0.z;
      s[3] += v[u][0].w - a0.w;
      s[4] += v[u][1].x - a1.x;
      s[5] += v[u][1].y - a1.y;
      s[6] += v[u][1].z - a1.z;
      s[7] += v[u][1].w - a1.w;
    }
  }
  const float inv_k = 1.0f / (float)K;
  *reinterpret_cast<float4*>(out + e0) =
      make_float4(a0.x + s[0] * inv_k, a0.y + s[1] * inv_k, a0.z + s[2] * inv_k, a0.w + s[3] * inv_k);
  *reinterpret_cast<float4*>(out + e0 + 4) =
      make_float4(a1.x + s[4] * inv_k, a1.y + s[5] * inv_k, a1.z + s[6] * inv_k, a1.w + s[7] * inv_k);
}

// any channel count: one thread per pixel, scalar loads
__device__ __forceinline__ void reduce_pixel_any_c(const float* const* member, int K, const ReduceJob& jb,
                                                   long long px) {
  const int C = jb.channels;
  const double inv_k = 1.0 / (double)K;
  float acc = 0.f;
  for (int c = 0; c < C; ++c) {
    const long long e = px * C + c;
    const float x0 = member[0][e];
    double s1 = 0.0, s2 = 0.0;
    for (int k = 1; k < K; ++k) {
      const double d = (double)member[k][e] - (double)x0;
      s1 += d;
      s2 += d * d;
    }
    if (jb.out_mean) jb.out_mean[e] = (float)((double)x0 + s1 * inv_k);
    double var = (s2 - s1 * s1 * inv_k) / (double)(K - 1);
    if (var < 0.0) var = 0.0;
    const float v = jb.spread_mode == UB_SPREAD_STD ? (float)sqrt(var) : (float)var;
    acc = c == 0 ? v : acc + v;
  }
  if (jb.spread_mode != UB_SPREAD_NONE && jb.out_spread) jb.out_spread[px] = C == 1 ? acc : acc / (float)C;
}

// Two instantiations share the launch: FLAT = true carries only the mean-only float32 stream path (about 40
// registers, full occupancy: most of a view's bytes go through it), FLAT = false the float64 spread paths (<= 80
// registers, 3 CTAs per SM).  One kernel for both left the streaming jobs at a quarter of the occupancy they need.
template <bool FLAT>
__global__ void __launch_bounds__(256, FLAT ? 6 : 3) reduce_members_batched_kernel(const __grid_constant__ ReduceBatch b) {
  const ReduceJob& jb = b.job[blockIdx.y];
  const float* const* member = b.member[blockIdx.y];
  const int K = b.num_members;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool is_flat = jb.spread_mode == UB_SPREAD_NONE && jb.vec_ok && jb.out_mean;
  if (is_flat != FLAT) return;
  if (FLAT) {
    const long long n = jb.num_pixels * jb.channels;
    const long long groups = n / 8;
    for (long long g = tid; g < groups; g += stride) reduce_flat_mean8(member, K, jb.out_mean, g * 8);
    const float inv_k = 1.0f / (float)K;
    for (long long e = groups * 8 + tid; e < n; e += stride) {
      const float x0 = member[0][e];
      float acc = 0.f;
      for (int k = 1; k < K; ++k) acc += member[k][e] - x0;
      jb.out_mean[e] = x0 + acc * inv_k;
    }
    return;
  }
  if constexpr (!FLAT) {
  if (jb.channels == 1) {
    const long long groups = (jb.num_pixels + 3) / 4;
    for (long long g = tid; g < groups; g += stride) {
      const long long pix0 = g * 4;
      const int npix = (int)min(4LL, jb.num_pixels - pix0);
      const bool vec = jb.vec_ok && npix == 4;
      if (K <= 6) reduce_pixels<1, 4, 6, false>(member, K, jb, pix0, npix, vec);
      else reduce_pixels<1, 4, 6, true>(member, K, jb, pix0, npix, vec);
    }
  } else if (jb.channels == 3) {
    const long long groups = (jb.num_pixels + 1) / 2;
    for (long long g = tid; g < groups; g += stride) {
      const long long pix0 = g * 2;
      const int npix = (int)min(2LL, jb.num_pixels - pix0);
      const bool vec = jb.vec_ok && npix == 2;
      if (K <= 5) reduce_pixels<3, 2, 5, false>(member, K, jb, pix0, npix, vec);
      else reduce_pixels<3, 2, 5, true>(member, K, jb, pix0, npix, vec);
    }
  } else {
    for (long long px = tid; px < jb.num_pixels; px += stride) reduce_pixel_any_c(member, K, jb, px);
  }
  }
}

static bool al16(const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

}  // namespace ub

extern "C" {

int ub_reduce_members_batched(const ub_reduce_job* jobs_host, int32_t num_jobs, int32_t num_members,
                              void* stream_v) {
  using namespace ub;
  UB_REQUIRE(jobs_host != nullptr && num_jobs >= 1, UB_ERR_BAD_ARG, "reduce_members: no jobs");
  UB_REQUIRE(num_jobs <= kMaxJobs, UB_ERR_UNSUPPORTED, "reduce_members: more than %d jobs per call", kMaxJobs);
  UB_REQUIRE(num_members >= 1 && num_members <= kMaxBatchMembers, UB_ERR_UNSUPPORTED,
             "reduce_members: num_members %d outside [1, %d]", num_members, kMaxBatchMembers);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  // two compacted job lists, one per kernel instantiation (a block of the wrong kind would only launch and exit)
  ReduceBatch flat{}, spread{};
  flat.num_members = spread.num_members = num_members;
  int n_flat = 0, n_spread = 0;
  long long flat_threads = 0, spread_threads = 0;
  for (int j = 0; j < num_jobs; ++j) {
    const ub_reduce_job& in = jobs_host[j];
    UB_REQUIRE(in.members_host != nullptr, UB_ERR_BAD_ARG, "reduce_members: job %d members_host is NULL", j);
    UB_REQUIRE(in.num_pixels >= 0 && in.channels >= 1, UB_ERR_BAD_ARG,
               "reduce_members: job %d bad shape N=%lld C=%d", j, (long long)in.num_pixels, in.channels);
    UB_REQUIRE(in.spread_mode >= UB_SPREAD_NONE && in.spread_mode <= UB_SPREAD_VAR, UB_ERR_BAD_ARG,
               "reduce_members: job %d bad spread_mode %d", j, in.spread_mode);
    UB_REQUIRE(in.spread_mode == UB_SPREAD_NONE || in.out_spread != nullptr, UB_ERR_BAD_ARG,
               "reduce_members: job %d spread requested but out_spread is NULL", j);
    bool vec_ok = al16(in.out_mean) && al16(in.out_spread);
    for (int k = 0; k < num_members; ++k) {
      UB_REQUIRE(in.members_host[k] != nullptr || in.num_pixels == 0, UB_ERR_BAD_ARG,
                 "reduce_members: job %d member %d is NULL", j, k);
      vec_ok = vec_ok && al16(in.members_host[k]);
    }
    const int mode = in.out_spread ? in.spread_mode : UB_SPREAD_NONE;
    const bool is_flat = mode == UB_SPREAD_NONE && vec_ok && in.out_mean != nullptr;
    ReduceBatch& b = is_flat ? flat : spread;
    const int slot = is_flat ? n_flat++ : n_spread++;
    for (int k = 0; k < num_members; ++k) b.member[slot][k] = in.members_host[k];
    b.job[slot].num_pixels = in.num_pixels;
    b.job[slot].channels = in.channels;
    b.job[slot].spread_mode = mode;
    b.job[slot].out_mean = in.out_mean;
    b.job[slot].out_spread = in.out_spread;
    b.job[slot].vec_ok = vec_ok ? 1 : 0;
    if (is_flat) {
      flat_threads = std::max(flat_threads, (long long)((in.num_pixels * in.channels + 7) / 8));
    } else {
      const long long t = in.channels == 1 ? (in.num_pixels + 3) / 4
                          : in.channels == 3 ? (in.num_pixels + 1) / 2 : in.num_pixels;
      spread_threads = std::max(spread_threads, t);
    }
  }
  const int sms = sm_count() > 0 ? sm_count() : 148;
  const long long cap = (long long)sms * 16;
  if (n_flat > 0 && flat_threads > 0) {
    dim3 grid((unsigned)std::min(cap, (flat_threads + 255) / 256), (unsigned)n_flat);
    reduce_members_batched_kernel<true><<<grid, 256, 0, stream>>>(flat);
  }
  if (n_spread > 0 && spread_threads > 0) {
    dim3 grid((unsigned)std::min(cap, (spread_threads + 255) / 256), (unsigned)n_spread);
    reduce_members_batched_kernel<false><<<grid, 256, 0, stream>>>(spread);
  }
  return check_launch("reduce_members");
}

int ub_reduce_members(const float* const* members_host, int32_t num_members, int64_t num_pixels,
                      int32_t channels, int32_t spread_mode, float* out_mean, float* out_spread,
                      void* stream_v) {
  using namespace ub;
  UB_REQUIRE(members_host != nullptr, UB_ERR_BAD_ARG, "reduce_members: members_host is NULL");
  // more members than one batched job carries: reduce in the single-job layout (same kernel)
  UB_REQUIRE(num_members >= 1 && num_members <= kMaxBatchMembers, UB_ERR_UNSUPPORTED,
             "reduce_members: num_members %d outside [1, %d]", num_members, kMaxBatchMembers);
  ub_reduce_job job;
  job.members_host = members_host;
  job.num_pixels = num_pixels;
  job.channels = channels;
  job.spread_mode = spread_mode;
  job.out_mean = out_mean;
  job.out_spread = out_spread;
  return ub_reduce_members_batched(&job, 1, num_members, stream_v);
}

}  // extern "C"
