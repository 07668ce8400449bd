// (B) Fused per-pixel mean / spread across K ensemble members or MC-dropout passes.
//
// Reference arithmetic (file:line under /root/reference/nerfuncertainty):
//   models/mcdropout/mcdropout_models.py:121-126   stack -> mean(0); std(0).mean(-1)[..., None]
//   models/ensemble/ensemble_pipeline.py:159-190   same, plus var(0).mean(-1) for the epistemic term
//
// One streaming pass: every member image is read exactly once through its own pointer (no
// torch.stack copy), each thread owns 4 consecutive pixels (128-bit loads), and the K values of an
// element never leave registers.  The spread uses shifted sums in float64 (shift = first member),
// which is exact for identical members and has no mean^2/var cancellation; torch's CPU kernel
// (Welford with float64 accumulators) agrees to float32 rounding.
#include "ub_common.cuh"

namespace ub {

struct ReduceParams {
  const float* member[UB_MAX_MEMBERS];
  int num_members;
  long long num_pixels;
  int spread_mode;
  float* out_mean;
  float* out_spread;
};

template <int C>
__device__ __forceinline__ void reduce_pixels(const ReduceParams& p, long long pix0, int npix,
                                              bool vec) {
  // npix pixels starting at pix0 (npix == 4 on the vector path)
  constexpr int E = 4 * C;  // elements per thread
  const int K = p.num_members;
  float x0[E];
  double s1[E], s2[E];
  const long long e0 = pix0 * C;
  const int ne = npix * C;
  if (vec) {
    const float4* src = reinterpret_cast<const float4*>(p.member[0] + e0);
#pragma unroll
    for (int j = 0; j < C; ++j) {
      float4 v = src[j];
      x0[4 * j + 0] = v.x;
      x0[4 * j + 1] = v.y;
      x0[4 * j + 2] = v.z;
      x0[4 * j + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < E; ++i) x0[i] = i < ne ? p.member[0][e0 + i] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < E; ++i) {
    s1[i] = 0.0;
    s2[i] = 0.0;
  }
  for (int k = 1; k < K; ++k) {
    float x[E];
    if (vec) {
      const float4* src = reinterpret_cast<const float4*>(p.member[k] + e0);
#pragma unroll
      for (int j = 0; j < C; ++j) {
        float4 v = src[j];
        x[4 * j + 0] = v.x;
        x[4 * j + 1] = v.y;
        x[4 * j + 2] = v.z;
        x[4 * j + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < E; ++i) x[i] = i < ne ? p.member[k][e0 + i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const double d = (double)x[i] - (double)x0[i];
      s1[i] += d;
      s2[i] += d * d;
    }
  }
  const double inv_k = 1.0 / (double)K;
  float mean[E];
#pragma unroll
  for (int i = 0; i < E; ++i) mean[i] = (float)((double)x0[i] + s1[i] * inv_k);
  if (p.out_mean) {
    if (vec) {
      float4* dst = reinterpret_cast<float4*>(p.out_mean + e0);
#pragma unroll
      for (int j = 0; j < C; ++j)
        dst[j] = make_float4(mean[4 * j], mean[4 * j + 1], mean[4 * j + 2], mean[4 * j + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < E; ++i)
        if (i < ne) p.out_mean[e0 + i] = mean[i];
    }
  }
  if (p.spread_mode != UB_SPREAD_NONE && p.out_spread) {
    float spread[4];
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int i = px * C + c;
        double var = (s2[i] - s1[i] * s1[i] * inv_k) / (double)(K - 1);  // K == 1 -> NaN like torch
        if (var < 0.0) var = 0.0;
        const float v = p.spread_mode == UB_SPREAD_STD ? (float)sqrt(var) : (float)var;
        acc = c == 0 ? v : acc + v;
      }
      spread[px] = C == 1 ? acc : acc / (float)C;
    }
    if (vec) {
      *reinterpret_cast<float4*>(p.out_spread + pix0) =
          make_float4(spread[0], spread[1], spread[2], spread[3]);
    } else {
#pragma unroll
      for (int px = 0; px < 4; ++px)
        if (px < npix) p.out_spread[pix0 + px] = spread[px];
    }
  }
}

template <int C>
__global__ void __launch_bounds__(256) reduce_members_kernel(const ReduceParams p, int vec_ok) {
  const long long groups = (p.num_pixels + 3) / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
    const long long pix0 = g * 4;
    const int npix = (int)min(4LL, p.num_pixels - pix0);
    reduce_pixels<C>(p, pix0, npix, vec_ok && npix == 4);
  }
}

// Any channel count: one thread per pixel, scalar loads.
__global__ void __launch_bounds__(256) reduce_members_any_c(const ReduceParams p, int C) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int K = p.num_members;
  const double inv_k = 1.0 / (double)K;
  for (long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x; px < p.num_pixels; px += stride) {
    float acc = 0.f;
    for (int c = 0; c < C; ++c) {
      const long long e = px * C + c;
      const float x0 = p.member[0][e];
      double s1 = 0.0, s2 = 0.0;
      for (int k = 1; k < K; ++k) {
        const double d = (double)p.member[k][e] - (double)x0;
        s1 += d;
        s2 += d * d;
      }
      if (p.out_mean) p.out_mean[e] = (float)((double)x0 + s1 * inv_k);
      double var = (s2 - s1 * s1 * inv_k) / (double)(K - 1);
      if (var < 0.0) var = 0.0;
      const float v = p.spread_mode == UB_SPREAD_STD ? (float)sqrt(var) : (float)var;
      acc = c == 0 ? v : acc + v;
    }
    if (p.spread_mode != UB_SPREAD_NONE && p.out_spread) p.out_spread[px] = C == 1 ? acc : acc / (float)C;
  }
}

}  // namespace ub

extern "C" int ub_reduce_members(const float* const* members_host, int32_t num_members,
                                 int64_t num_pixels, int32_t channels, int32_t spread_mode,
                                 float* out_mean, float* out_spread, void* stream_v) {
  using namespace ub;
  UB_REQUIRE(members_host != nullptr, UB_ERR_BAD_ARG, "reduce_members: members_host is NULL");
  UB_REQUIRE(num_members >= 1 && num_members <= UB_MAX_MEMBERS, UB_ERR_UNSUPPORTED,
             "reduce_members: num_members %d outside [1, %d]", num_members, UB_MAX_MEMBERS);
  UB_REQUIRE(num_pixels >= 0 && channels >= 1, UB_ERR_BAD_ARG, "reduce_members: bad shape N=%lld C=%d",
             (long long)num_pixels, channels);
  UB_REQUIRE(spread_mode >= UB_SPREAD_NONE && spread_mode <= UB_SPREAD_VAR, UB_ERR_BAD_ARG,
             "reduce_members: bad spread_mode %d", spread_mode);
  UB_REQUIRE(spread_mode == UB_SPREAD_NONE || out_spread != nullptr, UB_ERR_BAD_ARG,
             "reduce_members: spread requested but out_spread is NULL");
  if (num_pixels == 0) return UB_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  ReduceParams p{};
  bool vec_ok = true;
  for (int k = 0; k < num_members; ++k) {
    UB_REQUIRE(members_host[k] != nullptr, UB_ERR_BAD_ARG, "reduce_members: member %d is NULL", k);
    p.member[k] = members_host[k];
    vec_ok = vec_ok && (reinterpret_cast<uintptr_t>(members_host[k]) & 15u) == 0;
  }
  vec_ok = vec_ok && (out_mean == nullptr || (reinterpret_cast<uintptr_t>(out_mean) & 15u) == 0) &&
           (out_spread == nullptr || (reinterpret_cast<uintptr_t>(out_spread) & 15u) == 0);
  p.num_members = num_members;
  p.num_pixels = num_pixels;
  p.spread_mode = out_spread ? spread_mode : UB_SPREAD_NONE;
  p.out_mean = out_mean;
  p.out_spread = out_spread;
  const int sms = sm_count() > 0 ? sm_count() : 148;
  const long long cap = (long long)sms * 8;
  if (channels == 1 || channels == 3) {
    const long long groups = (num_pixels + 3) / 4;
    long long blocks = (groups + 255) / 256;
    if (blocks > cap) blocks = cap;
    if (channels == 1)
      reduce_members_kernel<1><<<(unsigned)blocks, 256, 0, stream>>>(p, vec_ok ? 1 : 0);
    else
      reduce_members_kernel<3><<<(unsigned)blocks, 256, 0, stream>>>(p, vec_ok ? 1 : 0);
  } else {
    long long blocks = (num_pixels + 255) / 256;
    if (blocks > cap) blocks = cap;
    reduce_members_any_c<<<(unsigned)blocks, 256, 0, stream>>>(p, channels);
  }
  return check_launch("reduce_members");
}
