// (A1') Mean volume-rendering weights over K Gaussian density draws (nerfacto-laplace, sampled density).
//
// Reference arithmetic (file:line under /root/reference/nerfuncertainty):
//   models/laplace/laplace_model.py:486-507
//     density_std = max(sqrt(density_var), 1e-10)  (NaN -> 1e-10)
//     sampled = relu(Normal(density, density_std).sample((100,)))      # [100, R, S, 1] materialised
//     weights = vmap(get_weights)(sampled).mean(dim=0)
// The reference materialises 100 x [R,S] tensors (629 MB per 32768-ray chunk); here a warp owns a ray,
// loops over the draws, keeps the running mean in registers and writes [R,S] once.  Draws are either
// read from a caller-provided standard-normal tensor [K,R,S] (bit-comparable with the oracle) or
// generated in-kernel with Philox4x32-10 (statistical parity only -- torch's global generator cannot be
// reproduced).
#include <curand_kernel.h>

#include "ub_common.cuh"

namespace ub {

constexpr int kMaxSlots = 8;  // samples per lane: supports S <= 256

__device__ __forceinline__ double warp_inclusive_scan_f64(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double up = shfl_up_double(FULL_MASK, v, o);
    if (lane >= o) v += up;
  }
  return v;
}

__global__ void __launch_bounds__(256)
sampled_weights_kernel(const float* __restrict__ density, const float* __restrict__ density_var,
                       const float* __restrict__ deltas, const float* __restrict__ noise,
                       long long num_rays, int S, int K, unsigned long long seed,
                       float* __restrict__ out_weights) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long num_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int slots = (S + 31) / 32;
  for (long long ray = warp_global; ray < num_rays; ray += num_warps) {
    const size_t row = (size_t)ray * S;
    float mu[kMaxSlots], sd[kMaxSlots], dl[kMaxSlots], acc[kMaxSlots];
#pragma unroll
    for (int j = 0; j < kMaxSlots; ++j) {
      const int i = j * 32 + lane;
      const bool ok = j < slots && i < S;
      mu[j] = ok ? density[row + i] : 0.f;
      dl[j] = ok ? deltas[row + i] : 0.f;
      float s = ok ? sqrtf(density_var[row + i]) : 0.f;
      s = fmaxf(s, 1e-10f);          // torch.maximum(std, 1e-10); fmaxf also maps NaN -> 1e-10 (:493-496)
      sd[j] = s;
      acc[j] = 0.f;
    }
    curandStatePhilox4_32_10_t rng;
    if (noise == nullptr) curand_init(seed, (unsigned long long)(ray * 32 + lane), 0ULL, &rng);
    for (int k = 0; k < K; ++k) {
      double carry = 0.0;
#pragma unroll
      for (int j = 0; j < kMaxSlots; ++j) {
        if (j < slots) {
          const int i = j * 32 + lane;
          const bool ok = i < S;
          float eps = 0.f;
          if (noise != nullptr) {
            if (ok) eps = noise[((size_t)k * num_rays + ray) * S + i];
          } else {
            eps = curand_normal(&rng);
          }
          // Normal(loc, scale).sample() == loc + scale * eps (fp32, product rounded first), then relu
          const float smp = fmaxf(__fadd_rn(__fmul_rn(eps, sd[j]), mu[j]), 0.f);
          const float dd = ok ? __fmul_rn(dl[j], smp) : 0.f;
          const double incl = warp_inclusive_scan_f64((double)dd, lane);
          const double prev = shfl_up_double(FULL_MASK, incl, 1);
          const double excl = lane == 0 ? carry : carry + prev;
          carry += shfl_double(FULL_MASK, incl, 31);
          if (ok) {
            const float w = nan_to_num((1.0f - expf(-dd)) * expf(-(float)excl));
            acc[j] += w;
          }
        }
      }
    }
    const float kf = (float)K;
#pragma unroll
    for (int j = 0; j < kMaxSlots; ++j) {
      const int i = j * 32 + lane;
      if (j < slots && i < S) out_weights[row + i] = acc[j] / kf;
    }
  }
}

}  // namespace ub

extern "C" int ub_average_sampled_weights(const float* density, const float* density_var,
                                          const float* deltas, const float* noise, int64_t num_rays,
                                          int32_t num_samples, int32_t num_draws, uint64_t seed,
                                          float* out_weights, void* stream_v) {
  using namespace ub;
  UB_REQUIRE(num_rays >= 0 && num_samples >= 1 && num_draws >= 1, UB_ERR_BAD_ARG,
             "average_sampled_weights: bad sizes");
  UB_REQUIRE(num_samples <= 32 * kMaxSlots, UB_ERR_UNSUPPORTED,
             "average_sampled_weights: num_samples must be <= %d", 32 * kMaxSlots);
  if (num_rays == 0) return UB_OK;
  UB_REQUIRE(density && density_var && deltas && out_weights, UB_ERR_BAD_ARG,
             "average_sampled_weights: NULL pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  long long blocks = (num_rays + 7) / 8;
  const long long cap = (long long)(sm_count() > 0 ? sm_count() : 148) * 8;
  if (blocks > cap) blocks = cap;
  sampled_weights_kernel<<<(unsigned)blocks, 256, 0, stream>>>(density, density_var, deltas, noise, num_rays,
                                                              num_samples, num_draws, seed, out_weights);
  return check_launch("average_sampled_weights");
}
