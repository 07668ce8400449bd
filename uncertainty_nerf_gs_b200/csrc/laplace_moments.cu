// (A3) Last-layer diagonal-Laplace MC moments for a linear head out = act(x W^T + b).
//
// Reference arithmetic (file:line under /root/reference/nerfuncertainty):
//   models/laplace/laplace_field.py:545-565   loop over the sampled parameter vectors:
//        pred = act(Linear_s(x)); pred_mu += pred; pred_mu2 += pred**2; /= n; sigma2 = mu2 - mu^2
//   callers :331-339 (density head 64 -> 1, trunc_exp) and :468-476 (rgb head 64 -> 3, sigmoid)
//
// This sub-step is fp32-FMA bound (2*H*out_dim*n_samples flop per point vs 4*H bytes), not HBM
// bound.  v1: all sampled parameters of the head are staged once per CTA in shared memory
// (100 x 195 floats = 78 KB for the rgb head); each thread keeps the H features of kPts points in
// registers and walks the samples with broadcast LDS.128 weight reads; the moments accumulate in
// float32 in sample order exactly like the reference loop.
#include "ub_common.cuh"

#include <stdlib.h>

namespace ub {

// tensor-core path (laplace_moments_tc.cu); UB_ERR_UNSUPPORTED when the shape does not fit it
int launch_laplace_tc(const float* x, long long num_points, const float* params, int n_samples, int out_dim, int act,
                      float* o_mean, float* o_mean2, float* o_sigma2, cudaStream_t stream);

constexpr int kLapThreads = 128;
constexpr int kLapH = 64;

__device__ __forceinline__ float apply_act(float z, int act) {
  if (act == UB_ACT_SIGMOID) return 1.0f / (1.0f + expf(-z));
  if (act == UB_ACT_EXP) return expf(z);
  return z;
}

template <int O, int PTS>
__global__ void __launch_bounds__(kLapThreads)
laplace_moments_kernel(const float* __restrict__ x, long long num_points,
                       const float* __restrict__ params, int n_samples, int samples_per_fill,
                       int act, float* __restrict__ o_mean, float* __restrict__ o_mean2,
                       float* __restrict__ o_sigma2) {
  extern __shared__ __align__(16) float s_par[];  // [samples_per_fill][O*H + O (padded to 4)]
  constexpr int H = kLapH;
  constexpr int NP = O * H + O;
  constexpr int NPP = (NP + 3) & ~3;  // row stride in shared memory, 16-byte aligned rows

  const long long groups = (num_points + PTS - 1) / PTS;
  for (long long g0 = (long long)blockIdx.x * kLapThreads; g0 < groups;
       g0 += (long long)gridDim.x * kLapThreads) {
    const long long g = g0 + threadIdx.x;
    float xr[PTS][H];
    bool ok[PTS];
#pragma unroll
    for (int q = 0; q < PTS; ++q) {
      const long long pt = g * PTS + q;
      ok[q] = g < groups && pt < num_points;
      const float4* row = reinterpret_cast<const float4*>(x + (size_t)(ok[q] ? pt : 0) * H);
#pragma unroll
      for (int j = 0; j < H / 4; ++j) {
        const float4 v = row[j];
        xr[q][4 * j] = v.x;
        xr[q][4 * j + 1] = v.y;
        xr[q][4 * j + 2] = v.z;
        xr[q][4 * j + 3] = v.w;
      }
    }
    float mu[PTS][O], mu2[PTS][O];
#pragma unroll
    for (int q = 0; q < PTS; ++q)
#pragma unroll
      for (int o = 0; o < O; ++o) {
        mu[q][o] = 0.f;
        mu2[q][o] = 0.f;
      }

    for (int s0 = 0; s0 < n_samples; s0 += samples_per_fill) {
      const int ns = min(samples_per_fill, n_samples - s0);
      __syncthreads();  // previous fill fully consumed
      for (int i = threadIdx.x; i < ns * NP; i += kLapThreads) {
        const int s = i / NP, k = i - s * NP;
        s_par[s * NPP + k] = params[(size_t)(s0 + s) * NP + k];
      }
      __syncthreads();
      for (int s = 0; s < ns; ++s) {
        const float* wrow = s_par + s * NPP;
        float z[PTS][O];
#pragma unroll
        for (int q = 0; q < PTS; ++q)
#pragma unroll
          for (int o = 0; o < O; ++o) z[q][o] = 0.f;
#pragma unroll
        for (int o = 0; o < O; ++o) {
          const float4* w4 = reinterpret_cast<const float4*>(wrow + o * H);
#pragma unroll
          for (int j = 0; j < H / 4; ++j) {
            const float4 w = w4[j];
#pragma unroll
            for (int q = 0; q < PTS; ++q) {
              z[q][o] = fmaf(xr[q][4 * j + 0], w.x, z[q][o]);
              z[q][o] = fmaf(xr[q][4 * j + 1], w.y, z[q][o]);
              z[q][o] = fmaf(xr[q][4 * j + 2], w.z, z[q][o]);
              z[q][o] = fmaf(xr[q][4 * j + 3], w.w, z[q][o]);
            }
          }
        }
#pragma unroll
        for (int o = 0; o < O; ++o) {
          const float b = wrow[O * H + o];
#pragma unroll
          for (int q = 0; q < PTS; ++q) {
            const float y = apply_act(z[q][o] + b, act);
            mu[q][o] += y;
            mu2[q][o] += y * y;
          }
        }
      }
    }
    const float nf = (float)n_samples;
#pragma unroll
    for (int q = 0; q < PTS; ++q) {
      if (!ok[q]) continue;
      const long long pt = g * PTS + q;
#pragma unroll
      for (int o = 0; o < O; ++o) {
        const float m = mu[q][o] / nf, m2 = mu2[q][o] / nf;
        if (o_mean) o_mean[pt * O + o] = m;
        if (o_mean2) o_mean2[pt * O + o] = m2;
        if (o_sigma2) o_sigma2[pt * O + o] = m2 - m * m;
      }
    }
  }
}

template <int O, int PTS>
static int launch_laplace(const float* x, long long num_points, const float* params, int n_samples,
                          int act, float* o_mean, float* o_mean2, float* o_sigma2,
                          cudaStream_t stream) {
  constexpr int NP = O * kLapH + O;
  constexpr int NPP = (NP + 3) & ~3;
  const size_t max_smem = 200 * 1024;
  int per_fill = (int)(max_smem / (NPP * sizeof(float)));
  if (per_fill > n_samples) per_fill = n_samples;
  const size_t smem = (size_t)per_fill * NPP * sizeof(float);
  auto kern = laplace_moments_kernel<O, PTS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("laplace_ll_moments: cannot reserve %zu B shared memory (%s)", smem, cudaGetErrorString(e));
    return UB_ERR_LAUNCH;
  }
  const long long groups = (num_points + PTS - 1) / PTS;
  long long blocks = (groups + kLapThreads - 1) / kLapThreads;
  // the parameter table is re-staged per grid-stride iteration: keep the grid persistent-sized
  // (three blocks per SM where the table is small: the density head's 27 KB; 162 registers x 128 threads allow no more)
  const long long cap = (long long)(sm_count() > 0 ? sm_count() : 148) * (smem > 100 * 1024 ? 1 : (smem > 60 * 1024 ? 2 : 3));
  if (blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, kLapThreads, smem, stream>>>(x, num_points, params, n_samples, per_fill, act,
                                                        o_mean, o_mean2, o_sigma2);
  return check_launch("laplace_ll_moments");
}

}  // namespace ub

extern "C" int ub_laplace_ll_moments(const float* x, int64_t num_points, int32_t hidden, int32_t out_dim,
                                     const float* sampled_params, int32_t n_samples, int32_t activation,
                                     float* out_mean, float* out_mean2, float* out_sigma2,
                                     void* stream_v) {
  using namespace ub;
  UB_REQUIRE(num_points >= 0 && n_samples >= 1, UB_ERR_BAD_ARG, "laplace_ll_moments: bad sizes");
  UB_REQUIRE(hidden == kLapH, UB_ERR_UNSUPPORTED,
             "laplace_ll_moments: hidden must be %d (nerfacto head width), got %d", kLapH, hidden);
  UB_REQUIRE(out_dim == 1 || out_dim == 3, UB_ERR_UNSUPPORTED,
             "laplace_ll_moments: out_dim must be 1 (density) or 3 (rgb), got %d", out_dim);
  const bool no_tc = (activation & UB_ACT_FLAG_NO_TENSOR_CORES) != 0;
  activation &= ~UB_ACT_FLAG_NO_TENSOR_CORES;
  UB_REQUIRE(activation >= UB_ACT_IDENTITY && activation <= UB_ACT_EXP, UB_ERR_BAD_ARG,
             "laplace_ll_moments: bad activation %d", activation);
  if (num_points == 0) return UB_OK;
  UB_REQUIRE(x && sampled_params, UB_ERR_BAD_ARG, "laplace_ll_moments: NULL input");
  UB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15u) == 0, UB_ERR_UNSUPPORTED,
             "laplace_ll_moments: x must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  // tcgen05 3xTF32 path unless UB_LAPLACE_FMA=1 or the caller's flag forces the fp32-FMA kernel, which also takes the
  // shapes the tensor-core kernel does not (> 304 columns).  rgb head: always.  Density head: the same GEMM with the
  // draws as its N dimension runs 2.4x faster there (0.20 vs 0.49 ms per 1.05 M points), but the tensor core adds
  // into its float32 accumulator with truncation and exp turns that absolute error of the pre-activation into a
  // relative error of the output: with one accumulator 1.1e-5 on E[y] at v ~ 7, with the products spread over three
  // accumulators (what the kernel does) 6.5e-6 on E[y] but 1.5e-5 on E[y^2] -- outside the 1e-5 contract, which the
  // FMA kernel keeps (7e-6 worst; sigmoid damps the error by >= 4).  Hence opt-in: UB_LAPLACE_TC_DENSITY=1.
  static const bool force_fma = [] { const char* e = getenv("UB_LAPLACE_FMA"); return e && atoi(e) != 0; }();
  static const bool tc_density = [] { const char* e = getenv("UB_LAPLACE_TC_DENSITY"); return e && atoi(e) != 0; }();
  if (!force_fma && !no_tc && (out_dim == 3 || tc_density)) {
    const int rc = launch_laplace_tc(x, num_points, sampled_params, n_samples, out_dim, activation, out_mean, out_mean2,
                                     out_sigma2, stream);
    if (rc != UB_ERR_UNSUPPORTED) return rc;
  }
  if (out_dim == 3)
    return launch_laplace<3, 2>(x, num_points, sampled_params, n_samples, activation, out_mean, out_mean2, out_sigma2,
                                stream);
  return launch_laplace<1, 2>(x, num_points, sampled_params, n_samples, activation, out_mean, out_mean2, out_sigma2,
                              stream);
}
