// (A1, backward) Gradients of the fused ray compositor w.r.t. density, sample colours and beta -- the
// "next" row f1 of SURVEY.md section 8: lets `ns-train` run active-nerfacto through the fused path
// (losses: models/activenerfacto/activenerfacto_model.py:155-191; interlevel / distortion losses consume
// the `weights` output, so its incoming gradient is an input here).
//
// With dd_i = delta_i sigma_i, T_i = exp(-sum_{j<i} dd_j), w_i = (1 - exp(-dd_i)) T_i:
//   dL/dw_i   = G_i = g_w[i] + g_rgb . (c_i - bg) + g_acc + g_exp (s_i - e) / (acc + 1e-10)
//                     + 2 g_var w_i beta_i + g_dvar (s_i - depth)^2
//   dL/ddd_k  = G_k T_{k+1} - sum_{i>k} G_i w_i          (dw_i/ddd_i = T_{i+1}, dw_i/ddd_k = -w_i for k < i)
//   dL/dsigma_k = delta_k dL/ddd_k,  dL/dc_i = g_rgb w_i (+ g_rgb (1 - acc) for the last sample when it is the
//   background),  dL/dbeta_i = g_var w_i^2,  with g_var / g_dvar including the sqrt chain of rgb_std / depth_std.
// The median depth is computed under no_grad in the reference (activenerfacto_model.py:99-100) and enters as
// a constant; the expected-depth gradient is zero where the forward clipped (torch.clip backward).
// The forward is recomputed from the inputs (nothing but the per-ray depth and the chunk bounds is saved);
// same 4-lanes-per-ray register layout as the forward kernel, operands read straight from global memory.
#include "ub_common.cuh"

namespace ub {

struct CompositeBwdParams {
  const float *density, *deltas, *starts, *ends, *rgb, *beta;
  long long num_rays;
  int bg_mode;
  float bg[3];
  long long rays_per_chunk;
  const float* depth;         // [R] forward median depth
  const unsigned* chunk_ws;   // forward workspace: per chunk {max key(steps), max ~key(steps), ...}
  const float *g_rgb, *g_acc, *g_exp, *g_var, *g_std, *g_dvar, *g_dstd, *g_w;
  float *o_g_density, *o_g_rgb, *o_g_beta;
};

__device__ __forceinline__ float reduce4f(float v) {
  v += __shfl_xor_sync(FULL_MASK, v, 1);
  v += __shfl_xor_sync(FULL_MASK, v, 2);
  return v;
}

template <int S>
__global__ void __launch_bounds__(256) composite_rays_bwd_kernel(const CompositeBwdParams p) {
  constexpr int P = S / 4, V = P / 4;
  const int lane = threadIdx.x & 31;
  const int q = lane & 3, group_base = lane & ~3;
  const long long rays_per_warp = 8;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long num_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long num_groups = (p.num_rays + rays_per_warp - 1) / rays_per_warp;

  for (long long grp = warp_global; grp < num_groups; grp += num_warps) {
    const long long ray = grp * rays_per_warp + (lane >> 2);
    const bool active = ray < p.num_rays;
    const size_t off = (size_t)(active ? ray : 0) * S + q * P;

    float dd[P], dl[P], step[P], bt[P], col[3 * P], gw[P];
    {
      const float4* a = reinterpret_cast<const float4*>(p.density + off);
      const float4* b = reinterpret_cast<const float4*>(p.deltas + off);
      const float4* c = reinterpret_cast<const float4*>(p.starts + off);
      const float4* d = reinterpret_cast<const float4*>(p.ends + off);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float4 x = a[v], y = b[v], s0 = c[v], s1 = d[v];
        dl[4 * v + 0] = y.x; dl[4 * v + 1] = y.y; dl[4 * v + 2] = y.z; dl[4 * v + 3] = y.w;
        dd[4 * v + 0] = __fmul_rn(y.x, x.x); dd[4 * v + 1] = __fmul_rn(y.y, x.y);
        dd[4 * v + 2] = __fmul_rn(y.z, x.z); dd[4 * v + 3] = __fmul_rn(y.w, x.w);
        step[4 * v + 0] = __fadd_rn(s0.x, s1.x) * 0.5f; step[4 * v + 1] = __fadd_rn(s0.y, s1.y) * 0.5f;
        step[4 * v + 2] = __fadd_rn(s0.z, s1.z) * 0.5f; step[4 * v + 3] = __fadd_rn(s0.w, s1.w) * 0.5f;
      }
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float4 x = p.beta ? reinterpret_cast<const float4*>(p.beta + off)[v] : make_float4(0.f, 0.f, 0.f, 0.f);
        bt[4 * v + 0] = x.x; bt[4 * v + 1] = x.y; bt[4 * v + 2] = x.z; bt[4 * v + 3] = x.w;
        float4 g = p.g_w ? reinterpret_cast<const float4*>(p.g_w + off)[v] : make_float4(0.f, 0.f, 0.f, 0.f);
        gw[4 * v + 0] = g.x; gw[4 * v + 1] = g.y; gw[4 * v + 2] = g.z; gw[4 * v + 3] = g.w;
      }
      const float4* f = reinterpret_cast<const float4*>(p.rgb + off * 3);
#pragma unroll
      for (int v = 0; v < 3 * V; ++v) {
        const float4 x = f[v];
        col[4 * v + 0] = x.x; col[4 * v + 1] = x.y; col[4 * v + 2] = x.z; col[4 * v + 3] = x.w;
      }
    }

    // ---- forward recompute: transmittance, weights ----
    double pre[P];
    double run = 0.0;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      pre[i] = run;
      run += (double)dd[i];
    }
    {
      const double t0 = shfl_double(FULL_MASK, run, group_base | 0), t1 = shfl_double(FULL_MASK, run, group_base | 1);
      const double t2 = shfl_double(FULL_MASK, run, group_base | 2);
      const double o = q == 0 ? 0.0 : (q == 1 ? t0 : (q == 2 ? t0 + t1 : (t0 + t1) + t2));
#pragma unroll
      for (int i = 0; i < P; ++i) pre[i] += o;
    }
    float wgt[P], tnext[P];
    float acc = 0.f, e_num = 0.f, var = 0.f;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      const float trans = expf(-(float)pre[i]);
      const float em = expf(-dd[i]);
      wgt[i] = __fmul_rn(1.0f - em, trans);
      tnext[i] = trans * em;  // T_{i+1}
      acc += wgt[i];
      e_num = fmaf(wgt[i], step[i], e_num);
      var = fmaf(wgt[i] * wgt[i], bt[i], var);
    }
    acc = reduce4f(acc);
    e_num = reduce4f(e_num);
    var = reduce4f(var);
    const float depth = active ? p.depth[ray] : 0.f;
    float dvar = 0.f;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      const float t = step[i] - depth;
      dvar = fmaf(wgt[i], t * t, dvar);
    }
    dvar = reduce4f(dvar) + 1e-5f;
    const float A = acc + 1e-10f;
    const float e = e_num / A;

    // ---- per-ray incoming gradients ----
    float gr0 = 0.f, gr1 = 0.f, gr2 = 0.f, g_acc = 0.f, g_exp = 0.f, g_var = 0.f, g_dvar = 0.f;
    if (active) {
      if (p.g_rgb) { gr0 = p.g_rgb[ray * 3 + 0]; gr1 = p.g_rgb[ray * 3 + 1]; gr2 = p.g_rgb[ray * 3 + 2]; }
      if (p.g_acc) g_acc = p.g_acc[ray];
      if (p.g_exp) {
        const long long chunk = p.rays_per_chunk > 0 ? ray / p.rays_per_chunk : 0;
        const float lo = order_key_inv(~p.chunk_ws[chunk * 4 + 1]), hi = order_key_inv(p.chunk_ws[chunk * 4 + 0]);
        g_exp = (e >= lo && e <= hi) ? p.g_exp[ray] : 0.f;
      }
      if (p.g_var) g_var = p.g_var[ray];
      if (p.g_std) g_var += p.g_std[ray] / (2.0f * sqrtf(var));
      if (p.g_dvar) g_dvar = p.g_dvar[ray];
      if (p.g_dstd) g_dvar += p.g_dstd[ray] / (2.0f * sqrtf(dvar));
    }
    float b0 = __shfl_sync(FULL_MASK, col[3 * P - 3], group_base | 3);
    float b1 = __shfl_sync(FULL_MASK, col[3 * P - 2], group_base | 3);
    float b2 = __shfl_sync(FULL_MASK, col[3 * P - 1], group_base | 3);
    if (p.bg_mode == UB_BG_FIXED) { b0 = p.bg[0]; b1 = p.bg[1]; b2 = p.bg[2]; }
    if (p.bg_mode == UB_BG_NONE) { b0 = b1 = b2 = 0.f; }
    const float rem = 1.0f - acc;
    const float g_exp_a = g_exp / A;

    // ---- dL/dw_i, suffix sums of G_i w_i, gradients ----
    float G[P];
    double sfx[P];
    run = 0.0;
#pragma unroll
    for (int i = P - 1; i >= 0; --i) {
      const float t = step[i] - depth;
      float g = gw[i] + g_acc + gr0 * (col[3 * i + 0] - b0) + gr1 * (col[3 * i + 1] - b1) + gr2 * (col[3 * i + 2] - b2);
      g = fmaf(g_exp_a, step[i] - e, g);
      g = fmaf(2.0f * g_var * wgt[i], bt[i], g);
      g = fmaf(g_dvar, t * t, g);
      G[i] = g;
      sfx[i] = run;  // sum over the later samples of this lane
      run += (double)(g * wgt[i]);
    }
    {
      // exclusive suffix over the lanes of the ray: totals of the lanes q' > q
      const double t1 = shfl_double(FULL_MASK, run, group_base | 1), t2 = shfl_double(FULL_MASK, run, group_base | 2);
      const double t3 = shfl_double(FULL_MASK, run, group_base | 3);
      const double o = q == 3 ? 0.0 : (q == 2 ? t3 : (q == 1 ? t3 + t2 : (t3 + t2) + t1));
#pragma unroll
      for (int i = 0; i < P; ++i) sfx[i] += o;
    }
    if (active) {
      if (p.o_g_density) {
        float4* out = reinterpret_cast<float4*>(p.o_g_density + off);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          float r4[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i = 4 * v + k;
            r4[k] = dl[i] * (fmaf(G[i], tnext[i], -(float)sfx[i]));
          }
          out[v] = make_float4(r4[0], r4[1], r4[2], r4[3]);
        }
      }
      if (p.o_g_beta) {
        float4* out = reinterpret_cast<float4*>(p.o_g_beta + off);
#pragma unroll
        for (int v = 0; v < V; ++v)
          out[v] = make_float4(g_var * wgt[4 * v] * wgt[4 * v], g_var * wgt[4 * v + 1] * wgt[4 * v + 1],
                               g_var * wgt[4 * v + 2] * wgt[4 * v + 2], g_var * wgt[4 * v + 3] * wgt[4 * v + 3]);
      }
      if (p.o_g_rgb) {
        float gc[3 * P];
#pragma unroll
        for (int i = 0; i < P; ++i) {
          gc[3 * i + 0] = gr0 * wgt[i];
          gc[3 * i + 1] = gr1 * wgt[i];
          gc[3 * i + 2] = gr2 * wgt[i];
        }
        if (p.bg_mode == UB_BG_LAST_SAMPLE && q == 3) {
          gc[3 * P - 3] += gr0 * rem;
          gc[3 * P - 2] += gr1 * rem;
          gc[3 * P - 1] += gr2 * rem;
        }
        float4* out = reinterpret_cast<float4*>(p.o_g_rgb + off * 3);
#pragma unroll
        for (int v = 0; v < 3 * V; ++v) out[v] = make_float4(gc[4 * v], gc[4 * v + 1], gc[4 * v + 2], gc[4 * v + 3]);
      }
    }
  }
}

template <int S>
static int launch_bwd(const CompositeBwdParams& p, cudaStream_t stream) {
  const long long groups = (p.num_rays + 7) / 8;
  long long blocks = (groups + 7) / 8;
  const long long cap = (long long)(sm_count() > 0 ? sm_count() : 148) * 8;
  if (blocks > cap) blocks = cap;
  composite_rays_bwd_kernel<S><<<(unsigned)blocks, 256, 0, stream>>>(p);
  return check_launch("composite_rays_backward");
}

}  // namespace ub

extern "C" int ub_composite_rays_backward(const ub_composite_rays_bwd_args* a, void* stream_v) {
  using namespace ub;
  UB_REQUIRE(a != nullptr, UB_ERR_BAD_ARG, "composite_rays_backward: args is NULL");
  UB_REQUIRE(a->num_rays >= 0 && a->num_samples >= 1, UB_ERR_BAD_ARG, "composite_rays_backward: bad shape");
  if (a->num_rays == 0) return UB_OK;
  UB_REQUIRE(a->density && a->deltas && a->starts && a->ends && a->rgb && a->depth, UB_ERR_BAD_ARG,
             "composite_rays_backward: forward inputs and the forward median depth must be non-NULL");
  UB_REQUIRE(a->g_expected_depth == nullptr || a->chunk_workspace != nullptr, UB_ERR_BAD_ARG,
             "composite_rays_backward: the expected-depth gradient needs the forward workspace (clip bounds)");
  UB_REQUIRE((a->g_rgb_var == nullptr && a->g_rgb_std == nullptr && a->out_g_beta == nullptr) || a->beta != nullptr,
             UB_ERR_BAD_ARG, "composite_rays_backward: beta gradients requested without beta");
  const void* ptrs[] = {a->density, a->deltas, a->starts, a->ends, a->rgb, a->beta, a->g_weights,
                        a->out_g_density, a->out_g_rgb, a->out_g_beta};
  for (const void* q : ptrs)
    UB_REQUIRE(q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15u) == 0, UB_ERR_UNSUPPORTED,
               "composite_rays_backward: per-sample tensors must be 16-byte aligned");
  CompositeBwdParams p{};
  p.density = a->density; p.deltas = a->deltas; p.starts = a->starts; p.ends = a->ends; p.rgb = a->rgb;
  p.beta = a->beta;
  p.num_rays = a->num_rays;
  p.bg_mode = a->background_mode;
  p.bg[0] = a->background_rgb[0]; p.bg[1] = a->background_rgb[1]; p.bg[2] = a->background_rgb[2];
  p.rays_per_chunk = a->rays_per_chunk;
  p.depth = a->depth;
  p.chunk_ws = static_cast<const unsigned*>(a->chunk_workspace);
  p.g_rgb = a->g_rgb; p.g_acc = a->g_accumulation; p.g_exp = a->g_expected_depth; p.g_var = a->g_rgb_var;
  p.g_std = a->g_rgb_std; p.g_dvar = a->g_depth_var; p.g_dstd = a->g_depth_std; p.g_w = a->g_weights;
  p.o_g_density = a->out_g_density; p.o_g_rgb = a->out_g_rgb; p.o_g_beta = a->out_g_beta;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  switch (a->num_samples) {
    case 16: return launch_bwd<16>(p, stream);
    case 32: return launch_bwd<32>(p, stream);
    case 48: return launch_bwd<48>(p, stream);
    case 64: return launch_bwd<64>(p, stream);
    case 96: return launch_bwd<96>(p, stream);
    default:
      set_error("composite_rays_backward: num_samples %d not supported (16, 32, 48, 64, 96)", a->num_samples);
      return UB_ERR_UNSUPPORTED;
  }
}
