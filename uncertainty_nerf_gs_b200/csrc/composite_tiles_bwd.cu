// (f1) Backward of the tile compositor: gradients of the composited planes / alpha with respect to the 2-D
// centres, conics, opacities and colour planes of the splats, for training active-splatfacto through the
// fused pass (reference losses: models/activesplatfacto/activesplatfacto_model.py:369-441 read
// outputs["rgb"] and outputs["uncertainty"], i.e. the rgb and beta planes of one fused pass).
//
// The arithmetic restates the published gsplat 0.1.11 `rasterize_backward` (not vendored: parity unpinned;
// checked against torch autograd through the oracle's differentiable rasteriser).  With
//   out_c = sum_i c_i a_i T_i + T_final bg_c,   alpha_out = 1 - T_final,   T_i = prod_{k<i} (1 - a_k),
// walking a pixel's list back to front with S_c = sum_{k>i} c_k a_k T_k:
//   dL/dc_i  = a_i T_i v_out_c
//   dL/da_i  = sum_c v_out_c (c_i T_i - S_c / (1 - a_i)) + (v_alpha - sum_c bg_c v_out_c) T_final / (1 - a_i)
//   a_i = min(0.999, o exp(-sigma)):  dL/dsigma = -o e^{-sigma} dL/da_i,  dL/do = e^{-sigma} dL/da_i  (0 if clamped)
//   sigma = 0.5 (A dx^2 + C dy^2) + B dx dy, dx = x_g - px:  dL/dA = 0.5 dx^2 dL/dsigma, dL/dB = dx dy dL/dsigma,
//   dL/dC = 0.5 dy^2 dL/dsigma,  dL/dx = (A dx + B dy) dL/dsigma,  dL/dy = (B dx + C dy) dL/dsigma.
//
// One CTA per 16x16 tile, one thread per pixel.  Phase 1 replays the forward geometry test to find each
// pixel's final transmittance and the end of its contributing range (so the forward pass needs no extra
// outputs); phase 2 walks the tile's list backwards in staged batches of 256 splats.  Per splat the
// 6 + CH partial gradients are reduced over the warp with shuffles (skipped when no lane of the warp is
// touched), accumulated per batch entry in shared memory and flushed with one global atomicAdd per value
// per (splat, tile) intersection.  float32 atomics: the summation order over tiles is not fixed.
#include "ub_common.cuh"

namespace ub {

constexpr int kBwdThreads = UB_TILE * UB_TILE;
constexpr int kBwdPlanes = UB_MAX_SPLAT_PLANES;

struct TileBwdParams {
  const float* xys;
  const float* conics;
  const float* opacities;
  const float* plane[kBwdPlanes];
  int plane_ch[kBwdPlanes];
  int plane_off[kBwdPlanes];
  int num_planes;
  const int32_t* gaussian_ids;
  const int32_t* tile_bins;
  int height, width, tiles_x;
  float background[UB_MAX_SPLAT_CHANNELS];
  const float* v_out[kBwdPlanes];  // [H, W, ch_p] or NULL (zero gradient)
  const float* v_alpha;            // [H, W] or NULL
  float* v_xys;                    // [G, 2]
  float* v_conics;                 // [G, 3]
  float* v_opacities;              // [G]
  float* v_plane[kBwdPlanes];      // [G, ch_p] or NULL
};

// Sum P (8 or 16) per-lane values over the warp with the value-halving butterfly: at every step a lane keeps half
// of its values and hands the other half to its partner, so P + 1 shuffles replace 5 P.  Afterwards the lanes
// (l, l ^ 1) both hold the warp total of value index(l); returns that total and sets `index`.
template <int P>
__device__ __forceinline__ float warp_sum_scatter(float (&v)[P], int lane, int& index) {
  static_assert(P == 8 || P == 16, "padded value count");
  constexpr int kFirst = P == 16 ? 16 : 8;  // lane bit used by the first halving step
  int idx = 0;
#pragma unroll
  for (int half = P / 2, bit = kFirst; half >= 1; half >>= 1, bit >>= 1) {
    const bool upper = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? v[i] : v[i + half];
      const float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(FULL_MASK, send, bit);
    }
    idx += upper ? half : 0;
  }
  // the remaining lane bits (1 for P = 16; 1, 2 for P = 8... folded below) hold partial sums of the same value
  float r = v[0];
#pragma unroll
  for (int bit = (P == 16 ? 1 : 1); bit >= 1; bit >>= 1) r += __shfl_xor_sync(FULL_MASK, r, bit);
  if (P == 8) r += __shfl_xor_sync(FULL_MASK, r, 16);
  index = idx;
  return r;
}

template <int CH>
__global__ void __launch_bounds__(kBwdThreads) composite_tiles_bwd_kernel(const TileBwdParams p) {
  constexpr int NCOLV = (CH + 2 + 3) / 4;
  constexpr int NV = 6 + CH;  // x, y, A, B, C, opacity, colours
  __shared__ float4 s_geo[kBwdThreads];
  __shared__ float4 s_rec[NCOLV][kBwdThreads];
  __shared__ int s_gid[kBwdThreads];
  __shared__ float4 s_box[kBwdThreads];  // warp-level cull, see splat_reach_box (ub_common.cuh)
  __shared__ unsigned char s_list[kBwdThreads / 32][kBwdThreads];  // per warp: staged splats that can touch its block
  __shared__ float s_acc[kBwdThreads][NV + 1];
  __shared__ int s_end;

  const int tile = blockIdx.y * p.tiles_x + blockIdx.x;
  int ti, tj;
  tile_pixel_of_thread(threadIdx.x, ti, tj);
  const int i = blockIdx.y * UB_TILE + ti, j = blockIdx.x * UB_TILE + tj;
  const float px = (float)j + 0.5f, py = (float)i + 0.5f;
  const bool inside = i < p.height && j < p.width;
  const int lo = p.tile_bins[2 * tile + 0], hi = p.tile_bins[2 * tile + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (hi <= lo) return;
  // pixel-centre box of this warp's 8 x 4 block
  const float blk_x_lo = (float)(blockIdx.x * UB_TILE + 8 * ((threadIdx.x >> 5) & 1)) + 0.5f, blk_x_hi = blk_x_lo + 7.0f;
  const float blk_y_lo = (float)(blockIdx.y * UB_TILE + 4 * (threadIdx.x >> 6)) + 0.5f, blk_y_hi = blk_y_lo + 3.0f;

  auto stage = [&](int batch, int limit) {
    const int idx = batch + threadIdx.x;
    if (idx < limit) {
      const int g = p.gaussian_ids[idx];
      const float2 xy = *reinterpret_cast<const float2*>(p.xys + 2 * (size_t)g);
      const float ca = p.conics[3 * (size_t)g + 0], cb = p.conics[3 * (size_t)g + 1], cc = p.conics[3 * (size_t)g + 2];
      float col[4 * NCOLV - 2];
#pragma unroll
      for (int c = 0; c < 4 * NCOLV - 2; ++c) col[c] = 0.f;
#pragma unroll
      for (int pl = 0; pl < kBwdPlanes; ++pl) {
        if (pl < p.num_planes) {
          const float* src = p.plane[pl] + (size_t)g * p.plane_ch[pl];
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const int k = c - p.plane_off[pl];
            if (k >= 0 && k < p.plane_ch[pl]) col[c] = src[k];
          }
        }
      }
      s_gid[threadIdx.x] = g;
      const float opac = p.opacities[g];
      s_box[threadIdx.x] = splat_reach_box(xy.x, xy.y, opac, ca, cb, cc);
      s_geo[threadIdx.x] = make_float4(xy.x, xy.y, opac, ca);
      s_rec[0][threadIdx.x] = make_float4(cb, cc, col[0], col[1]);
#pragma unroll
      for (int v = 1; v < NCOLV; ++v)
        s_rec[v][threadIdx.x] = make_float4(col[4 * v - 2], col[4 * v - 1], col[4 * v], col[4 * v + 1]);
    }
  };

  // in-order list of the staged splats (first n entries) whose reach box meets this warp's 8 x 4 pixel block
  auto build_list = [&](int n) {
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < kBwdThreads / 32; ++k) {
      const int t = k * 32 + lane;
      bool keep = false;
      if (t < n) {
        const float4 box = s_box[t];
        keep = !(box.x > blk_x_hi || box.y < blk_x_lo || box.z > blk_y_hi || box.w < blk_y_lo);
      }
      const unsigned m = __ballot_sync(FULL_MASK, keep);
      if (keep) s_list[warp][cnt + __popc(m & ((1u << lane) - 1u))] = (unsigned char)t;
      cnt += __popc(m);
    }
    __syncwarp();
    return cnt;
  };

  // ---- phase 1: final transmittance and end of the contributing range of every pixel ----
  float T = 1.0f;
  int last = lo;  // one past the last splat this pixel composited
  {
    bool done = !inside;
    for (int batch = lo; batch < hi; batch += kBwdThreads) {
      if (__syncthreads_count(done) >= kBwdThreads) break;
      stage(batch, hi);
      __syncthreads();
      const int n = min(kBwdThreads, hi - batch);
      if (__all_sync(FULL_MASK, done)) continue;
      const int cnt = build_list(n);
      for (int q = 0; q < cnt && !done; ++q) {
        const int t = s_list[warp][q];
        const float4 ga = s_geo[t];
        const float4 gb = s_rec[0][t];
        const float dx = ga.x - px, dy = ga.y - py;
        const float sigma = splat_sigma(ga.w, gb.x, gb.y, dx, dy);
        const float alpha = splat_alpha(ga.z, sigma);
        if (sigma < 0.0f || alpha < 1.0f / 255.0f) continue;
        const float next_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
        if (next_T <= 1e-4f) {
          done = true;
          break;
        }
        T = next_T;
        last = batch + t + 1;
      }
    }
  }
  const float T_final = T;

  // block-wide end of all contributing ranges
  if (threadIdx.x == 0) s_end = lo;
  __syncthreads();
  atomicMax(&s_end, inside ? last : lo);
  for (int k = threadIdx.x; k < kBwdThreads * (NV + 1); k += kBwdThreads) (&s_acc[0][0])[k] = 0.f;
  __syncthreads();
  const int end = s_end;
  if (end <= lo) return;

  // ---- per-pixel upstream gradients ----
  float vo[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) vo[c] = 0.f;
  float tail = 0.f;  // T_final (v_alpha - sum_c bg_c v_out_c)
  if (inside) {
    const size_t pix = (size_t)i * p.width + j;
#pragma unroll
    for (int pl = 0; pl < kBwdPlanes; ++pl) {
      if (pl < p.num_planes && p.v_out[pl]) {
        const float* src = p.v_out[pl] + pix * p.plane_ch[pl];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int k = c - p.plane_off[pl];
          if (k >= 0 && k < p.plane_ch[pl]) vo[c] = src[k];
        }
      }
    }
    float bg_dot = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) bg_dot += p.background[c] * vo[c];
    tail = T_final * ((p.v_alpha ? p.v_alpha[pix] : 0.f) - bg_dot);
  }

  // ---- phase 2: back to front ----
  float S[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) S[c] = 0.f;
  const int first_batch = lo + ((end - 1 - lo) / kBwdThreads) * kBwdThreads;
  for (int batch = first_batch; batch >= lo; batch -= kBwdThreads) {
    __syncthreads();  // previous batch flushed, staging buffers free
    stage(batch, end);
    __syncthreads();
    const int n = min(kBwdThreads, end - batch);
    // a warp none of whose pixels reached this batch has nothing to do in it
    const int cnt = __any_sync(FULL_MASK, inside && last > batch) ? build_list(n) : 0;
    for (int q = cnt - 1; q >= 0; --q) {
      const int t = s_list[warp][q];
      const int idx = batch + t;
      bool valid = inside && idx < last;
      float g[NV];
#pragma unroll
      for (int k = 0; k < NV; ++k) g[k] = 0.f;
      if (valid) {
        const float4 ga = s_geo[t];
        const float4 gb = s_rec[0][t];
        const float dx = ga.x - px, dy = ga.y - py;
        const float sigma = splat_sigma(ga.w, gb.x, gb.y, dx, dy);
        const float vis = splat_falloff(sigma);
        const float raw = __fmul_rn(ga.z, vis);
        const float alpha = fminf(0.999f, raw);
        valid = !(sigma < 0.0f || alpha < 1.0f / 255.0f);
        if (valid) {
          float col[4 * NCOLV - 2];
          col[0] = gb.z;
          col[1] = gb.w;
#pragma unroll
          for (int v = 1; v < NCOLV; ++v) {
            const float4 r = s_rec[v][t];
            col[4 * v - 2] = r.x;
            col[4 * v - 1] = r.y;
            col[4 * v] = r.z;
            col[4 * v + 1] = r.w;
          }
          const float ra = 1.0f / (1.0f - alpha);
          T *= ra;  // transmittance in front of this splat
          const float fac = alpha * T;
          float v_alpha = tail * ra;
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            g[6 + c] = fac * vo[c];
            v_alpha += (col[c] * T - S[c] * ra) * vo[c];
            S[c] += col[c] * fac;
          }
          if (raw <= 0.999f) {
            const float v_sigma = -raw * v_alpha;
            g[0] = v_sigma * (ga.w * dx + gb.x * dy);
            g[1] = v_sigma * (gb.x * dx + gb.y * dy);
            g[2] = 0.5f * v_sigma * dx * dx;
            g[3] = v_sigma * dx * dy;
            g[4] = 0.5f * v_sigma * dy * dy;
            g[5] = vis * v_alpha;
          }
        }
      }
      if (!__any_sync(FULL_MASK, valid)) continue;
      constexpr int P = NV <= 8 ? 8 : 16;
      float gp[P];
#pragma unroll
      for (int k = 0; k < P; ++k) gp[k] = k < NV ? g[k] : 0.f;
      int which;
      const float total = warp_sum_scatter<P>(gp, lane, which);
      // one lane per value index adds the warp total to the batch entry's accumulator
      const bool owner = P == 16 ? (lane & 1) == 0 : (lane & 17) == 0;
      if (owner && which < NV && total != 0.f) atomicAdd(&s_acc[t][which], total);
    }
    __syncthreads();
    // flush this batch: thread t owns entry t
    if ((int)threadIdx.x < n) {
      const int gid = s_gid[threadIdx.x];
      float a[NV];
      bool any = false;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        a[k] = s_acc[threadIdx.x][k];
        s_acc[threadIdx.x][k] = 0.f;
        any |= a[k] != 0.f;
      }
      if (any) {
        atomicAdd(&p.v_xys[2 * (size_t)gid + 0], a[0]);
        atomicAdd(&p.v_xys[2 * (size_t)gid + 1], a[1]);
        atomicAdd(&p.v_conics[3 * (size_t)gid + 0], a[2]);
        atomicAdd(&p.v_conics[3 * (size_t)gid + 1], a[3]);
        atomicAdd(&p.v_conics[3 * (size_t)gid + 2], a[4]);
        atomicAdd(&p.v_opacities[gid], a[5]);
#pragma unroll
        for (int pl = 0; pl < kBwdPlanes; ++pl) {
          if (pl < p.num_planes && p.v_plane[pl]) {
            float* dst = p.v_plane[pl] + (size_t)gid * p.plane_ch[pl];
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              const int k = c - p.plane_off[pl];
              if (k >= 0 && k < p.plane_ch[pl]) atomicAdd(&dst[k], a[6 + c]);
            }
          }
        }
      }
    }
  }
}

}  // namespace ub

extern "C" {

int ub_composite_tiles_planes_backward(const float* xys, const float* conics, const float* opacities,
                                       const float* const* planes_host, const int32_t* plane_channels_host,
                                       int32_t num_planes, const int32_t* gaussian_ids, const int32_t* tile_bins,
                                       int32_t img_height, int32_t img_width, const float* background_host,
                                       const float* const* v_outs_host, const float* v_alpha,
                                       int64_t num_gaussians, float* v_xys, float* v_conics, float* v_opacities,
                                       float* const* v_planes_host, void* stream_v) {
  using namespace ub;
  UB_REQUIRE(num_planes >= 1 && num_planes <= kBwdPlanes && planes_host && plane_channels_host && v_outs_host &&
                 v_planes_host,
             UB_ERR_BAD_ARG, "composite_tiles_backward: between 1 and %d colour planes", kBwdPlanes);
  UB_REQUIRE(img_height >= 1 && img_width >= 1 && num_gaussians >= 0, UB_ERR_BAD_ARG,
             "composite_tiles_backward: bad sizes");
  UB_REQUIRE(v_xys && v_conics && v_opacities, UB_ERR_BAD_ARG,
             "composite_tiles_backward: v_xys / v_conics / v_opacities must be non-NULL");
  UB_REQUIRE(tile_bins != nullptr, UB_ERR_BAD_ARG, "composite_tiles_backward: tile_bins is NULL");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  TileBwdParams p{};
  int total = 0;
  for (int pl = 0; pl < num_planes; ++pl) {
    const int ch = plane_channels_host[pl];
    UB_REQUIRE(ch >= 1 && planes_host[pl] != nullptr, UB_ERR_BAD_ARG, "composite_tiles_backward: bad plane %d", pl);
    p.plane[pl] = planes_host[pl];
    p.plane_ch[pl] = ch;
    p.plane_off[pl] = total;
    p.v_out[pl] = v_outs_host[pl];
    p.v_plane[pl] = v_planes_host[pl];
    total += ch;
  }
  UB_REQUIRE(total <= UB_MAX_SPLAT_CHANNELS, UB_ERR_UNSUPPORTED,
             "composite_tiles_backward: %d channels > %d", total, UB_MAX_SPLAT_CHANNELS);
  // zero the outputs (they are accumulated with atomics)
  const size_t g = (size_t)num_gaussians;
  bool ok = cudaMemsetAsync(v_xys, 0, g * 2 * sizeof(float), stream) == cudaSuccess &&
            cudaMemsetAsync(v_conics, 0, g * 3 * sizeof(float), stream) == cudaSuccess &&
            cudaMemsetAsync(v_opacities, 0, g * sizeof(float), stream) == cudaSuccess;
  for (int pl = 0; pl < num_planes && ok; ++pl)
    if (p.v_plane[pl])
      ok = cudaMemsetAsync(p.v_plane[pl], 0, g * p.plane_ch[pl] * sizeof(float), stream) == cudaSuccess;
  if (!ok) return check_launch("composite_tiles_backward memset");
  if (num_gaussians == 0 || gaussian_ids == nullptr) return UB_OK;
  UB_REQUIRE(xys && conics && opacities, UB_ERR_BAD_ARG, "composite_tiles_backward: geometry pointers are NULL");
  p.xys = xys;
  p.conics = conics;
  p.opacities = opacities;
  p.num_planes = num_planes;
  p.gaussian_ids = gaussian_ids;
  p.tile_bins = tile_bins;
  p.height = img_height;
  p.width = img_width;
  p.tiles_x = (img_width + UB_TILE - 1) / UB_TILE;
  for (int c = 0; c < UB_MAX_SPLAT_CHANNELS; ++c)
    p.background[c] = (background_host && c < total) ? background_host[c] : 0.f;
  p.v_alpha = v_alpha;
  p.v_xys = v_xys;
  p.v_conics = v_conics;
  p.v_opacities = v_opacities;
  dim3 grid((unsigned)p.tiles_x, (unsigned)((img_height + UB_TILE - 1) / UB_TILE));
  switch (total) {
    case 1: composite_tiles_bwd_kernel<1><<<grid, kBwdThreads, 0, stream>>>(p); break;
    case 2: composite_tiles_bwd_kernel<2><<<grid, kBwdThreads, 0, stream>>>(p); break;
    case 3: composite_tiles_bwd_kernel<3><<<grid, kBwdThreads, 0, stream>>>(p); break;
    case 4: composite_tiles_bwd_kernel<4><<<grid, kBwdThreads, 0, stream>>>(p); break;
    case 5: composite_tiles_bwd_kernel<5><<<grid, kBwdThreads, 0, stream>>>(p); break;
    case 6: composite_tiles_bwd_kernel<6><<<grid, kBwdThreads, 0, stream>>>(p); break;
    case 7: composite_tiles_bwd_kernel<7><<<grid, kBwdThreads, 0, stream>>>(p); break;
    default: composite_tiles_bwd_kernel<8><<<grid, kBwdThreads, 0, stream>>>(p); break;
  }
  return check_launch("composite_tiles_backward");
}

}  // extern "C"
