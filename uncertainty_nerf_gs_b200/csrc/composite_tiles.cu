// (A2) Front-to-back alpha compositing over pre-binned per-tile splat lists (16x16 tiles).
//
// Reference call sites (file:line under /root/reference/nerfuncertainty):
//   models/activesplatfacto/activesplatfacto_model.py:260-273   rgb + alpha
//   models/activesplatfacto/activesplatfacto_model.py:286-301   beta image (channel 0 of a x3 repeat)
//   models/activesplatfacto/activesplatfacto_model.py:306-318   depth image
//   models/activesplatfacto/activesplatfacto_model.py:343-355   depth-variance image
// each of which is one gsplat 0.1.11 `rasterize_gaussians` launch with 3-channel colours.  The
// per-pixel loop below restates gsplat's published `rasterize_forward` (not vendored in the
// reference: parity unpinned): pixel centre (j+0.5, i+0.5); sigma = 0.5(a dx^2 + c dy^2) + b dx dy;
// alpha = min(0.999, opac * exp(-sigma)); skip if sigma < 0 or alpha < 1/255; stop *before* a
// Gaussian that would bring T to <= 1e-4; out += colour * alpha * T.
//
// Here all channels that share the geometry (rgb, beta, depth = 5) go through ONE pass: a CTA owns
// a tile, stages 256 splats at a time (geometry + colours) in shared memory, and every pixel thread
// walks the staged list, so each intersection is read from HBM/L2 once per tile instead of once
// per tile per launch.
#include "ub_common.cuh"

namespace ub {

constexpr int kTileThreads = UB_TILE * UB_TILE;

struct TileParams {
  const float* xys;
  const float* conics;
  const float* opacities;
  const float* colors;
  const int32_t* gaussian_ids;
  const int32_t* tile_bins;
  int height, width, tiles_x;
  float background[UB_MAX_SPLAT_CHANNELS];
  float* out;
  float* out_alpha;
};

template <int CH>
__global__ void __launch_bounds__(kTileThreads) composite_tiles_kernel(const TileParams p) {
  __shared__ float s_x[kTileThreads], s_y[kTileThreads], s_op[kTileThreads];
  __shared__ float s_ca[kTileThreads], s_cb[kTileThreads], s_cc[kTileThreads];
  __shared__ float s_col[kTileThreads][CH];

  const int tile = blockIdx.y * p.tiles_x + blockIdx.x;
  const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;  // row / column inside the tile
  const int i = blockIdx.y * UB_TILE + ti, j = blockIdx.x * UB_TILE + tj;
  const float px = (float)j + 0.5f, py = (float)i + 0.5f;
  const bool inside = i < p.height && j < p.width;
  bool done = !inside;
  const int lo = p.tile_bins[2 * tile + 0], hi = p.tile_bins[2 * tile + 1];

  float T = 1.0f;
  float acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = 0.0f;

  for (int batch = lo; batch < hi; batch += kTileThreads) {
    if (__syncthreads_count(done) >= kTileThreads) break;
    const int idx = batch + threadIdx.x;
    if (idx < hi) {
      const int g = p.gaussian_ids[idx];
      s_x[threadIdx.x] = p.xys[2 * g + 0];
      s_y[threadIdx.x] = p.xys[2 * g + 1];
      s_op[threadIdx.x] = p.opacities[g];
      s_ca[threadIdx.x] = p.conics[3 * g + 0];
      s_cb[threadIdx.x] = p.conics[3 * g + 1];
      s_cc[threadIdx.x] = p.conics[3 * g + 2];
#pragma unroll
      for (int c = 0; c < CH; ++c) s_col[threadIdx.x][c] = p.colors[(size_t)g * CH + c];
    }
    __syncthreads();
    const int n = min(kTileThreads, hi - batch);
    for (int t = 0; t < n && !done; ++t) {
      const float dx = s_x[t] - px, dy = s_y[t] - py;
      const float sigma = 0.5f * (s_ca[t] * dx * dx + s_cc[t] * dy * dy) + s_cb[t] * dx * dy;
      const float alpha = fminf(0.999f, s_op[t] * expf(-sigma));
      if (sigma < 0.0f || alpha < 1.0f / 255.0f) continue;
      const float next_T = T * (1.0f - alpha);
      if (next_T <= 1e-4f) {
        done = true;
        break;
      }
      const float vis = alpha * T;
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] += s_col[t][c] * vis;
      T = next_T;
    }
  }
  if (inside) {
    const size_t pix = (size_t)i * p.width + j;
#pragma unroll
    for (int c = 0; c < CH; ++c) p.out[pix * CH + c] = acc[c] + T * p.background[c];
    if (p.out_alpha) p.out_alpha[pix] = 1.0f - T;
  }
}

}  // namespace ub

extern "C" int ub_composite_tiles(const float* xys, const float* conics, const float* opacities,
                                  const float* colors, int32_t channels, const int32_t* gaussian_ids,
                                  const int32_t* tile_bins, int32_t img_height, int32_t img_width,
                                  const float* background_host, float* out, float* out_alpha,
                                  void* stream_v) {
  using namespace ub;
  UB_REQUIRE(channels >= 1 && channels <= UB_MAX_SPLAT_CHANNELS, UB_ERR_UNSUPPORTED,
             "composite_tiles: channels must be in [1, %d]", UB_MAX_SPLAT_CHANNELS);
  UB_REQUIRE(img_height >= 1 && img_width >= 1, UB_ERR_BAD_ARG, "composite_tiles: bad image size");
  // gaussian_ids may be NULL when there is no intersection at all (every tile range empty)
  UB_REQUIRE(xys && conics && opacities && colors && tile_bins && out, UB_ERR_BAD_ARG,
             "composite_tiles: NULL input / output pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  TileParams p{};
  p.xys = xys;
  p.conics = conics;
  p.opacities = opacities;
  p.colors = colors;
  p.gaussian_ids = gaussian_ids;
  p.tile_bins = tile_bins;
  p.height = img_height;
  p.width = img_width;
  p.tiles_x = (img_width + UB_TILE - 1) / UB_TILE;
  for (int c = 0; c < UB_MAX_SPLAT_CHANNELS; ++c)
    p.background[c] = (background_host && c < channels) ? background_host[c] : 0.0f;
  p.out = out;
  p.out_alpha = out_alpha;
  dim3 grid((unsigned)p.tiles_x, (unsigned)((img_height + UB_TILE - 1) / UB_TILE));
  switch (channels) {
    case 1: composite_tiles_kernel<1><<<grid, kTileThreads, 0, stream>>>(p); break;
    case 2: composite_tiles_kernel<2><<<grid, kTileThreads, 0, stream>>>(p); break;
    case 3: composite_tiles_kernel<3><<<grid, kTileThreads, 0, stream>>>(p); break;
    case 4: composite_tiles_kernel<4><<<grid, kTileThreads, 0, stream>>>(p); break;
    case 5: composite_tiles_kernel<5><<<grid, kTileThreads, 0, stream>>>(p); break;
    case 6: composite_tiles_kernel<6><<<grid, kTileThreads, 0, stream>>>(p); break;
    case 7: composite_tiles_kernel<7><<<grid, kTileThreads, 0, stream>>>(p); break;
    default: composite_tiles_kernel<8><<<grid, kTileThreads, 0, stream>>>(p); break;
  }
  return check_launch("composite_tiles");
}
