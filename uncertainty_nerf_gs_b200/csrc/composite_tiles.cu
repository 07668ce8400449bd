// (A2) Front-to-back alpha compositing over pre-binned per-tile splat lists (16x16 tiles) and the
// active-splatfacto post-processing.
//
// Reference call sites (file:line under /root/reference/nerfuncertainty):
//   models/activesplatfacto/activesplatfacto_model.py:260-275   rgb + alpha, rgb = clamp(rgb, max=1)
//   models/activesplatfacto/activesplatfacto_model.py:286-301   beta image (channel 0 of a x3 repeat)
//   models/activesplatfacto/activesplatfacto_model.py:306-319   depth image, / alpha where alpha > 0 else max
//   models/activesplatfacto/activesplatfacto_model.py:325-341   per-Gaussian (depth - depth_im[centre pixel])^2
//   models/activesplatfacto/activesplatfacto_model.py:343-356   depth-variance image, / alpha else max
// each rasterisation being one gsplat 0.1.11 `rasterize_gaussians` launch with 3-channel colours.  The
// per-pixel loop restates gsplat's published `rasterize_forward` (not vendored in the reference):
// pixel centre (j+0.5, i+0.5); sigma = 0.5(a dx^2 + c dy^2) + b dx dy; alpha = min(0.999, opac * __expf(-sigma))
// (the fast intrinsic, as in gsplat's kernel); skip if sigma < 0 or alpha < 1/255; stop *before* a Gaussian that
// would bring T to <= 1e-4; out += colour * alpha * T; T *= 1 - alpha.  The rounding of sigma / alpha / T is pinned
// (ub_common.cuh: splat_sigma, splat_alpha), and `ub_tile_alpha_probe` hands the (sigma, alpha) values of any tile to
// the host, so that the CPU oracle replays every threshold decision in the same float32 arithmetic: the set of
// contributing splats of every pixel -- hence the alpha image -- is reproduced bit for bit (tests/test_gpu_splat_exact.py).
//
// Here all channels that share the geometry (rgb, beta, depth = 5) go through ONE pass.  A CTA owns a
// tile and stages 256 splats at a time in shared memory, packed as float4 records
// {x, y, opacity, conic.a} {conic.b, conic.c, c0, c1} {c2, c3, c4, c5} ...: the per-(pixel, splat) inner loop
// is bound by shared-memory load issue, so it costs two LDS.128 for the geometry test and one more for
// the colours of a splat that contributes (instead of eleven scalar loads).  Colours come from up to four
// "planes" (rgb [G,3], beta [G,1], depth [G,1]) and go to one image per plane, so no concatenated colour
// tensor and no strided channel views exist on the host side.  The per-channel maxima the reference's
// post-processing needs (`depth_im.max()`) are accumulated with one atomicMax per tile.
#include <cstdlib>

#include "ub_common.cuh"

namespace ub {

constexpr int kTileThreads = UB_TILE * UB_TILE;
constexpr int kMaxPlanes = UB_MAX_SPLAT_PLANES;

struct TileParams {
  const float* xys;
  const float* conics;
  const float* opacities;
  const float* plane[kMaxPlanes];
  int plane_ch[kMaxPlanes];
  int plane_off[kMaxPlanes];  // first channel of the plane
  int num_planes;
  const int32_t* gaussian_ids;
  const int32_t* tile_bins;
  int height, width, tiles_x;
  float background[UB_MAX_SPLAT_CHANNELS];
  float* out[kMaxPlanes];
  float* out_alpha;
  unsigned* channel_max_keys;  // [channels] order keys, or NULL
  int no_cull;                 // debugging aid (UB_TILES_NO_CULL=1): evaluate every (pixel, splat) pair
};

template <int CH>
__global__ void __launch_bounds__(kTileThreads) composite_tiles_kernel(const TileParams p) {
  constexpr int NCOLV = (CH + 2 + 3) / 4;  // float4 records after the first: {cb, cc, c0, c1}, {c2..c5}, ...
  // one record per staged splat, {x, y, opacity, conic.a} {conic.b, conic.c, c0, c1} {c2..c5} ...: the per-pixel loop
  // forms one address per splat and reads its float4s at constant offsets
  __shared__ float4 s_gr[kTileThreads][1 + NCOLV];
  __shared__ float4 s_box[kTileThreads];
  __shared__ unsigned char s_list[kTileThreads / 32][kTileThreads];  // per warp: staged splats that can touch its block
  __shared__ float s_max[kTileThreads / 32][CH];

  const int tile = blockIdx.y * p.tiles_x + blockIdx.x;
  int ti, tj;  // row / column inside the tile
  tile_pixel_of_thread(threadIdx.x, ti, tj);
  const int i = blockIdx.y * UB_TILE + ti, j = blockIdx.x * UB_TILE + tj;
  const float px = (float)j + 0.5f, py = (float)i + 0.5f;
  const bool inside = i < p.height && j < p.width;
  bool done = !inside;
  const int lo = p.tile_bins[2 * tile + 0], hi = p.tile_bins[2 * tile + 1];
  // pixel-centre rows of this warp (two rows of the tile) and columns of the tile
  // pixel-centre box of this warp's 8 x 4 block
  const float blk_x_lo = (float)(blockIdx.x * UB_TILE + 8 * ((threadIdx.x >> 5) & 1)) + 0.5f, blk_x_hi = blk_x_lo + 7.0f;
  const float blk_y_lo = (float)(blockIdx.y * UB_TILE + 4 * (threadIdx.x >> 6)) + 0.5f, blk_y_hi = blk_y_lo + 3.0f;

  float T = 1.0f;
  float acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = 0.0f;

  for (int batch = lo; batch < hi; batch += kTileThreads) {
    if (__syncthreads_count(done) >= kTileThreads) break;
    const int idx = batch + threadIdx.x;
    if (idx < hi) {
      const int g = p.gaussian_ids[idx];
      const float2 xy = *reinterpret_cast<const float2*>(p.xys + 2 * (size_t)g);
      const float ca = p.conics[3 * (size_t)g + 0], cb = p.conics[3 * (size_t)g + 1], cc = p.conics[3 * (size_t)g + 2];
      float col[4 * NCOLV - 2];
#pragma unroll
      for (int c = 0; c < 4 * NCOLV - 2; ++c) col[c] = 0.f;
#pragma unroll
      for (int pl = 0; pl < kMaxPlanes; ++pl) {
        if (pl < p.num_planes) {
          const float* src = p.plane[pl] + (size_t)g * p.plane_ch[pl];
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const int k = c - p.plane_off[pl];
            if (k >= 0 && k < p.plane_ch[pl]) col[c] = src[k];
          }
        }
      }
      const float opac = p.opacities[g];
      s_gr[threadIdx.x][0] = make_float4(xy.x, xy.y, opac, ca);
      s_gr[threadIdx.x][1] = make_float4(cb, cc, col[0], col[1]);
#pragma unroll
      for (int v = 1; v < NCOLV; ++v)
        s_gr[threadIdx.x][1 + v] = make_float4(col[4 * v - 2], col[4 * v - 1], col[4 * v], col[4 * v + 1]);
      s_box[threadIdx.x] = p.no_cull ? make_float4(-INFINITY, INFINITY, -INFINITY, INFINITY)
                                     : splat_reach_box(xy.x, xy.y, opac, ca, cb, cc);
    }
    __syncthreads();
    const int n = min(kTileThreads, hi - batch);
    // warp-level cull: compact, in order, the staged splats whose reach box meets this warp's 8 x 4 block; the
    // per-pixel loop below then never sees the others (about half of a tile's list for a typical scene)
    if (__all_sync(FULL_MASK, done)) continue;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < kTileThreads / 32; ++k) {
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
    // (the list is walked by its shared-memory address: one induction variable instead of a counter and a pointer)
    for (uint32_t lp = smem_u32(s_list[warp]), le = lp + (uint32_t)cnt; lp < le && !done; ++lp) {
      uint32_t t;
      asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t) : "r"(lp));
      const float4 ga = s_gr[t][0];
      const float4 gb = s_gr[t][1];
      const float dx = ga.x - px, dy = ga.y - py;
      const float sigma = splat_sigma(ga.w, gb.x, gb.y, dx, dy);
      const float alpha = splat_alpha(ga.z, sigma);
      if (sigma < 0.0f || alpha < 1.0f / 255.0f) continue;
      const float next_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
      if (next_T <= 1e-4f) {
        done = true;
        break;
      }
      const float vis = __fmul_rn(alpha, T);
      acc[0] += gb.z * vis;
      if (CH > 1) acc[1] += gb.w * vis;
#pragma unroll
      for (int v = 1; v < NCOLV; ++v) {
        const float4 r = s_gr[t][1 + v];
        if (4 * v - 2 < CH) acc[4 * v - 2] += r.x * vis;
        if (4 * v - 1 < CH) acc[4 * v - 1] += r.y * vis;
        if (4 * v + 0 < CH) acc[4 * v + 0] += r.z * vis;
        if (4 * v + 1 < CH) acc[4 * v + 1] += r.w * vis;
      }
      T = next_T;
    }
  }
  float val[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) val[c] = acc[c] + T * p.background[c];
  if (inside) {
    const size_t pix = (size_t)i * p.width + j;
#pragma unroll
    for (int pl = 0; pl < kMaxPlanes; ++pl) {
      if (pl < p.num_planes && p.out[pl]) {
        float* dst = p.out[pl] + pix * p.plane_ch[pl];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int k = c - p.plane_off[pl];
          if (k >= 0 && k < p.plane_ch[pl]) dst[k] = val[c];
        }
      }
    }
    if (p.out_alpha) p.out_alpha[pix] = 1.0f - T;
  }
  if (p.channel_max_keys) {  // per-channel maximum over the image: warp shuffle, shared, one atomic per tile
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      float m = inside ? val[c] : -INFINITY;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, o));
      if (lane == 0) s_max[warp][c] = m;
    }
    __syncthreads();
    if (threadIdx.x < CH) {
      float m = -INFINITY;
      for (int w = 0; w < kTileThreads / 32; ++w) m = fmaxf(m, s_max[w][threadIdx.x]);
      if (m > -INFINITY) atomicMax(&p.channel_max_keys[threadIdx.x], order_key(m));
    }
  }
}

// Diagnostic: (sigma, alpha) of `count` consecutive entries of a tile's list for the 256 pixels of the tile, exactly as
// composite_tiles_kernel evaluates them.  One block per list entry, thread = pixel (row-major inside the tile).
__global__ void __launch_bounds__(kTileThreads)
tile_alpha_probe_kernel(const float* xys, const float* conics, const float* opacities, const int32_t* gaussian_ids,
                        int first, int tile_x, int tile_y, float* out_sigma, float* out_alpha) {
  const int g = gaussian_ids[first + blockIdx.x];
  const int ti = threadIdx.x / UB_TILE, tj = threadIdx.x % UB_TILE;
  const float px = (float)(tile_x * UB_TILE + tj) + 0.5f, py = (float)(tile_y * UB_TILE + ti) + 0.5f;
  const float dx = xys[2 * (size_t)g] - px, dy = xys[2 * (size_t)g + 1] - py;
  const float sigma = splat_sigma(conics[3 * (size_t)g], conics[3 * (size_t)g + 1], conics[3 * (size_t)g + 2], dx, dy);
  out_sigma[(size_t)blockIdx.x * kTileThreads + threadIdx.x] = sigma;
  out_alpha[(size_t)blockIdx.x * kTileThreads + threadIdx.x] = splat_alpha(opacities[g], sigma);
}

static int launch_tiles(TileParams p, int channels, int img_height, cudaStream_t stream) {
  static const bool no_cull = [] { const char* e = getenv("UB_TILES_NO_CULL"); return e && e[0] == '1'; }();
  p.no_cull = no_cull ? 1 : 0;
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

// img[..., channel] = alpha > 0 ? img / alpha : max (reference :319, :356); rgb = min(rgb, 1) (:275)
__global__ void __launch_bounds__(256)
splat_normalize_kernel(float* img, int ch, const float* alpha, long long num_pixels, int clamp_max_one,
                       int divide_by_alpha, const unsigned* max_key, float* out_square, float* out_sqrt) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= num_pixels) return;
  if (clamp_max_one) {
    for (int c = 0; c < ch; ++c) {
      const float v = img[pix * ch + c];
      img[pix * ch + c] = v != v ? v : fminf(v, 1.0f);  // torch.clamp keeps NaN
    }
  }
  if (divide_by_alpha) {
    const float a = alpha[pix];
    const float mx = order_key_inv(*max_key);
    for (int c = 0; c < ch; ++c) img[pix * ch + c] = a > 0.0f ? img[pix * ch + c] / a : mx;
  }
  if (out_square || out_sqrt) {  // rgb_var = uncertainty ** 2 (:364), depth_std = depth_var.sqrt() (:367)
    for (int c = 0; c < ch; ++c) {
      const float v = img[pix * ch + c];
      if (out_square) out_square[pix * ch + c] = __fmul_rn(v, v);
      if (out_sqrt) out_sqrt[pix * ch + c] = sqrtf(v);
    }
  }
}

// out[g] = (depth_g - depth_im[floor(y), floor(x)])^2 for Gaussians whose centre pixel satisfies 0 < x < W,
// 0 < y < H (strict, as the reference's mask), else depth_g^2 (:325-341, :349)
__global__ void __launch_bounds__(256)
splat_depth_residual_kernel(const float* xys, const float* depths, const float* depth_img, int height, int width,
                            long long num_gaussians, float* out) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= num_gaussians) return;
  const long long x = (long long)floorf(xys[2 * g + 0]), y = (long long)floorf(xys[2 * g + 1]);
  float d = depths[g];
  if (x > 0 && x < width && y > 0 && y < height) d -= depth_img[y * width + x];
  out[g] = d * d;
}

}  // namespace ub

extern "C" {

int ub_composite_tiles_planes(const float* xys, const float* conics, const float* opacities,
                              const float* const* planes_host, const int32_t* plane_channels_host,
                              int32_t num_planes, const int32_t* gaussian_ids, const int32_t* tile_bins,
                              int32_t img_height, int32_t img_width, const float* background_host,
                              float* const* outs_host, float* out_alpha, uint32_t* channel_max_keys,
                              void* stream_v) {
  using namespace ub;
  UB_REQUIRE(num_planes >= 1 && num_planes <= kMaxPlanes && planes_host && plane_channels_host && outs_host,
             UB_ERR_BAD_ARG, "composite_tiles: between 1 and %d colour planes", kMaxPlanes);
  UB_REQUIRE(img_height >= 1 && img_width >= 1, UB_ERR_BAD_ARG, "composite_tiles: bad image size");
  // gaussian_ids may be NULL when there is no intersection at all (every tile range empty)
  UB_REQUIRE(xys && conics && opacities && tile_bins, UB_ERR_BAD_ARG, "composite_tiles: NULL input pointer");
  TileParams p{};
  int channels = 0;
  for (int pl = 0; pl < num_planes; ++pl) {
    UB_REQUIRE(planes_host[pl] != nullptr && plane_channels_host[pl] >= 1, UB_ERR_BAD_ARG,
               "composite_tiles: plane %d is NULL or has no channels", pl);
    p.plane[pl] = planes_host[pl];
    p.plane_ch[pl] = plane_channels_host[pl];
    p.plane_off[pl] = channels;
    p.out[pl] = outs_host[pl];
    channels += plane_channels_host[pl];
  }
  UB_REQUIRE(channels <= UB_MAX_SPLAT_CHANNELS, UB_ERR_UNSUPPORTED, "composite_tiles: at most %d channels in total",
             UB_MAX_SPLAT_CHANNELS);
  p.num_planes = num_planes;
  p.xys = xys;
  p.conics = conics;
  p.opacities = opacities;
  p.gaussian_ids = gaussian_ids;
  p.tile_bins = tile_bins;
  p.height = img_height;
  p.width = img_width;
  p.tiles_x = (img_width + UB_TILE - 1) / UB_TILE;
  for (int c = 0; c < UB_MAX_SPLAT_CHANNELS; ++c)
    p.background[c] = (background_host && c < channels) ? background_host[c] : 0.0f;
  p.out_alpha = out_alpha;
  p.channel_max_keys = channel_max_keys;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (channel_max_keys &&
      cudaMemsetAsync(channel_max_keys, 0, (size_t)channels * sizeof(uint32_t), stream) != cudaSuccess)
    return check_launch("composite_tiles max memset");
  return launch_tiles(p, channels, img_height, stream);
}

int ub_composite_tiles(const float* xys, const float* conics, const float* opacities, const float* colors,
                       int32_t channels, const int32_t* gaussian_ids, const int32_t* tile_bins,
                       int32_t img_height, int32_t img_width, const float* background_host, float* out,
                       float* out_alpha, void* stream_v) {
  using namespace ub;
  UB_REQUIRE(channels >= 1 && channels <= UB_MAX_SPLAT_CHANNELS, UB_ERR_UNSUPPORTED,
             "composite_tiles: channels must be in [1, %d]", UB_MAX_SPLAT_CHANNELS);
  UB_REQUIRE(colors && out, UB_ERR_BAD_ARG, "composite_tiles: NULL input / output pointer");
  const float* planes[1] = {colors};
  const int32_t chs[1] = {channels};
  float* outs[1] = {out};
  return ub_composite_tiles_planes(xys, conics, opacities, planes, chs, 1, gaussian_ids, tile_bins, img_height,
                                   img_width, background_host, outs, out_alpha, nullptr, stream_v);
}

int ub_tile_alpha_probe(const float* xys, const float* conics, const float* opacities, const int32_t* gaussian_ids,
                        int32_t first, int32_t count, int32_t tile_x, int32_t tile_y, float* out_sigma,
                        float* out_alpha, void* stream_v) {
  using namespace ub;
  UB_REQUIRE(xys && conics && opacities && gaussian_ids && out_sigma && out_alpha, UB_ERR_BAD_ARG,
             "tile_alpha_probe: NULL pointer");
  UB_REQUIRE(first >= 0 && count >= 0 && tile_x >= 0 && tile_y >= 0, UB_ERR_BAD_ARG, "tile_alpha_probe: bad range");
  if (count == 0) return UB_OK;
  tile_alpha_probe_kernel<<<(unsigned)count, kTileThreads, 0, static_cast<cudaStream_t>(stream_v)>>>(
      xys, conics, opacities, gaussian_ids, first, tile_x, tile_y, out_sigma, out_alpha);
  return check_launch("tile_alpha_probe");
}

int ub_splat_normalize(float* image, int32_t channels, const float* alpha, int64_t num_pixels,
                       int32_t clamp_max_one, int32_t divide_by_alpha, const uint32_t* max_key, float* out_square,
                       float* out_sqrt, void* stream_v) {
  using namespace ub;
  UB_REQUIRE(image != nullptr && channels >= 1 && num_pixels >= 0, UB_ERR_BAD_ARG, "splat_normalize: bad arguments");
  UB_REQUIRE(!divide_by_alpha || (alpha != nullptr && max_key != nullptr), UB_ERR_BAD_ARG,
             "splat_normalize: divide_by_alpha needs alpha and the channel maximum");
  if (num_pixels == 0) return UB_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  splat_normalize_kernel<<<(unsigned)((num_pixels + 255) / 256), 256, 0, stream>>>(
      image, channels, alpha, num_pixels, clamp_max_one, divide_by_alpha, max_key, out_square, out_sqrt);
  return check_launch("splat_normalize");
}

int ub_splat_depth_residual(const float* xys, const float* depths, const float* depth_image, int32_t img_height,
                            int32_t img_width, int64_t num_gaussians, float* out_sq_residual, void* stream_v) {
  using namespace ub;
  UB_REQUIRE(num_gaussians >= 0 && img_height >= 1 && img_width >= 1, UB_ERR_BAD_ARG, "splat_depth_residual: bad sizes");
  if (num_gaussians == 0) return UB_OK;
  UB_REQUIRE(xys && depths && depth_image && out_sq_residual, UB_ERR_BAD_ARG, "splat_depth_residual: NULL pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  splat_depth_residual_kernel<<<(unsigned)((num_gaussians + 255) / 256), 256, 0, stream>>>(
      xys, depths, depth_image, img_height, img_width, num_gaussians, out_sq_residual);
  return check_launch("splat_depth_residual");
}

}  // extern "C"
