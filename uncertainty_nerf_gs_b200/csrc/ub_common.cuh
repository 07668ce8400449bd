// Shared device / host helpers for the ub200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <float.h>

#include "../../include/ub200.h"

namespace ub {

// ---- host side: thread-local error record -------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaPeekAtLastError -> UB_ERR_LAUNCH
int sm_count();

#define UB_REQUIRE(cond, code, ...)      \
  do {                                   \
    if (!(cond)) {                       \
      ::ub::set_error(__VA_ARGS__);      \
      return (code);                     \
    }                                    \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- device side ---------------------------------------------------------------------------
constexpr unsigned FULL_MASK = 0xffffffffu;

// torch.nan_to_num with default arguments on float32: NaN -> 0, +-inf -> +-FLT_MAX.
__device__ __forceinline__ float nan_to_num(float x) {
  if (x != x) return 0.0f;
  if (x == INFINITY) return FLT_MAX;
  if (x == -INFINITY) return -FLT_MAX;
  return x;
}

// Order-preserving float32 -> uint32 key matching torch.sort on float32:
// -0.0 == +0.0, every NaN is the single largest key (ties broken by index => stable).
__device__ __forceinline__ uint32_t sort_key_from_float(float f) {
  if (f != f) return 0xFFFFFFFEu;
  if (f == 0.0f) f = 0.0f;  // canonicalise -0.0
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ uint32_t order_key(float f) {  // monotone, no canonicalisation
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float order_key_inv(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
  return __uint_as_float(b);
}

__device__ __forceinline__ double shfl_double(unsigned mask, double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(mask, lo, src);
  hi = __shfl_sync(mask, hi, src);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_xor_double(unsigned mask, double v, int lane_mask) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_xor_sync(mask, lo, lane_mask);
  hi = __shfl_xor_sync(mask, hi, lane_mask);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_up_double(unsigned mask, double v, int delta) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_up_sync(mask, lo, delta);
  hi = __shfl_up_sync(mask, hi, delta);
  return __hiloint2double(hi, lo);
}

// ---- mbarrier / bulk-copy (TMA) PTX wrappers -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
#ifdef UB_DEBUG_WAIT
__device__ unsigned long long g_ub_wait_debug[8];
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {
#ifdef UB_DEBUG_WAIT
  for (long long spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    if (spin > (1LL << 22)) {
      if (atomicAdd(&g_ub_wait_debug[0], 1ULL) == 0) {
        g_ub_wait_debug[1] = blockIdx.x;
        g_ub_wait_debug[2] = threadIdx.x;
        g_ub_wait_debug[3] = (unsigned long long)tag;
        g_ub_wait_debug[4] = parity;
        g_ub_wait_debug[5] = smem_u32(bar);
        printf("ub wait timeout: block %d thread %d tag %d parity %u bar %u\n", blockIdx.x, threadIdx.x,
               tag, parity, smem_u32(bar));
      }
      __trap();
    }
  }
#else
  (void)tag;
  while (!mbar_try_wait(bar, parity)) {
  }
#endif
}
// 1-D bulk asynchronous copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst / src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// Pixel-centre box {x_lo, x_hi, y_lo, y_hi} outside which a splat cannot reach alpha >= 1/255, for warp-level
// culling in the tile kernels: a warp covers an 8 x 4 pixel block of the tile, and a splat whose box misses the
// block cannot change any of its pixels, so the warp skips it with one uniform test instead of 32 per-pixel
// evaluations.  alpha >= 1/255 needs sigma <= L = ln(255 o); the ellipse {0.5 d^T Q d <= L} (Q = conic) spans
// |dx| <= sqrt(2 L C / det Q), |dy| <= sqrt(2 L A / det Q).  Slack (1e-3 on L, 1e-4 relative + 1e-3 px on the
// extents) keeps the test conservative under float32 rounding; a conic that is not positive definite or a NaN
// disables the cull; o < 1/255 makes the box empty.  Results are unchanged bit for bit.
__device__ __forceinline__ float4 splat_reach_box(float x, float y, float opac, float ca, float cb, float cc) {
  const float4 everywhere = make_float4(-INFINITY, INFINITY, -INFINITY, INFINITY);
  const float det = ca * cc - cb * cb;
  const float L = logf(255.0f * opac) + 1e-3f;
  if (!(det > 0.0f) || !(ca > 0.0f) || !(cc > 0.0f) || !(L == L) || !(x == x) || !(y == y)) return everywhere;
  if (L <= 0.0f) return make_float4(INFINITY, -INFINITY, INFINITY, -INFINITY);
  const float k = 2.0f * L / det;
  const float hy = sqrtf(k * ca) * (1.0f + 1e-4f) + 1e-3f;
  const float hx = sqrtf(k * cc) * (1.0f + 1e-4f) + 1e-3f;
  if (!(hy == hy) || !(hx == hx)) return everywhere;
  return make_float4(x - hx, x + hx, y - hy, y + hy);
}

// The per-(pixel, splat) test of gsplat 0.1.11's `rasterize_forward` with the rounding of every operation pinned,
// so that the forward kernel, its backward and the diagnostic probe (ub_tile_alpha_probe) agree bit for bit:
//   sigma = 0.5 (A dx^2 + C dy^2) + B dx dy        as  fma(B dx, dy, 0.5 * fma(C dy, dy, (A dx) dx))
//   alpha = min(0.999, opacity * __expf(-sigma))   __expf = ex2.approx(x log2 e), the intrinsic gsplat's kernel uses
// Decisions downstream (sigma < 0, alpha < 1/255, T (1 - alpha) <= 1e-4) are plain IEEE comparisons of these values.
__device__ __forceinline__ float splat_sigma(float ca, float cb, float cc, float dx, float dy) {
  const float a = __fmul_rn(__fmul_rn(ca, dx), dx);
  const float c = __fmaf_rn(__fmul_rn(cc, dy), dy, a);
  return __fmaf_rn(__fmul_rn(cb, dx), dy, __fmul_rn(0.5f, c));
}
// __expf(-sigma) the way gsplat's kernel gets it (built with --use_fast_math): ex2.approx.ftz of the float32 product with
// log2(e).  Without that flag nvcc wraps the same two instructions in a rescaling path for results below 2^-126
// (five instructions), which alpha >= 1/255 can never need; in the normal range both give the same bits.
__device__ __forceinline__ float splat_falloff(float sigma) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fmul_rn(-sigma, 1.4426950408889634f)));
  return r;
}
__device__ __forceinline__ float splat_alpha(float opac, float sigma) {
  return fminf(0.999f, __fmul_rn(opac, splat_falloff(sigma)));
}

// thread -> pixel of a 16 x 16 tile: warp w owns the 8 x 4 block at column 8 (w & 1), row 4 (w >> 1)
__device__ __forceinline__ void tile_pixel_of_thread(int tid, int& ti, int& tj) {
  const int w = tid >> 5, l = tid & 31;
  tj = 8 * (w & 1) + (l & 7);
  ti = 4 * (w >> 1) + (l >> 3);
}

}  // namespace ub
