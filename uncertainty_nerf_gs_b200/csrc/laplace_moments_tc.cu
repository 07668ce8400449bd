// (A3, tensor-core path) Last-layer diagonal-Laplace MC moments on tcgen05 with a 3xTF32 split.
//
// Reference arithmetic: models/laplace/laplace_field.py:545-565 (see laplace_moments.cu).  For the rgb head
// this is a [P,64] x [64,300] GEMM followed by sigmoid and a reduction over the 100 parameter draws --
// the one GEMM-shaped, compute-bound step of the path (38.4 kflop per point against 256 B).  The density head
// (one output, exp) is the same GEMM with the draws alone as its N dimension ([P,64] x [64,100]): template O = 1,
// only the first accumulator half is computed.
//
// One CTA per SM, 512 threads, a tile = 128 points:
//   * A (features) and B (all sampled weight rows, 3 per draw) are split x = hi + lo with hi = x truncated
//     to TF32 (top 19 bits) and lo = x - hi (exact); both parts live in shared memory in the canonical
//     K-major no-swizzle UMMA layout (8-row x 16-byte core matrices; LBO = 128 B between the K chunks of a
//     row group, SBO = 2048 B between 8-row groups).  B is converted once per CTA; the features of the next
//     tile are prefetched into registers while the current tile is in the tensor core.
//   * one elected thread issues tcgen05.mma.kind::tf32 (M = 128, N = 160 then 144, K = 8 per instruction):
//     hi*hi + lo*hi + hi*lo into float32 accumulators in tensor memory (the dropped lo*lo term is ~2^-22
//     relative); each accumulator half is committed to its own mbarrier, so the epilogue of the first half
//     overlaps the MMAs of the second;
//   * epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31 (one point per thread) and the 16-column chunks
//     c = w / 4 (mod 4): tcgen05.ld, bias and -log2(e) folded into one FFMA, ex2 + rcp on the MUFU pipe
//     (which bounds this phase), partial E[y], E[y^2] per channel; the four column groups of a point are
//     combined through shared memory (the A tile is dead by then).
#include "ub_common.cuh"

namespace ub {

constexpr int kTcH = 64;                 // features
constexpr int kTcM = 128;                // points per tile
constexpr int kTcThreads = 512;          // 16 warps: 4 TMEM lane quarters x 4 column groups
constexpr int kTcGroups = kTcThreads / kTcM;
constexpr int kTcChunksPerThread = kTcH / 4 / kTcGroups;  // 16-byte feature chunks a thread stages per tile
constexpr int kTcN0 = 160, kTcN1 = 144;  // accumulator halves (multiples of 16), 304 >= 300 columns
constexpr int kTcN = kTcN0 + kTcN1;
constexpr int kTcRowGroupBytes = (kTcH / 4) * 128;  // 2048: 16 K-chunks x (8 rows x 16 B)
constexpr uint32_t kTcTmemCols = 512;
constexpr uint32_t kTcTmemCol1 = 256;    // column offset of the second accumulator half

struct TcSmem {
  float a_hi[kTcM * kTcH];
  float a_lo[kTcM * kTcH];
  float b_hi[kTcN * kTcH];
  float b_lo[kTcN * kTcH];
  float bias[kTcN];        // sigmoid: -bias * log2(e); otherwise the plain bias
  uint64_t mma_bar[2];
  uint32_t tmem_base;
};

// byte offset of the 16-byte chunk (row r, K-chunk j) in the canonical K-major no-swizzle layout
__device__ __forceinline__ uint32_t canon_off(int r, int j) {
  return (uint32_t)((r >> 3) * kTcRowGroupBytes + j * 128 + (r & 7) * 16);
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ void split_store(float4 v, unsigned char* hi_base, unsigned char* lo_base, uint32_t off) {
  float4 hi, lo;
  hi.x = tf32_hi(v.x); hi.y = tf32_hi(v.y); hi.z = tf32_hi(v.z); hi.w = tf32_hi(v.w);
  lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
  *reinterpret_cast<float4*>(hi_base + off) = hi;
  *reinterpret_cast<float4*>(lo_base + off) = lo;
}

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  // start address >> 4 | LBO (128 B) >> 4 at bit 16 | SBO (2048 B) >> 4 at bit 32 | version 1 at bit 46 |
  // base offset 0 | layout type 0 (no swizzle) at bit 61
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) |
         ((uint64_t)(kTcRowGroupBytes >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
  // c_format F32 (1) at bit 4, a/b format TF32 (2) at bits 7 / 10, K-major A and B, N >> 3 at bit 17, M >> 4 at bit 24
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// one 16-column chunk of this thread's point: activation and per-channel partial sums (by i % O)
template <int ACT, int O, bool FULL>
__device__ __forceinline__ void epilogue_chunk(const float* v, const float* bias, int valid, float* t, float* t2) {
  const float4* b4 = reinterpret_cast<const float4*>(bias);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 b = b4[q];
    const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = 4 * q + e;
      float y;
      if (ACT == UB_ACT_SIGMOID) {
        // 1 / (1 + exp(-(v + bias))) with bias' = -bias * log2(e) precomputed
        y = rcp_approx(1.0f + ex2_approx(fmaf(v[i], -1.4426950408889634f, bb[e])));
      } else if (ACT == UB_ACT_EXP) {
        y = expf(v[i] + bb[e]);
      } else {
        y = v[i] + bb[e];
      }
      if (!FULL && i >= valid) y = 0.f;
      t[i % O] += y;
      t2[i % O] = fmaf(y, y, t2[i % O]);
    }
  }
}

template <int ACT, int O>
__global__ void __launch_bounds__(kTcThreads, 1)
laplace_moments_tc_kernel(const float* __restrict__ x, long long num_points,
                          const float* __restrict__ params, int n_samples,
                          float* __restrict__ o_mean, float* __restrict__ o_mean2,
                          float* __restrict__ o_sigma2) {
  extern __shared__ __align__(128) unsigned char tc_smem_raw[];
  TcSmem& sm = *reinterpret_cast<TcSmem*>(tc_smem_raw);
  static_assert(O == 1 || O == 3, "density or rgb head");
  constexpr int NP = O * kTcH + O;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int ncols = n_samples * O;  // <= 304 valid accumulator columns
  const bool two_halves = ncols > kTcN0;  // the density head's 100 columns fit the first accumulator half
  unsigned char* a_hi_b = reinterpret_cast<unsigned char*>(sm.a_hi);
  unsigned char* a_lo_b = reinterpret_cast<unsigned char*>(sm.a_lo);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)),
                 "r"(kTcTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    mbar_init(&sm.mma_bar[0], 1);
    mbar_init(&sm.mma_bar[1], 1);
    mbar_fence_init();
  }
  // B: every sampled weight row (column n = draw * 3 + channel) split into hi / lo, canonical layout
  for (int idx = tid; idx < kTcN * (kTcH / 4); idx += kTcThreads) {
    const int n = idx / (kTcH / 4), j = idx % (kTcH / 4);
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < ncols) {
      const float* src = params + (size_t)(n / O) * NP + (n % O) * kTcH + 4 * j;
      w = make_float4(src[0], src[1], src[2], src[3]);
    }
    split_store(w, reinterpret_cast<unsigned char*>(sm.b_hi), reinterpret_cast<unsigned char*>(sm.b_lo),
                canon_off(n, j));
  }
  for (int n = tid; n < kTcN; n += kTcThreads) {
    const float b = n < ncols ? params[(size_t)(n / O) * NP + O * kTcH + (n % O)] : 0.f;
    sm.bias[n] = ACT == UB_ACT_SIGMOID ? -b * 1.4426950408889634f : b;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;
  const uint32_t a_hi = smem_u32(sm.a_hi), a_lo = smem_u32(sm.a_lo);
  const uint32_t b_hi = smem_u32(sm.b_hi), b_lo = smem_u32(sm.b_lo);
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;  // this warp's TMEM lane quarter
  const int row = (warp & 3) * 32 + (tid & 31);                   // point of the tile this thread owns
  const int group = warp >> 2;                                    // column group / K-chunk quarter

  const long long num_tiles = (num_points + kTcM - 1) / kTcM;
  float4 xr[kTcChunksPerThread];
  auto prefetch = [&](long long tile) {
    const long long p = tile * kTcM + row;
    const bool in = tile < num_tiles && p < num_points;
    const float4* src = reinterpret_cast<const float4*>(x + (size_t)(in ? p : 0) * kTcH) + group * kTcChunksPerThread;
#pragma unroll
    for (int jj = 0; jj < kTcChunksPerThread; ++jj) xr[jj] = in ? src[jj] : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  prefetch(blockIdx.x);

  uint32_t phase = 0;
  for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const long long pt = tile * kTcM + row;
    const bool ok = pt < num_points;
    // ---- A tile from the prefetched registers: thread = (point row, quarter of the 16 K-chunks) ----
#pragma unroll
    for (int jj = 0; jj < kTcChunksPerThread; ++jj)
      split_store(xr[jj], a_hi_b, a_lo_b, canon_off(row, group * kTcChunksPerThread + jj));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
    __syncthreads();

    // ---- MMA: one thread issues 2 halves x 8 K-steps x 3 split products, one commit per half ----
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (O == 1) {
        // One output: the 100 draws fit one accumulator width, so three accumulators (TMEM columns 0 / 160 / 320) take
        // the products apart -- hi*hi of K-steps 0-3, hi*hi of K-steps 4-7, all cross terms -- and the epilogue adds them
        // in float32.  The tensor core adds into its accumulator with truncation; after exp an absolute error of the
        // pre-activation is a relative error of the output, and 24 accumulations into one accumulator cost 1.1e-5 at
        // v ~ 7.  Four per accumulator (the cross terms are 2^-11 smaller) stay inside the 1e-5 contract.
        const uint32_t idesc = umma_idesc_tf32(kTcM, kTcN0);
#pragma unroll
        for (int ks = 0; ks < kTcH / 8; ++ks) {
          const uint32_t ko = (uint32_t)ks * 256u;
          const uint64_t dah = umma_desc(a_hi + ko), dal = umma_desc(a_lo + ko);
          const uint64_t dbh = umma_desc(b_hi + ko), dbl = umma_desc(b_lo + ko);
          umma_tf32(tmem + (ks < 4 ? 0u : (uint32_t)kTcN0), dah, dbh, idesc, (ks & 3) ? 1u : 0u);
          umma_tf32(tmem + 2u * kTcN0, dal, dbh, idesc, ks > 0 ? 1u : 0u);
          umma_tf32(tmem + 2u * kTcN0, dah, dbl, idesc, 1u);
        }
        umma_commit(&sm.mma_bar[0]);
        umma_commit(&sm.mma_bar[1]);
      } else {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint32_t d = tmem + (half ? kTcTmemCol1 : 0u);
        const uint32_t idesc = half ? umma_idesc_tf32(kTcM, kTcN1) : umma_idesc_tf32(kTcM, kTcN0);
        const uint32_t brow = half ? (uint32_t)(kTcN0 / 8) * kTcRowGroupBytes : 0u;
#pragma unroll
        for (int ks = 0; ks < kTcH / 8 && (half == 0 || two_halves); ++ks) {
          const uint32_t ko = (uint32_t)ks * 256u;  // two 128-byte K chunks per instruction
          const uint64_t dah = umma_desc(a_hi + ko), dal = umma_desc(a_lo + ko);
          const uint64_t dbh = umma_desc(b_hi + brow + ko), dbl = umma_desc(b_lo + brow + ko);
          umma_tf32(d, dah, dbh, idesc, ks > 0 ? 1u : 0u);
          umma_tf32(d, dal, dbh, idesc, 1u);
          umma_tf32(d, dah, dbl, idesc, 1u);
        }
        umma_commit(&sm.mma_bar[half]);
      }
      }
    }
    prefetch(tile + gridDim.x);  // global loads of the next tile fly while this one is in the tensor core

    // ---- epilogue: this thread's point, 16-column chunks c = group, group + 4, ... ----
    float mu[3] = {0.f, 0.f, 0.f}, mu2[3] = {0.f, 0.f, 0.f};
    bool waited1 = false;
    mbar_wait(&sm.mma_bar[0], phase);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
    for (int c = group; c * 16 < ncols; c += kTcGroups) {
      const int c0 = c * 16;  // a 16-column chunk never straddles the two accumulator halves
      if (c0 >= kTcN0 && !waited1) {
        mbar_wait(&sm.mma_bar[1], phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        waited1 = true;
      }
      const uint32_t col = c0 < kTcN0 ? (uint32_t)c0 : kTcTmemCol1 + (uint32_t)(c0 - kTcN0);
      float v[16];
      tmem_ld16(tmem + lane_base + col, v);
      if (O == 1) {  // hi*hi of the second half of K, then the cross terms
        float v1[16], v2[16];
        tmem_ld16(tmem + lane_base + (uint32_t)kTcN0 + col, v1);
        tmem_ld16(tmem + lane_base + 2u * kTcN0 + col, v2);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __fadd_rn(__fadd_rn(v[i], v1[i]), v2[i]);
      }
      float t[3] = {0.f, 0.f, 0.f}, t2[3] = {0.f, 0.f, 0.f};  // by (i % O), relative to the chunk start
      if (c0 + 16 <= ncols) epilogue_chunk<ACT, O, true>(v, sm.bias + c0, 16, t, t2);
      else epilogue_chunk<ACT, O, false>(v, sm.bias + c0, ncols - c0, t, t2);
      const int ph = c % O;  // channel of the chunk's first column: (16 c) % 3 == c % 3 (0 for one output)
      if (ph == 0) {
        mu[0] += t[0]; mu[1] += t[1]; mu[2] += t[2]; mu2[0] += t2[0]; mu2[1] += t2[1]; mu2[2] += t2[2];
      } else if (ph == 1) {
        mu[1] += t[0]; mu[2] += t[1]; mu[0] += t[2]; mu2[1] += t2[0]; mu2[2] += t2[1]; mu2[0] += t2[2];
      } else {
        mu[2] += t[0]; mu[0] += t[1]; mu[1] += t[2]; mu2[2] += t2[0]; mu2[0] += t2[1]; mu2[1] += t2[2];
      }
    }
    if (!waited1) mbar_wait(&sm.mma_bar[1], phase);  // every thread observes both phases of this tile
    phase ^= 1u;
    // combine the column groups of a point (the A tile is no longer read: all MMAs have completed)
    float* red = sm.a_hi;  // [kTcGroups][kTcM][8]
    {
      float4* dst = reinterpret_cast<float4*>(red + ((size_t)group * kTcM + row) * 8);
      dst[0] = make_float4(mu[0], mu[1], mu[2], 0.f);
      dst[1] = make_float4(mu2[0], mu2[1], mu2[2], 0.f);
    }
    __syncthreads();
    if (group == 0 && ok) {
      float sm1[3] = {0.f, 0.f, 0.f}, sm2[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int g = 0; g < kTcGroups; ++g) {
        const float4 a = *reinterpret_cast<const float4*>(red + ((size_t)g * kTcM + row) * 8);
        const float4 b = *reinterpret_cast<const float4*>(red + ((size_t)g * kTcM + row) * 8 + 4);
        sm1[0] += a.x; sm1[1] += a.y; sm1[2] += a.z;
        sm2[0] += b.x; sm2[1] += b.y; sm2[2] += b.z;
      }
      const float nf = (float)n_samples;
#pragma unroll
      for (int o = 0; o < O; ++o) {
        const float m = sm1[o] / nf, m2 = sm2[o] / nf;
        if (o_mean) o_mean[pt * O + o] = m;
        if (o_mean2) o_mean2[pt * O + o] = m2;
        if (o_sigma2) o_sigma2[pt * O + o] = m2 - m * m;
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // accumulators, the A tile and the combine buffer may be overwritten
  }

  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTcTmemCols));
  }
}

template <int ACT, int O>
static int launch_tc(const float* x, long long num_points, const float* params, int n_samples, float* o_mean,
                     float* o_mean2, float* o_sigma2, cudaStream_t stream) {
  const size_t smem = sizeof(TcSmem) + 128;
  auto kern = laplace_moments_tc_kernel<ACT, O>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return UB_ERR_UNSUPPORTED;
  }
  const long long tiles = (num_points + kTcM - 1) / kTcM;
  long long blocks = sm_count() > 0 ? sm_count() : 148;
  if (tiles < blocks) blocks = tiles;
  kern<<<(unsigned)blocks, kTcThreads, smem, stream>>>(x, num_points, params, n_samples, o_mean, o_mean2, o_sigma2);
  return check_launch("laplace_ll_moments (tcgen05)");
}

// returns UB_ERR_UNSUPPORTED when the shape does not fit this path (caller falls back to the FMA kernel)
int launch_laplace_tc(const float* x, long long num_points, const float* params, int n_samples, int out_dim,
                      int act, float* o_mean, float* o_mean2, float* o_sigma2, cudaStream_t stream) {
  if (out_dim == 3 ? n_samples * 3 > kTcN : (out_dim != 1 || n_samples > kTcN0)) return UB_ERR_UNSUPPORTED;
#define UB_TC(A)                                                                                              \
  return out_dim == 3 ? launch_tc<A, 3>(x, num_points, params, n_samples, o_mean, o_mean2, o_sigma2, stream) \
                      : launch_tc<A, 1>(x, num_points, params, n_samples, o_mean, o_mean2, o_sigma2, stream)
  if (act == UB_ACT_SIGMOID) UB_TC(UB_ACT_SIGMOID);
  if (act == UB_ACT_EXP) UB_TC(UB_ACT_EXP);
  UB_TC(UB_ACT_IDENTITY);
#undef UB_TC
}

}  // namespace ub
