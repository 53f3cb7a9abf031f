// Two-site bond gradient on the fp16 tensor pipe (north_star piece 3), the 16-bit sibling of grad_tc.cu.
//   G[a, c] = sum_t P[t, a] * W[t, c],   P[t,(m,i)] = phi_near[t,m] * e_near[t,i],   W[t,(l,n,k)] = sres[t,(l,n)] * e_far[t,k]
// (optimizer.py:249, without the regulariser).  M = 2*Ds rows, N = 2*L*Ds columns, K = images.
//
// Why: the TF32 kernel is bound by shared-memory traffic (per 16-image chunk 144 KB of operand fetch + 64 KB of tile
// stores at 128 B/clk).  kind::f16 covers 16 images per instruction for the same operand bytes, and fp16 tiles are half
// as large, so the traffic per image halves.  FP32 accuracy is kept with hi/lo halves of both operands
// (h.h + l.h + h.l, 22 mantissa bits, like 3xTF32).
//
// Scaling.  The contraction index is the image, so each image contributes a rank-1 term and may be scaled freely as
// long as the two factors compensate: P^_t = P_t 2^-a_t with the row maximum in [2^12, 2^13), and
// W^_t = W_t 2^(a_t - g') with ONE global g' that puts the largest |W^| into [2^13, 2^14)
// (grad_scale_kernel: one warp per image, integer atomicMax -- deterministic).  Images far below the largest
// contribution lose relative precision in fp16's subnormal range, i.e. an absolute error 2^-38 of the largest term.
// G = 2^g' * sum_t P^_t (x) W^_t; the factor is applied by the kernel that sums the split-K partials, before the
// data-parallel all-reduce.
//
// Layout.  Both operands K-major SWIZZLE_64B: a stage holds 32 images = one 64-byte row per (m,i) / (l,n,k).
// A producer thread loads float4 = 4 consecutive rows for 8 consecutive images, transposes in registers and stores
// four 16-byte chunks (8 images of one row each).  The 4 rows of a quad q go to PHYSICAL rows e*64 + q (e = 0..3):
// the 32 lanes of a warp then store to 32 consecutive rows, which is bank-conflict free under the 64-byte swizzle
// (4 consecutive rows would hit 2 bank groups only).  The permutation is undone for free: the epilogue thread that
// owns accumulator row e*64+q writes logical row 4q+e, and reads the four column planes e' = 0..3 of 8 quads to write
// 32 consecutive logical columns.
//
// Warp roles (544 threads): warps 0-7 produce A tiles, warps 8-15 produce B tiles (warps 0-7 also drain TMEM at the
// end), warp 16 issues the MMAs; per-stage `full` mbarriers (512 arrivals) feed the issuer, tcgen05.commit recycles
// the stage.  One CTA = 256 x 256 output tile = two 128 x 256 accumulators (all 512 TMEM columns); split-K over the
// images, partials summed in fixed order.
#include <cuda_fp16.h>
#include "common.cuh"

namespace trmps {

namespace {

constexpr int GH_KCH = 32;                  // images per stage (two K=16 MMA steps)
constexpr int GH_STAGES = 3;
constexpr int GH_TM = 256, GH_TN = 256;     // CTA output tile
constexpr int GH_PART_BYTES = 256 * 64;     // one (hi | lo) operand tile: 256 rows x 32 fp16
constexpr int GH_STAGE_BYTES = 4 * GH_PART_BYTES;     // A_hi, A_lo, B_hi, B_lo = 64 KB
constexpr int GH_FORMERS = 512, GH_THREADS = GH_FORMERS + 32;
constexpr int GH_SMEM_BYTES = GH_STAGES * GH_STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int GH_TA = 12, GH_TW = 13;       // targets for the row maxima of P^ and for the global maximum of W^
constexpr int GH_BIAS = 1000;               // keeps the atomicMax'ed exponent positive (0 = nothing seen)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {      // sleeps in hardware (suspend-time hint)
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity), "r"(1000000u) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_64B descriptor: 64-byte rows, 8-row atoms 512 B apart
__device__ __forceinline__ uint64_t make_desc_k64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;                    // SWIZZLE_64B
  return d;
}
// byte offset of the 16-byte chunk `ch` (8 fp16 k-values) of row r
__device__ __forceinline__ uint32_t k64_chunk_off(int r, int ch) {
  return (uint32_t)((r >> 3) * 512 + (r & 7) * 64 + (((ch ^ (r >> 1)) & 3) << 4));
}
__device__ __forceinline__ float pow2f(int e) { return __int_as_float((e + 127) << 23); }     // -126 <= e <= 127
__device__ __forceinline__ int exponent_of(float x) { return (int)((__float_as_uint(x) >> 23) & 0xffu) - 127; }
__device__ __forceinline__ float scale_pow2(float x, int e) {
  e = max(-250, min(250, e));
  const int e1 = e / 2;
  return x * pow2f(e1) * pow2f(e - e1);
}

// ---- per-image exponents: a[t] scales P_t, atomicMax collects max_t (exponent of |W_t| + a[t]).  One warp per image.
__global__ void grad_scale_kernel(const float* __restrict__ e_near, const float* __restrict__ e_far,
                                  const float* __restrict__ phi_near, const float* __restrict__ sres, int64_t B, int Ds, int L2,
                                  int32_t* __restrict__ a_out, int32_t* __restrict__ gmax) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int n4 = Ds >> 2;                      // Ds is a multiple of 8: rows are read as float4 (two images per pass: four rows in flight)
  int best = 0;
  for (int64_t t0 = 2 * gw; t0 < B; t0 += 2 * nw) {
    float mn[2] = {0.f, 0.f}, mf[2] = {0.f, 0.f}, ms[2] = {0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int64_t t = t0 + q;
      if (t < B) {
        for (int i = lane; i < n4; i += 32) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(e_near + t * Ds) + i);
          const float4 b = __ldg(reinterpret_cast<const float4*>(e_far + t * Ds) + i);
          mn[q] = fmaxf(mn[q], fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
          mf[q] = fmaxf(mf[q], fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
        }
        for (int i = lane; i < L2; i += 32) ms[q] = fmaxf(ms[q], fabsf(__ldg(sres + t * L2 + i)));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        mn[q] = fmaxf(mn[q], __shfl_xor_sync(0xffffffffu, mn[q], o));
        mf[q] = fmaxf(mf[q], __shfl_xor_sync(0xffffffffu, mf[q], o));
        ms[q] = fmaxf(ms[q], __shfl_xor_sync(0xffffffffu, ms[q], o));
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int64_t t = t0 + q;
      if (t < B) {
        const float2 ph = __ldg(reinterpret_cast<const float2*>(phi_near) + t);
        const float mp = mn[q] * fmaxf(fabsf(ph.x), fabsf(ph.y)), mw = ms[q] * mf[q];
        int a = 0;
        if (mp > 0.f && mp < 3.0e38f) a = max(-100, min(100, exponent_of(mp) - GH_TA));
        if (lane == 0) a_out[t] = a;
        if (mw > 0.f && mw < 3.0e38f) best = max(best, exponent_of(mw) + a + GH_BIAS);
      }
    }
  }
  if (lane == 0 && best > 0) atomicMax(gmax, best);
}
__device__ __forceinline__ int grad_exponent(const int32_t* gmax) {          // g'
  const int v = __ldg(gmax);
  return v > 0 ? v - GH_BIAS - GH_TW : 0;
}

struct GradHArgs {
  const float* e_near; const float* e_far; const float* phi_near; const float* sres; const int32_t* a_t; const int32_t* gmax;
  float* gpart; int64_t B, imgs_per_split; int Ds, L, R, Cc;
};

__global__ void __launch_bounds__(GH_THREADS, 1) grad_h_kernel(GradHArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GH_STAGES * GH_STAGE_BYTES);   // empty[STAGES], done, full[STAGES]
  uint64_t* full = bars + GH_STAGES + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GH_STAGES + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Ds = g.Ds, L2 = 2 * g.L;
  const int c0 = blockIdx.x * GH_TN, a0 = blockIdx.y * GH_TM;
  const int64_t tbeg = (int64_t)blockIdx.z * g.imgs_per_split;
  int64_t tend = tbeg + g.imgs_per_split;
  if (tend > g.B) tend = g.B;
  const int nch = tend > tbeg ? (int)((tend - tbeg + GH_KCH - 1) / GH_KCH) : 0;

  if (tid == 0) {
    for (int s = 0; s <= GH_STAGES; ++s) mbar_init(&bars[s], 1);
    for (int s = 0; s < GH_STAGES; ++s) mbar_init(&full[s], GH_FORMERS);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (tid < GH_FORMERS) {
    // ===== producers: side 0 forms A (rows (m,i)), side 1 forms B (columns (l,n,k)) =====
    const int side = tid >> 8, p = tid & 255, q4 = p & 63, kgrp = p >> 6;
    const int first = (side == 0 ? a0 : c0) + 4 * q4;
    const bool valid = first < (side == 0 ? g.R : g.Cc);
    const int hiidx = valid ? first / Ds : 0, lowidx = valid ? first - hiidx * Ds : 0;      // (m | l*2+n) and (i | k)
    const float* env = side == 0 ? g.e_near : g.e_far;
    const float* fac = side == 0 ? g.phi_near : g.sres;
    const int fstride = side == 0 ? 2 : L2;
    const int gexp = grad_exponent(g.gmax);
    // Loads of a stage: branch-free (the image index is clamped, the tail is masked when the values are used) and nothing
    // is computed from them here -- with `if (t < tend) { load; scale }` every one of the eight images became a branch with a
    // dependent FMUL behind its loads, i.e. eight exposed L2 round trips per stage (ncu: 60 % of the kernel's stall samples on the
    // long scoreboard, tensor pipe 28 % active)
    float4 v[8];
    float fr[8];
    int ar[8];
    // (image t0 = first image of this thread in the stage; whole stages inside the range take the pointer path: one 64-bit
    //  multiply per stage instead of three per image -- the producers are issue bound once the loads overlap)
    auto load_stage = [&](int ch) {
      const int64_t t0 = tbeg + (int64_t)ch * GH_KCH + kgrp * 8;
      if (t0 + 8 <= tend) {
        const float* pe = env + t0 * Ds + lowidx;
        const float* pf = fac + t0 * fstride + hiidx;
        const int32_t* pa = g.a_t + t0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          v[u] = __ldg(reinterpret_cast<const float4*>(pe + u * Ds));
          ar[u] = __ldg(pa + u);
          fr[u] = __ldg(pf + u * fstride);
        }
      } else {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          int64_t t = t0 + u;
          t = t < tend ? t : tend - 1;
          v[u] = __ldg(reinterpret_cast<const float4*>(env + t * Ds + lowidx));
          ar[u] = __ldg(g.a_t + t);
          fr[u] = __ldg(fac + t * fstride + hiidx);
        }
      }
    };
    if (nch > 0) load_stage(0);
    for (int ch = 0; ch < nch; ++ch) {
      const int s = ch % GH_STAGES, it = ch / GH_STAGES;
      uint8_t* tile_hi = smem + s * GH_STAGE_BYTES + side * 2 * GH_PART_BYTES, *tile_lo = tile_hi + GH_PART_BYTES;
      if (ch >= GH_STAGES) mbar_wait(&bars[s], (uint32_t)((it - 1) & 1));   // MMAs that read this stage are done
      float f[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t t = tbeg + (int64_t)ch * GH_KCH + kgrp * 8 + u;
        const float sc = scale_pow2(fr[u], side == 0 ? -ar[u] : ar[u] - gexp);
        f[u] = (valid && t < tend) ? sc : 0.f;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        __align__(16) __half2 h2[4];
        __align__(16) __half2 l2[4];
#pragma unroll
        for (int u = 0; u < 8; u += 2) {
          const float x = (e == 0 ? v[u].x : e == 1 ? v[u].y : e == 2 ? v[u].z : v[u].w) * f[u];
          const float y = (e == 0 ? v[u + 1].x : e == 1 ? v[u + 1].y : e == 2 ? v[u + 1].z : v[u + 1].w) * f[u + 1];
          h2[u >> 1] = __floats2half2_rn(x, y);
          const float2 hf = __half22float2(h2[u >> 1]);
          l2[u >> 1] = __floats2half2_rn(x - hf.x, y - hf.y);
        }
        const uint32_t off = k64_chunk_off(e * 64 + q4, kgrp);        // physical row e*64 + q4
        *reinterpret_cast<uint4*>(tile_hi + off) = *reinterpret_cast<const uint4*>(h2);
        *reinterpret_cast<uint4*>(tile_lo + off) = *reinterpret_cast<const uint4*>(l2);
      }
      if (ch + 1 < nch) load_stage(ch + 1);      // in flight while the other warps convert and while this one waits
      fence_async_proxy();          // make the generic-proxy stores visible to the tensor core (async proxy)
      mbar_arrive(&full[s]);
    }
  } else if (tid >= GH_FORMERS) {
    // ===== MMA issuer: the whole warp runs the loop (uniform descriptor arithmetic), one elected lane issues =====
    uint32_t el_;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(el_));
    const bool issuer = el_ != 0;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(GH_TN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f16 x f16 -> f32, K-major
    for (int ch = 0; ch < nch; ++ch) {
      const int s = ch % GH_STAGES, it = ch / GH_STAGES;
      mbar_wait(&full[s], (uint32_t)(it & 1));
      tc_fence_after();
      const uint32_t base = smem_u32(smem + s * GH_STAGE_BYTES);
      const uint64_t a_hi = make_desc_k64(base), a_lo = make_desc_k64(base + GH_PART_BYTES);
      const uint64_t b_hi = make_desc_k64(base + 2 * GH_PART_BYTES), b_lo = make_desc_k64(base + 3 * GH_PART_BYTES);
#pragma unroll
      for (int ks = 0; ks < GH_KCH / 16; ++ks) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint64_t ro = (uint64_t)(mt * (128 * 64 >> 4) + 2 * ks);       // rows 128.. of the A tile; +2 units per K=16
          const uint32_t d = tmem_base + mt * GH_TN;
          if (issuer) umma_f16(d, a_hi + ro, b_hi + 2 * ks, idesc, (ch > 0 || ks > 0) ? 1u : 0u);
          if (issuer) umma_f16(d, a_lo + ro, b_hi + 2 * ks, idesc, 1u);
          if (issuer) umma_f16(d, a_hi + ro, b_lo + 2 * ks, idesc, 1u);
        }
      }
      if (issuer) umma_commit(&bars[s]);                          // stage s may be overwritten once these MMAs have read it
      if (ch == nch - 1 && issuer) umma_commit(&bars[GH_STAGES]);
    }
  }

  // ===== epilogue (warps 0-7): TMEM -> registers -> global partial tile, undoing the row/column permutation =====
  float* out = g.gpart + (size_t)blockIdx.z * g.R * g.Cc;
  if (warp < 8) {
    const int mt = warp >> 2, quad = warp & 3;
    const int phys = mt * 128 + quad * 32 + lane;                 // physical accumulator row e*64 + q
    const int row = a0 + 4 * (phys & 63) + (phys >> 6);
    if (nch > 0) {
      mbar_wait(&bars[GH_STAGES], 0);
      tc_fence_after();
#pragma unroll 1
      for (int qb = 0; qb < 8; ++qb) {             // 8 quads = 32 consecutive logical columns
        float pl[4][8];
#pragma unroll
        for (int e = 0; e < 4; ++e) tmem_ld8(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mt * GH_TN + e * 64 + qb * 8), pl[e]);
        tmem_ld_wait();
        if (row < g.R) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int col = c0 + 4 * (qb * 8 + q);
            if (col < g.Cc) *reinterpret_cast<float4*>(out + (size_t)row * g.Cc + col) = make_float4(pl[0][q], pl[1][q], pl[2][q], pl[3][q]);
          }
        }
      }
    } else if (row < g.R) {
      for (int col = c0; col < c0 + GH_TN && col < g.Cc; ++col) out[(size_t)row * g.Cc + col] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// out = 2^g' * sum of the split-K partials (fixed order)
__global__ void sum_parts_scaled_kernel(const float4* __restrict__ part, int nparts, int64_t n4, float4* __restrict__ out,
                                        const int32_t* __restrict__ gmax) {
  const int ex = grad_exponent(gmax);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 s = part[i];
    for (int p = 1; p < nparts; ++p) {
      const float4 v = part[(int64_t)p * n4 + i];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    out[i] = make_float4(scale_pow2(s.x, ex), scale_pow2(s.y, ex), scale_pow2(s.z, ex), scale_pow2(s.w, ex));
  }
}

}  // namespace

bool grad_h_supported(int64_t B, int Ds, int L) { return Ds % 4 == 0 && B + 8 <= (int64_t)2 * Ds * 2 * L * Ds; }

// gradient partials (fp16 pipe) + their scaled sum into `out`.  scratch: B + 1 int32 (per-image exponents, global maximum)
int launch_grad_h(const float* e_near, const float* e_far, const float* phi_near, const float* sres, float* gpart, int gsplit,
                  int64_t B, int Ds, int L, int32_t* scratch, float* out, cudaStream_t st) {
  int32_t* a_t = scratch;
  int32_t* gmax = scratch + B;
  TRMPS_CUDA_OK(cudaMemsetAsync(gmax, 0, sizeof(int32_t), st));
  grad_scale_kernel<<<592, 256, 0, st>>>(e_near, e_far, phi_near, sres, B, Ds, 2 * L, a_t, gmax);
  TRMPS_LAUNCH_OK("grad_scale_kernel");
  GradHArgs g;
  g.e_near = e_near; g.e_far = e_far; g.phi_near = phi_near; g.sres = sres; g.a_t = a_t; g.gmax = gmax; g.gpart = gpart;
  g.B = B; g.Ds = Ds; g.L = L; g.R = 2 * Ds; g.Cc = 2 * L * Ds;
  const int64_t per = ceil_div64(B, gsplit);
  g.imgs_per_split = ceil_div64(per, GH_KCH) * GH_KCH;
  static bool attr_set = false;
  if (!attr_set) {
    TRMPS_CUDA_OK(cudaFuncSetAttribute(grad_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GH_SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((unsigned)((g.Cc + GH_TN - 1) / GH_TN), (unsigned)((g.R + GH_TM - 1) / GH_TM), (unsigned)gsplit);
  grad_h_kernel<<<grid, GH_THREADS, GH_SMEM_BYTES, st>>>(g);
  TRMPS_LAUNCH_OK("grad_h_kernel");
  const int64_t n4 = (int64_t)g.R * g.Cc / 4;
  const int blocks = (int)(ceil_div64(n4, 256) < 1184 ? ceil_div64(n4, 256) : 1184);
  sum_parts_scaled_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(gpart), gsplit, n4, reinterpret_cast<float4*>(out), gmax);
  TRMPS_LAUNCH_OK("sum_parts_scaled_kernel");
  return TRMPS_OK;
}

}  // namespace trmps
