// Forward contraction f[t,l] of a bond matrix on the 5th-generation tensor cores (north_star piece 3;
// optimizer.py:222-227: f = tensordot(C, bond), without ever forming C).
//
//   X[t, (k; l,m,n)] = sum_i e_near[t,i] * Bm[(m,i), (l,n,k)]          GEMM  M = images, K = near bond, N = 4*L*Ds
//   f[t,l]           = sum_{k,m,n} X[t,(k;l,m,n)] * phi_near[t,m] phi_far[t,n] e_far[t,k]      epilogue
//
// The feature vectors are moved out of the GEMM into the epilogue weight, so the A operand is the raw
// environment tile (split once into TF32 hi/lo, K-major SWIZZLE_128B, resident in shared memory for the whole
// CTA) and the B operand is the bond itself.  The bond is re-laid-out once per call (prep kernel, 2.3 MB) into
// ready-made hi/lo shared-memory images whose columns are ordered (k; l,m,n): a 16*L-column tile then holds 4
// values of k times all 4*L (l,m,n) combinations, which makes the epilogue completely static -- every
// accumulator column has a compile-time (k_local, l, m, n), the weights are 4 + 4 registers per tile and the
// per-label sums live in registers.  3xTF32 (hi*hi + lo*hi + hi*lo) keeps FP32-level accuracy.
//
// Warp roles (192 threads): warps 0-3 epilogue (one TMEM lane quadrant each), warp 4 issues the MMAs, warp 5
// streams the B images with cp.async.bulk into a 3-stage mbarrier ring.  The accumulator is double buffered in
// TMEM so the epilogue of tile j overlaps the MMAs of tile j+1.
#include <cuda_fp16.h>
#include <cstdlib>
#include "common.cuh"

namespace trmps {

namespace {

constexpr int FT_THREADS = 192;
constexpr int FT_TM = 128;                 // images per CTA
constexpr int FT_KMAX = 128;               // near bond (padded) handled by the resident A tile
constexpr int FT_KCH = 16;                 // near-bond indices per B stage
constexpr int FT_STAGES = 3;
constexpr int FT_A_BYTES = FT_TM * FT_KMAX * 4;      // 64 KB per (hi | lo)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) { } }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tf32_round(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
// smem matrix descriptor: layout 2 = SWIZZLE_128B (K-major A), 1 = SWIZZLE_128B_BASE32B (MN-major B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// byte offset inside one B image (MN extent NT columns, 16 k): SWIZZLE_128B_BASE32B, groups of 4 k are NT/32 atoms apart
__host__ __device__ __forceinline__ uint32_t bimg_off(int col, int k, int nt) {
  const int blk = col >> 5, c32 = (col & 31) >> 3, e = col & 7, kg = k >> 2, k4 = k & 3;
  return (uint32_t)((kg * (nt / 32) + blk) * 512 + k4 * 128 + ((c32 ^ k4) << 5) + e * 4);
}

// ---------------------------------------------------------------------------------------------------
// prep: bond matrix M (R x Cc; rows (m,i), cols (l,n,k)) -> hi/lo B images, [tile j][chunk q][hi|lo][image bytes]
// tile j holds k = 4j .. 4j+3, column inside the tile = k_local*(4L) + l*4 + m*2 + n ; chunk q holds i = 16q .. 16q+15
template <int L>
__global__ void forward_prep_kernel(const float* __restrict__ M, float* __restrict__ bimg, int Ds, int ntiles, int nchunks) {
  constexpr int NT = 16 * L, IMG = 16 * NT;   // floats per (hi | lo) image
  const int64_t Cc = (int64_t)2 * L * Ds;
  const int64_t total = (int64_t)ntiles * nchunks * IMG;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(e % NT);
    const int il = (int)((e / NT) % 16);
    const int64_t jq = e / IMG;
    const int q = (int)(jq % nchunks), j = (int)(jq / nchunks);
    const int kl = col / (4 * L), r = col % (4 * L), l = r >> 2, m = (r >> 1) & 1, n = r & 1;
    const int k = 4 * j + kl, i = 16 * q + il;
    float v = 0.f;
    if (k < Ds && i < Ds) v = M[(int64_t)(m * Ds + i) * Cc + (int64_t)(l * 2 + n) * Ds + k];
    const float hi = tf32_round(v), lo = tf32_round(v - hi);
    float* img = bimg + jq * (2 * IMG);
    const uint32_t off = bimg_off(col, il, NT) >> 2;
    img[off] = hi;
    img[IMG + off] = lo;
  }
}

struct FwdTcArgs {
  const float* e_near; const float* e_far; const float* phi_near; const float* phi_far; const float* bimg;
  float* f; int64_t B; int Ds; const int32_t* d_near; const int32_t* d_far; int nchunks_max;
};

template <int L>
__global__ void __launch_bounds__(FT_THREADS, 1) forward_tc_kernel(FwdTcArgs a) {
  constexpr int NT = 16 * L;                         // accumulator columns per tile
  constexpr int IMG_BYTES = 16 * NT * 4;             // one (hi | lo) B image
  constexpr int STAGE_BYTES = 2 * IMG_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * NT <= 64 ? 64 : 2 * NT <= 128 ? 128 : 2 * NT <= 256 ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_hi = smem;
  uint8_t* a_lo = smem + FT_A_BYTES;
  uint8_t* bst = smem + 2 * FT_A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bst + FT_STAGES * STAGE_BYTES);
  uint64_t* b_full = bars, *b_empty = bars + FT_STAGES, *acc_full = bars + 2 * FT_STAGES, *acc_empty = bars + 2 * FT_STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * FT_STAGES + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Ds = a.Ds;
  const int64_t t0 = (int64_t)blockIdx.x * FT_TM;
  const int dn = __ldg(a.d_near), df = __ldg(a.d_far);
  const int nchunks = (dn + FT_KCH - 1) / FT_KCH;       // K chunks actually needed
  const int ntiles = (df + 3) / 4;                      // column tiles actually needed (4 k per tile)

  if (tid == 0) {
    for (int s = 0; s < FT_STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);

  // ---- A tile: e_near rows of my 128 images, split into tf32 hi / lo, K-major SWIZZLE_128B
  //      element (row t, k): atom (kb = k/32, tb = t/8) at (kb*16 + tb)*1024, row t%8 of 128 B, 16-byte chunk (k%32)/4 ^ (t%8)
  {
    const int kq = nchunks * FT_KCH / 4;      // float4 columns to fill (multiple of 4)
    for (int e = tid; e < FT_TM * kq; e += FT_THREADS) {
      const int t = e / kq, k4 = e - t * kq, k = 4 * k4;
      const int64_t tg = t0 + t;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tg < a.B && k < Ds) v = __ldg(reinterpret_cast<const float4*>(a.e_near + tg * Ds + k));
      float4 hi, lo;
      hi.x = tf32_round(v.x); hi.y = tf32_round(v.y); hi.z = tf32_round(v.z); hi.w = tf32_round(v.w);
      lo.x = tf32_round(v.x - hi.x); lo.y = tf32_round(v.y - hi.y); lo.z = tf32_round(v.z - hi.z); lo.w = tf32_round(v.w - hi.w);
      const int kb = k >> 5, tb = t >> 3, r = t & 7, ch = (k & 31) >> 2;
      const uint32_t off = (uint32_t)((kb * 16 + tb) * 1024 + r * 128 + ((ch ^ r) << 4));
      *reinterpret_cast<float4*>(a_hi + off) = hi;
      *reinterpret_cast<float4*>(a_lo + off) = lo;
    }
  }
  fence_async_proxy();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 5) {
    // ===== B loader: one elected lane streams (tile, chunk) images =====
    if (lane == 0) {
      int it = 0;
      for (int j = 0; j < ntiles; ++j) {
        for (int q = 0; q < nchunks; ++q, ++it) {
          const int s = it % FT_STAGES, ph = (it / FT_STAGES) & 1;
          if (it >= FT_STAGES) mbar_wait(&b_empty[s], (uint32_t)(ph ^ 1));
          mbar_expect_tx(&b_full[s], STAGE_BYTES);
          const float* src = a.bimg + ((int64_t)j * a.nchunks_max + q) * (STAGE_BYTES / 4);
          bulk_g2s(bst + s * STAGE_BYTES, src, STAGE_BYTES, &b_full[s]);
        }
      }
    }
  } else if (warp == 4) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(FT_TM >> 4) << 24);
      int it = 0;
      for (int j = 0; j < ntiles; ++j) {
        const int b = j & 1;
        if (j >= 2) mbar_wait(&acc_empty[b], (uint32_t)(((j >> 1) - 1) & 1));
        tc_fence_after();
        const uint32_t d = tmem_base + b * NT;
        for (int q = 0; q < nchunks; ++q, ++it) {
          const int s = it % FT_STAGES, ph = (it / FT_STAGES) & 1;
          mbar_wait(&b_full[s], (uint32_t)ph);
          tc_fence_after();
          const uint32_t bh = smem_u32(bst + s * STAGE_BYTES), bl = bh + IMG_BYTES;
#pragma unroll
          for (int ks = 0; ks < FT_KCH / 8; ++ks) {
            const int k0 = q * FT_KCH + ks * 8;                    // first near-bond index of this K=8 step
            const uint32_t aoff = (uint32_t)((k0 >> 5) * 16 * 1024 + ((k0 & 31) >> 3) * 32);
            const uint64_t dah = make_desc(smem_u32(a_hi) + aoff, 16, 1024, 2);
            const uint64_t dal = make_desc(smem_u32(a_lo) + aoff, 16, 1024, 2);
            const uint32_t boff = (uint32_t)(2 * ks * (NT / 32) * 512);   // two groups of 4 k per MMA
            const uint64_t dbh = make_desc(bh + boff, 512, (NT / 32) * 512, 1);
            const uint64_t dbl = make_desc(bl + boff, 512, (NT / 32) * 512, 1);
            const uint32_t acc = (q > 0 || ks > 0) ? 1u : 0u;
            umma_tf32(d, dah, dbh, idesc, acc);
            umma_tf32(d, dal, dbh, idesc, 1u);
            umma_tf32(d, dah, dbl, idesc, 1u);
          }
          umma_commit(&b_empty[s]);
        }
        umma_commit(&acc_full[b]);
      }
    }
  } else {
    // ===== epilogue warps 0-3: TMEM lanes 32*warp .. +31, one image row per thread =====
    const int row = warp * 32 + lane;
    const int64_t tg = t0 + row;
    const bool live = tg < a.B;
    float pp[4] = {0.f, 0.f, 0.f, 0.f};
    if (live) {
      const float2 p1 = __ldg(reinterpret_cast<const float2*>(a.phi_near) + tg);
      const float2 p2 = __ldg(reinterpret_cast<const float2*>(a.phi_far) + tg);
      pp[0] = p1.x * p2.x; pp[1] = p1.x * p2.y; pp[2] = p1.y * p2.x; pp[3] = p1.y * p2.y;   // index m*2+n
    }
    float facc[L];
#pragma unroll
    for (int l = 0; l < L; ++l) facc[l] = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      const int b = j & 1;
      float4 e4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live && 4 * j < Ds) e4 = __ldg(reinterpret_cast<const float4*>(a.e_far + tg * Ds + 4 * j));
      const float ek[4] = {e4.x, e4.y, e4.z, e4.w};
      mbar_wait(&acc_full[b], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
#pragma unroll
      for (int cb = 0; cb < NT / 32; ++cb) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(b * NT + cb * 32), v);
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int col = cb * 32 + e;                       // compile-time: (k_local; l, m, n)
          const int kl = col / (4 * L), r = col % (4 * L), l = r >> 2, mn = r & 3;
          facc[l] = fmaf(v[e], pp[mn] * ek[kl], facc[l]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[b]);
    }
    if (live) {
#pragma unroll
      for (int l = 0; l < L; ++l) a.f[tg * L + l] = facc[l];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

// =====================================================================================================
// fp16 hi/lo variant (kind::f16): the same contraction with both operands carried as two fp16 halves,
//     x = 2^e (h + l),  products h.h + l.h + h.l  (22 mantissa bits, like 3xTF32),
// at twice the TF32 rate (tools/probe/f16_probe.cu: an M=128 tcgen05.mma costs N/2 cycles for K=8 TF32 and for
// K=16 fp16 alike).  Exponents: one per image row of e_near (row maximum scaled into [2^12, 2^13)), one for the whole
// bond matrix (maximum into [2^13, 2^14)); both are exact powers of two and are undone in FP32 in the epilogue.
// Both operands K-major SWIZZLE_128B: A = [hi|lo][k atom][128 rows x 128 B], B image of (tile j, chunk q) =
// [hi|lo][160 rows (k_local; l,m,n) x 64 near-bond indices].
constexpr int FH_TROW = 12, FH_TB = 13;
constexpr int FH_A_BYTES = 4 * 16384;                       // hi ka0, hi ka1, lo ka0, lo ka1
constexpr int FH_KCH = 64;                                  // near-bond indices per B stage

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__host__ __device__ __forceinline__ uint32_t k128_chunk_off(int r, int ch) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + (((ch ^ r) & 7) << 4));
}
__device__ __forceinline__ float pow2f(int e) { return __int_as_float((e + 127) << 23); }     // -126 <= e <= 127
__device__ __forceinline__ int exponent_of(float x) { return (int)((__float_as_uint(x) >> 23) & 0xffu) - 127; }
__device__ __forceinline__ float scale_pow2(float x, int e) {
  e = max(-250, min(250, e));
  const int e1 = e / 2;
  return x * pow2f(e1) * pow2f(e - e1);
}

// max |M| as the bit pattern of a non-negative float (integer max is order preserving); *out zeroed by the caller
__global__ void absmax_kernel(const float* __restrict__ M, int64_t n, int32_t* __restrict__ out) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(M[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_int(m));
}
__device__ __forceinline__ int bond_exponent(const int32_t* maxbits) {
  const float m = __int_as_float(__ldg(maxbits));
  return (m > 0.f && m < 3.0e38f) ? exponent_of(m) - FH_TB : 0;
}

// bond matrix -> fp16 hi/lo images [tile j][chunk q][hi|lo][160 rows x 128 B]; one thread per (row, 8 near-bond indices)
template <int L>
__global__ void forward_prep_h_kernel(const float* __restrict__ M, uint8_t* __restrict__ bimg, const int32_t* __restrict__ maxbits,
                                      int Ds, int ntiles, int nchunks) {
  constexpr int NT = 16 * L, IMG_BYTES = NT * 128;
  const int64_t Cc = (int64_t)2 * L * Ds;
  const int ex = -bond_exponent(maxbits);
  const int64_t total = (int64_t)ntiles * nchunks * NT * 8;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(e % 8);
    const int col = (int)((e / 8) % NT);
    const int64_t jq = e / (8 * NT);
    const int q = (int)(jq % nchunks), j = (int)(jq / nchunks);
    const int kl = col / (4 * L), r = col % (4 * L), l = r >> 2, m = (r >> 1) & 1, n = r & 1;
    const int k = 4 * j + kl;
    __align__(16) __half hi[8];
    __align__(16) __half lo[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = q * FH_KCH + ch * 8 + u;
      float v = 0.f;
      if (k < Ds && i < Ds) v = M[(int64_t)(m * Ds + i) * Cc + (int64_t)(l * 2 + n) * Ds + k];
      v = scale_pow2(v, ex);
      hi[u] = __float2half_rn(v);
      lo[u] = __float2half_rn(v - __half2float(hi[u]));
    }
    uint8_t* img = bimg + jq * (2 * IMG_BYTES) + k128_chunk_off(col, ch);
    *reinterpret_cast<uint4*>(img) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(img + IMG_BYTES) = *reinterpret_cast<const uint4*>(lo);
  }
}

struct FwdHArgs {
  const float* e_near; const float* e_far; const float* phi_near; const float* phi_far; const uint8_t* bimg;
  const int32_t* maxbits; float* f; int64_t B; int Ds; const int32_t* d_near; const int32_t* d_far; int nchunks_max;
  // the image tiles beyond the last full wave are split by columns over `parts` CTAs each (tpp column tiles per CTA): their
  // partial logits go to pscr, the last CTA of an image tile to arrive (pcnt) adds them in a fixed order
  int n_full, parts, tpp; float* pscr; int32_t* pcnt;
};

// one lane of a converged warp (the lowest): the MMA warp runs its loop as a whole -- uniform control flow and descriptor
// arithmetic, which the compiler then keeps in the uniform datapath -- and this lane issues.  (A loop under `if (lane == 0)`
// made every descriptor a per-thread value: ELECT + 3 R2UR.BROADCAST per MMA, 154 cycles between MMAs against 95 of execution.)
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
  return p != 0;
}
constexpr int FH_PARTS_MAX = 6, FH_TAIL_MAX = 160, FH_PCNT = 256;   // column parts per tail tile, SMs sized for, counter words
constexpr int FH_THREADS = 320;        // warps 0-7: epilogue (TMEM lane quadrant x column half), 8: MMA issuer, 9: B loader

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// one accumulator tile, my half of its columns (16-column chunks [HALF*NCH, (HALF+1)*NCH)): every column has a
// compile-time (k_local; l, m, n).  ND = feature values of the far site carried by the columns: 2 for a two-site bond
// (4 far-bond indices per tile), 1 for the single-site form (columns (k_local; l, m), 8 far-bond indices per tile)
template <int L, int HALF, int ND>
__device__ __forceinline__ void fh_epilogue_tile(uint32_t taddr, const float (&pp)[4], const float (&ek)[8 / ND], float (&facc)[L]) {
  constexpr int NCH = (16 * L) / 32;          // 16-column chunks per half
  constexpr int PER = 2 * L * ND;             // columns per far-bond index
#pragma unroll
  for (int cb = 0; cb < NCH; ++cb) {
    float v[16];
    tmem_ld16(taddr + (uint32_t)((HALF * NCH + cb) * 16), v);
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int col = (HALF * NCH + cb) * 16 + e;
      const int kl = col / PER, r = col % PER, l = r / (2 * ND), mn = r % (2 * ND);
      facc[l] = fmaf(v[e], pp[mn] * ek[kl], facc[l]);
    }
  }
}

// single-site form of the forward (the logits of a bond that is still the product of its two sites, i.e. before the first
// update: f[t,l] = sum_{m,i,j} phi[t,m] e_near[t,i] S[l,m,i,j] z[t,j] with z the cached environment on the other side of
// the SAME site -- half the columns of the merged bond, half the MMAs).  Images of the special node S (L,2,Ds,Ds):
// [tile j][chunk q][hi|lo][160 rows (k_local; l, m) x 64 near-bond indices]; dir > 0: near index = S's third index
template <int L>
__global__ void forward_prep_site_kernel(const float* __restrict__ S, uint8_t* __restrict__ bimg, const int32_t* __restrict__ maxbits,
                                         int Ds, int dir, int ntiles, int nchunks) {
  constexpr int NT = 16 * L, IMG_BYTES = NT * 128, PER = 2 * L;
  const int ex = -bond_exponent(maxbits);
  const int64_t total = (int64_t)ntiles * nchunks * NT * 8;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(e % 8);
    const int col = (int)((e / 8) % NT);
    const int64_t jq = e / (8 * NT);
    const int q = (int)(jq % nchunks), j = (int)(jq / nchunks);
    const int kl = col / PER, r = col % PER, l = r >> 1, m = r & 1;
    const int k = 8 * j + kl;
    __align__(16) __half hi[8];
    __align__(16) __half lo[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = q * FH_KCH + ch * 8 + u;
      float v = 0.f;
      if (k < Ds && i < Ds) v = dir > 0 ? S[((int64_t)(l * 2 + m) * Ds + i) * Ds + k] : S[((int64_t)(l * 2 + m) * Ds + k) * Ds + i];
      v = scale_pow2(v, ex);
      hi[u] = __float2half_rn(v);
      lo[u] = __float2half_rn(v - __half2float(hi[u]));
    }
    uint8_t* img = bimg + jq * (2 * IMG_BYTES) + k128_chunk_off(col, ch);
    *reinterpret_cast<uint4*>(img) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(img + IMG_BYTES) = *reinterpret_cast<const uint4*>(lo);
  }
}

template <int L, int ND>
__global__ void __launch_bounds__(FH_THREADS, 1) forward_h_kernel(FwdHArgs a) {
  constexpr int NT = 16 * L, KPT = 8 / ND;             // KPT: far-bond indices per tile
  constexpr int IMG_BYTES = NT * 128;                // one (hi | lo) B image: NT rows x 64 fp16
  constexpr int STAGE_BYTES = 2 * IMG_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * NT <= 64 ? 64 : 2 * NT <= 128 ? 128 : 2 * NT <= 256 ? 256 : 512;
  constexpr int NSTAGE_WARPS = 9;                    // everyone but the loader stages the A tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_t = smem;
  uint8_t* bst = smem + FH_A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bst + FT_STAGES * STAGE_BYTES);
  uint64_t* b_full = bars, *b_empty = bars + FT_STAGES, *acc_full = bars + 2 * FT_STAGES, *acc_empty = bars + 2 * FT_STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * FT_STAGES + 4);
  int* rexp = reinterpret_cast<int*>(bars + 2 * FT_STAGES + 6);      // [128] row exponents
  float* xch = reinterpret_cast<float*>(rexp + FT_TM);               // [128][L] partial logits of the second column half
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Ds = a.Ds;
  int* s_last = reinterpret_cast<int*>(xch + FT_TM * L);
  const int dn = __ldg(a.d_near), df = __ldg(a.d_far);
  const int nchunks = (dn + FH_KCH - 1) / FH_KCH;
  const int ntiles_all = (df + KPT - 1) / KPT;
  // image tile and column-tile range [j0, j0 + ntiles) of this CTA
  int rt = (int)blockIdx.x, part = 0, j0 = 0, ntiles = ntiles_all;
  const bool partial = (int)blockIdx.x >= a.n_full && a.parts > 1;
  if (partial) {
    const int u = (int)blockIdx.x - a.n_full;
    rt = a.n_full + u / a.parts; part = u % a.parts;
    j0 = min(part * a.tpp, ntiles_all);
    ntiles = min(ntiles_all, j0 + a.tpp) - j0;
  }
  const int64_t t0 = (int64_t)rt * FT_TM;

  if (tid == 0) {
    for (int s = 0; s < FT_STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 8); }
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();               // barriers initialised, TMEM allocated: the B loader may start right away
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- A tile (warps 0-8; warp 9 is already streaming B images): one image row per (warp, iteration), lane = 4
  //      consecutive k; all global loads are issued before the first use, row maximum by shuffle, scaled fp16 hi/lo
  if (warp < NSTAGE_WARPS) {
    constexpr int ROWS_PER_WARP = (FT_TM + NSTAGE_WARPS - 1) / NSTAGE_WARPS;
    const int k = 4 * lane;
    float4 v[ROWS_PER_WARP];
#pragma unroll
    for (int u = 0; u < ROWS_PER_WARP; ++u) {
      const int t = warp + NSTAGE_WARPS * u;
      const int64_t tg = t0 + t;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < FT_TM && tg < a.B && k < Ds) v[u] = __ldg(reinterpret_cast<const float4*>(a.e_near + tg * Ds + k));
    }
#pragma unroll
    for (int u = 0; u < ROWS_PER_WARP; ++u) {
      const int t = warp + NSTAGE_WARPS * u;
      if (t >= FT_TM) continue;       // warp-uniform
      float mx = fmaxf(fmaxf(fabsf(v[u].x), fabsf(v[u].y)), fmaxf(fabsf(v[u].z), fabsf(v[u].w)));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const int ex = (mx > 0.f && mx < 3.0e38f) ? exponent_of(mx) - FH_TROW : 0;
      if (lane == 0) rexp[t] = ex;
      if (k < nchunks * FH_KCH) {
        const float x0 = scale_pow2(v[u].x, -ex), x1 = scale_pow2(v[u].y, -ex), x2 = scale_pow2(v[u].z, -ex), x3 = scale_pow2(v[u].w, -ex);
        const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(x0 - f01.x, x1 - f01.y), l23 = __floats2half2_rn(x2 - f23.x, x3 - f23.y);
        const uint32_t off = (uint32_t)((k >> 6) * 16384) + k128_chunk_off(t, (k & 63) >> 3) + (uint32_t)((lane & 1) * 8);
        uint2 hv, lv;
        hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
        lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
        *reinterpret_cast<uint2*>(a_t + off) = hv;
        *reinterpret_cast<uint2*>(a_t + 2 * 16384 + off) = lv;
      }
    }
    fence_async_proxy();
    // the MMA issuer waits for all staging warps on a named barrier (the loader does not take part)
    asm volatile("bar.sync 1, %0;" ::"n"(NSTAGE_WARPS * 32) : "memory");
  }

  if (warp == 9) {
    if (lane == 0) {
      int it = 0;
      for (int j = 0; j < ntiles; ++j) {
        for (int q = 0; q < nchunks; ++q, ++it) {
          const int s = it % FT_STAGES, ph = (it / FT_STAGES) & 1;
          if (it >= FT_STAGES) mbar_wait(&b_empty[s], (uint32_t)(ph ^ 1));
          mbar_expect_tx(&b_full[s], STAGE_BYTES);
          bulk_g2s(bst + s * STAGE_BYTES, a.bimg + ((int64_t)(j0 + j) * a.nchunks_max + q) * STAGE_BYTES, STAGE_BYTES, &b_full[s]);
        }
      }
    }
  } else if (warp == 8) {
    {
      // kind::f16, fp16 x fp16 -> fp32, both K-major, N = NT, M = 128
      const bool issuer = elect_one();
      const uint32_t idesc = (1u << 4) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(FT_TM >> 4) << 24);
      const uint64_t adesc0 = make_desc(smem_u32(a_t), 16, 1024, 2), bdesc0 = make_desc(smem_u32(bst), 16, 1024, 2);
      int it = 0;
      for (int j = 0; j < ntiles; ++j) {
        const int b = j & 1;
        if (j >= 2) mbar_wait(&acc_empty[b], (uint32_t)(((j >> 1) - 1) & 1));
        tc_fence_after();
        const uint32_t d = tmem_base + b * NT;
        for (int q = 0; q < nchunks; ++q, ++it) {
          const int s = it % FT_STAGES, ph = (it / FT_STAGES) & 1;
          mbar_wait(&b_full[s], (uint32_t)ph);
          tc_fence_after();
          // descriptor addresses in 16-byte units: +2 per K=16 step, +1024 per 16 KB atom of A, lo halves 2 atoms / one image on
          const uint64_t ahi = adesc0 + (uint64_t)(q * 1024), alo = ahi + 2048;
          const uint64_t bhi = bdesc0 + (uint64_t)(s * (STAGE_BYTES >> 4)), blo = bhi + (IMG_BYTES >> 4);
          const int ksteps = min(4, (dn - q * FH_KCH + 15) / 16);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (ks < ksteps && issuer) {
              umma_f16(d, ahi + 2 * ks, bhi + 2 * ks, idesc, (q > 0 || ks > 0) ? 1u : 0u);
              umma_f16(d, alo + 2 * ks, bhi + 2 * ks, idesc, 1u);
              umma_f16(d, ahi + 2 * ks, blo + 2 * ks, idesc, 1u);
            }
          }
          if (issuer) umma_commit(&b_empty[s]);
        }
        if (issuer) umma_commit(&acc_full[b]);
      }
    }
  } else {
    // ===== epilogue: warp = (column half, TMEM lane quadrant); one image row per thread, half of the tile's columns.
    //       One warp per scheduler was latency bound (2400 cycles per tile against 1920 of MMAs); two halve it.
    const int quad = warp & 3, half = warp >> 2;
    const int row = quad * 32 + lane;
    const int64_t tg = t0 + row;
    const bool live = tg < a.B;
    float pp[4] = {0.f, 0.f, 0.f, 0.f};
    if (live) {
      const float2 p1 = __ldg(reinterpret_cast<const float2*>(a.phi_near) + tg);
      if (ND == 2) {
        const float2 p2 = __ldg(reinterpret_cast<const float2*>(a.phi_far) + tg);
        pp[0] = p1.x * p2.x; pp[1] = p1.x * p2.y; pp[2] = p1.y * p2.x; pp[3] = p1.y * p2.y;   // index m*2+n
      } else {
        pp[0] = p1.x; pp[1] = p1.y;                                                             // index m
      }
    }
    float facc[L];
#pragma unroll
    for (int l = 0; l < L; ++l) facc[l] = 0.f;
    // far-environment entries of my row, fetched two tiles ahead: the load latency (~1 us) must not sit between the
    // accumulator becoming ready and its drain
    struct EK { float4 v[KPT / 4]; };
    auto load_e = [&](int j) {
      EK e;
#pragma unroll
      for (int u = 0; u < KPT / 4; ++u) {
        e.v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live && j < ntiles && KPT * (j0 + j) + 4 * u < Ds) e.v[u] = __ldg(reinterpret_cast<const float4*>(a.e_far + tg * Ds + KPT * (j0 + j) + 4 * u));
      }
      return e;
    };
    EK e_cur = load_e(0), e_nxt = load_e(1);
    for (int j = 0; j < ntiles; ++j) {
      const int b = j & 1;
      const EK e_nn = load_e(j + 2);
      float ek[KPT];
#pragma unroll
      for (int u = 0; u < KPT / 4; ++u) { ek[4 * u] = e_cur.v[u].x; ek[4 * u + 1] = e_cur.v[u].y; ek[4 * u + 2] = e_cur.v[u].z; ek[4 * u + 3] = e_cur.v[u].w; }
      e_cur = e_nxt; e_nxt = e_nn;
      mbar_wait(&acc_full[b], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * NT);
      if (half == 0) fh_epilogue_tile<L, 0, ND>(taddr, pp, ek, facc);
      else fh_epilogue_tile<L, 1, ND>(taddr, pp, ek, facc);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[b]);
    }
    // combine the two column halves, undo both scalings (exact powers of two)
    if (half == 1) {
#pragma unroll
      for (int l = 0; l < L; ++l) xch[row * L + l] = facc[l];
    }
    asm volatile("bar.sync 2, 256;" ::: "memory");
    if (half == 0) {
      const int ex = rexp[row] + bond_exponent(a.maxbits);
      if (!partial) {
        if (live) {
#pragma unroll
          for (int l = 0; l < L; ++l) a.f[tg * L + l] = scale_pow2(facc[l] + xch[row * L + l], ex);
        }
      } else {
        const int rloc = rt - a.n_full;
        float* mine = a.pscr + ((size_t)(rloc * a.parts + part) * FT_TM + row) * L;
        if (live) {
#pragma unroll
          for (int l = 0; l < L; ++l) __stcg(&mine[l], scale_pow2(facc[l] + xch[row * L + l], ex));
        }
        __threadfence();
        asm volatile("bar.sync 3, 128;" ::: "memory");
        if (tid == 0) {
          const int old = atomicAdd(&a.pcnt[rloc], 1);
          const int last = old == a.parts - 1;
          if (last) a.pcnt[rloc] = 0;               // ready for the next launch
          *s_last = last;
        }
        asm volatile("bar.sync 3, 128;" ::: "memory");
        if (*s_last && live) {
          __threadfence();
          float sum[L];
#pragma unroll
          for (int l = 0; l < L; ++l) sum[l] = 0.f;
          for (int pq = 0; pq < a.parts; ++pq) {           // fixed order, whoever arrives last
            const float* src = a.pscr + ((size_t)(rloc * a.parts + pq) * FT_TM + row) * L;
#pragma unroll
            for (int l = 0; l < L; ++l) sum[l] += __ldcg(&src[l]);
          }
#pragma unroll
          for (int l = 0; l < L; ++l) a.f[tg * L + l] = sum[l];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---- the same contraction by CTA PAIRS (tcgen05 cta_group::2): two CTAs of a cluster hold two image tiles (M = 256) and HALF
// of every B image each (80 of the 160 columns of a tile, 20 KB per stage instead of 40), the leader issues one MMA for
// both.  Per SM the shared-memory operand fetch drops from 115 to 81 B/clk and the TMA fill from 42 to 21 B/clk (the
// single-CTA kernel is bound by their sum against 128 B/clk and by a 3-stage ring that barely covers the L2 latency: the
// same shared memory now holds 6 stages).  Hand-shakes: each CTA's loader fills its own half and signals its own b_full;
// the peer's idle MMA warp relays its b_full to the leader's peer_full (a bulk copy cannot signal another CTA's mbarrier);
// tcgen05.commit multicasts b_empty / acc_full to both CTAs; the peer's epilogue warps arrive on the LEADER's acc_empty.
// Accumulators: each CTA's TMEM holds its own 128 rows, so the epilogue is unchanged.
constexpr int FH2_STAGES = 6;
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {       // acquire at cluster scope (remote arrivals)
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {          // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int L, int ND>
__global__ void __launch_bounds__(FH_THREADS, 1) forward_h2_kernel(FwdHArgs a) {
  constexpr int NT = 16 * L, KPT = 8 / ND;             // KPT: far-bond indices per tile
  constexpr int IMG_BYTES = NT * 128;                // one (hi | lo) B image: NT rows x 64 fp16
  constexpr int HALF_BYTES = IMG_BYTES / 2;          // this CTA's 80 rows of one (hi | lo) image
  constexpr int STAGE_BYTES = 2 * HALF_BYTES, STAGE_BYTES_G = 2 * IMG_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * NT <= 64 ? 64 : 2 * NT <= 128 ? 128 : 2 * NT <= 256 ? 256 : 512;
  constexpr int NSTAGE_WARPS = 9;                    // everyone but the loader stages the A tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_t = smem;
  uint8_t* bst = smem + FH_A_BYTES;
  constexpr int NS = FH2_STAGES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bst + NS * STAGE_BYTES);
  uint64_t* b_full = bars, *b_empty = bars + NS, *peer_full = bars + 2 * NS, *acc_full = bars + 3 * NS, *acc_empty = bars + 3 * NS + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * NS + 6);
  int* rexp = reinterpret_cast<int*>(bars + 3 * NS + 8);      // [128] row exponents
  float* xch = reinterpret_cast<float*>(rexp + FT_TM);               // [128][L] partial logits of the second column half
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Ds = a.Ds;
  int* s_last = reinterpret_cast<int*>(xch + FT_TM * L);
  const int dn = __ldg(a.d_near), df = __ldg(a.d_far);
  const int nchunks = (dn + FH_KCH - 1) / FH_KCH;
  const int ntiles_all = (df + KPT - 1) / KPT;
  // image tile and column-tile range [j0, j0 + ntiles) of this CTA
  const int crank = (int)cluster_ctarank();
  const bool leader = crank == 0;
  const int pair = (int)blockIdx.x >> 1;
  int rt = 2 * pair + crank, part = 0, j0 = 0, ntiles = ntiles_all;
  const bool partial = 2 * pair >= a.n_full && a.parts > 1;
  if (partial) {
    const int u = pair - a.n_full / 2;
    rt = a.n_full + 2 * (u / a.parts) + crank; part = u % a.parts;
    j0 = min(part * a.tpp, ntiles_all);
    ntiles = min(ntiles_all, j0 + a.tpp) - j0;
  }
  const int64_t t0 = (int64_t)rt * FT_TM;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); mbar_init(&peer_full[s], 1); }
    for (int b = 0; b < 3; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 16); }
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc2(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();               // barriers initialised, TMEM allocated: the B loader may start right away
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- A tile (warps 0-8; warp 9 is already streaming B images): one image row per (warp, iteration), lane = 4
  //      consecutive k; all global loads are issued before the first use, row maximum by shuffle, scaled fp16 hi/lo
  if (warp < NSTAGE_WARPS) {
    constexpr int ROWS_PER_WARP = (FT_TM + NSTAGE_WARPS - 1) / NSTAGE_WARPS;
    const int k = 4 * lane;
    float4 v[ROWS_PER_WARP];
#pragma unroll
    for (int u = 0; u < ROWS_PER_WARP; ++u) {
      const int t = warp + NSTAGE_WARPS * u;
      const int64_t tg = t0 + t;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < FT_TM && tg < a.B && k < Ds) v[u] = __ldg(reinterpret_cast<const float4*>(a.e_near + tg * Ds + k));
    }
#pragma unroll
    for (int u = 0; u < ROWS_PER_WARP; ++u) {
      const int t = warp + NSTAGE_WARPS * u;
      if (t >= FT_TM) continue;       // warp-uniform
      float mx = fmaxf(fmaxf(fabsf(v[u].x), fabsf(v[u].y)), fmaxf(fabsf(v[u].z), fabsf(v[u].w)));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const int ex = (mx > 0.f && mx < 3.0e38f) ? exponent_of(mx) - FH_TROW : 0;
      if (lane == 0) rexp[t] = ex;
      if (k < nchunks * FH_KCH) {
        const float x0 = scale_pow2(v[u].x, -ex), x1 = scale_pow2(v[u].y, -ex), x2 = scale_pow2(v[u].z, -ex), x3 = scale_pow2(v[u].w, -ex);
        const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(x0 - f01.x, x1 - f01.y), l23 = __floats2half2_rn(x2 - f23.x, x3 - f23.y);
        const uint32_t off = (uint32_t)((k >> 6) * 16384) + k128_chunk_off(t, (k & 63) >> 3) + (uint32_t)((lane & 1) * 8);
        uint2 hv, lv;
        hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
        lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
        *reinterpret_cast<uint2*>(a_t + off) = hv;
        *reinterpret_cast<uint2*>(a_t + 2 * 16384 + off) = lv;
      }
    }
    fence_async_proxy();
  }
  // both A tiles staged, every barrier of the pair initialised: one cluster barrier (the loader has already started: its
  // first stages only touch this CTA's own barriers)
  if (warp == 9 && lane == 0) {
    const int total = ntiles * nchunks;
    for (int it = 0; it < NS && it < total; ++it) {
      const int j = it / nchunks, q = it % nchunks;
      mbar_expect_tx(&b_full[it], STAGE_BYTES);
      const uint8_t* src = a.bimg + ((int64_t)(j0 + j) * a.nchunks_max + q) * STAGE_BYTES_G + (size_t)crank * HALF_BYTES;
      bulk_g2s(bst + it * STAGE_BYTES, src, HALF_BYTES, &b_full[it]);
      bulk_g2s(bst + it * STAGE_BYTES + HALF_BYTES, src + IMG_BYTES, HALF_BYTES, &b_full[it]);
    }
  }
  __syncwarp();
  cluster_sync_all();
  tc_fence_after();

  if (warp == 9) {
    if (lane == 0) {
      int it = 0;
      for (int j = 0; j < ntiles; ++j) {
        for (int q = 0; q < nchunks; ++q, ++it) {
          if (it < NS) continue;                           // issued before the cluster barrier
          const int s = it % NS, ph = (it / NS) & 1;
          mbar_wait(&b_empty[s], (uint32_t)(ph ^ 1));
          mbar_expect_tx(&b_full[s], STAGE_BYTES);
          const uint8_t* src = a.bimg + ((int64_t)(j0 + j) * a.nchunks_max + q) * STAGE_BYTES_G + (size_t)crank * HALF_BYTES;
          bulk_g2s(bst + s * STAGE_BYTES, src, HALF_BYTES, &b_full[s]);
          bulk_g2s(bst + s * STAGE_BYTES + HALF_BYTES, src + IMG_BYTES, HALF_BYTES, &b_full[s]);
        }
      }
    }
  } else if (warp == 8) {
    if (lane == 0 && !leader) {
      // relay: this CTA's half of a stage has landed -> the leader's peer_full
      const int total = ntiles * nchunks;
      for (int it = 0; it < total; ++it) {
        const int s = it % NS, ph = (it / NS) & 1;
        mbar_wait(&b_full[s], (uint32_t)ph);
        mbar_arrive_remote(mapa_u32(smem_u32(&peer_full[s]), 0));
      }
    } else if (leader) {
      // kind::f16, fp16 x fp16 -> fp32, both K-major, N = NT, M = 256 over the CTA pair
      const bool issuer = elect_one();
      const uint32_t idesc = (1u << 4) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint64_t adesc0 = make_desc(smem_u32(a_t), 16, 1024, 2), bdesc0 = make_desc(smem_u32(bst), 16, 1024, 2);
      int it = 0;
      for (int j = 0; j < ntiles; ++j) {
        const int b = j % 3;                 // three accumulators (480 of the 512 TMEM columns): the peer's drain is a cluster hop away
        if (j >= 3) mbar_wait_cluster(&acc_empty[b], (uint32_t)((j / 3 - 1) & 1));
        tc_fence_after();
        const uint32_t d = tmem_base + b * NT;
        for (int q = 0; q < nchunks; ++q, ++it) {
          const int s = it % NS, ph = (it / NS) & 1;
          mbar_wait(&b_full[s], (uint32_t)ph);
          mbar_wait_cluster(&peer_full[s], (uint32_t)ph);
          tc_fence_after();
          const uint64_t ahi = adesc0 + (uint64_t)(q * 1024), alo = ahi + 2048;
          const uint64_t bhi = bdesc0 + (uint64_t)(s * (STAGE_BYTES >> 4)), blo = bhi + (HALF_BYTES >> 4);
          const int ksteps = min(4, (dn - q * FH_KCH + 15) / 16);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (ks < ksteps && issuer) {
              umma2_f16(d, ahi + 2 * ks, bhi + 2 * ks, idesc, (q > 0 || ks > 0) ? 1u : 0u);
              umma2_f16(d, alo + 2 * ks, bhi + 2 * ks, idesc, 1u);
              umma2_f16(d, ahi + 2 * ks, blo + 2 * ks, idesc, 1u);
            }
          }
          if (issuer) umma2_commit(&b_empty[s]);
        }
        if (issuer) umma2_commit(&acc_full[b]);
      }
    }
  } else {
    // ===== epilogue: warp = (column half, TMEM lane quadrant); one image row per thread, half of the tile's columns.
    //       One warp per scheduler was latency bound (2400 cycles per tile against 1920 of MMAs); two halve it.
    const int quad = warp & 3, half = warp >> 2;
    const int row = quad * 32 + lane;
    const int64_t tg = t0 + row;
    const bool live = tg < a.B;
    float pp[4] = {0.f, 0.f, 0.f, 0.f};
    if (live) {
      const float2 p1 = __ldg(reinterpret_cast<const float2*>(a.phi_near) + tg);
      if (ND == 2) {
        const float2 p2 = __ldg(reinterpret_cast<const float2*>(a.phi_far) + tg);
        pp[0] = p1.x * p2.x; pp[1] = p1.x * p2.y; pp[2] = p1.y * p2.x; pp[3] = p1.y * p2.y;   // index m*2+n
      } else {
        pp[0] = p1.x; pp[1] = p1.y;                                                             // index m
      }
    }
    float facc[L];
#pragma unroll
    for (int l = 0; l < L; ++l) facc[l] = 0.f;
    // far-environment entries of my row, fetched two tiles ahead: the load latency (~1 us) must not sit between the
    // accumulator becoming ready and its drain
    struct EK { float4 v[KPT / 4]; };
    auto load_e = [&](int j) {
      EK e;
#pragma unroll
      for (int u = 0; u < KPT / 4; ++u) {
        e.v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live && j < ntiles && KPT * (j0 + j) + 4 * u < Ds) e.v[u] = __ldg(reinterpret_cast<const float4*>(a.e_far + tg * Ds + KPT * (j0 + j) + 4 * u));
      }
      return e;
    };
    EK e_cur = load_e(0), e_nxt = load_e(1);
    for (int j = 0; j < ntiles; ++j) {
      const int b = j % 3;
      const EK e_nn = load_e(j + 2);
      float ek[KPT];
#pragma unroll
      for (int u = 0; u < KPT / 4; ++u) { ek[4 * u] = e_cur.v[u].x; ek[4 * u + 1] = e_cur.v[u].y; ek[4 * u + 2] = e_cur.v[u].z; ek[4 * u + 3] = e_cur.v[u].w; }
      e_cur = e_nxt; e_nxt = e_nn;
      mbar_wait(&acc_full[b], (uint32_t)((j / 3) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * NT);
      if (half == 0) fh_epilogue_tile<L, 0, ND>(taddr, pp, ek, facc);
      else fh_epilogue_tile<L, 1, ND>(taddr, pp, ek, facc);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (leader) mbar_arrive(&acc_empty[b]); else mbar_arrive_remote(mapa_u32(smem_u32(&acc_empty[b]), 0)); }
    }
    // combine the two column halves, undo both scalings (exact powers of two)
    if (half == 1) {
#pragma unroll
      for (int l = 0; l < L; ++l) xch[row * L + l] = facc[l];
    }
    asm volatile("bar.sync 2, 256;" ::: "memory");
    if (half == 0) {
      const int ex = rexp[row] + bond_exponent(a.maxbits);
      if (!partial) {
        if (live) {
#pragma unroll
          for (int l = 0; l < L; ++l) a.f[tg * L + l] = scale_pow2(facc[l] + xch[row * L + l], ex);
        }
      } else {
        const int rloc = rt - a.n_full;
        float* mine = a.pscr + ((size_t)(rloc * a.parts + part) * FT_TM + row) * L;
        if (live) {
#pragma unroll
          for (int l = 0; l < L; ++l) __stcg(&mine[l], scale_pow2(facc[l] + xch[row * L + l], ex));
        }
        __threadfence();
        asm volatile("bar.sync 3, 128;" ::: "memory");
        if (tid == 0) {
          const int old = atomicAdd(&a.pcnt[rloc], 1);
          const int last = old == a.parts - 1;
          if (last) a.pcnt[rloc] = 0;               // ready for the next launch
          *s_last = last;
        }
        asm volatile("bar.sync 3, 128;" ::: "memory");
        if (*s_last && live) {
          __threadfence();
          float sum[L];
#pragma unroll
          for (int l = 0; l < L; ++l) sum[l] = 0.f;
          for (int pq = 0; pq < a.parts; ++pq) {           // fixed order, whoever arrives last
            const float* src = a.pscr + ((size_t)(rloc * a.parts + pq) * FT_TM + row) * L;
#pragma unroll
            for (int l = 0; l < L; ++l) sum[l] += __ldcg(&src[l]);
          }
#pragma unroll
          for (int l = 0; l < L; ++l) a.f[tg * L + l] = sum[l];
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tmem_base, TMEM_COLS);
}

template <int L>
size_t forward_h_smem() { return (size_t)FH_A_BYTES + (size_t)FT_STAGES * 2 * 16 * L * 128 + 1024 + (size_t)FT_TM * L * 4 + 1024; }

template <int L>
size_t forward_h2_smem() { return (size_t)FH_A_BYTES + (size_t)FH2_STAGES * 16 * L * 128 + 1024 + (size_t)FT_TM * L * 4 + 1024; }

template <int L>
size_t forward_tc_smem() { return (size_t)2 * FT_A_BYTES + (size_t)FT_STAGES * 2 * 16 * 16 * L * 4 + 256 + 1024; }

}  // namespace

bool forward_tc_supported(int Ds, int L, int nd) { return nd == 2 && L == 10 && Ds <= FT_KMAX && Ds % 8 == 0; }

// floats needed for the hi/lo B images of one bond
int64_t forward_tc_image_floats(int Ds, int L) {
  const int64_t ntiles = (Ds + 3) / 4, nchunks = (Ds + FT_KCH - 1) / FT_KCH;
  return ntiles * nchunks * 2 * 16 * 16 * L + 64 + FH_PCNT + (int64_t)FH_PARTS_MAX * FH_TAIL_MAX * FT_TM * L;
  // (the fp16 images need half of the first term; + the bond exponent word, the arrival counters and the partial logits of
  //  the column-split tail tiles)
}

// how the image tiles of a launch are spread over the SMs: full waves of whole tiles, then the remaining `rem` tiles split
// by columns over parts = floor(sms / rem) CTAs each (at most FH_PARTS_MAX and one column tile per CTA)
// pairs: the kernel runs as clusters of two CTAs (two image tiles per cluster): whole waves of sms / 2 pairs, the tail split per pair
static void forward_h_tail(int64_t B, int ntiles_cols, FwdHArgs& h, unsigned& grid, uint8_t* img_end, bool pairs = false) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  if (sms > FH_TAIL_MAX) sms = FH_TAIL_MAX;
  if (pairs) sms &= ~1;
  const int64_t rtiles = ceil_div64(B, FT_TM);
  const int rem = (int)(rtiles % sms);
  const int rem_units = pairs ? (rem + 1) / 2 : rem, units = pairs ? sms / 2 : sms;      // CTAs or CTA pairs
  int parts = rem_units > 0 ? units / rem_units : 1;
  if (parts > FH_PARTS_MAX) parts = FH_PARTS_MAX;
  if (parts > ntiles_cols) parts = ntiles_cols;
  if (parts < 1) parts = 1;
  h.n_full = (int)(rtiles - rem);
  h.parts = parts;
  h.tpp = (ntiles_cols + parts - 1) / parts;
  h.pcnt = reinterpret_cast<int32_t*>(img_end) + 16;                 // right behind the bond exponent word (cleared with it)
  h.pscr = reinterpret_cast<float*>(img_end) + 16 + FH_PCNT;
  grid = pairs ? (unsigned)(h.n_full + 2 * (int64_t)rem_units * parts) : (unsigned)(h.n_full + (int64_t)rem * parts);
}

// launch of the CTA-pair kernel (clusters of 2); false if this device cannot schedule it (the caller then takes the single-CTA kernel)
template <int ND>
static int launch_forward_h2(FwdHArgs& h, int64_t B, int ntiles_cols, uint8_t* tail_base, cudaStream_t st, bool& done) {
  done = false;
  static int ok = -1;
  const size_t smem = forward_h2_smem<10>();
  if (ok < 0) {
    ok = cudaFuncSetAttribute(forward_h2_kernel<10, ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess ? 1 : 0;
    (void)cudaGetLastError();
  }
  if (!ok) return TRMPS_OK;
  unsigned grid = 0;
  forward_h_tail(B, ntiles_cols, h, grid, tail_base, true);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(FH_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  TRMPS_CUDA_OK(cudaLaunchKernelEx(&cfg, forward_h2_kernel<10, ND>, h));
  count_launch();
  done = true;
  return TRMPS_OK;
}

// where the fp16 forward keeps the bits of the largest |bond entry| (two-site form): a kernel that produces the bond may fill
// it with an integer atomicMax and pass have_max = true to launch_forward_tc (null when the fp16 forward is not in use)
int32_t* forward_tc_maxbits(float* bimg, int Ds, int L) {
  if (option_value(1) != 0 || !forward_tc_supported(Ds, L, 2)) return nullptr;
  const int ntiles = (Ds + 3) / 4, nchunks = (Ds + FH_KCH - 1) / FH_KCH;
  return reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(bimg) + (size_t)ntiles * nchunks * 2 * 16 * L * 128);
}

int launch_forward_tc(const float* e_near, const float* e_far, const float* phi_near, const float* phi_far, const float* M,
                      float* bimg, float* f_out, int64_t B, int Ds, int L, const int32_t* d_near, const int32_t* d_far,
                      cudaStream_t st, bool have_max) {
  if (B <= 0) return TRMPS_OK;
  if (!forward_tc_supported(Ds, L, 2)) {
    set_error("forward_tc: unsupported shape Ds=%d L=%d", Ds, L);
    return TRMPS_E_ARG;
  }
  if (option_value(1) == 0) {
    // fp16 hi/lo operands (default)
    const int ntiles = (Ds + 3) / 4, nchunks = (Ds + FH_KCH - 1) / FH_KCH;
    uint8_t* img = reinterpret_cast<uint8_t*>(bimg);
    int32_t* maxbits = reinterpret_cast<int32_t*>(img + (size_t)ntiles * nchunks * 2 * 16 * L * 128);
    if (!have_max) {
      TRMPS_CUDA_OK(cudaMemsetAsync(maxbits, 0, sizeof(int32_t) * (16 + FH_PCNT), st));
      absmax_kernel<<<148, 256, 0, st>>>(M, (int64_t)2 * Ds * 2 * L * Ds, maxbits);
      TRMPS_LAUNCH_OK("absmax_kernel");
    }
    forward_prep_h_kernel<10><<<592, 256, 0, st>>>(M, img, maxbits, Ds, ntiles, nchunks);
    TRMPS_LAUNCH_OK("forward_prep_h_kernel");
    FwdHArgs h;
    h.e_near = e_near; h.e_far = e_far; h.phi_near = phi_near; h.phi_far = phi_far; h.bimg = img; h.maxbits = maxbits;
    h.f = f_out; h.B = B; h.Ds = Ds; h.d_near = d_near; h.d_far = d_far; h.nchunks_max = nchunks;
    const size_t smem_h = forward_h_smem<10>();
    static bool attr_h = false;
    if (!attr_h) {
      TRMPS_CUDA_OK(cudaFuncSetAttribute(forward_h_kernel<10, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_h));
      attr_h = true;
    }
    if (option_value(6) != 0) {
      bool done = false;
      int rc = launch_forward_h2<2>(h, B, ntiles, reinterpret_cast<uint8_t*>(maxbits), st, done);
      if (rc || done) return rc;
    }
    unsigned grid = 0;
    forward_h_tail(B, ntiles, h, grid, reinterpret_cast<uint8_t*>(maxbits));
    forward_h_kernel<10, 2><<<grid, FH_THREADS, smem_h, st>>>(h);
    TRMPS_LAUNCH_OK("forward_h_kernel");
    return TRMPS_OK;
  }
  const int ntiles = (Ds + 3) / 4, nchunks = (Ds + FT_KCH - 1) / FT_KCH;
  forward_prep_kernel<10><<<592, 256, 0, st>>>(M, bimg, Ds, ntiles, nchunks);
  TRMPS_LAUNCH_OK("forward_prep_kernel");
  FwdTcArgs a;
  a.e_near = e_near; a.e_far = e_far; a.phi_near = phi_near; a.phi_far = phi_far; a.bimg = bimg; a.f = f_out;
  a.B = B; a.Ds = Ds; a.d_near = d_near; a.d_far = d_far; a.nchunks_max = nchunks;
  const size_t smem = forward_tc_smem<10>();
  static bool attr_set = false;
  if (!attr_set) {
    TRMPS_CUDA_OK(cudaFuncSetAttribute(forward_tc_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  forward_tc_kernel<10><<<(unsigned)ceil_div64(B, FT_TM), FT_THREADS, smem, st>>>(a);
  TRMPS_LAUNCH_OK("forward_tc_kernel");
  return TRMPS_OK;
}

// single-site form of the first forward of a bond update (see forward_prep_site_kernel): special = the special node at the near
// site, e_far = the cached environment on the other side of the same site, d_far = the bond between the two sites
bool forward_site_tc_supported(int Ds, int L) { return option_value(1) == 0 && forward_tc_supported(Ds, L, 2); }

int launch_forward_site_tc(const float* e_near, const float* e_far, const float* phi_near, const float* special, int dir, float* bimg,
                           float* f_out, int64_t B, int Ds, int L, const int32_t* d_near, const int32_t* d_far, cudaStream_t st) {
  if (B <= 0) return TRMPS_OK;
  if (!forward_site_tc_supported(Ds, L)) {
    set_error("forward_site_tc: unsupported shape Ds=%d L=%d", Ds, L);
    return TRMPS_E_ARG;
  }
  const int ntiles = (Ds + 7) / 8, nchunks = (Ds + FH_KCH - 1) / FH_KCH;
  uint8_t* img = reinterpret_cast<uint8_t*>(bimg);
  int32_t* maxbits = reinterpret_cast<int32_t*>(img + (size_t)ntiles * nchunks * 2 * 16 * L * 128);
  TRMPS_CUDA_OK(cudaMemsetAsync(maxbits, 0, sizeof(int32_t), st));
  absmax_kernel<<<148, 256, 0, st>>>(special, (int64_t)2 * L * Ds * Ds, maxbits);
  TRMPS_LAUNCH_OK("absmax_kernel");
  forward_prep_site_kernel<10><<<296, 256, 0, st>>>(special, img, maxbits, Ds, dir, ntiles, nchunks);
  TRMPS_LAUNCH_OK("forward_prep_site_kernel");
  FwdHArgs h;
  h.e_near = e_near; h.e_far = e_far; h.phi_near = phi_near; h.phi_far = phi_near; h.bimg = img; h.maxbits = maxbits;
  h.f = f_out; h.B = B; h.Ds = Ds; h.d_near = d_near; h.d_far = d_far; h.nchunks_max = nchunks;
  const size_t smem_h = forward_h_smem<10>();
  static bool attr_h = false;
  if (!attr_h) {
    TRMPS_CUDA_OK(cudaFuncSetAttribute(forward_h_kernel<10, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_h));
    attr_h = true;
  }
  // (the arrival counters and the partial logits of the tail tiles sit behind the TWO-site image, whichever form runs)
  uint8_t* tail_base = img + (size_t)((Ds + 3) / 4) * nchunks * 2 * 16 * L * 128;
  if (option_value(6) != 0) {
    bool done = false;
    int rc = launch_forward_h2<1>(h, B, ntiles, tail_base, st, done);
    if (rc || done) return rc;
  }
  unsigned grid = 0;
  forward_h_tail(B, ntiles, h, grid, tail_base);
  forward_h_kernel<10, 1><<<grid, FH_THREADS, smem_h, st>>>(h);
  TRMPS_LAUNCH_OK("forward_h_kernel");
  return TRMPS_OK;
}

}  // namespace trmps
