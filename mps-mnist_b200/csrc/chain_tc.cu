// Fused on-chip environment chain on the 5th-generation tensor cores (north_star pieces 1 + 2, inference path).
// Replaces the loops of MPS.predict (mps.py:536-549: tf.while_loop over _chain_multiply_r / _chain_multiply_l),
// which the reference runs one site at a time with a (B x D x D) intermediate in memory.
//
// One CTA carries TILES tiles of 128 images through ALL sites of a chain without touching HBM except for the
// pixel stream: per site and tile
//     X[t,(m,j)] = sum_i S[t,i] * node[m,i,j]        tcgen05.mma kind::f16, M=128, K=KP, N=2*KP, accumulator in TMEM
//     S'[t,j]    = phi_0(x_t) X[t,(0,j)] + phi_1(x_t) X[t,(1,j)]      epilogue in registers (feature map fused)
// and S' goes straight back into the K-major SWIZZLE_128B shared-memory tile that is the A operand of the next site.
//
// FP32 emulation on the fp16 pipe (twice the TF32 rate).  Both operands are carried as hi + lo halves,
//     x = 2^e (h + l),  h = fp16(x 2^-e),  l = fp16(x 2^-e - h),
// and the product keeps  h.h + l.h + h.l  (the dropped l.l term is 2^-22 relative, the representation error of
// h + l is 2^-23): the same accuracy as 3xTF32.  fp16's narrow exponent is handled by exact power-of-two scaling:
// every image row of S carries its own integer exponent, every node matrix one exponent for the site (maximum in
// [2^13, 2^14)).  The row exponent of a step is PREDICTED from the growth seen in the previous step so that the
// epilogue converts in a single pass over TMEM; the realised row maximum is checked against a safe window
// (2^-2 .. 2^15: no overflow, error floor below 2^-23 of the row maximum) and the step is redone with the exact
// exponent in the rare case it falls outside; the exponents are added back in fp32 only when the final environment is written.
//
// Order of accumulation.  The tensor core adds every MMA instruction's 16-term sum into the fp32 accumulator with
// round-toward-zero, i.e. each instruction issued while the accumulator already has its final magnitude shrinks it by
// half an ulp on average.  In a single contraction that is invisible (a few 1e-8); in a chain it compounds as a common
// factor over the sites (measured: 1.4e-4 relative on the logits after 784 sites at KP=32 and after 196 sites at KP=128
// with the main terms first, tests/test_gpu_fullsize.py).  The small products (l.h, h.l: 2^-11 of the result) are
// therefore issued FIRST, while the accumulator is still small, and the h.h products last: KP/16 instructions round at
// full magnitude instead of 3 KP/16.
//
// Shapes: KP = 32 / 64 / 128 (padded bond), TILES = 4 / 4 / 2 image tiles per CTA (TMEM columns 2*KP*TILES <= 512).
// The tiles are staggered: while the tensor core works on one tile, the epilogue warpgroups of the others drain
// TMEM and rebuild their A tiles.  Node images (pre-split fp16 hi/lo, built once per call, L2 resident) are
// streamed in 128-byte-row pieces by one warp with cp.async.bulk through an mbarrier ring.
//
// Warp roles: warpgroup k = epilogue of tile k (one image row per thread), then one MMA-issuer warp and one
// node-loader warp.
#include <cuda_fp16.h>
#include <cstdlib>
#include "common.cuh"

namespace trmps {

namespace {

constexpr int CH_TM = 128;
constexpr int CH_TEXP = 13;                  // node maxima are scaled into [2^13, 2^14)
constexpr int CH_TROW = 8;                   // image rows aim at a maximum in [2^8, 2^9); accepted window 2^-2 .. 2^15

template <int KP> struct ChainCfg;
// TILES image tiles per CTA, WG epilogue warpgroups per tile (each owns KP/WG state columns of every row), NSTAGE
// ring stages of one piece each, PIECES pieces per site.  TILE_MAJOR: a tile runs through all pieces of a site
// before the next tile starts (needs the ring to hold a whole site); otherwise the tiles alternate on each piece.
template <> struct ChainCfg<32> {
  static constexpr int TILES = 4, WG = 1, NSTAGE = 4, PIECES = 1;
  static constexpr bool TILE_MAJOR = true;
  static constexpr int A_BYTES = 16384;                 // one atom column: row = [hi(32) | lo(32)] fp16
  static constexpr int PIECE_BYTES = 64 * 128;          // 64 rows (m,j), row = [hi(k 0..31) | lo(k 0..31)]
};
template <> struct ChainCfg<64> {
  static constexpr int TILES = 4, WG = 1, NSTAGE = 4, PIECES = 2;  // pieces: lo, hi
  static constexpr bool TILE_MAJOR = true;
  static constexpr int A_BYTES = 2 * 16384;             // [hi | lo] x one 64-wide k atom
  static constexpr int PIECE_BYTES = 128 * 128;
};
template <> struct ChainCfg<128> {
  // N = 256 per MMA: M=128 MMAs with smaller N are bound by the shared-memory operand fetch (8 KB per 64 cycles),
  // at N = 256 the tensor pipe itself is the limit.  Two warpgroups per tile halve the epilogue latency.
  static constexpr int TILES = 2, WG = 2, NSTAGE = 3, PIECES = 4;  // pieces: (ka0,lo) (ka1,lo) (ka0,hi) (ka1,hi)
  static constexpr bool TILE_MAJOR = false;
  static constexpr int A_BYTES = 4 * 16384;             // [hi ka0 | hi ka1 | lo ka0 | lo ka1]
  static constexpr int PIECE_BYTES = 256 * 128;
};
template <int KP> constexpr int chain_threads() { return ChainCfg<KP>::TILES * ChainCfg<KP>::WG * 128 + 64; }
constexpr int CH_AUX_BYTES = 256 + 2048;     // barriers + TMEM slot + clock base, then the row-maximum exchange
constexpr int CH_SMEM_MAX = 232448;          // 227 KB opt-in limit per CTA
template <int KP> __host__ __device__ constexpr int chain_smem() {
  // the operand tiles need 1024-byte alignment; ask for as much slack as fits (the dynamic window starts 1024-aligned
  // when the kernel has no static shared memory, which the kernel checks)
  constexpr int need = ChainCfg<KP>::TILES * ChainCfg<KP>::A_BYTES + ChainCfg<KP>::NSTAGE * ChainCfg<KP>::PIECE_BYTES + CH_AUX_BYTES;
  return need + (CH_SMEM_MAX - need < 1024 ? CH_SMEM_MAX - need : 1024);
}
template <int KP> constexpr int site_bytes() { return ChainCfg<KP>::PIECES * ChainCfg<KP>::PIECE_BYTES; }

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
// blocking wait; the suspend-time hint lets the waiting warp sleep in hardware (it is woken when the phase completes)
// instead of spinning through issue slots that the epilogue warps need
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
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
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 consecutive accumulator columns of my TMEM lane (issue only; pair with tmem_ld_wait)
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B descriptor: 128-byte rows, 8-row atoms 1024 B apart
__device__ __forceinline__ uint64_t make_desc_k128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                    // LBO (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;          // SBO
  d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}
// byte offset of the 16-byte chunk `ch` (8 fp16 k-values) of row `r` inside a 128-byte-row K-major SWIZZLE_128B block
__host__ __device__ __forceinline__ uint32_t k128_chunk_off(int r, int ch) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + (((ch ^ r) & 7) << 4));
}
__device__ __forceinline__ float pow2f(int e) { return __int_as_float((e + 127) << 23); }     // -126 <= e <= 127
__device__ __forceinline__ int exponent_of(float x) { return (int)((__float_as_uint(x) >> 23) & 0xffu) - 127; }
// x * 2^e for any e (two exact steps keep the intermediate factors in range)
__device__ __forceinline__ float scale_pow2(float x, int e) {
  e = max(-250, min(250, e));
  const int e1 = e / 2;
  return x * pow2f(e1) * pow2f(e - e1);
}

// ---- node slots -> scaled fp16 hi/lo images in chain order.  step s uses site first + s*dir.
//   dir=+1: W[k=i][(m,j)] = node[m][i][j]     dir=-1: W[k=j][(m,i)] = node[m][i][j]
// exps[s] = e_s - TEXP where 2^e_s <= max|node_s| < 2^(e_s+1): the image holds node * 2^-exps[s].
__global__ void chain_exp_kernel(const float* __restrict__ nodes, int32_t* __restrict__ exps, int Ds, int first, int dir) {
  __shared__ float red[32];
  const int s = blockIdx.x;
  const float* node = nodes + (size_t)(first + s * dir) * 2 * Ds * Ds;
  float m = 0.f;
  // 2 Ds^2 floats, Ds a multiple of 4: float4 loads, eight in flight per thread (one L2 round trip instead of 28)
  const float4* node4 = reinterpret_cast<const float4*>(node);
  const int n4 = (2 * Ds * Ds) >> 2;
  for (int e0 = threadIdx.x; e0 < n4; e0 += 8 * blockDim.x) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = e0 + u * blockDim.x;
      v[u] = e < n4 ? __ldg(node4 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) m = fmaxf(m, fmaxf(fmaxf(fabsf(v[u].x), fabsf(v[u].y)), fmaxf(fabsf(v[u].z), fabsf(v[u].w))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
    exps[s] = (m > 0.f && m < 3.0e38f) ? exponent_of(m) - CH_TEXP : 0;
  }
}

template <int KP>
__global__ void chain_prep_kernel(const float* __restrict__ nodes, uint8_t* __restrict__ img, const int32_t* __restrict__ exps,
                                  int Ds, int first, int dir, int nsteps) {
  using C = ChainCfg<KP>;
  constexpr int PER = 2 * KP * (KP / 8);     // (row n, 8-wide k chunk) pairs per site
  const int64_t total = (int64_t)nsteps * PER;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int s = (int)(e / PER), r = (int)(e % PER);
    const int n = r / (KP / 8), kc = r % (KP / 8), m = n / KP, j = n % KP;
    const float* node = nodes + (size_t)(first + s * dir) * 2 * Ds * Ds;
    const int ex = -exps[s];
    __align__(16) __half hi[8];
    __align__(16) __half lo[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int k = kc * 8 + q;
      float v = 0.f;
      if (k < Ds && j < Ds) v = dir > 0 ? node[((size_t)m * Ds + k) * Ds + j] : node[((size_t)m * Ds + j) * Ds + k];
      v = scale_pow2(v, ex);
      hi[q] = __float2half_rn(v);
      lo[q] = __float2half_rn(v - __half2float(hi[q]));
    }
    uint8_t* base = img + (size_t)s * (C::PIECES * C::PIECE_BYTES);
    uint8_t *ph, *pl;
    if (KP == 32) {          // one piece, row = [hi | lo]
      ph = base + k128_chunk_off(n, kc);
      pl = base + k128_chunk_off(n, 4 + kc);
    } else {                 // pieces (lo, ka = 0..KA-1) then (hi, ka = 0..KA-1): the small products are issued first
      const int ka = kc >> 3, ch = kc & 7;
      ph = base + (size_t)(KP / 64 + ka) * C::PIECE_BYTES + k128_chunk_off(n, ch);
      pl = base + (size_t)ka * C::PIECE_BYTES + k128_chunk_off(n, ch);
    }
    *reinterpret_cast<uint4*>(ph) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(pl) = *reinterpret_cast<const uint4*>(lo);
  }
}

__device__ float g_chain_dbg[64];     // [0..7] phase counters, [16..63] event timestamps of the middle step (cycles since kernel start)

struct ChainArgs {
  const float* feat; int fm; int64_t B; int Ds, L; int first, dir, nsteps; int start_kind;   // start_kind 0: [0_L,1_L], 1: [1_L,0_L]
  const uint8_t* img; const int32_t* exps; float* out;      // out: (B, Ds) final environment
  const float* in;          // start_kind 2: the chain starts from this environment (B, Ds) in HBM instead of a boundary vector
  float* out_all;           // not null: the environment after step s (s < nsteps - 1) is also written to out_all + s * out_stride
  int64_t out_stride;       //           (training: the cached environments C1s / C2s, optimizer.py:49-87)
  int dbg;       // development: 1 = epilogue skips TMEM reads and A-tile stores, 2 = loader skips the copies (rate experiments; results invalid)
};

// DBG compiles in the clock64 phase timers / event stamps and the rate-experiment switches (TRMPS_CHAIN_DEBUG)
template <int KP, bool DBG>
__global__ void __launch_bounds__(chain_threads<KP>(), 1) chain_tc_kernel(ChainArgs a) {
  using C = ChainCfg<KP>;
  constexpr int TILES = C::TILES, WG = C::WG, NSTAGE = C::NSTAGE, PIECES = C::PIECES, NCOL = 2 * KP, JW = KP / WG;
  constexpr int EPI_WARPS = TILES * WG * 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_base = smem;                                   // [tile][A_BYTES]
  uint8_t* n_base = smem + TILES * C::A_BYTES;              // [stage][PIECE_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(n_base + NSTAGE * C::PIECE_BYTES);
  uint64_t* node_full = bars, *node_empty = bars + NSTAGE, *a_ready = bars + 2 * NSTAGE, *x_ready = bars + 2 * NSTAGE + TILES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 2 * TILES);
  long long& t_base = *reinterpret_cast<long long*>(bars + 2 * NSTAGE + 2 * TILES + 1);
  float* mx_x = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [tile][wg][row] partial row maxima
  static_assert((2 * NSTAGE + 2 * TILES + 2) * 8 <= 256 && TILES * WG * CH_TM * 4 <= 2048, "aux region");
  constexpr int SMEM_TOTAL = chain_smem<KP>();
  if (reinterpret_cast<uint8_t*>(bars) + CH_AUX_BYTES > smem_raw + SMEM_TOTAL) __trap();      // alignment slack exhausted
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nsteps = a.nsteps;

  if (tid == 0) {
    if (DBG) t_base = clock64();
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&node_full[s], 1); mbar_init(&node_empty[s], 1); }
    for (int k = 0; k < TILES; ++k) { mbar_init(&a_ready[k], WG * 128); mbar_init(&x_ready[k], 1); }
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, TILES * NCOL);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == EPI_WARPS + 1) {
    // ===== node loader: one bulk copy per piece =====
    if (lane == 0) {
      const int npieces = nsteps * PIECES;
      for (int p = 0; p < npieces; ++p) {
        const int st = p % NSTAGE;
        if (p >= NSTAGE) mbar_wait(&node_empty[st], (uint32_t)(((p / NSTAGE) - 1) & 1));
        if (DBG && a.dbg == 2) { mbar_arrive(&node_full[st]); continue; }      // experiment: no node traffic (results invalid)
        mbar_expect_tx(&node_full[st], C::PIECE_BYTES);
        bulk_g2s(n_base + st * C::PIECE_BYTES, a.img + (size_t)p * C::PIECE_BYTES, C::PIECE_BYTES, &node_full[st]);
      }
    }
  } else if (warp == EPI_WARPS) {
    // ===== MMA issuer: the whole warp runs the schedule (uniform control flow and descriptor arithmetic stay in the uniform
    //       datapath), one elected lane issues the MMAs and commits =====
    {
      uint32_t el_;
      asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(el_));
      const bool issuer = el_ != 0;
      // kind::f16: fp16 x fp16 -> fp32, both operands K-major, N = 2*KP, M = 128
      const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(NCOL >> 3) << 17) | ((uint32_t)(CH_TM >> 4) << 24);
      const uint64_t adesc0 = make_desc_k128(smem_u32(a_base)), bdesc0 = make_desc_k128(smem_u32(n_base));
      const bool itimer = DBG && blockIdx.x == 0;
      long long iacc[3] = {0, 0, 0}, iprev = clock64();
      auto itick = [&](int q) { if (itimer) { long long now = clock64(); iacc[q] += now - iprev; iprev = now; } };
      // MMAs of (tile k, piece p of the site).  descriptor start addresses count 16-byte units: +2 per K=16 step
      // inside an atom, +1024 per 16 KB atom
      auto issue = [&](int k, int p, uint32_t st) {
        const uint32_t d = tmem_base + k * NCOL;
        const uint64_t ad = adesc0 + (uint64_t)(k * (C::A_BYTES >> 4));
        const uint64_t bd = bdesc0 + (uint64_t)(st * (C::PIECE_BYTES >> 4));
        if (KP == 32) {                                    // small terms first (see "Order of accumulation")
          if (issuer) umma_f16(d, ad + 4, bd + 0, idesc, 0u);          // lo . hi
          if (issuer) umma_f16(d, ad + 6, bd + 2, idesc, 1u);
          if (issuer) umma_f16(d, ad + 0, bd + 4, idesc, 1u);          // hi . lo
          if (issuer) umma_f16(d, ad + 2, bd + 6, idesc, 1u);
          if (issuer) umma_f16(d, ad + 0, bd + 0, idesc, 1u);          // hi . hi
          if (issuer) umma_f16(d, ad + 2, bd + 2, idesc, 1u);
        } else {
          constexpr int KA = KP / 64;
          const int ka = p < KA ? p : p - KA;
          const uint64_t ahi = ad + (uint64_t)(ka * 1024), alo = ad + (uint64_t)((KA + ka) * 1024);
          if (p < KA) {                                    // piece holds g_lo of atom ka: a_hi . g_lo
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) if (issuer) umma_f16(d, ahi + 2 * ks, bd + 2 * ks, idesc, (p > 0 || ks > 0) ? 1u : 0u);
          } else {                                         // piece holds g_hi of atom ka: a_lo . g_hi, then a_hi . g_hi
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) if (issuer) umma_f16(d, alo + 2 * ks, bd + 2 * ks, idesc, 1u);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) if (issuer) umma_f16(d, ahi + 2 * ks, bd + 2 * ks, idesc, 1u);
          }
        }
      };
      auto wait_a = [&](int k, int s) {
        mbar_wait(&a_ready[k], (uint32_t)(s & 1));
        tc_fence_after();
        itick(1);
      };
      for (int s = 0; s < nsteps; ++s) {
        const bool ev = itimer && s == nsteps / 2;
        if (ev && issuer) g_chain_dbg[16] = (float)(clock64() - t_base);
        if (C::TILE_MAJOR) {
          // tile 0 waits for a piece, the last tile releases it
#pragma unroll
          for (int k = 0; k < TILES; ++k) {
            wait_a(k, s);
#pragma unroll
            for (int p = 0; p < PIECES; ++p) {
              const uint32_t q = (uint32_t)(s * PIECES + p), st = q % NSTAGE;
              if (k == 0) { mbar_wait(&node_full[st], (q / NSTAGE) & 1u); itick(0); }
              issue(k, p, st);
              if (k == TILES - 1) if (issuer) umma_commit(&node_empty[st]);
            }
            if (issuer) umma_commit(&x_ready[k]);
            itick(2);
            if (ev && issuer && k < 4) g_chain_dbg[17 + k] = (float)(clock64() - t_base);      // issued tile k
          }
        } else {
          // Two tiles sharing a 3-stage ring.  Tile 0 runs two pieces ahead of tile 1,
          //     (T0,p0) (T0,p1) (T1,p0) (T0,p2) (T1,p1) (T0,p3) (T1,p2) (T1,p3),
          // which keeps at most two pieces live (+1 prefetch) and finishes tile 0 a third of a site before tile 1: the
          // epilogue of tile 0 runs under the last MMAs of tile 1, the epilogue of tile 1 under the first of tile 0.
          static_assert(C::TILE_MAJOR || (TILES == 2 && PIECES == 4), "staggered schedule is written for 2 tiles x 4 pieces");
          constexpr int SK[8] = {0, 0, 1, 0, 1, 0, 1, 1}, SP[8] = {0, 1, 0, 2, 1, 3, 2, 3};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int k = SK[i], p = SP[i];
            const uint32_t q = (uint32_t)(s * PIECES + p), st = q % NSTAGE;
            if (k == 0) { mbar_wait(&node_full[st], (q / NSTAGE) & 1u); itick(0); }
            if (p == 0) wait_a(k, s);
            issue(k, p, st);
            if (p == PIECES - 1) {
              if (issuer) umma_commit(&x_ready[k]);
              if (ev && issuer) g_chain_dbg[17 + k] = (float)(clock64() - t_base);
            }
            if (k == 1) if (issuer) umma_commit(&node_empty[st]);
            itick(2);
          }
        }
      }
      if (itimer && issuer) for (int q = 0; q < 3; ++q) g_chain_dbg[4 + q] = (float)iacc[q];
    }
  } else {
    // ===== epilogue: warpgroup g of tile k owns state columns [g*JW, (g+1)*JW) of every image row =====
    const int wgi = warp >> 2, k = wgi / WG, g = wgi % WG, quad = warp & 3;
    const int row = quad * 32 + lane;
    const int64_t t = ((int64_t)blockIdx.x * TILES + k) * CH_TM + row;
    const bool live = t < a.B;
    uint8_t* at = a_base + k * C::A_BYTES;
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(k * NCOL + g * JW);
    int E = 0;               // true state = stored state * 2^E
    int pos_prev = CH_TROW;  // exponent of the stored row maximum of the previous step
    int sh_pred = CH_TROW - (CH_TROW + CH_TEXP);     // first guess: growth factor one (node maximum 2^TEXP)

    // 8 scaled state values -> packed fp16 hi and lo
    auto convert8 = [&](const float* v, float sc, uint4& hi, uint4& lo) {
      __align__(16) __half2 h2[4];
      __align__(16) __half2 l2[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float x = v[2 * q] * sc, y = v[2 * q + 1] * sc;
        h2[q] = __floats2half2_rn(x, y);
        const float2 hf = __half22float2(h2[q]);
        l2[q] = __floats2half2_rn(x - hf.x, y - hf.y);
      }
      hi = *reinterpret_cast<const uint4*>(h2);
      lo = *reinterpret_cast<const uint4*>(l2);
    };
    // 16-byte chunk kc (state columns 8 kc .. 8 kc + 7) of my A row
    auto store8 = [&](int kc, const uint4& hi, const uint4& lo) {
      uint32_t oh, ol;
      if (KP == 32) { oh = k128_chunk_off(row, kc); ol = k128_chunk_off(row, 4 + kc); }
      else {
        constexpr int KA = KP / 64;
        oh = (uint32_t)((kc >> 3) * 16384) + k128_chunk_off(row, kc & 7);
        ol = (uint32_t)((KA + (kc >> 3)) * 16384) + k128_chunk_off(row, kc & 7);
      }
      *reinterpret_cast<uint4*>(at + oh) = hi;
      *reinterpret_cast<uint4*>(at + ol) = lo;
    };
    // write 32 scaled state values (state columns j0 .. j0+31) as fp16 hi/lo chunks of my A row
    auto store_chunk32 = [&](const float (&v)[32], int j0, float sc) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 hi, lo;
        convert8(&v[8 * c], sc, hi, lo);
        store8((j0 >> 3) + c, hi, lo);
      }
    };
    auto publish = [&]() {
      fence_async_proxy();
      tc_fence_before();
      mbar_arrive(&a_ready[k]);
    };
    // environment after a step in fp32 (training caches): 32 of my columns, true value = v * 2^e
    auto emit32 = [&](const float (&v)[32], int j0, int e, float* base) {
      if (!live) return;
      float* o = base + t * a.Ds + j0;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        if (j0 + j < a.Ds)
          *reinterpret_cast<float4*>(o + j) = make_float4(scale_pow2(v[j], e), scale_pow2(v[j + 1], e), scale_pow2(v[j + 2], e), scale_pow2(v[j + 3], e));
    };
    if (nsteps > 0 && a.start_kind == 2) {
      // start from an environment in HBM (one training step of the chain, _find_C1/_find_C2 baseoptimizer.py:160-201):
      // my JW columns of my row, row maximum over the warpgroups of the tile, exact power-of-two shift into the fp16 window
      float v[JW / 32][32];
      float mx = 0.f;
#pragma unroll
      for (int c = 0; c < JW / 32; ++c) {
        const int j0 = g * JW + c * 32;
        const float* src = a.in + t * a.Ds + j0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
          if (live && j0 + j < a.Ds) x = __ldg(reinterpret_cast<const float4*>(src + j));
          v[c][j] = x.x; v[c][j + 1] = x.y; v[c][j + 2] = x.z; v[c][j + 3] = x.w;
          mx = fmaxf(mx, fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w))));
        }
      }
      if (WG > 1) {
        mx_x[(k * WG + g) * CH_TM + row] = mx;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + k), "r"(WG * 128) : "memory");
#pragma unroll
        for (int o = 0; o < WG; ++o) mx = fmaxf(mx, mx_x[(k * WG + o) * CH_TM + row]);
      }
      const bool valid = mx > 0.f && mx < 3.0e38f;
      const int sh = valid ? max(-120, min(120, CH_TROW - exponent_of(mx))) : 0;
      const float sc = pow2f(sh);
#pragma unroll
      for (int c = 0; c < JW / 32; ++c) store_chunk32(v[c], g * JW + c * 32, sc);
      if (WG > 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + k), "r"(WG * 128) : "memory");      // mx_x is reused in the first step
      E = -sh;
      publish();
    } else if (nsteps > 0) {
      // start vector (mps.py:446-466): exactly representable, row maximum 1 -> shift by TROW
#pragma unroll 1
      for (int c = 0; c < JW / 32; ++c) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int jj = g * JW + c * 32 + j;
          v[j] = a.start_kind == 0 ? ((jj >= a.L && jj < 2 * a.L) ? 1.f : 0.f) : (jj < a.L ? 1.f : 0.f);
        }
        store_chunk32(v, g * JW + c * 32, pow2f(CH_TROW));
      }
      E = -CH_TROW;
      publish();
    }
    const bool timer = DBG && blockIdx.x == 0 && tid == 0;
    const bool evt = DBG && blockIdx.x == 0 && row == 0 && g == 0 && k < 2;      // event stamps of tiles 0 and 1
    long long tacc[4] = {0, 0, 0, 0}, tprev = timer ? clock64() : 0;
    auto tick = [&](int q) { if (timer) { long long now = clock64(); tacc[q] += now - tprev; tprev = now; } };
    // raw feature of a site for my image: pixel (in .x) or the precomputed phi; fetched TWO sites ahead so that the
    // load latency never meets the arithmetic that consumes it
    // running pointer to my feature of the site two steps ahead (one 64-bit add per step instead of a multiply)
    const int64_t feat_stride = feat_site_offset(a.fm, a.B, 1) * a.dir;
    const float* feat_ptr = a.feat + feat_site_offset(a.fm, a.B, a.first) + (a.fm == TRMPS_FM_PRECOMPUTED ? 2 : 1) * t;
    auto load_raw = [&](int step) -> float2 {          // called with step = 0, 1, 2, ... in order
      float2 r = make_float2(0.f, 0.f);
      if (live && step < nsteps) {
        if (a.fm == TRMPS_FM_PRECOMPUTED) r = __ldg(reinterpret_cast<const float2*>(feat_ptr));
        else r.x = __ldg(feat_ptr);
      }
      feat_ptr += feat_stride;
      return r;
    };
    auto phi_of_raw = [&](float2 raw) -> float2 {
      if (!live) return make_float2(0.f, 0.f);
      return a.fm == TRMPS_FM_PRECOMPUTED ? raw : phi_of_pixel(raw.x, a.fm);
    };
    float2 ph = phi_of_raw(load_raw(0));
    float2 raw1 = load_raw(1);
    for (int s = 0; s < nsteps; ++s) {
      const float2 raw2 = load_raw(s + 2);          // in flight for two steps
      const float2 ph_next = phi_of_raw(raw1);      // loaded one step ago
      raw1 = raw2;
      const int fs_exp = __ldg(a.exps + s);
      const bool last = s + 1 == nsteps;
      tick(0);
      mbar_wait(&x_ready[k], (uint32_t)(s & 1));
      tc_fence_after();
      if (evt && s == nsteps / 2) g_chain_dbg[32 + k * 8] = (float)(clock64() - t_base);        // x_ready seen
      tick(1);
      // new state v[j] = phi_0 X[(0,j)] + phi_1 X[(1,j)], 32 of my columns at a time
      // (sc is a power of two, so folding it into phi is exact)
      auto chunk = [&](int c, float (&v)[32], float sc = 1.f) {
        uint32_t x0[32], x1[32];
        tmem_ld32_issue(taddr + c * 32, x0);
        tmem_ld32_issue(taddr + KP + c * 32, x1);
        tmem_ld_wait();
        const float p0 = ph.x * sc, p1 = ph.y * sc;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(p0, __uint_as_float(x0[j]), p1 * __uint_as_float(x1[j]));
      };
      if (DBG && a.dbg == 1 && !last) {
        publish();
      } else if (last) {
        // final environment in fp32: undo both exponents
        const int e = E + fs_exp;
#pragma unroll 1
        for (int c = 0; c < JW / 32; ++c) {
          float v[32];
          chunk(c, v);
          if (live) {
            const int j0 = g * JW + c * 32;
            float* o = a.out + t * a.Ds + j0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (j0 + j < a.Ds)
                *reinterpret_cast<float4*>(o + j) = make_float4(scale_pow2(v[j], e), scale_pow2(v[j + 1], e), scale_pow2(v[j + 2], e), scale_pow2(v[j + 3], e));
          }
        }
        tick(2);
      } else {
        // single pass with the predicted shift; redo (warp-uniformly: tcgen05.ld is warp-collective) if the realised
        // row maximum leaves the safe window
        int sh = max(-120, min(120, sh_pred));
        float mx = 0.f;
        {
          const float sc = pow2f(sh);
#pragma unroll 1
          for (int c = 0; c < JW / 32; ++c) {
            float v[32];
            chunk(c, v, sc);                   // already scaled
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fabsf(v[j]));
            store_chunk32(v, g * JW + c * 32, 1.f);
            if (a.out_all) emit32(v, g * JW + c * 32, E + fs_exp - sh, a.out_all + (int64_t)s * a.out_stride);
          }
        }
        if (WG > 1) {
          // the row maximum spans the warpgroups of the tile: exchange the partial maxima (named barrier per tile)
          mx_x[(k * WG + g) * CH_TM + row] = mx;
          asm volatile("bar.sync %0, %1;" ::"r"(1 + k), "r"(WG * 128) : "memory");
#pragma unroll
          for (int o = 0; o < WG; ++o) mx = fmaxf(mx, mx_x[(k * WG + o) * CH_TM + row]);
        }
        const bool valid = mx > 0.f && mx < 3.0e38f;
        int pos = valid ? exponent_of(mx) : 0;   // exponent of the stored row maximum
        const int ex = pos - sh;                 // exponent of the raw maximum
        tick(2);
        if (__any_sync(0xffffffffu, valid && (pos > 14 || pos < CH_TROW - 10))) {
          if (valid) sh = max(-120, min(120, CH_TROW - ex));
          const float sc = pow2f(sh);
#pragma unroll 1
          for (int c = 0; c < JW / 32; ++c) {
            float v[32];
            chunk(c, v);
            store_chunk32(v, g * JW + c * 32, sc);
            if (a.out_all) emit32(v, g * JW + c * 32, E + fs_exp, a.out_all + (int64_t)s * a.out_stride);
          }
          pos = ex + sh;
        }
        if (WG > 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + k), "r"(WG * 128) : "memory");      // mx_x is reused next step
        E += fs_exp - sh;
        // next raw maximum ~ 2^pos * (growth seen now): growth = ex - pos_prev  (the node scale 2^13 is inside both)
        if (valid) { sh_pred = CH_TROW - (pos + ex - pos_prev); pos_prev = pos; }
        publish();
        if (evt && s == nsteps / 2) g_chain_dbg[32 + k * 8 + 4] = (float)(clock64() - t_base);          // published
      }
      ph = ph_next;
      tick(3);
    }
    if (timer) for (int q = 0; q < 4; ++q) g_chain_dbg[q] = (float)tacc[q];
    if (nsteps == 0 && live && g == 0) {
      float* o = a.out + t * a.Ds;
      for (int j = 0; j < a.Ds; ++j) o[j] = a.start_kind == 0 ? ((j >= a.L && j < 2 * a.L) ? 1.f : 0.f) : (j < a.L ? 1.f : 0.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TILES * NCOL);
}

template <int KP>
int launch_chain_kp(const ChainArgs& a, const float* nodes, uint8_t* img, int32_t* exps, cudaStream_t st) {
  if (a.nsteps > 0) {
    chain_exp_kernel<<<a.nsteps, 1024, 0, st>>>(nodes, exps, a.Ds, a.first, a.dir);
    TRMPS_LAUNCH_OK("chain_exp_kernel");
    chain_prep_kernel<KP><<<296, 256, 0, st>>>(nodes, img, exps, a.Ds, a.first, a.dir, a.nsteps);
    TRMPS_LAUNCH_OK("chain_prep_kernel");
  }
  const unsigned grid = (unsigned)ceil_div64(a.B, ChainCfg<KP>::TILES * CH_TM);
  if (a.dbg != 0) {
    static bool attr_set = false;
    if (!attr_set) {
      TRMPS_CUDA_OK(cudaFuncSetAttribute(chain_tc_kernel<KP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, chain_smem<KP>()));
      attr_set = true;
    }
    chain_tc_kernel<KP, true><<<grid, chain_threads<KP>(), chain_smem<KP>(), st>>>(a);
  } else {
    static bool attr_set = false;
    if (!attr_set) {
      TRMPS_CUDA_OK(cudaFuncSetAttribute(chain_tc_kernel<KP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, chain_smem<KP>()));
      attr_set = true;
    }
    chain_tc_kernel<KP, false><<<grid, chain_threads<KP>(), chain_smem<KP>(), st>>>(a);
  }
  TRMPS_LAUNCH_OK("chain_tc_kernel");
  return TRMPS_OK;
}

int chain_kp(int Ds) { return Ds <= 32 ? 32 : Ds <= 64 ? 64 : 128; }
int64_t chain_site_bytes(int Ds) {
  const int kp = chain_kp(Ds);
  return kp == 32 ? site_bytes<32>() : kp == 64 ? site_bytes<64>() : site_bytes<128>();
}

}  // namespace

bool chain_tc_supported(int Ds) { return Ds <= 128 && Ds % 4 == 0; }
// Expected relative shrink of the logits after `nsites` fused sites: the round-toward-zero accumulation of the KP/16
// h.h instructions per site ("Order of accumulation" above).  Per-site figures measured on dense random chains
// (tools/chain_accuracy.py: 4.0e-5 after 783 sites at KP=32, 6.0e-5 after 783 at KP=64, 7.4e-5 after 195 at KP=128).
double chain_tc_bias_estimate(int Ds, int nsites) {
  const double per_site = Ds <= 32 ? 5.2e-8 : Ds <= 64 ? 7.8e-8 : 3.8e-7;
  return per_site * (nsites > 0 ? nsites : 0);
}
int chain_tc_debug(float* out64) { return cudaMemcpyFromSymbol(out64, g_chain_dbg, 64 * sizeof(float)) == cudaSuccess ? 0 : -1; }
// scratch for the images of `nsteps` sites plus their exponents (in floats)
int64_t chain_tc_image_floats(int Ds, int nsteps) {
  if (nsteps < 1) nsteps = 1;
  return (int64_t)nsteps * (chain_site_bytes(Ds) / 4) + (((int64_t)nsteps + 3) / 4) * 4 + 8;
}

// runs `nsteps` chain steps starting at site `first` in direction dir from the boundary vector (start_kind 0 / 1) or from the
// environment in_env (start_kind 2); writes the final environment (B x Ds) to `out` and, when out_all is given, the
// environment after step s < nsteps - 1 to out_all + s * out_stride.  img: 16-byte aligned scratch of
// chain_tc_image_floats(Ds, nsteps) floats.
int launch_chain_tc(const float* feat, int fm, int64_t B, int Ds, int L, const float* nodes, int first, int dir, int nsteps,
                    int start_kind, float* img, float* out, cudaStream_t st, const float* in_env, float* out_all, int64_t out_stride) {
  if (B <= 0) return TRMPS_OK;
  if (!chain_tc_supported(Ds)) {
    set_error("chain_tc: Ds=%d not supported (<= 128, multiple of 4)", Ds);
    return TRMPS_E_ARG;
  }
  if (start_kind == 2 && !in_env) { set_error("chain_tc: start_kind 2 needs an input environment"); return TRMPS_E_ARG; }
  ChainArgs a;
  a.feat = feat; a.fm = fm; a.B = B; a.Ds = Ds; a.L = L; a.first = first; a.dir = dir; a.nsteps = nsteps;
  a.start_kind = start_kind; a.out = out; a.in = in_env; a.out_all = out_all; a.out_stride = out_stride;
  static const int dbg_mode = getenv("TRMPS_CHAIN_DEBUG") ? atoi(getenv("TRMPS_CHAIN_DEBUG")) : 0;
  a.dbg = dbg_mode;
  uint8_t* image = reinterpret_cast<uint8_t*>(img);
  int32_t* exps = reinterpret_cast<int32_t*>(image + (size_t)(nsteps < 1 ? 1 : nsteps) * chain_site_bytes(Ds));
  a.img = image; a.exps = exps;
  const int kp = chain_kp(Ds);
  return kp == 32 ? launch_chain_kp<32>(a, nodes, image, exps, st)
       : kp == 64 ? launch_chain_kp<64>(a, nodes, image, exps, st)
                  : launch_chain_kp<128>(a, nodes, image, exps, st);
}

}  // namespace trmps
