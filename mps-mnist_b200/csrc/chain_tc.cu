// Fused on-chip environment chain on the 5th-generation tensor cores (north_star pieces 1 + 2, inference path).
// Replaces the loops of MPS.predict (mps.py:536-549: tf.while_loop over _chain_multiply_r / _chain_multiply_l),
// which the reference runs one site at a time with a (B x D x D) intermediate in memory.
//
// One CTA carries 4 tiles of 128 images through ALL sites of a chain without touching HBM except for the pixel
// stream: per site and tile
//     X[t,(m,j)] = sum_i S[t,i] * node[m,i,j]        tcgen05.mma, M=128, K=KP, N=2*KP, 3xTF32, accumulator in TMEM
//     S'[t,j]    = phi_0(x_t) X[t,(0,j)] + phi_1(x_t) X[t,(1,j)]      epilogue in registers (feature map fused)
// and S' is re-split into TF32 hi/lo and written straight back into the K-major SWIZZLE_128B shared-memory tile
// that is the A operand of the next site.  The four tiles are staggered: while the tensor core works on one tile,
// the epilogue warpgroups of the others drain TMEM and rebuild their A tiles.  Node matrices (pre-split hi/lo
// images built once per call, L2 resident) are streamed by one warp with cp.async.bulk through an mbarrier ring.
//
// Warp roles (576 threads): warpgroups 0-3 = epilogue of tile 0-3, warp 16 = MMA issuer, warp 17 = node loader.
#include "common.cuh"

namespace trmps {

namespace {

constexpr int CH_KP = 32;                    // padded bond (K of the per-site GEMM); N = 2*KP
constexpr int CH_TILES = 4;                  // image tiles per CTA
constexpr int CH_TM = 128;
constexpr int CH_THREADS = CH_TILES * 128 + 64;
constexpr int CH_NSTAGE = 4;                 // node-image ring
constexpr int CH_A_BYTES = CH_TM * CH_KP * 4;            // 16 KB per (hi | lo)
constexpr int CH_IMG_BYTES = CH_KP * 2 * CH_KP * 4;      // 8 KB per (hi | lo)
constexpr int CH_SMEM = CH_TILES * 2 * CH_A_BYTES + CH_NSTAGE * 2 * CH_IMG_BYTES + 512 + 1024;

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
__device__ __forceinline__ float tf32_round(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
// B image (MN-major SWIZZLE_128B_BASE32B), N = 2*KP columns, KP rows (k): byte offset of element (col, k)
__host__ __device__ __forceinline__ uint32_t node_img_off(int col, int k) {
  const int blk = col >> 5, c32 = (col & 31) >> 3, e = col & 7, kg = k >> 2, k4 = k & 3;
  return (uint32_t)((kg * (2 * CH_KP / 32) + blk) * 512 + k4 * 128 + ((c32 ^ k4) << 5) + e * 4);
}

// node slots -> hi/lo B images in chain order.  step s uses site first + s*dir.
//   dir=+1: W[i][(m,j)] = node[m][i][j]     dir=-1: W[j][(m,i)] = node[m][i][j]
__global__ void chain_prep_kernel(const float* __restrict__ nodes, float* __restrict__ img, int Ds, int first, int dir, int nsteps) {
  constexpr int PER = CH_KP * 2 * CH_KP;      // elements per (hi | lo) image
  const int64_t total = (int64_t)nsteps * PER;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int s = (int)(e / PER), r = (int)(e % PER);
    const int k = r / (2 * CH_KP), col = r % (2 * CH_KP), m = col / CH_KP, n = col % CH_KP;
    const float* node = nodes + (size_t)(first + s * dir) * 2 * Ds * Ds;
    float v = 0.f;
    if (k < Ds && n < Ds) v = dir > 0 ? node[((size_t)m * Ds + k) * Ds + n] : node[((size_t)m * Ds + n) * Ds + k];
    const float hi = tf32_round(v), lo = v - hi;
    float* base = img + (size_t)s * 2 * PER;
    const uint32_t off = node_img_off(col, k) >> 2;
    base[off] = hi;
    base[PER + off] = lo;
  }
}

__device__ float g_chain_dbg[16];

struct ChainArgs {
  const float* feat; int fm; int64_t B; int Ds, L; int first, dir, nsteps; int start_kind;   // start_kind 0: [0_L,1_L], 1: [1_L,0_L]
  const float* img; float* out;      // out: (B, Ds) final environment
};

__global__ void __launch_bounds__(CH_THREADS, 1) chain_tc_kernel(ChainArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_base = smem;                                              // [tile][hi|lo][16 KB]
  uint8_t* n_base = smem + CH_TILES * 2 * CH_A_BYTES;                  // [stage][hi|lo][8 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(n_base + CH_NSTAGE * 2 * CH_IMG_BYTES);
  uint64_t* node_full = bars, *node_empty = bars + CH_NSTAGE, *a_ready = bars + 2 * CH_NSTAGE, *x_ready = bars + 2 * CH_NSTAGE + CH_TILES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * CH_NSTAGE + 2 * CH_TILES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nsteps = a.nsteps;

  if (tid == 0) {
    for (int s = 0; s < CH_NSTAGE; ++s) { mbar_init(&node_full[s], 1); mbar_init(&node_empty[s], 1); }
    for (int k = 0; k < CH_TILES; ++k) { mbar_init(&a_ready[k], 128); mbar_init(&x_ready[k], 1); }
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == CH_TILES * 4 + 1) {
    // ===== node loader =====
    if (lane == 0) {
      for (int s = 0; s < nsteps; ++s) {
        const int st = s % CH_NSTAGE;
        if (s >= CH_NSTAGE) mbar_wait(&node_empty[st], (uint32_t)(((s / CH_NSTAGE) - 1) & 1));
        mbar_expect_tx(&node_full[st], 2 * CH_IMG_BYTES);
        bulk_g2s(n_base + st * 2 * CH_IMG_BYTES, a.img + (size_t)s * (2 * CH_IMG_BYTES / 4), 2 * CH_IMG_BYTES, &node_full[st]);
      }
    }
  } else if (warp == CH_TILES * 4) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)((2 * CH_KP) >> 3) << 17) | ((uint32_t)(CH_TM >> 4) << 24);
      const bool itimer = blockIdx.x == 0;
      long long iacc[3] = {0, 0, 0}, iprev = clock64();
      auto itick = [&](int q) { if (itimer) { long long now = clock64(); iacc[q] += now - iprev; iprev = now; } };
      for (int s = 0; s < nsteps; ++s) {
        const int st = s % CH_NSTAGE;
        mbar_wait(&node_full[st], (uint32_t)((s / CH_NSTAGE) & 1));
        itick(0);
        const uint32_t bh = smem_u32(n_base + st * 2 * CH_IMG_BYTES), bl = bh + CH_IMG_BYTES;
        for (int k = 0; k < CH_TILES; ++k) {
          mbar_wait(&a_ready[k], (uint32_t)(s & 1));
          tc_fence_after();
          itick(1);
          const uint32_t ah = smem_u32(a_base + k * 2 * CH_A_BYTES), al = ah + CH_A_BYTES;
          const uint32_t d = tmem_base + k * 2 * CH_KP;
#pragma unroll
          for (int ks = 0; ks < CH_KP / 8; ++ks) {
            const uint64_t dah = make_desc(ah + ks * 32, 16, 1024, 2), dal = make_desc(al + ks * 32, 16, 1024, 2);
            const uint32_t boff = (uint32_t)(2 * ks * (2 * CH_KP / 32) * 512);
            const uint64_t dbh = make_desc(bh + boff, 512, (2 * CH_KP / 32) * 512, 1), dbl = make_desc(bl + boff, 512, (2 * CH_KP / 32) * 512, 1);
            umma_tf32(d, dah, dbh, idesc, ks > 0 ? 1u : 0u);
            umma_tf32(d, dal, dbh, idesc, 1u);
            umma_tf32(d, dah, dbl, idesc, 1u);
          }
          umma_commit(&x_ready[k]);
          itick(2);
        }
        umma_commit(&node_empty[st]);
      }
      if (itimer) for (int q = 0; q < 3; ++q) g_chain_dbg[4 + q] = (float)iacc[q];
    }
  } else {
    // ===== epilogue warpgroup of tile k: one image row per thread =====
    const int k = warp >> 2, quad = warp & 3;
    const int row = quad * 32 + lane;
    const int64_t t = ((int64_t)blockIdx.x * CH_TILES + k) * CH_TM + row;
    const bool live = t < a.B;
    uint8_t* ah = a_base + k * 2 * CH_A_BYTES, *al = ah + CH_A_BYTES;
    const uint32_t roff = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
    const int r8 = row & 7;
    float snew[CH_KP];
    // start vector (mps.py:446-466)
#pragma unroll
    for (int j = 0; j < CH_KP; ++j)
      snew[j] = a.start_kind == 0 ? ((j >= a.L && j < 2 * a.L) ? 1.f : 0.f) : (j < a.L ? 1.f : 0.f);
    auto store_state = [&]() {
#pragma unroll
      for (int c = 0; c < CH_KP / 4; ++c) {
        float4 hi, lo;
        hi.x = tf32_round(snew[4 * c]); hi.y = tf32_round(snew[4 * c + 1]); hi.z = tf32_round(snew[4 * c + 2]); hi.w = tf32_round(snew[4 * c + 3]);
        lo.x = snew[4 * c] - hi.x; lo.y = snew[4 * c + 1] - hi.y; lo.z = snew[4 * c + 2] - hi.z; lo.w = snew[4 * c + 3] - hi.w;
        const uint32_t off = roff + (uint32_t)((c ^ r8) << 4);
        *reinterpret_cast<float4*>(ah + off) = hi;
        *reinterpret_cast<float4*>(al + off) = lo;
      }
      fence_async_proxy();
      tc_fence_before();
      mbar_arrive(&a_ready[k]);
    };
    if (nsteps > 0) store_state();
    const bool timer = blockIdx.x == 0 && tid == 0;
    long long tacc[4] = {0, 0, 0, 0}, tprev = timer ? clock64() : 0;
    auto tick = [&](int q) { if (timer) { long long now = clock64(); tacc[q] += now - tprev; tprev = now; } };
    // raw feature of a site for my image: pixel (in .x) or the precomputed phi; fetched TWO sites ahead so that the
    // load latency never meets the sincos that consumes it
    auto load_raw = [&](int step) -> float2 {
      if (!live || step >= nsteps) return make_float2(0.f, 0.f);
      const float* fs = a.feat + feat_site_offset(a.fm, a.B, a.first + step * a.dir);
      if (a.fm == TRMPS_FM_PRECOMPUTED) return __ldg(reinterpret_cast<const float2*>(fs) + t);
      return make_float2(__ldg(fs + t), 0.f);
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
      tick(0);
      mbar_wait(&x_ready[k], (uint32_t)(s & 1));
      tc_fence_after();
      tick(1);
      float x0[32], x1[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(k * 2 * CH_KP);
      tmem_ld32(taddr, x0);
      tmem_ld32(taddr + CH_KP, x1);
#pragma unroll
      for (int j = 0; j < CH_KP; ++j) snew[j] = fmaf(ph.x, x0[j], ph.y * x1[j]);
      ph = ph_next;
      tick(2);
      if (s + 1 < nsteps) store_state();
      tick(3);
    }
    if (timer) for (int q = 0; q < 4; ++q) g_chain_dbg[q] = (float)tacc[q];
    if (live) {
      float* o = a.out + t * a.Ds;
      for (int j = 0; j < a.Ds; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(snew[j], snew[j + 1], snew[j + 2], snew[j + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace

bool chain_tc_supported(int Ds) { return Ds <= CH_KP && Ds % 4 == 0; }
int chain_tc_debug(float* out16) { return cudaMemcpyFromSymbol(out16, g_chain_dbg, 16 * sizeof(float)) == cudaSuccess ? 0 : -1; }
int64_t chain_tc_image_floats(int nsteps) { return (int64_t)nsteps * 2 * CH_KP * 2 * CH_KP; }

// runs `nsteps` chain steps starting at site `first` in direction dir from the boundary vector; writes the final
// environment (B x Ds).  img: scratch of chain_tc_image_floats(nsteps) floats.
int launch_chain_tc(const float* feat, int fm, int64_t B, int Ds, int L, const float* nodes, int first, int dir, int nsteps,
                    int start_kind, float* img, float* out, cudaStream_t st) {
  if (B <= 0) return TRMPS_OK;
  if (!chain_tc_supported(Ds)) {
    set_error("chain_tc: Ds=%d not supported (<= %d)", Ds, CH_KP);
    return TRMPS_E_ARG;
  }
  if (nsteps > 0) {
    chain_prep_kernel<<<296, 256, 0, st>>>(nodes, img, Ds, first, dir, nsteps);
    TRMPS_LAUNCH_OK("chain_prep_kernel");
  }
  ChainArgs a;
  a.feat = feat; a.fm = fm; a.B = B; a.Ds = Ds; a.L = L; a.first = first; a.dir = dir; a.nsteps = nsteps;
  a.start_kind = start_kind; a.img = img; a.out = out;
  static bool attr_set = false;
  if (!attr_set) {
    TRMPS_CUDA_OK(cudaFuncSetAttribute(chain_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM));
    attr_set = true;
  }
  chain_tc_kernel<<<(unsigned)ceil_div64(B, CH_TILES * CH_TM), CH_THREADS, CH_SMEM, st>>>(a);
  TRMPS_LAUNCH_OK("chain_tc_kernel");
  return TRMPS_OK;
}

}  // namespace trmps
