// Two-site bond gradient on the 5th-generation tensor cores (north_star piece 3).
//   G[a, c] = sum_t P[t, a] * W[t, c],   P[t,(m,i)] = phi_near[t,m] * e_near[t,i],   W[t,(l,n,k)] = sres[t,(l,n)] * e_far[t,k]
// (optimizer.py:249, without the regulariser).  M = 2*Ds rows, N = 2*L*Ds columns, K = images.
//
// tcgen05.mma kind::tf32 with FP32 accumulation in tensor memory; FP32 accuracy is kept with the 3xTF32 split
// a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo (dropped term ~2^-22), because plain TF32 (10-bit mantissa) misses
// the 1e-4 tolerance on this cancellation-heavy sum.
//
// Neither operand exists in memory: both are rank-1-per-image products.  512 producer threads of the CTA form the
// hi/lo tiles while staging global -> registers -> shared memory, directly in the MN-major 128-byte-swizzled
// canonical layout the MMA reads (both operands are "transposed": the contraction index t is the slow one in
// memory; for TF32 tcgen05 supports that only with the SWIZZLE_128B_BASE32B layout -- verified with
// tools/probe/tc_probe.cu), so every global load and every shared store is a float4.
// One CTA owns a 256 x 256 output tile = two 128 x 256 accumulators (all 512 TMEM columns) that share the W tile.
// A 3-stage ring of 64 KB stages is recycled through mbarriers signalled by tcgen05.commit; the MMAs of a stage
// are issued by one thread and run asynchronously while all threads build the next stage.
#include "common.cuh"

namespace trmps {

namespace {

constexpr int TC_FORMERS = 512;             // warps 0-15 form the operand tiles (warps 0-7 also drain TMEM at the end)
constexpr int TC_THREADS = TC_FORMERS + 32;  // + one warp that issues the MMAs
constexpr int TC_KCH = 16;                 // images per stage (two K=8 MMA steps)
constexpr int TC_STAGES = 3;
constexpr int TC_RPT = TC_KCH * 64 / TC_FORMERS;     // image rows of a chunk per producer thread
constexpr int TC_TM = 256, TC_TN = 256;    // CTA output tile
constexpr int TC_ATOM = 512;               // one swizzle atom: 4 k-rows x 128 B (32 MN elements), 32-byte units XORed with the row
constexpr int TC_A_BYTES = (TC_TM / 32) * (TC_KCH / 4) * TC_ATOM;   // 16 KB (hi) per stage
constexpr int TC_B_BYTES = (TC_TN / 32) * (TC_KCH / 4) * TC_ATOM;   // 16 KB
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;     // A_hi, A_lo, B_hi, B_lo = 64 KB
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;

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

// shared-memory matrix descriptor, MN-major, SWIZZLE_128B_BASE32B: LBO = stride between 32-element MN blocks,
// SBO = stride between groups of 4 k (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>)
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;      // descriptor version 1 (Blackwell)
  d |= (uint64_t)1 << 61;      // SWIZZLE_128B_BASE32B
  return d;
}

// tf32 x tf32 -> f32, both operands MN-major, M = 128, N = 256
__device__ __forceinline__ uint32_t make_idesc_tf32_mn(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                 // D format F32
  d |= 2u << 7;                 // A format TF32
  d |= 2u << 10;                // B format TF32
  d |= 1u << 15;                // A MN-major
  d |= 1u << 16;                // B MN-major
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

// round-to-nearest (ties away) to TF32 precision with two integer-pipe ops; cvt.rna.tf32.f32 goes through the
// slow conversion unit and was the bottleneck of the operand producers
__device__ __forceinline__ float tf32_round(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

struct GradTcArgs {
  const float* e_near; const float* e_far; const float* phi_near; const float* sres; float* gpart;
  int64_t B, imgs_per_split; int Ds, L, R, Cc; float* dbg;
};

// byte offset of the 16-byte chunk holding MN elements [mn, mn+4) of image-row k inside a tile whose groups of
// 4 k are `atoms_per_k` atoms apart (atoms_per_k = MN extent / 32)
__device__ __forceinline__ uint32_t tile_off(int mn, int k, int atoms_per_k) {
  const int blk = mn >> 5, c32 = (mn & 31) >> 3, half = (mn & 7) >> 2, kg = k >> 2, k4 = k & 3;
  return (uint32_t)((kg * atoms_per_k + blk) * TC_ATOM + k4 * 128 + ((c32 ^ k4) << 5) + half * 16);
}

// SQ: the diagonal Hessian (optimizer.py:229-237): the same contraction with squared operands, sum_t hres[t,(l,n)] (phi e_near)^2 e_far^2
template <bool SQ>
__global__ void __launch_bounds__(TC_THREADS, 1) grad_tc_kernel(GradTcArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);   // empty[TC_STAGES], done, full[TC_STAGES]
  uint64_t* full = bars + TC_STAGES + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Ds = g.Ds, L2 = 2 * g.L;
  const int c0 = blockIdx.x * TC_TN, a0 = blockIdx.y * TC_TM;
  const int64_t tbeg = (int64_t)blockIdx.z * g.imgs_per_split;
  int64_t tend = tbeg + g.imgs_per_split;
  if (tend > g.B) tend = g.B;
  const int nch = tend > tbeg ? (int)((tend - tbeg + TC_KCH - 1) / TC_KCH) : 0;

  if (tid == 0) {
    for (int s = 0; s <= TC_STAGES; ++s) mbar_init(&bars[s], 1);
    for (int s = 0; s < TC_STAGES; ++s) mbar_init(&full[s], TC_FORMERS);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // this thread stages one quad of A rows and one quad of B columns for image rows ksub + KS r, r = 0..TC_RPT-1
  constexpr int KS = TC_FORMERS / 64;
  const int q4 = tid & 63, ksub = tid >> 6;
  const int arow = a0 + 4 * q4, ccol = c0 + 4 * q4;
  const bool va = arow < g.R, vc = ccol < g.Cc;
  const int am = va ? arow / Ds : 0, ai = va ? arow - am * Ds : 0;
  const int cln = vc ? ccol / Ds : 0, ck = vc ? ccol - cln * Ds : 0;
  const uint32_t idesc = make_idesc_tf32_mn(128, TC_TN);

  const bool timer = g.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 32;
  long long tacc[6] = {0, 0, 0, 0, 0, 0}, tprev = timer ? clock64() : 0;
  auto tick = [&](int k) { if (timer) { long long now = clock64(); tacc[k] += now - tprev; tprev = now; } };

  // global -> registers for one chunk: RAW rows of e_near / e_far plus their per-image scale factors for image rows
  // ksub + 4 r.  Nothing here consumes a loaded value, so all 16 loads of a chunk stay in flight.
  auto load_chunk = [&](int ch, float4 (&ra)[TC_RPT], float4 (&rb)[TC_RPT], float (&pa)[TC_RPT], float (&pb)[TC_RPT]) {
#pragma unroll
    for (int r = 0; r < TC_RPT; ++r) {
      const int64_t t = tbeg + (int64_t)ch * TC_KCH + ksub + KS * r;
      ra[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      rb[r] = ra[r];
      pa[r] = 0.f;
      pb[r] = 0.f;
      if (t < tend) {
        if (va) {
          ra[r] = __ldg(reinterpret_cast<const float4*>(g.e_near + t * Ds + ai));
          pa[r] = __ldg(g.phi_near + 2 * t + am);
        }
        if (vc) {
          rb[r] = __ldg(reinterpret_cast<const float4*>(g.e_far + t * Ds + ck));
          pb[r] = __ldg(g.sres + t * L2 + cln);
        }
      }
    }
  };
  // registers -> tf32 hi/lo tiles in stage ch % TC_STAGES, then one thread issues the MMAs of the chunk
  auto process_chunk = [&](int ch, const float4 (&ra)[TC_RPT], const float4 (&rb)[TC_RPT], const float (&pa)[TC_RPT], const float (&pb)[TC_RPT]) {
    const int s = ch % TC_STAGES, it = ch / TC_STAGES;
    uint8_t* st = smem + s * TC_STAGE_BYTES;
    uint8_t* a_hi = st, *a_lo = st + TC_A_BYTES, *b_hi = st + 2 * TC_A_BYTES, *b_lo = st + 2 * TC_A_BYTES + TC_B_BYTES;
    tick(0);
    if (ch >= TC_STAGES) mbar_wait(&bars[s], (uint32_t)((it - 1) & 1));   // MMAs that read this stage are done
    tick(1);
#pragma unroll
    for (int r = 0; r < TC_RPT; ++r) {
      const int k = ksub + KS * r;
      {
        float4 hi, lo;
        float vx = ra[r].x * pa[r], vy = ra[r].y * pa[r], vz = ra[r].z * pa[r], vw = ra[r].w * pa[r];
        if (SQ) { vx *= vx; vy *= vy; vz *= vz; vw *= vw; }
        hi.x = tf32_round(vx); hi.y = tf32_round(vy); hi.z = tf32_round(vz); hi.w = tf32_round(vw);
        lo.x = vx - hi.x; lo.y = vy - hi.y; lo.z = vz - hi.z; lo.w = vw - hi.w;
        const uint32_t off = tile_off(4 * q4, k, TC_TM / 32);
        *reinterpret_cast<float4*>(a_hi + off) = hi;
        *reinterpret_cast<float4*>(a_lo + off) = lo;
      }
      {
        float4 hi, lo;
        float vx = rb[r].x * pb[r], vy = rb[r].y * pb[r], vz = rb[r].z * pb[r], vw = rb[r].w * pb[r];
        if (SQ) { vx *= rb[r].x; vy *= rb[r].y; vz *= rb[r].z; vw *= rb[r].w; }          // the weight (hres) enters once
        hi.x = tf32_round(vx); hi.y = tf32_round(vy); hi.z = tf32_round(vz); hi.w = tf32_round(vw);
        lo.x = vx - hi.x; lo.y = vy - hi.y; lo.z = vz - hi.z; lo.w = vw - hi.w;
        const uint32_t off = tile_off(4 * q4, k, TC_TN / 32);
        *reinterpret_cast<float4*>(b_hi + off) = hi;
        *reinterpret_cast<float4*>(b_lo + off) = lo;
      }
    }
    tick(2);
    fence_async_proxy();          // make the generic-proxy stores visible to the tensor core (async proxy)
    tick(3);
    mbar_arrive(&full[s]);        // no block barrier: the issuer warp waits for all producers on the stage's mbarrier
    tick(4);
  };
  // MMAs of one chunk: the whole issuer warp runs this (uniform control flow and descriptor arithmetic, which the compiler
  // keeps in the uniform datapath), one elected lane issues -- under `if (tid == TC_FORMERS)` every descriptor was a
  // per-thread value moved to the uniform registers MMA by MMA (ELECT + R2UR.BROADCAST), ~150 cycles between MMAs
  auto issue_chunk = [&](int ch, bool issuer) {
    const int s = ch % TC_STAGES, it = ch / TC_STAGES;
    uint8_t* st = smem + s * TC_STAGE_BYTES;
    uint8_t* a_hi = st, *a_lo = st + TC_A_BYTES, *b_hi = st + 2 * TC_A_BYTES, *b_lo = st + 2 * TC_A_BYTES + TC_B_BYTES;
    mbar_wait(&full[s], (uint32_t)(it & 1));
    tc_fence_after();
    const uint32_t acc0 = ch > 0 ? 1u : 0u;
#pragma unroll
    for (int kb = 0; kb < TC_KCH / 8; ++kb) {
      const uint32_t koffA = 2 * kb * (TC_TM / 32) * TC_ATOM, koffB = 2 * kb * (TC_TN / 32) * TC_ATOM;   // two 4-k groups per MMA
      const uint64_t dbh = make_desc_mn_sw128(smem_u32(b_hi) + koffB, TC_ATOM, (TC_TN / 32) * TC_ATOM);
      const uint64_t dbl = make_desc_mn_sw128(smem_u32(b_lo) + koffB, TC_ATOM, (TC_TN / 32) * TC_ATOM);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const uint32_t moff = mt * 4 * TC_ATOM;      // rows 128.. of the A tile: MN blocks 4..7
        const uint64_t dah = make_desc_mn_sw128(smem_u32(a_hi) + koffA + moff, TC_ATOM, (TC_TM / 32) * TC_ATOM);
        const uint64_t dal = make_desc_mn_sw128(smem_u32(a_lo) + koffA + moff, TC_ATOM, (TC_TM / 32) * TC_ATOM);
        const uint32_t d = tmem_base + mt * TC_TN;
        const uint32_t acc = (kb > 0) ? 1u : acc0;
        if (issuer) {
          umma_tf32(d, dah, dbh, idesc, acc);
          umma_tf32(d, dal, dbh, idesc, 1u);
          umma_tf32(d, dah, dbl, idesc, 1u);
        }
      }
    }
    if (issuer) {
      umma_commit(&bars[s]);                          // stage s may be overwritten once these MMAs have read it
      if (ch == nch - 1) umma_commit(&bars[TC_STAGES]);
    }
  };

  // software pipeline: the global loads of chunks ch+1 and ch+2 are in flight while chunk ch is converted and stored
  if (tid < TC_FORMERS) {
    float4 ra0[TC_RPT], rb0[TC_RPT], ra1[TC_RPT], rb1[TC_RPT];
    float pa0[TC_RPT], pb0[TC_RPT], pa1[TC_RPT], pb1[TC_RPT];
    if (nch > 0) load_chunk(0, ra0, rb0, pa0, pb0);
    if (nch > 1) load_chunk(1, ra1, rb1, pa1, pb1);
    for (int ch = 0; ch < nch; ch += 2) {
      process_chunk(ch, ra0, rb0, pa0, pb0);
      if (ch + 2 < nch) load_chunk(ch + 2, ra0, rb0, pa0, pb0);
      if (ch + 1 < nch) {
        process_chunk(ch + 1, ra1, rb1, pa1, pb1);
        if (ch + 3 < nch) load_chunk(ch + 3, ra1, rb1, pa1, pb1);
      }
    }
  } else {
    uint32_t el;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(el));
    const bool issuer = el != 0;
    for (int ch = 0; ch < nch; ++ch) issue_chunk(ch, issuer);
  }

  if (timer) for (int k = 0; k < 6; ++k) g.dbg[k] = (float)tacc[k];
  // epilogue: TMEM -> registers -> global partial tile.  Warps 0-3 read accumulator 0, warps 4-7 accumulator 1.
  float* out = g.gpart + (size_t)blockIdx.z * g.R * g.Cc;
  const int mt = warp >> 2, quad = warp & 3;
  const int row = a0 + mt * 128 + quad * 32 + lane;
  if (warp >= 8) {
    // the issuer warp has nothing to drain
  } else if (nch > 0) {
    mbar_wait(&bars[TC_STAGES], 0);
    tc_fence_after();
#pragma unroll 1
    for (int cb = 0; cb < TC_TN / 32; ++cb) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mt * TC_TN + cb * 32), v);
      if (row < g.R) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int col = c0 + cb * 32 + j;
          if (col < g.Cc) *reinterpret_cast<float4*>(out + (size_t)row * g.Cc + col) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
    }
  } else if (row < g.R) {
    for (int col = c0; col < c0 + TC_TN && col < g.Cc; ++col) out[(size_t)row * g.Cc + col] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// split-K factor the caller must size `gpart` for
int grad_tc_splits(int64_t B, int Ds, int L, int sms) {
  const int R = 2 * Ds, Cc = 2 * L * Ds;
  const int tiles = ((Cc + TC_TN - 1) / TC_TN) * ((R + TC_TM - 1) / TC_TM);
  int gs = sms / tiles;
  if (gs < 1) gs = 1;
  int64_t cap = B / 256;
  if (cap < 1) cap = 1;
  if (gs > cap) gs = (int)cap;
  if (gs > 64) gs = 64;
  return gs;
}

int launch_grad_tc(const float* e_near, const float* e_far, const float* phi_near, const float* sres, float* gpart, int gsplit,
                   int64_t B, int Ds, int L, cudaStream_t st, float* dbg, bool squared) {
  GradTcArgs g;
  g.e_near = e_near; g.e_far = e_far; g.phi_near = phi_near; g.sres = sres; g.gpart = gpart;
  g.B = B; g.Ds = Ds; g.L = L; g.R = 2 * Ds; g.Cc = 2 * L * Ds; g.dbg = dbg;
  int64_t per = ceil_div64(B, gsplit);
  g.imgs_per_split = ceil_div64(per, TC_KCH) * TC_KCH;
  static bool attr_set = false;
  if (!attr_set) {
    TRMPS_CUDA_OK(cudaFuncSetAttribute(grad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    TRMPS_CUDA_OK(cudaFuncSetAttribute(grad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((unsigned)((g.Cc + TC_TN - 1) / TC_TN), (unsigned)((g.R + TC_TM - 1) / TC_TM), (unsigned)gsplit);
  if (squared) grad_tc_kernel<true><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(g);
  else grad_tc_kernel<false><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(g);
  TRMPS_LAUNCH_OK("grad_tc_kernel");
  return TRMPS_OK;
}

}  // namespace trmps
