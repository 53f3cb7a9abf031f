// Standalone tcgen05 probe for kind::f16 (fp16 operands, fp32 accumulate), K-major SWIZZLE_128B operands:
//   1. correctness of D = A B^T for M=128, K=64 (4 MMAs of K=16), N in {64,128,256}
//   2. issue-to-completion cycles per MMA for kind::f16 (K=16) and kind::tf32 (K=8) at those N
// usage: f16_probe
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__host__ __device__ inline float aval(int m, int k) { return (float)((m * 7 + k * 3) % 17 - 8) * 0.125f; }
__host__ __device__ inline float bval(int n, int k) { return (float)((n * 5 + k * 11) % 13 - 6) * 0.25f; }

// K-major SW128, 2-byte elements: rows of 128 B (64 halfs), 8 rows per 1024 B atom, 16 B chunks XORed with row&7
__device__ inline uint32_t off_k_sw128_h(int r, int k) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + (k & 7) * 2);
}

// K-major SW64, 2-byte elements: rows of 64 B (32 halfs), 8 rows per 512 B atom, 16 B chunks XORed with (row>>1)&3
__device__ inline uint32_t off_k_sw64_h(int r, int k) {
  return (uint32_t)((r >> 3) * 512 + (r & 7) * 64 + ((((k >> 3) ^ (r >> 1)) & 3) << 4) + (k & 7) * 2);
}

struct Params { int N; int reps; int kind; int nacc; int b64; };   // b64: B operand in K-major SWIZZLE_64B (64-byte rows, K = 32)   // kind 0 = f16, 1 = tf32 (timing only)

__global__ void __launch_bounds__(128, 1) probe_kernel(Params p, float* out, long long* cyc) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;             // 128 rows x 128 B = 16 KB
  uint8_t* sb = smem + 16384;     // 256 rows x 128 B = 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 49152 + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 49152 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  __syncthreads();
  for (int e = tid; e < 128 * 64; e += 128) { int m = e / 64, k = e % 64; *reinterpret_cast<__half*>(sa + off_k_sw128_h(m, k)) = __float2half(aval(m, k)); }
  if (p.b64) { for (int e = tid; e < p.N * 32; e += 128) { int n = e / 32, k = e % 32; *reinterpret_cast<__half*>(sb + off_k_sw64_h(n, k)) = __float2half(bval(n, k)); } }
  else for (int e = tid; e < p.N * 64; e += 128) { int n = e / 64, k = e % 64; *reinterpret_cast<__half*>(sb + off_k_sw128_h(n, k)) = __float2half(bval(n, k)); }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = *slot;
  auto mk = [&](uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((16u >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
  };
  auto mk64 = [&](uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((512u >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
  };
  long long t0 = 0;
  if (tid == 0) {
    const uint32_t idesc_h = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_t = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (p.reps == 1) {     // correctness pass: 4 MMAs, first one overwrites
      for (int ks = 0; ks < (p.b64 ? 2 : 4); ++ks) {
        const uint64_t xa = mk(smem_u32(sa) + ks * 32), xb = p.b64 ? mk64(smem_u32(sb) + ks * 32) : mk(smem_u32(sb) + ks * 32);
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}"
                     ::"r"(tbase), "l"(xa), "l"(xb), "r"(idesc_h), "r"(ks > 0 ? 1u : 0u) : "memory");
      }
    }
    uint64_t da[4], db[4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) { da[ks] = mk(smem_u32(sa) + ks * 32); db[ks] = p.b64 ? mk64(smem_u32(sb) + (ks & 1) * 32 + (ks >> 1) * 16384) : mk(smem_u32(sb) + ks * 32); }
    const uint32_t idesc = p.kind == 0 ? idesc_h : idesc_t;
    const uint32_t amask = (uint32_t)(p.nacc - 1), nn = (uint32_t)p.N;
    t0 = clock64();
    if (p.kind == 0) {
#pragma unroll 1
      for (int rep = 0; rep < (p.reps == 1 ? 0 : p.reps); ++rep) {
        const uint32_t td = tbase + ((uint32_t)rep & amask) * nn;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}"
                       ::"r"(td), "l"(da[ks]), "l"(db[ks]), "r"(idesc), "r"(1u) : "memory");
      }
    } else {
#pragma unroll 1
      for (int rep = 0; rep < (p.reps == 1 ? 0 : p.reps); ++rep) {
        const uint32_t td = tbase + ((uint32_t)rep & amask) * nn;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, q;\n\t}"
                       ::"r"(td), "l"(da[ks]), "l"(db[ks]), "r"(idesc), "r"(1u) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  uint32_t ok = 0;
  long long tw = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(0u) : "memory");
    if (clock64() - tw > 2000000000LL) break;
  }
  if (tid == 0) cyc[0] = clock64() - t0;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int cb = 0; cb < p.N / 8; ++cb) {
    uint32_t r[8];
    uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16) + cb * 8;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) out[(warp * 32 + lane) * 256 + cb * 8 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
}

int run(int N, int reps, int kind, bool check, int nacc = 1, int b64 = 0) {
  float* out; long long* cyc;
  cudaMalloc(&out, 128 * 256 * 4); cudaMalloc(&cyc, 64);
  cudaMemset(out, 0xff, 128 * 256 * 4); cudaMemset(cyc, 0, 64);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 51200 + 1024);
  Params p{N, reps, kind, nacc, b64};
  probe_kernel<<<1, 128, 51200 + 1024>>>(p, out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  static float h[128 * 256]; long long hc[8];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(hc, cyc, 64, cudaMemcpyDeviceToHost);
  if (check) {
    int bad = 0; double maxerr = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
      double want = 0; for (int k = 0; k < (b64 ? 32 : 64); ++k) want += (double)aval(m, k) * bval(n, k);
      double err = fabs((double)h[m * 256 + n] - want);
      if (!(err < 1e-3)) ++bad;
      if (err > maxerr) maxerr = err;
    }
    printf("f16 K-major %s  N=%3d: err=%s bad=%d/%d maxabs=%.3g  D[0][0..2]=%g %g %g\n", b64 ? "A SW128 / B SW64" : "SW128", N, cudaGetErrorString(e), bad, 128 * N, maxerr, h[0], h[1], h[2]);
  } else {
    const int nm = reps * 4;
    const double macs = (double)nm * 128 * N * (kind == 0 ? 16 : 8);
    printf("%s%s nacc=%d N=%3d: %d MMAs in %lld cycles -> %.1f cycles/MMA, %.0f MAC/clk/SM  (%s)\n", kind == 0 ? "f16 " : "tf32", b64 ? " (B SW64)" : "", nacc, N, nm, hc[0],
           (double)hc[0] / nm, macs / (double)hc[0], cudaGetErrorString(e));
  }
  cudaFree(out); cudaFree(cyc);
  return e != cudaSuccess;
}

int main() {
  for (int N : {64, 128, 256}) if (run(N, 1, 0, true)) return 1;
  for (int N : {128, 256}) if (run(N, 1, 0, true, 1, 1)) return 1;
  for (int kind : {0, 1}) for (int N : {64, 128, 256}) if (run(N, 256, kind, false)) return 1;
  for (int kind : {0, 1}) for (int N : {64, 128}) for (int na : {2, 4}) if (run(N, 256, kind, false, na)) return 1;
  for (int kind : {0, 1}) if (run(256, 256, kind, false, 2)) return 1;
  for (int N : {128, 256}) if (run(N, 256, 0, false, 1, 1)) return 1;
  for (int N : {144, 160, 176, 192, 208, 224, 240}) if (run(N, 256, 0, false)) return 1;
  return 0;
}
