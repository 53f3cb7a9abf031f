// Standalone tcgen05 probe: one CTA, one or more tf32 MMAs on hand-filled smem tiles, prints mismatches.
// usage: tc_probe <variant>   (see main)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Params {
  int a_major, b_major;        // 0 = K-major, 1 = MN-major
  int layout;                  // UMMA layout type: 0 none, 2 = SW128
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  int N;                       // 128 or 256
  int use_fence;               // fence.proxy.async before MMA
};

// operand element value
__host__ __device__ inline float aval(int m, int k) { return (m % 8 == k) ? 1.0f + 0.001f * m : 0.0f; }
__host__ __device__ inline float bval(int n, int k) { return (float)(n + 1) + 1000.0f * k; }

// MN-major SW128 placement: element (mn, k) -> byte offset (k in 0..7)
__device__ inline uint32_t off_mn_sw128(int mn, int k) {
  int blk = mn >> 5, chunk = (mn & 31) >> 2, e = mn & 3;
  return blk * 1024 + k * 128 + ((chunk ^ k) << 4) + e * 4;
}
// MN-major SWIZZLE_128B_BASE32B (layout type 1): atom = 4 k-rows x 128 B, 32-byte units XORed with the row index
__device__ inline uint32_t off_mn_sw128_32b(int mn, int k, uint32_t lbo, uint32_t sbo) {
  int blk = mn >> 5, c32 = (mn & 31) >> 3, e = mn & 7, kg = k >> 2, k4 = k & 3;
  return blk * lbo + kg * sbo + k4 * 128 + ((c32 ^ k4) << 5) + e * 4;
}
// K-major SW128 placement (tf32, K = 8 occupies 32 B of the 128 B row): element (mn, k): rows of 128 B, 8 rows per atom
__device__ inline uint32_t off_k_sw128(int mn, int k) {
  int atom = mn >> 3, r = mn & 7, chunk = k >> 2, e = k & 3;
  return atom * 1024 + r * 128 + ((chunk ^ r) << 4) + e * 4;
}
// K-major no-swizzle (INTERLEAVE): core matrix = 8 rows x 16 B; K = 8 -> 2 core matrices along K (LBO apart), M blocks SBO apart
__device__ inline uint32_t off_k_none(int mn, int k, uint32_t lbo, uint32_t sbo) {
  int mb = mn >> 3, r = mn & 7, kc = k >> 2, e = k & 3;
  return mb * sbo + kc * lbo + r * 16 + e * 4;
}
// MN-major no-swizzle: core = 4 mn (16 B) x 8 k rows of 16 B; MN blocks of 4 at SBO, (K has one group of 8)
__device__ inline uint32_t off_mn_none(int mn, int k, uint32_t lbo, uint32_t sbo) {
  int mb = mn >> 2, e = mn & 3;
  return mb * sbo + k * 16 + e * 4;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(Params p, float* out, uint32_t* info) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;             // up to 16 KB
  uint8_t* sb = smem + 16384;     // up to 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 49152 + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 49152 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  __syncthreads();
  for (int e = tid; e < 128 * 8; e += 128) {
    int m = e / 8, k = e % 8;
    uint32_t o = p.a_major ? (p.layout == 1 ? off_mn_sw128_32b(m, k, p.a_lbo, p.a_sbo) : p.layout == 2 ? off_mn_sw128(m, k) : off_mn_none(m, k, p.a_lbo, p.a_sbo))
                           : (p.layout == 2 ? off_k_sw128(m, k) : off_k_none(m, k, p.a_lbo, p.a_sbo));
    *reinterpret_cast<float*>(sa + o) = aval(m, k);
  }
  for (int e = tid; e < p.N * 8; e += 128) {
    int n = e / 8, k = e % 8;
    uint32_t o = p.b_major ? (p.layout == 1 ? off_mn_sw128_32b(n, k, p.b_lbo, p.b_sbo) : p.layout == 2 ? off_mn_sw128(n, k) : off_mn_none(n, k, p.b_lbo, p.b_sbo))
                           : (p.layout == 2 ? off_k_sw128(n, k) : off_k_none(n, k, p.b_lbo, p.b_sbo));
    *reinterpret_cast<float*>(sb + o) = bval(n, k);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (p.use_fence) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = *slot;
  if (tid == 0) {
    info[0] = tbase;
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_major << 15) | ((uint32_t)p.b_major << 16) |
                     ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto mk = [&](uint32_t addr, uint32_t lbo, uint32_t sbo) {
      uint64_t d = 0;
      d |= (uint64_t)((addr >> 4) & 0x3FFF);
      d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
      d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)p.layout << 61;
      return d;
    };
    uint64_t da = mk(smem_u32(sa), p.a_lbo, p.a_sbo), db = mk(smem_u32(sb), p.b_lbo, p.b_sbo);
    info[1] = idesc; info[2] = (uint32_t)da; info[3] = (uint32_t)(da >> 32);
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, q;\n\t}"
                 ::"r"(tbase), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  // wait
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(0u) : "memory");
    if (clock64() - t0 > 2000000000LL) { if (tid == 0) info[4] = 0xdead; break; }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // read back: warp w reads lanes 32w..32w+31
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
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256u) : "memory");
}

int run(const char* name, Params p) {
  float* out; uint32_t* info;
  cudaMalloc(&out, 128 * 256 * 4); cudaMalloc(&info, 64);
  cudaMemset(out, 0xff, 128 * 256 * 4); cudaMemset(info, 0, 64);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 51200 + 1024);
  probe_kernel<<<1, 128, 51200 + 1024>>>(p, out, info);
  cudaError_t e = cudaDeviceSynchronize();
  static float h[128 * 256]; uint32_t hi[16];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(hi, info, 64, cudaMemcpyDeviceToHost);
  int bad = 0, zeros = 0; double maxerr = 0;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < p.N; ++n) {
    float want = 0; for (int k = 0; k < 8; ++k) want += aval(m, k) * bval(n, k);
    float got = h[m * 256 + n];
    double err = fabs((double)got - want) / fabs(want);
    if (!(err < 2e-3)) ++bad;
    if (got == 0.0f) ++zeros;
    if (err > maxerr) maxerr = err;
  }
  printf("%-34s err=%s tmem=0x%08x idesc=0x%08x timeout=%x  bad=%d/%d zeros=%d maxrel=%.3g  D[0][0..3]=%g %g %g %g  D[1][0]=%g D[9][5]=%g (want %g %g %g)\n",
         name, cudaGetErrorString(e), hi[0], hi[1], hi[4], bad, 128 * p.N, zeros, maxerr, h[0], h[1], h[2], h[3], h[256], h[9 * 256 + 5],
         aval(0, 0) * bval(0, 0), aval(1, 1) * bval(0, 1), aval(9, 1) * bval(5, 1));
  cudaFree(out); cudaFree(info);
  return e != cudaSuccess;
}

int main(int argc, char** argv) {
  int v = argc > 1 ? atoi(argv[1]) : -1;
  struct { const char* name; Params p; } cases[] = {
    {"K-major none  lbo128 sbo256 N128", {0, 0, 0, 128, 256, 128, 256, 128, 1}},
    {"K-major none  lbo256 sbo128 N128", {0, 0, 0, 256, 128, 256, 128, 128, 1}},
    {"K-major sw128 lbo16 sbo1024 N128", {0, 0, 2, 16, 1024, 16, 1024, 128, 1}},
    {"K-major sw128 lbo0  sbo1024 N256", {0, 0, 2, 0, 1024, 0, 1024, 256, 1}},
    {"MN-major sw128 lbo1024 sbo4096 N128", {1, 1, 2, 1024, 4096, 1024, 4096, 128, 1}},
    {"MN-major sw128 lbo4096 sbo1024 N128", {1, 1, 2, 4096, 1024, 4096, 1024, 128, 1}},
    {"MN-major sw128 lbo1024 sbo8192 N256", {1, 1, 2, 1024, 8192, 1024, 8192, 256, 1}},
    {"MN-major sw128 N256 nofence", {1, 1, 2, 1024, 8192, 1024, 8192, 256, 0}},
    {"MN-major none lbo128 sbo128 N128", {1, 1, 0, 128, 128, 128, 128, 128, 1}},
    {"MN sw128_32B lbo512 sbo2048 N128", {1, 1, 1, 512, 2048, 512, 2048, 128, 1}},
    {"MN sw128_32B lbo1024 sbo512 N128", {1, 1, 1, 1024, 512, 1024, 512, 128, 1}},
    {"MN sw128_32B A512/4096 B512/4096 N256", {1, 1, 1, 512, 4096, 512, 4096, 256, 1}},
    {"A MN sw128_32B, B K-major? (mixed) N128", {1, 0, 1, 512, 2048, 16, 1024, 128, 1}},
  };
  int n = sizeof(cases) / sizeof(cases[0]);
  for (int i = 0; i < n; ++i) {
    if (v >= 0 && v != i) continue;
    if (run(cases[i].name, cases[i].p)) { printf("CUDA error, stopping\n"); return 1; }
  }
  return 0;
}
