// Latency probe for the pieces of one Jacobi step (one warp, dependent repetitions): shuffle reduction, angle, apply.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o jac_probe jac_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float rsqrt_fast(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_fast(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr int REP = 256;
__global__ void probe(float* out, long long* cyc, float seed) {
  const int lane = threadIdx.x & 31;
  float v = seed + lane * 1e-3f;
  long long t0, t1;
  // (a) 5-level butterfly sum, dependent
  t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < REP; ++r) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    v *= 0.03f;
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = (t1 - t0) / REP;
  // (b) one shuffle + add, dependent
  t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < REP; ++r) { v += __shfl_xor_sync(0xffffffffu, v, 1); v *= 0.5f; }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = (t1 - t0) / REP;
  // (c) rsqrt chain
  float w = fabsf(v) + 1.5f;
  t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < REP; ++r) { w = rsqrt_fast(w) + 1.25f; }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = (t1 - t0) / REP;
  // (d) rcp chain
  t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < REP; ++r) { w = rcp_fast(w) + 1.25f; }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = (t1 - t0) / REP;
  // (e) fma chain
  t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < REP; ++r) { w = fmaf(w, 0.999f, 0.001f); w = fmaf(w, 0.999f, 0.001f); w = fmaf(w, 0.999f, 0.001f); w = fmaf(w, 0.999f, 0.001f); }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = (t1 - t0) / REP;
  // (f) full angle chain as in jacobi_angle
  float al = 1.f + w, be = 2.f + v * 1e-3f, ga = 0.3f;
  t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < REP; ++r) {
    const float rr = rsqrt_fast(al) * rsqrt_fast(be);
    const float d = (be - al) * rr, g = ga * rr;
    const float rh = rsqrt_fast(fmaf(d, d, 4.f * g * g));
    const float c2 = fabsf(d) * rh, xx = fmaf(0.5f, c2, 0.5f), rc = rsqrt_fast(xx), cc = xx * rc;
    const float s1 = ((d >= 0.f) ? g : -g) * rh * rc;
    const float tau = s1 * rcp_fast(1.f + cc);
    ga = fmaf(tau, 0.1f, 0.3f) + s1 * 1e-3f;          // feed back
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = (t1 - t0) / REP;
  // (g) smem float4 load -> store dependent round trip
  __shared__ float4 sm[64];
  sm[lane] = make_float4(v, w, ga, 1.f); sm[lane + 32] = sm[lane];
  __syncwarp();
  t0 = clock64();
  float4 q = sm[lane];
#pragma unroll 1
  for (int r = 0; r < REP; ++r) { q = sm[(lane + (int)q.w) & 63]; q.w = 1.f; }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = (t1 - t0) / REP;
  // (h) __syncthreads round
  t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < REP; ++r) { __syncthreads(); }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[7] = (t1 - t0) / REP;
  out[threadIdx.x] = v + w + ga + q.x;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
  for (int nt : {32, 256, 512}) {
    probe<<<1, nt>>>(out, cyc, 1.0f); cudaDeviceSynchronize();
    probe<<<1, nt>>>(out, cyc, 1.0f); cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    printf("threads %3d: 5-level shfl-sum %lld | shfl+add+mul %lld | rsqrt+add %lld | rcp+add %lld | 4 fma %lld | angle chain %lld | LDS.128 dep %lld | syncthreads %lld\n",
           nt, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
  }
  return 0;
}
