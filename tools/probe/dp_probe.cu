// FP64 latency/throughput probe: dependent DFMA chain, F2F round trip, 8 independent chains.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int REP = 512;
__global__ void probe(double* out, long long* cyc, double seed) {
  double v = seed + threadIdx.x * 1e-3;
  long long t0, t1;
  t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < REP; ++r) { v = fma(v, 0.999, 0.001); v = fma(v, 0.999, 0.001); v = fma(v, 0.999, 0.001); v = fma(v, 0.999, 0.001); }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = (t1 - t0) / REP;
  t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < REP; ++r) { float f = (float)v; f = f * 0.5f + 0.25f; v = (double)f; }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = (t1 - t0) / REP;
  double a[8];
  for (int k = 0; k < 8; ++k) a[k] = v + k;
  t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < REP; ++r) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fma(a[k], 0.999, 0.001);
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = (t1 - t0) / REP;
  t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < REP; ++r) { v = rsqrt(v + 1.5); }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = (t1 - t0) / REP;
  for (int k = 0; k < 8; ++k) v += a[k];
  out[threadIdx.x] = v;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8192); cudaMalloc(&cyc, 64);
  for (int nt : {32, 256, 1024}) {
    probe<<<1, nt>>>(out, cyc, 1.0); cudaDeviceSynchronize();
    long long h[4]; cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
    printf("threads %4d: 4 dependent DFMA %lld | F2F.F32.F64 + FFMA + F2F.F64.F32 %lld | 8 independent DFMA %lld | rsqrt(double)+add %lld\n", nt, h[0], h[1], h[2], h[3]);
  }
  return 0;
}
