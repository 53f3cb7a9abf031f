// latency of rotation_params (dependent chain) -- development probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ bool rotation_params(float al, float be, float ga, float tol, float& c, float& s, float& tau, float& relg) {
  const float big = fmaxf(al, be), small = fminf(al, be);
  relg = fabsf(ga) * rsqrtf(al) * rsqrtf(be);
  const bool rot = small > 1e-30f && small > 1e-24f * big && relg > tol;
  // scale by a power of two near 1/max(al,be) (exponent arithmetic, no special-function op) so that the squares
  // below neither underflow for tiny rows nor overflow for huge ones
  const int ebig = (__float_as_int(big) >> 23) & 0xff;
  const float scale = __int_as_float(max(1, min(253, 254 - ebig)) << 23);
  const float d = (be - al) * scale, gs = ga * scale;
  const float rh = rsqrtf(fmaxf(fmaf(d, d, 4.f * gs * gs), 1e-37f));   // 1 / hyp
  const float c2 = fabsf(d) * rh;                               // cos 2t  (>= 0)
  const float x = fmaf(0.5f, c2, 0.5f);                         // cos^2 t in [0.5, 1]
  const float rc = rsqrtf(x);                                   // 1 / c
  const float cc = x * rc;
  const float sg = (d >= 0.f) ? 1.f : -1.f;
  const float ss = sg * gs * rh * rc;                           // sin 2t / (2 c), sin 2t = sg * 2 ga / hyp
  c = rot ? cc : 1.f;
  s = rot ? ss : 0.f;
  tau = rot ? __fdividef(ss, 1.f + cc) : 0.f;
  relg = rot ? relg : 0.f;
  return rot;
}


__global__ void k(float* out, long long* cyc, float a0, float b0, float g0) {
  float al = a0 + threadIdx.x * 1e-3f, be = b0, ga = g0;
  float c, s, tau, relg;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
    rotation_params(al, be, ga, 3e-6f, c, s, tau, relg);
    al = al + c * 1e-3f; be = be + s * 1e-3f; ga = ga + tau * 1e-3f;      // dependent
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; }
  out[threadIdx.x] = al + be + ga + relg;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
  for (int threads : {32, 256}) {
    k<<<1, threads>>>(out, cyc, 2.0f, 1.0f, 0.3f); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("rotation_params dependent chain, %d threads: %.1f cycles per call\n", threads, h / 256.0);
  }
  return 0;
}
