// Helpers shared by the two split implementations (svd.cu: Jacobi on the rows of the bond matrix itself;
// split_factor.cu: Jacobi on the Cholesky factor of its Gram matrix).
#pragma once
#include "common.cuh"

namespace trmps {

__device__ __forceinline__ int compact_to_row(int p, int dn, int Ds) { return (p / dn) * Ds + (p % dn); }

// round-robin pairing of n players (n even): pair #j in step s
__host__ __device__ constexpr int rr_first(int n, int s, int j) { return j == 0 ? n - 1 : (s + j) % (n - 1); }
__host__ __device__ constexpr int rr_second(int n, int s, int j) { return j == 0 ? s : (s + n - 1 - j) % (n - 1); }

__device__ __forceinline__ float rsqrt_fast(float x) {      // one MUFU.RSQ, no denormal fix-up (callers keep x in range)
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// rotation that annihilates the (x,y) entry of a 2x2 Gram [[al, ga],[ga, be]]: x' = c x - s y, y' = s x + c y.
// Half-angle form on the NORMALISED Gram (divide by sqrt(al be): no exponent juggling, nothing can over- or
// underflow while the pair is rotated at all): cos 2t = |d| / hyp, sin 2t = sgn * 2|g| / hyp with
// d = (be - al) / sqrt(al be), g = ga / sqrt(al be), c = sqrt((1 + cos 2t)/2), s = sin 2t / (2c).
// The dependent chain is rsqrt -> 2 mul -> fma -> rsqrt -> mul -> fma -> rsqrt -> 2 mul (~110 cycles; the previous
// form with exponent arithmetic and IEEE-range rsqrtf took 250, tools/probe/rot_probe.cu).
__device__ __forceinline__ bool rotation_params(float al, float be, float ga, float tol, float& c, float& s, float& tau, float& relg) {
  const float big = fmaxf(al, be), small = fminf(al, be);
  const float rr = rsqrt_fast(al) * rsqrt_fast(be);
  const float g = ga * rr;                                      // cosine of the angle between the rows
  relg = fabsf(g);
  const bool rot = small > 1e-30f && small > 1e-24f * big && relg > tol;
  const float d = (be - al) * rr;                               // |d| <= sqrt(big / small) <= 1e12 when rot
  const float rh = rsqrt_fast(fmaf(d, d, 4.f * g * g));        // 1 / hyp
  const float c2 = fabsf(d) * rh;                               // cos 2t  (>= 0)
  const float x = fmaf(0.5f, c2, 0.5f);                         // cos^2 t in [0.5, 1]
  const float rc = rsqrt_fast(x);                               // 1 / c
  const float cc = x * rc;
  const float ss = ((d >= 0.f) ? g : -g) * rh * rc;             // sin 2t / (2 c), sin 2t = sgn * 2 g / hyp
  c = rot ? cc : 1.f;
  s = rot ? ss : 0.f;
  tau = rot ? __fdividef(ss, 1.f + cc) : 0.f;
  relg = rot ? relg : 0.f;
  return rot;
}

}  // namespace trmps
