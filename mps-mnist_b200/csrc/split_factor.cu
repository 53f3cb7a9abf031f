// Truncated bond split on a small factor that never leaves the chip  (north_star piece 4, second implementation).
// Replaces MPSOptimizer._bond_decomposition (optimizer.py:266-321) like svd.cu does, but instead of rotating the
// long rows of the bond matrix M (R x Cc, Cc = L * R) it works on an n x n factor, n = rows in use <= 256:
//
//   1. split_gram_kernel     G = M M^T in FP64 (the products of two floats are exact in double; every CTA owns a
//                            64 x 64 tile of G and a slice of the columns, the last CTA to finish a tile adds the
//                            slices in a fixed order -- no floating-point atomics, bit-identical on every rank)
//   2. split_chol_kernel     G = R^T R, blocked left-looking Cholesky in FP64 on one SM; R is rounded to FP32,
//                            scaled by a power of two so that its largest row has norm ~1, rows ranked by norm
//   3. split_jacobi_kernel   one-sided (Hestenes) Jacobi on the ROWS of R inside ONE thread-block cluster: every CTA
//                            holds two complete blocks of 16 rows in shared memory (a row pair is rotated by one warp:
//                            dot products and norms are warp-local, nothing is reduced across CTAs), blocks move
//                            between CTAs through distributed shared memory with one cluster barrier per round
//                            (round-robin tournament over blocks).  R R^T is much closer to diagonal than
//                            G (Drmac-Veselic preconditioning), and no rotations have to be accumulated:
//                            J^T R = diag(sigma) W^T with R^T R = G means the rows of W^T -- the converged rows
//                            divided by their norms -- are the eigenvectors of G, i.e. the left singular vectors of M.
//   4. split_rank_kernel     sigma sorted, m = clamp(#{sigma > min_sv}, min_size, max_size)   (optimizer.py:294-302)
//   5. split_project_kernel  a_j1 = U[:, :m]^T M written straight into the special-node layout, a_j = U[:, :m] into
//                            the node slot (optimizer.py:305-317 and the transposes at :180, :125-126)
//
// sigma_p is the norm of row p of U^T M.  Nothing is divided by a small sigma except the normalisation of rows that
// are numerically zero, which are set to zero like utils.check_nan does with the NaNs of tf.svd (utils.py:9-20).
// None of the kernels is a cooperative launch, so ncu can replay each of them.
#include <cooperative_groups.h>
#include "common.cuh"
#include "svd_common.cuh"

namespace cg = cooperative_groups;

namespace trmps {

// ------------------------------------------------------------------------------------------------ workspace
constexpr int SP_KS_MAX = 16;     // column slices of the Gram product
constexpr int SP_NMAX = 256;      // rows of the factor

struct SplitWs {
  double* gp;      // [SP_KS_MAX][R][R] Gram partials
  double* G;       // [R][R] Gram matrix (lower triangle used)
  double* Rd;      // [R][R] Cholesky factor, row k = row k of R (upper triangle used)
  float* Rf;       // [R][R] R in FP32 / scale
  float* X;        // [R + 2 JB][R] rows after the Jacobi iteration, normalised (by slot)
  float* sig;      // [SP_NMAX] row norms after the Jacobi iteration (by slot), then sigma
  float* scale;    // [4]: scale of Rf, largest diagonal of G
  int32_t* cnt;    // [64] tile arrival counters (zero between launches)
  int32_t* rowmap; // [SP_NMAX] rank -> row of R
};

int64_t split_ws_floats(int R) {
  const int64_t RR = (int64_t)R * R;
  return 2 * (SP_KS_MAX * RR) + 2 * RR + 2 * RR + RR + (int64_t)(R + 32) * R + SP_NMAX + 4 + 64 + SP_NMAX + 16;
}

static SplitWs split_ws(float* base, int R) {
  const int64_t RR = (int64_t)R * R;
  SplitWs w;
  double* d = reinterpret_cast<double*>(base);
  w.gp = d; d += SP_KS_MAX * RR;
  w.G = d; d += RR;
  w.Rd = d; d += RR;
  float* f = reinterpret_cast<float*>(d);
  w.Rf = f; f += RR;
  w.X = f; f += (int64_t)(R + 32) * R;
  w.sig = f; f += SP_NMAX;
  w.scale = f; f += 4;
  w.cnt = reinterpret_cast<int32_t*>(f); f += 64;
  w.rowmap = reinterpret_cast<int32_t*>(f);
  return w;
}

// ------------------------------------------------------------------------------------------------ 1. Gram matrix
constexpr int GT = 64, GK = 32, GLD = GT + 2;

struct GramArgs {
  const float* M; int R, Cc, Ds; const int32_t* d_near; int ks_cols, KS; double* gp; double* G; int32_t* cnt;
  int finalize;      // 0: leave the column-slice partials in gp (the cluster Cholesky adds them while it loads its blocks)
};

__global__ void __launch_bounds__(256) split_gram_kernel(GramArgs a) {
  __shared__ __align__(16) double As[GK][GLD];
  __shared__ __align__(16) double Bs[GK][GLD];
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int dn = *a.d_near, n = 2 * dn;
  // lower-triangular tile (ti >= tj) from the linear tile index
  int ti = 0, rem = blockIdx.x;
  while (rem > ti) { rem -= ti + 1; ++ti; }
  const int tj = rem;
  if (ti * GT >= n) return;                    // tile outside the rows in use: the factorisation never reads it
  const int ks = blockIdx.y;
  const int c_begin = ks * a.ks_cols, c_end = min(a.Cc, c_begin + a.ks_cols);
  const int tx = tid & 15, ty = tid >> 4;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  const int lrow = tid >> 2, lcg = (tid & 3) * 8;
  const int pa = ti * GT + lrow, pb = tj * GT + lrow;
  const float* rowa = pa < n ? a.M + (size_t)compact_to_row(pa, dn, a.Ds) * a.Cc : nullptr;
  const float* rowb = pb < n ? a.M + (size_t)compact_to_row(pb, dn, a.Ds) * a.Cc : nullptr;
  const bool diag = ti == tj;
  for (int c0 = c_begin; c0 < c_end; c0 += GK) {
    float va[8], vb[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = c0 + lcg + 4 * h;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
      if (c < c_end) {
        if (rowa) x = __ldg(reinterpret_cast<const float4*>(rowa + c));
        if (!diag && rowb) y = __ldg(reinterpret_cast<const float4*>(rowb + c));
      }
      va[4 * h] = x.x; va[4 * h + 1] = x.y; va[4 * h + 2] = x.z; va[4 * h + 3] = x.w;
      vb[4 * h] = y.x; vb[4 * h + 1] = y.y; vb[4 * h + 2] = y.z; vb[4 * h + 3] = y.w;
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      As[lcg + e][lrow] = (double)va[e];
      if (!diag) Bs[lcg + e][lrow] = (double)vb[e];
    }
    __syncthreads();
    const double (*Bp)[GLD] = diag ? As : Bs;
#pragma unroll 8
    for (int k = 0; k < GK; ++k) {
      const double2 a01 = *reinterpret_cast<const double2*>(&As[k][ty * 4]);
      const double2 a23 = *reinterpret_cast<const double2*>(&As[k][ty * 4 + 2]);
      const double2 b01 = *reinterpret_cast<const double2*>(&Bp[k][tx * 4]);
      const double2 b23 = *reinterpret_cast<const double2*>(&Bp[k][tx * 4 + 2]);
      const double av[4] = {a01.x, a01.y, a23.x, a23.y}, bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
  }
  // partial tile of this column slice
  const int R = a.R;
  double* gp = a.gp + (size_t)ks * R * R;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = ti * GT + ty * 4 + i;
    if (p < R) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = tj * GT + tx * 4 + j;
        if (q < R) __stcg(&gp[(size_t)p * R + q], acc[i][j]);
      }
    }
  }
  if (!a.finalize) return;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int old = atomicAdd(&a.cnt[blockIdx.x], 1);
    s_last = old == a.KS - 1;
    if (s_last) a.cnt[blockIdx.x] = 0;          // ready for the next launch
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last CTA of this tile: add the slices in a fixed order (all slices of a row of four entries are loaded first, so
  // the L2 round trips overlap)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = ti * GT + ty * 4 + i;
    const int q0 = tj * GT + tx * 4;
    if (p < R && q0 < R) {                       // R is a multiple of 4: a row of four entries is in or out as a whole
      double2 v[SP_KS_MAX][2];
#pragma unroll
      for (int k = 0; k < SP_KS_MAX; ++k) {
        if (k < a.KS) {
          const double2* src = reinterpret_cast<const double2*>(&a.gp[((size_t)k * R + p) * R + q0]);
          v[k][0] = __ldcg(src); v[k][1] = __ldcg(src + 1);
        } else {
          v[k][0] = make_double2(0.0, 0.0); v[k][1] = v[k][0];
        }
      }
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
      for (int k = 0; k < SP_KS_MAX; ++k) { s0 += v[k][0].x; s1 += v[k][0].y; s2 += v[k][1].x; s3 += v[k][1].y; }
      double2* dst = reinterpret_cast<double2*>(&a.G[(size_t)p * R + q0]);
      dst[0] = make_double2(s0, s1); dst[1] = make_double2(s2, s3);
      if (ti != tj) {                              // G is symmetric: store the mirror image too (the factorisation reads columns as rows)
        a.G[(size_t)q0 * R + p] = s0; a.G[(size_t)(q0 + 1) * R + p] = s1;
        a.G[(size_t)(q0 + 2) * R + p] = s2; a.G[(size_t)(q0 + 3) * R + p] = s3;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ 2. Cholesky
// Left-looking, panels of CW columns, one SM.  Rd[k][i] (i >= k) is stored row-wise, so thread i reads the entries of
// column i of all earlier rows with unit stride.  Per panel:
//   update   P[i][c] = G[i][c0+c] - sum_{k<c0} Rd[k][i] Rd[k][c0+c]     all 1024 threads (k split four ways), panel in registers
//   diagonal the 16 x 16 diagonal block is factorised by ONE warp in registers (lane = row, values exchanged by shuffles):
//            the only sequential part, 16 dependent steps without a block barrier
//   solve    every row below does its own forward substitution with the factor of the diagonal block (no barrier)
// (A version that kept the panel in shared memory and took a block barrier per column spent 1260 cycles per column on
// shared-memory traffic: 600 wavefronts per column.)
// A pivot that is not safely positive (rank-deficient bond: the reference's initial MPS has many exactly dependent rows) is
// replaced by a tiny one and its column is dropped: that row of R then carries a singular value ~1e-7 sigma_max, i.e.
// zero in FP32 terms.
constexpr int CW = 16, CT = 256, CNE = CW * SP_NMAX / CT;   // CNE: panel entries per thread

struct CholArgs {
  const double* G; double* Rd; float* Rf; int R; const int32_t* d_near; float* scale; int32_t* rowmap; float* dbg;
};

// 1/sqrt(x) for x in (1e-14, 2]: MUFU seed in single precision + one Newton step in double (relative error < 1e-13)
__device__ __forceinline__ double rsqrt_seeded(double x) {
  const double y = (double)rsqrt_fast((float)x);
  return y * fma(-0.5 * x * y, y, 1.5);
}

__global__ void __launch_bounds__(CT) split_chol_kernel(CholArgs a) {
  extern __shared__ __align__(16) double csm[];
  double (*D)[CW] = reinterpret_cast<double (*)[CW]>(csm);                                             // D[k][c] = Rd[k][c0 + c], k < c0
  double (*part)[CT] = reinterpret_cast<double (*)[CT]>(csm + CW * SP_NMAX);                           // part[c][thread]: partial sums of the k groups 1..
  double (*Gs)[SP_NMAX] = reinterpret_cast<double (*)[SP_NMAX]>(csm + CW * SP_NMAX + CW * CT);             // Gs[c][i] = G[i][c0 + c]
  __shared__ float nrm_part[CW][SP_NMAX / 32];
  __shared__ __align__(16) double Dg[CW][CW + 1];        // diagonal block before / its factor after
  __shared__ double rinv_s[CW];
  __shared__ double red[CT / 32];
  __shared__ float nrm[SP_NMAX];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = 2 * *a.d_near, R = a.R;
  // largest diagonal entry -> pivot threshold and the power-of-two scale of the FP32 copy
  double dmax = 0.0;
  for (int i = tid; i < n; i += CT) dmax = fmax(dmax, a.G[(size_t)i * R + i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  if (lane == 0) red[warp] = dmax;
  __syncthreads();
  dmax = 0.0;
  for (int w = 0; w < CT / 32; ++w) dmax = fmax(dmax, red[w]);
  const double tolp = (double)n * 2.3e-16 * dmax;
  const double sq_tolp = sqrt(tolp > 0.0 ? tolp : 1e-300), inv_dmax = dmax > 0.0 ? 1.0 / dmax : 0.0, rs_dmax = dmax > 0.0 ? rsqrt(dmax) : 0.0;
  int ex = 0;
  if (dmax > 0.0) frexp(sqrt(dmax), &ex);
  const double inv_scale = ldexp(1.0, -ex);                // Rf = Rd * 2^-ex: largest row norm in [0.5, ~16)
  const bool timer = a.dbg != nullptr && warp == 0;      // (BAR.SYNC defers its block to the next dependent instruction, so a
  long long tacc[6] = {0, 0, 0, 0, 0, 0}, tprev = timer ? clock64() : 0;   //  phase's wait shows up in the NEXT phase's counter)
  auto tick = [&](int k) { if (timer) { const long long now = clock64(); tacc[k] += now - tprev; tprev = now; } };
  for (int c0 = 0; c0 < n; c0 += CW) {
    const int w = min(CW, n - c0);
    // rows c0 .. n-1 of the panel, padded to whole warps; the k range of the update is split over as many thread groups as fit
    const int nr = n - c0, nrp = (nr + 31) & ~31, ngr = CT / nrp;
    const int g = tid / nrp, il = tid - g * nrp, irow = c0 + il;
    const bool upd = g < ngr && il < nr;                   // takes part in the update
    const bool mine = g == 0 && il < nr;                   // owns row irow of the panel
    auto rd_load = [&](int b, double (&r)[8]) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = g + ngr * (8 * b + u);
        r[u] = (upd && k < c0) ? a.Rd[(size_t)k * R + irow] : 0.0;      // L1-cached: a row of R is final once written and is re-read by every later panel
      }
    };
    // ---- columns c0 .. c0+w-1 of G (stored in full: read as rows, unit stride), D[k][c] = Rd[k][c0 + c] and the first batch
    //      of this thread's entries of R: all issued before the first one is used (an L2 round trip is ~800 cycles here)
    double r0[8], r1[8];
    {
      double gv[CNE], dv[CNE];
#pragma unroll
      for (int u = 0; u < CNE; ++u) {
        const int e = tid + u * CT, c = e / SP_NMAX, j = e % SP_NMAX;        // CW * SP_NMAX = CNE * CT entries
        gv[u] = (c < w && j >= c0 && j < n) ? __ldcg(&a.G[(size_t)(c0 + c) * R + j]) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < CNE; ++u) {
        const int e = tid + u * CT, k = e / CW, c = e % CW;
        dv[u] = (e < c0 * CW && c < w) ? a.Rd[(size_t)k * R + c0 + c] : 0.0;
      }
      rd_load(0, r0);
#pragma unroll
      for (int u = 0; u < CNE; ++u) {
        const int e = tid + u * CT;
        Gs[e / SP_NMAX][e % SP_NMAX] = gv[u];
        if (e < c0 * CW) D[e / CW][e % CW] = dv[u];
      }
    }
    __syncthreads();
    tick(0);
    // ---- update: acc[c] = sum over this group's k of Rd[k][irow] Rd[k][c0+c]; the next batch of loads is in flight while
    //      the current one is used
    double acc[CW];
#pragma unroll
    for (int c = 0; c < CW; ++c) acc[c] = 0.0;
    const int nbatch = (c0 + 8 * ngr - 1) / (8 * ngr);
    auto rd_use = [&](int b, const double (&r)[8]) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = min(g + ngr * (8 * b + u), c0 - 1);   // r[u] is zero beyond the range
#pragma unroll
        for (int c = 0; c < CW; ++c) acc[c] = fma(r[u], D[k][c], acc[c]);
      }
    };
    for (int b = 0; b < nbatch; b += 2) {
      rd_load(b + 1, r1);
      rd_use(b, r0);
      rd_load(b + 2, r0);
      rd_use(b + 1, r1);
    }
    if (g > 0 && upd) {
#pragma unroll
      for (int c = 0; c < CW; ++c) part[c][tid] = acc[c];
    }
    __syncthreads();
    tick(1);
    double p[CW];
    if (mine) {
#pragma unroll
      for (int c = 0; c < CW; ++c) {
        double sacc = acc[c];
        for (int gg = 1; gg < ngr; ++gg) sacc += part[c][gg * nrp + il];
        p[c] = Gs[c][irow] - sacc;
      }
      if (il < CW) {
#pragma unroll
        for (int c = 0; c < CW; ++c) Dg[il][c] = p[c];
      }
    }
    __syncthreads();
    tick(2);
    // ---- diagonal block: one warp, lane r holds row r, 16 dependent steps
    if (warp == 0) {
      const int r = lane & 15;
      double d[CW];
#pragma unroll
      for (int c = 0; c < CW; ++c) d[c] = (r < nr) ? Dg[r][c] : 0.0;
#pragma unroll
      for (int c = 0; c < CW; ++c) {
        const double piv = __shfl_sync(0xffffffffu, d[c], c);
        const bool drop = !(piv > tolp) || c >= w;
        const double rinv = drop ? 0.0 : rsqrt_seeded(piv * inv_dmax) * rs_dmax;
        const double lkk = drop ? sq_tolp : piv * rinv;
        const double l = r == c ? lkk : d[c] * rinv;       // l[r][c], meaningful for r >= c
        d[c] = l;
#pragma unroll
        for (int cc = c + 1; cc < CW; ++cc) {
          const double t = __shfl_sync(0xffffffffu, l, cc);   // l[cc][c]
          d[cc] = fma(-l, t, d[cc]);
        }
        if (lane == 0) rinv_s[c] = rinv;
      }
      if (lane < CW) {
#pragma unroll
        for (int c = 0; c < CW; ++c) Dg[r][c] = c <= r ? d[c] : 0.0;
      }
    }
    __syncthreads();
    tick(3);
    // ---- rows below the diagonal block: forward substitution; rows of the block: their row of its factor.  The finished
    //      rows go back into Gs (which already holds zeros left of the panel and beyond n)
    if (mine) {
      if (il >= CW) {
#pragma unroll
        for (int c = 0; c < CW; ++c) {
          const double l = p[c] * rinv_s[c];
          p[c] = l;
#pragma unroll
          for (int cc = c + 1; cc < CW; ++cc) p[cc] = fma(-l, Dg[cc][c], p[cc]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < CW; ++c) p[c] = Dg[il][c];
      }
#pragma unroll
      for (int c = 0; c < CW; ++c) Gs[c][irow] = p[c];
    }
    __syncthreads();
    // ---- rows c0 .. c0+w-1 of R out (unit stride); squared norms of the FP32 rows: warp sums, then the eight warps of a
    //      row in a fixed order
#pragma unroll
    for (int u = 0; u < CNE; ++u) {
      const int e = tid + u * CT, c = e / SP_NMAX, j = e % SP_NMAX;
      float f = 0.f;
      if (c < w && j < R) {
        const double v = Gs[c][j];
        __stcg(&a.Rd[(size_t)(c0 + c) * R + j], v);
        f = (float)(v * inv_scale);
        a.Rf[(size_t)(c0 + c) * R + j] = f;
      }
      float sq = f * f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      if (lane == 0) nrm_part[c][(e % SP_NMAX) / 32] = sq;
    }
    __syncthreads();
    if (tid < w) {
      float sq = 0.f;
#pragma unroll
      for (int q = 0; q < SP_NMAX / 32; ++q) sq += nrm_part[tid][q];
      nrm[c0 + tid] = sq;
    }
    tick(4);
  }
  // ---- ranking of the rows by norm
  __syncthreads();
  if (tid < n) {
    const float sp = nrm[tid];
    int rank = 0;
    for (int o = 0; o < n; ++o) {
      const float so = nrm[o];
      rank += (so > sp) || (so == sp && o < tid);
    }
    a.rowmap[rank] = tid;
  }
  tick(5);
  if (timer && lane == 0) for (int k = 0; k < 6; ++k) a.dbg[k] = (float)tacc[k];
  if (tid == 0) { a.scale[0] = (float)ldexp(1.0, ex); a.scale[1] = (float)dmax; }
}

// ------------------------------------------------------------------------------------------------ 3. Jacobi
constexpr int JB = 16;                 // rows per block
constexpr int JLD = SP_NMAX + 4;       // row stride in shared memory (floats): 16-byte rows, conflict-free float4 access;
                                       // float SP_NMAX of a row carries its squared norm through the rotations
constexpr int JCL_MAX = 16;            // CTAs per cluster = block pairs per round (16: non-portable cluster size)

struct JacArgs {
  const float* Rf; int R; const int32_t* rowmap; const int32_t* d_near; float* X; float* sig; int32_t* iscal;
  float tol, stop; int max_sweeps; int csize; float* dbg;   // dbg: 4 floats, cycles of CTA 0 in (rotations, cluster barrier, -, -)
};

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float s) {
  s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); return fmaf(a.w, b.w, s);
}
// packed FP32 (FFMA2, sm_100): two lanes of a float2 per instruction -- the Jacobi phase is bound by instruction issue
// and latency of a single warp per scheduler, so the long-row arithmetic (inner products, rotations) is issued in pairs
__device__ __forceinline__ float2 dot4_2(const float4& a, const float4& b, float2 s) {
  s = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), s);
  return __ffma2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w), s);
}
__device__ __forceinline__ void rot2(float& x0, float& x1, float& y0, float& y1, float2 sn, float2 nsn, float2 tau, float2 ntau) {
  // Rutishauser's form x' = x - s (y + tau x), y' = y + s (x - tau y): norm-preserving in FP32 even when cos rounds to 1
  const float2 u = make_float2(x0, x1), w = make_float2(y0, y1);
  const float2 xn = __ffma2_rn(nsn, __ffma2_rn(tau, u, w), u);
  const float2 yn = __ffma2_rn(sn, __ffma2_rn(ntau, w, u), w);
  x0 = xn.x; x1 = xn.y; y0 = yn.x; y1 = yn.y;
}
__device__ __forceinline__ void rot4(float4& x, float4& y, float sn, float tau) {
  const float2 s2 = make_float2(sn, sn), ns2 = make_float2(-sn, -sn), t2 = make_float2(tau, tau), nt2 = make_float2(-tau, -tau);
  rot2(x.x, x.y, y.x, y.y, s2, ns2, t2, nt2);
  rot2(x.z, x.w, y.z, y.w, s2, ns2, t2, nt2);
}

// ---- hand-over of the row blocks without a cluster-wide barrier: the rows are written into the next CTA's shared memory
// with st.async, which signals the bytes on an mbarrier in THAT CTA; the receiver waits for its own bytes only (its two
// ring neighbours), not for the slowest CTA of the cluster, and no release fence has to drain (the cluster barrier's
// MEMBAR + wait were 19 % of the kernel's stall samples, profiles/r2f_split_and_env_kernels_ncu.txt)
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, int cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void st_async_f4(uint32_t raddr, const float4& v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(raddr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)), "r"(rbar) : "memory");
}
__device__ __forceinline__ void st_async_f1(uint32_t raddr, float v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(__float_as_uint(v)), "r"(rbar) : "memory");
}
// one bulk copy of a whole row from this CTA's shared memory into another CTA's, bytes signalled on that CTA's mbarrier
// (per-thread st.async stores cost the receiver one mbarrier transaction per lane: 1040 per round, measured slower than the
// cluster barrier they replaced)
__device__ __forceinline__ void bulk_row_to_cta(uint32_t raddr, uint32_t local_addr, uint32_t bytes, uint32_t rbar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(raddr), "r"(local_addr), "r"(bytes), "r"(rbar) : "memory");
}
__device__ __forceinline__ void mbar_init_cta(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_cta(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_cta(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(smem_addr(bar)), "r"(parity), "r"(100000u) : "memory");
}

// a row of the factor spread over one warp: lane holds columns 4 lane .. 4 lane + 3 (and 128 + the same when WIDE);
// nrm = squared norm, carried through the rotations (al' = al - t ga, be' = be + t ga, exact for the exact Jacobi angle)
// and refreshed from the data once per sweep: it only steers the angle, so its rounding drift cannot bias the
// singular values.
template <bool WIDE>
struct JRow {
  float4 lo, hi; float nrm;
  static constexpr int NPOS = WIDE ? SP_NMAX : 128;       // float of the row that carries the squared norm (right behind the data)
  __device__ __forceinline__ void load(const float* p, int lane) {
    lo = *reinterpret_cast<const float4*>(p + lane * 4);
    if (WIDE) hi = *reinterpret_cast<const float4*>(p + 128 + lane * 4);
    nrm = p[NPOS];
  }
  __device__ __forceinline__ void store(float* p, int lane) const {
    *reinterpret_cast<float4*>(p + lane * 4) = lo;
    if (WIDE) *reinterpret_cast<float4*>(p + 128 + lane * 4) = hi;
    if (lane == 0) p[NPOS] = nrm;
  }
  // the same into another CTA's shared memory (shared::cluster address raddr), bytes signalled on its mbarrier rbar
  static constexpr uint32_t ASYNC_BYTES = (WIDE ? 1024u : 512u) + 4u;
  __device__ __forceinline__ void store_async(uint32_t raddr, uint32_t rbar, int lane) const {
    st_async_f4(raddr + (uint32_t)lane * 16u, lo, rbar);
    if (WIDE) st_async_f4(raddr + 512u + (uint32_t)lane * 16u, hi, rbar);
    if (lane == 0) st_async_f1(raddr + (uint32_t)NPOS * 4u, nrm, rbar);
  }
};

__device__ __forceinline__ float rcp_fast(float x) {        // one MUFU.RCP (callers keep x in [1, 2])
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Jacobi rotation of a row pair from its three inner products: sn, tau for Rutishauser's form, cs = cos, tg = tan * ga
// (the amount the squared norms change by), relg = |cos of the angle between the rows| before the rotation (0 when the
// pair is left alone: already orthogonal to rounding, or one row numerically zero).  Half-angle form, see rotation_params.
struct JRot { float sn, tau, cs, tg, relg; };
__device__ __forceinline__ JRot jacobi_angle(float al, float be, float ga, float tol) {
  const float big = fmaxf(al, be), small = fminf(al, be);
  const float rr = rsqrt_fast(al) * rsqrt_fast(be);
  const float d = (be - al) * rr;
  const bool can = small > 1e-30f && small > 1e-24f * big;
  const float g = ga * rr;                                      // cosine of the angle between the rows
  const float relg = fabsf(g);
  const bool rot = can && relg > tol;
  const float rh = rsqrt_fast(fmaf(d, d, 4.f * g * g));
  const float c2 = fabsf(d) * rh;
  const float xx = fmaf(0.5f, c2, 0.5f);
  const float rc = rsqrt_fast(xx);
  const float cc = xx * rc;
  const float s1 = ((d >= 0.f) ? g : -g) * rh * rc;
  JRot r;
  r.sn = rot ? s1 : 0.f;
  r.cs = rot ? cc : 1.f;
  r.tau = rot ? s1 * rcp_fast(1.f + cc) : 0.f;
  r.tg = rot ? s1 * rc * ga : 0.f;
  r.relg = rot ? relg : 0.f;
  return r;
}

template <bool WIDE>
__device__ __forceinline__ void jacobi_apply(JRow<WIDE>& x, JRow<WIDE>& y, const JRot& r) {
  rot4(x.lo, y.lo, r.sn, r.tau);
  if (WIDE) rot4(x.hi, y.hi, r.sn, r.tau);
  x.nrm = fmaxf(x.nrm - r.tg, 0.f);
  y.nrm = fmaxf(y.nrm + r.tg, 0.f);
}

template <bool WIDE>
__device__ __forceinline__ float jdot(const JRow<WIDE>& x, const JRow<WIDE>& y) {
  float2 s = dot4_2(x.lo, y.lo, make_float2(0.f, 0.f));
  if (WIDE) s = dot4_2(x.hi, y.hi, s);
  return s.x + s.y;
}

// Two steps on four rows held in registers: (xa,ya), (xb,yb), then (xa,yb), (xb,ya).  All six inner products of the four
// rows are reduced across the warp at once (six interleaved shuffle chains instead of two dependent rounds of two); the
// inner products the second step needs follow from the first step's rotations by linearity, so the second pair of
// angles does not wait for the rotated rows.  `single`: first step only.  Returns the largest |cos| rotated away.
template <bool WIDE>
__device__ __forceinline__ float jacobi_quad(JRow<WIDE>& xa, JRow<WIDE>& xb, JRow<WIDE>& ya, JRow<WIDE>& yb, float tol, bool single) {
  float p_aa = jdot(xa, ya), p_bb = jdot(xb, yb), p_ab = jdot(xa, yb), p_ba = jdot(xb, ya), p_xx = jdot(xa, xb), p_yy = jdot(ya, yb);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    p_aa += __shfl_xor_sync(0xffffffffu, p_aa, o);
    p_bb += __shfl_xor_sync(0xffffffffu, p_bb, o);
    p_ab += __shfl_xor_sync(0xffffffffu, p_ab, o);
    p_ba += __shfl_xor_sync(0xffffffffu, p_ba, o);
    p_xx += __shfl_xor_sync(0xffffffffu, p_xx, o);
    p_yy += __shfl_xor_sync(0xffffffffu, p_yy, o);
  }
  const JRot r1 = jacobi_angle(xa.nrm, ya.nrm, p_aa, tol);      // xa' = c1 xa - s1 ya, ya' = s1 xa + c1 ya
  const JRot r2 = jacobi_angle(xb.nrm, yb.nrm, p_bb, tol);      // xb' = c2 xb - s2 yb, yb' = s2 xb + c2 yb
  jacobi_apply<WIDE>(xa, ya, r1);
  jacobi_apply<WIDE>(xb, yb, r2);
  float worst = fmaxf(r1.relg, r2.relg);
  if (!single) {
    // <xa', yb'> and <xb', ya'> from the products before the first step
    const float q_ab = r1.cs * r2.sn * p_xx + r1.cs * r2.cs * p_ab - r1.sn * r2.sn * p_ba - r1.sn * r2.cs * p_yy;
    const float q_ba = r2.cs * r1.sn * p_xx + r2.cs * r1.cs * p_ba - r2.sn * r1.sn * p_ab - r2.sn * r1.cs * p_yy;
    const JRot r3 = jacobi_angle(xa.nrm, yb.nrm, q_ab, tol);
    const JRot r4 = jacobi_angle(xb.nrm, ya.nrm, q_ba, tol);
    jacobi_apply<WIDE>(xa, yb, r3);
    jacobi_apply<WIDE>(xb, ya, r4);
    worst = fmaxf(worst, fmaxf(r3.relg, r4.relg));
  }
  return worst;
}

// Four steps on EIGHT rows held in registers (x_0..3 from one block, y_0..3 from the other): step s rotates the pairs
// (x_i, y_(i+s)%4), i.e. all 16 cross pairs in the minimum of four steps.  All 28 inner products of the eight rows are formed
// ONCE per phase and reduced with a transposing butterfly (31 shuffles: after the five levels lane l holds the total of
// product l); from then on the 8 x 8 Gram matrix is carried through the rotations by linearity, distributed over the lanes
// as 2 x 2 blocks: with the y rows relabelled after every step so that the rotated pairs are always (x_i, y_i), lane 4i+j
// holds H_ij = [[x_i.x_j, x_i.y_j], [y_i.x_j, y_i.y_j]] and a step is H_ij <- J_i H_ij J_j^T -- lane-local, only the two
// rotations are fetched from the diagonal lanes 5i, 5j, which computed them (the four angles of a step are computed at
// once, one per lane, instead of one after the other on all lanes).  The relabelling y_i <- y_(i+1) moves three entries of
// every block to a fixed neighbour lane.  A step's dependent chain is angle -> fetch -> block update -> relabel (~250
// cycles) instead of products -> 5-level reduction -> angle per two steps, and a phase is 53 instructions per rotation
// instead of 105: the Jacobi kernel runs one warp per scheduler, so both count.  Squared norms (the diagonal of H) are
// carried as before (al' = al - t ga).  Returns, per lane, the largest |cos| it rotated away (diagonal lanes).
// Every lane holds ONE float4 of each row: at n <= 128 a warp covers the whole rows (PAIR = false); beyond, two warps share
// the eight rows, one column half each (PAIR = true): they add their partial products through shared memory behind a
// two-warp named barrier (fixed order lo + hi: both get the same bits), run the small Gram chain redundantly and rotate
// their own halves -- the long-row FFMA2s of a phase, which bound it, are spread over all four schedulers of the SM.
__device__ __forceinline__ float2 dot4p(const float4& a, const float4& b) {
  const float2 s = __ffma2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w), __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(0.f, 0.f)));
  return s;
}
template <bool PAIR>
__device__ __forceinline__ float jacobi_oct(float4 (&xs)[4], float4 (&ys)[4], float (&xn)[4], float (&yn)[4], float tol, int lane,
                                            float* exch_mine, const float* exch_other, int bar_id, bool first_half) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 d2 = dot4p(xs[i], ys[j]); v[4 * i + j] = d2.x + d2.y; }
  {
    int k = 16;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = i + 1; j < 4; ++j) {
        const float2 dx = dot4p(xs[i], xs[j]), dy = dot4p(ys[i], ys[j]);
        v[k] = dx.x + dx.y; v[k + 6] = dy.x + dy.y; ++k;
      }
  }
  v[28] = v[29] = v[30] = v[31] = 0.f;
  // transposing butterfly: level `off` halves the number of live values; the lane keeps the half selected by its own bit
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < off; ++k) {
      const float send = up ? v[k] : v[k + off];
      const float keep = up ? v[k + off] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  float tot = v[0];                             // total of product `lane` (over this warp's columns)
  if (PAIR) {
    exch_mine[lane] = tot;
    asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
    const float other = exch_other[lane];
    tot = first_half ? tot + other : other + tot;
  }
  // block (i, j) of this lane (lanes 16..31 mirror 0..15) and where its entries come from
  const int bi = (lane >> 2) & 3, bj = lane & 3;
  const int lo_ = min(bi, bj), hi_ = max(bi, bj);
  const int tri = ((lo_ * (7 - lo_)) >> 1) + (hi_ - lo_ - 1) + (bi == bj ? 1 : 0);   // (0,1)(0,2)(0,3)(1,2)(1,3)(2,3) -> 0..5 (diagonal: any valid lane)
  float hxy = __shfl_sync(0xffffffffu, tot, 4 * bi + bj);
  float hyx = __shfl_sync(0xffffffffu, tot, 4 * bj + bi);
  float hxx = __shfl_sync(0xffffffffu, tot, 16 + tri);
  float hyy = __shfl_sync(0xffffffffu, tot, 22 + tri);
  const bool dg = bi == bj;
  {
    const bool b0 = (bi & 1) != 0, b1 = (bi & 2) != 0;
    const float ax01 = b0 ? xn[1] : xn[0], ax23 = b0 ? xn[3] : xn[2];
    const float ay01 = b0 ? yn[1] : yn[0], ay23 = b0 ? yn[3] : yn[2];
    const float ax = b1 ? ax23 : ax01, ay = b1 ? ay23 : ay01;
    hxx = dg ? ax : hxx;
    hyy = dg ? ay : hyy;
  }
  const int src_i = 5 * bi, src_j = 5 * bj;
  const int nb_xy = 4 * bi + ((bj + 1) & 3), nb_yx = 4 * ((bi + 1) & 3) + bj, nb_yy = 4 * ((bi + 1) & 3) + ((bj + 1) & 3);
  float worst = 0.f;
#pragma unroll
  for (int st = 0; st < 4; ++st) {
    // the angle of pair i on lane 5i, from the unnormalised products: two dependent MUFU levels on the chain of the steps
    // (rsqrt(d^2 + 4 g^2), rsqrt((1 + cos 2 theta) / 2)); tan(theta / 2) and |cos| between the rows are off the chain
    // (products scaled by 1 / max norm, which only depends on the carried norms: tiny rows of a rank-deficient bond would
    // underflow in d^2 + 4 g^2 otherwise)
    const float al = hxx, be = hyy, ga = hxy;
    const float big = fmaxf(al, be), small = fminf(al, be);
    const float inv = rcp_fast(big);
    const float relg0 = fabsf(ga) * rsqrt_fast(al) * rsqrt_fast(be);
    const bool rot = small > 1e-30f && small > 1e-24f * big && relg0 > tol;
    const float d = (be - al) * inv, gs = ga * inv, g2 = gs + gs;
    const float rh = rsqrt_fast(fmaf(d, d, g2 * g2));
    const float xx = fmaf(0.5f * fabsf(d), rh, 0.5f);
    const float rc = rsqrt_fast(xx);
    const float cc0 = xx * rc;
    const float s0 = ((d >= 0.f) ? gs : -gs) * rh * rc;
    const float sn = rot ? s0 : 0.f, cs = rot ? cc0 : 1.f;
    const float tau = rot ? s0 * rcp_fast(1.f + cc0) : 0.f;
    const float tg = rot ? s0 * rc * ga : 0.f;
    worst = fmaxf(worst, (dg && rot) ? relg0 : 0.f);
    // the four rotations of the step, to every lane (selects only: a per-lane branch here costs a divergence stack round trip)
    float sk[4], tk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) sk[k] = __shfl_sync(0xffffffffu, sn, 5 * k);
    const float dxx = fmaxf(hxx - tg, 0.f), dyy = fmaxf(hyy + tg, 0.f);
    if (st < 3) {
      // H <- J_i H J_j^T, then relabel y_i <- y_(i+1)
      const float s_i = __shfl_sync(0xffffffffu, sn, src_i), c_i = __shfl_sync(0xffffffffu, cs, src_i);
      const float s_j = __shfl_sync(0xffffffffu, sn, src_j), c_j = __shfl_sync(0xffffffffu, cs, src_j);
      const float txx = fmaf(c_i, hxx, -s_i * hyx), txy = fmaf(c_i, hxy, -s_i * hyy);
      const float tyx = fmaf(s_i, hxx, c_i * hyx), tyy = fmaf(s_i, hxy, c_i * hyy);
      const float nxx = fmaf(c_j, txx, -s_j * txy), nxy = fmaf(s_j, txx, c_j * txy);
      const float nyx = fmaf(c_j, tyx, -s_j * tyy), nyy = fmaf(s_j, tyx, c_j * tyy);
      hxx = dg ? dxx : nxx;
      hxy = __shfl_sync(0xffffffffu, dg ? 0.f : nxy, nb_xy);
      hyx = __shfl_sync(0xffffffffu, dg ? 0.f : nyx, nb_yx);
      hyy = __shfl_sync(0xffffffffu, dg ? dyy : nyy, nb_yy);
    } else {
      hxx = dg ? dxx : hxx; hyy = dg ? dyy : hyy;
    }
    // the rows: pair k is (x_k, y_(k+st)%4)
#pragma unroll
    for (int k = 0; k < 4; ++k) tk[k] = __shfl_sync(0xffffffffu, tau, 5 * k);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      rot4(xs[k], ys[(k + st) & 3], sk[k], tk[k]);
    }
  }
  // carried squared norms back to the rows: after the last step lane 5i holds those of x_i and of y_(i+3)%4
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    xn[k] = __shfl_sync(0xffffffffu, hxx, 5 * k);
    yn[(k + 3) & 3] = __shfl_sync(0xffffffffu, hyy, 5 * k);
  }
  return worst;
}

__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ 2b. Cholesky in a cluster
// The same factorisation G = L L^T (R = L^T) spread over one thread-block cluster, right-looking by column blocks of 16:
// CTA J owns columns 16J .. 16J+15 of L (rows >= 16J), keeps that block of the trailing matrix in shared memory and
//   * for K = 0 .. J-1 waits for column block K of L (rows >= 16J), which CTA K pushes into THIS CTA's shared memory as one
//     bulk copy signalling an mbarrier here, and subtracts its contribution (rank-16 update; all earlier blocks are usually
//     in before block J-1 arrives, so only the last update is on the critical path);
//   * factorises its 16 x 16 diagonal block (one warp, registers + shuffles, the pivot rule of the single-CTA kernel),
//     solves the rows below and pushes rows >= 16J' of the result to every CTA J' > J (no flow control: a CTA has room for
//     all the pieces it will ever receive, at most J (n - 16J) rows <= 147 KB);
//   * writes its 16 rows of R = L^T in FP32 (scaled) and their squared norms; CTA 0 ranks the rows (de Rijk ordering).
// The critical path is 16 x (diagonal block + solve + push + one update) instead of the whole n^3/3 on one SM:
// measured 208 -> see DESIGN.md us at 240 x 240.
constexpr int CC_T = 256, CC_LD = 18;          // row stride of a block in shared memory (doubles): 144 B rows, conflict-free LDS.128

struct CholCArgs {
  const double* G; float* Rf; int R; const int32_t* d_near; float* scale; int32_t* rowmap; int csize;
  const double* gp; int KS;      // KS > 0: G = sum of the KS column-slice partials gp[ks][R][R] of split_gram_kernel (lower tiles), added here in order
};

size_t chol_cluster_smem(int R) {
  size_t best = 0;
  for (int J = 0; 16 * J < R; ++J) best = max(best, (size_t)J * (size_t)(R - 16 * J));
  return ((size_t)SP_NMAX + best) * CC_LD * sizeof(double);
}

__global__ void __launch_bounds__(CC_T) split_chol_cluster_kernel(CholCArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) double ccs[];
  __shared__ __align__(8) uint64_t pbar[16];
  __shared__ __align__(8) uint64_t hbar;                   // the HEAD of the last block to the left (this CTA's diagonal rows) has arrived
  __shared__ __align__(16) double Hd[CW * CC_LD];           // ... and the head itself
  __shared__ double rinv_s[CW];
  __shared__ double red[CC_T / 32];
  __shared__ float nrm_part[CW][SP_NMAX / 32];
  __shared__ float nrm_all[SP_NMAX];                       // CTA 0: squared norms of all rows of R
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int J = (int)cluster.block_rank();
  const int n = 2 * *a.d_near, R = a.R;
  const int NB = (n + CW - 1) / CW;
  const bool active = J < NB;
  const int r0 = CW * J, nrows = active ? n - r0 : 0, w = active ? min(CW, n - r0) : 0;
  double* A = ccs;                                         // A[li * CC_LD + c] = (trailing matrix / L)[r0 + li][r0 + c]
  double* P = ccs + (size_t)SP_NMAX * CC_LD;               // piece K: rows r0 .. n-1 of column block K, nrows x CC_LD
  // largest diagonal entry -> pivot threshold and the power-of-two scale of the FP32 copy (every CTA, same bits)
  // entry (i, j), i >= j or both inside one diagonal block, of the Gram matrix: the slices in a fixed order (all loads in flight)
  auto gram_at = [&](int i, int j) -> double {
    if (a.KS <= 0) return __ldcg(&a.G[(size_t)i * R + j]);
    double v[SP_KS_MAX];
#pragma unroll
    for (int ks = 0; ks < SP_KS_MAX; ++ks) v[ks] = ks < a.KS ? __ldcg(&a.gp[((size_t)ks * R + i) * R + j]) : 0.0;
    double sum = 0.0;
#pragma unroll
    for (int ks = 0; ks < SP_KS_MAX; ++ks) sum += v[ks];
    return sum;
  };
  double dmax = 0.0;
  for (int i = tid; i < n; i += CC_T) dmax = fmax(dmax, gram_at(i, i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  if (lane == 0) red[warp] = dmax;
  if (tid < CW && tid < J) {
    mbar_init_cta(&pbar[tid], 1);
  }
  if (tid == 0) mbar_init_cta(&hbar, 1);
  __syncthreads();
  dmax = 0.0;
  for (int q = 0; q < CC_T / 32; ++q) dmax = fmax(dmax, red[q]);
  const double tolp = (double)n * 2.3e-16 * dmax;
  const double sq_tolp = sqrt(tolp > 0.0 ? tolp : 1e-300), inv_dmax = dmax > 0.0 ? 1.0 / dmax : 0.0, rs_dmax = dmax > 0.0 ? rsqrt(dmax) : 0.0;
  int ex = 0;
  if (dmax > 0.0) frexp(sqrt(dmax), &ex);
  const double inv_scale = ldexp(1.0, -ex);
  if (active && tid < J) mbar_expect_tx_cta(&pbar[tid], (uint32_t)nrows * CC_LD * 8u);
  const int hrows = min(CW, nrows);                        // rows of the head piece (this CTA's diagonal rows)
  if (active && J > 0 && tid == 0) mbar_expect_tx_cta(&hbar, (uint32_t)hrows * CC_LD * 8u);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  // own block of G: rows r0 + li, columns r0 .. r0 + 15 (16 threads read one row's 128 bytes of every slice)
  if (a.KS > 0) {
    for (int e = tid; e < CW * SP_NMAX; e += CC_T) {
      const int c = e % CW, li = e / CW;
      if (li < nrows) A[li * CC_LD + c] = c < w ? gram_at(r0 + li, r0 + c) : 0.0;
    }
  } else {
    // (G stored in full and symmetric: column r0 + c read as row r0 + c, unit stride in li)
    for (int e = tid; e < CW * SP_NMAX; e += CC_T) {
      const int c = e / SP_NMAX, li = e % SP_NMAX;
      if (li < nrows) A[li * CC_LD + c] = c < w ? __ldcg(&a.G[(size_t)(r0 + c) * R + r0 + li]) : 0.0;
    }
  }
  __syncthreads();
  cluster_arrive_release();
  cluster_wait_acquire();
  if (active) {
    // ---- contributions of the column blocks to the left.  The last one (K = J - 1) is on the critical path of the whole
    //      factorisation -- CTA J - 1 has only just finished it -- so it comes in two parts: a HEAD (its rows of this CTA's
    //      diagonal block, 2 KB, sent first) that all 256 threads apply to the diagonal block at once, after which warp 0
    //      factorises the diagonal block while warps 1-7 apply the full piece to the rows below
    auto update_row = [&](const double* Pk, int li) {
      double pr[CW], acc[CW];
#pragma unroll
      for (int t = 0; t < CW; t += 2) {
        const double2 v = *reinterpret_cast<const double2*>(&Pk[li * CC_LD + t]);
        pr[t] = v.x; pr[t + 1] = v.y;
      }
#pragma unroll
      for (int c = 0; c < CW; ++c) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int t = 0; t < CW; t += 2) {
          const double2 v = *reinterpret_cast<const double2*>(&Pk[c * CC_LD + t]);      // L[r0 + c][16 K + t], broadcast
          s0 = fma(pr[t], v.x, s0); s1 = fma(pr[t + 1], v.y, s1);
        }
        acc[c] = s0 + s1;
      }
#pragma unroll
      for (int c = 0; c < CW; ++c) A[li * CC_LD + c] -= acc[c];
    };
    for (int K = 0; K + 1 < J; ++K) {
      mbar_wait_cta(&pbar[K], 0);
      if (tid < nrows) update_row(P + (size_t)K * nrows * CC_LD, tid);
    }
    if (J > 0) {
      mbar_wait_cta(&hbar, 0);
      __syncthreads();                                     // the earlier updates of the diagonal rows are in shared memory
      {
        const int r = tid >> 4, c = tid & 15;              // one entry of the diagonal block per thread
        if (r < hrows && c < hrows) {
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int t = 0; t < CW; t += 2) {
            const double2 x = *reinterpret_cast<const double2*>(&Hd[r * CC_LD + t]);
            const double2 y = *reinterpret_cast<const double2*>(&Hd[c * CC_LD + t]);
            s0 = fma(x.x, y.x, s0); s1 = fma(x.y, y.y, s1);
          }
          A[r * CC_LD + c] -= s0 + s1;
        }
      }
    }
    __syncthreads();
    // ---- diagonal block: one warp, lane r holds row r, 16 dependent steps (see split_chol_kernel)
    if (warp == 0) {
      const int r = lane & 15;
      double d[CW];
#pragma unroll
      for (int c = 0; c < CW; ++c) d[c] = (r < nrows) ? A[r * CC_LD + c] : 0.0;
#pragma unroll
      for (int c = 0; c < CW; ++c) {
        const double piv = __shfl_sync(0xffffffffu, d[c], c);
        const bool drop = !(piv > tolp) || c >= w;
        const double rinv = drop ? 0.0 : rsqrt_seeded(piv * inv_dmax) * rs_dmax;
        const double lkk = drop ? sq_tolp : piv * rinv;
        const double l = r == c ? lkk : d[c] * rinv;
        d[c] = l;
#pragma unroll
        for (int cc = c + 1; cc < CW; ++cc) {
          const double t = __shfl_sync(0xffffffffu, l, cc);
          d[cc] = fma(-l, t, d[cc]);
        }
        if (lane == 0) rinv_s[c] = rinv;
      }
      if (lane < CW && r < nrows) {
#pragma unroll
        for (int c = 0; c < CW; ++c) A[r * CC_LD + c] = c <= r ? d[c] : 0.0;
      }
    } else if (J > 0) {
      // warps 1-7 meanwhile: the full last piece applied to the rows below the diagonal block (at most 224 of them);
      // its first 16 rows are the head again (already applied to the diagonal block)
      mbar_wait_cta(&pbar[J - 1], 0);
      const int li = tid - 32 + CW;
      if (li < nrows) update_row(P + (size_t)(J - 1) * nrows * CC_LD, li);
    }
    __syncthreads();
    // ---- rows below: forward substitution with the factor of the diagonal block
    if (tid >= CW && tid < nrows) {
      double p[CW];
#pragma unroll
      for (int c = 0; c < CW; c += 2) {
        const double2 v = *reinterpret_cast<const double2*>(&A[tid * CC_LD + c]);
        p[c] = v.x; p[c + 1] = v.y;
      }
#pragma unroll
      for (int c = 0; c < CW; ++c) {
        const double l = p[c] * rinv_s[c];
        p[c] = l;
#pragma unroll
        for (int cc = c + 1; cc < CW; ++cc) p[cc] = fma(-l, A[cc * CC_LD + c], p[cc]);
      }
#pragma unroll
      for (int c = 0; c < CW; c += 2) *reinterpret_cast<double2*>(&A[tid * CC_LD + c]) = make_double2(p[c], p[c + 1]);
    }
    // ---- push rows >= 16 J' to every CTA J' > J (nearest first: it is the one waiting)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
      if (lane == 0 && J + 1 < NB) {
        // head: the next CTA's diagonal rows (rows 16 .. of this block), first in the queue
        const int hr = min(CW, n - CW * (J + 1));
        bulk_row_to_cta(map_to_cta(smem_addr(Hd), J + 1), smem_addr(A + (size_t)CW * CC_LD), (uint32_t)hr * CC_LD * 8u,
                        map_to_cta(smem_addr(&hbar), J + 1));
      }
      __syncwarp();
      const int Jd = J + 1 + lane;
      if (Jd < NB) {
        const int nr_d = n - CW * Jd;
        const uint32_t dst = map_to_cta(smem_addr(P + (size_t)J * nr_d * CC_LD), Jd);
        const uint32_t bar = map_to_cta(smem_addr(&pbar[J]), Jd);
        bulk_row_to_cta(dst, smem_addr(A + (size_t)(CW * (Jd - J)) * CC_LD), (uint32_t)nr_d * CC_LD * 8u, bar);
      }
    }
    // ---- rows r0 .. r0+w-1 of R = L^T in FP32 (scaled), zeros left of the block and beyond n; their squared norms
#pragma unroll
    for (int u = 0; u < CW * SP_NMAX / CC_T; ++u) {
      const int e = tid + u * CC_T, c = e / SP_NMAX, j = e % SP_NMAX;
      float f = 0.f;
      if (c < w && j < R) {
        if (j >= r0 && j < n) f = (float)(A[(j - r0) * CC_LD + c] * inv_scale);
        a.Rf[(size_t)(r0 + c) * R + j] = f;
      }
      float sq = f * f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      if (lane == 0) nrm_part[c][j / 32] = sq;
    }
    __syncthreads();
    if (tid < w) {
      float sq = 0.f;
#pragma unroll
      for (int q = 0; q < SP_NMAX / 32; ++q) sq += nrm_part[tid][q];
      *cluster.map_shared_rank(&nrm_all[r0 + tid], 0) = sq;
    }
  }
  // nobody leaves while a copy out of its shared memory may be in flight; the norms are with CTA 0
  cluster_arrive_release();
  cluster_wait_acquire();
  if (J == 0) {
    if (tid < n) {
      const float sp = nrm_all[tid];
      int rank = 0;
      for (int o = 0; o < n; ++o) {
        const float so = nrm_all[o];
        rank += (so > sp) || (so == sp && o < tid);
      }
      a.rowmap[rank] = tid;
    }
    if (tid == 0) { a.scale[0] = (float)ldexp(1.0, ex); a.scale[1] = (float)dmax; }
  }
}

// Inside a CTA the 32 rows of its two blocks are rotated by 8 warps with 2 x 2 register blocking: a warp holds two rows
// x_a, x_b in registers over several phases, fetches two rows y_a, y_b per phase from shared memory and rotates
// (x_a,y_a), (x_b,y_b), then (x_a,y_b), (x_b,y_a): four rotations per two rows moved through shared memory (a row per
// rotation made shared-memory bandwidth the limit: 64 KB per step).  Cross-block pairs: 8 phases.  Pairs inside a block
// (once per sweep): the same scheme applied to halves, quarters, eighths of a block (4 + 2 + 1 phases) and a last phase
// for the pairs inside the groups of two: 31 steps in all, the minimum for 32 rows.
// BR = rows per block: 16 (8 warps per CTA, clusters of up to 8) or 8 (4 warps per CTA -- one per scheduler, so a phase runs
// at a single warp's latency instead of two warps sharing the issue slots -- and clusters of up to 16 CTAs, a
// non-portable size the host enables when the device can schedule it).
template <bool WIDE, int BR>
__global__ void __launch_bounds__(16 * BR) split_jacobi_kernel(JacArgs a) {
  constexpr int JB = BR, JWARPS = BR / 2, JT = 16 * BR, JPH_CROSS = BR / 2, JPH_ALL = BR;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float jsm[];                 // [2 buffers][2 slots][JB][JLD]
  __shared__ float wmax[2][JCL_MAX];
  __shared__ float wwarp[JWARPS];
  __shared__ __align__(8) uint64_t rx_bar[2];                  // rx_bar[b]: the two blocks of buffer b have arrived
  __shared__ uchar4 tab_rows[JPH_ALL][JWARPS];
  __shared__ unsigned char tab_flags[JPH_ALL][JWARPS];         // 1: load x rows, 2: store x rows, 4: first half only
  __shared__ float oct_x[2][2][32];                            // partial inner products of the two column halves of an oct
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int crank = (int)cluster.block_rank();
  const int n = 2 * *a.d_near, R = a.R;
  const int nb = (n + JB - 1) / JB;
  int nbe = nb + (nb & 1);
  if (nbe < 2) nbe = 2;
  if (nbe > 2 * a.csize) nbe = 2 * a.csize;                    // (the host sizes the cluster for R, so this never binds)
  const int h = nbe / 2, rounds = nbe - 1;
  const bool active = crank < h;
  auto buf = [&](int b, int slot, int r) -> float* { return jsm + ((size_t)((b * 2 + slot) * JB + r)) * JLD; };
  if (tid < JPH_ALL * JWARPS) {
    // phase table: cross-block pairs first (x rows from block A, y rows from block B), then the pairs inside the blocks:
    // the same 2 x 2 scheme on groups of BR, BR/2, .. 4 rows (x from the first half of a group, y from the second), and
    // a last phase for the pairs inside the groups of two
    const int ph = tid / JWARPS, w = tid % JWARPS;
    int xa = 0, xb = 0, ya = 0, yb = 0, fl = 0;
    if (ph < JPH_CROSS) {
      const int t = ph, y = JB + 2 * ((w + t) % JWARPS);
      xa = 2 * w; xb = xa + 1; ya = y; yb = y + 1; fl = (t == 0 ? 1 : 0) | (t == JPH_CROSS - 1 ? 2 : 0);
    } else {
      int rest = ph - JPH_CROSS, S = JB;
      bool done = false;
      while (S >= 4 && !done) {
        const int np = S / 4;                    // phases of this level = warps per group
        if (rest < np) {
          const int t = rest, g = S * (w / np), q = w % np, y = g + S / 2 + 2 * ((q + t) % np);
          xa = g + 2 * q; xb = xa + 1; ya = y; yb = y + 1; fl = (t == 0 ? 1 : 0) | (t == np - 1 ? 2 : 0);
          done = true;
        }
        rest -= np; S /= 2;
      }
      if (!done) { const int g = 4 * w; xa = g; ya = g + 1; xb = g + 2; yb = g + 3; fl = 7; }
    }
    tab_rows[ph][w] = make_uchar4((unsigned char)xa, (unsigned char)xb, (unsigned char)ya, (unsigned char)yb);
    tab_flags[ph][w] = (unsigned char)fl;
  }
  // ---- load: position p holds block p; CTA j has positions j (slot 0) and nbe-1-j (slot 1)
  for (int e = tid; e < 2 * JB * (JLD / 4); e += JT) {
    const int c4 = e % (JLD / 4), r = (e / (JLD / 4)) % JB, slot = e / (JB * (JLD / 4));
    const int blk = slot == 0 ? crank : nbe - 1 - crank;
    const int rank = blk * JB + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active && rank < n && c4 * 4 < n) v = __ldg(reinterpret_cast<const float4*>(a.Rf + (size_t)a.rowmap[rank] * R) + c4);   // columns >= n of Rf are zero
    *reinterpret_cast<float4*>(buf(0, slot, r) + c4 * 4) = v;
    *reinterpret_cast<float4*>(buf(1, slot, r) + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (tid < 2 * JCL_MAX) (&wmax[0][0])[tid] = 0.f;
  if (tid < 2) mbar_init_cta(&rx_bar[tid], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  cluster_arrive_release();
  cluster_wait_acquire();
  // where this CTA's two blocks go after a round: position 0 stays, the others move one place round the ring 1 .. nbe-1
  int dcta0, dslot0, dcta1, dslot1;
  if (crank == 0) { dcta0 = 0; dslot0 = 0; } else if (crank + 1 < h) { dcta0 = crank + 1; dslot0 = 0; } else { dcta0 = h - 1; dslot0 = 1; }
  if (crank == 0) { dcta1 = h > 1 ? 1 : 0; dslot1 = h > 1 ? 0 : 1; } else { dcta1 = crank - 1; dslot1 = 1; }
  int cur = 0, sweeps_used = 0;
  uint32_t rx_phase = 0;                                       // parity of rx_bar[b] in bit b
  constexpr uint32_t ROW_BYTES = (JRow<WIDE>::NPOS + 4) * 4u, RX_BYTES = 2u * JB * ROW_BYTES;      // data + the float4 holding the carried norm
  constexpr uint32_t ROW_STRIDE = JLD * 4u;
  const bool timer = a.dbg != nullptr && crank == 0 && warp == 0;     // warp-uniform: a per-thread branch would make the shuffles divergence-checked
  long long tacc[2] = {0, 0}, tprev = timer ? clock64() : 0;
  auto tick = [&](int k) { if (timer) { const long long now = clock64(); tacc[k] += now - tprev; tprev = now; } };
  for (int sweep = 0; sweep < a.max_sweeps; ++sweep) {
    float worst = 0.f;
    for (int r = 0; r < rounds; ++r) {
      if (active) {
        float* base = jsm + (size_t)cur * 2 * JB * JLD;
        // the last phase of a round stores every row: it stores them where they are needed next (another CTA's shared
        // memory, st.async + that CTA's mbarrier); this CTA expects its own two blocks for the other buffer
        const bool xchg = rounds > 1;
        uint32_t rdst0 = 0, rdst1 = 0, rbar0 = 0, rbar1 = 0;
        if (xchg) {
          rdst0 = map_to_cta(smem_addr(buf(cur ^ 1, dslot0, 0)), dcta0); rbar0 = map_to_cta(smem_addr(&rx_bar[cur ^ 1]), dcta0);
          rdst1 = map_to_cta(smem_addr(buf(cur ^ 1, dslot1, 0)), dcta1); rbar1 = map_to_cta(smem_addr(&rx_bar[cur ^ 1]), dcta1);
          if (tid == 0) mbar_expect_tx_cta(&rx_bar[cur ^ 1], RX_BYTES);
        }
        // a row goes back into this CTA's buffer and from there, as one bulk copy, to its next owner
        auto send = [&](const JRow<WIDE>& row_v, int row) {
          float* loc = base + (size_t)row * JLD;
          row_v.store(loc, lane);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if (row < JB) bulk_row_to_cta(rdst0 + (uint32_t)row * ROW_STRIDE, smem_addr(loc), ROW_BYTES, rbar0);
            else bulk_row_to_cta(rdst1 + (uint32_t)(row - JB) * ROW_STRIDE, smem_addr(loc), ROW_BYTES, rbar1);
          }
        };
        if (r == 0) {
          // squared norms from the data (four rows per warp)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float* p = base + (size_t)(4 * warp + k) * JLD;
            const float4 x0 = *reinterpret_cast<const float4*>(p + lane * 4);
            float al = dot4(x0, x0, 0.f);
            if (WIDE) { const float4 x1 = *reinterpret_cast<const float4*>(p + 128 + lane * 4); al += dot4(x1, x1, 0.f); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) al += __shfl_xor_sync(0xffffffffu, al, o);
            if (lane == 0) p[JRow<WIDE>::NPOS] = al;
          }
          __syncthreads();
        }
        const int nph = r == 0 ? JPH_ALL : JPH_CROSS;
        int ph0 = 0;
        if constexpr (BR == 8) {
          // cross-block pairs: two phases of jacobi_oct by warps 0 and 1 (x rows 4w..4w+3 of block A stay in registers, the y
          // rows of block B alternate between the two warps); the pairs inside the blocks (round 0) follow on the 2 x 2 path
          ph0 = JPH_CROSS;
          // WIDE: warps w and w + 2 share an oct, one column half each.  Otherwise warps 2 and 3 shadow warps 0 and 1 (same
          // loads, same arithmetic, nothing stored): a per-thread branch around jacobi_oct would make every one of its
          // shuffles divergence-checked (WARPSYNC + ENDCOLLECTIVE per SHFL)
          const bool last_is_oct = r != 0;
          const int ow = warp & 1, half = WIDE ? (warp >> 1) : 0;
          const bool owner = WIDE || warp < 2;
          const int coff = 128 * half + lane * 4;
          constexpr int NPOS = JRow<WIDE>::NPOS;
          float4 xs[4], ys[4];
          float xn[4], yn[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float* pr = base + (size_t)(4 * ow + k) * JLD;
            xs[k] = *reinterpret_cast<const float4*>(pr + coff); xn[k] = pr[NPOS];
          }
#pragma unroll 1
          for (int t = 0; t < 2; ++t) {
            const int y0 = JB + 4 * ((ow + t) & 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float* pr = base + (size_t)(y0 + k) * JLD;
              ys[k] = *reinterpret_cast<const float4*>(pr + coff); yn[k] = pr[NPOS];
            }
            const float w8 = jacobi_oct<WIDE>(xs, ys, xn, yn, a.tol, lane, &oct_x[ow][half][0], &oct_x[ow][half ^ 1][0], 1 + ow, half == 0);
            worst = owner ? fmaxf(worst, w8) : worst;
            const bool hand_over = t == 1 && last_is_oct && xchg;
            if (owner) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float* pr = base + (size_t)(y0 + k) * JLD;
                *reinterpret_cast<float4*>(pr + coff) = ys[k];
                if (half == 0 && lane == 0) pr[NPOS] = yn[k];
              }
              if (t == 1) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  float* pr = base + (size_t)(4 * ow + k) * JLD;
                  *reinterpret_cast<float4*>(pr + coff) = xs[k];
                  if (half == 0 && lane == 0) pr[NPOS] = xn[k];
                }
              }
            }
            if (hand_over) {
              // every row of the round goes to its next owner as one bulk copy, once both column halves are in shared memory
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              if (WIDE) asm volatile("bar.sync %0, 64;" ::"r"(1 + ow) : "memory");
              else __syncwarp();
              if (owner && half == 0 && lane < 8) {
                // lanes 0..3: the y rows, lanes 4..7: the x rows (eight copies issued at once instead of one after the other)
                const int k = lane & 3;
                const bool isx = lane >= 4;
                const int row = isx ? 4 * ow + k : y0 + k;
                const uint32_t dst = isx ? rdst0 + (uint32_t)row * ROW_STRIDE : rdst1 + (uint32_t)(row - JB) * ROW_STRIDE;
                bulk_row_to_cta(dst, smem_addr(base + (size_t)row * JLD), ROW_BYTES, isx ? rbar0 : rbar1);
              }
            } else {
              __syncthreads();
            }
          }
        }
        JRow<WIDE> xa, xb, ya, yb;
        for (int ph = ph0; ph < nph; ++ph) {
          const uchar4 rw = tab_rows[ph][warp];
          const int fl = tab_flags[ph][warp];
          if (fl & 1) { xa.load(base + (size_t)rw.x * JLD, lane); xb.load(base + (size_t)rw.y * JLD, lane); }
          ya.load(base + (size_t)rw.z * JLD, lane); yb.load(base + (size_t)rw.w * JLD, lane);
          worst = fmaxf(worst, jacobi_quad<WIDE>(xa, xb, ya, yb, a.tol, (fl & 4) != 0));
          if (ph == nph - 1 && xchg) {
            send(ya, rw.z); send(yb, rw.w); send(xa, rw.x); send(xb, rw.y);
          } else if (ph == nph - 1) {
            ya.store(base + (size_t)rw.z * JLD, lane); yb.store(base + (size_t)rw.w * JLD, lane);
            xa.store(base + (size_t)rw.x * JLD, lane); xb.store(base + (size_t)rw.y * JLD, lane);
            __syncthreads();
          } else {
            ya.store(base + (size_t)rw.z * JLD, lane); yb.store(base + (size_t)rw.w * JLD, lane);
            if (fl & 2) { xa.store(base + (size_t)rw.x * JLD, lane); xb.store(base + (size_t)rw.y * JLD, lane); }
            __syncthreads();
          }
        }
      }
      if (r == rounds - 1) {
        // sweep maximum of |cos| over the rotated pairs -> every CTA
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) worst = fmaxf(worst, __shfl_xor_sync(0xffffffffu, worst, o));
        if (lane == 0) wwarp[warp] = worst;
        __syncthreads();
        if (tid < a.csize) {
          float m = 0.f;
          for (int w = 0; w < JWARPS; ++w) m = fmaxf(m, wwarp[w]);
          float* dstw = cluster.map_shared_rank(&wmax[sweep & 1][crank], tid);
          *dstw = active ? m : 0.f;
        }
      }
      tick(0);
      if (active && rounds > 1) {
        // my two blocks of the next round (from my ring neighbours); a sender cannot overwrite a buffer its receiver still
        // reads: it needs that receiver's previous output first (every edge of the ring is used in both directions)
        mbar_wait_cta(&rx_bar[cur ^ 1], (rx_phase >> (cur ^ 1)) & 1u);
        rx_phase ^= 1u << (cur ^ 1);
        cur ^= 1;
      }
      if (r == rounds - 1) {
        // once per sweep: the convergence flags of all CTAs (and, after the last sweep, nobody leaves while a store is in flight)
        cluster_arrive_release();
        cluster_wait_acquire();
      }
      tick(1);
    }
    sweeps_used = sweep + 1;
    float wall = 0.f;
    for (int k = 0; k < a.csize; ++k) wall = fmaxf(wall, wmax[sweep & 1][k]);
    if (wall < a.stop) break;             // quadratic convergence: what this sweep left behind is O(stop^2)
  }
  // ---- normalised rows (= eigenvectors of G) and their norms, by slot index
  if (active) {
    for (int rr = warp; rr < 2 * JB; rr += JWARPS) {
      const float* px = buf(cur, rr / JB, rr % JB) + lane * 4;
      float4 x0 = *reinterpret_cast<const float4*>(px), x1 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (WIDE) x1 = *reinterpret_cast<const float4*>(px + 128);
      float al = dot4(x0, x0, 0.f);
      if (WIDE) al = dot4(x1, x1, al);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) al += __shfl_xor_sync(0xffffffffu, al, o);
      const float nr = sqrtf(al);
      const float inv = nr > 1e-30f ? 1.f / nr : 0.f;           // utils.check_nan: a direction that does not exist is zero
      const int slot_row = crank * 2 * JB + rr;
      float* out = a.X + (size_t)slot_row * R;
      x0.x *= inv; x0.y *= inv; x0.z *= inv; x0.w *= inv;
      x1.x *= inv; x1.y *= inv; x1.z *= inv; x1.w *= inv;
      if (lane * 4 < R) *reinterpret_cast<float4*>(out + lane * 4) = x0;
      if (128 + lane * 4 < R) *reinterpret_cast<float4*>(out + 128 + lane * 4) = x1;
      if (lane == 0) a.sig[slot_row] = nr;
    }
  }
  if (timer && lane == 0) for (int k = 0; k < 2; ++k) a.dbg[k] = (float)tacc[k];
  if (crank == 0 && tid == 0) {
    a.iscal[TRMPS_I_SVD_SWEEPS] = sweeps_used;
    a.iscal[TRMPS_I_N_SWEEPS] += sweeps_used;
  }
}

// ------------------------------------------------------------------------------------------------ 3b. Jacobi, Gram form
// Same tournament, different work split inside a CTA (the row-pair-per-warp kernel above is bound by instruction issue:
// every rotation costs ~140 warp instructions, a third of them the angle arithmetic repeated on all 32 lanes).  Here
//   * the 32 x 32 Gram matrix G of the CTA's rows is formed once per round (shared-memory tiles, fixed summation order);
//   * a CHAIN group of 8 warps runs the round's steps on G alone: 16 lanes compute the 16 angles of a step, then every
//     thread applies the step's two rotations to one 2 x 2 block of G (G <- J^T G J).  These are exactly the rotations
//     Hestenes' method would apply to the rows: an angle only needs three inner products;
//   * an APPLY group of 4 warps holds the rows in registers (one or two columns per thread, all 32 rows) and replays the
//     recorded rotations as soon as a step's angles are published (one mbarrier per step): the 4 FMAs per element are
//     the only work done on the long rows, with compile-time register indices and no shared-memory traffic.
// Selected with trmps_opt.split_impl = 3.  Measured at 240 x 240 (tools/split_time.py): the same sweeps, 289k cycles per
// sweep against 303k of the kernel above -- the chain takes 650 cycles per step (two named barriers, and it shares the
// issue slots with the apply group), the Gram matrix 3.9k and the wait for the apply group + hand-over 4.3k cycles per
// round -- so it is not the default; it stays as the parity-tested alternative and as the starting point for a version
// that carries the diagonal Gram blocks with the rows.
constexpr int G2_APPLY = 128, G2_CHAIN = 256, G2_T = G2_APPLY + G2_CHAIN;
constexpr int G2_GLD = 33, G2_STEPS_ALL = 2 * JB - 1, G2_STEPS_X = JB, G2_TILES = 36, G2_KS = 8;

__device__ __forceinline__ void g2_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void g2_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void g2_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity), "r"(100000u) : "memory");
}

// replay of one step on the register-resident columns: pairs known at compile time (ALL: round-robin over the 32 rows,
// else cross pairs i <-> JB + (i + S) % JB)
template <int NC, bool ALL, int S>
__device__ __forceinline__ void g2_apply_step(float (&x)[2 * JB][NC], const float2* rec_s) {
#pragma unroll
  for (int k = 0; k < JB; ++k) {
    constexpr int dummy = 0; (void)dummy;
    const int p = ALL ? rr_first(2 * JB, S, k) : k;
    const int q = ALL ? rr_second(2 * JB, S, k) : JB + (k + S) % JB;
    const float2 r = rec_s[k];                  // (sn, tau)
    if (r.x != 0.f) {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const float u = x[p][c], w = x[q][c];
        x[p][c] = fmaf(-r.x, fmaf(r.y, u, w), u);
        x[q][c] = fmaf(r.x, fmaf(-r.y, w, u), w);
      }
    }
  }
}
template <int NC, bool ALL, int S, int NS>
struct G2ApplyLoop {
  static __device__ __forceinline__ void run(float (&x)[2 * JB][NC], const float2 (*rec)[JB], uint64_t* bars, uint32_t& phase) {
    g2_mbar_wait(&bars[S], (phase >> S) & 1u);
    phase ^= 1u << S;
    g2_apply_step<NC, ALL, S>(x, rec[S]);
    G2ApplyLoop<NC, ALL, S + 1, NS>::run(x, rec, bars, phase);
  }
};
template <int NC, bool ALL, int NS>
struct G2ApplyLoop<NC, ALL, NS, NS> {
  static __device__ __forceinline__ void run(float (&)[2 * JB][NC], const float2 (*)[JB], uint64_t*, uint32_t&) {}
};

template <int NC>
__global__ void __launch_bounds__(G2_T) split_jacobi_gram_kernel(JacArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float jsm[];                 // [2 buffers][2 slots][JB][JLD] rows, then the Gram partials
  float* gpart = jsm + 2 * 2 * JB * JLD;                       // [G2_KS][G2_TILES * 16]
  __shared__ float G[2][2 * JB][G2_GLD];
  __shared__ float2 rec[G2_STEPS_ALL][JB];                     // (sn, tau) of every rotation of the round
  __shared__ float2 coef[JB];                                  // (cos, sin) of the current step's rotations
  __shared__ unsigned char tp[G2_STEPS_ALL + G2_STEPS_X][JB], tq[G2_STEPS_ALL + G2_STEPS_X][JB];
  __shared__ unsigned char tile_i[G2_TILES], tile_j[G2_TILES];
  __shared__ __align__(8) uint64_t step_bar[G2_STEPS_ALL];
  __shared__ float wmax[2][JCL_MAX];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_apply = tid < G2_APPLY;
  const int ct = tid - G2_APPLY;                               // chain thread index
  const int crank = (int)cluster.block_rank();
  const int n = 2 * *a.d_near, R = a.R;
  const int nb = (n + JB - 1) / JB;
  int nbe = nb + (nb & 1);
  if (nbe < 2) nbe = 2;
  if (nbe > 2 * a.csize) nbe = 2 * a.csize;
  const int h = nbe / 2, rounds = nbe - 1;
  const bool active = crank < h;
  auto buf = [&](int b, int slot, int r) -> float* { return jsm + ((size_t)((b * 2 + slot) * JB + r)) * JLD; };
  for (int e = tid; e < (G2_STEPS_ALL + G2_STEPS_X) * JB; e += G2_T) {
    const int st = e / JB, j = e % JB;
    if (st < G2_STEPS_ALL) { tp[st][j] = (unsigned char)rr_first(2 * JB, st, j); tq[st][j] = (unsigned char)rr_second(2 * JB, st, j); }
    else { tp[st][j] = (unsigned char)j; tq[st][j] = (unsigned char)(JB + (j + st - G2_STEPS_ALL) % JB); }
  }
  if (tid < G2_TILES) {
    int ti = 0, rem = tid;
    while (rem >= 8 - ti) { rem -= 8 - ti; ++ti; }             // tiles (ti, tj), ti <= tj < 8
    tile_i[tid] = (unsigned char)ti; tile_j[tid] = (unsigned char)(ti + rem);
  }
  if (tid < G2_STEPS_ALL) g2_mbar_init(&step_bar[tid], 1);
  if (tid < 2 * JCL_MAX) (&wmax[0][0])[tid] = 0.f;
  // ---- load: position p holds block p; CTA j has positions j (slot 0) and nbe-1-j (slot 1)
  for (int e = tid; e < 2 * JB * (JLD / 4); e += G2_T) {
    const int c4 = e % (JLD / 4), r = (e / (JLD / 4)) % JB, slot = e / (JB * (JLD / 4));
    const int blk = slot == 0 ? crank : nbe - 1 - crank;
    const int rank = blk * JB + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active && rank < n && c4 * 4 < n) v = __ldg(reinterpret_cast<const float4*>(a.Rf + (size_t)a.rowmap[rank] * R) + c4);   // columns >= n of Rf are zero
    *reinterpret_cast<float4*>(buf(0, slot, r) + c4 * 4) = v;
    *reinterpret_cast<float4*>(buf(1, slot, r) + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  cluster_arrive_release();
  cluster_wait_acquire();
  int dcta0, dslot0, dcta1, dslot1;
  if (crank == 0) { dcta0 = 0; dslot0 = 0; } else if (crank + 1 < h) { dcta0 = crank + 1; dslot0 = 0; } else { dcta0 = h - 1; dslot0 = 1; }
  if (crank == 0) { dcta1 = h > 1 ? 1 : 0; dslot1 = h > 1 ? 0 : 1; } else { dcta1 = crank - 1; dslot1 = 1; }
  int cur = 0, sweeps_used = 0;
  uint32_t phase = 0;                                          // parity of every step barrier (one bit per step)
  const bool timer = a.dbg != nullptr && crank == 0 && warp == G2_APPLY / 32;      // first chain warp of CTA 0
  long long tacc[4] = {0, 0, 0, 0}, tprev = timer ? clock64() : 0;
  auto tick = [&](int k) { if (timer) { const long long now = clock64(); tacc[k] += now - tprev; tprev = now; } };
  for (int sweep = 0; sweep < a.max_sweeps; ++sweep) {
    float worst = 0.f;
    for (int r = 0; r < rounds; ++r) {
      const bool all = r == 0;                                 // first round of a sweep: every pair of the 32 rows
      const int nsteps = all ? G2_STEPS_ALL : G2_STEPS_X;
      if (active) {
        float* base = jsm + (size_t)cur * 2 * JB * JLD;
        // ---- Gram matrix: 4 x 4 register tiles (rows ti + 8 r, tj + 8 c: conflict-free float4 reads) over 8 column slices
        if (tid < G2_TILES * G2_KS) {
          const int tl = tid % G2_TILES, ks = tid / G2_TILES;
          const int ti = tile_i[tl], tj = tile_j[tl];
          float acc[4][4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
          const float* pa = base + (size_t)ti * JLD + ks * 32;
          const float* pb = base + (size_t)tj * JLD + ks * 32;
#pragma unroll 2
          for (int k4 = 0; k4 < 8; ++k4) {
            float4 av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              av[i] = *reinterpret_cast<const float4*>(pa + (size_t)(8 * i) * JLD + 4 * k4);
              bv[i] = *reinterpret_cast<const float4*>(pb + (size_t)(8 * i) * JLD + 4 * k4);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[i][j] = dot4(av[i], bv[j], acc[i][j]);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) gpart[(size_t)ks * (G2_TILES * 16) + tl * 16 + i * 4 + j] = acc[i][j];
        }
        __syncthreads();
        for (int e = tid; e < G2_TILES * 16; e += G2_T) {
          float sacc = 0.f;
#pragma unroll
          for (int ks = 0; ks < G2_KS; ++ks) sacc += gpart[(size_t)ks * (G2_TILES * 16) + e];
          const int tl = e >> 4, i = (e >> 2) & 3, j = e & 3;
          const int ra = tile_i[tl] + 8 * i, rb = tile_j[tl] + 8 * j;
          G[0][ra][rb] = sacc;
          G[0][rb][ra] = sacc;
        }
        __syncthreads();
        tick(0);
        if (is_apply) {
          // ---- the rows, column-wise in registers
          float x[2 * JB][NC];
#pragma unroll
          for (int rr = 0; rr < 2 * JB; ++rr)
#pragma unroll
            for (int c = 0; c < NC; ++c) x[rr][c] = base[(size_t)rr * JLD + tid + G2_APPLY * c];
          if (all) G2ApplyLoop<NC, true, 0, G2_STEPS_ALL>::run(x, rec, step_bar, phase);
          else G2ApplyLoop<NC, false, 0, G2_STEPS_X>::run(x, rec, step_bar, phase);
          // every row goes where it is needed next (another CTA's shared memory)
          float *dst0, *dst1;
          if (rounds > 1) { dst0 = cluster.map_shared_rank(buf(cur ^ 1, dslot0, 0), dcta0); dst1 = cluster.map_shared_rank(buf(cur ^ 1, dslot1, 0), dcta1); }
          else { dst0 = base; dst1 = base + (size_t)JB * JLD; }
#pragma unroll
          for (int rr = 0; rr < 2 * JB; ++rr)
#pragma unroll
            for (int c = 0; c < NC; ++c) ((rr < JB ? dst0 : dst1) + (size_t)(rr % JB) * JLD)[tid + G2_APPLY * c] = x[rr][c];
        } else {
          // ---- the chain of the round's steps on the Gram matrix
          const int iA = ct >> 4, iB = ct & 15;
          int gc = 0;
          for (int s = 0; s < nsteps; ++s) {
            const int st = all ? s : G2_STEPS_ALL + s;
            if (ct < JB) {
              const int p = tp[st][ct], q = tq[st][ct];
              const JRot rt = jacobi_angle(G[gc][p][p], G[gc][q][q], G[gc][p][q], a.tol);
              coef[ct] = make_float2(rt.cs, rt.sn);
              rec[s][ct] = make_float2(rt.sn, rt.tau);
              worst = fmaxf(worst, rt.relg);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(G2_CHAIN) : "memory");
            if (ct == 0) g2_mbar_arrive(&step_bar[s]);          // the apply group may replay step s
            {
              const int pA = tp[st][iA], qA = tq[st][iA], pB = tp[st][iB], qB = tq[st][iB];
              const float2 ca = coef[iA], cb = coef[iB];
              const float g00 = G[gc][pA][pB], g01 = G[gc][pA][qB], g10 = G[gc][qA][pB], g11 = G[gc][qA][qB];
              // rows: x_p' = c x_p - s x_q, x_q' = s x_p + c x_q; then the same on the columns
              const float r00 = ca.x * g00 - ca.y * g10, r01 = ca.x * g01 - ca.y * g11;
              const float r10 = ca.y * g00 + ca.x * g10, r11 = ca.y * g01 + ca.x * g11;
              G[gc ^ 1][pA][pB] = cb.x * r00 - cb.y * r01;
              G[gc ^ 1][pA][qB] = cb.y * r00 + cb.x * r01;
              G[gc ^ 1][qA][pB] = cb.x * r10 - cb.y * r11;
              G[gc ^ 1][qA][qB] = cb.y * r10 + cb.x * r11;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(G2_CHAIN) : "memory");
            gc ^= 1;
          }
          tick(1);
          // the chain threads keep the barrier parities in step with the apply group
          phase ^= all ? ((1u << G2_STEPS_ALL) - 1u) : ((1u << G2_STEPS_X) - 1u);
        }
      }
      if (r == rounds - 1) {
        // sweep maximum of |cos| over the rotated pairs (held by the 16 angle lanes) -> every CTA
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) worst = fmaxf(worst, __shfl_xor_sync(0xffffffffu, worst, o));
        if (tid == G2_APPLY) {
          for (int k = 0; k < a.csize; ++k) *cluster.map_shared_rank(&wmax[sweep & 1][crank], k) = active ? worst : 0.f;
        }
      }
      cluster_arrive_release();
      cluster_wait_acquire();
      tick(2);
      if (rounds > 1) cur ^= 1;
    }
    sweeps_used = sweep + 1;
    float wall = 0.f;
    for (int k = 0; k < a.csize; ++k) wall = fmaxf(wall, wmax[sweep & 1][k]);
    if (wall < a.stop) break;
  }
  if (timer && lane == 0) for (int k = 0; k < 3; ++k) a.dbg[k] = (float)tacc[k];
  // ---- normalised rows (= eigenvectors of G) and their norms, by slot index
  if (active) {
    for (int rr = warp; rr < 2 * JB; rr += G2_T / 32) {
      const float* px = buf(cur, rr / JB, rr % JB) + lane * 4;
      float4 x0 = *reinterpret_cast<const float4*>(px), x1 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (NC == 2) x1 = *reinterpret_cast<const float4*>(px + 128);
      float al = dot4(x0, x0, 0.f);
      if (NC == 2) al = dot4(x1, x1, al);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) al += __shfl_xor_sync(0xffffffffu, al, o);
      const float nr = sqrtf(al);
      const float inv = nr > 1e-30f ? 1.f / nr : 0.f;           // utils.check_nan: a direction that does not exist is zero
      const int slot_row = crank * 2 * JB + rr;
      float* out = a.X + (size_t)slot_row * R;
      x0.x *= inv; x0.y *= inv; x0.z *= inv; x0.w *= inv;
      x1.x *= inv; x1.y *= inv; x1.z *= inv; x1.w *= inv;
      if (lane * 4 < R) *reinterpret_cast<float4*>(out + lane * 4) = x0;
      if (128 + lane * 4 < R) *reinterpret_cast<float4*>(out + 128 + lane * 4) = x1;
      if (lane == 0) a.sig[slot_row] = nr;
    }
  }
  if (crank == 0 && tid == 0) {
    a.iscal[TRMPS_I_SVD_SWEEPS] = sweeps_used;
    a.iscal[TRMPS_I_N_SWEEPS] += sweeps_used;
  }
}

// ------------------------------------------------------------------------------------------------ 4. ranking, m
struct RankArgs {
  float* sig; const float* scale; int nslots; int block_rows; int R; const int32_t* d_near; const int32_t* d_far; int ncol_blocks;
  float* sv; int32_t* order; int32_t* iscal; int32_t* dim_out; float min_sv; int max_size, min_size;
};

__global__ void __launch_bounds__(SP_NMAX) split_rank_kernel(RankArgs a) {
  __shared__ float sg[SP_NMAX];
  __shared__ int s_count;
  const int tid = threadIdx.x;
  const int n = 2 * *a.d_near;
  const int nb = (n + a.block_rows - 1) / a.block_rows;
  int nbe = nb + (nb & 1);
  if (nbe < 2) nbe = 2;
  const int nslots = min(a.nslots, nbe * a.block_rows);
  if (tid == 0) s_count = 0;
  sg[tid] = tid < nslots ? a.sig[tid] * a.scale[0] : -1.f;
  __syncthreads();
  const int kmax = min(n, a.ncol_blocks * (*a.d_far));
  if (tid < nslots) {
    const float sp = sg[tid];
    int rank = 0;
    for (int o = 0; o < nslots; ++o) {
      const float so = sg[o];
      rank += (so > sp) || (so == sp && o < tid);
    }
    if (rank < a.R) { a.sv[rank] = rank < n ? sp : 0.f; a.order[rank] = rank < n ? tid : -1; }
    if (rank < kmax && sp > a.min_sv) atomicAdd(&s_count, 1);
  }
  for (int p = nslots + tid; p < a.R; p += SP_NMAX) { a.sv[p] = 0.f; a.order[p] = -1; }
  __syncthreads();
  if (tid == 0) {
    int m = s_count;
    if (m < a.min_size) m = a.min_size;
    else if (m > a.max_size) m = a.max_size;
    if (m > n) m = n;
    a.iscal[TRMPS_I_M] = m;
    *a.dim_out = m;
  }
}

// ------------------------------------------------------------------------------------------------ 5. projection
// a_j1[q][col] = sum_p U^T[q][p] M[row(p)][col], q < m, written into the special-node layout; a_j into the node slot.
constexpr int PT = 64, PK = 16, PLD = PT + 4;

struct ProjArgs {
  const float* M; const float* X; const int32_t* order; const int32_t* iscal; const int32_t* d_near;
  int R, Cc, Ds, L, dir; float* node; float* special; int ncol_tiles;
};

__global__ void __launch_bounds__(256) split_project_kernel(ProjArgs a) {
  __shared__ __align__(16) float Us[PK][PLD];
  __shared__ __align__(16) float Ms[PK][PLD];
  const int tid = threadIdx.x;
  const int Ds = a.Ds, dn = *a.d_near, n = 2 * dn, m = a.iscal[TRMPS_I_M];
  if ((int)blockIdx.x >= a.ncol_tiles) {
    // a_j: node[mm][x][y]: right -> (i = x, q = y); left -> (q = x, i = y)                  optimizer.py:305-307, :125
    const int64_t n_node = (int64_t)2 * Ds * Ds;
    const int nblk = gridDim.x - a.ncol_tiles;
    for (int64_t idx = (int64_t)((blockIdx.x - a.ncol_tiles) + blockIdx.y * nblk) * 256 + tid; idx < n_node;
         idx += (int64_t)nblk * gridDim.y * 256) {
      const int y = (int)(idx % Ds), x = (int)((idx / Ds) % Ds), mm = (int)(idx / ((int64_t)Ds * Ds));
      const int i = a.dir > 0 ? x : y, qq = a.dir > 0 ? y : x;
      float v = 0.f;
      if (qq < m && i < dn) v = a.X[(size_t)a.order[qq] * a.R + mm * dn + i];
      a.node[idx] = v;
    }
    return;
  }
  const int q0 = blockIdx.y * PT, col0 = blockIdx.x * PT;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // loaders: U tile: thread -> (q = tid / 4, 4 consecutive p); M tile: thread -> (k = tid / 16, 4 consecutive cols)
  const int uq = tid >> 2, up = (tid & 3) * 4;
  const int qglob = q0 + uq;
  const float* urow = (qglob < m) ? a.X + (size_t)a.order[qglob] * a.R : nullptr;
  const int mk = tid >> 4, mc = (tid & 15) * 4;
  for (int p0 = 0; p0 < n; p0 += PK) {
    float4 u = make_float4(0.f, 0.f, 0.f, 0.f), v = u;
    if (urow && p0 + up < n) u = *reinterpret_cast<const float4*>(urow + p0 + up);    // n even, tail entries of X rows are zero
    const int p = p0 + mk;
    if (p < n && col0 + mc < a.Cc) v = __ldg(reinterpret_cast<const float4*>(a.M + (size_t)compact_to_row(p, dn, Ds) * a.Cc + col0 + mc));
    __syncthreads();
    Us[up][uq] = u.x; Us[up + 1][uq] = u.y; Us[up + 2][uq] = u.z; Us[up + 3][uq] = u.w;
    *reinterpret_cast<float4*>(&Ms[mk][mc]) = v;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&Us[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Ms[k][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }
  // special[ln][x][y]: right -> (q = x, k = y); left -> (k = x, q = y)                         optimizer.py:309-317, :180, :126
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty * 4 + i;
    if (q >= Ds) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = col0 + tx * 4 + j;
      if (col >= a.Cc) continue;
      const int ln = col / Ds, k = col % Ds;
      const float v = q < m ? acc[i][j] : 0.f;
      const size_t o = a.dir > 0 ? ((size_t)ln * Ds + q) * Ds + k : ((size_t)ln * Ds + k) * Ds + q;
      a.special[o] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
bool split_factor_supported(int R) { return R <= SP_NMAX && R % 16 == 0; }

int split_factor(const trmps_sweep* s, const trmps_layout& lay, const BondGeom& g, const trmps_opt* opt, cudaStream_t st,
                 float* special_out, bool single) {
  const int Ds = s->Ds, L = s->L, R = 2 * Ds, Cc = 2 * L * Ds;
  SplitWs w = split_ws(s->work + lay.split_ws, R);
  float* M = s->work + lay.bondM;
  // 1. Gram
  GramArgs ga;
  ga.M = M; ga.R = R; ga.Cc = Cc; ga.Ds = Ds; ga.d_near = g.d_near; ga.gp = w.gp; ga.G = w.G; ga.cnt = w.cnt;
  const int T = (R + GT - 1) / GT, ntiles = T * (T + 1) / 2;
  const int chunks = (Cc + GK - 1) / GK;
  int ks = 160 / ntiles;
  if (ks < 1) ks = 1;
  if (ks > SP_KS_MAX) ks = SP_KS_MAX;
  if (ks > chunks) ks = chunks;
  const int per = (chunks + ks - 1) / ks;
  ga.ks_cols = per * GK;
  ga.KS = (chunks + per - 1) / per;
  // 2. Cholesky: spread over a cluster (split_impl 0, when the device can schedule the cluster), else on one SM
  const int chol_csize = R <= 16 ? 1 : R <= 32 ? 2 : R <= 64 ? 4 : R <= 128 ? 8 : 16;
  static int chol_cluster_ok[17] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
  const size_t chol_smem = chol_cluster_smem(SP_NMAX);
  if (chol_cluster_ok[chol_csize] < 0) {
    chol_cluster_ok[chol_csize] = 0;
    if (cudaFuncSetAttribute(split_chol_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
        cudaFuncSetAttribute(split_chol_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_smem) == cudaSuccess) {
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3(chol_csize); q.blockDim = dim3(CC_T); q.dynamicSmemBytes = chol_smem;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = chol_csize; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa; q.numAttrs = 1;
      int ncl = 0;
      if (cudaOccupancyMaxActiveClusters(&ncl, split_chol_cluster_kernel, &q) == cudaSuccess && ncl > 0) chol_cluster_ok[chol_csize] = 1;
    }
    (void)cudaGetLastError();
  }
  const bool chol_cluster = opt->split_impl == 0 && chol_cluster_ok[chol_csize] == 1;
  // (the cluster Cholesky adds the column slices of the Gram product itself: no last-CTA reduction pass, no G)
  ga.finalize = chol_cluster ? 0 : 1;
  split_gram_kernel<<<dim3(ntiles, ga.KS), 256, 0, st>>>(ga);
  TRMPS_LAUNCH_OK("split_gram_kernel");
  if (chol_cluster) {
    CholCArgs cc;
    cc.G = w.G; cc.Rf = w.Rf; cc.R = R; cc.d_near = g.d_near; cc.scale = w.scale; cc.rowmap = w.rowmap; cc.csize = chol_csize;
    cc.gp = w.gp; cc.KS = ga.KS;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(chol_csize); cfg.blockDim = dim3(CC_T); cfg.dynamicSmemBytes = chol_smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = chol_csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    TRMPS_CUDA_OK(cudaLaunchKernelEx(&cfg, split_chol_cluster_kernel, cc));
    count_launch();
  } else {
  CholArgs ca;
  ca.G = w.G; ca.Rd = w.Rd; ca.Rf = w.Rf; ca.R = R; ca.d_near = g.d_near; ca.scale = w.scale; ca.rowmap = w.rowmap; ca.dbg = s->work + lay.red + 8;
  {
    const size_t smem = (size_t)(2 * CW * SP_NMAX + CW * CT) * sizeof(double);
    static bool attr_done = false;
    if (!attr_done) {
      TRMPS_CUDA_OK(cudaFuncSetAttribute(split_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_done = true;
    }
    split_chol_kernel<<<1, CT, smem, st>>>(ca);
  }
  TRMPS_LAUNCH_OK("split_chol_kernel");
  }
  // 3. Jacobi in one cluster
  JacArgs ja;
  ja.Rf = w.Rf; ja.R = R; ja.rowmap = w.rowmap; ja.d_near = g.d_near; ja.X = w.X; ja.sig = w.sig; ja.iscal = s->iwork;
  ja.tol = 5.9604645e-8f * sqrtf((float)R);
  ja.stop = 1e-3f;              // quadratic convergence: a sweep whose largest |cos| was below 1e-3 leaves O(1e-6) behind
  ja.max_sweeps = opt->max_svd_sweeps > 0 ? (opt->max_svd_sweeps < 40 ? opt->max_svd_sweeps : 40) : 40;
  // rows per block: 8 when the cluster this needs (up to 16 CTAs for R = 256: a non-portable size) can be scheduled, else 16
  const bool gram = opt->split_impl == 3, wide = R > 128;
  static int big_cluster_ok = -1;                 // can a 16-CTA cluster of the 8-row kernel run on this device?
  if (big_cluster_ok < 0) {
    big_cluster_ok = 0;
    const size_t smem8 = (size_t)2 * 2 * 8 * JLD * sizeof(float);
    if (cudaFuncSetAttribute(split_jacobi_kernel<true, 8>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
        cudaFuncSetAttribute(split_jacobi_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8) == cudaSuccess) {
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3(16); q.blockDim = dim3(128); q.dynamicSmemBytes = smem8;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 16; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa; q.numAttrs = 1;
      int ncl = 0;
      if (cudaOccupancyMaxActiveClusters(&ncl, split_jacobi_kernel<true, 8>, &q) == cudaSuccess && ncl > 0) big_cluster_ok = 1;
    }
    (void)cudaGetLastError();
  }
  int br = (gram || opt->split_impl == 2) ? 16 : 8;
  int pairs = ((R + br - 1) / br + 1) / 2;
  if (br == 8 && pairs > 8 && !big_cluster_ok) { br = 16; pairs = ((R + br - 1) / br + 1) / 2; }
  ja.csize = pairs <= 1 ? 1 : pairs <= 2 ? 2 : pairs <= 4 ? 4 : pairs <= 8 ? 8 : 16;
  ja.dbg = s->work + lay.red;
  {
    const size_t smem = (size_t)2 * 2 * br * JLD * sizeof(float) + (gram ? (size_t)G2_KS * G2_TILES * 16 * sizeof(float) : 0);
    static bool attr_done = false;
    if (!attr_done) {
      const int rows16 = (int)((size_t)2 * 2 * 16 * JLD * sizeof(float)), rows8 = rows16 / 2;
      const int gram_b = rows16 + (int)((size_t)G2_KS * G2_TILES * 16 * sizeof(float));
      TRMPS_CUDA_OK(cudaFuncSetAttribute(split_jacobi_kernel<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, rows16));
      TRMPS_CUDA_OK(cudaFuncSetAttribute(split_jacobi_kernel<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, rows16));
      TRMPS_CUDA_OK(cudaFuncSetAttribute(split_jacobi_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, rows8));
      TRMPS_CUDA_OK(cudaFuncSetAttribute(split_jacobi_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, rows8));
      TRMPS_CUDA_OK(cudaFuncSetAttribute(split_jacobi_gram_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, gram_b));
      TRMPS_CUDA_OK(cudaFuncSetAttribute(split_jacobi_gram_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, gram_b));
      attr_done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ja.csize);
    cfg.blockDim = dim3(gram ? G2_T : 16 * br);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = ja.csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (gram) {
      if (wide) TRMPS_CUDA_OK(cudaLaunchKernelEx(&cfg, split_jacobi_gram_kernel<2>, ja));
      else TRMPS_CUDA_OK(cudaLaunchKernelEx(&cfg, split_jacobi_gram_kernel<1>, ja));
    } else if (br == 8) {
      if (wide) TRMPS_CUDA_OK(cudaLaunchKernelEx(&cfg, split_jacobi_kernel<true, 8>, ja));
      else TRMPS_CUDA_OK(cudaLaunchKernelEx(&cfg, split_jacobi_kernel<false, 8>, ja));
    } else {
      if (wide) TRMPS_CUDA_OK(cudaLaunchKernelEx(&cfg, split_jacobi_kernel<true, 16>, ja));
      else TRMPS_CUDA_OK(cudaLaunchKernelEx(&cfg, split_jacobi_kernel<false, 16>, ja));
    }
    count_launch();
  }
  // 4. ranking
  RankArgs ra;
  ra.sig = w.sig; ra.scale = w.scale; ra.nslots = ja.csize * 2 * br; ra.block_rows = br; ra.R = R; ra.d_near = g.d_near; ra.d_far = g.d_far;
  ra.ncol_blocks = single ? L : 2 * L;
  ra.sv = s->work + lay.sv;
  ra.order = reinterpret_cast<int32_t*>(s->work + lay.sv + R);
  ra.iscal = s->iwork;
  ra.dim_out = s->dims + (g.dir > 0 ? g.near_site + 1 : g.near_site);
  ra.min_sv = opt->min_singular_value;
  ra.max_size = opt->max_size < Ds ? opt->max_size : Ds;
  ra.min_size = opt->min_size > 0 ? opt->min_size : 3;
  split_rank_kernel<<<1, SP_NMAX, 0, st>>>(ra);
  TRMPS_LAUNCH_OK("split_rank_kernel");
  // 5. projection + scatter
  ProjArgs pa;
  pa.M = M; pa.X = ja.X; pa.order = ra.order; pa.iscal = s->iwork; pa.d_near = g.d_near;
  pa.R = R; pa.Cc = Cc; pa.Ds = Ds; pa.L = L; pa.dir = g.dir;
  pa.node = s->nodes + (size_t)g.near_site * 2 * Ds * Ds;
  pa.special = special_out ? special_out : s->special;
  pa.ncol_tiles = (Cc + PT - 1) / PT;
  const int qtiles = (Ds + PT - 1) / PT;
  split_project_kernel<<<dim3(pa.ncol_tiles + 8, qtiles), 256, 0, st>>>(pa);
  TRMPS_LAUNCH_OK("split_project_kernel");
  return TRMPS_OK;
}

}  // namespace trmps
