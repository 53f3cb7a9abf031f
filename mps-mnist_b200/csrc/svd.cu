// Truncated bond split by an on-device one-sided (Hestenes) Jacobi SVD  (north_star piece 4).
// Replaces MPSOptimizer._bond_decomposition (optimizer.py:266-321): transpose [1,3,0,2,4], tf.svd,
// m = clamp(#{sigma > min_singular_value}, 3, max_size), a_j = U[:, :m], a_j1 = diag(s[:m]) V[:, :m]^T,
// plus the layout fix-ups of _update_right/_update_left (optimizer.py:180, :125-126).
//
// The bond matrix M (R x Cc) already has the SVD layout.  Rows of M are orthogonalised in place by plane
// rotations, M <- J^T M, accumulating U^T <- J^T U^T.  At convergence row p of M is sigma_p v_p^T and
// row p of U^T is u_p^T, so a_j1 is read straight from the rows of M (no division by sigma, hence none of
// the NaNs utils.check_nan guards against, utils.py:9-20) and a_j from the rows of U^T.
//
// One persistent cooperative kernel: rows are grouped in blocks of BLK; each CTA owns one pair of blocks
// per round (round-robin tournament over blocks), keeps its 2*BLK rows in shared memory, rotates every
// row pair among them (inner tournament, BLK pairs at a time), writes the rows back and meets the other
// CTAs at a grid barrier.  No floating-point atomics: every rank of a data-parallel job computes
// bit-identical factors from the same all-reduced bond.
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace trmps {

struct SvdArgs {
  float* M; float* Ut; int R, Cc, Ds; const int32_t* d_near; int32_t* iscal; float tol; int max_sweeps;
};

__device__ __forceinline__ int compact_to_row(int p, int dn, int Ds) { return (p / dn) * Ds + (p % dn); }

__global__ void svd_init_kernel(SvdArgs a) {
  const int dn = *a.d_near, Ra = 2 * dn;
  const int64_t total = (int64_t)a.R * a.R;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int p = (int)(i / a.R), c = (int)(i % a.R);
    a.Ut[i] = (p < Ra && c == compact_to_row(p, dn, a.Ds)) ? 1.f : 0.f;
  }
  if (blockIdx.x == 0 && threadIdx.x < 64) a.iscal[TRMPS_I_SVD_ROT + threadIdx.x] = 0;
}

template <int BLK>
__global__ void __launch_bounds__(512) svd_jacobi_kernel(SvdArgs a) {
  cg::grid_group grid = cg::this_grid();
  constexpr int NR = 2 * BLK, NG = BLK, TPG = 512 / BLK, WPG = TPG / 32;
  extern __shared__ __align__(16) float smem[];
  const int Cc = a.Cc, R = a.R;
  float* rows = smem;                    // [NR][Cc]
  float* uts = rows + (size_t)NR * Cc;   // [NR][R]
  __shared__ float sred[NG][WPG][3];
  __shared__ int srow[NR];
  __shared__ int sflag;
  const int tid = threadIdx.x, g = tid / TPG, gt = tid % TPG, lane = tid & 31, wg = gt >> 5;
  const int dn = *a.d_near, Ra = 2 * dn;
  const int nb = (Ra + BLK - 1) / BLK;
  int nbe = nb + (nb & 1);
  if (nbe < 2) nbe = 2;
  const int rounds = nbe - 1;
  const int q = blockIdx.x;
  const bool active = q < nbe / 2;
  const int Cc4 = Cc / 4, R4 = R / 4;
  int sweeps_used = 0;

  for (int sweep = 0; sweep < a.max_sweeps; ++sweep) {
    for (int r = 0; r < rounds; ++r) {
      if (active) {
        int bi, bj;
        if (q == 0) { bi = nbe - 1; bj = r; }
        else { bi = (r + q) % (nbe - 1); bj = (r + nbe - 1 - q) % (nbe - 1); }
        if (tid < NR) {
          int blk = tid < BLK ? bi : bj;
          int p = blk * BLK + (tid % BLK);
          srow[tid] = (blk < nb && p < Ra) ? p : -1;
        }
        if (tid == 0) sflag = 0;
        __syncthreads();
        for (int e = tid; e < NR * Cc4; e += 512) {
          int x = e / Cc4, c4 = e - x * Cc4;
          int p = srow[x];
          if (p >= 0)
            reinterpret_cast<float4*>(rows + (size_t)x * Cc)[c4] =
                __ldcg(reinterpret_cast<const float4*>(a.M + (size_t)compact_to_row(p, dn, a.Ds) * Cc) + c4);
        }
        for (int e = tid; e < NR * R4; e += 512) {
          int x = e / R4, c4 = e - x * R4;
          int p = srow[x];
          if (p >= 0)
            reinterpret_cast<float4*>(uts + (size_t)x * R)[c4] = __ldcg(reinterpret_cast<const float4*>(a.Ut + (size_t)p * R) + c4);
        }
        __syncthreads();
        for (int step = 0; step < NR - 1; ++step) {
          int x, y;
          if (g == 0) { x = NR - 1; y = step; }
          else { x = (step + g) % (NR - 1); y = (step + NR - 1 - g) % (NR - 1); }
          const bool valid = srow[x] >= 0 && srow[y] >= 0;
          float4* rx = reinterpret_cast<float4*>(rows + (size_t)x * Cc);
          float4* ry = reinterpret_cast<float4*>(rows + (size_t)y * Cc);
          float al = 0.f, be = 0.f, ga = 0.f;
          if (valid) {
            for (int c4 = gt; c4 < Cc4; c4 += TPG) {
              float4 u = rx[c4], v = ry[c4];
              al = fmaf(u.x, u.x, al); al = fmaf(u.y, u.y, al); al = fmaf(u.z, u.z, al); al = fmaf(u.w, u.w, al);
              be = fmaf(v.x, v.x, be); be = fmaf(v.y, v.y, be); be = fmaf(v.z, v.z, be); be = fmaf(v.w, v.w, be);
              ga = fmaf(u.x, v.x, ga); ga = fmaf(u.y, v.y, ga); ga = fmaf(u.z, v.z, ga); ga = fmaf(u.w, v.w, ga);
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            al += __shfl_xor_sync(0xffffffffu, al, o);
            be += __shfl_xor_sync(0xffffffffu, be, o);
            ga += __shfl_xor_sync(0xffffffffu, ga, o);
          }
          if (lane == 0) { sred[g][wg][0] = al; sred[g][wg][1] = be; sred[g][wg][2] = ga; }
          asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(TPG) : "memory");
          al = 0.f; be = 0.f; ga = 0.f;
#pragma unroll
          for (int w = 0; w < WPG; ++w) { al += sred[g][w][0]; be += sred[g][w][1]; ga += sred[g][w][2]; }
          const float big = fmaxf(al, be), small = fminf(al, be);
          if (valid && small > 1e-30f && small > 1e-24f * big && fabsf(ga) > a.tol * sqrtf(al) * sqrtf(be)) {
            const float zeta = (be - al) / (2.f * ga);
            const float t = copysignf(1.f, zeta) / (fabsf(zeta) + sqrtf(fmaf(zeta, zeta, 1.f)));
            const float c = 1.f / sqrtf(fmaf(t, t, 1.f));
            const float s = c * t;
            for (int c4 = gt; c4 < Cc4; c4 += TPG) {
              float4 u = rx[c4], v = ry[c4], nu, nv;
              nu.x = c * u.x - s * v.x; nv.x = fmaf(s, u.x, c * v.x);
              nu.y = c * u.y - s * v.y; nv.y = fmaf(s, u.y, c * v.y);
              nu.z = c * u.z - s * v.z; nv.z = fmaf(s, u.z, c * v.z);
              nu.w = c * u.w - s * v.w; nv.w = fmaf(s, u.w, c * v.w);
              rx[c4] = nu; ry[c4] = nv;
            }
            float4* ux = reinterpret_cast<float4*>(uts + (size_t)x * R);
            float4* uy = reinterpret_cast<float4*>(uts + (size_t)y * R);
            for (int c4 = gt; c4 < R4; c4 += TPG) {
              float4 u = ux[c4], v = uy[c4], nu, nv;
              nu.x = c * u.x - s * v.x; nv.x = fmaf(s, u.x, c * v.x);
              nu.y = c * u.y - s * v.y; nv.y = fmaf(s, u.y, c * v.y);
              nu.z = c * u.z - s * v.z; nv.z = fmaf(s, u.z, c * v.z);
              nu.w = c * u.w - s * v.w; nv.w = fmaf(s, u.w, c * v.w);
              ux[c4] = nu; uy[c4] = nv;
            }
            if (gt == 0) sflag = 1;
          }
          __syncthreads();
        }
        for (int e = tid; e < NR * Cc4; e += 512) {
          int x = e / Cc4, c4 = e - x * Cc4;
          int p = srow[x];
          if (p >= 0)
            __stcg(reinterpret_cast<float4*>(a.M + (size_t)compact_to_row(p, dn, a.Ds) * Cc) + c4,
                   reinterpret_cast<const float4*>(rows + (size_t)x * Cc)[c4]);
        }
        for (int e = tid; e < NR * R4; e += 512) {
          int x = e / R4, c4 = e - x * R4;
          int p = srow[x];
          if (p >= 0) __stcg(reinterpret_cast<float4*>(a.Ut + (size_t)p * R) + c4, reinterpret_cast<const float4*>(uts + (size_t)x * R)[c4]);
        }
        if (tid == 0 && sflag) atomicAdd(&a.iscal[TRMPS_I_SVD_ROT + sweep], 1);
      }
      grid.sync();
    }
    sweeps_used = sweep + 1;
    const int nrot = __ldcg(&a.iscal[TRMPS_I_SVD_ROT + sweep]);
    if (nrot == 0) break;
  }
  if (blockIdx.x == 0 && tid == 0) a.iscal[TRMPS_I_SVD_SWEEPS] = sweeps_used;
}

// sigma, ordering, m, dims update                                     optimizer.py:294-302
struct SvdFinArgs {
  const float* M; int R, Cc, Ds; const int32_t* d_near; const int32_t* d_far; int L;
  float* sv; int32_t* order; int32_t* iscal; int32_t* dim_out;
  float min_sv; int max_size, min_size;
};

__global__ void __launch_bounds__(1024) svd_finalize_kernel(SvdFinArgs a) {
  extern __shared__ float sig[];   // [R]
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int dn = *a.d_near, Ra = 2 * dn;
  for (int p = w; p < Ra; p += 32) {
    const float4* row = reinterpret_cast<const float4*>(a.M + (size_t)compact_to_row(p, dn, a.Ds) * a.Cc);
    float s = 0.f;
    for (int c4 = lane; c4 < a.Cc / 4; c4 += 32) {
      float4 v = row[c4];
      s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) sig[p] = sqrtf(s);
  }
  __syncthreads();
  for (int p = tid; p < a.R; p += 1024) {
    if (p < Ra) {
      float sp = sig[p];
      int rank = 0;
      for (int o = 0; o < Ra; ++o) {
        float so = sig[o];
        rank += (so > sp) || (so == sp && o < p);
      }
      a.sv[rank] = sp;
      a.order[rank] = p;
    } else {
      a.sv[p] = 0.f;
      a.order[p] = -1;
    }
  }
  __syncthreads();
  __shared__ int s_count;
  if (tid == 0) s_count = 0;
  __syncthreads();
  const int kmax = min(Ra, 2 * a.L * (*a.d_far));
  int local = 0;
  for (int p = tid; p < Ra; p += 1024) {
    float sp = sig[p];
    int rank = 0;
    for (int o = 0; o < Ra; ++o) {
      float so = sig[o];
      rank += (so > sp) || (so == sp && o < p);
    }
    if (rank < kmax && sp > a.min_sv) ++local;
  }
  if (local) atomicAdd(&s_count, local);
  __syncthreads();
  if (tid == 0) {
    int m = s_count;
    if (m < a.min_size) m = a.min_size;
    else if (m > a.max_size) m = a.max_size;
    if (m > Ra) m = Ra;
    a.iscal[TRMPS_I_M] = m;
    *a.dim_out = m;
  }
}

// write a_j into the near site's node slot and a_j1 into the special node    optimizer.py:305-317, :180, :125-126
struct SvdScatterArgs {
  const float* M; const float* Ut; int R, Cc, Ds, L, dir; const int32_t* d_near; const int32_t* order;
  const int32_t* iscal; float* node; float* special;
};

__global__ void svd_scatter_kernel(SvdScatterArgs a) {
  const int Ds = a.Ds, dn = *a.d_near, m = a.iscal[TRMPS_I_M];
  const int64_t n_node = (int64_t)2 * Ds * Ds, n_sp = (int64_t)2 * a.L * Ds * Ds;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n_node + n_sp;
       idx += (int64_t)gridDim.x * blockDim.x) {
    if (idx < n_node) {
      // node[mm][x][y]: right -> (i = x, q = y); left -> (q = x, i = y)
      int y = (int)(idx % Ds), x = (int)((idx / Ds) % Ds), mm = (int)(idx / ((int64_t)Ds * Ds));
      int i = a.dir > 0 ? x : y, qq = a.dir > 0 ? y : x;
      float v = 0.f;
      if (qq < m && i < dn) v = a.Ut[(size_t)a.order[qq] * a.R + mm * Ds + i];
      a.node[idx] = v;
    } else {
      int64_t j = idx - n_node;
      // special[l][n][x][y]: right -> (q = x, k = y); left -> (k = x, q = y)
      int y = (int)(j % Ds), x = (int)((j / Ds) % Ds), ln = (int)(j / ((int64_t)Ds * Ds));
      int k = a.dir > 0 ? y : x, qq = a.dir > 0 ? x : y;
      float v = 0.f;
      if (qq < m) v = a.M[(size_t)compact_to_row(a.order[qq], dn, Ds) * a.Cc + (size_t)ln * Ds + k];
      a.special[j] = v;
    }
  }
}

static size_t svd_smem(int blk, int R, int Cc) { return (size_t)2 * blk * ((size_t)Cc + R) * sizeof(float); }

int svd_block_rows(int R, int Cc) {
  const size_t lim = 200 * 1024;
  if (svd_smem(8, R, Cc) <= lim) return 8;
  if (svd_smem(4, R, Cc) <= lim) return 4;
  if (svd_smem(2, R, Cc) <= lim) return 2;
  return 0;
}

int svd_split(const trmps_sweep* s, const trmps_layout& lay, const BondGeom& g, const trmps_opt* opt, cudaStream_t st) {
  const int Ds = s->Ds, L = s->L, R = 2 * Ds, Cc = 2 * L * Ds;
  SvdArgs a;
  a.M = s->work + lay.bondM;
  a.Ut = s->work + lay.Ut;
  a.R = R; a.Cc = Cc; a.Ds = Ds; a.d_near = g.d_near; a.iscal = s->iwork;
  a.tol = 5.9604645e-8f * sqrtf((float)Cc);
  a.max_sweeps = opt->max_svd_sweeps > 0 ? (opt->max_svd_sweeps < 60 ? opt->max_svd_sweeps : 60) : 30;
  svd_init_kernel<<<64, 256, 0, st>>>(a);
  TRMPS_LAUNCH_OK("svd_init_kernel");

  const int blk = lay.svd_block;
  if (blk == 0) {
    set_error("svd: bond stride Ds=%d does not fit shared memory", Ds);
    return TRMPS_E_ARG;
  }
  const int nb = (R + blk - 1) / blk;
  const int ncta = ((nb + (nb & 1)) / 2) > 1 ? (nb + (nb & 1)) / 2 : 1;
  size_t smem = svd_smem(blk, R, Cc);
  void* kargs[] = {&a};
  const void* fn = blk == 8 ? (const void*)svd_jacobi_kernel<8> : blk == 4 ? (const void*)svd_jacobi_kernel<4> : (const void*)svd_jacobi_kernel<2>;
  TRMPS_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TRMPS_CUDA_OK(cudaLaunchCooperativeKernel(fn, dim3(ncta), dim3(512), kargs, smem, st));

  SvdFinArgs f;
  f.M = a.M; f.R = R; f.Cc = Cc; f.Ds = Ds; f.d_near = g.d_near; f.d_far = g.d_far; f.L = L;
  f.sv = s->work + lay.sv;
  f.order = reinterpret_cast<int32_t*>(s->work + lay.sv + R);
  f.iscal = s->iwork;
  f.dim_out = s->dims + (g.dir > 0 ? g.near_site + 1 : g.near_site);
  f.min_sv = opt->min_singular_value;
  f.max_size = opt->max_size < Ds ? opt->max_size : Ds;
  f.min_size = opt->min_size > 0 ? opt->min_size : 3;
  svd_finalize_kernel<<<1, 1024, R * sizeof(float), st>>>(f);
  TRMPS_LAUNCH_OK("svd_finalize_kernel");

  SvdScatterArgs c;
  c.M = a.M; c.Ut = a.Ut; c.R = R; c.Cc = Cc; c.Ds = Ds; c.L = L; c.dir = g.dir; c.d_near = g.d_near;
  c.order = f.order; c.iscal = s->iwork;
  c.node = s->nodes + (size_t)g.near_site * 2 * Ds * Ds;
  c.special = s->special;
  svd_scatter_kernel<<<296, 256, 0, st>>>(c);
  TRMPS_LAUNCH_OK("svd_scatter_kernel");
  return TRMPS_OK;
}

}  // namespace trmps
