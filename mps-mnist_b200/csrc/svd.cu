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
// One persistent cooperative kernel.  Rows are visited in order of decreasing norm (de Rijk) and grouped
// in blocks of NR/2; each CTA owns one pair of blocks per round (round-robin tournament over blocks) and
// keeps its NR rows in REGISTERS, each thread holding a column slice of every row.  A step rotates NR/2
// disjoint row pairs at once: per-thread partial dot products, one deterministic block reduction, the
// rotation parameters computed once per pair and broadcast through shared memory, then the rotation is
// applied to the register-resident slices.  Cross-block pairs are rotated every round, pairs inside a
// block once per sweep.  Squared norms are carried through the rotations and refreshed from the data
// every round.  No floating-point atomics anywhere: every rank of a data-parallel job computes
// bit-identical factors from the same all-reduced bond.
#include <cooperative_groups.h>
#include <type_traits>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace trmps {

constexpr int SVD_THREADS = 512;
constexpr int SVD_WARPS = SVD_THREADS / 32;

struct SvdArgs {
  float* M; float* Ut; int R, Cc, Ds; const int32_t* d_near; int32_t* iscal; int32_t* rowmap; float* norms;
  float tol; int max_sweeps;
};

__device__ __forceinline__ int compact_to_row(int p, int dn, int Ds) { return (p / dn) * Ds + (p % dn); }

// squared norms of the active rows (one warp per row)
__global__ void svd_norms_kernel(SvdArgs a) {
  const int dn = *a.d_near, Ra = 2 * dn;
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int p = gw; p < a.R; p += nw) {
    float s = 0.f;
    if (p < Ra) {
      const float4* row = reinterpret_cast<const float4*>(a.M + (size_t)compact_to_row(p, dn, a.Ds) * a.Cc);
      for (int c4 = lane; c4 < a.Cc / 4; c4 += 32) {
        float4 v = row[c4];
        s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    if (lane == 0) a.norms[p] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x < 64) a.iscal[TRMPS_I_SVD_ROT + threadIdx.x] = 0;
}

// rowmap[rank] = padded row index of the rank-th largest row; U^T = permutation matrix
__global__ void svd_init_kernel(SvdArgs a) {
  const int dn = *a.d_near, Ra = 2 * dn;
  if (blockIdx.x == 0) {
    for (int p = threadIdx.x; p < a.R; p += blockDim.x) {
      if (p < Ra) {
        float sp = a.norms[p];
        int rank = 0;
        for (int o = 0; o < Ra; ++o) {
          float so = a.norms[o];
          rank += (so > sp) || (so == sp && o < p);
        }
        a.rowmap[rank] = compact_to_row(p, dn, a.Ds);
      }
    }
  }
  const int64_t total = (int64_t)a.R * a.R;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    a.Ut[i] = 0.f;
}

__global__ void svd_init_ut_kernel(SvdArgs a) {
  const int dn = *a.d_near, Ra = 2 * dn;
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < Ra) a.Ut[(size_t)p * a.R + a.rowmap[p]] = 1.f;
}

// round-robin pairing of n players (n even): partner pair #j in step s
__host__ __device__ constexpr int rr_first(int n, int s, int j) { return j == 0 ? n - 1 : (s + j) % (n - 1); }
__host__ __device__ constexpr int rr_second(int n, int s, int j) { return j == 0 ? s : (s + n - 1 - j) % (n - 1); }

// local row indices of pair p in a step.  kind 0: cross pairs (row p of block I with row (p+s)%BLK of block J);
// kind 1: pairs inside the blocks (first BLK/2 pairs in block I, the rest in block J)
template <int BLK>
__host__ __device__ constexpr int pair_x(int kind, int s, int p) {
  return kind == 0 ? p : (p < BLK / 2 ? rr_first(BLK, s, p) : BLK + rr_first(BLK, s, p - BLK / 2));
}
template <int BLK>
__host__ __device__ constexpr int pair_y(int kind, int s, int p) {
  return kind == 0 ? BLK + (p + s) % BLK : (p < BLK / 2 ? rr_second(BLK, s, p) : BLK + rr_second(BLK, s, p - BLK / 2));
}

// per-warp partial sums of K per-thread values -> part[warp][k]  (first stage of a deterministic block reduction).
// Transposing butterfly: each exchange halves the number of live values, so K values cost K-1+log2(32/K)
// shuffles instead of 5K.  K must be a power of two <= 32.
template <int K>
__device__ __forceinline__ void warp_partials(float (&x)[K], float (*part)[16], int lane, int warp) {
  int idx = 0;
  int o = 16;
#pragma unroll
  for (int n = K; n > 1; n >>= 1) {
    const bool hi = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float keep = hi ? x[i + n / 2] : x[i];
      const float send = hi ? x[i] : x[i + n / 2];
      x[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
    if (hi) idx += n / 2;
    o >>= 1;
  }
  const int rest = 2 * o - 1;   // lanes that differ only in these bits hold partials of the same value
  for (; o > 0; o >>= 1) x[0] += __shfl_xor_sync(0xffffffffu, x[0], o);
  if ((lane & rest) == 0) part[warp][idx] = x[0];
  __syncthreads();
}

template <int NR, int SLOTS>
__global__ void __launch_bounds__(SVD_THREADS) svd_jacobi_kernel(SvdArgs a) {
  cg::grid_group grid = cg::this_grid();
  constexpr int BLK = NR / 2, NP = BLK;
  extern __shared__ __align__(16) float uts[];   // [NR][R]
  __shared__ float part[SVD_WARPS][16];
  __shared__ float nrm[NR];
  __shared__ float rot_c[NP], rot_s[NP];
  __shared__ int srow[NR];
  __shared__ int sflag;
  __shared__ float smax;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Cc = a.Cc, R = a.R;
  const int dn = *a.d_near, Ra = 2 * dn;
  const int nb = (Ra + BLK - 1) / BLK;
  int nbe = nb + (nb & 1);
  if (nbe < 2) nbe = 2;
  const int rounds = nbe - 1;
  const int q = blockIdx.x;
  const bool active = q < nbe / 2;
  int sweeps_used = 0;
  float v[NR][SLOTS];

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
        if (tid == 0) { sflag = 0; smax = 0.f; }
        __syncthreads();
        // load the NR rows: thread owns columns tid + SVD_THREADS * slot
#pragma unroll
        for (int x = 0; x < NR; ++x) {
          const int p = srow[x];
          const float* src = p >= 0 ? a.M + (size_t)__ldcg(&a.rowmap[p]) * Cc : nullptr;
#pragma unroll
          for (int sl = 0; sl < SLOTS; ++sl) {
            int c = tid + SVD_THREADS * sl;
            v[x][sl] = (src && c < Cc) ? __ldcg(src + c) : 0.f;
          }
          if (tid < R) uts[x * R + tid] = p >= 0 ? __ldcg(a.Ut + (size_t)p * R + tid) : 0.f;
        }
        // refresh the squared norms from the data
        {
          float n2[NR];
#pragma unroll
          for (int x = 0; x < NR; ++x) {
            float s = 0.f;
#pragma unroll
            for (int sl = 0; sl < SLOTS; ++sl) s = fmaf(v[x][sl], v[x][sl], s);
            n2[x] = s;
          }
          warp_partials<NR>(n2, part, lane, warp);
          if (tid < NR * SVD_WARPS) {
            int j = tid / SVD_WARPS, w = tid % SVD_WARPS;
            float s = part[w][j];
#pragma unroll
            for (int o = SVD_WARPS / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (w == 0) nrm[j] = s;
          }
          __syncthreads();
        }

        auto do_steps = [&](auto kind_tag, int nsteps) {
          constexpr int KIND = decltype(kind_tag)::value;
#pragma unroll 1
          for (int step = 0; step < nsteps; ++step) {
            float g[NP];
            // (the row indices must be compile-time for the register file: switch over the step)
            auto dots = [&](auto step_tag) {
              constexpr int S = decltype(step_tag)::value;
#pragma unroll
              for (int p = 0; p < NP; ++p) {
                const int x = pair_x<BLK>(KIND, S, p), y = pair_y<BLK>(KIND, S, p);
                float s = 0.f;
#pragma unroll
                for (int sl = 0; sl < SLOTS; ++sl) s = fmaf(v[x][sl], v[y][sl], s);
                g[p] = s;
              }
            };
            auto apply = [&](auto step_tag) {
              constexpr int S = decltype(step_tag)::value;
#pragma unroll
              for (int p = 0; p < NP; ++p) {
                const int x = pair_x<BLK>(KIND, S, p), y = pair_y<BLK>(KIND, S, p);
                // Rutishauser's form x' = x - s (y + tau x), y' = y + s (x - tau y), tau = s / (1 + c): the
                // second-order term keeps the rotation norm-preserving even when c rounds to 1 in fp32
                const float tau = rot_c[p], s = rot_s[p];
                if (s != 0.f) {
#pragma unroll
                  for (int sl = 0; sl < SLOTS; ++sl) {
                    float u = v[x][sl], w = v[y][sl];
                    v[x][sl] = fmaf(-s, fmaf(tau, u, w), u);
                    v[y][sl] = fmaf(s, fmaf(-tau, w, u), w);
                  }
                  if (tid < R) {
                    float u = uts[x * R + tid], w = uts[y * R + tid];
                    uts[x * R + tid] = fmaf(-s, fmaf(tau, u, w), u);
                    uts[y * R + tid] = fmaf(s, fmaf(-tau, w, u), w);
                  }
                }
              }
            };
            auto dispatch = [&](auto&& fn) {
              switch (step) {
                case 0: fn(std::integral_constant<int, 0>{}); break;
                case 1: fn(std::integral_constant<int, 1>{}); break;
                case 2: if constexpr (BLK > 2) fn(std::integral_constant<int, 2>{}); break;
                case 3: if constexpr (BLK > 2) fn(std::integral_constant<int, 3>{}); break;
                case 4: if constexpr (BLK > 4) fn(std::integral_constant<int, 4>{}); break;
                case 5: if constexpr (BLK > 4) fn(std::integral_constant<int, 5>{}); break;
                case 6: if constexpr (BLK > 4) fn(std::integral_constant<int, 6>{}); break;
                case 7: if constexpr (BLK > 4) fn(std::integral_constant<int, 7>{}); break;
                default: break;
              }
            };
            dispatch(dots);
            warp_partials<NP>(g, part, lane, warp);
            if (tid < NP * SVD_WARPS) {
              const int j = tid / SVD_WARPS, w = tid % SVD_WARPS;
              float ga = part[w][j];
#pragma unroll
              for (int o = SVD_WARPS / 2; o > 0; o >>= 1) ga += __shfl_xor_sync(0xffffffffu, ga, o);
              if (w == 0) {
                const int x = pair_x<BLK>(KIND, step, j), y = pair_y<BLK>(KIND, step, j);
                const float al = nrm[x], be = nrm[y];
                float tau = 0.f, s = 0.f;
                const float big = fmaxf(al, be), small = fminf(al, be);
                if (small > 1e-30f && small > 1e-24f * big) {
                  const float relg = fabsf(ga) / (sqrtf(al) * sqrtf(be));
                  if (relg > a.tol) {
                    const float zeta = (be - al) / (2.f * ga);
                    const float t = copysignf(1.f, zeta) / (fabsf(zeta) + sqrtf(fmaf(zeta, zeta, 1.f)));
                    const float c = 1.f / sqrtf(fmaf(t, t, 1.f));
                    s = c * t;
                    tau = s / (1.f + c);
                    nrm[x] = al - t * ga;
                    nrm[y] = be + t * ga;
                    sflag = 1;
                    atomicMax(reinterpret_cast<int*>(&smax), __float_as_int(relg));   // relg >= 0: int order == float order
                  }
                }
                rot_c[j] = tau;
                rot_s[j] = s;
              }
            }
            __syncthreads();
            dispatch(apply);
          }
        };
        if (r == 0) do_steps(std::integral_constant<int, 1>{}, BLK - 1);   // pairs inside each block: once per sweep
        do_steps(std::integral_constant<int, 0>{}, BLK);                     // cross-block pairs
        __syncthreads();
        // write back
#pragma unroll
        for (int x = 0; x < NR; ++x) {
          const int p = srow[x];
          if (p >= 0) {
            float* dst = a.M + (size_t)__ldcg(&a.rowmap[p]) * Cc;
#pragma unroll
            for (int sl = 0; sl < SLOTS; ++sl) {
              int c = tid + SVD_THREADS * sl;
              if (c < Cc) __stcg(dst + c, v[x][sl]);
            }
            if (tid < R) __stcg(a.Ut + (size_t)p * R + tid, uts[x * R + tid]);
          }
        }
        if (tid == 0 && sflag) {
          atomicAdd(&a.iscal[TRMPS_I_SVD_ROT + sweep], 1);
          atomicMax(&a.iscal[TRMPS_I_SVD_ROT + 32 + (sweep & 31)], __float_as_int(smax));
        }
      }
      grid.sync();
    }
    sweeps_used = sweep + 1;
    const int nrot = __ldcg(&a.iscal[TRMPS_I_SVD_ROT + sweep]);
    const float worst = __int_as_float(__ldcg(&a.iscal[TRMPS_I_SVD_ROT + 32 + (sweep & 31)]));
    // quadratic convergence: a sweep whose largest relative off-diagonal was eps leaves O(eps^2) behind
    if (nrot == 0 || worst < 2e-4f) break;
  }
  if (blockIdx.x == 0 && tid == 0) a.iscal[TRMPS_I_SVD_SWEEPS] = sweeps_used;
}

// sigma, ordering, m, dims update                                     optimizer.py:294-302
struct SvdFinArgs {
  const float* M; int R, Cc, Ds; const int32_t* d_near; const int32_t* d_far; int L; const int32_t* rowmap;
  float* sv; int32_t* order; int32_t* iscal; int32_t* dim_out;
  float min_sv; int max_size, min_size;
};

__global__ void __launch_bounds__(1024) svd_finalize_kernel(SvdFinArgs a) {
  extern __shared__ float sig[];   // [R]
  __shared__ int s_count;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int dn = *a.d_near, Ra = 2 * dn;
  if (tid == 0) s_count = 0;
  for (int p = w; p < Ra; p += 32) {
    const float4* row = reinterpret_cast<const float4*>(a.M + (size_t)a.rowmap[p] * a.Cc);
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
  const int kmax = min(Ra, 2 * a.L * (*a.d_far));
  int local = 0;
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
      if (rank < kmax && sp > a.min_sv) ++local;
    } else {
      a.sv[p] = 0.f;
      a.order[p] = -1;
    }
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
  const int32_t* rowmap; const int32_t* iscal; float* node; float* special;
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
      if (qq < m) v = a.M[(size_t)a.rowmap[a.order[qq]] * a.Cc + (size_t)ln * Ds + k];
      a.special[j] = v;
    }
  }
}

// rows per Jacobi block (NR/2) for a given bond-matrix shape; 0 = unsupported
int svd_block_rows(int R, int Cc) {
  if (R > SVD_THREADS) return 0;
  const int slots = (Cc + SVD_THREADS - 1) / SVD_THREADS;
  if (slots <= 5) return 8;
  if (slots <= 10) return 4;
  return 0;
}

template <int NR, int SLOTS>
static int launch_jacobi(SvdArgs& a, int ncta, cudaStream_t st) {
  size_t smem = (size_t)NR * a.R * sizeof(float);
  void* kargs[] = {&a};
  const void* fn = (const void*)svd_jacobi_kernel<NR, SLOTS>;
  TRMPS_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TRMPS_CUDA_OK(cudaLaunchCooperativeKernel(fn, dim3(ncta), dim3(SVD_THREADS), kargs, smem, st));
  return TRMPS_OK;
}

int svd_split(const trmps_sweep* s, const trmps_layout& lay, const BondGeom& g, const trmps_opt* opt, cudaStream_t st) {
  const int Ds = s->Ds, L = s->L, R = 2 * Ds, Cc = 2 * L * Ds;
  const int blk = lay.svd_block;
  if (blk == 0) {
    set_error("svd: bond stride Ds=%d (R=%d, Cc=%d) is outside the supported range (Cc <= 5120, R <= 512)", Ds, R, Cc);
    return TRMPS_E_ARG;
  }
  SvdArgs a;
  a.M = s->work + lay.bondM;
  a.Ut = s->work + lay.Ut;
  a.R = R; a.Cc = Cc; a.Ds = Ds; a.d_near = g.d_near; a.iscal = s->iwork;
  a.norms = s->work + lay.sv;
  a.rowmap = reinterpret_cast<int32_t*>(s->work + lay.sv + 2 * R);
  a.tol = 5.9604645e-8f * sqrtf((float)Cc);
  a.max_sweeps = opt->max_svd_sweeps > 0 ? (opt->max_svd_sweeps < 30 ? opt->max_svd_sweeps : 30) : 30;
  svd_norms_kernel<<<(R * 32 + 255) / 256, 256, 0, st>>>(a);
  TRMPS_LAUNCH_OK("svd_norms_kernel");
  svd_init_kernel<<<32, 256, 0, st>>>(a);
  TRMPS_LAUNCH_OK("svd_init_kernel");
  svd_init_ut_kernel<<<(R + 255) / 256, 256, 0, st>>>(a);
  TRMPS_LAUNCH_OK("svd_init_ut_kernel");

  const int nb = (R + blk - 1) / blk;
  const int ncta = ((nb + (nb & 1)) / 2) > 1 ? (nb + (nb & 1)) / 2 : 1;
  const int slots = (Cc + SVD_THREADS - 1) / SVD_THREADS;
  int rc;
  if (blk == 8) {
    if (slots <= 1) rc = launch_jacobi<16, 1>(a, ncta, st);
    else if (slots <= 2) rc = launch_jacobi<16, 2>(a, ncta, st);
    else if (slots <= 3) rc = launch_jacobi<16, 3>(a, ncta, st);
    else rc = launch_jacobi<16, 5>(a, ncta, st);
  } else {
    if (slots <= 8) rc = launch_jacobi<8, 8>(a, ncta, st);
    else rc = launch_jacobi<8, 10>(a, ncta, st);
  }
  if (rc) return rc;

  SvdFinArgs f;
  f.M = a.M; f.R = R; f.Cc = Cc; f.Ds = Ds; f.d_near = g.d_near; f.d_far = g.d_far; f.L = L; f.rowmap = a.rowmap;
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
  c.order = f.order; c.rowmap = a.rowmap; c.iscal = s->iwork;
  c.node = s->nodes + (size_t)g.near_site * 2 * Ds * Ds;
  c.special = s->special;
  svd_scatter_kernel<<<296, 256, 0, st>>>(c);
  TRMPS_LAUNCH_OK("svd_scatter_kernel");
  return TRMPS_OK;
}

}  // namespace trmps
