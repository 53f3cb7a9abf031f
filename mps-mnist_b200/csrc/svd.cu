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
#include "svd_common.cuh"

namespace cg = cooperative_groups;

namespace trmps {

constexpr int SVD_THREADS = 512;
constexpr int SVD_WARPS = SVD_THREADS / 32;

struct SvdArgs {
  float* M; float* Ut; int R, Cc, Ds; const int32_t* d_near; int32_t* iscal; int32_t* rowmap; float* norms;
  float* norms2;     // second buffer of the current squared norms: sweep s writes buffer (s + 1) & 1 and reads buffer s & 1 (no CTA reads what another one is rewriting)
  float tol; int max_sweeps; int keep; float* dbg;   // keep: rows that can survive the truncation (max_size); dbg: 8 floats, phase cycle counters of CTA 0 (load, gram, rotate, replay, store, grid sync)
};


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


// per-warp partial sums of K per-thread values -> part[warp][k]  (first stage of a deterministic block reduction).
// Transposing butterfly: each exchange halves the number of live values, so K values cost K-1+log2(32/K)
// shuffles instead of 5K.  K must be a power of two <= 32.  No barrier inside.
template <int K>
__device__ __forceinline__ void warp_partials(float (&x)[K], float (*part)[32], int lane, int warp) {
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
  const int rest = o > 0 ? 2 * o - 1 : 0;   // lanes that differ only in these bits hold partials of the same value
  for (; o > 0; o >>= 1) x[0] += __shfl_xor_sync(0xffffffffu, x[0], o);
  if ((lane & rest) == 0) part[warp][idx] = x[0];
}

template <int N, typename F>
__device__ __forceinline__ void replay_all_steps(F&& f) {
  if constexpr (N > 0) {
    replay_all_steps<N - 1>(f);
    f(std::integral_constant<int, N - 1>{});
  }
}

// Block one-sided Jacobi.  Per round a CTA holds NR rows (two blocks) in registers and
//   1. forms their NR x NR Gram matrix (per-thread partials, deterministic block reduction),
//   2. runs the plane rotations of the round on the small Gram in shared memory (exactly the rotations
//      Hestenes' method would apply to the long rows, since a rotation's angle only needs the three
//      inner products), recording (s, tau) of every rotation,
//   3. replays the recorded rotations on the register-resident rows and on the rows of U^T, with no
//      synchronisation in between, in Rutishauser's form (norm-preserving in fp32 even when cos rounds to 1,
//      which would otherwise inflate every row by +t^2 per rotation).
template <int NR, int SLOTS>
__global__ void __launch_bounds__(SVD_THREADS) svd_jacobi_kernel(SvdArgs a) {
  cg::grid_group grid = cg::this_grid();
  constexpr int BLK = NR / 2, NP = NR / 2, NGRP = NR * NR / 16, LD = NR + 1;
  extern __shared__ __align__(16) float uts[];   // [NR][R]
  __shared__ float part[2][SVD_WARPS][32];
  __shared__ float Gs[2][NR][LD];
  __shared__ float rot_c[NR], rot_b[NR];      // per ROW of the current step: self coefficient c, partner coefficient (+-s)
  __shared__ float rec_s[NR - 1][NP], rec_tau[NR - 1][NP];   // recorded rotations of the round: (s, tau) per (step, pair)
  __shared__ int partner[NR];
  __shared__ int srow[NR];
  __shared__ int smrow[NR];   // padded row index in M of each local row
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
  const bool timer = blockIdx.x == 0 && tid == 0;
  long long tacc[6] = {0, 0, 0, 0, 0, 0}, tprev = timer ? clock64() : 0;
  auto tick = [&](int k) { if (timer) { long long now = clock64(); tacc[k] += now - tprev; tprev = now; } };

  for (int sweep = 0; sweep < a.max_sweeps; ++sweep) {
    for (int r = 0; r < rounds; ++r) {
      if (active) {
        int bi, bj;
        if (q == 0) { bi = nbe - 1; bj = r; }
        else { bi = (r + q) % (nbe - 1); bj = (r + nbe - 1 - q) % (nbe - 1); }
        if (tid < NR) {
          int blk = tid < BLK ? bi : bj;
          int p = blk * BLK + (tid % BLK);
          const bool ok = blk < nb && p < Ra;
          srow[tid] = ok ? p : -1;
          smrow[tid] = ok ? __ldcg(&a.rowmap[p]) : 0;
        }
        if (tid == 0) { sflag = 0; smax = 0.f; }
        __syncthreads();
        // ---- load the NR rows: thread owns columns tid + SVD_THREADS * slot
#pragma unroll
        for (int x = 0; x < NR; ++x) {
          const int p = srow[x];
          const float* src = p >= 0 ? a.M + (size_t)smrow[x] * Cc : nullptr;
#pragma unroll
          for (int sl = 0; sl < SLOTS; ++sl) {
            int c = tid + SVD_THREADS * sl;
            v[x][sl] = (src && c < Cc) ? __ldcg(src + c) : 0.f;
          }
          if (tid < R) uts[x * R + tid] = p >= 0 ? __ldcg(a.Ut + (size_t)p * R + tid) : 0.f;
        }
        tick(0);
        // ---- Gram matrix (upper triangle), 16 entries (16/NR rows) per block reduction
#pragma unroll
        for (int grp = 0; grp < NGRP; ++grp) {
          float pg[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int x = grp * (16 / NR) + k / NR, y = k % NR;
            float s = 0.f;
            if (y >= x) {
#pragma unroll
              for (int sl = 0; sl < SLOTS; ++sl) s = fmaf(v[x][sl], v[y][sl], s);
            }
            pg[k] = s;
          }
          warp_partials<16>(pg, part[grp & 1], lane, warp);
          __syncthreads();
          if (tid < 16 * SVD_WARPS) {
            const int j = tid / SVD_WARPS, w = tid % SVD_WARPS;   // 16 values x 16 warps
            float s = part[grp & 1][w][j];
#pragma unroll
            for (int o = SVD_WARPS / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const int x = grp * (16 / NR) + j / NR, y = j % NR;
            if (w == 0 && y >= x) { Gs[0][x][y] = s; Gs[0][y][x] = s; }
          }
        }
        __syncthreads();
        tick(1);
        // ---- rotations of this round on the small Gram
        const int nsteps = r == 0 ? NR - 1 : BLK;    // first round of a sweep: all pairs of the NR rows; else cross-block pairs
        int cur = 0;
        for (int step = 0; step < nsteps; ++step) {
          if (tid < NP) {
            int x, y;
            if (r == 0) { x = rr_first(NR, step, tid); y = rr_second(NR, step, tid); }
            else { x = tid; y = BLK + (tid + step) % BLK; }
            const float al = Gs[cur][x][x], be = Gs[cur][y][y], ga = Gs[cur][x][y];
            float c = 1.f, s = 0.f, tau = 0.f;
            const float big = fmaxf(al, be), small = fminf(al, be);
            if (srow[x] >= 0 && srow[y] >= 0 && small > 1e-30f && small > 1e-24f * big) {
              const float relg = fabsf(ga) / (sqrtf(al) * sqrtf(be));
              if (relg > a.tol) {
                const float zeta = (be - al) / (2.f * ga);
                const float t = copysignf(1.f, zeta) / (fabsf(zeta) + sqrtf(fmaf(zeta, zeta, 1.f)));
                c = 1.f / sqrtf(fmaf(t, t, 1.f));
                s = c * t;
                tau = s / (1.f + c);
                sflag = 1;
                atomicMax(reinterpret_cast<int*>(&smax), __float_as_int(relg));   // relg >= 0: int order == float order
              }
            }
            // x' = c x - s y ; y' = s x + c y
            partner[x] = y; rot_c[x] = c; rot_b[x] = -s;
            partner[y] = x; rot_c[y] = c; rot_b[y] = s;
            rec_s[step][tid] = s;
            rec_tau[step][tid] = tau;
          }
          __syncthreads();
          if (tid < NR * NR) {
            const int i = tid / NR, j = tid % NR, pi = partner[i], pj = partner[j];
            const float ai = rot_c[i], bi_ = rot_b[i], aj = rot_c[j], bj_ = rot_b[j];
            float g = ai * aj * Gs[cur][i][j];
            g = fmaf(ai * bj_, Gs[cur][i][pj], g);
            g = fmaf(bi_ * aj, Gs[cur][pi][j], g);
            g = fmaf(bi_ * bj_, Gs[cur][pi][pj], g);
            Gs[cur ^ 1][i][j] = g;
          }
          __syncthreads();
          cur ^= 1;
        }
        tick(2);
        // ---- replay the recorded rotations on the register-resident rows and on the rows of U^T, in Rutishauser's
        //      form x' = x - s (y + tau x), y' = y + s (x - tau y)  (norm-preserving even when cos rounds to 1)
        if (sflag) {
          auto replay = [&](auto kind_tag, auto step_tag) {
            constexpr int ALL = decltype(kind_tag)::value, S = decltype(step_tag)::value;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
              const int x = ALL ? rr_first(NR, S, p) : p;
              const int y = ALL ? rr_second(NR, S, p) : BLK + (p + S) % BLK;
              const float sn = rec_s[S][p], tau = rec_tau[S][p];
              if (sn != 0.f) {
#pragma unroll
                for (int sl = 0; sl < SLOTS; ++sl) {
                  const float u = v[x][sl], w = v[y][sl];
                  v[x][sl] = fmaf(-sn, fmaf(tau, u, w), u);
                  v[y][sl] = fmaf(sn, fmaf(-tau, w, u), w);
                }
                if (tid < R) {
                  const float u = uts[x * R + tid], w = uts[y * R + tid];
                  uts[x * R + tid] = fmaf(-sn, fmaf(tau, u, w), u);
                  uts[y * R + tid] = fmaf(sn, fmaf(-tau, w, u), w);
                }
              }
            }
          };
          if (r == 0) {
            replay_all_steps<NR - 1>([&](auto st) { replay(std::integral_constant<int, 1>{}, st); });
          } else {
            replay_all_steps<BLK>([&](auto st) { replay(std::integral_constant<int, 0>{}, st); });
          }
          tick(3);
          // ---- write back
#pragma unroll
          for (int x = 0; x < NR; ++x) {
            const int p = srow[x];
            if (p >= 0) {
              float* dst = a.M + (size_t)smrow[x] * Cc;
#pragma unroll
              for (int sl = 0; sl < SLOTS; ++sl) {
                int c = tid + SVD_THREADS * sl;
                if (c < Cc) __stcg(dst + c, v[x][sl]);
              }
              if (tid < R) __stcg(a.Ut + (size_t)p * R + tid, uts[x * R + tid]);
            }
          }
          if (tid == 0) {
            atomicAdd(&a.iscal[TRMPS_I_SVD_ROT + sweep], 1);
            atomicMax(&a.iscal[TRMPS_I_SVD_ROT + 32 + (sweep & 31)], __float_as_int(smax));
          }
        }
        tick(4);
      }
      grid.sync();
      tick(5);
    }
    sweeps_used = sweep + 1;
    const int nrot = __ldcg(&a.iscal[TRMPS_I_SVD_ROT + sweep]);
    const float worst = __int_as_float(__ldcg(&a.iscal[TRMPS_I_SVD_ROT + 32 + (sweep & 31)]));
    // quadratic convergence: a sweep whose largest relative off-diagonal was eps leaves O(eps^2) behind
    if (nrot == 0 || worst < 2e-4f) break;
  }
  if (blockIdx.x == 0 && tid == 0) {
    a.iscal[TRMPS_I_SVD_SWEEPS] = sweeps_used;
    a.iscal[TRMPS_I_N_SWEEPS] += sweeps_used;
    if (a.dbg) for (int k = 0; k < 6; ++k) a.dbg[k] = (float)tacc[k];
  }
}

// =====================================================================================================
// Cluster version (the one normally used): CL_SIZE CTAs form a thread-block cluster per row-block pair; each
// CTA owns a contiguous slice of the columns (and of the columns of U^T), so a round's load, Gram partials,
// rotation replay and store are spread over 8x more SMs.  The 16 x 16 Gram partials of the 8 CTAs are summed
// through distributed shared memory in a fixed order, so all CTAs of a cluster (and all ranks of a data-parallel
// job) hold bit-identical Gram matrices and take identical rotation decisions.  Rounds are chained by
// per-(block, slice) release/acquire flags in global memory instead of a grid barrier: CTA k of a cluster only
// waits for the CTAs with the same rank k that held its two blocks in the previous round.  One grid barrier per
// sweep decides convergence.
constexpr int CL_THREADS = 256, CL_WARPS = CL_THREADS / 32, CL_SIZE = 8, CL_NR = 16, CL_NG = CL_NR * (CL_NR + 1) / 2;
constexpr int CL_MAXU = 64;   // columns of U^T per CTA (R <= 512)

struct SvdClArgs {
  SvdArgs a; int cols_per_cta; int ucols_per_cta; int32_t* flags;   // flags: [blocks][CL_SIZE] rounds completed
  int32_t* wtop;     // [32] per-sweep worst |cos| (float bits) over pairs that involve a row that can survive the truncation
};

__device__ __forceinline__ int ld_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

// partner / role of `row` in step `step`; all = 1: round-robin over the 16 rows, else cross-block pairs
__device__ __forceinline__ void pair_of_row(int all, int step, int row, int& x, int& y, int& pidx) {
  constexpr int NR = CL_NR, BLK = CL_NR / 2;
  if (!all) {
    if (row < BLK) { x = row; y = BLK + (row + step) % BLK; pidx = row; }
    else { y = row; x = (row - BLK - step % BLK + BLK) % BLK; pidx = x; }
    return;
  }
  if (row == NR - 1) { x = NR - 1; y = step; pidx = 0; return; }
  if (row == step) { x = NR - 1; y = step; pidx = 0; return; }
  const int j = (row - step + (NR - 1)) % (NR - 1);     // 1 .. NR-2
  if (j <= (NR - 2) / 2) { pidx = j; x = row; y = (step + (NR - 1) - j) % (NR - 1); }
  else { pidx = (NR - 1) - j; y = row; x = (step + pidx) % (NR - 1); }
}



template <int SLOTS>
__global__ void __launch_bounds__(CL_THREADS) svd_jacobi_cluster_kernel(SvdClArgs ca) {
  cg::grid_group grid = cg::this_grid();
  cg::cluster_group cluster = cg::this_cluster();
  constexpr int NR = CL_NR, BLK = NR / 2, NP = NR / 2, NG = CL_NG, LD = NR + 1, NCH = (NG + 31) / 32;
  const SvdArgs& a = ca.a;
  __shared__ float uts[NR][CL_MAXU];
  __shared__ float part[CL_WARPS][NCH * 32];
  __shared__ float cta_g[2][NG];
  __shared__ float Gs[2][NR][LD];
  __shared__ float rec_s[NR - 1][NP], rec_tau[NR - 1][NP];
  __shared__ int srow[NR], smrow[NR];
  __shared__ int sflag;
  __shared__ float smax, smax_top, theta_s;
  __shared__ int nbot_s;
  __shared__ float cn_s[CL_SIZE * CL_MAXU];      // current squared norms by rank (R <= 512)
  __shared__ unsigned char tab_partner[2][NR - 1][NR], tab_first[2][NR - 1][NR], tab_pidx[2][NR - 1][NR];
  __shared__ unsigned char tab_x[2][NR - 1][NP], tab_y[2][NR - 1][NP];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < 2 * (NR - 1) * NR; e += CL_THREADS) {
    const int row = e % NR, step = (e / NR) % (NR - 1), all_ = e / (NR * (NR - 1));
    int x, y, pidx;
    pair_of_row(all_, step, row, x, y, pidx);
    tab_partner[all_][step][row] = (unsigned char)(row == x ? y : x);
    tab_first[all_][step][row] = (unsigned char)(row == x);
    tab_pidx[all_][step][row] = (unsigned char)pidx;
    if (row == x) { tab_x[all_][step][pidx] = (unsigned char)x; tab_y[all_][step][pidx] = (unsigned char)y; }
  }
  __syncthreads();
  const int crank = (int)cluster.block_rank();
  const int q = blockIdx.x / CL_SIZE;
  const int Cc = a.Cc, R = a.R;
  const int c0 = crank * ca.cols_per_cta, c1 = min(Cc, c0 + ca.cols_per_cta);
  const int u0 = crank * ca.ucols_per_cta, ucnt = max(0, min(R, u0 + ca.ucols_per_cta) - u0);
  const int dn = *a.d_near, Ra = 2 * dn;
  const int nb = (Ra + BLK - 1) / BLK;
  int nbe = nb + (nb & 1);
  if (nbe < 2) nbe = 2;
  const int rounds = nbe - 1;
  const bool active = q < nbe / 2;
  int sweeps_used = 0, ground = 0;     // ground: global round counter (flags hold the number of completed rounds)
  float v[NR][SLOTS];
  const bool timer = blockIdx.x == 0 && tid == 0;
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = timer ? clock64() : 0;
  auto tick = [&](int k) { if (timer) { long long now = clock64(); tacc[k] += now - tprev; tprev = now; } };

  int32_t* wtop = ca.wtop;
  for (int sweep = 0; sweep < a.max_sweeps; ++sweep) {
    // Rows whose squared norm is below 1/16 of the keep-th largest cannot survive the truncation (m <= max_size):
    // pairs of two such rows do not count towards convergence.  Safety: with n_bot such rows and all their mutual
    // cosines below w, their span has no singular value above sigma_keep * sqrt((1 + n_bot w) / 16), so the sweep
    // loop may only stop on the kept rows' criterion while n_bot * w < 7 (nothing there can belong to the top m).
    if (tid == 0) { theta_s = 0.f; nbot_s = 0; }
    if (sweep > 0 && a.keep < Ra) {
      const float* nsrc = (sweep & 1) ? a.norms2 : a.norms;          // written during the previous sweep
      for (int i = tid; i < Ra; i += CL_THREADS) cn_s[i] = __ldcg(&nsrc[i]);
      __syncthreads();
      for (int i = tid; i < Ra; i += CL_THREADS) {
        const float c = cn_s[i];
        int cnt = 0;
        for (int o = 0; o < Ra; ++o) cnt += (cn_s[o] > c) || (cn_s[o] == c && o < i);
        if (cnt == a.keep - 1) theta_s = 0.0625f * c;
      }
      __syncthreads();
      int mine = 0;
      for (int i = tid; i < Ra; i += CL_THREADS) mine += cn_s[i] < theta_s;
      if (mine) atomicAdd(&nbot_s, mine);
    }
    __syncthreads();
    for (int r = 0; r < rounds; ++r, ++ground) {
      if (!active) continue;
      int bi, bj;
      if (q == 0) { bi = nbe - 1; bj = r; }
      else { bi = (r + q) % (nbe - 1); bj = (r + nbe - 1 - q) % (nbe - 1); }
      if (tid < NR) {
        const int blk = tid < BLK ? bi : bj;
        const int p = blk * BLK + (tid % BLK);
        const bool ok = blk < nb && p < Ra;
        srow[tid] = ok ? p : -1;
        smrow[tid] = ok ? __ldcg(&a.rowmap[p]) : 0;
      }
      if (tid == 0) {
        sflag = 0; smax = 0.f; smax_top = 0.f;
        // wait until the previous holders of my two blocks (same column slice) have stored them
        if (bi < nb) while (ld_volatile(&ca.flags[bi * CL_SIZE + crank]) < ground) { }
        if (bj < nb) while (ld_volatile(&ca.flags[bj * CL_SIZE + crank]) < ground) { }
        __threadfence();
      }
      __syncthreads();
      tick(6);
      // ---- load my column slice of the NR rows (registers) and of their U^T rows (shared).  All global loads are
      //      issued before the first dependent instruction so that their latencies overlap.
      {
        float utmp[NR];
#pragma unroll
        for (int x = 0; x < NR; ++x) {
          const int p = srow[x];
          utmp[x] = (p >= 0 && tid < ucnt) ? __ldcg(a.Ut + (size_t)p * R + u0 + tid) : 0.f;
        }
#pragma unroll
        for (int x = 0; x < NR; ++x) {
          const int p = srow[x];
          const float* src = a.M + (size_t)smrow[x] * Cc;
#pragma unroll
          for (int sl = 0; sl < SLOTS; ++sl) {
            const int c = c0 + tid + CL_THREADS * sl;
            v[x][sl] = (p >= 0 && c < c1) ? __ldcg(src + c) : 0.f;
          }
        }
        if (tid < ucnt) {
#pragma unroll
          for (int x = 0; x < NR; ++x) uts[x][tid] = utmp[x];
        }
      }
      tick(0);
      // ---- Gram partials of my slice (upper triangle, row-major), reduced over the CTA
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        float pg[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) pg[k] = 0.f;
#pragma unroll
        for (int x = 0; x < NR; ++x) {
#pragma unroll
          for (int y = x; y < NR; ++y) {
            const int idx = x * NR - x * (x - 1) / 2 + (y - x);
            if (idx / 32 == ch) {
              float s = 0.f;
#pragma unroll
              for (int sl = 0; sl < SLOTS; ++sl) s = fmaf(v[x][sl], v[y][sl], s);
              pg[idx % 32] = s;
            }
          }
        }
        warp_partials<32>(pg, reinterpret_cast<float(*)[32]>(&part[0][0]), lane, warp * NCH + ch);
      }
      __syncthreads();
      const int par = ground & 1;
      if (tid < NG) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < CL_WARPS; ++w) s += part[w][tid];
        cta_g[par][tid] = s;
      }
      cluster.sync();
      if (tid < NG) {
        float s = 0.f;
#pragma unroll
        for (int rk = 0; rk < CL_SIZE; ++rk) s += cluster.map_shared_rank(&cta_g[par][0], rk)[tid];
        int x = 0, rem = tid;
        while (rem >= NR - x) { rem -= NR - x; ++x; }
        const int y = x + rem;
        Gs[0][x][y] = s;
        Gs[0][y][x] = s;
        if (x == y && crank == 0 && srow[x] >= 0) (((sweep + 1) & 1) ? a.norms2 : a.norms)[srow[x]] = s;      // current squared norm, by rank
      }
      __syncthreads();
      tick(1);
      // ---- rotations of this round on the small Gram.  Every warp recomputes the 8 rotations of the step (lane & 7),
      //      thread (i, j) fetches the two it needs by shuffle and updates G[i][j]: one barrier per step.
      const int all = r == 0 ? 1 : 0;
      const int nsteps = all ? NR - 1 : BLK;
      int cur = 0;
      {
        const int i = tid / NR, j = tid % NR, pl = lane & 7;
        float worst_rel = 0.f, worst_top = 0.f;
        const float theta = theta_s;
        // pairing data of the step (constant tables): fetched one step ahead, off the critical path
        int px = tab_x[all][0][pl], py = tab_y[all][0][pl];
        int qi = tab_pidx[all][0][i], qj = tab_pidx[all][0][j];
        int opi = tab_partner[all][0][i], opj = tab_partner[all][0][j];
        int fi = tab_first[all][0][i], fj = tab_first[all][0][j];
        for (int step = 0; step < nsteps; ++step) {
          const float al = Gs[cur][px][px], be = Gs[cur][py][py], ga = Gs[cur][px][py];
          const bool okp = srow[px] >= 0 && srow[py] >= 0;
          const float gij = Gs[cur][i][j], gipj = Gs[cur][i][opj], gpij = Gs[cur][opi][j], gpipj = Gs[cur][opi][opj];
          const int cqi = qi, cqj = qj, cfi = fi, cfj = fj;
          if (step + 1 < nsteps) {
            px = tab_x[all][step + 1][pl]; py = tab_y[all][step + 1][pl];
            qi = tab_pidx[all][step + 1][i]; qj = tab_pidx[all][step + 1][j];
            opi = tab_partner[all][step + 1][i]; opj = tab_partner[all][step + 1][j];
            fi = tab_first[all][step + 1][i]; fj = tab_first[all][step + 1][j];
          }
          float c, sn, tau, relg;
          bool rot = rotation_params(al, be, ga, a.tol, c, sn, tau, relg);
          if (!okp) { rot = false; c = 1.f; sn = 0.f; tau = 0.f; relg = 0.f; }
          const float ci = __shfl_sync(0xffffffffu, c, cqi), si = __shfl_sync(0xffffffffu, sn, cqi);
          const float cj = __shfl_sync(0xffffffffu, c, cqj), sj = __shfl_sync(0xffffffffu, sn, cqj);
          const float bi_ = cfi ? -si : si, bj_ = cfj ? -sj : sj;
          float g = ci * cj * gij;
          g = fmaf(ci * bj_, gipj, g);
          g = fmaf(bi_ * cj, gpij, g);
          g = fmaf(bi_ * bj_, gpipj, g);
          Gs[cur ^ 1][i][j] = g;
          if (tid < NP) {                 // warp 0, lanes 0..7: one lane per pair records the rotation
            rec_s[step][tid] = sn;
            rec_tau[step][tid] = tau;
            if (rot) sflag = 1;
            worst_rel = fmaxf(worst_rel, relg);
            if (fmaxf(al, be) >= theta) worst_top = fmaxf(worst_top, relg);
          }
          __syncthreads();
          cur ^= 1;
        }
        if (tid < NP && worst_rel > 0.f) atomicMax(reinterpret_cast<int*>(&smax), __float_as_int(worst_rel));
        if (tid < NP && worst_top > 0.f) atomicMax(reinterpret_cast<int*>(&smax_top), __float_as_int(worst_top));
        __syncthreads();
      }
      tick(2);
      // ---- replay the recorded rotations on my slice of the rows and of U^T (Rutishauser form), then store
      if (sflag) {
        auto replay = [&](auto kind_tag, auto step_tag) {
          constexpr int ALL = decltype(kind_tag)::value, S = decltype(step_tag)::value;
          float sn[NP], ta[NP];
#pragma unroll
          for (int p = 0; p < NP; ++p) { sn[p] = rec_s[S][p]; ta[p] = rec_tau[S][p]; }
#pragma unroll
          for (int p = 0; p < NP; ++p) {        // s = tau = 0 is an exact identity: no branch
            const int x = ALL ? rr_first(NR, S, p) : p;
            const int y = ALL ? rr_second(NR, S, p) : BLK + (p + S) % BLK;
#pragma unroll
            for (int sl = 0; sl < SLOTS; ++sl) {
              const float u = v[x][sl], w = v[y][sl];
              v[x][sl] = fmaf(-sn[p], fmaf(ta[p], u, w), u);
              v[y][sl] = fmaf(sn[p], fmaf(-ta[p], w, u), w);
            }
            if (tid < ucnt) {
              const float u = uts[x][tid], w = uts[y][tid];
              uts[x][tid] = fmaf(-sn[p], fmaf(ta[p], u, w), u);
              uts[y][tid] = fmaf(sn[p], fmaf(-ta[p], w, u), w);
            }
          }
        };
        if (all) replay_all_steps<NR - 1>([&](auto st) { replay(std::integral_constant<int, 1>{}, st); });
        else replay_all_steps<BLK>([&](auto st) { replay(std::integral_constant<int, 0>{}, st); });
        tick(3);
#pragma unroll
        for (int x = 0; x < NR; ++x) {
          const int p = srow[x];
          if (p >= 0) {
            float* dst = a.M + (size_t)smrow[x] * Cc;
#pragma unroll
            for (int sl = 0; sl < SLOTS; ++sl) {
              const int c = c0 + tid + CL_THREADS * sl;
              if (c < c1) __stcg(dst + c, v[x][sl]);
            }
            if (tid < ucnt) __stcg(a.Ut + (size_t)p * R + u0 + tid, uts[x][tid]);
          }
        }
        if (tid == 0 && crank == 0) {
          atomicAdd(&a.iscal[TRMPS_I_SVD_ROT + sweep], 1);
          atomicMax(&a.iscal[TRMPS_I_SVD_ROT + 32 + (sweep & 31)], __float_as_int(smax));
          atomicMax(&wtop[sweep & 31], __float_as_int(smax_top));
        }
      }
      // ---- publish: my slice of both blocks is up to date for round ground+1 (the barrier orders every thread's
      //      stores before thread 0's fence, which is cumulative)
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        if (bi < nb) *reinterpret_cast<volatile int*>(&ca.flags[bi * CL_SIZE + crank]) = ground + 1;
        if (bj < nb) *reinterpret_cast<volatile int*>(&ca.flags[bj * CL_SIZE + crank]) = ground + 1;
      }
      tick(4);
    }
    grid.sync();
    tick(5);
    sweeps_used = sweep + 1;
    const int nrot = __ldcg(&a.iscal[TRMPS_I_SVD_ROT + sweep]);
    const float worst = __int_as_float(__ldcg(&a.iscal[TRMPS_I_SVD_ROT + 32 + (sweep & 31)]));
    const float worst_kept = __int_as_float(__ldcg(&wtop[sweep & 31]));
    // quadratic convergence: what is left after this sweep is O(worst^2) <= 1e-6
    if (nrot == 0 || worst < 1e-3f || (worst_kept < 1e-3f && worst * (float)nbot_s < 7.f)) break;
  }
  if (blockIdx.x == 0 && tid == 0) {
    a.iscal[TRMPS_I_SVD_SWEEPS] = sweeps_used;
    a.iscal[TRMPS_I_N_SWEEPS] += sweeps_used;
    if (a.dbg) for (int k = 0; k < 8; ++k) a.dbg[k] = (float)tacc[k];
  }
}

__global__ void svd_zero_flags_kernel(int32_t* flags, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = 0;
}

template <int SLOTS>
static cudaError_t launch_jacobi_cluster(SvdClArgs& ca, int nclusters, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nclusters * CL_SIZE);
  cfg.blockDim = dim3(CL_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = CL_SIZE; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeCooperative;
  attrs[1].val.cooperative = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, svd_jacobi_cluster_kernel<SLOTS>, ca);
}

// sigma, ordering, m, dims update                                     optimizer.py:294-302
struct SvdFinArgs {
  const float* M; int R, Cc, Ds; const int32_t* d_near; const int32_t* d_far; int L; const int32_t* rowmap;
  float* sv; int32_t* order; int32_t* iscal; int32_t* dim_out;
  float min_sv; int max_size, min_size;
  int ncol_blocks;      // blocks of d_far columns that can be non-zero: 2L, or L for a bond with a unit far site
};

// squared norms of the converged rows, one warp per Jacobi rank (the whole grid reads the 2.4 MB bond; a single CTA
// took 48 us for it)
__global__ void svd_rank_norms_kernel(const float* __restrict__ M, const int32_t* __restrict__ rowmap, const int32_t* __restrict__ d_near,
                                      int Cc, float* __restrict__ out) {
  const int Ra = 2 * *d_near, lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int p = gw; p < Ra; p += nw) {
    const float4* row = reinterpret_cast<const float4*>(M + (size_t)rowmap[p] * Cc);
    float s = 0.f;
    for (int c4 = lane; c4 < Cc / 4; c4 += 32) {
      float4 v = row[c4];
      s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[p] = s;
  }
}

__global__ void __launch_bounds__(1024) svd_finalize_kernel(SvdFinArgs a) {
  extern __shared__ float sig[];   // [R]
  __shared__ int s_count;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int dn = *a.d_near, Ra = 2 * dn;
  if (tid == 0) s_count = 0;
  for (int p = tid; p < Ra; p += 1024) sig[p] = sqrtf(a.sv[p]);      // squared row norms by Jacobi rank (svd_rank_norms_kernel)
  __syncthreads();
  const int kmax = min(Ra, a.ncol_blocks * (*a.d_far));
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

int svd_split(const trmps_sweep* s, const trmps_layout& lay, const BondGeom& g, const trmps_opt* opt, cudaStream_t st,
              float* special_out, bool single) {
  const int Ds = s->Ds, L = s->L, R = 2 * Ds, Cc = 2 * L * Ds;
  if (opt->split_impl != 1 && split_factor_supported(R)) return split_factor(s, lay, g, opt, st, special_out, single);
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
  a.norms2 = s->work + lay.red + 1024;      // (red: 8 debug floats, the hand-over flags from +32, then this)
  a.rowmap = reinterpret_cast<int32_t*>(s->work + lay.sv + 2 * R);
  a.tol = 5.9604645e-8f * sqrtf((float)Cc);
  a.dbg = s->work + lay.red;
  a.keep = opt->max_size < Ds ? opt->max_size : Ds;
  a.max_sweeps = opt->max_svd_sweeps > 0 ? (opt->max_svd_sweeps < 40 ? opt->max_svd_sweeps : 40) : 40;
  svd_norms_kernel<<<(R * 32 + 255) / 256, 256, 0, st>>>(a);
  TRMPS_LAUNCH_OK("svd_norms_kernel");
  svd_init_kernel<<<32, 256, 0, st>>>(a);
  TRMPS_LAUNCH_OK("svd_init_kernel");
  svd_init_ut_kernel<<<(R + 255) / 256, 256, 0, st>>>(a);
  TRMPS_LAUNCH_OK("svd_init_ut_kernel");

  int rc = TRMPS_OK;
  bool done = false;
  // preferred: cluster kernel (8 CTAs per block pair)
  {
    SvdClArgs ca;
    ca.a = a;
    ca.cols_per_cta = (Cc + CL_SIZE - 1) / CL_SIZE;
    ca.ucols_per_cta = (R + CL_SIZE - 1) / CL_SIZE;
    ca.flags = reinterpret_cast<int32_t*>(s->work + lay.red + 32);
    const int slots = (ca.cols_per_cta + CL_THREADS - 1) / CL_THREADS;
    const int nbk = (R + CL_NR / 2 - 1) / (CL_NR / 2);
    ca.wtop = ca.flags + (nbk + 1) * CL_SIZE;
    const int ncl = ((nbk + (nbk & 1)) / 2) > 1 ? (nbk + (nbk & 1)) / 2 : 1;
    if (slots <= 4 && ca.ucols_per_cta <= CL_MAXU && (nbk + 1) * CL_SIZE <= 2048) {
      svd_zero_flags_kernel<<<((nbk + 1) * CL_SIZE + 40 + 255) / 256, 256, 0, st>>>(ca.flags, (nbk + 1) * CL_SIZE + 40);
      TRMPS_LAUNCH_OK("svd_zero_flags_kernel");
      // a refused cluster + cooperative launch is an error, not a reason to switch kernels: the single-CTA kernel
      // rotates in another order, so a rank that fell back alone would leave the replicated MPS of a data-parallel job
      cudaError_t e = slots <= 1 ? launch_jacobi_cluster<1>(ca, ncl, st)
                    : slots <= 2 ? launch_jacobi_cluster<2>(ca, ncl, st)
                    : slots <= 3 ? launch_jacobi_cluster<3>(ca, ncl, st)
                                 : launch_jacobi_cluster<4>(ca, ncl, st);
      if (e != cudaSuccess) return cuda_fail(e, "svd_jacobi_cluster_kernel (cluster + cooperative launch)");
      count_launch();
      done = true;
    }
  }
  if (!done) {
    const int nb = (R + blk - 1) / blk;
    const int ncta = ((nb + (nb & 1)) / 2) > 1 ? (nb + (nb & 1)) / 2 : 1;
    const int slots = (Cc + SVD_THREADS - 1) / SVD_THREADS;
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
  }

  SvdFinArgs f;
  f.M = a.M; f.R = R; f.Cc = Cc; f.Ds = Ds; f.d_near = g.d_near; f.d_far = g.d_far; f.L = L; f.rowmap = a.rowmap;
  f.sv = s->work + lay.sv;
  f.order = reinterpret_cast<int32_t*>(s->work + lay.sv + R);
  f.iscal = s->iwork;
  f.dim_out = s->dims + (g.dir > 0 ? g.near_site + 1 : g.near_site);
  f.min_sv = opt->min_singular_value;
  f.max_size = opt->max_size < Ds ? opt->max_size : Ds;
  f.min_size = opt->min_size > 0 ? opt->min_size : 3;
  f.ncol_blocks = single ? L : 2 * L;
  svd_rank_norms_kernel<<<(R * 32 + 255) / 256, 256, 0, st>>>(a.M, a.rowmap, a.d_near, Cc, f.sv);
  TRMPS_LAUNCH_OK("svd_rank_norms_kernel");
  svd_finalize_kernel<<<1, 1024, R * sizeof(float), st>>>(f);
  TRMPS_LAUNCH_OK("svd_finalize_kernel");

  SvdScatterArgs c;
  c.M = a.M; c.Ut = a.Ut; c.R = R; c.Cc = Cc; c.Ds = Ds; c.L = L; c.dir = g.dir; c.d_near = g.d_near;
  c.order = f.order; c.rowmap = a.rowmap; c.iscal = s->iwork;
  c.node = s->nodes + (size_t)g.near_site * 2 * Ds * Ds;
  c.special = special_out ? special_out : s->special;
  svd_scatter_kernel<<<296, 256, 0, st>>>(c);
  TRMPS_LAUNCH_OK("svd_scatter_kernel");
  return TRMPS_OK;
}

}  // namespace trmps
