// Bond-step contractions (north_star piece 3 and the scalar logic around it).
//
// The reference materialises the projected input C[t,m,n,i,k] = C1[t,i] C2[t,k] phi1[t,m] phi2[t,n]
// (optimizer.py:197-220; 13.8 GB at B=60k, D=120) and contracts it with the bond four times per update.
// Here C is never formed.  With P[t,(m,i)] = phi1[t,m] C1[t,i] and Q[t,(n,k)] = phi2[t,n] C2[t,k]:
//   forward   f[t,l]      = sum_c ( sum_a P[t,a] M[a,(l,n,k)] ) Q[t,(n,k)]        (optimizer.py:222-227)
//   gradient  G[a,(l,n,k)] = sum_t P[t,a] * (y-h)[t,l] Q[t,(n,k)]                  (optimizer.py:249)
// on the bond matrix M (R x Cc, the SVD layout).  The Armijo trials (baseoptimizer.py:299-322) use
// f(bond + lr*delta) = f(bond) + lr*f(delta), so one extra forward serves all 19 step sizes.
#include "common.cuh"

namespace trmps {

// =================================================================================================
// forward: CTA = TM images x (a range of 128-column tiles of M).  P tile resident in smem, M streamed
// with cp.async, epilogue weights the accumulators with Q and reduces over k into f[t,l].
struct FwdArgs {
  const float* e_near; const float* e_far; const float* phi_near; const float* phi_far; const float* M;
  float* fpart; int64_t B; int Ds, L, nd, Cc; const int32_t* d_near; int ntiles, tiles_per_split;
};

template <int RQ>
__global__ void __launch_bounds__(TRMPS_THREADS) forward_kernel(FwdArgs a) {
  constexpr int TM = 64 * RQ, TN = 128, KC = 16;
  extern __shared__ __align__(16) float smem[];
  const int Ds = a.Ds, L = a.L, nd = a.nd, Cc = a.Cc, LP = L + 1;
  float* Ps = smem;                       // [2*Ds][TM]
  float* Bs = Ps + (size_t)2 * Ds * TM;   // [2][KC][TN]
  float* E2s = Bs + 2 * KC * TN;          // [TM][Ds]
  float* ph = E2s + (size_t)TM * Ds;      // [4][TM]
  float* fsm = ph + 4 * TM;               // [TM][LP]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t t0 = (int64_t)blockIdx.x * TM;
  const int dn8 = align_up(__ldg(a.d_near), 8);
  const int nchunks = dn8 / 8;
  const int tile0 = blockIdx.y * a.tiles_per_split;
  int my_tiles = a.ntiles - tile0;
  if (my_tiles > a.tiles_per_split) my_tiles = a.tiles_per_split;
  if (my_tiles < 0) my_tiles = 0;

  for (int r = tid; r < TM; r += TRMPS_THREADS) {
    int64_t t = t0 + r;
    float2 p1 = make_float2(0.f, 0.f), p2 = make_float2(1.f, 1.f);
    if (t < a.B) {
      p1 = __ldg(reinterpret_cast<const float2*>(a.phi_near) + t);
      if (nd == 2) p2 = __ldg(reinterpret_cast<const float2*>(a.phi_far) + t);
    }
    ph[r] = p1.x; ph[TM + r] = p1.y; ph[2 * TM + r] = p2.x; ph[3 * TM + r] = p2.y;
  }
  for (int e = tid; e < TM * LP; e += TRMPS_THREADS) fsm[e] = 0.f;
  for (int e = tid; e < TM * (Ds / 4); e += TRMPS_THREADS) {
    int t = e / (Ds / 4), c4 = e % (Ds / 4);
    int64_t tg = t0 + t;
    float4 v = tg < a.B ? __ldg(reinterpret_cast<const float4*>(a.e_far + tg * Ds) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(E2s + (size_t)t * Ds + 4 * c4) = v;
  }
  __syncthreads();
  for (int e = tid; e < TM * (dn8 / 4); e += TRMPS_THREADS) {
    int t = e % TM, i4 = e / TM;
    int64_t tg = t0 + t;
    float4 v = tg < a.B ? __ldg(reinterpret_cast<const float4*>(a.e_near + tg * Ds) + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
    float p0 = ph[t], p1 = ph[TM + t];
    float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      Ps[(size_t)(4 * i4 + q) * TM + t] = p0 * vv[q];
      Ps[(size_t)(Ds + 4 * i4 + q) * TM + t] = p1 * vv[q];
    }
  }

  const int total_it = my_tiles * nchunks;
  auto issue = [&](int it) {
    int ch = it % nchunks, nt = tile0 + it / nchunks;
    float* base = Bs + (it & 1) * KC * TN;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      int e = tid + TRMPS_THREADS * r;
      int row = e >> 5, c4 = e & 31;
      int m = row >> 3, q = row & 7;
      int c = nt * TN + 4 * c4;
      bool valid = c < Cc;
      const float* src = valid ? a.M + (size_t)(m * Ds + ch * 8 + q) * Cc + c : a.M;
      cp_async16(base + row * TN + 4 * c4, src, valid);
    }
    cp_async_commit();
  };

  float acc[4 * RQ][8];
  if (total_it > 0) issue(0);
  __syncthreads();   // Ps complete
  for (int it = 0; it < total_it; ++it) {
    if (it + 1 < total_it) {
      issue(it + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int ch = it % nchunks, nt = tile0 + it / nchunks;
    if (ch == 0) {
#pragma unroll
      for (int i = 0; i < 4 * RQ; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    }
    const float* bs = Bs + (it & 1) * KC * TN;
    mma_rows<RQ, 8>(Ps + (size_t)(ch * 8) * TM, TM, bs, TN, ty * 4, TM / 2 + ty * 4, tx * 4, TN / 2 + tx * 4, acc);
    mma_rows<RQ, 8>(Ps + (size_t)(Ds + ch * 8) * TM, TM, bs + 8 * TN, TN, ty * 4, TM / 2 + ty * 4, tx * 4, TN / 2 + tx * 4, acc);
    __syncthreads();
    if (ch == nchunks - 1) {
      // epilogue: weight with Q[t,(n,k)] = phi_far[t,n] * e_far[t,k], reduce over the tile's columns into f[t,l]
#pragma unroll
      for (int qd = 0; qd < 2; ++qd) {
        const int cb = nt * TN + qd * 64;
        if (cb >= Cc) continue;
        const int c = cb + tx * 4;
        const bool valid = c < Cc;
        const int ln = valid ? c / Ds : 0;
        const int k = valid ? c - ln * Ds : 0;
        const int l = ln / nd, n = ln - l * nd;
        float part[4 * RQ];
#pragma unroll
        for (int r = 0; r < 4 * RQ; ++r) {
          int rr = r < 4 ? ty * 4 + r : TM / 2 + ty * 4 + (r - 4);
          float4 w = *reinterpret_cast<const float4*>(E2s + (size_t)rr * Ds + k);
          float pf = ph[(2 + n) * TM + rr];
          float d = acc[r][4 * qd] * w.x;
          d = fmaf(acc[r][4 * qd + 1], w.y, d);
          d = fmaf(acc[r][4 * qd + 2], w.z, d);
          d = fmaf(acc[r][4 * qd + 3], w.w, d);
          part[r] = valid ? pf * d : 0.f;
        }
        const int clast = (cb + 63 < Cc ? cb + 63 : Cc - 1);
        const int l_lo = cb / (nd * Ds), l_hi = clast / (nd * Ds);
        for (int ll = l_lo; ll <= l_hi; ++ll) {
#pragma unroll
          for (int r = 0; r < 4 * RQ; ++r) {
            float v = (valid && l == ll) ? part[r] : 0.f;
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            if (tx == 0) {
              int rr = r < 4 ? ty * 4 + r : TM / 2 + ty * 4 + (r - 4);
              fsm[rr * LP + ll] += v;
            }
          }
        }
      }
    }
  }
  __syncthreads();
  float* out = a.fpart + (size_t)blockIdx.y * a.B * L;
  for (int e = tid; e < TM * L; e += TRMPS_THREADS) {
    int t = e / L, l = e - t * L;
    int64_t tg = t0 + t;
    if (tg < a.B) out[tg * L + l] = fsm[t * LP + l];
  }
}

static size_t forward_smem(int RQ, int Ds, int L) {
  int TM = 64 * RQ;
  return ((size_t)2 * Ds * TM + 2 * 16 * 128 + (size_t)TM * Ds + 4 * TM + (size_t)TM * (L + 1)) * sizeof(float);
}

int forward_rq(int Ds, int L) {
  const size_t lim = 227 * 1024;
  if (forward_smem(2, Ds, L) <= lim) return 2;
  if (forward_smem(1, Ds, L) <= lim) return 1;
  return 0;
}

int launch_forward(const float* e_near, const float* e_far, const float* phi_near, const float* phi_far, const float* M,
                   float* fpart, int nsplit, int64_t B, int Ds, int L, int nd, const int32_t* d_near, cudaStream_t st) {
  if (B <= 0) return TRMPS_OK;
  FwdArgs a;
  a.e_near = e_near; a.e_far = e_far; a.phi_near = phi_near; a.phi_far = phi_far; a.M = M; a.fpart = fpart;
  a.B = B; a.Ds = Ds; a.L = L; a.nd = nd; a.Cc = nd * L * Ds; a.d_near = d_near;
  a.ntiles = (a.Cc + 127) / 128;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > a.ntiles) nsplit = a.ntiles;
  a.tiles_per_split = (a.ntiles + nsplit - 1) / nsplit;
  int rq = forward_rq(Ds, L);
  if (rq == 0) {
    set_error("forward: bond stride Ds=%d does not fit shared memory (max ~272)", Ds);
    return TRMPS_E_ARG;
  }
  size_t smem = forward_smem(rq, Ds, L);
  dim3 grid((unsigned)ceil_div64(B, 64 * rq), (unsigned)nsplit);
  if (rq == 2) {
    TRMPS_CUDA_OK(cudaFuncSetAttribute(forward_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    forward_kernel<2><<<grid, TRMPS_THREADS, smem, st>>>(a);
  } else {
    TRMPS_CUDA_OK(cudaFuncSetAttribute(forward_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    forward_kernel<1><<<grid, TRMPS_THREADS, smem, st>>>(a);
  }
  TRMPS_LAUNCH_OK("forward_kernel");
  // splits beyond the launched ones must read as zero: the caller sums `nsplit` partials it asked for
  return TRMPS_OK;
}

// =================================================================================================
// gradient: 128x128 output tile of G = P^T W per CTA, split-K over the images.  Both operands are formed
// on the fly while staging global -> registers -> shared (never materialised in HBM).
struct GradArgs {
  const float* e_near; const float* e_far; const float* phi_near; const float* sres; float* gpart;
  int64_t B, imgs_per_split; int Ds, L, R, Cc;
};

template <bool SQ>
__global__ void __launch_bounds__(TRMPS_THREADS, 2) grad_kernel(GradArgs g) {
  constexpr int KC = 16, T = 128;
  __shared__ __align__(16) float As[2][KC][T];
  __shared__ __align__(16) float Bs[2][KC][T];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int c0 = blockIdx.x * T, a0 = blockIdx.y * T;
  const int Ds = g.Ds, L2 = 2 * g.L;
  const int64_t tbeg = (int64_t)blockIdx.z * g.imgs_per_split;
  int64_t tend = tbeg + g.imgs_per_split;
  if (tend > g.B) tend = g.B;
  const int nchunks = tend > tbeg ? (int)((tend - tbeg + KC - 1) / KC) : 0;

  // this thread stages one quad of A columns and one quad of B columns for image rows kk and kk+8
  const int q4 = tid & 31, kk0 = tid >> 5;
  const int arow = a0 + 4 * q4, ccol = c0 + 4 * q4;
  const bool va = arow < g.R, vc = ccol < g.Cc;
  const int am = va ? arow / Ds : 0, ai = va ? arow - am * Ds : 0;
  const int cln = vc ? ccol / Ds : 0, ck = vc ? ccol - cln * Ds : 0;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto load = [&](int ch) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      int64_t t = tbeg + (int64_t)ch * KC + kk0 + 8 * r;
      ra[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      rb[r] = ra[r];
      if (t < tend) {
        if (va) {
          float4 v = __ldg(reinterpret_cast<const float4*>(g.e_near + t * Ds + ai));
          float p = __ldg(g.phi_near + 2 * t + am);
          v.x *= p; v.y *= p; v.z *= p; v.w *= p;
          if (SQ) { v.x *= v.x; v.y *= v.y; v.z *= v.z; v.w *= v.w; }
          ra[r] = v;
        }
        if (vc) {
          float4 v = __ldg(reinterpret_cast<const float4*>(g.e_far + t * Ds + ck));
          float s = __ldg(g.sres + t * L2 + cln);
          if (SQ) { v.x *= v.x; v.y *= v.y; v.z *= v.z; v.w *= v.w; }
          v.x *= s; v.y *= s; v.z *= s; v.w *= s;
          rb[r] = v;
        }
      }
    }
  };
  auto store = [&](int buf) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      *reinterpret_cast<float4*>(&As[buf][kk0 + 8 * r][4 * q4]) = ra[r];
      *reinterpret_cast<float4*>(&Bs[buf][kk0 + 8 * r][4 * q4]) = rb[r];
    }
  };

  if (nchunks > 0) {
    load(0);
    store(0);
  }
  __syncthreads();
  for (int ch = 0; ch < nchunks; ++ch) {
    const int buf = ch & 1;
    const bool more = ch + 1 < nchunks;
    if (more) load(ch + 1);
    mma_rows<2, KC>(&As[buf][0][0], T, &Bs[buf][0][0], T, ty * 4, 64 + ty * 4, tx * 4, 64 + tx * 4, acc);
    if (more) store(buf ^ 1);
    __syncthreads();
  }

  float* out = g.gpart + (size_t)blockIdx.z * g.R * g.Cc;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    int row = a0 + (r < 4 ? ty * 4 + r : 64 + ty * 4 + (r - 4));
    if (row >= g.R) continue;
#pragma unroll
    for (int qd = 0; qd < 2; ++qd) {
      int col = c0 + qd * 64 + tx * 4;
      if (col < g.Cc)
        *reinterpret_cast<float4*>(out + (size_t)row * g.Cc + col) =
            make_float4(acc[r][4 * qd], acc[r][4 * qd + 1], acc[r][4 * qd + 2], acc[r][4 * qd + 3]);
    }
  }
}

int launch_grad(const float* e_near, const float* e_far, const float* phi_near, const float* sres, float* gpart,
                int gsplit, int64_t B, int Ds, int L, bool squared, cudaStream_t st) {
  GradArgs g;
  g.e_near = e_near; g.e_far = e_far; g.phi_near = phi_near; g.sres = sres; g.gpart = gpart;
  g.B = B; g.Ds = Ds; g.L = L; g.R = 2 * Ds; g.Cc = 2 * L * Ds;
  int64_t per = ceil_div64(B, gsplit);
  g.imgs_per_split = ceil_div64(per, 16) * 16;
  dim3 grid((unsigned)((g.Cc + 127) / 128), (unsigned)((g.R + 127) / 128), (unsigned)gsplit);
  if (squared) grad_kernel<true><<<grid, TRMPS_THREADS, 0, st>>>(g);
  else grad_kernel<false><<<grid, TRMPS_THREADS, 0, st>>>(g);
  TRMPS_LAUNCH_OK("grad_kernel");
  return TRMPS_OK;
}

// Fixed-order sum of the per-CTA partials red[r][c] (r < nrows, c < ncols <= 32) by ONE CTA of >= 256 threads: 8 row groups
// per pass over 32 groups (rows r = g, g + 32, ..), then the groups in order -- the summation order of sum_red_kernel, so a
// kernel whose last CTA to arrive finishes its own reduction gives the same bits as the two-launch form.
__device__ __forceinline__ void fixed_order_reduce(const float* __restrict__ red, int nrows, int ncols, float* part /* [32][33] shared */) {
  const int c = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int g = w; g < 32; g += nw) {
    float s = 0.f;
    if (c < ncols)
      for (int r = g; r < nrows; r += 32) s += __ldcg(&red[r * 32 + c]);
    part[g * 33 + c] = s;
  }
  __syncthreads();
  if (w == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += part[k * 33 + c];
    part[c] = t;                 // row 0 is only read by this lane's own column
  }
  __syncthreads();
}
// arrival of this CTA at a self-resetting counter; true in every thread of the last CTA to arrive
__device__ __forceinline__ bool last_cta_arrives(int32_t* counter, int* s_flag) {
  if (threadIdx.x < 32) __threadfence();          // the per-CTA partials are written by threads of warp 0 only: the other warps' (large) output
  __syncthreads();                                //   stores need not drain before this CTA announces itself
  if (threadIdx.x == 0) {
    const int old = atomicAdd(counter, 1);
    const int last = old == (int)gridDim.x - 1;
    if (last) *counter = 0;
    *s_flag = last;
  }
  __syncthreads();
  const bool last = *s_flag != 0;
  if (last) __threadfence();
  return last;
}

// out[i] = sum_s part[s][i]   (fixed order -> bit-reproducible)
__global__ void sum_parts_kernel(const float4* __restrict__ part, int nparts, int64_t n4, float4* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 s = part[i];
    for (int p = 1; p < nparts; ++p) {
      float4 v = part[(int64_t)p * n4 + i];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    out[i] = s;
  }
}

int launch_sum_parts(const float* part, int nparts, int64_t n, float* out, cudaStream_t st) {
  int64_t n4 = n / 4;
  int blocks = (int)(ceil_div64(n4, 256) < 1184 ? ceil_div64(n4, 256) : 1184);
  sum_parts_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(part), nparts, n4, reinterpret_cast<float4*>(out));
  TRMPS_LAUNCH_OK("sum_parts_kernel");
  return TRMPS_OK;
}

// G (= sum of the split-K parts when `parts` is given, in a fixed order) -= 2*reg*M ; delta = G / H (Hessian on) ;
// <G, delta> / B_total -> scal[S_GDC] by the last CTA to arrive (fixed-order sum of the per-CTA partials) ;
// scal[S_COST0] = cost sum * cost_scale ; largest |delta| -> *maxbits (integer atomicMax: order independent), the exponent
// the next forward contraction scales delta by                                                     optimizer.py:249-253
struct GradFinishArgs {
  float* G; const float* M; const float* H; float* delta; const float* parts; int nparts; int64_t n;
  float reg, hess_add, gdc_scale, cost_scale; const float* cost_sum; float* red; float* scal; int32_t* arrive; int32_t* maxbits;
};
__global__ void __launch_bounds__(256) grad_finish_kernel(GradFinishArgs a) {
  __shared__ float sred[32];
  __shared__ float part[32 * 33];
  __shared__ int s_flag;
  float local = 0.f, amax = 0.f;
  const float two_reg = 2.f * a.reg;
  const int64_t n4 = a.n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 g;
    if (a.parts) {
      g = __ldcs(reinterpret_cast<const float4*>(a.parts) + i);
      for (int p = 1; p < a.nparts; ++p) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(a.parts) + (int64_t)p * n4 + i);
        g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
      }
    } else {
      g = reinterpret_cast<const float4*>(a.G)[i];
    }
    const float4 m = reinterpret_cast<const float4*>(a.M)[i];
    g.x -= two_reg * m.x; g.y -= two_reg * m.y; g.z -= two_reg * m.z; g.w -= two_reg * m.w;
    reinterpret_cast<float4*>(a.G)[i] = g;
    float4 d = g;
    if (a.H) {
      const float4 h = reinterpret_cast<const float4*>(a.H)[i];
      d.x = g.x / (h.x + a.hess_add); d.y = g.y / (h.y + a.hess_add); d.z = g.z / (h.z + a.hess_add); d.w = g.w / (h.w + a.hess_add);
      reinterpret_cast<float4*>(a.delta)[i] = d;
    }
    local = fmaf(g.x, d.x, local); local = fmaf(g.y, d.y, local); local = fmaf(g.z, d.z, local); local = fmaf(g.w, d.w, local);
    amax = fmaxf(amax, fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fmaxf(fabsf(d.z), fabsf(d.w))));
  }
  if (a.maxbits) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0 && amax > 0.f) atomicMax(a.maxbits, __float_as_int(amax));
  }
  const float s = block_sum(local, sred);
  if (threadIdx.x == 0) __stcg(&a.red[blockIdx.x * 32], s);
  if (!last_cta_arrives(a.arrive, &s_flag)) return;
  fixed_order_reduce(a.red, (int)gridDim.x, 1, part);
  if (threadIdx.x == 0) {
    a.scal[TRMPS_S_GDC] = part[0] * a.gdc_scale;
    a.scal[TRMPS_S_COST0] = __ldcg(a.cost_sum) * a.cost_scale;
  }
}

// dst[c] = scale * sum_r red[r][c]  for c < ncols <= 32: 32 row groups sum their rows (r = g, g + 32, ..) in order, then the groups
// are added in order -- a fixed summation order (every rank gets the same bits) with 32 loads in flight per column
// instead of one warp walking all rows
__global__ void __launch_bounds__(1024) sum_red_kernel(const float* __restrict__ red, int nrows, int ncols, float scale, float* __restrict__ dst) {
  __shared__ float part[32][33];
  const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  float s = 0.f;
  if (c < ncols)
    for (int r = g; r < nrows; r += 32) s += red[r * 32 + c];
  part[g][c] = s;
  __syncthreads();
  if (g == 0 && c < ncols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += part[k][c];
    dst[c] = t * scale;
  }
}

// =================================================================================================
// per-image loss / residual                                           optimizer.py:222-227, baseoptimizer.py:347-352

// one image's residual row (L floats, L even: float2 stores) and residual x feature row (2L floats: float4 stores)
__device__ __forceinline__ void store_resid_rows(float* __restrict__ resid, float* __restrict__ sres, const float (&r)[TRMPS_LMAX], float2 pf, int L) {
#pragma unroll
  for (int l = 0; l < TRMPS_LMAX; l += 2) {
    if (l < L) {
      *reinterpret_cast<float2*>(resid + l) = make_float2(r[l], r[l + 1]);
      *reinterpret_cast<float4*>(sres + 2 * l) = make_float4(r[l] * pf.x, r[l] * pf.y, r[l + 1] * pf.x, r[l + 1] * pf.y);
    }
  }
}

__global__ void __launch_bounds__(256) loss_kernel(LossArgs a) {
  __shared__ float sred[32];
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int L = a.L;
  float cost_t = 0.f;
  const bool even = (L & 1) == 0;          // rows of L floats are then 8-byte aligned: float2 / float4 accesses (L = 10: 5 instead of 10 per row)
  if (t < a.B) {
    float f[TRMPS_LMAX], y[TRMPS_LMAX];
    if (even) {
#pragma unroll
      for (int l = 0; l < TRMPS_LMAX; l += 2) {
        f[l] = f[l + 1] = 0.f; y[l] = y[l + 1] = 0.f;
        if (l < L) {
          float2 s;
          if (a.from_parts) {
            s = make_float2(0.f, 0.f);
            for (int p = 0; p < a.nsplit; ++p) {
              const float2 v = *reinterpret_cast<const float2*>(a.fpart + ((int64_t)p * a.B + t) * L + l);
              s.x += v.x; s.y += v.y;
            }
            *reinterpret_cast<float2*>(a.f0 + t * L + l) = s;
          } else {
            s = *reinterpret_cast<const float2*>(a.f0 + t * L + l);
          }
          f[l] = s.x; f[l + 1] = s.y;
          const float2 yy = __ldg(reinterpret_cast<const float2*>(a.labels + t * L + l));
          y[l] = yy.x; y[l + 1] = yy.y;
        }
      }
    } else {
#pragma unroll
    for (int l = 0; l < TRMPS_LMAX; ++l) {
      f[l] = 0.f; y[l] = 0.f;
      if (l < L) {
        if (a.from_parts) {
          float s = 0.f;
          for (int p = 0; p < a.nsplit; ++p) s += a.fpart[((int64_t)p * a.B + t) * L + l];
          f[l] = s;
          a.f0[t * L + l] = s;
        } else {
          f[l] = a.f0[t * L + l];
        }
        y[l] = __ldg(a.labels + t * L + l);
      }
    }
    }
    float2 pf = __ldg(reinterpret_cast<const float2*>(a.phi_far) + t);
    if (a.loss_kind == TRMPS_LOSS_SOFTMAX_XENT) {
      float mx = -INFINITY;
#pragma unroll
      for (int l = 0; l < TRMPS_LMAX; ++l) if (l < L) mx = fmaxf(mx, f[l]);
      float e[TRMPS_LMAX], se = 0.f;
#pragma unroll
      for (int l = 0; l < TRMPS_LMAX; ++l) { e[l] = l < L ? expf(f[l] - mx) : 0.f; se += e[l]; }
      const float lse = logf(se), inv = 1.f / se;
      float rr[TRMPS_LMAX];
#pragma unroll
      for (int l = 0; l < TRMPS_LMAX; ++l) {
        rr[l] = 0.f;
        if (l < L) {
          float h = e[l] * inv;
          float r = y[l] - h;
          rr[l] = r;
          cost_t += y[l] * (lse - (f[l] - mx));
          if (!even) {
            a.resid[t * L + l] = r;
            a.sres[t * 2 * L + 2 * l] = r * pf.x;
            a.sres[t * 2 * L + 2 * l + 1] = r * pf.y;
          }
          if (a.hres) {
            float hh = h * (1.f - h);
            a.hres[t * 2 * L + 2 * l] = hh * pf.x * pf.x;
            a.hres[t * 2 * L + 2 * l + 1] = hh * pf.y * pf.y;
          }
        }
      }
      if (even) store_resid_rows(a.resid + t * L, a.sres + t * 2 * L, rr, pf, L);
    } else {
      float rr[TRMPS_LMAX];
#pragma unroll
      for (int l = 0; l < TRMPS_LMAX; ++l) {
        rr[l] = 0.f;
        if (l < L) {
          float r = y[l] - f[l];
          rr[l] = r;
          cost_t += r * r;
          if (!even) {
            a.resid[t * L + l] = r;
            a.sres[t * 2 * L + 2 * l] = r * pf.x;
            a.sres[t * 2 * L + 2 * l + 1] = r * pf.y;
          }
        }
      }
      if (even) store_resid_rows(a.resid + t * L, a.sres + t * 2 * L, rr, pf, L);
    }
  }
  float s = block_sum(cost_t, sred);
  if (threadIdx.x == 0) __stcg(&a.red[blockIdx.x * 32], s);
  // the last CTA to arrive adds the per-CTA sums in a fixed order (one launch instead of loss + sum_red)
  __shared__ float part[32 * 33];
  __shared__ int s_flag;
  if (!last_cta_arrives(a.arrive, &s_flag)) return;
  fixed_order_reduce(a.red, (int)gridDim.x, 1, part);
  if (threadIdx.x == 0) {
    *a.cost_sum = part[0];
    if (a.clear_word) *a.clear_word = 0;          // the atomicMax word grad_finish_kernel fills
  }
}

// single-thread decision: Armijo acceptance + accept-only-if-cost-dropped     baseoptimizer.py:299-322, optimizer.py:257-262
__device__ __forceinline__ void decide_update(float* __restrict__ scal, const float* __restrict__ trial, int32_t* __restrict__ iscal,
                                              float lr, float armijo_coeff) {
  const float cost0 = scal[TRMPS_S_COST0], gdc = scal[TRMPS_S_GDC];
  float lr_c = lr, lr_star = 0.f, cost1 = cost0;
  int trials = 0;
  bool found = false;
  for (int c = 0; c < 19; ++c) {
    ++trials;
    float cost_c = trial[c];
    float target = cost0 - armijo_coeff * lr_c * gdc;
    if (!(cost_c > target)) { found = true; lr_star = lr_c; cost1 = cost_c; break; }
    lr_c *= 0.5f;
  }
  bool accept = found && (cost1 < cost0);
  if (!found) iscal[TRMPS_I_N_EXHAUST] += 1;
  if (!accept) { lr_star = 0.f; iscal[TRMPS_I_N_REVERT] += 1; }
  iscal[TRMPS_I_N_UPDATES] += 1;
  iscal[TRMPS_I_ACCEPT] = accept ? 1 : 0;
  iscal[TRMPS_I_TRIALS] = trials;
  scal[TRMPS_S_LR_STAR] = lr_star;
  scal[TRMPS_S_COST1] = accept ? cost1 : cost0;
  scal[TRMPS_S_LR] = lr;
}

// Armijo trial costs for lr, lr/2, ... (19 values) from f0 and fd = f(delta)      baseoptimizer.py:299-322

__global__ void __launch_bounds__(256) armijo_kernel(ArmijoArgs a) {
  __shared__ float swarp[8][20];
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int L = a.L;
  float f[TRMPS_LMAX], d[TRMPS_LMAX], y[TRMPS_LMAX];
  float ysum = 0.f;
  if ((L & 1) == 0) {
    // rows of L floats are 8-byte aligned: float2 accesses
#pragma unroll
    for (int l = 0; l < TRMPS_LMAX; l += 2) {
      f[l] = f[l + 1] = 0.f; d[l] = d[l + 1] = 0.f; y[l] = y[l + 1] = 0.f;
      if (t < a.B && l < L) {
        const float2 ff = *reinterpret_cast<const float2*>(a.f0 + t * L + l);
        float2 s = make_float2(0.f, 0.f);
        for (int p = 0; p < a.nsplit; ++p) {
          const float2 v = *reinterpret_cast<const float2*>(a.fdpart + ((int64_t)p * a.B + t) * L + l);
          s.x += v.x; s.y += v.y;
        }
        *reinterpret_cast<float2*>(a.fd + t * L + l) = s;
        const float2 yy = __ldg(reinterpret_cast<const float2*>(a.labels + t * L + l));
        f[l] = ff.x; f[l + 1] = ff.y; d[l] = s.x; d[l + 1] = s.y; y[l] = yy.x; y[l + 1] = yy.y;
        ysum += yy.x; ysum += yy.y;
      }
    }
  } else {
#pragma unroll
  for (int l = 0; l < TRMPS_LMAX; ++l) {
    f[l] = 0.f; d[l] = 0.f; y[l] = 0.f;
    if (t < a.B && l < L) {
      f[l] = a.f0[t * L + l];
      float s = 0.f;
      for (int p = 0; p < a.nsplit; ++p) s += a.fdpart[((int64_t)p * a.B + t) * L + l];
      d[l] = s;
      a.fd[t * L + l] = s;
      y[l] = __ldg(a.labels + t * L + l);
      ysum += y[l];
    }
  }
  }
  float lr = a.lr;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int c = 0; c < 19; ++c) {
    float cost_t = 0.f;
    if (t < a.B) {
      float z[TRMPS_LMAX];
      if (a.loss_kind == TRMPS_LOSS_SOFTMAX_XENT) {
        float mx = -INFINITY;
#pragma unroll
        for (int l = 0; l < TRMPS_LMAX; ++l) { z[l] = fmaf(lr, d[l], f[l]); if (l < L) mx = fmaxf(mx, z[l]); }
        float se = 0.f, yz = 0.f;
#pragma unroll
        for (int l = 0; l < TRMPS_LMAX; ++l) if (l < L) { se += expf(z[l] - mx); yz = fmaf(y[l], z[l] - mx, yz); }
        cost_t = ysum * logf(se) - yz;
      } else {
#pragma unroll
        for (int l = 0; l < TRMPS_LMAX; ++l) if (l < L) { float r = fmaf(lr, d[l], f[l]) - y[l]; cost_t = fmaf(r, r, cost_t); }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cost_t += __shfl_xor_sync(0xffffffffu, cost_t, o);
    if (lane == 0) swarp[w][c] = cost_t;
    lr *= 0.5f;
  }
  __syncthreads();
  if (threadIdx.x < 19) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += swarp[i][threadIdx.x];
    __stcg(&a.red[blockIdx.x * 32 + threadIdx.x], s);
  }
  // the last CTA to arrive: the 19 trial sums in a fixed order and, on a single rank, the decision itself
  __shared__ float part[32 * 33];
  __shared__ int s_flag;
  if (!last_cta_arrives(a.arrive, &s_flag)) return;
  fixed_order_reduce(a.red, (int)gridDim.x, 19, part);
  if (threadIdx.x < 19) a.trial[threadIdx.x] = part[threadIdx.x] * a.trial_scale;
  if (a.decide) {
    __syncthreads();
    if (threadIdx.x == 0) decide_update(a.scal, a.trial, a.iscal, a.lr, a.armijo_coeff);
  }
}

__global__ void decide_kernel(float* __restrict__ scal, const float* __restrict__ trial, int32_t* __restrict__ iscal,
                              float lr, float armijo_coeff) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  decide_update(scal, trial, iscal, lr, armijo_coeff);
}

// M += lr* delta ; f0 += lr* fd
__global__ void apply_update_kernel(float* __restrict__ M, const float* __restrict__ delta, int64_t n,
                                    float* __restrict__ f0, const float* __restrict__ fd, int64_t nf,
                                    const float* __restrict__ scal) {
  const float lr = scal[TRMPS_S_LR_STAR];
  if (lr == 0.f) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    M[i] = fmaf(lr, delta[i], M[i]);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += (int64_t)gridDim.x * blockDim.x)
    f0[i] = fmaf(lr, fd[i], f0[i]);
}

// =================================================================================================
// merge the special node with its neighbour into the bond matrix           optimizer.py:169 (right), :115 (left)
//   right: M[(m,i),(l,n,k)] = sum_j special[l,m,i,j] * far[n,j,k]
//   left : M[(m,i),(l,n,k)] = sum_j far[n,k,j]      * special[l,m,j,i]
__global__ void __launch_bounds__(256) merge_kernel(const float* __restrict__ sp, const float* __restrict__ far, float* __restrict__ M,
                                                    int Ds, int L, int dir, const int32_t* __restrict__ d_mid) {
  constexpr int T = 64, KC = 16;
  __shared__ float As[KC][T + 4];
  __shared__ float Bs[KC][T + 4];
  const int ln = blockIdx.z, l = ln >> 1, n = ln & 1;
  const int k0 = blockIdx.x * T, r0 = blockIdx.y * T;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int mid8 = align_up(__ldg(d_mid), 8);
  const int R = 2 * Ds;
  const int64_t Cc = (int64_t)2 * L * Ds;
  float acc[4][4] = {};
  // this thread's four elements of each tile (fixed row / column, four k-indices): addresses computed once, the loads of the
  // NEXT k-chunk are in flight while the current one is multiplied (the kernel was a chain of exposed L2 round trips)
  int a_row[4], a_jj[4], b_col[4], b_jj[4];
  const float* a_ptr[4];
  const float* b_ptr[4];
  int64_t a_step, b_step;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int e = tid + u * 256;
    if (dir > 0) { a_row[u] = e / KC; a_jj[u] = e % KC; b_col[u] = e % T; b_jj[u] = e / T; }
    else { a_row[u] = e % T; a_jj[u] = e / T; b_col[u] = e / KC; b_jj[u] = e % KC; }
    const int rg = r0 + a_row[u], kg = k0 + b_col[u];
    a_ptr[u] = nullptr; b_ptr[u] = nullptr;
    if (rg < R) {
      const int m = rg / Ds, i = rg - m * Ds;
      a_ptr[u] = dir > 0 ? sp + ((size_t)(l * 2 + m) * Ds + i) * Ds + a_jj[u] : sp + ((size_t)(l * 2 + m) * Ds + a_jj[u]) * Ds + i;
    }
    if (kg < Ds) b_ptr[u] = dir > 0 ? far + ((size_t)n * Ds + b_jj[u]) * Ds + kg : far + ((size_t)n * Ds + kg) * Ds + b_jj[u];
  }
  a_step = dir > 0 ? KC : (int64_t)KC * Ds;          // advance of the k-index by one chunk
  b_step = dir > 0 ? (int64_t)KC * Ds : KC;
  float ra[4], rb[4];
  auto fetch = [&](int j0) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ra[u] = (a_ptr[u] && j0 + a_jj[u] < Ds) ? __ldg(a_ptr[u] + (int64_t)(j0 / KC) * a_step) : 0.f;
      rb[u] = (b_ptr[u] && j0 + b_jj[u] < Ds) ? __ldg(b_ptr[u] + (int64_t)(j0 / KC) * b_step) : 0.f;
    }
  };
  if (mid8 > 0) fetch(0);
  for (int j0 = 0; j0 < mid8; j0 += KC) {
#pragma unroll
    for (int u = 0; u < 4; ++u) { As[a_jj[u]][a_row[u]] = ra[u]; Bs[b_jj[u]][b_col[u]] = rb[u]; }
    __syncthreads();
    if (j0 + KC < mid8) fetch(j0 + KC);
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { av[q] = As[kk][ty * 4 + q]; bv[q] = Bs[kk][tx * 4 + q]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int rg = r0 + ty * 4 + i;
    if (rg >= R) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int kg = k0 + tx * 4 + j;
      if (kg < Ds) M[(size_t)rg * Cc + (size_t)ln * Ds + kg] = acc[i][j];
    }
  }
}

int launch_merge(const float* special, const float* far_node, float* M, int Ds, int L, int dir, const int32_t* d_mid,
                 cudaStream_t st) {
  dim3 grid((unsigned)((Ds + 63) / 64), (unsigned)((2 * Ds + 63) / 64), (unsigned)(2 * L));
  merge_kernel<<<grid, 256, 0, st>>>(special, far_node, M, Ds, L, dir, d_mid);
  TRMPS_LAUNCH_OK("merge_kernel");
  return TRMPS_OK;
}

// launch helpers used by sweep.cu
int launch_loss(const LossArgs& a, cudaStream_t st) {
  if (a.B == 0) return TRMPS_OK;
  loss_kernel<<<(unsigned)ceil_div64(a.B, TRMPS_IMG_BLOCK), TRMPS_IMG_BLOCK, 0, st>>>(a);
  TRMPS_LAUNCH_OK("loss_kernel");
  return TRMPS_OK;
}
int launch_armijo(const ArmijoArgs& a, cudaStream_t st) {
  if (a.B == 0) return TRMPS_OK;
  armijo_kernel<<<(unsigned)ceil_div64(a.B, TRMPS_IMG_BLOCK), TRMPS_IMG_BLOCK, 0, st>>>(a);
  TRMPS_LAUNCH_OK("armijo_kernel");
  return TRMPS_OK;
}
int launch_sum_red(const float* red, int nrows, int ncols, float scale, float* dst, cudaStream_t st) {
  sum_red_kernel<<<1, 1024, 0, st>>>(red, nrows, ncols, scale, dst);
  TRMPS_LAUNCH_OK("sum_red_kernel");
  return TRMPS_OK;
}
int launch_grad_finish(float* G, const float* M, const float* H, float* delta, const float* parts, int nparts, int64_t n, float reg,
                       float hess_add, float gdc_scale, const float* cost_sum, float cost_scale, float* red, float* scal,
                       int32_t* arrive, int32_t* maxbits, int nblocks, cudaStream_t st) {
  GradFinishArgs a;
  a.G = G; a.M = M; a.H = H; a.delta = delta; a.parts = parts; a.nparts = nparts; a.n = n; a.reg = reg; a.hess_add = hess_add;
  a.gdc_scale = gdc_scale; a.cost_scale = cost_scale; a.cost_sum = cost_sum; a.red = red; a.scal = scal; a.arrive = arrive; a.maxbits = maxbits;
  grad_finish_kernel<<<nblocks, 256, 0, st>>>(a);
  TRMPS_LAUNCH_OK("grad_finish_kernel");
  return TRMPS_OK;
}
int launch_decide(float* scal, const float* trial, int32_t* iscal, float lr, float coeff, cudaStream_t st) {
  decide_kernel<<<1, 32, 0, st>>>(scal, trial, iscal, lr, coeff);
  TRMPS_LAUNCH_OK("decide_kernel");
  return TRMPS_OK;
}
int launch_apply_update(float* M, const float* delta, int64_t n, float* f0, const float* fd, int64_t nf, const float* scal,
                        cudaStream_t st) {
  apply_update_kernel<<<296, 256, 0, st>>>(M, delta, n, f0, fd, nf, scal);
  TRMPS_LAUNCH_OK("apply_update_kernel");
  return TRMPS_OK;
}

}  // namespace trmps
