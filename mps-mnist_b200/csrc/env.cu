// Environment chain: fused local feature map + site contraction (north_star pieces 1 and 2).
// Replaces MPS._chain_multiply_r/_l (mps.py:500-519), BaseOptimizer._find_C1/_find_C2
// (baseoptimizer.py:160-201) and the env advance at the end of _update_right/_update_left
// (optimizer.py:187-190, :132-135).  The reference materialises T[t,i,j] = sum_m phi[t,m] A[m,i,j]
// (B x Dl x Dr floats) per step; here one step is a (B x 2Dl) * (2Dl x Dr) product whose left operand
// phi[t,m]*C[t,i] is formed in shared memory and never written to HBM.
#include "common.cuh"

namespace trmps {

// ---------------------------------------------------------------------------------------------------
__global__ void fill_boundary_kernel(float* __restrict__ env, int64_t B, int Ds, int L, int which) {
  // which = 0: start vector [0_L, 1_L] (mps.py:446-454); which = 1: end vector [1_L, 0_L] (mps.py:456-466)
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = B * (Ds / 4);
  if (idx >= total) return;
  int j = (int)(idx % (Ds / 4)) * 4;
  float v[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int jj = j + q;
    v[q] = which == 0 ? ((jj >= L && jj < 2 * L) ? 1.f : 0.f) : (jj < L ? 1.f : 0.f);
  }
  reinterpret_cast<float4*>(env)[idx] = make_float4(v[0], v[1], v[2], v[3]);
}

int launch_fill_boundary(float* env, int64_t B, int Ds, int L, int which, cudaStream_t st) {
  int64_t total = B * (Ds / 4);
  if (total == 0) return TRMPS_OK;
  fill_boundary_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(env, B, Ds, L, which);
  TRMPS_LAUNCH_OK("fill_boundary_kernel");
  return TRMPS_OK;
}

// ---------------------------------------------------------------------------------------------------
// out[t, n] = sum_{m, a} phi_m[t] * in[t, a] * W[(m, a), n]
//   DIR=+1: W[(m,a), n] = node[m][a][n]   (contract the left bond, produce the right bond)
//   DIR=-1: W[(m,a), n] = node[m][n][a]   (contract the right bond, produce the left bond)
// CTA tile: TM images x TN outputs, 8x8 register tile per thread, K chunk = 8 bond indices x 2 features.
template <int TN, int DIR>
__global__ void __launch_bounds__(TRMPS_THREADS)
env_step_kernel(const float* __restrict__ feat_site, int fm, int64_t B, int Ds, const float* __restrict__ in_env,
                const float* __restrict__ node, const int32_t* __restrict__ din_dev, float* __restrict__ out_env) {
  constexpr int TX = TN / 8, TY = TRMPS_THREADS / TX, TM = TY * 8, KC = 16;
  constexpr int A_PER = TM * 2 / TRMPS_THREADS;
  constexpr int B_TOTAL = KC * TN / 4;
  constexpr int B_PER = (B_TOTAL + TRMPS_THREADS - 1) / TRMPS_THREADS;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                  // [2][KC][TM]
  float* Bs = As + 2 * KC * TM;      // [2][KC][TN]
  float* ph = Bs + 2 * KC * TN;      // [2][TM]
  const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
  const int64_t t0 = (int64_t)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  const int din8 = align_up(__ldg(din_dev), 8);
  const int nchunks = din8 / 8;

  for (int r = tid; r < TM; r += TRMPS_THREADS) {
    int64_t t = t0 + r;
    float2 p = t < B ? load_phi(feat_site, fm, t) : make_float2(0.f, 0.f);
    ph[r] = p.x;
    ph[TM + r] = p.y;
  }
  __syncthreads();

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ga[A_PER], gb[B_PER];
  auto load_A = [&](int ch) {
#pragma unroll
    for (int r = 0; r < A_PER; ++r) {
      int e = tid + TRMPS_THREADS * r;
      int t = e % TM, h = e / TM;
      int64_t tg = t0 + t;
      ga[r] = tg < B ? __ldg(reinterpret_cast<const float4*>(in_env + tg * Ds + ch * 8 + 4 * h))
                     : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_A = [&](int buf) {
    float* base = As + buf * KC * TM;
#pragma unroll
    for (int r = 0; r < A_PER; ++r) {
      int e = tid + TRMPS_THREADS * r;
      int t = e % TM, h = e / TM;
      float p0 = ph[t], p1 = ph[TM + t];
      float v[4] = {ga[r].x, ga[r].y, ga[r].z, ga[r].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        base[(4 * h + q) * TM + t] = p0 * v[q];
        base[(8 + 4 * h + q) * TM + t] = p1 * v[q];
      }
    }
  };
  auto load_B = [&](int ch) {
#pragma unroll
    for (int r = 0; r < B_PER; ++r) {
      int e = tid + TRMPS_THREADS * r;
      gb[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < B_TOTAL) {
        if (DIR > 0) {
          int row = e / (TN / 4), c4 = e % (TN / 4);
          int m = row >> 3, q = row & 7, n = n0 + 4 * c4;
          if (n < Ds) gb[r] = __ldg(reinterpret_cast<const float4*>(node + ((size_t)m * Ds + ch * 8 + q) * Ds + n));
        } else {
          int n = e % TN, mh = e / TN;
          int m = mh >> 1, h = mh & 1;
          if (n0 + n < Ds)
            gb[r] = __ldg(reinterpret_cast<const float4*>(node + ((size_t)m * Ds + n0 + n) * Ds + ch * 8 + 4 * h));
        }
      }
    }
  };
  auto store_B = [&](int buf) {
    float* base = Bs + buf * KC * TN;
#pragma unroll
    for (int r = 0; r < B_PER; ++r) {
      int e = tid + TRMPS_THREADS * r;
      if (e < B_TOTAL) {
        if (DIR > 0) {
          int row = e / (TN / 4), c4 = e % (TN / 4);
          *reinterpret_cast<float4*>(base + row * TN + 4 * c4) = gb[r];
        } else {
          int n = e % TN, mh = e / TN;
          int m = mh >> 1, h = mh & 1;
          base[(m * 8 + 4 * h + 0) * TN + n] = gb[r].x;
          base[(m * 8 + 4 * h + 1) * TN + n] = gb[r].y;
          base[(m * 8 + 4 * h + 2) * TN + n] = gb[r].z;
          base[(m * 8 + 4 * h + 3) * TN + n] = gb[r].w;
        }
      }
    }
  };

  if (nchunks > 0) {
    load_A(0);
    load_B(0);
    store_A(0);
    store_B(0);
  }
  __syncthreads();
  for (int ch = 0; ch < nchunks; ++ch) {
    const int buf = ch & 1;
    const bool more = ch + 1 < nchunks;
    if (more) {
      load_A(ch + 1);
      load_B(ch + 1);
    }
    mma_rows<2, KC>(As + buf * KC * TM, TM, Bs + buf * KC * TN, TN, ty * 4, TM / 2 + ty * 4, tx * 4, TN / 2 + tx * 4, acc);
    if (more) {
      store_A(buf ^ 1);
      store_B(buf ^ 1);
    }
    __syncthreads();
  }

#pragma unroll
  for (int r = 0; r < 8; ++r) {
    int rr = r < 4 ? ty * 4 + r : TM / 2 + ty * 4 + (r - 4);
    int64_t t = t0 + rr;
    if (t >= B) continue;
#pragma unroll
    for (int qd = 0; qd < 2; ++qd) {
      int n = n0 + (qd ? TN / 2 : 0) + tx * 4;
      if (n < Ds)
        *reinterpret_cast<float4*>(out_env + t * Ds + n) =
            make_float4(acc[r][4 * qd + 0], acc[r][4 * qd + 1], acc[r][4 * qd + 2], acc[r][4 * qd + 3]);
    }
  }
}

template <int TN, int DIR>
static int launch_env_step_t(const float* feat_site, int fm, int64_t B, int Ds, const float* in_env, const float* node,
                             const int32_t* din_dev, float* out_env, cudaStream_t st) {
  constexpr int TX = TN / 8, TY = TRMPS_THREADS / TX, TM = TY * 8, KC = 16;
  size_t smem = (size_t)(2 * KC * TM + 2 * KC * TN + 2 * TM) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    TRMPS_CUDA_OK(cudaFuncSetAttribute(env_step_kernel<TN, DIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid((unsigned)ceil_div64(B, TM), (unsigned)((Ds + TN - 1) / TN));
  env_step_kernel<TN, DIR><<<grid, TRMPS_THREADS, smem, st>>>(feat_site, fm, B, Ds, in_env, node, din_dev, out_env);
  TRMPS_LAUNCH_OK("env_step_kernel");
  return TRMPS_OK;
}

int launch_env_step(const float* feat_site, int fm, int64_t B, int Ds, int dir, const float* in_env, const float* node,
                    const int32_t* din_dev, float* out_env, cudaStream_t st) {
  if (B <= 0) return TRMPS_OK;
  if (Ds <= 32)
    return dir > 0 ? launch_env_step_t<32, 1>(feat_site, fm, B, Ds, in_env, node, din_dev, out_env, st)
                   : launch_env_step_t<32, -1>(feat_site, fm, B, Ds, in_env, node, din_dev, out_env, st);
  if (Ds <= 64)
    return dir > 0 ? launch_env_step_t<64, 1>(feat_site, fm, B, Ds, in_env, node, din_dev, out_env, st)
                   : launch_env_step_t<64, -1>(feat_site, fm, B, Ds, in_env, node, din_dev, out_env, st);
  return dir > 0 ? launch_env_step_t<128, 1>(feat_site, fm, B, Ds, in_env, node, din_dev, out_env, st)
                 : launch_env_step_t<128, -1>(feat_site, fm, B, Ds, in_env, node, din_dev, out_env, st);
}

// ---------------------------------------------------------------------------------------------------
// datasource-side feature map (MNISTpreprocessing.py:55): pixels -> (count, 2)
__global__ void feature_map_kernel(const float* __restrict__ pix, int64_t count, int fm, float2* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = phi_of_pixel(__ldg(pix + i), fm);
}

// special node (L, 2, Ds, Ds) -> matrix [(m, i)][(l*Ds + k)]  (the forward kernel's nd = 1 layout)
__global__ void special_to_matrix_kernel(const float* __restrict__ sp, float* __restrict__ M, int Ds, int L) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)L * 2 * Ds * Ds;
  if (idx >= total) return;
  int k = (int)(idx % Ds);
  int64_t r = idx / Ds;
  int i = (int)(r % Ds);
  r /= Ds;
  int m = (int)(r % 2);
  int l = (int)(r / 2);
  M[((int64_t)(m * Ds + i)) * ((int64_t)L * Ds) + (int64_t)l * Ds + k] = sp[idx];
}

// A special node as a bond with a unit far site: M[(m,a)][(l,n,b)] = special[l][m][a][b] for n = 0 (dir > 0; for
// dir < 0 the roles of the two bond indices swap: rows (m,b), columns (l,0,a)), zero for n = 1, and
// phi_far = (1, 0) for every image.  This is the layout the two-site contraction kernels consume; it serves the
// special-node contraction of predict (mps.py:552-557) and the single-site optimizer (singlesiteOptimizer.py:49-138).
__global__ void special_to_bond_kernel(const float* __restrict__ sp, float* __restrict__ M, float2* __restrict__ phi_far,
                                       int Ds, int L, int64_t B, int dir) {
  const int64_t total = (int64_t)L * 2 * Ds * Ds;
  const int64_t Cc = (int64_t)2 * L * Ds;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx % Ds);
    int64_t r = idx / Ds;
    const int a = (int)(r % Ds);
    r /= Ds;
    const int m = (int)(r % 2), l = (int)(r / 2);
    const int i = dir > 0 ? a : b, k = dir > 0 ? b : a;          // i: near bond (row), k: far bond (column)
    float* row = M + (int64_t)(m * Ds + i) * Cc + (int64_t)(l * 2) * Ds;
    row[k] = sp[idx];
    row[Ds + k] = 0.f;
  }
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < B; t += (int64_t)gridDim.x * blockDim.x)
    phi_far[t] = make_float2(1.f, 0.f);
}

int launch_special_to_bond(const float* special, float* M, float* phi_far, int Ds, int L, int64_t B, int dir, cudaStream_t st) {
  special_to_bond_kernel<<<296, 256, 0, st>>>(special, M, reinterpret_cast<float2*>(phi_far), Ds, L, B, dir);
  TRMPS_LAUNCH_OK("special_to_bond_kernel");
  return TRMPS_OK;
}

// Single-site move: the S V part of the split (tmp, special layout with n = 0: right move tmp[l][0][q][k], left move
// tmp[l][0][k][q]) is absorbed by the neighbouring node (singlesiteOptimizer.py:124-125 / :76-77):
//   right: special[l][n][q][k'] = sum_k tmp[l][0][q][k] node[n][k][k']     left: special[l][n][k'][q] = sum_k node[n][k'][k] tmp[l][0][k][q]
__global__ void site_absorb_kernel(const float* __restrict__ tmp, const float* __restrict__ node, float* __restrict__ special,
                                   int Ds, int L, int dir) {
  const int64_t total = (int64_t)L * 2 * Ds * Ds;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)(idx % Ds), x = (int)((idx / Ds) % Ds);
    const int n = (int)((idx / ((int64_t)Ds * Ds)) % 2), l = (int)(idx / ((int64_t)2 * Ds * Ds));
    const float* t0 = tmp + (int64_t)(l * 2) * Ds * Ds;
    const float* nd = node + (int64_t)n * Ds * Ds;
    float acc = 0.f;
    if (dir > 0) {
      for (int k = 0; k < Ds; ++k) acc = fmaf(t0[(int64_t)x * Ds + k], nd[(int64_t)k * Ds + y], acc);
    } else {
      for (int k = 0; k < Ds; ++k) acc = fmaf(nd[(int64_t)x * Ds + k], t0[(int64_t)k * Ds + y], acc);
    }
    special[idx] = acc;
  }
}

int launch_site_absorb(const float* tmp, const float* node, float* special, int Ds, int L, int dir, cudaStream_t st) {
  const int64_t total = (int64_t)L * 2 * Ds * Ds;
  site_absorb_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(tmp, node, special, Ds, L, dir);
  TRMPS_LAUNCH_OK("site_absorb_kernel");
  return TRMPS_OK;
}

int launch_special_to_matrix(const float* special, float* M, int Ds, int L, cudaStream_t st) {
  int64_t total = (int64_t)L * 2 * Ds * Ds;
  special_to_matrix_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(special, M, Ds, L);
  TRMPS_LAUNCH_OK("special_to_matrix_kernel");
  return TRMPS_OK;
}

__global__ void sum_round_logits_kernel(const float* __restrict__ fpart, int nsplit, int64_t BL, int round_output,
                                        float* __restrict__ logits) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BL) return;
  float s = 0.f;
  for (int p = 0; p < nsplit; ++p) s += fpart[(int64_t)p * BL + i];
  logits[i] = round_output ? rintf(s) : s;   // tf.round is round-half-to-even (mps.py:558-559)
}

int launch_sum_round_logits(const float* fpart, int nsplit, int64_t B, int L, int round_output, float* logits,
                            cudaStream_t st) {
  int64_t BL = B * L;
  if (BL == 0) return TRMPS_OK;
  sum_round_logits_kernel<<<(unsigned)ceil_div64(BL, 256), 256, 0, st>>>(fpart, nsplit, BL, round_output, logits);
  TRMPS_LAUNCH_OK("sum_round_logits_kernel");
  return TRMPS_OK;
}

__global__ void phi_site_kernel(const float* __restrict__ feat_site, int fm, int64_t B, float2* __restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < B) out[t] = load_phi(feat_site, fm, t);
}

// the feature vectors of both sites of a bond in one launch
__global__ void phi_pair_kernel(const float* __restrict__ feat_a, const float* __restrict__ feat_b, int fm, int64_t B,
                                float2* __restrict__ out_a, float2* __restrict__ out_b) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < B) { out_a[t] = load_phi(feat_a, fm, t); out_b[t] = load_phi(feat_b, fm, t); }
}

int launch_phi_pair(const float* feat_a, const float* feat_b, int fm, int64_t B, float* out_a, float* out_b, cudaStream_t st) {
  if (B == 0) return TRMPS_OK;
  phi_pair_kernel<<<(unsigned)ceil_div64(B, 256), 256, 0, st>>>(feat_a, feat_b, fm, B, reinterpret_cast<float2*>(out_a), reinterpret_cast<float2*>(out_b));
  TRMPS_LAUNCH_OK("phi_pair_kernel");
  return TRMPS_OK;
}

int launch_phi_site(const float* feat_site, int fm, int64_t B, float* out, cudaStream_t st) {
  if (B == 0) return TRMPS_OK;
  phi_site_kernel<<<(unsigned)ceil_div64(B, 256), 256, 0, st>>>(feat_site, fm, B, reinterpret_cast<float2*>(out));
  TRMPS_LAUNCH_OK("phi_site_kernel");
  return TRMPS_OK;
}

}  // namespace trmps

using namespace trmps;

extern "C" int trmps_feature_map(const float* pixels, int64_t count, int32_t feature_map, float* phi_out, void* stream) {
  if (!pixels || !phi_out || count < 0 || feature_map < TRMPS_FM_ONE_SQRT3 || feature_map > TRMPS_FM_COS_SIN) {
    set_error("trmps_feature_map: bad argument");
    return TRMPS_E_ARG;
  }
  if (count == 0) return TRMPS_OK;
  feature_map_kernel<<<(unsigned)ceil_div64(count, 256), 256, 0, (cudaStream_t)stream>>>(
      pixels, count, feature_map, reinterpret_cast<float2*>(phi_out));
  TRMPS_LAUNCH_OK("feature_map_kernel");
  return TRMPS_OK;
}

extern "C" int trmps_env_step(const float* feat_site, int32_t feature_map, int64_t B, int32_t Ds, int32_t dir,
                              const float* in_env, const float* node_slot, const int32_t* din_dev, float* out_env,
                              void* stream) {
  if (!feat_site || !in_env || !node_slot || !din_dev || !out_env || B < 0 || Ds <= 0 || Ds % 8 != 0 ||
      (dir != 1 && dir != -1)) {
    set_error("trmps_env_step: bad argument (Ds must be a positive multiple of 8, dir = +-1)");
    return TRMPS_E_ARG;
  }
  return launch_env_step(feat_site, feature_map, B, Ds, dir, in_env, node_slot, din_dev, out_env, (cudaStream_t)stream);
}

extern "C" int trmps_debug_chain(float* out64) { return trmps::chain_tc_debug(out64); }

extern "C" int64_t trmps_predict_work_floats(int64_t B, int32_t L, int32_t Ds, int32_t N) {
  // 4 environment buffers, the special-node matrix, phi of the special site, one partial-logit buffer, and the
  // fp16 node images of the longer of the two chains
  return 4 * B * Ds + (int64_t)2 * Ds * L * Ds + 2 * B + B * L + 4096 /*alignment*/ +
         (chain_tc_supported(Ds) ? chain_tc_image_floats(Ds, N) + 1024 : 0) +
         (forward_tc_supported(Ds, L, 2) ? (int64_t)4 * L * Ds * Ds + forward_tc_image_floats(Ds, L) + 1024 : 0);
}

extern "C" int trmps_predict(const trmps_predict_desc* p, void* stream) {
  if (!p || !p->feat || !p->nodes || !p->special || !p->dims || !p->logits || !p->work) {
    set_error("trmps_predict: null pointer");
    return TRMPS_E_ARG;
  }
  if (p->d != 2 || p->L < 1 || p->L > TRMPS_LMAX || p->Ds < 2 * p->L || p->Ds % 8 != 0 || p->N < 1 || p->loc < 0 ||
      p->loc >= p->N || p->B < 0) {
    set_error("trmps_predict: unsupported shape (need d=2, 1<=L<=16, Ds%%8==0, Ds>=2L, 0<=loc<N)");
    return TRMPS_E_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t B = p->B;
  const int Ds = p->Ds, L = p->L, N = p->N, fm = p->feature_map;
  if (B == 0) return TRMPS_OK;
  float* buf[4] = {p->work, p->work + B * Ds, p->work + 2 * B * Ds, p->work + 3 * B * Ds};
  float* M = p->work + 4 * B * Ds;
  float* phi_loc = M + (int64_t)2 * Ds * L * Ds;
  float* fpart = phi_loc + 2 * B;
  const size_t slot = (size_t)2 * Ds * Ds;
  int rc;
  // fused tensor-core chains unless their accumulated rounding bias would use up north_star's 1e-4 on the logits
  // (only chains longer than the BASELINE configurations: > 1500 sites at D <= 32, > 1000 at D <= 64, > 230 at D <= 128);
  // option 2 = 2 keeps them regardless
  const bool chain_fused = chain_tc_supported(Ds) &&
                           (option_value(2) == 2 || (option_value(2) == 0 && chain_tc_bias_estimate(Ds, N - 1) <= 9e-5));
  if (chain_fused) {
    // the environments never leave the SM between sites
    float* img = fpart + B * L;
    img += (256 - ((uintptr_t)img / sizeof(float)) % 256) % 256;      // bulk copies want 16-byte alignment; keep 1 KB
    if ((rc = launch_chain_tc(p->feat, fm, B, Ds, L, p->nodes, 0, +1, p->loc, 0, img, buf[0], st))) return rc;
    if ((rc = launch_chain_tc(p->feat, fm, B, Ds, L, p->nodes, N - 1, -1, N - 1 - p->loc, 1, img, buf[2], st))) return rc;
    if ((rc = trmps::launch_phi_site(p->feat + feat_site_offset(fm, B, p->loc), fm, B, phi_loc, st))) return rc;
    if (option_value(1) != 1 && forward_tc_supported(Ds, L, 2)) {
      // special-node contraction (mps.py:552-557) on the tensor cores: a bond whose far site is the identity
      float* M2 = img + chain_tc_image_floats(Ds, N);
      M2 += (256 - ((uintptr_t)M2 / sizeof(float)) % 256) % 256;
      float* bimg = M2 + (int64_t)4 * L * Ds * Ds;
      float* phi_far = buf[1];              // unused by the fused chains
      if ((rc = launch_special_to_bond(p->special, M2, phi_far, Ds, L, B, +1, st))) return rc;
      if ((rc = launch_forward_tc(buf[0], buf[2], phi_loc, phi_far, M2, bimg, fpart, B, Ds, L, p->dims + p->loc,
                                  p->dims + p->loc + 1, st)))
        return rc;
      return launch_sum_round_logits(fpart, 1, B, L, p->round_output, p->logits, st);
    }
    if ((rc = launch_special_to_matrix(p->special, M, Ds, L, st))) return rc;
    if ((rc = launch_forward(buf[0], buf[2], phi_loc, nullptr, M, fpart, 1, B, Ds, L, 1, p->dims + p->loc, st))) return rc;
    return launch_sum_round_logits(fpart, 1, B, L, p->round_output, p->logits, st);
  }
  // left chain (mps.py:536-541)
  int cur = 0;
  if ((rc = launch_fill_boundary(buf[cur], B, Ds, L, 0, st))) return rc;
  for (int c = 0; c < p->loc; ++c) {
    rc = launch_env_step(p->feat + feat_site_offset(fm, B, c), fm, B, Ds, +1, buf[cur], p->nodes + slot * c,
                         p->dims + c, buf[cur ^ 1], st);
    if (rc) return rc;
    cur ^= 1;
  }
  const float* C1 = buf[cur];
  // right chain (mps.py:544-549)
  int cur2 = 2;
  if ((rc = launch_fill_boundary(buf[cur2], B, Ds, L, 1, st))) return rc;
  for (int c = N - 1; c > p->loc; --c) {
    int nxt = cur2 == 2 ? 3 : 2;
    rc = launch_env_step(p->feat + feat_site_offset(fm, B, c), fm, B, Ds, -1, buf[cur2], p->nodes + slot * c,
                         p->dims + c + 1, buf[nxt], st);
    if (rc) return rc;
    cur2 = nxt;
  }
  const float* C2 = buf[cur2];
  // special node contraction (mps.py:552-557)
  if ((rc = launch_special_to_matrix(p->special, M, Ds, L, st))) return rc;
  if ((rc = trmps::launch_phi_site(p->feat + feat_site_offset(fm, B, p->loc), fm, B, phi_loc, st))) return rc;
  if ((rc = launch_forward(C1, C2, phi_loc, nullptr, M, fpart, 1, B, Ds, L, 1, p->dims + p->loc, st))) return rc;
  return launch_sum_round_logits(fpart, 1, B, L, p->round_output, p->logits, st);
}
