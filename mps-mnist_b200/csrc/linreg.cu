// Linear pre-training of the MPS (MPS._lin_reg, mps.py:331-399): plain gradient descent on a softmax regression
// (squared error for sqMPS, squaredDistanceMPS.py:29-30) over phi[..., 1], zero initial weights, one minibatch of
// `batch` consecutive samples (wrapping around, preprocessing.py:165-208) per iteration.
//
// The reference runs one sess.run per iteration; here ALL iterations run inside one persistent CTA: the model (N x L
// weights + L biases) lives in shared memory for the whole run, every iteration reads its minibatch from the
// device-resident dataset (L2 after the first touch) in chunks of 256 sites, once for the logits and once for the
// weight update.  The work per iteration is tiny (4 N L batch flops = 0.8 .. 3 MFLOP) and strictly sequential in the
// iteration index, so this is a latency-bound kernel: what matters is that there are no launches, no host round trips
// and no global-memory traffic for the model between iterations.  Sums run in a fixed order (deterministic).
#include "common.cuh"

namespace trmps {

namespace {

constexpr int LR_THREADS = 1024, LR_WARPS = LR_THREADS / 32, LR_CH = 256, LR_LDX = LR_CH + 1, LR_MAXB = 128, LR_GLD = 16, LR_TG = 4;

struct LinregArgs {
  const float* data; const float* labels; int64_t total; int N, L; int64_t start; int iterations, batch; float lr; int loss_kind;
  float* W; float* b;
};

// Shared-memory traffic is what bounds an iteration (one load instruction per clock and SM), so both passes are
// register tiled: the logits pass gives a warp four samples, splits the labels over its half-warps and lets the 16
// lanes of a half stride over the sites (10 loads per 24 FMAs at L = 10, shuffle reduction at the end), the update pass gives a thread one site and a quarter of the samples with
// the residual rows fetched as float4 broadcasts (4 loads per LT FMAs, partials summed in fixed order).
// LT = L rounded up to a multiple of 4.  Shared layout: Ws [N][LW] (LW odd: conflict-free), xs [B][257], zs/gs [B][16],
// pb [4][LT][256] partial updates, bs [16].
template <int LT>
__global__ void __launch_bounds__(LR_THREADS) linreg_kernel(LinregArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int N = a.N, L = a.L, B = a.batch, LW = L | 1;
  float* gs = sm;                                // [B][16], float4 aligned
  float* zs = gs + LR_MAXB * LR_GLD;             // [B][16]
  float* bs = zs + LR_MAXB * LR_GLD;             // [16]
  float* pb = bs + LR_GLD;                       // [LR_TG][LT][LR_CH]
  float* xs = pb + LR_TG * LT * LR_CH;           // [B][LR_LDX]
  float* Ws = xs + (size_t)B * LR_LDX;           // [N][LW]
  const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
  for (int e = tid; e < N * LW; e += LR_THREADS) Ws[e] = 0.f;         // tf.zeros, mps.py:336-337
  for (int e = tid; e < LR_MAXB * LR_GLD; e += LR_THREADS) gs[e] = 0.f;
  if (tid < LR_GLD) bs[tid] = 0.f;
  __syncthreads();
  const int nchunks = (N + LR_CH - 1) / LR_CH;
  // x of sample t, site n = phi[..., 1] of the dataset row (start + it * batch + t) mod total
  auto load_chunk = [&](int it, int c0) {
    const int w = min(LR_CH, N - c0), n = B * w;
    for (int e0 = tid; e0 < n; e0 += 8 * LR_THREADS) {     // eight independent loads in flight per thread
      float v[8];
      int dst[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int e = e0 + k * LR_THREADS;
        dst[k] = -1;
        if (e < n) {
          const int t = e / w, j = e - t * w;
          int64_t r = a.start + (int64_t)it * B + t;
          if (r >= a.total) r %= a.total;
          v[k] = __ldg(a.data + (r * N + c0 + j) * 2 + 1);
          dst[k] = t * LR_LDX + j;
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (dst[k] >= 0) xs[dst[k]] = v[k];
    }
  };
  for (int it = 0; it < a.iterations; ++it) {
    // ---- logits z = x W + b                                                     mps.py:349
    // warp wq owns samples 4 wq .. 4 wq + 3; its half-warps split the labels, the 16 lanes of a half stride over the sites
    constexpr int LH = LT / 2;
    const int half = lane >> 4, jl = lane & 15, t0 = 4 * wq;
    float acc[4][LH];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int l = 0; l < LH; ++l) acc[u][l] = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const int c0 = c * LR_CH, w = min(LR_CH, N - c0);
      __syncthreads();
      load_chunk(it, c0);
      __syncthreads();
      if (t0 < B) {
        const float* xr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) xr[u] = xs + min(t0 + u, B - 1) * LR_LDX;
        for (int j = jl; j < w; j += 16) {
          const float* wr = Ws + (size_t)(c0 + j) * LW + half * LH;
          float wv[LH];
#pragma unroll
          for (int l = 0; l < LH; ++l) wv[l] = half * LH + l < L ? wr[l] : 0.f;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float x = xr[u][j];
#pragma unroll
            for (int l = 0; l < LH; ++l) acc[u][l] = fmaf(x, wv[l], acc[u][l]);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int l = 0; l < LH; ++l) {
        float v = acc[u][l];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);      // within the half-warp, fixed order
        const int t = t0 + u, lab = half * LH + l;
        if (jl == 0 && t < B && lab < L) zs[t * LR_GLD + lab] = v + bs[lab];
      }
    __syncthreads();
    // ---- d cost / d z: (softmax(z) - y) / batch for the mean cross-entropy (mps.py:327-329), z - y for sqMPS
    if (tid < B) {
      int64_t r = a.start + (int64_t)it * B + tid;
      if (r >= a.total) r %= a.total;
      const float* y = a.labels + r * L;
      const float* z = zs + tid * LR_GLD;
      float* g = gs + tid * LR_GLD;
      if (a.loss_kind == TRMPS_LOSS_SQUARED) {
        for (int l = 0; l < L; ++l) g[l] = z[l] - __ldg(y + l);
      } else {
        float m = z[0];
        for (int l = 1; l < L; ++l) m = fmaxf(m, z[l]);
        float s = 0.f;
        for (int l = 0; l < L; ++l) s += expf(z[l] - m);
        const float inv = 1.f / s, ib = 1.f / (float)B;
        for (int l = 0; l < L; ++l) g[l] = (expf(z[l] - m) * inv - __ldg(y + l)) * ib;
      }
    }
    __syncthreads();
    if (tid < L) {                                   // bias -= lr * sum_t g
      float s = 0.f;
      for (int t = 0; t < B; ++t) s += gs[t * LR_GLD + tid];
      bs[tid] = fmaf(-a.lr, s, bs[tid]);
    }
    // ---- W -= lr * x^T g                                                        GradientDescentOptimizer, mps.py:352-353
    for (int c = 0; c < nchunks; ++c) {
      const int c0 = c * LR_CH, w = min(LR_CH, N - c0);
      if (nchunks > 1) {                             // a single chunk is still in shared memory from the logits pass
        __syncthreads();
        load_chunk(it, c0);
        __syncthreads();
      }
      {
        const int tg = wq / (LR_WARPS / LR_TG), j = (wq % (LR_WARPS / LR_TG)) * 32 + lane;     // 4 sample groups x 256 sites
        if (j < w) {
          float up[LT];
#pragma unroll
          for (int l = 0; l < LT; ++l) up[l] = 0.f;
          for (int t = tg; t < B; t += LR_TG) {
            const float x = xs[t * LR_LDX + j];
#pragma unroll
            for (int l4 = 0; l4 < LT / 4; ++l4) {
              const float4 g4 = *reinterpret_cast<const float4*>(gs + t * LR_GLD + 4 * l4);
              up[4 * l4 + 0] = fmaf(x, g4.x, up[4 * l4 + 0]);
              up[4 * l4 + 1] = fmaf(x, g4.y, up[4 * l4 + 1]);
              up[4 * l4 + 2] = fmaf(x, g4.z, up[4 * l4 + 2]);
              up[4 * l4 + 3] = fmaf(x, g4.w, up[4 * l4 + 3]);
            }
          }
#pragma unroll
          for (int l = 0; l < LT; ++l) pb[(tg * LT + l) * LR_CH + j] = up[l];
        }
      }
      __syncthreads();
      for (int e = tid; e < w * L; e += LR_THREADS) {
        const int l = e / w, j = e - l * w;
        float s = 0.f;
#pragma unroll
        for (int tg = 0; tg < LR_TG; ++tg) s += pb[(tg * LT + l) * LR_CH + j];
        Ws[(size_t)(c0 + j) * LW + l] = fmaf(-a.lr, s, Ws[(size_t)(c0 + j) * LW + l]);
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < N * L; e += LR_THREADS) a.W[e] = Ws[(size_t)(e / L) * LW + e % L];
  if (tid < L) a.b[tid] = bs[tid];
}

size_t linreg_smem(int N, int L, int B) {
  const int LT = (L + 3) / 4 * 4;
  return ((size_t)2 * LR_MAXB * LR_GLD + LR_GLD + (size_t)LR_TG * LT * LR_CH + (size_t)B * LR_LDX + (size_t)N * (L | 1)) * sizeof(float);
}

template <int LT>
int launch_linreg(const LinregArgs& a, size_t smem, cudaStream_t st) {
  TRMPS_CUDA_OK(cudaFuncSetAttribute(linreg_kernel<LT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  linreg_kernel<LT><<<1, LR_THREADS, smem, st>>>(a);
  TRMPS_LAUNCH_OK("linreg_kernel");
  return TRMPS_OK;
}

}  // namespace

}  // namespace trmps

using namespace trmps;

extern "C" int trmps_linreg_pretrain(const float* data, const float* labels, int64_t total, int32_t N, int32_t L, int64_t start,
                                     int32_t iterations, int32_t batch, float learning_rate, int32_t loss_kind, float* weight,
                                     float* bias, void* stream) {
  if (!data || !labels || !weight || !bias || total <= 0 || N <= 0 || L < 1 || L > TRMPS_LMAX || start < 0 || start >= total ||
      iterations < 0 || batch < 1 || batch > LR_MAXB || (loss_kind != TRMPS_LOSS_SOFTMAX_XENT && loss_kind != TRMPS_LOSS_SQUARED)) {
    set_error("trmps_linreg_pretrain: bad argument (need 1 <= L <= 16, 1 <= batch <= 128, 0 <= start < total)");
    return TRMPS_E_ARG;
  }
  const size_t smem = linreg_smem(N, L, batch);
  if (smem > 227 * 1024) {
    set_error("trmps_linreg_pretrain: N = %d, L = %d does not fit the shared-memory resident model (N * L up to about 12000)", N, L);
    return TRMPS_E_ARG;
  }
  LinregArgs a;
  a.data = data; a.labels = labels; a.total = total; a.N = N; a.L = L; a.start = start; a.iterations = iterations; a.batch = batch;
  a.lr = learning_rate; a.loss_kind = loss_kind; a.W = weight; a.b = bias;
  cudaStream_t st = (cudaStream_t)stream;
  const int LT = (L + 3) / 4 * 4;
  return LT == 4 ? launch_linreg<4>(a, smem, st) : LT == 8 ? launch_linreg<8>(a, smem, st)
       : LT == 12 ? launch_linreg<12>(a, smem, st) : launch_linreg<16>(a, smem, st);
}
