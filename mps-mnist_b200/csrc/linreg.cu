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

constexpr int LR_THREADS = 1024, LR_CH = 256, LR_MAXB = 128;

struct LinregArgs {
  const float* data; const float* labels; int64_t total; int N, L; int64_t start; int iterations, batch; float lr; int loss_kind;
  float* W; float* b;
};

__global__ void __launch_bounds__(LR_THREADS) linreg_kernel(LinregArgs a) {
  extern __shared__ float sm[];
  const int N = a.N, L = a.L, B = a.batch;
  float* Ws = sm;                               // [N][L]
  float* xs = Ws + (size_t)N * L;               // [B][LR_CH + 1]
  float* zs = xs + (size_t)LR_MAXB * (LR_CH + 1);   // [B][L]
  float* gs = zs + LR_MAXB * TRMPS_LMAX;        // [B][L]
  float* bs = gs + LR_MAXB * TRMPS_LMAX;        // [L]
  const int tid = threadIdx.x;
  for (int e = tid; e < N * L; e += LR_THREADS) Ws[e] = 0.f;          // tf.zeros, mps.py:336-337
  if (tid < L) bs[tid] = 0.f;
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
          dst[k] = t * (LR_CH + 1) + j;
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (dst[k] >= 0) xs[dst[k]] = v[k];
    }
  };
  for (int it = 0; it < a.iterations; ++it) {
    for (int e = tid; e < B * L; e += LR_THREADS) zs[e] = bs[e % L];
    // ---- logits z = x W + b                                                     mps.py:349
    for (int c = 0; c < nchunks; ++c) {
      const int c0 = c * LR_CH, w = min(LR_CH, N - c0);
      __syncthreads();
      load_chunk(it, c0);
      __syncthreads();
      if (tid < B * L) {
        const int t = tid / L, l = tid - t * L;
        const float* xr = xs + t * (LR_CH + 1);
        const float* wc = Ws + (size_t)c0 * L + l;
        float acc = 0.f;
        for (int j = 0; j < w; ++j) acc = fmaf(xr[j], wc[(size_t)j * L], acc);
        zs[tid] += acc;
      }
    }
    __syncthreads();
    // ---- d cost / d z: (softmax(z) - y) / batch for the mean cross-entropy (mps.py:327-329), z - y for sqMPS
    if (tid < B) {
      int64_t r = a.start + (int64_t)it * B + tid;
      if (r >= a.total) r %= a.total;
      const float* y = a.labels + r * L;
      float* z = zs + tid * L;
      float* g = gs + tid * L;
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
      for (int t = 0; t < B; ++t) s += gs[t * L + tid];
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
      for (int e = tid; e < w * L; e += LR_THREADS) {
        const int j = e / L, l = e - j * L;
        float acc = 0.f;
        for (int t = 0; t < B; ++t) acc = fmaf(xs[t * (LR_CH + 1) + j], gs[t * L + l], acc);
        Ws[(size_t)(c0 + j) * L + l] = fmaf(-a.lr, acc, Ws[(size_t)(c0 + j) * L + l]);
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < N * L; e += LR_THREADS) a.W[e] = Ws[e];
  if (tid < L) a.b[tid] = bs[tid];
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
  const size_t smem = ((size_t)N * L + (size_t)LR_MAXB * (LR_CH + 1) + 2 * LR_MAXB * TRMPS_LMAX + TRMPS_LMAX) * sizeof(float);
  if (smem > 227 * 1024) {
    set_error("trmps_linreg_pretrain: N*L = %d does not fit the shared-memory resident model (max about 21000)", N * L);
    return TRMPS_E_ARG;
  }
  TRMPS_CUDA_OK(cudaFuncSetAttribute(linreg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LinregArgs a;
  a.data = data; a.labels = labels; a.total = total; a.N = N; a.L = L; a.start = start; a.iterations = iterations; a.batch = batch;
  a.lr = learning_rate; a.loss_kind = loss_kind; a.W = weight; a.b = bias;
  linreg_kernel<<<1, LR_THREADS, smem, (cudaStream_t)stream>>>(a);
  TRMPS_LAUNCH_OK("linreg_kernel");
  return TRMPS_OK;
}
