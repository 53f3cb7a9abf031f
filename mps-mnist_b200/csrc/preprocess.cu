// Datasource side of the path, fused: what the reference does on the host with one sess.run per image
// (MNISTpreprocessing.py:14-78: 2x2 average pool, flatten, feature map) and one NumPy copy per training step
// (preprocessing.py:165-208: slice with wrap-around, swapaxes(0,1)) becomes one pass over the raw images that writes
// the site-major batch the sweep kernels read.
//
// HBM bound (no reuse): algorithmic bytes per image = rows*cols*sizeof(T) read + N*4 (pixels) or N*8 (phi) written.
// One CTA moves a tile of 128 images x one strip of the image: every image contributes a contiguous run of a few
// hundred bytes (whole 32-byte sectors, 16-byte vector loads), the transposition to site-major goes through a padded
// shared-memory tile, and every site row is written as 128 consecutive images (512 B / 1 KB per row).
#include "common.cuh"

namespace trmps {

namespace {

constexpr int PRE_TI = 128;          // images per tile
constexpr int PRE_LD = PRE_TI + 1;   // padded tile row (conflict-free transposed access)
constexpr int PRE_STRIP = 64;        // floats per image and strip in the non-pooled kernels
constexpr int PRE_MAXCOLS = 64;      // widest image the pooled kernel takes (sites per strip = cols/2 <= 32)

__device__ __forceinline__ int64_t wrap_row(int64_t start, int64_t t, int64_t total) {
  int64_t r = start + t;                       // next_training_data_batch wraps to index 0 (preprocessing.py:183-202)
  if (r >= total) r %= total;
  return r;
}

// Raw pixel formats.  U8 follows DataSet.__init__ of the TF-1.2 MNIST loader (float32(u) * float32(1/255));
// U8_DIV follows FashionMNISTpreprocessing.py:60 (u / 255.0 in float64, float32 when fed to the graph: the correctly
// rounded quotient, which the float32 division gives for every u in 0..255 -- 126 of the 256 values differ from U8).
struct U8Mul { uint8_t v; };
struct U8Div { uint8_t v; };
__device__ __forceinline__ float px(float v, const float*) { return v; }
// __fmul_rn / __fdiv_rn: rounded on their own (no contraction into the pooling adds)
__device__ __forceinline__ float px(uint8_t v, const U8Mul*) { return __fmul_rn((float)v, (float)(1.0 / 255.0)); }
__device__ __forceinline__ float px(uint8_t v, const U8Div*) { return __fdiv_rn((float)v, 255.0f); }

__device__ __forceinline__ float4 load4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
template <typename T>
__device__ __forceinline__ float4 load4(const T* p) {
  const uchar4 q = __ldcs(reinterpret_cast<const uchar4*>(p));
  return make_float4(px(q.x, p), px(q.y, p), px(q.z, p), px(q.w, p));
}
__device__ __forceinline__ float load1(const float* p) { return __ldcs(p); }
template <typename T>
__device__ __forceinline__ float load1(const T* p) { return px(__ldcs(reinterpret_cast<const uint8_t*>(p)), p); }

// tile[s][i] (s = site inside the strip, i = image inside the tile, row pitch TI + 1) -> out, site-major.
// A thread keeps its image and walks down the sites: one shared load, one store and a pointer bump per value.
template <int TI>
__device__ __forceinline__ void write_sites(const float* tile, int nsites, int64_t site0, int64_t t0, int64_t B, int out_map,
                                            float* __restrict__ out) {
  constexpr int STEP = 256 / TI;
  const int i = threadIdx.x % TI, s0 = threadIdx.x / TI;
  if (t0 + i >= B) return;
  const float* src = tile + s0 * (TI + 1) + i;
  const int64_t o = (site0 + s0) * B + t0 + i;
  if (out_map == TRMPS_FM_PRECOMPUTED) {
    float* po = out + o;
    for (int s = s0; s < nsites; s += STEP, src += STEP * (TI + 1), po += STEP * B) __stcs(po, *src);
  } else {
    float2* po = reinterpret_cast<float2*>(out) + o;
    for (int s = s0; s < nsites; s += STEP, src += STEP * (TI + 1), po += STEP * B) __stcs(po, phi_of_pixel(*src, out_map));
  }
}

// Work split shared by both kernels: blockIdx.x = tile * nstrips + strip, so that CTAs resident at the same time read
// neighbouring strips of the SAME images (whole DRAM pages instead of one short run per page), and every thread issues
// all the 16-byte loads of a pass before it touches the first result (UNROLL loads in flight per thread).
constexpr int PRE_UNROLL = 8;

// shrink: tf.nn.avg_pool(ksize 2x2, stride 2) then reshape to [rows/2 * cols/2] (MNISTpreprocessing.py:44-48), or the
// 2x2 block maximum of the Fashion-MNIST datasource (MAXP).
// strip = pooled image row: the two raw rows under it (2*cols consecutive pixels) of 128 images.
template <typename T, bool MAXP>
__global__ void __launch_bounds__(256) preprocess_pool_kernel(const T* __restrict__ images, int64_t total, int rows, int cols,
                                                              int64_t start, int64_t B, int out_map, float* __restrict__ out) {
  __shared__ float tile[(PRE_MAXCOLS / 2) * PRE_LD];
  const int nstrips = rows / 2;
  const int64_t t0 = (int64_t)(blockIdx.x / nstrips) * PRE_TI;
  const int pr = blockIdx.x % nstrips, Q = cols / 4;
  const int64_t img_stride = (int64_t)rows * cols;
  const int n_items = PRE_TI * Q, n_img = (int)min((int64_t)PRE_TI, B - t0);
  constexpr int U = PRE_UNROLL / 2;                    // two loads per item
  for (int base = threadIdx.x; base < n_items; base += 256 * U) {
    float4 a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * 256, i = idx / Q, q = idx - i * Q;
      if (idx < n_items && i < n_img) {
        const T* p = images + wrap_row(start, t0 + i, total) * img_stride + (int64_t)(2 * pr) * cols + 4 * q;
        a[u] = load4(p);
        b[u] = load4(p + cols);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * 256, i = idx / Q, q = idx - i * Q;
      if (idx < n_items && i < n_img) {
        if (MAXP) {      // skimage.measure.block_reduce(image, (2, 2), np.max), FashionMNISTpreprocessing.py:61-65
          tile[(2 * q) * PRE_LD + i] = fmaxf(fmaxf(a[u].x, a[u].y), fmaxf(b[u].x, b[u].y));
          tile[(2 * q + 1) * PRE_LD + i] = fmaxf(fmaxf(a[u].z, a[u].w), fmaxf(b[u].z, b[u].w));
        } else {
          tile[(2 * q) * PRE_LD + i] = ((a[u].x + a[u].y) + (b[u].x + b[u].y)) * 0.25f;
          tile[(2 * q + 1) * PRE_LD + i] = ((a[u].z + a[u].w) + (b[u].z + b[u].w)) * 0.25f;
        }
      }
    }
  }
  __syncthreads();
  write_sites<PRE_TI>(tile, cols / 2, (int64_t)pr * (cols / 2), t0, B, out_map, out);
}

// uint8 images, pooled: the byte format of the idx files.  4-byte loads per lane make the generic kernel instruction bound,
// so here one thread owns one image's strip (the two raw rows under a pooled row: 2 * COLS bytes, 8-byte aligned) and reads
// it with 8-byte loads (L1 keeps the sectors the neighbouring loads share); 256 images per CTA, sites written as 1 KB rows.
template <typename T, bool MAXP, int COLS>
__global__ void __launch_bounds__(256) preprocess_pool_bytes_kernel(const T* __restrict__ images, int64_t total, int rows, int64_t start,
                                                                    int64_t B, int out_map, float* __restrict__ out) {
  constexpr int TI = 256, NW = 2 * COLS / 8, NS = COLS / 2;
  static_assert(COLS % 4 == 0, "row pairs must be whole 8-byte words");
  __shared__ float tile[NS * (TI + 1)];
  const int nstrips = rows / 2;
  const int64_t t0 = (int64_t)(blockIdx.x / nstrips) * TI;
  const int pr = blockIdx.x % nstrips, i = threadIdx.x;
  if (t0 + i < B) {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(images) + wrap_row(start, t0 + i, total) * ((int64_t)rows * COLS) + (int64_t)(2 * pr) * COLS;
    uint2 w[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) w[k] = __ldg(reinterpret_cast<const uint2*>(p) + k);
    auto byte_at = [&](int b) -> float {              // b is a compile-time constant after unrolling
      const uint32_t word = (b & 4) ? w[b >> 3].y : w[b >> 3].x;
      return px((uint8_t)((word >> (8 * (b & 3))) & 0xffu), images);
    };
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const float a0 = byte_at(2 * s), a1 = byte_at(2 * s + 1), b0 = byte_at(COLS + 2 * s), b1 = byte_at(COLS + 2 * s + 1);
      tile[s * (TI + 1) + i] = MAXP ? fmaxf(fmaxf(a0, a1), fmaxf(b0, b1)) : ((a0 + a1) + (b0 + b1)) * 0.25f;
    }
  }
  __syncthreads();
  write_sites<TI>(tile, NS, (int64_t)pr * NS, t0, B, out_map, out);
}

// four values starting at p; the scalar variant serves rows whose length is not a multiple of 4 (no 16-byte alignment)
template <bool VEC, typename T>
__device__ __forceinline__ float4 load_quad(const T* p, int valid) {
  if (VEC) return load4(p);
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid > 0) r.x = load1(p);
  if (valid > 1) r.y = load1(p + 1);
  if (valid > 2) r.z = load1(p + 2);
  if (valid > 3) r.w = load1(p + 3);
  return r;
}

// shrink = False (MNISTpreprocessing.py:51-53), and the batch slice of an already mapped dataset
// (preprocessing.py:165-208) with C = 2 floats per site: rows of `len` values -> (len / C, B, C).
// One CTA: TI images x one strip of STRIP consecutive values of every row (TI * STRIP = 8192: eight quads per thread).
template <typename T, int C, int TI, int STRIP, bool VEC>
__global__ void __launch_bounds__(256) preprocess_flat_kernel(const T* __restrict__ data, int64_t total, int len, int nstrips,
                                                              int64_t start, int64_t B, int out_map, float* __restrict__ out) {
  // C == 1: tile[s][i] lives at s * TI + (i ^ swz(s)), swz(s) = VW * (s / 4), VW = 32 / (quads per image).  The quads a warp
  // transposes in one step land in different VW-bank groups (conflict-free stores) and VW consecutive images of a site
  // stay one aligned vector (conflict-free vector loads for the site-major write).  C == 2 keeps the padded tile.
  constexpr int Q = STRIP / 4, VW = 32 / Q, U = TI * Q / 256;
  constexpr int LD = C == 1 ? TI : TI + 1;
  extern __shared__ __align__(16) float tile[];          // STRIP * LD floats
  const int tid = threadIdx.x;
  const int64_t t0 = (int64_t)(blockIdx.x / nstrips) * TI;
  const int e0 = (blockIdx.x % nstrips) * STRIP;
  const int n = min(STRIP, len - e0);
  const int n_img = (int)min((int64_t)TI, B - t0);
  static_assert(C == 2 || VW == 2 || VW == 4, "8 or 16 quads per image");
  float4 a[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int idx = tid + u * 256, i = idx / Q, q = idx % Q;
    if (i < n_img && 4 * q < n)
      a[u] = load_quad<VEC>(data + wrap_row(start, t0 + i, total) * (int64_t)len + e0 + 4 * q, n - 4 * q);
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int idx = tid + u * 256, i = idx / Q, q = idx % Q;
    if (i < n_img && 4 * q < n) {
      float* col = tile + (4 * q) * LD + (C == 1 ? (i ^ (VW * q)) : i);      // rows 4q..4q+3 share swz = VW * q
      col[0] = a[u].x; col[LD] = a[u].y; col[2 * LD] = a[u].z; col[3 * LD] = a[u].w;
    }
  }
  __syncthreads();
  if (C == 1) {
    // TI / VW threads write the TI images of one site (VW images each), 256 * VW / TI sites per pass
    constexpr int TPS = TI / VW, SPP = 256 / TPS;
    const int j = VW * (tid % TPS);
    const int64_t t = t0 + j;
    const bool vec = (B % VW == 0) && j + VW - 1 < n_img;
    for (int s = tid / TPS; s < n; s += SPP) {
      float v[VW];
      const float* src = tile + s * LD + (j ^ (VW * (s >> 2)));
      if (VW == 4) {
        const float4 x = *reinterpret_cast<const float4*>(src);
        v[0] = x.x; v[1] = x.y; v[VW - 2] = x.z; v[VW - 1] = x.w;
      } else {
        const float2 x = *reinterpret_cast<const float2*>(src);
        v[0] = x.x; v[1] = x.y;
      }
      const int64_t o = (int64_t)(e0 + s) * B + t;
      if (out_map == TRMPS_FM_PRECOMPUTED) {
        if (vec) {
          if (VW == 4) __stcs(reinterpret_cast<float4*>(out + o), make_float4(v[0], v[1], v[VW - 2], v[VW - 1]));
          else __stcs(reinterpret_cast<float2*>(out + o), make_float2(v[0], v[1]));
        } else {
#pragma unroll
          for (int k = 0; k < VW; ++k)
            if (j + k < n_img) out[o + k] = v[k];
        }
      } else {
        float2* po = reinterpret_cast<float2*>(out) + o;
#pragma unroll
        for (int k = 0; k < VW; k += 2) {
          const float2 p0 = phi_of_pixel(v[k], out_map), p1 = phi_of_pixel(v[k + 1], out_map);
          if (vec) {
            __stcs(reinterpret_cast<float4*>(po + k), make_float4(p0.x, p0.y, p1.x, p1.y));
          } else {
            if (j + k < n_img) po[k] = p0;
            if (j + k + 1 < n_img) po[k + 1] = p1;
          }
        }
      }
    }
  } else {
    for (int idx = tid; idx < (n / 2) * TI; idx += 256) {
      const int s = idx / TI, i = idx % TI;
      if (t0 + i >= B) continue;
      __stcs(reinterpret_cast<float2*>(out) + (int64_t)(e0 / 2 + s) * B + t0 + i, make_float2(tile[(2 * s) * LD + i], tile[(2 * s + 1) * LD + i]));
    }
  }
}

int g_flat_variant = 0;      // development switch (trmps_set_option 3): tile shape of the C = 1 kernel

template <typename T, int C, int TI, int STRIP>
int launch_flat_shape(const T* data, int64_t total, int64_t len, int64_t start, int64_t B, int out_map, float* out, cudaStream_t st,
                      const char* who) {
  const int64_t tiles = ceil_div64(B, TI), nstrips = ceil_div64(len, STRIP);
  if (len > 0x7fffffff || tiles * nstrips > 0x7fffffff) {
    set_error("%s: batch too large", who);
    return TRMPS_E_ARG;
  }
  const unsigned grid = (unsigned)(tiles * nstrips);
  constexpr int smem = STRIP * (C == 1 ? TI : TI + 1) * (int)sizeof(float);
  auto kv = preprocess_flat_kernel<T, C, TI, STRIP, true>;
  auto ks = preprocess_flat_kernel<T, C, TI, STRIP, false>;
  if (smem > 48 * 1024) {
    TRMPS_CUDA_OK(cudaFuncSetAttribute(kv, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TRMPS_CUDA_OK(cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  if (len % 4 == 0)
    kv<<<grid, 256, smem, st>>>(data, total, (int)len, (int)nstrips, start, B, out_map, out);
  else
    ks<<<grid, 256, smem, st>>>(data, total, (int)len, (int)nstrips, start, B, out_map, out);
  TRMPS_LAUNCH_OK("preprocess_flat_kernel");
  return TRMPS_OK;
}

template <typename T, int C>
int launch_flat(const T* data, int64_t total, int64_t len, int64_t start, int64_t B, int out_map, float* out, cudaStream_t st,
                const char* who) {
  if (C == 2) return launch_flat_shape<T, C, 128, 64>(data, total, len, start, B, out_map, out, st, who);
  if (g_flat_variant == 1) return launch_flat_shape<T, 1, 256, 32>(data, total, len, start, B, out_map, out, st, who);
  if (g_flat_variant == 2) return launch_flat_shape<T, 1, 256, 64>(data, total, len, start, B, out_map, out, st, who);
  if (g_flat_variant == 3) return launch_flat_shape<T, 1, 512, 32>(data, total, len, start, B, out_map, out, st, who);
  return launch_flat_shape<T, 1, 128, 64>(data, total, len, start, B, out_map, out, st, who);
}

// labels of the same slice: (total, L) -> (B, L)
__global__ void gather_labels_kernel(const float* __restrict__ src, int64_t total, int L, int64_t start, int64_t B,
                                     float* __restrict__ dst) {
  const int64_t n = B * L;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = idx / L;
    dst[idx] = src[wrap_row(start, t, total) * L + (idx - t * L)];
  }
}

int bad(const char* msg) {
  set_error("%s", msg);
  return TRMPS_E_ARG;
}

}  // namespace

void set_flat_variant(int v) { g_flat_variant = v; }

}  // namespace trmps

using namespace trmps;

namespace {

template <typename T>
int preprocess_typed(const T* images, int64_t total, int rows, int cols, int pool, int64_t start, int64_t B, int out_map, float* out,
                     cudaStream_t st) {
  if (pool == TRMPS_POOL_NONE)
    return launch_flat<T, 1>(images, total, (int64_t)rows * cols, start, B, out_map, out, st, "trmps_preprocess_batch");
  if (rows % 2 || cols % 4 || cols > PRE_MAXCOLS)
    return bad("trmps_preprocess_batch: pooling needs an even number of rows and cols a multiple of 4, at most 64");
  if (sizeof(T) == 1 && cols == 28) {                 // MNIST / Fashion-MNIST bytes: one thread per image strip
    const int64_t tiles256 = ceil_div64(B, 256);
    if (tiles256 * (rows / 2) > 0x7fffffff) return bad("trmps_preprocess_batch: batch too large");
    const unsigned grid = (unsigned)(tiles256 * (rows / 2));
    if (pool == TRMPS_POOL_AVG) preprocess_pool_bytes_kernel<T, false, 28><<<grid, 256, 0, st>>>(images, total, rows, start, B, out_map, out);
    else preprocess_pool_bytes_kernel<T, true, 28><<<grid, 256, 0, st>>>(images, total, rows, start, B, out_map, out);
    TRMPS_LAUNCH_OK("preprocess_pool_bytes_kernel");
    return TRMPS_OK;
  }
  const int64_t tiles = ceil_div64(B, PRE_TI);
  if (tiles * (rows / 2) > 0x7fffffff) return bad("trmps_preprocess_batch: batch too large");
  const unsigned grid = (unsigned)(tiles * (rows / 2));
  if (pool == TRMPS_POOL_AVG) preprocess_pool_kernel<T, false><<<grid, 256, 0, st>>>(images, total, rows, cols, start, B, out_map, out);
  else preprocess_pool_kernel<T, true><<<grid, 256, 0, st>>>(images, total, rows, cols, start, B, out_map, out);
  TRMPS_LAUNCH_OK("preprocess_pool_kernel");
  return TRMPS_OK;
}

}  // namespace

extern "C" int trmps_preprocess_batch(const void* images, int32_t dtype, int64_t total, int32_t rows, int32_t cols, int32_t pool,
                                      int64_t start, int64_t B, int32_t out_map, float* out, void* stream) {
  if (!images || !out || total <= 0 || rows <= 0 || cols <= 0 || start < 0 || start >= total || B < 0)
    return bad("trmps_preprocess_batch: bad argument (need 0 <= start < total, B >= 0)");
  if (dtype < TRMPS_DT_F32 || dtype > TRMPS_DT_U8_DIV) return bad("trmps_preprocess_batch: dtype must be one of TRMPS_DT_*");
  if (pool < TRMPS_POOL_NONE || pool > TRMPS_POOL_MAX) return bad("trmps_preprocess_batch: pool must be one of TRMPS_POOL_*");
  if (out_map < TRMPS_FM_PRECOMPUTED || out_map > TRMPS_FM_COS_SIN) return bad("trmps_preprocess_batch: unknown feature map");
  if (B == 0) return TRMPS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TRMPS_DT_F32) return preprocess_typed((const float*)images, total, rows, cols, pool, start, B, out_map, out, st);
  if (dtype == TRMPS_DT_U8) return preprocess_typed((const U8Mul*)images, total, rows, cols, pool, start, B, out_map, out, st);
  return preprocess_typed((const U8Div*)images, total, rows, cols, pool, start, B, out_map, out, st);
}

extern "C" int trmps_batch_gather(const float* data, const float* labels, int64_t total, int32_t N, int32_t c, int32_t L,
                                  int64_t start, int64_t B, float* feat_out, float* labels_out, void* stream) {
  if (!data || !feat_out || total <= 0 || N <= 0 || (c != 1 && c != 2) || start < 0 || start >= total || B < 0)
    return bad("trmps_batch_gather: bad argument (c must be 1 or 2, 0 <= start < total)");
  if ((labels == nullptr) != (labels_out == nullptr) || (labels && L <= 0)) return bad("trmps_batch_gather: labels and labels_out go together");
  const int64_t len = (int64_t)N * c;
  if (B == 0) return TRMPS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int rc = c == 1 ? launch_flat<float, 1>(data, total, len, start, B, TRMPS_FM_PRECOMPUTED, feat_out, st, "trmps_batch_gather")
                        : launch_flat<float, 2>(data, total, len, start, B, TRMPS_FM_PRECOMPUTED, feat_out, st, "trmps_batch_gather");
  if (rc) return rc;
  if (labels) {
    const int64_t n = B * L;
    gather_labels_kernel<<<(unsigned)(ceil_div64(n, 256) < 1184 ? ceil_div64(n, 256) : 1184), 256, 0, st>>>(labels, total, L, start, B, labels_out);
    TRMPS_LAUNCH_OK("gather_labels_kernel");
  }
  return TRMPS_OK;
}
