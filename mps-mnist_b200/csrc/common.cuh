// Shared device helpers for libtrmps_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/trmps_b200.h"

#define TRMPS_THREADS 256
#define TRMPS_LMAX 16

namespace trmps {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define TRMPS_CUDA_OK(expr)                                              \
  do {                                                                   \
    cudaError_t _e = (expr);                                             \
    if (_e != cudaSuccess) return ::trmps::cuda_fail(_e, #expr);         \
  } while (0)

#define TRMPS_LAUNCH_OK(name)                                            \
  do {                                                                   \
    ::trmps::count_launch();                                             \
    cudaError_t _e = cudaGetLastError();                                 \
    if (_e != cudaSuccess) return ::trmps::cuda_fail(_e, name);          \
  } while (0)

void count_launch();

// optional per-category device timing (CUDA events on the launching stream); off unless trmps_profile_begin was called
enum ProfCat { PROF_ENV = 0, PROF_FORWARD, PROF_GRAD, PROF_SVD, PROF_MERGE, PROF_SCALAR, PROF_COMM, PROF_NCAT };
struct ProfScope {
  int cat; cudaStream_t st; bool on;
  ProfScope(int cat, cudaStream_t st);
  ~ProfScope();
};

__host__ __device__ __forceinline__ int align_up(int x, int a) { return (x + a - 1) / a * a; }
__host__ __device__ __forceinline__ int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// local feature map phi(x) for one pixel (d = 2).  Reference map: MNISTpreprocessing.py:55.
__device__ __forceinline__ float2 phi_of_pixel(float x, int kind) {
  if (kind == TRMPS_FM_ONE_SQRT3) return make_float2(1.0f, 1.7320508075688772f * (2.0f * x - 1.0f));
  if (kind == TRMPS_FM_ONE_X) return make_float2(1.0f, x);
  // TRMPS_FM_COS_SIN.  Pixels live in [0,1]: fold onto [0, pi/4] (1 - x is exact for x in [0.5,1]) and use the
  // classic single-precision minimax kernels (~1 ulp there); anything else takes the library path.
  float s, c;
  if (x >= 0.f && x <= 1.f) {
    const bool fold = x > 0.5f;
    const float t = 1.5707963267948966f * (fold ? 1.f - x : x), z = t * t;
    const float sp = fmaf(fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f) * z, t, t);
    const float cp = fmaf(fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f) * z, z,
                          fmaf(-0.5f, z, 1.f));
    s = fold ? cp : sp;
    c = fold ? sp : cp;
  } else {
    sincosf(1.5707963267948966f * x, &s, &c);
  }
  return make_float2(c, s);
}

// phi of image t at one site; feat_site points at this site's B pixels or B float2 phis
__device__ __forceinline__ float2 load_phi(const float* __restrict__ feat_site, int kind, int64_t t) {
  if (kind == TRMPS_FM_PRECOMPUTED) return __ldg(reinterpret_cast<const float2*>(feat_site) + t);
  return phi_of_pixel(__ldg(feat_site + t), kind);
}

__host__ __device__ __forceinline__ int64_t feat_site_offset(int kind, int64_t B, int site) {
  return (kind == TRMPS_FM_PRECOMPUTED ? 2 : 1) * B * (int64_t)site;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
  uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
  int bytes = valid ? 16 : 0;   // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// 8x8 (RQ=2) or 4x8 (RQ=1) register-tile outer-product step over KC k-rows.
// As: [k][TM] (rows ty*4.. and TM/2+ty*4..), Bs: [k][TN] (cols tx*4.. and TN/2+tx*4..)
template <int RQ, int KC>
__device__ __forceinline__ void mma_rows(const float* __restrict__ As, int lda, const float* __restrict__ Bs, int ldb,
                                         int a_off0, int a_off1, int b_off0, int b_off1, float (&acc)[4 * RQ][8]) {
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    float a[4 * RQ], b[8];
    float4 v = *reinterpret_cast<const float4*>(As + k * lda + a_off0);
    a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
    if (RQ == 2) {
      v = *reinterpret_cast<const float4*>(As + k * lda + a_off1);
      a[4 * (RQ - 1) + 0] = v.x; a[4 * (RQ - 1) + 1] = v.y; a[4 * (RQ - 1) + 2] = v.z; a[4 * (RQ - 1) + 3] = v.w;
    }
    v = *reinterpret_cast<const float4*>(Bs + k * ldb + b_off0);
    b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
    v = *reinterpret_cast<const float4*>(Bs + k * ldb + b_off1);
    b[4] = v.x; b[5] = v.y; b[6] = v.z; b[7] = v.w;
#pragma unroll
    for (int i = 0; i < 4 * RQ; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// deterministic block-wide sum; result valid in thread 0.  red must hold blockDim/32 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x == 0)
    for (int i = 0; i < nw; ++i) s += red[i];
  return s;
}

// ---- launch wrappers implemented in the .cu files (all asynchronous on `st`) -------------------------
struct BondGeom {        // the two sites of a bond update, direction-agnostic ("near" holds the special node)
  int near_site, far_site;
  int dir;
  const float* e_near;   // C1s[c] (right) or C2s[c] (left): (B, Ds)
  const float* e_far;    // C2s[c+1] (right) or C1s[c-1] (left)
  const int32_t* d_near; // device ptr: outer bond dim of the near site (rows i of the bond matrix)
  const int32_t* d_mid;  // device ptr: bond between the two sites (contracted by merge)
  const int32_t* d_far;  // device ptr: outer bond dim of the far site (cols k)
};

BondGeom bond_geom(const trmps_sweep* s, int site, int dir);

int launch_fill_boundary(float* env, int64_t B, int Ds, int L, int which, cudaStream_t st);
int launch_env_step(const float* feat_site, int fm, int64_t B, int Ds, int dir, const float* in_env,
                    const float* node, const int32_t* din_dev, float* out_env, cudaStream_t st);
int launch_forward(const float* e_near, const float* e_far, const float* phi_near, const float* phi_far,
                   const float* M, float* fpart, int nsplit, int64_t B, int Ds, int L, int nd,
                   const int32_t* d_near, cudaStream_t st);
int launch_special_to_matrix(const float* special, float* M, int Ds, int L, cudaStream_t st);
int launch_sum_round_logits(const float* fpart, int nsplit, int64_t B, int L, int round_output, float* logits, cudaStream_t st);

int launch_phi_site(const float* feat_site, int fm, int64_t B, float* out, cudaStream_t st);
int launch_phi_pair(const float* feat_a, const float* feat_b, int fm, int64_t B, float* out_a, float* out_b, cudaStream_t st);

constexpr int TRMPS_IMG_BLOCK = 256;      // images per CTA of the per-image kernels (loss, armijo); `red` has a row per CTA (128 was measured slower: scalar 17 -> 23 ms per cfg3 step)
struct LossArgs {
  const float* fpart; int nsplit; float* f0; const float* labels; const float* phi_far;
  float* resid; float* sres; float* hres; float* red; int64_t B; int L; int loss_kind; int from_parts;
  int32_t* arrive; float* cost_sum; int32_t* clear_word;     // last-CTA reduction: arrival counter, un-normalised cost sum; a word to zero (or null)
};
struct ArmijoArgs {
  const float* f0; const float* fdpart; int nsplit; float* fd; const float* labels; float* red;
  int64_t B; int L; int loss_kind; float lr;
  int32_t* arrive; float* trial; float trial_scale;          // last-CTA reduction: the 19 trial costs * trial_scale -> trial[]
  int decide; float* scal; int32_t* iscal; float armijo_coeff;   // decide != 0: the accept / revert decision in the same launch (single rank)
};
int launch_loss(const LossArgs& a, cudaStream_t st);
int launch_armijo(const ArmijoArgs& a, cudaStream_t st);
int launch_sum_red(const float* red, int nrows, int ncols, float scale, float* dst, cudaStream_t st);
int launch_grad_finish(float* G, const float* M, const float* H, float* delta, const float* parts, int nparts, int64_t n, float reg,
                       float hess_add, float gdc_scale, const float* cost_sum, float cost_scale, float* red, float* scal,
                       int32_t* arrive, int32_t* maxbits, int nblocks, cudaStream_t st);
int launch_decide(float* scal, const float* trial, int32_t* iscal, float lr, float coeff, cudaStream_t st);
int launch_apply_update(float* M, const float* delta, int64_t n, float* f0, const float* fd, int64_t nf, const float* scal,
                        cudaStream_t st);
int launch_merge(const float* special, const float* far_node, float* M, int Ds, int L, int dir, const int32_t* d_mid,
                 cudaStream_t st);
int launch_grad(const float* e_near, const float* e_far, const float* phi_near, const float* sres, float* gpart, int gsplit,
                int64_t B, int Ds, int L, bool squared, cudaStream_t st);
int launch_sum_parts(const float* part, int nparts, int64_t n, float* out, cudaStream_t st);
int launch_grad_tc(const float* e_near, const float* e_far, const float* phi_near, const float* sres, float* gpart, int gsplit,
                   int64_t B, int Ds, int L, cudaStream_t st, float* dbg = nullptr, bool squared = false);
int grad_tc_splits(int64_t B, int Ds, int L, int sms);
bool grad_h_supported(int64_t B, int Ds, int L);
int launch_grad_h(const float* e_near, const float* e_far, const float* phi_near, const float* sres, float* gpart, int gsplit,
                  int64_t B, int Ds, int L, int32_t* scratch, float* out, cudaStream_t st);
bool chain_tc_supported(int Ds);
double chain_tc_bias_estimate(int Ds, int nsites);
int chain_tc_debug(float* out64);
int64_t chain_tc_image_floats(int Ds, int nsteps);
int launch_chain_tc(const float* feat, int fm, int64_t B, int Ds, int L, const float* nodes, int first, int dir, int nsteps,
                    int start_kind, float* img, float* out, cudaStream_t st, const float* in_env = nullptr, float* out_all = nullptr,
                    int64_t out_stride = 0);
int option_value(int option);
bool forward_tc_supported(int Ds, int L, int nd);
int64_t forward_tc_image_floats(int Ds, int L);
int launch_forward_tc(const float* e_near, const float* e_far, const float* phi_near, const float* phi_far, const float* M,
                      float* bimg, float* f_out, int64_t B, int Ds, int L, const int32_t* d_near, const int32_t* d_far,
                      cudaStream_t st, bool have_max = false);
int32_t* forward_tc_maxbits(float* bimg, int Ds, int L);
bool forward_site_tc_supported(int Ds, int L);
int launch_forward_site_tc(const float* e_near, const float* e_far, const float* phi_near, const float* special, int dir, float* bimg,
                           float* f_out, int64_t B, int Ds, int L, const int32_t* d_near, const int32_t* d_far, cudaStream_t st);
int forward_rq(int Ds, int L);
int svd_block_rows(int R, int Cc);

// special_out: where the S V part goes (the special node, or a scratch buffer for the single-site move);
// single: the bond has a unit far site (columns with n = 1 are zero): at most L * d_far singular values are non-zero
int svd_split(const trmps_sweep* s, const trmps_layout& lay, const BondGeom& g, const trmps_opt* opt, cudaStream_t st,
              float* special_out = nullptr, bool single = false);
bool split_factor_supported(int R);
int64_t split_ws_floats(int R);
int split_factor(const trmps_sweep* s, const trmps_layout& lay, const BondGeom& g, const trmps_opt* opt, cudaStream_t st,
                 float* special_out, bool single);
int launch_special_to_bond(const float* special, float* M, float* phi_far, int Ds, int L, int64_t B, int dir, cudaStream_t st);
int launch_site_absorb(const float* tmp, const float* node, float* special, int Ds, int L, int dir, cudaStream_t st);
void set_flat_variant(int v);   // development switch: tile shape of the flat datasource kernel
int allreduce_sum(void* comm, float* buf, int64_t count, cudaStream_t st);

}  // namespace trmps
