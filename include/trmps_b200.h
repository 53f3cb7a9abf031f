/* trmps_b200.h -- C ABI of libtrmps_b200.so: the B200 (sm_100a) replacement for the TensorFlow graph
 * that TrMPS/MPS-MNIST builds for MPS inference and the two-site DMRG sweep.
 *
 * The reference has no FFI: its hot path is a TF-1.2 graph built by Python classes.  Each entry point
 * below names the reference graph-construction code it replaces (paths relative to the reference
 * root).  The Python classes in mps-mnist_b200/trmps/ keep the reference's signatures and call these
 * through ctypes.
 *
 * Rules of the boundary
 *   - plain C types only; every data buffer is a caller-owned DEVICE pointer (the Python side
 *     allocates them as torch CUDA tensors and passes .data_ptr()).
 *   - all work is enqueued asynchronously on the given cudaStream_t (passed as void*); no entry
 *     point synchronises with the host, so a sweep never returns to the host between bonds.
 *   - return 0 on success, negative TRMPS_E_* otherwise; message via trmps_last_error().
 *     Never throws, prints or exits.  Numerical guards of the reference (Armijo exhaustion,
 *     revert-if-cost-rose; baseoptimizer.py:311-315, optimizer.py:260-262) are counted in
 *     iscal[] instead of printed.
 *   - d_feature is fixed at 2 and d_output <= 16 in this build (every reference config on the
 *     path has d=2).
 *
 * Device data layout (floats; Ds = padded bond stride, a multiple of 8, >= every bond dim, >= 2L)
 *   node slot i     nodes + i*2*Ds*Ds      [m][a][b]      logical (d, Dl, Dr)   (mps.py:485-497)
 *   special node    special                [l][m][a][b]   logical (L, d, Dl, Dr)(mps.py:468-482)
 *   dims            int32[N+1]             dims[i] = left bond of site i; dims[0]=dims[N]=2L
 *   environments    envL/envR + i*B*Ds     [t][a]         C1s[i] / C2s[i]       (optimizer.py:49-87)
 *   bond matrices   R x Cc row-major, R = 2*Ds, Cc = 2*L*Ds
 *                   row = m*Ds + i (near feature, near outer bond), col = (l*2+n)*Ds + k
 *                   i.e. the reference bond (L,d,d,Dl,Dr) after transpose [1,3,0,2,4]
 *                   (optimizer.py:283) -- the SVD input layout, used for bond, gradient and delta.
 *   Everything outside the logical extents is kept at zero.
 */
#ifndef TRMPS_B200_H
#define TRMPS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRMPS_ABI_VERSION 1

/* error codes */
#define TRMPS_OK            0
#define TRMPS_E_ARG        (-1)   /* bad argument / unsupported shape */
#define TRMPS_E_CUDA       (-2)   /* CUDA runtime error (message has the cudaError string) */
#define TRMPS_E_NOGPU      (-3)   /* no usable CUDA device */
#define TRMPS_E_NCCL       (-4)   /* NCCL missing or failed */

/* feature maps applied inside the kernels */
#define TRMPS_FM_PRECOMPUTED 0    /* feat is phi, (N,B,2): what the reference model is fed (mps.py:27) */
#define TRMPS_FM_ONE_SQRT3   1    /* feat is pixels (N,B): [1, sqrt(3)(2x-1)]  MNISTpreprocessing.py:55 */
#define TRMPS_FM_ONE_X       2    /* [1, x]  FashionMNISTpreprocessing.py:73 */
#define TRMPS_FM_COS_SIN     3    /* [cos(pi x/2), sin(pi x/2)]  north_star benchmark map */

/* element types of raw image buffers handed to trmps_preprocess_batch */
#define TRMPS_DT_F32 0            /* float32 pixels in [0,1] (what the TF-1.2 MNIST loader returns) */
#define TRMPS_DT_U8  1            /* uint8 pixels as stored in the idx files; scaled by 1/255 in float32 like the loader does */
#define TRMPS_DT_U8_DIV 2         /* uint8 pixels divided by 255 (correctly rounded): FashionMNISTpreprocessing.py:60 */

/* pooling applied by trmps_preprocess_batch (the datasources' `shrink`) */
#define TRMPS_POOL_NONE 0         /* shrink=False */
#define TRMPS_POOL_AVG  1         /* 2x2 average, MNISTpreprocessing.py:44-45 */
#define TRMPS_POOL_MAX  2         /* 2x2 maximum, FashionMNISTpreprocessing.py:61-65 */

/* loss kinds */
#define TRMPS_LOSS_SOFTMAX_XENT 0 /* MPS.cost  mps.py:267-280, residual y - softmax(f) */
#define TRMPS_LOSS_SQUARED      1 /* sqMPS.cost squaredDistanceMPS.py:15-27, residual y - f */

/* indices into the float scalar block (work + layout.scal) */
enum {
  TRMPS_S_COST0 = 0,      /* cost before the bond update                 optimizer.py:242 */
  TRMPS_S_GDC = 1,        /* <G, delta>/B                                 optimizer.py:251-253 */
  TRMPS_S_LR_STAR = 2,    /* accepted step (0 if the bond was kept)       baseoptimizer.py:299-322 */
  TRMPS_S_COST1 = 3,      /* cost of the accepted trial                   optimizer.py:257 */
  TRMPS_S_LR = 4,         /* rate_of_change used for this update */
  TRMPS_S_COUNT = 32
};
/* indices into the int scalar block (iwork) */
enum {
  TRMPS_I_ACCEPT = 0,     /* 1 if the update was kept                     optimizer.py:260-262 */
  TRMPS_I_TRIALS = 1,     /* Armijo trials used (1..19) */
  TRMPS_I_M = 2,          /* bond dimension chosen by the last split      optimizer.py:294-302 */
  TRMPS_I_SVD_SWEEPS = 3, /* Jacobi sweeps used by the last split */
  TRMPS_I_N_REVERT = 4,   /* running count: updates rejected because cost did not drop */
  TRMPS_I_N_EXHAUST = 5,  /* running count: Armijo loops that ran out of trials */
  TRMPS_I_N_UPDATES = 6,  /* running count: bond updates performed */
  TRMPS_I_N_SWEEPS = 7,   /* running count: Jacobi sweeps over all splits */
  TRMPS_I_SVD_ROT = 8,    /* [8 .. 8+64): per-sweep rotation counters of the running SVD */
  TRMPS_I_BARRIER = 80,   /* [80 .. 84): grid-barrier words of the SVD kernel */
  TRMPS_I_ARRIVE = 84,    /* [84 .. 88): self-resetting arrival counters of the kernels whose last CTA finishes a reduction */
  TRMPS_I_COUNT = 96
};

/* MPSOptimizerParameters + max_size (parameterObjects.py:1-49, baseoptimizer.py:17) as the kernels see them */
typedef struct {
  float   lr;                 /* rate_of_change for this training step     baseoptimizer.py:109 */
  float   reg;                /* L2 regularisation                          optimizer.py:249 */
  float   armijo_coeff;       /*                                            baseoptimizer.py:306 */
  float   min_singular_value; /*                                            optimizer.py:294 */
  int32_t max_size;           /*                                            optimizer.py:299 */
  int32_t min_size;           /* 3 in the reference                         optimizer.py:266 */
  int32_t updates_per_step;   /*                                            baseoptimizer.py:324-332 */
  int32_t use_hessian;        /* diagonal Hessian scaling                   optimizer.py:229-237 */
  int32_t max_svd_sweeps;     /* 0 = default (40) */
  int32_t single_site;        /* 1: single-site DMRG updates (singlesiteOptimizer.py:49-138): trmps_bond_step(site, dir)
                                 optimises the special node at `site` alone, splits it and lets node site+dir absorb S V */
  int32_t split_impl;         /* 0: Jacobi on the Cholesky factor of the bond's Gram matrix inside one thread-block cluster
                                 (split_factor.cu; R = 2*Ds <= 256), falling back to 1 for larger bonds;
                                 1: Jacobi on the rows of the bond matrix itself (svd.cu);
                                 2: as 0 with blocks of 16 rows (clusters of at most 8 CTAs: what 0 falls back to when the device
                                 cannot schedule a 16-CTA cluster);
                                 3: as 0 with the Gram-form Jacobi kernel (rotation chain on the CTA's 32 x 32 Gram matrix by
                                 one warp group, rows replayed from registers by another) */
  int32_t reserved;
} trmps_opt;

/* offsets (in floats) of the library-defined regions inside the caller's `work` buffer */
typedef struct {
  int64_t bondM, gradM, deltaM, hessM;   /* R x Cc each (gradM has TRMPS_S_COUNT spare floats behind it) */
  int64_t gpart;                         /* gsplit x R x Cc split-K partials of the gradient */
  int64_t f0, fd;                        /* (B, L) logits of the bond and of delta */
  int64_t fpart;                         /* nsplit x (B, L) column-split partial logits (sized for the FP32 SIMT forward whatever kernel runs) */
  int64_t resid;                         /* (B, L)   y - h */
  int64_t sres;                          /* (B, 2L)  resid[t,l] * phi_far[t,n] */
  int64_t hres;                          /* (B, 2L)  h(1-h)[t,l] * phi_far[t,n]^2 (Hessian only) */
  int64_t phi2;                          /* 2 x (B, 2) feature vectors of the near and far site */
  int64_t Ut;                            /* R x R   rows = left singular vectors (U transposed) */
  int64_t sv;                            /* 3R: singular values (descending), int32 order, int32 row map */
  int64_t red;                           /* per-CTA partial sums of the scalar reductions */
  int64_t scal;                          /* TRMPS_S_COUNT floats */
  int64_t bimg;                          /* hi/lo shared-memory images of the bond for the tensor-core forward, the bond's power-of-two exponent word,
                                            arrival counters and the partial logits of the forward's column-split tail tiles */
  int64_t split_ws;                      /* FP64 Gram partials (added by the cluster Cholesky itself), Gram matrix and FP64 factor (single-SM Cholesky
                                            only), FP32 factor and the rotated factor of the split */
  int64_t env_img;                       /* fp16 hi/lo images of up to N node matrices for the tensor-core environment chain */
  int64_t total;                         /* floats the caller must allocate for `work` */
  int32_t gsplit, nsplit;                /* split-K factor of the gradient, column split of the forward */
  int32_t red_rows, svd_block;           /* rows of `red`; rows per Jacobi block */
  int32_t gsplit_tc, reserved0;          /* split-K factor of the tensor-core gradient kernel */
} trmps_layout;

/* one rank's view of a training batch + model; replaces the placeholders and TensorArrays that
 * BaseOptimizer.__init__ / MPSOptimizer._setup_optimization create (baseoptimizer.py:37-47, optimizer.py:49-87) */
typedef struct {
  int32_t N, d, L, Ds;
  int64_t B;                 /* images held by this rank */
  int64_t B_total;           /* images over all ranks (normalises cost and <G,delta>) */
  int32_t feature_map;       /* TRMPS_FM_* */
  int32_t loss_kind;         /* TRMPS_LOSS_* */
  const float* feat;         /* (N,B) pixels, or (N,B,2) phi when feature_map == PRECOMPUTED */
  const float* labels;       /* (B,L) one-hot float32 */
  float*   nodes;            /* N slots of 2*Ds*Ds */
  float*   special;          /* L*2*Ds*Ds */
  int32_t* dims;             /* N+1 */
  float*   envL;             /* N*B*Ds   C1s */
  float*   envR;             /* N*B*Ds   C2s */
  float*   work;             /* layout.total floats, ZERO-INITIALISED by the caller once (self-resetting arrival counters live in it) */
  int32_t* iwork;            /* TRMPS_I_COUNT ints, zero-initialised once */
  void*    comm;             /* trmps_comm* or NULL when world == 1 */
  int64_t  work_floats;      /* floats allocated behind `work`; every call checks it against the layout it computes */
} trmps_sweep;

/* inference: replaces the graph MPS.predict builds (mps.py:521-560) */
typedef struct {
  int32_t N, d, L, Ds;
  int64_t B;
  int32_t loc;               /* special node location */
  int32_t feature_map;
  int32_t round_output;      /* MPS(round=True): tf.round on the logits (mps.py:558-559) */
  int32_t reserved;
  const float* feat;
  const float* nodes;        /* N slots (slot `loc` unused) */
  const float* special;
  const int32_t* dims;
  float*   logits;           /* (B, L) */
  float*   work;             /* trmps_predict_work_floats() floats */
} trmps_predict_desc;

int         trmps_abi_version(void);
const char* trmps_last_error(void);
int         trmps_device_check(int device);             /* TRMPS_OK if `device` is an sm_100 GPU */
int         trmps_set_option(int32_t option, int32_t value); /* option 0: gradient kernel, 1: forward kernel, 2: inference chain; value 0 = tcgen05 (default for 1 and 2), 1 = FP32 SIMT;
                                                            option 0 value 2 (default): gradient with fp16 hi/lo operands on the fp16 tensor pipe, 0: 3xTF32 operands;
                                                            option 2 value 2: fused tensor-core chain even for chains so long that its accumulated
                                                            rounding bias is estimated above 9e-5 (value 0 then uses the per-site FP32 path);
                                                            option 3: tile shape of the flat datasource kernel (development switch, 0 = default);
                                                            option 4: environment chain of the sweep (initial build and the advance after every
                                                            split): 0 = tcgen05 (the inference chain started from / writing to HBM) while the
                                                            chain's rounding-bias estimate stays below 9e-5, 1 = FP32 SIMT steps;
                                                            option 5: first forward of a two-site update: 1 (default) = single-site form
                                                            (special node x cached environment of the same site, half the contraction; needs the
                                                            environments trmps_sweep_init_env / the sweeps keep), 0 = on the merged bond;
                                                            option 6: fp16 forward contraction by CTA pairs (tcgen05 cta_group::2, each CTA holds half of
                                                            every bond image; the peer's half is announced to the leader by a relayed mbarrier arrive): 0 (default) = one CTA per image
                                                            tile, 1 = CTA pairs (parity-tested alternative: 0.180 against 0.165 ms per cfg3 launch) */
int64_t     trmps_launch_count(void);                   /* kernels launched by this library so far (process-wide) */

/* optional device-side timing of the kernels a sweep launches (CUDA events on the caller's stream; replaces the
 * tf.RunOptions(FULL_TRACE) timeline of baseoptimizer.py:82-86).  Categories: 0 env chain, 1 forward contraction,
 * 2 gradient contraction, 3 SVD split, 4 bond merge, 5 scalar/elementwise, 6 all-reduce. */
#define TRMPS_PROF_NCAT 7
int trmps_debug_chain(float* out64);   /* development aid: phase cycle counters of the last fused-chain launch */
int trmps_profile_begin(int32_t max_records);
int trmps_profile_end(float* ms_per_cat, int32_t* count_per_cat, int32_t ncat);   /* synchronises the recorded events */

/* sizes */
int     trmps_sweep_layout(int32_t N, int64_t B, int32_t L, int32_t Ds, int32_t sm_count, trmps_layout* out);
int64_t trmps_predict_work_floats(int64_t B, int32_t L, int32_t Ds, int32_t N);

/* phi = feature_map(pixels): the datasource-side map, MNISTpreprocessing.py:55.  out (N*B, 2). */
int trmps_feature_map(const float* pixels, int64_t count, int32_t feature_map, float* phi_out, void* stream);

/* Datasource-side preprocessing fused with the batch fetch.  Replaces the per-image sess.run loop
 * of _preprocess_images (MNISTpreprocessing.py:14-78: 2x2 average pool :44-48 when `shrink`, flatten, feature map :55), the
 * Fashion-MNIST variant (FashionMNISTpreprocessing.py:56-77: /255, 2x2 maximum, [1, x]) and
 * the slice + wrap-around + swapaxes(0,1) of MPSDatasource.next_training_data_batch (preprocessing.py:165-208).
 * images: (total, rows*cols) row-major, float32 or uint8 (TRMPS_DT_*); pool: TRMPS_POOL_*.  The batch is
 * images[(start + t) % total], t < B.
 * out_map = TRMPS_FM_PRECOMPUTED: out is (N, B) pooled pixels (for the kernels' fused maps); otherwise out is
 * (N, B, 2) phi of that map.  N = rows*cols/4 with pooling (rows even, cols % 4 == 0, cols <= 64), rows*cols without
 * (any shape; rows*cols % 4 == 0 takes the 16-byte load path). */
int trmps_preprocess_batch(const void* images, int32_t dtype, int64_t total, int32_t rows, int32_t cols, int32_t pool,
                           int64_t start, int64_t B, int32_t out_map, float* out, void* stream);

/* MPSDatasource.next_training_data_batch (preprocessing.py:165-208) on a device-resident dataset in the reference's
 * storage layout: data (total, N, c) with c = 2 (phi) or 1 (pixels), labels (total, L) or NULL.
 * feat_out (N, B, c), labels_out (B, L); rows (start + t) % total. */
int trmps_batch_gather(const float* data, const float* labels, int64_t total, int32_t N, int32_t c, int32_t L,
                       int64_t start, int64_t B, float* feat_out, float* labels_out, void* stream);

/* MPS._lin_reg (mps.py:331-399): the linear pre-training that initialises the MPS -- `iterations` steps of plain gradient
 * descent on a softmax regression (loss_kind SOFTMAX_XENT: mean cross-entropy, mps.py:327-329; SQUARED: 0.5 * sum of squares,
 * squaredDistanceMPS.py:29-30) over phi[..., 1], zero initial model, minibatches of `batch` consecutive samples of a
 * device-resident dataset in the reference's storage layout: data (total, N, 2), labels (total, L), rows
 * (start + iteration * batch + t) % total.  All iterations run inside one launch.  weight: (N, L) row-major, bias: (L). */
int trmps_linreg_pretrain(const float* data, const float* labels, int64_t total, int32_t N, int32_t L, int64_t start,
                          int32_t iterations, int32_t batch, float learning_rate, int32_t loss_kind, float* weight,
                          float* bias, void* stream);

/* one environment step, _chain_multiply_r/_l (mps.py:500-519), _find_C1/_find_C2 (baseoptimizer.py:160-201).
 * dir=+1: out[t,j] = sum_{m,i} phi[t,m] in[t,i] node[m,i,j];  dir=-1: out[t,i] = sum_{m,j} phi[t,m] node[m,i,j] in[t,j].
 * feat_site points at this site's pixels (B) or phi (B,2). din = contracted bond dim. */
int trmps_env_step(const float* feat_site, int32_t feature_map, int64_t B, int32_t Ds, int32_t dir,
                   const float* in_env, const float* node_slot, const int32_t* din_dev, float* out_env, void* stream);

/* MPS.predict (mps.py:521-560) */
int trmps_predict(const trmps_predict_desc* p, void* stream);

/* MPSOptimizer._setup_optimization (optimizer.py:49-87): C1s[0..loc], C2s[loc+1..N-1] */
int trmps_sweep_init_env(const trmps_sweep* s, int32_t loc, void* stream);

/* MPSOptimizer._update_right / _update_left (optimizer.py:144-195 / :89-142): merge, repeated
 * Armijo-guarded gradient update (optimizer.py:239-264, baseoptimizer.py:299-332), truncated SVD split
 * (optimizer.py:266-321), environment advance.  dir=+1 updates bond (site, site+1) and moves the special
 * node to site+1; dir=-1 updates (site-1, site) and moves it to site-1. */
int trmps_bond_step(const trmps_sweep* s, int32_t site, int32_t dir, const trmps_opt* opt, void* stream);

/* BaseOptimizer._sweep_right / _sweep_left (baseoptimizer.py:246-297): bond_step for site = from .. to-dir */
int trmps_sweep_run(const trmps_sweep* s, int32_t from, int32_t to, int32_t dir, const trmps_opt* opt, void* stream);

/* BaseOptimizer.train_step (baseoptimizer.py:203-244): init_env, right loc->N-1, left N-1->0, right 0->loc */
int trmps_train_step(const trmps_sweep* s, int32_t loc, const trmps_opt* opt, void* stream);

/* pieces of bond_step, exposed so the parity tests can check them one by one */
int trmps_bond_merge(const trmps_sweep* s, int32_t site, int32_t dir, void* stream);                 /* optimizer.py:169 / :115 */
int trmps_bond_forward(const trmps_sweep* s, int32_t site, int32_t dir, int32_t use_delta, void* stream); /* optimizer.py:222-227 */
int trmps_bond_gradient(const trmps_sweep* s, int32_t site, int32_t dir, const trmps_opt* opt, void* stream); /* optimizer.py:249-253 */
int trmps_bond_split(const trmps_sweep* s, int32_t site, int32_t dir, const trmps_opt* opt, void* stream);   /* optimizer.py:266-321 */

/* data-parallel exchange (new; the reference is single-process).  One communicator per rank. */
int trmps_comm_unique_id(void* id128);                                   /* rank 0: fills 128 bytes */
int trmps_comm_create(const void* id128, int32_t rank, int32_t world, int32_t device, void** comm_out);
int trmps_comm_destroy(void* comm);

#ifdef __cplusplus
}
#endif
#endif /* TRMPS_B200_H */
