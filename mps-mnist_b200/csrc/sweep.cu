// Host-side orchestration behind the C ABI: buffer layout, one bond update, half-sweeps, a training step.
// Every call only enqueues kernels on the caller's stream; bond dimensions live in device memory
// (`dims`), so the truncated split can change them without the host ever looking.
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace trmps {

// ------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return TRMPS_E_CUDA;
}

// ------------------------------------------------------------------------------------------ counters / profiling
static long long g_launches = 0;
void count_launch() { ++g_launches; }

struct ProfRec { cudaEvent_t a, b; int cat; };
static bool g_prof_on = false;
static ProfRec* g_prof = nullptr;
static int g_prof_n = 0, g_prof_cap = 0;

ProfScope::ProfScope(int c, cudaStream_t s) : cat(c), st(s), on(false) {
  if (!g_prof_on || g_prof_n >= g_prof_cap) return;
  ProfRec& r = g_prof[g_prof_n];
  if (!r.a) {
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  }
  r.cat = c;
  cudaEventRecord(r.a, st);
  on = true;
}
ProfScope::~ProfScope() {
  if (!on) return;
  cudaEventRecord(g_prof[g_prof_n].b, st);
  ++g_prof_n;
}

static int g_grad_impl = 2;   // 2 = tcgen05 fp16 hi/lo (default; shapes it does not cover take 0), 0 = tcgen05 3xTF32, 1 = FP32 SIMT
static int g_fwd_impl = 0;    // same for the forward contraction
static int g_chain_impl = 0;  // 0 = fused tcgen05 chain for inference when Ds <= 32, 1 = per-site FP32 SIMT steps
static int g_fwd_site = 1;    // first forward of a two-site update in its single-site form (needs the cached environment of the same site: trmps_sweep_init_env / sweeps provide it)
static int g_env_impl = 0;    // environments of the sweep: 0 = tcgen05 chain from / to HBM, 1 = per-site FP32 SIMT steps
static int g_fwd_pair = 0;    // fp16 forward by CTA pairs (tcgen05 cta_group::2): parity-tested, measured slower than one CTA per tile
int option_value(int option) { return option == 0 ? g_grad_impl : option == 1 ? g_fwd_impl : option == 6 ? g_fwd_pair : g_chain_impl; }

// SM count of the CURRENT device (ranks above 0 of a multi-GPU process group run on other devices; the attribute query
// is a host-side table lookup)
static int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) return n;
  return 148;
}

// ------------------------------------------------------------------------------------------ layout
static int64_t a4(int64_t x) { return (x + 3) / 4 * 4; }

static int compute_layout(int32_t N, int64_t B, int32_t L, int32_t Ds, int32_t sms, trmps_layout* o) {
  if (B < 0 || L < 1 || L > TRMPS_LMAX || Ds < 2 * L || Ds % 8 != 0) {
    set_error("layout: unsupported shape (need 1<=L<=16, Ds%%8==0, Ds>=2L); got L=%d Ds=%d", L, Ds);
    return TRMPS_E_ARG;
  }
  if (sms <= 0) sms = 148;
  const int64_t R = 2 * (int64_t)Ds, Cc = 2 * (int64_t)L * Ds, RC = R * Cc;
  int rq = forward_rq(Ds, L);
  if (rq == 0) {
    set_error("layout: bond stride Ds=%d too large for the forward kernel's shared memory", Ds);
    return TRMPS_E_ARG;
  }
  // split-K of the gradient: one full wave of (2 CTAs per SM)
  int64_t tiles = ((R + 127) / 128) * ((Cc + 127) / 128);
  int64_t gs = (2 * (int64_t)sms) / tiles;
  if (gs < 1) gs = 1;
  int64_t cap = B / 256;
  if (cap < 1) cap = 1;
  if (gs > cap) gs = cap;
  if (gs > 64) gs = 64;
  // column split of the forward: smallest split whose last wave is >= 90% full, else the best one
  int64_t ctas_x = (B + 64 * rq - 1) / (64 * rq);
  if (ctas_x < 1) ctas_x = 1;
  int64_t ntiles = (Cc + 127) / 128;
  int best = 1;
  double best_eff = 0.0;
  for (int ns = 1; ns <= 8 && ns <= ntiles; ++ns) {
    int64_t total = ctas_x * ns;
    int64_t waves = (total + sms - 1) / sms;
    double eff = (double)total / (double)(waves * sms);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = ns; }
    if (eff >= 0.9) { best = ns; break; }
  }
  int64_t off = 0;
  o->bondM = off; off += a4(RC);
  o->gradM = off; off += a4(RC + TRMPS_S_COUNT);
  o->deltaM = off; off += a4(RC);
  o->hessM = off; off += a4(RC + TRMPS_S_COUNT);
  const int gs_tc = grad_tc_splits(B, Ds, L, sms);
  o->gpart = off; off += a4((gs > gs_tc ? gs : gs_tc) * RC);
  o->f0 = off; off += a4(B * L);
  o->fd = off; off += a4(B * L);
  o->fpart = off; off += a4((int64_t)best * B * L);
  o->resid = off; off += a4(B * L);
  o->sres = off; off += a4(B * 2 * L);
  o->hres = off; off += a4(B * 2 * L);
  o->phi2 = off; off += a4(4 * B);
  o->Ut = off; off += a4(R * R);
  o->sv = off; off += a4(3 * R);
  int64_t rr = (B + TRMPS_IMG_BLOCK - 1) / TRMPS_IMG_BLOCK;      // one row per CTA of the per-image kernels
  if (rr < 296) rr = 296;
  o->red = off; off += a4(rr * 32);
  o->scal = off; off += a4(2 * TRMPS_S_COUNT);
  o->bimg = off; off += a4(forward_tc_supported(Ds, L, 2) ? forward_tc_image_floats(Ds, L) : 0);
  o->split_ws = off; off += a4(split_factor_supported((int)R) ? split_ws_floats((int)R) : 0);
  o->env_img = off; off += a4(chain_tc_supported(Ds) ? chain_tc_image_floats(Ds, N > 1 ? N : 1) : 0);
  o->total = off;
  o->gsplit = (int32_t)gs;
  o->nsplit = best;
  o->red_rows = (int32_t)rr;
  o->svd_block = svd_block_rows((int)R, (int)Cc);
  o->gsplit_tc = gs_tc;
  o->reserved0 = 0;
  return TRMPS_OK;
}

BondGeom bond_geom(const trmps_sweep* s, int site, int dir) {
  BondGeom g;
  const int64_t stride = s->B * (int64_t)s->Ds;
  g.dir = dir;
  g.near_site = site;
  if (dir > 0) {
    g.far_site = site + 1;
    g.e_near = s->envL + stride * site;
    g.e_far = s->envR + stride * (site + 1);
    g.d_near = s->dims + site;
    g.d_mid = s->dims + site + 1;
    g.d_far = s->dims + site + 2;
  } else {
    g.far_site = site - 1;
    g.e_near = s->envR + stride * site;
    g.e_far = s->envL + stride * (site - 1);
    g.d_near = s->dims + site + 1;
    g.d_mid = s->dims + site;
    g.d_far = s->dims + site - 1;
  }
  return g;
}

// column split of the forward actually in use: the tensor-core forward writes complete logits (one part); the layout
// always reserves the parts of the FP32 SIMT kernel, so switching kernels with trmps_set_option never moves a buffer
static int eff_nsplit(const trmps_sweep* s, const trmps_layout& lay) {
  return (g_fwd_impl != 1 && forward_tc_supported(s->Ds, s->L, 2)) ? 1 : lay.nsplit;
}

static int check_sweep(const trmps_sweep* s, const char* who) {
  if (!s || !s->feat || !s->labels || !s->nodes || !s->special || !s->dims || !s->envL || !s->envR || !s->work || !s->iwork) {
    set_error("%s: null pointer in trmps_sweep", who);
    return TRMPS_E_ARG;
  }
  if (s->d != 2 || s->L < 1 || s->L > TRMPS_LMAX || s->Ds < 2 * s->L || s->Ds % 8 != 0 || s->N < 2 || s->B < 1 || s->B_total < s->B) {
    set_error("%s: unsupported shape (need d=2, 1<=L<=16, Ds%%8==0, Ds>=2L, N>=2, B>=1)", who);
    return TRMPS_E_ARG;
  }
  return TRMPS_OK;
}

static int check_bond(const trmps_sweep* s, int site, int dir, const char* who) {
  int rc = check_sweep(s, who);
  if (rc) return rc;
  if ((dir != 1 && dir != -1) || site < 0 || site >= s->N || site + dir < 0 || site + dir >= s->N) {
    set_error("%s: bond (site=%d, dir=%d) outside the chain of %d sites", who, site, dir, s->N);
    return TRMPS_E_ARG;
  }
  return TRMPS_OK;
}

// single-site update at `site` (singlesiteOptimizer.py:49-138): the "bond" is the special node alone; the far
// environment is the one on the other side of the same site and the far site is the identity
static BondGeom site_geom(const trmps_sweep* s, int site, int dir) {
  BondGeom g;
  const int64_t stride = s->B * (int64_t)s->Ds;
  g.dir = dir;
  g.near_site = site;
  g.far_site = site + dir;            // the node that absorbs S V
  if (dir > 0) {
    g.e_near = s->envL + stride * site;
    g.e_far = s->envR + stride * site;
    g.d_near = s->dims + site;
    g.d_mid = s->dims + site + 1;
    g.d_far = s->dims + site + 1;
  } else {
    g.e_near = s->envR + stride * site;
    g.e_far = s->envL + stride * site;
    g.d_near = s->dims + site + 1;
    g.d_mid = s->dims + site;
    g.d_far = s->dims + site;
  }
  return g;
}

static int phi_pair(const trmps_sweep* s, const trmps_layout& lay, const BondGeom& g, cudaStream_t st) {
  float* phi2 = s->work + lay.phi2;
  return launch_phi_pair(s->feat + feat_site_offset(s->feature_map, s->B, g.near_site),
                         s->feat + feat_site_offset(s->feature_map, s->B, g.far_site), s->feature_map, s->B, phi2, phi2 + 2 * s->B, st);
}

static int do_merge(const trmps_sweep* s, const trmps_layout& lay, const BondGeom& g, cudaStream_t st) {
  ProfScope ps(PROF_MERGE, st);
  return launch_merge(s->special, s->nodes + (size_t)g.far_site * 2 * s->Ds * s->Ds, s->work + lay.bondM, s->Ds, s->L, g.dir,
                      g.d_mid, st);
}

static int do_forward(const trmps_sweep* s, const trmps_layout& lay, const BondGeom& g, const float* M, cudaStream_t st,
                      bool have_max = false) {
  ProfScope ps(PROF_FORWARD, st);
  const float* phi2 = s->work + lay.phi2;
  if (g_fwd_impl != 1 && forward_tc_supported(s->Ds, s->L, 2))
    return launch_forward_tc(g.e_near, g.e_far, phi2, phi2 + 2 * s->B, M, s->work + lay.bimg, s->work + lay.fpart, s->B, s->Ds,
                             s->L, g.d_near, g.d_far, st, have_max);
  return launch_forward(g.e_near, g.e_far, phi2, phi2 + 2 * s->B, M, s->work + lay.fpart, lay.nsplit, s->B, s->Ds, s->L, 2,
                        g.d_near, st);
}

static float cost_norm(const trmps_sweep* s) {
  return s->loss_kind == TRMPS_LOSS_SOFTMAX_XENT ? 1.0f / (float)s->B_total : 0.5f / ((float)s->B_total * (float)s->L);
}

// loss + gradient (+Hessian) + all-reduce + finish; leaves G in gradM, delta in deltaM (Hessian) or gradM
static int do_gradient(const trmps_sweep* s, const trmps_layout& lay, const BondGeom& g, const trmps_opt* opt, int from_parts,
                       cudaStream_t st, int32_t* delta_maxbits) {
  static const bool grad_debug = getenv("TRMPS_GRAD_DEBUG") != nullptr;      // development: phase timers of grad_tc in hres
  const int Ds = s->Ds, L = s->L;
  const int64_t RC = (int64_t)2 * Ds * 2 * L * Ds;
  float* w = s->work;
  const bool hess = opt->use_hessian && s->loss_kind == TRMPS_LOSS_SOFTMAX_XENT;   // baseoptimizer.py:32-35
  LossArgs la;
  la.fpart = w + lay.fpart; la.nsplit = eff_nsplit(s, lay); la.f0 = w + lay.f0; la.labels = s->labels;
  la.phi_far = w + lay.phi2 + 2 * s->B; la.resid = w + lay.resid; la.sres = w + lay.sres;
  la.hres = hess ? w + lay.hres : nullptr; la.red = w + lay.red; la.B = s->B; la.L = L; la.loss_kind = s->loss_kind;
  la.from_parts = from_parts;
  int rc;
  float* tail = w + lay.gradM + RC;
  // the un-normalised cost sum goes behind the gradient (one all-reduce carries both); the last CTA of the loss kernel writes it
  la.arrive = s->iwork + TRMPS_I_ARRIVE; la.cost_sum = tail; la.clear_word = delta_maxbits;
  {
    ProfScope ps(PROF_SCALAR, st);
    if ((rc = launch_loss(la, st))) return rc;
  }
  const bool one_rank = s->comm == nullptr && !hess;      // (the Hessian pass reuses the split-K buffer: sum the gradient's parts first)
  const float* parts = nullptr;             // single rank: the split-K parts are summed by grad_finish_kernel itself
  int nparts = 0;
  {
    ProfScope ps(PROF_GRAD, st);
    if (g_grad_impl == 2 && grad_h_supported(s->B, Ds, L)) {
      // fp16 hi/lo operands (option 2; parity-green but not yet faster than 3xTF32: its producers' global loads are
      // single-buffered); the per-image exponents live in the (still unused) delta buffer
      if ((rc = launch_grad_h(g.e_near, g.e_far, w + lay.phi2, w + lay.sres, w + lay.gpart, lay.gsplit_tc, s->B, Ds, L,
                              reinterpret_cast<int32_t*>(w + lay.deltaM), w + lay.gradM, st)))
        return rc;
    } else if (g_grad_impl != 1) {
      if ((rc = launch_grad_tc(g.e_near, g.e_far, w + lay.phi2, w + lay.sres, w + lay.gpart, lay.gsplit_tc, s->B, Ds, L, st, grad_debug ? w + lay.hres : nullptr))) return rc;
      if (one_rank) { parts = w + lay.gpart; nparts = lay.gsplit_tc; }
      else if ((rc = launch_sum_parts(w + lay.gpart, lay.gsplit_tc, RC, w + lay.gradM, st))) return rc;
    } else {
      if ((rc = launch_grad(g.e_near, g.e_far, w + lay.phi2, w + lay.sres, w + lay.gpart, lay.gsplit, s->B, Ds, L, false, st))) return rc;
      if (one_rank) { parts = w + lay.gpart; nparts = lay.gsplit; }
      else if ((rc = launch_sum_parts(w + lay.gpart, lay.gsplit, RC, w + lay.gradM, st))) return rc;
    }
  }
  {
    ProfScope ps(PROF_COMM, st);
    if ((rc = allreduce_sum(s->comm, w + lay.gradM, RC + 4, st))) return rc;
  }
  if (hess) {
    // diagonal Hessian: the gradient contraction with squared operands (optimizer.py:229-237), on the same tensor-core kernel
    ProfScope ps(PROF_GRAD, st);
    if (g_grad_impl != 1) {
      if ((rc = launch_grad_tc(g.e_near, g.e_far, w + lay.phi2, w + lay.hres, w + lay.gpart, lay.gsplit_tc, s->B, Ds, L, st, nullptr, true))) return rc;
      if ((rc = launch_sum_parts(w + lay.gpart, lay.gsplit_tc, RC, w + lay.hessM, st))) return rc;
    } else {
      if ((rc = launch_grad(g.e_near, g.e_far, w + lay.phi2, w + lay.hres, w + lay.gpart, lay.gsplit, s->B, Ds, L, true, st))) return rc;
      if ((rc = launch_sum_parts(w + lay.gpart, lay.gsplit, RC, w + lay.hessM, st))) return rc;
    }
    if ((rc = allreduce_sum(s->comm, w + lay.hessM, RC, st))) return rc;
  }
  float* scal = w + lay.scal;
  ProfScope ps(PROF_SCALAR, st);
  // regulariser, delta, <G, delta> / B, the normalised cost and the exponent of delta for the next forward: one launch
  return launch_grad_finish(w + lay.gradM, w + lay.bondM, hess ? w + lay.hessM : nullptr, w + lay.deltaM, parts, nparts, RC, opt->reg,
                            2.f * opt->reg, 1.0f / (float)s->B_total, tail, cost_norm(s), w + lay.red, scal,
                            s->iwork + TRMPS_I_ARRIVE + 1, delta_maxbits, 296, st);
}

// gradient step(s) with Armijo backtracking and accept/revert on the bond matrix in bondM (optimizer.py:239-264,
// baseoptimizer.py:299-332); shared by the two-site and the single-site update
static int do_update_bond(const trmps_sweep* s, const trmps_layout& lay, const BondGeom& g, const trmps_opt* opt, cudaStream_t st,
                          bool site_form = false) {
  const int L = s->L;
  const int64_t RC = (int64_t)2 * s->Ds * 2 * L * s->Ds;
  float* w = s->work;
  float* scal = w + lay.scal;
  const bool hess = opt->use_hessian && s->loss_kind == TRMPS_LOSS_SOFTMAX_XENT;
  const float* delta = hess ? w + lay.deltaM : w + lay.gradM;
  int rc;
  if (site_form) {
    // the bond is still the product of its two sites: its logits are those of the special node alone with the cached
    // environment on the other side of the same site -- half the contraction (2 B L d Dl Dm instead of 2 B L d^2 Dl Dr flops)
    ProfScope ps(PROF_FORWARD, st);
    const int64_t stride = s->B * (int64_t)s->Ds;
    const float* e_same = g.dir > 0 ? s->envR + stride * g.near_site : s->envL + stride * g.near_site;
    if ((rc = launch_forward_site_tc(g.e_near, e_same, w + lay.phi2, s->special, g.dir, w + lay.bimg, w + lay.fpart, s->B, s->Ds, L,
                                     g.d_near, g.d_mid, st)))
      return rc;
  } else if ((rc = do_forward(s, lay, g, w + lay.bondM, st))) return rc;
  const int nupd = opt->updates_per_step > 0 ? opt->updates_per_step : 0;
  // the fp16 forward scales delta by one power of two: grad_finish_kernel finds it while it writes delta
  int32_t* delta_maxbits = forward_tc_maxbits(w + lay.bimg, s->Ds, L);
  for (int u = 0; u < nupd; ++u) {                                  // baseoptimizer.py:324-332
    if ((rc = do_gradient(s, lay, g, opt, u == 0, st, delta_maxbits))) return rc;
    if ((rc = do_forward(s, lay, g, delta, st, delta_maxbits != nullptr))) return rc;         // f(delta): serves every Armijo trial
    ArmijoArgs aa;
    aa.f0 = w + lay.f0; aa.fdpart = w + lay.fpart; aa.nsplit = eff_nsplit(s, lay); aa.fd = w + lay.fd; aa.labels = s->labels;
    aa.red = w + lay.red; aa.B = s->B; aa.L = L; aa.loss_kind = s->loss_kind; aa.lr = opt->lr;
    // the 19 trial sums by the kernel's last CTA; on a single rank the accept / revert decision too
    aa.arrive = s->iwork + TRMPS_I_ARRIVE + 2; aa.trial = scal + 8; aa.trial_scale = cost_norm(s);
    aa.decide = s->comm == nullptr; aa.scal = scal; aa.iscal = s->iwork; aa.armijo_coeff = opt->armijo_coeff;
    ProfScope ps(PROF_SCALAR, st);
    if ((rc = launch_armijo(aa, st))) return rc;
    if (!aa.decide) {
      if ((rc = allreduce_sum(s->comm, scal + 8, 20, st))) return rc;
      if ((rc = launch_decide(scal, scal + 8, s->iwork, opt->lr, opt->armijo_coeff, st))) return rc;
    }
    if ((rc = launch_apply_update(w + lay.bondM, delta, RC, w + lay.f0, w + lay.fd, s->B * L, scal, st))) return rc;
  }
  return TRMPS_OK;
}

// the sweep's environments run on the tensor cores (the inference chain started from / writing to HBM) while the rounding
// bias of the fused chain over the whole MPS stays inside the tolerance (chain_tc.cu, "Order of accumulation")
static bool env_on_tensor_cores(const trmps_sweep* s) {
  return g_env_impl == 0 && chain_tc_supported(s->Ds) && chain_tc_bias_estimate(s->Ds, s->N) <= 9e-5;
}

// environment advance over the site that just received its isometry (optimizer.py:187-190 / :132-135)
static int do_env_advance(const trmps_sweep* s, const trmps_layout& lay, int site, int dir, cudaStream_t st) {
  ProfScope ps_env(PROF_ENV, st);
  const int Ds = s->Ds;
  const int64_t stride = s->B * (int64_t)Ds;
  const float* in_env = dir > 0 ? s->envL + stride * site : s->envR + stride * site;
  float* out_env = dir > 0 ? s->envL + stride * (site + 1) : s->envR + stride * (site - 1);
  if (env_on_tensor_cores(s))
    return launch_chain_tc(s->feat, s->feature_map, s->B, Ds, s->L, s->nodes, site, dir, 1, 2, s->work + lay.env_img, out_env, st, in_env);
  const float* feat_site = s->feat + feat_site_offset(s->feature_map, s->B, site);
  const float* node = s->nodes + (size_t)site * 2 * Ds * Ds;
  return launch_env_step(feat_site, s->feature_map, s->B, Ds, dir, in_env, node, dir > 0 ? s->dims + site : s->dims + site + 1, out_env, st);
}

static int do_bond_step(const trmps_sweep* s, const trmps_layout& lay, int site, int dir, const trmps_opt* opt, cudaStream_t st) {
  const BondGeom g = bond_geom(s, site, dir);
  int rc;
  if ((rc = phi_pair(s, lay, g, st))) return rc;
  if ((rc = do_merge(s, lay, g, st))) return rc;
  const bool site_form = g_fwd_site != 0 && g_fwd_impl == 0 && forward_site_tc_supported(s->Ds, s->L);
  if ((rc = do_update_bond(s, lay, g, opt, st, site_form))) return rc;
  {
    ProfScope ps(PROF_SVD, st);
    if ((rc = svd_split(s, lay, g, opt, st))) return rc;
  }
  return do_env_advance(s, lay, site, dir, st);
}

// single-site update (singlesiteOptimizer.py:49-138): optimise the special node at `site`, split it into the
// isometry that stays and S V, let node site+dir absorb S V (it becomes the special node), advance the environment
static int do_site_step(const trmps_sweep* s, const trmps_layout& lay, int site, int dir, const trmps_opt* opt, cudaStream_t st) {
  const BondGeom g = site_geom(s, site, dir);
  const int Ds = s->Ds, L = s->L;
  float* w = s->work;
  if (opt->use_hessian) {
    set_error("single-site update: the Hessian variant is not implemented (nor is it in the reference, singlesiteOptimizer.py:161-172)");
    return TRMPS_E_ARG;
  }
  int rc;
  float* phi2 = w + lay.phi2;
  if ((rc = launch_phi_site(s->feat + feat_site_offset(s->feature_map, s->B, site), s->feature_map, s->B, phi2, st))) return rc;
  {
    ProfScope ps(PROF_MERGE, st);
    if ((rc = launch_special_to_bond(s->special, w + lay.bondM, phi2 + 2 * s->B, Ds, L, s->B, dir, st))) return rc;
  }
  if ((rc = do_update_bond(s, lay, g, opt, st))) return rc;
  float* tmp = w + lay.hessM;          // scratch for S V in special layout (the Hessian is off)
  {
    ProfScope ps(PROF_SVD, st);
    if ((rc = svd_split(s, lay, g, opt, st, tmp, true))) return rc;
  }
  {
    ProfScope ps(PROF_MERGE, st);
    if ((rc = launch_site_absorb(tmp, s->nodes + (size_t)(site + dir) * 2 * Ds * Ds, s->special, Ds, L, dir, st))) return rc;
  }
  return do_env_advance(s, lay, site, dir, st);
}

static int do_step(const trmps_sweep* s, const trmps_layout& lay, int site, int dir, const trmps_opt* opt, cudaStream_t st) {
  return opt->single_site ? do_site_step(s, lay, site, dir, opt, st) : do_bond_step(s, lay, site, dir, opt, st);
}

static int do_init_env(const trmps_sweep* s, const trmps_layout& lay, int loc, cudaStream_t st) {
  ProfScope ps(PROF_ENV, st);
  const int Ds = s->Ds, N = s->N;
  const int64_t stride = s->B * (int64_t)Ds;
  const size_t slot = (size_t)2 * Ds * Ds;
  int rc;
  if ((rc = launch_fill_boundary(s->envL, s->B, Ds, s->L, 0, st))) return rc;
  if ((rc = launch_fill_boundary(s->envR + stride * (N - 1), s->B, Ds, s->L, 1, st))) return rc;
  if (env_on_tensor_cores(s)) {
    // one fused chain per side: the state stays on chip, every site's environment is written to its cache slot on the way
    float* img = s->work + lay.env_img;
    if (loc > 0 && (rc = launch_chain_tc(s->feat, s->feature_map, s->B, Ds, s->L, s->nodes, 0, +1, loc, 0, img, s->envL + stride * loc, st,
                                         nullptr, s->envL + stride, stride)))
      return rc;
    // down to envR[loc]: the two-site sweep needs envR[loc+1], the single-site one (singlesiteOptimizer.py:38-47) envR[loc]
    if (N - 1 - loc > 0 && (rc = launch_chain_tc(s->feat, s->feature_map, s->B, Ds, s->L, s->nodes, N - 1, -1, N - 1 - loc, 1, img,
                                                 s->envR + stride * loc, st, nullptr, s->envR + stride * (N - 2), -stride)))
      return rc;
    return TRMPS_OK;
  }
  for (int c = 0; c < loc; ++c) {
    rc = launch_env_step(s->feat + feat_site_offset(s->feature_map, s->B, c), s->feature_map, s->B, Ds, +1, s->envL + stride * c,
                         s->nodes + slot * c, s->dims + c, s->envL + stride * (c + 1), st);
    if (rc) return rc;
  }
  for (int c = N - 1; c > loc; --c) {
    rc = launch_env_step(s->feat + feat_site_offset(s->feature_map, s->B, c), s->feature_map, s->B, Ds, -1, s->envR + stride * c,
                         s->nodes + slot * c, s->dims + c + 1, s->envR + stride * (c - 1), st);
    if (rc) return rc;
  }
  return TRMPS_OK;
}

// ------------------------------------------------------------------------------------------ NCCL (dlopen)
typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_init_rank)(void**, int, nccl_uid, int);
typedef int (*fn_destroy)(void*);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*fn_errstr)(int);

struct Nccl {
  void* h = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_destroy destroy = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_errstr errstr = nullptr;
};
static Nccl g_nccl;

static int nccl_load() {
  if (g_nccl.h) return TRMPS_OK;
  const char* env = getenv("TRMPS_NCCL_LIB");
  void* h = env ? dlopen(env, RTLD_NOW | RTLD_GLOBAL) : nullptr;
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_error("NCCL: cannot dlopen libnccl.so.2 (%s); set TRMPS_NCCL_LIB", dlerror());
    return TRMPS_E_NCCL;
  }
  g_nccl.get_uid = (fn_get_uid)dlsym(h, "ncclGetUniqueId");
  g_nccl.init_rank = (fn_init_rank)dlsym(h, "ncclCommInitRank");
  g_nccl.destroy = (fn_destroy)dlsym(h, "ncclCommDestroy");
  g_nccl.allreduce = (fn_allreduce)dlsym(h, "ncclAllReduce");
  g_nccl.errstr = (fn_errstr)dlsym(h, "ncclGetErrorString");
  if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.destroy || !g_nccl.allreduce) {
    set_error("NCCL: missing symbols in libnccl.so.2");
    return TRMPS_E_NCCL;
  }
  g_nccl.h = h;
  return TRMPS_OK;
}

static int nccl_fail(int code, const char* what) {
  set_error("NCCL error %d (%s) at %s", code, g_nccl.errstr ? g_nccl.errstr(code) : "?", what);
  return TRMPS_E_NCCL;
}

int allreduce_sum(void* comm, float* buf, int64_t count, cudaStream_t st) {
  if (!comm) return TRMPS_OK;
  int rc = g_nccl.allreduce(buf, buf, (size_t)count, /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm, st);
  if (rc != 0) return nccl_fail(rc, "ncclAllReduce");
  return TRMPS_OK;
}

}  // namespace trmps

using namespace trmps;

// ================================================================================================= C ABI
extern "C" int trmps_abi_version(void) { return TRMPS_ABI_VERSION; }
extern "C" int64_t trmps_launch_count(void) { return (int64_t)g_launches; }

extern "C" int trmps_set_option(int32_t option, int32_t value) {
  if (option == 0 && (value == 0 || value == 1 || value == 2)) { g_grad_impl = value; return TRMPS_OK; }
  if (option == 1 && (value == 0 || value == 1 || value == 2)) { g_fwd_impl = value; return TRMPS_OK; }
  if (option == 2 && (value == 0 || value == 1 || value == 2)) { g_chain_impl = value; return TRMPS_OK; }
  if (option == 3 && value >= 0 && value <= 3) { set_flat_variant(value); return TRMPS_OK; }
  if (option == 4 && (value == 0 || value == 1)) { g_env_impl = value; return TRMPS_OK; }
  if (option == 5 && (value == 0 || value == 1)) { g_fwd_site = value; return TRMPS_OK; }
  if (option == 6 && (value == 0 || value == 1)) { g_fwd_pair = value; return TRMPS_OK; }
  set_error("trmps_set_option: unknown option %d / value %d", option, value);
  return TRMPS_E_ARG;
}

extern "C" int trmps_profile_begin(int32_t max_records) {
  if (max_records <= 0) { set_error("trmps_profile_begin: max_records must be positive"); return TRMPS_E_ARG; }
  if (max_records > g_prof_cap) {
    ProfRec* n = (ProfRec*)calloc((size_t)max_records, sizeof(ProfRec));
    if (!n) { set_error("trmps_profile_begin: out of memory"); return TRMPS_E_ARG; }
    if (g_prof) { memcpy(n, g_prof, sizeof(ProfRec) * g_prof_cap); free(g_prof); }
    g_prof = n;
    g_prof_cap = max_records;
  }
  g_prof_n = 0;
  g_prof_on = true;
  return TRMPS_OK;
}

extern "C" int trmps_profile_end(float* ms_per_cat, int32_t* count_per_cat, int32_t ncat) {
  g_prof_on = false;
  if (!ms_per_cat || !count_per_cat || ncat < PROF_NCAT) { set_error("trmps_profile_end: need %d categories", (int)PROF_NCAT); return TRMPS_E_ARG; }
  for (int i = 0; i < ncat; ++i) { ms_per_cat[i] = 0.f; count_per_cat[i] = 0; }
  for (int i = 0; i < g_prof_n; ++i) {
    TRMPS_CUDA_OK(cudaEventSynchronize(g_prof[i].b));
    float ms = 0.f;
    TRMPS_CUDA_OK(cudaEventElapsedTime(&ms, g_prof[i].a, g_prof[i].b));
    ms_per_cat[g_prof[i].cat] += ms;
    count_per_cat[g_prof[i].cat] += 1;
  }
  g_prof_n = 0;
  return TRMPS_OK;
}
extern "C" const char* trmps_last_error(void) { return g_err; }

extern "C" int trmps_device_check(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    set_error("no CUDA device (%s)", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return TRMPS_E_NOGPU;
  }
  if (device < 0 || device >= n) {
    set_error("device %d out of range (0..%d)", device, n - 1);
    return TRMPS_E_ARG;
  }
  int major = 0;
  TRMPS_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) {
    set_error("device %d has compute capability %d.x; this library is built for sm_100a only", device, major);
    return TRMPS_E_NOGPU;
  }
  return TRMPS_OK;
}

extern "C" int trmps_sweep_layout(int32_t N, int64_t B, int32_t L, int32_t Ds, int32_t sms, trmps_layout* out) {
  if (!out) {
    set_error("trmps_sweep_layout: null output");
    return TRMPS_E_ARG;
  }
  memset(out, 0, sizeof(*out));
  return compute_layout(N, B, L, Ds, sms, out);
}

#define GET_LAYOUT(s, who)                                                          \
  trmps_layout lay;                                                                 \
  {                                                                                 \
    int _rc = compute_layout((s)->N, (s)->B, (s)->L, (s)->Ds, sm_count(), &lay);    \
    if (_rc) return _rc;                                                            \
    if ((s)->work_floats < lay.total) {                                             \
      set_error("%s: work buffer holds %lld floats, the layout needs %lld (trmps_sweep_layout)", who, (long long)(s)->work_floats, (long long)lay.total); \
      return TRMPS_E_ARG;                                                           \
    }                                                                               \
  }

extern "C" int trmps_sweep_init_env(const trmps_sweep* s, int32_t loc, void* stream) {
  int rc = check_sweep(s, "trmps_sweep_init_env");
  if (rc) return rc;
  if (loc < 0 || loc >= s->N) {
    set_error("trmps_sweep_init_env: loc=%d outside 0..%d", loc, s->N - 1);
    return TRMPS_E_ARG;
  }
  GET_LAYOUT(s, "trmps_sweep_init_env");
  return do_init_env(s, lay, loc, (cudaStream_t)stream);
}

extern "C" int trmps_bond_step(const trmps_sweep* s, int32_t site, int32_t dir, const trmps_opt* opt, void* stream) {
  int rc = check_bond(s, site, dir, "trmps_bond_step");
  if (rc) return rc;
  if (!opt) { set_error("trmps_bond_step: null opt"); return TRMPS_E_ARG; }
  GET_LAYOUT(s, "trmps_bond_step");
  return do_step(s, lay, site, dir, opt, (cudaStream_t)stream);
}

extern "C" int trmps_sweep_run(const trmps_sweep* s, int32_t from, int32_t to, int32_t dir, const trmps_opt* opt, void* stream) {
  int rc = check_sweep(s, "trmps_sweep_run");
  if (rc) return rc;
  if (!opt || (dir != 1 && dir != -1) || from < 0 || from >= s->N || to < 0 || to >= s->N || (to - from) * dir < 0) {
    set_error("trmps_sweep_run: bad range from=%d to=%d dir=%d", from, to, dir);
    return TRMPS_E_ARG;
  }
  GET_LAYOUT(s, "trmps_sweep_run");
  for (int c = from; c != to; c += dir) {
    rc = do_step(s, lay, c, dir, opt, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return TRMPS_OK;
}

extern "C" int trmps_train_step(const trmps_sweep* s, int32_t loc, const trmps_opt* opt, void* stream) {
  int rc = trmps_sweep_init_env(s, loc, stream);
  if (rc) return rc;
  if ((rc = trmps_sweep_run(s, loc, s->N - 1, +1, opt, stream))) return rc;    // baseoptimizer.py:219
  if ((rc = trmps_sweep_run(s, s->N - 1, 0, -1, opt, stream))) return rc;      // baseoptimizer.py:225
  return trmps_sweep_run(s, 0, loc, +1, opt, stream);                          // baseoptimizer.py:236
}

extern "C" int trmps_bond_merge(const trmps_sweep* s, int32_t site, int32_t dir, void* stream) {
  int rc = check_bond(s, site, dir, "trmps_bond_merge");
  if (rc) return rc;
  GET_LAYOUT(s, "trmps_bond_merge");
  const BondGeom g = bond_geom(s, site, dir);
  if ((rc = phi_pair(s, lay, g, (cudaStream_t)stream))) return rc;
  return do_merge(s, lay, g, (cudaStream_t)stream);
}

extern "C" int trmps_bond_forward(const trmps_sweep* s, int32_t site, int32_t dir, int32_t use_delta, void* stream) {
  int rc = check_bond(s, site, dir, "trmps_bond_forward");
  if (rc) return rc;
  GET_LAYOUT(s, "trmps_bond_forward");
  const BondGeom g = bond_geom(s, site, dir);
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = phi_pair(s, lay, g, st))) return rc;
  const float* M = s->work + (use_delta == 2 ? lay.deltaM : use_delta == 1 ? lay.gradM : lay.bondM);
  if ((rc = do_forward(s, lay, g, M, st))) return rc;
  float* dst = s->work + (use_delta ? lay.fd : lay.f0);
  return launch_sum_round_logits(s->work + lay.fpart, eff_nsplit(s, lay), s->B, s->L, 0, dst, st);
}

extern "C" int trmps_bond_gradient(const trmps_sweep* s, int32_t site, int32_t dir, const trmps_opt* opt, void* stream) {
  int rc = check_bond(s, site, dir, "trmps_bond_gradient");
  if (rc) return rc;
  if (!opt) { set_error("trmps_bond_gradient: null opt"); return TRMPS_E_ARG; }
  GET_LAYOUT(s, "trmps_bond_gradient");
  const BondGeom g = bond_geom(s, site, dir);
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = phi_pair(s, lay, g, st))) return rc;
  if ((rc = do_forward(s, lay, g, s->work + lay.bondM, st))) return rc;
  return do_gradient(s, lay, g, opt, 1, st, nullptr);
}

extern "C" int trmps_bond_split(const trmps_sweep* s, int32_t site, int32_t dir, const trmps_opt* opt, void* stream) {
  int rc = check_bond(s, site, dir, "trmps_bond_split");
  if (rc) return rc;
  if (!opt) { set_error("trmps_bond_split: null opt"); return TRMPS_E_ARG; }
  GET_LAYOUT(s, "trmps_bond_split");
  return svd_split(s, lay, bond_geom(s, site, dir), opt, (cudaStream_t)stream);
}

extern "C" int trmps_comm_unique_id(void* id128) {
  if (!id128) { set_error("trmps_comm_unique_id: null"); return TRMPS_E_ARG; }
  int rc = nccl_load();
  if (rc) return rc;
  nccl_uid uid;
  int e = g_nccl.get_uid(&uid);
  if (e != 0) return nccl_fail(e, "ncclGetUniqueId");
  memcpy(id128, &uid, sizeof(uid));
  return TRMPS_OK;
}

extern "C" int trmps_comm_create(const void* id128, int32_t rank, int32_t world, int32_t device, void** comm_out) {
  if (!id128 || !comm_out || world < 1 || rank < 0 || rank >= world) {
    set_error("trmps_comm_create: bad argument");
    return TRMPS_E_ARG;
  }
  int rc = nccl_load();
  if (rc) return rc;
  TRMPS_CUDA_OK(cudaSetDevice(device));
  nccl_uid uid;
  memcpy(&uid, id128, sizeof(uid));
  void* comm = nullptr;
  int e = g_nccl.init_rank(&comm, world, uid, rank);
  if (e != 0) return nccl_fail(e, "ncclCommInitRank");
  *comm_out = comm;
  return TRMPS_OK;
}

extern "C" int trmps_comm_destroy(void* comm) {
  if (!comm) return TRMPS_OK;
  if (!g_nccl.h) { set_error("trmps_comm_destroy: NCCL not loaded"); return TRMPS_E_NCCL; }
  int e = g_nccl.destroy(comm);
  if (e != 0) return nccl_fail(e, "ncclCommDestroy");
  return TRMPS_OK;
}
