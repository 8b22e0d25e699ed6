// C ABI of libl2hmc.so (see include/l2hmc.h).  Host side: context, parameter packing, launches.
#include "../../include/l2hmc.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "common.cuh"
#include "kernel_tile.cuh"
#include <cuda_fp16.h>
#include "kernel_tc.cuh"
#include "kernel_tc_s.cuh"
#include "kernel_small.cuh"
#include "layered.cuh"
#include "tc_gemm.cuh"
#include "tc_net.cuh"

#include <map>

using namespace l2hmc;

// ---------------------------------------------------------------------------------------------
// Context
// ---------------------------------------------------------------------------------------------
struct DevBuf {
  float *p = nullptr;
  size_t n = 0;
};

// layered engine (layered.cuh / layered_host.cuh)
struct LayNetView {
  const float *Wemb = nullptr, *tb = nullptr, *W4 = nullptr, *b4 = nullptr, *Wh = nullptr, *bh = nullptr,
              *es = nullptr, *eq = nullptr;
};
struct LayMlp {  // Linear / softplus stack (decoder of the energy, aux encoder of the nets)
  int n_layers = 0;
  std::vector<int> w, wp;  // widths and widths rounded up to 8
  DevBuf buf;
  std::vector<const float *> W, Wt, b;
};
struct LayTcWeight {  // one weight matrix pre-split / pre-tiled for tc_gemm_kernel: tf32 image, and fp16 image when allowed
  tcg::TcGemmB d, d16;
  DevBuf buf, buf16;
  bool has16 = false;
};
struct LayeredCtx {
  layered::LayDims dm;
  bool gemm_tc = true;                             // tcgen05 split GEMMs (false: fp32 FMA sgemm_kernel)
  bool gemm_f16 = true;                            // fp16 x3 operand split (default; L2HMC_LAYERED_GEMM=tf32 keeps tf32 x3)
  bool used_f16 = false;                           // what the last GEMM launch used
  std::map<const float *, LayTcWeight> tcw;        // keyed by the device pointer of the row-major weight
  int sms = 0;
  DevBuf net_buf[2];
  LayNetView net[2];
  LayMlp dec, enc;
  DevBuf x, v, x0, ab, hd, hA, hB, vec, eaux, auxp, tbias;
  std::vector<DevBuf> dact, eact;
  std::vector<DevBuf> aimg, gimg;                  // operand images of the decoder's activations / gradients (SplitImage)
  DevBuf ximg, abimg, hAimg, hBimg;                // ... of x, of the net input [a | b], of the nets' two hidden activations
  bool presplit = true;                            // L2HMC_LAYERED_PRESPLIT=0: every GEMM converts its own A operand
  int presplit_mode = 2;                           // 2: tc_gemm_pre_kernel (default); 1: the 256-row kernel with a TMA-fed A
  bool fused_net = true;                           // one kernel per S/T/Q net call (tc_net.cuh); L2HMC_LAYERED_FUSED_NET=0: three GEMMs
  long long ws_n = 0;
  cudaEvent_t ws_event = nullptr;  // recorded when a call's last workspace user is enqueued
  bool ws_recorded = false;
  float like_scale = 1.0f;     // weight of the Bernoulli log-likelihood in the decoder energy (AIS: beta)
  const float *aux = nullptr;  // bound by l2hmc_bind_aux for the component calls
  long long aux_n = 0;
};

struct l2hmc_ctx {
  l2hmc_config cfg;
  Shape sh;
  int kernel = L2HMC_KERNEL_TILE;
  std::string err;
  // device copies
  DevBuf net_packed[2];  // tile layout
  DevBuf net_raw[2];     // reference layout
  NetDev net_dev[2];
  NetRaw net_rawv[2];
  bool net_set[2] = {false, false};
  // host copies of the raw nets (eps / T changes do not need them, kept for re-packing)
  std::vector<float> net_host[2];
  DevBuf mask;
  DevBuf train_ws;       // scratch of l2hmc_loss_grad (train_host.cuh), grown on demand, kept until destroy
  float *train_part = nullptr;  // inside train_ws: the parts of the call's split reductions
  bool mask_set = false;
  DevBuf energy_buf, energy_buf2;  // energy_buf2: the second part of a mixed energy
  EnergyDev en;
  bool energy_set = false;
  int64_t launches = 0;
  // timing
  bool timing = false;
  std::vector<cudaEvent_t> ev;  // pairs
  size_t ev_used = 0;
  // host-call staging
  DevBuf hx, hv, hu, hxo, hvo, hpx, hxn;
  // tensor-core kernel: pre-split weight streams
  tc::TcDims td;
  DevBuf tc_buf[2], tc_gbuf, tc_hc[2];
  bool tc_used_f16 = false;  // the last tensor-core launch used the fp16 operand split
  size_t tc_g_h_off = 0;  // offset of the fp16 twin inside tc_gbuf
  float tc_wmax[3] = {0.f, 0.f, 0.f};  // largest |value| packed for the X net, the V net, the Gaussian grad (fp16 range check)
  std::vector<float> tc_head_raw[2];  // per net: bs | bt | bq | e^{scale_s} | e^{scale_q}, DP each (host copy for tc_pack_hc)
  tc::TcNet tc_net[2] = {};
  bool tc_ok = false;  // shape / energy kind inside what kernel_tc covers
  uint8_t *hdir = nullptr, *hacc = nullptr;
  size_t hdir_n = 0, hacc_n = 0;
  cudaStream_t hstream = nullptr;
  cudaStream_t hstreams[3] = {nullptr, nullptr, nullptr};  // chunk pipeline of l2hmc_transition_host
  LayeredCtx lay;
  DevBuf haux, diag;
  double host_stats0[2] = {0.0, 0.0};  // l2hmc_transition_host: the caller's stats before the call (for the tf32 repeat)
  DevBuf hstats;                        // device twin of the host stats accumulators (2 doubles in a float buffer of 4)
  // status word of this context: pinned host memory mapped into the device; kernels raise sticky bits with
  // atomicOr_system, the host reads it WITHOUT synchronising (l2hmc_status_flags, and before every tensor-core launch)
  unsigned int *status_h = nullptr, *status_d = nullptr;
  // cudaFuncAttributeMaxDynamicSharedMemorySize already granted on this context's device: [0] tc_s, [1] tc, [2] tile
  size_t smem_cfg[3] = {0, 0, 0};
};

static thread_local std::string g_err;

static int fail(l2hmc_ctx *ctx, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf; else g_err = buf;
  return code;
}

#define CUDA_TRY(ctx, expr)                                                                      \
  do {                                                                                           \
    cudaError_t e_ = (expr);                                                                     \
    if (e_ != cudaSuccess)                                                                       \
      return fail(ctx, L2HMC_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

static int ensure(l2hmc_ctx *ctx, DevBuf &b, size_t n) {
  if (b.n >= n && b.p) return L2HMC_OK;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.n = 0;
  CUDA_TRY(ctx, cudaMalloc(&b.p, n * sizeof(float)));
  b.n = n;
  return L2HMC_OK;
}

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

#include "layered_host.cuh"

// ---------------------------------------------------------------------------------------------
// Component kernels (Dynamics methods; not the hot path)
// ---------------------------------------------------------------------------------------------
__global__ void k_energy(EnergyDev en, Shape sh, long long n, const float *x, float *out) {
  long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g < n) out[g] = energy_chain(en, sh, x + g * sh.D, 1);
}
__global__ void k_grad(EnergyDev en, Shape sh, long long n, const float *x, float *out) {
  long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g < n) grad_chain(en, sh, x + g * sh.D, 1, out + g * sh.D, 1);
}
__global__ void k_kinetic(int D, long long n, const float *v, float *out) {
  long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g < n) {
    float s = 0.f;
    for (int d = 0; d < D; ++d) s = fmaf(v[g * D + d], v[g * D + d], s);
    out[g] = 0.5f * s;  // utils/dynamics.py:107-108
  }
}
__global__ void k_hamiltonian(EnergyDev en, Shape sh, long long n, const float *x, const float *v, float *out) {
  long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g < n) {
    float s = 0.f;
    for (int d = 0; d < sh.D; ++d) s = fmaf(v[g * sh.D + d], v[g * sh.D + d], s);
    out[g] = energy_chain(en, sh, x + g * sh.D, 1) + 0.5f * s;
  }
}
__global__ void k_p_accept(EnergyDev en, Shape sh, long long n, const float *x0, const float *v0,
                           const float *x1, const float *v1, const float *lj, float *out) {
  long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g < n) {
    float k0 = 0.f, k1 = 0.f;
    for (int d = 0; d < sh.D; ++d) {
      k0 = fmaf(v0[g * sh.D + d], v0[g * sh.D + d], k0);
      k1 = fmaf(v1[g * sh.D + d], v1[g * sh.D + d], k1);
    }
    const float e_new = energy_chain(en, sh, x1 + g * sh.D, 1) + 0.5f * k1;
    const float e_old = energy_chain(en, sh, x0 + g * sh.D, 1) + 0.5f * k0;
    out[g] = accept_prob(e_old, e_new, lj[g]);
  }
}

constexpr int NET_MAXH = 256;
__global__ void k_net_apply(NetRaw w, int D, int H, int T, int hmc, long long n, const float *a, const float *b,
                            float step, float *S, float *Tt, float *Q, const float *eaux, int lde) {
  long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g >= n) return;
  if (hmc) {
    for (int d = 0; d < D; ++d) S[g * D + d] = Tt[g * D + d] = Q[g * D + d] = 0.f;
    return;
  }
  float h1[NET_MAXH], h2[NET_MAXH];
  const float arg = 6.2831855f * step / (float)T;  // utils/dynamics.py:99-105
  const float ct = cosf(arg), st = sinf(arg);
  for (int j = 0; j < H; ++j) {
    float e1 = 0.f, e2 = 0.f;
    for (int i = 0; i < D; ++i) {
      e1 = fmaf(a[g * D + i], w.W1[i * H + j], e1);
      e2 = fmaf(b[g * D + i], w.W2[i * H + j], e2);
    }
    float e3 = fmaf(st, w.W3[H + j], ct * w.W3[j]);
    float s = (((0.f + (e1 + w.b1[j])) + (e2 + w.b2[j])) + (e3 + w.b3[j]));
    if (eaux) s += eaux[g * lde + j];  // 4th Zip entry: encoder_sampler(aux), mnist_vae.py:149
    h1[j] = fmaxf(s, 0.f);
  }
  for (int j = 0; j < H; ++j) {
    float s = 0.f;
    for (int i = 0; i < H; ++i) s = fmaf(h1[i], w.W4[i * H + j], s);
    h2[j] = fmaxf(s + w.b4[j], 0.f);
  }
  for (int d = 0; d < D; ++d) {
    float s = 0.f, t = 0.f, q = 0.f;
    for (int i = 0; i < H; ++i) {
      s = fmaf(h2[i], w.Ws[i * D + d], s);
      t = fmaf(h2[i], w.Wt[i * D + d], t);
      q = fmaf(h2[i], w.Wq[i * D + d], q);
    }
    S[g * D + d] = expf(w.ls[d]) * tanhf(s + w.bs[d]);
    Tt[g * D + d] = t + w.bt[d];
    Q[g * D + d] = expf(w.lq[d]) * tanhf(q + w.bq[d]);
  }
}

__global__ void k_accept(int D, long long n, long long off, const float *x, const float *Lx, const float *px,
                         const float *u, unsigned long long seed, unsigned long long counter, float *out,
                         uint8_t *accepted) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n * D) return;
  const long long g = i / D;
  float uu;
  if (u) uu = u[g];
  else {
    int dd;
    philox_dir_u(seed, counter, off + g, dd, uu);
  }
  const bool acc = (px[g] - uu >= 0.f);
  out[i] = acc ? Lx[i] : x[i];
  if (accepted && (i - g * D) == 0) accepted[g] = acc ? 1 : 0;
}

__global__ void k_philox_fill(int D, long long n, long long off, unsigned long long seed,
                              unsigned long long counter, float *v, uint8_t *dir, float *u) {
  long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g >= n) return;
  if (v) {
    for (int b = 0; b < (D + 3) / 4; ++b) {
      float z[4];
      philox_normals4(seed, counter, off + g, b, z);
      for (int q = 0; q < 4; ++q)
        if (4 * b + q < D) v[g * D + 4 * b + q] = z[q];
    }
  }
  if (dir || u) {
    int dbit;
    float uu;
    philox_dir_u(seed, counter, off + g, dbit, uu);
    if (dir) dir[g] = (uint8_t)dbit;
    if (u) u[g] = uu;
  }
}

// ---------------------------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------------------------
extern "C" const char *l2hmc_version(void) { return "l2hmc_b200 0.1 (sm_100a)"; }

extern "C" const char *l2hmc_last_error(const l2hmc_ctx *ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

// ---- tensor-core kernel: host-side operand preparation ------------------------------------------------
static inline float tf32_rna_host(float x) {  // cvt.rna.tf32.f32: round to nearest, ties away from zero
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

// Append the chunk stream of one B operand (B[n][k] = get(k, n), zero outside the real extent) to `out`:
// for every K=8 step one chunk = {hi slab, lo slab}, each slab Npad x 8 floats in the canonical K-major
// no-swizzle core-matrix order [k_core (2)][n_group (Npad/8)][row (8)][4 floats]  (LBO = Npad/8 * 128 B, SBO = 128 B).
template <class Get>
static void append_b_stream(std::vector<float> &out, int Kpad, int Npad, Get get) {
  const int NG = Npad / 8;
  for (int ks = 0; ks < Kpad / 8; ++ks) {
    const size_t base = out.size();
    out.resize(base + (size_t)16 * Npad, 0.f);
    float *hi = out.data() + base, *lo = hi + (size_t)8 * Npad;
    for (int n = 0; n < Npad; ++n)
      for (int kk = 0; kk < 8; ++kk) {
        const float w = get(8 * ks + kk, n);
        const float h = tf32_rna_host(w);
        const float l = tf32_rna_host(w - h);
        const size_t idx = ((size_t)(kk >> 2) * NG + (n >> 3)) * 32 + (size_t)(n & 7) * 4 + (kk & 3);
        hi[idx] = h;
        lo[idx] = l;
      }
  }
}

// fp16 twin of append_b_stream for the kind::f16 MMAs (K step = 16): per K step one chunk = {hi slab, lo slab}, each
// slab Npad x 16 halfs in the K-major no-swizzle core-matrix order [k_core (2)][n_group (Npad/8)][row (8)][8 halfs] --
// 16 * Npad 32-bit words per chunk, like one tf32 K step.  Returns the largest |value| (fp16 range check).
template <class Get>
static float append_b_stream_f16(std::vector<float> &out, int K, int Npad, Get get) {
  const int NG = Npad / 8, nsteps = (K + 15) / 16;
  float amax = 0.f;
  for (int ks = 0; ks < nsteps; ++ks) {
    const size_t base = out.size();
    out.resize(base + (size_t)16 * Npad, 0.f);
    uint16_t *hi = reinterpret_cast<uint16_t *>(out.data() + base), *lo = hi + (size_t)16 * Npad;
    for (int n = 0; n < Npad; ++n)
      for (int kk = 0; kk < 16; ++kk) {
        const int k = 16 * ks + kk;
        const float w = k < K ? get(k, n) : 0.f;
        amax = fmaxf(amax, fabsf(w));
        const __half h = __float2half_rn(w);
        const __half l = __float2half_rn(w - __half2float(h));
        const size_t idx = ((size_t)(kk >> 3) * NG + (n >> 3)) * 64 + (size_t)(n & 7) * 8 + (kk & 7);
        memcpy(&hi[idx], &h, 2);
        memcpy(&lo[idx], &l, 2);
      }
  }
  return amax;
}

static void tc_setup_dims(l2hmc_ctx *ctx) {
  const Shape &sh = ctx->sh;
  tc::TcDims &td = ctx->td;
  td.K1 = 2 * sh.DP;
  td.HK = round_up(sh.H, 8);
  td.N1 = round_up(td.HK, 16);
  td.N3 = round_up(3 * sh.DP, 16);
  td.KG = round_up(sh.DP, 8);
  td.NG = round_up(sh.DP, 16);
  int nmax = td.N1 > td.N3 ? td.N1 : td.N3;
  if (td.NG > nmax) nmax = td.NG;
  td.slot_floats = tc::KSLOT * 16 * nmax;
  const long long state_bytes = (long long)tc::make_tclay(sh.DP, sh.T).ring * 4;
  long long ns = (232448LL - 1024 - state_bytes) / ((long long)td.slot_floats * 4);
  td.nslot = ns > tc::MAX_SLOT ? tc::MAX_SLOT : (int)ns;
  // Tunables measured on B200 (profiles/r01_tc_*): 2 compute threads per chain beat 3 and 4 (10.6 / 11.1 / 12.1 ms
  // on config 2), and the ex2/rcp-based exp and tanh keep parity at the fp32 noise floor while saving 4%.
  const char *fm = getenv("L2HMC_TC_FAST_MATH");
  td.fast_math = (fm && fm[0] == '0') ? 0 : 1;
  const char *nqe = getenv("L2HMC_TC_NQ");
  td.nq = (nqe && nqe[0] >= '2' && nqe[0] <= '4') ? (nqe[0] - '0') : 2;
  // kernel_tc_s: biases as weight rows (needs two pad dimensions in the last 4-dim chunk and a pad hidden unit)
  const char *bg = getenv("L2HMC_TC_BIASG");
  // 1: the direction one-hot sits in the two pad dimensions of the last 4-dim chunk (x_dim <= DP - 2); 2: no pad dimensions
  // (x_dim = DP) -> a K step of its own behind the net input (needs whole 16-k steps before it: DP / 4 even, and room in the
  // A operand: DP / 4 <= 12).  L2HMC_TC_BIASG=0: explicit biases everywhere; =1: only the round-1 case (config 2's shape).
  td.biasg = 0;
  if (!(bg && bg[0] == '0') && sh.H <= td.HK - 1) {
    const int nqc = sh.DP / 4;
    if (sh.D <= sh.DP - 2) td.biasg = (bg && bg[0] == '1' && sh.DP != 52) ? 0 : 1;
    else if (nqc % 2 == 0 && nqc <= 12 && !(bg && bg[0] == '1')) td.biasg = 2;
  }
  const char *fe = getenv("L2HMC_TC_F16");
  td.f16 = (fe && fe[0] == '0') ? 0 : 1;  // fp16 split when every packed value is inside the fp16 range (checked at pack time)
  ctx->tc_ok = !sh.hmc && td.K1 <= 128 && td.HK <= 128 && td.N1 <= 192 && td.N3 <= 192 && td.nslot >= 4 && td.K1 % 8 == 0;
}

static int tc_pack_hc(l2hmc_ctx *ctx, int net_id);

static int tc_pack_net(l2hmc_ctx *ctx, int net_id, const l2hmc_net_params *p) {
  const Shape &sh = ctx->sh;
  const tc::TcDims &td = ctx->td;
  const int D = sh.D, H = sh.H, DP = sh.DP, T = sh.T;
  std::vector<float> img;
  append_b_stream(img, td.K1, td.N1, [&](int k, int n) -> float {  // embed: rows [a | b]
    if (n >= H) return 0.f;
    if (k < DP) return k < D ? p->W1[(size_t)k * H + n] : 0.f;
    const int d = k - DP;
    return d < D ? p->W2[(size_t)d * H + n] : 0.f;
  });
  append_b_stream(img, td.HK, td.N1, [&](int k, int n) -> float { return (k < H && n < H) ? p->W4[(size_t)k * H + n] : 0.f; });
  append_b_stream(img, td.HK, td.N3, [&](int k, int n) -> float {  // heads: columns S | T | Q blocks of DP
    if (k >= H || n >= 3 * DP) return 0.f;
    const int blk = n / DP, d = n - blk * DP;
    if (d >= D) return 0.f;
    const float *W = blk == 0 ? p->Ws : (blk == 1 ? p->Wt : p->Wq);
    return W[(size_t)k * D + d];
  });
  // the specialised kernel (kernel_tc_s.cuh): net input interleaved per 4-dim chunk (K step q = [a_{4q..4q+3} | b_{4q..4q+3}]),
  // heads split by dimensions into heads_a (first CA chunks) and heads_b (the rest), each with S | T | Q column blocks.
  // biasg: the biases ride in the GEMMs.  Hidden unit H (a pad column) is the constant 1: the embed produces it from the
  // direction one-hot the kernel puts into the pad dimensions DP-2 / DP-1 of the a-part (rows kf / kb of the last K
  // step, which also carry the time-embedding bias row of the chain's step: tb[it] forward, tb[T-1-it] backward -> one
  // image of that K step per leapfrog step, `emb_last`), the hidden layer passes it on and adds b4, the heads add bs/bt/bq.
  const bool biasg = td.biasg != 0;
  // K rows of the one-hot: the a-part pad dimensions DP-2, DP-1 in the interleaved order (biasg 1), or the first two rows of
  // a K step of its own behind the net input (biasg 2: that step holds nothing else)
  const int kf = td.biasg == 2 ? td.K1 : td.K1 - 6, kb = kf + 1;
  auto embed_w = [&](int k, int n) -> float {
    if (n >= H) return 0.f;
    const int d = 4 * (k / 8) + (k & 3);
    if (d >= D) return 0.f;
    return ((k & 7) < 4 ? p->W1 : p->W2)[(size_t)d * H + n];
  };
  auto tb_at = [&](int t, int j) -> float {
    const float arg = 6.2831855f * (float)t / (float)T;  // utils/dynamics.py:99-105 in fp32
    const float ct = cosf(arg), st = sinf(arg);
    return (p->b1[j] + p->b2[j]) + (fmaf(st, p->W3[H + j], ct * p->W3[j]) + p->b3[j]);
  };
  std::vector<float> img_s;
  append_b_stream(img_s, td.K1, td.N1, embed_w);
  append_b_stream(img_s, td.HK, td.N1, [&](int k, int n) -> float {
    if (biasg && k == H) return n < H ? p->b4[n] : (n == H ? 1.f : 0.f);
    return (k < H && n < H) ? p->W4[(size_t)k * H + n] : 0.f;
  });
  {
    const int nqc = DP / 4, ca = (nqc + 1) / 2, cb = nqc - ca;
    for (int part = 0; part < 2; ++part) {
      const int cp = part == 0 ? ca : cb, d0 = part == 0 ? 0 : 4 * ca, np = round_up(12 * cp, 16);
      if (cp == 0) continue;
      append_b_stream(img_s, td.HK, np, [&](int k, int n) -> float {
        if (k > H || (k == H && !biasg) || n >= 12 * cp) return 0.f;
        const int blk = n / (4 * cp), d = d0 + n - blk * 4 * cp;
        if (d >= D) return 0.f;
        if (k == H) return (blk == 0 ? p->bs : (blk == 1 ? p->bt : p->bq))[d];
        const float *W = blk == 0 ? p->Ws : (blk == 1 ? p->Wt : p->Wq);
        return W[(size_t)k * D + d];
      });
    }
  }
  const size_t nstream_s = img_s.size();
  if (biasg) {  // per leapfrog step: the last K step of the embed with the bias rows
    for (int t = 0; t < T; ++t) {
      std::vector<float> one;
      append_b_stream(one, 8, td.N1, [&](int kk, int n) -> float {
        const int k = (td.biasg == 2 ? td.K1 : td.K1 - 8) + kk;
        if (k == kf) return n < H ? tb_at(t, n) : (n == H ? 1.f : 0.f);
        if (k == kb) return n < H ? tb_at(T - 1 - t, n) : (n == H ? 1.f : 0.f);
        return embed_w(k, n);
      });
      img_s.insert(img_s.end(), one.begin(), one.end());
    }
  }
  const size_t nimg_s = img_s.size();
  // fp16 twins (K steps of 16; the embed's last K step covers k = K1-8 .. K1+7)
  std::vector<float> img_h;
  size_t nstream_h = 0;
  float wmax = 0.f;
  {
    wmax = fmaxf(wmax, append_b_stream_f16(img_h, td.K1, td.N1, embed_w));
    wmax = fmaxf(wmax, append_b_stream_f16(img_h, td.HK, td.N1, [&](int k, int n) -> float {
      if (biasg && k == H) return n < H ? p->b4[n] : (n == H ? 1.f : 0.f);
      return (k < H && n < H) ? p->W4[(size_t)k * H + n] : 0.f;
    }));
    const int nqc = DP / 4, ca = (nqc + 1) / 2, cb = nqc - ca;
    for (int part = 0; part < 2; ++part) {
      const int cp = part == 0 ? ca : cb, d0 = part == 0 ? 0 : 4 * ca, np = round_up(12 * cp, 16);
      if (cp == 0) continue;
      wmax = fmaxf(wmax, append_b_stream_f16(img_h, td.HK, np, [&](int k, int n) -> float {
        if (k > H || (k == H && !biasg) || n >= 12 * cp) return 0.f;
        const int blk = n / (4 * cp), d = d0 + n - blk * 4 * cp;
        if (d >= D) return 0.f;
        if (k == H) return (blk == 0 ? p->bs : (blk == 1 ? p->bt : p->bq))[d];
        const float *W = blk == 0 ? p->Ws : (blk == 1 ? p->Wt : p->Wq);
        return W[(size_t)k * D + d];
      }));
    }
    nstream_h = img_h.size();
    if (biasg) {
      const int k0 = td.biasg == 2 ? td.K1 : (td.K1 - 1) / 16 * 16;  // first k of the embed's last K = 16 step
      for (int t = 0; t < T; ++t)
        wmax = fmaxf(wmax, append_b_stream_f16(img_h, 16, td.N1, [&](int kk, int n) -> float {
          const int k = k0 + kk;
          if (k == kf) return n < H ? tb_at(t, n) : (n == H ? 1.f : 0.f);
          if (k == kb) return n < H ? tb_at(T - 1 - t, n) : (n == H ? 1.f : 0.f);
          if (k >= td.K1) return 0.f;
          return embed_w(k, n);
        }));
    }
  }
  ctx->tc_wmax[net_id] = wmax;
  const size_t nimg_h = img_h.size();
  const size_t nimg = img.size(), ntb = (size_t)T * td.N1, nb4 = td.N1, nbh = td.N3, nes = DP;
  std::vector<float> buf(nimg + ntb + nb4 + nbh + 2 * nes + nimg_s + nimg_h, 0.f);
  memcpy(buf.data(), img.data(), nimg * sizeof(float));
  memcpy(buf.data() + nimg + ntb + nb4 + nbh + 2 * nes, img_s.data(), nimg_s * sizeof(float));
  memcpy(buf.data() + nimg + ntb + nb4 + nbh + 2 * nes + nimg_s, img_h.data(), nimg_h * sizeof(float));
  float *tb = buf.data() + nimg, *b4 = tb + ntb, *bh = b4 + nb4, *es = bh + nbh, *eq = es + nes;
  for (int t = 0; t < T; ++t) {
    const float arg = 6.2831855f * (float)t / (float)T;  // utils/dynamics.py:99-105 in fp32
    const float ct = cosf(arg), st = sinf(arg);
    for (int j = 0; j < H; ++j) tb[(size_t)t * td.N1 + j] = (p->b1[j] + p->b2[j]) + (fmaf(st, p->W3[H + j], ct * p->W3[j]) + p->b3[j]);
  }
  for (int j = 0; j < H; ++j) b4[j] = p->b4[j];
  for (int d = 0; d < DP; ++d) {
    es[d] = eq[d] = 1.f;
    if (d < D) {
      bh[d] = p->bs[d];
      bh[DP + d] = p->bt[d];
      bh[2 * DP + d] = p->bq[d];
      es[d] = expf(p->scale_s[d]);
      eq[d] = expf(p->scale_q[d]);
    }
  }
  int rc = ensure(ctx, ctx->tc_buf[net_id], buf.size());
  if (rc) return rc;
  CUDA_TRY(ctx, cudaMemcpy(ctx->tc_buf[net_id].p, buf.data(), buf.size() * sizeof(float), cudaMemcpyHostToDevice));
  tc::TcNet &n = ctx->tc_net[net_id];
  n.img = ctx->tc_buf[net_id].p;
  n.tb = n.img + nimg;
  n.b4 = n.tb + ntb;
  n.bh = n.b4 + nb4;
  n.es = n.bh + nbh;
  n.eq = n.es + nes;
  n.img_s = n.eq + nes;
  n.emb_last = biasg ? n.img_s + nstream_s : nullptr;
  n.img_h = n.img_s + nimg_s;
  n.emb_last_h = biasg ? n.img_h + nstream_h : nullptr;
  std::vector<float> &raw = ctx->tc_head_raw[net_id];
  raw.assign((size_t)5 * DP, 0.f);
  for (int d = 0; d < D; ++d) {
    raw[d] = p->bs[d];
    raw[DP + d] = p->bt[d];
    raw[2 * DP + d] = p->bq[d];
    raw[3 * DP + d] = es[d];
    raw[4 * DP + d] = eq[d];
  }
  return tc_pack_hc(ctx, net_id);
}

// Per-dimension constants of the specialised kernel's heads epilogue (kernel_tc_s.cuh), pre-multiplied so that tanh and
// exp run in log2 units: per 4-dim chunk {bs2, bq2, n2cS, cS, n2cQ, cQ, bth} x 4 floats.  h = eps/2 for the V net
// (momentum half steps, utils/dynamics.py:121-125) and eps for the X net (:133-145); depends on eps -> repacked by l2hmc_set_eps.
static int tc_pack_hc(l2hmc_ctx *ctx, int net_id) {
  const int DP = ctx->sh.DP;
  const std::vector<float> &raw = ctx->tc_head_raw[net_id];
  if (raw.size() != (size_t)5 * DP) return L2HMC_OK;  // net not set yet
  const double L2E = 1.4426950408889634, eps = (double)ctx->sh.eps;
  const double h = net_id == L2HMC_VNET ? 0.5 * eps : eps;
  std::vector<float> hc((size_t)(DP / 4) * tc::HC_PER_CHUNK, 0.f);
  for (int d = 0; d < ctx->sh.D; ++d) {
    float *c = hc.data() + (size_t)(d / 4) * tc::HC_PER_CHUNK + (d & 3);
    const double bs = raw[d], bt = raw[DP + d], bq = raw[2 * DP + d], es = raw[3 * DP + d], eq = raw[4 * DP + d];
    const double cS = es * h * L2E, cQ = eq * eps * L2E;
    c[0] = (float)(bs * 2.0 * L2E);  // (the bias entries are unused when the biases ride in the GEMMs, td.biasg)
    c[4] = (float)(bq * 2.0 * L2E);
    c[8] = (float)(-2.0 * cS);
    c[12] = (float)cS;
    c[16] = (float)(-2.0 * cQ);
    c[20] = (float)cQ;
    c[24] = (float)(bt * h);
  }
  int rc = ensure(ctx, ctx->tc_hc[net_id], hc.size());
  if (rc) return rc;
  CUDA_TRY(ctx, cudaMemcpy(ctx->tc_hc[net_id].p, hc.data(), hc.size() * sizeof(float), cudaMemcpyHostToDevice));
  ctx->tc_net[net_id].hc = ctx->tc_hc[net_id].p;
  return L2HMC_OK;
}

static int tc_pack_gaussian(l2hmc_ctx *ctx, const float *Ssym_padded /* [DP][LDS] host */) {
  const Shape &sh = ctx->sh;
  const tc::TcDims &td = ctx->td;
  std::vector<float> img;
  append_b_stream(img, td.KG, td.NG, [&](int k, int n) -> float {  // g_n = sum_k d_k Ssym[k][n]
    return (k < sh.D && n < sh.D) ? Ssym_padded[(size_t)k * sh.LDS + n] : 0.f;
  });
  const size_t ng32 = img.size();
  ctx->tc_wmax[2] = append_b_stream_f16(img, td.KG, td.NG, [&](int k, int n) -> float {
    return (k < sh.D && n < sh.D) ? Ssym_padded[(size_t)k * sh.LDS + n] : 0.f;
  });
  int rc = ensure(ctx, ctx->tc_gbuf, img.size());
  if (rc) return rc;
  CUDA_TRY(ctx, cudaMemcpy(ctx->tc_gbuf.p, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice));
  ctx->tc_g_h_off = ng32;
  return L2HMC_OK;
}

// Which kernel a transition launches: the explicit request, or for AUTO
//   thread-per-chain kernel for tiny nets, tensor-core kernel for nets wide enough to fill 128x(>=32)x(>=16) MMAs on
//   the energies it covers, the generic FMA tile kernel otherwise.
static int resolve_kernel(l2hmc_ctx *ctx, int *out) {
  const Shape &sh = ctx->sh;
  int k = ctx->cfg.kernel;
  if (k == L2HMC_KERNEL_LAYERED_FMA) {  // the layered engine with fp32-FMA GEMMs
    ctx->lay.gemm_tc = false;
    k = L2HMC_KERNEL_LAYERED;
  }
  const bool small_ok = sh.D <= 4 && (sh.hmc || sh.H <= 16);
  const bool tc_energy = ctx->energy_set && ((ctx->en.kind == L2HMC_ENERGY_GAUSSIAN && ctx->en.ncomp == 1) ||
                                             ctx->en.kind == L2HMC_ENERGY_ROUGHWELL);
  const bool tc_auto = ctx->tc_ok && tc_energy && sh.D >= 8 && sh.H >= 32;
  // what only the layered engine covers: the decoder energy, aux-conditioned nets, shapes beyond one SM's tile
  const bool needs_layered = (ctx->energy_set && ctx->en.kind == L2HMC_ENERGY_DECODER) || ctx->lay.enc.n_layers > 0 ||
                             sh.DP > 64 || (!sh.hmc && sh.HP > 128);
  if (k == L2HMC_KERNEL_AUTO)
    k = needs_layered ? L2HMC_KERNEL_LAYERED
                      : (small_ok ? L2HMC_KERNEL_SMALL : (tc_auto ? L2HMC_KERNEL_TC : L2HMC_KERNEL_TILE));
  if (ctx->energy_set && ctx->en.kind == L2HMC_ENERGY_MIXED && k != L2HMC_KERNEL_SMALL && k != L2HMC_KERNEL_TILE)
    return fail(ctx, L2HMC_EUNSUPPORTED, "a mixed (annealed) energy is evaluated by the small and tile kernels (x_dim <= 64)");
  if (k != L2HMC_KERNEL_LAYERED && ((ctx->energy_set && ctx->en.kind == L2HMC_ENERGY_DECODER) || ctx->lay.enc.n_layers > 0))
    return fail(ctx, L2HMC_EUNSUPPORTED, "the decoder energy and aux-conditioned nets run on the layered engine only");
  if (k == L2HMC_KERNEL_LAYERED) {
    if (!sh.hmc && sh.H > 65536) return fail(ctx, L2HMC_EUNSUPPORTED, "layered engine: width too large");
  } else if (k == L2HMC_KERNEL_TILE) {
    if (sh.DP > 64 || (!sh.hmc && sh.HP > 128))
      return fail(ctx, L2HMC_EUNSUPPORTED, "tile kernel covers x_dim <= 64 and width <= 128 (got %d, %d)", sh.D, sh.H);
  } else if (k == L2HMC_KERNEL_SMALL) {
    if (!small_ok) return fail(ctx, L2HMC_EUNSUPPORTED, "small kernel covers x_dim <= 4 and width <= 16 (got %d, %d)", sh.D, sh.H);
  } else if (k == L2HMC_KERNEL_TC) {
    if (!ctx->tc_ok) return fail(ctx, L2HMC_EUNSUPPORTED, "tensor-core kernel does not cover this shape (x_dim %d, width %d, hmc %d)", sh.D, sh.H, sh.hmc);
    if (ctx->energy_set && !((ctx->en.kind == L2HMC_ENERGY_GAUSSIAN && ctx->en.ncomp == 1) || ctx->en.kind == L2HMC_ENERGY_ROUGHWELL))
      return fail(ctx, L2HMC_EUNSUPPORTED, "tensor-core kernel covers the Gaussian and RoughWell energies (kind %d given)", ctx->en.kind);
  } else {
    return fail(ctx, L2HMC_EUNSUPPORTED, "kernel kind %d not available in this build", k);
  }
  *out = k;
  return L2HMC_OK;
}

static int pick_kernel(l2hmc_ctx *ctx) {
  tc_setup_dims(ctx);
  int k = 0;
  int rc = resolve_kernel(ctx, &k);
  if (rc) return rc;
  ctx->kernel = k;
  return L2HMC_OK;
}

extern "C" int l2hmc_create(const l2hmc_config *cfg, l2hmc_ctx **out) {
  if (!cfg || !out) return fail(nullptr, L2HMC_EINVAL, "l2hmc_create: null argument");
  if (cfg->x_dim < 1 || cfg->T < 1 || (!cfg->hmc && cfg->width < 1) || !(cfg->eps > 0.f))
    return fail(nullptr, L2HMC_EINVAL, "l2hmc_create: need x_dim >= 1, T >= 1, width >= 1, eps > 0");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, L2HMC_ECUDA, "l2hmc_create: no CUDA device (%s)", cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev)
    return fail(nullptr, L2HMC_EINVAL, "l2hmc_create: device %d out of range (%d devices)", cfg->device, ndev);
  l2hmc_ctx *ctx = new (std::nothrow) l2hmc_ctx();
  if (!ctx) return fail(nullptr, L2HMC_ENOMEM, "l2hmc_create: out of host memory");
  ctx->cfg = *cfg;
  Shape &sh = ctx->sh;
  sh.D = cfg->x_dim;
  sh.DP = round_up(cfg->x_dim, 4);
  sh.H = cfg->hmc ? 4 : cfg->width;
  sh.HP = round_up(sh.H, 4);
  sh.T = cfg->T;
  sh.LDE = 128;
  sh.LDH = 192;
  sh.LDS = round_up(sh.DP, 128);
  sh.hmc = cfg->hmc ? 1 : 0;
  sh.eps = cfg->eps;
  ctx->en.kind = L2HMC_ENERGY_NONE;
  ctx->en.temperature = 1.0f;
  lay_setup_dims(ctx);
  {
    const char *gm = getenv("L2HMC_LAYERED_GEMM");  // "fma" forces the fp32 FMA GEMMs
    if (gm && gm[0] == 't') ctx->lay.gemm_f16 = false;                 // "tf32": the tf32 x3 split only
    else if (gm && gm[0] == 'f' && gm[1] == 'm') ctx->lay.gemm_tc = false;  // "fma"
    const char *ps = getenv("L2HMC_LAYERED_PRESPLIT");
    if (ps && ps[0] == '0') ctx->lay.presplit = false;
    if (ps && ps[0] == '1') ctx->lay.presplit_mode = 1;
    const char *fn = getenv("L2HMC_LAYERED_FUSED_NET");
    if (fn && fn[0] == '0') ctx->lay.fused_net = false;
    cudaDeviceGetAttribute(&ctx->lay.sms, cudaDevAttrMultiProcessorCount, cfg->device);
  }
  int rc = pick_kernel(ctx);
  if (rc != L2HMC_OK) {
    g_err = ctx->err;
    delete ctx;
    return rc;
  }
  if (cudaSetDevice(cfg->device) != cudaSuccess) {
    delete ctx;
    return fail(nullptr, L2HMC_ECUDA, "l2hmc_create: cudaSetDevice(%d) failed", cfg->device);
  }
  if (cudaHostAlloc((void **)&ctx->status_h, sizeof(unsigned int), cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer((void **)&ctx->status_d, ctx->status_h, 0) != cudaSuccess) {
    cudaGetLastError();
    if (ctx->status_h) cudaFreeHost(ctx->status_h);
    delete ctx;
    return fail(nullptr, L2HMC_ECUDA, "l2hmc_create: cannot allocate the mapped status word");
  }
  *ctx->status_h = 0u;
  *out = ctx;
  return L2HMC_OK;
}

extern "C" void l2hmc_destroy(l2hmc_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->cfg.device);
  DevBuf *bufs[] = {&ctx->net_packed[0], &ctx->net_packed[1], &ctx->net_raw[0], &ctx->net_raw[1], &ctx->mask,
                    &ctx->energy_buf, &ctx->energy_buf2, &ctx->hx, &ctx->hv, &ctx->hu, &ctx->hxo, &ctx->hvo, &ctx->hpx, &ctx->hxn,
                    &ctx->tc_buf[0], &ctx->tc_buf[1], &ctx->tc_gbuf, &ctx->tc_hc[0], &ctx->tc_hc[1], &ctx->train_ws, &ctx->hstats};
  for (DevBuf *b : bufs)
    if (b->p) cudaFree(b->p);
  {
    LayeredCtx &L = ctx->lay;
    DevBuf *lb[] = {&L.net_buf[0], &L.net_buf[1], &L.dec.buf, &L.enc.buf, &L.x, &L.v, &L.x0, &L.ab, &L.hd, &L.hA, &L.hB,
                    &L.vec, &L.eaux, &L.auxp, &L.tbias, &ctx->haux, &ctx->diag, &L.ximg, &L.abimg, &L.hAimg, &L.hBimg};
    for (DevBuf *b : lb)
      if (b->p) cudaFree(b->p);
    for (DevBuf &b : L.aimg)
      if (b.p) cudaFree(b.p);
    for (DevBuf &b : L.gimg)
      if (b.p) cudaFree(b.p);
    for (DevBuf &b : L.dact)
      if (b.p) cudaFree(b.p);
    for (DevBuf &b : L.eact)
      if (b.p) cudaFree(b.p);
    for (auto &kv : L.tcw)
    {
      if (kv.second.buf.p) cudaFree(kv.second.buf.p);
      if (kv.second.buf16.p) cudaFree(kv.second.buf16.p);
    }
    if (L.ws_event) cudaEventDestroy(L.ws_event);
  }
  if (ctx->hdir) cudaFree(ctx->hdir);
  if (ctx->hacc) cudaFree(ctx->hacc);
  if (ctx->status_h) cudaFreeHost(ctx->status_h);
  for (cudaEvent_t e : ctx->ev) cudaEventDestroy(e);
  if (ctx->hstream) cudaStreamDestroy(ctx->hstream);
  for (cudaStream_t hs : ctx->hstreams)
    if (hs) cudaStreamDestroy(hs);
  delete ctx;
}

// ---------------------------------------------------------------------------------------------
// parameters
// ---------------------------------------------------------------------------------------------
extern "C" int l2hmc_set_net(l2hmc_ctx *ctx, int net_id, const l2hmc_net_params *p) {
  if (!ctx || !p) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_net: null argument");
  if (net_id != L2HMC_XNET && net_id != L2HMC_VNET) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_net: bad net id %d", net_id);
  if (ctx->sh.hmc) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_net: context is hmc (nets are zero)");
  const float *ptrs[] = {p->W1, p->b1, p->W2, p->b2, p->W3, p->b3, p->W4, p->b4, p->Ws, p->bs,
                         p->Wt, p->bt, p->Wq, p->bq, p->scale_s, p->scale_q};
  for (const float *q : ptrs)
    if (!q) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_net: null weight pointer");
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  const Shape &sh = ctx->sh;
  const int D = sh.D, H = sh.H, DP = sh.DP, HP = sh.HP, T = sh.T;

  // ---- raw copy (reference layout) ------------------------------------------------------------
  const size_t sizes[] = {(size_t)D * H, (size_t)H, (size_t)D * H, (size_t)H, (size_t)2 * H, (size_t)H,
                          (size_t)H * H, (size_t)H, (size_t)H * D, (size_t)D, (size_t)H * D, (size_t)D,
                          (size_t)H * D, (size_t)D, (size_t)D, (size_t)D};
  size_t total = 0;
  for (size_t s : sizes) total += round_up((int)s, 4);
  std::vector<float> raw(total, 0.f);
  size_t off[16], o = 0;
  for (int i = 0; i < 16; ++i) {
    off[i] = o;
    memcpy(raw.data() + o, ptrs[i], sizes[i] * sizeof(float));
    o += round_up((int)sizes[i], 4);
  }
  for (float f : raw)
    if (!isfinite(f)) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_net: non-finite weight");
  int rc = ensure(ctx, ctx->net_raw[net_id], total);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaMemcpy(ctx->net_raw[net_id].p, raw.data(), total * sizeof(float), cudaMemcpyHostToDevice));
  {
    const float *b = ctx->net_raw[net_id].p;
    NetRaw &r = ctx->net_rawv[net_id];
    r.W1 = b + off[0]; r.b1 = b + off[1]; r.W2 = b + off[2]; r.b2 = b + off[3];
    r.W3 = b + off[4]; r.b3 = b + off[5]; r.W4 = b + off[6]; r.b4 = b + off[7];
    r.Ws = b + off[8]; r.bs = b + off[9]; r.Wt = b + off[10]; r.bt = b + off[11];
    r.Wq = b + off[12]; r.bq = b + off[13]; r.ls = b + off[14]; r.lq = b + off[15];
  }

  // ---- layered-engine layout (any shape) ----------------------------------------------------------
  rc = lay_pack_net(ctx, net_id, p);
  if (rc) return rc;
  if (sh.HP > sh.LDE || 3 * sh.DP > sh.LDH) {  // beyond the fused kernels' layouts: layered engine only
    ctx->net_set[net_id] = true;
    return L2HMC_OK;
  }

  // ---- packed copy (tile-kernel layout, see NetDev) ---------------------------------------------
  const int LDE = sh.LDE, LDH = sh.LDH;
  const size_t nWemb = (size_t)2 * DP * LDE, ntb = (size_t)T * LDE, nW4 = (size_t)HP * LDE, nb4 = LDE,
               nWh = (size_t)HP * LDH, nbh = LDH, nes = DP, neq = DP;
  const size_t ptotal = nWemb + ntb + nW4 + nb4 + nWh + nbh + nes + neq;
  std::vector<float> pk(ptotal, 0.f);
  float *Wemb = pk.data(), *tb = Wemb + nWemb, *W4 = tb + ntb, *b4 = W4 + nW4, *Wh = b4 + nb4, *bh = Wh + nWh,
        *es = bh + nbh, *eq = es + nes;
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < H; ++j) {
      Wemb[(size_t)i * LDE + j] = p->W1[(size_t)i * H + j];
      Wemb[(size_t)(DP + i) * LDE + j] = p->W2[(size_t)i * H + j];
    }
  for (int t = 0; t < T; ++t) {
    const float arg = 6.2831855f * (float)t / (float)T;  // utils/dynamics.py:99-105 in fp32
    const float ct = cosf(arg), st = sinf(arg);
    for (int j = 0; j < H; ++j) {
      const float e3 = fmaf(st, p->W3[H + j], ct * p->W3[j]) + p->b3[j];
      tb[(size_t)t * LDE + j] = (p->b1[j] + p->b2[j]) + e3;
    }
  }
  for (int i = 0; i < H; ++i)
    for (int j = 0; j < H; ++j) W4[(size_t)i * LDE + j] = p->W4[(size_t)i * H + j];
  for (int j = 0; j < H; ++j) b4[j] = p->b4[j];
  for (int i = 0; i < H; ++i)
    for (int d = 0; d < D; ++d) {
      Wh[(size_t)i * LDH + 3 * d + 0] = p->Ws[(size_t)i * D + d];
      Wh[(size_t)i * LDH + 3 * d + 1] = p->Wt[(size_t)i * D + d];
      Wh[(size_t)i * LDH + 3 * d + 2] = p->Wq[(size_t)i * D + d];
    }
  for (int d = 0; d < DP; ++d) {
    es[d] = eq[d] = 1.0f;
    if (d < D) {
      bh[3 * d + 0] = p->bs[d];
      bh[3 * d + 1] = p->bt[d];
      bh[3 * d + 2] = p->bq[d];
      es[d] = expf(p->scale_s[d]);  // utils/layers.py:84
      eq[d] = expf(p->scale_q[d]);
    }
  }
  rc = ensure(ctx, ctx->net_packed[net_id], ptotal);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaMemcpy(ctx->net_packed[net_id].p, pk.data(), ptotal * sizeof(float), cudaMemcpyHostToDevice));
  {
    const float *b = ctx->net_packed[net_id].p;
    NetDev &n = ctx->net_dev[net_id];
    n.Wemb = b; n.tb = n.Wemb + nWemb; n.W4 = n.tb + ntb; n.b4 = n.W4 + nW4;
    n.Wh = n.b4 + nb4; n.bh = n.Wh + nWh; n.es = n.bh + nbh; n.eq = n.es + nes;
  }
  if (ctx->tc_ok) {
    rc = tc_pack_net(ctx, net_id, p);
    if (rc) return rc;
  }
  ctx->net_set[net_id] = true;
  return L2HMC_OK;
}

extern "C" int l2hmc_set_masks(l2hmc_ctx *ctx, const float *mask) {
  if (!ctx || !mask) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_masks: null argument");
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  const Shape &sh = ctx->sh;
  std::vector<float> m((size_t)sh.T * sh.DP, 0.f);
  for (int t = 0; t < sh.T; ++t)
    for (int d = 0; d < sh.D; ++d) {
      const float v = mask[(size_t)t * sh.D + d];
      if (v != 0.f && v != 1.f) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_masks: mask[%d,%d] = %g is not 0/1", t, d, v);
      m[(size_t)t * sh.DP + d] = v;
    }
  int rc = ensure(ctx, ctx->mask, m.size());
  if (rc) return rc;
  CUDA_TRY(ctx, cudaMemcpy(ctx->mask.p, m.data(), m.size() * sizeof(float), cudaMemcpyHostToDevice));
  ctx->mask_set = true;
  return L2HMC_OK;
}

extern "C" int l2hmc_status_flags(l2hmc_ctx *ctx, uint32_t *flags, int clear) {
  if (!ctx || !flags) return fail(ctx, L2HMC_EINVAL, "l2hmc_status_flags: null argument");
  *flags = *(volatile unsigned int *)ctx->status_h;  // pinned host memory the kernels write with system-scope atomics
  if (clear) {
    *(volatile unsigned int *)ctx->status_h = 0u;
    const char *fe = getenv("L2HMC_TC_F16");
    ctx->td.f16 = (fe && fe[0] == '0') ? 0 : 1;  // the tensor-core kernel may try the fp16 operand split again
  }
  return L2HMC_OK;
}

extern "C" int l2hmc_set_eps(l2hmc_ctx *ctx, float eps) {
  if (!ctx) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_eps: null context");
  if (!(eps > 0.f) || !isfinite(eps)) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_eps: eps must be finite and > 0");
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));  // tc_pack_hc copies to this context's device
  ctx->sh.eps = eps;
  for (int net_id = 0; net_id < 2; ++net_id) {
    int rc = tc_pack_hc(ctx, net_id);
    if (rc) return rc;
  }
  return L2HMC_OK;
}

extern "C" int l2hmc_set_likelihood_scale(l2hmc_ctx *ctx, float beta) {
  if (!ctx) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_likelihood_scale: null context");
  if (!(beta >= 0.f) || !isfinite(beta)) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_likelihood_scale: must be finite and >= 0");
  ctx->lay.like_scale = beta;
  return L2HMC_OK;
}

extern "C" int l2hmc_set_temperature(l2hmc_ctx *ctx, float t) {
  if (!ctx) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_temperature: null context");
  if (!(t > 0.f) || !isfinite(t)) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_temperature: must be finite and > 0");
  ctx->en.temperature = t;
  return L2HMC_OK;
}

// one closed-form energy into a device buffer: fills kind-specific fields of `out` (ncomp, mu, Ssym, logc, s0, s1)
struct EnergyOne {
  int kind = -1, ncomp = 1;
  const float *mu = nullptr, *Ssym = nullptr, *logc = nullptr;
  float s0 = 0.f, s1 = 0.f;
  size_t nmu = 0;  // offset of Ssym inside the packed host image (for the tensor-core Gaussian image)
};

static int pack_energy_one(l2hmc_ctx *ctx, const char *who, int kind, int n_comp, const float *mu, const float *S, const float *logc,
                           const float *scalars, int n_scalars, DevBuf &dbuf, EnergyOne *out, std::vector<float> *host_image) {
  const Shape &sh = ctx->sh;
  EnergyOne e;
  e.kind = kind;
  if (kind == L2HMC_ENERGY_GAUSSIAN || kind == L2HMC_ENERGY_GMM) {
    if (kind == L2HMC_ENERGY_GAUSSIAN) n_comp = 1;
    if (n_comp < 1 || n_comp > MAX_COMP) return fail(ctx, L2HMC_EUNSUPPORTED, "%s: 1 <= n_comp <= %d", who, MAX_COMP);
    if (!mu || !S || (kind == L2HMC_ENERGY_GMM && !logc)) return fail(ctx, L2HMC_EINVAL, "%s: null parameter array", who);
    const size_t nmu = (size_t)n_comp * sh.DP, nS = (size_t)n_comp * sh.DP * sh.LDS, nc = round_up(n_comp, 4);
    std::vector<float> buf(nmu + nS + nc, 0.f);
    for (int c = 0; c < n_comp; ++c) {
      for (int i = 0; i < sh.D; ++i) buf[(size_t)c * sh.DP + i] = mu[(size_t)c * sh.D + i];
      const float *Sc = S + (size_t)c * sh.D * sh.D;
      for (int i = 0; i < sh.D; ++i)
        for (int j = 0; j < sh.D; ++j)  // 0.5 * (d S^T) + 0.5 * (d S) == d (0.5 S^T + 0.5 S)
          buf[nmu + ((size_t)c * sh.DP + i) * sh.LDS + j] = 0.5f * Sc[(size_t)i * sh.D + j] + 0.5f * Sc[(size_t)j * sh.D + i];
      if (logc) buf[nmu + nS + c] = logc[c];
    }
    for (float f : buf)
      if (!isfinite(f)) return fail(ctx, L2HMC_EINVAL, "%s: non-finite parameter", who);
    int rc = ensure(ctx, dbuf, buf.size());
    if (rc) return rc;
    CUDA_TRY(ctx, cudaMemcpy(dbuf.p, buf.data(), buf.size() * sizeof(float), cudaMemcpyHostToDevice));
    e.ncomp = n_comp;
    e.mu = dbuf.p;
    e.Ssym = e.mu + nmu;
    e.logc = e.Ssym + nS;
    e.nmu = nmu;
    if (host_image) host_image->swap(buf);
  } else if (kind == L2HMC_ENERGY_ROUGHWELL || kind == L2HMC_ENERGY_FUNNEL) {
    if (!scalars || n_scalars < 2) return fail(ctx, L2HMC_EINVAL, "%s: need 2 scalars", who);
    e.s0 = scalars[0];
    e.s1 = scalars[1];
    if (kind == L2HMC_ENERGY_FUNNEL && sh.D < 2) return fail(ctx, L2HMC_EINVAL, "%s: funnel needs x_dim >= 2", who);
  } else {
    return fail(ctx, L2HMC_EUNSUPPORTED, "%s: unknown energy kind %d", who, kind);
  }
  *out = e;
  return L2HMC_OK;
}

extern "C" int l2hmc_set_energy(l2hmc_ctx *ctx, int kind, int n_comp, const float *mu, const float *S,
                                const float *logc, const float *scalars, int n_scalars) {
  if (!ctx) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_energy: null context");
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  EnergyOne e;
  std::vector<float> img;
  int rc = pack_energy_one(ctx, "l2hmc_set_energy", kind, n_comp, mu, S, logc, scalars, n_scalars, ctx->energy_buf, &e, &img);
  if (rc) return rc;
  if (ctx->tc_ok && kind == L2HMC_ENERGY_GAUSSIAN) {
    rc = tc_pack_gaussian(ctx, img.data() + e.nmu);
    if (rc) return rc;
  }
  EnergyDev en = ctx->en;
  en.kind = e.kind; en.ncomp = e.ncomp; en.mu = e.mu; en.Ssym = e.Ssym; en.logc = e.logc; en.s0 = e.s0; en.s1 = e.s1;
  ctx->en = en;
  ctx->energy_set = true;
  {
    int k = ctx->kernel;
    if (resolve_kernel(ctx, &k) == L2HMC_OK) ctx->kernel = k;  // AUTO may now pick the tensor-core kernel
  }
  return L2HMC_OK;
}

extern "C" int l2hmc_set_energy_mixed(l2hmc_ctx *ctx, const l2hmc_energy_desc *a, const l2hmc_energy_desc *b, float beta) {
  if (!ctx || !a || !b) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_energy_mixed: null argument");
  if (!(beta >= 0.f && beta <= 1.f)) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_energy_mixed: beta must be in [0, 1]");
  if (ctx->sh.DP > MIX_MAXD) return fail(ctx, L2HMC_EUNSUPPORTED, "l2hmc_set_energy_mixed: x_dim <= %d", MIX_MAXD);
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  EnergyOne ea, eb;
  int rc = pack_energy_one(ctx, "l2hmc_set_energy_mixed (a)", a->kind, a->n_comp, a->mu, a->S, a->logc, a->scalars, a->n_scalars,
                           ctx->energy_buf, &ea, nullptr);
  if (rc) return rc;
  rc = pack_energy_one(ctx, "l2hmc_set_energy_mixed (b)", b->kind, b->n_comp, b->mu, b->S, b->logc, b->scalars, b->n_scalars,
                       ctx->energy_buf2, &eb, nullptr);
  if (rc) return rc;
  EnergyDev en = ctx->en;
  en.kind = L2HMC_ENERGY_MIXED;
  en.kind_a = ea.kind; en.ncomp = ea.ncomp; en.mu = ea.mu; en.Ssym = ea.Ssym; en.logc = ea.logc; en.s0 = ea.s0; en.s1 = ea.s1;
  en.kind_b = eb.kind; en.ncomp_b = eb.ncomp; en.mu_b = eb.mu; en.Ssym_b = eb.Ssym; en.logc_b = eb.logc; en.s0_b = eb.s0; en.s1_b = eb.s1;
  en.beta = beta;
  ctx->en = en;
  ctx->energy_set = true;
  int k = ctx->kernel;
  rc = resolve_kernel(ctx, &k);
  if (rc) return rc;
  ctx->kernel = k;
  return L2HMC_OK;
}

extern "C" int l2hmc_set_mix_beta(l2hmc_ctx *ctx, float beta) {
  if (!ctx) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_mix_beta: null context");
  if (!ctx->energy_set || ctx->en.kind != L2HMC_ENERGY_MIXED) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_mix_beta: the energy is not a mixed one");
  if (!(beta >= 0.f && beta <= 1.f)) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_mix_beta: beta must be in [0, 1]");
  ctx->en.beta = beta;
  return L2HMC_OK;
}

static int set_aux_dim(l2hmc_ctx *ctx, int aux_dim, const char *who) {
  layered::LayDims &dm = ctx->lay.dm;
  const bool other_uses = (std::string(who) == "l2hmc_set_energy_decoder") ? ctx->lay.enc.n_layers > 0
                                                                            : (ctx->energy_set && ctx->en.kind == L2HMC_ENERGY_DECODER);
  if (other_uses && dm.aux != aux_dim)
    return fail(ctx, L2HMC_EINVAL, "%s: aux_dim %d differs from the one already configured (%d)", who, aux_dim, dm.aux);
  dm.aux = aux_dim;
  dm.auxp = round_up(aux_dim, 8);
  return L2HMC_OK;
}

extern "C" int l2hmc_set_energy_decoder(l2hmc_ctx *ctx, int n_layers, const int32_t *widths, const float *const *W,
                                        const float *const *b) {
  if (!ctx) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_energy_decoder: null context");
  if (ctx->cfg.kernel != L2HMC_KERNEL_AUTO && ctx->cfg.kernel != L2HMC_KERNEL_LAYERED && ctx->cfg.kernel != L2HMC_KERNEL_LAYERED_FMA)
    return fail(ctx, L2HMC_EUNSUPPORTED, "l2hmc_set_energy_decoder: the decoder energy runs on the layered engine only");
  if (!widths || n_layers < 1) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_energy_decoder: bad argument");
  if (widths[0] != ctx->sh.D)
    return fail(ctx, L2HMC_EINVAL, "l2hmc_set_energy_decoder: first width %d is not x_dim %d", widths[0], ctx->sh.D);
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  int rc = set_aux_dim(ctx, widths[n_layers], "l2hmc_set_energy_decoder");
  if (rc) return rc;
  if ((rc = lay_pack_mlp(ctx, ctx->lay.dec, n_layers, widths, W, b, "l2hmc_set_energy_decoder"))) return rc;
  EnergyDev en = ctx->en;
  en.kind = L2HMC_ENERGY_DECODER;
  en.ncomp = 1;
  en.mu = en.Ssym = en.logc = nullptr;
  en.s0 = en.s1 = 0.f;
  ctx->en = en;
  ctx->energy_set = true;
  int k = ctx->kernel;
  if ((rc = resolve_kernel(ctx, &k))) return rc;
  ctx->kernel = k;
  return L2HMC_OK;
}

extern "C" int l2hmc_set_aux_encoder(l2hmc_ctx *ctx, int n_layers, const int32_t *widths, const float *const *W,
                                     const float *const *b) {
  if (!ctx) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_aux_encoder: null context");
  if (n_layers == 0) {
    ctx->lay.enc.n_layers = 0;
    int k = ctx->kernel;
    if (resolve_kernel(ctx, &k) == L2HMC_OK) ctx->kernel = k;
    return L2HMC_OK;
  }
  if (ctx->sh.hmc) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_aux_encoder: context is hmc (nets are zero)");
  if (ctx->cfg.kernel != L2HMC_KERNEL_AUTO && ctx->cfg.kernel != L2HMC_KERNEL_LAYERED && ctx->cfg.kernel != L2HMC_KERNEL_LAYERED_FMA)
    return fail(ctx, L2HMC_EUNSUPPORTED, "l2hmc_set_aux_encoder: aux-conditioned nets run on the layered engine only");
  if (!widths || n_layers < 1) return fail(ctx, L2HMC_EINVAL, "l2hmc_set_aux_encoder: bad argument");
  if (widths[n_layers] != ctx->sh.H)
    return fail(ctx, L2HMC_EINVAL, "l2hmc_set_aux_encoder: last width %d is not the net width %d", widths[n_layers], ctx->sh.H);
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  int rc = set_aux_dim(ctx, widths[0], "l2hmc_set_aux_encoder");
  if (rc) return rc;
  if ((rc = lay_pack_mlp(ctx, ctx->lay.enc, n_layers, widths, W, b, "l2hmc_set_aux_encoder"))) return rc;
  int k = ctx->kernel;
  if ((rc = resolve_kernel(ctx, &k))) return rc;
  ctx->kernel = k;
  return L2HMC_OK;
}

extern "C" int l2hmc_bind_aux(l2hmc_ctx *ctx, int64_t n, const float *aux) {
  if (!ctx) return fail(ctx, L2HMC_EINVAL, "l2hmc_bind_aux: null context");
  if (aux && n < 0) return fail(ctx, L2HMC_EINVAL, "l2hmc_bind_aux: n < 0");
  ctx->lay.aux = aux;
  ctx->lay.aux_n = aux ? n : 0;
  return L2HMC_OK;
}

// ---------------------------------------------------------------------------------------------
// hot path
// ---------------------------------------------------------------------------------------------
static int check_ready(l2hmc_ctx *ctx, const char *who) {
  if (!ctx) return fail(ctx, L2HMC_EINVAL, "%s: null context", who);
  if (!ctx->energy_set) return fail(ctx, L2HMC_EINVAL, "%s: energy not set (l2hmc_set_energy)", who);
  return L2HMC_OK;
}

static int validate_transition(l2hmc_ctx *ctx, const l2hmc_transition_args *a, bool host = false) {
  int rc = check_ready(ctx, "l2hmc_transition");
  if (rc) return rc;
  if (!a) return fail(ctx, L2HMC_EINVAL, "l2hmc_transition: null args");
  if (!ctx->mask_set) return fail(ctx, L2HMC_EINVAL, "l2hmc_transition: masks not set (l2hmc_set_masks)");
  if (!ctx->sh.hmc && !(ctx->net_set[0] && ctx->net_set[1]))
    return fail(ctx, L2HMC_EINVAL, "l2hmc_transition: XNet/VNet not set (l2hmc_set_net)");
  if (a->n < 0) return fail(ctx, L2HMC_EINVAL, "l2hmc_transition: n < 0");
  if (a->n_transitions < 1) return fail(ctx, L2HMC_EINVAL, "l2hmc_transition: n_transitions must be >= 1");
  if (a->n > 0 && (!a->x || (!host && (!a->x_out || !a->px_out))))
    return fail(ctx, L2HMC_EINVAL, "l2hmc_transition: x, x_out and px_out are required");
  if (a->dir_mode < 0 || a->dir_mode > 3) return fail(ctx, L2HMC_EINVAL, "l2hmc_transition: bad dir_mode %d", a->dir_mode);
  if (a->dir_mode == L2HMC_DIR_PER_CHAIN && !a->dir && a->n > 0) return fail(ctx, L2HMC_EINVAL, "l2hmc_transition: dir_mode PER_CHAIN needs dir");
  if (a->do_mh && !a->x_next && a->n > 0 && !host) return fail(ctx, L2HMC_EINVAL, "l2hmc_transition: do_mh needs x_next");
  if (a->n_transitions > 1 && !a->do_mh && !a->chain) return fail(ctx, L2HMC_EINVAL, "l2hmc_transition: n_transitions > 1 needs do_mh");
  if (a->chain && a->trace) return fail(ctx, L2HMC_EINVAL, "l2hmc_transition: chain mode has one Metropolis step: no per-transition trace");
  if (a->chain && host) return fail(ctx, L2HMC_EINVAL, "l2hmc_transition_host: chain mode is a device-resident call (use l2hmc_transition)");
  if (a->trace && !a->do_mh) return fail(ctx, L2HMC_EINVAL, "l2hmc_transition: trace records the Metropolis output and needs do_mh");
  if (a->trace && host) return fail(ctx, L2HMC_EINVAL, "l2hmc_transition_host: trace is a device-resident output (use l2hmc_transition)");
  if (a->counter + (uint64_t)a->n_transitions >= (1ull << 30)) return fail(ctx, L2HMC_EINVAL, "l2hmc_transition: counter must stay below 2^30");
  return L2HMC_OK;
}

static int launch_transition(l2hmc_ctx *ctx, const l2hmc_transition_args *a, cudaStream_t stream) {
  if (a->n == 0) return L2HMC_OK;
  KernelArgs K;
  K.sh = ctx->sh;
  K.xnet = ctx->net_dev[0];
  K.vnet = ctx->net_dev[1];
  K.en = ctx->en;
  K.mask = ctx->mask.p;
  TransitionIO &io = K.io;
  io.n = a->n; io.chain_offset = a->chain_offset;
  io.x = a->x; io.v = a->v; io.u = a->u; io.dir = a->dir;
  io.dir_mode = a->dir_mode; io.log_jac = a->log_jac; io.do_mh = a->do_mh; io.n_transitions = a->n_transitions;
  io.seed = a->seed; io.counter = a->counter;
  io.x_out = a->x_out; io.v_out = a->v_out; io.px_out = a->px_out; io.x_next = a->x_next; io.accepted = a->accepted;
  io.stats = a->stats; io.trace = a->trace; io.status = ctx->status_d;
  io.chain = a->chain ? 1 : 0; io.v0 = a->v0;

  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx->timing) {
    if (ctx->ev_used + 2 > ctx->ev.size()) {
      for (int i = 0; i < 2; ++i) {
        cudaEvent_t e;
        CUDA_TRY(ctx, cudaEventCreate(&e));
        ctx->ev.push_back(e);
      }
    }
    e0 = ctx->ev[ctx->ev_used];
    e1 = ctx->ev[ctx->ev_used + 1];
    ctx->ev_used += 2;
    CUDA_TRY(ctx, cudaEventRecord(e0, stream));
  }
  int kernel = 0;
  {
    int rc = resolve_kernel(ctx, &kernel);  // the energy kind may have been set after l2hmc_create
    if (rc) return rc;
    ctx->kernel = kernel;
  }
  if (kernel == L2HMC_KERNEL_TC) {
    tc::TcArgs TA;
    TA.sh = ctx->sh;
    TA.td = ctx->td;
    TA.xnet = ctx->tc_net[0];
    TA.vnet = ctx->tc_net[1];
    TA.gimg = ctx->tc_gbuf.p;
    TA.gimg_h = ctx->tc_gbuf.p ? ctx->tc_gbuf.p + ctx->tc_g_h_off : nullptr;
    TA.en = ctx->en;
    TA.mask = ctx->mask.p;
    TA.io = K.io;
    // shape-specialised compute path (kernel_tc_s.cuh) where an instantiation exists; L2HMC_TC_GENERIC=1 forces the generic one
    const int nqc = ctx->sh.DP / 4, nhc = ctx->td.HK / 8;
    const char *ge = getenv("L2HMC_TC_GENERIC");
    // fixed-shape instantiations for the benchmark shapes, the run-time-shape instantiation for everything else the
    // kernel's TMEM map holds (x_dim <= 52, width <= 104)
    const bool fixed_shape = (nqc == 13 && nhc == 13) || (nqc == 8 && nhc == 13);
    const bool spec = !(ge && ge[0] == '1') && ctx->td.nq == 2 && ctx->tc_net[0].hc && ctx->tc_net[1].hc &&
                      (fixed_shape || tc::tc_s_shape_fits(nqc, nhc));
    const unsigned blocks = (unsigned)((a->n + tc::MT - 1) / tc::MT);
    if (spec) {
      const long long state_bytes = (long long)tc::make_tclay_s(tc::tc_s_row_stride(ctx->sh.DP), ctx->sh.DP, ctx->sh.T).ring * 4;
      {  // ring slot = KSLOT K steps of the widest GEMM of this kernel's schedule (the heads are split in two)
        const int ca = (nqc + 1) / 2, n3a = round_up(12 * ca, 16);
        int nmax = ctx->td.N1 > n3a ? ctx->td.N1 : n3a;
        if (ctx->td.NG > nmax) nmax = ctx->td.NG;
        TA.td.slot_floats = tc::KSLOT_S * 16 * nmax;
      }
      long long ns = (232448LL - 1024 - state_bytes) / ((long long)TA.td.slot_floats * 4);
      TA.td.nslot = ns > tc::MAX_SLOT ? tc::MAX_SLOT : (int)ns;
      if (TA.td.nslot < 4) return fail(ctx, L2HMC_EINVAL, "tensor-core kernel: shared-memory ring too small for this shape");
      const size_t smem = tc::tc_s_smem_bytes(ctx->sh.DP, ctx->sh.T, TA.td.nslot, TA.td.slot_floats);
      size_t &tc_s_configured = ctx->smem_cfg[0];  // the attribute is per device: remembered per context, not per thread
      if (smem > tc_s_configured) {
#define L2HMC_TC_S_ATTR(Q, HH, F, B, X) \
  CUDA_TRY(ctx, cudaFuncSetAttribute(tc::tc_transition_kernel_s<Q, HH, F, B, X>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))
        L2HMC_TC_S_ATTR(13, 13, true, true, true); L2HMC_TC_S_ATTR(13, 13, true, true, false);
        L2HMC_TC_S_ATTR(13, 13, false, true, true); L2HMC_TC_S_ATTR(13, 13, false, true, false);
        L2HMC_TC_S_ATTR(13, 13, true, false, true); L2HMC_TC_S_ATTR(13, 13, true, false, false);
        L2HMC_TC_S_ATTR(13, 13, false, false, true); L2HMC_TC_S_ATTR(13, 13, false, false, false);
        L2HMC_TC_S_ATTR(8, 13, true, false, true); L2HMC_TC_S_ATTR(8, 13, true, false, false);
        L2HMC_TC_S_ATTR(8, 13, false, false, true); L2HMC_TC_S_ATTR(8, 13, false, false, false);
        L2HMC_TC_S_ATTR(8, 13, true, true, true); L2HMC_TC_S_ATTR(8, 13, true, true, false);
        L2HMC_TC_S_ATTR(8, 13, false, true, true); L2HMC_TC_S_ATTR(8, 13, false, true, false);
        L2HMC_TC_S_ATTR(0, 0, true, false, true); L2HMC_TC_S_ATTR(0, 0, true, false, false);
        L2HMC_TC_S_ATTR(0, 0, false, false, true); L2HMC_TC_S_ATTR(0, 0, false, false, false);
        L2HMC_TC_S_ATTR(0, 0, true, true, true); L2HMC_TC_S_ATTR(0, 0, true, true, false);
        L2HMC_TC_S_ATTR(0, 0, false, true, true); L2HMC_TC_S_ATTR(0, 0, false, true, false);
#undef L2HMC_TC_S_ATTR
        tc_s_configured = smem;
      }
      const unsigned nthreads = (unsigned)tc::TC_S_THREADS;  // 8 compute warps + MMA issuer + TMA producer + 2 idle (whole warpgroups: setmaxnreg)
      const bool fm = ctx->td.fast_math != 0, bg = ctx->td.biasg != 0;
      // fp16 split only when everything packed (weights, biases, time-embedding rows, precision matrix) is well inside
      // the fp16 range; activations are checked by the kernel (sticky flag, l2hmc_debug_counters[23])
      const float wmax = fmaxf(fmaxf(ctx->tc_wmax[0], ctx->tc_wmax[1]), ctx->en.kind == L2HMC_ENERGY_GAUSSIAN ? ctx->tc_wmax[2] : 0.f);
      // ... and an earlier launch of THIS context that met an out-of-range activation (sticky status bit, read from pinned
      // host memory without synchronising) puts the context on the tf32 split for good
      if (*(volatile unsigned int *)ctx->status_h & STATUS_F16_RANGE) ctx->td.f16 = 0;
      const bool h16 = ctx->td.f16 != 0 && wmax < 3.0e4f;
      TA.td.f16 = h16 ? 1 : 0;
      ctx->tc_used_f16 = h16;
#define L2HMC_TC_S_LAUNCH(Q, HH, F, B, X) tc::tc_transition_kernel_s<Q, HH, F, B, X><<<blocks, nthreads, smem, stream>>>(TA)
      if (nqc == 13 && nhc == 13) {
        if (bg) {
          if (fm) { if (h16) L2HMC_TC_S_LAUNCH(13, 13, true, true, true); else L2HMC_TC_S_LAUNCH(13, 13, true, true, false); }
          else { if (h16) L2HMC_TC_S_LAUNCH(13, 13, false, true, true); else L2HMC_TC_S_LAUNCH(13, 13, false, true, false); }
        } else {
          if (fm) { if (h16) L2HMC_TC_S_LAUNCH(13, 13, true, false, true); else L2HMC_TC_S_LAUNCH(13, 13, true, false, false); }
          else { if (h16) L2HMC_TC_S_LAUNCH(13, 13, false, false, true); else L2HMC_TC_S_LAUNCH(13, 13, false, false, false); }
        }
      } else if (nqc == 8 && nhc == 13) {
        if (bg) {
          if (fm) { if (h16) L2HMC_TC_S_LAUNCH(8, 13, true, true, true); else L2HMC_TC_S_LAUNCH(8, 13, true, true, false); }
          else { if (h16) L2HMC_TC_S_LAUNCH(8, 13, false, true, true); else L2HMC_TC_S_LAUNCH(8, 13, false, true, false); }
        } else {
          if (fm) { if (h16) L2HMC_TC_S_LAUNCH(8, 13, true, false, true); else L2HMC_TC_S_LAUNCH(8, 13, true, false, false); }
          else { if (h16) L2HMC_TC_S_LAUNCH(8, 13, false, false, true); else L2HMC_TC_S_LAUNCH(8, 13, false, false, false); }
        }
      } else {  // chunk counts read from the arguments
        if (bg) {
          if (fm) { if (h16) L2HMC_TC_S_LAUNCH(0, 0, true, true, true); else L2HMC_TC_S_LAUNCH(0, 0, true, true, false); }
          else { if (h16) L2HMC_TC_S_LAUNCH(0, 0, false, true, true); else L2HMC_TC_S_LAUNCH(0, 0, false, true, false); }
        } else {
          if (fm) { if (h16) L2HMC_TC_S_LAUNCH(0, 0, true, false, true); else L2HMC_TC_S_LAUNCH(0, 0, true, false, false); }
          else { if (h16) L2HMC_TC_S_LAUNCH(0, 0, false, false, true); else L2HMC_TC_S_LAUNCH(0, 0, false, false, false); }
        }
      }
#undef L2HMC_TC_S_LAUNCH
    } else {
    if (a->chain) return fail(ctx, L2HMC_EUNSUPPORTED, "chain mode: not in the generic tensor-core kernel (compose l2hmc_transition calls with log_jac = 1)");
    ctx->tc_used_f16 = false;
    const size_t smem = tc::tc_smem_bytes(ctx->sh.DP, ctx->sh.T, ctx->td.nslot, ctx->td.slot_floats);
    size_t &tc_configured = ctx->smem_cfg[1];
    if (smem > tc_configured) {
      CUDA_TRY(ctx, cudaFuncSetAttribute(tc::tc_transition_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CUDA_TRY(ctx, cudaFuncSetAttribute(tc::tc_transition_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CUDA_TRY(ctx, cudaFuncSetAttribute(tc::tc_transition_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CUDA_TRY(ctx, cudaFuncSetAttribute(tc::tc_transition_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CUDA_TRY(ctx, cudaFuncSetAttribute(tc::tc_transition_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CUDA_TRY(ctx, cudaFuncSetAttribute(tc::tc_transition_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      tc_configured = smem;
    }
    const unsigned nthreads = (unsigned)(tc::MT * ctx->td.nq + 64);
    if (ctx->td.nq == 2) {
      if (ctx->td.fast_math) tc::tc_transition_kernel<2, true><<<blocks, nthreads, smem, stream>>>(TA);
      else tc::tc_transition_kernel<2, false><<<blocks, nthreads, smem, stream>>>(TA);
    } else if (ctx->td.nq == 3) {
      if (ctx->td.fast_math) tc::tc_transition_kernel<3, true><<<blocks, nthreads, smem, stream>>>(TA);
      else tc::tc_transition_kernel<3, false><<<blocks, nthreads, smem, stream>>>(TA);
    } else {
      if (ctx->td.fast_math) tc::tc_transition_kernel<4, true><<<blocks, nthreads, smem, stream>>>(TA);
      else tc::tc_transition_kernel<4, false><<<blocks, nthreads, smem, stream>>>(TA);
    }
    }
  } else if (kernel == L2HMC_KERNEL_SMALL) {
    small::SmallArgs SA;
    SA.sh = ctx->sh;
    SA.xnet = ctx->net_rawv[0];
    SA.vnet = ctx->net_rawv[1];
    SA.tbx = ctx->net_dev[0].tb;
    SA.tbv = ctx->net_dev[1].tb;
    SA.en = ctx->en;
    SA.mask = ctx->mask.p;
    SA.io = K.io;
    const unsigned blocks = (unsigned)((a->n + small::NT - 1) / small::NT);
    // exp / tanh of the updates: ex2.approx / rcp.approx like the tensor-core epilogues (L2HMC_SMALL_FAST_MATH=0: expf / tanhf)
    const char *sfm = getenv("L2HMC_SMALL_FAST_MATH");
    const bool sfast = !(sfm && sfm[0] == '0');
    if (ctx->sh.D <= 2 && (ctx->sh.hmc || ctx->sh.H <= 10)) {
      if (sfast) small::small_transition_kernel<2, 10, true><<<blocks, small::NT, small::small_smem_bytes<2, 10>(ctx->sh.T), stream>>>(SA);
      else small::small_transition_kernel<2, 10, false><<<blocks, small::NT, small::small_smem_bytes<2, 10>(ctx->sh.T), stream>>>(SA);
    } else {
      if (sfast) small::small_transition_kernel<4, 16, true><<<blocks, small::NT, small::small_smem_bytes<4, 16>(ctx->sh.T), stream>>>(SA);
      else small::small_transition_kernel<4, 16, false><<<blocks, small::NT, small::small_smem_bytes<4, 16>(ctx->sh.T), stream>>>(SA);
    }
  } else if (kernel == L2HMC_KERNEL_TILE) {
    const size_t smem = tile::smem_bytes(ctx->sh.DP, ctx->sh.HP, ctx->sh.T);
    size_t &configured = ctx->smem_cfg[2];
    if (smem > configured) {
      CUDA_TRY(ctx, cudaFuncSetAttribute(tile::transition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    const long long blocks = (a->n + tile::M - 1) / tile::M;
    tile::transition_kernel<<<(unsigned)blocks, tile::NT, smem, stream>>>(K);
  } else if (kernel == L2HMC_KERNEL_LAYERED) {
    if (a->chain) return fail(ctx, L2HMC_EUNSUPPORTED, "chain mode: not in the layered engine (compose l2hmc_transition calls with log_jac = 1)");
    int rc = launch_layered(ctx, a, K.io, stream);  // counts its own launches
    if (rc) return rc;
    ctx->launches -= 1;
  } else {
    return fail(ctx, L2HMC_EUNSUPPORTED, "l2hmc_transition: kernel kind %d not available", ctx->kernel);
  }
  CUDA_TRY(ctx, cudaGetLastError());
  if (ctx->timing) CUDA_TRY(ctx, cudaEventRecord(e1, stream));
  ctx->launches += 1;
  return L2HMC_OK;
}

extern "C" int l2hmc_transition(l2hmc_ctx *ctx, const l2hmc_transition_args *a) {
  int rc = validate_transition(ctx, a);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  return launch_transition(ctx, a, (cudaStream_t)a->stream);
}

static int ensure_u8(l2hmc_ctx *ctx, uint8_t *&p, size_t &have, size_t n) {
  if (have >= n && p) return L2HMC_OK;
  if (p) cudaFree(p);
  p = nullptr;
  have = 0;
  CUDA_TRY(ctx, cudaMalloc(&p, n));
  have = n;
  return L2HMC_OK;
}

// l2hmc_transition_host for one transition of a large batch: chunks pipelined over 3 streams.
static int transition_host_chunked(l2hmc_ctx *ctx, const l2hmc_transition_args *a) {
  const size_t n = (size_t)a->n, D = (size_t)ctx->sh.D;
  int rc;
  for (int i = 0; i < 3; ++i)
    if (!ctx->hstreams[i]) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->hstreams[i], cudaStreamNonBlocking));
  if ((rc = ensure(ctx, ctx->hx, n * D))) return rc;
  if (a->v && (rc = ensure(ctx, ctx->hv, n * D))) return rc;
  if (a->u && (rc = ensure(ctx, ctx->hu, n))) return rc;
  if (a->dir && (rc = ensure_u8(ctx, ctx->hdir, ctx->hdir_n, n))) return rc;
  if ((rc = ensure(ctx, ctx->hxo, n * D))) return rc;
  if ((rc = ensure(ctx, ctx->hpx, n))) return rc;
  if (a->v_out && (rc = ensure(ctx, ctx->hvo, n * D))) return rc;
  const bool want_next = a->x_next || a->do_mh;
  if (want_next && (rc = ensure(ctx, ctx->hxn, n * D))) return rc;
  if (a->accepted && (rc = ensure_u8(ctx, ctx->hacc, ctx->hacc_n, n))) return rc;
  // chunks of two waves of 128-chain tiles (one tile per SM and wave), at least 4 chunks; the first and the last chunk are
  // one wave only: their copy in / copy out are the ones that do not hide under a kernel
  const size_t sms = ctx->lay.sms > 0 ? (size_t)ctx->lay.sms : 148;
  const size_t wave = sms * 128;
  std::vector<size_t> bounds;  // chunk c covers chains [bounds[c], bounds[c+1])
  bounds.push_back(0);
  if (n >= 8 * wave) {
    size_t lo = wave;
    bounds.push_back(lo);
    while (n - lo > 3 * wave) {
      lo += 2 * wave;
      bounds.push_back(lo);
    }
    if (n - lo > wave) {  // what is left: one more middle chunk, then a last chunk of at most one wave
      lo = n - wave;
      bounds.push_back(lo);
    }
    bounds.push_back(n);
  } else {
    const size_t per = ((n + 3) / 4 + 127) / 128 * 128;
    for (size_t lo = per; lo < n; lo += per) bounds.push_back(lo);
    bounds.push_back(n);
  }
  const int NCH = (int)bounds.size() - 1;
  double *dstats = nullptr;  // accept statistics: one device accumulator pair per stream, summed on the host at the end
  if (a->stats) {
    if ((rc = ensure(ctx, ctx->hstats, 12))) return rc;
    dstats = reinterpret_cast<double *>(ctx->hstats.p);
    ctx->host_stats0[0] = a->stats[0];
    ctx->host_stats0[1] = a->stats[1];
  }
  for (int c = 0; c < NCH; ++c) {
    const size_t lo = bounds[c];
    const size_t m = bounds[c + 1] - lo;
    cudaStream_t s = ctx->hstreams[c % 3];
    l2hmc_transition_args d = *a;
    d.stream = s;
    d.stats = dstats ? dstats + 2 * (c % 3) : nullptr;
    if (dstats && c < 3) CUDA_TRY(ctx, cudaMemsetAsync(d.stats, 0, 2 * sizeof(double), s));
    d.n = (int64_t)m;
    d.chain_offset = a->chain_offset + (int64_t)lo;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hx.p + lo * D, a->x + lo * D, m * D * sizeof(float), cudaMemcpyHostToDevice, s));
    d.x = ctx->hx.p + lo * D;
    if (a->v) {
      CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hv.p + lo * D, a->v + lo * D, m * D * sizeof(float), cudaMemcpyHostToDevice, s));
      d.v = ctx->hv.p + lo * D;
    }
    if (a->u) {
      CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hu.p + lo, a->u + lo, m * sizeof(float), cudaMemcpyHostToDevice, s));
      d.u = ctx->hu.p + lo;
    }
    if (a->dir) {
      CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hdir + lo, a->dir + lo, m, cudaMemcpyHostToDevice, s));
      d.dir = ctx->hdir + lo;
    }
    d.x_out = ctx->hxo.p + lo * D;
    d.px_out = ctx->hpx.p + lo;
    d.v_out = a->v_out ? ctx->hvo.p + lo * D : nullptr;
    d.x_next = want_next ? ctx->hxn.p + lo * D : nullptr;
    d.accepted = a->accepted ? ctx->hacc + lo : nullptr;
    if ((rc = launch_transition(ctx, &d, s))) return rc;
    if (a->x_out) CUDA_TRY(ctx, cudaMemcpyAsync(a->x_out + lo * D, d.x_out, m * D * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (a->px_out) CUDA_TRY(ctx, cudaMemcpyAsync(a->px_out + lo, d.px_out, m * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (a->v_out) CUDA_TRY(ctx, cudaMemcpyAsync(a->v_out + lo * D, d.v_out, m * D * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (a->x_next) CUDA_TRY(ctx, cudaMemcpyAsync(a->x_next + lo * D, d.x_next, m * D * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (a->accepted) CUDA_TRY(ctx, cudaMemcpyAsync(a->accepted + lo, d.accepted, m, cudaMemcpyDeviceToHost, s));
  }
  for (int i = 0; i < 3; ++i) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->hstreams[i]));
  if (dstats) {
    double h[6];
    const int used = NCH < 3 ? NCH : 3;
    CUDA_TRY(ctx, cudaMemcpy(h, dstats, (size_t)used * 2 * sizeof(double), cudaMemcpyDeviceToHost));
    for (int i = 0; i < used; ++i) {
      a->stats[0] += h[2 * i];
      a->stats[1] += h[2 * i + 1];
    }
  }
  return L2HMC_OK;
}

static int transition_host_once(l2hmc_ctx *ctx, const l2hmc_transition_args *a);

// The host-buffer entry point is synchronous: when its launch raised the fp16 range bit of the context's status word
// the call is repeated at once with the tf32 split (fp32 range).  Device-resident callers (l2hmc_transition) are
// asynchronous: the affected chains of THAT call carry non-finite proposals, which p_accept maps to probability 0
// (utils/dynamics.py:309), i.e. they are rejected and keep x; every later launch of the context sees the sticky bit
// (polled from pinned host memory, no synchronisation) and runs the tf32 split.  l2hmc_status_flags reports it.
extern "C" int l2hmc_transition_host(l2hmc_ctx *ctx, const l2hmc_transition_args *a) {
  const bool was_set = ctx && ctx->status_h && (*(volatile unsigned int *)ctx->status_h & STATUS_F16_RANGE);
  int rc = transition_host_once(ctx, a);
  if (rc != L2HMC_OK || !ctx || ctx->kernel != L2HMC_KERNEL_TC || !ctx->tc_used_f16 || was_set) return rc;
  if (!(*(volatile unsigned int *)ctx->status_h & STATUS_F16_RANGE)) return L2HMC_OK;
  ctx->td.f16 = 0;
  if (a->stats) {  // the repeated call must not count twice: the caller's accumulators are host memory here
    a->stats[0] = ctx->host_stats0[0];
    a->stats[1] = ctx->host_stats0[1];
  }
  return transition_host_once(ctx, a);
}

static int transition_host_once(l2hmc_ctx *ctx, const l2hmc_transition_args *a) {
  int rc = validate_transition(ctx, a, true);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  if (a->n == 0) return L2HMC_OK;
  if (!ctx->hstream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->hstream, cudaStreamNonBlocking));
  cudaStream_t s = ctx->hstream;
  const size_t n = (size_t)a->n, D = (size_t)ctx->sh.D, K = (size_t)a->n_transitions;
  l2hmc_transition_args d = *a;
  d.stream = s;
  {
    // Chains are independent and Philox is keyed by the global chain id, so a large batch can be cut into chunks whose
    // H2D copy, kernel and D2H copy run on separate streams: the copies of one chunk hide under the kernel of another.
    // (Fused kernels only: the layered engine's workspace belongs to the context.  Single transition only: the
    // multi-transition inputs are laid out [K][n].)
    int kernel = 0;
    if ((rc = resolve_kernel(ctx, &kernel))) return rc;
    if (kernel != L2HMC_KERNEL_LAYERED && K == 1 && n >= 32768) return transition_host_chunked(ctx, a);
  }
  if ((rc = ensure(ctx, ctx->hx, n * D))) return rc;
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hx.p, a->x, n * D * sizeof(float), cudaMemcpyHostToDevice, s));
  d.x = ctx->hx.p;
  if (a->v) {
    if ((rc = ensure(ctx, ctx->hv, K * n * D))) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hv.p, a->v, K * n * D * sizeof(float), cudaMemcpyHostToDevice, s));
    d.v = ctx->hv.p;
  }
  if (a->u) {
    if ((rc = ensure(ctx, ctx->hu, K * n))) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hu.p, a->u, K * n * sizeof(float), cudaMemcpyHostToDevice, s));
    d.u = ctx->hu.p;
  }
  if (a->dir) {
    if ((rc = ensure_u8(ctx, ctx->hdir, ctx->hdir_n, K * n))) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hdir, a->dir, K * n, cudaMemcpyHostToDevice, s));
    d.dir = ctx->hdir;
  }
  if (a->aux && ctx->lay.dm.aux > 0) {
    const size_t na = n * (size_t)ctx->lay.dm.aux;
    if ((rc = ensure(ctx, ctx->haux, na))) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->haux.p, a->aux, na * sizeof(float), cudaMemcpyHostToDevice, s));
    d.aux = ctx->haux.p;
  }
  if ((rc = ensure(ctx, ctx->hxo, n * D))) return rc;
  if ((rc = ensure(ctx, ctx->hpx, n))) return rc;
  d.x_out = ctx->hxo.p;
  d.px_out = ctx->hpx.p;
  if (a->v_out) {
    if ((rc = ensure(ctx, ctx->hvo, n * D))) return rc;
    d.v_out = ctx->hvo.p;
  }
  if (a->x_next || a->do_mh) {
    if ((rc = ensure(ctx, ctx->hxn, n * D))) return rc;
    d.x_next = ctx->hxn.p;
  }
  if (a->accepted) {
    if ((rc = ensure_u8(ctx, ctx->hacc, ctx->hacc_n, n))) return rc;
    d.accepted = ctx->hacc;
  }
  if (a->stats) {
    if ((rc = ensure(ctx, ctx->hstats, 12))) return rc;
    d.stats = reinterpret_cast<double *>(ctx->hstats.p);
    ctx->host_stats0[0] = a->stats[0];
    ctx->host_stats0[1] = a->stats[1];
    CUDA_TRY(ctx, cudaMemsetAsync(d.stats, 0, 2 * sizeof(double), s));
  }
  if ((rc = launch_transition(ctx, &d, s))) return rc;
  if (a->x_out) CUDA_TRY(ctx, cudaMemcpyAsync(a->x_out, d.x_out, n * D * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (a->px_out) CUDA_TRY(ctx, cudaMemcpyAsync(a->px_out, d.px_out, n * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (a->v_out) CUDA_TRY(ctx, cudaMemcpyAsync(a->v_out, d.v_out, n * D * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (a->x_next) CUDA_TRY(ctx, cudaMemcpyAsync(a->x_next, d.x_next, n * D * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (a->accepted) CUDA_TRY(ctx, cudaMemcpyAsync(a->accepted, d.accepted, n, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(ctx, cudaStreamSynchronize(s));
  if (a->stats) {
    double h[2];
    CUDA_TRY(ctx, cudaMemcpy(h, d.stats, sizeof(h), cudaMemcpyDeviceToHost));
    a->stats[0] += h[0];
    a->stats[1] += h[1];
  }
  return L2HMC_OK;
}

// ---------------------------------------------------------------------------------------------
// components
// ---------------------------------------------------------------------------------------------
#define GRID(n) (unsigned)(((n) + 127) / 128), 128

// Decoder target: U(x) -> workspace U (and grad U(x) -> ab[:, D:2D]) for caller rows x [n, D] and the bound aux rows.
static int lay_component_energy(l2hmc_ctx *ctx, cudaStream_t s, int64_t n, const float *x, int want_grad, const char *who) {
  if (!ctx->lay.aux || ctx->lay.aux_n != n)
    return fail(ctx, L2HMC_EINVAL, "%s: bind aux rows for these %lld chains first (l2hmc_bind_aux)", who, (long long)n);
  int rc = lay_ensure_ws(ctx, n, s);
  if (rc) return rc;
  ctx->lay.ws_n = n;
  const layered::LayDims &dm = ctx->lay.dm;
  const long long tot = n * dm.D;
  layered::k_lay_copy_rows<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(x, dm.D, ctx->lay.x.p, dm.Dp, dm.D, n);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  return lay_energy_grad(ctx, s, n, ctx->lay.aux, want_grad);
}

extern "C" int l2hmc_energy(l2hmc_ctx *ctx, int64_t n, const float *x, float *out, void *stream) {
  int rc = check_ready(ctx, "l2hmc_energy");
  if (rc) return rc;
  if (n <= 0) return L2HMC_OK;
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  if (ctx->en.kind == L2HMC_ENERGY_DECODER) {
    if ((rc = lay_component_energy(ctx, (cudaStream_t)stream, n, x, 0, "l2hmc_energy"))) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(out, lay_state(ctx, n).U, n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return lay_release(ctx, (cudaStream_t)stream);
  }
  k_energy<<<GRID(n), 0, (cudaStream_t)stream>>>(ctx->en, ctx->sh, n, x, out);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  return L2HMC_OK;
}
extern "C" int l2hmc_grad_energy(l2hmc_ctx *ctx, int64_t n, const float *x, float *out, void *stream) {
  int rc = check_ready(ctx, "l2hmc_grad_energy");
  if (rc) return rc;
  if (n <= 0) return L2HMC_OK;
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  if (ctx->en.kind == L2HMC_ENERGY_DECODER) {
    if ((rc = lay_component_energy(ctx, (cudaStream_t)stream, n, x, 1, "l2hmc_grad_energy"))) return rc;
    const layered::LayDims &dm = ctx->lay.dm;
    const long long tot = n * dm.D;
    layered::k_lay_copy_rows<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ctx->lay.ab.p + dm.D, dm.K1p, out, dm.D, dm.D, n);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return lay_release(ctx, (cudaStream_t)stream);
  }
  k_grad<<<GRID(n), 0, (cudaStream_t)stream>>>(ctx->en, ctx->sh, n, x, out);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  return L2HMC_OK;
}
extern "C" int l2hmc_kinetic(l2hmc_ctx *ctx, int64_t n, const float *v, float *out, void *stream) {
  if (!ctx) return fail(ctx, L2HMC_EINVAL, "l2hmc_kinetic: null context");
  if (n <= 0) return L2HMC_OK;
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  k_kinetic<<<GRID(n), 0, (cudaStream_t)stream>>>(ctx->sh.D, n, v, out);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  return L2HMC_OK;
}
extern "C" int l2hmc_hamiltonian(l2hmc_ctx *ctx, int64_t n, const float *x, const float *v, float *out, void *stream) {
  int rc = check_ready(ctx, "l2hmc_hamiltonian");
  if (rc) return rc;
  if (n <= 0) return L2HMC_OK;
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  if (ctx->en.kind == L2HMC_ENERGY_DECODER) {
    if ((rc = lay_component_energy(ctx, (cudaStream_t)stream, n, x, 0, "l2hmc_hamiltonian"))) return rc;
    layered::k_lay_hamiltonian<<<GRID(n), 0, (cudaStream_t)stream>>>(ctx->sh.D, n, lay_state(ctx, n).U, v, out);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return lay_release(ctx, (cudaStream_t)stream);
  }
  k_hamiltonian<<<GRID(n), 0, (cudaStream_t)stream>>>(ctx->en, ctx->sh, n, x, v, out);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  return L2HMC_OK;
}
extern "C" int l2hmc_p_accept(l2hmc_ctx *ctx, int64_t n, const float *x0, const float *v0, const float *x1,
                              const float *v1, const float *log_jac, float *out, void *stream) {
  int rc = check_ready(ctx, "l2hmc_p_accept");
  if (rc) return rc;
  if (n <= 0) return L2HMC_OK;
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  if (ctx->en.kind == L2HMC_ENERGY_DECODER) {
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = lay_component_energy(ctx, s, n, x0, 0, "l2hmc_p_accept"))) return rc;
    float *U = lay_state(ctx, n).U, *U0 = ctx->lay.vec.p + 6 * n;
    CUDA_TRY(ctx, cudaMemcpyAsync(U0, U, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if ((rc = lay_component_energy(ctx, s, n, x1, 0, "l2hmc_p_accept"))) return rc;
    layered::k_lay_p_accept<<<GRID(n), 0, s>>>(ctx->sh.D, n, U0, U, v0, v1, log_jac, out);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return lay_release(ctx, s);
  }
  k_p_accept<<<GRID(n), 0, (cudaStream_t)stream>>>(ctx->en, ctx->sh, n, x0, v0, x1, v1, log_jac, out);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  return L2HMC_OK;
}
extern "C" int l2hmc_net_apply(l2hmc_ctx *ctx, int net_id, int64_t n, const float *a, const float *b, float step,
                               float *S, float *T, float *Q, void *stream) {
  if (!ctx) return fail(ctx, L2HMC_EINVAL, "l2hmc_net_apply: null context");
  if (net_id != 0 && net_id != 1) return fail(ctx, L2HMC_EINVAL, "l2hmc_net_apply: bad net id");
  if (!ctx->sh.hmc && !ctx->net_set[net_id]) return fail(ctx, L2HMC_EINVAL, "l2hmc_net_apply: net not set");
  if (ctx->sh.H > NET_MAXH) return fail(ctx, L2HMC_EUNSUPPORTED, "l2hmc_net_apply: width > %d", NET_MAXH);
  if (n <= 0) return L2HMC_OK;
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  const float *eaux = nullptr;
  if (!ctx->sh.hmc && ctx->lay.enc.n_layers > 0) {
    if (!ctx->lay.aux || ctx->lay.aux_n != n)
      return fail(ctx, L2HMC_EINVAL, "l2hmc_net_apply: bind aux rows for these chains first (l2hmc_bind_aux)");
    int rc = lay_ensure_ws(ctx, n, (cudaStream_t)stream);
    if (rc) return rc;
    ctx->lay.ws_n = n;
    if ((rc = lay_encode_aux(ctx, (cudaStream_t)stream, n, ctx->lay.aux))) return rc;
    eaux = ctx->lay.eaux.p;
  }
  k_net_apply<<<GRID(n), 0, (cudaStream_t)stream>>>(ctx->net_rawv[net_id], ctx->sh.D, ctx->sh.H, ctx->sh.T, ctx->sh.hmc,
                                                     n, a, b, step, S, T, Q, eaux, ctx->lay.dm.Hp);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  if (eaux) return lay_release(ctx, (cudaStream_t)stream);
  return L2HMC_OK;
}
extern "C" int l2hmc_accept(l2hmc_ctx *ctx, int64_t n, int64_t chain_offset, const float *x, const float *Lx,
                            const float *px, const float *u, uint64_t seed, uint64_t counter, float *out,
                            uint8_t *accepted, void *stream) {
  if (!ctx) return fail(ctx, L2HMC_EINVAL, "l2hmc_accept: null context");
  if (n <= 0) return L2HMC_OK;
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  const long long tot = n * ctx->sh.D;
  k_accept<<<GRID(tot), 0, (cudaStream_t)stream>>>(ctx->sh.D, n, chain_offset, x, Lx, px, u, seed, counter, out, accepted);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  return L2HMC_OK;
}
extern "C" int l2hmc_philox_fill(l2hmc_ctx *ctx, int64_t n, int64_t chain_offset, uint64_t seed, uint64_t counter,
                                 float *v, uint8_t *dir, float *u, void *stream) {
  if (!ctx) return fail(ctx, L2HMC_EINVAL, "l2hmc_philox_fill: null context");
  if (n <= 0) return L2HMC_OK;
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  k_philox_fill<<<GRID(n), 0, (cudaStream_t)stream>>>(ctx->sh.D, n, chain_offset, seed, counter, v, dir, u);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  return L2HMC_OK;
}

// ---------------------------------------------------------------------------------------------
// diagnostics on the device-resident sample trace (utils/func_utils.py:45-54,114-120)
// ---------------------------------------------------------------------------------------------
// part[tau][c] = sum_{t < S - tau} sum_{e in chunk c} X[t, e] * X[t + tau, e], products and sums in fp64 (the
// reference's numpy keeps float32 products and per-step sums; the results agree to fp32 rounding).  Fixed reduction order: deterministic.
__global__ void k_autocov_partial(const float *X, long long S, long long E, long long chunk, int nchunks, double *part) {
  const long long tau = blockIdx.x;
  const int c = blockIdx.y;
  const long long e0 = c * chunk, e1 = (e0 + chunk < E) ? e0 + chunk : E;
  double s = 0.0;
  for (long long t = 0; t + tau < S; ++t) {
    const float *a = X + t * E, *b = X + (t + tau) * E;
    for (long long e = e0 + threadIdx.x; e < e1; e += blockDim.x) s += (double)a[e] * (double)b[e];
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[tau * nchunks + c] = sh[0];
}
// out[tau] = autocovariance(X / scale, tau) = (sum / n) / (S - tau) / scale^2
__global__ void k_autocov_final(const double *part, int nchunks, long long S, long long n, double scale, long long L, double *out) {
  const long long tau = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (tau >= L) return;
  double s = 0.0;
  for (int c = 0; c < nchunks; ++c) s += part[tau * nchunks + c];
  out[tau] = s / (double)n / (double)(S - tau) / (scale * scale);
}

extern "C" int l2hmc_acl_spectrum(l2hmc_ctx *ctx, int64_t n_steps, int64_t n, const float *trace, double scale,
                                  int64_t n_lags, double *out, void *stream) {
  if (!ctx) return fail(ctx, L2HMC_EINVAL, "l2hmc_acl_spectrum: null context");
  if (n_steps < 1 || n < 1 || !trace || !out) return fail(ctx, L2HMC_EINVAL, "l2hmc_acl_spectrum: bad argument");
  if (n_lags < 1 || n_lags > n_steps) return fail(ctx, L2HMC_EINVAL, "l2hmc_acl_spectrum: need 1 <= n_lags <= n_steps");
  if (!(scale > 0.0)) return fail(ctx, L2HMC_EINVAL, "l2hmc_acl_spectrum: scale must be > 0");
  if (n_lags > 2147483647LL) return fail(ctx, L2HMC_EUNSUPPORTED, "l2hmc_acl_spectrum: too many lags");
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  const long long E = n * (long long)ctx->sh.D;
  // enough chunks to fill the GPU when there are few lags, at least 4096 elements per chunk
  long long nch = (148LL * 8 + n_lags - 1) / n_lags;
  const long long max_ch = (E + 4095) / 4096;
  if (nch > max_ch) nch = max_ch;
  if (nch < 1) nch = 1;
  if (nch > 65535) nch = 65535;
  const long long chunk = (E + nch - 1) / nch;
  const size_t need = (size_t)n_lags * (size_t)nch * 2;  // doubles held in a float buffer
  int rc = ensure(ctx, ctx->diag, need);
  if (rc) return rc;
  double *part = reinterpret_cast<double *>(ctx->diag.p);
  cudaStream_t s = (cudaStream_t)stream;
  k_autocov_partial<<<dim3((unsigned)n_lags, (unsigned)nch), 256, 0, s>>>(trace, n_steps, E, chunk, (int)nch, part);
  CUDA_TRY(ctx, cudaGetLastError());
  k_autocov_final<<<(unsigned)((n_lags + 127) / 128), 128, 0, s>>>(part, (int)nch, n_steps, n, scale, n_lags, out);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches += 2;
  return L2HMC_OK;
}

// ---------------------------------------------------------------------------------------------
// introspection
// ---------------------------------------------------------------------------------------------
extern "C" const char *l2hmc_kernel_name(const l2hmc_ctx *ctx) {
  if (!ctx) return "none";
  switch (ctx->kernel) {
    case L2HMC_KERNEL_TILE: return "tile_fma";
    case L2HMC_KERNEL_SMALL: return "small_fma";
    case L2HMC_KERNEL_TC: return ctx->tc_used_f16 ? "tc_3xf16" : "tc_3xtf32";  // operand split of the last launch
    case L2HMC_KERNEL_LAYERED: return ctx->lay.gemm_tc ? ((ctx->lay.gemm_f16 && !(*(volatile unsigned int *)ctx->status_h & STATUS_F16_RANGE)) ? "layered_tc3xf16" : "layered_tc3xtf32") : "layered_fma";
    default: return "none";
  }
}
extern "C" int64_t l2hmc_launch_count(const l2hmc_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int l2hmc_debug_counters(l2hmc_ctx *ctx, int64_t *out, int n) {
  if (!ctx || !out || n < 0 || n > 56) return fail(ctx, L2HMC_EINVAL, "l2hmc_debug_counters: bad argument");
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  CUDA_TRY(ctx, cudaDeviceSynchronize());
  long long h[56];
  CUDA_TRY(ctx, cudaMemcpyFromSymbol(h, tc::g_tc_dbg, sizeof(h)));
  for (int i = 0; i < n; ++i) out[i] = (int64_t)h[i];
  return L2HMC_OK;
}

extern "C" int l2hmc_timing_enable(l2hmc_ctx *ctx, int on) {
  if (!ctx) return fail(ctx, L2HMC_EINVAL, "l2hmc_timing_enable: null context");
  ctx->timing = on != 0;
  ctx->ev_used = 0;
  return L2HMC_OK;
}
extern "C" int l2hmc_timing_read(l2hmc_ctx *ctx, double *avg_ms, int64_t *launches) {
  if (!ctx || !avg_ms || !launches) return fail(ctx, L2HMC_EINVAL, "l2hmc_timing_read: null argument");
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  double tot = 0.0;
  int64_t cnt = 0;
  for (size_t i = 0; i + 1 < ctx->ev_used; i += 2) {
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev[i + 1]));
    float ms = 0.f;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]));
    tot += ms;
    cnt++;
  }
  *avg_ms = cnt ? tot / cnt : 0.0;
  *launches = cnt;
  ctx->ev_used = 0;
  return L2HMC_OK;
}

// ---- training path (train.cuh / train_host.cuh; train_small.cuh: the fused kernel for small nets) ----------------
#include "train_host.cuh"
#include "train_small.cuh"

static NetRaw grads_as_raw(const l2hmc_net_grads &g) {
  NetRaw r;
  r.W1 = g.W1; r.b1 = g.b1; r.W2 = g.W2; r.b2 = g.b2; r.W3 = g.W3; r.b3 = g.b3; r.W4 = g.W4; r.b4 = g.b4;
  r.Ws = g.Ws; r.bs = g.bs; r.Wt = g.Wt; r.bt = g.bt; r.Wq = g.Wq; r.bq = g.bq; r.ls = g.scale_s; r.lq = g.scale_q;
  return r;
}

static int launch_small_train(l2hmc_ctx *ctx, const l2hmc_loss_grad_args *a) {
  small::SmallTrainArgs A;
  A.base.sh = ctx->sh;
  A.base.xnet = ctx->net_rawv[0];
  A.base.vnet = ctx->net_rawv[1];
  A.base.tbx = ctx->net_dev[0].tb;
  A.base.tbv = ctx->net_dev[1].tb;
  A.base.en = ctx->en;
  A.base.mask = ctx->mask.p;
  memset(&A.base.io, 0, sizeof(A.base.io));
  small::TrainIO &t = A.tio;
  t.n = a->n; t.x = a->x; t.v = a->v; t.dir = a->dir; t.scale = a->scale; t.inv_count = a->inv_count; t.loss_kind = a->loss_kind;
  t.loss = a->loss; t.d_eps = a->d_eps; t.gx = grads_as_raw(a->grad_xnet); t.gv = grads_as_raw(a->grad_vnet);
  t.x_out = a->x_out; t.px_out = a->px_out;
  cudaStream_t s = (cudaStream_t)a->stream;
  const unsigned blocks = (unsigned)((a->n + small::NT - 1) / small::NT);
  if (ctx->sh.D <= 2 && ctx->sh.H <= 10)
    small::small_train_kernel<2, 10, 32><<<blocks, small::NT, small::small_train_smem_bytes<2, 10>(ctx->sh.T), s>>>(A);
  else
    small::small_train_kernel<4, 16, 32><<<blocks, small::NT, small::small_train_smem_bytes<4, 16>(ctx->sh.T), s>>>(A);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches += 1;
  return L2HMC_OK;
}

extern "C" int l2hmc_loss_grad(l2hmc_ctx *ctx, const l2hmc_loss_grad_args *a) {
  int rc = check_ready(ctx, "l2hmc_loss_grad");
  if (rc) return rc;
  if (!a) return fail(ctx, L2HMC_EINVAL, "l2hmc_loss_grad: null args");
  if (ctx->sh.hmc) return fail(ctx, L2HMC_EINVAL, "l2hmc_loss_grad: an HMC-mode context has no parameters to train");
  if (!ctx->mask_set) return fail(ctx, L2HMC_EINVAL, "l2hmc_loss_grad: masks not set (l2hmc_set_masks)");
  if (!(ctx->net_set[0] && ctx->net_set[1])) return fail(ctx, L2HMC_EINVAL, "l2hmc_loss_grad: XNet/VNet not set (l2hmc_set_net)");
  const bool covered = ctx->en.kind == L2HMC_ENERGY_GAUSSIAN || ctx->en.kind == L2HMC_ENERGY_GMM ||
                       ctx->en.kind == L2HMC_ENERGY_ROUGHWELL || ctx->en.kind == L2HMC_ENERGY_FUNNEL;
  if (!covered || ctx->lay.enc.n_layers > 0)
    return fail(ctx, L2HMC_EUNSUPPORTED, "l2hmc_loss_grad: covers the closed-form energies without aux (kind %d given)",
                ctx->en.kind);
  if (a->n < 0) return fail(ctx, L2HMC_EINVAL, "l2hmc_loss_grad: n < 0");
  if (a->n == 0) return L2HMC_OK;
  if (!a->x || !a->v || !a->dir || !a->loss || !a->d_eps || !tr_grads_complete(a->grad_xnet) || !tr_grads_complete(a->grad_vnet))
    return fail(ctx, L2HMC_EINVAL, "l2hmc_loss_grad: x, v, dir, loss, d_eps and all 2 x 16 gradient tensors are required");
  if (a->loss_kind < 0 || a->loss_kind > 3) return fail(ctx, L2HMC_EINVAL, "l2hmc_loss_grad: loss_kind %d unknown", a->loss_kind);
  if (!(a->scale > 0.f) || !isfinite(a->scale) || !(a->inv_count > 0.f) || !isfinite(a->inv_count))
    return fail(ctx, L2HMC_EINVAL, "l2hmc_loss_grad: scale and inv_count must be finite and > 0");
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  // small nets (the notebook's: x_dim <= 4, width <= 16) with a separable loss: the whole batch in ONE launch, one chain
  // per thread (train_small.cuh); L2HMC_TRAIN_FUSED=0 keeps the launch sequence (the two are compared in the tests)
  {
    const char *fe = getenv("L2HMC_TRAIN_FUSED");
    const bool fused_ok = !(fe && fe[0] == '0') && ctx->sh.D <= 4 && ctx->sh.H <= 16 && ctx->sh.T <= 32 && a->loss_kind <= 1;
    if (fused_ok) return launch_small_train(ctx, a);
  }
  return tr_loss_grad(ctx, a);
}
