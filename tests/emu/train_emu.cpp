// TEST INFRASTRUCTURE (CPU suite only; never part of libl2hmc.so).
// Runs the SOURCE of the training-path kernels (l2hmc_b200/csrc/train.cuh) and of their host driver (train_host.cuh) on
// the host, so that `pytest -m "not gpu"` can check index arithmetic, strides, the GEMM tiling, the sweep order and
// the vector-Jacobian products against oracle/l2hmc_reverse.py without a GPU.  One fiber per CUDA thread of a block;
// blocks run one after another; __syncthreads / warp shuffles are scheduler barriers over the block / the warp.  The component kernels the driver borrows from l2hmc_api.cu (k_grad, k_hamiltonian: already checked on
// the GPU) are restated here for the two energies the training path covers.
//
// Build (tests/test_train_emu.py does this):  g++ -std=c++17 -O1 -shared -fPIC -DL2HMC_TRAIN_EMU ...
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <ucontext.h>

#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../../include/l2hmc.h"

// ---- CUDA vocabulary --------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static
#define __restrict__

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_ {
  unsigned x, y, z;
};
static uint3_ threadIdx, blockIdx;
static dim3 blockDim, gridDim;

typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0, cudaMemcpyDeviceToDevice = 3 };
// fresh device memory is filled with NaNs here, so a kernel that reads scratch it has not written shows up in the results
static inline cudaError_t cudaMalloc(void **p, size_t bytes) {
  *p = malloc(bytes);
  if (!*p) return 2;
  memset(*p, 0xFF, bytes);
  return 0;
}
static inline cudaError_t cudaFree(void *p) { free(p); return 0; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }

// Every CUDA thread of a block is a fiber (ucontext) on the calling OS thread; a barrier is a yield to a scheduler that
// releases a warp (shuffles) or the block (__syncthreads) once all of its live threads wait there.  Blocks run one after
// another.  Deterministic, and no OS threads are involved.
namespace emu {
constexpr int MAX_THREADS = 256;
constexpr size_t STACK_BYTES = 128 * 1024;
enum { RUN = 0, WAIT_WARP = 1, WAIT_BLOCK = 2 };
struct Fiber {
  ucontext_t ctx;
  bool done = true;
  int wait = RUN;
};
static Fiber fibers[MAX_THREADS];
static char *stacks = nullptr;
static ucontext_t sched_ctx;
static int cur = -1;                        // fiber that is running, -1: the scheduler / a plain call
static std::function<void()> job;
static float warp_buf[MAX_THREADS];
static dim3 cur_block, cur_grid;
static uint3_ cur_bidx;

static void set_ids(int t) {
  threadIdx = {(unsigned)t, 0, 0};
  blockIdx = cur_bidx;
  blockDim = cur_block;
  gridDim = cur_grid;
}
static void fiber_main() {
  job();
  fibers[cur].done = true;
  swapcontext(&fibers[cur].ctx, &sched_ctx);
}
static void yield(int kind) {
  if (cur < 0) abort();  // a kernel that synchronises was launched as one that does not
  fibers[cur].wait = kind;
  swapcontext(&fibers[cur].ctx, &sched_ctx);
}

static void run_block(int nt) {
  if (!stacks) stacks = static_cast<char *>(malloc(STACK_BYTES * MAX_THREADS));
  for (int t = 0; t < nt; ++t) {
    Fiber &f = fibers[t];
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = stacks + STACK_BYTES * t;
    f.ctx.uc_stack.ss_size = STACK_BYTES;
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, fiber_main, 0);
    f.done = false;
    f.wait = RUN;
  }
  for (;;) {
    bool progressed = false, alive = false;
    for (int t = 0; t < nt; ++t) {
      Fiber &f = fibers[t];
      if (f.done || f.wait != RUN) continue;
      cur = t;
      set_ids(t);
      swapcontext(&sched_ctx, &f.ctx);
      cur = -1;
      progressed = true;
    }
    for (int w = 0; w < nt / 32; ++w) {  // a warp whose live threads all wait on a shuffle moves on
      int live = 0, waiting = 0;
      for (int t = 32 * w; t < 32 * w + 32; ++t) {
        live += !fibers[t].done;
        waiting += !fibers[t].done && fibers[t].wait == WAIT_WARP;
      }
      if (live && waiting == live) {
        for (int t = 32 * w; t < 32 * w + 32; ++t) fibers[t].wait = RUN;
        progressed = true;
      }
    }
    int live = 0, waiting = 0;
    for (int t = 0; t < nt; ++t) {
      live += !fibers[t].done;
      waiting += !fibers[t].done && fibers[t].wait == WAIT_BLOCK;
    }
    alive = live > 0;
    if (live && waiting == live) {
      for (int t = 0; t < nt; ++t) fibers[t].wait = RUN;
      progressed = true;
    }
    if (!alive) return;
    if (!progressed) abort();  // deadlock: threads of one warp / block wait at different barriers
  }
}

// kernels that synchronise run as fibers; the others run their threads one after another as plain calls
static bool cooperative(const char *name) {
  for (const char *k : {"k_gemm", "k_update", "k_update_vjp", "k_loss", "k_loss_v", "k_loss_stats"}) {
    const char *p = strstr(name, k);
    if (p && strlen(p) == strlen(k) && (p == name || p[-1] == ':')) return true;
  }
  return false;
}

template <class F>
void launch(const char *name, dim3 grid, dim3 block, F &&body) {
  const int nt = (int)block.x;
  if (nt > MAX_THREADS || block.y != 1 || block.z != 1 || nt % 32 != 0) abort();
  const bool coop = cooperative(name);
  cur_block = block;
  cur_grid = grid;
  if (coop) job = body;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        cur_bidx = {bx, by, bz};
        if (coop) {
          run_block(nt);
        } else {
          for (int t = 0; t < nt; ++t) {
            set_ids(t);
            body();
          }
        }
      }
  job = nullptr;
}
}  // namespace emu

static inline void __syncthreads() { emu::yield(emu::WAIT_BLOCK); }
static inline float __shfl_xor_sync(unsigned, float v, int o) {
  const int t = emu::cur;
  if (t < 0) abort();
  emu::warp_buf[t] = v;
  emu::yield(emu::WAIT_WARP);
  const float r = emu::warp_buf[t ^ o];
  emu::yield(emu::WAIT_WARP);
  return r;
}
static inline float atomicAdd(float *p, float v) {
  const float old = *p;
  *p = old + v;
  return old;
}

// ---- the few types of common.cuh / l2hmc_api.cu the training path touches -------------------------------------------
namespace l2hmc {
constexpr int MAX_COMP = 8;
struct NetRaw {
  const float *W1, *b1, *W2, *b2, *W3, *b3, *W4, *b4, *Ws, *bs, *Wt, *bt, *Wq, *bq, *ls, *lq;
};
struct EnergyDev {
  int kind, ncomp;
  const float *mu, *Ssym, *logc;
  float s0, s1, temperature;
};
struct Shape {
  int D, DP, H, HP, T, LDE, LDH, LDS, hmc;
  float eps;
};
}  // namespace l2hmc
using l2hmc::EnergyDev;
using l2hmc::NetRaw;
using l2hmc::Shape;

struct DevBufEmu {
  float *p = nullptr;
  size_t n = 0;
};
struct l2hmc_ctx {
  Shape sh;
  EnergyDev en;
  DevBufEmu mask;
  NetRaw net_rawv[2];
  DevBufEmu train_ws;
  float *train_part = nullptr;
  long long launches = 0;
  std::string err;
  ~l2hmc_ctx() { free(train_ws.p); }
};

static int fail(l2hmc_ctx *ctx, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return code;
}
#define CUDA_TRY(ctx, expr)                                        \
  do {                                                             \
    if ((expr) != cudaSuccess) return fail(ctx, L2HMC_ECUDA, #expr); \
  } while (0)
#define GRID(n) (unsigned)(((n) + 127) / 128), 128

// grad U / T_emp and U / T_emp for the energies the training path covers (restated; the real kernels live in l2hmc_api.cu)
static float emu_quad(const float *mu, const float *S, const Shape &sh, const float *x, float *g) {
  float q = 0.f;
  for (int j = 0; j < sh.D; ++j) {
    float r = 0.f;
    for (int i = 0; i < sh.D; ++i) r = fmaf(x[i] - mu[i], S[i * sh.LDS + j], r);
    if (g) g[j] = r;
    q = fmaf(r, x[j] - mu[j], q);
  }
  return q;
}
static float emu_energy(const EnergyDev &en, const Shape &sh, const float *x) {
  float U = 0.f;
  if (en.kind == 0) {
    U = 0.5f * emu_quad(en.mu, en.Ssym, sh, x, nullptr);
  } else if (en.kind == 1) {
    float V[l2hmc::MAX_COMP], mx = -INFINITY, s = 0.f;
    for (int c = 0; c < en.ncomp; ++c) {
      V[c] = -0.5f * emu_quad(en.mu + c * sh.DP, en.Ssym + (size_t)c * sh.DP * sh.LDS, sh, x, nullptr) + en.logc[c];
      mx = fmaxf(mx, V[c]);
    }
    for (int c = 0; c < en.ncomp; ++c) s += expf(V[c] - mx);
    U = -(logf(s) + mx);
  } else if (en.kind == 2) {
    float n = 0.f, cs = 0.f;
    for (int i = 0; i < sh.D; ++i) {
      n = fmaf(x[i], x[i], n);
      cs += cosf(x[i] / en.s1);
    }
    U = 0.5f * n + en.s0 * cs;
  } else {
    const float sigma = en.s0, clip = en.s1, v = x[0];
    float ss = 0.f;
    for (int i = 1; i < sh.D; ++i) ss = fmaf(x[i], x[i], ss);
    float s = expf(v);
    if (v > clip) s = expf(clip);
    if (-clip > v) s = expf(-clip);
    U = 0.5f * ((v / sigma) * (v / sigma) + ss / s + (float)(sh.D - 1) * logf(6.2831855f * s));
  }
  return U / en.temperature;
}
static void k_grad(EnergyDev en, Shape sh, long long n, const float *x, float *out) {
  const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g >= n) return;
  const float *xr = x + g * sh.D;
  float *o = out + g * sh.D;
  if (en.kind == 0) {
    emu_quad(en.mu, en.Ssym, sh, xr, o);
  } else if (en.kind == 1) {
    float V[l2hmc::MAX_COMP], gc[l2hmc::MAX_COMP][64], mx = -INFINITY, s = 0.f;
    for (int c = 0; c < en.ncomp; ++c) {
      V[c] = -0.5f * emu_quad(en.mu + c * sh.DP, en.Ssym + (size_t)c * sh.DP * sh.LDS, sh, xr, gc[c]) + en.logc[c];
      mx = fmaxf(mx, V[c]);
    }
    for (int c = 0; c < en.ncomp; ++c) s += (V[c] = expf(V[c] - mx));
    for (int j = 0; j < sh.D; ++j) {
      float a = 0.f;
      for (int c = 0; c < en.ncomp; ++c) a = fmaf(V[c] / s, gc[c][j], a);
      o[j] = a;
    }
  } else if (en.kind == 2) {
    for (int j = 0; j < sh.D; ++j) o[j] = xr[j] - en.s0 * sinf(xr[j] / en.s1) / en.s1;
  } else {
    const float sigma = en.s0, clip = en.s1, v = xr[0];
    float ss = 0.f;
    for (int i = 1; i < sh.D; ++i) ss = fmaf(xr[i], xr[i], ss);
    const bool out_of_clip = v > clip || -clip > v;
    const float s = out_of_clip ? expf(v > clip ? clip : -clip) : expf(v);
    o[0] = v / (sigma * sigma) + (out_of_clip ? 0.f : 0.5f * (-ss / s + (float)(sh.D - 1)));
    for (int j = 1; j < sh.D; ++j) o[j] = xr[j] / s;
  }
  for (int j = 0; j < sh.D; ++j) o[j] /= en.temperature;
}
static void k_hamiltonian(EnergyDev en, Shape sh, long long n, const float *x, const float *v, float *out) {
  const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g >= n) return;
  float s = 0.f;
  for (int d = 0; d < sh.D; ++d) s = fmaf(v[g * sh.D + d], v[g * sh.D + d], s);
  out[g] = emu_energy(en, sh, x + g * sh.D) + 0.5f * s;
}

#include "../../l2hmc_b200/csrc/train_host.cuh"

struct EmuEnergy {  // padded copies in the layout of the context (mu [ncomp][DP], Ssym [ncomp][DP][LDS])
  std::vector<float> mu_p, S_p;
};
static void emu_setup(l2hmc_ctx &ctx, EmuEnergy &E, int D, int H, int T, float eps, float temperature, int kind, int ncomp,
                      const float *mu, const float *S, const float *logc, float s0, float s1) {
  Shape &sh = ctx.sh;
  sh.D = D; sh.DP = (D + 3) / 4 * 4; sh.H = H; sh.HP = (H + 3) / 4 * 4; sh.T = T; sh.LDE = 128; sh.LDH = 192;
  sh.LDS = (sh.DP + 127) / 128 * 128; sh.hmc = 0; sh.eps = eps;
  E.mu_p.assign((size_t)ncomp * sh.DP, 0.f);
  E.S_p.assign((size_t)ncomp * sh.DP * sh.LDS, 0.f);
  if (kind == 0 || kind == 1)
    for (int c = 0; c < ncomp; ++c)
      for (int i = 0; i < D; ++i) {
        E.mu_p[(size_t)c * sh.DP + i] = mu[c * D + i];
        for (int j = 0; j < D; ++j)
          E.S_p[((size_t)c * sh.DP + i) * sh.LDS + j] = 0.5f * (S[(c * D + i) * D + j] + S[(c * D + j) * D + i]);
      }
  ctx.en = EnergyDev{kind, ncomp, E.mu_p.data(), E.S_p.data(), logc, s0, s1, temperature};
}

// ---- entry points for tests/test_train_emu.py -----------------------------------------------------------------------
// Parameters in the reference layout (l2hmc_net_params order, host pointers); energy: kind 0 / 1 (mu [ncomp, D],
// S [ncomp, D, D], logc [ncomp]), 2 (scalars eps, denominator) or 3 (scalars sigma, clip).  Gradients are written to
// gx / gv (16 tensors each, caller-zeroed).
extern "C" int emu_loss_grad(int D, int H, int T, float eps, float temperature, int kind, int ncomp, const float *mu,
                             const float *S, const float *logc, float s0, float s1, const float *mask,
                             const l2hmc_net_params *xnet, const l2hmc_net_params *vnet, const l2hmc_loss_grad_args *a,
                             char *err, int err_len) {
  if (D > 64 || ncomp > l2hmc::MAX_COMP) return L2HMC_EUNSUPPORTED;
  l2hmc_ctx ctx;
  EmuEnergy E;
  emu_setup(ctx, E, D, H, T, eps, temperature, kind, ncomp, mu, S, logc, s0, s1);
  const Shape &sh = ctx.sh;
  std::vector<float> mask_p((size_t)T * sh.DP, 0.f);
  for (int t = 0; t < T; ++t)
    for (int d = 0; d < D; ++d) mask_p[(size_t)t * sh.DP + d] = mask[t * D + d];
  ctx.mask.p = mask_p.data();
  const l2hmc_net_params *ps[2] = {xnet, vnet};
  for (int i = 0; i < 2; ++i) {
    const l2hmc_net_params *p = ps[i];
    ctx.net_rawv[i] = NetRaw{p->W1, p->b1, p->W2, p->b2, p->W3, p->b3, p->W4, p->b4, p->Ws, p->bs,
                             p->Wt, p->bt, p->Wq, p->bq, p->scale_s, p->scale_q};
  }
  if (!tr_grads_complete(a->grad_xnet) || !tr_grads_complete(a->grad_vnet)) return L2HMC_EINVAL;
  const int rc = tr_loss_grad(&ctx, a);
  if (err && err_len > 0) snprintf(err, err_len, "%s", ctx.err.c_str());
  return rc;
}

// k_hvp alone: out [n, D] += w . d(grad U / T_emp)/dx at x
extern "C" int emu_hvp(int D, float temperature, int kind, int ncomp, const float *mu, const float *S, const float *logc,
                       float s0, float s1, long long n, const float *x, const float *w, float *out) {
  if (D > 64 || ncomp > l2hmc::MAX_COMP) return L2HMC_EUNSUPPORTED;
  l2hmc_ctx ctx;
  EmuEnergy E;
  emu_setup(ctx, E, D, 1, 1, 0.1f, temperature, kind, ncomp, mu, S, logc, s0, s1);
  const Shape &sh = ctx.sh;
  const EnergyDev &en = ctx.en;
  emu::launch("k_hvp", dim3((unsigned)((n + 127) / 128)), dim3(128), [&] { tr::k_hvp(en, sh, n, x, w, out); });
  return 0;
}
