// Shared device-side definitions for libl2hmc.so (sm_100a).
// Reference citations are into /root/reference (brain-research/l2hmc).
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>
#include <stdint.h>

namespace l2hmc {

constexpr int MAX_COMP = 8;  // GMM components the kernels hold in registers

// ---------------------------------------------------------------------------------------------
// Device-side parameter views (all pointers are device memory owned by the context)
// ---------------------------------------------------------------------------------------------

// Packed S/T/Q net for the tile kernel (SCGExperiment.ipynb:51-77 net, folded):
//   Wemb [2*DP][LDE]  rows 0..DP-1 = embed_1/W (input a), rows DP..2DP-1 = embed_2/W (input b)
//   tb   [T][LDE]     tb[t] = b1 + b2 + (tau(t) embed_3/W + b3), tau from utils/dynamics.py:99-105
//   W4   [HP][LDE], b4 [LDE]
//   Wh   [HP][LDH]    column 3*d + {0,1,2} = linear_{s,t,f}/W[:, d];  bh [LDH] likewise
//   es, eq [DP]       exp(scale_s), exp(scale_f)  (utils/layers.py:83-84)
struct NetDev {
  const float *Wemb, *tb, *W4, *b4, *Wh, *bh, *es, *eq;
};

// Raw (unpadded, reference-layout) copy used by the component kernels.
struct NetRaw {
  const float *W1, *b1, *W2, *b2, *W3, *b3, *W4, *b4, *Ws, *bs, *Wt, *bt, *Wq, *bq, *ls, *lq;
};

struct EnergyDev {
  int kind;           // l2hmc_energy_kind
  int ncomp;          // 1 for Gaussian
  const float *mu;    // [ncomp][DP]  zero padded
  const float *Ssym;  // [ncomp][DP][LDS]  0.5*(S + S^T), zero padded
  const float *logc;  // [ncomp]
  float s0, s1;       // ROUGHWELL: eps, denominator (eps or eps*eps) ; FUNNEL: sigma, clip
  float temperature;  // Dynamics.energy divides by it (utils/dynamics.py:203-212)
  // kind 5 (MIXED): U = (1 - beta) U_a + beta U_b, the annealed energy of utils/ais.py:44-45.  U_a: kind_a with the fields
  // above; U_b: kind_b with the fields below.
  int kind_a, kind_b, ncomp_b;
  const float *mu_b, *Ssym_b, *logc_b;
  float s0_b, s1_b, beta;
};
constexpr int ENERGY_KIND_MIXED = 5;
constexpr int MIX_MAXD = 64;  // the fused kernels that evaluate energies per chain cover x_dim <= 64

struct Shape {
  int D, DP;    // x_dim and x_dim rounded up to a multiple of 4
  int H, HP;    // width and width rounded up to a multiple of 4
  int T;
  int LDE;      // leading dim of Wemb/tb/W4 (multiple of 128)
  int LDH;      // leading dim of Wh (multiple of 192)
  int LDS;      // leading dim of Ssym (multiple of 128)
  int hmc;
  float eps;
};

struct TransitionIO {
  long long n;
  long long chain_offset;
  const float *x, *v, *u;
  const uint8_t *dir;
  int dir_mode, log_jac, do_mh, n_transitions;
  unsigned long long seed, counter;
  float *x_out, *v_out, *px_out, *x_next;
  uint8_t *accepted;
  double *stats;          // [2] or null: += sum of px, += number accepted, over every chain and transition of the launch
  float *trace;           // [n_transitions][n][D] or null: the Metropolis output x_next after every fused transition
  unsigned int *status;   // the context's status word (pinned, host-mapped): bit 0 = fp16 operand range exceeded
  // chain_operator (utils/sampler.py:57-85) in one launch: the n_transitions loop composes n_transitions SUB-PROPOSALS
  // (fresh direction and momentum each, log|J| accumulated, no Metropolis step in between) and closes with ONE accept
  // probability p_accept(x_in, v0, x_K, v_K, sum log|J|) and one Metropolis step against x_in.
  int chain;
  const float *v0;        // chain mode: init_v [n][D] paired with x_in in the first Hamiltonian (null: drawn, Philox)
};

// chain mode: the momentum of the first Hamiltonian when the caller gave none (tf.random_normal, utils/sampler.py:58-59):
// the normals stream at the call counter one past the last sub-proposal's
__device__ __forceinline__ unsigned long long chain_v0_counter(const TransitionIO &io) {
  return io.counter + (unsigned long long)io.n_transitions;
}

// status bits (l2hmc_status_flags)
constexpr unsigned int STATUS_F16_RANGE = 1u;

struct KernelArgs {
  Shape sh;
  NetDev xnet, vnet;
  EnergyDev en;
  const float *mask;  // [T][DP]
  TransitionIO io;
};

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  Host twin: l2hmc_b200/philox.py -- keep in lock step.
//   key     = (seed_lo, seed_hi)
//   counter = (chain_lo, chain_hi, block, (call_counter << 2) | stream)   [call_counter < 2^30]
//   stream 0: momentum normals, block b covers dims 4b..4b+3 (Box-Muller on word pairs)
//   stream 1: word 0 bit 0 = direction bit (1 = forward); word 1 >> 8 = accept uniform * 2^24
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifdef __CUDA_ARCH__
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
#else
  uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
  uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
  uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

__device__ __forceinline__ void philox_words(unsigned long long seed, unsigned long long counter,
                                             long long chain, uint32_t block, uint32_t stream,
                                             uint32_t (&out)[4]) {
  out[0] = (uint32_t)chain;
  out[1] = (uint32_t)((unsigned long long)chain >> 32);
  out[2] = block;
  out[3] = ((uint32_t)counter << 2) | stream;
  philox4x32_10(out, (uint32_t)seed, (uint32_t)(seed >> 32));
}

// Box-Muller on two 32-bit words -> two N(0,1) (same mapping in philox.py).
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float &n0, float &n1) {
  const float u1 = ((float)(a >> 8) + 0.5f) * 5.9604644775390625e-08f;  // (0,1), 2^-24 grid
  const float u2 = ((float)(b >> 8) + 0.5f) * 5.9604644775390625e-08f;
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincosf(6.2831855f * u2, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

__device__ __forceinline__ void philox_normals4(unsigned long long seed, unsigned long long counter,
                                                long long chain, int block, float (&z)[4]) {
  uint32_t w[4];
  philox_words(seed, counter, chain, (uint32_t)block, 0u, w);
  box_muller(w[0], w[1], z[0], z[1]);
  box_muller(w[2], w[3], z[2], z[3]);
}

__device__ __forceinline__ void philox_dir_u(unsigned long long seed, unsigned long long counter,
                                             long long chain, int &dir, float &u) {
  uint32_t w[4];
  philox_words(seed, counter, chain, 0u, 1u, w);
  dir = (int)(w[0] & 1u);
  u = (float)(w[1] >> 8) * 5.9604644775390625e-08f;  // [0,1)
}

// ---------------------------------------------------------------------------------------------
// Small helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// Accept statistics (utils/sampler.py:53-55 aggregated over the launch): called by the threads that hold one valid chain's
// (px, accepted) each; the calling lanes of a warp are reduced with shuffles and ONE lane issues the two atomics.
__device__ __forceinline__ void stats_add(double *stats, float px, int acc) {
  if (stats == nullptr) return;
  namespace cg = cooperative_groups;
  const cg::coalesced_group g = cg::coalesced_threads();
  const float s = cg::reduce(g, px, cg::plus<float>());
  const int a = cg::reduce(g, acc, cg::plus<int>());
  if (g.thread_rank() == 0) {
    atomicAdd(stats, (double)s);
    atomicAdd(stats + 1, (double)a);
  }
}

// p_accept tail (utils/dynamics.py:306-309): exp(min(v, 0)), non-finite -> 0
__device__ __forceinline__ float accept_prob(float e_old, float e_new, float log_jac) {
  float v = e_old - e_new + log_jac;
  float p = expf(fminf(v, 0.0f));
  // fminf drops a NaN operand; the reference's tf.minimum propagates it and then maps it to 0
  if (!(v == v)) p = 0.0f;
  return isfinite(p) ? p : 0.0f;
}

// ---------------------------------------------------------------------------------------------
// Per-chain (scalar) energies and gradients, x given as a strided column: x[d*stride].
// Used by the component kernels and by the tile kernel for the non-GEMM energy kinds.
// ---------------------------------------------------------------------------------------------
// One closed-form kind; `part` selects which parameter set of the descriptor it reads (0: the primary fields, 1: the _b fields
// of a mixed energy).  No division by the temperature here.
__device__ __noinline__ float energy_one(const EnergyDev &en, int part, const Shape &sh, const float *x, int stride) {
  const int D = sh.D;
  float U = 0.f;
  const int kind = part ? en.kind_b : (en.kind == ENERGY_KIND_MIXED ? en.kind_a : en.kind);
  const int ncomp = part ? en.ncomp_b : en.ncomp;
  const float *en_mu = part ? en.mu_b : en.mu, *en_S = part ? en.Ssym_b : en.Ssym, *en_logc = part ? en.logc_b : en.logc;
  const float s0 = part ? en.s0_b : en.s0, s1 = part ? en.s1_b : en.s1;
  switch (kind) {
    case 0: {  // Gaussian: 0.5 * d S d^T (utils/distributions.py:31-32)
      for (int j = 0; j < D; ++j) {
        float r = 0.f;
        for (int i = 0; i < D; ++i) r = fmaf(x[i * stride] - en_mu[i], en_S[i * sh.LDS + j], r);
        U = fmaf(r, x[j * stride] - en_mu[j], U);
      }
      U *= 0.5f;
    } break;
    case 1: {  // GMM: -logsumexp_i(-q_i + log c_i) (utils/distributions.py:125-134)
      float V[MAX_COMP];
      float mx = -INFINITY;
      for (int c = 0; c < ncomp; ++c) {
        const float *mu = en_mu + c * sh.DP;
        const float *S = en_S + (size_t)c * sh.DP * sh.LDS;
        float q = 0.f;
        for (int j = 0; j < D; ++j) {
          float r = 0.f;
          for (int i = 0; i < D; ++i) r = fmaf(x[i * stride] - mu[i], S[i * sh.LDS + j], r);
          q = fmaf(r, x[j * stride] - mu[j], q);
        }
        V[c] = -0.5f * q + en_logc[c];
        mx = fmaxf(mx, V[c]);
      }
      float s = 0.f;
      for (int c = 0; c < ncomp; ++c) s += expf(V[c] - mx);
      U = -(logf(s) + mx);
    } break;
    case 2: {  // RoughWell (utils/distributions.py:90-97)
      const float e = s0, den = s1;
      float n = 0.f, cs = 0.f;
      for (int i = 0; i < D; ++i) {
        float xi = x[i * stride];
        n = fmaf(xi, xi, n);
        cs += cosf(xi / den);
      }
      U = 0.5f * n + e * cs;
    } break;
    case 3: {  // GaussianFunnel (utils/distributions.py:161-180)
      const float sigma = s0, clip = s1;
      const float v = x[0];
      const float vs = v / sigma;
      const float lpv = vs * vs;
      float ss = 0.f;
      for (int i = 1; i < D; ++i) ss = fmaf(x[i * stride], x[i * stride], ss);
      const float n = (float)(D - 1);
      const float two_pi = 6.2831855f;
      float s = expf(v);
      if (v > clip) s = expf(clip);
      if (-clip > v) s = expf(-clip);
      U = 0.5f * (lpv + ss / s + n * logf(two_pi * s));
    } break;
    default: break;
  }
  return U;
}

// U(x) / temperature for the configured energy; kind 5 mixes two closed-form kinds (utils/ais.py:44-45), no recursion (the
// per-thread stack stays statically sized).
__device__ __forceinline__ float energy_chain(const EnergyDev &en, const Shape &sh, const float *x, int stride) {
  if (en.kind == ENERGY_KIND_MIXED)
    return ((1.f - en.beta) * energy_one(en, 0, sh, x, stride) + en.beta * energy_one(en, 1, sh, x, stride)) / en.temperature;
  return energy_one(en, 0, sh, x, stride) / en.temperature;
}

// g[d*gstride] = dU/dx_d / temperature
// mode 0: g = dU/dx / T (one closed-form kind).  Mixed energy: mode 1 writes g = w * dU_a/dx, mode 2 finishes
// g = (g + w * dU_b/dx) / T, so no per-thread scratch array is needed.
__device__ __forceinline__ void grad_emit(float *g, float val, float w, float T, int mode) {
  *g = mode == 0 ? val / T : (mode == 1 ? w * val : (*g + w * val) / T);
}

__device__ __noinline__ void grad_one(const EnergyDev &en, int part, const Shape &sh, const float *x, int stride,
                                      float *g, int gstride, float w, int mode) {
  const int D = sh.D;
  const float T = en.temperature;
  const int kind = part ? en.kind_b : (en.kind == ENERGY_KIND_MIXED ? en.kind_a : en.kind);
  const int ncomp = part ? en.ncomp_b : en.ncomp;
  const float *en_mu = part ? en.mu_b : en.mu, *en_S = part ? en.Ssym_b : en.Ssym, *en_logc = part ? en.logc_b : en.logc;
  const float s0 = part ? en.s0_b : en.s0, s1 = part ? en.s1_b : en.s1;
  switch (kind) {
    case 0: {
      for (int j = 0; j < D; ++j) {
        float r = 0.f;
        for (int i = 0; i < D; ++i) r = fmaf(x[i * stride] - en_mu[i], en_S[i * sh.LDS + j], r);
        grad_emit(g + j * gstride, r, w, T, mode);
      }
    } break;
    case 1: {
      float V[MAX_COMP];
      float mx = -INFINITY;
      for (int c = 0; c < ncomp; ++c) {
        const float *mu = en_mu + c * sh.DP;
        const float *S = en_S + (size_t)c * sh.DP * sh.LDS;
        float q = 0.f;
        for (int j = 0; j < D; ++j) {
          float r = 0.f;
          for (int i = 0; i < D; ++i) r = fmaf(x[i * stride] - mu[i], S[i * sh.LDS + j], r);
          q = fmaf(r, x[j * stride] - mu[j], q);
        }
        V[c] = -0.5f * q + en_logc[c];
        mx = fmaxf(mx, V[c]);
      }
      float s = 0.f;
      for (int c = 0; c < ncomp; ++c) {
        V[c] = expf(V[c] - mx);
        s += V[c];
      }
      for (int j = 0; j < D; ++j) {
        float acc = 0.f;
        for (int c = 0; c < ncomp; ++c) {
          const float *mu = en_mu + c * sh.DP;
          const float *S = en_S + (size_t)c * sh.DP * sh.LDS;
          float r = 0.f;
          for (int i = 0; i < D; ++i) r = fmaf(x[i * stride] - mu[i], S[i * sh.LDS + j], r);
          acc = fmaf(V[c] / s, r, acc);
        }
        grad_emit(g + j * gstride, acc, w, T, mode);
      }
    } break;
    case 2: {
      const float e = s0, den = s1;
      for (int i = 0; i < D; ++i) {
        float xi = x[i * stride];
        grad_emit(g + i * gstride, xi - e * sinf(xi / den) / den, w, T, mode);
      }
    } break;
    case 3: {
      const float sigma = s0, clip = s1;
      const float v = x[0];
      float ss = 0.f;
      for (int i = 1; i < D; ++i) ss = fmaf(x[i * stride], x[i * stride], ss);
      const float n = (float)(D - 1);
      const bool hi = v > clip, lo = -clip > v;
      float s = expf(v);
      float gv = v / (sigma * sigma) + 0.5f * (-ss / s + n);
      if (hi) { s = expf(clip); gv = v / (sigma * sigma); }
      if (lo) { s = expf(-clip); gv = v / (sigma * sigma); }
      grad_emit(g, gv, w, T, mode);
      for (int i = 1; i < D; ++i) grad_emit(g + i * gstride, x[i * stride] / s, w, T, mode);
    } break;
    default: break;
  }
}

// g[d*gstride] = dU/dx_d / temperature for the configured energy (kind 5: (1 - beta) dU_a + beta dU_b, utils/ais.py:44-45)
__device__ __forceinline__ void grad_chain(const EnergyDev &en, const Shape &sh, const float *x, int stride, float *g, int gstride) {
  if (en.kind == ENERGY_KIND_MIXED) {
    grad_one(en, 0, sh, x, stride, g, gstride, 1.f - en.beta, 1);
    grad_one(en, 1, sh, x, stride, g, gstride, en.beta, 2);
  } else {
    grad_one(en, 0, sh, x, stride, g, gstride, 1.f, 0);
  }
}

}  // namespace l2hmc
