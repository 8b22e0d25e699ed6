// One-chain-per-thread transition kernel for the small nets of the reference's notebook
// (SCGExperiment.ipynb: x_dim 2, width 10) and anything with x_dim <= 4, width <= 16.
//
// With D = 2 and H = 10 a net call is 200 MACs: there is nothing to tile.  Each thread keeps x, v, grad U and
// both hidden layers of its chain in registers for the whole transition; the (zero-padded) weights of both
// nets sit in shared memory and every lane reads the same word (broadcast), so a warp needs no shuffle, no
// barrier and no divergence except the per-chain direction predicate, which is folded arithmetically.
// Math follows utils/dynamics.py:115-201,246-309 and utils/sampler.py:28-55 exactly like kernel_tile.cuh.
#pragma once
#include "common.cuh"

namespace l2hmc {
namespace small {

constexpr int NT = 128;

// shared-memory image of one net, all extents padded to the template sizes (DM, HM)
template <int DM, int HM>
struct NetS {
  float W1[DM][HM], W2[DM][HM], W4[HM][HM], b4[HM];
  float Ws[HM][DM], Wt[HM][DM], Wq[HM][DM], bs[DM], bt[DM], bq[DM], es[DM], eq[DM];
};

struct SmallArgs {
  Shape sh;
  NetRaw xnet, vnet;   // reference-layout device copies
  const float *tbx, *tbv;  // [T][LDE] folded time/bias tables (b1 + b2 + tau(t) W3 + b3), NetDev::tb
  EnergyDev en;
  const float *mask;   // [T][DP]
  TransitionIO io;
};

template <int DM, int HM>
__device__ __forceinline__ void load_net(NetS<DM, HM> &s, const NetRaw &w, int D, int H) {
  for (int i = threadIdx.x; i < DM * HM; i += NT) {
    const int d = i / HM, j = i - d * HM;
    const bool ok = d < D && j < H;
    s.W1[d][j] = ok ? w.W1[d * H + j] : 0.f;
    s.W2[d][j] = ok ? w.W2[d * H + j] : 0.f;
    const int jj = i / DM, dd = i - jj * DM;  // [HM][DM] view of the same index range
    const bool ok2 = jj < H && dd < D;
    s.Ws[jj][dd] = ok2 ? w.Ws[jj * D + dd] : 0.f;
    s.Wt[jj][dd] = ok2 ? w.Wt[jj * D + dd] : 0.f;
    s.Wq[jj][dd] = ok2 ? w.Wq[jj * D + dd] : 0.f;
  }
  for (int i = threadIdx.x; i < HM * HM; i += NT) {
    const int a = i / HM, b = i - a * HM;
    s.W4[a][b] = (a < H && b < H) ? w.W4[a * H + b] : 0.f;
  }
  for (int i = threadIdx.x; i < HM; i += NT) s.b4[i] = i < H ? w.b4[i] : 0.f;
  for (int i = threadIdx.x; i < DM; i += NT) {
    const bool ok = i < D;
    s.bs[i] = ok ? w.bs[i] : 0.f;
    s.bt[i] = ok ? w.bt[i] : 0.f;
    s.bq[i] = ok ? w.bq[i] : 0.f;
    s.es[i] = ok ? expf(w.ls[i]) : 1.f;
    s.eq[i] = ok ? expf(w.lq[i]) : 1.f;
  }
}

// exp / tanh of the updates.  FAST: ex2.approx + rcp.approx as in the tensor-core kernels' epilogues (abs error ~1e-7 for the O(1)
// arguments here; tanhf / expf cost ~25 / ~8 instructions with branches, 32 calls per leapfrog step of a 2-d chain)
__device__ __forceinline__ float sm_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool FAST>
__device__ __forceinline__ float sm_exp(float x) {
  return FAST ? sm_ex2(x * 1.4426950408889634f) : expf(x);
}
template <bool FAST>
__device__ __forceinline__ float sm_tanh(float x) {
  if (!FAST) return tanhf(x);
  const float t = sm_ex2(x * 2.8853900817779268f);  // e^{2x}; inf gives 1 - 0, 0 gives 1 - 2
  return 1.f - __fdividef(2.f, t + 1.f);
}

#ifndef L2HMC_SMALL_F32X2
#define L2HMC_SMALL_F32X2 1  // packed fp32 FMAs (FFMA2, sm_100): two of the independent accumulators of a layer per instruction
#endif
// (h0, h1) += a * (w0, w1): the same two IEEE FMAs as fmaf, one issue slot (the kernel is bound by issue slots, not by the FMA pipe)
__device__ __forceinline__ void fma_pair(float a, float w0, float w1, float &h0, float &h1) {
#if L2HMC_SMALL_F32X2
  const float2 r = __ffma2_rn(make_float2(a, a), make_float2(w0, w1), make_float2(h0, h1));
  h0 = r.x;
  h1 = r.y;
#else
  h0 = fmaf(a, w0, h0);
  h1 = fmaf(a, w1, h1);
#endif
}

// [S, T, Q] = net([a, b, t]) with the time/bias row tb (already selected for this chain's direction)
template <int DM, int HM, bool FAST>
__device__ __forceinline__ void net_eval(const NetS<DM, HM> &n, const float *tb, const float (&a)[DM], const float (&b)[DM],
                                         float (&S)[DM], float (&T)[DM], float (&Q)[DM]) {
  float h1[HM], h2[HM];
#pragma unroll
  for (int j = 0; j < HM; ++j) h1[j] = tb[j];
  static_assert(HM % 2 == 0 && DM % 2 == 0, "pairs of accumulators");
#pragma unroll
  for (int d = 0; d < DM; ++d)
#pragma unroll
    for (int j = 0; j < HM; j += 2) {  // fmaf(b, W2, fmaf(a, W1, h1)) per element, as before
      fma_pair(a[d], n.W1[d][j], n.W1[d][j + 1], h1[j], h1[j + 1]);
      fma_pair(b[d], n.W2[d][j], n.W2[d][j + 1], h1[j], h1[j + 1]);
    }
#pragma unroll
  for (int j = 0; j < HM; ++j) {
    h1[j] = fmaxf(h1[j], 0.f);
    h2[j] = n.b4[j];
  }
#pragma unroll
  for (int i = 0; i < HM; ++i)
#pragma unroll
    for (int j = 0; j < HM; j += 2) fma_pair(h1[i], n.W4[i][j], n.W4[i][j + 1], h2[j], h2[j + 1]);
#pragma unroll
  for (int d = 0; d < DM; ++d) {
    S[d] = n.bs[d];
    T[d] = n.bt[d];
    Q[d] = n.bq[d];
  }
#pragma unroll
  for (int i = 0; i < HM; ++i) {
    const float h = fmaxf(h2[i], 0.f);
#pragma unroll
    for (int d = 0; d < DM; d += 2) {
      fma_pair(h, n.Ws[i][d], n.Ws[i][d + 1], S[d], S[d + 1]);
      fma_pair(h, n.Wt[i][d], n.Wt[i][d + 1], T[d], T[d + 1]);
      fma_pair(h, n.Wq[i][d], n.Wq[i][d + 1], Q[d], Q[d + 1]);
    }
  }
#pragma unroll
  for (int d = 0; d < DM; ++d) {
    S[d] = n.es[d] * sm_tanh<FAST>(S[d]);
    Q[d] = n.eq[d] * sm_tanh<FAST>(Q[d]);
  }
}

template <int DM>
__device__ __forceinline__ void grad_small(const EnergyDev &en, const Shape &sh, const float (&x)[DM], float (&g)[DM]) {
  // Gaussian and mixture targets inline, unrolled over the (at most DM) dimensions: the same operations in the same order as
  // grad_one (common.cuh), without the call, the strided local arrays and the run-time dimension loops -- grad U is evaluated
  // once per leapfrog step and the generic routine was ~10 % of the instructions of a step
  const int D = sh.D;
  if (en.kind == 0 || en.kind == 1) {
    const float T = en.temperature;
    const int ncomp = en.kind == 0 ? 1 : en.ncomp;
    float V[MAX_COMP];
    float mx = -INFINITY, s = 0.f;
    if (en.kind == 1) {
      for (int c = 0; c < ncomp; ++c) {
        const float *mu = en.mu + c * sh.DP;
        const float *S = en.Ssym + (size_t)c * sh.DP * sh.LDS;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < DM; ++j) {
          float r = 0.f;
#pragma unroll
          for (int i = 0; i < DM; ++i)
            if (i < D && j < D) r = fmaf(x[i] - mu[i], S[i * sh.LDS + j], r);
          if (j < D) q = fmaf(r, x[j] - mu[j], q);
        }
        V[c] = -0.5f * q + en.logc[c];
        mx = fmaxf(mx, V[c]);
      }
      for (int c = 0; c < ncomp; ++c) {
        V[c] = expf(V[c] - mx);
        s += V[c];
      }
    }
#pragma unroll
    for (int j = 0; j < DM; ++j) {
      float acc = 0.f;
      for (int c = 0; c < ncomp; ++c) {
        const float *mu = en.mu + c * sh.DP;
        const float *S = en.Ssym + (size_t)c * sh.DP * sh.LDS;
        float r = 0.f;
#pragma unroll
        for (int i = 0; i < DM; ++i)
          if (i < D && j < D) r = fmaf(x[i] - mu[i], S[i * sh.LDS + j], r);
        acc = en.kind == 0 ? r : fmaf(V[c] / s, r, acc);
      }
      g[j] = j < D ? acc / T : 0.f;
    }
    return;
  }
  float xs[DM], gl[DM];
#pragma unroll
  for (int d = 0; d < DM; ++d) { xs[d] = x[d]; gl[d] = 0.f; }
  grad_chain(en, sh, xs, 1, gl, 1);
#pragma unroll
  for (int d = 0; d < DM; ++d) g[d] = d < sh.D ? gl[d] : 0.f;
}

#ifndef L2HMC_SMALL_MINBLOCKS
#define L2HMC_SMALL_MINBLOCKS 6  // resident CTAs per SM the compiler leaves registers for (measured: 1 -> 0.189 / 0.707 ms, 6 -> 0.171 / 0.565, 8 -> 0.176 / 0.559 on configs 1 / 3)
#endif
template <int DM, int HM, bool FAST>
__global__ void __launch_bounds__(NT, L2HMC_SMALL_MINBLOCKS) small_transition_kernel(const __grid_constant__ SmallArgs A) {
  extern __shared__ __align__(16) float smem_small[];
  const Shape &sh = A.sh;
  const TransitionIO &io = A.io;
  const int D = sh.D, H = sh.H, T = sh.T;
  NetS<DM, HM> &NX = *reinterpret_cast<NetS<DM, HM> *>(smem_small);
  NetS<DM, HM> &NV = *(&NX + 1);
  float *tbx = reinterpret_cast<float *>(&NV + 1);  // [T][HM]
  float *tbv = tbx + T * HM;
  float *msk = tbv + T * HM;                         // [T][DM]
  if (!sh.hmc) {
    load_net(NX, A.xnet, D, H);
    load_net(NV, A.vnet, D, H);
    for (int i = threadIdx.x; i < T * HM; i += NT) {
      const int t = i / HM, j = i - t * HM;
      tbx[i] = j < H ? A.tbx[t * sh.LDE + j] : 0.f;  // rows of the tile-layout table (stride LDE)
      tbv[i] = j < H ? A.tbv[t * sh.LDE + j] : 0.f;
    }
  }
  for (int i = threadIdx.x; i < T * DM; i += NT) {
    const int t = i / DM, d = i - t * DM;
    msk[i] = d < D ? A.mask[t * sh.DP + d] : 0.f;
  }
  __syncthreads();

  const long long g = (long long)blockIdx.x * NT + threadIdx.x;
  if (g >= io.n) return;
  const float eps = sh.eps;
  float x[DM], v[DM], gr[DM];
#pragma unroll
  for (int d = 0; d < DM; ++d) x[d] = d < D ? io.x[g * D + d] : 0.f;

  float x0[DM], h_old = 0.f, lj = 0.f;
  for (int tr = 0; tr < io.n_transitions; ++tr) {
    const unsigned long long ctr = io.counter + (unsigned long long)tr;
    const bool first = !io.chain || tr == 0;              // chain mode: one Hamiltonian / log|J| / start point for all sub-proposals
    const bool closing = !io.chain || tr == io.n_transitions - 1;
    if (first) {
#pragma unroll
      for (int d = 0; d < DM; ++d) x0[d] = x[d];
    }
    if (io.v != nullptr) {
#pragma unroll
      for (int d = 0; d < DM; ++d) v[d] = d < D ? io.v[((long long)tr * io.n + g) * D + d] : 0.f;
    } else {
      float z[4];
      philox_normals4(io.seed, ctr, io.chain_offset + g, 0, z);
#pragma unroll
      for (int d = 0; d < DM; ++d) v[d] = d < D ? z[d] : 0.f;  // DM <= 4: one Philox block
    }
    int pd = 1;
    float pu = 0.f;
    if (io.dir_mode == 3 || (io.do_mh && io.u == nullptr)) philox_dir_u(io.seed, ctr, io.chain_offset + g, pd, pu);
    bool fwd = true;
    if (io.dir_mode == 1) fwd = false;
    else if (io.dir_mode == 2) fwd = io.dir[(long long)tr * io.n + g] != 0;
    else if (io.dir_mode == 3) fwd = pd != 0;
    if (io.do_mh && io.u != nullptr) pu = io.u[(io.chain ? 0ll : (long long)tr * io.n) + g];  // chain mode: one set of uniforms

    grad_small<DM>(A.en, sh, x, gr);
    float kin = 0.f;
    if (first) {
      if (io.chain) {  // H(x_in, init_v): the sub-proposals draw their own momenta (utils/sampler.py:35-36, 79)
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        if (io.v0 == nullptr) philox_normals4(io.seed, chain_v0_counter(io), io.chain_offset + g, 0, z);
#pragma unroll
        for (int d = 0; d < DM; ++d) {
          const float v0d = d < D ? (io.v0 ? io.v0[g * D + d] : z[d]) : 0.f;
          kin = fmaf(v0d, v0d, kin);
        }
      } else {
#pragma unroll
        for (int d = 0; d < DM; ++d) kin = fmaf(v[d], v[d], kin);
      }
      float xs0[DM];
#pragma unroll
      for (int d = 0; d < DM; ++d) xs0[d] = x[d];
      h_old = energy_chain(A.en, sh, xs0, 1) + 0.5f * kin;
      lj = 0.f;
    }

    for (int it = 0; it < T; ++it) {
      const int t = fwd ? it : T - 1 - it;
      // four sub-updates per leapfrog step: v half step, two masked x updates, v half step at the new x
      // (utils/dynamics.py:115-157 / :159-201).  One rolled loop: a single copy of the net evaluation in the instruction
      // stream (four inlined copies made a ~100 KB loop body and instruction fetch was 23 % of the stall samples).
#pragma unroll 1
      for (int call = 0; call < 4; ++call) {
        const bool isv = call == 0 || call == 3;
        if (call == 3) grad_small<DM>(A.en, sh, x, gr);
        float S[DM], Tt[DM], Q[DM], k[DM], a[DM], b[DM];
#pragma unroll
        for (int d = 0; d < DM; ++d) {
          const float m = msk[t * DM + d];
          k[d] = (fwd == (call == 1)) ? m : 1.f - m;  // used by the x updates only
          a[d] = isv ? x[d] : v[d];
          b[d] = isv ? gr[d] : k[d] * x[d];
        }
        if (sh.hmc) {
#pragma unroll
          for (int d = 0; d < DM; ++d) S[d] = Tt[d] = Q[d] = 0.f;
        } else {
          // ONE base address for the net of this sub-update, opaque to the compiler: with `isv ? NV : NX` it selected between
          // the two nets at every weight load (two UIADD3 per LDS: 19 % of the executed instructions of an issue-bound kernel)
          uint32_t noff = isv ? (uint32_t)sizeof(NetS<DM, HM>) : 0u;
          asm volatile("" : "+r"(noff));
          const NetS<DM, HM> &N = *reinterpret_cast<const NetS<DM, HM> *>(reinterpret_cast<const char *>(&NX) + noff);
          net_eval<DM, HM, FAST>(N, (isv ? tbv : tbx) + t * HM, a, b, S, Tt, Q);
        }
        if (isv) {
#pragma unroll
          for (int d = 0; d < DM; ++d) {
            const float sv = fwd ? (0.5f * eps) * S[d] : (-0.5f * eps) * S[d];
            const float cterm = (0.5f * eps) * (-(sm_exp<FAST>(eps * Q[d]) * gr[d]) + Tt[d]);
            const float e = sm_exp<FAST>(sv);
            v[d] = fwd ? (v[d] * e + cterm) : ((v[d] - cterm) * e);
            lj += sv;
          }
        } else {
#pragma unroll
          for (int d = 0; d < DM; ++d) {
            const float uu = 1.f - k[d];
            const float sx = fwd ? eps * S[d] : -eps * S[d];
            const float inner = eps * (sm_exp<FAST>(eps * Q[d]) * v[d] + Tt[d]);
            const float e = sm_exp<FAST>(sx);
            const float nx = fwd ? (x[d] * e + inner) : (e * (x[d] - inner));
            x[d] = k[d] * x[d] + uu * nx;
            lj += uu * sx;
          }
        }
      }
    }
    if (sh.hmc) lj = 0.f;
    if (!closing) continue;  // chain mode: the next sub-proposal starts from this proposal, no Metropolis step in between
    kin = 0.f;
#pragma unroll
    for (int d = 0; d < DM; ++d) kin = fmaf(v[d], v[d], kin);
    float xs1[DM];
#pragma unroll
    for (int d = 0; d < DM; ++d) xs1[d] = x[d];
    const float h_new = energy_chain(A.en, sh, xs1, 1) + 0.5f * kin;
    const float p = accept_prob(h_old, h_new, lj);
    const float px = io.log_jac ? lj : p;
    const bool acc = io.do_mh && (px - pu >= 0.f);
    const bool last = tr == io.n_transitions - 1;
    stats_add(io.stats, px, acc ? 1 : 0);
    if (io.trace) {
#pragma unroll
      for (int d = 0; d < DM; ++d)
        if (d < D) io.trace[((long long)tr * io.n + g) * D + d] = acc ? x[d] : x0[d];
    }
    if (last) {
      io.px_out[g] = px;
      if (io.accepted) io.accepted[g] = acc ? 1 : 0;
#pragma unroll
      for (int d = 0; d < DM; ++d)
        if (d < D) {
          io.x_out[g * D + d] = x[d];
          if (io.v_out) io.v_out[g * D + d] = v[d];
          if (io.do_mh) io.x_next[g * D + d] = acc ? x[d] : x0[d];
        }
    } else {
#pragma unroll
      for (int d = 0; d < DM; ++d) x[d] = acc ? x[d] : x0[d];
    }
  }
}

template <int DM, int HM>
inline size_t small_smem_bytes(int T) {
  return 2 * sizeof(NetS<DM, HM>) + sizeof(float) * ((size_t)2 * T * HM + (size_t)T * DM);
}

}  // namespace small
}  // namespace l2hmc
