// Fused training kernel for the notebook's small nets (x_dim <= 4, width <= 16: SCGExperiment.ipynb trains x_dim 2 / width 10 on
// 200 chains): value and gradient of one `propose` batch of the objective in ONE launch, one chain per thread.
//
// Same algorithm as train.cuh / oracle/l2hmc_reverse.py (forward sweep recording the state in front of each of the 4 T
// sub-updates, reverse sweep recomputing one net forward per sub-update and applying its vector-Jacobian product, Hessian-
// vector products of the energy, the p_accept clamp, the direction select) -- but where train.cuh runs ~200 launches per
// leapfrog step over [n, .] arrays in HBM (27 ms per optimiser step at 200 chains: launch-bound), here a thread keeps its
// chain's state, tape (local memory, 8 T x_dim floats) and cotangents to itself, both nets sit in shared memory (broadcast
// reads, as in kernel_small.cuh), and the parameter gradients are accumulated per thread and reduced once at the end
// (warp shuffles -> shared memory -> one atomicAdd per parameter and block).
// Reference: SCGExperiment.ipynb:159-188 (loss + optimizer.minimize), utils/losses.py:36-59, utils/sampler.py:34-44,
// utils/dynamics.py:115-201, 217-218, 246-309.  Covers the separable losses 'mixed' and 'standard'; 'inverse' / 'logsumexp'
// weigh chains by a batch statistic and stay on the launch-sequence path.
#pragma once
#include "kernel_small.cuh"

namespace l2hmc {
namespace small {

struct TrainIO {
  long long n;
  const float *x, *v;     // [n, D] start points, the momentum of each chain's direction
  const uint8_t *dir;     // [n] 1 = forward
  float scale, inv_count;
  int loss_kind;          // 0 mixed, 1 standard
  float *loss, *d_eps;    // [1] +=
  NetRaw gx, gv;          // gradient accumulators shaped like the nets (const-cast: written with atomicAdd), +=
  float *x_out, *px_out;  // Lx [n, D], px [n] or null
};

struct SmallTrainArgs {
  SmallArgs base;         // shape, nets, time/bias tables, masks, energy (io unused)
  TrainIO tio;
};

// per-net gradient accumulator of one thread, padded to the template sizes: ONE flat array with named offsets (the block
// reduction walks it by index; a struct of arrays walked through a float* alias is undefined behaviour the optimiser exploits)
template <int DM, int HM>
struct NetG {
  static constexpr int oW1 = 0, oW2 = oW1 + DM * HM, oW3 = oW2 + DM * HM, ob123 = oW3 + 2 * HM, oW4 = ob123 + HM,
                       ob4 = oW4 + HM * HM, oWs = ob4 + HM, oWt = oWs + HM * DM, oWq = oWt + HM * DM, obs = oWq + HM * DM,
                       obt = obs + DM, obq = obt + DM, ols = obq + DM, olq = ols + DM, COUNT = olq + DM;
  float g[COUNT];
  __device__ __forceinline__ float &W1(int d, int j) { return g[oW1 + d * HM + j]; }
  __device__ __forceinline__ float &W2(int d, int j) { return g[oW2 + d * HM + j]; }
  __device__ __forceinline__ float &W3(int r, int j) { return g[oW3 + r * HM + j]; }
  __device__ __forceinline__ float &b123(int j) { return g[ob123 + j]; }
  __device__ __forceinline__ float &W4(int i, int j) { return g[oW4 + i * HM + j]; }
  __device__ __forceinline__ float &b4(int j) { return g[ob4 + j]; }
  __device__ __forceinline__ float &Ws(int i, int d) { return g[oWs + i * DM + d]; }
  __device__ __forceinline__ float &Wt(int i, int d) { return g[oWt + i * DM + d]; }
  __device__ __forceinline__ float &Wq(int i, int d) { return g[oWq + i * DM + d]; }
  __device__ __forceinline__ float &bs(int d) { return g[obs + d]; }
  __device__ __forceinline__ float &bt(int d) { return g[obt + d]; }
  __device__ __forceinline__ float &bq(int d) { return g[obq + d]; }
  __device__ __forceinline__ float &ls(int d) { return g[ols + d]; }
  __device__ __forceinline__ float &lq(int d) { return g[olq + d]; }
};

// activations a reverse step needs: post-relu hidden layers (z > 0 <=> h > 0) and the two tanh
template <int DM, int HM>
struct Saved {
  float a[DM], b[DM], h1[HM], h2[HM], ts[DM], tq[DM];
};

template <int DM, int HM>
__device__ __forceinline__ void net_eval_saved(const NetS<DM, HM> &n, const float *tb, const float (&a)[DM], const float (&b)[DM],
                                               float (&S)[DM], float (&T)[DM], float (&Q)[DM], Saved<DM, HM> &sv) {
#pragma unroll
  for (int d = 0; d < DM; ++d) {
    sv.a[d] = a[d];
    sv.b[d] = b[d];
  }
#pragma unroll
  for (int j = 0; j < HM; ++j) sv.h1[j] = tb[j];
#pragma unroll
  for (int d = 0; d < DM; ++d)
#pragma unroll
    for (int j = 0; j < HM; ++j) sv.h1[j] = fmaf(b[d], n.W2[d][j], fmaf(a[d], n.W1[d][j], sv.h1[j]));
#pragma unroll
  for (int j = 0; j < HM; ++j) {
    sv.h1[j] = fmaxf(sv.h1[j], 0.f);
    sv.h2[j] = n.b4[j];
  }
#pragma unroll
  for (int i = 0; i < HM; ++i)
#pragma unroll
    for (int j = 0; j < HM; ++j) sv.h2[j] = fmaf(sv.h1[i], n.W4[i][j], sv.h2[j]);
#pragma unroll
  for (int d = 0; d < DM; ++d) {
    S[d] = n.bs[d];
    T[d] = n.bt[d];
    Q[d] = n.bq[d];
  }
#pragma unroll
  for (int i = 0; i < HM; ++i) {
    sv.h2[i] = fmaxf(sv.h2[i], 0.f);
#pragma unroll
    for (int d = 0; d < DM; ++d) {
      S[d] = fmaf(sv.h2[i], n.Ws[i][d], S[d]);
      T[d] = fmaf(sv.h2[i], n.Wt[i][d], T[d]);
      Q[d] = fmaf(sv.h2[i], n.Wq[i][d], Q[d]);
    }
  }
#pragma unroll
  for (int d = 0; d < DM; ++d) {
    sv.ts[d] = tanhf(S[d]);
    sv.tq[d] = tanhf(Q[d]);
    S[d] = n.es[d] * sv.ts[d];
    Q[d] = n.eq[d] * sv.tq[d];
  }
}

// cotangents (gS, gT, gQ) of the net outputs -> (ga, gb) of its two inputs; parameter gradients added into G
// (oracle/l2hmc_reverse.py net_vjp).  ct, st: the time input of this sub-update (utils/dynamics.py:99-105).
template <int DM, int HM>
__device__ __forceinline__ void net_vjp(const NetS<DM, HM> &n, const Saved<DM, HM> &sv, const float (&gS)[DM], const float (&gT)[DM],
                                        const float (&gQ)[DM], float ct, float st, NetG<DM, HM> &G, float (&ga)[DM], float (&gb)[DM]) {
  float g_us[DM], g_uq[DM];
#pragma unroll
  for (int d = 0; d < DM; ++d) {
    g_us[d] = gS[d] * n.es[d] * (1.f - sv.ts[d] * sv.ts[d]);
    g_uq[d] = gQ[d] * n.eq[d] * (1.f - sv.tq[d] * sv.tq[d]);
    G.ls(d) += gS[d] * n.es[d] * sv.ts[d];  // d(e^l tanh u)/dl = e^l tanh u
    G.lq(d) += gQ[d] * n.eq[d] * sv.tq[d];
    G.bs(d) += g_us[d];
    G.bt(d) += gT[d];
    G.bq(d) += g_uq[d];
  }
  float g_z2[HM];
#pragma unroll
  for (int i = 0; i < HM; ++i) {
    float g = 0.f;
#pragma unroll
    for (int d = 0; d < DM; ++d) {
      G.Ws(i, d) = fmaf(sv.h2[i], g_us[d], G.Ws(i, d));
      G.Wt(i, d) = fmaf(sv.h2[i], gT[d], G.Wt(i, d));
      G.Wq(i, d) = fmaf(sv.h2[i], g_uq[d], G.Wq(i, d));
      g = fmaf(g_us[d], n.Ws[i][d], fmaf(gT[d], n.Wt[i][d], fmaf(g_uq[d], n.Wq[i][d], g)));
    }
    g_z2[i] = sv.h2[i] > 0.f ? g : 0.f;
    G.b4(i) += g_z2[i];
  }
  float g_z1[HM];
#pragma unroll
  for (int i = 0; i < HM; ++i) {
    float g = 0.f;
#pragma unroll
    for (int j = 0; j < HM; ++j) {
      G.W4(i, j) = fmaf(sv.h1[i], g_z2[j], G.W4(i, j));
      g = fmaf(g_z2[j], n.W4[i][j], g);
    }
    g_z1[i] = sv.h1[i] > 0.f ? g : 0.f;
  }
#pragma unroll
  for (int d = 0; d < DM; ++d) ga[d] = gb[d] = 0.f;
#pragma unroll
  for (int j = 0; j < HM; ++j) {
    G.b123(j) += g_z1[j];
    G.W3(0, j) = fmaf(ct, g_z1[j], G.W3(0, j));
    G.W3(1, j) = fmaf(st, g_z1[j], G.W3(1, j));
#pragma unroll
    for (int d = 0; d < DM; ++d) {
      G.W1(d, j) = fmaf(sv.a[d], g_z1[j], G.W1(d, j));
      G.W2(d, j) = fmaf(sv.b[d], g_z1[j], G.W2(d, j));
      ga[d] = fmaf(g_z1[j], n.W1[d][j], ga[d]);
      gb[d] = fmaf(g_z1[j], n.W2[d][j], gb[d]);
    }
  }
}

// (w . Hessian of U at x) / temperature, closed form per energy kind (train::k_hvp for one chain), returned BY VALUE: the
// function is not inlined, and an output through a reference to the caller's register array came back without the
// caller's own earlier writes to that array (measured: gx lost its incoming value) -- values in, values out.
template <int DM>
struct VecD {
  float a[DM];
};
template <int DM>
__device__ __noinline__ VecD<DM> hvp_small(const EnergyDev &en, const Shape &sh, VecD<DM> xin, VecD<DM> win) {
  const int D = sh.D;
  float xr[DM], wr[DM], acc[DM];
#pragma unroll
  for (int d = 0; d < DM; ++d) {
    xr[d] = xin.a[d];
    wr[d] = win.a[d];
    acc[d] = 0.f;
  }
  if (en.kind == 0) {
    for (int j = 0; j < D; ++j) {
      float r = 0.f;
      for (int i = 0; i < D; ++i) r = fmaf(wr[i], en.Ssym[i * sh.LDS + j], r);
      acc[j] = r;
    }
  } else if (en.kind == 1) {
    float r[MAX_COMP], sc[MAX_COMP];
    float mx = -INFINITY;
    for (int c = 0; c < en.ncomp; ++c) {
      const float *mu = en.mu + c * sh.DP;
      const float *S = en.Ssym + (size_t)c * sh.DP * sh.LDS;
      float q = 0.f, s = 0.f;
      for (int j = 0; j < D; ++j) {
        float g = 0.f;
        for (int i = 0; i < D; ++i) g = fmaf(xr[i] - mu[i], S[i * sh.LDS + j], g);
        q = fmaf(g, xr[j] - mu[j], q);
        s = fmaf(g, wr[j], s);
      }
      r[c] = -0.5f * q + en.logc[c];
      sc[c] = s;
      mx = fmaxf(mx, r[c]);
    }
    float z = 0.f, gbw = 0.f;
    for (int c = 0; c < en.ncomp; ++c) {
      r[c] = expf(r[c] - mx);
      z += r[c];
    }
    for (int c = 0; c < en.ncomp; ++c) {
      r[c] /= z;
      gbw = fmaf(r[c], sc[c], gbw);
    }
    for (int j = 0; j < D; ++j) {
      float a = 0.f, gb = 0.f;
      for (int c = 0; c < en.ncomp; ++c) {
        const float *mu = en.mu + c * sh.DP;
        const float *S = en.Ssym + (size_t)c * sh.DP * sh.LDS;
        float g = 0.f, wa = 0.f;
        for (int i = 0; i < D; ++i) {
          g = fmaf(xr[i] - mu[i], S[i * sh.LDS + j], g);
          wa = fmaf(wr[i], S[i * sh.LDS + j], wa);
        }
        a = fmaf(r[c], wa - g * sc[c], a);
        gb = fmaf(r[c], g, gb);
      }
      acc[j] = a + gb * gbw;
    }
  } else if (en.kind == 2) {
    const float e = en.s0, den = en.s1;
    for (int j = 0; j < D; ++j) acc[j] = wr[j] * (1.f - e * cosf(xr[j] / den) / (den * den));
  } else {
    const float sigma = en.s0, clip = en.s1;
    const float v = xr[0];
    const bool out = (v > clip) || (-clip > v);
    const float s = out ? expf(v > clip ? clip : -clip) : expf(v);
    float ss = 0.f, wx = 0.f;
    for (int i = 1; i < D; ++i) {
      ss = fmaf(xr[i], xr[i], ss);
      wx = fmaf(wr[i], xr[i], wx);
    }
    const float h00 = 1.f / (sigma * sigma) + (out ? 0.f : 0.5f * ss / s);
    acc[0] = wr[0] * h00 - (out ? 0.f : wx / s);
    for (int j = 1; j < D; ++j) acc[j] = (wr[j] - (out ? 0.f : wr[0] * xr[j])) / s;
  }
  VecD<DM> out;
#pragma unroll
  for (int d = 0; d < DM; ++d) out.a[d] = d < D ? acc[d] / en.temperature : 0.f;
  return out;
}

template <int DM, int HM>
inline size_t small_train_smem_bytes(int T) {
  return small_smem_bytes<DM, HM>(T) + 2 * sizeof(float) * NetG<DM, HM>::COUNT + 2 * sizeof(float);
}

template <int DM, int HM, int TMAX>
__global__ void __launch_bounds__(NT) small_train_kernel(const __grid_constant__ SmallTrainArgs A) {
  extern __shared__ __align__(16) float smem_small[];
  const Shape &sh = A.base.sh;
  const TrainIO &io = A.tio;
  const EnergyDev &en = A.base.en;
  const int D = sh.D, H = sh.H, T = sh.T;
  NetS<DM, HM> &NX = *reinterpret_cast<NetS<DM, HM> *>(smem_small);
  NetS<DM, HM> &NV = *(&NX + 1);
  float *tbx = reinterpret_cast<float *>(&NV + 1);  // [T][HM]
  float *tbv = tbx + T * HM;
  float *msk = tbv + T * HM;                         // [T][DM]
  float *red = msk + T * DM;                         // [2][COUNT] block sums of the two nets' gradients, then loss, d_eps
  constexpr int NP = NetG<DM, HM>::COUNT;
  load_net(NX, A.base.xnet, D, H);
  load_net(NV, A.base.vnet, D, H);
  for (int i = threadIdx.x; i < T * HM; i += NT) {
    const int t = i / HM, j = i - t * HM;
    tbx[i] = j < H ? A.base.tbx[t * sh.LDE + j] : 0.f;
    tbv[i] = j < H ? A.base.tbv[t * sh.LDE + j] : 0.f;
  }
  for (int i = threadIdx.x; i < T * DM; i += NT) {
    const int t = i / DM, d = i - t * DM;
    msk[i] = d < D ? A.base.mask[t * sh.DP + d] : 0.f;
  }
  for (int i = threadIdx.x; i < 2 * NP + 2; i += NT) red[i] = 0.f;
  __syncthreads();

  const long long g = (long long)blockIdx.x * NT + threadIdx.x;
  const bool valid = g < io.n;
  const float eps = sh.eps;
  NetG<DM, HM> GX, GV;  // this thread's parameter-gradient accumulators (local memory)
  for (int i = 0; i < NP; ++i) GX.g[i] = GV.g[i] = 0.f;
  float loss_c = 0.f, deps_c = 0.f;

  if (valid) {
    float x[DM], v[DM], x0[DM], v0[DM];
#pragma unroll
    for (int d = 0; d < DM; ++d) {
      x[d] = x0[d] = d < D ? io.x[g * D + d] : 0.f;
      v[d] = v0[d] = d < D ? io.v[g * D + d] : 0.f;
    }
    const bool fwd = io.dir[g] != 0;
    const float sign = fwd ? 1.f : -1.f;
    float tape[4 * TMAX][2][DM];  // the state in front of each sub-update
    float lj = 0.f;

    // keep-mask of sub-update j (1, 2) of the step at time row t: forward m then 1 - m, backward 1 - m then m
    auto keep_of = [&](int t, int j, int d) -> float {
      const float m = msk[t * DM + d];
      const bool use_m = fwd == (j == 1);
      return use_m ? m : 1.f - m;
    };
    // ---- forward sweep (utils/dynamics.py:115-157 / :159-201) ------------------------------------------------------
    for (int it = 0; it < T; ++it) {
      const int t = fwd ? it : T - 1 - it;
      for (int j = 0; j < 4; ++j) {
        const int rec = it * 4 + j;
#pragma unroll
        for (int d = 0; d < DM; ++d) {
          tape[rec][0][d] = x[d];
          tape[rec][1][d] = v[d];
        }
        float S[DM], Tt[DM], Q[DM];
        Saved<DM, HM> sv;
        if (j == 0 || j == 3) {
          float gr[DM];
          grad_small<DM>(en, sh, x, gr);
          net_eval_saved(NV, tbv + t * HM, x, gr, S, Tt, Q, sv);
#pragma unroll
          for (int d = 0; d < DM; ++d) {
            const float s = 0.5f * sign * eps * S[d];
            const float shift = 0.5f * eps * (-(expf(eps * Q[d]) * gr[d]) + Tt[d]);
            v[d] = fwd ? v[d] * expf(s) + shift : (v[d] - shift) * expf(s);
            lj += s;
          }
        } else {
          float kx[DM];
#pragma unroll
          for (int d = 0; d < DM; ++d) kx[d] = keep_of(t, j, d) * x[d];
          net_eval_saved(NX, tbx + t * HM, v, kx, S, Tt, Q, sv);
#pragma unroll
          for (int d = 0; d < DM; ++d) {
            const float k = keep_of(t, j, d), upd = 1.f - k;
            const float s = sign * eps * S[d];
            const float shift = eps * (expf(eps * Q[d]) * v[d] + Tt[d]);
            const float nx = fwd ? x[d] * expf(s) + shift : expf(s) * (x[d] - shift);
            x[d] = k * x[d] + upd * nx;
            lj += upd * s;
          }
        }
      }
    }
    // ---- objective: p_accept (utils/dynamics.py:302-309), loss_vec and the loss (utils/losses.py:36-59) --------------
    float gX[DM], gV[DM];
    float glj;
    {
      float kin0 = 0.f, kin1 = 0.f, sq = 0.f;
#pragma unroll
      for (int d = 0; d < DM; ++d) {
        kin0 = fmaf(v0[d], v0[d], kin0);
        kin1 = fmaf(v[d], v[d], kin1);
        sq = fmaf(x0[d] - x[d], x0[d] - x[d], sq);
      }
      float xs[DM];
#pragma unroll
      for (int d = 0; d < DM; ++d) xs[d] = x0[d];
      const float H0 = energy_chain(en, sh, xs, 1) + 0.5f * kin0;
#pragma unroll
      for (int d = 0; d < DM; ++d) xs[d] = x[d];
      const float H1 = energy_chain(en, sh, xs, 1) + 0.5f * kin1;
      const float arg = H0 - H1 + lj;
      float p = expf(fminf(arg, 0.f));
      const bool ok = (arg == arg) && isfinite(p);
      if (!ok) p = 0.f;
      const float vv = fmaf(sq, p, 1e-4f);
      float g_v;
      if (io.loss_kind == 0) {  // scale mean(1 / v) - mean(v) / scale
        loss_c = (io.scale / vv - vv / io.scale) * io.inv_count;
        g_v = (-io.scale / (vv * vv) - 1.f / io.scale) * io.inv_count;
      } else {                  // -mean(v)
        loss_c = -vv * io.inv_count;
        g_v = -io.inv_count;
      }
      const float g_arg = (ok && arg < 0.f) ? g_v * sq * p : 0.f;
      float gU1[DM];
      grad_small<DM>(en, sh, x, gU1);
#pragma unroll
      for (int d = 0; d < DM; ++d) {
        gX[d] = g_v * p * 2.f * (x[d] - x0[d]) - g_arg * gU1[d];
        gV[d] = -g_arg * v[d];
      }
      glj = g_arg;
      if (io.x_out) {
#pragma unroll
        for (int d = 0; d < DM; ++d)
          if (d < D) io.x_out[g * D + d] = x[d];
      }
      if (io.px_out) io.px_out[g] = p;
    }
#if defined(L2HMC_DBG_TRAIN) && L2HMC_DBG_TRAIN == 1
#pragma unroll
    for (int d = 0; d < DM; ++d)
      if (d < D) io.x_out[g * D + d] = gX[d];
    io.px_out[g] = glj;
#endif
    // ---- reverse sweep (oracle/l2hmc_reverse.py transition_vjp) --------------------------------------------------------
    for (int it = T - 1; it >= 0; --it) {
      const int t = fwd ? it : T - 1 - it;
      const float targ = 6.2831855f * (float)t / (float)T;  // fp32(2 pi) * t / T (utils/dynamics.py:99-105)
      const float ct = cosf(targ), st = sinf(targ);
      for (int j = 3; j >= 0; --j) {
        const int rec = it * 4 + j;
#if defined(L2HMC_DBG_TRAIN) && L2HMC_DBG_TRAIN >= 2
        if (it == T - 1 && j == 4 - (L2HMC_DBG_TRAIN - 1)) {  // DBG 2: after sub-update j=3 ; 3: after j=2 ...
#pragma unroll
          for (int d = 0; d < DM; ++d)
            if (d < D) io.x_out[g * D + d] = gX[d];
          io.px_out[g] = gV[0];
        }
#endif
        float xs[DM], vs[DM];
#pragma unroll
        for (int d = 0; d < DM; ++d) {
          xs[d] = tape[rec][0][d];
          vs[d] = tape[rec][1][d];
        }
        float S[DM], Tt[DM], Q[DM], gS[DM], gTt[DM], gQ[DM], ga[DM], gb[DM];
        Saved<DM, HM> sv;
        if (j == 0 || j == 3) {
          float gr[DM], g_g[DM];
          grad_small<DM>(en, sh, xs, gr);
          net_eval_saved(NV, tbv + t * HM, xs, gr, S, Tt, Q, sv);
#pragma unroll
          for (int d = 0; d < DM; ++d) {
            const float s = 0.5f * sign * eps * S[d], f = eps * Q[d];
            const float es = expf(s), ef = expf(f);
            const float inner = -(ef * gr[d]) + Tt[d];
            const float shift = 0.5f * eps * inner;
            const float v_o = fwd ? vs[d] * es + shift : (vs[d] - shift) * es;
            const float gvo = gV[d];
            const float g_s = (fwd ? gvo * vs[d] * es : gvo * v_o) + glj;
            const float g_shift = fwd ? gvo : -gvo * es;
            gV[d] = gvo * es;
            const float g_inner = g_shift * (0.5f * eps);
            deps_c += g_shift * 0.5f * inner;
            const float g_f = g_inner * (-(ef * gr[d]));
            g_g[d] = g_inner * (-ef);
            gTt[d] = g_inner;
            gS[d] = g_s * (0.5f * sign * eps);
            deps_c += g_s * (0.5f * sign) * S[d];
            gQ[d] = g_f * eps;
            deps_c += g_f * Q[d];
          }
          net_vjp(NV, sv, gS, gTt, gQ, ct, st, GV, ga, gb);
#pragma unroll
          for (int d = 0; d < DM; ++d) {
            g_g[d] += gb[d];
            gX[d] += ga[d];
          }
          VecD<DM> hx, hw;
#pragma unroll
          for (int d = 0; d < DM; ++d) {
            hx.a[d] = xs[d];
            hw.a[d] = g_g[d];
          }
          const VecD<DM> hv = hvp_small<DM>(en, sh, hx, hw);
#pragma unroll
          for (int d = 0; d < DM; ++d) gX[d] += hv.a[d];
        } else {
          float kx[DM];
#pragma unroll
          for (int d = 0; d < DM; ++d) kx[d] = keep_of(t, j, d) * xs[d];
          net_eval_saved(NX, tbx + t * HM, vs, kx, S, Tt, Q, sv);
          float gxn[DM];
#pragma unroll
          for (int d = 0; d < DM; ++d) {
            const float k = keep_of(t, j, d), upd = 1.f - k;
            const float s = sign * eps * S[d], f = eps * Q[d];
            const float es = expf(s), ef = expf(f);
            const float inner = ef * vs[d] + Tt[d];
            const float gu = gX[d] * upd;
            gxn[d] = gX[d] * k + gu * es;
            const float g_s = (fwd ? gu * xs[d] * es : gu * es * (xs[d] - eps * inner)) + glj * upd;
            const float g_shift = fwd ? gu : -gu * es;
            const float g_inner = g_shift * eps;
            deps_c += g_shift * inner;
            const float g_f = g_inner * ef * vs[d];
            gV[d] += g_inner * ef;
            gTt[d] = g_inner;
            gS[d] = g_s * (sign * eps);
            deps_c += g_s * sign * S[d];
            gQ[d] = g_f * eps;
            deps_c += g_f * Q[d];
          }
          net_vjp(NX, sv, gS, gTt, gQ, ct, st, GX, ga, gb);
#pragma unroll
          for (int d = 0; d < DM; ++d) {
            gX[d] = gxn[d] + keep_of(t, j, d) * gb[d];
            gV[d] += ga[d];
          }
        }
      }
    }
  }

  // ---- reduce: warp shuffles -> shared memory -> one atomicAdd per parameter and block ---------------------------------
  const int lane = threadIdx.x & 31;
  {
    for (int i = 0; i < NP; ++i) {
      float a = GX.g[i], b = GV.g[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if (lane == 0) {
        atomicAdd(&red[i], a);
        atomicAdd(&red[NP + i], b);
      }
    }
    float l = loss_c, e = deps_c;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      l += __shfl_xor_sync(0xffffffffu, l, o);
      e += __shfl_xor_sync(0xffffffffu, e, o);
    }
    if (lane == 0) {
      atomicAdd(&red[2 * NP], l);
      atomicAdd(&red[2 * NP + 1], e);
    }
  }
  __syncthreads();
  // scatter the padded block sums to the reference-layout gradient tensors
  auto emit = [&](const float *r, const NetRaw &Gd) {
    using G_ = NetG<DM, HM>;
    for (int i = threadIdx.x; i < DM * HM; i += NT) {
      const int d = i / HM, j = i - d * HM;
      if (d < D && j < H) {
        atomicAdd(const_cast<float *>(Gd.W1) + d * H + j, r[G_::oW1 + d * HM + j]);
        atomicAdd(const_cast<float *>(Gd.W2) + d * H + j, r[G_::oW2 + d * HM + j]);
      }
      const int jj = i / DM, dd = i - jj * DM;
      if (jj < H && dd < D) {
        atomicAdd(const_cast<float *>(Gd.Ws) + jj * D + dd, r[G_::oWs + jj * DM + dd]);
        atomicAdd(const_cast<float *>(Gd.Wt) + jj * D + dd, r[G_::oWt + jj * DM + dd]);
        atomicAdd(const_cast<float *>(Gd.Wq) + jj * D + dd, r[G_::oWq + jj * DM + dd]);
      }
    }
    for (int i = threadIdx.x; i < HM * HM; i += NT) {
      const int a = i / HM, b = i - a * HM;
      if (a < H && b < H) atomicAdd(const_cast<float *>(Gd.W4) + a * H + b, r[G_::oW4 + a * HM + b]);
    }
    for (int i = threadIdx.x; i < HM; i += NT)
      if (i < H) {
        atomicAdd(const_cast<float *>(Gd.b1) + i, r[G_::ob123 + i]);
        atomicAdd(const_cast<float *>(Gd.b2) + i, r[G_::ob123 + i]);
        atomicAdd(const_cast<float *>(Gd.b3) + i, r[G_::ob123 + i]);
        atomicAdd(const_cast<float *>(Gd.b4) + i, r[G_::ob4 + i]);
        atomicAdd(const_cast<float *>(Gd.W3) + i, r[G_::oW3 + i]);
        atomicAdd(const_cast<float *>(Gd.W3) + H + i, r[G_::oW3 + HM + i]);
      }
    for (int i = threadIdx.x; i < DM; i += NT)
      if (i < D) {
        atomicAdd(const_cast<float *>(Gd.bs) + i, r[G_::obs + i]);
        atomicAdd(const_cast<float *>(Gd.bt) + i, r[G_::obt + i]);
        atomicAdd(const_cast<float *>(Gd.bq) + i, r[G_::obq + i]);
        atomicAdd(const_cast<float *>(Gd.ls) + i, r[G_::ols + i]);
        atomicAdd(const_cast<float *>(Gd.lq) + i, r[G_::olq + i]);
      }
  };
  emit(red, io.gx);
  emit(red + NP, io.gv);
  if (threadIdx.x == 0) {
    atomicAdd(io.loss, red[2 * NP]);
    atomicAdd(io.d_eps, red[2 * NP + 1]);
  }
}

}  // namespace small
}  // namespace l2hmc
