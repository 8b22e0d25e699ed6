// Generic fp32-FMA tile kernel: one CTA owns 64 chains for the WHOLE transition
// (T leapfrog steps + Hamiltonians + accept), so x, v, grad U, the hidden activations and
// log|J| never leave the SM between steps.
//
// Follows (does not translate) the reference math:
//   _forward_step / _backward_step   utils/dynamics.py:115-157 / :159-201
//   forward / backward loops         utils/dynamics.py:246-300
//   p_accept                         utils/dynamics.py:302-309
//   propose / tf_accept              utils/sampler.py:28-55
//   S/T/Q net                        SCGExperiment.ipynb:51-77, utils/layers.py:29-37,81-95
//
// Layout in shared memory (floats), M = 64 chains, "row" = one feature for all 64 chains:
//   xg [2*DP][M]   rows 0..DP-1 = x, rows DP..2DP-1 = grad U(x)     -> VNet input [x | g]
//   vx [2*DP][M]   rows 0..DP-1 = v, rows DP..2DP-1 = k (.) x       -> XNet input [v | masked x]
//   h  [HP][M]     hidden activations (h1 then h2 in place)
//   x0 [DP][M]     x at the start of the transition (for tf_accept)
//   wst            per-warp double-buffered weight slabs (cp.async from L2)
//
// Register tiling: thread (rg = lane & 7, cg = 4*warp + lane>>3) owns chains 8*rg..8*rg+7 and
// output columns 4*cg..4*cg+3 (embed / hidden / grad GEMMs) or dims 2*cg, 2*cg+1 x {S,T,Q} (heads),
// so the S/T/Q epilogue and the state update for a (chain, dim) pair happen in the thread that
// accumulated it.  Each chain runs only its selected direction; the direction is an elementwise
// predicate, so mixed-direction tiles do not diverge in the GEMMs.
#pragma once
#include "common.cuh"

namespace l2hmc {
namespace tile {

constexpr int M = 64;     // chains per CTA
constexpr int NT = 256;   // threads per CTA
constexpr int TM = 8;     // chains per thread
constexpr int KC = 8;     // k-rows per staged weight slab
constexpr int WS = 24;    // floats per staged row per warp (4 column groups x up to 6)
constexpr int WST_FLOATS = 8 /*warps*/ * 2 * KC * WS;

__host__ __device__ inline size_t smem_bytes(int DP, int HP, int T) {
  return sizeof(float) * ((size_t)(5 * DP + HP) * M + WST_FLOATS + (size_t)T * DP + 4 * M);
}

// acc[i][j] += sum_k in[k][8*rg + i] * W[k][col0 + cgl*TN + j]; weights streamed per warp with cp.async.
template <int TN>
__device__ __forceinline__ void gemm_core(float (&acc)[TM][TN], const float *__restrict__ Wg, int ldw,
                                          int col0, int K, const float *sIn, float *wbuf, int lane) {
  constexpr int V4 = TN;  // float4 per staged row = 4*TN/4
  const int cgl = lane >> 3;
  const int nchunks = (K + KC - 1) / KC;
  auto stage = [&](int c, int b) {
    const int k0 = c * KC;
    const int nk = min(KC, K - k0);
    for (int i = lane; i < nk * V4; i += 32) {
      const int r = i / V4, q = i - r * V4;
      cp_async16(wbuf + (b * KC + r) * WS + q * 4, Wg + (size_t)(k0 + r) * ldw + col0 + q * 4);
    }
    cp_async_commit();
  };
  stage(0, 0);
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) {
      stage(c + 1, (c + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
    const float *wb = wbuf + (c & 1) * KC * WS + cgl * TN;
    const float *ab = sIn + (size_t)c * KC * M;
    const int nk = min(KC, K - c * KC);
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      if (kk < nk) {
        const float4 a0 = *reinterpret_cast<const float4 *>(ab + kk * M);
        const float4 a1 = *reinterpret_cast<const float4 *>(ab + kk * M + 4);
        const float a[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float w[TN];
        if (TN == 4) {
          const float4 t = *reinterpret_cast<const float4 *>(wb + kk * WS);
          w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
        } else {
#pragma unroll
          for (int j = 0; j < TN; j += 2) {
            const float2 t = *reinterpret_cast<const float2 *>(wb + kk * WS + j);
            w[j] = t.x;
            w[j + 1] = t.y;
          }
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
      }
    }
    __syncwarp();
  }
}

struct Ctx {
  const KernelArgs &A;
  float *xg, *vx, *h, *x0, *wst, *smask, *h0, *su;
  int *sdir;
  int tid, warp, lane, rg, cg;
  unsigned dmask;  // bit i: chain 8*rg+i runs forward
  __device__ Ctx(const KernelArgs &a) : A(a) {}
};

// ---- embed: h = relu([a|b] Wemb + tb[t_chain]) -------------------------------------------------
__device__ __forceinline__ void phase_embed(Ctx &c, const NetDev &net, const float *sIn, int it) {
  const Shape &sh = c.A.sh;
  if (16 * c.warp < sh.HP) {  // warp-uniform: gemm_core uses __syncwarp
    float acc[TM][4];
    const int col = 4 * c.cg;  // < LDE always (padded)
    const float4 tf = *reinterpret_cast<const float4 *>(net.tb + (size_t)it * sh.LDE + col);
    const float4 tbk = *reinterpret_cast<const float4 *>(net.tb + (size_t)(sh.T - 1 - it) * sh.LDE + col);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const bool f = (c.dmask >> i) & 1u;
      acc[i][0] = f ? tf.x : tbk.x;
      acc[i][1] = f ? tf.y : tbk.y;
      acc[i][2] = f ? tf.z : tbk.z;
      acc[i][3] = f ? tf.w : tbk.w;
    }
    gemm_core<4>(acc, net.Wemb, sh.LDE, 16 * c.warp, 2 * sh.DP, sIn + 8 * c.rg,
                 c.wst + c.warp * 2 * KC * WS, c.lane);
    if (col < sh.HP) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float *o = c.h + (size_t)(col + j) * M + 8 * c.rg;
        *reinterpret_cast<float4 *>(o) = make_float4(fmaxf(acc[0][j], 0.f), fmaxf(acc[1][j], 0.f),
                                                     fmaxf(acc[2][j], 0.f), fmaxf(acc[3][j], 0.f));
        *reinterpret_cast<float4 *>(o + 4) = make_float4(fmaxf(acc[4][j], 0.f), fmaxf(acc[5][j], 0.f),
                                                         fmaxf(acc[6][j], 0.f), fmaxf(acc[7][j], 0.f));
      }
    }
  }
  __syncthreads();
}

// ---- hidden: h = relu(h W4 + b4), in place -----------------------------------------------------
__device__ __forceinline__ void phase_hidden(Ctx &c, const NetDev &net) {
  const Shape &sh = c.A.sh;
  float acc[TM][4];
  const int col = 4 * c.cg;
  const bool active = col < sh.HP;
  if (16 * c.warp < sh.HP) {  // warp-uniform
    const float4 b = *reinterpret_cast<const float4 *>(net.b4 + col);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      acc[i][0] = b.x; acc[i][1] = b.y; acc[i][2] = b.z; acc[i][3] = b.w;
    }
    gemm_core<4>(acc, net.W4, sh.LDE, 16 * c.warp, sh.HP, c.h + 8 * c.rg,
                 c.wst + c.warp * 2 * KC * WS, c.lane);
  }
  __syncthreads();  // everyone has finished reading h1
  if (active) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float *o = c.h + (size_t)(col + j) * M + 8 * c.rg;
      *reinterpret_cast<float4 *>(o) = make_float4(fmaxf(acc[0][j], 0.f), fmaxf(acc[1][j], 0.f),
                                                   fmaxf(acc[2][j], 0.f), fmaxf(acc[3][j], 0.f));
      *reinterpret_cast<float4 *>(o + 4) = make_float4(fmaxf(acc[4][j], 0.f), fmaxf(acc[5][j], 0.f),
                                                       fmaxf(acc[6][j], 0.f), fmaxf(acc[7][j], 0.f));
    }
  }
  __syncthreads();
}

// One (chain, dim) update.  MODE 0: momentum half-step (utils/dynamics.py:121-125,148-153 fwd;
// :166-171,193-199 bwd).  MODE 1: masked position update (:129-145 fwd; :173-190 bwd).
// k/uu are the keep / update masks as 0/1 floats; the arithmetic keeps the reference's
// mask*old + (1-mask)*new form so non-finite values propagate the same way.
template <int MODE>
__device__ __forceinline__ float update_elem(bool fwd, float eps, float S, float Tt, float Q, float &xv,
                                             float other, float k, float uu) {
  if (MODE == 0) {
    // xv = v, other = grad
    const float sv = fwd ? (0.5f * eps) * S : (-0.5f * eps) * S;
    const float fv = eps * Q;
    const float cterm = (0.5f * eps) * (-(expf(fv) * other) + Tt);
    const float e = expf(sv);
    xv = fwd ? (xv * e + cterm) : ((xv - cterm) * e);
    return sv;
  } else {
    // xv = x, other = v_h
    const float sx = fwd ? eps * S : -eps * S;
    const float fx = eps * Q;
    const float inner = eps * (expf(fx) * other + Tt);
    const float e = expf(sx);
    const float nx = fwd ? (xv * e + inner) : (e * (xv - inner));
    xv = k * xv + uu * nx;
    return uu * sx;
  }
}

// ---- heads: [S|T|Q] = h Wh + bh, then the fused state update ------------------------------------
template <int MODE>
__device__ __forceinline__ void phase_heads(Ctx &c, const NetDev &net, int it, int half, float (&lj)[TM]) {
  const Shape &sh = c.A.sh;
  const int d0 = 2 * c.cg;
  if (8 * c.warp < sh.DP) {  // warp-uniform
    float acc[TM][6];
    {
      const float2 b0 = *reinterpret_cast<const float2 *>(net.bh + 6 * c.cg);
      const float2 b1 = *reinterpret_cast<const float2 *>(net.bh + 6 * c.cg + 2);
      const float2 b2 = *reinterpret_cast<const float2 *>(net.bh + 6 * c.cg + 4);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        acc[i][0] = b0.x; acc[i][1] = b0.y; acc[i][2] = b1.x;
        acc[i][3] = b1.y; acc[i][4] = b2.x; acc[i][5] = b2.y;
      }
    }
    gemm_core<6>(acc, net.Wh, sh.LDH, 24 * c.warp, sh.HP, c.h + 8 * c.rg,
                 c.wst + c.warp * 2 * KC * WS, c.lane);
    const float eps = sh.eps;
#pragma unroll
    for (int dd = 0; dd < 2; ++dd) {
      const int d = d0 + dd;
      if (d >= sh.DP) break;  // DP is even: both dims or none
      const float es = net.es[d], eq = net.eq[d];
      float *px = (MODE == 0 ? c.vx : c.xg) + (size_t)d * M + 8 * c.rg;         // updated row (v or x)
      const float *po = (MODE == 0 ? c.xg + (size_t)(sh.DP + d) * M : c.vx + (size_t)d * M) + 8 * c.rg;
      float4 s0 = *reinterpret_cast<float4 *>(px), s1 = *reinterpret_cast<float4 *>(px + 4);
      const float4 o0 = *reinterpret_cast<const float4 *>(po), o1 = *reinterpret_cast<const float4 *>(po + 4);
      float sv[TM] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
      const float ov[TM] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
      float mF = 0.f, mB = 0.f;
      if (MODE == 1) {
        mF = c.smask[it * sh.DP + d];
        mB = c.smask[(sh.T - 1 - it) * sh.DP + d];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const bool fwd = (c.dmask >> i) & 1u;
        const float S = es * tanhf(acc[i][3 * dd + 0]);
        const float Tt = acc[i][3 * dd + 1];
        const float Q = eq * tanhf(acc[i][3 * dd + 2]);
        float k = 0.f, uu = 0.f;
        if (MODE == 1) {
          const float m = fwd ? mF : mB;
          // fwd: first half keeps m, second keeps 1-m; bwd: first keeps 1-m, second keeps m
          k = (fwd == (half == 0)) ? m : 1.f - m;
          uu = 1.f - k;
        }
        lj[i] += update_elem<MODE>(fwd, eps, S, Tt, Q, sv[i], ov[i], k, uu);
      }
      *reinterpret_cast<float4 *>(px) = make_float4(sv[0], sv[1], sv[2], sv[3]);
      *reinterpret_cast<float4 *>(px + 4) = make_float4(sv[4], sv[5], sv[6], sv[7]);
    }
  }
  __syncthreads();
}

// ---- hmc=True: nets are zero (utils/dynamics.py:73-76) -> plain elementwise update ---------------
template <int MODE>
__device__ __forceinline__ void phase_hmc(Ctx &c, int it, int half) {
  const Shape &sh = c.A.sh;
  for (int i = c.tid; i < sh.DP * M; i += NT) {
    const int d = i / M, ch = i - d * M;
    const bool fwd = c.sdir[ch] != 0;
    float k = 0.f, uu = 0.f;
    if (MODE == 1) {
      const float m = fwd ? c.smask[it * sh.DP + d] : c.smask[(sh.T - 1 - it) * sh.DP + d];
      k = (fwd == (half == 0)) ? m : 1.f - m;
      uu = 1.f - k;
    }
    float *px = (MODE == 0 ? c.vx : c.xg) + i;
    const float o = (MODE == 0) ? c.xg[(size_t)sh.DP * M + i] : c.vx[i];
    float xv = *px;
    update_elem<MODE>(fwd, sh.eps, 0.f, 0.f, 0.f, xv, o, k, uu);  // log|J| contribution is 0
    *px = xv;
  }
  __syncthreads();
}

// ---- masked copy of x for the XNet input: vx rows DP.. = k (.) x ---------------------------------
__device__ __forceinline__ void build_xm(Ctx &c, int it, int half) {
  const Shape &sh = c.A.sh;
  for (int i = c.tid; i < sh.DP * M; i += NT) {
    const int d = i / M, ch = i - d * M;
    const bool fwd = c.sdir[ch] != 0;
    const float m = fwd ? c.smask[it * sh.DP + d] : c.smask[(sh.T - 1 - it) * sh.DP + d];
    const float k = (fwd == (half == 0)) ? m : 1.f - m;
    c.vx[(size_t)sh.DP * M + i] = k * c.xg[i];
  }
  __syncthreads();
}

// ---- grad U(x) -> xg rows DP..2DP-1 ---------------------------------------------------------------
__device__ __forceinline__ void phase_grad(Ctx &c) {
  const Shape &sh = c.A.sh;
  const EnergyDev &en = c.A.en;
  if (en.kind == 0) {
    // d = x - mu into the (currently dead) masked-x rows, then g = d Ssym as a tile GEMM
    for (int i = c.tid; i < sh.DP * M; i += NT) c.vx[(size_t)sh.DP * M + i] = c.xg[i] - en.mu[i / M];
    __syncthreads();
    const int col = 4 * c.cg;
    if (16 * c.warp < sh.DP) {  // warp-uniform
      float acc[TM][4];
#pragma unroll
      for (int i = 0; i < TM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
      gemm_core<4>(acc, en.Ssym, sh.LDS, 16 * c.warp, sh.DP, c.vx + (size_t)sh.DP * M + 8 * c.rg,
                   c.wst + c.warp * 2 * KC * WS, c.lane);
      const float T = en.temperature;
      if (col < sh.DP)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float *o = c.xg + (size_t)(sh.DP + col + j) * M + 8 * c.rg;
        *reinterpret_cast<float4 *>(o) = make_float4(acc[0][j] / T, acc[1][j] / T, acc[2][j] / T, acc[3][j] / T);
        *reinterpret_cast<float4 *>(o + 4) = make_float4(acc[4][j] / T, acc[5][j] / T, acc[6][j] / T, acc[7][j] / T);
      }
    }
  } else if (en.kind == 2) {
    const float e = en.s0, den = en.s1;
    for (int i = c.tid; i < sh.DP * M; i += NT) {
      const float xi = c.xg[i];
      c.xg[(size_t)sh.DP * M + i] = (i / M < sh.D) ? (xi - e * sinf(xi / den) / den) / en.temperature : 0.f;
    }
  } else {
    if (c.tid < M) grad_chain(en, sh, c.xg + c.tid, M, c.xg + (size_t)sh.DP * M + c.tid, M);
  }
  __syncthreads();
}

// U(x) + 0.5|v|^2 for chain `ch` from the tile state; for the Gaussian kind it reuses d = x - mu
// (vx rows DP..) and g = d Ssym / T (xg rows DP..) left by the last phase_grad on the same x.
__device__ __forceinline__ float hamiltonian_chain(Ctx &c, int ch) {
  const Shape &sh = c.A.sh;
  const EnergyDev &en = c.A.en;
  float U;
  if (en.kind == 0) {
    float q = 0.f;
    for (int d = 0; d < sh.D; ++d)
      q = fmaf(c.vx[(size_t)(sh.DP + d) * M + ch], c.xg[(size_t)(sh.DP + d) * M + ch], q);
    U = 0.5f * q;  // g already carries 1/temperature
  } else {
    U = energy_chain(en, sh, c.xg + ch, M);
  }
  float kin = 0.f;
  for (int d = 0; d < sh.D; ++d) {
    const float v = c.vx[(size_t)d * M + ch];
    kin = fmaf(v, v, kin);
  }
  return U + 0.5f * kin;
}

__global__ void __launch_bounds__(NT, 2) transition_kernel(const __grid_constant__ KernelArgs A) {
  extern __shared__ __align__(16) float smem[];
  const Shape &sh = A.sh;
  const TransitionIO &io = A.io;
  Ctx c(A);
  c.xg = smem;
  c.vx = c.xg + (size_t)2 * sh.DP * M;
  c.h = c.vx + (size_t)2 * sh.DP * M;
  c.x0 = c.h + (size_t)sh.HP * M;
  c.wst = c.x0 + (size_t)sh.DP * M;
  c.smask = c.wst + WST_FLOATS;
  c.h0 = c.smask + (size_t)sh.T * sh.DP;
  c.su = c.h0 + M;
  c.sdir = reinterpret_cast<int *>(c.su + M);
  int *sacc = c.sdir + M;
  c.tid = threadIdx.x;
  c.warp = c.tid >> 5;
  c.lane = c.tid & 31;
  c.rg = c.lane & 7;
  c.cg = 4 * c.warp + (c.lane >> 3);

  const long long base = (long long)blockIdx.x * M;
  const int D = sh.D, DP = sh.DP;

  for (int i = c.tid; i < sh.T * DP; i += NT) c.smask[i] = A.mask[i];

  // x of the first transition (padded dims / chains are zero and stay zero)
  for (int i = c.tid; i < M * DP; i += NT) {
    const int ch = i / DP, d = i - ch * DP;
    const long long g = base + ch;
    c.xg[(size_t)d * M + ch] = (g < io.n && d < D) ? io.x[g * D + d] : 0.f;
  }
  __syncthreads();

  for (int tr = 0; tr < io.n_transitions; ++tr) {
    const unsigned long long ctr = io.counter + (unsigned long long)tr;
    const bool last = (tr == io.n_transitions - 1);
    // ---- x0, momentum, direction bit, accept uniform --------------------------------------------
    for (int i = c.tid; i < DP * M; i += NT) c.x0[i] = c.xg[i];
    if (io.v != nullptr) {
      for (int i = c.tid; i < M * DP; i += NT) {
        const int ch = i / DP, d = i - ch * DP;
        const long long g = base + ch;
        c.vx[(size_t)d * M + ch] = (g < io.n && d < D) ? io.v[((long long)tr * io.n + g) * D + d] : 0.f;
      }
    } else {
      for (int i = c.tid; i < M * (DP / 4); i += NT) {
        const int ch = i / (DP / 4), b = i - ch * (DP / 4);
        const long long g = base + ch;
        float z[4];
        philox_normals4(io.seed, ctr, io.chain_offset + g, b, z);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          c.vx[(size_t)(4 * b + q) * M + ch] = (g < io.n && 4 * b + q < D) ? z[q] : 0.f;
      }
    }
    if (c.tid < M) {
      const long long g = base + c.tid;
      int pd = 1;
      float pu = 0.f;
      if (io.dir_mode == 3 || (io.do_mh && io.u == nullptr))
        philox_dir_u(io.seed, ctr, io.chain_offset + g, pd, pu);
      int dbit = 1;
      if (io.dir_mode == 1) dbit = 0;
      else if (io.dir_mode == 2) dbit = (g < io.n) ? (io.dir[(long long)tr * io.n + g] != 0) : 1;
      else if (io.dir_mode == 3) dbit = pd;
      c.sdir[c.tid] = dbit;
      if (io.do_mh && io.u != nullptr) pu = (g < io.n) ? io.u[(long long)tr * io.n + g] : 0.f;
      c.su[c.tid] = pu;
    }
    __syncthreads();
    c.dmask = 0;
#pragma unroll
    for (int i = 0; i < TM; ++i) c.dmask |= (c.sdir[8 * c.rg + i] ? 1u : 0u) << i;

    float lj[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) lj[i] = 0.f;

    phase_grad(c);
    if (c.tid < M) c.h0[c.tid] = hamiltonian_chain(c, c.tid);
    __syncthreads();

    for (int it = 0; it < sh.T; ++it) {
      if (sh.hmc) {
        phase_hmc<0>(c, it, 0);
        phase_hmc<1>(c, it, 0);
        phase_hmc<1>(c, it, 1);
        phase_grad(c);
        phase_hmc<0>(c, it, 0);
      } else {
        // v half step: VNet([x, grad U(x), t])
        phase_embed(c, A.vnet, c.xg, it);
        phase_hidden(c, A.vnet);
        phase_heads<0>(c, A.vnet, it, 0, lj);
        // first masked x update: XNet([v_h, k1 (.) x, t])
        build_xm(c, it, 0);
        phase_embed(c, A.xnet, c.vx, it);
        phase_hidden(c, A.xnet);
        phase_heads<1>(c, A.xnet, it, 0, lj);
        // second masked x update
        build_xm(c, it, 1);
        phase_embed(c, A.xnet, c.vx, it);
        phase_hidden(c, A.xnet);
        phase_heads<1>(c, A.xnet, it, 1, lj);
        // v half step at the new x
        phase_grad(c);
        phase_embed(c, A.vnet, c.xg, it);
        phase_hidden(c, A.vnet);
        phase_heads<0>(c, A.vnet, it, 0, lj);
      }
    }

    // ---- log|J|: fixed-order reduction over the 2-dim column groups (h is dead: every
    // phase_heads ends with a barrier) -------------------------------------------------------------
    if (!sh.hmc && 2 * c.cg < DP) {
#pragma unroll
      for (int i = 0; i < TM; ++i) c.h[(size_t)c.cg * M + 8 * c.rg + i] = lj[i];
    }
    __syncthreads();

    if (c.tid < M) {
      const int ch = c.tid;
      const long long g = base + ch;
      float logj = 0.f;
      if (!sh.hmc)
        for (int q = 0; q < DP / 2; ++q) logj += c.h[(size_t)q * M + ch];
      const float h1 = hamiltonian_chain(c, ch);
      const float p = accept_prob(c.h0[ch], h1, logj);
      const float px = io.log_jac ? logj : p;
      int acc = 0;
      if (io.do_mh) acc = (px - c.su[ch] >= 0.f) ? 1 : 0;  // tf_accept, utils/sampler.py:53-55
      sacc[ch] = acc;
      if (g < io.n && last) {
        io.px_out[g] = px;
        if (io.accepted) io.accepted[g] = (uint8_t)acc;
      }
    }
    __syncthreads();

    if (last) {
      for (int i = c.tid; i < M * D; i += NT) {
        const int ch = i / D, d = i - ch * D;
        const long long g = base + ch;
        if (g < io.n) {
          const float lx = c.xg[(size_t)d * M + ch];
          io.x_out[g * D + d] = lx;
          if (io.v_out) io.v_out[g * D + d] = c.vx[(size_t)d * M + ch];
          if (io.do_mh) io.x_next[g * D + d] = sacc[ch] ? lx : c.x0[(size_t)d * M + ch];
        }
      }
    } else {
      for (int i = c.tid; i < DP * M; i += NT)
        if (!sacc[i % M]) c.xg[i] = c.x0[i];
      __syncthreads();
    }
  }
}

}  // namespace tile
}  // namespace l2hmc
