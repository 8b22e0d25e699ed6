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
//   xg  [2*DP][M]  rows 0..DP-1 = x, rows DP..2DP-1 = grad U(x)     -> VNet input [x | g]
//   vx  [2*DP][M]  rows 0..DP-1 = v, rows DP..2DP-1 = k (.) x       -> XNet input [v | masked x]
//   h   [HP][M]    hidden activations (h1 then h2 in place)
//   x0  [DP][M]    x at the start of the transition (for tf_accept)
//   ljs [DP/2][M]  per-thread partial sums of log|J| (one row per 2-dim column group)
//   wst            per-warp double-buffered weight slabs (cp.async from L2)
//
// Register tiling: thread (rg = lane & 7, cg = 4*warp + lane>>3) owns chains 4*rg..+3 and 32+4*rg..+3 and
// output columns 4*cg..4*cg+3 (embed / hidden / grad GEMMs) or dims 2*cg, 2*cg+1 x {S,T,Q} (heads),
// so the S/T/Q epilogue and the state update for a (chain, dim) pair happen in the thread that
// accumulated it.  Each chain runs only its selected direction; the direction is an elementwise
// predicate, so mixed-direction tiles do not diverge in the GEMMs.
//
// Every phase is a separate (noinline) function: the heads GEMM needs 48 accumulators per thread and the
// whole transition inlined into one body made ptxas spill them inside the FMA loop (profiles/r01).
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

// Offsets (in floats) of the shared-memory regions.
struct Lay {
  int xg, vx, h, x0, ljs, wst, smask, h0, su, sdir, sacc, end;
};

__host__ __device__ inline Lay make_lay(int DP, int HP, int T) {
  Lay l;
  l.xg = 0;
  l.vx = l.xg + 2 * DP * M;
  l.h = l.vx + 2 * DP * M;
  l.x0 = l.h + HP * M;
  l.ljs = l.x0 + DP * M;
  l.wst = l.ljs + (DP / 2) * M;
  l.smask = l.wst + WST_FLOATS;
  l.h0 = l.smask + T * DP;
  l.su = l.h0 + M;
  l.sdir = l.su + M;
  l.sacc = l.sdir + M;
  l.end = l.sacc + M;
  return l;
}

__host__ __device__ inline size_t smem_bytes(int DP, int HP, int T) { return sizeof(float) * (size_t)make_lay(DP, HP, T).end; }

extern __shared__ __align__(16) float smem[];

struct Tid {
  int tid, warp, lane, rg, cg;
  __device__ __forceinline__ Tid() {
    tid = threadIdx.x;
    warp = tid >> 5;
    lane = tid & 31;
    rg = lane & 7;
    cg = 4 * warp + (lane >> 3);
  }
};

// Chains owned by thread row-group rg: 4*rg..4*rg+3 and 32+4*rg..32+4*rg+3, so that the two float4
// activation loads of a warp's 8 row groups each cover 128 contiguous bytes (no bank conflicts).
__device__ __forceinline__ int chain_of(int rg, int i) { return (i < 4 ? 4 * rg : 28 + 4 * rg) + i; }

__device__ __forceinline__ unsigned dir_mask(const Lay &L, int rg) {
  const int4 a = *reinterpret_cast<const int4 *>(smem + L.sdir + 4 * rg);
  const int4 b = *reinterpret_cast<const int4 *>(smem + L.sdir + 32 + 4 * rg);
  return (a.x ? 1u : 0u) | (a.y ? 2u : 0u) | (a.z ? 4u : 0u) | (a.w ? 8u : 0u) | (b.x ? 16u : 0u) |
         (b.y ? 32u : 0u) | (b.z ? 64u : 0u) | (b.w ? 128u : 0u);
}

template <int TN, int ROWS>
__device__ __forceinline__ void stage_slab(const float *__restrict__ Wg, int ldw, int col0, int k0, float *dst, int lane) {
  constexpr int CNT = ROWS * TN;  // float4 copies in this slab (TN float4 per row)
#pragma unroll
  for (int base = 0; base < CNT; base += 32) {
    const int i = base + lane;
    if (CNT - base >= 32 || i < CNT) {
      const int r = i / TN, q = i - r * TN;
      cp_async16(dst + r * WS + q * 4, Wg + (size_t)(k0 + r) * ldw + col0 + q * 4);
    }
  }
  cp_async_commit();
}

template <int TN, int ROWS>
__device__ __forceinline__ void fma_slab(float (&acc)[TM][TN], const float *ab, const float *wb) {
#pragma unroll
  for (int kk = 0; kk < ROWS; ++kk) {
    const float4 a0 = *reinterpret_cast<const float4 *>(ab + kk * M);
    const float4 a1 = *reinterpret_cast<const float4 *>(ab + kk * M + 32);
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

// acc[i][j] += sum_k in[k][chain_of(rg,i)] * W[k][col0 + cgl*TN + j].  K % 4 == 0: full slabs of 8 k-rows
// plus one optional slab of 4.  Weights are streamed per warp with cp.async (double buffered), so the
// K loop needs no CTA-wide barrier.  sIn points at row 0, column 4*rg of the input rows.
template <int TN>
__device__ __forceinline__ void gemm_core(float (&acc)[TM][TN], const float *__restrict__ Wg, int ldw,
                                          int col0, int K, const float *sIn, float *wbuf, int lane) {
  const int cgl = lane >> 3;
  const int nfull = K >> 3;
  const bool tail = (K & 4) != 0;
  if (nfull > 0) stage_slab<TN, 8>(Wg, ldw, col0, 0, wbuf, lane);
  else stage_slab<TN, 4>(Wg, ldw, col0, 0, wbuf, lane);
  for (int c = 0; c < nfull; ++c) {
    float *nxt = wbuf + ((c + 1) & 1) * KC * WS;
    if (c + 1 < nfull) {
      stage_slab<TN, 8>(Wg, ldw, col0, (c + 1) * KC, nxt, lane);
      cp_async_wait<1>();
    } else if (tail) {
      stage_slab<TN, 4>(Wg, ldw, col0, (c + 1) * KC, nxt, lane);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
    fma_slab<TN, 8>(acc, sIn + (size_t)c * KC * M, wbuf + (c & 1) * KC * WS + cgl * TN);
    __syncwarp();
  }
  if (tail) {
    cp_async_wait<0>();
    __syncwarp();
    fma_slab<TN, 4>(acc, sIn + (size_t)nfull * KC * M, wbuf + (nfull & 1) * KC * WS + cgl * TN);
    __syncwarp();
  }
}

__device__ __forceinline__ void store_relu(float *o, const float (&acc)[TM][4], int j) {
  *reinterpret_cast<float4 *>(o) = make_float4(fmaxf(acc[0][j], 0.f), fmaxf(acc[1][j], 0.f), fmaxf(acc[2][j], 0.f), fmaxf(acc[3][j], 0.f));
  *reinterpret_cast<float4 *>(o + 32) = make_float4(fmaxf(acc[4][j], 0.f), fmaxf(acc[5][j], 0.f), fmaxf(acc[6][j], 0.f), fmaxf(acc[7][j], 0.f));
}

// ---- embed: h = relu([a|b] Wemb + tb[t_chain]) -------------------------------------------------
// which: 0 = XNet (input rows vx), 1 = VNet (input rows xg)
__device__ __noinline__ void phase_embed(const KernelArgs &A, int which, int it) {
  const Shape &sh = A.sh;
  const NetDev &net = which ? A.vnet : A.xnet;
  const Lay L = make_lay(sh.DP, sh.HP, sh.T);
  const Tid t;
  if (16 * t.warp < sh.HP) {  // warp-uniform: gemm_core uses __syncwarp
    float acc[TM][4];
    const int col = 4 * t.cg;  // < LDE always (padded)
    const unsigned dmask = dir_mask(L, t.rg);
    const float4 tf = *reinterpret_cast<const float4 *>(net.tb + (size_t)it * sh.LDE + col);
    const float4 tbk = *reinterpret_cast<const float4 *>(net.tb + (size_t)(sh.T - 1 - it) * sh.LDE + col);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const bool f = (dmask >> i) & 1u;
      acc[i][0] = f ? tf.x : tbk.x;
      acc[i][1] = f ? tf.y : tbk.y;
      acc[i][2] = f ? tf.z : tbk.z;
      acc[i][3] = f ? tf.w : tbk.w;
    }
    gemm_core<4>(acc, net.Wemb, sh.LDE, 16 * t.warp, 2 * sh.DP, smem + (which ? L.xg : L.vx) + 4 * t.rg,
                 smem + L.wst + t.warp * 2 * KC * WS, t.lane);
    if (col < sh.HP) {
#pragma unroll
      for (int j = 0; j < 4; ++j) store_relu(smem + L.h + (col + j) * M + 4 * t.rg, acc, j);
    }
  }
  __syncthreads();
}

// ---- hidden: h = relu(h W4 + b4), in place -----------------------------------------------------
__device__ __noinline__ void phase_hidden(const KernelArgs &A, int which) {
  const Shape &sh = A.sh;
  const NetDev &net = which ? A.vnet : A.xnet;
  const Lay L = make_lay(sh.DP, sh.HP, sh.T);
  const Tid t;
  float acc[TM][4];
  const int col = 4 * t.cg;
  if (16 * t.warp < sh.HP) {  // warp-uniform
    const float4 b = *reinterpret_cast<const float4 *>(net.b4 + col);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      acc[i][0] = b.x; acc[i][1] = b.y; acc[i][2] = b.z; acc[i][3] = b.w;
    }
    gemm_core<4>(acc, net.W4, sh.LDE, 16 * t.warp, sh.HP, smem + L.h + 4 * t.rg,
                 smem + L.wst + t.warp * 2 * KC * WS, t.lane);
  }
  __syncthreads();  // everyone has finished reading h1
  if (col < sh.HP) {
#pragma unroll
    for (int j = 0; j < 4; ++j) store_relu(smem + L.h + (col + j) * M + 4 * t.rg, acc, j);
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
// MODE 0: VNet heads -> momentum half step.  MODE 1: XNet heads -> masked position update (half 0 / 1).
template <int MODE>
__device__ __noinline__ void phase_heads(const KernelArgs &A, int it, int half) {
  const Shape &sh = A.sh;
  const NetDev &net = (MODE == 0) ? A.vnet : A.xnet;
  const Lay L = make_lay(sh.DP, sh.HP, sh.T);
  const Tid t;
  const int d0 = 2 * t.cg;
  if (8 * t.warp < sh.DP) {  // warp-uniform
    float acc[TM][6];
    {
      const float2 b0 = *reinterpret_cast<const float2 *>(net.bh + 6 * t.cg);
      const float2 b1 = *reinterpret_cast<const float2 *>(net.bh + 6 * t.cg + 2);
      const float2 b2 = *reinterpret_cast<const float2 *>(net.bh + 6 * t.cg + 4);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        acc[i][0] = b0.x; acc[i][1] = b0.y; acc[i][2] = b1.x;
        acc[i][3] = b1.y; acc[i][4] = b2.x; acc[i][5] = b2.y;
      }
    }
    gemm_core<6>(acc, net.Wh, sh.LDH, 24 * t.warp, sh.HP, smem + L.h + 4 * t.rg,
                 smem + L.wst + t.warp * 2 * KC * WS, t.lane);
    if (d0 < sh.DP) {  // DP is even: both dims of the pair or none
      const float eps = sh.eps;
      const unsigned dmask = dir_mask(L, t.rg);
      float lj[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) lj[i] = 0.f;
#pragma unroll
      for (int dd = 0; dd < 2; ++dd) {
        const int d = d0 + dd;
        const float es = net.es[d], eq = net.eq[d];
        float *px = smem + (MODE == 0 ? L.vx : L.xg) + d * M + 4 * t.rg;                         // updated row (v or x)
        const float *po = smem + (MODE == 0 ? L.xg + (sh.DP + d) * M : L.vx + d * M) + 4 * t.rg;  // grad row or v row
        const float4 s0 = *reinterpret_cast<float4 *>(px), s1 = *reinterpret_cast<float4 *>(px + 32);
        const float4 o0 = *reinterpret_cast<const float4 *>(po), o1 = *reinterpret_cast<const float4 *>(po + 32);
        float sv[TM] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const float ov[TM] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
        float mF = 0.f, mB = 0.f;
        if (MODE == 1) {
          mF = smem[L.smask + it * sh.DP + d];
          mB = smem[L.smask + (sh.T - 1 - it) * sh.DP + d];
        }
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          const bool fwd = (dmask >> i) & 1u;
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
        *reinterpret_cast<float4 *>(px + 32) = make_float4(sv[4], sv[5], sv[6], sv[7]);
      }
      // per-thread running sum of log|J| (this thread is the only writer of its 8 slots)
      float *pl = smem + L.ljs + t.cg * M + 4 * t.rg;
      float4 l0 = *reinterpret_cast<float4 *>(pl), l1 = *reinterpret_cast<float4 *>(pl + 32);
      l0.x += lj[0]; l0.y += lj[1]; l0.z += lj[2]; l0.w += lj[3];
      l1.x += lj[4]; l1.y += lj[5]; l1.z += lj[6]; l1.w += lj[7];
      *reinterpret_cast<float4 *>(pl) = l0;
      *reinterpret_cast<float4 *>(pl + 32) = l1;
    }
  }
  __syncthreads();
}

// ---- hmc=True: nets are zero (utils/dynamics.py:73-76) -> plain elementwise update ---------------
template <int MODE>
__device__ __noinline__ void phase_hmc(const KernelArgs &A, int it, int half) {
  const Shape &sh = A.sh;
  const Lay L = make_lay(sh.DP, sh.HP, sh.T);
  for (int i = threadIdx.x; i < sh.DP * M; i += NT) {
    const int d = i / M, ch = i - d * M;
    const bool fwd = reinterpret_cast<const int *>(smem + L.sdir)[ch] != 0;
    float k = 0.f, uu = 0.f;
    if (MODE == 1) {
      const float m = fwd ? smem[L.smask + it * sh.DP + d] : smem[L.smask + (sh.T - 1 - it) * sh.DP + d];
      k = (fwd == (half == 0)) ? m : 1.f - m;
      uu = 1.f - k;
    }
    float *px = smem + (MODE == 0 ? L.vx : L.xg) + i;
    const float o = (MODE == 0) ? smem[L.xg + sh.DP * M + i] : smem[L.vx + i];
    float xv = *px;
    update_elem<MODE>(fwd, sh.eps, 0.f, 0.f, 0.f, xv, o, k, uu);  // log|J| contribution is 0
    *px = xv;
  }
  __syncthreads();
}

// ---- masked copy of x for the XNet input: vx rows DP.. = k (.) x ---------------------------------
__device__ __noinline__ void build_xm(const KernelArgs &A, int it, int half) {
  const Shape &sh = A.sh;
  const Lay L = make_lay(sh.DP, sh.HP, sh.T);
  for (int i = threadIdx.x; i < sh.DP * M; i += NT) {
    const int d = i / M, ch = i - d * M;
    const bool fwd = reinterpret_cast<const int *>(smem + L.sdir)[ch] != 0;
    const float m = fwd ? smem[L.smask + it * sh.DP + d] : smem[L.smask + (sh.T - 1 - it) * sh.DP + d];
    const float k = (fwd == (half == 0)) ? m : 1.f - m;
    smem[L.vx + sh.DP * M + i] = k * smem[L.xg + i];
  }
  __syncthreads();
}

// ---- grad U(x) -> xg rows DP..2DP-1 ---------------------------------------------------------------
__device__ __noinline__ void phase_grad(const KernelArgs &A) {
  const Shape &sh = A.sh;
  const EnergyDev &en = A.en;
  const Lay L = make_lay(sh.DP, sh.HP, sh.T);
  const Tid t;
  if (en.kind == 0) {
    // d = x - mu into the (currently dead) masked-x rows, then g = d Ssym as a tile GEMM
    for (int i = t.tid; i < sh.DP * M; i += NT) smem[L.vx + sh.DP * M + i] = smem[L.xg + i] - en.mu[i / M];
    __syncthreads();
    const int col = 4 * t.cg;
    if (16 * t.warp < sh.DP) {  // warp-uniform
      float acc[TM][4];
#pragma unroll
      for (int i = 0; i < TM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
      gemm_core<4>(acc, en.Ssym, sh.LDS, 16 * t.warp, sh.DP, smem + L.vx + sh.DP * M + 4 * t.rg,
                   smem + L.wst + t.warp * 2 * KC * WS, t.lane);
      const float T = en.temperature;
      if (col < sh.DP) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float *o = smem + L.xg + (sh.DP + col + j) * M + 4 * t.rg;
          *reinterpret_cast<float4 *>(o) = make_float4(acc[0][j] / T, acc[1][j] / T, acc[2][j] / T, acc[3][j] / T);
          *reinterpret_cast<float4 *>(o + 32) = make_float4(acc[4][j] / T, acc[5][j] / T, acc[6][j] / T, acc[7][j] / T);
        }
      }
    }
  } else if (en.kind == 2) {
    const float e = en.s0, den = en.s1;
    for (int i = t.tid; i < sh.DP * M; i += NT) {
      const float xi = smem[L.xg + i];
      smem[L.xg + sh.DP * M + i] = (i / M < sh.D) ? (xi - e * sinf(xi / den) / den) / en.temperature : 0.f;
    }
  } else {
    if (t.tid < M) grad_chain(en, sh, smem + L.xg + t.tid, M, smem + L.xg + sh.DP * M + t.tid, M);
  }
  __syncthreads();
}

// U(x) + 0.5|v|^2 for chain `ch` from the tile state; for the Gaussian kind it reuses d = x - mu
// (vx rows DP..) and g = d Ssym / T (xg rows DP..) left by the last phase_grad on the same x.
__device__ __noinline__ float hamiltonian_chain(const KernelArgs &A, int ch, bool v0_chain = false) {
  const Shape &sh = A.sh;
  const EnergyDev &en = A.en;
  const Lay L = make_lay(sh.DP, sh.HP, sh.T);
  float U;
  if (en.kind == 0) {
    float q = 0.f;
    for (int d = 0; d < sh.D; ++d) q = fmaf(smem[L.vx + (sh.DP + d) * M + ch], smem[L.xg + (sh.DP + d) * M + ch], q);
    U = 0.5f * q;  // g already carries 1/temperature
  } else {
    U = energy_chain(en, sh, smem + L.xg + ch, M);
  }
  float kin = 0.f;
  if (v0_chain) {
    // chain mode, first Hamiltonian: H(x_in, init_v) -- the sub-proposals draw their own momenta (utils/sampler.py:35-36, 79)
    const TransitionIO &io = A.io;
    const long long g = (long long)blockIdx.x * M + ch;
    for (int b = 0; b < sh.DP / 4; ++b) {
      float z[4] = {0.f, 0.f, 0.f, 0.f};
      if (io.v0 == nullptr) philox_normals4(io.seed, chain_v0_counter(io), io.chain_offset + g, b, z);
      for (int q = 0; q < 4; ++q) {
        const int d = 4 * b + q;
        const float v = (g < io.n && d < sh.D) ? (io.v0 ? io.v0[g * sh.D + d] : z[q]) : 0.f;
        kin = fmaf(v, v, kin);
      }
    }
  } else {
    for (int d = 0; d < sh.D; ++d) {
      const float v = smem[L.vx + d * M + ch];
      kin = fmaf(v, v, kin);
    }
  }
  return U + 0.5f * kin;
}

// ---- per-transition setup: x0, momentum, direction bit, accept uniform ------------------------------
__device__ __noinline__ void phase_begin(const KernelArgs &A, int tr) {
  const Shape &sh = A.sh;
  const TransitionIO &io = A.io;
  const Lay L = make_lay(sh.DP, sh.HP, sh.T);
  const int tid = threadIdx.x, D = sh.D, DP = sh.DP;
  const long long base = (long long)blockIdx.x * M;
  const unsigned long long ctr = io.counter + (unsigned long long)tr;
  if (!io.chain || tr == 0) {  // chain mode: one start point and one log|J| accumulator for all sub-proposals
    for (int i = tid; i < DP * M; i += NT) smem[L.x0 + i] = smem[L.xg + i];
    for (int i = tid; i < (DP / 2) * M; i += NT) smem[L.ljs + i] = 0.f;
  }
  if (io.v != nullptr) {
    for (int i = tid; i < M * DP; i += NT) {
      const int ch = i / DP, d = i - ch * DP;
      const long long g = base + ch;
      smem[L.vx + d * M + ch] = (g < io.n && d < D) ? io.v[((long long)tr * io.n + g) * D + d] : 0.f;
    }
  } else {
    for (int i = tid; i < M * (DP / 4); i += NT) {
      const int ch = i / (DP / 4), b = i - ch * (DP / 4);
      const long long g = base + ch;
      float z[4];
      philox_normals4(io.seed, ctr, io.chain_offset + g, b, z);
#pragma unroll
      for (int q = 0; q < 4; ++q) smem[L.vx + (4 * b + q) * M + ch] = (g < io.n && 4 * b + q < D) ? z[q] : 0.f;
    }
  }
  if (tid < M) {
    const long long g = base + tid;
    int pd = 1;
    float pu = 0.f;
    if (io.dir_mode == 3 || (io.do_mh && io.u == nullptr)) philox_dir_u(io.seed, ctr, io.chain_offset + g, pd, pu);
    int dbit = 1;
    if (io.dir_mode == 1) dbit = 0;
    else if (io.dir_mode == 2) dbit = (g < io.n) ? (io.dir[(long long)tr * io.n + g] != 0) : 1;
    else if (io.dir_mode == 3) dbit = pd;
    reinterpret_cast<int *>(smem + L.sdir)[tid] = dbit;
    if (io.do_mh && io.u != nullptr) pu = (g < io.n) ? io.u[(io.chain ? 0ll : (long long)tr * io.n) + g] : 0.f;
    smem[L.su + tid] = pu;
  }
  __syncthreads();
}

// ---- log|J| reduction, Hamiltonian difference, accept, outputs ----------------------------------------
__device__ __noinline__ void phase_end(const KernelArgs &A, int tr) {
  const Shape &sh = A.sh;
  const TransitionIO &io = A.io;
  const Lay L = make_lay(sh.DP, sh.HP, sh.T);
  const int tid = threadIdx.x, D = sh.D, DP = sh.DP;
  const long long base = (long long)blockIdx.x * M;
  const bool last = (tr == io.n_transitions - 1);
  if (io.chain && !last) return;  // chain mode: the next sub-proposal starts from this proposal, no Metropolis step in between
  int *sacc = reinterpret_cast<int *>(smem + L.sacc);
  if (tid < M) {
    const int ch = tid;
    const long long g = base + ch;
    float logj = 0.f;
    if (!sh.hmc)
      for (int q = 0; q < DP / 2; ++q) logj += smem[L.ljs + q * M + ch];  // fixed order
    const float h1 = hamiltonian_chain(A, ch);
    const float p = accept_prob(smem[L.h0 + ch], h1, logj);
    const float px = io.log_jac ? logj : p;
    int acc = 0;
    if (io.do_mh) acc = (px - smem[L.su + ch] >= 0.f) ? 1 : 0;  // tf_accept, utils/sampler.py:53-55
    sacc[ch] = acc;
    if (g < io.n) stats_add(io.stats, px, acc);
    if (g < io.n && last) {
      io.px_out[g] = px;
      if (io.accepted) io.accepted[g] = (uint8_t)acc;
    }
  }
  __syncthreads();
  if (io.trace) {
    for (int i = tid; i < M * D; i += NT) {
      const int ch = i / D, d = i - ch * D;
      const long long g = base + ch;
      if (g < io.n) io.trace[((long long)tr * io.n + g) * D + d] = sacc[ch] ? smem[L.xg + d * M + ch] : smem[L.x0 + d * M + ch];
    }
  }
  if (last) {
    for (int i = tid; i < M * D; i += NT) {
      const int ch = i / D, d = i - ch * D;
      const long long g = base + ch;
      if (g < io.n) {
        const float lx = smem[L.xg + d * M + ch];
        io.x_out[g * D + d] = lx;
        if (io.v_out) io.v_out[g * D + d] = smem[L.vx + d * M + ch];
        if (io.do_mh) io.x_next[g * D + d] = sacc[ch] ? lx : smem[L.x0 + d * M + ch];
      }
    }
  } else {
    for (int i = tid; i < DP * M; i += NT)
      if (!sacc[i % M]) smem[L.xg + i] = smem[L.x0 + i];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(NT, 2) transition_kernel(const __grid_constant__ KernelArgs A) {
  const Shape &sh = A.sh;
  const TransitionIO &io = A.io;
  const Lay L = make_lay(sh.DP, sh.HP, sh.T);
  const int tid = threadIdx.x;
  const long long base = (long long)blockIdx.x * M;
  const int D = sh.D, DP = sh.DP;

  for (int i = tid; i < sh.T * DP; i += NT) smem[L.smask + i] = A.mask[i];
  // x of the first transition (padded dims / chains are zero and stay zero)
  for (int i = tid; i < M * DP; i += NT) {
    const int ch = i / DP, d = i - ch * DP;
    const long long g = base + ch;
    smem[L.xg + d * M + ch] = (g < io.n && d < D) ? io.x[g * D + d] : 0.f;
  }
  __syncthreads();

  for (int tr = 0; tr < io.n_transitions; ++tr) {
    phase_begin(A, tr);
    phase_grad(A);
    if (tid < M && (!io.chain || tr == 0)) smem[L.h0 + tid] = hamiltonian_chain(A, tid, io.chain != 0);
    __syncthreads();

    for (int it = 0; it < sh.T; ++it) {
      if (sh.hmc) {
        phase_hmc<0>(A, it, 0);
        phase_hmc<1>(A, it, 0);
        phase_hmc<1>(A, it, 1);
        phase_grad(A);
        phase_hmc<0>(A, it, 0);
      } else {
        // v half step: VNet([x, grad U(x), t])
        phase_embed(A, 1, it);
        phase_hidden(A, 1);
        phase_heads<0>(A, it, 0);
        // first masked x update: XNet([v_h, k1 (.) x, t])
        build_xm(A, it, 0);
        phase_embed(A, 0, it);
        phase_hidden(A, 0);
        phase_heads<1>(A, it, 0);
        // second masked x update
        build_xm(A, it, 1);
        phase_embed(A, 0, it);
        phase_hidden(A, 0);
        phase_heads<1>(A, it, 1);
        // v half step at the new x
        phase_grad(A);
        phase_embed(A, 1, it);
        phase_hidden(A, 1);
        phase_heads<0>(A, it, 0);
      }
    }
    phase_end(A, tr);
  }
}

}  // namespace tile
}  // namespace l2hmc
