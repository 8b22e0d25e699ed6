// Tensor-core transition kernel (sm_100a, tcgen05 / TMEM / TMA bulk copies).
//
// Same transition as kernel_tile.cuh (reference: utils/dynamics.py:115-201,246-309; utils/sampler.py:28-55;
// net SCGExperiment.ipynb:51-77) with the four GEMMs of every net call and the Gaussian grad-U on the 5th-gen
// tensor cores.  fp32 parity is kept with the 3xTF32 split: a = a_hi + a_lo (both tf32),
//   acc += a_lo * b_hi ;  acc += a_hi * b_lo ;  acc += a_hi * b_hi      (fp32 accumulate in TMEM)
// whose dropped term a_lo * b_lo is 2^-22 relative.
//
// One CTA = 128 chains = the 128 TMEM lanes (chain c <-> lane c, MMA M = 128), for the WHOLE transition.
//   warps 0-7 (256 compute threads): thread (c = 32*(w&3) + lane, half = w>>2) owns chain c and every second
//       4-dim / 8-column chunk.  It builds the A operand rows directly in TMEM (tcgen05.st), reads the fp32
//       accumulators back (tcgen05.ld) and does the relu / split / tanh / exp / leapfrog epilogue; x, v, grad U
//       live in shared memory feature-major, every (chain, dim) element is only ever touched by its owner.
//   warp 8 lane 0: MMA issuer  (tcgen05.mma kind::tf32, A from TMEM, B from the shared-memory ring)
//   warp 9 lane 0: TMA producer (cp.async.bulk global -> shared ring, mbarrier complete_tx)
// B operands (weights, pre-split hi/lo on the host, canonical K-major no-swizzle core-matrix layout) do not fit
// in shared memory for both nets (640 KB for config 2), so they stream from L2 through an 8-slot ring in
// consumption order; one slot = one K=8 step of one GEMM = {B_hi slab, B_lo slab}.
//
// TMEM columns: [0,192) accumulator, [192,320) A_hi, [320,448) A_lo.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace l2hmc {
namespace tc {

constexpr int MT = 128;               // chains per CTA
constexpr int NCT = 256;              // compute threads
constexpr int NTHREADS = 320;         // + MMA-issuer warp + producer warp
constexpr int MAX_SLOT = 16;          // ring slots (the host sizes the ring to what shared memory allows)
constexpr uint32_t T_ACC = 0, T_AHI = 192, T_ALO = 320;

struct TcDims {
  int K1;   // 2*DP            embed GEMM depth  (multiple of 8)
  int HK;   // H rounded to 8  hidden / heads GEMM depth
  int N1;   // HK rounded to 16: embed / hidden GEMM width
  int N3;   // 3*DP rounded to 16: heads GEMM width  (S | T | Q blocks of DP columns)
  int KG;   // DP rounded to 8: grad GEMM depth
  int NG;   // DP rounded to 16: grad GEMM width
  int nslot;        // ring slots (<= MAX_SLOT)
  int slot_floats;  // 16 * max(N1, N3): one K=8 step of the widest GEMM, hi + lo slabs
};

struct TcNet {
  const float *img;  // chunk stream: embed (K1/8 chunks of 2*N1*8 floats), hidden (HK/8), heads (HK/8 of 2*N3*8)
  const float *tb;   // [T][N1]
  const float *b4;   // [N1]
  const float *bh;   // [N3]  (S | T | Q blocks)
  const float *es, *eq;  // [DP]
};

struct TcArgs {
  Shape sh;
  TcDims td;
  TcNet xnet, vnet;
  const float *gimg;  // Gaussian: Ssym chunk stream (KG/8 chunks of 2*NG*8 floats)
  EnergyDev en;
  const float *mask;  // [T][DP]
  TransitionIO io;
};

struct TcLay {
  int xs, vs, gs, smask, h0, su, sdir, sacc, part, ring;
};
__host__ __device__ inline TcLay make_tclay(int DP, int T) {
  TcLay l;
  l.xs = 0;
  l.vs = l.xs + DP * MT;
  l.gs = l.vs + DP * MT;
  l.smask = l.gs + DP * MT;
  l.h0 = l.smask + ((T * DP + 3) & ~3);
  l.su = l.h0 + MT;
  l.sdir = l.su + MT;
  l.sacc = l.sdir + MT;
  l.part = l.sacc + MT;          // [3][2][MT] partial U, K, log|J| of the two column halves
  l.ring = (l.part + 6 * MT + 31) & ~31;  // 128-byte aligned
  return l;
}
__host__ __device__ inline size_t tc_smem_bytes(int DP, int T, int nslot, int slot_floats) {
  return sizeof(float) * ((size_t)make_tclay(DP, T).ring + (size_t)nslot * slot_floats);
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
               "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3]))
               : "memory");
}
__device__ __forceinline__ void compute_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// split 4 values into tf32 hi / lo and store them at column `col` of this thread's TMEM lane
__device__ __forceinline__ void put_a4(uint32_t lane_base, int col, const float (&a)[4]) {
  float hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    hi[j] = tf32_rna(a[j]);
    lo[j] = tf32_rna(a[j] - hi[j]);
  }
  tmem_st4(T_AHI + lane_base + col, hi);
  tmem_st4(T_ALO + lane_base + col, lo);
}

struct Sync {
  uint64_t *full, *empty, *a_ready, *acc_ready;
};

// ---- the GEMM schedule, walked identically by the producer, the MMA issuer and (structurally) the compute warps
// kind: 0 = grad (Gaussian), 1 = embed, 2 = hidden, 3 = heads ; net: 0 = X, 1 = V
template <class F>
__device__ __forceinline__ void walk_schedule(const TcArgs &A, F &&f) {
  const bool gauss = A.en.kind == 0;
  for (int tr = 0; tr < A.io.n_transitions; ++tr) {
    if (gauss) f(0, 0);
    for (int it = 0; it < A.sh.T; ++it) {
      f(1, 1); f(2, 1); f(3, 1);
      f(1, 0); f(2, 0); f(3, 0);
      f(1, 0); f(2, 0); f(3, 0);
      if (gauss) f(0, 0);
      f(1, 1); f(2, 1); f(3, 1);
    }
  }
}

struct GemmDesc {
  const float *src;  // first chunk in global memory
  int nsteps, n, chunk_floats;
};
__device__ __forceinline__ GemmDesc gemm_desc(const TcArgs &A, int kind, int net) {
  const TcDims &td = A.td;
  const TcNet &N = net ? A.vnet : A.xnet;
  GemmDesc g;
  if (kind == 0) {
    g.src = A.gimg; g.nsteps = td.KG / 8; g.n = td.NG;
  } else if (kind == 1) {
    g.src = N.img; g.nsteps = td.K1 / 8; g.n = td.N1;
  } else if (kind == 2) {
    g.src = N.img + (size_t)(td.K1 / 8) * 16 * td.N1; g.nsteps = td.HK / 8; g.n = td.N1;
  } else {
    g.src = N.img + (size_t)(td.K1 / 8 + td.HK / 8) * 16 * td.N1; g.nsteps = td.HK / 8; g.n = td.N3;
  }
  g.chunk_floats = 16 * g.n;  // hi slab (8 k x n) + lo slab
  return g;
}

__global__ void __launch_bounds__(NTHREADS, 1) tc_transition_kernel(const __grid_constant__ TcArgs A) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bars[2 * MAX_SLOT + 2];
  __shared__ uint32_t tmem_slot;
  const Shape &sh = A.sh;
  const TcDims &td = A.td;
  const TransitionIO &io = A.io;
  const TcLay L = make_tclay(sh.DP, sh.T);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = sh.D, DP = sh.DP;
  const long long base = (long long)blockIdx.x * MT;
  Sync S{bars, bars + MAX_SLOT, bars + 2 * MAX_SLOT, bars + 2 * MAX_SLOT + 1};
  float *ring = smem + L.ring;
  const uint32_t NSLOT = (uint32_t)td.nslot, SLOT_FLOATS = (uint32_t)td.slot_floats;

  if (tid == 0) {
    for (int s = 0; s < MAX_SLOT; ++s) {
      mbar_init(&S.full[s], 1);
      mbar_init(&S.empty[s], 1);
    }
    mbar_init(S.a_ready, NCT);
    mbar_init(S.acc_ready, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(&tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 9) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t n = 0;
      walk_schedule(A, [&](int kind, int net) {
        const GemmDesc g = gemm_desc(A, kind, net);
        const uint32_t bytes = (uint32_t)g.chunk_floats * 4u;
        for (int ks = 0; ks < g.nsteps; ++ks, ++n) {
          const uint32_t s = n % NSLOT;
          mbar_wait(&S.empty[s], ((n / NSLOT) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&S.full[s], bytes);
          bulk_g2s(ring + (size_t)s * SLOT_FLOATS, g.src + (size_t)ks * g.chunk_floats, bytes, &S.full[s]);
        }
      });
    }
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t n = 0, gi = 0;
      walk_schedule(A, [&](int kind, int net) {
        const GemmDesc g = gemm_desc(A, kind, net);
        const uint32_t idesc = make_idesc_tf32(128, g.n);
        const uint32_t lbo = (uint32_t)(g.n / 8) * 128u, sbo = 128u;
        mbar_wait(S.a_ready, gi & 1u);
        tcgen05_fence_after();
        for (int ks = 0; ks < g.nsteps; ++ks, ++n) {
          const uint32_t s = n % NSLOT;
          mbar_wait(&S.full[s], (n / NSLOT) & 1u);
          tcgen05_fence_after();
          const uint32_t bhi = smem_u32(ring + (size_t)s * SLOT_FLOATS);
          const uint32_t blo = bhi + (uint32_t)g.n * 32u;  // hi slab = n x 8 floats
          const uint64_t dhi = make_smem_desc(bhi, lbo, sbo), dlo = make_smem_desc(blo, lbo, sbo);
          const uint32_t ahi = tmem + T_AHI + 8u * ks, alo = tmem + T_ALO + 8u * ks;
          mma_tf32_ts(tmem + T_ACC, alo, dhi, idesc, ks > 0);
          mma_tf32_ts(tmem + T_ACC, ahi, dlo, idesc, true);
          mma_tf32_ts(tmem + T_ACC, ahi, dhi, idesc, true);
          tcgen05_commit(&S.empty[s]);
        }
        tcgen05_commit(S.acc_ready);
        ++gi;
      });
    }
  } else {
    // ===================== compute warps =====================
    const int c = 32 * (warp & 3) + lane;  // chain within the tile == TMEM lane
    const int half = warp >> 2;
    const uint32_t lb = tmem + (((uint32_t)(32 * (warp & 3))) << 16);
    const long long gch = base + c;
    const bool gauss = A.en.kind == 0;
    float *xs = smem + L.xs, *vs = smem + L.vs, *gs = smem + L.gs;
    int *sdir = reinterpret_cast<int *>(smem + L.sdir), *sacc = reinterpret_cast<int *>(smem + L.sacc);
    uint32_t gi = 0;  // GEMM counter (parity of a_ready / acc_ready)
    const float eps = sh.eps, Tm = A.en.temperature;
    const int nq = DP / 4;  // 4-dim chunks; this thread owns q with (q & 1) == half

    for (int i = tid; i < sh.T * DP; i += NCT) smem[L.smask + i] = A.mask[i];
    for (int i = tid; i < MT * DP; i += NCT) {
      const int ch = i / DP, d = i - ch * DP;
      const long long g = base + ch;
      xs[d * MT + ch] = (g < io.n && d < D) ? io.x[g * D + d] : 0.f;
    }
    compute_bar();

    for (int tr = 0; tr < io.n_transitions; ++tr) {
      const unsigned long long ctr = io.counter + (unsigned long long)tr;
      // ---- setup: x0, momentum, direction, uniform --------------------------------------------------
      if (io.v != nullptr) {
        for (int i = tid; i < MT * DP; i += NCT) {
          const int ch = i / DP, d = i - ch * DP;
          const long long g = base + ch;
          vs[d * MT + ch] = (g < io.n && d < D) ? io.v[((long long)tr * io.n + g) * D + d] : 0.f;
        }
      } else {
        for (int q = half; q < nq; q += 2) {
          float z[4];
          philox_normals4(io.seed, ctr, io.chain_offset + gch, q, z);
#pragma unroll
          for (int j = 0; j < 4; ++j) vs[(4 * q + j) * MT + c] = (gch < io.n && 4 * q + j < D) ? z[j] : 0.f;
        }
      }
      if (half == 0) {
        int pd = 1;
        float pu = 0.f;
        if (io.dir_mode == 3 || (io.do_mh && io.u == nullptr)) philox_dir_u(io.seed, ctr, io.chain_offset + gch, pd, pu);
        int dbit = 1;
        if (io.dir_mode == 1) dbit = 0;
        else if (io.dir_mode == 2) dbit = (gch < io.n) ? (io.dir[(long long)tr * io.n + gch] != 0) : 1;
        else if (io.dir_mode == 3) dbit = pd;
        sdir[c] = dbit;
        if (io.do_mh && io.u != nullptr) pu = (gch < io.n) ? io.u[(long long)tr * io.n + gch] : 0.f;
        smem[L.su + c] = pu;
      }
      compute_bar();
      const bool fwd = sdir[c] != 0;
      float lj = 0.f;

      // ---- grad U at the current x -> gs (own dims); Gaussian: tensor-core GEMM with Ssym ----------------
      auto grad_phase = [&]() {
        if (gauss) {
          for (int q = half; q < td.KG / 4; q += 2) {
            float a[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int d = 4 * q + j;
              a[j] = d < DP ? xs[d * MT + c] - A.en.mu[d] : 0.f;
            }
            put_a4(lb, 4 * q, a);
          }
          tmem_wait_st();
          tcgen05_fence_before();
          mbar_arrive(S.a_ready);
          mbar_wait(S.acc_ready, gi & 1u);
          ++gi;
          tcgen05_fence_after();
          for (int q = half; q < nq; q += 2) {
            float g4[4];
            tmem_ld4(lb + T_ACC + 4 * q, g4);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 4; ++j) gs[(4 * q + j) * MT + c] = g4[j] / Tm;
          }
          tcgen05_fence_before();
        } else {  // RoughWell (utils/distributions.py:90-97)
          const float e = A.en.s0, den = A.en.s1;
          for (int q = half; q < nq; q += 2)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int d = 4 * q + j;
              const float xi = xs[d * MT + c];
              gs[d * MT + c] = d < D ? (xi - e * sinf(xi / den) / den) / Tm : 0.f;
            }
        }
      };
      // partial Hamiltonian over this thread's dims (needs gs = grad U(x) for the Gaussian kind)
      auto ham_partial = [&](float &U, float &K) {
        U = 0.f; K = 0.f;
        for (int q = half; q < nq; q += 2)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int d = 4 * q + j;
            if (d < D) {
              const float xi = xs[d * MT + c], vi = vs[d * MT + c];
              K = fmaf(vi, vi, K);
              if (gauss) U = fmaf(xi - A.en.mu[d], gs[d * MT + c], U);  // g carries 1/temperature
              else U += (0.5f * xi * xi + A.en.s0 * cosf(xi / A.en.s1)) / Tm;
            }
          }
        if (gauss) U *= 0.5f;
        K *= 0.5f;
      };

      // one S/T/Q net call + fused update.  net: 0 X / 1 V ; mode: 0 momentum, 1 position (xhalf 0/1)
      auto net_call = [&](int net, int it, int mode, int xhalf) {
        const TcNet &N = net ? A.vnet : A.xnet;
        const int tF = it, tB = sh.T - 1 - it;
        // ---- A = [a | b] ----
        for (int q = half; q < nq; q += 2) {
          float a[4], b[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int d = 4 * q + j;
            if (net) {
              a[j] = xs[d * MT + c];
              b[j] = gs[d * MT + c];
            } else {
              const float m = fwd ? smem[L.smask + tF * DP + d] : smem[L.smask + tB * DP + d];
              const float k = (fwd == (xhalf == 0)) ? m : 1.f - m;
              a[j] = vs[d * MT + c];
              b[j] = k * xs[d * MT + c];
            }
          }
          put_a4(lb, 4 * q, a);
          put_a4(lb, DP + 4 * q, b);
        }
        tmem_wait_st();
        tcgen05_fence_before();
        mbar_arrive(S.a_ready);
        // ---- embed epilogue: h1 = relu(acc + tb[t_chain]) -> A ----
        mbar_wait(S.acc_ready, gi & 1u);
        ++gi;
        tcgen05_fence_after();
        const float *tbp = N.tb + (size_t)(fwd ? tF : tB) * td.N1;
        for (int q = half; q < td.HK / 8; q += 2) {
          float h[8];
          tmem_ld8(lb + T_ACC + 8 * q, h);
          tmem_wait_ld();
          float a0[4], a1[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            a0[j] = fmaxf(h[j] + tbp[8 * q + j], 0.f);
            a1[j] = fmaxf(h[4 + j] + tbp[8 * q + 4 + j], 0.f);
          }
          put_a4(lb, 8 * q, a0);
          put_a4(lb, 8 * q + 4, a1);
        }
        tmem_wait_st();
        tcgen05_fence_before();
        mbar_arrive(S.a_ready);
        // ---- hidden epilogue: h2 = relu(acc + b4) -> A ----
        mbar_wait(S.acc_ready, gi & 1u);
        ++gi;
        tcgen05_fence_after();
        for (int q = half; q < td.HK / 8; q += 2) {
          float h[8];
          tmem_ld8(lb + T_ACC + 8 * q, h);
          tmem_wait_ld();
          float a0[4], a1[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            a0[j] = fmaxf(h[j] + N.b4[8 * q + j], 0.f);
            a1[j] = fmaxf(h[4 + j] + N.b4[8 * q + 4 + j], 0.f);
          }
          put_a4(lb, 8 * q, a0);
          put_a4(lb, 8 * q + 4, a1);
        }
        tmem_wait_st();
        tcgen05_fence_before();
        mbar_arrive(S.a_ready);
        // ---- heads epilogue + fused state update (utils/dynamics.py:121-155 / :166-199) ----
        mbar_wait(S.acc_ready, gi & 1u);
        ++gi;
        tcgen05_fence_after();
        for (int q = half; q < nq; q += 2) {
          float s4[4], t4[4], q4[4];
          tmem_ld4(lb + T_ACC + 4 * q, s4);
          tmem_ld4(lb + T_ACC + DP + 4 * q, t4);
          tmem_ld4(lb + T_ACC + 2 * DP + 4 * q, q4);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int d = 4 * q + j;
            const float Sx = N.es[d] * tanhf(s4[j] + N.bh[d]);
            const float Tt = t4[j] + N.bh[DP + d];
            const float Qx = N.eq[d] * tanhf(q4[j] + N.bh[2 * DP + d]);
            if (mode == 0) {
              float v = vs[d * MT + c];
              const float g = gs[d * MT + c];
              const float sv = fwd ? (0.5f * eps) * Sx : (-0.5f * eps) * Sx;
              const float cterm = (0.5f * eps) * (-(expf(eps * Qx) * g) + Tt);
              const float e = expf(sv);
              v = fwd ? (v * e + cterm) : ((v - cterm) * e);
              vs[d * MT + c] = v;
              lj += sv;
            } else {
              const float m = fwd ? smem[L.smask + tF * DP + d] : smem[L.smask + tB * DP + d];
              const float k = (fwd == (xhalf == 0)) ? m : 1.f - m;
              const float uu = 1.f - k;
              float x = xs[d * MT + c];
              const float v = vs[d * MT + c];
              const float sx = fwd ? eps * Sx : -eps * Sx;
              const float inner = eps * (expf(eps * Qx) * v + Tt);
              const float e = expf(sx);
              const float nx = fwd ? (x * e + inner) : (e * (x - inner));
              xs[d * MT + c] = k * x + uu * nx;
              lj += uu * sx;
            }
          }
        }
        tcgen05_fence_before();
      };

      grad_phase();
      {
        float U, K;
        ham_partial(U, K);
        smem[L.part + half * MT + c] = U + K;
      }
      compute_bar();
      if (half == 0) smem[L.h0 + c] = smem[L.part + c] + smem[L.part + MT + c];

      for (int it = 0; it < sh.T; ++it) {
        net_call(1, it, 0, 0);
        net_call(0, it, 1, 0);
        net_call(0, it, 1, 1);
        grad_phase();
        net_call(1, it, 0, 0);
      }

      // ---- log|J|, Hamiltonian, accept ---------------------------------------------------------------
      {
        float U, K;
        ham_partial(U, K);
        smem[L.part + half * MT + c] = U + K;
        smem[L.part + (2 + half) * MT + c] = lj;
      }
      compute_bar();
      const bool last = (tr == io.n_transitions - 1);
      if (half == 0) {
        const float h1 = smem[L.part + c] + smem[L.part + MT + c];
        const float logj = smem[L.part + 2 * MT + c] + smem[L.part + 3 * MT + c];
        const float p = accept_prob(smem[L.h0 + c], h1, logj);
        const float px = io.log_jac ? logj : p;
        int acc = 0;
        if (io.do_mh) acc = (px - smem[L.su + c] >= 0.f) ? 1 : 0;
        sacc[c] = acc;
        if (gch < io.n && last) {
          io.px_out[gch] = px;
          if (io.accepted) io.accepted[gch] = (uint8_t)acc;
        }
      }
      compute_bar();
      if (last) {
        for (int i = tid; i < MT * D; i += NCT) {
          const int ch = i / D, d = i - ch * D;
          const long long g = base + ch;
          if (g < io.n) {
            const float lx = xs[d * MT + ch];
            io.x_out[g * D + d] = lx;
            if (io.v_out) io.v_out[g * D + d] = vs[d * MT + ch];
            // the state this transition started from: the caller's x, or the x_next written one transition ago
            if (io.do_mh) io.x_next[g * D + d] = sacc[ch] ? lx : (tr == 0 ? io.x[g * D + d] : io.x_next[g * D + d]);
          }
        }
      } else {
        // keep x_next in global memory between fused transitions (L2-resident, 2 x 4 D bytes per chain)
        for (int i = tid; i < MT * D; i += NCT) {
          const int ch = i / D, d = i - ch * D;
          const long long g = base + ch;
          if (g < io.n) {
            const float prev = tr == 0 ? io.x[g * D + d] : io.x_next[g * D + d];
            const float nx = sacc[ch] ? xs[d * MT + ch] : prev;
            io.x_next[g * D + d] = nx;
            xs[d * MT + ch] = nx;
          }
        }
        compute_bar();
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace tc
}  // namespace l2hmc
