// Tensor-core transition kernel (sm_100a, tcgen05 / TMEM / TMA bulk copies).
//
// Same transition as kernel_tile.cuh (reference: utils/dynamics.py:115-201,246-309; utils/sampler.py:28-55;
// net SCGExperiment.ipynb:51-77) with the three GEMMs of every net call and the Gaussian grad-U on the 5th-gen
// tensor cores.  fp32 parity is kept with the 3xTF32 split: a = a_hi + a_lo (both tf32),
//   acc += a_lo * b_hi ;  acc += a_hi * b_lo ;  acc += a_hi * b_hi      (fp32 accumulate in TMEM)
// whose dropped term a_lo * b_lo is 2^-22 relative.
//
// One CTA = 128 chains = the 128 TMEM lanes (chain c <-> lane c, MMA M = 128), for the WHOLE transition.
//   warps 0-15 (512 compute threads): thread (c = 32*(w&3) + lane, quarter = w>>2) owns chain c and every
//       fourth 4-dim / 8-column chunk.  It builds the A operand rows directly in TMEM (tcgen05.st), reads the
//       fp32 accumulators back (tcgen05.ld) and does the relu / split / tanh / exp / leapfrog epilogue; x, v,
//       grad U live in shared memory feature-major and every (chain, dim) element is only touched by its owner.
//   warp 16 lane 0: MMA issuer  (tcgen05.mma kind::tf32, A from TMEM, B from the shared-memory ring)
//   warp 17 lane 0: TMA producer (cp.async.bulk global -> shared ring, mbarrier complete_tx)
// B operands (weights, pre-split hi/lo on the host, canonical K-major no-swizzle core-matrix layout) do not fit
// in shared memory for both nets (640 KB for config 2), so they stream from L2 through a ring in consumption
// order; one slot = one K=8 step of one GEMM = {B_hi slab, B_lo slab}.  A bulk copy has ~2.3k cycles of latency
// (profiles/r01_tc_probe.txt), so the ring is as deep as shared memory allows (14 slots for config 2).
//
// TMEM columns: [0,192) accumulator, [192,320) A_hi, [320,448) A_lo.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"
#include <type_traits>

namespace l2hmc {
namespace tc {

constexpr int MT = 128;               // chains per CTA
constexpr int NQMAX = 4;              // compute threads per chain: template parameter NQ in {2, 4}
constexpr int MAX_SLOT = 16;          // ring slots (the host sizes the ring to what shared memory allows)
constexpr int KSLOT = 2;              // K=8 steps per ring slot (halves the mbarrier traffic of the MMA issuer)
constexpr uint32_t T_ACC = 0, T_AHI = 192, T_ALO = 320;

struct TcDims {
  int K1;   // 2*DP            embed GEMM depth  (multiple of 8)
  int HK;   // H rounded to 8  hidden / heads GEMM depth
  int N1;   // HK rounded to 16: embed / hidden GEMM width
  int N3;   // 3*DP rounded to 16: heads GEMM width  (S | T | Q blocks of DP columns)
  int KG;   // DP rounded to 8: grad GEMM depth
  int NG;   // DP rounded to 16: grad GEMM width
  int nslot;        // ring slots (<= MAX_SLOT)
  int slot_floats;  // KSLOT * 16 * max(N1, N3): KSLOT K=8 steps of the widest GEMM, hi + lo slabs each
  int fast_math;    // 1: ex2/rcp based exp and tanh in the epilogue (abs error ~1e-7)
  int nq;           // compute threads per chain (2 or 4) -> which instantiation the host launches
  int f16;          // kernel_tc_s: fp16 split instead of tf32 (all operands inside the fp16 range)
  int biasg;        // kernel_tc_s: biases ride in the GEMMs (weight rows that meet constant-1 / one-hot A columns): 0 no,
                    // 1 one-hot in the two pad dimensions of the last 4-dim chunk, 2 one-hot in a K step of its own (no pad dimensions)
};

struct TcNet {
  const float *img;  // chunk stream: embed (K1/8 chunks of 2*N1*8 floats), hidden (HK/8), heads (HK/8 of 2*N3*8)
  const float *tb;   // [T][N1]
  const float *b4;   // [N1]
  const float *bh;   // [N3]  (S | T | Q blocks)
  const float *es, *eq;  // [DP]
  const float *img_s;    // the same chunk stream with the embed rows interleaved per 4-dim chunk (kernel_tc_s.cuh)
  const float *img_h, *emb_last_h;  // fp16 twins of img_s / emb_last (kernel_tc_s, F16)
  const float *emb_last; // [T] chunks: last embed K step with the time-embedding bias rows of each leapfrog step (biasg)
  const float *hc;       // [DP/4][28]: pre-multiplied heads constants of the specialised kernel (kernel_tc_s.cuh)
};

struct TcArgs {
  Shape sh;
  TcDims td;
  TcNet xnet, vnet;
  const float *gimg;  // Gaussian: Ssym chunk stream (KG/8 chunks of 2*NG*8 floats)
  const float *gimg_h;  // its fp16 twin (K steps of 16)
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
  l.part = l.sacc + MT;                          // [2][NQMAX][MT]: partial Hamiltonian, partial log|J|
  l.ring = (l.part + 2 * NQMAX * MT + 31) & ~31; // 128-byte aligned
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
template <int NCT>
__device__ __forceinline__ void compute_bar_n() { asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory"); }
// waits that last thousands of cycles (compute warps waiting for a GEMM): back off so the spinning warps do not
// take issue slots from the single MMA-issuer / producer threads
#ifndef L2HMC_TC_SLEEP_NS
#define L2HMC_TC_SLEEP_NS 64
#endif
#ifndef L2HMC_TC_PRODUCER_SLEEP_NS
#define L2HMC_TC_PRODUCER_SLEEP_NS 64
#endif
template <int NS = L2HMC_TC_SLEEP_NS>
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t bar_u32 = smem_u32(bar);
  while (true) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(bar_u32), "r"(parity)
        : "memory");
    if (done) break;
    if (NS > 0) __nanosleep(NS);
  }
}

// split 4 values into tf32 hi / lo and store them at column `col` of this thread's TMEM lane
__device__ __forceinline__ void put_a4(uint32_t lane_base, int col, const float (&a)[4]) {
  float hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    // hi = a with the 13 low mantissa bits cleared (what the tf32 datapath reads anyway), lo = a - hi exactly; the MMA
    // keeps lo's top 19 bits.  2 instructions per element instead of 7 (cvt.rna.tf32 is FSETP + IADD + LOP3 in SASS);
    // the dropped terms stay <= 2^-20 |a b| (the weights keep round-to-nearest hi / lo from the host).
    hi[j] = __uint_as_float(__float_as_uint(a[j]) & 0xFFFFE000u);
    lo[j] = a[j] - hi[j];
  }
  tmem_st4(T_AHI + lane_base + col, hi);
  tmem_st4(T_ALO + lane_base + col, lo);
}

// exp / tanh of the epilogue.  FAST: ex2.approx + rcp.approx (abs error ~1e-7 for the O(1) arguments here).
__device__ __forceinline__ float ex2_approx(float x) {  // one MUFU.EX2 (exp2f adds a range check and two scalings)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ep_exp(float x, bool fast) {
  return fast ? ex2_approx(x * 1.4426950408889634f) : expf(x);
}
__device__ __forceinline__ float ep_tanh(float x, bool fast) {
  if (!fast) return tanhf(x);
  const float t = ex2_approx(x * 2.8853900817779268f);  // e^{2x}
  return 1.f - __fdividef(2.f, t + 1.f);
}

struct Sync {
  uint64_t *full, *empty, *a_ready, *acc_ready;
};

// Phase accounting of CTA 0 (clock64 cycles), read back through l2hmc_debug_counters():
// [0] issuer: waiting for A operands (tensor pipe idle, compute warps busy)   [1] issuer: waiting for TMA data
// [2] issuer: total   [3] compute thread 0: waiting for accumulators   [4] compute thread 0: total   [5] GEMMs issued
__device__ long long g_tc_dbg[56];  // [8..12] wait-for-A per GEMM kind, [13..17] wait-for-TMA per kind, [18..22] GEMMs per kind (kernel_tc_s), [23] fp16 range flag

// ---- the GEMM schedule, walked identically by the producer, the MMA issuer and (structurally) the compute warps
// kind: 0 = grad (Gaussian), 1 = embed, 2 = hidden, 3 = heads ; net: 0 = X, 1 = V
template <class F>
__device__ __forceinline__ void walk_schedule(const TcArgs &A, F &&f) {
  const bool gauss = A.en.kind == 0;
  for (int tr = 0; tr < A.io.n_transitions; ++tr) {
    if (gauss) f(0, 0);
    for (int it = 0; it < A.sh.T; ++it) {
      // 13 slots per leapfrog step: V (embed, hidden, heads), X, X, grad (Gaussian only), V.  One rolled loop so that
      // the body of f exists once: the issuer / producer code shares the instruction cache with the compute warps.
#pragma unroll 1
      for (int j = 0; j < 13; ++j) {
        if (j == 9) {
          if (gauss) f(0, 0);
          continue;
        }
        const int jj = j > 9 ? j - 1 : j;  // 0..11
        f(jj % 3 + 1, (jj < 3 || jj >= 9) ? 1 : 0);
      }
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

// ===================== TMA producer (one warp) =====================
// one ring slot = up to KSLOT consecutive K=8 steps of one GEMM (contiguous in the weight stream).
// The whole warp walks the schedule with warp-uniform values (kernel parameters and loop counters only), so
// ptxas keeps the addresses in uniform registers; elect.sync picks the lane that issues the copy.
__device__ __forceinline__ void producer_loop(const TcArgs &A, const Sync &S, float *ring, uint32_t NSLOT, uint32_t SLOT_FLOATS) {
  uint32_t s = 0, ph = 1;  // slot and the parity of the `empty` phase to wait for
  walk_schedule(A, [&](int kind, int net) {
    const GemmDesc g = gemm_desc(A, kind, net);
#pragma unroll 1
    for (int ks = 0; ks < g.nsteps; ks += KSLOT) {
      const uint32_t bytes = (uint32_t)g.chunk_floats * 4u * (uint32_t)min(KSLOT, g.nsteps - ks);
      mbar_wait_sleep(&S.empty[s], ph);  // the producer waits 97 % of the time: do not spin on issue slots
      if (elect_one()) {
        mbar_arrive_expect_tx(&S.full[s], bytes);
        bulk_g2s(ring + (size_t)s * SLOT_FLOATS, g.src + (size_t)ks * g.chunk_floats, bytes, &S.full[s]);
      }
      __syncwarp();
      if (++s == NSLOT) { s = 0; ph ^= 1u; }
    }
  });
}

// ===================== MMA issuer (one warp) =====================
// The whole warp walks the schedule; every operand of tcgen05.mma is computed from kernel parameters and loop
// counters (never from the thread index or a shared-memory load), so the descriptors stay in uniform registers
// and the MMAs of a slot issue back to back (SASS: UTCHMMA x6, UTCBAR) -- with a `lane == 0` leader ptxas
// wrapped every MMA in an ELECT loop fed by R2UR moves and the issue, not the tensor pipe, set the pace
// (113 cycles per 128x112x8 MMA against 56 of tensor work, profiles/r02_tc_issue.txt).
// The TMEM base of a 512-column allocation is column 0 / lane 0 (checked below), so it is a constant here.
__device__ __forceinline__ void issuer_loop(const TcArgs &A, const Sync &S, float *ring, uint32_t NSLOT, uint32_t SLOT_FLOATS, int lane) {
  {
    uint32_t s = 0, ph = 0, gi = 0;
    const uint32_t ring_u32 = smem_u32(ring);
    const uint32_t slot_bytes = SLOT_FLOATS * 4u;
#ifdef L2HMC_TC_PHASE_ACCOUNTING
    long long w_a = 0, w_f = 0;
    const long long t_begin = clock64();
#endif
    walk_schedule(A, [&](int kind, int net) {
      const GemmDesc g = gemm_desc(A, kind, net);
      const uint32_t idesc = make_idesc_tf32(128, g.n);
      // descriptor of a slab at shared address 0; the start-address field (bits 0-13, 16-byte units) is added per slab
      const uint64_t desc0 = make_smem_desc(0u, (uint32_t)(g.n / 8) * 128u, 128u);
      const uint32_t slab16 = (uint32_t)g.n * 2u;  // one slab (n x 8 floats) in 16-byte units
#ifdef L2HMC_TC_PHASE_ACCOUNTING
      long long t0 = clock64();
#endif
      mbar_wait(S.a_ready, gi & 1u);
#ifdef L2HMC_TC_PHASE_ACCOUNTING
      w_a += clock64() - t0;
#endif
      tcgen05_fence_after();
#pragma unroll 1
      for (int ks = 0; ks < g.nsteps; ks += KSLOT) {
#ifdef L2HMC_TC_PHASE_ACCOUNTING
        t0 = clock64();
#endif
        mbar_wait(&S.full[s], ph);
#ifdef L2HMC_TC_PHASE_ACCOUNTING
        w_f += clock64() - t0;
#endif
        const uint32_t b16 = (ring_u32 + s * slot_bytes) >> 4;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < KSLOT; ++kk) {
            if (ks + kk < g.nsteps) {
              const uint64_t dhi = desc0 + (uint64_t)(b16 + (2u * kk) * slab16);
              const uint64_t dlo = desc0 + (uint64_t)(b16 + (2u * kk + 1u) * slab16);
              const uint32_t ahi = T_AHI + 8u * (ks + kk), alo = T_ALO + 8u * (ks + kk);
              mma_tf32_ts(T_ACC, alo, dhi, idesc, (ks + kk) > 0);
              mma_tf32_ts(T_ACC, ahi, dlo, idesc, true);
              mma_tf32_ts(T_ACC, ahi, dhi, idesc, true);
            }
          }
          tcgen05_commit(&S.empty[s]);
        }
        __syncwarp();
        if (++s == NSLOT) { s = 0; ph ^= 1u; }
      }
      if (elect_one()) tcgen05_commit(S.acc_ready);
      __syncwarp();
      ++gi;
    });
#ifdef L2HMC_TC_PHASE_ACCOUNTING
    if (blockIdx.x == 0 && lane == 0) {
      g_tc_dbg[0] = w_a;
      g_tc_dbg[1] = w_f;
      g_tc_dbg[2] = clock64() - t_begin;
      g_tc_dbg[5] = gi;
    }
#endif
  }
}

template <int NQ, bool FAST>
__global__ void __launch_bounds__(MT * NQ + 64, 1) tc_transition_kernel(const __grid_constant__ TcArgs A) {
  constexpr int NCT = MT * NQ;                       // compute threads
  constexpr int W_MMA = NCT / 32, W_TMA = W_MMA + 1; // + MMA-issuer warp + producer warp
  auto compute_bar = []() { compute_bar_n<NCT>(); };
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bars[2 * MAX_SLOT + 2];
  __shared__ uint32_t tmem_slot;
  const Shape &sh = A.sh;
  const TcDims &td = A.td;
  const TransitionIO &io = A.io;
  const TcLay L = make_tclay(sh.DP, sh.T);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler (role dispatch stays on the uniform datapath)
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
  if (warp == W_MMA) tmem_alloc(&tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tmem != 0u) __trap();  // this CTA owns all 512 columns: the issuer uses column / lane 0 as a constant

  if (warp == W_TMA) {
    producer_loop(A, S, ring, NSLOT, SLOT_FLOATS);
  } else if (warp == W_MMA) {
    issuer_loop(A, S, ring, NSLOT, SLOT_FLOATS, lane);
  } else {
    // ===================== compute warps =====================
    const int c = 32 * (warp & 3) + lane;  // chain within the tile == TMEM lane
    const int qd = warp >> 2;              // which 1/NQ of the chunks this thread owns
    const uint32_t lb = tmem + (((uint32_t)(32 * (warp & 3))) << 16);
    const long long gch = base + c;
    const bool gauss = A.en.kind == 0;
    constexpr bool fast = FAST;
    float *xs = smem + L.xs, *vs = smem + L.vs, *gs = smem + L.gs;
    int *sdir = reinterpret_cast<int *>(smem + L.sdir), *sacc = reinterpret_cast<int *>(smem + L.sacc);
    uint32_t gi = 0;  // GEMM counter (parity of a_ready / acc_ready)
#ifdef L2HMC_TC_PHASE_ACCOUNTING
    long long w_acc = 0;
    const long long t_begin = clock64();
#endif
    auto wait_acc = [&]() {
#ifdef L2HMC_TC_PHASE_ACCOUNTING
      const long long t0 = clock64();
      mbar_wait_sleep(S.acc_ready, gi & 1u);
      w_acc += clock64() - t0;
#else
      mbar_wait_sleep(S.acc_ready, gi & 1u);
#endif
      ++gi;
    };
    const float eps = sh.eps, Tm = A.en.temperature;
    const int nq = DP / 4;      // 4-dim chunks; this thread owns q with (q & 3) == qd
    const int nh = td.HK / 8;   // 8-column chunks of the hidden layers

    for (int i = tid; i < sh.T * DP; i += NCT) smem[L.smask + i] = A.mask[i];
    for (int i = tid; i < MT * DP; i += NCT) {
      const int ch = i / DP, d = i - ch * DP;
      const long long g = base + ch;
      xs[d * MT + ch] = (g < io.n && d < D) ? io.x[g * D + d] : 0.f;
    }
    compute_bar();

    for (int tr = 0; tr < io.n_transitions; ++tr) {
      const unsigned long long ctr = io.counter + (unsigned long long)tr;
      // ---- setup: momentum, direction, uniform -------------------------------------------------------
      if (io.v != nullptr) {
        for (int i = tid; i < MT * DP; i += NCT) {
          const int ch = i / DP, d = i - ch * DP;
          const long long g = base + ch;
          vs[d * MT + ch] = (g < io.n && d < D) ? io.v[((long long)tr * io.n + g) * D + d] : 0.f;
        }
      } else {
        for (int q = qd; q < nq; q += NQ) {
          float z[4];
          philox_normals4(io.seed, ctr, io.chain_offset + gch, q, z);
#pragma unroll
          for (int j = 0; j < 4; ++j) vs[(4 * q + j) * MT + c] = (gch < io.n && 4 * q + j < D) ? z[j] : 0.f;
        }
      }
      if (qd == 0) {
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
          for (int q = qd; q < td.KG / 4; q += NQ) {
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
          wait_acc();
          tcgen05_fence_after();
#pragma unroll 1
          for (int q = qd; q < nq; q += NQ) {
            float g4[4];
            tmem_ld4(lb + T_ACC + 4 * q, g4);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 4; ++j) gs[(4 * q + j) * MT + c] = g4[j] / Tm;
          }
          tcgen05_fence_before();
        } else {  // RoughWell (utils/distributions.py:90-97)
          const float e = A.en.s0, den = A.en.s1;
          for (int q = qd; q < nq; q += NQ)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int d = 4 * q + j;
              const float xi = xs[d * MT + c];
              gs[d * MT + c] = d < D ? (xi - e * sinf(xi / den) / den) / Tm : 0.f;
            }
        }
      };
      // partial Hamiltonian over this thread's dims (needs gs = grad U(x) for the Gaussian kind)
      auto ham_partial = [&]() -> float {
        float U = 0.f, K = 0.f;
        for (int q = qd; q < nq; q += NQ)
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
        return U + 0.5f * K;
      };
      // relu(acc + bias) of this thread's 8-column chunks -> next A operand (bias may differ per direction)
      auto hidden_epilogue = [&](const float *bias) {
        wait_acc();
        tcgen05_fence_after();
        // software pipeline: the load of the next chunk is in flight while this one is processed
        float h[8], hn[8];
        if (qd < nh) tmem_ld8(lb + T_ACC + 8 * qd, h);
#pragma unroll 1
        for (int q = qd; q < nh; q += NQ) {
          tmem_wait_ld();
          if (q + NQ < nh) tmem_ld8(lb + T_ACC + 8 * (q + NQ), hn);
          const float4 b0 = __ldg(reinterpret_cast<const float4 *>(bias + 8 * q));
          const float4 b1 = __ldg(reinterpret_cast<const float4 *>(bias + 8 * q + 4));
          const float a0[4] = {fmaxf(h[0] + b0.x, 0.f), fmaxf(h[1] + b0.y, 0.f), fmaxf(h[2] + b0.z, 0.f), fmaxf(h[3] + b0.w, 0.f)};
          const float a1[4] = {fmaxf(h[4] + b1.x, 0.f), fmaxf(h[5] + b1.y, 0.f), fmaxf(h[6] + b1.z, 0.f), fmaxf(h[7] + b1.w, 0.f)};
          put_a4(lb, 8 * q, a0);
          put_a4(lb, 8 * q + 4, a1);
#pragma unroll
          for (int j = 0; j < 8; ++j) h[j] = hn[j];
        }
        tmem_wait_st();
        tcgen05_fence_before();
        mbar_arrive(S.a_ready);
      };

      // one S/T/Q net call + fused update.  net: 0 X / 1 V ; mode: 0 momentum, 1 position (xhalf 0/1)
      auto net_call = [&](int net, int it, int mode, int xhalf) {
        const TcNet &N = net ? A.vnet : A.xnet;
        const int tF = it, tB = sh.T - 1 - it;
        const float *mrow = smem + L.smask + (fwd ? tF : tB) * DP;
        // ---- A = [a | b] ----
        for (int q = qd; q < nq; q += NQ) {
          float a[4], b[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int d = 4 * q + j;
            if (net) {
              a[j] = xs[d * MT + c];
              b[j] = gs[d * MT + c];
            } else {
              const float m = mrow[d];
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
        hidden_epilogue(N.tb + (size_t)(fwd ? tF : tB) * td.N1);  // h1 = relu(acc + tb[t_chain])
        hidden_epilogue(N.b4);                                    // h2 = relu(acc + b4)
        // ---- heads epilogue + fused state update (utils/dynamics.py:121-155 / :166-199) ----
        wait_acc();
        tcgen05_fence_after();
        float s4[4], t4[4], q4[4], sn[4], tn[4], qn[4];
        if (qd < nq) {
          tmem_ld4(lb + T_ACC + 4 * qd, s4);
          tmem_ld4(lb + T_ACC + DP + 4 * qd, t4);
          tmem_ld4(lb + T_ACC + 2 * DP + 4 * qd, q4);
        }
#pragma unroll 1
        for (int q = qd; q < nq; q += NQ) {
          tmem_wait_ld();
          if (q + NQ < nq) {  // next chunk's accumulators travel while this chunk is processed
            tmem_ld4(lb + T_ACC + 4 * (q + NQ), sn);
            tmem_ld4(lb + T_ACC + DP + 4 * (q + NQ), tn);
            tmem_ld4(lb + T_ACC + 2 * DP + 4 * (q + NQ), qn);
          }
          {
            // per-dimension constants of this 4-dim chunk: one 16-byte load each (warp-uniform addresses)
            const float4 c_es = __ldg(reinterpret_cast<const float4 *>(N.es + 4 * q));
            const float4 c_eq = __ldg(reinterpret_cast<const float4 *>(N.eq + 4 * q));
            const float4 c_bs = __ldg(reinterpret_cast<const float4 *>(N.bh + 4 * q));
            const float4 c_bt = __ldg(reinterpret_cast<const float4 *>(N.bh + DP + 4 * q));
            const float4 c_bq = __ldg(reinterpret_cast<const float4 *>(N.bh + 2 * DP + 4 * q));
            const float es4[4] = {c_es.x, c_es.y, c_es.z, c_es.w}, eq4[4] = {c_eq.x, c_eq.y, c_eq.z, c_eq.w};
            const float bs4[4] = {c_bs.x, c_bs.y, c_bs.z, c_bs.w}, bt4[4] = {c_bt.x, c_bt.y, c_bt.z, c_bt.w};
            const float bq4[4] = {c_bq.x, c_bq.y, c_bq.z, c_bq.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int d = 4 * q + j;
              const float Sx = es4[j] * ep_tanh(s4[j] + bs4[j], fast);
              const float Tt = t4[j] + bt4[j];
              const float Qx = eq4[j] * ep_tanh(q4[j] + bq4[j], fast);
              if (mode == 0) {
                float v = vs[d * MT + c];
                const float g = gs[d * MT + c];
                const float sv = fwd ? (0.5f * eps) * Sx : (-0.5f * eps) * Sx;
                const float cterm = (0.5f * eps) * (-(ep_exp(eps * Qx, fast) * g) + Tt);
                const float e = ep_exp(sv, fast);
                v = fwd ? (v * e + cterm) : ((v - cterm) * e);
                vs[d * MT + c] = v;
                lj += sv;
              } else {
                const float m = mrow[d];
                const float k = (fwd == (xhalf == 0)) ? m : 1.f - m;
                const float uu = 1.f - k;
                float x = xs[d * MT + c];
                const float v = vs[d * MT + c];
                const float sx = fwd ? eps * Sx : -eps * Sx;
                const float inner = eps * (ep_exp(eps * Qx, fast) * v + Tt);
                const float e = ep_exp(sx, fast);
                const float nx = fwd ? (x * e + inner) : (e * (x - inner));
                xs[d * MT + c] = k * x + uu * nx;
                lj += uu * sx;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            s4[j] = sn[j];
            t4[j] = tn[j];
            q4[j] = qn[j];
          }
        }
        tcgen05_fence_before();
      };

      grad_phase();
      smem[L.part + qd * MT + c] = ham_partial();
      compute_bar();
      if (qd == 0) {
        float h = 0.f;
#pragma unroll
        for (int r = 0; r < NQ; ++r) h += smem[L.part + r * MT + c];
        smem[L.h0 + c] = h;
      }

      for (int it = 0; it < sh.T; ++it) {
        net_call(1, it, 0, 0);
        net_call(0, it, 1, 0);
        net_call(0, it, 1, 1);
        grad_phase();
        net_call(1, it, 0, 0);
      }

      // ---- log|J|, Hamiltonian, accept ---------------------------------------------------------------
      compute_bar();  // h0 readers are done with `part`
      smem[L.part + qd * MT + c] = ham_partial();
      smem[L.part + (NQ + qd) * MT + c] = lj;
      compute_bar();
      const bool last = (tr == io.n_transitions - 1);
      if (qd == 0) {
        float h1 = 0.f, logj = 0.f;
#pragma unroll
        for (int r = 0; r < NQ; ++r) {
          h1 += smem[L.part + r * MT + c];
          logj += smem[L.part + (NQ + r) * MT + c];
        }
        const float p = accept_prob(smem[L.h0 + c], h1, logj);
        const float px = io.log_jac ? logj : p;
        int acc = 0;
        if (io.do_mh) acc = (px - smem[L.su + c] >= 0.f) ? 1 : 0;
        sacc[c] = acc;
        if (gch < io.n) stats_add(io.stats, px, acc);
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
            if (io.do_mh) {
              const float nx = sacc[ch] ? lx : (tr == 0 ? io.x[g * D + d] : io.x_next[g * D + d]);
              io.x_next[g * D + d] = nx;
              if (io.trace) io.trace[((long long)tr * io.n + g) * D + d] = nx;
            }
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
            if (io.trace) io.trace[((long long)tr * io.n + g) * D + d] = nx;
            xs[d * MT + ch] = nx;
          }
        }
        compute_bar();
      }
    }
#ifdef L2HMC_TC_PHASE_ACCOUNTING
    if (blockIdx.x == 0 && tid == 0) {
      g_tc_dbg[3] = w_acc;
      g_tc_dbg[4] = clock64() - t_begin;
    }
#endif
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace tc
}  // namespace l2hmc
