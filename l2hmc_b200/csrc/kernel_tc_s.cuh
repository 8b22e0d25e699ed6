// Tensor-core transition kernel, shape-specialised compute path (sm_100a).
//
// Same launch shape and roles as kernel_tc.cuh (reference: utils/dynamics.py:115-201, 246-309; utils/sampler.py:28-55; net
// SCGExperiment.ipynb:51-77): one CTA = 128 chains for the whole transition, 8 compute warps, one MMA-issuer warp, one
// TMA-producer warp.  What differs is everything the ncu source page and the phase counters pointed at, one measured
// step at a time (profiles/r01_tc_s_history.txt):
//   * chunk counts are template parameters (NQC = DP/4 dimension chunks, NHC = HK/8 hidden chunks); the chunk loops stay
//     rolled, two chunks per iteration for the register double buffer (fully unrolled, the loop body was 178 KB and
//     instruction fetch took 40 % of the epilogue time);
//   * x, v, grad U live in shared memory chain-major (row stride RS floats, conflict-free 128-bit accesses): one
//     LDS.128 per 4 dimensions instead of 4 LDS.32;
//   * the per-dimension constants of the heads epilogue are pre-multiplied on the host (TcNet::hc) and copied to shared
//     memory once per CTA (as global loads their latency was ~18 % of the heads epilogue): tanh / exp of
//     the S and Q heads run in log2 units with one reciprocal for both tanh (5 MUFU per dimension instead of 6 --
//     MUFU is the floor of this epilogue: 16 lanes per clock per SM);
//   * the A operand of the NEXT GEMM (net input [a | b], or x - mu for the Gaussian grad) is produced inside the
//     heads / grad epilogue from values still in registers: no separate pass over the state;
//   * one mbarrier arrival per warp (count 8) instead of one per thread (count 256);
//   * biases ride in the GEMMs where the width leaves a pad hidden unit (BIASG): a constant-1 hidden unit, and a direction
//     one-hot in the net input that selects the time-embedding bias row of the chain's leapfrog step -- in the two pad
//     dimensions of the last chunk where x_dim leaves them (td.biasg 1), else in a K step of its own (td.biasg 2);
//   * fp16 split: the relu of the two hidden layers happens inside the fp16 conversion of the operand split
//     (cvt.rn.relu.f16x2.f32 on hi = a & mask and lo = a - hi, which carry the sign of a): no separate max(a, 0);
//   * the epilogue of GEMM k and the MMAs of GEMM k+1 overlap: three accumulator regions in TMEM, the A operand
//     handed over in slots of 16 k (the chunks i of both threads of a chain) through sub-barriers a_sub[0..NSUB),
//     chunks owned round-robin by the two threads of a chain so that they complete in K order, and the net input
//     interleaved per chunk ([a0..3 | b0..3] = one K step; the embed weight image is permuted to match on the host);
//   * the heads GEMM is split by dimensions into two GEMMs (first ceil(NQC/2) chunks | the rest, each with its S | T | Q
//     column blocks): two heads-sized accumulators do not fit beside the A operand in 512 TMEM columns, the halves do
//     (embed / heads_a: R1, hidden / grad: R2, heads_b: R3), and the state update of the first half runs while the
//     tensor pipe works on the second.  The A operand of the next GEMM is written during the second half's
//     epilogue only (for all chunks), so nothing overwrites A while heads_b still reads it.
//   * F16 = true: the same split with fp16 pairs instead of tf32 (a = a_hi + a_lo, b = b_hi + b_lo as fp16; three
//     kind::f16 MMAs per product, each covering K = 16 in the cycles a tf32 MMA needs for K = 8: half the tensor time;
//     measured error 5.9e-7 relative, profiles/r01_f16_probe.txt).  Two fp16 share a 32-bit TMEM column (low half = lower
//     k), the B images hold fp16.  fp16 stops at 65504: the host keeps tf32 when a weight is out of range, the kernel
//     records operand magnitudes and raises a sticky flag (l2hmc_debug_counters[23]) when an activation was.
// Shapes without an instantiation run the generic kernel of kernel_tc.cuh.
#pragma once
#include <cuda_fp16.h>
#include "kernel_tc.cuh"

namespace l2hmc {
namespace tc {

// TMEM columns of this kernel: A_hi [0,104), A_lo [104,208), accumulators R1 [208,320), R2 [320,432), R3 [432,512)
constexpr uint32_t S_AHI = 0, S_ALO = 104, S_R1 = 208, S_R2 = 320, S_R3 = 432;
#ifndef L2HMC_TC_KSLOT_S
#define L2HMC_TC_KSLOT_S 4
#endif
// K steps per ring slot (= per bulk copy) of this kernel.  A bulk copy takes ~590 cycles whatever its size up to 32 KB
// (profiles/r01_tc_probe.txt: 14 / 28 / 56 B/cycle for 8 / 16 / 32 KB), so the B stream (33 B/cycle per SM once every
// GEMM overlaps an epilogue) wants few large copies; the A operand is still handed over per slot of 16 k.
constexpr int KSLOT_S = L2HMC_TC_KSLOT_S;
#ifndef L2HMC_TC_F32X2
#define L2HMC_TC_F32X2 1  // packed fp32 arithmetic (FFMA2 / FMUL2 / FADD2) in the epilogues
#endif
#ifndef L2HMC_TC_SETMAXNREG
#define L2HMC_TC_SETMAXNREG (L2HMC_TC_S_NQ > 2)  // the last warpgroup (MMA issuer, TMA producer, two idle warps) hands registers to the compute warpgroups
#endif
#ifndef L2HMC_TC_RELU_PACK
#define L2HMC_TC_RELU_PACK 1  // fp16 split: the relu of the hidden layers happens in the fp16 conversion (cvt.rn.relu.f16x2.f32), not before it

#endif
static_assert(KSLOT_S % 2 == 0, "ring slot = whole A hand-over slots");
constexpr int NSUB_MAX = 8;  // sub-barriers of the A operand (one per K slot of 16 columns)
constexpr int HC_PER_CHUNK = 28;  // floats per 4-dim chunk of TcNet::hc: bs2, bq2, n2cS, cS, n2cQ, cQ, bth (4 each)


constexpr int HCS_PER_CHUNK = 20;  // shared-memory copy of the head constants: bs2, bq2, cS, cQ, bth (4 floats each)
struct TcLayS {
  int xs, vs, gs, smask, h0, su, sdir, sacc, part, hcs, ring;
};
__host__ __device__ inline TcLayS make_tclay_s(int RS, int DP, int T) {
  TcLayS l;
  l.xs = 0;
  l.vs = l.xs + RS * MT;
  l.gs = l.vs + RS * MT;
  l.smask = l.gs + RS * MT;
  l.h0 = l.smask + ((T * DP + 3) & ~3);
  l.su = l.h0 + MT;
  l.sdir = l.su + MT;
  l.sacc = l.sdir + MT;
  l.part = l.sacc + MT;                      // [4][MT]: partial Hamiltonian, then partial log|J| (one after the other), per thread group
  l.hcs = l.part + 4 * MT;                    // [2 nets][DP/4][HCS_PER_CHUNK]
  l.ring = (l.hcs + 2 * (DP / 4) * HCS_PER_CHUNK + 31) & ~31;  // 128-byte aligned
  return l;
}
__host__ __device__ inline int tc_s_row_stride(int DP) { return ((DP / 4) & 1) ? DP : DP + 4; }
// shapes the TMEM map of this kernel holds (A_hi / A_lo 104 columns each, accumulators 112 + 112 + 80)
__host__ __device__ inline bool tc_s_shape_fits(int nqc, int nhc) {
  const int ca = (nqc + 1) / 2, cb = nqc - ca;
  return nqc >= 2 && nhc >= 2 && 8 * nqc <= 104 && 8 * nhc <= 104 && 12 * ca <= 112 && 12 * cb <= 80 && (8 * nhc + 15) / 16 * 16 <= 112 &&
         ((nqc > nhc ? nqc : nhc) + 1) / 2 <= NSUB_MAX;
}
__host__ __device__ inline size_t tc_s_smem_bytes(int DP, int T, int nslot, int slot_floats) {
  return sizeof(float) * ((size_t)make_tclay_s(tc_s_row_stride(DP), DP, T).ring + (size_t)nslot * slot_floats);
}

__device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void sts4(float *p, const float (&a)[4]) {
  *reinterpret_cast<float4 *>(p) = make_float4(a[0], a[1], a[2], a[3]);
}
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_st4u(uint32_t taddr, const uint32_t (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
__device__ __forceinline__ void tmem_st2u(uint32_t taddr, const uint32_t (&v)[2]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(v[0]), "r"(v[1]) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float k_even, float k_odd) {  // low half = lower k (profiles/r01_f16_probe.txt)
  const __half2 h = __floats2half2_rn(k_even, k_odd);
  return *reinterpret_cast<const uint32_t *>(&h);
}
// the same with negative values clamped to +0 by the conversion itself (one F2FP.RELU instead of two FMNMX + F2FP)
__device__ __forceinline__ uint32_t pack_h2_relu(float k_even, float k_odd) {
  uint32_t r;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(k_odd), "f"(k_even));
  return r;
}
// tcgen05.mma kind::f16 (fp16 inputs, fp32 accumulate): D[tmem] (+)= A[tmem] * B[smem]^T, one K = 16 slice
__host__ __device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);  // A / B format 0 = f16, D format 1 = f32
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}

// split N values a[k0 .. k0+N) of the A operand into hi / lo and store them in this warp's TMEM lanes.
//   tf32: hi = a with the 13 low mantissa bits cleared, lo = a - hi (see put_a4); one 32-bit column per k.
//   fp16: the same hi (11 significant bits: exactly an fp16 inside its normal range) and lo, rounded to fp16 and packed
//         two per column (column k / 2); amax tracks |a| for the range check (fp16 overflows at 65504).
// RELU (fp16 only): the operand is relu(a).  hi = a & mask and lo = a - hi have the sign of a (|hi| <= |a|), so clamping both
// halves at the conversion gives exactly the split of relu(a) -- no separate max(a, 0); amax then also sees the negative
// pre-activations, which only makes the range check more conservative.
template <bool F16, int N, bool RELU = false>
__device__ __forceinline__ void put_a(uint32_t lane_base, int k0, const float (&a)[N], float &amax) {
  static_assert(N == 4 || N == 8, "4 or 8 values");
  static_assert(!RELU || F16, "relu in the conversion: fp16 split only");
  float hi[N], lo[N];
#pragma unroll
  for (int j = 0; j < N; j += 2) {
    hi[j] = __uint_as_float(__float_as_uint(a[j]) & 0xFFFFE000u);
    hi[j + 1] = __uint_as_float(__float_as_uint(a[j + 1]) & 0xFFFFE000u);
#if L2HMC_TC_F32X2
    const float2 l2 = __fadd2_rn(make_float2(a[j], a[j + 1]), make_float2(-hi[j], -hi[j + 1]));  // exact: one FADD2 for two residuals
    lo[j] = l2.x;
    lo[j + 1] = l2.y;
#else
    lo[j] = a[j] - hi[j];
    lo[j + 1] = a[j + 1] - hi[j + 1];
#endif
  }
  if (F16) {
    uint32_t h2[N / 2], l2[N / 2];
#pragma unroll
    for (int j = 0; j < N / 2; ++j) {
      h2[j] = RELU ? pack_h2_relu(hi[2 * j], hi[2 * j + 1]) : pack_h2(hi[2 * j], hi[2 * j + 1]);
      l2[j] = RELU ? pack_h2_relu(lo[2 * j], lo[2 * j + 1]) : pack_h2(lo[2 * j], lo[2 * j + 1]);
      amax = fmaxf(amax, fmaxf(fabsf(a[2 * j]), fabsf(a[2 * j + 1])));
    }
    if (N == 8) {
      const uint32_t h4[4] = {h2[0], h2[1], h2[2], h2[3]}, l4[4] = {l2[0], l2[1], l2[2], l2[3]};
      tmem_st4u(S_AHI + lane_base + k0 / 2, h4);
      tmem_st4u(S_ALO + lane_base + k0 / 2, l4);
    } else {
      const uint32_t hh[2] = {h2[0], h2[1]}, ll[2] = {l2[0], l2[1]};
      tmem_st2u(S_AHI + lane_base + k0 / 2, hh);
      tmem_st2u(S_ALO + lane_base + k0 / 2, ll);
    }
  } else {
    if (N == 8) {
      const float h8[8] = {hi[0], hi[1], hi[2], hi[3], hi[4], hi[5], hi[6], hi[7]}, l8[8] = {lo[0], lo[1], lo[2], lo[3], lo[4], lo[5], lo[6], lo[7]};
      tmem_st8(S_AHI + lane_base + k0, h8);
      tmem_st8(S_ALO + lane_base + k0, l8);
    } else {
      const float h4[4] = {hi[0], hi[1], hi[2], hi[3]}, l4[4] = {lo[0], lo[1], lo[2], lo[3]};
      tmem_st4(S_AHI + lane_base + k0, h4);
      tmem_st4(S_ALO + lane_base + k0, l4);
    }
  }
}

// GEMM kinds of this kernel: 0 grad (Gaussian), 1 embed, 2 hidden, 3 heads_a (first CA dimension chunks), 4 heads_b.
// Where each accumulates and what it needs before its first MMA:
//   embed -> R1, hidden -> R2, heads_a -> R1, heads_b -> R3, grad -> R2.  Consecutive GEMMs never share a region, the
//   region a GEMM writes was last read by an epilogue that has finished, so every GEMM follows the epilogue before it
//   K slot by K slot (grad: K step by K step, x - mu has 4 columns per chunk); heads_b reads the A operand heads_a has
//   already waited for.
__device__ __forceinline__ uint32_t s_region(int kind) {
  return (kind == 1 || kind == 3) ? S_R1 : (kind == 4 ? S_R3 : S_R2);
}
// 17 slots per leapfrog step: V (embed, hidden, heads_a, heads_b), X, X, grad (Gaussian only), V
template <class F>
__device__ __forceinline__ void walk_schedule_s(const TcArgs &A, F &&f) {
  const bool gauss = A.en.kind == 0;
  for (int tr = 0; tr < A.io.n_transitions; ++tr) {
    if (gauss) f(0, 0, 0);
    for (int it = 0; it < A.sh.T; ++it) {
#pragma unroll 1
      for (int j = 0; j < 17; ++j) {
        if (j == 12) {
          if (gauss) f(0, 0, it);
          continue;
        }
        const int jj = j > 12 ? j - 1 : j;  // 0..15
        f((jj & 3) + 1, (jj < 4 || jj >= 12) ? 1 : 0, it);
      }
    }
  }
}
// chunk stream of a GEMM in the specialised image [embed (interleaved rows) | hidden | heads_a | heads_b]; one chunk =
// one MMA K step (8 k in tf32, 16 k in fp16) = {hi slab, lo slab} = 16 * n 32-bit words either way
template <bool F16>
__device__ __forceinline__ GemmDesc gemm_desc_s(const TcArgs &A, const int NQC, int kind, int net) {
  const int CA = (NQC + 1) / 2, CB = NQC - CA;
  const int N3A = (12 * CA + 15) / 16 * 16, N3B = (12 * CB + 15) / 16 * 16;
  constexpr int KS = F16 ? 16 : 8;
  const TcDims &td = A.td;
  const TcNet &N = net ? A.vnet : A.xnet;
  const float *img = F16 ? N.img_h : N.img_s;
  const int ne = (td.K1 + KS - 1) / KS, nh = (td.HK + KS - 1) / KS, ng = (td.KG + KS - 1) / KS;
  GemmDesc g;
  const size_t o_hid = (size_t)ne * 16 * td.N1, o_ha = o_hid + (size_t)nh * 16 * td.N1;
  if (kind == 0) {
    g.src = F16 ? A.gimg_h : A.gimg; g.nsteps = ng; g.n = td.NG;
  } else if (kind == 1) {
    g.src = img; g.nsteps = ne + (td.biasg == 2 ? 1 : 0); g.n = td.N1;  // biasg 2: one more K step that holds the bias rows only
  } else if (kind == 2) {
    g.src = img + o_hid; g.nsteps = nh; g.n = td.N1;
  } else if (kind == 3) {
    g.src = img + o_ha; g.nsteps = nh; g.n = N3A;
  } else {
    g.src = img + o_ha + (size_t)nh * 16 * N3A; g.nsteps = nh; g.n = N3B;
  }
  g.chunk_floats = 16 * g.n;
  return g;
}

// ===================== TMA producer of this schedule (one warp) =====================
template <bool F16>
__device__ __forceinline__ void producer_loop_s(const TcArgs &A, const int NQC, const Sync &S, float *ring, uint32_t NSLOT, uint32_t SLOT_FLOATS) {
  uint32_t s = 0, ph = 1;
  walk_schedule_s(A, [&](int kind, int net, int it) {
    const GemmDesc g = gemm_desc_s<F16>(A, NQC, kind, net);
#pragma unroll 1
    for (int ks = 0; ks < g.nsteps; ks += KSLOT_S) {
      const uint32_t bytes = (uint32_t)g.chunk_floats * 4u * (uint32_t)min(KSLOT_S, g.nsteps - ks);
      mbar_wait_sleep<L2HMC_TC_PRODUCER_SLEEP_NS>(&S.empty[s], ph);  // the producer waits 97 % of the time and shares a scheduler with two compute warps
      if (elect_one()) {
        mbar_arrive_expect_tx(&S.full[s], bytes);
        const int nk = min(KSLOT_S, g.nsteps - ks);
        if (A.td.biasg && kind == 1 && ks + nk == g.nsteps) {
          // the last K step of the embed carries the time-embedding bias rows of leapfrog step `it` (forward and
          // backward chains): it comes from the per-step image, the K steps before it from the common stream
          const uint32_t cb = (uint32_t)g.chunk_floats * 4u;
          if (nk > 1) bulk_g2s(ring + (size_t)s * SLOT_FLOATS, g.src + (size_t)ks * g.chunk_floats, cb * (uint32_t)(nk - 1), &S.full[s]);
          bulk_g2s(ring + (size_t)s * SLOT_FLOATS + (size_t)(nk - 1) * g.chunk_floats,
                   (F16 ? (net ? A.vnet : A.xnet).emb_last_h : (net ? A.vnet : A.xnet).emb_last) + (size_t)it * g.chunk_floats, cb,
                   &S.full[s]);
        } else {
          bulk_g2s(ring + (size_t)s * SLOT_FLOATS, g.src + (size_t)ks * g.chunk_floats, bytes, &S.full[s]);
        }
      }
      __syncwarp();
      if (++s == NSLOT) { s = 0; ph ^= 1u; }
    }
  });
}

// ===================== MMA issuer of the overlapped schedule (one warp) =====================
// As issuer_loop (uniform datapath, elect.sync), plus: the accumulator region per GEMM, the A operand awaited per K slot
// (a_sub[K step / 2]; grad: a_sub[K step], its A operand has 4 columns per chunk; heads_b: not at all), and two
// accumulator-ready barriers used alternately (heads_a and heads_b complete without an epilogue in between).
template <bool F16>
__device__ __forceinline__ void issuer_loop_s(const TcArgs &A, const int NQC, const Sync &S, uint64_t *a_sub, uint64_t *acc_rdy, int nsub, float *ring,
                                              uint32_t NSLOT, uint32_t SLOT_FLOATS, int lane) {
  uint32_t s = 0, ph = 0, gi = 0, ai = 0;  // ring slot / phase, GEMM counter, A-operand generation counter
  const uint32_t ring_u32 = smem_u32(ring);
  const uint32_t slot_bytes = SLOT_FLOATS * 4u;
#ifdef L2HMC_TC_PHASE_ACCOUNTING
  long long w_a = 0, w_f = 0, k_a[5] = {0, 0, 0, 0, 0}, k_f[5] = {0, 0, 0, 0, 0}, k_n[5] = {0, 0, 0, 0, 0};
  const long long t_begin = clock64();
#endif
  walk_schedule_s(A, [&](int kind, int net, int) {
#ifdef L2HMC_TC_PHASE_ACCOUNTING
    const long long w_a0 = w_a, w_f0 = w_f;
#endif
    const GemmDesc g = gemm_desc_s<F16>(A, NQC, kind, net);
    const uint32_t acc = s_region(kind);
    const uint32_t idesc = F16 ? make_idesc_f16(128, g.n) : make_idesc_tf32(128, g.n);
    const uint64_t desc0 = make_smem_desc(0u, (uint32_t)(g.n / 8) * 128u, 128u);
    const uint32_t slab16 = (uint32_t)g.n * 2u;
    const uint32_t par = ai & 1u;
    const bool grad = kind == 0, nowait = kind == 4;
#pragma unroll 1
    for (int ks = 0; ks < g.nsteps; ks += KSLOT_S) {
#ifdef L2HMC_TC_PHASE_ACCOUNTING
      long long t0 = clock64();
#endif
      mbar_wait(&S.full[s], ph);
#ifdef L2HMC_TC_PHASE_ACCOUNTING
      w_f += clock64() - t0;
#endif
      const uint32_t b16 = (ring_u32 + s * slot_bytes) >> 4;
#pragma unroll
      for (int kk = 0; kk < KSLOT_S; ++kk) {
        if (ks + kk < g.nsteps) {
          // A hand-over slot i = the chunks i of both threads of a chain = 16 k (grad: 8 k, x - mu has 4 per chunk).
          // tf32 (K step = 8 k): slot = 2 K steps, grad 1.  fp16 (K step = 16 k): slot = 1 K step, grad: 2 slots.
          bool need;
          int sub;
          int sub0 = -1;  // fp16 grad: a K step (16 k) is FOUR x - mu chunks = two K slots, owned by different thread groups
          if (F16) {
            need = true;
            sub = grad ? min(2 * (ks + kk) + 1, nsub - 1) : ks + kk;
            if (grad) sub0 = min(2 * (ks + kk), nsub - 1);
          } else {
            need = grad || (kk & 1) == 0;
            sub = grad ? ks + kk : (ks + kk) >> 1;
          }
          if (!nowait && need) {
#ifdef L2HMC_TC_PHASE_ACCOUNTING
            t0 = clock64();
#endif
            if (sub0 >= 0) mbar_wait(&a_sub[sub0], par);
            mbar_wait(&a_sub[sub], par);
            tcgen05_fence_after();
#ifdef L2HMC_TC_PHASE_ACCOUNTING
            w_a += clock64() - t0;
#endif
          }
          if (elect_one()) {
            const uint64_t dhi = desc0 + (uint64_t)(b16 + (2u * kk) * slab16);
            const uint64_t dlo = desc0 + (uint64_t)(b16 + (2u * kk + 1u) * slab16);
            const uint32_t ahi = S_AHI + 8u * (ks + kk), alo = S_ALO + 8u * (ks + kk);
            if (F16) {
              mma_f16_ts(acc, alo, dhi, idesc, (ks + kk) > 0);
              mma_f16_ts(acc, ahi, dlo, idesc, true);
              mma_f16_ts(acc, ahi, dhi, idesc, true);
            } else {
              mma_tf32_ts(acc, alo, dhi, idesc, (ks + kk) > 0);
              mma_tf32_ts(acc, ahi, dlo, idesc, true);
              mma_tf32_ts(acc, ahi, dhi, idesc, true);
            }
          }
          __syncwarp();
        }
      }
      if (elect_one()) tcgen05_commit(&S.empty[s]);
      __syncwarp();
      if (++s == NSLOT) { s = 0; ph ^= 1u; }
    }
    if (elect_one()) tcgen05_commit(&acc_rdy[gi & 1u]);
    __syncwarp();
    ++gi;
    if (!nowait) ++ai;
#ifdef L2HMC_TC_PHASE_ACCOUNTING
#pragma unroll
    for (int k = 0; k < 5; ++k)
      if (k == kind) {
        k_a[k] += w_a - w_a0;
        k_f[k] += w_f - w_f0;
        k_n[k] += 1;
      }
#endif
  });
#ifdef L2HMC_TC_PHASE_ACCOUNTING
  if (blockIdx.x == 0 && lane == 0) {
    g_tc_dbg[0] = w_a;
    g_tc_dbg[1] = w_f;
    g_tc_dbg[2] = clock64() - t_begin;
    g_tc_dbg[5] = gi;
    for (int k = 0; k < 5; ++k) {
      g_tc_dbg[8 + k] = k_a[k];
      g_tc_dbg[13 + k] = k_f[k];
      g_tc_dbg[18 + k] = k_n[k];
    }
  }
#endif
}

// what the heads epilogue prepares for the GEMM that follows it
enum { NEXT_NONE = 0, NEXT_X1 = 1, NEXT_X2 = 2, NEXT_G = 3, NEXT_V = 4 };

// Register budget per warp role (setmaxnreg): the CTA is launched with 12 warps = 3 warpgroups at the compile-time count
// (<= 168: three warps share one scheduler's 16 K registers); warpgroup 2 (MMA issuer, TMA producer, two idle warps) gives
// registers back and the two compute warpgroups take them: per scheduler 2 x 232 + 40 <= 512 registers per lane.
#ifndef L2HMC_TC_REGS_COMPUTE
#define L2HMC_TC_REGS_COMPUTE (L2HMC_TC_S_NQ == 3 ? 152 : 232)  // per scheduler: NQ x compute + 40 <= 512 registers per lane
#endif
#ifndef L2HMC_TC_REGS_AUX
#define L2HMC_TC_REGS_AUX 40
#endif
// compute threads per chain: the chunks of a chain's epilogue are dealt round-robin to NQ_S threads (warp w, lane l: chain
// 32 (w % 4) + l, chunks q = w / 4, w / 4 + NQ_S, ...).  The epilogues sit on the critical path (GEMM k+1 follows epilogue k
// slot by slot).  3 threads per chain (12 compute warps, registers rebalanced with setmaxnreg) were expected to shorten each
// epilogue by ~30 %; measured, the kernel got 2 % SLOWER: the epilogues are bound by what the SM can issue to its fp32 / alu
// pipes in total, not by per-thread latency -- so the lever is fewer instructions (packed FFMA2 / FMUL2 / FADD2, no dead adds).
#ifndef L2HMC_TC_S_NQ
#define L2HMC_TC_S_NQ 2   // measured on config 2 (profiles/r02_tc_s_experiments.txt): 2 -> 2.95 ms, 3 (with setmaxnreg) -> 3.02 ms
#endif
constexpr int NQ_S = L2HMC_TC_S_NQ;
constexpr int TC_S_THREADS = L2HMC_TC_SETMAXNREG ? MT * NQ_S + 128 : MT * NQ_S + 64;
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// NQC_T / NHC_T > 0: chunk counts fixed at compile time (the benchmark shapes: 13 x 13 = config 2, 8 x 13 = config 4);
// 0 x 0: read from the launch arguments -- one instantiation for every other (x_dim, width) inside the TMEM map
// (x_dim <= 52, width <= 104), so that those shapes do not fall to the 2.2x slower generic kernel of kernel_tc.cuh.
template <int NQC_T, int NHC_T, bool FAST, bool BIASG, bool F16>
__global__ void __launch_bounds__(TC_S_THREADS, 1) tc_transition_kernel_s(const __grid_constant__ TcArgs A) {
  const int NQC = NQC_T > 0 ? NQC_T : A.sh.DP / 4;
  const int NHC = NHC_T > 0 ? NHC_T : A.td.HK / 8;
  constexpr int NQ = NQ_S;
  constexpr int NCT = MT * NQ;  // compute threads: NQ per chain
  constexpr int W_MMA = NCT / 32, W_TMA = W_MMA + 1;
  const int DP = 4 * NQC, RS = (NQC & 1) ? DP : DP + 4;  // row stride with RS/4 odd: 8 lanes x 16 B hit 32 distinct banks
  // K slots (16 k) of the deepest GEMM = sub-barriers (biasg 2: the embed has one more, the bias K step)
  const int NSUB0 = ((NQC > NHC ? NQC : NHC) + 1) / 2;
  const int NSUB = (BIASG && A.td.biasg == 2 && NQC / 2 + 1 > NSUB0) ? NQC / 2 + 1 : NSUB0;
  const int CA = (NQC + 1) / 2, CB = NQC - CA;  // dimension chunks of heads_a / heads_b
  // TMEM map of this kernel (the host checks the same for run-time shapes, tc_s_shape_fits)
  static_assert(NQC_T == 0 || (8 * NQC_T <= 104 && 12 * ((NQC_T + 1) / 2) <= 112 && 12 * (NQC_T / 2) <= 80), "TMEM map of kernel_tc_s.cuh");
  static_assert(NHC_T == 0 || (8 * NHC_T <= 104 && (8 * NHC_T + 15) / 16 * 16 <= 112), "TMEM map of kernel_tc_s.cuh");
  static_assert((NQC_T == 0) == (NHC_T == 0), "chunk counts: both fixed or both run-time");
  constexpr float L2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  auto compute_bar = []() { compute_bar_n<NCT>(); };
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bars[2 * MAX_SLOT + 2 + NSUB_MAX + 2];
  __shared__ uint32_t tmem_slot;
  const Shape &sh = A.sh;
  const TcDims &td = A.td;
  const TransitionIO &io = A.io;
  const TcLayS L = make_tclay_s(RS, DP, sh.T);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler
  const int D = sh.D;
  const long long base = (long long)blockIdx.x * MT;
  Sync S{bars, bars + MAX_SLOT, bars + 2 * MAX_SLOT, bars + 2 * MAX_SLOT + 1};
  uint64_t *a_sub = bars + 2 * MAX_SLOT + 2;
  uint64_t *acc_rdy = a_sub + NSUB_MAX;  // two accumulator-ready barriers, GEMM g commits to acc_rdy[g & 1]
  float *ring = smem + L.ring;
  const uint32_t NSLOT = (uint32_t)td.nslot, SLOT_FLOATS = (uint32_t)td.slot_floats;

  if (tid == 0) {
    for (int s = 0; s < MAX_SLOT; ++s) {
      mbar_init(&S.full[s], 1);
      mbar_init(&S.empty[s], 1);
    }
    mbar_init(S.a_ready, 1);  // unused here (a_sub instead)
    // K slot p = the chunks 2p and 2p + 1 (16 k of the operand): one arrival per warp of their two owner groups
    for (int p = 0; p < NSUB_MAX; ++p) mbar_init(&a_sub[p], 8);
    mbar_init(S.acc_ready, 1);  // unused here (acc_rdy instead)
    mbar_init(&acc_rdy[0], 1);
    mbar_init(&acc_rdy[1], 1);
    fence_mbar_init();
  }
  if (warp == W_MMA) tmem_alloc(&tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tmem != 0u) __trap();  // this CTA owns all 512 columns: column / lane 0 is a constant in the issuer and below

  if (warp >= W_MMA) {
    if (L2HMC_TC_SETMAXNREG) reg_dec<L2HMC_TC_REGS_AUX>();  // the whole third warpgroup (two of its warps have nothing else to do)
    if (warp == W_TMA) producer_loop_s<F16>(A, NQC, S, ring, NSLOT, SLOT_FLOATS);
    else if (warp == W_MMA) issuer_loop_s<F16>(A, NQC, S, a_sub, acc_rdy, NSUB, ring, NSLOT, SLOT_FLOATS, lane);
  } else {
    // ===================== compute warps =====================
    if (L2HMC_TC_SETMAXNREG) reg_inc<L2HMC_TC_REGS_COMPUTE>();
    using I0 = std::integral_constant<int, 0>;
    using I1 = std::integral_constant<int, 1>;
    const int c = 32 * (warp & 3) + lane;  // chain within the tile == TMEM lane
    const int qd = warp >> 2;              // 0 .. NQ-1: this thread owns the chunks q = qd, qd + NQ, ... (warp-uniform)
    const uint32_t lb = ((uint32_t)(32 * (warp & 3))) << 16;
    const int qn = (NQC - qd + NQ - 1) / NQ;  // its 4-dim chunks: q = qd + NQ i, i < qn
    const int hn = (NHC - qd + NQ - 1) / NQ;  // its 8-column hidden chunks: q = qd + NQ i, i < hn
    const long long gch = base + c;
    const bool gauss = A.en.kind == 0;
    float *xr = smem + L.xs + c * RS, *vr = smem + L.vs + c * RS, *gr = smem + L.gs + c * RS;
    int *sdir = reinterpret_cast<int *>(smem + L.sdir), *sacc = reinterpret_cast<int *>(smem + L.sacc);
    uint32_t gi = 0;  // GEMM counter (parity of acc_ready)
#ifdef L2HMC_TC_PHASE_ACCOUNTING
    long long w_acc = 0;
    const long long t_begin = clock64();
    // per epilogue kind (0 embed, 1 hidden, 2 / 3 heads part 0 / 1 of the V net, 4 / 5 of the X net, 6 grad): cycles spent
    // working [k] and waiting for the accumulator [8 + k]; written by CTA 0, thread 0 (first thread of a chain) and thread 128
    long long ph_c[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long ph_t0 = 0, ph_w0 = 0;
#define L2HMC_PH_BEGIN() do { ph_t0 = clock64(); ph_w0 = w_acc; } while (0)
#define L2HMC_PH_END(k) do { const long long dw_ = w_acc - ph_w0; ph_c[8 + (k)] += dw_; ph_c[(k)] += clock64() - ph_t0 - dw_; } while (0)
#else
#define L2HMC_PH_BEGIN() do {} while (0)
#define L2HMC_PH_END(k) do {} while (0)
#endif
    auto wait_acc = [&]() {
#ifdef L2HMC_TC_PHASE_ACCOUNTING
      const long long t0 = clock64();
      mbar_wait_sleep(&acc_rdy[gi & 1u], (gi >> 1) & 1u);
      w_acc += clock64() - t0;
#else
      mbar_wait_sleep(&acc_rdy[gi & 1u], (gi >> 1) & 1u);
#endif
      ++gi;
      tcgen05_fence_after();
    };
    // chunk q of the next A operand is complete for this warp's chains: every thread's tcgen05.st has landed, one
    // arrival per warp on the chunk's K slot
    auto slot_done = [&](int q) {
      tmem_wait_st();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_sub[q >> 1]);
    };
    // ... and the arrivals this warp owes for chunks it would own but that do not exist / are handed over whole: every
    // chunk index c >= from with c % NQ == qd, up to the last K slot (each slot is armed for two arrivals per lane quarter)
    auto a_done = [&](int from) {
      tmem_wait_st();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        int c = from + ((qd - from) % NQ + NQ) % NQ;  // first index >= from owned by this thread group
        for (; c < 2 * NSUB; c += NQ) mbar_arrive(&a_sub[c >> 1]);
      }
    };
    const float eps = sh.eps, Tm = A.en.temperature, rTm = 1.f / A.en.temperature;
    float amax = 0.f;  // F16: largest |value| this thread put into an A operand
    if (F16) {
      // TMEM is not cleared by the allocation: K tails of the A operand that no epilogue writes (k >= 8 NHC of the
      // embed / hidden operands, the grad GEMM's tail before the first net call) must not hold NaN / Inf patterns
      const uint32_t z4[4] = {0u, 0u, 0u, 0u};
      for (int cq = qd; cq < 14; cq += NQ) {
        tmem_st4u(S_AHI + lb + 4 * cq, z4);
        tmem_st4u(S_ALO + lb + 4 * cq, z4);
      }
      tmem_wait_st();
    }

    for (int i = tid; i < sh.T * DP; i += NCT) smem[L.smask + i] = A.mask[i];
    // head constants of both nets into shared memory (X net first): as global loads their latency was ~18 % of the
    // heads epilogue (first use right after the load, 2 warps per scheduler)
    for (int i = tid; i < 2 * NQC * HCS_PER_CHUNK; i += NCT) {
      const int net = i / (NQC * HCS_PER_CHUNK), r = i - net * NQC * HCS_PER_CHUNK;
      const int q = r / HCS_PER_CHUNK, e = r - q * HCS_PER_CHUNK;  // e: 0-3 bs2, 4-7 bq2, 8-11 cS, 12-15 cQ, 16-19 bth
      const int src = e < 8 ? e : (e < 12 ? e + 4 : (e < 16 ? e + 8 : e + 8));
      smem[L.hcs + i] = (net ? A.vnet.hc : A.xnet.hc)[HC_PER_CHUNK * q + src];
    }
    for (int i = tid; i < MT * DP; i += NCT) {
      const int ch = i / DP, d = i - ch * DP;
      const long long g = base + ch;
      smem[L.xs + ch * RS + d] = (g < io.n && d < D) ? io.x[g * D + d] : 0.f;
    }
    compute_bar();

    float ljl = 0.f;
    for (int tr = 0; tr < io.n_transitions; ++tr) {
      const unsigned long long ctr = io.counter + (unsigned long long)tr;
      // ---- setup: momentum, direction, uniform -------------------------------------------------------
      if (io.v != nullptr) {
        for (int i = tid; i < MT * DP; i += NCT) {
          const int ch = i / DP, d = i - ch * DP;
          const long long g = base + ch;
          smem[L.vs + ch * RS + d] = (g < io.n && d < D) ? io.v[((long long)tr * io.n + g) * D + d] : 0.f;
        }
      } else {
#pragma unroll 1
        for (int i = 0; i < qn; ++i) {
          const int q = qd + NQ * i;
          float z[4];
          philox_normals4(io.seed, ctr, io.chain_offset + gch, q, z);
#pragma unroll
          for (int j = 0; j < 4; ++j) z[j] = (gch < io.n && 4 * q + j < D) ? z[j] : 0.f;
          sts4(vr + 4 * q, z);
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
        if (io.do_mh && io.u != nullptr) pu = (gch < io.n) ? io.u[(io.chain ? 0ll : (long long)tr * io.n) + gch] : 0.f;
        smem[L.su + c] = pu;
      }
      compute_bar();
      const bool fwd = sdir[c] != 0;
      const float sg = fwd ? 1.f : -1.f;
      // chain mode (chain_operator in one launch): the sub-proposals share one start point, one first Hamiltonian (paired
      // with init_v, not with their own momenta) and one log|J| accumulator; one Metropolis step after the last
      const bool chain_first = !io.chain || tr == 0, chain_last = !io.chain || tr == io.n_transitions - 1;
      if (chain_first) ljl = 0.f;  // log|J| of this thread's dimensions, in log2 units

      // ---- A operands ------------------------------------------------------------------------------------
      // net input of one 4-dim chunk: [a0..3 | b0..3] = K step q of the embed GEMM (weight rows permuted to match)
      // BIASG: the two pad dimensions of the last chunk's a-part select the time-embedding bias row of this chain's
      // direction in the embed weights ([1, 0] forward, [0, 1] backward), see tc_pack_net
      // biasg 2 (no pad dimensions: x_dim = 4 NQC, NQC even): the one-hot sits in a K step of its own behind the net input
      // (k = 8 NQC, 8 NQC + 1); the owner of the last chunk rewrites it with every operand (the hidden activations of the
      // GEMMs in between use the same columns), before the arrival that releases that K step (a_done(NQC))
      auto put_ab = [&](int q, const float (&a)[4], const float (&b)[4]) {
        float ab[8] = {a[0], a[1], a[2], a[3], b[0], b[1], b[2], b[3]};
        if (BIASG && q == NQC - 1) {
          if (td.biasg == 2) {
            // 8 values: a whole tf32 K step (TMEM is not cleared by the allocation); fp16: the rest of the 16-k step holds
            // finite values of earlier operands (zeroed at the start) that meet zero weight rows
            const float oh[8] = {fwd ? 1.f : 0.f, fwd ? 0.f : 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            put_a<F16, 8>(lb, 8 * NQC, oh, amax);
          } else {
            ab[2] = fwd ? 1.f : 0.f;
            ab[3] = fwd ? 0.f : 1.f;
          }
        }
        put_a<F16, 8>(lb, 8 * q, ab, amax);
      };
      // Gaussian grad GEMM input x - mu of one chunk (the K tail beyond DP is zeroed once per GEMM by zero_gtail)
      auto put_xmu = [&](int q, const float (&x)[4]) {
        const float4 mu = ldg4(A.en.mu + 4 * q);
        const float a[4] = {x[0] - mu.x, x[1] - mu.y, x[2] - mu.z, x[3] - mu.w};
        put_a<F16, 4>(lb, 4 * q, a, amax);
      };
      auto zero_gtail = [&]() {  // KG = DP rounded to 8: one more 4-column chunk of zeros when NQC is odd
        if (!F16 && (NQC & 1) && qd == NQC % NQ) {  // fp16: the K tail holds finite values of earlier operands, its B rows are 0
          const float z[4] = {0.f, 0.f, 0.f, 0.f};
          put_a<F16, 4>(lb, DP, z, amax);
        }
      };
      // RoughWell grad U of one chunk (utils/distributions.py:90-97)
      auto roughwell_grad = [&](int q, const float (&x)[4], float (&g)[4]) {
        const float e = A.en.s0, den = A.en.s1;
#pragma unroll
        for (int j = 0; j < 4; ++j) g[j] = (4 * q + j < D) ? (x[j] - e * sinf(x[j] / den) / den) * rTm : 0.f;
      };

      // ---- grad U at the current x (start of a transition): fills gs, prepares the V-net input ------------------
      // partial Hamiltonian over this thread's dims, accumulated where x, v, g are in registers anyway
      auto ham_chunk = [&](const float (&x)[4], const float (&v)[4], const float (&g)[4], int q, float &U, float &K) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          K = fmaf(v[j], v[j], K);
          if (gauss) {
            U = fmaf(x[j] - __ldg(A.en.mu + 4 * q + j), g[j], U);  // g carries 1/temperature
          } else if (4 * q + j < D) {
            U += (0.5f * x[j] * x[j] + A.en.s0 * cosf(x[j] / A.en.s1)) * rTm;
          }
        }
      };
      // Gaussian: epilogue of the grad GEMM (accumulator at column `acc`) -> gs, V-net input [x | g] handed over per
      // K slot; optional Hamiltonian partial
      auto grad_epilogue = [&](uint32_t acc, bool want_h, float &Hpart) {
        wait_acc();
        float U = 0.f, K = 0.f;
#pragma unroll 1
        for (int i = 0; i < qn; ++i) {
          const int q = qd + NQ * i;
          float g4[4];
          tmem_ld4(lb + acc + 4 * q, g4);
          const float4 xv = lds4(xr + 4 * q);
          const float x4[4] = {xv.x, xv.y, xv.z, xv.w};
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 4; ++j) g4[j] *= rTm;
          sts4(gr + 4 * q, g4);
          if (want_h) {
            const float4 vv = lds4(vr + 4 * q);
            const float v4[4] = {vv.x, vv.y, vv.z, vv.w};
            ham_chunk(x4, v4, g4, q, U, K);
          }
          put_ab(q, x4, g4);
          slot_done(q);
        }
        a_done(NQC);
        Hpart = (gauss ? 0.5f * U : U) + 0.5f * K;
      };
      // start of a transition: grad U(x), H(x, v) partial, first V-net input
      auto first_grad = [&](float &Hpart) {
        if (gauss) {
#pragma unroll 1
          for (int i = 0; i < qn; ++i) {
            const int q = qd + NQ * i;
            const float4 xv = lds4(xr + 4 * q);
            const float x4[4] = {xv.x, xv.y, xv.z, xv.w};
            put_xmu(q, x4);
            slot_done(q);
          }
          zero_gtail();
          a_done(NQC);
          grad_epilogue(S_R2, true, Hpart);
        } else {
          float U = 0.f, K = 0.f;
#pragma unroll 1
          for (int i = 0; i < qn; ++i) {
            const int q = qd + NQ * i;
            const float4 xv = lds4(xr + 4 * q), vv = lds4(vr + 4 * q);
            const float x4[4] = {xv.x, xv.y, xv.z, xv.w}, v4[4] = {vv.x, vv.y, vv.z, vv.w};
            float g4[4];
            roughwell_grad(q, x4, g4);
            sts4(gr + 4 * q, g4);
            ham_chunk(x4, v4, g4, q, U, K);
            put_ab(q, x4, g4);
            slot_done(q);
          }
          a_done(NQC);
          Hpart = U + 0.5f * K;
        }
      };
      // end of a transition: H(x', v') partial (g = grad U(x') is in gs: the last V-net call did not move x)
      auto final_ham = [&]() -> float {
        float U = 0.f, K = 0.f;
#pragma unroll 1
        for (int i = 0; i < qn; ++i) {
          const int q = qd + NQ * i;
          const float4 xv = lds4(xr + 4 * q), vv = lds4(vr + 4 * q), gv = lds4(gr + 4 * q);
          const float x4[4] = {xv.x, xv.y, xv.z, xv.w}, v4[4] = {vv.x, vv.y, vv.z, vv.w}, g4[4] = {gv.x, gv.y, gv.z, gv.w};
          ham_chunk(x4, v4, g4, q, U, K);
        }
        return (gauss ? 0.5f * U : U) + 0.5f * K;
      };

      // ---- relu(acc + bias) of this thread's 8-column chunks -> K steps of the next A operand (bias row may differ per
      // lane).  `acc`: accumulator column of this GEMM; `handover`: arrive per K slot (the next GEMM follows slot by
      // slot) or once at the end (the next GEMM is serial).
      auto hidden_epilogue = [&](uint32_t acc, const float *__restrict__ bias, bool handover) {
        wait_acc();
        float h[2][8];
        if (hn > 0) tmem_ld8(lb + acc + 8 * qd, h[0]);
        auto chunk = [&](int i, auto buf_c) {
          constexpr int B = decltype(buf_c)::value;
          const int q = qd + NQ * i;
          float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
          if (!BIASG) {  // BIASG: the bias is a weight row that meets a constant-1 column of the A operand
            b0 = ldg4(bias + 8 * q);
            b1 = ldg4(bias + 8 * q + 4);
          }
          tmem_wait_ld();
          if (i + 1 < hn) tmem_ld8(lb + acc + 8 * (q + NQ), h[B ^ 1]);
          const float(&hh)[8] = h[B];
          float a[8];
          constexpr bool RP = F16 && L2HMC_TC_RELU_PACK;  // relu folded into the fp16 conversion of put_a
          if (BIASG) {  // no `+ 0.f`: the compiler must keep that add (it turns -0 into +0)
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = RP ? hh[j] : fmaxf(hh[j], 0.f);
          } else if (RP) {
            a[0] = hh[0] + b0.x; a[1] = hh[1] + b0.y; a[2] = hh[2] + b0.z; a[3] = hh[3] + b0.w;
            a[4] = hh[4] + b1.x; a[5] = hh[5] + b1.y; a[6] = hh[6] + b1.z; a[7] = hh[7] + b1.w;
          } else {
            a[0] = fmaxf(hh[0] + b0.x, 0.f); a[1] = fmaxf(hh[1] + b0.y, 0.f); a[2] = fmaxf(hh[2] + b0.z, 0.f); a[3] = fmaxf(hh[3] + b0.w, 0.f);
            a[4] = fmaxf(hh[4] + b1.x, 0.f); a[5] = fmaxf(hh[5] + b1.y, 0.f); a[6] = fmaxf(hh[6] + b1.z, 0.f); a[7] = fmaxf(hh[7] + b1.w, 0.f);
          }
          put_a<F16, 8, RP>(lb, 8 * q, a, amax);
          if (handover) slot_done(q);
        };
        // two chunks per iteration (the register double buffer needs static names); rolled: the fully unrolled
        // version of this kernel had a 178 KB loop body and spent 40 % of the epilogue time on instruction fetch
#pragma unroll 1
        for (int i = 0; i < hn; i += 2) {
          chunk(i, I0{});
          if (i + 1 < hn) chunk(i + 1, I1{});
        }
        a_done(handover ? NHC : 0);
      };

      // ---- heads epilogue + fused state update (utils/dynamics.py:121-155 / :166-199) + next A operand --------------
      // MODE 0: momentum half step (V net, scale 1/2 eps); MODE 1: position half step xh in {0, 1} (X net, scale eps).
      // In log2 units: svl = cS * tanh(s + bs), fql = cQ * tanh(q + bq), with cS = e^{scale_s} * h * log2(e),
      // cQ = e^{scale_q} * eps * log2(e), h = 1/2 eps or eps; exp(+-sv) = 2^{+-svl}; log|J| += +-svl * ln 2.
      // The A operand of the GEMM that follows, for one chunk whose NEW state is in x4 / v4 (+ g4, mask row m4):
      auto prep = [&](auto mode_c, const int xh, const int next, int q, int i, const float (&x4)[4], const float (&v4)[4],
                      const float (&g4)[4], const float (&m4)[4]) {
        constexpr int MODE = decltype(mode_c)::value;
        if (MODE == 0) {
          if (next != NEXT_NONE) {
            // NEXT_X1: X net, first half: [v | k x], k = m (fwd) or 1 - m (bwd); NEXT_V: V net again at the same [x | g]
            const bool nx1 = next == NEXT_X1;
            float a[4], b[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              a[j] = nx1 ? v4[j] : x4[j];
              b[j] = nx1 ? (fwd ? m4[j] : 1.f - m4[j]) * x4[j] : g4[j];
            }
            put_ab(q, a, b);
            slot_done(q);
          }
        } else {
          if (next == NEXT_X2) {  // X net, second half: its k is this half's 1 - k
            const bool flip = (fwd != (xh == 0));
            float b[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = (flip ? m4[j] : 1.f - m4[j]) * x4[j];
            put_ab(q, v4, b);
            slot_done(q);
          } else if (gauss) {  // NEXT_G: grad U at the new x (K step i of the grad GEMM = the chunks i of both threads)
            put_xmu(q, x4);
            slot_done(q);
          } else {
            float g[4];
            roughwell_grad(q, x4, g);
            sts4(gr + 4 * q, g);
            put_ab(q, x4, g);
            slot_done(q);
          }
        }
      };
      // part 0: chunks q < CA (accumulator of heads_a in R1), state update only; part 1: chunks q >= CA (heads_b in R3),
      // preceded by the A operand of the part-0 chunks (their new state is read back from shared memory) -- A is not
      // written while heads_b still reads it, and the tensor pipe works on heads_b during part 0.
      auto heads_epilogue = [&](auto mode_c, auto part_c, const int xh, const int next, const TcNet &N, const float *mrow) {
        constexpr int MODE = decltype(mode_c)::value, PART = decltype(part_c)::value;
        const int CP = PART == 0 ? CA : CB;                                   // chunks of this part
        const uint32_t cS = PART == 0 ? S_R1 : S_R3 - 4 * CA;                 // S column of chunk q: cS + 4 q
        const uint32_t cT = cS + 4 * CP, cQ = cS + 8 * CP;
        const float hc = MODE == 0 ? 0.5f * eps : eps;
        const float *hcs_net = smem + L.hcs + (MODE == 0 ? NQC * HCS_PER_CHUNK : 0);  // V net (MODE 0) second
        const bool flip = (fwd != (xh == 0));  // MODE 1: k = m, or 1 - m when flipped
        const int na = qd < CA ? (CA - qd + NQ - 1) / NQ : 0;  // this thread's chunks in part 0 (q < CA): i < na
        const int ib = PART == 0 ? 0 : na, ie = PART == 0 ? na : qn;
        wait_acc();
        float s4[2][1][4], t4[2][1][4], q4[2][1][4];  // [register buffer][.][dimension]
        if (ib < ie) {
          const int q = qd + NQ * ib;
          tmem_ld4(lb + cS + 4 * q, s4[0][0]);
          tmem_ld4(lb + cT + 4 * q, t4[0][0]);
          tmem_ld4(lb + cQ + 4 * q, q4[0][0]);
        }
        if (PART == 1 && next != NEXT_NONE) {
#pragma unroll 1
          for (int i = 0; i < na; ++i) {  // deferred A operand of the part-0 chunks
            const int q = qd + NQ * i;
            const float4 xv = lds4(xr + 4 * q), vv = lds4(vr + 4 * q), mv = lds4(mrow + 4 * q), gv = lds4(gr + 4 * q);
            const float x4[4] = {xv.x, xv.y, xv.z, xv.w}, v4[4] = {vv.x, vv.y, vv.z, vv.w};
            const float m4[4] = {mv.x, mv.y, mv.z, mv.w}, g4[4] = {gv.x, gv.y, gv.z, gv.w};
            prep(mode_c, xh, next, q, i, x4, v4, g4, m4);
          }
        }
        // inputs of one chunk: head constants, state, mask row, (grad U)
        struct ChunkIn {
          float bs2[4], bq2[4], cSc[4], cQc[4], bth[4], x4[4], v4[4], m4[4], g4[4];
        };
        auto load_in = [&](int q, ChunkIn &c) {
          const float *hcq = hcs_net + HCS_PER_CHUNK * q;
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 c_bs = BIASG ? z4 : lds4(hcq), c_bq = BIASG ? z4 : lds4(hcq + 4), c_cs = lds4(hcq + 8), c_cq = lds4(hcq + 12);
          const float4 c_bt = BIASG ? z4 : lds4(hcq + 16);
          c.bs2[0] = c_bs.x; c.bs2[1] = c_bs.y; c.bs2[2] = c_bs.z; c.bs2[3] = c_bs.w;
          c.bq2[0] = c_bq.x; c.bq2[1] = c_bq.y; c.bq2[2] = c_bq.z; c.bq2[3] = c_bq.w;
          c.cSc[0] = c_cs.x; c.cSc[1] = c_cs.y; c.cSc[2] = c_cs.z; c.cSc[3] = c_cs.w;
          c.cQc[0] = c_cq.x; c.cQc[1] = c_cq.y; c.cQc[2] = c_cq.z; c.cQc[3] = c_cq.w;
          c.bth[0] = c_bt.x; c.bth[1] = c_bt.y; c.bth[2] = c_bt.z; c.bth[3] = c_bt.w;
          const float4 xv = lds4(xr + 4 * q), vv = lds4(vr + 4 * q), mv = lds4(mrow + 4 * q);
          c.x4[0] = xv.x; c.x4[1] = xv.y; c.x4[2] = xv.z; c.x4[3] = xv.w;
          c.v4[0] = vv.x; c.v4[1] = vv.y; c.v4[2] = vv.z; c.v4[3] = vv.w;
          c.m4[0] = mv.x; c.m4[1] = mv.y; c.m4[2] = mv.z; c.m4[3] = mv.w;
          c.g4[0] = c.g4[1] = c.g4[2] = c.g4[3] = 0.f;
          if (MODE == 0) {
            const float4 gv = lds4(gr + 4 * q);
            c.g4[0] = gv.x; c.g4[1] = gv.y; c.g4[2] = gv.z; c.g4[3] = gv.w;
          }
        };
        // the update of one chunk (utils/dynamics.py:121-155 / :166-199) from its accumulators sa / ta / qa
        auto update = [&](ChunkIn &c, const float (&sa)[4], const float (&ta)[4], const float (&qa)[4], float &lj) {
#if L2HMC_TC_F32X2
          if (FAST) {
            // Packed fp32 (FFMA2 / FMUL2 / FADD2, sm_100): the three-register scalar FFMA / FMUL / FADD issue at half rate on
            // this architecture, and this epilogue is bound by the fp32 pipe (~30 scalar fp32 instructions per dimension,
            // 66-75 % pipe occupancy measured).  The 4 dimensions of a chunk are independent: two per instruction.  Same
            // IEEE operations in the same order as the scalar form below (bit-identical per dimension); the selects
            // w = fwd ? 1 : -e and k = flip ? 1 - m : m become exact FMAs with per-thread constants.
            const float2 c2l = make_float2(2.f * L2E, 2.f * L2E), one2 = make_float2(1.f, 1.f), m2two = make_float2(-2.f, -2.f);
            const float2 hc2 = make_float2(hc, hc), sg2 = make_float2(sg, sg);
            const float2 wa2 = fwd ? make_float2(0.f, 0.f) : make_float2(-1.f, -1.f), wb2 = fwd ? one2 : make_float2(0.f, 0.f);
            const float2 ka2 = flip ? make_float2(-1.f, -1.f) : one2, kb2 = flip ? one2 : make_float2(0.f, 0.f);
            float2 lj2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 4; j += 2) {
              const float2 aS = __ffma2_rn(make_float2(sa[j], sa[j + 1]), c2l, make_float2(c.bs2[j], c.bs2[j + 1]));
              const float2 aQ = __ffma2_rn(make_float2(qa[j], qa[j + 1]), c2l, make_float2(c.bq2[j], c.bq2[j + 1]));
              // tanh(z) = 1 - 2 / (e^{2z} + 1); one reciprocal serves both heads (arguments clamped so that the
              // product of the two denominators stays finite: tanh(19.7) == 1 in fp32)
              const float2 eS = make_float2(ex2_approx(fminf(aS.x, 57.f)), ex2_approx(fminf(aS.y, 57.f)));
              const float2 eQ = make_float2(ex2_approx(fminf(aQ.x, 57.f)), ex2_approx(fminf(aQ.y, 57.f)));
              const float2 dS = __fadd2_rn(eS, one2), dQ = __fadd2_rn(eQ, one2);
              const float2 den = __fmul2_rn(dS, dQ);
              const float2 r = make_float2(rcp_approx(den.x), rcp_approx(den.y));
              const float2 svl = __fmul2_rn(make_float2(c.cSc[j], c.cSc[j + 1]), __ffma2_rn(__fmul2_rn(r, dQ), m2two, one2));  // cS * tanh
              const float2 fql = __fmul2_rn(make_float2(c.cQc[j], c.cQc[j + 1]), __ffma2_rn(__fmul2_rn(r, dS), m2two, one2));
              const float2 Tt = __ffma2_rn(make_float2(ta[j], ta[j + 1]), hc2, make_float2(c.bth[j], c.bth[j + 1]));  // h * (t + bt)
              const float2 eQx = make_float2(ex2_approx(fql.x), ex2_approx(fql.y));
              const float2 svs = __fmul2_rn(svl, sg2);  // +- (scale * S) * log2(e)
              const float2 e = make_float2(ex2_approx(svs.x), ex2_approx(svs.y));
              const float2 w = __ffma2_rn(e, wa2, wb2);  // fwd ? 1 : -e
              if (MODE == 0) {
                // fwd: v e + h (T - e^{fq} g) ; bwd: (v - h (T - e^{fq} g)) e
                const float2 hg = __fmul2_rn(hc2, make_float2(c.g4[j], c.g4[j + 1]));
                const float2 tmp = __ffma2_rn(make_float2(-eQx.x, -eQx.y), hg, Tt);
                const float2 v2 = __ffma2_rn(make_float2(c.v4[j], c.v4[j + 1]), e, __fmul2_rn(tmp, w));
                c.v4[j] = v2.x;
                c.v4[j + 1] = v2.y;
                lj2 = __fadd2_rn(lj2, svs);
              } else {
                const float2 k = __ffma2_rn(make_float2(c.m4[j], c.m4[j + 1]), ka2, kb2);  // flip ? 1 - m : m
                const float2 uu = __ffma2_rn(k, make_float2(-1.f, -1.f), one2);           // 1 - k
                // fwd: x e + h (e^{fq} v + T) ; bwd: e (x - h (e^{fq} v + T))
                const float2 x2 = make_float2(c.x4[j], c.x4[j + 1]);
                const float2 inner = __ffma2_rn(eQx, __fmul2_rn(hc2, make_float2(c.v4[j], c.v4[j + 1])), Tt);
                const float2 nx = __ffma2_rn(x2, e, __fmul2_rn(inner, w));
                const float2 xo = __ffma2_rn(uu, nx, __fmul2_rn(k, x2));
                c.x4[j] = xo.x;
                c.x4[j + 1] = xo.y;
                lj2 = __ffma2_rn(uu, svs, lj2);
              }
            }
            lj += lj2.x + lj2.y;
            return;
          }
#endif
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float svl, fql;
            if (FAST) {
              // tanh(z) = 1 - 2 / (e^{2z} + 1); one reciprocal serves both heads (arguments clamped so that the
              // product of the two denominators stays finite: tanh(19.7) == 1 in fp32)
              const float eS = ex2_approx(fminf(fmaf(sa[j], 2.f * L2E, c.bs2[j]), 57.f));
              const float eQ = ex2_approx(fminf(fmaf(qa[j], 2.f * L2E, c.bq2[j]), 57.f));
              const float dS = eS + 1.f, dQ = eQ + 1.f;
              const float r = rcp_approx(dS * dQ);
              svl = c.cSc[j] * fmaf(r * dQ, -2.f, 1.f);  // cS * tanh
              fql = c.cQc[j] * fmaf(r * dS, -2.f, 1.f);
            } else {
              svl = c.cSc[j] * tanhf((sa[j] * (2.f * L2E) + c.bs2[j]) * (0.5f * LN2));
              fql = c.cQc[j] * tanhf((qa[j] * (2.f * L2E) + c.bq2[j]) * (0.5f * LN2));
            }
            const float Tt = fmaf(ta[j], hc, c.bth[j]);  // h * (t + bt)
            const float eQx = FAST ? ex2_approx(fql) : exp2f(fql);
            const float svs = svl * sg;                     // +- (scale * S) * log2(e)
            const float e = FAST ? ex2_approx(svs) : exp2f(svs);
            const float w = fwd ? 1.f : -e;
            if (MODE == 0) {
              // fwd: v e + h (T - e^{fq} g) ; bwd: (v - h (T - e^{fq} g)) e
              const float tmp = fmaf(-eQx, hc * c.g4[j], Tt);
              c.v4[j] = fmaf(c.v4[j], e, tmp * w);
              lj += svs;
            } else {
              const float k = flip ? 1.f - c.m4[j] : c.m4[j];
              const float uu = 1.f - k;
              // fwd: x e + h (e^{fq} v + T) ; bwd: e (x - h (e^{fq} v + T))
              const float inner = fmaf(eQx, hc * c.v4[j], Tt);
              const float nx = fmaf(c.x4[j], e, inner * w);
              c.x4[j] = k * c.x4[j] + uu * nx;
              lj = fmaf(uu, svs, lj);
            }
          }
        };
        auto finish = [&](int q, int i, ChunkIn &c) {
          if (MODE == 0) sts4(vr + 4 * q, c.v4);
          else sts4(xr + 4 * q, c.x4);
          if (PART == 1) prep(mode_c, xh, next, q, i, c.x4, c.v4, c.g4, c.m4);
        };
        auto chunk = [&](int i, auto buf_c) {
          constexpr int B = decltype(buf_c)::value;
          const int q = qd + NQ * i;
          ChunkIn c;
          load_in(q, c);
          tmem_wait_ld();
          if (i + 1 < ie) {  // next chunk's accumulators travel while this chunk is processed
            tmem_ld4(lb + cS + 4 * (q + NQ), s4[B ^ 1][0]);
            tmem_ld4(lb + cT + 4 * (q + NQ), t4[B ^ 1][0]);
            tmem_ld4(lb + cQ + 4 * (q + NQ), q4[B ^ 1][0]);
          }
          update(c, s4[B][0], t4[B][0], q4[B][0], ljl);
          finish(q, i, c);
        };
#pragma unroll 1
        for (int i = ib; i < ie; i += 2) {
          chunk(i, I0{});
          if (i + 1 < ie) chunk(i + 1, I1{});
        }
        if (PART == 1) {
          if (MODE == 1 && next == NEXT_G && gauss) zero_gtail();  // before this warp's arrival on the last K step
          if (next != NEXT_NONE) a_done(NQC);
          else tcgen05_fence_before();
        } else {
          tcgen05_fence_before();
        }
      };

      float hpart0 = 0.f;
      first_grad(hpart0);
      if (io.chain) {
        // H(x_in, init_v): replace the kinetic part of this thread's dimensions by init_v's (utils/sampler.py:58-59, 79)
        float Kv = 0.f, K0 = 0.f;
#pragma unroll 1
        for (int i = 0; i < qn; ++i) {
          const int q = qd + NQ * i;
          const float4 vv = lds4(vr + 4 * q);
          Kv = fmaf(vv.x, vv.x, fmaf(vv.y, vv.y, fmaf(vv.z, vv.z, fmaf(vv.w, vv.w, Kv))));
          float z[4] = {0.f, 0.f, 0.f, 0.f};
          if (io.v0 == nullptr) philox_normals4(io.seed, chain_v0_counter(io), io.chain_offset + gch, q, z);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float v0 = (gch < io.n && 4 * q + j < D) ? (io.v0 ? io.v0[gch * D + 4 * q + j] : z[j]) : 0.f;
            K0 = fmaf(v0, v0, K0);
          }
        }
        hpart0 += 0.5f * (K0 - Kv);
      }
      if (chain_first) smem[L.part + qd * MT + c] = hpart0;
      compute_bar();
      if (qd == 0 && chain_first) {
        float h0s = 0.f;
#pragma unroll
        for (int r = 0; r < NQ; ++r) h0s += smem[L.part + r * MT + c];
        smem[L.h0 + c] = h0s;
      }

#pragma unroll 1
      for (int it = 0; it < sh.T; ++it) {
        const int trow = fwd ? it : sh.T - 1 - it;  // the step index this chain is at (utils/dynamics.py:285)
        const float *mrow = smem + L.smask + trow * DP;
        // four net calls per leapfrog step: V (momentum half step), X, X (position half steps), V
#pragma unroll 1
        for (int ni = 0; ni < 4; ++ni) {
          const bool isv = (ni == 0 || ni == 3);
          const TcNet &N = isv ? A.vnet : A.xnet;
          const float *tb = N.tb + (size_t)trow * td.N1;
          L2HMC_PH_BEGIN();
          hidden_epilogue(S_R1, tb, true);    // embed accumulator -> A of the hidden GEMM
          L2HMC_PH_END(0);
          L2HMC_PH_BEGIN();
          hidden_epilogue(S_R2, N.b4, true);  // hidden accumulator -> A of heads_a / heads_b
          L2HMC_PH_END(1);
          if (isv) {
            const int next = ni == 0 ? NEXT_X1 : (it + 1 < sh.T ? NEXT_V : NEXT_NONE);
            L2HMC_PH_BEGIN();
            heads_epilogue(I0{}, I0{}, 0, next, N, mrow);
            L2HMC_PH_END(2);
            L2HMC_PH_BEGIN();
            heads_epilogue(I0{}, I1{}, 0, next, N, mrow);
            L2HMC_PH_END(3);
          } else {
            const int next = ni == 1 ? NEXT_X2 : NEXT_G;
            L2HMC_PH_BEGIN();
            heads_epilogue(I1{}, I0{}, ni - 1, next, N, mrow);
            L2HMC_PH_END(4);
            L2HMC_PH_BEGIN();
            heads_epilogue(I1{}, I1{}, ni - 1, next, N, mrow);
            L2HMC_PH_END(5);
            if (ni == 2 && gauss) {
              float dummy;
              L2HMC_PH_BEGIN();
              grad_epilogue(S_R2, false, dummy);
              L2HMC_PH_END(6);
            }
          }
        }
      }

      // ---- log|J|, Hamiltonian, accept ---------------------------------------------------------------
      compute_bar();  // h0 readers are done with `part`
      if (!chain_last) continue;  // chain mode: the next sub-proposal starts from this proposal, no Metropolis step in between
      smem[L.part + qd * MT + c] = final_ham();
      compute_bar();
      float h1 = 0.f;
#pragma unroll
      for (int r = 0; r < NQ; ++r) h1 += smem[L.part + r * MT + c];
      compute_bar();
      smem[L.part + qd * MT + c] = ljl * LN2;
      compute_bar();
      const bool last = (tr == io.n_transitions - 1);
      if (qd == 0) {
        float logj = 0.f;
#pragma unroll
        for (int r = 0; r < NQ; ++r) logj += smem[L.part + r * MT + c];
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
            const float lx = smem[L.xs + ch * RS + d];
            io.x_out[g * D + d] = lx;
            if (io.v_out) io.v_out[g * D + d] = smem[L.vs + ch * RS + d];
            // the state this transition started from: the caller's x, or the x_next written one transition ago
            if (io.do_mh) {
              const float nx = sacc[ch] ? lx : ((tr == 0 || io.chain) ? io.x[g * D + d] : io.x_next[g * D + d]);
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
            const float nx = sacc[ch] ? smem[L.xs + ch * RS + d] : prev;
            io.x_next[g * D + d] = nx;
            if (io.trace) io.trace[((long long)tr * io.n + g) * D + d] = nx;
            smem[L.xs + ch * RS + d] = nx;
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
    if (blockIdx.x == 0 && (tid == 0 || tid == MT))
      for (int k = 0; k < 16; ++k) g_tc_dbg[24 + (tid == 0 ? 0 : 16) + k] = ph_c[k];
#endif
#undef L2HMC_PH_BEGIN
#undef L2HMC_PH_END
    // an A operand left the fp16 range (or was not finite): raise the CONTEXT's sticky status bit (pinned host-mapped word
    // the host polls without a device synchronisation; the library then stays on the tf32 split)
    if (F16 && !(amax < 60000.f)) {
      if (io.status) atomicOr_system(io.status, STATUS_F16_RANGE);
    }
    (void)Tm;
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
