// One S/T/Q net call of the layered engine as ONE kernel (the net of SCGExperiment.ipynb:51-77 / mnist_vae.py:131-178 at widths
// that do not fit the fused transition kernels: config 5's width 200 with the aux-encoding row term):
//     hd = [S | T | Q]( relu( relu([a | b] Wemb + tb[t(dir)] + enc(aux)) W4 + b4 ) Wh + bh )
// per 128-chain tile three chained tcgen05 GEMMs (fp16 x3 split, SS form).  The two hidden activations never leave the SM: an
// epilogue turns the accumulator in TMEM into the NEXT GEMM's operand image (layered::SplitImage, one 128-row block column)
// in shared memory, in place of the operand the finished GEMM no longer needs.  Same split, same k order and the same epilogue
// code (tcg::epi_chunk) as the three tc_gemm_pre_kernel launches it replaces: bit-identical results, minus two image round
// trips through HBM and two launches per net call.
//   warps 0-15  epilogues (TMEM lane group warp % 4, 16-column chunks warp / 4, + 4, ...)
//   warp 16     TMA: the tile's [a | b] image blocks into the operand region, then the k-blocks of the three packed weight
//               images through a 3-stage ring
//   warp 17     MMA issuer
// TMEM: accumulator 1 / 3 at columns [0, 256), accumulator 2 at [256, 512).  One tile at a time per CTA (the operand region
// is single-buffered): GEMM 1 of the next tile waits for epilogue 3.
#pragma once
#include "tc_gemm.cuh"

namespace l2hmc {
namespace tcg {

constexpr int N_NS = 3;  // ring stages of the weight stream
#ifndef L2HMC_NET_SLEEP_NS
#define L2HMC_NET_SLEEP_NS 64
#endif

// mbarrier wait that backs off between polls: one tile is in flight per CTA, so at any time either the 16 epilogue warps or
// the MMA / TMA warps are only waiting -- spinning, they take issue slots from the warps that work on the same scheduler.
__device__ __forceinline__ void wait_backoff(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t a = tc::smem_u32(bar);
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) break;
    if (L2HMC_NET_SLEEP_NS > 0) __nanosleep(L2HMC_NET_SLEEP_NS);
  }
}

struct NetFusedArgs {
  const uint8_t *a_img;  // [a | b] operand image (SplitImage, K = K1p)
  int img_nmb;
  long long M;
  TcGemmB w1, w2, w3;    // packed fp16 images of Wemb [K1p][Hp], W4 [Hp][Hp], Wh [Hp][N3p]; one n-block each
  int N1, N3;            // Hp, N3p
  const float *bias1, *bias1_b;  // time-embedding bias row of forward / backward chains
  const int *dir;
  const float *R; int ldr;       // enc(aux) rows or null
  const float *bias2, *bias3;
  float *hd; int ldc;
  unsigned int *status;
};

__host__ __device__ inline size_t net_a_bytes(int nkb1, int nkb2) { return (size_t)(nkb1 > nkb2 ? nkb1 : nkb2) * layered::SplitImage::BLOCK_BYTES; }
__host__ __device__ inline size_t net_stage_bytes(int BN1, int BN3) { return b_block_floats(BN1 > BN3 ? BN1 : BN3) * 4; }
__host__ __device__ inline size_t tc_net_smem(int nkb1, int nkb2, int BN1, int BN3) {
  return net_a_bytes(nkb1, nkb2) + N_NS * net_stage_bytes(BN1, BN3) + 1024 + 256 + 4 * 256 * 4;
}

__global__ void __launch_bounds__(G_THREADS, 1) tc_net_kernel(const NetFusedArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int nkb[3] = {p.w1.nkb, p.w2.nkb, p.w3.nkb};
  const int BNs[3] = {p.w1.BN, p.w2.BN, p.w3.BN};
  const size_t ABYTES = net_a_bytes(nkb[0], nkb[1]);
  const size_t SB = net_stage_bytes(BNs[0], BNs[2]);
  uint8_t *areg = smem;            // operand region: [k-block][hi | lo] 16 KB blocks of the current GEMM's A
  uint8_t *ring = smem + ABYTES;
  uint64_t *bars = reinterpret_cast<uint64_t *>(ring + N_NS * SB);
  uint64_t *full = bars, *empty = bars + N_NS, *a_full = bars + 2 * N_NS, *a_free = a_full + 1, *acc_full = a_full + 2 /*[3]*/,
           *a_ready = a_full + 5 /*[2]*/, *e3_done = a_full + 7;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(a_full + 8);
  float *sbias = reinterpret_cast<float *>(ring + N_NS * SB + 256);  // [bias1 fwd | bias1 bwd | bias2 | bias3] x 256 floats

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long tiles = (p.M + GM - 1) / GM;
  const long long my_tiles = (tiles > blockIdx.x) ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < N_NS; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    tc::mbar_init(a_full, 1);
    tc::mbar_init(a_free, 1);
    for (int i = 0; i < 3; ++i) tc::mbar_init(&acc_full[i], 1);
    tc::mbar_init(&a_ready[0], 16);
    tc::mbar_init(&a_ready[1], 16);
    tc::mbar_init(e3_done, 16);
    tc::fence_mbar_init();
  }
  if (warp == W_MMA) tc::tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < 1024; i += G_THREADS) {
    const int which = i >> 8, j = i & 255;
    const float *src = which == 0 ? p.bias1 : (which == 1 ? p.bias1_b : (which == 2 ? p.bias2 : p.bias3));
    sbias[i] = (j < (which == 3 ? p.N3 : p.N1) && src != nullptr) ? src[j] : 0.f;
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 16) {
    // ---------------- epilogues ----------------
    const int lg = warp & 3, cpart = warp >> 2;
    const int r = lg * 32 + lane;  // row inside the tile = TMEM lane
    float omax = 0.f;
#ifdef L2HMC_NET_PHASES
    long long ep_wait[3] = {0, 0, 0}, ep_ld[3] = {0, 0, 0}, ep_work[3] = {0, 0, 0}, ep_math[2] = {0, 0}, ep_chunk[2] = {0, 0}, ep_fence[2] = {0, 0};
#endif
    for (long long ti = 0; ti < my_tiles; ++ti) {
      const long long mb = blockIdx.x + ti * gridDim.x;
      const long long m = mb * GM + r;
      const bool mok = m < p.M;
      const uint32_t ph = (uint32_t)ti & 1u;
      const uint32_t trow = tmem_base + ((uint32_t)(32 * lg) << 16);
      if (p.R != nullptr && mok) {  // the row term of epilogue 1 comes from HBM: ask for this warp's part of the row while GEMM 1 runs
        const float *rp = p.R + m * (long long)p.ldr;
        for (int c = cpart * 16; c < p.N1; c += 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + c));
      }
      // stages 1 and 2: accumulator -> relu tail -> operand image of the next GEMM in the operand region.  The arithmetic is
      // epi_chunk<EPI_RELU>'s, operation for operation ((acc + bias) + row term, max with 0; columns >= N are zero); what
      // differs is when the operands arrive: biases sit in shared memory, and the enc(aux) row values of chunk k + 1 are
      // requested before chunk k is processed (the first before the accumulator wait) -- an L2 / DRAM round trip per chunk
      // was most of this epilogue (ncu source page: stall_long_sb on the bias add, stall_lg on the row loads).
#pragma unroll 1
      for (int st = 0; st < 2; ++st) {
        const int BN = BNs[st], Nn = p.N1, kpad = nkb[st + 1] * 32;
        const uint32_t bias_s = tc::smem_u32(st == 0 ? ((p.dir != nullptr && mok && p.dir[m] == 0) ? sbias + 256 : sbias) : sbias + 512);
        const float *Rrow = (st == 0 && p.R != nullptr && mok) ? p.R + m * (long long)p.ldr : nullptr;
        auto load_r = [&](int c, float4(&rv)[4]) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            rv[q] = (Rrow != nullptr && c + 4 * q + 3 < Nn) ? __ldg(reinterpret_cast<const float4 *>(Rrow + c + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        float4 rcur[4], rnext[4];
        load_r(cpart * 16, rcur);
#ifdef L2HMC_NET_PHASES
        long long e0 = clock64();
#endif
        wait_backoff(&acc_full[st], ph);
        tc::tcgen05_fence_after();
#ifdef L2HMC_NET_PHASES
        ep_wait[st] += clock64() - e0;
        e0 = clock64();
#endif
        const uint32_t tcol = trow + (uint32_t)(st * 256);
        for (int c = cpart * 16; c < kpad; c += 64) {
          float acc[16];
#ifdef L2HMC_NET_PHASES
          const long long l0 = clock64();
#endif
          if (c < BN) tc::tmem_ld16(tcol + (uint32_t)c, acc);
          load_r(c + 64, rnext);
          if (c < BN) tc::tmem_wait_ld();
#ifdef L2HMC_NET_PHASES
          ep_ld[st] += clock64() - l0;
#endif
#ifdef L2HMC_NET_PHASES
          const long long m0_ = clock64();
#endif
          if (mok) {
            uint8_t *ip = areg + layered::SplitImage::piece(r, c, 1);
            if (c < BN && c < Nn) {
              // branch-free: biases (zero-padded to 256 in shared memory) as four 16-byte shared loads, columns >= N forced to zero
              const uint32_t ba = bias_s + (uint32_t)c * 4u;
              float bb[16];
#pragma unroll
              for (int q = 0; q < 4; ++q)
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bb[4 * q]), "=f"(bb[4 * q + 1]), "=f"(bb[4 * q + 2]), "=f"(bb[4 * q + 3]) : "r"(ba + 16u * q));
              const float rr[16] = {rcur[0].x, rcur[0].y, rcur[0].z, rcur[0].w, rcur[1].x, rcur[1].y, rcur[1].z, rcur[1].w,
                                    rcur[2].x, rcur[2].y, rcur[2].z, rcur[2].w, rcur[3].x, rcur[3].y, rcur[3].z, rcur[3].w};
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float o = fmaxf((acc[j] + bb[j]) + rr[j], 0.f);
                acc[j] = (c + j < Nn) ? o : 0.f;
              }
#ifdef L2HMC_NET_PHASES
              if (acc[3] == 123.456f) printf("x");  // keep the math before the clock read
              ep_math[st] += clock64() - m0_;
#endif
#pragma unroll
              for (int p8 = 0; p8 < 2; ++p8) {
                float a8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) a8[j] = acc[8 * p8 + j];
                uint4 hi, lo;
                layered::split8_to_half(a8, hi, lo, omax);
                *reinterpret_cast<uint4 *>(ip + p8 * 2048) = hi;
                *reinterpret_cast<uint4 *>(ip + 8192 + p8 * 2048) = lo;
              }
            } else {  // K padding of the next GEMM: zero columns
              const uint4 z = make_uint4(0u, 0u, 0u, 0u);
              *reinterpret_cast<uint4 *>(ip) = z;
              *reinterpret_cast<uint4 *>(ip + 2048) = z;
              *reinterpret_cast<uint4 *>(ip + 8192) = z;
              *reinterpret_cast<uint4 *>(ip + 8192 + 2048) = z;
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) rcur[q] = rnext[q];
#ifdef L2HMC_NET_PHASES
          ep_chunk[st] += clock64() - m0_;
#endif
        }
#ifdef L2HMC_NET_PHASES
        const long long f0_ = clock64();
#endif
        tc::tcgen05_fence_before();
        tc::fence_proxy_async_smem();  // the image was written through the generic proxy, the MMAs read it through the async one
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&a_ready[st]);
#ifdef L2HMC_NET_PHASES
        ep_work[st] += clock64() - e0;
        ep_fence[st] += clock64() - f0_;
#endif
      }
      // stage 3: heads -> hd (fp32)
#ifdef L2HMC_NET_PHASES
      long long e3 = clock64();
#endif
      wait_backoff(&acc_full[2], ph);
      tc::tcgen05_fence_after();
#ifdef L2HMC_NET_PHASES
      ep_wait[2] += clock64() - e3;
      e3 = clock64();
#endif
      {
        // acc + bias (epi_chunk<EPI_BIAS>'s arithmetic), bias from shared memory, 16-byte stores when the rows allow it
        const bool vec = ((p.ldc % 4) == 0 && (p.N3 % 4) == 0 && (reinterpret_cast<uintptr_t>(p.hd) % 16) == 0);
        const uint32_t b3 = tc::smem_u32(sbias + 768);
        float *Crow = p.hd + (mok ? m : 0) * (long long)p.ldc;
        for (int c = cpart * 16; c < BNs[2]; c += 64) {
          float acc[16];
          tc::tmem_ld16(trow + (uint32_t)c, acc);
          float bb[16];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bb[4 * q]), "=f"(bb[4 * q + 1]), "=f"(bb[4 * q + 2]), "=f"(bb[4 * q + 3]) : "r"(b3 + (uint32_t)c * 4u + 16u * q));
          tc::tmem_wait_ld();
          if (!mok || c >= p.N3) continue;
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] += bb[j];
          if (vec && c + 15 < p.N3) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<float4 *>(Crow + c + 4 * q) = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c + j < p.N3) Crow[c + j] = acc[j];
          }
        }
      }
      tc::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(e3_done);
#ifdef L2HMC_NET_PHASES
      ep_work[2] += clock64() - e3;
#endif
    }
#ifdef L2HMC_NET_PHASES
    if (blockIdx.x == 0 && tid == 0)
      printf("NETEPI wait %lld %lld %lld  work %lld %lld %lld  ld %lld %lld  math %lld %lld  chunk %lld %lld  fence %lld %lld\n", ep_wait[0],
             ep_wait[1], ep_wait[2], ep_work[0], ep_work[1], ep_work[2], ep_ld[0], ep_ld[1], ep_math[0], ep_math[1], ep_chunk[0], ep_chunk[1],
             ep_fence[0], ep_fence[1]);
#endif
    if (!(omax < 60000.f) && p.status) atomicOr_system(p.status, l2hmc::STATUS_F16_RANGE);
  } else if (warp == W_TMA) {
    if (lane == 0) {
      long long flat = 0;
      for (long long ti = 0; ti < my_tiles; ++ti) {
        const long long mb = blockIdx.x + ti * gridDim.x;
        wait_backoff(a_free, ((uint32_t)ti & 1u) ^ 1u);  // GEMM 3 of the previous tile has read the operand region
        tc::mbar_arrive_expect_tx(a_full, (uint32_t)(nkb[0] * layered::SplitImage::BLOCK_BYTES));
        for (int kb = 0; kb < nkb[0]; ++kb)
          tc::bulk_g2s(areg + (size_t)kb * layered::SplitImage::BLOCK_BYTES,
                       p.a_img + ((size_t)kb * p.img_nmb + (size_t)mb) * layered::SplitImage::BLOCK_BYTES, layered::SplitImage::BLOCK_BYTES, a_full);
        for (int gi = 0; gi < 3; ++gi) {
          const TcGemmB &w = gi == 0 ? p.w1 : (gi == 1 ? p.w2 : p.w3);
          const uint32_t bytes = (uint32_t)(b_block_floats(w.BN) * 4);
          for (int kb = 0; kb < w.nkb; ++kb, ++flat) {
            const int s = (int)(flat % N_NS);
            const uint32_t ph = (uint32_t)(flat / N_NS) & 1u;
            wait_backoff(&empty[s], ph ^ 1u);
            tc::mbar_arrive_expect_tx(&full[s], bytes);
            tc::bulk_g2s(ring + s * SB, w.pk + (size_t)kb * b_block_floats(w.BN), bytes, &full[s]);
          }
        }
      }
    }
  } else {
    if (lane == 0) {
      const uint32_t lbo_a = (GM / 8) * 128;
      long long flat = 0;
#ifdef L2HMC_NET_PHASES
      long long t_wait_a[3] = {0, 0, 0}, t_wait_b[3] = {0, 0, 0}, t_start = clock64();
#define NET_T0 const long long t0_ = clock64()
#define NET_ACC(x) x += clock64() - t0_
#else
#define NET_T0
#define NET_ACC(x)
#endif
      for (long long ti = 0; ti < my_tiles; ++ti) {
        const uint32_t ph = (uint32_t)ti & 1u;
        for (int gi = 0; gi < 3; ++gi) {
          const int BN = BNs[gi];
          const uint32_t idesc = make_idesc_f16g(GM, BN);
          const uint32_t lbo_b = (uint32_t)(BN / 8) * 128, B_HALF = (uint32_t)BN * GBK * 4;
          // the operand: TMA for GEMM 1 (and accumulator 1 drained by the previous tile's epilogue 3), epilogue images after
          {
            NET_T0;
            if (gi == 0) {
              wait_backoff(a_full, ph);
              wait_backoff(e3_done, ph ^ 1u);
            } else {
              wait_backoff(&a_ready[gi - 1], ph);
            }
            NET_ACC(t_wait_a[gi]);
          }
          tc::tcgen05_fence_after();
          const uint32_t dacc = tmem_base + (uint32_t)(gi == 1 ? 256 : 0);
          for (int kb = 0; kb < nkb[gi]; ++kb, ++flat) {
            const int s = (int)(flat % N_NS);
            const uint32_t sph = (uint32_t)(flat / N_NS) & 1u;
            {
              NET_T0;
              wait_backoff(&full[s], sph);
              NET_ACC(t_wait_b[gi]);
            }
            tc::tcgen05_fence_after();
            const uint32_t sa = tc::smem_u32(areg + (size_t)kb * layered::SplitImage::BLOCK_BYTES), sb = tc::smem_u32(ring + s * SB);
#pragma unroll
            for (int ks = 0; ks < GBK / 8; ++ks) {
              const uint64_t a_hi = tc::make_smem_desc(sa + ks * 2 * lbo_a, lbo_a, 128);
              const uint64_t a_lo = tc::make_smem_desc(sa + 8192 + ks * 2 * lbo_a, lbo_a, 128);
              const uint64_t b_hi = tc::make_smem_desc(sb + ks * 2 * lbo_b, lbo_b, 128);
              const uint64_t b_lo = tc::make_smem_desc(sb + B_HALF + ks * 2 * lbo_b, lbo_b, 128);
              mma_f16_ss(dacc, a_lo, b_hi, idesc, (kb | ks) != 0);
              mma_f16_ss(dacc, a_hi, b_lo, idesc, true);
              mma_f16_ss(dacc, a_hi, b_hi, idesc, true);
            }
            tc::tcgen05_commit(&empty[s]);
          }
          tc::tcgen05_commit(&acc_full[gi]);
          if (gi == 2) tc::tcgen05_commit(a_free);
        }
      }
#ifdef L2HMC_NET_PHASES
      if (blockIdx.x == 0)
        printf("NETPHASE tiles=%lld total=%lld wait_a: %lld %lld %lld  wait_b: %lld %lld %lld\n", my_tiles, clock64() - t_start, t_wait_a[0],
               t_wait_a[1], t_wait_a[2], t_wait_b[0], t_wait_b[1], t_wait_b[2]);
#endif
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc::tcgen05_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// fits: one n-block per weight, accumulators <= 256 columns, operand region + ring inside the shared-memory budget
inline bool tc_net_fits(const NetFusedArgs &p, size_t smem_limit) {
  return (p.N1 % 4) == 0 && p.w1.f16 && p.w2.f16 && p.w3.f16 && p.w1.nblk == 1 && p.w2.nblk == 1 && p.w3.nblk == 1 && p.w1.BN == p.w2.BN &&
         p.w2.nkb * 32 >= p.N1 && p.w3.nkb == p.w2.nkb && tc_net_smem(p.w1.nkb, p.w2.nkb, p.w1.BN, p.w3.BN) <= smem_limit;
}

inline cudaError_t launch_tc_net(const NetFusedArgs &p, int sms, cudaStream_t s) {
  const size_t smem = tc_net_smem(p.w1.nkb, p.w2.nkb, p.w1.BN, p.w3.BN);
  const long long tiles = (p.M + GM - 1) / GM;
  const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
  static size_t configured_dev[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  size_t &configured = configured_dev[dev & 63];
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_net_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  tc_net_kernel<<<grid, G_THREADS, smem, s>>>(p);
  return cudaGetLastError();
}

}  // namespace tcg
}  // namespace l2hmc
