// Tensor-core GEMM of the layered engine: C = epi(A B + bias [+ R]) with fp32 inputs / outputs and fp32-level
// accuracy from three tf32 tcgen05 MMAs per product (a_lo b_hi + a_hi b_lo + a_hi b_hi, fp32 accumulation in TMEM),
// the same split the fused kernel (kernel_tc.cuh) uses.  Same GemmArgs / epilogues as layered::sgemm_kernel.
//
//   Persistent CTAs (one per SM) walk 256 x BN output tiles (BN <= 256, multiple of 16) as two 128-row halves that
//   share every B stage, so the weight stream per flop -- the first version's limit: 40 KB of L2->SM traffic per 768
//   tensor cycles -- is halved.  18 warps:
//     warps 0-7   A path: LDG fp32 rows (three k-blocks in flight in registers) -> tf32 hi / lo -> STS into the UMMA
//                 K-major core-matrix layout.
//     warps 8-15  epilogue: warp e reads the 32 TMEM lanes (e % 4) of accumulator (e / 4) = 32 output rows, applies
//                 the fused layer tail and stores.
//     warp 16     B path: one cp.async.bulk (TMA) per k-block of the host-packed, pre-split weight image.
//     warp 17     one thread issues the tcgen05.mma's (SS form: A and B from shared memory) and commits each stage back
//                 to the producers; owns the TMEM allocation (2 accumulators x 256 columns).
//   3-stage ring of 64 KB, k-block = 16 (two K=8 MMA steps x 3 products x 2 row halves); mbarriers full_a / full_b /
//   empty per stage and acc_full / acc_empty for the accumulators.
//
// F16 = true (L2HMC_LAYERED_GEMM=f16): the same split with fp16 pairs and kind::f16 MMAs (K = 16 per MMA at the rate of a
// K = 8 tf32 MMA): a stage of the same 64 KB then holds a k-block of 32 instead of 16 -- half the tensor time AND half the
// shared-memory operand traffic per flop, which is what bounds the tf32 version (SS form).  fp16 range: |values| < 65504.
//
// Shared-memory operand layout (no swizzle, K-major): core matrix = 8 rows x 16 B (4 tf32) stored as 128 contiguous
// bytes; [k-chunk][row-group][8][16 B], so LBO (K-adjacent core matrices) = rows/8 * 128 B and SBO (adjacent 8-row
// groups) = 128 B -- the encoding verified by tools/tc_probe (profiles/r01_tc_probe.txt).
#pragma once
#include <cuda_fp16.h>

#include <vector>

#include "layered.cuh"
#include "tc_common.cuh"

namespace l2hmc {
namespace tcg {

constexpr int GM = 128;        // rows per MMA (one accumulator)
constexpr int GROWS = 2 * GM;  // rows per CTA tile
constexpr int GBK = 16;        // k-block per pipeline stage
constexpr int GNS = 3;         // stages
constexpr int G_THREADS = 576; // warps 0-7 A path, 8-15 epilogue, 16 TMA (B), 17 MMA
constexpr int W_TMA = 16, W_MMA = 17;

// Host-packed B: for n-block nb and k-block kb, a contiguous image [hi|lo][GBK/4][BN/8][8 rows][4 floats].
struct TcGemmB {
  const float *pk;
  int BN;    // columns per n-block (multiple of 16, <= 256)
  int nblk;  // n-blocks
  int nkb;   // k-blocks (K padded to a multiple of the k-block with zero rows)
  int f16;   // image holds fp16 hi / lo, k-block = 32 (else tf32, k-block = 16)
  unsigned int *status;  // the context's status word (host-mapped): the fp16 A path raises STATUS_F16_RANGE on an out-of-range activation
};

__host__ __device__ inline size_t b_block_floats(int BN) { return (size_t)2 * GBK * BN; }
__host__ __device__ inline size_t stage_bytes(int BN) { return (size_t)2 * GROWS * GBK * 4 + b_block_floats(BN) * 4; }
__host__ __device__ inline size_t tc_gemm_smem(int BN) { return GNS * stage_bytes(BN) + 1024 + 256; }

// D[tmem] (+)= A[smem] * B[smem]^T, one K=8 slice; issued by ONE thread.
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}

__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__host__ __device__ __forceinline__ uint32_t make_idesc_f16g(int M, int N) {  // kind::f16: A / B format 0 = f16, D format 1 = f32
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack_h2g(float k_even, float k_odd) {  // low half = lower k
  const __half2 h = __floats2half2_rn(k_even, k_odd);
  return *reinterpret_cast<const uint32_t *>(&h);
}

__device__ __forceinline__ void wait_spin(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t a = tc::smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  }
}

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// The fused layer tails (same meaning as in layered::sgemm_kernel); v already carries the bias.
template <int EPI>
__device__ __forceinline__ float epi_apply(float v, float aux, float scale) {
  if (EPI == layered::EPI_RELU) return fmaxf(v + aux, 0.f);                              // aux: + enc(aux) row term
  if (EPI == layered::EPI_SOFTPLUS) return fmaxf(v, 0.f) + layered::log1p_unit(expf(-fabsf(v)));   // tf.nn.softplus
  if (EPI == layered::EPI_DSOFTPLUS) return v * (1.f - expf(-aux));                       // aux: stored softplus output
  if (EPI == layered::EPI_ADD_SCALE) return (v + aux) * scale;                            // aux: z
  return v;
}

// One 16-column chunk of an output row: acc = A B for columns [nbase, nbase + 16) of row m (m < M, nbase < N) -> fused layer
// tail -> fp32 C and / or the operand image of the next GEMM.  Shared by the GEMM kernels below.
// img_row / img_nmb: the row and 128-row-block count used for the image address (default: the global row of g.c_img; the fused
// net kernel writes a one-block image in shared memory with the row inside the tile).
template <int EPI>
__device__ __forceinline__ void epi_chunk(const layered::GemmArgs &g, float (&acc)[16], long long m, int nbase, const float *bias,
                                          float &omax, long long img_row = -1, int img_nmb = 0) {
  float *Cp = g.C + m * (long long)g.ldc + nbase;
  const bool full = nbase + 15 < g.N;
  if (full && g.vec) {
    // fast path: whole 16-column chunk inside N, 16-byte aligned rows
    float aux[16];
    if (EPI == layered::EPI_DSOFTPLUS) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 t = *reinterpret_cast<const float4 *>(Cp + 4 * q);
        aux[4 * q] = t.x; aux[4 * q + 1] = t.y; aux[4 * q + 2] = t.z; aux[4 * q + 3] = t.w;
      }
    } else if (EPI == layered::EPI_ADD_SCALE || (EPI == layered::EPI_RELU && g.R != nullptr)) {
      const float *Rp = g.R + m * (long long)g.ldr + nbase;
#pragma unroll
      for (int j = 0; j < 16; ++j) aux[j] = Rp[j];
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) aux[j] = 0.f;
    }
    if (EPI != layered::EPI_DSOFTPLUS && EPI != layered::EPI_ADD_SCALE && bias != nullptr) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(bias + nbase + 4 * q));
        acc[4 * q] += t.x; acc[4 * q + 1] += t.y; acc[4 * q + 2] += t.z; acc[4 * q + 3] += t.w;
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = epi_apply<EPI>(acc[j], aux[j], g.scale);
    if (!g.no_c) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4 *>(Cp + 4 * q) = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int nn = nbase + j;
      float o = 0.f;
      if (nn < g.N) {
        float v = acc[j], a = 0.f;
        if (EPI == layered::EPI_DSOFTPLUS) a = Cp[j];
        else if (EPI == layered::EPI_ADD_SCALE || (EPI == layered::EPI_RELU && g.R != nullptr)) a = g.R[m * (long long)g.ldr + nn];
        if (EPI != layered::EPI_DSOFTPLUS && EPI != layered::EPI_ADD_SCALE && bias != nullptr) v += bias[nn];
        o = epi_apply<EPI>(v, a, g.scale);
        if (!g.no_c) Cp[j] = o;
      }
      acc[j] = o;  // columns >= N of the operand image are zero
    }
  }
  if (g.c_img != nullptr) {
    // the same values as the next GEMM's operand image: columns [nbase, nbase + 16) = two 16-byte pieces of hi and of lo;
    // the 32 lanes of the warp (consecutive rows) write 512 contiguous bytes per piece
    uint8_t *ip = g.c_img + (img_row >= 0 ? layered::SplitImage::piece(img_row, nbase, img_nmb) : layered::SplitImage::piece(m, nbase, g.img_nmb));
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
  }
}

// APRE (needs F16): A arrives as a pre-split operand image (layered::SplitImage, g.a_img) and is fetched by the TMA warp with
// one 32 KB bulk copy per k-block; nobody converts, and warps 0-7 join the epilogue (16 epilogue warps, alternate chunks).
template <int EPI, bool F16, bool APRE = false>
__global__ void __launch_bounds__(G_THREADS, 1) tc_gemm_kernel(const layered::GemmArgs g, const TcGemmB tb) {
  static_assert(!APRE || F16, "the operand image holds fp16 pairs");
  constexpr int KB = F16 ? 2 * GBK : GBK;  // K elements per k-block (the stage bytes are the same)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte aligned base (descriptors address in 16-byte units; keep stages well aligned)
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int BN = tb.BN;
  const size_t SB = stage_bytes(BN);
  const uint32_t A_IMG = GM * GBK * 4;   // bytes of one A image (128 rows, hi or lo); stage: [A0hi|A0lo|A1hi|A1lo|Bhi|Blo]
  const uint32_t A_ALL = 4 * A_IMG;
  const uint32_t B_HALF = (uint32_t)BN * GBK * 4;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + GNS * SB);
  uint64_t *full_a = bars, *full_b = bars + GNS, *empty = bars + 2 * GNS, *acc_full = bars + 3 * GNS,
           *acc_empty = bars + 3 * GNS + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * GNS + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = tb.nkb, nblk = tb.nblk;
  const long long mblocks = (g.M + GROWS - 1) / GROWS;
  const long long tiles = mblocks * nblk;
  // persistent CTA: tiles blockIdx.x, blockIdx.x + gridDim.x, ... ; n-block fastest so neighbouring CTAs share A rows in L2
  const long long my_tiles = (tiles > blockIdx.x) ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < GNS; ++s) {
      tc::mbar_init(&full_a[s], 8);
      tc::mbar_init(&full_b[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    tc::mbar_init(acc_full, 1);
    tc::mbar_init(acc_empty, APRE ? 16 : 8);
    tc::fence_mbar_init();
  }
  if (warp == W_MMA) tc::tmem_alloc(tmem_slot, 512);  // two 256-column accumulators (row halves of the tile)
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (!APRE && warp < 8) {
    // ---------------- A path: LDG (3 k-blocks in flight) -> tf32 hi / lo -> UMMA layout ----------------
    // element i of this thread: idx = i*256 + tid -> r8 = idx & 7, kc = (idx >> 3) & 3, mg = idx >> 5 (0..31)
    int mrow[4], kofs[4];
    uint32_t sofs[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = i * 256 + tid;
      const int r8 = idx & 7, kc = (idx >> 3) & 3, mg = idx >> 5;
      mrow[i] = mg * 8 + r8;
      kofs[i] = kc * 4;
      sofs[i] = (uint32_t)((mg >> 4) * 2 * A_IMG + (kc * (GM / 8) + (mg & 15)) * 128 + r8 * 16);
    }
    const long long total = my_tiles * nkb;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float amax = 0.f;  // F16: largest |activation| this thread converted (fp16 overflows at 65504)
    constexpr int NV = F16 ? 2 : 1;  // float4 per 16-byte piece of the operand image (8 halfs or 4 tf32)
    auto load_blk = [&](long long flat, float4(&dst)[4 * NV]) {
      if (flat >= total) return;
      const long long ti = flat / nkb;
      const int kb = (int)(flat - ti * nkb);
      const long long mb = (blockIdx.x + ti * gridDim.x) / nblk;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long m = mb * GROWS + mrow[i];
        const int k = kb * KB + kofs[i] * NV;
#pragma unroll
        for (int h = 0; h < NV; ++h)
          dst[NV * i + h] = (m < g.M && k + 4 * h < g.K) ? __ldg(reinterpret_cast<const float4 *>(g.A + m * (long long)g.lda + k + 4 * h)) : z4;
      }
    };
    auto put_blk = [&](long long flat, const float4(&src)[4 * NV]) {
      const int s = (int)(flat % GNS);
      const uint32_t ph = (uint32_t)(flat / GNS) & 1u;
      wait_spin(&empty[s], ph ^ 1u);
      uint8_t *st = smem + s * SB;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // hi = a with the 13 low mantissa bits cleared (what the tf32 datapath keeps; exactly an fp16 inside its normal
        // range), lo = a - hi exactly.  Two instructions per element: cvt.rna.tf32 lowers to ~7 (ncu: the A path, not the
        // tensor pipe, bounded the first version), and the dropped terms stay <= 2^-20 |a b|.
        if (F16) {
          const float a[8] = {src[2 * i].x, src[2 * i].y, src[2 * i].z, src[2 * i].w, src[2 * i + 1].x, src[2 * i + 1].y, src[2 * i + 1].z, src[2 * i + 1].w};
          uint32_t hp[4], lp[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float h0 = trunc_tf32(a[2 * j]), h1 = trunc_tf32(a[2 * j + 1]);
            amax = fmaxf(amax, fmaxf(fabsf(a[2 * j]), fabsf(a[2 * j + 1])));
            hp[j] = pack_h2g(h0, h1);
            lp[j] = pack_h2g(a[2 * j] - h0, a[2 * j + 1] - h1);
          }
          *reinterpret_cast<uint4 *>(st + sofs[i]) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
          *reinterpret_cast<uint4 *>(st + A_IMG + sofs[i]) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
        } else {
          float4 hi, lo;
          hi.x = trunc_tf32(src[i].x); lo.x = src[i].x - hi.x;
          hi.y = trunc_tf32(src[i].y); lo.y = src[i].y - hi.y;
          hi.z = trunc_tf32(src[i].z); lo.z = src[i].z - hi.z;
          hi.w = trunc_tf32(src[i].w); lo.w = src[i].w - hi.w;
          *reinterpret_cast<float4 *>(st + sofs[i]) = hi;
          *reinterpret_cast<float4 *>(st + A_IMG + sofs[i]) = lo;
        }
      }
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&full_a[s]);
    };
    if (F16) {  // two k-blocks in flight (a k-block is 32 floats per thread here)
      float4 b0[4 * NV], b1[4 * NV];
      load_blk(0, b0);
      load_blk(1, b1);
      for (long long flat = 0; flat < total; flat += 2) {
        put_blk(flat, b0);
        load_blk(flat + 2, b0);
        if (flat + 1 < total) {
          put_blk(flat + 1, b1);
          load_blk(flat + 3, b1);
        }
      }
    } else {
      float4 b0[4 * NV], b1[4 * NV], b2[4 * NV];
      load_blk(0, b0);
      load_blk(1, b1);
      load_blk(2, b2);
      for (long long flat = 0; flat < total; flat += 3) {
        put_blk(flat, b0);
        load_blk(flat + 3, b0);
        if (flat + 1 < total) {
          put_blk(flat + 1, b1);
          load_blk(flat + 4, b1);
        }
        if (flat + 2 < total) {
          put_blk(flat + 2, b2);
          load_blk(flat + 5, b2);
        }
      }
    }
    // an activation left the fp16 range (or was not finite): sticky bit in the context's status word; the host then uses
    // the tf32 images of the weights (kept beside the fp16 ones) for every later GEMM of the context
    if (F16 && !(amax < 60000.f) && tb.status) atomicOr_system(tb.status, l2hmc::STATUS_F16_RANGE);
  } else if (warp < 16) {
    // ---------------- epilogue warps: thread = output row of accumulator (e / 4), TMEM lane group (e % 4) ----------
    // (APRE: warps 0-7 take the even 16-column chunks, warps 8-15 the odd ones)
    const int e = warp & 7;
    const int lg = e & 3, half = e >> 2;
    const int cpart = APRE ? (warp >> 3) : 0, cstep = APRE ? 32 : 16;
    float omax = 0.f;  // largest |value| written to the operand image (fp16 range guard of the consumer)
    for (long long ti = 0; ti < my_tiles; ++ti) {
      const long long tile = blockIdx.x + ti * gridDim.x;
      const long long mb = tile / nblk;
      const int nb = (int)(tile - mb * nblk);
      const int n0 = nb * BN;
      wait_spin(acc_full, (uint32_t)ti & 1u);
      tc::tcgen05_fence_after();
      const long long m = mb * GROWS + half * GM + lg * 32 + lane;
      const bool mok = m < g.M;
      const float *bias = (EPI == layered::EPI_RELU && g.dir != nullptr && mok && g.dir[m] == 0) ? g.bias_b : g.bias;
      const uint32_t trow = tmem_base + (uint32_t)(half * 256) + ((uint32_t)(32 * lg) << 16);
      for (int c = cpart * 16; c < BN; c += cstep) {
        float acc[16];
        tc::tmem_ld16(trow + (uint32_t)c, acc);
        tc::tmem_wait_ld();
        const int nbase = n0 + c;
        if (!mok || nbase >= g.N) continue;
        epi_chunk<EPI>(g, acc, m, nbase, bias, omax);
      }
      tc::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(acc_empty);
    }
    if (g.c_img != nullptr && !(omax < 60000.f) && tb.status) atomicOr_system(tb.status, l2hmc::STATUS_F16_RANGE);
  } else if (warp == W_TMA) {
    // ---------------- B path (TMA bulk copies of the packed weight image) ----------------
    if (lane == 0) {
      const uint32_t bytes = 2 * B_HALF;
      long long flat = 0;
      for (long long ti = 0; ti < my_tiles; ++ti) {
        const long long tile = blockIdx.x + ti * gridDim.x;
        const int nb = (int)(tile % nblk);
        const long long mb = tile / nblk;
        const float *src = tb.pk + ((size_t)nb * nkb) * b_block_floats(BN);
        for (int kb = 0; kb < nkb; ++kb, ++flat) {
          const int s = (int)(flat % GNS);
          const uint32_t ph = (uint32_t)(flat / GNS) & 1u;
          wait_spin(&empty[s], ph ^ 1u);
          tc::mbar_arrive_expect_tx(&full_b[s], bytes + (APRE ? A_ALL : 0u));
          if (APRE)  // blocks (2 mb, kb) and (2 mb + 1, kb) of the image are adjacent: [A0 hi | A0 lo | A1 hi | A1 lo]
            tc::bulk_g2s(smem + s * SB, g.a_img + ((size_t)kb * g.img_nmb + (size_t)(2 * mb)) * layered::SplitImage::BLOCK_BYTES, A_ALL,
                         &full_b[s]);
          tc::bulk_g2s(smem + s * SB + A_ALL, src + (size_t)kb * b_block_floats(BN), bytes, &full_b[s]);
        }
      }
    }
  } else {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc = F16 ? make_idesc_f16g(GM, BN) : tc::make_idesc_tf32(GM, BN);
      const uint32_t lbo_a = (GM / 8) * 128, lbo_b = (uint32_t)(BN / 8) * 128;
      long long flat = 0;
      for (long long ti = 0; ti < my_tiles; ++ti) {
        wait_spin(acc_empty, ((uint32_t)ti & 1u) ^ 1u);
        tc::tcgen05_fence_after();
        for (int kb = 0; kb < nkb; ++kb, ++flat) {
          const int s = (int)(flat % GNS);
          const uint32_t ph = (uint32_t)(flat / GNS) & 1u;
          if (!APRE) wait_spin(&full_a[s], ph);
          wait_spin(&full_b[s], ph);
          tc::tcgen05_fence_after();
          const uint32_t sa = tc::smem_u32(smem + s * SB), sb = sa + A_ALL;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t dacc = tmem_base + (uint32_t)(h * 256);
            const uint32_t sah = sa + h * 2 * A_IMG;
#pragma unroll
            for (int ks = 0; ks < GBK / 8; ++ks) {
              const uint64_t a_hi = tc::make_smem_desc(sah + ks * 2 * lbo_a, lbo_a, 128);
              const uint64_t a_lo = tc::make_smem_desc(sah + A_IMG + ks * 2 * lbo_a, lbo_a, 128);
              const uint64_t b_hi = tc::make_smem_desc(sb + ks * 2 * lbo_b, lbo_b, 128);
              const uint64_t b_lo = tc::make_smem_desc(sb + B_HALF + ks * 2 * lbo_b, lbo_b, 128);
              if (F16) {
                mma_f16_ss(dacc, a_lo, b_hi, idesc, (kb | ks) != 0);  // small terms first
                mma_f16_ss(dacc, a_hi, b_lo, idesc, true);
                mma_f16_ss(dacc, a_hi, b_hi, idesc, true);
              } else {
                mma_tf32_ss(dacc, a_lo, b_hi, idesc, (kb | ks) != 0);  // small terms first
                mma_tf32_ss(dacc, a_hi, b_lo, idesc, true);
                mma_tf32_ss(dacc, a_hi, b_hi, idesc, true);
              }
            }
          }
          tc::tcgen05_commit(&empty[s]);
        }
        tc::tcgen05_commit(acc_full);
      }
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc::tcgen05_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// tc_gemm_pre_kernel: the GEMM for a pre-split A operand (layered::SplitImage), fp16 x3, with the epilogue off the critical
// path.  Persistent CTAs walk 128 x BN tiles; TMEM holds TWO accumulators of BN <= 256 columns, so the 16 epilogue warps
// drain tile i (fused layer tail, fp32 C and / or the next operand image) while the tensor pipe already accumulates tile
// i + 1 -- in the 256-row kernel above the pipe idles for the whole epilogue (ncu: 50 % tensor-active on the decoder GEMMs).
//   warps 0-15  epilogue: TMEM lane group warp % 4, 16-column chunks (warp / 4), (warp / 4) + 4, ...
//   warp 16     TMA: per k-block of 32 one 16 KB copy of the A image block and one of the packed B block
//   warp 17     MMA issuer (SS form), 6 MMAs (2 K=16 steps x 3 products) per stage
// 4-stage ring of 16 KB + 2 * BN * 64 B; mbarriers full / empty per stage, acc_full / acc_empty per accumulator.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int P_NS = 4;
__host__ __device__ inline size_t pre_stage_bytes(int BN) { return (size_t)2 * GM * GBK * 4 + b_block_floats(BN) * 4; }
__host__ __device__ inline size_t tc_gemm_pre_smem(int BN) { return P_NS * pre_stage_bytes(BN) + 1024 + 256; }

template <int EPI>
__global__ void __launch_bounds__(G_THREADS, 1) tc_gemm_pre_kernel(const layered::GemmArgs g, const TcGemmB tb) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int BN = tb.BN;
  const size_t SB = pre_stage_bytes(BN);
  const uint32_t A_IMG = GM * GBK * 4;  // 8 KB: hi or lo of one 128 x 32 block; stage: [A hi | A lo | B hi | B lo]
  const uint32_t A_ALL = 2 * A_IMG;
  const uint32_t B_HALF = (uint32_t)BN * GBK * 4;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + P_NS * SB);
  uint64_t *full = bars, *empty = bars + P_NS, *acc_full = bars + 2 * P_NS, *acc_empty = bars + 2 * P_NS + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * P_NS + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = tb.nkb, nblk = tb.nblk;
  const long long mblocks = (g.M + GM - 1) / GM;
  const long long tiles = mblocks * nblk;  // n-block fastest: neighbouring CTAs read the same A blocks through L2
  const long long my_tiles = (tiles > blockIdx.x) ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < P_NS; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&acc_full[b], 1);
      tc::mbar_init(&acc_empty[b], 16);
    }
    tc::fence_mbar_init();
  }
  if (warp == W_MMA) tc::tmem_alloc(tmem_slot, 512);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 16) {
    const int lg = warp & 3, cpart = warp >> 2;
    float omax = 0.f;
    for (long long ti = 0; ti < my_tiles; ++ti) {
      const long long tile = blockIdx.x + ti * gridDim.x;
      const long long mb = tile / nblk;
      const int nb = (int)(tile - mb * nblk);
      const int n0 = nb * BN;
      const int buf = (int)(ti & 1);
      wait_spin(&acc_full[buf], (uint32_t)(ti >> 1) & 1u);
      tc::tcgen05_fence_after();
      const long long m = mb * GM + lg * 32 + lane;
      const bool mok = m < g.M;
      const float *bias = (EPI == layered::EPI_RELU && g.dir != nullptr && mok && g.dir[m] == 0) ? g.bias_b : g.bias;
      const uint32_t trow = tmem_base + (uint32_t)(buf * 256) + ((uint32_t)(32 * lg) << 16);
      for (int c = cpart * 16; c < BN; c += 64) {
        float acc[16];
        tc::tmem_ld16(trow + (uint32_t)c, acc);
        tc::tmem_wait_ld();
        const int nbase = n0 + c;
        if (!mok || nbase >= g.N) continue;
        epi_chunk<EPI>(g, acc, m, nbase, bias, omax);
      }
      tc::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);
    }
    if (g.c_img != nullptr && !(omax < 60000.f) && tb.status) atomicOr_system(tb.status, l2hmc::STATUS_F16_RANGE);
  } else if (warp == W_TMA) {
    if (lane == 0) {
      const uint32_t bytes_b = 2 * B_HALF;
      long long flat = 0;
      for (long long ti = 0; ti < my_tiles; ++ti) {
        const long long tile = blockIdx.x + ti * gridDim.x;
        const long long mb = tile / nblk;
        const int nb = (int)(tile - mb * nblk);
        const float *src = tb.pk + ((size_t)nb * nkb) * b_block_floats(BN);
        for (int kb = 0; kb < nkb; ++kb, ++flat) {
          const int s = (int)(flat % P_NS);
          const uint32_t ph = (uint32_t)(flat / P_NS) & 1u;
          wait_spin(&empty[s], ph ^ 1u);
          tc::mbar_arrive_expect_tx(&full[s], bytes_b + A_ALL);
          tc::bulk_g2s(smem + s * SB, g.a_img + ((size_t)kb * g.img_nmb + (size_t)mb) * layered::SplitImage::BLOCK_BYTES, A_ALL, &full[s]);
          tc::bulk_g2s(smem + s * SB + A_ALL, src + (size_t)kb * b_block_floats(BN), bytes_b, &full[s]);
        }
      }
    }
  } else {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16g(GM, BN);
      const uint32_t lbo_a = (GM / 8) * 128, lbo_b = (uint32_t)(BN / 8) * 128;
      long long flat = 0;
      for (long long ti = 0; ti < my_tiles; ++ti) {
        const int buf = (int)(ti & 1);
        wait_spin(&acc_empty[buf], ((uint32_t)(ti >> 1) & 1u) ^ 1u);
        tc::tcgen05_fence_after();
        const uint32_t dacc = tmem_base + (uint32_t)(buf * 256);
        for (int kb = 0; kb < nkb; ++kb, ++flat) {
          const int s = (int)(flat % P_NS);
          const uint32_t ph = (uint32_t)(flat / P_NS) & 1u;
          wait_spin(&full[s], ph);
          tc::tcgen05_fence_after();
          const uint32_t sa = tc::smem_u32(smem + s * SB), sb = sa + A_ALL;
#pragma unroll
          for (int ks = 0; ks < GBK / 8; ++ks) {
            const uint64_t a_hi = tc::make_smem_desc(sa + ks * 2 * lbo_a, lbo_a, 128);
            const uint64_t a_lo = tc::make_smem_desc(sa + A_IMG + ks * 2 * lbo_a, lbo_a, 128);
            const uint64_t b_hi = tc::make_smem_desc(sb + ks * 2 * lbo_b, lbo_b, 128);
            const uint64_t b_lo = tc::make_smem_desc(sb + B_HALF + ks * 2 * lbo_b, lbo_b, 128);
            mma_f16_ss(dacc, a_lo, b_hi, idesc, (kb | ks) != 0);  // small terms first, as in tc_gemm_kernel
            mma_f16_ss(dacc, a_hi, b_lo, idesc, true);
            mma_f16_ss(dacc, a_hi, b_hi, idesc, true);
          }
          tc::tcgen05_commit(&empty[s]);
        }
        tc::tcgen05_commit(&acc_full[buf]);
      }
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc::tcgen05_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

inline cudaError_t launch_tc_gemm_pre(const layered::GemmArgs &g, const TcGemmB &tb, int sms, cudaStream_t s) {
  if (!tb.f16 || !g.a_img) return cudaErrorInvalidValue;
  const size_t smem = tc_gemm_pre_smem(tb.BN);
  const long long tiles = (long long)tb.nblk * ((g.M + GM - 1) / GM);
  const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
  cudaError_t e = cudaSuccess;
#define L2HMC_TCP_LAUNCH(E)                                                                                              \
  case E: {                                                                                                              \
    static size_t configured_dev[64] = {0};                                                                              \
    int dev_ = 0;                                                                                                        \
    cudaGetDevice(&dev_);                                                                                                \
    size_t &configured = configured_dev[dev_ & 63];                                                                      \
    if (smem > configured) {                                                                                             \
      e = cudaFuncSetAttribute(tc_gemm_pre_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
      if (e != cudaSuccess) return e;                                                                                    \
      configured = smem;                                                                                                 \
    }                                                                                                                    \
    tc_gemm_pre_kernel<E><<<grid, G_THREADS, smem, s>>>(g, tb);                                                          \
  } break;
  switch (g.epi) {
    L2HMC_TCP_LAUNCH(layered::EPI_BIAS)
    L2HMC_TCP_LAUNCH(layered::EPI_RELU)
    L2HMC_TCP_LAUNCH(layered::EPI_SOFTPLUS)
    L2HMC_TCP_LAUNCH(layered::EPI_DSOFTPLUS)
    L2HMC_TCP_LAUNCH(layered::EPI_ADD_SCALE)
    default: return cudaErrorInvalidValue;
  }
#undef L2HMC_TCP_LAUNCH
  return cudaGetLastError();
}

// Launch helper: picks the epilogue instantiation; grid = min(tiles, SMs).
inline cudaError_t launch_tc_gemm(const layered::GemmArgs &g, const TcGemmB &tb, int sms, cudaStream_t s) {
  if ((g.a_img || g.c_img) && !tb.f16) return cudaErrorInvalidValue;  // operand images are fp16 pairs
  const size_t smem = tc_gemm_smem(tb.BN);
  const long long tiles = (long long)tb.nblk * ((g.M + GROWS - 1) / GROWS);
  const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
  cudaError_t e = cudaSuccess;
#define L2HMC_TCG_LAUNCH(E)                                                                                              \
  case E: {                                                                                                              \
    static size_t configured_dev[64] = {0};                                                                              \
    int dev_ = 0;                                                                                                        \
    cudaGetDevice(&dev_);                                                                                                \
    size_t &configured = configured_dev[dev_ & 63];                                                                      \
    if (smem > configured) {                                                                                             \
      e = cudaFuncSetAttribute(tc_gemm_kernel<E, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
      if (e != cudaSuccess) return e;                                                                                    \
      e = cudaFuncSetAttribute(tc_gemm_kernel<E, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
      if (e != cudaSuccess) return e;                                                                                    \
      e = cudaFuncSetAttribute(tc_gemm_kernel<E, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
      if (e != cudaSuccess) return e;                                                                                    \
      configured = smem;                                                                                                 \
    }                                                                                                                    \
    if (tb.f16 && g.a_img) tc_gemm_kernel<E, true, true><<<grid, G_THREADS, smem, s>>>(g, tb);                           \
    else if (tb.f16) tc_gemm_kernel<E, true><<<grid, G_THREADS, smem, s>>>(g, tb);                                       \
    else tc_gemm_kernel<E, false><<<grid, G_THREADS, smem, s>>>(g, tb);                                                  \
  } break;
  switch (g.epi) {
    L2HMC_TCG_LAUNCH(layered::EPI_BIAS)
    L2HMC_TCG_LAUNCH(layered::EPI_RELU)
    L2HMC_TCG_LAUNCH(layered::EPI_SOFTPLUS)
    L2HMC_TCG_LAUNCH(layered::EPI_DSOFTPLUS)
    L2HMC_TCG_LAUNCH(layered::EPI_ADD_SCALE)
    default: return cudaErrorInvalidValue;
  }
#undef L2HMC_TCG_LAUNCH
  return cudaGetLastError();
}

// ---- host side: pre-split, pre-tiled weight image ---------------------------------------------------------------
inline float tf32_rna_h(float x) {  // cvt.rna.tf32.f32: round to nearest, ties away from zero
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return x;
  u += 0x1000u;
  u &= 0xFFFFE000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

inline void choose_bn(int N, int *BN, int *nblk) {
  int nb = (N + 255) / 256;
  int bn = ((N + nb - 1) / nb + 15) / 16 * 16;
  *BN = bn;
  *nblk = nb;
}

// B [K][ldb] row-major (columns [0, N)) -> packed image; returns floats written. K rows beyond `K` are zero.
// f16: fp16 hi / lo, k-block = 32, core-matrix rows of 8 halfs (the image of a k-block has the same size in bytes).
inline size_t pack_b(const float *B, int ldb, int K, int N, std::vector<float> &out, TcGemmB *desc, bool f16 = false) {
  int BN, nblk;
  choose_bn(N, &BN, &nblk);
  const int kbe = f16 ? 2 * GBK : GBK;
  const int nkb = (K + kbe - 1) / kbe;
  const size_t blk = b_block_floats(BN);
  const size_t base = out.size();
  out.resize(base + (size_t)nblk * nkb * blk, 0.f);
  float *o = out.data() + base;
  for (int nb = 0; nb < nblk; ++nb)
    for (int kb = 0; kb < nkb; ++kb) {
      float *hi = o + ((size_t)nb * nkb + kb) * blk, *lo = hi + (size_t)GBK * BN;
      if (f16) {
        uint16_t *hh = reinterpret_cast<uint16_t *>(hi), *lh = reinterpret_cast<uint16_t *>(lo);
        for (int kc = 0; kc < 4; ++kc)
          for (int ng = 0; ng < BN / 8; ++ng)
            for (int r = 0; r < 8; ++r)
              for (int e = 0; e < 8; ++e) {
                const int n = nb * BN + ng * 8 + r, k = kb * kbe + kc * 8 + e;
                const float w = (n < N && k < K) ? B[(size_t)k * ldb + n] : 0.f;
                const __half h = __float2half_rn(w);
                const __half l = __float2half_rn(w - __half2float(h));
                const size_t idx = ((size_t)(kc * (BN / 8) + ng) * 8 + r) * 8 + e;
                memcpy(&hh[idx], &h, 2);
                memcpy(&lh[idx], &l, 2);
              }
        continue;
      }
      for (int kc = 0; kc < GBK / 4; ++kc)
        for (int ng = 0; ng < BN / 8; ++ng)
          for (int r = 0; r < 8; ++r)
            for (int e = 0; e < 4; ++e) {
              const int n = nb * BN + ng * 8 + r, k = kb * GBK + kc * 4 + e;
              const float w = (n < N && k < K) ? B[(size_t)k * ldb + n] : 0.f;
              const float h = tf32_rna_h(w);
              const size_t idx = ((size_t)(kc * (BN / 8) + ng) * 8 + r) * 4 + e;
              hi[idx] = h;
              lo[idx] = tf32_rna_h(w - h);
            }
    }
  desc->pk = nullptr;  // caller sets the device pointer
  desc->status = nullptr;
  desc->BN = BN;
  desc->nblk = nblk;
  desc->nkb = nkb;
  desc->f16 = f16 ? 1 : 0;
  return out.size() - base;
}

}  // namespace tcg
}  // namespace l2hmc
