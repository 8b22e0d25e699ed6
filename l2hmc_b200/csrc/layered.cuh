// Layered engine of libl2hmc.so: the same transition as the fused kernels, run as a sequence of batched
// launches over all chains.  It exists for the shapes that do not fit on one SM -- the MNIST-VAE posterior
// target of the reference (mnist_vae.py:104-178: a 50 -> 1024 -> 1024 -> 784 softplus decoder inside the energy,
// width-200 S/T/Q nets conditioned on a 784 -> 512 -> 512 -> 200 encoding of the image) and any x_dim / width
// beyond the tile kernel's -- where the dominant cost is genuinely dense [chains, K] x [K, N] products.
//
//   sgemm_kernel     C = epi(A B + bias [+ R]) : register-tiled fp32 FMA GEMM (128 x 128|64 x 8 tiles, 8 x 8|4 per
//                    thread, double-buffered shared memory), with the layer's elementwise tail fused in:
//                    bias, per-direction time-embedding bias, relu, softplus, softplus' (.) backward product,
//                    "+ z" and 1/temperature of the decoder gradient.
//   k_lay_*          per-chain elementwise kernels (one warp per chain): momentum draw, the v / masked-x
//                    updates of utils/dynamics.py:115-201 with the log|J| row sums, Bernoulli log-likelihood
//                    and its logit gradient, Hamiltonians + accept.
//
// State lives in HBM between launches ([n, Dp] x, v; [n, K1p] net input; [n, Hp] activations), every buffer
// row-padded to a multiple of 8 floats with the pad columns held at zero so each GEMM runs without K tails.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace l2hmc {
namespace layered {

enum { EPI_BIAS = 0, EPI_RELU = 1, EPI_SOFTPLUS = 2, EPI_DSOFTPLUS = 3, EPI_ADD_SCALE = 4 };

struct GemmArgs {
  const float *A; int lda;      // [M][lda]; columns [0, K) are read, K % 8 == 0 (zero padded)
  const float *B; int ldb;      // [K][ldb] row-major; columns [0, Bn) readable, Bn % 4 == 0
  int Bn;
  float *C; int ldc;            // [M][ldc]; columns [0, N) are written
  long long M;
  int N, K;
  const float *bias;            // [>= N] or null
  const float *bias_b;          // bias for rows whose dir bit is 0 (time embedding of a backward chain)
  const int *dir;               // [M] or null
  const float *R; int ldr;      // optional row term [M][ldr]
  float scale;
  int epi;
  int vec;                      // 1: C rows 16-byte aligned and N % 4 == 0 -> float4 stores
  // tensor-core engine only (tc_gemm.cuh), all optional:
  const uint8_t *a_img = nullptr;  // A as a pre-split fp16 hi / lo operand image (SplitImage layout) instead of fp32 rows
  uint8_t *c_img = nullptr;        // also write epi(...) as the operand image the NEXT GEMM reads (its K = this N)
  int img_nmb = 0;                 // 128-row blocks of either image (both have M rows): 2 * ceil(M / 256)
  int no_c = 0;                    // 1: do not store the fp32 C (only c_img is wanted); DSOFTPLUS still reads C
};

// Pre-split operand image of an activation matrix X [M][K] for the fp16 x3 tensor-core GEMM: what tc_gemm_kernel's A path
// would put into shared memory, kept in HBM so that the consumer fetches it with bulk (TMA) copies and nobody converts an
// element more than once.  Block (mb, kb) = rows [128 mb, 128 mb + 128) x columns [32 kb, 32 kb + 32), 16 KB:
// [hi | lo] x [kc = 0..3][row group = 0..15][row in group = 0..7][8 halfs], the UMMA K-major no-swizzle core-matrix order;
// blocks are stored kb-major ((kb * nmb + mb) * 16 KB) so the two row halves of a 256-row tile are one 32 KB copy.
// Rows >= M and columns >= K of the image stay zero (the buffer is zero-initialised and only valid elements are written).
struct SplitImage {
  static constexpr int BLOCK_BYTES = 16384;
  __host__ __device__ static int nmb(long long M) { return (int)(2 * ((M + 255) / 256)); }
  __host__ __device__ static int nkb(int K) { return (K + 31) / 32; }
  __host__ __device__ static size_t bytes(long long M, int K) { return (size_t)nmb(M) * nkb(K) * BLOCK_BYTES; }
  // byte offset of the 16-byte piece holding columns [8 * (k / 8), +8) of row m in the hi part (lo: + 8192)
  __host__ __device__ static size_t piece(long long m, int k, int nmb_) {
    return ((size_t)(k >> 5) * nmb_ + (size_t)(m >> 7)) * BLOCK_BYTES + (size_t)((k & 31) >> 3) * 2048 + (size_t)((m & 127) >> 3) * 128 +
           (size_t)(m & 7) * 16;
  }
};

// fp32 -> (hi, lo) halfs exactly as tc_gemm_kernel's A path does: hi keeps the top 10 mantissa bits, lo = a - hi.
__device__ __forceinline__ void split8_to_half(const float (&a)[8], uint4 &hi, uint4 &lo, float &amax) {
  uint32_t hp[4], lp[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float h0 = __uint_as_float(__float_as_uint(a[2 * j]) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(a[2 * j + 1]) & 0xFFFFE000u);
    amax = fmaxf(amax, fmaxf(fabsf(a[2 * j]), fabsf(a[2 * j + 1])));
    const __half2 hh = __floats2half2_rn(h0, h1), ll = __floats2half2_rn(a[2 * j] - h0, a[2 * j + 1] - h1);
    hp[j] = *reinterpret_cast<const uint32_t *>(&hh);
    lp[j] = *reinterpret_cast<const uint32_t *>(&ll);
  }
  hi = make_uint4(hp[0], hp[1], hp[2], hp[3]);
  lo = make_uint4(lp[0], lp[1], lp[2], lp[3]);
}

__device__ __forceinline__ float softplus_f(float t) {  // tf.nn.softplus, overflow-free
  return fmaxf(t, 0.f) + log1pf(expf(-fabsf(t)));
}

// log(1 + e) for e in [0, 1]: e * p9(e), Chebyshev fit of log1p(e)/e (max relative error 1.8e-7 in fp32 Horner form);
// 10 FMAs where log1pf costs ~40 instructions -- the softplus / Bernoulli tails are instruction-bound.
__device__ __forceinline__ float log1p_unit(float e) {
  float p = -3.176057013e-03f;
  p = fmaf(p, e, 1.954252645e-02f);
  p = fmaf(p, e, -5.637361109e-02f);
  p = fmaf(p, e, 1.054362357e-01f);
  p = fmaf(p, e, -1.526966691e-01f);
  p = fmaf(p, e, 1.966327429e-01f);
  p = fmaf(p, e, -2.495161593e-01f);
  p = fmaf(p, e, 3.332971036e-01f);
  p = fmaf(p, e, -4.999989271e-01f);
  p = fmaf(p, e, 1.0f);
  return p * e;
}

template <int TN>
__global__ void __launch_bounds__(256, 2) sgemm_kernel(const GemmArgs g) {
  constexpr int BM = 128, BN = 16 * TN, BK = 8, LDA_S = BM + 4, BV = BN / 4;
  __shared__ __align__(16) float As[2][BK * LDA_S];
  __shared__ __align__(16) float Bs[2][BK * BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * BN;
  const long long m0 = (long long)blockIdx.y * BM;

  const int arow = tid >> 1, ak = (tid & 1) * 4;
  const bool a_ok = (m0 + arow) < g.M;
  const float *Ap = g.A + (m0 + arow) * (long long)g.lda + ak;
  const int brow = tid / BV, bc = (tid % BV) * 4;
  const bool b_ok = (tid < BK * BV) && (n0 + bc) < g.Bn;
  const float *Bp = g.B + (long long)brow * g.ldb + n0 + bc;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra = a_ok ? __ldg(reinterpret_cast<const float4 *>(Ap)) : zero4;
  float4 rb = b_ok ? __ldg(reinterpret_cast<const float4 *>(Bp)) : zero4;
  auto stage = [&](int buf) {
    As[buf][(ak + 0) * LDA_S + arow] = ra.x;
    As[buf][(ak + 1) * LDA_S + arow] = ra.y;
    As[buf][(ak + 2) * LDA_S + arow] = ra.z;
    As[buf][(ak + 3) * LDA_S + arow] = ra.w;
    if (tid < BK * BV) *reinterpret_cast<float4 *>(&Bs[buf][brow * BN + bc]) = rb;
  };
  stage(0);
  __syncthreads();

  const int nk = g.K / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      ra = a_ok ? __ldg(reinterpret_cast<const float4 *>(Ap + (kt + 1) * BK)) : zero4;
      rb = b_ok ? __ldg(reinterpret_cast<const float4 *>(Bp + (long long)(kt + 1) * BK * g.ldb)) : zero4;
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4 *>(&As[cur][k * LDA_S + ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4 *>(&As[cur][k * LDA_S + 64 + ty * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[TN];
      {
        const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[cur][k * BN + tx * 4]);
        bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
        if (TN == 8) {
          const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[cur][k * BN + 64 + tx * 4]);
          bv[TN - 4] = b1.x; bv[TN - 3] = b1.y; bv[TN - 2] = b1.z; bv[TN - 1] = b1.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) stage(cur ^ 1);
    __syncthreads();
  }

  // ---- fused epilogue --------------------------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
    const float *bias = (g.dir != nullptr && g.dir[m] == 0) ? g.bias_b : g.bias;
#pragma unroll
    for (int jj = 0; jj < TN / 4; ++jj) {
      const int n = n0 + jj * 64 + tx * 4;
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int nn = n + j;
        float v = acc[i][jj * 4 + j];
        if (nn < g.N) {
          if (g.epi == EPI_DSOFTPLUS) {
            const float h = g.C[m * g.ldc + nn];   // softplus'(pre) = sigmoid(pre) = 1 - exp(-softplus(pre))
            v = v * (-expm1f(-h));
          } else if (g.epi == EPI_ADD_SCALE) {
            v = (v + g.R[m * g.ldr + nn]) * g.scale;
          } else {
            if (bias) v += bias[nn];
            if (g.R) v += g.R[m * g.ldr + nn];
            if (g.epi == EPI_RELU) v = fmaxf(v, 0.f);
            else if (g.epi == EPI_SOFTPLUS) v = softplus_f(v);
          }
        }
        o[j] = v;
      }
      if (g.vec && n + 3 < g.N) {
        *reinterpret_cast<float4 *>(&g.C[m * g.ldc + n]) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < g.N) g.C[m * g.ldc + n + j] = o[j];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Elementwise kernels: one warp per chain
// ---------------------------------------------------------------------------------------------
struct LayDims {
  int D, Dp;     // x_dim, rounded up to 8
  int K1p;       // round_up(2 D, 8): net input [a | b | 0]
  int H, Hp;     // width, rounded up to 8
  int N3p;       // round_up(3 D, 8): head output [S | T | Q | 0]
  int T;
  int aux, auxp; // aux_dim, rounded up to 8 (0 when the target takes no aux)
  int ldm;       // row stride of the mask table
};

struct LayState {
  float *x, *v, *x0;   // [n][Dp]
  float *ab;           // [n][K1p]
  float *hd;           // [n][N3p]
  float *logj, *h0, *U, *u;  // [n]
  int *dir;            // [n]
  int *acc;            // [n]
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Per-transition setup: x (first transition: from the caller; later ones: already selected in place), momentum,
// direction bit, accept uniform, kinetic part of H(x0, v0).  Writes ab[:, :D] = x for the first VNet call.
__global__ void k_lay_begin(LayDims dm, LayState st, TransitionIO io, int tr) {
  const long long n = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= io.n) return;
  const int D = dm.D;
  const unsigned long long ctr = io.counter + (unsigned long long)tr;
  float kin = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float xv = (tr == 0) ? io.x[n * D + d] : st.x[n * dm.Dp + d];
    st.x[n * dm.Dp + d] = xv;
    st.x0[n * dm.Dp + d] = xv;
    st.ab[n * dm.K1p + d] = xv;
  }
  if (io.v != nullptr) {
    for (int d = lane; d < D; d += 32) {
      const float vv = io.v[((long long)tr * io.n + n) * D + d];
      st.v[n * dm.Dp + d] = vv;
      kin = fmaf(vv, vv, kin);
    }
  } else {
    for (int b = lane; 4 * b < D; b += 32) {
      float z[4];
      philox_normals4(io.seed, ctr, io.chain_offset + n, b, z);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (4 * b + q < D) {
          st.v[n * dm.Dp + 4 * b + q] = z[q];
          kin = fmaf(z[q], z[q], kin);
        }
    }
  }
  kin = warp_sum(kin);
  if (lane == 0) {
    int pd = 1;
    float pu = 0.f;
    if (io.dir_mode == 3 || (io.do_mh && io.u == nullptr)) philox_dir_u(io.seed, ctr, io.chain_offset + n, pd, pu);
    int dbit = 1;
    if (io.dir_mode == 1) dbit = 0;
    else if (io.dir_mode == 2) dbit = io.dir[(long long)tr * io.n + n] != 0;
    else if (io.dir_mode == 3) dbit = pd;
    st.dir[n] = dbit;
    if (io.do_mh && io.u != nullptr) pu = io.u[(long long)tr * io.n + n];
    st.u[n] = pu;
    st.logj[n] = 0.f;
    st.h0[n] = 0.5f * kin;
  }
}

// U(x), grad U(x) for the closed-form energies: one thread per chain.  g -> ab[:, D:2D].
__global__ void k_lay_grad_generic(LayDims dm, LayState st, EnergyDev en, Shape sh, long long n_chains, int want_grad) {
  const long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (n >= n_chains) return;
  const float *x = st.x + n * dm.Dp;
  st.U[n] = energy_chain(en, sh, x, 1);
  if (want_grad) grad_chain(en, sh, x, 1, st.ab + n * dm.K1p + dm.D, 1);
}

// Bernoulli decoder target (mnist_vae.py:122-126) given logits l = decoder(z):
//   U = like * sum_j [max(l,0) - l a + log(1 + exp(-|l|))] + 0.5 |z|^2, all / temperature;  l <- dU/dl = like * (sigmoid(l) - a).
// like = 1 for the sampler; the annealed energy of utils/ais.py:44-45 between the prior and this posterior is like = beta.
__global__ void k_lay_bce(LayDims dm, LayState st, float *logits, int ldl, const float *aux, float inv_temp, float like,
                          long long n_chains) {
  const long long n = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= n_chains) return;
  float s = 0.f;
  float *l = logits + n * ldl;
  const float *a = aux + n * dm.aux;
  for (int j = lane; j < dm.aux; j += 32) {
    const float lj = l[j], aj = a[j];
    const float e = expf(-fabsf(lj));
    s += fmaxf(lj, 0.f) - lj * aj + log1p_unit(e);
    const float r = __fdividef(1.f, 1.f + e);  // 1 + e in [1, 2]
    l[j] = like * (((lj >= 0.f) ? r : e * r) - aj);
  }
  float q = 0.f;
  for (int d = lane; d < dm.D; d += 32) {
    const float z = st.x[n * dm.Dp + d];
    q = fmaf(z, z, q);
  }
  s = warp_sum(s);
  q = warp_sum(q);
  if (lane == 0) st.U[n] = (like * s + 0.5f * q) * inv_temp;
}

// The same, with dU/dl written as the operand image of the first reverse GEMM (SplitImage; |dU/dl| <= like <= 1: no range
// guard) instead of in place.  A warp owns 8 consecutive chains: lane = (row r8 = lane & 7, column piece kc = lane >> 3), one
// iteration = a 32-column k-block, so its loads are 128 contiguous bytes per row and its image stores 128 contiguous bytes
// per piece (full 32-byte sectors -- 16-byte stores scattered over the image cost a sector fill each).
__global__ void __launch_bounds__(256) k_lay_bce_img(LayDims dm, LayState st, const float *logits, int ldl, const float *aux,
                                                     float inv_temp, float like, long long n_chains, uint8_t *img, int img_nmb) {
  const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, r8 = lane & 7, kc = lane >> 3;
  const long long n = w * 8 + r8;
  const bool ok = n < n_chains;
  const long long nn = ok ? n : 0;
  const float *l = logits + nn * ldl;
  const float *a = aux + nn * dm.aux;
  // 16-byte loads when the rows allow it (row strides multiples of 4 floats, 16-byte aligned bases)
  const bool vec = ((ldl | dm.aux) & 3) == 0 && ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(aux)) & 15) == 0;
  float s = 0.f, unused = 0.f;
  for (int j0 = 8 * kc; j0 < dm.aux; j0 += 32) {
    float l8[8], a8[8], d8[8];
    if (vec && j0 + 8 <= dm.aux) {
      const float4 l0 = __ldg(reinterpret_cast<const float4 *>(l + j0)), l1 = __ldg(reinterpret_cast<const float4 *>(l + j0 + 4));
      const float4 a0 = __ldg(reinterpret_cast<const float4 *>(a + j0)), a1 = __ldg(reinterpret_cast<const float4 *>(a + j0 + 4));
      l8[0] = l0.x; l8[1] = l0.y; l8[2] = l0.z; l8[3] = l0.w; l8[4] = l1.x; l8[5] = l1.y; l8[6] = l1.z; l8[7] = l1.w;
      a8[0] = a0.x; a8[1] = a0.y; a8[2] = a0.z; a8[3] = a0.w; a8[4] = a1.x; a8[5] = a1.y; a8[6] = a1.z; a8[7] = a1.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool in = j0 + j < dm.aux;
        l8[j] = in ? l[j0 + j] : 0.f;
        a8[j] = in ? a[j0 + j] : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float dj = 0.f;
      if (j0 + j < dm.aux) {
        const float lj = l8[j], aj = a8[j];
        const float e = expf(-fabsf(lj));
        s += fmaxf(lj, 0.f) - lj * aj + log1p_unit(e);
        const float r = __fdividef(1.f, 1.f + e);  // 1 + e in [1, 2]
        dj = like * (((lj >= 0.f) ? r : e * r) - aj);
      }
      d8[j] = dj;
    }
    if (ok) {
      uint4 hi, lo;
      split8_to_half(d8, hi, lo, unused);
      uint8_t *ip = img + SplitImage::piece(n, j0, img_nmb);
      *reinterpret_cast<uint4 *>(ip) = hi;
      *reinterpret_cast<uint4 *>(ip + 8192) = lo;
    }
  }
  float q = 0.f;
  for (int d = kc; d < dm.D; d += 4) {
    const float z = st.x[nn * dm.Dp + d];
    q = fmaf(z, z, q);
  }
  // the four column pieces of a row sit in lanes r8, r8 + 8, r8 + 16, r8 + 24
  s += __shfl_xor_sync(0xffffffffu, s, 8);
  s += __shfl_xor_sync(0xffffffffu, s, 16);
  q += __shfl_xor_sync(0xffffffffu, q, 8);
  q += __shfl_xor_sync(0xffffffffu, q, 16);
  if (ok && kc == 0) st.U[n] = (like * s + 0.5f * q) * inv_temp;
}

// fp32 rows [n][ld] (columns [0, K), K % 8 == 0, pad columns zero) -> operand image (SplitImage) for a GEMM whose A they are.
// Block = 128 rows x 2 column pieces at a time; a warp writes 512 contiguous bytes per piece.
__global__ void __launch_bounds__(256) k_lay_split(const float *src, int ld, int K, long long n, uint8_t *img, int img_nmb,
                                                   unsigned int *status) {
  const long long m = blockIdx.x * 128ll + (threadIdx.x & 127);
  if (m >= n) return;
  float amax = 0.f;
  const float *row = src + m * ld;
  for (int p = threadIdx.x >> 7; p < (K >> 3); p += 2) {
    const float4 v0 = *reinterpret_cast<const float4 *>(row + 8 * p), v1 = *reinterpret_cast<const float4 *>(row + 8 * p + 4);
    const float a8[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    uint4 hi, lo;
    split8_to_half(a8, hi, lo, amax);
    uint8_t *ip = img + SplitImage::piece(m, 8 * p, img_nmb);
    *reinterpret_cast<uint4 *>(ip) = hi;
    *reinterpret_cast<uint4 *>(ip + 8192) = lo;
  }
  if (!(amax < 60000.f) && status) atomicOr_system(status, l2hmc::STATUS_F16_RANGE);
}

__global__ void k_lay_add_h0(LayState st, long long n_chains) {
  const long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (n < n_chains) st.h0[n] += st.U[n];
}

// The fused state updates of utils/dynamics.py:115-201 given the raw head outputs hd = [S | T | Q] (pre-tanh).
// MODE 0: momentum half step (reads g from ab[:, D:2D]); build_next: write the XNet input [v_h | k (.) x].
// MODE 1: masked position update, half 0 / 1; writes the next net input's x part.
template <int MODE>
__global__ void k_lay_update(LayDims dm, LayState st, const float *es, const float *eq, const float *mask, float eps,
                             int it, int half, int build_next, int hmc, long long n_chains) {
  const long long n = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= n_chains) return;
  const int D = dm.D;
  const bool fwd = st.dir[n] != 0;
  const float *mrow = mask + (size_t)(fwd ? it : dm.T - 1 - it) * dm.ldm;
  const float *hd = st.hd + n * dm.N3p;
  float *ab = st.ab + n * dm.K1p;
  float lj = 0.f;
  for (int d = lane; d < D; d += 32) {
    float S = 0.f, Tt = 0.f, Q = 0.f;
    if (!hmc) {
      S = es[d] * tanhf(hd[d]);
      Tt = hd[D + d];
      Q = eq[d] * tanhf(hd[2 * D + d]);
    }
    const float m = mrow[d];
    if (MODE == 0) {
      float vv = st.v[n * dm.Dp + d];
      const float g = ab[D + d];
      const float sv = fwd ? (0.5f * eps) * S : (-0.5f * eps) * S;
      const float cterm = (0.5f * eps) * (-(expf(eps * Q) * g) + Tt);
      const float e = expf(sv);
      vv = fwd ? (vv * e + cterm) : ((vv - cterm) * e);
      st.v[n * dm.Dp + d] = vv;
      lj += sv;
      if (build_next) {
        const float k = fwd ? m : 1.f - m;  // keep mask of the first x update
        ab[d] = vv;
        ab[D + d] = k * st.x[n * dm.Dp + d];
      }
    } else {
      // fwd: first half keeps m, second keeps 1-m; bwd: first keeps 1-m, second keeps m
      const float k = (fwd == (half == 0)) ? m : 1.f - m;
      const float uu = 1.f - k;
      float xv = st.x[n * dm.Dp + d];
      const float vh = st.v[n * dm.Dp + d];
      const float sx = fwd ? eps * S : -eps * S;
      const float inner = eps * (expf(eps * Q) * vh + Tt);
      const float e = expf(sx);
      const float nx = fwd ? (xv * e + inner) : (e * (xv - inner));
      xv = k * xv + uu * nx;
      st.x[n * dm.Dp + d] = xv;
      lj += uu * sx;
      if (half == 0) ab[D + d] = uu * xv;  // second update keeps 1 - k
      else ab[d] = xv;                     // VNet([x_o, grad U(x_o)])
    }
  }
  lj = warp_sum(lj);
  if (lane == 0 && !hmc) st.logj[n] += lj;
}

// Hamiltonian difference, accept, outputs (utils/dynamics.py:302-309, utils/sampler.py:44-55).  U holds U(x1).
__global__ void k_lay_end(LayDims dm, LayState st, TransitionIO io, int tr) {
  const long long n = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= io.n) return;
  const int D = dm.D;
  const bool last = (tr == io.n_transitions - 1);
  float kin = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float vv = st.v[n * dm.Dp + d];
    kin = fmaf(vv, vv, kin);
  }
  kin = warp_sum(kin);
  const float logj = st.logj[n];
  const float p = accept_prob(st.h0[n], st.U[n] + 0.5f * kin, logj);
  const float px = io.log_jac ? logj : p;
  const int acc = io.do_mh ? ((px - st.u[n] >= 0.f) ? 1 : 0) : 0;
  if (lane == 0 && io.stats) {  // one warp per chain here: its lane 0 adds the chain (a launch of n atomics per transition)
    atomicAdd(io.stats, (double)px);
    atomicAdd(io.stats + 1, (double)acc);
  }
  if (io.trace)
    for (int d = lane; d < D; d += 32) io.trace[((long long)tr * io.n + n) * D + d] = acc ? st.x[n * dm.Dp + d] : st.x0[n * dm.Dp + d];
  if (last) {
    if (lane == 0) {
      io.px_out[n] = px;
      if (io.accepted) io.accepted[n] = (uint8_t)acc;
    }
    for (int d = lane; d < D; d += 32) {
      const float lx = st.x[n * dm.Dp + d];
      io.x_out[n * D + d] = lx;
      if (io.v_out) io.v_out[n * D + d] = st.v[n * dm.Dp + d];
      if (io.do_mh) io.x_next[n * D + d] = acc ? lx : st.x0[n * dm.Dp + d];
    }
  } else if (!acc) {
    for (int d = lane; d < D; d += 32) st.x[n * dm.Dp + d] = st.x0[n * dm.Dp + d];
  }
}

// Strided row copy: dst[r][0..w) = src[r][0..w) (other destination columns untouched).
__global__ void k_lay_copy_rows(const float *src, int lds, float *dst, int ldd, int w, long long n_rows) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_rows * w) return;
  const long long r = i / w;
  const int c = (int)(i - r * w);
  dst[r * ldd + c] = src[r * lds + c];
}

// H = U + 0.5 |v|^2 and the accept probability from precomputed energies (component calls on the decoder target).
__global__ void k_lay_hamiltonian(int D, long long n, const float *U, const float *v, float *out) {
  const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g >= n) return;
  float k = 0.f;
  for (int d = 0; d < D; ++d) k = fmaf(v[g * D + d], v[g * D + d], k);
  out[g] = U[g] + 0.5f * k;
}
__global__ void k_lay_p_accept(int D, long long n, const float *U0, const float *U1, const float *v0, const float *v1,
                               const float *log_jac, float *out) {
  const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g >= n) return;
  float k0 = 0.f, k1 = 0.f;
  for (int d = 0; d < D; ++d) {
    k0 = fmaf(v0[g * D + d], v0[g * D + d], k0);
    k1 = fmaf(v1[g * D + d], v1[g * D + d], k1);
  }
  out[g] = accept_prob(U0[g] + 0.5f * k0, U1[g] + 0.5f * k1, log_jac[g]);
}

// S, T, Q = ScaleTanh / identity of the raw head outputs (Dynamics net call as a component).
__global__ void k_lay_heads_out(LayDims dm, const float *hd, const float *es, const float *eq, float *S, float *T,
                                float *Q, long long n_chains) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_chains * dm.D) return;
  const long long n = i / dm.D;
  const int d = (int)(i - n * dm.D);
  const float *h = hd + n * dm.N3p;
  S[i] = es[d] * tanhf(h[d]);
  T[i] = h[dm.D + d];
  Q[i] = eq[d] * tanhf(h[2 * dm.D + d]);
}

}  // namespace layered
}  // namespace l2hmc
