// tcgen05 kind::f16 bring-up probe (sm_100a): D[128 x N] = A[128 x K] * B[N x K]^T in fp32-level accuracy from
// THREE fp16 MMAs per product (a = a_hi + a_lo, b = b_hi + b_lo as fp16 pairs: acc += a_lo*b_hi; a_hi*b_lo; a_hi*b_hi),
// A packed two fp16 per 32-bit TMEM column (tcgen05.st), B in shared memory (K-major, no swizzle, core matrix = 8 rows
// x 16 bytes = 8 halfs), fp32 accumulators in TMEM.  One fp16 MMA covers K = 16 in the cycles a tf32 MMA needs for
// K = 8, so the split costs half the tensor time of 3xTF32.  The probe settles (a) the instruction descriptor of
// kind::f16, (b) the order of the two halfs inside a TMEM column, (c) the achieved accuracy, (d) the MMA rate.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o f16_probe f16_probe.cu ; run: ./f16_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "tc_common.cuh"

using namespace l2hmc::tc;

// Instruction descriptor, kind::f16: [4,6) D format = 1 (f32); [7,10) A format = 0 (f16); [10,13) B format = 0 (f16);
// bits 15/16 A/B major = 0 (K); [17,23) N >> 3; [24,29) M >> 4.
__host__ __device__ inline uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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
__device__ __forceinline__ void tmem_st4u(uint32_t taddr, const uint32_t (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
__device__ __forceinline__ uint32_t pack2(float lo_k, float hi_k, int order) {  // order 0: low half = lower k
  const __half2 h = order == 0 ? __floats2half2_rn(lo_k, hi_k) : __floats2half2_rn(hi_k, lo_k);
  return *reinterpret_cast<const uint32_t *>(&h);
}

// K multiple of 16.  order: which half of a TMEM column holds the lower k.  terms: 1 = hi*hi only, 3 = split.
__global__ void __launch_bounds__(128, 1) probe_kernel(const float *A, const float *B, float *Dout, int N, int K, int order, int terms) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5;
  __half *Bhi = reinterpret_cast<__half *>(smem_raw);
  __half *Blo = Bhi + (size_t)N * K;
  if (warp == 0) tmem_alloc(&tmem_base_slot, 512);
  if (tid == 0) {
    mbar_init(&mbar, 1);
    fence_mbar_init();
  }
  // B -> K-major no-swizzle: per K=16 step [k_core (2)][n_group (N/8)][row (8)][8 halfs]
  const int NG = N / 8;
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i - n * K;
    const int ks = k >> 4, kk = k & 15;
    const size_t idx = (size_t)ks * 16 * N + ((size_t)(kk >> 3) * NG + (n >> 3)) * 64 + (size_t)(n & 7) * 8 + (kk & 7);
    const float b = B[i];
    const __half hi = __float2half_rn(b);
    const __half lo = __float2half_rn(b - __half2float(hi));
    Bhi[idx] = hi;
    Blo[idx] = lo;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_slot;
  const uint32_t t_acc = tmem, t_ahi = tmem + 256, t_alo = tmem + 384;
  {
    const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
    for (int k0 = 0; k0 < K; k0 += 8) {  // 8 k = 4 columns
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a0 = A[(size_t)tid * K + k0 + 2 * j], a1 = A[(size_t)tid * K + k0 + 2 * j + 1];
        const float h0 = __half2float(__float2half_rn(a0)), h1 = __half2float(__float2half_rn(a1));
        hi[j] = pack2(a0, a1, order);
        lo[j] = pack2(a0 - h0, a1 - h1, order);
      }
      tmem_st4u(t_ahi + lane_base + k0 / 2, hi);
      tmem_st4u(t_alo + lane_base + k0 / 2, lo);
    }
    tmem_wait_st();
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint32_t lbo = NG * 128, sbo = 128;
    const uint32_t kstep_bytes = 2 * NG * 128;  // two k-cores (8 halfs each) per MMA (K = 16)
    const uint32_t bhi0 = smem_u32(Bhi), blo0 = smem_u32(Blo);
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t dhi = make_smem_desc(bhi0 + ks * kstep_bytes, lbo, sbo);
      const uint64_t dlo = make_smem_desc(blo0 + ks * kstep_bytes, lbo, sbo);
      if (terms == 3) {
        mma_f16_ts(t_acc, t_alo + ks * 8, dhi, idesc, ks > 0);
        mma_f16_ts(t_acc, t_ahi + ks * 8, dlo, idesc, true);
        mma_f16_ts(t_acc, t_ahi + ks * 8, dhi, idesc, true);
      } else {
        mma_f16_ts(t_acc, t_ahi + ks * 8, dhi, idesc, ks > 0);
      }
    }
    tcgen05_commit(&mbar);
  }
  mbar_wait(&mbar, 0);
  tcgen05_fence_after();
  {
    const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
    for (int n0 = 0; n0 < N; n0 += 8) {
      float d[8];
      tmem_ld8(t_acc + lane_base + n0, d);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) Dout[(size_t)tid * N + n0 + j] = d[j];
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int nmma, long long *cycles_out) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5;
  __half *Bs = reinterpret_cast<__half *>(smem_raw);
  for (int i = tid; i < N * 16; i += 128) Bs[i] = __float2half(0.001f * (i % 17));
  if (warp == 0) tmem_alloc(&tmem_base_slot, 512);
  if (tid == 0) {
    mbar_init(&mbar, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_slot;
  {
    float z[8] = {1.f, 0.5f, 0.25f, 0.f, 0.f, 0.f, 0.f, 0.f};
    tmem_st8(tmem + 256 + (((uint32_t)(warp * 32)) << 16), z);
    tmem_wait_st();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint64_t d = make_smem_desc(smem_u32(Bs), (N / 8) * 128, 128);
    const long long t0 = clock64();
    for (int i = 0; i < nmma; ++i) mma_f16_ts(tmem, tmem + 256, d, idesc, i > 0);
    tcgen05_commit(&mbar);
    mbar_wait(&mbar, 0);
    cycles_out[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  const int shapes[3][2] = {{112, 112}, {96, 112}, {64, 64}};
  for (auto &s : shapes) {
    const int N = s[0], K = s[1];
    std::vector<float> A((size_t)128 * K), B((size_t)N * K);
    srand(1234);
    for (auto &v : A) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 8.f;   // activations up to 8
    for (auto &v : B) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.3f;  // weights up to 0.3
    for (int m = 0; m < 128; ++m) A[(size_t)m * K + 3] = 1e-6f * (m + 1);  // tiny values: fp16 subnormal range
    std::vector<double> ref((size_t)128 * N);
    double mr = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        double acc = 0;
        for (int k = 0; k < K; ++k) acc += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
        ref[(size_t)m * N + n] = acc;
        mr = fmax(mr, fabs(acc));
      }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4);
    cudaMalloc(&dB, B.size() * 4);
    cudaMalloc(&dD, (size_t)128 * N * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    for (int order = 0; order < 2; ++order)
      for (int terms : {1, 3}) {
        cudaMemset(dD, 0, (size_t)128 * N * 4);
        const size_t smem = (size_t)2 * N * K * 2;
        cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe_kernel<<<1, 128, smem>>>(dA, dB, dD, N, K, order, terms);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("N=%d K=%d order=%d terms=%d CUDA error: %s\n", N, K, order, terms, cudaGetErrorString(e));
          return 2;
        }
        std::vector<float> D((size_t)128 * N);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double err = 0;
        for (size_t i = 0; i < D.size(); ++i) err = fmax(err, fabs((double)D[i] - ref[i]));
        printf("F16 N=%d K=%d order=%d terms=%d  max_abs_err=%.3e (max|ref|=%.2f, rel %.2e)\n", N, K, order, terms, err, mr, err / mr);
      }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  long long *dc;
  cudaMalloc(&dc, 1024 * sizeof(long long));
  std::vector<long long> hc(1024);
  for (int N : {64, 80, 96, 112, 160, 256}) {
    const int nmma = 2000;
    cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    mma_rate_kernel<<<148, 128, 64 * 1024>>>(N, nmma, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mma_rate CUDA error: %s\n", cudaGetErrorString(e)); return 3; }
    cudaMemcpy(hc.data(), dc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = hc[i] > mx ? hc[i] : mx;
    printf("F16_MMA_RATE N=%d: %.1f cycles per 128xNx16 fp16 MMA (%.0f MAC/cycle/SM)\n", N, (double)mx / nmma, 128.0 * N * 16 * nmma / mx);
  }
  return 0;
}
