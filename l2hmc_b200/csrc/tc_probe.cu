// tcgen05 bring-up probe (sm_100a): D[128 x N] = A[128 x K] * B[N x K]^T with the 3xTF32 split,
// A staged in TMEM by the threads (tcgen05.st), B in shared memory (K-major, no swizzle, canonical
// core-matrix layout), accumulators in TMEM, read back with tcgen05.ld.  Prints the error against fp64
// for the layout / descriptor variants, so one GPU run settles the encodings used by kernel_tc.cuh.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu ; run: ./tc_probe
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "tc_common.cuh"

using namespace l2hmc::tc;

// layout_mode 0: B core matrices ordered [k_core][n_group]   -> LBO = (N/8)*128, SBO = 128
// layout_mode 1: B core matrices ordered [n_group][k_core]   -> LBO = 128, SBO = (K/4)*128
// terms: 1 = hi*hi only (plain TF32), 3 = 3xTF32
__global__ void __launch_bounds__(128, 1) probe_kernel(const float *A, const float *B, float *Dout, int N, int K,
                                                        int layout_mode, int terms, int swap_lbo_sbo) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5;
  float *Bhi = reinterpret_cast<float *>(smem_raw);
  float *Blo = Bhi + (size_t)N * K;

  if (warp == 0) tmem_alloc(&tmem_base_slot, 512);
  if (tid == 0) {
    mbar_init(&mbar, 1);
    fence_mbar_init();
  }
  // B -> canonical K-major no-swizzle layout: core matrix = 8 rows (n) x 16 bytes (4 k)
  const int NG = N / 8, KCORES = K / 4;
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i - n * K;
    const int ng = n >> 3, r = n & 7, kc = k >> 2, e = k & 3;
    const size_t core = layout_mode == 0 ? ((size_t)kc * NG + ng) : ((size_t)ng * KCORES + kc);
    const float b = B[i];
    const float hi = tf32_rna(b);
    const float lo = tf32_rna(b - hi);
    Bhi[core * 32 + r * 4 + e] = hi;
    Blo[core * 32 + r * 4 + e] = lo;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_slot;
  const uint32_t t_acc = tmem;            // columns [0, N)
  const uint32_t t_ahi = tmem + 256;      // columns [256, 256+K)
  const uint32_t t_alo = tmem + 256 + 128;

  // A row of this thread -> TMEM (lane = row)
  {
    const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
    for (int k0 = 0; k0 < K; k0 += 8) {
      float hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = A[(size_t)tid * K + k0 + j];
        hi[j] = tf32_rna(a);
        lo[j] = tf32_rna(a - hi[j]);
      }
      tmem_st8(t_ahi + lane_base + k0, hi);
      tmem_st8(t_alo + lane_base + k0, lo);
    }
    tmem_wait_st();
  }
  fence_proxy_async_smem();  // generic-proxy smem writes (B) -> visible to the tensor core's async proxy
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();

  if (tid == 0) {
    const uint32_t idesc = make_idesc_tf32(128, N);
    uint32_t lbo = layout_mode == 0 ? NG * 128 : 128;
    uint32_t sbo = layout_mode == 0 ? 128 : KCORES * 128;
    if (swap_lbo_sbo) { uint32_t t = lbo; lbo = sbo; sbo = t; }
    const uint32_t kstep_bytes = layout_mode == 0 ? 2 * NG * 128 : 2 * 128;  // two k-cores per MMA (K=8)
    const uint32_t bhi0 = smem_u32(Bhi), blo0 = smem_u32(Blo);
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint64_t dhi = make_smem_desc(bhi0 + ks * kstep_bytes, lbo, sbo);
      const uint64_t dlo = make_smem_desc(blo0 + ks * kstep_bytes, lbo, sbo);
      if (terms == 3) {
        mma_tf32_ts(t_acc, t_alo + ks * 8, dhi, idesc, ks > 0);
        mma_tf32_ts(t_acc, t_ahi + ks * 8, dlo, idesc, true);
        mma_tf32_ts(t_acc, t_ahi + ks * 8, dhi, idesc, true);
      } else {
        mma_tf32_ts(t_acc, t_ahi + ks * 8, dhi, idesc, ks > 0);
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

static double run(int N, int K, int layout_mode, int terms, int swap, const std::vector<float> &A, const std::vector<float> &B,
                  const std::vector<double> &ref, double *max_ref) {
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dB, B.size() * 4);
  cudaMalloc(&dD, (size_t)128 * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, (size_t)128 * N * 4);
  const size_t smem = (size_t)2 * N * K * 4;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 128, smem>>>(dA, dB, dD, N, K, layout_mode, terms, swap);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("  CUDA error: %s\n", cudaGetErrorString(e));
    exit(2);
  }
  std::vector<float> D((size_t)128 * N);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double err = 0, mr = 0;
  for (size_t i = 0; i < D.size(); ++i) {
    err = fmax(err, fabs((double)D[i] - ref[i]));
    mr = fmax(mr, fabs(ref[i]));
  }
  *max_ref = mr;
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return err;
}

// ---- micro-benchmark 1: tensor-pipe rate for back-to-back TS-mode tf32 MMAs of one tile shape ----------
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int nmma, long long *cycles_out) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5;
  float *Bs = reinterpret_cast<float *>(smem_raw);
  for (int i = tid; i < N * 8; i += 128) Bs[i] = 0.001f * (i % 17);
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
    const uint32_t idesc = make_idesc_tf32(128, N);
    const uint64_t d = make_smem_desc(smem_u32(Bs), (N / 8) * 128, 128);
    const long long t0 = clock64();
    for (int i = 0; i < nmma; ++i) mma_tf32_ts(tmem, tmem + 256, d, idesc, i > 0);
    tcgen05_commit(&mbar);
    mbar_wait(&mbar, 0);
    const long long t1 = clock64();
    cycles_out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- micro-benchmark 2: every CTA streams the same weight image from L2 into a shared-memory ring with
// 1-D TMA bulk copies (the pattern kernel_tc uses for the B operands) -------------------------------------
__global__ void __launch_bounds__(128, 1) stream_kernel(const uint8_t *img, int img_bytes, int chunk_bytes, int nslots,
                                                        int reps, long long *cycles_out) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[8];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < nslots; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    const int nchunks = img_bytes / chunk_bytes;
    const long long total = (long long)nchunks * reps;
    const long long t0 = clock64();
    long long issued = 0, done = 0;
    for (; issued < nslots && issued < total; ++issued) {
      const int s = (int)(issued % nslots);
      mbar_arrive_expect_tx(&full[s], chunk_bytes);
      bulk_g2s(smem_raw + (size_t)s * chunk_bytes, img + (size_t)(issued % nchunks) * chunk_bytes, chunk_bytes, &full[s]);
    }
    for (; done < total; ++done) {
      const int s = (int)(done % nslots);
      mbar_wait(&full[s], (uint32_t)((done / nslots) & 1));
      if (issued < total) {  // slot is free again (nobody consumes here): refill it
        mbar_arrive_expect_tx(&full[s], chunk_bytes);
        bulk_g2s(smem_raw + (size_t)s * chunk_bytes, img + (size_t)(issued % nchunks) * chunk_bytes, chunk_bytes, &full[s]);
        ++issued;
      }
    }
    cycles_out[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
}

static void micro() {
  long long *dc;
  cudaMalloc(&dc, 1024 * sizeof(long long));
  std::vector<long long> hc(1024);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  for (int N : {64, 80, 112, 160, 256}) {
    for (int grid : {1, 148}) {
      const int nmma = 2000;
      cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      mma_rate_kernel<<<grid, 128, 64 * 1024>>>(N, nmma, dc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mma_rate CUDA error: %s\n", cudaGetErrorString(e)); exit(3); }
      cudaMemcpy(hc.data(), dc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < grid; ++i) mx = hc[i] > mx ? hc[i] : mx;
      printf("MMA_RATE N=%d grid=%d: %.1f cycles per 128xNx8 tf32 MMA (%.0f MAC/cycle/SM)\n", N, grid, (double)mx / nmma,
             128.0 * N * 8 * nmma / mx);
    }
  }
  const int img_bytes = 320 * 1024;
  uint8_t *img;
  cudaMalloc(&img, img_bytes);
  cudaMemset(img, 1, img_bytes);
  for (int chunk : {8192, 16384, 32768}) {
    for (int grid : {1, 148, 296}) {
      const int nslots = 4, reps = 40;
      cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, nslots * chunk);
      stream_kernel<<<grid, 128, nslots * chunk>>>(img, img_bytes, chunk, nslots, reps, dc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("stream CUDA error: %s\n", cudaGetErrorString(e)); exit(4); }
      cudaMemcpy(hc.data(), dc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < grid; ++i) mx = hc[i] > mx ? hc[i] : mx;
      const double bytes = (double)img_bytes * reps;
      printf("STREAM chunk=%d slots=%d grid=%d: %.1f B/cycle per CTA, %.0f B/cycle chip-wide (max over CTAs %lld cycles)\n", chunk,
             nslots, grid, bytes / mx, bytes * grid / mx, mx);
    }
  }
  cudaFree(img);
  cudaFree(dc);
}

int main(int argc, char **argv) {
  if (argc > 1 && argv[1][0] == 'm') { micro(); return 0; }
  const int shapes[3][2] = {{112, 104}, {160, 104}, {64, 56}};
  for (auto &s : shapes) {
    const int N = s[0], K = s[1];
    std::vector<float> A((size_t)128 * K), B((size_t)N * K);
    srand(1234);
    for (auto &v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto &v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    std::vector<double> ref((size_t)128 * N);
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        double acc = 0;
        for (int k = 0; k < K; ++k) acc += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
        ref[(size_t)m * N + n] = acc;
      }
    for (int layout = 0; layout < 2; ++layout)
      for (int swap = 0; swap < 2; ++swap)
        for (int terms = 1; terms <= 3; terms += 2) {
          double mr;
          const double err = run(N, K, layout, terms, swap, A, B, ref, &mr);
          printf("N=%d K=%d layout=%d swap=%d terms=%d  max_abs_err=%.3e (max|ref|=%.2f)\n", N, K, layout, swap, terms, err, mr);
          fflush(stdout);
        }
  }
  return 0;
}
