// Bring-up / microbenchmark harness for tc_gemm.cuh (not part of libl2hmc.so):
//   gemm_probe check           correctness of every epilogue on ragged shapes against a double-precision host GEMM
//   gemm_probe time M N K      average kernel time and TFLOP/s, next to layered::sgemm_kernel on the same problem
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "tc_gemm.cuh"

using namespace l2hmc;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }
static int round_up(int a, int b) { return (a + b - 1) / b * b; }

struct Problem {
  int M, N, K, lda, ldb, ldc, epi;
  std::vector<float> A, B, C0, bias, bias_b, R;
  std::vector<int> dir;
  float *dA, *dB, *dC, *dbias, *dbias_b, *dR, *dpk;
  int *ddir;
  tcg::TcGemmB tb;
  layered::GemmArgs g;
};

static void build(Problem &p, int M, int N, int K, int epi, bool use_dir, bool use_R) {
  p.M = M; p.N = N; p.K = K; p.epi = epi;
  p.lda = round_up(K, 8); p.ldb = round_up(N, 8); p.ldc = round_up(N, 8) + 8;
  p.A.assign((size_t)M * p.lda, 0.f);
  p.B.assign((size_t)round_up(K, 16) * p.ldb, 0.f);
  p.C0.assign((size_t)M * p.ldc, 0.f);
  p.bias.assign(p.ldb, 0.f); p.bias_b.assign(p.ldb, 0.f); p.R.assign((size_t)M * p.ldc, 0.f); p.dir.assign(M, 1);
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) p.A[(size_t)m * p.lda + k] = frand();
  for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) p.B[(size_t)k * p.ldb + n] = frand() * 0.3f;
  for (int n = 0; n < N; ++n) { p.bias[n] = frand(); p.bias_b[n] = frand(); }
  for (size_t i = 0; i < p.C0.size(); ++i) { p.C0[i] = fabsf(frand()) + 0.01f; p.R[i] = frand(); }
  for (int m = 0; m < M; ++m) p.dir[m] = rand() & 1;
  std::vector<float> pk;
  tcg::pack_b(p.B.data(), p.ldb, K, N, pk, &p.tb);
  CK(cudaMalloc(&p.dA, p.A.size() * 4)); CK(cudaMalloc(&p.dB, p.B.size() * 4)); CK(cudaMalloc(&p.dC, p.C0.size() * 4));
  CK(cudaMalloc(&p.dbias, p.bias.size() * 4)); CK(cudaMalloc(&p.dbias_b, p.bias.size() * 4));
  CK(cudaMalloc(&p.dR, p.R.size() * 4)); CK(cudaMalloc(&p.ddir, M * 4)); CK(cudaMalloc(&p.dpk, pk.size() * 4));
  CK(cudaMemcpy(p.dA, p.A.data(), p.A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p.dB, p.B.data(), p.B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p.dbias, p.bias.data(), p.bias.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p.dbias_b, p.bias_b.data(), p.bias.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p.dR, p.R.data(), p.R.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p.ddir, p.dir.data(), M * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p.dpk, pk.data(), pk.size() * 4, cudaMemcpyHostToDevice));
  p.tb.pk = p.dpk;
  layered::GemmArgs &g = p.g;
  g.A = p.dA; g.lda = p.lda; g.B = p.dB; g.ldb = p.ldb; g.Bn = p.ldb; g.C = p.dC; g.ldc = p.ldc;
  g.M = M; g.N = N; g.K = round_up(K, 8);
  g.bias = p.dbias; g.bias_b = use_dir ? p.dbias_b : p.dbias; g.dir = use_dir ? p.ddir : nullptr;
  g.R = (use_R || epi == layered::EPI_ADD_SCALE) ? p.dR : nullptr; g.ldr = p.ldc; g.scale = 0.75f; g.epi = epi;
  g.vec = (p.ldc % 4 == 0 && N % 4 == 0) ? 1 : 0;
  if (epi == layered::EPI_DSOFTPLUS || epi == layered::EPI_ADD_SCALE) g.bias = g.bias_b = nullptr;
}

static void destroy(Problem &p) {
  cudaFree(p.dA); cudaFree(p.dB); cudaFree(p.dC); cudaFree(p.dbias); cudaFree(p.dbias_b); cudaFree(p.dR); cudaFree(p.ddir); cudaFree(p.dpk);
}

static void launch_tc(Problem &p) {
  static int sms = 0;
  if (!sms) CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(tcg::launch_tc_gemm(p.g, p.tb, sms, 0));
}
static void launch_fma(Problem &p) {
  layered::GemmArgs g = p.g;
  g.K = round_up(p.K, 8);
  const int bn8 = round_up(g.N, 128);
  layered::sgemm_kernel<8><<<dim3(bn8 / 128, (p.M + 127) / 128), 256>>>(g);
}

static double reference(const Problem &p, int m, int n) {
  double s = 0;
  for (int k = 0; k < p.K; ++k) s += (double)p.A[(size_t)m * p.lda + k] * (double)p.B[(size_t)k * p.ldb + n];
  const std::vector<float> &b = (p.g.dir && p.dir[m] == 0) ? p.bias_b : p.bias;
  switch (p.epi) {
    case layered::EPI_DSOFTPLUS: return s * (1.0 - exp(-(double)p.C0[(size_t)m * p.ldc + n]));
    case layered::EPI_ADD_SCALE: return (s + p.R[(size_t)m * p.ldc + n]) * 0.75;
    default: break;
  }
  if (p.g.bias) s += b[n];
  if (p.g.R) s += p.R[(size_t)m * p.ldc + n];
  if (p.epi == layered::EPI_RELU) s = s > 0 ? s : 0;
  if (p.epi == layered::EPI_SOFTPLUS) s = (s > 0 ? s : 0) + log1p(exp(-fabs(s)));
  return s;
}

static int check_one(int M, int N, int K, int epi, bool use_dir, bool use_R) {
  Problem p;
  build(p, M, N, K, epi, use_dir, use_R);
  int bad = 0;
  for (int which = 0; which < 2; ++which) {
    CK(cudaMemcpy(p.dC, p.C0.data(), p.C0.size() * 4, cudaMemcpyHostToDevice));
    if (which == 0) launch_tc(p); else launch_fma(p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> C(p.C0.size());
    CK(cudaMemcpy(C.data(), p.dC, C.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    long untouched_bad = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < p.ldc; ++n) {
        const float got = C[(size_t)m * p.ldc + n];
        if (n >= N) { if (got != p.C0[(size_t)m * p.ldc + n]) untouched_bad++; continue; }
        const double ref = reference(p, m, n);
        maxerr = fmax(maxerr, fabs(got - ref));
        maxref = fmax(maxref, fabs(ref));
      }
    const bool ok = maxerr <= 2e-5 * fmax(1.0, maxref) && untouched_bad == 0;
    printf("CHECK %s M=%d N=%d K=%d epi=%d dir=%d R=%d BN=%d nblk=%d: max_abs_err=%.3e (max|ref|=%.2f) pad_writes=%ld %s\n",
           which == 0 ? "tc " : "fma", M, N, K, epi, (int)use_dir, (int)use_R, p.tb.BN, p.tb.nblk, maxerr, maxref, untouched_bad,
           ok ? "ok" : "FAIL");
    bad += ok ? 0 : 1;
  }
  destroy(p);
  return bad;
}

static void time_one(int M, int N, int K, int epi) {
  Problem p;
  build(p, M, N, K, epi, false, false);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int which = 0; which < 2; ++which) {
    for (int i = 0; i < 3; ++i) { if (which == 0) launch_tc(p); else launch_fma(p); }
    CK(cudaDeviceSynchronize());
    const int reps = 10;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) { if (which == 0) launch_tc(p); else launch_fma(p); }
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    printf("TIME %s epi=%d M=%d N=%d K=%d: %.3f ms  %.1f TFLOP/s (fp32-equivalent 2MNK)\n", which == 0 ? "tc " : "fma", epi, M, N, K, ms,
           2.0 * M * N * K / (ms * 1e-3) / 1e12);
  }
  destroy(p);
}

int main(int argc, char **argv) {
  srand(1);
  if (argc >= 2 && !strcmp(argv[1], "check")) {
    int bad = 0;
    bad += check_one(128, 64, 16, layered::EPI_BIAS, false, false);
    bad += check_one(128, 256, 64, layered::EPI_BIAS, false, false);
    bad += check_one(300, 200, 104, layered::EPI_RELU, true, true);
    bad += check_one(131, 152, 200, layered::EPI_BIAS, false, false);
    bad += check_one(257, 784, 1024, layered::EPI_BIAS, false, false);
    bad += check_one(129, 1024, 56, layered::EPI_SOFTPLUS, false, false);
    bad += check_one(200, 1024, 784, layered::EPI_DSOFTPLUS, false, false);
    bad += check_one(77, 50, 1024, layered::EPI_ADD_SCALE, false, false);
    bad += check_one(64, 24, 24, layered::EPI_RELU, false, false);
    printf("%s\n", bad ? "SOME CHECKS FAILED" : "ALL CHECKS PASSED");
    return bad ? 1 : 0;
  }
  if (argc >= 5 && !strcmp(argv[1], "time")) {
    time_one(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), argc >= 6 ? atoi(argv[5]) : layered::EPI_SOFTPLUS);
    return 0;
  }
  printf("usage: gemm_probe check | time M N K\n");
  return 0;
}
