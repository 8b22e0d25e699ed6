// Host side of the training path (train.cuh): l2hmc_loss_grad.  Included by l2hmc_api.cu after the component kernels
// (k_grad, k_hamiltonian) and the context type.
#pragma once
#include "train.cuh"

namespace {

namespace tr = l2hmc::train;

// Device scratch of l2hmc_loss_grad: one allocation held by the context, grown on demand (no allocation and no
// synchronisation in the steady state); sub-buffers are carved out at 256-byte boundaries.
int tr_workspace(l2hmc_ctx *ctx, size_t n_floats, float **out) {
  if (ctx->train_ws.n < n_floats) {
    if (ctx->train_ws.p) cudaFree(ctx->train_ws.p);  // waits for work that still uses it
    ctx->train_ws.p = nullptr;
    ctx->train_ws.n = 0;
    void *p = nullptr;
    if (cudaMalloc(&p, n_floats * sizeof(float)) != cudaSuccess) {
      cudaGetLastError();
      return fail(ctx, L2HMC_ENOMEM, "l2hmc_loss_grad: out of device memory (%zu bytes of scratch; the record of the sub-updates takes %d x 2 x x_dim floats per chain)",
                  n_floats * sizeof(float), 4 * ctx->sh.T);
    }
    ctx->train_ws.p = static_cast<float *>(p);
    ctx->train_ws.n = n_floats;
  }
  *out = ctx->train_ws.p;
  return L2HMC_OK;
}

struct TrNetBufs {
  const float *ab;          // [n, 2D] net input
  const float *ct, *st;     // [n] time features
  float *h1, *h2, *hd;      // [n, H], [n, H], [n, 3D]
};

struct TrGB {
  dim3 g, b;
};
inline TrGB tr_gb(dim3 g, dim3 b) { return TrGB{g, b}; }

// One launch.  `gb` is a TrGB (tr_gb(GRID(n)) / TR_EGRID / TR_WGRID).  The second expansion runs the same kernel
// source on host threads (tests/emu/train_emu.cpp: the CPU suite checks this file and train.cuh against the oracle
// without a GPU); it is never part of libl2hmc.so.
#ifndef L2HMC_TRAIN_EMU
#define TR_KERNEL(kernel, gb, stream, ...) kernel<<<(gb).g, (gb).b, 0, stream>>>(__VA_ARGS__)
#else
#define TR_KERNEL(kernel, gb, stream, ...) emu::launch(#kernel, (gb).g, (gb).b, [&] { kernel(__VA_ARGS__); })
#endif
#define TR_LAUNCH(ctx, kernel, gb, stream, ...)        \
  do {                                                  \
    TR_KERNEL(kernel, gb, stream, __VA_ARGS__);         \
    CUDA_TRY(ctx, cudaGetLastError());                  \
    (ctx)->launches++;                                  \
  } while (0)

// Parts of a reduction over K elements with a smallest part size: at most L2HMC_TR_MAX_PARTS of them
inline long long tr_part_size(long long K, long long smallest) {
  const long long grown = (K + L2HMC_TR_MAX_PARTS - 1) / L2HMC_TR_MAX_PARTS;
  return grown > smallest ? grown : smallest;
}
// floats of the scratch the split reductions of one call need (largest product: [max(H, D)][H] per part)
inline size_t tr_part_floats(int D, int H) { return (size_t)L2HMC_TR_MAX_PARTS * (size_t)(H > D ? H : D) * (size_t)(H > D ? H : D); }

#define TR_EGRID(tot) tr_gb((unsigned)(((tot) + 255) / 256), 256)
#define TR_WGRID(n) tr_gb((unsigned)(((n) * 32 + 255) / 256), 256)

// mode 2 (C += A B with K = the chains): split K over CTAs into ctx's part buffer, then an ordered reduction
int tr_gemm(l2hmc_ctx *ctx, cudaStream_t s, tr::Gemm g) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return L2HMC_OK;
  const long long gy = (g.M + 63) / 64;
  long long gz = 1;
  float *dst = g.C;
  const long long dst_ld = g.ldc;
  if (g.mode == 2) {
    g.kchunk = tr_part_size(g.K, L2HMC_TR_KCHUNK);
    gz = (g.K + g.kchunk - 1) / g.kchunk;
    g.C = ctx->train_part;
    g.ldc = g.N;
  } else {
    g.kchunk = g.K;
  }
  if (gy > 65535 || gz > 65535) return fail(ctx, L2HMC_EUNSUPPORTED, "l2hmc_loss_grad: more than 4.1M chains per call");
#define TR_GEMM_GRID tr_gb(dim3((unsigned)((g.N + 63) / 64), (unsigned)gy, (unsigned)gz), 256)
  TR_LAUNCH(ctx, tr::k_gemm, TR_GEMM_GRID, s, g);
  if (g.mode == 2) TR_LAUNCH(ctx, tr::k_reduce_add, TR_EGRID(g.M * g.N), s, ctx->train_part, (int)gz, g.M, g.N, dst, dst_ld);
  return L2HMC_OK;
}

tr::Gemm tr_g(const float *A, long long sam, long long sak, const float *B, long long sbk, long long sbn, float *C,
              long long ldc, long long M, int N, long long K, int mode) {
  tr::Gemm g;
  g.A = A; g.sam = sam; g.sak = sak; g.B = B; g.sbk = sbk; g.sbn = sbn; g.C = C; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K; g.mode = mode; g.kchunk = K;
  return g;
}

// out[c] += sum_r w[r] A[r][c]: slabs of rows into the part buffer, then the ordered reduction
int tr_colsum(l2hmc_ctx *ctx, cudaStream_t s, const float *A, long long lda, long long n, int cols, const float *w, float *out) {
  const long long slab = tr_part_size(n, L2HMC_TR_SLAB);
  const long long parts = (n + slab - 1) / slab;
#define TR_COLSUM_GRID tr_gb(dim3((unsigned)((cols + 127) / 128), (unsigned)parts), 128)
  TR_LAUNCH(ctx, tr::k_colsum, TR_COLSUM_GRID, s, A, lda, n, cols, w, ctx->train_part, slab);
  TR_LAUNCH(ctx, tr::k_reduce_add, TR_EGRID(cols), s, ctx->train_part, (int)parts, 1ll, cols, out, (long long)cols);
  return L2HMC_OK;
}

// [S | T | Q] pre-activations of net([a, b, t]) for every chain (SCGExperiment.ipynb:51-77), activations kept
int tr_net_forward(l2hmc_ctx *ctx, cudaStream_t s, const NetRaw &w, long long n, const TrNetBufs &b) {
  const int D = ctx->sh.D, H = ctx->sh.H;
  int rc;
  if ((rc = tr_gemm(ctx, s, tr_g(b.ab, 2 * D, 1, w.W1, H, 1, b.h1, H, n, H, D, 0)))) return rc;
  if ((rc = tr_gemm(ctx, s, tr_g(b.ab + D, 2 * D, 1, w.W2, H, 1, b.h1, H, n, H, D, 1)))) return rc;
  TR_LAUNCH(ctx, tr::k_act1, TR_EGRID(n * H), s, n, H, b.h1, w.b1, w.b2, w.b3, w.W3, b.ct, b.st);
  if ((rc = tr_gemm(ctx, s, tr_g(b.h1, H, 1, w.W4, H, 1, b.h2, H, n, H, H, 0)))) return rc;
  TR_LAUNCH(ctx, tr::k_act2, TR_EGRID(n * H), s, n, H, b.h2, w.b4);
  if ((rc = tr_gemm(ctx, s, tr_g(b.h2, H, 1, w.Ws, D, 1, b.hd, 3 * D, n, D, H, 0)))) return rc;
  if ((rc = tr_gemm(ctx, s, tr_g(b.h2, H, 1, w.Wt, D, 1, b.hd + D, 3 * D, n, D, H, 0)))) return rc;
  return tr_gemm(ctx, s, tr_g(b.h2, H, 1, w.Wq, D, 1, b.hd + 2 * D, 3 * D, n, D, H, 0));
}

// reverse of tr_net_forward: ghd [n, 3D] (cotangents of the pre-activations) -> gab [n, 2D]; parameter gradients += G
int tr_net_vjp(l2hmc_ctx *ctx, cudaStream_t s, const NetRaw &w, const l2hmc_net_grads &G, long long n, const TrNetBufs &b,
               const float *ghd, const float *sc, float *gh2, float *gh1, float *gab) {
  const int D = ctx->sh.D, H = ctx->sh.H;
  int rc;
  // heads: biases, log-scales, weights
  if ((rc = tr_colsum(ctx, s, ghd, 3 * D, n, D, nullptr, G.bs))) return rc;
  if ((rc = tr_colsum(ctx, s, ghd + D, 3 * D, n, D, nullptr, G.bt))) return rc;
  if ((rc = tr_colsum(ctx, s, ghd + 2 * D, 3 * D, n, D, nullptr, G.bq))) return rc;
  if ((rc = tr_colsum(ctx, s, sc, 2 * D, n, D, nullptr, G.scale_s))) return rc;
  if ((rc = tr_colsum(ctx, s, sc + D, 2 * D, n, D, nullptr, G.scale_q))) return rc;
  if ((rc = tr_gemm(ctx, s, tr_g(b.h2, 1, H, ghd, 3 * D, 1, G.Ws, D, H, D, n, 2)))) return rc;
  if ((rc = tr_gemm(ctx, s, tr_g(b.h2, 1, H, ghd + D, 3 * D, 1, G.Wt, D, H, D, n, 2)))) return rc;
  if ((rc = tr_gemm(ctx, s, tr_g(b.h2, 1, H, ghd + 2 * D, 3 * D, 1, G.Wq, D, H, D, n, 2)))) return rc;
  // gh2 = ghd_s Ws^T + ghd_t Wt^T + ghd_q Wq^T, masked by relu
  if ((rc = tr_gemm(ctx, s, tr_g(ghd, 3 * D, 1, w.Ws, 1, D, gh2, H, n, H, D, 0)))) return rc;
  if ((rc = tr_gemm(ctx, s, tr_g(ghd + D, 3 * D, 1, w.Wt, 1, D, gh2, H, n, H, D, 1)))) return rc;
  if ((rc = tr_gemm(ctx, s, tr_g(ghd + 2 * D, 3 * D, 1, w.Wq, 1, D, gh2, H, n, H, D, 1)))) return rc;
  TR_LAUNCH(ctx, tr::k_relu_mask, TR_EGRID(n * H), s, n * H, gh2, b.h2);
  if ((rc = tr_colsum(ctx, s, gh2, H, n, H, nullptr, G.b4))) return rc;
  if ((rc = tr_gemm(ctx, s, tr_g(b.h1, 1, H, gh2, H, 1, G.W4, H, H, H, n, 2)))) return rc;
  // gh1 = gz2 W4^T, masked
  if ((rc = tr_gemm(ctx, s, tr_g(gh2, H, 1, w.W4, 1, H, gh1, H, n, H, H, 0)))) return rc;
  TR_LAUNCH(ctx, tr::k_relu_mask, TR_EGRID(n * H), s, n * H, gh1, b.h1);
  if ((rc = tr_colsum(ctx, s, gh1, H, n, H, nullptr, G.b1))) return rc;
  if ((rc = tr_colsum(ctx, s, gh1, H, n, H, nullptr, G.b2))) return rc;
  if ((rc = tr_colsum(ctx, s, gh1, H, n, H, nullptr, G.b3))) return rc;
  if ((rc = tr_colsum(ctx, s, gh1, H, n, H, b.ct, G.W3))) return rc;       // row 0 meets cos, row 1 sin (:99-105)
  if ((rc = tr_colsum(ctx, s, gh1, H, n, H, b.st, G.W3 + H))) return rc;
  if ((rc = tr_gemm(ctx, s, tr_g(b.ab, 1, 2 * D, gh1, H, 1, G.W1, H, D, H, n, 2)))) return rc;
  if ((rc = tr_gemm(ctx, s, tr_g(b.ab + D, 1, 2 * D, gh1, H, 1, G.W2, H, D, H, n, 2)))) return rc;
  // cotangents of the two inputs
  if ((rc = tr_gemm(ctx, s, tr_g(gh1, H, 1, w.W1, 1, H, gab, 2 * D, n, D, H, 0)))) return rc;
  return tr_gemm(ctx, s, tr_g(gh1, H, 1, w.W2, 1, H, gab + D, 2 * D, n, D, H, 0));
}

bool tr_grads_complete(const l2hmc_net_grads &g) {
  return g.W1 && g.b1 && g.W2 && g.b2 && g.W3 && g.b3 && g.W4 && g.b4 && g.Ws && g.bs && g.Wt && g.bt && g.Wq && g.bq &&
         g.scale_s && g.scale_q;
}

int tr_loss_grad(l2hmc_ctx *ctx, const l2hmc_loss_grad_args *a) {
  const Shape &sh = ctx->sh;
  const int D = sh.D, DP = sh.DP, H = sh.H, T = sh.T;
  const long long n = a->n;
  const float eps = sh.eps;
  cudaStream_t s = (cudaStream_t)a->stream;
  const size_t nD = (size_t)n * D, nH = (size_t)n * H;
  float *tape_x, *tape_v, *x, *v, *gU, *ab, *h1, *h2, *gh1, *gh2, *hd, *ghd, *sc, *gab, *gx, *gv, *gg, *vec, *part;
  struct { float **p; size_t n; } req[] = {
      {&tape_x, nD * 4 * T}, {&tape_v, nD * 4 * T}, {&x, nD}, {&v, nD}, {&gU, nD}, {&ab, 2 * nD}, {&h1, nH}, {&h2, nH},
      {&gh1, nH}, {&gh2, nH}, {&hd, 3 * nD}, {&ghd, 3 * nD}, {&sc, 2 * nD}, {&gab, 2 * nD}, {&gx, nD}, {&gv, nD}, {&gg, nD},
      {&vec, (size_t)n * 10 + 4}, {&part, tr_part_floats(D, H)}};
#ifndef L2HMC_TRAIN_EMU
  auto pad = [](size_t k) { return (k + 63) / 64 * 64; };
#else  // host emulation: a 64-float guard zone (NaN-filled by the emulated cudaMalloc) after each sub-buffer, checked at the end
  auto pad = [](size_t k) { return (k + 63) / 64 * 64 + 64; };
#endif
  size_t total = 0;
  for (auto &r : req) total += pad(r.n);
  float *base = nullptr;
  int rc = tr_workspace(ctx, total, &base);
  if (rc) return rc;
  for (auto &r : req) {
    *r.p = base;
    base += pad(r.n);
  }
  ctx->train_part = part;
  float *logj = vec, *H0 = vec + n, *H1 = vec + 2 * n, *lossv = vec + 3 * n, *px = vec + 4 * n, *glj = vec + 5 * n,
        *geps = vec + 6 * n, *ct = vec + 7 * n, *st = vec + 8 * n, *vv = vec + 9 * n, *stats = vec + 10 * n;
  CUDA_TRY(ctx, cudaMemcpyAsync(x, a->x, nD * sizeof(float), cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(ctx, cudaMemcpyAsync(v, a->v, nD * sizeof(float), cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(ctx, cudaMemsetAsync(vec, 0, ((size_t)n * 10 + 4) * sizeof(float), s));
  const tr::Heads heads[2] = {{ctx->net_rawv[0].bs, ctx->net_rawv[0].bt, ctx->net_rawv[0].bq, ctx->net_rawv[0].ls, ctx->net_rawv[0].lq},
                              {ctx->net_rawv[1].bs, ctx->net_rawv[1].bt, ctx->net_rawv[1].bq, ctx->net_rawv[1].ls, ctx->net_rawv[1].lq}};
  const TrNetBufs nb = {ab, ct, st, h1, h2, hd};
  // sub-update j of a leapfrog step: 0 and 3 are V (momentum) updates, 1 and 2 the two masked X (position) updates
  auto which_of = [](int j) { return (j == 0 || j == 3) ? 0 : j; };

  // ---- forward sweep, recording the state in front of every sub-update ------------------------------------------
  for (int it = 0; it < T; ++it) {
    TR_LAUNCH(ctx, tr::k_tau, TR_EGRID(n), s, n, a->dir, it, T, ct, st);
    for (int j = 0; j < 4; ++j) {
      const int which = which_of(j), net_id = which == 0 ? L2HMC_VNET : L2HMC_XNET;
      const size_t rec = ((size_t)it * 4 + j) * nD;
      CUDA_TRY(ctx, cudaMemcpyAsync(tape_x + rec, x, nD * sizeof(float), cudaMemcpyDeviceToDevice, s));
      CUDA_TRY(ctx, cudaMemcpyAsync(tape_v + rec, v, nD * sizeof(float), cudaMemcpyDeviceToDevice, s));
      if (which == 0) TR_LAUNCH(ctx, k_grad, tr_gb(GRID(n)), s, ctx->en, sh, n, x, gU);
      TR_LAUNCH(ctx, tr::k_build_ab, TR_EGRID(n * D), s, n, D, DP, T, it, which, a->dir, ctx->mask.p, x, v, gU, ab);
      if ((rc = tr_net_forward(ctx, s, ctx->net_rawv[net_id], n, nb))) return rc;
      TR_LAUNCH(ctx, tr::k_update, TR_WGRID(n), s, n, D, DP, T, it, which, a->dir, ctx->mask.p, heads[net_id], eps, hd, gU,
                                                           x, v, logj);
    }
  }
  // ---- objective and the cotangents of (X, V, log|J|) -----------------------------------------------------------
  TR_LAUNCH(ctx, k_hamiltonian, tr_gb(GRID(n)), s, ctx->en, sh, n, tape_x, tape_v, H0);
  TR_LAUNCH(ctx, k_hamiltonian, tr_gb(GRID(n)), s, ctx->en, sh, n, x, v, H1);
  TR_LAUNCH(ctx, k_grad, tr_gb(GRID(n)), s, ctx->en, sh, n, x, gU);
  if (a->loss_kind >= 2) {  // the inverse / logsumexp losses weigh each chain by a statistic of the whole batch
    TR_LAUNCH(ctx, tr::k_loss_v, TR_WGRID(n), s, n, D, tape_x, x, H0, H1, logj, vv);
    TR_LAUNCH(ctx, tr::k_loss_stats, tr_gb(1, 256), s, n, vv, stats);
  }
  TR_LAUNCH(ctx, tr::k_loss, TR_WGRID(n), s, a->loss_kind, n, D, tape_x, x, v, H0, H1, logj, gU, stats, a->scale, a->inv_count,
            lossv, px, glj, gx, gv);
  if ((rc = tr_colsum(ctx, s, lossv, 1, n, 1, nullptr, a->loss))) return rc;
  if (a->x_out) CUDA_TRY(ctx, cudaMemcpyAsync(a->x_out, x, nD * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if (a->px_out) CUDA_TRY(ctx, cudaMemcpyAsync(a->px_out, px, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, s));

  // ---- reverse sweep ---------------------------------------------------------------------------------------------
  for (int it = T - 1; it >= 0; --it) {
    TR_LAUNCH(ctx, tr::k_tau, TR_EGRID(n), s, n, a->dir, it, T, ct, st);
    for (int j = 3; j >= 0; --j) {
      const int which = which_of(j), net_id = which == 0 ? L2HMC_VNET : L2HMC_XNET;
      const size_t rec = ((size_t)it * 4 + j) * nD;
      const float *xs = tape_x + rec, *vs = tape_v + rec;
      if (which == 0) TR_LAUNCH(ctx, k_grad, tr_gb(GRID(n)), s, ctx->en, sh, n, xs, gU);
      TR_LAUNCH(ctx, tr::k_build_ab, TR_EGRID(n * D), s, n, D, DP, T, it, which, a->dir, ctx->mask.p, xs, vs, gU, ab);
      if ((rc = tr_net_forward(ctx, s, ctx->net_rawv[net_id], n, nb))) return rc;
      TR_LAUNCH(ctx, tr::k_update_vjp, TR_WGRID(n), s, n, D, DP, T, it, which, a->dir, ctx->mask.p, heads[net_id], eps, hd,
                                                               gU, xs, vs, glj, gx, gv, ghd, sc, gg, geps);
      if ((rc = tr_net_vjp(ctx, s, ctx->net_rawv[net_id], net_id == L2HMC_VNET ? a->grad_vnet : a->grad_xnet, n, nb, ghd, sc,
                           gh2, gh1, gab)))
        return rc;
      TR_LAUNCH(ctx, tr::k_scatter, TR_EGRID(n * D), s, n, D, DP, T, it, which, a->dir, ctx->mask.p, gab, gx, gv, gg);
      if (which == 0) TR_LAUNCH(ctx, tr::k_hvp, tr_gb(GRID(n)), s, ctx->en, sh, n, xs, gg, gx);
    }
  }
  if ((rc = tr_colsum(ctx, s, geps, 1, n, 1, nullptr, a->d_eps))) return rc;
#ifdef L2HMC_TRAIN_EMU
  for (auto &r : req) {
    const uint32_t *guard = reinterpret_cast<const uint32_t *>(*r.p + (r.n + 63) / 64 * 64);
    for (int i = 0; i < 64; ++i)
      if (guard[i] != 0xFFFFFFFFu) return fail(ctx, L2HMC_ECUDA, "emulation: a kernel wrote past the end of a scratch buffer");
  }
#endif
  return L2HMC_OK;  // asynchronous on the stream, like l2hmc_transition
}

}  // namespace
