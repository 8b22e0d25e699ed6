// Host side of the layered engine (see layered.cuh): parameter packing, workspace, launch sequence.
// Included by l2hmc_api.cu after l2hmc_ctx / fail / ensure / CUDA_TRY are defined.
#pragma once

namespace {

using l2hmc::layered::GemmArgs;
using l2hmc::layered::LayDims;
using l2hmc::layered::LayState;

inline void lay_setup_dims(l2hmc_ctx *ctx) {
  LayDims &dm = ctx->lay.dm;
  const Shape &sh = ctx->sh;
  dm.D = sh.D;
  dm.Dp = round_up(sh.D, 8);
  dm.K1p = round_up(2 * sh.D, 8);
  dm.H = sh.H;
  dm.Hp = round_up(sh.H, 8);
  dm.N3p = round_up(3 * sh.D, 8);
  dm.T = sh.T;
  dm.aux = dm.auxp = 0;
  dm.ldm = sh.DP;
}

// One S/T/Q net in the layered layout: Wemb [K1p][Hp] (rows 0..D-1 embed_1/W, D..2D-1 embed_2/W), tb [T][Hp] the
// folded time-embedding bias, W4 [Hp][Hp], b4 [Hp], Wh [Hp][N3p] with columns [S | T | Q], bh [N3p], es/eq [Dp].
// Pre-split (tf32 hi / lo) and pre-tile a row-major weight [K][ldb] for tc_gemm_kernel; keyed by its device pointer.
int lay_tc_register(l2hmc_ctx *ctx, const float *dev_B, const float *host_B, int ldb, int K, int N) {
  LayeredCtx &L = ctx->lay;
  LayTcWeight &w = L.tcw[dev_B];
  // the tf32 image always (fp32 exponent range); the fp16 image beside it when the fp16 split is enabled and every entry
  // of this weight is well inside the fp16 range -- lay_gemm picks per launch (status word of the context)
  {
    std::vector<float> pk;
    l2hmc::tcg::pack_b(host_B, ldb, K, N, pk, &w.d, false);
    int rc = ensure(ctx, w.buf, pk.size());
    if (rc) return rc;
    CUDA_TRY(ctx, cudaMemcpy(w.buf.p, pk.data(), pk.size() * sizeof(float), cudaMemcpyHostToDevice));
    w.d.pk = w.buf.p;
  }
  w.has16 = false;
  if (L.gemm_f16) {
    float wmax = 0.f;
    for (int k = 0; k < K; ++k)
      for (int n = 0; n < N; ++n) wmax = fmaxf(wmax, fabsf(host_B[(size_t)k * ldb + n]));
    if (wmax < 3.0e4f) {
      std::vector<float> pk;
      l2hmc::tcg::pack_b(host_B, ldb, K, N, pk, &w.d16, true);
      int rc = ensure(ctx, w.buf16, pk.size());
      if (rc) return rc;
      CUDA_TRY(ctx, cudaMemcpy(w.buf16.p, pk.data(), pk.size() * sizeof(float), cudaMemcpyHostToDevice));
      w.d16.pk = w.buf16.p;
      w.d16.status = ctx->status_d;
      w.has16 = true;
    }
  }
  return L2HMC_OK;
}

int lay_pack_net(l2hmc_ctx *ctx, int net_id, const l2hmc_net_params *p) {
  const LayDims &dm = ctx->lay.dm;
  const int D = dm.D, H = dm.H, Hp = dm.Hp, T = dm.T;
  const size_t nWemb = (size_t)dm.K1p * Hp, ntb = (size_t)T * Hp, nW4 = (size_t)Hp * Hp, nb4 = Hp,
               nWh = (size_t)Hp * dm.N3p, nbh = dm.N3p, nes = dm.Dp;
  std::vector<float> pk(nWemb + ntb + nW4 + nb4 + nWh + nbh + 2 * nes, 0.f);
  float *Wemb = pk.data(), *tb = Wemb + nWemb, *W4 = tb + ntb, *b4 = W4 + nW4, *Wh = b4 + nb4, *bh = Wh + nWh,
        *es = bh + nbh, *eq = es + nes;
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < H; ++j) {
      Wemb[(size_t)i * Hp + j] = p->W1[(size_t)i * H + j];
      Wemb[(size_t)(D + i) * Hp + j] = p->W2[(size_t)i * H + j];
    }
  for (int t = 0; t < T; ++t) {
    const float arg = 6.2831855f * (float)t / (float)T;  // utils/dynamics.py:99-105 in fp32
    const float ct = cosf(arg), st = sinf(arg);
    for (int j = 0; j < H; ++j)
      tb[(size_t)t * Hp + j] = (p->b1[j] + p->b2[j]) + (fmaf(st, p->W3[H + j], ct * p->W3[j]) + p->b3[j]);
  }
  for (int i = 0; i < H; ++i) {
    for (int j = 0; j < H; ++j) W4[(size_t)i * Hp + j] = p->W4[(size_t)i * H + j];
    for (int d = 0; d < D; ++d) {
      Wh[(size_t)i * dm.N3p + d] = p->Ws[(size_t)i * D + d];
      Wh[(size_t)i * dm.N3p + D + d] = p->Wt[(size_t)i * D + d];
      Wh[(size_t)i * dm.N3p + 2 * D + d] = p->Wq[(size_t)i * D + d];
    }
  }
  for (int j = 0; j < H; ++j) b4[j] = p->b4[j];
  for (int d = 0; d < D; ++d) {
    bh[d] = p->bs[d];
    bh[D + d] = p->bt[d];
    bh[2 * D + d] = p->bq[d];
    es[d] = expf(p->scale_s[d]);  // utils/layers.py:84
    eq[d] = expf(p->scale_q[d]);
  }
  int rc = ensure(ctx, ctx->lay.net_buf[net_id], pk.size());
  if (rc) return rc;
  CUDA_TRY(ctx, cudaMemcpy(ctx->lay.net_buf[net_id].p, pk.data(), pk.size() * sizeof(float), cudaMemcpyHostToDevice));
  LayNetView &v = ctx->lay.net[net_id];
  v.Wemb = ctx->lay.net_buf[net_id].p;
  v.tb = v.Wemb + nWemb;
  v.W4 = v.tb + ntb;
  v.b4 = v.W4 + nW4;
  v.Wh = v.b4 + nb4;
  v.bh = v.Wh + nWh;
  v.es = v.bh + nbh;
  v.eq = v.es + nes;
  if ((rc = lay_tc_register(ctx, v.Wemb, Wemb, Hp, dm.K1p, Hp))) return rc;
  if ((rc = lay_tc_register(ctx, v.W4, W4, Hp, Hp, Hp))) return rc;
  if ((rc = lay_tc_register(ctx, v.Wh, Wh, dm.N3p, Hp, dm.N3p))) return rc;
  return L2HMC_OK;
}

// Linear / softplus stack: per layer W [wp_i][wp_{i+1}], its transpose [wp_{i+1}][wp_i] (reverse mode), b [wp_{i+1}].
int lay_pack_mlp(l2hmc_ctx *ctx, LayMlp &m, int n_layers, const int32_t *widths, const float *const *W,
                 const float *const *b, const char *who) {
  if (n_layers < 1 || n_layers > 8) return fail(ctx, L2HMC_EUNSUPPORTED, "%s: 1 <= n_layers <= 8", who);
  if (!widths || !W || !b) return fail(ctx, L2HMC_EINVAL, "%s: null argument", who);
  m.w.assign(widths, widths + n_layers + 1);
  m.wp.resize(n_layers + 1);
  size_t total = 0;
  for (int i = 0; i <= n_layers; ++i) {
    if (m.w[i] < 1 || m.w[i] > 65536) return fail(ctx, L2HMC_EINVAL, "%s: width %d out of range", who, m.w[i]);
    m.wp[i] = round_up(m.w[i], 8);
  }
  for (int i = 0; i < n_layers; ++i) {
    if (!W[i] || !b[i]) return fail(ctx, L2HMC_EINVAL, "%s: null weight pointer", who);
    total += 2 * (size_t)m.wp[i] * m.wp[i + 1] + m.wp[i + 1];
  }
  std::vector<float> pk(total, 0.f);
  std::vector<size_t> oW(n_layers), oT(n_layers), ob(n_layers);
  size_t o = 0;
  for (int i = 0; i < n_layers; ++i) {
    const int wi = m.w[i], wo = m.w[i + 1], pi = m.wp[i], po = m.wp[i + 1];
    oW[i] = o; o += (size_t)pi * po;
    oT[i] = o; o += (size_t)po * pi;
    ob[i] = o; o += po;
    for (int r = 0; r < wi; ++r)
      for (int c = 0; c < wo; ++c) {
        const float v = W[i][(size_t)r * wo + c];
        if (!isfinite(v)) return fail(ctx, L2HMC_EINVAL, "%s: non-finite weight", who);
        pk[oW[i] + (size_t)r * po + c] = v;
        pk[oT[i] + (size_t)c * pi + r] = v;
      }
    for (int c = 0; c < wo; ++c) {
      if (!isfinite(b[i][c])) return fail(ctx, L2HMC_EINVAL, "%s: non-finite bias", who);
      pk[ob[i] + c] = b[i][c];
    }
  }
  int rc = ensure(ctx, m.buf, total);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaMemcpy(m.buf.p, pk.data(), total * sizeof(float), cudaMemcpyHostToDevice));
  m.W.resize(n_layers);
  m.Wt.resize(n_layers);
  m.b.resize(n_layers);
  for (int i = 0; i < n_layers; ++i) {
    m.W[i] = m.buf.p + oW[i];
    m.Wt[i] = m.buf.p + oT[i];
    m.b[i] = m.buf.p + ob[i];
    if ((rc = lay_tc_register(ctx, m.W[i], pk.data() + oW[i], m.wp[i + 1], m.wp[i], m.wp[i + 1]))) return rc;
    if ((rc = lay_tc_register(ctx, m.Wt[i], pk.data() + oT[i], m.wp[i], m.wp[i + 1], m.wp[i]))) return rc;
  }
  m.n_layers = n_layers;
  return L2HMC_OK;
}

int ensure_zero(l2hmc_ctx *ctx, DevBuf &b, size_t n) {
  if (b.n >= n && b.p) return L2HMC_OK;
  int rc = ensure(ctx, b, n);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaMemset(b.p, 0, n * sizeof(float)));
  return L2HMC_OK;
}

// Workspace for n chains.  Buffers whose pad columns are read as GEMM K columns are zero-initialised and
// only their first D / 2D / aux columns are ever written.
// The workspace belongs to the context, so calls that arrive on different streams must not overlap: every user
// first makes its stream wait for the previous user's release event (lay_ensure_ws) and records a new one when
// its last kernel is enqueued (lay_release).  Same-stream calls are ordered anyway.
int lay_release(l2hmc_ctx *ctx, cudaStream_t s) {
  LayeredCtx &L = ctx->lay;
  if (!L.ws_event) CUDA_TRY(ctx, cudaEventCreateWithFlags(&L.ws_event, cudaEventDisableTiming));
  CUDA_TRY(ctx, cudaEventRecord(L.ws_event, s));
  L.ws_recorded = true;
  return L2HMC_OK;
}

int lay_ensure_ws(l2hmc_ctx *ctx, long long n, cudaStream_t s) {
  LayeredCtx &L = ctx->lay;
  const LayDims &dm = L.dm;
  int rc;
  if (L.ws_recorded) CUDA_TRY(ctx, cudaStreamWaitEvent(s, L.ws_event, 0));
  const size_t N = (size_t)n;
  if ((rc = ensure_zero(ctx, L.x, N * dm.Dp))) return rc;
  if ((rc = ensure_zero(ctx, L.v, N * dm.Dp))) return rc;
  if ((rc = ensure_zero(ctx, L.x0, N * dm.Dp))) return rc;
  if ((rc = ensure_zero(ctx, L.ab, N * dm.K1p))) return rc;
  if ((rc = ensure(ctx, L.hd, N * dm.N3p))) return rc;
  if ((rc = ensure(ctx, L.hA, N * dm.Hp))) return rc;
  if ((rc = ensure(ctx, L.hB, N * dm.Hp))) return rc;
  if ((rc = ensure(ctx, L.vec, N * 8))) return rc;  // logj, h0, U, u, dir, acc, U0, U1
  if (L.enc.n_layers > 0) {
    if ((rc = ensure(ctx, L.eaux, N * dm.Hp))) return rc;
    L.eact.resize(L.enc.n_layers);
    for (int i = 1; i < L.enc.n_layers; ++i)
      if ((rc = ensure(ctx, L.eact[i], N * L.enc.wp[i]))) return rc;
  }
  if (ctx->en.kind == L2HMC_ENERGY_DECODER) {
    L.dact.resize(L.dec.n_layers + 1);
    for (int i = 1; i <= L.dec.n_layers; ++i)
      if ((rc = ensure(ctx, L.dact[i], N * L.dec.wp[i]))) return rc;
  }
  if (dm.aux != dm.auxp)
    if ((rc = ensure_zero(ctx, L.auxp, N * dm.auxp))) return rc;
  return L2HMC_OK;
}

LayState lay_state(l2hmc_ctx *ctx, long long n) {
  LayeredCtx &L = ctx->lay;
  LayState st;
  st.x = L.x.p; st.v = L.v.p; st.x0 = L.x0.p; st.ab = L.ab.p; st.hd = L.hd.p;
  st.logj = L.vec.p; st.h0 = st.logj + n; st.U = st.h0 + n; st.u = st.U + n;
  st.dir = reinterpret_cast<int *>(st.u + n);
  st.acc = st.dir + n;
  return st;
}

int lay_gemm(l2hmc_ctx *ctx, cudaStream_t s, GemmArgs g) {
  if (g.M <= 0 || g.N <= 0) return L2HMC_OK;
  if ((g.M + 127) / 128 > 65535) return fail(ctx, L2HMC_EUNSUPPORTED, "layered engine: more than 8.3M chains per call");
  g.vec = ((g.ldc % 4) == 0 && (g.N % 4) == 0 && (reinterpret_cast<uintptr_t>(g.C) % 16) == 0) ? 1 : 0;
  if (ctx->lay.gemm_tc) {
    // tensor-core path: the weight must have been registered, bias / C rows 16-byte aligned
    auto it = ctx->lay.tcw.find(g.B);
    const bool aligned = (reinterpret_cast<uintptr_t>(g.bias) % 16) == 0 && (reinterpret_cast<uintptr_t>(g.bias_b) % 16) == 0 &&
                         (reinterpret_cast<uintptr_t>(g.A) % 16) == 0 && (g.lda % 4) == 0;
    if (it != ctx->lay.tcw.end() && aligned) {
      // fp16 operand split unless this context has met an activation outside the fp16 range (sticky status bit, polled
      // from pinned host memory without synchronising) or the weight itself is out of range
      const bool imgs = g.a_img != nullptr || g.c_img != nullptr;  // decided by the caller under the same conditions
      const bool f16 = imgs || (ctx->lay.gemm_f16 && it->second.has16 && !(*(volatile unsigned int *)ctx->status_h & STATUS_F16_RANGE));
      ctx->lay.used_f16 = f16;
      if (g.a_img != nullptr && ctx->lay.presplit_mode == 2)  // 128-row tiles, two accumulators: the epilogue overlaps the MMAs
        CUDA_TRY(ctx, l2hmc::tcg::launch_tc_gemm_pre(g, it->second.d16, ctx->lay.sms, s));
      else
        CUDA_TRY(ctx, l2hmc::tcg::launch_tc_gemm(g, f16 ? it->second.d16 : it->second.d, ctx->lay.sms, s));
      ctx->launches++;
      return L2HMC_OK;
    }
  }
  if (g.a_img || g.c_img || g.no_c) return fail(ctx, L2HMC_EINVAL, "layered engine: operand images need the tensor-core GEMM");
  const int bn8 = round_up(g.N, 128), bn4 = round_up(g.N, 64);
  const unsigned my = (unsigned)((g.M + 127) / 128);
  // Measured on B200 (profiles/r01_vae_launches.txt): both tile widths run at 46-50% of the FMA peak per padded
  // column, so the one that pads N less wins (ties -> the wide tile).
  if (bn8 <= bn4) l2hmc::layered::sgemm_kernel<8><<<dim3(bn8 / 128, my), 256, 0, s>>>(g);
  else l2hmc::layered::sgemm_kernel<4><<<dim3(bn4 / 64, my), 256, 0, s>>>(g);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  return L2HMC_OK;
}

GemmArgs gemm_args(const float *A, int lda, const float *B, int ldb, float *C, int ldc, long long M, int N, int K,
                   const float *bias, int epi) {
  GemmArgs g;
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.Bn = ldb; g.C = C; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K;
  g.bias = bias; g.bias_b = bias; g.dir = nullptr; g.R = nullptr; g.ldr = 0; g.scale = 1.f; g.epi = epi; g.vec = 0;
  return g;
}

#define WGRID(n) (unsigned)(((n) * 32 + 255) / 256), 256

// aux rows as a GEMM operand (row stride a multiple of 8 floats)
int lay_aux_operand(l2hmc_ctx *ctx, cudaStream_t s, long long n, const float *aux, const float **out, int *ld) {
  LayeredCtx &L = ctx->lay;
  const LayDims &dm = L.dm;
  if (dm.aux == dm.auxp && (reinterpret_cast<uintptr_t>(aux) % 16) == 0) {
    *out = aux;
    *ld = dm.aux;
    return L2HMC_OK;
  }
  int rc = ensure_zero(ctx, L.auxp, (size_t)n * dm.auxp);
  if (rc) return rc;
  const long long tot = n * dm.aux;
  l2hmc::layered::k_lay_copy_rows<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(aux, dm.aux, L.auxp.p, dm.auxp, dm.aux, n);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  *out = L.auxp.p;
  *ld = dm.auxp;
  return L2HMC_OK;
}

// eaux = enc(aux) [n][Hp]  (mnist_vae.py:134-140,149): once per call, aux does not change along a trajectory
int lay_encode_aux(l2hmc_ctx *ctx, cudaStream_t s, long long n, const float *aux) {
  LayeredCtx &L = ctx->lay;
  const LayMlp &m = L.enc;
  const float *A;
  int lda;
  int rc = lay_aux_operand(ctx, s, n, aux, &A, &lda);
  if (rc) return rc;
  for (int i = 0; i < m.n_layers; ++i) {
    const bool last = (i + 1 == m.n_layers);
    float *C = last ? L.eaux.p : L.eact[i + 1].p;
    GemmArgs g = gemm_args(A, lda, m.W[i], m.wp[i + 1], C, m.wp[i + 1], n, m.wp[i + 1], m.wp[i], m.b[i],
                           last ? l2hmc::layered::EPI_BIAS : l2hmc::layered::EPI_SOFTPLUS);
    if ((rc = lay_gemm(ctx, s, g))) return rc;
    A = C;
    lda = m.wp[i + 1];
  }
  return L2HMC_OK;
}

// The fp16 x3 tensor-core GEMMs with pre-split operand images are usable for this context right now
bool lay_presplit_ok(l2hmc_ctx *ctx) {
  const LayeredCtx &L = ctx->lay;
  return L.gemm_tc && L.gemm_f16 && L.presplit && !(*(volatile unsigned int *)ctx->status_h & STATUS_F16_RANGE);
}
bool lay_has16(l2hmc_ctx *ctx, const float *dev_B) {
  auto it = ctx->lay.tcw.find(dev_B);
  return it != ctx->lay.tcw.end() && it->second.has16;
}
uint8_t *lay_img(DevBuf &b) { return reinterpret_cast<uint8_t *>(b.p); }

// fp32 rows -> operand image (for A operands no GEMM epilogue produces: the state rows x and [a | b])
int lay_split(l2hmc_ctx *ctx, cudaStream_t s, const float *src, int ld, int K, long long n, DevBuf &img) {
  int rc = ensure_zero(ctx, img, l2hmc::layered::SplitImage::bytes(n, K) / sizeof(float));
  if (rc) return rc;
  l2hmc::layered::k_lay_split<<<(unsigned)((n + 127) / 128), 256, 0, s>>>(src, ld, K, n, lay_img(img), l2hmc::layered::SplitImage::nmb(n),
                                                                          ctx->status_d);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  return L2HMC_OK;
}

// U(x) -> st.U and (want_grad) grad U(x) -> ab[:, D:2D], for the chains' current st.x
int lay_energy_grad(l2hmc_ctx *ctx, cudaStream_t s, long long n, const float *aux, int want_grad) {
  LayeredCtx &L = ctx->lay;
  const LayDims &dm = L.dm;
  LayState st = lay_state(ctx, L.ws_n);
  if (ctx->en.kind != L2HMC_ENERGY_DECODER) {
    l2hmc::layered::k_lay_grad_generic<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(dm, st, ctx->en, ctx->sh, n, want_grad);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return L2HMC_OK;
  }
  const LayMlp &m = L.dec;
  const int nl = m.n_layers;
  int rc;
  // Operand images (layered::SplitImage): with the fp16 x3 tensor-core GEMMs every activation / gradient that is the A
  // operand of the next GEMM is written ONCE, already split, by the kernel that produces it (GEMM epilogue or k_lay_bce) and
  // fetched by TMA; otherwise each of the next GEMM's N / 256 column tiles would convert it again.
  bool pre = lay_presplit_ok(ctx) && want_grad;
  for (int i = 0; pre && i < nl; ++i) pre = lay_has16(ctx, m.W[i]) && lay_has16(ctx, m.Wt[i]);
  const int nmb = l2hmc::layered::SplitImage::nmb(n);
  if (pre) {
    L.aimg.resize(nl + 1);
    L.gimg.resize(nl + 1);
    for (int i = 1; i <= nl; ++i) {
      const size_t fl = l2hmc::layered::SplitImage::bytes(n, m.wp[i]) / sizeof(float);
      if (i < nl && (rc = ensure_zero(ctx, L.aimg[i], fl))) return rc;
      if ((rc = ensure_zero(ctx, L.gimg[i], fl))) return rc;
    }
  }
  auto img = [](DevBuf &b) { return reinterpret_cast<uint8_t *>(b.p); };
  const float *A = st.x;
  int lda = dm.Dp;
  if (pre && (rc = lay_split(ctx, s, st.x, dm.Dp, dm.Dp, n, L.ximg))) return rc;
  for (int i = 0; i < nl; ++i) {
    const bool last = (i + 1 == nl);
    GemmArgs g = gemm_args(A, lda, m.W[i], m.wp[i + 1], L.dact[i + 1].p, m.wp[i + 1], n, m.wp[i + 1], m.wp[i], m.b[i],
                           last ? l2hmc::layered::EPI_BIAS : l2hmc::layered::EPI_SOFTPLUS);
    if (pre) {
      g.img_nmb = nmb;
      g.a_img = i >= 1 ? img(L.aimg[i]) : img(L.ximg);
      if (!last) g.c_img = img(L.aimg[i + 1]);
    }
    if ((rc = lay_gemm(ctx, s, g))) return rc;
    A = L.dact[i + 1].p;
    lda = m.wp[i + 1];
  }
  if (pre)
    l2hmc::layered::k_lay_bce_img<<<(unsigned)((n + 63) / 64), 256, 0, s>>>(dm, st, L.dact[nl].p, m.wp[nl], aux, 1.0f / ctx->en.temperature,
                                                                            L.like_scale, n, img(L.gimg[nl]), nmb);
  else
    l2hmc::layered::k_lay_bce<<<WGRID(n), 0, s>>>(dm, st, L.dact[nl].p, m.wp[nl], aux, 1.0f / ctx->en.temperature, L.like_scale, n);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  if (!want_grad) return L2HMC_OK;
  // reverse mode: d <- (d W_i^T) (.) softplus'(pre_{i-1}), in place on the stored activations
  for (int i = nl - 1; i >= 1; --i) {
    GemmArgs g = gemm_args(L.dact[i + 1].p, m.wp[i + 1], m.Wt[i], m.wp[i], L.dact[i].p, m.wp[i], n, m.wp[i], m.wp[i + 1],
                           nullptr, l2hmc::layered::EPI_DSOFTPLUS);
    if (pre) {  // gradient images only: the fp32 buffer keeps the activation (read by the softplus' factor)
      g.img_nmb = nmb;
      g.a_img = img(L.gimg[i + 1]);
      g.c_img = img(L.gimg[i]);
      g.no_c = 1;
    }
    if ((rc = lay_gemm(ctx, s, g))) return rc;
  }
  GemmArgs g = gemm_args(L.dact[1].p, m.wp[1], m.Wt[0], m.wp[0], st.ab + dm.D, dm.K1p, n, dm.D, m.wp[1], nullptr,
                         l2hmc::layered::EPI_ADD_SCALE);
  if (pre) {
    g.img_nmb = nmb;
    g.a_img = img(L.gimg[1]);
  }
  g.R = st.x;  // + z: gradient of the standard normal prior (mnist_vae.py:125)
  g.ldr = dm.Dp;
  g.scale = 1.0f / ctx->en.temperature;
  return lay_gemm(ctx, s, g);
}

// hd = raw [S | T | Q] of net([ab, t, aux]) for every chain; time index it (forward chains) / T-1-it (backward)
int lay_net_call(l2hmc_ctx *ctx, cudaStream_t s, int net_id, int it, long long n, const int *dir, const float *tbias) {
  LayeredCtx &L = ctx->lay;
  const LayDims &dm = L.dm;
  const LayNetView &w = L.net[net_id];
  LayState st = lay_state(ctx, L.ws_n);
  int rc;
  // operand images: [a | b] split once, the two hidden activations leave their GEMMs as images only (no fp32 copy)
  const bool pre = lay_presplit_ok(ctx) && lay_has16(ctx, w.Wemb) && lay_has16(ctx, w.W4) && lay_has16(ctx, w.Wh);
  const int nmb = l2hmc::layered::SplitImage::nmb(n);
  if (pre) {
    if ((rc = lay_split(ctx, s, st.ab, dm.K1p, dm.K1p, n, L.abimg))) return rc;
    const size_t fl = l2hmc::layered::SplitImage::bytes(n, dm.Hp) / sizeof(float);
    if ((rc = ensure_zero(ctx, L.hAimg, fl)) || (rc = ensure_zero(ctx, L.hBimg, fl))) return rc;
  }
  const float *bias_f = tbias ? tbias : w.tb + (size_t)it * dm.Hp;
  const float *bias_b = tbias ? tbias : w.tb + (size_t)(dm.T - 1 - it) * dm.Hp;
  if (pre && L.fused_net) {
    // the three GEMMs as one kernel (tc_net.cuh): the hidden activations stay on the SM
    l2hmc::tcg::NetFusedArgs p;
    p.a_img = lay_img(L.abimg); p.img_nmb = nmb; p.M = n;
    p.w1 = L.tcw[w.Wemb].d16; p.w2 = L.tcw[w.W4].d16; p.w3 = L.tcw[w.Wh].d16;
    p.N1 = dm.Hp; p.N3 = dm.N3p;
    p.bias1 = bias_f; p.bias1_b = bias_b; p.dir = dir;
    p.R = L.enc.n_layers > 0 ? L.eaux.p : nullptr; p.ldr = dm.Hp;
    p.bias2 = w.b4; p.bias3 = w.bh; p.hd = st.hd; p.ldc = dm.N3p; p.status = ctx->status_d;
    const bool aligned = ((reinterpret_cast<uintptr_t>(bias_f) | reinterpret_cast<uintptr_t>(bias_b) | reinterpret_cast<uintptr_t>(w.b4) |
                           reinterpret_cast<uintptr_t>(w.bh)) % 16) == 0;
    if (aligned && l2hmc::tcg::tc_net_fits(p, 227 * 1024)) {
      CUDA_TRY(ctx, l2hmc::tcg::launch_tc_net(p, L.sms, s));
      ctx->launches++;
      L.used_f16 = true;
      return L2HMC_OK;
    }
  }
  GemmArgs g = gemm_args(st.ab, dm.K1p, w.Wemb, dm.Hp, L.hA.p, dm.Hp, n, dm.Hp, dm.K1p, bias_f, l2hmc::layered::EPI_RELU);
  g.bias_b = bias_b;
  g.dir = dir;
  if (L.enc.n_layers > 0) {
    g.R = L.eaux.p;
    g.ldr = dm.Hp;
  }
  if (pre) { g.img_nmb = nmb; g.a_img = lay_img(L.abimg); g.c_img = lay_img(L.hAimg); g.no_c = 1; }
  if ((rc = lay_gemm(ctx, s, g))) return rc;
  g = gemm_args(L.hA.p, dm.Hp, w.W4, dm.Hp, L.hB.p, dm.Hp, n, dm.Hp, dm.Hp, w.b4, l2hmc::layered::EPI_RELU);
  if (pre) { g.img_nmb = nmb; g.a_img = lay_img(L.hAimg); g.c_img = lay_img(L.hBimg); g.no_c = 1; }
  if ((rc = lay_gemm(ctx, s, g))) return rc;
  g = gemm_args(L.hB.p, dm.Hp, w.Wh, dm.N3p, st.hd, dm.N3p, n, dm.N3p, dm.Hp, w.bh, l2hmc::layered::EPI_BIAS);
  if (pre) { g.img_nmb = nmb; g.a_img = lay_img(L.hBimg); }
  return lay_gemm(ctx, s, g);
}

int launch_layered(l2hmc_ctx *ctx, const l2hmc_transition_args *a, const TransitionIO &io, cudaStream_t s) {
  LayeredCtx &L = ctx->lay;
  const LayDims &dm = L.dm;
  const long long n = a->n;
  const bool needs_aux = ctx->en.kind == L2HMC_ENERGY_DECODER || L.enc.n_layers > 0;
  if (needs_aux && !a->aux) return fail(ctx, L2HMC_EINVAL, "l2hmc_transition: this target / these nets need aux rows");
  int rc = lay_ensure_ws(ctx, n, s);
  if (rc) return rc;
  L.ws_n = n;
  LayState st = lay_state(ctx, n);
  const int hmc = ctx->sh.hmc;
  const float eps = ctx->sh.eps;
  if (L.enc.n_layers > 0 && !hmc)
    if ((rc = lay_encode_aux(ctx, s, n, a->aux))) return rc;
  auto update = [&](int mode, int net_id, int it, int half, int build_next) -> int {
    const LayNetView &w = L.net[net_id];
    if (mode == 0)
      l2hmc::layered::k_lay_update<0><<<WGRID(n), 0, s>>>(dm, st, w.es, w.eq, ctx->mask.p, eps, it, half, build_next, hmc, n);
    else
      l2hmc::layered::k_lay_update<1><<<WGRID(n), 0, s>>>(dm, st, w.es, w.eq, ctx->mask.p, eps, it, half, build_next, hmc, n);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return L2HMC_OK;
  };
  for (int tr = 0; tr < a->n_transitions; ++tr) {
    l2hmc::layered::k_lay_begin<<<WGRID(n), 0, s>>>(dm, st, io, tr);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    if ((rc = lay_energy_grad(ctx, s, n, a->aux, 1))) return rc;
    l2hmc::layered::k_lay_add_h0<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(st, n);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    for (int it = 0; it < dm.T; ++it) {
      if (!hmc && (rc = lay_net_call(ctx, s, L2HMC_VNET, it, n, st.dir, nullptr))) return rc;
      if ((rc = update(0, L2HMC_VNET, it, 0, 1))) return rc;
      if (!hmc && (rc = lay_net_call(ctx, s, L2HMC_XNET, it, n, st.dir, nullptr))) return rc;
      if ((rc = update(1, L2HMC_XNET, it, 0, 0))) return rc;
      if (!hmc && (rc = lay_net_call(ctx, s, L2HMC_XNET, it, n, st.dir, nullptr))) return rc;
      if ((rc = update(1, L2HMC_XNET, it, 1, 0))) return rc;
      if ((rc = lay_energy_grad(ctx, s, n, a->aux, 1))) return rc;
      if (!hmc && (rc = lay_net_call(ctx, s, L2HMC_VNET, it, n, st.dir, nullptr))) return rc;
      if ((rc = update(0, L2HMC_VNET, it, 0, 0))) return rc;
    }
    l2hmc::layered::k_lay_end<<<WGRID(n), 0, s>>>(dm, st, io, tr);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
  }
  return lay_release(ctx, s);
}

}  // namespace
