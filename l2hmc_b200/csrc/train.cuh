// Training path of libl2hmc.so, FIRST-CORRECT version: value and gradient of the notebook's
// objective (SCGExperiment.ipynb:159-181; utils/losses.py:36-59) for one `propose` batch, by a hand-written reverse
// sweep through the unrolled leapfrog -- what TF1 autodiff does for the reference through tf.while_loop
// (utils/dynamics.py:246-300), tf.gradients(energy, x) (:217-218, i.e. Hessian-vector products on the way back),
// p_accept (:302-309) and the direction select of propose (utils/sampler.py:34-44).
//
// The algorithm is oracle/l2hmc_reverse.py, kernel for function:
//   forward sweep : 4 T sub-updates (V, X, X, V per leapfrog step); the state (x, v) in front of each one is recorded
//   reverse sweep : per sub-update, last to first: recompute its net forward from the record, apply the vector-Jacobian
//                   product of the update (k_update_vjp), of the S/T/Q net (GEMMs against the transposed weights, weight
//                   gradients as K = chains products, split over CTAs and reduced in a fixed order) and of grad U (k_hvp).
// Layout: chain-major unpadded fp32 rows ([n, D] states, [n, 2D] net input, [n, H] activations, [n, 3D] heads) against
// the reference-layout weights the context already holds (NetRaw).  GEMMs are a plain shared-memory fp32 FMA kernel with
// strided operands: this version is about the gradient being right, not about speed (DESIGN.md section 7.1 has the plan
// for the fused one).  Gaussian, mixture-of-Gaussians, RoughWell and funnel targets (closed-form Hessians), no aux.
#pragma once
#ifndef L2HMC_TRAIN_EMU  // tests/emu/train_emu.cpp supplies the few types it needs and runs these kernels on host threads
#include "common.cuh"
#endif

// Reductions over the chains (weight-gradient products, column sums) are split over CTAs and finished in a fixed order by
// k_reduce_add: deterministic, no atomics.  Smallest split sizes (the host grows them so that at most TR_MAX_PARTS parts
// exist per reduction); the CPU emulation shrinks both.
#ifndef L2HMC_TR_KCHUNK
#define L2HMC_TR_KCHUNK 128   // chains per CTA of a weight-gradient product (split K)
#endif
#ifndef L2HMC_TR_SLAB
#define L2HMC_TR_SLAB 64      // rows per CTA of a column sum
#endif
#define L2HMC_TR_MAX_PARTS 64

namespace l2hmc {
namespace train {

// C[m][n] (=, +=, or per-part slices) sum_k A(m,k) B(k,n);  A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn]
struct Gemm {
  const float *A; long long sam, sak;
  const float *B; long long sbk, sbn;
  float *C; long long ldc;
  long long M; int N; long long K;
  int mode;          // 0 store, 1 add (one writer per element), 2 split K: part blockIdx.z stores its [M][N] slice of C (ldc = N)
  long long kchunk;  // K range per blockIdx.z
};

__global__ void __launch_bounds__(256) k_gemm(const Gemm g) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 1];
  __shared__ float Bs[BK][BN + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const long long k_begin = (long long)blockIdx.z * g.kchunk;
  const long long k_end = (k_begin + g.kchunk < g.K) ? k_begin + g.kchunk : g.K;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_kfast = (g.sak == 1), b_nfast = (g.sbn == 1);
  for (long long k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = tid + 256 * r;  // 1024 elements of each tile
      int mm, kk;
      if (a_kfast) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
      const long long m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < g.M && k < k_end) ? g.A[m * g.sam + k * g.sak] : 0.f;
      int nn, kb;
      if (b_nfast) { nn = e & 63; kb = e >> 6; } else { kb = e & 15; nn = e >> 4; }
      const long long kq = k0 + kb;
      const int n = n0 + nn;
      Bs[kb][nn] = (n < g.N && kq < k_end) ? g.B[kq * g.sbk + (long long)n * g.sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float *c = g.C + m * g.ldc + n;
      if (g.mode == 0) *c = acc[i][j];
      else if (g.mode == 1) *c += acc[i][j];
      else c[(long long)blockIdx.z * g.M * g.ldc] = acc[i][j];
    }
  }
}

// part[blockIdx.y][c] = sum over the slab's rows of w[r] * A[r*lda + c]  (w == null: plain column sums)
__global__ void k_colsum(const float *A, long long lda, long long n_rows, int n_cols, const float *w, float *part, long long slab) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cols) return;
  const long long r0 = (long long)blockIdx.y * slab;
  const long long r1 = (r0 + slab < n_rows) ? r0 + slab : n_rows;
  float s = 0.f;
  for (long long r = r0; r < r1; ++r) s = fmaf(w ? w[r] : 1.f, A[r * lda + c], s);
  part[(long long)blockIdx.y * n_cols + c] = s;
}

// dst[m*ldc + n] += part[0][m][n] + part[1][m][n] + ... (parts added in index order: the same bits on every run)
__global__ void k_reduce_add(const float *part, int n_parts, long long M, int N, float *dst, long long ldc) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  float s = 0.f;
  for (int z = 0; z < n_parts; ++z) s += part[(long long)z * M * N + i];
  const long long m = i / N;
  dst[m * ldc + (i - m * N)] += s;
}

// per chain: the step index it is at and its time features (utils/dynamics.py:99-105; backward chains count down, :285)
__global__ void k_tau(long long n, const uint8_t *dir, int it, int T, float *ct, float *st) {
  const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g >= n) return;
  const int t = dir[g] ? it : T - 1 - it;
  const float arg = 6.2831855f * (float)t / (float)T;
  ct[g] = cosf(arg);
  st[g] = sinf(arg);
}

// h1 = relu(z1 + b1 + b2 + b3 + tau W3), in place
__global__ void k_act1(long long n, int H, float *z, const float *b1, const float *b2, const float *b3, const float *W3,
                       const float *ct, const float *st) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n * H) return;
  const long long g = i / H;
  const int j = (int)(i - g * H);
  const float e3 = fmaf(st[g], W3[H + j], ct[g] * W3[j]);
  z[i] = fmaxf(z[i] + b1[j] + b2[j] + (e3 + b3[j]), 0.f);
}
__global__ void k_act2(long long n, int H, float *z, const float *b4) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n * H) return;
  z[i] = fmaxf(z[i] + b4[(int)(i % H)], 0.f);
}
// g <- g * [h > 0]
__global__ void k_relu_mask(long long tot, float *g, const float *h) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < tot && !(h[i] > 0.f)) g[i] = 0.f;
}

// net input [a | b]: V update: [x | grad U(x)]; X update: [v | keep * x]
// which: 0 = V, 1 = first X of the step, 2 = second X
__device__ __forceinline__ float keep_of(int which, bool fwd, float m) {
  // forward: first X keeps m, second keeps 1 - m (utils/dynamics.py:131,140); backward: the other way round (:173,182)
  const bool keep_m = (which == 1) == fwd;
  return keep_m ? m : 1.f - m;
}
__global__ void k_build_ab(long long n, int D, int DP, int T, int it, int which, const uint8_t *dir, const float *mask,
                           const float *x, const float *v, const float *gU, float *ab) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n * D) return;
  const long long g = i / D;
  const int d = (int)(i - g * D);
  float a, b;
  if (which == 0) {
    a = x[i];
    b = gU[i];
  } else {
    const bool fwd = dir[g] != 0;
    const int t = fwd ? it : T - 1 - it;
    a = v[i];
    b = keep_of(which, fwd, mask[(size_t)t * DP + d]) * x[i];
  }
  ab[g * 2 * D + d] = a;
  ab[g * 2 * D + D + d] = b;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct Heads {  // raw head parameters of one net (reference layout)
  const float *bs, *bt, *bq, *ls, *lq;
};

// One sub-update, forward (utils/dynamics.py:117-128,146-153 / :131-144 and their inverses :161-199); one warp per
// chain; x or v updated in place, logj[n] += the sub-update's log|J| row sum.
__global__ void k_update(long long n, int D, int DP, int T, int it, int which, const uint8_t *dir, const float *mask,
                         Heads hp, float eps, const float *hd, const float *gU, float *x, float *v, float *logj) {
  const long long g = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= n) return;
  const bool fwd = dir[g] != 0;
  const float sign = fwd ? 1.f : -1.f;
  const int t = fwd ? it : T - 1 - it;
  float lj = 0.f;
  for (int d = lane; d < D; d += 32) {
    const long long i = g * D + d;
    const float S = expf(hp.ls[d]) * tanhf(hd[g * 3 * D + d] + hp.bs[d]);
    const float Tt = hd[g * 3 * D + D + d] + hp.bt[d];
    const float Q = expf(hp.lq[d]) * tanhf(hd[g * 3 * D + 2 * D + d] + hp.bq[d]);
    const float f = eps * Q;
    if (which == 0) {
      const float s = 0.5f * sign * eps * S;
      const float shift = 0.5f * eps * (-(expf(f) * gU[i]) + Tt);
      v[i] = fwd ? v[i] * expf(s) + shift : (v[i] - shift) * expf(s);
      lj += s;
    } else {
      const float keep = keep_of(which, fwd, mask[(size_t)t * DP + d]), upd = 1.f - keep;
      const float s = sign * eps * S;
      const float shift = eps * (expf(f) * v[i] + Tt);
      const float nx = fwd ? x[i] * expf(s) + shift : expf(s) * (x[i] - shift);
      x[i] = keep * x[i] + upd * nx;
      lj += upd * s;
    }
  }
  lj = warp_sum(lj);
  if (lane == 0) logj[g] += lj;
}

// Reverse of one sub-update at its recorded state (x, v) (oracle/l2hmc_reverse.py v_update_vjp / x_update_vjp).
// in/out: gx, gv [n, D] (cotangents of the sub-update's outputs -> of its inputs, the net / grad-U paths excluded);
// out: ghd [n, 3D] cotangents of the raw heads, sc [n, 2D] = (gS S | gQ Q) for the two log-scales, gg [n, D] the direct
// cotangent of grad U (V only), geps[n] += d/d eps.
__global__ void k_update_vjp(long long n, int D, int DP, int T, int it, int which, const uint8_t *dir, const float *mask,
                             Heads hp, float eps, const float *hd, const float *gU, const float *x, const float *v,
                             const float *glj, float *gx, float *gv, float *ghd, float *sc, float *gg, float *geps) {
  const long long g = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= n) return;
  const bool fwd = dir[g] != 0;
  const float sign = fwd ? 1.f : -1.f;
  const int t = fwd ? it : T - 1 - it;
  const float gl = glj[g];
  float ge = 0.f;
  for (int d = lane; d < D; d += 32) {
    const long long i = g * D + d;
    const float els = expf(hp.ls[d]), elq = expf(hp.lq[d]);
    const float ts = tanhf(hd[g * 3 * D + d] + hp.bs[d]);
    const float tq = tanhf(hd[g * 3 * D + 2 * D + d] + hp.bq[d]);
    const float S = els * ts, Q = elq * tq;
    const float Tt = hd[g * 3 * D + D + d] + hp.bt[d];
    const float f = eps * Q, ef = expf(f);
    float g_s, g_f, gT, dsde;  // dsde: d s / d eps divided by S
    if (which == 0) {
      const float s = 0.5f * sign * eps * S, es = expf(s);
      const float inner = -(ef * gU[i]) + Tt;  // shift = 0.5 eps inner
      const float gvo = gv[i];
      float g_shift;
      if (fwd) {
        g_s = gvo * v[i] * es + gl;
        g_shift = gvo;
      } else {
        const float v_o = (v[i] - 0.5f * eps * inner) * es;
        g_s = gvo * v_o + gl;
        g_shift = -gvo * es;
      }
      gv[i] = gvo * es;
      const float g_inner = g_shift * (0.5f * eps);
      ge = fmaf(g_shift * 0.5f, inner, ge);
      g_f = g_inner * (-(ef * gU[i]));
      gg[i] = g_inner * (-ef);
      gT = g_inner;
      dsde = 0.5f * sign;
    } else {
      const float keep = keep_of(which, fwd, mask[(size_t)t * DP + d]), upd = 1.f - keep;
      const float s = sign * eps * S, es = expf(s);
      const float inner = ef * v[i] + Tt;  // shift = eps inner
      const float gxo = gx[i], gu = gxo * upd;
      float g_shift;
      if (fwd) {
        g_s = gu * x[i] * es + gl * upd;
        g_shift = gu;
      } else {
        g_s = gu * es * (x[i] - eps * inner) + gl * upd;
        g_shift = -gu * es;
      }
      gx[i] = gxo * keep + gu * es;
      const float g_inner = g_shift * eps;
      ge = fmaf(g_shift, inner, ge);
      g_f = g_inner * ef * v[i];
      gv[i] += g_inner * ef;
      gT = g_inner;
      dsde = sign;
    }
    const float gS = g_s * dsde * eps;
    ge = fmaf(g_s * dsde, S, ge);
    const float gQ = g_f * eps;
    ge = fmaf(g_f, Q, ge);
    ghd[g * 3 * D + d] = gS * els * (1.f - ts * ts);
    ghd[g * 3 * D + D + d] = gT;
    ghd[g * 3 * D + 2 * D + d] = gQ * elq * (1.f - tq * tq);
    sc[g * 2 * D + d] = gS * S;
    sc[g * 2 * D + D + d] = gQ * Q;
  }
  ge = warp_sum(ge);
  if (lane == 0) geps[g] += ge;
}

// cotangents of the net input back onto the state.  V: gx += ga, w = gg + gb (then gx += Hessian-vector product of w);
// X: gv += ga, gx += keep * gb.
__global__ void k_scatter(long long n, int D, int DP, int T, int it, int which, const uint8_t *dir, const float *mask,
                          const float *gab, float *gx, float *gv, float *gg) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n * D) return;
  const long long g = i / D;
  const int d = (int)(i - g * D);
  const float ga = gab[g * 2 * D + d], gb = gab[g * 2 * D + D + d];
  if (which == 0) {
    gx[i] += ga;
    gg[i] += gb;
  } else {
    const bool fwd = dir[g] != 0;
    const int t = fwd ? it : T - 1 - it;
    gv[i] += ga;
    gx[i] += keep_of(which, fwd, mask[(size_t)t * DP + d]) * gb;
  }
}

// gx += w . d(grad U / T_emp)/dx at x  (what differentiating through tf.gradients(energy, x) yields); one thread per chain
__global__ void k_hvp(EnergyDev en, Shape sh, long long n, const float *x, const float *w, float *gx) {
  const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g >= n) return;
  const int D = sh.D;
  const float *xr = x + g * D, *wr = w + g * D;
  float *o = gx + g * D;
  if (en.kind == 0) {  // Gaussian: Hessian 0.5 (S + S^T)   utils/distributions.py:50-57
    for (int j = 0; j < D; ++j) {
      float r = 0.f;
      for (int i = 0; i < D; ++i) r = fmaf(wr[i], en.Ssym[i * sh.LDS + j], r);
      o[j] += r / en.temperature;
    }
  } else if (en.kind == 1) {
    // Mixture (utils/distributions.py:125-134): U = -logsumexp_c a_c, a_c = -q_c + log c_c, grad U = sum_c r_c g_c with
    // r = softmax(a), g_c = A_c (x - mu_c), A_c = (S_c + S_c^T) / 2.  Hessian = sum_c r_c A_c - sum_c r_c g_c g_c^T + gb gb^T,
    // gb = grad U.  Two passes so that no [components, D] array is held: (1) a_c and s_c = g_c . w, (2) the rows.
    float r[MAX_COMP], sc[MAX_COMP];
    float mx = -INFINITY;
    for (int c = 0; c < en.ncomp; ++c) {
      const float *mu = en.mu + c * sh.DP;
      const float *S = en.Ssym + (size_t)c * sh.DP * sh.LDS;
      float q = 0.f, s = 0.f;
      for (int j = 0; j < D; ++j) {
        float g = 0.f;
        for (int i = 0; i < D; ++i) g = fmaf(xr[i] - mu[i], S[i * sh.LDS + j], g);
        q = fmaf(g, xr[j] - mu[j], q);
        s = fmaf(g, wr[j], s);
      }
      r[c] = -0.5f * q + en.logc[c];
      sc[c] = s;
      mx = fmaxf(mx, r[c]);
    }
    float z = 0.f, gbw = 0.f;
    for (int c = 0; c < en.ncomp; ++c) {
      r[c] = expf(r[c] - mx);
      z += r[c];
    }
    for (int c = 0; c < en.ncomp; ++c) {
      r[c] /= z;
      gbw = fmaf(r[c], sc[c], gbw);  // grad U . w
    }
    for (int j = 0; j < D; ++j) {
      float acc = 0.f, gb = 0.f;
      for (int c = 0; c < en.ncomp; ++c) {
        const float *mu = en.mu + c * sh.DP;
        const float *S = en.Ssym + (size_t)c * sh.DP * sh.LDS;
        float g = 0.f, wa = 0.f;
        for (int i = 0; i < D; ++i) {
          g = fmaf(xr[i] - mu[i], S[i * sh.LDS + j], g);
          wa = fmaf(wr[i], S[i * sh.LDS + j], wa);
        }
        acc = fmaf(r[c], wa - g * sc[c], acc);
        gb = fmaf(r[c], g, gb);
      }
      o[j] += (acc + gb * gbw) / en.temperature;
    }
  } else if (en.kind == 2) {  // RoughWell: grad = x - e sin(x / den) / den, diagonal Hessian 1 - e cos(x / den) / den^2   :90-97
    const float e = en.s0, den = en.s1;
    for (int j = 0; j < D; ++j) o[j] += wr[j] * (1.f - e * cosf(xr[j] / den) / (den * den)) / en.temperature;
  } else {
    // GaussianFunnel (utils/distributions.py:161-180), v = x_0, s = e^v: grad = (v / sigma^2 + (n - |x_1:|^2 / s) / 2, x_1: / s)
    // inside the clip; outside it s is a constant and the v-coupling drops out (the tf.where branches)
    const float sigma = en.s0, clip = en.s1;
    const float v = xr[0];
    const bool out = (v > clip) || (-clip > v);
    const float s = out ? expf(v > clip ? clip : -clip) : expf(v);
    float ss = 0.f, wx = 0.f;
    for (int i = 1; i < D; ++i) {
      ss = fmaf(xr[i], xr[i], ss);
      wx = fmaf(wr[i], xr[i], wx);
    }
    const float h00 = 1.f / (sigma * sigma) + (out ? 0.f : 0.5f * ss / s);
    o[0] += (wr[0] * h00 - (out ? 0.f : wx / s)) / en.temperature;
    for (int j = 1; j < D; ++j) o[j] += ((wr[j] - (out ? 0.f : wr[0] * xr[j])) / s) / en.temperature;
  }
}

// ---- objective (utils/losses.py:36-59) ---------------------------------------------------------------------------
// p = exp(min(H0 - H1 + logJ, 0)), non-finite -> 0 (utils/dynamics.py:302-309); vv = |x0 - X|^2 p + 1e-4 (loss_vec).
// One warp per chain; returns sq = |x0 - X|^2 to every lane.
__device__ __forceinline__ float loss_vec_chain(long long g, int lane, int D, const float *x0, const float *X, const float *H0,
                                               const float *H1, const float *logj, float &p, bool &differentiable) {
  float sq = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float dx = x0[g * D + d] - X[g * D + d];
    sq = fmaf(dx, dx, sq);
  }
  sq = warp_sum(sq);
  const float arg = H0[g] - H1[g] + logj[g];
  p = expf(fminf(arg, 0.f));
  const bool ok = isfinite(p);
  if (!ok) p = 0.f;
  differentiable = ok && arg < 0.f;   // the min() clamp and the non-finite guard pass no gradient
  return sq;
}

// vv[n] alone, for the losses whose d loss / d v needs a statistic of the whole batch first
__global__ void k_loss_v(long long n, int D, const float *x0, const float *X, const float *H0, const float *H1,
                         const float *logj, float *vv) {
  const long long g = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= n) return;
  float p;
  bool diff;
  const float sq = loss_vec_chain(g, lane, D, x0, X, H0, H1, logj, p, diff);
  if (lane == 0) vv[g] = sq * p + 1e-4f;
}

// stats[0] = sum 1 / (v + 1e-4) (loss_inverse); stats[1] = max(-v), stats[2] = sum exp(-v - max) (loss_logsumexp).
// One block of 256 threads.
__global__ void __launch_bounds__(256) k_loss_stats(long long n, const float *vv, float *stats) {
  __shared__ float red[256];
  const int t = threadIdx.x;
  float a = 0.f, mx = -INFINITY;
  for (long long i = t; i < n; i += 256) {
    a += 1.f / (vv[i] + 1e-4f);
    mx = fmaxf(mx, -vv[i]);
  }
  red[t] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) red[t] = fmaxf(red[t], red[t + o]);
    __syncthreads();
  }
  mx = red[0];
  __syncthreads();
  float e = 0.f;
  for (long long i = t; i < n; i += 256) e += expf(-vv[i] - mx);
  red[t] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) red[t] += red[t + o];
    __syncthreads();
  }
  a = red[0];
  __syncthreads();
  red[t] = e;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) red[t] += red[t + o];
    __syncthreads();
  }
  if (t == 0) {
    stats[0] = a;
    stats[1] = mx;
    stats[2] = red[0];
  }
}

// Loss terms of the batch and the cotangents that start the reverse sweep.  kind (get_loss names, utils/losses.py:26-34):
//   0 'mixed'     scale mean(1/v) - mean(v)/scale   (:53-59; the notebook's objective, SCGExperiment.ipynb:171-181)
//   1 'standard'  -mean(v)                          (:49-51)
//   2 'inverse'   -1 / mean(1 / (v + 1e-4))         (:44-47)      needs stats[0]
//   3 'logsumexp' logsumexp(-v) - log N             (:39-42)      needs stats[1], stats[2]
// means run over 1 / inv_count chains.  in: gU1 = grad U(X) / T_emp.  out: lossv[n] (terms that sum to the loss), px[n],
// glj[n], gx = d/dX, gv = d/dV.
__global__ void k_loss(int kind, long long n, int D, const float *x0, const float *X, const float *V, const float *H0,
                       const float *H1, const float *logj, const float *gU1, const float *stats, float scale, float inv_count,
                       float *lossv, float *px, float *glj, float *gx, float *gv) {
  const long long g = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= n) return;
  float p;
  bool diff;
  const float sq = loss_vec_chain(g, lane, D, x0, X, H0, H1, logj, p, diff);
  const float vv = sq * p + 1e-4f;
  float g_v, term;
  if (kind == 0) {
    g_v = (-scale / (vv * vv) - 1.f / scale) * inv_count;
    term = (scale / vv - vv / scale) * inv_count;
  } else if (kind == 1) {
    g_v = -inv_count;
    term = -vv * inv_count;
  } else if (kind == 2) {
    const float m = stats[0] * inv_count, q = vv + 1e-4f;
    g_v = -inv_count / (m * m * q * q);
    term = g == 0 ? -1.f / m : 0.f;
  } else {
    g_v = -expf(-vv - stats[1]) / stats[2];
    term = g == 0 ? stats[1] + logf(stats[2]) + logf(inv_count) : 0.f;
  }
  const float g_arg = diff ? g_v * sq * p : 0.f;
  for (int d = lane; d < D; d += 32) {
    const long long i = g * D + d;
    gx[i] = g_v * p * 2.f * (X[i] - x0[i]) - g_arg * gU1[i];
    gv[i] = -g_arg * V[i];
  }
  if (lane == 0) {
    lossv[g] = term;
    px[g] = p;
    glj[g] = g_arg;
  }
}

}  // namespace train
}  // namespace l2hmc
