/* Plain-C restatement of the L2HMC sampling path.  TEST INFRASTRUCTURE ONLY (see l2hmc_oracle.c).
 * Included twice by l2hmc_oracle.c: REAL = float (suffix _f32) and REAL = double (suffix _f64).
 * Scalar loops in natural index order; every function cites the reference lines it follows
 * (/root/reference).  Pinned to vectors produced by the reference's own sources (tests/test_reference_pin.py). */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

/* ---- S/T/Q net: SCGExperiment.ipynb:51-77, utils/layers.py:29-37,81-95 ------------------------------ */
static void FN(net_apply)(const oracle_problem *P, const float *const *w, const REAL *a, const REAL *b,
                          REAL tcos, REAL tsin, REAL *S, REAL *T, REAL *Q) {
  const int D = P->D, H = P->H;
  if (w == NULL) { /* hmc: zeros (utils/dynamics.py:73-76) */
    for (int d = 0; d < D; ++d) S[d] = T[d] = Q[d] = 0;
    return;
  }
  REAL h1[ORACLE_MAXH], h2[ORACLE_MAXH];
  for (int j = 0; j < H; ++j) {
    REAL e1 = 0, e2 = 0, e3 = 0;
    for (int i = 0; i < D; ++i) e1 += a[i] * (REAL)w[W1][i * H + j];
    e1 += (REAL)w[B1][j];
    for (int i = 0; i < D; ++i) e2 += b[i] * (REAL)w[W2][i * H + j];
    e2 += (REAL)w[B2][j];
    e3 = tcos * (REAL)w[W3][j] + tsin * (REAL)w[W3][H + j];
    e3 += (REAL)w[B3][j];
    REAL s = (((0 + e1) + e2) + e3) + 0; /* python sum over the Zip outputs, aux branch = 0. */
    h1[j] = s > 0 ? s : 0;
  }
  for (int j = 0; j < H; ++j) {
    REAL s = 0;
    for (int i = 0; i < H; ++i) s += h1[i] * (REAL)w[W4][i * H + j];
    s += (REAL)w[B4][j];
    h2[j] = s > 0 ? s : 0;
  }
  for (int d = 0; d < D; ++d) {
    REAL s = 0, t = 0, q = 0;
    for (int i = 0; i < H; ++i) {
      s += h2[i] * (REAL)w[WS][i * D + d];
      t += h2[i] * (REAL)w[WT][i * D + d];
      q += h2[i] * (REAL)w[WQ][i * D + d];
    }
    S[d] = EXP((REAL)w[LS][d]) * TANH(s + (REAL)w[BS][d]);
    T[d] = t + (REAL)w[BT][d];
    Q[d] = EXP((REAL)w[LQ][d]) * TANH(q + (REAL)w[BQ][d]);
  }
}

/* ---- energies: utils/distributions.py ------------------------------------------------------------------ */
static REAL FN(quad)(const oracle_problem *P, int c, const REAL *x) { /* quadratic_gaussian :31-32 */
  const int D = P->D;
  const float *mu = P->mu + c * D, *S = P->S + c * D * D;
  REAL q = 0;
  for (int j = 0; j < D; ++j) {
    REAL r = 0;
    for (int i = 0; i < D; ++i) r += (x[i] - (REAL)mu[i]) * (REAL)S[i * D + j];
    q += r * (x[j] - (REAL)mu[j]);
  }
  return (REAL)0.5 * q;
}

static void FN(quad_grad)(const oracle_problem *P, int c, const REAL *x, REAL *g) { /* 0.5 d S^T + 0.5 d S */
  const int D = P->D;
  const float *mu = P->mu + c * D, *S = P->S + c * D * D;
  for (int j = 0; j < D; ++j) {
    REAL a = 0, b = 0;
    for (int i = 0; i < D; ++i) {
      a += (x[i] - (REAL)mu[i]) * (REAL)S[j * D + i];
      b += (x[i] - (REAL)mu[i]) * (REAL)S[i * D + j];
    }
    g[j] = (REAL)0.5 * a + (REAL)0.5 * b;
  }
}

static REAL FN(energy)(const oracle_problem *P, const REAL *x) { /* Dynamics.energy utils/dynamics.py:203-212 */
  const int D = P->D;
  REAL U = 0;
  switch (P->energy_kind) {
    case 0: U = FN(quad)(P, 0, x); break; /* Gaussian :50-57 */
    case 1: {                             /* GMM :125-134 */
      REAL V[ORACLE_MAXCOMP], mx = -INFINITY, s = 0;
      for (int c = 0; c < P->ncomp; ++c) {
        V[c] = -FN(quad)(P, c, x) + (REAL)P->logc[c];
        if (V[c] > mx) mx = V[c];
      }
      for (int c = 0; c < P->ncomp; ++c) s += EXP(V[c] - mx);
      U = -(LOG(s) + mx);
    } break;
    case 2: { /* RoughWell :90-97 ; s0 = eps, s1 = denominator */
      REAL n = 0, cs = 0;
      for (int i = 0; i < D; ++i) {
        n += x[i] * x[i];
        cs += COS(x[i] / (REAL)P->s1);
      }
      U = (REAL)0.5 * n + (REAL)P->s0 * cs;
    } break;
    case 3: { /* GaussianFunnel :161-180 ; s0 = sigma, s1 = clip */
      const REAL sigma = P->s0, clip = P->s1, v = x[0], two_pi = (REAL)(float)6.283185307179586;
      REAL ss = 0;
      for (int i = 1; i < D; ++i) ss += x[i] * x[i];
      const REAL n = (REAL)(D - 1), lpv = (v / sigma) * (v / sigma);
      REAL s = EXP(v);
      if (v > clip) s = EXP(clip);
      if (-clip > v) s = EXP(-clip);
      U = (REAL)0.5 * (lpv + ss / s + n * LOG(two_pi * s));
    } break;
  }
  return U / (REAL)P->temperature;
}

static void FN(grad_energy)(const oracle_problem *P, const REAL *x, REAL *g) { /* utils/dynamics.py:217-218 */
  const int D = P->D;
  switch (P->energy_kind) {
    case 0: FN(quad_grad)(P, 0, x, g); break;
    case 1: {
      REAL V[ORACLE_MAXCOMP], mx = -INFINITY, s = 0, gq[ORACLE_MAXD];
      for (int c = 0; c < P->ncomp; ++c) {
        V[c] = -FN(quad)(P, c, x) + (REAL)P->logc[c];
        if (V[c] > mx) mx = V[c];
      }
      for (int c = 0; c < P->ncomp; ++c) { V[c] = EXP(V[c] - mx); s += V[c]; }
      for (int d = 0; d < D; ++d) g[d] = 0;
      for (int c = 0; c < P->ncomp; ++c) {
        FN(quad_grad)(P, c, x, gq);
        for (int d = 0; d < D; ++d) g[d] += (V[c] / s) * gq[d];
      }
    } break;
    case 2:
      for (int i = 0; i < D; ++i) g[i] = x[i] - (REAL)P->s0 * SIN(x[i] / (REAL)P->s1) / (REAL)P->s1;
      break;
    case 3: {
      const REAL sigma = P->s0, clip = P->s1, v = x[0];
      REAL ss = 0;
      for (int i = 1; i < D; ++i) ss += x[i] * x[i];
      const REAL n = (REAL)(D - 1);
      REAL s = EXP(v), gv = v / (sigma * sigma) + (REAL)0.5 * (-ss / s + n);
      if (v > clip) { s = EXP(clip); gv = v / (sigma * sigma); }
      if (-clip > v) { s = EXP(-clip); gv = v / (sigma * sigma); }
      g[0] = gv;
      for (int i = 1; i < D; ++i) g[i] = x[i] / s;
    } break;
  }
  for (int d = 0; d < D; ++d) g[d] = g[d] / (REAL)P->temperature;
}

static REAL FN(hamiltonian)(const oracle_problem *P, const REAL *x, const REAL *v) { /* :214-215, :107-108 */
  REAL k = 0;
  for (int d = 0; d < P->D; ++d) k += v[d] * v[d];
  return FN(energy)(P, x) + (REAL)0.5 * k;
}

static void FN(time_embed)(const oracle_problem *P, int step, REAL *c, REAL *s) { /* _format_time :99-105 */
  const REAL arg = (REAL)(float)6.283185307179586 * (REAL)step / (REAL)P->T;
  *c = COS(arg);
  *s = SIN(arg);
}

/* ---- one chain, one step ------------------------------------------------------------------------------- */
static REAL FN(forward_step)(const oracle_problem *P, REAL *x, REAL *v, int step) { /* utils/dynamics.py:115-157 */
  const int D = P->D;
  const REAL eps = P->eps;
  const float *m = P->mask + step * D;
  REAL tc, ts, g[ORACLE_MAXD], S[ORACLE_MAXD], T[ORACLE_MAXD], Q[ORACLE_MAXD], in[ORACLE_MAXD], lj = 0;
  FN(time_embed)(P, step, &tc, &ts);
  FN(grad_energy)(P, x, g);
  FN(net_apply)(P, P->hmc ? NULL : P->vnet, x, g, tc, ts, S, T, Q);
  for (int d = 0; d < D; ++d) {
    const REAL sv1 = (REAL)0.5 * eps * S[d], fv1 = eps * Q[d];
    v[d] = v[d] * EXP(sv1) + (REAL)0.5 * eps * (-(EXP(fv1) * g[d]) + T[d]);
    lj += sv1;
  }
  for (int d = 0; d < D; ++d) in[d] = (REAL)m[d] * x[d];
  FN(net_apply)(P, P->hmc ? NULL : P->xnet, v, in, tc, ts, S, T, Q);
  for (int d = 0; d < D; ++d) {
    const REAL mm = m[d], mb = (REAL)1 - mm, sx1 = eps * S[d], fx1 = eps * Q[d];
    x[d] = mm * x[d] + mb * (x[d] * EXP(sx1) + eps * (EXP(fx1) * v[d] + T[d]));
    lj += mb * sx1;
  }
  for (int d = 0; d < D; ++d) in[d] = ((REAL)1 - (REAL)m[d]) * x[d];
  FN(net_apply)(P, P->hmc ? NULL : P->xnet, v, in, tc, ts, S, T, Q);
  for (int d = 0; d < D; ++d) {
    const REAL mm = m[d], mb = (REAL)1 - mm, sx2 = eps * S[d], fx2 = eps * Q[d];
    x[d] = mb * x[d] + mm * (x[d] * EXP(sx2) + eps * (EXP(fx2) * v[d] + T[d]));
    lj += mm * sx2;
  }
  FN(grad_energy)(P, x, g);
  FN(net_apply)(P, P->hmc ? NULL : P->vnet, x, g, tc, ts, S, T, Q);
  for (int d = 0; d < D; ++d) {
    const REAL sv2 = (REAL)0.5 * eps * S[d], fv2 = eps * Q[d];
    v[d] = v[d] * EXP(sv2) + (REAL)0.5 * eps * (-(EXP(fv2) * g[d]) + T[d]);
    lj += sv2;
  }
  return lj;
}

static REAL FN(backward_step)(const oracle_problem *P, REAL *x, REAL *v, int step) { /* utils/dynamics.py:159-201 */
  const int D = P->D;
  const REAL eps = P->eps;
  const float *m = P->mask + step * D;
  REAL tc, ts, g[ORACLE_MAXD], S[ORACLE_MAXD], T[ORACLE_MAXD], Q[ORACLE_MAXD], in[ORACLE_MAXD], lj = 0;
  FN(time_embed)(P, step, &tc, &ts);
  FN(grad_energy)(P, x, g);
  FN(net_apply)(P, P->hmc ? NULL : P->vnet, x, g, tc, ts, S, T, Q);
  for (int d = 0; d < D; ++d) {
    const REAL sv2 = (REAL)-0.5 * eps * S[d], fv2 = eps * Q[d];
    v[d] = (v[d] - (REAL)0.5 * eps * (-(EXP(fv2) * g[d]) + T[d])) * EXP(sv2);
    lj += sv2;
  }
  for (int d = 0; d < D; ++d) in[d] = ((REAL)1 - (REAL)m[d]) * x[d];
  FN(net_apply)(P, P->hmc ? NULL : P->xnet, v, in, tc, ts, S, T, Q);
  for (int d = 0; d < D; ++d) {
    const REAL mm = m[d], mb = (REAL)1 - mm, sx2 = -eps * S[d], fx2 = eps * Q[d];
    x[d] = mb * x[d] + mm * (EXP(sx2) * (x[d] - eps * (EXP(fx2) * v[d] + T[d])));
    lj += mm * sx2;
  }
  for (int d = 0; d < D; ++d) in[d] = (REAL)m[d] * x[d];
  FN(net_apply)(P, P->hmc ? NULL : P->xnet, v, in, tc, ts, S, T, Q);
  for (int d = 0; d < D; ++d) {
    const REAL mm = m[d], mb = (REAL)1 - mm, sx1 = -eps * S[d], fx1 = eps * Q[d];
    x[d] = mm * x[d] + mb * (EXP(sx1) * (x[d] - eps * (EXP(fx1) * v[d] + T[d])));
    lj += mb * sx1;
  }
  FN(grad_energy)(P, x, g);
  FN(net_apply)(P, P->hmc ? NULL : P->vnet, x, g, tc, ts, S, T, Q);
  for (int d = 0; d < D; ++d) {
    const REAL sv1 = (REAL)-0.5 * eps * S[d], fv1 = eps * Q[d];
    v[d] = EXP(sv1) * (v[d] - (REAL)0.5 * eps * (-(EXP(fv1) * g[d]) + T[d]));
    lj += sv1;
  }
  return lj;
}

/* forward / backward / p_accept: utils/dynamics.py:246-309.  dirflag 1 = forward, 0 = backward.
 * Writes X, V [n,D], logj [n], p [n] (any output may be NULL). */
void FN(oracle_trajectory)(const oracle_problem *P, int n, int dirflag, const REAL *x0, const REAL *v0, REAL *X,
                           REAL *V, REAL *logj, REAL *p) {
  const int D = P->D;
#pragma omp parallel for schedule(static)
  for (int c = 0; c < n; ++c) {
    REAL x[ORACLE_MAXD], v[ORACLE_MAXD], lj = 0;
    for (int d = 0; d < D; ++d) { x[d] = x0[c * D + d]; v[d] = v0[c * D + d]; }
    for (int t = 0; t < P->T; ++t)
      lj += dirflag ? FN(forward_step)(P, x, v, t) : FN(backward_step)(P, x, v, P->T - t - 1);
    if (X) for (int d = 0; d < D; ++d) X[c * D + d] = x[d];
    if (V) for (int d = 0; d < D; ++d) V[c * D + d] = v[d];
    if (logj) logj[c] = lj;
    if (p) { /* p_accept :302-309 */
      const REAL e_new = FN(hamiltonian)(P, x, v), e_old = FN(hamiltonian)(P, x0 + c * D, v0 + c * D);
      const REAL a = e_old - e_new + lj;
      REAL pr = EXP(a < 0 ? a : (a == a ? (REAL)0 : a));
      p[c] = isfinite(pr) ? pr : 0;
    }
  }
}

/* propose + tf_accept: utils/sampler.py:28-55 -- both directions for every chain, blended. */
void FN(oracle_propose)(const oracle_problem *P, int n, const REAL *x, const REAL *v_f, const REAL *v_b,
                        const unsigned char *dir, const REAL *u, int log_jac, REAL *Lx, REAL *Lv, REAL *px,
                        REAL *x_next, REAL *scratch /* 2*(2*n*D + n) REALs */) {
  const int D = P->D;
  REAL *X1 = scratch, *V1 = X1 + n * D, *P1 = V1 + n * D, *X2 = P1 + n, *V2 = X2 + n * D, *P2 = V2 + n * D;
  FN(oracle_trajectory)(P, n, 1, x, v_f, X1, V1, log_jac ? P1 : NULL, log_jac ? NULL : P1);
  if (!P->hmc) FN(oracle_trajectory)(P, n, 0, x, v_b, X2, V2, log_jac ? P2 : NULL, log_jac ? NULL : P2);
  for (int c = 0; c < n; ++c) {
    const REAL mk = P->hmc ? (REAL)1 : (REAL)(dir[c] != 0);
    for (int d = 0; d < D; ++d) {
      Lx[c * D + d] = P->hmc ? X1[c * D + d] : mk * X1[c * D + d] + ((REAL)1 - mk) * X2[c * D + d];
      Lv[c * D + d] = P->hmc ? V1[c * D + d] : mk * V1[c * D + d] + ((REAL)1 - mk) * V2[c * D + d];
    }
    px[c] = P->hmc ? P1[c] : mk * P1[c] + ((REAL)1 - mk) * P2[c];
    if (x_next) {
      const int acc = (px[c] - u[c] >= 0);
      for (int d = 0; d < D; ++d) x_next[c * D + d] = acc ? Lx[c * D + d] : x[c * D + d];
    }
  }
}

#undef FN
#undef CAT
#undef CAT_
