/* Plain-C restatement of the L2HMC augmented-leapfrog sampling path.  TEST INFRASTRUCTURE ONLY.
 *
 * Second, independently written CPU statement of the reference algorithm (the first is
 * oracle/l2hmc_oracle.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product (l2hmc_b200/) never does.
 *
 * PARITY PINNED (round 2): tests/test_reference_pin.py checks this file (and the torch oracle) against vectors the
 * UNMODIFIED reference sources produced when run on the eager TensorFlow stand-in oracle/tf_shim
 * (tests/golden/make_ref_golden.py -> tests/golden/ref_*.npz).  The reference itself holds no golden vectors, tests or
 * seeds.
 *
 * Follows: utils/dynamics.py:95-108,115-218,246-309; utils/sampler.py:28-55; utils/layers.py:29-37,81-95;
 * net wiring SCGExperiment.ipynb:51-77; energies utils/distributions.py:31-32,50-57,90-97,125-134,161-180.
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC -> oracle/_build/libl2hmc_oracle.so)
 */
#include <math.h>
#include <stddef.h>

#define ORACLE_MAXD 128
#define ORACLE_MAXH 512
#define ORACLE_MAXCOMP 16

enum { W1, B1, W2, B2, W3, B3, W4, B4, WS, BS, WT, BT, WQ, BQ, LS, LQ, NET_NPARAM };

typedef struct {
  int D, H, T, hmc;
  float eps;                        /* exp(log(eps)) in fp32 (utils/dynamics.py:50-58) */
  const float *mask;                /* [T][D] (utils/dynamics.py:84-97) */
  const float *xnet[NET_NPARAM];    /* W [in][out] row-major, b, log-scales */
  const float *vnet[NET_NPARAM];
  int energy_kind, ncomp;           /* 0 Gaussian, 1 GMM, 2 RoughWell, 3 Funnel */
  const float *mu, *S, *logc;       /* [K][D], [K][D][D], [K] */
  float s0, s1;                     /* RoughWell: eps, denominator ; Funnel: sigma, clip */
  float temperature;
} oracle_problem;

#define REAL float
#define SUFFIX _f32
#define EXP expf
#define LOG logf
#define TANH tanhf
#define SIN sinf
#define COS cosf
#include "l2hmc_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef EXP
#undef LOG
#undef TANH
#undef SIN
#undef COS

#define REAL double
#define SUFFIX _f64
#define EXP exp
#define LOG log
#define TANH tanh
#define SIN sin
#define COS cos
#include "l2hmc_oracle_impl.h"
