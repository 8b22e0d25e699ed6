/*
 * l2hmc.h -- C ABI of libl2hmc.so, the B200 (sm_100a) implementation of the L2HMC
 * augmented-leapfrog sampling path.
 *
 * The reference (brain-research/l2hmc, TF1/Python) has no FFI of its own; its boundary for
 * this path is the Python call surface Dynamics / propose / tf_accept / chain_operator and the
 * distribution energy callbacks.  Every entry point below names the
 * reference interface it sits under (file:line into /root/reference).  The Python package
 * l2hmc_b200 binds these with ctypes and keeps the reference's names and signatures.
 *
 * Conventions
 *  - All tensors are IEEE fp32, row-major [n, D] (chain-major) or [n]; direction / accept flags
 *    are uint8 [n].  (Reference: TF_FLOAT = tf.float32, utils/dynamics.py:27.)
 *  - Pointers in l2hmc_transition_args and the component calls are DEVICE pointers and are
 *    BORROWED for the duration of the (stream-ordered) call; the caller (PyTorch) owns storage.
 *    Pointers passed to the l2hmc_set_* calls are HOST pointers and are copied before return.
 *    l2hmc_transition_host takes HOST pointers for everything and does the copies itself.
 *  - Every call returns an l2hmc_status; the message for the last failure on a context is
 *    l2hmc_last_error(ctx) (l2hmc_last_error(NULL) for failures of l2hmc_create).
 *    No C++ exception crosses this boundary; CUDA errors are captured and translated.
 *  - A context is bound to one device and is not thread-safe.  Calls are asynchronous on the
 *    caller's stream (cudaStream_t passed as void*; NULL = legacy default stream).
 *  - Randomness: when v / dir / u pointers are non-NULL they are used verbatim; when NULL the
 *    kernel draws them from Philox4x32-10 keyed by (seed, counter, global chain id), so results
 *    do not depend on how chains are sharded over GPUs.  l2hmc_b200/philox.py is the host twin.
 */
#ifndef L2HMC_H_
#define L2HMC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct l2hmc_ctx l2hmc_ctx;

typedef enum {
  L2HMC_OK = 0,
  L2HMC_EINVAL = 1,        /* bad argument / state (message says which)            */
  L2HMC_ECUDA = 2,         /* a CUDA runtime call failed                           */
  L2HMC_EUNSUPPORTED = 3,  /* shape / energy kind outside what the kernels cover   */
  L2HMC_ENOMEM = 4
} l2hmc_status;

/* utils/distributions.py energy closures that the kernels evaluate analytically. */
typedef enum {
  L2HMC_ENERGY_NONE = -1,
  L2HMC_ENERGY_GAUSSIAN = 0,  /* Gaussian.get_energy_function       utils/distributions.py:50-57   */
  L2HMC_ENERGY_GMM = 1,       /* GMM.get_energy_function            utils/distributions.py:125-134 */
  L2HMC_ENERGY_ROUGHWELL = 2, /* RoughWell.get_energy_function      utils/distributions.py:90-97   */
  L2HMC_ENERGY_FUNNEL = 3,    /* GaussianFunnel.get_energy_function utils/distributions.py:161-180 */
  L2HMC_ENERGY_DECODER = 4,   /* energy(z, aux) of the VAE posterior target   mnist_vae.py:104-126 (l2hmc_set_energy_decoder) */
  L2HMC_ENERGY_MIXED = 5      /* (1 - beta) U_a + beta U_b of two closed-form kinds: curr_energy of utils/ais.py:44-45 (l2hmc_set_energy_mixed) */
} l2hmc_energy_kind;

typedef enum { L2HMC_XNET = 0, L2HMC_VNET = 1 } l2hmc_net_id; /* utils/dynamics.py:78-79 */

typedef enum {
  L2HMC_DIR_FORWARD = 0,   /* Dynamics.forward   utils/dynamics.py:246-272                       */
  L2HMC_DIR_BACKWARD = 1,  /* Dynamics.backward  utils/dynamics.py:274-300                       */
  L2HMC_DIR_PER_CHAIN = 2, /* propose: dir[n] given (1 = forward)  utils/sampler.py:34-44        */
  L2HMC_DIR_RANDOM = 3     /* propose: direction bit drawn in-kernel (Philox)                    */
} l2hmc_dir_mode;

typedef enum {
  L2HMC_KERNEL_AUTO = 0,
  L2HMC_KERNEL_TILE = 1,   /* generic fp32-FMA tile kernel (any D <= 64, H <= 128)               */
  L2HMC_KERNEL_SMALL = 2,  /* one chain per thread, nets in registers (D <= 4, H <= 16)          */
  L2HMC_KERNEL_TC = 3,     /* tcgen05 3xTF32 tensor-core kernel                                  */
  L2HMC_KERNEL_LAYERED = 4, /* batched GEMM + elementwise launches over all chains: any x_dim / width,
                               the decoder energy and aux-conditioned nets (state in HBM between launches);
                               GEMMs on tcgen05 (3xTF32, fp32-level accuracy)                              */
  L2HMC_KERNEL_LAYERED_FMA = 5 /* the same engine with fp32-FMA GEMMs                                     */
} l2hmc_kernel_kind;

/* Dynamics.__init__(x_dim, energy_function, T, eps, hmc, net_factory, ...)  utils/dynamics.py:35-81 */
typedef struct {
  int32_t x_dim;   /* D                                                                          */
  int32_t width;   /* H, hidden width of the S/T/Q nets (ignored when hmc != 0)                  */
  int32_t T;       /* leapfrog steps per trajectory (reference default 25)                       */
  int32_t hmc;     /* 1: both nets return zeros (utils/dynamics.py:73-76)                        */
  int32_t device;  /* CUDA device ordinal                                                        */
  int32_t kernel;  /* l2hmc_kernel_kind                                                          */
  float eps;       /* step size; the reference holds alpha = log(eps), eps = exp(alpha) (:50-58) */
} l2hmc_config;

/* One S/T/Q net as built by net_factory (SCGExperiment.ipynb:51-77): variables
 * {embed_1,embed_2,embed_3,linear_1,linear_s,linear_t,linear_f}/{W,b}, {scale_s,scale_f}/scale.
 * W matrices are [in, out] row-major like utils/layers.py:33. HOST pointers. */
typedef struct {
  const float *W1, *b1; /* embed_1  [D,H], [H]                                                   */
  const float *W2, *b2; /* embed_2  [D,H], [H]                                                   */
  const float *W3, *b3; /* embed_3  [2,H], [H]                                                   */
  const float *W4, *b4; /* linear_1 [H,H], [H]                                                   */
  const float *Ws, *bs; /* linear_s [H,D], [D]                                                   */
  const float *Wt, *bt; /* linear_t [H,D], [D]                                                   */
  const float *Wq, *bq; /* linear_f [H,D], [D]                                                   */
  const float *scale_s; /* scale_s/scale [D] (log-scale; the layer multiplies by exp(scale))     */
  const float *scale_q; /* scale_f/scale [D]                                                     */
} l2hmc_net_params;

/* One transition = propose (+ optional tf_accept):  utils/sampler.py:28-55.
 * For n_transitions > 1 the kernel iterates x <- x_next on-chip (the reference's host loop of
 * sess.run, SCGExperiment.ipynb:291-298) and requires do_mh != 0 and Philox randomness. */
typedef struct {
  int64_t n;              /* chains in this call                                                 */
  int64_t chain_offset;   /* global id of chain 0 (Philox keying when chains are sharded)        */
  const float *x;         /* [n,D]                                                               */
  const float *v;         /* [n,D] momentum to use, or NULL: draw N(0,I)                         */
  const uint8_t *dir;     /* [n], dir_mode == L2HMC_DIR_PER_CHAIN                                */
  const float *u;         /* [n] accept uniforms, or NULL: draw U[0,1) (only read if do_mh)      */
  int32_t dir_mode;       /* l2hmc_dir_mode                                                      */
  int32_t log_jac;        /* 1: px_out = accumulated log|J| ; 0: px_out = p_accept               */
  int32_t do_mh;          /* 1: also write x_next = tf_accept(x, Lx, px)                         */
  int32_t n_transitions;  /* >= 1                                                                */
  uint64_t seed;          /* Philox key                                                          */
  uint64_t counter;       /* Philox call counter (advance by n_transitions per call)             */
  float *x_out;           /* Lx [n,D]                                                            */
  float *v_out;           /* Lv [n,D] or NULL                                                    */
  float *px_out;          /* [n]                                                                 */
  float *x_next;          /* [n,D] or NULL (required when do_mh)                                 */
  uint8_t *accepted;      /* [n] or NULL                                                         */
  void *stream;           /* cudaStream_t                                                        */
  const float *aux;       /* [n,aux_dim] conditioning rows (propose(..., aux=) utils/sampler.py:28; the
                             image batch `inp` of mnist_vae.py:196,204), or NULL when the target takes none */
  double *stats;          /* [2] or NULL, ACCUMULATED (+=): stats[0] += sum of px_out values, stats[1] += number of
                             accepted proposals, over every chain and every fused transition of the call -- the
                             acceptance statistics the notebook prints from np.mean(px_) (SCGExperiment.ipynb:268),
                             reduced inside the kernel (warp shuffle + one atomic pair per warp).  DEVICE pointer for
                             l2hmc_transition, HOST pointer for l2hmc_transition_host.                            */
  float *trace;           /* [n_transitions,n,D] or NULL: the Metropolis output after EVERY fused transition (the
                             notebook's final_samples list, SCGExperiment.ipynb:291-298); needs do_mh; device only */
  int32_t chain;          /* 1: chain_operator (utils/sampler.py:57-85) in ONE launch.  n_transitions is then nb_steps: that
                             many SUB-PROPOSALS are composed -- each with a fresh direction bit and fresh momentum
                             (dir [n_transitions,n], v [n_transitions,n,D] when given, else Philox at counter + s), log|J|
                             accumulated, no Metropolis step in between -- and the call closes with
                             px_out = p_accept(x, v0, x_K, v_K, sum log|J|)  (or the summed log|J| when log_jac), ONE set
                             of uniforms u [n] and x_next = tf_accept(x, x_K, px).  The reference's quirk is kept: the
                             sub-proposals ignore the carried momentum, the final Hamiltonians pair x with v0 = init_v
                             and x_K with the LAST sub-proposal's momentum.  Fused kernels only (small, tile, the
                             shape-specialised tensor-core kernel); others: L2HMC_EUNSUPPORTED.                       */
  const float *v0;        /* chain mode: init_v [n,D] (utils/sampler.py:58-59), or NULL: drawn (Philox at counter + n_transitions) */
} l2hmc_transition_args;

/* ---- lifetime ------------------------------------------------------------------------------ */
int l2hmc_create(const l2hmc_config *cfg, l2hmc_ctx **out); /* Dynamics.__init__ utils/dynamics.py:35 */
void l2hmc_destroy(l2hmc_ctx *ctx);
const char *l2hmc_last_error(const l2hmc_ctx *ctx);
const char *l2hmc_version(void);

/* ---- parameters (host pointers, copied) ---------------------------------------------------- */
int l2hmc_set_net(l2hmc_ctx *ctx, int net_id, const l2hmc_net_params *p); /* XNet/VNet utils/dynamics.py:78-79 */
int l2hmc_set_masks(l2hmc_ctx *ctx, const float *mask /* [T,D] of {0,1} */); /* Dynamics.mask utils/dynamics.py:84-97 */
int l2hmc_set_eps(l2hmc_ctx *ctx, float eps);                 /* Dynamics.eps utils/dynamics.py:58 */
int l2hmc_set_temperature(l2hmc_ctx *ctx, float temperature); /* Dynamics.temperature utils/dynamics.py:47,203-207 */
/* Decoder energy only: U = beta * sum_pixels BCE + 0.5 |z|^2 -- the annealed energy (1-beta) prior + beta posterior of
 * utils/ais.py:44-45 / eval_vae.py:52-62 (beta = 1: the sampler's energy, mnist_vae.py:122-126). */
int l2hmc_set_likelihood_scale(l2hmc_ctx *ctx, float beta);
/* Energy closure parameters (utils/distributions.py):
 *  GAUSSIAN : mu [D], S [D,D] (= i_sigma as fp32), n_comp = 1
 *  GMM      : mu [K,D], S [K,D,D], logc [K] (= log of the fp32 constants, :120-123), n_comp = K <= 8
 *  ROUGHWELL: scalars[0] = eps, scalars[1] = the cosine's denominator: eps if easy else eps*eps
 *             (evaluated in double, rounded once to fp32, as python-float * tensor does in TF)
 *  FUNNEL   : scalars[0] = sigma (2.0), scalars[1] = clip (8.0) */
int l2hmc_set_energy(l2hmc_ctx *ctx, int kind, int n_comp, const float *mu, const float *S,
                     const float *logc, const float *scalars, int n_scalars);

/* The annealed energy of annealed importance sampling, utils/ais.py:44-45:
 *   curr_energy(z) = (1 - beta) * init_energy(z) + beta * final_energy(z)
 * for two closed-form energies (kinds GAUSSIAN / GMM / ROUGHWELL / FUNNEL, each described like the arguments of
 * l2hmc_set_energy).  Evaluated per chain inside the fused kernels (small / tile; HMC-mode Dynamics as utils/ais.py:58
 * builds them); l2hmc_set_mix_beta moves beta between annealing steps without re-sending the parameters. */
typedef struct {
  int32_t kind, n_comp;
  const float *mu, *S, *logc, *scalars;
  int32_t n_scalars;
} l2hmc_energy_desc;
int l2hmc_set_energy_mixed(l2hmc_ctx *ctx, const l2hmc_energy_desc *a, const l2hmc_energy_desc *b, float beta);
int l2hmc_set_mix_beta(l2hmc_ctx *ctx, float beta);

/* energy(z, aux) = sum_pix sigmoid_cross_entropy_with_logits(labels=aux, logits=decoder(z)) + 0.5 |z|^2
 * (mnist_vae.py:122-126) with decoder = Linear, softplus, ..., Linear (mnist_vae.py:104-111).
 * widths [n_layers+1] = {x_dim, ..., aux_dim}; W[i] is [widths[i], widths[i+1]] row-major, b[i] [widths[i+1]].
 * Runs on the layered engine; every transition / component call then needs aux rows. */
int l2hmc_set_energy_decoder(l2hmc_ctx *ctx, int n_layers, const int32_t *widths, const float *const *W,
                             const float *const *b);
/* The aux branch of the S/T/Q nets' first stage: h1 = relu(embed_1(a) + embed_2(b) + embed_3(t) + enc(aux)),
 * enc = Linear, softplus, ..., Linear shared by XNet and VNet (encoder_sampler, mnist_vae.py:134-149).
 * widths [n_layers+1] = {aux_dim, ..., width}.  n_layers = 0 removes it (`lambda _: 0.`, SCGExperiment.ipynb:58). */
int l2hmc_set_aux_encoder(l2hmc_ctx *ctx, int n_layers, const int32_t *widths, const float *const *W,
                          const float *const *b);
/* aux rows (DEVICE pointer [n,aux_dim], borrowed until rebound) used by the component calls below
 * (Dynamics.energy(x, aux=aux) etc., utils/dynamics.py:203-218,302); NULL unbinds. */
int l2hmc_bind_aux(l2hmc_ctx *ctx, int64_t n, const float *aux);

/* ---- the hot path -------------------------------------------------------------------------- */
int l2hmc_transition(l2hmc_ctx *ctx, const l2hmc_transition_args *a); /* propose utils/sampler.py:28-51 */
/* Same call with HOST buffers (pageable or pinned): H2D of inputs, kernel, D2H of outputs, stream
 * synchronised before return.  This is what a reference-side sess.run replacement measures. */
int l2hmc_transition_host(l2hmc_ctx *ctx, const l2hmc_transition_args *a);

/* ---- components (Dynamics methods), device pointers ---------------------------------------- */
int l2hmc_energy(l2hmc_ctx *ctx, int64_t n, const float *x, float *out, void *stream);       /* Dynamics.energy      utils/dynamics.py:203-212 */
int l2hmc_grad_energy(l2hmc_ctx *ctx, int64_t n, const float *x, float *out, void *stream);  /* Dynamics.grad_energy utils/dynamics.py:217-218 */
int l2hmc_kinetic(l2hmc_ctx *ctx, int64_t n, const float *v, float *out, void *stream);      /* Dynamics.kinetic     utils/dynamics.py:107-108 */
int l2hmc_hamiltonian(l2hmc_ctx *ctx, int64_t n, const float *x, const float *v, float *out, void *stream); /* utils/dynamics.py:214-215 */
int l2hmc_p_accept(l2hmc_ctx *ctx, int64_t n, const float *x0, const float *v0, const float *x1,
                   const float *v1, const float *log_jac, float *out, void *stream);         /* Dynamics.p_accept    utils/dynamics.py:302-309 */
/* net([a, b, t, aux]) -> [S, T, Q]; t is the scalar step index the reference feeds _format_time (utils/dynamics.py:99-105). */
int l2hmc_net_apply(l2hmc_ctx *ctx, int net_id, int64_t n, const float *a, const float *b, float step,
                    float *S, float *T, float *Q, void *stream);
/* tf_accept(x, Lx, px): where(px - u >= 0, Lx, x)  utils/sampler.py:53-55. u NULL => Philox(seed, counter). */
int l2hmc_accept(l2hmc_ctx *ctx, int64_t n, int64_t chain_offset, const float *x, const float *Lx, const float *px,
                 const float *u, uint64_t seed, uint64_t counter, float *out, uint8_t *accepted, void *stream);
/* The in-kernel generator, exported so tests can compare it with the host twin and re-inject it:
 * v [n,D] normals, dir [n] bits, u [n] uniforms (any may be NULL). */
int l2hmc_philox_fill(l2hmc_ctx *ctx, int64_t n, int64_t chain_offset, uint64_t seed, uint64_t counter,
                      float *v, uint8_t *dir, float *u, void *stream);

/* ---- diagnostics on a device-resident sample trace -----------------------------------------
 * acl_spectrum(X, scale) = [autocovariance(X / scale, tau) for tau in range(n_lags)]  (utils/func_utils.py:45-54,
 * 114-116; the notebook calls it with n_lags = n_steps - 1, SCGExperiment.ipynb:331-334), where
 * autocovariance(X, tau) = mean_t( sum_{chain, dim} X[t] * X[t + tau] / n ).  trace: DEVICE fp32 [n_steps, n, x_dim]
 * (the samples of consecutive transitions, never copied to the host); out: DEVICE fp64 [n_lags]; products and sums
 * are fp64 here (the reference's numpy keeps float32 products and per-step sums; the two agree to fp32 rounding).  ESS (utils/func_utils.py:118-120) is a 2-line reduction of `out`. */
int l2hmc_acl_spectrum(l2hmc_ctx *ctx, int64_t n_steps, int64_t n, const float *trace, double scale, int64_t n_lags,
                       double *out, void *stream);

/* ---- introspection ------------------------------------------------------------------------- */
/* Sticky status bits of the context, raised by its kernels in pinned host-mapped memory and read here WITHOUT a device
 * synchronisation.  L2HMC_STATUS_F16_RANGE: a launch of the tensor-core kernel met an activation outside the fp16 range
 * while using the fp16 operand split; the chains of that launch that were affected carry non-finite proposals (accept
 * probability 0: rejected, x kept -- utils/dynamics.py:309), every later launch of the context uses the tf32 split, and
 * l2hmc_transition_host repeats its call itself.  clear != 0 resets the word (and allows the fp16 split again). */
#define L2HMC_STATUS_F16_RANGE 1u
int l2hmc_status_flags(l2hmc_ctx *ctx, uint32_t *flags, int clear);
/* ---- training path (first-correct version) -------------------------------------------------
 * Replaces, for one `propose` batch, what the reference obtains from TF1 autodiff: `tf.gradients(loss, params)` behind
 * `AdamOptimizer.minimize(loss)` (SCGExperiment.ipynb:183-188) with
 *   v = sum((x - Lx)^2) * px + 1e-4 ; loss = scale * mean(1 / v) - mean(v) / scale       (SCGExperiment.ipynb:171-181,
 *   utils/losses.py:36-59),
 * back-propagated through propose's selected direction (utils/sampler.py:34-44), p_accept (utils/dynamics.py:302-309),
 * the unrolled leapfrog (:246-300) and tf.gradients(energy, x) inside it (:217-218).
 * Gradient tensors have the shapes of l2hmc_net_params; every output is ACCUMULATED (+=) so that the caller zeroes once
 * and adds the `x` batch and the `z` batch of the notebook objective.  Device pointers throughout; asynchronous on the
 * stream (scratch, 4*T*2*x_dim floats per chain for the record plus activations, is held by the context).  Covers the closed-form energies (Gaussian, GMM, RoughWell, funnel) without aux; the decoder target and
 * aux-conditioned nets: L2HMC_EUNSUPPORTED.
 * Determinism: on the launch-sequence path (any shape) the sums over chains -- weight-gradient products, bias column sums,
 *   loss -- are split over CTAs into per-part slices and added in part order by a second kernel: two calls on the same
 *   inputs return the same bits.  The fused kernel for the notebook's small nets (x_dim <= 4, width <= 16, T <= 32)
 *   adds one partial sum per block with fp32 atomicAdd: there two calls agree to fp32 rounding of another summation
 *   order (~1e-6 relative), not bit for bit (L2HMC_TRAIN_FUSED=0 selects the launch sequence); loss, Lx and px_out are
 *   deterministic on both.  (TF1's reductions on the reference's GPU path are not run-to-run deterministic either.)
 * Streams: the scratch belongs to the context -- calls on ONE context must be issued on one stream (or ordered by the
 *   caller); use one context per stream for concurrent batches. */
typedef struct {
  float *W1, *b1, *W2, *b2, *W3, *b3, *W4, *b4, *Ws, *bs, *Wt, *bt, *Wq, *bq, *scale_s, *scale_q;
} l2hmc_net_grads;

typedef enum {                 /* get_loss(name), utils/losses.py:26-34, on v = loss_vec(x, Lx, px) (:36-37)        */
  L2HMC_LOSS_MIXED = 0,        /* loss_mixed     scale * mean(1 / v) - mean(v) / scale   utils/losses.py:53-59      */
  L2HMC_LOSS_STANDARD = 1,     /* loss_std       -mean(v)                                utils/losses.py:49-51      */
  L2HMC_LOSS_INVERSE = 2,      /* loss_inverse   -1 / mean(1 / (v + 1e-4))               utils/losses.py:44-47      */
  L2HMC_LOSS_LOGSUMEXP = 3     /* loss_logsumexp logsumexp(-v) - log N                   utils/losses.py:39-42      */
} l2hmc_loss_kind;             /* kinds 2 and 3 are not additive over calls: one call = one batch mean             */

typedef struct {
  int64_t n;              /* chains of this batch                                                */
  const float *x;         /* [n,D] start points                                                  */
  const float *v;         /* [n,D] the fresh momentum of each chain's direction (utils/sampler.py:35-36) */
  const uint8_t *dir;     /* [n] 1 = forward (the `mask` of utils/sampler.py:34)                 */
  float scale;            /* the notebook's `scale` (0.1)                                        */
  float inv_count;        /* 1 / (number of chains the means run over)                           */
  float *loss;            /* [1]  +=                                                             */
  float *d_eps;           /* [1]  += d loss / d eps (the reference trains alpha = log eps: times eps) */
  l2hmc_net_grads grad_xnet, grad_vnet; /* += */
  float *x_out;           /* Lx [n,D] or NULL                                                    */
  float *px_out;          /* [n] or NULL                                                         */
  void *stream;
  int32_t loss_kind;      /* l2hmc_loss_kind; 0 = the notebook's objective                        */
} l2hmc_loss_grad_args;

int l2hmc_loss_grad(l2hmc_ctx *ctx, const l2hmc_loss_grad_args *a);

const char *l2hmc_kernel_name(const l2hmc_ctx *ctx); /* kernel the next l2hmc_transition will launch */
int64_t l2hmc_launch_count(const l2hmc_ctx *ctx);    /* kernels launched by this context so far      */
/* CUDA-event timing of the hot kernel on the launching stream: average ms over launches since reset. */
int l2hmc_timing_enable(l2hmc_ctx *ctx, int on);
int l2hmc_timing_read(l2hmc_ctx *ctx, double *avg_ms, int64_t *launches); /* synchronises */
/* Phase accounting of the tensor-core kernel's CTA 0 in SM clock cycles (synchronises the device):
 * out[0] MMA issuer waiting for A operands, [1] waiting for TMA weight slabs, [2] issuer total,
 * [3] compute thread 0 waiting for accumulators, [4] compute thread 0 total, [5] GEMMs issued; [8..22] per GEMM kind;
 * [24 + 16 g + k] compute thread g in {0: first, 1: second thread of chain 0}: cycles working in epilogue kind k (0 embed,
 * 1 hidden, 2 / 3 heads part 0 / 1 of the V net, 4 / 5 of the X net, 6 grad), [24 + 16 g + 8 + k] waiting for its accumulator
 * (development builds with -DL2HMC_TC_PHASE_ACCOUNTING only; zeros otherwise). n <= 56. */
int l2hmc_debug_counters(l2hmc_ctx *ctx, int64_t *out, int n);

#ifdef __cplusplus
}
#endif
#endif /* L2HMC_H_ */
