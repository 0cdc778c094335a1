/*
 * d4s.h -- C ABI of the B200-native tall-data diffusion posterior sampler (libd4s.so).
 *
 * This is the drop-in boundary for the hot path of JuliaLinhart/diffusions-for-sbi:
 *     NSE.ddim -> NSE.ddim_step -> NSE.factorized_score -> tweedies_approximation
 *     (+ vp_diffused_priors prior scores, PF_NSE partial factorisation)
 * Every entry point names the reference interface it replaces (file:line relative to the reference
 * checkout).  The reference has no FFI of its own (it is pure Python/torch); the binding a maintainer
 * adds is a ctypes stub, shown in INTEGRATION.md and shipped as diffusions-for-sbi_b200/d4s/_cabi.py.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  All tensors are fp32, row-major, contiguous.
 *   - every pointer named *_dev / inside D4sStep / D4sProblem is a DEVICE pointer owned by the caller;
 *     the library allocates nothing except the packed net (d4s_net_create/d4s_net_free) and its
 *     CUDA-graph cache (d4s_graph_cache_clear).
 *   - all work is enqueued on the caller's stream (a cudaStream_t passed as void*); no host
 *     synchronisation, no allocation inside a launch => capturable in a CUDA graph.
 *   - return value 0 = ok, negative = error; the message is available from d4s_last_error().
 *     Nothing throws across the ABI.  There is NO CPU fallback: without a CUDA device every compute
 *     entry point returns D4S_ERR_CUDA.
 */
#ifndef D4S_H_
#define D4S_H_

#include <stdint.h>

#if defined(__GNUC__)
#define D4S_API __attribute__((visibility("default")))
#else
#define D4S_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define D4S_ABI_VERSION 15

#define D4S_OK 0
#define D4S_ERR_ARG -1      /* bad argument / unsupported shape */
#define D4S_ERR_CUDA -2     /* CUDA runtime error (message holds cudaGetErrorString) */
#define D4S_ERR_WORKSPACE -3 /* caller workspace too small */

#define D4S_MAX_LAYERS 8    /* linear layers of the score MLP (zuko MLP: 4) */
#define D4S_MAX_THETA 10    /* theta_dim m supported by the register-resident solvers */
#define D4S_MAX_THETA_WIDE 32 /* theta_dim of the generic path (torch scores + native kernels 3/4: d4s_reduce_pairs, d4s_finalize_wide) */
#define D4S_MAX_WIDTH 256   /* widest hidden layer */
#define D4S_SCHED_STRIDE 16 /* floats per step in the schedule table */
#define D4S_MAX_RANKS 8     /* ranks (GPUs of one NVSwitch box) of an observation-sharded run */
#define D4S_IPC_HANDLE_BYTES 64 /* opaque handle of an exchange buffer (sizeof(cudaIpcMemHandle_t)) */

/* covariance mode of tweedies_approximation (tall_posterior_sampler.py:46-135) */
#define D4S_COV_GAUSS 0
#define D4S_COV_JAC 1
/* prior kind (vp_diffused_priors.py:5-74); NONE = score without prior term (single observation) */
#define D4S_PRIOR_NONE 0
#define D4S_PRIOR_GAUSSIAN 1
#define D4S_PRIOR_UNIFORM 2
#define D4S_PRIOR_EXTERNAL 3 /* generic path only: prior score / precision tensors evaluated by the caller's closure */
/* how the per-sample total score is assembled (nse.py:628-654) */
#define D4S_COMBINE_SUM 0        /* (1-n) s0 + sum_j s_ij                    (nse.py:642, gate closed / geffner) */
#define D4S_COMBINE_SHARED 1     /* Lambda^-1 w with one Lambda per step     (GAUSS + Gaussian prior, nse.py:628-638) */
#define D4S_COMBINE_PER_SAMPLE 2 /* per-sample Lambda_i, LU solve            (JAC, or uniform prior inside the gate) */

/* schedule table columns (fp32, one row of D4S_SCHED_STRIDE floats per DDIM step; computed by the
 * host in fp64 from NSE.alpha/NSE.sigma, nse.py:163-170, and the DDIM coefficients nse.py:283-298) */
#define D4S_S_T 0          /* t */
#define D4S_S_ALPHA 1      /* alpha(t) */
#define D4S_S_SIGMA 2      /* sigma(t) */
#define D4S_S_INV_SIGMA 3  /* 1/sigma(t): score = -eps/sigma (nse.py:145) */
#define D4S_S_A_OVER_S2 4  /* alpha/sigma^2 (tall_posterior_sampler.py:42) */
#define D4S_S_C_THETA 5    /* theta' = C_THETA*theta + C_SCORE*score + BRIDGE_STD*z   (nse.py:289-298) */
#define D4S_S_C_SCORE 6
#define D4S_S_BRIDGE_STD 7
#define D4S_S_SQRT_ALPHA 8
#define D4S_S_COMBINE 9    /* D4S_COMBINE_* for this step (the alpha-gate of nse.py:643 is decided on the host) */
#define D4S_S_ONE_MINUS_N 10 /* (1 - n_obs) as float */
/* generalised update  theta' = C_THETA theta + c_eff * score + BRIDGE_STD z  shared by DDIM and the Langevin family
 * (annealed Langevin nse.py:418-475, tamed ULA nse.py:301-338, deterministic Geffner nse.py:658-705): */
#define D4S_S_S1_WEIGHT 11  /* D4S_COMBINE_SUM: score = (1-n) s0 + S1_WEIGHT * sum_j s_ij  (nse.py:688-694: 1/sqrt(gamma)) */
#define D4S_S_TAME 12       /* > 0: tamed step c_eff = TAME / (1 + TAME * ||score||)  (nse.py:336), else c_eff = C_SCORE */
#define D4S_S_CLIP 13       /* 1: apply theta_clipping_range after this step (Langevin clips once per outer step) */

typedef struct D4sNet D4sNet; /* packed, device-resident score network */

/* Host-side description of the final score MLP (nse.py:87-90 `build_net`, zuko MLP = Linear/ReLU stack).
 * Layer 0's input columns are [theta features (theta_cols) | everything else]; only the theta block is
 * kept for the kernel (the observation / time / one-hot blocks are folded by the host into the
 * per-observation and per-step layer-1 partials `obs_feat` and `time_feat`, see D4sProblem). */
typedef struct {
    int32_t n_layers;                    /* number of Linear layers, >= 2 */
    int32_t dims[D4S_MAX_LAYERS + 1];    /* dims[0] = input width, dims[n_layers] = theta_dim */
    int32_t theta_cols;                  /* leading input columns fed by (embedded) theta */
    const float* weight[D4S_MAX_LAYERS]; /* HOST pointers, [dims[l+1], dims[l]] row-major (torch layout) */
    const float* bias[D4S_MAX_LAYERS];   /* HOST pointers, [dims[l+1]] */
    /* optional embedding MLP of theta (nse.py:31,129 `embedding_nn_theta`, a Linear/ReLU stack like the main net):
     * emb_layers == 0 means nn.Identity (then theta_cols == theta_dim); otherwise emb_dims[0] == theta_dim and
     * emb_dims[emb_layers] == theta_cols.  Widths <= 256. */
    int32_t emb_layers;
    int32_t emb_dims[D4S_MAX_LAYERS + 1];
    const float* emb_weight[D4S_MAX_LAYERS];
    const float* emb_bias[D4S_MAX_LAYERS];
    /* output head of the toy nets (embedding_nets.py:127-160, FakeFNet/EpsilonNet):
     * eps = out_scale * act(W_L h + b_L) [+ analytic affine term, see D4sStep.affine]; out_act 0 = identity, 1 = tanh.
     * out_scale == 0 is read as 1 (plain MLP). */
    int32_t out_act;
    float out_scale;
} D4sNetDesc;

/* Everything that is constant along one trajectory / one factorized_score call. */
typedef struct {
    int32_t n_samples;       /* N: rows of theta */
    int32_t n_obs;           /* n: observations (or K subsets for PF_NSE, nse.py:587) */
    int32_t theta_dim;       /* m */
    int32_t cov_mode;        /* D4S_COV_* */
    int32_t prior_kind;      /* D4S_PRIOR_* */
    int32_t paired;          /* k >= 1: sample i has its OWN observation, row i / k of obs_feat, and the plain score is used (no
                                factorisation); n_obs = ceil(N / k).  k = 1: nse.py:253 `x.shape[0] == shape[0]`; k = S: the
                                covariance pre-pass vmap(ddim) over the observations, S samples each
                                (sbibm_posterior_estimation.py:306-314) (ABI v15: k > 1) */
    int32_t obs_chunk;       /* observations per tile (0 = library default) */
    int32_t n_obs_total;     /* n used in (1-n) and in Lambda (== n_obs unless observations are sharded over ranks) */
    const float* obs_feat;   /* [n_obs, H1p] layer-1 partial of observation j: W1[:, x-block] e_x(x_j) (+ one-hot block) */
    const float* obs_prec;   /* GAUSS: [n_obs, m, m] inv(Sigma_j)  (tall_posterior_sampler.py:42); else NULL */
    const float* prior_low;  /* uniform prior: [m] (vp_diffused_priors.py:5) */
    const float* prior_high; /* uniform prior: [m] */
    float clip_lo, clip_hi;  /* theta_clipping_range (nse.py:265-266); lo > hi disables */
    uint64_t seed;           /* Philox key for the DDIM noise (used when D4sStep.noise == NULL) */
    int64_t sample_offset;   /* global index of local sample 0 (sample sharding keeps the noise G-invariant) */
    float* workspace;        /* partial sums, >= d4s_workspace_floats() floats */
    int64_t workspace_floats;
    const float* obs_raw;    /* [n_obs, x_dim] raw observations, only read when D4sStep.affine != NULL */
    int32_t x_dim;
    int32_t ctx_samples;     /* 0: every sample sees all n_obs observations.  S > 0 (batched contexts, replaces
                                vmap(sampler) over calibration sets, jrnnm_posterior_estimation.py:286-294): samples
                                [c S, (c+1) S) belong to context c and see observations [c n_obs, (c+1) n_obs) of
                                obs_feat / obs_prec / obs_raw (n_samples must be a multiple of S) */
    const float* obs_weight; /* NULL (all 1), or [n_obs] weights of the observations in every sum over j.  Classifier-free
                                guidance (nse.py:598-625, 532-538) is the tall score with one extra pseudo-observation
                                x = 0 (the net's own prior score) of weight (1 - n). */
    const float* ctx_qsum;   /* batched contexts, GAUSS: [n_samples / ctx_samples, m, m] sum_j inv(Sigma_cj) of every context,
                                added to the step's Lambda base (which then carries no observation term); NULL otherwise */
} D4sProblem;

/* Per-step inputs. */
typedef struct {
    const float* sched;     /* [D4S_SCHED_STRIDE] this step's row of the schedule table */
    const float* time_feat; /* [H1p] layer-1 partial of the time embedding + bias: W1[:, t-block] temb(t) + b1 */
    const float* mats;      /* per-step small matrices, see d4s_mats_floats(): Lmat[m*m], G[m*m], mu_t[m] */
    const float* noise;     /* [N, m] injected z for this step, or NULL => Philox(seed, step, sample, dim) */
    int32_t step_index;     /* Philox counter word; also selects nothing else */
    int32_t update;         /* 1: theta <- DDIM update (nse.py:289-298); 0: only write the score */
    const float* affine;    /* NULL, or [m*m + m*x_dim + m]: analytic part of eps, eps += M_t theta_i - G_t x_j - g_t
                               (the closed-form Gaussian posterior score of gen_gaussian_gaussian.py:59-72) */
} D4sStep;

/* Observation sharding over the GPUs of one box (the reference cannot run this shape: it
 * materialises [N, n, m, m], tall_posterior_sampler.py:114-128).  Rank r evaluates the score network on ITS slice of
 * the observations for all N samples.  The one exchange step of the path -- the sum over ranks of the per-sample
 * vectors S1_i = sum_j s_ij and S2_i = sum_j inv(Sigma_j) s_ij -- happens INSIDE the trajectory kernel: a CTA stores
 * its partial sums into every peer's receive buffer through mapped peer memory (NVLink / NVSwitch), raises a flag
 * there, waits for the peers' flags in its own buffer and adds the contributions in rank order, so that every rank
 * continues with bit-identical theta and no broadcast is needed.  One launch still runs every DDIM step.
 * Buffers come from d4s_xchg_alloc (cudaMalloc + IPC handle); peers map them with d4s_xchg_open.
 *   data [r]: float  [2][world][N][2 m]      (slot = (seq_base + step) & 1, second index = source rank)
 *   flags[r]: uint32 [world][n_groups]       zero-initialised once; n_groups from d4s_xchg_layout()
 * seq_base must be the same on every rank and grow by `steps` from one trajectory to the next on the same buffers. */
typedef struct {
    int32_t world;                 /* ranks taking part (1 .. D4S_MAX_RANKS) */
    int32_t rank;                  /* this process */
    uint32_t seq_base;             /* sequence number of step 0 */
    int32_t reserved;
    float* data[D4S_MAX_RANKS];    /* data[r]  = rank r's receive buffer as mapped in THIS process (data[rank] is local) */
    uint32_t* flags[D4S_MAX_RANKS]; /* flags[r] = rank r's flag array as mapped in this process */
} D4sExchange;

/* ---- library ---------------------------------------------------------------------------------- */
D4S_API int d4s_abi_version(void);
D4S_API const char* d4s_last_error(void);
/* number of visible CUDA devices (0 => every compute call fails with D4S_ERR_CUDA) */
D4S_API int d4s_device_count(void);

/* ---- packed net (replaces the nn.Module weights held by NSE.net, nse.py:87-90) ------------------ */
D4S_API int d4s_net_create(D4sNet** out, const D4sNetDesc* desc, void* stream);
D4S_API void d4s_net_free(D4sNet* net);
/* padded width of hidden layer 0 (row stride of obs_feat / time_feat) */
D4S_API int d4s_net_h1_padded(const D4sNet* net);

/* ---- sizes -------------------------------------------------------------------------------------- */
D4S_API int64_t d4s_workspace_floats(const D4sNet* net, const D4sProblem* prob);
/* floats of the leading partial-sum block [N, C, W] of the workspace (what observation sharding all-reduces) */
D4S_API int64_t d4s_partials_floats(const D4sNet* net, const D4sProblem* prob);
D4S_API int32_t d4s_mats_floats(int32_t theta_dim);

/* ---- kernel 1+2: score network on the (sample x observation) grid, Tweedie precisions, partial sums.
 * Replaces tweedies_approximation (tall_posterior_sampler.py:46-135) as called from
 * NSE.factorized_score (nse.py:588-596).  Writes per-(sample, chunk) partial sums to prob->workspace;
 * optionally materialises scores[N,n,m] (and JAC precisions prec[N,n,m,m]) for the B2 hook. */
D4S_API int d4s_score_partials(const D4sNet* net, const D4sProblem* prob, const D4sStep* step, const float* theta_dev,
                       float* scores_out_dev, float* prec_out_dev, void* stream);

/* ---- kernel 3+4: reduction over observation chunks, diffused prior score, Lambda solve, DDIM update.
 * Replaces the tail of NSE.factorized_score (nse.py:627-656) + NSE.ddim_step (nse.py:281-298) +
 * vp_diffused_priors (vp_diffused_priors.py:20-72) + tweedies_approximation_prior
 * (tall_posterior_sampler.py:138-155).  theta is updated in place when step->update; score_out may be NULL. */
D4S_API int d4s_finalize(const D4sNet* net, const D4sProblem* prob, const D4sStep* step, float* theta_dev,
                 float* score_out_dev, void* stream);

/* ---- whole trajectory: replaces the loop of NSE.ddim (nse.py:263-266).  sched_dev[steps, 16],
 * time_feat_dev[steps, H1p], mats_dev[steps, d4s_mats_floats(m)], noise_dev NULL or [steps, N, m].
 * affine_dev: NULL or [steps, m*m + m*x_dim + m] (see D4sStep.affine).
 * theta_dev holds the initial draw on entry and the samples on return (stream-ordered).
 * step_cov_mode_host: NULL, or a HOST array [steps] of D4S_COV_* overriding prob->cov_mode per step (a JAC
 * run needs no Jacobian on the steps whose combine mode is D4S_COMBINE_SUM: nse.py:642-643 discards it).
 * use_graph != 0 captures the steps into a CUDA graph (cached by argument identity). */
D4S_API int d4s_ddim_run(const D4sNet* net, const D4sProblem* prob, int32_t steps, const float* sched_dev,
                 const float* time_feat_dev, const float* mats_dev, const float* noise_dev, float* theta_dev,
                 const int32_t* step_cov_mode_host, const float* affine_dev, int32_t use_graph, void* stream);
D4S_API void d4s_graph_cache_clear(void);

/* ---- observation-sharded trajectory with the in-kernel exchange (see D4sExchange) ------------------------------
 * Same arguments as d4s_ddim_run; prob describes THIS rank's observation slice (n_obs local, n_obs_total global; the
 * schedule / matrix tables are built from the global n).  GAUSS-type steps on the fused tcgen05 trajectory kernel only
 * (returns D4S_ERR_ARG otherwise: callers then use d4s_score_partials + a collective + d4s_finalize per step). */
D4S_API int d4s_ddim_run_sharded(const D4sNet* net, const D4sProblem* prob, int32_t steps, const float* sched_dev,
                         const float* time_feat_dev, const float* mats_dev, const float* noise_dev, float* theta_dev,
                         const D4sExchange* xchg, void* stream);
/* floats of one rank's data buffer and uint32 words of its flag array for this problem (both ranks-symmetric) */
D4S_API int d4s_xchg_layout(const D4sNet* net, const D4sProblem* prob, int32_t world, int64_t* data_floats,
                    int64_t* flag_words);
/* Exchange buffers: plain device allocations of this process that peers can map (cudaIpcGetMemHandle /
 * cudaIpcOpenMemHandle).  alloc zero-fills; handle_out receives D4S_IPC_HANDLE_BYTES opaque bytes to send to the peers. */
D4S_API int d4s_xchg_alloc(int64_t bytes, void** dev_ptr_out, uint8_t* handle_out);
D4S_API int d4s_xchg_free(void* dev_ptr);
D4S_API int d4s_xchg_open(const uint8_t* handle, void** peer_ptr_out);
D4S_API int d4s_xchg_close(void* peer_ptr);
/* Library options: "path" = 0 auto | 1 per-step fp32 SIMT kernels only | 2 require the fused tcgen05 trajectory
 * kernel (bf16x3 split on the tensor cores; GAUSS, tall grid); "tc_steps_per_launch" = DDIM steps per tcgen05
 * launch (0 = whole trajectory in one launch); "profile" = 1 enables the in-kernel phase timers; "tc_pairing" = 1
 * (default) lets a CTA that owns two sample groups interleave them step by step, 0 runs them one after the other
 * (same arithmetic, bit-identical results); "jac_tc" = 1 (default) runs JAC steps (theta_dim <= 5) on the tcgen05 score +
 * Jacobian kernel (fp16x3 products, cross terms accumulated first, accumulator debias: same tolerance as the fp32 SIMT
 * kernel, ~5x faster), 0 = fp32 SIMT kernel; "paired_tc" = 1 (default) runs the paired mode (D4sProblem.paired) on its
 * tcgen05 trajectory kernel, 0 = fp32 SIMT kernels; "tc_cta_pair" / "jac_cta_pair" = 1 launch the trajectory kernel / the
 * JAC step kernel as clusters of two CTAs on one cta_group::2 instruction stream (default 0: bit-identical, measured
 * slower). */
D4S_API int d4s_set_option(const char* key, int32_t value);
/* "profile" = 1 makes CTA 0 of the tcgen05 kernel accumulate clock64() cycles per phase; this call synchronises and
 * copies the 16 counters of the last launch: 0 layer-0 sample part, 1 layer 0, 2 prior+noise, 3 tensor-core wait,
 * 4 epilogues, 5 scores, 6 inv(Sigma) s + sums, 7 finalize; 8 / 9: MMA issue -> commit of the inner / last hidden layer
 * (measured by the MMA warp). */
D4S_API int d4s_get_profile(int64_t* out16);

/* ---- stand-alone pieces of kernel 3/4 (API parity with the reference's public helpers) ---------------- */
/* prior_score_fn(theta, t) of vp_diffused_priors.py:20-46 (uniform: also d score / d theta, the diagonal
 * that tweedies_approximation_prior obtains by jacrev, tall_posterior_sampler.py:147) and :55-72
 * (gaussian: mats_dev row must hold G = -C^-T and mu_t, layout of D4sStep.mats). */
D4S_API int d4s_prior_score(int32_t prior_kind, const float* theta_dev, int64_t n_rows, int32_t m, const float* sched_dev,
                    const float* mats_dev, const float* low_dev, const float* high_dev, float* score_out_dev,
                    float* jacdiag_out_dev, void* stream);
/* NSE.ddim_step with a caller-supplied score (nse.py:281-298): theta <- C_THETA theta + C_SCORE score +
 * BRIDGE_STD z, then clip.  noise_dev NULL => Philox(seed, step_index, sample_offset + i, k). */
D4S_API int d4s_ddim_update(float* theta_dev, const float* score_dev, int64_t n_rows, int32_t m, const float* sched_dev,
                    const float* noise_dev, uint64_t seed, int32_t step_index, int64_t sample_offset, float clip_lo,
                    float clip_hi, void* stream);

/* ---- generic path: score networks that cannot be fused (a Python closure inside FakeFNet, gen_gaussian_gaussian.py:59-72;
 * modules other than Linear/ReLU stacks; theta_dim up to D4S_MAX_THETA_WIDE, gen_gaussian_gaussian.py:24).  torch evaluates
 * the scores (and, in JAC mode, the per-pair precisions) on the (sample x observation) grid; kernels 3 and 4 stay native.
 *
 * d4s_reduce_pairs = kernel 3a, the reduction over observations (nse.py:631-636, nse_toy.py:619-622 do it with torch on
 * materialised [N, n, m, m] tensors): scores_dev[N, n, m];
 *   pair_prec_dev != NULL ([N, n, m, m], JAC): part_out[N, m + m*m] = sum_j w_j P_ij s_ij | sum_j w_j P_ij
 *   else (GAUSS):                              part_out[N, 2m]      = sum_j w_j s_ij | sum_j w_j Q_j s_ij, Q = obs_prec_dev[n, m, m] or NULL
 * obs_weight_dev NULL (all 1) or [n]. */
D4S_API int d4s_reduce_pairs(const float* scores_dev, const float* pair_prec_dev, const float* obs_prec_dev,
                     const float* obs_weight_dev, int64_t n_samples, int32_t n_obs, int32_t theta_dim, float* part_out_dev,
                     void* stream);
/* d4s_finalize_wide = kernel 3b + 4 for theta_dim <= D4S_MAX_THETA_WIDE on the buffer written by d4s_reduce_pairs (one warp
 * per sample, LU with partial pivoting in fp64).  Same step semantics as d4s_finalize (sched / mats rows, prior kinds, combine
 * modes, update rule, Philox keying); in addition prior_kind == D4S_PRIOR_EXTERNAL takes the prior score ext_s0_dev[N, m] and
 * (weighted steps) the prior precision ext_p0_dev[N, m, m] that the caller's prior closure returned (nse_toy.py:618):
 * g = (1-n) P0 s0, Lambda += (1-n) P0.  mats_dev may be NULL when no step matrix applies. */
D4S_API int d4s_finalize_wide(const float* part_dev, int64_t n_samples, int32_t theta_dim, int32_t layout_jac,
                      const float* sched_dev, const float* mats_dev, int32_t prior_kind, const float* low_dev,
                      const float* high_dev, const float* ext_s0_dev, const float* ext_p0_dev, const float* noise_dev,
                      uint64_t seed, int32_t step_index, int64_t sample_offset, float clip_lo, float clip_hi, int32_t update,
                      float* theta_dev, float* score_out_dev, void* stream);

/* ---- utilities ---------------------------------------------------------------------------------- */
/* out[b] = inv(a[b]) for b < batch, m x m, LU with partial pivoting (torch.linalg.inv at
 * tall_posterior_sampler.py:42,94,154).  a and out may alias. */
D4S_API int d4s_batched_inverse(const float* a_dev, float* out_dev, int32_t batch, int32_t m, void* stream);
/* z[i, k] ~ N(0,1) from Philox4x32-10(seed; step_index, sample_offset + i, k) -- the generator the
 * fused update uses, exposed so that a trajectory can be replayed against the reference with the
 * same noise (torch.randn_like at nse.py:298; DiagNormal.sample at nse.py:261 uses step_index = -1). */
D4S_API int d4s_philox_normal(float* out_dev, int64_t n_rows, int32_t m, uint64_t seed, int32_t step_index,
                      int64_t sample_offset, void* stream);
/* kernels launched by this library since load (the bench reports it as gpu_launches) */
D4S_API int64_t d4s_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* D4S_H_ */
