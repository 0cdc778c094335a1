// Shared device/host helpers of libd4s (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/d4s.h"

namespace d4s {

// ------------------------------------------------------------------------------------------------
// tiling constants of the SIMT score kernel
// ------------------------------------------------------------------------------------------------
constexpr int TM = 64;      // rows (sample x observation [x tangent] pairs) per tile
constexpr int NT = 256;     // threads per CTA: 8 warps, warp w owns rows 8w..8w+7
constexpr int ASTR = 68;    // activation buffer row stride: act[k][ASTR] (k-major), conflict-free STS.128
constexpr int KC = 8;       // weight k-chunk staged per pipeline stage
constexpr int HMAX = D4S_MAX_WIDTH;
constexpr int MMAX = D4S_MAX_THETA;
constexpr int D4S_ACT_SHIFT16 = 4;  // fp16x3 operands hold 2^4 x activation (d4s_jac_tc.cu)

// Device view of the packed network.  Hidden widths are zero-padded to 64/128/256 so that padded
// units produce relu(0) = 0 and contribute nothing downstream.
struct NetDev {
    int L;                          // linear layers
    int m;                          // theta_dim
    int theta_cols;                 // leading input columns of layer 0 fed by theta
    int Hp[D4S_MAX_LAYERS];         // padded width of hidden layer l (l = 0 .. L-2)
    const float* A;                 // [theta_cols][Hp[0]]  layer-0 theta block, k-major
    const float* Wt[D4S_MAX_LAYERS];   // l = 1 .. L-2: [Hp[l-1]][Hp[l]] k-major (transposed torch weight)
    const float* bias[D4S_MAX_LAYERS]; // l = 1 .. L-2: [Hp[l]]
    const float* Wout;              // [Hp[L-2]][MMAX] k-major, zero padded columns
    const float* bout;              // [MMAX]
    const float* WoutT;             // [m][Hp[L-2]] output layer, torch layout (tcgen05 path epilogue)
    int out_act;                    // output head: 0 identity, 1 tanh (EpsilonNet, embedding_nets.py:135)
    float out_scale;                // eps = out_scale * act(.)   (FakeFNet.eps_net_max, embedding_nets.py:160)
    // optional theta-embedding MLP (nse.py:129): emb_L linear layers, k-major fp32, widths as given (<= 256)
    int emb_L;
    int emb_dims[D4S_MAX_LAYERS + 1];
    const float* emb_Wt[D4S_MAX_LAYERS];    // [in][out]
    const float* emb_bias[D4S_MAX_LAYERS];  // [out]
    const uint8_t* wimg[D4S_MAX_LAYERS];  // l = 1 .. L-2: bf16 hi/lo UMMA images, [Hp[l-1]/16][hi | lo][2][Hp[l]][8]
    const uint8_t* wimg2[D4S_MAX_LAYERS];  // the bf16 images again, every K-step split by output rows for a CTA pair: [Hp[l-1]/16][rank][hi | lo][2][Hp[l]/2][8]
    const uint8_t* wimg16p[D4S_MAX_LAYERS];  // the fp16 images split by output rows for a CTA pair (layout of wimg2)
    const uint8_t* wimg16[D4S_MAX_LAYERS];  // same layout, fp16 hi / lo of 2^sw16[l] W_l (JAC kernel: fp16x3 split)
    int sw16[D4S_MAX_LAYERS];               // power-of-two weight scale that keeps the fp16 lo parts out of the subnormals
    const float* bias16[D4S_MAX_LAYERS];    // 2^(D4S_ACT_SHIFT16 + sw16[l]) bias_l: the JAC kernel's accumulators are scaled
};


// ------------------------------------------------------------------------------------------------
// kernel parameter blocks (passed by value) and launchers implemented in the kernel translation units
// ------------------------------------------------------------------------------------------------
struct ScoreParams {
    NetDev net;
    int N, n, C, SG, OC, P;  // P = SG*OC pairs per tile
    int jac, paired;
    int W;                   // floats per (sample, chunk) partial: GAUSS 2m, JAC m + m*m
    const float* theta;      // [N][m]
    const float* base_pre;   // optional [N][Hp0]: A e_theta(theta_i) from the embedding kernel (else A theta_i in-kernel)
    const float* tan_pre;    // optional [N][m][Hp0]: A d e_theta / d theta_k (JAC tangent seeds)
    const float* obs_feat;   // [n][Hp0]
    const float* time_feat;  // [Hp0]
    const float* obs_prec;   // GAUSS: [n][m][m]
    const float* sched;      // [16]
    const float* affine;     // optional [m*m + m*d + m]: eps += M theta_i - G x_j - g   (analytic Gaussian posterior part)
    const float* obs_raw;    // [n][d] raw observations (read with `affine`)
    int d;
    const float* obs_weight; // optional [n] weights of the observations in the sums (classifier-free guidance)
    int ctx_samples;         // batched contexts: samples [c S, (c+1) S) see observations [c n, (c+1) n); 0 = one context
    float* part;             // [N][C][W]
    float* scores_out;       // optional [N][n][m]
    float* prec_out;         // optional JAC [N][n][m][m]
};

struct FinalizeParams {
    int N, C, W, m;
    int jac, prior_kind, update;
    const float* part;   // [N][C][W]
    const float* sched;  // [16]
    const float* mats;   // Lmat[m*m], G[m*m], mu_t[m]
    const float* low;    // uniform prior
    const float* high;
    const float* noise;  // [N][m] or NULL
    uint64_t seed;
    int step_index;
    long long sample_offset;
    float clip_lo, clip_hi;
    float* theta;      // [N][m] in/out
    float* score_out;  // optional [N][m]
    int ctx_samples;          // batched contexts (see ScoreParams)
    const float* ctx_qsum;    // [N / ctx_samples][m][m] sum_j inv(Sigma_cj), added to the Lambda base of the step; or NULL
};

// Observation sharding inside the trajectory kernel (D4sExchange, include/d4s.h): every rank evaluates its slice of the
// observations for ALL samples; after the per-sample sums of a step the CTA stores them straight into every peer's
// receive buffer (mapped peer memory, NVLink) and raises a flag there, waits for the peers' flags in its own buffer and
// adds the contributions in rank order (identical on every rank => theta stays bit-identical without a broadcast).
//   data [rank r]: float [2 slots][world][N][2m]   slot = (seq_base + step) & 1, second index = source rank
//   flags[rank r]: uint32 [world][n_groups]        last sequence number published by a source rank for a sample group
struct XchgDev {
    int world, rank;        // world == 0: no exchange
    unsigned seq_base;      // sequence number of step 0 of this trajectory (monotonic over the life of the buffers)
    float* data[D4S_MAX_RANKS];
    unsigned* flags[D4S_MAX_RANKS];
};

// fused tcgen05 trajectory kernel (d4s_traj_tc.cu)
struct TrajParams {
    NetDev net;
    int N, n, C, SG, OC, P;
    int RS;                  // tile row of (sample si, observation oj) = si * RS + oj; RS == OC, or OC rounded up to a power of two
    int step_begin, step_end;
    int pairing;             // 1: a CTA that owns >= 2 sample groups interleaves them step by step (software pipeline)
    int cta_pair;            // 1: launched as clusters of 2 CTAs sharing one cta_group::2 MMA stream (traj_tc_kernel<MT, true>)
    const float* obs_feat;   // [n][Hp0]
    const float* obs_prec;   // [n][m][m] or NULL
    const float* obs_weight; // [n] or NULL
    const float* sched;      // [steps][16]
    const float* time_feat;  // [steps][Hp0]
    const float* mats;       // [steps][2 m^2 + m]
    const float* noise;      // [steps][N][m] or NULL
    FinalizeParams fin;      // constant part of the per-step finalize parameters
    float* theta;            // [N][m] in/out
    float* part_out;         // export mode (per-step API / observation sharding): [N][2m] sums S1 | S2, no update
    XchgDev xg;              // in-kernel exchange of the per-sample sums between the ranks of an observation-sharded run
    int* error_flag;
    long long* prof;         // optional [16] phase cycle counters (CTA 0)
    long long* trace;        // -DD4S_TRACE builds only: [20 warps][256][tag, clock] event timeline of CTA 0 (tests/gpu_trace_tc.py)
};
cudaError_t launch_traj_tc(const TrajParams& p, int n_ctas, cudaStream_t stream);
int traj_tc_smem_bytes();

// tcgen05 score + Jacobian kernel of one JAC step (d4s_jac_tc.cu).  Units = (sample, chunk of OC observations), UT
// units per 128-row tile; a pair takes 1 + m tile rows, PW pairs per TMEM lane quarter.
struct JacTcParams {
    NetDev net;
    int N, n, C, OC, UT, PW;
    int n_tiles;
    int cta_pair;             // 1: clusters of 2 CTAs sharing one cta_group::2 MMA stream (jac_tc_kernel<true>)
    int ctx_samples;          // batched contexts (D4sProblem.ctx_samples): > 0 = samples per context, obs_feat is [B n][Hp0]
    const float* theta;       // [N][m]
    const float* obs_feat;    // [n][Hp0]
    const float* time_feat;   // [Hp0] row of this step
    const float* sched;       // [16] row of this step
    const float* obs_weight;  // [n] or NULL
    float* part;              // [N][C][m + m^2] sums over the chunk's observations of P s | P
    const float* base_pre;    // nets with a theta-embedding MLP: [N][Hp0] A e(theta_i) and
    const float* tan_pre;     //   [N][m][Hp0] A d e / d theta_k from theta_embed_kernel (NULL otherwise)
    int* error_flag;
    long long* prof;          // optional [16] cycle counters of CTA 0 (d4s_set_option("profile", 1); tests/gpu_debug_jac_prof.py)
};
cudaError_t launch_jac_tc(const JacTcParams& p, int n_ctas, cudaStream_t stream);

// tcgen05 trajectory kernel of the paired mode (d4s_paired_tc.cu): row i = (sample i, observation i / obs_div), plain score
struct PairedTcParams {
    NetDev net;
    long long N;              // rows = posterior samples
    int obs_div;              // row i reads obs_feat[i / obs_div] (1: one observation row per sample)
    int step_begin, step_end;
    const float* obs_feat;    // [ceil(N / obs_div)][Hp0]
    const float* time_feat;   // [steps][Hp0]
    const float* sched;       // [steps][16]
    const float* noise;       // [steps][N][m] or NULL (Philox)
    float* theta;             // [N][m] in/out
    unsigned long long seed;
    long long sample_offset;
    float clip_lo, clip_hi;
    int* error_flag;
};
cudaError_t launch_paired_tc(const PairedTcParams& p, int n_ctas, cudaStream_t stream);

// generic path (d4s_generic.cu): kernel 3b + 4 for theta_dim <= D4S_MAX_THETA_WIDE
struct WideParams {
    long long N;
    int m, layout_jac, prior_kind, update;
    const float* part;    // [N][2m] or [N][m + m*m]
    const float* sched;   // [16]
    const float* mats;    // Lmat | G | mu_t, or NULL
    const float* low;
    const float* high;
    const float* ext_s0;  // [N][m] prior score from the caller's closure (D4S_PRIOR_EXTERNAL)
    const float* ext_p0;  // [N][m][m] prior precision (weighted steps), or NULL
    const float* noise;   // [N][m] or NULL
    uint64_t seed;
    int step_index;
    long long sample_offset;
    float clip_lo, clip_hi;
    float* theta;
    float* score_out;
};
cudaError_t launch_finalize_wide(const WideParams& p, cudaStream_t stream);
cudaError_t launch_reduce_pairs(const float* scores, const float* pair_prec, const float* obs_prec, const float* obs_weight,
                                long long N, int n, int m, float* part, cudaStream_t stream);

cudaError_t launch_theta_embed(const NetDev& net, const float* theta, int N, int jac, float* base_pre, float* tan_pre,
                               cudaStream_t stream);
cudaError_t launch_score_simt(const ScoreParams& p, int tiles, cudaStream_t stream);
cudaError_t launch_finalize(const FinalizeParams& p, cudaStream_t stream);
cudaError_t launch_batched_inverse(const float* a, float* out, int batch, int m, cudaStream_t stream);
cudaError_t launch_prior_score(int kind, const float* theta, long long N, int m, const float* sched, const float* mats,
                               const float* low, const float* high, float* score_out, float* jacdiag_out,
                               cudaStream_t stream);
cudaError_t launch_ddim_update(float* theta, const float* score, long long N, int m, const float* sched,
                               const float* noise, uint64_t seed, int step, long long sample_offset, float clip_lo,
                               float clip_hi, cudaStream_t stream);
cudaError_t launch_philox_normal(float* out, long long n_rows, int m, uint64_t seed, int step, long long sample_offset,
                                 cudaStream_t stream);

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller.  Counter = (dim_pair, sample_lo, sample_hi, step), key = seed.
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
#else
    uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    uint32_t n0 = hi1 ^ c[1] ^ k[0];
    uint32_t n1 = lo1;
    uint32_t n2 = hi0 ^ c[3] ^ k[1];
    uint32_t n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
}

__host__ __device__ inline void philox4x32_10(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                             uint32_t (&out)[4]) {
    uint32_t c[4] = {c0, c1, c2, c3};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int r = 0; r < 10; ++r) philox_round(c, k);
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// Four standard normals for (step, sample, quad) where quad covers dims 4*quad .. 4*quad+3.
__device__ inline void philox_normal4(uint64_t seed, int step, int64_t sample, int quad, float (&z)[4]) {
    uint32_t r[4];
    philox4x32_10(seed, (uint32_t)quad, (uint32_t)(uint64_t)sample, (uint32_t)((uint64_t)sample >> 32),
                  (uint32_t)step, r);
    const float two_pow_m32 = 2.3283064365386963e-10f;
    // u in (0, 1]: (r + 0.5) * 2^-32 keeps log() finite
    float u0 = ((float)r[0] + 0.5f) * two_pow_m32;
    float u1 = ((float)r[1] + 0.5f) * two_pow_m32;
    float u2 = ((float)r[2] + 0.5f) * two_pow_m32;
    float u3 = ((float)r[3] + 0.5f) * two_pow_m32;
    u0 = fminf(u0, 1.0f); u2 = fminf(u2, 1.0f);
    float rad0 = sqrtf(-2.0f * logf(u0)), rad1 = sqrtf(-2.0f * logf(u2));
    float s0, c0, s1, c1;
    sincospif(2.0f * u1, &s0, &c0);
    sincospif(2.0f * u3, &s1, &c1);
    z[0] = rad0 * c0; z[1] = rad0 * s0; z[2] = rad1 * c1; z[3] = rad1 * s1;
}

// ------------------------------------------------------------------------------------------------
// Small dense linear algebra in registers (M compile-time, fully unrolled).  Gauss-Jordan inverse with
// partial pivoting -- the pivoting strategy of the LU behind torch.linalg.inv that the reference calls
// (tall_posterior_sampler.py:42,94,154); row swaps are predicated selects so every index stays a
// compile-time constant.
// ------------------------------------------------------------------------------------------------
template <typename T, int M>
__device__ __forceinline__ void gj_inverse_t(T (&a)[M][M], T (&inv)[M][M]) {
#pragma unroll
    for (int r = 0; r < M; ++r)
#pragma unroll
        for (int c = 0; c < M; ++c) inv[r][c] = (r == c) ? T(1) : T(0);
#pragma unroll
    for (int k = 0; k < M; ++k) {
        int p = k;
        T best = a[k][k] < T(0) ? -a[k][k] : a[k][k];
#pragma unroll
        for (int r = k + 1; r < M; ++r) {
            T v = a[r][k] < T(0) ? -a[r][k] : a[r][k];
            if (v > best) { best = v; p = r; }
        }
#pragma unroll
        for (int r = k + 1; r < M; ++r) {
            if (r == p) {
#pragma unroll
                for (int c = 0; c < M; ++c) {
                    T t = a[k][c]; a[k][c] = a[r][c]; a[r][c] = t;
                    T u = inv[k][c]; inv[k][c] = inv[r][c]; inv[r][c] = u;
                }
            }
        }
        const T d = T(1) / a[k][k];
#pragma unroll
        for (int c = 0; c < M; ++c) { a[k][c] *= d; inv[k][c] *= d; }
#pragma unroll
        for (int r = 0; r < M; ++r) {
            if (r != k) {
                const T f = a[r][k];
#pragma unroll
                for (int c = 0; c < M; ++c) {
                    a[r][c] -= f * a[k][c];
                    inv[r][c] -= f * inv[k][c];
                }
            }
        }
    }
}

}  // namespace d4s
