// Kernel 3+4: per-sample reduction of the observation-chunk partial sums, diffused prior score,
// assembly and solve of the m x m system Lambda_i out_i = w_i, and the fused DDIM update with Philox noise.
//
// Replaces (reference file:line):
//   prior_score_fn(theta, t)                         vp_diffused_priors.py:20-46 (uniform), :55-72 (gaussian)
//   tweedies_approximation_prior                     tall_posterior_sampler.py:138-155 (closed form, diagonal)
//   Lambda, weighted scores, linalg.solve            nse.py:628-638 / 640-654
//   ddim_step: mean_pred, bridge_mean, randn_like    nse.py:281-298, 199-221 ; clip nse.py:265-266
//
// One thread per posterior sample; the m x m work is register resident (M is a template parameter).
// The small solves and the erf-based prior score run in fp64 (a few hundred flops per sample): they are
// the ill-conditioned part of the path (SURVEY.md Finding B) and cost nothing next to kernel 1.
#include "d4s_common.cuh"

namespace d4s {


template <int M>
__device__ __forceinline__ void lu_solve_d(double (&a)[M][M], double (&b)[M]) {
#pragma unroll
    for (int k = 0; k < M; ++k) {
        int p = k;
        double best = fabs(a[k][k]);
#pragma unroll
        for (int r = k + 1; r < M; ++r) {
            double v = fabs(a[r][k]);
            if (v > best) { best = v; p = r; }
        }
#pragma unroll
        for (int r = k + 1; r < M; ++r) {
            if (r == p) {
#pragma unroll
                for (int c = 0; c < M; ++c) { double t = a[k][c]; a[k][c] = a[r][c]; a[r][c] = t; }
                double t = b[k]; b[k] = b[r]; b[r] = t;
            }
        }
        const double inv = 1.0 / a[k][k];
#pragma unroll
        for (int r = k + 1; r < M; ++r) {
            const double f = a[r][k] * inv;
#pragma unroll
            for (int c = k + 1; c < M; ++c) a[r][c] = fma(-f, a[k][c], a[r][c]);
            b[r] = fma(-f, b[k], b[r]);
        }
    }
#pragma unroll
    for (int k = M - 1; k >= 0; --k) {
        double s = b[k];
#pragma unroll
        for (int c = k + 1; c < M; ++c) s = fma(-a[k][c], b[c], s);
        b[k] = s / a[k][k];
    }
}

// Phi(ub) - Phi(ua) evaluated on the tail that avoids cancellation.
__device__ __forceinline__ double cdf_diff(double ua, double ub) {
    const double r = 0.70710678118654752440;
    if (ua > 0.0) return 0.5 * (erfc(ua * r) - erfc(ub * r));
    if (ub < 0.0) return 0.5 * (erfc(-ub * r) - erfc(-ua * r));
    return 0.5 * (erf(ub * r) - erf(ua * r));
}

// Diffused-uniform prior, one dimension: score s0 = f'/(f + 1e-6) and its derivative d s0 / d theta
// (vp_diffused_priors.py:29-46; the derivative is what jacrev yields at tall_posterior_sampler.py:147).
__device__ __forceinline__ void uniform_prior_terms(double th, double lo, double hi, double s, double sig, double& s0,
                                                    double& jac) {
    const double inv_sqrt_2pi = 0.39894228040143267794;
    const double ub = (hi * s - th) / sig;
    const double ua = (lo * s - th) / sig;
    const double f = cdf_diff(ua, ub) / s;
    const double pb = inv_sqrt_2pi * exp(-0.5 * ub * ub), pa = inv_sqrt_2pi * exp(-0.5 * ua * ua);
    const double fp = -(pb - pa) / (sig * s);
    const double den = f + 1e-6;  // vp_diffused_priors.py:44
    s0 = fp / den;
    const double fpp = -(ub * pb - ua * pa) / (sig * sig * s);
    jac = fpp / den - fp * fp / (den * den);
}

template <int M>
__global__ void __launch_bounds__(128) finalize_kernel(const FinalizeParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N) return;
    constexpr int WG = 2 * M, WJ = M + M * M;
    const int combine = (int)p.sched[D4S_S_COMBINE];
    const double aos2 = (double)p.sched[D4S_S_A_OVER_S2];
    const double one_minus_n = (double)p.sched[D4S_S_ONE_MINUS_N];

    float th[M];
#pragma unroll
    for (int k = 0; k < M; ++k) th[k] = p.theta[(size_t)i * M + k];

    // ---- reduce the chunk partials (fixed order) ---------------------------------------------------
    double S1[M], S2[M];  // GAUSS: sum s, sum inv(Sigma_j) s ; JAC: S2 = sum P s, S1 unused
    double Lam[M][M];
#pragma unroll
    for (int k = 0; k < M; ++k) { S1[k] = 0.0; S2[k] = 0.0; }
    const bool per_sample = (combine == D4S_COMBINE_PER_SAMPLE);
    if (per_sample) {
#pragma unroll
        for (int r = 0; r < M; ++r)
#pragma unroll
            for (int c = 0; c < M; ++c) Lam[r][c] = (double)p.mats[r * M + c];
    }
    const float* pp = p.part + (size_t)i * p.C * p.W;
    if (!p.jac) {
        for (int c = 0; c < p.C; ++c) {
#pragma unroll
            for (int k = 0; k < M; ++k) {
                S1[k] += (double)pp[c * WG + k];
                S2[k] += (double)pp[c * WG + M + k];
            }
        }
    } else {
        for (int c = 0; c < p.C; ++c) {
#pragma unroll
            for (int k = 0; k < M; ++k) S2[k] += (double)pp[c * WJ + k];
            if (per_sample) {
#pragma unroll
                for (int r = 0; r < M; ++r)
#pragma unroll
                    for (int cc = 0; cc < M; ++cc) Lam[r][cc] += (double)pp[c * WJ + M + r * M + cc];
            }
        }
    }

    // ---- prior term ------------------------------------------------------------------------------------
    double g[M];  // SUM: (1-n) s0 ; weighted modes: (1-n) P0 s0
#pragma unroll
    for (int k = 0; k < M; ++k) g[k] = 0.0;
    if (p.prior_kind == D4S_PRIOR_GAUSSIAN) {
        const float* G = p.mats + M * M;
        const float* mu = p.mats + 2 * M * M;
        double d[M];
#pragma unroll
        for (int k = 0; k < M; ++k) d[k] = (double)th[k] - (double)mu[k];
#pragma unroll
        for (int r = 0; r < M; ++r) {
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < M; ++c) acc = fma((double)G[r * M + c], d[c], acc);
            g[r] = acc;
        }
    } else if (p.prior_kind == D4S_PRIOR_UNIFORM) {
        const double s = (double)p.sched[D4S_S_SQRT_ALPHA];
        const double sig = (double)p.sched[D4S_S_SIGMA];
#pragma unroll
        for (int k = 0; k < M; ++k) {
            double s0, jac;
            uniform_prior_terms((double)th[k], (double)p.low[k], (double)p.high[k], s, sig, s0, jac);
            if (per_sample) {
                const double p0 = aos2 / (1.0 + sig * sig * jac);  // inv(sigma^2/alpha (1 + sigma^2 J0))
                Lam[k][k] += one_minus_n * p0;
                g[k] = one_minus_n * p0 * s0;
            } else {
                g[k] = one_minus_n * s0;
            }
        }
    }

    // ---- combine -----------------------------------------------------------------------------------------
    double out[M];
    if (combine == D4S_COMBINE_SUM) {
#pragma unroll
        for (int k = 0; k < M; ++k) out[k] = g[k] + (p.jac ? 0.0 : S1[k]);
    } else {
        double w[M];
#pragma unroll
        for (int k = 0; k < M; ++k) w[k] = (p.jac ? S2[k] : fma(aos2, S1[k], S2[k])) + g[k];
        if (combine == D4S_COMBINE_SHARED) {
#pragma unroll
            for (int r = 0; r < M; ++r) {
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < M; ++c) acc = fma((double)p.mats[r * M + c], w[c], acc);
                out[r] = acc;
            }
        } else {
            lu_solve_d<M>(Lam, w);
#pragma unroll
            for (int k = 0; k < M; ++k) out[k] = w[k];
        }
    }
    if (p.score_out) {
#pragma unroll
        for (int k = 0; k < M; ++k) p.score_out[(size_t)i * M + k] = (float)out[k];
    }

    // ---- DDIM update: theta' = c_theta theta + c_score score + bridge_std z -----------------------------
    if (p.update) {
        const double c_theta = (double)p.sched[D4S_S_C_THETA], c_score = (double)p.sched[D4S_S_C_SCORE];
        const float bstd = p.sched[D4S_S_BRIDGE_STD];
        float z[(M + 3) / 4 * 4];
        if (p.noise) {
#pragma unroll
            for (int k = 0; k < M; ++k) z[k] = p.noise[(size_t)i * M + k];
        } else if (bstd != 0.f) {
#pragma unroll
            for (int qd = 0; qd < (M + 3) / 4; ++qd) {
                float zz[4];
                philox_normal4(p.seed, p.step_index, p.sample_offset + i, qd, zz);
                z[4 * qd] = zz[0]; z[4 * qd + 1] = zz[1]; z[4 * qd + 2] = zz[2]; z[4 * qd + 3] = zz[3];
            }
        } else {
#pragma unroll
            for (int k = 0; k < M; ++k) z[k] = 0.f;
        }
        const bool clip = p.clip_lo <= p.clip_hi;
#pragma unroll
        for (int k = 0; k < M; ++k) {
            float v = (float)(c_theta * (double)th[k] + c_score * out[k] + (double)bstd * (double)z[k]);
            if (clip) v = fminf(fmaxf(v, p.clip_lo), p.clip_hi);
            p.theta[(size_t)i * M + k] = v;
        }
    }
}

cudaError_t launch_finalize(const FinalizeParams& p, cudaStream_t stream) {
    const int threads = 128;
    const int blocks = (p.N + threads - 1) / threads;
    switch (p.m) {
        case 1: finalize_kernel<1><<<blocks, threads, 0, stream>>>(p); break;
        case 2: finalize_kernel<2><<<blocks, threads, 0, stream>>>(p); break;
        case 3: finalize_kernel<3><<<blocks, threads, 0, stream>>>(p); break;
        case 4: finalize_kernel<4><<<blocks, threads, 0, stream>>>(p); break;
        case 5: finalize_kernel<5><<<blocks, threads, 0, stream>>>(p); break;
        case 6: finalize_kernel<6><<<blocks, threads, 0, stream>>>(p); break;
        case 7: finalize_kernel<7><<<blocks, threads, 0, stream>>>(p); break;
        case 8: finalize_kernel<8><<<blocks, threads, 0, stream>>>(p); break;
        case 9: finalize_kernel<9><<<blocks, threads, 0, stream>>>(p); break;
        case 10: finalize_kernel<10><<<blocks, threads, 0, stream>>>(p); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// ---- standalone prior score (API parity with the reference's prior_score_fn closures) ---------------------
__global__ void prior_score_kernel(int kind, const float* __restrict__ theta, long long N, int m,
                                   const float* __restrict__ sched, const float* __restrict__ mats,
                                   const float* __restrict__ low, const float* __restrict__ high,
                                   float* __restrict__ score_out, float* __restrict__ jacdiag_out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * m) return;
    const long long i = e / m;
    const int k = (int)(e - i * m);
    if (kind == D4S_PRIOR_UNIFORM) {
        double s0, jac;
        uniform_prior_terms((double)theta[e], (double)low[k], (double)high[k], (double)sched[D4S_S_SQRT_ALPHA],
                            (double)sched[D4S_S_SIGMA], s0, jac);
        score_out[e] = (float)s0;
        if (jacdiag_out) jacdiag_out[e] = (float)jac;
    } else {  // gaussian: score = G (theta - mu_t) with G = -C^-T written by the host
        const float* G = mats + m * m;
        const float* mu = mats + 2 * m * m;
        double acc = 0.0;
        for (int c = 0; c < m; ++c) acc = fma((double)G[k * m + c], (double)theta[i * m + c] - (double)mu[c], acc);
        score_out[e] = (float)acc;
    }
}

cudaError_t launch_prior_score(int kind, const float* theta, long long N, int m, const float* sched, const float* mats,
                               const float* low, const float* high, float* score_out, float* jacdiag_out,
                               cudaStream_t stream) {
    const int threads = 256;
    const long long blocks = (N * m + threads - 1) / threads;
    prior_score_kernel<<<(unsigned)blocks, threads, 0, stream>>>(kind, theta, N, m, sched, mats, low, high, score_out,
                                                                 jacdiag_out);
    return cudaGetLastError();
}

// ---- standalone DDIM update (NSE.ddim_step with a caller-supplied score, nse.py:281-298) ---------------------
__global__ void ddim_update_kernel(float* __restrict__ theta, const float* __restrict__ score, long long N, int m,
                                   const float* __restrict__ sched, const float* __restrict__ noise, uint64_t seed,
                                   int step, long long sample_offset, float clip_lo, float clip_hi) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double c_theta = (double)sched[D4S_S_C_THETA], c_score = (double)sched[D4S_S_C_SCORE];
    const float bstd = sched[D4S_S_BRIDGE_STD];
    const bool clip = clip_lo <= clip_hi;
    for (int qd = 0; qd < (m + 3) / 4; ++qd) {
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        if (!noise && bstd != 0.f) philox_normal4(seed, step, sample_offset + i, qd, z);
        for (int j = 0; j < 4; ++j) {
            const int k = 4 * qd + j;
            if (k >= m) break;
            const double zz = noise ? (double)noise[i * m + k] : (double)z[j];
            float v = (float)(c_theta * (double)theta[i * m + k] + c_score * (double)score[i * m + k] + (double)bstd * zz);
            if (clip) v = fminf(fmaxf(v, clip_lo), clip_hi);
            theta[i * m + k] = v;
        }
    }
}

cudaError_t launch_ddim_update(float* theta, const float* score, long long N, int m, const float* sched,
                               const float* noise, uint64_t seed, int step, long long sample_offset, float clip_lo,
                               float clip_hi, cudaStream_t stream) {
    const int threads = 128;
    const long long blocks = (N + threads - 1) / threads;
    ddim_update_kernel<<<(unsigned)blocks, threads, 0, stream>>>(theta, score, N, m, sched, noise, seed, step,
                                                                 sample_offset, clip_lo, clip_hi);
    return cudaGetLastError();
}

// ---- utilities ------------------------------------------------------------------------------------------
template <int M>
__global__ void batched_inverse_kernel(const float* __restrict__ a, float* __restrict__ out, int batch) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    double A[M][M], I[M][M];
#pragma unroll
    for (int r = 0; r < M; ++r)
#pragma unroll
        for (int c = 0; c < M; ++c) A[r][c] = (double)a[(size_t)b * M * M + r * M + c];
    gj_inverse_t<double, M>(A, I);
#pragma unroll
    for (int r = 0; r < M; ++r)
#pragma unroll
        for (int c = 0; c < M; ++c) out[(size_t)b * M * M + r * M + c] = (float)I[r][c];
}

cudaError_t launch_batched_inverse(const float* a, float* out, int batch, int m, cudaStream_t stream) {
    const int threads = 64, blocks = (batch + threads - 1) / threads;
    switch (m) {
        case 1: batched_inverse_kernel<1><<<blocks, threads, 0, stream>>>(a, out, batch); break;
        case 2: batched_inverse_kernel<2><<<blocks, threads, 0, stream>>>(a, out, batch); break;
        case 3: batched_inverse_kernel<3><<<blocks, threads, 0, stream>>>(a, out, batch); break;
        case 4: batched_inverse_kernel<4><<<blocks, threads, 0, stream>>>(a, out, batch); break;
        case 5: batched_inverse_kernel<5><<<blocks, threads, 0, stream>>>(a, out, batch); break;
        case 6: batched_inverse_kernel<6><<<blocks, threads, 0, stream>>>(a, out, batch); break;
        case 7: batched_inverse_kernel<7><<<blocks, threads, 0, stream>>>(a, out, batch); break;
        case 8: batched_inverse_kernel<8><<<blocks, threads, 0, stream>>>(a, out, batch); break;
        case 9: batched_inverse_kernel<9><<<blocks, threads, 0, stream>>>(a, out, batch); break;
        case 10: batched_inverse_kernel<10><<<blocks, threads, 0, stream>>>(a, out, batch); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

__global__ void philox_normal_kernel(float* __restrict__ out, long long n_rows, int m, uint64_t seed, int step,
                                     long long sample_offset) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    for (int qd = 0; qd < (m + 3) / 4; ++qd) {
        float z[4];
        philox_normal4(seed, step, sample_offset + i, qd, z);
        for (int k = 0; k < 4; ++k)
            if (4 * qd + k < m) out[i * m + 4 * qd + k] = z[k];
    }
}

cudaError_t launch_philox_normal(float* out, long long n_rows, int m, uint64_t seed, int step, long long sample_offset,
                                 cudaStream_t stream) {
    const int threads = 256;
    const long long blocks = (n_rows + threads - 1) / threads;
    philox_normal_kernel<<<(unsigned)blocks, threads, 0, stream>>>(out, n_rows, m, seed, step, sample_offset);
    return cudaGetLastError();
}

}  // namespace d4s
