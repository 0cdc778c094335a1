// Per-sample finalize core shared by the stand-alone finalize kernel (d4s_finalize.cu) and the fused
// tcgen05 trajectory kernel (d4s_traj_tc.cu).  See d4s_finalize.cu for the reference citations.
#pragma once
#include "d4s_common.cuh"

namespace d4s {

// LU with partial pivoting, fully unrolled (row swaps as predicated moves: everything stays in registers).
template <typename T, int M>
__device__ __forceinline__ void lu_solve_t(T (&a)[M][M], T (&b)[M]) {
#pragma unroll
    for (int k = 0; k < M; ++k) {
        int p = k;
        T best = fabs(a[k][k]);
#pragma unroll
        for (int r = k + 1; r < M; ++r) {
            T v = fabs(a[r][k]);
            if (v > best) { best = v; p = r; }
        }
#pragma unroll
        for (int r = k + 1; r < M; ++r) {
            if (r == p) {
#pragma unroll
                for (int c = 0; c < M; ++c) { T t = a[k][c]; a[k][c] = a[r][c]; a[r][c] = t; }
                T t = b[k]; b[k] = b[r]; b[r] = t;
            }
        }
        const T inv = T(1) / a[k][k];
#pragma unroll
        for (int r = k + 1; r < M; ++r) {
            const T f = a[r][k] * inv;
#pragma unroll
            for (int c = k + 1; c < M; ++c) a[r][c] = fma(-f, a[k][c], a[r][c]);
            b[r] = fma(-f, b[k], b[r]);
        }
    }
#pragma unroll
    for (int k = M - 1; k >= 0; --k) {
        T s = b[k];
#pragma unroll
        for (int c = k + 1; c < M; ++c) s = fma(-a[k][c], b[c], s);
        b[k] = s / a[k][k];
    }
}
template <int M>
__device__ __forceinline__ void lu_solve_d(double (&a)[M][M], double (&b)[M]) { lu_solve_t<double, M>(a, b); }

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

// Per-sample tail of the step: reduce chunk partials, prior term, combine/solve, optional DDIM update.
//   part      : this sample's partial sums, [C][W]
//   theta_io  : this sample's theta[M] (read; written when p.update)
//   score_out : optional [M]
//   noise     : optional injected z[M]
//   pre_prior : optional [2][M] doubles {s0[k]}, {jac[k]} of the uniform prior, evaluated by other threads
//   ctx_qsum  : optional [M][M] observation part of Lambda of this sample's context (batched contexts)
template <int M>
__device__ __forceinline__ void finalize_sample(const FinalizeParams& p, const float* part, float* theta_io,
                                                float* score_out, const float* noise, long long global_sample,
                                                const double* pre_prior = nullptr, const float* ctx_qsum = nullptr) {
    constexpr int WG = 2 * M, WJ = M + M * M;
    const int combine = (int)p.sched[D4S_S_COMBINE];
    const double aos2 = (double)p.sched[D4S_S_A_OVER_S2];
    const double one_minus_n = (double)p.sched[D4S_S_ONE_MINUS_N];

    float th[M];
#pragma unroll
    for (int k = 0; k < M; ++k) th[k] = theta_io[k];

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
            for (int c = 0; c < M; ++c)  // batched contexts: + sum_j inv(Sigma_cj) of the sample's context
                Lam[r][c] = (double)p.mats[r * M + c] + (ctx_qsum ? (double)ctx_qsum[r * M + c] : 0.0);
    }
    const float* pp = part;
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
            if (pre_prior) {
                s0 = pre_prior[k];
                jac = pre_prior[M + k];
            } else {
                uniform_prior_terms((double)th[k], (double)p.low[k], (double)p.high[k], s, sig, s0, jac);
            }
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
        for (int k = 0; k < M; ++k) out[k] = g[k] + (p.jac ? 0.0 : (double)p.sched[D4S_S_S1_WEIGHT] * S1[k]);
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
    if (score_out) {
#pragma unroll
        for (int k = 0; k < M; ++k) score_out[k] = (float)out[k];
    }

    // ---- DDIM update: theta' = c_theta theta + c_score score + bridge_std z -----------------------------
    if (p.update) {
        const double c_theta = (double)p.sched[D4S_S_C_THETA];
        double c_score = (double)p.sched[D4S_S_C_SCORE];
        const double tame = (double)p.sched[D4S_S_TAME];
        if (tame > 0.0) {  // tamed ULA step (nse.py:336): eps / (1 + eps ||g||)
            double nrm = 0.0;
#pragma unroll
            for (int k = 0; k < M; ++k) nrm = fma(out[k], out[k], nrm);
            c_score = tame / (1.0 + tame * sqrt(nrm));
        }
        const float bstd = p.sched[D4S_S_BRIDGE_STD];
        float z[(M + 3) / 4 * 4];
        if (noise) {
#pragma unroll
            for (int k = 0; k < M; ++k) z[k] = noise[k];
        } else if (bstd != 0.f) {
#pragma unroll
            for (int qd = 0; qd < (M + 3) / 4; ++qd) {
                float zz[4];
                philox_normal4(p.seed, p.step_index, global_sample, qd, zz);
                z[4 * qd] = zz[0]; z[4 * qd + 1] = zz[1]; z[4 * qd + 2] = zz[2]; z[4 * qd + 3] = zz[3];
            }
        } else {
#pragma unroll
            for (int k = 0; k < M; ++k) z[k] = 0.f;
        }
        const bool clip = (p.clip_lo <= p.clip_hi) && p.sched[D4S_S_CLIP] != 0.f;
#pragma unroll
        for (int k = 0; k < M; ++k) {
            float v = (float)(c_theta * (double)th[k] + c_score * out[k] + (double)bstd * (double)z[k]);
            if (clip) v = fminf(fmaxf(v, p.clip_lo), p.clip_hi);
            theta_io[k] = v;
        }
    }
}

}  // namespace d4s
