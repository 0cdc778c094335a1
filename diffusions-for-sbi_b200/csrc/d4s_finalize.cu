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
#include "d4s_finalize_core.cuh"

namespace d4s {


template <int M>
__global__ void __launch_bounds__(128) finalize_kernel(const FinalizeParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N) return;
    finalize_sample<M>(p, p.part + (size_t)i * p.C * p.W, p.theta + (size_t)i * M,
                       p.score_out ? p.score_out + (size_t)i * M : nullptr,
                       p.noise ? p.noise + (size_t)i * M : nullptr, p.sample_offset + i, nullptr,
                       (p.ctx_qsum && p.ctx_samples > 0) ? p.ctx_qsum + (size_t)(i / p.ctx_samples) * M * M : nullptr);
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
    const double c_theta = (double)sched[D4S_S_C_THETA];
    double c_score = (double)sched[D4S_S_C_SCORE];
    const double tame = (double)sched[D4S_S_TAME];
    if (tame > 0.0) {  // tamed ULA (nse.py:336)
        double nrm = 0.0;
        for (int k = 0; k < m; ++k) nrm = fma((double)score[i * m + k], (double)score[i * m + k], nrm);
        c_score = tame / (1.0 + tame * sqrt(nrm));
    }
    const float bstd = sched[D4S_S_BRIDGE_STD];
    const bool clip = clip_lo <= clip_hi && sched[D4S_S_CLIP] != 0.f;
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
