// Generic path (SURVEY.md section 7 "arbitrary nse.net", section 8 a13): score networks that cannot be fused -- a Python
// closure inside FakeFNet (gen_gaussian_gaussian.py:59-72), modules other than Linear/ReLU stacks, theta_dim > 10 -- are
// evaluated by torch on the (sample x observation) grid; kernels 3 and 4 of the path stay native:
//
//   reduce_pairs_kernel   kernel 3a: the reduction over observations of the precision-weighted scores
//                         (tall_posterior_sampler.py:114-128 materialises [N, n, m, m] and nse.py:631-636 / nse_toy.py:619-622
//                         reduce it with torch; here one pass over the pair tensors, coalesced, fixed summation order)
//   finalize_wide_kernel  kernel 3b + 4 for theta_dim <= 32: prior term (closed forms, or score / precision tensors
//                         returned by a caller's prior closure: nse_toy.py:618), Lambda solve (LU with partial pivoting,
//                         one warp per sample, one matrix row per lane, fp64), DDIM / Langevin update with Philox noise
//                         (nse.py:281-298, nse_toy.py:388-402).
#include "d4s_finalize_core.cuh"

namespace d4s {

constexpr int RP_THREADS = 256;
constexpr int RP_JB = 32;  // observations staged per chunk
constexpr int RP_Q = (D4S_MAX_THETA_WIDE * D4S_MAX_THETA_WIDE + RP_THREADS - 1) / RP_THREADS;  // matrix entries per thread

// One CTA per sample.  JAC layout (pair_prec != NULL): part[i] = [ sum_j w_j P_ij s_ij (m) | sum_j w_j P_ij (m*m) ];
// GAUSS layout: part[i] = [ sum_j w_j s_ij (m) | sum_j w_j Q_j s_ij (m) ] with Q = obs_prec (zeros when NULL).
__global__ void __launch_bounds__(RP_THREADS) reduce_pairs_kernel(const float* __restrict__ scores, const float* __restrict__ pair_prec,
                                                                  const float* __restrict__ obs_prec, const float* __restrict__ obs_weight,
                                                                  int n, int m, float* __restrict__ part) {
    __shared__ float s_sc[RP_JB * D4S_MAX_THETA_WIDE];
    __shared__ float s_red[D4S_MAX_THETA_WIDE * D4S_MAX_THETA_WIDE];
    const long long i = blockIdx.x;
    const int tid = threadIdx.x, mm = m * m;
    const bool jac = pair_prec != nullptr;
    const int W = jac ? m + mm : 2 * m;
    const float* sc_i = scores + i * (long long)n * m;
    const float* mat_i = jac ? pair_prec + i * (long long)n * mm : obs_prec;
    float acc_p[RP_Q], acc_ps[RP_Q];
#pragma unroll
    for (int q = 0; q < RP_Q; ++q) acc_p[q] = acc_ps[q] = 0.f;
    float acc_s1 = 0.f;  // GAUSS: thread r < m
    for (int j0 = 0; j0 < n; j0 += RP_JB) {
        const int nj = min(RP_JB, n - j0);
        __syncthreads();
        for (int e = tid; e < nj * m; e += RP_THREADS) s_sc[e] = sc_i[(long long)j0 * m + e];
        __syncthreads();
        for (int jj = 0; jj < nj; ++jj) {
            const float wj = obs_weight ? __ldg(obs_weight + j0 + jj) : 1.f;
            if (!jac && tid < m) acc_s1 = fmaf(wj, s_sc[jj * m + tid], acc_s1);
            if (mat_i) {
                const float* mj = mat_i + (long long)(j0 + jj) * mm;
#pragma unroll
                for (int q = 0; q < RP_Q; ++q) {
                    const int c = tid + q * RP_THREADS;
                    if (c < mm) {
                        const float pv = wj * __ldg(mj + c);
                        acc_p[q] += pv;
                        acc_ps[q] = fmaf(pv, s_sc[jj * m + (c % m)], acc_ps[q]);
                    }
                }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < RP_Q; ++q) {
        const int c = tid + q * RP_THREADS;
        if (c < mm) s_red[c] = acc_ps[q];
    }
    __syncthreads();
    float* out = part + i * (long long)W;
    if (tid < m) {
        float ps = 0.f;
        for (int k = 0; k < m; ++k) ps += s_red[tid * m + k];  // fixed order
        if (jac) {
            out[tid] = ps;
        } else {
            out[tid] = acc_s1;
            out[m + tid] = ps;
        }
    }
    if (jac) {
#pragma unroll
        for (int q = 0; q < RP_Q; ++q) {
            const int c = tid + q * RP_THREADS;
            if (c < mm) out[m + c] = acc_p[q];
        }
    }
}

cudaError_t launch_reduce_pairs(const float* scores, const float* pair_prec, const float* obs_prec, const float* obs_weight,
                                long long N, int n, int m, float* part, cudaStream_t stream) {
    reduce_pairs_kernel<<<(unsigned)N, RP_THREADS, 0, stream>>>(scores, pair_prec, obs_prec, obs_weight, n, m, part);
    return cudaGetLastError();
}

// ---- kernel 3b + 4 for theta_dim <= 32 ---------------------------------------------------------------------------------------
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// One warp per sample, lane r = component r of theta / row r of Lambda.  MW = 16 or 32 (compile-time width of the row
// held in registers); m <= MW at run time, loops are unrolled over MW and guarded by the (warp-uniform) bound m.
template <int MW>
__global__ void __launch_bounds__(128) finalize_wide_kernel(const WideParams p) {
    const int lane = threadIdx.x & 31;
    const long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= p.N) return;  // whole warps leave together
    const int m = p.m, r = lane;
    const bool act = r < m;
    const int W = p.layout_jac ? m + m * m : 2 * m;
    const float* part = p.part + i * (long long)W;
    const int combine = (int)p.sched[D4S_S_COMBINE];
    const double aos2 = (double)p.sched[D4S_S_A_OVER_S2];
    const double omn = (double)p.sched[D4S_S_ONE_MINUS_N];
    const bool per_sample = combine == D4S_COMBINE_PER_SAMPLE;
    const double th = act ? (double)p.theta[i * m + r] : 0.0;

    double S1 = 0.0, S2 = 0.0;
    double L[MW];
#pragma unroll
    for (int k = 0; k < MW; ++k) L[k] = 0.0;
    if (act) {
        if (!p.layout_jac) {
            S1 = (double)part[r];
            S2 = (double)part[m + r];
        } else {
            S2 = (double)part[r];
        }
        if (per_sample) {
#pragma unroll
            for (int k = 0; k < MW; ++k)
                if (k < m)
                    L[k] = (p.layout_jac ? (double)part[m + r * m + k] : 0.0) + (p.mats ? (double)p.mats[r * m + k] : 0.0);
        }
    }
    // ---- prior term: g = (1-n) s0 (SUM) or (1-n) P0 s0 (weighted); Lambda += (1-n) P0 where it depends on the sample ----
    double g = 0.0;
    if (p.prior_kind == D4S_PRIOR_GAUSSIAN) {  // g = G (theta - mu_t), G = -(1-n) [P0] C^-T written by the host
        const double d = act ? th - (double)p.mats[2 * m * m + r] : 0.0;
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < MW; ++c) {
            if (c < m) {
                const double dc = shfl_d(d, c);
                if (act) acc = fma((double)p.mats[m * m + r * m + c], dc, acc);
            }
        }
        g = acc;
    } else if (p.prior_kind == D4S_PRIOR_UNIFORM) {
        if (act) {
            double s0, jac;
            const double sig = (double)p.sched[D4S_S_SIGMA];
            uniform_prior_terms(th, (double)p.low[r], (double)p.high[r], (double)p.sched[D4S_S_SQRT_ALPHA], sig, s0, jac);
            if (per_sample) {
                const double p0 = aos2 / (1.0 + sig * sig * jac);
#pragma unroll
                for (int k = 0; k < MW; ++k)
                    if (k == r) L[k] += omn * p0;
                g = omn * p0 * s0;
            } else {
                g = omn * s0;
            }
        }
    } else if (p.prior_kind == D4S_PRIOR_EXTERNAL) {  // prior closure evaluated by the caller (nse_toy.py:618)
        const double s0 = act ? (double)p.ext_s0[i * m + r] : 0.0;
        if (combine == D4S_COMBINE_SUM || !p.ext_p0) {
            g = omn * s0;
        } else {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < MW; ++k) {
                if (k < m) {
                    const double sk = shfl_d(s0, k);
                    if (act) {
                        const double pk = (double)p.ext_p0[(i * m + r) * m + k];
                        acc = fma(pk, sk, acc);
                        if (per_sample) L[k] += omn * pk;
                    }
                }
            }
            g = omn * acc;
        }
    }
    // ---- combine (nse.py:628-654, nse_toy.py:619-625) -----------------------------------------------------------------
    double out;
    if (combine == D4S_COMBINE_SUM) {
        out = g + (p.layout_jac ? 0.0 : (double)p.sched[D4S_S_S1_WEIGHT] * S1);  // (plain-sum steps use the GAUSS layout)
    } else {
        double w = (p.layout_jac ? S2 : fma(aos2, S1, S2)) + g;
        if (combine == D4S_COMBINE_SHARED) {
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < MW; ++c) {
                if (c < m) {
                    const double wc = shfl_d(w, c);
                    if (act) acc = fma((double)p.mats[r * m + c], wc, acc);
                }
            }
            out = acc;
        } else {  // Lambda_i x = w: LU with partial pivoting, row r in lane r
#pragma unroll
            for (int k = 0; k < MW; ++k) {
                if (k < m) {
                    double v = (r >= k && act) ? fabs(L[k]) : -1.0;
                    int idx = r;
#pragma unroll
                    for (int sh = 16; sh > 0; sh >>= 1) {
                        const double ov = __shfl_xor_sync(0xffffffffu, v, sh);
                        const int oi = __shfl_xor_sync(0xffffffffu, idx, sh);
                        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
                    }
                    if (idx != k) {  // swap rows k and idx (uniform decision)
                        const int src = (r == k) ? idx : (r == idx) ? k : r;
#pragma unroll
                        for (int c = 0; c < MW; ++c)
                            if (c < m) L[c] = shfl_d(L[c], src);
                        w = shfl_d(w, src);
                    }
                    const double inv = 1.0 / shfl_d(L[k], k);
                    const double f = (r > k && act) ? L[k] * inv : 0.0;
#pragma unroll
                    for (int c = 0; c < MW; ++c) {
                        if (c > k && c < m) {
                            const double pc = shfl_d(L[c], k);
                            L[c] = fma(-f, pc, L[c]);
                        }
                    }
                    w = fma(-f, shfl_d(w, k), w);
                }
            }
#pragma unroll
            for (int k = MW - 1; k >= 0; --k) {
                if (k < m) {
                    const double xk = shfl_d(w / L[k], k);
                    if (r < k) w = fma(-L[k], xk, w);
                    if (r == k) w = xk;
                }
            }
            out = w;
        }
    }
    if (p.score_out && act) p.score_out[i * m + r] = (float)out;
    // ---- update: theta' = c_theta theta + c_score score + bridge_std z (nse.py:289-298) ----------------------------------
    if (p.update) {
        double c_score = (double)p.sched[D4S_S_C_SCORE];
        const double tame = (double)p.sched[D4S_S_TAME];
        if (tame > 0.0) {
            double nrm = act ? out * out : 0.0;
#pragma unroll
            for (int sh = 16; sh > 0; sh >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, sh);
            c_score = tame / (1.0 + tame * sqrt(nrm));
        }
        const float bstd = p.sched[D4S_S_BRIDGE_STD];
        float z = 0.f;
        if (act) {
            if (p.noise) {
                z = p.noise[i * m + r];
            } else if (bstd != 0.f) {
                float zz[4];
                philox_normal4(p.seed, p.step_index, p.sample_offset + i, r >> 2, zz);
                z = zz[r & 3];
            }
            float v = (float)((double)p.sched[D4S_S_C_THETA] * th + c_score * out + (double)bstd * (double)z);
            if (p.clip_lo <= p.clip_hi && p.sched[D4S_S_CLIP] != 0.f) v = fminf(fmaxf(v, p.clip_lo), p.clip_hi);
            p.theta[i * m + r] = v;
        }
    }
}

cudaError_t launch_finalize_wide(const WideParams& p, cudaStream_t stream) {
    const int warps = 4;
    const long long blocks = (p.N + warps - 1) / warps;
    if (p.m <= 16) finalize_wide_kernel<16><<<(unsigned)blocks, warps * 32, 0, stream>>>(p);
    else finalize_wide_kernel<32><<<(unsigned)blocks, warps * 32, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace d4s
