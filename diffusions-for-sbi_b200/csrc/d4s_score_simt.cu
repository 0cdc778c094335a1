// Kernel 1+2 (fp32 SIMT variant): fused score-network forward on a tile of (sample x observation) pairs,
// optional forward-mode Jacobian (JAC), Tweedie precisions, and the per-sample partial sums over the
// tile's observations.  Nothing of shape [N, n, *] is materialised unless the caller asks for it.
//
// Replaces (reference file:line):
//   NSE.forward / NSE.score on the grid           nse.py:112-145 via tall_posterior_sampler.py:114-121
//   vmap(vmap(jacrev(score)))                      tall_posterior_sampler.py:76-85  (forward mode here)
//   cov -> inv (JAC)                               tall_posterior_sampler.py:91-94
//   prec @ scores, .sum(dim=1)                     nse.py:631-636 / 647-652
//
// Layer 0 is separable: z0[i,j,:] = A theta_i + obs_feat_j + time_feat (SURVEY.md section 7 identity 1),
// so it costs m FMAs per element instead of a GEMM.  Hidden layers run as a register-tiled fp32 GEMM:
// warp w owns rows 8w..8w+7 of the tile and all output columns (lane l -> columns l, l+32, ...), the
// activations live in shared memory k-major and are updated in place, the weights stream from L2 through
// a cp.async ring.  JAC rows (one primal + m tangents per pair) share the GEMM; tangents are masked by the
// primal's ReLU pattern, published per layer as ballot words.
#include "d4s_finalize_core.cuh"

namespace d4s {

constexpr int NSTAGE = 3;


__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Output columns of a lane in the hidden layers: groups of G = min(NC, 4) consecutive columns, so that a lane reads its
// weights of one k with one vector load per group: col(lane, j) = G lane + j % G + 32 G (j / G).
template <int NC>
__device__ __forceinline__ int col_of(int lane, int j) {
    constexpr int G = NC >= 4 ? 4 : NC;
    return G * lane + (j % G) + 32 * G * (j / G);
}

// One hidden layer: acc[q][j] = sum_k act[k][8*warp+q] * Wt[k][col_of(lane, j)]
template <int NC>
__device__ __forceinline__ void layer_gemm(float (&acc)[8][NC], const float* __restrict__ Wt, int K,
                                           const float* sAct, float* sW, int warp, int lane, int tid) {
    constexpr int Hout = NC * 32;
    constexpr int CHUNK_F4 = KC * Hout / 4;  // float4 per stage
    // accumulators as packed fp32 pairs over two rows: one FFMA2 does (acc[q][j], acc[q+1][j]) += (a[q], a[q+1]) * w[j],
    // which halves the issue slots of the inner loop (the kernel is issue bound, not FMA-pipe bound)
    float2 acc2[4][NC];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < NC; ++j) acc2[q][j] = make_float2(0.f, 0.f);

    __syncthreads();  // every warp is done with the previous layer's weight ring before it is refilled
    const int nchunks = K / KC;
    auto issue = [&](int ch) {
        if (ch < nchunks) {
            const float4* src = reinterpret_cast<const float4*>(Wt + (size_t)ch * KC * Hout);
            float4* dst = reinterpret_cast<float4*>(sW + (ch % NSTAGE) * KC * HMAX);
            for (int i = tid; i < CHUNK_F4; i += NT) cp_async16(dst + i, src + i);
        }
        cp_async_commit();
    };
    issue(0);
    issue(1);
    for (int ch = 0; ch < nchunks; ++ch) {
        cp_async_wait<1>();
        __syncthreads();
        issue(ch + 2);
        const float* w = sW + (ch % NSTAGE) * KC * HMAX;
        const float* a = sAct + (size_t)ch * KC * ASTR + 8 * warp;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            float4 a0 = *reinterpret_cast<const float4*>(a + kk * ASTR);
            float4 a1 = *reinterpret_cast<const float4*>(a + kk * ASTR + 4);
            const float2 av[4] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w), make_float2(a1.x, a1.y), make_float2(a1.z, a1.w)};
            float2 bv[NC];
            constexpr int G = NC >= 4 ? 4 : NC;
#pragma unroll
            for (int g = 0; g < NC / G; ++g) {
                const float* wp = w + kk * Hout + G * lane + 32 * G * g;
                if constexpr (G == 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(wp);
                    bv[4 * g] = make_float2(b4.x, b4.x);
                    bv[4 * g + 1] = make_float2(b4.y, b4.y);
                    bv[4 * g + 2] = make_float2(b4.z, b4.z);
                    bv[4 * g + 3] = make_float2(b4.w, b4.w);
                } else {
                    const float2 b2 = *reinterpret_cast<const float2*>(wp);
                    bv[2 * g] = make_float2(b2.x, b2.x);
                    bv[2 * g + 1] = make_float2(b2.y, b2.y);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < NC; ++j) acc2[q][j] = __ffma2_rn(av[q], bv[j], acc2[q][j]);
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            acc[2 * q][j] = acc2[q][j].x;
            acc[2 * q + 1][j] = acc2[q][j].y;
        }
}

// bias + ReLU (+ JAC masking) and in-place write-back of one hidden layer's activations
template <int NC>
__device__ __forceinline__ void layer_epilogue(float (&acc)[8][NC], const float* __restrict__ bias, float* sAct,
                                               unsigned* sMask, const int* sKind, const int* sPair, int jac,
                                               int warp, int lane) {
    float bj[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) bj[j] = __ldg(bias + col_of<NC>(lane, j));
    if (jac) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int r = 8 * warp + q;
            if (sKind[r] == 0) {  // primal row: publish the ReLU pattern (warp-uniform branch)
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    unsigned word = __ballot_sync(0xffffffffu, acc[q][j] + bj[j] > 0.f);
                    if (lane == 0) sMask[sPair[r] * 8 + j] = word;
                }
            }
        }
        __syncthreads();
    } else {
        __syncwarp();
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int r = 8 * warp + q;
            const int kind = sKind[r];
            if (kind <= 0) {
                v[q] = fmaxf(acc[q][j] + bj[j], 0.f);  // primal (or padding row: harmless)
            } else {
                unsigned word = sMask[sPair[r] * 8 + j];
                v[q] = ((word >> lane) & 1u) ? acc[q][j] : 0.f;
            }
        }
        float* dst = sAct + (size_t)col_of<NC>(lane, j) * ASTR + 8 * warp;
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncwarp();
}

template <int NC>
__device__ __forceinline__ void hidden_layer(const float* __restrict__ Wt, const float* __restrict__ bias, int K,
                                             float* sAct, float* sW, unsigned* sMask, const int* sKind,
                                             const int* sPair, int jac, int warp, int lane, int tid) {
    float acc[8][NC];
    layer_gemm<NC>(acc, Wt, K, sAct, sW, warp, lane, tid);
    layer_epilogue<NC>(acc, bias, sAct, sMask, sKind, sPair, jac, warp, lane);
}

// JAC per-pair: Pm = (alpha/sigma^2) inv(I - sigma Jeps), v = Pm s   (tall_posterior_sampler.py:91-94).
// One thread per (pair, column c): it solves (I - sigma Jeps) x = e_c by LU with partial pivoting and hands back column c
// of Pm and its contribution Pm[:, c] s[c] to v.  (One thread per PAIR inverting the whole matrix left all but <= 10
// threads of the CTA idle and, for theta_dim = 10, ran out of registers into local memory.)
// fp64: the covariance I - sigma*Jeps of a learned (asymmetric, possibly indefinite) Jacobian is the ill-conditioned
// step of the JAC path; m^3 flops per pair are free next to the MLP.
template <int M>
__device__ __noinline__ void jac_pair_col(const float* sJp, const float* sSp, float sigma, float a_over_s2, int c, float* outp,
                                          float* prec_out, double* contrib) {
    double a[M][M], b[M];
#pragma unroll
    for (int o = 0; o < M; ++o) {
#pragma unroll
        for (int k = 0; k < M; ++k) a[o][k] = ((o == k) ? 1.0 : 0.0) - (double)sigma * (double)sJp[o * M + k];
        b[o] = (o == c) ? 1.0 : 0.0;
    }
    lu_solve_d<M>(a, b);
    const double sc = (double)sSp[c];
#pragma unroll
    for (int r = 0; r < M; ++r) {
        const double pr = (double)a_over_s2 * b[r];
        outp[M + r * M + c] = (float)pr;
        if (prec_out) prec_out[r * M + c] = (float)pr;
        contrib[r] = pr * sc;
    }
}

__device__ __forceinline__ void jac_pair_col_dispatch(int m, const float* J, const float* S, float sigma, float aos2, int c,
                                                      float* outp, float* po, double* contrib) {
    switch (m) {
        case 1: jac_pair_col<1>(J, S, sigma, aos2, c, outp, po, contrib); break;
        case 2: jac_pair_col<2>(J, S, sigma, aos2, c, outp, po, contrib); break;
        case 3: jac_pair_col<3>(J, S, sigma, aos2, c, outp, po, contrib); break;
        case 4: jac_pair_col<4>(J, S, sigma, aos2, c, outp, po, contrib); break;
        case 5: jac_pair_col<5>(J, S, sigma, aos2, c, outp, po, contrib); break;
        case 6: jac_pair_col<6>(J, S, sigma, aos2, c, outp, po, contrib); break;
        case 7: jac_pair_col<7>(J, S, sigma, aos2, c, outp, po, contrib); break;
        case 8: jac_pair_col<8>(J, S, sigma, aos2, c, outp, po, contrib); break;
        case 9: jac_pair_col<9>(J, S, sigma, aos2, c, outp, po, contrib); break;
        default: jac_pair_col<10>(J, S, sigma, aos2, c, outp, po, contrib); break;
    }
}

__global__ void __launch_bounds__(NT, 2) score_simt_kernel(const ScoreParams p) {
    extern __shared__ __align__(16) float smem[];
    float* sAct = smem;                           // [HMAX][ASTR]
    float* sW = sAct + HMAX * ASTR;               // [NSTAGE][KC][HMAX]
    float* sS = sW + NSTAGE * KC * HMAX;          // [TM][MMAX] scores of primal rows
    float* sQS = sS + TM * MMAX;                  // [TM][MMAX] GAUSS: inv(Sigma_j) s ; JAC: per-pair outputs
    float* sJ = sQS + TM * MMAX;                  // [P][m][m] d eps / d theta
    float* sTheta = sJ + TM * MMAX;               // [SG][MMAX]
    float* sDact = sTheta + TM * MMAX;            // [TM][MMAX] d eps / d (output pre-activation) of primal rows
    unsigned* sMask = reinterpret_cast<unsigned*>(sDact + TM * MMAX);  // [32][8]
    int* sKind = reinterpret_cast<int*>(sMask + 32 * 8);                // [TM] -1 invalid, 0 primal, 1+k tangent k
    int* sPair = sKind + TM;                                            // [TM] pair index within tile
    int* sSamp = sPair + TM;                                            // [TM] sample index within tile (per pair)
    int* sObs = sSamp + TM;                                             // [TM] global observation index (per pair)

    const NetDev& net = p.net;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m = net.m;
    const int tile = blockIdx.x;
    const int grp = tile / p.C, chunk = tile - grp * p.C;
    const int i0 = grp * p.SG;
    const int j0 = chunk * p.OC;
    const int nsamp = min(p.SG, p.N - i0);
    const int nobs = p.paired ? 1 : min(p.OC, p.n - j0);
    const float inv_sigma = p.sched[D4S_S_INV_SIGMA];
    const float sigma = p.sched[D4S_S_SIGMA];
    const float a_over_s2 = p.sched[D4S_S_A_OVER_S2];

    // ---- tile metadata -------------------------------------------------------------------------
    if (tid < TM) {
        const int r = tid;
        int kind = -1, pair = 0;
        if (r < p.P) { kind = 0; pair = r; }
        else if (p.jac && r < p.P * (1 + m)) { kind = 1 + (r - p.P) / p.P; pair = r % p.P; }
        const int si = pair / p.OC, oj = pair - si * p.OC;
        if (kind >= 0 && (si >= nsamp || oj >= nobs)) kind = -1;
        sKind[r] = kind;
        sPair[r] = pair;
        if (r < p.P) {
            sSamp[r] = si;
            // batched contexts: the sample's context owns observations [c n, (c + 1) n)
            const int ctx0 = p.ctx_samples > 0 ? (min(i0 + si, p.N - 1) / p.ctx_samples) * p.n : 0;
            sObs[r] = p.paired ? min((i0 + si) / p.paired, p.n - 1) : ctx0 + min(j0 + oj, p.n - 1);  // paired = samples per observation row
        }
    }
    for (int e = tid; e < p.SG * m; e += NT) {
        const int si = e / m, k = e - si * m;
        sTheta[si * MMAX + k] = (si < nsamp) ? p.theta[(size_t)(i0 + si) * m + k] : 0.f;
    }
    __syncthreads();

    // ---- layer 0 (separable) -------------------------------------------------------------------
    {
        const int H0 = net.Hp[0];
        const bool pre = (p.base_pre != nullptr);  // theta goes through an embedding MLP (separate kernel)
        for (int j = 0; j < H0 / 32; ++j) {
            const int col = lane + 32 * j;
            const float tf = __ldg(p.time_feat + col);
            float acol[MMAX];
#pragma unroll
            for (int k = 0; k < MMAX; ++k) acol[k] = (!pre && k < m) ? __ldg(net.A + (size_t)k * H0 + col) : 0.f;
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int r = 8 * warp + q;
                const int kind = sKind[r];
                const int pr = sPair[r];
                const int si = sSamp[pr];
                const size_t gi = (size_t)min(i0 + si, p.N - 1);
                float z = tf + __ldg(p.obs_feat + (size_t)sObs[pr] * H0 + col);
                if (pre) {
                    z += __ldg(p.base_pre + gi * H0 + col);
                } else {
#pragma unroll
                    for (int k = 0; k < MMAX; ++k)
                        if (k < m) z = fmaf(acol[k], sTheta[si * MMAX + k], z);
                }
                float out = 0.f;
                if (kind == 0) out = fmaxf(z, 0.f);
                else if (kind > 0) {
                    float t = 0.f;
                    if (pre) {
                        t = __ldg(p.tan_pre + (gi * m + (kind - 1)) * H0 + col);
                    } else {
#pragma unroll
                        for (int k = 0; k < MMAX; ++k) t = (k == kind - 1) ? acol[k] : t;
                    }
                    out = (z > 0.f) ? t : 0.f;
                }
                v[q] = out;
            }
            float* dst = sAct + (size_t)col * ASTR + 8 * warp;
            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
    __syncwarp();

    // ---- hidden layers 1 .. L-2 ------------------------------------------------------------------
    for (int l = 1; l <= net.L - 2; ++l) {
        const int K = net.Hp[l - 1];
        switch (net.Hp[l]) {
            case 64: hidden_layer<2>(net.Wt[l], net.bias[l], K, sAct, sW, sMask, sKind, sPair, p.jac, warp, lane, tid); break;
            case 128: hidden_layer<4>(net.Wt[l], net.bias[l], K, sAct, sW, sMask, sKind, sPair, p.jac, warp, lane, tid); break;
            default: hidden_layer<8>(net.Wt[l], net.bias[l], K, sAct, sW, sMask, sKind, sPair, p.jac, warp, lane, tid); break;
        }
    }

    // ---- output layer: eps[r][o] = bout[o] + sum_k Wout[k][o] act[k][r] ------------------------------
    {
        const int K = net.Hp[net.L - 2];
        for (int e = lane; e < 8 * m; e += 32) {
            const int q = e & 7, o = e >> 3;
            const int r = 8 * warp + q;
            const float* a = sAct + 8 * warp + q;
            const float* w = net.Wout + o;
            float d0 = 0.f, d1 = 0.f;
            for (int k = 0; k < K; k += 2) {
                d0 = fmaf(a[(size_t)k * ASTR], __ldg(w + (size_t)k * MMAX), d0);
                d1 = fmaf(a[(size_t)(k + 1) * ASTR], __ldg(w + (size_t)(k + 1) * MMAX), d1);
            }
            const float dot = d0 + d1;
            const int kind = sKind[r];
            if (kind == 0) {
                const int pr = sPair[r];
                float z = dot + __ldg(net.bout + o);
                float dact = net.out_scale;
                if (net.out_act == 1) {  // tanh head of the toy perturbation net (embedding_nets.py:135)
                    z = tanhf(z);
                    dact *= 1.f - z * z;
                }
                float eps = net.out_scale * z;
                if (p.affine) {  // analytic Gaussian part: eps += M theta_i - G x_j - g  (gen_gaussian_gaussian.py:59-72)
                    const float* M = p.affine + o * m;
                    const float* G = p.affine + m * m + o * p.d;
                    const float* xr = p.obs_raw + (size_t)sObs[pr] * p.d;
                    float aff = -__ldg(p.affine + m * m + m * p.d + o);
                    for (int k = 0; k < m; ++k) aff = fmaf(__ldg(M + k), sTheta[sSamp[pr] * MMAX + k], aff);
                    for (int k = 0; k < p.d; ++k) aff = fmaf(-__ldg(G + k), __ldg(xr + k), aff);
                    eps += aff;
                }
                const float sc = -eps * inv_sigma;  // nse.py:145
                sS[r * MMAX + o] = sc;
                sDact[r * MMAX + o] = dact;
                if (p.scores_out) {
                    const size_t gi = (size_t)(i0 + sSamp[pr]);
                    const size_t gj = p.paired ? 0 : (size_t)(p.ctx_samples > 0 ? sObs[pr] % p.n : sObs[pr]);
                    const size_t nn = p.paired ? 1 : (size_t)p.n;
                    p.scores_out[(gi * nn + gj) * m + o] = sc;
                }
            } else if (kind > 0) {
                sJ[(sPair[r] * m + o) * m + (kind - 1)] = dot;  // d (output pre-activation)_o / d theta_k
            }
        }
    }
    __syncthreads();
    if (p.jac && (net.out_act != 0 || net.out_scale != 1.f || p.affine)) {
        // chain rule through the output head: d eps_o / d theta_k = dact_o * dz_o/dtheta_k + M[o][k]
        for (int e = tid; e < p.P * m * m; e += NT) {
            const int pr = e / (m * m), ok = e - pr * m * m, o = ok / m;
            float v = sDact[pr * MMAX + o] * sJ[e];
            if (p.affine) v += __ldg(p.affine + ok);
            sJ[e] = v;
        }
    }
    __syncthreads();

    // ---- per-pair Tweedie terms -----------------------------------------------------------------------
    if (!p.jac) {
        // GAUSS: inv(Sigma_j) s_ij ; the alpha/sigma^2 I part of the precision is applied after the
        // reduction (it is the same for every observation).
        for (int e = tid; e < p.P * m; e += NT) {
            const int r = e / m, o = e - r * m;
            float acc = 0.f;
            if (sKind[r] == 0 && p.obs_prec) {
                const float* Q = p.obs_prec + ((size_t)sObs[r] * m + o) * m;
                for (int k = 0; k < m; ++k) acc = fmaf(__ldg(Q + k), sS[r * MMAX + k], acc);
            }
            sQS[r * MMAX + o] = acc;
        }
    } else {
        double* sC = reinterpret_cast<double*>(sAct);  // [P][m][m]: column c's contribution Pm[:, c] s[c] (activations are dead)
        for (int e = tid; e < p.P * m; e += NT) {
            const int r = e / m, c = e - r * m;
            if (sKind[r] == 0) {
                float* po = nullptr;
                if (p.prec_out) {
                    const size_t gi = (size_t)(i0 + sSamp[r]);
                    const size_t gj = p.paired ? 0 : (size_t)(p.ctx_samples > 0 ? sObs[r] % p.n : sObs[r]);
                    const size_t nn = p.paired ? 1 : (size_t)p.n;
                    po = p.prec_out + (gi * nn + gj) * m * m;
                }
                jac_pair_col_dispatch(m, sJ + r * m * m, sS + r * MMAX, sigma, a_over_s2, c, sQS + r * (m + m * m), po,
                                      sC + (size_t)(r * m + c) * m);
            }
        }
        __syncthreads();
        for (int e = tid; e < p.P * m; e += NT) {  // v = Pm s: the columns' contributions in a fixed order
            const int r = e / m, o = e - r * m;
            if (sKind[r] == 0) {
                double v = 0.0;
                for (int c = 0; c < m; ++c) v += sC[(size_t)(r * m + c) * m + o];
                sQS[r * (m + m * m) + o] = (float)v;
            }
        }
    }
    __syncthreads();

    // ---- reduction over the tile's observations (fixed order => deterministic) ------------------------
    const int W = p.W;
    for (int e = tid; e < nsamp * W; e += NT) {
        const int si = e / W, c = e - si * W;
        float acc = 0.f;
        for (int oj = 0; oj < nobs; ++oj) {
            const int r = si * p.OC + oj;
            float v;
            if (!p.jac) v = (c < m) ? sS[r * MMAX + c] : sQS[r * MMAX + (c - m)];
            else v = sQS[r * W + c];
            if (p.obs_weight) v *= __ldg(p.obs_weight + (p.ctx_samples > 0 ? sObs[r] % p.n : sObs[r]));
            acc += v;
        }
        if (p.part) p.part[((size_t)(i0 + si) * p.C + chunk) * W + c] = acc;
    }
}

// ---- kernel 0: theta-embedding MLP (JR-NMM nets, nse.py:129) + projection on the layer-0 theta block -------------------
// One CTA per posterior sample.  Rows: the primal e(theta_i) and, for JAC, the m tangents d e / d theta_k (forward
// mode, masked by the primal's ReLU pattern).  Output: base_pre[i][col] = sum_f A[f][col] e_f,
// tan_pre[i][k][col] = sum_f A[f][col] d e_f / d theta_k.
__global__ void __launch_bounds__(256) theta_embed_kernel(const NetDev net, const float* __restrict__ theta, int N, int jac,
                                                         float* __restrict__ base_pre, float* __restrict__ tan_pre) {
    __shared__ float act[2][1 + MMAX][HMAX];  // ping-pong activations: row 0 primal, rows 1..m tangents
    const int i = blockIdx.x, tid = threadIdx.x;
    const int m = net.m, rows = jac ? 1 + m : 1;
    if (i >= N) return;
    for (int e = tid; e < (1 + MMAX) * HMAX; e += blockDim.x) {
        const int r = e / HMAX, c = e - r * HMAX;
        float v = 0.f;
        if (c < m) v = (r == 0) ? theta[(size_t)i * m + c] : ((c == r - 1) ? 1.f : 0.f);  // tangent seeds = unit vectors
        act[0][r][c] = v;
    }
    __syncthreads();
    int cur = 0;
    for (int l = 0; l < net.emb_L; ++l) {
        const int K = net.emb_dims[l], Nn = net.emb_dims[l + 1];
        const bool relu = (l < net.emb_L - 1);
        for (int o = tid; o < Nn; o += blockDim.x) {
            float acc[1 + MMAX];
#pragma unroll
            for (int r = 0; r < 1 + MMAX; ++r) acc[r] = 0.f;
            for (int k = 0; k < K; ++k) {
                const float w = __ldg(net.emb_Wt[l] + (size_t)k * Nn + o);
#pragma unroll
                for (int r = 0; r < 1 + MMAX; ++r)
                    if (r < rows) acc[r] = fmaf(w, act[cur][r][k], acc[r]);
            }
            const float z = acc[0] + __ldg(net.emb_bias[l] + o);
            const bool on = !relu || z > 0.f;
            act[cur ^ 1][0][o] = relu ? fmaxf(z, 0.f) : z;
#pragma unroll
            for (int r = 1; r < 1 + MMAX; ++r)
                if (r < rows) act[cur ^ 1][r][o] = on ? acc[r] : 0.f;
        }
        __syncthreads();
        cur ^= 1;
    }
    const int F = net.theta_cols, H0 = net.Hp[0];
    for (int col = tid; col < H0; col += blockDim.x) {
        float acc[1 + MMAX];
#pragma unroll
        for (int r = 0; r < 1 + MMAX; ++r) acc[r] = 0.f;
        for (int f = 0; f < F; ++f) {
            const float a = __ldg(net.A + (size_t)f * H0 + col);
#pragma unroll
            for (int r = 0; r < 1 + MMAX; ++r)
                if (r < rows) acc[r] = fmaf(a, act[cur][r][f], acc[r]);
        }
        base_pre[(size_t)i * H0 + col] = acc[0];
        if (jac) {
#pragma unroll
            for (int r = 1; r < 1 + MMAX; ++r)
                if (r < rows) tan_pre[((size_t)i * m + (r - 1)) * H0 + col] = acc[r];
        }
    }
}

cudaError_t launch_theta_embed(const NetDev& net, const float* theta, int N, int jac, float* base_pre, float* tan_pre,
                               cudaStream_t stream) {
    theta_embed_kernel<<<N, 256, 0, stream>>>(net, theta, N, jac, base_pre, tan_pre);
    return cudaGetLastError();
}

size_t score_simt_smem_bytes() {
    return sizeof(float) * (HMAX * ASTR + NSTAGE * KC * HMAX + 5 * TM * MMAX) + sizeof(unsigned) * 32 * 8 +
           sizeof(int) * 4 * TM;
}

cudaError_t launch_score_simt(const ScoreParams& p, int tiles, cudaStream_t stream) {
    static bool configured[64] = {};  // the attribute is per device
    const size_t smem = score_simt_smem_bytes();
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(score_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    score_simt_kernel<<<tiles, NT, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace d4s
