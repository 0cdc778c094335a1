// Fused tcgen05 trajectory kernel (GAUSS): one launch runs EVERY DDIM step of the trajectory.
//
// A CTA owns a group of SG posterior samples for the whole trajectory (samples never interact, so there is
// no inter-CTA dependency and no per-step launch).  Per step and per chunk of OC observations it evaluates the
// score MLP on the 128-row tile (sample x observation pairs):
//   layer 0   separable fp32 SIMT:  z0 = A theta_i + obs_feat_j + time_feat              (8 compute warps)
//   layer l   D[128 x N] (TMEM, fp32) = act[128 x K] (smem) x W_l^T (smem ring)  on tcgen05.mma kind::f16
//             with the bf16x3 split  a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo   (fp32-level parity, 3 MMAs)
//             weights: pre-packed canonical K-major bf16 images, streamed L2 -> smem by 1-D TMA bulk copies
//             epilogue: tcgen05.ld -> +bias, ReLU -> hi/lo bf16 -> next layer's A operand in smem
//   output    fused into the last epilogue in fp32, then score = -eps/sigma, inv(Sigma_j) s, sums over j
//   finalize  prior score, Lambda solve, DDIM update with Philox noise (shared code: d4s_finalize_core.cuh)
// theta stays in shared memory across all steps.
//
// Replaces the same reference lines as kernels 1-4 (nse.py:263-298, 544-656; tall_posterior_sampler.py:95-128).
//
// Warp roles: warps 0-7 compute/epilogue (TMEM lane quarter = warp % 4, column half = warp / 4),
//             warp 8 = TMA producer (one lane), warp 9 = TMEM allocator + MMA issuer (one lane).
#include <cuda_bf16.h>

#include "d4s_finalize_core.cuh"

namespace d4s {

constexpr int TC_M = 128;                      // tile rows = TMEM lanes
constexpr int TC_CW = 8;                       // compute warps
constexpr int TC_THREADS = (TC_CW + 2) * 32;   // 320
constexpr int TC_NSTG = 4;                     // weight ring stages
constexpr int TC_STAGE_BYTES = 256 * 64;       // one K=16 step of W_hi | W_lo for N <= 256
constexpr int TC_A_BYTES = (256 / 16) * 4096;  // 64 KB: act[128 x 256] bf16 in K16-step images
constexpr int TC_SG_MAX = 8;                   // samples per tile (layer-0 staging buffer)
constexpr long long SPIN_LIMIT_CYCLES = 4000000000LL;  // ~2 s: bounded mbarrier waits trap instead of hanging

// ---- shared memory carve-up (bytes) ---------------------------------------------------------------------------
constexpr int OFF_AHI = 0;
constexpr int OFF_ALO = OFF_AHI + TC_A_BYTES;
constexpr int OFF_RING = OFF_ALO + TC_A_BYTES;
constexpr int OFF_S = OFF_RING + TC_NSTG * TC_STAGE_BYTES;   // float [128][MMAX] scores
constexpr int OFF_QS = OFF_S + TC_M * MMAX * 4;               // float [128][MMAX] inv(Sigma_j) s
constexpr int OFF_HALF = OFF_QS + TC_M * MMAX * 4;            // float [128][MMAX] output-layer partials (upper half)
constexpr int OFF_BASE = OFF_HALF + TC_M * MMAX * 4;          // float [SG_MAX][256] layer-0 per-sample part
constexpr int OFF_ACC = OFF_BASE + TC_SG_MAX * 256 * 4;       // float [SG_MAX][2*MMAX]
constexpr int OFF_THETA = OFF_ACC + TC_SG_MAX * 2 * MMAX * 4; // float [SG_MAX][MMAX]
constexpr int OFF_WOUT = OFF_THETA + TC_SG_MAX * MMAX * 4;    // float [MMAX][256] output layer (torch layout)
constexpr int OFF_NOISE = OFF_WOUT + MMAX * 256 * 4;          // float [SG_MAX][MMAX] DDIM noise of the current step
constexpr int OFF_PRIOR = OFF_NOISE + TC_SG_MAX * MMAX * 4;   // float [SG_MAX][2][MMAX] uniform-prior score / derivative
constexpr int OFF_BAR = OFF_PRIOR + TC_SG_MAX * 2 * MMAX * 4; // mbarriers
constexpr int OFF_TMEM = OFF_BAR + (2 * TC_NSTG + 2) * 8;
constexpr int TC_SMEM_BYTES = OFF_TMEM + 16;

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > SPIN_LIMIT_CYCLES) {  // a protocol bug must not hang the GPU: fail loudly instead
            if (err) atomicExch(err, 1);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void compute_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, no swizzle ("interleave"): 8x16B core matrices, 128 B each.
//   lbo = byte distance between the two K halves of a K=16 step, sbo = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor: D fp32, A/B bf16, both K-major, dense, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}

// splits 8 fp32 activations into bf16 hi / lo and stores the two 16-byte K-groups
__device__ __forceinline__ void store_split8(const float (&v)[8], uint8_t* a_hi, uint8_t* a_lo, int byte_off) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
        const float2 hf = __bfloat1622float2(h);
        const __nv_bfloat162 l = __floats2bfloat162_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
        hi[e] = *reinterpret_cast<const uint32_t*>(&h);
        lo[e] = *reinterpret_cast<const uint32_t*>(&l);
    }
    *reinterpret_cast<uint4*>(a_hi + byte_off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(a_lo + byte_off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// byte offset of (row r, k-group kg of 8 columns) inside an A image: K16 step = kg/2, half = kg%2
__device__ __forceinline__ int a_off(int r, int kg) { return (kg >> 1) * 4096 + (kg & 1) * 2048 + r * 16; }

template <int M>
__device__ __noinline__ void finalize_one(const FinalizeParams& fp, const float* part, float* theta, const float* noise,
                                          long long gs, const float* pre) {
    double loc[2 * M];
    if (pre) {
#pragma unroll
        for (int k = 0; k < M; ++k) { loc[k] = (double)pre[k]; loc[M + k] = (double)pre[MMAX + k]; }
    }
    finalize_sample<M>(fp, part, theta, nullptr, noise, gs, pre ? loc : nullptr);
}

__device__ __forceinline__ void finalize_dispatch(int m, const FinalizeParams& fp, const float* part, float* theta,
                                                  const float* noise, long long gs, const float* pre) {
    switch (m) {
        case 1: finalize_one<1>(fp, part, theta, noise, gs, pre); break;
        case 2: finalize_one<2>(fp, part, theta, noise, gs, pre); break;
        case 3: finalize_one<3>(fp, part, theta, noise, gs, pre); break;
        case 4: finalize_one<4>(fp, part, theta, noise, gs, pre); break;
        case 5: finalize_one<5>(fp, part, theta, noise, gs, pre); break;
        case 6: finalize_one<6>(fp, part, theta, noise, gs, pre); break;
        case 7: finalize_one<7>(fp, part, theta, noise, gs, pre); break;
        case 8: finalize_one<8>(fp, part, theta, noise, gs, pre); break;
        case 9: finalize_one<9>(fp, part, theta, noise, gs, pre); break;
        default: finalize_one<10>(fp, part, theta, noise, gs, pre); break;
    }
}

__global__ void __launch_bounds__(TC_THREADS, 1) traj_tc_kernel(const TrajParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sAhi = smem + OFF_AHI;
    uint8_t* sAlo = smem + OFF_ALO;
    float* sS = reinterpret_cast<float*>(smem + OFF_S);
    float* sQS = reinterpret_cast<float*>(smem + OFF_QS);
    float* sHalf = reinterpret_cast<float*>(smem + OFF_HALF);
    float* sBase = reinterpret_cast<float*>(smem + OFF_BASE);
    float* sAcc = reinterpret_cast<float*>(smem + OFF_ACC);
    float* sTheta = reinterpret_cast<float*>(smem + OFF_THETA);
    float* sWout = reinterpret_cast<float*>(smem + OFF_WOUT);
    float* sNoise = reinterpret_cast<float*>(smem + OFF_NOISE);
    float* sPrior = reinterpret_cast<float*>(smem + OFF_PRIOR);
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
    const uint32_t bar0 = smem_u32(smem + OFF_BAR);
    auto bar_full = [&](int s) { return bar0 + 8u * s; };
    auto bar_empty = [&](int s) { return bar0 + 8u * (TC_NSTG + s); };
    const uint32_t bar_a = bar0 + 8u * (2 * TC_NSTG);      // A operand (and TMEM D) ready for the MMA warp
    const uint32_t bar_d = bar0 + 8u * (2 * TC_NSTG + 1);  // accumulator ready for the epilogue

    const NetDev& net = p.net;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m = net.m, L = net.L;
    const int H0 = net.Hp[0];
    const int n_groups = (p.N + p.SG - 1) / p.SG;

    if (tid == 0) {
        for (int s = 0; s < TC_NSTG; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_empty(s), 1);
        }
        mbar_init(bar_a, TC_CW * 32);
        mbar_init(bar_d, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_CW + 1) {  // TMEM: 256 fp32 accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *sTmem;

    const int n_steps = p.step_end - p.step_begin;
    int my_groups = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) ++my_groups;
    const long long tile_iters = (long long)my_groups * n_steps * p.C;  // tile evaluations by this CTA

    if (warp == TC_CW) {
        // ================= TMA producer: streams the weight images in consumption order ===========================
        if (lane == 0) {
            unsigned it = 0;
            for (long long ti = 0; ti < tile_iters; ++ti) {
                for (int l = 1; l <= L - 2; ++l) {
                    const int K = net.Hp[l - 1], Nn = net.Hp[l];
                    const uint32_t bytes = (uint32_t)Nn * 64u;
                    const uint8_t* src = net.wimg[l];
                    for (int ks = 0; ks < K / 16; ++ks, ++it) {
                        const int s = it % TC_NSTG;
                        const uint32_t ph = (it / TC_NSTG) & 1u;
                        mbar_wait(bar_empty(s), ph ^ 1u, p.error_flag);
                        mbar_expect_tx(bar_full(s), bytes);
                        tma_bulk_g2s(smem_u32(smem + OFF_RING + s * TC_STAGE_BYTES), src + (size_t)ks * bytes, bytes,
                                     bar_full(s));
                    }
                }
            }
        }
    } else if (warp == TC_CW + 1) {
        // ================= MMA issuer ===================================================================================
        if (lane == 0) {
            unsigned it = 0, a_cnt = 0;
            const uint32_t a_hi0 = smem_u32(sAhi), a_lo0 = smem_u32(sAlo);
            for (long long ti = 0; ti < tile_iters; ++ti) {
                for (int l = 1; l <= L - 2; ++l) {
                    const int K = net.Hp[l - 1], Nn = net.Hp[l];
                    const uint32_t idesc = umma_idesc(Nn);
                    const uint32_t b_lbo = p.swap_lbo_sbo ? 128u : (uint32_t)Nn * 16u;
                    const uint32_t b_sbo = p.swap_lbo_sbo ? (uint32_t)Nn * 16u : 128u;
                    const uint32_t a_lbo = p.swap_lbo_sbo ? 128u : 2048u;
                    const uint32_t a_sbo = p.swap_lbo_sbo ? 2048u : 128u;
                    mbar_wait(bar_a, a_cnt & 1u, p.error_flag);
                    ++a_cnt;
                    tc_fence_after();
                    for (int ks = 0; ks < K / 16; ++ks, ++it) {
                        const int s = it % TC_NSTG;
                        const uint32_t ph = (it / TC_NSTG) & 1u;
                        mbar_wait(bar_full(s), ph, p.error_flag);
                        tc_fence_after();
                        const uint32_t wb = smem_u32(smem + OFF_RING + s * TC_STAGE_BYTES);
                        const uint64_t b_hi = umma_desc(wb, b_lbo, b_sbo);
                        const uint64_t b_lo = umma_desc(wb + (uint32_t)Nn * 32u, b_lbo, b_sbo);
                        const uint64_t a_hi = umma_desc(a_hi0 + ks * 4096, a_lbo, a_sbo);
                        const uint64_t a_lo = umma_desc(a_lo0 + ks * 4096, a_lbo, a_sbo);
                        tc_mma_bf16(tmem_d, a_hi, b_hi, idesc, ks > 0 ? 1u : 0u);
                        tc_mma_bf16(tmem_d, a_lo, b_hi, idesc, 1u);
                        tc_mma_bf16(tmem_d, a_hi, b_lo, idesc, 1u);
                        tc_commit(bar_empty(s));  // ring slot is free once these MMAs have read it
                    }
                    tc_commit(bar_d);  // accumulator complete
                }
            }
        }
    } else {
        // ================= compute warps ==================================================================================
        const int q = warp & 3, ch = warp >> 2;
        const int row = 32 * q + lane;  // the TMEM lane / tile row this thread owns in the epilogues
        unsigned d_cnt = 0;
        const bool uni = (p.fin.prior_kind == D4S_PRIOR_UNIFORM);
        const int Kout = net.Hp[L - 2];
        // rows of the tile that hold no (sample, observation) pair stay zero for the whole kernel
        for (int e = tid; e < 2 * TC_A_BYTES / 16; e += 256) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0, 0, 0, 0);
        for (int e = tid; e < m * Kout; e += 256) sWout[(e / Kout) * 256 + (e % Kout)] = __ldg(net.WoutT + e);
        for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
            const int i0 = g * p.SG;
            const int nsamp = min(p.SG, p.N - i0);
            for (int e = tid; e < TC_SG_MAX * MMAX; e += 256) {
                const int si = e / MMAX, k = e - si * MMAX;
                sTheta[e] = (si < nsamp && k < m) ? p.theta[(size_t)(i0 + si) * m + k] : 0.f;
            }
            compute_bar();
            for (int step = p.step_begin; step < p.step_end; ++step) {
                const float* sched = p.sched + (size_t)step * D4S_SCHED_STRIDE;
                const float inv_sigma = __ldg(sched + D4S_S_INV_SIGMA);
                // ---- layer-0 per-sample part: base[si][col] = time_feat[col] + sum_k A[k][col] theta[si][k] ----------
                {
                    const float* tf = p.time_feat + (size_t)step * H0;
                    for (int col = tid; col < H0; col += 256) {
                        float acol[MMAX];
#pragma unroll
                        for (int k = 0; k < MMAX; ++k) acol[k] = (k < m) ? __ldg(net.A + (size_t)k * H0 + col) : 0.f;
                        const float t0 = __ldg(tf + col);
                        for (int si = 0; si < p.SG; ++si) {
                            float z = t0;
#pragma unroll
                            for (int k = 0; k < MMAX; ++k)
                                if (k < m) z = fmaf(acol[k], sTheta[si * MMAX + k], z);
                            sBase[si * 256 + col] = z;
                        }
                    }
                }
                compute_bar();
                for (int chunk = 0; chunk < p.C; ++chunk) {
                    const int j0 = chunk * p.OC;
                    const int nobs = min(p.OC, p.n - j0);
                    // ---- layer 0: act0[r][col] = relu(base[si][col] + obs_feat[j][col]) -> bf16 hi/lo A images -----------
                    // item = (observation oj, k-group kg): one 32-byte read of obs_feat serves the rows of all SG samples;
                    // all global loads of a thread are issued before their first use.
                    {
                        const int nkg = H0 / 8;
                        const int items = p.OC * nkg;
                        constexpr int IPT = 4;  // items per thread per pass
                        for (int it0 = 0; it0 < items; it0 += 256 * IPT) {
                            float4 u[IPT][2];
#pragma unroll
                            for (int j = 0; j < IPT; ++j) {
                                const int item = it0 + j * 256 + tid;
                                const int kg = item / p.OC, oj = item - kg * p.OC;
                                if (item < items && oj < nobs) {
                                    const float* up = p.obs_feat + (size_t)(j0 + oj) * H0 + 8 * kg;
                                    u[j][0] = __ldg(reinterpret_cast<const float4*>(up));
                                    u[j][1] = __ldg(reinterpret_cast<const float4*>(up + 4));
                                } else {
                                    u[j][0] = u[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                                }
                            }
#pragma unroll
                            for (int j = 0; j < IPT; ++j) {
                                const int item = it0 + j * 256 + tid;
                                if (item >= items) break;
                                const int kg = item / p.OC, oj = item - kg * p.OC;
                                const bool ovalid = oj < nobs;
                                for (int si = 0; si < p.SG; ++si) {
                                    float v[8];
                                    if (ovalid && si < nsamp) {
                                        const float* brow = sBase + si * 256 + 8 * kg;
                                        const float4 b0 = *reinterpret_cast<const float4*>(brow);
                                        const float4 b1 = *reinterpret_cast<const float4*>(brow + 4);
                                        v[0] = fmaxf(u[j][0].x + b0.x, 0.f); v[1] = fmaxf(u[j][0].y + b0.y, 0.f);
                                        v[2] = fmaxf(u[j][0].z + b0.z, 0.f); v[3] = fmaxf(u[j][0].w + b0.w, 0.f);
                                        v[4] = fmaxf(u[j][1].x + b1.x, 0.f); v[5] = fmaxf(u[j][1].y + b1.y, 0.f);
                                        v[6] = fmaxf(u[j][1].z + b1.z, 0.f); v[7] = fmaxf(u[j][1].w + b1.w, 0.f);
                                    } else {
#pragma unroll
                                        for (int e = 0; e < 8; ++e) v[e] = 0.f;
                                    }
                                    store_split8(v, sAhi, sAlo, a_off(si * p.OC + oj, kg));
                                }
                            }
                        }
                        // rows beyond the SG*OC pairs of the tile are never written: zero them once per kernel (below)
                    }
                    fence_proxy_async();
                    mbar_arrive(bar_a);

                    // ---- theta-only work of this step, hidden behind the first tensor-core layer -----------------------------
                    if (chunk == 0) {
                        if (uni && tid < nsamp * m) {  // erf/exp of the diffused-uniform score: one thread per (sample, dim)
                            const int si = tid / m, k = tid - si * m;
                            double s0, jac;
                            uniform_prior_terms((double)sTheta[si * MMAX + k], (double)__ldg(p.fin.low + k),
                                                (double)__ldg(p.fin.high + k), (double)__ldg(sched + D4S_S_SQRT_ALPHA),
                                                (double)__ldg(sched + D4S_S_SIGMA), s0, jac);
                            sPrior[si * 2 * MMAX + k] = (float)s0;
                            sPrior[si * 2 * MMAX + MMAX + k] = (float)jac;
                        }
                        const int nq = (m + 3) / 4;
                        if (tid >= 128 && tid < 128 + nsamp * nq) {  // DDIM noise: one thread per (sample, 4 dims)
                            const int e = tid - 128, si = e / nq, qd = e - si * nq;
                            float z[4] = {0.f, 0.f, 0.f, 0.f};
                            if (p.noise) {
                                const float* nz = p.noise + ((size_t)step * p.N + (i0 + si)) * m;
                                for (int k = 0; k < 4; ++k)
                                    if (4 * qd + k < m) z[k] = __ldg(nz + 4 * qd + k);
                            } else {
                                philox_normal4(p.fin.seed, step, p.fin.sample_offset + i0 + si, qd, z);
                            }
                            for (int k = 0; k < 4; ++k)
                                if (4 * qd + k < MMAX) sNoise[si * MMAX + 4 * qd + k] = z[k];
                        }
                    }

                    // ---- hidden layers on the tensor core; epilogues here ----------------------------------------------------
                    float po[MMAX];  // output-layer partial dot products of this thread's row / column half
#pragma unroll
                    for (int o = 0; o < MMAX; ++o) po[o] = 0.f;
                    for (int l = 1; l <= L - 2; ++l) {
                        const int Nn = net.Hp[l];
                        const bool last = (l == L - 2);
                        const int half_cols = Nn >> 1;
                        const int nblk = half_cols / 32;
                        const float* bias_l = net.bias[l] + ch * half_cols;
                        float4 bcur[8], bnxt[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) bcur[e] = __ldg(reinterpret_cast<const float4*>(bias_l) + e);  // hidden behind the MMA
                        mbar_wait(bar_d, d_cnt & 1u, p.error_flag);
                        ++d_cnt;
                        tc_fence_after();
                        for (int cb = 0; cb < nblk; ++cb) {
                            const int col0 = ch * half_cols + 32 * cb;
                            uint32_t d[32];
                            tc_ld32(tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)col0, d);
                            if (cb + 1 < nblk) {
#pragma unroll
                                for (int e = 0; e < 8; ++e)
                                    bnxt[e] = __ldg(reinterpret_cast<const float4*>(bias_l + 32 * (cb + 1)) + e);
                            }
                            tc_wait_ld();
                            float a[32];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                a[4 * e] = fmaxf(__uint_as_float(d[4 * e]) + bcur[e].x, 0.f);
                                a[4 * e + 1] = fmaxf(__uint_as_float(d[4 * e + 1]) + bcur[e].y, 0.f);
                                a[4 * e + 2] = fmaxf(__uint_as_float(d[4 * e + 2]) + bcur[e].z, 0.f);
                                a[4 * e + 3] = fmaxf(__uint_as_float(d[4 * e + 3]) + bcur[e].w, 0.f);
                            }
                            if (!last) {
#pragma unroll
                                for (int k8 = 0; k8 < 4; ++k8) {
                                    float v[8];
#pragma unroll
                                    for (int e = 0; e < 8; ++e) v[e] = a[8 * k8 + e];
                                    store_split8(v, sAhi, sAlo, a_off(row, (col0 >> 3) + k8));
                                }
                            } else {
                                // output layer in fp32: po[o] += sum_c act[c] * Wout[o][c]   (Wout in smem, broadcast reads)
#pragma unroll
                                for (int o = 0; o < MMAX; ++o) {
                                    if (o < m) {
                                        const float* w = sWout + o * 256 + col0;
                                        float acc = po[o];
#pragma unroll
                                        for (int e = 0; e < 32; e += 4) {
                                            const float4 w4 = *reinterpret_cast<const float4*>(w + e);
                                            acc = fmaf(a[e], w4.x, acc);
                                            acc = fmaf(a[e + 1], w4.y, acc);
                                            acc = fmaf(a[e + 2], w4.z, acc);
                                            acc = fmaf(a[e + 3], w4.w, acc);
                                        }
                                        po[o] = acc;
                                    }
                                }
                            }
#pragma unroll
                            for (int e = 0; e < 8; ++e) bcur[e] = bnxt[e];
                        }
                        if (!last) {
                            fence_proxy_async();
                            tc_fence_before();
                            mbar_arrive(bar_a);
                        }
                    }
                    // single-hidden-layer nets (L == 2) have no tensor-core layer: output layer straight from layer 0
                    // is not supported here (the host routes them to the SIMT path).

                    // ---- scores: combine the two column halves, s = -(eps + b)/sigma ----------------------------------------
                    if (ch == 1) {
#pragma unroll
                        for (int o = 0; o < MMAX; ++o)
                            if (o < m) sHalf[row * MMAX + o] = po[o];
                    }
                    tc_fence_before();
                    compute_bar();
                    if (ch == 0) {
#pragma unroll
                        for (int o = 0; o < MMAX; ++o)
                            if (o < m) sS[row * MMAX + o] = -(po[o] + sHalf[row * MMAX + o] + __ldg(net.bout + o)) * inv_sigma;
                    }
                    compute_bar();
                    // ---- inv(Sigma_j) s per pair -----------------------------------------------------------------------------
                    for (int e = tid; e < p.P * m; e += 256) {
                        const int r = e / m, o = e - r * m;
                        const int si = r / p.OC, oj = r - si * p.OC;
                        float acc = 0.f;
                        if (p.obs_prec && si < nsamp && oj < nobs) {
                            const float* Q = p.obs_prec + ((size_t)(j0 + oj) * m + o) * m;
                            for (int k = 0; k < m; ++k) acc = fmaf(__ldg(Q + k), sS[r * MMAX + k], acc);
                        }
                        sQS[r * MMAX + o] = acc;
                    }
                    compute_bar();
                    // ---- sums over the chunk's observations (fixed order) ------------------------------------------------------
                    for (int e = tid; e < nsamp * 2 * m; e += 256) {
                        const int si = e / (2 * m), c = e - si * 2 * m;
                        float acc = (chunk == 0) ? 0.f : sAcc[si * 2 * m + c];
                        for (int oj = 0; oj < nobs; ++oj) {
                            const int r = si * p.OC + oj;
                            acc += (c < m) ? sS[r * MMAX + c] : sQS[r * MMAX + (c - m)];
                        }
                        sAcc[si * 2 * m + c] = acc;
                    }
                    compute_bar();
                }  // chunks
                // ---- finalize: prior, Lambda solve, DDIM update (theta stays in smem) ---------------------------------------------
                if (tid < nsamp) {
                    FinalizeParams fp = p.fin;
                    fp.C = 1;
                    fp.W = 2 * m;
                    fp.sched = sched;
                    fp.mats = p.mats + (size_t)step * (2 * m * m + m);
                    fp.step_index = step;
                    finalize_dispatch(m, fp, sAcc + tid * 2 * m, sTheta + tid * MMAX, sNoise + tid * MMAX,
                                      fp.sample_offset + i0 + tid, uni ? sPrior + tid * 2 * MMAX : nullptr);
                }
                compute_bar();
            }  // steps
            for (int e = tid; e < nsamp * m; e += 256) {
                const int si = e / m, k = e - si * m;
                p.theta[(size_t)(i0 + si) * m + k] = sTheta[si * MMAX + k];
            }
            compute_bar();
        }  // groups
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_CW + 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(256));
    }
}

cudaError_t launch_traj_tc(const TrajParams& p, int n_ctas, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(traj_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    traj_tc_kernel<<<n_ctas, TC_THREADS, TC_SMEM_BYTES, stream>>>(p);
    return cudaGetLastError();
}

int traj_tc_smem_bytes() { return TC_SMEM_BYTES; }

}  // namespace d4s
