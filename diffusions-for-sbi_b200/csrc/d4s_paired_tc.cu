// tcgen05 trajectory kernel of the PAIRED mode: row i of the grid is (posterior sample i, its own observation) -- the plain
// single-observation DDIM sampler the reference runs when `x.shape[0] == shape[0]` (nse.py:253-254, 263-266) and, vmapped over
// the observations, as the covariance pre-pass of the GAUSS mode (sbibm_posterior_estimation.py:306-314:
// `vmap(lambda x: score_network.ddim(shape=(1000,), x=x, steps=100, eta=0.5))(x_obs)` followed by `torch.cov`).
//
// There is nothing to share between the rows of a tile (no separable layer 0, no sums over observations, no prior, no solve),
// so the fused tall-data kernel (d4s_traj_tc.cu: <= 8 samples per tile, per-sample state in shared memory) does not fit: with
// one observation per sample its tiles would be 6 % full.  Here a tile is 128 independent rows:
//   layer 0   z = time_feat(t) + A theta_i + obs_feat[i / obs_div]      fp32, per row (m FMAs per element)
//   layers 1 .. L-2 on the tensor cores, bf16x3 split as in d4s_traj_tc.cu (first layer SS form from shared memory, further
//             layers TS form: the epilogue converts the accumulator in place in tensor memory, handed over block by block)
//   output layer in fp32 fused into the last epilogue, score = -(eps + b) / sigma (nse.py:166)
//   DDIM update of the row: theta' = c_theta theta + c_score score + bridge_std z (nse.py:289-298), Philox noise keyed by
//             (seed; step, global sample, dim quad) exactly like kernel 4 (d4s_finalize_core.cuh), optional clipping.
// One launch runs every step.  A CTA owns the tiles b, b + grid, ... for the whole trajectory and walks the items
// (step s, tile j) in that order; consecutive items belong to different tiles (when it owns >= 2), so the last epilogue + update
// of item i - 1 run under the first tensor-core layer of item i and layer 0 of item i + 1 under its last layer, like the tile
// loop of d4s_jac_tc.cu.  theta lives in global memory (L2) between the steps: 128 x m floats per item.
#include "d4s_tc_common.cuh"

namespace d4s {

constexpr int PR_THREADS = (TC_CW + 4) * 32;  // 640: 16 compute warps + TMA producer + MMA issuer + 2 idle (full warpgroup)
constexpr int PR_NSTG = 4;                    // weight ring stages
constexpr int PR_MMAX = 5;                    // largest theta_dim of this kernel

constexpr int POFF_AHI = 0;
constexpr int POFF_ALO = POFF_AHI + TC_A_BYTES;
constexpr int POFF_RING = POFF_ALO + TC_A_BYTES;
constexpr int POFF_A = POFF_RING + PR_NSTG * TC_STAGE_BYTES;  // float [m][256] layer-0 theta block
constexpr int POFF_WOUT = POFF_A + PR_MMAX * 256 * 4;         // float [m][256] output layer
constexpr int POFF_RED = POFF_WOUT + PR_MMAX * 256 * 4;       // float [3][128][m] output partials of column quarters 1..3
constexpr int POFF_TH = POFF_RED + 3 * TC_M * PR_MMAX * 4;    // float [2][128][m] theta of the tile (double buffered by item)
constexpr int POFF_BOUT = POFF_TH + 2 * TC_M * PR_MMAX * 4;   // float [8]
constexpr int POFF_BAR = POFF_BOUT + 32;                      // mbarriers: full[4] | empty[4] | a | dn[2] | c[4]
constexpr int POFF_TMEM = POFF_BAR + (2 * PR_NSTG + 3 + 4) * 8;
constexpr int PR_SMEM_BYTES = POFF_TMEM + 16;
static_assert(PR_SMEM_BYTES <= 232448, "shared memory budget (227 KB per CTA) exceeded");
static_assert(POFF_BAR % 8 == 0 && POFF_TH % 16 == 0, "alignment");

__global__ void __launch_bounds__(PR_THREADS, 1) paired_tc_kernel(const __grid_constant__ PairedTcParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sAhi = smem + POFF_AHI;
    uint8_t* sAlo = smem + POFF_ALO;
    float* sA = reinterpret_cast<float*>(smem + POFF_A);
    float* sWout = reinterpret_cast<float*>(smem + POFF_WOUT);
    float* sRed = reinterpret_cast<float*>(smem + POFF_RED);
    float* sTh = reinterpret_cast<float*>(smem + POFF_TH);
    float* sBout = reinterpret_cast<float*>(smem + POFF_BOUT);
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(smem + POFF_TMEM);
    const uint32_t bar0 = smem_u32(smem + POFF_BAR);
    auto bar_full = [&](int s) { return bar0 + 8u * s; };
    auto bar_empty = [&](int s) { return bar0 + 8u * (PR_NSTG + s); };
    const uint32_t bar_a = bar0 + 8u * (2 * PR_NSTG);  // layer-0 operand images ready for the MMA warp
    auto bar_dn = [&](unsigned n) { return bar0 + 8u * (2 * PR_NSTG + 1 + (n & 1u)); };  // accumulator complete (alternating)
    auto bar_c = [&](int cb) { return bar0 + 8u * (2 * PR_NSTG + 3 + cb); };             // block cb of a mid epilogue in tensor memory
    auto ks_of = [](int i, int nblk) { return (i & 3) * nblk + (i >> 2); };              // issue order of a TS-form layer's K-steps

    const NetDev& net = p.net;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m = net.m, L = net.L, Lh = L - 2;
    const int H0 = net.Hp[0];

    if (tid == 0) {
        for (int s = 0; s < PR_NSTG; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_empty(s), 1);
        }
        mbar_init(bar_a, TC_CW * 32);
        mbar_init(bar_dn(0), 1);
        mbar_init(bar_dn(1), 1);
        for (int cb = 0; cb < 4; ++cb) mbar_init(bar_c(cb), TC_CW * 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_CW + 1) {  // TMEM: two 256-column fp32 accumulators, used alternately layer by layer
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    for (int e = tid; e < 2 * TC_A_BYTES / 16; e += PR_THREADS) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0, 0, 0, 0);
    {
        const int Kout = net.Hp[L - 2];
        for (int e = tid; e < m * Kout; e += PR_THREADS) sWout[(e / Kout) * 256 + (e % Kout)] = __ldg(net.WoutT + e);
        for (int e = tid; e < m * H0; e += PR_THREADS) sA[(e / H0) * 256 + (e % H0)] = __ldg(net.A + e);
        if (tid < 8) sBout[tid] = (tid < m) ? __ldg(net.bout + tid) : 0.f;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *sTmem;

    const int n_tiles = (int)((p.N + TC_M - 1) / TC_M);
    int my_tiles = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) ++my_tiles;
    const int n_steps = p.step_end - p.step_begin;
    const long long n_items = (long long)my_tiles * n_steps;

    if (warp >= TC_CW + 2) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");  // idle warps: complete the service warpgroup
    } else if (warp == TC_CW) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        // ================= TMA producer: streams the weight images in consumption order ===========================
        if (lane == 0) {
            unsigned it = 0;
            for (long long ti = 0; ti < n_items; ++ti) {
                for (int l = 1; l <= Lh; ++l) {
                    const int K = net.Hp[l - 1], Nn = net.Hp[l];
                    const uint32_t bytes = (uint32_t)Nn * 64u;
                    const uint8_t* src = net.wimg[l];
                    for (int i = 0; i < K / 16; ++i, ++it) {
                        const int ks = (l > 1) ? ks_of(i, K / 64) : i;
                        const int s = it % PR_NSTG;
                        const uint32_t ph = (it / PR_NSTG) & 1u;
                        mbar_wait(bar_empty(s), ph ^ 1u, p.error_flag);
                        mbar_expect_tx(bar_full(s), bytes);
                        tma_bulk_g2s(smem_u32(smem + POFF_RING + s * TC_STAGE_BYTES), src + (size_t)ks * bytes, bytes, bar_full(s));
                    }
                }
            }
        }
    } else if (warp == TC_CW + 1) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        // ================= MMA issuer ===================================================================================
        if (lane == 0) {
            unsigned it = 0, a_cnt = 0, lc = 0, c_par = 0;
            const uint32_t a_hi0 = smem_u32(sAhi), a_lo0 = smem_u32(sAlo);
            for (long long ti = 0; ti < n_items; ++ti) {
                for (int l = 1; l <= Lh; ++l) {
                    const int K = net.Hp[l - 1], Nn = net.Hp[l];
                    const uint32_t idesc = umma_idesc(Nn);
                    const uint32_t b_lbo = (uint32_t)Nn * 16u, b_sbo = 128u;
                    const uint32_t a_lbo = 2048u, a_sbo = 128u;
                    if (l == 1) {
                        mbar_wait(bar_a, a_cnt & 1u, p.error_flag);
                        ++a_cnt;
                        // this layer's accumulator is the region the previous item's last (TS-form) layer reads its A operand from
                        if (lc > 0 && Lh >= 2) mbar_wait(bar_dn(lc - 1u), ((lc - 1u) >> 1) & 1u, p.error_flag);
                    }
                    tc_fence_after();
                    const uint32_t dacc = tmem_d + ((lc & 1u) ? 256u : 0u);
                    const uint32_t dprev = tmem_d + ((lc & 1u) ? 0u : 256u);
                    ++lc;
                    for (int i = 0; i < K / 16; ++i, ++it) {
                        const int ks = (l > 1) ? ks_of(i, K / 64) : i;
                        if (l > 1 && (i & 3) == 0) {
                            mbar_wait(bar_c(i >> 2), (c_par >> (i >> 2)) & 1u, p.error_flag);
                            c_par ^= 1u << (i >> 2);
                            tc_fence_after();
                        }
                        const int s = it % PR_NSTG;
                        const uint32_t ph = (it / PR_NSTG) & 1u;
                        mbar_wait(bar_full(s), ph, p.error_flag);
                        tc_fence_after();
                        const uint32_t wb = smem_u32(smem + POFF_RING + s * TC_STAGE_BYTES);
                        const uint64_t b_hi = umma_desc(wb, b_lbo, b_sbo);
                        const uint64_t b_lo = umma_desc(wb + (uint32_t)Nn * 32u, b_lbo, b_sbo);
                        if (l > 1) {
                            const uint32_t a_t = dprev + (uint32_t)ks * 16u;  // hi in columns 16 ks .. +7, lo in +8 .. +15
                            tc_mma_bf16_ts(dacc, a_t, b_hi, idesc, i > 0 ? 1u : 0u);
                            tc_mma_bf16_ts(dacc, a_t + 8u, b_hi, idesc, 1u);
                            tc_mma_bf16_ts(dacc, a_t, b_lo, idesc, 1u);
                        } else {
                            const uint64_t a_hi = umma_desc(a_hi0 + ks * 4096, a_lbo, a_sbo);
                            const uint64_t a_lo = umma_desc(a_lo0 + ks * 4096, a_lbo, a_sbo);
                            tc_mma_bf16(dacc, a_hi, b_hi, idesc, i > 0 ? 1u : 0u);
                            tc_mma_bf16(dacc, a_lo, b_hi, idesc, 1u);
                            tc_mma_bf16(dacc, a_hi, b_lo, idesc, 1u);
                        }
                        tc_commit(bar_empty(s));
                    }
                    tc_commit(bar_dn(lc - 1u));
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
        // ================= compute warps ==================================================================================
        const int q = warp & 3, cq = warp >> 2;  // TMEM lane quarter, column quarter
        const int row = 32 * q + lane;           // the TMEM lane / tile row this thread owns in the epilogues
        const int nkg = H0 / 8;
        unsigned d_cnt = 0;

        auto item_tile = [&](long long it) { return blockIdx.x + (int)(it % my_tiles) * (int)gridDim.x; };
        auto item_step = [&](long long it) { return p.step_begin + (int)(it / my_tiles); };
        // theta of the item's rows -> shared memory (rows past N: zeros)
        auto load_theta = [&](long long it, int buf) {
            const long long r0 = (long long)item_tile(it) * TC_M;
            for (int e = tid; e < TC_M * m; e += TC_CT) {
                const int r = e / m, k = e - r * m;
                const long long g = r0 + r;
                sTh[(buf * TC_M + r) * PR_MMAX + k] = (g < p.N) ? p.theta[(size_t)g * m + k] : 0.f;
            }
        };
        // layer 0: lanes = rows of a lane quarter, warp = k-groups warp, warp + 16, all four row quarters in turn
        auto do_layer0 = [&](long long it, int buf) {
            const long long r0 = (long long)item_tile(it) * TC_M;
            const float* tf = p.time_feat + (size_t)item_step(it) * H0;
            for (int rq = 0; rq < 4; ++rq) {
                const int r = 32 * rq + lane;
                long long g = r0 + r;
                if (g >= p.N) g = p.N - 1;  // (padding rows: computed, never written back)
                const float* up = p.obs_feat + (size_t)(g / p.obs_div) * H0;
                float th[PR_MMAX];
#pragma unroll
                for (int k = 0; k < PR_MMAX; ++k) th[k] = (k < m) ? sTh[(buf * TC_M + r) * PR_MMAX + k] : 0.f;
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int kg = warp + 16 * jj;
                    if (kg < nkg) {
                        const float4 u0 = __ldg(reinterpret_cast<const float4*>(up + 8 * kg));
                        const float4 u1 = __ldg(reinterpret_cast<const float4*>(up + 8 * kg + 4));
                        const float4 t0 = __ldg(reinterpret_cast<const float4*>(tf + 8 * kg));
                        const float4 t1 = __ldg(reinterpret_cast<const float4*>(tf + 8 * kg + 4));
                        float z[8] = {u0.x + t0.x, u0.y + t0.y, u0.z + t0.z, u0.w + t0.w,
                                      u1.x + t1.x, u1.y + t1.y, u1.z + t1.z, u1.w + t1.w};
#pragma unroll
                        for (int k = 0; k < PR_MMAX; ++k) {
                            if (k < m) {
                                const float4 a0 = *reinterpret_cast<const float4*>(sA + k * 256 + 8 * kg);
                                const float4 a1 = *reinterpret_cast<const float4*>(sA + k * 256 + 8 * kg + 4);
                                z[0] = fmaf(a0.x, th[k], z[0]); z[1] = fmaf(a0.y, th[k], z[1]);
                                z[2] = fmaf(a0.z, th[k], z[2]); z[3] = fmaf(a0.w, th[k], z[3]);
                                z[4] = fmaf(a1.x, th[k], z[4]); z[5] = fmaf(a1.y, th[k], z[5]);
                                z[6] = fmaf(a1.z, th[k], z[6]); z[7] = fmaf(a1.w, th[k], z[7]);
                            }
                        }
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = fmaxf(z[e], 0.f);
                        store_split8(v, sAhi, sAlo, a_off(r, kg));
                    }
                }
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(bar_a);
        };
        auto wait_layer = [&]() -> uint32_t {  // next tensor-core layer in issue order; returns its accumulator
            const uint32_t dacc = tmem_d + ((d_cnt & 1u) ? 256u : 0u);
            mbar_wait(bar_dn(d_cnt), (d_cnt >> 1) & 1u, p.error_flag);
            ++d_cnt;
            tc_fence_after();
            return dacc;
        };
        // one 16-column block of an accumulator row: bias + ReLU
        auto ld_act = [&](uint32_t dacc, int col0, const float* bias16, float (&a)[16]) {
            uint32_t d[16];
            tc_ld16(dacc + ((uint32_t)(32 * q) << 16) + (uint32_t)col0, d);
            float4 bb[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) bb[e] = __ldg(reinterpret_cast<const float4*>(bias16) + e);
            tc_wait_ld();
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                a[4 * e] = fmaxf(__uint_as_float(d[4 * e]) + bb[e].x, 0.f);
                a[4 * e + 1] = fmaxf(__uint_as_float(d[4 * e + 1]) + bb[e].y, 0.f);
                a[4 * e + 2] = fmaxf(__uint_as_float(d[4 * e + 2]) + bb[e].z, 0.f);
                a[4 * e + 3] = fmaxf(__uint_as_float(d[4 * e + 3]) + bb[e].w, 0.f);
            }
        };
        // last epilogue + fp32 output layer + score + DDIM update of the item's rows
        auto do_finish = [&](uint32_t dacc, long long it, int buf) {
            const int Nn = net.Hp[L - 2];
            const int qcols = Nn >> 2, nblk = qcols / 16;
            const float* bias_l = net.bias[L - 2] + cq * qcols;
            float po[PR_MMAX];
#pragma unroll
            for (int o = 0; o < PR_MMAX; ++o) po[o] = 0.f;
            for (int cb = 0; cb < nblk; ++cb) {
                const int col0 = cq * qcols + 16 * cb;
                float a[16];
                ld_act(dacc, col0, bias_l + 16 * cb, a);
#pragma unroll
                for (int e = 0; e < 16; e += 4) {
                    float4 w4[PR_MMAX];
#pragma unroll
                    for (int o = 0; o < PR_MMAX; ++o) w4[o] = *reinterpret_cast<const float4*>(sWout + o * 256 + col0 + e);
#pragma unroll
                    for (int o = 0; o < PR_MMAX; ++o) {
                        float acc = po[o];
                        acc = fmaf(a[e], w4[o].x, acc);
                        acc = fmaf(a[e + 1], w4[o].y, acc);
                        acc = fmaf(a[e + 2], w4[o].z, acc);
                        acc = fmaf(a[e + 3], w4[o].w, acc);
                        po[o] = acc;
                    }
                }
            }
            if (cq != 0) {
#pragma unroll
                for (int o = 0; o < PR_MMAX; ++o)
                    if (o < m) sRed[((cq - 1) * TC_M + row) * PR_MMAX + o] = po[o];
            }
            tc_fence_before();
            compute_bar();
            if (cq == 0) {
                const int step = item_step(it);
                const long long g = (long long)item_tile(it) * TC_M + row;
                const float* sc = p.sched + (size_t)step * D4S_SCHED_STRIDE;
                const float inv_sigma = __ldg(sc + D4S_S_INV_SIGMA);
                const float w1 = __ldg(sc + D4S_S_S1_WEIGHT);
                float score[PR_MMAX];
#pragma unroll
                for (int o = 0; o < PR_MMAX; ++o) {
                    score[o] = 0.f;
                    if (o < m) {
                        const float v = (po[o] + sRed[row * PR_MMAX + o]) +
                                        (sRed[(TC_M + row) * PR_MMAX + o] + sRed[(2 * TC_M + row) * PR_MMAX + o]);
                        score[o] = w1 * (-(v + sBout[o]) * inv_sigma);  // score = -(eps + b) / sigma (nse.py:166)
                    }
                }
                if (g < p.N) {
                    // DDIM update of kernel 4 (d4s_finalize_core.cuh): theta' = c_theta theta + c_score score + bridge_std z
                    const double c_theta = (double)__ldg(sc + D4S_S_C_THETA);
                    double c_score = (double)__ldg(sc + D4S_S_C_SCORE);
                    const double tame = (double)__ldg(sc + D4S_S_TAME);
                    if (tame > 0.0) {
                        double nrm = 0.0;
#pragma unroll
                        for (int k = 0; k < PR_MMAX; ++k) nrm = fma((double)score[k], (double)score[k], nrm);
                        c_score = tame / (1.0 + tame * sqrt(nrm));
                    }
                    const float bstd = __ldg(sc + D4S_S_BRIDGE_STD);
                    float z[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) z[k] = 0.f;
                    if (p.noise) {
                        for (int k = 0; k < m; ++k) z[k] = __ldg(p.noise + ((size_t)step * p.N + (size_t)g) * m + k);
                    } else if (bstd != 0.f) {
#pragma unroll
                        for (int qd = 0; qd < 2; ++qd) {
                            if (4 * qd < m) {
                                float zz[4];
                                philox_normal4(p.seed, step, p.sample_offset + g, qd, zz);
                                z[4 * qd] = zz[0]; z[4 * qd + 1] = zz[1]; z[4 * qd + 2] = zz[2]; z[4 * qd + 3] = zz[3];
                            }
                        }
                    }
                    const bool clip = (p.clip_lo <= p.clip_hi) && __ldg(sc + D4S_S_CLIP) != 0.f;
#pragma unroll
                    for (int k = 0; k < PR_MMAX; ++k) {
                        if (k < m) {
                            const float th = sTh[(buf * TC_M + row) * PR_MMAX + k];
                            float v = (float)(c_theta * (double)th + c_score * (double)score[k] + (double)bstd * (double)z[k]);
                            if (clip) v = fminf(fmaxf(v, p.clip_lo), p.clip_hi);
                            p.theta[(size_t)g * m + k] = v;
                        }
                    }
                }
            }
            compute_bar();  // the row's new theta is in global memory before any thread loads it for the tile's next step
        };

        // With >= 2 tiles the items i and i + 1 belong to different tiles: layer 0 of item i + 1 is written while the last layer of
        // item i runs, the finish of item i under the first layer of item i + 1.  A single tile: every item depends on the update
        // of the previous one, the loop runs in program order.
        const bool overlap = my_tiles >= 2;
        if (n_items > 0) {
            load_theta(0, 0);
            compute_bar();
            do_layer0(0, 0);
        }
        uint32_t dacc_last = 0;
        bool pending = false;  // the finish of item it - 1 is still to do
        for (long long it = 0; it < n_items; ++it) {
            const int buf = (int)(it & 1), nbuf = buf ^ 1;
            if (pending) {
                do_finish(dacc_last, it - 1, nbuf);
                pending = false;
            }
            for (int l = 1; l < Lh; ++l) {  // hidden layers that feed another tensor-core layer: converted in place (TS form)
                const int Nn = net.Hp[l];
                const int qcols = Nn >> 2, nblk = qcols / 16;
                const float* bias_l = net.bias[l] + cq * qcols;
                const uint32_t dacc = wait_layer();
                for (int cb = 0; cb < nblk; ++cb) {
                    const int col0 = cq * qcols + 16 * cb;
                    float a[16];
                    ld_act(dacc, col0, bias_l + 16 * cb, a);
                    uint32_t w[16];
#pragma unroll
                    for (int k8 = 0; k8 < 2; ++k8) {
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = a[8 * k8 + e];
                        split8_regs(v, w + 4 * k8, w + 8 + 4 * k8);
                    }
                    tc_st16(dacc + ((uint32_t)(32 * q) << 16) + (uint32_t)col0, w);
                    tc_wait_st();
                    tc_fence_before();
                    mbar_arrive(bar_c(cb));
                }
            }
            if (Lh < 2) dacc_last = wait_layer();  // a single tensor-core layer: the operand images are free once it completed
            if (overlap) {
                if (it + 1 < n_items) {
                    load_theta(it + 1, nbuf);
                    compute_bar();
                    do_layer0(it + 1, nbuf);
                }
                if (Lh >= 2) dacc_last = wait_layer();
                pending = true;
            } else {
                if (Lh >= 2) dacc_last = wait_layer();
                do_finish(dacc_last, it, buf);
                if (it + 1 < n_items) {
                    load_theta(it + 1, nbuf);
                    compute_bar();
                    do_layer0(it + 1, nbuf);
                }
            }
        }
        if (pending) do_finish(dacc_last, n_items - 1, (int)((n_items - 1) & 1));
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_CW + 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512));
    }
}

cudaError_t launch_paired_tc(const PairedTcParams& p, int n_ctas, cudaStream_t stream) {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(paired_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PR_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, paired_tc_kernel);
        if (e != cudaSuccess) return e;
        if (fa.numRegs < 96) return cudaErrorLaunchOutOfResources;  // setmaxnreg needs the launch allocation (see d4s_traj_tc.cu)
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    paired_tc_kernel<<<n_ctas, PR_THREADS, PR_SMEM_BYTES, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace d4s
