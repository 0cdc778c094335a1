// tcgen05 score + Jacobian kernel of one JAC step (tall_posterior_sampler.py:75-94 with mode="JAC": the score of every
// (posterior sample, observation) pair, d score / d theta by vmap(jacrev), the Tweedie precision
// P = inv(sigma^2/alpha (I + sigma^2 J)); nse.py:599-625 then sums P s and P over the observations).
//
// The Jacobian is propagated in forward mode: every pair owns 1 + m consecutive tile rows, the primal activations and
// the m tangents d act / d theta_k.  All rows go through the same bf16x3 tensor-core GEMMs as in d4s_traj_tc.cu; the
// epilogues apply the ReLU of the PRIMAL row to the pair's tangent rows (h = [pre > 0], exchanged through a per-warp
// shared-memory slot: a pair never straddles a warp).  After the fp32 output layer the primal row holds eps, tangent
// row k holds column k of d eps / d theta; one thread per (pair, column) then solves (I - sigma J_eps) x = e_c
// (LU with partial pivoting in registers) and the per-unit sums over observations are written as partial sums
// part[N][C][m + m^2] = [sum_j P s | sum_j P], the layout kernel 3 (d4s_finalize.cu) reduces.
//
// Tiles: pairs are packed by units (sample i, observation chunk c) of OC observations, UT units per tile, so a unit's
// sums complete inside one tile (no atomics, fixed summation order).  The tile loop is software pipelined like the
// interleaved schedule of the trajectory kernel: the last epilogue / pair algebra of tile t-1 run while the tensor
// core works on tile t (two TMEM accumulators used alternately).
//
// Warp roles: 16 compute warps (TMEM lane quarter = warp % 4, column quarter = warp / 4), one TMA producer warp
// (weight images, 3-stage ring), one MMA issuer warp.  theta_dim <= 5 (every JAC configuration of the reference:
// SLCP 5, Lotka-Volterra 4, SIR 2); larger theta_dim runs on the fp32 SIMT kernel.
#include "d4s_finalize_core.cuh"
#include "d4s_tc_common.cuh"

#ifndef JT_TWO_PASS
#define JT_TWO_PASS 1
#endif

namespace d4s {

constexpr int JT_THREADS = (TC_CW + 4) * 32;  // 640: 16 compute warps + TMA producer + MMA issuer + 2 idle (full warpgroup)
constexpr int JT_NSTG = 3;                    // weight ring stages
constexpr int JT_MMAX = 5;                    // largest theta_dim of this kernel
constexpr int JT_PMAX = 32;                   // pairs per tile <= 4 * 8
constexpr int JT_WMAX = JT_MMAX + JT_MMAX * JT_MMAX;
// fp16 hi / lo parts lose precision below 2^-14 (subnormals: absolute quantum 2^-24), and the lo part of v is ~2^-12 v.
// So operands are kept in a scaled domain: activations x 2^4, weights x 2^sw16[l] (max |w'| in [2^9, 2^10), chosen on
// the host), accumulators therefore x 2^(4 + sw16[l]); biases are pre-scaled, the ReLU indicator is scale invariant
// and carries the 2^-sw16[l] that brings the next operand back to 2^4 x activation.  All factors are powers of two.
constexpr float ACT_SC = float(1 << D4S_ACT_SHIFT16);

constexpr int JOFF_AHI = 0;
constexpr int JOFF_ALO = JOFF_AHI + TC_A_BYTES;
constexpr int JOFF_RING = JOFF_ALO + TC_A_BYTES;
constexpr int JOFF_A = JOFF_RING + JT_NSTG * TC_STAGE_BYTES;  // float [m][256] layer-0 theta block
constexpr int JOFF_AIMG = JOFF_A + JT_MMAX * 256 * 4;         // uint4 [m][32][2] its fp16 hi / lo K-groups
constexpr int JOFF_WOUT = JOFF_AIMG + JT_MMAX * 32 * 2 * 16;  // float [m][256] output layer
constexpr int JOFF_TF = JOFF_WOUT + JT_MMAX * 256 * 4;        // float [256] time feature row (+ layer-0 bias)
constexpr int JOFF_RED = JOFF_TF + 256 * 4;                   // float [3][128][m] output partials of column quarters 1..3
constexpr int JOFF_OUT = JOFF_RED + 3 * TC_M * JT_MMAX * 4;   // float [128][m] row outputs: score (primal) / J column
constexpr int JOFF_H = JOFF_OUT + TC_M * JT_MMAX * 4;         // float [16 warps][8 pairs][16] ReLU indicators
constexpr int JOFF_PAIR = JOFF_H + TC_CW * 8 * 16 * 4;        // float [32][m + m^2] per-pair P s | P
constexpr int JT_MBUF = 3;  // pair tables of the tiles t - 1 (being finished), t and t + 1 (prepared ahead)
constexpr int JOFF_TH = JOFF_PAIR + JT_PMAX * JT_WMAX * 4;          // float [3][32][m] theta of the pair's sample
constexpr int JOFF_PI = JOFF_TH + JT_MBUF * JT_PMAX * JT_MMAX * 4;  // int [3][32] sample of the pair (-1: empty slot)
constexpr int JOFF_PJ = JOFF_PI + JT_MBUF * JT_PMAX * 4;            // int [3][32] observation of the pair
constexpr int JOFF_PS = JOFF_PJ + JT_MBUF * JT_PMAX * 4;            // float [3][32] power-of-two scale of the pair's primal row
constexpr int JOFF_BOUT = JOFF_PS + JT_MBUF * JT_PMAX * 4;          // float [8]
constexpr int JOFF_BAR = JOFF_BOUT + 32;                            // mbarriers: full[3] | empty[3] | a | dn[2] | c[4]
constexpr int JOFF_TMEM = JOFF_BAR + (2 * JT_NSTG + 3 + 4) * 8;
// CTA-pair instantiation: the ring region holds 6 half-size stages (a CTA streams half of every weight tile); its barriers and
// the leader's copies for the peer's arrivals: full[6] | empty[6] | fullp[6] | ap | cp[4]
constexpr int JT_NSTG_PAIR = 6;
constexpr int JOFF_BARP = JOFF_TMEM + 16;
constexpr int JT_SMEM_BYTES = JOFF_BARP + (3 * JT_NSTG_PAIR + 1 + 4) * 8;
static_assert(JT_SMEM_BYTES <= 232448, "shared memory budget (227 KB per CTA) exceeded");
static_assert(JOFF_BAR % 8 == 0 && JOFF_AIMG % 16 == 0 && JOFF_H % 16 == 0, "alignment");

// One thread per (pair, column c): column c of P = (alpha/sigma^2) inv(I - sigma J_eps), LU with partial pivoting in
// registers.  fp32: the entries of J_eps carry the tensor core's ~1e-6 error already, so fp64 (as in the SIMT kernel's
// jac_pair) would only add a few thousand cycles of latency per tile.
template <int M>
__device__ __noinline__ void jac_pair_col(const float* out_rows, int row0, float sigma, float a_over_s2, int c, float* pair_out) {
    float a[M][M], b[M];
#pragma unroll
    for (int o = 0; o < M; ++o) {
#pragma unroll
        for (int k = 0; k < M; ++k)  // tangent row 1 + k holds d eps_o / d theta_k
            a[o][k] = ((o == k) ? 1.f : 0.f) - sigma * out_rows[(row0 + 1 + k) * M + o];
        b[o] = (o == c) ? 1.f : 0.f;
    }
    lu_solve_t<float, M>(a, b);
#pragma unroll
    for (int r = 0; r < M; ++r) pair_out[M + r * M + c] = a_over_s2 * b[r];
}

__device__ __forceinline__ void jac_pair_col_dispatch(int m, const float* out_rows, int row0, float sigma, float aos2, int c,
                                                      float* pair_out) {
    switch (m) {
        case 1: jac_pair_col<1>(out_rows, row0, sigma, aos2, c, pair_out); break;
        case 2: jac_pair_col<2>(out_rows, row0, sigma, aos2, c, pair_out); break;
        case 3: jac_pair_col<3>(out_rows, row0, sigma, aos2, c, pair_out); break;
        case 4: jac_pair_col<4>(out_rows, row0, sigma, aos2, c, pair_out); break;
        default: jac_pair_col<5>(out_rows, row0, sigma, aos2, c, pair_out); break;
    }
}

// PAIR: clusters of two CTAs share one cta_group::2 tensor-core stream, exactly like traj_tc_kernel<MT, true> (d4s_traj_tc.cu):
// the leader issues every tcgen05.mma over M = 256 rows (its tile + the peer's), each CTA streams half of every weight tile from
// the row-split fp16 image.  This kernel needs it more than the trajectory kernel does: two passes over the hi images make it
// pull 384 KB per layer and tile through a 3-stage ring (40 B/clk at the tensor core's pace), which is what bounded its MMA
// phases (15-17k cycles per layer against ~9.7k for the instructions themselves, profiles/r02_jac_tc_notes.md).
template <bool PAIR>
__global__ void __launch_bounds__(JT_THREADS, 1) jac_tc_kernel(const JacTcParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sAhi = smem + JOFF_AHI;
    uint8_t* sAlo = smem + JOFF_ALO;
    float* sA = reinterpret_cast<float*>(smem + JOFF_A);
    uint4* sAimg = reinterpret_cast<uint4*>(smem + JOFF_AIMG);
    float* sWout = reinterpret_cast<float*>(smem + JOFF_WOUT);
    float* sTf = reinterpret_cast<float*>(smem + JOFF_TF);
    float* sRed = reinterpret_cast<float*>(smem + JOFF_RED);
    float* sOut = reinterpret_cast<float*>(smem + JOFF_OUT);
    float* sH = reinterpret_cast<float*>(smem + JOFF_H);
    float* sPair = reinterpret_cast<float*>(smem + JOFF_PAIR);
    float* sTh = reinterpret_cast<float*>(smem + JOFF_TH);
    int* sPI = reinterpret_cast<int*>(smem + JOFF_PI);
    int* sPJ = reinterpret_cast<int*>(smem + JOFF_PJ);
    float* sPS = reinterpret_cast<float*>(smem + JOFF_PS);
    float* sBout = reinterpret_cast<float*>(smem + JOFF_BOUT);
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(smem + JOFF_TMEM);
    const uint32_t bar0 = smem_u32(smem + JOFF_BAR);
    constexpr int NSTG = PAIR ? JT_NSTG_PAIR : JT_NSTG;
    constexpr int STG_BYTES = PAIR ? TC_STAGE_BYTES / 2 : TC_STAGE_BYTES;
    const uint32_t barp0 = smem_u32(smem + JOFF_BARP);
    auto bar_full = [&](int s) { return PAIR ? barp0 + 8u * s : bar0 + 8u * s; };
    auto bar_empty = [&](int s) { return PAIR ? barp0 + 8u * (NSTG + s) : bar0 + 8u * (JT_NSTG + s); };
    const uint32_t bar_a = bar0 + 8u * (2 * JT_NSTG);      // A operand ready for the MMA warp
    // accumulator ready for the epilogue: two barriers used alternately, layer by layer (the MMA warp may complete the last layer
    // of a tile and the first layer of the next one before the compute warps look at the first)
    auto bar_dn = [&](unsigned n) { return bar0 + 8u * (2 * JT_NSTG + 1 + (n & 1u)); };
    // block cb of a mid epilogue is in tensor memory (every compute warp has converted its cb-th 16-column block): K-steps
    // {cq * nblk + cb, cq = 0..3} of the next layer may be issued
    auto bar_c = [&](int cb) { return bar0 + 8u * (2 * JT_NSTG + 3 + cb); };
    // i-th K-step of a TS-form layer in issue order (K = 64 nblk): block index first, column quarter second
    auto ks_of = [](int i, int nblk) { return (i & 3) * nblk + (i >> 2); };
    // CTA pair: the leader's copies of the barriers the peer CTA arrives on
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = crank == 0u;
    auto bar_fullp = [&](int s) { return barp0 + 8u * (2 * NSTG + s); };
    const uint32_t bar_ap = barp0 + 8u * (3 * NSTG);
    auto bar_cp = [&](int cb) { return barp0 + 8u * (3 * NSTG + 1 + cb); };
    const uint32_t arr_a = (PAIR && !leader) ? mapa_shared(bar_ap, 0u) : bar_a;
    auto arrive_c = [&](int cb) {
        if (PAIR && !leader) mbar_arrive_cluster(mapa_shared(bar_cp(cb), 0u));
        else mbar_arrive(bar_c(cb));
    };
    auto arrive_a = [&]() {
        if (PAIR && !leader) mbar_arrive_cluster(arr_a);
        else mbar_arrive(bar_a);
    };

    const NetDev& net = p.net;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m = net.m, L = net.L, Lh = L - 2;
    const int H0 = net.Hp[0];
    const int R = 1 + m;        // tile rows per pair
    const int PW = p.PW;        // pairs per TMEM lane quarter (32 / R, at most 8)
    const int PT = 4 * PW;      // pairs per tile
    const int W = m + m * m;

    if (tid == 0) {
        for (int s = 0; s < NSTG; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_empty(s), 1);
            if (PAIR) mbar_init(bar_fullp(s), 1);
        }
        mbar_init(bar_a, TC_CW * 32);
        if (PAIR) mbar_init(bar_ap, TC_CW * 32);
        mbar_init(bar_dn(0), 1);
        mbar_init(bar_dn(1), 1);
        for (int cb = 0; cb < 4; ++cb) {
            mbar_init(bar_c(cb), TC_CW * 32);
            if (PAIR) mbar_init(bar_cp(cb), TC_CW * 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_CW + 1) {  // TMEM: two 256-column fp32 accumulators, used alternately layer by layer
        if (PAIR) {  // (the same warp of both CTAs of the pair)
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        }
    }
    // constant tables; operand images zeroed once (rows that belong to no pair are never written again)
    for (int e = tid; e < 2 * TC_A_BYTES / 16; e += JT_THREADS) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0, 0, 0, 0);
    {
        const int Kout = net.Hp[L - 2];
        for (int e = tid; e < m * Kout; e += JT_THREADS) sWout[(e / Kout) * 256 + (e % Kout)] = __ldg(net.WoutT + e);
        // the operands carry 2^4 x activation (ACT_SC), see the note at store_split8_f16: layer 0 is built in that scale
        // (nets with a theta-embedding MLP: the layer-0 sample part and its tangents come per sample from theta_embed_kernel)
        const bool emb = p.base_pre != nullptr;
        for (int e = tid; e < (emb ? 0 : m * H0); e += JT_THREADS) sA[(e / H0) * 256 + (e % H0)] = ACT_SC * __ldg(net.A + e);
        for (int e = tid; e < H0; e += JT_THREADS) sTf[e] = ACT_SC * __ldg(p.time_feat + e);
        if (tid < 8) sBout[tid] = (tid < m) ? __ldg(net.bout + tid) : 0.f;
        // fp16 hi / lo K-groups of the theta block (the layer-0 tangents are masked copies of them)
        for (int e = tid; e < (emb ? 0 : m * (H0 / 8)); e += JT_THREADS) {
            const int k = e / (H0 / 8), kg = e - k * (H0 / 8);
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int q2 = 0; q2 < 4; ++q2) {
                const float x0 = ACT_SC * __ldg(net.A + (size_t)k * H0 + 8 * kg + 2 * q2);
                const float x1 = ACT_SC * __ldg(net.A + (size_t)k * H0 + 8 * kg + 2 * q2 + 1);
                const __half2 h16 = __floats2half2_rn(x0, x1);
                const float2 hf = __half22float2(h16);
                const __half2 l16 = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
                hi[q2] = *reinterpret_cast<const uint32_t*>(&h16);
                lo[q2] = *reinterpret_cast<const uint32_t*>(&l16);
            }
            sAimg[(k * 32 + kg) * 2] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            sAimg[(k * 32 + kg) * 2 + 1] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_d = *sTmem;

    // (pair: both CTAs run the leader's number of tiles -- a tile index past the end has no valid pair slot and writes nothing)
    int my_tiles = 0;
    for (int t = PAIR ? (int)(blockIdx.x & ~1u) : (int)blockIdx.x; t < p.n_tiles; t += gridDim.x) ++my_tiles;

    // setmaxnreg: launched with 96 registers per thread; the service warpgroup keeps 32, the compute warps grow to 112
    if (warp >= TC_CW + 2) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");  // idle warps: complete the service warpgroup
    } else if (warp == TC_CW) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        // ================= TMA producer: streams the weight images in consumption order ===========================
        if (lane == 0) {
            unsigned it = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                for (int l = 1; l <= Lh; ++l) {
                    const int K = net.Hp[l - 1], Nn = net.Hp[l];
                    const uint32_t bytes = (uint32_t)Nn * 64u;
                    // fp16 hi | lo; pair: the image split by output rows, this CTA's half of a K-step contiguous (hi | lo)
                    const uint8_t* src = PAIR ? net.wimg16p[l] + crank * (bytes / 2) : net.wimg16[l];
                    const uint32_t mine = PAIR ? bytes / 2 : bytes;
                    // two passes over K (see the MMA issuer): cross terms first (hi | lo images), then the main terms (hi image only)
                    for (int pass = 0; pass < (JT_TWO_PASS ? 2 : 1); ++pass) {
                        const uint32_t nb = (JT_TWO_PASS && pass == 1) ? mine / 2 : mine;
                        for (int i = 0; i < K / 16; ++i, ++it) {
                            const int ks = (l > 1) ? ks_of(i, K / 64) : i;  // the order the MMA warp consumes them in
                            const int s = it % NSTG;
                            const uint32_t ph = (it / NSTG) & 1u;
                            mbar_wait(bar_empty(s), ph ^ 1u, p.error_flag);
                            mbar_expect_tx(bar_full(s), nb);
                            tma_bulk_g2s(smem_u32(smem + JOFF_RING + s * STG_BYTES), src + (size_t)ks * bytes, nb, bar_full(s));
                        }
                    }
                }
            }
        }
    } else if (warp == TC_CW + 1) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        // ================= MMA issuer ===================================================================================
        if (PAIR && !leader) {
            // peer CTA of a pair: lane 0 forwards "ring stage full" to the leader in consumption order
            if (lane == 0) {
                unsigned it = 0;
                const uint32_t fwd0 = mapa_shared(bar_fullp(0), 0u);
                for (int ti = 0; ti < my_tiles; ++ti)
                    for (int l = 1; l <= Lh; ++l)
                        for (int i = 0; i < (JT_TWO_PASS ? 2 : 1) * (net.Hp[l - 1] / 16); ++i, ++it) {
                            const int s = it % NSTG;
                            mbar_wait(bar_full(s), (it / NSTG) & 1u, p.error_flag);
                            mbar_arrive_cluster(fwd0 + 8u * s);
                        }
            }
        } else
        if (lane == 0) {
            unsigned it = 0, a_cnt = 0, lc = 0, c_par = 0;  // c_par: bit cb = phase parity of bar_c(cb) to wait for next
            const uint32_t a_hi0 = smem_u32(sAhi), a_lo0 = smem_u32(sAlo);
            for (int ti = 0; ti < my_tiles; ++ti) {
                for (int l = 1; l <= Lh; ++l) {
                    const int K = net.Hp[l - 1], Nn = net.Hp[l];
                    // fp16x3: a w ~= a_lo w_hi + a_hi w_lo + a_hi w_hi with fp16 hi / lo parts (~2^-22 per product; the bf16x3
                    // split of the GAUSS kernel is not enough here: the Jacobian goes through inv(I - sigma J), which
                    // amplifies the GEMM error by its condition number)
                    const uint32_t idesc = umma_idesc_fmt(Nn, false, false, PAIR ? 256 : TC_M);
                    const int Nb = PAIR ? Nn / 2 : Nn;  // B rows held by this CTA
                    const uint32_t b_lbo = (uint32_t)Nb * 16u, b_sbo = 128u;
                    const uint32_t a_lbo = 2048u, a_sbo = 128u;
                    const bool prof_on = p.prof != nullptr && blockIdx.x == 0;
                    long long pt0 = prof_on ? clock64() : 0;
                    if (l == 1) {  // layer 0 of the tile is in shared memory (layers l > 1 take their operand from tensor memory)
                        mbar_wait(bar_a, a_cnt & 1u, p.error_flag);
                        if (PAIR) mbar_wait(bar_ap, a_cnt & 1u, p.error_flag);  // the peer's operand image
                        ++a_cnt;
                        // this layer's accumulator is the region the previous tile's last (TS-form) layer reads its A operand from
                        if (lc > 0 && Lh >= 2) mbar_wait(bar_dn(lc - 1u), ((lc - 1u) >> 1) & 1u, p.error_flag);
                    }
                    long long pt1 = 0, pfull = 0;
                    if (prof_on) {
                        pt1 = clock64();
                        p.prof[0] += pt1 - pt0;  // MMA warp: waiting for the operand image
                    }
                    tc_fence_after();
                    const uint32_t dacc = tmem_d + ((lc & 1u) ? 256u : 0u);
                    const uint32_t dprev = tmem_d + ((lc & 1u) ? 0u : 256u);  // accumulator of the previous layer = TS-form A operand
                    ++lc;
                    // The tensor core truncates when it adds into the accumulator: every accumulating instruction loses a fraction of an
                    // ulp of the RUNNING SUM.  All cross terms (2^-11 of the result) are therefore summed first, while the accumulator
                    // is still small, and the K / 16 main-term instructions follow: a third as many truncations at full magnitude (the
                    // remaining, now uniform, deficit is what acc_debias() corrects in the epilogues).  Costs one more pass of the hi images.
                    for (int pass = 0; pass < (JT_TWO_PASS ? 2 : 1); ++pass) {
                        for (int i = 0; i < K / 16; ++i, ++it) {
                            const int ks = (l > 1) ? ks_of(i, K / 64) : i;
                            if (l > 1 && pass == 0 && (i & 3) == 0) {  // block i / 4 of the previous layer's epilogue is in tensor memory
                                mbar_wait(bar_c(i >> 2), (c_par >> (i >> 2)) & 1u, p.error_flag);
                                if (PAIR) mbar_wait(bar_cp(i >> 2), (c_par >> (i >> 2)) & 1u, p.error_flag);
                                c_par ^= 1u << (i >> 2);
                                tc_fence_after();
                            }
                            const bool first = pass == 0 && i == 0;  // first instruction of the layer overwrites the accumulator
                            const int s = it % NSTG;
                            const uint32_t ph = (it / NSTG) & 1u;
                            const long long pf0 = prof_on ? clock64() : 0;
                            mbar_wait(bar_full(s), ph, p.error_flag);
                            if (PAIR) mbar_wait(bar_fullp(s), ph, p.error_flag);  // the peer's half of the tile
                            if (prof_on) pfull += clock64() - pf0;
                            tc_fence_after();
                            const uint32_t wb = smem_u32(smem + JOFF_RING + s * STG_BYTES);
                            const uint64_t b_hi = umma_desc(wb, b_lbo, b_sbo);
                            const uint64_t b_lo = umma_desc(wb + (uint32_t)Nb * 32u, b_lbo, b_sbo);
                            const uint64_t a_hi = umma_desc(a_hi0 + ks * 4096, a_lbo, a_sbo);
                            const uint64_t a_lo = umma_desc(a_lo0 + ks * 4096, a_lbo, a_sbo);
                            // layers l > 1: A = the previous layer's accumulator region, converted in place by the mid epilogue
                            // (K-step ks: fp16 hi parts in columns 16 ks .. +7, lo parts in +8 .. +15)
                            const uint32_t a_t = dprev + (uint32_t)ks * 16u;
                            if (PAIR) {
                                if (!JT_TWO_PASS || pass == 0) {
                                    if (l > 1) {
                                        tc_mma2_bf16_ts(dacc, a_t + 8u, b_hi, idesc, first ? 0u : 1u);
                                        tc_mma2_bf16_ts(dacc, a_t, b_lo, idesc, 1u);
                                    } else {
                                        tc_mma2_bf16(dacc, a_lo, b_hi, idesc, first ? 0u : 1u);
                                        tc_mma2_bf16(dacc, a_hi, b_lo, idesc, 1u);
                                    }
                                }
                                if (!JT_TWO_PASS || pass == 1) {
                                    if (l > 1) tc_mma2_bf16_ts(dacc, a_t, b_hi, idesc, 1u);
                                    else tc_mma2_bf16(dacc, a_hi, b_hi, idesc, 1u);
                                }
                                tc_commit2(bar_empty(s));
                            } else {
                                if (!JT_TWO_PASS || pass == 0) {
                                    if (l > 1) {
                                        tc_mma_bf16_ts(dacc, a_t + 8u, b_hi, idesc, first ? 0u : 1u);
                                        tc_mma_bf16_ts(dacc, a_t, b_lo, idesc, 1u);
                                    } else {
                                        tc_mma_bf16(dacc, a_lo, b_hi, idesc, first ? 0u : 1u);
                                        tc_mma_bf16(dacc, a_hi, b_lo, idesc, 1u);
                                    }
                                }
                                if (!JT_TWO_PASS || pass == 1) {
                                    if (l > 1) tc_mma_bf16_ts(dacc, a_t, b_hi, idesc, 1u);
                                    else tc_mma_bf16(dacc, a_hi, b_hi, idesc, 1u);
                                }
                                tc_commit(bar_empty(s));
                            }
                        }
                    }
                    if (PAIR) tc_commit2(bar_dn(lc - 1u));
                    else tc_commit(bar_dn(lc - 1u));
                    if (prof_on) {  // profiling only: drain, so that slot 3 is issue -> complete of this layer
                        const long long pt2 = clock64();
                        mbar_wait(bar_dn(lc - 1u), ((lc - 1u) >> 1) & 1u, p.error_flag);
                        p.prof[1] += pfull;                 // waiting for ring stages
                        p.prof[2] += pt2 - pt1;             // issue loop (including those waits)
                        p.prof[3] += clock64() - pt1;       // operand ready -> accumulator complete
                        p.prof[4] += 1;                     // layers
                    }
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
        // ================= compute warps ==================================================================================
        const int q = warp & 3, cq = warp >> 2;  // TMEM lane quarter, column quarter
        const int row = 32 * q + lane;           // the TMEM lane / tile row this thread owns in the epilogues
        const int pl = lane / R;                 // pair slot inside the lane quarter
        const int rc = lane - pl * R;            // 0 = primal row, 1 + k = tangent k
        const bool rvalid = pl < PW;             // (lanes past PW * R own no row)
        const float bsel = (rvalid && rc == 0) ? 1.f : 0.f;
        float* hslot = sH + (warp * 8 + (rvalid ? pl : 0)) * 16;
        const int nkg = H0 / 8;
        unsigned d_cnt = 0;
        const float sigma = __ldg(p.sched + D4S_S_SIGMA), inv_sigma = __ldg(p.sched + D4S_S_INV_SIGMA);
        const float aos2 = __ldg(p.sched + D4S_S_A_OVER_S2);

        // pair slot -> (sample, observation) of a tile, and the sample's theta (double buffered: tile t-1 is still
        // being finished while tile t starts)
        auto prepare_meta = [&](int tile, int buf) {
            if (tid < PT) {
                const int u = tid / p.OC, o = tid - u * p.OC;
                const long long gu = (long long)tile * p.UT + u;
                int i = -1, j = 0;
                if (u < p.UT && gu < (long long)p.N * p.C) {
                    const int ii = (int)(gu / p.C), ch = (int)(gu - (long long)ii * p.C);
                    j = ch * p.OC + o;
                    if (j < p.n) i = ii;
                    // batched contexts: sample ii belongs to context ii / ctx_samples, whose observation block starts at row
                    // (ii / ctx_samples) * n of obs_feat
                    if (p.ctx_samples > 0) j += (ii / p.ctx_samples) * p.n;
                }
                sPI[buf * JT_PMAX + tid] = i;
                sPJ[buf * JT_PMAX + tid] = j;
                // The fp16 operands hold |value| < 65504 / 2^4.  A primal row is positively homogeneous in its activations (ReLU;
                // biases scaled alongside), so a sample far outside the usual range of normalised parameters (max |theta_k| > 16:
                // diverged trajectories of untrained nets) runs with its primal row scaled by a power of two and is unscaled
                // after the output layer; tangent rows are derivatives and do not scale.  |theta| <= 16: scale 1, nothing changes.
                float mx = 0.f;
                if (i >= 0)
                    for (int k = 0; k < m; ++k) mx = fmaxf(mx, fabsf(__ldg(p.theta + (size_t)i * m + k)));
                int ex = 0;
                if (mx > 16.f && mx < 3.0e38f) {
                    (void)frexpf(mx, &ex);  // mx = f 2^ex, f in [0.5, 1)
                    ex -= 4;
                }
                sPS[buf * JT_PMAX + tid] = ldexpf(1.f, -ex);
            }
            for (int e = tid; e < PT * m; e += TC_CT) {
                const int pp = e / m, k = e - pp * m;
                const int u = pp / p.OC;
                const long long gu = (long long)tile * p.UT + u;
                float v = 0.f;
                if (u < p.UT && gu < (long long)p.N * p.C) v = __ldg(p.theta + (size_t)(gu / p.C) * m + k);
                sTh[(buf * JT_PMAX + pp) * JT_MMAX + k] = v;
            }
        };
        // layer 0: lane = pair, warp = k-groups warp, warp + 16.  z = time_feat + A theta_i + obs_feat_j; the primal row
        // stores relu(z), tangent row k the theta-block column masked by [z > 0].
        auto do_layer0 = [&](int buf) {
            const int pp = lane;
            const int i = (pp < PT) ? sPI[buf * JT_PMAX + pp] : -1;
            if (i >= 0) {
                const float psc = sPS[buf * JT_PMAX + pp];
                const int j = sPJ[buf * JT_PMAX + pp];
                const int prow = 32 * (pp / PW) + (pp % PW) * R;  // tile row of the pair's primal
                const float* up = p.obs_feat + (size_t)j * H0;
                float th[JT_MMAX];
#pragma unroll
                for (int k = 0; k < JT_MMAX; ++k) th[k] = sTh[(buf * JT_PMAX + pp) * JT_MMAX + k];
                float4 u[2][2];
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int kg = warp + 16 * jj;
                    if (kg < nkg) {
                        u[jj][0] = __ldg(reinterpret_cast<const float4*>(up + 8 * kg));
                        u[jj][1] = __ldg(reinterpret_cast<const float4*>(up + 8 * kg + 4));
                    }
                }
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int kg = warp + 16 * jj;
                    if (kg < nkg) {
                        float z[8];
                        const float4 t0 = *reinterpret_cast<const float4*>(sTf + 8 * kg);
                        const float4 t1 = *reinterpret_cast<const float4*>(sTf + 8 * kg + 4);
                        z[0] = fmaf(u[jj][0].x, ACT_SC, t0.x); z[1] = fmaf(u[jj][0].y, ACT_SC, t0.y);
                        z[2] = fmaf(u[jj][0].z, ACT_SC, t0.z); z[3] = fmaf(u[jj][0].w, ACT_SC, t0.w);
                        z[4] = fmaf(u[jj][1].x, ACT_SC, t1.x); z[5] = fmaf(u[jj][1].y, ACT_SC, t1.y);
                        z[6] = fmaf(u[jj][1].z, ACT_SC, t1.z); z[7] = fmaf(u[jj][1].w, ACT_SC, t1.w);
                        if (p.base_pre) {  // embedding net: A e(theta_i), precomputed per sample
                            const float* bp = p.base_pre + (size_t)i * H0 + 8 * kg;
                            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bp));
                            const float4 b1 = __ldg(reinterpret_cast<const float4*>(bp + 4));
                            z[0] = fmaf(b0.x, ACT_SC, z[0]); z[1] = fmaf(b0.y, ACT_SC, z[1]);
                            z[2] = fmaf(b0.z, ACT_SC, z[2]); z[3] = fmaf(b0.w, ACT_SC, z[3]);
                            z[4] = fmaf(b1.x, ACT_SC, z[4]); z[5] = fmaf(b1.y, ACT_SC, z[5]);
                            z[6] = fmaf(b1.z, ACT_SC, z[6]); z[7] = fmaf(b1.w, ACT_SC, z[7]);
                        } else {
#pragma unroll
                            for (int k = 0; k < JT_MMAX; ++k) {
                                if (k < m) {
                                    const float4 a0 = *reinterpret_cast<const float4*>(sA + k * 256 + 8 * kg);
                                    const float4 a1 = *reinterpret_cast<const float4*>(sA + k * 256 + 8 * kg + 4);
                                    z[0] = fmaf(a0.x, th[k], z[0]); z[1] = fmaf(a0.y, th[k], z[1]);
                                    z[2] = fmaf(a0.z, th[k], z[2]); z[3] = fmaf(a0.w, th[k], z[3]);
                                    z[4] = fmaf(a1.x, th[k], z[4]); z[5] = fmaf(a1.y, th[k], z[5]);
                                    z[6] = fmaf(a1.z, th[k], z[6]); z[7] = fmaf(a1.w, th[k], z[7]);
                                }
                            }
                        }
                        uint32_t mw[4];  // 16-bit lane masks of the bf16 pairs
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            mw[e] = (z[2 * e] > 0.f ? 0x0000FFFFu : 0u) | (z[2 * e + 1] > 0.f ? 0xFFFF0000u : 0u);
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = fmaxf(z[e], 0.f) * psc;
                        store_split8_f16(v, sAhi, sAlo, a_off(prow, kg));
                        if (p.tan_pre) {  // embedding net: tangent k of the layer-0 input = A d e / d theta_k of this sample, masked
                            for (int k = 0; k < m; ++k) {
                                const float* tp = p.tan_pre + ((size_t)i * m + k) * H0 + 8 * kg;
                                const float4 t0 = __ldg(reinterpret_cast<const float4*>(tp));
                                const float4 t1 = __ldg(reinterpret_cast<const float4*>(tp + 4));
                                float tv[8];
                                tv[0] = z[0] > 0.f ? ACT_SC * t0.x : 0.f; tv[1] = z[1] > 0.f ? ACT_SC * t0.y : 0.f;
                                tv[2] = z[2] > 0.f ? ACT_SC * t0.z : 0.f; tv[3] = z[3] > 0.f ? ACT_SC * t0.w : 0.f;
                                tv[4] = z[4] > 0.f ? ACT_SC * t1.x : 0.f; tv[5] = z[5] > 0.f ? ACT_SC * t1.y : 0.f;
                                tv[6] = z[6] > 0.f ? ACT_SC * t1.z : 0.f; tv[7] = z[7] > 0.f ? ACT_SC * t1.w : 0.f;
                                store_split8_f16(tv, sAhi, sAlo, a_off(prow + 1 + k, kg));
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < JT_MMAX; ++k) {
                                if (k < m) {
                                    const uint4 hi = sAimg[(k * 32 + kg) * 2], lo = sAimg[(k * 32 + kg) * 2 + 1];
                                    const int off = a_off(prow + 1 + k, kg);
                                    *reinterpret_cast<uint4*>(sAhi + off) = make_uint4(hi.x & mw[0], hi.y & mw[1], hi.z & mw[2], hi.w & mw[3]);
                                    *reinterpret_cast<uint4*>(sAlo + off) = make_uint4(lo.x & mw[0], lo.y & mw[1], lo.z & mw[2], lo.w & mw[3]);
                                }
                            }
                        }
                    }
                }
            } else if (pp < PT) {  // empty pair slot: its rows must not carry the previous tile's values
                const int prow = 32 * (pp / PW) + (pp % PW) * R;
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int kg = warp + 16 * jj;
                    if (kg < nkg)
                        for (int r = 0; r < R; ++r) {
                            const int off = a_off(prow + r, kg);
                            *reinterpret_cast<uint4*>(sAhi + off) = make_uint4(0, 0, 0, 0);
                            *reinterpret_cast<uint4*>(sAlo + off) = make_uint4(0, 0, 0, 0);
                        }
                }
            }
            fence_proxy_async();
            tc_fence_before();
            arrive_a();
        };
        // one 16-column block of an accumulator row.  Primal: x = d + bias, a = relu(x).  Tangent: a = d * [x_primal > 0].
        // `hval` = what the indicator holds for an active unit: 2^-sw16[l] (operand of the next layer) or 1 (last layer).
        auto ld_act = [&](uint32_t dacc, int col0, const float* bias16, float hval, float deb, float psb, float (&a)[16]) {
            uint32_t d[16];
            tc_ld16(dacc + ((uint32_t)(32 * q) << 16) + (uint32_t)col0, d);
            float4 bb[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) bb[e] = __ldg(reinterpret_cast<const float4*>(bias16) + e);
            tc_wait_ld();
            float x[16];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                x[4 * e] = fmaf(bb[e].x, psb, __uint_as_float(d[4 * e]) * deb);
                x[4 * e + 1] = fmaf(bb[e].y, psb, __uint_as_float(d[4 * e + 1]) * deb);
                x[4 * e + 2] = fmaf(bb[e].z, psb, __uint_as_float(d[4 * e + 2]) * deb);
                x[4 * e + 3] = fmaf(bb[e].w, psb, __uint_as_float(d[4 * e + 3]) * deb);
            }
            __syncwarp();  // the previous block's indicators have been read by every lane
            if (bsel != 0.f) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    *reinterpret_cast<float4*>(hslot + 4 * e) =
                        make_float4(x[4 * e] > 0.f ? hval : 0.f, x[4 * e + 1] > 0.f ? hval : 0.f, x[4 * e + 2] > 0.f ? hval : 0.f,
                                    x[4 * e + 3] > 0.f ? hval : 0.f);
            }
            __syncwarp();
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float4 h = *reinterpret_cast<const float4*>(hslot + 4 * e);
                a[4 * e] = rvalid ? x[4 * e] * h.x : 0.f;
                a[4 * e + 1] = rvalid ? x[4 * e + 1] * h.y : 0.f;
                a[4 * e + 2] = rvalid ? x[4 * e + 2] * h.z : 0.f;
                a[4 * e + 3] = rvalid ? x[4 * e + 3] * h.w : 0.f;
            }
        };
        // epilogue of the last hidden layer + output layer (fp32) -> sOut[row][o]: score of the primal row, column k of
        // d eps / d theta for tangent row k
        auto do_last_epilogue = [&](uint32_t dacc, int buf) {
            const float prs = rvalid ? sPS[buf * JT_PMAX + q * PW + pl] : 1.f;  // scale of this row's pair (primal rows use it)
            const float psb = bsel * prs, inv_prs = 1.f / prs;
            const int Nn = net.Hp[L - 2];
            const int qcols = Nn >> 2, nblk = qcols / 16;
            const float* bias_l = net.bias16[L - 2] + cq * qcols;
            const float uns = exp2f(-(float)(D4S_ACT_SHIFT16 + net.sw16[L - 2]));  // accumulator scale of the last hidden layer
            const float deb_last = acc_debias(net.Hp[L - 3]);  // (truncating accumulation of the tensor core, d4s_tc_common.cuh)
            float po[JT_MMAX];
#pragma unroll
            for (int o = 0; o < JT_MMAX; ++o) po[o] = 0.f;
            for (int cb = 0; cb < nblk; ++cb) {
                const int col0 = cq * qcols + 16 * cb;
                float a[16];
                ld_act(dacc, col0, bias_l + 16 * cb, 1.f, deb_last, psb, a);
#pragma unroll
                for (int e = 0; e < 16; e += 4) {
                    float4 w4[JT_MMAX];
#pragma unroll
                    for (int o = 0; o < JT_MMAX; ++o) w4[o] = *reinterpret_cast<const float4*>(sWout + o * 256 + col0 + e);
#pragma unroll
                    for (int o = 0; o < JT_MMAX; ++o) {
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
                for (int o = 0; o < JT_MMAX; ++o)
                    if (o < m) sRed[((cq - 1) * TC_M + row) * JT_MMAX + o] = po[o];
            }
            tc_fence_before();
            compute_bar();
            if (cq == 0) {
#pragma unroll
                for (int o = 0; o < JT_MMAX; ++o) {
                    if (o < m) {
                        const float v = uns * ((po[o] + sRed[row * JT_MMAX + o]) +
                                               (sRed[(TC_M + row) * JT_MMAX + o] + sRed[(2 * TC_M + row) * JT_MMAX + o]));
                        // primal: score = -(eps + b) / sigma (nse.py:166); tangent k: d eps_o / d theta_k
                        sOut[row * m + o] = (rc == 0) ? -(v * inv_prs + sBout[o]) * inv_sigma : v;
                    }
                }
            }
            compute_bar();
        };
        // per-pair Tweedie precision and P s, then the per-unit sums over observations -> part[N][C][W]
        auto do_pairs = [&](int tile, int buf) {
            if (tid < PT * m) {
                const int pp = tid / m, c = tid - pp * m;
                if (sPI[buf * JT_PMAX + pp] >= 0) {
                    const int prow = 32 * (pp / PW) + (pp % PW) * R;
                    jac_pair_col_dispatch(m, sOut, prow, sigma, aos2, c, sPair + pp * W);
                }
            }
            compute_bar();
            if (tid < PT * m) {
                const int pp = tid / m, r = tid - pp * m;
                if (sPI[buf * JT_PMAX + pp] >= 0) {
                    const int prow = 32 * (pp / PW) + (pp % PW) * R;
                    float v = 0.f;
                    for (int k = 0; k < m; ++k) v = fmaf(sPair[pp * W + m + r * m + k], sOut[prow * m + k], v);
                    sPair[pp * W + r] = v;
                }
            }
            compute_bar();
            for (int e = tid; e < p.UT * W; e += TC_CT) {
                const int u = e / W, w = e - u * W;
                const long long gu = (long long)tile * p.UT + u;
                if (gu < (long long)p.N * p.C) {
                    float acc = 0.f;
                    for (int o = 0; o < p.OC; ++o) {
                        const int pp = u * p.OC + o;
                        if (sPI[buf * JT_PMAX + pp] >= 0) {
                            const int pj = sPJ[buf * JT_PMAX + pp];
                            const float wj = p.obs_weight ? __ldg(p.obs_weight + (p.ctx_samples > 0 ? pj % p.n : pj)) : 1.f;
                            acc = fmaf(wj, sPair[pp * W + w], acc);
                        }
                    }
                    p.part[(size_t)gu * W + w] = acc;  // gu = i * C + chunk
                }
            }
            compute_bar();  // the pair tables of this buffer are rewritten next
        };

        // Schedule (one tile = layer 0 -> tensor-core layers 1 .. Lh -> fp32 output layer -> pair algebra).  The tensor core is the
        // critical resource, so everything else runs underneath it:
        //   while layer 1 of tile t runs:        last epilogue + output layer + pair algebra of tile t - 1
        //   mid epilogue of tile t:               converted IN PLACE in tensor memory (the next layer takes its A operand from there,
        //                                         TS form), handed over block by block: the next layer starts after the first block
        //   while the last layer of tile t runs:  layer 0 of tile t + 1 into the shared-memory images (free since layer 1 completed),
        //                                         pair tables of tile t + 2
        // so layer 1 of tile t + 1 starts the moment the last layer of tile t has completed.
        if (my_tiles > 0) prepare_meta(blockIdx.x, 0);
        compute_bar();
        if (my_tiles > 0) do_layer0(0);
        if (my_tiles > 1) prepare_meta(blockIdx.x + (int)gridDim.x, 1);
        compute_bar();
        const bool cprof = p.prof != nullptr && blockIdx.x == 0 && tid == 0;
        long long ct = cprof ? clock64() : 0;
#define JPROF(slot)                            \
    do {                                       \
        if (cprof) {                           \
            const long long now_ = clock64();  \
            p.prof[slot] += now_ - ct;         \
            ct = now_;                         \
        }                                      \
    } while (0)
        auto wait_layer = [&]() -> uint32_t {  // next tensor-core layer in issue order; returns its accumulator
            const uint32_t dacc = tmem_d + ((d_cnt & 1u) ? 256u : 0u);
            mbar_wait(bar_dn(d_cnt), (d_cnt >> 1) & 1u, p.error_flag);
            ++d_cnt;
            tc_fence_after();
            return dacc;
        };
        uint32_t dacc_last = 0;  // accumulator of the last layer of the previous tile
        for (int ps = 0; ps < my_tiles; ++ps) {
            const int tile = blockIdx.x + ps * (int)gridDim.x;
            const int buf = ps % JT_MBUF, pbuf = (ps + JT_MBUF - 1) % JT_MBUF;
            if (ps > 0) {
                do_last_epilogue(dacc_last, pbuf);
                JPROF(7);
                do_pairs(tile - (int)gridDim.x, pbuf);
                JPROF(10);
            }
            for (int l = 1; l < Lh; ++l) {  // hidden layers that feed another tensor-core layer
                const int Nn = net.Hp[l];
                const int qcols = Nn >> 2, nblk = qcols / 16;
                const float* bias_l = net.bias16[l] + cq * qcols;
                const float hval = exp2f(-(float)net.sw16[l]);
                const float deb_l = acc_debias(net.Hp[l - 1]);
                const float psb_cur = bsel * (rvalid ? sPS[buf * JT_PMAX + q * PW + pl] : 1.f);
                const uint32_t dacc = wait_layer();
                JPROF(8);  // waiting for a hidden layer
                for (int cb = 0; cb < nblk; ++cb) {
                    const int col0 = cq * qcols + 16 * cb;
                    float a[16];
                    ld_act(dacc, col0, bias_l + 16 * cb, hval, deb_l, psb_cur, a);
                    // the next layer's A operand goes back to TENSOR MEMORY, in place over the 16 accumulator columns just read:
                    // columns col0 .. +7 = fp16 hi parts of K elements col0 .. col0 + 15, columns +8 .. +15 = lo parts
                    uint32_t w[16];
#pragma unroll
                    for (int k8 = 0; k8 < 2; ++k8) {
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = a[8 * k8 + e];
                        split8_regs_f16(v, w + 4 * k8, w + 8 + 4 * k8);
                    }
                    tc_st16(dacc + ((uint32_t)(32 * q) << 16) + (uint32_t)col0, w);
                    tc_wait_st();
                    tc_fence_before();
                    arrive_c(cb);
                }
                JPROF(9);  // mid epilogue
            }
            if (Lh < 2) {  // a single tensor-core layer: the shared-memory images are free once it has completed
                dacc_last = wait_layer();
                JPROF(5);
            }
            if (ps + 1 < my_tiles) do_layer0((ps + 1) % JT_MBUF);
            JPROF(6);
            if (ps + 2 < my_tiles) prepare_meta(tile + 2 * (int)gridDim.x, (ps + 2) % JT_MBUF);
            JPROF(11);
            compute_bar();
            JPROF(12);
            if (Lh >= 2) {
                dacc_last = wait_layer();  // the last layer of this tile
                JPROF(5);
            }
        }
        if (my_tiles > 0) {
            const int pbuf = (my_tiles - 1) % JT_MBUF;
            do_last_epilogue(dacc_last, pbuf);
            do_pairs(blockIdx.x + (my_tiles - 1) * (int)gridDim.x, pbuf);
        }
#undef JPROF
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();  // neither CTA leaves (or frees tensor memory) while the other may still use its shared memory
    if (warp == TC_CW + 1) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512));
    }
}

template <bool PAIR>
static cudaError_t launch_jac_tc_t(const JacTcParams& p, int n_ctas, cudaStream_t stream) {
    // the attribute is per device: keep one flag per device ordinal (a process may drive several GPUs)
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(jac_tc_kernel<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, JT_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        // setmaxnreg moves registers inside the CTA's launch allocation: with fewer than 96 registers per thread at launch
        // the compute warps' `inc` could never be served (the kernel would hang), so refuse to launch instead
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, jac_tc_kernel<PAIR>);
        if (e != cudaSuccess) return e;
        if (fa.numRegs < 96) return cudaErrorLaunchOutOfResources;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    if (!PAIR) {
        jac_tc_kernel<false><<<n_ctas, JT_THREADS, JT_SMEM_BYTES, stream>>>(p);
        return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n_ctas, 1, 1);
    cfg.blockDim = dim3(JT_THREADS, 1, 1);
    cfg.dynamicSmemBytes = JT_SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, jac_tc_kernel<true>, p);
}

// p.cta_pair: clusters of 2 CTAs (n_ctas even)
cudaError_t launch_jac_tc(const JacTcParams& p, int n_ctas, cudaStream_t stream) {
    return p.cta_pair ? launch_jac_tc_t<true>(p, n_ctas, stream) : launch_jac_tc_t<false>(p, n_ctas, stream);
}

}  // namespace d4s
