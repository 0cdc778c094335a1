// Fused tcgen05 trajectory kernel (GAUSS): one launch runs EVERY DDIM step of the trajectory.
//
// A CTA owns a group of SG posterior samples for the whole trajectory (samples never interact, so there is
// no inter-CTA dependency and no per-step launch).  Per step and per chunk of OC observations it evaluates the
// score MLP on the 128-row tile (sample x observation pairs):
//   layer 0   separable fp32 SIMT:  z0 = A theta_i + obs_feat_j + time_feat   (theta through its embedding MLP, if any)
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
// Warp roles (640 threads): warps 0-15 compute / epilogue (TMEM lane quarter = warp % 4, column quarter = warp / 4),
//             warp 16 = TMA producer (one lane), warp 17 = TMEM allocator + MMA issuer (one lane),
//             warp 18 = scalar helper: per-step schedule / matrix rows -> smem, fp64 diffused-prior terms, per-sample
//             Lambda inverses, Philox noise (everything that depends only on theta at the start of the step, off the
//             critical path), warp 19 idle (completes the service warpgroup for setmaxnreg: service warps 64 registers,
//             compute warps 104).
// Schedule: a CTA that owns two sample groups interleaves them step by step, so that the last epilogue / scores / sums /
// tail of one group overlap the tensor-core layers of the other (two TMEM accumulators); observation chunks of one item
// are pipelined the same way.  DESIGN.md section 4.1, profiles/r01_tc_notes.md.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "d4s_finalize_core.cuh"
#include "d4s_tc_common.cuh"

namespace d4s {

constexpr int TC_THREADS = (TC_CW + 4) * 32;   // 640: + TMA producer, MMA issuer, scalar helper, one idle warp (full warpgroup)
constexpr int TC_NSTG = 4;                     // weight ring stages
#ifndef D4S_PAIR_NSTG
#define D4S_PAIR_NSTG 7
#endif
// CTA-pair instantiation: a CTA streams half of every weight tile (<= 8 KB per K-step), so the same 64 KB hold 7 half-size
// stages + one spare half-stage for the ring's mbarriers and the leader's peer barriers
constexpr int TC_NSTG_PAIR = D4S_PAIR_NSTG;
static_assert(TC_NSTG_PAIR >= 2 && TC_NSTG_PAIR <= 7, "pair ring: 2..7 half-size stages");
constexpr int TC_SG_MAX = 8;                   // samples per tile (warp w < SG sums sample w)
constexpr int TC_EMB_STRIDE = (TC_M * MMAX) / TC_SG_MAX;  // 160: widest theta-embedding layer (ping-pong in sS / sQS)
constexpr int TC_MATS_MAX = 2 * MMAX * MMAX + MMAX + 2;

// ---- shared memory carve-up (bytes) ---------------------------------------------------------------------------
constexpr int OFF_AHI = 0;
constexpr int OFF_ALO = OFF_AHI + TC_A_BYTES;
constexpr int OFF_RING = OFF_ALO + TC_A_BYTES;
constexpr int OFF_S = OFF_RING + TC_NSTG * TC_STAGE_BYTES;    // float [128][MMAX] scores (also the half-exchange buffer)
constexpr int OFF_QS = OFF_S + TC_M * MMAX * 4;                // float [128][MMAX] inv(Sigma_j) s
constexpr int OFF_BASE = OFF_QS + TC_M * MMAX * 4;             // float [SG_MAX][256] layer-0 per-sample part
constexpr int OFF_WOUT = OFF_BASE + TC_SG_MAX * 256 * 4;       // float [MMAX][256] output layer (torch layout)
constexpr int OFF_ACC = OFF_WOUT + MMAX * 256 * 4;             // float [SG_MAX][2*MMAX] sums over observations
constexpr int OFF_THETA = OFF_ACC + TC_SG_MAX * 2 * MMAX * 4;  // float [2][SG_MAX][MMAX] (two interleaved groups)
constexpr int OFF_NOISE = OFF_THETA + 2 * TC_SG_MAX * MMAX * 4;  // float [SG_MAX][MMAX] DDIM noise of the step
constexpr int OFF_PRIOR = OFF_NOISE + TC_SG_MAX * MMAX * 4;    // float [SG_MAX][2][MMAX] uniform-prior score, derivative
constexpr int OFF_SCHED = OFF_PRIOR + TC_SG_MAX * 2 * MMAX * 4;  // float [16] schedule row of the step
constexpr int OFF_MATS = OFF_SCHED + 16 * 4;                   // float [TC_MATS_MAX] Lmat | G | mu_t of the step
constexpr int OFF_BOUT = OFF_MATS + TC_MATS_MAX * 4;           // float [MMAX + 2]
constexpr int OFF_ROWS = OFF_BOUT + (MMAX + 2) * 4;            // uint8 [128] sample of row, uint8 [128] observation of row
constexpr int OFF_MINV = OFF_ROWS + 256;                       // float [SG_MAX][MMAX*MMAX] per-sample inverse of Lambda_i
constexpr int OFF_G = OFF_MINV + TC_SG_MAX * MMAX * MMAX * 4;  // float [SG_MAX][MMAX] prior term of the step
constexpr int OFF_BAR = OFF_G + TC_SG_MAX * MMAX * 4;          // mbarriers
constexpr int OFF_TMEM = OFF_BAR + (2 * TC_NSTG + 3 + 4) * 8;
constexpr int TC_SMEM_BYTES = OFF_TMEM + 16;
static_assert(TC_SMEM_BYTES <= 232448, "shared memory budget (227 KB per CTA) exceeded");
static_assert(OFF_BAR % 8 == 0 && OFF_SCHED % 16 == 0, "alignment");

// Optional phase timing (d4s_set_option("profile", 1)): thread 0 of CTA 0 accumulates clock64() deltas per phase.
#define PROF_MARK(slot)                                                     \
    do {                                                                    \
        TRACE(slot);                                                        \
        if (prof_on) {                                                      \
            const long long now_ = clock64();                               \
            p.prof[slot] += now_ - prof_t;                                  \
            prof_t = now_;                                                  \
        }                                                                   \
    } while (0)

// Event timeline of CTA 0 (-DD4S_TRACE variant builds, tests/build_variant.py + tests/gpu_trace_tc.py): lane 0 of every warp
// appends (tag, clock64) pairs for the items [TRACE_FROM, TRACE_TO); launch_traj_tc() writes the buffer to $D4S_TRACE_FILE.
// Tags: the PROF_MARK slots at the END of a phase; 20 item start, 21 / 22 before the wait for the last / a mid tensor-core layer;
// MMA warp: 30 + l first MMA of layer l issued, 40 + l all MMAs of layer l issued.
constexpr int TRACE_MAX = 256, TRACE_FROM = 40, TRACE_TO = 46;
#ifdef D4S_TRACE
#define TRACE(tag)                                                                          \
    do {                                                                                    \
        if (tr_go && lane == 0 && tr_n < TRACE_MAX) {                                       \
            p.trace[(warp * TRACE_MAX + tr_n) * 2] = (tag);                                 \
            p.trace[(warp * TRACE_MAX + tr_n) * 2 + 1] = clock64();                         \
            ++tr_n;                                                                         \
        }                                                                                   \
    } while (0)
#else
#define TRACE(tag) do { } while (0)
#endif

// Helper warp: per-sample Lambda_i = Lbase + (1-n) diag(p0_i) and its inverse, in fp64 (nse.py:647-654 builds and
// solves this system; here the inverse is prepared off the critical path, the solve is a mat-vec in the tail).
// One lane per (sample, column c): it solves Lambda_i x = e_c by LU with partial pivoting, everything in registers
// for the usual theta_dim <= 5 (an in-thread inverse with local-memory arrays costs tens of thousands of cycles
// because the L1 is almost entirely carved out as shared memory).
// Helper-warp variants of lu_solve_d / uniform_prior_terms (d4s_finalize_core.cuh) with the fp64 divisions replaced by
// reciprocals that are computed once: the helper warp is bound by the SM's fp64 issue rate (ncu: 66 % of its samples on
// the math-pipe throttle) and is on the critical path of the per-sample Lambda steps; an fp64 division is ~30 fp64-pipe
// instructions.  Results differ from the division forms by a few fp64 ulps, far below the fp32 outputs.
template <int M>
__device__ __forceinline__ void lu_solve_recip(double (&a)[M][M], double (&b)[M]) {
    double inv[M];
#pragma unroll
    for (int k = 0; k < M; ++k) {
        int p = k;
        double best = fabs(a[k][k]);
#pragma unroll
        for (int r = k + 1; r < M; ++r) {
            const double v = fabs(a[r][k]);
            if (v > best) { best = v; p = r; }
        }
#pragma unroll
        for (int r = k + 1; r < M; ++r) {
            if (r == p) {
#pragma unroll
                for (int c = 0; c < M; ++c) { const double t = a[k][c]; a[k][c] = a[r][c]; a[r][c] = t; }
                const double t = b[k]; b[k] = b[r]; b[r] = t;
            }
        }
        inv[k] = 1.0 / a[k][k];
#pragma unroll
        for (int r = k + 1; r < M; ++r) {
            const double f = a[r][k] * inv[k];
#pragma unroll
            for (int c = k + 1; c < M; ++c) a[r][c] = fma(-f, a[k][c], a[r][c]);
            b[r] = fma(-f, b[k], b[r]);
        }
    }
#pragma unroll
    for (int k = M - 1; k >= 0; --k) {
        double sacc = b[k];
#pragma unroll
        for (int c = k + 1; c < M; ++c) sacc = fma(-a[k][c], b[c], sacc);
        b[k] = sacc * inv[k];
    }
}

// inv_sig = 1 / sigma, inv_s = 1 / sqrt(alpha); p0 = a/s^2 / (1 + sigma^2 jac) is formed by the caller when needed
__device__ __noinline__ void uniform_prior_terms_recip(double th, double lo, double hi, double s, double sig, double inv_s,
                                                          double inv_sig, double& s0, double& jac) {
    const double inv_sqrt_2pi = 0.39894228040143267794;
    const double ub = (hi * s - th) * inv_sig;
    const double ua = (lo * s - th) * inv_sig;
    const double f = cdf_diff(ua, ub) * inv_s;
    const double pb = inv_sqrt_2pi * exp(-0.5 * ub * ub), pa = inv_sqrt_2pi * exp(-0.5 * ua * ua);
    const double iss = inv_sig * inv_s;
    const double fp = -(pb - pa) * iss;
    const double inv_den = 1.0 / (f + 1e-6);  // vp_diffused_priors.py:44
    s0 = fp * inv_den;
    const double fpp = -(ub * pb - ua * pa) * (iss * inv_sig);
    jac = (fpp - fp * s0) * inv_den;  // fpp / den - fp^2 / den^2
}

template <int M>
__device__ __noinline__ void lambda_inverse_col(const float* lbase, const float* qsum, const float* p0, double one_minus_n, int c,
                                                float* out) {
    double a[M][M], b[M];
#pragma unroll
    for (int r = 0; r < M; ++r) {
#pragma unroll
        for (int cc = 0; cc < M; ++cc)  // (+ the observation part of the sample's context, batched contexts only)
            a[r][cc] = (double)lbase[r * M + cc] + (qsum ? (double)__ldg(qsum + r * M + cc) : 0.0) +
                       ((r == cc) ? one_minus_n * (double)p0[r] : 0.0);
        b[r] = (r == c) ? 1.0 : 0.0;
    }
    lu_solve_recip<M>(a, b);
#pragma unroll
    for (int r = 0; r < M; ++r) out[r * M + c] = (float)b[r];
}

__device__ __forceinline__ void lambda_inverse_dispatch(int m, const float* lbase, const float* qsum, const float* p0, double omn, int c,
                                                        float* out) {
    switch (m) {
        case 1: lambda_inverse_col<1>(lbase, qsum, p0, omn, c, out); break;
        case 2: lambda_inverse_col<2>(lbase, qsum, p0, omn, c, out); break;
        case 3: lambda_inverse_col<3>(lbase, qsum, p0, omn, c, out); break;
        case 4: lambda_inverse_col<4>(lbase, qsum, p0, omn, c, out); break;
        case 5: lambda_inverse_col<5>(lbase, qsum, p0, omn, c, out); break;
        case 6: lambda_inverse_col<6>(lbase, qsum, p0, omn, c, out); break;
        case 7: lambda_inverse_col<7>(lbase, qsum, p0, omn, c, out); break;
        case 8: lambda_inverse_col<8>(lbase, qsum, p0, omn, c, out); break;
        case 9: lambda_inverse_col<9>(lbase, qsum, p0, omn, c, out); break;
        default: lambda_inverse_col<10>(lbase, qsum, p0, omn, c, out); break;
    }
}

// Layer 0 for batched contexts (D4sProblem.ctx_samples > 0): every sample of the tile has its own observation block, so
// the obs_feat rows are loaded per (sample, observation) instead of once per observation.  Not inlined (see below).
__device__ __noinline__ void layer0_contexts(const float* __restrict__ obs_feat, const float* s_base, uint8_t* a_hi, uint8_t* a_lo,
                                             int H0, int OC, int SG, int nsamp, int n, int j0, int nobs, int ctx_samples, int i0,
                                             int warp, int lane) {
    const int nkg = H0 / 8;
    for (int oj = lane; oj < OC; oj += 32) {
        const bool ovalid = oj < nobs;
        for (int si = 0; si < SG; ++si) {
            const bool valid = ovalid && si < nsamp;
            const float* up = obs_feat + ((size_t)((i0 + (valid ? si : 0)) / ctx_samples) * n + j0 + (ovalid ? oj : 0)) * H0;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int kg = warp + 16 * j;
                if (kg < nkg) {
                    float v[8];
                    if (valid) {
                        const float4 u0 = __ldg(reinterpret_cast<const float4*>(up + 8 * kg));
                        const float4 u1 = __ldg(reinterpret_cast<const float4*>(up + 8 * kg + 4));
                        const float* brow = s_base + si * 256 + 8 * kg;
                        const float4 b0 = *reinterpret_cast<const float4*>(brow);
                        const float4 b1 = *reinterpret_cast<const float4*>(brow + 4);
                        v[0] = fmaxf(u0.x + b0.x, 0.f); v[1] = fmaxf(u0.y + b0.y, 0.f);
                        v[2] = fmaxf(u0.z + b0.z, 0.f); v[3] = fmaxf(u0.w + b0.w, 0.f);
                        v[4] = fmaxf(u1.x + b1.x, 0.f); v[5] = fmaxf(u1.y + b1.y, 0.f);
                        v[6] = fmaxf(u1.z + b1.z, 0.f); v[7] = fmaxf(u1.w + b1.w, 0.f);
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = 0.f;
                    }
                    store_split8(v, a_hi, a_lo, a_off(si * OC + oj, kg));
                }
            }
        }
    }
}

// Layer-0 sample part for nets whose theta goes through an embedding MLP (nse.py:129, the JR-NMM nets): the MLP of the
// group's samples (ReLU between layers, last layer linear; all compute warps, ping-pong in two free smem buffers of
// [SG][TC_EMB_STRIDE] floats), then base[si][col] = t0[col] + sum_f A[f][col] e_f(theta_si).  Not inlined: it must not
// weigh on the register allocation of the plain-MLP path.
__device__ __noinline__ void theta_embed_base(const NetDev& net, const float* th, float* buf_a, float* buf_b, float* s_base,
                                              int SG, int H0, float t0, int tid) {
    const float* in = th;
    int in_stride = MMAX;
    for (int l = 0; l < net.emb_L; ++l) {
        const int K = net.emb_dims[l], Nn = net.emb_dims[l + 1];
        float* out = (l & 1) ? buf_b : buf_a;
        const float* Wt = net.emb_Wt[l];
        const bool relu = l < net.emb_L - 1;
        for (int idx = tid; idx < SG * Nn; idx += TC_CT) {
            const int si = idx / Nn, o = idx - si * Nn;
            float acc = __ldg(net.emb_bias[l] + o);
            // (eight weight loads in flight per step of the dependent FMA chain: one load per FMA left the chain waiting a
            // cache round trip per term; same summation order)
            int k = 0;
            for (; k + 8 <= K; k += 8) {
                float w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) w[j] = __ldg(Wt + (size_t)(k + j) * Nn + o);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc = fmaf(w[j], in[si * in_stride + k + j], acc);
            }
            for (; k < K; ++k) acc = fmaf(__ldg(Wt + (size_t)k * Nn + o), in[si * in_stride + k], acc);
            out[si * TC_EMB_STRIDE + o] = relu ? fmaxf(acc, 0.f) : acc;
        }
        compute_bar();
        in = out;
        in_stride = TC_EMB_STRIDE;
    }
    if (tid < H0) {
        float z[TC_SG_MAX];
#pragma unroll
        for (int si = 0; si < TC_SG_MAX; ++si) z[si] = t0;
        int f = 0;
        for (; f + 8 <= net.theta_cols; f += 8) {
            float a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = __ldg(net.A + (size_t)(f + j) * H0 + tid);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
#pragma unroll
                for (int si = 0; si < TC_SG_MAX; ++si) z[si] = fmaf(a[j], in[si * TC_EMB_STRIDE + f + j], z[si]);
            }
        }
        for (; f < net.theta_cols; ++f) {
            const float a = __ldg(net.A + (size_t)f * H0 + tid);
#pragma unroll
            for (int si = 0; si < TC_SG_MAX; ++si) z[si] = fmaf(a, in[si * TC_EMB_STRIDE + f], z[si]);
        }
#pragma unroll
        for (int si = 0; si < TC_SG_MAX; ++si)
            if (si < SG) s_base[si * 256 + tid] = z[si];
    }
}

// Observation sharding over ranks (TrajParams.xg): the group's sums S1 | S2 over THIS rank's observations go to every
// rank's receive buffer (plain stores to mapped peer memory over NVLink), then one flag per peer; once every rank's flag
// for this (group, step) has arrived the contributions are added in rank order -> s_acc holds the sums over ALL
// observations, bit-identical on every rank.  Slots alternate with the sequence number: a peer can be at most one step
// ahead (it needs our flag of step s before it can publish step s + 1).  Called by all compute threads.  Not inlined: it
// must not weigh on the register allocation of the single-GPU path.
__device__ __noinline__ void xchg_group_sums(const XchgDev& xg, float* s_acc, int N, int m, int n_groups, int SG, int step,
                                             int i0, int nsamp, int tid, int* error_flag) {
    const int W2 = 2 * m, cnt = nsamp * W2;
    const unsigned seq = xg.seq_base + (unsigned)step;
    const size_t per_rank = (size_t)N * W2;
    const size_t slot_off = (size_t)(seq & 1u) * xg.world * per_rank;
    const int gidx = i0 / SG;
    for (int e = tid; e < cnt * xg.world; e += TC_CT) {
        const int r = e / cnt, k = e - r * cnt;
        xg.data[r][slot_off + (size_t)xg.rank * per_rank + (size_t)i0 * W2 + k] = s_acc[k];
    }
    __threadfence_system();
    compute_bar();
    if (tid < xg.world) {
        st_release_sys(xg.flags[tid] + (size_t)xg.rank * n_groups + gidx, seq + 1u);
        const unsigned* mine = xg.flags[xg.rank] + (size_t)tid * n_groups + gidx;
        if ((int)(ld_acquire_sys(mine) - (seq + 1u)) < 0) {
            const long long t0 = clock64();
            unsigned polls = 0;
            while ((int)(ld_acquire_sys(mine) - (seq + 1u)) < 0) {
                if ((++polls & 1023u) == 0u && clock64() - t0 > XCHG_SPIN_LIMIT_CYCLES) {
                    if (error_flag) atomicExch(error_flag, 2);  // a missing peer must not hang the GPU
                    __trap();
                }
            }
        }
    }
    compute_bar();
    if (tid < cnt) {
        const float* src = xg.data[xg.rank] + slot_off + (size_t)i0 * W2 + tid;
        float acc = 0.f;
        for (int r = 0; r < xg.world; ++r) acc += __ldcg(src + (size_t)r * per_rank);
        s_acc[tid] = acc;
    }
    compute_bar();
}

// MT: theta_dim as a compile-time constant (0 = read it from the net).  With a runtime theta_dim every small loop of the kernel
// (layer-0 sample part, Q_j products, sums, tail) is a chain of predicated steps over MMAX = 10 slots; the instantiations for the
// benchmark tasks' dimensions unroll them exactly.
//
// PAIR: the CTA runs as one half of a CTA pair (cluster of 2, `cta_group::2`).  Both CTAs keep their own sample groups, operand
// images, tensor memory and compute / helper warps exactly as in the single-CTA kernel; what is shared is the tensor-core
// instruction stream: the leader (cluster rank 0) issues every tcgen05.mma over M = 256 rows (its own tile + the peer's), each CTA
// streams only HALF of every weight tile (rows [rank N/2, (rank+1) N/2) of the K-major image) into its ring.  Per SM that halves
// the bytes pulled from L2 and the B-operand reads of shared memory, the two things the MMA phases of the single-CTA kernel are
// bound by (profiles/r02_tc_notes.md).  Protocol: the peer's compute warps arrive on the LEADER's "peer" copies of bar_a /
// bar_c (remote mbarrier arrive), the peer's otherwise idle MMA warp forwards its ring's full barriers to the leader, and the
// leader's tcgen05.commit arrives on bar_empty / bar_dn of BOTH CTAs (multicast).  Needs the same number of items on both CTAs
// (the host only selects it for an even number of sample groups) and the plain fused mode (no export, no exchange, C == 1).
template <int MT, bool PAIR>
__global__ void __launch_bounds__(TC_THREADS, 1) traj_tc_kernel(const __grid_constant__ TrajParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sAhi = smem + OFF_AHI;
    uint8_t* sAlo = smem + OFF_ALO;
    float* sS = reinterpret_cast<float*>(smem + OFF_S);
    float* sQS = reinterpret_cast<float*>(smem + OFF_QS);
    float* sBase = reinterpret_cast<float*>(smem + OFF_BASE);
    float* sWout = reinterpret_cast<float*>(smem + OFF_WOUT);
    float* sQ = sBase + 1280;
    float* sA = sWout + (MMAX / 2) * 256;  // [m][256] layer-0 theta block, only when m <= MMAX / 2
    float* sAcc = reinterpret_cast<float*>(smem + OFF_ACC);
    float* sTheta = reinterpret_cast<float*>(smem + OFF_THETA);
    float* sNoise = reinterpret_cast<float*>(smem + OFF_NOISE);
    float* sPrior = reinterpret_cast<float*>(smem + OFF_PRIOR);
    float* sSched = reinterpret_cast<float*>(smem + OFF_SCHED);
    float* sMats = reinterpret_cast<float*>(smem + OFF_MATS);
    float* sBout = reinterpret_cast<float*>(smem + OFF_BOUT);
    float* sMinv = reinterpret_cast<float*>(smem + OFF_MINV);
    float* sG = reinterpret_cast<float*>(smem + OFF_G);
    uint8_t* sRowSi = smem + OFF_ROWS;
    uint8_t* sRowOj = smem + OFF_ROWS + 128;
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
    const uint32_t bar0 = smem_u32(smem + OFF_BAR);
    constexpr int NSTG = PAIR ? TC_NSTG_PAIR : TC_NSTG;                      // ring stages
    constexpr int STG_BYTES = PAIR ? TC_STAGE_BYTES / 2 : TC_STAGE_BYTES;    // bytes per stage
    // CTA pair: the ring's barriers and the leader's copies for the peer CTA's arrivals live in the spare half-stage
    const uint32_t barp0 = smem_u32(smem + OFF_RING + 7 * (TC_STAGE_BYTES / 2));
    auto bar_full = [&](int s) { return PAIR ? barp0 + 8u * s : bar0 + 8u * s; };
    auto bar_empty = [&](int s) { return PAIR ? barp0 + 8u * (NSTG + s) : bar0 + 8u * (TC_NSTG + s); };
    const uint32_t bar_a = bar0 + 8u * (2 * TC_NSTG);      // layer-0 A operand (and TMEM D) ready for the MMA warp
    // (handing layer 0 over in two K halves, like the mid epilogue below, measured -3 % .. -5 %: a second fence.proxy.async per
    // pass, and the early MMAs compete with the rest of layer 0 for the shared-memory pipe)
    // accumulator ready for the epilogue: two barriers used alternately, layer by layer (the MMA warp may complete two layers
    // -- the last layer of an item and the first layer of the next one -- before the compute warps look at the first)
    auto bar_dn = [&](unsigned n) { return bar0 + 8u * (2 * TC_NSTG + 1 + (n & 1u)); };
    // block cb of a mid epilogue is in tensor memory (every compute warp has converted its cb-th 16-column block): K-steps
    // {cq * nblk + cb, cq = 0..3} of the next layer may be issued
    auto bar_c = [&](int cb) { return bar0 + 8u * (2 * TC_NSTG + 3 + cb); };
    // CTA pair: the leader's copies of the barriers the peer CTA arrives on
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = crank == 0u;
    auto bar_fullp = [&](int s) { return barp0 + 8u * (2 * NSTG + s); };
    const uint32_t bar_ap = barp0 + 8u * (3 * NSTG);
    auto bar_cp = [&](int cb) { return barp0 + 8u * (3 * NSTG + 1 + cb); };
    // what the compute warps arrive on: their own CTA's barrier (leader, single-CTA kernel) or the leader's peer copy
    const uint32_t arr_a = (PAIR && !leader) ? mapa_shared(bar_ap, 0u) : bar_a;
    auto arrive_a = [&]() {
        if (PAIR && !leader) mbar_arrive_cluster(arr_a);
        else mbar_arrive(bar_a);
    };
    auto arrive_c = [&](int cb) {
        if (PAIR && !leader) mbar_arrive_cluster(mapa_shared(bar_cp(cb), 0u));
        else mbar_arrive(bar_c(cb));
    };
    // i-th K-step of a TS-form layer in issue order (K = 64 nblk): block index first, column quarter second
    auto ks_of = [](int i, int nblk) { return (i & 3) * nblk + (i >> 2); };

    const NetDev& net = p.net;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m = MT > 0 ? MT : net.m, L = net.L;
    const int H0 = net.Hp[0];
    const bool a_in_smem = (m <= MMAX / 2);
    // sBase holds SG <= 8 rows of 256 floats, and its first 1280 floats double as scratch of the scores phase; with
    // SG <= 5 the last 768 floats are free: enough for the inv(Sigma_j) of a short observation set (n = 30, m = 5: 750),
    // which saves the scores phase its dependent global loads.
    // (row stride odd: lanes = observations read Q_j[r][k] at the same (r, k), an even stride would be a 16-way bank conflict)
    const int qstr = (m * m) | 1;
    const bool q_in_smem = p.obs_prec != nullptr && p.C == 1 && p.SG <= 5 && p.n * qstr <= 768 && p.fin.ctx_samples == 0;
    // needs the aligned tile layout (row stride RS of a sample = power of two <= 32, chosen by the host when it costs no rows)
    const bool fused_sums = m <= 5 && (p.obs_prec == nullptr || q_in_smem) && p.fin.ctx_samples == 0 && p.C == 1 &&
                            p.RS <= 32 && (p.RS & (p.RS - 1)) == 0;
    const int n_groups = (p.N + p.SG - 1) / p.SG;
    const int mats_floats = 2 * m * m + m;
    constexpr int NB23 = TC_CT + 32;  // participants of named barriers 2 and 3

    if (tid == 0) {
        for (int s = 0; s < NSTG; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_empty(s), 1);
        }
        mbar_init(bar_a, TC_CW * 32);
        mbar_init(bar_dn(0), 1);
        mbar_init(bar_dn(1), 1);
        for (int cb = 0; cb < 4; ++cb) mbar_init(bar_c(cb), TC_CW * 32);
        if (PAIR) {
            for (int s = 0; s < NSTG; ++s) mbar_init(bar_fullp(s), 1);
            mbar_init(bar_ap, TC_CW * 32);
            for (int cb = 0; cb < 4; ++cb) mbar_init(bar_cp(cb), TC_CW * 32);
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
    // constant tables: output layer, its bias, row -> (sample, observation) maps, zeroed operand images
    for (int e = tid; e < 2 * TC_A_BYTES / 16; e += TC_THREADS) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0, 0, 0, 0);
    {
        const int Kout = net.Hp[L - 2];
        for (int e = tid; e < m * Kout; e += TC_THREADS) sWout[(e / Kout) * 256 + (e % Kout)] = __ldg(net.WoutT + e);
        if (a_in_smem)  // layer-0 theta block next to the output layer when both fit (theta_dim <= 5)
            for (int e = tid; e < m * H0; e += TC_THREADS) sA[(e / H0) * 256 + (e % H0)] = __ldg(net.A + e);
        if (tid < MMAX) sBout[tid] = (tid < m) ? __ldg(net.bout + tid) : 0.f;
        if (q_in_smem)  // inv(Sigma_j) of every observation in the tail of sBase (see q_in_smem)
            for (int e = tid; e < p.n * m * m; e += TC_THREADS) sQ[(e / (m * m)) * qstr + e % (m * m)] = __ldg(p.obs_prec + e);
        if (tid < TC_M) {
            const int si = tid / p.RS, oj = tid - si * p.RS;  // row = si * RS + oj (RS >= OC: rows oj >= OC of a sample are padding)
            const bool used = si < p.SG && oj < p.OC;
            sRowSi[tid] = (uint8_t)(used ? si : 255);
            sRowOj[tid] = (uint8_t)(used ? oj : 255);
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_d = *sTmem;

    const int n_steps = p.step_end - p.step_begin;
    int my_groups = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) ++my_groups;
    const long long tile_iters = (long long)my_groups * n_steps * p.C;  // tile evaluations by this CTA
    // Two sample groups of one CTA are interleaved step by step: the sums / tail of one group and the layer-0 sample part of
    // the other run while the tensor core works (needs the whole observation set in one tile and the in-kernel tail).
    const bool pairing = p.pairing && p.C == 1 && p.part_out == nullptr;

    // Register reallocation between the warpgroups (setmaxnreg): the CTA is launched with 96 registers per thread; the
    // four service warps give 32 of theirs back, the compute warps grow to 112.
    if (warp == TC_CW + 3) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");  // idle warp: completes the service warpgroup
    } else if (warp == TC_CW) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        // ================= TMA producer: streams the weight images in consumption order ===========================
        if (lane == 0) {
            unsigned it = 0;
            int tr_n = 0;
            bool tr_go = false;
            (void)tr_n; (void)tr_go;
            for (long long ti = 0; ti < tile_iters; ++ti) {
                tr_go = blockIdx.x == 0 && p.trace != nullptr && ti >= TRACE_FROM && ti < TRACE_TO;
                for (int l = 1; l <= L - 2; ++l) {
                    const int K = net.Hp[l - 1], Nn = net.Hp[l];
                    const uint32_t bytes = (uint32_t)Nn * 64u;
                    const uint8_t* src = net.wimg[l];
                    for (int i = 0; i < K / 16; ++i, ++it) {
                        const int ks = (l > 1) ? ks_of(i, K / 64) : i;  // the order the MMA warp consumes them in
                        const int s = it % NSTG;
                        const uint32_t ph = (it / NSTG) & 1u;
                        mbar_wait(bar_empty(s), ph ^ 1u, p.error_flag);
                        TRACE(60 + l);  // ring slot free = the MMAs of the K-step 4 stages back have completed
                        if (PAIR) {
                            // this CTA's half of the tile: rows [rank N/2, (rank+1) N/2) of the K-step, hi | lo, contiguous in
                            // the pair image
                            mbar_expect_tx(bar_full(s), bytes / 2);
                            tma_bulk_g2s(smem_u32(smem + OFF_RING + s * STG_BYTES),
                                         net.wimg2[l] + (size_t)ks * bytes + crank * (bytes / 2), bytes / 2, bar_full(s));
                        } else {
                            mbar_expect_tx(bar_full(s), bytes);
                            tma_bulk_g2s(smem_u32(smem + OFF_RING + s * TC_STAGE_BYTES), src + (size_t)ks * bytes, bytes,
                                         bar_full(s));
                        }
                    }
                }
            }
        }
    } else if (warp == TC_CW + 1) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        // ================= MMA issuer ===================================================================================
        if (PAIR && !leader) {
            // peer CTA of a pair: no instructions to issue.  Lane 0 forwards "ring stage full" to the leader, in the order the
            // leader consumes the stages (the leader's MMAs read this CTA's half of every weight tile).
            if (lane == 0) {
                unsigned it = 0;
                const uint32_t fwd0 = mapa_shared(bar_fullp(0), 0u);
                for (long long ti = 0; ti < tile_iters; ++ti)
                    for (int l = 1; l <= L - 2; ++l)
                        for (int i = 0; i < net.Hp[l - 1] / 16; ++i, ++it) {
                            const int s = it % NSTG;
                            mbar_wait(bar_full(s), (it / NSTG) & 1u, p.error_flag);
                            mbar_arrive_cluster(fwd0 + 8u * s);
                        }
            }
        } else
        if (lane == 0) {
            int tr_n = 0;
            bool tr_go = false;
            (void)tr_n; (void)tr_go;
            unsigned it = 0, a_cnt = 0, lc = 0, c_par = 0;  // c_par: bit cb = phase parity of bar_c(cb) to wait for next
            const uint32_t a_hi0 = smem_u32(sAhi), a_lo0 = smem_u32(sAlo);
            for (long long ti = 0; ti < tile_iters; ++ti) {
                tr_go = blockIdx.x == 0 && p.trace != nullptr && ti >= TRACE_FROM && ti < TRACE_TO;
                for (int l = 1; l <= L - 2; ++l) {
                    const int K = net.Hp[l - 1], Nn = net.Hp[l];
                    const uint32_t idesc = umma_idesc(Nn, PAIR ? 256 : TC_M);
                    const int Nb = PAIR ? Nn / 2 : Nn;                        // B rows held by this CTA
                    const uint32_t b_lbo = (uint32_t)Nb * 16u, b_sbo = 128u;  // K halves Nb*16 B apart, 8-row groups 128 B
                    const uint32_t a_lbo = 2048u, a_sbo = 128u;
                    if (l == 1)
                    {
                        mbar_wait(bar_a, a_cnt & 1u, p.error_flag);
                        if (PAIR) mbar_wait(bar_ap, a_cnt & 1u, p.error_flag);  // the peer's layer 0
                        ++a_cnt;
                    }
                    TRACE(50 + l);
                    // early layer 0: this layer's accumulator is the region the previous (TS-form) layer reads its A operand
                    // from; wait until that layer has completed before overwriting it
                    if (l == 1 && lc > 0 && L - 2 >= 2) mbar_wait(bar_dn(lc - 1u), ((lc - 1u) >> 1) & 1u, p.error_flag);
                    tc_fence_after();
                    const uint32_t dacc = tmem_d + ((lc & 1u) ? 256u : 0u);  // the other accumulator may still be read by an epilogue
                    const uint32_t dprev = tmem_d + ((lc & 1u) ? 0u : 256u);  // accumulator of the previous layer (TS form: its A operand)
                    (void)dprev;
                    const long long mma_t0 = (p.prof && blockIdx.x == 0) ? clock64() : 0;
                    ++lc;
                    for (int i = 0; i < K / 16; ++i, ++it) {
                        const int ks = (l > 1) ? ks_of(i, K / 64) : i;
                        if (l > 1 && (i & 3) == 0) {  // block i / 4 of the previous layer's epilogue is in tensor memory
                            mbar_wait(bar_c(i >> 2), (c_par >> (i >> 2)) & 1u, p.error_flag);
                            if (PAIR) mbar_wait(bar_cp(i >> 2), (c_par >> (i >> 2)) & 1u, p.error_flag);
                            c_par ^= 1u << (i >> 2);  // (layers of different widths use different numbers of block barriers)
                            tc_fence_after();
                        }
                        const int s = it % NSTG;
                        const uint32_t ph = (it / NSTG) & 1u;
                        mbar_wait(bar_full(s), ph, p.error_flag);
                        if (PAIR) mbar_wait(bar_fullp(s), ph, p.error_flag);  // the peer's half of the tile
                        tc_fence_after();
                        const uint32_t wb = smem_u32(smem + OFF_RING + s * STG_BYTES);
                        const uint64_t b_hi = umma_desc(wb, b_lbo, b_sbo);
                        const uint64_t b_lo = umma_desc(wb + (uint32_t)Nb * 32u, b_lbo, b_sbo);
                        if (l > 1) {  // A = the previous layer's accumulator region, converted in place by the epilogue
                            const uint32_t a_t = dprev + (uint32_t)ks * 16u;  // K-step ks: hi in columns 16 ks .. +7, lo in +8 .. +15
                            if (PAIR) {
                                tc_mma2_bf16_ts(dacc, a_t, b_hi, idesc, i > 0 ? 1u : 0u);
                                tc_mma2_bf16_ts(dacc, a_t + 8u, b_hi, idesc, 1u);
                                tc_mma2_bf16_ts(dacc, a_t, b_lo, idesc, 1u);
                            } else {
                                tc_mma_bf16_ts(dacc, a_t, b_hi, idesc, i > 0 ? 1u : 0u);
                                tc_mma_bf16_ts(dacc, a_t + 8u, b_hi, idesc, 1u);
                                tc_mma_bf16_ts(dacc, a_t, b_lo, idesc, 1u);
                            }
                        } else
                        {
                            const uint64_t a_hi = umma_desc(a_hi0 + ks * 4096, a_lbo, a_sbo);
                            const uint64_t a_lo = umma_desc(a_lo0 + ks * 4096, a_lbo, a_sbo);
                            if (PAIR) {
                                tc_mma2_bf16(dacc, a_hi, b_hi, idesc, i > 0 ? 1u : 0u);
                                tc_mma2_bf16(dacc, a_lo, b_hi, idesc, 1u);
                                tc_mma2_bf16(dacc, a_hi, b_lo, idesc, 1u);
                            } else {
                                tc_mma_bf16(dacc, a_hi, b_hi, idesc, i > 0 ? 1u : 0u);
                                tc_mma_bf16(dacc, a_lo, b_hi, idesc, 1u);
                                tc_mma_bf16(dacc, a_hi, b_lo, idesc, 1u);
                            }
                        }
                        // ring slot is free once these MMAs have read it (pair: the slot of both CTAs)
                        if (PAIR) tc_commit2(bar_empty(s));
                        else tc_commit(bar_empty(s));
                        if (i == 0) TRACE(30 + l);
                    }
                    TRACE(40 + l);
                    // accumulator complete (pair: in the tensor memory of both CTAs)
                    if (PAIR) tc_commit2(bar_dn(lc - 1u));
                    else tc_commit(bar_dn(lc - 1u));
                    if (p.prof && blockIdx.x == 0) {  // profiling: duration of this layer's MMAs (slots 8 / 9: first / last layer)
                        mbar_wait(bar_dn(lc - 1u), ((lc - 1u) >> 1) & 1u, p.error_flag);
                        p.prof[(l == L - 2) ? 9 : 8] += clock64() - mma_t0;
                    }
                }
            }
        }
    } else if (warp == TC_CW + 2) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        // ================= scalar helper: everything that only needs theta at the start of the step =================
        const bool uni = (p.fin.prior_kind == D4S_PRIOR_UNIFORM);
        const int nq = (m + 3) / 4;
        for (int gk = 0; gk < my_groups;) {
            const int ngr = (pairing && gk + 1 < my_groups) ? 2 : 1;  // groups interleaved step by step (see compute warps)
            const int total = ngr * n_steps;
            for (int it = 0; it < total; ++it) {
                const int gi = (ngr == 2) ? (it & 1) : 0;
                const int step = p.step_begin + ((ngr == 2) ? (it >> 1) : it);
                const int i0 = (blockIdx.x + (gk + gi) * (int)gridDim.x) * p.SG;
                const int nsamp = min(p.SG, p.N - i0);
                const float* th = sTheta + gi * TC_SG_MAX * MMAX;
                bar_sync(2, NB23);  // theta of this item is final; the previous item's consumers are done with our buffers
                if (p.part_out) {  // export mode: the caller reduces over ranks and finalizes; nothing to prepare
                    bar_arrive(3, NB23);
                    continue;
                }
                const float* sched = p.sched + (size_t)step * D4S_SCHED_STRIDE;
                if (lane < D4S_SCHED_STRIDE) sSched[lane] = __ldg(sched + lane);
                const float* mrow = p.mats + (size_t)step * mats_floats;
                for (int e = lane; e < mats_floats; e += 32) sMats[e] = __ldg(mrow + e);
                __syncwarp();
                const int combine = (int)sSched[D4S_S_COMBINE];
                const double omn = (double)sSched[D4S_S_ONE_MINUS_N], aos2 = (double)sSched[D4S_S_A_OVER_S2];
                const bool per_sample = (combine == D4S_COMBINE_PER_SAMPLE);
                if (uni) {  // erf/exp of the diffused-uniform score in fp64: one lane per (sample, dim)
                    const double sa = (double)sSched[D4S_S_SQRT_ALPHA], sg = (double)sSched[D4S_S_SIGMA];
                    const double inv_sa = 1.0 / sa, inv_sg = 1.0 / sg;
                    for (int e = lane; e < nsamp * m; e += 32) {
                        const int si = e / m, k = e - si * m;
                        double s0, jac;
                        uniform_prior_terms_recip((double)th[si * MMAX + k], (double)__ldg(p.fin.low + k),
                                                  (double)__ldg(p.fin.high + k), sa, sg, inv_sa, inv_sg, s0, jac);
                        if (per_sample) {
                            const double p0 = aos2 / (1.0 + sg * sg * jac);  // inv(sigma^2/alpha (1 + sigma^2 J0))
                            sPrior[si * MMAX + k] = (float)p0;
                            sG[si * MMAX + k] = (float)(omn * p0 * s0);
                        } else {
                            sG[si * MMAX + k] = (float)(omn * s0);
                        }
                    }
                } else if (p.fin.prior_kind == D4S_PRIOR_GAUSSIAN) {  // g = G (theta - mu_t), G from the step's matrix row
                    const float* G = sMats + m * m;
                    const float* mu = sMats + 2 * m * m;
                    for (int e = lane; e < nsamp * m; e += 32) {
                        const int si = e / m, r = e - si * m;
                        double acc = 0.0;
                        for (int c = 0; c < m; ++c) acc = fma((double)G[r * m + c], (double)th[si * MMAX + c] - (double)mu[c], acc);
                        sG[si * MMAX + r] = (float)acc;
                    }
                } else {
                    for (int e = lane; e < TC_SG_MAX * MMAX; e += 32) sG[e] = 0.f;
                }
                __syncwarp();
                if (per_sample) {
                    if (!uni)  // no per-sample prior precision: invert the step's Lambda as it is
                        for (int e = lane; e < nsamp * MMAX; e += 32) sPrior[e] = 0.f;
                    __syncwarp();
                    for (int e = lane; e < nsamp * m; e += 32) {  // one lane per (sample, column of the inverse)
                        const int si = e / m, c = e - si * m;
                        const float* qsum = (p.fin.ctx_qsum && p.fin.ctx_samples > 0)
                                                ? p.fin.ctx_qsum + (size_t)((i0 + si) / p.fin.ctx_samples) * m * m : nullptr;
                        lambda_inverse_dispatch(m, sMats, qsum, sPrior + si * MMAX, uni ? omn : 0.0, c, sMinv + si * MMAX * MMAX);
                    }
                }
                for (int e = lane; e < nsamp * nq; e += 32) {  // DDIM noise: one lane per (sample, 4 dims)
                    const int si = e / nq, qd = e - si * nq;
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
                __syncwarp();
                bar_arrive(3, NB23);
            }
            gk += ngr;
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ================= compute warps ==================================================================================
        const int q = warp & 3, cq = warp >> 2;  // TMEM lane quarter, column quarter
        const int row = 32 * q + lane;           // the TMEM lane / tile row this thread owns in the epilogues
        unsigned d_cnt = 0;
        const bool prof_on = (p.prof != nullptr) && blockIdx.x == 0 && tid == 0;
        int tr_n = 0;
        bool tr_go = false;
        (void)tr_n; (void)tr_go;
        long long prof_t = prof_on ? clock64() : 0;
        const int nkg = H0 / 8;
        const int bcol = tid & 255, bhalf = tid >> 8;  // layer-0 sample part: column, and which half of the samples
        float tf_next = (bcol < H0 && n_steps > 0) ? __ldg(p.time_feat + (size_t)p.step_begin * H0 + bcol) : 0.f;
        // ---- phases shared by the two schedules (one group after the other / two groups interleaved) ----------------------
        // layer-0 per-sample part of an item: base[si][col] = time_feat[col] + sum_k A[k][col] theta[si][k]
        auto do_base = [&](const float* th, bool fresh, int step, int next_step) {
            if (net.emb_L > 0) {  // nets with a theta embedding MLP (JR-NMM): separate, non-inlined code path
                float t0e = 0.f;
                if (tid < H0) {
                    if (fresh) {
                        t0e = tf_next;
                        tf_next = __ldg(p.time_feat + (size_t)next_step * H0 + tid);
                    } else {
                        t0e = __ldg(p.time_feat + (size_t)step * H0 + tid);
                    }
                }
                theta_embed_base(net, th, sS, sQS, sBase, p.SG, H0, t0e, tid);
                return;
            }
            if (bcol < H0) {  // all 16 compute warps: thread = (column, half of the group's samples)
                float t0;
                if (fresh) {
                    t0 = tf_next;
                    tf_next = __ldg(p.time_feat + (size_t)next_step * H0 + bcol);
                } else {
                    t0 = __ldg(p.time_feat + (size_t)step * H0 + bcol);
                }
                float ak[MMAX];
#pragma unroll
                for (int k = 0; k < MMAX; ++k)
                    ak[k] = (k < m) ? (a_in_smem ? sA[k * 256 + bcol] : __ldg(net.A + (size_t)k * H0 + bcol)) : 0.f;
                for (int si = bhalf; si < p.SG; si += 2) {
                    float z = t0;
#pragma unroll
                    for (int k = 0; k < MMAX; ++k)
                        if (k < m) z = fmaf(ak[k], th[si * MMAX + k], z);
                    sBase[si * 256 + bcol] = z;
                }
            }
        };
        // per-sample sums over the chunk's observations.  item = (sample si, component c < 2m): S1[c] = sum_j s_ij[c]
        // (c < m), S2[c-m] = sum_j (inv(Sigma_j) s_ij)[c-m] (rows of sQS); a warp takes one item at a time, its lanes run over the
        // observations, one butterfly per item (fixed order => deterministic).
        auto do_sums = [&](int nsamp, int j0, int nobs, bool first_chunk) {
            for (int item = warp; item < nsamp * 2 * m; item += TC_CW) {
                const int si = item / (2 * m), c = item - si * 2 * m;
                const float* src = (c < m) ? sS + c : sQS + (c - m);
                float acc = 0.f;
                if (c < m || p.obs_prec) {
                    for (int oj = lane; oj < nobs; oj += 32) {
                        const float wj = p.obs_weight ? __ldg(p.obs_weight + j0 + oj) : 1.f;
                        acc = fmaf(wj, src[(si * p.RS + oj) * MMAX], acc);
                    }
                }
#pragma unroll
                for (int sh = 16; sh > 0; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
                if (lane == 0) sAcc[si * 2 * m + c] = (first_chunk ? 0.f : sAcc[si * 2 * m + c]) + acc;
            }
        };
        // tail of an item: prior, Lambda solve, DDIM update (theta stays in smem); warp -> sample, lane -> theta component
        auto do_tail = [&](float* th, int nsamp) {
            if (warp < nsamp) {
                const int si = warp, r = lane;
                const bool act = lane < m;
                const int combine = (int)sSched[D4S_S_COMBINE];
                const float* acc = sAcc + si * 2 * m;
                float out = 0.f;
                if (act) {
                    if (combine == D4S_COMBINE_SUM) {
                        out = sG[si * MMAX + r] + sSched[D4S_S_S1_WEIGHT] * acc[r];  // (1-n) s0 + w sum_j s_ij (nse.py:642)
                    } else {
                        const float aos2 = sSched[D4S_S_A_OVER_S2];
                        const float* Minv = (combine == D4S_COMBINE_SHARED) ? sMats : sMinv + si * MMAX * MMAX;
                        for (int c = 0; c < m; ++c) {
                            const float w = fmaf(aos2, acc[c], acc[m + c]) + sG[si * MMAX + c];  // nse.py:631-636 / 647-652
                            out = fmaf(Minv[r * m + c], w, out);  // Lambda^-1 w   (nse.py:638 / 654)
                        }
                    }
                }
                float c_score = sSched[D4S_S_C_SCORE];
                const float tame = sSched[D4S_S_TAME];
                if (tame > 0.f) {  // tamed ULA (nse.py:336); the whole warp takes part, lanes >= m contribute 0
                    float nrm = out * out;
#pragma unroll
                    for (int sh = 16; sh > 0; sh >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, sh);
                    c_score = tame / (1.f + tame * sqrtf(nrm));
                }
                if (act) {
                    float v = sSched[D4S_S_C_THETA] * th[si * MMAX + r] + c_score * out +
                              sSched[D4S_S_BRIDGE_STD] * sNoise[si * MMAX + r];  // nse.py:289-298
                    if (p.fin.clip_lo <= p.fin.clip_hi && sSched[D4S_S_CLIP] != 0.f)
                        v = fminf(fmaxf(v, p.fin.clip_lo), p.fin.clip_hi);
                    th[si * MMAX + r] = v;
                }
            }
        };

        // layer 0 of a pass: act0[r][col] = relu(base[si][col] + obs_feat[j][col]) -> bf16 hi/lo A images, then launch layer 1
        // obs_feat rows of this thread's first layer-0 iteration (lane = observation, k-groups warp and warp + 16)
        auto l0_load = [&](float4 (&u)[2][2], int j0, int nobs, int oj) {
            const bool ovalid = oj < nobs;
            const float* up = p.obs_feat + (size_t)(j0 + (ovalid ? oj : 0)) * H0;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int kg = warp + 16 * j;
                if (ovalid && kg < nkg) {
                    u[j][0] = __ldg(reinterpret_cast<const float4*>(up + 8 * kg));
                    u[j][1] = __ldg(reinterpret_cast<const float4*>(up + 8 * kg + 4));
                } else {
                    u[j][0] = u[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        };
        auto do_layer0 = [&](int i0, int nsamp, int j0, int nobs, float4 (&upre)[2][2], bool have_pre) {
            if (p.fin.ctx_samples > 0) {
                layer0_contexts(p.obs_feat, sBase, sAhi, sAlo, H0, p.OC, p.SG, nsamp, p.n, j0, nobs, p.fin.ctx_samples, i0, warp, lane);  // (RS == OC there)
                fence_proxy_async();
                tc_fence_before();
                arrive_a();
                return;
            }
            // ---- layer 0: act0[r][col] = relu(base[si][col] + obs_feat[j][col]) -> bf16 hi/lo A images -----------
            // warp w owns k-groups w, w+16; lanes run over the observations: one 32-byte read of obs_feat
            // serves the rows of all SG samples, and every global load is issued before its first use.
            for (int oj = lane; oj < p.OC; oj += 32) {
                const bool ovalid = oj < nobs;
                float4 u[2][2];
                if (have_pre && oj == lane) {  // loaded before the wait for the tensor core
#pragma unroll
                    for (int j = 0; j < 2; ++j) { u[j][0] = upre[j][0]; u[j][1] = upre[j][1]; }
                } else {
                    l0_load(u, j0, nobs, oj);
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int kg = warp + 16 * j;
                    if (kg < nkg) {
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
                            store_split8(v, sAhi, sAlo, a_off(si * p.RS + oj, kg));
                        }
                    }
                }
            }
            fence_proxy_async();
            tc_fence_before();
            arrive_a();
        };
        // one 16-column block of an accumulator row: bias + ReLU
        auto ld_act = [&](uint32_t dacc, int col0, const float4 (&bb)[4], float (&a)[16]) {
            uint32_t d[16];
            tc_ld16(dacc + ((uint32_t)(32 * q) << 16) + (uint32_t)col0, d);
            tc_wait_ld();
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                a[4 * e] = fmaxf(__uint_as_float(d[4 * e]) + bb[e].x, 0.f);
                a[4 * e + 1] = fmaxf(__uint_as_float(d[4 * e + 1]) + bb[e].y, 0.f);
                a[4 * e + 2] = fmaxf(__uint_as_float(d[4 * e + 2]) + bb[e].z, 0.f);
                a[4 * e + 3] = fmaxf(__uint_as_float(d[4 * e + 3]) + bb[e].w, 0.f);
            }
        };
        // epilogue of the last hidden layer + output layer in fp32 + scores of a pass -> sS
        auto do_last_epilogue = [&](uint32_t dacc, float inv_sigma, int j0, int nobs, int i0, int nsamp, bool first_chunk) {
            const int Nn = net.Hp[L - 2];
            const int qcols = Nn >> 2, nblk = qcols / 16;
            const float* bias_l = net.bias[L - 2] + cq * qcols;
            float po[MMAX];  // output-layer partial dot products of this thread's row / column quarter
            if constexpr (MT == 0 || MT > MMAX / 2) {
                // wide theta (and the generic instantiation): plain FMAs, five outputs at a time -- the packed / pipelined form
                // below needs 2 x MMAX accumulators + two weight sets and spills at the kernel's 96 registers (stress shape,
                // theta_dim = 10: +6 % per step, profiles/r02_tc_notes.md)
#pragma unroll
                for (int o = 0; o < MMAX; ++o) po[o] = 0.f;
                for (int cb = 0; cb < nblk; ++cb) {
                    const int col0 = cq * qcols + 16 * cb;
                    float4 bb[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) bb[e] = __ldg(reinterpret_cast<const float4*>(bias_l + 16 * cb) + e);
                    float a[16];
                    ld_act(dacc, col0, bb, a);
#pragma unroll
                    for (int oh = 0; oh < MMAX; oh += 5) {
                        if (oh < m) {
#pragma unroll
                            for (int e = 0; e < 16; e += 4) {
                                float4 w4[5];
#pragma unroll
                                for (int o = 0; o < 5; ++o)
                                    w4[o] = *reinterpret_cast<const float4*>(sWout + (oh + o) * 256 + col0 + e);  // rows >= m: unused, in bounds
#pragma unroll
                                for (int o = 0; o < 5; ++o) {
                                    float acc = po[oh + o];
                                    acc = fmaf(a[e], w4[o].x, acc);
                                    acc = fmaf(a[e + 1], w4[o].y, acc);
                                    acc = fmaf(a[e + 2], w4[o].z, acc);
                                    acc = fmaf(a[e + 3], w4[o].w, acc);
                                    po[oh + o] = acc;
                                }
                            }
                        }
                    }
                }
            } else {
                // output-layer partial dot products of this thread's row / column quarter, as packed pairs (even | odd columns): the
                // products run as FFMA2 over column pairs, po[o] = po2[o].x + po2[o].y at the end
                float2 po2[MMAX];
#pragma unroll
                for (int o = 0; o < MMAX; ++o) po2[o] = make_float2(0.f, 0.f);
                auto load_w = [&](float4 (&w)[4], int o, int col0) {  // Wout[o][col0 .. col0 + 15] (smem, broadcast reads)
#pragma unroll
                    for (int e = 0; e < 4; ++e) w[e] = *reinterpret_cast<const float4*>(sWout + o * 256 + col0 + 4 * e);
                };
                for (int cb = 0; cb < nblk; ++cb) {
                    const int col0 = cq * qcols + 16 * cb;
                    float4 bb[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) bb[e] = __ldg(reinterpret_cast<const float4*>(bias_l + 16 * cb) + e);
                    // The weight loads of output o + 1 are issued before the products of output o, those of output 0 before the wait for
                    // the accumulator block: a shared-memory load issued right before its use waited ~40 % of this phase's cycles (the
                    // MMA operand reads and the TMA writes of the other item's layer keep the shared-memory pipe busy).
                    uint32_t d[16];
                    tc_ld16(dacc + ((uint32_t)(32 * q) << 16) + (uint32_t)col0, d);
                    float4 wn[4];
                    load_w(wn, 0, col0);
                    tc_wait_ld();
                    float2 a2[8];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        a2[2 * e] = make_float2(fmaxf(__uint_as_float(d[4 * e]) + bb[e].x, 0.f), fmaxf(__uint_as_float(d[4 * e + 1]) + bb[e].y, 0.f));
                        a2[2 * e + 1] = make_float2(fmaxf(__uint_as_float(d[4 * e + 2]) + bb[e].z, 0.f), fmaxf(__uint_as_float(d[4 * e + 3]) + bb[e].w, 0.f));
                    }
#pragma unroll
                    for (int o = 0; o < MMAX; ++o) {
                        if (o < m) {
                            float4 wc[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) wc[e] = wn[e];
                            if (o + 1 < m) load_w(wn, o + 1, col0);
                            float2 acc = po2[o];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                acc = __ffma2_rn(a2[2 * e], make_float2(wc[e].x, wc[e].y), acc);
                                acc = __ffma2_rn(a2[2 * e + 1], make_float2(wc[e].z, wc[e].w), acc);
                            }
                            po2[o] = acc;
                        }
                    }
                }
#pragma unroll
                for (int o = 0; o < MMAX; ++o) po[o] = po2[o].x + po2[o].y;
            }
            PROF_MARK(4);  // epilogues
            if (fused_sums) {
                // The sums over the observations are LINEAR in the output-layer partials: S1 = -1/sigma sum_rows sum_quarters (po + b),
                // S2 = -1/sigma sum_rows Q_j sum_quarters (po + b).  Every thread applies inv(Sigma_j) to its own partial (bias added by
                // quarter 0), the lanes of one sample are reduced with shuffles, 10 floats per (warp, sample) are parked and a
                // fixed-order sum over the column quarters finishes -- no per-row scores in shared memory, one barrier instead of four.
                const int si = sRowSi[row], oj = sRowOj[row];
                const bool in_tile = si != 255;
                float val[10];
#pragma unroll
                for (int e = 0; e < 10; ++e) val[e] = 0.f;
                if (in_tile && oj < nobs) {
                    const float wj = p.obs_weight ? __ldg(p.obs_weight + j0 + oj) : 1.f;
#pragma unroll
                    for (int o = 0; o < 5; ++o)
                        if (o < m) val[o] = (po[o] + (cq == 0 ? sBout[o] : 0.f)) * wj;
                    if (p.obs_prec) {
                        const float* Q = sQ + oj * qstr;
#pragma unroll
                        for (int r = 0; r < 5; ++r) {
                            if (r < m) {
                                float acc = 0.f;
#pragma unroll
                                for (int k = 0; k < 5; ++k)
                                    if (k < m) acc = fmaf(Q[r * m + k], val[k], acc);
                                val[5 + r] = acc;
                            }
                        }
                    }
                }
                // Sum over the RS lanes of a sample (an aligned power-of-two segment of the warp; padding rows hold zeros) by recursive
                // halving: at distance d the lower lanes keep the lower half of the value slots and receive the partner's, the upper
                // lanes the upper half -- 8 + 4 + 2 + 1 (+ 1) shuffles for 16 slots (10 used) instead of 10 per tree level, and the
                // order depends only on a row's offset inside its segment (=> independent of the sample's slot in the tile).
                float v[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = e < 10 ? val[e] : 0.f;
                int prefix = 0, cnt = 16;  // this lane ends up with slots prefix * cnt .. prefix * cnt + cnt - 1
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int dist = p.RS >> (k + 1);
                    if (dist > 0) {
                        const bool upper = (lane & dist) != 0;
#pragma unroll
                        for (int i = 0; i < (8 >> k); ++i) {
                            const float send = upper ? v[i] : v[(8 >> k) + i];
                            const float keep = upper ? v[(8 >> k) + i] : v[i];
                            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, dist);
                        }
                        prefix = 2 * prefix + (upper ? 1 : 0);
                        cnt = 8 >> k;
                    }
                }
                if (p.RS == 32) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);  // (one slot per lane pair left)
                // (the lane that ends up with a slot may be a padding row of its sample: what counts is the segment, not the lane)
                const int seg = lane / p.RS;  // segment of this lane inside the warp = sample (32 q) / RS + seg of the tile
                if ((32 * q) / p.RS + seg < p.SG && (p.RS < 32 || (lane & 1) == 0)) {  // lane quarter q, column quarter cq
                    float* dst = sS + (((q * 4 + cq) * TC_SG_MAX) + seg) * 10;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int e = prefix * cnt + j;
                        if (j < cnt && e < 10) dst[e] = v[j];
                    }
                }
                tc_fence_before();
                PROF_MARK(11);  // per-thread contributions + segmented shuffle tree (the barrier wait below stays in slot 5)
                compute_bar();
                if (tid < nsamp * 2 * m) {
                    const int s_i = tid / (2 * m), c = tid - s_i * 2 * m;
                    const int e = c < m ? c : 5 + (c - m);
                    // (RS is a power of two <= 32: the rows of a sample sit in ONE lane quarter, at a fixed offset pattern, so
                    // the summation order of a sample does not depend on its slot in the tile => sample sharding stays bit-exact)
                    const int qq = (s_i * p.RS) >> 5;
                    const float* src = sS + ((qq * 4 * TC_SG_MAX) + (s_i - (32 * qq) / p.RS)) * 10 + e;
                    float acc = 0.f;
#pragma unroll
                    for (int c2 = 0; c2 < 4; ++c2) acc += src[c2 * TC_SG_MAX * 10];
                    sAcc[tid] = (first_chunk ? 0.f : sAcc[tid]) - inv_sigma * acc;
                }
                PROF_MARK(5);
                return;
            }
            // ---- scores: the column quarters 1..3 park their partial dot products in sS / sQS / sBase (free here), quarter 0
            // adds them up as (q0 + q1) + (q2 + q3), s = -(eps + b) / sigma, and applies inv(Sigma_j) to its row's score.
            // (when layer 0 of the next pass ran just before this epilogue, other warps may still be reading sBase there:
            // nothing else orders them against the scratch writes below -- with one short tensor-core layer this raced)
            compute_bar();
            if (cq != 0) {
                float* dst = (cq == 1) ? sS : (cq == 2) ? sQS : sBase;
#pragma unroll
                for (int o = 0; o < MMAX; ++o)
                    if (o < m) dst[row * MMAX + o] = po[o];
            }
            tc_fence_before();
            compute_bar();
            if (cq == 0) {
                float s[MMAX];
#pragma unroll
                for (int o = 0; o < MMAX; ++o) {
                    s[o] = 0.f;
                    if (o < m) {
                        const float v = (po[o] + sS[row * MMAX + o]) + (sQS[row * MMAX + o] + sBase[row * MMAX + o]);
                        s[o] = -(v + sBout[o]) * inv_sigma;
                        sS[row * MMAX + o] = s[o];
                    }
                }
                const int oj = sRowOj[row];
                if (p.obs_prec && oj < nobs) {  // (rows past the tile's pairs carry oj = 255)
                    if (q_in_smem) {
                        const float* Q = sQ + oj * qstr;
                        for (int r = 0; r < m; ++r) {
                            float qs = 0.f;
#pragma unroll
                            for (int k = 0; k < MMAX; ++k)
                                if (k < m) qs = fmaf(Q[r * m + k], s[k], qs);
                            sQS[row * MMAX + r] = qs;
                        }
                    } else {
                        const int ctx0 = p.fin.ctx_samples > 0 ? ((i0 + sRowSi[row]) / p.fin.ctx_samples) * p.n : 0;  // batched contexts
                        const float* Q = p.obs_prec + (size_t)(ctx0 + j0 + oj) * m * m;
                        for (int r = 0; r < m; ++r) {
                            float qs = 0.f;
#pragma unroll
                            for (int k = 0; k < MMAX; ++k)
                                if (k < m) qs = fmaf(__ldg(Q + r * m + k), s[k], qs);
                            sQS[row * MMAX + r] = qs;
                        }
                    }
                }
            }
            compute_bar();
            PROF_MARK(5);  // scores
        };

        const int Lh = L - 2;  // hidden layers on the tensor core
        // Layer 0 reads, per thread, 16 floats of obs_feat (its observation = lane, its two k-groups): with the whole observation set
        // in one tile of <= 32 observations per sample they are the SAME 16 floats for every item of the trajectory -> loaded once.
        // (Per-item loads were 30 KB per item through a 28 KB L1 that also has to hold the local-memory lines.)
        // (instantiations for theta_dim <= 5 only: the generic / wide ones are short of registers, 16 more live values cost the tall
        // stress shape 9 % per step)
        const bool obs_in_regs = (MT != 0 && MT <= MMAX / 2) && p.C == 1 && p.OC <= 32 && p.fin.ctx_samples == 0;
        float4 upre[2][2];
        if (obs_in_regs) l0_load(upre, 0, min(p.OC, p.n), lane);
        for (int gk = 0; gk < my_groups;) {
            const int ngr = (pairing && gk + 1 < my_groups) ? 2 : 1;
            // Two groups interleaved: the last epilogue / scores of a pass run while the tensor core is on layer 1 of the
            // next pass (other accumulator), its sums / tail and the layer-0 sample part of the pass after next while
            // the tensor core is on the last layer.  One group: the same loop, everything in program order.
            const bool defer = (ngr == 2);
            const int total = ngr * n_steps;
            const int total_p = total * p.C;  // passes (defer => C == 1)
            // (scalars + selects, not arrays indexed by the group: those would live in local memory, and a local-memory miss
            // costs an L2 round trip here -- the L1 is almost entirely carved out as shared memory)
            const int i0g0 = (blockIdx.x + gk * (int)gridDim.x) * p.SG;
            const int i0g1 = (blockIdx.x + (gk + ngr - 1) * (int)gridDim.x) * p.SG;
            const int nsg0 = min(p.SG, p.N - i0g0), nsg1 = min(p.SG, p.N - i0g1);
            auto i0g = [&](int gi) { return gi ? i0g1 : i0g0; };
            auto nsg = [&](int gi) { return gi ? nsg1 : nsg0; };
            for (int e = tid; e < ngr * TC_SG_MAX * MMAX; e += TC_CT) {
                const int gi = e / (TC_SG_MAX * MMAX), r = e - gi * TC_SG_MAX * MMAX;
                const int si = r / MMAX, k = r - si * MMAX;
                sTheta[e] = (si < nsg(gi) && k < m) ? p.theta[(size_t)(i0g(gi) + si) * m + k] : 0.f;
            }
            compute_bar();
            if (total > 0) bar_arrive(2, NB23);  // theta of item 0 is final: the helper warp may start on it
            auto step_of = [&](int item) { return (item < total) ? p.step_begin + (defer ? (item >> 1) : item) : p.step_begin; };
            if (defer && total > 0) {  // (later items: inside the previous item's last tensor-core wait)
                do_base(sTheta, true, 0, step_of(1));
                compute_bar();
            }
            for (int ps = 0; ps <= total_p; ++ps) {  // one extra iteration finishes the last pass
                tr_go = blockIdx.x == 0 && p.trace != nullptr && ps >= TRACE_FROM && ps < TRACE_TO;
                TRACE(20);
                const bool have = ps < total_p, prev = ps > 0;
                const int it = ps / p.C, chunk = ps - it * p.C;
                const int gi = defer ? (it & 1) : 0;
                const int j0 = chunk * p.OC, nobs = min(p.OC, p.n - j0);
                const int pit = prev ? (ps - 1) / p.C : 0, pchunk = prev ? (ps - 1) - pit * p.C : 0;
                const int pgi = defer ? (pit & 1) : 0;
                const int pj0 = pchunk * p.OC, pnobs = min(p.OC, p.n - pj0);
                float* th = sTheta + gi * TC_SG_MAX * MMAX;
                float* pth = sTheta + pgi * TC_SG_MAX * MMAX;
                float4 bcur[4];
                float inv_sigma_prev = 0.f;
                uint32_t dacc_prev = 0;
                // layer 0 of this pass goes first when it does not depend on the previous pass: interleaved groups, or a
                // later observation chunk of the same item (its sample part was recomputed at the end of the last pass)
                const bool early = have && (defer || chunk > 0);
                const bool have_pre = !obs_in_regs && early && (ps == 0 || Lh < 2) && lane < p.OC && p.fin.ctx_samples == 0;
                if (have_pre) l0_load(upre, j0, nobs, lane);  // layer 0 follows the wait directly: hide its global loads
                if (prev) {  // the last layer of the previous pass: once it is done the operand images are free again
                    inv_sigma_prev = __ldg(p.sched + (size_t)step_of(pit) * D4S_SCHED_STRIDE + D4S_S_INV_SIGMA);
                    dacc_prev = tmem_d + ((d_cnt & 1u) ? 256u : 0u);
                    TRACE(21);
                    mbar_wait(bar_dn(d_cnt), (d_cnt >> 1) & 1u, p.error_flag);
                    ++d_cnt;
                    tc_fence_after();
                    PROF_MARK(3);  // waiting for the tensor core
                }
                const bool l0_early = Lh >= 2;  // the operand image in shared memory is free once layer 1 has completed
                if (early && (ps == 0 || !l0_early)) {
                    do_layer0(i0g(gi), nsg(gi), j0, nobs, upre, have_pre || obs_in_regs);
                    PROF_MARK(1);  // layer 0
                }
                if (prev) do_last_epilogue(dacc_prev, inv_sigma_prev, pj0, pnobs, i0g(pgi), nsg(pgi), pchunk == 0);
                if (!defer) {
                    if (prev) {
                        if (!fused_sums) do_sums(nsg(pgi), pj0, pnobs, pchunk == 0);
                        compute_bar();
                        PROF_MARK(6);  // inv(Sigma) s + sums over observations
                        if (pchunk == p.C - 1) {
                            bar_sync(3, NB23);  // schedule/matrix rows, prior terms and noise of this step are in smem
                            PROF_MARK(2);       // (time spent waiting for the helper warp, normally zero)
                            if (p.part_out) {
                                for (int e = tid; e < nsg(pgi) * 2 * m; e += TC_CT) p.part_out[(size_t)i0g(pgi) * 2 * m + e] = sAcc[e];
                            } else {
                                if (p.xg.world > 0) xchg_group_sums(p.xg, sAcc, p.N, m, n_groups, p.SG, step_of(pit), i0g(pgi), nsg(pgi), tid, p.error_flag);
                                do_tail(pth, nsg(pgi));
                            }
                            compute_bar();
                            if (have) bar_arrive(2, NB23);
                            PROF_MARK(7);  // finalize
                        }
                    }
                    if (have && !early) {  // first chunk of an item: theta comes from the tail above
                        do_base(th, true, 0, step_of(it + 1));
                        compute_bar();
                        PROF_MARK(0);  // per-sample layer-0 part
                        do_layer0(i0g(gi), nsg(gi), j0, nobs, upre, obs_in_regs);
                        PROF_MARK(1);  // layer 0
                    }
                }
                if (have) {
                    for (int l = 1; l < Lh; ++l) {  // hidden layers that feed another tensor-core layer
                        const int Nn = net.Hp[l];
                        const int qcols = Nn >> 2, nblk = qcols / 16;
                        const float* bias_l = net.bias[l] + cq * qcols;
#pragma unroll
                        for (int e = 0; e < 4; ++e) bcur[e] = __ldg(reinterpret_cast<const float4*>(bias_l) + e);  // hidden behind the MMA
                        const uint32_t dacc = tmem_d + ((d_cnt & 1u) ? 256u : 0u);
                        TRACE(22);
                        mbar_wait(bar_dn(d_cnt), (d_cnt >> 1) & 1u, p.error_flag);
                        ++d_cnt;
                        tc_fence_after();
                        PROF_MARK(3);  // waiting for the tensor core
                        for (int cb = 0; cb < nblk; ++cb) {
                            const int col0 = cq * qcols + 16 * cb;
                            float4 bnxt[4];
                            if (cb + 1 < nblk) {
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    bnxt[e] = __ldg(reinterpret_cast<const float4*>(bias_l + 16 * (cb + 1)) + e);
                            }
                            float a[16];
                            ld_act(dacc, col0, bcur, a);
                            {   // next layer's A operand goes back to TENSOR MEMORY, in place over the 16 accumulator columns just
                                // read: columns col0 .. +7 = bf16 hi of K elements col0 .. col0 + 15, columns +8 .. +15 = lo
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
                                arrive_c(cb);
                            }
                            if (cb + 1 < nblk) {
#pragma unroll
                                for (int e = 0; e < 4; ++e) bcur[e] = bnxt[e];
                            }
                        }
                        PROF_MARK(4);  // epilogues
                    }
                }
                if (!defer && have && chunk + 1 < p.C) {  // next pass = next chunk of this item: the scores used sBase as scratch
                    do_base(th, false, step_of(it), 0);
                    compute_bar();
                    PROF_MARK(0);
                    if (l0_early) {  // its layer 0 now, under the last tensor-core layer of this chunk
                        const int nj0 = (chunk + 1) * p.OC;
                        do_layer0(i0g(gi), nsg(gi), nj0, min(p.OC, p.n - nj0), upre, false);
                        PROF_MARK(1);
                    }
                }
                if (defer) {
                    const bool pre_next = !obs_in_regs && l0_early && ps + 1 < total_p && lane < p.OC && p.fin.ctx_samples == 0;
                    TRACE(23);
                    if (pre_next) l0_load(upre, 0, min(p.OC, p.n), lane);  // obs_feat rows of the next item's layer 0: latency hidden
                    TRACE(24);
                    if (prev) {
                        if (!fused_sums) do_sums(nsg(pgi), 0, pnobs, true);
                        TRACE(25);
                        compute_bar();  // (redundant: every compute thread also takes part in barrier 3 below)
                        PROF_MARK(6);
                        bar_sync(3, NB23);  // schedule / matrix rows, prior terms and noise of the previous item are in smem
                        PROF_MARK(2);  // (time spent waiting for the helper warp)
                        if (p.xg.world > 0) xchg_group_sums(p.xg, sAcc, p.N, m, n_groups, p.SG, step_of(pit), i0g(pgi), nsg(pgi), tid, p.error_flag);
                        do_tail(pth, nsg(pgi));
                        PROF_MARK(10);  // tail proper (the barrier wait below stays in slot 6)
                        compute_bar();
                        if (have) bar_arrive(2, NB23);  // helper warp: buffers are free, go on with this item
                        PROF_MARK(6);
                    }
                    if (ps + 1 < total_p) {
                        // first pass of a block: nothing since layer 0 of item 0 has synchronised the compute warps (later passes
                        // go through the barriers of the last epilogue / tail) -- warps without layer-0 work (narrow first layer)
                        // would overwrite the sample part while the others still read it
                        if (!prev) compute_bar();
                        do_base(sTheta + (gi ^ 1) * TC_SG_MAX * MMAX, true, 0, step_of(it + 2));
                        compute_bar();
                        PROF_MARK(0);
                        if (l0_early) {  // layer 0 of the next item now, under the last tensor-core layer of this one
                            do_layer0(i0g(gi ^ 1), nsg(gi ^ 1), 0, min(p.OC, p.n), upre, pre_next || obs_in_regs);
                            PROF_MARK(1);
                        }
                    }
                }
            }  // passes
            if (!p.part_out) {
                for (int e = tid; e < ngr * TC_SG_MAX * MMAX; e += TC_CT) {
                    const int gi = e / (TC_SG_MAX * MMAX), r = e - gi * TC_SG_MAX * MMAX;
                    const int si = r / MMAX, k = r - si * MMAX;
                    if (si < nsg(gi) && k < m) p.theta[(size_t)(i0g(gi) + si) * m + k] = sTheta[e];
                }
            }
            compute_bar();
            gk += ngr;
        }  // group blocks
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();  // neither CTA leaves (or frees tensor memory) while the other may still use its shared memory
    if (warp == TC_CW + 1) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512));
    }
}

// plain launch, or a cluster of 2 CTAs (the two SMs of a TPC) for the CTA-pair instantiation
template <int MT, bool PAIR>
static cudaError_t launch_kernel(const TrajParams& p, int n_ctas, cudaStream_t stream) {
    if (!PAIR) {
        traj_tc_kernel<MT, false><<<n_ctas, TC_THREADS, TC_SMEM_BYTES, stream>>>(p);
        return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n_ctas, 1, 1);
    cfg.blockDim = dim3(TC_THREADS, 1, 1);
    cfg.dynamicSmemBytes = TC_SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, traj_tc_kernel<MT, PAIR>, p);
}

template <int MT, bool PAIR>
static cudaError_t launch_traj_tc_m(const TrajParams& p, int n_ctas, cudaStream_t stream) {
    // the attribute is per device: keep one flag per device ordinal (a process may drive several GPUs)
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(traj_tc_kernel<MT, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        // setmaxnreg moves registers inside the CTA's launch allocation: with fewer than 96 registers per thread at launch
        // the compute warps' `inc` could never be served (the kernel would hang), so refuse to launch instead
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, traj_tc_kernel<MT, PAIR>);
        if (e != cudaSuccess) return e;
        if (fa.numRegs < 96) return cudaErrorLaunchOutOfResources;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
#ifdef D4S_TRACE
    static long long* trace_buf = nullptr;
    const size_t trace_bytes = (size_t)(TC_THREADS / 32) * TRACE_MAX * 2 * sizeof(long long);
    if (const char* path = getenv("D4S_TRACE_FILE")) {
        if (!trace_buf) cudaMalloc(&trace_buf, trace_bytes);
        cudaMemsetAsync(trace_buf, 0, trace_bytes, stream);
        TrajParams q = p;
        q.trace = trace_buf;
        cudaError_t e = launch_kernel<MT, PAIR>(q, n_ctas, stream);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return e;
        std::vector<long long> host(trace_bytes / sizeof(long long));
        cudaMemcpy(host.data(), trace_buf, trace_bytes, cudaMemcpyDeviceToHost);
        if (FILE* f = fopen(path, "wb")) {
            fwrite(host.data(), 1, trace_bytes, f);
            fclose(f);
        }
        return cudaGetLastError();
    }
#endif
    return launch_kernel<MT, PAIR>(p, n_ctas, stream);
}

cudaError_t launch_traj_tc(const TrajParams& p, int n_ctas, cudaStream_t stream) {
    if (p.cta_pair) {  // (chosen by the host only for the instantiated dimensions and an even grid)
        switch (p.net.m) {
            case 2: return launch_traj_tc_m<2, true>(p, n_ctas, stream);
            case 4: return launch_traj_tc_m<4, true>(p, n_ctas, stream);
            case 5: return launch_traj_tc_m<5, true>(p, n_ctas, stream);
            default: return cudaErrorInvalidValue;
        }
    }
    switch (p.net.m) {  // theta_dim of the benchmark tasks: SIR 2, Lotka-Volterra / JR-NMM 4, SLCP 5; anything else: generic
        case 2: return launch_traj_tc_m<2, false>(p, n_ctas, stream);
        case 4: return launch_traj_tc_m<4, false>(p, n_ctas, stream);
        case 5: return launch_traj_tc_m<5, false>(p, n_ctas, stream);
        case 10: return launch_traj_tc_m<10, false>(p, n_ctas, stream);  // the tall-data stress shape (BASELINE config 5)
        default: return launch_traj_tc_m<0, false>(p, n_ctas, stream);
    }
}

int traj_tc_smem_bytes() { return TC_SMEM_BYTES; }

}  // namespace d4s
