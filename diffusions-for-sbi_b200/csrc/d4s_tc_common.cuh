// Shared pieces of the tcgen05 kernels (d4s_traj_tc.cu: GAUSS trajectory, d4s_jac_tc.cu: JAC step): PTX wrappers for
// mbarriers, 1-D TMA bulk copies, tcgen05.mma / ld / commit, the UMMA descriptors of the canonical K-major no-swizzle
// layout, and the bf16 hi/lo operand split.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "d4s_common.cuh"

namespace d4s {

constexpr int TC_M = 128;                      // tile rows = TMEM lanes
constexpr int TC_CW = 16;                      // compute warps
constexpr int TC_CT = TC_CW * 32;              // compute threads
constexpr int TC_STAGE_BYTES = 256 * 64;       // one K=16 step of W_hi | W_lo for N <= 256
constexpr int TC_A_BYTES = (256 / 16) * 4096;  // 64 KB: act[128 x 256] bf16 in K16-step images
constexpr long long SPIN_LIMIT_CYCLES = 4000000000LL;  // ~2 s: bounded mbarrier waits trap instead of hanging

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
    unsigned polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++polls & 255u) == 0u && clock64() - t0 > SPIN_LIMIT_CYCLES) {
            if (err) atomicExch(err, 1);  // a protocol bug must not hang the GPU: fail loudly instead
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
__device__ __forceinline__ void compute_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

// tcgen05.mma with the A operand in TENSOR MEMORY (D[tmem] (+)= A[tmem] x B[smem]): A row r sits in TMEM lane r, its K = 16
// bf16 elements of one instruction in 8 consecutive 32-bit columns (two elements per column, K-contiguous).
__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane <- registers
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- CTA pair (cta_group::2): two CTAs of one cluster (the two SMs of a TPC) run ONE tcgen05.mma over M = 256 rows; each
// CTA supplies its own 128 A rows (shared or tensor memory, same offsets in both CTAs) and HALF of the B tile (N / 2 rows of
// the K-major image), and receives its 128 accumulator rows in its own tensor memory.  Per SM that halves the weight bytes
// streamed from L2 and the B-operand reads of shared memory.  Only the leader (cluster rank 0) issues the instructions; its
// tcgen05.commit signals the mbarrier at the same offset in both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of both CTAs
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // (default .release.cta semantics, as CUTLASS' ClusterBarrier::arrive(cta_id): what the arrival publishes -- this CTA's operand
    // image in its OWN shared / tensor memory, read by the pair's MMA through the async proxy -- never crosses to the peer's
    // generic proxy; `.release.cluster` compiles to MEMBAR.ALL.GPU + CCTL.IVALL per arrival, i.e. an L1 flush per thread)
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tc_commit2(uint32_t bar) {  // arrives on `bar` of BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc_mma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_mma2_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// system-scope flag accesses of the in-kernel exchange between ranks (peer memory over NVLink)
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
constexpr long long XCHG_SPIN_LIMIT_CYCLES = 40000000000LL;  // ~20 s: a missing peer makes the kernel trap, not hang

// UMMA shared-memory descriptor, K-major, no swizzle ("interleave"): 8x16B core matrices, 128 B each.
//   lbo = byte distance between the two K halves of a K=16 step, sbo = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor: D fp32, A/B bf16, both K-major, dense, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc(int n, int m_rows = TC_M) {  // m_rows = 256: CTA pair (cta_group::2)
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m_rows >> 4) << 24);
}
// same with the operand formats chosen per instruction: bit 7 / bit 10 set = A / B is bf16, clear = fp16
__device__ __forceinline__ uint32_t umma_idesc_fmt(int n, bool a_bf16, bool b_bf16, int m_rows = TC_M) {
    return (1u << 4) | (a_bf16 ? (1u << 7) : 0u) | (b_bf16 ? (1u << 10) : 0u) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m_rows >> 4) << 24);
}

// named barriers: 1 = compute warps only; 2 = theta of the step is final (compute arrive, helper sync);
//                 3 = helper results ready (helper arrive, compute sync)
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// Splits 8 fp32 activations into bf16 hi / lo and stores the two 16-byte K-groups.  hi = RN_bf16(v) by Veltkamp
// splitting (c = v * (2^16 + 1); hi = c - (c - v): three FMA-pipe ops, no conversion-unit instruction), lo = v - hi
// exactly, truncated to bf16 when packed (its dropped bits are < 2^-17 |v|).  Packing is one PRMT per pair.
__device__ __forceinline__ void store_split8(const float (&v)[8], uint8_t* a_hi, uint8_t* a_lo, int byte_off) {
    uint32_t hi[4], lo[4];
    const float2 k = make_float2(65537.0f, 65537.0f), neg1 = make_float2(-1.0f, -1.0f);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        // packed fp32 pairs (FMUL2 / FFMA2): a - b = fma(b, -1, a) rounds once, exactly like a subtraction
        const float2 x = make_float2(v[2 * e], v[2 * e + 1]);
        const float2 c = __fmul2_rn(x, k);
        const float2 d = __ffma2_rn(x, neg1, c);  // c - x
        const float2 h = __ffma2_rn(d, neg1, c);  // hi = c - (c - x)
        const float2 l = __ffma2_rn(h, neg1, x);  // lo = x - hi (exact)
        hi[e] = __byte_perm(__float_as_uint(h.x), __float_as_uint(h.y), 0x7632);
        lo[e] = __byte_perm(__float_as_uint(l.x), __float_as_uint(l.y), 0x7632);
    }
    *reinterpret_cast<uint4*>(a_hi + byte_off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(a_lo + byte_off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// the same split into registers: hi[e] / lo[e] = packed bf16 pair of elements (2e, 2e + 1)
__device__ __forceinline__ void split8_regs(const float (&v)[8], uint32_t* hi, uint32_t* lo) {
    const float2 k = make_float2(65537.0f, 65537.0f), neg1 = make_float2(-1.0f, -1.0f);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 x = make_float2(v[2 * e], v[2 * e + 1]);
        const float2 c = __fmul2_rn(x, k);
        const float2 d = __ffma2_rn(x, neg1, c);
        const float2 h = __ffma2_rn(d, neg1, c);
        const float2 l = __ffma2_rn(h, neg1, x);
        hi[e] = __byte_perm(__float_as_uint(h.x), __float_as_uint(h.y), 0x7632);
        lo[e] = __byte_perm(__float_as_uint(l.x), __float_as_uint(l.y), 0x7632);
    }
}

// two floats -> packed fp16 pair (x0 in the low half), round to nearest, saturating at +-65504 instead of overflowing to inf
__device__ __forceinline__ uint32_t pack_f16x2_sat(float x0, float x1) {
    uint32_t d;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(x1), "f"(x0));
    return d;
}

// fp16 hi / lo split: hi = v rounded to 11 significant bits (Veltkamp with 2^13 + 1), lo = v - hi rounded to fp16.
// hi + lo carries ~22 bits, so a_hi w_hi + a_lo w_hi + a_hi w_lo is an fp32-grade product.  Range: |v| < 65504; below 6e-5 the parts are fp16
// subnormals and the absolute error stays <= 2^-25.  Values beyond the fp16 range saturate (only diverged samples get there).
__device__ __forceinline__ void store_split8_f16(const float (&v)[8], uint8_t* a_hi, uint8_t* a_lo, int byte_off) {
    uint32_t hi[4], lo[4];
    const float2 k = make_float2(8193.0f, 8193.0f), neg1 = make_float2(-1.0f, -1.0f);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 x = make_float2(v[2 * e], v[2 * e + 1]);
        const float2 c = __fmul2_rn(x, k);
        const float2 d = __ffma2_rn(x, neg1, c);
        const float2 h = __ffma2_rn(d, neg1, c);
        const float2 l = __ffma2_rn(h, neg1, x);
        hi[e] = pack_f16x2_sat(h.x, h.y);
        lo[e] = pack_f16x2_sat(l.x, l.y);
    }
    *reinterpret_cast<uint4*>(a_hi + byte_off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(a_lo + byte_off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// the same split into registers: hi[e] / lo[e] = packed fp16 pair of elements (2e, 2e + 1)
__device__ __forceinline__ void split8_regs_f16(const float (&v)[8], uint32_t* hi, uint32_t* lo) {
    const float2 k = make_float2(8193.0f, 8193.0f), neg1 = make_float2(-1.0f, -1.0f);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 x = make_float2(v[2 * e], v[2 * e + 1]);
        const float2 c = __fmul2_rn(x, k);
        const float2 d = __ffma2_rn(x, neg1, c);
        const float2 h = __ffma2_rn(d, neg1, c);
        const float2 l = __ffma2_rn(h, neg1, x);
        hi[e] = pack_f16x2_sat(h.x, h.y);
        lo[e] = pack_f16x2_sat(l.x, l.y);
    }
}

// The tensor core accumulates by TRUNCATION: every tcgen05.mma that adds into an fp32 accumulator loses, on average, a fixed
// fraction of an ulp of the running sum, always towards zero -- a systematic relative deficit of the result that grows with the
// number of accumulating instructions at full magnitude.  Measured on the JAC fixtures, where inv(I - sigma J) amplifies it: the
// error of the Jacobian sums is linear in a correction factor c and vanishes at c = +600 .. +900 ppb (K = 256) when the three
// products of a K-step are accumulated together (48 instructions per layer), at c = +250 ppb when only the K / 16 main-term
// instructions run at full magnitude (d4s_jac_tc.cu; profiles/r02_jac_tc_notes.md).  The JAC kernel's epilogues multiply the
// accumulator by 1 + c K / 256; the GAUSS trajectory kernel does not (its error is the bf16x3 product truncation, the correction
// changes nothing measurable there).
#ifndef D4S_ACC_DEBIAS_PPB
#define D4S_ACC_DEBIAS_PPB 250
#endif
__device__ __forceinline__ float acc_debias(int K) { return 1.f + (1e-9f * D4S_ACC_DEBIAS_PPB / 256.f) * (float)K; }

// byte offset of (row r, k-group kg of 8 columns) inside an A image: K16 step = kg/2, half = kg%2
__device__ __forceinline__ int a_off(int r, int kg) { return (kg >> 1) * 4096 + (kg & 1) * 2048 + r * 16; }

}  // namespace d4s
