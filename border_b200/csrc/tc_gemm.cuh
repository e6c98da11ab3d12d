// tc_gemm.cuh -- tcgen05 (5th-gen tensor core) tile GEMM with gather operands, 3xTF32 split.
//
// Same contraction as gemm.cuh,  C[M,N] = sum_k A(m,k) B(k,n)  with the same separable-gather
// operands, but the multiply runs on the tensor cores:
//
//   * 8 producer warps load the A / B tiles from global memory (dense rows, im2col gather, u8
//     frames), split every fp32 value x into  hi = x & 0xffffe000  (exact TF32) and  lo = x - hi,
//     and store both into shared memory in the canonical UMMA K-major SWIZZLE_128B layout
//     (row r at r*128 B, 16-byte chunk c at ((c ^ (r & 7)) << 4)); generic-proxy stores are made
//     visible to the tensor core with fence.proxy.async before the mbarrier arrive.
//   * one thread of warp 8 issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) three times per
//     k-slice -- hi*hi, lo*hi, hi*lo -- accumulating fp32 in TMEM (BN columns); that is the
//     "3xTF32" scheme: error ~2^-21 relative, which keeps the 1e-4 loss parity with fp32.
//     tcgen05.commit releases the shared-memory stage back to the producers.
//   * after the last k-slice the 8 producer warps become the epilogue: tcgen05.ld (32x32b) their
//     lane quarter / column half of the accumulator, apply bias / ReLU / ReLU-mask, store C
//     (or the split-K partial).
//
// Every mbarrier wait is bounded: a mis-programmed pipeline raises g_tc_error instead of hanging
// the GPU.  Tile = 128 x BN x 32 (BN in {32, 64, 128}), 3 stages, 288 threads, 1 CTA/SM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "gemm.cuh"

namespace bb {

static __device__ int g_tc_error = 0;  // per translation unit (tc_gemm.cu reads its own copy)
// the library's pinned, mapped failure flag (common.cuh: device_error_flag), published per device at first launch: a
// timed-out pipeline is seen by every agent entry point, not only by the test hooks
static __device__ int* g_tc_err_flag = nullptr;
// debug trace (BB_TC_DEBUG bit 16): clock64 stamps of CTA (0,0,0)'s producer thread 0 and MMA thread
static __device__ long long g_tc_trace[2][64][8];  // per translation unit

namespace tc {

constexpr int BM = 128, BK = 32, NPROD = 256, NTHREADS = 288;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.b32 %0, 1, 0, P1;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: returns false (and flags the error) instead of spinning forever
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 22); ++it)
        if (mbar_try_wait(bar, parity)) return true;
    atomicExch(&g_tc_error, 1);
    if (g_tc_err_flag) { *reinterpret_cast<volatile int*>(g_tc_err_flag) = 17; __threadfence_system(); }
    return false;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    // UMMA shared-memory descriptor, K-major, SWIZZLE_128B: start address >> 4, LBO (unused for a
    // swizzled K-major operand) = 1, SBO = 8 rows * 128 B = 1024 B >> 4, version 1, layout type 2.
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void split_store(uint32_t hi_addr, uint32_t lo_addr, float4 v) {
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(hi_addr), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo_addr), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
}
__device__ __forceinline__ void split_store1(uint32_t hi_addr, uint32_t lo_addr, float x) {
    float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(hi_addr), "f"(h) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(lo_addr), "f"(x - h) : "memory");
}

__device__ __forceinline__ float4 u8x4_to_float4(uint32_t w) {
    const float s = 1.0f / 255.0f;
    return make_float4((float)(w & 0xff) * s, (float)((w >> 8) & 0xff) * s, (float)((w >> 16) & 0xff) * s,
                       (float)(w >> 24) * s);
}

// 16 accumulator columns at taddr plus the 16 columns `second` further on, summed (the two accumulator halves of
// the stacked 3xTF32 scheme)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
        "%14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16_sum(uint32_t taddr, uint32_t second, uint32_t* r) {
    uint32_t r2[16];
    tmem_ld16(taddr, r);
    tmem_ld16(taddr + second, r2);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
}

// swizzled byte offset of 16-byte chunk c (0..7) of row r inside a [rows][128 B] K-major tile
__device__ __forceinline__ uint32_t sw128(uint32_t r, uint32_t c) { return r * 128u + ((c ^ (r & 7u)) << 4); }

}  // namespace tc

// A operand in tensor memory (used by conv1_tc.cu): register -> TMEM stores and the [a_tmem] MMA form
namespace tc3 {
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::
            "r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
}  // namespace tc3

// A_KSRC: A(m,k) is k-contiguous (dense rows or gather rowbase[m] + koff[k]); else A(m,k) = A[k*lda + m].
// B_KSRC: B(k,n) = B[n*ldb + k]; else n-contiguous (dense B[k*ldb + n] or gather rowbase[k] + noff[n]).
// STAGES shared-memory stages, PF register sets of prefetched operand data per producer thread,
// MINB co-resident CTAs per SM (two small CTAs overlap one's prologue/epilogue with the other's
// main loop).
// FAST: the host guarantees whole tiles (M % 128 == 0, N % BN == 0, K % 32 == 0), float operands and 16-byte
// aligned groups (dense leading dimensions or gather tables that are multiples of 4), so the producers' loads
// carry no bounds / alignment branches -- the loop is paced by the producer warps' instruction count.
template <int BN, int STAGES, int PF, int MINB, bool A_KSRC, bool B_KSRC, bool A_U8, bool B_U8, bool FAST = false>
__global__ void __launch_bounds__(tc::NTHREADS, MINB) tc_gemm_kernel(GemmArgs g) {
    using namespace tc;
    constexpr uint32_t A_TILE = BM * 128, B_TILE = BN * 128;       // bytes per hi (or lo) tile
    constexpr uint32_t STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;
    constexpr int A_LD = BM * 8 / NPROD;                             // float4 per producer thread per stage (4)
    constexpr int B_LD = (BN * 8 + NPROD - 1) / NPROD;               // 1, 2 or 4
    // accumulator: columns [0, BN) = hi*hi + lo*hi, columns [BN, 2 BN) = hi*lo (see the MMA issuer); summed in the epilogue
    constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
    extern __shared__ uint8_t smem_dyn[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int k_begin = blockIdx.z * g.k_per_split;
    const int k_end = min(g.K, k_begin + g.k_per_split);
    const int nks = k_end > k_begin ? (k_end - k_begin + BK - 1) / BK : 0;
    const uint32_t tiles = (smem_u32(smem_dyn) + 1023u) & ~1023u;   // SWIZZLE_128B tiles need 1024 B alignment
    const bool trace_cta = (g.fence_mode & 256) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
    const bool trace0 = trace_cta && tid == 0;
    if (trace0) g_tc_trace[1][56][0] = clock64();

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), NPROD / 32);  // one arrive per producer warp
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        mbar_init(smem_u32(&accum_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    pdl_sync();  // barriers / tensor memory are set up while the previous kernel of the stream drains
    if (trace0) g_tc_trace[1][57][0] = clock64();

    if (warp < 8) {
        // ================================================================ producers
        const float* Af = reinterpret_cast<const float*>(g.A);
        const uint8_t* Au = reinterpret_cast<const uint8_t*>(g.A);
        const float* Bf = reinterpret_cast<const float*>(g.B);
        const uint8_t* Bu = reinterpret_cast<const uint8_t*>(g.B);
        const bool a_vec = A_U8 || (((g.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0));
        const bool b_vec = B_U8 || (((g.ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0));
        // k-contiguous operands: a thread owns chunk (tid & 7) of rows (tid >> 3) + 32 i
        long a_base[A_LD], b_base[B_LD];
        if (A_KSRC) {
#pragma unroll
            for (int i = 0; i < A_LD; ++i) {
                int m = m0 + (tid >> 3) + 32 * i;
                a_base[i] = m < g.M ? (g.a_rowbase ? (long)g.a_rowbase[m] : (long)m * g.lda) : -1;
            }
        }
        if (B_KSRC) {
#pragma unroll
            for (int i = 0; i < B_LD; ++i) {
                int r = (tid >> 3) + 32 * i;
                int n = n0 + r;
                b_base[i] = (r < BN && n < g.N) ? (long)n * g.ldb : -1;
            }
        }
        // m-contiguous A: address = koff-part(k) + m-part(m); dense is k*lda + m, a gather (the transposed
        // im2col matrix of a conv weight gradient) takes the m-part from a_rowbase[m], the k-part from a_koff[k]
        long a_moff[A_LD];
        if (!A_KSRC) {
#pragma unroll
            for (int i = 0; i < A_LD; ++i) {
                int m = m0 + (warp + 8 * i) * 4;
                a_moff[i] = m < g.M ? (g.a_rowbase ? (long)g.a_rowbase[m] : (long)m) : 0;
            }
        }
        // The only per-stage table entries are a_koff[k] (k-contiguous gather A) and b_rowbase[k]
        // (n-contiguous gather B); they are fetched two stages ahead of the data loads that depend on
        // them, and the data loads run two stages ahead of the shared-memory stores (3 register sets).
        long b_noff_r[B_LD];
        if (!B_KSRC) {
#pragma unroll
            for (int i = 0; i < B_LD; ++i) {
                int n = n0 + (warp + 8 * i) * 4;
                b_noff_r[i] = (n < g.N) ? (g.b_noff ? (long)g.b_noff[n] : (long)n) : 0;
            }
        }
        float4 ra[PF][A_LD], rb[PF][B_LD];
        long ta[PF], tb[PF];

        auto load_tab = [&](long& oa, long& ob, int ks) {
            const int k0 = k_begin + ks * BK;
            oa = 0; ob = 0;
            if (FAST) {
                const int ka = k0 + (A_KSRC ? (tid & 7) * 4 : lane);
                oa = g.a_koff ? (long)__ldg(g.a_koff + ka) : (A_KSRC ? (long)ka : (long)ka * g.lda);
                if (!B_KSRC) ob = g.b_rowbase ? (long)__ldg(g.b_rowbase + k0 + lane) : (long)(k0 + lane) * g.ldb;
                return;
            }
            if (A_KSRC) {
                int k = k0 + (tid & 7) * 4;
                if (k < k_end) oa = g.a_koff ? (long)__ldg(g.a_koff + k) : (long)k;
            } else {
                int k = k0 + lane;
                if (k < k_end) oa = g.a_koff ? (long)__ldg(g.a_koff + k) : (long)k * g.lda;
            }
            if (!B_KSRC) {
                int k = k0 + lane;
                if (k < k_end) ob = g.b_rowbase ? (long)__ldg(g.b_rowbase + k) : (long)k * g.ldb;
            }
        };

        // one 16-byte group of the A tile (group i of this thread) of k-slice ks
        auto load_a = [&](int i, int ks, long tabA) -> float4 {
            if (!FAST && (g.fence_mode & 16)) return make_float4(1.f, 1.f, 1.f, 1.f);
            if (FAST && !A_U8) {
                if (A_KSRC) return __ldg(reinterpret_cast<const float4*>(Af + a_base[i] + tabA));
                return __ldg(reinterpret_cast<const float4*>(Af + tabA + a_moff[i]));
            }
            const int k0 = k_begin + ks * BK;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (A_KSRC) {
                int k = k0 + (tid & 7) * 4;
                if (a_base[i] >= 0 && k < k_end) {
                    long off = a_base[i] + tabA;
                    if (A_U8) {
                        v = u8x4_to_float4(__ldg(reinterpret_cast<const uint32_t*>(Au + off)));
                        if (k + 1 >= k_end) v.y = 0.f;
                        if (k + 2 >= k_end) v.z = 0.f;
                        if (k + 3 >= k_end) v.w = 0.f;
                    } else if (a_vec && k + 3 < k_end && ((off & 3) == 0)) {
                        v = __ldg(reinterpret_cast<const float4*>(Af + off));
                    } else {
                        v.x = __ldg(Af + off);
                        if (k + 1 < k_end) v.y = __ldg(Af + off + (g.a_koff ? g.a_koff[k + 1] - g.a_koff[k] : 1));
                        if (k + 2 < k_end) v.z = __ldg(Af + off + (g.a_koff ? g.a_koff[k + 2] - g.a_koff[k] : 2));
                        if (k + 3 < k_end) v.w = __ldg(Af + off + (g.a_koff ? g.a_koff[k + 3] - g.a_koff[k] : 3));
                    }
                }
            } else {  // 4 consecutive m at one k: lanes run along k (conflict-free transposing stores)
                int k = k0 + lane;
                int m = m0 + (warp + 8 * i) * 4;
                if (k < k_end && m < g.M) {
                    long off = tabA + a_moff[i];
                    if (A_U8) {
                        v = u8x4_to_float4(__ldg(reinterpret_cast<const uint32_t*>(Au + off)));
                        if (m + 1 >= g.M) v.y = 0.f;
                        if (m + 2 >= g.M) v.z = 0.f;
                        if (m + 3 >= g.M) v.w = 0.f;
                    } else if (a_vec && m + 3 < g.M && ((off & 3) == 0)) {
                        v = __ldg(reinterpret_cast<const float4*>(Af + off));
                    } else {
                        v.x = __ldg(Af + off);
                        if (m + 1 < g.M) v.y = __ldg(Af + off + (g.a_rowbase ? g.a_rowbase[m + 1] - g.a_rowbase[m] : 1));
                        if (m + 2 < g.M) v.z = __ldg(Af + off + (g.a_rowbase ? g.a_rowbase[m + 2] - g.a_rowbase[m] : 2));
                        if (m + 3 < g.M) v.w = __ldg(Af + off + (g.a_rowbase ? g.a_rowbase[m + 3] - g.a_rowbase[m] : 3));
                    }
                }
            }
            return v;
        };
        auto load_b = [&](int i, int ks, long tabB) -> float4 {
            if (!FAST && (g.fence_mode & 16)) return make_float4(1.f, 1.f, 1.f, 1.f);
            if (FAST && !B_U8) {
                if (B_KSRC) return __ldg(reinterpret_cast<const float4*>(Bf + b_base[i] + (k_begin + ks * BK + (tid & 7) * 4)));
                return __ldg(reinterpret_cast<const float4*>(Bf + tabB + b_noff_r[i]));
            }
            const int k0 = k_begin + ks * BK;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (B_KSRC) {
                int k = k0 + (tid & 7) * 4;
                if (b_base[i] >= 0 && k < k_end) {
                    long off = b_base[i] + k;
                    if (b_vec && k + 3 < k_end) v = __ldg(reinterpret_cast<const float4*>(Bf + off));
                    else {
                        v.x = __ldg(Bf + off);
                        if (k + 1 < k_end) v.y = __ldg(Bf + off + 1);
                        if (k + 2 < k_end) v.z = __ldg(Bf + off + 2);
                        if (k + 3 < k_end) v.w = __ldg(Bf + off + 3);
                    }
                }
            } else {
                int k = k0 + lane;
                int n4 = warp + 8 * i;          // group of 4 consecutive n
                int n = n0 + n4 * 4;
                if (n4 * 4 < BN && k < k_end && n < g.N) {
                    long off = tabB + b_noff_r[i];
                    if (B_U8) {
                        v = u8x4_to_float4(__ldg(reinterpret_cast<const uint32_t*>(Bu + off)));
                    } else if (b_vec && n + 3 < g.N && ((off & 3) == 0)) {
                        v = __ldg(reinterpret_cast<const float4*>(Bf + off));
                    } else {
                        v.x = __ldg(Bf + off);
                        if (n + 1 < g.N) v.y = __ldg(Bf + off + (g.b_noff ? g.b_noff[n + 1] - g.b_noff[n] : 1));
                        if (n + 2 < g.N) v.z = __ldg(Bf + off + (g.b_noff ? g.b_noff[n + 2] - g.b_noff[n] : 2));
                        if (n + 3 < g.N) v.w = __ldg(Bf + off + (g.b_noff ? g.b_noff[n + 3] - g.b_noff[n] : 3));
                    }
                }
            }
            return v;
        };
        // the matching split + swizzled stores of one group
        auto store_a = [&](int i, const float4& v4, uint32_t a_hi, uint32_t a_lo) {
            if (A_KSRC) {
                uint32_t off = sw128((uint32_t)(tid >> 3) + 32u * i, (uint32_t)(tid & 7));
                split_store(a_hi + off, a_lo + off, v4);
            } else {
                uint32_t r = (uint32_t)(warp + 8 * i) * 4u;
                const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t off = sw128(r + j, (uint32_t)lane >> 2) + ((uint32_t)lane & 3u) * 4u;
                    split_store1(a_hi + off, a_lo + off, v[j]);
                }
            }
        };
        auto store_b = [&](int i, const float4& v4, uint32_t b_hi, uint32_t b_lo) {
            if (B_KSRC) {
                uint32_t r = (uint32_t)(tid >> 3) + 32u * i;
                if (r < (uint32_t)BN) {
                    uint32_t off = sw128(r, (uint32_t)(tid & 7));
                    split_store(b_hi + off, b_lo + off, v4);
                }
            } else {
                uint32_t r = (uint32_t)(warp + 8 * i) * 4u;
                if (r < (uint32_t)BN) {
                    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t off = sw128(r + j, (uint32_t)lane >> 2) + ((uint32_t)lane & 3u) * 4u;
                        split_store1(b_hi + off, b_lo + off, v[j]);
                    }
                }
            }
        };

        // Software pipeline with PF register sets and the loads issued PF k-slices ahead: every 16-byte group of
        // k-slice k is split + stored and its register is IMMEDIATELY reloaded with the same group of k-slice k+PF
        // (the trace showed the producers, not the MMAs, pacing the loop: with the reload deferred to the next
        // iteration only one slice of loads was in flight).  Gather-table entries run one slice further ahead.
        bool alive = true;
        long nta = 0, ntb = 0;   // table entries of the next k-slice to be loaded
#pragma unroll
        for (int j = 0; j < PF; ++j)
            if (j < nks) load_tab(ta[j], tb[j], j);
#pragma unroll
        for (int j = 0; j < PF; ++j)
            if (j < nks) {
#pragma unroll
                for (int i = 0; i < A_LD; ++i) ra[j][i] = load_a(i, j, ta[j]);
#pragma unroll
                for (int i = 0; i < B_LD; ++i) rb[j][i] = load_b(i, j, tb[j]);
            }
        if (PF < nks) load_tab(nta, ntb, PF);
        // one k-slice: wait for the stage, split + store every group, reload it (MORE) with the slice PF ahead
        auto slice = [&](auto more_tag, float4* pa, float4* pb, int k) {
            constexpr bool MORE = decltype(more_tag)::value;
            const int s = k % STAGES;
            const uint32_t ph = (uint32_t)(k / STAGES) & 1u;
            if (!FAST && trace0 && k < 56) g_tc_trace[0][k][0] = clock64();
            if (alive && !mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u)) alive = false;
            if (!FAST && trace0 && k < 56) g_tc_trace[0][k][2] = clock64();
            const uint32_t a_hi = tiles + s * STAGE_BYTES, a_lo = a_hi + A_TILE, b_hi = a_lo + A_TILE, b_lo = b_hi + B_TILE;
#pragma unroll
            for (int i = 0; i < A_LD; ++i) {
                store_a(i, pa[i], a_hi, a_lo);
                if (MORE) pa[i] = load_a(i, k + PF, nta);
            }
#pragma unroll
            for (int i = 0; i < B_LD; ++i) {
                store_b(i, pb[i], b_hi, b_lo);
                if (MORE) pb[i] = load_b(i, k + PF, ntb);
            }
            // generic stores -> async proxy (UMMA).  A fence waits for ALL of the thread's outstanding
            // memory operations, including the prefetched global loads, so the writer-side fence
            // serialises the load latency into every stage; fence_mode 1 moves it to the consumer.
            if ((g.fence_mode & 1) == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();  // orders the warp's st.shared before lane 0's release-arrive (256 arrives/stage were costly)
            if (lane == 0) mbar_arrive(smem_u32(&full_bar[s]));
            if (!FAST && trace0 && k < 56) g_tc_trace[0][k][3] = clock64();
            if (MORE && k + PF + 1 < nks) load_tab(nta, ntb, k + PF + 1);
        };
        for (int ks = 0; ks < nks; ks += PF) {
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int k = ks + u;
                if (k < nks) {
                    if (k + PF < nks) slice(std::true_type{}, ra[u], rb[u], k);
                    else slice(std::false_type{}, ra[u], rb[u], k);
                }
            }
        }

        // ================================================================ epilogue
        if (trace0) g_tc_trace[1][58][0] = clock64();
        const bool direct = g.split_k <= 1;
        float* out = direct ? g.C : g.workspace + (size_t)blockIdx.z * g.M * g.N;
        const int ldo = direct ? g.ldc : (g.trans_out ? g.M : g.N);
        const int q = warp & 3;                     // TMEM lane quarter this warp may read
        const int m = m0 + q * 32 + lane;
        const bool mapped = direct && g.c_rowoff != nullptr;
        const size_t rowo = m < g.M ? (mapped ? (size_t)g.c_rowoff[m] : (size_t)m * ldo) : 0;
        constexpr int HALF = BN / 2 < 16 ? 16 : BN / 2;   // columns per warp-pair member
        const int c_begin = (warp >> 2) * HALF;
        // FAST, whole-tile problems with 16-byte aligned rows: the bias / ReLU-mask groups of this thread's columns are
        // fetched as float4 BEFORE the accumulator is waited for (they do not depend on it), and stored as float4.
        constexpr bool EPI_VEC = FAST && BN <= 64;
        bool epi_vec = false;
        if (EPI_VEC)
            epi_vec = !g.trans_out && (ldo & 3) == 0 && (!mapped || g.tables_vec4) &&
                      ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(g.bias) | reinterpret_cast<uintptr_t>(g.mask)) & 15) == 0;
        if (EPI_VEC && epi_vec) {
            constexpr int G = HALF / 4;
            const bool use_bias = direct && g.bias != nullptr, use_mask = direct && g.mask != nullptr;
            float4 b4[G], k4[G];
            size_t colo[G];
#pragma unroll
            for (int gi = 0; gi < G; ++gi) {
                const int n = n0 + c_begin + 4 * gi;
                colo[gi] = (mapped && g.c_coloff) ? (size_t)g.c_coloff[n] : (size_t)n;
                b4[gi] = use_bias ? __ldg(reinterpret_cast<const float4*>(g.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
                k4[gi] = use_mask ? __ldg(reinterpret_cast<const float4*>(g.mask + rowo + colo[gi])) : make_float4(1.f, 1.f, 1.f, 1.f);
            }
            if (nks > 0 && alive) alive = mbar_wait(smem_u32(&accum_bar), 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (trace0) g_tc_trace[1][59][0] = clock64();
#pragma unroll
            for (int c0 = 0; c0 < HALF; c0 += 16) {
                uint32_t r[16];
                if (nks > 0 && alive) {
                    tmem_ld16_sum(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c_begin + c0), (uint32_t)BN, r);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) r[j] = 0u;
                }
#pragma unroll
                for (int j4 = 0; j4 < 16; j4 += 4) {
                    const int gi = (c0 + j4) / 4;
                    float4 v = make_float4(__uint_as_float(r[j4]), __uint_as_float(r[j4 + 1]), __uint_as_float(r[j4 + 2]),
                                           __uint_as_float(r[j4 + 3]));
                    if (direct) {
                        v.x += b4[gi].x; v.y += b4[gi].y; v.z += b4[gi].z; v.w += b4[gi].w;
                        if (g.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                        v.x = k4[gi].x > 0.f ? v.x : 0.f; v.y = k4[gi].y > 0.f ? v.y : 0.f;
                        v.z = k4[gi].z > 0.f ? v.z : 0.f; v.w = k4[gi].w > 0.f ? v.w : 0.f;
                    }
                    *reinterpret_cast<float4*>(out + rowo + colo[gi]) = v;
                    if (direct && g.c_plane)
                        *reinterpret_cast<float4*>(out + g.c_plane + rowo + colo[gi]) =
                            make_float4(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u), v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u),
                                        v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u), v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u));
                }
            }
        } else {
        if (nks > 0 && alive) alive = mbar_wait(smem_u32(&accum_bar), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (trace0) g_tc_trace[1][59][0] = clock64();
        if (c_begin < BN) {
#pragma unroll
            for (int c0 = 0; c0 < HALF; c0 += 16) {
                const int col = c_begin + c0;
                if (col >= BN) break;
                uint32_t r[16];
                if (nks > 0 && alive) {
                    tmem_ld16_sum(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col, (uint32_t)BN, r);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) r[j] = 0u;
                }
                if (m < g.M) {
#pragma unroll
                    for (int j4 = 0; j4 < 16; j4 += 4) {
                        const int n = n0 + col + j4;
                        if (n >= g.N) break;
                        float v[4];
                        const size_t colo = (mapped && g.c_coloff) ? (size_t)g.c_coloff[n] : (size_t)n;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            v[j] = __uint_as_float(r[j4 + j]);
                            if (direct && n + j < g.N) {
                                if (g.bias) v[j] += g.bias[n + j];
                                if (g.relu) v[j] = fmaxf(v[j], 0.f);
                                if (g.mask) v[j] = g.mask[rowo + colo + j] > 0.f ? v[j] : 0.f;
                            }
                        }
                        if (g.trans_out) {  // C^T: consecutive lanes (rows m) write consecutive addresses
                            for (int j = 0; j < 4; ++j)
                                if (n + j < g.N) out[(size_t)(n + j) * ldo + m] = v[j];
                            continue;
                        }
                        float* dst = out + rowo + colo;
                        if (direct && g.c_plane)
                            for (int j = 0; j < 4; ++j)
                                if (n + j < g.N) dst[g.c_plane + j] = v[j] - __uint_as_float(__float_as_uint(v[j]) & 0xffffe000u);
                        if (n + 3 < g.N && (((rowo + colo) & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0))
                            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                        else
                            for (int j = 0; j < 4; ++j)
                                if (n + j < g.N) dst[j] = v[j];
                    }
                }
            }
        }
        }
        if (trace0) g_tc_trace[1][60][0] = clock64();
    } else if (lane == 0) {
        // ================================================================ MMA issuer (one thread)
        const bool trace = trace_cta;
        // instruction descriptor: D = F32 (bit 4), A = B = TF32 (2 << 7, 2 << 10), both K-major,
        // N >> 3 at bit 17, M >> 4 at bit 24
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        // "stacked" 3xTF32: the hi and lo tiles of B are adjacent in shared memory, so ONE descriptor spans [B_hi; B_lo]
        // as 2 BN rows.  A_hi x [B_hi; B_lo] (N = 2 BN) yields hi*hi in columns [0, BN) and hi*lo in [BN, 2 BN);
        // A_lo x B_hi (N = BN) accumulates lo*hi into [0, BN).  8 MMAs per k-slice instead of 12, A_hi read once
        // instead of twice (56 KB instead of 72 KB of shared-memory operand reads per 128x64x32 slice).
        const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        bool alive = true;
        for (int ks = 0; ks < nks && alive; ++ks) {
            const int s = ks % STAGES;
            const uint32_t ph = (uint32_t)(ks / STAGES) & 1u;
            if (trace && ks < 56) g_tc_trace[1][ks][0] = clock64();
            if (!mbar_wait(smem_u32(&full_bar[s]), ph)) { alive = false; break; }
            if (trace && ks < 56) g_tc_trace[1][ks][1] = clock64();
            if ((g.fence_mode & 1) == 1 && !(g.fence_mode & 64)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = tiles + s * STAGE_BYTES, a_lo = a_hi + A_TILE, b_hi = a_lo + A_TILE;  // b_lo = b_hi + B_TILE follows
            const uint64_t da_hi = make_desc(a_hi), da_lo = make_desc(a_lo), db_hi = make_desc(b_hi);
#pragma unroll
            for (int k4 = 0; k4 < BK / 8; ++k4) {
                const uint64_t adv = (uint64_t)(k4 * 2);  // 8 tf32 = 32 B = 2 x 16 B along K inside the swizzle atom
                if (g.fence_mode & 32) continue;
                mma_tf32(tmem_base, da_hi + adv, db_hi + adv, idesc2, (ks | k4) ? 1u : 0u);
                if (g.fence_mode & 128) continue;  // debug: skip the lo*hi pass (timing floor, wrong numerics)
                mma_tf32(tmem_base, da_lo + adv, db_hi + adv, idesc, 1u);
            }
            mma_commit(smem_u32(&empty_bar[s]));  // frees the stage once the MMAs above have read it
            if (trace && ks < 56) g_tc_trace[1][ks][3] = clock64();
        }
        if (nks > 0) mma_commit(smem_u32(&accum_bar));
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
    if (trace0) g_tc_trace[1][61][0] = clock64();
}

}  // namespace bb
