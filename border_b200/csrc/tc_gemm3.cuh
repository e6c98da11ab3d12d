// tc_gemm3.cuh -- tcgen05 3xTF32 GEMM with the A operand in TENSOR MEMORY.
//
// tc_gemm.cuh stages both operands in shared memory; per 128x64x32 stage that is 48 KB of hi/lo
// st.shared plus 72 KB of UMMA operand reads, and ncu / the BB_TC_DEBUG bisect show the kernel
// bound by exactly that shared-memory traffic.  Here the producers keep A in registers after the
// global load, split it into hi / lo and write it straight into TMEM with tcgen05.st (thread =
// matrix row = TMEM lane, 32 k-values = 32 columns); tcgen05.mma then takes A from TMEM
// ("[a_tmem]" form) and only the small B tile goes through shared memory.  Shared-memory traffic
// per stage drops from ~144 KB to ~64 KB.
//
//   warps 0-3  producers of the even k-stages, warps 4-7 of the odd ones (warp w owns rows
//              32*(w%4).. of the tile, the only TMEM lanes it may touch); loads for a stage are
//              issued before waiting for its slot, so two stages per CTA and two CTAs per SM keep
//              four load streams in flight per SM.
//   warp 8     TMEM owner + MMA issuer: per k-slice  D += A_hi B_hi + A_lo B_hi + A_hi B_lo.
//   epilogue   by warps 0-7 after the last stage (lane quarter w%4, column half w/4).
//
// TMEM columns: [0, 64) accumulator, then 3 stages x (32 hi + 32 lo) = 256 columns per CTA.
#pragma once
#include "tc_gemm.cuh"

namespace bb {
namespace tc3 {
constexpr int STAGES = 3, NTHREADS = 288, ACC_COLS = 64, A_STAGE_COLS = 64, TMEM_COLS = 256;

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

template <int BN, bool A_KSRC, bool B_KSRC, bool A_U8, bool B_U8>
__global__ void __launch_bounds__(tc3::NTHREADS, 2) tc_gemm_tmem_kernel(GemmArgs g) {
    using namespace tc;
    using namespace tc3;
    static_assert(BN == 32 || BN == 64, "BN <= 64");
    constexpr uint32_t B_TILE = BN * 128;              // bytes of one hi (or lo) B tile
    constexpr uint32_t STAGE_BYTES = 2 * B_TILE;
    constexpr int B_LD = BN * 8 / 128;                  // float4 per thread of a 128-thread producer group (2 or 4)
    extern __shared__ uint8_t smem_dyn[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int k_begin = blockIdx.z * g.k_per_split;
    const int k_end = min(g.K, k_begin + g.k_per_split);
    const int nks = k_end > k_begin ? (k_end - k_begin + BK - 1) / BK : 0;
    const uint32_t tiles = (smem_u32(smem_dyn) + 1023u) & ~1023u;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 4);   // the 4 warps of one producer group
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        mbar_init(smem_u32(&accum_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp < 8) {
        // ================================================================ producers
        const int grp = warp >> 2;                 // 0: even stages, 1: odd stages
        const int q = warp & 3;                    // TMEM lane quarter
        const int row = q * 32 + lane;             // A row owned by this thread
        const int gt = tid & 127;                  // thread index inside the group
        const float* Af = reinterpret_cast<const float*>(g.A);
        const uint8_t* Au = reinterpret_cast<const uint8_t*>(g.A);
        const float* Bf = reinterpret_cast<const float*>(g.B);
        const uint8_t* Bu = reinterpret_cast<const uint8_t*>(g.B);
        const bool a_vec = A_U8 || (((g.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0));
        const bool b_vec = B_U8 || (((g.ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0));
        const int m = m0 + row;
        long a_base = -1;
        if (A_KSRC && m < g.M) a_base = g.a_rowbase ? (long)__ldg(g.a_rowbase + m) : (long)m * g.lda;
        long b_noff_r[B_LD];
        if (!B_KSRC) {
#pragma unroll
            for (int i = 0; i < B_LD; ++i) {
                int n = n0 + (q + 4 * i) * 4;
                b_noff_r[i] = (n < g.N) ? (g.b_noff ? (long)__ldg(g.b_noff + n) : (long)n) : 0;
            }
        }
        bool alive = true;
        for (int ks = grp; ks < nks; ks += 2) {
            const int k0 = k_begin + ks * BK;
            float av[32];
            float4 bv[B_LD];
            // ---------------- A: 32 k-values of this thread's row
            if (A_KSRC) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    int k = k0 + c * 4;
                    if (a_base >= 0 && k < k_end) {
                        long off = a_base + (g.a_koff ? (long)__ldg(g.a_koff + k) : (long)k);
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
                    av[c * 4 + 0] = v.x; av[c * 4 + 1] = v.y; av[c * 4 + 2] = v.z; av[c * 4 + 3] = v.w;
                }
            } else {  // A(m,k) = A[k*lda + m]: lanes are consecutive m -> coalesced 4-byte loads
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    int k = k0 + j;
                    av[j] = (m < g.M && k < k_end) ? __ldg(Af + (long)k * g.lda + m) : 0.f;
                }
            }
            // ---------------- B: BN x 32 tile through shared memory (hi / lo, SWIZZLE_128B K-major)
#pragma unroll
            for (int i = 0; i < B_LD; ++i) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (B_KSRC) {
                    int k = k0 + (gt & 7) * 4;
                    int r = (gt >> 3) + 16 * i;
                    int n = n0 + r;
                    if (n < g.N && k < k_end) {
                        long off = (long)n * g.ldb + k;
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
                    int n = n0 + (q + 4 * i) * 4;
                    if (k < k_end && n < g.N) {
                        long off = (g.b_rowbase ? (long)__ldg(g.b_rowbase + k) : (long)k * g.ldb) + b_noff_r[i];
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
                bv[i] = v;
            }
            // ---------------- wait for the slot, then publish
            const int s = ks % STAGES;
            const uint32_t ph = (uint32_t)(ks / STAGES) & 1u;
            if (alive && !mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u)) alive = false;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                hi[j] = __float_as_uint(av[j]) & 0xffffe000u;
                lo[j] = __float_as_uint(av[j] - __uint_as_float(hi[j]));
            }
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + ACC_COLS + (uint32_t)s * A_STAGE_COLS;
            tmem_st16(ta + 0, hi);
            tmem_st16(ta + 16, hi + 16);
            tmem_st16(ta + 32, lo);
            tmem_st16(ta + 48, lo + 16);
            const uint32_t b_hi = tiles + s * STAGE_BYTES, b_lo = b_hi + B_TILE;
#pragma unroll
            for (int i = 0; i < B_LD; ++i) {
                if (B_KSRC) {
                    uint32_t off = sw128((uint32_t)(gt >> 3) + 16u * i, (uint32_t)(gt & 7));
                    split_store(b_hi + off, b_lo + off, bv[i]);
                } else {
                    uint32_t r = (uint32_t)(q + 4 * i) * 4u;
                    const float v[4] = {bv[i].x, bv[i].y, bv[i].z, bv[i].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t off = sw128(r + j, (uint32_t)lane >> 2) + ((uint32_t)lane & 3u) * 4u;
                        split_store1(b_hi + off, b_lo + off, v[j]);
                    }
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&full_bar[s]));
        }

        // ================================================================ epilogue
        if (nks > 0 && alive) alive = mbar_wait(smem_u32(&accum_bar), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const bool direct = g.split_k <= 1;
        float* out = direct ? g.C : g.workspace + (size_t)blockIdx.z * g.M * g.N;
        const int ldo = direct ? g.ldc : g.N;
        constexpr int HALF = BN / 2;
        const int c_begin = grp * HALF;
#pragma unroll
        for (int c0 = 0; c0 < HALF; c0 += 16) {
            const int col = c_begin + c0;
            uint32_t r[16];
            if (nks > 0 && alive) {
                uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
                    "%14, %15}, [%16];\n"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = 0u;
            }
            if (m < g.M) {
#pragma unroll
                for (int j4 = 0; j4 < 16; j4 += 4) {
                    const int n = n0 + col + j4;
                    if (n < g.N) {
                        float v[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            v[j] = __uint_as_float(r[j4 + j]);
                            if (direct && n + j < g.N) {
                                if (g.bias) v[j] += g.bias[n + j];
                                if (g.relu) v[j] = fmaxf(v[j], 0.f);
                                if (g.mask) v[j] = g.mask[(size_t)m * g.ldc + n + j] > 0.f ? v[j] : 0.f;
                            }
                        }
                        float* dst = out + (size_t)m * ldo + n;
                        if (n + 3 < g.N && ((ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0))
                            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                        else
                            for (int j = 0; j < 4; ++j)
                                if (n + j < g.N) dst[j] = v[j];
                    }
                }
            }
        }
    } else if (lane == 0) {
        // ================================================================ MMA issuer
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        for (int ks = 0; ks < nks; ++ks) {
            const int s = ks % STAGES;
            const uint32_t ph = (uint32_t)(ks / STAGES) & 1u;
            if (!mbar_wait(smem_u32(&full_bar[s]), ph)) break;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // B tile: generic st.shared -> UMMA reads
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = tmem_base + ACC_COLS + (uint32_t)s * A_STAGE_COLS, a_lo = a_hi + 32;
            const uint32_t b_hi = tiles + s * STAGE_BYTES, b_lo = b_hi + B_TILE;
            const uint64_t db_hi = make_desc(b_hi), db_lo = make_desc(b_lo);
#pragma unroll
            for (int k4 = 0; k4 < BK / 8; ++k4) {
                const uint64_t adv = (uint64_t)(k4 * 2);
                const uint32_t ac = (uint32_t)(k4 * 8);  // 8 tf32 = 8 TMEM columns
                tc3::mma_tf32_ts(tmem_base, a_hi + ac, db_hi + adv, idesc, (ks | k4) ? 1u : 0u);
                tc3::mma_tf32_ts(tmem_base, a_lo + ac, db_hi + adv, idesc, 1u);
                tc3::mma_tf32_ts(tmem_base, a_hi + ac, db_lo + adv, idesc, 1u);
            }
            mma_commit(smem_u32(&empty_bar[s]));
        }
        if (nks > 0) mma_commit(smem_u32(&accum_bar));
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

}  // namespace bb
