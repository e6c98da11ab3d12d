// tc_gemm5.cuh -- tcgen05 3xTF32 tile GEMM: A operand in TENSOR MEMORY, cp.async operand staging.
//
// What the per-stage clock64 trace of tc_gemm4.cuh showed (tools/tc_trace.py, B200): a 128x64x32
// stage costs ~1900 cycles although its 12 MMAs need 384, because (a) 3xTF32 with both operands
// in shared memory moves 24 KB raw + 48 KB hi/lo writes + 72 KB UMMA operand reads per stage
// (168 KB at ~88 B/clk = the shared-memory pipe), and (b) every producer warp walks every stage,
// so the per-stage dependency chain (copy -> wait -> read -> split -> store -> arrive) is never
// overlapped with itself.  This kernel removes both:
//
//   * A (128 rows x 32 k) never touches shared memory as an MMA operand: thread = matrix row =
//     TMEM lane, the thread splits its 32 k-values into TF32 hi / lo in registers and writes them
//     to TMEM with tcgen05.st; tcgen05.mma reads A from TMEM ("[a_tmem]" form).  Only the small
//     B tile (BN x 32) is staged hi / lo in shared memory.  Shared-memory traffic per stage drops
//     to 24 KB raw (cp.async in, ld.shared out) + 16 KB B split + 24 KB UMMA B reads.
//   * two producer groups of 4 warps own alternating k-stages, so one group's copy / split chain
//     overlaps the other's, and two CTAs per SM overlap prologue / epilogue with main loops.
//   * operands arrive through cp.async (no registers held while in flight): each group keeps RD
//     raw buffers; the copy of its next stage is issued as soon as the current one has been read.
//
//   warps 0-3 / 4-7  producer groups (even / odd k-stages); warp w owns TMEM lane quarter w % 4
//   warp 8           TMEM owner + MMA issuer: per k-slice  D += A_hi B_hi + A_lo B_hi + A_hi B_lo
//   epilogue         warps 0-7 (lane quarter w % 4, column half w / 4)
//
// TMEM columns: [0, 64) accumulator, then S stages x (32 hi + 32 lo).
#pragma once
#include "tc_gemm.cuh"
#include "tc_gemm3.cuh"
#include "tc_gemm4.cuh"

namespace bb {
namespace tc5 {
constexpr int NTHREADS = 288, ACC_COLS = 64, A_STAGE_COLS = 64;
constexpr uint32_t A_RAW = 128 * 128;  // bytes of one raw A buffer (128 rows x 32 floats)
__host__ __device__ constexpr uint32_t tmem_cols(int S) { return (ACC_COLS + S * A_STAGE_COLS) <= 256 ? 256u : 512u; }
__host__ __device__ constexpr size_t smem_bytes(int BN, int S, int RD) {
    return (size_t)S * 2 * BN * 128 + (size_t)2 * RD * (A_RAW + BN * 128) + 1024;
}
__device__ __forceinline__ void group_bar(int grp) {
    asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
}
}  // namespace tc5

// S: TMEM-A / shared-B stages; RD: raw cp.async buffers per producer group; MINB: CTAs per SM.
template <int BN, int S, int RD, int MINB, bool A_KSRC, bool B_KSRC, bool A_U8>
__global__ void __launch_bounds__(tc5::NTHREADS, MINB) tc_gemm_ta_kernel(GemmArgs g) {
    using namespace tc;
    using namespace tc4;
    using namespace tc5;
    static_assert(BN == 32 || BN == 64, "BN <= 64");
    static_assert(!A_U8 || A_KSRC, "u8 A is k-contiguous only");
    constexpr uint32_t B_TILE = BN * 128;              // bytes of one hi (or lo) B tile = bytes of one raw B tile
    constexpr uint32_t STAGE_BYTES = 2 * B_TILE;
    constexpr uint32_t RAW_BUF = A_RAW + B_TILE;
    constexpr int B_LD = BN * 8 / 128;                  // 16-byte chunks per thread of a 128-thread group (2 or 4)
    constexpr uint32_t TMEM_COLS = tmem_cols(S);
    extern __shared__ uint8_t smem_dyn[];
    __shared__ uint64_t full_bar[S], empty_bar[S], accum_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int k_begin = blockIdx.z * g.k_per_split;
    const int k_end = min(g.K, k_begin + g.k_per_split);
    const int nks = k_end > k_begin ? (k_end - k_begin + BK - 1) / BK : 0;
    const uint32_t tiles = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const uint32_t raw0 = tiles + S * STAGE_BYTES;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 4);   // the 4 warps of one producer group
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

    if (warp < 8) {
        // ================================================================ producers
        const int grp = warp >> 2;                 // 0: even stages, 1: odd stages
        const int q = warp & 3;                    // TMEM lane quarter
        const int row = q * 32 + lane;             // A row owned by this thread
        const int gt = tid & 127;                  // thread index inside the group
        const float* Af = reinterpret_cast<const float*>(g.A);
        const uint8_t* Au = reinterpret_cast<const uint8_t*>(g.A);
        const float* Bf = reinterpret_cast<const float*>(g.B);
        const bool a_vec = A_U8 || (((g.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0));
        const bool b_vec = ((g.ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0);
        const int m = m0 + row;
        // k-contiguous A: this thread copies (and later reads) its own row
        long a_base = -1;
        if (A_KSRC && m < g.M) a_base = g.a_rowbase ? (long)__ldg(g.a_rowbase + m) : (long)m * g.lda;
        // m-contiguous A: the group copies [32 k][128 m]; this thread's chunks are 4 m at m0 + 4*(gt & 31)
        const int am = m0 + 4 * (gt & 31);
        long a_moff = 0;
        if (!A_KSRC && am < g.M) a_moff = g.a_rowbase ? (long)__ldg(g.a_rowbase + am) : (long)am;
        long b_noff_r[B_LD];
        if (!B_KSRC) {
#pragma unroll
            for (int i = 0; i < B_LD; ++i) {
                int n = n0 + (q + 4 * i) * 4;
                b_noff_r[i] = (n < g.N) ? (g.b_noff ? (long)__ldg(g.b_noff + n) : (long)n) : 0;
            }
        }
        const uint32_t raw_grp = raw0 + (uint32_t)grp * (RD * RAW_BUF);

        // cp.async the operands of k-slice ks into raw buffer `rb` of this group
        auto issue = [&](int ks, int rb) {
            const int k0 = k_begin + ks * BK;
            const uint32_t ra = raw_grp + (uint32_t)rb * RAW_BUF, rbB = ra + A_RAW;
            if (A_KSRC) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int k = k0 + c * 4;
                    const bool valid = a_base >= 0 && k < k_end;
                    const uint32_t dst = ra + sw128((uint32_t)row, (uint32_t)c);
                    long off = 0;
                    if (valid) off = a_base + (g.a_koff ? (long)__ldg(g.a_koff + k) : (long)k);
                    if (A_U8) {
                        cp4(dst, Au + off, valid);
                    } else if (!valid || (a_vec && k + 3 < k_end && ((off & 3) == 0))) {
                        cp16(dst, Af + off, valid);
                    } else {
                        cp4(dst, Af + off, true);
#pragma unroll
                        for (int j = 1; j < 4; ++j) {
                            const bool vj = k + j < k_end;
                            const long d = vj ? (g.a_koff ? (long)(g.a_koff[k + j] - g.a_koff[k]) : (long)j) : 0;
                            cp4(dst + 4u * j, Af + off + d, vj);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int kk = (gt >> 5) + 4 * i;   // 0..31
                    const int k = k0 + kk;
                    const bool valid = k < k_end && am < g.M;
                    const uint32_t dst = ra + (uint32_t)kk * 512u + (uint32_t)(gt & 31) * 16u;
                    long off = 0;
                    if (valid) off = (g.a_koff ? (long)__ldg(g.a_koff + k) : (long)k * g.lda) + a_moff;
                    if (!valid || (a_vec && am + 3 < g.M && ((off & 3) == 0))) {
                        cp16(dst, Af + off, valid);
                    } else {
                        cp4(dst, Af + off, true);
#pragma unroll
                        for (int j = 1; j < 4; ++j) {
                            const bool vj = am + j < g.M;
                            const long d = vj ? (g.a_rowbase ? (long)(g.a_rowbase[am + j] - g.a_rowbase[am]) : (long)j) : 0;
                            cp4(dst + 4u * j, Af + off + d, vj);
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < B_LD; ++i) {
                const uint32_t dst = rbB + (uint32_t)(gt + 128 * i) * 16u;
                if (B_KSRC) {
                    const int k = k0 + (gt & 7) * 4;
                    const int n = n0 + (gt >> 3) + 16 * i;
                    const bool valid = n < g.N && k < k_end;
                    const long off = valid ? (long)n * g.ldb + k : 0;
                    if (!valid || (b_vec && k + 3 < k_end)) {
                        cp16(dst, Bf + off, valid);
                    } else {
                        cp4(dst, Bf + off, true);
#pragma unroll
                        for (int j = 1; j < 4; ++j) cp4(dst + 4u * j, Bf + off + (k + j < k_end ? j : 0), k + j < k_end);
                    }
                } else {
                    const int k = k0 + lane;
                    const int n = n0 + (q + 4 * i) * 4;
                    const bool valid = k < k_end && n < g.N;
                    long off = 0;
                    if (valid) off = (g.b_rowbase ? (long)__ldg(g.b_rowbase + k) : (long)k * g.ldb) + b_noff_r[i];
                    if (!valid || (b_vec && n + 3 < g.N && ((off & 3) == 0))) {
                        cp16(dst, Bf + off, valid);
                    } else {
                        cp4(dst, Bf + off, true);
#pragma unroll
                        for (int j = 1; j < 4; ++j) {
                            const bool vj = n + j < g.N;
                            const long d = vj ? (g.b_noff ? (long)(g.b_noff[n + j] - g.b_noff[n]) : (long)j) : 0;
                            cp4(dst + 4u * j, Bf + off + d, vj);
                        }
                    }
                }
            }
        };

        // prologue: the first RD stages of this group
#pragma unroll
        for (int j = 0; j < RD; ++j) {
            if (grp + 2 * j < nks) issue(grp + 2 * j, j);
            cp_commit();
        }
        bool alive = true;
        int it = 0;
        const bool trace = (g.fence_mode & 256) && gt == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
        for (int ks = grp; ks < nks; ks += 2, ++it) {
            const int k0 = k_begin + ks * BK;
            const int rb = it % RD;
            const uint32_t ra = raw_grp + (uint32_t)rb * RAW_BUF, rbB = ra + A_RAW;
            if (trace && ks < 64) g_tc_trace[0][ks][0] = clock64();
            cp_wait<RD - 1>();
            if (!A_KSRC) group_bar(grp);   // the [k][m] tile was copied by the whole group
            if (trace && ks < 64) g_tc_trace[0][ks][1] = clock64();
            const int s = ks % S;
            const uint32_t ph = (uint32_t)(ks / S) & 1u;
            if (alive && !mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u)) alive = false;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (trace && ks < 64) g_tc_trace[0][ks][2] = clock64();
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + ACC_COLS + (uint32_t)s * A_STAGE_COLS;
            // ---------------- A: 32 k-values of this thread's row -> hi / lo -> TMEM
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float av[16];
                if (A_KSRC) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t src = ra + sw128((uint32_t)row, (uint32_t)(h * 4 + c));
                        float4 v;
                        if (A_U8) {
                            v = u8x4_to_float4(lds32(src));
                            const int k = k0 + (h * 4 + c) * 4;
                            if (k + 1 >= k_end) v.y = 0.f;
                            if (k + 2 >= k_end) v.z = 0.f;
                            if (k + 3 >= k_end) v.w = 0.f;
                        } else {
                            v = lds128(src);
                        }
                        av[c * 4 + 0] = v.x; av[c * 4 + 1] = v.y; av[c * 4 + 2] = v.z; av[c * 4 + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        av[j] = __uint_as_float(lds32(ra + (uint32_t)(h * 16 + j) * 512u + (uint32_t)row * 4u));
                }
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    hi[j] = __float_as_uint(av[j]) & 0xffffe000u;
                    lo[j] = __float_as_uint(av[j] - __uint_as_float(hi[j]));
                }
                tc3::tmem_st16(ta + h * 16, hi);
                tc3::tmem_st16(ta + 32 + h * 16, lo);
            }
            if (trace && ks < 64) g_tc_trace[0][ks][4] = clock64();
            // ---------------- B: own raw chunks -> hi / lo -> swizzled UMMA tiles
            const uint32_t b_hi = tiles + s * STAGE_BYTES, b_lo = b_hi + B_TILE;
#pragma unroll
            for (int i = 0; i < B_LD; ++i) {
                const float4 v = lds128(rbB + (uint32_t)(gt + 128 * i) * 16u);
                if (B_KSRC) {
                    uint32_t off = sw128((uint32_t)(gt >> 3) + 16u * i, (uint32_t)(gt & 7));
                    split_store(b_hi + off, b_lo + off, v);
                } else {
                    uint32_t r = (uint32_t)(q + 4 * i) * 4u;
                    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t off = sw128(r + j, (uint32_t)lane >> 2) + ((uint32_t)lane & 3u) * 4u;
                        split_store1(b_hi + off, b_lo + off, vv[j]);
                    }
                }
            }
            // ---------------- refill this raw buffer with the group's stage RD ahead
            if (trace && ks < 64) g_tc_trace[0][ks][5] = clock64();
            if (!A_KSRC) group_bar(grp);   // everyone has read the shared [k][m] tile
            if (ks + 2 * RD < nks) issue(ks + 2 * RD, rb);
            cp_commit();
            if (trace && ks < 64) g_tc_trace[0][ks][6] = clock64();
            // ---------------- publish
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            if (trace && ks < 64) g_tc_trace[0][ks][7] = clock64();
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&full_bar[s]));
            if (trace && ks < 64) g_tc_trace[0][ks][3] = clock64();
        }
        cp_wait<0>();

        // ================================================================ epilogue
        if (nks > 0 && alive) alive = mbar_wait(smem_u32(&accum_bar), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const bool direct = g.split_k <= 1;
        float* out = direct ? g.C : g.workspace + (size_t)blockIdx.z * g.M * g.N;
        const int ldo = direct ? g.ldc : (g.trans_out ? g.M : g.N);
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
                        if (g.trans_out) {  // C^T: consecutive lanes (rows m) write consecutive addresses
                            for (int j = 0; j < 4; ++j)
                                if (n + j < g.N) out[(size_t)(n + j) * ldo + m] = v[j];
                        } else {
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
        }
    } else if (lane == 0) {
        // ================================================================ MMA issuer
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        const bool trace = (g.fence_mode & 256) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
        for (int ks = 0; ks < nks; ++ks) {
            const int s = ks % S;
            const uint32_t ph = (uint32_t)(ks / S) & 1u;
            if (trace && ks < 64) g_tc_trace[1][ks][0] = clock64();
            if (!mbar_wait(smem_u32(&full_bar[s]), ph)) break;
            if (trace && ks < 64) g_tc_trace[1][ks][1] = clock64();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // B tile: generic st.shared -> UMMA reads
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (trace && ks < 64) g_tc_trace[1][ks][2] = clock64();
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
            if (trace && ks < 64) g_tc_trace[1][ks][3] = clock64();
        }
        if (nks > 0) mma_commit(smem_u32(&accum_bar));
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace bb
