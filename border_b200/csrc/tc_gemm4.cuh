// tc_gemm4.cuh -- tcgen05 3xTF32 tile GEMM with an asynchronous (cp.async) operand pipeline.
//
// Same contraction, operands, tile shape, MMA issue and epilogue as tc_gemm.cuh; what changes is
// how the operand tiles reach shared memory.  tc_gemm.cuh stages them through registers
// (global -> registers -> hi/lo split -> st.shared): the register budget allows one stage of loads
// in flight, and ncu shows the kernel waiting on exactly that (long_scoreboard 40 % of all stalls,
// ~4000 cycles per 32-wide k-slice; profiles/r01_summary.md).  Here every producer thread copies
// ITS chunks of the next DEPTH-1 k-slices with cp.async into a raw staging ring (no registers
// held, DEPTH * 24 KB in flight per CTA), waits for the oldest group, reads its own chunks back,
// splits them into the TF32 hi / lo parts and writes the two swizzled UMMA tiles.  A thread only
// ever reads raw chunks it copied itself, so cp.async.wait_group is the only synchronisation of
// the raw ring; the split tiles keep the mbarrier full/empty handshake with the MMA thread.
#pragma once
#include "tc_gemm.cuh"

namespace bb {


namespace tc4 {

using namespace tc;

__device__ __forceinline__ void cp16(uint32_t dst, const void* src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid ? 16u : 0u) : "memory");
}
__device__ __forceinline__ void cp4(uint32_t dst, const void* src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(valid ? 4u : 0u) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

}  // namespace tc4

// STAGES split (hi/lo UMMA) stages, DEPTH raw cp.async stages, MINB co-resident CTAs per SM.
template <int BN, int STAGES, int DEPTH, int MINB, bool A_KSRC, bool B_KSRC, bool A_U8, bool B_U8>
__global__ void __launch_bounds__(tc::NTHREADS, MINB) tc_gemm_async_kernel(GemmArgs g) {
    using namespace tc;
    using namespace tc4;
    constexpr uint32_t A_TILE = BM * 128, B_TILE = BN * 128;       // bytes per hi (or lo) tile
    constexpr uint32_t STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;
    constexpr int A_LD = BM * 8 / NPROD;                             // 16-byte chunks per producer thread per stage (4)
    constexpr int B_LD = (BN * 8 + NPROD - 1) / NPROD;               // 1, 2 or 4
    constexpr uint32_t RAW_STAGE = (uint32_t)(A_LD + B_LD) * NPROD * 16u;
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
    extern __shared__ uint8_t smem_dyn[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int k_begin = blockIdx.z * g.k_per_split;
    const int k_end = min(g.K, k_begin + g.k_per_split);
    const int nks = k_end > k_begin ? (k_end - k_begin + BK - 1) / BK : 0;
    const uint32_t tiles = (smem_u32(smem_dyn) + 1023u) & ~1023u;   // SWIZZLE_128B tiles need 1024 B alignment
    const uint32_t raw = tiles + STAGES * STAGE_BYTES;
    const bool trace0 = (g.fence_mode & 256) && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
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
        // m-contiguous A: a thread owns 4 consecutive m (rows (warp + 8 i) * 4 ..) at k = k0 + lane
        long a_moff[A_LD];
        if (!A_KSRC) {
#pragma unroll
            for (int i = 0; i < A_LD; ++i) {
                int m = m0 + (warp + 8 * i) * 4;
                a_moff[i] = m < g.M ? (g.a_rowbase ? (long)g.a_rowbase[m] : (long)m) : 0;
            }
        }
        long b_noff_r[B_LD];
        if (!B_KSRC) {
#pragma unroll
            for (int i = 0; i < B_LD; ++i) {
                int n = n0 + (warp + 8 * i) * 4;
                b_noff_r[i] = (n < g.N) ? (g.b_noff ? (long)g.b_noff[n] : (long)n) : 0;
            }
        }

        // per-stage gather table entries: a_koff[k] (k-contiguous gather A) / b_rowbase[k] (n-contiguous B)
        auto load_tab = [&](long& oa, long& ob, int ks) {
            const int k0 = k_begin + ks * BK;
            oa = 0; ob = 0;
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

        // issue the cp.async copies of k-slice ks into raw stage ks % DEPTH
        auto issue = [&](int ks, long tabA, long tabB) {
            if (g.fence_mode & 16) return;  // debug: no global loads
            const int k0 = k_begin + ks * BK;
            const uint32_t rs = raw + (uint32_t)(ks % DEPTH) * RAW_STAGE + (uint32_t)tid * 16u;
#pragma unroll
            for (int i = 0; i < A_LD; ++i) {
                const uint32_t dst = rs + (uint32_t)i * (NPROD * 16u);
                if (A_KSRC) {
                    const int k = k0 + (tid & 7) * 4;
                    const bool valid = a_base[i] >= 0 && k < k_end;
                    const long off = valid ? a_base[i] + tabA : 0;
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
                } else {
                    const int k = k0 + lane;
                    const int m = m0 + (warp + 8 * i) * 4;
                    const bool valid = k < k_end && m < g.M;
                    const long off = valid ? tabA + a_moff[i] : 0;
                    if (A_U8) {
                        cp4(dst, Au + off, valid);
                    } else if (!valid || (a_vec && m + 3 < g.M && ((off & 3) == 0))) {
                        cp16(dst, Af + off, valid);
                    } else {
                        cp4(dst, Af + off, true);
#pragma unroll
                        for (int j = 1; j < 4; ++j) {
                            const bool vj = m + j < g.M;
                            const long d = vj ? (g.a_rowbase ? (long)(g.a_rowbase[m + j] - g.a_rowbase[m]) : (long)j) : 0;
                            cp4(dst + 4u * j, Af + off + d, vj);
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < B_LD; ++i) {
                const uint32_t dst = rs + (uint32_t)(A_LD + i) * (NPROD * 16u);
                if (B_KSRC) {
                    const int k = k0 + (tid & 7) * 4;
                    const bool valid = b_base[i] >= 0 && k < k_end;
                    const long off = valid ? b_base[i] + k : 0;
                    if (!valid || (b_vec && k + 3 < k_end)) {
                        cp16(dst, Bf + off, valid);
                    } else {
                        cp4(dst, Bf + off, true);
#pragma unroll
                        for (int j = 1; j < 4; ++j) cp4(dst + 4u * j, Bf + off + (k + j < k_end ? j : 0), k + j < k_end);
                    }
                } else {
                    const int k = k0 + lane;
                    const int n4 = warp + 8 * i;
                    const int n = n0 + n4 * 4;
                    const bool valid = n4 * 4 < BN && k < k_end && n < g.N;
                    const long off = valid ? tabB + b_noff_r[i] : 0;
                    if (B_U8) {
                        cp4(dst, Bu + off, valid);
                    } else if (!valid || (b_vec && n + 3 < g.N && ((off & 3) == 0))) {
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

        bool alive = true;
        // raw stage ks % DEPTH (own chunks) -> hi / lo split -> UMMA stage ks % STAGES
        const bool trace = (g.fence_mode & 256) && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
        auto convert = [&](int ks) {
            if (trace && ks < 64) g_tc_trace[0][ks][1] = clock64();
            const int k0 = k_begin + ks * BK;
            const uint32_t rs = raw + (uint32_t)(ks % DEPTH) * RAW_STAGE + (uint32_t)tid * 16u;
            float4 pa[A_LD], pb[B_LD];
#pragma unroll
            for (int i = 0; i < A_LD; ++i) {
                const uint32_t src = rs + (uint32_t)i * (NPROD * 16u);
                if (A_U8) {
                    float4 v = u8x4_to_float4(lds32(src));
                    if (A_KSRC) {
                        const int k = k0 + (tid & 7) * 4;
                        if (k + 1 >= k_end) v.y = 0.f;
                        if (k + 2 >= k_end) v.z = 0.f;
                        if (k + 3 >= k_end) v.w = 0.f;
                    } else {
                        const int m = m0 + (warp + 8 * i) * 4;
                        if (m + 1 >= g.M) v.y = 0.f;
                        if (m + 2 >= g.M) v.z = 0.f;
                        if (m + 3 >= g.M) v.w = 0.f;
                    }
                    pa[i] = v;
                } else {
                    pa[i] = lds128(src);
                }
            }
#pragma unroll
            for (int i = 0; i < B_LD; ++i) {
                const uint32_t src = rs + (uint32_t)(A_LD + i) * (NPROD * 16u);
                if (!B_KSRC && B_U8) pb[i] = u8x4_to_float4(lds32(src));
                else pb[i] = lds128(src);
            }
            const int s = ks % STAGES;
            const uint32_t ph = (uint32_t)(ks / STAGES) & 1u;
            if (alive && !mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u)) alive = false;
            if (trace && ks < 64) g_tc_trace[0][ks][2] = clock64();
            const uint32_t a_hi = tiles + s * STAGE_BYTES, a_lo = a_hi + A_TILE, b_hi = a_lo + A_TILE, b_lo = b_hi + B_TILE;
            if (g.fence_mode & 64) {  // debug: no split stores
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&full_bar[s]));
                return;
            }
#pragma unroll
            for (int i = 0; i < A_LD; ++i) {
                if (A_KSRC) {
                    uint32_t off = sw128((uint32_t)(tid >> 3) + 32u * i, (uint32_t)(tid & 7));
                    split_store(a_hi + off, a_lo + off, pa[i]);
                } else {
                    uint32_t r = (uint32_t)(warp + 8 * i) * 4u;
                    const float v[4] = {pa[i].x, pa[i].y, pa[i].z, pa[i].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t off = sw128(r + j, (uint32_t)lane >> 2) + ((uint32_t)lane & 3u) * 4u;
                        split_store1(a_hi + off, a_lo + off, v[j]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < B_LD; ++i) {
                if (B_KSRC) {
                    uint32_t r = (uint32_t)(tid >> 3) + 32u * i;
                    if (r < (uint32_t)BN) {
                        uint32_t off = sw128(r, (uint32_t)(tid & 7));
                        split_store(b_hi + off, b_lo + off, pb[i]);
                    }
                } else {
                    uint32_t r = (uint32_t)(warp + 8 * i) * 4u;
                    if (r < (uint32_t)BN) {
                        const float v[4] = {pb[i].x, pb[i].y, pb[i].z, pb[i].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint32_t off = sw128(r + j, (uint32_t)lane >> 2) + ((uint32_t)lane & 3u) * 4u;
                            split_store1(b_hi + off, b_lo + off, v[j]);
                        }
                    }
                }
            }
            // the generic->async proxy fence runs in the MMA thread after it acquires the full barrier
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&full_bar[s]));
            if (trace && ks < 64) g_tc_trace[0][ks][3] = clock64();
        };

        // prologue: DEPTH-1 slices in flight; the table entries run one slice ahead of their issue
        long ta, tb;
        load_tab(ta, tb, 0);
#pragma unroll
        for (int j = 0; j < DEPTH - 1; ++j) {
            if (j < nks) {
                long na = 0, nb = 0;
                if (j + 1 < nks) load_tab(na, nb, j + 1);
                issue(j, ta, tb);
                ta = na; tb = nb;
            }
            cp_commit();
        }
        for (int ks = 0; ks < nks; ++ks) {
            const int nx = ks + DEPTH - 1;
            if (nx < nks) {
                long na = 0, nb = 0;
                if (nx + 1 < nks) load_tab(na, nb, nx + 1);
                issue(nx, ta, tb);
                ta = na; tb = nb;
            }
            cp_commit();
            if (trace && ks < 64) g_tc_trace[0][ks][0] = clock64();
            cp_wait<DEPTH - 1>();
            convert(ks);
        }

        // ================================================================ epilogue
        if (trace0) g_tc_trace[1][58][0] = clock64();
        if (nks > 0 && alive) alive = mbar_wait(smem_u32(&accum_bar), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (trace0) g_tc_trace[1][59][0] = clock64();
        const bool direct = g.split_k <= 1;
        float* out = direct ? g.C : g.workspace + (size_t)blockIdx.z * g.M * g.N;
        const int ldo = direct ? g.ldc : (g.trans_out ? g.M : g.N);
        const int q = warp & 3;                     // TMEM lane quarter this warp may read
        const int m = m0 + q * 32 + lane;
        constexpr int HALF = BN / 2 < 16 ? 16 : BN / 2;   // columns per warp-pair member
        const int c_begin = (warp >> 2) * HALF;
        if (c_begin < BN) {
#pragma unroll
            for (int c0 = 0; c0 < HALF; c0 += 16) {
                const int col = c_begin + c0;
                if (col >= BN) break;
                uint32_t r[16];
                if (nks > 0 && alive) {
                    uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col;
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
                        "%14, %15}, [%16];\n"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
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
                        if (n >= g.N) break;
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
                            continue;
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
        // ================================================================ MMA issuer (one thread)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        bool alive = true;
        const bool trace = (g.fence_mode & 256) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
        for (int ks = 0; ks < nks && alive; ++ks) {
            const int s = ks % STAGES;
            const uint32_t ph = (uint32_t)(ks / STAGES) & 1u;
            if (trace && ks < 64) g_tc_trace[1][ks][0] = clock64();
            if (!mbar_wait(smem_u32(&full_bar[s]), ph)) { alive = false; break; }
            if (trace && ks < 64) g_tc_trace[1][ks][1] = clock64();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (trace && ks < 64) g_tc_trace[1][ks][2] = clock64();
            const uint32_t a_hi = tiles + s * STAGE_BYTES, a_lo = a_hi + A_TILE, b_hi = a_lo + A_TILE, b_lo = b_hi + B_TILE;
            const uint64_t da_hi = make_desc(a_hi), da_lo = make_desc(a_lo), db_hi = make_desc(b_hi), db_lo = make_desc(b_lo);
#pragma unroll
            for (int k4 = 0; k4 < BK / 8; ++k4) {
                const uint64_t adv = (uint64_t)(k4 * 2);  // 8 tf32 = 32 B = 2 x 16 B along K inside the swizzle atom
                if (g.fence_mode & 32) continue;  // debug: no MMA
                mma_tf32(tmem_base, da_hi + adv, db_hi + adv, idesc, (ks | k4) ? 1u : 0u);
                if (g.fence_mode & 128) continue;  // debug: one pass only
                mma_tf32(tmem_base, da_lo + adv, db_hi + adv, idesc, 1u);
                mma_tf32(tmem_base, da_hi + adv, db_lo + adv, idesc, 1u);
            }
            mma_commit(smem_u32(&empty_bar[s]));  // frees the stage once the MMAs above have read it
            if (trace && ks < 64) g_tc_trace[1][ks][3] = clock64();
        }
        if (nks > 0) mma_commit(smem_u32(&accum_bar));
    }

    if (trace0) g_tc_trace[1][60][0] = clock64();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
    if (trace0) g_tc_trace[1][61][0] = clock64();
}

}  // namespace bb
