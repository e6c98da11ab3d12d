// tc_gemm2.cuh -- persistent, fully pipelined variant of the tcgen05 3xTF32 GEMM (tc_gemm.cuh).
//
// One CTA per SM loops over output tiles (tile = blockIdx.x, += gridDim.x).  Three roles run
// decoupled, connected only by mbarriers, so the load latency of tile t+1, the tensor-core work of
// tile t and the epilogue of tile t-1 overlap:
//
//   warps 0-7   producers: a FLAT software pipeline over all (tile, k-slice) stages of this CTA --
//               gather tables PF stages ahead, operand data PF-1 stages ahead (register sets), then
//               hi/lo split + swizzled st.shared + fence.proxy.async + mbarrier arrive.  The stream
//               never drains at a tile boundary.
//   warp 8      TMEM owner and MMA issuer: tcgen05.mma.kind::tf32 x3 per k-slice into one of TWO
//               accumulators (2*BN TMEM columns), tcgen05.commit to free the smem stage / publish
//               the accumulator.
//   warps 9-12  epilogue: tcgen05.ld the finished accumulator (warp w reads lane quarter w % 4),
//               release it to the MMA warp, then bias / ReLU / mask and global stores.
//
// Same operand patterns, swizzle and descriptors as tc_gemm.cuh (shared helpers in namespace tc).
#pragma once
#include "tc_gemm.cuh"

namespace bb {
namespace tc2 {
constexpr int NPROD = 256, NEPI = 128, NTHREADS = NPROD + 32 + NEPI;

struct Cursor {
    int tile, ks, nks, m0, n0, k_begin, k_end;
    bool valid;
};
}  // namespace tc2

template <int BN, int STAGES, int PF, bool A_KSRC, bool B_KSRC, bool A_U8, bool B_U8>
__global__ void __launch_bounds__(tc2::NTHREADS, 1)
    tc_gemm_persist_kernel(GemmArgs g, int tiles_m, int tiles_n, int total_tiles) {
    using namespace tc;
    using tc2::Cursor;
    constexpr int NPROD = tc2::NPROD;
    constexpr uint32_t A_TILE = BM * 128, B_TILE = BN * 128;
    constexpr uint32_t STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;
    constexpr int A_LD = BM * 8 / NPROD;
    constexpr int B_LD = (BN * 8 + NPROD - 1) / NPROD;
    constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
    extern __shared__ uint8_t smem_dyn[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tiles = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const int mn_tiles = tiles_m * tiles_n;

    auto set_tile = [&](Cursor& c, int tile) {
        c.tile = tile;
        c.valid = tile < total_tiles;
        c.ks = 0;
        if (c.valid) {
            int z = tile / mn_tiles, r = tile - z * mn_tiles;
            int tmi = r / tiles_n, tni = r - tmi * tiles_n;
            c.m0 = tmi * BM; c.n0 = tni * BN;
            c.k_begin = z * g.k_per_split;
            c.k_end = min(g.K, c.k_begin + g.k_per_split);
            c.nks = (c.k_end - c.k_begin + BK - 1) / BK;
        } else {
            c.m0 = c.n0 = c.k_begin = c.k_end = 0; c.nks = 1;
        }
    };
    auto advance = [&](Cursor& c) {
        if (++c.ks >= c.nks) set_tile(c, c.tile + (int)gridDim.x);
    };

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), NPROD / 32);  // one arrive per producer warp
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&acc_full[b]), 1);
            mbar_init(smem_u32(&acc_empty[b]), tc2::NEPI / 32);
        }
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
        const float* Af = reinterpret_cast<const float*>(g.A);
        const uint8_t* Au = reinterpret_cast<const uint8_t*>(g.A);
        const float* Bf = reinterpret_cast<const float*>(g.B);
        const uint8_t* Bu = reinterpret_cast<const uint8_t*>(g.B);
        const bool a_vec = A_U8 || (((g.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0));
        const bool b_vec = B_U8 || (((g.ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0));
        long a_base[A_LD], a_base_n[A_LD];    // row bases of the k-contiguous A operand (current / prefetched tile)
        long b_noff_r[B_LD], b_noff_n[B_LD];  // column offsets of the n-contiguous B operand
        float4 ra[PF][A_LD], rb[PF][B_LD];
        long ta[PF], tb[PF];
        Cursor cT, cD, cS;
        set_tile(cT, (int)blockIdx.x);
        cD = cT; cS = cT;

        // table stream: per-stage gather entries, plus the per-tile bases when a new tile starts
        auto load_tab = [&](long& oa, long& ob, const Cursor& c) {
            const int k0 = c.k_begin + c.ks * BK;
            oa = 0; ob = 0;
            if (A_KSRC) {
                int k = k0 + (tid & 7) * 4;
                if (k < c.k_end) oa = g.a_koff ? (long)__ldg(g.a_koff + k) : (long)k;
            }
            if (!B_KSRC) {
                int k = k0 + lane;
                if (k < c.k_end) ob = g.b_rowbase ? (long)__ldg(g.b_rowbase + k) : (long)k * g.ldb;
            }
            if (c.ks == 0) {
                if (A_KSRC) {
#pragma unroll
                    for (int i = 0; i < A_LD; ++i) {
                        int m = c.m0 + (tid >> 3) + 32 * i;
                        a_base_n[i] = m < g.M ? (g.a_rowbase ? (long)__ldg(g.a_rowbase + m) : (long)m * g.lda) : -1;
                    }
                }
                if (!B_KSRC) {
#pragma unroll
                    for (int i = 0; i < B_LD; ++i) {
                        int n = c.n0 + (warp + 8 * i) * 4;
                        b_noff_n[i] = (n < g.N) ? (g.b_noff ? (long)__ldg(g.b_noff + n) : (long)n) : 0;
                    }
                }
            }
        };

        auto load = [&](float4* pa, float4* pb, const Cursor& c, long tabA, long tabB) {
            if (g.fence_mode & 16) {
#pragma unroll
                for (int i = 0; i < A_LD; ++i) pa[i] = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
                for (int i = 0; i < B_LD; ++i) pb[i] = make_float4(1.f, 1.f, 1.f, 1.f);
                return;
            }
            if (c.ks == 0) {  // the data stream enters a new tile: adopt the prefetched bases
#pragma unroll
                for (int i = 0; i < A_LD; ++i) a_base[i] = a_base_n[i];
#pragma unroll
                for (int i = 0; i < B_LD; ++i) b_noff_r[i] = b_noff_n[i];
            }
            const int k0 = c.k_begin + c.ks * BK;
            const int k_end = c.k_end;
#pragma unroll
            for (int i = 0; i < A_LD; ++i) {
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
                } else {
                    int k = k0 + lane;
                    int m = c.m0 + (warp + 8 * i) * 4;
                    if (k < k_end && m < g.M) {
                        long off = (long)k * g.lda + m;
                        if (a_vec && m + 3 < g.M) v = __ldg(reinterpret_cast<const float4*>(Af + off));
                        else {
                            v.x = __ldg(Af + off);
                            if (m + 1 < g.M) v.y = __ldg(Af + off + 1);
                            if (m + 2 < g.M) v.z = __ldg(Af + off + 2);
                            if (m + 3 < g.M) v.w = __ldg(Af + off + 3);
                        }
                    }
                }
                pa[i] = v;
            }
#pragma unroll
            for (int i = 0; i < B_LD; ++i) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (B_KSRC) {
                    int k = k0 + (tid & 7) * 4;
                    int r = (tid >> 3) + 32 * i;
                    int n = c.n0 + r;
                    if (r < BN && n < g.N && k < k_end) {
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
                    int n4 = warp + 8 * i;
                    int n = c.n0 + n4 * 4;
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
                pb[i] = v;
            }
        };

        bool alive = true;
        uint32_t gs = 0;  // stages stored so far by this CTA
        auto store = [&](const float4* pa, const float4* pb) {
            const uint32_t s = gs % STAGES;
            const uint32_t ph = (gs / STAGES) & 1u;
            if (alive && !mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u)) alive = false;
            const uint32_t a_hi = tiles + s * STAGE_BYTES, a_lo = a_hi + A_TILE, b_hi = a_lo + A_TILE, b_lo = b_hi + B_TILE;
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
            if ((g.fence_mode & 1) == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();  // orders the warp's st.shared before lane 0's release-arrive (256 arrives/stage were costly)
            if (lane == 0) mbar_arrive(smem_u32(&full_bar[s]));
            gs += 1;
        };

        // prologue: tables for stages 0..PF-1, data for stages 0..PF-2 (data lags tables by one stage)
#pragma unroll
        for (int j = 0; j < PF; ++j) {
            if (j >= 1 && cD.valid) { load(ra[j - 1], rb[j - 1], cD, ta[j - 1], tb[j - 1]); advance(cD); }
            if (cT.valid) { load_tab(ta[j], tb[j], cT); advance(cT); }
        }
        while (cS.valid) {
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                if (cS.valid) {
                    if (cD.valid) {
                        load(ra[(u + PF - 1) % PF], rb[(u + PF - 1) % PF], cD, ta[(u + PF - 1) % PF], tb[(u + PF - 1) % PF]);
                        advance(cD);
                    }
                    if (cT.valid) { load_tab(ta[u], tb[u], cT); advance(cT); }
                    store(ra[u], rb[u]);
                    advance(cS);
                }
            }
        }
    } else if (warp == 8) {
        // ================================================================ MMA issuer (one thread)
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            Cursor c;
            set_tile(c, (int)blockIdx.x);
            uint32_t gs = 0, lt = 0;  // global stage counter, local tile counter
            bool alive = true;
            while (c.valid && alive) {
                const uint32_t buf = lt & 1u;
                // the epilogue must have drained this accumulator (two tiles ago)
                if (!mbar_wait(smem_u32(&acc_empty[buf]), ((lt >> 1) & 1u) ^ 1u)) { alive = false; break; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + buf * BN;
                const int nks = c.nks;
                for (int ks = 0; ks < nks; ++ks) {
                    const uint32_t s = gs % STAGES;
                    const uint32_t ph = (gs / STAGES) & 1u;
                    if (!mbar_wait(smem_u32(&full_bar[s]), ph)) { alive = false; break; }
                    if ((g.fence_mode & 1) == 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = tiles + s * STAGE_BYTES, a_lo = a_hi + A_TILE, b_hi = a_lo + A_TILE, b_lo = b_hi + B_TILE;
                    const uint64_t da_hi = make_desc(a_hi), da_lo = make_desc(a_lo), db_hi = make_desc(b_hi), db_lo = make_desc(b_lo);
#pragma unroll
                    for (int k4 = 0; k4 < BK / 8; ++k4) {
                        const uint64_t adv = (uint64_t)(k4 * 2);
                        if (g.fence_mode & 32) continue;
                        mma_tf32(tmem_d, da_hi + adv, db_hi + adv, idesc, (ks | k4) ? 1u : 0u);
                        mma_tf32(tmem_d, da_lo + adv, db_hi + adv, idesc, 1u);
                        mma_tf32(tmem_d, da_hi + adv, db_lo + adv, idesc, 1u);
                    }
                    mma_commit(smem_u32(&empty_bar[s]));
                    gs += 1;
                }
                if (!alive) break;
                mma_commit(smem_u32(&acc_full[buf]));
                lt += 1;
                set_tile(c, c.tile + (int)gridDim.x);
            }
        }
    } else {
        // ================================================================ epilogue warps 9..12
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        Cursor c;
        set_tile(c, (int)blockIdx.x);
        uint32_t lt = 0;
        bool alive = true;
        const bool direct = g.split_k <= 1;
        const int ldo = direct ? g.ldc : g.N;
        while (c.valid) {
            const uint32_t buf = lt & 1u;
            if (alive && !mbar_wait(smem_u32(&acc_full[buf]), (lt >> 1) & 1u)) alive = false;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t r[BN];
#pragma unroll
            for (int c0 = 0; c0 < BN; c0 += 16) {
                uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + (uint32_t)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
                    "%14, %15}, [%16];\n"
                    : "=r"(r[c0 + 0]), "=r"(r[c0 + 1]), "=r"(r[c0 + 2]), "=r"(r[c0 + 3]), "=r"(r[c0 + 4]), "=r"(r[c0 + 5]),
                      "=r"(r[c0 + 6]), "=r"(r[c0 + 7]), "=r"(r[c0 + 8]), "=r"(r[c0 + 9]), "=r"(r[c0 + 10]), "=r"(r[c0 + 11]),
                      "=r"(r[c0 + 12]), "=r"(r[c0 + 13]), "=r"(r[c0 + 14]), "=r"(r[c0 + 15])
                    : "r"(taddr)
                    : "memory");
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));  // accumulator is in registers: hand it back
            const int z = c.tile / mn_tiles;
            float* out = direct ? g.C : g.workspace + (size_t)z * g.M * g.N;
            const int m = c.m0 + q * 32 + lane;
            if (m < g.M) {
#pragma unroll
                for (int j4 = 0; j4 < BN; j4 += 4) {
                    const int n = c.n0 + j4;
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
            lt += 1;
            set_tile(c, c.tile + (int)gridDim.x);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace bb
