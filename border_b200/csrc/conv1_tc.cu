// conv1_tc.cu -- the first AtariCnn layer (cnn/base.rs:26-28: x.float()/255 -> conv2d(C->32, k8, s4)
// -> ReLU) on the u8 frame stack of the replay batch, as a persistent tcgen05 kernel.
//
// Why a dedicated kernel: through the generic implicit GEMM (tc_gemm.cuh) this layer is the most
// expensive forward of the step (46 us at B=256, twice per step) although it is 0.84 GFLOP over
// 7 MB of input: the generic producer gathers the u8 patch with scattered 4-byte requests, splits
// every value into TF32 hi/lo and feeds three MMA passes from shared memory.  Here:
//
//   * u8 pixels are EXACT in TF32 (8 significant bits), so A needs no lo part; the 1/255 moves
//     into the weights (W' = W/255, split hi/lo once per CTA):  y = x (W'_hi + W'_lo)  is two MMA
//     passes instead of three and differs from (x/255) W by one fp32 rounding of W/255.
//   * thread = output pixel = TMEM lane: a thread reads its 8x8xC patch with 4-byte loads that
//     are contiguous across the warp (adjacent ow are 4 bytes apart), widens each byte with one
//     PRMT + one FADD, and writes the values straight into tensor memory (tcgen05.st); the MMA
//     takes A from TMEM.  No shared-memory staging of A at all.
//   * the whole weight matrix (32 x 64C, hi and lo, 16C KB) lives in shared memory for the
//     lifetime of the CTA as K-major SWIZZLE_128B tiles; CTAs are persistent (one per SM) and walk
//     the 128-row tiles, with double-buffered accumulators so the epilogue of tile t overlaps the
//     main loop of tile t+1.
//
//   warps 0-3 / 4-7   producer groups: channel stages alternate between them (lane quarter w % 4)
//   warps 8-11        epilogue: TMEM -> +bias -> ReLU -> NHWC rows (128 B per thread)
//   warp 12           TMEM owner + MMA issuer
//
// TMEM: [0,32) [32,64) accumulators, then S stages x 64 columns (one channel = 64 k-values).
#include <stdlib.h>
#include <atomic>
#include <algorithm>
#include <vector>
#include "nn.cuh"
#include "tc_gemm.cuh"

namespace bb {
static __device__ int g_tc_error_c1 = 0;
static __device__ int* g_err_flag_c1 = nullptr;   // the library's pinned, mapped failure flag (common.cuh), set per device at first launch

namespace c1 {
using namespace tc;
constexpr int NTHREADS = 13 * 32, S = 3, OC = 32, TMEM_COLS = 256, A_COL0 = 64, A_STAGE = 64;

struct Args {
    const uint8_t* X;      // [B][C][H][W] u8
    const float* Wt;       // [32][C*64] (OIHW)
    const float* bias;     // [32]
    float* Y;              // [M][32] NHWC
    long y_plane;          // != 0: also store lo(Y) = Y - tf32_trunc(Y) at Y + y_plane (operand plane of the next layer's TMA loads)
    const int* rowbase;    // [M]
    // != null: image b of the batch is row ix[b] of X (the replay ring itself, base.rs:388-395 gather fused into the loader:
    // no materialised batch); OHW = output positions per image
    const unsigned long long* ix;
    int OHW;
    int M, C, HW, W, relu;
    int n_tiles;
    int dbg;   // bench bisect: 1 no MMA, 2 no global loads, 4 no TMEM stores
};

__device__ __forceinline__ bool wait_bar(uint32_t bar, uint32_t parity) {
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 22); ++it)
        if (mbar_try_wait(bar, parity)) return true;
    atomicExch(&g_tc_error_c1, 1);
    if (g_err_flag_c1) { *reinterpret_cast<volatile int*>(g_err_flag_c1) = 16; __threadfence_system(); }
    return false;
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// byte i of w -> exact float: (0x4B000000 | b) is 8388608 + b
__device__ __forceinline__ uint32_t widen(uint32_t w, int i) {
    uint32_t v = __byte_perm(w, 0x4B000000u, 0x7540u | (uint32_t)i);
    return __float_as_uint(__uint_as_float(v) - 8388608.0f);
}

__global__ void __launch_bounds__(NTHREADS, 2) conv1_fwd_kernel(Args g) {
    extern __shared__ uint8_t smem_dyn[];
    __shared__ uint64_t full_bar[S], empty_bar[S], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = g.C * 64, nslice = K / 32;
    const uint32_t tiles = (smem_u32(smem_dyn) + 1023u) & ~1023u;   // hi slices, then lo slices (4 KB each)
    const uint32_t lo0 = tiles + (uint32_t)nslice * 4096u;
    pdl_sync();

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(smem_u32(&full_bar[s]), 4); mbar_init(smem_u32(&empty_bar[s]), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&acc_full[b]), 1); mbar_init(smem_u32(&acc_empty[b]), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // W' = W / 255 -> hi / lo, K-major SWIZZLE_128B slices of 32 k: element (n, k) of slice k / 32
    for (int i = tid; i < OC * K / 4; i += NTHREADS) {
        const int n = i / (K / 4), k = (i % (K / 4)) * 4;
        float4 v = __ldg(reinterpret_cast<const float4*>(g.Wt + (size_t)n * K + k));
        const float s = 1.0f / 255.0f;
        v.x = v.x * s; v.y = v.y * s; v.z = v.z * s; v.w = v.w * s;   // the reference divides x by 255: a multiply by
        const uint32_t off = (uint32_t)(k >> 5) * 4096u + sw128((uint32_t)n, (uint32_t)((k & 31) >> 2));
        split_store(tiles + off, lo0 + off, v);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic st.shared -> UMMA reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp < 8) {
        // ================================================================ producers
        const int grp = warp >> 2, q = warp & 3, row = q * 32 + lane;
        bool alive = true;
        // this group's stages are n = grp, grp + 2, ... of the CTA's (tile, channel) sequence; the patch
        // words of stage n + 2 are requested before stage n is converted (register prefetch)
        const uint32_t my_tiles = (g.n_tiles > (int)blockIdx.x) ? (uint32_t)(g.n_tiles - 1 - (int)blockIdx.x) / gridDim.x + 1u : 0u;
        const uint32_t n_end = my_tiles * (uint32_t)g.C;
        auto load = [&](uint32_t n, uint32_t* w) {
            const int t = (int)blockIdx.x + (int)(n / (uint32_t)g.C) * (int)gridDim.x;
            const int c = (int)(n % (uint32_t)g.C);
            const int m = t * 128 + row;
            if (m < g.M && !(g.dbg & 2)) {
                const uint8_t* pc = g.X + (size_t)__ldg(g.rowbase + m) + (size_t)c * g.HW;
                if (g.ix) {
                    const int b = m / g.OHW;
                    pc += ((long long)__ldg(g.ix + b) - (long long)b) * ((long long)g.C * g.HW);
                }
#pragma unroll
                for (int kh = 0; kh < 8; ++kh) {
                    const uint32_t* pr = reinterpret_cast<const uint32_t*>(pc + kh * g.W);
                    w[2 * kh] = __ldg(pr);
                    w[2 * kh + 1] = __ldg(pr + 1);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) w[j] = 0u;
            }
        };
        uint32_t w[16], wn[16];
        if ((uint32_t)grp < n_end) load((uint32_t)grp, w);
        for (uint32_t n = (uint32_t)grp; n < n_end; n += 2) {
            if (n + 2 < n_end) load(n + 2, wn);
            const uint32_t s = n % S, ph = (n / S) & 1u;
            if (alive && !wait_bar(smem_u32(&empty_bar[s]), ph ^ 1u)) alive = false;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + A_COL0 + s * A_STAGE;
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                uint32_t f[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) f[j * 4 + i] = widen(w[h * 4 + j], i);
                }
                if (!(g.dbg & 4)) tc3::tmem_st16(ta + h * 16, f);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&full_bar[s]));
#pragma unroll
            for (int j = 0; j < 16; ++j) w[j] = wn[j];
        }
    } else if (warp < 12) {
        // ================================================================ epilogue
        const int q = warp & 3, row = q * 32 + lane;
        uint32_t it = 0;
        bool alive = true;
        for (int t = blockIdx.x; t < g.n_tiles; t += gridDim.x, ++it) {
            const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
            if (alive && !wait_bar(smem_u32(&acc_full[buf]), ph)) alive = false;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * 32u, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
            const int m = t * 128 + row;
            if (m < g.M && alive) {
                float4* dst = reinterpret_cast<float4*>(g.Y + (size_t)m * OC);
#pragma unroll
                for (int j = 0; j < OC; j += 4) {
                    const float4 bj = g.bias ? __ldg(reinterpret_cast<const float4*>(g.bias + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 v;
                    v.x = __uint_as_float(r[j]) + bj.x; v.y = __uint_as_float(r[j + 1]) + bj.y;
                    v.z = __uint_as_float(r[j + 2]) + bj.z; v.w = __uint_as_float(r[j + 3]) + bj.w;
                    if (g.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    dst[j >> 2] = v;
                    if (g.y_plane)
                        reinterpret_cast<float4*>(g.Y + g.y_plane + (size_t)m * OC)[j >> 2] =
                            make_float4(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u), v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u),
                                        v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u), v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u));
                }
            }
        }
    } else if (lane == 0) {
        // ================================================================ MMA issuer
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(OC >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        uint32_t n = 0, it = 0;
        bool alive = true;
        for (int t = blockIdx.x; t < g.n_tiles && alive; t += gridDim.x, ++it) {
            const uint32_t buf = it & 1u, aph = (it >> 1) & 1u;
            if (!wait_bar(smem_u32(&acc_empty[buf]), aph ^ 1u)) { alive = false; break; }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d = tmem_base + buf * 32u;
            for (int c = 0; c < g.C; ++c, ++n) {
                const uint32_t s = n % S, ph = (n / S) & 1u;
                if (!wait_bar(smem_u32(&full_bar[s]), ph)) { alive = false; break; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = tmem_base + A_COL0 + s * A_STAGE;
#pragma unroll
                for (int k8 = 0; k8 < 8; ++k8) {
                    const uint32_t sl = (uint32_t)(c * 2 + (k8 >> 2)) * 4096u;
                    const uint64_t adv = (uint64_t)((k8 & 3) * 2);
                    const uint64_t dhi = make_desc(tiles + sl) + adv, dlo = make_desc(lo0 + sl) + adv;
                    if (g.dbg & 1) continue;
                    tc3::mma_tf32_ts(d, a0 + k8 * 8, dhi, idesc, (c | k8) ? 1u : 0u);
                    tc3::mma_tf32_ts(d, a0 + k8 * 8, dlo, idesc, 1u);
                }
                mma_commit(smem_u32(&empty_bar[s]));
            }
            if (alive) mma_commit(smem_u32(&acc_full[buf]));
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 12) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}
}  // namespace c1


// ---------------------------------------------------------------------------------------------
// Weight gradient of the same layer:  dW[oc][k] = (1/255) sum_m dY[m][oc] x[m][k],  m = (b, oh, ow),
// k = (c, kh, kw).  Computed transposed, D[k][oc] = sum_m A(k, m) B(m, oc), so that the 64C patch
// positions fill the 128 MMA rows (two M-tiles for C = 4) and the contraction runs over pixels:
//
//   * A(k, m) is the u8 pixel X[b][c][4 oh + kh][4 ow + kw]: exact in TF32, no lo part.  Thread =
//     patch position k = TMEM lane; one stage = two output rows (oh, oh + 1) of one sample = 40
//     pixels = 5 MMA k-steps.  The thread reads its two input rows (20 words each, the 8 kw lanes
//     of a row share addresses), picks byte kw of every word (PRMT + FADD) and stores 40 columns
//     to tensor memory.
//   * B(m, oc) = dY rows, split hi / lo by warps 8-11 into K-major SWIZZLE_128B tiles (transposing
//     stores, conflict-free with lanes along m).  D += A B_hi + A B_lo: two passes.
//   * CTAs are persistent: each takes a contiguous range of stages (split-K over the batch),
//     accumulates in TMEM for its whole lifetime and writes one partial dW^T; a deterministic
//     reduce (nn.cu splitk_reduce8_kernel) adds the partials.
//
//   warps 0-3 / 4-7  A producers of M-tile 0 / 1 (lane quarter w % 4), epilogue at the end
//   warps 8-11       B producers
//   warp 12          TMEM owner + MMA issuer
// TMEM: [0,32) [32,64) accumulators of the two M-tiles, then S stages x 96 columns (40 + pad per tile).
namespace c1w {
using namespace tc;
constexpr int NTHREADS = 13 * 32, S = 4, OC = 32, TMEM_COLS = 512, A_COL0 = 64, A_STAGE = 96, A_TILE1 = 48;
constexpr uint32_t B_STAGE = 2 * 2 * 4096;   // hi (2 slices) + lo (2 slices)

struct Args {
    const uint8_t* X;   // [B][C][H][W] u8
    const float* dY;    // [B*OH*OW][32]
    float* part;        // [ctas][32][K] partial dW (already scaled by 1/255)
    int n_stages;       // B * OH / 2
    const unsigned long long* ix;   // != null: image b is row ix[b] of X (see c1::Args)
    int C, HW, W, OHW /* OH*OW */, OW, OH2 /* OH/2 */;
    int dbg;
};

__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3])
                 : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1) conv1_wgrad_kernel(Args g) {
    extern __shared__ uint8_t smem_dyn[];
    __shared__ uint64_t full_bar[S], empty_bar[S], done_bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tiles = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const int n_mt = g.C / 2;                      // M-tiles of 128 patch positions
    // this CTA's stage range
    const int s_begin = (int)((long)g.n_stages * blockIdx.x / gridDim.x);
    const int s_end = (int)((long)g.n_stages * (blockIdx.x + 1) / gridDim.x);
    const int ns = s_end - s_begin;
    pdl_sync();

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(smem_u32(&full_bar[s]), (uint32_t)(4 * n_mt + 4)); mbar_init(smem_u32(&empty_bar[s]), 1); }
        mbar_init(smem_u32(&done_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12) {
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
        // ================================================================ A producers (+ epilogue)
        const int mt = warp >> 2, q = warp & 3;
        const int r = warp * 32 + lane;            // patch position k = (c, kh, kw)
        const bool active = mt < n_mt;
        const int c = r >> 6, kh = (r >> 3) & 7, kw = r & 7;
        const uint32_t sel = 0x7540u | (uint32_t)(kw & 3);
        const size_t plane = (size_t)c * g.HW + (size_t)kh * g.W + (size_t)(kw >> 2) * 4;
        bool alive = true;
        auto load = [&](int st, uint32_t* w) {
            const int b = st / g.OH2, oh0 = (st % g.OH2) * 2;
            const uint8_t* p0 = g.X + (size_t)(g.ix ? __ldg(g.ix + b) : (unsigned long long)b) * g.C * g.HW + plane + (size_t)(oh0 * 4) * g.W;
            const uint32_t* r0 = reinterpret_cast<const uint32_t*>(p0);
            const uint32_t* r1 = reinterpret_cast<const uint32_t*>(p0 + 4 * g.W);
#pragma unroll
            for (int j = 0; j < 20; ++j) { w[j] = __ldg(r0 + j); w[20 + j] = __ldg(r1 + j); }
        };
        if (active) {
            uint32_t w[40], wn[40];
            if (ns > 0 && !(g.dbg & 2)) load(s_begin, w);
            for (int i = 0; i < ns; ++i) {
                if (i + 1 < ns && !(g.dbg & 2)) load(s_begin + i + 1, wn);
                const uint32_t s = (uint32_t)i % S, ph = ((uint32_t)i / S) & 1u;
                if (alive && !c1::wait_bar(smem_u32(&empty_bar[s]), ph ^ 1u)) alive = false;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + A_COL0 + s * A_STAGE + (uint32_t)mt * A_TILE1;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t f[20];
#pragma unroll
                    for (int j = 0; j < 20; ++j) {
                        uint32_t v = __byte_perm(w[h * 20 + j], 0x4B000000u, sel);
                        f[j] = __float_as_uint(__uint_as_float(v) - 8388608.0f);
                    }
                    if (!(g.dbg & 4)) {
                        tc3::tmem_st16(ta + h * 20, f);
                        tmem_st4(ta + h * 20 + 16, f + 16);
                    }
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&full_bar[s]));
#pragma unroll
                for (int j = 0; j < 40; ++j) w[j] = wn[j];
            }
            // ------------------------------------------------------------ epilogue: D[k][oc] -> part[oc][k] / 255
            if (ns > 0 && alive) alive = c1::wait_bar(smem_u32(&done_bar), 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t d[32];
            if (ns > 0 && alive) {
                c1::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)mt * 32u, d);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) d[j] = 0u;
            }
            const int K = g.C * 64;
            float* out = g.part + (size_t)blockIdx.x * OC * K + r;
            const float sc = 1.0f / 255.0f;
#pragma unroll
            for (int oc = 0; oc < OC; ++oc) out[(size_t)oc * K] = __uint_as_float(d[oc]) * sc;
        }
    } else if (warp < 12) {
        // ================================================================ B producers: dY rows -> hi / lo, transposed
        const int t = tid - 256;
        bool alive = true;
        for (int i = 0; i < ns; ++i) {
            const int st = s_begin + i;
            const int b = st / g.OH2, oh0 = (st % g.OH2) * 2;
            const float* src = g.dY + ((size_t)b * g.OHW + (size_t)oh0 * g.OW) * OC;
            float4 v[3];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int idx = t + 128 * u;       // mm = idx % 40 (pixel), oc4 = idx / 40
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx < 320) v[u] = __ldg(reinterpret_cast<const float4*>(src + (size_t)(idx % 40) * OC + (idx / 40) * 4));
            }
            const uint32_t s = (uint32_t)i % S, ph = ((uint32_t)i / S) & 1u;
            if (alive && !c1::wait_bar(smem_u32(&empty_bar[s]), ph ^ 1u)) alive = false;
            const uint32_t bhi = tiles + s * B_STAGE, blo = bhi + 2 * 4096;
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int idx = t + 128 * u;
                if (idx < 320) {
                    const int mm = idx % 40, oc4 = idx / 40;
                    const float vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t oc = (uint32_t)(oc4 * 4 + j), kk = (uint32_t)(mm & 31);
                        const uint32_t off = (uint32_t)(mm >> 5) * 4096u + sw128(oc, kk >> 2) + (kk & 3u) * 4u;
                        split_store1(bhi + off, blo + off, vv[j]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&full_bar[s]));
        }
    } else if (lane == 0) {
        // ================================================================ MMA issuer
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(OC >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        bool alive = true;
        for (int i = 0; i < ns && alive; ++i) {
            const uint32_t s = (uint32_t)i % S, ph = ((uint32_t)i / S) & 1u;
            if (!c1::wait_bar(smem_u32(&full_bar[s]), ph)) { alive = false; break; }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t bhi = tiles + s * B_STAGE, blo = bhi + 2 * 4096;
            for (int mt = 0; mt < n_mt; ++mt) {
                const uint32_t d = tmem_base + (uint32_t)mt * 32u;
                const uint32_t a0 = tmem_base + A_COL0 + s * A_STAGE + (uint32_t)mt * A_TILE1;
#pragma unroll
                for (int k8 = 0; k8 < 5; ++k8) {
                    if (g.dbg & 1) continue;
                    const uint32_t sl = (uint32_t)(k8 >> 2) * 4096u;
                    const uint64_t adv = (uint64_t)((k8 & 3) * 2);
                    tc3::mma_tf32_ts(d, a0 + k8 * 8, make_desc(bhi + sl) + adv, idesc, (i | k8) ? 1u : 0u);
                    tc3::mma_tf32_ts(d, a0 + k8 * 8, make_desc(blo + sl) + adv, idesc, 1u);
                }
            }
            mma_commit(smem_u32(&empty_bar[s]));
        }
        if (ns > 0 && alive) mma_commit(smem_u32(&done_bar));
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 12) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}
}  // namespace c1w

__global__ void splitk_reduce8_kernel(const float* __restrict__ ws, float* __restrict__ C, int M, int N, int ldc,
                                      int splits, const float* __restrict__ bias, int relu, const float* __restrict__ mask, long c_plane);

static void publish_error_flag_c1() {
    static std::atomic<uint32_t> done{0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (done.load() & (1u << (dev & 31))) return;
    int* f = device_error_flag();
    BB_CUDA(cudaMemcpyToSymbol(g_err_flag_c1, &f, sizeof(f)));
    done.fetch_or(1u << (dev & 31));
}

// Can BOTH dedicated kernels take this first layer?  Then the replay batch need not be materialised: they read the ring rows
// through the sampled index list (ConvGeom::in_ix).
bool conv1_direct_ok(const ConvGeom& g) {
    const bool on = (getenv("BB_CONV1_TC") ? atoi(getenv("BB_CONV1_TC")) : 1) && (getenv("BB_TC") ? atoi(getenv("BB_TC")) : 1);
    return on && g.u8_chw && g.KH == 8 && g.KW == 8 && g.S == 4 && g.OC == 32 && (g.C == 2 || g.C == 4) && (g.W & 3) == 0 &&
           (g.OH & 1) == 0 && g.OW == 20 && g.M() >= 1024;
}

// dW[32][64C] of the AtariCnn first layer; false => geometry not handled (generic path).
bool conv1_wgrad_tc(const Ctx& c, const ConvGeom& g, const float* dY, const void* X, float* dW) {
    static const int on = getenv("BB_CONV1_TC") ? atoi(getenv("BB_CONV1_TC")) : 1;
    if (!on || !g.u8_chw || g.KH != 8 || g.KW != 8 || g.S != 4 || g.OC != 32 || (g.C != 2 && g.C != 4) || (g.W & 3) || (g.OH & 1) ||
        g.OW != 20 || g.M() < 1024)
        return false;
    c1w::Args a;
    a.X = (const uint8_t*)X; a.dY = dY; a.part = c.ws; a.n_stages = g.B * g.OH / 2; a.ix = g.in_ix; a.C = g.C; a.HW = g.H * g.W; a.W = g.W;
    a.OHW = g.OH * g.OW; a.OW = g.OW; a.OH2 = g.OH / 2;
    static const int dbg = getenv("BB_CONV1_DEBUG") ? atoi(getenv("BB_CONV1_DEBUG")) : 0;
    a.dbg = dbg;
    const int K = g.K();
    const int ctas = std::min(a.n_stages, c.sms);
    if ((size_t)ctas * 32 * K > c.ws_floats - 1024) return false;
    const size_t smem = (size_t)c1w::S * c1w::B_STAGE + 1024;
    BB_ENSURE_SMEM(c1w::conv1_wgrad_kernel, smem);
    publish_error_flag_c1();
    launch_pdl(c1w::conv1_wgrad_kernel, dim3(ctas), dim3(c1w::NTHREADS), smem, c.stream, a);
    BB_LAUNCHED();
    c.mark("tc_conv1_wgrad");
    const size_t total = (size_t)32 * K;
    const int blocks = (int)std::min<size_t>((total * 8 + 255) / 256, (size_t)c.sms * 8);
    launch_pdl(splitk_reduce8_kernel, dim3(blocks), dim3(256), 0, c.stream, c.ws, dW, 32, K, K, ctas, nullptr, 0, nullptr, 0L);
    BB_LAUNCHED();
    c.mark("splitk_reduce");
    return true;
}

// Returns false when the geometry is not the AtariCnn first layer (the caller falls back to the
// generic implicit GEMM).
bool conv1_fwd_tc(const Ctx& c, const ConvGeom& g, const void* X, const float* W, const float* b, float* Y, bool relu) {
    static const int on = getenv("BB_CONV1_TC") ? atoi(getenv("BB_CONV1_TC")) : 1;
    if (!on || !g.u8_chw || g.KH != 8 || g.KW != 8 || g.S != 4 || g.OC != 32 || g.C < 1 || g.C > 8 || (g.W & 3) || g.M() < 1024)
        return false;
    c1::Args a;
    a.X = (const uint8_t*)X; a.Wt = W; a.bias = b; a.Y = Y; a.y_plane = g.y_plane; a.rowbase = g.rowbase; a.ix = g.in_ix; a.OHW = g.OH * g.OW; a.M = g.M(); a.C = g.C;
    a.HW = g.H * g.W; a.W = g.W; a.relu = relu ? 1 : 0; a.n_tiles = (a.M + 127) / 128;
    static const int dbg = getenv("BB_CONV1_DEBUG") ? atoi(getenv("BB_CONV1_DEBUG")) : 0;
    a.dbg = dbg;
    const size_t smem = (size_t)g.C * 2 * 2 * 4096 + 1024;   // 2 slices per channel, hi + lo
    BB_ENSURE_SMEM(c1::conv1_fwd_kernel, 8 * 2 * 2 * 4096 + 1024);
    publish_error_flag_c1();
    const int ctas = std::min(a.n_tiles, 2 * c.sms);
    launch_pdl(c1::conv1_fwd_kernel, dim3(ctas), dim3(c1::NTHREADS), smem, c.stream, a);
    BB_LAUNCHED();
    c.mark("tc_conv1_fwd");
    return true;
}

int conv1_error_flag() {
    int e = 0;
    cudaMemcpyFromSymbol(&e, g_tc_error_c1, sizeof(int));
    return e;
}

}  // namespace bb

// Timing hook: `iters` back-to-back launches of the AtariCnn first-layer forward on a synthetic
// u8 batch [B][C][84][84]; mean milliseconds per launch (CUDA events on the launching stream).
extern "C" int32_t bb_bench_conv1(int32_t device, int32_t B, int32_t C, int32_t iters, float* ms_out) {
    BB_API_BEGIN
    using namespace bb;
    DeviceGuard dg(device);
    Ctx c;
    c.device = device; c.sms = num_sms(device); c.stream = device_stream(device);
    ConvGeom g{};
    g.B = B; g.C = C; g.H = 84; g.W = 84; g.OC = 32; g.KH = 8; g.KW = 8; g.S = 4; g.OH = 20; g.OW = 20; g.u8_chw = true;
    const size_t M = (size_t)g.M();
    std::vector<int> rb(M);
    for (int b = 0; b < B; ++b)
        for (int oh = 0; oh < 20; ++oh)
            for (int ow = 0; ow < 20; ++ow) rb[((size_t)b * 20 + oh) * 20 + ow] = b * C * 84 * 84 + oh * 4 * 84 + ow * 4;
    int* d_rb = dev_alloc<int>(M);
    BB_CUDA(cudaMemcpy(d_rb, rb.data(), M * sizeof(int), cudaMemcpyHostToDevice));
    g.rowbase = d_rb;
    uint8_t* X = dev_alloc<uint8_t>((size_t)B * C * 84 * 84);
    BB_CUDA(cudaMemsetAsync(X, 37, (size_t)B * C * 84 * 84, c.stream));
    float* W = dev_alloc<float>((size_t)32 * C * 64);
    float* bias = dev_alloc_zero<float>(32, c.stream);
    float* Y = dev_alloc<float>(M * 32);
    fill_uniform(c, W, (size_t)32 * C * 64, 0.1f, 3);
    cudaEvent_t e0, e1;
    BB_CUDA(cudaEventCreate(&e0));
    BB_CUDA(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) BB_CHECK(conv1_fwd_tc(c, g, X, W, bias, Y, true), "conv1_fwd_tc declined the geometry");
    BB_CUDA(cudaEventRecord(e0, c.stream));
    for (int i = 0; i < iters; ++i) conv1_fwd_tc(c, g, X, W, bias, Y, true);
    BB_CUDA(cudaEventRecord(e1, c.stream));
    BB_CUDA(cudaStreamSynchronize(c.stream));
    float ms = 0.f;
    BB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *ms_out = ms / iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_rb); cudaFree(X); cudaFree(W); cudaFree(bias); cudaFree(Y);
    BB_CHECK(conv1_error_flag() == 0, "conv1 tcgen05 pipeline timed out");
    BB_API_END
}
