// tma_gemm.cuh -- TMA-fed tcgen05 tile GEMM, 3xTF32, warp-specialised (the round-2 replacement of tc_gemm.cuh's
// SIMT producers).
//
// Same contraction as gemm.cuh / tc_gemm.cuh,  C[M,N] = sum_k A(m,k) B(k,n),  for the layers of border-tch-agent's
// networks (cnn/base.rs:23-36 conv2d/linear, mlp/base.rs:13-41 linear) and the backward passes libtorch autograd
// derives for opt.rs:74-83.  What changed is WHO moves the operands:
//
//   * ONE thread issues cp.async.bulk.tensor (TMA) loads that land the fp32 operand tiles in shared memory in the
//     canonical swizzled UMMA layouts -- tiled boxes for dense operands, im2col boxes (cuTensorMapEncodeIm2col) for the
//     implicit-GEMM rows of a convolution -- and arms an mbarrier with the byte count.  No SIMT global loads, no
//     address tables.
//   * 3xTF32 from ONE copy of the data: the tensor core ignores the low 13 mantissa bits of a TF32 operand (measured:
//     bit-identical to an explicit truncation), so the fp32 tile itself IS the "hi" operand.  While the ring runs, the
//     four epilogue warps -- idle until the accumulator is complete -- derive the "lo" tiles  lo = x - tf32_trunc(x)
//     from the landed tile with a purely elementwise shared->shared pass (same swizzled position, so no layout
//     knowledge), fence to the async proxy and hand the stage to the MMA thread.  x*y ~= x.y + lo(x).y + x.lo(y) with three
//     tcgen05.mma.kind::tf32 products accumulated in fp32 (error ~2^-21, what keeps the 1e-4 loss parity with the
//     libtorch-CPU oracle).  (Round-2 measurement: shipping a second lo plane through L2 instead made every layer
//     L2-bandwidth bound -- 48 KB per 128x64x32 slice and CTA, ~1.3 GB per DQN step.)
//   * ONE thread issues the MMAs (M = 128, K = 8 per instruction).  K-major operands (k contiguous in memory) and
//     MN-major operands (m or n contiguous: weight gradients, the linear data gradient) both feed the tensor core
//     directly: the instruction descriptor's major bits select the layout, nothing is transposed.
//     "Stacked" 3xTF32: the hi and lo tiles of B are adjacent, so one descriptor spans [B_hi; B_lo] as 2*BN columns:
//     A x [B_hi; B_lo] gives x.y in TMEM columns [0,BN) and x.lo(y) in [BN,2BN); lo(A) x B_hi accumulates into
//     [0,BN).  tcgen05.commit hands the stage back to the TMA thread.
//   * the eight split warps then read the accumulator (tcgen05.ld 32x32b), add the halves, apply bias / ReLU /
//     ReLU-mask, and store C (row-major, transposed, or through a separable output map).
//   * split-K without a second launch: every split stores its partial tile, the LAST CTA of a tile (atomic ticket)
//     adds the partials in split order -- deterministic -- and runs the epilogue.
//
// Tile 128 x BN x 32, STAGES-deep TMA ring, 192 threads.  Every mbarrier wait is bounded: a mis-programmed
// pipeline raises the error flag (checked by the host after every update) instead of hanging the GPU.
#pragma once
#include <type_traits>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"

namespace bb {
namespace tg {

constexpr int BM = 128, BK = 32, NTHREADS = 320;   // TMA warp, MMA warp, 2 x 4 split / epilogue warps

enum OpKind {
    OP_K_TILED = 0,    // k contiguous, dense rows:            map (k, row, plane), box (32, rows, planes)
    OP_K_IM2COL = 1,   // k contiguous, conv rows (A only):     im2col map (c, w, h, n), box 32 channels x 128 pixels
    OP_MN_TILED = 2,   // m / n contiguous, dense:              map (32, k, mn/32, plane), box (32, 32, blocks, planes)
    OP_MN_IM2COL = 3,  // m contiguous conv rows (A only):      im2col map, box 32 channels x 32 pixels per 32-row block
};

// Pixel traversal of an NHWC tensor the way an im2col tensor map walks it.
struct Im2col {
    int ow, ohw;         // filter positions per image row / per image
    int stride;          // traversal stride (the convolution's stride)
    int kw;              // filter taps per filter row (tap = kh * kw + kw)
    int cblocks;         // 32-channel blocks per tap
    int flip_w, flip_h;  // < 0: offset = tap coordinate; >= 0: offset = flip - tap coordinate (transposed convolution)
    int nblocks;         // OP_MN_IM2COL: number of 32-row blocks (taps * cblocks)
    int pad;             // filter bases start at -pad (padding = the tensor map's out-of-bounds zero fill)
};

struct Args {
    int M, N, K;
    int slices_per_split;  // k-slices of 32 per blockIdx.z
    int split_k;
    float* C;
    long c_plane;          // != 0: also store lo(C) at C + c_plane
    int ldc;
    const float* bias;     // [N] or null
    const float* mask;     // C *= (mask > 0), read through the same addressing as C; or null
    int relu;
    int trans_out;         // element (m, n) at C[n*ldc + m]
    const int* c_rowoff;   // separable output map: element (m, n) at C[c_rowoff[m] + c_coloff[n]]
    const int* c_coloff;
    float* workspace;      // [split_k][M][N] partial tiles
    unsigned int* counters;  // [tiles] tickets, zero between launches
    int* error;            // device flag: a bounded wait timed out
    long long* trace;      // debug (BB_TMA_TRACE=1): clock64 stamps of CTA (0,0,0): [role 0..2][slice < 64][4], then per CTA
                           // (first 1024) %globaltimer at entry / after the dependency wait / at exit and its SM id
    Im2col ga;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of the (converged) warp: the single-thread instructions (TMA, tcgen05.mma, tcgen05.commit) go under it
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.b32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* err, int code = 1) {
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 24); ++it)
        if (mbar_try_wait(bar, parity)) return true;
    *reinterpret_cast<volatile int*>(err) = code;  // pinned, mapped host memory (common.cuh: device_error_flag)
    __threadfence_system();
    return false;
}

// For the warp-uniform loops of the TMA / MMA warps ALL lanes poll.  (Letting one lane poll and broadcasting the result
// makes the loop body divergent for ptxas again: the TMA / MMA operands fall back to vector registers and every UTCHMMA is
// wrapped in an R2UR + ELECT loop -- measured 1030 instead of 480 cycles for the 8 MMAs of a k-slice.)
__device__ __forceinline__ bool mbar_wait_warp(uint32_t bar, uint32_t parity, int* err, int code) {
    return mbar_wait(bar, parity, err, code);
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_im2col(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h, int n,
                                                uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], "
        "{%7, %8};" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
        : "memory");
}

// UMMA shared-memory descriptors, descriptor version 1.
//   K-major : SWIZZLE_128B (layout type 2; TMA CU_TENSOR_MAP_SWIZZLE_128B): rows of 128 B (32 tf32 along k), 8-row groups
//             1024 B apart (SBO); LBO unused.  k-step (8 tf32) = +32 B.
//   MN-major: 32-bit operands have exactly one legal MN-major layout, SWIZZLE_128B_BASE32B (layout type 1; TMA
//             CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B: 32-byte chunks permuted inside each 128-byte row by (row & 3); measured on
//             B200: with the 16-byte-atom swizzle the MMA returns zeros).  Atoms of 4 k-rows x 128 B (32 tf32 along m/n); next
//             32 m/n = +LBO (4096 B: a [32 k][128 B] block), next 4 k = +SBO (512 B).  k-step (8 k-rows) = +1024 B.
__device__ __forceinline__ uint64_t desc_k(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)1 << 61);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* f) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::
            "r"(taddr),
        "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]), "f"(f[4]), "f"(f[5]), "f"(f[6]), "f"(f[7]), "f"(f[8]), "f"(f[9]), "f"(f[10]),
        "f"(f[11]), "f"(f[12]), "f"(f[13]), "f"(f[14]), "f"(f[15])
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
        "%14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// (n, h, w) of the filter base of linear position p
__device__ __forceinline__ void pixel_of(const Im2col& g, int p, int& n, int& h, int& w) {
    n = p / g.ohw;
    const int r = p - n * g.ohw;
    const int oh = r / g.ow;
    h = oh * g.stride - g.pad;
    w = (r - oh * g.ow) * g.stride - g.pad;
}

}  // namespace tg

// PASSES = 3: 3xTF32 (fp32 parity); PASSES = 1: one TF32 product per fp32 product (the `fast` precision mode; only the
// hi planes are loaded).
template <int BN, int STAGES, int AK, int BKIND, int PASSES, int MINB>
__global__ void __launch_bounds__(tg::NTHREADS, MINB)
tma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, tg::Args g) {
    using namespace tg;
    static_assert(BN == 32 || BN == 64 || BN == 128, "BN");
    constexpr bool A_MN = AK == OP_MN_TILED || AK == OP_MN_IM2COL;
    constexpr bool B_MN = BKIND == OP_MN_TILED;
    constexpr uint32_t A_TILE = BM * 128, B_TILE = BN * 128;        // bytes of one fp32 tile
    // shared-memory stage: [A][B][lo(B)] -- TMA fills A and B, the split warps derive lo(B) next to B (one descriptor
    // then spans [B; lo(B)]) and put lo(A) into TENSOR memory (32 columns per stage: the MMA takes it from there)
    constexpr uint32_t STAGE_BYTES = A_TILE + (PASSES == 3 ? 2 : 1) * B_TILE;
    constexpr uint32_t TX_BYTES = A_TILE + B_TILE;
    // accumulator columns: [0,BN) x.y + lo(x).y, [BN,2BN) x.lo(y); the epilogue adds the halves
    constexpr uint32_t ACC_COLS = PASSES == 3 ? 2 * BN : BN;
    constexpr uint32_t ALO_COL0 = ACC_COLS;                          // first tensor-memory column of the lo(A) stages
    constexpr uint32_t USED_COLS = ACC_COLS + (PASSES == 3 ? STAGES * 32 : 0);
    constexpr uint32_t TMEM_COLS = USED_COLS <= 32 ? 32 : USED_COLS <= 64 ? 64 : USED_COLS <= 128 ? 128 : USED_COLS <= 256 ? 256 : 512;
    extern __shared__ uint8_t smem_dyn[];
    __shared__ uint64_t full_bar[STAGES], ready_bar[STAGES], empty_bar[STAGES], accum_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ int s_last;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int nks_total = (g.K + BK - 1) / BK;
    const int ks0 = blockIdx.z * g.slices_per_split;
    const int nks = max(0, min(nks_total - ks0, g.slices_per_split));
    const uint32_t tiles = (smem_u32(smem_dyn) + 1023u) & ~1023u;   // SWIZZLE_128B atoms are 1024 B aligned

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&ready_bar[s]), 4);   // one arrive per split warp
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        mbar_init(smem_u32(&accum_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const unsigned cta_lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    if (g.trace && tid == 0 && cta_lin < 1024) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); g.trace[768 + cta_lin * 4 + 0] = (long long)t; }
    pdl_sync();  // barriers / tensor memory are set up while the previous kernel of the stream drains
    if (g.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0) g.trace[(1 * 64 + 63) * 4 + 0] = clock64();
    if (g.trace && tid == 0 && cta_lin < 1024) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); g.trace[768 + cta_lin * 4 + 1] = (long long)t; }

    if (warp == 0) {
        // ================================================================ TMA producer
        // The WHOLE warp walks the loop so that every coordinate / address is warp-uniform (uniform registers feed UTMALDG
        // directly); only the instructions themselves sit under elect.sync.  Issued from inside an `if (lane == 0)` branch the
        // operands live in vector registers and ptxas wraps every UTMALDG / UTCHMMA in an R2UR + ELECT "waterfall" loop:
        // measured ~108 cycles per tcgen05.mma instead of ~55-64 (tools/tma_probe.cu, probe 4 vs probe 5).
        int pn = 0, ph = 0, pw = 0;                 // OP_K_IM2COL: filter base of the tile's first row
        if (AK == OP_K_IM2COL) pixel_of(g.ga, m0, pn, ph, pw);
        int bc[4] = {0, 0, 0, 0};                    // OP_MN_IM2COL: channel / tap offsets of the tile's four 32-row blocks
        int bw[4] = {0, 0, 0, 0}, bh[4] = {0, 0, 0, 0};
        if (AK == OP_MN_IM2COL) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int q = min(m0 / 32 + j, g.ga.nblocks - 1);   // rows past M repeat the last block (never stored)
                const int tap = q / g.ga.cblocks;
                bc[j] = (q - tap * g.ga.cblocks) * 32;
                const int th = tap / g.ga.kw, tw = tap - th * g.ga.kw;
                bw[j] = g.ga.flip_w >= 0 ? g.ga.flip_w - tw : tw;
                bh[j] = g.ga.flip_h >= 0 ? g.ga.flip_h - th : th;
            }
        }
        bool alive = true;
        for (int i = 0; i < nks && alive; ++i) {
            const int s = i % STAGES;
            const uint32_t phs = (uint32_t)(i / STAGES) & 1u;
            const bool tr0 = g.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && i < 64 && lane == 0;
            if (tr0) g.trace[(0 * 64 + i) * 4 + 0] = clock64();
            if (!mbar_wait_warp(smem_u32(&empty_bar[s]), phs ^ 1u, g.error, 11)) { alive = false; break; }
            if (tr0) g.trace[(0 * 64 + i) * 4 + 1] = clock64();
            const uint32_t bar = smem_u32(&full_bar[s]);
            const int ks = ks0 + i;
            const uint32_t a_hi = tiles + s * STAGE_BYTES, b_hi = a_hi + A_TILE;
            int c = 0, ow_ = 0, oh_ = 0, n_ = 0, h_ = 0, w_ = 0;
            if (AK == OP_K_IM2COL) {
                const int tap = ks / g.ga.cblocks;
                c = (ks - tap * g.ga.cblocks) * 32;
                const int th = tap / g.ga.kw, tw = tap - th * g.ga.kw;
                ow_ = g.ga.flip_w >= 0 ? g.ga.flip_w - tw : tw;
                oh_ = g.ga.flip_h >= 0 ? g.ga.flip_h - th : th;
            } else if (AK == OP_MN_IM2COL) {   // contraction over pixels, 32 per k-slice
                pixel_of(g.ga, ks * BK, n_, h_, w_);
            }
            if (elect_one()) {
                mbar_expect_tx(bar, TX_BYTES);
                if (AK == OP_K_TILED) {
                    tma_load_2d(a_hi, &tmA, bar, ks * BK, m0);
                } else if (AK == OP_MN_TILED) {
                    tma_load_3d(a_hi, &tmA, bar, 0, ks * BK, m0 / 32);
                } else if (AK == OP_K_IM2COL) {
                    tma_load_im2col(a_hi, &tmA, bar, c, pw, ph, pn, (uint16_t)ow_, (uint16_t)oh_);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) tma_load_im2col(a_hi + j * 4096, &tmA, bar, bc[j], w_, h_, n_, (uint16_t)bw[j], (uint16_t)bh[j]);
                }
                if (BKIND == OP_K_TILED) tma_load_2d(b_hi, &tmB, bar, ks * BK, n0);
                else tma_load_3d(b_hi, &tmB, bar, 0, ks * BK, n0 / 32);
            }
            __syncwarp();
            if (tr0) g.trace[(0 * 64 + i) * 4 + 2] = clock64();
        }
        // Producer tail: every tcgen05.commit of the ring must have ARRIVED before this CTA may exit.  The accumulator
        // barrier the epilogue waits on is a different barrier: its arrival can be observed while the last stages' "empty"
        // arrivals are still in flight, and they would then land in the shared memory of the NEXT CTA scheduled on this SM
        // (same barrier addresses), flipping its phases -- measured as sporadic timeouts / wrong results whenever a grid had
        // more CTAs than fit at once.
        for (int i = max(0, nks - STAGES); i < nks && alive; ++i)
            alive = mbar_wait_warp(smem_u32(&empty_bar[i % STAGES]), (uint32_t)(i / STAGES) & 1u, g.error, 15);
    } else if (warp == 1) {
        // ================================================================ MMA issuer (warp-uniform loop, one elected lane issues)
        // instruction descriptor: D = F32 (bit 4), A = B = TF32 (2 << 7, 2 << 10), major bits 15 / 16 (1 = MN-major),
        // N >> 3 at bit 17, M >> 4 at bit 24
        constexpr uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                        ((uint32_t)(BM >> 4) << 24);
        constexpr uint32_t idesc1 = idesc_base | ((uint32_t)(BN >> 3) << 17);
        constexpr uint32_t idesc2 = idesc_base | ((uint32_t)((2 * BN) >> 3) << 17);
        // lo(A) comes from tensor memory (lane = row, column = k): no A-major bit
        constexpr uint32_t idesc_lo = (1u << 4) | (2u << 7) | (2u << 10) | ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BM >> 4) << 24) |
                                      ((uint32_t)(BN >> 3) << 17);
        constexpr uint64_t a_adv = A_MN ? (1024 >> 4) : 2, b_adv = B_MN ? (1024 >> 4) : 2;
        bool alive = true;
        for (int i = 0; i < nks && alive; ++i) {
            const int s = i % STAGES;
            const uint32_t phs = (uint32_t)(i / STAGES) & 1u;
            const bool tr1 = g.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && i < 64 && lane == 0;
            if (tr1) g.trace[(1 * 64 + i) * 4 + 0] = clock64();
            // 3 passes: the stage is ready once the split warps have derived its lo operands; 1 pass: as soon as the TMA data landed
            if (!mbar_wait_warp(smem_u32(PASSES == 3 ? &ready_bar[s] : &full_bar[s]), phs, g.error, 12)) { alive = false; break; }
            if (tr1) g.trace[(1 * 64 + i) * 4 + 1] = clock64();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = tiles + s * STAGE_BYTES, b_hi = a_hi + A_TILE;
            const uint64_t da_hi = A_MN ? desc_mn(a_hi) : desc_k(a_hi);
            const uint64_t db = B_MN ? desc_mn(b_hi) : desc_k(b_hi);
            const uint32_t a_lo = tmem_base + ALO_COL0 + (uint32_t)s * 32u;   // lo(A): 128 lanes x 32 columns of this stage
            const uint32_t ebar = smem_u32(&empty_bar[s]);
            if (elect_one()) {
#pragma unroll
                for (int k4 = 0; k4 < BK / 8; ++k4) {
                    const uint32_t acc = (i | k4) ? 1u : 0u;
                    if (PASSES == 3) {
                        mma_tf32(tmem_base, da_hi + a_adv * k4, db + b_adv * k4, idesc2, acc);            // x . [y; lo(y)]
                        mma_tf32_ts(tmem_base, a_lo + (uint32_t)k4 * 8u, db + b_adv * k4, idesc_lo, 1u);   // lo(x) . y
                    } else {
                        mma_tf32(tmem_base, da_hi + a_adv * k4, db + b_adv * k4, idesc1, acc);
                    }
                }
                mma_commit(ebar);  // frees the stage once the MMAs above have read it
            }
            __syncwarp();
            if (tr1) g.trace[(1 * 64 + i) * 4 + 2] = clock64();
        }
        if (nks > 0 && elect_one()) mma_commit(smem_u32(&accum_bar));
        __syncwarp();
    } else {
        // ================================================================ split warps, then epilogue (2 groups of 4 warps)
        // Measured (clock64 trace, tools/tma_trace.py): one group of four warps needs ~850 cycles per k-slice for the lo pass
        // (loads, 48 subtractions, stores, tcgen05.wait::st + proxy fence + arrive) while the 8 MMAs of a slice issue in
        // ~440: two groups take alternate slices.
        const int q = warp & 3;                      // the TMEM lane quarter this warp may touch
        const int grp = (warp - 2) >> 2;             // split group 0 / 1; in the epilogue: which half of the tile's columns
        if (PASSES == 3) {
            // lo operands of every landed stage.  lo(B): an elementwise pass over the B tile (16 bytes per thread and trip,
            // consecutive lanes on consecutive addresses: conflict-free), same offset in the tile after it.  lo(A): thread =
            // tile row = tensor-memory lane; it reads its 32 k-values of the swizzled tile and stores their lo parts into the
            // stage's 32 tensor-memory columns.
            const int te = (tid - 64) & 127;
            const uint32_t row = (uint32_t)(q * 32 + lane);
            bool alive_s = true;
            for (int i = 0; i < nks && alive_s; ++i) {
                const int s = i % STAGES;
                const uint32_t phs = (uint32_t)(i / STAGES) & 1u;
                const bool tr2 = g.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && i < 64 && tid == 64;
                if (tr2) g.trace[(2 * 64 + i) * 4 + 0] = clock64();
                // BOTH groups observe every phase of every "full" barrier, although a group only works on alternate slices: a
                // parity wait can tell a phase from its neighbours only, and with an odd ring depth a group that skipped the
                // other group's phase of a stage would take the completion of slice i - 2*STAGES for that of slice i (measured:
                // sporadic stale tiles / timeouts as soon as CTAs queued behind each other)
                if (!mbar_wait_warp(smem_u32(&full_bar[s]), phs, g.error, 13)) { alive_s = false; break; }
                if ((i & 1) != grp) continue;
                if (tr2) g.trace[(2 * 64 + i) * 4 + 1] = clock64();
                const uint32_t a_hi = tiles + s * STAGE_BYTES, b_hi = a_hi + A_TILE;
                constexpr int B4 = B_TILE / 16 / 128;   // 16-byte groups of B per thread
                float4 vb[B4];
                float fa[32];
#pragma unroll
                for (int j = 0; j < B4; ++j)
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(vb[j].x), "=f"(vb[j].y), "=f"(vb[j].z), "=f"(vb[j].w) : "r"(b_hi + (uint32_t)(te + 128 * j) * 16u));
                if (!A_MN) {
                    // K-major SWIZZLE_128B: 16-byte chunk c of row r sits at r*128 + ((c ^ (r & 7)) << 4)
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(fa[4 * c]), "=f"(fa[4 * c + 1]), "=f"(fa[4 * c + 2]), "=f"(fa[4 * c + 3])
                                     : "r"(a_hi + row * 128u + ((((uint32_t)c) ^ (row & 7u)) << 4)));
                } else {
                    // MN-major, 32-byte-atom swizzle: element (m, k) of 32-row block b = m / 32 sits at
                    // b*4096 + k*128 + ((((m % 32) / 8) ^ (k & 3)) * 32) + (m % 8) * 4
                    const uint32_t base = a_hi + (row >> 5) * 4096u + (row & 7u) * 4u, ch = (row & 31u) >> 3;
#pragma unroll
                    for (int k = 0; k < 32; ++k)
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(fa[k]) : "r"(base + (uint32_t)k * 128u + ((ch ^ ((uint32_t)k & 3u)) << 5)));
                }
#pragma unroll
                for (int j = 0; j < B4; ++j)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(b_hi + B_TILE + (uint32_t)(te + 128 * j) * 16u), "f"(tf32_lo(vb[j].x)),
                                 "f"(tf32_lo(vb[j].y)), "f"(tf32_lo(vb[j].z)), "f"(tf32_lo(vb[j].w))
                                 : "memory");
#pragma unroll
                for (int k = 0; k < 32; ++k) fa[k] = tf32_lo(fa[k]);
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + ALO_COL0 + (uint32_t)s * 32u;
                tmem_st16(ta, fa);
                tmem_st16(ta + 16, fa + 16);
                if (tr2) g.trace[(2 * 64 + i) * 4 + 2] = clock64();
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores of lo(B) -> the tensor core's async proxy
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&ready_bar[s])) : "memory");
                if (tr2) g.trace[(2 * 64 + i) * 4 + 3] = clock64();
            }
        }
        // ================================================================ epilogue (8 warps)
        // Phase A: accumulator (tcgen05.ld, halves added) -> a staging tile in the now idle ring memory; thread = row, group =
        // column half.  Phase B: a coalesced pass over the staged tile (16 bytes per thread and trip, consecutive lanes on
        // consecutive addresses of C) applies bias / ReLU / ReLU-mask and stores C -- or the split-K partial.  (Storing straight
        // from the row-per-thread registers cost 8000 cycles per 128x64 tile: every lane wrote its own 256-byte-strided row.)
        // The storage space is (rr, cc) with cc contiguous: (m, n), or (n, m) for a transposed output.
        const bool split = g.split_k > 1;
        const bool tr = g.trans_out != 0;
        const bool mapped = g.c_rowoff != nullptr;
        const int te = tid - 64;                                 // 0..255
        constexpr int LDN = BN + 4, LDT = BM + 4;                // padded row lengths of the staging tile (floats)
        const uint32_t stg = tiles;
        const int RT = tr ? BN : BM, CT = tr ? BM : BN;          // tile extent in storage space
        const int R0 = tr ? n0 : m0, C0 = tr ? m0 : n0;
        const int Rmax = tr ? g.N : g.M, Cmax = tr ? g.M : g.N;
        const int G = CT / 4;                                    // 16-byte groups per staged row
        const int ldp = tr ? g.M : g.N;                          // leading dimension of a split-K partial
        const bool ov4 = (mapped || (g.ldc & 3) == 0) && (g.c_plane & 3) == 0 &&
                         ((reinterpret_cast<uintptr_t>(g.C) | reinterpret_cast<uintptr_t>(g.mask)) & 15) == 0;

        const bool bias_v4 = (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0;   // (cc is a multiple of 4)
        // final values of storage-space group (rr, cc .. cc+nv-1) -> C (+ lo plane)
        auto store_final = [&](int rr, int cc, float4 a, int nv) {
            float x[4] = {a.x, a.y, a.z, a.w};
            const size_t o = mapped ? (size_t)g.c_rowoff[rr] + (size_t)g.c_coloff[cc] : (size_t)rr * g.ldc + cc;
            if (g.bias) {
                if (tr) {
                    const float b = __ldg(g.bias + rr);
                    x[0] += b; x[1] += b; x[2] += b; x[3] += b;
                } else if (nv == 4 && bias_v4) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + cc));
                    x[0] += b.x; x[1] += b.y; x[2] += b.z; x[3] += b.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (j < nv) x[j] += __ldg(g.bias + cc + j);
                }
            }
            if (g.relu) { x[0] = fmaxf(x[0], 0.f); x[1] = fmaxf(x[1], 0.f); x[2] = fmaxf(x[2], 0.f); x[3] = fmaxf(x[3], 0.f); }
            if (ov4 && nv == 4) {
                if (g.mask) {
                    const float4 k = __ldg(reinterpret_cast<const float4*>(g.mask + o));
                    x[0] = k.x > 0.f ? x[0] : 0.f; x[1] = k.y > 0.f ? x[1] : 0.f; x[2] = k.z > 0.f ? x[2] : 0.f; x[3] = k.w > 0.f ? x[3] : 0.f;
                }
                *reinterpret_cast<float4*>(g.C + o) = make_float4(x[0], x[1], x[2], x[3]);
                if (g.c_plane) *reinterpret_cast<float4*>(g.C + g.c_plane + o) = make_float4(tf32_lo(x[0]), tf32_lo(x[1]), tf32_lo(x[2]), tf32_lo(x[3]));
            } else {
                for (int j = 0; j < nv; ++j) {
                    const size_t oj = mapped ? (size_t)g.c_rowoff[rr] + (size_t)g.c_coloff[cc + j] : o + j;
                    float y = x[j];
                    if (g.mask) y = g.mask[oj] > 0.f ? y : 0.f;
                    g.C[oj] = y;
                    if (g.c_plane) g.C[g.c_plane + oj] = tf32_lo(y);
                }
            }
        };

        bool alive = true;
        const bool tr3 = g.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 64;
        if (tr3) g.trace[(2 * 64 + 63) * 4 + 0] = clock64();
        if (nks > 0) alive = mbar_wait_warp(smem_u32(&accum_bar), 0, g.error, 14);
        if (tr3) g.trace[(2 * 64 + 63) * 4 + 1] = clock64();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- phase A
        const uint32_t row = (uint32_t)(q * 32 + lane);
        constexpr int HALF = BN / 2;
#pragma unroll
        for (int cc = 0; cc < HALF; cc += 16) {
            const int c0 = grp * HALF + cc;
            uint32_t r[16];
            if (nks > 0 && alive) {
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
                tmem_ld16(ta, r);
                if (PASSES == 3) {
                    uint32_t r2[16];
                    tmem_ld16(ta + BN, r2);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
                } else {
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = 0u;
            }
            if (!tr) {
#pragma unroll
                for (int j4 = 0; j4 < 16; j4 += 4)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (row * LDN + (uint32_t)(c0 + j4)) * 4u), "r"(r[j4]), "r"(r[j4 + 1]),
                                 "r"(r[j4 + 2]), "r"(r[j4 + 3])
                                 : "memory");
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(stg + ((uint32_t)(c0 + j) * LDT + row) * 4u), "r"(r[j]) : "memory");
            }
        }
        if (tr3) g.trace[(2 * 64 + 62) * 4 + 0] = clock64();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tr3) g.trace[(2 * 64 + 62) * 4 + 1] = clock64();
        // ---- phase B: compile-time tile geometry per output orientation (the index arithmetic and the loop were ~100
        // instructions per 16-byte group with runtime extents: 4400 cycles per tile), 4 groups in flight per thread
        float* part = split ? g.workspace + (size_t)blockIdx.z * g.M * g.N : nullptr;
        const bool pv4 = (ldp & 3) == 0;
        auto phase_b = [&](auto tr_tag) {
            constexpr bool kTr = decltype(tr_tag)::value;
            constexpr int kRT = kTr ? BN : BM, kG = (kTr ? BM : BN) / 4, kLd = kTr ? LDT : LDN;
            constexpr int kIters = kRT * kG / 256;
            static_assert(kRT * kG % 256 == 0, "tile groups are a multiple of the epilogue threads");
#pragma unroll 4
            for (int it = 0; it < kIters; ++it) {
                const int idx = te + it * 256;
                const int rl = idx / kG, cl = (idx % kG) * 4;
                const int rr = R0 + rl, cc = C0 + cl;
                if (rr >= Rmax || cc >= Cmax) continue;
                const int nv = min(4, Cmax - cc);
                float4 x;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(stg + ((uint32_t)rl * kLd + (uint32_t)cl) * 4u));
                if (!split) {
                    store_final(rr, cc, x, nv);
                } else {
                    float* pz = part + (size_t)rr * ldp + cc;
                    if (pv4 && nv == 4) *reinterpret_cast<float4*>(pz) = x;
                    else {
                        pz[0] = x.x;
                        if (nv > 1) pz[1] = x.y;
                        if (nv > 2) pz[2] = x.z;
                        if (nv > 3) pz[3] = x.w;
                    }
                }
            }
        };
        if (tr) phase_b(std::true_type{}); else phase_b(std::false_type{});
        if (split) {
            // The last CTA of this tile to arrive (atomic ticket) folds the partials in split order -- deterministic -- and runs
            // the final epilogue: a coalesced pass with 16 independent 16-byte loads in flight per thread (4 positions x 4 splits).
            __threadfence();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const unsigned int tile = blockIdx.y * gridDim.x + blockIdx.x;
            if (tid == 64) {
                const unsigned int t = atomicAdd(&g.counters[tile], 1u);
                s_last = (t == (unsigned int)g.split_k - 1u) ? 1 : 0;
                if (s_last) g.counters[tile] = 0u;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (s_last) {
                __threadfence();
                const size_t zs = (size_t)g.M * g.N;
                for (int base = 0; base < RT * G; base += 256 * 4) {
                    float4 acc[4];
                    const float* src[4];
                    int rr[4], cc[4], nv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int idx = base + te + 256 * i;
                        rr[i] = R0 + idx / G;
                        cc[i] = C0 + (idx % G) * 4;
                        nv[i] = (idx < RT * G && rr[i] < Rmax) ? min(4, Cmax - cc[i]) : 0;   // valid elements of the group
                        src[i] = g.workspace + (size_t)rr[i] * ldp + cc[i];
                        acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    for (int z0 = 0; z0 < g.split_k; z0 += 4) {
                        float4 t[4][4];
#pragma unroll
                        for (int dz = 0; dz < 4; ++dz)
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                t[dz][i] = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (z0 + dz < g.split_k && nv[i] > 0) {
                                    const float* pz = src[i] + (size_t)(z0 + dz) * zs;
                                    if (pv4 && nv[i] == 4) t[dz][i] = __ldcg(reinterpret_cast<const float4*>(pz));
                                    else {
                                        t[dz][i].x = __ldcg(pz);
                                        if (nv[i] > 1) t[dz][i].y = __ldcg(pz + 1);
                                        if (nv[i] > 2) t[dz][i].z = __ldcg(pz + 2);
                                        if (nv[i] > 3) t[dz][i].w = __ldcg(pz + 3);
                                    }
                                }
                            }
#pragma unroll
                        for (int dz = 0; dz < 4; ++dz)
#pragma unroll
                            for (int i = 0; i < 4; ++i) { acc[i].x += t[dz][i].x; acc[i].y += t[dz][i].y; acc[i].z += t[dz][i].z; acc[i].w += t[dz][i].w; }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (nv[i] > 0) store_final(rr[i], cc[i], acc[i], nv[i]);
                }
            }
        }
    }

    if (g.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 64) g.trace[(2 * 64 + 63) * 4 + 2] = clock64();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (g.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 64) g.trace[(2 * 64 + 63) * 4 + 3] = clock64();
    if (g.trace && tid == 0 && cta_lin < 1024) {   // whole-grid view: when did every CTA start / end, on which SM
        unsigned long long t; unsigned sm;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
        g.trace[768 + cta_lin * 4 + 2] = (long long)t; g.trace[768 + cta_lin * 4 + 3] = sm;
    }
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace bb
