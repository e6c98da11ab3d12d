// tma_probe.cu -- scratch hardware probe (not part of the library): what do 4-D tiled / im2col TMA boxes put in shared
// memory, and which UMMA descriptor conventions does tcgen05.mma.kind::tf32 honour for MN-major operands?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I border_b200/csrc tools/tma_probe.cu -o build/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include "tma_gemm.cuh"

using namespace bb::tg;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static void* drv(const char* n) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint(n, &f, cudaEnableDefault, &q);
    return f;
}

// ---- probe 1/2: one TMA load of `bytes` bytes into shared memory, dumped raw
__global__ void load_dump_kernel(const __grid_constant__ CUtensorMap tm, int kind, int c0, int c1, int c2, int c3, int ow, int oh,
                                 uint32_t bytes, float* out, int* err) {
    extern __shared__ uint8_t smem_dyn[];
    __shared__ uint64_t bar;
    const uint32_t tiles = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    float* t = reinterpret_cast<float*>(smem_dyn + (tiles - smem_u32(smem_dyn)));
    for (uint32_t i = threadIdx.x; i < bytes / 4; i += blockDim.x) t[i] = -7777.f;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(smem_u32(&bar), bytes);
        if (kind == 0) tma_load_3d(tiles, &tm, smem_u32(&bar), c0, c1, c2);
        else if (kind == 1) tma_load_4d(tiles, &tm, smem_u32(&bar), c0, c1, c2, c3);
        else tma_load_im2col(tiles, &tm, smem_u32(&bar), c0, c1, c2, c3, (uint16_t)ow, (uint16_t)oh);
        mbar_wait(smem_u32(&bar), 0, err);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = t[i];
}

// de-swizzle: element e (0..31) of 128-byte row r of a SWIZZLE_128B tile
static float at(const std::vector<float>& d, int r, int e) {
    int chunk = e / 4, w = e % 4;
    return d[(size_t)r * 32 + ((chunk ^ (r & 7)) * 4) + w];
}
// the same for the 128B-span / 32B-atom swizzle (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, UMMA layout type 1 = Swizzle<2,5,2>)
static float at32(const std::vector<float>& d, int r, int e) {
    int chunk = e / 8, w = e % 8;
    return d[(size_t)r * 32 + ((chunk ^ (r & 3)) * 8) + w];
}

// ---- probe 3: one 128 x 64 x 32 tile product from shared-memory images prepared by the host, any major-ness / strides
struct MmaProbe {
    int a_mn, b_mn;
    int a_lt, b_lt;                        // descriptor layout type (2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B)
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;   // bytes
    uint32_t a_kadv, b_kadv;               // bytes per k-step of 8
};
__global__ void __launch_bounds__(128) mma_probe_kernel(const float* a_img, const float* b_img, MmaProbe p, float* out, int* err) {
    extern __shared__ uint8_t smem_dyn[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t tiles = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    float* sa = reinterpret_cast<float*>(smem_dyn + (tiles - smem_u32(smem_dyn)));
    float* sb = sa + 4096;  // A image 16 KB, B image 8 KB
    for (int i = threadIdx.x; i < 4096; i += 128) sa[i] = a_img[i];
    for (int i = threadIdx.x; i < 2048; i += 128) sb[i] = b_img[i];
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) {
        auto mk = [](uint32_t addr, uint32_t lbo, uint32_t sbo, int lt) {
            return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) |
                   ((uint64_t)lt << 61);
        };
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) | ((64u >> 3) << 17) |
                               ((128u >> 4) << 24);
        for (int k4 = 0; k4 < 4; ++k4) {
            const uint64_t da = mk(tiles + k4 * p.a_kadv, p.a_lbo, p.a_sbo, p.a_lt);
            const uint64_t db = mk(tiles + 16384 + k4 * p.b_kadv, p.b_lbo, p.b_sbo, p.b_lt);
            mma_tf32(tmem, da, db, idesc, k4 ? 1u : 0u);
        }
        mma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0, err);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + c0, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) out[(q * 32 + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

// ---- probe 4: tcgen05.mma issue/execute throughput: R back-to-back MMAs on fixed shared-memory tiles, one CTA
__global__ void __launch_bounds__(128) mma_rate_kernel(int N, int kind_bf16, int a_tmem, int reps, long long* cycles, int* err) {
    extern __shared__ uint8_t smem_dyn[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t tiles = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    float* sa = reinterpret_cast<float*>(smem_dyn + (tiles - smem_u32(smem_dyn)));
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) sa[i] = 0.f;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) {
        const uint32_t fmt = kind_bf16 ? 1u : 2u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t da = desc_k(tiles), db = desc_k(tiles + 16384);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint64_t adv = (uint64_t)((r & 3) * 2);
            if (a_tmem) {
                if (kind_bf16)
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem), "r"(tmem + 128u + (r & 3) * 8u), "l"(db + adv), "r"(idesc), "r"(1u) : "memory");
                else
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem), "r"(tmem + 128u + (r & 3) * 8u), "l"(db + adv), "r"(idesc), "r"(1u) : "memory");
            } else {
                if (kind_bf16)
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(da + adv), "l"(db + adv), "r"(idesc), "r"(1u) : "memory");
                else
                    mma_tf32(tmem, da + adv, db + adv, idesc, 1u);
            }
        }
        long long t1 = clock64();
        mma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0, err);
        long long t2 = clock64();
        cycles[0] = t1 - t0;
        cycles[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

// ---- probe 5: TWO warps of one CTA issuing tcgen05.mma concurrently into disjoint accumulator columns
__global__ void __launch_bounds__(128) mma_rate2_kernel(int N, int issuers, int reps, long long* cycles, int* err) {
    extern __shared__ uint8_t smem_dyn[];
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tmem_base_s;
    const uint32_t tiles = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    float* sa = reinterpret_cast<float*>(smem_dyn + (tiles - smem_u32(smem_dyn)));
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) sa[i] = 0.f;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar[0]), 1);
        mbar_init(smem_u32(&bar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && w < issuers) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t da = desc_k(tiles), db = desc_k(tiles + 16384);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint64_t adv = (uint64_t)((r & 3) * 2);
            mma_tf32(tmem + (uint32_t)w * 128u, da + adv, db + adv, idesc, 1u);
        }
        mma_commit(smem_u32(&bar[w]));
        mbar_wait(smem_u32(&bar[w]), 0, err);
        long long t2 = clock64();
        cycles[w] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

int main() {
    EncodeTiledFn enc_t = (EncodeTiledFn)drv("cuTensorMapEncodeTiled");
    EncodeIm2colFn enc_i = (EncodeIm2colFn)drv("cuTensorMapEncodeIm2col");
    if (!enc_t || !enc_i) { printf("no driver entry points\n"); return 1; }
    int* err;
    CK(cudaMalloc(&err, 4));
    CK(cudaMemset(err, 0, 4));
    CK(cudaFuncSetAttribute(load_dump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    float* d_out;
    CK(cudaMalloc(&d_out, 1 << 20));

    // ---------------- probe 1: 4-D "MN" view of a row-major [K][N] matrix (+ second plane)
    {
        const int K = 64, N = 128;
        std::vector<float> X(2 * K * N);
        for (int pl = 0; pl < 2; ++pl)
            for (int k = 0; k < K; ++k)
                for (int n = 0; n < N; ++n) X[(size_t)pl * K * N + k * N + n] = pl * 100000.f + k * 1000.f + n;
        float* dX;
        CK(cudaMalloc(&dX, X.size() * 4));
        CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
        cuuint64_t gd[4] = {32, (cuuint64_t)K, (cuuint64_t)N / 32, 2};
        cuuint64_t gs[3] = {(cuuint64_t)N * 4, 128, (cuuint64_t)K * N * 4};
        cuuint32_t bx[4] = {32, 32, 2, 2}, es[4] = {1, 1, 1, 1};
        for (int sw32 = 0; sw32 < 2; ++sw32) {
        CUtensorMap tm;
        CUresult r = enc_t(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dX, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           sw32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("probe1 encode 4d-mn (atom32 = %d): %d\n", sw32, (int)r);
        const uint32_t bytes = 32 * 32 * 2 * 2 * 4;
        load_dump_kernel<<<1, 128, 48 * 1024>>>(tm, 1, 0, 32, 1, 0, 0, 0, bytes, d_out, err);
        CK(cudaDeviceSynchronize());
        std::vector<float> d(bytes / 4);
        CK(cudaMemcpy(d.data(), d_out, bytes, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int pl = 0; pl < 2; ++pl)
            for (int blk = 0; blk < 2; ++blk)
                for (int k = 0; k < 32; ++k)
                    for (int n = 0; n < 32; ++n) {
                        float want = pl * 100000.f + (32 + k) * 1000.f + 32 * (1 + blk) + n;
                        float got = sw32 ? at32(d, (pl * 2 + blk) * 32 + k, n) : at(d, (pl * 2 + blk) * 32 + k, n);
                        if (got != want && bad++ < 6) printf("  probe1 mismatch pl %d blk %d k %d n %d: got %.0f want %.0f\n", pl, blk, k, n, got, want);
                    }
        printf("probe1 (4-D tiled MN box, atom32 = %d): %d mismatches; row 1: %.0f %.0f ... %.0f %.0f ... %.0f\n", sw32, bad, d[32], d[33], d[40], d[41], d[48]);
        }
        cudaFree(dX);
    }
    // ---------------- probe 2: im2col box
    for (int variant = 0; variant < 3; ++variant) {
        const int Nn = 3, H = 6, W = 6, Cc = 32, KH = 2, KW = 2, S = 2;
        std::vector<float> X((size_t)Nn * H * W * Cc);
        for (int n = 0; n < Nn; ++n)
            for (int h = 0; h < H; ++h)
                for (int w = 0; w < W; ++w)
                    for (int c = 0; c < Cc; ++c) X[(((size_t)n * H + h) * W + w) * Cc + c] = n * 10000.f + h * 1000.f + w * 100.f + c;
        float* dX;
        CK(cudaMalloc(&dX, X.size() * 4));
        CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
        cuuint64_t gd[4] = {(cuuint64_t)Cc, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Nn};
        cuuint64_t gs[3] = {(cuuint64_t)Cc * 4, (cuuint64_t)W * Cc * 4, (cuuint64_t)H * W * Cc * 4};
        int lower[2] = {0, 0}, upper[2] = {-(KW - 1), -(KH - 1)};
        cuuint32_t es[4] = {1, (cuuint32_t)S, (cuuint32_t)S, 1};
        const int pixels = 16;
        CUtensorMap tm;
        CUresult r = enc_i(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dX, gd, gs, lower, upper, 32, pixels, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (variant == 1) reinterpret_cast<uint64_t*>(&tm)[1] &= ~(1ull << 21);   // CUTLASS's small-tensor correction
        printf("probe2 variant %d encode im2col: %d   (word1 bit21 = %d)\n", variant, (int)r, (int)((reinterpret_cast<uint64_t*>(&tm)[1] >> 21) & 1));
        const uint32_t bytes = 32 * pixels * 4;
        // start at image 0, output position (oh 1, ow 2) -> base (h 2, w 4); tap offset (kh 1, kw 1)
        int cw = 4, ch = 2, offw = 1, offh = 1;
        if (variant == 2) { cw = 0; ch = 0; offw = 0; offh = 0; }
        load_dump_kernel<<<1, 128, 48 * 1024>>>(tm, 2, 0, cw, ch, 0, offw, offh, bytes, d_out, err);
        CK(cudaDeviceSynchronize());
        std::vector<float> d(bytes / 4);
        CK(cudaMemcpy(d.data(), d_out, bytes, cudaMemcpyDeviceToHost));
        printf("  rows (n,h,w of channel 0): ");
        for (int p = 0; p < pixels; ++p) printf("%.0f ", at(d, p, 0));
        printf("\n  channel 5 of row 0: %.0f\n", at(d, 0, 5));
        cudaFree(dX);
    }
    // ---------------- probe 3: MMA layouts.  A[128][32], B[64][32] (values small integers: exact in tf32)
    {
        std::vector<float> A(128 * 32), B(64 * 32), ref(128 * 64);
        for (int m = 0; m < 128; ++m)
            for (int k = 0; k < 32; ++k) A[m * 32 + k] = (float)((m * 7 + k * 3) % 11 - 5);
        for (int n = 0; n < 64; ++n)
            for (int k = 0; k < 32; ++k) B[n * 32 + k] = (float)((n * 5 + k * 2) % 13 - 6);
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n) {
                float s = 0;
                for (int k = 0; k < 32; ++k) s += A[m * 32 + k] * B[n * 32 + k];
                ref[m * 64 + n] = s;
            }
        auto sw = [](int row, int e) { return row * 32 + (((e / 4) ^ (row & 7)) * 4) + e % 4; };
        // K-major images: row = m (or n), 32 k per row
        std::vector<float> ak(4096), bk(2048);
        for (int m = 0; m < 128; ++m)
            for (int k = 0; k < 32; ++k) ak[sw(m, k)] = A[m * 32 + k];
        for (int n = 0; n < 64; ++n)
            for (int k = 0; k < 32; ++k) bk[sw(n, k)] = B[n * 32 + k];
        // MN-major images, layout "P" (what a [blk][k][128 B] TMA box gives): row = blk*32 + k, 32 m per row
        std::vector<float> amP(4096), bmP(2048), amQ(4096), bmQ(2048);
        for (int m = 0; m < 128; ++m)
            for (int k = 0; k < 32; ++k) amP[sw((m / 32) * 32 + k, m % 32)] = A[m * 32 + k];
        for (int n = 0; n < 64; ++n)
            for (int k = 0; k < 32; ++k) bmP[sw((n / 32) * 32 + k, n % 32)] = B[n * 32 + k];
        // layout "Q" (CUTLASS's default tiling of the MN_SW128 atom): atom = 8 k-rows x 128 B; atoms along MN first (1024 B apart),
        // then along k: row = (k/8) * (nblk*8) + blk*8 + k%8
        for (int m = 0; m < 128; ++m)
            for (int k = 0; k < 32; ++k) amQ[sw((k / 8) * 32 + (m / 32) * 8 + k % 8, m % 32)] = A[m * 32 + k];
        for (int n = 0; n < 64; ++n)
            for (int k = 0; k < 32; ++k) bmQ[sw((k / 8) * 16 + (n / 32) * 8 + k % 8, n % 32)] = B[n * 32 + k];
        // 32-byte-atom swizzle images (UMMA layout type 1): rows of 128 B, k-rows 128 B apart, 32 k-rows per MN block
        auto sw32 = [](int row, int e) { return row * 32 + (((e / 8) ^ (row & 3)) * 8) + e % 8; };
        std::vector<float> amR(4096), bmR(2048);
        for (int m = 0; m < 128; ++m)
            for (int k = 0; k < 32; ++k) amR[sw32((m / 32) * 32 + k, m % 32)] = A[m * 32 + k];
        for (int n = 0; n < 64; ++n)
            for (int k = 0; k < 32; ++k) bmR[sw32((n / 32) * 32 + k, n % 32)] = B[n * 32 + k];
        float *da, *db;
        CK(cudaMalloc(&da, 16384));
        CK(cudaMalloc(&db, 8192));
        struct Case { const char* name; const std::vector<float>*a, *b; MmaProbe p; };
        std::vector<Case> cases = {
            {"K x K", &ak, &bk, {0, 0, 2, 2, 16, 1024, 16, 1024, 32, 32}},
            {"K x MN(type1: lbo 4096 sbo 512 kadv 1024)", &ak, &bmR, {0, 1, 2, 1, 16, 1024, 4096, 512, 32, 1024}},
            {"K x MN(type1, lbo/sbo swapped)", &ak, &bmR, {0, 1, 2, 1, 16, 1024, 512, 4096, 32, 1024}},
            {"MN(type1) x K", &amR, &bk, {1, 0, 1, 2, 4096, 512, 16, 1024, 1024, 32}},
            {"MN(type1, swapped) x K", &amR, &bk, {1, 0, 1, 2, 512, 4096, 16, 1024, 1024, 32}},
            {"MN(type1) x MN(type1)", &amR, &bmR, {1, 1, 1, 1, 4096, 512, 4096, 512, 1024, 1024}},
            {"K x MN(type2 image, type2 desc)", &ak, &bmP, {0, 1, 2, 2, 16, 1024, 4096, 1024, 32, 1024}},
        };
        for (auto& c : cases) {
            CK(cudaMemcpy(da, c.a->data(), 16384, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(db, c.b->data(), 8192, cudaMemcpyHostToDevice));
            CK(cudaMemset(d_out, 0, 128 * 64 * 4));
            mma_probe_kernel<<<1, 128, 40 * 1024>>>(da, db, c.p, d_out, err);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("probe3 %-45s CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
            std::vector<float> o(128 * 64);
            CK(cudaMemcpy(o.data(), d_out, o.size() * 4, cudaMemcpyDeviceToHost));
            int bad = 0, zeros = 0;
            for (int i = 0; i < 128 * 64; ++i) { if (o[i] != ref[i]) ++bad; if (o[i] == 0.f) ++zeros; }
            printf("probe3 %-45s mismatches %5d / 8192  (zeros %d)  o[0..3] = %.0f %.0f %.0f %.0f  ref = %.0f %.0f %.0f %.0f\n", c.name, bad, zeros, o[0],
                   o[1], o[2], o[3], ref[0], ref[1], ref[2], ref[3]);
        }
    }
    // ---------------- probe 4: MMA rate
    {
        CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        long long* dc;
        CK(cudaMalloc(&dc, 16));
        const int reps = 2048;
        for (int bf = 0; bf < 2; ++bf)
            for (int at = 0; at < 2; ++at)
                for (int N : {64, 128, 256}) {
                    for (int warm = 0; warm < 2; ++warm) {
                        mma_rate_kernel<<<1, 128, 56 * 1024>>>(N, bf, at, reps, dc, err);
                        CK(cudaDeviceSynchronize());
                    }
                    long long hc[2];
                    CK(cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost));
                    const double macs = 128.0 * N * (bf ? 16 : 8);
                    printf("probe4 %s A-%s N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA -> %.0f MAC/cycle\n", bf ? "bf16" : "tf32", at ? "tmem" : "smem", N,
                           (double)hc[0] / reps, (double)hc[1] / reps, macs / ((double)hc[1] / reps));
                }
        // one vs two CTAs per SM (wall time by events): is the ~108-cycle floor per issuing thread or per SM?
        {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int N : {64, 128}) {   // (N <= 128: accumulator + A columns fit 256 tensor-memory columns, two CTAs per SM)
                for (int ctas : {148, 296, 444}) {
                    mma_rate_kernel<<<ctas, 128, 56 * 1024>>>(N, 0, 0, reps, dc, err);
                    CK(cudaDeviceSynchronize());
                    cudaEventRecord(e0);
                    mma_rate_kernel<<<ctas, 128, 56 * 1024>>>(N, 0, 0, reps, dc, err);
                    cudaEventRecord(e1);
                    CK(cudaDeviceSynchronize());
                    float ms = 0;
                    cudaEventElapsedTime(&ms, e0, e1);
                    long long hc[2];
                    CK(cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost));
                    printf("probe4 tf32 N=%3d, %d CTAs: wall %.1f us, CTA 0: %.1f cyc/MMA\n", N, ctas, ms * 1e3, (double)hc[1] / reps);
                }
            }
        }
        CK(cudaFuncSetAttribute(mma_rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        for (int N : {64, 128})
            for (int iss : {1, 2}) {
                for (int warm = 0; warm < 2; ++warm) {
                    mma_rate2_kernel<<<1, 128, 56 * 1024>>>(N, iss, reps, dc, err);
                    CK(cudaDeviceSynchronize());
                }
                long long hc[2];
                CK(cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost));
                printf("probe5 tf32 N=%3d, %d issuing warps in ONE CTA: %.1f / %.1f cyc per MMA per issuer\n", N, iss, (double)hc[0] / reps, (double)hc[iss - 1] / reps);
            }
        // many CTAs at once (one per SM): does the rate hold chip-wide (power / clocks)?
        for (int N : {128, 256}) {
            mma_rate_kernel<<<148, 128, 56 * 1024>>>(N, 0, 0, reps, dc, err);
            CK(cudaDeviceSynchronize());
            long long hc[2];
            CK(cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost));
            printf("probe4 tf32 A-smem N=%3d x148 CTAs: complete %.1f cyc/MMA\n", N, (double)hc[1] / reps);
        }
    }
    int herr = 0;
    CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
    printf("error flag: %d\n", herr);
    return 0;
}
