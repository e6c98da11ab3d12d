// gemm.cuh -- fp32 CUDA-core tile GEMM with gather ("implicit im2col") operands.
//
// Every dense contraction of the agents (border-tch-agent/src/cnn/base.rs:23-36 conv2d/linear,
// mlp/base.rs:13-41 linear, and their backward passes that libtorch autograd derives for
// opt.rs:74-83 backward_step) is one of three operand patterns of  C[M,N] = sum_k A(m,k) B(k,n):
//
//   forward  : A k-contiguous (dense rows or im2col gather), B = W[N][K] k-contiguous
//   dgrad    : A = dY[M][N'] k-contiguous,                   B = W[N'][K']  n-contiguous
//   wgrad    : A = dY[m][n] read as A(n, m) m-contiguous,    B n-contiguous (dense rows or gather)
//
// A gather operand is separable: element (r, c) lives at  base[rowbase[r] + coloff[c]]  -- the
// im2col matrix of a convolution is exactly that, so no im2col buffer is ever materialised in the
// forward / wgrad passes.  u8 gather operands (Atari frames straight from the replay batch) are
// widened to float and scaled by 1/255 in the loader (cnn/base.rs:26).
//
// Tiles: BM x BN x 16, 256 threads, 4x4 (or 8x4 ...) register tile per thread, global->register
// prefetch of the next k-tile while the current one is multiplied out of shared memory.
// Split-K (blockIdx.z) writes partial tiles to a workspace; splitk_reduce_kernel finishes with the
// epilogue.  All accumulation is fp32 FFMA: this is the parity path (1e-4 rel on losses).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"

namespace bb {

// An NHWC float tensor walked as the implicit-GEMM operand of a convolution by TMA im2col loads (tma_gemm.cu):
// rows = filter positions (n, oh, ow), columns = (kh, kw, c).
struct TmaConv {
    int N, H, W, C;   // tensor dims
    int KH, KW, S;    // filter taps and traversal stride
    int flip;         // 1: tap offsets run backwards (transposed convolution over a zero-padded gradient)
    int pad;          // zero padding on every side, supplied by the TMA unit's out-of-bounds fill (the tensor itself is
                      // unpadded): filter bases start at -pad
};

struct GemmArgs {
    const void* A;       // float (or u8 when a_u8)
    const void* B;       // float (or u8 when b_u8)
    float* C;
    int M, N, K;
    long lda, ldb;       // dense leading dimensions (elements)
    int ldc;
    // gather tables (element offsets); null => dense
    const int* a_rowbase;  // [M]  (A k-contiguous)
    const int* a_koff;     // [K]
    const int* b_rowbase;  // [K]  (B n-contiguous)
    const int* b_noff;     // [N]
    // epilogue
    const float* bias;     // [N] or null
    const float* mask;     // [M][ldc] or null: C *= (mask > 0)
    int relu;
    // split-K
    int split_k;           // >= 1
    int k_per_split;       // multiple of BK
    float* workspace;      // [split_k][M][N] when split_k > 1
    // tcgen05 path: where the generic->async proxy fence runs (0 = in every producer thread after
    // its st.shared, 1 = in the MMA-issuing thread after it has acquired the full barrier)
    int fence_mode;
    // tcgen05 path only: store C transposed (element (m, n) at C[n*ldc + m]); used by the conv weight
    // gradient, which is computed as dW^T = im2col^T dY so that the long K*K*C axis fills the 128 MMA rows
    int trans_out;
    // tcgen05 path only, direct (unsplit) stores: separable output map, element (m, n) at
    // C[c_rowoff[m] + c_coloff[n]] (the mask is read through the same map); null => m*ldc / n.  Groups of
    // 4 consecutive n (n % 4 == 0) must stay contiguous.  Used by the gather-form conv data gradient,
    // whose GEMM rows / columns are (image, h/S, w/S) / (h%S, w%S, channel).
    const int* c_rowoff;
    const int* c_coloff;
    // tcgen05 path only: every gather-table entry in use is a multiple of 4 elements (conv tables with C % 4 == 0), so
    // whole-tile problems may take the branch-free producer loads (tc_gemm_kernel<..., FAST>)
    int tables_vec4;
    // TMA path (tma_gemm.cu): element offset of the operand's "lo" plane (x - tf32_trunc(x)) from its fp32 plane; 0 = the
    // tensor has none (the SIMT-producer kernels are used).  c_plane != 0: the epilogue also stores lo(C) at C + c_plane.
    long a_plane, b_plane, c_plane;
    const TmaConv* a_conv;  // A's gather tables describe this convolution (im2col tensor maps instead of the tables)
};

constexpr int kBK = 16;

template <int BM, int BN, int TM, int TN, bool A_KMAJOR, bool B_KMAJOR, bool A_U8, bool B_U8>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) gemm_kernel(GemmArgs g) {
    constexpr int NT = (BM / TM) * (BN / TN);  // threads per CTA
    static_assert(NT % 32 == 0 && NT <= 1024, "whole warps");
    pdl_sync();
    static_assert(TM % 4 == 0 && TN % 4 == 0, "float4 register tiles");
    constexpr int BK = kBK;
    constexpr int PAD = 4;
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int k_begin = blockIdx.z * g.k_per_split;
    const int k_end = min(g.K, k_begin + g.k_per_split);

    constexpr int A_LD = (BM * BK / 4 + NT - 1) / NT;  // float4 groups per thread
    constexpr int B_LD = (BN * BK / 4 + NT - 1) / NT;
    float4 a_reg[A_LD], b_reg[B_LD];

    const float* Af = reinterpret_cast<const float*>(g.A);
    const uint8_t* Au = reinterpret_cast<const uint8_t*>(g.A);
    const float* Bf = reinterpret_cast<const float*>(g.B);
    const uint8_t* Bu = reinterpret_cast<const uint8_t*>(g.B);
    const bool a_vec = A_U8 || (((g.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0));
    const bool b_vec = B_U8 || (((g.ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0));

    // per-thread row bases of the k-contiguous A operand are fixed for the whole k loop
    long a_base[A_LD];
    if (A_KMAJOR) {
#pragma unroll
        for (int i = 0; i < A_LD; ++i) {
            int l = tid + i * NT;
            int m = m0 + l / (BK / 4);
            a_base[i] = -1;
            if (l < BM * BK / 4 && m < g.M) a_base[i] = g.a_rowbase ? (long)g.a_rowbase[m] : (long)m * g.lda;
        }
    }
    long b_base[B_LD];
    if (B_KMAJOR) {
#pragma unroll
        for (int i = 0; i < B_LD; ++i) {
            int l = tid + i * NT;
            int n = n0 + l / (BK / 4);
            b_base[i] = -1;
            if (l < BN * BK / 4 && n < g.N) b_base[i] = (long)n * g.ldb;
        }
    }

    auto load_tiles = [&](int k0) {
        // ---- A
#pragma unroll
        for (int i = 0; i < A_LD; ++i) {
            int l = tid + i * NT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (l < BM * BK / 4) {
                if (A_KMAJOR) {
                    int k = k0 + (l % (BK / 4)) * 4;
                    if (a_base[i] >= 0 && k < k_end) {
                        long off = a_base[i] + (g.a_koff ? (long)g.a_koff[k] : (long)k);
                        if (A_U8) {
                            uint32_t w = *reinterpret_cast<const uint32_t*>(Au + off);
                            const float s = 1.0f / 255.0f;
                            v.x = (float)(w & 0xff) * s; v.y = (float)((w >> 8) & 0xff) * s;
                            v.z = (float)((w >> 16) & 0xff) * s; v.w = (float)(w >> 24) * s;
                            if (k + 3 >= k_end) {
                                if (k + 1 >= k_end) v.y = 0.f;
                                if (k + 2 >= k_end) v.z = 0.f;
                                v.w = 0.f;
                            }
                        } else if (a_vec && k + 3 < k_end && ((off & 3) == 0)) {
                            v = __ldg(reinterpret_cast<const float4*>(Af + off));
                        } else {
                            v.x = __ldg(Af + off);
                            if (k + 1 < k_end) v.y = __ldg(Af + off + (g.a_koff ? g.a_koff[k + 1] - g.a_koff[k] : 1));
                            if (k + 2 < k_end) v.z = __ldg(Af + off + (g.a_koff ? g.a_koff[k + 2] - g.a_koff[k] : 2));
                            if (k + 3 < k_end) v.w = __ldg(Af + off + (g.a_koff ? g.a_koff[k + 3] - g.a_koff[k] : 3));
                        }
                    }
                } else {  // A(m,k) = A[k*lda + m], m contiguous
                    int k = k0 + l / (BM / 4);
                    int m = m0 + (l % (BM / 4)) * 4;
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
            }
            a_reg[i] = v;
        }
        // ---- B
#pragma unroll
        for (int i = 0; i < B_LD; ++i) {
            int l = tid + i * NT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (l < BN * BK / 4) {
                if (B_KMAJOR) {  // B(k,n) = B[n*ldb + k]
                    int k = k0 + (l % (BK / 4)) * 4;
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
                } else {  // n contiguous: dense B[k*ldb + n] or gather rowbase[k] + noff[n]
                    int k = k0 + l / (BN / 4);
                    int n = n0 + (l % (BN / 4)) * 4;
                    if (k < k_end && n < g.N) {
                        long off = (g.b_rowbase ? (long)g.b_rowbase[k] : (long)k * g.ldb) +
                                   (g.b_noff ? (long)g.b_noff[n] : (long)n);
                        if (B_U8) {
                            uint32_t w = *reinterpret_cast<const uint32_t*>(Bu + off);
                            const float s = 1.0f / 255.0f;
                            v.x = (float)(w & 0xff) * s; v.y = (float)((w >> 8) & 0xff) * s;
                            v.z = (float)((w >> 16) & 0xff) * s; v.w = (float)(w >> 24) * s;
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
            }
            b_reg[i] = v;
        }
    };

    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_LD; ++i) {
            int l = tid + i * NT;
            if (l < BM * BK / 4) {
                if (A_KMAJOR) {
                    int k = (l % (BK / 4)) * 4, m = l / (BK / 4);
                    As[buf][k + 0][m] = a_reg[i].x; As[buf][k + 1][m] = a_reg[i].y;
                    As[buf][k + 2][m] = a_reg[i].z; As[buf][k + 3][m] = a_reg[i].w;
                } else {
                    int k = l / (BM / 4), m = (l % (BM / 4)) * 4;
                    *reinterpret_cast<float4*>(&As[buf][k][m]) = a_reg[i];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < B_LD; ++i) {
            int l = tid + i * NT;
            if (l < BN * BK / 4) {
                if (B_KMAJOR) {
                    int k = (l % (BK / 4)) * 4, n = l / (BK / 4);
                    Bs[buf][k + 0][n] = b_reg[i].x; Bs[buf][k + 1][n] = b_reg[i].y;
                    Bs[buf][k + 2][n] = b_reg[i].z; Bs[buf][k + 3][n] = b_reg[i].w;
                } else {
                    int k = l / (BN / 4), n = (l % (BN / 4)) * 4;
                    *reinterpret_cast<float4*>(&Bs[buf][k][n]) = b_reg[i];
                }
            }
        }
    };

    constexpr int RM = TM / 4, RN = TN / 4;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    int buf = 0;
    if (k_begin < k_end) {
        load_tiles(k_begin);
        store_tiles(0);
    }
    __syncthreads();
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        const bool has_next = k0 + BK < k_end;
        if (has_next) load_tiles(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int r = 0; r < RM; ++r) {
                float4 v = *reinterpret_cast<const float4*>(&As[buf][k][r * (BM / RM) + ty * 4]);
                a[r * 4 + 0] = v.x; a[r * 4 + 1] = v.y; a[r * 4 + 2] = v.z; a[r * 4 + 3] = v.w;
            }
#pragma unroll
            for (int r = 0; r < RN; ++r) {
                float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][r * (BN / RN) + tx * 4]);
                b[r * 4 + 0] = v.x; b[r * 4 + 1] = v.y; b[r * 4 + 2] = v.z; b[r * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (has_next) {
            store_tiles(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }

    // ---- epilogue
    const bool direct = g.split_k <= 1;
    float* out = direct ? g.C : g.workspace + (size_t)blockIdx.z * g.M * g.N;
    const int ldo = direct ? g.ldc : g.N;
#pragma unroll
    for (int rm = 0; rm < RM; ++rm)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int m = m0 + rm * (BM / RM) + ty * 4 + i;
            if (m >= g.M) continue;
#pragma unroll
            for (int rn = 0; rn < RN; ++rn) {
                int n = n0 + rn * (BN / RN) + tx * 4;
                if (n >= g.N) continue;
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = acc[rm * 4 + i][rn * 4 + j];
                if (direct) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (n + j < g.N) {
                            if (g.bias) v[j] += g.bias[n + j];
                            if (g.relu) v[j] = fmaxf(v[j], 0.f);
                            if (g.mask) v[j] = g.mask[(size_t)m * g.ldc + n + j] > 0.f ? v[j] : 0.f;
                        }
                    }
                }
                float* dst = out + (size_t)m * ldo + n;
                if (direct && g.c_plane)
                    for (int j = 0; j < 4; ++j)
                        if (n + j < g.N) dst[g.c_plane + j] = v[j] - __uint_as_float(__float_as_uint(v[j]) & 0xffffe000u);
                if (n + 3 < g.N && ((ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0))
                    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                else
                    for (int j = 0; j < 4; ++j)
                        if (n + j < g.N) dst[j] = v[j];
            }
        }
}

// Finishes a split-K GEMM: sums the partial tiles in a fixed order (deterministic) and applies
// the bias / ReLU / ReLU-mask epilogue.
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, float* __restrict__ C, int M, int N, int ldc,
                                     int splits, const float* __restrict__ bias, int relu,
                                     const float* __restrict__ mask, long c_plane);

}  // namespace bb
