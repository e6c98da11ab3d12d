// nn.cu -- see nn.cuh.  Layer primitives on top of gemm.cuh plus the elementwise kernels
// (col2im, column sums, fused Adam, Polyak update, init).
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include "gemm.cuh"
#include <cooperative_groups.h>
#include "nn.cuh"

namespace bb {

void Profiler::clear() {
    for (auto& m : marks) cudaEventDestroy(m.second);
    marks.clear();
}
void Ctx::mark(const char* kernel) const {
    if (!prof) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, stream);
    prof->marks.emplace_back(phase + ":" + layer + ":" + kernel, e);
}

void Ctx::alloc_scratch(size_t floats) {
    ws_floats = floats;
    ws = dev_alloc_zero<float>(ws_floats, stream);  // the tail holds colsum's block counters (must start at 0)
    n_tickets = 4096;
    tickets = dev_alloc_zero<unsigned int>(n_tickets, stream);
}
void Ctx::free_scratch() {
    cudaFree(ws); cudaFree(tickets);
    ws = nullptr; tickets = nullptr;
}

void Ctx::fork_to(const Ctx& s) const {
    BB_CUDA(cudaEventRecord(ev, stream));
    BB_CUDA(cudaStreamWaitEvent(s.stream, ev, 0));
}
void Ctx::join_from(const Ctx& s) const {
    BB_CUDA(cudaEventRecord(s.ev, s.stream));
    BB_CUDA(cudaStreamWaitEvent(stream, s.ev, 0));
}

// ------------------------------------------------------------------------------- split-K reduce

// Few splits of a large tile: one thread per output element, coalesced across elements.
__device__ __forceinline__ float lo_of(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

__global__ void splitk_reduce_kernel(const float* __restrict__ ws, float* __restrict__ C, int M, int N, int ldc,
                                     int splits, const float* __restrict__ bias, int relu,
                                     const float* __restrict__ mask, long c_plane) {
    pdl_sync();
    size_t total = (size_t)M * N;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int m = (int)(i / N), n = (int)(i % N);
        float s = 0.f;
        for (int z = 0; z < splits; ++z) s += ws[(size_t)z * total + i];
        if (bias) s += bias[n];
        if (relu) s = fmaxf(s, 0.f);
        if (mask) s = mask[(size_t)m * ldc + n] > 0.f ? s : 0.f;
        C[(size_t)m * ldc + n] = s;
        if (c_plane) C[c_plane + (size_t)m * ldc + n] = lo_of(s);
    }
}

// Many splits of a small tile (weight gradients): a CTA takes 32 consecutive output elements x 8 split groups; warp g
// sums splits g, g+8, ... of its 32 elements (coalesced 128-byte loads, four independent partial sums in flight), the
// eight group sums are folded in group order through shared memory (deterministic).
__global__ void splitk_reduce8_kernel(const float* __restrict__ ws, float* __restrict__ C, int M, int N, int ldc,
                                      int splits, const float* __restrict__ bias, int relu,
                                      const float* __restrict__ mask, long c_plane) {
    pdl_sync();
    __shared__ float part[8][33];
    const size_t total = (size_t)M * N;
    const int e = threadIdx.x & 31, g = threadIdx.x >> 5;
    for (size_t c0 = (size_t)blockIdx.x * 32; c0 < total; c0 += (size_t)gridDim.x * 32) {
        const size_t i = c0 + e;
        const bool ok = i < total;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (ok) {
            int z = g;
            for (; z + 24 < splits; z += 32) {
                a0 += ws[(size_t)z * total + i]; a1 += ws[(size_t)(z + 8) * total + i];
                a2 += ws[(size_t)(z + 16) * total + i]; a3 += ws[(size_t)(z + 24) * total + i];
            }
            for (; z < splits; z += 8) a0 += ws[(size_t)z * total + i];
        }
        __syncthreads();  // the previous chunk's readers are done with `part`
        part[g][e] = (a0 + a1) + (a2 + a3);
        __syncthreads();
        if (g == 0 && ok) {
            float s = part[0][e];
#pragma unroll
            for (int k = 1; k < 8; ++k) s += part[k][e];
            int m = (int)(i / N), n = (int)(i % N);
            if (bias) s += bias[n];
            if (relu) s = fmaxf(s, 0.f);
            if (mask) s = mask[(size_t)m * ldc + n] > 0.f ? s : 0.f;
            C[(size_t)m * ldc + n] = s;
            if (c_plane) C[c_plane + (size_t)m * ldc + n] = lo_of(s);
        }
    }
}


template <int BM, int BN, int TM, int TN>
static void launch_cfg(GemmMode mode, const GemmArgs& a, dim3 grid, cudaStream_t s) {
    constexpr int NT = (BM / TM) * (BN / TN);
    switch (mode) {
        case G_FWD: launch_pdl(gemm_kernel<BM, BN, TM, TN, true, true, false, false>, grid, dim3(NT), 0, s, a); break;
        case G_FWD_U8: launch_pdl(gemm_kernel<BM, BN, TM, TN, true, true, true, false>, grid, dim3(NT), 0, s, a); break;
        case G_NN: launch_pdl(gemm_kernel<BM, BN, TM, TN, true, false, false, false>, grid, dim3(NT), 0, s, a); break;
        case G_WGRAD: launch_pdl(gemm_kernel<BM, BN, TM, TN, false, false, false, false>, grid, dim3(NT), 0, s, a); break;
        case G_WGRAD_U8: launch_pdl(gemm_kernel<BM, BN, TM, TN, false, false, false, true>, grid, dim3(NT), 0, s, a); break;
        case G_WGRAD_AU8: throw Error("G_WGRAD_AU8 exists on the tcgen05 path only");
    }
    BB_LAUNCHED();
}

// Tuning knobs (read once): BB_GEMM_VARIANT selects the register-tile shape per CTA-tile family,
// BB_GEMM_FILL the number of CTAs per SM that split-K aims for.
static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

// ---- skinny contractions: one dimension <= 16 (the A-wide output layer of a Q net, the 1-wide critic output, the
// 2x8 actor heads).  The tile kernels need split-K + a reduce launch to fill the GPU on these (9 + 4 us for the DQN
// output layer, 22 + 13 us for the SAC critic's); a warp (or a thread) per output element does them in one ~3 us
// launch.  Dense fp32 operands only; same epilogue semantics (bias, ReLU, mask).
// forward: C[m][n] = act(sum_k A[m*lda+k] B[n*ldb+k] + bias[n]), N <= 16: a warp per output element
__global__ void __launch_bounds__(256) skinny_fwd_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                         float* __restrict__ C, int M, int N, int K, long lda, long ldb,
                                                         int ldc, const float* __restrict__ bias, int relu,
                                                         const float* __restrict__ mask, long c_plane) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const long gw = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = (long)gridDim.x * (blockDim.x >> 5);
    for (long o = gw; o < (long)M * N; o += nw) {
        const int m = (int)(o / N), n = (int)(o % N);
        const float* a = A + (size_t)m * lda;
        const float* b = B + (size_t)n * ldb;
        float acc = 0.f;
        for (int k = lane; k < K; k += 32) acc = fmaf(a[k], b[k], acc);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        if (lane == 0) {
            if (bias) acc += bias[n];
            if (relu) acc = fmaxf(acc, 0.f);
            if (mask) acc = mask[(size_t)m * ldc + n] > 0.f ? acc : 0.f;
            C[(size_t)m * ldc + n] = acc;
            if (c_plane) C[c_plane + (size_t)m * ldc + n] = lo_of(acc);
        }
    }
}
// data gradient: C[m][n] = (sum_k A[m*lda+k] B[k*ldb+n]) * (mask > 0), K <= 16: a thread per output element
__global__ void __launch_bounds__(256) skinny_dgrad_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                           float* __restrict__ C, int M, int N, int K, long lda, long ldb,
                                                           int ldc, const float* __restrict__ mask, long c_plane) {
    pdl_sync();
    const size_t total = (size_t)M * N;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(i / N), n = (int)(i % N);
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc = fmaf(A[(size_t)m * lda + k], B[(size_t)k * ldb + n], acc);
        if (mask) acc = mask[(size_t)m * ldc + n] > 0.f ? acc : 0.f;
        C[(size_t)m * ldc + n] = acc;
        if (c_plane) C[c_plane + (size_t)m * ldc + n] = lo_of(acc);
    }
}
// weight gradient: C[m][n] = sum_k A[k*lda+m] B[k*ldb+n], M <= 16: 8 columns x 32 k-groups per CTA (N/8 CTAs: with 32
// columns per CTA a 256-wide layer kept only 8 SMs busy, 18-23 us), the skinny operand staged in shared memory, partial
// sums folded in k-group order (deterministic)
__global__ void __launch_bounds__(256) skinny_wgrad_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                           float* __restrict__ C, int M, int N, int K, long lda, long ldb,
                                                           int ldc) {
    pdl_sync();
    __shared__ float part[32][16][9];
    __shared__ float sa[256][16];   // one 256-row slab of A, reused by all 8 columns
    const int col = threadIdx.x & 7, kg = threadIdx.x >> 3;
    const int n = blockIdx.x * 8 + col;
    float acc[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) acc[m] = 0.f;
    for (int k0 = 0; k0 < K; k0 += 256) {
        const int kn = min(256, K - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < kn * M; i += 256) sa[i / M][i % M] = A[(size_t)(k0 + i / M) * lda + i % M];
        __syncthreads();
        if (n < N) {
            // all (up to 8) k of this thread in one trip: independent loads in flight
            float b[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) b[u] = (kg + 32 * u < kn) ? B[(size_t)(k0 + kg + 32 * u) * ldb + n] : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (kg + 32 * u < kn) {
#pragma unroll
                    for (int m = 0; m < 16; ++m)
                        if (m < M) acc[m] = fmaf(sa[kg + 32 * u][m], b[u], acc[m]);
                }
        }
    }
#pragma unroll
    for (int m = 0; m < 16; ++m) part[kg][m][col] = acc[m];
    __syncthreads();
    // thread (col, kg) finishes row m = kg
    if (kg < M && n < N) {
        float t = 0.f;
#pragma unroll
        for (int g2 = 0; g2 < 32; ++g2) t += part[g2][kg][col];
        C[(size_t)kg * ldc + n] = t;
    }
}

static bool gemm_skinny(const Ctx& c, GemmMode mode, const GemmArgs& a) {
    if (!env_int("BB_SKINNY", 1)) return false;
    if (a.a_rowbase || a.a_koff || a.b_rowbase || a.b_noff || a.trans_out || a.c_rowoff) return false;
    const float* A = reinterpret_cast<const float*>(a.A);
    const float* B = reinterpret_cast<const float*>(a.B);
    if (mode == G_FWD && a.N <= 16 && a.K >= 32) {
        long warps = (long)a.M * a.N;
        int blocks = (int)std::min<long>((warps + 7) / 8, (long)c.sms * 8);
        launch_pdl(skinny_fwd_kernel, dim3(blocks), dim3(256), 0, c.stream, A, B, a.C, a.M, a.N, a.K, a.lda, a.ldb, a.ldc,
                   a.bias, a.relu, a.mask, a.c_plane);
        BB_LAUNCHED();
        c.mark("skinny_fwd");
        return true;
    }
    if (mode == G_NN && a.K <= 16 && !a.bias && !a.relu) {
        size_t total = (size_t)a.M * a.N;
        int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)c.sms * 8);
        launch_pdl(skinny_dgrad_kernel, dim3(blocks), dim3(256), 0, c.stream, A, B, a.C, a.M, a.N, a.K, a.lda, a.ldb, a.ldc,
                   a.mask, a.c_plane);
        BB_LAUNCHED();
        c.mark("skinny_dgrad");
        return true;
    }
    // (short contractions only: each CTA walks all of K; IQN's 4 x 512 x 16384 took 486 us here against 25 + 8 us as split-K tiles)
    if (mode == G_WGRAD && a.M <= 16 && a.K <= 2048 && !a.bias && !a.relu && !a.mask) {
        launch_pdl(skinny_wgrad_kernel, dim3((a.N + 7) / 8), dim3(256), 0, c.stream, A, B, a.C, a.M, a.N, a.K, a.lda, a.ldb,
                   a.ldc);
        BB_LAUNCHED();
        c.mark("skinny_wgrad");
        return true;
    }
    return false;
}

void gemm(const Ctx& c, GemmMode mode, GemmArgs a) {
    if (a.M <= 0 || a.N <= 0) return;
    if (gemm_skinny(c, mode, a)) return;
    const int use_tc = env_int("BB_TC", 1);  // read per call so tests can flip it (0 = fp32 CUDA-core tiles)
    // tcgen05 paths; tiny problems (policy forward, the 6-wide output layer) stay on CUDA cores.  TMA-fed kernel
    // (tma_gemm.cu) when both operands carry lo planes, else the SIMT-producer kernel (tc_gemm.cu).
    const bool tc_size = a.M >= 64 && (long)a.M * a.N * a.K >= (1L << 22) && a.N >= 16;
    if (use_tc && tc_size && tma_gemm(c, mode, a)) return;
    if (use_tc && tc_size && tc_gemm(c, mode, a)) return;
    gemm_simt(c, mode, a);
}

// fp32 CUDA-core path (gemm.cuh): the parity reference for the tensor-core path and the fallback
// for tiny M.
void gemm_simt(const Ctx& c, GemmMode mode, GemmArgs a) {
    if (a.M <= 0 || a.N <= 0) return;
    static const int variant = env_int("BB_GEMM_VARIANT", 0);  // A/B on B200 (profiles/r01_summary.md): 0 wins
    static const int fill = env_int("BB_GEMM_FILL", 3);
    int BM, BN, cfg;
    if (a.N <= 32) { cfg = 1; BM = 128; BN = 32; }
    else if (a.M <= 32) { cfg = 2; BM = 32; BN = 128; }
    else if (variant == 3 && a.M >= 4096) { cfg = 3; BM = 128; BN = 64; }
    else { cfg = 0; BM = 64; BN = 64; }
    int tm = (a.M + BM - 1) / BM, tn = (a.N + BN - 1) / BN;
    long tiles = (long)tm * tn;
    int split = 1;
    int kt = (a.K + kBK - 1) / kBK;  // k tiles
    long want = (long)fill * c.sms;
    if (tiles < want && kt >= 8) {  // (raising the threshold to K >= 512 made the SAC step slower: 368 -> 457 us; the k loop, not the reduce launch, is the cost)
        split = (int)std::min<long>((want + tiles - 1) / tiles, kt / 4);
        size_t per = (size_t)a.M * a.N;
        const size_t usable = c.ws_floats - 1024;  // the last 1024 words hold colsum's block counters
        if (per * split > usable) split = (int)(usable / per);
        if (split < 1) split = 1;
    }
    int kps = ((kt + split - 1) / split) * kBK;
    split = (a.K + kps - 1) / kps;
    a.split_k = split;
    a.k_per_split = kps;
    a.workspace = c.ws;
    dim3 grid(tn, tm, split);
    const char* tag = "gemm";
    if (variant == 0) {
        switch (cfg) {
            case 0: launch_cfg<64, 64, 4, 4>(mode, a, grid, c.stream); tag = "gemm64x64"; break;
            case 1: launch_cfg<128, 32, 4, 4>(mode, a, grid, c.stream); tag = "gemm128x32"; break;
            case 2: launch_cfg<32, 128, 4, 4>(mode, a, grid, c.stream); tag = "gemm32x128"; break;
        }
    } else if (variant == 2) {
        switch (cfg) {
            case 0: launch_cfg<64, 64, 8, 8>(mode, a, grid, c.stream); tag = "gemm64x64"; break;
            case 1: launch_cfg<128, 32, 8, 8>(mode, a, grid, c.stream); tag = "gemm128x32"; break;
            case 2: launch_cfg<32, 128, 8, 8>(mode, a, grid, c.stream); tag = "gemm32x128"; break;
        }
    } else {
        switch (cfg) {
            case 0: launch_cfg<64, 64, 8, 4>(mode, a, grid, c.stream); tag = "gemm64x64"; break;
            case 1: launch_cfg<128, 32, 8, 4>(mode, a, grid, c.stream); tag = "gemm128x32"; break;
            case 2: launch_cfg<32, 128, 4, 8>(mode, a, grid, c.stream); tag = "gemm32x128"; break;
            case 3: launch_cfg<128, 64, 8, 8>(mode, a, grid, c.stream); tag = "gemm128x64"; break;
        }
    }
    c.mark(tag);
    if (split > 1) {
        size_t total = (size_t)a.M * a.N;
        if (split >= 16) {
            int blocks = (int)std::min<size_t>((total * 8 + 255) / 256, (size_t)c.sms * 8);
            launch_pdl(splitk_reduce8_kernel, dim3(blocks), dim3(256), 0, c.stream, c.ws, a.C, a.M, a.N, a.ldc, split, a.bias, a.relu, a.mask, a.c_plane);
        } else {
            int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)c.sms * 8);
            launch_pdl(splitk_reduce_kernel, dim3(blocks), dim3(256), 0, c.stream, c.ws, a.C, a.M, a.N, a.ldc, split, a.bias, a.relu, a.mask, a.c_plane);
        }
        BB_LAUNCHED();
        c.mark("splitk_reduce");
    }
}

GemmArgs zero_args() {
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.split_k = 1;
    return a;
}

// ------------------------------------------------------------------------------- column sums

__global__ void colsum_partial_kernel(const float* __restrict__ Y, float* __restrict__ part, int M, int N,
                                      int rows_per_block) {
    __shared__ float s[8][33];
    int n = blockIdx.x * 32 + threadIdx.x;
    int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
    float acc = 0.f;
    if (n < N)
        for (int m = r0 + threadIdx.y; m < r1; m += 8) acc += Y[(size_t)m * N + n];
    s[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += s[i][threadIdx.x];
        part[(size_t)blockIdx.y * N + n] = t;
    }
}
// Column sums in ONE launch: block (chunk, r) sums its row slice of 32 columns; the last block of a
// chunk to finish (counter in the workspace tail) adds the R partials in index order (deterministic).
__global__ void colsum_fused_kernel(const float* __restrict__ Y, float* __restrict__ part, unsigned int* __restrict__ counters,
                                    float* __restrict__ out, int M, int N, int rows_per_block, int R) {
    pdl_sync();
    __shared__ float s[8][33];
    __shared__ bool s_last;
    int n = blockIdx.x * 32 + threadIdx.x;
    int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
    float acc = 0.f;
    if (n < N) {
        // four independent partial sums: four loads in flight per thread (the single chain paid one load latency per row:
        // 22 us for the 102400 x 32 conv1 bias gradient)
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int m = r0 + threadIdx.y;
        for (; m + 24 < r1; m += 32) {
            a0 += Y[(size_t)m * N + n]; a1 += Y[(size_t)(m + 8) * N + n];
            a2 += Y[(size_t)(m + 16) * N + n]; a3 += Y[(size_t)(m + 24) * N + n];
        }
        for (; m < r1; m += 8) a0 += Y[(size_t)m * N + n];
        acc = (a0 + a1) + (a2 + a3);
    }
    s[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += s[i][threadIdx.x];
        part[(size_t)blockIdx.y * N + n] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) s_last = (atomicAdd(&counters[blockIdx.x], 1u) == (unsigned)R - 1u);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // 8 row-groups x 32 columns: thread (x, y) sums partials y, y+8, ... of column n, then the 8 groups in order
    float t = 0.f;
    if (n < N)
        for (int r = threadIdx.y; r < R; r += 8) t += part[(size_t)r * N + n];
    s[threadIdx.y][threadIdx.x] = t;
    __syncthreads();
    if (threadIdx.y == 0 && n < N) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) v += s[i][threadIdx.x];
        out[n] = v;
    }
    if (threadIdx.x == 0 && threadIdx.y == 0) counters[blockIdx.x] = 0;
}

void colsum(const Ctx& c, const float* dY, float* db, int M, int N) {
    int chunks = (N + 31) / 32;
    int R = std::max(1, std::min(std::min(256, (M + 31) / 32), (2 * c.sms + chunks - 1) / chunks));
    const size_t counter_floats = 1024;  // tail of the workspace holds the per-chunk counters (zeroed at creation)
    while ((size_t)R * N > c.ws_floats - counter_floats && R > 1) R /= 2;
    int rpb = (M + R - 1) / R;
    R = (M + rpb - 1) / rpb;
    unsigned int* counters = reinterpret_cast<unsigned int*>(c.ws + c.ws_floats - counter_floats);
    BB_CHECK(chunks <= (int)counter_floats, "colsum: too many column chunks");
    launch_pdl(colsum_fused_kernel, dim3(chunks, R), dim3(32, 8), 0, c.stream, dY, c.ws, counters, db, M, N, rpb, R);
    BB_LAUNCHED();
    c.mark("colsum");
}

// ------------------------------------------------------------------------------- linear

void linear_fwd(const Ctx& c, const float* X, long ldx, const float* W, const float* b, float* Y, int M, int N, int K,
                bool relu, long x_plane, long w_plane, long y_plane) {
    GemmArgs a = zero_args();
    a.A = X; a.lda = ldx; a.B = W; a.ldb = K; a.C = Y; a.ldc = N; a.M = M; a.N = N; a.K = K;
    a.bias = b; a.relu = relu;
    a.a_plane = x_plane; a.b_plane = w_plane; a.c_plane = y_plane;
    gemm(c, G_FWD, a);
}

void linear_bwd_data(const Ctx& c, const float* dY, const float* W, float* dX, long lddx, int M, int N, int K,
                     const float* mask, long dy_plane, long w_plane, long dx_plane) {
    // dX[M][K] = dY[M][N] * W[N][K]
    GemmArgs a = zero_args();
    a.A = dY; a.lda = N; a.B = W; a.ldb = K; a.C = dX; a.ldc = (int)lddx; a.M = M; a.N = K; a.K = N;
    a.mask = mask;
    a.a_plane = dy_plane; a.b_plane = w_plane; a.c_plane = dx_plane;
    gemm(c, G_NN, a);
}

void linear_bwd_weight(const Ctx& c, const float* dY, const float* X, long ldx, float* dW, float* db, int M, int N,
                       int K, long dy_plane, long x_plane) {
    // dW[N][K] = sum_m dY[m][N]^T X[m][K]
    GemmArgs a = zero_args();
    a.A = dY; a.lda = N; a.B = X; a.ldb = ldx; a.C = dW; a.ldc = K; a.M = N; a.N = K; a.K = M;
    a.a_plane = dy_plane; a.b_plane = x_plane;
    gemm(c, G_WGRAD, a);
    if (db) colsum(c, dY, db, M, N);
}

// ------------------------------------------------------------------------------- conv

static TmaConv tma_conv_of(const ConvGeom& g) { return TmaConv{g.B, g.H, g.W, g.C, g.KH, g.KW, g.S, 0}; }

void conv_fwd(const Ctx& c, const ConvGeom& g, const void* X, const float* W, const float* b, float* Y, bool relu) {
    if (g.u8_chw && env_int("BB_TC", 1) && conv1_fwd_tc(c, g, X, W, b, Y, relu)) return;
    BB_CHECK(!g.in_ix, "indexed input reached the generic convolution path");
    GemmArgs a = zero_args();
    const TmaConv cv = tma_conv_of(g);
    if (!g.u8_chw) { a.a_conv = &cv; a.a_plane = g.x_plane; a.b_plane = g.w_plane; }
    a.c_plane = g.y_plane;
    a.A = X; a.a_rowbase = g.rowbase; a.a_koff = g.koff; a.B = W; a.ldb = g.K(); a.C = Y; a.ldc = g.OC;
    a.M = g.M(); a.N = g.OC; a.K = g.K(); a.bias = b; a.relu = relu;
    a.tables_vec4 = !g.u8_chw && g.C % 4 == 0;
    gemm(c, g.u8_chw ? G_FWD_U8 : G_FWD, a);
}

void conv_bwd_weight(const Ctx& c, const ConvGeom& g, const float* dY, const void* X, float* dW, float* db) {
    // dW[OC][K] = sum_m dY[m][OC]^T im2col(X)[m][K]
    if (g.u8_chw && env_int("BB_TC", 1) && conv1_wgrad_tc(c, g, dY, X, dW)) {  // dedicated kernel (conv1_tc.cu)
        if (db) colsum(c, dY, db, g.M(), g.OC);
        return;
    }
    BB_CHECK(!g.in_ix, "indexed input reached the generic convolution path");
    // (float inputs only: for the u8 conv1 frames the transposed 4-byte gathers are slower than the
    // CUDA-core kernel -- 115 us vs 67 us at B=256)
    if (!g.u8_chw && env_int("BB_TC", 1) && env_int("BB_TC_WGRAD_T", 1)) {
        // tensor-core form: dW^T[K][OC] = im2col(X)^T dY, stored transposed.  The long K = KH*KW*C axis
        // fills the 128 MMA rows (OC is only 32 / 64), A is the transposed gather (m-part from koff,
        // k-part from rowbase), B = dY rows.
        GemmArgs t = zero_args();
        t.A = X; t.a_rowbase = g.koff; t.a_koff = g.rowbase; t.B = dY; t.ldb = g.OC; t.C = dW; t.ldc = g.K();
        t.M = g.K(); t.N = g.OC; t.K = g.M(); t.trans_out = 1;
        t.tables_vec4 = g.C % 4 == 0;
        const TmaConv cv = tma_conv_of(g);
        t.a_conv = &cv; t.a_plane = g.x_plane; t.b_plane = g.y_plane;
        if (tma_gemm(c, G_WGRAD, t) || tc_gemm(c, g.u8_chw ? G_WGRAD_AU8 : G_WGRAD, t)) {
            if (db) colsum(c, dY, db, g.M(), g.OC);
            return;
        }
    }
    GemmArgs a = zero_args();
    a.A = dY; a.lda = g.OC; a.B = X; a.b_rowbase = g.rowbase; a.b_noff = g.koff; a.C = dW; a.ldc = g.K();
    a.M = g.OC; a.N = g.K(); a.K = g.M();
    gemm(c, g.u8_chw ? G_WGRAD_U8 : G_WGRAD, a);
    if (db) colsum(c, dY, db, g.M(), g.OC);
}

// dX[b][h][w][c] = sum over the kernel taps that touch (h, w) of col[(b,oh,ow)][(kh,kw,c)]
__global__ void col2im_nhwc_kernel(const float* __restrict__ col, float* __restrict__ dX, const float* __restrict__ mask,
                                   int B, int C, int H, int W, int KH, int KW, int S, int OH, int OW, long dx_plane) {
    pdl_sync();
    const int C4 = C / 4;
    size_t total = (size_t)B * H * W * C4;
    const int K = KH * KW * C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int c4 = (int)(i % C4);
        size_t t = i / C4;
        int w = (int)(t % W); t /= W;
        int h = (int)(t % H);
        int b = (int)(t / H);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int kh = h % S; kh < KH; kh += S) {
            int oh = (h - kh) / S;
            if (h - kh < 0 || oh >= OH) continue;
            for (int kw = w % S; kw < KW; kw += S) {
                int ow = (w - kw) / S;
                if (w - kw < 0 || ow >= OW) continue;
                size_t m = ((size_t)b * OH + oh) * OW + ow;
                float4 v = __ldg(reinterpret_cast<const float4*>(col + m * K + (kh * KW + kw) * C) + c4);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        size_t o = (((size_t)b * H + h) * W + w) * C + c4 * 4;
        if (mask) {
            float4 mk = __ldg(reinterpret_cast<const float4*>(mask + o));
            acc.x = mk.x > 0.f ? acc.x : 0.f; acc.y = mk.y > 0.f ? acc.y : 0.f;
            acc.z = mk.z > 0.f ? acc.z : 0.f; acc.w = mk.w > 0.f ? acc.w : 0.f;
        }
        *reinterpret_cast<float4*>(dX + o) = acc;
        if (dx_plane) *reinterpret_cast<float4*>(dX + dx_plane + o) = make_float4(lo_of(acc.x), lo_of(acc.y), lo_of(acc.z), lo_of(acc.w));
    }
}

// Operand preparation of the gather-form data gradient, one launch:
//   blocks [0, pad_blocks): interior of the zero-padded dY copy, dypad[b][oh + Jh-1][ow + Jw-1][:] = dY[b][oh][ow][:]
//   the rest:               Wt[n][k] = W[brow[k] + bnoff[n]]  (weights re-laid so that the GEMM's B operand is k-contiguous)
__global__ void dgrad_prep_kernel(const float4* __restrict__ dY, float4* __restrict__ pad, int B, int OH, int OW, int OC4,
                                  int OHp, int OWp, int offh, int offw, int pad_blocks, const float* __restrict__ W,
                                  float* __restrict__ Wt, const int* __restrict__ brow, const int* __restrict__ bnoff,
                                  int N, int K, long dy_plane4, long pad_plane4, long w_plane, long wt_plane) {
    if ((int)blockIdx.x < pad_blocks) {
        size_t total = (size_t)B * OH * OW * OC4;
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)pad_blocks * blockDim.x) {
            int c = (int)(i % OC4);
            size_t t = i / OC4;
            int ow = (int)(t % OW); t /= OW;
            int oh = (int)(t % OH);
            size_t b = t / OH;
            const size_t o = ((b * OHp + oh + offh) * OWp + ow + offw) * OC4 + c;
            pad[o] = dY[i];
            if (pad_plane4) pad[pad_plane4 + o] = dY[dy_plane4 + i];   // the lo plane travels with it
        }
    } else {
        const int wb = gridDim.x - pad_blocks;
        const int total = N * K;
        for (int i = (blockIdx.x - pad_blocks) * blockDim.x + threadIdx.x; i < total; i += wb * blockDim.x) {
            int k = i % K, n = i / K;
            const int src = brow[k] + bnoff[n];
            Wt[i] = W[src];
            if (wt_plane) Wt[wt_plane + i] = w_plane ? W[w_plane + src] : lo_of(W[src]);
        }
    }
}

bool dgrad_gather_path(const ConvGeom& g) {
    return g.dypad && g.dg_rowbase && g.dg_wt && env_int("BB_TC", 1) && env_int("BB_DGRAD_GATHER", 1);
}

// Wt[n][k] = W[brow[k] + bnoff[n]]: the weights of the gather-form data gradient, k-contiguous (the weight-only blocks of
// dgrad_prep_kernel).  Depends on the parameters only, so it runs beside the start of the backward pass.
void conv_dgrad_prepare_weights(const Ctx& c, const ConvGeom& g, const float* W) {
    const int Jh = g.KH / g.S, Jw = g.KW / g.S;
    const int N = g.S * g.S * g.C, K = Jh * Jw * g.OC;
    const int w_blocks = std::min((N * K + 255) / 256, c.sms * 2);
    dgrad_prep_kernel<<<w_blocks, 256, 0, c.stream>>>(nullptr, nullptr, g.B, g.OH, g.OW, g.OC / 4, g.dg_hp(), g.dg_wp(), Jh - 1, Jw - 1,
                                                      0, W, g.dg_wt, g.dg_brow, g.dg_bnoff, N, K, 0, 0, g.w_plane, g.wt_plane);
    BB_LAUNCHED();
    c.mark("dgrad_weights");
}

void conv_bwd_data(const Ctx& c, const ConvGeom& g, const float* dY, const float* W, float* col, float* dX,
                   const float* mask) {
    BB_CHECK(!g.u8_chw && g.C % 4 == 0, "conv_bwd_data expects an NHWC float input with C % 4 == 0");
    // Default since the whole-tile producer path of the tcgen05 kernel: DQN step 415 -> 382 us at B = 256 (c2: 9 + 41 us
    // instead of 46 + 22 us for dY*W + col2im, c3: 9 + 36 us instead of 35 + 18 us).  BB_DGRAD_GATHER=0 keeps the col2im path.
    if (dgrad_gather_path(g)) {
        // dX[b][S hq + ph][S wq + pw][c] = sum_{jh,jw,oc} dYpad[b][hq + Jh-1-jh][wq + Jw-1-jw][oc] W[oc][S jh + ph][S jw + pw][c]:
        // one tcgen05 GEMM, no [M][K] column buffer and no col2im pass (the buffer was 42 MB for c2 at B = 256)
        const int Jh = g.KH / g.S, Jw = g.KW / g.S;
        const int N = g.S * g.S * g.C, K = Jh * Jw * g.OC;
        if (!g.wt_ready) conv_dgrad_prepare_weights(c, g, W);   // (Net::backward does it on a side stream at the start of the pass)
        GemmArgs a = zero_args();
        a.a_rowbase = g.dg_rowbase; a.a_koff = g.dg_koff;
        a.B = g.dg_wt; a.ldb = K;
        a.C = dX; a.ldc = N; a.c_rowoff = g.dg_crow; a.c_coloff = g.dg_ccol; a.mask = mask;
        a.M = g.B * (g.H / g.S) * (g.W / g.S); a.N = N; a.K = K;
        a.tables_vec4 = 1;  // OC % 4 == 0 and C % 4 == 0 (dgrad_gather_ok)
        a.b_plane = g.wt_plane; a.c_plane = g.dx_plane;
        // TMA form 1: a stride-1 im2col walk of dY ITSELF with the tap offsets running backwards; the zero halo of the
        // transposed convolution is the tensor map's out-of-bounds fill -- no padded copy of dY (DQN step 283.6 -> 282.0 us;
        // BB_DGRAD_DIRECT=0 restores the copy)
        if (Jh == Jw && env_int("BB_DGRAD_DIRECT", 1)) {
            const TmaConv cv{g.B, g.OH, g.OW, g.OC, Jh, Jw, 1, 1, Jh - 1};
            a.A = dY; a.a_conv = &cv; a.a_plane = g.y_plane;
            if (tma_gemm(c, G_FWD, a)) return;
        }
        // form 2 / SIMT-fed fallback: over a zero-padded copy of dY
        size_t total = (size_t)g.B * g.OH * g.OW * (g.OC / 4);
        int pad_blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)c.sms * 6);
        dgrad_prep_kernel<<<pad_blocks, 256, 0, c.stream>>>(
            reinterpret_cast<const float4*>(dY), reinterpret_cast<float4*>(g.dypad), g.B, g.OH, g.OW, g.OC / 4, g.dg_hp(),
            g.dg_wp(), Jh - 1, Jw - 1, pad_blocks, W, g.dg_wt, g.dg_brow, g.dg_bnoff, N, K,
            g.y_plane && g.dypad_plane ? g.y_plane / 4 : 0, g.y_plane && g.dypad_plane ? g.dypad_plane / 4 : 0, g.w_plane, g.wt_plane);
        BB_LAUNCHED();
        c.mark("dgrad_prep");
        const TmaConv cv{g.B, g.dg_hp(), g.dg_wp(), g.OC, Jh, Jw, 1, 1, 0};
        a.A = g.dypad; a.a_conv = &cv; a.a_plane = g.dypad_plane;
        if (tma_gemm(c, G_FWD, a)) return;
        if (tc_gemm(c, G_FWD, a)) return;
    }
    BB_CHECK(col != nullptr, "conv_bwd_data: no column buffer for the col2im path");
    GemmArgs a = zero_args();
    a.A = dY; a.lda = g.OC; a.B = W; a.ldb = g.K(); a.C = col; a.ldc = g.K(); a.M = g.M(); a.N = g.K(); a.K = g.OC;
    a.a_plane = g.y_plane; a.b_plane = g.w_plane;
    gemm(c, G_NN, a);
    size_t total = (size_t)g.B * g.H * g.W * (g.C / 4);
    int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)c.sms * 16);
    launch_pdl(col2im_nhwc_kernel, dim3(blocks), dim3(256), 0, c.stream, col, dX, mask, g.B, g.C, g.H, g.W, g.KH, g.KW, g.S, g.OH, g.OW, g.dx_plane);
    BB_LAUNCHED();
    c.mark("col2im");
}

// ------------------------------------------------------------------------------- Adam / track

struct PeerPtrs { const float* p[8]; };

// torch::optim::Adam::step (what tch's nn::Adam binds, opt.rs:35,77):
//   m = m*b1 + (1-b1) g ; v = v*b2 + (1-b2) g*g ; denom = sqrt(v)/sqrt(1-b2^t) + eps ;
//   p = p - lr/(1-b1^t) * m/denom.     AdamW: p *= (1 - lr*wd) first; Adam: g += wd*p.
// With world > 1 the gradient is the mean over the ranks' buffers read through peer pointers
// (NVLink P2P loads): all-reduce and optimizer step in one kernel, no parameter broadcast needed.
__device__ __forceinline__ void adam_elem(float& pi, float gi, float& mi, float& vi, float b1, float b2, float one_m_b1,
                                          float one_m_b2, float eps, float bc2_sqrt, float neg_step, float wd, float decay,
                                          int adamw) {
    if (adamw) pi = pi * decay;
    else if (wd != 0.f) gi = gi + wd * pi;
    mi = mi * b1 + one_m_b1 * gi;
    vi = vi * b2 + one_m_b2 * gi * gi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi = pi + neg_step * (mi / denom);
}

// Mean over ranks as a pairwise tree in rank order, ((g0+g1)+(g2+g3))+..., then / world: the same value on every
// rank, and for a power-of-two world of IDENTICAL gradients exactly the gradient itself (2g, 4g, 8g and the
// division are exact), which is what lets tests/mgpu_check.py compare a data-parallel run with one GPU bit for bit.
__device__ __forceinline__ float4 mean_over_ranks4(const PeerPtrs& peers, size_t i, int world) {
    float4 t[8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
        t[r] = r < world ? __ldcv(reinterpret_cast<const float4*>(peers.p[r]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 1; s < 8; s *= 2)
#pragma unroll
        for (int r = 0; r + s < 8; r += 2 * s)
            if (r + s < world) { t[r].x += t[r + s].x; t[r].y += t[r + s].y; t[r].z += t[r + s].z; t[r].w += t[r + s].w; }
    const float w = (float)world;
    return make_float4(t[0].x / w, t[0].y / w, t[0].z / w, t[0].w / w);
}
__device__ __forceinline__ float mean_over_ranks1(const PeerPtrs& peers, size_t i, int world) {
    float t[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) t[r] = r < world ? peers.p[r][i] : 0.f;
#pragma unroll
    for (int s = 1; s < 8; s *= 2)
#pragma unroll
        for (int r = 0; r + s < 8; r += 2 * s)
            if (r + s < world) t[r] += t[r + s];
    return t[0] / (float)world;
}

__device__ __forceinline__ unsigned int gtimer_lo() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return (unsigned int)t;
}

// Spin until every rank's slot of `kind` in this rank's flag array has reached `epoch` (threads 0..world-1 of the block).
// No fence here: a system-scope fence costs 3-7 us on a busy GPU (measured through the %globaltimer stamps below), and the
// data the flags guard is read with L1-bypassing loads (__ldcv) issued after the barrier -- the writer fenced before it
// raised the flag, so by the time the flag is visible the data is in the owner's L2, where peer and local loads both go.
__device__ __forceinline__ void wait_flags(const Exchange& x, int kind, unsigned int epoch) {
    if ((int)threadIdx.x < x.world) {
        volatile unsigned int* f = x.flags[x.rank] + xflag(kind, (int)threadIdx.x);
        const long long t0 = clock64();
        while ((int)(*f - epoch) < 0) {
            if (clock64() - t0 > x.timeout_cycles) {   // a peer never arrived: raise the sticky failure flag, do not hang
                *reinterpret_cast<volatile int*>(x.err) = 21;
                __threadfence_system();
                break;
            }
        }
    }
    __syncthreads();
}

// n4 = n / 4 float4 groups (every tensor of the flat vector is 16 B aligned and padded), tail scalars after.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float b1, float b2, float one_m_b1, float one_m_b2,
                            float eps, float bc2_sqrt, float neg_step, float wd, float decay, int adamw,
                            PeerPtrs peers, int world, float* __restrict__ p_lo, Exchange xw, int wait_regions,
                            const AdamScalars* __restrict__ dev_scalars) {
    if (dev_scalars) { bc2_sqrt = dev_scalars->bc2_sqrt; neg_step = dev_scalars->neg_step; decay = dev_scalars->decay; }
    if (wait_regions) {   // every rank's slice of the mean gradient has landed in this rank's buffer
        if (blockIdx.x == 0 && threadIdx.x == 0) xw.ctr[16] = gtimer_lo();
        if (wait_regions & 1) wait_flags(xw, 1, xw.ctr[0]);
        if (wait_regions & 2) wait_flags(xw, 3, xw.ctr[1]);
        if (blockIdx.x == 0 && threadIdx.x == 0) xw.ctr[17] = gtimer_lo();
    }
    const size_t n4 = n >> 2;
    const float inv_world = 1.0f;
    (void)inv_world;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 gi;
        if (world > 1) {
            gi = mean_over_ranks4(peers, i, world);
        } else {
            // (after an exchange the buffer was written by the peers: read it past L1, see wait_flags)
            gi = wait_regions ? __ldcv(reinterpret_cast<const float4*>(g) + i) : reinterpret_cast<const float4*>(g)[i];
        }
        float4 pi = reinterpret_cast<float4*>(p)[i], mi = reinterpret_cast<float4*>(m)[i], vi = reinterpret_cast<float4*>(v)[i];
        adam_elem(pi.x, gi.x, mi.x, vi.x, b1, b2, one_m_b1, one_m_b2, eps, bc2_sqrt, neg_step, wd, decay, adamw);
        adam_elem(pi.y, gi.y, mi.y, vi.y, b1, b2, one_m_b1, one_m_b2, eps, bc2_sqrt, neg_step, wd, decay, adamw);
        adam_elem(pi.z, gi.z, mi.z, vi.z, b1, b2, one_m_b1, one_m_b2, eps, bc2_sqrt, neg_step, wd, decay, adamw);
        adam_elem(pi.w, gi.w, mi.w, vi.w, b1, b2, one_m_b1, one_m_b2, eps, bc2_sqrt, neg_step, wd, decay, adamw);
        reinterpret_cast<float4*>(m)[i] = mi; reinterpret_cast<float4*>(v)[i] = vi; reinterpret_cast<float4*>(p)[i] = pi;
        if (p_lo) reinterpret_cast<float4*>(p_lo)[i] = make_float4(lo_of(pi.x), lo_of(pi.y), lo_of(pi.z), lo_of(pi.w));
    }
    for (size_t i = (n4 << 2) + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float gi;
        if (world > 1) {
            gi = mean_over_ranks1(peers, i, world);
        } else {
            gi = wait_regions ? __ldcv(g + i) : g[i];
        }
        float pi = p[i], mi = m[i], vi = v[i];
        adam_elem(pi, gi, mi, vi, b1, b2, one_m_b1, one_m_b2, eps, bc2_sqrt, neg_step, wd, decay, adamw);
        m[i] = mi; v[i] = vi; p[i] = pi;
        if (p_lo) p_lo[i] = lo_of(pi);
    }
}

// Sharded gradient exchange for world >= 4: rank r owns float4 groups [r*per, (r+1)*per).  It reads
// that slice of EVERY rank's gradient through the peer pointers, forms the mean (mean_over_ranks4: the same
// arithmetic as the fused kernel, so all ranks end bit-identical) and stores it back into
// that slice of every rank's buffer.  Per GPU that is 2*(world-1)/world of the vector over NVLink instead of the
// (world-1) full vectors the all-read form moves (47 MB -> 11.8 MB at world = 8); a plain local Adam follows.
__global__ void grad_reduce_scatter_kernel(PeerPtrs peers, size_t n, int rank, int world) {
    const size_t n4 = n >> 2;
    const size_t per = (n4 + world - 1) / world;
    const size_t lo = (size_t)rank * per, hi = lo + per < n4 ? lo + per : n4;
    for (size_t i = lo + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < hi; i += (size_t)gridDim.x * blockDim.x) {
        const float4 gi = mean_over_ranks4(peers, i, world);
        for (int r = 0; r < world; ++r) reinterpret_cast<float4*>(const_cast<float*>(peers.p[r]))[i] = gi;
    }
    if (rank == 0)  // tail scalars (the flat vector is padded to float4 per tensor, so normally none)
        for (size_t i = (n4 << 2) + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            const float gi = mean_over_ranks1(peers, i, world);
            for (int r = 0; r < world; ++r) const_cast<float*>(peers.p[r])[i] = gi;
        }
}

// One launch per region.  Each thread keeps U float4 of every rank in flight (the peer loads are NVLink round trips of
// ~1 us: bandwidth comes from bytes in flight, not from loop iterations), reduces with the pairwise tree of mean_over_ranks4
// and stores the mean to every rank.  W (the world size) is a template parameter so that every rank index is static: with a
// runtime index the pointer table goes to local memory and ptxas serialises the loads (measured: 21 us per loop iteration).
// ctr[8 + 4*region ..] receive %globaltimer stamps (start / after the rendezvous / end) for bench.py's exchange trace.
template <int W>
__global__ void __launch_bounds__(256) grad_exchange_kernel(Exchange x, size_t lo4, size_t hi4, int region, int dbg) {
    constexpr int U = W <= 2 ? 8 : W <= 4 ? 4 : 2;
    __shared__ bool s_last;
    const unsigned int epoch = x.ctr[region] + 1;   // (bumped by the last block of THIS launch, after every block has read it)
    // (this rank's gradient kernels completed before this launch: their output is in this GPU's L2, where peer loads go)
    if (blockIdx.x == 0 && (int)threadIdx.x < W)
        *reinterpret_cast<volatile unsigned int*>(x.flags[threadIdx.x] + xflag(2 * region, x.rank)) = epoch;
    if (blockIdx.x == 0 && threadIdx.x == 0) x.ctr[8 + 4 * region] = gtimer_lo();
    wait_flags(x, 2 * region, epoch);
    if (blockIdx.x == 0 && threadIdx.x == 0) x.ctr[9 + 4 * region] = gtimer_lo();
    const size_t per = (hi4 - lo4 + W - 1) / W;
    const size_t a = lo4 + (size_t)x.rank * per, b = (dbg & 1) ? a : (a + per < hi4 ? a + per : hi4);   // dbg 1: rendezvous only (timing experiments)
    for (size_t base = a + (size_t)blockIdx.x * (256 * U); base < b; base += (size_t)gridDim.x * (256 * U)) {
        float4 t[U][W];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = base + u * 256 + threadIdx.x;
#pragma unroll
            for (int r = 0; r < W; ++r)
                t[u][r] = i < b ? __ldcv(reinterpret_cast<const float4*>(x.grads[r]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int s = 1; s < W; s *= 2)
#pragma unroll
                for (int r = 0; r + s < W; r += 2 * s) {
                    t[u][r].x += t[u][r + s].x; t[u][r].y += t[u][r + s].y; t[u][r].z += t[u][r + s].z; t[u][r].w += t[u][r + s].w;
                }
            const float w = (float)W;
            t[u][0] = make_float4(t[u][0].x / w, t[u][0].y / w, t[u][0].z / w, t[u][0].w / w);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = base + u * 256 + threadIdx.x;
            if (i < b) {
#pragma unroll
                for (int r = 0; r < W; ++r) reinterpret_cast<float4*>(const_cast<float*>(x.grads[r]))[i] = t[u][0];
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) x.ctr[11 + 4 * region] = gtimer_lo();
    // the ONE system fence of the exchange: this thread's peer stores have landed.  (acq_rel, not __threadfence_system():
    // that one is fence.sc.sys, which also joins the global order of all sc fences and costs several us more)
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&x.ctr[2 + region], 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;      // (every other block fenced before its ticket, and this block drew the last ticket after them)
    if ((int)threadIdx.x < W) *reinterpret_cast<volatile unsigned int*>(x.flags[threadIdx.x] + xflag(2 * region + 1, x.rank)) = epoch;
    if (threadIdx.x == 0) { x.ctr[2 + region] = 0u; x.ctr[region] = epoch; x.ctr[10 + 4 * region] = gtimer_lo(); }
}

void grad_exchange(const Ctx& c, const Exchange& x, size_t lo, size_t hi, int region, int max_blocks) {
    BB_CHECK((lo & 3) == 0 && (hi & 3) == 0, "grad_exchange: regions are whole 16-byte groups");
    BB_CHECK(x.world >= 2 && x.world <= 8, "grad_exchange: 2..8 ranks");
    const size_t per = ((hi - lo) / 4 + x.world - 1) / x.world;
    const int U = x.world <= 2 ? 8 : x.world <= 4 ? 4 : 2;
    // under the backward pass (max_blocks > 0) a bounded number of blocks keeps NVLink busy without taking SMs from the
    // GEMMs (a resident exchange block costs a GEMM CTA its slot); alone on the GPU the kernel spreads over every SM
    static const int mult = getenv("BB_XCHG_BLOCKS_PER_SM") ? atoi(getenv("BB_XCHG_BLOCKS_PER_SM")) : 2;
    const size_t cap = max_blocks > 0 ? (size_t)max_blocks : (size_t)c.sms * mult;
    const int blocks = (int)std::max<size_t>(1, std::min<size_t>((per + 256 * U - 1) / (256 * U), cap));
    static const int dbg = getenv("BB_XCHG_DBG") ? atoi(getenv("BB_XCHG_DBG")) : 0;
    switch (x.world) {
#define BB_XCHG_CASE(W) case W: grad_exchange_kernel<W><<<blocks, 256, 0, c.stream>>>(x, lo / 4, hi / 4, region, dbg); break;
        BB_XCHG_CASE(2) BB_XCHG_CASE(3) BB_XCHG_CASE(4) BB_XCHG_CASE(5) BB_XCHG_CASE(6) BB_XCHG_CASE(7) BB_XCHG_CASE(8)
#undef BB_XCHG_CASE
    }
    BB_LAUNCHED();
    c.layer = region ? "conv" : "fc";
    c.mark("grad_exchange");
}

template <int W>
__global__ void __launch_bounds__(256) grad_exchange_ll_kernel(Exchange x, size_t lo4, size_t hi4, int k) {
    __shared__ bool s_last;
    const unsigned int epoch = x.ctr[4 + 2 * k] + 1;
    if (blockIdx.x == 0 && threadIdx.x == 0) x.ctr[20 + 4 * k] = gtimer_lo();
    float* mine = const_cast<float*>(x.grads[x.rank]);
    const size_t slot = x.ll_cap / 2;   // uint4 entries per sender slot (one entry carries two floats)
    const long long t0 = clock64();
    for (size_t q = lo4 + (size_t)blockIdx.x * 256 + threadIdx.x; q < hi4; q += (size_t)gridDim.x * 256) {
        const float4 own = reinterpret_cast<const float4*>(mine)[q];
        const size_t e = q * 2;   // (entries are addressed by absolute position: concurrent exchanges use disjoint ranges)
        const uint4 a = make_uint4(__float_as_uint(own.x), epoch, __float_as_uint(own.y), epoch);
        const uint4 b = make_uint4(__float_as_uint(own.z), epoch, __float_as_uint(own.w), epoch);
#pragma unroll
        for (int r = 0; r < W; ++r)
            if (r != x.rank) {
                uint4* dst = reinterpret_cast<uint4*>(const_cast<float*>(x.grads[r]) + x.ll_off) + (size_t)x.rank * slot + e;
                dst[0] = a; dst[1] = b;
            }
        float4 t[W];
#pragma unroll
        for (int r = 0; r < W; ++r) {
            if (r == x.rank) { t[r] = own; continue; }
            const uint4* src = reinterpret_cast<const uint4*>(mine + x.ll_off) + (size_t)r * slot + e;
            uint4 u, v;
            for (;;) {
                asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(src));
                asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + 1));
                if (u.y == epoch && u.w == epoch && v.y == epoch && v.w == epoch) break;
                if (clock64() - t0 > x.timeout_cycles) {   // a peer never sent: raise the sticky failure flag, do not hang
                    *reinterpret_cast<volatile int*>(x.err) = 21;
                    break;
                }
            }
            t[r] = make_float4(__uint_as_float(u.x), __uint_as_float(u.z), __uint_as_float(v.x), __uint_as_float(v.z));
        }
#pragma unroll
        for (int s = 1; s < W; s *= 2)
#pragma unroll
            for (int r = 0; r + s < W; r += 2 * s) { t[r].x += t[r + s].x; t[r].y += t[r + s].y; t[r].z += t[r + s].z; t[r].w += t[r + s].w; }
        const float w = (float)W;
        reinterpret_cast<float4*>(mine)[q] = make_float4(t[0].x / w, t[0].y / w, t[0].z / w, t[0].w / w);
    }
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&x.ctr[5 + 2 * k], 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last && threadIdx.x == 0) { x.ctr[5 + 2 * k] = 0u; x.ctr[4 + 2 * k] = epoch; x.ctr[21 + 4 * k] = gtimer_lo(); }
}

void grad_exchange_ll(const Ctx& c, const Exchange& x, size_t lo, size_t hi, int k) {
    BB_CHECK((lo & 3) == 0 && (hi & 3) == 0 && hi <= x.ll_cap && (k == 0 || k == 1), "grad_exchange_ll: region does not fit the receive area");
    BB_CHECK(x.world >= 2 && x.world <= 8, "grad_exchange_ll: 2..8 ranks");
    // every block must be resident (a block's sends are what its peers' twin blocks wait for): at most 4 light blocks per SM
    const int blocks = (int)std::max<size_t>(1, std::min<size_t>(((hi - lo) / 4 + 255) / 256, (size_t)c.sms * 4));
    static const int dbg = getenv("BB_XCHG_DBG") ? atoi(getenv("BB_XCHG_DBG")) : 0;
    if (dbg & 2) return;   // timing experiments only: no exchange at all
    switch (x.world) {
#define BB_XCHG_CASE(W) case W: grad_exchange_ll_kernel<W><<<blocks, 256, 0, c.stream>>>(x, lo / 4, hi / 4, k); break;
        BB_XCHG_CASE(2) BB_XCHG_CASE(3) BB_XCHG_CASE(4) BB_XCHG_CASE(5) BB_XCHG_CASE(6) BB_XCHG_CASE(7) BB_XCHG_CASE(8)
#undef BB_XCHG_CASE
    }
    BB_LAUNCHED();
    c.layer = k ? "conv.mid" : "conv";
    c.mark("grad_exchange_ll");
}

void grad_reduce_scatter(const Ctx& c, const float* const* peer_grads, size_t n, int rank, int world) {
    PeerPtrs pp;
    for (int r = 0; r < 8; ++r) pp.p[r] = r < world ? peer_grads[r] : nullptr;
    size_t per = ((n >> 2) + world - 1) / world;
    int blocks = (int)std::min<size_t>((per + 255) / 256 + 1, (size_t)c.sms * 4);
    grad_reduce_scatter_kernel<<<blocks, 256, 0, c.stream>>>(pp, n, rank, world);
    BB_LAUNCHED();
    c.layer = "";
    c.mark("grad_reduce_scatter");
}

AdamScalars adam_scalars(const AdamHyper& h, uint64_t step) {
    double bc1 = 1.0 - pow(h.beta1, (double)step);
    double bc2 = 1.0 - pow(h.beta2, (double)step);
    double step_size = h.lr / bc1;
    return AdamScalars{(float)sqrt(bc2), (float)(-step_size), (float)(1.0 - h.lr * h.wd), 0.f};
}

void adam_step(const Ctx& c, float* p, const float* g, float* m, float* v, size_t n, const AdamHyper& h,
               uint64_t step, const float* const* peer_grads, int world, float* p_lo, const Exchange* wait, int wait_regions,
               const AdamScalars* dev_scalars) {
    const AdamScalars sc = adam_scalars(h, step);
    PeerPtrs pp;
    for (int r = 0; r < 8; ++r) pp.p[r] = (peer_grads && r < world) ? peer_grads[r] : nullptr;
    BB_CHECK(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
               reinterpret_cast<uintptr_t>(v)) & 15) == 0, "adam_step: buffers must be 16 B aligned");
    int blocks = (int)std::min<size_t>((n / 4 + 255) / 256 + 1, (size_t)c.sms * 8);
    adam_kernel<<<blocks, 256, 0, c.stream>>>(p, g, m, v, n, (float)h.beta1, (float)h.beta2, (float)(1.0 - h.beta1),
                                              (float)(1.0 - h.beta2), (float)h.eps, sc.bc2_sqrt, sc.neg_step, (float)h.wd, sc.decay,
                                              h.adamw ? 1 : 0, pp, peer_grads ? world : 1, p_lo, wait ? *wait : Exchange{},
                                              wait ? wait_regions : 0, dev_scalars);
    BB_LAUNCHED();
    c.layer = "";
    c.mark("adam");
}

__global__ void track_kernel(float* __restrict__ dest, const float* __restrict__ src, size_t n, float tau,
                             float one_m_tau, float* __restrict__ dest_lo) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float d = __fadd_rn(__fmul_rn(tau, src[i]), __fmul_rn(one_m_tau, dest[i]));
        dest[i] = d;
        if (dest_lo) dest_lo[i] = lo_of(d);
    }
}
void track(const Ctx& c, float* dest, const float* src, size_t n, double tau, float* dest_lo) {
    int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)c.sms * 8);
    track_kernel<<<blocks, 256, 0, c.stream>>>(dest, src, n, (float)tau, (float)(1.0 - tau), dest_lo);
    BB_LAUNCHED();
    c.layer = "";
    c.mark("track");
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__global__ void fill_uniform_kernel(float* p, size_t n, float bound, unsigned long long seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float u = (float)(uint32_t)(mix64(seed + i) >> 40) * (1.0f / 16777216.0f);
        p[i] = (2.f * u - 1.f) * bound;
    }
}
__global__ void fill_const_kernel(float* p, size_t n, float v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
void fill_uniform(const Ctx& c, float* p, size_t n, float bound, uint64_t seed) {
    if (!n) return;
    fill_uniform_kernel<<<(int)std::min<size_t>((n + 255) / 256, 1024), 256, 0, c.stream>>>(p, n, bound, seed);
    BB_LAUNCHED();
}
void fill_const(const Ctx& c, float* p, size_t n, float v) {
    if (!n) return;
    fill_const_kernel<<<(int)std::min<size_t>((n + 255) / 256, 1024), 256, 0, c.stream>>>(p, n, v);
    BB_LAUNCHED();
}

// ------------------------------------------------------------------------------- Net

void NetWorkspace::release() {
    for (auto p : act) cudaFree(p);
    for (auto p : dact) cudaFree(p);
    for (auto p : rowbase) cudaFree(p);
    for (auto p : dg_rowbase) cudaFree(p);
    for (auto p : dg_crow) cudaFree(p);
    for (auto p : dypad) cudaFree(p);
    for (auto p : dg_wt) cudaFree(p);
    cudaFree(col);
    act.clear(); dact.clear(); rowbase.clear(); dg_rowbase.clear(); dg_crow.clear(); dypad.clear(); dg_wt.clear(); col = nullptr;
    plane.clear(); dypad_plane.clear(); wt_plane.clear();
}

static void add_param(Net& n, const std::string& name, std::vector<int64_t> shape, int perm, int pc, int ph, int pw,
                      int fan_in) {
    ParamInfo pi;
    pi.name = name; pi.shape = shape; pi.offset = n.n_params; pi.perm = perm; pi.pc = pc; pi.ph = ph; pi.pw = pw;
    pi.fan_in = fan_in;
    pi.numel = 1;
    for (auto d : shape) pi.numel *= (size_t)d;
    n.n_params += (pi.numel + 3) / 4 * 4;  // keep every tensor 16 B aligned
    n.params.push_back(pi);
}

static void add_linear(Net& n, const std::string& name, int in, int out, bool relu, int perm = 0, int pc = 0, int ph = 0,
                       int pw = 0) {
    Layer l{};
    l.type = 0; l.in_dim = in; l.out_dim = out; l.relu = relu; l.out_elems_per_sample = out;
    l.w_off = n.n_params;
    add_param(n, name + ".weight", {out, in}, perm, pc, ph, pw, in);
    l.b_off = n.n_params;
    add_param(n, name + ".bias", {out}, 0, 0, 0, 0, in);
    n.layers.push_back(l);
}

static void add_conv(Net& n, const std::string& name, int C, int H, int W, int OC, int k, int s, bool u8) {
    Layer l{};
    l.type = 1; l.relu = true;
    ConvGeom& g = l.geom;
    g.B = 0; g.C = C; g.H = H; g.W = W; g.OC = OC; g.KH = k; g.KW = k; g.S = s;
    g.OH = (H - k) / s + 1; g.OW = (W - k) / s + 1; g.u8_chw = u8; g.rowbase = nullptr; g.koff = nullptr;
    l.out_elems_per_sample = (size_t)g.OH * g.OW * OC;
    l.w_off = n.n_params;
    add_param(n, name + ".weight", {OC, C, k, k}, u8 ? 0 : 1, C, k, k, C * k * k);
    l.b_off = n.n_params;
    add_param(n, name + ".bias", {OC}, 0, 0, 0, 0, C * k * k);
    n.layers.push_back(l);
}

void Net::add_linear_layer(const std::string& name, int in, int out, bool relu) {
    if (layers.empty()) in_elems = in;
    add_linear(*this, name, in, out, relu);
    out_dim = out;
}

void Net::add_twin_heads(const std::string& name1, const std::string& name2, int in, int out) {
    if (layers.empty()) in_elems = in;
    Layer l{};
    l.type = 0; l.in_dim = in; l.out_dim = 2 * out; l.relu = false; l.out_elems_per_sample = 2 * out;
    auto push = [&](const std::string& nm, std::vector<int64_t> shape, size_t off) {
        ParamInfo pi;
        pi.name = nm; pi.shape = shape; pi.offset = off; pi.perm = 0; pi.pc = pi.ph = pi.pw = 0; pi.fan_in = in;
        pi.numel = 1;
        for (auto d : shape) pi.numel *= (size_t)d;
        params.push_back(pi);
    };
    l.w_off = n_params;
    push(name1 + ".weight", {out, in}, n_params);
    push(name2 + ".weight", {out, in}, n_params + (size_t)out * in);
    n_params += ((size_t)2 * out * in + 3) / 4 * 4;
    l.b_off = n_params;
    push(name1 + ".bias", {out}, n_params);
    push(name2 + ".bias", {out}, n_params + out);
    n_params += ((size_t)2 * out + 3) / 4 * 4;
    layers.push_back(l);
    out_dim = 2 * out;
}

void Net::build(const bb_net_cfg& cfg, const std::string& prefix) {
    layers.clear(); params.clear(); n_params = 0;
    if (cfg.kind == BB_NET_ATARI_CNN) {  // cnn/base.rs:23-48
        u8_input = true;
        in_elems = cfg.n_stack * 84 * 84;
        add_conv(*this, prefix + "c1", cfg.n_stack, 84, 84, 32, 8, 4, true);
        add_conv(*this, prefix + "c2", 32, 20, 20, 64, 4, 2, false);
        add_conv(*this, prefix + "c3", 64, 9, 9, 64, 3, 1, false);
        if (!cfg.skip_linear) {
            add_linear(*this, prefix + "l1", 3136, 512, true, 2, 64, 7, 7);
            add_linear(*this, prefix + "l2", 512, cfg.out_dim, false);
            out_dim = cfg.out_dim;
        } else {
            out_dim = 3136;
        }
    } else if (cfg.kind == BB_NET_MLP) {  // mlp/base.rs:13-41, var names mlp.ln{i}
        u8_input = false;
        in_elems = cfg.in_dim;
        int in = cfg.in_dim;
        BB_CHECK(cfg.n_units >= 0 && cfg.n_units <= 8, "MlpConfig.units: at most 8 hidden layers");
        for (int i = 0; i < cfg.n_units; ++i) {
            add_linear(*this, prefix + "mlp.ln" + std::to_string(i), in, cfg.units[i], true);
            in = cfg.units[i];
        }
        add_linear(*this, prefix + "mlp.ln" + std::to_string(cfg.n_units), in, cfg.out_dim, cfg.activation_out != 0);
        out_dim = cfg.out_dim;
    } else {
        throw Error("unknown network kind");
    }
}

void Net::init_tables(int device) {
    (void)device;
    free_tables();
    for (auto& l : layers) {
        if (l.type != 1) continue;
        ConvGeom& g = l.geom;
        std::vector<int> h(g.K());
        if (g.u8_chw) {
            for (int c = 0; c < g.C; ++c)
                for (int kh = 0; kh < g.KH; ++kh)
                    for (int kw = 0; kw < g.KW; ++kw) h[(c * g.KH + kh) * g.KW + kw] = c * g.H * g.W + kh * g.W + kw;
        } else {
            for (int kh = 0; kh < g.KH; ++kh)
                for (int kw = 0; kw < g.KW; ++kw)
                    for (int c = 0; c < g.C; ++c) h[(kh * g.KW + kw) * g.C + c] = (kh * g.W + kw) * g.C + c;
        }
        int* d = dev_alloc<int>(h.size());
        BB_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice));
        koff.push_back(d);
        g.koff = d;
        if (&l != &layers[0] && g.dgrad_gather_ok()) {
            // gather-form data gradient (see ConvGeom): k = (jh, jw, oc), n = (ph, pw, c)
            const int S = g.S, Jh = g.KH / S, Jw = g.KW / S, OWp = g.dg_wp();
            std::vector<int> ko((size_t)Jh * Jw * g.OC), br(ko.size()), bn((size_t)S * S * g.C), cc(bn.size());
            for (int jh = 0; jh < Jh; ++jh)
                for (int jw = 0; jw < Jw; ++jw)
                    for (int oc = 0; oc < g.OC; ++oc) {
                        size_t k = ((size_t)jh * Jw + jw) * g.OC + oc;
                        ko[k] = ((Jh - 1 - jh) * OWp + (Jw - 1 - jw)) * g.OC + oc;   // oh + Jh-1 = h/S + (Jh-1-jh)
                        br[k] = ((oc * g.KH + S * jh) * g.KW + S * jw) * g.C;       // W[oc][S jh + ph][S jw + pw][c]
                    }
            for (int ph = 0; ph < S; ++ph)
                for (int pw = 0; pw < S; ++pw)
                    for (int c = 0; c < g.C; ++c) {
                        size_t n = ((size_t)ph * S + pw) * g.C + c;
                        bn[n] = (ph * g.KW + pw) * g.C + c;
                        cc[n] = (ph * g.W + pw) * g.C + c;
                    }
            auto up = [&](const std::vector<int>& v) {
                int* dp = dev_alloc<int>(v.size());
                BB_CUDA(cudaMemcpy(dp, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice));
                dg_tables.push_back(dp);
                return dp;
            };
            g.dg_koff = up(ko); g.dg_brow = up(br); g.dg_bnoff = up(bn); g.dg_ccol = up(cc);
        }
    }
}

void Net::free_tables() {
    for (auto p : koff) cudaFree(p);
    for (auto p : dg_tables) cudaFree(p);
    koff.clear();
    dg_tables.clear();
}

void Net::alloc_workspace(NetWorkspace& w, int max_batch, bool with_grad) const {
    w.release();
    w.max_batch = max_batch;
    w.with_grad = with_grad;
    size_t col = 0;
    for (size_t i = 0; i < layers.size(); ++i) {
        const Layer& l = layers[i];
        // every buffer is followed by its lo plane (x - tf32_trunc(x)), the second operand plane of the TMA-fed GEMMs
        size_t n = ((size_t)max_batch * l.out_elems_per_sample + 3) / 4 * 4;
        w.plane.push_back((long)n);
        w.act.push_back(dev_alloc<float>(2 * n));
        w.dact.push_back(with_grad ? dev_alloc<float>(2 * n) : nullptr);
        if (l.type == 1) {
            const ConvGeom& g = l.geom;
            size_t M = (size_t)max_batch * g.OH * g.OW;
            BB_CHECK(M * 4 < (1ull << 31) && (size_t)max_batch * g.C * g.H * g.W < (1ull << 31), "batch too large for int32 gather tables");
            std::vector<int> h(M);
            for (int b = 0; b < max_batch; ++b)
                for (int oh = 0; oh < g.OH; ++oh)
                    for (int ow = 0; ow < g.OW; ++ow) {
                        size_t m = ((size_t)b * g.OH + oh) * g.OW + ow;
                        h[m] = g.u8_chw ? b * g.C * g.H * g.W + oh * g.S * g.W + ow * g.S
                                        : ((b * g.H + oh * g.S) * g.W + ow * g.S) * g.C;
                    }
            int* d = dev_alloc<int>(M);
            BB_CUDA(cudaMemcpy(d, h.data(), M * sizeof(int), cudaMemcpyHostToDevice));
            w.rowbase.push_back(d);
            int *dgr = nullptr, *dgc = nullptr;
            float *pad = nullptr, *wt = nullptr;
            long pad_plane = 0, wt_plane_ = 0;
            if (with_grad && i > 0 && g.dg_koff) {  // gather-form data gradient: per-batch row tables + padded dY
                const int Hq = g.H / g.S, Wq = g.W / g.S, OHp = g.dg_hp(), OWp = g.dg_wp();
                size_t Mq = (size_t)max_batch * Hq * Wq;
                BB_CHECK((size_t)max_batch * OHp * OWp * g.OC < (1ull << 31), "batch too large for int32 gather tables");
                std::vector<int> rb(Mq), cr(Mq);
                for (int b = 0; b < max_batch; ++b)
                    for (int hq = 0; hq < Hq; ++hq)
                        for (int wq = 0; wq < Wq; ++wq) {
                            size_t m = ((size_t)b * Hq + hq) * Wq + wq;
                            rb[m] = ((b * OHp + hq) * OWp + wq) * g.OC;
                            cr[m] = ((b * g.H + g.S * hq) * g.W + g.S * wq) * g.C;
                        }
                dgr = dev_alloc<int>(Mq); dgc = dev_alloc<int>(Mq);
                BB_CUDA(cudaMemcpy(dgr, rb.data(), Mq * sizeof(int), cudaMemcpyHostToDevice));
                BB_CUDA(cudaMemcpy(dgc, cr.data(), Mq * sizeof(int), cudaMemcpyHostToDevice));
                size_t pn = (size_t)max_batch * OHp * OWp * g.OC;   // OC % 4 == 0
                pad = dev_alloc<float>(2 * pn);
                BB_CUDA(cudaMemset(pad, 0, 2 * pn * sizeof(float)));
                wt = dev_alloc<float>(2 * (size_t)g.K() * g.OC);
                pad_plane = (long)pn; wt_plane_ = (long)g.K() * g.OC;
            }
            // the col2im path (CUDA-core fallback, BB_TC=0 / BB_DGRAD_GATHER=0) keeps its column buffer
            if (with_grad && i > 0) col = std::max(col, M * (size_t)g.K());
            w.dg_rowbase.push_back(dgr); w.dg_crow.push_back(dgc); w.dypad.push_back(pad); w.dg_wt.push_back(wt);
            w.dypad_plane.push_back(pad_plane); w.wt_plane.push_back(wt_plane_);
        } else {
            w.rowbase.push_back(nullptr);
            w.dg_rowbase.push_back(nullptr); w.dg_crow.push_back(nullptr); w.dypad.push_back(nullptr); w.dg_wt.push_back(nullptr);
            w.dypad_plane.push_back(0); w.wt_plane.push_back(0);
        }
    }
    w.col_floats = col;
    w.col = col ? dev_alloc<float>(col) : nullptr;
}

void Net::init_params(const Ctx& c, float* p, uint64_t seed) const {
    // tch defaults: weights Kaiming-uniform(fan_in, ReLU gain) = U(+-sqrt(6/fan_in)); linear bias
    // U(+-1/sqrt(fan_in)); conv bias 0.  (The reference draws from libtorch's global generator, so
    // initial weights are never reproducible across implementations: parity tests load weights.)
    fill_const(c, p, n_params, 0.f);
    for (size_t i = 0; i < params.size(); ++i) {
        const ParamInfo& pi = params[i];
        bool is_w = pi.shape.size() > 1;
        bool conv = pi.shape.size() == 4;
        if (is_w) fill_uniform(c, p + pi.offset, pi.numel, sqrtf(6.0f / (float)pi.fan_in), seed * 1315423911ull + i * 0x51ed27ull);
        else if (!conv && !(i > 0 && params[i - 1].shape.size() == 4))
            fill_uniform(c, p + pi.offset, pi.numel, 1.0f / sqrtf((float)pi.fan_in), seed * 1315423911ull + i * 0x51ed27ull);
    }
}

std::string Net::layer_name(size_t i) const {
    // parameter names are "<layer>.weight": strip the suffix of the layer's weight tensor
    for (auto& pi : params)
        if (pi.offset == layers[i].w_off) return pi.name.substr(0, pi.name.rfind('.'));
    return "layer" + std::to_string(i);
}

bool Net::direct_input_ok(int B) const {
    if (layers.empty() || layers[0].type != 1) return false;
    ConvGeom g = layers[0].geom;
    g.B = B;
    return conv1_direct_ok(g);
}

const float* Net::forward(const Ctx& c, const float* p, const void* input, long ld_in, int B, NetWorkspace& w, long p_plane,
                          const unsigned long long* in_ix) const {
    BB_CHECK(B <= w.max_batch, "batch larger than the workspace");
    BB_CHECK(!in_ix || direct_input_ok(B), "indexed input needs the dedicated first-layer kernels");
    const void* x = input;
    long ldx = ld_in;
    for (size_t i = 0; i < layers.size(); ++i) {
        const Layer& l = layers[i];
        c.layer = layer_name(i) + ".fwd";
        // lo planes: the network input has none; every layer output gets one when the parameters carry one
        const long x_plane = (p_plane && i) ? w.plane[i - 1] : 0, y_plane = p_plane ? w.plane[i] : 0;
        if (l.type == 1) {
            ConvGeom g = l.geom;
            g.B = B; g.rowbase = w.rowbase[i];
            g.x_plane = x_plane; g.w_plane = p_plane; g.y_plane = y_plane;
            g.in_ix = i == 0 ? in_ix : nullptr;
            conv_fwd(c, g, x, p + l.w_off, p + l.b_off, w.act[i], l.relu);
        } else {
            linear_fwd(c, (const float*)x, ldx, p + l.w_off, p + l.b_off, w.act[i], B, l.out_dim, l.in_dim, l.relu, x_plane,
                       p_plane, y_plane);
        }
        x = w.act[i];
        ldx = (long)l.out_elems_per_sample;
        if (p_plane && env_int("BB_DEBUG_CHECK_LO", 0)) {  // debugging aid: is the lo plane every producer must write really there?
            const size_t n = (size_t)B * l.out_elems_per_sample;
            std::vector<float> hx(n), hl(n);
            BB_CUDA(cudaStreamSynchronize(c.stream));
            BB_CUDA(cudaMemcpy(hx.data(), w.act[i], n * 4, cudaMemcpyDeviceToHost));
            BB_CUDA(cudaMemcpy(hl.data(), w.act[i] + w.plane[i], n * 4, cudaMemcpyDeviceToHost));
            size_t bad = 0, first = 0;
            for (size_t j = 0; j < n; ++j) {
                uint32_t u; memcpy(&u, &hx[j], 4); u &= 0xffffe000u;
                float hi; memcpy(&hi, &u, 4);
                if (hl[j] != hx[j] - hi) { if (!bad) first = j; ++bad; }
            }
            double sum = 0, sq = 0;
            for (size_t j = 0; j < n; ++j) { sum += hx[j]; sq += (double)hx[j] * hx[j]; }
            fprintf(stderr, "[check_lo] %s (stream %p): %zu / %zu lo values wrong (first at %zu: x %.9g lo %.9g)  sum %.12g sumsq %.12g  x[0..2] %.9g %.9g %.9g\n",
                    c.layer.c_str(), (void*)c.stream, bad, n, first, bad ? hx[first] : 0.f, bad ? hl[first] : 0.f, sum, sq, hx[0], hx[1], hx[2]);
        }
    }
    return w.act.back();
}

// ---- small-batch forward (Policy::sample): one cooperative kernel for the whole net ------------------------------
struct SmallLayer {
    int type, M, N, K, relu, u8;   // type 0 linear, 1 conv (gather); output [M][N]
    const int* rowbase;            // conv: [M]
    const int* koff;               // conv: [K]
    const void* in;
    float* out;
    long ld_in;                    // linear: row stride of the input
    size_t w_off, b_off;
};
struct SmallNet {
    SmallLayer l[8];
    int n_layers;
    const float* p;
};

// A warp per output element; layers with fewer outputs than warps (the 512-wide fully connected layer: 98 dependent
// 128-byte steps per warp, 15 us of the 35 us forward) split the contraction over 2 / 4 / 8 warps of a CTA and fold the
// partial sums through shared memory.
__global__ void __launch_bounds__(256) small_forward_kernel(SmallNet P) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    __shared__ float s_part[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long nw = (long)gridDim.x * 8;
    for (int li = 0; li < P.n_layers; ++li) {
        const SmallLayer& L = P.l[li];
        const float* W = P.p + L.w_off;
        const float* bias = P.p + L.b_off;
        const long total = (long)L.M * L.N;
        int wpo = 1;                                   // warps per output
        while (wpo < 8 && total * (wpo * 2) <= nw) wpo *= 2;
        const int opc = 8 / wpo;                       // outputs per CTA and trip
        const int part = warp % wpo;
        const int kspan = ((L.K + wpo * 32 - 1) / (wpo * 32)) * 32;   // contraction elements per warp (multiple of 32)
        for (long og = blockIdx.x; og * opc < total; og += gridDim.x) {
            const long o = og * opc + warp / wpo;
            float acc = 0.f;
            if (o < total) {
                const int m = (int)(o / L.N), n = (int)(o % L.N);
                const float* w = W + (size_t)n * L.K;
                const int k0 = part * kspan, k1 = min(L.K, k0 + kspan);
                if (L.type == 1) {
                    const long base = L.rowbase[m];
                    if (L.u8) {
                        const uint8_t* x = reinterpret_cast<const uint8_t*>(L.in) + base;
                        const float s = 1.0f / 255.0f;
                        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;   // four independent gather chains in flight
                        int k = k0 + lane;
                        for (; k + 96 < k1; k += 128) {
                            const int o0 = L.koff[k], o1 = L.koff[k + 32], o2 = L.koff[k + 64], o3 = L.koff[k + 96];
                            a0 = fmaf((float)x[o0] * s, w[k], a0); a1 = fmaf((float)x[o1] * s, w[k + 32], a1);
                            a2 = fmaf((float)x[o2] * s, w[k + 64], a2); a3 = fmaf((float)x[o3] * s, w[k + 96], a3);
                        }
                        for (; k < k1; k += 32) a0 = fmaf((float)x[L.koff[k]] * s, w[k], a0);
                        acc = (a0 + a1) + (a2 + a3);
                    } else {
                        const float* x = reinterpret_cast<const float*>(L.in) + base;
                        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                        int k = k0 + lane;
                        for (; k + 96 < k1; k += 128) {
                            const int o0 = L.koff[k], o1 = L.koff[k + 32], o2 = L.koff[k + 64], o3 = L.koff[k + 96];
                            a0 = fmaf(x[o0], w[k], a0); a1 = fmaf(x[o1], w[k + 32], a1);
                            a2 = fmaf(x[o2], w[k + 64], a2); a3 = fmaf(x[o3], w[k + 96], a3);
                        }
                        for (; k < k1; k += 32) a0 = fmaf(x[L.koff[k]], w[k], a0);
                        acc = (a0 + a1) + (a2 + a3);
                    }
                } else {
                    const float* x = reinterpret_cast<const float*>(L.in) + (size_t)m * L.ld_in;
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;   // four independent chains: loads of 4 steps in flight
                    int k = k0 + lane;
                    for (; k + 96 < k1; k += 128) {
                        a0 = fmaf(x[k], w[k], a0); a1 = fmaf(x[k + 32], w[k + 32], a1);
                        a2 = fmaf(x[k + 64], w[k + 64], a2); a3 = fmaf(x[k + 96], w[k + 96], a3);
                    }
                    for (; k < k1; k += 32) a0 = fmaf(x[k], w[k], a0);
                    acc = (a0 + a1) + (a2 + a3);
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
            }
            if (wpo > 1) {                              // (uniform per layer: every warp of the CTA takes the barriers)
                if (lane == 0) s_part[warp] = acc;
                __syncthreads();
                if (part == 0 && lane == 0) {
                    acc = 0.f;
                    for (int j = 0; j < wpo; ++j) acc += s_part[warp + j];
                }
                __syncthreads();
            }
            if (o < total && part == 0 && lane == 0) {
                const int n = (int)(o % L.N);
                acc += bias[n];
                if (L.relu) acc = fmaxf(acc, 0.f);
                L.out[o] = acc;
            }
        }
        if (li + 1 < P.n_layers) grid.sync();
    }
}

const float* Net::forward_small(const Ctx& c, const float* p, const void* input, long ld_in, int B, NetWorkspace& w) const {
    if (B > 8 || layers.size() > 8 || layers.empty() || B > w.max_batch || !env_int("BB_POLICY_FUSED", 1)) return nullptr;
    SmallNet P;
    memset(&P, 0, sizeof(P));
    P.n_layers = (int)layers.size();
    P.p = p;
    const void* x = input;
    long ldx = ld_in;
    for (size_t i = 0; i < layers.size(); ++i) {
        const Layer& l = layers[i];
        SmallLayer& s = P.l[i];
        s.relu = l.relu ? 1 : 0; s.in = x; s.out = w.act[i]; s.w_off = l.w_off; s.b_off = l.b_off; s.ld_in = ldx;
        if (l.type == 1) {
            const ConvGeom& g = l.geom;
            s.type = 1; s.M = B * g.OH * g.OW; s.N = g.OC; s.K = g.K(); s.u8 = g.u8_chw ? 1 : 0;
            s.rowbase = w.rowbase[i]; s.koff = g.koff;
        } else {
            s.type = 0; s.M = B; s.N = l.out_dim; s.K = l.in_dim;
        }
        x = w.act[i];
        ldx = (long)l.out_elems_per_sample;
    }
    static int ctas_per_sm = 0;
    if (!ctas_per_sm) {
        BB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, small_forward_kernel, 256, 0));
        ctas_per_sm = std::max(1, std::min(ctas_per_sm, 4));
    }
    void* args[] = {&P};
    BB_CUDA(cudaLaunchCooperativeKernel((void*)small_forward_kernel, dim3(c.sms * ctas_per_sm), dim3(256), args, 0, c.stream));
    BB_LAUNCHED();
    c.phase = "policy"; c.layer = "forward";
    c.mark("small_forward");
    return w.act.back();
}

__global__ void relu_mask_kernel(float* __restrict__ d, const float* __restrict__ y, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        d[i] = y[i] > 0.f ? d[i] : 0.f;
}

void Net::backward(const Ctx& c, const float* p, float* g, const void* input, long ld_in, int B, NetWorkspace& w,
                   float* d_input, long ld_din, long p_plane, const unsigned long long* in_ix,
                   const std::function<void(int)>* hook) const {
    BB_CHECK(w.with_grad, "workspace was allocated without gradient buffers");
    int L = (int)layers.size();
    // d(output) arrives in w.dact[L-1] as the gradient wrt the post-activation output
    if (layers[L - 1].relu) {
        size_t n = (size_t)B * layers[L - 1].out_elems_per_sample;
        relu_mask_kernel<<<(int)std::min<size_t>((n + 255) / 256, (size_t)c.sms * 8), 256, 0, c.stream>>>(w.dact[L - 1], w.act[L - 1], n);
        BB_LAUNCHED();
    }
    // d(output) comes from a loss kernel without a lo plane: a wide last layer (tensor-core contraction) needs one
    bool dy_lo = false;   // does dact[i] carry a valid lo plane?
    if (p_plane && layers[L - 1].out_elems_per_sample > 16) {
        make_lo(c, w.dact[L - 1], w.dact[L - 1] + w.plane[L - 1], (size_t)B * layers[L - 1].out_elems_per_sample);
        dy_lo = true;
    }
    const bool conc = c.concurrent() && g;
    int n_side = 0;
    auto conv_geom = [&](int i) {
        ConvGeom cg = layers[i].geom;
        cg.B = B; cg.rowbase = w.rowbase[i];
        cg.dg_rowbase = w.dg_rowbase[i]; cg.dg_crow = w.dg_crow[i]; cg.dypad = w.dypad[i]; cg.dg_wt = w.dg_wt[i];
        return cg;
    };
    // The re-laid weights of the gather-form data gradients depend on the parameters only, so they could run beside the start
    // of the pass -- measured SLOWER (DQN step 283.6 -> 286.8 us: the extra fork / join costs more than the 2 x 3 us launches
    // it takes off the critical path), so it is opt-in (BB_DGRAD_WT_SIDE=1).
    bool wt_side = false, wt_joined = false;
    if (conc && env_int("BB_DGRAD_WT_SIDE", 0)) {
        for (int i = L - 1; i >= 1; --i) {
            if (layers[i].type != 1) continue;
            ConvGeom cg = conv_geom(i);
            cg.w_plane = p_plane; cg.wt_plane = p_plane ? w.wt_plane[i] : 0;
            if (!dgrad_gather_path(cg)) continue;
            if (!wt_side) { c.fork_to(*c.side[0]); wt_side = true; }
            c.side[0]->layer = layer_name(i) + ".dgrad";
            conv_dgrad_prepare_weights(*c.side[0], cg, p + layers[i].w_off);
        }
    }
    for (int i = L - 1; i >= 0; --i) {
        const Layer& l = layers[i];
        const void* x = i ? (const void*)w.act[i - 1] : input;
        long ldx = i ? (long)layers[i - 1].out_elems_per_sample : ld_in;
        float* dx = i ? w.dact[i - 1] : d_input;
        long lddx = i ? (long)layers[i - 1].out_elems_per_sample : ld_din;
        // the mask folds the previous layer's ReLU backward into this layer's data-grad epilogue
        const float* mask = (i && layers[i - 1].relu) ? w.act[i - 1] : nullptr;
        // weight gradient: off the critical path (it only feeds the optimizer) unless nothing follows it
        const Ctx* wc = &c;
        const Ctx* bc = nullptr;   // where the bias gradient (column sums) runs; null = with the weight gradient
        if (conc && dx) {
            wc = c.side[n_side++ & 1];
            c.fork_to(*wc);  // dact[i] is complete on c.stream here
        } else if (conc && l.type == 1) {
            // last layer of the chain: the weight gradient stays on c.stream, its bias gradient goes aside
            bc = c.side[n_side++ & 1];
            c.fork_to(*bc);
        }
        // lo planes of this layer's operands: X = act[i-1] (forward wrote it), dY = dact[i], dX = dact[i-1]
        const long x_plane = (p_plane && i) ? w.plane[i - 1] : 0, dy_plane = dy_lo ? w.plane[i] : 0;
        const long dx_plane = (p_plane && i) ? w.plane[i - 1] : 0;
        if (l.type == 1) {
            ConvGeom cg = conv_geom(i);
            cg.wt_ready = (wt_side && dx) ? 1 : 0;
            if (cg.wt_ready && !wt_joined) { c.join_from(*c.side[0]); wt_joined = true; }
            cg.x_plane = x_plane; cg.y_plane = dy_plane; cg.w_plane = p_plane; cg.dx_plane = dx_plane;
            cg.dypad_plane = p_plane ? w.dypad_plane[i] : 0; cg.wt_plane = p_plane ? w.wt_plane[i] : 0;
            cg.in_ix = i == 0 ? in_ix : nullptr;
            if (g) {
                c.layer = layer_name(i) + ".wgrad";
                if (bc) colsum(*bc, w.dact[i], g + l.b_off, cg.M(), cg.OC);
                conv_bwd_weight(*wc, cg, w.dact[i], x, g + l.w_off, bc ? nullptr : g + l.b_off);
            }
            c.layer = layer_name(i) + ".dgrad";
            if (dx) conv_bwd_data(c, cg, w.dact[i], p + l.w_off, w.col, dx, mask);
        } else {
            if (g) {
                c.layer = layer_name(i) + ".wgrad";
                linear_bwd_weight(*wc, w.dact[i], (const float*)x, ldx, g + l.w_off, g + l.b_off, B, l.out_dim, l.in_dim, dy_plane,
                                  x_plane);
            }
            c.layer = layer_name(i) + ".dgrad";
            if (dx) linear_bwd_data(c, w.dact[i], p + l.w_off, dx, lddx, B, l.out_dim, l.in_dim, mask, dy_plane, p_plane, dx_plane);
        }
        dy_lo = p_plane != 0;  // every data-gradient epilogue above wrote the lo plane of dact[i-1]
        // (after the data gradient is enqueued too: an exchange that starts while it runs takes SMs from the critical path)
        if (hook && g) (*hook)(i);
    }
    if (n_side > 0) c.join_from(*c.side[0]);
    if (n_side > 1) c.join_from(*c.side[1]);
}

// ------------------------------------------------------------------------------- layout conversion

void param_to_internal(const ParamInfo& pi, const float* ref, float* internal) {
    if (pi.perm == 1) {  // OIHW -> O(HWI)
        int O = (int)pi.shape[0], I = pi.pc, H = pi.ph, W = pi.pw;
        for (int o = 0; o < O; ++o)
            for (int i = 0; i < I; ++i)
                for (int h = 0; h < H; ++h)
                    for (int w = 0; w < W; ++w)
                        internal[((size_t)(o * H + h) * W + w) * I + i] = ref[((size_t)(o * I + i) * H + h) * W + w];
    } else if (pi.perm == 2) {  // [out][C*H*W] -> [out][H*W*C]
        int O = (int)pi.shape[0], C = pi.pc, H = pi.ph, W = pi.pw;
        size_t in = (size_t)C * H * W;
        for (int o = 0; o < O; ++o)
            for (int ch = 0; ch < C; ++ch)
                for (int h = 0; h < H; ++h)
                    for (int w = 0; w < W; ++w)
                        internal[o * in + ((size_t)h * W + w) * C + ch] = ref[o * in + ((size_t)ch * H + h) * W + w];
    } else if (pi.perm == 3 || pi.perm == 4) {  // rows (or elements) of a [C*H*W][cols] tensor -> [H*W*C][cols]
        int C = pi.pc, H = pi.ph, W = pi.pw;
        size_t cols = pi.perm == 3 ? (size_t)pi.shape[1] : 1;
        for (int ch = 0; ch < C; ++ch)
            for (int h = 0; h < H; ++h)
                for (int w = 0; w < W; ++w)
                    std::copy(ref + (((size_t)ch * H + h) * W + w) * cols, ref + (((size_t)ch * H + h) * W + w + 1) * cols,
                              internal + (((size_t)h * W + w) * C + ch) * cols);
    } else {
        std::copy(ref, ref + pi.numel, internal);
    }
}

void param_to_reference(const ParamInfo& pi, const float* internal, float* ref) {
    if (pi.perm == 1) {
        int O = (int)pi.shape[0], I = pi.pc, H = pi.ph, W = pi.pw;
        for (int o = 0; o < O; ++o)
            for (int i = 0; i < I; ++i)
                for (int h = 0; h < H; ++h)
                    for (int w = 0; w < W; ++w)
                        ref[((size_t)(o * I + i) * H + h) * W + w] = internal[((size_t)(o * H + h) * W + w) * I + i];
    } else if (pi.perm == 2) {
        int O = (int)pi.shape[0], C = pi.pc, H = pi.ph, W = pi.pw;
        size_t in = (size_t)C * H * W;
        for (int o = 0; o < O; ++o)
            for (int ch = 0; ch < C; ++ch)
                for (int h = 0; h < H; ++h)
                    for (int w = 0; w < W; ++w)
                        ref[o * in + ((size_t)ch * H + h) * W + w] = internal[o * in + ((size_t)h * W + w) * C + ch];
    } else if (pi.perm == 3 || pi.perm == 4) {
        int C = pi.pc, H = pi.ph, W = pi.pw;
        size_t cols = pi.perm == 3 ? (size_t)pi.shape[1] : 1;
        for (int ch = 0; ch < C; ++ch)
            for (int h = 0; h < H; ++h)
                for (int w = 0; w < W; ++w)
                    std::copy(internal + (((size_t)h * W + w) * C + ch) * cols, internal + (((size_t)h * W + w) * C + ch + 1) * cols,
                              ref + (((size_t)ch * H + h) * W + w) * cols);
    } else {
        std::copy(internal, internal + pi.numel, ref);
    }
}

}  // namespace bb

// Test hook: ONE convolution layer (NHWC float input, [OC][KH][KW][C] weights) through the layer primitives, with
// (use_tma = 1) or without the operands' lo planes, i.e. through the TMA-fed im2col kernels or the SIMT-producer ones.
//   mode 0: Y[B][OH][OW][OC] = conv(X, W) + bias (no ReLU)      mode 1: dW[OC][KH][KW][C] from (dY, X)
//   mode 2: dX[B][H][W][C] from (dY, W)
// The layer is built as the second layer of a two-layer Net so that the production table / workspace code is what runs.
extern "C" int32_t bb_test_conv(int32_t device, int32_t mode, int32_t use_tma, int32_t B, int32_t C, int32_t H, int32_t W,
                                int32_t OC, int32_t k, int32_t s, const float* X, const float* Wt, const float* bias,
                                const float* dY, float* out) {
    BB_API_BEGIN
    using namespace bb;
    DeviceGuard dg(device);
    Ctx c;
    c.device = device; c.sms = num_sms(device); c.stream = device_stream(device);
    c.alloc_scratch(8u << 20);
    Net net;
    add_conv(net, "a", C, H, W, C, 1, 1, false);   // placeholder producing the [B][H][W][C] input; never run
    add_conv(net, "b", C, H, W, OC, k, s, false);
    net.init_tables(device);
    NetWorkspace w;
    net.alloc_workspace(w, B, true);
    const Layer& l = net.layers[1];
    ConvGeom g = l.geom;
    g.B = B; g.rowbase = w.rowbase[1];
    g.dg_rowbase = w.dg_rowbase[1]; g.dg_crow = w.dg_crow[1]; g.dypad = w.dypad[1]; g.dg_wt = w.dg_wt[1];
    const size_t nx = (size_t)B * H * W * C, ny = (size_t)B * g.OH * g.OW * OC, nw = (size_t)OC * k * k * C;
    const size_t nwp = (nw + 3) / 4 * 4;
    float* dW = dev_alloc<float>(2 * nwp);
    float* dG = dev_alloc_zero<float>(nwp, c.stream);
    float* dB = dev_alloc_zero<float>(OC, c.stream);
    BB_CUDA(cudaMemcpyAsync(w.act[0], X, nx * 4, cudaMemcpyHostToDevice, c.stream));
    BB_CUDA(cudaMemcpyAsync(dW, Wt, nw * 4, cudaMemcpyHostToDevice, c.stream));
    if (bias) BB_CUDA(cudaMemcpyAsync(dB, bias, (size_t)OC * 4, cudaMemcpyHostToDevice, c.stream));
    if (dY) BB_CUDA(cudaMemcpyAsync(w.dact[1], dY, ny * 4, cudaMemcpyHostToDevice, c.stream));
    if (use_tma) {
        make_lo(c, w.act[0], w.act[0] + w.plane[0], nx);
        make_lo(c, dW, dW + nwp, nw);
        if (dY) make_lo(c, w.dact[1], w.dact[1] + w.plane[1], ny);
        g.x_plane = w.plane[0]; g.w_plane = (long)nwp; g.y_plane = w.plane[1]; g.dx_plane = w.plane[0];
        g.dypad_plane = w.dypad_plane[1]; g.wt_plane = w.wt_plane[1];
    }
    size_t n_out = 0;
    const float* src = nullptr;
    auto run = [&]() {
        if (mode == 0) {
            conv_fwd(c, g, w.act[0], dW, bias ? dB : nullptr, w.act[1], false);
            src = w.act[1]; n_out = ny;
        } else if (mode == 1) {
            conv_bwd_weight(c, g, w.dact[1], w.act[0], dG, nullptr);
            src = dG; n_out = nw;
        } else if (mode == 2) {
            conv_bwd_data(c, g, w.dact[1], dW, w.col, w.dact[0], nullptr);
            src = w.dact[0]; n_out = nx;
        } else {
            throw Error("bb_test_conv: mode must be 0, 1 or 2");
        }
    };
    run();
    if (int iters = env_int("BB_CONV_ITERS", 0)) {   // scratch timing of the layer (stderr)
        cudaEvent_t e0, e1;
        BB_CUDA(cudaEventCreate(&e0)); BB_CUDA(cudaEventCreate(&e1));
        BB_CUDA(cudaEventRecord(e0, c.stream));
        for (int i = 0; i < iters; ++i) run();
        BB_CUDA(cudaEventRecord(e1, c.stream));
        BB_CUDA(cudaStreamSynchronize(c.stream));
        float ms = 0.f;
        BB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        fprintf(stderr, "[conv timing] mode %d tma %d B %d C %d HxW %dx%d OC %d k %d s %d: %.1f us\n", mode, use_tma, B, C, H, W, OC, k, s,
                ms * 1e3f / iters);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    BB_CUDA(cudaMemcpyAsync(out, src, n_out * 4, cudaMemcpyDeviceToHost, c.stream));
    cudaError_t e = cudaStreamSynchronize(c.stream);
    w.release(); net.free_tables();
    cudaFree(dW); cudaFree(dG); cudaFree(dB); c.free_scratch();
    BB_CUDA(e);
    check_device_error("bb_test_conv");
    BB_API_END
}
