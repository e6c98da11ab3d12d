// tc_gemm.cu -- launcher of the tcgen05 GEMM (tc_gemm.cuh) and its test hook.
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "nn.cuh"
#include "tc_gemm.cuh"

namespace bb {

template <int BN, int STAGES, int PF, int MINB>
static void tc_launch(GemmMode mode, const GemmArgs& a, dim3 grid, cudaStream_t s, bool fast = false) {
    constexpr size_t smem = (size_t)STAGES * (2 * tc::BM * 128 + 2 * BN * 128) + 1024;
#define BB_TC_LAUNCH(AK, BK_, AU, BU, FAST_)                                                                    \
    do {                                                                                                        \
        auto kern = tc_gemm_kernel<BN, STAGES, PF, MINB, AK, BK_, AU, BU, FAST_>;                               \
        BB_ENSURE_SMEM(kern, smem);                                                                             \
        launch_pdl(kern, grid, dim3(tc::NTHREADS), smem, s, a);                                                 \
    } while (0)
    switch (mode) {
        case G_FWD: if (fast) BB_TC_LAUNCH(true, true, false, false, true); else BB_TC_LAUNCH(true, true, false, false, false); break;
        case G_FWD_U8: BB_TC_LAUNCH(true, true, true, false, false); break;
        case G_NN: if (fast) BB_TC_LAUNCH(true, false, false, false, true); else BB_TC_LAUNCH(true, false, false, false, false); break;
        case G_WGRAD: if (fast) BB_TC_LAUNCH(false, false, false, false, true); else BB_TC_LAUNCH(false, false, false, false, false); break;
        case G_WGRAD_U8: BB_TC_LAUNCH(false, false, false, true, false); break;
        case G_WGRAD_AU8: BB_TC_LAUNCH(false, false, true, false, false); break;
    }
#undef BB_TC_LAUNCH
    BB_LAUNCHED();
}

__global__ void splitk_reduce_kernel(const float* __restrict__ ws, float* __restrict__ C, int M, int N, int ldc,
                                     int splits, const float* __restrict__ bias, int relu, const float* __restrict__ mask, long c_plane);
__global__ void splitk_reduce8_kernel(const float* __restrict__ ws, float* __restrict__ C, int M, int N, int ldc,
                                      int splits, const float* __restrict__ bias, int relu, const float* __restrict__ mask, long c_plane);

static void publish_error_flag_tc() {
    static std::atomic<uint32_t> done{0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (done.load() & (1u << (dev & 31))) return;
    int* f = device_error_flag();
    BB_CUDA(cudaMemcpyToSymbol(g_tc_err_flag, &f, sizeof(f)));
    done.fetch_or(1u << (dev & 31));
}

bool tc_gemm(const Ctx& c, GemmMode mode, GemmArgs& a) {
    publish_error_flag_tc();
    // target CTAs of a split-K launch, % of SMs.  A/B on the DQN step (us): 40: 361.5, 50: 357.2, 60: 350.2, 75: 349.1,
    // 100: 354.7, 150: 368.4 -- the split launches share the GPU with the other streams' kernels
    static const int fill_pct = getenv("BB_TC_FILL") ? atoi(getenv("BB_TC_FILL")) : 75;
    // BN <= 64 with two CTAs/SM measured fastest on B200 (3 stages x 1 CTA/SM and 128-wide tiles lose on these shapes)
    int BN = a.N <= 32 ? 32 : 64;
    // Both operands of an untransposed weight gradient are stored transposed (scalar st.shared): the producer cost per
    // k-slice is per ROW loaded, so the wider 128x128 tile (3 stages, 1 CTA/SM) wins there -- l1.wgrad 30.8 -> 24.6 us,
    // 64x512x20736 71 -> 49 us on the box -- while it loses on the forward / dgrad shapes (l1.fwd 24.5 -> 35.9 us).
    const bool wgrad_mode = mode == G_WGRAD || mode == G_WGRAD_U8 || mode == G_WGRAD_AU8;
    // Only for long contractions (IQN's 512x3136x16384: 1561 -> 913 us): the 1-CTA/SM, 200 KB tile keeps other streams'
    // CTAs off its SMs, which cost the DQN step 26 us when its small l1.wgrad (K = 256, side stream) took it.
    if (wgrad_mode && !a.trans_out && a.N >= 128 && a.M >= 64 && a.K >= 4096) BN = 128;
    int tm = (a.M + tc::BM - 1) / tc::BM, tn = (a.N + BN - 1) / BN;
    long tiles = (long)tm * tn;
    int kt = (a.K + tc::BK - 1) / tc::BK;
    int split = 1;
    long want = (long)c.sms * fill_pct / 100;
    const bool mapped_out = a.c_rowoff != nullptr;  // the output map lives in the direct epilogue only
    if (tiles * 2 <= want && kt >= 8 && !mapped_out) {  // a split costs a reduce launch: only when under half the SMs would work
        split = (int)std::min<long>((want + tiles - 1) / tiles, kt / 4);
        size_t per = (size_t)a.M * a.N;
        const size_t usable = c.ws_floats - 1024;  // the last 1024 words hold colsum's block counters
        if (per * split > usable) split = (int)(usable / per);
        if (split < 1) split = 1;
    }
    int kps = ((kt + split - 1) / split) * tc::BK;
    split = (a.K + kps - 1) / kps;
    a.split_k = split;
    a.k_per_split = kps;
    a.workspace = c.ws;
    static const int fence_mode = getenv("BB_TC_FENCE") ? atoi(getenv("BB_TC_FENCE")) : 1;
    static const int debug = getenv("BB_TC_DEBUG") ? atoi(getenv("BB_TC_DEBUG")) : 0;  // 1: no global loads, 2: no MMA
    a.fence_mode = fence_mode | (debug << 4);
    dim3 grid(tn, tm, split);
    {
        // branch-free producer loads when every tile and k-slice is whole and every 4-group is 16-byte aligned
        static const int fast_env = getenv("BB_TC_FAST") ? atoi(getenv("BB_TC_FAST")) : 1;
        const bool a_tab = a.a_rowbase || a.a_koff, b_tab = a.b_rowbase || a.b_noff;
        const bool fast = fast_env && (mode == G_FWD || mode == G_NN || mode == G_WGRAD) && a.M % tc::BM == 0 &&
                          a.N % BN == 0 && a.K % tc::BK == 0 && ((reinterpret_cast<uintptr_t>(a.A) | reinterpret_cast<uintptr_t>(a.B)) & 15) == 0 &&
                          (a_tab ? a.tables_vec4 != 0 : (a.lda & 3) == 0) && (b_tab ? a.tables_vec4 != 0 : (a.ldb & 3) == 0);
        switch (BN) {
            case 32: tc_launch<32, 2, 2, 2>(mode, a, grid, c.stream, fast); break;
            case 64: tc_launch<64, 2, 2, 2>(mode, a, grid, c.stream, fast); break;
            default: tc_launch<128, 3, 3, 1>(mode, a, grid, c.stream, fast); break;
        }
    }
    c.mark((BN == 32 ? "tc_gemm128x32" : (BN == 64 ? "tc_gemm128x64" : "tc_gemm128x128")));
    if (split > 1) {
        size_t total = (size_t)a.M * a.N;
        const int rows = a.trans_out ? a.N : a.M, cols = a.trans_out ? a.M : a.N;  // layout of the partials = layout of C
        if (split >= 16) {
            int blocks = (int)std::min<size_t>((total * 8 + 255) / 256, (size_t)c.sms * 8);
            launch_pdl(splitk_reduce8_kernel, dim3(blocks), dim3(256), 0, c.stream, c.ws, a.C, rows, cols, a.ldc, split, a.bias, a.relu, a.mask, a.c_plane);
        } else {
            int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)c.sms * 8);
            launch_pdl(splitk_reduce_kernel, dim3(blocks), dim3(256), 0, c.stream, c.ws, a.C, rows, cols, a.ldc, split, a.bias, a.relu, a.mask, a.c_plane);
        }
        BB_LAUNCHED();
        c.mark("splitk_reduce");
    }
    return true;
}

void tc_trace_read(long long* out) { cudaMemcpyFromSymbol(out, g_tc_trace, sizeof(long long) * 2 * 64 * 8); }

int tc_error_flag() {
    int e = 0;
    cudaMemcpyFromSymbol(&e, g_tc_error, sizeof(int));
    return e;
}

}  // namespace bb

// Test hook: one dense GEMM on the device through either path.
//   mode 0 (forward): C[M][N] = A[M][K] B[N][K]^T (+bias, relu)
//   mode 2 (dgrad):   C[M][N] = A[M][K] B[K][N]
//   mode 3 (wgrad):   C[M][N] = A[K][M]^T B[K][N]
extern "C" int32_t bb_test_gemm(int32_t device, int32_t mode, int32_t use_tc, int32_t M, int32_t N, int32_t K,
                                const float* A, const float* B, const float* bias, int32_t relu, float* C_out) {
    BB_API_BEGIN
    using namespace bb;
    DeviceGuard g(device);
    Ctx c;
    c.device = device; c.sms = num_sms(device); c.stream = device_stream(device);
    c.alloc_scratch(8u << 20);
    // every operand is followed by its lo plane (use_tc = 3: the TMA-fed kernel needs them; C's is checked below)
    const size_t na = ((size_t)M * K + 3) / 4 * 4, nb = ((size_t)N * K + 3) / 4 * 4, nc = ((size_t)M * N + 3) / 4 * 4;
    float *dA = dev_alloc<float>(2 * na), *dB = dev_alloc<float>(2 * nb), *dC = dev_alloc_zero<float>(2 * nc, c.stream), *dbias = nullptr;
    BB_CUDA(cudaMemcpyAsync(dA, A, (size_t)M * K * 4, cudaMemcpyHostToDevice, c.stream));
    BB_CUDA(cudaMemcpyAsync(dB, B, (size_t)N * K * 4, cudaMemcpyHostToDevice, c.stream));
    make_lo(c, dA, dA + na, (size_t)M * K);
    make_lo(c, dB, dB + nb, (size_t)N * K);
    if (bias) {
        dbias = dev_alloc<float>(N);
        BB_CUDA(cudaMemcpyAsync(dbias, bias, (size_t)N * 4, cudaMemcpyHostToDevice, c.stream));
    }
    GemmArgs a = zero_args();
    a.A = dA; a.B = dB; a.C = dC; a.M = M; a.N = N; a.K = K; a.ldc = N; a.bias = dbias; a.relu = relu;
    GemmMode gm;
    if (mode == 0) { gm = G_FWD; a.lda = K; a.ldb = K; }
    else if (mode == 2) { gm = G_NN; a.lda = K; a.ldb = N; }
    else if (mode == 3) { gm = G_WGRAD; a.lda = M; a.ldb = N; }
    else throw Error("bb_test_gemm: mode must be 0, 2 or 3");
    if (use_tc == 3) { a.a_plane = (long)na; a.b_plane = (long)nb; }
    if (use_tc == 3 || (use_tc == 2 && mode != 3)) a.c_plane = (long)nc;   // (weight gradients feed Adam only: the skinny wgrad kernel writes no lo plane)
    if (use_tc == 3) {
        BB_CHECK(tma_gemm(c, gm, a), "bb_test_gemm: the TMA path declined this problem");
    } else if (use_tc == 2) {
        gemm(c, gm, a);  // the dispatcher the layers use: skinny kernels, tcgen05 tiles or CUDA-core tiles by shape
    } else if (use_tc) {
        tc_gemm(c, gm, a);
    } else {
        gemm_simt(c, gm, a);
    }
    BB_CUDA(cudaMemcpyAsync(C_out, dC, (size_t)M * N * 4, cudaMemcpyDeviceToHost, c.stream));
    std::vector<float> lo;
    if (a.c_plane) {
        lo.resize((size_t)M * N);
        BB_CUDA(cudaMemcpyAsync(lo.data(), dC + nc, (size_t)M * N * 4, cudaMemcpyDeviceToHost, c.stream));
    }
    cudaError_t e = cudaStreamSynchronize(c.stream);
    int flag = tc_error_flag();
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dbias); c.free_scratch();
    BB_CUDA(e);
    BB_CHECK(flag == 0, "tcgen05 pipeline timed out (g_tc_error)");
    check_device_error("bb_test_gemm");
    for (size_t i = 0; i < lo.size(); ++i) {   // every writer of a planed tensor must also write its lo plane
        uint32_t u;
        memcpy(&u, &C_out[i], 4);
        u &= 0xffffe000u;
        float hi;
        memcpy(&hi, &u, 4);
        BB_CHECK(lo[i] == C_out[i] - hi, "bb_test_gemm: the lo plane of C is not C - tf32_trunc(C)");
    }
    BB_API_END
}

// Timing hook: `iters` back-to-back launches of one dense GEMM (device-resident operands),
// CUDA events on the launching stream; returns the mean milliseconds per GEMM.
extern "C" int32_t bb_bench_gemm(int32_t device, int32_t mode, int32_t use_tc, int32_t M, int32_t N, int32_t K,
                                 int32_t iters, float* ms_out) {
    BB_API_BEGIN
    using namespace bb;
    DeviceGuard g(device);
    Ctx c;
    c.device = device; c.sms = num_sms(device); c.stream = device_stream(device);
    c.alloc_scratch(8u << 20);
    size_t na = ((size_t)M * K + 3) / 4 * 4, nb = ((size_t)N * K + 3) / 4 * 4, nc = ((size_t)M * N + 3) / 4 * 4;
    float *dA = dev_alloc<float>(2 * na), *dB = dev_alloc<float>(2 * nb), *dC = dev_alloc_zero<float>(2 * nc, c.stream);
    fill_uniform(c, dA, na, 1.0f, 1);
    fill_uniform(c, dB, nb, 1.0f, 2);
    make_lo(c, dA, dA + na, na);
    make_lo(c, dB, dB + nb, nb);
    GemmArgs a = zero_args();
    a.A = dA; a.B = dB; a.C = dC; a.M = M; a.N = N; a.K = K; a.ldc = N;
    if (use_tc == 3) { a.a_plane = (long)na; a.b_plane = (long)nb; a.c_plane = (long)nc; }
    GemmMode gm;
    if (mode == 0) { gm = G_FWD; a.lda = K; a.ldb = K; }
    else if (mode == 2) { gm = G_NN; a.lda = K; a.ldb = N; }
    else { gm = G_WGRAD; a.lda = M; a.ldb = N; }
    cudaEvent_t e0, e1;
    BB_CUDA(cudaEventCreate(&e0));
    BB_CUDA(cudaEventCreate(&e1));
    auto run = [&]() {
        GemmArgs b = a;
        if (use_tc == 3) BB_CHECK(tma_gemm(c, gm, b), "bb_bench_gemm: the TMA path declined this problem");
        else if (use_tc) tc_gemm(c, gm, b);
        else gemm_simt(c, gm, b);
    };
    for (int i = 0; i < 3; ++i) run();
    BB_CUDA(cudaEventRecord(e0, c.stream));
    for (int i = 0; i < iters; ++i) run();
    BB_CUDA(cudaEventRecord(e1, c.stream));
    BB_CUDA(cudaStreamSynchronize(c.stream));
    float ms = 0.f;
    BB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *ms_out = ms / iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(dA); cudaFree(dB); cudaFree(dC); c.free_scratch();
    BB_CHECK(tc_error_flag() == 0, "tcgen05 pipeline timed out (g_tc_error)");
    check_device_error("bb_bench_gemm");
    BB_API_END
}

// Debug: clock64 stamps of the last traced tcgen05 launch (BB_TC_DEBUG bit 16): [2 roles][64 k-slices][4]
extern "C" int32_t bb_debug_tc_trace(int64_t* out) {
    BB_API_BEGIN
    bb::tc_trace_read(reinterpret_cast<long long*>(out));
    BB_API_END
}
