// tc_gemm.cu -- launcher of the tcgen05 GEMM (tc_gemm.cuh) and its test hook.
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "nn.cuh"
#include "tc_gemm.cuh"

namespace bb {

template <int BN>
static void tc_launch(GemmMode mode, const GemmArgs& a, dim3 grid, cudaStream_t s) {
    constexpr size_t smem = (size_t)tc::STAGES * (2 * tc::BM * 128 + 2 * BN * 128) + 1024;
#define BB_TC_LAUNCH(AK, BK_, AU, BU)                                                                           \
    do {                                                                                                        \
        auto kern = tc_gemm_kernel<BN, AK, BK_, AU, BU>;                                                        \
        static bool configured = false;                                                                         \
        if (!configured) {                                                                                      \
            BB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
            configured = true;                                                                                  \
        }                                                                                                       \
        kern<<<grid, tc::NTHREADS, smem, s>>>(a);                                                               \
    } while (0)
    switch (mode) {
        case G_FWD: BB_TC_LAUNCH(true, true, false, false); break;
        case G_FWD_U8: BB_TC_LAUNCH(true, true, true, false); break;
        case G_NN: BB_TC_LAUNCH(true, false, false, false); break;
        case G_WGRAD: BB_TC_LAUNCH(false, false, false, false); break;
        case G_WGRAD_U8: BB_TC_LAUNCH(false, false, false, true); break;
    }
#undef BB_TC_LAUNCH
    BB_LAUNCHED();
}

__global__ void splitk_reduce_kernel(const float* __restrict__ ws, float* __restrict__ C, int M, int N, int ldc,
                                     int splits, const float* __restrict__ bias, int relu, const float* __restrict__ mask);
__global__ void splitk_reduce8_kernel(const float* __restrict__ ws, float* __restrict__ C, int M, int N, int ldc,
                                      int splits, const float* __restrict__ bias, int relu, const float* __restrict__ mask);

bool tc_gemm(const Ctx& c, GemmMode mode, GemmArgs& a) {
    static const int fill_pct = getenv("BB_TC_FILL") ? atoi(getenv("BB_TC_FILL")) : 100;  // target CTAs, % of SMs
    int BN = a.N <= 32 ? 32 : (a.N <= 64 ? 64 : 128);
    int tm = (a.M + tc::BM - 1) / tc::BM, tn = (a.N + BN - 1) / BN;
    long tiles = (long)tm * tn;
    int kt = (a.K + tc::BK - 1) / tc::BK;
    int split = 1;
    long want = (long)c.sms * fill_pct / 100;
    if (tiles < want && kt >= 8) {
        split = (int)std::min<long>((want + tiles - 1) / tiles, kt / 4);
        size_t per = (size_t)a.M * a.N;
        if (per * split > c.ws_floats) split = (int)(c.ws_floats / per);
        if (split < 1) split = 1;
    }
    int kps = ((kt + split - 1) / split) * tc::BK;
    split = (a.K + kps - 1) / kps;
    a.split_k = split;
    a.k_per_split = kps;
    a.workspace = c.ws;
    dim3 grid(tn, tm, split);
    switch (BN) {
        case 32: tc_launch<32>(mode, a, grid, c.stream); break;
        case 64: tc_launch<64>(mode, a, grid, c.stream); break;
        default: tc_launch<128>(mode, a, grid, c.stream); break;
    }
    c.mark(BN == 32 ? "tc_gemm128x32" : (BN == 64 ? "tc_gemm128x64" : "tc_gemm128x128"));
    if (split > 1) {
        size_t total = (size_t)a.M * a.N;
        if (split >= 16) {
            int blocks = (int)std::min<size_t>((total * 8 + 255) / 256, (size_t)c.sms * 8);
            splitk_reduce8_kernel<<<blocks, 256, 0, c.stream>>>(c.ws, a.C, a.M, a.N, a.ldc, split, a.bias, a.relu, a.mask);
        } else {
            int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)c.sms * 8);
            splitk_reduce_kernel<<<blocks, 256, 0, c.stream>>>(c.ws, a.C, a.M, a.N, a.ldc, split, a.bias, a.relu, a.mask);
        }
        BB_LAUNCHED();
        c.mark("splitk_reduce");
    }
    return true;
}

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
    c.ws_floats = 8u << 20;
    c.ws = dev_alloc<float>(c.ws_floats);
    size_t na = (size_t)M * K, nb = (size_t)N * K, nc = (size_t)M * N;
    float *dA = dev_alloc<float>(na), *dB = dev_alloc<float>(nb), *dC = dev_alloc_zero<float>(nc, c.stream), *dbias = nullptr;
    BB_CUDA(cudaMemcpyAsync(dA, A, na * 4, cudaMemcpyHostToDevice, c.stream));
    BB_CUDA(cudaMemcpyAsync(dB, B, nb * 4, cudaMemcpyHostToDevice, c.stream));
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
    if (use_tc) {
        tc_gemm(c, gm, a);
    } else {
        gemm_simt(c, gm, a);
    }
    BB_CUDA(cudaMemcpyAsync(C_out, dC, nc * 4, cudaMemcpyDeviceToHost, c.stream));
    cudaError_t e = cudaStreamSynchronize(c.stream);
    int flag = tc_error_flag();
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dbias); cudaFree(c.ws);
    BB_CUDA(e);
    BB_CHECK(flag == 0, "tcgen05 pipeline timed out (g_tc_error)");
    BB_API_END
}
