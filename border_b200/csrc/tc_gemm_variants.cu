// tc_gemm_variants.cu -- launchers of the experimental tcgen05 GEMM variants (BB_TC_CFG = 3..6), kept as measured
// A/B evidence: persistent flat-pipelined kernel (tc_gemm2.cuh), A operand in tensor memory (tc_gemm3.cuh), cp.async
// raw ring + split pass (tc_gemm4.cuh), tensor-memory A + cp.async (tc_gemm5.cuh).  All parity-green, all slower than
// the default kernel of tc_gemm.cuh on B200 (profiles/r01_summary.md).  Own translation unit: they are 3/4 of the
// instantiations and would otherwise serialise the build.
#include <stdlib.h>
#include <algorithm>
#include "nn.cuh"
#include "tc_gemm.cuh"
#include "tc_gemm2.cuh"
#include "tc_gemm3.cuh"
#include "tc_gemm4.cuh"
#include "tc_gemm5.cuh"

namespace bb {

template <int BN, int STAGES, int PF>
static void tc_launch_persist(GemmMode mode, const GemmArgs& a, int tm, int tn, int total, int ctas, cudaStream_t s) {
    constexpr size_t smem = (size_t)STAGES * (2 * tc::BM * 128 + 2 * BN * 128) + 1024;
#define BB_TC_LAUNCH(AK, BK_, AU, BU)                                                                           \
    do {                                                                                                        \
        auto kern = tc_gemm_persist_kernel<BN, STAGES, PF, AK, BK_, AU, BU>;                                    \
        static bool configured = false;                                                                         \
        if (!configured) {                                                                                      \
            BB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
            configured = true;                                                                                  \
        }                                                                                                       \
        kern<<<ctas, tc2::NTHREADS, smem, s>>>(a, tm, tn, total);                                               \
    } while (0)
    switch (mode) {
        case G_FWD: BB_TC_LAUNCH(true, true, false, false); break;
        case G_FWD_U8: BB_TC_LAUNCH(true, true, true, false); break;
        case G_NN: BB_TC_LAUNCH(true, false, false, false); break;
        case G_WGRAD: BB_TC_LAUNCH(false, false, false, false); break;
        case G_WGRAD_U8: BB_TC_LAUNCH(false, false, false, true); break;
        case G_WGRAD_AU8: BB_TC_LAUNCH(false, false, true, false); break;
    }
#undef BB_TC_LAUNCH
    BB_LAUNCHED();
}

template <int BN>
static void tc_launch_tmem(GemmMode mode, const GemmArgs& a, dim3 grid, cudaStream_t s) {
    constexpr size_t smem = (size_t)tc3::STAGES * (2 * BN * 128) + 1024;
#define BB_TC_LAUNCH(AK, BK_, AU, BU)                                                                           \
    do {                                                                                                        \
        auto kern = tc_gemm_tmem_kernel<BN, AK, BK_, AU, BU>;                                                   \
        static bool configured = false;                                                                         \
        if (!configured) {                                                                                      \
            BB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
            configured = true;                                                                                  \
        }                                                                                                       \
        kern<<<grid, tc3::NTHREADS, smem, s>>>(a);                                                              \
    } while (0)
    switch (mode) {
        case G_FWD: BB_TC_LAUNCH(true, true, false, false); break;
        case G_FWD_U8: BB_TC_LAUNCH(true, true, true, false); break;
        case G_NN: BB_TC_LAUNCH(true, false, false, false); break;
        case G_WGRAD: BB_TC_LAUNCH(false, false, false, false); break;
        case G_WGRAD_U8: BB_TC_LAUNCH(false, false, false, true); break;
        case G_WGRAD_AU8: BB_TC_LAUNCH(false, false, true, false); break;
    }
#undef BB_TC_LAUNCH
    BB_LAUNCHED();
}

template <int BN, int STAGES, int DEPTH, int MINB>
static void tc_launch_async(GemmMode mode, const GemmArgs& a, dim3 grid, cudaStream_t s) {
    constexpr int B_LD = (BN * 8 + tc::NPROD - 1) / tc::NPROD;
    constexpr size_t smem = (size_t)STAGES * (2 * tc::BM * 128 + 2 * BN * 128) + (size_t)DEPTH * (4 + B_LD) * tc::NPROD * 16 + 1024;
#define BB_TC_LAUNCH(AK, BK_, AU, BU)                                                                           \
    do {                                                                                                        \
        auto kern = tc_gemm_async_kernel<BN, STAGES, DEPTH, MINB, AK, BK_, AU, BU>;                             \
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
        case G_WGRAD_AU8: BB_TC_LAUNCH(false, false, true, false); break;
    }
#undef BB_TC_LAUNCH
    BB_LAUNCHED();
}

template <int BN, int S, int RD, int MINB>
static bool tc_launch_ta(GemmMode mode, const GemmArgs& a, dim3 grid, cudaStream_t s) {
    constexpr size_t smem = tc5::smem_bytes(BN, S, RD);
#define BB_TC_LAUNCH(AK, BK_, AU)                                                                               \
    do {                                                                                                        \
        auto kern = tc_gemm_ta_kernel<BN, S, RD, MINB, AK, BK_, AU>;                                            \
        static bool configured = false;                                                                         \
        if (!configured) {                                                                                      \
            BB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
            configured = true;                                                                                  \
        }                                                                                                       \
        kern<<<grid, tc5::NTHREADS, smem, s>>>(a);                                                              \
    } while (0)
    switch (mode) {
        case G_FWD: BB_TC_LAUNCH(true, true, false); break;
        case G_FWD_U8: BB_TC_LAUNCH(true, true, true); break;
        case G_NN: BB_TC_LAUNCH(true, false, false); break;
        case G_WGRAD: BB_TC_LAUNCH(false, false, false); break;
        default: return false;
    }
#undef BB_TC_LAUNCH
    BB_LAUNCHED();
    return true;
}


bool tc_gemm_variant(const Ctx& c, int cfg2, int avar, GemmMode mode, const GemmArgs& a, dim3 grid, int BN, int tm, int tn,
                     int split, bool v1_only, bool mapped_out, const char** tag) {
    if (BN != 32 && BN != 64) return false;
    if (cfg2 == 6 && !mapped_out) {  // A in tensor memory + cp.async staging (tc_gemm5.cuh)
        bool done;
        if (avar == 0) done = BN == 32 ? tc_launch_ta<32, 3, 1, 2>(mode, a, grid, c.stream) : tc_launch_ta<64, 3, 1, 2>(mode, a, grid, c.stream);
        else done = BN == 32 ? tc_launch_ta<32, 6, 2, 1>(mode, a, grid, c.stream) : tc_launch_ta<64, 6, 2, 1>(mode, a, grid, c.stream);
        if (done) { *tag = BN == 32 ? "tc_ta128x32" : "tc_ta128x64"; return true; }
        return false;
    }
    if (cfg2 == 5 && !mapped_out) {  // cp.async operand pipeline (tc_gemm4.cuh)
        if (avar == 0) {
            if (BN == 32) tc_launch_async<32, 2, 4, 1>(mode, a, grid, c.stream);
            else tc_launch_async<64, 2, 4, 1>(mode, a, grid, c.stream);
        } else if (avar == 1) {
            if (BN == 32) tc_launch_async<32, 1, 2, 2>(mode, a, grid, c.stream);
            else tc_launch_async<64, 1, 2, 2>(mode, a, grid, c.stream);
        } else {
            if (BN == 32) tc_launch_async<32, 2, 3, 1>(mode, a, grid, c.stream);
            else tc_launch_async<64, 2, 3, 1>(mode, a, grid, c.stream);
        }
        *tag = BN == 32 ? "tc_async128x32" : "tc_async128x64";
        return true;
    }
    if (cfg2 == 4 && !v1_only) {  // A operand in tensor memory (tc_gemm3.cuh)
        if (BN == 32) tc_launch_tmem<32>(mode, a, grid, c.stream);
        else tc_launch_tmem<64>(mode, a, grid, c.stream);
        *tag = BN == 32 ? "tc_tmem128x32" : "tc_tmem128x64";
        return true;
    }
    if (cfg2 == 3 && !v1_only) {  // persistent, flat-pipelined kernel (tc_gemm2.cuh)
        int total = tm * tn * split;
        int ctas = std::min(total, c.sms);
        if (BN == 32) tc_launch_persist<32, 4, 3>(mode, a, tm, tn, total, ctas, c.stream);
        else tc_launch_persist<64, 4, 3>(mode, a, tm, tn, total, ctas, c.stream);
        *tag = BN == 32 ? "tc_persist128x32" : "tc_persist128x64";
        return true;
    }
    return false;
}

void tc_trace_read_variants(long long* out) { cudaMemcpyFromSymbol(out, g_tc_trace, sizeof(long long) * 2 * 64 * 8); }

}  // namespace bb

// Debug: clock64 stamps of the last traced VARIANT launch (tools/tc_trace.py)
extern "C" int32_t bb_debug_tc_trace_variants(int64_t* out) {
    BB_API_BEGIN
    bb::tc_trace_read_variants(reinterpret_cast<long long*>(out));
    BB_API_END
}
