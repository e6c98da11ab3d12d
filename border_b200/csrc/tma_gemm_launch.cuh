// tma_gemm_launch.cuh -- launch templates of the TMA-fed tcgen05 GEMM, shared by the translation units that instantiate
// it (tma_gemm.cu: 3xTF32; tma_gemm_fast.cu: the single-pass `fast` precision mode) so that they compile in parallel.
#pragma once
#include <atomic>
#include "gemm.cuh"
#include "nn.cuh"
#include "tma_gemm.cuh"

namespace bb {

template <int BN, int STAGES, int AK, int BKIND, int PASSES, int MINB>
static void launch_one(const CUtensorMap& ta, const CUtensorMap& tb, const tg::Args& g, dim3 grid, cudaStream_t s) {
    constexpr size_t smem = (size_t)STAGES * (tg::BM * 128 + (PASSES == 3 ? 2 : 1) * BN * 128) + 1024;
    auto kern = tma_gemm_kernel<BN, STAGES, AK, BKIND, PASSES, MINB>;
    static std::atomic<uint32_t> configured{0};  // bit per device: the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(configured.load() & (1u << dev))) {
        BB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.fetch_or(1u << dev);
    }
    launch_pdl(kern, grid, dim3(tg::NTHREADS), smem, s, ta, tb, g);
    BB_LAUNCHED();
}

template <int AK, int BKIND, int PASSES>
static void launch_cfg(int BN, int cfg, const CUtensorMap& ta, const CUtensorMap& tb, const tg::Args& g, dim3 grid,
                       cudaStream_t s) {
    // BN <= 64: a stage is 32 KB ([A][B][lo(B)]; lo(A) lives in tensor memory), so two CTAs with 3-stage rings share an SM
    // (one's prologue / epilogue overlaps the other's main loop, and a grid slightly over 148 tiles is still one wave);
    // cfg 1: one CTA per SM with a 6-stage ring.  BN = 128: 48 KB stages, 4 of them, one CTA per SM.
    if (BN == 32) {
        if (cfg == 1) launch_one<32, 6, AK, BKIND, PASSES, 1>(ta, tb, g, grid, s);
        else launch_one<32, 3, AK, BKIND, PASSES, 2>(ta, tb, g, grid, s);
    } else if (BN == 64) {
        if (cfg == 1) launch_one<64, 6, AK, BKIND, PASSES, 1>(ta, tb, g, grid, s);
        else launch_one<64, 3, AK, BKIND, PASSES, 2>(ta, tb, g, grid, s);
    } else {
        launch_one<128, 4, AK, BKIND, PASSES, 1>(ta, tb, g, grid, s);
    }
}


// one (A kind, B kind) combination of a precision mode; false => unknown combination
template <int PASSES>
static bool launch_combo(int AK, int BKIND, int BN, int cfg, const CUtensorMap& ta, const CUtensorMap& tb,
                         const tg::Args& g, dim3 grid, cudaStream_t s) {
    using namespace tg;
    if (AK == OP_K_TILED && BKIND == OP_K_TILED) launch_cfg<OP_K_TILED, OP_K_TILED, PASSES>(BN, cfg, ta, tb, g, grid, s);
    else if (AK == OP_K_IM2COL && BKIND == OP_K_TILED) launch_cfg<OP_K_IM2COL, OP_K_TILED, PASSES>(BN, cfg, ta, tb, g, grid, s);
    else if (AK == OP_K_TILED && BKIND == OP_MN_TILED) launch_cfg<OP_K_TILED, OP_MN_TILED, PASSES>(BN, cfg, ta, tb, g, grid, s);
    else if (AK == OP_MN_TILED && BKIND == OP_MN_TILED) launch_cfg<OP_MN_TILED, OP_MN_TILED, PASSES>(BN, cfg, ta, tb, g, grid, s);
    else if (AK == OP_MN_IM2COL && BKIND == OP_MN_TILED) launch_cfg<OP_MN_IM2COL, OP_MN_TILED, PASSES>(BN, cfg, ta, tb, g, grid, s);
    else return false;
    return true;
}

bool tma_launch_fast(int AK, int BKIND, int BN, int cfg, const CUtensorMap& ta, const CUtensorMap& tb,
                     const tg::Args& g, dim3 grid, cudaStream_t s);   // tma_gemm_fast.cu

}  // namespace bb
