// tma_gemm_fast.cu -- the single-pass instantiations of the TMA-fed tcgen05 GEMM (one TF32 product per fp32 product: the
// `fast` precision mode, BB_TMA_PASSES=1), in their own translation unit so that they compile beside tma_gemm.cu.
#include "tma_gemm_launch.cuh"

namespace bb {
bool tma_launch_fast(int AK, int BKIND, int BN, int cfg, const CUtensorMap& ta, const CUtensorMap& tb,
                     const tg::Args& g, dim3 grid, cudaStream_t s) {
    return launch_combo<1>(AK, BKIND, BN, cfg, ta, tb, g, grid, s);
}
}  // namespace bb
