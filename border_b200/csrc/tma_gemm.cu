// tma_gemm.cu -- host side of the TMA-fed tcgen05 GEMM (tma_gemm.cuh): tensor-map construction and cache,
// eligibility, tile / split-K choice, launch; plus the lo-plane helper kernels.
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <mutex>
#include <vector>
#include "tma_gemm_launch.cuh"

namespace bb {

// ------------------------------------------------------------------------------- driver entry points

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void* driver_fn(const char* name) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return fn;
}
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn f = (EncodeTiledFn)driver_fn("cuTensorMapEncodeTiled");
    return f;
}
static EncodeIm2colFn encode_im2col() {
    static EncodeIm2colFn f = (EncodeIm2colFn)driver_fn("cuTensorMapEncodeIm2col");
    return f;
}

// ------------------------------------------------------------------------------- tensor-map cache
// A map depends only on (base pointer, geometry); the buffers of a workspace never move, so the ~30 maps of an update are
// built once (first eager steps) and the CUDA graph keeps its own copies (they are __grid_constant__ kernel parameters).

struct MapKey {
    const void* p;
    long v[16];
    bool operator<(const MapKey& o) const {
        if (p != o.p) return p < o.p;
        return memcmp(v, o.v, sizeof(v)) < 0;
    }
};
static std::map<MapKey, CUtensorMap>& map_cache() {
    static std::map<MapKey, CUtensorMap> c;
    return c;
}
static std::mutex g_map_mu;
std::atomic<uint64_t> g_tma_launches{0}, g_tma_rejects{0};

// dims/strides innermost first; strides in BYTES for dims 1..rank-1
static bool tiled_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
                      const uint32_t* box, bool mn_major = false) {
    MapKey k;
    memset(&k, 0, sizeof(k));
    k.p = base; k.v[0] = 100 + rank + (mn_major ? 50 : 0);
    for (int i = 0; i < rank; ++i) { k.v[1 + i] = (long)dims[i]; k.v[6 + i] = (long)box[i]; }
    for (int i = 0; i + 1 < rank; ++i) k.v[11 + i] = (long)strides_b[i];
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = map_cache().find(k);
    if (it != map_cache().end()) { *out = it->second; return true; }
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    cuuint64_t gd[5], gs[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_b[i];
    CUtensorMap m;
    // MN-major 32-bit UMMA operands live in the 32-byte-atom flavour of the 128-byte swizzle (see tma_gemm.cuh: desc_mn)
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    map_cache()[k] = m;
    *out = m;
    return true;
}

// NHWC float tensor [N][H][W][C] walked by a KH x KW filter with traversal stride S (no padding): box = 32 channels x
// `pixels` filter positions.  Bounding box of the filter base: lower corner 0, upper corner -(K-1).
static bool im2col_map(CUtensorMap* out, const float* base, const TmaConv& cv, int pixels, bool mn_major = false) {
    MapKey k;
    memset(&k, 0, sizeof(k));
    k.p = base; k.v[0] = mn_major ? 250 : 200; k.v[1] = cv.N; k.v[2] = cv.H; k.v[3] = cv.W; k.v[4] = cv.C; k.v[5] = cv.KH; k.v[6] = cv.KW;
    k.v[7] = cv.S; k.v[8] = pixels; k.v[9] = cv.pad;
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = map_cache().find(k);
    if (it != map_cache().end()) { *out = it->second; return true; }
    EncodeIm2colFn enc = encode_im2col();
    if (!enc) return false;
    cuuint64_t gd[4] = {(cuuint64_t)cv.C, (cuuint64_t)cv.W, (cuuint64_t)cv.H, (cuuint64_t)cv.N};
    cuuint64_t gs[3] = {(cuuint64_t)cv.C * 4, (cuuint64_t)cv.W * cv.C * 4, (cuuint64_t)cv.H * cv.W * cv.C * 4};
    // bounding box of the filter bases: [-pad, W + pad - (KW - 1)) -- elements outside the tensor read as zero
    int lower[2] = {-cv.pad, -cv.pad};
    int upper[2] = {cv.pad - (cv.KW - 1), cv.pad - (cv.KH - 1)};
    cuuint32_t es[4] = {1, (cuuint32_t)cv.S, (cuuint32_t)cv.S, 1};
    CUtensorMap m;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gd, gs, lower, upper, 32, (cuuint32_t)pixels, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    // Drivers up to 13.1 mis-encode im2col maps of tensors under 128 KiB (the same correction CUTLASS applies in
    // cute/atom/copy_traits_sm90_im2col.hpp): clear bit 21 of the second descriptor word.
    int drv = 0;
    if (cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010 && (size_t)cv.N * cv.H * cv.W * cv.C * 4 < 131072)
        reinterpret_cast<uint64_t*>(&m)[1] &= ~(1ull << 21);
    map_cache()[k] = m;
    *out = m;
    return true;
}

void tma_forget_maps() {  // buffers were freed: their addresses may be reused with another geometry (keys include the geometry, so
                          // stale entries are harmless; this only bounds the cache)
    std::lock_guard<std::mutex> lk(g_map_mu);
    if (map_cache().size() > 4096) map_cache().clear();
}

// ------------------------------------------------------------------------------- lo planes

__global__ void make_lo_kernel(const float* __restrict__ x, float* __restrict__ lo, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) lo[i] = tg::tf32_lo(x[i]);
}
void make_lo(const Ctx& c, const float* x, float* lo, size_t n) {
    if (!n) return;
    make_lo_kernel<<<(int)std::min<size_t>((n + 255) / 256, (size_t)c.sms * 8), 256, 0, c.stream>>>(x, lo, n);
    BB_LAUNCHED();
    c.mark("make_lo");
}

// ------------------------------------------------------------------------------- debug trace
static long long* tma_trace_buffer() {
    static long long* buf = nullptr;
    if (!buf) {
        BB_CUDA(cudaMalloc(&buf, (3 * 64 * 4 + 4 * 1024) * sizeof(long long)));
        BB_CUDA(cudaMemset(buf, 0, (3 * 64 * 4 + 4 * 1024) * sizeof(long long)));
    }
    return buf;
}

// ------------------------------------------------------------------------------- launch

static int env_i(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

// false => not handled (the caller falls back to the SIMT-producer tcgen05 kernel or the CUDA-core tiles)
bool tma_gemm(const Ctx& c, GemmMode mode, GemmArgs& a) {
    using namespace tg;
    if (!env_i("BB_TMA", 1)) return false;
    const int passes = env_i("BB_TMA_PASSES", c.passes) == 1 ? 1 : 3;
    if (mode != G_FWD && mode != G_NN && mode != G_WGRAD) return false;
    if (!c.tickets) return false;
    if ((reinterpret_cast<uintptr_t>(a.A) | reinterpret_cast<uintptr_t>(a.B)) & 15) return false;
    if (a.c_rowoff && !a.tables_vec4) return false;
    const bool a_gather = a.a_rowbase || a.a_koff;
    if (a_gather && !a.a_conv) return false;
    if (a.b_rowbase || a.b_noff) return false;
    const float* A = reinterpret_cast<const float*>(a.A);
    const float* B = reinterpret_cast<const float*>(a.B);

    int AK, BKIND;
    CUtensorMap ta, tb;
    memset(&ta, 0, sizeof(ta)); memset(&tb, 0, sizeof(tb));
    Args g;
    memset(&g, 0, sizeof(g));
    g.ga.flip_w = g.ga.flip_h = -1;
    // One tcgen05.mma costs the issuing thread ~55-64 cycles whatever its width up to N = 128 (tools/tma_probe.cu), so wide
    // layers take 128-column tiles: x.[y; lo(y)] becomes ONE N = 256 instruction.  (BB_TMA_BN=64 forces the narrow tile.)
    // Only for grids of several waves, though: on the DQN shapes (a few hundred 128x64 tiles) the narrow tile measured faster
    // (l1.fwd 24.8 vs 28.8 us, c2.dgrad 37.7 vs 40.9), on 8192^2 x 1024 the wide one (659 vs 719 us).
    int BN = a.N >= 64 ? 64 : 32;
    if (a.N >= 128 && a.N % 128 == 0 && env_i("BB_TMA_BN", 128) == 128 &&
        (long)((a.M + BM - 1) / BM) * ((a.N + 63) / 64) > 4L * c.sms)
        BN = 128;

    // ---- A
    if (mode == G_FWD || mode == G_NN) {
        if (a.a_conv) {
            const TmaConv& cv = *a.a_conv;
            if (cv.C % 32) return false;
            AK = OP_K_IM2COL;
            const int OH = (cv.H + 2 * cv.pad - cv.KH) / cv.S + 1, OW = (cv.W + 2 * cv.pad - cv.KW) / cv.S + 1;
            if (a.M != cv.N * OH * OW || a.K != cv.KH * cv.KW * cv.C) return false;
            if (!im2col_map(&ta, A, cv, 128)) { g_tma_rejects++; return false; }
            g.ga.ow = OW; g.ga.ohw = OH * OW; g.ga.stride = cv.S; g.ga.kw = cv.KW; g.ga.cblocks = cv.C / 32; g.ga.pad = cv.pad;
            if (cv.flip) { g.ga.flip_w = cv.KW - 1; g.ga.flip_h = cv.KH - 1; }
        } else {
            if (a.lda & 3) return false;
            AK = OP_K_TILED;
            uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.M};
            uint64_t st[1] = {(uint64_t)a.lda * 4};
            uint32_t box[2] = {32, 128};
            if (!tiled_map(&ta, A, 2, dims, st, box)) { g_tma_rejects++; return false; }
        }
    } else {  // G_WGRAD: A(m, k) m-contiguous
        if (a.a_conv) {
            // transposed im2col matrix: GEMM rows = (kh, kw, c), contraction over filter positions
            const TmaConv& cv = *a.a_conv;
            if (cv.C % 32 || cv.pad) return false;
            AK = OP_MN_IM2COL;
            const int OH = (cv.H - cv.KH) / cv.S + 1, OW = (cv.W - cv.KW) / cv.S + 1;
            if (a.K != cv.N * OH * OW || a.M != cv.KH * cv.KW * cv.C) return false;
            if (!im2col_map(&ta, A, cv, 32, true)) { g_tma_rejects++; return false; }
            g.ga.ow = OW; g.ga.ohw = OH * OW; g.ga.stride = cv.S; g.ga.kw = cv.KW; g.ga.cblocks = cv.C / 32;
            g.ga.nblocks = cv.KH * cv.KW * (cv.C / 32);
        } else {
            if ((a.lda & 3) || (a.M & 31)) return false;
            AK = OP_MN_TILED;
            uint64_t dims[3] = {32, (uint64_t)a.K, (uint64_t)a.M / 32};
            uint64_t st[2] = {(uint64_t)a.lda * 4, 128};
            uint32_t box[3] = {32, 32, 4};
            if (!tiled_map(&ta, A, 3, dims, st, box, true)) { g_tma_rejects++; return false; }
        }
    }
    // ---- B
    if (mode == G_FWD) {
        if (a.ldb & 3) return false;
        BKIND = OP_K_TILED;
        uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.N};
        uint64_t st[1] = {(uint64_t)a.ldb * 4};
        uint32_t box[2] = {32, (uint32_t)BN};
        if (!tiled_map(&tb, B, 2, dims, st, box)) { g_tma_rejects++; return false; }
    } else {
        if ((a.ldb & 3) || (a.N & 31)) return false;
        BKIND = OP_MN_TILED;
        uint64_t dims[3] = {32, (uint64_t)a.K, (uint64_t)a.N / 32};
        uint64_t st[2] = {(uint64_t)a.ldb * 4, 128};
        uint32_t box[3] = {32, 32, (uint32_t)BN / 32};
        if (!tiled_map(&tb, B, 3, dims, st, box, true)) { g_tma_rejects++; return false; }
    }

    // ---- tiles and split-K (finished inside the kernel by the last CTA of each tile)
    const int tm = (a.M + BM - 1) / BM, tn = (a.N + BN - 1) / BN;
    const long tiles = (long)tm * tn;
    const int nks = (a.K + BK - 1) / BK;
    // The fp32 operand planes make these kernels L2-bandwidth bound (48 KB per 128x64x32 slice and CTA): beyond about half
    // the SMs more CTAs add no throughput, only partial tiles for the finishing CTA to fold -- at most 16 splits.
    static const int fill_pct = env_i("BB_TMA_FILL", 50);
    const long want = (long)c.sms * fill_pct / 100;
    int split = 1;
    if (tiles * 2 <= want && nks >= 8 && !a.c_rowoff) {
        // as many splits as keep every CTA of the launch resident at once (one per SM with the deep ring): a second,
        // mostly empty wave doubled l1.fwd's time (160 CTAs on 148 SMs)
        static const int split_cap = env_i("BB_TMA_SPLIT_CAP", 16);
        split = (int)std::min<long>(std::min<long>(want / tiles, split_cap), nks / 4);
        const size_t per = (size_t)a.M * a.N;
        const size_t usable = c.ws_floats - 1024;
        if (per * split > usable) split = (int)(usable / per);
        if (split < 1) split = 1;
    }
    if (tiles > (long)c.n_tickets) split = 1;
    int sps = (nks + split - 1) / split;
    split = (nks + sps - 1) / sps;
    g.M = a.M; g.N = a.N; g.K = a.K;
    g.slices_per_split = sps; g.split_k = split;
    g.C = a.C; g.c_plane = a.c_plane; g.ldc = a.ldc;
    g.bias = a.bias; g.mask = a.mask; g.relu = a.relu; g.trans_out = a.trans_out;
    g.c_rowoff = a.c_rowoff; g.c_coloff = a.c_coloff;
    g.workspace = c.ws; g.counters = c.tickets; g.error = device_error_flag();
    g.trace = env_i("BB_TMA_TRACE", 0) ? tma_trace_buffer() : nullptr;
    dim3 grid(tn, tm, split);
    // one CTA per SM anyway (small or split-K grids): give it the 6-stage ring -- with 3 stages a k-slice's round trip
    // (TMA 650 + lo split 380 + MMA 700 cycles) bounds the loop at ~750 cycles per slice, MMA issue alone is ~490
    static const int cfg_env = env_i("BB_TMA_CFG", -1);
    const int cfg = cfg_env >= 0 ? cfg_env : (tiles * split <= (long)c.sms ? 1 : 0);
    // debugging aid: BB_TMA_MASK bit i enables operand combination i (dense forward, conv forward / data gradient,
    // linear data gradient, linear weight gradient, conv weight gradient, conv data gradient); the others fall back to tc_gemm.cu
    const int combo = (AK == OP_K_TILED && BKIND == OP_K_TILED) ? 0 : (AK == OP_K_IM2COL && BKIND == OP_K_TILED) ? (g.ga.flip_w >= 0 ? 5 : 1)
                    : (AK == OP_K_TILED && BKIND == OP_MN_TILED) ? 2 : (AK == OP_MN_TILED && BKIND == OP_MN_TILED) ? 3
                    : (AK == OP_MN_IM2COL && BKIND == OP_MN_TILED) ? 4 : 6;
    if (!((env_i("BB_TMA_MASK", 0x3f) >> combo) & 1)) return false;

    if (passes == 3) {
        if (!launch_combo<3>(AK, BKIND, BN, cfg, ta, tb, g, grid, c.stream)) return false;
    } else if (!tma_launch_fast(AK, BKIND, BN, cfg, ta, tb, g, grid, c.stream)) {
        return false;
    }
    g_tma_launches++;
    c.mark(BN == 32 ? "tma_gemm128x32" : "tma_gemm128x64");
    return true;
}

}  // namespace bb

// clock64 stamps of CTA (0,0,0) of the last traced launch (BB_TMA_TRACE=1): [3 roles][64 slices][4]
extern "C" int32_t bb_tma_trace(int64_t* out) {
    BB_API_BEGIN
    BB_CUDA(cudaDeviceSynchronize());
    BB_CUDA(cudaMemcpy(out, bb::tma_trace_buffer(), 3 * 64 * 4 * sizeof(long long), cudaMemcpyDeviceToHost));
    BB_API_END
}
// ... and the per-CTA part: out[1024][4] = {entry ns, after the dependency wait, exit ns, SM id} (zero rows: CTA did not exist)
extern "C" int32_t bb_tma_trace_ctas(int64_t* out) {
    BB_API_BEGIN
    BB_CHECK(out, "null argument");
    BB_CUDA(cudaDeviceSynchronize());
    BB_CUDA(cudaMemcpy(out, bb::tma_trace_buffer() + 3 * 64 * 4, 4 * 1024 * sizeof(long long), cudaMemcpyDeviceToHost));
    BB_CUDA(cudaMemset(bb::tma_trace_buffer() + 3 * 64 * 4, 0, 4 * 1024 * sizeof(long long)));
    BB_API_END
}

// launches that took the TMA path / tensor-map constructions the driver rejected (tests assert the path is live)
extern "C" int32_t bb_tma_stats(uint64_t* launches, uint64_t* rejects, int32_t reset) {
    BB_API_BEGIN
    if (launches) *launches = bb::g_tma_launches.load();
    if (rejects) *rejects = bb::g_tma_rejects.load();
    if (reset) { bb::g_tma_launches.store(0); bb::g_tma_rejects.store(0); }
    BB_API_END
}
