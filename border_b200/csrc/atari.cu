// atari.cu -- the Atari observation pipeline of border-atari-env on the device.
//
//   skip_and_max's max-pool   border-atari-env/src/env.rs:139-145   element-wise max of two RGB24 frames
//   warp_and_grayscale        env.rs:161-185   image::imageops::resize(84, 84, Triangle) (image 0.23.14: a vertical and a
//                                              horizontal pass, each rounded back to u8) + the 0.299/0.587/0.114 grey
//   stack_frame               env.rs:187-199   frames[1..4] = frames[0..3], frames[0] = new  (newest first)
//   reset                     env.rs:288-297   all four frames = the first warped render
//   clip_reward               env.rs:149-159   sign(r) in training
//
// One launch per env step: every CTA owns a band of output rows, runs the vertical pass of its rows into shared memory
// (reading max(a, b) of the two uploaded frames), the horizontal pass + grey from there, and writes its rows of the new
// frame stack -- the three older frames are copied from the previous stack (two stacks, ping-pong), so the [4][84][84]
// observation the policy / replay push consume stays in HBM and never returns to the host.
// Arithmetic: f32 multiply and add kept separate (__fmul_rn / __fadd_rn), the weight tables are computed on the host in
// f32 exactly as sample.rs does: bit-identical to oracle/atari_oracle.py (tests/test_atari_gpu.py).
#include <math.h>
#include <vector>
#include "common.cuh"
#include "../../include/border_b200.h"

namespace bb {

constexpr int AT_OUT = 84;
constexpr int AT_MAXT = 16;    // taps per output sample (ratio <= 7)
constexpr int AT_ROWS = 4;     // output rows per CTA (84 = 21 x 4)
constexpr int AT_MAXW = 320;   // source width bound (shared-memory band: AT_ROWS x W x 3 bytes)

struct AtTaps {
    int left[AT_OUT];
    int n[AT_OUT];
    float w[AT_OUT][AT_MAXT];
};

// imageops/sample.rs (image 0.23.14), one axis: see oracle/atari_oracle.py:_taps
static bool make_taps(int n_in, AtTaps& t) {
    const float ratio = (float)n_in / (float)AT_OUT;
    const float sratio = ratio < 1.0f ? 1.0f : ratio;
    const float support = 1.0f * sratio;
    for (int o = 0; o < AT_OUT; ++o) {
        const float centre = ((float)o + 0.5f) * ratio;
        long left = (long)floorf(centre - support);
        left = std::min<long>(std::max<long>(left, 0), n_in - 1);
        long right = (long)ceilf(centre + support);
        right = std::min<long>(std::max<long>(right, left + 1), n_in);
        if (right - left > AT_MAXT) return false;
        const float c = centre - 0.5f;
        float sum = 0.0f;
        for (long i = left; i < right; ++i) {
            const float x = ((float)i - c) / sratio;
            const float w = fabsf(x) < 1.0f ? 1.0f - fabsf(x) : 0.0f;
            t.w[o][i - left] = w;
            sum += w;
        }
        for (long i = left; i < right; ++i) t.w[o][i - left] /= sum;
        t.left[o] = (int)left;
        t.n[o] = (int)(right - left);
    }
    return true;
}

__device__ __forceinline__ unsigned char at_round_u8(float t) {
    t = fminf(fmaxf(t, 0.0f), 255.0f);
    return (unsigned char)floorf(__fadd_rn(t, 0.5f));   // f32::round on a non-negative value
}

__global__ void __launch_bounds__(512) atari_step_kernel(const unsigned char* __restrict__ fa, const unsigned char* __restrict__ fb,
                                                         int W, int H, const AtTaps* __restrict__ tv, const AtTaps* __restrict__ th,
                                                         const unsigned char* __restrict__ stack_in, unsigned char* __restrict__ stack_out,
                                                         int fill_all) {
    extern __shared__ unsigned char band[];   // [AT_ROWS][W][3] after the vertical pass
    const int r0 = blockIdx.x * AT_ROWS;
    const int W3 = W * 3;
    // vertical pass (rows r0 .. r0+3): out[r][x][c] = round(sum_i max(a, b)[left + i][x][c] * w_i)
    for (int e = threadIdx.x; e < AT_ROWS * W3; e += blockDim.x) {
        const int r = e / W3, xc = e - r * W3;
        const int o = r0 + r;
        const int left = tv->left[o], n = tv->n[o];
        float t = 0.0f;
        for (int i = 0; i < n; ++i) {
            const size_t src = (size_t)(left + i) * W3 + xc;
            const unsigned char a = fa[src], b = fb[src];
            t = __fadd_rn(t, __fmul_rn((float)(a > b ? a : b), tv->w[o][i]));
        }
        band[e] = at_round_u8(t);
    }
    __syncthreads();
    // horizontal pass + grey, one thread per output pixel of the band
    for (int e = threadIdx.x; e < AT_ROWS * AT_OUT; e += blockDim.x) {
        const int r = e / AT_OUT, x = e - r * AT_OUT;
        const int left = th->left[x], n = th->n[x];
        float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f;
        for (int i = 0; i < n; ++i) {
            const unsigned char* p = band + (size_t)r * W3 + (left + i) * 3;
            const float w = th->w[x][i];
            t0 = __fadd_rn(t0, __fmul_rn((float)p[0], w));
            t1 = __fadd_rn(t1, __fmul_rn((float)p[1], w));
            t2 = __fadd_rn(t2, __fmul_rn((float)p[2], w));
        }
        const float c0 = (float)at_round_u8(t0), c1 = (float)at_round_u8(t1), c2 = (float)at_round_u8(t2);
        // ((0.299 * r) + (0.587 * g) + (0.114 * b)) as u8 with (b, g, r) = bytes 0, 1, 2 of the pixel (env.rs:168-176)
        float g = __fadd_rn(__fmul_rn(0.299f, c2), __fmul_rn(0.587f, c1));
        g = __fadd_rn(g, __fmul_rn(0.114f, c0));
        const unsigned char grey = (unsigned char)fminf(fmaxf(truncf(g), 0.0f), 255.0f);
        const int o = (r0 + r) * AT_OUT + x;
        stack_out[o] = grey;
        for (int k = 1; k < 4; ++k)
            stack_out[k * AT_OUT * AT_OUT + o] = fill_all ? grey : stack_in[(k - 1) * AT_OUT * AT_OUT + o];
    }
}

struct Atari {
    int device = 0, W = 160, H = 210, train = 1;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    unsigned char* d_frames = nullptr;   // two RGB24 frames
    unsigned char* h_frames = nullptr;   // pinned staging
    unsigned char* d_stack[2] = {nullptr, nullptr};
    int cur = 0;                         // which stack holds the current observation
    AtTaps *d_tv = nullptr, *d_th = nullptr;
    cudaEvent_t ev = nullptr;
    size_t frame_bytes() const { return (size_t)W * H * 3; }

    void launch(const unsigned char* a_host, const unsigned char* b_host, bool fill_all) {
        DeviceGuard g(device);
        const size_t fb = frame_bytes();
        // (the previous call's copy has been consumed: the stream was synchronised or an event waited on since)
        BB_CUDA(cudaEventSynchronize(ev));
        memcpy(h_frames, a_host, fb);
        memcpy(h_frames + fb, b_host, fb);
        BB_CUDA(cudaMemcpyAsync(d_frames, h_frames, 2 * fb, cudaMemcpyHostToDevice, stream));
        const int nxt = cur ^ 1;
        atari_step_kernel<<<AT_OUT / AT_ROWS, 512, (size_t)AT_ROWS * W * 3, stream>>>(d_frames, d_frames + fb, W, H, d_tv, d_th,
                                                                                      d_stack[cur], d_stack[nxt], fill_all ? 1 : 0);
        BB_LAUNCHED();
        BB_CUDA(cudaEventRecord(ev, stream));
        cur = nxt;
    }
};

}  // namespace bb

struct bb_atari { bb::Atari impl; };

extern "C" {

int32_t bb_atari_create(int32_t device, int32_t width, int32_t height, int32_t train, bb_atari** out) {
    BB_API_BEGIN
    BB_CHECK(out, "null argument");
    BB_CHECK(width >= bb::AT_OUT && width <= bb::AT_MAXW && height >= bb::AT_OUT && height <= 4096, "frame size out of range");
    bb::DeviceGuard g(device);
    auto* st = new bb_atari();
    bb::Atari& a = st->impl;
    a.device = device; a.W = width; a.H = height; a.train = train;
    std::vector<bb::AtTaps> t(2);
    BB_CHECK(bb::make_taps(height, t[0]) && bb::make_taps(width, t[1]), "resize ratio too large for the tap table");
    BB_CUDA(cudaStreamCreateWithFlags(&a.stream, cudaStreamNonBlocking));
    a.own_stream = true;
    a.d_frames = bb::dev_alloc<unsigned char>(2 * a.frame_bytes());
    BB_CUDA(cudaMallocHost(&a.h_frames, 2 * a.frame_bytes()));
    for (int k = 0; k < 2; ++k) a.d_stack[k] = bb::dev_alloc_zero<unsigned char>(4 * bb::AT_OUT * bb::AT_OUT + 16, a.stream);
    a.d_tv = bb::dev_alloc<bb::AtTaps>(1);
    a.d_th = bb::dev_alloc<bb::AtTaps>(1);
    bb::h2d_sync(a.d_tv, &t[0], sizeof(bb::AtTaps), a.stream);
    bb::h2d_sync(a.d_th, &t[1], sizeof(bb::AtTaps), a.stream);
    BB_CUDA(cudaEventCreateWithFlags(&a.ev, cudaEventDisableTiming));
    BB_CUDA(cudaEventRecord(a.ev, a.stream));
    BB_CUDA(cudaStreamSynchronize(a.stream));
    *out = st;
    BB_API_END
}
int32_t bb_atari_destroy(bb_atari* st) {
    BB_API_BEGIN
    if (!st) return 0;
    bb::Atari& a = st->impl;
    bb::DeviceGuard g(a.device);
    cudaStreamSynchronize(a.stream);
    cudaFree(a.d_frames); cudaFree(a.d_stack[0]); cudaFree(a.d_stack[1]); cudaFree(a.d_tv); cudaFree(a.d_th);
    if (a.h_frames) cudaFreeHost(a.h_frames);
    if (a.ev) cudaEventDestroy(a.ev);
    if (a.own_stream) cudaStreamDestroy(a.stream);
    delete st;
    BB_API_END
}
int32_t bb_atari_set_stream(bb_atari* st, void* cuda_stream) {
    BB_API_BEGIN
    BB_CHECK(st, "null handle");
    bb::Atari& a = st->impl;
    bb::DeviceGuard g(a.device);
    BB_CUDA(cudaStreamSynchronize(a.stream));
    if (a.own_stream) { cudaStreamDestroy(a.stream); a.own_stream = false; }
    a.stream = (cudaStream_t)cuda_stream;
    BB_CUDA(cudaEventRecord(a.ev, a.stream));
    BB_API_END
}
int32_t bb_atari_reset(bb_atari* st, const uint8_t* rgb) {
    BB_API_BEGIN
    BB_CHECK(st && rgb, "null argument");
    st->impl.launch(rgb, rgb, true);
    BB_API_END
}
int32_t bb_atari_step(bb_atari* st, const uint8_t* rgb_a, const uint8_t* rgb_b, float reward, float* reward_out) {
    BB_API_BEGIN
    BB_CHECK(st && rgb_a && rgb_b, "null argument");
    st->impl.launch(rgb_a, rgb_b, false);
    if (reward_out) *reward_out = st->impl.train ? (reward == 0.0f ? 0.0f : (reward > 0.0f ? 1.0f : -1.0f)) : reward;
    BB_API_END
}
int32_t bb_atari_obs_device(bb_atari* st, void* consumer_stream, const uint8_t** dev_ptr) {
    BB_API_BEGIN
    BB_CHECK(st && dev_ptr, "null argument");
    bb::Atari& a = st->impl;
    bb::DeviceGuard g(a.device);
    if ((cudaStream_t)consumer_stream != a.stream) BB_CUDA(cudaStreamWaitEvent((cudaStream_t)consumer_stream, a.ev, 0));
    *dev_ptr = a.d_stack[a.cur];
    BB_API_END
}
int32_t bb_atari_obs_host(bb_atari* st, uint8_t* out) {
    BB_API_BEGIN
    BB_CHECK(st && out, "null argument");
    bb::Atari& a = st->impl;
    bb::DeviceGuard g(a.device);
    BB_CUDA(cudaMemcpyAsync(out, a.d_stack[a.cur], 4 * bb::AT_OUT * bb::AT_OUT, cudaMemcpyDeviceToHost, a.stream));
    BB_CUDA(cudaStreamSynchronize(a.stream));
    BB_API_END
}

}  // extern "C"
