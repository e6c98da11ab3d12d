// common.cuh -- error handling, launch accounting and small device helpers shared by the library.
#pragma once
#include <stdlib.h>
#include <utility>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <stdexcept>
#include <string>

namespace bb {

// thread-local message behind bb_last_error()
std::string& last_error();
void set_error(const std::string& msg);

extern std::atomic<uint64_t> g_launch_count;  // kernels launched by this library

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define BB_CUDA(expr)                                                                           \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            char _b[512];                                                                       \
            snprintf(_b, sizeof(_b), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,             \
                     cudaGetErrorString(_e));                                                   \
            throw bb::Error(_b);                                                                \
        }                                                                                       \
    } while (0)

#define BB_CHECK(cond, msg)                                                                     \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            char _b[512];                                                                       \
            snprintf(_b, sizeof(_b), "%s:%d: %s", __FILE__, __LINE__, (msg));                  \
            throw bb::Error(_b);                                                                \
        }                                                                                       \
    } while (0)

// Counts the launch and surfaces launch-configuration errors immediately.
#define BB_LAUNCHED()                                                                           \
    do {                                                                                        \
        bb::g_launch_count.fetch_add(1, std::memory_order_relaxed);                             \
        BB_CUDA(cudaGetLastError());                                                            \
    } while (0)

// Wraps the body of every extern "C" entry point: nothing unwinds across the ABI.
#define BB_API_BEGIN try {
#define BB_API_END                                                                              \
    return 0;                                                                                   \
    }                                                                                           \
    catch (const std::exception& e) {                                                           \
        bb::set_error(e.what());                                                                \
        return 1;                                                                               \
    }                                                                                           \
    catch (...) {                                                                               \
        bb::set_error("unknown C++ exception");                                                 \
        return 2;                                                                               \
    }

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <class T>
T* dev_alloc(size_t n) {
    T* p = nullptr;
    BB_CUDA(cudaMalloc(&p, (n ? n : 1) * sizeof(T)));
    return p;
}
template <class T>
T* dev_alloc_zero(size_t n, cudaStream_t s = 0) {
    T* p = dev_alloc<T>(n);
    BB_CUDA(cudaMemsetAsync(p, 0, (n ? n : 1) * sizeof(T), s));
    return p;
}

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// The update is a chain of ~20 dependent kernels of 5-50 us each; launched with the programmatic-stream-
// serialization attribute, kernel N+1 is scheduled as soon as every CTA of kernel N has passed pdl_sync(), and
// its CTAs then block in griddepcontrol.wait until kernel N has completed and flushed: the launch latency and
// block scheduling of N+1 overlap N's execution.  Every kernel launched through launch_pdl() calls pdl_sync()
// before it touches global memory.  Captured into the CUDA graph as programmatic edges.  BB_PDL=0 disables.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif
// Measured on B200 (tools/quick_bench.py): back-to-back replay sample+gather calls 10.3 -> 8.4 us with PDL, but the
// graph-replayed DQN update 428 -> 433 us (graph nodes already launch back to back, and early-resident dependents take
// SM slots from the side-stream branches), so BB_PDL: unset = replay kernels only, 1 = every converted kernel, 0 = none.
inline int pdl_mode() {
    static const int m = getenv("BB_PDL") ? atoi(getenv("BB_PDL")) : -1;
    return m;
}
inline bool pdl_enabled() { return pdl_mode() == 1; }          // agent / GEMM kernels
inline bool pdl_replay_enabled() { return pdl_mode() != 0; }   // replay kernels
template <typename... KArgs, typename... Args>
inline void launch_pdl_if(bool allow, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = allow ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    BB_CUDA(cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(std::forward<Args>(args))...));
}

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    launch_pdl_if(pdl_enabled(), kern, grid, block, smem, s, std::forward<Args>(args)...);
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, DEVICE): the attribute is per device, so a
// process-wide "configured" flag left every GPU but the first one a handle touched with the 48 KB default.
#define BB_ENSURE_SMEM(kern, bytes)                                                                        \
    do {                                                                                                   \
        static std::atomic<uint32_t> _done{0};                                                             \
        int _dev = 0;                                                                                      \
        cudaGetDevice(&_dev);                                                                              \
        if (!(_done.load(std::memory_order_acquire) & (1u << (_dev & 31)))) {                              \
            BB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            _done.fetch_or(1u << (_dev & 31), std::memory_order_release);                                  \
        }                                                                                                  \
    } while (0)

// Sticky device-side failure flag (a bounded mbarrier / peer-barrier wait timed out): ONE int in pinned, mapped host
// memory, so kernels store to it directly and every host entry point can test it without a copy or a sync.
int* device_error_flag();
void check_device_error(const char* where);  // throws bb::Error when the flag is set

cudaStream_t device_stream(int device);  // one non-blocking stream per (host thread, device)
// Host -> device copy that later work on `s` (a NON-BLOCKING stream) is guaranteed to see.  A plain cudaMemcpy from
// pageable memory returns once the data sits in the driver's staging buffer: the DMA itself is ordered on the legacy
// default stream, which non-blocking streams do not wait for -- a kernel launched right after (the lo-plane refresh of
// set_param) read the OLD parameter values.  Copy in stream order and wait for the stream.
inline void h2d_sync(void* dst, const void* src, size_t bytes, cudaStream_t s) {
    BB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
    BB_CUDA(cudaStreamSynchronize(s));
}
void stream_wait(cudaStream_t waiter, cudaStream_t signaler);

inline int num_sms(int device) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
    return n > 0 ? n : 148;
}

}  // namespace bb
