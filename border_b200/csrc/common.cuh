// common.cuh -- error handling, launch accounting and small device helpers shared by the library.
#pragma once
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

cudaStream_t device_stream(int device);  // one non-blocking stream per (host thread, device)
void stream_wait(cudaStream_t waiter, cudaStream_t signaler);

inline int num_sms(int device) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
    return n > 0 ? n : 148;
}

}  // namespace bb
