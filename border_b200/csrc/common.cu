// common.cu -- process-wide state of the library: thread-local error string, launch counter.
#include "common.cuh"
#include <map>
#include <mutex>
#include <string.h>
#include "../../include/border_b200.h"

namespace bb {
std::string& last_error() {
    static thread_local std::string e;
    return e;
}
void set_error(const std::string& msg) { last_error() = msg; }
std::atomic<uint64_t> g_launch_count{0};

int* device_error_flag() {
    static int* flag = nullptr;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!flag) {
        BB_CUDA(cudaHostAlloc(&flag, 64, cudaHostAllocMapped | cudaHostAllocPortable));
        memset(flag, 0, 64);
    }
    return flag;
}
void check_device_error(const char* where) {
    volatile int* f = device_error_flag();
    if (*f != 0) {
        char b[256];
        snprintf(b, sizeof(b), "%s: a device-side bounded wait timed out (code %d): results of this handle are invalid", where, *f);
        throw Error(b);
    }
}

// ------------------------------------------------------------------------------- streams

// One non-blocking stream per (host thread, device): handles created on one thread (a learner's agent + replay)
// share it, which is what lets the update be captured into a CUDA graph; handles created on other threads (the
// actors of the async trainer build their own agents, actor/base.rs:127-136) get their own, so their calls never
// land on a stream another thread is capturing.
cudaStream_t device_stream(int device) {
    static thread_local std::map<int, cudaStream_t> streams;
    auto it = streams.find(device);
    if (it != streams.end()) return it->second;
    DeviceGuard g(device);
    cudaStream_t s;
    BB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    streams[device] = s;
    return s;
}

void stream_wait(cudaStream_t waiter, cudaStream_t signaler) {
    cudaEvent_t e;
    BB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    BB_CUDA(cudaEventRecord(e, signaler));
    BB_CUDA(cudaStreamWaitEvent(waiter, e, 0));
    BB_CUDA(cudaEventDestroy(e));  // released once the wait has been satisfied
}

}  // namespace bb

extern "C" {
const char* bb_last_error(void) { return bb::last_error().c_str(); }
int32_t bb_abi_version(void) { return BB_ABI_VERSION; }
int32_t bb_device_count(int32_t* out) {
    BB_API_BEGIN
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { n = 0; cudaGetLastError(); }
    *out = n;
    BB_API_END
}
int32_t bb_device_error(int32_t* out, int32_t reset) {
    BB_API_BEGIN
    volatile int* f = bb::device_error_flag();
    if (out) *out = *f;
    if (reset) *f = 0;
    BB_API_END
}
int32_t bb_kernel_launch_count(uint64_t* out, int32_t reset) {
    BB_API_BEGIN
    if (out) *out = bb::g_launch_count.load();
    if (reset) bb::g_launch_count.store(0);
    BB_API_END
}
}
