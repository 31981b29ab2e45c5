// Shared plumbing of the C-ABI translation units: error channel, exception guard, device buffers.
#pragma once
#include <string>

#include "../../include/smcb200.h"
#include "engine.cuh"

namespace smcb_abi {

std::string& last_error();   // thread-local message behind smcb200_last_error()

inline int fail(int code, const std::string& msg) { last_error() = msg; return code; }

template <class F>
int guarded(F&& f) {
    try {
        return f();
    } catch (const smcb::CudaError& e) {
        return fail(SMCB200_ERR_CUDA, e.what());
    } catch (const std::invalid_argument& e) {
        return fail(SMCB200_ERR_INVALID, e.what());
    } catch (const std::exception& e) {
        return fail(SMCB200_ERR_STATE, e.what());
    }
}

inline void require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n < 1)
        throw smcb::CudaError(std::string("no CUDA device available (this library has no CPU fallback): ") +
                              (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
}

struct DevBuf {
    void* p = nullptr;
    DevBuf() {}
    explicit DevBuf(size_t bytes) { SMCB_CUDA(cudaMalloc(&p, bytes ? bytes : 1)); }
    ~DevBuf() { if (p) cudaFree(p); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p) { o.p = nullptr; }
    DevBuf& operator=(DevBuf&& o) noexcept { if (p) cudaFree(p); p = o.p; o.p = nullptr; return *this; }
    template <class T> T* as() const { return (T*)p; }
};

}  // namespace smcb_abi
