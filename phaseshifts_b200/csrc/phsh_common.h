// phsh_common.h -- host-side helpers shared by the translation units of libphsh_b200.so
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/phsh_b200.h"

namespace phsh {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define PHSH_CUDA(call)                                              \
    do {                                                             \
        cudaError_t _e = (call);                                     \
        if (_e != cudaSuccess) return ::phsh::cuda_fail(_e, #call);  \
    } while (0)

// make `device` current and check that it is a Blackwell sm_100 part; returns 0 or a PHSH_B200_E* code
int select_device(int device);
int check_current_device();

// stream-ordered scratch that is retained by the default mempool between calls
struct DeviceBuf {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    ~DeviceBuf() {
        if (p) cudaFreeAsync(p, s);
    }
    int alloc(size_t bytes, cudaStream_t stream);
};

}  // namespace phsh
