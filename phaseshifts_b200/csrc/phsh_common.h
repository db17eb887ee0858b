// phsh_common.h -- host-side helpers shared by the translation units of libphsh_b200.so
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/phsh_b200.h"

namespace phsh {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define PHSH_CUDA(call)                                              \
    do {                                                             \
        cudaError_t _e = (call);                                     \
        if (_e != cudaSuccess) return ::phsh::cuda_fail(_e, #call);  \
    } while (0)

// Stage timing of the file entry points: with PHSH_B200_TRACE=1 in the environment every stage boundary prints one line
// "[phsh_b200] <entry>: <stage> <ms>" to stderr (wall clock; the device stages include their synchronisation).
struct Trace {
    const char* entry;
    bool on;
    double t0, last;
    explicit Trace(const char* entry_name);
    void mark(const char* stage);
    ~Trace();
};

// make `device` current and check that it is a Blackwell sm_100 part; returns 0 or a PHSH_B200_E* code
int select_device(int device);
int check_current_device();

// stream-ordered scratch from the library's PRIVATE per-device memory pool (never the device's default pool, which
// the embedding process -- e.g. torch's allocator -- may be sharing).  The pool keeps up to PHSH_B200_POOL_KEEP_MB
// (default 4096) cached between calls; phsh_b200_trim() hands everything back to the driver.
// With PHSH_B200_GUARD=1 in the environment every buffer gets a 256-byte canary band in front of and behind it, checked
// when the buffer is released (compute-sanitizer is not available on the target pool; this is the library's own
// out-of-bounds-write detector for its scratch and staging buffers, see phsh_b200_guard_violations).
struct DeviceBuf {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    size_t guarded_bytes = 0;  // > 0: p sits kGuardBand bytes inside a larger allocation
    static constexpr size_t kGuardBand = 256;
    ~DeviceBuf() { release(); }
    int alloc(size_t bytes, cudaStream_t stream);
    void release();
};

// What a "_host" / "_file" entry point needs besides device memory: streams, events and (for pageable host buffers)
// a pinned staging ring.  Created once per device and recycled between calls; concurrent host threads each get a
// lane of their own, so the entry points stay re-entrant.
struct HostLane {
    static constexpr int kSlots = 4;
    static constexpr size_t kSlotBytes = 8u << 20;
    int dev = -1;
    cudaStream_t sc = nullptr, sc2 = nullptr, sd = nullptr;  // two compute streams + one copy-back stream
    std::vector<cudaEvent_t> ev;                              // recycled, timing disabled
    size_t ev_next = 0;
    void* stage[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t stage_ev[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    int event(cudaEvent_t* out);  // next recycled event of this call
    int ensure_stage();           // allocates the pinned ring on first use
    // dst (HOST, pinned or pageable) <- src (DEVICE), ordered after everything already enqueued on `sd`.  Pinned
    // destinations: one asynchronous copy.  Pageable destinations: chunks through the pinned ring with the host
    // memcpy of chunk k overlapping the DMA of chunk k+1; returns when the bytes are in dst.
    int copy_back(void* dst, const void* src, size_t bytes, bool dst_pinned);
};
bool host_pointer_is_pinned(const void* p);

// RAII lease of a lane of the CURRENT device (call select_device first).  The destructor drains the lane's three
// streams before the lane is handed back, so an early error return can never leave work in flight on buffers that
// are about to be released: declare the DeviceBufs of a multi-stream entry point BEFORE its lease.
struct StreamGuard {
    cudaStream_t s = nullptr;  // = lane->sc
    HostLane* lane = nullptr;
    int acquire();
    ~StreamGuard();
};

}  // namespace phsh
