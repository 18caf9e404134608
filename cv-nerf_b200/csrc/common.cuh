// Shared helpers for the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "nerf_b200.h"

namespace nerf {

void set_last_error(const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

inline int arg_error(const char* what) {
    set_last_error("bad argument: %s", what);
    return NERF_ERR_ARG;
}

// Every entry point launches on the caller's stream; the CUDA *current device* must be the one that
// owns the buffers (grid sizing reads its SM count, and a launch on another device's stream fails).
// The guard looks the device up from an output pointer and switches to it for the duration of the
// call -- a caller holding tensors on cuda:1 while cuda:0 is current gets the right launch.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(const void* device_ptr) {
        if (!device_ptr) return;
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, device_ptr) != cudaSuccess) { cudaGetLastError(); return; }
        if (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged) return;
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        if (attr.device != prev && cudaSetDevice(attr.device) == cudaSuccess) switched = true;
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

inline unsigned blocks_for(long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace nerf
