// Shared between the C-ABI translation units (orb_capi.cu, search_capi.cu): error reporting and the
// matcher handle.  Not installed; include/orb_b200.h is the only public header.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "../../include/orb_b200.h"

int orb_fail(int code, const char* fmt, ...);
#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) return orb_fail(ORB_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

struct orb_matcher {
    int device = 0;
    cudaStream_t stream = nullptr;
    // grow-only device scratch for the host-buffer entry points
    static const int kSlots = 16;
    void* buf[kSlots] = {};
    size_t cap[kSlots] = {};
};

// Grow-only device scratch slot of at least `bytes`.
int orb_matcher_scratch(orb_matcher* m, int slot, size_t bytes, void** out);
