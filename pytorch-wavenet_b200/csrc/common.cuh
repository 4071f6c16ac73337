// common.cuh -- error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <cstdint>
#include "../../include/wavenet_b200.h"

namespace wn {

// last error text, per host thread (returned by wn_last_error_string)
char* err_buf();
int   set_err(int code, const char* fmt, ...);

#define WN_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return ::wn::set_err((int)e_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                                 __FILE__, __LINE__);                                         \
    } while (0)

#define WN_REQUIRE(cond, code, ...)                                                           \
    do {                                                                                      \
        if (!(cond)) return ::wn::set_err((code), __VA_ARGS__);                               \
    } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace wn
