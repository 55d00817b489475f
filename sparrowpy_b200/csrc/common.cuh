// Shared helpers for the sparrow_b200 CUDA translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/sparrow_b200.h"

namespace spb {

// thread-local message of the last failed call (returned by spb_last_error)
std::string &last_error();

inline int fail(int code, const char *what, const char *detail = "") {
    last_error() = std::string(what) + (detail[0] ? ": " : "") + detail;
    return code;
}

inline int check_launch(const char *kernel) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(-2, kernel, cudaGetErrorString(e));
    return 0;
}

#define SPB_CUDA(call)                                                       \
    do {                                                                     \
        cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) return spb::fail(-2, #call, cudaGetErrorString(e__)); \
    } while (0)

#define SPB_REQUIRE(cond, msg)                                               \
    do {                                                                     \
        if (!(cond)) return spb::fail(-1, "invalid argument", msg);          \
    } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

}  // namespace spb
