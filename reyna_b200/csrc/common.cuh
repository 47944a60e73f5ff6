// Shared host/device helpers for the reyna_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/reyna_b200.h"

namespace rb {

void set_error(const char* fmt, ...);
void count_launch();   // every kernel launch of the library is counted (reyna_b200_launch_count)

inline int check_cuda(cudaError_t e, const char* what) {
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

#define RB_CUDA(call)                                     \
    do {                                                  \
        int _rc = rb::check_cuda((call), #call);          \
        if (_rc) return _rc;                              \
    } while (0)

#define RB_LAUNCH_CHECK(name)                             \
    do {                                                  \
        rb::count_launch();                               \
        int _rc = rb::check_cuda(cudaGetLastError(), name); \
        if (_rc) return _rc;                              \
    } while (0)

#define RB_REQUIRE(cond, code, msg)                       \
    do {                                                  \
        if (!(cond)) {                                    \
            rb::set_error("%s", msg);                     \
            return (code);                                \
        }                                                 \
    } while (0)

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs; grids are sized in multiples of this

__host__ __device__ inline int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grid size for a grid-stride kernel: enough blocks to cover n items, capped at `cap`
inline int grid_for(int64_t n, int per_block, int cap) {
    int64_t b = div_up(n > 1 ? n : 1, per_block);
    return (int)(b < cap ? b : cap);
}

// Exclusive scan of n int32 values (in != out allowed to alias); total written to *total_dev when non-null.
// scratch must hold scan_scratch_ints(n) ints.
size_t scan_scratch_ints(int64_t n);
int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* scratch, int32_t* total_dev,
                       cudaStream_t st);

}  // namespace rb
