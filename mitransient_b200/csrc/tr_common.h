// Error plumbing shared by the C-ABI translation units.  No exceptions cross the C boundary:
// every entry point returns a TrStatus and records a message retrievable with tr_last_error().
#pragma once
#include <cuda_runtime.h>

#include <new>
#include <string>

#include "../../include/mitransient_b200.h"

namespace tr {
int fail(int code, const std::string &msg);   // records msg (thread-local) and returns code
int fail_cuda(cudaError_t e, const char *what, const char *file, int line);
}  // namespace tr

#define TR_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) return tr::fail_cuda(_e, #expr, __FILE__, __LINE__);        \
    } while (0)

#define TR_CUDA_OR(expr, cleanup)                                                          \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) { cleanup; return tr::fail_cuda(_e, #expr, __FILE__, __LINE__); } \
    } while (0)
