// Host-side helpers shared by the C-ABI entry points: error reporting and TMA tensor-map encoding.
// The driver API is resolved lazily through cudaGetDriverEntryPoint so the library loads (and its
// symbols can be inspected) on a machine without libcuda.so.1.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

namespace scm {

// error codes of the C ABI (see include/scmgan.h)
enum : int { SCM_OK = 0, SCM_EINVAL = -1, SCM_EUNSUPPORTED = -2, SCM_ECUDA = -3 };

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define SCM_CUDA(expr)                                                   \
    do {                                                                 \
        cudaError_t _e = (expr);                                         \
        if (_e != cudaSuccess) return ::scm::cuda_fail(_e, #expr);       \
    } while (0)

#define SCM_REQUIRE(cond, ...)                                           \
    do {                                                                 \
        if (!(cond)) {                                                   \
            ::scm::set_error(__VA_ARGS__);                               \
            return ::scm::SCM_EINVAL;                                    \
        }                                                                \
    } while (0)

// rank<=4 tiled bf16 tensor map.  dims/box are innermost-first; strides_bytes has rank-1 entries
// (stride of dim1.. in bytes, multiples of 16).  swizzle_bytes in {0, 32, 64, 128}.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

int num_sms();

}  // namespace scm
