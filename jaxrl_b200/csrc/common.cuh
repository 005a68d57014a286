// Shared host-side helpers for the C-ABI entry points (error reporting, launch checks).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mapox_b200.h"

namespace mpx {

void set_error(const char* fmt, ...);   // abi.cu: thread-local message for mpx_last_error()

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MPX_ERR_CUDA;
  }
  return MPX_OK;
}

#define MPX_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ::mpx::set_error(__VA_ARGS__);    \
      return MPX_ERR_INVALID;           \
    }                                   \
  } while (0)

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

int num_sms();   // abi.cu (cached cudaDevAttrMultiProcessorCount of the current device)

// Bring-up aid: mpx_debug_set_stamps(ptr) hands the fused kernels a device buffer of 64-bit slots; CTA 0 then records
// %globaltimer (ns) at its phase boundaries.  NULL (the default) compiles to one predicated-off branch per stamp.
extern unsigned long long* g_dbg_stamps;
#ifdef __CUDACC__
__device__ __forceinline__ void dbg_stamp(unsigned long long* buf, int slot) {
  if (buf) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    buf[slot] = t;
  }
}
#endif

}  // namespace mpx
