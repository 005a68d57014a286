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

// Launch with optional thread-block cluster (x dimension) and programmatic dependent launch (see ptx.cuh).
extern int g_use_pdl;   // mpx_debug_set(6, mask); default 12.  Bit per kernel family (the `pdl` argument of launch_ex):
                        // 1 = everything else, 2 = obs encoder, 4 = decode attention (+ rollout-sized GEMM / rmsnorm launches, M <= 8192),
                        // 8 = fused GLU, 16 = policy head.
                        // Measured on the whole PPO step (trips 76 / 81, 10 timed steps, +-0.03 ms run to run): 4 -> -1.4 ms
                        // of rollout, 4|8 -> -2.0 ms, 2 -> +3.5 ms, 16 -> +3.1 ms (their early CTAs take the SM slots the
                        // predecessor's tail needs).
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                             int cluster_x, int pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 0) {   // 0 = not a cluster kernel; >= 1 = launch as a cluster (the kernel uses shared::cluster addressing)
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl & g_use_pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

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
