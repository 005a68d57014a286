// Host-side TMA tensor-map encoding. cuTensorMapEncodeTiled is resolved at run time through
// cudaGetDriverEntryPoint so the library has no link-time dependency on libcuda (the build box has no driver).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpx {

// Row-major 2-D bf16 matrix [rows, cols] with row pitch ld (elements); box = box_cols x box_rows elements,
// 128-byte swizzle (box_cols * 2 bytes must be <= 128).  Out-of-bounds elements are zero-filled.
// Returns 0 on success.
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_cols, uint32_t box_rows);
// The same with the 64-byte swizzle (box_cols * 2 bytes must be <= 64).
int make_tmap_bf16_2d_sw64(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                           uint32_t box_cols, uint32_t box_rows);

}  // namespace mpx
