// Geometry, weight-image layout and one-hot code specification shared by the fused observation-encoder kernels
// (obs_fused.cu: rollout step; obs_train.cu: training forward / backward) for the standard grid_cnn encoder
// (reference mapox_trainer/model/observation.py:63-96): 11x11xK uint8 view -> one-hot -> conv 3x3/2 (16) -> gelu ->
// conv 3x3/1 (32) -> gelu -> conv 3x3/1 (H = 128).
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace mpx {

constexpr int OF_IH = 11, OF_IW = 11, OF_P1 = 5, OF_C1 = 16, OF_C2 = 32, OF_H = 128;
constexpr int OF_K3 = 9 * OF_C2;       // 288
constexpr int OF_MAXK = 4;             // observation components per cell
constexpr int OF_MAXJ = 48;            // joint one-hot codes per cell (product of (size_k + 1))

// Weight image (bytes): [W3 core-matrix image 73728][W2 tap images 9216][conv1 joint table fp32][biases bf16 512]
//   W3: element (n, k) at (n/8)*4608 + (k/8)*128 + (n%8)*16 + (k%8)*2        (k = pos*32 + cin)
//   W2: element (tap, n, c) at tap*1024 + (n/8)*256 + (c/8)*128 + (n%8)*16 + (c%8)*2
//   T1: [tap][code][copy 0|1][16] fp32 = sum over components k with digit_k < sizes[k] of W1[tap*C + off_k + digit_k][:]
//       (row pitch 128 B: copy 0 sits in banks 0-15, copy 1 in banks 16-31, so two lookups never collide); the table
//       is packed with the ACTUAL number of joint codes NJ, the region is sized for OF_MAXJ
//   biases: b1[16] b2[32] b3[128] bf16
constexpr int OF_W3_BYTES = 16 * (OF_K3 / 8) * 128;    // 128 n x 576 B
constexpr int OF_W2_BYTES = 9 * 1024;
constexpr int OF_T1_BYTES = 9 * OF_MAXJ * 2 * OF_C1 * 4;   // 55296
constexpr int OF_BIAS_BYTES = 512;
constexpr int OF_IMG_BYTES = OF_W3_BYTES + OF_W2_BYTES + OF_T1_BYTES + OF_BIAS_BYTES;

// One joint code per cell: sum_k min(value_k, sizes[k]) * radix[k]; digit sizes[k] = out of range = all-zero one-hot.
struct ObsCodeSpec {
  int K;                          // components per cell
  int NJ;                         // joint codes per cell = prod(sizes[k] + 1)
  int sizes[OF_MAXK];             // one-hot size per component
  int radix[OF_MAXK];
};

inline int obs_fill_spec(ObsCodeSpec& s, const int* sizes, int K, const char* who) {
  MPX_REQUIRE(sizes && K >= 1 && K <= OF_MAXK, "%s: K=%d components unsupported", who, K);
  s.K = K;
  int nj = 1;
  for (int i = 0; i < OF_MAXK; ++i) {
    s.sizes[i] = i < K ? sizes[i] : 0;
    s.radix[i] = i < K ? nj : 0;
    if (i < K) {
      MPX_REQUIRE(sizes[i] > 0 && sizes[i] < 255, "%s: bad one-hot size %d", who, sizes[i]);
      nj *= sizes[i] + 1;
    }
  }
  MPX_REQUIRE(nj <= OF_MAXJ, "%s: %d joint one-hot codes per cell > %d (use the unfused encoder)", who, nj, OF_MAXJ);
  s.NJ = nj;
  return MPX_OK;
}

// obs_train.cu: the backward image (data-gradient operands of conv2 / conv3) that mpx_obs_fused_prep appends to the
// forward image at offset OF_IMG_BYTES
size_t obs_train_bwd_image_bytes();
#ifdef __CUDACC__
int obs_train_prep_bwd(const __nv_bfloat16* wt2, const __nv_bfloat16* wt3, uint8_t* image_b, cudaStream_t stream);
__device__ __forceinline__ uint64_t umma_smem_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;                                                  // layout type 0: no swizzle
}
__device__ __forceinline__ void of_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
#endif

}  // namespace mpx
