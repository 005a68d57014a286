// Persistent TMA + tcgen05 GEMM family for the policy's dense contractions (bf16 x bf16 -> fp32 in TMEM).
//
//   forward / dgrad :  C[M,N] = A[M,K] . Bt[N,K]^T          both operands K-major, SWIZZLE_128B
//   wgrad           :  dW[K,N] += X[M,K]^T . dY[M,N]        both operands MN-major (read straight from the
//                                                           row-major activations), SWIZZLE_128B, split over M
//
// Structure (one CTA per SM, 320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc),
// warps 2-9 = epilogue (TMEM -> registers -> fused epilogue -> global; two warps per TMEM lane quarter, each
// taking every other 32-column chunk, so math-heavy epilogues are not limited to one warp per SM sub-partition).  4-stage smem ring (full/empty
// mbarriers) and 2 TMEM accumulator stages (tmem_full/tmem_empty) so the epilogue of tile i overlaps the MMAs
// of tile i+1.  Tile = 128 x BLOCK_N, BLOCK_K = 64 bf16 = one 128-byte swizzle atom.
//
// Epilogues (what the reference does as separate XLA fusions):
//   EPI_STORE : bf16 round, + bias, gelu, + residual (contract #1, #6)           -> nnx.Linear / LinearGeneral
//   EPI_QKV   : bf16 round, RoPE on q/k heads (positional_embeddings.py:14-27), KV ring-slot write
//               (attention.py:108-114)
//   EPI_GLU   : h = bf16(bf16(gelu(u)) * g) (feed_forward.py:73-77), optional save of u, g for backward
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.h"

namespace mpx {

int g_debug_swap_lbo_sbo = 0;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int STAGES = 4;
constexpr int WGRAD_THREADS = 192;  // weight-gradient kernel: warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int GEMM_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB

enum { EPI_STORE = 0, EPI_QKV = 1, EPI_GLU = 2, EPI_PAIRSUM = 3 };

struct GemmParams {
  int M, N, K;
  // EPI_STORE
  const __nv_bfloat16* bias;
  const __nv_bfloat16* residual;
  long long ldr;
  __nv_bfloat16* y;
  long long ldy;
  float* y_f32;
  long long ldy32;
  int act;
  __nv_bfloat16* y_pre;      // optional: bf16 pre-activation (after bias), same pitch as y
  const float* bias_f32;     // EPI_PAIRSUM
  const float* norm_scale;   // EPI_QKV, K == 128: RMSNorm(x; scale, eps) applied to the A tile in shared memory first
  float norm_eps;
  int tma_store;             // EPI_STORE: bf16 outputs (y, y_pre) leave through per-warp TMA bulk stores (tm_y, tm_ypre)
  // EPI_QKV
  const int32_t* pos;
  int pos_len;
  const float* timescale;
  __nv_bfloat16* q;
  __nv_bfloat16* k;
  __nv_bfloat16* v;
  long long q_row_stride, k_row_stride, v_row_stride;
  int cache_S, Nq, Nkv;
  // EPI_GLU
  __nv_bfloat16* h;
  __nv_bfloat16* u_save;
  __nv_bfloat16* g_save;
  int F;
};

__host__ __device__ constexpr uint32_t tmem_cols_for(int block_n) {
  return 2 * block_n <= 32 ? 32 : 2 * block_n <= 64 ? 64 : 2 * block_n <= 128 ? 128 : 2 * block_n <= 256 ? 256 : 512;
}

template <int BLOCK_N>
struct SmemLayout {
  static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int NORM_OFFSET = BAR_OFFSET + 256;    // 2 x 128 partial sums of squares + 128 scales (pre-norm prologue)
  static constexpr int XPOSE_OFFSET = BAR_OFFSET + 2048;  // 2 x 2 KB per epilogue warp: row-per-lane <-> coalesced transposes,
                                                          // and the (SWIZZLE_64B) source tiles of the TMA output stores
  static constexpr int TOTAL = XPOSE_OFFSET + 8 * 4096 + 1024;    // + 1 KB alignment slack
};

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&f)[32]) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 v;
    v.x = pack_bf16(f[8 * i + 0], f[8 * i + 1]);
    v.y = pack_bf16(f[8 * i + 2], f[8 * i + 3]);
    v.z = pack_bf16(f[8 * i + 4], f[8 * i + 5]);
    v.w = pack_bf16(f[8 * i + 6], f[8 * i + 7]);
    d4[i] = v;
  }
}
__device__ __forceinline__ void load_bf16x32(const __nv_bfloat16* src, float (&f)[32]) {
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 v = s4[i];
    f[8 * i + 0] = bf16_lo(v.x); f[8 * i + 1] = bf16_hi(v.x);
    f[8 * i + 2] = bf16_lo(v.y); f[8 * i + 3] = bf16_hi(v.y);
    f[8 * i + 4] = bf16_lo(v.z); f[8 * i + 5] = bf16_hi(v.z);
    f[8 * i + 6] = bf16_lo(v.w); f[8 * i + 7] = bf16_hi(v.w);
  }
}


// ---- row-per-lane <-> coalesced re-layout of a 32-row x 32-column bf16 tile through a 2 KB per-warp smem buffer.
// The epilogue holds one output ROW per lane (TMEM lane == row).  Writing that straight to global memory makes
// every warp store touch 32 different 128-byte lines with 16 bytes each, and the L1 tag stage (one line per cycle)
// becomes the bottleneck of epilogue-heavy GEMMs.  After the transpose, lane = (row r8, 16-byte piece j): one warp
// store covers 8 rows x 64 contiguous bytes.  The XOR swizzle keeps both smem phases bank-conflict free.
// tile: global address of (first row of the warp, first column of the chunk); rows_valid in [0, 32];
// xp: shared-space address of this warp's 2 KB transpose buffer; v: this lane's row as 16 packed bf16 pairs.
__device__ __forceinline__ void warp_store_tile(uint32_t xp, __nv_bfloat16* tile, long long ld, int rows_valid,
                                                const uint32_t (&w)[16], int lane) {
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    sts_v4(xp + lane * 64 + ((j ^ sw) << 4), make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]));
  __syncwarp();
  const int r8 = lane >> 2, j = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int R = 8 * i + r8;
    const uint4 x = lds_v4(xp + R * 64 + ((j ^ ((R >> 1) & 3)) << 4));
    if (R < rows_valid) *reinterpret_cast<uint4*>(tile + (long long)R * ld + j * 8) = x;
  }
  __syncwarp();
}
// The same tile through a TMA bulk store: the row-per-lane write above IS the SWIZZLE_64B layout of a 32-column box
// (16-byte chunk index XOR address bits 7-8), so the 4 shared loads + 4 global stores per lane become one instruction
// per warp.  Two buffers per warp; a buffer is rewritten only after the store issued from it two tiles ago has read it.
__device__ __forceinline__ void warp_store_tile_tma(uint32_t xp, uint32_t& sel, const CUtensorMap* tm, int col0, int row0,
                                                    const uint32_t (&w)[16], int lane) {
  const uint32_t buf = xp + sel * 2048;
  sel ^= 1;
  if (lane == 0) bulk_wait_group_read1();
  __syncwarp();
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    sts_v4(buf + lane * 64 + ((j ^ sw) << 4), make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]));
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d_s(tm, buf, col0, row0);
    bulk_commit_group();
  }
}
// coalesced global read of the warp's 32 x 32 tile into registers (lane = (r8, j) layout) ...
__device__ __forceinline__ void warp_fetch_tile(const __nv_bfloat16* tile, long long ld, int rows_valid, uint4 (&c)[4],
                                                int lane) {
  const int r8 = lane >> 2, j = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int R = 8 * i + r8;
    c[i] = (R < rows_valid) ? *reinterpret_cast<const uint4*>(tile + (long long)R * ld + j * 8) : make_uint4(0, 0, 0, 0);
  }
}
// ... and its conversion to one row per lane (16 packed bf16 pairs)
__device__ __forceinline__ void warp_rows_from_fetched(uint32_t xp, const uint4 (&c)[4], uint32_t (&w)[16], int lane) {
  const int r8 = lane >> 2, j = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int R = 8 * i + r8;
    sts_v4(xp + R * 64 + ((j ^ ((R >> 1) & 3)) << 4), c[i]);
  }
  __syncwarp();
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const uint4 v = lds_v4(xp + lane * 64 + ((jj ^ sw) << 4));
    w[4 * jj] = v.x; w[4 * jj + 1] = v.y; w[4 * jj + 2] = v.z; w[4 * jj + 3] = v.w;
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------ epilogues
// Each epilogue thread owns one output row; `acc` holds 32 consecutive fp32 columns starting at `col0`.
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}

// Warp-cooperative: every lane of the warp must call (rows beyond M are masked inside).  row0 = first row of the warp,
// rpre: this chunk's residual already fetched with warp_fetch_tile (when use_pre).
// Values are carried as packed bf16 pairs between the rounding points (one F2FP per pair).
__device__ __forceinline__ void epi_store(const GemmParams& p, int row0, int lane, int col0, const uint32_t (&acc)[32],
                                          uint32_t xp, const uint4 (&rpre)[4], bool use_pre, const CUtensorMap* tm_y,
                                          const CUtensorMap* tm_ypre, uint32_t& sel) {
  if (row0 >= p.M || col0 >= p.N) return;                 // warp-uniform
  const int row = row0 + lane;
  const int rows_valid = min(32, p.M - row0);
  const bool full = (col0 + 32 <= p.N);
  if (p.y_f32 && row < p.M) {
    float* d = p.y_f32 + (long long)row * p.ldy32 + col0;
    if (full && ((p.ldy32 & 3) == 0)) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        reinterpret_cast<float4*>(d)[i] = make_float4(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1]),
                                                      __uint_as_float(acc[4 * i + 2]), __uint_as_float(acc[4 * i + 3]));
    } else {
      for (int i = 0; i < 32 && col0 + i < p.N; ++i) d[i] = __uint_as_float(acc[i]);
    }
  }
  if (!p.y) return;
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = pack_bf16(__uint_as_float(acc[2 * i]), __uint_as_float(acc[2 * i + 1]));
  if (p.bias) {
    if (full && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 bv = reinterpret_cast<const uint4*>(p.bias + col0)[i];
        const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) w[4 * i + e] = bf16x2_add(w[4 * i + e], bw[e]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float b0 = (col0 + 2 * i < p.N) ? __bfloat162float(p.bias[col0 + 2 * i]) : 0.f;
        const float b1 = (col0 + 2 * i + 1 < p.N) ? __bfloat162float(p.bias[col0 + 2 * i + 1]) : 0.f;
        w[i] = pack_bf16(bf16_lo(w[i]) + b0, bf16_hi(w[i]) + b1);
      }
    }
  }
  // 16-byte pieces everywhere: pitches are multiples of 8 elements and the bases 16-byte aligned
  const bool vec_y = full && ((p.ldy & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
  if (p.y_pre) {
    if (p.tma_store) {
      warp_store_tile_tma(xp, sel, tm_ypre, col0, row0, w, lane);
    } else if (vec_y && ((reinterpret_cast<uintptr_t>(p.y_pre) & 15) == 0)) {
      warp_store_tile(xp, p.y_pre + (long long)row0 * p.ldy + col0, p.ldy, rows_valid, w, lane);
    } else if (row < p.M) {
      __nv_bfloat16* dp = p.y_pre + (long long)row * p.ldy + col0;
      for (int i = 0; i < 32 && col0 + i < p.N; ++i)
        dp[i] = __ushort_as_bfloat16((unsigned short)((i & 1) ? (w[i >> 1] >> 16) : (w[i >> 1] & 0xFFFFu)));
    }
  }
  if (p.act == 1) {
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = gelu_tanh_bf16x2(w[i]);
  }
  if (p.residual) {
    const bool vec_r = full && ((p.ldr & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0);
    if (use_pre || vec_r) {
      uint4 c[4];
      uint32_t rw[16];
      if (use_pre) {
#pragma unroll
        for (int i = 0; i < 4; ++i) c[i] = rpre[i];
      } else {
        warp_fetch_tile(p.residual + (long long)row0 * p.ldr + col0, p.ldr, rows_valid, c, lane);
      }
      warp_rows_from_fetched(xp, c, rw, lane);
      if (p.act == 2) {                                   // gelu transpose: `residual` is the saved pre-activation z
#pragma unroll
        for (int i = 0; i < 16; ++i)
          w[i] = pack_bf16(bf16_lo(w[i]) * gelu_tanh_grad(bf16_lo(rw[i])), bf16_hi(w[i]) * gelu_tanh_grad(bf16_hi(rw[i])));
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = bf16x2_add(w[i], rw[i]);
      }
    } else if (row < p.M) {
      const __nv_bfloat16* r = p.residual + (long long)row * p.ldr + col0;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float r0 = (col0 + 2 * i < p.N) ? __bfloat162float(r[2 * i]) : 0.f;
        const float r1 = (col0 + 2 * i + 1 < p.N) ? __bfloat162float(r[2 * i + 1]) : 0.f;
        if (p.act == 2) w[i] = pack_bf16(bf16_lo(w[i]) * gelu_tanh_grad(r0), bf16_hi(w[i]) * gelu_tanh_grad(r1));
        else w[i] = pack_bf16(bf16_lo(w[i]) + r0, bf16_hi(w[i]) + r1);
      }
    }
  }
  if (p.tma_store) {
    warp_store_tile_tma(xp, sel, tm_y, col0, row0, w, lane);
  } else if (vec_y) {
    warp_store_tile(xp, p.y + (long long)row0 * p.ldy + col0, p.ldy, rows_valid, w, lane);
  } else if (row < p.M) {
    __nv_bfloat16* d = p.y + (long long)row * p.ldy + col0;
    if ((p.ldy & 7) == 0 && (p.N & 7) == 0 && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0)) {   // ragged last chunk
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (col0 + 8 * i < p.N) reinterpret_cast<uint4*>(d)[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    } else {
      for (int i = 0; i < 32 && col0 + i < p.N; ++i)
        d[i] = __ushort_as_bfloat16((unsigned short)((i & 1) ? (w[i >> 1] >> 16) : (w[i >> 1] & 0xFFFFu)));
    }
  }
}

// One 32-column chunk == one attention head (head_dim 32).  Warp-cooperative like epi_store.
__device__ __forceinline__ void epi_qkv(const GemmParams& p, int row0, int lane, int col0, const uint32_t (&acc)[32],
                                        const float (&sn)[16], const float (&cs)[16], uint32_t xp) {
  if (row0 >= p.M || col0 >= p.N) return;
  const int rows_valid = min(32, p.M - row0);
  const int head = col0 >> 5;
  float f[32];
#pragma unroll
  for (int i = 0; i < 32; i += 2) {                        // projection output is bf16
    f[i] = __uint_as_float(acc[i]);
    f[i + 1] = __uint_as_float(acc[i + 1]);
    bf16_round2(f[i], f[i + 1]);
  }
  if (head < p.Nq + p.Nkv) {
    float o[32];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      o[i] = f[i] * cs[i] - f[i + 16] * sn[i];
      o[i + 16] = f[i + 16] * cs[i] + f[i] * sn[i];
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = o[i];
  }
  __nv_bfloat16* dst;
  long long ld;
  if (head < p.Nq) {
    dst = p.q + (long long)row0 * p.q_row_stride + head * 32;
    ld = p.q_row_stride;
  } else {
    long long slot_off = 0;
    if (p.cache_S > 0) {
      int p0 = p.pos[0] % p.cache_S;
      if (p0 < 0) p0 += p.cache_S;
      slot_off = (long long)p0 * (p.Nkv * 32);
    }
    if (head < p.Nq + p.Nkv) {
      dst = p.k + (long long)row0 * p.k_row_stride + slot_off + (head - p.Nq) * 32;
      ld = p.k_row_stride;
    } else {
      dst = p.v + (long long)row0 * p.v_row_stride + slot_off + (head - p.Nq - p.Nkv) * 32;
      ld = p.v_row_stride;
    }
  }
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = pack_bf16(f[2 * i], f[2 * i + 1]);
  if ((ld & 7) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    warp_store_tile(xp, dst, ld, rows_valid, w, lane);
  } else if (lane < rows_valid) {
    store_bf16x32(dst + (long long)lane * ld, f);
  }
}

__device__ __forceinline__ void epi_glu(const GemmParams& p, int row0, int lane, int hcol0, const uint32_t (&u)[32],
                                        const uint32_t (&g)[32], uint32_t xp) {
  if (row0 >= p.M || hcol0 >= p.F) return;
  const int rows_valid = min(32, p.M - row0);
  uint32_t uw[16], gw[16], hw[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    uw[i] = pack_bf16(__uint_as_float(u[2 * i]), __uint_as_float(u[2 * i + 1]));
    gw[i] = pack_bf16(__uint_as_float(g[2 * i]), __uint_as_float(g[2 * i + 1]));
    const uint32_t aw = gelu_tanh_bf16x2(uw[i]);
    hw[i] = bf16x2_mul(aw, gw[i]);
  }
  const long long o = (long long)row0 * p.F + hcol0;
  warp_store_tile(xp, p.h + o, p.F, rows_valid, hw, lane);
  if (p.u_save) warp_store_tile(xp, p.u_save + o, p.F, rows_valid, uw, lane);
  if (p.g_save) warp_store_tile(xp, p.g_save + o, p.F, rows_valid, gw, lane);
}

// fp32-weight Linear emulated with a bf16 hi/lo split of the weight: out[j] = acc[j] + acc[64 + j] + bias[j].
__device__ __forceinline__ void epi_pairsum(const GemmParams& p, int row, int col0, const uint32_t (&hi)[32],
                                            const uint32_t (&lo)[32]) {
  if (row >= p.M) return;
  float* d = p.y_f32 + (long long)row * p.ldy32 + col0;
  for (int i = 0; i < 32; ++i) {
    if (col0 + i < p.N) {
      float v = __uint_as_float(hi[i]) + __uint_as_float(lo[i]);
      if (p.bias_f32) v += p.bias_f32[col0 + i];
      d[i] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ kernel
template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const GemmParams p, const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
               const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_ypre) {
  using L = SmemLayout<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint64_t* normed = bars + 2 * STAGES + 5;           // [EPI_QKV pre-norm] both K blocks of the A tile are normalised

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = tmem_cols_for(BLOCK_N);

  const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (EPI == EPI_PAIRSUM) ? 1 : (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;

  griddep_launch();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 8);
    }
    mbar_init(normed, 8);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_wait();                              // operands / residual / positions may come from the previous kernel
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          mbar_arrive_expect_tx(&full[stage], L::STAGE_BYTES);
          tma_load_2d(&tm_a, &full[stage], sa, kb * BLOCK_K, m_blk * BLOCK_M);
          tma_load_2d(&tm_b, &full[stage], sb, kb * BLOCK_K, n_blk * BLOCK_N);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // warp-uniform control flow, one elected lane per tcgen05 instruction (descriptors stay in uniform registers)
    {
      constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        if constexpr (EPI == EPI_QKV) {
          if (p.norm_scale) mbar_wait(normed, it & 1);      // the epilogue warps have normalised this tile's rows
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t sb = sa + A_STAGE_BYTES;
          const uint64_t adesc = umma_smem_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = umma_smem_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
            umma_bf16_1(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_1(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_1(&tfull[as]);
      }
    }
  } else {
    // ================================ epilogue warps ================================
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;        // which of the two warps of this quarter: takes chunks with (c/32)&1 == half
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t xp = smem_u32(smem + L::XPOSE_OFFSET + (warp - 2) * 4096);
    uint32_t xsel = 0;                      // which of the warp's two 2 KB buffers the next TMA store tile takes
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int row0 = m_blk * BLOCK_M + quarter * 32;      // first row of this warp
      const int row = row0 + lane;
      float sn[16], cs[16];
      if constexpr (EPI == EPI_QKV) {
        if (p.norm_scale) {
          // ---- pre-norm (network.py:160: history_norm) on the A tile, in place in the swizzled TMA layout: this warp
          //      owns K block `half` of its 32 rows; the two halves of a row exchange their sums of squares via smem
          float* sSq = reinterpret_cast<float*>(smem + L::NORM_OFFSET);
          float* sScale = sSq + 256;
          if (it == 0 && warp == 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) sScale[lane + 32 * i] = p.norm_scale[lane + 32 * i];
          }
          const int lin = 2 * it + half;                       // running k-block counter of the producer (num_kb == 2)
          mbar_wait(&full[lin % STAGES], (lin / STAGES) & 1);
          const uint32_t xrow = smem_u32(smem + (lin % STAGES) * L::STAGE_BYTES + row_in_tile * 128);
          uint4 xv[8];
          float ss = 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            xv[c] = lds_v4(xrow + ((c ^ (row_in_tile & 7)) << 4));
            const uint32_t w4[4] = {xv[c].x, xv[c].y, xv[c].z, xv[c].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = bf16_lo(w4[e]), b2 = bf16_hi(w4[e]);
              ss += a * a + b2 * b2;
            }
          }
          sSq[half * 128 + row_in_tile] = ss;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const float inv = rsqrtf((sSq[row_in_tile] + sSq[128 + row_in_tile]) * (1.0f / 128.0f) + p.norm_eps);
          const float4* sc4 = reinterpret_cast<const float4*>(sScale + half * 64);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 s0 = sc4[2 * c], s1 = sc4[2 * c + 1];
            uint4 o;
            o.x = pack_bf16(bf16_lo(xv[c].x) * (inv * s0.x), bf16_hi(xv[c].x) * (inv * s0.y));
            o.y = pack_bf16(bf16_lo(xv[c].y) * (inv * s0.z), bf16_hi(xv[c].y) * (inv * s0.w));
            o.z = pack_bf16(bf16_lo(xv[c].z) * (inv * s1.x), bf16_hi(xv[c].z) * (inv * s1.y));
            o.w = pack_bf16(bf16_lo(xv[c].w) * (inv * s1.z), bf16_hi(xv[c].w) * (inv * s1.w));
            sts_v4(xrow + ((c ^ (row_in_tile & 7)) << 4), o);
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, 256;" ::: "memory");       // sSq may be rewritten by the next tile
          if (lane == 0) mbar_arrive(normed);
        }
        if (row < p.M) {
          const float pf = (float)p.pos[row % p.pos_len];
#pragma unroll
          for (int i = 0; i < 16; ++i) rope_sincos(pf / p.timescale[i], &sn[i], &cs[i]);
        }
      }
      // residual rows of this tile are fetched BEFORE waiting for the accumulator (hides the global-load latency)
      constexpr int kMine = (BLOCK_N / 32 + 1) / 2;                  // 32-column chunks owned by this warp (at most)
      constexpr int kPre = (EPI == EPI_STORE && BLOCK_N <= 128) ? kMine : 1;
      uint4 rpre[kPre][4];
      bool have_pre = false;
      if constexpr (EPI == EPI_STORE && BLOCK_N <= 128) {
        if (p.residual && p.y && row0 < p.M && (p.ldr & 7) == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0 &&
            (n_blk + 1) * BLOCK_N <= p.N) {
          const int rows_valid = min(32, p.M - row0);
#pragma unroll
          for (int j = 0; j < kMine; ++j)
            if ((2 * j + half) * 32 < BLOCK_N)
              warp_fetch_tile(p.residual + (long long)row0 * p.ldr + n_blk * BLOCK_N + (2 * j + half) * 32, p.ldr,
                              rows_valid, rpre[j], lane);
          have_pre = true;
        }
      }
      // wide tiles (BLOCK_N > 128): the residual / pre-activation tile of a chunk is fetched one chunk ahead instead
      uint4 rnext[4];
      bool pre_wide = false;
      if constexpr (EPI == EPI_STORE && BLOCK_N > 128) {
        pre_wide = p.residual && p.y && row0 < p.M && (p.ldr & 7) == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0;
        if (pre_wide && n_blk * BLOCK_N + 32 * half + 32 <= p.N)
          warp_fetch_tile(p.residual + (long long)row0 * p.ldr + n_blk * BLOCK_N + 32 * half, p.ldr, min(32, p.M - row0),
                          rnext, lane);
      }
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * BLOCK_N;
      if constexpr (EPI == EPI_GLU) {
        static_assert(EPI != EPI_GLU || BLOCK_N == 256, "GLU tile is 128 up + 128 gate columns");
#pragma unroll 1
        for (int c = 32 * half; c < 128; c += 64) {
          uint32_t u[32], g[32];
          tmem_ld_32x32(t_row + c, u);
          tmem_ld_32x32(t_row + 128 + c, g);
          tmem_ld_wait();
          epi_glu(p, row0, lane, n_blk * 128 + c, u, g, xp);
        }
      } else if constexpr (EPI == EPI_PAIRSUM) {
        {
          const int c = 32 * half;
          uint32_t hi[32], lo[32];
          tmem_ld_32x32(t_row + c, hi);
          tmem_ld_32x32(t_row + 64 + c, lo);
          tmem_ld_wait();
          epi_pairsum(p, row, c, hi, lo);
        }
      } else if constexpr (EPI == EPI_STORE && BLOCK_N <= 128) {
#pragma unroll
        for (int j = 0; j < kMine; ++j) {
          const int c = (2 * j + half) * 32;
          if (c < BLOCK_N) {
            uint32_t acc[32];
            tmem_ld_32x32(t_row + c, acc);
            tmem_ld_wait();
            epi_store(p, row0, lane, n_blk * BLOCK_N + c, acc, xp, rpre[j], have_pre, &tm_y, &tm_ypre, xsel);
          }
        }
      } else if constexpr (EPI == EPI_STORE) {
#pragma unroll 1
        for (int c = 32 * half; c < BLOCK_N; c += 64) {
          const int col0 = n_blk * BLOCK_N + c;
          const bool have = pre_wide && col0 + 32 <= p.N;
          uint4 rcur[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) rcur[i] = rnext[i];
          if (pre_wide && c + 64 < BLOCK_N && col0 + 96 <= p.N)
            warp_fetch_tile(p.residual + (long long)row0 * p.ldr + col0 + 64, p.ldr, min(32, p.M - row0), rnext, lane);
          uint32_t acc[32];
          tmem_ld_32x32(t_row + c, acc);
          tmem_ld_wait();
          epi_store(p, row0, lane, col0, acc, xp, rcur, have, &tm_y, &tm_ypre, xsel);
        }
      } else {
#pragma unroll 1
        for (int c = 32 * half; c < BLOCK_N; c += 64) {
          uint32_t acc[32];
          tmem_ld_32x32(t_row + c, acc);
          tmem_ld_wait();
          epi_qkv(p, row0, lane, n_blk * BLOCK_N + c, acc, sn, cs, xp);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
    }
    if (p.tma_store && lane == 0) bulk_wait_group0();       // shared memory must outlive the stores reading it
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ wgrad kernel
// dW[K,N] (fp32, red.add) += X[M,K]^T . dY[M,N].  Output tile 128 (K rows) x WG_N; the contraction runs over a
// chunk of the M tokens, 64 tokens per stage.  A = X^T and B = dY^T are MN-major: the smem tile of a stage is
// [64 tokens][64 features] x (features/64) boxes, each box 8 KB with the 128-byte swizzle.
template <int WG_N>
__global__ void __launch_bounds__(WGRAD_THREADS, 1)
gemm_wgrad_kernel(int M, int N, int K, float* __restrict__ dw, long long lddw, float* __restrict__ partial,
                  int tokens_per_cta, int swap_lbo_sbo, const __grid_constant__ CUtensorMap tm_x,
                  const __grid_constant__ CUtensorMap tm_dy) {
  constexpr int A_BYTES = 2 * 8192;            // 128 features = 2 boxes
  constexpr int B_BYTES = (WG_N / 64) * 8192;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = WG_N <= 32 ? 32 : WG_N <= 64 ? 64 : WG_N <= 128 ? 128 : 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_blk = blockIdx.x;   // 128-row block of dW (input features)
  const int n_blk = blockIdx.y;   // WG_N-col block of dW (output features)
  const int tok0 = blockIdx.z * tokens_per_cta;
  const int tok1 = min(M, tok0 + tokens_per_cta);
  const int num_it = (tok1 - tok0 + 63) / 64;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_dy);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&tfull[0], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < num_it; ++it) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        uint8_t* sb = sa + A_BYTES;
        mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
        const int tok = tok0 + it * 64;   // rows past tok1 but < M belong to the next CTA: masked by zero weight? no:
                                          // tokens_per_cta is a multiple of 64, so only the global M tail is ragged
                                          // and TMA zero-fills rows >= M.
#pragma unroll
        for (int b = 0; b < 2; ++b) tma_load_2d(&tm_x, &full[stage], sa + b * 8192, k_blk * 128 + b * 64, tok);
#pragma unroll
        for (int b = 0; b < WG_N / 64; ++b) tma_load_2d(&tm_dy, &full[stage], sb + b * 8192, n_blk * WG_N + b * 64, tok);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, WG_N, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      // MN-major SW128 canonical layout ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO)) in elements:
      // LBO = byte distance between 64-element blocks along MN (8192), SBO = between 8-row groups along K (1024).
      const uint32_t lbo = swap_lbo_sbo ? 1024 : 8192;
      const uint32_t sbo = swap_lbo_sbo ? 8192 : 1024;
      for (int it = 0; it < num_it; ++it) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // 16 tokens along K = 16 rows of 128 bytes = 2048 bytes
          const uint64_t adesc = umma_smem_desc_sw128(sa + k * 2048, lbo, sbo);
          const uint64_t bdesc = umma_smem_desc_sw128(sb + k * 2048, lbo, sbo);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tfull[0]);
    }
  } else {
    const int quarter = warp & 3;
    const int krow = k_blk * 128 + quarter * 32 + lane;
    mbar_wait(&tfull[0], 0);
    tc_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < WG_N; c += 32) {
      uint32_t acc[32];
      tmem_ld_32x32(t_row + c, acc);
      tmem_ld_wait();
      if (partial) {
        // split-K partial tile -> workspace [split][K_pad][N_pad] (plain stores; reduced in fixed order afterwards)
        const long long kp = (long long)gridDim.x * 128, np = (long long)gridDim.y * WG_N;
        float4* d4 = reinterpret_cast<float4*>(partial + ((long long)blockIdx.z * kp + (k_blk * 128 + quarter * 32 + lane)) * np +
                                               n_blk * WG_N + c);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          d4[i] = num_it > 0 ? make_float4(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1]),
                                           __uint_as_float(acc[4 * i + 2]), __uint_as_float(acc[4 * i + 3]))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
      } else if (krow < K && num_it > 0) {
        float* d = dw + (long long)krow * lddw + n_blk * WG_N + c;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (n_blk * WG_N + c + i < N) atomicAdd(d + i, __uint_as_float(acc[i]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// dw[k][n] += sum over splits of partial[s][k][n]   (fixed summation order: deterministic)
// Block = 8 warps: each warp owns a strided subset of the splits for the same 32 float4 (512 contiguous bytes), with
// four independent loads in flight; the 8 per-warp sums are combined through shared memory in warp order.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int kp, int np,
                                                           float* __restrict__ dw, long long lddw, int K, int N) {
  __shared__ float4 s_part[8][32];
  const int n4 = np / 4;
  const int total = K * n4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long split_stride = (long long)kp * np;
  for (int base = blockIdx.x * 32; base < total; base += gridDim.x * 32) {
    const int i = base + lane;
    const bool live = i < total;
    const int k = live ? i / n4 : 0, n = live ? (i % n4) * 4 : 0;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live && n < N) {
      const float* p = partial + (long long)k * np + n;
      int s = warp;
      for (; s + 24 < splits; s += 32) {
        const float4 v0 = *reinterpret_cast<const float4*>(p + (long long)s * split_stride);
        const float4 v1 = *reinterpret_cast<const float4*>(p + (long long)(s + 8) * split_stride);
        const float4 v2 = *reinterpret_cast<const float4*>(p + (long long)(s + 16) * split_stride);
        const float4 v3 = *reinterpret_cast<const float4*>(p + (long long)(s + 24) * split_stride);
        a.x += (v0.x + v1.x) + (v2.x + v3.x); a.y += (v0.y + v1.y) + (v2.y + v3.y);
        a.z += (v0.z + v1.z) + (v2.z + v3.z); a.w += (v0.w + v1.w) + (v2.w + v3.w);
      }
      for (; s < splits; s += 8) {
        const float4 v = *reinterpret_cast<const float4*>(p + (long long)s * split_stride);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
    }
    s_part[warp][lane] = a;
    __syncthreads();
    if (warp == 0 && live && n < N) {
#pragma unroll
      for (int w = 1; w < 8; ++w) {
        const float4 v = s_part[w][lane];
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
      float* d = dw + (long long)k * lddw + n;
      d[0] += a.x;
      if (n + 1 < N) d[1] += a.y;
      if (n + 2 < N) d[2] += a.z;
      if (n + 3 < N) d[3] += a.w;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ one-hot wgrad
// Weight gradient of the first observation conv (reference GridCnnObsEncoder, observation.py:66-96, through autodiff):
//   dW1[(tap, channel c)][ch] += sum over output rows (agent-step m, position q) of onehot(obs)[row][(tap, c)] * dz1[row][ch]
// i.e. the same X^T . dY contraction as gemm_wgrad_kernel, except that X (the one-hot im2col matrix, 566 MB per
// minibatch if materialised) is never written to memory: four builder warps generate each [64 rows][128 features]
// operand tile straight into shared memory in the MN-major SW128 layout the MMA reads (zeros + at most taps*K ones per
// row) from the uint8 observation, while warp 0 streams the matching dz1 tile with TMA.
struct OneHotWgradParams {
  const uint8_t* obs;      // (M, IH, IW, K)
  long long rows;          // M * OH * OW
  int IH, IW, OH, OW, K, C, kh, kw, sh, sw;
  int offs[5];             // prefix sums of the one-hot sizes
  float* partial;          // [splits][128][64]
  int rows_per_cta;        // multiple of 64
};

constexpr int OHW_BUILDERS = 8;                              // builder warps
constexpr int OHW_THREADS = 64 + OHW_BUILDERS * 32;

// KC = observation components per cell; the receptive field is 3x3 (9 taps).
template <int KC>
__global__ void __launch_bounds__(OHW_THREADS, 1)
onehot_wgrad_kernel(const OneHotWgradParams p, const __grid_constant__ CUtensorMap tm_dy) {
  constexpr int A_BYTES = 2 * 8192, B_BYTES = 8192, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)blockIdx.x * p.rows_per_cta;
  const long long row1 = min(p.rows, row0 + p.rows_per_cta);
  const int num_it = row1 > row0 ? (int)((row1 - row0 + 63) / 64) : 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_dy);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1 + OHW_BUILDERS);   // TMA producer (expect_tx) + builder warps
      mbar_init(&empty[s], 1);
    }
    mbar_init(&tfull[0], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < num_it; ++it) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], B_BYTES);
        // rows past row1 but < rows belong to the next CTA: their one-hot rows are built as zeros, so they add nothing
        tma_load_2d(&tm_dy, &full[stage], smem + stage * STAGE_BYTES + A_BYTES, 0, (int)(row0 + (long long)it * 64));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < num_it; ++it) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base, umma_smem_desc_sw128(sa + k * 2048, 8192, 1024), umma_smem_desc_sw128(sb + k * 2048, 8192, 1024),
                    idesc, (it | k) != 0 ? 1u : 0u);
        umma_commit(&empty[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tfull[0]);
    }
  } else {
    // ---- builders: thread = (row of the stage tl, 64-feature box, half): zero-fills 4 of the 8 chunks of its box
    //      row and turns on the ones of taps [5*half, 5*half+5) that fall into its box.  The observation bytes of the
    //      NEXT stage are requested before this stage is written, so their latency overlaps the stores.
    const int tb = threadIdx.x - 64;
    const int tl = tb & 63, box = (tb >> 7) & 1, half = (tb >> 6) & 1;
    const int tap0 = half * 5, tap1 = half ? 9 : 5;
    const int ohw = p.OH * p.OW;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t cur[5], nxt[5];
    auto fetch = [&](int it, uint32_t (&dst)[5]) {
      const long long row = row0 + (long long)it * 64 + tl;
#pragma unroll
      for (int i = 0; i < 5; ++i) dst[i] = 0xFFFFFFFFu;        // "out of range" for every component: no ones
      if (it < num_it && row < row1) {
        const long long m = row / ohw;
        const int q = (int)(row % ohw);
        const int oy = q / p.OW, ox = q % p.OW;
        const uint8_t* cell0 = p.obs + ((m * p.IH + oy * p.sh) * p.IW + ox * p.sw) * KC;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const int tap = tap0 + i;
          if (tap < tap1) {
            const uint8_t* cell = cell0 + ((tap / 3) * p.IW + (tap % 3)) * KC;
            uint32_t v = 0;
            if (KC == 2) {
              v = *reinterpret_cast<const unsigned short*>(cell);   // cells are 2-byte aligned when KC == 2
            } else if (KC == 4) {
              v = *reinterpret_cast<const uint32_t*>(cell);
            } else {
#pragma unroll
              for (int kc = 0; kc < KC; ++kc) v |= (uint32_t)cell[kc] << (8 * kc);
            }
            dst[i] = v;
          }
        }
      }
    };
    fetch(0, cur);
    for (int it = 0; it < num_it; ++it) {
      fetch(it + 1, nxt);
      mbar_wait(&empty[stage], phase ^ 1);
      const uint32_t rowaddr = smem_u32(smem + stage * STAGE_BYTES + box * 8192 + tl * 128);
#pragma unroll
      for (int c = 0; c < 4; ++c) sts_v4(rowaddr + (((4 * half + c) ^ (tl & 7)) << 4), make_uint4(0, 0, 0, 0));
      // the partner thread (other half) zero-fills the other 4 chunks: both must be done before any one is set
      asm volatile("bar.sync 1, %0;" ::"r"(OHW_BUILDERS * 32) : "memory");
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const int tap = tap0 + i;
        if (tap < tap1) {
#pragma unroll
          for (int kc = 0; kc < KC; ++kc) {
            const int val = (cur[i] >> (8 * kc)) & 255;
            const int f = tap * p.C + p.offs[kc] + val;
            if (val < p.offs[kc + 1] - p.offs[kc] && (f >> 6) == box)
              asm volatile("st.shared.u16 [%0], %1;" ::"r"(rowaddr + ((((f & 63) >> 3) ^ (tl & 7)) << 4) + (f & 7) * 2),
                           "h"((unsigned short)0x3F80) : "memory");
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[stage]);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
#pragma unroll
      for (int i = 0; i < 5; ++i) cur[i] = nxt[i];
    }
    // ---- epilogue (warps 2-5): TMEM (128 features x 64 columns) -> partial[split][128][64]
    if (warp < 6) {
      const int quarter = warp & 3;
      if (num_it > 0) {
        mbar_wait(&tfull[0], 0);
        tc_fence_after();
      }
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        uint32_t acc[32];
        if (num_it > 0) {
          tmem_ld_32x32(t_row + c, acc);
          tmem_ld_wait();
        }
        float4* d4 = reinterpret_cast<float4*>(p.partial + ((long long)blockIdx.x * 128 + quarter * 32 + lane) * 64 + c);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          d4[i] = num_it > 0 ? make_float4(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1]),
                                           __uint_as_float(acc[4 * i + 2]), __uint_as_float(acc[4 * i + 3]))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 64);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int make_tmap_bf16_2d_sw(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                                uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle swz);
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_cols, uint32_t box_rows) {
  return make_tmap_bf16_2d_sw(out, base, rows, cols, ld, box_cols, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}
int make_tmap_bf16_2d_sw64(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                           uint32_t box_cols, uint32_t box_rows) {
  return make_tmap_bf16_2d_sw(out, base, rows, cols, ld, box_cols, box_rows, CU_TENSOR_MAP_SWIZZLE_64B);
}
static int make_tmap_bf16_2d_sw(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                                uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return MPX_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld * 2) & 15) != 0) {
    set_error("TMA operand must be 16-byte aligned with a 16-byte-multiple row pitch (base=%p ld=%llu)", base,
              (unsigned long long)ld);
    return MPX_ERR_INVALID;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu box=%ux%u)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_cols, box_rows);
    return MPX_ERR_CUDA;
  }
  return MPX_OK;
}

template <int BLOCK_N, int EPI>
static int launch_tn(const GemmParams& p, const void* a, long long lda, const void* bt, long long ldb,
                     cudaStream_t stream) {
  using L = SmemLayout<BLOCK_N>;
  CUtensorMap tm_a, tm_b;
  int rc = make_tmap_bf16_2d(&tm_a, a, p.M, p.K, lda, BLOCK_K, BLOCK_M);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_b, bt, EPI == EPI_PAIRSUM ? 128 : p.N, p.K, ldb, BLOCK_K, BLOCK_N);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_kernel<BLOCK_N, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         L::TOTAL);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm_tn): %s", cudaGetErrorString(e));
      return MPX_ERR_CUDA;
    }
    attr_set = true;
  }
  const int tiles = ceil_div(p.M, BLOCK_M) * (EPI == EPI_PAIRSUM ? 1 : ceil_div(p.N, BLOCK_N));
  const int grid = tiles < num_sms() ? tiles : num_sms();
  GemmParams pp = p;
  // bf16 outputs through TMA stores (32 x 32 boxes, SWIZZLE_64B) when nothing else needs the per-warp transposer
  CUtensorMap tm_y = tm_a, tm_ypre = tm_a;
  pp.tma_store = 0;
  if (EPI == EPI_STORE && p.y && !p.residual && !p.y_f32 && (p.ldy & 7) == 0 &&
      (reinterpret_cast<uintptr_t>(p.y) & 15) == 0 && (!p.y_pre || (reinterpret_cast<uintptr_t>(p.y_pre) & 15) == 0)) {
    rc = make_tmap_bf16_2d_sw(&tm_y, p.y, p.M, p.N, p.ldy, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    if (p.y_pre) {
      rc = make_tmap_bf16_2d_sw(&tm_ypre, p.y_pre, p.M, p.N, p.ldy, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
    }
    pp.tma_store = 1;
  }
  // rollout-sized GEMMs (one decode step of the per-op chain, hidden 256) launch programmatically like the fused decode
  // kernels (family 4): -3 us per layer of a multitask step; training-sized ones stay in the "everything else" family
  const int pdl_family = p.M <= 8192 ? 4 : 1;
  cudaError_t le = launch_ex(gemm_tn_kernel<BLOCK_N, EPI>, dim3(grid), dim3(GEMM_THREADS), L::TOTAL, stream, 0, pdl_family, pp, tm_a,
                             tm_b, tm_y, tm_ypre);
  if (le != cudaSuccess) {
    set_error("launch gemm_tn_kernel: %s", cudaGetErrorString(le));
    return MPX_ERR_CUDA;
  }
  return check_launch("gemm_tn_kernel");
}

template <int WG_N>
static int launch_wgrad(const mpx_wgrad_args* a, cudaStream_t stream) {
  constexpr int STAGE_BYTES = 2 * 8192 + (WG_N / 64) * 8192;
  constexpr int SMEM = STAGES * STAGE_BYTES + 256 + 1024;
  CUtensorMap tm_x, tm_dy;
  int rc = make_tmap_bf16_2d(&tm_x, a->x, a->M, a->K, a->ldx, 64, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_dy, a->dy, a->M, a->N, a->ldy, 64, 64);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_wgrad_kernel<WG_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm_wgrad): %s", cudaGetErrorString(e));
      return MPX_ERR_CUDA;
    }
    attr_set = true;
  }
  const int kb = ceil_div(a->K, 128), nb = ceil_div(a->N, WG_N);
  // split the token dimension so that the grid covers the machine about twice
  int splits = ceil_div(2 * num_sms(), kb * nb);
  int tokens_per_cta = ceil_div(ceil_div(a->M, splits), 64) * 64;
  if (tokens_per_cta < 64) tokens_per_cta = 64;
  splits = ceil_div(a->M, tokens_per_cta);
  dim3 grid(kb, nb, splits);
  const size_t need = (size_t)splits * kb * 128 * nb * WG_N * sizeof(float);
  float* partial = (a->workspace && a->workspace_bytes >= need && splits > 1) ? reinterpret_cast<float*>(a->workspace) : nullptr;
  gemm_wgrad_kernel<WG_N><<<grid, WGRAD_THREADS, SMEM, stream>>>(a->M, a->N, a->K, a->dw, a->lddw, partial, tokens_per_cta,
                                                               g_debug_swap_lbo_sbo, tm_x, tm_dy);
  if (partial) {
    const int total = a->K * (nb * WG_N / 4);
    int rg = ceil_div(total, 32);
    if (rg > num_sms() * 8) rg = num_sms() * 8;
    wgrad_reduce_kernel<<<rg, 256, 0, stream>>>(partial, splits, kb * 128, nb * WG_N, a->dw, a->lddw, a->K, a->N);
  }
  return check_launch("gemm_wgrad_kernel");
}

}  // namespace mpx

using namespace mpx;

extern "C" int mpx_linear(const mpx_linear_args* a, mpx_stream_t stream) {
  MPX_REQUIRE(a && a->x && a->wt && (a->y || a->y_f32), "mpx_linear: null pointer");
  MPX_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "mpx_linear: empty shape M=%d N=%d K=%d", a->M, a->N, a->K);
  MPX_REQUIRE(a->act == 0 || a->act == 1 || a->act == 2, "mpx_linear: unknown activation %d", a->act);
  MPX_REQUIRE(a->act != 2 || (a->residual && a->y && !a->bias && !a->y_pre),
              "mpx_linear: act 2 (gelu transpose) needs y and the pre-activation in `residual`, and takes no bias / y_pre");
  GemmParams p = {};
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(a->bias);
  p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual);
  p.ldr = a->ldr;
  p.y = reinterpret_cast<__nv_bfloat16*>(a->y);
  p.ldy = a->ldy;
  p.y_f32 = a->y_f32;
  p.ldy32 = a->ldy32;
  p.act = a->act;
  p.y_pre = reinterpret_cast<__nv_bfloat16*>(a->y_pre);
  MPX_REQUIRE(!a->y_pre || a->y, "mpx_linear: y_pre requires y");
  cudaStream_t s = (cudaStream_t)stream;
  // widest N tile that still gives the machine enough CTAs (small-M decode GEMMs are latency bound: more, narrower
  // tiles shorten both the K loop's exposed latency and the epilogue)
  const int mt = ceil_div(a->M, BLOCK_M);
  const int want = num_sms();
  if (a->N > 128 && mt * ceil_div(a->N, 256) >= want) return launch_tn<256, EPI_STORE>(p, a->x, a->ldx, a->wt, a->ldw, s);
  if (a->N > 64 && mt * ceil_div(a->N, 128) >= want) return launch_tn<128, EPI_STORE>(p, a->x, a->ldx, a->wt, a->ldw, s);
  if (a->N > 32 && mt * ceil_div(a->N, 64) >= want) return launch_tn<64, EPI_STORE>(p, a->x, a->ldx, a->wt, a->ldw, s);
  return launch_tn<32, EPI_STORE>(p, a->x, a->ldx, a->wt, a->ldw, s);
}

extern "C" size_t mpx_linear_wgrad_workspace_bytes(int M, int N, int K) {
  const int wg_n = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  const int kb = ceil_div(K, 128), nb = ceil_div(N, wg_n);
  int splits = ceil_div(2 * num_sms(), kb * nb);
  int tokens_per_cta = ceil_div(ceil_div(M, splits), 64) * 64;
  if (tokens_per_cta < 64) tokens_per_cta = 64;
  splits = ceil_div(M, tokens_per_cta);
  return (size_t)splits * kb * 128 * nb * wg_n * sizeof(float);
}

extern "C" int mpx_linear_wgrad(const mpx_wgrad_args* a, mpx_stream_t stream) {
  MPX_REQUIRE(a && a->x && a->dy && a->dw, "mpx_linear_wgrad: null pointer");
  MPX_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "mpx_linear_wgrad: empty shape");
  MPX_REQUIRE(a->ldx % 8 == 0 && a->ldy % 8 == 0, "mpx_linear_wgrad: ldx and ldy must be multiples of 8");
  cudaStream_t s = (cudaStream_t)stream;
  if (a->N <= 64) return launch_wgrad<64>(a, s);
  if (a->N <= 128) return launch_wgrad<128>(a, s);
  return launch_wgrad<256>(a, s);
}

extern "C" int mpx_qkv_rope(const mpx_qkv_rope_args* a, mpx_stream_t stream) {
  MPX_REQUIRE(a && a->x && a->wt_qkv && a->pos && a->rope_timescale && a->q && a->k && a->v,
              "mpx_qkv_rope: null pointer");
  MPX_REQUIRE(a->D == 32, "mpx_qkv_rope: head_dim %d unsupported (only 32)", a->D);
  MPX_REQUIRE(a->M > 0 && a->H > 0 && a->Nq > 0 && a->Nkv > 0 && a->pos_len > 0, "mpx_qkv_rope: bad shape");
  const int N = (a->Nq + 2 * a->Nkv) * a->D;
  GemmParams p = {};
  p.M = a->M; p.N = N; p.K = a->H;
  p.pos = a->pos; p.pos_len = a->pos_len; p.timescale = a->rope_timescale;
  p.q = reinterpret_cast<__nv_bfloat16*>(a->q);
  p.k = reinterpret_cast<__nv_bfloat16*>(a->k);
  p.v = reinterpret_cast<__nv_bfloat16*>(a->v);
  p.q_row_stride = a->q_row_stride; p.k_row_stride = a->k_row_stride; p.v_row_stride = a->v_row_stride;
  p.cache_S = a->cache_S; p.Nq = a->Nq; p.Nkv = a->Nkv;
  p.norm_scale = a->norm_scale; p.norm_eps = a->norm_eps;
  MPX_REQUIRE(!a->norm_scale || a->H == 128, "mpx_qkv_rope: the fused pre-norm needs hidden_features == 128 (got %d)", a->H);
  MPX_REQUIRE((a->q_row_stride % 8) == 0 && (a->k_row_stride % 8) == 0 && (a->v_row_stride % 8) == 0,
              "mpx_qkv_rope: q/k/v row strides must be multiples of 8");
  cudaStream_t s = (cudaStream_t)stream;
  const int mt = ceil_div(a->M, BLOCK_M);
  const int want = (num_sms() * 3) / 5;
  if (N > 192 && mt * ceil_div(N, 256) >= want) return launch_tn<256, EPI_QKV>(p, a->x, a->ldx, a->wt_qkv, a->H, s);
  if (N > 128 && N <= 192 && mt >= want) return launch_tn<192, EPI_QKV>(p, a->x, a->ldx, a->wt_qkv, a->H, s);
  if (N > 64 && mt * ceil_div(N, 128) >= want) return launch_tn<128, EPI_QKV>(p, a->x, a->ldx, a->wt_qkv, a->H, s);
  if (N > 32 && mt * ceil_div(N, 64) >= want) return launch_tn<64, EPI_QKV>(p, a->x, a->ldx, a->wt_qkv, a->H, s);
  return launch_tn<32, EPI_QKV>(p, a->x, a->ldx, a->wt_qkv, a->H, s);
}

extern "C" int mpx_glu_up(const mpx_glu_up_args* a, mpx_stream_t stream) {
  MPX_REQUIRE(a && a->x && a->wt_ug && a->h, "mpx_glu_up: null pointer");
  MPX_REQUIRE(a->M > 0 && a->K > 0 && a->F > 0 && a->F % 128 == 0, "mpx_glu_up: F=%d must be a positive multiple of 128", a->F);
  GemmParams p = {};
  p.M = a->M; p.N = 2 * a->F; p.K = a->K; p.F = a->F;
  p.h = reinterpret_cast<__nv_bfloat16*>(a->h);
  p.u_save = reinterpret_cast<__nv_bfloat16*>(a->u_save);
  p.g_save = reinterpret_cast<__nv_bfloat16*>(a->g_save);
  return launch_tn<256, EPI_GLU>(p, a->x, a->ldx, a->wt_ug, a->K, (cudaStream_t)stream);
}

extern "C" int mpx_linear_f32w(const mpx_bf16* x, long long ldx, const mpx_bf16* wt_hilo, const float* bias, float* y,
                               long long ldy, int M, int N, int K, mpx_stream_t stream) {
  MPX_REQUIRE(x && wt_hilo && y, "mpx_linear_f32w: null pointer");
  MPX_REQUIRE(M > 0 && K > 0 && N > 0 && N <= 64, "mpx_linear_f32w: N=%d must be in [1,64]", N);
  GemmParams p = {};
  p.M = M; p.N = N; p.K = K;
  p.y_f32 = y; p.ldy32 = ldy; p.bias_f32 = bias;
  return launch_tn<128, EPI_PAIRSUM>(p, x, ldx, wt_hilo, K, (cudaStream_t)stream);
}

static int onehot_wgrad_splits(long long rows, int* rows_per_cta) {
  int splits = 2 * num_sms();
  long long per = ((rows + splits - 1) / splits + 63) / 64 * 64;
  if (per < 64) per = 64;
  *rows_per_cta = (int)per;
  return (int)((rows + per - 1) / per);
}

extern "C" size_t mpx_conv1_onehot_wgrad_workspace_bytes(long long M, int IH, int IW, int kh, int kw, int sh, int sw) {
  const int OH = (IH - kh) / sh + 1, OW = (IW - kw) / sw + 1;
  int per;
  const int splits = onehot_wgrad_splits(M * OH * OW, &per);
  return (size_t)splits * 128 * 64 * sizeof(float);
}

extern "C" int mpx_conv1_onehot_wgrad(const uint8_t* obs, const mpx_bf16* dz, float* dw, long long M, int IH, int IW, int K,
                                      const int* sizes, int kh, int kw, int sh, int sw, int cout, void* workspace,
                                      size_t workspace_bytes, mpx_stream_t stream) {
  MPX_REQUIRE(obs && dz && dw && sizes && workspace && M > 0, "mpx_conv1_onehot_wgrad: bad arguments");
  MPX_REQUIRE(K >= 1 && K <= 4 && kh == 3 && kw == 3, "mpx_conv1_onehot_wgrad: K=%d / %dx%d taps unsupported (3x3, K <= 4)", K, kh, kw);
  MPX_REQUIRE(K != 2 || (reinterpret_cast<uintptr_t>(obs) & 1) == 0, "mpx_conv1_onehot_wgrad: obs must be 2-byte aligned for K=2");
  MPX_REQUIRE(K != 4 || (reinterpret_cast<uintptr_t>(obs) & 3) == 0, "mpx_conv1_onehot_wgrad: obs must be 4-byte aligned for K=4");
  MPX_REQUIRE(cout > 0 && cout <= 64 && cout % 8 == 0, "mpx_conv1_onehot_wgrad: cout=%d must be a multiple of 8, <= 64", cout);
  OneHotWgradParams p = {};
  p.obs = obs; p.IH = IH; p.IW = IW; p.K = K; p.kh = kh; p.kw = kw; p.sh = sh; p.sw = sw;
  p.OH = (IH - kh) / sh + 1; p.OW = (IW - kw) / sw + 1;
  p.offs[0] = 0;
  for (int i = 0; i < 4; ++i) p.offs[i + 1] = p.offs[i] + (i < K ? sizes[i] : 0);
  p.C = p.offs[K];
  MPX_REQUIRE(kh * kw * p.C <= 128, "mpx_conv1_onehot_wgrad: %d one-hot features > 128", kh * kw * p.C);
  p.rows = M * p.OH * p.OW;
  MPX_REQUIRE(p.rows < (1LL << 31), "mpx_conv1_onehot_wgrad: too many rows");
  const int splits = onehot_wgrad_splits(p.rows, &p.rows_per_cta);
  MPX_REQUIRE(workspace_bytes >= (size_t)splits * 128 * 64 * sizeof(float), "mpx_conv1_onehot_wgrad: workspace too small");
  p.partial = reinterpret_cast<float*>(workspace);
  CUtensorMap tm_dy;
  int rc = make_tmap_bf16_2d(&tm_dy, dz, p.rows, cout, cout, 64, 64);
  if (rc) return rc;
  constexpr int SMEM = STAGES * (3 * 8192) + 256 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(onehot_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(onehot_wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(onehot_wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(onehot_wgrad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(onehot_wgrad): %s", cudaGetErrorString(e));
      return MPX_ERR_CUDA;
    }
    attr_set = true;
  }
  cudaStream_t s = (cudaStream_t)stream;
  switch (K) {
    case 1: onehot_wgrad_kernel<1><<<splits, OHW_THREADS, SMEM, s>>>(p, tm_dy); break;
    case 2: onehot_wgrad_kernel<2><<<splits, OHW_THREADS, SMEM, s>>>(p, tm_dy); break;
    case 3: onehot_wgrad_kernel<3><<<splits, OHW_THREADS, SMEM, s>>>(p, tm_dy); break;
    default: onehot_wgrad_kernel<4><<<splits, OHW_THREADS, SMEM, s>>>(p, tm_dy); break;
  }
  const int feats = kh * kw * p.C;
  int rg = ceil_div(feats * 16, 32);
  wgrad_reduce_kernel<<<rg, 256, 0, s>>>(p.partial, splits, 128, 64, dw, cout, feats, cout);
  return check_launch("onehot_wgrad_kernel");
}
