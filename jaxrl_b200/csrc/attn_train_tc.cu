// PPO-update attention forward on 5th-generation tensor cores: causal GQA flash attention with both contractions
// (S = Q.K^T and O = P.V) issued as tcgen05.mma, S and O accumulators in TMEM, operands staged by TMA.
// Restates the training branch of AttentionBlock.__call__ (reference mapox_trainer/model/attention.py:151-169);
// same arithmetic as attn_fwd_kernel in attn_train.cu (fp32 scores/softmax, un-normalised probabilities rounded to
// bf16 before P.V, fp32 accumulation, one bf16 rounding of O) and the same lse output for the backward pass.
//
// CTA = (batch row b, query head h, 128-query tile qt); 128 threads, thread r owns query row r (= TMEM lane r).
//   smem:  Q tile   [128 q  x 64 cols] SW128 (cols h*32.. of the (M, Nq*D+2*D) qkv buffer; only 32 are used)
//          KV tiles [keys   x 64 cols] SW128: each row is [k (32) | v (32)] of one position
//          P tile   [128 q  x 64 keys] SW128 bf16, written by the softmax threads, A operand of the second MMA
//   TMEM:  S0 [0,64)  S1 [64,128)  (double buffered: S(kt+1) is computed while softmax(kt) runs)   O [128,192)
//   MMA1:  S(128x64)  = Q (K-major, K = 32)        . K_tile (K-major)            2 x (M128 N64 K16)
//   MMA2:  O(128x64) += P (K-major, K = 64 keys)   . [K|V]_tile (MN-major)       4 x (M128 N64 K16)
//          columns [0,32) of O accumulate P.K (ignored), columns [32,64) accumulate P.V.
// Requires Nkv == 1 (k and v adjacent in the qkv buffer), head_dim 32, T % 128 == 0.  Other shapes use the
// mma.sync kernel in attn_train.cu.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.h"

namespace mpx {

constexpr int TC_QT = 128;   // queries per CTA
constexpr int TC_KT = 64;    // keys per inner tile

__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(128, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ o, long long ldo,
                   float* __restrict__ lse, int T, int Nq, int kv_col, float scale_log2, unsigned long long* dbg_buf) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nqt = T / TC_QT;
  const int qt = nqt - 1 - (int)blockIdx.x;      // heavy (late) query tiles first
  const int head = blockIdx.y, b = blockIdx.z;
  const int nkt = 2 * (qt + 1);                  // key tiles of 64 visible to this query tile
  const int nkv_boxes = qt + 1;                  // 128-row TMA boxes of [k|v]
  uint8_t* sQ = smem;                            // 16 KB
  uint8_t* sP = smem + 16384;                    // 16 KB
  uint8_t* sKV = smem + 32768;                   // nkv_boxes * 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + (size_t)nqt * 16384);
  uint64_t* bar_q = bars;            // Q tile landed
  uint64_t* bar_kv = bars + 1;       // [nqt] one per [k|v] box
  uint64_t* bar_s = bars + 1 + 8;    // [2] S buffer written by MMA1
  uint64_t* bar_o = bars + 1 + 8 + 2;  // MMA2 finished (O updated, P tile free)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp index, provably uniform
  const long long tok0 = (long long)b * T;
  unsigned long long* dbg = (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 32) ? dbg_buf : nullptr;
  dbg_stamp(dbg, 0);

  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv);
    mbar_init(bar_q, 1);
    for (int j = 0; j < nqt; ++j) mbar_init(&bar_kv[j], 1);
    mbar_init(&bar_s[0], 1);
    mbar_init(&bar_s[1], 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  dbg_stamp(dbg, 1);
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t tS[2] = {tmem_base, tmem_base + 64};
  const uint32_t tO = tmem_base + 128;

  constexpr uint32_t idesc1 = umma_idesc_bf16(128, 64, 0, 0);   // Q (K-major) x K (K-major)
  constexpr uint32_t idesc2 = umma_idesc_bf16(128, 64, 0, 1);   // P (K-major) x [K|V] (MN-major)

  // MMAs are issued by warp 0 as a whole (uniform control flow, one elected lane per instruction) so that the
  // descriptors stay in uniform registers
  auto issue_mma1 = [&](int kt) {   // S[kt & 1] = Q . K_kt^T     (warp 0)
    const uint32_t a = smem_u32(sQ);
    const uint32_t bb = smem_u32(sKV) + kt * (TC_KT * 128);
#pragma unroll
    for (int k = 0; k < 2; ++k)
      umma_bf16_1(tS[kt & 1], umma_smem_desc_sw128(a + 32 * k, 16, 1024), umma_smem_desc_sw128(bb + 32 * k, 16, 1024),
                  idesc1, k);
    umma_commit_1(&bar_s[kt & 1]);
  };

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, 16384);
    tma_load_2d(&tm_qkv, bar_q, sQ, head * 32, (int)(tok0 + qt * TC_QT));
    for (int j = 0; j < nkv_boxes; ++j) {
      mbar_arrive_expect_tx(&bar_kv[j], 16384);
      tma_load_2d(&tm_qkv, &bar_kv[j], sKV + j * 16384, kv_col, (int)(tok0 + j * 128));
    }
  }
  if (warp == 0) {
    mbar_wait(bar_q, 0);
    mbar_wait(&bar_kv[0], 0);
    tc_fence_after();
    issue_mma1(0);
  }

  const int qrow = qt * TC_QT + tid;                       // position of this thread's query in the sequence
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;   // TMEM lane quarter of this warp
  float m = -INFINITY, l = 0.f;

  for (int kt = 0; kt < nkt; ++kt) {
    if (warp == 0 && kt + 1 < nkt) {
      if (((kt + 1) & 1) == 0) mbar_wait(&bar_kv[(kt + 1) >> 1], 0);   // first use of a new [k|v] box
      tc_fence_after();
      issue_mma1(kt + 1);
    }
    mbar_wait(&bar_s[kt & 1], (kt >> 1) & 1);
    tc_fence_after();
    if (kt < 8) dbg_stamp(dbg, 2 + 3 * kt);
    uint32_t sraw[2][32];
    tmem_ld_32x32(tS[kt & 1] + lane_off, sraw[0]);
    tmem_ld_32x32(tS[kt & 1] + lane_off + 32, sraw[1]);
    tmem_ld_wait();
    float s[64];
    float mx = -INFINITY;
    const bool diag = (kt >= 2 * qt);                      // tiles that straddle the causal boundary
#pragma unroll
    for (int j = 0; j < 64; ++j) {
      float v = __uint_as_float(sraw[j >> 5][j & 31]) * scale_log2;
      if (diag && kt * TC_KT + j > qrow) v = -INFINITY;
      s[j] = v;
      mx = fmaxf(mx, v);
    }
    const float mn = fmaxf(m, mx);
    // rows of the upper half of the last tile see no key at all in it: keep them inert (mn stays finite from kt = 0)
    const float alpha = exp2f(m - mn);
    m = mn;
    l *= alpha;
    if (kt > 0) {
      mbar_wait(bar_o, (kt - 1) & 1);                      // MMA2(kt-1) done: O readable, P tile reusable
      tc_fence_after();
      if (kt < 8) dbg_stamp(dbg, 3 + 3 * kt);
      if (__any_sync(0xffffffffu, alpha != 1.0f)) {
        uint32_t ov[32];
        tmem_ld_32x32(tO + lane_off + 32, ov);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) ov[j] = __float_as_uint(__uint_as_float(ov[j]) * alpha);
        tmem_st_32x32(tO + lane_off + 32, ov);
        tmem_st_wait();
      }
    }
    // P = exp2(s - m) -> bf16, swizzled K-major row of the P tile (128 bytes = 8 x 16-byte pieces)
    uint8_t* prow = sP + tid * 128;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p0 = exp2f(s[c * 8 + 2 * e] - m), p1 = exp2f(s[c * 8 + 2 * e + 1] - m);
        w[e] = pack_bf16(p0, p1);
        l += bf16_lo(w[e]) + bf16_hi(w[e]);
      }
      *reinterpret_cast<uint4*>(prow + ((c ^ (tid & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (kt < 8) dbg_stamp(dbg, 4 + 3 * kt);
    if (warp == 0) {
      tc_fence_after();
      const uint32_t a = smem_u32(sP);
      const uint32_t bb = smem_u32(sKV) + kt * (TC_KT * 128);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_1(tO, umma_smem_desc_sw128(a + 32 * k, 16, 1024), umma_smem_desc_sw128(bb + 2048 * k, 8192, 1024),
                    idesc2, (kt | k) != 0 ? 1u : 0u);
      umma_commit_1(bar_o);
    }
  }
  mbar_wait(bar_o, (nkt - 1) & 1);
  tc_fence_after();
  uint32_t ov[32];
  tmem_ld_32x32(tO + lane_off + 32, ov);
  tmem_ld_wait();
  const float inv = 1.0f / l;
  __nv_bfloat16* orow = o + (tok0 + qrow) * ldo + head * 32;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 v;
    v.x = pack_bf16(__uint_as_float(ov[8 * c + 0]) * inv, __uint_as_float(ov[8 * c + 1]) * inv);
    v.y = pack_bf16(__uint_as_float(ov[8 * c + 2]) * inv, __uint_as_float(ov[8 * c + 3]) * inv);
    v.z = pack_bf16(__uint_as_float(ov[8 * c + 4]) * inv, __uint_as_float(ov[8 * c + 5]) * inv);
    v.w = pack_bf16(__uint_as_float(ov[8 * c + 6]) * inv, __uint_as_float(ov[8 * c + 7]) * inv);
    reinterpret_cast<uint4*>(orow)[c] = v;
  }
  lse[(tok0 + qrow) * Nq + head] = m + log2f(l);
  dbg_stamp(dbg, 30);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ---------------------------------------------------------------------------------------------------------------
// Four query heads per CTA (the GQA group of the shipped configs: Nq = 4 on one kv head).  CTA = (batch row b, 128-query
// tile qt); 512 threads = 4 softmax warpgroups, warpgroup h owns head h and runs the pipeline of the kernel above on its
// own Q tile, P tile, TMEM columns and barriers; the [k|v] boxes are loaded ONCE for the four heads.  16 softmax warps per
// SM instead of 8 (two single-head CTAs): the kernel is bound by the latency of the per-tile chain
// (MMA1 -> TMEM load -> max -> exp2 -> P tile -> MMA2), not by the tensor or SFU pipes, so the extra independent chains
// fill the idle issue slots.  TMEM: 4 x (S 64 + O 64) = 512 columns, so S is single buffered: S(kt+1) is issued as soon as
// the warpgroup holds S(kt) in registers, i.e. it still runs under the exponentials of tile kt.
// 2^x for x <= 0 on the FMA pipe: round-to-nearest split x = n + f, f in [-0.5, 0.5], cubic minimax for 2^f (relative error
// 7.5e-5, far inside the bf16 rounding of the probabilities), exponent patched in with one integer add.
__device__ __forceinline__ float ex2_poly3(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;                 // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.0551716685f, 0.2426111251f);
  p = fmaf(p, f, 0.6932609677f);
  p = fmaf(p, f, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

__device__ __forceinline__ void wg_sync(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }

__global__ void __launch_bounds__(512, 1)
attn_fwd_tc4_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ o, long long ldo,
                    float* __restrict__ lse, int T, int kv_col, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nqt = T / TC_QT;
  const int qt = nqt - 1 - (int)blockIdx.x;      // heavy (late) query tiles first
  const int b = blockIdx.y;
  const int nkt = 2 * (qt + 1);
  const int nkv_boxes = qt + 1;
  const int tid = threadIdx.x;
  const int wg = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 7), 0);          // = query head, warp uniform
  const int wwarp = __shfl_sync(0xffffffffu, (int)((threadIdx.x >> 5) & 3), 0);  // warp inside the warpgroup
  const int wt = tid & 127;                                                       // query row of the tile = TMEM lane
  uint8_t* sQ = smem + wg * 16384;               // 4 x 16 KB
  uint8_t* sP = smem + 65536 + wg * 16384;       // 4 x 16 KB
  uint8_t* sKV = smem + 131072;                  // nqt x 16 KB, shared by the four heads
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + (size_t)nqt * 16384);
  uint64_t* bar_kv = bars;                 // [4] one per [k|v] box
  uint64_t* bar_q = bars + 4 + wg;         // Q tile of this head landed
  uint64_t* bar_s = bars + 8 + wg;         // S written by MMA1
  uint64_t* bar_o = bars + 12 + wg;        // MMA2 finished (O updated, P tile free)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const long long tok0 = (long long)b * T;

  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv);
    for (int j = 0; j < 16; ++j) mbar_init(&bars[j], 1);
    fence_barrier_init();
  }
  if (tid < 32) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const uint32_t tS = tmem_base + wg * 128;
  const uint32_t tO = tS + 64;
  constexpr uint32_t idesc1 = umma_idesc_bf16(128, 64, 0, 0);   // Q (K-major) x K (K-major)
  constexpr uint32_t idesc2 = umma_idesc_bf16(128, 64, 0, 1);   // P (K-major) x [K|V] (MN-major)

  auto issue_mma1 = [&](int kt) {   // S = Q . K_kt^T   (warp 0 of the warpgroup, converged)
    const uint32_t a = smem_u32(sQ);
    const uint32_t bb = smem_u32(sKV) + kt * (TC_KT * 128);
#pragma unroll
    for (int k = 0; k < 2; ++k)
      umma_bf16_1(tS, umma_smem_desc_sw128(a + 32 * k, 16, 1024), umma_smem_desc_sw128(bb + 32 * k, 16, 1024), idesc1, k);
    umma_commit_1(bar_s);
  };

  if (tid == 0) {
    for (int j = 0; j < nkv_boxes; ++j) {
      mbar_arrive_expect_tx(&bar_kv[j], 16384);
      tma_load_2d(&tm_qkv, &bar_kv[j], sKV + j * 16384, kv_col, (int)(tok0 + j * 128));
    }
  }
  if (wt == 0) {
    mbar_arrive_expect_tx(bar_q, 16384);
    tma_load_2d(&tm_qkv, bar_q, sQ, wg * 32, (int)(tok0 + qt * TC_QT));
  }
  if (wwarp == 0) {
    mbar_wait(bar_q, 0);
    mbar_wait(&bar_kv[0], 0);
    tc_fence_after();
    issue_mma1(0);
  }

  const int qrow = qt * TC_QT + wt;
  const uint32_t lane_off = (uint32_t)(wwarp * 32) << 16;   // TMEM lane quarter of this warp (warp index % 4)
  float m = -INFINITY, l = 0.f;

  for (int kt = 0; kt < nkt; ++kt) {
    mbar_wait(bar_s, kt & 1);
    tc_fence_after();
    uint32_t sraw[2][32];
    tmem_ld_32x32(tS + lane_off, sraw[0]);
    tmem_ld_32x32(tS + lane_off + 32, sraw[1]);
    tmem_ld_wait();
    tc_fence_before();
    wg_sync(wg);                                   // the whole warpgroup holds S(kt): the accumulator may be overwritten
    if (wwarp == 0 && kt + 1 < nkt) {
      if (((kt + 1) & 1) == 0) mbar_wait(&bar_kv[(kt + 1) >> 1], 0);   // first use of a new [k|v] box
      tc_fence_after();
      issue_mma1(kt + 1);
    }
    // Instruction diet (the loop is issue-bound): the row maximum is taken on the RAW scores (scale > 0 commutes with
    // max), the scale and the subtraction of the maximum ride in one FFMA in front of ex2.approx, the row sum uses the
    // fp32 probabilities before their bf16 rounding.  Causal masking costs instructions on the diagonal tiles only.
    if (kt >= 2 * qt) {
#pragma unroll
      for (int j = 0; j < 64; ++j)
        if (kt * TC_KT + j > qrow) sraw[j >> 5][j & 31] = 0xff800000u;   // -inf
    }
    float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};       // four independent chains (fixed-latency stalls),
#pragma unroll                                                          // two scores per FMNMX3
    for (int j = 0; j < 64; j += 2)
      mx4[(j >> 1) & 3] = fmax3(mx4[(j >> 1) & 3], __uint_as_float(sraw[j >> 5][j & 31]), __uint_as_float(sraw[j >> 5][(j & 31) + 1]));
    const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
    const float mn = fmaxf(m, mx * scale_log2);
    const float alpha = ex2_approx(m - mn);
    m = mn;
    l *= alpha;
    if (kt > 0) {
      mbar_wait(bar_o, (kt - 1) & 1);              // MMA2(kt-1) done: O readable, P tile reusable
      tc_fence_after();
      if (__any_sync(0xffffffffu, alpha != 1.0f)) {
        uint32_t ov[32];
        tmem_ld_32x32(tO + lane_off + 32, ov);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) ov[j] = __float_as_uint(__uint_as_float(ov[j]) * alpha);
        tmem_st_32x32(tO + lane_off + 32, ov);
        tmem_st_wait();
      }
    }
    const float neg_m = -m;
    // packed f32x2 arithmetic: one FFMA2 scales two scores and subtracts the maximum, one FADD2 adds two probabilities to a
    // pair of partial row sums (four independent pair chains)
    const unsigned long long sc2 = f32x2_pack(scale_log2, scale_log2), nm2 = f32x2_pack(neg_m, neg_m);
    unsigned long long l4[4] = {0ull, 0ull, 0ull, 0ull};
    uint8_t* prow = sP + wt * 128;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float pv[8];
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        const int j = c * 8 + e;
        float x0, x1;
        f32x2_unpack(f32x2_fma(f32x2_pack(__uint_as_float(sraw[j >> 5][j & 31]), __uint_as_float(sraw[j >> 5][(j & 31) + 1])), sc2, nm2),
                     x0, x1);
        // every 4th exponential runs as a polynomial on the FMA pipe (the MUFU pipe is the busiest unit of this loop);
        // evaluating the two of a group together as f32x2 costs registers the 128-register budget does not have (spills:
        // 102 instead of 94 us)
        pv[e] = ex2_approx(x0);
        pv[e + 1] = (e & 3) == 2 ? ex2_poly3(x1) : ex2_approx(x1);
      }
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        w[e] = pack_bf16(pv[2 * e], pv[2 * e + 1]);
        l4[e] = f32x2_add(l4[e], f32x2_pack(pv[2 * e], pv[2 * e + 1]));
      }
      *reinterpret_cast<uint4*>(prow + ((c ^ (wt & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    {
      float a0, a1;
      f32x2_unpack(f32x2_add(f32x2_add(l4[0], l4[1]), f32x2_add(l4[2], l4[3])), a0, a1);
      l += a0 + a1;
    }
    fence_proxy_async_smem();
    tc_fence_before();
    wg_sync(wg);
    if (wwarp == 0) {
      tc_fence_after();
      const uint32_t a = smem_u32(sP);
      const uint32_t bb = smem_u32(sKV) + kt * (TC_KT * 128);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_1(tO, umma_smem_desc_sw128(a + 32 * k, 16, 1024), umma_smem_desc_sw128(bb + 2048 * k, 8192, 1024),
                    idesc2, (kt | k) != 0 ? 1u : 0u);
      umma_commit_1(bar_o);
    }
  }
  mbar_wait(bar_o, (nkt - 1) & 1);
  tc_fence_after();
  uint32_t ov[32];
  tmem_ld_32x32(tO + lane_off + 32, ov);
  tmem_ld_wait();
  const float inv = 1.0f / l;
  __nv_bfloat16* orow = o + (tok0 + qrow) * ldo + wg * 32;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 v;
    v.x = pack_bf16(__uint_as_float(ov[8 * c + 0]) * inv, __uint_as_float(ov[8 * c + 1]) * inv);
    v.y = pack_bf16(__uint_as_float(ov[8 * c + 2]) * inv, __uint_as_float(ov[8 * c + 3]) * inv);
    v.z = pack_bf16(__uint_as_float(ov[8 * c + 4]) * inv, __uint_as_float(ov[8 * c + 5]) * inv);
    v.w = pack_bf16(__uint_as_float(ov[8 * c + 6]) * inv, __uint_as_float(ov[8 * c + 7]) * inv);
    reinterpret_cast<uint4*>(orow)[c] = v;
  }
  lse[(tok0 + qrow) * 4 + wg] = m + log2f(l);
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_base, 512);
}

int g_attn_train_impl = 0;   // 0 = tcgen05 forward when the shape allows (four heads per CTA when Nq == 4), 1 = mma.sync
                             // everywhere, 2 = tcgen05 with one head per CTA (the round-1 kernel)

// Returns MPX_ERR_UNSUPPORTED (without setting an error) when the shape is not covered; the caller falls back.
int attn_fwd_tc_launch(const mpx_attn_train_args* a, cudaStream_t stream) {
  const long long QD = (long long)a->Nq * 32;
  const bool packed = a->Nkv == 1 && a->D == 32 && a->T % TC_QT == 0 && a->T / TC_QT <= 8 && a->ldq == a->ldk &&
                      a->ldk == a->ldv && a->k == a->q + QD && a->v == a->k + 32 && a->ldq >= QD + 64 &&
                      (a->ldq % 8) == 0 && (a->ldo % 8) == 0;
  if (!packed || g_attn_train_impl == 1) return MPX_ERR_UNSUPPORTED;
  CUtensorMap tm;
  const long long M = (long long)a->B * a->T;
  int rc = make_tmap_bf16_2d(&tm, a->q, (uint64_t)M, (uint64_t)a->ldq, (uint64_t)a->ldq, 64, 128);
  if (rc) return rc;
  const int nqt = a->T / TC_QT;
  const float scale_log2 = (1.0f / sqrtf((float)a->D)) * 1.4426950408889634f;
  if (a->Nq == 4 && nqt <= 4 && g_attn_train_impl == 0) {
    const size_t smem4 = 1024 + 131072 + (size_t)nqt * 16384 + 256;
    static size_t attr4 = 0;
    if (smem4 > attr4) {
      cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4);
      if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(attn_fwd_tc4): %s", cudaGetErrorString(e));
        return MPX_ERR_CUDA;
      }
      attr4 = smem4;
    }
    attn_fwd_tc4_kernel<<<dim3(nqt, a->B), 512, smem4, stream>>>(tm, reinterpret_cast<__nv_bfloat16*>(a->o), a->ldo, a->lse,
                                                                 a->T, (int)QD, scale_log2);
    return check_launch("attn_fwd_tc4_kernel");
  }
  const size_t smem = 1024 + 32768 + (size_t)nqt * 16384 + 256;
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attn_fwd_tc): %s", cudaGetErrorString(e));
      return MPX_ERR_CUDA;
    }
    attr = smem;
  }
  dim3 grid(nqt, a->Nq, a->B);
  attn_fwd_tc_kernel<<<grid, 128, smem, stream>>>(tm, reinterpret_cast<__nv_bfloat16*>(a->o), a->ldo, a->lse, a->T, a->Nq,
                                                 (int)QD, scale_log2, g_dbg_stamps);
  return check_launch("attn_fwd_tc_kernel");
}

}  // namespace mpx
