// PPO-update attention BACKWARD on 5th-generation tensor cores: the five contractions of the flash-attention
// transpose (autodiff of the training branch of AttentionBlock.__call__, reference mapox_trainer/model/attention.py:
// 151-169) issued as tcgen05.mma with every accumulator in TMEM.  Same arithmetic as attn_bwd_kernel in
// attn_train.cu (fp32 P and dS from the saved base-2 log-sum-exp and delta = rowsum(dO*O); P and dS rounded to bf16
// before they feed the next contraction; fp32 accumulation).
//
// CTA = (batch row b, 128-key tile j); it owns dK_j and dV_j (accumulated in TMEM over the 4 query heads of the GQA
// group and every 128-query tile i >= j) and adds its dQ contribution to the fp32 accumulator with red.add (finished
// by attn_bwd_dq_kernel, which also applies the inverse RoPE).  Per (i, h) iteration, the key tile in two 64-key halves:
//   S  [q x 64] = Q_h  . K_half^T    A = Q box  (K-major, K = 32, SW64)  B = [k|v] box rows of the half (K-major, k part)
//   dP [q x 64] = dO_h . V_half^T    A = dO box (K-major, K = 32, SW64)  B = [k|v] box (K-major, +64 B: the v part)
//   P = exp2(S*scale*log2e - lse), dS = P * (dP - delta) * scale         (128 threads, thread = query row)
// and, once both halves of the P / dS tiles are written,
//   dV [k x d] += P^T  . dO_h        A = P  tile (MN-major: M = keys)    B = dO box (MN-major, N = 32, SW64)
//   dK [k x d] += dS^T . Q_h         A = dS tile (MN-major)              B = Q box  (MN-major, N = 32, SW64)
//   dQ [q x d]  = dS   . K_j         A = dS tile (K-major, K = 128 keys) B = [k|v] box (MN-major, N = 32: the k part)
// Two CTAs are resident per SM.  The per-iteration chain of one CTA is serial and every CTA pays ~6 us of prologue /
// epilogue, so ONE resident CTA (round 1: 320 threads, 148 KB, 512 TMEM columns, full-tile S and dP) left the SM idle for
// half of the time: 217 us per layer-minibatch.  A CTA now fits in half an SM -
//   shared memory 112 KB: [k|v] 16 KB; Q and dO as [128 x 32] SWIZZLE_64B boxes (8 KB each, double buffered: the old 64-column
//                         SWIZZLE_128B boxes carried 32 unused columns); P and dS tiles 2 x 32 KB
//   TMEM 256 columns:     S | dP of ONE 64-key half (64 + 64), dV, dK, dQ (3 x 32)
//   192 threads:          warps 0-3 softmax / epilogue (thread = query row, the 64 keys of the half), warp 4 TMA producer,
//                         warp 5 MMA issuer / TMEM owner; <= 168 registers per thread
// - S / dP of half 1 are issued as soon as the softmax warps hold half 0 in registers, half 0 of the next iteration after
// half 1 (in front of this iteration's 24 MMAs if its boxes have landed, behind them otherwise), and the Q / dO stage is
// released after the 16 dV / dK MMAs, before the 8 dQ ones.  CTAs are ordered longest-first (key tile 0 of every batch
// row, then tile 1, ...) so the short ones fill the tail.  ~180 us per layer-minibatch; what bounds it now is the tensor
// core's operand fetch: 32 small MMAs per iteration (N = 32 / 64) each re-read a 4 KB A operand from shared memory
// (measured: ~30-55 ns per MMA, issue-to-issue), and the red.add traffic of dQ (35 us; profiles/r02_attn_train_kernels.md).
// Requires Nkv == 1, packed q|k|v rows, head_dim 32, T % 128 == 0, T <= 1024.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.h"

namespace mpx {

extern int g_attn_train_impl;

constexpr int BT = 128;            // queries per tile == keys per tile

struct AttnBwdTcParams {
  const float* lse;                // (M, Nq) base-2 log-sum-exp (scale folded in)
  const float* delta;              // (M, Nq)
  float* dq_acc;                   // (M, Nq*32) fp32
  __nv_bfloat16* dk;
  long long lddk;
  __nv_bfloat16* dv;
  long long lddv;
  const int32_t* pos;
  int pos_len;
  const float* timescale;
  int T, Nq, kv_col;
  float scale, scale_log2;
  unsigned long long* dbg;
};

constexpr int B2_SM_WARPS = 4;
constexpr int B2_THREADS = (B2_SM_WARPS + 2) * 32;
constexpr int B2_SMEM = 16384 * 7 + 128;

__device__ __forceinline__ uint64_t umma_smem_desc_sw64(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;                    // SWIZZLE_64B
  return d;
}

__global__ void __launch_bounds__(B2_THREADS, 2)
attn_bwd_tc2_kernel(const AttnBwdTcParams p, const __grid_constant__ CUtensorMap tm_kv,
                    const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sKV = smem;                       // 16 KB  [128 keys x (k | v)], SWIZZLE_128B
  uint8_t* sQ = smem + 16384;                // 2 x 8 KB  [128 q x 32], SWIZZLE_64B
  uint8_t* sDO = smem + 16384 * 2;           // 2 x 8 KB
  uint8_t* sP = smem + 16384 * 3;            // 32 KB  [128 q x 128 keys] bf16: 2 boxes of 64 keys, SWIZZLE_128B
  uint8_t* sDS = smem + 16384 * 5;           // 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 16384 * 7);
  uint64_t* kv_full = bars;
  uint64_t* in_full = bars + 1;              // [2] Q + dO boxes of an iteration landed
  uint64_t* in_empty = bars + 3;             // [2] all MMAs reading them are done
  uint64_t* s_full = bars + 5;               // S and dP of a 64-key half written
  uint64_t* s_empty = bars + 6;              // ... and read by the softmax warps
  uint64_t* p_full = bars + 7;               // P and dS tiles (both halves) written
  uint64_t* p_empty = bars + 8;              // dV/dK/dQ MMAs done: tiles reusable, dQ readable
  uint64_t* dq_empty = bars + 9;             // dQ accumulator read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int nt = p.T / BT;
  const int nb = (int)gridDim.x / nt;
  const int kt = (int)blockIdx.x / nb;        // longest first: key tile 0 sees every query tile
  const int b = (int)blockIdx.x % nb;
  const long long tok0 = (long long)b * p.T;
  const int n_iter = (nt - kt) * p.Nq;        // (query tile i = kt .. nt-1) x heads

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();     // the swizzled tiles assume a 1024-byte aligned window
    tma_prefetch_desc(&tm_kv);
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_do);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&in_full[s], 1);
      mbar_init(&in_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, B2_SM_WARPS);
    mbar_init(p_full, B2_SM_WARPS);
    mbar_init(p_empty, 1);
    mbar_init(dq_empty, B2_SM_WARPS);
    fence_barrier_init();
  }
  if (warp == B2_SM_WARPS + 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t tS = tmem_base, tDP = tmem_base + 64, tDV = tmem_base + 128, tDK = tmem_base + 160, tDQ = tmem_base + 192;

  if (warp == B2_SM_WARPS) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 16384);
      tma_load_2d(&tm_kv, kv_full, sKV, p.kv_col, (int)(tok0 + kt * BT));
      for (int it = 0; it < n_iter; ++it) {
        const int s = it & 1;
        const int i = kt + it / p.Nq, h = it % p.Nq;
        mbar_wait(&in_empty[s], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&in_full[s], 16384);
        tma_load_2d(&tm_q, &in_full[s], sQ + s * 8192, h * 32, (int)(tok0 + i * BT));
        tma_load_2d(&tm_do, &in_full[s], sDO + s * 8192, h * 32, (int)(tok0 + i * BT));
      }
    }
  } else if (warp == B2_SM_WARPS + 1) {
    // ================================ MMA issuer ================================
    // warp-uniform control flow, one elected lane per tcgen05 instruction (see the note in the file header of gemm_tc.cu)
    constexpr uint32_t id_s = umma_idesc_bf16(128, 64, 0, 0);     // K-major x K-major, 64 keys
    constexpr uint32_t id_kd = umma_idesc_bf16(128, 32, 1, 1);    // MN-major x MN-major  (dV, dK)
    constexpr uint32_t id_q = umma_idesc_bf16(128, 32, 0, 1);     // K-major x MN-major   (dQ)
    mbar_wait(kv_full, 0);
    const uint32_t kv = smem_u32(sKV);
    auto mma_s = [&](int u, bool wait_in) {                   // S and dP of sub-iteration u = 2 * it + half
      const int it = u >> 1, half = u & 1, s = it & 1;
      const uint32_t q = smem_u32(sQ + s * 8192), d_o = smem_u32(sDO + s * 8192);
      if (wait_in) mbar_wait(&in_full[s], (it >> 1) & 1);
      mbar_wait(s_empty, (u & 1) ^ 1);                        // sub-iteration u-1 has been pulled out of TMEM
      tc_fence_after();
      const uint32_t kh = kv + half * 8192;                   // keys 64*half .. +63 of the tile
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        umma_bf16_1(tS, umma_smem_desc_sw64(q + 32 * k, 16, 512), umma_smem_desc_sw128(kh + 32 * k, 16, 1024), id_s, k);
        umma_bf16_1(tDP, umma_smem_desc_sw64(d_o + 32 * k, 16, 512), umma_smem_desc_sw128(kh + 64 + 32 * k, 16, 1024), id_s, k);
      }
      umma_commit_1(s_full);
    };
    mma_s(0, true);
    for (int it = 0; it < n_iter; ++it) {
      const int s = it & 1;
      const uint32_t q = smem_u32(sQ + s * 8192), d_o = smem_u32(sDO + s * 8192);
      mma_s(2 * it + 1, false);
      // half 0 of the next iteration goes in FRONT of this iteration's 24 MMAs when its Q / dO boxes have already landed
      // (their stage was released by the MMAs of iteration it-1); otherwise it must not hold those MMAs back
      bool early = false;
      if (it + 1 < n_iter) {
        mbar_wait(s_empty, 1);                                // half 1 (an odd sub-iteration) pulled; always precedes p_full
        early = mbar_try_wait(&in_full[s ^ 1], ((it + 1) >> 1) & 1);
        if (early) mma_s(2 * it + 2, false);
      }
      mbar_wait(p_full, it & 1);                              // P and dS tiles of this iteration are in smem
      mbar_wait(dq_empty, (it & 1) ^ 1);                      // previous dQ has been drained
      tc_fence_after();
      const uint32_t pp = smem_u32(sP), ds = smem_u32(sDS);
#pragma unroll
      for (int k = 0; k < 8; ++k) {                           // contraction over the 128 queries: 16 rows per step
        umma_bf16_1(tDV, umma_smem_desc_sw128(pp + 2048 * k, 16384, 1024), umma_smem_desc_sw64(d_o + 1024 * k, 4096, 512),
                    id_kd, (it | k) != 0 ? 1u : 0u);
        umma_bf16_1(tDK, umma_smem_desc_sw128(ds + 2048 * k, 16384, 1024), umma_smem_desc_sw64(q + 1024 * k, 4096, 512),
                    id_kd, (it | k) != 0 ? 1u : 0u);
      }
      umma_commit_1(&in_empty[s]);                            // Q / dO of this stage are free: the loads two iterations ahead start
#pragma unroll
      for (int k = 0; k < 8; ++k)                             // contraction over the 128 keys: two 64-key boxes
        umma_bf16_1(tDQ, umma_smem_desc_sw128(ds + (k >> 2) * 16384 + 32 * (k & 3), 16, 1024),
                    umma_smem_desc_sw128(kv + 2048 * k, 8192, 1024), id_q, k != 0 ? 1u : 0u);
      umma_commit_1(p_empty);
      if (it + 1 < n_iter && !early) mma_s(2 * it + 2, true);
    }
  } else {
    // ================================ softmax / epilogue warps ================================
    const int r = warp * 32 + lane;                          // query row of the tile == TMEM lane (key row at the end)
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    unsigned long long* dbg = (blockIdx.x == 0 && threadIdx.x == 0) ? p.dbg : nullptr;
    dbg_stamp(dbg, 0);
    auto drain_dq = [&](int it_done) {                       // my row of dQ(it_done) -> fp32 accumulator
      const int pi = kt + it_done / p.Nq, ph = it_done % p.Nq;
      uint32_t dq[32];
      tmem_ld_32x32(tDQ + lane_off, dq);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_empty);
      float* dst = p.dq_acc + (tok0 + (long long)pi * BT + r) * (long long)(p.Nq * 32) + ph * 32;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * c), "f"(__uint_as_float(dq[4 * c])),
                     "f"(__uint_as_float(dq[4 * c + 1])), "f"(__uint_as_float(dq[4 * c + 2])),
                     "f"(__uint_as_float(dq[4 * c + 3]))
                     : "memory");
    };
    float lse_n = p.lse[(tok0 + (long long)kt * BT + r) * p.Nq], dl_n = p.delta[(tok0 + (long long)kt * BT + r) * p.Nq];
    for (int it = 0; it < n_iter; ++it) {
      const int i = kt + it / p.Nq;
      const float lse_r = lse_n, dl_r = dl_n;
      if (it + 1 < n_iter) {
        const int i2 = kt + (it + 1) / p.Nq, h2 = (it + 1) % p.Nq;
        const long long q2 = tok0 + (long long)i2 * BT + r;
        lse_n = p.lse[q2 * p.Nq + h2];
        dl_n = p.delta[q2 * p.Nq + h2];
      }
      const bool diag = (i == kt);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        mbar_wait(s_full, half);                              // sub-iteration 2 * it + half
        tc_fence_after();
        if (it < 6) dbg_stamp(dbg, 1 + 10 * it + 5 * half);
        uint32_t sv[2][32], dpv[2][32];
        tmem_ld_32x32(tS + lane_off, sv[0]);
        tmem_ld_32x32(tDP + lane_off, dpv[0]);
        tmem_ld_32x32(tS + lane_off + 32, sv[1]);
        tmem_ld_32x32(tDP + lane_off + 32, dpv[1]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty);                  // S / dP are in registers: the next half may be issued
        uint32_t pw[32], dw[32];
        // packed f32x2 arithmetic, two keys per instruction: x = fma(s, scale_log2, -lse), dS = (P * (dP - delta)) * scale
        const unsigned long long sc2 = f32x2_pack(p.scale_log2, p.scale_log2), nl2 = f32x2_pack(-lse_r, -lse_r);
        const unsigned long long nd2 = f32x2_pack(-dl_r, -dl_r), s2 = f32x2_pack(p.scale, p.scale);
#pragma unroll
        for (int cc = 0; cc < 2; ++cc)
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float x0, x1;
            f32x2_unpack(f32x2_fma(f32x2_pack(__uint_as_float(sv[cc][2 * e]), __uint_as_float(sv[cc][2 * e + 1])), sc2, nl2), x0, x1);
            float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
            if (diag) {
              if (64 * half + 32 * cc + 2 * e > r) p0 = 0.f;
              if (64 * half + 32 * cc + 2 * e + 1 > r) p1 = 0.f;
            }
            pw[16 * cc + e] = pack_bf16(p0, p1);
            const unsigned long long t = f32x2_add(f32x2_pack(__uint_as_float(dpv[cc][2 * e]), __uint_as_float(dpv[cc][2 * e + 1])), nd2);
            float d0, d1;
            f32x2_unpack(f32x2_mul(f32x2_mul(f32x2_pack(p0, p1), t), s2), d0, d1);
            dw[16 * cc + e] = pack_bf16(d0, d1);
          }
        if (it < 6) dbg_stamp(dbg, 2 + 10 * it + 5 * half);
        if (half == 0) {
          mbar_wait(p_empty, (it & 1) ^ 1);                   // MMAs of iteration it-1 are done: tiles free, dQ(it-1) readable
          tc_fence_after();
          if (it < 6) dbg_stamp(dbg, 3 + 10 * it);
        }
        // key columns [64*half, +64) = box `half`, all 8 16-byte chunks of my row (128-byte swizzle)
        const uint32_t prow = smem_u32(sP) + half * 16384 + r * 128, dsrow = smem_u32(sDS) + half * 16384 + r * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint32_t off = (ch ^ (r & 7)) << 4;
          sts_v4(prow + off, make_uint4(pw[4 * ch], pw[4 * ch + 1], pw[4 * ch + 2], pw[4 * ch + 3]));
          sts_v4(dsrow + off, make_uint4(dw[4 * ch], dw[4 * ch + 1], dw[4 * ch + 2], dw[4 * ch + 3]));
        }
        if (it < 6) dbg_stamp(dbg, 4 + 10 * it + 5 * half);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (it > 0) drain_dq(it - 1);
      if (it < 6) dbg_stamp(dbg, 5 + 10 * it);
    }
    // ---- last dQ, then dK and dV of this key tile (thread = key row)
    mbar_wait(p_empty, (n_iter & 1) ^ 1);
    tc_fence_after();
    drain_dq(n_iter - 1);
    const long long ktok = tok0 + (long long)kt * BT + r;
    {
      uint32_t dkv[32];
      tmem_ld_32x32(tDK + lane_off, dkv);
      tmem_ld_wait();
      // inverse RoPE on dK (positional_embeddings.py:14-27 transposed): pairs (d, d + 16)
      const float pf = (float)p.pos[ktok % p.pos_len];
      uint32_t klo[8], khi[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float o_lo[2], o_hi[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int d = 2 * e + u;
          float sn, cs;
          rope_sincos(pf / p.timescale[d], &sn, &cs);
          const float a = __uint_as_float(dkv[d]), bb = __uint_as_float(dkv[d + 16]);
          o_lo[u] = a * cs + bb * sn;
          o_hi[u] = bb * cs - a * sn;
        }
        klo[e] = pack_bf16(o_lo[0], o_lo[1]);
        khi[e] = pack_bf16(o_hi[0], o_hi[1]);
      }
      uint4* dkr = reinterpret_cast<uint4*>(p.dk + ktok * p.lddk);
      dkr[0] = make_uint4(klo[0], klo[1], klo[2], klo[3]);
      dkr[1] = make_uint4(klo[4], klo[5], klo[6], klo[7]);
      dkr[2] = make_uint4(khi[0], khi[1], khi[2], khi[3]);
      dkr[3] = make_uint4(khi[4], khi[5], khi[6], khi[7]);
    }
    {
      uint32_t dvv[32], vw[16];
      tmem_ld_32x32(tDV + lane_off, dvv);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) vw[e] = pack_bf16(__uint_as_float(dvv[2 * e]), __uint_as_float(dvv[2 * e + 1]));
      uint4* dvr = reinterpret_cast<uint4*>(p.dv + ktok * p.lddv);
#pragma unroll
      for (int c = 0; c < 4; ++c) dvr[c] = make_uint4(vw[4 * c], vw[4 * c + 1], vw[4 * c + 2], vw[4 * c + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == B2_SM_WARPS + 1) tmem_dealloc(tmem_base, 256);
}

// Returns MPX_ERR_UNSUPPORTED (without setting an error) when the shape is not covered; the caller falls back.
int attn_bwd_tc_launch(const mpx_attn_train_bwd_args* a, cudaStream_t stream) {
  const long long QD = (long long)a->Nq * 32;
  const bool packed = a->Nkv == 1 && a->D == 32 && a->T % BT == 0 && a->T <= 1024 && a->ldq == a->ldk && a->ldk == a->ldv &&
                      a->k == a->q + QD && a->v == a->k + 32 && a->ldq >= QD + 64 && (a->ldq % 8) == 0 &&
                      (a->lddo % 8) == 0 && a->lddo >= QD && (a->lddk % 8) == 0 && (a->lddv % 8) == 0 &&
                      (reinterpret_cast<uintptr_t>(a->dk) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->dv) & 15) == 0;
  if (!packed || g_attn_train_impl != 0) return MPX_ERR_UNSUPPORTED;
  const long long M = (long long)a->B * a->T;
  CUtensorMap tm_kv, tm_q, tm_do;
  int rc = make_tmap_bf16_2d(&tm_kv, a->q, (uint64_t)M, (uint64_t)a->ldq, (uint64_t)a->ldq, 64, 128);        // [k | v] boxes
  if (rc) return rc;
  rc = make_tmap_bf16_2d_sw64(&tm_q, a->q, (uint64_t)M, (uint64_t)a->ldq, (uint64_t)a->ldq, 32, 128);  // one head of q
  if (rc) return rc;
  rc = make_tmap_bf16_2d_sw64(&tm_do, a->dout, (uint64_t)M, (uint64_t)QD, (uint64_t)a->lddo, 32, 128);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_tc2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attn_bwd_tc2): %s", cudaGetErrorString(e));
      return MPX_ERR_CUDA;
    }
    attr_set = true;
  }
  AttnBwdTcParams p;
  p.lse = a->lse; p.delta = a->delta; p.dq_acc = a->dq_acc;
  p.dk = reinterpret_cast<__nv_bfloat16*>(a->dk); p.lddk = a->lddk;
  p.dv = reinterpret_cast<__nv_bfloat16*>(a->dv); p.lddv = a->lddv;
  p.pos = a->pos; p.pos_len = a->pos_len; p.timescale = a->rope_timescale;
  p.T = a->T; p.Nq = a->Nq; p.kv_col = (int)QD;
  p.scale = 1.0f / sqrtf((float)a->D);
  p.scale_log2 = p.scale * 1.4426950408889634f;
  p.dbg = g_dbg_stamps;
  attn_bwd_tc2_kernel<<<dim3((unsigned)((a->T / BT) * a->B)), B2_THREADS, B2_SMEM, stream>>>(p, tm_kv, tm_q, tm_do);
  return check_launch("attn_bwd_tc2_kernel");
}

}  // namespace mpx
