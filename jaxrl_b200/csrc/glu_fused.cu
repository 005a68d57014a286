// Fused GLU feed-forward block (reference GLUBlock.__call__, mapox_trainer/model/feed_forward.py:73-78, plus the
// residual add of TransformerBlock, network.py:167-170):
//     out = bf16( residual + bf16( bf16( bf16(gelu(bf16(x@Wup))) * bf16(x@Wgate) ) @ Wdown ) )
// in ONE kernel: the (M, F) hidden activation never leaves the SM.  Two chained tcgen05 GEMMs per 64-wide chunk of
// the hidden dimension (FlashAttention-style): MMA1 writes [u_c | g_c] (128 x 128 fp32) to TMEM, the epilogue warps
// turn it into the bf16 h_c tile in shared memory (SW128 K-major, the A operand of MMA2), MMA2 accumulates
// h_c @ Wdown_c^T into a second TMEM accumulator.  MMA1 of chunk c+1 overlaps the activation math of chunk c.
//
// Roles (352 threads): warp 0 TMA producer (x tile + weight ring), warp 1 issues MMA1 (and owns TMEM), warp 10 issues
// MMA2, warps 2-9 activation + output epilogue (two warps per TMEM lane quarter; thread = row, each warp half of the
// columns).  Two issuing warps: in one thread's program order MMA1 of chunk c+1 would queue behind "h of chunk c-1 is
// written", although it only needs the TMEM buffer that chunk c-1's activation warps drained at their start.
// H (model width) must be 128; F % 64 == 0.
//
// Rollout batches are small (M = agents, 32 row tiles at 4096 agents), so one CTA per row tile would leave most
// of the 148 SMs idle while each CTA walks a serial MMA1 -> activation -> MMA2 chain over all F/64 chunks.  The
// kernel therefore runs as a thread-block cluster of CL CTAs per row tile: CTA k of the cluster owns chunks
// [k*F/64/CL, (k+1)*F/64/CL), leaves its fp32 partial of the output tile in its own shared memory, and the CL
// partials are summed over distributed shared memory - CTA k finishes output columns [k*128/CL, (k+1)*128/CL).
//
// Last layer of a rollout step (mpx_glu_fused_head): the cluster of 4 reduces BY ROWS instead - CTA k receives the four
// partials of rows [32k, 32k+32) with all 128 columns, adds x1 (every CTA holds the whole x1 tile), stores the block output
// and, holding complete rows, runs the policy head on them (output RMSNorm, tied logits, mask, Gumbel arg-max, log-prob:
// policy_head.cuh, 8 lanes per row = the 256 epilogue threads) - the separate policy_head_sample launch disappears.
#include "common.cuh"
#include "policy_head.cuh"
#include "ptx.cuh"
#include "tmap.h"

namespace mpx {

constexpr int GF_STAGES = 4;        // weight ring depth of the cluster / projection variants
constexpr int GF_MAX_STAGES = 10;   // single-CTA variant: the ring also takes the (then unused) partial and x1 regions
constexpr int GF_STAGE_BYTES = 16384;   // [128 rows x 64 cols] bf16, SW128
constexpr int GF_THREADS = 352;
constexpr int GF_EPI_WARPS = 8;
constexpr int GF_PART_BYTES = 65536;    // fp32 partial of the 128 x 128 output tile, [32 float4 columns][128 rows]

int g_glu_fused_cluster = 0;   // mpx_debug_set(3, v): 0 = auto, 1/2/4 = force that cluster size (tests)

struct GluFusedParams {
  int M, F;
  const __nv_bfloat16* residual;
  __nv_bfloat16* out;
  const float* norm_scale;        // (128) fp32: RMSNorm applied to the x tile in shared memory before MMA1; or null
  float eps;
  int head_on;                    // 1: cluster of 4 with the projection prologue; the partials are reduced BY ROWS and the owner
  PolicyHeadDev head;             //    of a row finishes it and runs the policy head on it (last layer of a rollout step)
  unsigned long long* dbg;
};

__device__ __forceinline__ void gf_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// One instantiation per mode - PROJ: out-projection prologue; CLUSTER: F split over a cluster of 2 or 4 CTAs; HOUT: h kept
// for the weight gradients (training); HEAD: policy-head epilogue (last layer of a rollout step).  The single runtime-mode
// kernel was 87 KB of SASS: a latency-bound launch pays for every instruction line it has to fetch (the rollout alternates
// five kernels, so little of it stays in the instruction caches) - measured 1.3 us per launch between the kernel with and
// without the head code compiled in.
template <bool PROJ, bool CLUSTER, bool HOUT, bool HEAD>
__global__ void __launch_bounds__(GF_THREADS, 1)
glu_fused_kernel(const GluFusedParams p, const __grid_constant__ CUtensorMap tm_x,
                 const __grid_constant__ CUtensorMap tm_wug, const __grid_constant__ CUtensorMap tm_wd,
                 const __grid_constant__ CUtensorMap tm_wo, const __grid_constant__ CUtensorMap tm_h) {
  extern __shared__ uint8_t smem_raw[];
  // the dynamic window starts at the same shared::cta offset in every CTA of the cluster, so the aligned
  // carve-up below is identical across the cluster (mapa relies on that)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;                         // 2 x 16 KB: x tile, K blocks 0 and 1
  uint8_t* sH = smem + 32768;                 // 2 x 16 KB: h chunk double buffer
  uint8_t* sW = smem + 65536;                 // GF_STAGES x 16 KB weight ring
  uint8_t* sP = sW + GF_STAGES * GF_STAGE_BYTES;   // fp32 partial (cluster reduction staging)
  uint8_t* sX1 = sP + GF_PART_BYTES;          // 32 KB: x1 rows (bf16, 256 B each, 16-byte chunks XOR-swizzled by row) [proj]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX1 + 32768);
  uint64_t* w_full = bars;                    // [GF_MAX_STAGES]
  uint64_t* w_empty = bars + GF_MAX_STAGES;   // [GF_MAX_STAGES]
  uint64_t* x_full = bars + 2 * GF_MAX_STAGES;
  uint64_t* x_empty = x_full + 1;
  uint64_t* ug_full = x_full + 2;             // [2]
  uint64_t* ug_empty = x_full + 4;            // [2]
  uint64_t* h_full = x_full + 6;              // [2]
  uint64_t* h_empty = x_full + 8;             // [2]
  uint64_t* d_full = x_full + 10;
  uint64_t* d_empty = x_full + 11;
  uint64_t* part_full = x_full + 12;          // all partials of the cluster written (remote arrivals)
  uint64_t* part_empty = x_full + 13;         // all readers of MY partial done (remote arrivals)
  uint64_t* xn_ready = x_full + 14;           // x tile landed and (optionally) normalised in place
  uint64_t* d0_full = x_full + 15;            // [proj] x @ Wo^T accumulated in the tD columns
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_full + 16);
  float* sSq = reinterpret_cast<float*>(x_full + 18);   // 2 x 128 partial sums of squares (RMSNorm prologue)
  float* sScale = sSq + 256;                            // norm scale (128)

  // shfl from a constant lane makes the warp index provably warp-uniform for the compiler: the role branches below and
  // everything computed inside them can live on the uniform datapath (UTCHMMA takes uniform-register descriptors)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int ncta = CLUSTER ? (int)cluster_nctarank() : 1, krank = CLUSTER ? (int)cluster_ctarank() : 0;
  if (CLUSTER) __builtin_assume(ncta > 1);
  const int nstages = (ncta == 1 && !PROJ) ? GF_MAX_STAGES : GF_STAGES;
  const int nchunks = p.F / 64 / ncta;        // chunks of the hidden dimension owned by this CTA
  const int chunk0 = krank * nchunks;
  const int m_tiles = (p.M + 127) / 128;

  unsigned long long* dbg = (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 64) ? p.dbg : nullptr;   // warp 2 lane 0
  dbg_stamp(dbg, 0);
  griddep_launch();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_wug);
    tma_prefetch_desc(&tm_wd);
    if (PROJ) tma_prefetch_desc(&tm_wo);
    for (int s = 0; s < GF_MAX_STAGES; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&ug_full[b], 1);
      mbar_init(&ug_empty[b], GF_EPI_WARPS);
      mbar_init(&h_full[b], GF_EPI_WARPS);
      mbar_init(&h_empty[b], 1);
    }
    mbar_init(d_full, 1);
    mbar_init(d_empty, GF_EPI_WARPS);
    mbar_init(part_full, 1);                   // one arrive.expect_tx per tile; the pushed bytes complete it
    mbar_init(part_empty, GF_EPI_WARPS * ncta);
    mbar_init(xn_ready, GF_EPI_WARPS);
    mbar_init(d0_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // the norm scale is a parameter (written by the optimizer / weight staging, never by the kernel right before this
  // one), so it and the CTA / cluster set-up barriers run under the preceding kernel's tail
  if (p.norm_scale && threadIdx.x >= 64 && threadIdx.x < 96)
    reinterpret_cast<float4*>(sScale)[threadIdx.x - 64] = reinterpret_cast<const float4*>(p.norm_scale)[threadIdx.x - 64];
  tc_fence_before();
  __syncthreads();
  if (ncta > 1) cluster_sync_all();            // remote CTAs' barriers are initialised before anyone arrives on them
  tc_fence_after();
  griddep_wait();                              // activations read from global memory below may come from earlier kernels
  dbg_stamp(dbg, 1);
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t tUG[2] = {tmem_base, tmem_base + 128};
  const uint32_t tD = tmem_base + 256;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t wphase = 0, xphase = 0;
      auto load_w = [&](const CUtensorMap* tm, int c0, int c1) {
        mbar_wait(&w_empty[stage], wphase ^ 1);
        mbar_arrive_expect_tx(&w_full[stage], GF_STAGE_BYTES);
        tma_load_2d(tm, &w_full[stage], sW + stage * GF_STAGE_BYTES, c0, c1);
        if (++stage == nstages) { stage = 0; wphase ^= 1; }
      };
      for (int tile = blockIdx.y; tile < m_tiles; tile += gridDim.y) {
        mbar_wait(x_empty, xphase ^ 1);
        mbar_arrive_expect_tx(x_full, 32768);
        tma_load_2d(&tm_x, x_full, sX, 0, tile * 128);
        tma_load_2d(&tm_x, x_full, sX + 16384, 64, tile * 128);
        xphase ^= 1;
        // weight order == MMA issue order: [Wo K blocks], UG_0, then for each c: UG_{c+1} (if any), D_c
        if (PROJ) {
          load_w(&tm_wo, 0, 0);
          load_w(&tm_wo, 64, 0);
        }
        load_w(&tm_wug, 0, chunk0 * 128);
        load_w(&tm_wug, 64, chunk0 * 128);
        for (int c = 0; c < nchunks; ++c) {
          if (c + 1 < nchunks) {
            load_w(&tm_wug, 0, (chunk0 + c + 1) * 128);
            load_w(&tm_wug, 64, (chunk0 + c + 1) * 128);
          }
          load_w(&tm_wd, (chunk0 + c) * 64, 0);
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ================================ MMA issuers ================================
    // The whole warp runs the control flow (waits, ring positions, descriptor arithmetic stay warp-uniform); one elected
    // lane issues each tcgen05.mma / commit.  Weight tiles are consumed in the producer's order: per tile
    // [Wo a, Wo b], UG_0 a, b, then for each c: UG_{c+1} a, b (if any), D_c - both warps derive ring positions from that.
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
    const int P = PROJ ? 2 : 0;
    const int per_tile = P + 3 * nchunks;
    auto w_wait = [&](int n) -> uint32_t {    // wait for weight tile number n (running count) and return its address
      const int st = n % nstages;
      mbar_wait(&w_full[st], (uint32_t)(n / nstages) & 1u);
      return smem_u32(sW + st * GF_STAGE_BYTES);
    };
    auto w_release = [&](int n) { umma_commit_1(&w_empty[n % nstages]); };
    if (warp == 1) {
      uint32_t xphase = 0, dphase = 0, ug_ph0 = 0, ug_ph1 = 0;
      unsigned long long* mdbg = (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) ? p.dbg : nullptr;   // slots 32..63
      int base = 0;
      for (int tile = blockIdx.y; tile < m_tiles; tile += gridDim.y, base += per_tile) {
        if (PROJ) {
          // x1 accumulator = ao . Wo^T in the tD columns (free: the previous tile's output has been drained)
          mbar_wait(x_full, xphase);
          mbar_wait(d_empty, dphase ^ 1);
          tc_fence_after();
          for (int kb = 0; kb < 2; ++kb) {
            const uint32_t sb = w_wait(base + kb);
            tc_fence_after();
            const uint32_t sa = smem_u32(sX + kb * 16384);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_1(tD, umma_smem_desc_sw128(sa + 32 * k, 16, 1024), umma_smem_desc_sw128(sb + 32 * k, 16, 1024),
                          idesc, (kb | k) != 0 ? 1u : 0u);
            w_release(base + kb);
          }
          umma_commit_1(d0_full);
        }
        dphase ^= 1;
        mbar_wait(xn_ready, xphase);
        xphase ^= 1;
        tc_fence_after();
        for (int c = 0; c < nchunks; ++c) {   // [u_c | g_c] = x . Wug_c^T
          const int b = c & 1;
          if (c < 8) dbg_stamp(mdbg, 32 + 2 * c);
          if (b) { mbar_wait(&ug_empty[1], ug_ph1 ^ 1); ug_ph1 ^= 1; }
          else   { mbar_wait(&ug_empty[0], ug_ph0 ^ 1); ug_ph0 ^= 1; }
          tc_fence_after();
          const int n0 = base + (c == 0 ? P : P + 3 * c - 1);
          for (int kb = 0; kb < 2; ++kb) {
            const uint32_t sb = w_wait(n0 + kb);
            tc_fence_after();
            const uint32_t sa = smem_u32(sX + kb * 16384);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_1(tUG[b], umma_smem_desc_sw128(sa + 32 * k, 16, 1024), umma_smem_desc_sw128(sb + 32 * k, 16, 1024),
                          idesc, (kb | k) != 0 ? 1u : 0u);
            w_release(n0 + kb);
          }
          umma_commit_1(&ug_full[b]);
          if (c + 1 == nchunks) umma_commit_1(x_empty);
          if (c < 8) dbg_stamp(mdbg, 33 + 2 * c);
        }
      }
    } else {
      uint32_t dphase = 0, h_ph0 = 0, h_ph1 = 0;
      int base = 0;
      for (int tile = blockIdx.y; tile < m_tiles; tile += gridDim.y, base += per_tile) {
        mbar_wait(d_empty, dphase ^ 1);       // previous tile's output accumulator has been drained
        dphase ^= 1;
        for (int c = 0; c < nchunks; ++c) {   // acc_d += h_c . Wd_c^T
          const int b = c & 1;
          if (b) { mbar_wait(&h_full[1], h_ph1); h_ph1 ^= 1; }
          else   { mbar_wait(&h_full[0], h_ph0); h_ph0 ^= 1; }
          const int n = base + (c + 1 < nchunks ? P + 4 + 3 * c : P + 2 + 3 * c);
          const uint32_t sb = w_wait(n);
          tc_fence_after();
          const uint32_t sa = smem_u32(sH + b * 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_1(tD, umma_smem_desc_sw128(sa + 32 * k, 16, 1024), umma_smem_desc_sw128(sb + 32 * k, 16, 1024), idesc,
                        (c | k) != 0 ? 1u : 0u);
          w_release(n);
          umma_commit_1(&h_empty[b]);
          if (c + 1 == nchunks) umma_commit_1(d_full);
        }
      }
    }
  } else {
    // ================================ activation / output warps ================================
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;                        // which 32 of the chunk's 64 hidden columns
    const int r = quarter * 32 + lane;                       // row within the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    uint32_t ug_ph[2] = {0, 0}, h_ph[2] = {0, 0}, dphase = 0, pf_phase = 0, pe_phase = 0, xphase = 0;
    unsigned long long head_sid = 0;
    if constexpr (HEAD) head_sid = p.head.stream_id + (p.head.stream_off ? p.head.stream_off[0] : 0ull);
    // output columns this thread finishes: single CTA -> its half of the row; cluster -> its half of the CTA's slice
    const int slice = 128 / ncta;                            // 128, 64 or 32 columns per CTA
    const int ocol0 = krank * slice + half * (slice / 2);
    const int ncol4 = slice / 8;                             // float4 groups (of 4 columns) per thread: 16, 8 or 4
    for (int tile = blockIdx.y; tile < m_tiles; tile += gridDim.y) {
      const long long row = (long long)tile * 128 + r;
      if (ncta > 1 && warp == 2 && lane == 0) mbar_arrive_expect_tx(part_full, GF_PART_BYTES);   // 128 rows x slice x 4 B x CL
      // ---- optional RMSNorm of the x tile, in place in the swizzled layout (this warp: K block `half` of its rows)
      if (PROJ) {
        // ---- x1 = bf16(bf16(ao @ Wo) + x): kept in smem for the final residual; RMSNorm(x1) becomes the MMA1 operand
        if (tile != (int)blockIdx.y) gf_bar_sync(1, GF_EPI_WARPS * 32);   // every reader of the previous tile's x1 rows is done
        uint4 rx[8];
#pragma unroll
        for (int c = 0; c < 8; ++c)
          rx[c] = row < p.M ? *reinterpret_cast<const uint4*>(p.residual + row * 128 + 64 * half + 8 * c) : make_uint4(0, 0, 0, 0);
        mbar_wait(d0_full, xphase);
        tc_fence_after();
        dbg_stamp(dbg, 2);
        uint32_t a0[32], a1[32];
        tmem_ld_32x32(tD + lane_off + 64 * half, a0);
        tmem_ld_32x32(tD + lane_off + 64 * half + 32, a1);
        tmem_ld_wait();
        tc_fence_before();
        uint32_t xw[32];
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const uint32_t* acc = i < 16 ? a0 : a1;
          const int j = i & 15;
          const uint32_t yw = pack_bf16(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]));
          const uint32_t rw = (i & 3) == 0 ? rx[i >> 2].x : (i & 3) == 1 ? rx[i >> 2].y : (i & 3) == 2 ? rx[i >> 2].z : rx[i >> 2].w;
          xw[i] = bf16x2_add(yw, rw);
          const float a = bf16_lo(xw[i]), b2 = bf16_hi(xw[i]);
          ss += a * a + b2 * b2;
        }
        const uint32_t x1row = smem_u32(sX1 + r * 256);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          sts_v4(x1row + (((8 * half + c) ^ (r & 7)) << 4), make_uint4(xw[4 * c], xw[4 * c + 1], xw[4 * c + 2], xw[4 * c + 3]));
        sSq[half * 128 + r] = ss;
        gf_bar_sync(1, GF_EPI_WARPS * 32);
        const float inv = rsqrtf((sSq[r] + sSq[128 + r]) * (1.0f / 128.0f) + p.eps);
        const unsigned long long inv2 = f32x2_pack(inv, inv);
        const float4* sc4 = reinterpret_cast<const float4*>(sScale + half * 64);
        const uint32_t xrow = smem_u32(sX + half * 16384 + r * 128);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 s0 = sc4[2 * c], s1 = sc4[2 * c + 1];
          uint4 o;
          o.x = bf16x2_norm_scale(xw[4 * c], inv2, s0.x, s0.y);
          o.y = bf16x2_norm_scale(xw[4 * c + 1], inv2, s0.z, s0.w);
          o.z = bf16x2_norm_scale(xw[4 * c + 2], inv2, s1.x, s1.y);
          o.w = bf16x2_norm_scale(xw[4 * c + 3], inv2, s1.z, s1.w);
          sts_v4(xrow + ((c ^ (r & 7)) << 4), o);
        }
        fence_proxy_async_smem();
        gf_bar_sync(1, GF_EPI_WARPS * 32);                     // sSq / sX1 complete for every row
        xphase ^= 1;
      } else {
      mbar_wait(x_full, xphase);
      xphase ^= 1;
      dbg_stamp(dbg, 2);
      if (p.norm_scale) {
        const uint32_t xrow = smem_u32(sX + half * 16384 + r * 128);
        uint4 xv[8];
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          xv[c] = lds_v4(xrow + ((c ^ (r & 7)) << 4));
          const uint32_t w4[4] = {xv[c].x, xv[c].y, xv[c].z, xv[c].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a = bf16_lo(w4[e]), bb = bf16_hi(w4[e]);
            ss += a * a + bb * bb;
          }
        }
        sSq[half * 128 + r] = ss;
        gf_bar_sync(1, GF_EPI_WARPS * 32);
        const float inv = rsqrtf((sSq[r] + sSq[128 + r]) * (1.0f / 128.0f) + p.eps);
        const float4* sc4 = reinterpret_cast<const float4*>(sScale + half * 64);      // same address for the whole warp
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 s0 = sc4[2 * c], s1 = sc4[2 * c + 1];
          uint4 o;
          o.x = pack_bf16(bf16_lo(xv[c].x) * inv * s0.x, bf16_hi(xv[c].x) * inv * s0.y);
          o.y = pack_bf16(bf16_lo(xv[c].y) * inv * s0.z, bf16_hi(xv[c].y) * inv * s0.w);
          o.z = pack_bf16(bf16_lo(xv[c].z) * inv * s1.x, bf16_hi(xv[c].z) * inv * s1.y);
          o.w = pack_bf16(bf16_lo(xv[c].w) * inv * s1.z, bf16_hi(xv[c].w) * inv * s1.w);
          sts_v4(xrow + ((c ^ (r & 7)) << 4), o);
        }
        fence_proxy_async_smem();
        gf_bar_sync(1, GF_EPI_WARPS * 32);                     // sSq may be rewritten by the next tile
      }
      }
      __syncwarp();
      dbg_stamp(dbg, 3);
      if (lane == 0) mbar_arrive(xn_ready);
      // Half-chunk software pipeline: while the 16 columns in registers are turned into h, the TMEM load of the next
      // 16 (second half of this chunk, or first half of the next chunk) is in flight - TMEM reads (64 B/clk per SM) and
      // the activation math of a warp overlap instead of alternating.
      uint32_t uA[16], gA[16], uB[16], gB[16];
      auto act16 = [&](const uint32_t (&u)[16], const uint32_t (&g)[16], uint32_t (&hw)[8]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t uw = pack_bf16(__uint_as_float(u[2 * i]), __uint_as_float(u[2 * i + 1]));
          const uint32_t gw = pack_bf16(__uint_as_float(g[2 * i]), __uint_as_float(g[2 * i + 1]));
          const uint32_t aw = gelu_tanh_bf16x2(uw);
          hw[i] = bf16x2_mul(aw, gw);
        }
      };
      mbar_wait(&ug_full[0], ug_ph[0]);
      ug_ph[0] ^= 1;
      tc_fence_after();
      tmem_ld_32x16(tUG[0] + lane_off + 32 * half, uA);
      tmem_ld_32x16(tUG[0] + lane_off + 64 + 32 * half, gA);
      tmem_ld_wait16(uA);
      tmem_ld_wait16(gA);
      for (int c = 0; c < nchunks; ++c) {
        const int b = c & 1;
        if (c < 8) dbg_stamp(dbg, 4 + 2 * c);
        tmem_ld_32x16(tUG[b] + lane_off + 32 * half + 16, uB);
        tmem_ld_32x16(tUG[b] + lane_off + 64 + 32 * half + 16, gB);
        uint32_t hw[8];
        act16(uA, gA, hw);
        mbar_wait(&h_empty[b], h_ph[b] ^ 1);                   // MMA2 of chunk c-2 no longer reads this buffer
        h_ph[b] ^= 1;
        if (HOUT) {                                         // ... and neither does its bulk store
          if (half == 0 && lane == 0) bulk_wait_group_read1();
          gf_bar_sync(2 + quarter, 64);
        }
        const uint32_t hrow = smem_u32(sH + b * 16384 + r * 128);
        sts_v4(hrow + (((4 * half) ^ (r & 7)) << 4), make_uint4(hw[0], hw[1], hw[2], hw[3]));
        sts_v4(hrow + (((4 * half + 1) ^ (r & 7)) << 4), make_uint4(hw[4], hw[5], hw[6], hw[7]));
        tmem_ld_wait16(uB);
        tmem_ld_wait16(gB);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ug_empty[b]);              // both halves of chunk c are in registers
        if (c + 1 < nchunks) {
          mbar_wait(&ug_full[b ^ 1], ug_ph[b ^ 1]);
          ug_ph[b ^ 1] ^= 1;
          tc_fence_after();
          tmem_ld_32x16(tUG[b ^ 1] + lane_off + 32 * half, uA);
          tmem_ld_32x16(tUG[b ^ 1] + lane_off + 64 + 32 * half, gA);
        }
        act16(uB, gB, hw);
        sts_v4(hrow + (((4 * half + 2) ^ (r & 7)) << 4), make_uint4(hw[0], hw[1], hw[2], hw[3]));
        sts_v4(hrow + (((4 * half + 3) ^ (r & 7)) << 4), make_uint4(hw[4], hw[5], hw[6], hw[7]));
        fence_proxy_async_smem();
        __syncwarp();
        if (c < 8) dbg_stamp(dbg, 5 + 2 * c);
        if (lane == 0) mbar_arrive(&h_full[b]);
        if (HOUT) {                                         // rows [32 quarter, +32) of h_c leave from the operand tile
          gf_bar_sync(2 + quarter, 64);
          if (half == 0 && lane == 0) {
            tma_store_2d(&tm_h, sH + b * 16384 + quarter * 4096, (chunk0 + c) * 64, tile * 128 + quarter * 32);
            bulk_commit_group();
          }
        }
        if (c + 1 < nchunks) {
          tmem_ld_wait16(uA);
          tmem_ld_wait16(gA);
        }
      }
      uint4 rpre[8];                                           // single CTA: this thread's 64 residual columns, in flight
      if (ncta == 1 && !PROJ && row < p.M) {                 // while the last chunks drain
#pragma unroll
        for (int i = 0; i < 8; ++i) rpre[i] = *reinterpret_cast<const uint4*>(p.residual + row * 128 + 64 * half + 8 * i);
      }
      mbar_wait(d_full, dphase);
      dphase ^= 1;
      tc_fence_after();
      dbg_stamp(dbg, 20);
      if (ncta == 1) {
        // output: bf16(acc) + residual -> bf16, straight from TMEM
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t acc[32];
          tmem_ld_32x32(tD + lane_off + 64 * half + 32 * cc, acc);
          tmem_ld_wait();
          if (row < p.M) {
            uint4* o4 = reinterpret_cast<uint4*>(p.out + row * 128 + 64 * half + 32 * cc);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint4 rv = PROJ ? lds_v4(smem_u32(sX1 + r * 256) + (((8 * half + 4 * cc + i) ^ (r & 7)) << 4)) : rpre[4 * cc + i];
              const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
              uint32_t ow[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                ow[e] = bf16x2_add(pack_bf16(__uint_as_float(acc[8 * i + 2 * e]), __uint_as_float(acc[8 * i + 2 * e + 1])), rw[e]);
              }
              o4[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d_empty);
      } else if constexpr (HEAD) {
        // 1'. PUSH my fp32 partial BY ROWS: CTA j finishes rows [32 j, 32 j + 32) and receives them from every CTA of the
        //     cluster into recv[src][row][32 float4 columns] (st.async completes the owner's transaction barrier)
        const bool last_tile = tile + (int)gridDim.y >= m_tiles;
        if (tile != (int)blockIdx.y) mbar_wait_cluster(part_empty, pe_phase ^ 1);   // owners consumed the previous tile
        pe_phase ^= 1;
        const uint32_t sp = smem_u32(sP);
        // the h-chunk buffers are free once the last MMA2 has completed (d_full): the policy head's parameters (action table,
        // padded per lane slot like policy_head_sample_kernel's; output norm scale) are staged there while the partials travel
        float* sT = reinterpret_cast<float*>(sH);
        {
          const int t = (int)threadIdx.x - 64;
          for (int i = t; i < p.head.A * 32; i += GF_EPI_WARPS * 32) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p.head.table) + i);
            *reinterpret_cast<float4*>(sT + (i >> 5) * 160 + ((i & 31) >> 2) * 20 + (i & 3) * 4) = v;
          }
          if (t < 32) *reinterpret_cast<float4*>(sT + 8 * 160 + 4 * t) = __ldg(reinterpret_cast<const float4*>(p.head.norm_scale) + t);
        }
        {
          uint32_t acc0[32], acc1[32];
          tmem_ld_32x32(tD + lane_off + 64 * half, acc0);
          tmem_ld_32x32(tD + lane_off + 64 * half + 32, acc1);
          tmem_ld_wait();
          // recv[src][float4 column c4][row slot]: the 32 lanes of a warp (consecutive rows, same owner = r / 32) write 512
          // contiguous bytes per instruction; the row slot is XOR-ed with (c4 / 4) & 7 so that the 8 lanes that later read one
          // row's 8 column groups hit 8 different bank groups
          const uint32_t dst = mapa_u32(sp + (uint32_t)((krank * 32 + 16 * half) * 32) * 16u, quarter);
          const uint32_t bar = mapa_u32(smem_u32(part_full), quarter);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t* a4 = (i < 8) ? &acc0[4 * i] : &acc1[4 * (i - 8)];
            const int c4 = 16 * half + i;
            st_async_v4(dst + (uint32_t)(i * 32 + (lane ^ ((c4 >> 2) & 7))) * 16u, make_uint4(a4[0], a4[1], a4[2], a4[3]), bar);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d_empty);                   // TMEM accumulator drained
        dbg_stamp(dbg, 21);
        mbar_wait(part_full, pf_phase);                        // async-proxy completion, like a TMA load
        pf_phase ^= 1;
        gf_bar_sync(1, GF_EPI_WARPS * 32);                     // the staged head parameters are visible to every warp
        dbg_stamp(dbg, 22);
        {
          // 2'. my 32 rows, 8 lanes per row: lane `sub` sums columns [16 sub, +16) in CTA-rank order, adds x1, stores the
          //     block output and feeds the policy head
          const int rl = (warp - 2) * 4 + (lane >> 3), sub = lane & 7;
          const int rt = krank * 32 + rl;                      // row inside the 128-row tile
          float4 sm[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) sm[q] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 a = lds_v4(sp + (uint32_t)(((j * 32 + 4 * sub + q) * 32 + (rl ^ sub)) * 16));
              sm[q].x += __uint_as_float(a.x); sm[q].y += __uint_as_float(a.y);
              sm[q].z += __uint_as_float(a.z); sm[q].w += __uint_as_float(a.w);
            }
          const uint32_t x1row = smem_u32(sX1 + rt * 256);
          const uint4 rv0 = lds_v4(x1row + (((2 * sub) ^ (rt & 7)) << 4)), rv1 = lds_v4(x1row + (((2 * sub + 1) ^ (rt & 7)) << 4));
          uint32_t xw[8];
          xw[0] = bf16x2_add(pack_bf16(sm[0].x, sm[0].y), rv0.x);
          xw[1] = bf16x2_add(pack_bf16(sm[0].z, sm[0].w), rv0.y);
          xw[2] = bf16x2_add(pack_bf16(sm[1].x, sm[1].y), rv0.z);
          xw[3] = bf16x2_add(pack_bf16(sm[1].z, sm[1].w), rv0.w);
          xw[4] = bf16x2_add(pack_bf16(sm[2].x, sm[2].y), rv1.x);
          xw[5] = bf16x2_add(pack_bf16(sm[2].z, sm[2].w), rv1.y);
          xw[6] = bf16x2_add(pack_bf16(sm[3].x, sm[3].y), rv1.z);
          xw[7] = bf16x2_add(pack_bf16(sm[3].z, sm[3].w), rv1.w);
          const long long grow = (long long)tile * 128 + rt;
          const bool live = grow < p.M;
          if (live) {
            uint4* o4 = reinterpret_cast<uint4*>(p.out + grow * 128 + 16 * sub);
            o4[0] = make_uint4(xw[0], xw[1], xw[2], xw[3]);
            o4[1] = make_uint4(xw[4], xw[5], xw[6], xw[7]);
          }
          float xv[16];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            xv[2 * e] = bf16_lo(xw[e]);
            xv[2 * e + 1] = bf16_hi(xw[e]);
          }
          dbg_stamp(dbg, 25);
          policy_head_row16(xv, live, grow, sub, p.head, head_sid, sT, 160, 20, sT + 8 * 160, dbg);
        }
        __syncwarp();
        dbg_stamp(dbg, 23);
        if (lane == 0 && !last_tile) {                         // my receive buffer may be overwritten by the next tile
          const uint32_t pe = smem_u32(part_empty);
          for (int j = 0; j < ncta; ++j) mbar_arrive_cluster(mapa_u32(pe, j));
        }
      } else {
        // 1. PUSH my fp32 partial to its owners: CTA j finishes output columns [j*slice, (j+1)*slice) and receives, from
        //    every CTA of the cluster, those columns into recv[src][c4 local][row] (st.async: the data itself completes
        //    the owner's transaction barrier - no fences, no acknowledgement on the critical path)
        const bool last_tile = tile + (int)gridDim.y >= m_tiles;
        if (tile != (int)blockIdx.y) mbar_wait_cluster(part_empty, pe_phase ^ 1);   // owners consumed the previous tile
        pe_phase ^= 1;
        const uint32_t sp = smem_u32(sP);
        const int per_owner = slice / 4;                       // float4 column groups per owner: 8 (CL 4) or 16 (CL 2)
        // residual of my output slice: in flight while the partials are exchanged
        uint4 rres[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
        if (ncta == 4 && row < p.M && !PROJ) {
          rres[0] = *reinterpret_cast<const uint4*>(p.residual + row * 128 + ocol0);
          rres[1] = *reinterpret_cast<const uint4*>(p.residual + row * 128 + ocol0 + 8);
        }
        {
          uint32_t acc0[32], acc1[32];
          tmem_ld_32x32(tD + lane_off + 64 * half, acc0);
          tmem_ld_32x32(tD + lane_off + 64 * half + 32, acc1);
          tmem_ld_wait();
          const uint32_t pf = smem_u32(part_full);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c4 = 16 * half + i;                      // float4 column group of the full 128-column row
            const int owner = c4 / per_owner, c4l = c4 % per_owner;
            const uint32_t dst = sp + (uint32_t)((krank * per_owner + c4l) * 128 + r) * 16u;
            const uint32_t* a4 = (i < 8) ? &acc0[4 * i] : &acc1[4 * (i - 8)];
            st_async_v4(mapa_u32(dst, owner), make_uint4(a4[0], a4[1], a4[2], a4[3]), mapa_u32(pf, owner));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d_empty);                   // TMEM accumulator drained
        dbg_stamp(dbg, 21);
        // 2. all CL partials of my slice have landed in MY shared memory: sum in CTA-rank order, add the residual, store
        mbar_wait(part_full, pf_phase);                        // async-proxy completion, like a TMA load
        pf_phase ^= 1;
        dbg_stamp(dbg, 22);
#pragma unroll
        for (int i = 0; i < 8; i += 2) {                       // 8 output columns per iteration (ncol4 = 4 or 8)
          if (i >= ncol4) break;
          const int c4l = half * ncol4 + i;                    // my first float4 group inside the slice
          float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < ncta) {
              const uint4 a = lds_v4(sp + (uint32_t)((j * per_owner + c4l) * 128 + r) * 16u);
              const uint4 bb = lds_v4(sp + (uint32_t)((j * per_owner + c4l + 1) * 128 + r) * 16u);
              s0.x += __uint_as_float(a.x); s0.y += __uint_as_float(a.y); s0.z += __uint_as_float(a.z); s0.w += __uint_as_float(a.w);
              s1.x += __uint_as_float(bb.x); s1.y += __uint_as_float(bb.y); s1.z += __uint_as_float(bb.z); s1.w += __uint_as_float(bb.w);
            }
          }
          if (row < p.M) {
            const long long o = row * 128 + ocol0 + 4 * i;
            const uint4 rv = PROJ ? lds_v4(smem_u32(sX1 + r * 256) + ((((ocol0 + 4 * i) >> 3) ^ (r & 7)) << 4))
                                    : (ncta == 4) ? rres[(i >> 1) & 1] : *reinterpret_cast<const uint4*>(p.residual + o);
            uint4 ov;
            ov.x = bf16x2_add(pack_bf16(s0.x, s0.y), rv.x);
            ov.y = bf16x2_add(pack_bf16(s0.z, s0.w), rv.y);
            ov.z = bf16x2_add(pack_bf16(s1.x, s1.y), rv.z);
            ov.w = bf16x2_add(pack_bf16(s1.z, s1.w), rv.w);
            *reinterpret_cast<uint4*>(p.out + o) = ov;
          }
        }
        __syncwarp();
        dbg_stamp(dbg, 23);
        if (lane == 0 && !last_tile) {                         // my receive buffer may be overwritten by the next tile
          const uint32_t pe = smem_u32(part_empty);
          for (int j = 0; j < ncta; ++j) mbar_arrive_cluster(mapa_u32(pe, j));
        }
      }
    }
    if (HOUT && half == 0 && lane == 0) bulk_wait_group0();   // shared memory must outlive the stores reading it
    // Exit protocol: a CTA leaves only after ITS receive barrier completed (all pushes into its shared memory have
    // landed) and nobody signals part_empty after the last tile, so no peer touches an exited CTA.
  }
  tc_fence_before();
  __syncthreads();
  dbg_stamp(dbg, 24);
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace mpx

using namespace mpx;

// head != nullptr: fold the policy head in when the launch is a cluster of 4 with the projection prologue and A <= 8;
// *head_done tells the caller whether that happened.
static int glu_fused_launch(const mpx_bf16* x, const mpx_bf16* wt_ug64, const mpx_bf16* wt_d, const mpx_bf16* residual,
                            mpx_bf16* out, int M, int H, int F, const float* norm_scale, float eps, const mpx_bf16* wt_o,
                            mpx_bf16* h_out, const PolicyHeadDev* head, bool* head_done, mpx_stream_t stream) {
  MPX_REQUIRE(x && wt_ug64 && wt_d && residual && out && M > 0, "mpx_glu_fused: bad arguments");
  MPX_REQUIRE(H == 128, "mpx_glu_fused: hidden_features=%d unsupported (128 only; use mpx_glu_up + mpx_linear)", H);
  MPX_REQUIRE(F > 0 && F % 64 == 0, "mpx_glu_fused: F=%d must be a positive multiple of 64", F);
  MPX_REQUIRE(!wt_o || norm_scale, "mpx_glu_fused: the projection prologue needs norm_scale");
  CUtensorMap tm_x, tm_wug, tm_wd, tm_wo, tm_h;
  int rc = make_tmap_bf16_2d(&tm_x, x, M, H, H, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_wug, wt_ug64, 2 * F, H, H, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_wd, wt_d, H, F, F, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_wo, wt_o ? wt_o : wt_d, H, H, wt_o ? H : F, 64, 128);
  if (rc) return rc;
  rc = h_out ? make_tmap_bf16_2d(&tm_h, h_out, M, F, F, 64, 32) : make_tmap_bf16_2d(&tm_h, wt_d, H, F, F, 64, 32);
  if (rc) return rc;
  const int smem = 1024 + 65536 + GF_STAGES * GF_STAGE_BYTES + GF_PART_BYTES + 32768 + 384 + 1024 + 512;
  GluFusedParams p;
  p.M = M; p.F = F;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.norm_scale = norm_scale;
  p.eps = eps;
  p.dbg = g_dbg_stamps;
  MPX_REQUIRE(!norm_scale || (reinterpret_cast<uintptr_t>(norm_scale) & 15) == 0, "mpx_glu_fused: norm_scale must be 16-byte aligned");
  const int tiles = ceil_div(M, 128);
  const int nchunks = F / 64;
  // largest cluster (split of the hidden dimension) that still fits one wave
  int cl = 1;
  if (nchunks % 4 == 0 && tiles * 4 <= num_sms()) cl = 4;
  else if (nchunks % 2 == 0 && tiles * 2 <= num_sms()) cl = 2;
  if (g_glu_fused_cluster == 1 || ((g_glu_fused_cluster == 2 || g_glu_fused_cluster == 4) && nchunks % g_glu_fused_cluster == 0))
    cl = g_glu_fused_cluster;
  int gy = num_sms() / cl;
  if (gy > tiles) gy = tiles;
  p.head_on = 0;
  if (head && cl == 4 && wt_o && !h_out && head->A <= 8) {
    p.head = *head;
    p.head_on = 1;
  }
  if (head_done) *head_done = p.head_on != 0;
  using Kern = void (*)(GluFusedParams, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap);
  static const Kern kerns[9] = {
      glu_fused_kernel<false, false, false, false>, glu_fused_kernel<true, false, false, false>,
      glu_fused_kernel<false, true, false, false>,  glu_fused_kernel<true, true, false, false>,
      glu_fused_kernel<false, false, true, false>,  glu_fused_kernel<true, false, true, false>,
      glu_fused_kernel<false, true, true, false>,   glu_fused_kernel<true, true, true, false>,
      glu_fused_kernel<true, true, false, true>};
  const int ki = p.head_on ? 8 : (wt_o ? 1 : 0) + (cl > 1 ? 2 : 0) + (h_out ? 4 : 0);
  static bool attr_set[9] = {};
  if (!attr_set[ki]) {
    cudaError_t e = cudaFuncSetAttribute(kerns[ki], cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(glu_fused): %s", cudaGetErrorString(e));
      return MPX_ERR_CUDA;
    }
    attr_set[ki] = true;
  }
  cudaError_t le = launch_ex(kerns[ki], dim3(cl, gy, 1), dim3(GF_THREADS), smem, (cudaStream_t)stream, cl, 8, p, tm_x, tm_wug, tm_wd,
                             tm_wo, tm_h);
  if (le != cudaSuccess) {
    set_error("cudaLaunchKernelEx(glu_fused, cluster %d): %s", cl, cudaGetErrorString(le));
    return MPX_ERR_CUDA;
  }
  return check_launch("glu_fused_kernel");
}

extern "C" int mpx_glu_fused(const mpx_bf16* x, const mpx_bf16* wt_ug64, const mpx_bf16* wt_d, const mpx_bf16* residual,
                             mpx_bf16* out, int M, int H, int F, const float* norm_scale, float eps, const mpx_bf16* wt_o,
                             mpx_bf16* h_out, mpx_stream_t stream) {
  return glu_fused_launch(x, wt_ug64, wt_d, residual, out, M, H, F, norm_scale, eps, wt_o, h_out, nullptr, nullptr, stream);
}

extern "C" int mpx_glu_fused_head(const mpx_bf16* x, const mpx_bf16* wt_ug64, const mpx_bf16* wt_d, const mpx_bf16* residual,
                                  mpx_bf16* out, int M, int H, int F, const float* norm_scale, float eps, const mpx_bf16* wt_o,
                                  const mpx_policy_head_args* h, mpx_stream_t stream) {
  MPX_REQUIRE(h && h->norm_scale && h->table && h->action && h->log_prob && h->A > 0, "mpx_glu_fused_head: bad head arguments");
  MPX_REQUIRE((reinterpret_cast<uintptr_t>(h->norm_scale) & 15) == 0 && (reinterpret_cast<uintptr_t>(h->table) & 15) == 0,
              "mpx_glu_fused_head: norm_scale and table must be 16-byte aligned");
  PolicyHeadDev d;
  d.norm_scale = h->norm_scale; d.eps = h->eps; d.table = h->table; d.mask = h->action_mask; d.gumbel_in = h->gumbel_in;
  d.seed = h->seed; d.stream_id = h->stream_id; d.stream_off = h->stream_offset; d.logits_out = h->logits_out;
  d.xf_out = reinterpret_cast<__nv_bfloat16*>(h->xf_out); d.action = h->action; d.log_prob = h->log_prob; d.A = h->A;
  bool done = false;
  const int rc = glu_fused_launch(x, wt_ug64, wt_d, residual, out, M, H, F, norm_scale, eps, wt_o, nullptr, &d, &done, stream);
  if (rc || done) return rc;
  // shapes the fused epilogue does not cover (other cluster sizes, A > 8): the same arithmetic as two launches
  return mpx_policy_head_sample(out, h->norm_scale, h->eps, h->table, h->action_mask, h->gumbel_in, h->seed, h->stream_id,
                                h->stream_offset, h->logits_out, h->xf_out, h->action, h->log_prob, M, H, h->A, stream);
}
