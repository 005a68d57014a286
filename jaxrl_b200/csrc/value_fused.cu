// Fused value path of the rollout step (TransformerActorCritic.__call__, mapox_trainer/model/network.py:321,335-341,
// and HlGaussValue.get_value, values.py:40-45):
//     xf      = RMSNorm(x)                              (optional: skipped when the caller passes normalised rows)
//     pv      = bf16(gelu(bf16(bf16(xf @ Wvm) + b_vm)))  (value MLP, H -> Vh)
//     vlogits = f32(pv) @ Wvh + b_vh                     (fp32 Linear, Vh -> NL; Wvh carried as a bf16 hi/lo pair)
//     value   = sum softmax(vlogits) * centers
// in ONE kernel: the (M, Vh) hidden activation and the logits never leave the SM.  Same skeleton as glu_fused.cu:
// two chained tcgen05 GEMMs per 64-wide chunk of Vh (MMA1 -> TMEM -> bias+gelu in registers -> bf16 smem tile ->
// MMA2 accumulating [hi | lo] logits in TMEM), and a thread-block cluster splits the Vh chunks of one 128-row tile
// over CL CTAs so that small rollout batches still fill the machine; the CL fp32 partials are summed over
// distributed shared memory, CTA k finishing rows [k*128/CL, (k+1)*128/CL) of the tile (softmax needs whole rows).
//
// Roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer / TMEM owner, warps 2-9 activation + epilogue
// (two warps per TMEM lane quarter; thread = row, each warp half of the chunk's columns).  H must be 128, Vh % 64 == 0,
// NL <= 64.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.h"

namespace mpx {

constexpr int VF_STAGES = 4;
constexpr int VF_STAGE_BYTES = 16384;
constexpr int VF_THREADS = 320;
constexpr int VF_EPI_WARPS = 8;
#ifndef VF_X_TMEM
#define VF_X_TMEM 1             // 1: the (normalised) x tile, A operand of every MMA1 of the tile, is copied to TMEM once per tile
#endif
#ifndef VF_H_TMEM
#define VF_H_TMEM 1             // 1: the pv chunk (A operand of MMA2) goes registers -> TMEM (tcgen05.st), never through shared memory
#endif

int g_value_fused_cluster = 0;   // mpx_debug_set(4, v): 0 = auto, 1/2/4 = force that cluster size (tests)

struct ValueFusedParams {
  int M, Vh, NL;
  const float* norm_scale;        // (128) or null
  float eps;
  const __nv_bfloat16* b_vm;      // (Vh)
  const float* b_vh;              // (NL)
  const float* centers;           // (NL)
  float* value;                   // (M)
  float* vlogits;                 // (M, NL) or null
  unsigned long long* dbg;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(VF_THREADS, 1)
value_fused_kernel(const ValueFusedParams p, const __grid_constant__ CUtensorMap tm_x,
                   const __grid_constant__ CUtensorMap tm_wvm, const __grid_constant__ CUtensorMap tm_whl) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;                         // 2 x 16 KB: x tile, K blocks 0 and 1
#if !VF_H_TMEM
  uint8_t* sH = smem + 32768;                 // 2 x 16 KB: pv chunk double buffer (the TMEM form leaves this region unused)
#endif
  uint8_t* sW = smem + 65536;                 // VF_STAGES x 16 KB weight ring
  uint8_t* sP = sW + VF_STAGES * VF_STAGE_BYTES;   // 32 KB: fp32 partial logits [16 float4 columns][128 rows]
  uint8_t* sR = sP + 32768;                   // 8 KB: summed logits of one 32-row group [16 float4 columns][32 rows]
  float* sSq = reinterpret_cast<float*>(sR + 8192);   // 2 x 128 partial sums of squares (RMSNorm prologue)
  float* sScale = sSq + 256;                          // norm scale (128)
  float* sBvh = sScale + 128;                         // value head bias (64)
  float* sCen = sBvh + 64;                            // bin centres (64)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sR + 8192 + 2048);
  uint64_t* w_full = bars;                    // [VF_STAGES]
  uint64_t* w_empty = bars + VF_STAGES;       // [VF_STAGES]
  uint64_t* x_full = bars + 2 * VF_STAGES;
  uint64_t* x_empty = x_full + 1;
  uint64_t* xn_ready = x_full + 2;
  uint64_t* z_full = x_full + 3;              // [2]
  uint64_t* z_empty = x_full + 5;             // [2]
  uint64_t* h_full = x_full + 7;              // [2]
  uint64_t* h_empty = x_full + 9;             // [2]
  uint64_t* d_full = x_full + 11;
  uint64_t* d_empty = x_full + 12;
  uint64_t* part_full = x_full + 13;
  uint64_t* part_empty = x_full + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_full + 15);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int ncta = (int)cluster_nctarank(), krank = (int)cluster_ctarank();
  const int nchunks = p.Vh / 64 / ncta;
  const int chunk0 = krank * nchunks;
  const int m_tiles = (p.M + 127) / 128;

  griddep_launch();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_wvm);
    tma_prefetch_desc(&tm_whl);
    for (int s = 0; s < VF_STAGES; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    mbar_init(xn_ready, VF_EPI_WARPS);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&z_full[b], 1);
      mbar_init(&z_empty[b], VF_EPI_WARPS);
      mbar_init(&h_full[b], VF_EPI_WARPS);
      mbar_init(&h_empty[b], 1);
    }
    mbar_init(d_full, 1);
    mbar_init(d_empty, VF_EPI_WARPS);
    mbar_init(part_full, 1);                   // one arrive.expect_tx per tile; the pushed bytes complete it
    mbar_init(part_empty, VF_EPI_WARPS * ncta);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (VF_H_TMEM || VF_X_TMEM) ? 512 : 256);
    tmem_relinquish();
  }
  griddep_wait();                              // everything read from global memory below may come from earlier kernels
  if (threadIdx.x >= 64 && threadIdx.x < 192) {
    const int i = threadIdx.x - 64;
    if (p.norm_scale) sScale[i] = p.norm_scale[i];
    if (i < 64) {
      sBvh[i] = i < p.NL ? p.b_vh[i] : 0.f;
      sCen[i] = i < p.NL ? p.centers[i] : 0.f;
    }
  }
  unsigned long long* dbg = (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 64) ? p.dbg : nullptr;   // warp 2 lane 0
  dbg_stamp(dbg, 0);
  tc_fence_before();
  __syncthreads();
  if (ncta > 1) cluster_sync_all();
  tc_fence_after();
  dbg_stamp(dbg, 1);
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t tZ[2] = {tmem_base, tmem_base + 64};
  const uint32_t tD = tmem_base + 128;
  const uint32_t tX = tmem_base + 320;                         // [VF_X_TMEM] x tile: 128 lanes x 64 columns (K = 128 as bf16 pairs)
  const uint32_t tH[2] = {tmem_base + 256, tmem_base + 288};   // [VF_H_TMEM] pv chunk double buffer: 128 lanes x 32 columns (bf16 pairs)

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t wphase = 0, xphase = 0;
      auto load_vm = [&](int chunk) {           // Wvm^T rows [chunk*64, +64), both K blocks: 2 x 8 KB
        mbar_wait(&w_empty[stage], wphase ^ 1);
        mbar_arrive_expect_tx(&w_full[stage], VF_STAGE_BYTES);
        tma_load_2d(&tm_wvm, &w_full[stage], sW + stage * VF_STAGE_BYTES, 0, chunk * 64);
        tma_load_2d(&tm_wvm, &w_full[stage], sW + stage * VF_STAGE_BYTES + 8192, 64, chunk * 64);
        if (++stage == VF_STAGES) { stage = 0; wphase ^= 1; }
      };
      auto load_hl = [&](int chunk) {           // [hi | lo] rows 0..127, Vh columns [chunk*64, +64): 16 KB
        mbar_wait(&w_empty[stage], wphase ^ 1);
        mbar_arrive_expect_tx(&w_full[stage], VF_STAGE_BYTES);
        tma_load_2d(&tm_whl, &w_full[stage], sW + stage * VF_STAGE_BYTES, chunk * 64, 0);
        if (++stage == VF_STAGES) { stage = 0; wphase ^= 1; }
      };
      for (int tile = blockIdx.y; tile < m_tiles; tile += gridDim.y) {
        mbar_wait(x_empty, xphase ^ 1);
        mbar_arrive_expect_tx(x_full, 32768);
        tma_load_2d(&tm_x, x_full, sX, 0, tile * 128);
        tma_load_2d(&tm_x, x_full, sX + 16384, 64, tile * 128);
        xphase ^= 1;
        load_vm(chunk0);
        for (int c = 0; c < nchunks; ++c) {
          if (c + 1 < nchunks) load_vm(chunk0 + c + 1);
          load_hl(chunk0 + c);
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // warp-uniform control flow, one elected lane per tcgen05 instruction (descriptors stay in uniform registers)
    {
      constexpr uint32_t idesc1 = umma_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idesc2 = umma_idesc_bf16(128, 128, 0, 0);
      int stage = 0;
      uint32_t wphase = 0, xnphase = 0, dphase = 0;
      uint32_t z_ph[2] = {0, 0}, h_ph[2] = {0, 0};
      auto mma1 = [&](int c, bool last) {     // z_c (128 x 64) = xf . Wvm_c^T
        const int b = c & 1;
        mbar_wait(&z_empty[b], z_ph[b] ^ 1);
        mbar_wait(&w_full[stage], wphase);
        tc_fence_after();
        const uint32_t sb = smem_u32(sW + stage * VF_STAGE_BYTES);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
#if VF_X_TMEM
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ts_1(tZ[b], tX + 32 * kb + 8 * k, umma_smem_desc_sw128(sb + kb * 8192 + 32 * k, 16, 1024), idesc1,
                           (kb | k) != 0 ? 1u : 0u);
#else
          const uint32_t sa = smem_u32(sX + kb * 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_1(tZ[b], umma_smem_desc_sw128(sa + 32 * k, 16, 1024),
                      umma_smem_desc_sw128(sb + kb * 8192 + 32 * k, 16, 1024), idesc1, (kb | k) != 0 ? 1u : 0u);
#endif
        }
        umma_commit_1(&w_empty[stage]);
        if (++stage == VF_STAGES) { stage = 0; wphase ^= 1; }
        umma_commit_1(&z_full[b]);
        if (last) umma_commit_1(x_empty);
        z_ph[b] ^= 1;
      };
      auto mma2 = [&](int c, bool last) {     // [hi | lo] += pv_c . Whl_c^T
        const int b = c & 1;
        mbar_wait(&h_full[b], h_ph[b]);
        h_ph[b] ^= 1;
        mbar_wait(&w_full[stage], wphase);
        tc_fence_after();
        const uint32_t sb = smem_u32(sW + stage * VF_STAGE_BYTES);
#if VF_H_TMEM
#pragma unroll
        for (int k = 0; k < 4; ++k)                            // 16 bf16 of K = 8 TMEM columns per step
          umma_bf16_ts_1(tD, tH[b] + 8 * k, umma_smem_desc_sw128(sb + 32 * k, 16, 1024), idesc2, (c | k) != 0 ? 1u : 0u);
#else
        const uint32_t sa = smem_u32(sH + b * 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_1(tD, umma_smem_desc_sw128(sa + 32 * k, 16, 1024), umma_smem_desc_sw128(sb + 32 * k, 16, 1024), idesc2,
                    (c | k) != 0 ? 1u : 0u);
#endif
        umma_commit_1(&w_empty[stage]);
        if (++stage == VF_STAGES) { stage = 0; wphase ^= 1; }
        umma_commit_1(&h_empty[b]);
        if (last) umma_commit_1(d_full);
      };
      for (int tile = blockIdx.y; tile < m_tiles; tile += gridDim.y) {
        mbar_wait(xn_ready, xnphase);          // x tile landed and (optionally) normalised in place
        xnphase ^= 1;
        mbar_wait(d_empty, dphase ^ 1);
        dphase ^= 1;
        tc_fence_after();
        mma1(0, nchunks == 1);
        for (int c = 0; c < nchunks; ++c) {
          if (c + 1 < nchunks) mma1(c + 1, c + 2 == nchunks);
          mma2(c, c + 1 == nchunks);
        }
      }
    }
  } else {
    // ================================ activation / epilogue warps ================================
    const int ew = warp - 2;                                 // 0..7
    const int quarter = warp & 3;
    const int half = ew >> 2;
    const int r = quarter * 32 + lane;                       // row within the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    uint32_t z_ph[2] = {0, 0}, h_ph[2] = {0, 0}, dphase = 0, pf_phase = 0, pe_phase = 0, xphase = 0;
    for (int tile = blockIdx.y; tile < m_tiles; tile += gridDim.y) {
      if (ew == 0 && lane == 0) mbar_arrive_expect_tx(part_full, 32768);   // CL x (128/CL) rows x 64 logits x 4 B
      // ---- RMSNorm prologue on the x tile, in place in the swizzled smem layout (this warp: K block `half`)
      mbar_wait(x_full, xphase);
      xphase ^= 1;
      dbg_stamp(dbg, 2);
#if VF_X_TMEM
      {
        // my row's 64 features of K block `half` (8 swizzled 16-byte chunks) -> registers -> (RMSNorm) -> TMEM columns
        // [32 half, +32) of the x operand; MMA1 of the previous tile finished with tX before x_full could complete again
        const uint32_t xrow = smem_u32(sX + half * 16384 + r * 128);
        uint32_t xw[32];
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 v = lds_v4(xrow + ((c ^ (r & 7)) << 4));
          xw[4 * c] = v.x; xw[4 * c + 1] = v.y; xw[4 * c + 2] = v.z; xw[4 * c + 3] = v.w;
        }
        if (p.norm_scale) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float a = bf16_lo(xw[i]), b = bf16_hi(xw[i]);
            ss += a * a + b * b;
          }
          sSq[half * 128 + r] = ss;
          named_bar_sync(1, VF_EPI_WARPS * 32);
          const float inv = rsqrtf((sSq[r] + sSq[128 + r]) * (1.0f / 128.0f) + p.eps);
          const unsigned long long inv2 = f32x2_pack(inv, inv);
          const float2* sc2 = reinterpret_cast<const float2*>(sScale + half * 64);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float2 s2 = sc2[i];
            xw[i] = bf16x2_norm_scale(xw[i], inv2, s2.x, s2.y);
          }
        }
        tc_fence_after();
        tmem_st_32x32(tX + lane_off + 32 * half, xw);
        tmem_st_wait_all();
        tc_fence_before();
        if (p.norm_scale) named_bar_sync(1, VF_EPI_WARPS * 32);   // sSq may be rewritten by the next tile
      }
#else
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
            const float a = bf16_lo(w4[e]), b = bf16_hi(w4[e]);
            ss += a * a + b * b;
          }
        }
        sSq[half * 128 + r] = ss;
        named_bar_sync(1, VF_EPI_WARPS * 32);
        const float inv = rsqrtf((sSq[r] + sSq[128 + r]) * (1.0f / 128.0f) + p.eps);
        const float4* sc4 = reinterpret_cast<const float4*>(sScale + half * 64);
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
        named_bar_sync(1, VF_EPI_WARPS * 32);                 // sSq may be rewritten by the next tile
      }
#endif
      __syncwarp();
      dbg_stamp(dbg, 3);
      if (lane == 0) mbar_arrive(xn_ready);

      // ---- value MLP chunks: bias + gelu on the MMA1 accumulator, bf16 tile for MMA2
      for (int c = 0; c < nchunks; ++c) {
        const int b = c & 1;
        mbar_wait(&z_full[b], z_ph[b]);
        z_ph[b] ^= 1;
        tc_fence_after();
        dbg_stamp(dbg, 4 + 2 * c);
        uint32_t z[32];
        tmem_ld_32x32(tZ[b] + lane_off + 32 * half, z);
        const uint4* b4 = reinterpret_cast<const uint4*>(p.b_vm + (chunk0 + c) * 64 + 32 * half);
        uint4 bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) bv[i] = b4[i];
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&z_empty[b]);
        uint32_t hw[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t bw[4] = {bv[i].x, bv[i].y, bv[i].z, bv[i].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t zw = pack_bf16(__uint_as_float(z[8 * i + 2 * e]), __uint_as_float(z[8 * i + 2 * e + 1]));
            const uint32_t zb = bf16x2_add(zw, bw[e]);
            hw[4 * i + e] = gelu_tanh_bf16x2(zb);
          }
        }
        mbar_wait(&h_empty[b], h_ph[b] ^ 1);
        h_ph[b] ^= 1;
#if VF_H_TMEM
        tc_fence_after();                                     // MMA2 of chunk c-2 (the buffer's last reader) has completed
        tmem_st_32x16(tH[b] + lane_off + 16 * half, hw);        // my row, K columns [32 half, +32) as 16 bf16 pairs
        tmem_st_wait_all();
        tc_fence_before();
        __syncwarp();
#else
        const uint32_t hrow = smem_u32(sH + b * 16384 + r * 128);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int pc = 4 * half + i;
          sts_v4(hrow + ((pc ^ (r & 7)) << 4), make_uint4(hw[4 * i], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]));
        }
        fence_proxy_async_smem();
        __syncwarp();
#endif
        dbg_stamp(dbg, 5 + 2 * c);
        if (lane == 0) mbar_arrive(&h_full[b]);
      }

      // ---- logits: PUSH the hi + lo partial (fp32) of each row to the CTA that finishes it (rows [j*R, (j+1)*R),
      //      R = 128/CL), into recv[src][row][16 float4 columns, XOR-swizzled by row]; the data completes the owner's
      //      transaction barrier (st.async), so no fence / acknowledgement sits on the critical path
      mbar_wait(d_full, dphase);
      dphase ^= 1;
      tc_fence_after();
      dbg_stamp(dbg, 20);
      const bool last_tile = tile + (int)gridDim.y >= m_tiles;
      if (tile != (int)blockIdx.y) mbar_wait_cluster(part_empty, pe_phase ^ 1);   // owners consumed the previous tile
      pe_phase ^= 1;
      const uint32_t sp = smem_u32(sP);
      const int R = 128 / ncta;                              // rows finished per CTA
      {
        uint32_t hi[32], lo[32];
        tmem_ld_32x32(tD + lane_off + 32 * half, hi);
        tmem_ld_32x32(tD + lane_off + 64 + 32 * half, lo);
        tmem_ld_wait();
        const int owner = r / R, rl = r % R;
        const uint32_t dst = mapa_u32(sp + (uint32_t)((krank * R + rl) * 16) * 16u, owner);
        const uint32_t bar = mapa_u32(smem_u32(part_full), owner);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c4 = (8 * half + i) ^ ((rl & 7) << 1);
          st_async_v4(dst + c4 * 16,
                      make_uint4(__float_as_uint(__uint_as_float(hi[4 * i]) + __uint_as_float(lo[4 * i])),
                                 __float_as_uint(__uint_as_float(hi[4 * i + 1]) + __uint_as_float(lo[4 * i + 1])),
                                 __float_as_uint(__uint_as_float(hi[4 * i + 2]) + __uint_as_float(lo[4 * i + 2])),
                                 __float_as_uint(__uint_as_float(hi[4 * i + 3]) + __uint_as_float(lo[4 * i + 3]))),
                      bar);
        }
      }
      tc_fence_before();
      __syncwarp();
      dbg_stamp(dbg, 21);
      if (lane == 0) mbar_arrive(d_empty);
      mbar_wait(part_full, pf_phase);                        // async-proxy completion, like a TMA load
      pf_phase ^= 1;
      dbg_stamp(dbg, 22);
      // ---- my R rows: 8 lanes per row, lane `sub` owns logits [8*sub, 8*sub+8); bias, softmax, expectation
      {
        const int sub = lane & 7;
        for (int rl = ew * 4 + (lane >> 3); rl < R; rl += VF_EPI_WARPS * 4) {
          float lg[8];
          {
            float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
            const int c4 = (2 * sub) ^ ((rl & 7) << 1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (j < ncta) {
                const uint4 a = lds_v4(sp + (uint32_t)((j * R + rl) * 16 + c4) * 16u);
                const uint4 bq = lds_v4(sp + (uint32_t)((j * R + rl) * 16 + c4 + 1) * 16u);
                s0.x += __uint_as_float(a.x); s0.y += __uint_as_float(a.y); s0.z += __uint_as_float(a.z); s0.w += __uint_as_float(a.w);
                s1.x += __uint_as_float(bq.x); s1.y += __uint_as_float(bq.y); s1.z += __uint_as_float(bq.z); s1.w += __uint_as_float(bq.w);
              }
            }
            lg[0] = s0.x; lg[1] = s0.y; lg[2] = s0.z; lg[3] = s0.w; lg[4] = s1.x; lg[5] = s1.y; lg[6] = s1.z; lg[7] = s1.w;
          }
          float m = -INFINITY;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int j = 8 * sub + e;
            lg[e] = (j < p.NL) ? lg[e] + sBvh[j] : -INFINITY;
            m = fmaxf(m, lg[e]);
          }
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
          float se = 0.f, sc = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int j = 8 * sub + e;
            if (j < p.NL) {
              const float ex = __expf(lg[e] - m);
              se += ex;
              sc = fmaf(ex, sCen[j], sc);
            }
          }
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            se += __shfl_xor_sync(0xffffffffu, se, o);
            sc += __shfl_xor_sync(0xffffffffu, sc, o);
          }
          const long long row = (long long)tile * 128 + krank * R + rl;
          if (row < p.M) {
            if (sub == 0) p.value[row] = sc / se;
            if (p.vlogits) {
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (8 * sub + e < p.NL) p.vlogits[row * p.NL + 8 * sub + e] = lg[e];
            }
          }
        }
      }
      __syncwarp();
      dbg_stamp(dbg, 24);
      if (lane == 0 && !last_tile) {                         // my receive buffer may be overwritten by the next tile
        const uint32_t pe = smem_u32(part_empty);
        for (int j = 0; j < ncta; ++j) mbar_arrive_cluster(mapa_u32(pe, j));
      }
    }
    // exit protocol as in glu_fused.cu: my receive barrier completed, nobody signals after the last tile
  }
  tc_fence_before();
  __syncthreads();
  dbg_stamp(dbg, 25);
  if (warp == 1) tmem_dealloc(tmem_base, (VF_H_TMEM || VF_X_TMEM) ? 512 : 256);
}

}  // namespace mpx

using namespace mpx;

extern "C" int mpx_value_head_fused(const mpx_bf16* x, const float* norm_scale, float eps, const mpx_bf16* wt_vm,
                                    const mpx_bf16* b_vm, const mpx_bf16* wt_vh_hilo, const float* b_vh,
                                    const float* centers, float* value, float* vlogits, int M, int H, int Vh, int NL,
                                    mpx_stream_t stream) {
  MPX_REQUIRE(x && wt_vm && b_vm && wt_vh_hilo && b_vh && centers && value && M > 0, "mpx_value_head_fused: bad arguments");
  MPX_REQUIRE(H == 128, "mpx_value_head_fused: hidden_features=%d unsupported (128 only; use mpx_linear + mpx_linear_f32w)", H);
  MPX_REQUIRE(Vh > 0 && Vh % 64 == 0, "mpx_value_head_fused: value_hidden_dim=%d must be a positive multiple of 64", Vh);
  MPX_REQUIRE(NL > 0 && NL <= 64, "mpx_value_head_fused: n_logits=%d not in [1,64]", NL);
  MPX_REQUIRE((reinterpret_cast<uintptr_t>(b_vm) & 15) == 0 && (!norm_scale || (reinterpret_cast<uintptr_t>(norm_scale) & 15) == 0),
              "mpx_value_head_fused: b_vm / norm_scale must be 16-byte aligned");
  CUtensorMap tm_x, tm_wvm, tm_whl;
  int rc = make_tmap_bf16_2d(&tm_x, x, M, H, H, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_wvm, wt_vm, Vh, H, H, 64, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_whl, wt_vh_hilo, 128, Vh, Vh, 64, 128);
  if (rc) return rc;
  const int smem = 1024 + 65536 + VF_STAGES * VF_STAGE_BYTES + 32768 + 8192 + 2048 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(value_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(value_fused): %s", cudaGetErrorString(e));
      return MPX_ERR_CUDA;
    }
    attr_set = true;
  }
  ValueFusedParams p;
  p.M = M; p.Vh = Vh; p.NL = NL;
  p.norm_scale = norm_scale;
  p.eps = eps;
  p.b_vm = reinterpret_cast<const __nv_bfloat16*>(b_vm);
  p.b_vh = b_vh;
  p.centers = centers;
  p.value = value;
  p.vlogits = vlogits;
  p.dbg = g_dbg_stamps;
  const int tiles = ceil_div(M, 128);
  const int nchunks = Vh / 64;
  int cl = 1;
  if (nchunks % 4 == 0 && tiles * 4 <= num_sms()) cl = 4;
  else if (nchunks % 2 == 0 && tiles * 2 <= num_sms()) cl = 2;
  if (g_value_fused_cluster == 1 || ((g_value_fused_cluster == 2 || g_value_fused_cluster == 4) && nchunks % g_value_fused_cluster == 0))
    cl = g_value_fused_cluster;
  int gy = num_sms() / cl;
  if (gy > tiles) gy = tiles;
  cudaError_t le = launch_ex(value_fused_kernel, dim3(cl, gy, 1), dim3(VF_THREADS), smem, (cudaStream_t)stream, cl, true, p, tm_x, tm_wvm, tm_whl);
  if (le != cudaSuccess) {
    set_error("cudaLaunchKernelEx(value_fused, cluster %d): %s", cl, cudaGetErrorString(le));
    return MPX_ERR_CUDA;
  }
  return check_launch("value_fused_kernel");
}
