// Fused backward of the GLU feed-forward block (reference GLUBlock.__call__, mapox_trainer/model/feed_forward.py:73-78,
// differentiated by jax.grad in train.py:180-230).  With xn = RMSNorm(x1) and dy = d(block output):
//     u = bf16(xn Wup)   g = bf16(xn Wgate)   a = bf16(gelu(u))   h = bf16(a g)            (recomputed, not stored)
//     dh = bf16(dy Wdown^T)   dg = bf16(dh a)   du = bf16(bf16(dh g) gelu'(u))
//     dxn = bf16(du Wup^T + dg Wgate^T)
// in ONE kernel per 128-row tile: the up/gate pre-activations and dh live only in TMEM.  The forward pass therefore
// keeps nothing but xn and h (mpx_glu_fused with h_out; h feeds dWdown = h^T dy), and what leaves this kernel is
// dug = [du | dg] (the operand of dWup/dWgate = xn^T dug) and dxn.
//
// Per 64-wide chunk c of the hidden dimension three tcgen05 GEMMs are chained (FlashAttention-backward style):
//   MMA1  [u_c | g_c] (128x128 fp32, TMEM) = xn . Wug_c^T          A, B K-major
//   MMA2  dh_c        (128x64  fp32, TMEM) = dy . Wdown_c          B = the forward's Wdown^T tile read MN-major
//   -- activation warps: u,g,dh -> du_c, dg_c (bf16) into shared memory (SW128 K-major) --
//   MMA3  dxn        += [du_c | dg_c] . Wug_c                       B = the SAME Wug_c tile as MMA1, read MN-major
// so both weight tiles are the forward kernel's (wt_ug64, wt_d): nothing extra is staged.  du_c/dg_c leave through
// TMA bulk stores straight from the operand tile.  MMA1/MMA2 of chunk c+1 overlap the activation math of chunk c.
// Weights stream from L2 through two rings: Wug_c lives until MMA3(c) (3 x 32 KB, so that the load of chunk c+2
// starts one whole chunk before it is needed), Wdown_c only until MMA2(c) (2 x 16 KB).
//
// Roles (352 threads): warp 0 TMA producer, warp 1 issues MMA1/MMA2 (and owns TMEM), warp 10 issues MMA3, warps 2-9
// are the activation warps (two per TMEM lane quarter; thread = row, each warp 32 of the chunk's 64 hidden columns).
// Two issuing warps because one thread's program order would chain MMA1/MMA2 of chunk c+1 behind "du|dg of chunk c-1
// are written", although they only need the TMEM buffer that chunk c-1's activation warps drained at their START.
// Both issuers run warp-uniform control flow with one elected lane per tcgen05 instruction, so descriptor arithmetic
// stays on the uniform datapath.  H must be 128, F % 64 == 0.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.h"

namespace mpx {

constexpr int GB_THREADS = 352;
constexpr int GB_EPI_WARPS = 8;
constexpr int GB_UG_STAGES = 3;
constexpr int GB_UG_BYTES = 32768;  // Wug_c: K blocks 0 and 1, 16 KB each
constexpr int GB_D_BYTES = 16384;   // Wdown_c

struct GluBwdParams {
  int M, F;
  __nv_bfloat16* dxn;
  unsigned long long* dbg;
};

__global__ void __launch_bounds__(GB_THREADS, 1)
glu_bwd_fused_kernel(const GluBwdParams p, const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_dy,
                     const __grid_constant__ CUtensorMap tm_wug, const __grid_constant__ CUtensorMap tm_wd,
                     const __grid_constant__ CUtensorMap tm_dug) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;                         // 2 x 16 KB: xn tile, K blocks 0 and 1
  uint8_t* sDY = smem + 32768;                // 2 x 16 KB: dy tile
  uint8_t* sWug = smem + 65536;               // 3 x 32 KB ring
  uint8_t* sWd = sWug + GB_UG_STAGES * GB_UG_BYTES;   // 2 x 16 KB ring
  uint8_t* sDUG = sWd + 2 * GB_D_BYTES;       // 2 x 16 KB: du_c | dg_c  (A operand of MMA3, source of the dug stores)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDUG + 32768);
  uint64_t* ug_full = bars;                   // [3]
  uint64_t* ug_empty = bars + 3;              // [3]  MMA3 of the chunk has read Wug_c
  uint64_t* wd_full = bars + 6;               // [2]
  uint64_t* wd_empty = bars + 8;              // [2]  MMA2 of the chunk has read Wdown_c
  uint64_t* x_full = bars + 10;
  uint64_t* x_empty = bars + 11;
  uint64_t* acc_full = bars + 12;             // [2]  u|g and dh of a chunk are in TMEM
  uint64_t* acc_empty = bars + 14;            // [2]
  uint64_t* dug_full = bars + 16;             // du|dg operand written by all activation warps
  uint64_t* dug_empty = bars + 17;            // MMA3 has read it
  uint64_t* dx_full = bars + 18;
  uint64_t* dx_empty = bars + 19;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int nchunks = p.F / 64;
  const int m_tiles = (p.M + 127) / 128;

  unsigned long long* dbg = (blockIdx.x == 0 && threadIdx.x == 64) ? p.dbg : nullptr;   // warp 2 lane 0
  dbg_stamp(dbg, 0);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_dy);
    tma_prefetch_desc(&tm_wug);
    tma_prefetch_desc(&tm_wd);
    tma_prefetch_desc(&tm_dug);
    for (int b = 0; b < GB_UG_STAGES; ++b) {
      mbar_init(&ug_full[b], 1);
      mbar_init(&ug_empty[b], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&wd_full[b], 1);
      mbar_init(&wd_empty[b], 1);
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], GB_EPI_WARPS);
    }
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    mbar_init(dug_full, GB_EPI_WARPS);
    mbar_init(dug_empty, 1);
    mbar_init(dx_full, 1);
    mbar_init(dx_empty, GB_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  dbg_stamp(dbg, 1);
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  auto tUG = [&](int b) { return tmem_base + 128u * b; };
  auto tDH = [&](int b) { return tmem_base + 256u + 64u * b; };
  const uint32_t tDX = tmem_base + 384;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      uint32_t xphase = 0, ugphase = 0, dphase = 0;
      int us = 0, ds = 0;                     // ring cursors
      for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
        mbar_wait(x_empty, xphase ^ 1);
        mbar_arrive_expect_tx(x_full, 65536);
        tma_load_2d(&tm_x, x_full, sX, 0, tile * 128);
        tma_load_2d(&tm_x, x_full, sX + 16384, 64, tile * 128);
        tma_load_2d(&tm_dy, x_full, sDY, 0, tile * 128);
        tma_load_2d(&tm_dy, x_full, sDY + 16384, 64, tile * 128);
        xphase ^= 1;
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(&ug_empty[us], ugphase ^ 1);
          mbar_arrive_expect_tx(&ug_full[us], GB_UG_BYTES);
          tma_load_2d(&tm_wug, &ug_full[us], sWug + us * GB_UG_BYTES, 0, c * 128);
          tma_load_2d(&tm_wug, &ug_full[us], sWug + us * GB_UG_BYTES + 16384, 64, c * 128);
          if (++us == GB_UG_STAGES) { us = 0; ugphase ^= 1; }
          mbar_wait(&wd_empty[ds], dphase ^ 1);
          mbar_arrive_expect_tx(&wd_full[ds], GB_D_BYTES);
          tma_load_2d(&tm_wd, &wd_full[ds], sWd + ds * GB_D_BYTES, c * 64, 0);
          if (++ds == 2) { ds = 0; dphase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA1 / MMA2 issuer ================================
    {
      constexpr uint32_t id_ug = umma_idesc_bf16(128, 128, 0, 0);   // K-major x K-major
      constexpr uint32_t id_dh = umma_idesc_bf16(128, 64, 0, 1);    // K-major x MN-major
      uint32_t xphase = 0, acc_ph0 = 0, acc_ph1 = 0, ugphase = 0, dphase = 0;
      int us = 0, ds = 0;                     // ring cursors: Wug for MMA1, Wdown for MMA2
      const uint32_t xa = smem_u32(sX), da = smem_u32(sDY);
      for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
        mbar_wait(x_full, xphase);
        xphase ^= 1;
        for (int c = 0; c < nchunks; ++c) {
          const int b = c & 1;
          mbar_wait(&ug_full[us], ugphase);
          mbar_wait(&wd_full[ds], dphase);
          if (b) { mbar_wait(&acc_empty[1], acc_ph1 ^ 1); acc_ph1 ^= 1; }
          else   { mbar_wait(&acc_empty[0], acc_ph0 ^ 1); acc_ph0 ^= 1; }
          tc_fence_after();
          const uint32_t w = smem_u32(sWug + us * GB_UG_BYTES);
          const uint32_t wd = smem_u32(sWd + ds * GB_D_BYTES);
#pragma unroll
          for (int k = 0; k < 8; ++k)       // [u | g] = xn . Wug_c^T
            umma_bf16_1(tUG(b), umma_smem_desc_sw128(xa + (k >> 2) * 16384 + 32 * (k & 3), 16, 1024),
                        umma_smem_desc_sw128(w + (k >> 2) * 16384 + 32 * (k & 3), 16, 1024), id_ug, k != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 8; ++k)       // dh = dy . Wdown_c : B[n = feature, k = hidden] from the (hidden x feature) tile
            umma_bf16_1(tDH(b), umma_smem_desc_sw128(da + (k >> 2) * 16384 + 32 * (k & 3), 16, 1024),
                        umma_smem_desc_sw128(wd + 2048 * k, 8192, 1024), id_dh, k != 0 ? 1u : 0u);
          umma_commit_1(&wd_empty[ds]);
          umma_commit_1(&acc_full[b]);
          if (c + 1 == nchunks) umma_commit_1(x_empty);
          if (++us == GB_UG_STAGES) { us = 0; ugphase ^= 1; }
          if (++ds == 2) { ds = 0; dphase ^= 1; }
        }
      }
    }
  } else if (warp == 10) {
    // ================================ MMA3 issuer ================================
    {
      constexpr uint32_t id_dx = umma_idesc_bf16(128, 128, 0, 1);   // K-major x MN-major
      uint32_t dxphase = 0, dug_ph = 0, ugphase = 0;
      int us3 = 0;                            // ring cursor: Wug for MMA3
      const uint32_t a = smem_u32(sDUG);
      for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
        mbar_wait(dx_empty, dxphase ^ 1);    // previous tile's dxn accumulator has been drained
        dxphase ^= 1;
        for (int c = 0; c < nchunks; ++c) {  // dxn += [du | dg] . Wug_c : B[n = hidden, k = feature]
          mbar_wait(&ug_full[us3], ugphase); // (long since complete: MMA1 of this chunk read the same tile)
          mbar_wait(dug_full, dug_ph);
          dug_ph ^= 1;
          tc_fence_after();
          const uint32_t w = smem_u32(sWug + us3 * GB_UG_BYTES);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16_1(tDX, umma_smem_desc_sw128(a + (k >> 2) * 16384 + 32 * (k & 3), 16, 1024),
                        umma_smem_desc_sw128(w + 2048 * k, 16384, 1024), id_dx, (c | k) != 0 ? 1u : 0u);
          umma_commit_1(&ug_empty[us3]);
          umma_commit_1(dug_empty);
          if (c + 1 == nchunks) umma_commit_1(dx_full);
          if (++us3 == GB_UG_STAGES) { us3 = 0; ugphase ^= 1; }
        }
      }
    }
  } else {
    // ================================ activation warps ================================
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;                        // which 32 of the chunk's 64 hidden columns
    const int r = quarter * 32 + lane;                       // row within the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const bool storer = half == 0 && lane == 0;              // issues the bulk stores of this quarter's 32 rows
    uint32_t acc_ph[2] = {0, 0}, dug_ph = 0, dxphase = 0;
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
      const long long row = (long long)tile * 128 + r;
      for (int c = 0; c < nchunks; ++c) {
        const int b = c & 1;
        mbar_wait(&acc_full[b], acc_ph[b]);
        acc_ph[b] ^= 1;
        tc_fence_after();
        if (c < 8) dbg_stamp(dbg, 2 + 3 * c);
        uint32_t u[32], g[32], dh[32];
        tmem_ld_32x32(tUG(b) + lane_off + 32 * half, u);
        tmem_ld_32x32(tUG(b) + lane_off + 64 + 32 * half, g);
        tmem_ld_32x32(tDH(b) + lane_off + 32 * half, dh);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[b]);
        uint32_t duw[16], dgw[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t uw = pack_bf16(__uint_as_float(u[2 * i]), __uint_as_float(u[2 * i + 1]));
          const uint32_t gw = pack_bf16(__uint_as_float(g[2 * i]), __uint_as_float(g[2 * i + 1]));
          const uint32_t dw = pack_bf16(__uint_as_float(dh[2 * i]), __uint_as_float(dh[2 * i + 1]));
          float av[2], gr[2];
          gelu_tanh_val_grad_bf16x2(uw, av, gr);                // gelu and its derivative share one sigmoid
          const uint32_t aw = pack_bf16(av[0], av[1]);
          dgw[i] = bf16x2_mul(dw, aw);
          const uint32_t tw = bf16x2_mul(dw, gw);
          duw[i] = pack_bf16(bf16_lo(tw) * gr[0], bf16_hi(tw) * gr[1]);
        }
        if (c < 8) dbg_stamp(dbg, 3 + 3 * c);
        // the operand tiles are free once MMA3 of the previous chunk retired and the previous bulk stores have read them
        mbar_wait(dug_empty, dug_ph ^ 1);
        dug_ph ^= 1;
        if (storer) bulk_wait_group_read0();
        asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
        const uint32_t durow = smem_u32(sDUG + r * 128), dgrow = durow + 16384;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t off = (uint32_t)(((4 * half + i) ^ (r & 7)) << 4);
          sts_v4(durow + off, make_uint4(duw[4 * i], duw[4 * i + 1], duw[4 * i + 2], duw[4 * i + 3]));
          sts_v4(dgrow + off, make_uint4(dgw[4 * i], dgw[4 * i + 1], dgw[4 * i + 2], dgw[4 * i + 3]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(dug_full);
        if (c < 8) dbg_stamp(dbg, 4 + 3 * c);
        asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
        if (storer) {                                          // rows [32 quarter, +32) x 64 columns of du_c and dg_c
          const int r0 = tile * 128 + quarter * 32;
          tma_store_2d(&tm_dug, sDUG + quarter * 4096, c * 64, r0);
          tma_store_2d(&tm_dug, sDUG + 16384 + quarter * 4096, p.F + c * 64, r0);
          bulk_commit_group();
        }
      }
      // ---- dxn tile: bf16(acc), this thread's 64 columns of its row
      dbg_stamp(dbg, 30);
      mbar_wait(dx_full, dxphase);
      dxphase ^= 1;
      tc_fence_after();
      dbg_stamp(dbg, 31);
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t acc[32];
        tmem_ld_32x32(tDX + lane_off + 64 * half + 32 * cc, acc);
        tmem_ld_wait();
        if (row < p.M) {
          uint4* o4 = reinterpret_cast<uint4*>(p.dxn + row * 128 + 64 * half + 32 * cc);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            o4[i] = make_uint4(pack_bf16(__uint_as_float(acc[8 * i]), __uint_as_float(acc[8 * i + 1])),
                               pack_bf16(__uint_as_float(acc[8 * i + 2]), __uint_as_float(acc[8 * i + 3])),
                               pack_bf16(__uint_as_float(acc[8 * i + 4]), __uint_as_float(acc[8 * i + 5])),
                               pack_bf16(__uint_as_float(acc[8 * i + 6]), __uint_as_float(acc[8 * i + 7])));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dx_empty);
      dbg_stamp(dbg, 32);
    }
    if (storer) bulk_wait_group0();                            // shared memory must outlive the stores reading it
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace mpx

using namespace mpx;

extern "C" int mpx_glu_bwd_fused(const mpx_bf16* xn, const mpx_bf16* dy, const mpx_bf16* wt_ug64, const mpx_bf16* wt_d,
                                 mpx_bf16* dug, mpx_bf16* dxn, int M, int H, int F, mpx_stream_t stream) {
  MPX_REQUIRE(xn && dy && wt_ug64 && wt_d && dug && dxn && M > 0, "mpx_glu_bwd_fused: bad arguments");
  MPX_REQUIRE(H == 128, "mpx_glu_bwd_fused: hidden_features=%d unsupported (128 only; use mpx_linear + mpx_glu_bwd)", H);
  MPX_REQUIRE(F > 0 && F % 64 == 0, "mpx_glu_bwd_fused: F=%d must be a positive multiple of 64", F);
  CUtensorMap tm_x, tm_dy, tm_wug, tm_wd, tm_dug;
  int rc = make_tmap_bf16_2d(&tm_x, xn, M, H, H, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_dy, dy, M, H, H, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_wug, wt_ug64, 2 * F, H, H, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_wd, wt_d, H, F, F, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_dug, dug, M, 2 * F, 2 * F, 64, 32);
  if (rc) return rc;
  const int smem = 1024 + 65536 + GB_UG_STAGES * GB_UG_BYTES + 2 * GB_D_BYTES + 32768 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(glu_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(glu_bwd_fused): %s", cudaGetErrorString(e));
      return MPX_ERR_CUDA;
    }
    attr_set = true;
  }
  GluBwdParams p;
  p.M = M; p.F = F;
  p.dxn = reinterpret_cast<__nv_bfloat16*>(dxn);
  p.dbg = g_dbg_stamps;
  const int tiles = ceil_div(M, 128);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  glu_bwd_fused_kernel<<<grid, GB_THREADS, smem, (cudaStream_t)stream>>>(p, tm_x, tm_dy, tm_wug, tm_wd, tm_dug);
  return check_launch("glu_bwd_fused_kernel");
}
