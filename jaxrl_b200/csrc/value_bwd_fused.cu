// Fused backward of the value path (TransformerActorCritic.__call__, mapox_trainer/model/network.py:335-341, and the
// fp32 HL-Gauss head, values.py:34,40-41, differentiated by nnx.grad in train.py:236-238).  With xf = the normalised
// trunk output and dvl = bf16(d loss / d value logits) (M, 64: logits padded to 64 columns):
//     z  = bf16(bf16(xf Wvm) + b_vm)      pv = bf16(gelu(z))                 (recomputed - the forward keeps nothing)
//     dpv = bf16(dvl Wvh^T)               dz = bf16(dpv gelu'(z))
//     dxf = bf16(dz Wvm^T)                d b_vm = column sums of dz
// in ONE kernel per 128-row tile: the pre-activation and dpv live only in TMEM.  What leaves the kernel is pv and dz
// (operands of the weight gradients dWvh = pv^T dvl, dWvm = xf^T dz; TMA bulk stores straight from the operand tiles),
// dxf and per-CTA partial sums of d b_vm (fixed summation order).
//
// Per 64-wide chunk c of the value-MLP width three tcgen05 GEMMs are chained (same skeleton as glu_bwd_fused.cu):
//   MMA1  z_c   (128x64 fp32, TMEM) = xf . Wvm_c^T          A, B K-major
//   MMA2  dpv_c (128x64 fp32, TMEM) = dvl . Wvh_c^T         A, B K-major (K = 64 padded logits)
//   -- activation warps: z, dpv -> pv_c, dz_c (bf16) into shared memory (SW128 K-major) --
//   MMA3  dxf  += dz_c . Wvm_c                              B = the SAME Wvm_c tile as MMA1, read MN-major
// Roles (384 threads): warp 0 TMA producer, warp 1 issues MMA1/MMA2 (and owns TMEM), warps 2-9 activation warps (two per
// TMEM lane quarter; thread = row, each warp 32 of the chunk's 64 columns), warp 10 issues MMA3, warp 11 sums the columns
// of every dz tile (bias gradient).  H must be 128, Vh % 64 == 0.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.h"

namespace mpx {

constexpr int VB_THREADS = 384;
constexpr int VB_EPI_WARPS = 8;
constexpr int VB_VM_STAGES = 3;
constexpr int VB_VM_BYTES = 16384;   // Wvm_c: [64 rows x 128 h] = K blocks 0 and 1, 8 KB each
constexpr int VB_VH_BYTES = 8192;    // Wvh_c: [64 rows x 64 logits]
constexpr int VB_MAX_VH = 1024;

struct ValueBwdParams {
  int M, Vh;
  const __nv_bfloat16* b_vm;      // (Vh)
  __nv_bfloat16* dxf;             // (M, 128)
  float* bias_part;               // (gridDim.x, Vh)
  unsigned long long* dbg;
};

__global__ void __launch_bounds__(VB_THREADS, 1)
value_bwd_fused_kernel(const ValueBwdParams p, const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_dvl,
                       const __grid_constant__ CUtensorMap tm_wvm, const __grid_constant__ CUtensorMap tm_wvh,
                       const __grid_constant__ CUtensorMap tm_pv, const __grid_constant__ CUtensorMap tm_dz) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;                         // 2 x 16 KB: xf tile, K blocks 0 and 1
  uint8_t* sDVL = smem + 32768;               // 16 KB: dvl tile
  uint8_t* sWvm = smem + 49152;               // 3 x 16 KB ring
  uint8_t* sWvh = sWvm + VB_VM_STAGES * VB_VM_BYTES;   // 2 x 8 KB ring
  uint8_t* sPV = sWvh + 2 * VB_VH_BYTES;      // 2 x 16 KB: pv_c (source of the pv stores), double buffered by chunk parity
  uint8_t* sDZ = sPV + 32768;                 // 2 x 16 KB: dz_c (A operand of MMA3, source of the dz stores)
  float* sBias = reinterpret_cast<float*>(sDZ + 32768);   // VB_MAX_VH column sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDZ + 32768 + VB_MAX_VH * 4);
  uint64_t* vm_full = bars;                   // [3]
  uint64_t* vm_empty = bars + 3;              // [3]  MMA3 of the chunk has read Wvm_c
  uint64_t* vh_full = bars + 6;               // [2]
  uint64_t* vh_empty = bars + 8;              // [2]  MMA2 of the chunk has read Wvh_c
  uint64_t* x_full = bars + 10;
  uint64_t* x_empty = bars + 11;
  uint64_t* acc_full = bars + 12;             // [2]  z and dpv of a chunk are in TMEM
  uint64_t* acc_empty = bars + 14;            // [2]
  uint64_t* dz_full = bars + 16;              // [2]  pv / dz tiles written by all activation warps
  uint64_t* dz_empty = bars + 18;             // [2]  MMA3 and the bias warp have read the dz tile
  uint64_t* dx_full = bars + 20;
  uint64_t* dx_empty = bars + 21;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int nchunks = p.Vh / 64;
  const int m_tiles = (p.M + 127) / 128;

  unsigned long long* dbg = (blockIdx.x == 0 && threadIdx.x == 64) ? p.dbg : nullptr;   // warp 2 lane 0
  dbg_stamp(dbg, 0);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_dvl);
    tma_prefetch_desc(&tm_wvm);
    tma_prefetch_desc(&tm_wvh);
    tma_prefetch_desc(&tm_pv);
    tma_prefetch_desc(&tm_dz);
    for (int b = 0; b < VB_VM_STAGES; ++b) {
      mbar_init(&vm_full[b], 1);
      mbar_init(&vm_empty[b], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&vh_full[b], 1);
      mbar_init(&vh_empty[b], 1);
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], VB_EPI_WARPS);
      mbar_init(&dz_full[b], VB_EPI_WARPS);
      mbar_init(&dz_empty[b], 2);              // MMA3 commit + bias warp
    }
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    mbar_init(dx_full, 1);
    mbar_init(dx_empty, VB_EPI_WARPS);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < VB_MAX_VH; i += VB_THREADS) sBias[i] = 0.f;
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  dbg_stamp(dbg, 1);
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  auto tZ = [&](int b) { return tmem_base + 64u * b; };
  auto tP = [&](int b) { return tmem_base + 128u + 64u * b; };
  const uint32_t tDX = tmem_base + 256;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      uint32_t xphase = 0, vmphase = 0, vhphase = 0;
      int ms = 0, hs = 0;                     // ring cursors
      for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
        mbar_wait(x_empty, xphase ^ 1);
        mbar_arrive_expect_tx(x_full, 49152);
        tma_load_2d(&tm_x, x_full, sX, 0, tile * 128);
        tma_load_2d(&tm_x, x_full, sX + 16384, 64, tile * 128);
        tma_load_2d(&tm_dvl, x_full, sDVL, 0, tile * 128);
        xphase ^= 1;
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(&vm_empty[ms], vmphase ^ 1);
          mbar_arrive_expect_tx(&vm_full[ms], VB_VM_BYTES);
          tma_load_2d(&tm_wvm, &vm_full[ms], sWvm + ms * VB_VM_BYTES, 0, c * 64);
          tma_load_2d(&tm_wvm, &vm_full[ms], sWvm + ms * VB_VM_BYTES + 8192, 64, c * 64);
          if (++ms == VB_VM_STAGES) { ms = 0; vmphase ^= 1; }
          mbar_wait(&vh_empty[hs], vhphase ^ 1);
          mbar_arrive_expect_tx(&vh_full[hs], VB_VH_BYTES);
          tma_load_2d(&tm_wvh, &vh_full[hs], sWvh + hs * VB_VH_BYTES, 0, c * 64);
          if (++hs == 2) { hs = 0; vhphase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA1 / MMA2 issuer ================================
    constexpr uint32_t id_z = umma_idesc_bf16(128, 64, 0, 0);     // K-major x K-major
    uint32_t xphase = 0, acc_ph0 = 0, acc_ph1 = 0, vmphase = 0, vhphase = 0;
    int ms = 0, hs = 0;
    const uint32_t xa = smem_u32(sX), da = smem_u32(sDVL);
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
      mbar_wait(x_full, xphase);
      xphase ^= 1;
      for (int c = 0; c < nchunks; ++c) {
        const int b = c & 1;
        mbar_wait(&vm_full[ms], vmphase);
        mbar_wait(&vh_full[hs], vhphase);
        if (b) { mbar_wait(&acc_empty[1], acc_ph1 ^ 1); acc_ph1 ^= 1; }
        else   { mbar_wait(&acc_empty[0], acc_ph0 ^ 1); acc_ph0 ^= 1; }
        tc_fence_after();
        const uint32_t w = smem_u32(sWvm + ms * VB_VM_BYTES);
        const uint32_t wh = smem_u32(sWvh + hs * VB_VH_BYTES);
#pragma unroll
        for (int k = 0; k < 8; ++k)       // z_c = xf . Wvm_c^T
          umma_bf16_1(tZ(b), umma_smem_desc_sw128(xa + (k >> 2) * 16384 + 32 * (k & 3), 16, 1024),
                      umma_smem_desc_sw128(w + (k >> 2) * 8192 + 32 * (k & 3), 16, 1024), id_z, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)       // dpv_c = dvl . Wvh_c^T
          umma_bf16_1(tP(b), umma_smem_desc_sw128(da + 32 * k, 16, 1024), umma_smem_desc_sw128(wh + 32 * k, 16, 1024), id_z,
                      k != 0 ? 1u : 0u);
        umma_commit_1(&vh_empty[hs]);
        umma_commit_1(&acc_full[b]);
        if (c + 1 == nchunks) umma_commit_1(x_empty);
        if (++ms == VB_VM_STAGES) { ms = 0; vmphase ^= 1; }
        if (++hs == 2) { hs = 0; vhphase ^= 1; }
      }
    }
  } else if (warp == 10) {
    // ================================ MMA3 issuer ================================
    constexpr uint32_t id_dx = umma_idesc_bf16(128, 128, 0, 1);   // K-major x MN-major
    uint32_t dxphase = 0, dz_ph[2] = {0, 0}, vmphase = 0;
    int ms3 = 0, it = 0;                   // it: running chunk count = parity of the pv / dz tile buffer
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
      mbar_wait(dx_empty, dxphase ^ 1);    // previous tile's dxf accumulator has been drained
      dxphase ^= 1;
      for (int c = 0; c < nchunks; ++c) {  // dxf += dz_c . Wvm_c : B[n = h, k = value-MLP column]
        mbar_wait(&vm_full[ms3], vmphase); // (long since complete: MMA1 of this chunk read the same tile)
        const int b = (it++) & 1;
        mbar_wait(&dz_full[b], dz_ph[b]);
        dz_ph[b] ^= 1;
        tc_fence_after();
        const uint32_t a = smem_u32(sDZ + b * 16384);
        const uint32_t w = smem_u32(sWvm + ms3 * VB_VM_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_1(tDX, umma_smem_desc_sw128(a + 32 * k, 16, 1024), umma_smem_desc_sw128(w + 2048 * k, 8192, 1024), id_dx,
                      (c | k) != 0 ? 1u : 0u);
        umma_commit_1(&vm_empty[ms3]);
        umma_commit_1(&dz_empty[b]);
        if (c + 1 == nchunks) umma_commit_1(dx_full);
        if (++ms3 == VB_VM_STAGES) { ms3 = 0; vmphase ^= 1; }
      }
    }
  } else if (warp == 11) {
    // ================================ bias-gradient warp: column sums of every dz tile ================================
    // lane = two adjacent columns (one 32-bit word of each 128-byte row; the 32 lanes read the 32 words of a row: no
    // bank conflicts); rows past M were zero-filled by the TMA loads of xf / dvl and contribute 0
    uint32_t dz_ph[2] = {0, 0};
    int it = 0;
    const int ch = lane >> 2;
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
      for (int c = 0; c < nchunks; ++c) {
        const int b = (it++) & 1;
        mbar_wait(&dz_full[b], dz_ph[b]);
        dz_ph[b] ^= 1;
        const uint32_t base = smem_u32(sDZ + b * 16384) + (lane & 3) * 4;
        float s0 = 0.f, s1 = 0.f, t0 = 0.f, t1 = 0.f;
#pragma unroll 4
        for (int r = 0; r < 128; r += 2) {
          uint32_t w0, w1;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w0) : "r"(base + r * 128 + ((ch ^ (r & 7)) << 4)));
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w1) : "r"(base + (r + 1) * 128 + ((ch ^ ((r + 1) & 7)) << 4)));
          s0 += bf16_lo(w0); s1 += bf16_hi(w0);
          t0 += bf16_lo(w1); t1 += bf16_hi(w1);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&dz_empty[b]);
        sBias[c * 64 + 2 * lane] += s0 + t0;
        sBias[c * 64 + 2 * lane + 1] += s1 + t1;
      }
    }
    __syncwarp();
    for (int i = lane; i < p.Vh; i += 32) p.bias_part[(long long)blockIdx.x * p.Vh + i] = sBias[i];
  } else {
    // ================================ activation warps ================================
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;                        // which 32 of the chunk's 64 columns
    const int r = quarter * 32 + lane;                       // row within the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const bool storer = half == 0 && lane == 0;              // issues the bulk stores of this quarter's 32 rows
    uint32_t acc_ph[2] = {0, 0}, dz_ph[2] = {0, 0}, dxphase = 0;
    int it = 0;                                              // running chunk count = parity of the pv / dz tile buffer
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
      const long long row = (long long)tile * 128 + r;
      for (int c = 0; c < nchunks; ++c) {
        const int b = c & 1;
        const uint4* b4 = reinterpret_cast<const uint4*>(p.b_vm + c * 64 + 32 * half);
        uint4 bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) bv[i] = b4[i];
        mbar_wait(&acc_full[b], acc_ph[b]);
        acc_ph[b] ^= 1;
        tc_fence_after();
        if (c < 8) dbg_stamp(dbg, 2 + 3 * c);
        uint32_t z[32], dp[32];
        tmem_ld_32x32(tZ(b) + lane_off + 32 * half, z);
        tmem_ld_32x32(tP(b) + lane_off + 32 * half, dp);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[b]);
        uint32_t pvw[16], dzw[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t bw = i & 3 ? (i & 3) == 1 ? bv[i >> 2].y : (i & 3) == 2 ? bv[i >> 2].z : bv[i >> 2].w : bv[i >> 2].x;
          const uint32_t zw = pack_bf16(__uint_as_float(z[2 * i]), __uint_as_float(z[2 * i + 1]));
          const uint32_t zb = bf16x2_add(zw, bw);
          const uint32_t dw = pack_bf16(__uint_as_float(dp[2 * i]), __uint_as_float(dp[2 * i + 1]));
          float av[2], gr[2];
          gelu_tanh_val_grad_bf16x2(zb, av, gr);                // gelu and its derivative share one sigmoid
          pvw[i] = pack_bf16(av[0], av[1]);
          dzw[i] = pack_bf16(bf16_lo(dw) * gr[0], bf16_hi(dw) * gr[1]);
        }
        if (c < 8) dbg_stamp(dbg, 3 + 3 * c);
        // buffer b is free once MMA3 and the bias warp have read dz_(c-2) and the bulk stores of chunk c-2 have read both
        const int tb = (it++) & 1;
        mbar_wait(&dz_empty[tb], dz_ph[tb] ^ 1);
        dz_ph[tb] ^= 1;
        if (storer) bulk_wait_group_read1();
        asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
        const uint32_t pvrow = smem_u32(sPV + tb * 16384 + r * 128), dzrow = smem_u32(sDZ + tb * 16384 + r * 128);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t off = (uint32_t)(((4 * half + i) ^ (r & 7)) << 4);
          sts_v4(pvrow + off, make_uint4(pvw[4 * i], pvw[4 * i + 1], pvw[4 * i + 2], pvw[4 * i + 3]));
          sts_v4(dzrow + off, make_uint4(dzw[4 * i], dzw[4 * i + 1], dzw[4 * i + 2], dzw[4 * i + 3]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dz_full[tb]);
        if (c < 8) dbg_stamp(dbg, 4 + 3 * c);
        asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
        if (storer) {                                          // rows [32 quarter, +32) x 64 columns of pv_c and dz_c
          const int r0 = tile * 128 + quarter * 32;
          tma_store_2d(&tm_pv, sPV + tb * 16384 + quarter * 4096, c * 64, r0);
          tma_store_2d(&tm_dz, sDZ + tb * 16384 + quarter * 4096, c * 64, r0);
          bulk_commit_group();
        }
      }
      // ---- dxf tile: bf16(acc), this thread's 64 columns of its row
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
          uint4* o4 = reinterpret_cast<uint4*>(p.dxf + row * 128 + 64 * half + 32 * cc);
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

// db (Vh) += sum over CTAs of partial[cta][:]; fixed order: deterministic
__global__ void __launch_bounds__(256) value_bias_reduce_kernel(const float* __restrict__ part, int nparts, int Vh,
                                                                float* __restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Vh) return;
  float s = 0.f;
  for (int c = 0; c < nparts; ++c) s += part[(long long)c * Vh + i];
  db[i] += s;
}

}  // namespace mpx

using namespace mpx;

extern "C" size_t mpx_value_bwd_fused_workspace_bytes(int Vh) { return (size_t)num_sms() * (size_t)Vh * sizeof(float); }

extern "C" int mpx_value_bwd_fused(const mpx_bf16* xf, const mpx_bf16* dvl, const mpx_bf16* wt_vm, const mpx_bf16* b_vm,
                                   const mpx_bf16* w_vh_pad, mpx_bf16* pv, mpx_bf16* dz, mpx_bf16* dxf, float* db_vm, int M,
                                   int H, int Vh, void* workspace, size_t workspace_bytes, mpx_stream_t stream) {
  MPX_REQUIRE(xf && dvl && wt_vm && b_vm && w_vh_pad && pv && dz && dxf && db_vm && workspace && M > 0,
              "mpx_value_bwd_fused: bad arguments");
  MPX_REQUIRE(H == 128, "mpx_value_bwd_fused: hidden_features=%d unsupported (128 only)", H);
  MPX_REQUIRE(Vh > 0 && Vh % 64 == 0 && Vh <= VB_MAX_VH, "mpx_value_bwd_fused: Vh=%d must be a multiple of 64, at most %d", Vh,
              VB_MAX_VH);
  MPX_REQUIRE(workspace_bytes >= mpx_value_bwd_fused_workspace_bytes(Vh), "mpx_value_bwd_fused: workspace of %zu bytes required",
              mpx_value_bwd_fused_workspace_bytes(Vh));
  MPX_REQUIRE((reinterpret_cast<uintptr_t>(b_vm) & 15) == 0 && (reinterpret_cast<uintptr_t>(dxf) & 15) == 0,
              "mpx_value_bwd_fused: b_vm / dxf must be 16-byte aligned");
  CUtensorMap tm_x, tm_dvl, tm_wvm, tm_wvh, tm_pv, tm_dz;
  int rc = make_tmap_bf16_2d(&tm_x, xf, M, H, H, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_dvl, dvl, M, 64, 64, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_wvm, wt_vm, Vh, H, H, 64, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_wvh, w_vh_pad, Vh, 64, 64, 64, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_pv, pv, M, Vh, Vh, 64, 32);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_dz, dz, M, Vh, Vh, 64, 32);
  if (rc) return rc;
  const int smem = 1024 + 49152 + VB_VM_STAGES * VB_VM_BYTES + 2 * VB_VH_BYTES + 65536 + VB_MAX_VH * 4 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(value_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(value_bwd_fused): %s", cudaGetErrorString(e));
      return MPX_ERR_CUDA;
    }
    attr_set = true;
  }
  ValueBwdParams p;
  p.M = M; p.Vh = Vh;
  p.b_vm = reinterpret_cast<const __nv_bfloat16*>(b_vm);
  p.dxf = reinterpret_cast<__nv_bfloat16*>(dxf);
  p.bias_part = reinterpret_cast<float*>(workspace);
  p.dbg = g_dbg_stamps;
  const int tiles = ceil_div(M, 128);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  value_bwd_fused_kernel<<<grid, VB_THREADS, smem, (cudaStream_t)stream>>>(p, tm_x, tm_dvl, tm_wvm, tm_wvh, tm_pv, tm_dz);
  value_bias_reduce_kernel<<<ceil_div(Vh, 256), 256, 0, (cudaStream_t)stream>>>(p.bias_part, grid, Vh, db_vm);
  return check_launch("value_bwd_fused_kernel");
}
