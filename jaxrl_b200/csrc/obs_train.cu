// Training-side observation encoder + input embedding (reference GridCnnObsEncoder.__call__,
// mapox_trainer/model/observation.py:84-96, and the embedding sum of TransformerActorCritic.__call__, network.py:303-309,
// as nnx.grad(ppo_loss) evaluates and differentiates them, train.py:180-192,236-238) for the standard geometry
// (obs_common.cuh).  Three kernels, none of which builds an im2col matrix; one tile = 128 tokens = the M of every
// tcgen05.mma.
//
// obs_train_fwd_kernel (persistent CTAs):
//   conv1  one-hot convolution == gather-sum over the joint per-cell table (8 worker warps, fp32), z1 = bf16(bf16(acc)+b1),
//          a1 = gelu(z1) written to shared memory as 25 position tiles [128 tokens x 16 ch] (no-swizzle K-major core
//          matrices) and, with z1, to HBM for the backward;
//   conv2  81 tcgen05.mma (M128 N32 K16): D[p] += a1 tile (p + tap) . W2[tap] - the A operand of (output position, tap)
//          IS the position tile; epilogue: z2 = bf16(bf16(acc)+b2) saved, a2 = gelu(z2) into the conv3 A operand tile;
//   conv3  the 3x3 map flattened is a Linear with K = 288: 18 tcgen05.mma (M128 N128 K16); the a2 tile is copied out
//          row-major (operand of the conv3 weight gradient) while they run; epilogue: + b3 and the reward / last-action /
//          task embedding sum, coalesced 32-byte stores of x.
//   Shared memory is the constraint (a1 100 KB + W3 72 KB + table up to 46 KB): a2 and the fp32 output staging alias a1,
//   and the W3 image shares its region with the per-cell codes - it is re-fetched from L2 by one bulk copy per tile
//   while the conv2 MMAs run.
// obs_train_bwd_kernel (data gradients): da2 = dx0 . W3^T (16 MMAs M128 N144 K16), dz2 = bf16(da2) * gelu'(z2) ->
//   operand tiles; conv2 data gradient one row of the 5x5 map at a time: D[qy] (128 x 80) += dz2[py,px] . W2[qy-py]^T
//   (N = 48 at column 16 px, the buffer zeroed by an MMA on an all-zero operand), dz1 = bf16(da1) * gelu'(z1) staged
//   through shared memory and written row-major for the one-hot weight gradient; bias gradients as per-CTA partials.
// obs_conv2_wgrad_kernel: dW2[tap] (16 x 32) = sum_p a1[p + tap]^T dz2[p] with mma.sync m16n8k16 straight from the
//   position tiles (ldmatrix.trans), cp.async double-buffered half tiles.
//
// Private layouts (only these kernels read them; a thread's natural (row, 16-byte piece) access is coalesced):
//   z1t / a1t : [tile][q 25][half 2][row 128][8 ch]      z2t / dz2t : [tile][p 9][piece 4][row 128][8 ch]
#include "obs_common.cuh"

namespace mpx {

constexpr int OT_TOK = 128;             // tokens per tile
constexpr int OT_WORKERS = 8;
constexpr int OT_THREADS = (OT_WORKERS + 1) * 32;
constexpr int OT_A_BYTES = 25 * 4096;   // 25 position tiles [128 x 16] bf16; a2 (73728) / fp32 staging (65536) alias it
constexpr int OT_B_BYTES = OF_W3_BYTES; // W3 image; the codes (128 * 121 B) live here during the gather
static_assert(OT_TOK * OF_IH * OF_IW <= OT_B_BYTES, "codes must fit in the W3 region");
constexpr long long OT_Z1_TILE = 25 * 2 * 128 * 8;   // elements per tile of z1t / a1t
constexpr long long OT_Z2_TILE = 9 * 4 * 128 * 8;    // elements per tile of z2t / dz2t

// ---- shared-space accessors (a pointer that went through the 1 KB alignment arithmetic is generic to the compiler)
__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t saddr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_v2(uint32_t saddr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(saddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t saddr, uint32_t a) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(saddr), "r"(a) : "memory");
}
__device__ __forceinline__ void sts_f4(uint32_t saddr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g, bool valid) {
  const int n = valid ? 16 : 0;                               // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(g), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

// Ties registers written by an earlier (asynchronous) tcgen05.ld to the program point after the wait: volatile asm
// statements keep their order, so arithmetic on r cannot be scheduled above the tcgen05.wait::ld that precedes this.
__device__ __forceinline__ void reg_fence8(uint32_t (&r)[8]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]));
}

struct ObsTrainFwdParams {
  const uint8_t* obs;             // (M, 11, 11, K)
  ObsCodeSpec spec;
  const uint8_t* image;           // obs_common.cuh layout (mpx_obs_fused_prep)
  const float* reward;            // (M)
  const int32_t* last_action;     // (M)
  const int32_t* task_ids;        // (M) or null
  const float* w_r;               // (H)
  const float* b_r;               // (H)
  const float* a_table;           // (A, H)
  int A;
  const float* t_table;           // (n_tasks, H) or null
  int n_tasks;
  __nv_bfloat16* z1t;             // tile layout: pre-activation of conv1 (+bias)
  __nv_bfloat16* a1t;             // tile layout: gelu(z1)
  __nv_bfloat16* z2t;             // tile layout: pre-activation of conv2 (+bias)
  __nv_bfloat16* a2;              // (M, 288) row-major gelu(z2)
  __nv_bfloat16* x;               // (M, H)
  __nv_bfloat16* obs_emb;         // (M, H) or null
  long long M;
  unsigned long long* dbg;
};

// Joint codes of 4 consecutive cells from their 4 KK bytes (little-endian words w); KK compile-time: static indices only
template <int KK>
__device__ __forceinline__ uint32_t ot_codes4(const uint32_t (&w)[OF_MAXK], const ObsCodeSpec& s) {
  uint32_t packed = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    int code = 0;
#pragma unroll
    for (int k = 0; k < KK; ++k) {
      const int b = e * KK + k;                               // byte index inside the group
      const int val = (int)((w[b >> 2] >> (8 * (b & 3))) & 0xFFu);
      code += min(val, s.sizes[k]) * s.radix[k];
    }
    packed |= (uint32_t)code << (8 * e);
  }
  return packed;
}

// One joint code per cell straight from the uint8 observation: 4 cells (4 KK bytes) per thread and step; a thread's
// loads of half a tile (8 steps of 256 threads) are all in flight before the first is used.
template <int KK>
__device__ __forceinline__ void ot_codes_phase(const uint8_t* __restrict__ src, int ncell, uint32_t sCode_s, int wtid,
                                               const ObsCodeSpec& spec) {
  constexpr int STEPS = 8;
  const bool words = (reinterpret_cast<uintptr_t>(src) & 3) == 0;
#pragma unroll 1
  for (int base = 0; base * 4 < ncell; base += STEPS * OT_WORKERS * 32) {
    uint32_t w[STEPS][OF_MAXK];
#pragma unroll
    for (int it = 0; it < STEPS; ++it) {
      const int i4 = base + wtid + it * OT_WORKERS * 32;
#pragma unroll
      for (int k = 0; k < OF_MAXK; ++k) w[it][k] = 0;
      if (i4 * 4 < ncell) {
        if (words && i4 * 4 + 4 <= ncell) {
#pragma unroll
          for (int k = 0; k < KK; ++k) w[it][k] = reinterpret_cast<const uint32_t*>(src)[i4 * KK + k];
        } else {
          const int nb = min(4, ncell - i4 * 4) * KK;
#pragma unroll
          for (int b = 0; b < 4 * KK; ++b)
            if (b < nb) w[it][b >> 2] |= (uint32_t)src[i4 * 4 * KK + b] << (8 * (b & 3));
        }
      }
    }
#pragma unroll
    for (int it = 0; it < STEPS; ++it) {
      const int i4 = base + wtid + it * OT_WORKERS * 32;
      if (i4 * 4 < ncell) sts_u32(sCode_s + i4 * 4, ot_codes4<KK>(w[it], spec));
    }
  }
}

__global__ void __launch_bounds__(OT_THREADS, 1) obs_train_fwd_kernel(const ObsTrainFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int t1_bytes = 9 * p.spec.NJ * 2 * OF_C1 * 4;
  uint8_t* sA = smem;                               // a1 tiles | a2 operand tile | fp32 output staging
  uint8_t* sW3 = sA + OT_A_BYTES;                   // W3 image | codes
  uint8_t* sW2 = sW3 + OT_B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW2 + OF_W2_BYTES + t1_bytes + OF_BIAS_BYTES);
  uint64_t* a1_ready = bars;
  uint64_t* c2_full = bars + 1;
  uint64_t* a2_ready = bars + 2;
  uint64_t* c3_full = bars + 3;
  uint64_t* const_full = bars + 4;
  uint64_t* w3_full = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  const uint32_t sA_s = smem_u32(sA), sCode_s = smem_u32(sW3), sT1_s = smem_u32(sW2) + OF_W2_BYTES;
  const uint32_t sB_s = sT1_s + t1_bytes;          // b1[16] b2[32] b3[128] bf16

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int tid = threadIdx.x;
  const int n_tiles = (int)((p.M + OT_TOK - 1) / OT_TOK);

  if (tid == 0) {
    mbar_init(a1_ready, OT_WORKERS);
    mbar_init(c2_full, 1);
    mbar_init(a2_ready, OT_WORKERS);
    mbar_init(c3_full, 1);
    mbar_init(const_full, 1);
    mbar_init(w3_full, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(const_full, OF_W2_BYTES + t1_bytes + OF_BIAS_BYTES);   // W2 taps + joint table, biases
    bulk_load_1d(sW2, p.image + OF_W3_BYTES, OF_W2_BYTES + t1_bytes, const_full);
    bulk_load_1d(sW2 + OF_W2_BYTES + t1_bytes, p.image + OF_W3_BYTES + OF_W2_BYTES + OF_T1_BYTES, OF_BIAS_BYTES, const_full);
  }
  if (warp == OT_WORKERS) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t tD2 = tmem_base;            // 9 x 32 columns
  const uint32_t tD3 = tmem_base + 320;      // 128 columns
  const uint32_t sbo3 = (OF_K3 / 8) * 128;   // 4608

  if (warp == OT_WORKERS) {
    // ================================ MMA issuer (whole warp runs the control flow, one elected lane issues) ==========
    constexpr uint32_t idesc2 = umma_idesc_bf16(128, OF_C2, 0, 0);
    constexpr uint32_t idesc3 = umma_idesc_bf16(128, OF_H, 0, 0);
    uint32_t ph = 0;
    mbar_wait(const_full, 0);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      mbar_wait(a1_ready, ph);               // a1 tiles written, the codes are dead
      tc_fence_after();
      if (lane == 0) {
        fence_proxy_async_smem();
        mbar_arrive_expect_tx(w3_full, OF_W3_BYTES);
        bulk_load_1d(sW3, p.image, OF_W3_BYTES, w3_full);
      }
      __syncwarp();
      const uint32_t w2 = smem_u32(sW2);
#pragma unroll 1
      for (int pp = 0; pp < 9; ++pp) {
        const int q00 = (pp / 3) * OF_P1 + (pp % 3);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int q = q00 + (tap / 3) * OF_P1 + (tap % 3);
          umma_bf16_1(tD2 + pp * OF_C2, umma_smem_desc_nosw(sA_s + q * 4096, 128u, 256u),
                      umma_smem_desc_nosw(w2 + tap * 1024, 128u, 256u), idesc2, tap != 0 ? 1u : 0u);
        }
      }
      umma_commit_1(c2_full);
      mbar_wait(a2_ready, ph);
      mbar_wait(w3_full, ph);
      tc_fence_after();
      const uint32_t w3 = smem_u32(sW3);
#pragma unroll
      for (int s = 0; s < OF_K3 / 16; ++s)
        umma_bf16_1(tD3, umma_smem_desc_nosw(sA_s + s * 256, 128u, sbo3), umma_smem_desc_nosw(w3 + s * 256, 128u, sbo3),
                    idesc3, s != 0 ? 1u : 0u);
      umma_commit_1(c3_full);
      ph ^= 1;
    }
  } else {
    // ================================ worker warps ================================
    const int wtid = tid;                                   // 0..255
    const int quarter = warp & 3, chalf = warp >> 2;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int row = quarter * 32 + lane;                    // this thread's TMEM lane = token row of the tile
    const float sqH = bf16_round(sqrtf((float)OF_H));
    const int K = p.spec.K, NJ = p.spec.NJ;
    const int cell_bytes = OF_IH * OF_IW * K;
    const bool has_task = p.task_ids && p.t_table;
    // embedding-sum role: rows (wtid >> 3) + 32 r, columns [16 ecg, +16)
    const int erow = wtid >> 3, ecg = wtid & 7;
    uint32_t ph = 0;
    unsigned long long* dbg = (blockIdx.x == 0 && tid == 0) ? p.dbg : nullptr;
    // reward projection of this thread's 16 columns (tile-invariant)
    uint32_t wr_p[8], br_p[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 wv = reinterpret_cast<const float4*>(p.w_r + ecg * 16)[i], bq = reinterpret_cast<const float4*>(p.b_r + ecg * 16)[i];
      wr_p[2 * i] = pack_bf16(wv.x, wv.y); wr_p[2 * i + 1] = pack_bf16(wv.z, wv.w);
      br_p[2 * i] = pack_bf16(bq.x, bq.y); br_p[2 * i + 1] = pack_bf16(bq.z, bq.w);
    }
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long m0 = (long long)tile * OT_TOK;
      const int nag = (int)min((long long)OT_TOK, p.M - m0);
      dbg_stamp(dbg, 0);
      // reward / last action / task id of this thread's 4 embedding rows: fetched now, used after conv2 - two dependent
      // L2 round trips per row (index, then table row) used to sit in front of the conv3 epilogue
      float rw_r[4];
      int ea_r[4], et_r[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int er = erow + 32 * r;
        rw_r[r] = 0.f; ea_r[r] = -1; et_r[r] = -1;
        if (er < nag) {
          const long long m = m0 + er;
          rw_r[r] = p.reward[m];
          ea_r[r] = p.last_action[m];
          if (has_task) et_r[r] = p.task_ids[m];
        }
      }
      {
        const int ncell = nag * OF_IH * OF_IW;
        const uint8_t* src = p.obs + m0 * cell_bytes;
        if (K == 1) ot_codes_phase<1>(src, ncell, sCode_s, wtid, p.spec);
        else if (K == 2) ot_codes_phase<2>(src, ncell, sCode_s, wtid, p.spec);
        else if (K == 3) ot_codes_phase<3>(src, ncell, sCode_s, wtid, p.spec);
        else ot_codes_phase<4>(src, ncell, sCode_s, wtid, p.spec);
      }
      if (tile == (int)blockIdx.x) mbar_wait(const_full, 0);   // joint table and biases have landed
      of_bar_sync(1, OT_WORKERS * 32);
      dbg_stamp(dbg, 1);
      // ---- conv1: lane = (task tsub of 8, channel quad v); 9 conflict-free float4 lookups per task
      {
        const int tsub = lane >> 2, v = lane & 3;
        const uint32_t tbase = sT1_s + ((tsub & 1) * OF_C1 + 4 * v) * 4;
        const uint2 bv = lds_v2(sB_s + 8 * v);
        __nv_bfloat16* z1_tile = p.z1t + (long long)tile * OT_Z1_TILE + (v >> 1) * 1024 + (v & 1) * 4;
        __nv_bfloat16* a1_tile = p.a1t + (long long)tile * OT_Z1_TILE + (v >> 1) * 1024 + (v & 1) * 4;
        constexpr int NG = OT_TOK * 25 / 8;                    // 400 task groups
        for (int g0 = warp; g0 < NG; g0 += 2 * OT_WORKERS) {   // two task groups in flight per iteration
          float4 acc[2];
          int agq[2][2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int g = g0 + u * OT_WORKERS;
            const int task = g * 8 + tsub;
            const int ag = task & (OT_TOK - 1), q = task >> 7;
            agq[u][0] = ag; agq[u][1] = q;
            acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g < NG && ag < nag) {
              const int oy = q / OF_P1, ox = q % OF_P1;
              const uint32_t c0 = sCode_s + ag * (OF_IH * OF_IW) + (oy * 2) * OF_IW + ox * 2;
              uint32_t code[9];
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) code[tap] = lds_u8(c0 + (tap / 3) * OF_IW + (tap % 3));
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const float4 w = lds_f4(tbase + (tap * NJ + code[tap]) * (2 * OF_C1 * 4));
                acc[u].x += w.x; acc[u].y += w.y; acc[u].z += w.z; acc[u].w += w.w;
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int ag = agq[u][0], q = agq[u][1];
            if (g0 + u * OT_WORKERS < NG && ag < nag) {
              const uint32_t z0 = pack_bf16(acc[u].x, acc[u].y), z1 = pack_bf16(acc[u].z, acc[u].w);
              uint2 zz;
              zz.x = bf16x2_add(z0, bv.x);
              zz.y = bf16x2_add(z1, bv.y);
              uint2 o;
              o.x = gelu_tanh_bf16x2(zz.x);
              o.y = gelu_tanh_bf16x2(zz.y);
              sts_v2(sA_s + q * 4096 + (ag >> 3) * 256 + (v >> 1) * 128 + (ag & 7) * 16 + (v & 1) * 8, o.x, o.y);
              const int go = q * 2048 + ag * 8;                // [q][half][row][8]: 8 tokens x 16 B contiguous per half
              *reinterpret_cast<uint2*>(z1_tile + go) = zz;
              *reinterpret_cast<uint2*>(a1_tile + go) = o;
            }
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      dbg_stamp(dbg, 2);
      if (lane == 0) mbar_arrive(a1_ready);
      // ---- conv2 epilogue: row = token, this warp's 16 of the 32 channels -> z2 (HBM) and the conv3 A operand tile
      mbar_wait(c2_full, ph);
      tc_fence_after();
      dbg_stamp(dbg, 3);
      {
        const uint4 bq0 = lds_v4(sB_s + 32 + chalf * 32), bq1 = lds_v4(sB_s + 32 + chalf * 32 + 16);
        const uint32_t bw[8] = {bq0.x, bq0.y, bq0.z, bq0.w, bq1.x, bq1.y, bq1.z, bq1.w};
        __nv_bfloat16* z2_tile = p.z2t + (long long)tile * OT_Z2_TILE + (chalf * 2) * 1024 + row * 8;
        uint32_t d[2][16];
        tmem_ld_32x16(tD2 + lane_off + chalf * 16, d[0]);
#pragma unroll
        for (int pos = 0; pos < 9; ++pos) {
          tmem_ld_wait16(d[pos & 1]);
          if (pos + 1 < 9) tmem_ld_32x16(tD2 + lane_off + (pos + 1) * OF_C2 + chalf * 16, d[(pos + 1) & 1]);
          uint32_t zb[8], o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const uint32_t zw = pack_bf16(__uint_as_float(d[pos & 1][2 * e]), __uint_as_float(d[pos & 1][2 * e + 1]));
            zb[e] = bf16x2_add(zw, bw[e]);
            o[e] = gelu_tanh_bf16x2(zb[e]);
          }
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int kk = pos * 4 + chalf * 2 + hh;          // 8-wide K core matrix index: k = pos*32 + chalf*16 + hh*8
            sts_v4(sA_s + (row >> 3) * sbo3 + kk * 128 + (row & 7) * 16, make_uint4(o[4 * hh], o[4 * hh + 1], o[4 * hh + 2], o[4 * hh + 3]));
            *reinterpret_cast<uint4*>(z2_tile + (pos * 4 + hh) * 1024) = make_uint4(zb[4 * hh], zb[4 * hh + 1], zb[4 * hh + 2], zb[4 * hh + 3]);
          }
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      dbg_stamp(dbg, 4);
      if (lane == 0) mbar_arrive(a2_ready);
      // ---- a2 tile -> HBM row-major while conv3 runs: a warp instruction moves 8 rows x 64 B (conflict-free 128-byte
      //      core-matrix columns in shared memory, two full sectors per row in HBM)
      of_bar_sync(1, OT_WORKERS * 32);
      {
        const int r8 = lane & 7, pc = lane >> 3;
#pragma unroll 6
        for (int u = warp; u < 16 * 9; u += OT_WORKERS) {
          const int rg = u / 9, pb = u % 9;
          const int r = rg * 8 + r8;
          const uint4 v = lds_v4(sA_s + rg * sbo3 + (pb * 4 + pc) * 128 + r8 * 16);
          if (r < nag) *reinterpret_cast<uint4*>(p.a2 + (m0 + r) * OF_K3 + (pb * 4 + pc) * 8) = v;
        }
      }
      // ---- embedding addends of this thread's 4 (row, 16 columns) items (network.py:304-309): the table rows of all four
      //      rows are requested before the first one is packed (one round trip, not four)
      uint32_t ae_w[4][8], te_w[4][8];
#pragma unroll
      for (int r0 = 0; r0 < 4; r0 += 2) {                      // two rows per batch: 8 table loads in flight
        float4 av[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          int a = ea_r[r0 + u];
          if (erow + 32 * (r0 + u) >= nag) a = -1;
          else {
            if (a < 0) a += p.A;
            if (a >= p.A) a = -1;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
            av[u][i] = a >= 0 ? reinterpret_cast<const float4*>(p.a_table + (long long)a * OF_H + ecg * 16)[i]
                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            ae_w[r0 + u][2 * i] = pack_bf16(bf16_round(av[u][i].x) * sqH, bf16_round(av[u][i].y) * sqH);
            ae_w[r0 + u][2 * i + 1] = pack_bf16(bf16_round(av[u][i].z) * sqH, bf16_round(av[u][i].w) * sqH);
          }
      }
      if (has_task) {
#pragma unroll
        for (int r0 = 0; r0 < 4; r0 += 2) {
          float4 tv[2][4];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            int tk = et_r[r0 + u];
            if (erow + 32 * (r0 + u) >= nag) tk = -1;
            else {
              if (tk < 0) tk += p.n_tasks;
              if (tk >= p.n_tasks) tk = -1;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
              tv[u][i] = tk >= 0 ? reinterpret_cast<const float4*>(p.t_table + (long long)tk * OF_H + ecg * 16)[i]
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              te_w[r0 + u][2 * i] = pack_bf16(bf16_round(tv[u][i].x) * sqH, bf16_round(tv[u][i].y) * sqH);
              te_w[r0 + u][2 * i + 1] = pack_bf16(bf16_round(tv[u][i].z) * sqH, bf16_round(tv[u][i].w) * sqH);
            }
        }
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int h = 0; h < 8; ++h) te_w[r][h] = 0u;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) rw_r[r] = bf16_round(rw_r[r]);
      // ---- conv3 accumulator -> fp32 staging (aliases a1 / a2, both dead once c3_full fires and the copy-out is done)
      dbg_stamp(dbg, 5);
      mbar_wait(c3_full, ph);
      tc_fence_after();
      of_bar_sync(1, OT_WORKERS * 32);                       // every warp's share of the a2 copy-out has been read
      dbg_stamp(dbg, 6);
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t acc[32];
        tmem_ld_32x32(tD3 + lane_off + chalf * 64 + cc * 32, acc);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i)
          sts_f4(sA_s + (row * OF_H + chalf * 64 + cc * 32 + 4 * (i ^ (row & 7))) * 4, __uint_as_float(acc[4 * i]),
                 __uint_as_float(acc[4 * i + 1]), __uint_as_float(acc[4 * i + 2]), __uint_as_float(acc[4 * i + 3]));
      }
      tc_fence_before();
      of_bar_sync(1, OT_WORKERS * 32);
      dbg_stamp(dbg, 7);
      // ---- bias + embedding sum (network.py:303-309)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int er = erow + 32 * r;
        if (er < nag) {
          const long long m = m0 + er;
          uint32_t ow[8], ew[8];
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const int c4 = ecg * 4 + i4;                         // float4 index within the row
            const int blk = c4 >> 3, i = c4 & 7;                 // 32-column block, piece inside it (swizzled)
            const float4 vq = lds_f4(sA_s + (er * OF_H + blk * 32 + 4 * (i ^ (er & 7))) * 4);
            const uint2 b3 = lds_v2(sB_s + 96 + c4 * 8);
            const uint32_t c0 = pack_bf16(vq.x, vq.y), c1 = pack_bf16(vq.z, vq.w);              // conv3 output -> bf16
            const uint32_t e0 = bf16x2_add(c0, b3.x);   // + bias
            const uint32_t e1 = bf16x2_add(c1, b3.y);
            ew[2 * i4] = e0; ew[2 * i4 + 1] = e1;
            const float rw = rw_r[r];
            const uint32_t re0 = pack_bf16(bf16_round(rw * bf16_lo(wr_p[2 * i4])) + bf16_lo(br_p[2 * i4]),
                                           bf16_round(rw * bf16_hi(wr_p[2 * i4])) + bf16_hi(br_p[2 * i4]));
            const uint32_t re1 = pack_bf16(bf16_round(rw * bf16_lo(wr_p[2 * i4 + 1])) + bf16_lo(br_p[2 * i4 + 1]),
                                           bf16_round(rw * bf16_hi(wr_p[2 * i4 + 1])) + bf16_hi(br_p[2 * i4 + 1]));
            uint32_t x0 = bf16x2_add(e0, re0);
            uint32_t x1 = bf16x2_add(e1, re1);
            x0 = bf16x2_add(x0, ae_w[r][2 * i4]);
            x1 = bf16x2_add(x1, ae_w[r][2 * i4 + 1]);
            if (has_task) {
              x0 = bf16x2_add(x0, te_w[r][2 * i4]);
              x1 = bf16x2_add(x1, te_w[r][2 * i4 + 1]);
            }
            ow[2 * i4] = x0; ow[2 * i4 + 1] = x1;
          }
          uint4* dst = reinterpret_cast<uint4*>(p.x + m * OF_H + ecg * 16);
          dst[0] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
          dst[1] = make_uint4(ow[4], ow[5], ow[6], ow[7]);
          if (p.obs_emb) {
            uint4* de = reinterpret_cast<uint4*>(p.obs_emb + m * OF_H + ecg * 16);
            de[0] = make_uint4(ew[0], ew[1], ew[2], ew[3]);
            de[1] = make_uint4(ew[4], ew[5], ew[6], ew[7]);
          }
        }
      }
      of_bar_sync(1, OT_WORKERS * 32);                       // the staging (aliasing a1) and the codes' region are free again
      dbg_stamp(dbg, 8);
      ph ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == OT_WORKERS) tmem_dealloc(tmem_base, 512);
}

// ================================================================================================ backward: data gradients
// Backward weight image (appended to the forward image by mpx_obs_fused_prep), no-swizzle K-major core matrices:
//   W3d: B[n = pos*32 + c2][k = h] = W3[n][h]          at (n/8)*2048 + (k/8)*128 + (n%8)*16 + (k%8)*2        (73728 B)
//   W2d: per ky, B[n = kx*16 + c1][k = c2] = W2[ky,kx,c1,c2] at ky*3072 + (n/8)*512 + (k/8)*128 + (n%8)*16 + (k%8)*2 (9216 B)
//   zero block (2560 B): the all-zero B operand of the accumulator-clearing MMA
constexpr int OB_W3D_BYTES = 36 * 2048, OB_W2D_BYTES = 3 * 3072, OB_ZERO_BYTES = 10 * 256;
constexpr int OB_IMG_BYTES = OB_W3D_BYTES + OB_W2D_BYTES + OB_ZERO_BYTES;             // 85504
constexpr int OB_DX_BYTES = 16 * 2048;          // dx0 tile [128 x 128] bf16
constexpr int OB_DZ2_BYTES = 9 * 8192;          // 9 position tiles [128 x 32] bf16
constexpr int OB_STAGE_BYTES = 128 * 160;       // one row of the 5x5 map: [128 tokens][5 positions x 16 ch] bf16
constexpr int OB_SMEM = OB_IMG_BYTES + OB_DX_BYTES + OB_DZ2_BYTES + OB_STAGE_BYTES + 256 + 1024;
static_assert(OB_SMEM <= 227 * 1024, "backward kernel shared memory");
constexpr int OB_NBIAS = 48;                    // per-CTA partials: db2[32], db1[16]

struct ObsTrainBwdParams {
  const __nv_bfloat16* dx0;       // (M, 128) cotangent of the encoder output
  const __nv_bfloat16* z1t;       // tile layouts saved by the forward
  const __nv_bfloat16* z2t;
  const uint8_t* image_b;         // backward image (above)
  __nv_bfloat16* dz2t;            // tile layout, operand of the conv2 weight gradient
  __nv_bfloat16* dz1;             // (M, 25, 16) row-major, operand of the one-hot weight gradient
  float* bias_part;               // (gridDim.x, 48)
  long long M;
  unsigned long long* dbg;
};

__global__ void __launch_bounds__(OT_THREADS, 1) obs_train_bwd_kernel(const ObsTrainBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sImg = smem;
  uint8_t* sDx = sImg + OB_IMG_BYTES;
  uint8_t* sDz2 = sDx + OB_DX_BYTES;
  uint8_t* sStage = sDz2 + OB_DZ2_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + OB_STAGE_BYTES);
  uint64_t* img_full = bars;
  uint64_t* dx_ready = bars + 1;
  uint64_t* da2_full = bars + 2;
  uint64_t* dz2_ready = bars + 3;
  uint64_t* row_full = bars + 4;      // [5]
  uint64_t* row_free = bars + 9;      // [5]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  float* sRed = reinterpret_cast<float*>(sStage);            // end-of-kernel bias reduction scratch [8 warps][24]
  const uint32_t sW3d_s = smem_u32(sImg), sW2d_s = sW3d_s + OB_W3D_BYTES, sZero_s = sW2d_s + OB_W2D_BYTES;
  const uint32_t sDx_s = smem_u32(sDx), sDz2_s = smem_u32(sDz2), sStage_s = smem_u32(sStage);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int tid = threadIdx.x;
  const int n_tiles = (int)((p.M + OT_TOK - 1) / OT_TOK);

  if (tid == 0) {
    mbar_init(img_full, 1);
    mbar_init(dx_ready, OT_WORKERS);
    mbar_init(da2_full, 1);
    mbar_init(dz2_ready, OT_WORKERS);
    for (int i = 0; i < 5; ++i) {
      mbar_init(row_full + i, 1);
      mbar_init(row_free + i, OT_WORKERS);
    }
    fence_barrier_init();
    mbar_arrive_expect_tx(img_full, OB_IMG_BYTES);
    bulk_load_1d(sImg, p.image_b, OB_IMG_BYTES, img_full);
  }
  if (warp == OT_WORKERS) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t tDa2 = tmem_base;            // 288 columns
  const uint32_t tRow = tmem_base + 288;      // 2 x 80 columns

  if (warp == OT_WORKERS) {
    // ================================ MMA issuer ================================
    constexpr uint32_t idesc_a2 = umma_idesc_bf16(128, 144, 0, 0);
    constexpr uint32_t idesc_z = umma_idesc_bf16(128, 80, 0, 0);
    constexpr uint32_t idesc_r = umma_idesc_bf16(128, 48, 0, 0);
    uint32_t ph = 0;
    mbar_wait(img_full, 0);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      mbar_wait(dx_ready, ph);
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int s = 0; s < 8; ++s)
          umma_bf16_1(tDa2 + half * 144, umma_smem_desc_nosw(sDx_s + s * 256, 128u, 2048u),
                      umma_smem_desc_nosw(sW3d_s + half * 18 * 2048 + s * 256, 128u, 2048u), idesc_a2, s != 0 ? 1u : 0u);
      }
      umma_commit_1(da2_full);
      mbar_wait(dz2_ready, ph);
      tc_fence_after();
#pragma unroll 1
      for (int qy = 0; qy < 5; ++qy) {
        if (qy >= 2) {                                       // the buffer's previous row has been read
          mbar_wait(row_free + (qy - 2), ph);
          tc_fence_after();
        }
        const uint32_t tb = tRow + (qy & 1) * 80;
        umma_bf16_1(tb, umma_smem_desc_nosw(sDz2_s, 128u, 512u), umma_smem_desc_nosw(sZero_s, 128u, 256u), idesc_z, 0u);
        const int py0 = qy > 2 ? qy - 2 : 0, py1 = qy < 2 ? qy : 2;
        for (int py = py0; py <= py1; ++py) {
          const int ky = qy - py;
#pragma unroll
          for (int px = 0; px < 3; ++px) {
#pragma unroll
            for (int s = 0; s < 2; ++s)
              umma_bf16_1(tb + 16 * px, umma_smem_desc_nosw(sDz2_s + (py * 3 + px) * 8192 + s * 256, 128u, 512u),
                          umma_smem_desc_nosw(sW2d_s + ky * 3072 + s * 256, 128u, 512u), idesc_r, 1u);
          }
        }
        umma_commit_1(row_full + qy);
      }
      ph ^= 1;
    }
  } else {
    // ================================ worker warps ================================
    const int wtid = tid;
    const int quarter = warp & 3, chalf = warp >> 2;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int row = quarter * 32 + lane;
    float db2[16], db1[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) db2[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) db1[i] = 0.f;
    uint32_t ph = 0;
    unsigned long long* dbg = (blockIdx.x == 0 && tid == 0) ? p.dbg : nullptr;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long m0 = (long long)tile * OT_TOK;
      const int nag = (int)min((long long)OT_TOK, p.M - m0);
      const bool live = row < nag;
      dbg_stamp(dbg, 0);
      // ---- dx0 tile -> K-major core matrices: a warp instruction moves 8 rows x 64 B
      {
        const int r8 = lane & 7, pc = lane >> 3;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int u = warp + OT_WORKERS * i;                // 64 units = 16 row groups x 4 piece blocks
          const int rg = u >> 2, c = (u & 3) * 4 + pc, r = rg * 8 + r8;
          cp_async16(sDx_s + rg * 2048 + c * 128 + r8 * 16, p.dx0 + (m0 + (r < nag ? r : 0)) * OF_H + c * 8, r < nag);
        }
        cp_async_commit_group();
      }
      // ---- this thread's z2 pieces (18 x 16 B, coalesced in the tile layout), in flight under the dx0 copy and the MMAs
      const __nv_bfloat16* z2_tile = p.z2t + (long long)tile * OT_Z2_TILE + (chalf * 2) * 1024 + row * 8;
      __nv_bfloat16* dz2_tile = p.dz2t + (long long)tile * OT_Z2_TILE + (chalf * 2) * 1024 + row * 8;
      uint4 zq[9][2];
#pragma unroll
      for (int pos = 0; pos < 9; ++pos) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) zq[pos][hh] = *reinterpret_cast<const uint4*>(z2_tile + (pos * 4 + hh) * 1024);
      }
      cp_async_wait_group<0>();
      fence_proxy_async_smem();
      __syncwarp();
      dbg_stamp(dbg, 1);
      if (lane == 0) mbar_arrive(dx_ready);
      // ---- dz2 = bf16(bf16(da2) * gelu'(z2)): operand tiles for the conv2 data gradient + HBM (tile layout)
      mbar_wait(da2_full, ph);
      tc_fence_after();
      dbg_stamp(dbg, 2);
      {
        uint32_t d[2][16];
        tmem_ld_32x16(tDa2 + lane_off + chalf * 16, d[0]);
#pragma unroll
        for (int pos = 0; pos < 9; ++pos) {
          tmem_ld_wait16(d[pos & 1]);
          if (pos + 1 < 9) tmem_ld_32x16(tDa2 + lane_off + (pos + 1) * OF_C2 + chalf * 16, d[(pos + 1) & 1]);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const uint32_t zw[4] = {zq[pos][hh].x, zq[pos][hh].y, zq[pos][hh].z, zq[pos][hh].w};
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const uint32_t da = pack_bf16(__uint_as_float(d[pos & 1][8 * hh + 2 * e]), __uint_as_float(d[pos & 1][8 * hh + 2 * e + 1]));
              o[e] = pack_bf16(bf16_lo(da) * gelu_tanh_grad(bf16_lo(zw[e])), bf16_hi(da) * gelu_tanh_grad(bf16_hi(zw[e])));
              if (live) {
                db2[8 * hh + 2 * e] += bf16_lo(o[e]);
                db2[8 * hh + 2 * e + 1] += bf16_hi(o[e]);
              }
            }
            const uint4 ov = make_uint4(o[0], o[1], o[2], o[3]);
            sts_v4(sDz2_s + pos * 8192 + (row >> 3) * 512 + (chalf * 2 + hh) * 128 + (row & 7) * 16, ov);
            if (live) *reinterpret_cast<uint4*>(dz2_tile + (pos * 4 + hh) * 1024) = ov;
          }
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      dbg_stamp(dbg, 3);
      if (lane == 0) mbar_arrive(dz2_ready);
      const __nv_bfloat16* z1_tile = p.z1t + (long long)tile * OT_Z1_TILE + chalf * 1024 + row * 8;
      uint4 z1q[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) z1q[j] = *reinterpret_cast<const uint4*>(z1_tile + j * 2048);
      // ---- conv2 data gradient, one row of the 5x5 map at a time: dz1 = bf16(bf16(da1) * gelu'(z1)), staged row-major
#pragma unroll 1
      for (int qy = 0; qy < 5; ++qy) {
        uint4 z1n[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) z1n[j] = make_uint4(0u, 0u, 0u, 0u);
        if (qy + 1 < 5) {
#pragma unroll
          for (int j = 0; j < 5; ++j) z1n[j] = *reinterpret_cast<const uint4*>(z1_tile + ((qy + 1) * 5 + j) * 2048);
        }
        mbar_wait(row_full + qy, ph);
        tc_fence_after();
        const uint32_t tb = tRow + (qy & 1) * 80 + lane_off + chalf * 8;
        uint32_t d[5][8];
#pragma unroll
        for (int j = 0; j < 5; ++j) tmem_ld_32x8(tb + 16 * j, d[j]);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 5; ++j) reg_fence8(d[j]);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(row_free + qy);           // the accumulator buffer may be overwritten
        if (qy > 0) of_bar_sync(1, OT_WORKERS * 32);         // the previous row's copy-out has read the staging buffer
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const uint32_t zw[4] = {z1q[j].x, z1q[j].y, z1q[j].z, z1q[j].w};
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t da = pack_bf16(__uint_as_float(d[j][2 * e]), __uint_as_float(d[j][2 * e + 1]));
            o[e] = pack_bf16(bf16_lo(da) * gelu_tanh_grad(bf16_lo(zw[e])), bf16_hi(da) * gelu_tanh_grad(bf16_hi(zw[e])));
            if (live) {
              db1[2 * e] += bf16_lo(o[e]);
              db1[2 * e + 1] += bf16_hi(o[e]);
            }
          }
          sts_v4(sStage_s + row * 160 + j * 32 + chalf * 16, make_uint4(o[0], o[1], o[2], o[3]));
        }
        of_bar_sync(1, OT_WORKERS * 32);
        // copy-out: the staging buffer is 128 x 160 B, row r -> dz1 + ((m0 + r) * 25 + qy * 5) * 16 elements
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const int idx = wtid + 256 * i;                     // 1280 pieces
          const int r = idx / 10, pc = idx % 10;
          if (r < nag)
            *reinterpret_cast<uint4*>(p.dz1 + ((m0 + r) * 25 + qy * 5) * OF_C1 + pc * 8) = lds_v4(sStage_s + idx * 16);
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) z1q[j] = z1n[j];
      }
      of_bar_sync(1, OT_WORKERS * 32);                       // staging buffer and dx0 tile free for the next tile
      dbg_stamp(dbg, 10);
      ph ^= 1;
    }
    // ---- bias gradient partials of this CTA: sum over the 32 rows of a warp, then over the 4 row quarters
#pragma unroll
    for (int i = 0; i < 16; ++i) db2[i] = warp_sum(db2[i]);
#pragma unroll
    for (int i = 0; i < 8; ++i) db1[i] = warp_sum(db1[i]);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) sRed[warp * 24 + i] = db2[i];
#pragma unroll
      for (int i = 0; i < 8; ++i) sRed[warp * 24 + 16 + i] = db1[i];
    }
    of_bar_sync(1, OT_WORKERS * 32);
    if (wtid < OB_NBIAS) {
      // db2 channel c: chalf = c / 16, slot c % 16;  db1 channel c: chalf = c / 8, slot 16 + c % 8
      const int ch = wtid < 32 ? wtid : wtid - 32;
      const int half = wtid < 32 ? ch >> 4 : ch >> 3;
      const int slot = wtid < 32 ? (ch & 15) : 16 + (ch & 7);
      float s = 0.f;
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) s += sRed[(half * 4 + qd) * 24 + slot];
      p.bias_part[(long long)blockIdx.x * OB_NBIAS + wtid] = s;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == OT_WORKERS) tmem_dealloc(tmem_base, 512);
}

// db2 (32) += sum over CTAs of partial[:, 0:32], db1 (16) += partial[:, 32:48]; fixed order: deterministic
__global__ void obs_bias_reduce_kernel(const float* __restrict__ part, int nparts, float* __restrict__ db1, float* __restrict__ db2) {
  const int t = threadIdx.x;
  if (t >= OB_NBIAS) return;
  float s = 0.f;
  for (int i = 0; i < nparts; ++i) s += part[i * OB_NBIAS + t];
  if (t < 32) db2[t] += s;
  else db1[t - 32] += s;
}

// ================================================================================================ conv2 weight gradient
// dW2[(ky,kx,c1)][c2] += sum over tokens and output positions p of a1[p + (ky,kx)][c1] * dz2[p][c2].
// 12 warps = 3 kernel rows ky x 4 token slices; a warp holds the 3 taps of its row (3 x 16 x 32 fp32 accumulators), per
// (p, 16-token step): B fragments of dz2[p] once (2 ldmatrix.x4.trans), one A fragment per tap from the position tile
// p + tap (ldmatrix.x4.trans: memory rows are tokens = K, the 8 channels of a 16-byte piece are M), 12 mma.sync.
// Stages of 64 tokens, cp.async double buffered; rows past M are zero-filled.
constexpr int CW_THREADS = 384;
constexpr int CW_A_BYTES = 25 * 2 * 64 * 16;     // 51200
constexpr int CW_D_BYTES = 9 * 4 * 64 * 16;      // 36864
constexpr int CW_STAGE = CW_A_BYTES + CW_D_BYTES;
constexpr int CW_SMEM = 2 * CW_STAGE + 128;
constexpr int CW_OUT = 144 * 32;

__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(CW_THREADS, 1) obs_conv2_wgrad_kernel(const __nv_bfloat16* __restrict__ a1t,
                                                                        const __nv_bfloat16* __restrict__ dz2t,
                                                                        float* __restrict__ partial, long long M) {
  extern __shared__ __align__(128) uint8_t cw_smem[];
  const uint32_t s0 = smem_u32(cw_smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ky = warp >> 2, kq = warp & 3;
  const int n_tiles = (int)((M + OT_TOK - 1) / OT_TOK);
  const int my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int n_stages = 2 * my_tiles;

  auto issue = [&](int st) {
    const int tile = (int)blockIdx.x + (st >> 1) * (int)gridDim.x, hs = st & 1;
    const long long m0 = (long long)tile * OT_TOK + hs * 64;
    const uint32_t sb = s0 + (st & 1) * CW_STAGE;
    const __nv_bfloat16* ga = a1t + (long long)tile * OT_Z1_TILE + hs * 64 * 8;
    const __nv_bfloat16* gd = dz2t + (long long)tile * OT_Z2_TILE + hs * 64 * 8;
    for (int i = tid; i < 50 * 64; i += CW_THREADS) {          // (q, half) block b, row r: 64 rows x 16 B contiguous
      const int b = i >> 6, r = i & 63;
      cp_async16(sb + i * 16, ga + b * 1024 + r * 8, m0 + r < M);
    }
    for (int i = tid; i < 36 * 64; i += CW_THREADS) {
      const int b = i >> 6, r = i & 63;
      cp_async16(sb + CW_A_BYTES + i * 16, gd + b * 1024 + r * 8, m0 + r < M);
    }
    cp_async_commit_group();
  };

  float acc[3][4][4];
#pragma unroll
  for (int t = 0; t < 3; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[t][j][e] = 0.f;

  if (n_stages > 0) issue(0);
  // ldmatrix row addresses of this lane: matrix mi = lane >> 3 supplies row lane & 7
  //   A: mi -> (token group mi >> 1, channel half mi & 1)       B: mi -> (piece mi >> 1, token group mi & 1)
  const int mi = lane >> 3, r8 = lane & 7;
  const uint32_t a_lane = (uint32_t)(((mi & 1) * 64 + kq * 16 + (mi >> 1) * 8 + r8) * 16);      // + (q*2)*64*16
  const uint32_t b_lane = (uint32_t)(((mi >> 1) * 64 + kq * 16 + (mi & 1) * 8 + r8) * 16);      // + (p*4 + 2*jj)*64*16
  for (int st = 0; st < n_stages; ++st) {
    if (st + 1 < n_stages) {
      issue(st + 1);
      cp_async_wait_group<1>();
    } else {
      cp_async_wait_group<0>();
    }
    __syncthreads();
    const uint32_t sa = s0 + (st & 1) * CW_STAGE, sd = sa + CW_A_BYTES;
#pragma unroll 1
    for (int pp = 0; pp < 9; ++pp) {
      uint32_t b01[4], b23[4];
      ldsm_x4_t(b01, sd + (pp * 4) * 1024 + b_lane);          // n tiles 0,1: (k lo, piece 0), (k hi, piece 0), (k lo, 1), (k hi, 1)
      ldsm_x4_t(b23, sd + (pp * 4 + 2) * 1024 + b_lane);
      const int q0 = ((pp / 3) + ky) * OF_P1 + (pp % 3);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        uint32_t a[4];
        ldsm_x4_t(a, sa + ((q0 + kx) * 2) * 1024 + a_lane);
        mma_16816(acc[kx][0], a, b01[0], b01[1]);
        mma_16816(acc[kx][1], a, b01[2], b01[3]);
        mma_16816(acc[kx][2], a, b23[0], b23[1]);
        mma_16816(acc[kx][3], a, b23[2], b23[3]);
      }
    }
    __syncthreads();                                          // this buffer is refilled by the next iteration's issue
  }
  // ---- sum the 4 token slices of each kernel row through shared memory, write this CTA's partial (144 x 32)
  float* red = reinterpret_cast<float*>(cw_smem);             // [12 warps][3 kx][16][32]
  {
    const int m = lane >> 2, n2 = (lane & 3) * 2;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float* o = red + ((warp * 3 + kx) * 16) * 32 + j * 8 + n2;
        o[m * 32] = acc[kx][j][0]; o[m * 32 + 1] = acc[kx][j][1];
        o[(m + 8) * 32] = acc[kx][j][2]; o[(m + 8) * 32 + 1] = acc[kx][j][3];
      }
  }
  __syncthreads();
  for (int i = tid; i < CW_OUT; i += CW_THREADS) {            // i = ((ky*3 + kx)*16 + c1)*32 + c2
    const int kyy = i / (3 * 512), rem = i % (3 * 512);
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) s += red[(kyy * 4 + w) * (3 * 512) + rem];
    partial[(long long)blockIdx.x * CW_OUT + i] = s;
  }
}

__global__ void __launch_bounds__(256) obs_conv2_wgrad_reduce_kernel(const float* __restrict__ partial, int nparts,
                                                                     float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= CW_OUT) return;
  float s = 0.f;
  for (int c = 0; c < nparts; ++c) s += partial[(long long)c * CW_OUT + i];
  dw[i] += s;
}

// Backward image from the transposed kernels wt2 (32,144) / wt3 (128,288) (mpx_obs_fused_prep appends it to the image)
__global__ void __launch_bounds__(256) obs_train_prep_bwd_kernel(const __nv_bfloat16* __restrict__ wt2,
                                                                 const __nv_bfloat16* __restrict__ wt3,
                                                                 uint8_t* __restrict__ image_b) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
  __nv_bfloat16* w3d = reinterpret_cast<__nv_bfloat16*>(image_b);
  for (int i = tid; i < OF_K3 * OF_H; i += nthr) {              // (n, k): W3[n][k] = wt3[k][n]
    const int n = i / OF_H, k = i % OF_H;
    w3d[((n >> 3) * 2048 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) >> 1] = wt3[k * OF_K3 + n];
  }
  __nv_bfloat16* w2d = reinterpret_cast<__nv_bfloat16*>(image_b + OB_W3D_BYTES);
  for (int i = tid; i < 3 * 48 * OF_C2; i += nthr) {            // (ky, n = kx*16 + c1, k = c2): wt2[c2][(ky*3+kx)*16 + c1]
    const int k = i % OF_C2, n = (i / OF_C2) % 48, ky = i / (OF_C2 * 48);
    w2d[(ky * 3072 + (n >> 3) * 512 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) >> 1] = wt2[k * 144 + ky * 48 + n];
  }
  uint32_t* z = reinterpret_cast<uint32_t*>(image_b + OB_W3D_BYTES + OB_W2D_BYTES);
  for (int i = tid; i < OB_ZERO_BYTES / 4; i += nthr) z[i] = 0u;
}

int obs_train_prep_bwd(const __nv_bfloat16* wt2, const __nv_bfloat16* wt3, uint8_t* image_b, cudaStream_t stream) {
  obs_train_prep_bwd_kernel<<<32, 256, 0, stream>>>(wt2, wt3, image_b);
  return check_launch("obs_train_prep_bwd_kernel");
}
size_t obs_train_bwd_image_bytes() { return (size_t)OB_IMG_BYTES; }

}  // namespace mpx

using namespace mpx;

static size_t ot_fwd_smem(int NJ) {
  return (size_t)OT_A_BYTES + OT_B_BYTES + OF_W2_BYTES + (size_t)9 * NJ * 2 * OF_C1 * 4 + OF_BIAS_BYTES + 128 + 1024;
}

extern "C" int mpx_obs_train_supported(const int* sizes, int K) {
  ObsCodeSpec s;
  if (obs_fill_spec(s, sizes, K, "mpx_obs_train_fwd")) return 0;
  return ot_fwd_smem(s.NJ) <= 227 * 1024 ? 1 : 0;
}

extern "C" long long mpx_obs_train_tiles(long long M) { return (M + OT_TOK - 1) / OT_TOK; }

extern "C" int mpx_obs_train_fwd(const uint8_t* obs, const int* sizes, int K, const void* image, const float* reward,
                                 const int32_t* last_action, const int32_t* task_ids, const float* w_reward,
                                 const float* b_reward, const float* action_table, int A, const float* task_table,
                                 int n_tasks, mpx_bf16* z1t, mpx_bf16* a1t, mpx_bf16* z2t, mpx_bf16* a2, mpx_bf16* x,
                                 mpx_bf16* obs_emb, long long M, int IH, int IW, int H, mpx_stream_t stream) {
  MPX_REQUIRE(obs && image && reward && last_action && w_reward && b_reward && action_table && z1t && a1t && z2t && a2 &&
                  x && M > 0 && A > 0,
              "mpx_obs_train_fwd: bad arguments");
  MPX_REQUIRE(IH == OF_IH && IW == OF_IW && H == OF_H, "mpx_obs_train_fwd: only the 11x11 view / H=128 geometry is fused");
  MPX_REQUIRE((reinterpret_cast<uintptr_t>(image) & 15) == 0, "mpx_obs_train_fwd: image must be 16-byte aligned");
  MPX_REQUIRE(((reinterpret_cast<uintptr_t>(z1t) | reinterpret_cast<uintptr_t>(a1t) | reinterpret_cast<uintptr_t>(z2t) |
                reinterpret_cast<uintptr_t>(a2) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(obs_emb)) & 15) == 0,
              "mpx_obs_train_fwd: outputs must be 16-byte aligned");
  ObsTrainFwdParams p = {};
  int rc = obs_fill_spec(p.spec, sizes, K, "mpx_obs_train_fwd");
  if (rc) return rc;
  const size_t smem = ot_fwd_smem(p.spec.NJ);
  MPX_REQUIRE(smem <= 227 * 1024, "mpx_obs_train_fwd: %d joint one-hot codes per cell need %zu B of shared memory (use the unfused encoder)",
              p.spec.NJ, smem);
  p.obs = obs;
  p.image = reinterpret_cast<const uint8_t*>(image);
  p.reward = reward; p.last_action = last_action; p.task_ids = task_ids;
  p.w_r = w_reward; p.b_r = b_reward; p.a_table = action_table; p.A = A;
  p.t_table = task_table; p.n_tasks = n_tasks;
  p.z1t = reinterpret_cast<__nv_bfloat16*>(z1t);
  p.a1t = reinterpret_cast<__nv_bfloat16*>(a1t);
  p.z2t = reinterpret_cast<__nv_bfloat16*>(z2t);
  p.a2 = reinterpret_cast<__nv_bfloat16*>(a2);
  p.x = reinterpret_cast<__nv_bfloat16*>(x);
  p.obs_emb = reinterpret_cast<__nv_bfloat16*>(obs_emb);
  p.M = M;
  p.dbg = g_dbg_stamps;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(obs_train_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(obs_train_fwd): %s", cudaGetErrorString(e));
      return MPX_ERR_CUDA;
    }
    attr_set = true;
  }
  const long long tiles = (M + OT_TOK - 1) / OT_TOK;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  obs_train_fwd_kernel<<<grid, OT_THREADS, smem, (cudaStream_t)stream>>>(p);
  return check_launch("obs_train_fwd_kernel");
}

extern "C" size_t mpx_obs_train_bwd_workspace_bytes(void) {
  return (size_t)num_sms() * (OB_NBIAS + CW_OUT) * sizeof(float);
}

extern "C" int mpx_obs_train_bwd(const mpx_bf16* dx0, const mpx_bf16* z1t, const mpx_bf16* z2t, const void* image,
                                 mpx_bf16* dz2t, mpx_bf16* dz1, float* db1, float* db2, long long M, void* workspace,
                                 size_t workspace_bytes, mpx_stream_t stream) {
  MPX_REQUIRE(dx0 && z1t && z2t && image && dz2t && dz1 && db1 && db2 && workspace && M > 0, "mpx_obs_train_bwd: bad arguments");
  MPX_REQUIRE(workspace_bytes >= mpx_obs_train_bwd_workspace_bytes(), "mpx_obs_train_bwd: workspace of %zu bytes required",
              mpx_obs_train_bwd_workspace_bytes());
  MPX_REQUIRE(((reinterpret_cast<uintptr_t>(dx0) | reinterpret_cast<uintptr_t>(z1t) | reinterpret_cast<uintptr_t>(z2t) |
                reinterpret_cast<uintptr_t>(image) | reinterpret_cast<uintptr_t>(dz2t) | reinterpret_cast<uintptr_t>(dz1)) & 15) == 0,
              "mpx_obs_train_bwd: pointers must be 16-byte aligned");
  ObsTrainBwdParams p = {};
  p.dx0 = reinterpret_cast<const __nv_bfloat16*>(dx0);
  p.z1t = reinterpret_cast<const __nv_bfloat16*>(z1t);
  p.z2t = reinterpret_cast<const __nv_bfloat16*>(z2t);
  p.image_b = reinterpret_cast<const uint8_t*>(image) + OF_IMG_BYTES;
  p.dz2t = reinterpret_cast<__nv_bfloat16*>(dz2t);
  p.dz1 = reinterpret_cast<__nv_bfloat16*>(dz1);
  p.bias_part = reinterpret_cast<float*>(workspace);
  p.M = M;
  p.dbg = g_dbg_stamps;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(obs_train_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OB_SMEM);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(obs_train_bwd): %s", cudaGetErrorString(e));
      return MPX_ERR_CUDA;
    }
    attr_set = true;
  }
  const long long tiles = (M + OT_TOK - 1) / OT_TOK;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  obs_train_bwd_kernel<<<grid, OT_THREADS, OB_SMEM, (cudaStream_t)stream>>>(p);
  obs_bias_reduce_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(p.bias_part, grid, db1, db2);
  return check_launch("obs_train_bwd_kernel");
}

extern "C" int mpx_obs_conv2_wgrad(const mpx_bf16* a1t, const mpx_bf16* dz2t, float* dw, long long M, void* workspace,
                                   size_t workspace_bytes, mpx_stream_t stream) {
  MPX_REQUIRE(a1t && dz2t && dw && workspace && M > 0, "mpx_obs_conv2_wgrad: bad arguments");
  MPX_REQUIRE(workspace_bytes >= mpx_obs_train_bwd_workspace_bytes(), "mpx_obs_conv2_wgrad: workspace of %zu bytes required",
              mpx_obs_train_bwd_workspace_bytes());
  MPX_REQUIRE(((reinterpret_cast<uintptr_t>(a1t) | reinterpret_cast<uintptr_t>(dz2t)) & 15) == 0,
              "mpx_obs_conv2_wgrad: pointers must be 16-byte aligned");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(obs_conv2_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CW_SMEM);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(obs_conv2_wgrad): %s", cudaGetErrorString(e));
      return MPX_ERR_CUDA;
    }
    attr_set = true;
  }
  float* partial = reinterpret_cast<float*>(workspace) + (size_t)num_sms() * OB_NBIAS;
  const long long tiles = (M + OT_TOK - 1) / OT_TOK;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  obs_conv2_wgrad_kernel<<<grid, CW_THREADS, CW_SMEM, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(a1t), reinterpret_cast<const __nv_bfloat16*>(dz2t), partial, M);
  obs_conv2_wgrad_reduce_kernel<<<(CW_OUT + 255) / 256, 256, 0, (cudaStream_t)stream>>>(partial, grid, dw);
  return check_launch("obs_conv2_wgrad_kernel");
}
