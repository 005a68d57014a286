// Fused observation encoder + input embedding of the rollout step (reference GridCnnObsEncoder.__call__,
// mapox_trainer/model/observation.py:84-96, and the embedding sum of TransformerActorCritic.__call__,
// network.py:303-309) for the standard geometry: 11x11xK uint8 view -> one-hot -> conv 3x3/2 (16) -> gelu ->
// conv 3x3/1 (32) -> gelu -> conv 3x3/1 (H = 128) -> + reward / last-action / task embeddings.
//
// One CTA per tile of 32 agents; nothing but the uint8 observation is read from and nothing but the (B, H) bf16
// embedding is written to HBM:
//   conv1  one-hot convolution == a gather-sum of weight rows (SIMT, 8 worker warps), result a1 (25 positions x 16 ch)
//          written to shared memory directly in the tcgen05 no-swizzle K-major core-matrix layout;
//   conv2  27 tcgen05.mma (M128 N32 K16): for output row py, tap (ky,kx): D[py] += A(q) . W2[tap], where the A
//          operand is simply the a1 tile of input position q = (py+ky)*5 + kx - rows 0..95 of the MMA are
//          (px, agent) because consecutive px are consecutive q tiles; no im2col copy exists anywhere;
//   conv3  the 3x3 map flattened is a Linear with K = 288: 18 tcgen05.mma (M128 N128 K16) on a2 (bias+gelu epilogue
//          of conv2, written by the worker warps in core-matrix layout); only rows 0..31 of the tile are agents.
// conv1's per-cell lookups use a JOINT table: one code per cell combines all K components (an extra code per component
// stands for an out-of-range value = all-zero one-hot), so a 3x3 receptive field costs 9 lookups of a 16-channel fp32 row.
// The joint table and the core-matrix images of W2 / W3 are built once per weight update by obs_fused_prep_kernel into
// one contiguous device buffer, which each CTA pulls into shared memory with a single bulk copy.
#include "obs_common.cuh"

namespace mpx {


// Shared-space loads for the conv1 gather: through the generic pointers derived from the aligned dynamic window the compiler
// emits generic LD.E (address-space check per access, lower rate); volatile keeps them behind the named barriers.
__device__ __forceinline__ uint32_t of_lds_u8(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float4 of_lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

constexpr int OF_AG = 32;              // agents per tile
constexpr int OF_WORKERS = 8;          // worker warps
#ifndef OF_NU
#define OF_NU 2                        // conv1 task groups (8 tasks x 4 channel quads) in flight per warp and iteration
#endif
constexpr int OF_THREADS = (OF_WORKERS + 1) * 32;

constexpr int OF_A1_BYTES = 26 * 1024;                 // 25 position tiles [32 agents x 16 ch] + 1 tile of slack
constexpr int OF_A2_BYTES = 16 * (OF_K3 / 8) * 128;    // 128 rows x 576 B = 73728
constexpr int OF_OBS_BYTES = OF_AG * OF_IH * OF_IW * OF_MAXK;   // 15488
constexpr int OF_CODE_BYTES = OF_AG * OF_IH * OF_IW;            // 3872: one joint code per cell
// Only rows 0..31 of the conv3 A operand are agents; the other 96 rows of the M=128 MMA are don't-care, so a1, the
// observation tile and the codes live inside that part of the a2 region (row groups 4..15).
constexpr int OF_A1_OFF = 4 * (OF_K3 / 8) * 128;                // 18432
constexpr int OF_OBS_OFF = OF_A1_OFF + OF_A1_BYTES;             // 45056
constexpr int OF_CODE_OFF = OF_OBS_OFF + OF_OBS_BYTES;          // 60544
static_assert(OF_CODE_OFF + OF_CODE_BYTES <= OF_A2_BYTES, "aliased buffers must fit in the unused a2 rows");
constexpr int OF_SMEM = OF_A2_BYTES + OF_IMG_BYTES + 128 + 1024;

struct ObsFusedParams {
  const uint8_t* obs;             // (M, 11, 11, K)
  int K;                          // components per cell
  int NJ;                         // joint codes per cell = prod(sizes[k] + 1)
  int sizes[OF_MAXK];             // one-hot size per component
  int radix[OF_MAXK];             // joint code = sum_k min(val_k, sizes[k]) * radix[k]
  const uint8_t* image;           // OF_IMG_BYTES, built by obs_fused_prep_kernel
  const float* reward;            // (M)
  const int32_t* last_action;     // (M)
  const int32_t* task_ids;        // (M) or null
  const float* w_r;               // (H)
  const float* b_r;               // (H)
  const float* a_table;           // (A, H)
  int A;
  const float* t_table;           // (n_tasks, H) or null
  int n_tasks;
  __nv_bfloat16* x;               // (M, H)
  __nv_bfloat16* obs_emb;         // (M, H) or null: conv3 output before the embedding sum (tests)
  long long M;
  unsigned long long* dbg;
};

// Builds the weight image (layout: obs_common.cuh)
__global__ void __launch_bounds__(256) obs_fused_prep_kernel(const __nv_bfloat16* __restrict__ w1, const __nv_bfloat16* __restrict__ b1,
                                                             const __nv_bfloat16* __restrict__ wt2, const __nv_bfloat16* __restrict__ b2,
                                                             const __nv_bfloat16* __restrict__ wt3, const __nv_bfloat16* __restrict__ b3,
                                                             ObsFusedParams p, uint8_t* __restrict__ image) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
  uint4* img4 = reinterpret_cast<uint4*>(image);
  for (int i = tid; i < OF_H * (OF_K3 / 8); i += nthr) {                  // wt3 (128, 288): 16-byte piece (n, kk)
    const int n = i / (OF_K3 / 8), kk = i % (OF_K3 / 8);
    img4[((n >> 3) * ((OF_K3 / 8) * 128) + kk * 128 + (n & 7) * 16) >> 4] =
        *reinterpret_cast<const uint4*>(wt3 + (long long)n * OF_K3 + kk * 8);
  }
  for (int i = tid; i < OF_C2 * 18; i += nthr) {                          // wt2 (32, 144): piece (n, kk), tap = kk / 2
    const int n = i / 18, kk = i % 18;
    img4[(OF_W3_BYTES + (kk >> 1) * 1024 + (n >> 3) * 256 + (kk & 1) * 128 + (n & 7) * 16) >> 4] =
        *reinterpret_cast<const uint4*>(wt2 + n * 144 + kk * 8);
  }
  float* t1 = reinterpret_cast<float*>(image + OF_W3_BYTES + OF_W2_BYTES);
  int C = 0;
  for (int k = 0; k < p.K; ++k) C += p.sizes[k];
  for (int i = tid; i < 9 * p.NJ * OF_C1; i += nthr) {
    const int ch = i % OF_C1, code = (i / OF_C1) % p.NJ, tap = i / (OF_C1 * p.NJ);
    float acc = 0.f;
    int off = 0, rem = code;
    for (int k = 0; k < p.K; ++k) {
      const int digit = rem % (p.sizes[k] + 1);
      rem /= (p.sizes[k] + 1);
      if (digit < p.sizes[k]) acc += __bfloat162float(w1[(tap * C + off + digit) * OF_C1 + ch]);
      off += p.sizes[k];
    }
    t1[((tap * p.NJ + code) * 2) * OF_C1 + ch] = acc;
    t1[((tap * p.NJ + code) * 2 + 1) * OF_C1 + ch] = acc;
  }
  __nv_bfloat16* bb = reinterpret_cast<__nv_bfloat16*>(image + OF_W3_BYTES + OF_W2_BYTES + OF_T1_BYTES);
  for (int i = tid; i < OF_C1 + OF_C2 + OF_H; i += nthr)
    bb[i] = i < OF_C1 ? b1[i] : i < OF_C1 + OF_C2 ? b2[i - OF_C1] : b3[i - OF_C1 - OF_C2];
}

__global__ void __launch_bounds__(OF_THREADS, 1) obs_fused_kernel(const ObsFusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA2 = smem;
  uint8_t* sA1 = sA2 + OF_A1_OFF;
  uint8_t* sImg = sA2 + OF_A2_BYTES;
  uint8_t* sW3 = sImg;
  uint8_t* sW2 = sW3 + OF_W3_BYTES;
  const float* sT1 = reinterpret_cast<const float*>(sW2 + OF_W2_BYTES);
  const __nv_bfloat16* sB = reinterpret_cast<const __nv_bfloat16*>(sW2 + OF_W2_BYTES + OF_T1_BYTES);   // b1[16] b2[32] b3[128]
  uint8_t* sObs = sA2 + OF_OBS_OFF;
  uint8_t* sCode = sA2 + OF_CODE_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sImg + OF_IMG_BYTES);
  uint64_t* a1_ready = bars;
  uint64_t* c2_full = bars + 1;
  uint64_t* a2_ready = bars + 2;
  uint64_t* c3_full = bars + 3;
  uint64_t* img_full = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  float* sOut = reinterpret_cast<float*>(sA1);               // conv3 accumulator staging, aliases a1 (dead by then)

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int tid = threadIdx.x;
  const int n_tiles = (int)((p.M + OF_AG - 1) / OF_AG);
  const int cell_bytes = OF_IH * OF_IW * p.K;

  unsigned long long* dbg = (blockIdx.x == 0 && tid == 0) ? p.dbg : nullptr;
  dbg_stamp(dbg, 0);
  griddep_launch();
  if (tid == 0) {
    mbar_init(a1_ready, OF_WORKERS);
    mbar_init(c2_full, 1);
    mbar_init(a2_ready, OF_WORKERS);
    mbar_init(c3_full, 1);
    mbar_init(img_full, 1);
    fence_barrier_init();
    griddep_wait();                                          // the image may have been rebuilt by the previous kernel
    mbar_arrive_expect_tx(img_full, OF_IMG_BYTES);           // weights + tables: one bulk copy per CTA
    bulk_load_1d(sImg, p.image, OF_IMG_BYTES, img_full);
  }
  if (warp == OF_WORKERS) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  dbg_stamp(dbg, 1);
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t tD2 = tmem_base;            // 3 x 32 columns
  const uint32_t tD3 = tmem_base + 128;      // 128 columns
  const uint32_t lboA = 128u, sboA1 = 256u;
  const uint32_t sbo3 = (OF_K3 / 8) * 128;   // 4608

  if (warp == OF_WORKERS) {
    // ================================ MMA issuer ================================
    // the whole warp runs the (uniform) control flow; one elected lane issues each of the 45 MMAs of a tile
    {
      constexpr uint32_t idesc2 = umma_idesc_bf16(128, OF_C2, 0, 0);
      constexpr uint32_t idesc3 = umma_idesc_bf16(128, OF_H, 0, 0);
      uint32_t ph = 0;
      mbar_wait(img_full, 0);
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        mbar_wait(a1_ready, ph);
        tc_fence_after();
        const uint32_t a1 = smem_u32(sA1), w2 = smem_u32(sW2);
#pragma unroll 1
        for (int j = 0; j < 3; ++j) {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int q0 = (j + tap / 3) * OF_P1 + (tap % 3);
            umma_bf16_1(tD2 + j * OF_C2, umma_smem_desc_nosw(a1 + q0 * 1024, lboA, sboA1),
                      umma_smem_desc_nosw(w2 + tap * 1024, lboA, sboA1), idesc2, tap != 0 ? 1u : 0u);
          }
        }
        umma_commit_1(c2_full);
        mbar_wait(a2_ready, ph);
        tc_fence_after();
        const uint32_t a2 = smem_u32(sA2), w3 = smem_u32(sW3);
#pragma unroll
        for (int s = 0; s < OF_K3 / 16; ++s)
          umma_bf16_1(tD3, umma_smem_desc_nosw(a2 + s * 256, 128u, sbo3),
                    umma_smem_desc_nosw(w3 + s * 256, 128u, sbo3), idesc3, s != 0 ? 1u : 0u);
        umma_commit_1(c3_full);
        ph ^= 1;
      }
    }
  } else {
    // ================================ worker warps ================================
    const int wtid = tid;                                   // 0..255
    const int quarter = warp & 3, chalf = warp >> 2;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float sqH = bf16_round(sqrtf((float)OF_H));
    uint32_t ph = 0;
    // embedding-sum role of this thread: (agent eag, columns [16*ecg, +16)); reward encoder weights are tile-invariant
    const int eag = wtid >> 3, ecg = wtid & 7;
    griddep_wait();                                          // observations / rewards / actions come from earlier kernels
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long m0 = (long long)tile * OF_AG;
      const int nag = (int)min((long long)OF_AG, p.M - m0);
      // ---- this thread's embedding inputs (network.py:304-309): scalars first, so the table rows can be requested
      //      as soon as the observation tile is on its way
      float e_rw = 0.f;
      int e_a = -1, e_tk = -1;
      if (eag < nag) {
        const long long m = m0 + eag;
        e_rw = p.reward[m];
        e_a = p.last_action[m];
        if (p.task_ids && p.t_table) e_tk = p.task_ids[m];
      }
      // ---- observation tile -> smem (contiguous bytes)
      {
        const int nbytes = nag * cell_bytes;
        const uint8_t* src = p.obs + m0 * cell_bytes;
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
          const int nw = nbytes >> 4;
          for (int i = wtid; i < nw; i += OF_WORKERS * 32)
            reinterpret_cast<uint4*>(sObs)[i] = reinterpret_cast<const uint4*>(src)[i];
          for (int i = (nw << 4) + wtid; i < nbytes; i += OF_WORKERS * 32) sObs[i] = src[i];
        } else {
          for (int i = wtid; i < nbytes; i += OF_WORKERS * 32) sObs[i] = src[i];
        }
      }
      of_bar_sync(1, OF_WORKERS * 32);
      dbg_stamp(dbg, 2);
      // ---- one joint code per cell: sum_k min(value_k, size_k) * radix_k  (size_k = "out of range" digit); 4 cells / thread
      {
        const int ncell = nag * OF_IH * OF_IW;
        const int K = p.K;
        const int s0 = p.sizes[0], s1 = p.sizes[1], s2 = p.sizes[2], s3 = p.sizes[3];
        const int r1 = p.radix[1], r2 = p.radix[2], r3 = p.radix[3];
        for (int i4 = wtid; i4 * 4 < ncell; i4 += OF_WORKERS * 32) {
          uint32_t packed = 0;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = i4 * 4 + e;
            if (i < ncell) {
              const uint8_t* cell = sObs + i * K;
              int code = min((int)cell[0], s0);
              if (K > 1) code += min((int)cell[1], s1) * r1;
              if (K > 2) code += min((int)cell[2], s2) * r2;
              if (K > 3) code += min((int)cell[3], s3) * r3;
              packed |= (uint32_t)code << (8 * e);
            }
          }
          reinterpret_cast<uint32_t*>(sCode)[i4] = packed;
        }
      }
      if (tile == (int)blockIdx.x) mbar_wait(img_full, 0);   // tables and weights have landed
      of_bar_sync(1, OF_WORKERS * 32);
      dbg_stamp(dbg, 3);
      // ---- this thread's embedding addends (network.py:304-309): the table rows are REQUESTED here - after the observation
      //      tile and the cell codes, so that the dependent round trip (last action -> table row) is off the start-up path -
      //      and consumed after conv1, whose gather they fly under
      float4 e_av[4], e_tv[4], e_wv[4], e_bq[4];
      {
        int a = e_a, tk = e_tk;
        if (eag < nag) {
          if (a < 0) a += p.A;
          if (a >= p.A) a = -1;
          if (p.task_ids && p.t_table) {
            if (tk < 0) tk += p.n_tasks;
            if (tk >= p.n_tasks) tk = -1;
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          e_av[i] = e_tv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a >= 0) e_av[i] = reinterpret_cast<const float4*>(p.a_table + (long long)a * OF_H + ecg * 16)[i];
          if (tk >= 0) e_tv[i] = reinterpret_cast<const float4*>(p.t_table + (long long)tk * OF_H + ecg * 16)[i];
          e_wv[i] = reinterpret_cast<const float4*>(p.w_r + ecg * 16)[i];
          e_bq[i] = reinterpret_cast<const float4*>(p.b_r + ecg * 16)[i];
        }
      }
      // ---- conv1: lane = (task tsub of 8, channel quad v); 9 conflict-free float4 lookups per task
      {
        const int tsub = lane >> 2, v = lane & 3;
        const uint32_t tbase = smem_u32(sT1) + ((tsub & 1) * OF_C1 + 4 * v) * 4;
        const uint32_t code_s = smem_u32(sCode);
        const int NJ = p.NJ;
        const uint2 bv = *reinterpret_cast<const uint2*>(sB + 4 * v);
        for (int g0 = warp; g0 < OF_AG * 25 / 8; g0 += OF_NU * OF_WORKERS) {  // OF_NU task groups in flight per iteration
          // all code bytes first, then all table rows, then the sums: the loads are volatile (ordered), so a tap-by-tap
          // code -> row -> add sequence would serialise 18 dependent shared-memory round trips per iteration
          float4 acc[OF_NU];
          int agq[OF_NU][2];
          uint32_t codes[OF_NU][9];
          bool on[OF_NU];
#pragma unroll
          for (int u = 0; u < OF_NU; ++u) {
            const int g = g0 + u * OF_WORKERS;
            const int task = g * 8 + tsub;
            const int ag = task & 31, q = task >> 5;
            agq[u][0] = ag; agq[u][1] = q;
            on[u] = g < OF_AG * 25 / 8 && ag < nag;
            const int oy = q / OF_P1, ox = q % OF_P1;
            const uint32_t c0 = code_s + (on[u] ? ag * (OF_IH * OF_IW) + (oy * 2) * OF_IW + ox * 2 : 0);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) codes[u][tap] = of_lds_u8(c0 + (tap / 3) * OF_IW + (tap % 3));
          }
          float4 wrow[OF_NU][9];
#pragma unroll
          for (int u = 0; u < OF_NU; ++u)
#pragma unroll
            for (int tap = 0; tap < 9; ++tap)
              wrow[u][tap] = of_lds_f4(tbase + (uint32_t)(tap * NJ + (on[u] ? (int)codes[u][tap] : 0)) * (2 * OF_C1 * 4));
#pragma unroll
          for (int u = 0; u < OF_NU; ++u) {
            acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (on[u]) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                acc[u].x += wrow[u][tap].x; acc[u].y += wrow[u][tap].y; acc[u].z += wrow[u][tap].z; acc[u].w += wrow[u][tap].w;
              }
            }
          }
#pragma unroll
          for (int u = 0; u < OF_NU; ++u) {
            if (g0 + u * OF_WORKERS < OF_AG * 25 / 8) {
              const int ag = agq[u][0], q = agq[u][1];
              const uint32_t z0 = pack_bf16(acc[u].x, acc[u].y), z1 = pack_bf16(acc[u].z, acc[u].w);
              const uint32_t y0 = bf16x2_add(z0, bv.x);
              const uint32_t y1 = bf16x2_add(z1, bv.y);
              uint2 o;
              o.x = gelu_tanh_bf16x2(y0);
              o.y = gelu_tanh_bf16x2(y1);
              *reinterpret_cast<uint2*>(sA1 + q * 1024 + (ag >> 3) * 256 + (v >> 1) * 128 + (ag & 7) * 16 + (v & 1) * 8) = o;
            }
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      dbg_stamp(dbg, 4);
      if (lane == 0) mbar_arrive(a1_ready);
      // ---- the embedding addends, packed (their loads were issued in front of the conv1 gather)
      uint32_t re_w[8], ae_w[8], te_w[8];
      {
        const float rw = bf16_round(e_rw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 av = e_av[i], tv = e_tv[i], wv = e_wv[i], bq = e_bq[i];
          const float avv[4] = {av.x, av.y, av.z, av.w}, tvv[4] = {tv.x, tv.y, tv.z, tv.w};
          const float wr[4] = {bf16_round(wv.x), bf16_round(wv.y), bf16_round(wv.z), bf16_round(wv.w)};
          const float br[4] = {bf16_round(bq.x), bf16_round(bq.y), bf16_round(bq.z), bf16_round(bq.w)};
#pragma unroll
          for (int e = 0; e < 4; e += 2) {
            const int h = 4 * i + e;
            re_w[h >> 1] = pack_bf16(bf16_round(rw * wr[e]) + br[e], bf16_round(rw * wr[e + 1]) + br[e + 1]);
            ae_w[h >> 1] = pack_bf16(bf16_round(avv[e]) * sqH, bf16_round(avv[e + 1]) * sqH);
            te_w[h >> 1] = pack_bf16(bf16_round(tvv[e]) * sqH, bf16_round(tvv[e + 1]) * sqH);
          }
        }
      }
      // ---- conv2 epilogue: rows (px = quarter, agent = lane), this warp's 16 of the 32 channels -> a2
      mbar_wait(c2_full, ph);
      tc_fence_after();
      dbg_stamp(dbg, 5);
      if (quarter < 3) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          uint32_t d[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]),
                "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
              : "r"(tD2 + lane_off + j * OF_C2 + chalf * 16)
              : "memory");
          tmem_ld_wait();
          const uint4* b4 = reinterpret_cast<const uint4*>(sB + 16 + chalf * 16);
          const int pos = 3 * j + quarter;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const uint4 bv = b4[hh];
            const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const uint32_t zw = pack_bf16(__uint_as_float(d[8 * hh + 2 * e]), __uint_as_float(d[8 * hh + 2 * e + 1]));
              const uint32_t zb = bf16x2_add(zw, bw[e]);
              o[e] = gelu_tanh_bf16x2(zb);
            }
            const int kk = pos * 4 + chalf * 2 + hh;        // 8-wide K core matrix index: k = pos*32 + chalf*16 + hh*8
            *reinterpret_cast<uint4*>(sA2 + (lane >> 3) * sbo3 + kk * 128 + (lane & 7) * 16) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      dbg_stamp(dbg, 6);
      if (lane == 0) mbar_arrive(a2_ready);
      // ---- conv3 accumulator (TMEM lanes 0..31 = agents) -> fp32 staging
      mbar_wait(c3_full, ph);
      tc_fence_after();
      dbg_stamp(dbg, 7);
      if (quarter == 0) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t acc[32];
          tmem_ld_32x32(tD3 + chalf * 64 + cc * 32, acc);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(sOut + lane * OF_H + chalf * 64 + cc * 32 + 4 * (i ^ (lane & 7))) =
                make_float4(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1]), __uint_as_float(acc[4 * i + 2]),
                            __uint_as_float(acc[4 * i + 3]));
        }
        tc_fence_before();
      }
      of_bar_sync(1, OF_WORKERS * 32);
      dbg_stamp(dbg, 8);
      // ---- bias + embedding sum (network.py:303-309): thread = (agent eag, 16 columns)
      if (eag < nag) {
        const long long m = m0 + eag;
        uint32_t ow[8], ew[8];
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int c4 = ecg * 4 + i4;                         // float4 index within the row
          const int blk = c4 >> 3, i = c4 & 7;                 // 32-column block, piece inside it (swizzled)
          const float4 vq = *reinterpret_cast<const float4*>(sOut + eag * OF_H + blk * 32 + 4 * (i ^ (eag & 7)));
          const uint2 b3 = *reinterpret_cast<const uint2*>(sB + 48 + c4 * 4);
          const uint32_t c0 = pack_bf16(vq.x, vq.y), c1 = pack_bf16(vq.z, vq.w);              // conv3 output -> bf16
          const uint32_t e0 = bf16x2_add(c0, b3.x);   // + bias
          const uint32_t e1 = bf16x2_add(c1, b3.y);
          ew[2 * i4] = e0; ew[2 * i4 + 1] = e1;
          uint32_t x0 = bf16x2_add(e0, re_w[2 * i4]);
          uint32_t x1 = bf16x2_add(e1, re_w[2 * i4 + 1]);
          x0 = bf16x2_add(x0, ae_w[2 * i4]);
          x1 = bf16x2_add(x1, ae_w[2 * i4 + 1]);
          if (p.task_ids && p.t_table) {
            x0 = bf16x2_add(x0, te_w[2 * i4]);
            x1 = bf16x2_add(x1, te_w[2 * i4 + 1]);
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
      of_bar_sync(1, OF_WORKERS * 32);                       // sOut (aliasing a1) and sObs are free for the next tile
      dbg_stamp(dbg, 9);
      ph ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == OF_WORKERS) tmem_dealloc(tmem_base, 256);
}

}  // namespace mpx

using namespace mpx;

static int of_fill_spec(ObsFusedParams& p, const int* sizes, int K) {
  ObsCodeSpec s;
  int rc = obs_fill_spec(s, sizes, K, "mpx_obs_embed_fused");
  if (rc) return rc;
  p.K = s.K;
  p.NJ = s.NJ;
  for (int i = 0; i < OF_MAXK; ++i) {
    p.sizes[i] = s.sizes[i];
    p.radix[i] = s.radix[i];
  }
  return MPX_OK;
}

extern "C" size_t mpx_obs_fused_image_bytes(void) { return (size_t)OF_IMG_BYTES + obs_train_bwd_image_bytes(); }

extern "C" int mpx_obs_fused_prep(const int* sizes, int K, const mpx_bf16* w1, const mpx_bf16* b1, const mpx_bf16* wt2,
                                  const mpx_bf16* b2, const mpx_bf16* wt3, const mpx_bf16* b3, void* image,
                                  mpx_stream_t stream) {
  MPX_REQUIRE(w1 && b1 && wt2 && b2 && wt3 && b3 && image, "mpx_obs_fused_prep: null pointer");
  MPX_REQUIRE((reinterpret_cast<uintptr_t>(image) & 15) == 0 && (reinterpret_cast<uintptr_t>(wt2) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(wt3) & 15) == 0,
              "mpx_obs_fused_prep: image / wt2 / wt3 must be 16-byte aligned");
  ObsFusedParams p = {};
  int rc = of_fill_spec(p, sizes, K);
  if (rc) return rc;
  obs_fused_prep_kernel<<<32, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(w1), reinterpret_cast<const __nv_bfloat16*>(b1),
      reinterpret_cast<const __nv_bfloat16*>(wt2), reinterpret_cast<const __nv_bfloat16*>(b2),
      reinterpret_cast<const __nv_bfloat16*>(wt3), reinterpret_cast<const __nv_bfloat16*>(b3), p,
      reinterpret_cast<uint8_t*>(image));
  rc = check_launch("obs_fused_prep_kernel");
  if (rc) return rc;
  return obs_train_prep_bwd(reinterpret_cast<const __nv_bfloat16*>(wt2), reinterpret_cast<const __nv_bfloat16*>(wt3),
                            reinterpret_cast<uint8_t*>(image) + OF_IMG_BYTES, (cudaStream_t)stream);
}

extern "C" int mpx_obs_embed_fused(const uint8_t* obs, const int* sizes, int K, const void* image, const float* reward,
                                   const int32_t* last_action, const int32_t* task_ids, const float* w_reward,
                                   const float* b_reward, const float* action_table, int A, const float* task_table,
                                   int n_tasks, mpx_bf16* x, mpx_bf16* obs_emb, long long M, int IH, int IW, int H,
                                   mpx_stream_t stream) {
  MPX_REQUIRE(obs && image && reward && last_action && w_reward && b_reward && action_table && x && M > 0 && A > 0,
              "mpx_obs_embed_fused: bad arguments");
  MPX_REQUIRE(IH == OF_IH && IW == OF_IW && H == OF_H, "mpx_obs_embed_fused: only the 11x11 view / H=128 geometry is fused");
  MPX_REQUIRE((reinterpret_cast<uintptr_t>(image) & 15) == 0, "mpx_obs_embed_fused: image must be 16-byte aligned");
  ObsFusedParams p = {};
  int rc = of_fill_spec(p, sizes, K);
  if (rc) return rc;
  p.obs = obs;
  p.image = reinterpret_cast<const uint8_t*>(image);
  p.reward = reward; p.last_action = last_action; p.task_ids = task_ids;
  p.w_r = w_reward; p.b_r = b_reward; p.a_table = action_table; p.A = A;
  p.t_table = task_table; p.n_tasks = n_tasks;
  p.x = reinterpret_cast<__nv_bfloat16*>(x);
  p.obs_emb = reinterpret_cast<__nv_bfloat16*>(obs_emb);
  p.M = M;
  p.dbg = g_dbg_stamps;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(obs_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OF_SMEM);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(obs_fused): %s", cudaGetErrorString(e));
      return MPX_ERR_CUDA;
    }
    attr_set = true;
  }
  const long long tiles = (M + OF_AG - 1) / OF_AG;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  cudaError_t le = launch_ex(obs_fused_kernel, dim3(grid), dim3(OF_THREADS), OF_SMEM, (cudaStream_t)stream, 0, 2, p);
  if (le != cudaSuccess) {
    set_error("launch obs_fused_kernel: %s", cudaGetErrorString(le));
    return MPX_ERR_CUDA;
  }
  return check_launch("obs_fused_kernel");
}
