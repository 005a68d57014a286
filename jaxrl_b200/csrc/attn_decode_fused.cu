// Rollout decode step, attention half of a TransformerBlock in ONE kernel (reference network.py:158-165 with the
// decode branch of AttentionBlock.__call__, attention.py:96-150):
//     xn = RMSNorm(x)   q,k,v = xn Wq, xn Wk, xn Wv   q,k = RoPE(q,k; pos)   cache[pos % S] = k,v
//     out = softmax(q . K[0:kv_len] / sqrt(D)) . V[0:kv_len]            kv_len = min(pos[0]+1, S)
// The separate projection kernel (mpx_qkv_rope) is folded into the front of the cache-streaming kernel
// (attn_decode_online_kernel, attn_decode.cu): the K/V stream of the agent's ring starts at kernel entry and the
// projection of the CTA's agents runs underneath it - one m16 x n192 x k128 mma.sync GEMM per round, the weight
// fragments held in registers for the whole kernel.  The new k,v go to the ring in global memory (for later steps)
// AND are patched into the shared-memory stage of the chunk that holds slot pos % S, so this step attends to them
// without a round trip.  Shapes: hidden 128 (return_*_layers, blog_32_layers) or 256 (multitask), 4 query heads on 1 kv
// head, head_dim 32; everything else takes mpx_qkv_rope + mpx_attn_decode.  At hidden 256 the weight fragments (64
// registers per lane) are re-read from L2 at the top of every round (a launch has two rounds at 4096 agents) and are dead
// before the cache loop begins, so the CTA keeps its 14 warps.
//
// One warp per agent, W (12..14) agents per CTA per round.  Shared memory = W rings of 4 stages x 4 KB (K chunk |
// V chunk of 32 positions).  Stage 3 of the first rings doubles as the projection scratch (normalised x rows - 4 KB at
// hidden 128, 8 KB over two rings at hidden 256 -, q rows, new k|v rows) while only stages 0-2 are in flight.
#include "common.cuh"
#include "ptx.cuh"

namespace mpx {

namespace {

constexpr int AF_STAGES = 4;
constexpr int AF_CHUNK = 32;
constexpr int AF_HALF = AF_CHUNK * 64;          // K (or V) part of a stage
constexpr int AF_STAGE_BYTES = 2 * AF_HALF;
constexpr int AF_RING = AF_STAGES * AF_STAGE_BYTES;

__device__ __forceinline__ void af_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void af_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint32_t af_off64(int row, int piece) { return row * 64 + ((piece ^ ((row >> 1) & 3)) << 4); }
__device__ __forceinline__ void af_ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void af_ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void af_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t af_movmatrix_t(uint32_t x) {
  uint32_t y;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ uint32_t af_lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void af_sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void af_sts64(uint32_t addr, uint32_t v0, uint32_t v1) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(v0), "r"(v1) : "memory");
}

struct AttnFusedParams {
  const __nv_bfloat16* x;          // (B,H) block input (residual stream), H = 16 HK
  const float* norm_scale;         // (H) history_norm
  float eps;
  const __nv_bfloat16* wt_qkv;     // (192,H): rows = [4 q heads x 32 | k 32 | v 32] output features
  const int32_t* pos;              // (B)
  const float* timescale;          // (16)
  __nv_bfloat16* kcache;           // (B,S,1,32)
  __nv_bfloat16* vcache;
  __nv_bfloat16* out;              // (B,128)
  int B, S;
  float scale_log2;
  int early_stream;                // 1: ring slots other than pos % S are stable - their stream may start before griddepcontrol.wait
};

// one 32-position chunk of an agent's K and V rings -> stage (c % 4) of the warp's shared-memory ring (one commit group)
__device__ __forceinline__ void af_issue_chunk(uint32_t ring, const __nv_bfloat16* kbase, const __nv_bfloat16* vbase, int c,
                                               int nchunks, int kv_len, bool active, int lane) {
  if (active && c < nchunks) {
    const uint32_t st = ring + (c % AF_STAGES) * AF_STAGE_BYTES;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int pi = lane + 32 * (i & 3);
      const int r = pi >> 2, j = pi & 3;
      const int pp = c * AF_CHUNK + r;
      const bool valid = pp < kv_len;
      const __nv_bfloat16* gp = (i < 4 ? kbase : vbase) + (size_t)(valid ? pp : 0) * 32 + j * 8;
      const uint32_t n = valid ? 16u : 0u;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;"
                   ::"r"(st + (i < 4 ? 0 : AF_HALF) + af_off64(r, j)), "l"(gp), "r"(n) : "memory");
    }
  }
  af_commit();
}

// HK = hidden / 16 (k-steps of the projection): 8 or 16
template <int HK>
__global__ void __launch_bounds__(448) attn_decode_fused_kernel(const AttnFusedParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int H = 16 * HK;
  constexpr int XROW = 2 * H;                   // bytes of a normalised x row
  constexpr int XRINGS = HK / 8;                // rings whose stage 3 holds x rows (8 rows of 512 B or 16 rows of 256 B each)
  constexpr int XROWS_PER_RING = AF_STAGE_BYTES / XROW;
  const int warps = blockDim.x >> 5;            // 12..14
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem0 = smem_u32(smem_raw);
  const uint32_t ring = smem0 + warp * AF_RING;
  // normalised x: 16 rows x XROW bytes, 16-byte chunk c of row r at chunk position c ^ xs_swz(r): the 8 lanes of a
  // quarter warp (rows 2q, 2q+1; 4 chunks each) of the A-fragment loads hit 8 different bank groups
  auto xs_swz = [](int r) { return ((r & 1) << 2) | ((r >> 1) & 3); };
  auto xs_row = [&](int r) { return smem0 + (r / XROWS_PER_RING) * AF_RING + 3 * AF_STAGE_BYTES + (r % XROWS_PER_RING) * XROW; };
  const uint32_t qs = smem0 + XRINGS * AF_RING + 3 * AF_STAGE_BYTES;          // W rows x 256 B (q, 4 heads x 32)
  const uint32_t kvs = smem0 + (XRINGS + 1) * AF_RING + 3 * AF_STAGE_BYTES;   // W rows x 128 B (new k | new v)
  const int g = lane >> 2, t = lane & 3, mi = lane >> 3, r8 = lane & 7;
  constexpr int G = 4;

  // ---- projection weights: warp w < 12 owns output columns [n0, n0+8) and [n0+16, n0+24) of head-slot w/2 (the two
  //      halves of a RoPE pair live in the same lane).  B fragments stay in registers for the whole kernel.
  const int he = warp >> 1, jp = warp & 1;      // head slot: 0-3 q heads, 4 k, 5 v
  // B fragments as 16-byte loads: lane (g,t) holds W[n][32 kb + 8t .. +7] per 32-wide k block kb - words 0,1 are the
  // fragment of the block's first m16n8k16 step, words 2,3 of its second.  That is a permutation of k inside the block
  // (logical k 2t+i -> 8t+i, 2t+8+i -> 8t+2+i; second step +4); the A fragments below read x under the same permutation
  // (one 16-byte shared load per row and block), and a contraction does not care.
  uint4 bw[2][HK / 2];
  auto load_w = [&]() {
    if (warp < 12) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const __nv_bfloat16* wrow = p.wt_qkv + (size_t)(he * 32 + 8 * jp + 16 * j + g) * H + 8 * t;
#pragma unroll
        for (int kb = 0; kb < HK / 2; ++kb) bw[j][kb] = __ldg(reinterpret_cast<const uint4*>(wrow + 32 * kb));
      }
    }
  };
  load_w();                                      // weights are parameters: never produced by the preceding kernel
  float nscale[HK / 2];                          // this lane's HK/2 consecutive features
#pragma unroll
  for (int i = 0; i < HK / 8; ++i) {
    const float4 v4 = *reinterpret_cast<const float4*>(p.norm_scale + (HK / 2) * lane + 4 * i);
    nscale[4 * i] = v4.x; nscale[4 * i + 1] = v4.y; nscale[4 * i + 2] = v4.z; nscale[4 * i + 3] = v4.w;
  }
  float ts0 = 0.f, ts1 = 0.f;
  if (warp < 10) {                               // q and k columns: rotation frequencies 8 jp + 2t, +1
    ts0 = p.timescale[8 * jp + 2 * t];
    ts1 = p.timescale[8 * jp + 2 * t + 1];
  }
  const int pos0 = p.pos[0];                     // positions are step inputs, never produced by the preceding kernel
  griddep_launch();
  const int per_round = gridDim.x * warps;
  const int nrounds = (p.B + per_round - 1) / per_round;
  uint32_t xv[HK / 4];                           // HK/2 bf16 of this warp's row
#pragma unroll
  for (int i = 0; i < HK / 4; ++i) xv[i] = 0u;
  float rsn[2][2], rcs[2][2];
  auto fetch_x = [&](int round) {
    const int b0 = round * per_round + blockIdx.x * warps;
    if (round < nrounds && b0 + warp < p.B) {
      const __nv_bfloat16* xp = p.x + (size_t)(b0 + warp) * H + (HK / 2) * lane;
      if constexpr (HK == 8) {
        const uint2 v = *reinterpret_cast<const uint2*>(xp);
        xv[0] = v.x; xv[1] = v.y;
      } else {
        const uint4 v = *reinterpret_cast<const uint4*>(xp);
        xv[0] = v.x; xv[1] = v.y; xv[2] = v.z; xv[3] = v.w;
      }
    }
  };
  auto fetch_angles = [&](int round) {
    const int b0 = round * per_round + blockIdx.x * warps;
    if (round < nrounds && warp < 10) {
#pragma unroll
      for (int rs = 0; rs < 2; ++rs) {
        const int ab = b0 + g + 8 * rs;
        const float pf = (float)p.pos[ab < p.B ? ab : 0];
        rope_sincos(pf / ts0, &rsn[rs][0], &rcs[rs][0]);
        rope_sincos(pf / ts1, &rsn[rs][1], &rcs[rs][1]);
      }
    }
  };
  fetch_angles(0);                               // positions only: runs under the preceding kernel's tail
  const int kv_len = min(pos0 + 1, p.S);
  const int nchunks = (kv_len + AF_CHUNK - 1) / AF_CHUNK;
  int slot = pos0 % p.S;
  if (slot < 0) slot += p.S;
  // The caller vouches (early_stream) that every ring slot except pos % S was written by kernels that completed
  // before this one could start (the rollout: earlier decode steps, with fully serialised launches in between).  Round
  // 0's first chunks then leave for shared memory while the preceding kernel is still draining; the slot this step
  // writes is patched into its stage below, whatever the early load brought for it.
  if (p.early_stream) {
    const int b = blockIdx.x * warps + warp;
    const bool active = b < p.B;
    const __nv_bfloat16* kbase = p.kcache + (size_t)(active ? b : 0) * p.S * 32;
    const __nv_bfloat16* vbase = p.vcache + (size_t)(active ? b : 0) * p.S * 32;
#pragma unroll
    for (int i = 0; i < AF_STAGES - 1; ++i) af_issue_chunk(ring, kbase, vbase, i, nchunks, kv_len, active, lane);
  }
  griddep_wait();                                // x and the rings are written by the previous kernels
  const int c_new = slot >> 5, r_new = slot & 31;
  const bool head_lane = (2 * t) < G;
  const bool head1 = (2 * t + 1) < G;

  // this warp's input row and the rotation angles of the projection rows it finishes (agents g, g + 8 of the CTA), for
  // the round about to start: fetched one round ahead so that their latency hides behind the previous round's stream
  fetch_x(0);

  for (int round = 0; round < nrounds; ++round) {
    const int cta_b0 = round * per_round + blockIdx.x * warps;      // first agent of this CTA in this round
    const int b = cta_b0 + warp;
    const bool active = b < p.B;
    const __nv_bfloat16* kbase = p.kcache + (size_t)(active ? b : 0) * p.S * 32;
    const __nv_bfloat16* vbase = p.vcache + (size_t)(active ? b : 0) * p.S * 32;
    auto issue = [&](int c) { af_issue_chunk(ring, kbase, vbase, c, nchunks, kv_len, active, lane); };
    // ---- 1. the cache stream starts first: chunks 0..2 (the stale row at slot pos % S is patched below)
    if (!(round == 0 && p.early_stream)) {
#pragma unroll
      for (int i = 0; i < AF_STAGES - 1; ++i) issue(i);
    }
    // every warp has left the previous round's loop: stage 3 of the first rings is free for the scratch rows
    if (round > 0) __syncthreads();
    // ---- 2. RMSNorm of this agent's row -> xs[warp] (the row and the rotation angles were fetched a round ahead)
    {
      uint32_t y[HK / 4];
#pragma unroll
      for (int i = 0; i < HK / 4; ++i) y[i] = 0u;
      if (active) {
        float ssl = 0.f;
#pragma unroll
        for (int i = 0; i < HK / 4; ++i) {
          const float a = bf16_lo(xv[i]), b = bf16_hi(xv[i]);
          ssl += a * a + b * b;
        }
        const float ss = warp_sum(ssl);
        const float inv = rsqrtf(ss * (1.0f / (float)H) + p.eps);
#pragma unroll
        for (int i = 0; i < HK / 4; ++i)
          y[i] = pack_bf16(bf16_lo(xv[i]) * (inv * nscale[2 * i]), bf16_hi(xv[i]) * (inv * nscale[2 * i + 1]));
      }
      if constexpr (HK == 8) af_sts64(xs_row(warp) + (((lane >> 1) ^ xs_swz(warp)) << 4) + 8 * (lane & 1), y[0], y[1]);
      else sts_v4(xs_row(warp) + ((lane ^ xs_swz(warp)) << 4), make_uint4(y[0], y[1], y[2], y[3]));
    }
    __syncthreads();
    // ---- 3. q,k,v of the CTA's agents: C[16 agents, 16 columns] per warp
    if (warp < 12) {
      float acc[2][4];
#pragma unroll
      for (int j = 0; j < 2; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
      const bool row_hi_ok = g + 8 < warps;
      const uint32_t xr0 = xs_row(g), xr1 = xs_row(g + 8);
      const int fg = xs_swz(g);                           // == xs_swz(g + 8)
#pragma unroll
      for (int kb = 0; kb < HK / 2; ++kb) {
        const uint32_t c = (uint32_t)(((4 * kb + t) ^ fg) << 4);
        const uint4 lo = lds_v4(xr0 + c);
        const uint4 hi = row_hi_ok ? lds_v4(xr1 + c) : make_uint4(0u, 0u, 0u, 0u);
        const uint32_t a0[4] = {lo.x, hi.x, lo.y, hi.y}, a1[4] = {lo.z, hi.z, lo.w, hi.w};
        af_mma(acc[0], a0, bw[0][kb].x, bw[0][kb].y);
        af_mma(acc[1], a0, bw[1][kb].x, bw[1][kb].y);
        af_mma(acc[0], a1, bw[0][kb].z, bw[0][kb].w);
        af_mma(acc[1], a1, bw[1][kb].z, bw[1][kb].w);
      }
#pragma unroll
      for (int rs = 0; rs < 2; ++rs) {                    // agent rows g and g + 8
        const int arow = g + 8 * rs;
        const int ab = cta_b0 + arow;
        if (arow < warps && ab < p.B) {
          float lo0 = acc[0][2 * rs], lo1 = acc[0][2 * rs + 1], hi0 = acc[1][2 * rs], hi1 = acc[1][2 * rs + 1];
          bf16_round2(lo0, lo1);                          // the projection output is bf16 (attention.py:100-107)
          bf16_round2(hi0, hi1);
          if (he < 5) {                                   // RoPE on q and k: (d, d + 16) pairs
            const float s0 = rsn[rs][0], c0 = rcs[rs][0], s1 = rsn[rs][1], c1 = rcs[rs][1];
            const float a0 = lo0 * c0 - hi0 * s0, b0 = hi0 * c0 + lo0 * s0;
            const float a1 = lo1 * c1 - hi1 * s1, b1 = hi1 * c1 + lo1 * s1;
            lo0 = a0; hi0 = b0; lo1 = a1; hi1 = b1;
          }
          const uint32_t wlo = pack_bf16(lo0, lo1), whi = pack_bf16(hi0, hi1);
          const int d = 8 * jp + 2 * t;                   // column inside the head (low half; high half = d + 16)
          if (he < 4) {
            af_sts32(qs + arow * 256 + he * 64 + d * 2, wlo);
            af_sts32(qs + arow * 256 + he * 64 + (d + 16) * 2, whi);
          } else {
            const uint32_t kv = kvs + arow * 128 + (he - 4) * 64;
            af_sts32(kv + d * 2, wlo);
            af_sts32(kv + (d + 16) * 2, whi);
            __nv_bfloat16* crow = (he == 4 ? p.kcache : p.vcache) + ((size_t)ab * p.S + slot) * 32;
            *reinterpret_cast<uint32_t*>(crow + d) = wlo;
            *reinterpret_cast<uint32_t*>(crow + d + 16) = whi;
          }
        }
      }
    }
    __syncthreads();
    // ---- 4. this warp's agent: q fragments and the new k|v row (8 lanes x 16 B) into registers
    uint32_t qb[2][2];
    uint4 kvnew = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      qb[ks][0] = g < G ? af_lds32(qs + warp * 256 + g * 64 + (16 * ks + 2 * t) * 2) : 0u;
      qb[ks][1] = g < G ? af_lds32(qs + warp * 256 + g * 64 + (16 * ks + 8 + 2 * t) * 2) : 0u;
    }
    if (lane < 8) kvnew = lds_v4(kvs + warp * 128 + lane * 16);
    __syncthreads();                                       // scratch rows are dead: stage 3 may be refilled
    fetch_x(round + 1);
    fetch_angles(round + 1);

    if (active) {
      float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;       // heads 2t, 2t+1 (log2 domain)
      float oacc[2][4];
#pragma unroll
      for (int md = 0; md < 2; ++md) oacc[md][0] = oacc[md][1] = oacc[md][2] = oacc[md][3] = 0.f;
      for (int c = 0; c < nchunks; ++c) {
        af_wait<AF_STAGES - 2>();
        __syncwarp();
        const uint32_t st = ring + (c % AF_STAGES) * AF_STAGE_BYTES;
        if (c == c_new) {                                  // this step's k,v replace whatever the stream brought for the slot
          if (lane < 8) sts_v4(st + (lane < 4 ? 0 : AF_HALF) + af_off64(r_new, lane & 3), kvnew);
          __syncwarp();
        }
        float sacc[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          sacc[mt][0] = sacc[mt][1] = sacc[mt][2] = sacc[mt][3] = 0.f;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            uint32_t ka[4];
            af_ldsm_x4(ka, st + af_off64(16 * mt + (mi & 1) * 8 + r8, ks * 2 + (mi >> 1)));
            af_mma(sacc[mt], ka, qb[ks][0], qb[ks][1]);
          }
        }
        float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int pa = c * AF_CHUNK + 16 * mt + g;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int pp = pa + ((e & 2) ? 8 : 0);
            float v = sacc[mt][e] * p.scale_log2;
            if (pp >= kv_len) v = -INFINITY;
            sacc[mt][e] = v;
          }
          cm0 = fmaxf(cm0, fmaxf(sacc[mt][0], sacc[mt][2]));
          cm1 = fmaxf(cm1, fmaxf(sacc[mt][1], sacc[mt][3]));
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, o));
          cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, o));
        }
        const float mn0 = fmaxf(m0, cm0), mn1 = fmaxf(m1, cm1);
        const float a0 = exp2f(m0 - mn0), a1 = exp2f(m1 - mn1);     // first chunk: exp2(-inf) = 0
        m0 = mn0; m1 = mn1;
        l0 *= a0; l1 *= a1;
#pragma unroll
        for (int md = 0; md < 2; ++md) {
          oacc[md][0] *= a0; oacc[md][2] *= a0;
          oacc[md][1] *= a1; oacc[md][3] *= a1;
        }
        uint32_t pb[2][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const float p0 = head_lane ? exp2f(sacc[mt][0] - m0) : 0.f, p1 = head1 ? exp2f(sacc[mt][1] - m1) : 0.f;
          const float p2 = head_lane ? exp2f(sacc[mt][2] - m0) : 0.f, p3 = head1 ? exp2f(sacc[mt][3] - m1) : 0.f;
          const uint32_t lo = pack_bf16(p0, p1), hi = pack_bf16(p2, p3);
          l0 += bf16_lo(lo) + bf16_lo(hi);
          l1 += bf16_hi(lo) + bf16_hi(hi);
          pb[mt][0] = af_movmatrix_t(lo);
          pb[mt][1] = af_movmatrix_t(hi);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
          for (int md = 0; md < 2; ++md) {
            uint32_t va[4];
            af_ldsm_x4_t(va, st + AF_HALF + af_off64(16 * j + (mi >> 1) * 8 + r8, 2 * md + (mi & 1)));
            af_mma(oacc[md], va, pb[j][0], pb[j][1]);
          }
        }
        __syncwarp();
        issue(c + AF_STAGES - 1);
      }
      af_wait<0>();
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        l0 += __shfl_xor_sync(0xffffffffu, l0, o);
        l1 += __shfl_xor_sync(0xffffffffu, l1, o);
      }
      if (head_lane) {
        const float i0 = 1.0f / l0, i1 = head1 ? 1.0f / l1 : 0.f;
        __nv_bfloat16* orow = p.out + (size_t)b * 128;
#pragma unroll
        for (int md = 0; md < 2; ++md) {
          const int d0 = 16 * md + g;
          orow[(2 * t) * 32 + d0] = __float2bfloat16_rn(oacc[md][0] * i0);
          orow[(2 * t) * 32 + d0 + 8] = __float2bfloat16_rn(oacc[md][2] * i0);
          if (head1) {
            orow[(2 * t + 1) * 32 + d0] = __float2bfloat16_rn(oacc[md][1] * i1);
            orow[(2 * t + 1) * 32 + d0 + 8] = __float2bfloat16_rn(oacc[md][3] * i1);
          }
        }
      }
    } else {
      af_wait<0>();
    }
    // hidden 256: the weight fragments do not stay in registers across the cache loop - re-read them (L2) for the next
    // round; unconditional, so that nothing of the previous copy is live inside the loop above
    if (HK > 8) load_w();
  }
}

}  // namespace
}  // namespace mpx

using namespace mpx;

extern "C" int mpx_attn_decode_fused(const mpx_attn_decode_fused_args* a, mpx_stream_t stream) {
  MPX_REQUIRE(a && a->x && a->norm_scale && a->wt_qkv && a->pos && a->rope_timescale && a->kcache && a->vcache && a->out,
              "mpx_attn_decode_fused: null pointer");
  MPX_REQUIRE(a->B > 0 && a->S > 0, "mpx_attn_decode_fused: bad shape");
  MPX_REQUIRE((a->H == 128 || a->H == 256) && a->Nq == 4 && a->Nkv == 1 && a->D == 32,
              "mpx_attn_decode_fused: only hidden 128 / 256, 4 query heads on 1 kv head, head_dim 32 (got H=%d Nq=%d Nkv=%d D=%d); "
              "use mpx_qkv_rope + mpx_attn_decode", a->H, a->Nq, a->Nkv, a->D);
  MPX_REQUIRE((reinterpret_cast<uintptr_t>(a->x) & (a->H == 256 ? 15 : 7)) == 0 && (reinterpret_cast<uintptr_t>(a->norm_scale) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(a->kcache) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->vcache) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(a->wt_qkv) & 15) == 0, "mpx_attn_decode_fused: misaligned pointer");
  const int sms = num_sms();
  int best_w = 14;
  double best_eff = -1.0;
  for (int w = 14; w >= 12; --w) {
    const int ctas = ceil_div(a->B, w) < sms ? ceil_div(a->B, w) : sms;
    const int rounds = ceil_div(a->B, ctas * w);
    const double eff = (double)a->B / ((double)sms * rounds * w);
    if (eff > best_eff + 1e-9) { best_eff = eff; best_w = w; }
  }
  const size_t smem = (size_t)AF_RING * best_w;
  const int wide = a->H == 256 ? 1 : 0;
  static size_t attr_smem[2] = {0, 0};
  if (smem > attr_smem[wide]) {
    cudaError_t e = wide ? cudaFuncSetAttribute(attn_decode_fused_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                         : cudaFuncSetAttribute(attn_decode_fused_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attn_decode_fused): %s", cudaGetErrorString(e));
      return MPX_ERR_CUDA;
    }
    attr_smem[wide] = smem;
  }
  const int grid = ceil_div(a->B, best_w) < sms ? ceil_div(a->B, best_w) : sms;
  AttnFusedParams p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(a->x);
  p.norm_scale = a->norm_scale;
  p.eps = a->eps;
  p.wt_qkv = reinterpret_cast<const __nv_bfloat16*>(a->wt_qkv);
  p.pos = a->pos;
  p.timescale = a->rope_timescale;
  p.kcache = reinterpret_cast<__nv_bfloat16*>(a->kcache);
  p.vcache = reinterpret_cast<__nv_bfloat16*>(a->vcache);
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.B = a->B; p.S = a->S;
  p.scale_log2 = (1.0f / sqrtf((float)a->D)) * 1.4426950408889634f;
  p.early_stream = (a->flags & 1) ? 1 : 0;
  cudaError_t le = wide ? launch_ex(attn_decode_fused_kernel<16>, dim3(grid), dim3(best_w * 32), smem, (cudaStream_t)stream, 0, 4, p)
                        : launch_ex(attn_decode_fused_kernel<8>, dim3(grid), dim3(best_w * 32), smem, (cudaStream_t)stream, 0, 4, p);
  if (le != cudaSuccess) {
    set_error("launch attn_decode_fused_kernel: %s", cudaGetErrorString(le));
    return MPX_ERR_CUDA;
  }
  return check_launch("attn_decode_fused_kernel");
}
