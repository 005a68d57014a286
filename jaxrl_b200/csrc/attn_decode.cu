// Rollout decode attention: one query token per agent against that agent's KV ring cache.
// Restates the decode branch of AttentionBlock.__call__ (reference mapox_trainer/model/attention.py:136-150)
// with jax.nn.dot_product_attention's "xla" arithmetic (contract #4):
//   kv_len = min(pos[0]+1, S) for every row (one scalar for the whole batch, attention.py:141-143);
//   logits = (q . k) * 1/sqrt(D) in fp32; softmax over [0,kv_len) in fp32 (exp(x-max)/sum); probs -> bf16;
//   out = probs . v (fp32 accumulate) -> bf16;  query head n reads kv head n / (Nq/Nkv).
// HBM-bound: (kv_len * Nkv * D * 2 * 2) bytes of cache per agent per layer, read exactly once: only slots
// [0,kv_len) are touched (the XLA reference path scans all S slots).
//
// One warp per (agent, kv head).  K then V stream through a 4-stage shared-memory ring in 32-position chunks
// (16-byte cp.async, fully coalesced for Nkv == 1; rows are XOR-swizzled so both access patterns below are
// bank-conflict free).  Pass 1: lane = position, fp32 dot products for the G query heads, scores kept in smem;
// warp-shuffle max / sum.  Pass 2: lane = (position parity, d pair), probabilities broadcast from smem.
#include "common.cuh"
#include "ptx.cuh"

namespace mpx {

constexpr int AD_WARPS = 4;
constexpr int AD_STAGES = 4;
constexpr int AD_CHUNK = 32;                  // positions per chunk
constexpr int AD_STAGE_BYTES = AD_CHUNK * 64; // head_dim 32 bf16 = 64 B per position

__device__ __forceinline__ void cp_async_16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t n = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(n)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int G>
__device__ __forceinline__ void st_heads(float* dst, const float (&s)[G]) {
  if constexpr (G == 4) *reinterpret_cast<float4*>(dst) = make_float4(s[0], s[1], s[2], s[3]);
  else if constexpr (G == 2) *reinterpret_cast<float2*>(dst) = make_float2(s[0], s[1]);
  else {
#pragma unroll
    for (int h = 0; h < G; ++h) dst[h] = s[h];
  }
}
template <int G>
__device__ __forceinline__ void ld_heads(const float* src, float (&s)[G]) {
  if constexpr (G == 4) {
    const float4 v = *reinterpret_cast<const float4*>(src);
    s[0] = v.x; s[1] = v.y; s[2] = v.z; s[3] = v.w;
  } else if constexpr (G == 2) {
    const float2 v = *reinterpret_cast<const float2*>(src);
    s[0] = v.x; s[1] = v.y;
  } else {
#pragma unroll
    for (int h = 0; h < G; ++h) s[h] = src[h];
  }
}

template <int G>
__global__ void __launch_bounds__(AD_WARPS * 32) attn_decode_kernel(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kcache,
    const __nv_bfloat16* __restrict__ vcache, const int32_t* __restrict__ pos, __nv_bfloat16* __restrict__ out,
    int B, int S, int Nkv, float scale) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = G * 32 * 4 + S * G * 4 + AD_STAGES * AD_STAGE_BYTES;
  uint8_t* base = smem_raw + warp * per_warp;
  float* q32 = reinterpret_cast<float*>(base);                       // [G][32]
  float* sc = reinterpret_cast<float*>(base + G * 32 * 4);           // [S][G]
  uint8_t* ring = base + G * 32 * 4 + S * G * 4;                     // [AD_STAGES][32 rows][64 B]

  const int wg = blockIdx.x * AD_WARPS + warp;
  if (wg >= B * Nkv) return;
  const int b = wg / Nkv, kvh = wg % Nkv;
  const int p0 = pos[0];
  const int kv_len = min(p0 + 1, S);
  const int nchunks = (kv_len + AD_CHUNK - 1) / AD_CHUNK;
  const int Nq = Nkv * G;

  // q -> fp32 in smem
  {
    const __nv_bfloat16* qrow = q + (size_t)b * Nq * 32 + kvh * G * 32;
    for (int i = lane; i < G * 32; i += 32) q32[i] = __bfloat162float(qrow[i]);
  }

  const size_t row_stride = (size_t)Nkv * 32;   // elements between consecutive positions of this kv head
  const __nv_bfloat16* kbase = kcache + ((size_t)b * S * Nkv + kvh) * 32;
  const __nv_bfloat16* vbase = vcache + ((size_t)b * S * Nkv + kvh) * 32;

  auto issue = [&](int ci) {
    if (ci < 2 * nchunks) {
      const bool is_v = ci >= nchunks;
      const int c = is_v ? ci - nchunks : ci;
      const __nv_bfloat16* src = is_v ? vbase : kbase;
      uint8_t* st = ring + (ci % AD_STAGES) * AD_STAGE_BYTES;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int pi = lane + 32 * i;
        const int r = pi >> 2, j = pi & 3;
        const int p = c * AD_CHUNK + r;
        const bool valid = p < kv_len;
        const __nv_bfloat16* g = src + (size_t)(valid ? p : 0) * row_stride + j * 8;
        cp_async_16_zfill(st + r * 64 + ((j ^ ((r >> 1) & 3)) * 16), g, valid);
      }
    }
    cp_async_commit();
  };

#pragma unroll
  for (int i = 0; i < AD_STAGES - 1; ++i) issue(i);
  __syncwarp();

  float mx[G];
#pragma unroll
  for (int h = 0; h < G; ++h) mx[h] = -INFINITY;

  int ci = 0;
  // ---------------- pass 1: scores ----------------
  for (int c = 0; c < nchunks; ++c, ++ci) {
    cp_async_wait<AD_STAGES - 2>();
    __syncwarp();
    const uint8_t* st = ring + (ci % AD_STAGES) * AD_STAGE_BYTES;
    const int r = lane;
    float s[G];
#pragma unroll
    for (int h = 0; h < G; ++h) s[h] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 kv = *reinterpret_cast<const uint4*>(st + r * 64 + ((j ^ ((r >> 1) & 3)) * 16));
      float kf[8];
      kf[0] = bf16_lo(kv.x); kf[1] = bf16_hi(kv.x); kf[2] = bf16_lo(kv.y); kf[3] = bf16_hi(kv.y);
      kf[4] = bf16_lo(kv.z); kf[5] = bf16_hi(kv.z); kf[6] = bf16_lo(kv.w); kf[7] = bf16_hi(kv.w);
#pragma unroll
      for (int h = 0; h < G; ++h) {
        const float4 qa = *reinterpret_cast<const float4*>(q32 + h * 32 + j * 8);
        const float4 qb = *reinterpret_cast<const float4*>(q32 + h * 32 + j * 8 + 4);
        s[h] = fmaf(qa.x, kf[0], s[h]); s[h] = fmaf(qa.y, kf[1], s[h]);
        s[h] = fmaf(qa.z, kf[2], s[h]); s[h] = fmaf(qa.w, kf[3], s[h]);
        s[h] = fmaf(qb.x, kf[4], s[h]); s[h] = fmaf(qb.y, kf[5], s[h]);
        s[h] = fmaf(qb.z, kf[6], s[h]); s[h] = fmaf(qb.w, kf[7], s[h]);
      }
    }
    const int p = c * AD_CHUNK + r;
    if (p < kv_len) {
#pragma unroll
      for (int h = 0; h < G; ++h) {
        s[h] *= scale;
        mx[h] = fmaxf(mx[h], s[h]);
      }
      st_heads<G>(sc + p * G, s);
    }
    __syncwarp();
    issue(ci + AD_STAGES - 1);
  }
  // ---------------- softmax statistics ----------------
  float l[G];
#pragma unroll
  for (int h = 0; h < G; ++h) {
    mx[h] = warp_max(mx[h]);
    l[h] = 0.f;
  }
  __syncwarp();
  for (int p = lane; p < kv_len; p += 32) {
    float e[G];
    ld_heads<G>(sc + p * G, e);
#pragma unroll
    for (int h = 0; h < G; ++h) {
      e[h] = expf(e[h] - mx[h]);
      l[h] += e[h];
    }
    st_heads<G>(sc + p * G, e);
  }
#pragma unroll
  for (int h = 0; h < G; ++h) l[h] = warp_sum(l[h]);
  for (int p = lane; p < kv_len; p += 32) {
    float e[G];
    ld_heads<G>(sc + p * G, e);
#pragma unroll
    for (int h = 0; h < G; ++h) e[h] = bf16_round(e[h] / l[h]);   // probs cast to bf16 before PV
    st_heads<G>(sc + p * G, e);
  }
  __syncwarp();
  // ---------------- pass 2: probs . V ----------------
  const int half = lane >> 4, dp = lane & 15;
  float acc[G][2];
#pragma unroll
  for (int h = 0; h < G; ++h) acc[h][0] = acc[h][1] = 0.f;
  for (int c = 0; c < nchunks; ++c, ++ci) {
    cp_async_wait<AD_STAGES - 2>();
    __syncwarp();
    const uint8_t* st = ring + (ci % AD_STAGES) * AD_STAGE_BYTES;
    const int rows = min(AD_CHUNK, kv_len - c * AD_CHUNK);
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
      const int r = 2 * i + half;
      if (r < rows) {
        const int j = dp >> 2;
        const uint32_t v2 =
            *reinterpret_cast<const uint32_t*>(st + r * 64 + ((j ^ ((r >> 1) & 3)) * 16) + (dp & 3) * 4);
        const float v0 = bf16_lo(v2), v1 = bf16_hi(v2);
        float pr[G];
        ld_heads<G>(sc + (c * AD_CHUNK + r) * G, pr);
#pragma unroll
        for (int h = 0; h < G; ++h) {
          acc[h][0] = fmaf(pr[h], v0, acc[h][0]);
          acc[h][1] = fmaf(pr[h], v1, acc[h][1]);
        }
      }
    }
    __syncwarp();
    issue(ci + AD_STAGES - 1);
  }
  cp_async_wait<0>();
#pragma unroll
  for (int h = 0; h < G; ++h) {
    acc[h][0] += __shfl_xor_sync(0xffffffffu, acc[h][0], 16);
    acc[h][1] += __shfl_xor_sync(0xffffffffu, acc[h][1], 16);
  }
  if (half == 0) {
    __nv_bfloat16* orow = out + (size_t)b * Nq * 32 + kvh * G * 32;
#pragma unroll
    for (int h = 0; h < G; ++h)
      *reinterpret_cast<uint32_t*>(orow + h * 32 + 2 * dp) = pack_bf16(acc[h][0], acc[h][1]);
  }
}

// ================================================================================================ v2: mma.sync
// Same arithmetic, tensor-core inner products.  Per 32-position chunk:
//   pass 1  S^T[pos x heads] = K[pos x d] . Q^T[d x heads]   (A = K rows via ldmatrix, B = Q in registers)
//   pass 2  O^T[d x heads]  += V^T[d x pos] . P^T[pos x heads] (A = V via ldmatrix.trans, B = bf16 probs from smem)
// The G (<= 8) query heads of the group occupy the N = 8 dimension of m16n8k16; scores stay fp32 in shared memory
// ([head][pos], row pitch S+4 floats: conflict-free), the normalised probabilities are rounded to bf16 IN PLACE
// over the front half of each score row.  Warps are persistent and stride over (agent, kv head) pairs.
constexpr int AD2_STAGES = 3;

__device__ __forceinline__ uint32_t ad_off64(int row, int piece) { return row * 64 + ((piece ^ ((row >> 1) & 3)) << 4); }
__device__ __forceinline__ void ad_ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ad_ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ad_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int G>
__global__ void __launch_bounds__(512) attn_decode_mma_kernel(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kcache,
    const __nv_bfloat16* __restrict__ vcache, const int32_t* __restrict__ pos, __nv_bfloat16* __restrict__ out,
    int B, int S, int Nkv, float scale) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Sp = S + 4;                                            // score row pitch (floats)
  const int per_warp = ((G * Sp * 4 + 127) & ~127) + AD2_STAGES * AD_STAGE_BYTES;
  uint8_t* base = smem_raw + warp * per_warp;
  float* sc = reinterpret_cast<float*>(base);                      // [G][Sp] fp32 scores -> exp -> (front half) bf16 probs
  const uint32_t ring = smem_u32(base + ((G * Sp * 4 + 127) & ~127));
  const int g = lane >> 2, t = lane & 3, mi = lane >> 3, r8 = lane & 7;
  const int p0 = pos[0];
  const int kv_len = min(p0 + 1, S);
  const int nchunks = (kv_len + AD_CHUNK - 1) / AD_CHUNK;
  const int Nq = Nkv * G;
  const size_t row_stride = (size_t)Nkv * 32;
  const bool head_lane = (2 * t) < G;                              // this lane's C columns (heads 2t, 2t+1) are real
  const int total = B * Nkv;

  for (int wg = blockIdx.x * warps + warp; wg < total; wg += gridDim.x * warps) {
    const int b = wg / Nkv, kvh = wg % Nkv;
    const __nv_bfloat16* kbase = kcache + ((size_t)b * S * Nkv + kvh) * 32;
    const __nv_bfloat16* vbase = vcache + ((size_t)b * S * Nkv + kvh) * 32;
    // B fragments of Q^T (k = d, n = head): b0 = Q[g][16ks + 2t..], b1 = Q[g][16ks + 8 + 2t..]
    uint32_t qb[2][2];
    {
      const __nv_bfloat16* qrow = q + (size_t)b * Nq * 32 + (size_t)kvh * G * 32 + g * 32;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        qb[ks][0] = g < G ? *reinterpret_cast<const uint32_t*>(qrow + 16 * ks + 2 * t) : 0u;
        qb[ks][1] = g < G ? *reinterpret_cast<const uint32_t*>(qrow + 16 * ks + 8 + 2 * t) : 0u;
      }
    }
    auto issue = [&](int ci) {
      if (ci < 2 * nchunks) {
        const bool is_v = ci >= nchunks;
        const int c = is_v ? ci - nchunks : ci;
        const __nv_bfloat16* src = is_v ? vbase : kbase;
        const uint32_t st = ring + (ci % AD2_STAGES) * AD_STAGE_BYTES;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int pi = lane + 32 * i;
          const int r = pi >> 2, j = pi & 3;
          const int p = c * AD_CHUNK + r;
          const bool valid = p < kv_len;
          const __nv_bfloat16* gp = src + (size_t)(valid ? p : 0) * row_stride + j * 8;
          const uint32_t n = valid ? 16u : 0u;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(st + ad_off64(r, j)), "l"(gp), "r"(n) : "memory");
        }
      }
      cp_async_commit();
    };
    __syncwarp();                                                  // previous agent's smem reads are complete
#pragma unroll
    for (int i = 0; i < AD2_STAGES - 1; ++i) issue(i);

    float mx0 = -INFINITY, mx1 = -INFINITY;                        // running max of heads 2t, 2t+1 (this lane's rows)
    int ci = 0;
    // ---------------- pass 1: scores ----------------
    for (int c = 0; c < nchunks; ++c, ++ci) {
      cp_async_wait<AD2_STAGES - 2>();
      __syncwarp();
      const uint32_t st = ring + (ci % AD2_STAGES) * AD_STAGE_BYTES;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        float cacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t ka[4];
          ad_ldsm_x4(ka, st + ad_off64(16 * mt + (mi & 1) * 8 + r8, ks * 2 + (mi >> 1)));
          ad_mma(cacc, ka, qb[ks][0], qb[ks][1]);
        }
        if (head_lane) {
          const int pa = c * AD_CHUNK + 16 * mt + g, pb = pa + 8;
          if (pa < kv_len) {
            const float s0 = cacc[0] * scale, s1 = cacc[1] * scale;
            sc[(2 * t) * Sp + pa] = s0;
            mx0 = fmaxf(mx0, s0);
            if (2 * t + 1 < G) { sc[(2 * t + 1) * Sp + pa] = s1; mx1 = fmaxf(mx1, s1); }
          }
          if (pb < kv_len) {
            const float s2 = cacc[2] * scale, s3 = cacc[3] * scale;
            sc[(2 * t) * Sp + pb] = s2;
            mx0 = fmaxf(mx0, s2);
            if (2 * t + 1 < G) { sc[(2 * t + 1) * Sp + pb] = s3; mx1 = fmaxf(mx1, s3); }
          }
        }
      }
      __syncwarp();
      issue(ci + AD2_STAGES - 1);
    }
    // ---------------- softmax statistics ----------------
    // reduce the running maxima over the 8 position lanes (lane bits 2..4); lanes with equal t share heads 2t, 2t+1
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
    }
    float mxh[G], lh[G];
#pragma unroll
    for (int h = 0; h < G; ++h) {
      mxh[h] = __shfl_sync(0xffffffffu, (h & 1) ? mx1 : mx0, h >> 1);   // lane t = h/2 (g = 0) holds head h
      lh[h] = 0.f;
    }
    __syncwarp();
#pragma unroll
    for (int h = 0; h < G; ++h) {
      float* row = sc + h * Sp;
      for (int p = lane; p < kv_len; p += 32) {
        const float e = expf(row[p] - mxh[h]);
        row[p] = e;
        lh[h] += e;
      }
      lh[h] = warp_sum(lh[h]);
    }
    __syncwarp();
    // normalise, round to bf16 (probs are cast to bf16 before PV) and store in place over the front of the row:
    // iteration k reads positions [64k, 64k+64) and overwrites the bytes of positions [32k, 32k+32) only.
    const int kv_pad = nchunks * AD_CHUNK;
#pragma unroll
    for (int h = 0; h < G; ++h) {
      float* row = sc + h * Sp;
      uint32_t* prow = reinterpret_cast<uint32_t*>(row);
      for (int k = 0; 64 * k < kv_pad; ++k) {
        const int p2 = 64 * k + 2 * lane;
        const float e0 = p2 < kv_len ? row[p2] : 0.f;
        const float e1 = p2 + 1 < kv_len ? row[p2 + 1] : 0.f;
        __syncwarp();
        if (p2 < kv_pad) prow[32 * k + lane] = pack_bf16(e0 / lh[h], e1 / lh[h]);
        __syncwarp();
      }
    }
    __syncwarp();
    // ---------------- pass 2: O^T = V^T . P^T ----------------
    float oacc[2][4];
#pragma unroll
    for (int md = 0; md < 2; ++md) oacc[md][0] = oacc[md][1] = oacc[md][2] = oacc[md][3] = 0.f;
    const uint32_t* prow_g = reinterpret_cast<const uint32_t*>(sc + (g < G ? g : 0) * Sp);
    for (int c = 0; c < nchunks; ++c, ++ci) {
      cp_async_wait<AD2_STAGES - 2>();
      __syncwarp();
      const uint32_t st = ring + (ci % AD2_STAGES) * AD_STAGE_BYTES;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int pbase = c * AD_CHUNK + 16 * j;                    // bf16 pair index = position / 2
        const uint32_t b0 = g < G ? prow_g[(pbase >> 1) + t] : 0u;
        const uint32_t b1 = g < G ? prow_g[(pbase >> 1) + 4 + t] : 0u;
#pragma unroll
        for (int md = 0; md < 2; ++md) {
          uint32_t va[4];
          ad_ldsm_x4_t(va, st + ad_off64(16 * j + (mi >> 1) * 8 + r8, 2 * md + (mi & 1)));
          ad_mma(oacc[md], va, b0, b1);
        }
      }
      __syncwarp();
      issue(ci + AD2_STAGES - 1);
    }
    cp_async_wait<0>();
    if (head_lane) {
      __nv_bfloat16* orow = out + (size_t)b * Nq * 32 + (size_t)kvh * G * 32;
#pragma unroll
      for (int md = 0; md < 2; ++md) {
        const int d0 = 16 * md + g;
        orow[(2 * t) * 32 + d0] = __float2bfloat16_rn(oacc[md][0]);
        orow[(2 * t) * 32 + d0 + 8] = __float2bfloat16_rn(oacc[md][2]);
        if (2 * t + 1 < G) {
          orow[(2 * t + 1) * 32 + d0] = __float2bfloat16_rn(oacc[md][1]);
          orow[(2 * t + 1) * 32 + d0 + 8] = __float2bfloat16_rn(oacc[md][3]);
        }
      }
    }
  }
}

// ================================================================================================ v3: online softmax
// Single pass (flash-decoding): K and V chunks travel together (4 KB per stage, 4 stages -> 12 KB in flight per
// warp), no score buffer.  Per 32-position chunk: S^T = K.Q^T (mma), running max / rescale in the C-fragment lanes
// that own heads (2t, 2t+1), P -> bf16, C-layout -> B-layout with movmatrix.trans, O^T += V^T.P^T (mma).
// Rounds the UN-normalised probabilities to bf16 (like the training kernel and cuDNN's fused SDPA); the exact
// two-pass kernels above remain selectable (mpx_debug_set(1, 1|2)).
constexpr int AD3_STAGES = 4;
constexpr int AD3_STAGE_BYTES = 2 * AD_STAGE_BYTES;   // K chunk then V chunk

__device__ __forceinline__ uint32_t ad_movmatrix_t(uint32_t x) {
  uint32_t y;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

template <int G>
__global__ void __launch_bounds__(512) attn_decode_online_kernel(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kcache,
    const __nv_bfloat16* __restrict__ vcache, const int32_t* __restrict__ pos, __nv_bfloat16* __restrict__ out,
    int B, int S, int Nkv, float scale_log2) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ring = smem_u32(smem_raw) + warp * (AD3_STAGES * AD3_STAGE_BYTES);
  const int g = lane >> 2, t = lane & 3, mi = lane >> 3, r8 = lane & 7;
  griddep_launch();
  griddep_wait();                              // q, the KV ring and pos are written by the previous kernels
  const int kv_len = min(pos[0] + 1, S);
  const int nchunks = (kv_len + AD_CHUNK - 1) / AD_CHUNK;
  const int Nq = Nkv * G;
  const size_t row_stride = (size_t)Nkv * 32;
  const bool head_lane = (2 * t) < G;
  const bool head1 = (2 * t + 1) < G;
  const int total = B * Nkv;

  for (int wg = blockIdx.x * warps + warp; wg < total; wg += gridDim.x * warps) {
    const int b = wg / Nkv, kvh = wg % Nkv;
    const __nv_bfloat16* kbase = kcache + ((size_t)b * S * Nkv + kvh) * 32;
    const __nv_bfloat16* vbase = vcache + ((size_t)b * S * Nkv + kvh) * 32;
    uint32_t qb[2][2];
    {
      const __nv_bfloat16* qrow = q + (size_t)b * Nq * 32 + (size_t)kvh * G * 32 + g * 32;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        qb[ks][0] = g < G ? *reinterpret_cast<const uint32_t*>(qrow + 16 * ks + 2 * t) : 0u;
        qb[ks][1] = g < G ? *reinterpret_cast<const uint32_t*>(qrow + 16 * ks + 8 + 2 * t) : 0u;
      }
    }
    auto issue = [&](int c) {
      if (c < nchunks) {
        const uint32_t st = ring + (c % AD3_STAGES) * AD3_STAGE_BYTES;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int pi = lane + 32 * (i & 3);
          const int r = pi >> 2, j = pi & 3;
          const int p = c * AD_CHUNK + r;
          const bool valid = p < kv_len;
          const __nv_bfloat16* gp = (i < 4 ? kbase : vbase) + (size_t)(valid ? p : 0) * row_stride + j * 8;
          const uint32_t n = valid ? 16u : 0u;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;"
                       ::"r"(st + (i < 4 ? 0 : AD_STAGE_BYTES) + ad_off64(r, j)), "l"(gp), "r"(n) : "memory");
        }
      }
      cp_async_commit();
    };
    __syncwarp();
#pragma unroll
    for (int i = 0; i < AD3_STAGES - 1; ++i) issue(i);

    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;       // heads 2t, 2t+1 (log2 domain)
    float oacc[2][4];
#pragma unroll
    for (int md = 0; md < 2; ++md) oacc[md][0] = oacc[md][1] = oacc[md][2] = oacc[md][3] = 0.f;

    for (int c = 0; c < nchunks; ++c) {
      cp_async_wait<AD3_STAGES - 2>();
      __syncwarp();
      const uint32_t st = ring + (c % AD3_STAGES) * AD3_STAGE_BYTES;
      float sacc[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        sacc[mt][0] = sacc[mt][1] = sacc[mt][2] = sacc[mt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t ka[4];
          ad_ldsm_x4(ka, st + ad_off64(16 * mt + (mi & 1) * 8 + r8, ks * 2 + (mi >> 1)));
          ad_mma(sacc[mt], ka, qb[ks][0], qb[ks][1]);
        }
      }
      // scale, mask the ragged tail, chunk max over the 32 positions (lanes differing in g)
      float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int pa = c * AD_CHUNK + 16 * mt + g;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int pp = pa + ((e & 2) ? 8 : 0);
          float v = sacc[mt][e] * scale_log2;
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
      // P (bf16), row sums of the ROUNDED values, and the C-layout -> B-layout transposition
      uint32_t pb[2][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const float p0 = head_lane ? exp2f(sacc[mt][0] - m0) : 0.f, p1 = head1 ? exp2f(sacc[mt][1] - m1) : 0.f;
        const float p2 = head_lane ? exp2f(sacc[mt][2] - m0) : 0.f, p3 = head1 ? exp2f(sacc[mt][3] - m1) : 0.f;
        const uint32_t lo = pack_bf16(p0, p1), hi = pack_bf16(p2, p3);     // (pos g | g+8, heads 2t, 2t+1)
        l0 += bf16_lo(lo) + bf16_lo(hi);
        l1 += bf16_hi(lo) + bf16_hi(hi);
        pb[mt][0] = ad_movmatrix_t(lo);                                    // (k = pos 2t,2t+1 ; n = head g)
        pb[mt][1] = ad_movmatrix_t(hi);                                    // (k = pos 8+2t.. ; n = head g)
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
#pragma unroll
        for (int md = 0; md < 2; ++md) {
          uint32_t va[4];
          ad_ldsm_x4_t(va, st + AD_STAGE_BYTES + ad_off64(16 * j + (mi >> 1) * 8 + r8, 2 * md + (mi & 1)));
          ad_mma(oacc[md], va, pb[j][0], pb[j][1]);
        }
      }
      __syncwarp();
      issue(c + AD3_STAGES - 1);
    }
    cp_async_wait<0>();
    // total row sums: partial sums live in the 8 position lanes of each head pair
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      l0 += __shfl_xor_sync(0xffffffffu, l0, o);
      l1 += __shfl_xor_sync(0xffffffffu, l1, o);
    }
    if (head_lane) {
      const float i0 = 1.0f / l0, i1 = head1 ? 1.0f / l1 : 0.f;
      __nv_bfloat16* orow = out + (size_t)b * Nq * 32 + (size_t)kvh * G * 32;
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
  }
}

template <int G>
static int launch_attn_decode_online(const mpx_attn_decode_args* a, cudaStream_t stream) {
  const size_t per_warp = AD3_STAGES * AD3_STAGE_BYTES;
  const int wmax = 14;
  const int total = a->B * a->Nkv, sms = num_sms();
  int best_w = wmax;
  double best_eff = -1.0;
  for (int w = wmax; w >= 8; --w) {
    const int ctas = ceil_div(total, w) < sms ? ceil_div(total, w) : sms;
    const int rounds = ceil_div(total, ctas * w);
    const double eff = (double)total / ((double)sms * rounds * w);
    if (eff > best_eff + 1e-9) { best_eff = eff; best_w = w; }
  }
  const size_t smem = per_warp * best_w;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(attn_decode_online_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attn_decode_online): %s", cudaGetErrorString(e));
      return MPX_ERR_CUDA;
    }
    attr_smem = smem;
  }
  const int grid = ceil_div(total, best_w) < sms ? ceil_div(total, best_w) : sms;
  const float scale_log2 = (1.0f / sqrtf((float)a->D)) * 1.4426950408889634f;
  cudaError_t le = launch_ex(attn_decode_online_kernel<G>, dim3(grid), dim3(best_w * 32), smem, stream, 0, 4,
                             reinterpret_cast<const __nv_bfloat16*>(a->q), reinterpret_cast<const __nv_bfloat16*>(a->kcache),
                             reinterpret_cast<const __nv_bfloat16*>(a->vcache), a->pos,
                             reinterpret_cast<__nv_bfloat16*>(a->out), a->B, a->S, a->Nkv, scale_log2);
  if (le != cudaSuccess) {
    set_error("launch attn_decode_online_kernel: %s", cudaGetErrorString(le));
    return MPX_ERR_CUDA;
  }
  return check_launch("attn_decode_online_kernel");
}

// 0 = online-softmax mma (default), 1 = SIMT exact two-pass, 2 = mma exact two-pass
int g_attn_decode_impl = 0;

template <int G>
static int launch_attn_decode_mma(const mpx_attn_decode_args* a, cudaStream_t stream) {
  const size_t per_warp = (((size_t)G * (a->S + 4) * 4 + 127) & ~(size_t)127) + AD2_STAGES * AD_STAGE_BYTES;
  int wmax = (int)((227 * 1024) / per_warp);
  if (wmax > 16) wmax = 16;
  if (wmax < 1) {
    set_error("mpx_attn_decode: S=%d needs %zu bytes of shared memory per warp (> 227 KB)", a->S, per_warp);
    return MPX_ERR_UNSUPPORTED;
  }
  // warps per CTA: best balance of (agent, kv-head) pairs over one persistent CTA per SM
  const int total = a->B * a->Nkv, sms = num_sms();
  int best_w = wmax;
  double best_eff = -1.0;
  for (int w = wmax; w >= (wmax > 4 ? wmax - 6 : 1) && w >= 1; --w) {
    const int ctas = ceil_div(total, w) < sms ? ceil_div(total, w) : sms;
    const int rounds = ceil_div(total, ctas * w);
    const double eff = (double)total / ((double)sms * rounds * w);
    if (eff > best_eff + 1e-9) { best_eff = eff; best_w = w; }
  }
  const size_t smem = per_warp * best_w;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(attn_decode_mma_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attn_decode_mma): %s", cudaGetErrorString(e));
      return MPX_ERR_CUDA;
    }
    attr_smem = smem;
  }
  const int grid = ceil_div(total, best_w) < sms ? ceil_div(total, best_w) : sms;
  const float scale = 1.0f / sqrtf((float)a->D);
  attn_decode_mma_kernel<G><<<grid, best_w * 32, smem, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(a->q), reinterpret_cast<const __nv_bfloat16*>(a->kcache),
      reinterpret_cast<const __nv_bfloat16*>(a->vcache), a->pos, reinterpret_cast<__nv_bfloat16*>(a->out), a->B, a->S,
      a->Nkv, scale);
  return check_launch("attn_decode_mma_kernel");
}

template <int G>
static int launch_attn_decode(const mpx_attn_decode_args* a, cudaStream_t stream) {
  const size_t per_warp = (size_t)G * 32 * 4 + (size_t)a->S * G * 4 + AD_STAGES * AD_STAGE_BYTES;
  const size_t smem = per_warp * AD_WARPS;
  if (smem > 227 * 1024) {
    set_error("mpx_attn_decode: S=%d needs %zu bytes of shared memory per CTA (> 227 KB)", a->S, smem);
    return MPX_ERR_UNSUPPORTED;
  }
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(attn_decode_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attn_decode): %s", cudaGetErrorString(e));
      return MPX_ERR_CUDA;
    }
    attr_smem = smem;
  }
  const int warps = a->B * a->Nkv;
  const int grid = ceil_div(warps, AD_WARPS);
  const float scale = 1.0f / sqrtf((float)a->D);
  attn_decode_kernel<G><<<grid, AD_WARPS * 32, smem, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(a->q), reinterpret_cast<const __nv_bfloat16*>(a->kcache),
      reinterpret_cast<const __nv_bfloat16*>(a->vcache), a->pos, reinterpret_cast<__nv_bfloat16*>(a->out), a->B, a->S,
      a->Nkv, scale);
  return check_launch("attn_decode_kernel");
}

}  // namespace mpx

using namespace mpx;

extern "C" int mpx_attn_decode(const mpx_attn_decode_args* a, mpx_stream_t stream) {
  MPX_REQUIRE(a && a->q && a->kcache && a->vcache && a->pos && a->out, "mpx_attn_decode: null pointer");
  MPX_REQUIRE(a->B > 0 && a->S > 0 && a->Nq > 0 && a->Nkv > 0, "mpx_attn_decode: bad shape");
  MPX_REQUIRE(a->D == 32, "mpx_attn_decode: head_dim %d unsupported (only 32)", a->D);
  MPX_REQUIRE(a->Nq % a->Nkv == 0, "mpx_attn_decode: Nq=%d not a multiple of Nkv=%d", a->Nq, a->Nkv);
  cudaStream_t s = (cudaStream_t)stream;
  if (g_attn_decode_impl == 0) {
    switch (a->Nq / a->Nkv) {
      case 1: return launch_attn_decode_online<1>(a, s);
      case 2: return launch_attn_decode_online<2>(a, s);
      case 4: return launch_attn_decode_online<4>(a, s);
      case 8: return launch_attn_decode_online<8>(a, s);
      default:
        set_error("mpx_attn_decode: group size %d unsupported (1,2,4,8)", a->Nq / a->Nkv);
        return MPX_ERR_UNSUPPORTED;
    }
  }
  if (g_attn_decode_impl == 2) {
    switch (a->Nq / a->Nkv) {
      case 1: return launch_attn_decode_mma<1>(a, s);
      case 2: return launch_attn_decode_mma<2>(a, s);
      case 4: return launch_attn_decode_mma<4>(a, s);
      case 8: return launch_attn_decode_mma<8>(a, s);
      default:
        set_error("mpx_attn_decode: group size %d unsupported (1,2,4,8)", a->Nq / a->Nkv);
        return MPX_ERR_UNSUPPORTED;
    }
  }
  switch (a->Nq / a->Nkv) {
    case 1: return launch_attn_decode<1>(a, s);
    case 2: return launch_attn_decode<2>(a, s);
    case 4: return launch_attn_decode<4>(a, s);
    case 8: return launch_attn_decode<8>(a, s);
    default:
      set_error("mpx_attn_decode: group size %d unsupported (1,2,4,8)", a->Nq / a->Nkv);
      return MPX_ERR_UNSUPPORTED;
  }
}
