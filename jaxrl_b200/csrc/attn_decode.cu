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
