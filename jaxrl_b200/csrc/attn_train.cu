// PPO-update attention over full trajectory windows: causal GQA flash-style forward and backward.
// Restates the training branch of AttentionBlock.__call__ (reference mapox_trainer/model/attention.py:151-169,
// jax.nn.dot_product_attention(is_causal=True)) and its autodiff transpose (train.py:236-238 nnx.grad):
//   S = (q . k) / sqrt(D) (fp32), causal mask s <= t, P = softmax(S) (fp32), P -> bf16, O = P . v -> bf16.
// q/k arrive already rotated (RoPE fused in mpx_qkv_rope); the backward applies the inverse rotation to dq/dk
// so its outputs are cotangents of the *projection* outputs.
//
// v1 data path: bf16 m16n8k16 tensor-core MMAs with fp32 accumulation, K/V of one (batch, kv-head) staged in
// shared memory once per CTA and shared by the G = Nq/Nkv query heads of the group (one 64-row query tile per
// CTA, one warp per 16 rows per head).  head_dim = 32, T % 64 == 0.
#include "common.cuh"
#include "ptx.cuh"

namespace mpx {

constexpr int AT_TILE = 64;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp16(uint32_t smem_addr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// [rows][32 bf16] tile, 64-byte rows, 16-byte pieces XOR-swizzled so 8 consecutive rows hit 8 distinct bank groups.
__device__ __forceinline__ uint32_t off64(int row, int piece) { return row * 64 + ((piece ^ ((row >> 1) & 3)) << 4); }
// [rows][64 bf16] tile, 128-byte rows, standard 128-byte swizzle.
__device__ __forceinline__ uint32_t off128(int row, int piece) { return row * 128 + ((piece ^ (row & 7)) << 4); }

// cooperative async copy of `rows` rows of 32 bf16 (global row stride ld elements) into a swizzled 64 B tile
__device__ __forceinline__ void load_tile64(uint32_t smem_base, const __nv_bfloat16* g, long long ld, int rows, int tid,
                                            int nthreads) {
  for (int i = tid; i < rows * 4; i += nthreads) {
    const int r = i >> 2, p = i & 3;
    cp16(smem_base + off64(r, p), g + (long long)r * ld + p * 8);
  }
}

// ================================================================================================ forward
// grid (T/64, Nkv, B); block G*4 warps: warp w -> head (w / 4) of the group, query rows 16*(w % 4) of the tile.
template <int G>
__global__ void __launch_bounds__(G * 128) attn_fwd_kernel(
    const __nv_bfloat16* __restrict__ q, long long ldq, const __nv_bfloat16* __restrict__ k, long long ldk,
    const __nv_bfloat16* __restrict__ v, long long ldv, __nv_bfloat16* __restrict__ o, long long ldo,
    float* __restrict__ lse, int T, int Nkv, float scale_log2) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int nqt = T / AT_TILE;
  const int qt = nqt - 1 - blockIdx.x;   // heavy tiles first
  const int kvh = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkeys = (qt + 1) * AT_TILE;
  const uint32_t sK = smem_u32(smem);
  const uint32_t sV = sK + T * 64;
  const uint32_t sQ = sV + T * 64;       // [G][64 rows][64 B]
  const long long tok0 = (long long)b * T;
  const int Nq = Nkv * G;

  load_tile64(sK, k + tok0 * ldk + kvh * 32, ldk, nkeys, tid, G * 128);
  load_tile64(sV, v + tok0 * ldv + kvh * 32, ldv, nkeys, tid, G * 128);
  for (int h = 0; h < G; ++h)
    load_tile64(sQ + h * 4096, q + (tok0 + qt * AT_TILE) * ldq + (kvh * G + h) * 32, ldq, AT_TILE, tid, G * 128);
  cp_commit();
  cp_wait_all();
  __syncthreads();

  const int hw = warp >> 2;           // head within group
  const int rw = (warp & 3) * 16;     // row offset in the q tile
  const int g = lane >> 2, t = lane & 3;
  const int mi = lane >> 3, r8 = lane & 7;

  uint32_t qa[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
    ldsm_x4(qa[ks], sQ + hw * 4096 + off64(rw + (mi & 1) * 8 + r8, ks * 2 + (mi >> 1)));

  float oacc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) oacc[i][j] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int row0 = qt * AT_TILE + rw + g, row1 = row0 + 8;

  for (int kt = 0; kt <= qt; ++kt) {
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      uint32_t bf[4];
      ldsm_x4(bf, sK + off64(kt * AT_TILE + nt * 8 + r8, mi));
      mma16816(s[nt], qa[0], bf[0], bf[1]);
      mma16816(s[nt], qa[1], bf[2], bf[3]);
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int key = kt * AT_TILE + nt * 8 + 2 * t;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kk = key + (e & 1);
        const int rr = (e < 2) ? row0 : row1;
        float val = s[nt][e] * scale_log2;
        if (kt == qt && kk > rr) val = -INFINITY;
        s[nt][e] = val;
      }
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float a0 = exp2f(m0 - mn0), a1 = exp2f(m1 - mn1);
    m0 = mn0; m1 = mn1;
    l0 *= a0; l1 *= a1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      oacc[i][0] *= a0; oacc[i][1] *= a0; oacc[i][2] *= a1; oacc[i][3] *= a1;
    }
    uint32_t pa[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(s[nt][0] - m0), p1 = exp2f(s[nt][1] - m0);
      const float p2 = exp2f(s[nt][2] - m1), p3 = exp2f(s[nt][3] - m1);
      // the bf16-rounded probabilities are what multiplies V, so the normaliser sums the rounded values too
      const uint32_t lo = pack_bf16(p0, p1), hi = pack_bf16(p2, p3);
      l0 += bf16_lo(lo) + bf16_hi(lo);
      l1 += bf16_lo(hi) + bf16_hi(hi);
      pa[nt >> 1][(nt & 1) * 2 + 0] = lo;
      pa[nt >> 1][(nt & 1) * 2 + 1] = hi;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int dp = 0; dp < 2; ++dp) {
        uint32_t bf[4];
        ldsm_x4_t(bf, sV + off64(kt * AT_TILE + j * 16 + (mi & 1) * 8 + r8, dp * 2 + (mi >> 1)));
        mma16816(oacc[dp * 2], pa[j], bf[0], bf[1]);
        mma16816(oacc[dp * 2 + 1], pa[j], bf[2], bf[3]);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.f / l0, inv1 = 1.f / l1;
  const int head = kvh * G + hw;
  __nv_bfloat16* o0 = o + (tok0 + row0) * ldo + head * 32;
  __nv_bfloat16* o1 = o + (tok0 + row1) * ldo + head * 32;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    *reinterpret_cast<uint32_t*>(o0 + nt * 8 + 2 * t) = pack_bf16(oacc[nt][0] * inv0, oacc[nt][1] * inv0);
    *reinterpret_cast<uint32_t*>(o1 + nt * 8 + 2 * t) = pack_bf16(oacc[nt][2] * inv1, oacc[nt][3] * inv1);
  }
  if (t == 0) {
    lse[(tok0 + row0) * Nq + head] = m0 + log2f(l0);    // log2 units of the scaled scores
    lse[(tok0 + row1) * Nq + head] = m1 + log2f(l1);
  }
}

// ================================================================================================ backward
// delta[token, head] = sum_d dO * O  (fp32)
// 4 lanes per (token, head): 16-byte loads, two xor-shuffles.
__global__ void attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ dout, long long lddo,
                                      const __nv_bfloat16* __restrict__ o, long long ldo, float* __restrict__ delta,
                                      long long M, int Nq) {
  const int lane = threadIdx.x & 31;
  const int piece = lane & 3;
  const long long total = M * Nq;
  long long idx = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 8 + (lane >> 2);
  const long long stride = (long long)gridDim.x * (blockDim.x >> 5) * 8;
  for (long long base = idx - (lane >> 2); base < total; base += stride, idx += stride) {
    float a = 0.f;
    if (idx < total) {
      const long long tok = idx / Nq;
      const int h = (int)(idx % Nq);
      const uint4 dv = *reinterpret_cast<const uint4*>(dout + tok * lddo + h * 32 + piece * 8);
      const uint4 ov = *reinterpret_cast<const uint4*>(o + tok * ldo + h * 32 + piece * 8);
      a = bf16_lo(dv.x) * bf16_lo(ov.x) + bf16_hi(dv.x) * bf16_hi(ov.x) + bf16_lo(dv.y) * bf16_lo(ov.y) +
          bf16_hi(dv.y) * bf16_hi(ov.y) + bf16_lo(dv.z) * bf16_lo(ov.z) + bf16_hi(dv.z) * bf16_hi(ov.z) +
          bf16_lo(dv.w) * bf16_lo(ov.w) + bf16_hi(dv.w) * bf16_hi(ov.w);
    }
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    if (piece == 0 && idx < total) delta[idx] = a;
  }
}

// grid (T/64, Nkv, B); 4 warps; warp w owns keys 16w..16w+15 of key tile kt.  Loops over the G heads of the group
// and the query tiles qt >= kt.  dK/dV accumulate in registers (no atomics); dQ is accumulated in fp32 global
// memory with red.add and finished by attn_bwd_dq_kernel.
template <int G>
__global__ void __launch_bounds__(128) attn_bwd_kernel(
    const __nv_bfloat16* __restrict__ q, long long ldq, const __nv_bfloat16* __restrict__ k, long long ldk,
    const __nv_bfloat16* __restrict__ v, long long ldv, const __nv_bfloat16* __restrict__ dout, long long lddo,
    const float* __restrict__ lse, const float* __restrict__ delta, float* __restrict__ dq_acc,
    __nv_bfloat16* __restrict__ dk, long long lddk, __nv_bfloat16* __restrict__ dv, long long lddv,
    const int32_t* __restrict__ pos, int pos_len, const float* __restrict__ timescale, int T, int Nkv, float scale,
    float scale_log2) {
  __shared__ __align__(128) uint8_t smem[4096 * 4 + 8192 + 512];
  const int nqt = T / AT_TILE;
  const int kt = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3, mi = lane >> 3, r8 = lane & 7;
  const uint32_t sK = smem_u32(smem), sV = sK + 4096, sQ = sV + 4096, sDO = sQ + 4096, sDS = sDO + 4096;
  float* s_lse = reinterpret_cast<float*>(smem + 4096 * 4 + 8192);
  float* s_delta = s_lse + 64;
  const long long tok0 = (long long)b * T;
  const int Nq = Nkv * G;

  load_tile64(sK, k + (tok0 + kt * AT_TILE) * ldk + kvh * 32, ldk, AT_TILE, tid, 128);
  load_tile64(sV, v + (tok0 + kt * AT_TILE) * ldv + kvh * 32, ldv, AT_TILE, tid, 128);
  cp_commit();
  cp_wait_all();
  __syncthreads();

  const int kw = warp * 16;   // this warp's key rows within the tile
  uint32_t ka[2][4], va[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    ldsm_x4(ka[ks], sK + off64(kw + (mi & 1) * 8 + r8, ks * 2 + (mi >> 1)));
    ldsm_x4(va[ks], sV + off64(kw + (mi & 1) * 8 + r8, ks * 2 + (mi >> 1)));
  }
  float dkacc[4][4], dvacc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dkacc[i][j] = dvacc[i][j] = 0.f;
  const int key0 = kt * AT_TILE + kw + g, key1 = key0 + 8;

  for (int h = 0; h < G; ++h) {
    const int head = kvh * G + h;
    for (int qt = kt; qt < nqt; ++qt) {
      __syncthreads();   // previous iteration finished with sQ / sDO / sDS / s_lse
      const long long qtok = tok0 + qt * AT_TILE;
      load_tile64(sQ, q + qtok * ldq + head * 32, ldq, AT_TILE, tid, 128);
      load_tile64(sDO, dout + qtok * lddo + head * 32, lddo, AT_TILE, tid, 128);
      cp_commit();
      if (tid < 64) {
        s_lse[tid] = lse[(qtok + tid) * Nq + head];
        s_delta[tid] = delta[(qtok + tid) * Nq + head];
      }
      cp_wait_all();
      __syncthreads();

      // S^T = K_w . Q^T  and dP^T = V_w . dO^T   (16 keys x 64 queries each)
      float st[8][4], dpt[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        st[nt][0] = st[nt][1] = st[nt][2] = st[nt][3] = 0.f;
        dpt[nt][0] = dpt[nt][1] = dpt[nt][2] = dpt[nt][3] = 0.f;
        uint32_t bf[4];
        ldsm_x4(bf, sQ + off64(nt * 8 + r8, mi));
        mma16816(st[nt], ka[0], bf[0], bf[1]);
        mma16816(st[nt], ka[1], bf[2], bf[3]);
        ldsm_x4(bf, sDO + off64(nt * 8 + r8, mi));
        mma16816(dpt[nt], va[0], bf[0], bf[1]);
        mma16816(dpt[nt], va[1], bf[2], bf[3]);
      }
      // P^T = exp2(S^T * scale_log2 - lse[q]),  dS^T = P^T * (dP^T - delta[q]) * scale
      uint32_t pa[4][4], dsa[4][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        float p[4], ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qi = nt * 8 + 2 * t + (e & 1);
          const int key = (e < 2) ? key0 : key1;
          float pv = exp2f(st[nt][e] * scale_log2 - s_lse[qi]);
          if (qt == kt && key > qt * AT_TILE + qi) pv = 0.f;
          p[e] = pv;
          ds[e] = pv * (dpt[nt][e] - s_delta[qi]) * scale;
        }
        pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p[0], p[1]);
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p[2], p[3]);
        const uint32_t d01 = pack_bf16(ds[0], ds[1]), d23 = pack_bf16(ds[2], ds[3]);
        dsa[nt >> 1][(nt & 1) * 2 + 0] = d01;
        dsa[nt >> 1][(nt & 1) * 2 + 1] = d23;
        // dS^T tile to smem, [key][64 queries] with 128 B swizzle, for the dQ product
        const int qcol = nt * 8 + 2 * t;
        *reinterpret_cast<uint32_t*>(smem + 4096 * 4 + off128(kw + g, qcol >> 3) + (qcol & 7) * 2) = d01;
        *reinterpret_cast<uint32_t*>(smem + 4096 * 4 + off128(kw + g + 8, qcol >> 3) + (qcol & 7) * 2) = d23;
      }
      // dV_w += P^T . dO ; dK_w += dS^T . Q     (contraction over the 64 queries)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int dp = 0; dp < 2; ++dp) {
          uint32_t bf[4];
          ldsm_x4_t(bf, sDO + off64(j * 16 + (mi & 1) * 8 + r8, dp * 2 + (mi >> 1)));
          mma16816(dvacc[dp * 2], pa[j], bf[0], bf[1]);
          mma16816(dvacc[dp * 2 + 1], pa[j], bf[2], bf[3]);
          ldsm_x4_t(bf, sQ + off64(j * 16 + (mi & 1) * 8 + r8, dp * 2 + (mi >> 1)));
          mma16816(dkacc[dp * 2], dsa[j], bf[0], bf[1]);
          mma16816(dkacc[dp * 2 + 1], dsa[j], bf[2], bf[3]);
        }
      }
      __syncthreads();   // dS^T tile complete
      // dQ_w (queries 16w..16w+15) = dS . K   (contraction over the 64 keys of this tile)
      float dqa[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) dqa[i][0] = dqa[i][1] = dqa[i][2] = dqa[i][3] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t af[4];
        // A = dS (rows = queries, k = keys) read transposed from the [key][query] tile
        ldsm_x4_t(af, sDS + off128(j * 16 + (mi >> 1) * 8 + r8, ((warp * 16) >> 3) + (mi & 1)));
#pragma unroll
        for (int dp = 0; dp < 2; ++dp) {
          uint32_t bf[4];
          ldsm_x4_t(bf, sK + off64(j * 16 + (mi & 1) * 8 + r8, dp * 2 + (mi >> 1)));
          mma16816(dqa[dp * 2], af, bf[0], bf[1]);
          mma16816(dqa[dp * 2 + 1], af, bf[2], bf[3]);
        }
      }
      float* dq0 = dq_acc + (qtok + warp * 16 + g) * (long long)(Nq * 32) + head * 32;
      float* dq1 = dq0 + 8LL * (Nq * 32);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        atomicAdd(dq0 + nt * 8 + 2 * t, dqa[nt][0]);
        atomicAdd(dq0 + nt * 8 + 2 * t + 1, dqa[nt][1]);
        atomicAdd(dq1 + nt * 8 + 2 * t, dqa[nt][2]);
        atomicAdd(dq1 + nt * 8 + 2 * t + 1, dqa[nt][3]);
      }
    }
  }
  // epilogue: inverse RoPE on dK (pairs (d, d+16) live in n-tiles j and j+2 of the same thread), bf16 stores
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int key = half ? key1 : key0;
    const long long tok = tok0 + key;
    const float pf = (float)pos[tok % pos_len];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      float o_lo[2], o_hi[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = nt * 8 + 2 * t + e;   // rotation index 0..15
        float sn, cs;
        rope_sincos(pf / timescale[i], &sn, &cs);
        const float a = dkacc[nt][half * 2 + e], bb = dkacc[nt + 2][half * 2 + e];
        o_lo[e] = a * cs + bb * sn;
        o_hi[e] = bb * cs - a * sn;
      }
      __nv_bfloat16* dkr = dk + tok * lddk + kvh * 32;
      *reinterpret_cast<uint32_t*>(dkr + nt * 8 + 2 * t) = pack_bf16(o_lo[0], o_lo[1]);
      *reinterpret_cast<uint32_t*>(dkr + 16 + nt * 8 + 2 * t) = pack_bf16(o_hi[0], o_hi[1]);
    }
    __nv_bfloat16* dvr = dv + tok * lddv + kvh * 32;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
      *reinterpret_cast<uint32_t*>(dvr + nt * 8 + 2 * t) = pack_bf16(dvacc[nt][half * 2], dvacc[nt][half * 2 + 1]);
  }
}

// dq (bf16, inverse RoPE) from the fp32 accumulator: one thread per (token, two rotation pairs); the rotation angles
// depend on the token only, so one sincosf pair serves all Nq heads (the kernel was issue-bound on sincosf, not on HBM).
__global__ void attn_bwd_dq_kernel(const float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dq, long long lddq,
                                   const int32_t* __restrict__ pos, int pos_len, const float* __restrict__ timescale,
                                   long long M, int Nq) {
  const long long total = M * 8;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int sub = (int)(i & 7);
    const long long tok = i >> 3;
    const float pf = (float)pos[tok % pos_len];
    float s0, c0, s1, c1;
    rope_sincos(pf / timescale[sub * 2], &s0, &c0);
    rope_sincos(pf / timescale[sub * 2 + 1], &s1, &c1);
    const float* src = dq_acc + tok * (long long)(Nq * 32) + sub * 2;
    __nv_bfloat16* d = dq + tok * lddq + sub * 2;
    for (int head = 0; head < Nq; ++head, src += 32, d += 32) {
      const float2 a = *reinterpret_cast<const float2*>(src);
      const float2 bb = *reinterpret_cast<const float2*>(src + 16);
      *reinterpret_cast<uint32_t*>(d) = pack_bf16(a.x * c0 + bb.x * s0, a.y * c1 + bb.y * s1);
      *reinterpret_cast<uint32_t*>(d + 16) = pack_bf16(bb.x * c0 - a.x * s0, bb.y * c1 - a.y * s1);
    }
  }
}

template <int G>
static int launch_fwd(const mpx_attn_train_args* a, cudaStream_t s) {
  const size_t smem = (size_t)a->T * 128 + (size_t)G * 4096;
  if (smem > 227 * 1024) {
    set_error("mpx_attn_train_fwd: T=%d needs %zu B of shared memory", a->T, smem);
    return MPX_ERR_UNSUPPORTED;
  }
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attn_fwd): %s", cudaGetErrorString(e));
      return MPX_ERR_CUDA;
    }
    attr = smem;
  }
  const float scale_log2 = (1.0f / sqrtf((float)a->D)) * 1.4426950408889634f;
  dim3 grid(a->T / AT_TILE, a->Nkv, a->B);
  attn_fwd_kernel<G><<<grid, G * 128, smem, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(a->q), a->ldq, reinterpret_cast<const __nv_bfloat16*>(a->k), a->ldk,
      reinterpret_cast<const __nv_bfloat16*>(a->v), a->ldv, reinterpret_cast<__nv_bfloat16*>(a->o), a->ldo, a->lse,
      a->T, a->Nkv, scale_log2);
  return check_launch("attn_fwd_kernel");
}

int attn_bwd_tc_launch(const mpx_attn_train_bwd_args* a, cudaStream_t stream);   // attn_train_bwd_tc.cu

template <int G>
static int launch_bwd(const mpx_attn_train_bwd_args* a, cudaStream_t s) {
  const long long M = (long long)a->B * a->T;
  const float scale = 1.0f / sqrtf((float)a->D);
  cudaError_t e = cudaMemsetAsync(a->dq_acc, 0, (size_t)M * a->Nq * 32 * sizeof(float), s);
  if (e != cudaSuccess) {
    set_error("cudaMemsetAsync(dq_acc): %s", cudaGetErrorString(e));
    return MPX_ERR_CUDA;
  }
  long long nw = (M * a->Nq + 7) / 8;   // warps: 8 (token, head) pairs each
  int dgrid = (int)((nw + 7) / 8 < (long long)num_sms() * 16 ? (nw + 7) / 8 : (long long)num_sms() * 16);
  attn_bwd_delta_kernel<<<dgrid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(a->dout), a->lddo,
                                              reinterpret_cast<const __nv_bfloat16*>(a->o), a->ldo, a->delta, M, a->Nq);
  const int tc = attn_bwd_tc_launch(a, s);          // tcgen05 / TMEM kernel when the layout allows (attn_train_bwd_tc.cu)
  if (tc != MPX_OK && tc != MPX_ERR_UNSUPPORTED) return tc;
  dim3 grid(a->T / AT_TILE, a->Nkv, a->B);
  if (tc == MPX_ERR_UNSUPPORTED)
    attn_bwd_kernel<G><<<grid, 128, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(a->q), a->ldq, reinterpret_cast<const __nv_bfloat16*>(a->k), a->ldk,
      reinterpret_cast<const __nv_bfloat16*>(a->v), a->ldv, reinterpret_cast<const __nv_bfloat16*>(a->dout), a->lddo,
      a->lse, a->delta, a->dq_acc, reinterpret_cast<__nv_bfloat16*>(a->dk), a->lddk,
      reinterpret_cast<__nv_bfloat16*>(a->dv), a->lddv, a->pos, a->pos_len, a->rope_timescale, a->T, a->Nkv, scale,
      scale * 1.4426950408889634f);
  const long long tot = M * 8;
  int qgrid = (int)((tot + 255) / 256 < (long long)num_sms() * 16 ? (tot + 255) / 256 : (long long)num_sms() * 16);
  attn_bwd_dq_kernel<<<qgrid, 256, 0, s>>>(a->dq_acc, reinterpret_cast<__nv_bfloat16*>(a->dq), a->lddq, a->pos,
                                           a->pos_len, a->rope_timescale, M, a->Nq);
  return check_launch("attn_bwd_kernel");
}

}  // namespace mpx

using namespace mpx;

namespace mpx { int attn_fwd_tc_launch(const mpx_attn_train_args* a, cudaStream_t stream); }

extern "C" int mpx_attn_train_fwd(const mpx_attn_train_args* a, mpx_stream_t stream) {
  MPX_REQUIRE(a && a->q && a->k && a->v && a->o && a->lse, "mpx_attn_train_fwd: null pointer");
  MPX_REQUIRE(a->D == 32, "mpx_attn_train_fwd: head_dim %d unsupported (only 32)", a->D);
  MPX_REQUIRE(a->B > 0 && a->T > 0 && a->T % AT_TILE == 0, "mpx_attn_train_fwd: T=%d must be a positive multiple of 64", a->T);
  MPX_REQUIRE(a->Nkv > 0 && a->Nq % a->Nkv == 0, "mpx_attn_train_fwd: bad head counts");
  MPX_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 2 == 0, "mpx_attn_train_fwd: strides must be multiples of 8");
  cudaStream_t s = (cudaStream_t)stream;
  {  // tcgen05 / TMEM kernel when q, k, v are one packed (M, Nq*D + 2*D) buffer (Nkv == 1); else mma.sync
    const int rc = attn_fwd_tc_launch(a, s);
    if (rc != MPX_ERR_UNSUPPORTED) return rc;
  }
  switch (a->Nq / a->Nkv) {
    case 1: return launch_fwd<1>(a, s);
    case 2: return launch_fwd<2>(a, s);
    case 4: return launch_fwd<4>(a, s);
    default: set_error("mpx_attn_train_fwd: group size %d unsupported (1,2,4)", a->Nq / a->Nkv); return MPX_ERR_UNSUPPORTED;
  }
}

extern "C" int mpx_attn_train_bwd(const mpx_attn_train_bwd_args* a, mpx_stream_t stream) {
  MPX_REQUIRE(a && a->q && a->k && a->v && a->o && a->dout && a->lse && a->delta && a->dq_acc && a->dq && a->dk && a->dv &&
                  a->pos && a->rope_timescale, "mpx_attn_train_bwd: null pointer");
  MPX_REQUIRE(a->D == 32, "mpx_attn_train_bwd: head_dim %d unsupported (only 32)", a->D);
  MPX_REQUIRE(a->lddo % 8 == 0 && a->ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(a->dout) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(a->o) & 15) == 0,
              "mpx_attn_train_bwd: o / dout must be 16-byte aligned with pitches that are multiples of 8");
  MPX_REQUIRE(a->B > 0 && a->T > 0 && a->T % AT_TILE == 0, "mpx_attn_train_bwd: T=%d must be a positive multiple of 64", a->T);
  MPX_REQUIRE(a->Nkv > 0 && a->Nq % a->Nkv == 0 && a->pos_len > 0, "mpx_attn_train_bwd: bad head counts");
  cudaStream_t s = (cudaStream_t)stream;
  switch (a->Nq / a->Nkv) {
    case 1: return launch_bwd<1>(a, s);
    case 2: return launch_bwd<2>(a, s);
    case 4: return launch_bwd<4>(a, s);
    default: set_error("mpx_attn_train_bwd: group size %d unsupported (1,2,4)", a->Nq / a->Nkv); return MPX_ERR_UNSUPPORTED;
  }
}
